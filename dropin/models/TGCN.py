"""Drop-in for the reference's models/TGCN.py: re-exports the graphecho_b200 implementation."""
from graphecho_b200.models.TGCN import *  # noqa: F401,F403
from graphecho_b200.models import TGCN as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
