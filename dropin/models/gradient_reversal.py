"""Drop-in for the reference's models/gradient_reversal.py: re-exports the graphecho_b200 implementation."""
from graphecho_b200.models.gradient_reversal import *  # noqa: F401,F403
from graphecho_b200.models import gradient_reversal as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
