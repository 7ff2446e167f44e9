"""Drop-in for the reference's models/affinity_layer.py: re-exports the graphecho_b200 implementation."""
from graphecho_b200.models.affinity_layer import *  # noqa: F401,F403
from graphecho_b200.models import affinity_layer as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
