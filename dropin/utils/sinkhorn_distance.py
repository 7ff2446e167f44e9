"""Drop-in for the reference's utils/sinkhorn_distance.py: re-exports the graphecho_b200 implementation."""
from graphecho_b200.utils.sinkhorn_distance import *  # noqa: F401,F403
from graphecho_b200.utils import sinkhorn_distance as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
