"""Drop-in for the reference's models/graph_matching.py: re-exports the graphecho_b200 implementation."""
from graphecho_b200.models.graph_matching import *  # noqa: F401,F403
from graphecho_b200.models import graph_matching as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
