"""Alias of clane_b200.similarity (drop-in for the reference module of the same name)."""
from clane_b200.similarity import *  # noqa: F401,F403
from clane_b200 import similarity as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
