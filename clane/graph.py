"""Alias of clane_b200.graph (drop-in for the reference module of the same name)."""
from clane_b200.graph import *  # noqa: F401,F403
from clane_b200 import graph as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
