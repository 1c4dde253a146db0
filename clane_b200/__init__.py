"""clane_b200 -- B200-native (sm_100a) implementation of CLANE's iterative embedding update.

Drop-in for the reference's Python surface on that path:
``clane_b200.graph.Graph``, ``clane_b200.similarity.*``, ``clane_b200.embedder.Embedder`` and
``python -m clane_b200`` mirror ``clane.graph`` / ``clane.similarity`` / ``clane.embedder`` /
``python -m clane`` of helloybz/CLANE; the arithmetic runs in hand-written CUDA kernels behind
the C-ABI of ``include/clane_b200.h`` (libclane_b200.so).
"""
__version__ = "0.1.0"
