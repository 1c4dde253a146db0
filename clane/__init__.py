"""Alias package: lets code written against helloybz/CLANE (``from clane.graph import Graph``,
``python -m clane``) run on the B200 implementation unchanged.  Everything lives in clane_b200."""
