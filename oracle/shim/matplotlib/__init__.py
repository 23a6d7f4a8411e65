"""Empty stub: the reference imports matplotlib at module top (Core/gp.py:50,
Core/tools.py:25) but the hot path never plots."""
