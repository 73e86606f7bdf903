"""Stand-in for fvcore: the reference imports four FLOP-counting names at module scope (vmamba.py:14) and only calls them
from its `flops()` helpers, which are outside the training / evaluation path."""
