"""bwtb3m_b200: B200-native BWT construction by balanced block merging behind the gt1/bwtb3m
surface.  The compute path is libb3m.so (hand-written CUDA for sm_100a, C ABI in include/b3m.h);
this package is the thin host-side mirror used by tests, bench.py and the multi-GPU driver."""
from ._lib import INPUT_TYPES, LIB_PATH, lib  # noqa: F401
from .engine import B3MError, Engine, MultiEngine  # noqa: F401
