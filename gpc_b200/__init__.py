"""gpc_b200 -- B200-native (sm_100a) exact-GP hot path behind GPc's CKern / CGp / CMatrix interface.

The compute lives in libgpc_b200.so (hand-written CUDA behind the C ABI in include/gpc_b200.h); this package is the
host-side mirror of the reference classes used by tests and bench.py.  There is no CPU fallback."""
from ._lib import GpcError, MatrixNonPosDef, LIB_PATH, lib  # noqa: F401
from .kern import (CKern, CWhiteKern, CBiasKern, CRbfKern, CRbfardKern, CMatern32Kern, CMatern52Kern, CLinKern,  # noqa: F401
                   CPolyKern, CCmpndKern, DeviceContext, make_kern)
from .gp import CGp, CGplvm  # noqa: F401
from . import matrix  # noqa: F401
from .io import read_svml  # noqa: F401
