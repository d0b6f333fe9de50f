"""gpc_b200 -- B200-native (sm_100a) exact-GP hot path behind GPc's CKern / CGp / CMatrix interface.

The compute lives in libgpc_b200.so (hand-written CUDA behind the C ABI in include/gpc_b200.h); this package is the
host-side mirror of the reference classes used by tests and bench.py.  There is no CPU fallback."""
import os as _os

# one evaluation uses ~20 CUDA streams; with the default 8 hardware work queues small chain kernels queue behind bulk
# launches of other streams (csrc/api.cu, gpc_more_hw_queues).  Read by the driver when the context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._lib import GpcError, MatrixNonPosDef, LIB_PATH, lib  # noqa: F401
from .kern import (CKern, CWhiteKern, CBiasKern, CRbfKern, CRbfardKern, CMatern32Kern, CMatern52Kern, CLinKern,  # noqa: F401
                   CPolyKern, CCmpndKern, DeviceContext, make_kern)
from .gp import CGp, CGplvm  # noqa: F401
from . import matrix  # noqa: F401
from .io import read_svml  # noqa: F401
