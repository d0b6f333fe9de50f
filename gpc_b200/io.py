"""SVM-light input, as CClctrl::readSvmlDataFile reads it (CClctrl.cpp:55-171): the native reader of the library
(gpc_svml_dims / gpc_svml_read); no parsing is done in Python."""
import ctypes as C

import numpy as np

from ._lib import check, i64, lib, ptr


def read_svml(path):
    """returns (X, y): X is N x D column-major (D = the largest feature index in the file), y is N x 1."""
    n, d = i64(0), C.c_int(0)
    check(lib().gpc_svml_dims(str(path).encode(), C.byref(n), C.byref(d)))
    X = np.zeros((n.value, d.value), order="F")
    y = np.zeros((n.value, 1), order="F")
    if n.value:
        check(lib().gpc_svml_read(str(path).encode(), ptr(X), max(n.value, 1), ptr(y), n.value, d.value))
    return X, y
