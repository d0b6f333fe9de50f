"""Host-side scaled-conjugate-gradient driver: the *caller* of the hot path (COptimisable::scgOptimise,
COptimisable.cpp:246-396, after Moller 1993).  Out of the GPU scope (O(P) vector work), but BASELINE config 5
("100 SCG iterations end-to-end") needs it, so it is restated here step for step -- including the reference's quirks
that shape the trajectory: step 3 adds lambdaDiff*|p| (not |p|^2) to delta (:313), and the convergence test looks at
CMatrix::max(), which because of a loop-body bug is max(p[0], p[-1]) (CMatrix.cpp:568-577)."""
import numpy as np


def scgOptimise(model, maxIters=1000, paramTol=1e-6, objectiveTol=1e-6, verbosity=0, log=None):
    """model: anything with getOptParams / setOptParams / computeObjectiveVal / computeObjectiveGradParams
    (the COptimisable interface, COptimisable.h:15-239).  Returns the number of iterations run."""
    w = np.array(model.getOptParams(), dtype=np.float64)
    nParams = w.size
    m_step, m_reg = 1.0e-4, 1.0
    lam, lamBar = m_reg, 0.0
    success = True
    g, oldObj = model.computeObjectiveGradParams()
    r = -np.asarray(g, dtype=np.float64)
    p = r.copy()
    s = np.zeros(nParams)
    delta = 0.0
    newObj = oldObj
    for it in range(1, maxIters + 1):
        normp = float(np.sqrt(p @ p))
        normp2 = normp * normp
        if success:  # 2
            sigma = m_step / normp
            model.setOptParams(w + sigma * p)
            gs, _ = model.computeObjectiveGradParams()
            s = (np.asarray(gs, dtype=np.float64) + r) / sigma
            delta = float(s @ p)
        lamDiff = lam - lamBar  # 3
        s = s + lamDiff * p
        delta += lamDiff * normp
        if delta <= 0.0:  # 4
            dn = delta / normp2
            s = s + (lam - 2.0 * dn) * p
            lamBar = 2.0 * (lam - dn)
            delta = lam * normp2 - delta
            lam = lamBar
        mu = float(p @ r)  # 5
        alpha = mu / delta
        wPlus = w + alpha * p  # 6
        model.setOptParams(wPlus)
        newObj = model.computeObjectiveVal()
        Delta = 2.0 * delta * (oldObj - newObj) / (mu * mu)
        if Delta >= 0.0:  # 7
            w = wPlus
            oldObj = newObj
            grp, _ = model.computeObjectiveGradParams()
            rp = -np.asarray(grp, dtype=np.float64)
            lamBar = 0.0
            success = True
            if it % nParams == 0:
                p = rp.copy()
            else:
                beta = (float(rp @ rp) - float(r @ rp)) / mu
                p = beta * p + rp
            r = rp
            if Delta >= 0.75:
                lam *= 0.5
            if lam < 1e-15:
                lam = 1e-15
        else:
            model.setOptParams(w)
            lamBar = lam
            success = False
        if Delta < 0.25:  # 8
            lam *= 4.0
        if log is not None:
            log.append(oldObj)
        if verbosity > 2:
            print("Iteration: %d Error: %.6g Scale: %.6g" % (it, oldObj, lam))
        pmax = max(p[0], p[-1])  # CMatrix::max() as implemented
        if success and abs(pmax * alpha) < paramTol and abs(newObj - oldObj) < objectiveTol:  # 9
            return it
    return maxIters


def scgOptimiseNative(model, maxIters=1000, paramTol=1e-6, objectiveTol=1e-6, log=None):
    """The same loop run natively (gpc_scg_minimise in libgpc_b200.so): Python is entered once per DISTINCT point, to
    evaluate objective and gradient together.  Returns (iterations, evaluations)."""
    import ctypes as C

    from ._lib import OBJECTIVE_FN, check, lib

    w = np.ascontiguousarray(model.getOptParams(), dtype=np.float64)
    n = w.size
    err = []

    def fn(user, wp, nn, obj, grad):
        try:
            model.setOptParams(np.ctypeslib.as_array(wp, shape=(nn,)).copy())
            g, f = model.computeObjectiveGradParams()
            obj[0] = float(f)
            np.ctypeslib.as_array(grad, shape=(nn,))[:] = np.asarray(g, dtype=np.float64)
            return 0
        except Exception as e:  # no exception may cross the C ABI
            err.append(e)
            return -1

    cb = OBJECTIVE_FN(fn)
    trace = np.zeros(max(int(maxIters), 1))
    it, ev = C.c_int(0), C.c_int(0)
    rc = lib().gpc_scg_minimise(cb, None, w.ctypes.data_as(C.c_void_p), n, int(maxIters), paramTol, objectiveTol,
                                trace.ctypes.data_as(C.c_void_p), C.byref(it), C.byref(ev))
    if err:
        raise err[0]
    check(rc)
    model.setOptParams(w)
    if log is not None:
        log.extend(trace[:it.value].tolist())
    return it.value, ev.value
