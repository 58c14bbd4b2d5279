"""gradcheck on device arrays: the reference's finite-difference instrument (test/gradcheck.jl:35-104)
restated against the C ABI.  Chosen arguments are wrapped in Param, `@diff gcsum(f, ...)` gives the
analytic gradient; up to `nsample` random entries per array are perturbed by +-delta (sign follows
the entry) and the one-sided difference is compared with isapprox(nd, ad; rtol, atol)."""
import numpy as np

from .core import Param, diff, grad, value
from .darray import DArray, isnum
from .rules import full, sum_


def _gcsum(f, *x, **kw):
    y = f(*x, **kw)
    v = value(y)
    if isnum(v) or (isinstance(v, DArray) and v.shape == ()):
        return y
    if v.size == 0:
        return 0
    return sum_(y)


def gradcheck(f, *x, kw=None, args=None, nsample=10, rtol=0.05, atol=0.01, delta=1e-4, rng=None):
    kw = kw or {}
    rng = rng or np.random.default_rng(0)
    args = list(range(len(x))) if args is None else list(args)
    xrec = list(x)
    for i in args:
        xrec[i] = Param(xrec[i])
    result = diff(lambda: _gcsum(f, *xrec, **kw))
    f0 = float(value(result))
    ok = True
    xptr = list(x)
    for i in args:
        g = full(grad(result, xrec[i]))
        xi = xptr[i]
        if isnum(xi):
            x1 = xi + delta if xi >= 0 else xi - delta
            xptr[i] = x1
            f1 = float(value(_gcsum(f, *xptr, **kw)))
            xptr[i] = xi
            nd = (f1 - f0) / (x1 - xi)
            ad = 0.0 if g is None else float(g)
            ok &= bool(np.isclose(nd, ad, rtol=rtol, atol=atol))
            continue
        host = xi.to_numpy()
        flat = host.reshape(-1, order="F")
        gflat = None if g is None else (np.full(flat.shape, float(g)) if isnum(g) else
                                         g.to_numpy().reshape(-1, order="F"))
        n = flat.size
        for j in (range(n) if n <= nsample else rng.integers(0, n, nsample)):
            old = flat[j]
            d = host.dtype.type(delta)
            pert = flat.copy()
            pert[j] = old + d if old >= 0 else old - d
            xptr[i] = DArray.from_numpy(pert.reshape(host.shape, order="F"))
            f1 = float(value(_gcsum(f, *xptr, **kw)))
            xptr[i] = xi
            nd = (f1 - f0) / (float(pert[j]) - float(old))
            ad = 0.0 if gflat is None else float(gflat[j])
            ok &= bool(np.isclose(nd, ad, rtol=rtol, atol=atol))
    return ok
