"""pygps_b200 - the exact-GP hot path of pyGPs on a B200, behind pyGPs' own plugin API.

    import pygps_b200 as pyGPs
    model = pyGPs.GPR()
    model.setPrior(kernel=pyGPs.cov.RBF(np.log(2.), 0.))
    model.optimize(x, y)
    ym, ys2, fm, fs2, lp = model.predict(xs)

Drop-in for `pyGPs.GPR / GPR_FITC`, `pyGPs.cov.RBF / RBFard / Matern`,
`pyGPs.inf.Exact / FITC_Exact`, `getPosterior / optimize / predict`
(/root/reference/pyGPs/__init__.py:1-9 re-exports the same names).  The arithmetic runs in
libgpk.so (hand-written sm_100a CUDA, include/gpk.h); there is no CPU fallback.
"""
from . import cov, inf, lik, mean, opt, tools   # noqa: F401
from .gp import GP, GPR, GPC, GP_FITC, GPR_FITC  # noqa: F401
from . import gp                                # noqa: F401

__version__ = "0.1.0"
