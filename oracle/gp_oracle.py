"""CPU oracle for the exact-GP hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the one path of marionmari/pyGPs
that this repository accelerates (kernel-matrix build -> Cholesky of K/sn2+I ->
triangular solves -> nlZ / alpha / dnlZ / predictive mean+variance).  It exists
so the CUDA path can be checked on a box where `/root/reference` is absent.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it; nothing under `pygps_b200/` does, and the product never falls back to
it.

Parity status: PINNED.  `oracle/gen_golden.py` runs the unmodified reference
(imported from /root/reference through the stubs in `oracle/shim/`) on the
reference's own fixtures and on the synthetic BASELINE configurations and
freezes the outputs into `tests/golden/*.npz`; `tests/test_oracle.py` checks
every function below against those files (<= 1e-12 relative).

It keeps the reference's *call structure* on purpose (cdist + exp, dpotrf on a
Fortran copy, two general `np.linalg.solve` on the triangular factor) so that
timing it is a fair stand-in for timing the reference (`cpu_baseline.kind =
"port"`).  All citations are to files under /root/reference/pyGPs/.

A covariance function is described by a plain tuple
    ("rbf",    [log_ell, log_sf])                 Core/cov.py:786-808
    ("rbfard", [log_ell_1..log_ell_D, log_sf])    Core/cov.py:872-904
    ("matern", [log_ell, log_sf], d)              Core/cov.py:1078-1148
    ("rbfunit", [log_ell]) :832 | ("rq", [log_ell, log_sf, log_alpha]) :1304 | ("rqard", [ell.., log_sf, log_alpha]) :1356
    ("periodic", [log_ell, log_p, log_sf]) :1186 | ("piecepoly", [log_ell, log_sf], v) :683 | ("gabor", [log_ell, log_p]) :392
    ("noise", [log_s]) :1254 | ("const", [log_s]) :941 | ("linear", [log_s]) :986 | ("poly", [log_c, log_sf], d) :623
    ("pre", M1, M2) :1429
    ("sum", k1, k2) :265 | ("prod", k1, k2) :230 | ("scale", c, k) :299      (composites; hyp lists concatenate)
and a mean function by
    ("zero",) | ("one",) | ("const", c) | ("linear", [a_1..a_D])   Core/mean.py:279-369
"""
import numpy as np
import scipy.linalg.lapack as _lapack
from scipy.spatial.distance import cdist as _cdist

LOG2PI = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------
# mean functions  (Core/mean.py:279-369)
# --------------------------------------------------------------------------
def mean_vec(mean, x):
    """Prior mean m(x) as an (n,1) column.  Core/mean.py:285,303,323,356."""
    n = x.shape[0]
    kind = mean[0]
    if kind == "zero":
        return np.zeros((n, 1))
    if kind == "one":
        return np.ones((n, 1))
    if kind == "const":
        return float(mean[1]) * np.ones((n, 1))
    if kind == "linear":
        return x @ np.asarray(mean[1], dtype=float).reshape(-1, 1)
    raise ValueError(kind)


def mean_hyp(mean):
    """Flattened trainable parameters of a mean spec."""
    if mean[0] == "const":
        return [float(mean[1])]
    if mean[0] == "linear":
        return [float(v) for v in mean[1]]
    return []


def mean_der(mean, x, i):
    """d m(x) / d hyp_i as an (n,1) column.  Core/mean.py:290,308,328-335,361-369."""
    n = x.shape[0]
    if mean[0] == "const" and i == 0:
        return np.ones((n, 1))
    if mean[0] == "linear" and i < x.shape[1]:
        return x[:, i].reshape(n, 1).astype(float)
    return np.zeros((n, 1))


# --------------------------------------------------------------------------
# covariance functions
# --------------------------------------------------------------------------
def _matern_d(d):
    """Core/cov.py:1128-1136: d is rounded, anything outside {1,3,5,7} becomes 3."""
    if abs(d - round(d)) < 1e-8:
        d = int(round(d))
    d = int(d)
    return d if d in (1, 3, 5, 7) else 3


def _matern_poly(d, t):
    """Core/cov.py:1094-1104 (func)."""
    if d == 1:
        return 1.0 + 0.0 * t
    if d == 3:
        return 1.0 + t
    if d == 5:
        return 1.0 + t + t * t / 3.0
    return 1.0 + t + 2.0 * t * t / 5.0 + t * t * t / 15.0


def _matern_dpoly(d, t):
    """func - dfunc/dt, the factor of the length-scale derivative (Core/cov.py:1106-1116).
    For d=7 the reference has (t + 3t^2 + t^3)/15; the true value of func - func' is
    (3t + 3t^2 + t^3)/15 (finite-difference checked in tests/test_oracle.py), used here."""
    if d == 1:
        return 1.0 + 0.0 * t
    if d == 3:
        return t
    if d == 5:
        return (t + t * t) / 3.0
    return (3.0 * t + 3.0 * t * t + t * t * t) / 15.0


def _scaled_inputs(cov, x, z):
    """Inputs after the kernel's own length-scale transform, as the reference
    forms them before calling cdist (RBF divides, Core/cov.py:804; RBFard
    multiplies by 1/exp(hyp), :893,899; Matern multiplies by sqrt(d)/ell, :1141)."""
    kind, hyp = cov[0], cov[1]
    if kind == "rbf":
        ell = np.exp(hyp[0])
        f = lambda a: a / ell
    elif kind == "rbfard":
        D = (x if x is not None else z).shape[1]
        inv = 1.0 / np.exp(np.asarray(hyp[:D], dtype=float))
        f = lambda a: a * inv[None, :]
    elif kind == "matern":
        ell = np.exp(hyp[0])
        d = _matern_d(cov[2])
        f = lambda a: np.sqrt(d) * a / ell
    else:
        raise ValueError(kind)
    return (None if x is None else f(x)), (None if z is None else f(z))


def _sf2(cov, D):
    hyp = cov[1]
    return np.exp(2.0 * (hyp[D] if cov[0] == "rbfard" else hyp[1]))


def _native_cov_matrix(cov, x=None, z=None, mode=None):
    """getCovMatrix for RBF / RBFard / Matern.  Core/cov.py:796-808, 887-904, 1124-1148.

    mode 'train' -> (n,n), 'cross' -> (n,m), 'self_test' -> (m,1)."""
    if mode is None:
        raise Exception("Specify the mode: 'train' or 'cross'")
    if x is None and z is None:
        raise Exception("Specify at least one: training input (x) or test input (z) or both.")
    if mode == "cross" and (x is None or z is None):
        raise Exception("Specify both: training input (x) and test input (z) for cross covariance.")
    D = (x if x is not None else z).shape[1]
    sf2 = _sf2(cov, D)
    if mode == "self_test":
        A = np.zeros((z.shape[0], 1))
    else:
        xs, zs = _scaled_inputs(cov, x, z if mode == "cross" else None)
        A = _cdist(xs, xs if mode == "train" else zs, "sqeuclidean")
    if cov[0] == "matern":
        d = _matern_d(cov[2])
        t = np.sqrt(A)
        return sf2 * _matern_poly(d, t) * np.exp(-t)
    return sf2 * np.exp(-0.5 * A)


def _native_cov_der_matrix(cov, x=None, z=None, mode=None, der=None):
    """getDerMatrix: dK/dhyp_der.  RBF Core/cov.py:811-828, RBFard :906-938.

    For Matern this is the mathematically CORRECT derivative (sf2*dmfunc(t) for
    the length scale); the reference's Core/cov.py:1173-1177 overwrites the
    distance with K before using it and is wrong (SURVEY section 7, quirk list),
    so Matern dnlZ is checked by finite differences, not against the reference."""
    D = (x if x is not None else z).shape[1]
    sf2 = _sf2(cov, D)
    kind = cov[0]
    nh = D + 1 if kind == "rbfard" else 2
    if der is None:
        raise Exception("Specify the index of parameters of the derivatives.")
    if der < 0 or der >= nh:
        raise Exception("Wrong derivative index")
    if mode == "self_test":
        A = np.zeros((z.shape[0], 1))
    else:
        xs, zs = _scaled_inputs(cov, x, z if mode == "cross" else None)
        A = _cdist(xs, xs if mode == "train" else zs, "sqeuclidean")
    if kind == "matern":
        d = _matern_d(cov[2])
        t = np.sqrt(A)
        if der == 0:
            return sf2 * _matern_dpoly(d, t) * t * np.exp(-t)
        return 2.0 * sf2 * _matern_poly(d, t) * np.exp(-t)
    K = sf2 * np.exp(-0.5 * A)
    if kind == "rbf":
        return K * A if der == 0 else 2.0 * K
    # rbfard
    if der == D:
        return 2.0 * K
    if mode == "self_test":
        return K * 0.0
    inv = 1.0 / np.exp(cov[1][der])
    a = (x[:, der] * inv).reshape(-1, 1)
    b = a if mode == "train" else (z[:, der] * inv).reshape(-1, 1)
    return K * _cdist(a, b, "sqeuclidean")


# ---- the other kernels and the composites -----------------------------------------------------------------------
_NATIVE = ("rbf", "rbfard", "matern")


def cov_nhyp(cov, D):
    """Number of hyper-parameters of a covariance spec (the composite's flat list is the concatenation)."""
    k = cov[0]
    if k in ("sum", "prod"):
        return cov_nhyp(cov[1], D) + cov_nhyp(cov[2], D)
    if k == "scale":
        return 1 + cov_nhyp(cov[2], D)
    if k == "pre":
        return 0
    return len(cov[1])


def _dist2(cov_scale, x, z, mode):
    """cdist(x*s, z*s, 'sqeuclidean') in the reference's mode convention; zeros for self-test."""
    if mode == "self_test":
        return np.zeros((z.shape[0], 1))
    a = cov_scale(x)
    return _cdist(a, a if mode == "train" else cov_scale(z), "sqeuclidean")


def _pp(v, r, j, deriv=False):
    """PiecePoly pp / dpp.  Core/cov.py:696-727."""
    m = np.maximum(1.0 - r, 0.0)
    if v == 0:
        f, df = 1.0 + 0.0 * r, 0.0 * r
    elif v == 1:
        f, df = 1.0 + (j + 1) * r, (j + 1) + 0.0 * r
    elif v == 2:
        f = 1.0 + (j + 2) * r + (j * j + 4.0 * j + 3) / 3.0 * r * r
        df = (j + 2) + 2.0 * (j * j + 4.0 * j + 3.0) / 3.0 * r
    else:
        f = (1.0 + (j + 3) * r + (6.0 * j * j + 36.0 * j + 45.0) / 15.0 * r * r
             + (j * j * j + 9.0 * j * j + 23.0 * j + 15.0) / 15.0 * r * r * r)
        df = ((j + 3) + 2.0 * (6.0 * j * j + 36.0 * j + 45.0) / 15.0 * r
              + (j * j * j + 9.0 * j * j + 23.0 * j + 15.0) / 5.0 * r * r)
    if not deriv:
        return f * m ** (j + v)
    return m ** (j + v - 1) * r * ((j + v) * f - m * df)


def _leaf(cov, x, z, mode, der):
    """Value (der None) or derivative matrix of one non-native leaf kernel, formulas as in the reference."""
    k, hyp = cov[0], cov[1]
    D = (x if x is not None else z).shape[1]
    n = None if x is None else x.shape[0]
    nn = None if z is None else z.shape[0]
    if k == "rbfunit":                                           # Core/cov.py:842-868
        ell = np.exp(hyp[0])
        A = _dist2(lambda a: a / ell, x, z, mode)
        return np.exp(-0.5 * A) if der is None else np.exp(-0.5 * A) * A
    if k in ("rq", "rqard"):                                     # Core/cov.py:1316-1352, 1373-1425
        if k == "rq":
            ell = np.exp(hyp[0]); sf2 = np.exp(2.0 * hyp[1]); al = np.exp(hyp[2])
            D2 = _dist2(lambda a: a / ell, x, z, mode)
            nl = 1
        else:
            inv = 1.0 / np.exp(np.asarray(hyp[:D], dtype=float)); sf2 = np.exp(2.0 * hyp[D]); al = np.exp(hyp[D + 1])
            D2 = _dist2(lambda a: a * inv[None, :], x, z, mode)
            nl = D
        base = 1.0 + 0.5 * D2 / al
        if der is None:
            return sf2 * base ** (-al)
        if der < nl:
            if k == "rq":
                return sf2 * base ** (-al - 1) * D2
            # TRUE derivative: the reference's (:1413-1418) is identically zero in train mode (distance of two ROW vectors)
            if mode == "self_test":
                return D2 * 0
            a = (x[:, der] * inv[der]).reshape(-1, 1)
            b = a if mode == "train" else (z[:, der] * inv[der]).reshape(-1, 1)
            return sf2 * base ** (-al - 1) * _cdist(a, b, "sqeuclidean")
        if der == nl:
            return 2.0 * sf2 * base ** (-al)
        return sf2 * base ** (-al) * (0.5 * D2 / base - al * np.log(base))
    if k == "periodic":                                          # Core/cov.py:1198-1250
        ell = np.exp(hyp[0]); p = np.exp(hyp[1]); sf2 = np.exp(2.0 * hyp[2])
        A = np.pi * np.sqrt(_dist2(lambda a: a, x, z, mode)) / p
        R = np.sin(A) / ell
        if der is None:
            return sf2 * np.exp(-2.0 * R * R)
        if der == 0:
            return 4.0 * sf2 * np.exp(-2.0 * R * R) * R * R
        if der == 1:
            return 4.0 * sf2 / ell * np.exp(-2.0 * R * R) * R * np.cos(A) * A
        return 2.0 * sf2 * np.exp(-2.0 * R * R)
    if k == "piecepoly":                                         # Core/cov.py:729-782
        ell = np.exp(hyp[0]); sf2 = np.exp(2.0 * hyp[1]); v = int(round(cov[2]))
        j = np.floor(0.5 * D) + v + 1
        r = np.sqrt(_dist2(lambda a: a / ell, x, z, mode))
        if der is None:
            return sf2 * _pp(v, r, j)
        if der == 0:
            return sf2 * _pp(v, r, j, deriv=True)
        if der == 1:
            return 2.0 * sf2 * _pp(v, r, j)
        return r * 0
    if k == "gabor":                                             # Core/cov.py:413-448 (derivatives as the reference has them)
        ell = np.exp(hyp[0]); p = np.exp(2.0 * hyp[1])
        A = _dist2(lambda a: a / ell, x, z, mode)
        dp = 2 * np.pi * np.sqrt(A) * ell / p
        K = np.exp(-0.5 * A) * np.cos(dp)
        if der is None:
            return K
        return dp * K if der == 0 else dp * np.exp(-0.5 * A) * np.sin(dp)
    if k == "noise":                                             # Core/cov.py:1265-1300
        s2 = np.exp(2.0 * hyp[0])
        if mode == "self_test":
            A = np.zeros((nn, 1))
        elif mode == "train":
            A = np.eye(n)
        else:
            A = (_cdist(x, z, "sqeuclidean") < 1e-9).astype(float)
        return s2 * A if der is None else 2.0 * s2 * A
    if k == "const":                                             # Core/cov.py:950-982
        sf2 = np.exp(hyp[0])
        if mode == "self_test":
            A = np.ones((nn, 1))
        elif mode == "train":
            A = np.ones((n, n))
        else:
            A = np.ones((n, nn))
        if der is None:
            return sf2 * A + (np.eye(n) * 1e-10 if mode == "train" else 0.0)
        return 2.0 * sf2 * A
    if k in ("linear", "poly"):                                  # Core/cov.py:995-1024, 635-679
        if mode == "self_test":
            A = np.sum(z * z, 1).reshape(nn, 1)
        elif mode == "train":
            A = np.dot(x, x.T)
        else:
            A = np.dot(x, z.T)
        if k == "linear":
            sf2 = np.exp(hyp[0])
            if der is None:
                return sf2 * (A + (np.eye(n) * 1e-10 if mode == "train" else 0.0))
            return 2.0 * sf2 * (A + (np.eye(n) * 1e-16 if mode == "train" else 0.0))
        c = np.exp(hyp[0]); sf2 = np.exp(2.0 * hyp[1]); o = int(round(cov[2]))
        if der is None:
            return sf2 * (c + A + (np.eye(n) * 1e-10 if mode == "train" else 0.0)) ** o
        if der == 0:
            return c * o * sf2 * (c + A) ** (o - 1)
        if der == 1:
            return 2.0 * sf2 * (c + A) ** o
        return A * 0
    raise ValueError(k)


def cov_matrix(cov, x=None, z=None, mode=None):
    """getCovMatrix of any supported covariance spec, composites included (Core/cov.py:247-250, 281-284, 315-319)."""
    k = cov[0]
    if k in _NATIVE:
        return _native_cov_matrix(cov, x=x, z=z, mode=mode)
    if k == "sum":
        return cov_matrix(cov[1], x, z, mode) + cov_matrix(cov[2], x, z, mode)
    if k == "prod":
        return cov_matrix(cov[1], x, z, mode) * cov_matrix(cov[2], x, z, mode)
    if k == "scale":
        return np.exp(cov[1]) * cov_matrix(cov[2], x, z, mode)
    if k == "pre":                                               # Core/cov.py:1442-1450
        M1, M2 = cov[1], cov[2]
        if mode == "self_test":
            return M1[-1, :].reshape(-1, 1)
        return M2 if mode == "train" else M1[:-1, :]
    return _leaf(cov, x, z, mode, None)


def cov_der_matrix(cov, x=None, z=None, mode=None, der=None):
    """getDerMatrix of any supported covariance spec (Core/cov.py:252-261, 286-295, 321-328)."""
    k = cov[0]
    D = (x if x is not None else z).shape[1]
    if k in _NATIVE:
        return _native_cov_der_matrix(cov, x=x, z=z, mode=mode, der=der)
    if k in ("sum", "prod"):
        n1 = cov_nhyp(cov[1], D)
        if der < n1:
            A = cov_der_matrix(cov[1], x, z, mode, der)
            return A if k == "sum" else A * cov_matrix(cov[2], x, z, mode)
        A = cov_der_matrix(cov[2], x, z, mode, der - n1)
        return A if k == "sum" else A * cov_matrix(cov[1], x, z, mode)
    if k == "scale":
        sf2 = np.exp(cov[1])
        if der == 0:
            return 2.0 * sf2 * cov_matrix(cov[2], x, z, mode)
        return sf2 * cov_der_matrix(cov[2], x, z, mode, der - 1)
    return _leaf(cov, x, z, mode, der)


def fitc_cov_matrix(cov, xu, x=None, z=None, mode=None):
    """FITCOfKernel.getCovMatrix.  Core/cov.py:352-370.
    'train' -> (diagK (n,1), Kuu (M,M), Ku (M,n)); 'cross' -> K(xu,z); 'self_test' -> diag."""
    if x is not None and xu.shape[1] != x.shape[1]:
        raise Exception("Dimensionality of inducing inputs must match training inputs")
    if mode == "self_test":
        return cov_matrix(cov, z=z, mode="self_test")
    if mode == "train":
        return (cov_matrix(cov, z=x, mode="self_test"),
                cov_matrix(cov, x=xu, mode="train"),
                cov_matrix(cov, x=xu, z=x, mode="cross"))
    if mode == "cross":
        return cov_matrix(cov, x=xu, z=z, mode="cross")
    raise Exception("Specify the mode: 'train' or 'cross'")


def fitc_cov_der_matrix(cov, xu, x, der):
    """FITCOfKernel.getDerMatrix(mode='train').  Core/cov.py:372-390."""
    return (cov_der_matrix(cov, z=x, mode="self_test", der=der),
            cov_der_matrix(cov, x=xu, mode="train", der=der),
            cov_der_matrix(cov, x=xu, z=x, mode="cross", der=der))


# --------------------------------------------------------------------------
# dense linear-algebra helpers  (Core/tools.py:31-97)
# --------------------------------------------------------------------------
def jitchol(A):
    """Lower Cholesky factor through LAPACK dpotrf on a Fortran-ordered copy.
    Core/tools.py:60-77.  The reference's jitter retry is dead code (it calls
    np.linalg.cholesky with a non-existent `lower=` kwarg inside a bare except),
    so the observable contract is: not PD -> np.linalg.LinAlgError."""
    A = np.asfortranarray(A)
    L, info = _lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    if np.any(np.diag(A) <= 0.0):
        raise np.linalg.LinAlgError(
            "kernel matrix not positive definite: non-positive diagonal elements")
    raise np.linalg.LinAlgError("kernel matrix not positive definite, even with jitter.")


def solve_chol(R, B):
    """X = (R'R)^-1 B with R upper, done as the reference does it: two GENERAL
    solves (dgesv) on the triangular factor.  Core/tools.py:92-97."""
    if not (R.shape[0] == R.shape[1] and R.shape[0] == B.shape[0]):
        raise Exception("Wrong sizes of matrix arguments in solve_chol.py")
    return np.linalg.solve(R, np.linalg.solve(R.T, B))


# --------------------------------------------------------------------------
# inference engines
# --------------------------------------------------------------------------
def exact_evaluate(mean, cov, log_sn, x, y, nargout=1):
    """inf.Exact.evaluate.  Core/inf.py:353-384.

    Returns post dict {alpha (n,1), sW (n,1), L (n,n) upper} [, nlZ [, dnlZ dict
    {mean:[], cov:[], lik:[]}]]."""
    n = x.shape[0]
    K = cov_matrix(cov, x=x, mode="train")
    m = mean_vec(mean, x)
    sn2 = np.exp(2.0 * log_sn)
    R = jitchol(K / sn2 + np.eye(n)).T
    alpha = solve_chol(R, y - m) / sn2
    post = {"alpha": alpha, "sW": np.ones((n, 1)) / np.sqrt(sn2), "L": R}
    if nargout == 1:
        return post
    nlZ = (np.dot((y - m).T, alpha) / 2.0 + np.log(np.diag(R)).sum()
           + n * np.log(2.0 * np.pi * sn2) / 2.0)[0, 0]
    if nargout == 2:
        return post, nlZ
    Q = solve_chol(R, np.eye(n)) / sn2 - np.dot(alpha, alpha.T)
    nh = cov_nhyp(cov, x.shape[1])
    dn = {"lik": [sn2 * np.trace(Q)],
          "cov": [(Q * cov_der_matrix(cov, x=x, mode="train", der=i)).sum() / 2.0
                  for i in range(nh)],
          "mean": [np.dot(-mean_der(mean, x, i).T, alpha)[0, 0]
                   for i in range(len(mean_hyp(mean)))]}
    return post, nlZ, dn


def exact_evaluate_fair(cov, log_sn, x, y):
    """NOT the reference: the same nlZ through scipy's cho_factor/cho_solve
    (SURVEY 8(d) 'fair CPU' line), so the reported speed-up is not credited for
    the reference's LU-on-a-triangular-matrix waste."""
    import scipy.linalg as sla
    n = x.shape[0]
    sn2 = np.exp(2.0 * log_sn)
    A = cov_matrix(cov, x=x, mode="train")
    A /= sn2
    A[np.diag_indices(n)] += 1.0
    c = sla.cho_factor(A, lower=True, overwrite_a=True, check_finite=False)
    alpha = sla.cho_solve(c, y, check_finite=False) / sn2
    return (np.dot(y.T, alpha) / 2.0 + np.log(np.diag(c[0])).sum()
            + n * np.log(2.0 * np.pi * sn2) / 2.0)[0, 0]


def fitc_evaluate(mean, cov, xu, log_sn, x, y, nargout=1):
    """inf.FITC_Exact.evaluate.  Core/inf.py:398-455."""
    diagK, Kuu, Ku = fitc_cov_matrix(cov, xu, x=x, mode="train")
    m = mean_vec(mean, x)
    n = x.shape[0]
    nu = Kuu.shape[0]
    sn2 = np.exp(2.0 * log_sn)
    snu2 = 1.0e-6 * sn2                                    # :410
    Ruu = jitchol(Kuu + snu2 * np.eye(nu)).T               # :412
    V = np.linalg.solve(Ruu.T, Ku)                         # :413
    g = diagK + sn2 - (V * V).sum(axis=0).reshape(n, 1)    # :415
    Ru = jitchol(np.eye(nu) + np.dot(V / g.T, V.T)).T      # :417
    r = (y - m) / np.sqrt(g)
    be = np.linalg.solve(Ru.T, np.dot(V, r / np.sqrt(g)))  # :419
    iKuu = solve_chol(Ruu, np.eye(nu))                     # :420
    post = {"alpha": np.linalg.solve(Ruu, np.linalg.solve(Ru, be)),
            "L": solve_chol(np.dot(Ru, Ruu), np.eye(nu)) - iKuu,
            "sW": np.ones((n, 1)) / np.sqrt(sn2)}
    if nargout == 1:
        return post
    nlZ = (np.log(np.diag(Ru)).sum()
           + (np.log(g).sum() + n * LOG2PI + np.dot(r.T, r) - np.dot(be.T, be)) / 2.0)[0, 0]
    if nargout == 2:
        return post, nlZ
    al = r / np.sqrt(g) - np.dot(V.T, np.linalg.solve(Ru, be)) / g      # :431
    B = np.dot(iKuu, Ku)
    w = np.dot(B, al)
    W = np.linalg.solve(Ru.T, V / g.T)
    WW = (W * W).sum(axis=0).reshape(1, n)
    BW = np.dot(B, W.T)
    dcov = []
    for i in range(len(cov[1])):
        ddiag, dKuu, dKu = fitc_cov_der_matrix(cov, xu, x, i)
        Rm = 2.0 * dKu - np.dot(dKuu, B)
        v = ddiag - (Rm * B).sum(axis=0).reshape(n, 1)
        val = (np.dot(ddiag.T, 1.0 / g) + np.dot(w.T, np.dot(dKuu, w) - 2.0 * np.dot(dKu, al))
               - np.dot(al.T, v * al) - np.dot(WW, v) - (np.dot(Rm, W.T) * BW).sum()) / 2.0
        dcov.append(val[0, 0])
    dlik = sn2 * ((1.0 / g).sum() - WW.sum() - np.dot(al.T, al))
    dKuu_s = 2.0 * snu2
    Rm = -dKuu_s * B
    v = -(Rm * B).sum(axis=0).reshape(n, 1)
    dlik = dlik + (np.dot(w.T, dKuu_s * w) - np.dot(al.T, v * al) - np.dot(WW, v)
                   - (np.dot(Rm, W.T) * BW).sum()) / 2.0
    dn = {"cov": dcov, "lik": [dlik[0, 0]],
          "mean": [np.dot(-mean_der(mean, x, i).T, al)[0, 0]
                   for i in range(len(mean_hyp(mean)))]}
    return post, nlZ, dn


# --------------------------------------------------------------------------
# prediction  (Core/gp.py:388-437 with lik.Gauss, Core/lik.py:135-158)
# --------------------------------------------------------------------------
def predict(mean, cov, log_sn, x, post, xs, ys=None, xu=None, nperbatch=1000):
    """GP.predict for a Gaussian likelihood.  Returns (ymu, ys2, fmu, fs2, lp|None).
    `xu` not None selects the FITC branch (dense post['L'], Core/gp.py:418)."""
    alpha, R, sW = post["alpha"], post["L"], post["sW"]
    tri = bool(np.all(np.tril(R, -1) == 0))                 # Core/gp.py:393
    ns = xs.shape[0]
    sn2 = np.exp(2.0 * log_sn)
    fmu = np.zeros((ns, 1))
    fs2 = np.zeros((ns, 1))
    for lo in range(0, ns, nperbatch):
        sl = slice(lo, min(lo + nperbatch, ns))
        zb = xs[sl]
        if xu is None:
            kss = cov_matrix(cov, z=zb, mode="self_test")
            Ks = cov_matrix(cov, x=x, z=zb, mode="cross")
        else:
            kss = fitc_cov_matrix(cov, xu, z=zb, mode="self_test")
            Ks = fitc_cov_matrix(cov, xu, x=x, z=zb, mode="cross")
        fmu[sl] = mean_vec(mean, zb) + np.dot(Ks.T, alpha)
        if tri:
            V = np.linalg.solve(R.T, sW * Ks)                # Core/gp.py:415
            fs2[sl] = kss - (V * V).sum(axis=0).reshape(-1, 1)
        else:
            fs2[sl] = kss + (Ks * np.dot(R, Ks)).sum(axis=0).reshape(-1, 1)
        fs2[sl] = np.maximum(fs2[sl], 0.0)
    ymu = fmu.copy()
    ys2 = fs2 + sn2                                          # Core/lik.py:150-154
    lp = None
    if ys is not None:
        # Core/lik.py:147-148 -> EP-mode Gauss log partition, :160
        lp = -(ys - fmu) ** 2 / (sn2 + fs2) / 2.0 - np.log(2.0 * np.pi * (sn2 + fs2)) / 2.0
    return ymu, ys2, fmu, fs2, lp


# --------------------------------------------------------------------------
# synthetic workloads of BASELINE.json / SURVEY section 8(d)
# --------------------------------------------------------------------------
def synth_regression(N, D, seed=0):
    """`rng=default_rng(seed); X=standard_normal((N,D)); y=sin(X.sum(1))+0.1*standard_normal`."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    return X, y


# --------------------------------------------------------------------------
# lik.Erf + inf.EP  (BASELINE config 5; Core/lik.py:236-366, Core/inf.py:174-189, 723-806)
# --------------------------------------------------------------------------
from scipy.special import erf as _erf  # noqa: E402


def erf_logphi(z, p):
    """Safe log of the normal cdf.  Core/lik.py:354-366 (thresholds -6.2 / -5.5)."""
    z = np.asarray(z, dtype=float)
    lp = np.zeros_like(z)
    zmin, zmax = -6.2, -5.5
    ok = z > zmax
    bd = z < zmin
    nok = ~ok
    ip = nok & ~bd
    lam = 1.0 / (1.0 + np.exp(25.0 * (0.5 - (z[ip] - zmin) / (zmax - zmin))))
    lp[ok] = np.log(p[ok])
    lp[nok] = -np.log(np.pi) / 2.0 - z[nok] ** 2 / 2.0 - np.log(np.sqrt(z[nok] ** 2 / 2.0 + 2.0) - z[nok] / np.sqrt(2.0))
    lp[ip] = (1 - lam) * lp[ip] + lam * np.log(p[ip])
    return lp


def erf_cum_gauss(y, f):
    """(p, log p) of the probit likelihood.  Core/lik.py:328-339."""
    yf = y * f
    p = (1.0 + _erf(yf / np.sqrt(2.0))) / 2.0
    return p, erf_logphi(yf, p)


def erf_gau_over_cum_gauss(f, p):
    """N(f)/Phi(f) with the asymptotic branch below -6 and interpolation on [-6,-5].  Core/lik.py:341-352."""
    f = np.asarray(f, dtype=float)
    n_p = np.zeros_like(f)
    ok = f > -5
    n_p[ok] = (np.exp(-f[ok] ** 2 / 2) / np.sqrt(2 * np.pi)) / p[ok]
    bd = f < -6
    n_p[bd] = np.sqrt(f[bd] ** 2 / 4 + 1) - f[bd] / 2
    it = ~ok & ~bd
    tmp = f[it]
    lam = -5.0 - f[it]
    n_p[it] = (1 - lam) * (np.exp(-tmp ** 2 / 2) / np.sqrt(2 * np.pi)) / p[it] + lam * (np.sqrt(tmp ** 2 / 4 + 1) - tmp / 2)
    return n_p


def erf_ep_moments(y, mu, s2, nargout=1):
    """lZ [, dlZ, d2lZ] of int Phi(y f) N(f|mu,s2) df.  Core/lik.py:295-311."""
    y = np.sign(np.asarray(y, dtype=float))
    y = np.where(y == 0, 1.0, y)
    z = mu / np.sqrt(1 + s2)
    _, lZ = erf_cum_gauss(y, z)
    if nargout == 1:
        return lZ
    z = z * y
    n_p = erf_gau_over_cum_gauss(z, np.exp(lZ))
    dlZ = y * n_p / np.sqrt(1.0 + s2)
    if nargout == 2:
        return lZ, dlZ
    return lZ, dlZ, -n_p * (z + n_p) / (1.0 + s2)


def erf_predict(ys, fmu, fs2):
    """Prediction mode of lik.Erf (Core/lik.py:253-271): returns (lp, ymu, ys2)."""
    y = np.ones_like(fmu) if ys is None else np.where(np.sign(ys) == 0, 1.0, np.sign(ys)) * np.ones_like(fmu)
    if fs2 is not None and np.linalg.norm(fs2) > 0:
        lp = erf_ep_moments(y, fmu, fs2, 1)
        p = np.exp(lp)
    else:
        p, lp = erf_cum_gauss(y, fmu)
    return lp, 2 * p - 1, 4 * p * (1 - p)


def _ep_compute_params(K, y, ttau, tnu, m):
    """Core/inf.py:174-189."""
    n = len(y)
    ssi = np.sqrt(ttau)
    R = jitchol(np.eye(n) + np.dot(ssi, ssi.T) * K).T
    V = np.linalg.solve(R.T, np.tile(ssi, (1, n)) * K)
    Sigma = K - np.dot(V.T, V)
    mu = np.dot(Sigma, tnu)
    Ds = np.diag(Sigma).reshape(-1, 1)
    tau_n = 1 / Ds - ttau
    nu_n = mu / Ds - tnu + m * tau_n
    lZ = erf_ep_moments(y, nu_n / tau_n, 1 / tau_n, 1)
    nlZ = (np.log(np.diag(R)).sum() - lZ.sum() - np.dot(tnu.T, np.dot(Sigma, tnu)) / 2
           - np.dot((nu_n - m * tau_n).T, ((ttau / tau_n * (nu_n - m * tau_n) - 2 * tnu) / (ttau + tau_n))) / 2
           + (tnu ** 2 / (tau_n + ttau)).sum() / 2.0 - np.log(1.0 + ttau / tau_n).sum() / 2.0)
    return Sigma, mu, nlZ[0], R


def ep_evaluate(mean, cov, x, y, nargout=2, last=None):
    """inf.EP.evaluate with lik.Erf.  Core/inf.py:731-806.  `last` = (ttau, tnu) warm start or None.
    Returns post {alpha, sW, L}, nlZ [, dnlZ], and the site parameters (ttau, tnu) and sweep count."""
    tol, max_sweep, min_sweep = 1e-4, 10, 2
    n = x.shape[0]
    K = cov_matrix(cov, x=x, mode="train")
    m = mean_vec(mean, x)
    nlZ0 = -erf_ep_moments(y, m, np.diag(K).reshape(-1, 1), 1).sum()
    if last is None:
        ttau, tnu = np.zeros((n, 1)), np.zeros((n, 1))
        Sigma, mu, nlZ = K.copy(), np.zeros((n, 1)), nlZ0
    else:
        ttau, tnu = last[0].copy(), last[1].copy()
        Sigma, mu, nlZ, R = _ep_compute_params(K, y, ttau, tnu, m)
        if nlZ > nlZ0:
            ttau, tnu = np.zeros((n, 1)), np.zeros((n, 1))
            Sigma, mu, nlZ = K.copy(), np.zeros((n, 1)), nlZ0
    nlZ_old, sweep = np.inf, 0
    while (np.abs(nlZ - nlZ_old) > tol and sweep < max_sweep) or sweep < min_sweep:
        nlZ_old = nlZ
        sweep += 1
        for i in range(n):
            tau_ni = 1 / Sigma[i, i] - ttau[i]
            nu_ni = mu[i] / Sigma[i, i] + m[i] * tau_ni - tnu[i]
            lZ, dlZ, d2lZ = erf_ep_moments(y[i], nu_ni / tau_ni, 1 / tau_ni, 3)
            ttau_old = ttau[i].copy()
            ttau[i] = max(-d2lZ / (1.0 + d2lZ / tau_ni), 0)
            tnu[i] = (dlZ + (m[i] - nu_ni / tau_ni) * d2lZ) / (1.0 + d2lZ / tau_ni)
            ds2 = ttau[i] - ttau_old
            si = Sigma[:, i].reshape(-1, 1)
            Sigma = Sigma - ds2 / (1.0 + ds2 * si[i]) * np.dot(si, si.T)
            mu = np.dot(Sigma, tnu)
        Sigma, mu, nlZ, R = _ep_compute_params(K, y, ttau, tnu, m)
    sW = np.sqrt(ttau)
    alpha = tnu - sW * solve_chol(R, sW * np.dot(K, tnu))
    post = {"alpha": alpha, "sW": sW, "L": R}
    extra = {"ttau": ttau, "tnu": tnu, "sweeps": sweep}
    if nargout <= 2:
        return post, nlZ, extra
    V = np.linalg.solve(R.T, np.tile(sW, (1, n)) * K)
    Sigma = K - np.dot(V.T, V)
    mu = np.dot(Sigma, tnu)
    Ds = np.diag(Sigma).reshape(-1, 1)
    tau_n = 1 / Ds - ttau
    nu_n = mu / Ds - tnu
    F = np.dot(alpha, alpha.T) - np.tile(sW, (1, n)) * solve_chol(R, np.diag(sW.reshape(-1)))
    dn = {"cov": [-(F * cov_der_matrix(cov, x=x, mode="train", der=j)).sum() / 2.0 for j in range(len(cov[1]))],
          "lik": [], "mean": []}
    _, dlZ = erf_ep_moments(y, nu_n / tau_n, 1 / tau_n, 2)
    for i in range(len(mean_hyp(mean))):
        dn["mean"].append(-np.dot(dlZ.T, mean_der(mean, x, i))[0, 0])
    return post, nlZ, dn, extra


def predict_class(mean, cov, x, post, xs, ys=None, nperbatch=1000):
    """GP.predict for GPC (EP posterior, lik.Erf).  Core/gp.py:388-437."""
    alpha, R, sW = post["alpha"], post["L"], post["sW"]
    ns = xs.shape[0]
    fmu = np.zeros((ns, 1)); fs2 = np.zeros((ns, 1))
    for lo in range(0, ns, nperbatch):
        sl = slice(lo, min(lo + nperbatch, ns))
        kss = cov_matrix(cov, z=xs[sl], mode="self_test")
        Ks = cov_matrix(cov, x=x, z=xs[sl], mode="cross")
        fmu[sl] = mean_vec(mean, xs[sl]) + np.dot(Ks.T, alpha)
        V = np.linalg.solve(R.T, sW * Ks)
        fs2[sl] = np.maximum(kss - (V * V).sum(axis=0).reshape(-1, 1), 0)
    lp, ymu, ys2 = erf_predict(ys, fmu, fs2)
    return ymu, ys2, fmu, fs2, (lp if ys is not None else None), lp
