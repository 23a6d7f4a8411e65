"""Model facade: GP, GPR, GP_FITC, GPR_FITC.

Drop-in for the regression models of pyGPs.Core.gp (/root/reference/pyGPs/Core/gp.py:
GP :62-527, GPR :533-635, GP_FITC :934-1009, GPR_FITC :1015-1114): same attributes
(nlZ, dnlZ, posterior, ym, ys2, fm, fs2, lp), same setData / setPrior / setNoise /
setOptimizer / getPosterior / optimize / predict / predict_with_posterior signatures
and return arities.  Classification models, Laplace/EP engines and plotting are
outside the accelerated path and are not provided.
"""
import itertools
import logging
from copy import deepcopy

import numpy as np

from . import cov, inf, lik, mean, opt
from .cov import FITCOfKernel
from .tools import jitchol, solve_chol


def _as_col(a):
    if a is not None and a.ndim == 1:
        a = np.reshape(a, (a.shape[0], 1))
    return a


class GP(object):
    """Base class for GP models (Core/gp.py:62-527)."""

    def __init__(self):
        super(GP, self).__init__()
        self.usingDefaultMean = True
        self.meanfunc = None
        self.covfunc = None
        self.likfunc = None
        self.inffunc = None
        self.optimizer = None
        self.nlZ = None
        self.dnlZ = None
        self.posterior = None
        self.x = None
        self.y = None
        self.xs = None
        self.ys = None
        self.ym = None
        self.ys2 = None
        self.fm = None
        self.fs2 = None
        self.lp = None
        self.devices = None          # see setDevices
        self.logger = logging.getLogger(__name__)

    def __str__(self):
        return ('To get the properties of the model use:\n'
                'model.nlZ          # negative log marginal likelihood\n'
                'model.dnlZ.cov     # derivatives of cov func of negative log marginal likelihood\n'
                'model.dnlZ.lik     # derivatives of lik func of negative log marginal likelihood\n'
                'model.dnlZ.mean    # derivatives of mean func of negative log marginal likelihood\n'
                'model.posterior    # posterior structure\n'
                'model.covfunc.hyp  # hyperparameters of cov func\n'
                'model.meanfunc.hyp # hyperparameters of mean func\n'
                'model.likfunc.hyp  # hyperparameters of lik func\n'
                'model.fm           # latent mean\n'
                'model.fs2          # latent variance\n'
                'model.ym           # predictive mean\n'
                'model.ys2          # predictive variance\n'
                'model.lp           # log predictive probability')

    def __repr__(self):
        return str(type(self)) + ': ' + self.__str__()

    # ------------------------------------------------------------------ data / prior
    def setData(self, x, y):
        """Set training inputs/targets; 1-d arrays become columns; the default Zero mean
        becomes mean.Const(mean(y)) (Core/gp.py:131-156)."""
        assert x.shape[0] == y.shape[0], "number of inputs and labels does not match"
        self.x = _as_col(x)
        self.y = _as_col(y)
        if self.usingDefaultMean:
            self.meanfunc = mean.Const(np.mean(y))

    def setPrior(self, mean=None, kernel=None):
        """Core/gp.py:205-222."""
        from . import mean as mean_mod
        if mean is not None:
            assert isinstance(mean, mean_mod.Mean), "mean function is not an instance of pyGPs.mean.Mean"
            self.meanfunc = mean
            self.usingDefaultMean = False
        if kernel is not None:
            assert isinstance(kernel, cov.Kernel), "cov function is not an instance of pyGPs.cov.Kernel"
            self.covfunc = kernel

    def setOptimizer(self, method, num_restarts=None, min_threshold=None, meanRange=None, covRange=None,
                     likRange=None):
        pass

    def setDevices(self, devices=None, shard=None):
        """Multi-GPU use of this model (no counterpart in the reference, which has no parallelism at all):
        `devices` = CUDA ordinals (None: every visible GPU).  The optimizers run their random restarts
        (Core/opt.py:301-327) concurrently, one per GPU; an exact evaluation whose factor does not fit one GPU - or every
        evaluation with shard=True - is sharded over all of them (gpk_exact_eval_dist, NCCL panel broadcasts)."""
        self.devices = None if devices is None else [int(d) for d in devices]
        if self.inffunc is not None:
            self.inffunc.devices = self.devices
            self.inffunc.shard = shard
            self.inffunc._engine = None

    def _take_xy(self, x, y):
        if x is not None and y is not None:
            assert x.shape[0] == y.shape[0], "number of inputs and labels does not match"
        if x is not None:
            self.x = _as_col(x)
        if y is not None:
            self.y = _as_col(y)
        if self.usingDefaultMean and self.meanfunc is None:
            self.meanfunc = mean.Const(np.mean(y))

    # ------------------------------------------------------------------ training
    def optimize(self, x=None, y=None, numIterations=40):
        """Train hyper-parameters with the model's optimizer (Core/gp.py:251-285)."""
        self._take_xy(x, y)
        optimalHyp, optimalNlZ = self.optimizer.findMin(self.x, self.y, numIters=numIterations)
        self.nlZ = optimalNlZ
        self.optimizer._apply_in_objects(optimalHyp)
        self.getPosterior()

    def getPosterior(self, x=None, y=None, der=True):
        """nlZ, dnlZ, post = getPosterior(x, y, der=True);  nlZ, post = getPosterior(x, y, der=False)
        (Core/gp.py:289-345)."""
        self._take_xy(x, y)
        if not der:
            post, nlZ = self.inffunc.evaluate(self.meanfunc, self.covfunc, self.likfunc, self.x, self.y, 2)
            self.nlZ = nlZ
            self.posterior = deepcopy(post)
            return nlZ, post
        post, nlZ, dnlZ = self.inffunc.evaluate(self.meanfunc, self.covfunc, self.likfunc, self.x, self.y, 3)
        self.nlZ = nlZ
        self.dnlZ = deepcopy(dnlZ)
        self.posterior = deepcopy(post)
        return nlZ, dnlZ, post

    # ------------------------------------------------------------------ prediction
    def predict(self, xs, ys=None):
        """ym, ys2, fm, fs2, lp = predict(xs[, ys])  (Core/gp.py:349-437)."""
        xs = _as_col(xs)
        self.xs = xs
        if ys is not None:
            ys = _as_col(ys)
            self.ys = ys
        if self.posterior is None:
            self.getPosterior()
        out = self._predict_core(self.posterior, xs, ys)
        self.ym, self.ys2, self.fm, self.fs2, self.lp = out[0], out[1], out[2], out[3], out[5]
        return out[:5]

    def predict_with_posterior(self, post, xs, ys=None):
        """Same as predict with a caller-supplied posterior (Core/gp.py:441-527)."""
        xs = _as_col(xs)
        self.xs = xs
        if ys is not None:
            ys = _as_col(ys)
            self.ys = ys
        self.posterior = deepcopy(post)
        out = self._predict_core(post, xs, ys)
        self.ym, self.ys2, self.fm, self.fs2, self.lp = out[0], out[1], out[2], out[3], out[5]
        return out[:5]

    def _posterior_matches(self, post, kindname):
        """Is `post` the posterior the engine still holds, built from the CURRENT hyper-parameters?"""
        spec = getattr(post, '_spec', None)
        if spec is None or spec[0] != kindname or not post._resident():
            return False
        # the GPU posterior stands in for `post` only if post still IS that evaluation's result for these inputs
        if post._vsig is None or inf._vec_signature(post) != post._vsig:
            return False
        if self.x is None or post._xsig is None or inf._x_signature(self.x) != post._xsig:
            return False
        dev = self.covfunc._device_spec()
        if dev is None:
            prog = self.covfunc._device_prog() if not isinstance(self.covfunc, FITCOfKernel) else None
            if prog is None:
                return False
            dev = ('prog', tuple(prog[0]), prog[1])
        kind, md, hyp = dev
        lik0 = float(self.likfunc.hyp[0]) if self.likfunc.hyp else 0.0
        return (spec[1] == kind and spec[2] == md and tuple(hyp) == spec[3] and lik0 == spec[4])

    def _predict_core(self, post, xs, ys):
        meanfunc, covfunc, likfunc = self.meanfunc, self.covfunc, self.likfunc
        x = self.x
        ns = xs.shape[0]
        fitc = isinstance(covfunc, FITCOfKernel)
        kindname = 'fitc' if fitc else ('ep' if isinstance(self.inffunc, inf.EP) else 'exact')
        if self._posterior_matches(post, kindname):
            # fast path: cross-covariances, the triangular solves and the column reductions
            # all happen next to the resident factor (gpk_predict / gpk_fitc_predict)
            eng = post._engine
            ka, fs2 = eng.fitc_predict(xs) if fitc else eng.predict(xs)
            fmu = meanfunc.getMean(xs) + ka
        else:
            fmu, fs2 = self._predict_host_posterior(post, xs)
        if ys is None:
            lp_all, ymu, ys2 = likfunc.evaluate(None, fmu, fs2, None, None, 3)
        else:
            lp_all, ymu, ys2 = likfunc.evaluate(ys, fmu, fs2, None, None, 3)
        lp_all = np.reshape(lp_all, (ns, 1))
        return ymu, ys2, fmu, fs2, (None if ys is None else lp_all), lp_all

    def _predict_host_posterior(self, post, xs):
        """Posterior given as host arrays (predict_with_posterior, a caller-edited posterior): the reference's
        batch loop (Core/gp.py:395-419); the factor is uploaded ONCE (gpk_set_factor) and every batch is two
        triangular sweeps on the GPU (gpk_potrs)."""
        from . import _lib
        covfunc, meanfunc, x = self.covfunc, self.meanfunc, self.x
        alpha, L, sW = post.alpha, post.L, post.sW
        if len(L) == 0:
            K = covfunc.getCovMatrix(x=x, mode='train')
            L = jitchol((np.eye(x.shape[0]) + np.dot(sW, sW.T) * K).T).T
        Ltril = np.all(np.tril(L, -1) == 0)
        eng = None
        if Ltril:
            eng = _lib.shared_engine()
            eng.set_factor(np.asarray(L, dtype=np.float64))
        ns = xs.shape[0]
        fmu = np.zeros((ns, 1))
        fs2 = np.zeros((ns, 1))
        nperbatch = 1000
        for lo in range(0, ns, nperbatch):
            ids = slice(lo, min(lo + nperbatch, ns))
            kss = covfunc.getCovMatrix(z=xs[ids, :], mode='self_test')
            Ks = covfunc.getCovMatrix(x=x, z=xs[ids, :], mode='cross')
            fmu[ids] = meanfunc.getMean(xs[ids, :]) + np.dot(Ks.T, alpha)
            if Ltril:
                # colsum(V*V) with V = L'^-1 (sW*Ks) equals colsum(B * (L'L)^-1 B), B = sW*Ks
                B = sW * Ks
                fs2[ids] = kss - (B * eng.potrs(B)).sum(axis=0).reshape(-1, 1)
            else:
                fs2[ids] = kss + (Ks * np.dot(L, Ks)).sum(axis=0).reshape(-1, 1)
            fs2[ids] = np.maximum(fs2[ids], 0)
        return fmu, fs2


class GPR(GP):
    """Gaussian-process regression (Core/gp.py:533-635).  `devices` / `shard`: see GP.setDevices."""

    def __init__(self, devices=None, shard=None):
        super(GPR, self).__init__()
        self.meanfunc = mean.Zero()
        self.covfunc = cov.RBF()
        self.likfunc = lik.Gauss()
        self.inffunc = inf.Exact()
        self.optimizer = opt.Minimize(self)
        if devices is not None or shard is not None:
            self.setDevices(devices, shard)

    def setNoise(self, log_sigma):
        """Replace the default noise (log 0.1) - Core/gp.py:547-553."""
        self.likfunc = lik.Gauss(log_sigma)

    def setOptimizer(self, method, num_restarts=None, min_threshold=None, meanRange=None, covRange=None,
                     likRange=None):
        """Core/gp.py:556-582."""
        conf = None
        if (num_restarts is not None) or (min_threshold is not None):
            conf = opt.random_init_conf(self.meanfunc, self.covfunc, self.likfunc)
            conf.num_restarts = num_restarts
            conf.min_threshold = min_threshold
            if meanRange is not None:
                conf.meanRange = meanRange
            if covRange is not None:
                conf.covRange = covRange
            if likRange is not None:
                conf.likRange = likRange
        table = {"Minimize": opt.Minimize, "SCG": opt.SCG, "CG": opt.CG, "BFGS": opt.BFGS,
                 "Nelder-Mead": opt.Simplex}
        if method not in table:
            raise Exception('Optimization method is not set correctly in setOptimizer')
        self.optimizer = table[method](self, conf)

    def useInference(self, newInf):
        raise Exception('Only exact inference is on the accelerated path; "Laplace" and "EP" are out of scope.')

    def useLikelihood(self, newLik):
        raise Exception('Only the Gaussian likelihood is on the accelerated path.')


class GPC(GP):
    """Gaussian-process binary classification: lik.Erf + inf.EP (Core/gp.py:641-732)."""

    def __init__(self):
        super(GPC, self).__init__()
        self.meanfunc = mean.Zero()
        self.covfunc = cov.RBF()
        self.likfunc = lik.Erf()
        self.inffunc = inf.EP()
        self.optimizer = opt.Minimize(self)

    def getPosterior(self, x=None, y=None, der=True):
        self._take_xy(x, y)
        uy = np.unique(self.y)                       # labels must be +1 / -1 (Core/gp.py:329-333)
        if np.any((uy != 1) & (uy != -1)):
            raise Exception('You attempt classification using labels different from {+1,-1}')
        return super(GPC, self).getPosterior(der=der)

    def setOptimizer(self, method, num_restarts=None, min_threshold=None, meanRange=None, covRange=None,
                     likRange=None):
        conf = None
        if (num_restarts is not None) or (min_threshold is not None):
            conf = opt.random_init_conf(self.meanfunc, self.covfunc, self.likfunc)
            conf.num_restarts = num_restarts
            conf.min_threshold = min_threshold
            if meanRange is not None:
                conf.meanRange = meanRange
            if covRange is not None:
                conf.covRange = covRange
            if likRange is not None:
                conf.likRange = likRange
        table = {"Minimize": opt.Minimize, "SCG": opt.SCG, "CG": opt.CG, "BFGS": opt.BFGS}
        if method in table:
            self.optimizer = table[method](self, conf)

    def useInference(self, newInf):
        raise Exception('Only EP inference is on the accelerated classification path ("Laplace" is out of scope).')

    def useLikelihood(self, newLik):
        if newLik == "Logistic":
            raise Exception("Logistic likelihood is currently not implemented.")
        raise Exception('Possible lik values are "Logistic".')


class GP_FITC(GP):
    """Base class for FITC models (Core/gp.py:934-1009)."""

    def __init__(self):
        super(GP_FITC, self).__init__()
        self.u = None

    def setData(self, x, y, value_per_axis=5):
        """As GP.setData, plus a default inducing grid of `value_per_axis` values per input
        dimension spanning the data range (Core/gp.py:944-983)."""
        assert x.shape[0] == y.shape[0], "number of inputs and labels does not match"
        x = _as_col(x)
        y = _as_col(y)
        self.x = x
        self.y = y
        if self.usingDefaultMean:
            self.meanfunc = mean.Const(np.mean(y))
        if self.u is None:
            axes = [np.linspace(np.min(x[:, d]), np.max(x[:, d]), value_per_axis) for d in range(x.shape[1])]
            self.u = np.array(list(itertools.product(*axes)))
            self.covfunc = self.covfunc.fitc(self.u)

    def setPrior(self, mean=None, kernel=None, inducing_points=None):
        """Core/gp.py:987-1009."""
        if kernel is not None:
            if inducing_points is not None:
                self.covfunc = kernel.fitc(inducing_points)
                self.u = inducing_points
            else:
                if self.u is not None:
                    self.covfunc = kernel.fitc(self.u)
                else:
                    raise Exception("To use default inducing points, please call setData() first!")
        if mean is not None:
            self.meanfunc = mean
            self.usingDefaultMean = False


class GPR_FITC(GP_FITC):
    """FITC regression (Core/gp.py:1015-1114)."""

    def __init__(self):
        super(GPR_FITC, self).__init__()
        self.meanfunc = mean.Zero()
        self.covfunc = cov.RBF()
        self.likfunc = lik.Gauss()
        self.inffunc = inf.FITC_Exact()
        self.optimizer = opt.Minimize(self)
        self.u = None

    def setNoise(self, log_sigma):
        self.likfunc = lik.Gauss(log_sigma)

    def setOptimizer(self, method, num_restarts=None, min_threshold=None, meanRange=None, covRange=None,
                     likRange=None):
        conf = None
        if (num_restarts is not None) or (min_threshold is not None):
            conf = opt.random_init_conf(self.meanfunc, self.covfunc, self.likfunc)
            conf.num_restarts = num_restarts
            conf.min_threshold = min_threshold
            if meanRange is not None:
                conf.meanRange = meanRange
            if covRange is not None:
                conf.covRange = covRange
            if likRange is not None:
                conf.likRange = likRange
        table = {"Minimize": opt.Minimize, "SCG": opt.SCG, "CG": opt.CG, "BFGS": opt.BFGS}
        if method in table:
            self.optimizer = table[method](self, conf)

    def useInference(self, newInf):
        raise Exception('Only FITC exact inference is on the accelerated path.')

    def useLikelihood(self, newLik):
        raise Exception('Only the Gaussian likelihood is on the accelerated path.')
