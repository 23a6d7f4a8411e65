"""Optimizer drivers: the callers that turn nlZ evaluations/s into wall-clock.

Same classes and `findMin(x, y, numIters) -> (optimalHyp, funcValue)` contract as
pyGPs.Core.opt (/root/reference/pyGPs/Core/opt.py:35-383) and the same random-restart
semantics (`random_init_conf`, /root/reference/pyGPs/Optimization/conf.py:18-54).  All of
this is host scalar logic; every objective evaluation is one GPU call through
`model.getPosterior`.

`minimize()` below is a fresh implementation of Carl Rasmussen's published
conjugate-gradient routine (Polack-Ribiere directions, cubic/quadratic line search under
the Wolfe-Powell conditions) that the reference ships as Optimization/minimize.py:41-172;
`scg()` is Moller's scaled conjugate gradient (the algorithm of Optimization/scg.py).
CG / BFGS / Nelder-Mead delegate to scipy.optimize exactly as the reference does.
"""
import logging
from copy import deepcopy

import numpy as np
from scipy.optimize import fmin as _simplex
from scipy.optimize import fmin_bfgs as _bfgs
from scipy.optimize import fmin_cg as _cg


class random_init_conf(object):
    """Ranges for random restarts; default (-5, 5) per hyper-parameter."""

    def __init__(self, mean, cov, lik):
        self.num_restarts = None
        self.min_threshold = None
        self.mean = mean
        self.cov = cov
        self.lik = lik
        self._meanRange = [(-5, 5) for i in mean.hyp]
        self._covRange = [(-5, 5) for i in cov.hyp]
        self._likRange = [(-5, 5) for i in lik.hyp]

    def _checked(self, value, ref, what):
        if len(value) != len(ref):
            raise Exception('The length of %s is not consistent with number of hyparameters' % what)
        return value

    meanRange = property(lambda s: s._meanRange,
                         lambda s, v: setattr(s, '_meanRange', s._checked(v, s.mean.hyp, 'meanRange')))
    covRange = property(lambda s: s._covRange,
                        lambda s, v: setattr(s, '_covRange', s._checked(v, s.cov.hyp, 'covRange')))
    likRange = property(lambda s: s._likRange,
                        lambda s, v: setattr(s, '_likRange', s._checked(v, s.lik.hyp, 'likRange')))


# --------------------------------------------------------------------------------------
def minimize(f, X, length, red=1.0):
    """Nonlinear conjugate gradients with an interpolating/extrapolating line search.

    f(X) -> (value, gradient).  length > 0: maximum number of line searches; length < 0:
    maximum number of function evaluations.  Returns (X, [f values], iterations).
    Constants as published: INT 0.1, EXT 3.0, MAX 20, RATIO 10, SIG 0.1, RHO SIG/2."""
    INT, EXT, MAXEV, RATIO, SIG = 0.1, 3.0, 20, 10.0, 0.1
    RHO = SIG / 2.0
    tiny = np.finfo(float).tiny
    by_evals = length < 0
    budget = abs(length)

    count = 0
    failed_before = False
    f0, df0 = f(X)
    history = [f0]
    count += 1 if by_evals else 0
    s = -df0
    d0 = -np.dot(s, s)
    x3 = red / (1.0 - d0)

    while count < budget:
        count += 0 if by_evals else 1
        Xbest, Fbest, dFbest = X, f0, df0
        M = min(MAXEV, budget - count) if by_evals else MAXEV

        # ---- extrapolation ----
        while True:
            x2, f2, d2 = 0.0, f0, d0
            f3, df3 = f0, df0
            ok = False
            while not ok and M > 0:
                try:
                    M -= 1
                    count += 1 if by_evals else 0
                    f3, df3 = f(X + x3 * s)
                    if np.isnan(f3) or np.isinf(f3) or np.any(np.isnan(df3) + np.isinf(df3)):
                        return None                       # the reference gives up here (bare `return`)
                    ok = True
                except Exception:
                    x3 = (x2 + x3) / 2.0                   # bisect and retry
            if f3 < Fbest:
                Xbest, Fbest, dFbest = X + x3 * s, f3, df3
            d3 = np.dot(df3, s)
            if d3 > SIG * d0 or f3 > f0 + x3 * RHO * d0 or M == 0:
                break
            x1, f1, d1 = x2, f2, d2
            x2, f2, d2 = x3, f3, d3
            A = 6.0 * (f1 - f2) + 3.0 * (d2 + d1) * (x2 - x1)
            B = 3.0 * (f2 - f1) - (2.0 * d1 + d2) * (x2 - x1)
            Z = B + np.sqrt(complex(B * B - A * d1 * (x2 - x1)))
            x3 = x1 - d1 * (x2 - x1) ** 2 / Z if Z != 0.0 else np.inf
            if (not np.isreal(x3)) or np.isnan(x3) or np.isinf(x3) or (x3 < 0):
                x3 = x2 * EXT
            elif x3 > x2 * EXT:
                x3 = x2 * EXT
            elif x3 < x2 + INT * (x2 - x1):
                x3 = x2 + INT * (x2 - x1)
            x3 = np.real(x3)

        # ---- interpolation ----
        x4 = f4 = d4 = None
        while (abs(d3) > -SIG * d0 or f3 > f0 + x3 * RHO * d0) and M > 0:
            if d3 > 0 or f3 > f0 + x3 * RHO * d0:
                x4, f4, d4 = x3, f3, d3
            else:
                x2, f2, d2 = x3, f3, d3
            if f4 > f0:
                x3 = x2 - (0.5 * d2 * (x4 - x2) ** 2) / (f4 - f2 - d2 * (x4 - x2))
            else:
                A = 6.0 * (f2 - f4) / (x4 - x2) + 3.0 * (d4 + d2)
                B = 3.0 * (f4 - f2) - (2.0 * d2 + d4) * (x4 - x2)
                x3 = x2 + (np.sqrt(B * B - A * d2 * (x4 - x2) ** 2) - B) / A if A != 0 else np.inf
            if np.isnan(x3) or np.isinf(x3):
                x3 = (x2 + x4) / 2.0
            x3 = max(min(x3, x4 - INT * (x4 - x2)), x2 + INT * (x4 - x2))
            f3, df3 = f(X + x3 * s)
            if f3 < Fbest:
                Xbest, Fbest, dFbest = X + x3 * s, f3, df3
            M -= 1
            count += 1 if by_evals else 0
            d3 = np.dot(df3, s)

        if abs(d3) < -SIG * d0 and f3 < f0 + x3 * RHO * d0:          # line search succeeded
            X = X + x3 * s
            f0 = f3
            history.append(f0)
            s = (np.dot(df3, df3) - np.dot(df0, df3)) / np.dot(df0, df0) * s - df3
            df0 = df3
            d3, d0 = d0, np.dot(df0, s)
            if d0 > 0:
                s = -df0
                d0 = -np.dot(s, s)
            x3 = x3 * min(RATIO, d3 / (d0 - tiny))
            failed_before = False
        else:
            X, f0, df0 = Xbest, Fbest, dFbest
            if failed_before or count > budget:
                break
            s = -df0
            d0 = -np.dot(s, s)
            x3 = 1.0 / (1.0 - d0)
            failed_before = True
    return X, history, count


def scg(f, x, niters=100, gradcheck=False, display=False, xtol=1e-6, ftol=1e-6):
    """Moller's scaled conjugate gradient.  f(x) -> (value, gradient).
    Returns (x, [f values], iterations)."""
    sigma0 = 1.0e-4
    fold, gradnew = f(x)
    fnow = fold
    gradold = gradnew.copy()
    d = -gradnew
    success = True
    nsuccess = 0
    beta, betamin, betamax = 1.0, 1.0e-15, 1.0e100
    nparams = len(x)
    flog = [fold]
    j = 1
    mu = kappa = theta = 0.0
    while j <= niters:
        if success:
            mu = np.dot(d, gradnew)
            if mu >= 0:
                d = -gradnew
                mu = np.dot(d, gradnew)
            kappa = np.dot(d, d)
            if kappa < np.finfo(float).eps:
                return x, flog, j
            sigma = sigma0 / np.sqrt(kappa)
            _, gplus = f(x + sigma * d)
            theta = np.dot(d, gplus - gradnew) / sigma
        delta = theta + beta * kappa
        if delta <= 0:
            delta = beta * kappa
            beta = beta - theta / kappa
        alpha = -mu / delta
        xnew = x + alpha * d
        fnew, gnew_candidate = f(xnew)
        Delta = 2 * (fnew - fold) / (alpha * mu)
        if Delta >= 0:
            success = True
            nsuccess += 1
            x = xnew
            fnow = fnew
        else:
            success = False
            fnow = fold
        flog.append(fnow)
        if success:
            if max(abs(alpha * d)) < xtol and abs(fnew - fold) < ftol:
                return x, flog, j
            fold = fnew
            gradold = gradnew
            gradnew = gnew_candidate
            if np.dot(gradnew, gradnew) == 0:
                return x, flog, j
        if Delta < 0.25:
            beta = min(4.0 * beta, betamax)
        if Delta > 0.75:
            beta = max(0.5 * beta, betamin)
        if nsuccess == nparams:
            d = -gradnew
            nsuccess = 0
        elif success:
            gamma = np.dot(gradold - gradnew, gradnew) / mu
            d = gamma * d - gradnew
        j += 1
    return x, flog, j


# --------------------------------------------------------------------------------------
class Optimizer(object):
    """Base class (Core/opt.py:35-89): hyper-parameter (un)flattening in the order
    mean + cov + lik and the objective callbacks."""
    _label = 'Optimizer'
    _failtext = 'optimizer'

    def __init__(self, model=None, searchConfig=None):
        self.model = model
        self.searchConfig = searchConfig
        self.trailsCounter = 0
        self.errorCounter = 0
        self.logger = logging.getLogger(__name__)

    def findMin(self, x, y, numIters):
        hyp0 = self._convert_to_array()
        return self._search(hyp0, numIters)

    def _nlml(self, hypInArray):
        self._apply_in_objects(hypInArray)
        nlZ, post = self.model.getPosterior(der=False)
        return nlZ

    def _dnlml(self, hypInArray):
        self._apply_in_objects(hypInArray)
        nlZ, dnlZ, post = self.model.getPosterior()
        return np.array(dnlZ.mean + dnlZ.cov + dnlZ.lik)

    def _nlzAnddnlz(self, hypInArray):
        self._apply_in_objects(hypInArray)
        nlZ, dnlZ, post = self.model.getPosterior()
        return nlZ, np.array(dnlZ.mean + dnlZ.cov + dnlZ.lik)

    def _convert_to_array(self):
        m = self.model
        return np.array(m.meanfunc.hyp + m.covfunc.hyp + m.likfunc.hyp)

    def _apply_in_objects(self, hypInArray):
        m = self.model
        Lm = len(m.meanfunc.hyp)
        Lc = len(m.covfunc.hyp)
        flat = hypInArray.tolist()
        m.meanfunc.hyp = flat[:Lm]
        m.covfunc.hyp = flat[Lm:(Lm + Lc)]
        m.likfunc.hyp = flat[(Lm + Lc):]

    # one optimisation run from `hyp`; returns (hyp, value)
    def _run_once(self, hyp, numIters, first):
        raise NotImplementedError

    # -- multi-GPU random restarts ------------------------------------------------------------------------------
    def _restart_devices(self):
        """CUDA ordinals the random restarts may run on (model.setDevices, else every GPU this process sees)."""
        devs = getattr(self.model, 'devices', None)
        if devs:
            return list(devs)
        try:
            from . import _lib
            return _lib.visible_devices()
        except Exception:
            return [0]

    def _worker_for(self, device, slot=0):
        """A private copy of the model (and of this optimizer) whose evaluations run on `device`: the restarts of
        Core/opt.py:301-327 are independent minimisations, so each GPU gets its own copy and its own libgpk handle
        (a device listed twice gets two copies: two evaluation streams share that GPU)."""
        if not hasattr(self, '_workers'):
            self._workers = {}
        key = (device, slot)
        w = self._workers.get(key)
        if w is None:
            workers, self._workers = self._workers, {}          # do not copy the copies
            try:
                m = deepcopy(self.model)
            finally:
                self._workers = workers
            m.devices = [device]
            m.inffunc.devices = [device]
            m.inffunc.shard = False
            m.inffunc._engine = None
            m.optimizer.searchConfig = None
            w = self._workers[key] = m
        return w

    def _run_trials(self, jobs, numIters):
        """jobs: [(hyp, first)] -> [(hyp, value) | Exception] in the same order.  One GPU: in order on the model itself;
        several: concurrently, one model copy per GPU (libgpk calls release the GIL), results kept in job order so that
        the bookkeeping below sees exactly the sequence a serial run would."""
        devs = self._restart_devices()
        out = [None] * len(jobs)
        if len(devs) < 2 or len(jobs) < 2:
            for i, (hyp, first) in enumerate(jobs):
                try:
                    h, v = self._run_once(hyp, numIters, first)
                    out[i] = (deepcopy(h), v)
                except Exception as e:                       # a failed trial, as Core/opt.py:317-318
                    out[i] = e
            self.devices_used = sorted(set(getattr(self, 'devices_used', [])) | set(devs[:1]))
            return out
        import queue
        import threading
        todo = queue.Queue()
        for i, job in enumerate(jobs):
            todo.put((i, job))
        used = set()

        copies = []
        for slot, d in enumerate(devs[:len(jobs)]):          # model copies are made here, on the calling thread
            try:
                copies.append(self._worker_for(d, devs[:slot].count(d)))
            except Exception as e:
                copies.append(e)

        def work(slot, dev):
            m = copies[slot]
            while True:
                try:
                    i, (hyp, first) = todo.get_nowait()
                except queue.Empty:
                    return
                if isinstance(m, Exception):
                    out[i] = m
                    continue
                try:
                    h, v = m.optimizer._run_once(np.array(hyp, dtype=float), numIters, first)
                    out[i] = (deepcopy(h), v)
                    used.add(dev)
                except Exception as e:
                    out[i] = e
        threads = [threading.Thread(target=work, args=(slot, d)) for slot, d in enumerate(devs[:len(jobs)])]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        self.devices_used = sorted(set(getattr(self, 'devices_used', [])) | used)
        return out

    def _search(self, hyp0, numIters):
        """First run from the current hyper-parameters, then (with a searchConfig) uniform random
        restarts until num_restarts trials or min_threshold is reached; exceptions inside a trial
        count as failed trials and more than num_restarts/2 failures abort (Core/opt.py:100-149 etc.).
        The trials are independent: with several GPUs they run concurrently (one per GPU), drawn and
        accounted for in the serial order."""
        conf = self.searchConfig
        optimalHyp = funcValue = None
        if not conf:
            r = self._run_trials([(hyp0, True)], numIters)[0]
            self.trailsCounter += 1
            if isinstance(r, Exception):
                self.errorCounter += 1
                raise Exception("Can not learn hyperparamters using %s." % self._failtext)
            return r
        ranges = conf.meanRange + conf.covRange + conf.likRange
        if not (conf.num_restarts or conf.min_threshold):
            raise Exception('Specify at least one of the stop conditions')
        ndev = max(1, len(self._restart_devices()))
        hyp = np.array(hyp0, dtype=float)
        first_pending = True
        while True:
            # one wave: as many trials as can still be needed (all remaining restarts, or one per GPU when only
            # min_threshold bounds the search); random starts are drawn in trial order exactly like the serial loop
            if conf.num_restarts:
                left = int(conf.num_restarts) - self.trailsCounter
                wave = max(1, left)
            else:
                wave = ndev
            jobs = []
            if first_pending:
                jobs.append((np.array(hyp0, dtype=float), True))
            while len(jobs) < wave:
                for i in range(hyp.shape[0]):
                    hyp[i] = np.random.uniform(low=ranges[i][0], high=ranges[i][1])
                jobs.append((hyp.copy(), False))
            results = self._run_trials(jobs, numIters)
            for (job_hyp, first), r in zip(jobs, results):
                self.trailsCounter += 1
                if isinstance(r, Exception):
                    self.errorCounter += 1
                else:
                    h, v = r
                    if funcValue is None or v < funcValue:
                        funcValue, optimalHyp = v, h
                if first:
                    first_pending = False
                    continue                                  # the stop conditions are checked after restarts only
                if conf.num_restarts and self.errorCounter > conf.num_restarts / 2:
                    self.logger.warning("[%s] %d out of %d trails failed during optimization", self._label,
                                        self.errorCounter, self.trailsCounter)
                    raise Exception("Over half of the trails failed for %s" % self._failtext)
                done = (conf.num_restarts and self.trailsCounter > conf.num_restarts - 1) or \
                       (conf.min_threshold and funcValue is not None and funcValue <= conf.min_threshold)
                if done:
                    self.logger.warning("[%s] %d out of %d trails failed during optimization", self._label,
                                        self.errorCounter, self.trailsCounter)
                    return optimalHyp, funcValue


class Minimize(Optimizer):
    """minimize by Carl Rasmussen - the default optimizer (Core/opt.py:273-328)."""
    _label, _failtext = 'Minimize', 'minimize'

    def __init__(self, model, searchConfig=None):
        super(Minimize, self).__init__(model, searchConfig)

    def findMin(self, x, y, numIters=200):
        return self._search(self._convert_to_array(), numIters)

    def _run_once(self, hyp, numIters, first):
        out = minimize(self._nlzAnddnlz, np.array(hyp, dtype=float), length=numIters)
        self.logger.warning("Number of line searches %g", out[2])
        return out[0], out[1][-1]


class SCG(Optimizer):
    """Scaled conjugate gradient (Core/opt.py:332-383)."""
    _label, _failtext = 'SCG', 'Scaled conjugate gradient'

    def __init__(self, model, searchConfig=None):
        super(SCG, self).__init__(model, searchConfig)

    def findMin(self, x, y, numIters=100):
        return self._search(self._convert_to_array(), numIters)

    def _run_once(self, hyp, numIters, first):
        out = scg(self._nlzAnddnlz, np.array(hyp, dtype=float), niters=numIters if first else 100)
        return out[0], out[1][-1]


class CG(Optimizer):
    """Conjugate gradient through scipy.optimize.fmin_cg (Core/opt.py:152-207)."""
    _label, _failtext = 'CG', 'conjugate gradient'

    def __init__(self, model, searchConfig=None):
        super(CG, self).__init__(model, searchConfig)

    def findMin(self, x, y, numIters=100):
        return self._search(self._convert_to_array(), numIters)

    def _run_once(self, hyp, numIters, first):
        out = _cg(self._nlml, np.array(hyp, dtype=float), self._dnlml, maxiter=numIters, disp=False,
                  full_output=True)
        if out[4] == 1:
            self.logger.warning("Maximum number of iterations exceeded.")
        elif out[4] == 2:
            self.logger.warning("Gradient and/or function calls not changing.")
        return out[0], out[1]


class BFGS(Optimizer):
    """Quasi-Newton BFGS through scipy.optimize.fmin_bfgs (Core/opt.py:211-269)."""
    _label, _failtext = 'BFGS', 'BFGS'

    def __init__(self, model, searchConfig=None):
        super(BFGS, self).__init__(model, searchConfig)

    def findMin(self, x, y, numIters=100):
        return self._search(self._convert_to_array(), numIters)

    def _run_once(self, hyp, numIters, first):
        out = _bfgs(self._nlml, np.array(hyp, dtype=float), self._dnlml, maxiter=numIters, disp=False,
                    full_output=True)
        if out[6] == 1:
            self.logger.warning("Maximum number of iterations exceeded.")
        elif out[6] == 2:
            self.logger.warning("Gradient and/or function calls not changing.")
        return out[0], out[1]


class Simplex(Optimizer):
    """Nelder-Mead downhill simplex through scipy.optimize.fmin (Core/opt.py:92-149)."""
    _label, _failtext = 'Simplex', 'Nelder-Mead'

    def __init__(self, model, searchConfig=None):
        super(Simplex, self).__init__(model, searchConfig)

    def findMin(self, x, y, numIters=100):
        return self._search(self._convert_to_array(), numIters)

    def _run_once(self, hyp, numIters, first):
        out = _simplex(self._nlml, np.array(hyp, dtype=float), maxiter=numIters, disp=False, full_output=True)
        if out[4] == 1:
            self.logger.warning("Maximum number of function evaluations made")
        elif out[4] == 2:
            self.logger.warning("Maximum number of iterations exceeded.")
        return out[0], out[1]
