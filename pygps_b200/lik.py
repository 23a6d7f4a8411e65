"""Likelihood functions (host side).

Only lik.Gauss is on the accelerated path (inf.Exact / inf.FITC_Exact refuse anything
else, /root/reference/pyGPs/Core/inf.py:354,400).  It supplies the noise variance
sn2 = exp(2*hyp[0]) and the O(n) predictive moments ymu = fmu, ys2 = fs2 + sn2 and
log-probabilities (Core/lik.py:123-198); those stay on the host.
"""
import logging

import numpy as np


class Likelihood(object):
    """Base class (Core/lik.py:44-119)."""

    def __init__(self):
        self.hyp = []
        self.logger = logging.getLogger(__name__)

    def evaluate(self, y=None, mu=None, s2=None, inffunc=None, der=None, nargout=1):
        pass


class Gauss(Likelihood):
    """Gaussian likelihood for regression.  hyp = [log_sigma] (default log 0.1)."""

    def __init__(self, log_sigma=np.log(0.1)):
        self.hyp = [log_sigma]

    def _ep_moments(self, y, mu, s2, sn2, der, nargout):
        # log partition function of N(y|f,sn2) N(f|mu,s2) and its mu-derivatives (Core/lik.py:158-172)
        if der is not None:
            return ((y - mu) ** 2 / (sn2 + s2) - 1) / (1 + s2 / sn2)
        lZ = -(y - mu) ** 2 / (sn2 + s2) / 2. - np.log(2 * np.pi * (sn2 + s2)) / 2.
        if nargout == 1:
            return lZ
        dlZ = (y - mu) / (sn2 + s2)
        if nargout == 2:
            return lZ, dlZ
        return lZ, dlZ, -1 / (sn2 + s2)

    def evaluate(self, y=None, mu=None, s2=None, inffunc=None, der=None, nargout=1):
        sn2 = np.exp(2. * self.hyp[0])
        if inffunc is None:                                  # prediction mode (Core/lik.py:137-157)
            if y is None:
                y = np.zeros_like(mu)
            if s2 is not None and np.linalg.norm(s2) > 0:
                lp = self._ep_moments(y, mu, s2, sn2, None, 1)
            else:
                lp = -(y - mu) ** 2 / sn2 / 2 - np.log(2. * np.pi * sn2) / 2.
                s2 = np.zeros_like(s2)
            if nargout == 1:
                return lp
            if nargout == 2:
                return lp, mu
            return lp, mu, s2 + sn2
        name = type(inffunc).__name__
        if name in ('EP', 'FITC_EP'):
            return self._ep_moments(y, mu, s2, sn2, der, nargout)
        raise Exception('lik.Gauss: inference mode %s is outside the accelerated path' % name)


class Erf(Likelihood):
    """Probit (cumulative Gaussian) likelihood for binary classification; no hyper-parameters.

    Mirror of pyGPs.Core.lik.Erf (/root/reference/pyGPs/Core/lik.py:236-366).  During inference (inf.EP) the
    EP moments are evaluated on the GPU inside gpk_ep_eval; this class supplies the O(ns) prediction mode
    (lp, ymu = 2p-1, ys2 = 4p(1-p), Core/lik.py:253-271) and the same safe special functions on the host."""

    def __init__(self):
        self.hyp = []

    # -- special functions (thresholds as in the reference) ---------------------------------------
    def logphi(self, z, p):
        z = np.asarray(z, dtype=float)
        lp = np.zeros_like(z)
        zmin, zmax = -6.2, -5.5
        ok = z > zmax
        nok = ~ok
        ip = nok & ~(z < zmin)
        lam = 1. / (1. + np.exp(25. * (0.5 - (z[ip] - zmin) / (zmax - zmin))))
        lp[ok] = np.log(p[ok])
        lp[nok] = -np.log(np.pi) / 2. - z[nok] ** 2 / 2. - np.log(np.sqrt(z[nok] ** 2 / 2. + 2.) - z[nok] / np.sqrt(2.))
        lp[ip] = (1 - lam) * lp[ip] + lam * np.log(p[ip])
        return lp

    def cumGauss(self, y=None, f=None, nargout=1):
        from scipy.special import erf
        yf = y * f if y is not None else f
        p = (1. + erf(yf / np.sqrt(2.))) / 2.
        if nargout > 1:
            return p, self.logphi(yf, p)
        return p

    def gauOverCumGauss(self, f, p):
        f = np.asarray(f, dtype=float)
        n_p = np.zeros_like(f)
        ok = f > -5
        n_p[ok] = (np.exp(-f[ok] ** 2 / 2) / np.sqrt(2 * np.pi)) / p[ok]
        bd = f < -6
        n_p[bd] = np.sqrt(f[bd] ** 2 / 4 + 1) - f[bd] / 2
        it = ~ok & ~bd
        lam = -5. - f[it]
        n_p[it] = (1 - lam) * (np.exp(-f[it] ** 2 / 2) / np.sqrt(2 * np.pi)) / p[it] + \
            lam * (np.sqrt(f[it] ** 2 / 4 + 1) - f[it] / 2)
        return n_p

    def _ep(self, y, mu, s2, nargout):
        z = mu / np.sqrt(1 + s2)
        junk, lZ = self.cumGauss(y, z, 2)
        if nargout == 1:
            return lZ
        z = z * y
        n_p = self.gauOverCumGauss(z, np.exp(lZ))
        dlZ = y * n_p / np.sqrt(1. + s2)
        if nargout == 2:
            return lZ, dlZ
        return lZ, dlZ, -n_p * (z + n_p) / (1. + s2)

    def evaluate(self, y=None, mu=None, s2=None, inffunc=None, der=None, nargout=1):
        if y is not None:
            y = np.sign(y)
            y[y == 0] = 1
        else:
            y = 1
        if inffunc is None:                                   # prediction mode
            y = y * np.ones_like(mu)
            if s2 is not None and np.linalg.norm(s2) > 0:
                lp = self._ep(y, mu, s2, 1)
                p = np.exp(lp)
            else:
                p, lp = self.cumGauss(y, mu, 2)
            if nargout == 1:
                return lp
            if nargout == 2:
                return lp, 2 * p - 1
            return lp, 2 * p - 1, 4 * p * (1 - p)
        if type(inffunc).__name__ in ('EP', 'FITC_EP'):
            if der is not None:
                return []                                     # no hyper-parameters
            return self._ep(y * np.ones_like(mu), mu, s2, nargout)
        raise Exception('lik.Erf: inference mode %s is outside the accelerated path' % type(inffunc).__name__)
