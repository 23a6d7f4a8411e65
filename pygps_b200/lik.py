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
