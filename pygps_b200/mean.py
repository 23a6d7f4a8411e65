"""Prior mean functions (host side; O(nD) vectors feeding the GPU path).

Mirror of pyGPs.Core.mean (/root/reference/pyGPs/Core/mean.py): Zero :279, One :297,
Const :315, Linear :339 and the +, *, scalar*, ** composites (:140-276).  They stay on
the host: the reference evaluates them once per call (Core/inf.py:358) and they are
vectors, not matrices.
"""
import logging

import numpy as np


class Mean(object):
    """Base class with the operator overloads of Core/mean.py:47-136."""

    def __init__(self):
        self.hyp = []
        self.para = []
        self.logger = logging.getLogger(__name__)

    def __repr__(self):
        return (str(type(self)) + ': to get the mean vector or mean derviatives use: \n'
                'model.meanfunc.getMean()\nmodel.meanfunc.getDerMatrix()')

    def __add__(self, mean):
        return SumOfMean(self, mean)

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return ScaleOfMean(self, other)
        if isinstance(other, Mean):
            return ProductOfMean(self, other)
        logging.getLogger(__name__).error("only numbers and Means are allowed for *")

    __rmul__ = __mul__

    def __pow__(self, number):
        if isinstance(number, int) and number > 0:
            return PowerOfMean(self, number)
        logging.getLogger(__name__).error("only non-zero integers are supported for **")

    def getMean(self, x=None):
        pass

    def getDerMatrix(self, x=None, der=None):
        pass


def _col(n, v):
    return np.full((n, 1), float(v))


class Zero(Mean):
    """m(x) = 0."""

    def __init__(self):
        self.hyp = []
        self.name = '0'

    def getMean(self, x=None):
        return _col(x.shape[0], 0.)

    def getDerMatrix(self, x=None, der=None):
        return _col(x.shape[0], 0.)


class One(Mean):
    """m(x) = 1."""

    def __init__(self):
        self.hyp = []
        self.name = '1'

    def getMean(self, x=None):
        return _col(x.shape[0], 1.)

    def getDerMatrix(self, x=None, der=None):
        return _col(x.shape[0], 0.)


class Const(Mean):
    """m(x) = c.  hyp = [c]."""

    def __init__(self, c=5.):
        self.hyp = [c]

    def getMean(self, x=None):
        return self.hyp[0] * np.ones((x.shape[0], 1))

    def getDerMatrix(self, x=None, der=None):
        return _col(x.shape[0], 1. if der == 0 else 0.)


class Linear(Mean):
    """m(x) = x . a.  hyp = alpha_list (default 0.5 per dimension)."""

    def __init__(self, D=None, alpha_list=None):
        if alpha_list is None:
            self.hyp = [0.5] if D is None else [0.5 for i in range(D)]
        else:
            self.hyp = alpha_list

    def getMean(self, x=None):
        a = np.array(self.hyp, dtype=float).reshape(-1, 1)
        return np.dot(x, a)

    def getDerMatrix(self, x=None, der=None):
        n, D = x.shape
        if isinstance(der, (int, np.integer)) and der < D:
            return np.reshape(x[:, der], (n, 1))
        return _col(n, 0.)


class _PairOfMean(Mean):
    def __init__(self, mean1, mean2):
        self.mean1 = mean1
        self.mean2 = mean2
        self._hyp = list(mean1.hyp) + list(mean2.hyp)

    def _setHyp(self, hyp):
        assert len(hyp) == len(self._hyp)
        k = len(self.mean1.hyp)
        self._hyp = hyp
        self.mean1.hyp = self._hyp[:k]
        self.mean2.hyp = self._hyp[k:]

    def _getHyp(self):
        return self._hyp
    hyp = property(_getHyp, _setHyp)


class ProductOfMean(_PairOfMean):
    def getMean(self, x=None):
        return self.mean1.getMean(x) * self.mean2.getMean(x)

    def getDerMatrix(self, x=None, der=None):
        k = len(self.mean1.hyp)
        if der < k:
            return self.mean1.getDerMatrix(x, der) * self.mean2.getMean(x)
        if der < len(self.hyp):
            return self.mean2.getDerMatrix(x, der - k) * self.mean1.getMean(x)
        raise Exception("Error: der out of range for meanProduct")


class SumOfMean(_PairOfMean):
    def getMean(self, x=None):
        return self.mean1.getMean(x) + self.mean2.getMean(x)

    def getDerMatrix(self, x=None, der=None):
        k = len(self.mean1.hyp)
        if der < k:
            return self.mean1.getDerMatrix(x, der)
        if der < len(self.hyp):
            return self.mean2.getDerMatrix(x, der - k)
        raise Exception("Error: der out of range for meanSum")


class _WrapOfMean(Mean):
    def __init__(self, mean, first):
        self.mean = mean
        self._hyp = [first] + list(mean.hyp)

    def _setHyp(self, hyp):
        assert len(hyp) == len(self._hyp)
        self._hyp = hyp
        self.mean.hyp = self._hyp[1:]

    def _getHyp(self):
        return self._hyp
    hyp = property(_getHyp, _setHyp)


class ScaleOfMean(_WrapOfMean):
    def getMean(self, x=None):
        return self.hyp[0] * self.mean.getMean(x)

    def getDerMatrix(self, x=None, der=None):
        if der == 0:
            return self.mean.getMean(x)
        return self.hyp[0] * self.mean.getDerMatrix(x, der - 1)


class PowerOfMean(_WrapOfMean):
    def _d(self):
        return max(np.abs(np.floor(self.hyp[0])), 1)

    def getMean(self, x=None):
        return self.mean.getMean(x) ** self._d()

    def getDerMatrix(self, x=None, der=None):
        d = self._d()
        if der == 0:
            a = self.mean.getMean(x)
            return a ** d * np.log(a)
        return d * self.mean.getMean(x) ** (d - 1) * self.mean.getDerMatrix(x, der - 1)
