"""Covariance functions - host side of the plugin contract of pyGPs.Core.cov.

Same class names, constructor arguments, `.hyp` lists of LOG hyper-parameters,
`getCovMatrix(x, z, mode)` / `getDerMatrix(x, z, mode, der)` signatures, modes
('train' | 'cross' | 'self_test') and error behaviour as the reference
(/root/reference/pyGPs/Core/cov.py:61-226 base, :786-828 RBF, :872-938 RBFard,
:1078-1182 Matern, :332-390 FITCOfKernel, :230-328 composites).  All matrix
arithmetic runs in libgpk.so (csrc/kbuild.cu); nothing here computes a distance.

Only the kernels on the accelerated path are provided natively.  Composites
(+, *, scalar *) combine device-built matrices element-wise on the host - they
are the "next tier" of SURVEY section 8(f3).
"""
import logging

import numpy as np

from . import _lib


class Kernel(object):
    """Base class: defines the interface and the operator overloads (Core/cov.py:61-202)."""

    def __init__(self):
        self.hyp = []
        self.para = []
        self.logger = logging.getLogger(__name__)

    def __repr__(self):
        return (str(type(self)) + ': to get the kernel matrix or kernel derviatives use: \n'
                'model.covfunc.getCovMatrix()\nmodel.covfunc.getDerMatrix()')

    def getCovMatrix(self, x=None, z=None, mode=None):
        pass

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        pass

    # -- argument checks: same messages as Core/cov.py:114-152 ----------------
    def checkInputGetCovMatrix(self, x, z, mode):
        if mode is None:
            raise Exception("Specify the mode: 'train' or 'cross'")
        if x is None and z is None:
            raise Exception("Specify at least one: training input (x) or test input (z) or both.")
        if mode == 'cross':
            if x is None or z is None:
                raise Exception("Specify both: training input (x) and test input (z) for cross covariance.")

    def checkInputGetDerMatrix(self, x, z, mode, der):
        self.checkInputGetCovMatrix(x, z, mode)
        if der is None:
            raise Exception("Specify the index of parameters of the derivatives.")

    # -- operators -------------------------------------------------------------
    def __add__(self, cov):
        return SumOfKernel(self, cov)

    def __mul__(self, other):
        if isinstance(other, (int, float)):
            return ScaleOfKernel(self, other)
        if isinstance(other, Kernel):
            return ProductOfKernel(self, other)
        logging.getLogger(__name__).error("only numbers and Kernels are supported operand types for *")

    __rmul__ = __mul__

    def fitc(self, inducingInput):
        """Wrap for the FITC approximation (Core/cov.py:194-202)."""
        return FITCOfKernel(self, inducingInput)

    # -- device description ------------------------------------------------------
    def _device_spec(self):
        """(kind, matern_d, hyp) when libgpk has a native fused path for this kernel, else None."""
        return None


class _NativeKernel(Kernel):
    """RBF / RBFard / Matern: matrices come from gpk_cov_matrix."""
    _kind = None

    def _matern_d(self):
        return 3

    def _device_spec(self):
        return (self._kind, self._matern_d(), [float(v) for v in self.hyp])

    def _nder(self):
        return len(self.hyp)

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        kind, d, hyp = self._device_spec()
        if mode not in ('train', 'cross', 'self_test'):
            return None                      # the reference falls through and returns nothing useful
        if mode == 'train' and x is None:
            raise Exception("Specify training input (x) for mode 'train'")
        if mode == 'self_test' and z is None:
            raise Exception("Specify test input (z) for mode 'self_test'")
        return _lib.shared_engine().cov_matrix(kind, d, hyp, x, z, mode, -1)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        kind, d, hyp = self._device_spec()
        if not (isinstance(der, (int, np.integer)) and 0 <= der < self._nder()):
            return self._bad_der(x, z, mode, der)
        return _lib.shared_engine().cov_matrix(kind, d, hyp, x, z, mode, int(der))

    def _bad_der(self, x, z, mode, der):
        raise Exception("Calling for a derivative in %s that does not exist" % type(self).__name__)


class RBF(_NativeKernel):
    """Squared exponential, isotropic.  hyp = [log_ell, log_sigma]  (Core/cov.py:786-828)."""
    _kind = _lib.COV_RBF

    def __init__(self, log_ell=0., log_sigma=0.):
        self.hyp = [log_ell, log_sigma]
        self.para = []


class RBFard(_NativeKernel):
    """Squared exponential with ARD.  hyp = log_ell_list + [log_sigma]  (Core/cov.py:872-938)."""
    _kind = _lib.COV_RBFARD

    def __init__(self, D=None, log_ell_list=None, log_sigma=0.):
        if log_ell_list is None:
            self.hyp = [0. for i in range(D)] + [log_sigma]
        else:
            self.hyp = log_ell_list + [log_sigma]
        self.para = []

    def _bad_der(self, x, z, mode, der):
        raise Exception("Wrong derivative index in RDFard")


class Matern(_NativeKernel):
    """Matern, nu = d/2, d in {1,3,5,7}.  hyp = [log_ell, log_sigma], para = [d]  (Core/cov.py:1078-1182).

    `getDerMatrix(der=0)` returns the mathematically correct length-scale derivative; the
    reference's (Core/cov.py:1173-1177) reuses K as the distance and is wrong."""
    _kind = _lib.COV_MATERN

    def __init__(self, log_ell=0., d=3, log_sigma=0.):
        self.hyp = [log_ell, log_sigma]
        self.para = [d]

    def _matern_d(self):
        d = self.para[0]
        if np.abs(d - np.round(d)) < 1e-8:
            d = int(round(d))
        d = int(d)
        if d not in (1, 3, 5, 7):
            logging.getLogger(__name__).warning("d is neither 1,3,5 nor 7. We set it to d=3. ")
            d = 3
        return d

    def _nder(self):
        return 3

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        if der == 2:                         # d is not learned (Core/cov.py:1178-1179)
            kind, d, hyp = self._device_spec()
            return np.zeros_like(_lib.shared_engine().cov_matrix(kind, d, hyp, x, z, mode, -1))
        if der not in (0, 1):
            raise Exception("Wrong derivative value in Matern")
        kind, d, hyp = self._device_spec()
        return _lib.shared_engine().cov_matrix(kind, d, hyp, x, z, mode, int(der))


class FITCOfKernel(Kernel):
    """Covariances against inducing inputs for FITC (Core/cov.py:332-390).
    mode='train' returns the triple (diag K (n,1), Kuu (M,M), Ku (M,n))."""

    def __init__(self, cov, inducingInput):
        self.inducingInput = inducingInput
        self.covfunc = cov
        self._hyp = cov.hyp
        self.para = []

    def _getHyp(self):
        return self._hyp

    def _setHyp(self, hyp):
        self._hyp = hyp
        self.covfunc.hyp = hyp
    hyp = property(_getHyp, _setHyp)

    def _check_dim(self, x):
        if x is not None and self.inducingInput.shape[1] != x.shape[1]:
            raise Exception('Dimensionality of inducing inputs must match training inputs')

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        xu = self.inducingInput
        self._check_dim(x)
        if mode == 'self_test':
            return self.covfunc.getCovMatrix(z=z, mode='self_test')
        if mode == 'train':
            return (self.covfunc.getCovMatrix(z=x, mode='self_test'),
                    self.covfunc.getCovMatrix(x=xu, mode='train'),
                    self.covfunc.getCovMatrix(x=xu, z=x, mode='cross'))
        if mode == 'cross':
            return self.covfunc.getCovMatrix(x=xu, z=z, mode='cross')

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        xu = self.inducingInput
        self._check_dim(x)
        if mode == 'self_test':
            return self.covfunc.getDerMatrix(z=z, mode='self_test', der=der)
        if mode == 'train':
            return (self.covfunc.getDerMatrix(z=x, mode='self_test', der=der),
                    self.covfunc.getDerMatrix(x=xu, mode='train', der=der),
                    self.covfunc.getDerMatrix(x=xu, z=x, mode='cross', der=der))
        if mode == 'cross':
            return self.covfunc.getDerMatrix(x=xu, z=z, mode='cross', der=der)

    def _device_spec(self):
        return self.covfunc._device_spec()


class _Pair(Kernel):
    def __init__(self, cov1, cov2):
        self.cov1 = cov1
        self.cov2 = cov2
        self._hyp = cov1.hyp + cov2.hyp
        self.para = []

    def _setHyp(self, hyp):
        assert len(hyp) == len(self._hyp)
        len1 = len(self.cov1.hyp)
        self._hyp = hyp
        self.cov1.hyp = self._hyp[:len1]
        self.cov2.hyp = self._hyp[len1:]

    def _getHyp(self):
        return self._hyp
    hyp = property(_getHyp, _setHyp)


class SumOfKernel(_Pair):
    """k1 + k2 (Core/cov.py:265-295)."""

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        return self.cov1.getCovMatrix(x, z, mode) + self.cov2.getCovMatrix(x, z, mode)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        n1 = len(self.cov1.hyp)
        if der < n1:
            return self.cov1.getDerMatrix(x, z, mode, der)
        if der < len(self.hyp):
            return self.cov2.getDerMatrix(x, z, mode, der - n1)
        raise Exception("Error: der out of range for covSum")


class ProductOfKernel(_Pair):
    """k1 * k2 (Core/cov.py:230-261)."""

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        return self.cov1.getCovMatrix(x, z, mode) * self.cov2.getCovMatrix(x, z, mode)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        n1 = len(self.cov1.hyp)
        if der < n1:
            return self.cov1.getDerMatrix(x, z, mode, der) * self.cov2.getCovMatrix(x, z, mode)
        if der < len(self.hyp):
            return self.cov2.getDerMatrix(x, z, mode, der - n1) * self.cov1.getCovMatrix(x, z, mode)
        raise Exception("Error: der out of range for covProduct")


class ScaleOfKernel(Kernel):
    """scalar * k.  As in the reference (Core/cov.py:299-328) the scalar is stored raw in
    hyp[0] and enters as exp(hyp[0]); der=0 returns 2*exp(hyp[0])*K (reference behaviour)."""

    def __init__(self, cov, scalar):
        self.cov = cov
        if cov.hyp:
            self._hyp = [scalar] + cov.hyp
        else:
            self._hyp = [scalar]
        self.para = []

    def _setHyp(self, hyp):
        assert len(hyp) == len(self._hyp)
        self._hyp = hyp
        self.cov.hyp = self._hyp[1:]

    def _getHyp(self):
        return self._hyp
    hyp = property(_getHyp, _setHyp)

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        return np.exp(self.hyp[0]) * self.cov.getCovMatrix(x, z, mode)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        sf2 = np.exp(self.hyp[0])
        if der == 0:
            return 2. * sf2 * self.cov.getCovMatrix(x, z, mode)
        return sf2 * self.cov.getDerMatrix(x, z, mode, der - 1)
