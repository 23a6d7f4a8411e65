"""Covariance functions - host side of the plugin contract of pyGPs.Core.cov.

Same class names, constructor arguments, `.hyp` lists of LOG hyper-parameters,
`getCovMatrix(x, z, mode)` / `getDerMatrix(x, z, mode, der)` signatures, modes
('train' | 'cross' | 'self_test') and error behaviour as the reference
(/root/reference/pyGPs/Core/cov.py:61-226 base, :786-828 RBF, :872-938 RBFard,
:1078-1182 Matern, :332-390 FITCOfKernel, :230-328 composites).  All matrix
arithmetic runs in libgpk.so (csrc/kbuild.cu); nothing here computes a distance.

RBF / RBFard / Matern have a dedicated fused build (gpk_cov_matrix).  Every other kernel here and every composite
(+, *, scalar *; Core/cov.py:230-328) is a PROGRAM evaluated on the device (csrc/covprog.cu): the expression tree in
post-order, one pass over the pair's coordinates, all components and all hyper-parameter derivatives at once -
`_device_prog()` below emits it.  cov.Pre (precomputed matrices, :1429-1455) uploads its training matrix once.
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

    _op = None                 # GPK_OP_* of a leaf kernel

    def _para(self):
        return 0.0

    def _emit(self, nodes, h0):
        """Append this kernel's nodes (post-order) to `nodes`; return the index of its root.  h0 = index of this
        kernel's first hyper-parameter in the composite's flat list.  Leaves override nothing but _op / _para."""
        if self._op is None:
            raise NotImplementedError("%s has no device implementation" % type(self).__name__)
        nodes.append((self._op, -1, -1, h0, float(self._para())))
        return len(nodes) - 1

    def _device_prog(self):
        """(nodes, hyp): the covariance program of this kernel (csrc/covprog.cu), or None if some component has no
        device implementation.  nodes: list of (op, a, b, hyp0, para)."""
        nodes = []
        try:
            self._emit(nodes, 0)
        except NotImplementedError:
            return None
        if len(nodes) > 32 or len(self.hyp) > 96:
            return None
        return nodes, [float(v) for v in self.hyp]

    def _pre_leaves(self):
        return []

    def _prog_matrix(self, x, z, mode, der):
        """getCovMatrix / getDerMatrix through the device program."""
        prog = self._device_prog()
        if prog is None:
            raise Exception("%s: no device implementation of this kernel" % type(self).__name__)
        nodes, hyp = prog
        if mode not in ('train', 'cross', 'self_test'):
            return None
        if mode == 'train' and x is None:
            raise Exception("Specify training input (x) for mode 'train'")
        if mode == 'self_test' and z is None:
            raise Exception("Specify test input (z) for mode 'self_test'")
        return _lib.shared_engine().cov_matrix_prog(nodes, hyp, x, z, mode, -1 if der is None else int(der))


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
    _op = _lib.OP_RBF

    def __init__(self, log_ell=0., log_sigma=0.):
        self.hyp = [log_ell, log_sigma]
        self.para = []


class RBFard(_NativeKernel):
    """Squared exponential with ARD.  hyp = log_ell_list + [log_sigma]  (Core/cov.py:872-938)."""
    _kind = _lib.COV_RBFARD
    _op = _lib.OP_RBFARD

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
    _op = _lib.OP_MATERN

    def __init__(self, log_ell=0., d=3, log_sigma=0.):
        self.hyp = [log_ell, log_sigma]
        self.para = [d]

    def _para(self):
        return self._matern_d()

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


class _ProgKernel(Kernel):
    """Leaf kernels evaluated by the device program (csrc/covprog.cu)."""
    _nder_extra = 0            # trailing non-learned parameters the reference accepts as `der` and answers with zeros

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        self._check_inputs(x, z)
        return self._prog_matrix(x, z, mode, None)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        self._check_inputs(x, z)
        if isinstance(der, (int, np.integer)) and len(self.hyp) <= der < len(self.hyp) + self._nder_extra:
            return np.zeros_like(self._prog_matrix(x, z, mode, None))
        if not (isinstance(der, (int, np.integer)) and 0 <= der < len(self.hyp)):
            raise Exception("Wrong derivative index in %s" % type(self).__name__)
        return self._prog_matrix(x, z, mode, der)

    def _check_inputs(self, x, z):
        pass


class RBFunit(_ProgKernel):
    """Squared exponential with unit magnitude.  hyp = [log_ell]  (Core/cov.py:832-868)."""
    _op = _lib.OP_RBFUNIT

    def __init__(self, log_ell=0.):
        self.hyp = [log_ell]
        self.para = []


class RQ(_ProgKernel):
    """Rational quadratic, isotropic.  hyp = [log_ell, log_sigma, log_alpha]  (Core/cov.py:1304-1352)."""
    _op = _lib.OP_RQ

    def __init__(self, log_ell=0., log_sigma=0., log_alpha=0.):
        self.hyp = [log_ell, log_sigma, log_alpha]
        self.para = []


class RQard(_ProgKernel):
    """Rational quadratic with ARD.  hyp = log_ell_list + [log_sigma, log_alpha]  (Core/cov.py:1356-1425).
    The length-scale derivatives are the true ones: the reference's (:1413-1418) take the distance between two
    (1,n) ROW vectors and are identically zero in train mode."""
    _op = _lib.OP_RQARD

    def __init__(self, D=None, log_ell_list=None, log_sigma=0., log_alpha=0.):
        if log_ell_list is None:
            self.hyp = [0. for i in range(D)] + [log_sigma, log_alpha]
        else:
            self.hyp = log_ell_list + [log_sigma, log_alpha]
        self.para = []


class Periodic(_ProgKernel):
    """Smooth periodic kernel for 1-d inputs.  hyp = [log_ell, log_p, log_sigma]  (Core/cov.py:1186-1250)."""
    _op = _lib.OP_PERIODIC

    def __init__(self, log_ell=0., log_p=0., log_sigma=0.):
        self.hyp = [log_ell, log_p, log_sigma]
        self.para = []

    def _check_inputs(self, x, z):
        if x is not None:
            assert x.shape[1] == 1, 'periodic covariance can only be used for 1d data'
        if z is not None:
            assert z.shape[1] == 1, 'periodic covariance can only be used for 1d data'


class PiecePoly(_ProgKernel):
    """Piecewise polynomial with compact support, degree v in {0,1,2,3}.  hyp = [log_ell, log_sigma], para = [v]
    (Core/cov.py:683-782)."""
    _op = _lib.OP_PIECEPOLY
    _nder_extra = 1

    def __init__(self, log_ell=0., v=2, log_sigma=0.):
        self.hyp = [log_ell, log_sigma]
        self.para = [v]

    def _para(self):
        v = self.para[0]
        if np.abs(v - np.round(v)) < 1e-8:
            v = int(round(v))
        assert int(v) in range(4)
        return int(v)


class Gabor(_ProgKernel):
    """Gabor kernel h(t) = exp(-t^2/(2 ell^2)) cos(2 pi t / p).  hyp = [log_ell, log_p]; as in the reference the
    period enters as exp(2*log_p) and the derivative matrices are the reference's dp*K and tan(dp)*dp*K
    (Core/cov.py:392-448)."""
    _op = _lib.OP_GABOR

    def __init__(self, log_ell=0., log_p=0.):
        self.hyp = [log_ell, log_p]
        self.para = []


class Noise(_ProgKernel):
    """White noise.  hyp = [log_sigma].  Train mode: s2*I; cross mode: s2 where the points coincide (sq. distance
    < 1e-9); self-test mode: 0, as the reference returns (Core/cov.py:1254-1300)."""
    _op = _lib.OP_NOISE

    def __init__(self, log_sigma=0.):
        self.hyp = [log_sigma]
        self.para = []


class Const(_ProgKernel):
    """Constant kernel.  hyp = [log_sigma]; the reference uses sf2 = exp(hyp[0]) (Core/cov.py:941-982)."""
    _op = _lib.OP_CONST

    def __init__(self, log_sigma=0.):
        self.hyp = [log_sigma]
        self.para = []


class Linear(_ProgKernel):
    """Linear kernel sf2 * x z'.  hyp = [log_sigma]; sf2 = exp(hyp[0]) as in the reference (Core/cov.py:986-1024)."""
    _op = _lib.OP_LINEAR

    def __init__(self, log_sigma=0.):
        self.hyp = [log_sigma]
        self.para = []


class Poly(_ProgKernel):
    """Polynomial kernel sf2 (c + x z')^d.  hyp = [log_c, log_sigma], para = [d]  (Core/cov.py:623-679)."""
    _op = _lib.OP_POLY
    _nder_extra = 1

    def __init__(self, log_c=0., d=2, log_sigma=0.):
        self.hyp = [log_c, log_sigma]
        self.para = [d]

    def _para(self):
        o = self.para[0]
        if np.abs(o - np.round(o)) < 1e-8:
            o = int(round(o))
        assert o >= 1.
        return int(o)


class Pre(Kernel):
    """Precomputed kernel matrices (Core/cov.py:1429-1455): M1 = (train+1) x test, cross-covariances with the test
    points' self-covariances in the last row; M2 = train x train.  No hyper-parameters.  On the device the training
    matrix is uploaded once per evaluation (gpk_set_pre) and read by the program's PRE leaf."""
    _op = _lib.OP_PRE

    def __init__(self, M1, M2):
        self.M1 = M1
        self.M2 = M2
        self.hyp = []
        self.para = []

    def getCovMatrix(self, x=None, z=None, mode=None):
        if mode == 'self_test':
            A = self.M1[-1, :]
            return np.reshape(A, (A.shape[0], 1))
        if mode == 'train':
            return self.M2
        if mode == 'cross':
            return self.M1[:-1, :]

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        if der is not None:
            raise Exception("Error: NO optimization in precomputed kernel matrix")
        return 0

    def _pre_leaves(self):
        return [self]


def _host_combine(kernel):
    """Composites that contain a Pre leaf are combined on the host from their parts (the precomputed parts ARE host
    matrices; the other parts still come from the device)."""
    return len(kernel._pre_leaves()) > 0


class _Pair(Kernel):
    _pair_op = None

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

    def _emit(self, nodes, h0):
        a = self.cov1._emit(nodes, h0)
        b = self.cov2._emit(nodes, h0 + len(self.cov1.hyp))
        nodes.append((self._pair_op, a, b, -1, 0.0))
        return len(nodes) - 1

    def _pre_leaves(self):
        return self.cov1._pre_leaves() + self.cov2._pre_leaves()


class SumOfKernel(_Pair):
    """k1 + k2 (Core/cov.py:265-295)."""
    _pair_op = _lib.OP_SUM

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        if _host_combine(self):
            return self.cov1.getCovMatrix(x, z, mode) + self.cov2.getCovMatrix(x, z, mode)
        return self._prog_matrix(x, z, mode, None)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        n1 = len(self.cov1.hyp)
        if not der < len(self.hyp):
            raise Exception("Error: der out of range for covSum")
        if _host_combine(self):
            if der < n1:
                return self.cov1.getDerMatrix(x, z, mode, der)
            return self.cov2.getDerMatrix(x, z, mode, der - n1)
        return self._prog_matrix(x, z, mode, der)


class ProductOfKernel(_Pair):
    """k1 * k2 (Core/cov.py:230-261)."""
    _pair_op = _lib.OP_PROD

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        if _host_combine(self):
            return self.cov1.getCovMatrix(x, z, mode) * self.cov2.getCovMatrix(x, z, mode)
        return self._prog_matrix(x, z, mode, None)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        n1 = len(self.cov1.hyp)
        if not der < len(self.hyp):
            raise Exception("Error: der out of range for covProduct")
        if _host_combine(self):
            if der < n1:
                return self.cov1.getDerMatrix(x, z, mode, der) * self.cov2.getCovMatrix(x, z, mode)
            return self.cov2.getDerMatrix(x, z, mode, der - n1) * self.cov1.getCovMatrix(x, z, mode)
        return self._prog_matrix(x, z, mode, der)


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

    def _emit(self, nodes, h0):
        a = self.cov._emit(nodes, h0 + 1)
        nodes.append((_lib.OP_SCALE, a, -1, h0, 0.0))
        return len(nodes) - 1

    def _pre_leaves(self):
        return self.cov._pre_leaves()

    def getCovMatrix(self, x=None, z=None, mode=None):
        self.checkInputGetCovMatrix(x, z, mode)
        if _host_combine(self):
            return np.exp(self.hyp[0]) * self.cov.getCovMatrix(x, z, mode)
        return self._prog_matrix(x, z, mode, None)

    def getDerMatrix(self, x=None, z=None, mode=None, der=None):
        self.checkInputGetDerMatrix(x, z, mode, der)
        if _host_combine(self):
            sf2 = np.exp(self.hyp[0])
            if der == 0:
                return 2. * sf2 * self.cov.getCovMatrix(x, z, mode)
            return sf2 * self.cov.getDerMatrix(x, z, mode, der - 1)
        return self._prog_matrix(x, z, mode, der)
