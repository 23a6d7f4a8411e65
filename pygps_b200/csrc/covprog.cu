// covprog.cu - composite covariance functions evaluated on the device.
//
// The reference builds composite kernels (k1 + k2, k1 * k2, c * k; Core/cov.py:230-328) by materialising one n x n
// matrix per leaf on the host and combining them with numpy, and its derivative loop (Core/inf.py:376-377) materialises
// one more n x n matrix per hyper-parameter.  Here a composite is a small PROGRAM - the expression tree in post-order -
// that every thread evaluates for its own matrix entries from ONE pass over the pair's coordinates (squared distance, dot
// product, ARD-weighted distances).  The same program, run backwards, gives d k / d leaf for every leaf, so the
// fused reduction  dnlZ.cov[h] = 1/2 sum_ij Q_ij dK_ij/dhyp_h  covers all hyper-parameters of all components in one
// pass over Q without materialising a single derivative matrix.
//
// Leaves (reference classes, Core/cov.py): RBF :786, RBFunit :832, RBFard :872, Matern :1078, RQ :1304, RQard :1356,
// Periodic :1186, PiecePoly :683, Gabor :392, Noise :1254, Const :941, Linear :986, Poly :623, Pre :1429 (train mode).
// Formulas - including the reference's conventions (Const/ScaleOfKernel use exp(hyp), their derivative carries a factor 2;
// Gabor's period is exp(2 hyp); Noise is exactly the identity in train mode and 0 in self-test mode; Linear/Poly/Const add
// 1e-10 to the training diagonal) - follow the reference line by line; Matern's length-scale derivative is the true one
// (the reference's is broken, see kbuild.cu).
#include <cmath>
#include "gpk_internal.cuh"

namespace gpk {

struct PairGeom {
  double r2;                       // sum_d (x_d - z_d)^2, raw inputs
  double dot;                      // sum_d x_d z_d
  double ard2[PROG_MAX_ARD];       // sum_d w_d (x_d - z_d)^2 per ARD leaf
};

__device__ __forceinline__ double ipow(double b, int e) {
  double r = 1.0;
  for (; e > 0; --e) r *= b;
  return r;
}

__device__ __forceinline__ double matern_f(int d, double t) {
  switch (d) {
    case 1: return 1.0;
    case 3: return 1.0 + t;
    case 5: return 1.0 + t + t * t / 3.0;
    default: return 1.0 + t + 2.0 * t * t / 5.0 + t * t * t / 15.0;
  }
}
__device__ __forceinline__ double matern_df(int d, double t) {
  switch (d) {
    case 1: return 1.0;
    case 3: return t;
    case 5: return (t + t * t) / 3.0;
    default: return (3.0 * t + 3.0 * t * t + t * t * t) / 15.0;
  }
}
// PiecePoly (Core/cov.py:696-727)
__device__ __forceinline__ double pp_f(int v, double r, double j) {
  switch (v) {
    case 0: return 1.0;
    case 1: return 1.0 + (j + 1.0) * r;
    case 2: return 1.0 + (j + 2.0) * r + (j * j + 4.0 * j + 3.0) / 3.0 * r * r;
    default: return 1.0 + (j + 3.0) * r + (6.0 * j * j + 36.0 * j + 45.0) / 15.0 * r * r +
                    (j * j * j + 9.0 * j * j + 23.0 * j + 15.0) / 15.0 * r * r * r;
  }
}
__device__ __forceinline__ double pp_df(int v, double r, double j) {
  switch (v) {
    case 0: return 0.0;
    case 1: return j + 1.0;
    case 2: return (j + 2.0) + 2.0 * (j * j + 4.0 * j + 3.0) / 3.0 * r;
    default: return (j + 3.0) + 2.0 * (6.0 * j * j + 36.0 * j + 45.0) / 15.0 * r +
                    (j * j * j + 9.0 * j * j + 23.0 * j + 15.0) / 5.0 * r * r;
  }
}

// value of one leaf.  diag: train mode and the two points are the same training point; train: train mode;
// selft: self-test mode (k(z,z) vector); pre: the uploaded entry for OP_PRE
__device__ double leaf_value(const ProgNode& nd, const PairGeom& g, bool diag, bool train, bool selft, double pre) {
  switch (nd.op) {
    case OP_RBF: return nd.p1 * exp(-0.5 * g.r2 * nd.p0);
    case OP_RBFUNIT: return exp(-0.5 * g.r2 * nd.p0);
    case OP_RBFARD: return nd.p1 * exp(-0.5 * g.ard2[nd.ard]);
    case OP_MATERN: {
      const double t = sqrt(g.r2 * nd.p0);
      return nd.p1 * matern_f(nd.ipar, t) * exp(-t);
    }
    case OP_RQ: return nd.p1 * exp(-nd.p2 * log(1.0 + 0.5 * g.r2 * nd.p0 / nd.p2));
    case OP_RQARD: return nd.p1 * exp(-nd.p2 * log(1.0 + 0.5 * g.ard2[nd.ard] / nd.p2));
    case OP_PERIODIC: {
      const double s = sin(M_PI * sqrt(g.r2) / nd.p1) * nd.p0;
      return nd.p2 * exp(-2.0 * s * s);
    }
    case OP_PIECEPOLY: {
      const double r = sqrt(g.r2 * nd.p0);
      const double m = fmax(1.0 - r, 0.0);
      return nd.p1 * pp_f(nd.ipar, r, nd.p2) * ipow(m, (int)nd.p2 + nd.ipar);
    }
    case OP_GABOR: {
      const double d2 = g.r2 * nd.p0;
      return exp(-0.5 * d2) * cos(2.0 * M_PI * sqrt(d2) * nd.p2 / nd.p1);
    }
    case OP_NOISE: return selft ? 0.0 : (train ? (diag ? nd.p0 : 0.0) : (g.r2 < 1.0e-9 ? nd.p0 : 0.0));
    case OP_CONST: return nd.p0 + (diag ? 1.0e-10 : 0.0);
    case OP_LINEAR: return nd.p0 * (g.dot + (diag ? 1.0e-10 : 0.0));
    case OP_POLY: return nd.p1 * ipow(nd.p0 + g.dot + (diag ? 1.0e-10 : 0.0), nd.ipar);
    case OP_PRE: return pre;
    default: return 0.0;
  }
}

// forward pass: val[i] for every node; returns the root
__device__ double prog_forward(const CovProg& P, const PairGeom& g, bool diag, bool train, bool selft, double pre,
                               double* val) {
  for (int i = 0; i < P.n_nodes; ++i) {
    const ProgNode& nd = P.node[i];
    double v;
    if (nd.op == OP_SUM) v = val[nd.a] + val[nd.b];
    else if (nd.op == OP_PROD) v = val[nd.a] * val[nd.b];
    else if (nd.op == OP_SCALE) v = nd.p0 * val[nd.a];
    else v = leaf_value(nd, g, diag, train, selft, pre);
    val[i] = v;
  }
  return val[P.n_nodes - 1];
}

// backward pass: adj[i] = d root / d val[i]
__device__ void prog_backward(const CovProg& P, const double* val, double* adj) {
  for (int i = 0; i < P.n_nodes; ++i) adj[i] = 0.0;
  adj[P.n_nodes - 1] = 1.0;
  for (int i = P.n_nodes - 1; i >= 0; --i) {
    const ProgNode& nd = P.node[i];
    const double a = adj[i];
    if (nd.op == OP_SUM) { adj[nd.a] += a; adj[nd.b] += a; }
    else if (nd.op == OP_PROD) { adj[nd.a] += a * val[nd.b]; adj[nd.b] += a * val[nd.a]; }
    else if (nd.op == OP_SCALE) adj[nd.a] += a * nd.p0;
  }
}

// d leaf / d (its own hyper-parameters, in the reference's order), times `w`, accumulated into acc[h0 ...].  The per-dimension
// ARD derivatives need the pair's coordinates: xi, xj (D doubles each).  diag as in leaf_value (train mode only: the
// derivative reduction runs over the training matrix).
__device__ void leaf_grad(const CovProg& P, const ProgNode& nd, const PairGeom& g, bool diag, bool train, double w,
                          const double* xi, const double* xj, double* acc) {
  switch (nd.op) {
    case OP_RBF: {
      const double d2 = g.r2 * nd.p0, k = nd.p1 * exp(-0.5 * d2);
      acc[nd.h0] += w * k * d2;
      acc[nd.h0 + 1] += w * 2.0 * k;
      break;
    }
    case OP_RBFUNIT: {
      const double d2 = g.r2 * nd.p0;
      acc[nd.h0] += w * exp(-0.5 * d2) * d2;
      break;
    }
    case OP_RBFARD: {
      const double k = nd.p1 * exp(-0.5 * g.ard2[nd.ard]);
      for (int d = 0; d < P.D; ++d) {
        const double df = xi[d] - xj[d];
        acc[nd.h0 + d] += w * k * P.ardw[nd.ard][d] * df * df;
      }
      acc[nd.h0 + P.D] += w * 2.0 * k;
      break;
    }
    case OP_MATERN: {
      const double t = sqrt(g.r2 * nd.p0), e = exp(-t);
      acc[nd.h0] += w * nd.p1 * matern_df(nd.ipar, t) * t * e;
      acc[nd.h0 + 1] += w * 2.0 * nd.p1 * matern_f(nd.ipar, t) * e;
      break;
    }
    case OP_RQ:
    case OP_RQARD: {
      const bool ard = nd.op == OP_RQARD;
      const double d2 = ard ? g.ard2[nd.ard] : g.r2 * nd.p0;
      const double base = 1.0 + 0.5 * d2 / nd.p2, lb = log(base);
      const double K = nd.p1 * exp(-nd.p2 * lb);                 // sf2 base^-alpha
      const double K1 = K / base;                                 // sf2 base^(-alpha-1)
      int hs = nd.h0 + 1;
      if (ard) {
        for (int d = 0; d < P.D; ++d) {
          const double df = xi[d] - xj[d];
          acc[nd.h0 + d] += w * K1 * P.ardw[nd.ard][d] * df * df;
        }
        hs = nd.h0 + P.D;
      } else {
        acc[nd.h0] += w * K1 * d2;
      }
      acc[hs] += w * 2.0 * K;
      acc[hs + 1] += w * K * (0.5 * d2 / base - nd.p2 * lb);
      break;
    }
    case OP_PERIODIC: {
      const double a = M_PI * sqrt(g.r2) / nd.p1;
      const double R = sin(a) * nd.p0, e = exp(-2.0 * R * R);
      acc[nd.h0] += w * 4.0 * nd.p2 * e * R * R;
      acc[nd.h0 + 1] += w * 4.0 * nd.p2 * nd.p0 * e * R * cos(a) * a;
      acc[nd.h0 + 2] += w * 2.0 * nd.p2 * e;
      break;
    }
    case OP_PIECEPOLY: {
      const double r = sqrt(g.r2 * nd.p0), m = fmax(1.0 - r, 0.0);
      const int e = (int)nd.p2 + nd.ipar;
      const double f = pp_f(nd.ipar, r, nd.p2);
      acc[nd.h0] += w * nd.p1 * ipow(m, e - 1) * r * ((double)e * f - m * pp_df(nd.ipar, r, nd.p2));
      acc[nd.h0 + 1] += w * 2.0 * nd.p1 * f * ipow(m, e);
      break;
    }
    case OP_GABOR: {
      // the reference's derivative formulas (Core/cov.py:441-446): dp*K and tan(dp)*dp*K
      const double d2 = g.r2 * nd.p0, dp = 2.0 * M_PI * sqrt(d2) * nd.p2 / nd.p1;
      const double e = exp(-0.5 * d2);
      acc[nd.h0] += w * dp * e * cos(dp);
      acc[nd.h0 + 1] += w * dp * e * sin(dp);
      break;
    }
    case OP_NOISE: acc[nd.h0] += w * 2.0 * ((train ? diag : (g.r2 < 1.0e-9)) ? nd.p0 : 0.0); break;
    case OP_CONST: acc[nd.h0] += w * 2.0 * nd.p0; break;
    case OP_LINEAR: acc[nd.h0] += w * 2.0 * nd.p0 * (g.dot + (diag ? 1.0e-16 : 0.0)); break;
    case OP_POLY: {
      const double b = nd.p0 + g.dot;
      acc[nd.h0] += w * nd.p0 * (double)nd.ipar * nd.p1 * ipow(b, nd.ipar - 1);
      acc[nd.h0 + 1] += w * 2.0 * nd.p1 * ipow(b, nd.ipar);
      break;
    }
    default: break;
  }
}

__device__ __forceinline__ void pair_geom(const CovProg& P, const double* xi, const double* xj, PairGeom& g) {
  double r2 = 0.0, dot = 0.0, a0 = 0.0, a1 = 0.0;
  for (int d = 0; d < P.D; ++d) {
    const double u = xi[d], v = xj[d], df = u - v, q = df * df;
    r2 += q;
    dot = fma(u, v, dot);
    if (P.n_ard > 0) a0 = fma(P.ardw[0][d], q, a0);
    if (P.n_ard > 1) a1 = fma(P.ardw[1][d], q, a1);
  }
  g.r2 = r2; g.dot = dot; g.ard2[0] = a0; g.ard2[1] = a1;
}

constexpr int PT = 64;     // tile edge of the program kernels

// out[f + s*ld] over a 64x64 tile; same flags as cov_kernel (scale, diag_add, lower_only, pad_identity, s_bstride);
// a.prog_der >= 0: the derivative matrix w.r.t. hyper-parameter prog_der instead of the covariance (getDerMatrix).
__global__ void __launch_bounds__(256) cov_prog_kernel(const CovArgs a) {
  extern __shared__ double psm[];
  __shared__ CovProg P;
  const int D = a.D;
  double* Fs = psm;                 // [64][D]
  double* Ss = psm + PT * D;        // [64][D]
  const int tid = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(a.prog);
    int* dst = reinterpret_cast<int*>(&P);
    for (int i = tid; i < (int)(sizeof(CovProg) / sizeof(int)); i += 256) dst[i] = src[i];
  }
  const int64_t f0 = (int64_t)blockIdx.x * PT, s0 = (int64_t)blockIdx.y * PT;
  const int64_t gs0 = (a.s_bstride > 0) ? (s0 / NB) * a.s_bstride * NB + (int64_t)a.s_boff * NB + s0 % NB : s0;
  if (a.lower_only && f0 + PT - 1 < gs0) return;
  for (int idx = tid; idx < PT * D; idx += 256) {
    const int p = idx / D, d = idx - p * D;
    Fs[idx] = (f0 + p < a.nF) ? a.F[(f0 + p) * D + d] : 0.0;
    Ss[idx] = (gs0 + p < a.nS) ? a.S[(gs0 + p) * D + d] : 0.0;
  }
  __syncthreads();
  const bool train = a.same_set != 0;
  double val[PROG_MAX_NODES], adj[PROG_MAX_NODES], gacc[PROG_MAX_HYP];
  const int tf = tid & 63, ts0 = tid >> 6;          // f fastest across lanes: coalesced stores
  const int64_t f = f0 + tf;
  for (int j = ts0; j < PT; j += 4) {
    const int64_t s = s0 + j, gs = gs0 + j;
    if (s >= a.pS || f >= a.pF) continue;
    double v;
    if (f < a.nF && gs < a.nS) {
      const bool diag = train && f == gs;
      PairGeom g;
      pair_geom(P, Fs + tf * D, Ss + j * D, g);
      const double pre = a.pre ? a.pre[f + gs * a.pre_ld] : 0.0;
      v = prog_forward(P, g, diag, train, false, pre, val);
      if (a.prog_der1 > 0) {
        prog_backward(P, val, adj);
        for (int h = 0; h < P.nhyp; ++h) gacc[h] = 0.0;
        for (int i = 0; i < P.n_nodes; ++i) {
          const ProgNode& nd = P.node[i];
          if (nd.op == OP_SCALE) gacc[nd.h0] += adj[i] * 2.0 * nd.p0 * val[nd.a];   // Core/cov.py:323-325
          else if (nd.op < OP_SUM) leaf_grad(P, nd, g, diag, train, adj[i], Fs + tf * D, Ss + j * D, gacc);
        }
        v = gacc[a.prog_der1 - 1];
      }
      v *= a.scale;
      if (train && f == gs) v += a.diag_add;
      if (a.lower_only && f < gs) v = 0.0;
    } else {
      v = (a.pad_identity && f == gs) ? 1.0 : 0.0;
    }
    a.out[f + s * a.ld] = v;
  }
}

// self-test mode: out[i] = k(z_i, z_i) with the reference's self-test conventions (distance 0, Noise -> 0, no 1e-10)
__global__ void cov_prog_diag_kernel(const CovProg* __restrict__ prog, const double* __restrict__ Z, int64_t m, int D,
                                     int der, double* __restrict__ out) {
  __shared__ CovProg P;
  {
    const int* src = reinterpret_cast<const int*>(prog);
    int* dst = reinterpret_cast<int*>(&P);
    for (int i = threadIdx.x; i < (int)(sizeof(CovProg) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double val[PROG_MAX_NODES], adj[PROG_MAX_NODES], gacc[PROG_MAX_HYP];
  PairGeom g;
  g.r2 = 0.0; g.ard2[0] = 0.0; g.ard2[1] = 0.0;
  double dot = 0.0;
  for (int d = 0; d < D; ++d) dot = fma(Z[i * D + d], Z[i * D + d], dot);
  g.dot = dot;
  double v = prog_forward(P, g, false, false, true, 0.0, val);
  if (der >= 0) {
    prog_backward(P, val, adj);
    for (int h = 0; h < P.nhyp; ++h) gacc[h] = 0.0;
    for (int k = 0; k < P.n_nodes; ++k) {
      const ProgNode& nd = P.node[k];
      if (nd.op == OP_SCALE) gacc[nd.h0] += adj[k] * 2.0 * nd.p0 * val[nd.a];
      else if (nd.op == OP_NOISE) { /* self-test derivative of Noise is 0 (Core/cov.py:1286-1288) */ }
      else if (nd.op < OP_SUM) leaf_grad(P, nd, g, false, true, adj[k], Z + i * D, Z + i * D, gacc);
    }
    v = gacc[der];
  }
  out[i] = v;
}

// ---------------------------------------------------------------------------
// fused dnlZ reduction for programs:  part[cta][h] = sum_{i>=j} w_ij Q_ij dK_ij/dhyp_h , part[cta][nhyp] = sum_i Q_ii
// with Q = Ainv*inv_sn2 - alpha alpha' (Core/inf.py:373-377); one 64x64 tile of the lower triangle per CTA.
// ---------------------------------------------------------------------------
struct DnlzProgArgs {
  const CovProg* prog; const double* X; int64_t n; int D;
  const double* Ainv; int64_t ld; const double* alpha; double inv_sn2;
  const double* pre; int64_t pre_ld;
  double* part;
};

__global__ void __launch_bounds__(256) dnlz_prog_kernel(const DnlzProgArgs a) {
  extern __shared__ double psm[];
  __shared__ CovProg P;
  __shared__ double red[32];
  const int tid = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(a.prog);
    int* dst = reinterpret_cast<int*>(&P);
    for (int i = tid; i < (int)(sizeof(CovProg) / sizeof(int)); i += 256) dst[i] = src[i];
  }
  const int bi = blockIdx.x, bj = blockIdx.y;
  const int cta = bi + bj * gridDim.x;
  const int D = a.D;
  double acc[PROG_MAX_HYP + 1];
  __syncthreads();
  const int nh = P.nhyp;
  for (int h = 0; h <= nh; ++h) acc[h] = 0.0;
  if (bi >= bj) {
    double* Xi = psm;
    double* Xj = psm + PT * D;
    const int64_t i0 = (int64_t)bi * PT, j0 = (int64_t)bj * PT;
    for (int idx = tid; idx < PT * D; idx += 256) {
      const int p = idx / D, d = idx - p * D;
      Xi[idx] = (i0 + p < a.n) ? a.X[(i0 + p) * D + d] : 0.0;
      Xj[idx] = (j0 + p < a.n) ? a.X[(j0 + p) * D + d] : 0.0;
    }
    __syncthreads();
    double val[PROG_MAX_NODES], adj[PROG_MAX_NODES];
    const int ti = tid & 63, tj0 = tid >> 6;
    const int64_t i = i0 + ti;
    const double ai = (i < a.n) ? a.alpha[i] : 0.0;
    for (int jj = tj0; jj < PT; jj += 4) {
      const int64_t j = j0 + jj;
      if (i >= a.n || j >= a.n || i < j) continue;
      const double q = a.Ainv[i + j * a.ld] * a.inv_sn2 - ai * a.alpha[j];
      const double w = (i == j) ? q : 2.0 * q;
      const bool diag = (i == j);
      PairGeom g;
      pair_geom(P, Xi + ti * D, Xj + jj * D, g);
      const double pre = a.pre ? a.pre[i + j * a.pre_ld] : 0.0;
      prog_forward(P, g, diag, true, false, pre, val);
      prog_backward(P, val, adj);
      for (int k = 0; k < P.n_nodes; ++k) {
        const ProgNode& nd = P.node[k];
        if (nd.op == OP_SCALE) acc[nd.h0] += w * adj[k] * 2.0 * nd.p0 * val[nd.a];
        else if (nd.op < OP_SUM) leaf_grad(P, nd, g, diag, true, w * adj[k], Xi + ti * D, Xj + jj * D, acc);
      }
      if (diag) acc[nh] += q;
    }
  }
  for (int h = 0; h <= nh; ++h) {
    double v = acc[h];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      double t = (tid < 8) ? red[tid] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (tid == 0) a.part[(int64_t)cta * (nh + 1) + h] = t;
    }
  }
}

__global__ void __launch_bounds__(1024) dnlz_prog_finish_kernel(const double* __restrict__ part, int64_t nctas, int nacc,
                                                                double* __restrict__ res) {
  __shared__ double sh[32];
  for (int q = 0; q < nacc; ++q) {
    double s = 0.0;
    for (int64_t c = threadIdx.x; c < nctas; c += blockDim.x) s += part[c * nacc + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double t = sh[threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) res[q] = t;
    }
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int nhyp_of(int op, int D) {
  switch (op) {
    case OP_RBF: return 2;
    case OP_RBFUNIT: return 1;
    case OP_RBFARD: return D + 1;
    case OP_MATERN: return 2;
    case OP_RQ: return 3;
    case OP_RQARD: return D + 2;
    case OP_PERIODIC: return 3;
    case OP_PIECEPOLY: return 2;
    case OP_GABOR: return 2;
    case OP_NOISE: return 1;
    case OP_CONST: return 1;
    case OP_LINEAR: return 1;
    case OP_POLY: return 2;
    case OP_PRE: return 0;
    case OP_SCALE: return 1;
    default: return 0;
  }
}

// nodes (post-order, root last) + the flat LOG hyper-parameter vector -> the device program.  Leaf hyper-parameters are
// read at hyp[node.hyp0 ...] in the reference's order for that class; a ScaleOfKernel node's scalar is hyp[hyp0].
int prog_compile(const gpk_cov_node* nodes, int nnodes, const double* hyp, int nhyp, int D, CovProg* out) {
  if (!nodes || !out || nnodes < 1 || nnodes > PROG_MAX_NODES || nhyp < 0 || nhyp > PROG_MAX_HYP || D < 1) return GPK_ERR_ARG;
  if (nhyp > 0 && !hyp) return GPK_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  out->n_nodes = nnodes; out->nhyp = nhyp; out->D = D; out->n_ard = 0;
  for (int i = 0; i < nnodes; ++i) {
    const gpk_cov_node& s = nodes[i];
    ProgNode& nd = out->node[i];
    nd.op = s.op; nd.a = s.a; nd.b = s.b; nd.h0 = s.hyp0; nd.ard = 0; nd.ipar = 0;
    const int need = nhyp_of(s.op, D);
    if (need > 0 && (s.hyp0 < 0 || s.hyp0 + need > nhyp)) return GPK_ERR_ARG;
    const double* hp = (need > 0) ? hyp + s.hyp0 : nullptr;
    if (s.op == OP_SUM || s.op == OP_PROD) {
      if (s.a < 0 || s.a >= i || s.b < 0 || s.b >= i) return GPK_ERR_ARG;
      continue;
    }
    if (s.op == OP_SCALE) {
      if (s.a < 0 || s.a >= i) return GPK_ERR_ARG;
      nd.p0 = std::exp(hp[0]);                                  // Core/cov.py:317
      continue;
    }
    switch (s.op) {
      case OP_RBF: nd.p0 = std::exp(-2.0 * hp[0]); nd.p1 = std::exp(2.0 * hp[1]); break;
      case OP_RBFUNIT: nd.p0 = std::exp(-2.0 * hp[0]); break;
      case OP_RBFARD:
      case OP_RQARD: {
        if (D > PROG_MAX_D || out->n_ard >= PROG_MAX_ARD) return GPK_ERR_ARG;
        nd.ard = out->n_ard++;
        for (int d = 0; d < D; ++d) out->ardw[nd.ard][d] = std::exp(-2.0 * hp[d]);
        nd.p1 = std::exp(2.0 * hp[D]);
        if (s.op == OP_RQARD) nd.p2 = std::exp(hp[D + 1]);
        break;
      }
      case OP_MATERN: {
        int d = (int)std::lround(s.para);
        if (!(d == 1 || d == 3 || d == 5 || d == 7)) d = 3;     // Core/cov.py:1132-1136
        nd.ipar = d;
        nd.p0 = (double)d * std::exp(-2.0 * hp[0]);             // t^2 = d r^2 / ell^2
        nd.p1 = std::exp(2.0 * hp[1]);
        break;
      }
      case OP_RQ: nd.p0 = std::exp(-2.0 * hp[0]); nd.p1 = std::exp(2.0 * hp[1]); nd.p2 = std::exp(hp[2]); break;
      case OP_PERIODIC:
        if (D != 1) return GPK_ERR_ARG;                         // Core/cov.py:1201: 1-d data only
        nd.p0 = std::exp(-hp[0]); nd.p1 = std::exp(hp[1]); nd.p2 = std::exp(2.0 * hp[2]);
        break;
      case OP_PIECEPOLY: {
        const int v = (int)std::lround(s.para);
        if (v < 0 || v > 3) return GPK_ERR_ARG;
        nd.ipar = v;
        nd.p0 = std::exp(-2.0 * hp[0]); nd.p1 = std::exp(2.0 * hp[1]);
        nd.p2 = std::floor(0.5 * D) + v + 1.0;                  // j, Core/cov.py:742
        break;
      }
      case OP_GABOR: nd.p0 = std::exp(-2.0 * hp[0]); nd.p1 = std::exp(2.0 * hp[1]); nd.p2 = std::exp(hp[0]); break;
      case OP_NOISE: nd.p0 = std::exp(2.0 * hp[0]); break;
      case OP_CONST: nd.p0 = std::exp(hp[0]); break;
      case OP_LINEAR: nd.p0 = std::exp(hp[0]); break;
      case OP_POLY: {
        const int o = (int)std::lround(s.para);
        if (o < 1 || o > 16) return GPK_ERR_ARG;
        nd.ipar = o; nd.p0 = std::exp(hp[0]); nd.p1 = std::exp(2.0 * hp[1]);
        break;
      }
      case OP_PRE: break;
      default: return GPK_ERR_ARG;
    }
  }
  return 0;
}

bool prog_has_op(const CovProg& p, int op) {
  for (int i = 0; i < p.n_nodes; ++i)
    if (p.node[i].op == op) return true;
  return false;
}

int prog_upload(Handle* h, cudaStream_t st, const CovProg& p) {
  if (!h->dProg) GPK_CK(h, cudaMalloc((void**)&h->dProg, sizeof(CovProg)));
  if (!h->hProgPinned) GPK_CK(h, cudaMallocHost((void**)&h->hProgPinned, sizeof(CovProg)));
  std::memcpy(h->hProgPinned, &p, sizeof(CovProg));
  GPK_CK(h, cudaMemcpyAsync(h->dProg, h->hProgPinned, sizeof(CovProg), cudaMemcpyHostToDevice, st));
  return 0;
}

int launch_cov_prog(Handle* h, cudaStream_t st, const CovArgs& a) {
  const int64_t gf = (a.pF + PT - 1) / PT, gs = (a.pS + PT - 1) / PT;
  if (gf <= 0 || gs <= 0) return 0;
  if (gs > 65535) return GPK_ERR_ARG;
  const size_t smem = (size_t)2 * PT * a.D * sizeof(double);
  GPK_SMEM_ATTR(h, cov_prog_kernel, (size_t)2 * PT * PROG_MAX_D * sizeof(double));
  cov_prog_kernel<<<dim3((unsigned)gf, (unsigned)gs), 256, smem, st>>>(a);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int launch_cov_prog_diag(Handle* h, cudaStream_t st, const CovProg* dprog, const double* Z, int64_t m, int D, int der,
                         double* out) {
  cov_prog_diag_kernel<<<(unsigned)((m + 127) / 128), 128, 0, st>>>(dprog, Z, m, D, der, out);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// res[0..nhyp-1] = sum Q o dK_h (before the 1/2), res[nhyp] = trace(Q)
int launch_dnlz_prog(Handle* h, cudaStream_t st, const CovProg* dprog, int nhyp, const double* X, int64_t n, int D,
                     const double* Ainv, int64_t ld, const double* alpha, double inv_sn2, const double* pre,
                     int64_t pre_ld, double* part, int64_t part_cap, double* res) {
  const int64_t g = (n + PT - 1) / PT;
  if (g > 65535) return GPK_ERR_ARG;
  if (g * g * (nhyp + 1) > part_cap) return GPK_ERR_ARG;
  DnlzProgArgs a{dprog, X, n, D, Ainv, ld, alpha, inv_sn2, pre, pre_ld, part};
  const size_t smem = (size_t)2 * PT * D * sizeof(double);
  GPK_SMEM_ATTR(h, dnlz_prog_kernel, (size_t)2 * PT * PROG_MAX_D * sizeof(double));
  dnlz_prog_kernel<<<dim3((unsigned)g, (unsigned)g), 256, smem, st>>>(a);
  dnlz_prog_finish_kernel<<<1, 1024, 0, st>>>(part, g * g, nhyp + 1, res);
  h->stats.launches += 2;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
