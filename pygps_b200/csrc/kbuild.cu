// kbuild.cu - covariance-matrix build: pairwise squared distance + kernel map,
// fused with the scaling / +I / lower-triangle-only store the factorisation wants.
//
// Replaces cov.RBF/RBFard/Matern.getCovMatrix and getDerMatrix
// (/root/reference/pyGPs/Core/cov.py:796-828, 887-938, 1124-1182) including
// scipy.spatial.distance.cdist(...,'sqeuclidean') and np.exp.
//
// The squared distance is the DIRECT sum_d (a_d-b_d)^2 that cdist computes, not
// the |a|^2+|b|^2-2ab expansion: it is exact 0 on the diagonal and bit-symmetric
// like the reference's, and with D = 8..32 the contraction is < 1% of the work
// of the fp64 exp() in the epilogue, so there is nothing for a tensor core to
// win here (hard part 2).  The kernel is bound by the fp64 pipe
// (exp ~ 30 DFMA) and by the HBM write of the matrix, not by the distance.
#include "gpk_internal.cuh"

namespace gpk {

constexpr int CT = 64;   // output tile edge
constexpr int CDK = 16;  // input-dimension chunk staged in shared memory

// out[p, d] = X[p, d] (*|/) scale[d] for p < n, 0 for n <= p < np.
// divide=1 reproduces `x/ell` (Core/cov.py:804) and, with premul=sqrt(d), `sqrt(d)*x/ell` (:1141);
// divide=0 reproduces `x*ell_inv` (:899).
__global__ void prescale_kernel(const double* __restrict__ X, int64_t n, int64_t np, int D,
                                const double* __restrict__ scale, int divide, double premul,
                                double* __restrict__ out) {
  const int64_t total = np * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / D;
    const int d = (int)(i - p * D);
    double v = 0.0;
    if (p < n) v = divide ? (premul * X[i]) / scale[d] : X[i] * scale[d];
    out[i] = v;
  }
}

__device__ __forceinline__ double cov_value(const CovArgs& a, double d2, double dd2) {
  // d2: scaled squared distance; dd2: squared scaled difference along ard_dim (EPI_DER_ARD only)
  if (a.kind == GPK_COV_MATERN) {
    const double t = sqrt(d2);
    const double e = exp(-t);
    double f, df;
    switch (a.matern_d) {
      case 1: f = 1.0; df = 1.0; break;
      case 3: f = 1.0 + t; df = t; break;
      case 5: f = 1.0 + t + t * t / 3.0; df = (t + t * t) / 3.0; break;
      default: f = 1.0 + t + 2.0 * t * t / 5.0 + t * t * t / 15.0; df = (3.0 * t + 3.0 * t * t + t * t * t) / 15.0; break;
    }
    if (a.epi == EPI_COV) return a.sf2 * f * e;
    if (a.epi == EPI_DER_SF) return 2.0 * a.sf2 * f * e;
    return a.sf2 * df * t * e;  // d/dlog(ell), mathematically correct: the reference reuses K as the distance
                              // (:1173-1177) and its d=7 polynomial (:1114) has t where 3t belongs
  }
  const double k = a.sf2 * exp(-0.5 * d2);
  switch (a.epi) {
    case EPI_COV: return k;
    case EPI_DER_ELL: return k * d2;
    case EPI_DER_SF: return 2.0 * k;
    default: return k * dd2;
  }
}

// out[f + s*ld] for a 64x64 tile; thread (tf = tid%16, ts = tid/16) owns f = f0+tf+16a, s = s0+4ts+b.
__global__ void __launch_bounds__(256) cov_kernel(const CovArgs a) {
  __shared__ double Fs[CDK][CT + 1];
  __shared__ double Ss[CDK][CT + 1];
  const int bf = blockIdx.x, bs = blockIdx.y;
  const int tid = threadIdx.x;
  const int tf = tid & 15, ts = tid >> 4;
  const int64_t f0 = (int64_t)bf * CT, s0 = (int64_t)bs * CT;
  // global index of the tile's first slow point (block-cyclic column ownership on multi-GPU runs)
  const int64_t gs0 = (a.s_bstride > 0) ? (s0 / NB) * a.s_bstride * NB + (int64_t)a.s_boff * NB + s0 % NB : s0;
  if (a.lower_only && f0 + CT - 1 < gs0) return;

  double acc[4][4];
  double accd[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.0; accd[i][j] = 0.0; }

  for (int d0 = 0; d0 < a.D; d0 += CDK) {
    const int dc = min(CDK, a.D - d0);
    for (int idx = tid; idx < CT * CDK; idx += 256) {
      const int p = idx / CDK, d = idx % CDK;
      double vf = 0.0, vs = 0.0;
      if (d < dc) {
        if (f0 + p < a.nF) vf = a.F[(f0 + p) * a.D + d0 + d];
        if (gs0 + p < a.nS) vs = a.S[(gs0 + p) * a.D + d0 + d];
      }
      Fs[d][p] = vf;
      Ss[d][p] = vs;
    }
    __syncthreads();
    for (int d = 0; d < dc; ++d) {
      double fv[4], sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) fv[i] = Fs[d][tf + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) sv[j] = Ss[d][ts * 4 + j];
      const bool isd = (a.epi == EPI_DER_ARD) && (d0 + d == a.ard_dim);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double df = fv[i] - sv[j];
          acc[i][j] = fma(df, df, acc[i][j]);
          if (isd) accd[i][j] = df * df;
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t s = s0 + ts * 4 + j;
    const int64_t gs = gs0 + ts * 4 + j;
    if (s >= a.pS) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t f = f0 + tf + 16 * i;
      if (f >= a.pF) continue;
      double v;
      if (f < a.nF && gs < a.nS) {
        v = cov_value(a, acc[i][j], accd[i][j]) * a.scale;
        if (a.same_set && f == gs) v += a.diag_add;
        if (a.lower_only && f < gs) v = 0.0;
      } else {
        v = (a.pad_identity && f == gs) ? 1.0 : 0.0;
      }
      a.out[f + s * a.ld] = v;
    }
  }
}

int launch_prescale(Handle* h, cudaStream_t st, const double* X, int64_t n, int64_t np, int D, const double* scale,
                    int divide, double premul, double* out) {
  const int64_t total = np * D;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  prescale_kernel<<<blocks, 256, 0, st>>>(X, n, np, D, scale, divide, premul, out);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

int launch_cov(Handle* h, cudaStream_t st, const CovArgs& a) {
  const int64_t gf = (a.pF + CT - 1) / CT, gs = (a.pS + CT - 1) / CT;
  if (gf <= 0 || gs <= 0) return 0;
  if (gs > 65535) return GPK_ERR_ARG;
  dim3 grid((unsigned)gf, (unsigned)gs);
  cov_kernel<<<grid, 256, 0, st>>>(a);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
