// fitc.cu - inf.FITC_Exact.evaluate and the FITC branch of GP.predict on the GPU.
//
// Replaces /root/reference/pyGPs/Core/inf.py:398-455 and Core/gp.py:418 (with cov.FITCOfKernel.getCovMatrix,
// Core/cov.py:352-370).  Every O(M^2 n) step is the DMMA GEMM of gemm_nt.cu:
//   V  = Luu^-1 Ku           -> transposed sweep  Vt = Kut * Luu^-T            (n x M, one data point per row)
//   A2 = I + V diag(1/g) V'  -> SYRK over n after a transpose-and-scale        (Vs = diag(g^-1/2) V, M x n)
//   dnlZ: B = iKuu Ku, W = Lu^-1 (V/g), R = 2 dKu - dKuu B, R W', B W'         -> GEMMs on the same layouts
// Ku is never kept: its storage becomes V.  The two M x M factorisations reuse potrf_device.
//
// Multi-GPU (BASELINE config 4): after gpk_dist_init the data points are SHARDED over the ranks (each rank passes its
// own rows of X and y-m); U, Kuu, Luu are replicated.  The exchanges are sum-all-reduces of the M x M SYRK partial, a
// few M-vectors and scalars; the M x M tail (Lu, alpha, post.L) is computed redundantly on every rank.
#include <cmath>
#include "gpk_internal.cuh"

namespace gpk {

constexpr unsigned FULLF = 0xffffffffu;

// g_i = base - sum_u Vt[i,u]^2 ; rs_i = g_i^-1/2 ; r_i = ymm_i * rs_i  (rows >= n: r = 0)
__global__ void fitc_g_kernel(const double* __restrict__ Vt, int64_t ld, int64_t rows, int64_t cols, int64_t n,
                              double base, const double* __restrict__ ymm, double* __restrict__ g,
                              double* __restrict__ rs, double* __restrict__ r) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s0 = 0.0, s1 = 0.0;
  int64_t u = 0;
  for (; u + 1 < cols; u += 2) {
    const double a = Vt[i + u * ld], b = Vt[i + (u + 1) * ld];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  for (; u < cols; ++u) { const double a = Vt[i + u * ld]; s0 = fma(a, a, s0); }
  const double gi = base - (s0 + s1);
  const double q = 1.0 / sqrt(gi);
  g[i] = gi;
  rs[i] = q;
  r[i] = (i < n) ? ymm[i] * q : 0.0;
}

// out[c + r*ldo] = in[r + c*ldi] * (rowscale ? rowscale[r] : 1)   for r < rows, c < cols  (32x32 tiles)
__global__ void transpose_scale_kernel(const double* __restrict__ in, int64_t ldi, int64_t rows, int64_t cols,
                                       const double* __restrict__ rowscale, double* __restrict__ out, int64_t ldo) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t r = r0 + threadIdx.x, c = c0 + j;
    double v = 0.0;
    if (r < rows && c < cols) v = in[r + c * ldi] * (rowscale ? rowscale[r] : 1.0);
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + threadIdx.x, r = r0 + j;
    if (r < rows && c < cols) out[c + r * ldo] = tile[threadIdx.x][j];
  }
}

// rows scaled in place or into out: out[i + u*ld] = in[i + u*ld] * s[i]
__global__ void rowscale_kernel(const double* __restrict__ in, int64_t ld, int64_t rows, int64_t cols,
                                const double* __restrict__ s, double* __restrict__ out) {
  const int64_t total = rows * cols;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = k % rows, u = k / rows;
    out[i + u * ld] = in[i + u * ld] * s[i];
  }
}

__device__ __forceinline__ double block_sum_f(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLF, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < nw) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLF, t, o);
  }
  return t;
}

// generic deterministic single-CTA reductions: res[q] = sum_i f_q(i)
//   q=0: log(a[i]) i<na ; q=1: b[i]^2 i<nb ; q=2: c[i]^2 i<nc ; q=3: d[i] i<nd ; q=4: 1/a[i] ; q=5: e[i]*f2[i] i<ne
__global__ void __launch_bounds__(1024) fitc_reduce_kernel(const double* a, int64_t na, const double* b, int64_t nb,
                                                           const double* c, int64_t nc, const double* d, int64_t nd,
                                                           const double* e, const double* f2, int64_t ne,
                                                           double* __restrict__ res) {
  __shared__ double sh[32];
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = threadIdx.x; i < na; i += blockDim.x) { s[0] += log(a[i]); s[4] += 1.0 / a[i]; }
  for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) s[1] = fma(b[i], b[i], s[1]);
  for (int64_t i = threadIdx.x; i < nc; i += blockDim.x) s[2] = fma(c[i], c[i], s[2]);
  for (int64_t i = threadIdx.x; i < nd; i += blockDim.x) s[3] += d[i];
  for (int64_t i = threadIdx.x; i < ne; i += blockDim.x) s[5] = fma(e[i], f2[i], s[5]);
  for (int q = 0; q < 6; ++q) {
    const double t = block_sum_f(s[q], sh);
    if (threadIdx.x == 0) res[q] = t;
  }
}

// out (full, pitch ldo) = sym(lower of S, pitch ld) - sym(lower of K, pitch ld)   for an m x m block
__global__ void sym_diff_kernel(const double* __restrict__ S, const double* __restrict__ K, int64_t ld, int64_t m,
                                double* __restrict__ out, int64_t ldo) {
  const int64_t total = m * m;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % m, c = k / m;
    const int64_t hi = r >= c ? r : c, lo = r >= c ? c : r;
    out[r + c * ldo] = S[hi + lo * ld] - K[hi + lo * ld];
  }
}

// out (full m x m, pitch ldo) from the lower triangle of S (pitch ld)
__global__ void sym_fill_kernel(const double* __restrict__ S, int64_t ld, int64_t m, double* __restrict__ out,
                                int64_t ldo) {
  const int64_t total = m * m;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % m, c = k / m;
    out[r + c * ldo] = (r >= c) ? S[r + c * ld] : S[c + r * ld];
  }
}

// per-row products of two (rows x cols) matrices: out[i] = post(sum_u P[i,u]*Q[i,u])
//   mode 0: sum ; mode 1: max(kss + sum, 0)
__global__ void rowdot2_kernel(const double* __restrict__ P, const double* __restrict__ Q, int64_t ld, int64_t rows,
                               int64_t cols, int mode, double kss, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s = 0.0;
  for (int64_t u = 0; u < cols; ++u) s = fma(P[i + u * ld], Q[i + u * ld], s);
  out[i] = (mode == 0) ? s : fmax(kss + s, 0.0);
}

// out[u] = sum_i P[i + u*ld] * v[i]   (one CTA per column, deterministic)
__global__ void __launch_bounds__(256) coldot_kernel(const double* __restrict__ P, int64_t ld, int64_t rows,
                                                     const double* __restrict__ v, const double* __restrict__ v2,
                                                     double* __restrict__ out) {
  __shared__ double sh[32];
  const int64_t u = blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) s = fma(P[i + u * ld] * (v2 ? v2[i] : 1.0), v[i], s);
  const double t = block_sum_f(s, sh);
  if (threadIdx.x == 0) out[u] = t;
}

// al_i = (ymm_i - t_i)/g_i   (rows >= n: 0)                                     Core/inf.py:431
__global__ void fitc_al_kernel(const double* ymm, const double* g, const double* t, int64_t n, int64_t rows,
                               double* al) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= rows) return;
  al[i] = (i < n) ? (ymm[i] - t[i]) / g[i] : 0.0;
}

// T <- 2*dKu - T   (element-wise; T = B' dKuu comes in, R' goes out)              :437
__global__ void fitc_r_kernel(const double* __restrict__ dKu, double* __restrict__ T, int64_t total) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x)
    T[k] = 2.0 * dKu[k] - T[k];
}

// deterministic two-stage sum of A .* B over `total` elements: part[cta] then a single-CTA finish
__global__ void __launch_bounds__(256) bigdot_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                     int64_t total, double* __restrict__ part) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x)
    s = fma(A[k], B[k], s);
  const double t = block_sum_f(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(1024) sum_finish_kernel(const double* __restrict__ part, int nparts,
                                                          double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int k = threadIdx.x; k < nparts; k += blockDim.x) s += part[k];
  const double t = block_sum_f(s, sh);
  if (threadIdx.x == 0) out[0] = t;
}

// per-hyper-parameter scalars (single CTA):
//   out[0] = w'(dKuu w) ; out[1] = (dKu' w).al ; out[2] = sum_i al_i^2 v_i ; out[3] = sum_i ww_i v_i
//   with v_i = ddiag - vscale*rb_i                                               :438-440
__global__ void __launch_bounds__(1024) fitc_hyp_scalars_kernel(const double* wv, const double* tw, int64_t Mp,
                                                                const double* dkw, const double* al,
                                                                const double* ww, const double* rb, double ddiag,
                                                                double vscale, int64_t n, double* __restrict__ out) {
  __shared__ double sh[32];
  double s[4] = {0, 0, 0, 0};
  for (int64_t u = threadIdx.x; u < Mp; u += blockDim.x) s[0] = fma(wv[u], tw[u], s[0]);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = ddiag - vscale * rb[i];
    if (dkw) s[1] = fma(dkw[i], al[i], s[1]);
    s[2] = fma(al[i] * al[i], v, s[2]);
    s[3] = fma(ww[i], v, s[3]);
  }
  for (int q = 0; q < 4; ++q) {
    const double t = block_sum_f(s[q], sh);
    if (threadIdx.x == 0) out[q] = t;
  }
}

__global__ void res_n_kernel(double* res, double n) { if (threadIdx.x == 0) { res[6] = n; res[7] = 0.0; } }

// X[i] *= s for i in [0,count)  (used to keep rank-replicated scalars from being summed G times)
__global__ void scale_small_kernel(double* __restrict__ X, int count, double s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) X[i] *= s;
}

static inline int grid1(int64_t total, int bs = 256) {
  int64_t b = (total + bs - 1) / bs;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (int)b;
}

static int transpose_scale(Handle* h, cudaStream_t st, const double* in, int64_t ldi, int64_t rows, int64_t cols,
                           const double* rowscale, double* out, int64_t ldo) {
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
  if (grid.y > 65535) return GPK_ERR_ARG;
  transpose_scale_kernel<<<grid, block, 0, st>>>(in, ldi, rows, cols, rowscale, out, ldo);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

static int vec_fwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* work,
                   double* z, int T) {
  for (int k = 0; k < T; ++k) GPK_TRY(launch_trsv_fwd(h, st, A, lda, Dinv, work, z, k, T));
  return 0;
}
static int vec_bwd(Handle* h, cudaStream_t st, const double* A, int64_t lda, const double* Dinv, double* work,
                   double* x, int T) {
  for (int k = T - 1; k >= 0; --k) GPK_TRY(launch_trsv_bwd(h, st, A, lda, Dinv, work, x, k, T));
  return 0;
}

}  // namespace gpk

using namespace gpk;

extern "C" {

int gpk_fitc_eval(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, double log_sn, const double* U,
                  int64_t M, const double* ymm, int want_der, double* nlZ, double* alpha, double* Lpost, double* dcov,
                  double* dlik, double* al_out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->dX || h->n <= 0) return GPK_ERR_STATE;
  if (!hyp || !U || M <= 0 || !ymm || !nlZ || !alpha || !Lpost) return GPK_ERR_ARG;
  if (want_der && (!dcov || !dlik || !al_out)) return GPK_ERR_ARG;
  const int64_t n = h->n, np = h->np;
  const int D = h->D;
  const int64_t Mp = round_up(M, NB);
  const int Tm = (int)(Mp / NB), Tn = (int)(np / NB);
  std::vector<double> scale;
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  GPK_TRY(kind_scale(kind, matern_d, hyp, nhyp, D, scale, &divide, &premul, &sf2));
  if (D > 200) return GPK_ERR_ARG;   // pinned staging holds 24 + 8*nhyp result scalars
  const double sn2 = std::exp(2.0 * log_sn), snu2 = 1.0e-6 * sn2;   // Core/inf.py:409-410
  stats_begin(h);
  h->has_post = false; h->has_fitc = false; h->pn = 0;
  cudaStream_t st = h->s_main;

  GPK_TRY(ensure(h, &h->dUin, &h->capUin, M * D));
  GPK_TRY(ensure(h, &h->dUs, &h->capUs, Mp * D));
  GPK_TRY(ensure(h, &h->fKuu, &h->cKuu, Mp * Mp));
  GPK_TRY(ensure_zero(h, &h->fDinvU, &h->cDinvU, Mp * NB));
  GPK_TRY(ensure(h, &h->fA2, &h->cA2, Mp * Mp));
  GPK_TRY(ensure_zero(h, &h->fDinv2, &h->cDinv2, Mp * NB));
  GPK_TRY(ensure(h, &h->fVt, &h->cVt, np * Mp));
  GPK_TRY(ensure(h, &h->fVs, &h->cVs, np * Mp));
  GPK_TRY(ensure(h, &h->dAlphaU, &h->capAlphaU, Mp));
  GPK_TRY(ensure(h, &h->dLpost, &h->capLpost, Mp * Mp));
  GPK_TRY(ensure(h, &h->dU, &h->capU, Mp * Mp));
  GPK_TRY(ensure(h, &h->dW, &h->capW, Mp * Mp));
  GPK_TRY(ensure(h, &h->dP, &h->capP, Mp * Mp));
  GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, Mp * Mp));
  const int nsplit = 8;
  const int64_t big = (np > Mp) ? np : Mp;
  // vector workspace: 8 x (np) | 6 x (Mp) | parts (2*Tm) | res (24 + 8*nhyp scalars) | partial sums
  const int64_t nres_max = 32 + 8 * (int64_t)nhyp;
  const int64_t part_need = ((int64_t)nsplit * big > 148 * 8) ? (int64_t)nsplit * big : 148 * 8;
  const int64_t vec_need = 8 * np + 6 * Mp + 2 * Tm + nres_max + part_need;
  GPK_TRY(ensure(h, &h->fVec, &h->cVec, vec_need));
  double* g = h->fVec;         // g_sn2
  double* rs = g + np;         // g^-1/2
  double* r = rs + np;         // (y-m) g^-1/2
  double* al = r + np;         // (Kt + sn2 I)^-1 (y-m)
  double* ww = al + np;        // colsum(W*W)
  double* va = ww + np;        // scratch (np)
  double* vb = va + np;        // scratch (np)
  double* vb2 = vb + np;       // scratch (np)
  double* tv = vb2 + np;       // (Mp) V r/sqrt(g)
  double* be = tv + Mp;
  double* x1 = be + Mp;        // Lu^-T be
  double* wk = x1 + Mp;        // solve work vector
  double* wv = wk + Mp;        // w = B al
  double* tw = wv + Mp;        // dKuu w
  double* partsU = tw + Mp;    // (Tm) log-det parts of Luu
  double* parts2 = partsU + Tm;
  double* res = parts2 + Tm;   // scalars
  double* part = res + nres_max;   // partial sums

  GPK_CK(h, cudaEventRecord(h->t0, st));
  std::memcpy(h->hPinned, scale.data(), D * sizeof(double));
  GPK_CK(h, cudaMemcpyAsync(h->dScale, h->hPinned, D * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  GPK_CK(h, cudaMemcpyAsync(h->dUin, U, (size_t)M * D * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->dR, 0, (size_t)np * sizeof(double), st));
  GPK_CK(h, cudaMemcpyAsync(h->dR, ymm, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  h->stats.h2d_bytes = (M * D + n + D) * (int64_t)sizeof(double);
  GPK_TRY(launch_prescale(h, st, h->dX, n, np, D, h->dScale, divide, premul, h->dXs));
  GPK_TRY(launch_prescale(h, st, h->dUin, M, Mp, D, h->dScale, divide, premul, h->dUs));

  CovArgs c{};
  c.D = D; c.kind = kind; c.matern_d = matern_d; c.epi = EPI_COV; c.sf2 = sf2; c.scale = 1.0;
  // 1. Kuu + snu2*I  (lower, identity padding)                                  Core/inf.py:412
  c.F = h->dUs; c.S = h->dUs; c.out = h->fKuu; c.ld = Mp; c.nF = M; c.nS = M; c.pF = Mp; c.pS = Mp;
  c.diag_add = snu2; c.same_set = 1; c.lower_only = 1; c.pad_identity = 1;
  GPK_TRY(launch_cov(h, st, c));
  // 3a. Ku transposed: (np x Mp), one data point per row                        Core/cov.py:366
  c.F = h->dXs; c.S = h->dUs; c.out = h->fVt; c.ld = np; c.nF = n; c.nS = M; c.pF = np; c.pS = Mp;
  c.diag_add = 0.0; c.same_set = 0; c.lower_only = 0; c.pad_identity = 0;
  GPK_TRY(launch_cov(h, st, c));
  GPK_CK(h, cudaEventRecord(h->t1, st));
  // 2. Luu
  GPK_TRY(potrf_device(h, h->fKuu, Mp, h->fDinvU, partsU, h->dInfo, nullptr, nullptr));
  // 3b. V = Luu^-1 Ku   as   Vt = Kut * Luu^-T                                   :413
  // Both O(M^2 n) products of the nlZ path run on the int8 tensor cores (error-free fp64 split, ozaki.cu) when there
  // are enough inducing points for it to pay; GPK_OZAKI_FITC=0 keeps them on DMMA.
  const bool fitc_oz = env_int("GPK_OZAKI", 1) && env_int("GPK_OZAKI_FITC", 1) && Tm >= 16;
  if (fitc_oz) GPK_TRY(sweep_forward_oz(h, st, h->fVt, np, Tn, h->fKuu, Mp, h->fDinvU, Tm));
  else GPK_TRY(sweep_forward(h, st, h->fVt, np, Tn, h->fKuu, Mp, h->fDinvU, Tm));
  // 4. g_sn2 = diagK + sn2 - colsum(V*V) ; r = (y-m)/sqrt(g)                      :415,418
  fitc_g_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(h->fVt, np, np, Mp, n, sf2 + sn2, h->dR, g, rs, r);
  // 5. Vs = diag(g^-1/2) V  (Mp x np)
  GPK_TRY(transpose_scale(h, st, h->fVt, np, np, Mp, rs, h->fVs, Mp));
  // 6. A2 = I + Vs Vs'  (lower)                                                   :417
  const bool sharded = (h->world > 1 && h->nccl_comm);
  if (!sharded || h->rank == 0) GPK_TRY(launch_set_identity(h, st, h->fA2, Mp, Mp, Mp));
  else GPK_CK(h, cudaMemsetAsync(h->fA2, 0, (size_t)Mp * Mp * sizeof(double), st));
  if (fitc_oz) {
    // SYRK over the data points in chunks of 16384 (int32 accumulators: 7 * 16384 * 2^14 < 2^31); every chunk is
    // sliced with its own row scales and added to A2 by the update kernel's epilogue
    const int64_t KC = 16384;
    GPK_TRY(oz_ensure(h, 0, Mp, (int)(np < KC ? np : KC)));
    for (int64_t c0 = 0; c0 < np; c0 += KC) {
      const int kw = (int)((np - c0 < KC) ? np - c0 : KC);
      GPK_TRY(launch_oz_slice(h, 0, st, h->fVs + c0 * Mp, Mp, (int)Mp, kw));
      GPK_TRY(launch_oz_ex(h, 0, st, h->fA2, Mp, (int)Mp, kw, 0, Tm, 0, 0, 0, /*C +=*/ 2));
    }
  } else {
    GemmArgs a{};
    a.A = h->fVs; a.B = h->fVs; a.C = h->fA2; a.lda = Mp; a.ldb = Mp; a.ldc = Mp; a.K = (int)np; a.tri = 1;
    GPK_TRY(launch_gemm_nt(h, st, 2, a, Tm, Tm));
  }
  GPK_TRY(dist_allreduce_sum(h, h->fA2, (size_t)Mp * Mp, st));          // the one big exchange: M x M partial
  // 7. Lu
  GPK_TRY(potrf_device(h, h->fA2, Mp, h->fDinv2, parts2, h->dInfo + 1, nullptr, nullptr));
  // 8. be = Lu^-1 (V r/sqrt(g)) = Lu^-1 (Vs r)                                    :419
  GPK_TRY(launch_rowdot(h, st, h->fVs, Mp, Mp, np, r, 0, 1.0, 0.0, part, nsplit, tv, Mp));
  GPK_TRY(dist_allreduce_sum(h, tv, (size_t)Mp, st));
  GPK_CK(h, cudaMemcpyAsync(wk, tv, (size_t)Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
  GPK_TRY(vec_fwd(h, st, h->fA2, Mp, h->fDinv2, wk, be, Tm));
  // 9. scalars of nlZ                                                             :428
  fitc_reduce_kernel<<<1, 1024, 0, st>>>(g, n, r, n, be, Mp, parts2, Tm, nullptr, nullptr, 0, res);
  if (sharded) {
    // res[0]=sum log g and res[1]=r'r are sums over the LOCAL data; res[2]=be'be and res[3]=logdet Lu are replicated
    res_n_kernel<<<1, 32, 0, st>>>(res, (double)n);                      // res[6] = local n
    if (h->rank != 0) scale_small_kernel<<<1, 32, 0, st>>>(res + 2, 2, 0.0);
    GPK_TRY(dist_allreduce_sum(h, res, 8, st));
  }
  // 10. post.alpha = Luu^-T Lu^-T be                                              :422
  GPK_CK(h, cudaMemcpyAsync(wk, be, (size_t)Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
  GPK_TRY(vec_bwd(h, st, h->fA2, Mp, h->fDinv2, wk, x1, Tm));
  GPK_CK(h, cudaMemcpyAsync(wk, x1, (size_t)Mp * sizeof(double), cudaMemcpyDeviceToDevice, st));
  GPK_TRY(vec_bwd(h, st, h->fKuu, Mp, h->fDinvU, wk, h->dAlphaU, Tm));
  GPK_CK(h, cudaEventRecord(h->t2, st));
  // 11. post.L = (Lu' Luu')^-1 (..)^-T - iKuu = (Uuu Uu)(Uuu Uu)' - Uuu Uuu'        :420,423
  GPK_TRY(inverse_factor_T(h, st, h->dU, h->fKuu, Mp, h->fDinvU));          // Uuu = Luu^-T (upper)
  GPK_TRY(inverse_factor_T(h, st, h->dW, h->fA2, Mp, h->fDinv2));           // Uu  = Lu^-T
  GPK_TRY(transpose_scale(h, st, h->dW, Mp, Mp, Mp, nullptr, h->dP, Mp));   // Wu = Uu' = Lu^-1 (lower)
  {
    GemmArgs a{};
    a.A = h->dU; a.B = h->dP; a.C = h->dTmp; a.lda = Mp; a.ldb = Mp; a.ldc = Mp; a.K = (int)Mp; a.tri = 0;
    GPK_TRY(launch_gemm_nt(h, st, 0, a, Tm, Tm));                            // Pm = Uuu * Uu
    GemmArgs b{};
    b.A = h->dTmp; b.B = h->dTmp; b.C = h->dW; b.lda = Mp; b.ldb = Mp; b.ldc = Mp; b.K = (int)Mp; b.tri = 1;
    GPK_TRY(launch_gemm_nt(h, st, 0, b, Tm, Tm));                            // Sigma (lower) = Pm Pm'
    GemmArgs k{};
    k.A = h->dU; k.B = h->dU; k.C = h->dP; k.lda = Mp; k.ldb = Mp; k.ldc = Mp; k.K = (int)Mp; k.tri = 2;
    GPK_TRY(launch_gemm_nt(h, st, 0, k, Tm, Tm));                            // iKuu (lower) = Uuu Uuu'
  }
  sym_diff_kernel<<<grid1(Mp * Mp), 256, 0, st>>>(h->dW, h->dP, Mp, Mp, h->dLpost, Mp);
  h->stats.launches += 3;
  GPK_CK(h, cudaGetLastError());
  GPK_CK(h, cudaEventRecord(h->t3, st));

  int nres = 8;
  if (want_der) {
    // ---- derivative block, Core/inf.py:429-451, in the data-major (np x Mp) layout -------------------------
    GPK_TRY(ensure(h, &h->fWt, &h->cWt, 5 * np * Mp));
    double* Bt = h->fWt;                // B' = (iKuu Ku)'            (np x Mp)
    double* Wt = Bt + np * Mp;          // W' = (Lu^-1 (V/g))'        (np x Mp)
    double* Gt = Wt + np * Mp;          // G' = ((B W') W)'           (np x Mp)
    double* Rt = Gt + np * Mp;          // dKu' then R'               (np x Mp)
    double* Wm = Rt + np * Mp;          // W                          (Mp x np)
    double* Bm = h->fVs;                // B (Mp x np); Vs is dead after step 8
    // al = r/sqrt(g) - V'(Lu^-T be)/g                                                   :431
    GPK_TRY(launch_rowdot(h, st, h->fVt, np, np, Mp, x1, 0, 1.0, 0.0, part, nsplit, va, np));
    fitc_al_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(h->dR, g, va, n, np, al);
    // B = iKuu Ku = Luu^-T V   ->   Bt = Vt * Uuu'                                      :432
    {
      GemmArgs a{};
      a.A = h->fVt; a.B = h->dU; a.C = Bt; a.lda = np; a.ldb = Mp; a.ldc = np; a.K = (int)Mp; a.tri = 0;
      if (fitc_oz) GPK_TRY(oz_gemm_nt(h, st, Bt, np, h->fVt, np, (int)np, h->dU, Mp, (int)Mp, (int)Mp, 1));
      else GPK_TRY(launch_gemm_nt(h, st, 0, a, Tn, Tm));
    }
    // w = B al                                                                           :433
    coldot_kernel<<<(unsigned)Mp, 256, 0, st>>>(Bt, np, np, al, nullptr, wv);
    GPK_TRY(dist_allreduce_sum(h, wv, (size_t)Mp, st));
    // W = Lu^-1 (V/g)   ->   Wt = diag(1/g) Vt * Lu^-T                                    :434
    rowscale_kernel<<<grid1(np * Mp), 256, 0, st>>>(h->fVt, np, np, Mp, rs, Wt);
    rowscale_kernel<<<grid1(np * Mp), 256, 0, st>>>(Wt, np, np, Mp, rs, Wt);
    if (fitc_oz) GPK_TRY(sweep_forward_oz(h, st, Wt, np, Tn, h->fA2, Mp, h->fDinv2, Tm));
    else GPK_TRY(sweep_forward(h, st, Wt, np, Tn, h->fA2, Mp, h->fDinv2, Tm));
    // ww = colsum(W*W) ; bb = colsum(B*B)
    rowdot2_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(Wt, Wt, np, np, Mp, 0, 0.0, ww);
    rowdot2_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(Bt, Bt, np, np, Mp, 0, 0.0, vb);
    // Q = B W' (Mp x Mp) needs both operands inducing-major; then G = Q W, used as sum(R .* G) = sum(RW' .* BW')
    GPK_TRY(transpose_scale(h, st, Bt, np, np, Mp, nullptr, Bm, Mp));
    GPK_TRY(transpose_scale(h, st, Wt, np, np, Mp, nullptr, Wm, Mp));
    {
      GemmArgs a{};
      a.A = Bm; a.B = Wm; a.C = h->dTmp; a.lda = Mp; a.ldb = Mp; a.ldc = Mp; a.K = (int)np; a.tri = 0;
      if (fitc_oz) GPK_TRY(oz_gemm_nt(h, st, h->dTmp, Mp, Bm, Mp, (int)Mp, Wm, Mp, (int)Mp, (int)np, 1));
      else GPK_TRY(launch_gemm_nt(h, st, 0, a, Tm, Tm));
      GPK_TRY(dist_allreduce_sum(h, h->dTmp, (size_t)Mp * Mp, st));        // Q = B W' summed over all data shards
      GemmArgs q{};
      q.A = Wt; q.B = h->dTmp; q.C = Gt; q.lda = np; q.ldb = Mp; q.ldc = np; q.K = (int)Mp; q.tri = 0;
      if (fitc_oz) GPK_TRY(oz_gemm_nt(h, st, Gt, np, Wt, np, (int)np, h->dTmp, Mp, (int)Mp, (int)Mp, 1));
      else GPK_TRY(launch_gemm_nt(h, st, 0, q, Tn, Tm));
    }
    // scalars shared by all hyper-parameters: res[8]=sum log g (unused) [9]=al'al [11]=sum ww [12]=sum 1/g
    fitc_reduce_kernel<<<1, 1024, 0, st>>>(g, n, al, n, nullptr, 0, ww, n, nullptr, nullptr, 0, res + 8);
    const int nbig = 148 * 8;
    double* bigpart = part;   // >= nbig entries
    // inducing-noise term of dnlZ.lik (:444-448): dKuu = 2 snu2 I, dKu = 0  =>  R = -2 snu2 B, v = 2 snu2 colsum(B*B)
    {
      double* out = res + 16;
      fitc_hyp_scalars_kernel<<<1, 1024, 0, st>>>(wv, wv, Mp, nullptr, al, ww, vb, 0.0, -1.0, n, out);
      bigdot_kernel<<<nbig, 256, 0, st>>>(Bt, Gt, np * Mp, bigpart);
      sum_finish_kernel<<<1, 1024, 0, st>>>(bigpart, nbig, out + 4);
    }
    h->stats.launches += 12;
    GPK_CK(h, cudaGetLastError());
    for (int ii = 0; ii < nhyp; ++ii) {
      double* out = res + 24 + 8 * ii;
      int epi, ard_dim = 0;
      if (kind == GPK_COV_RBFARD) { if (ii < D) { epi = EPI_DER_ARD; ard_dim = ii; } else epi = EPI_DER_SF; }
      else epi = (ii == 0) ? EPI_DER_ELL : EPI_DER_SF;
      const double ddiag = (epi == EPI_DER_SF) ? 2.0 * sf2 : 0.0;     // d k(x,x) / d hyp
      CovArgs d = c;
      d.epi = epi; d.ard_dim = ard_dim; d.scale = 1.0; d.diag_add = 0.0; d.lower_only = 0; d.pad_identity = 0;
      d.F = h->dUs; d.S = h->dUs; d.out = h->dW; d.ld = Mp; d.nF = M; d.nS = M; d.pF = Mp; d.pS = Mp; d.same_set = 1;
      GPK_TRY(launch_cov(h, st, d));                                   // dKuu (full, zero padded) -> dW
      d.F = h->dXs; d.S = h->dUs; d.out = Rt; d.ld = np; d.nF = n; d.nS = M; d.pF = np; d.pS = Mp; d.same_set = 0;
      GPK_TRY(launch_cov(h, st, d));                                   // dKu' (np x Mp) -> Rt
      GPK_TRY(launch_rowdot(h, st, Rt, np, np, Mp, wv, 0, 1.0, 0.0, part, nsplit, va, np));      // (dKu' w)
      GPK_TRY(launch_rowdot(h, st, h->dW, Mp, Mp, Mp, wv, 0, 1.0, 0.0, part, nsplit, tw, Mp));   // dKuu w
      {
        GemmArgs a{};                                                  // T' = B' dKuu  (dKuu symmetric)
        a.A = Bt; a.B = h->dW; a.C = h->fVt; a.lda = np; a.ldb = Mp; a.ldc = np; a.K = (int)Mp; a.tri = 0;
        // (V' is no longer needed once B', W' and al exist: its storage is scratch from here on)
        if (fitc_oz) GPK_TRY(oz_gemm_nt(h, st, h->fVt, np, Bt, np, (int)np, h->dW, Mp, (int)Mp, (int)Mp, 1));
        else GPK_TRY(launch_gemm_nt(h, st, 0, a, Tn, Tm));
      }
      fitc_r_kernel<<<grid1(np * Mp), 256, 0, st>>>(Rt, h->fVt, np * Mp);                         // fVt <- R'
      rowdot2_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(h->fVt, Bt, np, np, Mp, 0, 0.0, vb2);
      fitc_hyp_scalars_kernel<<<1, 1024, 0, st>>>(wv, tw, Mp, va, al, ww, vb2, ddiag, 1.0, n, out);
      bigdot_kernel<<<nbig, 256, 0, st>>>(h->fVt, Gt, np * Mp, bigpart);
      sum_finish_kernel<<<1, 1024, 0, st>>>(bigpart, nbig, out + 4);
      h->stats.launches += 5;
      GPK_CK(h, cudaGetLastError());
    }
    nres = 24 + 8 * nhyp;
    if (sharded) {
      // every scalar from res[8] on is a sum over local data EXCEPT the w'(dKuu w) entries (slot 0 of each 8-block)
      if (h->rank != 0) {
        scale_small_kernel<<<1, 32, 0, st>>>(res + 16, 1, 0.0);
        for (int ii = 0; ii < nhyp; ++ii) scale_small_kernel<<<1, 32, 0, st>>>(res + 24 + 8 * ii, 1, 0.0);
      }
      GPK_TRY(dist_allreduce_sum(h, res + 8, (size_t)(nres - 8), st));
    }
    GPK_CK(h, cudaMemcpyAsync(al_out, al, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  GPK_CK(h, cudaEventRecord(h->t4, st));

  GPK_CK(h, cudaMemcpyAsync(h->hPinned, res, nres * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(alpha, h->dAlphaU, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpy2DAsync(Lpost, (size_t)M * sizeof(double), h->dLpost, (size_t)Mp * sizeof(double),
                              (size_t)M * sizeof(double), (size_t)M, cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  h->stats.d2h_bytes = (M + M * M + nres + (want_der ? n : 0)) * (int64_t)sizeof(double);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t4); h->stats.total_ms = ms;
  cudaEventElapsedTime(&ms, h->t0, h->t1); h->stats.kbuild_ms = ms;
  cudaEventElapsedTime(&ms, h->t1, h->t2); h->stats.potrf_ms = ms;
  cudaEventElapsedTime(&ms, h->t2, h->t3); h->stats.solve_ms = ms;
  cudaEventElapsedTime(&ms, h->t3, h->t4); h->stats.deriv_ms = ms;

  const int* infos = reinterpret_cast<const int*>(h->hPinned + 2048);
  const double* R = h->hPinned;
  // nlZ = sum(log diag Lu) + (sum(log g) + n log 2pi + r'r - be'be)/2                    Core/inf.py:428
  const double n_total = (h->world > 1 && h->nccl_comm) ? R[6] : (double)n;
  *nlZ = R[3] + (R[0] + n_total * std::log(2.0 * M_PI) + R[1] - R[2]) / 2.0;
  if (want_der) {
    const double alal = R[9], sumww = R[11], suminvg = R[12];
    for (int ii = 0; ii < nhyp; ++ii) {
      const double* o = R + 24 + 8 * ii;
      int epi_sf = (kind == GPK_COV_RBFARD) ? (ii == D) : (ii == 1);
      const double ddiag = epi_sf ? 2.0 * sf2 : 0.0;
      dcov[ii] = (ddiag * suminvg + o[0] - 2.0 * o[1] - o[2] - o[3] - o[4]) / 2.0;        // :441-442
    }
    const double* o = R + 16;
    // R = -2 snu2 B  =>  sum(RW' .* BW') = -2 snu2 sum(B' .* G') ; w'(dKuu w) = 2 snu2 w'w
    dlik[0] = sn2 * (suminvg - sumww - alal)
              + (2.0 * snu2 * o[0] - 2.0 * snu2 * o[2] - 2.0 * snu2 * o[3] + 2.0 * snu2 * o[4]) / 2.0;
  }
  h->kind = kind; h->matern_d = matern_d; h->nhyp = nhyp; h->sn2 = sn2; h->sf2 = sf2;
  h->hyp.assign(hyp, hyp + nhyp);
  h->M = M; h->Mp = Mp;
  if (infos[0] != 0) return infos[0];
  if (infos[1] != 0) return infos[1];
  h->has_fitc = true;
  return 0;
}

int gpk_fitc_predict(gpk_handle hh, const double* Xs, int64_t ns, double* ks_alpha, double* fs2) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->has_fitc) return GPK_ERR_STATE;
  if (!Xs || ns <= 0 || !ks_alpha || !fs2) return GPK_ERR_ARG;
  const int D = h->D;
  const int64_t M = h->M, Mp = h->Mp;
  const int Tm = (int)(Mp / NB);
  cudaStream_t st = h->s_main;
  stats_begin(h);
  std::vector<double> scale;
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  GPK_TRY(kind_scale(h->kind, h->matern_d, h->hyp.data(), h->nhyp, D, scale, &divide, &premul, &sf2));
  const int64_t chunk = 16384;
  const int nsplit = 8;
  const int64_t cp_max = round_up(ns < chunk ? ns : chunk, NB);
  // Kst (cp x Mp) and Z = Kst * post.L (cp x Mp) in dXtmp; inputs, partials and outputs in dP
  GPK_TRY(ensure(h, &h->dXtmp, &h->capXtmp, 2 * cp_max * Mp));
  GPK_TRY(ensure(h, &h->dP, &h->capP, 2 * cp_max * D + (int64_t)nsplit * cp_max + 2 * cp_max));
  double* Kst = h->dXtmp;
  double* Z = h->dXtmp + cp_max * Mp;
  double* dXraw = h->dP;
  double* dXsc = dXraw + cp_max * D;
  double* dPart = dXsc + cp_max * D;
  double* dOut = dPart + (int64_t)nsplit * cp_max;
  GPK_CK(h, cudaEventRecord(h->t0, st));
  for (int64_t lo = 0; lo < ns; lo += chunk) {
    const int64_t m = (ns - lo < chunk) ? ns - lo : chunk;
    const int64_t mp = round_up(m, NB);
    GPK_CK(h, cudaMemcpyAsync(dXraw, Xs + lo * D, (size_t)m * D * sizeof(double), cudaMemcpyHostToDevice, st));
    GPK_TRY(launch_prescale(h, st, dXraw, m, mp, D, h->dScale, divide, premul, dXsc));
    CovArgs c{};
    c.F = dXsc; c.S = h->dUs; c.out = Kst; c.ld = mp; c.nF = m; c.nS = M; c.pF = mp; c.pS = Mp; c.D = D;
    c.kind = h->kind; c.matern_d = h->matern_d; c.epi = EPI_COV; c.sf2 = sf2; c.scale = 1.0;
    GPK_TRY(launch_cov(h, st, c));                                      // Ks' : K(xu, xs) transposed (Core/cov.py:369)
    GPK_TRY(launch_rowdot(h, st, Kst, mp, mp, Mp, h->dAlphaU, 0, 1.0, 0.0, dPart, nsplit, dOut, m));   // Ks' alpha
    GemmArgs a{};
    a.A = Kst; a.B = h->dLpost; a.C = Z; a.lda = mp; a.ldb = Mp; a.ldc = mp; a.K = (int)Mp; a.tri = 0;
    GPK_TRY(launch_gemm_nt(h, st, 0, a, (int)(mp / NB), Tm));           // Z = Ks' L  (L symmetric)
    // fs2 = max(kss + colsum(Ks .* (L Ks)), 0)                          Core/gp.py:418-419
    rowdot2_kernel<<<(unsigned)((mp + 127) / 128), 128, 0, st>>>(Kst, Z, mp, mp, Mp, 1, sf2, dOut + mp);
    h->stats.launches++;
    GPK_CK(h, cudaMemcpyAsync(ks_alpha + lo, dOut, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPK_CK(h, cudaMemcpyAsync(fs2 + lo, dOut + mp, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPK_CK(h, cudaStreamSynchronize(st));
  }
  GPK_CK(h, cudaEventRecord(h->t1, st));
  GPK_CK(h, cudaEventSynchronize(h->t1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t1);
  h->stats.total_ms = ms;
  h->stats.h2d_bytes = ns * D * (int64_t)sizeof(double);
  h->stats.d2h_bytes = 2 * ns * (int64_t)sizeof(double);
  return 0;
}

}  // extern "C"
