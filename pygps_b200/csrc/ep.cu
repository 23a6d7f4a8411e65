// ep.cu - inf.EP.evaluate with lik.Erf (binary GP classification, BASELINE config 5) on the GPU.
//
// Replaces /root/reference/pyGPs/Core/inf.py:731-806 (+ _epComputeParams :174-189) and the EP moments of
// lik.Erf (Core/lik.py:295-311 with cumGauss :328, gauOverCumGauss :341, logphi :354 - thresholds copied exactly).
//
// The site loop keeps the reference's FIXED sequential order (range(n), Core/inf.py:757-758).  Per site:
//   ep_site_kernel   (1 CTA): cavity, probit moments, new tilde parameters, gathers s = Sigma[:,i] from the
//                    lower-triangular storage and applies the O(n) update of mu that is algebraically the
//                    reference's `mu = Sigma*tnu`:  mu += s*(dtnu*(1-c*s_i) - c*mu_i)
//   ep_rank1_kernel  (grid) : Sigma -= c*s*s' on the lower triangle only (the reference's line :769, its "70%")
// After every sweep Sigma, mu and nlZ are rebuilt from scratch exactly like _epComputeParams, with the blocked
// Cholesky, the transposed multi-right-hand-side sweep and the DMMA SYRK of gemm_nt.cu.
#include <cmath>
#include <cstdlib>
#include "gpk_internal.cuh"

namespace gpk {

// ---------------------------------------------------------------- probit pieces (Core/lik.py:328-366)
__device__ __forceinline__ double d_logphi(double z, double p) {
  const double zmin = -6.2, zmax = -5.5;
  if (z > zmax) return log(p);
  const double asym = -log(M_PI) / 2.0 - z * z / 2.0 - log(sqrt(z * z / 2.0 + 2.0) - z / sqrt(2.0));
  if (z < zmin) return asym;
  const double lam = 1.0 / (1.0 + exp(25.0 * (0.5 - (z - zmin) / (zmax - zmin))));
  return (1.0 - lam) * asym + lam * log(p);
}

__device__ __forceinline__ double d_gau_over_cum(double f, double p) {
  const double naive = (exp(-f * f / 2.0) / sqrt(2.0 * M_PI)) / p;
  if (f > -5.0) return naive;
  const double bound = sqrt(f * f / 4.0 + 1.0) - f / 2.0;
  if (f < -6.0) return bound;
  const double lam = -5.0 - f;
  return (1.0 - lam) * naive + lam * bound;
}

// lZ, dlZ, d2lZ of int Phi(y f) N(f|mu,s2) df     (Core/lik.py:295-311); y is +-1
__device__ __forceinline__ void d_erf_moments(double y, double mu, double s2, double& lZ, double& dlZ, double& d2lZ) {
  const double z = mu / sqrt(1.0 + s2);
  const double yz = y * z;
  const double p = (1.0 + erf(yz / sqrt(2.0))) / 2.0;
  lZ = d_logphi(yz, p);
  const double n_p = d_gau_over_cum(yz, exp(lZ));
  dlZ = y * n_p / sqrt(1.0 + s2);
  d2lZ = -n_p * (yz + n_p) / (1.0 + s2);
}

__device__ __forceinline__ double d_sign1(double y) { return (y < 0.0) ? -1.0 : 1.0; }   // sign(y), 0 -> +1

constexpr unsigned FULLE = 0xffffffffu;
__device__ __forceinline__ double block_sum_e(double v, double* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLE, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < nw) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULLE, t, o);
  }
  return t;
}

// res[0] = -sum_i lZ(y_i, m_i, K_ii)                                               Core/inf.py:737
__global__ void __launch_bounds__(1024) ep_nlz0_kernel(const double* __restrict__ y, const double* __restrict__ m,
                                                       double kdiag, int64_t n, double* __restrict__ res) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double lZ, a, b;
    d_erf_moments(d_sign1(y[i]), m[i], kdiag, lZ, a, b);
    s -= lZ;
  }
  const double t = block_sum_e(s, sh);
  if (threadIdx.x == 0) res[0] = t;
}

// one EP site update (Core/inf.py:759-770); Sigma holds its lower triangle (pitch ld)
__global__ void __launch_bounds__(512) ep_site_kernel(const double* __restrict__ Sig, int64_t ld, int64_t n, int i,
                                                      const double* __restrict__ y, const double* __restrict__ m,
                                                      double* __restrict__ ttau, double* __restrict__ tnu,
                                                      double* __restrict__ mu, double* __restrict__ sbuf,
                                                      double* __restrict__ cbuf) {
  __shared__ double s_coef;
  if (threadIdx.x == 0) {
    const double Sii = Sig[i + (int64_t)i * ld];
    const double tau_ni = 1.0 / Sii - ttau[i];
    const double nu_ni = mu[i] / Sii + m[i] * tau_ni - tnu[i];
    double lZ, dlZ, d2lZ;
    d_erf_moments(d_sign1(y[i]), nu_ni / tau_ni, 1.0 / tau_ni, lZ, dlZ, d2lZ);
    const double ttau_old = ttau[i], tnu_old = tnu[i];
    double tt = -d2lZ / (1.0 + d2lZ / tau_ni);
    tt = fmax(tt, 0.0);
    const double tn = (dlZ + (m[i] - nu_ni / tau_ni) * d2lZ) / (1.0 + d2lZ / tau_ni);
    ttau[i] = tt;
    tnu[i] = tn;
    const double ds2 = tt - ttau_old;
    const double c = ds2 / (1.0 + ds2 * Sii);
    cbuf[0] = c;
    s_coef = (tn - tnu_old) * (1.0 - c * Sii) - c * mu[i];
  }
  __syncthreads();
  const double coef = s_coef;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
    const double s = (r >= i) ? Sig[r + (int64_t)i * ld] : Sig[i + r * ld];
    sbuf[r] = s;
    mu[r] += s * coef;
  }
}

// Blocked ("delayed update") form of the same site update.  Within a block of EPB consecutive sites the rank-1
// corrections are NOT applied to Sigma; site t of the block reconstructs its column of the current Sigma as
//     s = Sigma0[:, i] - S[:, 0:t] * w ,  w_j = c_j * S[i, j]
// (S holds the earlier columns of the block, c their coefficients), which is O(n*t) instead of O(n^2).  After the
// block, Sigma0 -= (S diag c) S' is ONE rank-EPB update on the tensor pipe.  Same arithmetic as :759-770 up to
// rounding; the per-sweep rebuild (_epComputeParams) bounds any drift exactly as in the reference.
// grid = ceil(n/256) CTAs; every CTA recomputes the (deterministic) site scalars, each thread owns one row.
constexpr int EPB = 128;
__global__ void ep_set_int_kernel(int* p, int v) { *p = v; }
// The site index is i = *i0p + t with t fixed per launch, so the EPB launches of a block are IDENTICAL from block to
// block and are replayed as one CUDA graph (the per-site cost is launch latency, not work); launches past the last
// site (ragged last block) return at once.
__global__ void __launch_bounds__(256) ep_site_blk_kernel(const double* __restrict__ Sig, int64_t ld, int64_t n,
                                                          const int* __restrict__ i0p, int t,
                                                          const double* __restrict__ y,
                                                          const double* __restrict__ m,
                                                          const double* __restrict__ ttau,
                                                          const double* __restrict__ tnu, double* __restrict__ mu,
                                                          double* __restrict__ S, double* __restrict__ Sc,
                                                          double* __restrict__ cvec, double* __restrict__ mun,
                                                          double* __restrict__ ttau_new, double* __restrict__ tnu_new) {
  // ttau / tnu are READ by every CTA of this launch (each recomputes the site scalars); the new site parameters go to
  // the staging vectors ttau_new / tnu_new and are committed once per sweep - a CTA scheduled late can therefore never
  // see the values CTA 0 writes at the end of the same launch.
  const int i = *i0p + t;
  if (i >= n) return;
  const double* mu_in = mun + (i & 1);
  double* mu_out = mun + ((i + 1) & 1);
  __shared__ double w[EPB];
  __shared__ double sh[8];
  __shared__ double s_c, s_coef, s_tt, s_tn;
  const int tid = threadIdx.x;
  // w_j = c_j * S[i, j]  and  Sii = Sigma0[i,i] - sum_j w_j S[i,j]
  double part = 0.0;
  if (tid < t) {
    const double sij = S[i + (int64_t)tid * ld];
    const double wj = cvec[tid] * sij;
    w[tid] = wj;
    part = wj * sij;
  }
  // deterministic block sum of `part` over the first EPB threads
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) sh[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    const double corr = (sh[0] + sh[1]) + (sh[2] + sh[3]);
    const double Sii = Sig[i + (int64_t)i * ld] - corr;
    const double mui = mu_in[0];
    const double tau_ni = 1.0 / Sii - ttau[i];
    const double nu_ni = mui / Sii + m[i] * tau_ni - tnu[i];
    double lZ, dlZ, d2lZ;
    d_erf_moments(d_sign1(y[i]), nu_ni / tau_ni, 1.0 / tau_ni, lZ, dlZ, d2lZ);
    const double ttau_old = ttau[i], tnu_old = tnu[i];
    double tt = -d2lZ / (1.0 + d2lZ / tau_ni);
    tt = fmax(tt, 0.0);
    const double tn = (dlZ + (m[i] - nu_ni / tau_ni) * d2lZ) / (1.0 + d2lZ / tau_ni);
    const double ds2 = tt - ttau_old;
    const double c = ds2 / (1.0 + ds2 * Sii);
    s_c = c; s_tt = tt; s_tn = tn;
    s_coef = (tn - tnu_old) * (1.0 - c * Sii) - c * mui;
  }
  __syncthreads();
  const double c = s_c, coef = s_coef;
  const int64_t r = (int64_t)blockIdx.x * 256 + tid;
  if (r < n) {
    double s = (r >= i) ? Sig[r + (int64_t)i * ld] : Sig[i + r * ld];
    // eight loads in flight per thread: this loop is pure L2 latency (one launch per site, 32 CTAs at n = 8192)
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.0;
    int j = 0;
    for (; j + 8 <= t; j += 8) {
      double x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = S[r + (int64_t)(j + q) * ld];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(x[q], w[j + q], a[q]);
    }
    double tail = 0.0;
    for (; j < t; ++j) tail = fma(S[r + (int64_t)j * ld], w[j], tail);
    s -= (((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))) + tail;
    S[r + (int64_t)t * ld] = s;
    Sc[r + (int64_t)t * ld] = c * s;
    const double munew = mu[r] + s * coef;
    mu[r] = munew;
    if (r == i + 1) mu_out[0] = munew;           // mu of the NEXT site, read by its launch (two alternating slots)
  }
  if (blockIdx.x == 0 && tid == 0) {
    ttau_new[i] = s_tt;
    tnu_new[i] = s_tn;
    cvec[t] = c;
  }
}

// Sigma -= c * s s'   on the lower triangle (64x64 tiles)
__global__ void __launch_bounds__(256) ep_rank1_kernel(double* __restrict__ Sig, int64_t ld, int64_t n,
                                                       const double* __restrict__ sbuf,
                                                       const double* __restrict__ cbuf) {
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bi < bj) return;
  const double c = cbuf[0];
  if (c == 0.0) return;
  __shared__ double sr[64], sc[64];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)bi * 64, c0 = (int64_t)bj * 64;
  if (tid < 64) sr[tid] = (r0 + tid < n) ? sbuf[r0 + tid] : 0.0;
  else if (tid < 128) sc[tid - 64] = (c0 + tid - 64 < n) ? sbuf[c0 + tid - 64] * c : 0.0;
  __syncthreads();
  const int tr = tid & 63, tc0 = tid >> 6;
  const int64_t r = r0 + tr;
  if (r >= n) return;
#pragma unroll 4
  for (int cc = tc0; cc < 64; cc += 4) {
    const int64_t col = c0 + cc;
    if (col < n && r >= col) Sig[r + col * ld] -= sr[tr] * sc[cc];
  }
}

// B (lower, identity padding) = I + ssi ssi' .* K    ; ssi = sqrt(ttau)             Core/inf.py:176-178
__global__ void ep_build_b_kernel(const double* __restrict__ K, const double* __restrict__ ttau, int64_t ld, int64_t n,
                                  int64_t np, double* __restrict__ B) {
  const int64_t total = np * np;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % np, c = k / np;
    if (r < c) continue;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < n && c < n) v += sqrt(ttau[r]) * sqrt(ttau[c]) * K[r + c * ld];
    B[r + c * ld] = v;
  }
}

// P[r,c] = K[r,c] * sqrt(ttau[c])   (the transposed right-hand sides of V = L^-1 (ssi .* K))      :179
__global__ void ep_colscale_kernel(const double* __restrict__ K, const double* __restrict__ ttau, int64_t ld, int64_t n,
                                   int64_t np, double* __restrict__ P) {
  const int64_t total = np * np;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % np, c = k / np;
    P[k] = (r < n && c < n) ? K[r + c * ld] * sqrt(ttau[c]) : 0.0;
  }
}

// upper <- lower (full symmetric matrix from its lower triangle)
__global__ void mirror_kernel(double* __restrict__ A, int64_t ld, int64_t n) {
  const int64_t total = n * n;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % n, c = k / n;
    if (r < c) A[r + c * ld] = A[c + r * ld];
  }
}

// nlZ pieces of _epComputeParams (:182-188) and the cavity vectors; res[0] = nlZ - sum(log diag L)
//   with_m = 1: nu_n includes + m*tau_n (:184); with_m = 0: the derivative block's variant (:787)
__global__ void __launch_bounds__(1024) ep_terms_kernel(const double* __restrict__ Sig, int64_t ld, int64_t n,
                                                        const double* __restrict__ mu, const double* __restrict__ ttau,
                                                        const double* __restrict__ tnu, const double* __restrict__ m,
                                                        const double* __restrict__ y, int with_m,
                                                        double* __restrict__ dlz_out, double* __restrict__ res) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double Ds = Sig[i + i * ld];
    const double tau_n = 1.0 / Ds - ttau[i];
    const double nu_n = mu[i] / Ds - tnu[i] + (with_m ? m[i] * tau_n : 0.0);
    double lZ, dlZ, d2lZ;
    d_erf_moments(d_sign1(y[i]), nu_n / tau_n, 1.0 / tau_n, lZ, dlZ, d2lZ);
    if (dlz_out) dlz_out[i] = dlZ;
    const double a = nu_n - m[i] * tau_n;
    s += -lZ - tnu[i] * mu[i] / 2.0 - a * ((ttau[i] / tau_n * a - 2.0 * tnu[i]) / (ttau[i] + tau_n)) / 2.0 +
         tnu[i] * tnu[i] / (tau_n + ttau[i]) / 2.0 - log(1.0 + ttau[i] / tau_n) / 2.0;
  }
  const double t = block_sum_e(s, sh);
  if (threadIdx.x == 0) res[0] = t;
}

// v = sW .* (K tnu) ; sW = sqrt(ttau)
__global__ void ep_sw_mul_kernel(const double* __restrict__ ttau, const double* __restrict__ ktnu, int64_t np, int64_t n,
                                 double* __restrict__ sw, double* __restrict__ v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  const double s = (i < n) ? sqrt(ttau[i]) : 0.0;
  sw[i] = s;
  v[i] = s * ((i < n) ? ktnu[i] : 0.0);
}
// alpha = tnu - sW .* w                                                           Core/inf.py:777
__global__ void ep_alpha_kernel(const double* __restrict__ tnu, const double* __restrict__ sw,
                                const double* __restrict__ w, int64_t np, int64_t n, double* __restrict__ alpha) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= np) return;
  alpha[i] = (i < n) ? tnu[i] - sw[i] * w[i] : 0.0;
}

// P[c, r] *= s[r]  for a (rows x cols) column-major matrix: scale COLUMN r by s[r]
__global__ void colscale_inplace_kernel(double* __restrict__ P, int64_t ld, int64_t rows, int64_t cols,
                                        const double* __restrict__ s) {
  const int64_t total = rows * cols;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k % rows, c = k / rows;
    P[r + c * ld] *= s[c];
  }
}

static inline int grid_e(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_colscale_inplace(Handle* h, cudaStream_t st, double* P, int64_t ld, int64_t rows, int64_t cols,
                            const double* s) {
  colscale_inplace_kernel<<<grid_e(rows * cols), 256, 0, st>>>(P, ld, rows, cols, s);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

struct EpBuf {
  double *K, *Sig, *P;                        // (np x np)
  double *y, *m, *ttau, *tnu, *mu, *sbuf, *ktnu, *v, *w, *wk, *dlz, *parts, *res, *part, *cbuf;
};

// Sigma, mu, nlZ from (ttau, tnu): _epComputeParams.  Leaves the factor of B in dA/dDinv.  nlZ -> *nlz_host.
static int ep_compute_params(Handle* h, cudaStream_t st, const EpBuf& b, int64_t n, int64_t np, double* nlz_host,
                             int* info_host) {
  const int T = (int)(np / NB);
  const int nsplit = 8;
  GPK_CK(h, cudaMemsetAsync(h->dInfo, 0, 4 * sizeof(int), st));
  ep_build_b_kernel<<<grid_e(np * np), 256, 0, st>>>(b.K, b.ttau, np, n, np, h->dA);
  GPK_TRY(potrf_device(h, h->dA, np, h->dDinv, b.parts, h->dInfo, nullptr, nullptr));
  ep_colscale_kernel<<<grid_e(np * np), 256, 0, st>>>(b.K, b.ttau, np, n, np, b.P);
  // the two O(n^3) products of the rebuild run on the int8 tensor cores (error-free fp64 split, ozaki.cu) when the
  // problem is large enough for it to pay and the int32 accumulators cannot overflow; GPK_OZAKI_EP=0: fp64 DMMA
  const bool ep_oz = env_int("GPK_OZAKI", 1) && env_int("GPK_OZAKI_EP", 1) && T >= 16 && np < 18432;
  if (ep_oz) GPK_TRY(sweep_forward_oz(h, st, b.P, np, T, h->dA, np, h->dDinv, T)); // P <- V'
  else GPK_TRY(sweep_forward(h, st, b.P, np, T, h->dA, np, h->dDinv, T));
  GPK_CK(h, cudaMemcpyAsync(b.Sig, b.K, (size_t)np * np * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (ep_oz) {
    GPK_TRY(oz_ensure(h, 1, np, (int)np));
    GPK_TRY(launch_oz_slice(h, 1, st, b.P, np, (int)np, (int)np));
    GPK_TRY(launch_oz_syrk(h, 1, st, b.Sig, np, (int)np, (int)np, 0, T));          // Sigma = K - V'V (lower)
  } else {
    GemmArgs a{};
    a.A = b.P; a.B = b.P; a.C = b.Sig; a.lda = np; a.ldb = np; a.ldc = np; a.K = (int)np; a.tri = 1;
    GPK_TRY(launch_gemm_nt(h, st, 1, a, T, T));                                      // Sigma = K - V'V (lower)
  }
  mirror_kernel<<<grid_e(np * np), 256, 0, st>>>(b.Sig, np, np);
  GPK_TRY(launch_rowdot(h, st, b.Sig, np, np, np, b.tnu, 0, 1.0, 0.0, b.part, nsplit, b.mu, np));   // mu = Sigma tnu
  ep_terms_kernel<<<1, 1024, 0, st>>>(b.Sig, np, n, b.mu, b.ttau, b.tnu, b.m, b.y, 1, nullptr, b.res);
  GPK_TRY(launch_sum_parts(h, st, b.parts, T, b.res + 1));
  h->stats.launches += 5;
  GPK_CK(h, cudaGetLastError());
  GPK_CK(h, cudaMemcpyAsync(h->hPinned, b.res, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 2048, h->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  *nlz_host = h->hPinned[0] + h->hPinned[1];
  *info_host = *reinterpret_cast<int*>(h->hPinned + 2048);
  return 0;
}

}  // namespace gpk

using namespace gpk;

extern "C" {

int gpk_ep_eval(gpk_handle hh, int kind, int matern_d, const double* hyp, int nhyp, const double* mvec,
                const double* y, double* ttau_io, double* tnu_io, int use_last, int want_der, double* nlZ,
                double* alpha, double* sW, double* dcov, double* dlz_out, int* sweeps_out) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!h->dX || h->n <= 0) return GPK_ERR_STATE;
  if (!hyp || !mvec || !y || !ttau_io || !tnu_io || !nlZ || !alpha || !sW) return GPK_ERR_ARG;
  if (want_der && (!dcov || !dlz_out)) return GPK_ERR_ARG;
  const int64_t n = h->n, np = h->np;
  const int D = h->D, T = (int)(np / NB);
  std::vector<double> scale;
  int divide = 0;
  double premul = 1.0, sf2 = 1.0;
  GPK_TRY(kind_scale(kind, matern_d, hyp, nhyp, D, scale, &divide, &premul, &sf2));
  if (D > 1900) return GPK_ERR_ARG;
  stats_begin(h);
  h->has_post = false; h->has_fitc = false; h->pn = 0; h->post_ep = false;
  cudaStream_t st = h->s_main;
  const double tol = 1e-4;
  const int max_sweep = 10, min_sweep = 2;          // Core/inf.py:732

  GPK_TRY(ensure(h, &h->dA, &h->capA, np * np));
  GPK_TRY(ensure(h, &h->eK, &h->ceK, np * np));
  GPK_TRY(ensure(h, &h->eSig, &h->ceSig, np * np));
  GPK_TRY(ensure(h, &h->dP, &h->capP, np * np));
  const int nsplit = 8;
  GPK_TRY(ensure(h, &h->eVec, &h->ceVec, 14 * np + T + 128 + 4 * (int64_t)D + (int64_t)nsplit * np + 256));
  GPK_TRY(ensure(h, &h->dU, &h->capU, 2 * np * EPB));
  EpBuf b;
  b.K = h->eK; b.Sig = h->eSig; b.P = h->dP;
  double* v0 = h->eVec;
  b.y = v0; b.m = v0 + np; b.ttau = v0 + 2 * np; b.tnu = v0 + 3 * np; b.mu = v0 + 4 * np; b.sbuf = v0 + 5 * np;
  b.ktnu = v0 + 6 * np; b.v = v0 + 7 * np; b.w = v0 + 8 * np; b.wk = v0 + 9 * np; b.dlz = v0 + 10 * np;
  double* sw = v0 + 11 * np;
  double* ttau_new = v0 + 12 * np;   // staging of the site parameters of the running sweep (blocked path)
  double* tnu_new = v0 + 13 * np;
  // res: [0,1] nlZ pieces, [8] nlZ0, [12] scratch, [14] cbuf, [16..) dnlZ results (+ ARD scratch)
  b.parts = v0 + 14 * np; b.res = b.parts + T; b.cbuf = b.res + 14; b.part = b.res + 128 + 4 * D;
  double* cvec = b.part + (int64_t)nsplit * np;   // EPB coefficients of the current block
  double* mun = cvec + EPB;                       // two alternating slots: mu of the next site
  const bool naive = getenv("GPK_EP_NAIVE") != nullptr;   // the unblocked rank-1 form, kept for A/B checks

  GPK_CK(h, cudaEventRecord(h->t0, st));
  std::memcpy(h->hPinned, scale.data(), D * sizeof(double));
  GPK_CK(h, cudaMemcpyAsync(h->dScale, h->hPinned, D * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemsetAsync(h->eVec, 0, (size_t)(14 * np) * sizeof(double), st));
  GPK_CK(h, cudaMemcpyAsync(b.y, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_CK(h, cudaMemcpyAsync(b.m, mvec, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  GPK_TRY(launch_prescale(h, st, h->dX, n, np, D, h->dScale, divide, premul, h->dXs));
  {
    CovArgs c{};
    c.F = h->dXs; c.S = h->dXs; c.out = b.K; c.ld = np; c.nF = n; c.nS = n; c.pF = np; c.pS = np; c.D = D;
    c.kind = kind; c.matern_d = matern_d; c.epi = EPI_COV; c.sf2 = sf2; c.scale = 1.0; c.same_set = 1;
    GPK_TRY(launch_cov(h, st, c));
  }
  GPK_CK(h, cudaEventRecord(h->t1, st));
  // nlZ0 = -sum lZ(y, m, diag K); diag K = sf2 for all three kernels
  ep_nlz0_kernel<<<1, 1024, 0, st>>>(b.y, b.m, sf2, n, b.res + 8);
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 16, b.res + 8, sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  const double nlZ0 = h->hPinned[16];
  double nlz = nlZ0;
  int info = 0;
  bool zero_start = true;
  if (use_last) {
    GPK_CK(h, cudaMemcpyAsync(b.ttau, ttau_io, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    GPK_CK(h, cudaMemcpyAsync(b.tnu, tnu_io, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    double cand = 0.0;
    GPK_TRY(ep_compute_params(h, st, b, n, np, &cand, &info));
    if (info == 0 && !(cand > nlZ0)) { nlz = cand; zero_start = false; }   // keep the warm start (Core/inf.py:746-753)
  }
  if (zero_start) {
    GPK_CK(h, cudaMemsetAsync(b.ttau, 0, (size_t)np * sizeof(double), st));
    GPK_CK(h, cudaMemsetAsync(b.tnu, 0, (size_t)np * sizeof(double), st));
    GPK_CK(h, cudaMemsetAsync(b.mu, 0, (size_t)np * sizeof(double), st));
    GPK_CK(h, cudaMemcpyAsync(b.Sig, b.K, (size_t)np * np * sizeof(double), cudaMemcpyDeviceToDevice, st));
    nlz = nlZ0;
  }
  double nlz_old = INFINITY;
  int sweep = 0;
  const int g64 = (int)((n + 63) / 64);
  while ((std::fabs(nlz - nlz_old) > tol && sweep < max_sweep) || sweep < min_sweep) {
    nlz_old = nlz;
    ++sweep;
    if (naive) {
      for (int i = 0; i < (int)n; ++i) {
        ep_site_kernel<<<1, 512, 0, st>>>(b.Sig, np, n, i, b.y, b.m, b.ttau, b.tnu, b.mu, b.sbuf, b.cbuf);
        ep_rank1_kernel<<<dim3(g64, g64), 256, 0, st>>>(b.Sig, np, n, b.sbuf, b.cbuf);
      }
      h->stats.launches += 2 * n;
    } else {
      double* S = h->dU;
      double* Sc = h->dU + np * EPB;
      GPK_CK(h, cudaMemcpyAsync(mun, b.mu, sizeof(double), cudaMemcpyDeviceToDevice, st));   // slot 0 = mu[0]
      const unsigned gsite = (unsigned)((n + 255) / 256);
      if (!h->dFlags) {
        GPK_CK(h, cudaMalloc((void**)&h->dFlags, 1024 * sizeof(int)));
        GPK_CK(h, cudaMemsetAsync(h->dFlags, 0, 1024 * sizeof(int), st));
        h->flag_epoch = 0;
      }
      int* i0p = h->dFlags + 1000;                // (the first T entries are the backward substitution's ready flags)
      // the EPB site launches of one block as a graph, rebuilt only when a buffer or the size changes
      const void* sig[8] = {b.Sig, b.y, i0p, b.ttau, b.mu, S, cvec, (const void*)(intptr_t)n};
      const bool use_graph = env_int("GPK_EP_GRAPH", 1) != 0;
      if (use_graph && (!h->epGraphExec || std::memcmp(sig, h->epGraphSig, sizeof(sig)) != 0)) {
        if (h->epGraphExec) { cudaGraphExecDestroy((cudaGraphExec_t)h->epGraphExec); h->epGraphExec = nullptr; }
        cudaGraph_t g = nullptr;
        GPK_CK(h, cudaStreamSynchronize(st));
        GPK_CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int t = 0; t < EPB; ++t)
          ep_site_blk_kernel<<<gsite, 256, 0, st>>>(b.Sig, np, n, i0p, t, b.y, b.m, b.ttau, b.tnu, b.mu, S, Sc, cvec, mun, ttau_new, tnu_new);
        GPK_CK(h, cudaStreamEndCapture(st, &g));
        cudaGraphExec_t ge = nullptr;
        GPK_CK(h, cudaGraphInstantiate(&ge, g, 0));
        cudaGraphDestroy(g);
        h->epGraphExec = ge;
        std::memcpy(h->epGraphSig, sig, sizeof(sig));
      }
      for (int i0 = 0; i0 < (int)n; i0 += EPB) {
        const int bl = ((int)n - i0 < EPB) ? (int)n - i0 : EPB;
        GPK_CK(h, cudaMemsetAsync(S, 0, (size_t)(2 * np * EPB) * sizeof(double), st));
        ep_set_int_kernel<<<1, 1, 0, st>>>(i0p, i0);
        if (use_graph) {
          GPK_CK(h, cudaGraphLaunch((cudaGraphExec_t)h->epGraphExec, st));
        } else {
          for (int t = 0; t < bl; ++t)
            ep_site_blk_kernel<<<gsite, 256, 0, st>>>(b.Sig, np, n, i0p, t, b.y, b.m, b.ttau, b.tnu, b.mu, S, Sc, cvec, mun, ttau_new, tnu_new);
        }
        GemmArgs a{};                                 // Sigma0 -= (S diag c) S'  on the lower triangle
        a.A = Sc; a.B = S; a.C = b.Sig; a.lda = np; a.ldb = np; a.ldc = np; a.K = EPB; a.tri = 1;
        GPK_TRY(launch_gemm_nt(h, st, 1, a, T, T));
        h->stats.launches += bl;
      }
      // every site has been visited once: commit the sweep's site parameters
      GPK_CK(h, cudaMemcpyAsync(b.ttau, ttau_new, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      GPK_CK(h, cudaMemcpyAsync(b.tnu, tnu_new, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    GPK_CK(h, cudaGetLastError());
    GPK_TRY(ep_compute_params(h, st, b, n, np, &nlz, &info));
    if (info != 0) return info;
    if (!std::isfinite(nlz)) break;
  }
  GPK_CK(h, cudaEventRecord(h->t2, st));
  // posterior: sW = sqrt(ttau), alpha = tnu - sW .* solve_chol(L, sW .* (K tnu))          Core/inf.py:777
  GPK_TRY(launch_rowdot(h, st, b.K, np, np, np, b.tnu, 0, 1.0, 0.0, b.part, nsplit, b.ktnu, np));
  ep_sw_mul_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(b.ttau, b.ktnu, np, n, sw, b.v);
  for (int k = 0; k < T; ++k) GPK_TRY(launch_trsv_fwd(h, st, h->dA, np, h->dDinv, b.v, b.wk, k, T));
  for (int k = T - 1; k >= 0; --k) GPK_TRY(launch_trsv_bwd(h, st, h->dA, np, h->dDinv, b.wk, b.w, k, T));
  ep_alpha_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(b.tnu, sw, b.w, np, n, h->dAlpha);
  GPK_CK(h, cudaEventRecord(h->t3, st));
  if (want_der) {
    // F = alpha alpha' - sW sW' .* B^-1 ; dnlZ.cov = -sum(F .* dK)/2                      :788-792
    GPK_TRY(ensure(h, &h->dU, &h->capU, np * np));
    GPK_TRY(ensure(h, &h->dW, &h->capW, np * np));
    const int64_t g = (n + 63) / 64;
    GPK_TRY(ensure(h, &h->dTmp, &h->capTmp, g * g * 34));
    GPK_TRY(inverse_factor_T(h, st, h->dU, h->dA, np, h->dDinv));
    GemmArgs v{};
    v.A = h->dU; v.B = h->dU; v.C = h->dW; v.lda = np; v.ldb = np; v.ldc = np; v.K = (int)np; v.tri = 2;
    GPK_TRY(launch_gemm_nt(h, st, 0, v, T, T));
    GPK_TRY(launch_dnlz_sw(h, st, h->dXs, n, D, h->dW, np, h->dAlpha, sw, sf2, kind, matern_d, h->dTmp, h->capTmp,
                           b.res + 16));
    // dlZ for dnlZ.mean with the derivative block's cavity (no +m*tau_n term, :787)
    ep_terms_kernel<<<1, 1024, 0, st>>>(b.Sig, np, n, b.mu, b.ttau, b.tnu, b.m, b.y, 0, b.dlz, b.res + 12);
    GPK_CK(h, cudaMemcpyAsync(dlz_out, b.dlz, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  GPK_CK(h, cudaEventRecord(h->t4, st));
  GPK_CK(h, cudaMemcpyAsync(h->hPinned + 32, b.res + 16, (nhyp + 2) * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(alpha, h->dAlpha, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(sW, sw, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(ttau_io, b.ttau, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaMemcpyAsync(tnu_io, b.tnu, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  GPK_CK(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->t0, h->t4); h->stats.total_ms = ms;
  cudaEventElapsedTime(&ms, h->t0, h->t1); h->stats.kbuild_ms = ms;
  cudaEventElapsedTime(&ms, h->t1, h->t2); h->stats.potrf_ms = ms;   // sweeps + per-sweep refactorisation
  cudaEventElapsedTime(&ms, h->t2, h->t3); h->stats.solve_ms = ms;
  cudaEventElapsedTime(&ms, h->t3, h->t4); h->stats.deriv_ms = ms;
  h->stats.h2d_bytes = (4 * n + D) * (int64_t)sizeof(double);
  h->stats.d2h_bytes = (4 * n + (want_der ? n : 0)) * (int64_t)sizeof(double);
  *nlZ = nlz;
  if (sweeps_out) *sweeps_out = sweep;
  if (want_der)
    for (int i = 0; i < nhyp; ++i) dcov[i] = h->hPinned[32 + i] / 2.0;
  h->kind = kind; h->matern_d = matern_d; h->nhyp = nhyp; h->sn2 = 1.0; h->sf2 = sf2;
  h->hyp.assign(hyp, hyp + nhyp);
  // keep sW on the handle for predict
  GPK_TRY(ensure(h, &h->eSW, &h->ceSW, np));
  GPK_CK(h, cudaMemcpy(h->eSW, sw, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice));
  h->has_post = true;
  h->post_ep = true;
  return 0;
}

}  // extern "C"
