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
#include "tc_common.cuh"

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

// ---------------------------------------------------------------------------
// cov_tile_kernel - the covariance build the factorisation and the predictions use (EPI_COV, padded inputs).
//
// One 128x128 output tile per CTA (the tile of the blocked Cholesky: in lower-only mode whole tiles above the
// diagonal are never launched into work), 256 threads.  The two input blocks (128 points x D doubles each, contiguous
// rows of the padded (np,D) arrays) arrive by ONE cp.async.bulk each (TMA engine, mbarrier completion) and are
// transposed in shared memory to [d][point] so that the inner loop reads them with conflict-free 16-byte loads.
// A thread owns f = {2l, 2l+1, 64+2l, 64+2l+1} x 16 slow indices (4 passes of 4): 16 accumulators, every store is a
// 16-byte store and a warp's store instruction covers 512 contiguous bytes of one matrix column.
// exp(-y) is evaluated inline: y = n ln2/32 + r, exp = 2^(n/32) (1 + r q(r)) with a 32-entry table and a degree-5
// q - 12 fp64 operations instead of libdevice's ~25, max error 1.8 units of roundoff (the kernel is bound by the fp64
// pipe: 2 D operations per pair for the direct squared distance, then the exponential).
// ---------------------------------------------------------------------------
// 2^(j/32), j = 0..31, correctly rounded
__constant__ double c_exp2_tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467,
    1.4768261459394993, 1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,
    1.9152065613971474, 1.9571441241754002};

constexpr int FT = 128;                 // tile edge
constexpr int FT_LD = FT + 2;           // [d][point] pitch: rows stay 16-byte aligned
constexpr int FT_MAXD = 32;

__device__ __forceinline__ double exp_fast(double y, const double* __restrict__ tab) {
  // n = rint(y * 32/ln2) by the 1.5*2^52 trick; r = y - n ln2/32 (ln2/32 split so that n*hi is exact)
  const double t = fma(y, 46.16624130844683, 6755399441055744.0);
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -0.02166084938653512, y);
  r = fma(nf, -5.9631716539705866e-12, r);
  double q = 1.0 / 720.0;
  q = fma(q, r, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double tj = tab[n & 31];
  const double v = fma(tj, r * q, tj);
  const double sc = __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
  return (y < -700.0) ? 0.0 : sc;       // below 1e-304: flushed (the reference reaches denormals there)
}

template <int MAT /*0: exp(-d2/2) (RBF, RBFard); 1,3,5,7: Matern d*/>
__device__ __forceinline__ double cov_tile_value(double d2, double sf2, const double* tab) {
  if (MAT == 0) return sf2 * exp_fast(-0.5 * d2, tab);
  const double t = sqrt(d2);
  const double e = exp_fast(-t, tab);
  double f;
  if (MAT == 1) f = 1.0;
  else if (MAT == 3) f = 1.0 + t;
  else if (MAT == 5) f = 1.0 + t + t * t / 3.0;
  else f = 1.0 + t + 2.0 * t * t / 5.0 + t * t * t / 15.0;
  return sf2 * f * e;
}

template <int MAT>
__global__ void __launch_bounds__(256, 2) cov_tile_kernel(const CovArgs a) {
  extern __shared__ __align__(128) double csm[];
  __shared__ uint64_t bar;
  __shared__ double tab[32];
  const int D = a.D;
  double* land = csm;                         // [2][128*D] raw rows as they arrive (F block, S block)
  double* Fs = csm + 2 * FT * D;              // [D][FT_LD]
  double* Ss = Fs + D * FT_LD;                // [D][FT_LD]
  const int bf = blockIdx.x, bs = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f0 = (int64_t)bf * FT, s0 = (int64_t)bs * FT;
  // global index of the tile's first slow point (block-cyclic column ownership on multi-GPU runs)
  const int64_t gs0 = (a.s_bstride > 0) ? ((int64_t)bs * a.s_bstride + a.s_boff) * FT : s0;
  if (a.lower_only && f0 < gs0) return;       // tile entirely above the diagonal
  const bool same_tile = a.same_set && f0 == gs0;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const uint32_t bytes = (uint32_t)(FT * D * sizeof(double));
    mbar_expect_tx(&bar, same_tile ? bytes : 2 * bytes);
    bulk_g2s(land, a.F + f0 * D, bytes, &bar);
    if (!same_tile) bulk_g2s(land + FT * D, a.S + gs0 * D, bytes, &bar);
  }
  if (tid < 32) tab[tid] = c_exp2_tab[tid];
  __syncthreads();                            // barrier initialised, table written
  mbar_wait(&bar, 0);
  for (int idx = tid; idx < FT * D; idx += 256) {
    const int p = idx / D, d = idx - p * D;
    const double vf = land[idx];
    Fs[d * FT_LD + p] = vf;
    Ss[d * FT_LD + p] = same_tile ? vf : land[FT * D + idx];
  }
  __syncthreads();

  const int fa = 2 * lane;                    // this thread's fast indices: fa, fa+1, 64+fa, 64+fa+1
  const double sf2 = a.sf2, scale = a.scale;
  const bool interior = (f0 + FT <= a.nF) && (gs0 + FT <= a.nS) && (f0 + FT <= a.pF) && (s0 + FT <= a.pS) &&
                        (a.same_set ? (f0 != gs0) : true) && (a.lower_only ? (f0 > gs0) : true);
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const int sl = warp * 16 + q * 4;         // slow indices sl .. sl+3 (tile-local)
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const double2 f01 = *reinterpret_cast<const double2*>(Fs + d * FT_LD + fa);
      const double2 f23 = *reinterpret_cast<const double2*>(Fs + d * FT_LD + 64 + fa);
      const double2 s01 = *reinterpret_cast<const double2*>(Ss + d * FT_LD + sl);
      const double2 s23 = *reinterpret_cast<const double2*>(Ss + d * FT_LD + sl + 2);
      const double fv[4] = {f01.x, f01.y, f23.x, f23.y};
      const double sv[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double df = fv[i] - sv[j];
          acc[i][j] = fma(df, df, acc[i][j]);
        }
    }
    if (interior) {
      // the common case - a whole tile strictly below the diagonal, no padding: no per-entry masks at all
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double* col = a.out + (s0 + sl + j) * a.ld + f0 + fa;
        *reinterpret_cast<double2*>(col) = make_double2(cov_tile_value<MAT>(acc[0][j], sf2, tab) * scale,
                                                        cov_tile_value<MAT>(acc[1][j], sf2, tab) * scale);
        *reinterpret_cast<double2*>(col + 64) = make_double2(cov_tile_value<MAT>(acc[2][j], sf2, tab) * scale,
                                                             cov_tile_value<MAT>(acc[3][j], sf2, tab) * scale);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t s = s0 + sl + j, gs = gs0 + sl + j;
      if (s >= a.pS) continue;
      double v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t f = f0 + fa + (i & 1) + 64 * (i >> 1);
        double x;
        if (f < a.nF && gs < a.nS) {
          x = cov_tile_value<MAT>(acc[i][j], sf2, tab) * scale;
          if (a.same_set && f == gs) x += a.diag_add;
          if (a.lower_only && f < gs) x = 0.0;
        } else {
          x = (a.pad_identity && f == gs) ? 1.0 : 0.0;
        }
        v[i] = x;
      }
      double* col = a.out + s * a.ld + f0 + fa;
      if (f0 + FT <= a.pF) {
        *reinterpret_cast<double2*>(col) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(col + 64) = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (f0 + fa + (i & 1) + 64 * (i >> 1) < a.pF) col[(i & 1) + 64 * (i >> 1)] = v[i];
      }
    }
  }
}

static bool cov_tile_ok(const CovArgs& a) {
  return a.padded128 && a.epi == EPI_COV && a.D <= FT_MAXD && (a.ld % 2 == 0) &&
         ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && env_int("GPK_COV_TILE", 1) != 0;
}

template <int MAT>
static int launch_cov_tile(Handle* h, cudaStream_t st, const CovArgs& a) {
  const int64_t gf = (a.pF + FT - 1) / FT, gs = (a.pS + FT - 1) / FT;
  if (gs > 65535) return GPK_ERR_ARG;
  const size_t smem = (size_t)(2 * FT * a.D + 2 * a.D * FT_LD) * sizeof(double);
  GPK_SMEM_ATTR(h, cov_tile_kernel<MAT>, (size_t)(2 * FT * FT_MAXD + 2 * FT_MAXD * FT_LD) * sizeof(double));
  cov_tile_kernel<MAT><<<dim3((unsigned)gf, (unsigned)gs), 256, smem, st>>>(a);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
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
  if (a.prog) return launch_cov_prog(h, st, a);
  const int64_t gf = (a.pF + CT - 1) / CT, gs = (a.pS + CT - 1) / CT;
  if (gf <= 0 || gs <= 0) return 0;
  if (cov_tile_ok(a)) {
    if (a.kind != GPK_COV_MATERN) return launch_cov_tile<0>(h, st, a);
    switch (a.matern_d) {
      case 1: return launch_cov_tile<1>(h, st, a);
      case 3: return launch_cov_tile<3>(h, st, a);
      case 5: return launch_cov_tile<5>(h, st, a);
      default: return launch_cov_tile<7>(h, st, a);
    }
  }
  if (gs > 65535) return GPK_ERR_ARG;
  dim3 grid((unsigned)gf, (unsigned)gs);
  cov_kernel<<<grid, 256, 0, st>>>(a);
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk
