// gemm_nt.cu - the hot kernel: fp64 C (+)= A * B^T on the tensor pipe (DMMA).
//
// Every O(N^3) step of the path is this one kernel:
//   * trailing SYRK/GEMM update of the right-looking Cholesky  (replaces the
//     dpotrf call in jitchol, /root/reference/pyGPs/Core/tools.py:61)
//   * the panel TRSM, done as a product with the inverted diagonal block
//   * the triangular solves with many right-hand sides (predict, Core/gp.py:415;
//     solve_chol(L, eye(n)), Core/inf.py:373)
//
// tcgen05.mma has no f64 kind, so the fp64 contraction is issued as warp-level
// mma.sync.m16n8k8.f64 (SASS: DMMA).  Operands are staged global->shared with
// cp.async 16-byte copies through a 4-stage ring; the shared layout is the
// global one (column-major slabs of BK=16 columns) with a pitch of 132 doubles,
// which makes every A/B fragment load bank-conflict free (pitch == 4 mod 16).
//
// Tile: 128x64 per CTA (default), 8 warps as 4(M) x 2(N), warp tile 32x32 = 2x4
// MMA tiles, 32 fp64 accumulators per thread, two CTAs per SM; a 128x128 /
// one-CTA-per-SM variant is kept for comparison (GPK_GEMM_BN=128).
#include <cstdlib>
#include "gpk_internal.cuh"

namespace gpk {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// D(16x8) += A(16x8,row) * B(8x8,col).  Fragment layout (lane = 4*g + t):
//   a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4) ; b0:(k=t,n=g) b1:(k=t+4,n=g)
//   c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], double b0, double b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}

// MODE 0: C = A*B^T          (C is not read)
// MODE 1: C = C - A*B^T      (accumulators start at -C, result is negated on store)
// MODE 2: C = C + A*B^T
// Variants (template parameters; picked at run time by GPK_GEMM_VARIANT for A/B measurements):
//   BN   column width of the CTA tile (rows are always 128)
//   WN   warps along N (4 along M): CTA = 128*WN threads, warp tile 32 x BN/WN
//   BK   k-slab per pipeline stage, ST stages, MINB CTAs per SM the register budget is sized for
//   WM   warps along M: the CTA tile has BM = 32*WM rows (4 -> 128 rows; 1 -> 32-row strips for the latency-critical,
//        in-place panel TRSM: four times as many CTAs per 128-row tile, each reading and writing only its own rows)
template <int MODE, int BN, int WN, int BK, int ST, int MINB, int WM = 4>
__global__ void __launch_bounds__(32 * WM * WN, MINB) dgemm_nt_kernel(const GemmArgs p) {
  constexpr int THREADS = 32 * WM * WN;
  constexpr int BM = 32 * WM;
  constexpr int LDA_S = BM + 4;          // == 4 mod 16 for 32, 64, 128
  constexpr int WTN = BN / WN;           // warp tile width
  constexpr int NT = WTN / 8;            // 8-wide MMA tiles per warp along N
  constexpr int LDB = BN + 4;            // == 4 mod 16 for 64 and 128
  constexpr int A_STAGE = BK * LDA_S;
  constexpr int B_STAGE = BK * LDB;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + ST * A_STAGE;

  const int ti = blockIdx.x, tjs = blockIdx.y;          // tjs counts BN-wide column tiles
  const int grow0 = ti * BM + p.ti_off * NB;           // first global row of this tile
  const int gi = grow0 / NB;                            // its 128-row tile index
  constexpr int SUBS = NB / BN;                         // BN-wide sub-tiles per 128-wide column tile
  const int tc = tjs / SUBS, sub = tjs % SUBS;
  const int cs = (p.cstride > 1) ? p.cstride : 1;
  const int gcol0 = (p.tj_off + tc * cs) * NB + sub * BN;   // first global column of this tile
  if (p.tri && grow0 + BM - 1 < gcol0) return;          // tile entirely above the diagonal
  const bool diag_tile = (p.tri != 0) && (gcol0 + BN - 1 > grow0);   // crosses the diagonal
  const int dshift = gcol0 - grow0;                     // store (r,c) iff r >= c + dshift

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WM, wn = warp / WM;
  const int g = lane >> 2, t = lane & 3;

  const double* __restrict__ Ag = p.A + (int64_t)ti * BM;
  const double* __restrict__ Bg = p.B + (int64_t)(tc * cs) * NB + sub * BN;
  double* __restrict__ Cg = p.C + (int64_t)ti * BM + (int64_t)tjs * BN * p.ldc;

  const int kbeg = (p.tri == 2) ? gi * NB : 0;
  const int nkt = (p.K - kbeg) / BK;

  auto load_stage = [&](int slot, int kt) {
    const int k0 = kbeg + kt * BK;
#pragma unroll
    for (int i = 0; i < (BK * BM / 2) / THREADS; ++i) {
      const int c = tid + i * THREADS;       // BM/2 16-byte chunks per column of A
      const int col = c / (BM / 2);
      const int r = (c % (BM / 2)) * 2;
      cp_async16(As + slot * A_STAGE + col * LDA_S + r, Ag + r + (int64_t)(k0 + col) * p.lda);
    }
#pragma unroll
    for (int i = 0; i < (BK * BN / 2) / THREADS; ++i) {
      const int c = tid + i * THREADS;       // BN/2 chunks per column of B
      const int col = c / (BN / 2);
      const int r = (c % (BN / 2)) * 2;
      cp_async16(Bs + slot * B_STAGE + col * LDB + r, Bg + r + (int64_t)(k0 + col) * p.ldb);
    }
  };

  // start the pipeline before touching C so the loads overlap
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (s < nkt) load_stage(s, s);
    cp_async_commit();
  }

  double acc[2][NT][4];
  const int row_base = wm * 32 + g;
  const int col_base = wn * WTN + 2 * t;
  if (MODE != 0) {
    const double isg = (MODE == 1) ? -1.0 : 1.0;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        const int r = row_base + mi * 16, c = col_base + ni * 8;
        const double* cp = Cg + r + (int64_t)c * p.ldc;
        acc[mi][ni][0] = isg * cp[0];
        acc[mi][ni][1] = isg * cp[p.ldc];
        acc[mi][ni][2] = isg * cp[8];
        acc[mi][ni][3] = isg * cp[8 + p.ldc];
      }
  } else {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < NT; ++ni)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.0;
  }

  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    {
      const int nk = kt + ST - 1;
      if (nk < nkt) load_stage(nk % ST, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % ST) * A_STAGE;
    const double* bs = Bs + (kt % ST) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 8; ++kk) {
      double a[2][4];
      const double* a_lo = as + (kk * 8 + t) * LDA_S + wm * 32 + g;
      const double* a_hi = a_lo + 4 * LDA_S;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        a[mi][0] = a_lo[mi * 16];
        a[mi][1] = a_lo[mi * 16 + 8];
        a[mi][2] = a_hi[mi * 16];
        a[mi][3] = a_hi[mi * 16 + 8];
      }
      const double* b_lo = bs + (kk * 8 + t) * LDB + wn * WTN + g;
      const double* b_hi = b_lo + 4 * LDB;
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        const double b0 = b_lo[ni * 8], b1 = b_hi[ni * 8];
        dmma_16x8x8(acc[0][ni], a[0], b0, b1);
        dmma_16x8x8(acc[1][ni], a[1], b0, b1);
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: registers -> global (each store instruction covers 4 columns x 64 contiguous bytes)
  const double sgn = (MODE == 1) ? -1.0 : 1.0;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      const int r = row_base + mi * 16, c = col_base + ni * 8;
      double* cp = Cg + r + (int64_t)c * p.ldc;
      if (!diag_tile) {
        cp[0] = sgn * acc[mi][ni][0];
        cp[p.ldc] = sgn * acc[mi][ni][1];
        cp[8] = sgn * acc[mi][ni][2];
        cp[8 + p.ldc] = sgn * acc[mi][ni][3];
      } else {
        const int cc = c + dshift;
        if (r >= cc) cp[0] = sgn * acc[mi][ni][0];
        if (r >= cc + 1) cp[p.ldc] = sgn * acc[mi][ni][1];
        if (r + 8 >= cc) cp[8] = sgn * acc[mi][ni][2];
        if (r + 8 >= cc + 1) cp[8 + p.ldc] = sgn * acc[mi][ni][3];
      }
    }
}

template <int BN, int BK, int ST, int WM = 4>
constexpr size_t gemm_smem() {
  return size_t(ST) * BK * (32 * WM + 4 + BN + 4) * sizeof(double);
}

// variant table:      BN   WN  BK  ST  MINB
#define GPK_V0 64, 2, 16, 4, 2      /* 128x64, 8 warps, two CTAs per SM (default)        */
#define GPK_V1 128, 2, 16, 4, 1     /* 128x128, 8 warps, one CTA per SM (round-1 first cut) */
#define GPK_V2 128, 4, 16, 4, 1     /* 128x128, 16 warps, one CTA per SM                  */
#define GPK_V3 128, 4, 32, 3, 1     /* 128x128, 16 warps, 32-deep k-slabs, 3 stages       */
#define GPK_V4 64, 2, 32, 2, 2      /* 128x64, 8 warps, two CTAs per SM, 32-deep k-slabs  */
#define GPK_V5 128, 4, 16, 4, 2, 1  /* 32x128 row strips, 4 warps, two CTAs per SM: in-place products (panel TRSM) */

static int g_gemm_variant = 0;
static int g_trsm_strip = 1;     // in-place products in 32-row strips (V5) instead of whole 128-row tiles (V1)

template <int BN, int WN, int BK, int ST, int MINB, int WM = 4>
static int variant_init(Handle* h) {
  GPK_CK(h, cudaFuncSetAttribute(dgemm_nt_kernel<0, BN, WN, BK, ST, MINB, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)gemm_smem<BN, BK, ST, WM>()));
  GPK_CK(h, cudaFuncSetAttribute(dgemm_nt_kernel<1, BN, WN, BK, ST, MINB, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)gemm_smem<BN, BK, ST, WM>()));
  GPK_CK(h, cudaFuncSetAttribute(dgemm_nt_kernel<2, BN, WN, BK, ST, MINB, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)gemm_smem<BN, BK, ST, WM>()));
  return 0;
}

template <int BN, int WN, int BK, int ST, int MINB, int WM = 4>
static void variant_launch(cudaStream_t st, int mode, const GemmArgs& a, int tiles_m, int tiles_n) {
  dim3 grid((unsigned)(tiles_m * (4 / WM)), (unsigned)(tiles_n * (NB / BN)));
  constexpr int TH = 32 * WM * WN;
  if (mode == 0)
    dgemm_nt_kernel<0, BN, WN, BK, ST, MINB, WM><<<grid, TH, gemm_smem<BN, BK, ST, WM>(), st>>>(a);
  else if (mode == 1)
    dgemm_nt_kernel<1, BN, WN, BK, ST, MINB, WM><<<grid, TH, gemm_smem<BN, BK, ST, WM>(), st>>>(a);
  else
    dgemm_nt_kernel<2, BN, WN, BK, ST, MINB, WM><<<grid, TH, gemm_smem<BN, BK, ST, WM>(), st>>>(a);
}

int gemm_init(Handle* h) {
  if (const char* e = getenv("GPK_GEMM_VARIANT")) g_gemm_variant = atoi(e);
  if (g_gemm_variant < 0 || g_gemm_variant > 4) g_gemm_variant = 0;
  GPK_TRY((variant_init<GPK_V0>(h)));
  GPK_TRY((variant_init<GPK_V1>(h)));
  GPK_TRY((variant_init<GPK_V2>(h)));
  GPK_TRY((variant_init<GPK_V3>(h)));
  GPK_TRY((variant_init<GPK_V4>(h)));
  GPK_TRY((variant_init<GPK_V5>(h)));
  if (const char* e = getenv("GPK_TRSM_STRIP")) g_trsm_strip = atoi(e);
  return 0;
}

int launch_gemm_nt(Handle* h, cudaStream_t st, int mode, const GemmArgs& a, int tiles_m, int tiles_n) {
  if (tiles_m <= 0 || tiles_n <= 0) return 0;
  if (a.K % 32 != 0 || tiles_n > 32000) return GPK_ERR_ARG;
  // In-place products (C overwrites A: the panel TRSM and the first step of the multi-right-hand-side sweeps)
  // are only safe when ONE CTA owns a whole 128-row tile of A: it finishes reading the tile before its
  // epilogue writes it.  Column-split tiles (BN=64) would let a sibling CTA overwrite columns still being read.
  const bool inplace = (static_cast<const double*>(a.C) == a.A);
  // Either ONE CTA owns the whole 128-row tile (V1), or - the default - four CTAs own 32-row strips of it (V5): a
  // strip's CTA reads only its own rows of A, so the in-place product is still race-free, and the latency-critical
  // TRSM of the Cholesky panel chain spreads over four times as many SMs.
  switch (inplace ? (g_trsm_strip ? 5 : 1) : (a.strips ? 5 : g_gemm_variant)) {
    case 5: variant_launch<GPK_V5>(st, mode, a, tiles_m, tiles_n); break;
    case 1: variant_launch<GPK_V1>(st, mode, a, tiles_m, tiles_n); break;
    case 2: variant_launch<GPK_V2>(st, mode, a, tiles_m, tiles_n); break;
    case 3: variant_launch<GPK_V3>(st, mode, a, tiles_m, tiles_n); break;
    case 4: variant_launch<GPK_V4>(st, mode, a, tiles_m, tiles_n); break;
    default: variant_launch<GPK_V0>(st, mode, a, tiles_m, tiles_n); break;
  }
  h->stats.launches++;
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// fp64 pipe micro-benchmarks (the roofline denominators MEASURED_PEAKS.json lacks)
// ---------------------------------------------------------------------------
template <int SHAPE>
__global__ void __launch_bounds__(512) dmma_peak_kernel(double* out, int iters) {
  const int lane = threadIdx.x & 31;
  double seed = 1.0 + 1e-9 * lane;
  if (SHAPE == 4) {  // plain DFMA, 16 independent chains
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = seed * i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
    return;
  }
  double c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[i][q] = 0.0;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * 1e-3 * (i + 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = seed * 1e-3 * (i + 2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
            "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
            : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
            : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
              "d"(b[1]), "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) s += c[i][q];
  if (s == 123.456) out[0] = s;
}

// shape 5: even warps run DMMA chains, odd warps run DFMA chains: do the two fp64 paths overlap?
__global__ void __launch_bounds__(512) mixed_peak_kernel(double* out, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double seed = 1.0 + 1e-9 * lane;
  if (warp & 1) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = seed * i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
  } else {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[i][q] = 0.0;
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = seed * 1e-3 * (i + 1);
    b[0] = seed * 2e-3; b[1] = seed * 3e-3;
    // same flop count per iteration as the DFMA warps: 16 FMA/lane = 512 flop... DMMA m16n8k8 = 2048 flop/instr
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) s += c[i][q];
    if (s == 123.456) out[0] = s;
  }
}

int bench_dmma(Handle* h, int shape, int warps, int iters, double* tflops, double* ms_out) {
  if (shape == 5) {
    if (warps < 2 || warps > 16 || iters < 1) return GPK_ERR_ARG;
    int sms = 0;
    GPK_CK(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    double* d = nullptr;
    GPK_CK(h, cudaMalloc(&d, 64));
    dim3 grid(sms * 2), block(warps * 32);
    mixed_peak_kernel<<<grid, block, 0, h->s_main>>>(d, iters / 10 + 1);
    GPK_CK(h, cudaStreamSynchronize(h->s_main));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      GPK_CK(h, cudaEventRecord(h->t0, h->s_main));
      mixed_peak_kernel<<<grid, block, 0, h->s_main>>>(d, iters);
      GPK_CK(h, cudaEventRecord(h->t1, h->s_main));
      GPK_CK(h, cudaEventSynchronize(h->t1));
      float ms = 0;
      GPK_CK(h, cudaEventElapsedTime(&ms, h->t0, h->t1));
      if (ms < best) best = ms;
    }
    GPK_CK(h, cudaGetLastError());
    cudaFree(d);
    // per iteration: DMMA warp 8 x 2048 flop, DFMA warp 16 x 64 flop ; the kernel ends when the slower half ends
    const double total = ((warps / 2) * 8.0 * 2048.0 + (warps / 2) * 16.0 * 64.0) * iters * (double)grid.x;
    *ms_out = best;
    *tflops = total / (best * 1e-3) / 1e12;
    return 0;
  }
  if (shape < 0 || shape > 4 || warps < 1 || warps > 16 || iters < 1) return GPK_ERR_ARG;
  int sms = 0;
  GPK_CK(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  double* d = nullptr;
  GPK_CK(h, cudaMalloc(&d, 64));
  dim3 grid(sms * 2), block(warps * 32);
  auto run = [&](int it) {
    switch (shape) {
      case 0: dmma_peak_kernel<0><<<grid, block, 0, h->s_main>>>(d, it); break;
      case 1: dmma_peak_kernel<1><<<grid, block, 0, h->s_main>>>(d, it); break;
      case 2: dmma_peak_kernel<2><<<grid, block, 0, h->s_main>>>(d, it); break;
      case 3: dmma_peak_kernel<3><<<grid, block, 0, h->s_main>>>(d, it); break;
      default: dmma_peak_kernel<4><<<grid, block, 0, h->s_main>>>(d, it); break;
    }
  };
  run(iters / 10 + 1);
  GPK_CK(h, cudaStreamSynchronize(h->s_main));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    GPK_CK(h, cudaEventRecord(h->t0, h->s_main));
    run(iters);
    GPK_CK(h, cudaEventRecord(h->t1, h->s_main));
    GPK_CK(h, cudaEventSynchronize(h->t1));
    float ms = 0;
    GPK_CK(h, cudaEventElapsedTime(&ms, h->t0, h->t1));
    if (ms < best) best = ms;
  }
  GPK_CK(h, cudaGetLastError());
  cudaFree(d);
  static const double flops_per[5] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16, 2.0 * 32};
  const double per_warp_iter = (shape == 4 ? 16.0 : 8.0) * flops_per[shape];
  const double total = per_warp_iter * iters * warps * (double)grid.x;
  *ms_out = best;
  *tflops = total / (best * 1e-3) / 1e12;
  return 0;
}

}  // namespace gpk
