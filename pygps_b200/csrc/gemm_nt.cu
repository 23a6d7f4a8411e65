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
// Tile: 128x128 per CTA, 8 warps as 4(M) x 2(N), warp tile 32x64 = 2x8 MMA
// tiles, 64 fp64 accumulators (128 registers) per thread, one CTA per SM.
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
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_nt_kernel(const GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + GEMM_STAGES * GEMM_BK * GEMM_LDS;

  const int ti = blockIdx.x, tj = blockIdx.y;
  const int gi = ti + p.ti_off, gj = tj + p.tj_off;
  if (p.tri && gi < gj) return;
  const bool diag_tile = (p.tri != 0) && (gi == gj);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;

  const double* __restrict__ Ag = p.A + (int64_t)ti * NB;
  const double* __restrict__ Bg = p.B + (int64_t)tj * NB;
  double* __restrict__ Cg = p.C + (int64_t)ti * NB + (int64_t)tj * NB * p.ldc;

  const int kbeg = (p.tri == 2) ? gi * NB : 0;
  const int nkt = (p.K - kbeg) / GEMM_BK;

  auto load_stage = [&](int slot, int kt) {
    const int k0 = kbeg + kt * GEMM_BK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * GEMM_THREADS;  // 0..1023 16-byte chunks per operand
      const int col = c >> 6;                // 0..15
      const int r = (c & 63) * 2;            // 0..126
      cp_async16(As + (slot * GEMM_BK + col) * GEMM_LDS + r, Ag + r + (int64_t)(k0 + col) * p.lda);
      cp_async16(Bs + (slot * GEMM_BK + col) * GEMM_LDS + r, Bg + r + (int64_t)(k0 + col) * p.ldb);
    }
  };

  // start the pipeline before touching C so the loads overlap
#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < nkt) load_stage(s, s);
    cp_async_commit();
  }

  double acc[2][8][4];
  const int row_base = wm * 32 + g;
  const int col_base = wn * 64 + 2 * t;
  if (MODE == 1) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        const int r = row_base + mi * 16, c = col_base + ni * 8;
        const double* cp = Cg + r + (int64_t)c * p.ldc;
        acc[mi][ni][0] = -cp[0];
        acc[mi][ni][1] = -cp[p.ldc];
        acc[mi][ni][2] = -cp[8];
        acc[mi][ni][3] = -cp[8 + p.ldc];
      }
  } else {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.0;
  }

  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + GEMM_STAGES - 1;
      if (nk < nkt) load_stage(nk % GEMM_STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % GEMM_STAGES) * GEMM_BK * GEMM_LDS;
    const double* bs = Bs + (kt % GEMM_STAGES) * GEMM_BK * GEMM_LDS;
#pragma unroll
    for (int kk = 0; kk < GEMM_BK / 8; ++kk) {
      double a[2][4];
      const double* a_lo = as + (kk * 8 + t) * GEMM_LDS + wm * 32 + g;
      const double* a_hi = a_lo + 4 * GEMM_LDS;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        a[mi][0] = a_lo[mi * 16];
        a[mi][1] = a_lo[mi * 16 + 8];
        a[mi][2] = a_hi[mi * 16];
        a[mi][3] = a_hi[mi * 16 + 8];
      }
      const double* b_lo = bs + (kk * 8 + t) * GEMM_LDS + wn * 64 + g;
      const double* b_hi = b_lo + 4 * GEMM_LDS;
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
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
    for (int ni = 0; ni < 8; ++ni) {
      const int r = row_base + mi * 16, c = col_base + ni * 8;
      double* cp = Cg + r + (int64_t)c * p.ldc;
      if (!diag_tile) {
        cp[0] = sgn * acc[mi][ni][0];
        cp[p.ldc] = sgn * acc[mi][ni][1];
        cp[8] = sgn * acc[mi][ni][2];
        cp[8 + p.ldc] = sgn * acc[mi][ni][3];
      } else {
        if (r >= c) cp[0] = sgn * acc[mi][ni][0];
        if (r >= c + 1) cp[p.ldc] = sgn * acc[mi][ni][1];
        if (r + 8 >= c) cp[8] = sgn * acc[mi][ni][2];
        if (r + 8 >= c + 1) cp[8 + p.ldc] = sgn * acc[mi][ni][3];
      }
    }
}

int gemm_init(Handle* h) {
  GPK_CK(h, cudaFuncSetAttribute(dgemm_nt_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
  GPK_CK(h, cudaFuncSetAttribute(dgemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
  return 0;
}

int launch_gemm_nt(Handle* h, cudaStream_t st, int mode, const GemmArgs& a, int tiles_m, int tiles_n) {
  if (tiles_m <= 0 || tiles_n <= 0) return 0;
  if (a.K % GEMM_BK != 0 || tiles_n > 65535) return GPK_ERR_ARG;
  dim3 grid((unsigned)tiles_m, (unsigned)tiles_n);
  if (mode == 0)
    dgemm_nt_kernel<0><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(a);
  else
    dgemm_nt_kernel<1><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(a);
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

int bench_dmma(Handle* h, int shape, int warps, int iters, double* tflops, double* ms_out) {
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
