// ozaki.cu - the fp64 trailing update C -= P·P' of the blocked Cholesky, computed on the int8 tensor cores.
//
// Why: B200's fp64 pipe peaks at 37 TFLOP/s (DMMA and DFMA share it; profiles/r1_peaks_first_cut.json), and the
// rank-(W*128) update is >90 % of gp.GPR.getPosterior() at N=16384.  `tcgen05.mma.kind::i8` runs at 4.5 POP/s with
// EXACT int32 accumulation in TMEM.  The update is therefore computed by an error-free split (Ozaki scheme):
//
//   row i of the panel:  P[i,:] = 2^e_i * sum_{t=1..S} d_t[i,:] * 2^(-7t)  + O(2^(e_i-7S-1)),   d_t int8 in [-64,64]
//   (P P')[i,j]          = 2^(e_i+e_j) * sum_{m=2..S+1} 2^(-7m) * G_m[i,j],   G_m = sum_{t+u=m} d_t[i,:]·d_u[j,:]
//
// Each G_m is an exact integer (|G_m| <= 8 * 384 * 64^2 < 2^31), the S accumulators G_2..G_{S+1} live in TMEM
// (S x 64 columns), and only the final weighted sum is rounded, in fp64.  Products with t+u > S+1 are below 2^(-7(S+2)) relative to
// the row scales and are dropped.  With S=8 the split keeps 56 bits per entry relative to the row maximum, i.e.
// the result is at least as accurate as an fp64 DMMA accumulation of the same contraction.
//
// Kernels:
//   oz_rowexp_kernel : e_i = exponent of the largest |P[i,k]| over the panel row
//   oz_slice_kernel  : writes the int8 slices directly in the tensor core's canonical K-major "core matrix" order
//                      [k-step 32][row group 8][slice t][k chunk 2][row 8][16 B], so that a (rows x 32 k) stage of
//                      ALL slices of a tile is ONE contiguous block: a single cp.async.bulk (TMA) per operand per stage
//   oz_syrk_kernel   : one CTA per 128x64 tile of C; warp 5 = TMA producer, warp 4 = MMA issuer, warps 0-3 = epilogue
//                      (TMEM -> fp64 -> C).  S(S+1)/2 MMAs of 128x64x32 per k-step.
#include <cstdint>
#include <cstdlib>
#include "gpk_internal.cuh"
#include "tc_common.cuh"

namespace gpk {

constexpr int OZ_BN = 64;        // tile columns
constexpr int OZ_ST = 4;         // pipeline stages
constexpr int OZ_EPI_WARPS = 8;  // epilogue warps 0..7, then the MMA issuer warp and the TMA producer warp
constexpr int OZ_THREADS = (OZ_EPI_WARPS + 2) * 32;

__global__ void __launch_bounds__(256) oz_rowexp_kernel(const double* __restrict__ P, int64_t lda, int n, int kw,
                                                         int* __restrict__ ex) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double m = 0.0;
#pragma unroll 8
  for (int k = 0; k < kw; ++k) m = fmax(m, fabs(P[i + (int64_t)k * lda]));
  // |x| < 2^(ilogb+1)  ->  |x| * 2^-(ilogb+2) < 0.5
  int e = 0;
  if (m > 0.0 && m < 1.0e300) e = ilogb(m) + 2;
  ex[i] = e;
}

template <int S>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ P, int64_t lda, int n, int kw,
                                                        const int* __restrict__ ex, int8_t* __restrict__ sl) {
  // CTA: 32 rows x 32 k (one k-step).  Output block = 4 row groups x S x 256 B, contiguous in `sl`.
  __shared__ __align__(16) int8_t out[4][S][2][8][16];
  const int r0 = blockIdx.x * 32, ks = blockIdx.y;
  const int r = threadIdx.x & 31, kq = threadIdx.x >> 5;
  const int e = ex[r0 + r];
  for (int kk = kq; kk < 32; kk += 8) {
    double x = scalbn(P[(r0 + r) + (int64_t)(ks * 32 + kk) * lda], -e);   // exact; |x| < 0.5
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const double y = x * 128.0;
      const double d = rint(y);          // in [-64, 64]
      x = y - d;                         // exact, |x| <= 0.5
      out[r >> 3][t][kk >> 4][r & 7][kk & 15] = (int8_t)(int)d;
    }
  }
  __syncthreads();
  int8_t* dst = sl + (size_t)ks * ((size_t)n * S * 32) + (size_t)r0 * S * 32;
  const int4* src4 = reinterpret_cast<const int4*>(&out[0][0][0][0][0]);
  int4* dst4 = reinterpret_cast<int4*>(dst);
  for (int c = threadIdx.x; c < 4 * S * 16; c += 256) dst4[c] = src4[c];
}

struct OzArgs {
  const int8_t* sl;   // slices of the panel rows [kw/32][n/8][S][2][8][16]
  const int* ex;      // row exponents
  double* C;          // trailing matrix origin (row 0 / col 0 of the sliced rows), column-major
  int64_t ldc;
  int n, kw;          // sliced rows, contraction length
  int cj0;            // first 64-column tile of this launch
};

template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_syrk_kernel(OzArgs a) {
  constexpr uint32_t A_BYTES = 128 * S * 32, B_BYTES = OZ_BN * S * 32, STAGE = A_BYTES + B_BYTES;
  constexpr uint32_t TCOLS = 512;
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[OZ_ST], empty[OZ_ST], done;
  __shared__ uint32_t tmem_base;
  __shared__ double scol[OZ_BN];     // 2^(e_j - 7) of the tile's columns

  const int ti = blockIdx.x, tj = a.cj0 + blockIdx.y;
  const int row0 = ti * 128, col0 = tj * OZ_BN;
  if (row0 + 128 <= col0) return;                       // tile entirely above the diagonal
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = a.kw / 32;
  const size_t kstride = (size_t)a.n * S * 32;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < OZ_ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid >= 64 && tid < 64 + OZ_BN) scol[tid - 64] = scalbn(1.0, a.ex[col0 + tid - 64] - 7);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;

  if (warp == OZ_EPI_WARPS + 1) {
    if (elect_one()) {
      const int8_t* gA = a.sl + (size_t)row0 * S * 32;
      const int8_t* gB = a.sl + (size_t)col0 * S * 32;
      for (int ks = 0; ks < nk; ++ks) {
        const int slot = ks % OZ_ST;
        if (ks >= OZ_ST) mbar_wait(&empty[slot], ((ks / OZ_ST) - 1) & 1);
        mbar_expect_tx(&full[slot], STAGE);
        uint8_t* s = sm + (size_t)slot * STAGE;
        bulk_g2s(s, gA + ks * kstride, A_BYTES, &full[slot]);
        bulk_g2s(s + A_BYTES, gB + ks * kstride, B_BYTES, &full[slot]);
      }
    }
  } else if (warp == OZ_EPI_WARPS) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_i8(128, OZ_BN);
      // descriptors differ only in the 14-bit start-address field: slice t sits 256 B (= 16 units) further
      const uint64_t dbase = make_smem_desc(0, 128, S * 256);
      for (int ks = 0; ks < nk; ++ks) {
        const int slot = ks % OZ_ST;
        mbar_wait(&full[slot], (ks / OZ_ST) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(sm + (size_t)slot * STAGE);
        const uint64_t da = dbase | (uint64_t)(sa >> 4), db = dbase | (uint64_t)((sa + A_BYTES) >> 4);
#pragma unroll
        for (int t = 0; t < S; ++t) {
#pragma unroll
          for (int u = 0; u < S - t; ++u)
            tc_mma_i8(tbase + (uint32_t)(t + u) * OZ_BN, da + 16 * t, db + 16 * u, idesc, (ks > 0 || t > 0) ? 1u : 0u);
        }
        tc_commit(&empty[slot]);
      }
      tc_commit(&done);
    }
  } else {
    // epilogue warps: TMEM lanes 32*(warp%4).. = rows of the tile; warps 4..7 take the upper 32 columns.
    // The C tile is fetched BEFORE the accumulators are complete, so its latency hides under the MMA loop.
    const int q4 = warp & 3, chalf = (warp >> 2) * (OZ_BN / 2);
    const int gi = row0 + q4 * 32 + lane;
    const double si = scalbn(1.0, a.ex[gi] - 7);
    double* crow = a.C + gi + (int64_t)(col0 + chalf) * a.ldc;
    double cv[OZ_BN / 2];
#pragma unroll
    for (int q = 0; q < OZ_BN / 2; ++q) cv[q] = (gi >= col0 + chalf + q) ? crow[(int64_t)q * a.ldc] : 0.0;
    mbar_wait(&done, 0);
    tc_fence_after();
    const uint32_t tw = tbase + ((uint32_t)(q4 * 32) << 16) + chalf;
#pragma unroll
    for (int c0 = 0; c0 < OZ_BN / 2; c0 += 16) {
      double acc[16];
      uint32_t v[16];
      tc_ld16(tw + (uint32_t)(S - 1) * OZ_BN + c0, v);
      tc_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = (double)(int)v[q];
#pragma unroll
      for (int m = S - 2; m >= 0; --m) {
        tc_ld16(tw + (uint32_t)m * OZ_BN + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = fma(acc[q], 0.0078125, (double)(int)v[q]);
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (gi >= col0 + chalf + c0 + q)
          crow[(int64_t)(c0 + q) * a.ldc] = cv[c0 + q] - (acc[q] * si) * scol[chalf + c0 + q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(TCOLS) : "memory");
}

static int oz_slices() {
  static int s = -1;
  if (s < 0) {
    s = 8;
    if (const char* e = getenv("GPK_OZAKI_SLICES")) { const int v = atoi(e); if (v == 7 || v == 8) s = v; }
  }
  return s;
}

int oz_ensure(Handle* h, int64_t n, int kw) {
  const size_t need = (size_t)n * kw * 8;
  if (h->ozCap < need) {
    if (h->ozSl) cudaFree(h->ozSl);
    h->ozSl = nullptr; h->ozCap = 0;
    GPK_CK(h, cudaMalloc((void**)&h->ozSl, need));
    h->ozCap = need;
  }
  if (h->ozExCap < (size_t)n) {
    if (h->ozEx) cudaFree(h->ozEx);
    h->ozEx = nullptr; h->ozExCap = 0;
    GPK_CK(h, cudaMalloc((void**)&h->ozEx, (size_t)n * sizeof(int)));
    h->ozExCap = (size_t)n;
  }
  return 0;
}

// slice the panel P (n rows x kw columns, column-major, lda) into h->ozSl / h->ozEx
int launch_oz_slice(Handle* h, cudaStream_t st, const double* P, int64_t lda, int n, int kw) {
  if (n % 128 != 0 || kw % 32 != 0) return GPK_ERR_ARG;
  const int S = oz_slices();
  oz_rowexp_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, lda, n, kw, h->ozEx);
  dim3 g(n / 32, kw / 32);
  if (S == 8) oz_slice_kernel<8><<<g, 256, 0, st>>>(P, lda, n, kw, h->ozEx, h->ozSl);
  else oz_slice_kernel<7><<<g, 256, 0, st>>>(P, lda, n, kw, h->ozEx, h->ozSl);
  GPK_CK(h, cudaGetLastError());
  return 0;
}

// C(lower tiles, 64-column tiles [cj0, cj1)) -= P·P' from the current slices
int launch_oz_syrk(Handle* h, cudaStream_t st, double* C, int64_t ldc, int n, int kw, int cj0, int cj1) {
  const int S = oz_slices();
  static bool attr_done = false;
  const size_t smem8 = (size_t)OZ_ST * (128 + OZ_BN) * 8 * 32, smem7 = (size_t)OZ_ST * (128 + OZ_BN) * 7 * 32;
  if (!attr_done) {
    GPK_CK(h, cudaFuncSetAttribute(oz_syrk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
    GPK_CK(h, cudaFuncSetAttribute(oz_syrk_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7));
    attr_done = true;
  }
  OzArgs a{h->ozSl, h->ozEx, C, ldc, n, kw, cj0};
  dim3 g(n / 128, cj1 - cj0);
  if (S == 8) oz_syrk_kernel<8><<<g, OZ_THREADS, smem8, st>>>(a);
  else oz_syrk_kernel<7><<<g, OZ_THREADS, smem7, st>>>(a);
  GPK_CK(h, cudaGetLastError());
  return 0;
}

}  // namespace gpk

using namespace gpk;

// C (n x n, column-major, lower triangle) -= P (n x kw) · P'; host buffers; mode 0: int8 tensor-core path, 1: DMMA path.
// `ms` (optional) receives the device time of the update alone (slicing included for mode 0), averaged over `reps`.
extern "C" int gpk_dbg_oz_syrk(gpk_handle hh, int64_t n, int kw, const double* P, double* C, int mode, int reps,
                               double* ms) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!P || !C || n <= 0 || n % 128 != 0 || kw <= 0 || kw % 128 != 0 || reps < 1) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  double *dP = nullptr, *dC = nullptr;
  GPK_CK(h, cudaMalloc((void**)&dP, (size_t)n * kw * 8));
  GPK_CK(h, cudaMalloc((void**)&dC, (size_t)n * n * 8));
  int rc = 0;
  cudaMemcpyAsync(dP, P, (size_t)n * kw * 8, cudaMemcpyHostToDevice, st);
  if (mode == 0) rc = oz_ensure(h, n, kw);
  float total = 0.f;
  for (int r = 0; r < reps && rc == 0; ++r) {
    cudaMemcpyAsync(dC, C, (size_t)n * n * 8, cudaMemcpyHostToDevice, st);
    cudaEventRecord(h->t0, st);
    if (mode == 0) {
      rc = launch_oz_slice(h, st, dP, n, (int)n, kw);
      if (rc == 0) rc = launch_oz_syrk(h, st, dC, n, (int)n, kw, 0, (int)(n / OZ_BN));
    } else {
      GemmArgs u{};
      u.A = dP; u.B = dP; u.C = dC; u.lda = n; u.ldb = n; u.ldc = n; u.K = kw; u.tri = 1;
      rc = launch_gemm_nt(h, st, 1, u, (int)(n / NB), (int)(n / NB));
    }
    cudaEventRecord(h->t1, st);
    cudaStreamSynchronize(st);
    float t = 0.f;
    cudaEventElapsedTime(&t, h->t0, h->t1);
    total += t;
  }
  if (rc == 0) cudaMemcpyAsync(C, dC, (size_t)n * n * 8, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dP); cudaFree(dC);
  if (rc) return rc;
  GPK_CK(h, e);
  GPK_CK(h, cudaGetLastError());
  if (ms) *ms = total / reps;
  return 0;
}
