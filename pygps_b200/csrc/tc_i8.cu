// tc_i8.cu - tcgen05 (5th-generation tensor core) int8 building blocks: the road past the fp64 roofline.
//
// B200's fp64 pipe tops out at 37 TFLOP/s (DMMA and DFMA share it, measured), while `tcgen05.mma.kind::i8`
// delivers 4.5 POP/s with exact int32 accumulation in TMEM.  An Ozaki-style split of the fp64 panel into int8
// slices turns the trailing update into exact integer GEMMs.  This file holds the verified primitives:
// shared-memory matrix descriptors (K-major, no swizzle), the instruction descriptor, TMEM alloc/ld and a single-CTA
// tile product used by the tests.
#include <cstdint>
#include "gpk_internal.cuh"
#include "tc_common.cuh"

namespace gpk {

// C(128 x N, int32, row-major) = A(128 x K, int8, row-major) * B(N x K, int8, row-major)^T ; one CTA, 128 threads.
template <int N>
__global__ void __launch_bounds__(128) i8_tile_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                      int32_t* __restrict__ C, int K, int fmt) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kch = K / 16;                       // 16-byte chunks along K
  uint8_t* sA = sm;                             // [kch][128 rows][16 B]
  uint8_t* sB = sm + (size_t)kch * 128 * 16;    // [kch][N rows][16 B]
  for (int c = tid; c < kch * 128; c += 128) {
    const int kc = c / 128, r = c % 128;
    *reinterpret_cast<int4*>(sA + ((size_t)kc * 128 + r) * 16) = *reinterpret_cast<const int4*>(A + (size_t)r * K + kc * 16);
  }
  for (int c = tid; c < kch * N; c += 128) {
    const int kc = c / N, r = c % N;
    *reinterpret_cast<int4*>(sB + ((size_t)kc * N + r) * 16) = *reinterpret_cast<const int4*>(B + (size_t)r * K + kc * 16);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = tmem_base;
  if (tid == 0) {
    // fmt bit 0: A holds unsigned bytes, bit 1: B holds unsigned bytes (operand format fields of the descriptor)
    const uint32_t idesc = make_idesc_i8(128, N) & ~(((fmt & 1) ? (1u << 7) : 0u) | ((fmt & 2) ? (1u << 10) : 0u));
    for (int ks = 0; ks < K / 32; ++ks) {       // one MMA consumes 32 bytes of K = 2 chunks
      const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 8 * 16);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 8 * 16);
      if (fmt & 12) {
        // A staged through TMEM columns 256.. by tcgen05.cp; bit 3: every k-step reuses the same 8 columns
        const uint32_t ta = tbase + 256 + ((fmt & 8) ? 0 : ks * 8);
        tc_cp_128x256b(ta, ad);
        tc_mma_i8_ts(tbase, ta, bd, idesc, ks > 0 ? 1u : 0u);
      } else {
        tc_mma_i8(tbase, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
    }
    tc_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // warp w reads TMEM lanes 32w..32w+31 (= rows of D), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tc_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    const int row = warp * 32 + lane;
#pragma unroll
    for (int q = 0; q < 8; ++q) C[(size_t)row * N + c0 + q] = (int32_t)v[q];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(512) : "memory");
}

// issue-rate probe: `iters` back-to-back MMAs (128 x N x 32, int8) from one thread per CTA, operands anywhere in a
// 64 KB shared window, accumulators rotating over the 512 TMEM columns; clocks per MMA written per CTA.
__global__ void __launch_bounds__(128) i8_rate_kernel(int N, int iters, uint32_t lbo, uint32_t sbo, uint32_t astep,
                                                      int same_acc, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int c = tid; c < 65536 / 16; c += 128) reinterpret_cast<int4*>(sm)[c] = make_int4(0x01010101, 0x01010101, 0x01010101, 0x01010101);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) { mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_i8(128, N);
    const uint32_t amask = same_acc ? 0u : (uint32_t)(512 / N - 1);   // 512/N is a power of two for N in {64,128,256}
    const uint32_t sa = smem_u32(sm);
    uint64_t ad[8], bd[8];
    uint32_t dcol[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      ad[q] = make_smem_desc(sa + q * astep, lbo, sbo);
      bd[q] = make_smem_desc(sa + 32768 + q * astep, lbo, sbo);
      dcol[q] = tbase + ((uint32_t)q & amask) * N;
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) tc_mma_i8(dcol[q], ad[q], bd[q], idesc, 1u);
    }
    tc_commit(&mbar);
    mbar_wait(&mbar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = (float)(t1 - t0) / iters;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(512) : "memory");
}

}  // namespace gpk

using namespace gpk;

extern "C" int gpk_bench_i8_rate(gpk_handle hh, int N, int iters, int lbo, int sbo, int astep, int same_acc, int ctas,
                                 double* clk_per_mma, double* tops) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!clk_per_mma || N < 16 || N > 256 || N % 16 || iters < 1 || ctas < 1 || ctas > 1024) return GPK_ERR_ARG;
  float* d = nullptr;
  GPK_CK(h, cudaMalloc((void**)&d, ctas * sizeof(float)));
  GPK_CK(h, cudaFuncSetAttribute(i8_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  i8_rate_kernel<<<ctas, 128, 65536, h->s_main>>>(N, iters, (uint32_t)lbo, (uint32_t)sbo, (uint32_t)astep, same_acc, d);   // warm-up
  cudaEventRecord(h->t0, h->s_main);
  i8_rate_kernel<<<ctas, 128, 65536, h->s_main>>>(N, iters, (uint32_t)lbo, (uint32_t)sbo, (uint32_t)astep, same_acc, d);
  cudaEventRecord(h->t1, h->s_main);
  std::vector<float> v(ctas);
  cudaError_t e = cudaMemcpyAsync(v.data(), d, ctas * sizeof(float), cudaMemcpyDeviceToHost, h->s_main);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_main);
  cudaFree(d);
  GPK_CK(h, e);
  double m = 0;
  for (float x : v) m = x > m ? x : m;
  *clk_per_mma = m;
  if (tops) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->t0, h->t1);
    *tops = 2.0 * 128.0 * N * 32.0 * (double)iters * ctas / (ms * 1e-3) / 1e12;
  }
  return 0;
}

extern "C" int gpk_dbg_i8_tile(gpk_handle hh, int N, int K, const int8_t* A, const int8_t* B, int32_t* C, int fmt) {
  Handle* h;
  GPK_TRY(check_handle(hh, &h));
  if (!A || !B || !C || (N != 64 && N != 128 && N != 256) || K % 32 != 0 || K <= 0 || K > 256) return GPK_ERR_ARG;
  cudaStream_t st = h->s_main;
  int8_t *dA = nullptr, *dB = nullptr;
  int32_t* dC = nullptr;
  GPK_CK(h, cudaMalloc((void**)&dA, 128 * K));
  GPK_CK(h, cudaMalloc((void**)&dB, (size_t)N * K));
  GPK_CK(h, cudaMalloc((void**)&dC, (size_t)128 * N * 4));
  cudaMemcpyAsync(dA, A, 128 * K, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(dB, B, (size_t)N * K, cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(dC, 0xff, (size_t)128 * N * 4, st);
  const size_t smem = (size_t)(K / 16) * (128 + N) * 16;
  if (N == 64) {
    cudaFuncSetAttribute(i8_tile_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_tile_kernel<64><<<1, 128, smem, st>>>(dA, dB, dC, K, fmt);
  } else if (N == 128) {
    cudaFuncSetAttribute(i8_tile_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_tile_kernel<128><<<1, 128, smem, st>>>(dA, dB, dC, K, fmt);
  } else {
    cudaFuncSetAttribute(i8_tile_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_tile_kernel<256><<<1, 128, smem, st>>>(dA, dB, dC, K, fmt);
  }
  cudaMemcpyAsync(C, dC, (size_t)128 * N * 4, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  GPK_CK(h, e);
  GPK_CK(h, cudaGetLastError());
  return 0;
}
