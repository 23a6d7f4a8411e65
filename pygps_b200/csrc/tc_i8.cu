// tc_i8.cu - tcgen05 (5th-generation tensor core) int8 building blocks: the road past the fp64 roofline.
//
// B200's fp64 pipe tops out at 37 TFLOP/s (DMMA and DFMA share it, measured), while `tcgen05.mma.kind::i8`
// delivers 4.5 POP/s with exact int32 accumulation in TMEM.  An Ozaki-style split of the fp64 panel into int8
// slices turns the trailing update into exact integer GEMMs.  This file holds the verified primitives:
// shared-memory matrix descriptors (K-major, no swizzle), the instruction descriptor, TMEM alloc/ld and a single-CTA
// tile product used by the tests.
#include <cstdint>
#include "gpk_internal.cuh"

namespace gpk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE ("interleave") canonical layout, in 16-byte units: ((8,n),2):((1,SBO),LBO)
//   address(row, kchunk) = start + (row%8)*16 + (row/8)*SBO + kchunk*LBO
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// kind::i8, signed x signed -> s32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(mbar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}

// C(128 x N, int32, row-major) = A(128 x K, int8, row-major) * B(N x K, int8, row-major)^T ; one CTA, 128 threads.
template <int N>
__global__ void __launch_bounds__(128) i8_tile_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                      int32_t* __restrict__ C, int K) {
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N) : "memory");
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
    const uint32_t idesc = make_idesc_i8(128, N);
    for (int ks = 0; ks < K / 32; ++ks) {       // one MMA consumes 32 bytes of K = 2 chunks
      const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 8 * 16);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 8 * 16);
      tc_mma_i8(tbase, ad, bd, idesc, ks > 0 ? 1u : 0u);
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "n"(N < 32 ? 32 : N) : "memory");
}

}  // namespace gpk

using namespace gpk;

extern "C" int gpk_dbg_i8_tile(gpk_handle hh, int N, int K, const int8_t* A, const int8_t* B, int32_t* C) {
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
    i8_tile_kernel<64><<<1, 128, smem, st>>>(dA, dB, dC, K);
  } else if (N == 128) {
    cudaFuncSetAttribute(i8_tile_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_tile_kernel<128><<<1, 128, smem, st>>>(dA, dB, dC, K);
  } else {
    cudaFuncSetAttribute(i8_tile_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_tile_kernel<256><<<1, 128, smem, st>>>(dA, dB, dC, K);
  }
  cudaMemcpyAsync(C, dC, (size_t)128 * N * 4, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  GPK_CK(h, e);
  GPK_CK(h, cudaGetLastError());
  return 0;
}
