// lat_probe.cu - single-warp latency / issue-rate probe for the instructions on the critical path of the
// 128x128 diagonal-block kernel (potrf_diag.cu): DFMA, SHFL, MUFU.RCP64H, LDS broadcast, STS->LDS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/lat_probe scripts/lat_probe.cu && scripts/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr int IT = 256;
__global__ void probe(long long* out, double* sink, int warps_active) {
  __shared__ double sm[64];
  __shared__ int chase[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_active) return;
  sm[lane] = 1.0 + lane; sm[lane + 32] = 2.0;
  chase[lane] = (lane + 1) & 31;
  __syncwarp();
  double x = 1.0 + 1e-9 * lane, y = 0.999999, z = 1e-12;
  long long t0, t1;
  int k = 0;
  // 1 dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) x = fma(x, y, z);
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  // 2 eight independent DFMA chains
  double c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = x + j;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT / 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = fma(c[j], y, z);
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
#pragma unroll
  for (int j = 0; j < 8; ++j) x += c[j];
  // 3 dependent SHFL chain (32-bit)
  int v = lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  // 4 independent SHFLs (fixed source lanes, like the Cholesky broadcast)
  int acc = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) acc += __shfl_sync(0xffffffffu, v + i, i & 31);
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  v += acc;
  // 5 dependent rcp.approx.ftz.f64 + DFMA (pair latency)
  double r = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = fma(q, y, 1.5); }
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  x += r;
  // 6 dependent LDS chain
  int p = lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) p = chase[p];
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  v += p;
  // 7 independent broadcast LDS.128 (all lanes same address)
  double s = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) { double wx, wy; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(wx), "=d"(wy) : "r"((unsigned)__cvta_generic_to_shared(&sm[(2 * i) & 62]))); s += wx + wy; }
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  x += s;
  // 8 STS -> syncwarp -> broadcast LDS round trip, dependent
  double u = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) { sm[lane] = u; __syncwarp(); u = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  x += u;
  // 9 dependent 64-bit shuffle + DFMA (the Cholesky inner dependency)
  double g = x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i) { double b = __shfl_sync(0xffffffffu, g, i & 31); g = fma(b, y, z); }
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  x += g;
  // 10 DMMA m16n8k8 dependent chain
  double d4[4] = {x, x, x, x};
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < IT; ++i)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d4[0]), "+d"(d4[1]), "+d"(d4[2]), "+d"(d4[3]) : "d"(y), "d"(y), "d"(y), "d"(y), "d"(z), "d"(z));
  t1 = clock64(); if (threadIdx.x == 0) out[k] = t1 - t0; ++k;
  x += d4[0] + d4[1] + d4[2] + d4[3];
  sink[threadIdx.x] = x + v;
}
int main() {
  long long* d; double* s;
  cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 1024 * 8);
  const char* names[] = {"DFMA dependent", "DFMA 8 chains (per op)", "SHFL dependent", "SHFL independent (per op)",
                         "RCP64H+DFMA dependent pair", "LDS dependent", "LDS.128 broadcast independent (per op)",
                         "STS->syncwarp->LDS->DADD round trip", "SHFL64+DFMA dependent pair", "DMMA m16n8k8 dependent"};
  for (int warps : {1, 4, 8}) {
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 256>>>(d, s, warps);
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps active in the CTA: %d  (%s)\n", warps, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 10; ++i) printf("  %-42s %7.1f clk\n", names[i], (double)h[i] / IT);
  }
  return 0;
}
