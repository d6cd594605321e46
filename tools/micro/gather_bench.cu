// Micro-benchmark (tools/micro): what a random 16-byte gather from a large array costs with different load flavours.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ld_nc(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_plain(const float4* p) { return *p; }
__device__ __forceinline__ float4 ld_cg(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cs(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_lu(const float4* p) { return __ldlu(p); }
__device__ __forceinline__ float4 ld_noalloc(const float4* p)
{
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_l2_64(const float4* p)
{
  float4 v;
  asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_evict_first(const float4* p)
{
  float4 v;
  asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int kMode>
__global__ void k_gather(const float4* __restrict__ a, const unsigned* __restrict__ idx, unsigned long long n, float* __restrict__ out)
{
  float acc = 0.f;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
  {
    const float4* p = a + idx[i];
    float4 v;
    if (kMode == 0) v = ld_nc(p);
    if (kMode == 1) v = ld_plain(p);
    if (kMode == 2) v = ld_cg(p);
    if (kMode == 3) v = ld_cs(p);
    if (kMode == 4) v = ld_lu(p);
    if (kMode == 5) v = ld_noalloc(p);
    if (kMode == 6) v = ld_l2_64(p);
    if (kMode == 7) v = ld_evict_first(p);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f)
    out[0] = acc;
}

int main()
{
  const unsigned long long N = 25ull << 20;       // 25 M float4 = 400 MB
  const unsigned long long G = 150ull << 20;      // gathers
  float4* a;
  unsigned* idx;
  float* out;
  cudaMalloc(&a, N * sizeof(float4));
  cudaMalloc(&idx, G * sizeof(unsigned));
  cudaMalloc(&out, 4);
  cudaMemset(a, 0, N * sizeof(float4));
  unsigned* h = (unsigned*)malloc(G * sizeof(unsigned));
  unsigned long long x = 88172645463325252ull;
  for (unsigned long long i = 0; i < G; ++i)
  {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    h[i] = (unsigned)(x % N);
  }
  cudaMemcpy(idx, h, G * sizeof(unsigned), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[8] = { "ld.global.nc (__ldg)", "ld.global", "ld.global.cg", "ld.global.cs", "ld.global.lu", "nc.L1::no_allocate", "nc.L2::64B", "nc.L1::evict_first" };
  for (int mode = 0; mode < 8; ++mode)
  {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
      cudaEventRecord(e0);
      switch (mode)
      {
        case 0: k_gather<0><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 1: k_gather<1><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 2: k_gather<2><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 3: k_gather<3><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 4: k_gather<4><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 5: k_gather<5><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 6: k_gather<6><<<148 * 8, 256>>>(a, idx, G, out); break;
        case 7: k_gather<7><<<148 * 8, 256>>>(a, idx, G, out); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("%-24s %8.3f ms  %6.1f G gathers/s  (%5.1f B of HBM time per gather at 6.5 TB/s)\n", names[mode], best, G / best / 1e6, best * 1e-3 * 6.5e12 / G);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
