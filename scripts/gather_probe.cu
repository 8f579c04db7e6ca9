// Random-gather probe: what does a B200 pay in DRAM traffic/time for isolated 4..128-byte gathers?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu ; ./gather_probe [granularity]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31; return z; }
__device__ __forceinline__ void ld256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); }
__device__ __forceinline__ void ld256_na(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); }
// MODE: 0 = 4 B (one int), 1 = 32 B as 1x256, 2 = 32 B as 2x128, 3 = 64 B as 2x256, 4 = 96 B as 3x256 (stride 96),
//       5 = 128 B as 4x256, 6 = 16 B as 1x128, 7 = 32 B 256 L1::no_allocate, 8 = 96 B stride 128, 9 = 8 B
template <int MODE>
__global__ void probe(const char *tab, size_t nrec, size_t stride, size_t nq, uint64_t *out)
{
  uint64_t acc = 0;
  const size_t step = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += step)
    {
      const size_t r = mix(q) % nrec;
      const char *p = tab + r * stride;
      uint64_t a = 0, b = 0, c = 0, d = 0;
      if (MODE == 0) acc += __ldg((const int *) p);
      if (MODE == 9) acc += __ldg((const unsigned long long *) p);
      if (MODE == 6) { ulonglong2 v = __ldg((const ulonglong2 *) p); acc += v.x ^ v.y; }
      if (MODE == 1) { ld256(p, a, b, c, d); acc += a ^ b ^ c ^ d; }
      if (MODE == 7) { ld256_na(p, a, b, c, d); acc += a ^ b ^ c ^ d; }
      if (MODE == 2) { ulonglong2 v = __ldg((const ulonglong2 *) p), w = __ldg((const ulonglong2 *) p + 1); acc += v.x ^ v.y ^ w.x ^ w.y; }
      if (MODE == 3) { ld256(p, a, b, c, d); acc += a ^ b ^ c ^ d; ld256(p + 32, a, b, c, d); acc += a ^ b ^ c ^ d; }
      if (MODE == 4 || MODE == 8) { ld256(p, a, b, c, d); acc += a ^ b ^ c ^ d; ld256(p + 32, a, b, c, d); acc += a ^ b ^ c ^ d; ld256(p + 64, a, b, c, d); acc += a ^ b ^ c ^ d; }
      if (MODE == 5) { for (int k = 0; k < 4; k++) { ld256(p + 32 * k, a, b, c, d); acc += a ^ b ^ c ^ d; } }
    }
  if (acc == 0x123456789abcdefull) out[0] = acc;
}
template <int MODE> void run(const char *name, const char *tab, size_t bytes, size_t stride, size_t useful, size_t nq, uint64_t *out)
{
  const size_t nrec = bytes / stride;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<MODE><<<148 * 8, 256>>>(tab, nrec, stride, nq, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<MODE><<<148 * 8, 256>>>(tab, nrec, stride, nq, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-34s table %5.0f MB  %7.3f ms  %7.2f Ggather/s  useful %7.1f GB/s\n", name, bytes / 1e6, ms, nq / ms / 1e6, nq * useful / ms / 1e6);
}
int main(int argc, char **argv)
{
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  size_t tab_mb = argc > 2 ? atoi(argv[2]) : 4096;
  if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set granularity %d -> %s\n", gran, cudaGetErrorString(e)); }
  size_t g = 0; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity)); printf("L2 fetch granularity limit: %zu\n", g);
  const size_t bytes = tab_mb << 20, nq = 200000000;
  char *tab; uint64_t *out; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 1, bytes)); CK(cudaMalloc(&out, 8));
  run<0>("4 B  (int)             stride 4", tab, bytes, 4, 4, nq, out);
  run<9>("8 B  (u64)             stride 8", tab, bytes, 8, 8, nq, out);
  run<6>("16 B (1x128)           stride 16", tab, bytes, 16, 16, nq, out);
  run<1>("32 B (1x256)           stride 32", tab, bytes, 32, 32, nq, out);
  run<7>("32 B (1x256 no_alloc)  stride 32", tab, bytes, 32, 32, nq, out);
  run<2>("32 B (2x128)           stride 32", tab, bytes, 32, 32, nq, out);
  run<3>("64 B (2x256)           stride 64", tab, bytes, 64, 64, nq, out);
  run<4>("96 B (3x256)           stride 96", tab, bytes, 96, 96, nq, out);
  run<8>("96 B (3x256)           stride 128", tab, bytes, 128, 96, nq, out);
  run<5>("128 B (4x256)          stride 128", tab, bytes, 128, 128, nq, out);
  return 0;
}
