// Probe 2: does a warp-cooperative gather (several lanes of ONE load instruction covering the
// sectors of one random line) cost one miss-queue entry per line or one per sector?
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31; return z; }
__device__ __forceinline__ void ld256(const void *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); }
// LPR lanes cooperate on one record of LPR*BYTES bytes (BYTES = 32: ld256, 16: ld128, 8: ld64) in ONE instruction.
// REPS records are fetched per group per loop iteration with REPS independent instructions.
template <int LPR, int BYTES, int REPS>
__global__ void probe(const char *tab, size_t nrec, size_t nrecords_total, uint64_t *out)
{
  uint64_t acc = 0;
  const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t) gridDim.x * blockDim.x;
  const size_t group = tid / LPR, ngroups = nthreads / LPR;
  const int sub = (int) (tid % LPR);
  const size_t stride = (size_t) LPR * BYTES;
  for (size_t r0 = group * REPS; r0 < nrecords_total; r0 += ngroups * REPS)
    {
#pragma unroll
      for (int k = 0; k < REPS; k++)
        {
          const size_t r = mix(r0 + k) % nrec;
          const char *p = tab + r * stride + (size_t) sub * BYTES;
          if (BYTES == 32) { uint64_t a, b, c, d; ld256(p, a, b, c, d); acc += a ^ b ^ c ^ d; }
          if (BYTES == 16) { ulonglong2 v = __ldg((const ulonglong2 *) p); acc += v.x ^ v.y; }
          if (BYTES == 8) { acc += __ldg((const unsigned long long *) p); }
          if (BYTES == 4) { acc += __ldg((const unsigned *) p); }
        }
    }
  if (acc == 0x123456789abcdefull) out[0] = acc;
}
template <int LPR, int BYTES, int REPS> void run(const char *name, const char *tab, size_t bytes, size_t nrecords, uint64_t *out)
{
  const size_t stride = (size_t) LPR * BYTES, nrec = bytes / stride;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<LPR, BYTES, REPS><<<148 * 8, 256>>>(tab, nrec, nrecords, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  probe<LPR, BYTES, REPS><<<148 * 8, 256>>>(tab, nrec, nrecords, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-44s %7.3f ms  %7.2f Grecords/s  %7.1f Gsectors/s  useful %7.1f GB/s\n", name, ms, nrecords / ms / 1e6,
         nrecords * ((stride + 31) / 32) / ms / 1e6, nrecords * stride / ms / 1e6);
}
int main(int argc, char **argv)
{
  size_t tab_mb = argc > 1 ? atoi(argv[1]) : 4096;
  const size_t bytes = tab_mb << 20, nr = 100000000;
  char *tab; uint64_t *out; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 1, bytes)); CK(cudaMalloc(&out, 8));
  printf("table %zu MB\n", tab_mb);
  run<1, 32, 1>("32 B record : 1 lane  x ld256", tab, bytes, nr, out);
  run<1, 32, 4>("32 B record : 1 lane  x ld256, 4 in flight", tab, bytes, nr, out);
  run<2, 32, 1>("64 B record : 2 lanes x ld256 (1 instr)", tab, bytes, nr, out);
  run<2, 32, 4>("64 B record : 2 lanes x ld256, 4 in flight", tab, bytes, nr, out);
  run<4, 32, 1>("128 B record: 4 lanes x ld256 (1 instr)", tab, bytes, nr, out);
  run<4, 32, 4>("128 B record: 4 lanes x ld256, 4 in flight", tab, bytes, nr, out);
  run<4, 16, 1>("64 B record : 4 lanes x ld128 (1 instr)", tab, bytes, nr, out);
  run<8, 16, 1>("128 B record: 8 lanes x ld128 (1 instr)", tab, bytes, nr, out);
  run<8, 16, 4>("128 B record: 8 lanes x ld128, 4 in flight", tab, bytes, nr, out);
  run<16, 8, 1>("128 B record: 16 lanes x ld64 (1 instr)", tab, bytes, nr, out);
  run<32, 4, 1>("128 B record: 32 lanes x ld32 (1 instr)", tab, bytes, nr, out);
  run<32, 4, 4>("128 B record: 32 lanes x ld32, 4 in flight", tab, bytes, nr, out);
  return 0;
}
