// Which load flavour makes a B200 fetch only the 32-byte sector it needs?  Random 32-byte gathers
// from a table that exceeds L2 but stays inside the TLB reach (default 224 MB), one flavour per
// launch; run under ncu for lts sectors / dram bytes per gather.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe3 gather_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31; return z; }
#define LD256(NAME, PTX) __device__ __forceinline__ uint64_t NAME(const void *p) { uint64_t a, b, c, d; asm volatile(PTX " {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p)); return a ^ b ^ c ^ d; }
LD256(ld_nc, "ld.global.nc.v4.u64")
LD256(ld_ca, "ld.global.ca.v4.u64")
LD256(ld_cg, "ld.global.cg.v4.u64")
LD256(ld_cs, "ld.global.cs.v4.u64")
LD256(ld_lu, "ld.global.lu.v4.u64")
LD256(ld_cv, "ld.global.cv.v4.u64")
LD256(ld_nc_na, "ld.global.nc.L1::no_allocate.v4.u64")
LD256(ld_nc_64, "ld.global.nc.L2::64B.v4.u64")
LD256(ld_nc_128, "ld.global.nc.L2::128B.v4.u64")
LD256(ld_nc_ef, "ld.global.nc.L1::evict_first.v4.u64")
LD256(ld_plain, "ld.global.v4.u64")
__device__ __forceinline__ uint64_t ld_4(const void *p) { return __ldg((const unsigned *) p); }
__device__ __forceinline__ uint64_t ld_4cg(const void *p) { return __ldcg((const unsigned *) p); }
template <int MODE>
__global__ void probe(const char *tab, size_t nrec, size_t nq, uint64_t *out)
{
  uint64_t acc = 0;
  const size_t step = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += step)
    {
      const char *p = tab + (mix(q) % nrec) * 32;
      if (MODE == 0) acc += ld_nc(p);
      if (MODE == 1) acc += ld_ca(p);
      if (MODE == 2) acc += ld_cg(p);
      if (MODE == 3) acc += ld_cs(p);
      if (MODE == 4) acc += ld_lu(p);
      if (MODE == 5) acc += ld_cv(p);
      if (MODE == 6) acc += ld_nc_na(p);
      if (MODE == 7) acc += ld_nc_64(p);
      if (MODE == 8) acc += ld_nc_128(p);
      if (MODE == 9) acc += ld_nc_ef(p);
      if (MODE == 10) acc += ld_plain(p);
      if (MODE == 11) acc += ld_4(p);
      if (MODE == 12) acc += ld_4cg(p);
    }
  if (acc == 0x123456789abcdefull) out[0] = acc;
}
template <int MODE> void run(const char *name, const char *tab, size_t bytes, size_t nq, uint64_t *out)
{
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe<MODE><<<148 * 8, 256>>>(tab, bytes / 32, nq, out);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("mode %2d %-40s %7.3f ms  %7.2f Ggather/s\n", MODE, name, ms, nq / ms / 1e6);
}
int main(int argc, char **argv)
{
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  size_t tab_mb = argc > 2 ? atoi(argv[2]) : 224;
  if (gran) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); printf("set granularity %d -> %s\n", gran, cudaGetErrorString(e)); }
  size_t g = 0; CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity)); printf("L2 fetch granularity limit: %zu, table %zu MB\n", g, tab_mb);
  const size_t bytes = tab_mb << 20, nq = 100000000;
  char *tab; uint64_t *out; CK(cudaMalloc(&tab, bytes)); CK(cudaMemset(tab, 1, bytes)); CK(cudaMalloc(&out, 8));
  run<0>("warm", tab, bytes, nq, out);
  run<0>("ld.global.nc.v4.u64", tab, bytes, nq, out);
  run<1>("ld.global.ca.v4.u64", tab, bytes, nq, out);
  run<2>("ld.global.cg.v4.u64", tab, bytes, nq, out);
  run<3>("ld.global.cs.v4.u64", tab, bytes, nq, out);
  run<4>("ld.global.lu.v4.u64", tab, bytes, nq, out);
  run<5>("ld.global.cv.v4.u64", tab, bytes, nq, out);
  run<6>("ld.global.nc.L1::no_allocate.v4.u64", tab, bytes, nq, out);
  run<7>("ld.global.nc.L2::64B.v4.u64", tab, bytes, nq, out);
  run<8>("ld.global.nc.L2::128B.v4.u64", tab, bytes, nq, out);
  run<9>("ld.global.nc.L1::evict_first.v4.u64", tab, bytes, nq, out);
  run<10>("ld.global.v4.u64", tab, bytes, nq, out);
  run<11>("__ldg 4 B", tab, bytes, nq, out);
  run<12>("__ldcg 4 B", tab, bytes, nq, out);
  return 0;
}
