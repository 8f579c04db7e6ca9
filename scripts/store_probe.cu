// How fast can B200 take the raster kernel's store pattern?  P planes of N doubles each; thread q
// writes element q of every plane (the field-major output of eval_raster with 16 fields), plus
// one int plane.  Variants: plain stores, st.cs, two consecutive elements per thread (v2 stores).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
template <int MODE, int P>
__global__ void __launch_bounds__(256) k(double *val, int *leaf, size_t n)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  if (MODE < 2)
    for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
      {
        const double v = (double) q;
        if (MODE == 0) { leaf[q] = (int) q; 
#pragma unroll
          for (int f = 0; f < P; f++) val[(size_t) f * n + q] = v + f; }
        else { __stcs(leaf + q, (int) q);
#pragma unroll
          for (int f = 0; f < P; f++) __stcs(val + (size_t) f * n + q, v + f); }
      }
  else
    for (size_t q = 2 * ((size_t) blockIdx.x * blockDim.x + threadIdx.x); q + 1 < n; q += 2 * stride)
      {
        const double v = (double) q;
        *reinterpret_cast<int2 *>(leaf + q) = make_int2((int) q, (int) q + 1);
#pragma unroll
        for (int f = 0; f < P; f++) *reinterpret_cast<double2 *>(val + (size_t) f * n + q) = make_double2(v + f, v + f + 1);
      }
}
template <int MODE, int P> void run(const char *name, double *val, int *leaf, size_t n, int ctas)
{
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<MODE, P><<<ctas, 256>>>(val, leaf, n); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 3; i++) k<MODE, P><<<ctas, 256>>>(val, leaf, n);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  printf("%-28s planes %2d ctas %5d  %7.3f ms  %7.1f GB/s\n", name, P, ctas, ms, n * (8.0 * P + 4) / ms / 1e6);
}
int main()
{
  const size_t n = (size_t) 8192 * 8192;
  double *val; int *leaf; CK(cudaMalloc(&val, n * 16 * 8)); CK(cudaMalloc(&leaf, n * 4));
  for (int ctas : {148 * 4, 148 * 8, 148 * 16, 148 * 64})
    {
      run<0, 16>("plain", val, leaf, n, ctas);
      run<1, 16>("st.cs", val, leaf, n, ctas);
      run<2, 16>("v2 (two pixels per thread)", val, leaf, n, ctas);
    }
  run<0, 1>("plain", val, leaf, n, 148 * 8);
  run<0, 4>("plain", val, leaf, n, 148 * 8);
  return 0;
}
