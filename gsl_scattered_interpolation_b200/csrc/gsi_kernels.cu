/* gsi_kernels.cu -- sm_100a kernels of libgsl_scattered_cuda.
 *
 *   k_locate  fast path: seed from the uniform grid, walk across the edge with the most
 *             negative barycentric weight, accept the leaf only when all three weights
 *             clear the leaf's safety margin, then blend.  The weights are the reference's
 *             own (calculate_bary_coords, linear_simplex.c:607-651) so the accepted result
 *             is bit-identical to find_leaf + interp_point; everything else is deferred.
 *   k_exact   exact replay of find_leaf/_find_leaf (linear_simplex.c:331-402) on the
 *             flattened history DAG + interp_point (linear_simplex.c:678-711): the
 *             reference's edge/vertex tie-breaks, cage leaves, stale-coords quirk.  Runs on
 *             the deferred list, on grid-cell centres (seeding), or on every query.
 *
 * This path is gather/latency bound: no tensor cores, no FMA contraction (gsi_core.h).
 */
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "gsi_core.h"
#include "gsi_kernels.h"
#include "gsi_layout.h"

namespace
{

__device__ __forceinline__ void
fetch_rec(gsi_rec &r, const gsi_rec *base, int idx)
{
  const double2 *p = reinterpret_cast<const double2 *>(base + idx);
  const double2 a = __ldg(p);
  const double2 b = __ldg(p + 1);
  const double2 c = __ldg(p + 2);
  const int4 d = __ldg(reinterpret_cast<const int4 *>(p + 3));
  r.x0x = a.x; r.x0y = a.y;
  r.u00 = b.x; r.u01 = b.y;
  r.u11 = c.x; r.l10 = c.y;
  r.lnk[0] = d.x; r.lnk[1] = d.y; r.lnk[2] = d.z;
  r.meta = (uint32_t) d.w;
}

#define GSI_FETCH_DEV(dst, idx) fetch_rec((dst), dag, (idx))
GSI_DEFINE_DESCEND(descend_dev, __device__ __forceinline__, const gsi_rec *, GSI_FETCH_DEV)

/* ---- query sources -------------------------------------------------------------------- */
struct src_arrays
{
  const double *x, *y;
  __device__ __forceinline__ void get(size_t q, double &qx, double &qy) const
  {
    qx = __ldcs(x + q); /* streamed once: do not let it evict the triangulation */
    qy = __ldcs(y + q);
  }
};

/* the reference's raster driver, scattered_interp_example.c:184-199: x = min + step*i with
 * two roundings, x outer / y inner, result x-major */
struct src_raster
{
  double x0, dx, y0, dy;
  unsigned long long ny, first;
  __device__ __forceinline__ void get(size_t q, double &qx, double &qy) const
  {
    const unsigned long long p = first + q, i = p / ny, j = p - i * ny;
    qx = __dadd_rn(x0, __dmul_rn(dx, (double) i));
    qy = __dadd_rn(y0, __dmul_rn(dy, (double) j));
  }
};

struct src_cells /* grid-cell centres, for seeding */
{
  gsi_view V;
  __device__ __forceinline__ void get(size_t q, double &qx, double &qy) const
  {
    const unsigned g = (unsigned) V.grid_g;
    const unsigned iy = (unsigned) (q / g), ix = (unsigned) (q - (size_t) iy * g);
    const double u = V.u0 + ((double) ix + 0.5) / V.inv_du;
    const double v = V.v0 + ((double) iy + 0.5) / V.inv_dv;
    qx = V.xf.cx + u / V.xf.sx;
    qy = V.xf.cy + v / V.xf.sy;
  }
};

/* ---- blend + store (interp_point, linear_simplex.c:689-710), F fields, field-major out --- */
__device__ __forceinline__ void
blend_store(const gsi_view &V, const int4 ax, double c0, double c1, double *val, size_t ld,
            size_t q)
{
  const int F = V.n_fields;
  const int ha = ax.x >= 0, hb = ax.y >= 0, hx = ax.z >= 0;
  const double *za = V.z + (size_t) (ha ? ax.x : 0) * F;
  const double *zb = V.z + (size_t) (hb ? ax.y : 0) * F;
  const double *zx = V.z + (size_t) (hx ? ax.z : 0) * F;
  if (F == 1)
    {
      const double a = ha ? __ldg(za) : 0.0, b = hb ? __ldg(zb) : 0.0, x = hx ? __ldg(zx) : 0.0;
      __stcs(val + q, gsi_blend(c0, c1, ha, a, hb, b, hx, x));
      return;
    }
  if ((F & 1) == 0)
    {
      const double2 *pa = reinterpret_cast<const double2 *>(za);
      const double2 *pb = reinterpret_cast<const double2 *>(zb);
      const double2 *px = reinterpret_cast<const double2 *>(zx);
#pragma unroll 4
      for (int f = 0; f < F; f += 2)
        {
          double2 a = make_double2(0, 0), b = a, x = a;
          if (ha) a = __ldg(pa + (f >> 1));
          if (hb) b = __ldg(pb + (f >> 1));
          if (hx) x = __ldg(px + (f >> 1));
          __stcs(val + (size_t) f * ld + q, gsi_blend(c0, c1, ha, a.x, hb, b.x, hx, x.x));
          __stcs(val + (size_t) (f + 1) * ld + q, gsi_blend(c0, c1, ha, a.y, hb, b.y, hx, x.y));
        }
      return;
    }
  for (int f = 0; f < F; f++)
    {
      const double a = ha ? __ldg(za + f) : 0.0, b = hb ? __ldg(zb + f) : 0.0,
                   x = hx ? __ldg(zx + f) : 0.0;
      __stcs(val + (size_t) f * ld + q, gsi_blend(c0, c1, ha, a, hb, b, hx, x));
    }
}

__device__ __forceinline__ void
store_nan(const gsi_view &V, double *val, size_t ld, size_t q)
{
  for (int f = 0; f < V.n_fields; f++) val[(size_t) f * ld + q] = CUDART_NAN;
}

/* ---- exact replay of one query ----------------------------------------------------------- */
template <bool WANT_SLOT>
__device__ __forceinline__ int
exact_one(const gsi_view &V, double qx, double qy, double *val, int *leaf_out, size_t ld,
          size_t q, unsigned *counters)
{
  double c0 = 0, c1 = 0;
  int last = -1;
  gsi_rec lr;
  const int node = descend_dev(V.dag, &V.xf, qx, qy, &c0, &c1, &last, &lr);
  if (node < 0)
    {
      if (WANT_SLOT) return 0;
      if (leaf_out) leaf_out[q] = -1;
      if (val) store_nan(V, val, ld, q);
      atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
      return -1;
    }
  const int slot = lr.lnk[0];
  if (WANT_SLOT) return slot;
  /* interp_point recomputes the leaf's coordinates unless the leaf is singular, in which
     case accel->coords keeps whatever the descent left there (linear_simplex.c:641-642,687) */
  if (!(lr.meta & GSI_META_SINGULAR) && last != node)
    gsi_solve(&V.xf, &lr, (int) (lr.meta & GSI_META_SWAP), qx, qy, &c0, &c1);
  if (leaf_out) leaf_out[q] = node;
  if (val)
    {
      const int4 ax = __ldg(reinterpret_cast<const int4 *>(V.aux) + slot);
      blend_store(V, ax, c0, c1, val, ld, q);
    }
  return slot;
}

/* exact replay on every query of a source */
template <class SRC>
__global__ void __launch_bounds__(128)
k_exact_all(const gsi_view V, const SRC src, size_t nq, double *val, int *leaf_out, size_t ld,
            unsigned *counters)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride)
    {
      double qx, qy;
      src.get(q, qx, qy);
      exact_one<false>(V, qx, qy, val, leaf_out, ld, q, counters);
    }
}

/* exact replay on the deferred list written by k_locate */
template <class SRC>
__global__ void __launch_bounds__(128)
k_exact_list(const gsi_view V, const SRC src, const unsigned *list, double *val, int *leaf_out,
             size_t ld, unsigned *counters)
{
  const unsigned count = counters[GSI_CNT_DEFERRED];
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    {
      const size_t q = list[t];
      double qx, qy;
      src.get(q, qx, qy);
      exact_one<false>(V, qx, qy, val, leaf_out, ld, q, counters);
    }
}

/* seeds[cell] = leaf slot containing the cell centre */
__global__ void __launch_bounds__(128)
k_seed_cells(const gsi_view V, int32_t *seeds, unsigned *counters)
{
  const src_cells src{V};
  const size_t ncell = (size_t) V.grid_g * V.grid_g;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += stride)
    {
      double qx, qy;
      src.get(q, qx, qy);
      seeds[q] = exact_one<true>(V, qx, qy, nullptr, nullptr, 0, q, counters);
    }
}

/* ---- fast path ----------------------------------------------------------------------------- */
template <class SRC, bool STATS>
__global__ void __launch_bounds__(GSI_LOCATE_THREADS)
k_locate(const gsi_view V, const SRC src, size_t nq, double *val, int *leaf_out, size_t ld,
         unsigned *list, unsigned *counters, int max_steps)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31u;
  /* warp-uniform trip count so the ballot below is always converged */
  for (size_t base = (size_t) blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < nq;
       base += stride)
    {
      const size_t q = base + lane;
      const bool live = q < nq;
      bool defer = false;
      if (live)
        {
          double qx, qy;
          src.get(q, qx, qy);
          /* grid cell in standardised coordinates (a hint only: no exactness needed) */
          const double u = (qx - V.xf.cx) * V.xf.sx, v = (qy - V.xf.cy) * V.xf.sy;
          const double gmax = (double) (V.grid_g - 1);
          const double fx = fmin(fmax((u - V.u0) * V.inv_du, 0.0), gmax);
          const double fy = fmin(fmax((v - V.v0) * V.inv_dv, 0.0), gmax);
          /* fmin/fmax drop NaN, so non-finite queries get cell 0 and are deferred below */
          int slot = __ldg(V.seeds + (size_t) __double2int_rd(fy) * V.grid_g + __double2int_rd(fx));
          defer = true;
          if (isfinite(qx) && isfinite(qy))
            {
              for (int step = 0; step < max_steps; step++)
                {
                  gsi_rec r;
                  fetch_rec(r, V.leaf, slot);
                  const double tau = (double) __uint_as_float(r.meta & 0x7fffffffu);
                  double c0, c1;
                  gsi_solve(&V.xf, &r, (int) (r.meta >> 31), qx, qy, &c0, &c1);
                  const double c2 = __dsub_rn(1.0, __dadd_rn(__dadd_rn(0.0, c0), c1));
                  if ((c0 >= tau) && (c1 >= tau) && (c2 >= tau))
                    {
                      /* strictly inside, `delta` away from every edge: the reference
                         descent ends in this leaf */
                      const int4 ax = __ldg(reinterpret_cast<const int4 *>(V.aux) + slot);
                      if (leaf_out) __stcs(leaf_out + q, ax.w);
                      if (val) blend_store(V, ax, c0, c1, val, ld, q);
                      defer = false;
                      if (STATS)
                        atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS),
                                  (unsigned long long) step + 1ull);
                      break;
                    }
                  int e = 0;
                  double m = c0;
                  if (c1 < m) { m = c1; e = 1; }
                  if (c2 < m) { m = c2; e = 2; }
                  if (!(m <= -tau)) break; /* within the margin of an edge (or NaN): defer */
                  const int nb = GSI_LNK(r, e);
                  if (nb < 0) break;       /* beyond the cage hull: exact path decides */
                  slot = nb;
                }
            }
        }
      /* warp-aggregated append to the deferred list */
      const unsigned mask = __ballot_sync(0xffffffffu, defer);
      if (mask)
        {
          unsigned pos = 0;
          if (lane == (unsigned) (__ffs(mask) - 1))
            pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(mask));
          pos = __shfl_sync(0xffffffffu, pos, __ffs(mask) - 1);
          if (defer) list[pos + __popc(mask & ((1u << lane) - 1u))] = (unsigned) q;
        }
    }
}

/* ---- synthetic inputs / verification helpers ------------------------------------------------ */
__global__ void __launch_bounds__(256)
k_gen_uniform(uint64_t seed, uint64_t first, size_t nq, double xlo, double xw, double ylo,
              double yw, double *xq, double *yq)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride)
    {
      const uint64_t k = first + q;
      xq[q] = __dadd_rn(xlo, __dmul_rn(xw, gsi_uniform(seed, 2 * k)));
      yq[q] = __dadd_rn(ylo, __dmul_rn(yw, gsi_uniform(seed, 2 * k + 1)));
    }
}

/* order-independent checksum: sum of leaf ids and wrapping sum of value bit patterns */
__global__ void __launch_bounds__(256)
k_checksum(const int *leaf, const double *val, size_t n, unsigned long long *out)
{
  unsigned long long s_leaf = 0, s_val = 0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    {
      if (leaf) s_leaf += (unsigned long long) (long long) leaf[q];
      if (val) s_val += (unsigned long long) __double_as_longlong(val[q]);
    }
  for (int o = 16; o > 0; o >>= 1)
    {
      s_leaf += __shfl_xor_sync(0xffffffffu, s_leaf, o);
      s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
    }
  if ((threadIdx.x & 31) == 0)
    {
      atomicAdd(out, s_leaf);
      atomicAdd(out + 1, s_val);
    }
}

/* count mismatches between two result sets (fast path vs exact replay cross-check) */
__global__ void __launch_bounds__(256)
k_compare(const int *leaf_a, const int *leaf_b, const double *val_a, const double *val_b,
          size_t n, size_t n_val, unsigned long long *out)
{
  unsigned long long bad_leaf = 0, bad_val = 0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride)
    if (leaf_a[q] != leaf_b[q]) bad_leaf++;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < n_val; q += stride)
    if (__double_as_longlong(val_a[q]) != __double_as_longlong(val_b[q])) bad_val++;
  for (int o = 16; o > 0; o >>= 1)
    {
      bad_leaf += __shfl_xor_sync(0xffffffffu, bad_leaf, o);
      bad_val += __shfl_xor_sync(0xffffffffu, bad_val, o);
    }
  if ((threadIdx.x & 31) == 0)
    {
      if (bad_leaf) atomicAdd(out, bad_leaf);
      if (bad_val) atomicAdd(out + 1, bad_val);
    }
}

__global__ void
k_accumulate_deferred(unsigned *counters)
{
  const unsigned c = counters[GSI_CNT_DEFERRED];
  unsigned long long *tot = reinterpret_cast<unsigned long long *>(counters + GSI_CNT_DEFERRED_TOTAL);
  *tot += c;
}

inline int
grid_for(size_t n, int threads, int sm_count, int per_sm)
{
  const size_t want = (n + threads - 1) / threads;
  const size_t cap = (size_t) sm_count * per_sm;
  return (int) (want < cap ? (want ? want : 1) : cap);
}

template <class SRC>
cudaError_t
launch_eval(const gsi_view &V, const SRC &src, size_t nq, double *val, int *leaf_out, size_t ld,
            const gsi_launch_cfg &cfg, cudaStream_t st)
{
  if (nq == 0) return cudaSuccess;
  if (nq > cfg.list_capacity && cfg.mode != GSI_MODE_EXACT) return cudaErrorInvalidValue;
  if (cfg.mode == GSI_MODE_EXACT)
    {
      k_exact_all<SRC><<<grid_for(nq, 128, cfg.sm_count, 16), 128, 0, st>>>(
          V, src, nq, val, leaf_out, ld, cfg.counters);
      return cudaGetLastError();
    }
  cudaError_t e = cudaMemsetAsync(cfg.counters + GSI_CNT_DEFERRED, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  const int g1 = grid_for(nq, GSI_LOCATE_THREADS, cfg.sm_count, cfg.locate_ctas_per_sm);
  if (cfg.stats)
    k_locate<SRC, true><<<g1, GSI_LOCATE_THREADS, 0, st>>>(V, src, nq, val, leaf_out, ld, cfg.list,
                                                          cfg.counters, cfg.max_steps);
  else
    k_locate<SRC, false><<<g1, GSI_LOCATE_THREADS, 0, st>>>(V, src, nq, val, leaf_out, ld,
                                                           cfg.list, cfg.counters, cfg.max_steps);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_exact_list<SRC><<<cfg.sm_count * 4, 128, 0, st>>>(V, src, cfg.list, val, leaf_out, ld,
                                                      cfg.counters);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_accumulate_deferred<<<1, 1, 0, st>>>(cfg.counters);
  return cudaGetLastError();
}

} // namespace

cudaError_t
gsi_launch_eval_arrays(const gsi_view &V, const double *xq, const double *yq, size_t nq,
                       double *val, int *leaf_out, size_t ld, const gsi_launch_cfg &cfg,
                       cudaStream_t st)
{
  return launch_eval(V, src_arrays{xq, yq}, nq, val, leaf_out, ld, cfg, st);
}

cudaError_t
gsi_launch_eval_raster(const gsi_view &V, double x0, double dx, double y0, double dy,
                       unsigned long long ny, size_t q_first, size_t nq, double *val,
                       int *leaf_out, size_t ld, const gsi_launch_cfg &cfg, cudaStream_t st)
{
  return launch_eval(V, src_raster{x0, dx, y0, dy, ny, (unsigned long long) q_first}, nq, val,
                     leaf_out, ld, cfg, st);
}

cudaError_t
gsi_launch_seed_cells(const gsi_view &V, int32_t *seeds, unsigned *counters, int sm_count,
                      cudaStream_t st)
{
  const size_t ncell = (size_t) V.grid_g * V.grid_g;
  k_seed_cells<<<grid_for(ncell, 128, sm_count, 16), 128, 0, st>>>(V, seeds, counters);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_gen_uniform(uint64_t seed, uint64_t first, size_t nq, double xlo, double xhi,
                       double ylo, double yhi, double *xq, double *yq, int sm_count,
                       cudaStream_t st)
{
  if (nq == 0) return cudaSuccess;
  k_gen_uniform<<<grid_for(nq, 256, sm_count, 8), 256, 0, st>>>(seed, first, nq, xlo, xhi - xlo,
                                                                 ylo, yhi - ylo, xq, yq);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_checksum(const int *leaf, const double *val, size_t n, unsigned long long *out2,
                    int sm_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), st);
  if (e != cudaSuccess || n == 0) return e;
  k_checksum<<<grid_for(n, 256, sm_count, 8), 256, 0, st>>>(leaf, val, n, out2);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_compare(const int *leaf_a, const int *leaf_b, const double *val_a, const double *val_b,
                   size_t n, size_t n_val, unsigned long long *out2, int sm_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  const size_t m = n > n_val ? n : n_val;
  if (m == 0) return cudaSuccess;
  k_compare<<<grid_for(m, 256, sm_count, 8), 256, 0, st>>>(leaf_a, leaf_b, val_a, val_b, n, n_val,
                                                           out2);
  return cudaGetLastError();
}
