/* gsi_kernels.cu -- sm_100a kernels of libgsl_scattered_cuda.
 *
 *   k_walkc        fast path for query arrays (default): lanes decoupled from queries by a warp-wide
 *                  queue, one walk step per iteration for whatever query each lane holds; the WARP
 *                  copies the 32 leaf records its lanes visit into shared memory (cp.async, one
 *                  128-byte line per record), the seed-block bases sit in shared memory.  Seed from
 *                  the uniform grid, cross the edge with the most negative barycentric weight,
 *                  accept the leaf only when all three weights clear the leaf's safety margin, then
 *                  blend.  The weights are the reference's own (calculate_bary_coords,
 *                  linear_simplex.c:607-651), so an accepted result is bit-identical to find_leaf +
 *                  interp_point; everything else is deferred.
 *   k_walkf / k_walk / k_walkp
 *                  opt-in variants of the same step (walk_variant 3 / 0 / 1): fresh batches + a ring
 *                  of walkers; the round-1 kernel with per-lane 256-bit gathers; its prefetching form.
 *   k_locate       the same per-query logic, one thread per query over compact pixel tiles: the
 *                  default for rasters (coherent queries).
 *   k_exact_*      exact replay of find_leaf/_find_leaf (linear_simplex.c:331-402) on the flattened
 *                  history DAG + interp_point (linear_simplex.c:678-711): the reference's
 *                  edge/vertex tie-breaks, cage leaves, stale-coords quirk.  Runs on the deferred
 *                  list (four lanes per query, k_exact_list4 / k_exact_overflow4), on grid-cell
 *                  centres (seeding), on the data nodes (k_memo_fill) or on every query
 *                  (k_exact_all); starts at the cell's certified entry node (k_entry_cells).
 *   k_memo_list    queries of the deferred list that coincide with a data node are answered from the
 *                  vertex memo, the others compacted for the replay.
 *   k_seed_cells, k_pack_seeds, k_entry_cells, k_leaf_rcp, k_memo_fill
 *                  the acceleration structure, filled on the device at init (gsi_layout.h).
 *   k_bin -> k_tile_walk -> k_unpermute
 *                  reorder pipeline for large incoherent query batches on triangulations beyond
 *                  the L2 (chunk-local counting sort by spatial tile, tile-major walk out of
 *                  TMA-staged shared memory, inverse permutation); opt-in.
 *   k_deriv, k_build_locate, k_validate_slab, k_gen_uniform, k_checksum, k_compare
 *                  gradients, the builder's batched point location, slab validation, bench helpers.
 *
 * This path is gather/latency bound: no tensor cores, no FMA contraction (gsi_core.h).
 * Measured bounds and the experiments that did not pay off are in DESIGN.md section 4.
 */
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "gsi_core.h"
#include "gsi_eval_core.h"
#include "gsi_kernels.h"
#include "gsi_layout.h"

namespace
{

/* ---- query sources --------------------------------------------------------------------
 * get(t, q, x, y): launch index t -> output index q and coordinates; false = padding (no query).
 * n_launch(nq)   : number of launch indices needed to cover nq queries. */
struct src_arrays
{
  static constexpr bool coherent = false; /* arbitrary order: lanes-decoupled kernel by default */
  const double *x, *y;
  __device__ __forceinline__ bool get(size_t t, size_t &q, double &qx, double &qy) const
  {
    q = t;
    qx = GSI_LD_QUERY(x + t);
    qy = GSI_LD_QUERY(y + t);
    return true;
  }
  __device__ __forceinline__ void get_linear(size_t q, double &qx, double &qy) const
  {
    qx = GSI_LD_QUERY(x + q);
    qy = GSI_LD_QUERY(y + q);
  }
  __host__ __device__ size_t n_launch(size_t nq) const { return nq; }
};

/* the reference's raster driver, scattered_interp_example.c:184-199: x = min + step*i with
 * two roundings, x outer / y inner, result x-major (pixel = i*ny + j).  A launch covers the
 * columns [i_first, i_first + n_cols); launch indices enumerate 8 x 32 pixel tiles (8 columns
 * x 32 consecutive rows) so that a 256-thread CTA works on a compact patch of the plane
 * (leaf and value lines are reused from L1) while every warp still writes 32 consecutive
 * pixels of one column (full-sector stores). */
struct src_raster
{
  static constexpr bool coherent = true; /* neighbouring indices are neighbouring points */
  double x0, dx, y0, dy;
  unsigned long long ny, i_first, n_cols;
  __device__ __forceinline__ void coords(unsigned long long i, unsigned long long j, double &qx,
                                         double &qy) const
  {
    qx = __dadd_rn(x0, __dmul_rn(dx, (double) (i_first + i)));
    qy = __dadd_rn(y0, __dmul_rn(dy, (double) j));
  }
  __device__ __forceinline__ bool get(size_t t, size_t &q, double &qx, double &qy) const
  {
    const unsigned long long tiles_j = (ny + 31) >> 5;
    const unsigned long long tile = t >> 8, within = t & 255;
    const unsigned long long ti = tile / tiles_j, tj = tile - ti * tiles_j;
    const unsigned long long i = (ti << 3) + (within >> 5), j = (tj << 5) + (within & 31);
    if (i >= n_cols || j >= ny) return false;
    q = i * ny + j;
    coords(i, j, qx, qy);
    return true;
  }
  __device__ __forceinline__ void get_linear(size_t q, double &qx, double &qy) const
  {
    const unsigned long long i = q / ny, j = q - i * ny;
    coords(i, j, qx, qy);
  }
  __host__ __device__ size_t n_launch(size_t nq) const
  {
    (void) nq;
    return (size_t) (((n_cols + 7) >> 3) * ((ny + 31) >> 5)) << 8;
  }
};

/* a deferred query that coincides with a data node: interp_point on the memoised leaf (the leaf is
   not singular, so interp_point solves it afresh, linear_simplex.c:687) */
__device__ __forceinline__ void
memo_answer(const gsi_view &V, int node, int slot, double qx, double qy, double *val, int *leaf_out, size_t ld, size_t q)
{
  if (leaf_out) leaf_out[q] = node;
  if (val)
    {
      const gsi_leaf *lf = gsi_leaf_at(&V, slot);
      gsi_rec r;
      gsi_fetch_rec(&r, lf);
      double c0, c1;
      gsi_solve(&V.xf, &r, (int) (r.meta >> 31), qx, qy, &c0, &c1);
      gsi_leaf_tail t;
      gsi_fetch_tail(&t, lf);
      gsi_blend_store(&V, &t, c0, c1, val, ld, q);
    }
}

/* exact replay on every query of a source */
template <class SRC>
__global__ void __launch_bounds__(128, 4)
k_exact_all(const gsi_view V, const SRC src, size_t nq, double *val, int *leaf_out, size_t ld,
            unsigned *counters)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride)
    {
      double qx, qy;
      src.get_linear(q, qx, qy);
      if (gsi_exact_one(&V, qx, qy, val, leaf_out, ld, q) < 0)
        atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
    }
}

/* exact replay on a deferred list (the one the walk kernels wrote, or what k_memo_list left of it) */
template <class SRC>
__global__ void __launch_bounds__(128, 4)
k_exact_list(const gsi_view V, const SRC src, const unsigned *list, unsigned list_cap, int cnt_word, double *val,
             int *leaf_out, size_t ld, unsigned *counters)
{
  const unsigned count = min(counters[cnt_word], list_cap);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride)
    {
      const size_t q = list[t];
      double qx, qy;
      src.get_linear(q, qx, qy);
      if (gsi_exact_one(&V, qx, qy, val, leaf_out, ld, q) < 0)
        atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
    }
}

/* vertex-memo pass over the deferred list: a query that coincides with a data node gets its answer
 * from the hash table; the others are compacted into list2 for the replay kernels.  (Probing inside
 * the four-lane replay instead leaves the groups of a warp out of step with one another: measured
 * 48 % slower on queries that miss, and no faster on a half-and-half mix, scripts/profile_cfg3_classes.py.) */
template <class SRC>
__global__ void __launch_bounds__(128)
k_memo_list(const gsi_view V, const SRC src, const unsigned *list, unsigned list_cap, unsigned *list2, double *val,
            int *leaf_out, size_t ld, unsigned *counters)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned count = min(counters[GSI_CNT_DEFERRED], list_cap);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned t0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); t0 < count; t0 += stride)
    {
      const unsigned t = t0 + lane;
      bool miss = false;
      unsigned qi = 0;
      if (t < count)
        {
          qi = list[t];
          double qx, qy;
          src.get_linear((size_t) qi, qx, qy);
          int mnode, mslot;
          if (gsi_memo_find(&V, qx, qy, &mnode, &mslot))
            memo_answer(V, mnode, mslot, qx, qy, val, leaf_out, ld, (size_t) qi);
          else
            miss = true;
        }
      const unsigned mm = __ballot_sync(FULL, miss);
      if (mm)
        {
          const int leader = __ffs(mm) - 1;
          unsigned pos = 0;
          if (lane == (unsigned) leader) pos = atomicAdd(counters + GSI_CNT_DEFERRED2, (unsigned) __popc(mm));
          pos = __shfl_sync(FULL, pos, leader);
          if (miss) list2[pos + __popc(mm & ((1u << lane) - 1u))] = qi;
        }
    }
}

/* exact replay of the queries marked in the overflow bitmask (deferred list was full) */
template <class SRC>
__global__ void __launch_bounds__(128, 4)
k_exact_overflow(const gsi_view V, const SRC src, size_t nq, unsigned *mask, double *val,
                 int *leaf_out, size_t ld, unsigned *counters)
{
  if (counters[GSI_CNT_OVERFLOW] == 0) return;
  const size_t n_words = (nq + 31) >> 5;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t w = (size_t) blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride)
    {
      unsigned m = mask[w];
      if (!m) continue;
      mask[w] = 0;
      while (m)
        {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const size_t q = (w << 5) + (size_t) b;
          double qx, qy;
          src.get_linear(q, qx, qy);
          int mnode, mslot;
          if (gsi_memo_find(&V, qx, qy, &mnode, &mslot))
            memo_answer(V, mnode, mslot, qx, qy, val, leaf_out, ld, q);
          else if (gsi_exact_one(&V, qx, qy, val, leaf_out, ld, q) < 0)
            atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
        }
    }
}

/* ---- exact replay, four lanes per query ---------------------------------------------------------------
 * gsi_descend (one thread per query) keeps the current record and up to three child records in
 * registers (128 registers, spills, 16 warps per SM) and tests the children one after the other:
 * the deferred queries of a mixed workload (BASELINE.json configs[2]: 20 % on nodes and on edge
 * midpoints) spend ~50 dependent levels like that.  Here a group of four lanes replays ONE query:
 * lane j fetches and solves child j (lane 3 idles), a ballot finds the first passing child in link
 * order, and only when no child passes are the coordinates gathered to replay the reference's
 * best-worst fallback with its stale-coordinate rule.  One record per lane, the three solves of a
 * level side by side, twice the warps per SM.  The semantics are those of GSI_DEFINE_DESCEND
 * (gsi_core.h), statement for statement; k_exact_all keeps the one-thread form, and the tests
 * compare the two on 100 % of their queries.
 *
 * All four lanes of a group hold the same state (node, accel->coords, last) and take the same
 * branches, so the group-masked shuffles below are always converged. */
__device__ __forceinline__ double
shfl_f64(unsigned mask, double v, int src)
{
  const int lo = __shfl_sync(mask, __double2loint(v), src), hi = __shfl_sync(mask, __double2hiint(v), src);
  return __hiloint2double(hi, lo);
}

/* Returns the leaf node id (>= 0), GSI_DESCEND_OUTSIDE or GSI_DESCEND_LOOP; *slot = the leaf's compact
 * slot, (*c0, *c1) = accel->coords as interp_point will use them (recomputed at the leaf unless the
 * leaf is singular or was the last simplex solved, linear_simplex.c:641-642,687). */
__device__ __forceinline__ int
descend4(const gsi_view &V, unsigned gmask, int gbase, int sub, int start, double qx, double qy, double *c0, double *c1,
         int *slot)
{
  const gsi_rec *dag = V.dag;
  const gsi_xform *t = &V.xf;
  gsi_rec r; /* the record this lane fetched last: the current node's in lane `holder` (in every lane at the start) */
  GSI_FETCH_DAG(r, start);
  if (r.meta & GSI_META_SINGULAR) return GSI_DESCEND_OUTSIDE;
  gsi_solve(t, &r, (int) (r.meta & GSI_META_SWAP), qx, qy, c0, c1);
  int last = start, node = start, guard = 0, holder = sub; /* every lane holds the start record */
  if (start == 0 && !gsi_inside(*c0, *c1)) return GSI_DESCEND_OUTSIDE;
  int k0 = r.lnk[0], k1 = r.lnk[1], k2 = r.lnk[2];
  uint32_t meta = r.meta;
  for (;;)
    {
      const uint32_t type = meta & GSI_META_TYPE_MASK;
      if (type == GSI_LEAF)
        {
          *slot = k0;
          /* interp_point solves the leaf again unless it is singular or was solved last */
          if (!(meta & GSI_META_SINGULAR) && last != node)
            {
              double a, b;
              gsi_solve(t, &r, (int) (r.meta & GSI_META_SWAP), qx, qy, &a, &b); /* meaningful in lane `holder` */
              *c0 = shfl_f64(gmask, a, gbase + (holder & 3));
              *c1 = shfl_f64(gmask, b, gbase + (holder & 3));
            }
          return node;
        }
      if (++guard > GSI_MAX_DESCENT) return GSI_DESCEND_LOOP;
      const int nch = (type == GSI_SUB_DPLUS1) ? 3 : 2;
      if (nch == 2) k2 = 0;
      const int k = sub == 0 ? k0 : (sub == 1 ? k1 : k2);
      const bool active = sub < nch && k != 0;
      bool valid = false;
      double a = 0, b = 0;
      int pass = 0;
      if (active)
        {
          GSI_FETCH_DAG(r, k);
          valid = !(r.meta & GSI_META_SINGULAR);
          if (valid)
            {
              gsi_solve(t, &r, (int) (r.meta & GSI_META_SWAP), qx, qy, &a, &b);
              pass = gsi_inside(a, b);
            }
        }
      const unsigned pm = (__ballot_sync(gmask, pass) >> gbase) & 0xfu;
      int chosen;
      if (pm)
        {
          /* first passing child in link order; the children behind it are never evaluated, so
             accel->coords and `last` are this child's */
          chosen = __ffs(pm) - 1;
          *c0 = shfl_f64(gmask, a, gbase + chosen);
          *c1 = shfl_f64(gmask, b, gbase + chosen);
          node = chosen == 0 ? k0 : (chosen == 1 ? k1 : k2);
          last = node;
        }
      else
        {
          /* no child passes: every child was evaluated in order; a child with link 0 or a singular
             matrix leaves accel->coords as they were (linear_simplex.c:375-395 after :641-642) */
          const unsigned vm = (__ballot_sync(gmask, valid) >> gbase) & 0xfu;
          int best_match = 0;
          double best_worst = -1;
          double e0 = *c0, e1 = *c1;
#pragma unroll
          for (int i = 0; i < 3; i++)
            {
              const double ai = shfl_f64(gmask, a, gbase + i), bi = shfl_f64(gmask, b, gbase + i);
              if (i < nch)
                {
                  if (vm & (1u << i))
                    {
                      e0 = ai;
                      e1 = bi;
                      last = i == 0 ? k0 : (i == 1 ? k1 : k2);
                    }
                  const double worst = gsi_worst(e0, e1);
                  if ((best_worst < 0) || (worst < best_worst))
                    {
                      best_worst = worst;
                      best_match = i;
                    }
                }
            }
          *c0 = e0;
          *c1 = e1;
          chosen = best_match;
          node = chosen == 0 ? k0 : (chosen == 1 ? k1 : k2);
        }
      if (node == 0)
        {
          /* link 0: the reference would re-enter the root */
          GSI_FETCH_DAG(r, 0);
          holder = sub;
          k0 = r.lnk[0]; k1 = r.lnk[1]; k2 = r.lnk[2];
          meta = r.meta;
        }
      else
        {
          /* the chosen child's record sits in lane `chosen`: its links and meta word to everyone */
          holder = chosen;
          k0 = __shfl_sync(gmask, r.lnk[0], gbase + chosen);
          k1 = __shfl_sync(gmask, r.lnk[1], gbase + chosen);
          k2 = __shfl_sync(gmask, r.lnk[2], gbase + chosen);
          meta = (uint32_t) __shfl_sync(gmask, (int) r.meta, gbase + chosen);
        }
    }
}

/* one deferred query per group of four lanes; lane 0 of the group writes the result */
template <bool MEMO, class SRC>
__device__ __forceinline__ void
exact_one4(const gsi_view &V, const SRC &src, size_t q, double *val, int *leaf_out, size_t ld, unsigned *counters,
           unsigned gmask, int gbase, int sub)
{
  double qx, qy, c0 = 0, c1 = 0;
  src.get_linear(q, qx, qy);
  int slot = 0;
  if (MEMO)
    {
    /* a query that coincides with a data node: the four lanes probe the memo alike (overflow path only:
       the list kernels run behind k_memo_list) */
    int mnode, mslot;
    if (gsi_memo_find(&V, qx, qy, &mnode, &mslot))
      {
        if (sub == 0) memo_answer(V, mnode, mslot, qx, qy, val, leaf_out, ld, q);
        return;
      }
  }
  const int start = V.entry ? GSI_LDG(V.entry + gsi_cell_of(&V, qx, qy)) : 0;
  const int node = descend4(V, gmask, gbase, sub, start, qx, qy, &c0, &c1, &slot);
  if (sub != 0) return;
  if (node < 0)
    {
      if (leaf_out) leaf_out[q] = -1;
      if (val)
        for (int f = 0; f < V.n_fields; f++) val[(size_t) f * ld + q] = GSI_NAN;
      atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
      return;
    }
  if (leaf_out) leaf_out[q] = node;
  if (val)
    {
      gsi_leaf_tail tl;
      gsi_fetch_tail(&tl, gsi_leaf_at(&V, slot));
      gsi_blend_store(&V, &tl, c0, c1, val, ld, q);
    }
}

template <class SRC>
__global__ void __launch_bounds__(128, 8)
k_exact_list4(const gsi_view V, const SRC src, const unsigned *list, unsigned list_cap, int cnt_word, double *val,
              int *leaf_out, size_t ld, unsigned *counters)
{
  const unsigned count = min(counters[cnt_word], list_cap);
  const unsigned lane = threadIdx.x & 31u;
  const int sub = (int) (lane & 3u), gbase = (int) (lane & ~3u);
  const unsigned gmask = 0xfu << gbase;
  const unsigned n_groups = (gridDim.x * blockDim.x) >> 2;
  for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; g < count; g += n_groups)
    exact_one4<false>(V, src, (size_t) list[g], val, leaf_out, ld, counters, gmask, gbase, sub);
}

#ifndef GSI_EXACT4P_CTAS
#define GSI_EXACT4P_CTAS 7 /* 72 registers: no spills */
#endif
/* ---- exact replay, four lanes per query, groups refilled as they finish (exact_coop = 2, opt-in) ----
 * In k_exact_list4 the eight groups of a warp start a query together and the warp moves on when the
 * LAST of the eight descents ends: descents differ in length (the deferred queries of a mixed workload
 * take 10 to 40 levels from their entry nodes), so on average only 4.3 of the 8 groups are at work
 * (12.8 active lanes per instruction, ncu).  Here every trip of the loop is ONE level for whatever query
 * each group holds; a group whose query reached its leaf writes the result and takes the next entry of
 * the list at once.  A fresh query enters as a level with a single virtual child, its entry node: the
 * same fetch / solve / ballot instructions as a real level, so fresh and running groups stay in step.
 * Semantics: those of descend4 / GSI_DEFINE_DESCEND, statement for statement (first passing child in
 * link order, best-worst fallback with the stale-coordinate rule, singular or link-0 children fail
 * without a test, link 0 re-enters the root); the tests compare it with the one-thread replay.
 * Measured (scripts/gpu_r2_aa.sh, config 3): exact, but SLOWER than k_exact_list4 (1.97 vs 1.48 ms per 10^6
 * descents, 6 / 7 / 8 CTAs per SM alike): the refill and the leaf paths (dependent loads) run while the other
 * groups of the warp wait, once per trip instead of once per query.  Kept opt-in. */
template <class SRC>
__global__ void __launch_bounds__(128, GSI_EXACT4P_CTAS)
k_exact_list4p(const gsi_view V, const SRC src, const unsigned *list, unsigned list_cap, int cnt_word, double *val,
               int *leaf_out, size_t ld, unsigned *counters)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned count = min(counters[cnt_word], list_cap);
  const unsigned lane = threadIdx.x & 31u;
  const int sub = (int) (lane & 3u), gbase = (int) (lane & ~3u);
  const unsigned gmask = 0xfu << gbase;
  const unsigned n_groups = (gridDim.x * blockDim.x) >> 2;
  const gsi_rec *dag = V.dag;
  const gsi_xform *t = &V.xf;
  unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;

  bool have = false, fresh = false;
  size_t q = 0;
  double qx = 0, qy = 0, c0 = 0, c1 = 0;
  int node = 0, last = -1, holder = 0, guard = 0, k0 = 0, k1 = 0, k2 = 0;
  uint32_t meta = 0;
  gsi_rec r; /* the record this lane fetched last: the current node's in lane `holder` */
  r.meta = GSI_META_SINGULAR;
  for (;;)
    {
      if (!have && g < count)
        {
          q = (size_t) list[g];
          g += n_groups;
          src.get_linear(q, qx, qy);
          /* a virtual node whose only child is the entry node of the query's cell */
          k0 = V.entry ? GSI_LDG(V.entry + gsi_cell_of(&V, qx, qy)) : 0;
          k1 = k2 = 0;
          c0 = c1 = 0;
          last = -1;
          node = -1;
          guard = 0;
          have = fresh = true;
        }
      if (!__any_sync(FULL, have)) break;
      if (!have) continue;

      const uint32_t type = meta & GSI_META_TYPE_MASK;
      if (!fresh && type == GSI_LEAF)
        {
          /* interp_point solves the leaf again unless it is singular or was solved last */
          if (!(meta & GSI_META_SINGULAR) && last != node)
            {
              double a, b;
              gsi_solve(t, &r, (int) (r.meta & GSI_META_SWAP), qx, qy, &a, &b); /* meaningful in lane `holder` */
              c0 = shfl_f64(gmask, a, gbase + (holder & 3));
              c1 = shfl_f64(gmask, b, gbase + (holder & 3));
            }
          if (sub == 0)
            {
              if (leaf_out) leaf_out[q] = node;
              if (val)
                {
                  gsi_leaf_tail tl;
                  gsi_fetch_tail(&tl, gsi_leaf_at(&V, k0)); /* a leaf's lnk[0] is its compact slot */
                  gsi_blend_store(&V, &tl, c0, c1, val, ld, q);
                }
            }
          have = false;
          continue;
        }
      bool outside = !fresh && ++guard > GSI_MAX_DESCENT; /* GSI_DESCEND_LOOP */
      const int nch = fresh ? 1 : ((type == GSI_SUB_DPLUS1) ? 3 : 2);
      if (nch < 3) k2 = 0;
      const int k = sub == 0 ? k0 : (sub == 1 ? k1 : k2);
      const bool active = !outside && sub < nch && (fresh || k != 0);
      bool valid = false;
      double a = 0, b = 0;
      int pass = 0;
      if (active)
        {
          GSI_FETCH_DAG(r, k);
          valid = !(r.meta & GSI_META_SINGULAR);
          if (valid)
            {
              gsi_solve(t, &r, (int) (r.meta & GSI_META_SWAP), qx, qy, &a, &b);
              /* find_leaf tests the root (linear_simplex.c:343-347); a later entry node is a node the
                 descent is known to pass: it is taken with its coordinates, no test */
              pass = (fresh && k != 0) ? 1 : gsi_inside(a, b);
            }
        }
      const unsigned pm = (__ballot_sync(gmask, pass) >> gbase) & 0xfu;
      int chosen = 0;
      if (fresh)
        {
          /* singular entry record, or a root that does not contain the query: outside the cage */
          if (!pm) outside = true;
          else
            {
              c0 = shfl_f64(gmask, a, gbase);
              c1 = shfl_f64(gmask, b, gbase);
              node = k0;
              last = node;
            }
        }
      else if (outside)
        ;
      else if (pm)
        {
          /* first passing child in link order; the children behind it are never evaluated, so
             accel->coords and `last` are this child's */
          chosen = __ffs(pm) - 1;
          c0 = shfl_f64(gmask, a, gbase + chosen);
          c1 = shfl_f64(gmask, b, gbase + chosen);
          node = chosen == 0 ? k0 : (chosen == 1 ? k1 : k2);
          last = node;
        }
      else
        {
          /* no child passes: every child was evaluated in order; a child with link 0 or a singular
             matrix leaves accel->coords as they were (linear_simplex.c:375-395 after :641-642) */
          const unsigned vm = (__ballot_sync(gmask, valid) >> gbase) & 0xfu;
          int best_match = 0;
          double best_worst = -1;
          double e0 = c0, e1 = c1;
#pragma unroll
          for (int i = 0; i < 3; i++)
            {
              const double ai = shfl_f64(gmask, a, gbase + i), bi = shfl_f64(gmask, b, gbase + i);
              if (i < nch)
                {
                  if (vm & (1u << i))
                    {
                      e0 = ai;
                      e1 = bi;
                      last = i == 0 ? k0 : (i == 1 ? k1 : k2);
                    }
                  const double worst = gsi_worst(e0, e1);
                  if ((best_worst < 0) || (worst < best_worst))
                    {
                      best_worst = worst;
                      best_match = i;
                    }
                }
            }
          c0 = e0;
          c1 = e1;
          chosen = best_match;
          node = chosen == 0 ? k0 : (chosen == 1 ? k1 : k2);
        }
      if (outside)
        {
          if (sub == 0)
            {
              if (leaf_out) leaf_out[q] = -1;
              if (val)
                for (int f = 0; f < V.n_fields; f++) val[(size_t) f * ld + q] = GSI_NAN;
              atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
            }
          have = false;
          continue;
        }
      if (node == 0 && !fresh)
        {
          /* link 0: the reference would re-enter the root */
          GSI_FETCH_DAG(r, 0);
          holder = sub;
          k0 = r.lnk[0]; k1 = r.lnk[1]; k2 = r.lnk[2];
          meta = r.meta;
        }
      else
        {
          /* the chosen child's record sits in lane `chosen`: its links and meta word to everyone */
          holder = chosen;
          k0 = __shfl_sync(gmask, r.lnk[0], gbase + chosen);
          k1 = __shfl_sync(gmask, r.lnk[1], gbase + chosen);
          k2 = __shfl_sync(gmask, r.lnk[2], gbase + chosen);
          meta = (uint32_t) __shfl_sync(gmask, (int) r.meta, gbase + chosen);
        }
      fresh = false;
    }
}

/* exact replay of the queries marked in the overflow bitmask, four lanes per query: a warp takes 32
 * mask words at a time (one per lane) and, round after round, the first eight lanes that still hold a
 * marked query hand one each to the warp's eight groups */
template <class SRC>
__global__ void __launch_bounds__(128, 8)
k_exact_overflow4(const gsi_view V, const SRC src, size_t nq, unsigned *mask, double *val, int *leaf_out, size_t ld,
                  unsigned *counters)
{
  if (counters[GSI_CNT_OVERFLOW] == 0) return;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const int sub = (int) (lane & 3u), gbase = (int) (lane & ~3u);
  const unsigned gmask = 0xfu << gbase, group = lane >> 2;
  const size_t n_words = (nq + 31) >> 5;
  const size_t wstride = (((size_t) gridDim.x * blockDim.x) >> 5) * 32;
  for (size_t w0 = ((((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5); w0 < n_words; w0 += wstride)
    {
      const size_t w = w0 + lane;
      unsigned m = w < n_words ? mask[w] : 0u;
      if (m) mask[w] = 0;
      for (;;)
        {
          const unsigned have = __ballot_sync(FULL, m != 0);
          if (!have) break;
          /* the lane that serves this group in this round: the (group+1)-th lane with a marked query */
          const unsigned srcl = __fns(have, 0, (int) group + 1);
          unsigned bit = 0;
          const bool giver = m != 0 && (unsigned) __popc(have & ((1u << lane) - 1u)) < 8u;
          if (giver)
            {
              bit = (unsigned) __ffs(m) - 1u;
              m &= m - 1u;
            }
          const unsigned sl = srcl < 32u ? srcl : 0u;
          const unsigned b = __shfl_sync(FULL, bit, (int) sl);
          if (srcl < 32u)
            exact_one4<true>(V, src, ((w0 + sl) << 5) + (size_t) b, val, leaf_out, ld, counters, gmask, gbase, sub);
        }
    }
}

/* seeds[cell] = leaf slot containing the cell centre (exact replay of the centre) */
__global__ void __launch_bounds__(128)
k_seed_cells(const gsi_view V, int32_t *seeds)
{
  const unsigned g = (unsigned) V.grid_g;
  const size_t ncell = (size_t) g * g;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += stride)
    {
      double qx, qy;
      gsi_cell_centre(&V, q, &qx, &qy);
      const int slot = gsi_exact_one(&V, qx, qy, nullptr, nullptr, 0, q);
      seeds[q] = slot < 0 ? 0 : slot;
    }
}

/* compact seed table (gsi_layout.h): one WARP per block of B x B cells (B = 2^seed_shift).  The base
 * of a block is the seed of its centre cell minus 32768, so the 16-bit window covers the block's own
 * leaves and their neighbours in Morton order on both sides.  A cell whose seed (the leaf holding its
 * centre) lies outside the window -- a leaf whose centroid falls into a block far away along the Morton
 * curve, typically just across a coarse Morton boundary -- gets the in-window seed of the nearest cell of
 * its row or column instead: the seed is only the START of the walk, which then needs a step or two
 * more for the queries of that cell, and no query ever waits for a second, dependent table look-up.
 * 0xffff (escape to the int32 table) remains for cells with no in-window seed in their row and column. */
__global__ void __launch_bounds__(128)
k_pack_seeds(const gsi_view V, const int32_t *seeds, uint16_t *seed16, int32_t *seedbase)
{
  const unsigned g = (unsigned) V.grid_g, gb = (unsigned) V.grid_gb, sh = (unsigned) V.seed_shift, B = 1u << sh;
  const unsigned lane = threadIdx.x & 31u;
  const size_t nblk = (size_t) gb * gb;
  const size_t wstride = ((size_t) gridDim.x * blockDim.x) >> 5;
  for (size_t b = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nblk; b += wstride)
    {
      const unsigned by = (unsigned) (b / gb), bx = (unsigned) (b - (size_t) by * gb);
      const unsigned x0 = bx << sh, y0 = by << sh, x1 = min(x0 + B, g), y1 = min(y0 + B, g);
      const unsigned w = x1 - x0, ncell = w * (y1 - y0);
      const int base = max(seeds[(size_t) ((y0 + y1) >> 1) * g + ((x0 + x1) >> 1)] - 32768, 0);
      if (lane == 0) seedbase[b] = base;
      for (unsigned k = lane; k < ncell; k += 32)
        {
          const unsigned y = y0 + k / w, x = x0 + k % w;
          int d = seeds[(size_t) y * g + x] - base;
          for (unsigned r = 1; !(d >= 0 && d < (int) GSI_SEED16_ESC) && r < B; r++)
            {
              int c;
              if (x >= x0 + r && (c = seeds[(size_t) y * g + (x - r)] - base) >= 0 && c < (int) GSI_SEED16_ESC) d = c;
              else if (x + r < x1 && (c = seeds[(size_t) y * g + (x + r)] - base) >= 0 && c < (int) GSI_SEED16_ESC) d = c;
              else if (y >= y0 + r && (c = seeds[(size_t) (y - r) * g + x] - base) >= 0 && c < (int) GSI_SEED16_ESC) d = c;
              else if (y + r < y1 && (c = seeds[(size_t) (y + r) * g + x] - base) >= 0 && c < (int) GSI_SEED16_ESC) d = c;
            }
          seed16[(size_t) y * g + x] = (d >= 0 && d < (int) GSI_SEED16_ESC) ? (uint16_t) d : (uint16_t) GSI_SEED16_ESC;
        }
    }
}

/* entry[cell] = deepest DAG node every query of the cell is certain to pass (gsi_eval_core.h) */
__global__ void __launch_bounds__(128)
k_entry_cells(const gsi_view V, int32_t *entry)
{
  const unsigned g = (unsigned) V.grid_g;
  const size_t ncell = (size_t) g * g;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += stride)
    entry[q] = gsi_entry_of_cell(&V, q);
}

/* vertex memo: the exact replay's answer for every key the flattener wrote (gsi_layout.h) */
__global__ void __launch_bounds__(128)
k_memo_fill(const gsi_view V, gsi_memo *memo)
{
  const size_t n = (size_t) V.memo_mask + 1;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      gsi_memo *e = memo + i;
      if (e->state != GSI_MEMO_PENDING) continue;
      const double qx = e->x, qy = e->y;
      double c0 = 0, c1 = 0;
      int last = -1;
      gsi_rec lr;
      const int start = V.entry ? GSI_LDG(V.entry + gsi_cell_of(&V, qx, qy)) : 0;
      const int node = gsi_descend(V.dag, &V.xf, start, qx, qy, &c0, &c1, &last, &lr, 0);
      /* a singular leaf keeps the descent's stale coordinates (linear_simplex.c:641-642): not memoised */
      if (node < 0 || (lr.meta & GSI_META_SINGULAR))
        e->state = GSI_MEMO_SKIP;
      else
        {
          e->node = node;
          e->slot = lr.lnk[0];
          e->state = GSI_MEMO_FULL;
        }
    }
}

/* rare: the deferred list is full -- mark the query in the overflow bitmask; k_exact_overflow
 * resolves the marked queries after the main kernel (no exact replay inside the hot loops) */
__device__ __forceinline__ void
mark_overflow(unsigned *mask, size_t q, unsigned *counters)
{
  atomicOr(mask + (q >> 5), 1u << (unsigned) (q & 31));
  atomicAdd(counters + GSI_CNT_OVERFLOW, 1u);
}

/* ---- fast path, one thread per query (coherent query sets: rasters, sorted queries) --------------
 * Two phases per batch of 32 queries so that the expensive part stays converged: (1) the walk
 * (lanes may take different numbers of steps), (2) tail fetch + blend + stores for all accepted
 * lanes together -- consecutive lanes are consecutive queries, so every store instruction
 * writes full 32-byte sectors of one output plane. */
template <class SRC, bool STATS>
__global__ void __launch_bounds__(GSI_LOCATE_THREADS, 4)
k_locate(const gsi_view V, const SRC src, size_t nq, double *val, int *leaf_out, size_t ld,
         unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters, int max_steps)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long steps_acc = 0;
  /* warp-uniform trip count so the ballot below is always converged */
  const size_t n_launch = src.n_launch(nq);
  for (size_t base = (size_t) blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_launch;
       base += stride)
    {
      const size_t t = base + lane;
      size_t q = 0;
      int status = GSI_STEP_MOVE; /* = not a live query */
      int slot = 0;
      double qx = 0, qy = 0, c0 = 0, c1 = 0;
      if (t < n_launch && src.get(t, q, qx, qy))
        {
          status = GSI_STEP_DEFER;
          if (isfinite(qx) && isfinite(qy))
            {
              slot = gsi_seed_of(&V, qx, qy);
              for (int step = 0; step < max_steps; step++)
                {
                  gsi_rec r;
                  int at = slot;
                  gsi_fetch_rec(&r, gsi_leaf_at(&V, at));
                  const int st = gsi_walk_step(&V.xf, &r, qx, qy, &slot, &c0, &c1);
                  if (st != GSI_STEP_MOVE)
                    {
                      status = st;
                      slot = at;
                      if (STATS && st == GSI_STEP_ACCEPT) steps_acc += (unsigned) step + 1u;
                      break;
                    }
                }
            }
        }
      /* independent thread scheduling does not promise that the lanes which left the walk early
         wait here for the others: without this barrier phase 2 runs several times with partial
         lane masks (measured: 15 of 32 lanes active on average) */
      __syncwarp();
      if (status == GSI_STEP_ACCEPT)
        {
          gsi_leaf_tail t;
          gsi_fetch_tail(&t, gsi_leaf_at(&V, slot));
          if (leaf_out) leaf_out[q] = t.g.node;
          if (val) gsi_blend_store(&V, &t, c0, c1, val, ld, q);
        }
      const bool defer = status == GSI_STEP_DEFER;
      /* warp-aggregated append to the deferred list */
      const unsigned mask = __ballot_sync(0xffffffffu, defer);
      if (mask)
        {
          const int leader = __ffs(mask) - 1;
          unsigned pos = 0;
          if (lane == (unsigned) leader)
            pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(mask));
          pos = __shfl_sync(0xffffffffu, pos, leader);
          if (defer)
            {
              const unsigned at = pos + __popc(mask & ((1u << lane) - 1u));
              if (at < list_cap)
                list[at] = (unsigned) q;
              else
                mark_overflow(ovf_mask, q, counters);
            }
        }
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
}

/* ---- fast path, lanes decoupled from queries -------------------------------------------------
 * One warp owns a contiguous chunk of queries and consumes it through a warp-wide work queue:
 * every loop iteration is ONE walk step (one 96-byte leaf fetch = three 256-bit loads) for
 * whatever query each lane currently holds; a lane that accepted or deferred its query takes the
 * next one from the queue at once instead of idling until the slowest lane of the warp finished
 * its walk.  Each lane keeps two queries in flight behind the current one -- `nn` (coordinates
 * being loaded) and `nxt` (seed cell being loaded) -- so the only exposed latency per iteration
 * is the leaf fetch itself. */
template <class SRC, bool STATS, bool ONE_FIELD>
__global__ void __launch_bounds__(GSI_WALK_THREADS, 4)
k_walk(const gsi_view V, const SRC src, size_t nq, unsigned chunk, double *val, int *leaf_out,
       size_t ld, unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters,
       int max_steps)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t warp_id = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t) gridDim.x * blockDim.x) >> 5;
  const size_t n_chunks = (nq + chunk - 1) / chunk;
  unsigned long long steps_acc = 0;
  for (size_t ch = warp_id; ch < n_chunks; ch += n_warps)
    {
      const size_t q0 = ch * (size_t) chunk;
      const unsigned total = (unsigned) (nq - q0 < (size_t) chunk ? nq - q0 : (size_t) chunk);
      unsigned next = 0; /* warp-uniform: queries of this chunk handed out so far */
      bool cur_v = false, nxt_v = false, nn_v = false;
      unsigned cur_q = 0, nxt_q = 0, nn_q = 0;
      double cur_x = 0, cur_y = 0, nxt_x = 0, nxt_y = 0, nn_x = 0, nn_y = 0;
      int cur_slot = 0, nxt_seed = 0, cur_steps = 0;
      for (;;)
        {
          /* refill: promote nxt -> cur, nn -> nxt (issue its seed load), take a new nn */
          const bool need = !cur_v;
          if (need)
            {
              cur_v = nxt_v; cur_q = nxt_q; cur_x = nxt_x; cur_y = nxt_y;
              cur_slot = nxt_seed; cur_steps = 0;
              nxt_v = nn_v; nxt_q = nn_q; nxt_x = nn_x; nxt_y = nn_y;
              if (nxt_v) nxt_seed = gsi_seed_of(&V, nxt_x, nxt_y);
            }
          const unsigned take = __ballot_sync(FULL, need);
          if (need)
            {
              const unsigned off = next + __popc(take & lt_mask);
              nn_v = off < total;
              if (nn_v)
                {
                  nn_q = off;
                  src.get_linear(q0 + off, nn_x, nn_y);
                }
            }
          next += __popc(take);
          if (next > total) next = total;
          if (!__any_sync(FULL, cur_v || nxt_v || nn_v)) break;

          bool defer = false;
          if (cur_v)
            {
              if (!(isfinite(cur_x) && isfinite(cur_y)))
                defer = true;
              else
                {
                  const gsi_leaf *lf = gsi_leaf_at(&V, cur_slot);
                  gsi_rec r;
                  gsi_leaf_tail t;
                  gsi_fetch_rec(&r, lf);
                  gsi_fetch_tail(&t, lf); /* speculative: only used when this leaf is accepted */
                  double c0, c1;
                  const int st = gsi_walk_step(&V.xf, &r, cur_x, cur_y, &cur_slot, &c0, &c1);
                  ++cur_steps;
                  if (st == GSI_STEP_ACCEPT)
                    {
                      const size_t q = q0 + cur_q;
                      if (leaf_out) GSI_ST_RESULT(leaf_out + q, t.g.node);
                      if (val)
                        {
                          if (ONE_FIELD)
                            GSI_ST_RESULT(val + q, gsi_blend(c0, c1, (int) (t.s.vflags & 1u), t.s.z[0],
                                               (int) (t.s.vflags & 2u), t.s.z[1],
                                               (int) (t.s.vflags & 4u), t.s.z[2]));
                          else
                            gsi_blend_store(&V, &t, c0, c1, val, ld, q);
                        }
                      if (STATS) steps_acc += (unsigned) cur_steps;
                      cur_v = false;
                    }
                  else if (st == GSI_STEP_DEFER || cur_steps >= max_steps)
                    defer = true;
                }
              if (defer) cur_v = false;
            }
          /* warp-aggregated append to the deferred list */
          const unsigned dm = __ballot_sync(FULL, defer);
          if (dm)
            {
              const int leader = __ffs(dm) - 1;
              unsigned pos = 0;
              if (lane == (unsigned) leader)
                pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
              pos = __shfl_sync(FULL, pos, leader);
              if (defer)
                {
                  const size_t q = q0 + cur_q;
                  const unsigned at = pos + __popc(dm & lt_mask);
                  if (at < list_cap)
                    list[at] = (unsigned) q;
                  else
                    mark_overflow(ovf_mask, q, counters);
                }
            }
        }
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
}

/* ---- fast path, two walks per lane in rotation behind L2 prefetches ----------------------------
 * k_walk exposes one full DRAM latency per loop iteration: the leaf a lane needs next (its
 * seed, or the neighbour it just moved to) is only requested when the lane gets there.  Here
 * every lane owns TWO ready queries, P and Q, and alternates between them: one iteration is a
 * walk step of P; when it ends the lane knows which leaf P needs next (the neighbour it moved
 * to, or the seed of the query that replaces it) and asks the L2 for that line
 * (prefetch.global.L2 -- no destination register, nothing to wait for), then turns to Q, whose
 * line was requested one iteration ago.  The record loads themselves then hit the L2, and the
 * HBM fill of one query overlaps the arithmetic of the other.  Behind P and Q sit the two feeder
 * stages of k_walk (N: seed cell being loaded, M: coordinates being loaded).
 *
 * The queue is continuous: a warp draws chunks of `chunk` consecutive queries from a device
 * counter (the id of the chunk after the current one is always in flight) and moves on to the
 * next chunk in the middle of a refill, so there is no drain/fill bubble per chunk and a CTA that
 * starts late or runs on a slower SM simply takes fewer chunks.  `trace` (optional, zeroed by the
 * caller) receives {smid, start, end in globaltimer ns, queries handed out} per CTA. */
/* how the line(s) of a leaf record are requested ahead of time (experiment knob "walk_pf"):
 *   0 nothing; 1 prefetch.global.L2 on the first byte (+ the last sector when the record ends in the
 *   next line); 2 prefetch.global.L2 on each of the three sectors; 3 one TMA bulk prefetch of the 96
 *   bytes; 4 a 4-byte demand load per line (returned in *sink so that it stays in flight) */
__device__ __forceinline__ void
gsi_prefetch_leaf_l2(const gsi_view &V, int slot, int how, unsigned *sink)
{
  const char *p = reinterpret_cast<const char *>(gsi_leaf_at(&V, slot));
  const bool straddles = (((unsigned) (size_t) p) & 127u) > 128u - (unsigned) sizeof(gsi_leaf);
  if (how == 1)
    {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (straddles) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + sizeof(gsi_leaf) - 32));
    }
  else if (how == 2)
    {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64));
    }
  else if (how == 3)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((unsigned) sizeof(gsi_leaf)) : "memory");
  else if (how == 4)
    {
      unsigned a, b = 0;
      asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
      if (straddles) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(b) : "l"(p + sizeof(gsi_leaf) - 32));
      *sink = a ^ b;
    }
}

template <class SRC, bool STATS, bool ONE_FIELD>
__global__ void __launch_bounds__(GSI_WALK_THREADS, 5)
k_walkp(const gsi_view V, const SRC src, size_t nq, unsigned chunk, double *val, int *leaf_out,
        size_t ld, unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters,
        int max_steps, int dynamic, int pf, unsigned long long *trace)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned n_chunks = (unsigned) ((nq + chunk - 1) / chunk);
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  unsigned *work = counters + GSI_CNT_WORK;
  unsigned long long steps_acc = 0, t_start = 0;
  unsigned handed = 0;
  if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

  /* warp-uniform queue state: queries [base + next, base + total) of the current chunk are still
     to be handed out; lane 0 keeps the id of the following chunk in `ahead` (atomic in flight) */
  unsigned base = 0, next = 0, total = 0, ahead = 0;
  bool dry = false; /* the counter ran past the last chunk */
  /* dynamic = 0 deals the chunks round-robin instead (warp w: w, w + n_warps, ...) */
  if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

  bool p_v = false, q_v = false, n_v = false, m_v = false;
  unsigned p_q = 0, q_q = 0, n_q = 0, m_q = 0;
  double p_x = 0, p_y = 0, q_x = 0, q_y = 0, n_x = 0, n_y = 0, m_x = 0, m_y = 0;
  int p_slot = 0, q_slot = 0, n_seed = 0, p_steps = 0, q_steps = 0;
  unsigned p_sink = 0, q_sink = 0, sink_acc = 0;
  for (;;)
    {
      /* ---- one walk step of P (its leaf line was requested an iteration ago) */
      bool defer = false;
      if (pf == 4) sink_acc ^= p_sink;
      if (p_v)
        {
          if (!(isfinite(p_x) && isfinite(p_y)))
            defer = true;
          else
            {
              const gsi_leaf *lf = gsi_leaf_at(&V, p_slot);
              gsi_rec r;
              gsi_leaf_tail t;
              gsi_fetch_rec(&r, lf);
              gsi_fetch_tail(&t, lf); /* speculative: only used when this leaf is accepted */
              double c0, c1;
              const int st = gsi_walk_step(&V.xf, &r, p_x, p_y, &p_slot, &c0, &c1);
              ++p_steps;
              if (st == GSI_STEP_ACCEPT)
                {
                  const size_t q = p_q;
                  if (leaf_out) GSI_ST_RESULT(leaf_out + q, t.g.node);
                  if (val)
                    {
                      if (ONE_FIELD)
                        GSI_ST_RESULT(val + q, gsi_blend(c0, c1, (int) (t.s.vflags & 1u), t.s.z[0],
                                           (int) (t.s.vflags & 2u), t.s.z[1],
                                           (int) (t.s.vflags & 4u), t.s.z[2]));
                      else
                        gsi_blend_store(&V, &t, c0, c1, val, ld, q);
                    }
                  if (STATS) steps_acc += (unsigned) p_steps;
                  p_v = false;
                }
              else if (st == GSI_STEP_DEFER || p_steps >= max_steps)
                defer = true;
            }
          if (defer) p_v = false;
        }
      /* warp-aggregated append to the deferred list */
      const unsigned dm = __ballot_sync(FULL, defer);
      if (dm)
        {
          const int leader = __ffs(dm) - 1;
          unsigned pos = 0;
          if (lane == (unsigned) leader)
            pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
          pos = __shfl_sync(FULL, pos, leader);
          if (defer)
            {
              const unsigned at = pos + __popc(dm & lt_mask);
              if (at < list_cap)
                list[at] = p_q;
              else
                mark_overflow(ovf_mask, (size_t) p_q, counters);
            }
        }

      /* ---- refill: P <- N (seed known by now), N <- M (issue its seed load), M <- queue */
      const bool need = !p_v;
      if (need)
        {
          p_v = n_v; p_q = n_q; p_x = n_x; p_y = n_y; p_slot = n_seed; p_steps = 0;
          n_v = m_v; n_q = m_q; n_x = m_x; n_y = m_y;
          if (n_v && isfinite(n_x) && isfinite(n_y)) n_seed = gsi_seed_of(&V, n_x, n_y);
          m_v = false;
        }
      const unsigned take = __ballot_sync(FULL, need);
      if (take && !dry)
        {
          const unsigned cnt = (unsigned) __popc(take), rank = (unsigned) __popc(take & lt_mask);
          unsigned rem = total - next;
          if (cnt > rem)
            {
              /* the lanes ranked below `rem` finish the current chunk, the others start the next */
              const unsigned ch = __shfl_sync(FULL, ahead, 0);
              if (need && rank < rem)
                {
                  m_v = true;
                  m_q = base + next + rank;
                }
              if (ch < n_chunks)
                {
                  if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : ch + n_warps;
                  base = ch * chunk;
                  total = (unsigned) (nq - (size_t) base < (size_t) chunk ? nq - (size_t) base : (size_t) chunk);
                  const unsigned want = cnt - rem;
                  if (need && rank >= rem && rank - rem < total)
                    {
                      m_v = true;
                      m_q = base + (rank - rem);
                    }
                  next = want < total ? want : total;
                }
              else
                {
                  dry = true;
                  next = total;
                }
            }
          else
            {
              if (need)
                {
                  m_v = true;
                  m_q = base + next + rank;
                }
              next += cnt;
            }
          if (need && m_v) src.get_linear((size_t) m_q, m_x, m_y);
          if (trace) handed += (unsigned) __popc(__ballot_sync(FULL, need && m_v));
        }
      /* ask the L2 for the line P needs two iterations from now (moved: the neighbour; refilled:
         the seed leaf) */
      if (pf && p_v && isfinite(p_x) && isfinite(p_y)) gsi_prefetch_leaf_l2(V, p_slot, pf, &p_sink);
      if (!__any_sync(FULL, p_v || q_v || n_v || m_v)) break;

      /* ---- rotate: Q is processed next */
      {
        const bool tv = p_v; p_v = q_v; q_v = tv;
        const unsigned tq = p_q; p_q = q_q; q_q = tq;
        double td = p_x; p_x = q_x; q_x = td;
        td = p_y; p_y = q_y; q_y = td;
        int ti = p_slot; p_slot = q_slot; q_slot = ti;
        ti = p_steps; p_steps = q_steps; q_steps = ti;
        const unsigned ts = p_sink; p_sink = q_sink; q_sink = ts;
      }
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
  if (pf == 4 && sink_acc == 0x9e3779b9u && nq == 0) counters[GSI_CNT_WORDS - 1] = sink_acc; /* keeps the loads alive */
  if (trace)
    {
      __syncthreads();
      if (threadIdx.x == 0)
        {
          unsigned long long t_end;
          unsigned smid;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          trace[4 * (size_t) blockIdx.x + 0] = smid;
          trace[4 * (size_t) blockIdx.x + 1] = t_start;
          trace[4 * (size_t) blockIdx.x + 2] = t_end;
        }
      if (lane == 0) atomicAdd(trace + 4 * (size_t) blockIdx.x + 3, (unsigned long long) handed);
    }
}

/* ---- fast path, warp-cooperative staging of the visited leaves (walk_variant = 2) ----------------
 * The walk kernels above gather a leaf record with three fully divergent 256-bit loads per lane:
 * 96 L1 tag look-ups per warp and iteration for 32 records, and that tag rate (about 0.6 sectors per
 * clock and SM, whatever the hit rate) is what bounds them.  Here the leaves are 128-byte records on
 * 128-byte lines (slab with leaf_stride = 128) and the WARP fetches them: eight lanes copy one record
 * as eight 16-byte pieces of one line (cp.async, global -> shared memory without registers), four
 * records per instruction, so a visit costs ONE tag look-up instead of three.  Every lane then reads
 * its own record from shared memory (112-byte pitch: conflict-free 128-bit reads).  The spare 32 bytes
 * of the line carry the reciprocals of the record's two divisors (k_leaf_rcp), so the two binary64
 * divisions of the solve cost three operations each (gsi_walk_step_rcp; a quotient the hoisted form
 * cannot vouch for defers the query).  With the record in shared memory the kernel fits 64
 * registers: 8 CTAs = 32 warps per SM.  Queue as in k_walkp (continuous, chunks from a counter). */
#ifndef GSI_WALKC_THREADS
#define GSI_WALKC_THREADS 128 /* CTA size of k_walkc; 1024 / GSI_WALKC_THREADS CTAs per SM */
#endif
#ifndef GSI_WALKC_SBASE
#define GSI_WALKC_SBASE (GSI_SEED_GB_MAX * GSI_SEED_GB_MAX)
#endif
#define GSI_WALKC_PITCH 112 /* 7 x 16 bytes; 28 words: conflict-free 128-bit reads, 8 CTAs in 164 KB of shared memory */
#ifndef GSI_WALKC_CP
#define GSI_WALKC_CP "cg" /* L2 only: the staged lines are not read through L1 again */
#endif

struct gsi_warp_queue
{
  unsigned base = 0, next = 0, total = 0, ahead = 0; /* ahead: lane 0 only */
  bool dry = false;
  unsigned chunk, n_chunks, n_warps;
  size_t nq;
  unsigned *work;
  int dynamic;

  __device__ __forceinline__ void init(size_t nq_, unsigned chunk_, unsigned *work_, int dynamic_, unsigned lane)
  {
    nq = nq_; chunk = chunk_; work = work_; dynamic = dynamic_;
    n_chunks = (unsigned) ((nq + chunk - 1) / chunk);
    n_warps = (gridDim.x * blockDim.x) >> 5;
    if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  }
  /* warp-wide: every lane with `need` gets the next query of the stream in *q; false = none left for it */
  __device__ __forceinline__ bool take(bool need, unsigned lane, unsigned lt_mask, unsigned *q)
  {
    const unsigned FULL = 0xffffffffu;
    const unsigned mask = __ballot_sync(FULL, need);
    bool got = false;
    if (!mask || dry) return false;
    const unsigned cnt = (unsigned) __popc(mask), rank = (unsigned) __popc(mask & lt_mask);
    const unsigned rem = total - next;
    if (cnt > rem)
      {
        const unsigned ch = __shfl_sync(FULL, ahead, 0);
        if (need && rank < rem)
          {
            got = true;
            *q = base + next + rank;
          }
        if (ch < n_chunks)
          {
            if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : ch + n_warps;
            base = ch * chunk;
            total = (unsigned) (nq - (size_t) base < (size_t) chunk ? nq - (size_t) base : (size_t) chunk);
            const unsigned want = cnt - rem;
            if (need && rank >= rem && rank - rem < total)
              {
                got = true;
                *q = base + (rank - rem);
              }
            next = want < total ? want : total;
          }
        else
          {
            dry = true;
            next = total;
          }
      }
    else
      {
        if (need)
          {
            got = true;
            *q = base + next + rank;
          }
        next += cnt;
      }
    return got;
  }
};

/* gsi_seed_of with the cell computed by one fused multiply-add per axis and an integer clamp (the cell
 * is a hint: any leaf slot is a correct start of the walk; NaN converts to 0, +-inf saturate) */
struct gsi_cell_map
{
  double ax, bx, ay, by;
};

/* issue the load of a query's 16-bit seed offset (nothing waits for it here) and note its seed block */
__device__ __forceinline__ void
gsi_seed_issue(const gsi_view &V, const gsi_cell_map &C, double qx, double qy, unsigned *off, int *blk)
{
  const int ix = min(max(__double2int_rd(__fma_rn(qx, C.ax, C.bx)), 0), V.grid_g - 1);
  const int iy = min(max(__double2int_rd(__fma_rn(qy, C.ay, C.by)), 0), V.grid_g - 1);
  const unsigned cell = (unsigned) iy * (unsigned) V.grid_g + (unsigned) ix;
  *off = gsi_ld_seed16(V.seed16 + cell);
  *blk = (int) ((unsigned) (iy >> V.seed_shift) * (unsigned) V.grid_gb + (unsigned) (ix >> V.seed_shift));
}

/* ... and turn it into the seed slot an iteration later: block base from the copy of the base table in
   shared memory (escaped cells, if any: the int32 table) */
__device__ __forceinline__ int
gsi_seed_resolve(const gsi_view &V, const gsi_cell_map &C, const int *sbase, double qx, double qy, unsigned off, int blk)
{
  if (blk < 0) return -1; /* non-finite query: no seed, the caller defers it */
  if (off != GSI_SEED16_ESC) return sbase[blk] + (int) off;
  const int ix = min(max(__double2int_rd(__fma_rn(qx, C.ax, C.bx)), 0), V.grid_g - 1);
  const int iy = min(max(__double2int_rd(__fma_rn(qy, C.ay, C.by)), 0), V.grid_g - 1);
  return gsi_ld_seed(V.seeds + ((unsigned) iy * (unsigned) V.grid_g + (unsigned) ix));
}

__device__ __forceinline__ void
gsi_lds128(unsigned addr, unsigned long long &a, unsigned long long &b)
{
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

template <class SRC, bool STATS, bool ONE_FIELD>
__global__ void __launch_bounds__(GSI_WALKC_THREADS, 1024 / GSI_WALKC_THREADS)
k_walkc(const gsi_view V, const gsi_cell_map C, const SRC src, size_t nq, unsigned chunk, double *val, int *leaf_out,
        size_t ld, unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters,
        int max_steps, int dynamic, int l1_alloc, unsigned long long *trace)
{
  __shared__ __align__(16) unsigned char stage_all[(GSI_WALKC_THREADS / 32) * 32 * GSI_WALKC_PITCH];
  __shared__ int sbase[GSI_WALKC_SBASE]; /* the seed-block bases: at most 4 KB */
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned stage = (unsigned) __cvta_generic_to_shared(stage_all) + (threadIdx.x >> 5) * (32u * GSI_WALKC_PITCH);
  {
    const unsigned nb = (unsigned) V.grid_gb * (unsigned) V.grid_gb;
    for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) sbase[i] = __ldg(V.seedbase + i);
  }
  __syncthreads();
  unsigned long long steps_acc = 0, t_start = 0;
  unsigned handed = 0;
  if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  gsi_warp_queue Q;
  Q.init(nq, chunk, counters + GSI_CNT_WORK, dynamic, lane);

  bool cur_v = false, nxt_v = false, nn_v = false;
  unsigned cur_q = 0, nxt_q = 0, nn_q = 0;
  double cur_x = 0, cur_y = 0, nxt_x = 0, nxt_y = 0, nn_x = 0, nn_y = 0;
  int cur_slot = 0, nxt_base = 0, cur_steps = 0;
  unsigned nxt_off = 0;
  for (;;)
    {
      /* refill: promote nxt -> cur (its seed loads were issued a refill ago and are only consumed
         now), nn -> nxt (issue its seed loads), take a new nn */
      const bool need = !cur_v;
      if (need)
        {
          cur_v = nxt_v; cur_q = nxt_q; cur_x = nxt_x; cur_y = nxt_y;
          cur_steps = 0;
          if (cur_v) cur_slot = gsi_seed_resolve(V, C, sbase, cur_x, cur_y, nxt_off, nxt_base);
          nxt_v = nn_v; nxt_q = nn_q; nxt_x = nn_x; nxt_y = nn_y;
          /* a non-finite query gets block -1, hence slot -1: nothing is staged for it and it is deferred below */
          nxt_off = 0;
          nxt_base = -1;
          if (nxt_v && isfinite(nxt_x) && isfinite(nxt_y)) gsi_seed_issue(V, C, nxt_x, nxt_y, &nxt_off, &nxt_base);
          nn_v = false;
        }
      {
        unsigned q = 0;
        const bool got = Q.take(need, lane, lt_mask, &q);
        if (got)
          {
            nn_v = true;
            nn_q = q;
            src.get_linear((size_t) q, nn_x, nn_y);
          }
        if (trace) handed += (unsigned) __popc(__ballot_sync(FULL, got));
      }
      if (!__any_sync(FULL, cur_v || nxt_v || nn_v)) break;

      /* the warp copies the records its lanes are about to visit: lanes 8j .. 8j+7 move record j of
         each group of four as 16-byte pieces of one line (the last piece is padding) */
      const bool live = cur_v && cur_slot >= 0;
      const int my = live ? cur_slot : -1;
      __syncwarp(); /* last iteration's reads of the staging area are done */
#pragma unroll
      for (int r = 0; r < 8; r++)
        {
          const unsigned qq = 4u * (unsigned) r + (lane >> 3);
          const int sl = __shfl_sync(FULL, my, (int) qq);
          if (sl >= 0 && (lane & 7u) != 7u)
            {
              const char *g = V.leaf + (size_t) sl * 128u + (lane & 7u) * 16u;
              const unsigned d = stage + qq * GSI_WALKC_PITCH + (lane & 7u) * 16u;
              /* .cg: the line goes to shared memory only (best when every visit is a miss anyway);
                 .ca: it also allocates in L1, which takes the pressure off the L2 slice that owns a
                 HOT leaf (cage leaves, the big leaves of sparse regions): launch_eval picks */
              if (l1_alloc)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
              else
                asm volatile("cp.async." GSI_WALKC_CP ".shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();

      bool defer = false;
      if (cur_v)
        {
          if (!live)
            defer = true;
          else
            {
              const unsigned mine = stage + lane * GSI_WALKC_PITCH;
              unsigned long long a, b;
              gsi_rec r;
              gsi_lds128(mine, a, b);
              r.x0x = __longlong_as_double((long long) a); r.x0y = __longlong_as_double((long long) b);
              gsi_lds128(mine + 16, a, b);
              r.u00 = __longlong_as_double((long long) a); r.u01 = __longlong_as_double((long long) b);
              gsi_lds128(mine + 32, a, b);
              r.u11 = __longlong_as_double((long long) a); r.l10 = __longlong_as_double((long long) b);
              gsi_lds128(mine + 48, a, b);
              r.lnk[0] = (int32_t) (unsigned) a; r.lnk[1] = (int32_t) (unsigned) (a >> 32);
              r.lnk[2] = (int32_t) (unsigned) b; r.meta = (uint32_t) (b >> 32);
              gsi_lds128(mine + 96, a, b);
              const double r11 = __longlong_as_double((long long) a), r00 = __longlong_as_double((long long) b);
              double c0, c1;
              const int st = gsi_walk_step_rcp(&V.xf, &r, r11, r00, cur_x, cur_y, &cur_slot, &c0, &c1);
              ++cur_steps;
              if (st == GSI_STEP_ACCEPT)
                {
                  gsi_leaf_tail t;
                  gsi_lds128(mine + 64, a, b);
                  t.s.node = (int32_t) (unsigned) a; t.s.vflags = (uint32_t) (a >> 32);
                  t.s.z[0] = __longlong_as_double((long long) b);
                  gsi_lds128(mine + 80, a, b);
                  t.s.z[1] = __longlong_as_double((long long) a); t.s.z[2] = __longlong_as_double((long long) b);
                  const size_t q = cur_q;
                  if (leaf_out) GSI_ST_RESULT(leaf_out + q, t.g.node);
                  if (val)
                    {
                      if (ONE_FIELD)
                        GSI_ST_RESULT(val + q, gsi_blend(c0, c1, (int) (t.s.vflags & 1u), t.s.z[0],
                                           (int) (t.s.vflags & 2u), t.s.z[1],
                                           (int) (t.s.vflags & 4u), t.s.z[2]));
                      else
                        gsi_blend_store(&V, &t, c0, c1, val, ld, q);
                    }
                  if (STATS) steps_acc += (unsigned) cur_steps;
                  cur_v = false;
                }
              else if (st == GSI_STEP_DEFER || cur_steps >= max_steps)
                defer = true;
            }
          if (defer) cur_v = false;
        }
      /* warp-aggregated append to the deferred list */
      const unsigned dm = __ballot_sync(FULL, defer);
      if (dm)
        {
          const int leader = __ffs(dm) - 1;
          unsigned pos = 0;
          if (lane == (unsigned) leader)
            pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
          pos = __shfl_sync(FULL, pos, leader);
          if (defer)
            {
              const unsigned at = pos + __popc(dm & lt_mask);
              if (at < list_cap)
                list[at] = cur_q;
              else
                mark_overflow(ovf_mask, (size_t) cur_q, counters);
            }
        }
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
  if (trace)
    {
      __syncthreads();
      if (threadIdx.x == 0)
        {
          unsigned long long t_end;
          unsigned smid;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
          asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
          trace[4 * (size_t) blockIdx.x + 0] = smid;
          trace[4 * (size_t) blockIdx.x + 1] = t_start;
          trace[4 * (size_t) blockIdx.x + 2] = t_end;
        }
      if (lane == 0) atomicAdd(trace + 4 * (size_t) blockIdx.x + 3, (unsigned long long) handed);
    }
}

/* ---- fast path, fresh batches + a ring of walkers (walk_variant = 3) ------------------------------
 * k_walkc decouples lanes from queries with a warp-wide queue and pays for it on EVERY visit: about
 * 15 warp instructions per query, two divergent seed-table gathers and scattered result stores, and it
 * runs at the L1 wavefront rate (one divergent line = about two cycles of the SM's L1 pipe: 7.5
 * cycles per query and SM whether the structure is L2-resident or not).  Three out of four queries
 * are accepted by the first leaf they visit, so this kernel gives that visit a straight path:
 *   - a warp takes 32 CONSECUTIVE queries (a batch): coalesced coordinate loads two batches ahead,
 *     the 16-bit seed offsets one batch ahead, the block bases from a <= 4 KB table in SHARED memory
 *     (one divergent global access per seed instead of two), the 32 seed leaves staged by the warp
 *     (cp.async, as in k_walkc), one walk step per lane, results of the accepted lanes written to
 *     consecutive addresses;
 *   - a lane whose query has to move on pushes {x, y, query, next slot, steps} onto the warp's ring
 *     in shared memory; whenever the ring holds 32 walkers (or no batch is left) the warp pops 32 of
 *     them and gives them one visit with the same code.  Lanes stay full, no per-lane queue state.
 * Deferred queries go to the same list as in the other kernels.  Needs 128-byte leaf records. */
#ifndef GSI_WALKF_PITCH
#define GSI_WALKF_PITCH 112 /* 7 x 16 bytes of a record; 28 words: conflict-free 128-bit reads */
#endif
#define GSI_WALKF_RING 64
#define GSI_WALKF_WARP_BYTES (32 * GSI_WALKF_PITCH + GSI_WALKF_RING * 24 + GSI_WALKF_RING + 64)
#define GSI_WALKF_SBASE_BYTES (GSI_SEED_GB_MAX * GSI_SEED_GB_MAX * 4)
#define GSI_WALKF_SMEM (GSI_WALKF_SBASE_BYTES + (GSI_WALK_THREADS / 32) * GSI_WALKF_WARP_BYTES)

template <class SRC, bool STATS, bool ONE_FIELD>
__global__ void __launch_bounds__(GSI_WALK_THREADS, 8)
k_walkf(const gsi_view V, const gsi_cell_map C, const SRC src, size_t nq, unsigned chunk, double *val, int *leaf_out,
        size_t ld, unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters,
        int max_steps, int dynamic)
{
  extern __shared__ __align__(16) unsigned char walkf_smem[];
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  int *sbase = reinterpret_cast<int *>(walkf_smem);
  unsigned char *wmem = walkf_smem + GSI_WALKF_SBASE_BYTES + (threadIdx.x >> 5) * GSI_WALKF_WARP_BYTES;
  const unsigned stage = (unsigned) __cvta_generic_to_shared(wmem);
  double *rx = reinterpret_cast<double *>(wmem + 32 * GSI_WALKF_PITCH);
  double *ry = rx + GSI_WALKF_RING;
  unsigned *rq = reinterpret_cast<unsigned *>(ry + GSI_WALKF_RING);
  int *rs = reinterpret_cast<int *>(rq + GSI_WALKF_RING);
  unsigned char *rn = reinterpret_cast<unsigned char *>(rs + GSI_WALKF_RING);
  {
    const unsigned nb = (unsigned) V.grid_gb * (unsigned) V.grid_gb;
    for (unsigned i = threadIdx.x; i < nb; i += blockDim.x) sbase[i] = __ldg(V.seedbase + i);
  }
  __syncthreads();
  const int step_cap = max_steps < 255 ? max_steps : 255;
  unsigned long long steps_acc = 0;

  /* chunks of consecutive queries, dealt by a device counter; batches of 32 within a chunk */
  unsigned *work = counters + GSI_CNT_WORK;
  const unsigned n_chunks = (unsigned) ((nq + chunk - 1) / chunk);
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  unsigned ahead = 0, cbase = 0, cleft = 0;
  bool dry = false;
  if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;

  /* stage A: coordinates loaded; stage B: seed offset requested */
  unsigned a_n = 0, a_q0 = 0, b_n = 0, b_q0 = 0, b_off = 0, b_blk = 0;
  double a_x = 0, a_y = 0, b_x = 0, b_y = 0;
  unsigned head = 0, tail = 0; /* ring positions (free-running) */

  /* advance the batch pipeline: returns stage B as the current batch, B <- A, A <- a new batch */
#define GSI_WALKF_ADVANCE()                                                                        \
  do                                                                                               \
    {                                                                                              \
      b_n = a_n; b_q0 = a_q0; b_x = a_x; b_y = a_y;                                                \
      b_off = GSI_SEED16_ESC; b_blk = 0xffffffffu;                                                 \
      if (lane < b_n && isfinite(b_x) && isfinite(b_y))                                            \
        {                                                                                          \
          const int ix = min(max(__double2int_rd(__fma_rn(b_x, C.ax, C.bx)), 0), V.grid_g - 1);    \
          const int iy = min(max(__double2int_rd(__fma_rn(b_y, C.ay, C.by)), 0), V.grid_g - 1);    \
          b_off = gsi_ld_seed16(V.seed16 + ((unsigned) iy * (unsigned) V.grid_g + (unsigned) ix)); \
          b_blk = (unsigned) (iy >> V.seed_shift) * (unsigned) V.grid_gb + (unsigned) (ix >> V.seed_shift); \
        }                                                                                          \
      a_n = 0;                                                                                     \
      if (cleft == 0 && !dry)                                                                      \
        {                                                                                          \
          const unsigned ch = __shfl_sync(FULL, ahead, 0);                                         \
          if (ch >= n_chunks)                                                                      \
            dry = true;                                                                            \
          else                                                                                     \
            {                                                                                      \
              if (lane == 0) ahead = dynamic ? atomicAdd(work, 1u) : ch + n_warps;                 \
              cbase = ch * chunk;                                                                  \
              cleft = (unsigned) (nq - (size_t) cbase < (size_t) chunk ? nq - (size_t) cbase : (size_t) chunk); \
            }                                                                                      \
        }                                                                                          \
      if (cleft)                                                                                   \
        {                                                                                          \
          a_n = cleft < 32u ? cleft : 32u;                                                         \
          a_q0 = cbase;                                                                            \
          cbase += a_n; cleft -= a_n;                                                              \
          if (lane < a_n) src.get_linear((size_t) a_q0 + lane, a_x, a_y);                          \
        }                                                                                          \
    }                                                                                              \
  while (0)

  GSI_WALKF_ADVANCE(); /* A filled */
  GSI_WALKF_ADVANCE(); /* B filled, A refilled */

  for (;;)
    {
      __syncwarp(); /* ring entries pushed in the last iteration are visible */
      const unsigned cnt = tail - head;
      const bool have_fresh = (a_n | b_n) != 0;
      bool v = false;
      unsigned q = 0;
      double x = 0, y = 0;
      int slot = -1, steps = 0;
      if (cnt >= 32u || (!have_fresh && cnt))
        {
          const unsigned n = cnt < 32u ? cnt : 32u;
          v = lane < n;
          const unsigned idx = (head + lane) & (GSI_WALKF_RING - 1);
          if (v)
            {
              x = rx[idx]; y = ry[idx]; q = rq[idx]; slot = rs[idx]; steps = rn[idx];
            }
          head += n;
        }
      else if (have_fresh)
        {
          v = lane < b_n;
          q = b_q0 + lane;
          x = b_x; y = b_y;
          if (v && b_blk != 0xffffffffu)
            {
              if (b_off != GSI_SEED16_ESC)
                slot = sbase[b_blk] + (int) b_off;
              else
                {
                  const int ix = min(max(__double2int_rd(__fma_rn(x, C.ax, C.bx)), 0), V.grid_g - 1);
                  const int iy = min(max(__double2int_rd(__fma_rn(y, C.ay, C.by)), 0), V.grid_g - 1);
                  slot = gsi_ld_seed(V.seeds + ((unsigned) iy * (unsigned) V.grid_g + (unsigned) ix));
                }
            }
          GSI_WALKF_ADVANCE();
        }
      else
        break;

      /* the warp copies the 32 records its lanes visit (k_walkc's scheme, 112-byte pitch) */
      const bool live = v && slot >= 0;
      const int my = live ? slot : -1;
      __syncwarp(); /* last iteration's reads of the staging area are done */
#pragma unroll
      for (int r = 0; r < 8; r++)
        {
          const unsigned qq = 4u * (unsigned) r + (lane >> 3);
          const int sl = __shfl_sync(FULL, my, (int) qq);
          if (sl >= 0 && (lane & 7u) != 7u)
            {
              const char *g = V.leaf + (size_t) sl * 128u + (lane & 7u) * 16u;
              const unsigned d = stage + qq * GSI_WALKF_PITCH + (lane & 7u) * 16u;
              asm volatile("cp.async." GSI_WALKC_CP ".shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();

      bool defer = false, moved = false;
      if (v)
        {
          if (!live)
            defer = true; /* non-finite coordinates */
          else
            {
              const unsigned mine = stage + lane * GSI_WALKF_PITCH;
              unsigned long long a, b;
              gsi_rec r;
              gsi_lds128(mine, a, b);
              r.x0x = __longlong_as_double((long long) a); r.x0y = __longlong_as_double((long long) b);
              gsi_lds128(mine + 16, a, b);
              r.u00 = __longlong_as_double((long long) a); r.u01 = __longlong_as_double((long long) b);
              gsi_lds128(mine + 32, a, b);
              r.u11 = __longlong_as_double((long long) a); r.l10 = __longlong_as_double((long long) b);
              gsi_lds128(mine + 48, a, b);
              r.lnk[0] = (int32_t) (unsigned) a; r.lnk[1] = (int32_t) (unsigned) (a >> 32);
              r.lnk[2] = (int32_t) (unsigned) b; r.meta = (uint32_t) (b >> 32);
              gsi_lds128(mine + 96, a, b);
              const double r11 = __longlong_as_double((long long) a), r00 = __longlong_as_double((long long) b);
              double c0, c1;
              const int st = gsi_walk_step_rcp(&V.xf, &r, r11, r00, x, y, &slot, &c0, &c1);
              ++steps;
              if (st == GSI_STEP_ACCEPT)
                {
                  gsi_leaf_tail t;
                  gsi_lds128(mine + 64, a, b);
                  t.s.node = (int32_t) (unsigned) a; t.s.vflags = (uint32_t) (a >> 32);
                  t.s.z[0] = __longlong_as_double((long long) b);
                  gsi_lds128(mine + 80, a, b);
                  t.s.z[1] = __longlong_as_double((long long) a); t.s.z[2] = __longlong_as_double((long long) b);
                  const size_t qi = q;
                  if (leaf_out) GSI_ST_RESULT(leaf_out + qi, t.g.node);
                  if (val)
                    {
                      if (ONE_FIELD)
                        GSI_ST_RESULT(val + qi, gsi_blend(c0, c1, (int) (t.s.vflags & 1u), t.s.z[0],
                                                          (int) (t.s.vflags & 2u), t.s.z[1],
                                                          (int) (t.s.vflags & 4u), t.s.z[2]));
                      else
                        gsi_blend_store(&V, &t, c0, c1, val, ld, qi);
                    }
                  if (STATS) steps_acc += (unsigned) steps;
                }
              else if (st == GSI_STEP_DEFER || steps >= step_cap)
                defer = true;
              else
                moved = true;
            }
        }
      /* walkers go (back) onto the ring: at most 31 were left, at most 32 arrive */
      const unsigned mm = __ballot_sync(FULL, moved);
      if (mm)
        {
          if (moved)
            {
              const unsigned idx = (tail + (unsigned) __popc(mm & lt_mask)) & (GSI_WALKF_RING - 1);
              rx[idx] = x; ry[idx] = y; rq[idx] = q; rs[idx] = slot; rn[idx] = (unsigned char) steps;
            }
          tail += (unsigned) __popc(mm);
        }
      /* warp-aggregated append to the deferred list */
      const unsigned dm = __ballot_sync(FULL, defer);
      if (dm)
        {
          const int leader = __ffs(dm) - 1;
          unsigned pos = 0;
          if (lane == (unsigned) leader)
            pos = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
          pos = __shfl_sync(FULL, pos, leader);
          if (defer)
            {
              const unsigned at = pos + __popc(dm & lt_mask);
              if (at < list_cap)
                list[at] = q;
              else
                mark_overflow(ovf_mask, (size_t) q, counters);
            }
        }
    }
#undef GSI_WALKF_ADVANCE
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
}

/* reciprocals of the two divisors of every 128-byte leaf record, into the record's padding */
__global__ void __launch_bounds__(256)
k_leaf_rcp(const gsi_view V, char *leaf_base)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t s = (size_t) blockIdx.x * blockDim.x + threadIdx.x; s < (size_t) V.n_leaves; s += stride)
    {
      const gsi_rec *r = reinterpret_cast<const gsi_rec *>(leaf_base + s * 128u);
      double *pad = reinterpret_cast<double *>(leaf_base + s * 128u + 96u);
      const double u11 = r->u11, u00 = r->u00;
      pad[0] = gsi_div_rcp_usable(u11) ? gsi_div_rcp(u11) : CUDART_NAN;
      pad[1] = gsi_div_rcp_usable(u00) ? gsi_div_rcp(u00) : CUDART_NAN;
    }
}

/* ---- reorder pipeline: k_bin -> k_tile_walk -> k_unpermute -----------------------------------
 * Uniformly random queries make every leaf fetch an isolated L2/DRAM line fill: at 10^6 nodes the
 * direct walk moves ~220 bytes per query through HBM for 28 algorithmic bytes.  The pipeline
 * trades those fills for three streaming passes (~90 bytes per query) and keeps the structure
 * on chip:
 *   k_bin          one CTA per chunk of <= 8192 consecutive queries: counting sort by spatial tile
 *                  (the leaf order's own Morton-key prefix, so a tile's leaves are one contiguous
 *                  slot range) entirely in shared memory, then coalesced stores of the binned
 *                  coordinates, the 16-bit in-chunk index of every binned query and the chunk's
 *                  runs (bin-major table: one contiguous row per tile);
 *   k_tile_walk    tile-major, persistent CTAs pulling (tile, range) work items from a counter:
 *                  the tile's leaf records arrive in shared memory by one TMA bulk copy
 *                  (cp.async.bulk + mbarrier), its seed cells beside them; every warp flattens the
 *                  tile's runs of 32 chunks at a time into a dense position list and consumes it
 *                  through a warp-wide work queue (one walk step per iteration for whatever query
 *                  each lane holds -- a lane that finished takes the next query at once), reading
 *                  seeds, records and tails from shared memory;
 *   k_unpermute    one CTA per chunk: the inverse permutation through shared memory, coalesced
 *                  loads and stores.
 * The arithmetic per query is the one of every other fast-path kernel (gsi_walk_step), so
 * accepted results are bit-identical; deferred queries go to the same list (original query index)
 * and are replayed exactly after k_unpermute. */
__device__ __forceinline__ unsigned
spread16(unsigned x) /* the low 16 bits of x on the even bit positions */
{
  x &= 0x0000ffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

/* Tiles are the 2^P prefixes of the 32-bit Morton key gsi_flatten.cpp sorts the leaves' centroids
 * by (16 bits per axis, y on the odd bits): the prefix holds ceil(P/2) bits of y and floor(P/2) of
 * x.  k_bin numbers its bins row-major, bin = (ty << bx_bits) | tx, from one fused multiply-add per
 * axis; the tile kernel turns a bin into the Morton prefix that indexes V.tiles.  Binning is a
 * heuristic only: a query binned next door still walks to its leaf through global memory.  NaN
 * converts to 0 and +-inf saturate, so every query gets a bin. */
struct gsi_tile_map
{
  double ax, bx, ay, by; /* tx = floor(qx*ax + bx), ty = floor(qy*ay + by) before clamping */
  int bx_bits, by_bits;
};

__device__ __forceinline__ unsigned
bin_of(const gsi_tile_map &M, double qx, double qy)
{
  const int tx = min(max(__double2int_rd(__fma_rn(qx, M.ax, M.bx)), 0), (1 << M.bx_bits) - 1);
  const int ty = min(max(__double2int_rd(__fma_rn(qy, M.ay, M.by)), 0), (1 << M.by_bits) - 1);
  return ((unsigned) ty << M.bx_bits) | (unsigned) tx;
}

__device__ __forceinline__ unsigned
bin_to_prefix(const gsi_tile_map &M, unsigned bin)
{
  const unsigned tx = bin & ((1u << M.bx_bits) - 1u), ty = bin >> M.bx_bits;
  /* odd P: the lowest prefix bit belongs to y */
  return (M.by_bits > M.bx_bits) ? (spread16(ty) | (spread16(tx) << 1)) : (spread16(tx) | (spread16(ty) << 1));
}

/* One coordinate is staged at a time (96 KB of shared memory for 8192-query chunks) so that two
 * CTAs share an SM and one CTA's global loads overlap the other's scan / copy-out. */
__global__ void __launch_bounds__(GSI_BIN_THREADS, 2)
k_bin(const gsi_tile_map M, const double *__restrict__ xq, const double *__restrict__ yq, size_t nq, int S,
      int nt, double *__restrict__ bx, double *__restrict__ by, unsigned short *__restrict__ bi,
      unsigned *__restrict__ runs)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sbuf = reinterpret_cast<double *>(smem_raw);
  unsigned *hist = reinterpret_cast<unsigned *>(sbuf + S);
  unsigned short *si = reinterpret_cast<unsigned short *>(hist + nt);
  constexpr unsigned T = GSI_BIN_THREADS, WPC = T / 32;
  constexpr int KPT = GSI_BIN_CHUNK_MAX / T;
  constexpr unsigned PER = 4096 / T; /* histogram entries per thread in the scan, nt <= 4096 */
  __shared__ unsigned warp_sum[WPC];
  const unsigned FULL = 0xffffffffu;
  const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const size_t n_chunks = (nq + (size_t) S - 1) / (size_t) S;
  const unsigned per = ((unsigned) nt + T - 1) / T;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x)
    {
      const size_t q0 = c * (size_t) S;
      const unsigned n = (unsigned) (nq - q0 < (size_t) S ? nq - q0 : (size_t) S);
      /* the chunk this CTA sorts next: ask the L2 for its coordinates now (TMA bulk prefetch), so that the
         loads of the next round hit the L2 instead of waiting for HBM between two barriers */
      if (tid < 2 && c + gridDim.x < n_chunks)
        {
          const size_t qn = (c + gridDim.x) * (size_t) S;
          const size_t nn = nq - qn < (size_t) S ? nq - qn : (size_t) S;
          const double *src = (tid == 0 ? xq : yq) + qn;
          const size_t lo = ((size_t) src + 15) & ~(size_t) 15, hi = ((size_t) (src + nn)) & ~(size_t) 15;
          if (hi > lo)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned) (hi - lo)) : "memory");
        }
      for (unsigned i = tid; i < (unsigned) nt; i += T) hist[i] = 0;
      __syncthreads();
      unsigned pk[KPT];
#pragma unroll
      for (int k = 0; k < KPT; k++)
        {
          const unsigned i = (unsigned) k * T + tid;
          pk[k] = 0xffffffffu;
          if (i < n)
            {
              const unsigned t = bin_of(M, __ldg(xq + q0 + i), __ldg(yq + q0 + i));
              pk[k] = t | (atomicAdd(hist + t, 1u) << 12);
            }
        }
      __syncthreads();
      unsigned loc[PER], sum = 0;
      const unsigned b = tid * per;
#pragma unroll
      for (unsigned j = 0; j < PER; j++)
        {
          loc[j] = (j < per && b + j < (unsigned) nt) ? hist[b + j] : 0u;
          sum += loc[j];
        }
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        {
          const unsigned v = __shfl_up_sync(FULL, inc, o);
          if (lane >= (unsigned) o) inc += v;
        }
      if (lane == 31) warp_sum[wid] = inc;
      __syncthreads();
      unsigned run = inc - sum;
#pragma unroll
      for (unsigned w = 0; w < WPC; w++)
        if (w < wid) run += warp_sum[w];
#pragma unroll
      for (unsigned j = 0; j < PER; j++)
        if (j < per && b + j < (unsigned) nt)
          {
            hist[b + j] = run;
            /* bin-major run table: neighbouring chunks (sorted by neighbouring CTAs at about the
               same time) fill neighbouring words of a row, so the sectors merge in L2 */
            runs[(size_t) (b + j) * n_chunks + c] = run | ((run + loc[j]) << 16);
            run += loc[j];
          }
      __syncthreads();
      /* x: permute in shared memory, copy out; then y through the same buffer */
#pragma unroll
      for (int k = 0; k < KPT; k++)
        if (pk[k] != 0xffffffffu)
          {
            const unsigned i = (unsigned) k * T + tid;
            const unsigned pos = hist[pk[k] & 0xfffu] + (pk[k] >> 12);
            pk[k] = pos;
            sbuf[pos] = __ldg(xq + q0 + i);
            si[pos] = (unsigned short) i;
          }
      __syncthreads();
      for (unsigned i = tid; i < n; i += T)
        {
          __stcs(bx + q0 + i, sbuf[i]);
          bi[q0 + i] = si[i];
        }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KPT; k++)
        if (pk[k] != 0xffffffffu)
          {
            const unsigned i = (unsigned) k * T + tid;
            sbuf[pk[k]] = __ldg(yq + q0 + i);
          }
      __syncthreads();
      for (unsigned i = tid; i < n; i += T) __stcs(by + q0 + i, sbuf[i]);
      __syncthreads();
    }
}

/* ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk; SASS: UBLKCP, SYNCS) ---------------------------- */
__device__ __forceinline__ unsigned
smem_u32(const void *p)
{
  return (unsigned) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool
mbar_try_wait(unsigned long long *bar, unsigned parity)
{
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(done)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return done != 0;
}

/* global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is
 * signalled on `bar` as a transaction count */
__device__ __forceinline__ void
bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void
fence_proxy_async()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void
fetch_rec_smem(gsi_rec *r, const uint4 *p)
{
  const uint4 a = p[0], b = p[1], c = p[2], d = p[3];
  r->x0x = __hiloint2double((int) a.y, (int) a.x);
  r->x0y = __hiloint2double((int) a.w, (int) a.z);
  r->u00 = __hiloint2double((int) b.y, (int) b.x);
  r->u01 = __hiloint2double((int) b.w, (int) b.z);
  r->u11 = __hiloint2double((int) c.y, (int) c.x);
  r->l10 = __hiloint2double((int) c.w, (int) c.z);
  r->lnk[0] = (int32_t) d.x;
  r->lnk[1] = (int32_t) d.y;
  r->lnk[2] = (int32_t) d.z;
  r->meta = d.w;
}

__device__ __forceinline__ void
fetch_tail_smem(gsi_leaf_tail *t, const uint4 *p)
{
  const uint4 e = p[4], f = p[5];
  t->s.node = (int32_t) e.x;
  t->s.vflags = e.y;
  t->s.z[0] = __hiloint2double((int) e.w, (int) e.z);
  t->s.z[1] = __hiloint2double((int) f.y, (int) f.x);
  t->s.z[2] = __hiloint2double((int) f.w, (int) f.z);
}

struct gsi_tile_args
{
  gsi_tile_map M;
  double cax, cbx, cay, cby; /* grid cell of a query as one fused multiply-add per axis (a hint, like the bin) */
  const double *bx, *by;
  const unsigned short *bi;
  const unsigned *runs; /* [n_bins][n_chunks] */
  unsigned *work;       /* work-item counter (zero at launch) */
  unsigned n_chunks;
  int S, nt, ranges;
  int leaf_cap, seed_rows, seed_pitch;
  double *zb;
  int *lb;
  size_t ldb;
  unsigned *list;
  unsigned list_cap;
  unsigned *ovf_mask;
  unsigned *counters;
  int max_steps;
};

#define GSI_SEED_FAR 0xffffu /* staged seed cell whose leaf lies outside the staged slot range */
#define GSI_TILE_RING 64     /* per-warp ring of queries whose walk left the staged tile */

/* shared memory of k_tile_walk: leaf records | reciprocals of their two divisors | seed cells | per
 * warp: a dense position list and a ring of stragglers (position, leaf slot) */
__host__ __device__ inline size_t
tile_smem_seeds_at(int leaf_cap)
{
  return (size_t) leaf_cap * (96 + 16);
}

__host__ __device__ inline size_t
tile_smem_warp_at(int leaf_cap, int seed_rows, int seed_pitch)
{
  return (tile_smem_seeds_at(leaf_cap) + (size_t) seed_rows * seed_pitch * 2 + 15) & ~(size_t) 15;
}

#define GSI_TILE_WARP_WORDS (GSI_TILE_SPOS + 2 * GSI_TILE_RING)

/* Predicated streaming loads INTO an existing register: no temporary and no move, so nothing
 * touches the value before its consumer one loop iteration later (a compiler-placed move right
 * behind a plain load makes the warp wait for HBM at once -- measured: 19 % of all stall samples). */
__device__ __forceinline__ void
ld_stream_if(bool p, double &dst, const double *src)
{
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.cs.f64 %0, [%1];\n\t}"
               : "+d"(dst)
               : "l"(src), "r"((unsigned) p)
               : "memory");
}

__device__ __forceinline__ void
ld_nc_if(bool p, unsigned &dst, const unsigned *src)
{
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
               : "+r"(dst)
               : "l"(src), "r"((unsigned) p)
               : "memory");
}

template <bool STATS, bool ONE_FIELD>
__global__ void __launch_bounds__(GSI_TILE_THREADS, 1)
k_tile_walk(const gsi_view V, const gsi_tile_args A)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned long long mbar;
  __shared__ unsigned s_item, s_next_block;
  uint4 *sleaf = reinterpret_cast<uint4 *>(smem_raw);
  double2 *srcp = reinterpret_cast<double2 *>(smem_raw + (size_t) A.leaf_cap * 96);
  unsigned short *sseed = reinterpret_cast<unsigned short *>(smem_raw + tile_smem_seeds_at(A.leaf_cap));
  const unsigned FULL = 0xffffffffu;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned *spos = reinterpret_cast<unsigned *>(smem_raw + tile_smem_warp_at(A.leaf_cap, A.seed_rows, A.seed_pitch))
                   + (size_t) warp * GSI_TILE_WARP_WORDS;
  unsigned *sg_pos = spos + GSI_TILE_SPOS;
  int *sg_slot = reinterpret_cast<int *>(sg_pos + GSI_TILE_RING);
  const unsigned n_blocks = (A.n_chunks + 31u) >> 5;
  const unsigned n_items = (unsigned) A.nt * (unsigned) A.ranges;
  unsigned long long steps_acc = 0;
  unsigned parity = 0;
  if (tid == 0) mbar_init(&mbar, 1);

  /* warp-aggregated append to the deferred list (original query index of binned position pos) */
  auto append_deferred = [&](bool defer, unsigned pos) {
    const unsigned dm = __ballot_sync(FULL, defer);
    if (!dm) return;
    const int leader = __ffs(dm) - 1;
    unsigned at0 = 0;
    if (lane == (unsigned) leader) at0 = atomicAdd(A.counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
    at0 = __shfl_sync(FULL, at0, leader);
    if (defer)
      {
        const size_t q = (size_t) (pos / (unsigned) A.S) * (size_t) A.S + A.bi[pos];
        const unsigned at = at0 + __popc(dm & lt_mask);
        if (at < A.list_cap)
          A.list[at] = (unsigned) q;
        else
          mark_overflow(A.ovf_mask, q, A.counters);
      }
  };
  /* result of an accepted query */
  auto store_result = [&](unsigned pos, const gsi_leaf_tail &tl, double c0, double c1) {
    if (A.lb) __stcs(A.lb + pos, tl.g.node);
    if (A.zb)
      {
        if (ONE_FIELD)
          __stcs(A.zb + pos, gsi_blend(c0, c1, (int) (tl.s.vflags & 1u), tl.s.z[0], (int) (tl.s.vflags & 2u), tl.s.z[1],
                                       (int) (tl.s.vflags & 4u), tl.s.z[2]));
        else
          gsi_blend_store(&V, &tl, c0, c1, A.zb, A.ldb, pos);
      }
  };

  for (;;)
    {
      if (tid == 0) s_item = atomicAdd(A.work, 1u);
      __syncthreads(); /* every warp is done with the previous item's shared memory */
      const unsigned item = s_item;
      if (item >= n_items) break;
      const unsigned bin = item / (unsigned) A.ranges, range = item - bin * (unsigned) A.ranges;
      const unsigned prefix = bin_to_prefix(A.M, bin);
      const int slot0 = V.tiles[prefix];
      const int n_staged = min(V.tiles[prefix + 1] - slot0, A.leaf_cap);
      const bool packed = V.leaf_stride == (int) sizeof(gsi_leaf);
      /* ---- stage the tile: leaf records by TMA bulk copy, seed cells by the threads ---- */
      if (tid == 0)
        {
          s_next_block = 0;
          if (packed && n_staged > 0)
            {
              const unsigned bytes = (unsigned) n_staged * 96u;
              const char *src = V.leaf + (size_t) slot0 * sizeof(gsi_leaf);
              fence_proxy_async();
              mbar_arrive_expect_tx(&mbar, bytes);
              for (unsigned o = 0; o < bytes; o += 32768u)
                bulk_load(reinterpret_cast<char *>(sleaf) + o, src + o, min(32768u, bytes - o), &mbar);
            }
        }
      if (!packed)
        for (int i = tid; i < n_staged * 6; i += blockDim.x)
          sleaf[i] = __ldg(reinterpret_cast<const uint4 *>(V.leaf + (size_t) (slot0 + i / 6) * (size_t) V.leaf_stride) + i % 6);
      /* the tile's rectangle of grid cells, one cell of margin */
      const unsigned tx = bin & ((1u << A.M.bx_bits) - 1u), ty = bin >> A.M.bx_bits;
      const int cx0 = max(0, (int) (((unsigned long long) tx * (unsigned) V.grid_g) >> A.M.bx_bits) - 1);
      const int cy0 = max(0, (int) (((unsigned long long) ty * (unsigned) V.grid_g) >> A.M.by_bits) - 1);
      const int ncx = min(A.seed_pitch, V.grid_g - cx0), ncy = min(A.seed_rows, V.grid_g - cy0);
      for (int i = tid; i < ncy * A.seed_pitch; i += blockDim.x)
        {
          const int ry = i / A.seed_pitch, rx = i - ry * A.seed_pitch;
          if (rx < ncx)
            {
              const unsigned rel = (unsigned) (__ldg(V.seeds + (size_t) (cy0 + ry) * (size_t) V.grid_g + (size_t) (cx0 + rx)) - slot0);
              sseed[i] = rel < (unsigned) n_staged && rel < GSI_SEED_FAR ? (unsigned short) rel : (unsigned short) GSI_SEED_FAR;
            }
        }
      if (packed && n_staged > 0)
        {
          while (!mbar_try_wait(&mbar, parity)) {}
          parity ^= 1u;
        }
      else
        __syncthreads(); /* the fallback copy of the records is visible */
      /* reciprocals of the two divisors of every staged record (gsi_div_rcp): the walk then divides
         with three operations per quotient.  A divisor the hoisted division cannot serve makes the
         leaf defer always (tau = +inf), like a singular leaf. */
      for (int i = tid; i < n_staged; i += blockDim.x)
        {
          const uint4 b = sleaf[(size_t) i * 6 + 1], c = sleaf[(size_t) i * 6 + 2];
          const double u00 = __hiloint2double((int) b.y, (int) b.x), u11 = __hiloint2double((int) c.y, (int) c.x);
          srcp[i] = make_double2(gsi_div_rcp(u11), gsi_div_rcp(u00));
          if (!(gsi_div_rcp_usable(u00) && gsi_div_rcp_usable(u11)))
            {
              unsigned *meta = reinterpret_cast<unsigned *>(sleaf + (size_t) i * 6 + 3) + 3;
              *meta = (*meta & 0x80000000u) | 0x7f800000u;
            }
        }
      __syncthreads(); /* seeds, records, reciprocals and the block counter are visible */

      /* ---- walk the tile's queries ----
       * The item's 32-chunk blocks (block k of this item = chunks of block range + k * ranges) are
       * handed to the warps by a shared counter; the run word of a warp's next block is loaded while
       * it still works on the current one. */
      const unsigned *runs_row = A.runs + (size_t) bin * A.n_chunks;
      unsigned cb_nxt = 0, run_nxt = 0;
      auto grab_block = [&]() {
        unsigned k = 0;
        if (lane == 0) k = atomicAdd(&s_next_block, 1u);
        k = __shfl_sync(FULL, k, 0);
        cb_nxt = range + k * (unsigned) A.ranges;
        const unsigned c = (cb_nxt << 5) + lane;
        run_nxt = 0;
        ld_nc_if(cb_nxt < n_blocks && c < A.n_chunks, run_nxt, runs_row + c);
      };
      grab_block();
      unsigned r_pos0 = 0, r_len = 0, r_excl = 0; /* this lane's run of the current block */
      unsigned b_total = 0, win_lo = 0, win_n = 0, next = 0; /* block items, window [win_lo, win_lo + win_n), cursor */
      unsigned sg_head = 0, sg_count = 0;                     /* ring of stragglers (warp-uniform) */
      /* a lane holds two queries: cur (walking) and nxt (coordinates on their way from HBM, requested one
         loop iteration -- longer than the HBM latency -- before they are needed) */
      bool cur_v = false, nxt_v = false;
      double cur_x = 0, cur_y = 0, nxt_x = 0, nxt_y = 0;
      unsigned cur_pos = 0, nxt_pos = 0;
      int cur_slot = 0, cur_steps = 0;

      /* Queries whose walk left the staged slot range (seed or neighbour in another tile; ~6 %) wait
         in the ring until a full batch can take the global-memory path together: one lane fetching
         from HBM while 31 wait was 13 % of the kernel. */
      auto run_stragglers = [&](unsigned n) {
        bool act = lane < n, dfr = false;
        unsigned pos = 0;
        int slot = 0, steps = 0;
        double x = 0, y = 0;
        if (act)
          {
            const unsigned e = (sg_head + lane) & (GSI_TILE_RING - 1);
            pos = sg_pos[e];
            slot = sg_slot[e];
            x = __ldcs(A.bx + pos);
            y = __ldcs(A.by + pos);
            if (slot < 0)
              {
                const int ix = min(max(__double2int_rd(__fma_rn(x, A.cax, A.cbx)), 0), V.grid_g - 1);
                const int iy = min(max(__double2int_rd(__fma_rn(y, A.cay, A.cby)), 0), V.grid_g - 1);
                slot = gsi_ld_seed(V.seeds + (size_t) iy * (size_t) V.grid_g + (size_t) ix);
              }
          }
        sg_head = (sg_head + n) & (GSI_TILE_RING - 1);
        sg_count -= n;
        while (__any_sync(FULL, act))
          {
            if (act)
              {
                gsi_rec rec;
                const int at = slot;
                double c0, c1;
                gsi_fetch_rec(&rec, gsi_leaf_at(&V, at));
                const int st = gsi_walk_step(&V.xf, &rec, x, y, &slot, &c0, &c1);
                ++steps;
                if (st == GSI_STEP_ACCEPT)
                  {
                    gsi_leaf_tail tl;
                    gsi_fetch_tail(&tl, gsi_leaf_at(&V, at));
                    store_result(pos, tl, c0, c1);
                    if (STATS) steps_acc += (unsigned) steps;
                    act = false;
                  }
                else if (st == GSI_STEP_DEFER || steps >= A.max_steps)
                  {
                    dfr = true;
                    act = false;
                  }
              }
          }
        append_deferred(dfr, pos);
        __syncwarp();
      };

      for (;;)
        {
          if (next >= win_n) /* warp-uniform: the dense list is used up */
            {
              bool have = true;
              win_lo += win_n; /* items of the current block handed out so far */
              if (win_lo < b_total)
                ; /* next window of a block that exceeds the list */
              else if (cb_nxt < n_blocks)
                {
                  const unsigned c = (cb_nxt << 5) + lane;
                  const unsigned start = run_nxt & 0xffffu;
                  r_len = (run_nxt >> 16) - start;
                  r_pos0 = c * (unsigned) A.S + start;
                  unsigned incl = r_len;
#pragma unroll
                  for (int o = 1; o < 32; o <<= 1)
                    {
                      const unsigned v = __shfl_up_sync(FULL, incl, o);
                      if (lane >= (unsigned) o) incl += v;
                    }
                  r_excl = incl - r_len;
                  b_total = __shfl_sync(FULL, incl, 31);
                  win_lo = 0;
                  grab_block();
                }
              else
                have = false;
              next = 0;
              win_n = 0;
              if (have)
                {
                  win_n = min((unsigned) GSI_TILE_SPOS, b_total - win_lo);
                  /* items [win_lo, win_lo + win_n) of the block's concatenated runs */
                  const unsigned k0 = win_lo > r_excl ? win_lo - r_excl : 0u;
                  const unsigned k1 = min(r_len, win_lo + win_n > r_excl ? win_lo + win_n - r_excl : 0u);
                  for (unsigned k = k0; k < k1; k++) spos[r_excl + k - win_lo] = r_pos0 + k;
                  __syncwarp();
                  if (win_n == 0) continue; /* empty block */
                }
              else if (!__any_sync(FULL, cur_v || nxt_v))
                break;
            }
          /* refill: a lane without a query takes over its prefetched one and requests the next position of the list */
          bool defer = false;
          const bool need = !cur_v;
          const unsigned take = __ballot_sync(FULL, need);
          if (need)
            {
              cur_v = nxt_v; cur_pos = nxt_pos; cur_x = nxt_x; cur_y = nxt_y;
              cur_steps = 0;
              const unsigned off = next + __popc(take & lt_mask);
              nxt_v = off < win_n;
              if (nxt_v) nxt_pos = spos[off];
              ld_stream_if(nxt_v, nxt_x, A.bx + nxt_pos);
              ld_stream_if(nxt_v, nxt_y, A.by + nxt_pos);
              if (cur_v)
                {
                  /* grid cell (one fused multiply-add per axis; NaN -> 0, +-inf saturate) and its seed from
                     the staged block when it is inside */
                  const int ix = min(max(__double2int_rd(__fma_rn(cur_x, A.cax, A.cbx)), 0), V.grid_g - 1);
                  const int iy = min(max(__double2int_rd(__fma_rn(cur_y, A.cay, A.cby)), 0), V.grid_g - 1);
                  const unsigned rx = (unsigned) (ix - cx0), ry = (unsigned) (iy - cy0);
                  unsigned rel = GSI_SEED_FAR;
                  if (rx < (unsigned) ncx && ry < (unsigned) ncy) rel = sseed[ry * A.seed_pitch + rx];
                  /* a seed outside the staged range is looked up by the straggler path (slot -1) */
                  cur_slot = rel != GSI_SEED_FAR ? slot0 + (int) rel : -1;
                  if (!(isfinite(cur_x) && isfinite(cur_y)))
                    {
                      defer = true;
                      cur_v = false;
                    }
                }
            }
          next = min(next + (unsigned) __popc(take), win_n);
          /* one walk step out of shared memory for every lane that holds a query inside the staged tile */
          const unsigned rel = (unsigned) (cur_slot - slot0);
          const bool staged = rel < (unsigned) n_staged;
          if (cur_v && staged)
            {
              gsi_rec rec;
              double c0, c1;
              fetch_rec_smem(&rec, sleaf + (size_t) rel * 6);
              const double2 rc = srcp[rel];
              const int st = gsi_walk_step_rcp(&V.xf, &rec, rc.x, rc.y, cur_x, cur_y, &cur_slot, &c0, &c1);
              ++cur_steps;
              if (st == GSI_STEP_ACCEPT)
                {
                  gsi_leaf_tail tl;
                  fetch_tail_smem(&tl, sleaf + (size_t) rel * 6);
                  store_result(cur_pos, tl, c0, c1);
                  if (STATS) steps_acc += (unsigned) cur_steps;
                  cur_v = false;
                }
              else if (st == GSI_STEP_DEFER || cur_steps >= A.max_steps)
                {
                  defer = true;
                  cur_v = false;
                }
            }
          /* lanes whose leaf is not staged hand their query to the ring */
          const bool push = cur_v && !staged;
          const unsigned pm = __ballot_sync(FULL, push);
          if (pm)
            {
              if (push)
                {
                  const unsigned e = (sg_head + sg_count + __popc(pm & lt_mask)) & (GSI_TILE_RING - 1);
                  sg_pos[e] = cur_pos;
                  sg_slot[e] = cur_slot;
                  cur_v = false;
                }
              sg_count += __popc(pm);
              __syncwarp();
            }
          append_deferred(defer, cur_pos);
          if (sg_count >= 32) run_stragglers(32);
        }
      while (sg_count > 0) run_stragglers(min(sg_count, 32u));
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(A.counters + GSI_CNT_STEPS), steps_acc);
}

__global__ void __launch_bounds__(GSI_UNPERMUTE_THREADS, 2)
k_unpermute(const double *__restrict__ zb, const int *__restrict__ lb, const unsigned short *__restrict__ bi,
            size_t nq, int S, int n_fields, size_t ldb, double *__restrict__ val, size_t ld,
            int *__restrict__ leaf_out)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sz = reinterpret_cast<double *>(smem_raw);
  int *sl = reinterpret_cast<int *>(sz + S);
  const unsigned tid = threadIdx.x;
  const size_t n_chunks = (nq + (size_t) S - 1) / (size_t) S;
  constexpr int KPT = GSI_BIN_CHUNK_MAX / GSI_UNPERMUTE_THREADS;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x)
    {
      const size_t q0 = c * (size_t) S;
      const unsigned n = (unsigned) (nq - q0 < (size_t) S ? nq - q0 : (size_t) S);
      unsigned short idx[KPT];
#pragma unroll
      for (int k = 0; k < KPT; k++)
        {
          const unsigned i = (unsigned) k * GSI_UNPERMUTE_THREADS + tid;
          idx[k] = 0;
          if (i < n)
            {
              idx[k] = bi[q0 + i];
              if (leaf_out) sl[idx[k]] = __ldcs(lb + q0 + i);
            }
        }
      const int passes = val ? n_fields : 1;
      for (int f = 0; f < passes; f++)
        {
          if (val)
            {
#pragma unroll
              for (int k = 0; k < KPT; k++)
                {
                  const unsigned i = (unsigned) k * GSI_UNPERMUTE_THREADS + tid;
                  if (i < n) sz[idx[k]] = __ldcs(zb + (size_t) f * ldb + q0 + i);
                }
            }
          __syncthreads();
          for (unsigned i = tid; i < n; i += GSI_UNPERMUTE_THREADS)
            {
              if (val) __stcs(val + (size_t) f * ld + q0 + i, sz[i]);
              if (f == 0 && leaf_out) __stcs(leaf_out + q0 + i, sl[i]);
            }
          __syncthreads();
        }
    }
}

/* derivative evaluators: gradient of the leaf's plane for leaf ids produced by a locate pass */
__global__ void __launch_bounds__(256)
k_deriv(const gsi_view V, const int *__restrict__ leaf, size_t nq, double *dzdx, double *dzdy, size_t ld)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride)
    gsi_deriv_one(&V, __ldg(leaf + q), dzdx, dzdy, ld, q);
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

/* Range check of every index a kernel will follow in an adopted slab image (file, peer process):
 * DAG links, leaf slots, neighbour slots, vertex slots, node ids, seed cells, entry nodes, tile
 * boundaries.  A foreign or damaged image is rejected with GSL_SCATTERED_EINVAL instead of
 * faulting the context. */
__global__ void __launch_bounds__(256)
k_validate_slab(const gsi_view V, unsigned long long n_points, unsigned long long *bad)
{
  const size_t cells = (size_t) V.grid_g * (size_t) V.grid_g;
  const size_t nt1 = ((size_t) 1 << V.tile_bits) + 1;
  size_t n = (size_t) V.n_dag;
  if ((size_t) V.n_leaves > n) n = (size_t) V.n_leaves;
  if (cells > n) n = cells;
  if (nt1 > n) n = nt1;
  if (V.memo && (size_t) V.memo_mask + 1 > n) n = (size_t) V.memo_mask + 1;
  unsigned long long mine = 0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      if (i < (size_t) V.n_dag)
        {
          const gsi_rec &r = V.dag[i];
          const uint32_t type = r.meta & GSI_META_TYPE_MASK;
          if (type == GSI_LEAF)
            mine += (unsigned) r.lnk[0] >= (unsigned) V.n_leaves;
          else
            for (int k = 0; k < 3; k++) mine += (unsigned) r.lnk[k] >= (unsigned) V.n_dag;
        }
      if (i < (size_t) V.n_leaves)
        {
          const gsi_leaf *lf = gsi_leaf_at(&V, (int) i);
          for (int k = 0; k < 3; k++)
            {
              mine += lf->r.lnk[k] < -1 || lf->r.lnk[k] >= V.n_leaves;
              if (V.n_fields > 1) mine += lf->t.g.v[k] < -1 || (lf->t.g.v[k] >= 0 && (unsigned long long) lf->t.g.v[k] >= n_points);
            }
          mine += (unsigned) lf->t.g.node >= (unsigned) V.n_dag;
        }
      if (i < cells)
        {
          mine += (unsigned) V.seeds[i] >= (unsigned) V.n_leaves;
          {
            const size_t iy = i / (size_t) V.grid_g, ix = i - iy * (size_t) V.grid_g;
            const unsigned off = V.seed16[i];
            if (off != GSI_SEED16_ESC)
              mine += (unsigned) (V.seedbase[(iy >> V.seed_shift) * (size_t) V.grid_gb + (ix >> V.seed_shift)] + (int) off) >= (unsigned) V.n_leaves;
          }
          if (V.entry) mine += (unsigned) V.entry[i] >= (unsigned) V.n_dag;
        }
      if (V.memo && i <= (size_t) V.memo_mask)
        {
          const gsi_memo *e = V.memo + i;
          mine += e->state > GSI_MEMO_SKIP
                  || (e->state == GSI_MEMO_FULL && ((unsigned) e->node >= (unsigned) V.n_dag || (unsigned) e->slot >= (unsigned) V.n_leaves));
        }
      if (i < nt1)
        {
          const int t = V.tiles[i];
          mine += t < 0 || t > V.n_leaves || (i + 1 < nt1 && V.tiles[i + 1] < t) || (i + 1 == nt1 && t != V.n_leaves);
        }
    }
  if (mine) atomicAdd(bad, mine);
}

/* self-test of the hoisted-reciprocal division (gsi_core.h): operand pairs from four families
 * (data-like magnitudes, raw random bit patterns, mixed, scaled over the whole exponent range);
 * out[0] = pairs the fast path vouched for, out[1] = of those, quotients that differ from __ddiv_rn */
__global__ void __launch_bounds__(256)
k_selftest_division(uint64_t seed, size_t n, unsigned long long *out)
{
  unsigned long long ok = 0, bad = 0;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      uint64_t z[3];
      for (int k = 0; k < 3; k++)
        {
          uint64_t v = seed + (3 * i + k + 1) * 0x9E3779B97F4A7C15ull;
          v ^= v >> 30; v *= 0xBF58476D1CE4E5B9ull;
          v ^= v >> 27; v *= 0x94D049BB133111EBull;
          v ^= v >> 31;
          z[k] = v;
        }
      const double ua = ((double) (z[0] >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 8.0;
      const double ub = ((double) (z[1] >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 8.0;
      double a, b;
      switch (i & 3)
        {
        case 0: a = ua; b = ub; break;
        case 1: a = __longlong_as_double((long long) z[0]); b = __longlong_as_double((long long) z[1]); break;
        case 2: a = __longlong_as_double((long long) z[0]); b = ub; break;
        default:
          a = scalbn(ua, (int) (z[2] % 2200u) - 1100);
          b = scalbn(ub, (int) ((z[2] >> 32) % 2200u) - 1100);
          break;
        }
      if (!gsi_div_rcp_usable(b)) continue;
      unsigned flag = 0;
      const double q = gsi_div_by_rcp(a, b, gsi_div_rcp(b), &flag);
      if (flag) continue;
      ok++;
      bad += __double_as_longlong(q) != __double_as_longlong(__ddiv_rn(a, b));
    }
  for (int o = 16; o > 0; o >>= 1)
    {
      ok += __shfl_xor_sync(0xffffffffu, ok, o);
      bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
  if ((threadIdx.x & 31) == 0)
    {
      atomicAdd(out, ok);
      if (bad) atomicAdd(out + 1, bad);
    }
}

/* ---- batched point location for the host builder -----------------------
 * The builder (gsi_builder.cpp) locates every point it inserts with find_leaf's descent.  A node's
 * children are final once it has them, so the descent of point i through a SNAPSHOT of the growing
 * tree is the prefix of its real descent: one thread per point replays it on the snapshot (the
 * same GSI_DEFINE_DESCEND as the evaluation replay and the host builder: identical operations,
 * identical results) and returns the snapshot leaf it ends in plus the state it arrived with; the
 * host finishes the few levels the tree has grown since.  The snapshot lives in three arrays as on
 * the host: LU records (written once per node), links and types (re-uploaded every round). */
struct gsi_build_view
{
  const gsi_rec *rec;
  const int32_t *lnk;
  const uint8_t *type;
};

#define GSI_FETCH_BUILD(dst, idx)                                                              \
  do                                                                                           \
    {                                                                                          \
      const size_t i_ = (size_t) (idx);                                                        \
      gsi_fetch_rec(&(dst), dag->rec + i_);                                                    \
      (dst).lnk[0] = __ldg(dag->lnk + 3 * i_);                                                 \
      (dst).lnk[1] = __ldg(dag->lnk + 3 * i_ + 1);                                             \
      (dst).lnk[2] = __ldg(dag->lnk + 3 * i_ + 2);                                             \
      (dst).meta = ((dst).meta & (GSI_META_SWAP | GSI_META_SINGULAR)) | (uint32_t) __ldg(dag->type + i_); \
    }                                                                                          \
  while (0)
GSI_DEFINE_DESCEND(descend_build, __device__ __forceinline__, const gsi_build_view *, GSI_FETCH_BUILD)

__global__ void __launch_bounds__(128, 4)
k_build_locate(const gsi_build_view B, const gsi_xform xf, const double *__restrict__ px,
               const double *__restrict__ py, size_t first, size_t count, int32_t *node, double *c0,
               double *c1, int32_t *last)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride)
    {
      double a = 0, b = 0;
      int l = -1;
      gsi_rec leaf_rec;
      node[k] = descend_build(&B, &xf, 0, __ldg(px + first + k), __ldg(py + first + k), &a, &b, &l, &leaf_rec, 0);
      c0[k] = a;
      c1[k] = b;
      last[k] = l;
    }
}

__global__ void
k_accumulate_deferred(unsigned *counters)
{
  const unsigned c = counters[GSI_CNT_DEFERRED];
  unsigned long long *tot = reinterpret_cast<unsigned long long *>(counters + GSI_CNT_DEFERRED_TOTAL);
  *tot += c;
  counters[GSI_CNT_OVERFLOW] = 0;
  counters[GSI_CNT_WORK] = 0;
  counters[GSI_CNT_DEFERRED2] = 0;
}

inline int
grid_for(size_t n, int threads, int sm_count, int per_sm)
{
  const size_t want = (n + threads - 1) / threads;
  const size_t cap = (size_t) sm_count * per_sm;
  return (int) (want < cap ? (want ? want : 1) : cap);
}

/* the reorder pipeline applies to query arrays only (rasters are coherent by construction) */
inline bool
launch_reordered(const gsi_view &, const src_raster &, size_t, double *, int *, size_t,
                 const gsi_launch_cfg &, unsigned, cudaStream_t)
{
  return false;
}

inline bool
launch_reordered(const gsi_view &V, const src_arrays &src, size_t nq, double *val, int *leaf_out,
                 size_t ld, const gsi_launch_cfg &cfg, unsigned list_cap, cudaStream_t st)
{
  const gsi_reorder_scratch *R = cfg.reorder;
  if (!R || nq > R->cap || R->chunk < 32 || R->chunk > GSI_BIN_CHUNK_MAX || V.tile_bits < 1 || V.tile_bits > 12)
    return false;
  const int S = R->chunk, P = V.tile_bits, nt = 1 << P;
  const size_t n_chunks = (nq + (size_t) S - 1) / (size_t) S;
  gsi_tile_map M;
  M.bx_bits = P >> 1;
  M.by_bits = (P + 1) >> 1;
  {
    /* tx = floor(((qx - cx) * sx - u0) * inv_du * 2^bx_bits / G) as one fused multiply-add */
    const double kx = V.inv_du * (double) (1 << M.bx_bits) / (double) V.grid_g;
    const double ky = V.inv_dv * (double) (1 << M.by_bits) / (double) V.grid_g;
    M.ax = V.xf.sx * kx;
    M.bx = (-V.xf.cx * V.xf.sx - V.u0) * kx;
    M.ay = V.xf.sy * ky;
    M.by = (-V.xf.cy * V.xf.sy - V.v0) * ky;
  }
  const size_t smem_bin = (size_t) S * 10 + (size_t) nt * 4, smem_unp = (size_t) S * 12;
  /* per device (the current one), so not cached in a process-wide flag */
  cudaFuncSetAttribute(k_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, GSI_BIN_CHUNK_MAX * 10 + 4096 * 4);
  cudaFuncSetAttribute(k_unpermute, cudaFuncAttributeMaxDynamicSharedMemorySize, GSI_BIN_CHUNK_MAX * 12);
  const int g_bin = (int) (n_chunks < (size_t) (2 * cfg.sm_count) ? n_chunks : (size_t) (2 * cfg.sm_count));
  k_bin<<<g_bin, GSI_BIN_THREADS, smem_bin, st>>>(M, src.x, src.y, nq, S, nt, R->bx, R->by, R->bi, R->runs);

  /* tile kernel: one persistent CTA per SM; shared memory = the largest tile's records (capped) +
     the seed cells under a tile's rectangle (16-bit, relative to the tile's first slot) + a dense
     position list per warp */
  int threads = R->tile_threads < 32 ? 32 : (R->tile_threads > GSI_TILE_THREADS ? GSI_TILE_THREADS : R->tile_threads);
  threads &= ~31;
  const int wpc = threads / 32;
  gsi_tile_args A;
  A.M = M;
  A.bx = R->bx; A.by = R->by; A.bi = R->bi; A.runs = R->runs; A.work = R->work;
  A.n_chunks = (unsigned) n_chunks;
  A.S = S; A.nt = nt;
  A.seed_pitch = (V.grid_g >> M.bx_bits) + 4;
  A.seed_rows = (V.grid_g >> M.by_bits) + 4;
  const size_t budget = 227 * 1024 - 512;
  const size_t spos_bytes = (size_t) wpc * GSI_TILE_WARP_WORDS * 4; /* dense list + straggler ring per warp */
  size_t seed_bytes = (size_t) A.seed_pitch * A.seed_rows * 2;
  if (seed_bytes + spos_bytes + 4096 > budget)
    {
      /* a grid far finer than the tiles (explicit "grid_g"): seeds come from global memory */
      A.seed_pitch = A.seed_rows = 0;
      seed_bytes = 0;
    }
  int leaf_cap = R->tile_max_leaves < 1 ? 1 : R->tile_max_leaves;
  const size_t room = (budget - seed_bytes - spos_bytes - 16) / (96 + 16);
  if ((size_t) leaf_cap > room) leaf_cap = (int) room;
  if (leaf_cap > 65000) leaf_cap = 65000; /* 16-bit relative seeds */
  A.leaf_cap = leaf_cap;
  const size_t smem_tile = tile_smem_warp_at(leaf_cap, A.seed_rows, A.seed_pitch) + spos_bytes;
  /* cell = floor(((q - c) * s - u0) * inv_du) as one fused multiply-add per axis */
  A.cax = V.xf.sx * V.inv_du;
  A.cbx = (-V.xf.cx * V.xf.sx - V.u0) * V.inv_du;
  A.cay = V.xf.sy * V.inv_dv;
  A.cby = (-V.xf.cy * V.xf.sy - V.v0) * V.inv_dv;
  /* work items per tile: a few dozen items per CTA for balance, at least two 32-chunk blocks per warp */
  const size_t n_blocks = (n_chunks + 31) / 32;
  int ranges = R->ranges;
  if (ranges <= 0) ranges = (24 * cfg.sm_count + nt - 1) / nt;
  const int ranges_max = (int) (n_blocks / (size_t) (2 * wpc));
  if (ranges > ranges_max) ranges = ranges_max;
  if (ranges < 1) ranges = 1;
  A.ranges = ranges;
  A.zb = val ? R->zb : nullptr;
  A.lb = leaf_out ? R->lb : nullptr;
  A.ldb = R->cap;
  A.list = cfg.list; A.list_cap = list_cap; A.ovf_mask = cfg.ovf_mask; A.counters = cfg.counters;
  A.max_steps = cfg.max_steps;
  cudaMemsetAsync(R->work, 0, sizeof(unsigned), st);
  const unsigned n_items = (unsigned) nt * (unsigned) ranges;
  const int g_tile = (int) (n_items < (unsigned) cfg.sm_count ? n_items : (unsigned) cfg.sm_count);
#define GSI_LAUNCH_TILE(STATS, ONE)                                                             \
  do                                                                                            \
    {                                                                                           \
      cudaFuncSetAttribute(k_tile_walk<STATS, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget); \
      k_tile_walk<STATS, ONE><<<g_tile, threads, smem_tile, st>>>(V, A);                        \
    }                                                                                           \
  while (0)
  if (cfg.stats && V.n_fields == 1) GSI_LAUNCH_TILE(true, true);
  else if (cfg.stats) GSI_LAUNCH_TILE(true, false);
  else if (V.n_fields == 1) GSI_LAUNCH_TILE(false, true);
  else GSI_LAUNCH_TILE(false, false);
#undef GSI_LAUNCH_TILE
  const int g_unp = (int) (n_chunks < (size_t) (2 * cfg.sm_count) ? n_chunks : (size_t) (2 * cfg.sm_count));
  k_unpermute<<<g_unp, GSI_UNPERMUTE_THREADS, smem_unp, st>>>(R->zb, R->lb, R->bi, nq, S, V.n_fields, R->cap, val, ld,
                                                              leaf_out);
  return true;
}

template <class SRC>
cudaError_t
launch_eval(const gsi_view &V, const SRC &src, size_t nq, double *val, int *leaf_out, size_t ld,
            const gsi_launch_cfg &cfg, cudaStream_t st)
{
  if (nq == 0) return cudaSuccess;
  if (nq > 0xfffffff0ull) return cudaErrorInvalidValue; /* list entries are 32-bit */
  const unsigned list_cap = (unsigned) (cfg.list_capacity < 0xfffffff0ull ? cfg.list_capacity : 0xfffffff0ull);
  if (cfg.mode == GSI_MODE_EXACT)
    {
      if (cfg.ev_exact0) cudaEventRecord(cfg.ev_exact0, st);
      k_exact_all<SRC><<<grid_for(nq, 128, cfg.sm_count, 16), 128, 0, st>>>(
          V, src, nq, val, leaf_out, ld, cfg.counters);
      if (cfg.ev_exact1) cudaEventRecord(cfg.ev_exact1, st);
      return cudaGetLastError();
    }
  cudaError_t e = cudaMemsetAsync(cfg.counters + GSI_CNT_DEFERRED, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  if (cfg.ev_locate0) cudaEventRecord(cfg.ev_locate0, st);
  if (launch_reordered(V, src, nq, val, leaf_out, ld, cfg, list_cap, st))
    ;
  else if (cfg.kernel == GSI_KERNEL_WALK || (cfg.kernel == GSI_KERNEL_AUTO && !SRC::coherent))
    {
      const int ctas = cfg.sm_count * cfg.walk_ctas_per_sm;
      const size_t n_warps = (size_t) ctas * (GSI_WALK_THREADS / 32);
      /* per-warp chunk: ~8 chunks per warp for balance, 256..8192 queries, multiple of 32; the
         continuous queue of k_walkp pays nothing per chunk, so its chunks stay small */
      size_t chunk = (nq / (n_warps * 8) + 31) & ~(size_t) 31;
      if (chunk < 256) chunk = 256;
      if (chunk > (cfg.walk_variant ? 2048u : 8192u)) chunk = cfg.walk_variant ? 2048 : 8192;
      if (cfg.walk_chunk) chunk = cfg.walk_chunk;
      const size_t n_chunks = (nq + chunk - 1) / chunk;
      const int g = (int) ((n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32) < (size_t) ctas
                               ? (n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32)
                               : (size_t) ctas);
      if (cfg.walk_variant == 2 && V.leaf_stride == 128)
        {
          /* leaf lines through L1?  Yes when the whole leaf section is small enough for hot leaves to matter
             more than the L1 space of the in-flight misses (profiles/r2z_*): up to 64 MB of records */
          const int l1_alloc = cfg.walk_l1 >= 0 ? cfg.walk_l1 : ((size_t) V.n_leaves * 128u <= ((size_t) 64 << 20));
          const int wpc = GSI_WALKC_THREADS / 32; /* warps per CTA */
          const int ctas_c = GSI_WALKC_THREADS == 128 ? cfg.sm_count * (cfg.walk_ctas_per_sm > 5 ? cfg.walk_ctas_per_sm : 8)
                                                      : cfg.sm_count * (1024 / GSI_WALKC_THREADS);
          const int gc = (int) ((n_chunks + wpc - 1) / wpc < (size_t) ctas_c ? (n_chunks + wpc - 1) / wpc : (size_t) ctas_c);
          unsigned long long *trace = (size_t) gc <= cfg.walk_trace_ctas ? cfg.walk_trace : nullptr;
          if (trace) cudaMemsetAsync(trace, 0, cfg.walk_trace_ctas * 4 * sizeof(unsigned long long), st);
          gsi_cell_map CM;
          CM.ax = V.xf.sx * V.inv_du;
          CM.bx = (-V.xf.cx * V.xf.sx - V.u0) * V.inv_du;
          CM.ay = V.xf.sy * V.inv_dv;
          CM.by = (-V.xf.cy * V.xf.sy - V.v0) * V.inv_dv;
#define GSI_LAUNCH_WALKC(STATS, ONE)                                                            \
  do                                                                                            \
    {                                                                                           \
      if (cfg.walk_carveout >= 0)                                                               \
        cudaFuncSetAttribute(k_walkc<SRC, STATS, ONE>, cudaFuncAttributePreferredSharedMemoryCarveout, cfg.walk_carveout); \
      k_walkc<SRC, STATS, ONE><<<gc, GSI_WALKC_THREADS, 0, st>>>(V, CM, src, nq, (unsigned) chunk, val, leaf_out, \
                                                                ld, cfg.list, list_cap, cfg.ovf_mask,      \
                                                                cfg.counters, cfg.max_steps, cfg.walk_dynamic, \
                                                                l1_alloc, trace);                          \
    }                                                                                           \
  while (0)
          if (cfg.stats && V.n_fields == 1) GSI_LAUNCH_WALKC(true, true);
          else if (cfg.stats) GSI_LAUNCH_WALKC(true, false);
          else if (V.n_fields == 1) GSI_LAUNCH_WALKC(false, true);
          else GSI_LAUNCH_WALKC(false, false);
#undef GSI_LAUNCH_WALKC
        }
      else if (cfg.walk_variant == 3 && V.leaf_stride == 128)
        {
          const int ctas_f = cfg.sm_count * (cfg.walk_ctas_per_sm > 5 ? cfg.walk_ctas_per_sm : 8);
          const int gf = (int) ((n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32) < (size_t) ctas_f
                                    ? (n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32)
                                    : (size_t) ctas_f);
          gsi_cell_map CM;
          CM.ax = V.xf.sx * V.inv_du;
          CM.bx = (-V.xf.cx * V.xf.sx - V.u0) * V.inv_du;
          CM.ay = V.xf.sy * V.inv_dv;
          CM.by = (-V.xf.cy * V.xf.sy - V.v0) * V.inv_dv;
#define GSI_LAUNCH_WALKF(STATS, ONE)                                                            \
  do                                                                                            \
    {                                                                                           \
      cudaFuncSetAttribute(k_walkf<SRC, STATS, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GSI_WALKF_SMEM); \
      if (cfg.walk_carveout >= 0)                                                               \
        cudaFuncSetAttribute(k_walkf<SRC, STATS, ONE>, cudaFuncAttributePreferredSharedMemoryCarveout, cfg.walk_carveout); \
      k_walkf<SRC, STATS, ONE><<<gf, GSI_WALK_THREADS, GSI_WALKF_SMEM, st>>>(V, CM, src, nq, (unsigned) chunk, val, \
                                                                            leaf_out, ld, cfg.list, list_cap,     \
                                                                            cfg.ovf_mask, cfg.counters,           \
                                                                            cfg.max_steps, cfg.walk_dynamic);     \
    }                                                                                           \
  while (0)
          if (cfg.stats && V.n_fields == 1) GSI_LAUNCH_WALKF(true, true);
          else if (cfg.stats) GSI_LAUNCH_WALKF(true, false);
          else if (V.n_fields == 1) GSI_LAUNCH_WALKF(false, true);
          else GSI_LAUNCH_WALKF(false, false);
#undef GSI_LAUNCH_WALKF
        }
      else if (cfg.walk_variant == 1)
        {
          unsigned long long *trace = (size_t) g <= cfg.walk_trace_ctas ? cfg.walk_trace : nullptr;
          if (trace) cudaMemsetAsync(trace, 0, cfg.walk_trace_ctas * 4 * sizeof(unsigned long long), st);
#define GSI_LAUNCH_WALKP(STATS, ONE)                                                            \
  k_walkp<SRC, STATS, ONE><<<g, GSI_WALK_THREADS, 0, st>>>(V, src, nq, (unsigned) chunk, val, leaf_out, \
                                                           ld, cfg.list, list_cap, cfg.ovf_mask,      \
                                                           cfg.counters, cfg.max_steps, cfg.walk_dynamic, cfg.walk_pf, trace)
          if (cfg.stats && V.n_fields == 1) GSI_LAUNCH_WALKP(true, true);
          else if (cfg.stats) GSI_LAUNCH_WALKP(true, false);
          else if (V.n_fields == 1) GSI_LAUNCH_WALKP(false, true);
          else GSI_LAUNCH_WALKP(false, false);
#undef GSI_LAUNCH_WALKP
        }
      else
        {
#define GSI_LAUNCH_WALK(STATS, ONE)                                                             \
  k_walk<SRC, STATS, ONE><<<g, GSI_WALK_THREADS, 0, st>>>(V, src, nq, (unsigned) chunk, val, leaf_out,  \
                                                          ld, cfg.list, list_cap, cfg.ovf_mask,      \
                                                          cfg.counters, cfg.max_steps)
      if (cfg.stats && V.n_fields == 1) GSI_LAUNCH_WALK(true, true);
      else if (cfg.stats) GSI_LAUNCH_WALK(true, false);
      else if (V.n_fields == 1) GSI_LAUNCH_WALK(false, true);
      else GSI_LAUNCH_WALK(false, false);
#undef GSI_LAUNCH_WALK
        }
    }
  else if (cfg.stats)
    k_locate<SRC, true><<<grid_for(src.n_launch(nq), GSI_LOCATE_THREADS, cfg.sm_count, cfg.locate_ctas_per_sm),
                          GSI_LOCATE_THREADS, 0, st>>>(V, src, nq, val, leaf_out, ld, cfg.list, list_cap,
                                                       cfg.ovf_mask, cfg.counters, cfg.max_steps);
  else
    k_locate<SRC, false><<<grid_for(src.n_launch(nq), GSI_LOCATE_THREADS, cfg.sm_count, cfg.locate_ctas_per_sm),
                           GSI_LOCATE_THREADS, 0, st>>>(V, src, nq, val, leaf_out, ld, cfg.list,
                                                        list_cap, cfg.ovf_mask, cfg.counters,
                                                        cfg.max_steps);
  if (cfg.ev_locate1) cudaEventRecord(cfg.ev_locate1, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (cfg.ev_exact0) cudaEventRecord(cfg.ev_exact0, st);
  const unsigned *rlist = cfg.list;
  int cnt_word = GSI_CNT_DEFERRED;
  if (V.memo && cfg.list2)
    {
      k_memo_list<SRC><<<cfg.sm_count * 8, 128, 0, st>>>(V, src, cfg.list, list_cap, cfg.list2, val, leaf_out, ld, cfg.counters);
      rlist = cfg.list2;
      cnt_word = GSI_CNT_DEFERRED2;
    }
  if (cfg.exact_coop == 2)
    k_exact_list4p<SRC><<<cfg.sm_count * GSI_EXACT4P_CTAS, 128, 0, st>>>(V, src, rlist, list_cap, cnt_word, val, leaf_out, ld, cfg.counters);
  else if (cfg.exact_coop)
    k_exact_list4<SRC><<<cfg.sm_count * 8, 128, 0, st>>>(V, src, rlist, list_cap, cnt_word, val, leaf_out, ld, cfg.counters);
  else
    k_exact_list<SRC><<<cfg.sm_count * 4, 128, 0, st>>>(V, src, rlist, list_cap, cnt_word, val, leaf_out, ld,
                                                        cfg.counters);
  if (cfg.exact_coop)
    k_exact_overflow4<SRC><<<cfg.sm_count * 8, 128, 0, st>>>(V, src, nq, cfg.ovf_mask, val, leaf_out, ld, cfg.counters);
  else
    k_exact_overflow<SRC><<<cfg.sm_count * 4, 128, 0, st>>>(V, src, nq, cfg.ovf_mask, val, leaf_out, ld,
                                                            cfg.counters);
  if (cfg.ev_exact1) cudaEventRecord(cfg.ev_exact1, st);
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
                       unsigned long long ny, unsigned long long i_first, unsigned long long n_cols,
                       double *val, int *leaf_out, size_t ld, const gsi_launch_cfg &cfg,
                       cudaStream_t st)
{
  return launch_eval(V, src_raster{x0, dx, y0, dy, ny, i_first, n_cols}, (size_t) (n_cols * ny), val,
                     leaf_out, ld, cfg, st);
}

cudaError_t
gsi_launch_seed_cells(const gsi_view &V, int32_t *seeds, int sm_count, cudaStream_t st)
{
  const size_t ncell = (size_t) V.grid_g * V.grid_g;
  k_seed_cells<<<grid_for(ncell, 128, sm_count, 16), 128, 0, st>>>(V, seeds);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_pack_seeds(const gsi_view &V, const int32_t *seeds, uint16_t *seed16, int32_t *seedbase, int sm_count,
                      cudaStream_t st)
{
  const size_t nblk = (size_t) V.grid_gb * V.grid_gb; /* one warp per block */
  k_pack_seeds<<<grid_for(nblk * 32, 128, sm_count, 16), 128, 0, st>>>(V, seeds, seed16, seedbase);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_entry_cells(const gsi_view &V, int32_t *entry, int sm_count, cudaStream_t st)
{
  const size_t ncell = (size_t) V.grid_g * V.grid_g;
  k_entry_cells<<<grid_for(ncell, 128, sm_count, 16), 128, 0, st>>>(V, entry);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_memo_fill(const gsi_view &V, gsi_memo *memo, int sm_count, cudaStream_t st)
{
  if (!V.memo) return cudaSuccess;
  k_memo_fill<<<grid_for((size_t) V.memo_mask + 1, 128, sm_count, 16), 128, 0, st>>>(V, memo);
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

cudaError_t
gsi_launch_validate_slab(const gsi_view &V, unsigned long long *bad, int sm_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(bad, 0, sizeof *bad, st);
  if (e != cudaSuccess) return e;
  const unsigned long long n_points = V.n_points;
  k_validate_slab<<<grid_for((size_t) V.n_dag, 256, sm_count, 8), 256, 0, st>>>(V, n_points, bad);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_deriv(const gsi_view &V, const int *leaf, size_t nq, double *dzdx, double *dzdy, size_t ld, int sm_count,
                 cudaStream_t st)
{
  if (nq == 0) return cudaSuccess;
  k_deriv<<<grid_for(nq, 256, sm_count, 8), 256, 0, st>>>(V, leaf, nq, dzdx, dzdy, ld);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_leaf_rcp(const gsi_view &V, char *leaf_base, int sm_count, cudaStream_t st)
{
  if (V.leaf_stride != 128 || V.n_leaves <= 0) return cudaSuccess;
  k_leaf_rcp<<<grid_for((size_t) V.n_leaves, 256, sm_count, 8), 256, 0, st>>>(V, leaf_base);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_build_locate(const gsi_rec *rec, const int32_t *lnk, const uint8_t *type, const gsi_xform &xf,
                        const double *px, const double *py, size_t first, size_t count, int32_t *node, double *c0,
                        double *c1, int32_t *last, int sm_count, cudaStream_t st)
{
  if (count == 0) return cudaSuccess;
  const gsi_build_view B{rec, lnk, type};
  k_build_locate<<<grid_for(count, 128, sm_count, 16), 128, 0, st>>>(B, xf, px, py, first, count, node, c0, c1, last);
  return cudaGetLastError();
}

cudaError_t
gsi_launch_selftest_division(uint64_t seed, size_t n, unsigned long long *out2, int sm_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), st);
  if (e != cudaSuccess || n == 0) return e;
  k_selftest_division<<<grid_for(n, 256, sm_count, 8), 256, 0, st>>>(seed, n, out2);
  return cudaGetLastError();
}
