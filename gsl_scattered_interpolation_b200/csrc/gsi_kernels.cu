/* gsi_kernels.cu -- sm_100a kernels of libgsl_scattered_cuda.
 *
 *   k_walk         fast path for query arrays (default): a warp consumes a chunk of queries through
 *                  a warp-wide work queue, one walk step per iteration for whatever query each lane
 *                  holds.  Seed from the uniform grid, cross the edge with the most negative
 *                  barycentric weight, accept the leaf only when all three weights clear the leaf's
 *                  safety margin, then blend.  The weights are the reference's own
 *                  (calculate_bary_coords, linear_simplex.c:607-651), so an accepted result is
 *                  bit-identical to find_leaf + interp_point; everything else is deferred.
 *   k_locate       the same per-query logic, one thread per query over compact pixel tiles: the
 *                  default for rasters (coherent queries).
 *   k_exact_*      exact replay of find_leaf/_find_leaf (linear_simplex.c:331-402) on the flattened
 *                  history DAG + interp_point (linear_simplex.c:678-711): the reference's
 *                  edge/vertex tie-breaks, cage leaves, stale-coords quirk.  Runs on the deferred
 *                  list, on grid-cell centres (seeding), or on every query; starts at the cell's
 *                  certified entry node (k_entry_cells).
 *   k_bin -> k_tile_locate -> k_unpermute
 *                  opt-in reorder pipeline for incoherent query arrays (chunk-local counting sort
 *                  by spatial tile, tile-major walk, inverse permutation).
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

/* exact replay on the deferred list written by k_locate */
template <class SRC>
__global__ void __launch_bounds__(128, 4)
k_exact_list(const gsi_view V, const SRC src, const unsigned *list, unsigned list_cap, double *val,
             int *leaf_out, size_t ld, unsigned *counters)
{
  const unsigned count = min(counters[GSI_CNT_DEFERRED], list_cap);
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
          if (gsi_exact_one(&V, qx, qy, val, leaf_out, ld, q) < 0)
            atomicAdd(counters + GSI_CNT_OUTSIDE, 1u);
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
              slot = gsi_ld_seed(V.seeds + gsi_cell_of(&V, qx, qy));
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
              if (nxt_v) nxt_seed = gsi_ld_seed(V.seeds + gsi_cell_of(&V, nxt_x, nxt_y));
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

/* ---- reorder pipeline: k_bin -> k_tile_locate -> k_unpermute ---------------------------------
 * Uniformly random queries make every leaf fetch an isolated L2/DRAM access.  The pipeline
 * gives the walk spatial coherence without a global sort and without scattered global writes:
 *   k_bin          one CTA per chunk of <= 8192 consecutive queries: counting sort by spatial
 *                  tile (Morton-key prefix of the query position: the leaf order's own key, so a
 *                  tile's leaves are one contiguous slot range) entirely in shared memory,
 *                  then coalesced stores of the binned coordinates, the 16-bit in-chunk index
 *                  of every binned query and the chunk's run table (start of each tile's run);
 *   k_tile_locate  tile-major: the CTAs of tile t gather t's run from every chunk (32 chunks
 *                  per warp step, lane-parallel prefix over the run lengths) and walk them, so
 *                  the tile's ~1000 leaves and its seed cells stay in that SM's L1; results go
 *                  to the binned slots (neighbouring tiles are resident together, so the runs'
 *                  shared sectors merge in L2);
 *   k_unpermute    one CTA per chunk: the inverse permutation through shared memory, coalesced
 *                  loads and stores.
 * The arithmetic per query is the one of k_locate (gsi_walk_step), so accepted results are
 * bit-identical; deferred queries go to the same list (original query index) and are replayed
 * after k_unpermute. */
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

__device__ __forceinline__ unsigned
compact16(unsigned x) /* inverse of spread16 */
{
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0f0f0f0fu;
  x = (x | (x >> 4)) & 0x00ff00ffu;
  x = (x | (x >> 8)) & 0x0000ffffu;
  return x;
}

/* Tile of a query = the tile_bits-bit prefix of the 32-bit Morton key of its position in the data
 * bounding box (16 bits per axis, y on the odd bits) -- the key gsi_flatten.cpp sorts the leaves'
 * centroids by, so tile t's leaves are the slots [V.tiles[t], V.tiles[t+1]).  Non-finite and
 * outside coordinates are clamped (they are deferred by the walk anyway). */
__device__ __forceinline__ unsigned
tile_of(const gsi_view &V, double qx, double qy)
{
  const double u = (qx - V.xf.cx) * V.xf.sx, v = (qy - V.xf.cy) * V.xf.sy;
  const double k = 65536.0 / (double) V.grid_g; /* (u - u0) * inv_du is in grid cells */
  const unsigned ix = (unsigned) __double2int_rd(fmin(fmax((u - V.u0) * V.inv_du * k, 0.0), 65535.0));
  const unsigned iy = (unsigned) __double2int_rd(fmin(fmax((v - V.v0) * V.inv_dv * k, 0.0), 65535.0));
  return (spread16(ix) | (spread16(iy) << 1)) >> (32 - V.tile_bits);
}

/* One coordinate is staged at a time (96 KB of shared memory for 8192-query chunks) so that two
 * CTAs share an SM and one CTA's global loads overlap the other's scan / copy-out. */
__global__ void __launch_bounds__(GSI_BIN_THREADS, 2)
k_bin(const gsi_view V, const double *__restrict__ xq, const double *__restrict__ yq, size_t nq, int S,
       int nt, double *__restrict__ bx, double *__restrict__ by,
       unsigned short *__restrict__ bi, unsigned short *__restrict__ off)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
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
  const size_t ntp = (size_t) nt + 2;
  const unsigned per = ((unsigned) nt + T - 1) / T;
  for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x)
    {
      const size_t q0 = c * (size_t) S;
      const unsigned n = (unsigned) (nq - q0 < (size_t) S ? nq - q0 : (size_t) S);
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
              const unsigned t = tile_of(V, __ldg(xq + q0 + i), __ldg(yq + q0 + i));
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
      unsigned short *row = off + c * ntp;
#pragma unroll
      for (unsigned j = 0; j < PER; j++)
        if (j < per && b + j < (unsigned) nt)
          {
            hist[b + j] = run;
            row[b + j] = (unsigned short) run;
            run += loc[j];
          }
      if (tid == 0) row[nt] = (unsigned short) n;
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

/* one converged warp step of the tile kernels: every live lane walks the binned query at slot
 * `pos` of the window (chunk base `cbase`) and stores its result to the binned outputs;
 * deferred queries are appended to the list under their original index */
template <bool STATS>
__device__ __forceinline__ void
tile_direct_step(const gsi_view &V, const double *__restrict__ bx, const double *__restrict__ by,
                 const unsigned short *__restrict__ bi, bool live, size_t pos, size_t cbase, double *zb,
                 int *lb, size_t ldb, unsigned *list, unsigned list_cap, unsigned *ovf_mask,
                 unsigned *counters, int max_steps, unsigned long long &steps_acc)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  int status = GSI_STEP_MOVE; /* = not a live query */
  int slot = 0;
  double c0 = 0, c1 = 0;
  if (live)
    {
      const double qx = __ldcs(bx + pos), qy = __ldcs(by + pos);
      status = GSI_STEP_DEFER;
      if (isfinite(qx) && isfinite(qy))
        {
          slot = gsi_ld_seed(V.seeds + gsi_cell_of(&V, qx, qy));
          for (int step = 0; step < max_steps; step++)
            {
              gsi_rec rec;
              const int at = slot;
              gsi_fetch_rec(&rec, gsi_leaf_at(&V, at));
              const int st = gsi_walk_step(&V.xf, &rec, qx, qy, &slot, &c0, &c1);
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
  __syncwarp();
  if (status == GSI_STEP_ACCEPT)
    {
      gsi_leaf_tail tl;
      gsi_fetch_tail(&tl, gsi_leaf_at(&V, slot));
      if (lb) lb[pos] = tl.g.node;
      if (zb) gsi_blend_store(&V, &tl, c0, c1, zb, ldb, pos);
    }
  const bool defer = status == GSI_STEP_DEFER;
  const unsigned dm = __ballot_sync(FULL, defer);
  if (dm)
    {
      const int leader = __ffs(dm) - 1;
      unsigned at0 = 0;
      if (lane == (unsigned) leader) at0 = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
      at0 = __shfl_sync(FULL, at0, leader);
      if (defer)
        {
          const size_t q = cbase + bi[pos];
          const unsigned at = at0 + __popc(dm & ((1u << lane) - 1u));
          if (at < list_cap)
            list[at] = (unsigned) q;
          else
            mark_overflow(ovf_mask, q, counters);
        }
    }
}

/* lane <-> chunk: load tile t's run in chunk (cb*32 + lane) and the inclusive prefix of the run
 * lengths over the warp; returns the warp's total */
__device__ __forceinline__ unsigned
tile_runs(const unsigned short *__restrict__ off, size_t ntp, unsigned t, unsigned cb, unsigned n_blocks,
          unsigned n_chunks, unsigned &o0, unsigned &len, unsigned &incl)
{
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned c = (cb << 5) + lane;
  o0 = 0;
  len = 0;
  if (cb < n_blocks && c < n_chunks)
    {
      const unsigned short *row = off + (size_t) c * ntp + t;
      o0 = row[0];
      len = (unsigned) row[1] - o0;
    }
  incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned v = __shfl_up_sync(FULL, incl, o);
      if (lane >= (unsigned) o) incl += v;
    }
  return __shfl_sync(FULL, incl, 31);
}

/* item k (< total) of the warp's concatenated runs -> binned slot of the window; *cbase = first
 * slot of the item's chunk.  Must be called by the whole warp. */
__device__ __forceinline__ size_t
tile_item(unsigned k, unsigned cb, int S, unsigned o0, unsigned len, unsigned incl, size_t *cbase)
{
  const unsigned FULL = 0xffffffffu;
  unsigned j = 0; /* number of chunks whose inclusive prefix is <= k = the chunk holding item k */
#pragma unroll
  for (unsigned step = 16; step; step >>= 1)
    {
      const unsigned v = __shfl_sync(FULL, incl, j + step - 1);
      if (v <= k) j += step;
    }
  const unsigned incl_j = __shfl_sync(FULL, incl, j), len_j = __shfl_sync(FULL, len, j),
                 o0_j = __shfl_sync(FULL, o0, j);
  *cbase = (size_t) ((cb << 5) + j) * (size_t) S;
  return *cbase + o0_j + (k - (incl_j - len_j));
}

template <bool STATS>
__global__ void __launch_bounds__(GSI_TILE_THREADS, 2)
k_tile_locate(const gsi_view V, const double *__restrict__ bx, const double *__restrict__ by,
              const unsigned short *__restrict__ bi, const unsigned short *__restrict__ off,
              unsigned n_chunks, int S, int nt, int ranges, double *zb, int *lb, size_t ldb,
              unsigned *list, unsigned list_cap, unsigned *ovf_mask, unsigned *counters, int max_steps)
{
  const unsigned t = blockIdx.x / (unsigned) ranges, r = blockIdx.x - t * (unsigned) ranges;
  constexpr unsigned WPC = GSI_TILE_THREADS / 32;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned n_blocks = (n_chunks + 31u) >> 5;
  const size_t ntp = (size_t) nt + 2;
  unsigned long long steps_acc = 0;
  for (unsigned cb = r * WPC + warp; cb < n_blocks; cb += (unsigned) ranges * WPC)
    {
      unsigned o0, len, incl;
      const unsigned total = tile_runs(off, ntp, t, cb, n_blocks, n_chunks, o0, len, incl);
      for (unsigned base = 0; base < total; base += 32)
        {
          const bool live = base + lane < total;
          size_t cbase;
          const size_t pos = tile_item(live ? base + lane : total - 1, cb, S, o0, len, incl, &cbase);
          tile_direct_step<STATS>(V, bx, by, bi, live, pos, cbase, zb, lb, ldb, list, list_cap, ovf_mask,
                                  counters, max_steps, steps_acc);
        }
    }
  if (STATS && steps_acc)
    atomicAdd(reinterpret_cast<unsigned long long *>(counters + GSI_CNT_STEPS), steps_acc);
}

/* ---- tile kernel with the tile staged in shared memory ---------------------------------------------
 * k_tile_locate gathers every leaf sector through L1, and a divergent L1 access costs the same
 * L1 cycles whether it hits or not (0.77 sectors/clk/SM measured).  Here the CTA first copies its
 * tile -- the contiguous slot range [V.tiles[t], V.tiles[t+1]) of 96-byte leaf records and the
 * block of seed cells under the tile's rectangle -- into shared memory with coalesced loads, and
 * the walk reads seeds, records and tails from there; only leaves outside the tile (walks that
 * leave it, tiles larger than the staging area) and the query coordinates come from global
 * memory.  Same per-query operations as every other fast-path kernel. */
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

template <bool STATS>
__global__ void __launch_bounds__(GSI_TILE_THREADS, 2)
k_tile_locate_smem(const gsi_view V, const double *__restrict__ bx, const double *__restrict__ by,
                   const unsigned short *__restrict__ bi, const unsigned short *__restrict__ off,
                   unsigned n_chunks, int S, int nt, int ranges, int leaf_cap, int seed_rows, int seed_pitch,
                   double *zb, int *lb, size_t ldb, unsigned *list, unsigned list_cap, unsigned *ovf_mask,
                   unsigned *counters, int max_steps)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4 *sleaf = reinterpret_cast<uint4 *>(smem_raw);
  int *sseed = reinterpret_cast<int *>(sleaf + (size_t) leaf_cap * 6);
  const unsigned FULL = 0xffffffffu;
  const unsigned t = blockIdx.x / (unsigned) ranges, r = blockIdx.x - t * (unsigned) ranges;
  constexpr unsigned WPC = GSI_TILE_THREADS / 32;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned n_blocks = (n_chunks + 31u) >> 5;
  const size_t ntp = (size_t) nt + 2;

  /* ---- stage the tile ---- */
  const int slot0 = V.tiles[t];
  const int n_staged = min(V.tiles[t + 1] - slot0, leaf_cap);
  if (V.leaf_stride == (int) sizeof(gsi_leaf))
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(V.leaf + (size_t) slot0 * sizeof(gsi_leaf));
      for (int i = threadIdx.x; i < n_staged * 6; i += GSI_TILE_THREADS) sleaf[i] = __ldg(src + i);
    }
  else
    for (int i = threadIdx.x; i < n_staged * 6; i += GSI_TILE_THREADS)
      sleaf[i] = __ldg(reinterpret_cast<const uint4 *>(V.leaf + (size_t) (slot0 + i / 6) * (size_t) V.leaf_stride) + i % 6);
  /* the tile's rectangle in 16-bit Morton coordinates -> block of seed cells */
  const int by_bits = (V.tile_bits + 1) >> 1, bx_bits = V.tile_bits >> 1;
  const unsigned key_hi = t << (32 - V.tile_bits);
  const unsigned qx_lo = compact16(key_hi), qy_lo = compact16(key_hi >> 1);
  const double cells_per_q = (double) V.grid_g / 65536.0;
  const int cx0 = max(0, (int) floor((double) qx_lo * cells_per_q) - 1);
  const int cy0 = max(0, (int) floor((double) qy_lo * cells_per_q) - 1);
  const int ncx = min(seed_pitch, V.grid_g - cx0), ncy = min(seed_rows, V.grid_g - cy0);
  (void) bx_bits;
  (void) by_bits;
  for (int i = threadIdx.x; i < ncy * seed_pitch; i += GSI_TILE_THREADS)
    {
      const int ry = i / seed_pitch, rx = i - ry * seed_pitch;
      if (rx < ncx) sseed[i] = __ldg(V.seeds + (size_t) (cy0 + ry) * (size_t) V.grid_g + (size_t) (cx0 + rx));
    }
  __syncthreads();

  const double gmax = (double) (V.grid_g - 1);
  unsigned long long steps_acc = 0;
  for (unsigned cb = r * WPC + warp; cb < n_blocks; cb += (unsigned) ranges * WPC)
    {
      unsigned o0, len, incl;
      const unsigned total = tile_runs(off, ntp, t, cb, n_blocks, n_chunks, o0, len, incl);
      for (unsigned base = 0; base < total; base += 32)
        {
          const bool live = base + lane < total;
          size_t cbase;
          const size_t pos = tile_item(live ? base + lane : total - 1, cb, S, o0, len, incl, &cbase);
          int status = GSI_STEP_MOVE; /* = not a live query */
          int slot = 0;
          double c0 = 0, c1 = 0;
          if (live)
            {
              const double qx = __ldcs(bx + pos), qy = __ldcs(by + pos);
              status = GSI_STEP_DEFER;
              if (isfinite(qx) && isfinite(qy))
                {
                  /* the cell of gsi_cell_of; its seed from the staged block when it is inside */
                  const double u = (qx - V.xf.cx) * V.xf.sx, v = (qy - V.xf.cy) * V.xf.sy;
                  const int ix = __double2int_rd(fmin(fmax((u - V.u0) * V.inv_du, 0.0), gmax));
                  const int iy = __double2int_rd(fmin(fmax((v - V.v0) * V.inv_dv, 0.0), gmax));
                  const unsigned rx = (unsigned) (ix - cx0), ry = (unsigned) (iy - cy0);
                  slot = (rx < (unsigned) ncx && ry < (unsigned) ncy)
                             ? sseed[ry * seed_pitch + rx]
                             : gsi_ld_seed(V.seeds + (size_t) iy * (size_t) V.grid_g + (size_t) ix);
                  for (int step = 0; step < max_steps; step++)
                    {
                      gsi_rec rec;
                      const int at = slot;
                      const unsigned rel = (unsigned) (at - slot0);
                      if (rel < (unsigned) n_staged)
                        fetch_rec_smem(&rec, sleaf + (size_t) rel * 6);
                      else
                        gsi_fetch_rec(&rec, gsi_leaf_at(&V, at));
                      const int st = gsi_walk_step(&V.xf, &rec, qx, qy, &slot, &c0, &c1);
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
          __syncwarp();
          if (status == GSI_STEP_ACCEPT)
            {
              gsi_leaf_tail tl;
              const unsigned rel = (unsigned) (slot - slot0);
              if (rel < (unsigned) n_staged)
                fetch_tail_smem(&tl, sleaf + (size_t) rel * 6);
              else
                gsi_fetch_tail(&tl, gsi_leaf_at(&V, slot));
              if (lb) lb[pos] = tl.g.node;
              if (zb) gsi_blend_store(&V, &tl, c0, c1, zb, ldb, pos);
            }
          const bool defer = status == GSI_STEP_DEFER;
          const unsigned dm = __ballot_sync(FULL, defer);
          if (dm)
            {
              const int leader = __ffs(dm) - 1;
              unsigned at0 = 0;
              if (lane == (unsigned) leader) at0 = atomicAdd(counters + GSI_CNT_DEFERRED, (unsigned) __popc(dm));
              at0 = __shfl_sync(FULL, at0, leader);
              if (defer)
                {
                  const size_t q = cbase + bi[pos];
                  const unsigned at = at0 + __popc(dm & lt_mask);
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

__global__ void __launch_bounds__(GSI_UNPERMUTE_THREADS, 2)
k_unpermute(const double *__restrict__ zb, const int *__restrict__ lb, const unsigned short *__restrict__ bi,
            size_t nq, int S, int n_fields, size_t ldb, double *__restrict__ val, size_t ld,
            int *__restrict__ leaf_out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
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
  counters[GSI_CNT_OVERFLOW] = 0;
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
  const int S = R->chunk, nt = 1 << V.tile_bits;
  const size_t n_chunks = (nq + (size_t) S - 1) / (size_t) S;
  const size_t smem_bin = (size_t) S * 10 + (size_t) nt * 4, smem_unp = (size_t) S * 12;
  /* staging area of the tile kernel: the largest tile's records (capped) + the seed cells under a
     tile's rectangle (x gets floor(P/2) bits of the prefix, y the rest), two cells of margin */
  const int bx_bits = V.tile_bits >> 1, by_bits = (V.tile_bits + 1) >> 1;
  const int seed_pitch = (V.grid_g >> bx_bits) + 4, seed_rows = (V.grid_g >> by_bits) + 4;
  const size_t seed_bytes = (size_t) seed_pitch * seed_rows * 4;
  int leaf_cap = R->tile_max_leaves < 1 ? 1 : R->tile_max_leaves;
  const size_t budget = 110 * 1024; /* two CTAs per SM */
  if (seed_bytes + 96 > budget) leaf_cap = 0;
  else if ((size_t) leaf_cap * 96 + seed_bytes > budget) leaf_cap = (int) ((budget - seed_bytes) / 96);
  const size_t smem_tile = (size_t) leaf_cap * 96 + seed_bytes;
  /* per device (the current one), so not cached in a process-wide flag */
  cudaFuncSetAttribute(k_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, GSI_BIN_CHUNK_MAX * 10 + 4096 * 4);
  cudaFuncSetAttribute(k_unpermute, cudaFuncAttributeMaxDynamicSharedMemorySize, GSI_BIN_CHUNK_MAX * 12);
  const int g_bin = (int) (n_chunks < (size_t) (2 * cfg.sm_count) ? n_chunks : (size_t) (2 * cfg.sm_count));
  k_bin<<<g_bin, GSI_BIN_THREADS, smem_bin, st>>>(V, src.x, src.y, nq, S, nt, R->bx, R->by, R->bi, R->off);
  /* CTAs per tile: every warp of a CTA should get a few 32-chunk blocks, else one CTA per tile */
  const size_t n_blocks = (n_chunks + 31) / 32;
  int ranges = (int) (n_blocks / (4 * (GSI_TILE_THREADS / 32)));
  if (ranges > R->ranges) ranges = R->ranges;
  if (ranges < 1) ranges = 1;
  double *zb = val ? R->zb : nullptr;
  int *lb = leaf_out ? R->lb : nullptr;
  if (R->tile_smem && leaf_cap > 0)
    {
      cudaFuncSetAttribute(k_tile_locate_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget);
      cudaFuncSetAttribute(k_tile_locate_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget);
      if (cfg.stats)
        k_tile_locate_smem<true><<<nt * ranges, GSI_TILE_THREADS, smem_tile, st>>>(
            V, R->bx, R->by, R->bi, R->off, (unsigned) n_chunks, S, nt, ranges, leaf_cap, seed_rows, seed_pitch, zb,
            lb, R->cap, cfg.list, list_cap, cfg.ovf_mask, cfg.counters, cfg.max_steps);
      else
        k_tile_locate_smem<false><<<nt * ranges, GSI_TILE_THREADS, smem_tile, st>>>(
            V, R->bx, R->by, R->bi, R->off, (unsigned) n_chunks, S, nt, ranges, leaf_cap, seed_rows, seed_pitch, zb,
            lb, R->cap, cfg.list, list_cap, cfg.ovf_mask, cfg.counters, cfg.max_steps);
    }
  else if (cfg.stats)
    k_tile_locate<true><<<nt * ranges, GSI_TILE_THREADS, 0, st>>>(
        V, R->bx, R->by, R->bi, R->off, (unsigned) n_chunks, S, nt, ranges, zb, lb, R->cap, cfg.list, list_cap,
        cfg.ovf_mask, cfg.counters, cfg.max_steps);
  else
    k_tile_locate<false><<<nt * ranges, GSI_TILE_THREADS, 0, st>>>(
        V, R->bx, R->by, R->bi, R->off, (unsigned) n_chunks, S, nt, ranges, zb, lb, R->cap, cfg.list, list_cap,
        cfg.ovf_mask, cfg.counters, cfg.max_steps);
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
      /* per-warp chunk: ~8 chunks per warp for balance, 256..8192 queries, multiple of 32 */
      size_t chunk = (nq / (n_warps * 8) + 31) & ~(size_t) 31;
      if (chunk < 256) chunk = 256;
      if (chunk > 8192) chunk = 8192;
      if (cfg.walk_chunk) chunk = cfg.walk_chunk;
      const size_t n_chunks = (nq + chunk - 1) / chunk;
      const int g = (int) ((n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32) < (size_t) ctas
                               ? (n_chunks + GSI_WALK_THREADS / 32 - 1) / (GSI_WALK_THREADS / 32)
                               : (size_t) ctas);
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
  k_exact_list<SRC><<<cfg.sm_count * 4, 128, 0, st>>>(V, src, cfg.list, list_cap, val, leaf_out, ld,
                                                      cfg.counters);
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
gsi_launch_entry_cells(const gsi_view &V, int32_t *entry, int sm_count, cudaStream_t st)
{
  const size_t ncell = (size_t) V.grid_g * V.grid_g;
  k_entry_cells<<<grid_for(ncell, 128, sm_count, 16), 128, 0, st>>>(V, entry);
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
