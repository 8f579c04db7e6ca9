/* gsi_layout.h -- the device triangulation layout ("slab") of libgsl_scattered_cuda.
 *
 * One contiguous, position-independent allocation per GPU (offsets only, so it can be
 * replicated with a plain peer copy / NCCL broadcast).  Sections, all 256-byte aligned:
 *
 *   header   gsi_slab_header
 *   dag      gsi_rec[n_dag]     history-DAG nodes in the reference's node-id order (exact
 *                               replay kernel; a leaf's lnk[0] holds its compact slot)
 *   leaf     gsi_leaf[n_leaves] leaves only (96 B), sorted along a Morton curve of their
 *                               centroids; r.lnk[] = slot of the neighbour opposite vertex i,
 *                               -1 = none (hull of the cage); r.meta = float margin | swap
 *                               sign bit; tail = reference node id + vertex slots (or, with
 *                               one field, the three vertex values themselves)
 *   z        double[n_points][n_fields]  vertex values, vertex-major so the F fields of a
 *                               vertex are contiguous (F = 16 -> one 128-byte line, four
 *                               256-bit loads), vertices sorted along the same Morton curve
 *   seeds    int32[G*G]         uniform grid over the data bounding box in standardised
 *                               coordinates: leaf slot containing the cell centre
 *                               (filled on the device by the exact-replay kernel)
 *   seed16   uint16[G*G] +      the same table as the fast path reads it: seed = base[block] + seed16[cell]
 *   seedbase int32[GB*GB]       with one base (the smallest seed) per block of B x B cells, B = 2^seed_shift,
 *                               GB = ceil(G/B) <= 32.  Leaves are in Morton order and the grid covers the
 *                               box the Morton keys quantise, so the seeds of a block are a narrow slot
 *                               range; 0xffff = escape to the int32 table (blocks that straddle a
 *                               coarse Morton boundary, dense clusters).  2 bytes per cell instead of 4:
 *                               a table that stays L2-resident beside ~250 MB of randomly gathered
 *                               leaf records (16 MB instead of 32 MB at 10^6 nodes).  The base table is
 *                               at most 4 KB, so the walk kernel keeps it in shared memory: a query's
 *                               seed costs ONE divergent global access (seed16), not two.  The
 *                               flattener rounds G to a multiple of 8 B so that the coarse Morton
 *                               boundaries (multiples of 1/8 of the box) fall between blocks.
 *   tiles    int32[2^P + 1]     the leaves are sorted by the Morton key of their centroid, so the
 *                               leaves whose key starts with the P-bit prefix t are the slot
 *                               range [tiles[t], tiles[t+1]): a spatial tile whose records can be
 *                               staged in shared memory with one contiguous copy (reorder pipeline)
 *   memo     gsi_memo[2^m]      open-addressing hash table keyed by the BIT PATTERN of a data node's raw
 *                               coordinates: the exact replay's answer (leaf node id + slot) for a query
 *                               that coincides with that node, computed once at init by the exact replay
 *                               itself (k_memo_fill).  Such queries can never be accepted by the fast path
 *                               (two weights are within rounding of 0) and their leaf is history-dependent;
 *                               evaluating an interpolant AT its nodes is a common workload (residuals,
 *                               BASELINE.json configs[2]).  The deferred-query kernels probe it before they
 *                               descend the DAG; k_exact_all (the verification path) never does.
 *   entry    int32[G*G]         per grid cell, the deepest history-DAG node that the reference's
 *                               descent provably passes through for EVERY point of the cell
 *                               (0 = root): where the exact replay of a query of that cell
 *                               starts (filled on the device, gsi_entry_of_cell)
 *
 * Reference data model this flattens: simplex_tree / simplex_tree_node pools,
 * /root/reference/interpolation/linear_simplex.h:16-65 (POINT/LINK/LEAF/DATA_POINT).
 */
#ifndef GSI_LAYOUT_H
#define GSI_LAYOUT_H

#include <stdint.h>

#include "gsi_core.h"

#define GSI_SLAB_MAGIC 0x3142414c53495347ull /* "GSISLAB1" */
#define GSI_SLAB_VERSION 7u
#define GSI_SEED_GB_MAX 32 /* seed blocks per side: the base table is at most 32 x 32 x 4 = 4 KB */
#define GSI_SEED16_ESC 0xffffu

/* One compacted leaf: 96 bytes = three 32-byte sectors, each fetched with one 256-bit load
 * (LDG.E.256 on sm_100a).  A walk step reads r (sectors 0-1); only the accepted leaf's tail
 * (sector 2) is read.  With one field the three vertex values live in the tail itself, so an
 * accepted query costs no further gather. */
typedef union
{
  struct
  {
    int32_t node; /* reference simplex_index (history-DAG node id) */
    int32_t v[3]; /* vertex slots into z (vertex-major), -1 = cage seed vertex */
    double pad[2];
  } g; /* n_fields > 1 */
  struct
  {
    int32_t node;
    uint32_t vflags; /* bit i set: POINT(leaf,i) is a data vertex (not a cage seed) */
    double z[3];     /* values of POINT(leaf,0..2) */
  } s; /* n_fields == 1 */
} gsi_leaf_tail;

typedef struct
#if defined(__CUDACC__)
    __align__(32)
#endif
    gsi_leaf
{
  gsi_rec r; /* lnk[] = neighbour slot opposite vertex i (-1 = none), meta = tau | swap bit */
  gsi_leaf_tail t;
} gsi_leaf;

#define GSI_MEMO_EMPTY 0u   /* ends a probe sequence */
#define GSI_MEMO_FULL 1u    /* key and answer */
#define GSI_MEMO_PENDING 2u /* key written by the flattener, answer not yet filled in on the device */
#define GSI_MEMO_SKIP 3u    /* key without an answer (outside the cage / singular leaf): keep probing */
#define GSI_MEMO_MAX_PROBES 64

typedef struct
{
  uint32_t state;
  int32_t node; /* reference simplex_index of the leaf */
  double x, y;  /* raw coordinates of the data node (compared bit for bit) */
  int32_t slot; /* the leaf's compact slot */
  uint32_t pad;
} gsi_memo;

typedef struct
{
  uint64_t magic;
  uint32_t version;
  uint32_t n_fields;
  uint64_t n_points;
  uint64_t n_dag;
  uint64_t n_leaves;
  uint32_t grid_g;
  uint32_t leaf_stride; /* bytes between leaf records: 96 (packed) or 128 (one cache line each) */
  gsi_xform xf;
  /* grid mapping in standardised coordinates: cell = floor((u - u0) * inv_du) */
  double u0, v0, inv_du, inv_dv;
  double delta; /* absolute safety margin (standardised units) behind the per-leaf tau */
  uint64_t off_dag, off_leaf, off_z, off_seeds;
  uint64_t total_bytes;
  uint64_t reserved[1]; /* [0] = image checksum in files (gsl_scattered_cuda_fwrite) */
  uint64_t off_memo;    /* gsi_memo[2^memo_bits] (memo_bits = 0: no memo) */
  uint64_t off_seed16;  /* uint16[grid_g^2] */
  uint64_t off_seedbase; /* int32[ceil(grid_g / 2^seed_shift)^2] */
  uint32_t tile_bits;       /* tiles = the 2^tile_bits prefixes of the leaves' 32-bit Morton keys */
  uint32_t tile_max_leaves; /* largest tile */
  uint64_t off_tiles;       /* int32[2^tile_bits + 1]: first leaf slot of every tile */
  uint32_t seed_shift;      /* seed blocks are 2^seed_shift cells wide (>= 3) */
  uint32_t memo_bits;
  uint64_t off_entry;
} gsi_slab_header;

/* seed blocks: the smallest power-of-two width (>= 8 cells) that leaves at most GSI_SEED_GB_MAX per side */
static inline uint32_t
gsi_seed_shift_for(uint32_t grid_g)
{
  uint32_t sh = 3;
  while (((grid_g + (1u << sh) - 1u) >> sh) > GSI_SEED_GB_MAX) sh++;
  return sh;
}

static inline uint32_t
gsi_seed_blocks(uint32_t grid_g, uint32_t seed_shift)
{
  return (grid_g + (1u << seed_shift) - 1u) >> seed_shift;
}

/* what the kernels receive by value */
typedef struct
{
  gsi_xform xf;
  const gsi_rec *dag;
  const char *leaf; /* gsi_leaf records, leaf_stride bytes apart */
  const double *z;
  const int32_t *seeds;
  const uint16_t *seed16;   /* compact seeds: base[block] + seed16[cell], GSI_SEED16_ESC -> seeds[cell] */
  const int32_t *seedbase;
  const int32_t *entry; /* NULL: exact replay always starts at the root */
  const int32_t *tiles; /* first leaf slot per Morton-prefix tile, 2^tile_bits + 1 entries */
  const gsi_memo *memo; /* NULL: no vertex memo */
  uint32_t memo_mask;   /* entries - 1 */
  int32_t tile_bits;
  double u0, v0, inv_du, inv_dv;
  double delta;
  int32_t grid_g;
  int32_t grid_gb; /* seed blocks per side */
  int32_t seed_shift;
  int32_t n_fields;
  int32_t n_dag;
  int32_t n_leaves;
  int32_t leaf_stride;
  uint64_t n_points;
} gsi_view;

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
static inline
#endif
const gsi_leaf *
gsi_leaf_at(const gsi_view *v, int slot)
{
  return (const gsi_leaf *) (v->leaf + (size_t) slot * (size_t) v->leaf_stride);
}

static inline gsi_view
gsi_view_of(const void *slab_base, const gsi_slab_header *h)
{
  const char *b = (const char *) slab_base;
  gsi_view v;
  v.xf = h->xf;
  v.dag = (const gsi_rec *) (b + h->off_dag);
  v.leaf = b + h->off_leaf;
  v.leaf_stride = (int32_t) h->leaf_stride;
  v.z = (const double *) (b + h->off_z);
  v.seeds = (const int32_t *) (b + h->off_seeds);
  v.seed16 = (const uint16_t *) (b + h->off_seed16);
  v.seedbase = (const int32_t *) (b + h->off_seedbase);
  v.seed_shift = (int32_t) h->seed_shift;
  v.grid_gb = (int32_t) gsi_seed_blocks(h->grid_g, h->seed_shift);
  v.entry = (const int32_t *) (b + h->off_entry);
  v.tiles = (const int32_t *) (b + h->off_tiles);
  v.tile_bits = (int32_t) h->tile_bits;
  v.memo = h->memo_bits ? (const gsi_memo *) (b + h->off_memo) : (const gsi_memo *) 0;
  v.memo_mask = h->memo_bits ? (((uint32_t) 1) << h->memo_bits) - 1u : 0u;
  v.delta = h->delta;
  v.u0 = h->u0; v.v0 = h->v0; v.inv_du = h->inv_du; v.inv_dv = h->inv_dv;
  v.grid_g = (int32_t) h->grid_g;
  v.n_fields = (int32_t) h->n_fields;
  v.n_dag = (int32_t) h->n_dag;
  v.n_leaves = (int32_t) h->n_leaves;
  v.n_points = h->n_points;
  return v;
}

#endif
