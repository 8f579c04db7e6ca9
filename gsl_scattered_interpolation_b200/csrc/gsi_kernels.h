/* gsi_kernels.h -- host-callable launchers of the sm_100a kernels (gsi_kernels.cu). */
#ifndef GSI_KERNELS_H
#define GSI_KERNELS_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "gsi_layout.h"

#define GSI_LOCATE_THREADS 256
#define GSI_WALK_THREADS 128
#define GSI_BIN_THREADS 512        /* k_bin: one CTA sorts one chunk in shared memory, two CTAs per SM */
#define GSI_UNPERMUTE_THREADS 1024 /* k_unpermute */
#define GSI_BIN_CHUNK_MAX 8192  /* queries per chunk (16-bit in-chunk indices, 96 KB of shared memory) */
#define GSI_TILE_THREADS 768    /* k_tile_walk: one persistent CTA per SM, 24 warps, <= 85 registers */
#define GSI_TILE_SPOS 256       /* k_tile_walk: entries of a warp's dense position list */

/* device counters (unsigned[16]) */
#define GSI_CNT_DEFERRED 0       /* u32: length of the deferred list of the current launch */
#define GSI_CNT_OUTSIDE 1        /* u32: queries outside the cage since the last reset */
#define GSI_CNT_DEFERRED_TOTAL 2 /* u64: deferred queries since the last reset */
#define GSI_CNT_STEPS 4          /* u64: walk steps of accepted queries (stats launches only) */
#define GSI_CNT_OVERFLOW 8       /* u32: deferred queries of the current launch that did not fit the list */
#define GSI_CNT_WORK 9           /* u32: chunk counter of k_walkp (zero between launches) */
#define GSI_CNT_DEFERRED2 10     /* u32: deferred queries left after the vertex-memo pass (zero between launches) */
#define GSI_CNT_WORDS 16

#define GSI_MODE_FAST 0  /* grid-seeded walk + exact replay of the deferred queries */
#define GSI_MODE_EXACT 1 /* exact DAG replay on every query */

#define GSI_KERNEL_WALK 0   /* k_walk: lanes decoupled from queries (default) */
#define GSI_KERNEL_SIMPLE 1 /* k_locate: one thread per query, two-phase (coherent query sets) */
#define GSI_KERNEL_AUTO 2   /* k_walk for query arrays, k_locate for rasters (default) */

/* Scratch of the reorder pipeline (k_bin -> k_tile_walk -> k_unpermute), one per evaluation
 * context.  Queries are processed in windows of `cap` queries; a window is cut into chunks of
 * `chunk` consecutive queries and every chunk is counting-sorted by spatial tile (Morton-key
 * prefix, the tiles of the slab's tile table) inside shared memory, so both the reorder and its
 * inverse are chunk-local permutations with coalesced global traffic. */
struct gsi_reorder_scratch
{
  double *bx = nullptr, *by = nullptr; /* query coordinates, tile order inside every chunk */
  unsigned short *bi = nullptr;        /* position of the binned query inside its chunk */
  unsigned *runs = nullptr;            /* [n_bins][n_chunks] (bin-major, so the tile kernel reads one
                                          contiguous row per tile): start | end << 16 of the bin's run
                                          inside the chunk */
  unsigned *work = nullptr;            /* dynamic work counter of the tile kernel */
  double *zb = nullptr;                /* results in binned order, field-major, ld = cap */
  int *lb = nullptr;
  size_t cap = 0;     /* queries per window */
  int chunk = 0;      /* queries per chunk */
  int tile_max_leaves = 0; /* leaves of the largest tile (slab header): size of the staging area */
  int ranges = 0;     /* tile-kernel work items per tile (0 = derive from the launch size) */
  int tile_threads = GSI_TILE_THREADS;
};

struct gsi_launch_cfg
{
  int mode = GSI_MODE_FAST;
  int stats = 0;
  int max_steps = 256;
  int sm_count = 148;
  int locate_ctas_per_sm = 8;
  int kernel = GSI_KERNEL_AUTO;
  int walk_ctas_per_sm = 5;
  int exact_coop = 1; /* deferred list: 1 = four lanes per query, eight queries per warp (k_exact_list4, default),
                         2 = groups refilled as they finish (k_exact_list4p: measured slower), 0 = one thread (k_exact_list) */
  size_t walk_chunk = 0; /* 0 = derive from the launch size */
  int walk_variant = 2;  /* 0 = k_walk, 1 = k_walkp (two walks per lane in rotation behind L2 prefetches),
                            2 = k_walkc (warp-cooperative cp.async staging; needs 128-byte leaf records) */
  int walk_dynamic = 1;  /* k_walkp: chunks dealt by a device counter (0: round-robin) */
  int walk_l1 = -1;       /* k_walkc: leaf lines allocate in L1 (cp.async.ca) 1, not (cp.async.cg) 0, -1 = by structure size */
  int walk_carveout = -1; /* k_walkc / k_walkf: cudaFuncAttributePreferredSharedMemoryCarveout in percent, -1 = default */
  int walk_pf = 1;       /* k_walkp: how the next leaf line is requested ahead (gsi_prefetch_leaf_l2) */
  unsigned long long *walk_trace = nullptr; /* k_walkp / k_walkc: 4 words per CTA {smid, start ns, end ns, queries}, zeroed here */
  size_t walk_trace_ctas = 0;               /* CTAs the trace buffer has room for */
  unsigned *list = nullptr;     /* deferred list, list_capacity entries */
  unsigned *list2 = nullptr;    /* list_capacity entries more: what is left of the list after the vertex-memo pass */
  size_t list_capacity = 0;     /* deferred-list entries; overflow goes to the bitmask */
  unsigned *ovf_mask = nullptr; /* one bit per query of a launch (zero between launches) */
  unsigned *counters = nullptr; /* GSI_CNT_WORDS words */
  const gsi_reorder_scratch *reorder = nullptr; /* non-NULL: query arrays go through the reorder pipeline */
  /* optional profiling events recorded around the two kernels of a launch */
  cudaEvent_t ev_locate0 = nullptr, ev_locate1 = nullptr, ev_exact0 = nullptr, ev_exact1 = nullptr;
};

/* locate + blend nq queries given as device arrays; val is field-major with leading
 * dimension ld (val[f*ld + q]); val and leaf_out may each be NULL */
cudaError_t gsi_launch_eval_arrays(const gsi_view &V, const double *xq, const double *yq,
                                   size_t nq, double *val, int *leaf_out, size_t ld,
                                   const gsi_launch_cfg &cfg, cudaStream_t st);

/* same for the columns [i_first, i_first + n_cols) of an nx x ny raster (x-major, pixel =
 * i*ny + j); val / leaf_out point at pixel i_first*ny */
cudaError_t gsi_launch_eval_raster(const gsi_view &V, double x0, double dx, double y0, double dy,
                                   unsigned long long ny, unsigned long long i_first,
                                   unsigned long long n_cols, double *val, int *leaf_out, size_t ld,
                                   const gsi_launch_cfg &cfg, cudaStream_t st);

cudaError_t gsi_launch_seed_cells(const gsi_view &V, int32_t *seeds, int sm_count,
                                  cudaStream_t st);

/* vertex memo: answers of the exact replay for the keys the flattener wrote; run after the entry nodes */
cudaError_t gsi_launch_memo_fill(const gsi_view &V, gsi_memo *memo, int sm_count, cudaStream_t st);
/* fill the compact seed table (seed16 + seedbase) from the int32 seeds; run after the seeding */
cudaError_t gsi_launch_pack_seeds(const gsi_view &V, const int32_t *seeds, uint16_t *seed16, int32_t *seedbase,
                                  int sm_count, cudaStream_t st);

/* fill the entry section (per-cell start node of the exact replay); run before the seeding */
cudaError_t gsi_launch_entry_cells(const gsi_view &V, int32_t *entry, int sm_count,
                                   cudaStream_t st);

/* gradient of the interpolant for the leaf ids of a locate pass (gsi_deriv_one); dzdx / dzdy are
 * field-major with leading dimension ld and may each be NULL */
cudaError_t gsi_launch_deriv(const gsi_view &V, const int *leaf, size_t nq, double *dzdx, double *dzdy, size_t ld,
                             int sm_count, cudaStream_t st);

/* count out-of-range indices in an adopted slab image (*bad, device memory) */
cudaError_t gsi_launch_validate_slab(const gsi_view &V, unsigned long long *bad, int sm_count, cudaStream_t st);

/* hoisted-reciprocal division against __ddiv_rn on n operand pairs: out2[0] = pairs vouched for,
 * out2[1] = mismatches among them (device memory) */
cudaError_t gsi_launch_selftest_division(uint64_t seed, size_t n, unsigned long long *out2, int sm_count, cudaStream_t st);

/* slabs with 128-byte leaf records: fill the reciprocals of each record's two divisors into its padding
 * (bytes 96..111), which k_walkc divides with; no-op for 96-byte records */
cudaError_t gsi_launch_leaf_rcp(const gsi_view &V, char *leaf_base, int sm_count, cudaStream_t st);

/* batched locate of the builder's points [first, first + count) on a snapshot of the growing tree
 * (device arrays as in gsi_host_tree + the LU records); outputs are device arrays of `count` */
cudaError_t gsi_launch_build_locate(const gsi_rec *rec, const int32_t *lnk, const uint8_t *type, const gsi_xform &xf,
                                    const double *px, const double *py, size_t first, size_t count, int32_t *node,
                                    double *c0, double *c1, int32_t *last, int sm_count, cudaStream_t st);

cudaError_t gsi_launch_gen_uniform(uint64_t seed, uint64_t first, size_t nq, double xlo,
                                   double xhi, double ylo, double yhi, double *xq, double *yq,
                                   int sm_count, cudaStream_t st);

cudaError_t gsi_launch_checksum(const int *leaf, const double *val, size_t n,
                                unsigned long long *out2, int sm_count, cudaStream_t st);

cudaError_t gsi_launch_compare(const int *leaf_a, const int *leaf_b, const double *val_a,
                               const double *val_b, size_t n, size_t n_val,
                               unsigned long long *out2, int sm_count, cudaStream_t st);

#endif
