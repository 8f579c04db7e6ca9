/* gsl_scattered_cuda.h -- C ABI of libgsl_scattered_cuda (B200 / sm_100a).
 *
 * Batched scattered-data interpolation on the GPU: locate each 2-D query point in the
 * Delaunay triangulation and blend the containing triangle's vertex values with
 * barycentric weights.  It is a drop-in for the evaluation loop of the GSL
 * scattered-interpolation branch,
 *
 *     leaf = find_leaf(tree, data, point, accel);                       (linear_simplex.c:331-350)
 *     val  = interp_point(tree, leaf, data, response, point, accel);    (linear_simplex.c:678-711)
 *
 * as driven by /root/reference/interpolation/scattered_interp_example.c:158-160,171-173,
 * 184-199, and returns exactly what that loop returns: the containing leaf as the reference's
 * simplex_index (node id in the history DAG, linear_simplex.h:6) bit-exact -- including the
 * reference's edge/vertex tie-breaks and its cage leaves for points outside the data hull --
 * and the interpolated values bit-identical in binary64.
 *
 * Conventions follow gsl_interp2d / gsl_spline2d (interpolation/gsl_interp2d.h:37-80,
 * interpolation/spline2d.c:30-110): opaque workspace, alloc returns NULL on failure, init
 * returns a status, eval returns the value and eval_e a status with out-parameters, free(NULL)
 * is a no-op.  Status codes are GSL's numeric values (err/gsl_errno.h:38-108).  The header is
 * plain C89, has no GSL or CUDA types in any signature and can be bound over FFI (see
 * INTEGRATION.md for the reference-side stub).
 */
#ifndef GSL_SCATTERED_CUDA_H
#define GSL_SCATTERED_CUDA_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 64-bit unsigned counter / seed type.  Strict C89 has no long long; on LP64 (the only data model
 * CUDA supports on Linux) unsigned long is the same 64-bit type at the ABI level. */
#if defined(__cplusplus) || (defined(__STDC_VERSION__) && __STDC_VERSION__ >= 199901L)
typedef unsigned long long gsl_scattered_cuda_u64;
#else
typedef unsigned long gsl_scattered_cuda_u64;
typedef char gsl_scattered_cuda_u64_must_have_64_bits[sizeof (unsigned long) == 8 ? 1 : -1];
#endif

/* status codes, numerically identical to err/gsl_errno.h:38-75 */
#define GSL_SCATTERED_SUCCESS 0
#define GSL_SCATTERED_EDOM 1     /* at least one query lies outside the cage (find_leaf would assert) */
#define GSL_SCATTERED_EINVAL 4   /* invalid argument */
#define GSL_SCATTERED_EFAILED 5  /* CUDA failure / inconsistent tree */
#define GSL_SCATTERED_ENOMEM 8   /* host or device allocation failed */
#define GSL_SCATTERED_EBADLEN 19 /* array lengths do not conform */

/* init_flags, identical to SIMPLEX_TREE_* (interpolation/linear_simplex.h:109-112) */
#define GSL_SCATTERED_DEFAULT 0
#define GSL_SCATTERED_NOSTANDARDIZE (1 << 0)
#define GSL_SCATTERED_ISOSCALE (1 << 1)

typedef struct gsl_scattered_cuda gsl_scattered_cuda;

/* Error hook with the signature of gsl_error_handler_t (err/gsl_errno.h:110-123).  Every
 * failing call reports through it before returning its status; pass gsl_error itself to get
 * GSL's policy (print + abort unless gsl_set_error_handler_off).  Default: none (status only).
 * Returns the previous handler. */
typedef void gsl_scattered_cuda_error_handler_t(const char *reason, const char *file, int line,
                                                int gsl_errno);
gsl_scattered_cuda_error_handler_t *
gsl_scattered_cuda_set_error_handler(gsl_scattered_cuda_error_handler_t *handler);

/* Number of usable CUDA devices (0 if none / no driver). */
int gsl_scattered_cuda_device_count(void);

/* ---- workspace --------------------------------------------------------------------------
 * replaces simplex_tree_alloc + simplex_tree_accel_alloc (linear_simplex.c:53-106,298-307)
 * and simplex_tree_free (linear_simplex.c:108-132). */
gsl_scattered_cuda *gsl_scattered_cuda_alloc(size_t n_nodes, size_t n_fields);
void gsl_scattered_cuda_free(gsl_scattered_cuda *w);

/* Select the GPUs of this process that hold a replica of the triangulation (default:
 * device 0 only).  Call before init/upload.  eval_e / eval_raster_e split their queries
 * into one contiguous slice per device; no collective is involved. */
int gsl_scattered_cuda_set_devices(gsl_scattered_cuda *w, const int dev[], size_t ndev);

/* Tuning knobs (before init/upload).  Known names: "cells_per_leaf" (grid resolution,
 * default 4), "grid_g" (explicit grid size), "max_steps" (walk cap, default 256),
 * "delta" (edge safety margin in standardised units, default 2^-28), "mode" (0 = fast
 * path + exact replay of deferred queries, 1 = exact replay of every query),
 * "stats" (1 = count walk steps), "profile" (1 = time every kernel launch with CUDA events on
 * its own stream), "chunk" (queries per launch), "host_chunk" (queries per pipeline stage
 * of eval_e), "list_fraction" (deferred-list size / launch size), "ctas_per_sm",
 * "walk_ctas_per_sm", "walk_chunk", "kernel" (0 = k_walk, 1 = k_locate, 2 = k_walk for query
 * arrays and k_locate for rasters, default), "leaf_stride" (96 or 128), "l2_fetch".
 * Exact replay: "exact_entry" (1 = start at the grid cell's certified entry node, default;
 * 0 = at the root; may be changed at any time); "vertex_memo" (1 = keep the exact answers for
 * queries that coincide bit for bit with a data node in a hash table built at init, default;
 * before init/upload); "exact_coop" (1 = four lanes per deferred query, default).
 * Walk kernel: "walk_variant" (2 = k_walkc, default; 3 = k_walkf; 0 / 1 = the round-1 kernels),
 * "walk_carveout" (preferred shared-memory carve-out of the walk kernel in percent, default 72;
 * -1 = the driver's choice).
 * Reorder pipeline for incoherent query arrays (chunk-local counting sort by spatial tile ->
 * tile-major walk out of TMA-staged shared memory -> inverse permutation): "reorder" (0 = off,
 * default; 1 = always; 2 = auto: batches of at least "reorder_min" queries (2^24) with at most
 * "reorder_max_fields" fields (2) on triangulations of at least "reorder_min_leaves" leaves
 * (600000: leaf records beyond the reach of the L2)), "reorder_window" (queries per window,
 * default 2^27), "reorder_chunk" (queries per sorted chunk, <= 8192), "tile_leaves" (target
 * leaves per Morton-prefix tile, default 1000; fixed at init/upload), "tile_ranges" (work items
 * per tile, 0 = derived from the launch size), "tile_threads" (threads of the tile kernel's CTA).
 * Host pipeline: "copy_threads" (host threads that move pageable caller arrays through pinned
 * staging, default: three quarters of the host's hardware threads, 4 to 16), "pageable_min" (smaller pageable batches go straight to cudaMemcpyAsync).
 * Parameters are read when a call starts: change them between calls, not during one. */
int gsl_scattered_cuda_set_param(gsl_scattered_cuda *w, const char *name, double value);

/* order[] for init: what gsl_ran_shuffle(rng, shuffle, n, sizeof(size_t)) leaves in
 * tree->shuffle for rng = gsl_rng_mt19937 seeded with `seed` (linear_simplex.c:280-281,
 * randist/shuffle.c:69-79, rng/mt.c).  seed 0 = gsl_rng_default's default. */
int gsl_scattered_cuda_shuffle_mt19937(unsigned long seed, size_t n, size_t order[]);

/* ---- init -------------------------------------------------------------------------------
 * replaces simplex_tree_init (linear_simplex.c:134-296) with min = max = NULL.
 * xa, ya: node coordinates; za: n_fields x n_nodes values, field-major (field f of node i at
 * za[f*n_nodes + i]); order: insertion order (NULL = data order, i.e. rng == NULL).  The
 * triangulation is built on the host by the library's own restatement of the reference's
 * incremental builder (same DAG, node for node), flattened and uploaded to every selected
 * GPU.  Inputs are copied; nothing is borrowed after return.
 *
 * init == build (host only, timed separately) followed by upload (needs a GPU). */
int gsl_scattered_cuda_init(gsl_scattered_cuda *w, const double xa[], const double ya[],
                            const double za[], size_t n_nodes, int init_flags,
                            const size_t order[]);
int gsl_scattered_cuda_build(gsl_scattered_cuda *w, const double xa[], const double ya[],
                             size_t n_nodes, int init_flags, const size_t order[]);
int gsl_scattered_cuda_upload(gsl_scattered_cuda *w, const double za[], size_t n_fields);

/* True drop-in: adopt a tree built by the reference's own simplex_tree_init.  The arrays are
 * the flattened simplex_tree pools (linear_simplex.h:31-59): for node k, pts3[3k+i] =
 * POINT(k,i), lnk3[3k+i] = LINK(k,i), type[k] = SIMP(k)->type; shuffle = tree->shuffle->data;
 * shift/scale = tree->shift/scale; seed_points = tree->seed_points (3x2 row-major);
 * data = the n_nodes x 2 gsl_matrix the tree was built on (row stride data_tda).
 * INTEGRATION.md shows the 20-line stub that fills these from a simplex_tree*. */
int gsl_scattered_cuda_init_from_tree(gsl_scattered_cuda *w, size_t n_simplexes,
                                      const int pts3[], const int lnk3[], const int type[],
                                      const size_t shuffle[], const double shift[2],
                                      const double scale[2], const double seed_points[6],
                                      const double data[], size_t data_tda, size_t n_nodes,
                                      const double za[], size_t n_fields);

/* ---- eval -------------------------------------------------------------------------------
 * replaces the loop  find_leaf + n_fields x interp_point  over nq points.
 * xq, yq: HOST arrays of nq query coordinates.  zq: n_fields x nq values, field-major, may
 * be NULL.  leaf: nq reference simplex_index values, may be NULL.  A query outside the cage
 * (the reference asserts there, linear_simplex.c:343-347) gets leaf -1 and NaN values and
 * makes the call return GSL_SCATTERED_EDOM after all other queries have been evaluated. */
int gsl_scattered_cuda_eval_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                              size_t nq, double zq[], int leaf[]);

/* one point, field 0 (NaN outside the cage) */
double gsl_scattered_cuda_eval(const gsl_scattered_cuda *w, double x, double y);

/* The reference's raster driver (scattered_interp_example.c:176-199): pixel (i,j) is the
 * point (x0 + dx*i, y0 + dy*j) -- two roundings each, as the reference computes them --
 * results x-major: zq[f*nx*ny + i*ny + j], leaf[i*ny + j].  HOST outputs. */
int gsl_scattered_cuda_eval_raster_e(const gsl_scattered_cuda *w, double x0, double dx, size_t nx,
                                     double y0, double dy, size_t ny, double zq[], int leaf[]);

/* Same two entry points on DEVICE pointers of device `dev` (must be one of the selected
 * devices), asynchronous on `stream` (a cudaStream_t passed as void*, NULL = the library's
 * own stream followed by a synchronize).  ld = leading dimension of d_zq (>= nq).
 * With a caller's stream the call returns as soon as the work is enqueued, hence always
 * GSL_SCATTERED_SUCCESS for well-formed arguments: queries outside the cage then show up only as
 * leaf -1 / NaN in the outputs and in gsl_scattered_cuda_info.n_outside (read it after
 * synchronising the stream); GSL_SCATTERED_EDOM is reported by the synchronous form (NULL). */
int gsl_scattered_cuda_eval_device_e(const gsl_scattered_cuda *w, int dev, const double *d_xq,
                                     const double *d_yq, size_t nq, double *d_zq, size_t ld,
                                     int *d_leaf, void *stream);
int gsl_scattered_cuda_eval_raster_device_e(const gsl_scattered_cuda *w, int dev, double x0,
                                            double dx, size_t nx, double y0, double dy, size_t ny,
                                            double *d_zq, int *d_leaf, void *stream);

/* ---- derivatives and the type descriptor ----------------------------
 * The counterpart of gsl_interp2d_eval_deriv_x / _y (interpolation/gsl_interp2d.h:37-49,
 * interpolation/spline2d.c:30-110) for the scattered method.  The interpolant is affine on every
 * leaf, so the derivatives are the (piecewise constant) gradient of the containing leaf's plane:
 * the derivatives of the barycentric coordinates come from the same LU solve as the coordinates
 * (calculate_bary_coords, linear_simplex.c:607-651) with right-hand side scale_k e_k, blended with
 * the vertex values like interp_point does (linear_simplex.c:689-710: cage vertices contribute
 * nothing, last weight 1 - tot -> 0 - tot).  The containing leaf is the one find_leaf returns,
 * tie-breaks included.  Outside the cage and on singular leaves: NaN (status GSL_SCATTERED_EDOM
 * for a point outside the cage).  Second derivatives are zero wherever the first exist.
 * dzdx, dzdy: n_fields x nq, field-major, HOST arrays, either may be NULL. */
int gsl_scattered_cuda_eval_deriv_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                    size_t nq, double dzdx[], double dzdy[]);
int gsl_scattered_cuda_eval_deriv_x_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                      size_t nq, double d[]);
int gsl_scattered_cuda_eval_deriv_y_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                      size_t nq, double d[]);
int gsl_scattered_cuda_eval_deriv_xx_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                       size_t nq, double d[]);
int gsl_scattered_cuda_eval_deriv_yy_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                       size_t nq, double d[]);
int gsl_scattered_cuda_eval_deriv_xy_e(const gsl_scattered_cuda *w, const double xq[], const double yq[],
                                       size_t nq, double d[]);
/* one point, field 0 */
double gsl_scattered_cuda_eval_deriv_x(const gsl_scattered_cuda *w, double x, double y);
double gsl_scattered_cuda_eval_deriv_y(const gsl_scattered_cuda *w, double x, double y);
/* DEVICE pointers, asynchronous on `stream` like eval_device_e; d_leaf (nq ints) receives the
 * containing leaves (it may be NULL: the library then uses scratch of its own). */
int gsl_scattered_cuda_eval_deriv_device_e(const gsl_scattered_cuda *w, int dev, const double *d_xq,
                                           const double *d_yq, size_t nq, double *d_dzdx,
                                           double *d_dzdy, size_t ld, int *d_leaf, void *stream);

/* Method descriptor in the style of gsl_interp2d_type (interpolation/gsl_interp2d.h:37-49), so
 * that code written against "a 2-D interpolation type with alloc / init / eval / eval_deriv_* /
 * free members" can hold the scattered GPU method next to gsl_interp2d_bilinear / _bicubic.  The
 * members are the batched entry points above (a GPU method is called on arrays, not point by
 * point; the one-point forms are gsl_scattered_cuda_eval / _eval_deriv_x / _eval_deriv_y). */
typedef struct
{
  const char *name;
  unsigned int min_size; /* smallest node count that yields a data triangle (fewer are accepted, as in
                            the reference's trivial_test, and interpolate inside the cage only) */
  gsl_scattered_cuda *(*alloc)(size_t n_nodes, size_t n_fields);
  int (*init)(gsl_scattered_cuda *w, const double xa[], const double ya[], const double za[],
              size_t n_nodes, int init_flags, const size_t order[]);
  int (*eval_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq,
                double zq[], int leaf[]);
  int (*eval_deriv_x_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[]);
  int (*eval_deriv_y_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[]);
  int (*eval_deriv_xx_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[]);
  int (*eval_deriv_xy_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[]);
  int (*eval_deriv_yy_e)(const gsl_scattered_cuda *w, const double xq[], const double yq[], size_t nq, double d[]);
  void (*free)(gsl_scattered_cuda *w);
} gsl_scattered_cuda_type;

/* the one method of this library: piecewise-linear interpolation on the Delaunay triangulation */
extern const gsl_scattered_cuda_type *gsl_scattered_cuda_linear_simplex;
const gsl_scattered_cuda_type *gsl_scattered_cuda_type_linear_simplex(void); /* same, for FFI without data symbols */
const char *gsl_scattered_cuda_name(const gsl_scattered_cuda *w);            /* cf. gsl_interp2d_name */
unsigned int gsl_scattered_cuda_min_size(const gsl_scattered_cuda *w);       /* cf. gsl_interp2d_min_size */
unsigned int gsl_scattered_cuda_type_min_size(const gsl_scattered_cuda_type *T);

/* ---- introspection / replication / measurement helpers ------------------------------------- */
typedef struct
{
  size_t n_nodes;          /* data points */
  size_t n_fields;
  size_t n_simplexes;      /* history-DAG nodes (tree->n_simplexes) */
  size_t n_leaves;
  size_t grid_g;           /* acceleration grid is grid_g x grid_g */
  size_t slab_bytes;       /* device bytes per replica */
  size_t n_devices;
  double build_seconds;    /* host Delaunay build */
  double flatten_seconds;  /* host flatten */
  double upload_seconds;   /* H2D + device seeding + peer replication */
  gsl_scattered_cuda_u64 n_deferred; /* queries resolved by the exact replay since reset */
  gsl_scattered_cuda_u64 n_outside;  /* queries outside the cage since reset */
  gsl_scattered_cuda_u64 n_steps;    /* walk steps (only with param "stats") */
  gsl_scattered_cuda_u64 n_kernel_launches; /* eval kernels launched since reset */
  double locate_seconds;   /* device time in the fast-path kernel (only with param "profile") */
  double exact_seconds;    /* device time in the exact-replay kernel (only with param "profile") */
  gsl_scattered_cuda_u64 n_reordered; /* queries that went through the reorder pipeline since reset */
  size_t n_contexts;       /* evaluation contexts created so far (one per concurrent eval call and device) */
} gsl_scattered_cuda_info;
int gsl_scattered_cuda_get_info(const gsl_scattered_cuda *w, gsl_scattered_cuda_info *info);
int gsl_scattered_cuda_reset_counters(const gsl_scattered_cuda *w);

/* Copy out the host tree in the format of init_from_tree (for parity checks against the
 * reference builder).  Any pointer may be NULL. */
int gsl_scattered_cuda_get_tree(const gsl_scattered_cuda *w, int pts3[], int lnk3[], int type[],
                                size_t shuffle[], double shift[2], double scale[2],
                                double seed_points[6]);

/* The device layout is one position-independent slab: replicate it to another process /
 * GPU by copying the bytes (NCCL broadcast, peer copy) and adopting them there. */
size_t gsl_scattered_cuda_slab_bytes(const gsl_scattered_cuda *w);
int gsl_scattered_cuda_slab_export_device(const gsl_scattered_cuda *w, int dev, void *d_dst,
                                          void *stream);
int gsl_scattered_cuda_init_from_slab_device(gsl_scattered_cuda *w, const void *d_src,
                                             size_t bytes);

/* ---- serialisation of the flattened triangulation -----------------
 * The slab (header, history DAG, leaves, vertex values, seed grid) written as one binary
 * image, in the style of gsl_block_fwrite / gsl_block_fread (block/fwrite_source.c:23-60):
 * native byte order, GSL_SUCCESS or GSL_SCATTERED_EFAILED ("fwrite failed" / "fread failed").
 * fread checks magic, layout version, size and the image checksum (GSL_SCATTERED_EINVAL on
 * mismatch), needs w allocated with the same n_fields, and uploads to the selected devices:
 * a run can skip the host Delaunay build (42 s at 10^6 nodes with the reference's builder).
 * A workspace restored this way has no host tree (get_tree fails), like init_from_slab_device.
 * save/load are the same on a path (FILE* is awkward across FFI boundaries). */
int gsl_scattered_cuda_fwrite(FILE *stream, const gsl_scattered_cuda *w);
int gsl_scattered_cuda_fread(FILE *stream, gsl_scattered_cuda *w);
int gsl_scattered_cuda_save(const gsl_scattered_cuda *w, const char *path);
int gsl_scattered_cuda_load(gsl_scattered_cuda *w, const char *path);

/* Pinned host memory for zero-staging eval_e calls. */
void *gsl_scattered_cuda_host_alloc(size_t bytes);
void gsl_scattered_cuda_host_free(void *p);

/* Deterministic synthetic queries, generated on the device:
 * query k of the stream `seed` is (xlo + (xhi-xlo)*u(seed,2k), ylo + (yhi-ylo)*u(seed,2k+1))
 * with u = splitmix64-based uniform; fills k in [first, first+nq). */
int gsl_scattered_cuda_generate_uniform_device(const gsl_scattered_cuda *w, int dev,
                                               gsl_scattered_cuda_u64 seed, gsl_scattered_cuda_u64 first,
                                               size_t nq, double xlo, double xhi, double ylo,
                                               double yhi, double *d_xq, double *d_yq,
                                               void *stream);
/* same generator on the host (bit-identical) */
void gsl_scattered_cuda_generate_uniform_host(gsl_scattered_cuda_u64 seed, gsl_scattered_cuda_u64 first,
                                              size_t nq, double xlo, double xhi, double ylo,
                                              double yhi, double xq[], double yq[]);

/* out[0] = wrapping sum of leaf ids, out[1] = wrapping sum of the value bit patterns */
int gsl_scattered_cuda_checksum_device(const gsl_scattered_cuda *w, int dev, const int *d_leaf,
                                       const double *d_val, size_t n, gsl_scattered_cuda_u64 out[2],
                                       void *stream);
/* out[0] = number of differing leaf ids, out[1] = number of differing value bit patterns */
int gsl_scattered_cuda_compare_device(const gsl_scattered_cuda *w, int dev, const int *d_leaf_a,
                                      const int *d_leaf_b, const double *d_val_a,
                                      const double *d_val_b, size_t n, size_t n_val,
                                      gsl_scattered_cuda_u64 out[2], void *stream);

/* Diagnostic: the tile kernel of the reorder pipeline divides by per-leaf constants with the
 * divisor's reciprocal hoisted out (csrc/gsi_core.h: the operations of nvcc's own binary64 division,
 * split at the point where they stop depending on the divisor only).  This runs n operand pairs
 * (data-like, raw bit patterns, the whole exponent range) on device `dev`: out[0] = pairs the
 * hoisted form vouched for, out[1] = quotients among them that differ from __ddiv_rn (must be 0). */
int gsl_scattered_cuda_selftest_division(int dev, gsl_scattered_cuda_u64 seed, size_t n,
                                         gsl_scattered_cuda_u64 out[2]);

/* The builder locates every point it inserts with the evaluation path's own descent
 * (linear_simplex.c:283-293 -> find_leaf); with a CUDA device those locations run there in batches
 * against snapshots of the growing tree (set_param "build_device" = 0: host only).  The tree is the
 * same node for node.  Reports how much of the last build went that way. */
int gsl_scattered_cuda_build_stats(const gsl_scattered_cuda *w, double *batched_seconds, size_t *batched_points,
                                   size_t *rounds);

/* Diagnostic (set_param "walk_trace" = 1 before evaluating): four words per thread block of the most
 * recent fast-path launch on device `dev` -- {SM id, start, end (globaltimer ns), queries handed
 * out}.  Shows whether a slow launch is slow everywhere or on some SMs only.  *n_words = words
 * written (<= cap_words). */
int gsl_scattered_cuda_debug_walk_trace(const gsl_scattered_cuda *w, int dev, gsl_scattered_cuda_u64 *out,
                                        size_t cap_words, size_t *n_words);

const char *gsl_scattered_cuda_version(void);

#ifdef __cplusplus
}
#endif

#endif /* GSL_SCATTERED_CUDA_H */
