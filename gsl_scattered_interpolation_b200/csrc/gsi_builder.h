/* gsi_builder.h -- host-side triangulation builder of libgsl_scattered_cuda.
 *
 * The Delaunay builder "stays on the host and is timed separately" (BASELINE.json
 * north_star).  This is a from-scratch C++ restatement, specialised to dim = 2, of the
 * reference's incremental builder
 *   simplex_tree_init   /root/reference/interpolation/linear_simplex.c:134-296
 *   insert_point        linear_simplex.c:404-492
 *   in_hypersphere*     linear_simplex.c:495-605
 *   delaunay + helpers  /root/reference/interpolation/edge_flip.c:17-320
 *   gsl_ran_shuffle     /root/reference/randist/shuffle.c:69-79 (+ rng/mt.c, rng/gsl_rng.h:188-215)
 * It must reproduce the reference's history DAG node for node (ids, vertex order, links,
 * types): the "triangle index" the eval path returns is the DAG node id, and the
 * edge/vertex tie-breaks depend on the DAG.  tests/test_builder.py compares it with the
 * tree built by the reference itself (oracle/_ref).
 */
#ifndef GSI_BUILDER_H
#define GSI_BUILDER_H

#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "gsi_core.h"

#define GSI_FLAG_NOSTANDARDIZE 1 /* SIMPLEX_TREE_NOSTANDARDIZE, linear_simplex.h:111 */
#define GSI_FLAG_ISOSCALE 2      /* SIMPLEX_TREE_ISOSCALE,      linear_simplex.h:112 */

struct gsi_host_tree
{
  size_t n_points = 0;          /* data points */
  std::vector<int32_t> pts;     /* 3 per node: POINT(node,i); negative = cage seed -(k+1) */
  std::vector<int32_t> lnk;     /* 3 per node: LINK(node,i) */
  std::vector<uint8_t> type;    /* node_type */
  std::vector<int64_t> shuffle; /* tree->shuffle: point id -> row of the user's data */
  gsi_xform xf;                 /* scale / shift */
  double seed[6];               /* tree->seed_points, row-major 3x2, raw coordinates */
  double build_seconds = 0;
  /* batched locate (gsi_build_locator): points located through it, rounds, seconds inside it */
  size_t batched_points = 0, batched_rounds = 0;
  double batched_seconds = 0;

  size_t n_nodes() const { return type.size(); }
};

/* Batched point location for the builder.
 *
 * insert_point's caller locates every new point with the same descent the evaluation path uses
 * (linear_simplex.c:283-293 -> find_leaf, :331-402), one point after the other, and that descent
 * is about half of the build.  A node's children never change once it has them, so the descent
 * of point i through the nodes that existed when point s <= i was inserted is the same whenever it
 * is run: the builder hands the points [s, e) to a locator together with the tree as of point s,
 * gets back for each the node where that descent leaves the snapshot (a leaf of the snapshot) and
 * the state it arrived with (accel->coords, the node they belong to), and finishes each descent
 * on the live tree from there (GSI_DEFINE_DESCEND, resume = 1) right before inserting the point --
 * a few levels instead of ~45 at 10^6 nodes.  The result is the reference's tree node for node.
 * The CUDA implementation lives in gsi_capi.cu / gsi_kernels.cu (k_build_locate). */
struct gsi_build_locator
{
  virtual ~gsi_build_locator() {}
  /* called once before the first round: the points in insertion order, the transform, and an
     upper bound of the number of nodes the tree will have */
  virtual int begin(const gsi_xform &xf, const double *px, const double *py, size_t n, size_t max_nodes) = 0;
  /* The tree now has n_nodes nodes, n_prev of which were there at the previous call (0 at the
     first).  rec[i] = LU half + swap/singular bits of node i (immutable once made), lnk / type as
     in gsi_host_tree (a node's links are final once its type is not GSI_LEAF).  Locate the points
     [first, first + count): node[k] = leaf of this snapshot reached by point first + k (< 0:
     GSI_DESCEND_OUTSIDE / GSI_DESCEND_LOOP), c0/c1/last = the state on arrival. */
  virtual int locate(const gsi_rec *rec, const int32_t *lnk, const uint8_t *type, size_t n_nodes, size_t n_prev,
                     size_t first, size_t count, int32_t *node, double *c0, double *c1, int32_t *last) = 0;
};

/* Fill `order` with gsl_ran_shuffle(gsl_rng_mt19937 seeded `seed`) applied to the identity. */
void gsi_shuffle_mt19937(unsigned long seed, size_t n, int64_t *order);

/* Build the tree of the points (xa[i], ya[i]).  order == NULL: data order.
 * Returns 0 on success; on failure returns a GSL error code and sets *err. */
int gsi_build_tree(const double *xa, const double *ya, size_t n, int flags,
                   const int64_t *order, gsi_host_tree *out, std::string *err,
                   gsi_build_locator *locator = nullptr);

/* a locator that runs the descent on the host over a COPY of the snapshot (tests of the batched
 * protocol without a GPU; never used by the library itself) */
gsi_build_locator *gsi_make_host_snapshot_locator();

/* Raw coordinates of vertex id `p` (negative = seed), DATA_POINT of linear_simplex.h:82-93. */
static inline void
gsi_vertex_xy(const gsi_host_tree &t, const double *xa, const double *ya, int32_t p, double *x,
              double *y)
{
  if (p < 0)
    {
      *x = t.seed[2 * (-p - 1)];
      *y = t.seed[2 * (-p - 1) + 1];
    }
  else
    {
      *x = xa[t.shuffle[p]];
      *y = ya[t.shuffle[p]];
    }
}

/* LU records for every DAG node (meta = type|swap|singular, lnk = LINK(node, 0..2)). */
void gsi_make_dag_records(const gsi_host_tree &t, const double *xa, const double *ya,
                          std::vector<gsi_rec> *recs);

#endif
