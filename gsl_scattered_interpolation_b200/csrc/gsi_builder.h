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

  size_t n_nodes() const { return type.size(); }
};

/* Fill `order` with gsl_ran_shuffle(gsl_rng_mt19937 seeded `seed`) applied to the identity. */
void gsi_shuffle_mt19937(unsigned long seed, size_t n, int64_t *order);

/* Build the tree of the points (xa[i], ya[i]).  order == NULL: data order.
 * Returns 0 on success; on failure returns a GSL error code and sets *err. */
int gsi_build_tree(const double *xa, const double *ya, size_t n, int flags,
                   const int64_t *order, gsi_host_tree *out, std::string *err);

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
