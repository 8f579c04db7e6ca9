/* gsi_layout.h -- the device triangulation layout ("slab") of libgsl_scattered_cuda.
 *
 * One contiguous, position-independent allocation per GPU (offsets only, so it can be
 * replicated with a plain peer copy / NCCL broadcast).  Sections, all 256-byte aligned:
 *
 *   header   gsi_slab_header
 *   dag      gsi_rec[n_dag]     history-DAG nodes in the reference's node-id order (exact
 *                               replay kernel; a leaf's lnk[0] holds its compact slot)
 *   leaf     gsi_rec[n_leaves]  leaves only, sorted along a Morton curve of their centroids;
 *                               lnk[] = slot of the neighbour opposite vertex i, -1 = none
 *                               (hull of the cage); meta = float margin | swap sign bit
 *   aux      int32[4][n_leaves] {vertex slot of POINT(leaf,0..2) or -1 for a cage seed,
 *                               reference node id}  -- read once per query, 128-bit load
 *   z        double[n_points][n_fields]  vertex values, vertex-major so the F fields of a
 *                               vertex are contiguous (F = 16 -> one 128-byte line), vertices
 *                               sorted along the same Morton curve
 *   seeds    int32[G*G]         uniform grid over the data bounding box in standardised
 *                               coordinates: leaf slot containing the cell centre
 *                               (filled on the device by the exact-replay kernel)
 *
 * Reference data model this flattens: simplex_tree / simplex_tree_node pools,
 * /root/reference/interpolation/linear_simplex.h:16-65 (POINT/LINK/LEAF/DATA_POINT).
 */
#ifndef GSI_LAYOUT_H
#define GSI_LAYOUT_H

#include <stdint.h>

#include "gsi_core.h"

#define GSI_SLAB_MAGIC 0x3142414c53495347ull /* "GSISLAB1" */
#define GSI_SLAB_VERSION 1u

typedef struct
{
  uint64_t magic;
  uint32_t version;
  uint32_t n_fields;
  uint64_t n_points;
  uint64_t n_dag;
  uint64_t n_leaves;
  uint32_t grid_g;
  uint32_t reserved0;
  gsi_xform xf;
  /* grid mapping in standardised coordinates: cell = floor((u - u0) * inv_du) */
  double u0, v0, inv_du, inv_dv;
  double delta; /* absolute safety margin (standardised units) behind the per-leaf tau */
  uint64_t off_dag, off_leaf, off_aux, off_z, off_seeds;
  uint64_t total_bytes;
  uint64_t reserved[8];
} gsi_slab_header;

/* what the kernels receive by value */
typedef struct
{
  gsi_xform xf;
  const gsi_rec *dag;
  const gsi_rec *leaf;
  const int32_t *aux; /* 4 per leaf */
  const double *z;
  const int32_t *seeds;
  double u0, v0, inv_du, inv_dv;
  int32_t grid_g;
  int32_t n_fields;
  int32_t n_dag;
  int32_t n_leaves;
} gsi_view;

static inline gsi_view
gsi_view_of(const void *slab_base, const gsi_slab_header *h)
{
  const char *b = (const char *) slab_base;
  gsi_view v;
  v.xf = h->xf;
  v.dag = (const gsi_rec *) (b + h->off_dag);
  v.leaf = (const gsi_rec *) (b + h->off_leaf);
  v.aux = (const int32_t *) (b + h->off_aux);
  v.z = (const double *) (b + h->off_z);
  v.seeds = (const int32_t *) (b + h->off_seeds);
  v.u0 = h->u0; v.v0 = h->v0; v.inv_du = h->inv_du; v.inv_dv = h->inv_dv;
  v.grid_g = (int32_t) h->grid_g;
  v.n_fields = (int32_t) h->n_fields;
  v.n_dag = (int32_t) h->n_dag;
  v.n_leaves = (int32_t) h->n_leaves;
  return v;
}

#endif
