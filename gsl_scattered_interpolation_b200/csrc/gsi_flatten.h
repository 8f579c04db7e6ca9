/* gsi_flatten.h -- host tree -> device slab image (layout in gsi_layout.h). */
#ifndef GSI_FLATTEN_H
#define GSI_FLATTEN_H

#include <string>
#include <vector>

#include "gsi_builder.h"
#include "gsi_layout.h"

struct gsi_flatten_params
{
  double cells_per_leaf = 4.0;  /* grid resolution: G*G ~ cells_per_leaf * n_leaves */
  uint32_t grid_g = 0;          /* explicit G (0 = derive from cells_per_leaf) */
  double delta = 3.725290298461914e-09; /* 2^-28: absolute edge margin, standardised units */
  double tau_floor = 9.094947017729282e-13; /* 2^-40 */
  uint32_t leaf_stride = 128;   /* 128 = one 128-byte line per leaf (default: k_walkc fetches a leaf as one line and keeps
                                   the reciprocals of its divisors in the spare bytes), 96 = packed records */
  uint32_t tile_leaves = 1000;  /* target leaves per Morton-prefix tile (reorder pipeline) */
  int vertex_memo = 1;          /* hash table of the data nodes' coordinates for the exact answers at the nodes */
};

/* za: n_fields x n_points, field-major (may be NULL: values left zero).
 * The seeds, seed16, seedbase and entry sections are left zero (entry 0 = root); they are filled on
 * the device (seed16 all zero + seedbase all zero is a valid, if useless, compact table: slot 0).  The
 * memo section receives the keys (state PENDING); k_memo_fill computes the answers on the device. */
int gsi_flatten(const gsi_host_tree &t, const double *xa, const double *ya, const double *za,
                size_t n_fields, const gsi_flatten_params &prm, std::vector<char> *image,
                std::string *err);

#endif
