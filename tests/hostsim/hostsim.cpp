/* tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Runs the per-query logic of the CUDA kernels (csrc/gsi_eval_core.h, the same header the
 * kernels are compiled from) on the host, over a slab image produced by the library's own
 * builder + flatten code.  It exists so that the CPU test-suite (-m "not gpu") can check
 * the kernel LOGIC (exact DAG replay, walk, safety-margin acceptance, deferred list,
 * seeding) against the reference before a GPU is available.  It is not linked into
 * libgsl_scattered_cuda.so and is not an evaluation fallback.
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "gsi_builder.h"
#include "gsi_eval_core.h"
#include "gsi_flatten.h"

struct hostsim
{
  std::vector<char> image;
  gsi_slab_header hdr;
  gsi_view view;
  mutable long long memo_hits = 0; /* deferred queries answered from the vertex memo */
};

extern "C" {

hostsim *
hostsim_create(const double *xa, const double *ya, const double *za, size_t n, size_t n_fields,
               int flags, const size_t *order, double cells_per_leaf, double delta, int exact_entry,
               int tile_leaves)
{
  hostsim *h = new hostsim();
  gsi_host_tree tree;
  std::string err;
  std::vector<int64_t> ord;
  if (order)
    {
      ord.resize(n);
      for (size_t i = 0; i < n; i++) ord[i] = (int64_t) order[i];
    }
  if (gsi_build_tree(xa, ya, n, flags, order ? ord.data() : nullptr, &tree, &err))
    {
      delete h;
      return nullptr;
    }
  gsi_flatten_params prm;
  if (cells_per_leaf > 0) prm.cells_per_leaf = cells_per_leaf;
  if (delta > 0) prm.delta = delta;
  if (tile_leaves > 0) prm.tile_leaves = (uint32_t) tile_leaves;
  if (gsi_flatten(tree, xa, ya, za, n_fields, prm, &h->image, &err))
    {
      delete h;
      return nullptr;
    }
  std::memcpy(&h->hdr, h->image.data(), sizeof h->hdr);
  h->view = gsi_view_of(h->image.data(), &h->hdr);
  const size_t ncell = (size_t) h->hdr.grid_g * h->hdr.grid_g;
  /* what k_entry_cells does on the device (exact_entry = 0: every replay starts at the root) */
  int32_t *entry = (int32_t *) (h->image.data() + h->hdr.off_entry);
  if (exact_entry)
    for (size_t c = 0; c < ncell; c++) entry[c] = gsi_entry_of_cell(&h->view, c);
  else
    h->view.entry = nullptr;
  /* what k_memo_fill does on the device */
  if (h->hdr.memo_bits)
    {
      gsi_memo *memo = (gsi_memo *) (h->image.data() + h->hdr.off_memo);
      for (size_t i = 0; i <= (size_t) h->view.memo_mask; i++)
        {
          gsi_memo *e = memo + i;
          if (e->state != GSI_MEMO_PENDING) continue;
          double c0 = 0, c1 = 0;
          int last = -1;
          gsi_rec lr;
          const int start = h->view.entry ? h->view.entry[gsi_cell_of(&h->view, e->x, e->y)] : 0;
          const int node = gsi_descend(h->view.dag, &h->view.xf, start, e->x, e->y, &c0, &c1, &last, &lr, 0);
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
  /* what k_seed_cells does on the device */
  int32_t *seeds = (int32_t *) (h->image.data() + h->hdr.off_seeds);
  for (size_t c = 0; c < ncell; c++)
    {
      double qx, qy;
      gsi_cell_centre(&h->view, c, &qx, &qy);
      const int slot = gsi_exact_one(&h->view, qx, qy, nullptr, nullptr, 0, c);
      seeds[c] = slot < 0 ? 0 : slot;
    }
  /* what k_pack_seeds does on the device: base = seed of the block's centre cell - 32768; a cell whose seed
     is outside the 16-bit window takes the in-window seed of the nearest cell of its row / column */
  {
    const size_t g = h->hdr.grid_g, B = (size_t) 1 << h->hdr.seed_shift, gb = (g + B - 1) / B;
    uint16_t *s16 = (uint16_t *) (h->image.data() + h->hdr.off_seed16);
    int32_t *sb = (int32_t *) (h->image.data() + h->hdr.off_seedbase);
    const int ESC = (int) GSI_SEED16_ESC;
    for (size_t by = 0; by < gb; by++)
      for (size_t bx = 0; bx < gb; bx++)
        {
          const size_t x0 = bx * B, y0 = by * B, x1 = std::min(x0 + B, g), y1 = std::min(y0 + B, g);
          const int base = std::max(seeds[((y0 + y1) >> 1) * g + ((x0 + x1) >> 1)] - 32768, 0);
          sb[by * gb + bx] = base;
          for (size_t y = y0; y < y1; y++)
            for (size_t x = x0; x < x1; x++)
              {
                int d = seeds[y * g + x] - base;
                for (size_t r = 1; !(d >= 0 && d < ESC) && r < B; r++)
                  {
                    int c;
                    if (x >= x0 + r && (c = seeds[y * g + (x - r)] - base) >= 0 && c < ESC) d = c;
                    else if (x + r < x1 && (c = seeds[y * g + (x + r)] - base) >= 0 && c < ESC) d = c;
                    else if (y >= y0 + r && (c = seeds[(y - r) * g + x] - base) >= 0 && c < ESC) d = c;
                    else if (y + r < y1 && (c = seeds[(y + r) * g + x] - base) >= 0 && c < ESC) d = c;
                  }
                s16[y * g + x] = (d >= 0 && d < ESC) ? (uint16_t) d : (uint16_t) GSI_SEED16_ESC;
              }
        }
  }
  return h;
}

void
hostsim_free(hostsim *h)
{
  delete h;
}

/* cells of the compact seed table that escape to the int32 table */
size_t
hostsim_seed_escapes(const hostsim *h)
{
  const uint16_t *s16 = (const uint16_t *) (h->image.data() + h->hdr.off_seed16);
  const size_t cells = (size_t) h->hdr.grid_g * h->hdr.grid_g;
  size_t n = 0;
  for (size_t c = 0; c < cells; c++) n += s16[c] == GSI_SEED16_ESC;
  return n;
}
int hostsim_seed_shift(const hostsim *h) { return (int) h->hdr.seed_shift; }
long long hostsim_memo_hits(const hostsim *h) { return h->memo_hits; }
size_t hostsim_n_leaves(const hostsim *h) { return (size_t) h->hdr.n_leaves; }
size_t hostsim_grid_g(const hostsim *h) { return (size_t) h->hdr.grid_g; }
const int32_t *hostsim_tiles(const hostsim *h) { return (const int32_t *) (h->image.data() + h->hdr.off_tiles); }
int hostsim_tile_bits(const hostsim *h) { return (int) h->hdr.tile_bits; }
int hostsim_tile_max_leaves(const hostsim *h) { return (int) h->hdr.tile_max_leaves; }
const int32_t *hostsim_entry(const hostsim *h) { return (const int32_t *) (h->image.data() + h->hdr.off_entry); }
/* depth of DAG node `node` below the root is not stored; report the node type instead */
int hostsim_node_is_leaf(const hostsim *h, int node) { return (h->view.dag[node].meta & GSI_META_TYPE_MASK) == GSI_LEAF; }

/* mode 0: k_locate + k_exact_list protocol; mode 1: k_exact_all.
 * stats[0] = deferred, stats[1] = walk steps of accepted queries, stats[2] = outside cage */
int
hostsim_eval(const hostsim *h, const double *xq, const double *yq, size_t nq, double *zq, int *leaf,
             int mode, int max_steps, long long stats[3])
{
  stats[0] = stats[1] = stats[2] = 0;
  std::vector<size_t> deferred;
  for (size_t q = 0; q < nq; q++)
    {
      if (mode == 1)
        {
          if (gsi_exact_one(&h->view, xq[q], yq[q], zq, leaf, nq, q) < 0) stats[2]++;
          continue;
        }
      int steps = 0;
      if (gsi_locate_one(&h->view, xq[q], yq[q], max_steps, zq, leaf, nq, q, &steps))
        stats[1] += steps;
      else
        deferred.push_back(q);
    }
  for (size_t q : deferred)
    {
      /* the deferred-query kernels probe the vertex memo first (memo_answer in gsi_kernels.cu) */
      int mnode, mslot;
      if (gsi_memo_find(&h->view, xq[q], yq[q], &mnode, &mslot))
        {
          const gsi_leaf *lf = gsi_leaf_at(&h->view, mslot);
          gsi_rec r;
          gsi_fetch_rec(&r, lf);
          double c0, c1;
          gsi_solve(&h->view.xf, &r, (int) (r.meta >> 31), xq[q], yq[q], &c0, &c1);
          gsi_leaf_tail t;
          gsi_fetch_tail(&t, lf);
          if (leaf) leaf[q] = mnode;
          if (zq) gsi_blend_store(&h->view, &t, c0, c1, zq, nq, q);
          h->memo_hits++;
        }
      else if (gsi_exact_one(&h->view, xq[q], yq[q], zq, leaf, nq, q) < 0)
        stats[2]++;
    }
  stats[0] = (long long) deferred.size();
  return 0;
}

/* what k_deriv does on the device: gradient of every field for the leaf ids of a locate pass */
int
hostsim_deriv(const hostsim *h, const int *leaf, size_t nq, double *dzdx, double *dzdy)
{
  for (size_t q = 0; q < nq; q++) gsi_deriv_one(&h->view, leaf[q], dzdx, dzdy, nq, q);
  return 0;
}

/* experiment: how deep is the certified entry node of a LEAF's inflated bounding box, compared with the
 * depth of the leaf itself (certified descent of a degenerate box at the centroid)? */
int
hostsim_leaf_entry_depths(const hostsim *h, double grow, int *depth_entry, int *depth_leaf, int *entry_node)
{
  const gsi_view *V = &h->view;
  for (int s = 0; s < V->n_leaves; s++)
    {
      const gsi_leaf *lf = gsi_leaf_at(V, s);
      gsi_rec r = lf->r;
      double cx[4], cy[4];
      if (!(r.u00 != 0 && r.u11 != 0)) { depth_entry[s] = depth_leaf[s] = -1; entry_node[s] = 0; continue; }
      gsi_leaf_box(V, &r, (int) (r.meta >> 31), grow, cx, cy);
      entry_node[s] = gsi_entry_of_box(V, cx, cy, &depth_entry[s]);
      /* centroid */
      double mx = 0.25 * (cx[0] + cx[1] + cx[2] + cx[3]), my = 0.25 * (cy[0] + cy[1] + cy[2] + cy[3]);
      double px[4] = {mx, mx, mx, mx}, py[4] = {my, my, my, my};
      gsi_entry_of_box(V, px, py, &depth_leaf[s]);
    }
  return 0;
}

/* The builder's batched-locate protocol (gsi_build_locator) with the locator
 * running on the host over a copy of each snapshot: must give the tree of the plain build node for
 * node.  Returns 1 = identical, 0 = different, < 0 = a build failed. */
int
hostsim_build_batched_equal(const double *xa, const double *ya, size_t n, int flags, const size_t *order,
                            size_t *batched_points, size_t *rounds)
{
  std::vector<int64_t> ord;
  if (order)
    {
      ord.resize(n);
      for (size_t i = 0; i < n; i++) ord[i] = (int64_t) order[i];
    }
  gsi_host_tree a, b;
  std::string err;
  if (gsi_build_tree(xa, ya, n, flags, order ? ord.data() : nullptr, &a, &err)) return -1;
  gsi_build_locator *loc = gsi_make_host_snapshot_locator();
  const int st = gsi_build_tree(xa, ya, n, flags, order ? ord.data() : nullptr, &b, &err, loc);
  delete loc;
  if (st) return -2;
  if (batched_points) *batched_points = b.batched_points;
  if (rounds) *rounds = b.batched_rounds;
  return a.pts == b.pts && a.lnk == b.lnk && a.type == b.type && a.shuffle == b.shuffle;
}
}
