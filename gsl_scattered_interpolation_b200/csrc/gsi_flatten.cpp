/* gsi_flatten.cpp -- host tree -> device slab image (see gsi_layout.h).
 *
 * Resolves the reference's indirections once (POINT -> shuffle -> data row,
 * linear_simplex.h:82-93), compacts the leaves of the history DAG, orders leaves and
 * vertices along a Morton curve, and precomputes the query-independent LU half of
 * calculate_bary_coords for every DAG node and leaf (linear_simplex.c:616-639).
 */
#include "gsi_flatten.h"
#include "gsi_eval_core.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace
{

inline uint64_t
align256(uint64_t x)
{
  return (x + 255) & ~(uint64_t) 255;
}

inline uint32_t
part1by1(uint32_t x) /* spread the low 16 bits */
{
  x &= 0x0000ffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

inline uint32_t
morton16(double fu, double fv) /* fu, fv in [0,1] */
{
  auto q = [](double f) -> uint32_t {
    if (!(f > 0)) return 0;
    if (f >= 1) return 65535u;
    return (uint32_t) (f * 65536.0);
  };
  return part1by1(q(fu)) | (part1by1(q(fv)) << 1);
}

inline float
float_up(double x) /* smallest float >= x */
{
  float f = (float) x;
  if ((double) f < x) f = std::nextafter(f, std::numeric_limits<float>::infinity());
  return f;
}

} // namespace

int
gsi_flatten(const gsi_host_tree &t, const double *xa, const double *ya, const double *za,
            size_t n_fields, const gsi_flatten_params &prm, std::vector<char> *image,
            std::string *err)
{
  const size_t n = t.n_points, N = t.n_nodes();
  if (N == 0)
    {
      if (err) *err = "empty tree";
      return 4;
    }
  if (n_fields == 0) n_fields = 1;
  const gsi_xform &xf = t.xf;

  /* --- data bounding box in standardised coordinates (grid domain) ------------------- */
  double umin = -0.5, umax = 0.5, vmin = -0.5, vmax = 0.5;
  if (n > 0)
    {
      umin = vmin = std::numeric_limits<double>::infinity();
      umax = vmax = -umin;
      for (size_t i = 0; i < n; i++)
        {
          const double u = xf.sx * (xa[i] - xf.cx), v = xf.sy * (ya[i] - xf.cy);
          umin = std::min(umin, u); umax = std::max(umax, u);
          vmin = std::min(vmin, v); vmax = std::max(vmax, v);
        }
      if (!(umax > umin)) { umin -= 0.5; umax += 0.5; }
      if (!(vmax > vmin)) { vmin -= 0.5; vmax += 0.5; }
    }
  const double du = umax - umin, dv = vmax - vmin;

  /* --- leaves, Morton order ------------------------------------------------------------ */
  std::vector<int32_t> leaf_nodes;
  leaf_nodes.reserve(2 * n + 8);
  for (size_t i = 0; i < N; i++)
    if (t.type[i] == GSI_LEAF) leaf_nodes.push_back((int32_t) i);
  const size_t L = leaf_nodes.size();

  auto std_xy = [&](int32_t p, double *u, double *v) {
    double x, y;
    gsi_vertex_xy(t, xa, ya, p, &x, &y);
    *u = xf.sx * (x - xf.cx);
    *v = xf.sy * (y - xf.cy);
  };

  std::vector<std::pair<uint32_t, int32_t>> keyed(L);
  for (size_t k = 0; k < L; k++)
    {
      const size_t node = (size_t) leaf_nodes[k];
      double cu = 0, cv = 0;
      for (int i = 0; i < 3; i++)
        {
          double u, v;
          std_xy(t.pts[3 * node + i], &u, &v);
          /* cage vertices are ~2700 units away: clamp so cage leaves sort next to the hull */
          cu += std::min(std::max(u, umin - du), umax + du);
          cv += std::min(std::max(v, vmin - dv), vmax + dv);
        }
      keyed[k] = std::make_pair(morton16((cu / 3 - umin) / du, (cv / 3 - vmin) / dv),
                                (int32_t) node);
    }
  std::sort(keyed.begin(), keyed.end());
  std::vector<int32_t> slot_of_node(N, -1);
  for (size_t k = 0; k < L; k++) slot_of_node[(size_t) keyed[k].second] = (int32_t) k;

  /* --- vertices, Morton order ---------------------------------------------------------- */
  std::vector<std::pair<uint32_t, int64_t>> vkey(n);
  for (size_t i = 0; i < n; i++)
    {
      const double u = xf.sx * (xa[i] - xf.cx), v = xf.sy * (ya[i] - xf.cy);
      vkey[i] = std::make_pair(morton16((u - umin) / du, (v - vmin) / dv), (int64_t) i);
    }
  std::sort(vkey.begin(), vkey.end());
  std::vector<int32_t> vslot_of_row(n);
  for (size_t s = 0; s < n; s++) vslot_of_row[(size_t) vkey[s].second] = (int32_t) s;

  /* --- grid ---------------------------------------------------------------------------- */
  uint32_t G = prm.grid_g;
  if (G == 0)
    {
      const double want = std::sqrt(std::max(1.0, prm.cells_per_leaf * (double) L));
      G = (uint32_t) std::min(16384.0, std::max(1.0, std::ceil(want)));
      /* large grids: a multiple of eight seed blocks per side, so that the coarse Morton boundaries
         (where neighbouring cells' seeds are more than 64K slots apart) fall between blocks */
      if (G >= 256)
        {
          const uint32_t B = 1u << gsi_seed_shift_for(G);
          uint32_t gb = ((G + 4 * B) / (8 * B)) * 8;
          if (gb < 8) gb = 8;
          if (gb > GSI_SEED_GB_MAX) gb = GSI_SEED_GB_MAX;
          G = gb * B;
        }
    }
  const uint32_t seed_shift = gsi_seed_shift_for(G);

  /* --- sizes --------------------------------------------------------------------------- */
  gsi_slab_header h;
  std::memset(&h, 0, sizeof h);
  h.magic = GSI_SLAB_MAGIC;
  h.version = GSI_SLAB_VERSION;
  h.n_fields = (uint32_t) n_fields;
  h.n_points = n;
  h.n_dag = N;
  h.n_leaves = L;
  h.grid_g = G;
  h.seed_shift = seed_shift;
  h.xf = xf;
  h.u0 = umin; h.v0 = vmin;
  h.inv_du = (double) G / du;
  h.inv_dv = (double) G / dv;
  h.delta = prm.delta;
  uint64_t off = align256(sizeof h);
  h.off_dag = off;   off = align256(off + N * sizeof(gsi_rec));
  const size_t leaf_stride = prm.leaf_stride == 128 ? 128 : sizeof(gsi_leaf);
  h.leaf_stride = (uint32_t) leaf_stride;
  h.off_leaf = off;  off = align256(off + L * leaf_stride);
  h.off_z = off;     off = align256(off + std::max<size_t>(1, n) * n_fields * sizeof(double));
  h.off_seeds = off; off = align256(off + (uint64_t) G * G * sizeof(int32_t));
  h.off_entry = off; off = align256(off + (uint64_t) G * G * sizeof(int32_t));
  h.off_seed16 = off; off = align256(off + (uint64_t) G * G * sizeof(uint16_t));
  {
    const uint64_t gb = gsi_seed_blocks(G, seed_shift);
    h.off_seedbase = off; off = align256(off + gb * gb * sizeof(int32_t));
  }
  /* Morton-prefix tiles of about tile_leaves leaves: P bits, 2 <= P <= 12 */
  uint32_t P = 2;
  while (P < 12 && (double) L / (double) (1u << P) > (double) std::max<uint32_t>(1, prm.tile_leaves) * 1.05) P++;
  h.tile_bits = P;
  h.off_tiles = off; off = align256(off + ((uint64_t) (1u << P) + 1) * sizeof(int32_t));
  /* vertex memo: a power of two of at least 2 n entries */
  uint32_t memo_bits = 0;
  if (prm.vertex_memo && n > 0 && xa && ya)
    {
      memo_bits = 4;
      while (memo_bits < 31 && ((uint64_t) 1 << memo_bits) < 2 * (uint64_t) n) memo_bits++;
      h.off_memo = off; off = align256(off + ((uint64_t) 1 << memo_bits) * sizeof(gsi_memo));
    }
  h.memo_bits = memo_bits;
  h.total_bytes = off;

  image->assign((size_t) h.total_bytes, 0);
  char *base = image->data();
  std::memcpy(base, &h, sizeof h);
  gsi_rec *dag = (gsi_rec *) (base + h.off_dag);
  char *leaf_base = base + h.off_leaf;
  double *z = (double *) (base + h.off_z);

  /* --- DAG records --------------------------------------------------------------------- */
  for (size_t i = 0; i < N; i++)
    {
      double a[2], b[2], x0[2];
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i], &a[0], &a[1]);
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i + 1], &b[0], &b[1]);
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i + 2], &x0[0], &x0[1]);
      gsi_rec &r = dag[i];
      const uint32_t m = gsi_make_lu(&xf, a[0], a[1], b[0], b[1], x0[0], x0[1], &r);
      r.meta = m | t.type[i];
      if (t.type[i] == GSI_LEAF)
        {
          r.lnk[0] = slot_of_node[i];
          r.lnk[1] = r.lnk[2] = 0;
        }
      else
        {
          r.lnk[0] = t.lnk[3 * i];
          r.lnk[1] = t.lnk[3 * i + 1];
          r.lnk[2] = t.lnk[3 * i + 2];
        }
    }

  /* --- leaf records -------------------------------------------------------------------- */
  for (size_t k = 0; k < L; k++)
    {
      const size_t node = (size_t) keyed[k].second;
      gsi_rec r = dag[node];
      const uint32_t m = r.meta;
      double tau;
      if (m & GSI_META_SINGULAR)
        tau = std::numeric_limits<double>::infinity();
      else
        {
          /* margin: every barycentric weight must exceed delta * Lmax / |det| so that the
             query is at least `delta` (standardised units) away from every edge line of
             the leaf -- see DESIGN.md "why the fast path is exact" */
          double u[3], v[3];
          for (int i = 0; i < 3; i++) std_xy(t.pts[3 * node + i], &u[i], &v[i]);
          double lmax = 0;
          for (int i = 0; i < 3; i++)
            {
              const int j = (i + 1) % 3;
              lmax = std::max(lmax, std::hypot(u[i] - u[j], v[i] - v[j]));
            }
          const double det = std::fabs(r.u00 * r.u11);
          tau = prm.delta * lmax / det;
          if (!(tau >= prm.tau_floor)) tau = prm.tau_floor; /* also catches NaN */
          if (!(tau < 0.25)) tau = std::numeric_limits<double>::infinity();
        }
      float tf = float_up(tau);
      uint32_t bits;
      std::memcpy(&bits, &tf, 4);
      if (m & GSI_META_SWAP) bits |= 0x80000000u;
      r.meta = bits;
      gsi_leaf_tail tail;
      std::memset(&tail, 0, sizeof tail);
      for (int i = 0; i < 3; i++)
        {
          const int32_t nb = t.lnk[3 * node + i];
          r.lnk[i] = nb ? slot_of_node[(size_t) nb] : -1;
          const int32_t p = t.pts[3 * node + i];
          if (n_fields == 1)
            {
              if (p >= 0)
                {
                  tail.s.vflags |= 1u << i;
                  tail.s.z[i] = za ? za[(size_t) t.shuffle[(size_t) p]] : 0.0;
                }
            }
          else
            tail.g.v[i] = p < 0 ? -1 : vslot_of_row[(size_t) t.shuffle[(size_t) p]];
        }
      tail.g.node = (int32_t) node; /* same offset in both views */
      gsi_leaf lf;
      lf.r = r;
      lf.t = tail;
      std::memcpy(leaf_base + k * leaf_stride, &lf, sizeof lf);
    }

  /* --- tiles: first leaf slot of every Morton prefix ------------------------------------ */
  {
    int32_t *tiles = (int32_t *) (base + h.off_tiles);
    const uint32_t nt = 1u << h.tile_bits, sh = 32 - h.tile_bits;
    size_t k = 0;
    uint32_t biggest = 0;
    for (uint32_t tl = 0; tl < nt; tl++)
      {
        while (k < L && (keyed[k].first >> sh) < tl) k++;
        tiles[tl] = (int32_t) k;
        if (tl) biggest = std::max<uint32_t>(biggest, (uint32_t) (tiles[tl] - tiles[tl - 1]));
      }
    tiles[nt] = (int32_t) L;
    biggest = std::max<uint32_t>(biggest, (uint32_t) (tiles[nt] - tiles[nt - 1]));
    gsi_slab_header *hp = (gsi_slab_header *) base;
    hp->tile_max_leaves = biggest;
  }

  /* --- vertex memo: the keys (answers are filled in on the device) ----------------------- */
  if (memo_bits)
    {
      gsi_memo *memo = (gsi_memo *) (base + h.off_memo);
      const uint32_t mask = ((uint32_t) 1 << memo_bits) - 1u;
      for (size_t i = 0; i < n; i++)
        {
          if (!(std::isfinite(xa[i]) && std::isfinite(ya[i]))) continue;
          const uint64_t xb = gsi_f64_bits(xa[i]), yb = gsi_f64_bits(ya[i]);
          uint32_t k = gsi_memo_hash(xb, yb) & mask;
          for (int probe = 0; probe < GSI_MEMO_MAX_PROBES; probe++, k = (k + 1u) & mask)
            {
              gsi_memo &e = memo[k];
              if (e.state == GSI_MEMO_EMPTY)
                {
                  e.state = GSI_MEMO_PENDING;
                  e.x = xa[i];
                  e.y = ya[i];
                  break;
                }
              if (gsi_f64_bits(e.x) == xb && gsi_f64_bits(e.y) == yb) break; /* duplicate node */
            }
        }
    }

  /* --- values -------------------------------------------------------------------------- */
  if (za)
    for (size_t s = 0; s < n; s++)
      {
        const size_t row = (size_t) vkey[s].second;
        for (size_t f = 0; f < n_fields; f++) z[s * n_fields + f] = za[f * n + row];
      }
  return 0;
}
