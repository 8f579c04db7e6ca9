/* gsi_builder.cpp -- host builder of the history DAG ("simplex tree"), dim = 2.
 * See gsi_builder.h for the reference functions this restates.  Compiled with
 * -ffp-contract=off: every floating-point expression below is written in the reference's
 * operation order and must round operation by operation.
 */
#include "gsi_builder.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <utility>

namespace
{

const double kDblEps = 2.2204460492503131e-16;      /* GSL_DBL_EPSILON, gsl_machine.h:17 */
const double kRoot5DblEps = 7.4009597974140505e-04; /* GSL_ROOT5_DBL_EPSILON, gsl_machine.h:21 */

/* ---- gsl_rng_mt19937 (rng/mt.c:64-151, the 2002 seeding) ------------------------------- */
struct mt19937
{
  uint32_t mt[624];
  int mti;

  explicit mt19937(unsigned long s)
  {
    if (s == 0) s = 4357; /* rng/mt.c:136-137 */
    mt[0] = (uint32_t) (s & 0xffffffffUL);
    for (int i = 1; i < 624; i++)
      mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t) i;
    mti = 624;
  }

  uint32_t get()
  {
    if (mti >= 624)
      {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++)
          {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
          }
        for (; kk < 623; kk++)
          {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
          }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mti = 0;
      }
    uint32_t k = mt[mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680u;
    k ^= (k << 15) & 0xefc60000u;
    k ^= (k >> 18);
    return k;
  }

  /* gsl_rng_uniform_int, rng/gsl_rng.h:188-211 (min = 0, max = 0xffffffff) */
  unsigned long uniform_int(unsigned long n)
  {
    const unsigned long range = 0xffffffffUL;
    const unsigned long scale = range / n;
    unsigned long k;
    do
      k = (unsigned long) get() / scale;
    while (k >= n);
    return k;
  }
};

/* ---- 2x2 LU with partial pivoting + solve (linalg/lu.c:59-201, source_trsv_r.h:33-79) -- */
struct lu2
{
  double u00, u01, l10, u11;
  bool swap, singular;
};

inline lu2
lu2_decomp(double m00, double m01, double m10, double m11)
{
  lu2 r;
  r.swap = std::fabs(m10) > std::fabs(m00);
  if (r.swap)
    {
      std::swap(m00, m10);
      std::swap(m01, m11);
    }
  if (m00 != 0.0)
    {
      m10 = m10 / m00;
      m11 = m11 - m10 * m01;
    }
  r.u00 = m00; r.u01 = m01; r.l10 = m10; r.u11 = m11;
  r.singular = (m00 == 0) || (m11 == 0);
  return r;
}

inline void
lu2_solve(const lu2 &f, double b0, double b1, double *x0, double *x1)
{
  if (f.swap) std::swap(b0, b1);
  b1 = b1 - f.l10 * b0;
  b1 = b1 / f.u11;
  b0 = (b0 - f.u01 * b1) / f.u00;
  *x0 = b0;
  *x1 = b1;
}

/* ---- cblas level 1 for N = 2, unit stride ---------------------------------------------- */
inline double
nrm2_2(const double x[2]) /* cblas/source_nrm2_r.h */
{
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < 2; i++)
    {
      const double v = x[i];
      if (v != 0.0)
        {
          const double ax = std::fabs(v);
          if (scale < ax)
            {
              ssq = 1.0 + ssq * (scale / ax) * (scale / ax);
              scale = ax;
            }
          else
            ssq += (ax / scale) * (ax / scale);
        }
    }
  return scale * std::sqrt(ssq);
}

inline double
dot_2(const double x[2], const double y[2]) /* cblas/source_dot_r.h */
{
  double r = 0.0;
  r += x[0] * y[0];
  r += x[1] * y[1];
  return r;
}

/* live view of the growing tree for the descent: leaf links mutate while the tree grows
 * (children links of internal nodes are final), so records are assembled on the fly. */
struct tree_view
{
  const std::vector<gsi_rec> *rec;
  const std::vector<int32_t> *lnk;
  const std::vector<uint8_t> *type;
  gsi_rec rec_at(size_t i) const
  {
    gsi_rec r = (*rec)[i];
    r.lnk[0] = (*lnk)[3 * i];
    r.lnk[1] = (*lnk)[3 * i + 1];
    r.lnk[2] = (*lnk)[3 * i + 2];
    r.meta = (r.meta & ~GSI_META_TYPE_MASK) | (*type)[i];
    return r;
  }
};

#define GSI_FETCH_VIEW(dst, idx) (dst) = dag->rec_at((size_t) (idx))
GSI_DEFINE_DESCEND(descend_view, static inline, const tree_view *, GSI_FETCH_VIEW)

struct frame
{
  int leaf, ismplx, i;
};

class builder
{
public:
  builder(gsi_host_tree &t, std::string *err) : T(t), err_(err) {}
  int run(const double *xa, const double *ya, size_t n, int flags, const int64_t *order,
          gsi_build_locator *locator);

private:
  gsi_host_tree &T;
  std::string *err_;
  std::vector<double> px_, py_; /* raw coordinates in insertion (= point id) order */
  std::vector<gsi_rec> rec_;    /* LU cache per node (transparent, see gsi_core.h) */
  double c0_ = 0, c1_ = 0;      /* accel->coords */
  int new_simplexes_[3] = {0, 0, 0};
  int left_out_[2] = {0, 0};
  int n_inserted_ = 0;

  int fail(const char *msg)
  {
    if (err_ && err_->empty()) *err_ = msg;
    return 5; /* GSL_EFAILED */
  }

  int32_t &POINT(int node, int i) { return T.pts[3 * (size_t) node + i]; }
  int32_t &LINK(int node, int i) { return T.lnk[3 * (size_t) node + i]; }
  bool LEAF(int node) const { return T.type[node] == GSI_LEAF; }

  void P(int32_t id, double v[2]) const
  {
    if (id < 0)
      {
        v[0] = T.seed[2 * (-id - 1)];
        v[1] = T.seed[2 * (-id - 1) + 1];
      }
    else
      {
        v[0] = px_[id];
        v[1] = py_[id];
      }
  }

  int node_alloc() /* simplex_tree_node_alloc, linear_simplex.c:17-51 */
  {
    const int id = (int) T.type.size();
    T.type.push_back(GSI_LEAF);
    T.pts.insert(T.pts.end(), 3, 0);
    T.lnk.insert(T.lnk.end(), 3, 0);
    rec_.push_back(gsi_rec());
    return id;
  }

  void make_rec(int node)
  {
    double a[2], b[2], x0[2];
    P(POINT(node, 0), a);
    P(POINT(node, 1), b);
    P(POINT(node, 2), x0);
    rec_[node].meta = gsi_make_lu(&T.xf, a[0], a[1], b[0], b[1], x0[0], x0[1], &rec_[node]);
  }

  bool point_in_simplex(int node, int32_t point) /* linear_simplex.c:318-329 */
  {
    return POINT(node, 0) == point || POINT(node, 1) == point || POINT(node, 2) == point;
  }

  int find_link(int node, int target) /* the FIND macro, linear_simplex.h:95-102 */
  {
    for (int j = 0; j < 3; j++)
      if (LINK(node, j) == target) return j;
    return -1;
  }

  int insert_point(int leaf);
  bool in_hypersphere(int node, int32_t idx);
  bool flippable(int leaf, int face, int neighbor, int far);
  int set_external_links(const int old_neighbors[2], int neighbor, int neighbor_set,
                         int new_simplex, int32_t point_left_out);
  int try_flip(int leaf, int face, bool *flipped);
  int delaunay(int leaf, int face);
};

/* in_hypersphere_points + calculate_hypersphere_points, linear_simplex.c:507-605 */
bool
builder::in_hypersphere(int node, int32_t idx)
{
  const double s[2] = {T.xf.sx, T.xf.sy}, sh[2] = {T.xf.cx, T.xf.cy};
  double v[3][2], pt[2];
  for (int i = 0; i < 3; i++) P(POINT(node, i), v[i]);
  P(idx, pt);

  double coords[2], M[2][2];
  for (int i = 0; i < 2; i++)
    {
      coords[i] = 0;
      for (int j = 0; j < 2; j++)
        {
          const double pij = s[j] * (v[i][j] - sh[j]);
          const double pij1 = s[j] * (v[i + 1][j] - sh[j]);
          coords[i] = coords[i] + pij * pij - pij1 * pij1;
          M[i][j] = pij - pij1;
        }
      coords[i] = 0.5 * coords[i];
    }
  const lu2 f = lu2_decomp(M[0][0], M[0][1], M[1][0], M[1][1]);
  if (f.singular) return true; /* degenerate points: "assume inside", :517-521 */

  double x0[2];
  lu2_solve(f, coords[0], coords[1], &x0[0], &x0[1]);

  double r2 = 0;
  for (int j = 0; j < 2; j++)
    {
      double c = v[0][j];
      c = c - sh[j];
      c = c * s[j];
      c = c - x0[j];
      r2 += c * c;
    }
  double dist2 = 0;
  for (int j = 0; j < 2; j++)
    {
      const double comp = s[j] * (pt[j] - sh[j]);
      const double val = comp - x0[j];
      dist2 += val * val;
    }
  return dist2 < (r2 * (1 - 10 * kDblEps));
}

/* flippable + orthonormalize, edge_flip.c:39-95, linear_simplex_util.h:43-70 (raw coords) */
bool
builder::flippable(int leaf, int face, int neighbor, int far)
{
  double p_face[2], p_far[2];
  P(POINT(leaf, face), p_face);
  P(POINT(neighbor, far), p_far);
  bool ok = true;
  for (int ismplx = 0; ismplx < 2; ismplx++)
    {
      double m[2][2], p[2];
      /* row 0: the face vertex that is not left out; row 1: the left-out vertex */
      P(POINT(leaf, left_out_[1 - ismplx]), p);
      m[0][0] = p[0] - p_face[0];
      m[0][1] = p[1] - p_face[1];
      P(POINT(leaf, left_out_[ismplx]), p);
      m[1][0] = p[0] - p_face[0];
      m[1][1] = p[1] - p_face[1];

      /* orthonormalize(tmp_mat) */
      double scale = -1;
      {
        double mag = nrm2_2(m[0]);
        if (scale < mag) scale = mag;
        if (mag < scale * 100 * kDblEps) return true;
        const double alpha = 1 / mag;
        m[0][0] *= alpha;
        m[0][1] *= alpha;
        const double proj = dot_2(m[0], m[1]);
        const double a = -proj;
        if (a != 0.0) /* cblas daxpy returns early for alpha == 0 */
          {
            m[1][0] += a * m[0][0];
            m[1][1] += a * m[0][1];
          }
        mag = nrm2_2(m[1]);
        if (scale < mag) scale = mag;
        if (mag < scale * 100 * kDblEps) return true;
        const double beta = 1 / mag;
        m[1][0] *= beta;
        m[1][1] *= beta;
      }
      double v[2] = {p_far[0] - p_face[0], p_far[1] - p_face[1]};
      const double proj = dot_2(m[1], v);
      ok = ok && (proj > 0);
      if (!ok) break;
    }
  return ok;
}

/* set_external_links, edge_flip.c:149-183 */
int
builder::set_external_links(const int old_neighbors[2], int neighbor, int neighbor_set,
                            int new_simplex, int32_t point_left_out)
{
  int jsmplx, null_option = 0;
  for (jsmplx = 0; jsmplx < 2; jsmplx++)
    {
      if (!old_neighbors[jsmplx])
        {
          null_option = 1;
          continue;
        }
      if (!point_in_simplex(old_neighbors[jsmplx], point_left_out)) break;
    }
  if (!(jsmplx < 2 || null_option)) return fail("no proper external link found");
  int ext;
  if (jsmplx < 2)
    ext = LINK(new_simplex, neighbor_set) = old_neighbors[jsmplx];
  else
    ext = LINK(new_simplex, neighbor_set) = 0;
  if (ext)
    {
      const int j = find_link(ext, neighbor);
      if (j < 0) return fail("no reverse link found");
      LINK(ext, j) = new_simplex;
    }
  return 0;
}

/* the non-recursive part of delaunay(), edge_flip.c:211-303 */
int
builder::try_flip(int leaf, int face, bool *flipped)
{
  *flipped = false;
  if (!leaf) return fail("delaunay() on the initial simplex");
  if (!LINK(leaf, face)) return 0;
  if (!LEAF(leaf)) return fail("checking if a flip is necessary on a non-leaf simplex");
  const int neighbor = LINK(leaf, face);
  if (!LEAF(neighbor)) return fail("neighbor to leaf not a leaf");
  const int far = find_link(neighbor, leaf);
  if (far < 0) return fail("reverse link not found");

  if (!in_hypersphere(leaf, POINT(neighbor, far))) return 0;

  /* set_left_out, edge_flip.c:17-35: the two face vertices in index order */
  for (int i = 0, k = 0; i < 3; i++)
    if (i != face) left_out_[k++] = i;
  if (!flippable(leaf, face, neighbor, far)) return 0;
  if (!in_hypersphere(neighbor, POINT(leaf, face))) return fail("in_hypersphere not reciprocal");

  T.type[leaf] = GSI_SUB_D;
  T.type[neighbor] = GSI_SUB_D;

  /* save_current_neighbors, edge_flip.c:98-114 */
  int old1[2], old2[2], k = 0, ok = 0;
  for (int i = 0; i < 3; i++)
    if (LINK(leaf, i) != neighbor) { if (k < 2) old1[k] = LINK(leaf, i); k++; }
    else ok++;
  if (ok != 1) return fail("didn't find one and only one reverse link");
  k = 0; ok = 0;
  for (int i = 0; i < 3; i++)
    if (LINK(neighbor, i) != leaf) { if (k < 2) old2[k] = LINK(neighbor, i); k++; }
    else ok++;
  if (ok != 1) return fail("didn't find one and only one reverse link");

  for (int s = 0; s < 2; s++) new_simplexes_[s] = node_alloc();

  /* set_points, edge_flip.c:117-146 */
  for (int s = 0; s < 2; s++)
    {
      const int ns = new_simplexes_[s];
      POINT(ns, 0) = POINT(leaf, face);
      POINT(ns, 1) = POINT(neighbor, far);
      POINT(ns, 2) = POINT(leaf, left_out_[1 - s]);
      make_rec(ns);
    }
  for (int s = 0; s < 2; s++)
    {
      int st = set_external_links(old2, neighbor, 0, new_simplexes_[s], POINT(leaf, left_out_[s]));
      if (st) return st;
      st = set_external_links(old1, leaf, 1, new_simplexes_[s], POINT(leaf, left_out_[s]));
      if (st) return st;
    }
  /* set_internal_links, edge_flip.c:186-207 (i = 2 only when dim = 2) */
  for (int s = 0; s < 2; s++)
    {
      int js;
      for (js = 0; js < 2; js++)
        {
          if (s == js) continue;
          if (!point_in_simplex(new_simplexes_[js], POINT(new_simplexes_[s], 2))) break;
        }
      LINK(new_simplexes_[s], 2) = new_simplexes_[js]; /* js == 2 reads the stale third slot */
    }
  for (int i = 0; i < 2; i++)
    {
      LINK(leaf, i) = new_simplexes_[i];
      LINK(neighbor, i) = new_simplexes_[i];
    }
  LINK(leaf, 2) = neighbor;
  LINK(neighbor, 2) = leaf;
  *flipped = true;
  return 0;
}

/* delaunay(), edge_flip.c:211-320, with the recursion of :307-316 unrolled onto an explicit
 * stack (same visiting order; the loop conditions are re-read after every nested flip). */
int
builder::delaunay(int leaf0, int face0)
{
  bool flipped;
  int st = try_flip(leaf0, face0, &flipped);
  if (st || !flipped) return st;
  std::vector<frame> stack;
  stack.push_back(frame{leaf0, 0, 0});
  while (!stack.empty())
    {
      frame &f = stack.back();
      if (f.ismplx >= 2)
        {
          stack.pop_back();
          continue;
        }
      const int child = LINK(f.leaf, f.ismplx);
      if (f.i >= 3 || !LEAF(child))
        {
          f.ismplx++;
          f.i = 0;
          continue;
        }
      const int i = f.i++;
      if (!LINK(child, i)) continue;
      st = try_flip(child, i, &flipped);
      if (st) return st;
      if (flipped) stack.push_back(frame{child, 0, 0});
    }
  return 0;
}

/* insert_point, linear_simplex.c:404-492 */
int
builder::insert_point(int leaf)
{
  if (!LEAF(leaf)) return fail("you can only insert a point into a leaf");
  T.type[leaf] = GSI_SUB_DPLUS1;
  for (int s = 0; s < 3; s++) new_simplexes_[s] = node_alloc();
  for (int i = 0; i < 3; i++)
    {
      const int ns = new_simplexes_[i];
      POINT(ns, 0) = n_inserted_;
      int k = 1;
      for (int j = 0; j < 3; j++)
        {
          if (j == i) continue;
          POINT(ns, k++) = POINT(leaf, j);
        }
      make_rec(ns);
    }
  for (int i = 0; i < 3; i++)
    {
      const int neighbor = LINK(leaf, i);
      if (neighbor && !LEAF(neighbor)) return fail("inconsistent tree: neighbor not a leaf");
      if (neighbor && point_in_simplex(neighbor, POINT(leaf, i)))
        return fail("inconsistent tree: opposite vertex inside neighbor");
      LINK(new_simplexes_[i], 0) = neighbor;
      if (neighbor)
        {
          const int j = find_link(neighbor, leaf);
          if (j < 0) return fail("no reverse link");
          LINK(neighbor, j) = new_simplexes_[i];
        }
    }
  for (int s = 0; s < 3; s++)
    for (int i = 1; i < 3; i++)
      {
        int js;
        for (js = 0; js < 3; js++)
          {
            if (s == js) continue;
            if (!point_in_simplex(new_simplexes_[js], POINT(new_simplexes_[s], i))) break;
          }
        if (js >= 3) return fail("proper internal simplex not found");
        LINK(new_simplexes_[s], i) = new_simplexes_[js];
      }
  for (int i = 0; i < 3; i++) LINK(leaf, i) = new_simplexes_[i];
  n_inserted_++;

  for (int i = 0; i < 3; i++)
    {
      if (!LEAF(LINK(leaf, i))) continue;
      const int st = delaunay(LINK(leaf, i), 0);
      if (st) return st;
    }
  return 0;
}

/* simplex_tree_init, linear_simplex.c:134-296 (min == max == NULL) */
int
builder::run(const double *xa, const double *ya, size_t n, int flags, const int64_t *order,
             gsi_build_locator *locator)
{
  double mn[2], mx[2];
  if (flags & GSI_FLAG_NOSTANDARDIZE || n == 0)
    {
      mn[0] = mn[1] = -0.5;
      mx[0] = mx[1] = +0.5;
    }
  else
    {
      mn[0] = mx[0] = xa[0];
      mn[1] = mx[1] = ya[0];
      for (size_t i = 1; i < n; i++)
        {
          if (xa[i] < mn[0]) mn[0] = xa[i];
          if (xa[i] > mx[0]) mx[0] = xa[i];
          if (ya[i] < mn[1]) mn[1] = ya[i];
          if (ya[i] > mx[1]) mx[1] = ya[i];
        }
    }
  double shift[2], scale[2];
  for (int i = 0; i < 2; i++)
    {
      shift[i] = (mn[i] + mx[i]) / 2.0;
      if (mx[i] - mn[i] <= 0)
        scale[i] = 1.0;
      else
        scale[i] = 1.0 / (mx[i] - mn[i]);
    }
  if (!(flags & GSI_FLAG_NOSTANDARDIZE) && (flags & GSI_FLAG_ISOSCALE))
    {
      double min_comp = scale[0];
      if (min_comp > scale[1]) min_comp = scale[1];
      scale[0] = scale[1] = min_comp;
    }
  T.xf.sx = scale[0]; T.xf.sy = scale[1];
  T.xf.cx = shift[0]; T.xf.cy = shift[1];

  /* regular simplex cage, linear_simplex.c:215-260; entry (0,1) is never written by the
     reference and is taken as 0 (what it reads from fresh zeroed memory) */
  double sp[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  for (int i = 0; i < 2; i++)
    {
      double tot2 = 0;
      for (int j = 0; j < i; j++)
        {
          const double comp = sp[i][j];
          tot2 += comp * comp;
        }
      const double chosen = std::sqrt(1 - tot2);
      sp[i][i] = chosen;
      const double others = -(1.0 / 2 + tot2) / chosen;
      for (int j = i + 1; j < 3; j++) sp[j][i] = others;
    }
  const double altitude = sp[0][0] - sp[1][0];
  const double radius = altitude / 3;
  const double factor = 1 / (kRoot5DblEps * radius);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 2; j++) sp[i][j] *= factor;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 2; j++)
      {
        sp[i][j] /= scale[j];
        sp[i][j] += shift[j];
        T.seed[2 * i + j] = sp[i][j];
      }

  T.n_points = n;
  T.pts.clear(); T.lnk.clear(); T.type.clear(); rec_.clear();
  T.pts.reserve(27 * n + 3); T.lnk.reserve(27 * n + 3); T.type.reserve(9 * n + 1);
  rec_.reserve(9 * n + 1);
  const int root = node_alloc();
  for (int i = 0; i < 3; i++)
    {
      POINT(root, i) = -(i + 1);
      LINK(root, i) = 0;
    }
  make_rec(root);

  T.shuffle.resize(n);
  for (size_t i = 0; i < n; i++) T.shuffle[i] = order ? order[i] : (int64_t) i;
  px_.resize(n); py_.resize(n);
  for (size_t i = 0; i < n; i++)
    {
      if (T.shuffle[i] < 0 || (size_t) T.shuffle[i] >= n) return fail("insertion order is not a permutation");
      px_[i] = xa[T.shuffle[i]];
      py_[i] = ya[T.shuffle[i]];
    }

  tree_view view{&rec_, &T.lnk, &T.type};
  T.batched_points = T.batched_rounds = 0;
  T.batched_seconds = 0;
  /* Rounds of the batched locate: the first kFirstBatch points are located on the host (the tree
     is small and hot in cache); afterwards every round covers as many points as are already in
     the tree (the snapshot doubles), so the descent that is left for the host is the few levels
     the tree grew by during the round. */
  const size_t kFirstBatch = 4096;
  size_t n_prev_nodes = 0;
  std::vector<int32_t> b_node, b_last;
  std::vector<double> b_c0, b_c1;
  if (locator && n > kFirstBatch)
    {
      const int st = locator->begin(T.xf, px_.data(), py_.data(), n, 9 * n + 16);
      if (st) locator = nullptr; /* no device: plain host build */
    }
  size_t i = 0;
  while (i < n)
    {
      if (!locator || i < kFirstBatch)
        {
          int last = -1;
          gsi_rec leaf_rec;
          const int leaf = descend_view(&view, &T.xf, 0, px_[i], py_[i], &c0_, &c1_, &last, &leaf_rec, 0);
          if (leaf < 0) return fail("given point outside the domain");
          const int st = insert_point(leaf);
          if (st) return st;
          i++;
          continue;
        }
      const size_t count = std::min(n - i, i);
      b_node.resize(count); b_last.resize(count); b_c0.resize(count); b_c1.resize(count);
      const auto t0 = std::chrono::steady_clock::now();
      const size_t n_nodes = T.type.size();
      const int ls = locator->locate(rec_.data(), T.lnk.data(), T.type.data(), n_nodes, n_prev_nodes, i, count,
                                     b_node.data(), b_c0.data(), b_c1.data(), b_last.data());
      T.batched_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (ls) return fail("batched locate failed");
      n_prev_nodes = n_nodes;
      T.batched_rounds++;
      T.batched_points += count;
      for (size_t k = 0; k < count; k++)
        {
          int leaf = b_node[k];
          if (leaf < 0 || (size_t) leaf >= n_nodes) return fail("given point outside the domain");
          c0_ = b_c0[k];
          c1_ = b_c1[k];
          if (!LEAF(leaf))
            {
              /* the snapshot's leaf has been subdivided since: finish the descent on the live tree */
              int last = b_last[k];
              gsi_rec leaf_rec;
              leaf = descend_view(&view, &T.xf, leaf, px_[i + k], py_[i + k], &c0_, &c1_, &last, &leaf_rec, 1);
              if (leaf < 0) return fail("given point outside the domain");
            }
          const int st = insert_point(leaf);
          if (st) return st;
        }
      i += count;
    }
  return 0;
}

/* the batched protocol on the host: descends on a private copy of the snapshot */
class host_snapshot_locator : public gsi_build_locator
{
public:
  int begin(const gsi_xform &xf, const double *px, const double *py, size_t, size_t) override
  {
    xf_ = xf; px_ = px; py_ = py;
    rec_.clear(); lnk_.clear(); type_.clear();
    return 0;
  }
  int locate(const gsi_rec *rec, const int32_t *lnk, const uint8_t *type, size_t n_nodes, size_t n_prev, size_t first,
             size_t count, int32_t *node, double *c0, double *c1, int32_t *last) override
  {
    rec_.insert(rec_.end(), rec + n_prev, rec + n_nodes); /* records are immutable: only the new ones */
    lnk_.assign(lnk, lnk + 3 * n_nodes);
    type_.assign(type, type + n_nodes);
    tree_view view{&rec_, &lnk_, &type_};
    for (size_t k = 0; k < count; k++)
      {
        double a = 0, b = 0;
        int l = -1;
        gsi_rec leaf_rec;
        node[k] = descend_view(&view, &xf_, 0, px_[first + k], py_[first + k], &a, &b, &l, &leaf_rec, 0);
        c0[k] = a; c1[k] = b; last[k] = l;
      }
    return 0;
  }

private:
  gsi_xform xf_;
  const double *px_ = nullptr, *py_ = nullptr;
  std::vector<gsi_rec> rec_;
  std::vector<int32_t> lnk_;
  std::vector<uint8_t> type_;
};

} // namespace

void
gsi_shuffle_mt19937(unsigned long seed, size_t n, int64_t *order)
{
  for (size_t i = 0; i < n; i++) order[i] = (int64_t) i;
  if (n < 2) return;
  mt19937 rng(seed);
  for (size_t i = n - 1; i > 0; i--) /* randist/shuffle.c:69-79 */
    {
      const size_t j = rng.uniform_int(i + 1);
      std::swap(order[i], order[j]);
    }
}

int
gsi_build_tree(const double *xa, const double *ya, size_t n, int flags, const int64_t *order,
               gsi_host_tree *out, std::string *err, gsi_build_locator *locator)
{
  if (n > 0 && (!xa || !ya)) return 4; /* GSL_EINVAL */
  if (n > (size_t) 200000000) return 4; /* node ids are int32: 9n must fit */
  const auto t0 = std::chrono::steady_clock::now();
  builder b(*out, err);
  const int st = b.run(xa, ya, n, flags, order, locator);
  out->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return st;
}

gsi_build_locator *
gsi_make_host_snapshot_locator()
{
  return new host_snapshot_locator();
}

void
gsi_make_dag_records(const gsi_host_tree &t, const double *xa, const double *ya,
                     std::vector<gsi_rec> *recs)
{
  const size_t N = t.n_nodes();
  recs->resize(N);
  for (size_t i = 0; i < N; i++)
    {
      double a[2], b[2], x0[2];
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i], &a[0], &a[1]);
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i + 1], &b[0], &b[1]);
      gsi_vertex_xy(t, xa, ya, t.pts[3 * i + 2], &x0[0], &x0[1]);
      gsi_rec &r = (*recs)[i];
      const uint32_t m = gsi_make_lu(&t.xf, a[0], a[1], b[0], b[1], x0[0], x0[1], &r);
      r.meta = m | t.type[i];
      r.lnk[0] = t.lnk[3 * i];
      r.lnk[1] = t.lnk[3 * i + 1];
      r.lnk[2] = t.lnk[3 * i + 2];
    }
}
