/* gsi_eval_core.h -- per-query logic of the two evaluation kernels, written once for the
 * device (gsi_kernels.cu) and for the host-side logic simulator used by the CPU tests
 * (tests/hostsim/hostsim.cpp; never part of the shipped library's evaluation path).
 *
 *   gsi_exact_one   exact replay of find_leaf + interp_point  (linear_simplex.c:331-402,678-711)
 *   gsi_locate_one  grid-seeded barycentric walk with the safety-margin acceptance test
 */
#ifndef GSI_EVAL_CORE_H
#define GSI_EVAL_CORE_H

#include <math.h>
#include <string.h>

#include "gsi_core.h"
#include "gsi_layout.h"

#if defined(__CUDA_ARCH__)
#define GSI_LDG(p) __ldg(p)
#define GSI_STREAM_STORE(p, v) (*(p) = (v))
#define GSI_NAN __longlong_as_double(0x7ff8000000000000LL)
#define GSI_FLOOR_INT(x) __double2int_rd(x)
#define GSI_BITS_TO_FLOAT(u) __uint_as_float(u)
#else
#define GSI_LDG(p) (*(p))
#define GSI_STREAM_STORE(p, v) (*(p) = (v))
#define GSI_NAN (NAN)
#define GSI_FLOOR_INT(x) ((int) floor(x))
static inline float
gsi_bits_to_float_host(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}
#define GSI_BITS_TO_FLOAT(u) gsi_bits_to_float_host(u)
#endif

#if defined(__CUDACC__)
/* L2 fill-size hint of the structure gathers (leaf records, seed cells).  Measured on B200
 * (scripts/gather_probe3.cu, profiles/): without a hint every L2 miss fills a whole 128-byte
 * line from HBM whatever cudaLimitMaxL2FetchGranularity says; ".L2::64B" halves the fill. */
#ifndef GSI_GATHER_HINT
#define GSI_GATHER_HINT ""
#endif
/* one 32-byte sector with a single 256-bit load (LDG.E.256, sm_100+); read-only path */
__device__ __forceinline__ void
gsi_ld256(const void *p, unsigned long long &a, unsigned long long &b, unsigned long long &c,
          unsigned long long &d)
{
  asm volatile("ld.global.nc" GSI_GATHER_HINT ".v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(p));
}

__device__ __forceinline__ int
gsi_ld_seed(const int32_t *p)
{
  int v;
  asm volatile("ld.global.nc" GSI_GATHER_HINT ".s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
/* the 16-bit seed offset of a cell (walk kernels; L1::no_allocate on this load and on the query loads was
   measured 20-35 % slower, profiles/r2u_carveout_and_pitch.txt section 4) */
__device__ __forceinline__ unsigned
gsi_ld_seed16(const uint16_t *p)
{
  return __ldg(p);
}
#if defined(GSI_STREAM_HINTS)
#define GSI_LD_QUERY(p) __ldcs(p)
#define GSI_ST_RESULT(p, v) __stcs((p), (v))
#else
#define GSI_LD_QUERY(p) __ldg(p)
#define GSI_ST_RESULT(p, v) (*(p) = (v))
#endif
#endif

/* the 64-byte LU/link part of a record (DAG node or leaf): two sectors */
GSI_HD void
gsi_fetch_rec(gsi_rec *r, const void *rec)
{
#if defined(__CUDA_ARCH__)
  unsigned long long a, b, c, d, e, f, g, h;
  gsi_ld256(rec, a, b, c, d);
  gsi_ld256(reinterpret_cast<const char *>(rec) + 32, e, f, g, h);
  r->x0x = __longlong_as_double((long long) a);
  r->x0y = __longlong_as_double((long long) b);
  r->u00 = __longlong_as_double((long long) c);
  r->u01 = __longlong_as_double((long long) d);
  r->u11 = __longlong_as_double((long long) e);
  r->l10 = __longlong_as_double((long long) f);
  r->lnk[0] = (int32_t) (unsigned) g;
  r->lnk[1] = (int32_t) (unsigned) (g >> 32);
  r->lnk[2] = (int32_t) (unsigned) h;
  r->meta = (uint32_t) (h >> 32);
#else
  *r = *(const gsi_rec *) rec;
#endif
}

/* the 32-byte tail of a leaf: one sector */
GSI_HD void
gsi_fetch_tail(gsi_leaf_tail *t, const gsi_leaf *leaf)
{
#if defined(__CUDA_ARCH__)
  unsigned long long a, b, c, d;
  gsi_ld256(&leaf->t, a, b, c, d);
  t->s.node = (int32_t) (unsigned) a;
  t->s.vflags = (uint32_t) (a >> 32);
  t->s.z[0] = __longlong_as_double((long long) b);
  t->s.z[1] = __longlong_as_double((long long) c);
  t->s.z[2] = __longlong_as_double((long long) d);
#else
  *t = leaf->t;
#endif
}

#define GSI_FETCH_DAG(dst, idx) gsi_fetch_rec(&(dst), dag + (idx))
GSI_DEFINE_DESCEND(gsi_descend, GSI_HD, const gsi_rec *, GSI_FETCH_DAG)

/* interp_point for the F fields of one query, field-major output val[f*ld + q].
 * `t` is the accepted leaf's tail: with one field it carries the three vertex values, else the
 * vertex slots of the vertex-major value table (gathered with 256-/128-bit loads). */
GSI_HD void
gsi_blend_store(const gsi_view *V, const gsi_leaf_tail *t, double c0, double c1, double *val,
                size_t ld, size_t q)
{
  const int F = V->n_fields;
  if (F == 1)
    {
      const uint32_t vf = t->s.vflags;
      GSI_STREAM_STORE(val + q, gsi_blend(c0, c1, (int) (vf & 1u), t->s.z[0], (int) (vf & 2u),
                                          t->s.z[1], (int) (vf & 4u), t->s.z[2]));
      return;
    }
  const int ha = t->g.v[0] >= 0, hb = t->g.v[1] >= 0, hx = t->g.v[2] >= 0;
  const double *za = V->z + (size_t) (ha ? t->g.v[0] : 0) * F;
  const double *zb = V->z + (size_t) (hb ? t->g.v[1] : 0) * F;
  const double *zx = V->z + (size_t) (hx ? t->g.v[2] : 0) * F;
#if defined(__CUDA_ARCH__)
  if ((F & 3) == 0)
    {
      /* the F values of a vertex are contiguous and 32-byte aligned: 256-bit gathers */
      double *out = val + q;
      if (ha & hb & hx)
        {
          /* common case (no cage vertex): straight-line code, same operations as gsi_blend:
             ((0 + c0*za) + c1*zb) + (1 - (c0 + c1))*zx.  Not unrolled: measured on the 8192^2 x
             16-field raster, occupancy (78 vs 92 registers) beats a second group of loads in
             flight (3.99 ms vs 4.61 ms). */
          const double w2 = GSI_SUB(1.0, GSI_ADD(GSI_ADD(0.0, c0), c1));
#pragma unroll 1
          for (int f = 0; f < F; f += 4)
            {
              unsigned long long a[4], b[4], x[4];
              gsi_ld256(za + f, a[0], a[1], a[2], a[3]);
              gsi_ld256(zb + f, b[0], b[1], b[2], b[3]);
              gsi_ld256(zx + f, x[0], x[1], x[2], x[3]);
#pragma unroll
              for (int k = 0; k < 4; k++)
                {
                  double v = GSI_ADD(0.0, GSI_MUL(c0, __longlong_as_double((long long) a[k])));
                  v = GSI_ADD(v, GSI_MUL(c1, __longlong_as_double((long long) b[k])));
                  v = GSI_ADD(v, GSI_MUL(w2, __longlong_as_double((long long) x[k])));
                  *out = v;
                  out += ld;
                }
            }
          return;
        }
#pragma unroll 1
      for (int f = 0; f < F; f += 4)
        {
          unsigned long long a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, x[4] = {0, 0, 0, 0};
          if (ha) gsi_ld256(za + f, a[0], a[1], a[2], a[3]);
          if (hb) gsi_ld256(zb + f, b[0], b[1], b[2], b[3]);
          if (hx) gsi_ld256(zx + f, x[0], x[1], x[2], x[3]);
#pragma unroll
          for (int k = 0; k < 4; k++)
            {
              *out = gsi_blend(c0, c1, ha, __longlong_as_double((long long) a[k]), hb,
                               __longlong_as_double((long long) b[k]), hx,
                               __longlong_as_double((long long) x[k]));
              out += ld;
            }
        }
      return;
    }
  if ((F & 1) == 0)
    {
      const double2 *pa = reinterpret_cast<const double2 *>(za);
      const double2 *pb = reinterpret_cast<const double2 *>(zb);
      const double2 *px = reinterpret_cast<const double2 *>(zx);
#pragma unroll 1
      for (int f = 0; f < F; f += 2)
        {
          double2 a = make_double2(0, 0), b = a, x = a;
          if (ha) a = __ldg(pa + (f >> 1));
          if (hb) b = __ldg(pb + (f >> 1));
          if (hx) x = __ldg(px + (f >> 1));
          val[(size_t) f * ld + q] = gsi_blend(c0, c1, ha, a.x, hb, b.x, hx, x.x);
          val[(size_t) (f + 1) * ld + q] = gsi_blend(c0, c1, ha, a.y, hb, b.y, hx, x.y);
        }
      return;
    }
#endif
  for (int f = 0; f < F; f++)
    {
      const double a = ha ? GSI_LDG(za + f) : 0.0, b = hb ? GSI_LDG(zb + f) : 0.0,
                   x = hx ? GSI_LDG(zx + f) : 0.0;
      GSI_STREAM_STORE(val + (size_t) f * ld + q, gsi_blend(c0, c1, ha, a, hb, b, hx, x));
    }
}

/* grid cell of a query in standardised coordinates (a hint only: no exactness needed;
 * fmin/fmax drop NaN so non-finite queries land in cell 0) */
GSI_HD size_t
gsi_cell_of(const gsi_view *V, double qx, double qy)
{
  const double u = (qx - V->xf.cx) * V->xf.sx, v = (qy - V->xf.cy) * V->xf.sy;
  const double gmax = (double) (V->grid_g - 1);
  const double fx = fmin(fmax((u - V->u0) * V->inv_du, 0.0), gmax);
  const double fy = fmin(fmax((v - V->v0) * V->inv_dv, 0.0), gmax);
  return (size_t) GSI_FLOOR_INT(fy) * (size_t) V->grid_g + (size_t) GSI_FLOOR_INT(fx);
}

/* Seed of the fast path for a query: the compact table (two independent loads that stay in L2),
 * the int32 table only for escaped cells.  Like the cell itself this is a hint: any leaf slot is
 * a correct starting point of the walk. */
GSI_HD int
gsi_seed_of(const gsi_view *V, double qx, double qy)
{
  const double u = (qx - V->xf.cx) * V->xf.sx, v = (qy - V->xf.cy) * V->xf.sy;
  const double gmax = (double) (V->grid_g - 1);
  const int ix = GSI_FLOOR_INT(fmin(fmax((u - V->u0) * V->inv_du, 0.0), gmax));
  const int iy = GSI_FLOOR_INT(fmin(fmax((v - V->v0) * V->inv_dv, 0.0), gmax));
  const size_t cell = (size_t) iy * (size_t) V->grid_g + (size_t) ix;
#if defined(__CUDA_ARCH__)
  const unsigned off = __ldg(V->seed16 + cell);
  const int base = __ldg(V->seedbase + (size_t) (iy >> V->seed_shift) * (size_t) V->grid_gb + (size_t) (ix >> V->seed_shift));
  if (off != GSI_SEED16_ESC) return base + (int) off;
  return gsi_ld_seed(V->seeds + cell);
#else
  const unsigned off = V->seed16[cell];
  if (off != GSI_SEED16_ESC) return V->seedbase[(size_t) (iy >> V->seed_shift) * (size_t) V->grid_gb + (size_t) (ix >> V->seed_shift)] + (int) off;
  return V->seeds[cell];
#endif
}

/* centre of grid cell `cell` in raw coordinates (seeding) */
GSI_HD void
gsi_cell_centre(const gsi_view *V, size_t cell, double *qx, double *qy)
{
  const size_t g = (size_t) V->grid_g;
  const size_t iy = cell / g, ix = cell - iy * g;
  const double u = V->u0 + ((double) ix + 0.5) / V->inv_du;
  const double v = V->v0 + ((double) iy + 0.5) / V->inv_dv;
  *qx = V->xf.cx + u / V->xf.sx;
  *qy = V->xf.cy + v / V->xf.sy;
}

/* ---- entry nodes of the exact replay ------------------------------------------------------------
 * The reference descends ~50 history-DAG levels per query at 10^6 nodes, and every level costs
 * up to three divergent 64-byte record fetches.  Most of those levels decide nothing for a
 * small neighbourhood: if a whole grid cell lies inside child j of a node with a safety margin,
 * and outside every earlier sibling with the same margin, then the reference's test sequence
 * (children in link order, first closed-interval pass wins, linear_simplex.c:363-373) takes
 * child j for EVERY point of the cell.  gsi_entry_of_cell follows such certified steps from the
 * root and returns the last node reached; a query of that cell may start its replay there with
 * accel->coords = that node's coordinates, exactly the state the reference arrives with.
 *
 * Certificates (barycentric weights are affine, so the four corners of the slightly enlarged
 * cell bound the whole cell): PASS = all three weights of all four corners >= 2*tau, FAIL = one
 * weight index whose value is <= -2*tau at all four corners, tau = delta * Lmax / |det| being
 * the weight change of a `delta` move towards an edge line (same margin as the fast path's
 * acceptance; delta = 2^-28 standardised units, four orders above the rounding error of the
 * solve).  Children that are singular or have link 0 fail in the reference without a test
 * (linear_simplex.c:366, 641-642).  Anything else stops the walk.  Border cells also receive
 * the clamped queries from outside the grid (and the non-finite ones), so they keep the root. */
GSI_HD double
gsi_rec_tau(const gsi_rec *r, double delta)
{
  /* columns of the edge matrix (a - x0, b - x0) up to the row swap, which does not change
     lengths: M = L*U */
  const double m00 = r->u00, m01 = r->u01;
  const double m10 = r->l10 * r->u00, m11 = r->u11 + r->l10 * r->u01;
  const double e0 = sqrt(m00 * m00 + m10 * m10), e1 = sqrt(m01 * m01 + m11 * m11);
  const double dx = m00 - m01, dy = m10 - m11;
  const double e2 = sqrt(dx * dx + dy * dy);
  double lmax = e0 > e1 ? e0 : e1;
  if (e2 > lmax) lmax = e2;
  const double det = fabs(r->u00 * r->u11);
  return delta * lmax / det; /* NaN / inf for degenerate records: never certified */
}

/* 1 = the cell certainly passes record r, -1 = certainly fails, 0 = undecided */
GSI_HD int
gsi_cell_vs_rec(const gsi_xform *xf, const gsi_rec *r, double delta, const double cx[4],
                const double cy[4])
{
  const double tau2 = 2.0 * gsi_rec_tau(r, delta);
  if (!(tau2 < 0.5)) return 0;
  int all_in = 1, out0 = 1, out1 = 1, out2 = 1;
  for (int k = 0; k < 4; k++)
    {
      double c0, c1;
      gsi_solve(xf, r, (int) (r->meta & GSI_META_SWAP), cx[k], cy[k], &c0, &c1);
      const double c2 = 1.0 - (c0 + c1);
      if (!(c0 >= tau2 && c1 >= tau2 && c2 >= tau2)) all_in = 0;
      if (!(c0 <= -tau2)) out0 = 0;
      if (!(c1 <= -tau2)) out1 = 0;
      if (!(c2 <= -tau2)) out2 = 0;
    }
  if (all_in) return 1;
  return (out0 | out1 | out2) ? -1 : 0;
}

/* deepest DAG node every point of the box with corners (cx[k], cy[k]) is certain to pass; *depth
 * (may be NULL) = levels descended */
GSI_HD int
gsi_entry_of_box(const gsi_view *V, const double cx[4], const double cy[4], int *depth)
{
  const gsi_rec *dag = V->dag;
  gsi_rec cur;
  if (depth) *depth = 0;
  GSI_FETCH_DAG(cur, 0);
  if (cur.meta & GSI_META_SINGULAR) return 0;
  if (gsi_cell_vs_rec(&V->xf, &cur, V->delta, cx, cy) != 1) return 0;
  int node = 0;
  for (int guard = 0; guard < GSI_MAX_DESCENT; guard++)
    {
      const uint32_t type = cur.meta & GSI_META_TYPE_MASK;
      if (type == GSI_LEAF) break;
      const int nch = (type == GSI_SUB_DPLUS1) ? 3 : 2;
      int chosen = 0;
      gsi_rec rc;
      for (int i = 0; i < nch; i++)
        {
          const int k = GSI_LNK(cur, i);
          if (!k) continue;                          /* link 0: fails without a test */
          GSI_FETCH_DAG(rc, k);
          if (rc.meta & GSI_META_SINGULAR) continue; /* singular: fails without a test */
          const int verdict = gsi_cell_vs_rec(&V->xf, &rc, V->delta, cx, cy);
          if (verdict == 1) chosen = k;
          if (verdict != -1) break;                  /* taken, or undecided: stop looking */
        }
      if (!chosen) break;
      node = chosen;
      cur = rc;
      if (depth) ++*depth;
    }
  return node;
}

GSI_HD int
gsi_entry_of_cell(const gsi_view *V, size_t cell)
{
  const size_t g = (size_t) V->grid_g;
  const size_t iy = cell / g, ix = cell - iy * g;
  if (ix == 0 || iy == 0 || ix + 1 >= g || iy + 1 >= g) return 0;
  const double pad = 1.0 / 1024; /* of a cell: covers the rounding of gsi_cell_of by far */
  double cx[4], cy[4];
  for (int k = 0; k < 4; k++)
    {
      const double u = V->u0 + ((double) ix + ((k & 1) ? 1.0 + pad : -pad)) / V->inv_du;
      const double v = V->v0 + ((double) iy + ((k & 2) ? 1.0 + pad : -pad)) / V->inv_dv;
      cx[k] = V->xf.cx + u / V->xf.sx;
      cy[k] = V->xf.cy + v / V->xf.sy;
    }
  return gsi_entry_of_box(V, cx, cy, (int *) 0);
}

/* The bounding box, in raw coordinates, of the leaf whose record is r (vertices x0, x0 + column 0,
 * x0 + column 1 of the edge matrix M = L U, rows un-swapped), inflated by `grow` of its size plus
 * twice the absolute margin delta: every point the fast path can defer AT this leaf -- all three
 * weights above -tau, i.e. at most delta outside an edge -- lies inside. */
GSI_HD void
gsi_leaf_box(const gsi_view *V, const gsi_rec *r, int swap, double grow, double cx[4], double cy[4])
{
  const double m00 = r->u00, m01 = r->u01;
  const double m10 = r->l10 * r->u00, m11 = r->u11 + r->l10 * r->u01;
  /* standardised edge vectors a - x0, b - x0 (row 0 = x unless the rows were swapped) */
  const double ax = swap ? m10 : m00, ay = swap ? m00 : m10;
  const double bx = swap ? m11 : m01, by = swap ? m01 : m11;
  const double u0 = (r->x0x - V->xf.cx) * V->xf.sx, v0 = (r->x0y - V->xf.cy) * V->xf.sy;
  double ulo = fmin(u0, fmin(u0 + ax, u0 + bx)), uhi = fmax(u0, fmax(u0 + ax, u0 + bx));
  double vlo = fmin(v0, fmin(v0 + ay, v0 + by)), vhi = fmax(v0, fmax(v0 + ay, v0 + by));
  const double pu = grow * (uhi - ulo) + 4.0 * V->delta, pv = grow * (vhi - vlo) + 4.0 * V->delta;
  ulo -= pu; uhi += pu; vlo -= pv; vhi += pv;
  for (int k = 0; k < 4; k++)
    {
      cx[k] = V->xf.cx + ((k & 1) ? uhi : ulo) / V->xf.sx;
      cy[k] = V->xf.cy + ((k & 2) ? vhi : vlo) / V->xf.sy;
    }
}

/* ---- vertex memo (gsi_layout.h) ------------------------------------------------------------------ */
GSI_HD uint64_t
gsi_f64_bits(double v)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t) __double_as_longlong(v);
#else
  uint64_t u;
  memcpy(&u, &v, 8);
  return u;
#endif
}

GSI_HD uint32_t
gsi_memo_hash(uint64_t xb, uint64_t yb)
{
  uint64_t z = xb * 0x9E3779B97F4A7C15ull + (yb ^ (yb >> 29)) * 0xBF58476D1CE4E5B9ull;
  z ^= z >> 31; z *= 0x94D049BB133111EBull;
  z ^= z >> 29;
  return (uint32_t) z;
}

/* the memoised answer for a query that coincides bit for bit with a data node: 1 = found */
GSI_HD int
gsi_memo_find(const gsi_view *V, double qx, double qy, int *node, int *slot)
{
  if (!V->memo) return 0;
  const uint64_t xb = gsi_f64_bits(qx), yb = gsi_f64_bits(qy);
  uint32_t h = gsi_memo_hash(xb, yb) & V->memo_mask;
  for (int probe = 0; probe < GSI_MEMO_MAX_PROBES; probe++, h = (h + 1u) & V->memo_mask)
    {
#if defined(__CUDA_ARCH__)
      unsigned long long a, b, c, d;
      gsi_ld256(V->memo + h, a, b, c, d);
      const uint32_t state = (uint32_t) a;
      if (state == GSI_MEMO_EMPTY) return 0;
      if (b == xb && c == yb)
        {
          if (state != GSI_MEMO_FULL) return 0;
          *node = (int32_t) (uint32_t) (a >> 32);
          *slot = (int32_t) (uint32_t) d;
          return 1;
        }
#else
      const gsi_memo *e = V->memo + h;
      if (e->state == GSI_MEMO_EMPTY) return 0;
      if (gsi_f64_bits(e->x) == xb && gsi_f64_bits(e->y) == yb)
        {
          if (e->state != GSI_MEMO_FULL) return 0;
          *node = e->node;
          *slot = e->slot;
          return 1;
        }
#endif
    }
  return 0;
}

#define GSI_EXACT_OUTSIDE (-1)

/* Exact replay of one query.  Returns the leaf's compact slot (>= 0) or GSI_EXACT_OUTSIDE;
 * writes leaf_out[q] (reference node id, -1 outside) and val (NaN outside) when non-NULL. */
GSI_HD int
gsi_exact_one(const gsi_view *V, double qx, double qy, double *val, int *leaf_out, size_t ld,
              size_t q)
{
  double c0 = 0, c1 = 0;
  int last = -1;
  gsi_rec lr;
  const int start = V->entry ? GSI_LDG(V->entry + gsi_cell_of(V, qx, qy)) : 0;
  const int node = gsi_descend(V->dag, &V->xf, start, qx, qy, &c0, &c1, &last, &lr, 0);
  if (node < 0)
    {
      if (leaf_out) leaf_out[q] = -1;
      if (val)
        for (int f = 0; f < V->n_fields; f++) val[(size_t) f * ld + q] = GSI_NAN;
      return GSI_EXACT_OUTSIDE;
    }
  const int slot = lr.lnk[0];
  /* interp_point recomputes the leaf's coordinates unless the leaf is singular, in which
     case accel->coords keeps whatever the descent left there (linear_simplex.c:641-642,687) */
  if (!(lr.meta & GSI_META_SINGULAR) && last != node)
    gsi_solve(&V->xf, &lr, (int) (lr.meta & GSI_META_SWAP), qx, qy, &c0, &c1);
  if (leaf_out) leaf_out[q] = node;
  if (val)
    {
      gsi_leaf_tail t;
      gsi_fetch_tail(&t, gsi_leaf_at(V, slot));
      gsi_blend_store(V, &t, c0, c1, val, ld, q);
    }
  return slot;
}

/* Gradient of the interpolant on the leaf with reference node id `node` (as returned by the
 * locate kernels; < 0 = outside the cage): writes d/dx and d/dy of every field, field-major
 * (out[f*ld + q]); NaN outside the cage and on singular leaves (where gsl_linalg_LU_solve raises
 * GSL_EDOM, linalg/lu.c:145-148).  Either output may be NULL. */
GSI_HD void
gsi_deriv_one(const gsi_view *V, int node, double *dzdx, double *dzdy, size_t ld, size_t q)
{
  const int F = V->n_fields;
  gsi_rec r;
  int ok = node >= 0 && node < V->n_dag;
  if (ok)
    {
      gsi_fetch_rec(&r, V->dag + node);
      ok = (r.meta & GSI_META_TYPE_MASK) == GSI_LEAF && !(r.meta & GSI_META_SINGULAR);
    }
  if (!ok)
    {
      for (int f = 0; f < F; f++)
        {
          if (dzdx) dzdx[(size_t) f * ld + q] = GSI_NAN;
          if (dzdy) dzdy[(size_t) f * ld + q] = GSI_NAN;
        }
      return;
    }
  const int swap = (int) (r.meta & GSI_META_SWAP);
  double ax, bx, ay, by; /* d(c0, c1)/dx and d(c0, c1)/dy */
  gsi_solve_rhs(&r, swap, V->xf.sx, 0.0, &ax, &bx);
  gsi_solve_rhs(&r, swap, 0.0, V->xf.sy, &ay, &by);
  gsi_leaf_tail t;
  gsi_fetch_tail(&t, gsi_leaf_at(V, r.lnk[0]));
  if (F == 1)
    {
      const int ha = (int) (t.s.vflags & 1u), hb = (int) (t.s.vflags & 2u), hx = (int) (t.s.vflags & 4u);
      if (dzdx) dzdx[q] = gsi_blend_deriv(ax, bx, ha, t.s.z[0], hb, t.s.z[1], hx, t.s.z[2]);
      if (dzdy) dzdy[q] = gsi_blend_deriv(ay, by, ha, t.s.z[0], hb, t.s.z[1], hx, t.s.z[2]);
      return;
    }
  const int ha = t.g.v[0] >= 0, hb = t.g.v[1] >= 0, hx = t.g.v[2] >= 0;
  const double *za = V->z + (size_t) (ha ? t.g.v[0] : 0) * F;
  const double *zb = V->z + (size_t) (hb ? t.g.v[1] : 0) * F;
  const double *zx = V->z + (size_t) (hx ? t.g.v[2] : 0) * F;
  for (int f = 0; f < F; f++)
    {
      const double a = ha ? GSI_LDG(za + f) : 0.0, b = hb ? GSI_LDG(zb + f) : 0.0, x = hx ? GSI_LDG(zx + f) : 0.0;
      if (dzdx) dzdx[(size_t) f * ld + q] = gsi_blend_deriv(ax, bx, ha, a, hb, b, hx, x);
      if (dzdy) dzdy[(size_t) f * ld + q] = gsi_blend_deriv(ay, by, ha, a, hb, b, hx, x);
    }
}

#define GSI_STEP_ACCEPT 0 /* strictly inside with margin: the leaf is the answer */
#define GSI_STEP_MOVE 1   /* clearly outside across one edge: *slot = neighbour */
#define GSI_STEP_DEFER 2  /* cannot be proven here: exact replay decides */

/* One step of the fast path on the leaf record `r` of `*slot`.
 *
 * Acceptance: all three reference weights c0, c1, 1-(c0+c1) of the visited leaf are >= the
 * leaf's margin tau, i.e. the query is at least `delta` (standardised units) inside every
 * edge; then the reference's DAG descent provably ends in this leaf and
 * the weights are by construction the ones interp_point uses.  Otherwise cross the edge of the
 * most negative weight if it is clearly negative (<= -tau), else defer. */
GSI_HD int
gsi_walk_step(const gsi_xform *xf, const gsi_rec *r, double qx, double qy, int *slot, double *c0,
              double *c1)
{
  const double tau = (double) GSI_BITS_TO_FLOAT(r->meta & 0x7fffffffu);
  gsi_solve(xf, r, (int) (r->meta >> 31), qx, qy, c0, c1);
  const double c2 = GSI_SUB(1.0, GSI_ADD(GSI_ADD(0.0, *c0), *c1));
  if ((*c0 >= tau) && (*c1 >= tau) && (c2 >= tau)) return GSI_STEP_ACCEPT;
  int e = 0;
  double m = *c0;
  if (*c1 < m) { m = *c1; e = 1; }
  if (c2 < m) { m = c2; e = 2; }
  if (!(m <= -tau)) return GSI_STEP_DEFER; /* within the margin of an edge (or NaN) */
  const int nb = GSI_LNK(*r, e);
  if (nb < 0) return GSI_STEP_DEFER;       /* beyond the cage hull */
  *slot = nb;
  return GSI_STEP_MOVE;
}

#if defined(__CUDACC__)
/* gsi_walk_step for a record whose divisors' reciprocals (r11 for u11, r00 for u00, gsi_div_rcp) are
 * at hand: the two divisions of the solve cost three operations each.  Same operations otherwise;
 * when the hoisted division cannot vouch for a quotient the step defers (exact replay decides). */
__device__ __forceinline__ int
gsi_walk_step_rcp(const gsi_xform *xf, const gsi_rec *r, double r11, double r00, double qx, double qy, int *slot,
                  double *c0, double *c1)
{
  const double tau = (double) GSI_BITS_TO_FLOAT(r->meta & 0x7fffffffu);
  double b0 = GSI_MUL(GSI_SUB(qx, r->x0x), xf->sx);
  double b1 = GSI_MUL(GSI_SUB(qy, r->x0y), xf->sy);
  if (r->meta >> 31)
    {
      const double s = b0; b0 = b1; b1 = s;
    }
  unsigned bad = 0;
  b1 = GSI_SUB(b1, GSI_MUL(r->l10, b0));
  b1 = gsi_div_by_rcp(b1, r->u11, r11, &bad);
  b0 = gsi_div_by_rcp(GSI_SUB(b0, GSI_MUL(r->u01, b1)), r->u00, r00, &bad);
  *c0 = b0;
  *c1 = b1;
  if (bad) return GSI_STEP_DEFER;
  const double c2 = GSI_SUB(1.0, GSI_ADD(GSI_ADD(0.0, b0), b1));
  if ((b0 >= tau) && (b1 >= tau) && (c2 >= tau)) return GSI_STEP_ACCEPT;
  int e = 0;
  double m = b0;
  if (b1 < m) { m = b1; e = 1; }
  if (c2 < m) { m = c2; e = 2; }
  if (!(m <= -tau)) return GSI_STEP_DEFER;
  const int nb = GSI_LNK(*r, e);
  if (nb < 0) return GSI_STEP_DEFER;
  *slot = nb;
  return GSI_STEP_MOVE;
}
#endif

/* Fast path for one query, start to finish (host logic simulator; the device kernel runs
 * the same gsi_walk_step with lanes decoupled from queries).  Returns 1 when the query was
 * resolved (outputs written), 0 when it must be deferred.  *steps = leaves visited. */
GSI_HD int
gsi_locate_one(const gsi_view *V, double qx, double qy, int max_steps, double *val, int *leaf_out,
               size_t ld, size_t q, int *steps)
{
  *steps = 0;
  if (!(isfinite(qx) && isfinite(qy))) return 0;
  int slot = gsi_seed_of(V, qx, qy);
  for (int step = 0; step < max_steps; step++)
    {
      gsi_rec r;
      double c0, c1;
      const gsi_leaf *lf = gsi_leaf_at(V, slot);
      gsi_fetch_rec(&r, lf);
      *steps = step + 1;
      const int st = gsi_walk_step(&V->xf, &r, qx, qy, &slot, &c0, &c1);
      if (st == GSI_STEP_ACCEPT)
        {
          gsi_leaf_tail t;
          gsi_fetch_tail(&t, lf);
          if (leaf_out) GSI_STREAM_STORE(leaf_out + q, t.g.node);
          if (val) gsi_blend_store(V, &t, c0, c1, val, ld, q);
          return 1;
        }
      if (st == GSI_STEP_DEFER) return 0;
    }
  return 0;
}

#endif /* GSI_EVAL_CORE_H */
