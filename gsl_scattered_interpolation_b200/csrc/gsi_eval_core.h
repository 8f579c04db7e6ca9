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
#define GSI_STREAM_STORE(p, v) __stcs((p), (v))
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

typedef struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    gsi_int4
{
  int32_t x, y, z, w;
} gsi_int4;

typedef struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    gsi_double2
{
  double x, y;
} gsi_double2;

GSI_HD void
gsi_fetch_rec(gsi_rec *r, const gsi_rec *base, int idx)
{
#if defined(__CUDA_ARCH__)
  const double2 *p = reinterpret_cast<const double2 *>(base + idx);
  const double2 a = __ldg(p);
  const double2 b = __ldg(p + 1);
  const double2 c = __ldg(p + 2);
  const int4 d = __ldg(reinterpret_cast<const int4 *>(p + 3));
  r->x0x = a.x; r->x0y = a.y;
  r->u00 = b.x; r->u01 = b.y;
  r->u11 = c.x; r->l10 = c.y;
  r->lnk[0] = d.x; r->lnk[1] = d.y; r->lnk[2] = d.z;
  r->meta = (uint32_t) d.w;
#else
  *r = base[idx];
#endif
}

GSI_HD gsi_int4
gsi_fetch_aux(const int32_t *aux, int slot)
{
#if defined(__CUDA_ARCH__)
  const int4 v = __ldg(reinterpret_cast<const int4 *>(aux) + slot);
  gsi_int4 o;
  o.x = v.x; o.y = v.y; o.z = v.z; o.w = v.w;
  return o;
#else
  gsi_int4 o;
  o.x = aux[4 * slot]; o.y = aux[4 * slot + 1]; o.z = aux[4 * slot + 2]; o.w = aux[4 * slot + 3];
  return o;
#endif
}

#define GSI_FETCH_DAG(dst, idx) gsi_fetch_rec(&(dst), dag, (idx))
GSI_DEFINE_DESCEND(gsi_descend, GSI_HD, const gsi_rec *, GSI_FETCH_DAG)

/* interp_point for the F fields of one query, field-major output val[f*ld + q] */
GSI_HD void
gsi_blend_store(const gsi_view *V, const gsi_int4 ax, double c0, double c1, double *val, size_t ld,
                size_t q)
{
  const int F = V->n_fields;
  const int ha = ax.x >= 0, hb = ax.y >= 0, hx = ax.z >= 0;
  const double *za = V->z + (size_t) (ha ? ax.x : 0) * F;
  const double *zb = V->z + (size_t) (hb ? ax.y : 0) * F;
  const double *zx = V->z + (size_t) (hx ? ax.z : 0) * F;
  if (F == 1)
    {
      const double a = ha ? GSI_LDG(za) : 0.0, b = hb ? GSI_LDG(zb) : 0.0,
                   x = hx ? GSI_LDG(zx) : 0.0;
      GSI_STREAM_STORE(val + q, gsi_blend(c0, c1, ha, a, hb, b, hx, x));
      return;
    }
#if defined(__CUDA_ARCH__)
  if ((F & 1) == 0)
    {
      /* 128-bit gathers: the F values of a vertex are contiguous and 16-byte aligned */
      const double2 *pa = reinterpret_cast<const double2 *>(za);
      const double2 *pb = reinterpret_cast<const double2 *>(zb);
      const double2 *px = reinterpret_cast<const double2 *>(zx);
#pragma unroll 4
      for (int f = 0; f < F; f += 2)
        {
          double2 a = make_double2(0, 0), b = a, x = a;
          if (ha) a = __ldg(pa + (f >> 1));
          if (hb) b = __ldg(pb + (f >> 1));
          if (hx) x = __ldg(px + (f >> 1));
          __stcs(val + (size_t) f * ld + q, gsi_blend(c0, c1, ha, a.x, hb, b.x, hx, x.x));
          __stcs(val + (size_t) (f + 1) * ld + q, gsi_blend(c0, c1, ha, a.y, hb, b.y, hx, x.y));
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
  const int node = gsi_descend(V->dag, &V->xf, qx, qy, &c0, &c1, &last, &lr);
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
  if (val) gsi_blend_store(V, gsi_fetch_aux(V->aux, slot), c0, c1, val, ld, q);
  return slot;
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

/* Fast path for one query.  Returns 1 when the query was resolved (outputs written),
 * 0 when it must be deferred to the exact replay.  *steps = leaves visited.
 *
 * Acceptance: all three reference weights c0, c1, 1-(c0+c1) of the visited leaf are >= the
 * leaf's margin tau, i.e. the query is at least `delta` (standardised units) inside every
 * edge; then the reference's DAG descent provably ends in this leaf and the
 * weights are by construction the ones interp_point uses.  Otherwise cross the edge of the
 * most negative weight if it is clearly negative (<= -tau), else defer. */
GSI_HD int
gsi_locate_one(const gsi_view *V, double qx, double qy, int max_steps, double *val, int *leaf_out,
               size_t ld, size_t q, int *steps)
{
  *steps = 0;
  if (!(isfinite(qx) && isfinite(qy))) return 0;
  int slot = GSI_LDG(V->seeds + gsi_cell_of(V, qx, qy));
  for (int step = 0; step < max_steps; step++)
    {
      gsi_rec r;
      gsi_fetch_rec(&r, V->leaf, slot);
      const double tau = (double) GSI_BITS_TO_FLOAT(r.meta & 0x7fffffffu);
      double c0, c1;
      gsi_solve(&V->xf, &r, (int) (r.meta >> 31), qx, qy, &c0, &c1);
      const double c2 = GSI_SUB(1.0, GSI_ADD(GSI_ADD(0.0, c0), c1));
      *steps = step + 1;
      if ((c0 >= tau) && (c1 >= tau) && (c2 >= tau))
        {
          const gsi_int4 ax = gsi_fetch_aux(V->aux, slot);
          if (leaf_out) GSI_STREAM_STORE(leaf_out + q, ax.w);
          if (val) gsi_blend_store(V, ax, c0, c1, val, ld, q);
          return 1;
        }
      int e = 0;
      double m = c0;
      if (c1 < m) { m = c1; e = 1; }
      if (c2 < m) { m = c2; e = 2; }
      if (!(m <= -tau)) return 0; /* within the margin of an edge (or NaN): defer */
      const int nb = GSI_LNK(r, e);
      if (nb < 0) return 0; /* beyond the cage hull: the exact path decides */
      slot = nb;
    }
  return 0;
}

#endif /* GSI_EVAL_CORE_H */
