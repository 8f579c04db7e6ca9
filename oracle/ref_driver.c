/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline).
 *
 * A thin C driver around the UNMODIFIED reference implementation, which oracle/Makefile
 * compiles from the sources where they lie under /root/reference (nothing is copied).
 * It exposes the reference's own public API
 *     simplex_tree_alloc / simplex_tree_init / find_leaf / interp_point / simplex_tree_free
 *     (/root/reference/interpolation/linear_simplex.h:105-179)
 * through plain-pointer entry points that Python (ctypes) can call, so that
 *   - tests/ can compare the CUDA path and the C restatement (oracle/simplex_oracle.c)
 *     with what the reference itself returns on the same inputs,
 *   - tests/golden/make_golden.py can freeze golden vectors,
 *   - bench.py's cpu_baseline / --impl reference leg can time the reference's own eval loop
 *     (find_leaf + interp_point, linear_simplex.c:331-402,678-711) on the host cores.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 *
 * Driver hygiene: seed_points(0,1) is never written by
 * simplex_tree_init (gsl_matrix_alloc, linear_simplex.c:58,217-232) so it is zeroed here
 * before init; min and max are never both passed; one simplex_tree_accel per evaluating
 * process; asserts stay enabled; "outside the cage" (find_leaf would assert(0),
 * linear_simplex.c:343-347) is detected first with the reference's own contains_point on
 * the root node and reported as leaf = -1 / value = NaN.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

#include <gsl/gsl_math.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_linalg.h>
#include <gsl/gsl_rng.h>
#include "linear_simplex.h"

typedef struct ref_handle
{
  simplex_tree *tree;
  gsl_matrix *data;      /* n x 2, owned copy */
  int n;
  double build_seconds;
} ref_handle;

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Build with the reference builder.  xy: n x 2 row-major.  use_rng != 0 -> gsl_rng_mt19937
 * seeded with `seed` is passed to simplex_tree_init (random insertion order). */
ref_handle *ref_build(const double *xy, int n, int flags, int use_rng, unsigned long seed)
{
  ref_handle *h = calloc(1, sizeof(*h));
  int i;
  h->n = n;
  h->data = gsl_matrix_alloc(n > 0 ? n : 1, 2);
  for (i = 0; i < n; i++)
    {
      gsl_matrix_set(h->data, i, 0, xy[2*i]);
      gsl_matrix_set(h->data, i, 1, xy[2*i+1]);
    }
  h->tree = simplex_tree_alloc(2, n > 0 ? n : 1);
  gsl_matrix_set_zero(h->tree->seed_points);   /* step around the uninitialised (0,1) entry */
  gsl_rng *rng = NULL;
  if (use_rng)
    {
      rng = gsl_rng_alloc(gsl_rng_mt19937);
      gsl_rng_set(rng, seed);
    }
  double t0 = now_s();
  int ret;
  if (n > 0)
    ret = simplex_tree_init(h->tree, h->data, NULL, NULL, flags, rng);
  else
    ret = simplex_tree_init(h->tree, NULL, NULL, NULL, flags | SIMPLEX_TREE_NOSTANDARDIZE, NULL);
  h->build_seconds = now_s() - t0;
  if (rng) gsl_rng_free(rng);
  if (ret != GSL_SUCCESS)
    {
      fprintf(stderr, "ref_build: simplex_tree_init failed (%d)\n", ret);
      return NULL;
    }
  return h;
}

void ref_free(ref_handle *h)
{
  if (!h) return;
  simplex_tree_free(h->tree);
  gsl_matrix_free(h->data);
  free(h);
}

double ref_build_seconds(const ref_handle *h) { return h->build_seconds; }
int ref_n_nodes(const ref_handle *h) { return h->tree->n_simplexes; }
int ref_n_points(const ref_handle *h) { return h->tree->n_points; }

/* Flattened export of the history DAG: per node 3 vertex ids (negative = cage seed),
 * 3 links (children for internal nodes, neighbours-opposite-vertex for leaves), type. */
void ref_export(const ref_handle *h, int *pts3, int *lnk3, int *type,
                long long *shuffle, double *shift2, double *scale2, double *seed6)
{
  simplex_tree *tree = h->tree;
  int i, k;
  for (i = 0; i < tree->n_simplexes; i++)
    {
      for (k = 0; k < 3; k++)
        {
          pts3[3*i+k] = POINT(i, k);
          lnk3[3*i+k] = LINK(i, k);
        }
      type[i] = (int) SIMP(i)->type;
    }
  for (i = 0; i < h->n; i++)
    shuffle[i] = (long long) gsl_permutation_get(tree->shuffle, i);
  for (k = 0; k < 2; k++)
    {
      shift2[k] = gsl_vector_get(tree->shift, k);
      scale2[k] = gsl_vector_get(tree->scale, k);
    }
  for (i = 0; i < 3; i++)
    for (k = 0; k < 2; k++)
      seed6[2*i+k] = gsl_matrix_get(tree->seed_points, i, k);
}

/* The reference eval loop on queries [q0,q1): find_leaf + n_fields x interp_point.
 * z is field-major (n_fields x n).  val is field-major (n_fields x nq). */
static long eval_range(const ref_handle *h, const double *xq, const double *yq,
                       long q0, long q1, long nq, const double *z, int n_fields,
                       int *leaf_out, double *val_out)
{
  simplex_tree *tree = h->tree;
  simplex_tree_accel *accel = simplex_tree_accel_alloc(2);
  double pt[2];
  gsl_vector_view point = gsl_vector_view_array(pt, 2);
  long outside = 0, q;
  int f;
  for (q = q0; q < q1; q++)
    {
      pt[0] = xq[q];
      pt[1] = yq[q];
      if (!contains_point(tree, 0, h->data, &point.vector, accel))
        {
          outside++;
          if (leaf_out) leaf_out[q] = -1;
          if (val_out) for (f = 0; f < n_fields; f++) val_out[(long)f*nq + q] = NAN;
          continue;
        }
      simplex_index leaf = find_leaf(tree, h->data, &point.vector, accel);
      if (leaf_out) leaf_out[q] = leaf;
      if (val_out)
        for (f = 0; f < n_fields; f++)
          {
            gsl_vector_const_view resp = gsl_vector_const_view_array(z + (long)f*h->n, h->n > 0 ? h->n : 1);
            val_out[(long)f*nq + q] = interp_point(tree, leaf, h->data, (gsl_vector *)&resp.vector,
                                                   &point.vector, accel);
          }
    }
  simplex_tree_accel_free(accel);
  return outside;
}

long ref_eval(const ref_handle *h, const double *xq, const double *yq, long nq,
              const double *z, int n_fields, int *leaf_out, double *val_out)
{
  return eval_range(h, xq, yq, 0, nq, nq, z, n_fields, leaf_out, val_out);
}

/* Same loop on nproc forked workers (the tree is not thread-safe -- tree->tmp_vec1 is shared
 * scratch, linear_simplex.c:644 -- so processes, not threads; the children share the tree
 * copy-on-write and each uses its own accel).  Returns wall seconds of the parallel region;
 * outputs may be NULL (timing only). */
double ref_eval_fork(const ref_handle *h, const double *xq, const double *yq, long nq,
                     const double *z, int n_fields, int *leaf_out, double *val_out, int nproc)
{
  if (nproc < 1) nproc = 1;
  size_t lbytes = sizeof(int) * (size_t)nq, vbytes = sizeof(double) * (size_t)nq * n_fields;
  int *sl = mmap(NULL, lbytes ? lbytes : 1, PROT_READ|PROT_WRITE, MAP_SHARED|MAP_ANONYMOUS, -1, 0);
  double *sv = mmap(NULL, vbytes ? vbytes : 1, PROT_READ|PROT_WRITE, MAP_SHARED|MAP_ANONYMOUS, -1, 0);
  if (sl == MAP_FAILED || sv == MAP_FAILED) return -1.0;
  double t0 = now_s();
  int p;
  for (p = 0; p < nproc; p++)
    {
      long q0 = nq * p / nproc, q1 = nq * (p + 1) / nproc;
      pid_t pid = fork();
      if (pid == 0)
        {
          eval_range(h, xq, yq, q0, q1, nq, z, n_fields, sl, sv);
          _exit(0);
        }
      else if (pid < 0)
        return -1.0;
    }
  int status, bad = 0;
  for (p = 0; p < nproc; p++)
    {
      wait(&status);
      if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = 1;
    }
  double dt = now_s() - t0;
  if (leaf_out) memcpy(leaf_out, sl, lbytes);
  if (val_out) memcpy(val_out, sv, vbytes);
  munmap(sl, lbytes ? lbytes : 1);
  munmap(sv, vbytes ? vbytes : 1);
  return bad ? -1.0 : dt;
}

/* Oracle-side definition of the derivative evaluators (the reference has
 * none -- README:14 "converging to the standard GSL interpolation interface", whose 2-D types carry
 * eval_deriv_x / eval_deriv_y, interpolation/gsl_interp2d.h:37-49).  The gradient of the leaf's
 * barycentric plane, built only from the reference's own pieces: calculate_bary_coords
 * (linear_simplex.c:607-651) leaves the leaf's LU factors and permutation in the accel; the
 * derivative of the coordinates along axis k is the same gsl_linalg_LU_solve applied to
 * d/dx_k [(point - x0) * scale] = scale_k e_k; the blend is interp_point's loop
 * (linear_simplex.c:689-710) with d/dx_k (1 - tot) = 0 - tot.  Singular leaf or outside the
 * cage: NaN. */
static double deriv_point(simplex_tree *tree, simplex_index leaf, gsl_matrix *data, const double *z,
                          gsl_vector *point, simplex_tree_accel *accel, int axis)
{
  double rhs_[2] = {0.0, 0.0}, dc_[2] = {0.0, 0.0};
  gsl_vector_view rhs = gsl_vector_view_array(rhs_, 2), dc = gsl_vector_view_array(dc_, 2);
  double tot = 0, d = 0;
  int i, idx;
  if (GSL_SUCCESS != calculate_bary_coords(tree, leaf, data, point, accel)) return NAN;
  rhs_[axis] = gsl_vector_get(tree->scale, axis);
  gsl_linalg_LU_solve(accel->simplex_matrix, accel->perm, &rhs.vector, &dc.vector);
  for (i = 0; i < 2; i++)
    {
      double c = dc_[i];
      tot += c;
      idx = POINT(leaf, i);
      if (idx >= 0) d += c * z[gsl_permutation_get(tree->shuffle, idx)];
    }
  idx = POINT(leaf, 2);
  if (idx >= 0) d += (0.0 - tot) * z[gsl_permutation_get(tree->shuffle, idx)];
  return d;
}

/* find_leaf + the gradient of every field; dzdx / dzdy field-major (n_fields x nq), may be NULL */
long ref_eval_deriv(const ref_handle *h, const double *xq, const double *yq, long nq, const double *z,
                    int n_fields, int *leaf_out, double *dzdx, double *dzdy)
{
  simplex_tree *tree = h->tree;
  simplex_tree_accel *accel = simplex_tree_accel_alloc(2);
  double pt[2];
  gsl_vector_view point = gsl_vector_view_array(pt, 2);
  long outside = 0, q;
  int f;
  for (q = 0; q < nq; q++)
    {
      pt[0] = xq[q];
      pt[1] = yq[q];
      if (!contains_point(tree, 0, h->data, &point.vector, accel))
        {
          outside++;
          if (leaf_out) leaf_out[q] = -1;
          for (f = 0; f < n_fields; f++)
            {
              if (dzdx) dzdx[(long)f*nq + q] = NAN;
              if (dzdy) dzdy[(long)f*nq + q] = NAN;
            }
          continue;
        }
      simplex_index leaf = find_leaf(tree, h->data, &point.vector, accel);
      if (leaf_out) leaf_out[q] = leaf;
      for (f = 0; f < n_fields; f++)
        {
          if (dzdx) dzdx[(long)f*nq + q] = deriv_point(tree, leaf, h->data, z + (long)f*h->n, &point.vector, accel, 0);
          if (dzdy) dzdy[(long)f*nq + q] = deriv_point(tree, leaf, h->data, z + (long)f*h->n, &point.vector, accel, 1);
        }
    }
  simplex_tree_accel_free(accel);
  return outside;
}

/* The reference's raster driver shape (scattered_interp_example.c:176-199):
 * x = x0 + dx*i (outer), y = y0 + dy*j (inner), result stored x-major [i][j].
 * Two roundings per coordinate (this file is compiled with -ffp-contract=off). */
long ref_eval_raster(const ref_handle *h, double x0, double dx, long nx,
                     double y0, double dy, long ny, const double *z, int n_fields,
                     int *leaf_out, double *val_out)
{
  long nq = nx * ny, i, j;
  double *xq = malloc(sizeof(double) * (size_t)nq), *yq = malloc(sizeof(double) * (size_t)nq);
  for (i = 0; i < nx; i++)
    {
      double x = x0 + dx * i;
      for (j = 0; j < ny; j++)
        {
          double y = y0 + dy * j;
          xq[i*ny + j] = x;
          yq[i*ny + j] = y;
        }
    }
  long out = eval_range(h, xq, yq, 0, nq, nq, z, n_fields, leaf_out, val_out);
  free(xq); free(yq);
  return out;
}

/* The 50-station fixture of the reference example, read from where it lies
 * (interpolation/weather_data.c is a bare initialiser meant to be #included,
 * scattered_interp_example.c:135). out: 50 x 3 (lon, lat, temperature). */
void ref_weather_data(double *out150)
{
#include "weather_data.c"
  memcpy(out150, data_vec, sizeof(double) * 150);
}

#ifdef REF_CHECKED
#include "linear_simplex_integrity_check.h"
/* Full structural + brute-force Delaunay check of the reference (n <~ 300 only). */
int ref_check(const ref_handle *h)
{
  check_leaf_nodes(h->tree, NULL);
  return check_delaunay(h->tree, h->data);
}
#endif
