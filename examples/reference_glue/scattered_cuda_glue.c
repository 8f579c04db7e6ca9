/* examples/reference_glue/scattered_cuda_glue.c -- the translation unit a maintainer of the
 * reference adds next to interpolation/scattered_interp_example.c.
 *
 * It includes the REFERENCE's headers (linear_simplex.h and GSL's) and the C ABI header of
 * libgsl_scattered_cuda, flattens the simplex_tree pools (linear_simplex.h:31-65) and uploads
 * them with gsl_scattered_cuda_init_from_tree.  Nothing else of the reference changes.
 * Built only where the reference sources exist (oracle/Makefile target `glue`). */
#include <stdlib.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_linalg.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_permutation.h>
#include <gsl/gsl_rng.h>
#include "linear_simplex.h"
#include <gsl_scattered_cuda.h>

/* response: n_fields x n, field-major (one gsl_vector per field, concatenated) */
gsl_scattered_cuda *
simplex_tree_to_cuda(simplex_tree *tree, const gsl_matrix *data, const double *response,
                     size_t n_fields)
{
  const size_t N = tree->n_simplexes, n = tree->n_points;
  int *pts = malloc(3 * N * sizeof(int)), *lnk = malloc(3 * N * sizeof(int));
  int *type = malloc(N * sizeof(int));
  size_t k;
  int i;
  gsl_scattered_cuda *w;
  for (k = 0; k < N; k++)
    {
      for (i = 0; i < 3; i++)
        {
          pts[3 * k + i] = POINT(k, i); /* tree->pidx[i + SIMP(k)->points] */
          lnk[3 * k + i] = LINK(k, i);  /* tree->links[i + SIMP(k)->links] */
        }
      type[k] = SIMP(k)->type;          /* 2-bit bitfield -> int */
    }
  gsl_scattered_cuda_set_error_handler(gsl_error); /* honour the user's GSL policy */
  w = gsl_scattered_cuda_alloc(n, n_fields);
  if (w
      && gsl_scattered_cuda_init_from_tree(w, N, pts, lnk, type, tree->shuffle->data, tree->shift->data,
                                           tree->scale->data, tree->seed_points->data, data->data,
                                           data->tda, n, response, n_fields)
             != GSL_SUCCESS)
    {
      gsl_scattered_cuda_free(w);
      w = NULL;
    }
  free(pts);
  free(lnk);
  free(type);
  return w;
}
