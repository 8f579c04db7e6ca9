/* examples/reference_glue/glue_demo.c -- TEST PROGRAM for the glue: the reference builds the tree
 * (its own simplex_tree_init), the glue hands it to libgsl_scattered_cuda, and the same process
 * evaluates a raster and a batch of points BOTH ways -- the reference's loop
 * (find_leaf + interp_point, scattered_interp_example.c:184-199) and one C-ABI call each -- and
 * counts differing leaf ids and differing value bit patterns.  Exit status 0 = identical.
 *
 *   glue_demo n_nodes seed n_grid
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_linalg.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_permutation.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_vector.h>
#include "linear_simplex.h"
#include <gsl_scattered_cuda.h>

gsl_scattered_cuda *simplex_tree_to_cuda(simplex_tree *tree, const gsl_matrix *data,
                                         const double *response, size_t n_fields);

int
main(int argc, char **argv)
{
  const size_t n = argc > 1 ? (size_t) atol(argv[1]) : 500;
  const unsigned long seed = argc > 2 ? (unsigned long) atol(argv[2]) : 12345;
  const size_t n_grid = argc > 3 ? (size_t) atol(argv[3]) : 128;
  const size_t npix = n_grid * n_grid;
  gsl_rng *gen = gsl_rng_alloc(gsl_rng_taus), *rng = gsl_rng_alloc(gsl_rng_mt19937);
  gsl_matrix *data = gsl_matrix_alloc(n, 2);
  gsl_vector *response = gsl_vector_alloc(n);
  double *z_gpu = malloc(npix * sizeof(double)), *z_ref = malloc(npix * sizeof(double));
  int *leaf_gpu = malloc(npix * sizeof(int)), *leaf_ref = malloc(npix * sizeof(int));
  size_t i, j, bad_leaf = 0, bad_val = 0;
  double pt[2];
  gsl_vector_view point = gsl_vector_view_array(pt, 2);
  simplex_tree *tree;
  simplex_tree_accel *accel;
  gsl_scattered_cuda *w;
  double min[2] = {-0.05, -0.05}, step;
  int status;

  gsl_rng_set(gen, 7);
  for (i = 0; i < n; i++)
    {
      gsl_matrix_set(data, i, 0, gsl_rng_uniform(gen));
      gsl_matrix_set(data, i, 1, gsl_rng_uniform(gen));
      gsl_vector_set(response, i, gsl_rng_uniform(gen));
    }
  /* the reference builds the triangulation */
  gsl_rng_set(rng, seed);
  tree = simplex_tree_alloc(2, (int) n);
  gsl_matrix_set_zero(tree->seed_points); /* entry (0,1) is never written by simplex_tree_init */
  if (simplex_tree_init(tree, data, NULL, NULL, 0, rng) != GSL_SUCCESS)
    return 2;
  accel = simplex_tree_accel_alloc(2);

  /* the GPU library adopts it */
  if (gsl_scattered_cuda_device_count() <= 0)
    {
      fprintf(stderr, "no CUDA device: libgsl_scattered_cuda has no CPU evaluation path\n");
      return 3;
    }
  w = simplex_tree_to_cuda(tree, data, response->data, 1);
  if (!w)
    return 1;

  /* the raster, both ways; the box sticks out of the data hull (cage leaves) */
  step = 1.1 / (double) n_grid;
  status = gsl_scattered_cuda_eval_raster_e(w, min[0], step, n_grid, min[1], step, n_grid, z_gpu, leaf_gpu);
  if (status != GSL_SUCCESS)
    return 1;
  for (i = 0; i < n_grid; i++)
    {
      pt[0] = min[0] + step * i;
      for (j = 0; j < n_grid; j++)
        {
          pt[1] = min[1] + step * j;
          leaf_ref[i * n_grid + j] = find_leaf(tree, data, &point.vector, accel);
          z_ref[i * n_grid + j] = interp_point(tree, leaf_ref[i * n_grid + j], data, response, &point.vector, accel);
        }
    }
  for (i = 0; i < npix; i++)
    {
      bad_leaf += leaf_gpu[i] != leaf_ref[i];
      bad_val += memcmp(z_gpu + i, z_ref + i, sizeof(double)) != 0;
    }
  /* the data nodes themselves (on-vertex queries: the reference's tie-breaks), as a batch */
  {
    double *xq = malloc(n * sizeof(double)), *yq = malloc(n * sizeof(double));
    for (i = 0; i < n; i++)
      {
        xq[i] = gsl_matrix_get(data, i, 0);
        yq[i] = gsl_matrix_get(data, i, 1);
      }
    status = gsl_scattered_cuda_eval_e(w, xq, yq, n, z_gpu, leaf_gpu);
    for (i = 0; i < n && status == GSL_SUCCESS; i++)
      {
        int lf;
        double v;
        pt[0] = xq[i];
        pt[1] = yq[i];
        lf = find_leaf(tree, data, &point.vector, accel);
        v = interp_point(tree, lf, data, response, &point.vector, accel);
        bad_leaf += lf != leaf_gpu[i];
        bad_val += memcmp(&v, z_gpu + i, sizeof(double)) != 0;
      }
    free(xq);
    free(yq);
  }
  printf("nodes %lu simplexes %d pixels %lu leaf_mismatches %lu value_bit_mismatches %lu\n", (unsigned long) n,
         tree->n_simplexes, (unsigned long) npix, (unsigned long) bad_leaf, (unsigned long) bad_val);
  gsl_scattered_cuda_free(w);
  simplex_tree_accel_free(accel);
  simplex_tree_free(tree);
  return (status == GSL_SUCCESS && bad_leaf == 0 && bad_val == 0) ? 0 : 1;
}
