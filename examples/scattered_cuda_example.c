/* examples/scattered_cuda_example.c -- a plain C (C89) host program on libgsl_scattered_cuda.
 *
 * The evaluation part of scattered_data() in the reference's driver
 * (interpolation/scattered_interp_example.c:130-199) through the C ABI: build the triangulation of
 * scattered nodes, interpolate one point, then resample onto an n_grid x n_grid raster with the
 * reference's own pixel coordinates (x = min[0] + xstep*i, y = min[1] + ystep*j, x-major result).
 *
 *   scattered_cuda_example nodes.txt seed n_grid xmin xmax ymin ymax [px py]
 *
 * nodes.txt holds one "x y z" triple per line; seed is the gsl_rng_mt19937 seed of the
 * insertion-order shuffle (negative: data order, i.e. rng == NULL).  Prints the value at (px,py)
 * and the raster, one "%.17g" value per line (exactly round-trippable), so that tests can compare
 * the output bit for bit with the reference's.
 *
 *   gcc -std=c89 -pedantic -Wall -I include examples/scattered_cuda_example.c \
 *       -L gsl_scattered_interpolation_b200/lib -lgsl_scattered_cuda -o scattered_cuda_example
 */
#include <stdio.h>
#include <stdlib.h>

#include <gsl_scattered_cuda.h>

static void
report (const char *reason, const char *file, int line, int gsl_errno)
{
  fprintf (stderr, "gsl_scattered_cuda: %s:%d: ERROR %d: %s\n", file, line, gsl_errno, reason);
}

int
main (int argc, char **argv)
{
  FILE *f;
  double *xa, *ya, *za, *grid;
  size_t *order = NULL;
  size_t n = 0, cap = 64, i, n_grid;
  long seed;
  double min[2], max[2], xstep, ystep, px = -88, py = 41, x, y, z;
  gsl_scattered_cuda *w;
  int status;

  if (argc < 8)
    {
      fprintf (stderr, "usage: %s nodes.txt seed n_grid xmin xmax ymin ymax [px py]\n", argv[0]);
      return 2;
    }
  f = fopen (argv[1], "r");
  if (!f)
    {
      perror (argv[1]);
      return 2;
    }
  xa = (double *) malloc (cap * sizeof (double));
  ya = (double *) malloc (cap * sizeof (double));
  za = (double *) malloc (cap * sizeof (double));
  while (fscanf (f, "%lf %lf %lf", &x, &y, &z) == 3)
    {
      if (n == cap)
        {
          cap *= 2;
          xa = (double *) realloc (xa, cap * sizeof (double));
          ya = (double *) realloc (ya, cap * sizeof (double));
          za = (double *) realloc (za, cap * sizeof (double));
        }
      xa[n] = x;
      ya[n] = y;
      za[n] = z;
      n++;
    }
  fclose (f);
  seed = atol (argv[2]);
  n_grid = (size_t) atol (argv[3]);
  min[0] = atof (argv[4]);
  max[0] = atof (argv[5]);
  min[1] = atof (argv[6]);
  max[1] = atof (argv[7]);
  if (argc >= 10)
    {
      px = atof (argv[8]);
      py = atof (argv[9]);
    }

  gsl_scattered_cuda_set_error_handler (report);
  if (gsl_scattered_cuda_device_count () <= 0)
    {
      fprintf (stderr, "no CUDA device: libgsl_scattered_cuda has no CPU evaluation path\n");
      return 3;
    }
  /* simplex_tree_alloc + simplex_tree_init(tree, data, NULL, NULL, 0, rng) */
  w = gsl_scattered_cuda_alloc (n, 1);
  if (!w)
    return 1;
  if (seed >= 0)
    {
      order = (size_t *) malloc ((n ? n : 1) * sizeof (size_t));
      gsl_scattered_cuda_shuffle_mt19937 ((unsigned long) seed, n, order);
    }
  status = gsl_scattered_cuda_init (w, xa, ya, za, n, GSL_SCATTERED_DEFAULT, order);
  if (status != GSL_SCATTERED_SUCCESS)
    return 1;

  /* leaf = find_leaf(...); res = interp_point(...) for one point */
  printf ("%.17g\n", gsl_scattered_cuda_eval (w, px, py));

  /* the gridding loop */
  grid = (double *) malloc ((n_grid ? n_grid * n_grid : 1) * sizeof (double));
  xstep = (max[0] - min[0]) / (double) n_grid;
  ystep = (max[1] - min[1]) / (double) n_grid;
  status = gsl_scattered_cuda_eval_raster_e (w, min[0], xstep, n_grid, min[1], ystep, n_grid, grid, NULL);
  if (status != GSL_SCATTERED_SUCCESS && status != GSL_SCATTERED_EDOM)
    return 1;
  for (i = 0; i < n_grid * n_grid; i++)
    printf ("%.17g\n", grid[i]);

  gsl_scattered_cuda_free (w);
  free (grid);
  free (order);
  free (xa);
  free (ya);
  free (za);
  return 0;
}
