/* oracle/ref_stub_integrity.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Link-time stand-in for the reference's debug translation unit
 * /root/reference/interpolation/linear_simplex_integrity_check.c.  The reference calls
 * check_delaunay() after every insertion (linear_simplex.c:489, an O(leaves x points) brute
 * force) and output_triangulation() -- a gnuplot dump to /tmp -- on every edge flip
 * (edge_flip.c:302-303,314-315), which makes its builder unusable above a few hundred
 * nodes.  Neither hook changes the tree, so replacing the TU (no reference source is
 * modified) yields bit-identical trees; oracle/_ref/libref_simplex_checked.so keeps the
 * real TU for small-n structure tests and tests/test_oracle_ref.py checks both agree.
 */
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_permutation.h>
#include <gsl/gsl_rng.h>
#include "linear_simplex.h"

int check_delaunay(simplex_tree *tree, gsl_matrix *data)
{
  (void) tree; (void) data;
  return 1;
}

void output_triangulation(simplex_tree *tree, gsl_matrix *data, gsl_vector *response,
                          int standardize_output, char lines_filename[],
                          char points_filename[], char circles_filename[])
{
  (void) tree; (void) data; (void) response; (void) standardize_output;
  (void) lines_filename; (void) points_filename; (void) circles_filename;
}
