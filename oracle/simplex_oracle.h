/* oracle/simplex_oracle.h -- TEST INFRASTRUCTURE ONLY (see simplex_oracle.c). */
#ifndef SIMPLEX_ORACLE_H
#define SIMPLEX_ORACLE_H

#include <stddef.h>

typedef struct so_tree so_tree;

/* Arrays are copied.  pts3/lnk3/type: flattened simplex_tree pools (POINT, LINK, type per
 * node); shuffle: tree->shuffle; seed6: tree->seed_points (3x2); data: n x 2, row stride tda. */
so_tree *so_tree_create(int n_simplexes, const int *pts3, const int *lnk3, const int *type,
                        const unsigned long long *shuffle, const double shift[2],
                        const double scale[2], const double seed6[6], const double *data,
                        size_t tda, int n);
void so_tree_free(so_tree *t);

/* find_leaf + n_fields x interp_point for nq points; z field-major (n_fields x n), val
 * field-major (n_fields x nq).  Outside the cage: leaf -1, value NaN.  Returns the number
 * of outside-cage queries. */
long so_eval(const so_tree *t, const double *xq, const double *yq, long nq, const double *z,
             int n_fields, int *leaf, double *val);

/* find_leaf + the gradient of every field (definition: oracle/ref_driver.c:deriv_point);
 * dzdx / dzdy field-major (n_fields x nq), may be NULL; NaN outside the cage / singular leaf */
long so_eval_deriv(const so_tree *t, const double *xq, const double *yq, long nq, const double *z,
                   int n_fields, int *leaf, double *dzdx, double *dzdy);

/* so_eval on n_threads POSIX threads sharing the (read-only) tree: thread-safe, unlike the
 * reference (tree->tmp_vec1, linear_simplex.c:644).  SURVEY.md section 8f.4. */
long so_eval_mt(const so_tree *t, const double *xq, const double *yq, long nq, const double *z,
                int n_fields, int *leaf, double *val, int n_threads);

/* per-query descent statistics: depth[q] = DAG levels, tests[q] = contains_point calls */
long so_eval_stats(const so_tree *t, const double *xq, const double *yq, long nq, int *depth,
                   int *tests);

#endif
