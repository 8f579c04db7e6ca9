// host-side sanitizer run: builder + flatten + kernel logic (hostsim pieces) under ASAN/UBSAN
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "gsi_builder.h"
#include "gsi_eval_core.h"
#include "gsi_flatten.h"
int main()
{
  std::mt19937_64 rng(5);
  std::uniform_real_distribution<double> U(0, 1);
  for (int trial = 0; trial < 12; trial++)
    {
      const size_t n = trial == 0 ? 0 : (trial == 1 ? 1 : (trial == 2 ? 3 : 50 + 400 * trial));
      std::vector<double> xa(n), ya(n), za(2 * n);
      for (size_t i = 0; i < n; i++) { xa[i] = trial % 3 == 0 ? (double) (i % 17) : U(rng); ya[i] = trial % 3 == 0 ? (double) (i / 17) : U(rng) * (trial % 2 ? 1e-3 : 1.0); za[i] = U(rng); za[n + i] = U(rng); }
      std::vector<int64_t> ord(n);
      gsi_shuffle_mt19937(123 + trial, n, ord.data());
      gsi_host_tree tree; std::string err;
      const int flags = trial % 4 == 3 ? 1 : 0;
      if (gsi_build_tree(xa.data(), ya.data(), n, flags, trial % 2 ? ord.data() : nullptr, &tree, &err)) { printf("trial %d build failed: %s\n", trial, err.c_str()); continue; }
      gsi_flatten_params prm; prm.cells_per_leaf = trial % 2 ? 4.0 : 0.3; prm.tile_leaves = trial % 2 ? 1000 : 5;
      std::vector<char> image;
      if (gsi_flatten(tree, xa.data(), ya.data(), za.data(), 2, prm, &image, &err)) { printf("trial %d flatten failed\n", trial); continue; }
      gsi_slab_header h; memcpy(&h, image.data(), sizeof h);
      gsi_view V = gsi_view_of(image.data(), &h);
      const size_t ncell = (size_t) h.grid_g * h.grid_g;
      int32_t *entry = (int32_t *) (image.data() + h.off_entry), *seeds = (int32_t *) (image.data() + h.off_seeds);
      for (size_t c = 0; c < ncell; c++) entry[c] = gsi_entry_of_cell(&V, c);
      for (size_t c = 0; c < ncell; c++) { double qx, qy; gsi_cell_centre(&V, c, &qx, &qy); int s = gsi_exact_one(&V, qx, qy, nullptr, nullptr, 0, c); seeds[c] = s < 0 ? 0 : s; }
      const size_t nq = 20000;
      std::vector<double> z(2 * nq); std::vector<int> leaf(nq);
      long deferred = 0, mism = 0;
      for (size_t q = 0; q < nq; q++)
        {
          double qx = -0.3 + 1.6 * U(rng) * (trial % 3 == 0 ? 17 : 1), qy = -0.3 + 1.6 * U(rng) * (trial % 3 == 0 ? 17 : 1);
          if (q % 7 == 0 && n) { qx = xa[q % n]; qy = ya[q % n]; }
          if (q == 5) qx = NAN; if (q == 6) qy = INFINITY; if (q == 7) qx = 1e300;
          int steps; double z2[2]; int l2;
          gsi_exact_one(&V, qx, qy, z2, &l2, 1, 0);
          if (!gsi_locate_one(&V, qx, qy, 64, z.data(), leaf.data(), nq, q, &steps)) { deferred++; continue; }
          if (leaf[q] != l2 || memcmp(&z[q], &z2[0], 8) || memcmp(&z[nq + q], &z2[1], 8)) mism++;
        }
      printf("trial %d n %zu nodes %zu leaves %llu G %u tiles 2^%u deferred %ld mismatches %ld\n", trial, n, tree.n_nodes(), (unsigned long long) h.n_leaves, h.grid_g, h.tile_bits, deferred, mism);
      if (mism) return 1;
    }
  printf("sanitizer run ok\n");
  return 0;
}
