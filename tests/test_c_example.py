"""The drop-in boundary is a C ABI: a plain C89 host program (examples/scattered_cuda_example.c, the
evaluation part of the reference's scattered_data() driver, scattered_interp_example.c:130-199)
must compile against include/gsl_scattered_cuda.h with a strict C89 compiler, link to the shared
library, fail loudly without a GPU, and on a GPU print exactly the reference's numbers."""
import os
import subprocess

import numpy as np
import pytest

from gsl_scattered_interpolation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "scattered_cuda_example.c")


def _build(tmp_path, std="c89"):
    exe = str(tmp_path / f"scattered_cuda_example_{std}")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", f"-std={std}", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", libdir, "-lgsl_scattered_cuda", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("std", ["c89", "c99", "c11"])
def test_c_host_program_compiles_and_links(tmp_path, std, have_gpu):
    exe = _build(tmp_path, std)
    if not have_gpu:
        nodes = tmp_path / "nodes.txt"
        nodes.write_text("0 0 1\n1 0 2\n0 1 3\n")
        r = subprocess.run([exe, str(nodes), "-1", "4", "0", "1", "0", "1"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU evaluation path" in r.stderr      # no silent fallback


@pytest.mark.gpu
def test_c_host_program_reproduces_the_reference_raster(tmp_path, golden):
    exe = _build(tmp_path)
    wd = golden["weather"]
    nodes = tmp_path / "weather.txt"
    nodes.write_text("".join(f"{x!r} {y!r} {z!r}\n" for x, y, z in wd.tolist()))
    x0, dx, nx, y0, dy, ny = golden["weatherB_raster_spec"]
    nx = int(nx)
    # the driver's own box: min + step * n_grid (scattered_interp_example.c:176-183)
    args = [exe, str(nodes), "0", str(nx), repr(float(x0)), repr(float(x0 + dx * nx)), repr(float(y0)),
            repr(float(y0 + dy * nx)), "-88", "41"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.array([float(t) for t in r.stdout.split()])
    assert out.size == 1 + nx * nx
    assert out[0] == 274.08608080148747
    want = golden["weatherB_raster_val"].reshape(-1)
    step_ok = (float(x0 + dx * nx) - float(x0)) / nx == dx and (float(y0 + dy * nx) - float(y0)) / nx == dy
    if step_ok:                                                               # same xstep / ystep bits
        assert np.array_equal(out[1:].view(np.uint64), want.view(np.uint64))
    else:
        assert np.allclose(out[1:], want, rtol=1e-9, equal_nan=True)


GLUE = os.path.join(ROOT, "oracle", "_ref", "glue_demo")


def test_reference_side_glue_builds_against_the_real_reference(have_gpu):
    """INTEGRATION.md section 2 as a real translation unit (examples/reference_glue/): compiled by
    oracle/Makefile against the reference's own headers and objects.  Without a GPU the demo
    builds the tree with the reference and then refuses to evaluate (no CPU path)."""
    if not os.path.exists(GLUE):
        pytest.skip("oracle/_ref/glue_demo not built (needs the reference sources at build time)")
    if have_gpu:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([GLUE, "200", "3", "16"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU evaluation path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_grid", [(50, 1, 100), (2000, 12345, 256), (20000, 7, 512)])
def test_true_drop_in_reference_tree_evaluated_on_the_gpu(n, seed, n_grid):
    """One C process: simplex_tree_init of the UNMODIFIED reference builds the tree, the glue hands
    it to gsl_scattered_cuda_init_from_tree, and a raster (sticking out of the hull) plus the data
    nodes themselves are evaluated both by the reference's find_leaf + interp_point loop and by
    the GPU library: leaf ids and value bits must be identical."""
    if not os.path.exists(GLUE):
        pytest.skip("oracle/_ref/glue_demo not built")
    r = subprocess.run([GLUE, str(n), str(seed), str(n_grid)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "leaf_mismatches 0 value_bit_mismatches 0" in r.stdout
