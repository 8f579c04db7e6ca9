"""derivative evaluators and the gsl_interp2d_type-style descriptor.

The reference has no derivative entry point.  The definition is oracle/ref_driver.c:ref_eval_deriv
(the reference's own calculate_bary_coords + gsl_linalg_LU_solve + interp_point's blend loop);
tests/golden/golden_v2_deriv.npz freezes what it returns (tests/golden/make_golden_deriv.py).
CPU part: oracle/_ref and the port reproduce the golden vectors, the kernels' per-query logic
(csrc/gsi_eval_core.h:gsi_deriv_one, run by the host simulator) does too, and the gradient
reproduces a linear field.  GPU part (-m gpu): the CUDA path through the C ABI."""
import ctypes
import os

import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
import hostsim_py
import refimpl
from gsl_scattered_interpolation_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold2():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v2_deriv.npz"))


def _cfg1():
    xa, ya = synth.uniform_nodes(1000)
    return xa, ya, synth.values(1000, 1), g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000)


def _cfg3(golden):
    n = 3000
    xa, ya = synth.clustered_nodes(n)
    xq, yq = synth.mixed_queries(20000, xa, ya, golden["cfg3_leaf_rows"])
    return xa, ya, synth.values(n, 3), g.shuffle_mt19937(synth.SEED_SHUFFLE, n), xq, yq


def test_oracle_and_port_reproduce_the_golden_derivatives(golden, gold2):
    xa, ya, za, _ = _cfg1()
    xq, yq = synth.uniform_queries(10000)
    tree = {k: golden["cfg1_" + k] for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}
    pl, pdx, pdy = refimpl.PortTree(tree, xa, ya).eval_deriv(xq, yq, za)
    assert np.array_equal(pl, golden["cfg1_leaf"])
    assert refimpl.same_bits(pdx, gold2["cfg1_dzdx"]) and refimpl.same_bits(pdy, gold2["cfg1_dzdy"])
    if refimpl.have_ref():
        rt = refimpl.RefTree(xa, ya, 0, synth.SEED_SHUFFLE)
        rl, rdx, rdy = rt.eval_deriv(xq, yq, za)
        assert np.array_equal(rl, golden["cfg1_leaf"])
        assert refimpl.same_bits(rdx, gold2["cfg1_dzdx"]) and refimpl.same_bits(rdy, gold2["cfg1_dzdy"])
        rt.close()
        w = golden["weather"]
        tb = refimpl.RefTree(w[:, 0].copy(), w[:, 1].copy(), 0, seed=0)
        q = golden["weatherB_q"]
        _, dx, dy = tb.eval_deriv(q[:, 0], q[:, 1], w[:, 2].copy())
        assert refimpl.same_bits(dx, gold2["weatherB_dzdx"]) and refimpl.same_bits(dy, gold2["weatherB_dzdy"])
        assert np.isnan(dx[0, 8])        # outside the cage; the NaN query passes every test (linear_simplex.c:665-675)
        assert np.isfinite(dx[0, 9])     # and lands in a leaf whose gradient does not depend on the point
        tb.close()


def test_kernel_logic_of_the_gradient_on_the_golden_sets(golden, gold2):
    """gsi_deriv_one (what k_deriv runs per query) on the host, on config 1 and the config-3 mix
    (3 fields; on-vertex / on-edge / outside-hull queries, cage vertices dropped from the blend)"""
    xa, ya, za, order = _cfg1()
    hs = hostsim_py.HostSim(xa, ya, za, 0, order)
    dx, dy = hs.deriv(golden["cfg1_leaf"])
    assert refimpl.same_bits(dx, gold2["cfg1_dzdx"]) and refimpl.same_bits(dy, gold2["cfg1_dzdy"])
    dx, dy = hs.deriv(np.array([-1, 10 ** 8, 0], dtype=np.int32))   # outside, out of range, root (not a leaf)
    assert np.isnan(dx).all() and np.isnan(dy).all()
    hs.close()
    xa, ya, za, order, xq, yq = _cfg3(golden)
    hs = hostsim_py.HostSim(xa, ya, za, 0, order)
    dx, dy = hs.deriv(golden["cfg3_leaf"])
    assert refimpl.same_bits(dx, gold2["cfg3_dzdx"]) and refimpl.same_bits(dy, gold2["cfg3_dzdy"])
    hs.close()


def test_gradient_of_a_linear_field_and_finite_differences(golden, gold2):
    """linear precision: z = 2 + 3x - 5y has gradient (3, -5) in every leaf without a cage vertex;
    and the golden gradient agrees with finite differences of the golden interpolant"""
    xa, ya, _, order = _cfg1()
    za = (2.0 + 3.0 * xa - 5.0 * ya)[None, :]
    tree = {k: golden["cfg1_" + k] for k in ("pts3", "lnk3", "type", "shuffle", "shift", "scale", "seed_points")}
    port = refimpl.PortTree(tree, xa, ya)
    xq, yq = synth.uniform_queries(10000)
    leaf, dx, dy = port.eval_deriv(xq, yq, za)
    real = (tree["pts3"].reshape(-1, 3)[leaf] >= 0).all(axis=1)
    assert real.mean() > 0.95
    assert np.abs(dx[0, real] - 3.0).max() < 1e-9 and np.abs(dy[0, real] + 5.0).max() < 1e-9
    z1 = synth.values(1000, 1)
    h = 1e-7
    l0, v0, _ = port.eval(xq, yq, z1)
    l1, v1, _ = port.eval(xq + h, yq, z1)
    same = l0 == l1
    assert same.mean() > 0.99
    assert np.abs((v1[0] - v0[0])[same] / h - gold2["cfg1_dzdx"][0][same]).max() < 1e-3


def test_type_descriptor_is_exported_and_wired():
    T = g.type_descriptor()
    assert T.name == b"linear-simplex-cuda" and T.min_size == 3
    L = _lib.load()
    assert L.gsl_scattered_cuda_type_min_size(ctypes.byref(T)) == 3
    data_sym = ctypes.c_void_p.in_dll(L, "gsl_scattered_cuda_linear_simplex")     # the extern data symbol
    assert data_sym.value == L.gsl_scattered_cuda_type_linear_simplex()
    for member, sym in (("alloc", "gsl_scattered_cuda_alloc"), ("init", "gsl_scattered_cuda_init"),
                        ("eval_e", "gsl_scattered_cuda_eval_e"), ("eval_deriv_x_e", "gsl_scattered_cuda_eval_deriv_x_e"),
                        ("eval_deriv_y_e", "gsl_scattered_cuda_eval_deriv_y_e"),
                        ("eval_deriv_xx_e", "gsl_scattered_cuda_eval_deriv_xx_e"),
                        ("eval_deriv_xy_e", "gsl_scattered_cuda_eval_deriv_xy_e"),
                        ("eval_deriv_yy_e", "gsl_scattered_cuda_eval_deriv_yy_e"), ("free", "gsl_scattered_cuda_free")):
        got = ctypes.cast(getattr(T, member), ctypes.c_void_p).value
        want = ctypes.cast(getattr(L, sym), ctypes.c_void_p).value
        assert got == want, member
    # through the descriptor, host side only: alloc / name / min_size / free
    w = T.alloc(10, 1)
    assert w and L.gsl_scattered_cuda_name(w) == b"linear-simplex-cuda" and L.gsl_scattered_cuda_min_size(w) == 3
    x = np.zeros(3)
    assert T.eval_deriv_x_e(w, x.ctypes.data, x.ctypes.data, 3, x.ctypes.data) == g.EFAILED   # not initialised
    T.free(w)


# ---- GPU ----------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_derivatives_match_the_golden_vectors(golden, gold2):
    wd = golden["weather"]
    w = g.ScatteredCuda(50, 1)
    w.init(wd[:, 0], wd[:, 1], wd[:, 2][None, :], g.DEFAULT, g.shuffle_mt19937(0, 50))
    q = golden["weatherB_q"]
    st, dx, dy = w.eval_deriv_e(q[:, 0], q[:, 1])
    assert st == g.EDOM                                  # (1e6, 0) is outside the cage
    assert refimpl.same_bits(dx, gold2["weatherB_dzdx"]) and refimpl.same_bits(dy, gold2["weatherB_dzdy"])
    assert w.eval_deriv_x(-87.5, 42.0) == gold2["weatherB_dzdx"][0, 2]
    assert w.eval_deriv_y(-87.5, 42.0) == gold2["weatherB_dzdy"][0, 2]
    assert np.isnan(w.eval_deriv_x(1e6, 0.0))
    st, d2 = w.eval_deriv2_e(q[:, 0], q[:, 1], "xy")
    assert st == g.EDOM and np.array_equal(np.isnan(d2[0]), golden["weatherB_leaf"] < 0)
    assert (d2[0][golden["weatherB_leaf"] >= 0] == 0).all()
    assert w.name() == "linear-simplex-cuda" and w.min_size() == 3
    w.free()
    xa, ya, za, order = _cfg1()
    xq, yq = synth.uniform_queries(10000)
    for params in ({}, {"mode": 1}, {"host_chunk": 777}):
        w = g.ScatteredCuda(1000, 1)
        for k, v in params.items():
            w.set_param(k, v)
        w.init(xa, ya, za, g.DEFAULT, order)
        st, dx, dy = w.eval_deriv_e(xq, yq)
        assert st == g.SUCCESS
        assert refimpl.same_bits(dx, gold2["cfg1_dzdx"]) and refimpl.same_bits(dy, gold2["cfg1_dzdy"])
        st, dx, _ = w.eval_deriv_e(xq, yq, want_y=False)
        assert refimpl.same_bits(dx, gold2["cfg1_dzdx"])
        w.free()
    xa, ya, za, order, xq, yq = _cfg3(golden)
    w = g.ScatteredCuda(3000, 3)
    w.init(xa, ya, za, g.DEFAULT, order)
    st, dx, dy = w.eval_deriv_e(xq, yq)
    assert refimpl.same_bits(dx, gold2["cfg3_dzdx"]) and refimpl.same_bits(dy, gold2["cfg3_dzdy"])
    w.free()


@pytest.mark.gpu
def test_gpu_derivatives_through_the_descriptor_and_on_device_arrays(golden, gold2):
    """the gsl_interp2d_type-style members called as function pointers, and the device-pointer form"""
    import torch
    xa, ya, za, order = _cfg1()
    xq, yq = synth.uniform_queries(10000)
    T = g.type_descriptor()
    w = T.alloc(1000, 1)
    order = np.ascontiguousarray(order, dtype=np.uint64)
    assert T.init(w, xa.ctypes.data, ya.ctypes.data, za.ctypes.data, 1000, g.DEFAULT, order.ctypes.data) == g.SUCCESS
    zq = np.empty(10000)
    leaf = np.empty(10000, dtype=np.int32)
    d = np.empty(10000)
    assert T.eval_e(w, xq.ctypes.data, yq.ctypes.data, 10000, zq.ctypes.data, leaf.ctypes.data) == g.SUCCESS
    assert np.array_equal(leaf, golden["cfg1_leaf"]) and refimpl.same_bits(zq[None, :], golden["cfg1_val"])
    assert T.eval_deriv_x_e(w, xq.ctypes.data, yq.ctypes.data, 10000, d.ctypes.data) == g.SUCCESS
    assert refimpl.same_bits(d[None, :], gold2["cfg1_dzdx"])
    assert T.eval_deriv_y_e(w, xq.ctypes.data, yq.ctypes.data, 10000, d.ctypes.data) == g.SUCCESS
    assert refimpl.same_bits(d[None, :], gold2["cfg1_dzdy"])
    assert T.eval_deriv_xx_e(w, xq.ctypes.data, yq.ctypes.data, 10000, d.ctypes.data) == g.SUCCESS and (d == 0).all()
    T.free(w)
    w = g.ScatteredCuda(1000, 1)
    w.init(xa, ya, za, g.DEFAULT, order)
    dev = torch.device("cuda:0")
    dxq, dyq = torch.from_numpy(xq).to(dev), torch.from_numpy(yq).to(dev)
    ddx = torch.empty(10000, dtype=torch.float64, device=dev)
    ddy = torch.empty(10000, dtype=torch.float64, device=dev)
    dl = torch.empty(10000, dtype=torch.int32, device=dev)
    for leaf_ptr in (dl.data_ptr(), None):
        ddx.zero_(); ddy.zero_()
        w.eval_deriv_device_e(0, dxq.data_ptr(), dyq.data_ptr(), 10000, ddx.data_ptr(), ddy.data_ptr(), 10000, leaf_ptr)
        assert refimpl.same_bits(ddx.cpu().numpy()[None, :], gold2["cfg1_dzdx"])
        assert refimpl.same_bits(ddy.cpu().numpy()[None, :], gold2["cfg1_dzdy"])
    assert np.array_equal(dl.cpu().numpy(), golden["cfg1_leaf"])
    w.free()
