"""The C-ABI library loads, exports every symbol include/gsl_scattered_cuda.h declares, and
its host-side behaviour (argument checks, error hook, no-GPU failure mode, generators)
follows the header.  No compute calls need a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

import gsl_scattered_interpolation_b200 as g
from gsl_scattered_interpolation_b200 import _lib, synth


def test_every_declared_symbol_is_exported():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(gsl_scattered_cuda_[a-z0-9_]+)\s*\(", text))
    declared.discard("gsl_scattered_cuda_error_handler_t")
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_library_is_sm100a_native():
    """the shipped .so carries sm_100a SASS (no PTX-JIT, no other arch)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert not re.search(r"sm_(?!100a)\d+", out.stdout)


def test_alloc_free_and_argument_checks():
    L = _lib.load()
    seen = []

    @_lib.ERROR_HANDLER
    def handler(reason, file, line, code):
        seen.append((reason.decode(), code))

    old = L.gsl_scattered_cuda_set_error_handler(ctypes.cast(handler, ctypes.c_void_p))
    try:
        L.gsl_scattered_cuda_free(None)                       # free(NULL) is a no-op
        assert not L.gsl_scattered_cuda_alloc(10, 0)          # n_fields == 0
        assert seen and seen[-1][1] == g.EINVAL
        w = L.gsl_scattered_cuda_alloc(10, 2)
        assert w
        assert L.gsl_scattered_cuda_set_param(w, b"no_such_knob", 1.0) == g.EINVAL
        assert L.gsl_scattered_cuda_set_param(w, b"cells_per_leaf", 2.0) == g.SUCCESS
        assert L.gsl_scattered_cuda_upload(w, None, 2) == g.EINVAL       # upload before build
        x = np.zeros(3)
        assert L.gsl_scattered_cuda_eval_e(w, x.ctypes.data, x.ctypes.data, 3, None, None) == g.EFAILED
        assert np.isnan(L.gsl_scattered_cuda_eval(w, 0.0, 0.0))
        dev = (ctypes.c_int * 2)(0, 0)
        assert L.gsl_scattered_cuda_set_devices(w, dev, 2) == g.EINVAL   # duplicate ids
        L.gsl_scattered_cuda_free(w)
    finally:
        L.gsl_scattered_cuda_set_error_handler(old)
    assert len(seen) >= 5


def test_no_gpu_means_failure_not_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    xa, ya = synth.uniform_nodes(100)
    w = g.ScatteredCuda(100, 1)
    w.build(xa, ya)                                   # host build works without a GPU
    with pytest.raises(g.GslError) as e:
        w.upload(synth.values(100, 1))                # ... but there is no CPU evaluation path
    assert e.value.status == g.EFAILED
    w.free()


def test_host_generator_matches_numpy_generator():
    xq, yq = g.generate_uniform_host(synth.SEED_QUERIES, 1234, 5000, 0.0, 1.0, 0.0, 1.0)
    nx, ny = synth.uniform_queries(5000, first=1234)
    assert np.array_equal(xq, nx) and np.array_equal(yq, ny)
    xq, yq = g.generate_uniform_host(7, 0, 1000, -3.0, 5.0, 10.0, 10.5)
    nx, ny = synth.uniform_queries(1000, seed=7, box=(-3.0, 5.0, 10.0, 10.5))
    assert np.array_equal(xq, nx) and np.array_equal(yq, ny)


def test_build_reports_info_without_gpu():
    xa, ya = synth.uniform_nodes(1000)
    w = g.ScatteredCuda(1000, 1)
    w.build(xa, ya, g.DEFAULT, g.shuffle_mt19937(synth.SEED_SHUFFLE, 1000))
    info = w.info()
    assert info["n_nodes"] == 1000 and info["n_simplexes"] == 8967 and info["n_devices"] == 0
    assert info["build_seconds"] > 0
    w.free()


def test_slab_file_argument_and_format_checks(tmp_path):
    """gsl_scattered_cuda_save/load: the host-side checks need no GPU"""
    L = _lib.load()
    w = L.gsl_scattered_cuda_alloc(10, 1)
    assert L.gsl_scattered_cuda_save(w, os.fsencode(tmp_path / "x.gsltri")) == g.EINVAL   # nothing uploaded
    assert L.gsl_scattered_cuda_load(w, os.fsencode(tmp_path / "missing.gsltri")) == g.EFAILED
    junk = tmp_path / "junk.gsltri"
    junk.write_bytes(b"\x00" * 4096)
    assert L.gsl_scattered_cuda_load(w, os.fsencode(junk)) == g.EINVAL                     # bad magic
    short = tmp_path / "short.gsltri"
    short.write_bytes(b"GSISLAB1")
    assert L.gsl_scattered_cuda_load(w, os.fsencode(short)) == g.EFAILED                   # "fread failed"
    assert L.gsl_scattered_cuda_save(None, b"x") == g.EINVAL
    L.gsl_scattered_cuda_free(w)
