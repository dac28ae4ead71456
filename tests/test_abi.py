"""CPU: the C-ABI library loads and exports every symbol include/remap_b200.h declares; the
product path fails loudly without a GPU (no CPU fallback); the host build of the kernel bodies
(tests/emul) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import remap_b200
from remap_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "remap_b200.h")).read()
    body = header[header.index('extern "C" {'):]
    declared = set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", body))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.rb_abi_version() == 4


def test_default_config_is_the_reference_constants():
    lib = _lib.load()
    cfg = _lib.RbConfig()
    lib.rb_default_config(C.byref(cfg), 320, 224, 100)
    # src/frc.hpp:22-24 and :32-33
    assert (cfg.grid_w, cfg.grid_h, cfg.overlap, cfg.weight_switch, cfg.region_votes) == (4, 2, 16, 10, 3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(remap_b200.RemapError) as e:
        remap_b200.Registrar(320, 224, 16)
    assert e.value.code == _lib.RB_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pkg = os.path.join(ROOT, "remap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert not re.search(r"#\s*include[^\n]*oracle", text), f
                assert "libremap_oracle" not in text and "_ref/ref_harness" not in text, f


def test_host_side_nibble_packing_for_the_pcie_copy():
    """rb_register_host_async ships frames as 4 bit/pixel: the host packer (rb_hostpack.cpp, AVX2 + scalar
    tail, multi-threaded) against numpy, including odd widths, padding and dirty high nibbles."""
    lib = _lib.load()
    fn = lib.rb_hostpack_frames
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_uint32, C.c_int]
    rng = np.random.default_rng(5)
    for W, H, n in ((320, 224, 3), (323, 227, 2), (33, 9, 5), (8, 8, 1), (1, 4, 2)):
        frames = rng.integers(0, 256, size=(n, H, W), dtype=np.uint8)
        pitch4 = ((W + 1) // 2 + 15) // 16 * 16
        out = np.full((n, H, pitch4), 0xAB, np.uint8)
        fn(frames.ctypes.data_as(C.c_void_p), W, H, n, out.ctypes.data_as(C.c_void_p), pitch4, 3)
        lo = frames & 15
        padded = np.zeros((n, H, 2 * pitch4), np.uint8)
        padded[:, :, :W] = lo
        want = padded[:, :, 0::2] | (padded[:, :, 1::2] << 4)
        assert np.array_equal(out, want), (W, H)
