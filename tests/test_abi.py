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
    assert lib.rb_abi_version() == 2


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
