#!/usr/bin/env python3
"""Build the REAL reference (kataklinger/remap) hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
product path (remap_b200/, include/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it.

The reference is a header-only MSVC C++20 program.  Its registration path (kpe/kpr/kpm/frc,
plus fde::generate_mask and nic for the full frc loop) compiles under g++ 13 after a small
mechanical MSVC->GCC patch.  This script

  1. copies /root/reference/src/*.hpp into a throw-away temp dir (never into the repo),
  2. applies the patch there (every substitution is asserted to hit),
  3. compiles oracle/ref_harness.cpp (OUR harness, which #includes the patched headers)
     into oracle/_ref/ref_harness  (git-ignored, shipped to the GPU box by gpurun),
  4. deletes the temp dir.

If /root/reference is absent (GPU box) the prebuilt oracle/_ref/ref_harness is used as is.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("REMAP_REFERENCE_SRC", "/root/reference/src")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_BIN = os.path.join(OUT_DIR, "ref_harness")
SHIM_BIN = os.path.join(OUT_DIR, "shim_harness")
REPO = os.path.dirname(HERE)

# (file, old, new, expected_count) -- semantics-preserving MSVC->GCC fixes only.
PATCHES = [
    # GCC rejects a standard attribute in the middle of the decl-specifier-seq.
    ("all.hpp", "inline [[nodiscard]] bool", "[[nodiscard]] inline bool", 2),
    # copy-list-initialisation through an explicit constructor.
    ("all.hpp", "*current_ = {previous_->total_used() << 1};",
     "*current_ = memory_pool{previous_->total_used() << 1};", 1),
    # `{}` is not an __m128i for GCC.
    ("kpe.hpp", "_mm256_castsi128_si256({})", "_mm256_castsi128_si256(_mm_setzero_si128())", 2),
    # unaligned buffers: MSVC emits unaligned moves for __m256i*, GCC does not.
    ("fde.hpp", "using mm_type = __m256i;", "using mm_type = __m256i_u;", 1),
    ("aws.hpp", "using mm_t = __m256i;", "using mm_t = __m256i_u;", 1),
    # arf.hpp: a uint8_t non-type template parameter cannot be deduced from std::array's size_t extent (MSVC lets it).
    ("arf.hpp", "    shift(arr, std::make_index_sequence<Size>{});", "    shift<Size>(arr, std::make_index_sequence<Size>{});", 1),
    ("arf.hpp", "      shift(data_);", "      shift<units_count>(data_);", 1),
    ("arf.hpp", "      return hash_impl(buf.data(),", "      return hash_impl<buffer<Size>::units_count>(buf.data(),", 1),
    # arf.hpp: unaligned buffers behind __m256i* / __m256* (see fde.hpp above).
    ("arf.hpp", "*reinterpret_cast<__m256i const*>(a)", "*reinterpret_cast<__m256i_u const*>(a)", 1),
    ("arf.hpp", "*reinterpret_cast<__m256i const*>(b)", "*reinterpret_cast<__m256i_u const*>(b)", 1),
    ("arf.hpp", "*reinterpret_cast<__m256*>(out)", "*reinterpret_cast<__m256_u*>(out)", 1),
    # AVX-512VL/BW-only spelling of an unaligned 256-bit load; AVX2 spelling is identical.
    ("fde.hpp", "_mm256_loadu_epi8(bcur)",
     "_mm256_loadu_si256(reinterpret_cast<__m256i_u const*>(bcur))", 1),
]

# Dead members of kpe::extractor that MSVC never instantiates (they index details::vec_unit
# with [] which has no operator[]): remove the three get_unit* definitions.
DEAD_MEMBER_RE = re.compile(
    r"\n  inline \[\[nodiscard\]\] __m256i get_unit(?:_low|_hi)?\(.*?\n  \}\n", re.S)

FILES = ["all.hpp", "arf.hpp", "aws.hpp", "cdt.hpp", "cpl.hpp", "cte.hpp", "ctr.hpp", "fde.hpp", "fdf.hpp", "fgm.hpp", "fgs.hpp",
         "frc.hpp", "icd.hpp", "ifd.hpp", "kpe.hpp", "kpm.hpp", "kpr.hpp", "mpb.hpp", "mrl.hpp",
         "nic.hpp", "sid.hpp"]

CXXFLAGS = ["-std=c++20", "-O2", "-mavx2", "-fpermissive", "-w", "-pthread",
            "-include", "functional"]


def patched_tree(dst):
    for name in FILES:
        with open(os.path.join(REF_SRC, name), "r", encoding="utf-8-sig") as f:
            text = f.read()
        for fname, old, new, cnt in PATCHES:
            if fname != name:
                continue
            assert text.count(old) == cnt, (name, old, text.count(old))
            text = text.replace(old, new)
        if name == "kpe.hpp":
            text, n = DEAD_MEMBER_RE.subn("\n", text)
            assert n == 3, n
        with open(os.path.join(dst, name), "w") as f:
            f.write(text)
    # <intrin.h> is the MSVC umbrella header
    with open(os.path.join(dst, "intrin.h"), "w") as f:
        f.write("#pragma once\n#include <immintrin.h>\n")


def build_shim(verbose=True):
    """oracle/_ref/shim_harness: the reference's frc::collector next to include/frc_b200.hpp (the
    collector-shaped front of the C ABI), linked against remap_b200/libremap_b200.so.  Returns the
    binary's path, or None if it cannot be had (no reference sources and no prebuilt binary)."""
    lib = os.path.join(REPO, "remap_b200", "libremap_b200.so")
    if not os.path.isdir(REF_SRC) or not os.path.exists(lib):
        return SHIM_BIN if os.path.exists(SHIM_BIN) else None
    harness = os.path.join(HERE, "shim_harness.cpp")
    srcs = [harness, os.path.join(REPO, "include", "frc_b200.hpp"), os.path.join(REPO, "include", "fdf_b200.hpp"),
            os.path.join(REPO, "include", "fgs_b200.hpp"),
            os.path.join(REPO, "include", "remap_b200.h"), lib, __file__]
    if os.path.exists(SHIM_BIN) and os.path.getmtime(SHIM_BIN) >= max(os.path.getmtime(p) for p in srcs):
        return SHIM_BIN
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="remap_ref_")
    try:
        patched_tree(tmp)
        cmd = (["g++"] + CXXFLAGS + ["-I", tmp, "-I", os.path.join(REPO, "include"), harness, "-o", SHIM_BIN,
                                     "-L", os.path.dirname(lib), "-lremap_b200",
                                     "-Wl,-rpath,$ORIGIN/../../remap_b200"])
        if verbose:
            print("[oracle/_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return SHIM_BIN


PIPE_BIN = os.path.join(OUT_DIR, "pipeline_harness")

# INTEGRATION.md's three substitutions, applied mechanically to a copy of the reference's mpb.hpp in the temp dir
MPB_SUBST = [
    ("namespace mpb {", "namespace mpb_subst {", 1),
    ("} // namespace mpb", "} // namespace mpb_subst", 1),
    ("frc::collector collector{window};", "frc_b200::collector collector{window};", 1),
    ("fgs::splice(fragments.begin(), fragments.end())", "fgs_b200::splice(fragments.begin(), fragments.end())", 1),
    ("fdf::filter(fragments, window, adapter_.get_compression(), cb())",
     "fdf_b200::filter(fragments, window, adapter_.get_compression(), cb())", 1),
    ("#pragma once", '#pragma once\n#include "frc_b200.hpp"\n#include "fgs_b200.hpp"\n#include "fdf_b200.hpp"', 1),
]


def build_pipeline(verbose=True):
    """oracle/_ref/pipeline_harness: mpb::builder::build as it is next to the same header with the three B200 shims
    substituted, on the same frames.  Returns the binary's path or None."""
    lib = os.path.join(REPO, "remap_b200", "libremap_b200.so")
    if not os.path.isdir(REF_SRC) or not os.path.exists(lib):
        return PIPE_BIN if os.path.exists(PIPE_BIN) else None
    harness = os.path.join(HERE, "pipeline_harness.cpp")
    srcs = [harness, lib, __file__] + [os.path.join(REPO, "include", f) for f in
                                      ("frc_b200.hpp", "fgs_b200.hpp", "fdf_b200.hpp", "mpb_b200.hpp", "remap_b200.h")]
    if os.path.exists(PIPE_BIN) and os.path.getmtime(PIPE_BIN) >= max(os.path.getmtime(p) for p in srcs):
        return PIPE_BIN
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="remap_ref_")
    try:
        patched_tree(tmp)
        with open(os.path.join(tmp, "mpb.hpp")) as f:
            text = f.read()
        for old, new, cnt in MPB_SUBST:
            assert text.count(old) == cnt, (old, text.count(old))
            text = text.replace(old, new)
        with open(os.path.join(tmp, "mpb_subst.hpp"), "w") as f:
            f.write(text)
        cmd = (["g++"] + CXXFLAGS + ["-I", tmp, "-I", os.path.join(REPO, "include"), harness, "-o", PIPE_BIN,
                                     "-L", os.path.dirname(lib), "-lremap_b200",
                                     "-Wl,-rpath,$ORIGIN/../../remap_b200"])
        if verbose:
            print("[oracle/_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return PIPE_BIN


def build(verbose=True):
    """Returns the path of the harness binary, or None if it cannot be had."""
    if not os.path.isdir(REF_SRC):
        return OUT_BIN if os.path.exists(OUT_BIN) else None
    harness = os.path.join(HERE, "ref_harness.cpp")
    if os.path.exists(OUT_BIN) and os.path.getmtime(OUT_BIN) >= max(
            os.path.getmtime(harness), os.path.getmtime(__file__)):
        return OUT_BIN
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="remap_ref_")
    try:
        patched_tree(tmp)
        cmd = ["g++"] + CXXFLAGS + ["-I", tmp, harness, "-o", OUT_BIN]
        if verbose:
            print("[oracle/_ref]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT_BIN


if __name__ == "__main__":
    p = build()
    print(build_shim())
    print(build_pipeline())
    print(p if p else "reference sources not available and no prebuilt oracle/_ref")
    sys.exit(0 if p else 1)
