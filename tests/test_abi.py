"""The C-ABI library builds, loads on a CPU-only box, exports every symbol include/tinsel_b200.h
declares, and fails loudly (no CPU fallback) when asked to render without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if build.needs_build():
        build.build_native()
    return tb.load_library()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "tinsel_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tb200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(abi.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # sizes the header implies (all 4-byte fields, pointers 8 bytes)
    assert C.sizeof(abi.Transform) == 32
    assert C.sizeof(abi.Material) == 84
    assert C.sizeof(abi.Primitive) == 32 + 32 + 4 + 4 + 16 + 4 + 84 + 4
    assert C.sizeof(abi.BvhNode) == 32          # identical to the reference BVHNode (bvh.h:21)
    assert C.sizeof(abi.Camera) == 40           # identical to the reference Camera
    assert C.sizeof(abi.Options) == 48          # identical to the reference Options


def test_sample_seed_is_bijective_within_a_frame(lib):
    seeds = {tb.sample_seed(p, 3) for p in range(1 << 16)}
    assert len(seeds) == 1 << 16
    assert tb.sample_seed(10, 0) != tb.sample_seed(10, 1)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    snap = tb.Snapshot(tb.scene_path("cornell"))
    with pytest.raises(tb.TinselB200Error):
        tb.Renderer(snap.scene)
    assert "no CUDA device" in tb.last_error()
    snap.close()


def test_product_does_not_reference_oracle():
    """The product path must not import / link / include anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tinsel_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    s = line.strip()
                    if s.startswith(("//", "#", "*", '"""')) and "include" not in s and "import" not in s:
                        continue
                    if re.search(r'(#include\s+"[^"]*oracle|import\s+refdrv|from\s+oracle|libtinsel_oracle|libtinsel_ref)', s):
                        bad.append((f, s))
    assert not bad, bad
