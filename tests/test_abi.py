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


def test_create_validates_the_scene_before_touching_the_device(lib):
    """Indices the kernels would follow are range-checked in tb200_create (runs before the CUDA
    device is opened, so this works on a CPU-only box too)."""
    snap = tb.Snapshot(tb.scene_path("glass"))
    base = snap.scene.contents

    def attempt(mutate):
        sc = abi.Scene()
        C.memmove(C.byref(sc), C.byref(base), C.sizeof(abi.Scene))
        keep = mutate(sc)
        h = lib.tb200_create(C.byref(sc), 0)
        if h:
            lib.tb200_destroy(h)
        return bool(h), tb.last_error(), keep

    def bad_prim_mesh(sc):
        prims = (abi.Primitive * sc.numPrimitives)()
        C.memmove(prims, sc.primitives, C.sizeof(prims))
        for p in prims:
            if p.type == 2:
                p.mesh = sc.numMeshes + 3
        sc.primitives = C.cast(prims, C.POINTER(abi.Primitive))
        return prims

    def bad_bvh(sc):
        nodes = (abi.BvhNode * sc.numBvhNodes)()
        C.memmove(nodes, sc.bvhNodes, C.sizeof(nodes))
        nodes[0].left = 10 ** 6
        nodes[0].right_leaf &= 0x7fffffff
        sc.bvhNodes = C.cast(nodes, C.POINTER(abi.BvhNode))
        return nodes

    def bad_index(sc):
        meshes = (abi.Mesh * sc.numMeshes)()
        C.memmove(meshes, sc.meshes, C.sizeof(meshes))
        idx = (C.c_int32 * meshes[0].numIndices)()
        C.memmove(idx, meshes[0].indices, C.sizeof(idx))
        idx[5] = meshes[0].numVertices
        meshes[0].indices = C.cast(idx, C.POINTER(C.c_int32))
        sc.meshes = C.cast(meshes, C.POINTER(abi.Mesh))
        return meshes, idx

    def no_prims(sc):
        sc.numPrimitives = 0

    for mutate, expect in ((bad_prim_mesh, "mesh index out of range"), (bad_bvh, "points outside its array"),
                           (bad_index, "vertex index out of range"), (no_prims, "no primitives")):
        ok, err, _ = attempt(mutate)
        assert not ok and "invalid scene" in err and expect in err, (mutate.__name__, err)
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


@pytest.mark.gpu
def test_error_behaviour_through_the_c_abi(lib):
    """The reference interface has no error channel (void returns, no exceptions): failures come
    back as -1 with a sticky message in tb200_last_error(), the caller's buffer is left untouched
    and the renderer stays usable."""
    snap = tb.Snapshot(tb.scene_path("cornell"))
    cam, opt = snap.camera, snap.options
    opt.width, opt.height = 32, 24
    r = tb.Renderer(snap.scene)
    out = np.full((24, 32, 4), 7.0, np.float32)
    fp = out.ctypes.data_as(C.POINTER(C.c_float))
    # Render before Init
    assert lib.tb200_render(r.h, C.byref(cam), C.byref(opt), fp) == -1
    assert "tb200_init" in tb.last_error() and (out == 7.0).all()
    assert lib.tb200_finish(r.h, 1.0, 1.5, fp, None) == -1
    r.Init(32, 24)
    # options that disagree with the last Init
    opt.width = 40
    assert lib.tb200_render(r.h, C.byref(cam), C.byref(opt), fp) == -1
    assert "differ" in tb.last_error() and (out == 7.0).all()
    opt.width = 32
    # null arguments, bad sizes, bad shards, bad row ranges
    assert lib.tb200_render(r.h, None, C.byref(opt), fp) == -1
    assert lib.tb200_render(r.h, C.byref(cam), C.byref(opt), None) == -1
    assert lib.tb200_init(r.h, 0, 10) == -1
    assert lib.tb200_set_shard(r.h, 2, 2) == -1
    assert lib.tb200_render_device(r.h, C.byref(cam), C.byref(opt), 1, 20, 10) == -1
    assert lib.tb200_render_n(r.h, C.byref(cam), C.byref(opt), 0, fp) == -1
    assert lib.tb200_nlm(r.h, 200.0, 1, fp) == -1            # nothing finished yet
    assert (out == 7.0).all()
    # eComplexity is a no-op (render.cpp:516-519); the renderer still works afterwards
    opt.mode = abi.MODE_COMPLEXITY
    assert lib.tb200_render(r.h, C.byref(cam), C.byref(opt), fp) == 0 and (out == 7.0).all()
    opt.mode = abi.MODE_PATHTRACE
    r.Render(cam, opt, out)
    assert np.isfinite(out).all() and abs(float(out[2:-2, 2:-2, 3].mean()) / float(out[12, 16, 3]) - 1.0) < 0.5
    assert r.stats().frames == 1 and r.stats().samples == 32 * 24
    # a bad device ordinal fails in the factory
    bad = lib.tb200_create(snap.scene, 99)
    assert not bad and "device" in tb.last_error()
    # re-Init with another size resets the accumulation
    r.Init(16, 16)
    opt.width = opt.height = 16
    small = np.zeros((16, 16, 4), np.float32)
    r.Render(cam, opt, small)
    assert r.stats().frames == 1
    r.close()
    snap.close()
