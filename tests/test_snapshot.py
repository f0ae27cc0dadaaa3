"""Scene snapshots: round trip, and (build container only) equality with what the reference's
own loader produces, including the probe importance tables the snapshot loader rebuilds."""
import ctypes as C
import os

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import abi
import refdrv


def _arr(ptr, n, dtype=np.float32):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)


def _nodes(ptr, n):
    """BVHNode words with the right-child index of leaves masked: the reference builder leaves
    it uninitialised (bvh.h leaf nodes only set leftIndex and the leaf bit)."""
    w = _arr(ptr, n * 8, np.uint32).reshape(n, 8).copy()
    leaf = (w[:, 7] >> 31) == 1
    w[leaf, 7] = 0x80000000
    return w


def test_snapshot_round_trip(tmp_path):
    snap = tb.Snapshot(tb.scene_path("cornell"))
    out = str(tmp_path / "copy.tsnap")
    cam, opt = snap.camera, snap.options
    rc = snap.lib.tb200_snapshot_save(out.encode(), snap.scene, C.byref(cam), C.byref(opt))
    assert rc == 0
    assert open(out, "rb").read() == open(tb.scene_path("cornell"), "rb").read()
    snap.close()


def test_cornell_snapshot_contents():
    snap = tb.Snapshot(tb.scene_path("cornell"))
    s = snap.scene.contents
    assert s.numPrimitives == 8 and s.numMeshes == 1 and s.numBvhNodes == 15
    types = [s.primitives[i].type for i in range(8)]
    assert types == [abi.PLANE] * 5 + [abi.MESH, abi.SPHERE, abi.SPHERE]
    assert s.primitives[5].lightSamples == 1
    o = snap.options
    # loader quirk: `filter gaussian 1.0 1.0` keeps the default Filter(0.75,1.0) offset exp(-0.5625)
    assert o.filterWidth == 1.0 and o.filterFalloff == 1.0
    assert abs(o.filterOffset - 0.569782853) < 1e-7
    assert o.maxDepth == 4
    snap.close()


@pytest.mark.skipif(not (refdrv.have_reference_tree() and refdrv.have_ref("literal")), reason="needs /root/reference")
@pytest.mark.parametrize("name,tin", [("cornell", "cornell.tin"), ("veach", "veach.tin"), ("glass", "glass.tin")])
def test_snapshot_equals_reference_loader(name, tin):
    ref = refdrv.RefScene.from_tin(os.path.join(refdrv.REFERENCE_ROOT, "data", tin), flavour="literal")
    snap = tb.Snapshot(tb.scene_path(name))
    a, b = ref.scene.contents, snap.scene.contents
    assert (a.numPrimitives, a.numMeshes, a.numBvhNodes) == (b.numPrimitives, b.numMeshes, b.numBvhNodes)
    n = a.numPrimitives * C.sizeof(abi.Primitive)
    assert bytes(_arr(a.primitives, n, np.uint8)) == bytes(_arr(b.primitives, n, np.uint8))
    assert np.array_equal(_nodes(a.bvhNodes, a.numBvhNodes), _nodes(b.bvhNodes, b.numBvhNodes))
    for m in range(a.numMeshes):
        ma, mb = a.meshes[m], b.meshes[m]
        assert (ma.numVertices, ma.numIndices, ma.numNodes, ma.area) == (mb.numVertices, mb.numIndices, mb.numNodes, mb.area)
        assert np.array_equal(_arr(ma.positions, ma.numVertices * 3), _arr(mb.positions, mb.numVertices * 3))
        assert np.array_equal(_nodes(ma.nodes, ma.numNodes), _nodes(mb.nodes, mb.numNodes))
        assert np.array_equal(_arr(ma.cdf, ma.numIndices // 3), _arr(mb.cdf, mb.numIndices // 3))
    ref.close()
    snap.close()


@pytest.mark.skipif(not (refdrv.have_reference_tree() and refdrv.have_ref("literal") and os.path.exists(tb.scene_path("env"))),
                    reason="needs /root/reference and scenes/env.tsnap")
def test_probe_tables_rebuilt_bit_exact():
    """snapshot.cpp restates Probe::BuildCDF (probe.h:31-79); compare with the reference's tables."""
    ref = refdrv.RefScene.from_tin(os.path.join(refdrv.REFERENCE_ROOT, "data", "env.tin"), flavour="literal")
    snap = tb.Snapshot(tb.scene_path("env"))
    a, b = ref.scene.contents.sky, snap.scene.contents.sky
    assert (a.probeValid, a.probeWidth, a.probeHeight) == (b.probeValid, b.probeWidth, b.probeHeight)
    n = a.probeWidth * a.probeHeight
    assert np.array_equal(_arr(a.probeData, n * 4), _arr(b.probeData, n * 4))
    for f, cnt in (("pdfValuesX", n), ("cdfValuesX", n), ("pdfValuesY", a.probeHeight), ("cdfValuesY", a.probeHeight)):
        assert np.array_equal(_arr(getattr(a, f), cnt).view(np.uint32), _arr(getattr(b, f), cnt).view(np.uint32)), f
    ref.close()
    snap.close()


def test_damaged_snapshots_are_refused_not_crashed(tmp_path):
    """tb200_snapshot_load checks every count against the bytes left in the file before sizing an array from
    it, and lets no exception cross the C ABI: truncated files, huge or negative counts give NULL + an error."""
    import struct
    data = open(tb.scene_path("glass"), "rb").read()
    lib = tb.load_library()

    def attempt(blob):
        p = str(tmp_path / "bad.tsnap")
        open(p, "wb").write(blob)
        h = lib.tb200_snapshot_load(p.encode())
        if h:
            lib.tb200_snapshot_free(h)
        return bool(h), tb.last_error()

    assert attempt(data)[0]
    for cut in (4, 20, 60, len(data) // 3, len(data) - 5):
        ok, err = attempt(data[:cut])
        assert not ok and "snapshot" in err, cut
    # header: magic(8) + 6 x uint32 {primitives, meshes, bvh nodes, probe flag, probe w, probe h}
    for field, value in ((0, 0x7FFFFFFF), (1, 0x40000000), (2, 0xFFFFFFFF), (3, 1)):
        hdr = bytearray(data)
        struct.pack_into("<I", hdr, 8 + 4 * field, value)
        if field == 3:
            struct.pack_into("<II", hdr, 8 + 16, 60000, 60000)   # a probe far larger than the file
        ok, err = attempt(bytes(hdr))
        assert not ok, field
    # a mesh with a negative vertex count: first mesh record follows the primitives and the scene BVH
    nprim, nmesh, nnodes = struct.unpack_from("<III", data, 8)
    off = 8 + 24 + 40 + 48 + 24 + nprim * 180 + nnodes * 32   # sizeof(tb200_primitive) == 180
    bad = bytearray(data)
    struct.pack_into("<i", bad, off, -5)
    ok, err = attempt(bytes(bad))
    assert not ok
