"""tinsel's binary mesh cache (src/mesh.cpp:809-880) through the C ABI (tb200_mesh_bin_*): files
written by the reference's own ExportMeshToBin load into the same arrays the scene holds; files
written by tb200_mesh_bin_save are byte-identical to the reference writer's and load in the
reference's ImportMeshFromBin.  CPU only (host I/O)."""
import ctypes as C
import os

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import abi
import refdrv


def _arrays(mesh):
    nv, ni, nn = mesh.numVertices, mesh.numIndices, mesh.numNodes
    as_np = np.ctypeslib.as_array
    return {
        "positions": as_np(mesh.positions, (nv * 3,)).copy(),
        "normals": as_np(mesh.normals, (nv * 3,)).copy(),
        "indices": as_np(mesh.indices, (ni,)).copy(),
        "nodes": np.frombuffer(C.string_at(mesh.nodes, nn * 32), np.uint8).copy(),
        "cdf": as_np(mesh.cdf, (ni // 3,)).copy(),
        "area": np.float32(mesh.area),
    }


@pytest.mark.parametrize("name", ["glass", "cornell", "meshlight"])
def test_bin_mesh_round_trips_against_the_reference_writer_and_reader(name, tmp_path):
    if not refdrv.have_ref("literal"):
        pytest.skip("oracle/_ref not built")
    lib = tb.load_library()
    ref = refdrv.RefScene.from_snapshot(tb.scene_path(name), "literal")
    snap = tb.Snapshot(tb.scene_path(name))
    scene = snap.scene.contents
    assert scene.numMeshes >= 1
    for m in range(scene.numMeshes):
        want = _arrays(scene.meshes[m])
        ref_file = str(tmp_path / ("ref%d.bin" % m)).encode()
        our_file = str(tmp_path / ("our%d.bin" % m)).encode()
        assert ref.lib.ref_mesh_bin_export(ref.h, m, ref_file) == 0          # the reference's ExportMeshToBin
        # our reader on the reference's file
        f = lib.tb200_mesh_bin_load(ref_file)
        assert f, tb.last_error()
        got = _arrays(lib.tb200_mesh_bin_mesh(f).contents)
        for k in want:
            assert np.array_equal(got[k].view(np.uint8) if got[k].ndim else got[k], want[k].view(np.uint8) if want[k].ndim else want[k]), k
        # our writer: byte-identical file
        assert lib.tb200_mesh_bin_save(our_file, lib.tb200_mesh_bin_mesh(f)) == 0
        assert open(our_file, "rb").read() == open(ref_file, "rb").read()
        lib.tb200_mesh_bin_free(f)
        # the reference's reader on our file
        rm = ref.lib.ref_mesh_bin_import(our_file)
        assert rm
        counts = (C.c_int * 3)()
        area = C.c_float()
        ref.lib.ref_mesh_bin_info(rm, counts, C.byref(area))
        assert list(counts) == [scene.meshes[m].numVertices, scene.meshes[m].numIndices, scene.meshes[m].numNodes]
        assert np.float32(area.value) == want["area"]
        pos = np.empty(counts[0] * 3, np.float32)
        nrm = np.empty(counts[0] * 3, np.float32)
        idx = np.empty(counts[1], np.int32)
        nodes = np.empty(counts[2] * 32, np.uint8)
        cdf = np.empty(counts[1] // 3, np.float32)
        ref.lib.ref_mesh_bin_copy(rm, pos.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)),
                                  idx.ctypes.data_as(C.POINTER(C.c_int)), nodes.ctypes.data_as(C.c_void_p),
                                  cdf.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(pos.view(np.uint32), want["positions"].view(np.uint32))
        assert np.array_equal(idx, want["indices"]) and np.array_equal(nodes, want["nodes"])
        assert np.array_equal(cdf.view(np.uint32), want["cdf"].view(np.uint32))
        ref.lib.ref_mesh_bin_free(rm)
    ref.close()
    snap.close()


def test_bin_mesh_rejects_malformed_files(tmp_path):
    lib = tb.load_library()
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"\x03\x00\x00\x00\x03\x00\x00\x00\x01\x00\x00\x00" + b"\x00" * 10)   # header promises more than the file holds
    assert not lib.tb200_mesh_bin_load(str(bad).encode())
    assert "not a tinsel .bin mesh" in tb.last_error()
    assert not lib.tb200_mesh_bin_load(str(tmp_path / "missing.bin").encode())
    assert "cannot open" in tb.last_error()
    assert lib.tb200_mesh_bin_save(str(tmp_path / "x.bin").encode(), None) == -1


def test_bin_mesh_feeds_a_scene(tmp_path):
    """A tb200_scene assembled from a cached .bin mesh is the scene the loader would have built:
    saving both as snapshots gives identical files."""
    lib = tb.load_library()
    snap = tb.Snapshot(tb.scene_path("glass"))
    scene = snap.scene.contents
    path = str(tmp_path / "m0.bin").encode()
    assert lib.tb200_mesh_bin_save(path, C.byref(scene.meshes[0])) == 0
    f = lib.tb200_mesh_bin_load(path)
    assert f
    meshes = (abi.Mesh * scene.numMeshes)()
    for m in range(scene.numMeshes):
        C.memmove(C.byref(meshes[m]), C.byref(scene.meshes[m]), C.sizeof(abi.Mesh))
    C.memmove(C.byref(meshes[0]), lib.tb200_mesh_bin_mesh(f), C.sizeof(abi.Mesh))
    rebuilt = abi.Scene()
    C.memmove(C.byref(rebuilt), C.byref(scene), C.sizeof(abi.Scene))
    rebuilt.meshes = C.cast(meshes, C.POINTER(abi.Mesh))
    a, b = str(tmp_path / "a.tsnap").encode(), str(tmp_path / "b.tsnap").encode()
    cam, opt = snap.camera, snap.options
    assert lib.tb200_snapshot_save(a, C.byref(scene), C.byref(cam), C.byref(opt)) == 0
    assert lib.tb200_snapshot_save(b, C.byref(rebuilt), C.byref(cam), C.byref(opt)) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    lib.tb200_mesh_bin_free(f)
    snap.close()
