"""Pins the oracle: the CPU restatement (oracle/tinsel_oracle.cpp) against
 (1) tests/golden/*.npz, produced by the reference's own src/render.cpp (oracle/_ref detmath
     flavour, tools/make_golden.py) -- always available, also on the GPU box;
 (2) oracle/_ref itself when it is present (per sample, bit exact), and the literal glibc flavour
     statistically."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import abi
import refdrv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENES = ["cornell", "veach", "glass", "meshlight", "motionblur", "gloss", "emitter", "furnace", "conservation", "ajax", "env", "many", "mini", "envmini", "table", "simple"]
f32p = C.POINTER(C.c_float)


def fp(a):
    return a.ctypes.data_as(f32p)


@pytest.fixture(scope="module", autouse=True)
def port_built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(refdrv.ROOT, "oracle"), "port"])


def _bits_equal(a, b):
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("name", SCENES)
def test_port_matches_golden_radiance(name):
    if not os.path.exists(tb.scene_path(name)):
        pytest.skip("snapshot not present")
    g = np.load(os.path.join(GOLD, "scene_%s.npz" % name))
    port = refdrv.PortScene.from_snapshot(tb.scene_path(name))
    port.set_size(int(g["width"]), int(g["height"]))
    for f in g["frames"]:
        rad, ras = port.trace_frame(int(f), 4)
        assert np.array_equal(ras, g["raster_%d" % f])
        assert _bits_equal(rad, g["radiance_%d" % f]), "frame %d" % f
    acc = port.render_seeded(0, 4, 1)
    assert _bits_equal(acc, g["accum_4spp"])
    assert _bits_equal(port.render_normals(), g["normals"])
    port.close()


def test_port_kats_match_golden():
    g = np.load(os.path.join(GOLD, "kats.npz"))
    lib = refdrv.load_port()
    # Random
    for i, s in enumerate(g["rng_seeds"]):
        u = np.zeros(16, np.uint32)
        f = np.zeros(16, np.float32)
        lib.oracle_random_u32(int(s), 16, u.ctypes.data_as(C.POINTER(C.c_uint32)))
        lib.oracle_random_f32(int(s), 16, fp(f))
        assert np.array_equal(u, g["rng_u32"][i]) and np.array_equal(f, g["rng_f32"][i])
    # SURVEY.md 8a KATs: Random(0) -> 4286347612, 3588832032, ...; Random(12345).Rand() = 1048047690
    assert list(g["rng_u32"][0][:4]) == [4286347612, 3588832032, 2916816470, 2627245810]
    assert g["rng_u32"][2][0] == 1048047690
    # BSDF
    mats = []
    for row in g["materials"]:
        m = abi.Material()
        C.memmove(C.byref(m), row.astype(np.float32).ctypes.data, C.sizeof(m))
        mats.append(m)
    for i, m in enumerate(mats):
        assert lib.oracle_material_ior(C.byref(m)) == g["ior"][i]
    for rec in g["bsdf"]:
        m = mats[int(rec[0])]
        etaI, etaO = np.float32(rec[1]), np.float32(rec[2])
        n, v, l = (rec[3:6].astype(np.float32), rec[6:9].astype(np.float32), rec[9:12].astype(np.float32))
        f = np.zeros(3, np.float32)
        pdf = C.c_float()
        lib.oracle_bsdf_eval(C.byref(m), etaI, etaO, fp(n), fp(v), fp(l), fp(f), C.byref(pdf))
        assert _bits_equal(f, rec[12:15].astype(np.float32)) and _bits_equal(np.float32(pdf.value), np.float32(rec[15]))
        ls = np.zeros(3, np.float32)
        spdf, stype, after = C.c_float(), C.c_int(), (C.c_uint32 * 2)()
        lib.oracle_bsdf_sample(C.byref(m), etaI, etaO, fp(n), fp(v), int(rec[16]), fp(ls), C.byref(spdf), C.byref(stype), after)
        assert _bits_equal(np.float32(spdf.value), np.float32(rec[20]))
        assert (after[0], after[1]) == (int(rec[22]), int(rec[23]))
        if spdf.value > 0:
            assert _bits_equal(ls, rec[17:20].astype(np.float32)) and stype.value == int(rec[21])
    # GenerateRay (first row is the SURVEY KAT: d = (-0.0665700957, 0.217102483, -0.973876297))
    cam = abi.Camera()
    cam.position[:] = [0.0, 1.0, 4.0]
    cam.rotation[:] = [0.0, 0.0, 0.0, 1.0]
    cam.fov = float(np.float32(np.deg2rad(np.float32(35.0))))
    for row in g["rays"][:4]:
        o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib.oracle_generate_ray(C.byref(cam), 256, 256, np.float32(row[0]), np.float32(row[1]), fp(o), fp(d))
        assert _bits_equal(o, row[2:5].astype(np.float32)) and _bits_equal(d, row[5:8].astype(np.float32))
    assert np.allclose(g["rays"][0][5:8], [-0.0665700957, 0.217102483, -0.973876297], atol=1e-7)
    cam2 = abi.Camera()
    cam2.position[:] = [1.5, 2.0, -3.0]
    cam2.rotation[:] = list(g["ray_cam2_rot"])
    cam2.fov = 0.9
    for row in g["rays"][4:]:
        o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib.oracle_generate_ray(C.byref(cam2), 640, 480, np.float32(row[0]), np.float32(row[1]), fp(o), fp(d))
        assert _bits_equal(o, row[2:5].astype(np.float32)) and _bits_equal(d, row[5:8].astype(np.float32))
    # Filter (SURVEY KAT: Filter(gauss,.75,1).Eval(.3,-.2) = 0.134564266)
    for row in g["filter"]:
        got = lib.oracle_filter_eval(1, *[np.float32(x) for x in row[:5]])
        assert _bits_equal(np.float32(got), np.float32(row[5]))
    assert abs(g["filter"][0][5] - 0.134564266) < 1e-8


@pytest.mark.skipif(not refdrv.have_ref("detmath"), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["cornell", "veach", "glass", "env"])
def test_port_matches_reference_code_per_sample(name):
    if not os.path.exists(tb.scene_path(name)):
        pytest.skip("snapshot not present")
    ref = refdrv.RefScene.from_snapshot(tb.scene_path(name), "detmath")
    port = refdrv.PortScene.from_snapshot(tb.scene_path(name))
    ref.set_size(80, 60)
    port.set_size(80, 60)
    for f in (1, 9):
        a, ar = ref.trace_frame(f, 4)
        b, br = port.trace_frame(f, 4)
        assert np.array_equal(ar, br) and _bits_equal(a, b)
    ref.close()
    port.close()


@pytest.mark.skipif(not refdrv.have_ref("literal"), reason="oracle/_ref not built")
def test_literal_reference_agrees_statistically():
    """The reference as shipped (glibc libm) vs the detmath restatement: libm rounding differs in
    ~1 % of sinf/cosf calls, so samples differ in the last bits but no path may change."""
    ref = refdrv.RefScene.from_snapshot(tb.scene_path("cornell"), "literal")
    port = refdrv.PortScene.from_snapshot(tb.scene_path("cornell"))
    ref.set_size(96, 96)
    port.set_size(96, 96)
    a = ref.render_seeded(0, 8, 4)
    b = port.render_seeded(0, 8, 4)
    rel = np.linalg.norm((a - b).astype(np.float64)) / np.linalg.norm(a.astype(np.float64))
    assert rel < 1e-5, rel
    # and the literal CpuRenderer (single sequential stream) agrees in the mean
    lit = ref.render_literal(8)
    ma, mb = lit[..., :3].sum() / lit[..., 3].sum(), b[..., :3].sum() / b[..., 3].sum()
    assert abs(ma - mb) / mb < 0.02
    ref.close()
    port.close()


def test_furnace_energy_bound():
    """furnace.tin: sphere inside a radius-5 emitter of radiance 0.5; the estimate must stay
    close to 0.5 (the Disney lobes are not exactly conserving)."""
    port = refdrv.PortScene.from_snapshot(tb.scene_path("furnace"))
    port.set_size(48, 48)
    port.options.maxDepth = 8
    img = port.render_seeded(0, 8, 4)
    mean = img[..., :3].sum() / (3 * img[..., 3].sum())
    assert 0.4 < mean < 0.6, mean
    port.close()


@pytest.mark.skipif(not refdrv.have_ref("detmath"), reason="oracle/_ref not built")
def test_port_sampling_functions_match_reference_code():
    """Function-level known answers straight from the reference's compiled code: ProbeSample /
    Sky::Eval + ProbePdf (probe.h:105-236, scene.h:168-178) on the HDR-probe scene, PrimitiveSample /
    PrimitiveArea (intersection.h:833-904) on sphere and mesh lights, and Trace (render.cpp:17-62)
    for rays the image-level tests never shoot (from inside, grazing, axis-aligned, zero components)."""
    def f3():
        return np.zeros(3, np.float32)

    ref = refdrv.RefScene.from_snapshot(tb.scene_path("envmini"), "detmath")
    port = refdrv.PortScene.from_snapshot(tb.scene_path("envmini"))
    for seed in range(200):
        rd, rc, pd, pc = f3(), f3(), f3(), f3()
        rp, pp = C.c_float(), C.c_float()
        ref.lib.ref_probe_sample(ref.h, seed, fp(rd), fp(rc), C.byref(rp))
        port.lib.oracle_probe_sample(port.h, seed, fp(pd), fp(pc), C.byref(pp))
        assert _bits_equal(rd, pd) and _bits_equal(rc, pc) and np.float32(rp.value) == np.float32(pp.value), seed
    rng = np.random.RandomState(2)
    dirs = rng.normal(size=(300, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    dirs = np.concatenate([dirs, np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, -1], [0.6, 0, 0.8]], np.float32)])
    for d in dirs:
        rc, pc = f3(), f3()
        rp, pp = C.c_float(), C.c_float()
        ref.lib.ref_sky_eval(ref.h, fp(d), fp(rc), C.byref(rp))
        port.lib.oracle_sky_eval(port.h, fp(d), fp(pc), C.byref(pp))
        assert _bits_equal(rc, pc) and np.float32(rp.value) == np.float32(pp.value), d
    ref.close()
    port.close()

    for name, prims in (("veach", range(0, 12)), ("meshlight", range(0, 4)), ("cornell", range(0, 8))):
        ref = refdrv.RefScene.from_snapshot(tb.scene_path(name), "detmath")
        port = refdrv.PortScene.from_snapshot(tb.scene_path(name))
        n = ref.scene.contents.numPrimitives
        for prim in prims:
            if prim >= n or ref.scene.contents.primitives[prim].type == abi.PLANE:
                continue                                  # planes cannot be sampled (assert(0) in the reference)
            for seed in (0, 1, 77):
                ra, pa = C.c_float(), C.c_float()
                rpos, rn, ppos, pn = f3(), f3(), f3(), f3()
                ref.lib.ref_primitive_sample(ref.h, prim, 0.3, seed, fp(rpos), fp(rn), C.byref(ra))
                port.lib.oracle_primitive_sample(port.h, prim, 0.3, seed, fp(ppos), fp(pn), C.byref(pa))
                assert _bits_equal(rpos, ppos) and _bits_equal(rn, pn) and np.float32(ra.value) == np.float32(pa.value), (name, prim, seed)
        # Trace: closest hit, primitive and face-forwarded normal
        rays = [((0.0, 1.0, 0.0), (0.0, 1.0, 0.0)), ((0.0, 1.0, 0.0), (1.0, 0.0, 0.0)), ((0.3, 0.5, 0.2), (0.0, 0.0, -1.0)),
                ((0.0, 1.0, 3.9), (0.0, -0.001, -1.0)), ((0.35, 0.5, 0.0), (0.57735, 0.57735, 0.57735))]
        for o, d in rays + [(tuple(rng.uniform(-0.9, 0.9, 3) + (0, 1, 0)), tuple(x)) for x in dirs[:40]]:
            a = ref.trace(o, d, 0.5)
            b = port.trace(o, d, 0.5)
            assert a[0] == b[0] and np.float32(a[1]) == np.float32(b[1]) and _bits_equal(a[2], b[2]), (name, o, d)
        ref.close()
        port.close()
