#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own code (oracle/_ref, detmath flavour for
anything that touches libm, literal flavour for the RNG-free / libm-free vectors).

Runs in the build container only.  The fixtures pin the oracle port (tests/test_oracle.py) and the
CUDA path (tests/test_golden_gpu.py) to outputs of tinsel's src/render.cpp, and travel to the GPU
box with the repo.  Re-run after changing the seed rule or include/tb200_detmath.h.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdrv  # noqa: E402
from tinsel_b200 import abi  # noqa: E402
import tinsel_b200 as tb  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# scene -> (width, height, frames)
FRAMES = {
    "cornell": (64, 64, (0, 7)),
    "veach": (64, 48, (0, 3)),
    "glass": (48, 48, (0, 3)),
    "meshlight": (48, 48, (1,)),
    "motionblur": (48, 48, (2,)),
    "gloss": (48, 48, (0,)),
    "emitter": (32, 32, (0,)),
    "furnace": (32, 32, (0,)),
    "conservation": (32, 16, (0,)),
    "ajax": (48, 48, (0, 4)),
    "env": (48, 48, (0, 4)),
    "many": (48, 32, (0, 2)),
    "mini": (48, 32, (0, 2)),
    "envmini": (64, 48, (0, 3)),
    "table": (64, 40, (0, 2)),
    "simple": (48, 24, (0,)),
}

f32p = C.POINTER(C.c_float)


def fp(a):
    return a.ctypes.data_as(f32p)


def scene_fixtures(only=()):
    for name, (w, h, frames) in FRAMES.items():
        if only and name not in only:
            continue
        path = tb.scene_path(name)
        if not os.path.exists(path):
            print("skip", name)
            continue
        ref = refdrv.RefScene.from_snapshot(path, "detmath")
        ref.set_size(w, h)
        data = {"width": w, "height": h, "frames": np.array(frames)}
        for f in frames:
            rad, ras = ref.trace_frame(f, 8)
            data["radiance_%d" % f] = rad
            data["raster_%d" % f] = ras
        data["accum_4spp"] = ref.render_seeded(0, 4, 1)
        lit = refdrv.RefScene.from_snapshot(path, "literal")
        lit.set_size(w, h)
        lit.set_mode(abi.MODE_NORMALS)
        data["normals"] = lit.render_literal(1)
        np.savez_compressed(os.path.join(GOLD, "scene_%s.npz" % name), **data)
        print("wrote scene_%s.npz" % name)
        ref.close()
        lit.close()


def material_set():
    mats = []
    rng = np.random.RandomState(7)
    base = abi.Material.default()
    mats.append(base)
    for _ in range(23):
        m = abi.Material.default()
        m.color[:] = list(rng.rand(3))
        m.metallic = float(rng.rand() < 0.3) * rng.rand()
        m.subsurface = float(rng.rand() < 0.3) * rng.rand()
        m.specular = rng.rand()
        m.roughness = rng.rand()
        m.specularTint = rng.rand()
        m.clearcoat = float(rng.rand() < 0.4) * rng.rand()
        m.clearcoatGloss = rng.rand()
        m.transmission = float(rng.rand() < 0.5) * rng.rand()
        if rng.rand() < 0.3:
            m.transmission = 1.0
        m.eta = 0.0 if rng.rand() < 0.5 else 1.0 + rng.rand()
        mats.append(m)
    return mats


def unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def kat_fixtures():
    lib = refdrv.load_ref("detmath")
    out = {}
    # Random (maths.h:1036-1091)
    seeds = np.array([0, 1, 12345, -7, 2 ** 31 - 1], np.int32)
    u = np.zeros((len(seeds), 16), np.uint32)
    f = np.zeros((len(seeds), 16), np.float32)
    for i, s in enumerate(seeds):
        lib.ref_random_u32(int(s), 16, u[i].ctypes.data_as(C.POINTER(C.c_uint32)))
        lib.ref_random_f32(int(s), 16, fp(f[i]))
    out.update(rng_seeds=seeds, rng_u32=u, rng_f32=f)

    # BSDF eval / pdf / sample (disney.h)
    mats = material_set()
    rng = np.random.RandomState(11)
    recs = []
    for mi, m in enumerate(mats):
        for k in range(24):
            n = unit(rng.randn(3))
            v = unit(rng.randn(3))
            if np.dot(v, n) < 0:
                v = -v
            l = unit(rng.randn(3))
            etaI, etaO = (1.0, lib.ref_material_ior(C.byref(m))) if k % 3 else (lib.ref_material_ior(C.byref(m)), 1.0)
            fv = np.zeros(3, np.float32)
            pdf = C.c_float()
            lib.ref_bsdf_eval(C.byref(m), etaI, etaO, fp(n), fp(v), fp(l), fp(fv), C.byref(pdf))
            ls = np.zeros(3, np.float32)
            spdf = C.c_float()
            stype = C.c_int()
            after = (C.c_uint32 * 2)()
            seed = int(rng.randint(0, 2 ** 31 - 1))
            lib.ref_bsdf_sample(C.byref(m), etaI, etaO, fp(n), fp(v), seed, fp(ls), C.byref(spdf), C.byref(stype), after)
            recs.append((mi, etaI, etaO, *n, *v, *l, *fv, pdf.value, seed, *ls, spdf.value, stype.value, after[0], after[1]))
    out["bsdf"] = np.array(recs, np.float64)
    out["materials"] = np.array([[*m.emission, *m.color, *m.absorption, m.eta, m.metallic, m.subsurface, m.specular,
                                  m.roughness, m.specularTint, m.anisotropic, m.sheen, m.sheenTint, m.clearcoat,
                                  m.clearcoatGloss, m.transmission] for m in mats], np.float32)
    out["ior"] = np.array([lib.ref_material_ior(C.byref(m)) for m in mats], np.float32)

    # GenerateRay (util.h:49-79)
    cam = abi.Camera()
    cam.position[:] = [0.0, 1.0, 4.0]
    cam.rotation[:] = [0.0, 0.0, 0.0, 1.0]
    cam.fov = float(np.float32(np.deg2rad(np.float32(35.0))))
    rays = []
    for (x, y) in [(100.25, 37.5), (0.0, 0.0), (255.9, 255.1), (128.0, 128.0)]:
        o = np.zeros(3, np.float32)
        d = np.zeros(3, np.float32)
        lib.ref_generate_ray(C.byref(cam), 256, 256, x, y, fp(o), fp(d))
        rays.append((x, y, *o, *d))
    cam2 = abi.Camera()
    cam2.position[:] = [1.5, 2.0, -3.0]
    q = unit([0.1, 0.7, -0.2, 0.6])
    cam2.rotation[:] = list(q)
    cam2.fov = 0.9
    for (x, y) in [(10.5, 400.25), (639.0, 1.0)]:
        o = np.zeros(3, np.float32)
        d = np.zeros(3, np.float32)
        lib.ref_generate_ray(C.byref(cam2), 640, 480, x, y, fp(o), fp(d))
        rays.append((x, y, *o, *d))
    out["rays"] = np.array(rays, np.float64)
    out["ray_cam2_rot"] = q

    # Filter::Eval (render.h:13-39) with the loader's stale offset (loader.cpp:75)
    fl = []
    for (w, fo, off) in [(0.75, 1.0, 0.569782853), (1.0, 1.0, 0.569782853), (1.0, 2.0, 0.135335283)]:
        for (x, y) in [(0.3, -0.2), (0.0, 0.0), (0.9, 0.1), (-0.74, 0.74), (1.2, 0.0)]:
            fl.append((w, fo, off, x, y, lib.ref_filter_eval(1, w, fo, off, x, y)))
    out["filter"] = np.array(fl, np.float64)
    np.savez_compressed(os.path.join(GOLD, "kats.npz"), **out)
    print("wrote kats.npz")


def finish_inputs():
    """Accumulator values for the finish-step fixtures: a rendered image plus adversarial sums
    (zero weight -> inf/NaN scale, negative, huge, denormal, values around the 0.004 toe)."""
    g = np.load(os.path.join(GOLD, "scene_cornell.npz"))
    img = g["accum_4spp"].astype(np.float32)                    # 64x64x4
    rng = np.random.RandomState(11)
    adv = np.zeros((16, 64, 4), np.float32)
    adv[..., 3] = rng.uniform(0.5, 6.0, adv.shape[:2])
    adv[..., :3] = rng.uniform(0.0, 1.0, adv.shape[:2] + (3,)) ** 4 * adv[..., 3:4] * rng.choice([0.01, 0.3, 1.0, 8.0, 200.0], adv.shape[:2] + (1,))
    adv[0, :8] = 0.0                                            # w == 0: exposure / 0 = inf, 0 * inf = NaN
    adv[0, 8:16, :3] = 1.0
    adv[0, 8:16, 3] = 0.0                                       # positive sums over zero weight: inf
    adv[1, :8, :3] = -0.5                                       # negative radiance sums
    adv[1, 8:16, :3] = 1e30
    adv[1, 16:24, :3] = 1e-42                                   # denormal
    adv[1, 24:32, :3] = np.float32(0.004) * adv[1, 24:32, 3:4]  # the toe of the filmic curve
    adv[1, 32:40, 0] = np.nan
    adv[1, 40:48, 3] = -1.0                                     # negative weight
    return np.concatenate([img, adv], axis=0)                  # 80 x 64 x 4


def finish_fixtures():
    import tempfile
    pixels = finish_inputs()
    data = {"pixels": pixels, "exposures": np.array([1.0, 0.25, 3.5], np.float32)}
    with tempfile.TemporaryDirectory() as tmp:
        for k, e in enumerate(data["exposures"]):
            f = refdrv.ref_finish(pixels, float(e))
            data["filtered_%d" % k] = f
            data["rgb8_%d" % k] = refdrv.ref_png_bytes(f, os.path.join(tmp, "f%d.png" % k))
        # NonLocalMeansFilter on the finished image (src/main.cpp:275: falloff 200, radius 1), plus a wider window
        img = data["filtered_0"][:64]
        data["nlm_r1"] = refdrv.ref_nlm(img, 200.0, 1)
        data["nlm_r3"] = refdrv.ref_nlm(img, 35.0, 3)
        data["nlm_r0"] = refdrv.ref_nlm(img, 200.0, 0)
    np.savez_compressed(os.path.join(GOLD, "finish.npz"), **data)
    print("wrote finish.npz")


if __name__ == "__main__":
    # python tools/make_golden.py [scene ...]: no arguments regenerates everything
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ["finish"]:
        finish_fixtures()
        sys.exit(0)
    if len(sys.argv) == 1:
        kat_fixtures()
    scene_fixtures(sys.argv[1:])
    if len(sys.argv) == 1:
        finish_fixtures()
