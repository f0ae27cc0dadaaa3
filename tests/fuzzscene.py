"""Random small scenes (planes, static and moving spheres, sphere lights, the whole Disney parameter
space, gradient sky) written as .tsnap snapshots, for the fuzz parity tests: the same file is loaded
into a reference `Scene` (oracle/_ref), the CPU restatement and the CUDA renderer.  The scene-level
BVH is a median-split tree built here; any valid BVH gives the three implementations the same walk."""
import ctypes as C

import numpy as np

import tinsel_b200 as tb
from tinsel_b200 import abi

LEAF = 1 << 31


def _material(rng, light):
    m = abi.Material.default()
    if light:
        e = float(rng.uniform(4.0, 40.0))
        m.emission[:] = [e, e * float(rng.uniform(0.6, 1.0)), e * float(rng.uniform(0.4, 1.0))]
        m.color[:] = [0.0, 0.0, 0.0]
        m.specular = 0.0
        return m
    m.color[:] = [float(x) for x in rng.uniform(0.05, 1.0, 3)]
    m.metallic = float(rng.choice([0.0, 1.0, rng.uniform()]))
    m.subsurface = float(rng.choice([0.0, rng.uniform()]))
    m.specular = float(rng.uniform())
    m.roughness = float(rng.uniform(0.02, 1.0))
    m.specularTint = float(rng.uniform())
    m.sheen = float(rng.choice([0.0, rng.uniform()]))
    m.sheenTint = float(rng.uniform())
    m.clearcoat = float(rng.choice([0.0, rng.uniform()]))
    m.clearcoatGloss = float(rng.uniform())
    m.transmission = float(rng.choice([0.0, 0.0, rng.uniform(), 1.0]))
    if m.transmission > 0.0:
        m.eta = float(rng.uniform(1.1, 1.9))
        m.absorption[:] = [float(x) for x in rng.choice([0.0, 1.0]) * rng.uniform(0.0, 2.0, 3)]
    return m


def _xf(t, p, q, s):
    t.p[:] = [float(x) for x in p]
    t.r[:] = [float(x) for x in q]
    t.s = float(s)


def _build_bvh(bounds):
    """Median split over the x centre; preorder array of (lower, upper, left, right_leaf)."""
    nodes = []

    def rec(items):
        me = len(nodes)
        nodes.append(None)
        lo = np.min([bounds[i][0] for i in items], axis=0)
        hi = np.max([bounds[i][1] for i in items], axis=0)
        if len(items) == 1:
            nodes[me] = (lo, hi, items[0], LEAF)
            return me
        items = sorted(items, key=lambda i: float(bounds[i][0][0] + bounds[i][1][0]))
        half = len(items) // 2
        left = rec(items[:half])
        right = rec(items[half:])
        nodes[me] = (lo, hi, left, right)
        return me

    rec(list(range(len(bounds))))
    return nodes


def make_scene(seed, path, width=40, height=30):
    """Writes a random scene to `path` (a .tsnap) and returns a short description."""
    rng = np.random.RandomState(seed)
    lib = tb.load_library()
    nspheres = int(rng.randint(3, 7))
    nlights = int(rng.randint(1, 3))
    nplanes = int(rng.randint(1, 3))
    n = nplanes + nspheres + nlights
    prims = (abi.Primitive * n)()
    bounds = []
    k = 0
    for i in range(nplanes):
        p = prims[k]
        p.type = abi.PLANE
        p.plane[:] = [0.0, 1.0, 0.0, 0.0] if i == 0 else [0.0, 0.0, 1.0, 3.0]
        _xf(p.start, (0, 0, 0), (0, 0, 0, 1), 1.0)
        _xf(p.end, (0, 0, 0), (0, 0, 0, 1), 1.0)
        p.mesh = -1
        p.material = _material(rng, False)
        p.material.transmission = 0.0
        bounds.append((np.full(3, -1.0e8, np.float32), np.full(3, 1.0e8, np.float32)))   # PrimitiveBounds, intersection.h:919-924
        k += 1
    for i in range(nspheres + nlights):
        light = i >= nspheres
        p = prims[k]
        p.type = abi.SPHERE
        p.radius = float(rng.uniform(0.2, 0.7)) if not light else float(rng.uniform(0.15, 0.4))
        pos = np.array([rng.uniform(-2.2, 2.2), rng.uniform(0.3, 1.8) if not light else rng.uniform(2.0, 3.5), rng.uniform(-2.0, 1.5)])
        q = rng.normal(size=4)
        q = q / np.linalg.norm(q) if rng.uniform() < 0.5 else np.array([0.0, 0.0, 0.0, 1.0])
        s = float(rng.choice([1.0, rng.uniform(0.6, 1.5)]))
        end = pos + (rng.uniform(-0.4, 0.4, 3) if (rng.uniform() < 0.3 and not light) else 0.0)
        _xf(p.start, pos, q, s)
        _xf(p.end, end, q, s)
        p.mesh = -1
        p.material = _material(rng, light)
        p.lightSamples = int(rng.randint(1, 3)) if light else 0
        r = p.radius * s * 1.001 + 1e-4
        bounds.append((np.minimum(pos, end).astype(np.float32) - np.float32(r), np.maximum(pos, end).astype(np.float32) + np.float32(r)))
        k += 1
    tree = _build_bvh(bounds)
    nodes = (abi.BvhNode * len(tree))()
    for i, (lo, hi, left, right) in enumerate(tree):
        nodes[i].lower[:] = [float(x) for x in lo]
        nodes[i].upper[:] = [float(x) for x in hi]
        nodes[i].left = int(left)
        nodes[i].right_leaf = int(right)
    scene = abi.Scene()
    scene.primitives = C.cast(prims, C.POINTER(abi.Primitive))
    scene.numPrimitives = n
    scene.numMeshes = 0
    scene.bvhNodes = C.cast(nodes, C.POINTER(abi.BvhNode))
    scene.numBvhNodes = len(tree)
    scene.sky.horizon[:] = [float(x) for x in rng.uniform(0.0, 1.0, 3)]
    scene.sky.zenith[:] = [float(x) for x in rng.uniform(0.0, 1.0, 3)]
    base = tb.Snapshot(tb.scene_path("cornell"))
    cam, opt = abi.copy_struct(base.camera), abi.copy_struct(base.options)
    base.close()
    cam.position[:] = [0.0, 1.2, 6.0]
    cam.rotation[:] = [0.0, 0.0, 0.0, 1.0]
    cam.fov = float(np.radians(rng.uniform(35.0, 60.0))) if cam.fov < 3.2 else float(rng.uniform(35.0, 60.0))
    cam.shutterStart, cam.shutterEnd = 0.0, 1.0
    opt.width, opt.height = width, height
    opt.maxDepth = int(rng.randint(2, 6))
    opt.filterType, opt.filterWidth, opt.filterFalloff, opt.filterOffset = abi.FILTER_GAUSSIAN, 0.75, 1.0, 0.5697828531265259
    if lib.tb200_snapshot_save(str(path).encode(), C.byref(scene), C.byref(cam), C.byref(opt)) != 0:
        raise RuntimeError(tb.last_error())
    return "seed %d: %d planes, %d spheres, %d lights, maxDepth %d" % (seed, nplanes, nspheres, nlights, opt.maxDepth)
