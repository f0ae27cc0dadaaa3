#!/usr/bin/env python
"""Generate scenes/*.tsnap from the reference's own .tin scenes.

Runs ONLY in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`): the reference's LoadTin + Scene::Build (SAH BVH, mesh import, probe
load) produce the Scene, oracle/ref_driver.cpp flattens it, tinsel_b200/csrc/snapshot.cpp
writes it.  Small snapshots are committed; ajax/env are large (53 MB / 96 MB) and are
git-ignored but still travel to the GPU box with the working tree.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdrv  # noqa: E402

# name -> (tin file, width, height): BASELINE.json configs + the parity-only scenes
SCENES = {
    "cornell": ("cornell.tin", 256, 256),
    "veach": ("veach.tin", 0, 0),
    "ajax": ("ajax.tin", 0, 0),
    "env": ("env.tin", 0, 0),
    "glass": ("glass.tin", 0, 0),
    "transmission": ("transmission.tin", 0, 0),
    "furnace": ("furnace.tin", 0, 0),
    "conservation": ("conservation.tin", 0, 0),
    "meshlight": ("meshlight.tin", 0, 0),
    "motionblur": ("motionblur.tin", 0, 0),
    "gloss": ("gloss.tin", 0, 0),
    "emitter": ("emitter.tin", 0, 0),
    "table": ("table.tin", 0, 0),
    "simple": ("simple.tin", 0, 0),
    # this repository's own .tin scenes (tests/data): >16 primitives / a glass sphere under a gradient sky
    "many": ("@tests/data/many.tin", 0, 0),
    "mini": ("@tests/data/mini0.tin", 0, 0),
    # small synthetic HDR probe (tools/make_test_probe.py): the image-based-lighting path in a snapshot small enough to commit
    "envmini": ("@tests/data/envmini.tin", 0, 0),
}


def main(names):
    os.makedirs(os.path.join(ROOT, "scenes"), exist_ok=True)
    for name in names:
        tin, w, h = SCENES[name]
        path = os.path.join(ROOT, tin[1:]) if tin.startswith("@") else os.path.join(refdrv.REFERENCE_ROOT, "data", tin)
        rs = refdrv.RefScene.from_tin(path, w, h, flavour="literal")
        out = os.path.join(ROOT, "scenes", name + ".tsnap")
        rs.save_snapshot(out)
        s = rs.scene.contents
        print("%-14s %4dx%-4d prims=%d meshes=%d sceneNodes=%d probe=%d -> %s (%d bytes)" % (
            name, rs.options.width, rs.options.height, s.numPrimitives, s.numMeshes, s.numBvhNodes,
            s.sky.probeValid, out, os.path.getsize(out)))


if __name__ == "__main__":
    main(sys.argv[1:] or list(SCENES))
