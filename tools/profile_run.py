#!/usr/bin/env python
"""Minimal timing / ncu driver: one warm-up launch, then `reps` launches of `spp` samples per pixel
on the scene/size given on the command line; prints the median and the best.
    python tools/profile_run.py [scene] [w] [h] [spp] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinsel_b200 as tb  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 8
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
snap = tb.Snapshot(tb.scene_path(scene))
cam, opt = snap.camera, snap.options
opt.width, opt.height = w, h
r = tb.Renderer(snap.scene)
r.Init(w, h)
r.render_device(cam, opt, 1)
ms = []
for _ in range(reps):
    r.render_device(cam, opt, spp)
    ms.append(r.stats().gpuMs)
ms.sort()
med = ms[len(ms) // 2]
print("%s %dx%d spp=%d: median %.3f ms (%.1f Msamples/s), best %.1f Msamples/s over %d" % (
    scene, w, h, spp, med, w * h * spp / med / 1e3, w * h * spp / ms[0] / 1e3, reps))
r.close()
