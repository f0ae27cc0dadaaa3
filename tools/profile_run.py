#!/usr/bin/env python
"""Minimal driver for ncu captures: one warm-up launch, then one k_wavefront launch of `spp` samples
per pixel on scene/size given on the command line (default cornell 1024x1024, 8 spp)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinsel_b200 as tb  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 8
snap = tb.Snapshot(tb.scene_path(scene))
cam, opt = snap.camera, snap.options
opt.width, opt.height = w, h
r = tb.Renderer(snap.scene)
r.Init(w, h)
r.render_device(cam, opt, 1)
r.render_device(cam, opt, spp)
print("%s %dx%d spp=%d: %.3f ms (%.1f Msamples/s)" % (scene, w, h, spp, r.stats().gpuMs, w * h * spp / r.stats().gpuMs / 1e3))
r.close()
