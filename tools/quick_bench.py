#!/usr/bin/env python
"""Quick device-side timing of the two pipelines (development aid; bench.py is the contract)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinsel_b200 as tb  # noqa: E402


def run(name, w, h, spp, pipeline):
    os.environ["TINSEL_B200_PIPELINE"] = pipeline
    snap = tb.Snapshot(tb.scene_path(name))
    cam, opt = snap.camera, snap.options
    opt.width, opt.height = w, h
    r = tb.Renderer(snap.scene)
    r.Init(w, h)
    r.render_device(cam, opt, 2)
    t0 = time.time()
    r.render_device(cam, opt, spp)
    wall = time.time() - t0
    ms = r.stats().gpuMs
    print("%-8s %-9s %dx%d spp=%d  gpu %.2f ms  wall %.2f ms  %.1f Msamples/s" % (
        name, pipeline, w, h, spp, ms, wall * 1e3, w * h * spp / ms / 1e3), flush=True)
    out = np.zeros((h, w, 4), np.float32)
    t0 = time.time()
    n = 8
    for _ in range(n):
        r.Render(cam, opt, out)
    wall = (time.time() - t0) / n
    print("         Render() e2e %.2f ms/spp -> %.1f Msamples/s" % (wall * 1e3, w * h / wall / 1e6), flush=True)
    r.close()
    snap.close()


if __name__ == "__main__":
    scenes = sys.argv[1:] or ["cornell"]
    for name in scenes:
        if not os.path.exists(tb.scene_path(name)):
            continue
        for pipe in ("mega", "wavefront"):
            run(name, 1024, 1024, 16, pipe)
