#!/usr/bin/env python
"""Per-call cost of Renderer::Render on small frames (what one device of an N-GPU renderer sees):
kernel ms (CUDA events) and wall ms per call for cornell at 1024 x H, 1 spp per call."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import tinsel_b200 as tb  # noqa: E402

snap = tb.Snapshot(tb.scene_path("cornell"))
cam, opt = snap.camera, snap.options
for h in (1024, 512, 256, 132, 64):
    opt.width, opt.height = 1024, h
    r = tb.Renderer(snap.scene)
    r.Init(1024, h)
    host = np.zeros((h, 1024, 4), np.float32)
    r.pin_output(host)
    for _ in range(5):
        r.Render(cam, opt, host)
    n = 200
    ms = []
    t0 = time.time()
    for _ in range(n):
        r.Render(cam, opt, host)
        ms.append(r.stats().gpuMs)
    wall = (time.time() - t0) / n * 1e3
    ms.sort()
    print("1024x%-4d 1 spp per call: kernel %.3f ms (median), wall %.3f ms per Render(), %.0f Msamples/s" % (h, ms[n // 2], wall, 1024 * h / wall / 1e3))
    r.close()
