#!/usr/bin/env python
"""Times one displayed frame of tinsel's main loop (src/main.cpp:242-271: 16 x Render, then the
exposure/tone-map/sRGB loop over W*H pixels, then -- at the end -- WritePng's quantisation) three ways:
  (a) the reference's shape: 16 x tb200_render (16 read-backs) + the finish loop on the host (oracle port, 1 thread)
  (b) 16 x tb200_render + tb200_finish on the device
  (c) tb200_render_n(16) + tb200_finish(rgb8 only): 3 B/pixel read back
and the finish kernel alone.   python tools/finish_bench.py [scene] [w] [h]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tinsel_b200 as tb  # noqa: E402
import refdrv  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
snap = tb.Snapshot(tb.scene_path(scene))
cam, opt = snap.camera, snap.options
opt.width, opt.height = w, h
r = tb.Renderer(snap.scene)
r.Init(w, h)
out = np.zeros((h, w, 4), np.float32)
for _ in range(3):
    r.Render(cam, opt, out)
r.finish(1.0, 1.5)


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2] * 1e3


def frame_a():
    for _ in range(16):
        r.Render(cam, opt, out)
    refdrv.port_finish(out, 1.0)


def frame_b():
    for _ in range(16):
        r.Render(cam, opt, out)
    r.finish(1.0, 1.5, filtered=True, rgb8=False)


def frame_c():
    r.render_n(cam, opt, 16, out)
    r.finish(1.0, 1.5, filtered=False, rgb8=True)


res = {"a_16xRender+host_finish_ms": timed(frame_a, 3), "b_16xRender+device_finish_ms": timed(frame_b), "c_render_n16+device_rgb8_ms": timed(frame_c)}
ks = []
for _ in range(5):
    r.finish(1.0, 1.5, filtered=True, rgb8=True)
    ks.append(r.stats().gpuMs)
res["k_finish_ms"] = sorted(ks)[2]
res["k_finish_GBps"] = w * h * (16 + 16 + 3 + 8) / (res["k_finish_ms"] * 1e-3) / 1e9
t0 = time.perf_counter()
f = refdrv.port_finish(out, 1.0)
res["host_finish_1thread_ms"] = (time.perf_counter() - t0) * 1e3
r.finish(1.0, 1.5, filtered=False, rgb8=False)
ks = []
for _ in range(5):
    r.nlm(200.0, 1)
    ks.append(r.stats().gpuMs)
res["k_nlm_r1_ms"] = sorted(ks)[2]
res["nlm_r1_call_ms"] = timed(lambda: r.nlm(200.0, 1))
t0 = time.perf_counter()
refdrv.port_nlm(f, 200.0, 1)
res["host_nlm_r1_1thread_ms"] = (time.perf_counter() - t0) * 1e3
print(scene, w, h, res)
r.close()
