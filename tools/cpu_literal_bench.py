#!/usr/bin/env python
"""the LITERAL reference -- tinsel's CpuRenderer exactly as it ships
(one sequential RNG stream, one thread: render.cpp:399,462-464) -- timed on this box's host CPU with
the reference's own flags (-O3 -DNDEBUG -ffast-math, makefile:4: oracle/_ref/libtinsel_ref_fast.so)
and with the strict flags the parity oracle uses (-O2, no fast-math).  Msamples/s = W*H*spp / wall s.
    python tools/cpu_literal_bench.py [scene w h spp]..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tinsel_b200 as tb  # noqa: E402
import refdrv  # noqa: E402

jobs = [a.split() for a in sys.argv[1:]] or [["cornell", "256", "256", "16"], ["ajax", "256", "256", "4"], ["veach", "480", "270", "2"]]
for scene, w, h, spp in jobs:
    if not os.path.exists(tb.scene_path(scene)):
        continue
    row = []
    for flavour in ("fast", "literal"):
        if not refdrv.have_ref(flavour):
            row.append("%s: not built" % flavour)
            continue
        rs = refdrv.RefScene.from_snapshot(tb.scene_path(scene), flavour)
        rs.set_size(int(w), int(h))
        t0 = time.perf_counter()
        rs.render_literal(int(spp))
        dt = time.perf_counter() - t0
        row.append("%s %.3f Msamples/s (%.2f s)" % ("-O3 -ffast-math" if flavour == "fast" else "-O2 strict", int(w) * int(h) * int(spp) / dt / 1e6, dt))
        rs.close()
    print("%s %sx%s spp %s, 1 thread: %s" % (scene, w, h, spp, "; ".join(row)))
