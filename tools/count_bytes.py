#!/usr/bin/env python
"""Counts the ALGORITHMIC bytes per camera sample of a workload on the REFERENCE traversal order,
with the instrumented oracle (oracle/libtinsel_oracle_count.so, `make -C oracle count`).

Formula:
  B = 64*V_int + 48*T_tri + 136*T_prim + 48*H_mesh + 128*H + B_nee + B_probe + B_miss + 32*P_fb
  V_int   interior-node visits, both BVH levels (two 32-byte child boxes each)
  T_tri   triangle tests (3 indices + 3 vertices)
  T_prim  scene-leaf primitive tests (first 136 bytes of Primitive: transforms + type + geometry)
  H_mesh  mesh hits (3 indices + 3 vertex normals)
  H       shaded surface hits (Material, 128 bytes)
  B_nee   136 bytes per light sample (+ cdf search and triangle fetch for mesh lights)
  B_probe probe-sample table reads; B_miss sky / probe-pdf reads on misses
  P_fb    framebuffer pixels touched by the filter footprint (16-byte read + write each)
Traversal-only bytes are the first three terms.  Writes tools/algo_bytes.json (read by bench.py).
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import refdrv  # noqa: E402
import tinsel_b200 as tb  # noqa: E402

NAMES = ["V_int", "T_tri", "T_prim", "H_mesh", "H", "B_nee", "B_probe", "B_miss", "P_fb", "rays", "samples"]


def count(scene, w, h, spp, threads=8):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "count"])
    refdrv.PORT_PATH = os.path.join(ROOT, "oracle", "libtinsel_oracle_count.so")
    refdrv._port = None
    lib = refdrv.load_port()
    lib.oracle_counters.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    port = refdrv.PortScene.from_snapshot(tb.scene_path(scene))
    port.set_size(w, h)
    buf = (C.c_ulonglong * 16)()
    lib.oracle_counters(buf, 1)
    port.render_seeded(0, spp, threads)
    lib.oracle_counters(buf, 1)
    c = dict(zip(NAMES, [int(x) for x in buf[:len(NAMES)]]))
    n = float(c["samples"])
    per = {k: c[k] / n for k in NAMES}
    trav = 64 * per["V_int"] + 48 * per["T_tri"] + 136 * per["T_prim"]
    total = trav + 48 * per["H_mesh"] + 128 * per["H"] + per["B_nee"] + per["B_probe"] + per["B_miss"] + 32 * per["P_fb"]
    return {"scene": scene, "width": w, "height": h, "spp_counted": spp, "per_sample": per,
            "traversal_bytes_per_sample": trav, "bytes_per_sample": total}


if __name__ == "__main__":
    scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    spp = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 2
    res = count(scene, w, h, spp)
    print(json.dumps(res, indent=1))
    if "--write" in sys.argv:
        # merge into tools/algo_bytes.json (bench.py reads it); the ncu-derived constants of the entry stay
        out = os.path.join(ROOT, "tools", "algo_bytes.json")
        doc = json.load(open(out)) if os.path.exists(out) else {"scenes": {}}
        entry = doc["scenes"].setdefault(scene, {})
        entry.update({"counted_at": "%dx%d, %d spp" % (w, h, spp), "per_sample": res["per_sample"],
                      "traversal_bytes_per_sample": res["traversal_bytes_per_sample"], "bytes_per_sample": res["bytes_per_sample"]})
        json.dump(doc, open(out, "w"), indent=1)
        print("wrote", out)
