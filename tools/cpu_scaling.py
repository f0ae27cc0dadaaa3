#!/usr/bin/env python
"""How the CPU arm (oracle/_ref, the reference's PathTrace under the work-stealing driver) scales with
the thread count on this host, next to the cgroup CPU quota: prints Msamples/s per thread count."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdrv  # noqa: E402
import tinsel_b200 as tb  # noqa: E402

for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
print("os.cpu_count", os.cpu_count(), "sched_getaffinity", len(os.sched_getaffinity(0)))
sc = refdrv.RefScene.from_snapshot(tb.scene_path(sys.argv[1] if len(sys.argv) > 1 else "cornell"), "literal")
sc.set_size(512, 512)
t0 = time.time()
while time.time() - t0 < 1.5:
    sc.render_pool(1000, 1, os.cpu_count())
for n in (1, 4, 8, 16, 32, 64, 96, 128, 192, 256):
    if n > 2 * (os.cpu_count() or 1):
        break
    spp = 1 if n < 8 else 4
    t0 = time.time()
    sc.render_pool(0, spp, n)
    dt = time.time() - t0
    print("threads %3d: %.2f Msamples/s" % (n, 512 * 512 * spp / dt / 1e6), flush=True)
