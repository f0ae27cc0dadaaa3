"""N>1 host logic on CPU (gloo, world_size 2): each rank renders ITS shard of the image with the
oracle (the product has no CPU path), the accumulators are sum-reduced onto rank 0 exactly as
bench.py does with NCCL, and the result must equal the unsharded oracle image."""
import os
import socket
import sys

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import sharding
import refdrv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_image():
    for h in (1, 3, 4, 5, 70, 256, 1080):
        for n in (1, 2, 3, 8):
            rows = sorted(r for s in range(n) for r in sharding.shard_rows(h, s, n))
            assert rows == list(range(h)), (h, n)
    assert sharding.shard_rows(10, 0, 2) == [0, 1, 2, 3, 8, 9]
    assert sharding.shard_rows(10, 1, 2) == [4, 5, 6, 7]
    assert sum(sharding.shard_sample_count(100, 70, s, 3, 2) for s in range(3)) == 100 * 70 * 2


def _worker(rank, world, port, w, h, spp, out_path):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = refdrv.PortScene.from_snapshot(tb.scene_path("cornell"))
    oracle.set_size(w, h)
    # render only this rank's rows: per-sample radiance -> filter splat of the owned samples
    acc = np.zeros((h, w, 4), np.float32)
    mine = set(sharding.shard_rows(h, rank, world))
    lib = oracle.lib
    for k in range(spp):
        rad, ras = oracle.trace_frame(k, 2)
        for j in sorted(mine):
            for i in range(w):
                _splat(acc, oracle.options, ras[j, i, 0], ras[j, i, 1], rad[j, i], lib)
    t = torch.from_numpy(acc)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()
    oracle.close()


def _splat(acc, o, rx, ry, c, lib):
    """AddSample (render.cpp:401-445) in numpy-free python, using the oracle's filter hook."""
    h, w = acc.shape[:2]
    fw = o.filterWidth
    for x in range(max(0, int(rx - fw)), min(int(rx + fw), w - 1) + 1):
        for y in range(max(0, int(ry - fw)), min(int(ry + fw), h - 1) + 1):
            wt = lib.oracle_filter_eval(o.filterType, fw, o.filterFalloff, o.filterOffset,
                                        np.float32(np.float32(x) - rx), np.float32(np.float32(y) - ry))
            acc[y, x, :3] += np.float32(wt) * c
            acc[y, x, 3] += np.float32(wt)


def test_two_rank_gloo_reduce_equals_full_image(tmp_path):
    import torch.multiprocessing as mp
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    w, h, spp = 24, 18, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, port, w, h, spp, out), nprocs=2, join=True)
    got = np.load(out)
    oracle = refdrv.PortScene.from_snapshot(tb.scene_path("cornell"))
    oracle.set_size(w, h)
    full = oracle.render_seeded(0, spp, 1)
    oracle.close()
    assert np.allclose(got, full, rtol=1e-5, atol=1e-6)
    assert got[..., 3].sum() > 0
