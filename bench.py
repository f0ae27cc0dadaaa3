#!/usr/bin/env python
"""bench.py -- headline benchmark of the tinsel_b200 path tracer (see BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arm

Metric: Msamples/s (camera paths incl. all bounces and shadow rays per second), whole job.
Workload: BASELINE.json configs[1] = data/cornell.tin, 1024x1024, Gaussian filter, maxDepth 4,
256 spp in total at the default K=8 steps; a "step" is one pass of the hot path adding
SPP_PER_STEP samples to every pixel of the image.  Image rows are sharded over the N ranks
(interleaved 4-row tile rows, fixed total work => "strong" scaling) and the accumulated radiance
is summed onto rank 0 with one NCCL reduce at the end of the timed region.

`value`   device-resident: scene and accumulator in HBM, timed with CUDA events on the launching
          stream (torch's current stream, bound with tb200_set_stream), max over ranks.
`e2e`     the same metric through the reference-facing call: tb200_render() = Renderer::Render
          (1 spp per call, HOST output buffer, device->host copy of the full accumulator inside
          every call), timed on the host clock.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = "cornell"               # --scene: other BASELINE.json configs (ajax 1024^2, veach 1920x1080, env 2048^2), not the default
WIDTH = HEIGHT = 1024
SPP_PER_STEP = 32
# the other BASELINE.json configurations at their own sizes (--scene): size and spp per step
OTHER_WORKLOADS = {"ajax": (1024, 1024, 16), "veach": (1920, 1080, 16), "env": (2048, 2048, 16)}
E2E_CALLS_PER_STEP = 8          # Render() calls (1 spp each) per e2e step

# Algorithmic bytes per camera sample for this workload, counted on the REFERENCE traversal order by
# the instrumented oracle (tools/count_bytes.py; formula of SURVEY.md 8d, restated in DESIGN.md):
#   B = 64*V_int + 48*T_tri + 136*T_prim + 48*H_mesh + 128*H + B_nee + B_fb
ALGO_BYTES_PER_SAMPLE = None    # filled from tools/algo_bytes.json
ALGO_JSON = os.path.join(ROOT, "tools", "algo_bytes.json")


def load_algo_bytes():
    with open(ALGO_JSON) as f:
        d = json.load(f)
    if SCENE != "cornell":
        o = d["other_scenes"][SCENE]
        return {"bytes_per_sample": o["bytes_per_sample"], "traversal_bytes_per_sample": o.get("traversal_bytes_per_sample"),
                "dram_traffic_bytes_per_launch": None}
    return d


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs (nvidia-smi equivalent via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            self.ok = True
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.error = str(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload_config(n_gpus):
    return {
        "workload": "tinsel data/%s.tin %dx%d, Gaussian filter, maxDepth from the scene file, %d spp per step" % (SCENE, WIDTH, HEIGHT, SPP_PER_STEP),
        "scene": "scenes/%s.tsnap (snapshot of the reference loader's Scene)" % SCENE,
        "width": WIDTH, "height": HEIGHT, "spp_per_step": SPP_PER_STEP,
        "sharding": "interleaved 4-row tile rows over %d rank(s); one NCCL sum-reduce of the accumulator at the end" % n_gpus,
        "l2": "flushed between timed steps (256 MiB device write outside the per-step events)",
    }


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref = tinsel src/render.cpp compiled
# unmodified, literal glibc flavour) with the per-sample-seeded multi-threaded driver
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(sample_w, sample_h, sample_spp, threads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refdrv
    import tinsel_b200 as tb
    if refdrv.have_ref("literal"):
        kind = "reference"
        sc = refdrv.RefScene.from_snapshot(tb.scene_path(SCENE), "literal")
        sc.set_size(sample_w, sample_h)
    else:
        kind = "port"
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        sc = refdrv.PortScene.from_snapshot(tb.scene_path(SCENE))
        sc.set_size(sample_w, sample_h)
    t0 = time.time()
    sc.render_seeded(0, sample_spp, threads)
    dt = time.time() - t0
    sc.close()
    return kind, sample_w * sample_h * sample_spp / dt / 1e6, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # each step: a bounded sample of the workload (full 1024x1024 image, 1 spp) on all host threads
    vals = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        kind, msps, dt = cpu_reference_run(WIDTH, HEIGHT, 1, cores)
        if i >= args.warmup:
            vals.append((msps, dt))
    total_dt = sum(d for _, d in vals)
    value = WIDTH * HEIGHT * len(vals) / total_dt / 1e6
    line = {
        "impl": "reference",
        "metric": "Msamples/sec (paths x spp / s)", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_dt / len(vals) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (scene snapshot, random paths)",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
                         "sample": "%dx%d image, 1 spp per step, per-sample-seeded driver over the reference's PathTrace" % (WIDTH, HEIGHT)},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import tinsel_b200 as tb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    os.environ.pop("TINSEL_B200_PIPELINE", None)
    snap = tb.Snapshot(tb.scene_path(SCENE))
    cam, opt = snap.camera, snap.options
    opt.width, opt.height = WIDTH, HEIGHT
    r = tb.Renderer(snap.scene, device=local)
    r.Init(WIDTH, HEIGHT)
    r.set_shard(rank, world)
    accum = torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float32, device="cuda")
    r.bind_accumulator(accum.data_ptr())
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    for _ in range(args.warmup):
        r.render_device(cam, opt, SPP_PER_STEP)
    if dist is not None:
        dist.reduce(accum.clone(), dst=0)   # warm the communicator
    accum.zero_()
    r.set_frame(0)
    barrier()

    clocks = ClockSampler(local)
    clocks.start()
    launches0 = r.stats().kernelLaunches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                       # evict L2 between timed steps (not timed)
        ev[i][0].record(stream)
        r.render_device(cam, opt, SPP_PER_STEP)     # synchronous: returns when the step's kernels are done
        ev[i][1].record(stream)
    ev[args.steps][0].record(stream)
    if dist is not None:
        dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
    ev[args.steps][1].record(stream)
    barrier()
    wall = time.time() - wall0
    clocks.stop_flag = True
    clocks.join(timeout=2)
    step_ms = [a.elapsed_time(b) for a, b in ev[:args.steps]]
    reduce_ms = ev[args.steps][0].elapsed_time(ev[args.steps][1])
    dev_ms = sum(step_ms) + reduce_ms
    launches = r.stats().kernelLaunches - launches0
    t = torch.tensor([dev_ms, float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms = float(tmax[0])
        launches = int(tsum[1])
    total_samples = WIDTH * HEIGHT * SPP_PER_STEP * args.steps
    value = total_samples / dev_ms / 1e3

    # sanity: the image is a plausible cornell box (filter weights sum to spp per interior pixel)
    if rank == 0:
        img = accum.cpu().numpy()
        # interior pixels collect, per spp, the integral of the filter over its support
        xs = (np.arange(4000) + 0.5) / 4000.0 * 2.0 - 1.0
        g = np.maximum(0.0, np.exp(-opt.filterFalloff * xs * xs) - opt.filterOffset)
        expect = float((g.sum() * (2.0 / 4000.0)) ** 2) * SPP_PER_STEP * args.steps
        wsum = float(img[4:-4, 4:-4, 3].mean())
        assert abs(wsum / expect - 1.0) < 0.01, (wsum, expect)
        assert np.isfinite(img).all() and float(img[..., :3].sum()) > 0.0

    # ---- e2e: Renderer::Render through the C ABI with a HOST output buffer ------------------------
    r.set_stream(None)
    r.bind_accumulator(None)
    r.Init(WIDTH, HEIGHT)
    r.set_shard(rank, world)
    host = np.zeros((HEIGHT, WIDTH, 4), np.float32)
    if world > 1:
        r.bind_accumulator(accum.data_ptr())
        accum.zero_()
        part = torch.zeros_like(accum)
        pinned = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32, pin_memory=True)
        host = pinned.numpy()

    def e2e_call():
        if world == 1:
            r.Render(cam, opt, host)                        # kernels + D2H of W*H*16 bytes
        else:
            # rank-local shard -> NCCL sum onto rank 0 -> rank 0 copies to the host buffer
            r.render_device(cam, opt, 1)
            part.copy_(accum)
            dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                pinned.copy_(part, non_blocking=True)   # into `host` (pinned), W*H*16 bytes
            torch.cuda.synchronize()

    for _ in range(3):
        e2e_call()
    barrier()
    t0 = time.time()
    calls = E2E_CALLS_PER_STEP * args.steps
    for _ in range(calls):
        e2e_call()
    barrier()
    e2e_dt = time.time() - t0
    if dist is not None:
        td = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e_dt = float(td[0])
    e2e_value = WIDTH * HEIGHT * calls / e2e_dt / 1e6

    if rank == 0:
        algo = load_algo_bytes()
        peak, peak_src = read_peaks()
        # dominant kernel: k_wavefront2 (one launch per step per rank); its share of the step is ~100 %
        per_launch_samples = WIDTH * HEIGHT * SPP_PER_STEP / world
        avg_launch_s = (sum(step_ms) / len(step_ms)) / 1e3
        achieved = algo["bytes_per_sample"] * per_launch_samples / avg_launch_s / 1e9
        cores = os.cpu_count() or 1
        kind, cpu_msps, cpu_dt = cpu_reference_run(WIDTH, HEIGHT, 2, cores)
        line = {
            "metric": "Msamples/sec (paths x spp / s)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (scene snapshot, random paths)",
            "config": workload_config(world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": E2E_CALLS_PER_STEP * (C.sizeof(tb.Camera) + C.sizeof(tb.Options)),
                    "d2h_bytes_per_step": E2E_CALLS_PER_STEP * WIDTH * HEIGHT * 16,
                    "calls_per_step": E2E_CALLS_PER_STEP, "ms_per_call": e2e_dt / calls * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": algo.get("dram_traffic_bytes_per_launch"),
                         "kernel": "k_wavefront2", "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": algo["bytes_per_sample"],
                         "traversal_bytes_per_sample": algo.get("traversal_bytes_per_sample"),
                         "note": "algorithmic bytes on the reference traversal order; the working set is SMEM/L2 resident so DRAM traffic is far below it"},
            "cpu_baseline": {"value": cpu_msps, "unit": "Msamples/s", "cores": cores, "kind": kind,
                             "sample": "%dx%d image, 2 spp (%.1f s), per-sample-seeded driver over the reference's PathTrace" % (WIDTH, HEIGHT, cpu_dt)},
            "wall_s_timed_region": wall, "reduce_ms": reduce_ms,
        }
        print(json.dumps(line), flush=True)
    r.close()
    snap.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="cornell", choices=["cornell"] + sorted(OTHER_WORKLOADS),
                    help="BASELINE.json configuration to run (default: configs[1], cornell 1024x1024)")
    args = ap.parse_args()
    if args.scene != "cornell":
        global SCENE, WIDTH, HEIGHT, SPP_PER_STEP
        SCENE = args.scene
        WIDTH, HEIGHT, SPP_PER_STEP = OTHER_WORKLOADS[args.scene]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
