"""GPU parity tests proper: the CUDA path through the C ABI vs the reference's own arithmetic
(oracle/_ref detmath flavour = tinsel's src/render.cpp compiled with the deterministic libm the
kernels use), per sample and bit-exact, plus image-level checks through Renderer.Render."""
import os

import numpy as np
import pytest

import tinsel_b200 as tb
from tinsel_b200 import abi
import refdrv

pytestmark = pytest.mark.gpu

# scene -> (width, height) used for the per-sample tests (None keeps the .tin's size)
SCENES = {
    "cornell": (256, 256),
    "veach": (192, 128),
    "glass": (128, 128),
    "meshlight": (128, 128),
    "motionblur": (128, 128),
    "gloss": (128, 128),
    "emitter": (128, 128),
    "furnace": (96, 96),
    "conservation": (128, 64),
    "ajax": (160, 160),
    "env": (160, 160),
    "many": (120, 80),     # 22 primitives: reference-order BVH walk instead of the flat scene program
    "mini": (96, 64),      # glass sphere (specular transmission, Beer-Lambert medium) under a gradient sky
    "table": (160, 100),   # 15 primitives, 5 meshes (prisms, cubes, quads) on a table: glass, metal, many instances
    "simple": (96, 48),    # two primitives
    "envmini": (128, 96),  # HDR-probe lighting only (synthetic 128x64 probe): ProbeSample / ProbePdf / ProbeEval
}


def _available(name):
    return os.path.exists(tb.scene_path(name))


def _setup(name, pipeline, size=None, flavour="detmath"):
    if not _available(name):
        pytest.skip("snapshot scenes/%s.tsnap not present" % name)
    if not refdrv.have_ref(flavour):
        pytest.skip("oracle/_ref not built")
    os.environ["TINSEL_B200_PIPELINE"] = pipeline
    snap = tb.Snapshot(tb.scene_path(name))
    cam, opt = snap.camera, snap.options
    w, h = size or SCENES[name]
    opt.width, opt.height = w, h
    ref = refdrv.RefScene.from_snapshot(tb.scene_path(name), flavour=flavour)
    ref.set_size(w, h)
    r = tb.Renderer(snap.scene)
    r.Init(w, h)
    return snap, cam, opt, ref, r


def test_sample_seed_matches_host():
    # host copy (snapshot.cpp) vs the values the device copy is expected to produce are compared
    # implicitly by every per-sample test; here pin the host function itself
    assert tb.sample_seed(0, 0) != tb.sample_seed(1, 0)
    assert tb.sample_seed(5, 7) == tb.sample_seed(5, 7)


@pytest.mark.parametrize("pipeline", ["mega", "wavefront"])
@pytest.mark.parametrize("name", list(SCENES))
def test_per_sample_radiance_bit_exact(name, pipeline):
    snap, cam, opt, ref, r = _setup(name, pipeline)
    for frame in (0, 5):
        rad, ras = r.trace_frame(cam, opt, frame)
        rrad, rras = ref.trace_frame(frame, nthreads=8)
        assert np.array_equal(ras, rras), "raster positions differ"
        same = (rad.view(np.uint32) == rrad.view(np.uint32)).all(-1) | (np.isnan(rad).any(-1) & np.isnan(rrad).any(-1))
        bad = int((~same).sum())
        assert bad == 0, "%s/%s frame %d: %d of %d samples differ (max abs %g)" % (
            name, pipeline, frame, bad, same.size, float(np.nanmax(np.abs(rad - rrad))))
    r.close()
    ref.close()
    snap.close()


@pytest.mark.parametrize("pipeline", ["mega", "wavefront"])
@pytest.mark.parametrize("name", ["cornell", "veach", "glass", "ajax", "env", "envmini"])
def test_render_matches_seeded_oracle(name, pipeline):
    """Renderer.Render x spp vs oracle B at matched spp and seeds: rel-L2 <= 1e-4 (BASELINE.json)."""
    snap, cam, opt, ref, r = _setup(name, pipeline)
    spp = 4
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    for _ in range(spp):
        r.Render(cam, opt, out)
    oracle = ref.render_seeded(0, spp, nthreads=8)
    num = np.linalg.norm((out - oracle).astype(np.float64))
    den = np.linalg.norm(oracle.astype(np.float64))
    assert num / den <= 1e-4, "rel L2 %g" % (num / den)
    # the filter-weight channel must agree to fp32 summation noise
    assert np.allclose(out[..., 3], oracle[..., 3], rtol=1e-5, atol=1e-6)
    assert r.stats().frames == spp
    r.close()
    ref.close()
    snap.close()


@pytest.mark.parametrize("name", ["cornell", "veach", "meshlight", "ajax"])
def test_normals_mode_bit_exact_vs_literal_reference(name):
    """eNormals (render.cpp:494-515) has no RNG and no libm: exact match against the literal reference."""
    snap, cam, opt, ref, r = _setup(name, "wavefront", flavour="literal")
    opt.mode = abi.MODE_NORMALS
    ref.set_mode(abi.MODE_NORMALS)
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    r.Render(cam, opt, out)
    expect = ref.render_literal(1)
    assert np.array_equal(out.view(np.uint32), expect.view(np.uint32)), \
        "%d pixels differ" % int((out != expect).any(-1).sum())
    r.close()
    ref.close()
    snap.close()


def test_batched_render_equals_repeated_render():
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=(128, 128))
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    for _ in range(3):
        r.Render(cam, opt, out)
    r2 = tb.Renderer(snap.scene)
    r2.Init(opt.width, opt.height)
    r2.render_device(cam, opt, 3)
    out2 = r2.read_accumulator()
    assert np.allclose(out, out2, rtol=2e-6, atol=1e-6)
    # row-sharded rendering (image-plane DP) sums to the same image
    r3 = tb.Renderer(snap.scene)
    r3.Init(opt.width, opt.height)
    r3.render_device(cam, opt, 3, 0, 50)
    r3.set_frame(0)
    r3.render_device(cam, opt, 3, 50, opt.height - 50)
    out3 = r3.read_accumulator()
    assert np.allclose(out, out3, rtol=2e-6, atol=1e-6)
    for x in (r, r2, r3):
        x.close()
    ref.close()
    snap.close()


def test_literal_reference_statistical_agreement():
    """Against the reference as shipped (glibc libm): no per-sample exactness is possible in
    principle (libm rounding), but at matched seeds the images must agree to ~1e-6."""
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=(128, 128), flavour="literal")
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    for _ in range(8):
        r.Render(cam, opt, out)
    oracle = ref.render_seeded(0, 8, nthreads=8)
    rel = np.linalg.norm((out - oracle).astype(np.float64)) / np.linalg.norm(oracle.astype(np.float64))
    assert rel <= 1e-4, rel
    r.close()
    ref.close()
    snap.close()


def test_interleaved_shards_sum_to_full_image():
    """Image-plane sharding (tb200_set_shard): the shard accumulators sum to the unsharded image."""
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=(100, 70))
    r.render_device(cam, opt, 3)
    full = r.read_accumulator()
    total = np.zeros_like(full)
    samples = 0
    for shard in range(3):
        rs = tb.Renderer(snap.scene)
        rs.Init(opt.width, opt.height)
        rs.set_shard(shard, 3)
        rs.render_device(cam, opt, 3)
        total += rs.read_accumulator()
        samples += rs.stats().samples
        rs.close()
    assert samples == 3 * opt.width * opt.height
    assert np.allclose(full, total, rtol=2e-6, atol=1e-6)
    r.close()
    ref.close()
    snap.close()


@pytest.mark.parametrize("size", [(1, 1), (3, 100), (37, 23), (8, 4), (129, 65)])
def test_ragged_image_sizes(size):
    """Image sizes that are not multiples of the 8x4 sample tiles (and the degenerate 1x1)."""
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=size)
    rad, ras = r.trace_frame(cam, opt, 2)
    rrad, rras = ref.trace_frame(2, nthreads=2)
    assert np.array_equal(ras, rras)
    assert (rad.view(np.uint32) == rrad.view(np.uint32)).all()
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    r.Render(cam, opt, out)
    r.Render(cam, opt, out)
    oracle = ref.render_seeded(0, 2, nthreads=1)
    assert np.allclose(out, oracle, rtol=1e-5, atol=1e-6)
    assert r.stats().samples == 2 * size[0] * size[1]
    r.close()
    ref.close()
    snap.close()


@pytest.mark.parametrize("pipeline", ["mega", "wavefront"])
@pytest.mark.parametrize("depth", [0, 1, 2, 7])
def test_max_depth_edge_cases(depth, pipeline):
    """maxDepth 0 (no trace at all), 1 (camera ray + NEE only), deeper than the default 4."""
    snap, cam, opt, ref, r = _setup("glass", pipeline, size=(64, 48))
    opt.maxDepth = depth
    ref.set_max_depth(depth)
    rad, _ = r.trace_frame(cam, opt, 1)
    rrad, _ = ref.trace_frame(1, nthreads=4)
    assert (rad.view(np.uint32) == rrad.view(np.uint32)).all()
    if depth == 0:
        assert not rad.any()
    r.close()
    ref.close()
    snap.close()


@pytest.mark.parametrize("cta", ["512", "768"])
@pytest.mark.parametrize("sched", ["hard", "free", "ring", "split", "treelet", "offload", "offload1", "offload100"])
def test_both_scheduling_modes_bit_exact(sched, cta):
    """Every scheduler variant x CTA size of the wavefront kernel (the per-scene heuristics of
    api.cu build_scene pick one of them) produces the same bits: hard phases, free running, free
    running with the split trace queue forced on for every scene that has a mesh, and the mesh-walk
    offload (shader CTAs + walker CTAs, wavefront_walk.cuh) forced on for every mesh, with the
    default number of walker CTAs, a single one, and as many as fit; "treelet" = free running with the top
    of the biggest mesh's BVH staged in shared memory for the inline walk.  "free" at 512 threads runs the
    lane-owned slot layout with bit-set queues on every scene without a split queue, "ring" the ring queues."""
    offload = sched.startswith("offload")
    os.environ["TINSEL_B200_SCHED"] = "free" if sched in ("split", "treelet", "ring") or offload else sched
    os.environ["TINSEL_B200_QUEUES"] = "ring" if sched == "ring" else "lanes"
    os.environ["TINSEL_B200_TREELET"] = "1" if sched == "treelet" else "0"   # top of the biggest mesh's BVH TMA-staged for the inline walk
    os.environ["TINSEL_B200_SPLIT"] = "1" if sched == "split" else "0"
    os.environ["TINSEL_B200_OFFLOAD"] = "2" if offload else "0"
    if sched in ("offload1", "offload100"):
        os.environ["TINSEL_B200_WALKERS"] = sched[len("offload"):]
    os.environ["TINSEL_B200_CTA"] = cta
    try:
        for name in ("veach", "meshlight", "many", "envmini", "glass", "table", "ajax", "motionblur"):
            if not _available(name):
                continue
            snap, cam, opt, ref, r = _setup(name, "wavefront")
            for frame in (3, 4):   # two launches: the offload queues' counters and lap tags run on across launches
                rad, _ = r.trace_frame(cam, opt, frame)
                rrad, _ = ref.trace_frame(frame, nthreads=8)
                assert (rad.view(np.uint32) == rrad.view(np.uint32)).all(), (name, frame)
            r.close()
            ref.close()
            snap.close()
    finally:
        for k in ("TINSEL_B200_SCHED", "TINSEL_B200_SPLIT", "TINSEL_B200_CTA", "TINSEL_B200_OFFLOAD", "TINSEL_B200_WALKERS", "TINSEL_B200_TREELET",
                  "TINSEL_B200_QUEUES"):
            os.environ.pop(k, None)


def test_box_filter_and_clamp():
    """eFilterBox (render.cpp:405-423) and a tight radiance clamp (ClampLength, maths.h:1577-1589)."""
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=(64, 64))
    opt.filterType = abi.FILTER_BOX
    opt.clamp = 0.75
    o = ref.options
    # the reference driver has no setter for these: compare against the CPU restatement instead
    port = refdrv.PortScene.from_snapshot(tb.scene_path("cornell"))
    port.set_size(64, 64)
    port.options.filterType = abi.FILTER_BOX
    port.options.clamp = 0.75
    out = np.zeros((64, 64, 4), np.float32)
    for _ in range(3):
        r.Render(cam, opt, out)
    expect = port.render_seeded(0, 3, 1)
    assert np.allclose(out, expect, rtol=1e-5, atol=1e-6)
    assert float(np.linalg.norm(out[..., :3], axis=-1).max()) <= 0.75 * out[..., 3].max() * 1.001 + 1e-3
    r.close()
    ref.close()
    port.close()
    snap.close()


def test_rotated_camera_and_fov():
    snap, cam, opt, ref, r = _setup("cornell", "wavefront", size=(80, 60))
    q = np.array([0.05, 0.2, -0.03, 0.97], np.float32)
    q /= np.linalg.norm(q)
    port = refdrv.PortScene.from_snapshot(tb.scene_path("cornell"))
    port.set_size(80, 60)
    for c in (cam, port.camera):
        c.rotation[:] = [float(x) for x in q]
        c.position[:] = [0.2, 1.1, 3.5]
        c.fov = 0.9
    rad, ras = r.trace_frame(cam, opt, 0)
    prad, pras = port.trace_frame(0, 4)
    assert np.array_equal(ras, pras)
    assert (rad.view(np.uint32) == prad.view(np.uint32)).all()
    r.close()
    ref.close()
    port.close()
    snap.close()


@pytest.mark.parametrize("case", [
    ("cornell", (1024, 1024), 1, 0.0),    # bench size: 64 bands of 256 KiB
    ("cornell", (1021, 515), 1, 0.0),     # ragged: tile padding retires through the band counters too
    ("veach", (640, 2048), 1, 0.0),       # hard-phase scheduling, tall image (64 bands)
    ("cornell", (512, 512), 3, 0.0),      # rank 1 of 3: interleaved tile rows
    ("cornell", (300, 900), 1, 3.5),      # wide box filter: rows stay open 4-5 rows behind the frontier
    ("cornell", (5, 3), 1, 0.0),          # fewer rows than one tile
])
def test_streamed_readback_delivers_final_rows(case):
    """tb200_render copies finished pixel rows to the host while the kernel is still running
    (api.cu render_streamed / wavefront2.cuh wf2_band_report).  Whatever the timing, the host
    buffer must equal the device accumulator bit for bit after every call, and must equal what
    the plain launch-then-copy path delivers up to the accumulation order."""
    name, size, shards, boxw = case
    snap, cam, opt, ref, r = _setup(name, "wavefront", size=size)
    if boxw > 0.0:
        opt.filterType = abi.FILTER_BOX
        opt.filterWidth = boxw
    r.set_shard(shards - 1 if shards > 1 else 0, shards)
    out = np.full((opt.height, opt.width, 4), -1.0, np.float32)
    for frame in range(4):
        r.Render(cam, opt, out)
        dev = r.read_accumulator()
        assert (out.view(np.uint32) == dev.view(np.uint32)).all(), "frame %d: stale rows in the host buffer" % frame
    os.environ["TINSEL_B200_READBACK"] = "plain"
    try:
        rp = tb.Renderer(snap.scene)
    finally:
        del os.environ["TINSEL_B200_READBACK"]
    rp.Init(opt.width, opt.height)
    rp.set_shard(shards - 1 if shards > 1 else 0, shards)
    outp = np.zeros_like(out)
    for frame in range(4):
        rp.Render(cam, opt, outp)
    assert np.allclose(out, outp, rtol=2e-5, atol=1e-6)
    assert r.stats().samples == rp.stats().samples
    rp.close()
    r.close()
    ref.close()
    snap.close()


def test_lane_queue_layout_on_small_frames(monkeypatch):
    """Launches of fewer than six samples per slot run the ring queues by default (kernels.cu:
    launch_wavefront2), so the small-frame feature tests above never see the lane-owned layout: force it
    (TINSEL_B200_QUEUES=lanes) and repeat the ones that exercise ragged tiles, tile padding, the band counters
    of the streamed read-back, interleaved shards and batched frames."""
    monkeypatch.setenv("TINSEL_B200_QUEUES", "lanes")
    for size in [(1, 1), (37, 23), (129, 65)]:
        test_ragged_image_sizes(size)
    test_batched_render_equals_repeated_render()
    test_interleaved_shards_sum_to_full_image()
    for case in [("cornell", (1021, 515), 1, 0.0), ("cornell", (512, 512), 3, 0.0), ("cornell", (300, 900), 1, 3.5), ("cornell", (5, 3), 1, 0.0)]:
        test_streamed_readback_delivers_final_rows(case)


# BASELINE.json configs at their full image sizes (C2 cornell 1024^2, C3 ajax 1024^2, C4 veach
# 1920x1080 with clamp 4, C5 env 2048^2): one frame of every config traced per sample on the GPU
# and by the reference's PathTrace on all host cores -- bit-exact -- and the same frame through
# Render() against the seeded oracle image.
@pytest.mark.parametrize("name,size", [("cornell", (1024, 1024)), ("ajax", (1024, 1024)), ("veach", (1920, 1080)),
                                       ("env", (2048, 2048))])
def test_baseline_configs_full_size(name, size):
    snap, cam, opt, ref, r = _setup(name, "wavefront", size=size)
    threads = min(os.cpu_count() or 8, 32)   # the box runs under a 16-CPU cgroup quota: more threads only thrash
    frame = 3
    rad, ras = r.trace_frame(cam, opt, frame)
    rrad, rras = ref.trace_frame(frame, nthreads=threads)
    assert np.array_equal(ras, rras)
    same = (rad.view(np.uint32) == rrad.view(np.uint32)).all(-1) | (np.isnan(rad).any(-1) & np.isnan(rrad).any(-1))
    assert bool(same.all()), "%d of %d samples differ" % (int((~same).sum()), same.size)
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    r.Render(cam, opt, out)
    r.Render(cam, opt, out)
    oracle = ref.render_seeded(0, 2, nthreads=threads)
    num = np.linalg.norm((out - oracle).astype(np.float64))
    den = np.linalg.norm(oracle.astype(np.float64))
    assert num / den <= 1e-4, "rel L2 %g" % (num / den)
    r.close()
    ref.close()
    snap.close()


# north_star's tolerance is against src/render.cpp AS SHIPPED (glibc libm, `flavour="literal"`:
# render.cpp:230-388 compiled unmodified), not against the deterministic-libm pin used above.  Per-sample
# exactness is impossible there in principle (glibc's sinf/cosf/expf/acosf/atan2f differ from the
# correctly rounded results in the last bit for 0.06-16 % of arguments), so: the image at matched
# seeds and spp must agree to rel-L2 <= 1e-4 on every BASELINE.json GPU configuration at its full
# size, and the per-sample differences are counted and printed (last-bit differences vs. flipped paths).
# A last-bit difference that flips a branch or a probe texel replaces one sample, so the image error
# falls as 1/sqrt(spp): env (4000x2000 HDR probe, nearest-texel lookups through acosf/atan2f: ~2e-5 of
# the samples land on another texel) measures 1.6e-4 on rgb/w at 8 spp, and is tested at 32 spp -- its
# BASELINE configuration runs 2048 spp, where the same error is ~1e-5.
LITERAL_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "literal_parity.txt")


# glass (not a BASELINE configuration; specular transmission): ONE path that takes another branch carries a
# caustic's unbounded radiance, 10^3 x the mean -- 1 such sample in 262 144 puts the 8-spp image at 1.6e-3.
# Its bound is therefore on the pixels: all but a 1e-4 fraction of them agree to 1e-3, and the image to 5e-3.
@pytest.mark.parametrize("name,size,spp,tol", [("cornell", (1024, 1024), 8, 1e-4), ("ajax", (1024, 1024), 8, 1e-4),
                                               ("veach", (1920, 1080), 8, 1e-4), ("env", (2048, 2048), 32, 1e-4),
                                               ("glass", (512, 512), 8, 5e-3)])
def test_literal_reference_rel_l2_at_baseline_sizes(name, size, spp, tol):
    snap, cam, opt, ref, r = _setup(name, "wavefront", size=size, flavour="literal")
    threads = min(os.cpu_count() or 8, 32)
    out = np.zeros((opt.height, opt.width, 4), np.float32)
    r.render_n(cam, opt, spp, out)
    oracle = ref.render_pool(0, spp, threads)
    num = np.linalg.norm((out - oracle).astype(np.float64))
    den = np.linalg.norm(oracle.astype(np.float64))
    # normalised image (rgb / w), the quantity main.cpp displays
    wgt = np.maximum(oracle[..., 3:], 1e-20)
    img_o, img_g = oracle[..., :3] / wgt, out[..., :3] / np.maximum(out[..., 3:], 1e-20)
    rel_img = np.linalg.norm((img_g - img_o).astype(np.float64)) / np.linalg.norm(img_o.astype(np.float64))
    # per-sample bookkeeping on one frame
    rad, _ = r.trace_frame(cam, opt, 1)
    rrad, _ = ref.trace_frame(1, nthreads=threads)
    same = (rad.view(np.uint32) == rrad.view(np.uint32)).all(-1)
    mag = np.maximum(np.abs(rrad).max(-1), 1e-6)
    flipped = (np.abs(rad - rrad).max(-1) / mag) > 1e-3
    msg = "%-8s %4dx%-4d %d spp: rel-L2 sums %.3e, rel-L2 rgb/w %.3e; frame 1: %d of %d samples differ in the last bits, %d by more than 1e-3 (flipped paths)" % (
        name, size[0], size[1], spp, num / den, rel_img, int((~same).sum()), same.size, int(flipped.sum()))
    print(msg)
    try:
        os.makedirs(os.path.dirname(LITERAL_REPORT), exist_ok=True)
        with open(LITERAL_REPORT, "a") as f:
            f.write(msg + "\n")
    except OSError:
        pass
    assert num / den <= tol, msg
    assert rel_img <= tol, msg
    if tol > 1e-4:
        # the looser image bound comes with a bound on how many pixels may be off at all
        pix = np.abs(img_g - img_o).max(-1) / np.maximum(np.abs(img_o).max(-1), 1e-6)
        assert float((pix > 1e-2).mean()) <= 1e-3, msg
    r.close()
    ref.close()
    snap.close()
