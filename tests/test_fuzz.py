"""Fuzz parity over random scenes (tests/fuzzscene.py): the whole Disney parameter space, moving and
scaled spheres, sphere lights with 1-2 samples, gradient skies.  CPU: the restatement against the
reference's own compiled PathTrace, per sample, bit for bit.  GPU: the CUDA path against the same."""
import os

import numpy as np
import pytest

import tinsel_b200 as tb
import refdrv
import fuzzscene

SEEDS = list(range(12))            # GPU part (validated on the B200)
CPU_SEEDS = list(range(48))        # restatement vs the reference's compiled code


def _same(a, b):
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.skipif(not refdrv.have_ref("detmath"), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", CPU_SEEDS)
def test_port_matches_reference_on_random_scenes(seed, tmp_path):
    path = str(tmp_path / "fuzz.tsnap")
    what = fuzzscene.make_scene(seed, path)
    ref = refdrv.RefScene.from_snapshot(path, "detmath")
    port = refdrv.PortScene.from_snapshot(path)
    for frame in (0, 3):
        a, ar = ref.trace_frame(frame, 4)
        b, br = port.trace_frame(frame, 4)
        assert np.array_equal(ar, br), what
        assert _same(a, b), what
    assert float(np.nan_to_num(a).sum()) > 0.0, what      # the scene is lit
    ref.close()
    port.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_gpu_matches_oracle_on_random_scenes(seed, tmp_path):
    os.environ.pop("TINSEL_B200_PIPELINE", None)
    path = str(tmp_path / "fuzz.tsnap")
    what = fuzzscene.make_scene(seed, path, width=96, height=64)
    port = refdrv.PortScene.from_snapshot(path)
    snap = tb.Snapshot(path)
    cam, opt = snap.camera, snap.options
    r = tb.Renderer(snap.scene)
    r.Init(opt.width, opt.height)
    for frame in (0, 3):
        rad, ras = r.trace_frame(cam, opt, frame)
        prad, pras = port.trace_frame(frame, 8)
        assert np.array_equal(ras, pras), what
        bad = int((~((rad.view(np.uint32) == prad.view(np.uint32)) | (np.isnan(rad) & np.isnan(prad))).all(-1)).sum())
        assert bad == 0, "%s: %d of %d samples differ" % (what, bad, rad.shape[0] * rad.shape[1])
    r.close()
    port.close()
    snap.close()
