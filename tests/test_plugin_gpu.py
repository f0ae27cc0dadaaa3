"""The reference-facing C++ adapter (tinsel_b200/plugin): a reference `Scene` (rebuilt from a
snapshot by oracle/_ref) goes through the Itanium-mangled CreateGpuWavefrontRenderer(const Scene*),
Renderer::Init and Renderer::Render exactly as tinsel's main.cpp would call them."""
import ctypes as C
import os

import numpy as np
import pytest

import tinsel_b200 as tb
import refdrv

PLUGIN = os.path.join(tb.ROOT, "tinsel_b200", "plugin", "libtinsel_b200_plugin.so")


def test_plugin_exports_reference_factory_symbol():
    if not os.path.exists(PLUGIN):
        pytest.skip("plugin not built (needs the reference headers; build container only)")
    lib = C.CDLL(PLUGIN)
    assert hasattr(lib, "_Z26CreateGpuWavefrontRendererPK5Scene")   # Renderer* CreateGpuWavefrontRenderer(const Scene*)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell", "veach"])
def test_plugin_render_matches_seeded_oracle(name):
    if not os.path.exists(PLUGIN) or not refdrv.have_ref("detmath"):
        pytest.skip("plugin or oracle/_ref not built")
    os.environ.pop("TINSEL_B200_PIPELINE", None)
    ref = refdrv.RefScene.from_snapshot(tb.scene_path(name), "detmath")
    ref.set_size(96, 80)
    lib = C.CDLL(PLUGIN)
    lib.tb200_plugin_render.restype = C.c_int
    lib.tb200_plugin_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    out = np.zeros((80, 96, 4), np.float32)
    rc = lib.tb200_plugin_render(ref.lib.ref_native_scene(ref.h), ref.lib.ref_native_camera(ref.h),
                                 ref.lib.ref_native_options(ref.h), 4, out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    oracle = ref.render_seeded(0, 4, 4)
    rel = np.linalg.norm((out - oracle).astype(np.float64)) / np.linalg.norm(oracle.astype(np.float64))
    assert rel <= 1e-4, rel
    ref.close()
