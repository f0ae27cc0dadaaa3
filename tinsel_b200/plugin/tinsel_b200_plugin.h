// tinsel_b200_plugin.h -- what a tinsel host may include next to src/render.h.
//
// The drop-in itself needs nothing from here: CreateGpuWavefrontRenderer is already declared in
// src/render.h:78.  The two functions below are OPTIONAL fast paths for src/main.cpp's frame loop
// (see INTEGRATION.md): they only act on renderers made by CreateGpuWavefrontRenderer and return
// false for any other, so the host keeps its existing code as the fallback.
#pragma once

#include "render.h"   // the reference's src/render.h

// `n` x Renderer::Render with one read-back (src/main.cpp:242-251 calls Render 16 times per frame).
bool TinselB200RenderN(Renderer* r, const Camera& camera, const Options& options, int n, Color* output);

// The display loop of src/main.cpp:258-271 on the device, from the renderer's own running sums:
//   filtered[i] = LinearToSrgb(ToneMap(pixels[i] * (options.exposure / pixels[i].w), options.limit))
// and/or the 8-bit RGB buffer WritePng would build from it (src/png.cpp:329-343, same dither).
// Either pointer may be null.
bool TinselB200Finish(Renderer* r, const Options& options, Color* filtered, unsigned char* rgb8);

// NonLocalMeansFilter(g_filtered, g_exposed, w, h, falloff, radius) of src/main.cpp:273-277 on the
// device, applied to the image the last TinselB200Finish produced (src/nlm.cpp:36-73, same sums).
bool TinselB200Nlm(Renderer* r, float falloff, int radius, Color* out);
