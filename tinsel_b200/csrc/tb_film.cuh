// tb_film.cuh -- camera ray generation and filtered framebuffer accumulation.
#pragma once

#include "tb_shade.cuh"

// CameraSampler (util.h:45-83) after the host has built the two matrices (api.cu, camera_setup).
struct DCamera {
    float r2w[16];      // rasterToWorld, column major (cols[c][r] = r2w[c*4+r])
    V3 origin;          // cameraToWorld.GetCol(3)
    float shutterStart, shutterEnd;
};

struct DFilm {
    int width, height;
    int filterType;
    float filterWidth, filterFalloff, filterOffset;
    float clamp;
    int maxDepth;
    // arguments below this give expf(arg) < offset by a wide margin, i.e. Gaussian() == +0 exactly
    // (host: log(offset*(1-1e-3)); -inf disables the shortcut)
    float filterArgZero;
    // Owner-computes row slabs (multi-GPU): splats land only in pixel rows [rowLo, rowHi).  A device
    // that owns a slab also traces the samples of the rows within the filter's reach outside it, so
    // its rows receive exactly the contributions the whole image would give them.  Default: all rows.
    int rowLo, rowHi;
};

// GenerateRay, util.h:73-79 with TransformPoint(Mat44, Vec3(x,y,0)), maths.h:923-930
TB_DEV void generate_ray(const DCamera& cam, float rx, float ry, V3& origin, V3& dir)
{
    const float* m = cam.r2w;
    V3 p;
    p.x = m[0] * rx + m[4] * ry + m[8] * 0.0f + m[12];
    p.y = m[1] * rx + m[5] * ry + m[9] * 0.0f + m[13];
    p.z = m[2] * rx + m[6] * ry + m[10] * 0.0f + m[14];
    origin = cam.origin;
    dir = normalize(p - origin);
}

// Filter::Gaussian, render.h:29-32 (offset is taken as handed in, never recomputed)
TB_DEV float filter_gaussian(const DFilm& f, float x)
{
    const float arg = -f.filterFalloff * x * x;
    if (arg < f.filterArgZero) return 0.0f;   // Max(0, expf(arg) - offset) with expf(arg) < offset
    return tb_max(0.0f, tb_expf(arg) - f.filterOffset);
}

// one 16-byte vector REDUCTION per touched pixel: red.global.add.v4.f32 (REDG.E.ADD.F32x4).  CUDA's
// atomicAdd(float4*) compiles to the value-returning form (ATOMG), which makes the warp wait a round trip
// to L2 for a result nobody reads; the reduction is fire-and-forget.  Same adder, same rounding.  (The L2
// vector adder flushes denormal sums to zero -- ".FTZ" in the SASS mnemonic -- where the CPU's `+=` keeps
// them: the one place this renderer is not IEEE-exact; irrelevant at the 1e-4 image tolerance, and
// per-sample radiance never passes through it.)
TB_DEV void accum_add(float4* accum, int idx, float r, float g, float b, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(accum + idx), "f"(r), "f"(g), "f"(b), "f"(w) : "memory");
}

// CpuRenderer::AddSample, render.cpp:401-445.  Pixels whose weight is exactly zero are skipped
// when the sample is finite (adding +0 changes nothing); a non-finite sample takes the literal
// path so that it poisons the same pixels as in the reference.
TB_DEV void add_sample(const DFilm& f, float4* accum, float rasterX, float rasterY, V3 sample)
{
    const int startX = tb_max(0, int(rasterX - f.filterWidth));
    const int startY = tb_max(f.rowLo, int(rasterY - f.filterWidth));
    const int endX = tb_min(int(rasterX + f.filterWidth), f.width - 1);
    const int endY = tb_min(int(rasterY + f.filterWidth), f.rowHi - 1);

    // ClampLength, maths.h:1577-1589
    V3 c = sample;
    const float l = length(sample);
    if (l > f.clamp) c = sample * (f.clamp / l);

    if (f.filterType == TB200_FILTER_BOX) {
        for (int x = startX; x <= endX; ++x)
            for (int y = startY; y <= endY; ++y) accum_add(accum, y * f.width + x, c.x, c.y, c.z, 1.0f);
        return;
    }
    const bool finite = isfinite(c.x) && isfinite(c.y) && isfinite(c.z);
    for (int x = startX; x <= endX; ++x) {
        const float gx = filter_gaussian(f, x - rasterX);
        if (gx == 0.0f && finite) continue;
        for (int y = startY; y <= endY; ++y) {
            const float w = gx * filter_gaussian(f, y - rasterY);
            if (w == 0.0f && finite) continue;
            accum_add(accum, y * f.width + x, c.x * w, c.y * w, c.z * w, w);
        }
    }
}
