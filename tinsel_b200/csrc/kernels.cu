// kernels.cu -- sm_100a kernels of the tinsel_b200 path tracer.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true
//        -ftz=false -lineinfo  (see tb_math.cuh for why contraction is off)
#include "tb_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// Work decomposition.  A launch covers numFrames x (the 8x4 pixel tiles of this shard inside the
// row range).  Consecutive sample indices walk one tile (32 samples = one warp's worth of coherent
// camera rays), then the next tile of the row, then the shard's next tile row.
// ------------------------------------------------------------------------------------------------
void finalize_params(LaunchParams* p)
{
    if (p->numShards < 1) p->numShards = 1;
    p->tilesX = (p->film.width + 7) >> 3;
    const int tileRowsAll = (p->numRows + 3) >> 2;
    // tile rows t in [0, tileRowsAll) with t % numShards == shard
    p->tileRows = tileRowsAll > p->shard ? (tileRowsAll - p->shard + p->numShards - 1) / p->numShards : 0;
    p->samplesPerFrame = (unsigned long long)p->tilesX * (unsigned long long)p->tileRows * 32ull;
}

// (launches are split on the host so that idx and samplesPerFrame fit 31 bits: 32-bit division)
TB_DEV bool decode_sample(const LaunchParams& P, unsigned long long idx64, int& px, int& py, int& frame)
{
    const uint32_t idx = (uint32_t)idx64, spf = (uint32_t)P.samplesPerFrame;
    const uint32_t f = idx / spf;
    frame = P.frame0 + (int)f;
    const uint32_t local = idx - f * spf;
    const uint32_t tile = local >> 5, in = local & 31u;
    px = (int)((tile % (uint32_t)P.tilesX) * 8u + (in & 7u));
    const int tileRow = (int)(tile / (uint32_t)P.tilesX) * P.numShards + P.shard;
    const int ry = tileRow * 4 + (int)(in >> 3);
    py = P.firstRow + ry;
    return px < P.film.width && ry < P.numRows;
}

// ------------------------------------------------------------------------------------------------
// Shared per-sample prologue: seed Random per (pixel, frame), draw x,y,t in the oracle's order
// (render.cpp:476-482), generate the camera ray.
// ------------------------------------------------------------------------------------------------
TB_DEV void sample_begin(const LaunchParams& P, int px, int py, int frame, PathState& ps, float& rasterX, float& rasterY)
{
    Rng rng = rng_seed(tb_sample_seed((uint32_t)(py * P.film.width + px), (uint32_t)frame));
    float x = rng_float(rng);
    float y = rng_float(rng);
    const float t = rng_float(rng);
    const float time = tb_lerp(P.camera.shutterStart, P.camera.shutterEnd, t);
    x += px;
    y += py;
    V3 origin, dir;
    generate_ray(P.camera, x, y, origin, dir);
    path_init(ps, origin, dir, time, rng);
    rasterX = x;
    rasterY = y;
}

TB_DEV void sample_end(const LaunchParams& P, int px, int py, float rasterX, float rasterY, V3 radiance)
{
    if (P.outRadiance) {
        const size_t p = (size_t)py * P.film.width + px;
        P.outRadiance[p * 3 + 0] = radiance.x;
        P.outRadiance[p * 3 + 1] = radiance.y;
        P.outRadiance[p * 3 + 2] = radiance.z;
        P.outRaster[p * 2 + 0] = rasterX;
        P.outRaster[p * 2 + 1] = rasterY;
    } else {
        add_sample(P.film, P.accum, rasterX, rasterY, radiance);
    }
}

// ------------------------------------------------------------------------------------------------
// Validation pipeline: one thread = one path, PathTrace's loop (render.cpp:250-385) verbatim in
// terms of the stage functions.  Kept as the simplest possible arrangement of the same device
// functions the wavefront kernel uses, so that a parity failure can be bisected between "math"
// and "scheduling".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mega(LaunchParams P)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.samplesPerFrame * (unsigned long long)P.numFrames) return;
    int px, py, frame;
    if (!decode_sample(P, idx, px, py, frame)) return;

    PathState ps;
    float rasterX, rasterY;
    sample_begin(P, px, py, frame, ps, rasterX, rasterY);

    for (int bounce = 0; bounce < P.film.maxDepth; ++bounce) {
        const Hit h = trace_closest(P.scene, ps.o, ps.d, ps.time, true);
        if (h.prim < 0) {
            path_miss(P.scene, ps, bounce);
            break;
        }
        Surface sf;
        path_hit(P.scene, ps, h, bounce, sf);

        NeeCursor cur;
        nee_begin(cur);
        ShadowRay sr;
        while (nee_generate(P.scene, sf, ps.time, cur, ps.rng, sr)) {
            const Hit sh = trace_closest(P.scene, sr.o, sr.d, ps.time, false);
            nee_connect(P.scene, sf, sr, sh, cur);
        }
        if (!path_scatter(P.scene, ps, sf, cur.sum)) break;
    }
    sample_end(P, px, py, rasterX, rasterY, ps.L);
}

void launch_mega(const LaunchParams& p, cudaStream_t stream, unsigned long long* launchCount)
{
    const unsigned long long total = p.samplesPerFrame * (unsigned long long)p.numFrames;
    if (total == 0ull) return;
    const int block = 128;
    const unsigned long long grid = (total + block - 1) / block;
    k_mega<<<(unsigned)grid, block, 0, stream>>>(p);
    if (launchCount) ++*launchCount;
}

// ------------------------------------------------------------------------------------------------
// eNormals (render.cpp:494-515): ray through the integer raster position, time 1.0, overwrite.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_normals(LaunchParams P)
{
    const int W = P.film.width;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.numRows * W) return;
    const int py = P.firstRow + idx / W;
    const int px = idx % W;
    V3 origin, dir;
    generate_ray(P.camera, (float)px, (float)py, origin, dir);
    const Hit h = trace_closest(P.scene, origin, dir, 1.0f, true);
    float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (h.prim >= 0) {
        const V3 n = h.n * 0.5f + v3s(0.5f);
        out = make_float4(n.x, n.y, n.z, 1.0f);
    }
    P.accum[py * W + px] = out;
}

void launch_normals(const LaunchParams& p, cudaStream_t stream, unsigned long long* launchCount)
{
    const int total = p.numRows * p.film.width;
    if (total <= 0) return;
    k_normals<<<(total + 127) / 128, 128, 0, stream>>>(p);
    if (launchCount) ++*launchCount;
}

// ------------------------------------------------------------------------------------------------
// Wavefront pipeline: see wavefront2.cuh
// ------------------------------------------------------------------------------------------------
#include "wavefront2.cuh"
