// tb_kernels.cuh -- kernel launch interface shared by api.cu (host side) and kernels.cu.
#pragma once

#include "tb_film.cuh"

// Mesh-walk offload of the wavefront kernel (wavefront_walk.cuh): the queues between shader CTAs and
// walker CTAs, in global memory.  numWalkers == 0: the launch does not use it.
// path slots per CTA of the wavefront kernel, at most (the layouts of kernels.cu use 1024 and 2048): the host
// sizes the cold slot state and the offload queues with it
#define TB_WF2_SLOTS 2048
#define TB_WF2_MAX_CTAS_PER_SM 2   // the launch never places more CTAs per SM (the host sizes the cold state with it)
struct WalkParams {
    uint4* reqRing;                // (1 << reqLog2) cells x 3 chunks of 16 bytes
    unsigned int reqLog2;
    unsigned int* reqTail;
    unsigned int* reqHead;
    uint4* ansRing;                // [shader CTA][2][1024] cells x 3 chunks
    unsigned int* ansTail;         // [shader CTA][2]
    unsigned int* shadersDone;     // shader CTAs that have exited (this launch)
    unsigned int* walkersDone;
    unsigned int* abortFlag;       // set by a watchdog: everybody leaves, the host reports an error
    int numShaders, numWalkers;
    int treeletMesh;               // the mesh whose top-of-tree the walkers stage in shared memory (pairs in BFS order)
    int treeletPairs;              // its number of pairs
};

struct LaunchParams {
    DScene scene;
    DCamera camera;
    DFilm film;
    float4* accum;        // width*height running sums
    unsigned long long* sampleCounter;   // per-renderer work counter of the wavefront kernel
    int hardPhases;       // wavefront scheduling mode: 1 = block-synchronous stages, 0 = free-running warps
    int wideCta;          // 1 = 768-thread CTAs (deep mesh BVHs), 0 = 512
    int laneQueues;       // lane-owned slots with bit-set stage queues (free-running on-chip scenes): 0 never, 1 for big launches, 2 always
    int frame0;           // first frame (sample index per pixel)
    int numFrames;        // frames to trace in this launch
    int firstRow;         // pixel rows [firstRow, firstRow+numRows) are traced ...
    int numRows;
    int shard, numShards; // ... restricted to 4-row tile rows t with t % numShards == shard
    int tilesX;           // 8-pixel tile columns
    int tileRows;         // tile rows owned by this shard inside the row range
    unsigned long long samplesPerFrame;   // tilesX * tileRows * 32 (includes padding outside the image)
    // tb200_trace_frame outputs (nullptr in normal rendering)
    float* outRadiance;   // 3 floats per pixel
    float* outRaster;     // 2 floats per pixel
    // streamed read-back (tb200_render, single-frame launches; nullptr/0 otherwise): the sample range
    // is cut into bands of `bandSamples` consecutive indices (whole tile rows); the thread that
    // retires the last sample of a band stores `bandTag` into bandFlags[band] (mapped host memory),
    // so the host can copy finished rows while the kernel is still tracing the rest.
    unsigned int* bandCount;          // device counters, zeroed before the launch
    volatile unsigned int* bandFlags; // device pointer to pinned, mapped host memory
    unsigned int bandSamples;
    unsigned int bandTag;
    WalkParams walk;
    // cold half of the wavefront's slot state: 8 x float4 per slot, TB_WF2_PATHS slots per CTA of the launch
    float4* cold;
    int treeletBytes;     // shared memory behind the slot arrays that the prologue fills with the top of DScene::treeletMesh's BVH
};

#define TB_MAX_BANDS 64

// fills tilesX / tileRows / samplesPerFrame from the row range and shard
void finalize_params(LaunchParams* p);

// one thread per (pixel, frame): the whole path in registers.  Validation / fallback pipeline.
void launch_mega(const LaunchParams& p, cudaStream_t stream, unsigned long long* launchCount);
// persistent CTA-resident wavefront pipeline with a dedicated traversal stage (the product path)
void launch_wavefront2(const LaunchParams& p, int numSMs, cudaStream_t stream, unsigned long long* launchCount);
// eNormals mode (render.cpp:494-515)
void launch_normals(const LaunchParams& p, cudaStream_t stream, unsigned long long* launchCount);

// Display/finish step (src/main.cpp:258-271, src/png.cpp:329-343), see tb200_finish.
struct FinishParams {
    const float4* accum;      // width*height running sums
    int numPixels;
    float exposure;
    float4* filtered;         // device, or nullptr
    unsigned char* rgb8;      // device (numPixels*3 bytes, allocated in multiples of 12), or nullptr
    const uint2* ditherState; // WritePng's Random state in front of each pixel's six draws (when rgb8)
};
void launch_finish(const FinishParams& p, cudaStream_t stream, unsigned long long* launchCount);

// NonLocalMeansFilter (src/nlm.cpp:4-73) on the finished image; see tb200_nlm.
struct NlmParams {
    const float4* in;     // finished image (k_finish output)
    float4* means;        // scratch: AverageFilter output
    float4* out;
    int width, height, radius;
    float falloff;
};
void launch_nlm(const NlmParams& p, cudaStream_t stream, unsigned long long* launchCount);
