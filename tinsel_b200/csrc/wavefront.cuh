// wavefront.cuh -- the product pipeline: a persistent, CTA-resident wavefront path tracer.
//
// One CTA per SM stays resident for the whole launch and owns TB_WF_PATHS path slots whose
// state lives in SHARED MEMORY (SoA, 28 words per path).  The usual global-memory wavefront
// (tinsel's own unfinished wavefront.cu keeps ~150 B/path in HBM and re-reads it in five
// kernels per bounce, wavefront.cu:765-796,1357-1375) would make queue traffic, not the scene,
// the HBM consumer: at 1 Msample per wave that is gigabytes per spp.  Keeping the wavefront in
// the 227 KB of SMEM leaves only the scene (L1/L2 resident) and the framebuffer reductions on
// the memory system.
//
// Each iteration of the CTA loop advances every live path by one bounce in two stages:
//   stage 1  all slots   : regenerate dead slots from the global sample counter (warp-
//                          aggregated), trace the extension ray, shade the miss branch
//                          (sky, splat) or the hit prologue (absorption, emission MIS) and
//                          append surviving slots to the shading queue (ballot + prefix sum).
//   stage 2  compacted   : next-event estimation (probe + area lights, shadow rays traced in
//                          place), BSDF sampling, throughput update; finished paths splat and
//                          free their slot.
// Path regeneration keeps the slot array full until the sample counter runs dry, so both stages
// run on (nearly) full warps at every bounce depth.
#pragma once

#ifndef TB_WF_THREADS
#define TB_WF_THREADS 512
#endif
#ifndef TB_WF_PATHS
#define TB_WF_PATHS 1536
#endif
#define TB_WF_MAX_PRIMS 48      // scene tables up to this size are staged in shared memory
#define TB_WF_MAX_PAIRS 48

struct WfShared {
    // path state, SoA
    float ox[TB_WF_PATHS], oy[TB_WF_PATHS], oz[TB_WF_PATHS];
    float dx[TB_WF_PATHS], dy[TB_WF_PATHS], dz[TB_WF_PATHS];
    float time[TB_WF_PATHS];
    float Tx[TB_WF_PATHS], Ty[TB_WF_PATHS], Tz[TB_WF_PATHS];
    float Lx[TB_WF_PATHS], Ly[TB_WF_PATHS], Lz[TB_WF_PATHS];
    float eta[TB_WF_PATHS];
    float ax[TB_WF_PATHS], ay[TB_WF_PATHS], az[TB_WF_PATHS];
    float bsdfPdf[TB_WF_PATHS];
    uint32_t rng1[TB_WF_PATHS], rng2[TB_WF_PATHS];
    uint32_t sample[TB_WF_PATHS];     // sample index within the launch (pixel + frame)
    uint32_t flags[TB_WF_PATHS];      // bit0 alive, bits 1-2 rayType, bits 8.. bounce
    // hit record handed from stage 1 to stage 2
    float ht[TB_WF_PATHS], hnx[TB_WF_PATHS], hny[TB_WF_PATHS], hnz[TB_WF_PATHS];
    int hprim[TB_WF_PATHS];
    // shading queue
    uint16_t queue[TB_WF_PATHS];
    int queueCount;
    int contCount;                    // paths that survive stage 2
    // work distribution
    int exhausted;
    // scene tables staged on chip
    DPrim prims[TB_WF_MAX_PRIMS];
    BvhPair pairs[TB_WF_MAX_PAIRS];
    FlatNode flat[32];
};


TB_DEV void wf_load(const WfShared& S, int s, PathState& ps)
{
    ps.o = v3(S.ox[s], S.oy[s], S.oz[s]);
    ps.d = v3(S.dx[s], S.dy[s], S.dz[s]);
    ps.time = S.time[s];
    ps.T = v3(S.Tx[s], S.Ty[s], S.Tz[s]);
    ps.L = v3(S.Lx[s], S.Ly[s], S.Lz[s]);
    ps.eta = S.eta[s];
    ps.absorb = v3(S.ax[s], S.ay[s], S.az[s]);
    ps.bsdfPdf = S.bsdfPdf[s];
    ps.rng.s1 = S.rng1[s];
    ps.rng.s2 = S.rng2[s];
    ps.rayType = (int)((S.flags[s] >> 1) & 3u);
}

TB_DEV void wf_store(WfShared& S, int s, const PathState& ps, int bounce)
{
    S.ox[s] = ps.o.x; S.oy[s] = ps.o.y; S.oz[s] = ps.o.z;
    S.dx[s] = ps.d.x; S.dy[s] = ps.d.y; S.dz[s] = ps.d.z;
    S.time[s] = ps.time;
    S.Tx[s] = ps.T.x; S.Ty[s] = ps.T.y; S.Tz[s] = ps.T.z;
    S.Lx[s] = ps.L.x; S.Ly[s] = ps.L.y; S.Lz[s] = ps.L.z;
    S.eta[s] = ps.eta;
    S.ax[s] = ps.absorb.x; S.ay[s] = ps.absorb.y; S.az[s] = ps.absorb.z;
    S.bsdfPdf[s] = ps.bsdfPdf;
    S.rng1[s] = ps.rng.s1;
    S.rng2[s] = ps.rng.s2;
    S.flags[s] = 1u | ((uint32_t)ps.rayType << 1) | ((uint32_t)bounce << 8);
}

// raster position of a sample: its first two RNG draws (render.cpp:476,481-482)
TB_DEV void wf_raster_of(const LaunchParams& P, int px, int py, int frame, float& rx, float& ry)
{
    Rng rng = rng_seed(tb_sample_seed((uint32_t)(py * P.film.width + px), (uint32_t)frame));
    rx = rng_float(rng);
    ry = rng_float(rng);
    rx += px;
    ry += py;
}

TB_DEV void wf_finish(const LaunchParams& P, WfShared& S, int s, V3 radiance)
{
    int px, py, frame;
    decode_sample(P, (unsigned long long)S.sample[s], px, py, frame);
    float rx, ry;
    wf_raster_of(P, px, py, frame, rx, ry);
    sample_end(P, px, py, rx, ry, radiance);
    S.flags[s] = 0u;
}

// Claims one sample index for every lane with want==true (one global atomic per warp, ballot +
// prefix rank); returns false for lanes that got none because the launch's samples ran out.
TB_DEV bool wf_claim(const LaunchParams& P, WfShared& S, unsigned long long total, bool want, unsigned long long& idx)
{
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0u) return false;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0ull;
    if (lane == leader) {
        base = atomicAdd(P.sampleCounter, (unsigned long long)__popc(mask));
        if (base + __popc(mask) >= total) *(volatile int*)&S.exhausted = 1;
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    idx = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
    return want && idx < total;
}

__global__ void __launch_bounds__(TB_WF_THREADS, 1) k_wavefront(LaunchParams P, unsigned long long total)
{
    extern __shared__ __align__(16) unsigned char wf_smem_raw[];
    WfShared& S = *reinterpret_cast<WfShared*>(wf_smem_raw);
    const int tid = threadIdx.x;

    // ---- prologue: stage the scene tables on chip, clear the slots ---------------------------
    DScene sc = P.scene;
    if (sc.numPrims <= TB_WF_MAX_PRIMS) {
        const int words = sc.numPrims * (int)(sizeof(DPrim) / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.scene.prims);
        uint32_t* dst = reinterpret_cast<uint32_t*>(S.prims);
        for (int i = tid; i < words; i += TB_WF_THREADS) dst[i] = src[i];
        sc.prims = S.prims;
    }
    {
        const int numPairs = sc.numPairs;
        if (numPairs <= TB_WF_MAX_PAIRS) {
            const int words = numPairs * (int)(sizeof(BvhPair) / 4);
            const uint32_t* src = reinterpret_cast<const uint32_t*>(P.scene.pairs);
            uint32_t* dst = reinterpret_cast<uint32_t*>(S.pairs);
            for (int i = tid; i < words; i += TB_WF_THREADS) dst[i] = src[i];
            sc.pairs = S.pairs;
        }
    }
    if (sc.numFlat > 0 && sc.numFlat <= 32) {
        const int words = sc.numFlat * (int)(sizeof(FlatNode) / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.scene.flat);
        uint32_t* dst = reinterpret_cast<uint32_t*>(S.flat);
        for (int i = tid; i < words; i += TB_WF_THREADS) dst[i] = src[i];
        sc.flat = S.flat;
    }
    for (int s = tid; s < TB_WF_PATHS; s += TB_WF_THREADS) S.flags[s] = 0u;
    if (tid == 0) {
        S.queueCount = 0;
        S.contCount = 0;
        S.exhausted = 0;
    }
    __syncthreads();

    const int maxDepth = P.film.maxDepth;

    for (;;) {
        // ================= stage 1: regenerate + extend + miss / hit prologue ==================
        for (int s0 = 0; s0 < TB_WF_PATHS; s0 += TB_WF_THREADS) {
            const int s = s0 + tid;   // TB_WF_PATHS is a multiple of TB_WF_THREADS
            bool alive = (S.flags[s] & 1u) != 0u;
            PathState ps;
            int bounce = 0;

            // regenerate dead slots
            unsigned long long idx = 0ull;
            bool fresh = false;
            {
                bool want = !alive && !*(volatile int*)&S.exhausted;
                // a claimed index can fall on tile padding outside the image: claim again
                for (int attempt = 0; attempt < 64; ++attempt) {
                    const bool got = wf_claim(P, S, total, want, idx);
                    if (!__any_sync(0xffffffffu, want)) break;
                    if (want && got) {
                        int px, py, frame;
                        if (decode_sample(P, idx, px, py, frame)) {
                            float rx, ry;
                            sample_begin(P, px, py, frame, ps, rx, ry);
                            S.sample[s] = (uint32_t)idx;
                            fresh = true;
                            want = false;
                        }
                    } else if (want && !got) {
                        want = !*(volatile int*)&S.exhausted;
                    }
                }
            }
            if (fresh) {
                alive = true;
                bounce = 0;
            } else if (alive) {
                wf_load(S, s, ps);
                bounce = (int)(S.flags[s] >> 8);
            }

            bool toShade = false;
            if (alive) {
                if (maxDepth <= 0) {
                    wf_finish(P, S, s, ps.L);
                } else {
                    const Hit h = trace_closest(sc, ps.o, ps.d, ps.time, true);
                    if (h.prim < 0) {
                        path_miss(sc, ps, bounce);
                        wf_finish(P, S, s, ps.L);
                    } else {
                        // hit prologue (path_hit) is applied in stage 2 from the stored hit; here
                        // only the record is written so the slot can move to another thread
                        if (fresh) wf_store(S, s, ps, bounce);
                        S.ht[s] = h.t;
                        S.hnx[s] = h.n.x;
                        S.hny[s] = h.n.y;
                        S.hnz[s] = h.n.z;
                        S.hprim[s] = h.prim;
                        toShade = true;
                    }
                }
            }
            // append to the shading queue: ballot + prefix sum, one shared atomic per warp
            const unsigned m = __ballot_sync(0xffffffffu, toShade);
            if (m) {
                const int lane = tid & 31;
                int base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&S.queueCount, __popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (toShade) S.queue[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)s;
            }
        }
        __syncthreads();

        // ================= stage 2: NEE + scatter on the compacted queue =======================
        const int qn = S.queueCount;
        for (int q0 = 0; q0 < qn; q0 += TB_WF_THREADS) {
            const int q = q0 + tid;
            bool cont = false;
            if (q < qn) {
                const int s = S.queue[q];
                PathState ps;
                wf_load(S, s, ps);
                const int bounce = (int)(S.flags[s] >> 8);
                Hit h;
                h.t = S.ht[s];
                h.n = v3(S.hnx[s], S.hny[s], S.hnz[s]);
                h.prim = S.hprim[s];

                Surface sf;
                path_hit(sc, ps, h, bounce, sf);

                NeeCursor cur;
                nee_begin(cur);
                ShadowRay sr;
                while (nee_generate(sc, sf, ps.time, cur, ps.rng, sr)) {
                    const Hit sh = trace_closest(sc, sr.o, sr.d, ps.time, false);
                    nee_connect(sc, sf, sr, sh, cur);
                }
                const bool last = (bounce + 1 >= maxDepth);
                bool go;
                if (last) {
                    // the scattered ray of the final bounce is never traced (render.cpp:250):
                    // only the NEE sum is folded in
                    ps.L = ps.L + ps.T * cur.sum;
                    go = false;
                } else {
                    go = path_scatter(sc, ps, sf, cur.sum);
                }
                if (go) {
                    wf_store(S, s, ps, bounce + 1);
                    cont = true;
                } else {
                    wf_finish(P, S, s, ps.L);
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, cont);
            if (m && (tid & 31) == __ffs(m) - 1) atomicAdd(&S.contCount, __popc(m));
        }
        __syncthreads();
        const int continuing = S.contCount;
        const int done = *(volatile int*)&S.exhausted;
        __syncthreads();
        if (tid == 0) {
            S.queueCount = 0;
            S.contCount = 0;
        }
        __syncthreads();
        if (continuing == 0 && done) break;
    }
}

void launch_wavefront(const LaunchParams& p, int numSMs, cudaStream_t stream, unsigned long long* launchCount)
{
    const unsigned long long total = p.samplesPerFrame * (unsigned long long)p.numFrames;
    if (total == 0ull) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WfShared));
        configured = true;
    }
    cudaMemsetAsync(p.sampleCounter, 0, sizeof(unsigned long long), stream);
    // one resident CTA per SM; small jobs use fewer CTAs so that every CTA has a full slot array
    unsigned long long want = (total + TB_WF_PATHS - 1) / TB_WF_PATHS;
    int grid = numSMs > 0 ? numSMs : 148;
    if (want < (unsigned long long)grid) grid = (int)want;
    if (grid < 1) grid = 1;
    k_wavefront<<<grid, TB_WF_THREADS, sizeof(WfShared), stream>>>(p, total);
    if (launchCount) ++*launchCount;
}
