// wavefront2.cuh -- the product pipeline: a persistent, CTA-resident wavefront path tracer with a
// dedicated traversal stage.
//
// One CTA per SM stays resident for the whole launch and owns TB_WF2_PATHS path slots whose ENTIRE
// inter-stage state lives in SHARED MEMORY (SoA, 45 words per path: path state, pending shadow ray,
// hit records, next-event-estimation cursor).  A global-memory wavefront (tinsel's own unfinished
// wavefront.cu keeps ~150 B/path in DRAM and re-reads it in five kernels per bounce,
// wavefront.cu:765-796,1357-1375) would make queue traffic, not the scene, the HBM consumer; here
// HBM only sees scene misses and the framebuffer reductions.
//
// Every slot that is alive has exactly one pending ray: its extension ray or one shadow ray, and
// sits in exactly one CTA-shared stage queue.  Each WARP claims a chunk of up to 32 entries from a
// stage queue, processes it and appends the slots to the queues of their next stages:
//
//   R  finish + regenerate: splat the finished sample into the accumulator, claim a new sample
//      index from the global counter (one atomic per warp), generate its camera ray      -> T
//   T  trace the pending ray: closest hit -> hit record                        -> A (extension) / B (shadow)
//   A  extension result: miss -> sky -> R; hit -> absorption, emission MIS, first NEE sample -> T
//   B  shadow result: connect the NEE sample; next NEE sample -> T, or BSDF sample, throughput
//      update and next extension ray -> T, or termination -> R
//
// The queues are lock-free rings in shared memory: producers reserve cells with one warp-
// aggregated atomicAdd on `tail` (ballot + prefix sum) and store lap-tagged slot ids, so a chunk
// is always full while the queue holds >= 32 entries (compaction is implicit in the queue).
// Regeneration keeps the slots populated until the sample counter runs dry; the CTA exits when
// its last slot dies.
//
// Who runs what when is the scheduler's business, and there are three of them (template MODE,
// chosen per scene on the host, see the comment at the kernel): free-running warps with no block
// barrier at all and a shared stage preference that keeps them loosely in step (they share the
// instruction cache: 32 KB of L1.5 against ~150 KB of kernel); block-synchronous hard phases with
// ticket claims for scenes where every hit spawns several shadow rays; and free-running with a
// second trace queue for rays that enter a big mesh.
//
// This header is compiled THREE TIMES by kernels.cu, in three namespaces, for the three layouts of the slot
// state and its queues (chosen with TB_WF2_PATHS / TB_WF2_COLD_SMEM / TB_WF2_LANEQ before each inclusion):
//   wf2_smem  1024 slots per CTA, the whole state in shared memory, ring queues -- scenes held on chip under
//             the hard-phase scheduler (several shadow rays per hit), 512-thread CTAs
//   wf2_lq    the same state with LANE-OWNED slots and bit-set queues (see wf2_push) -- scenes held on chip
//             under the free-running scheduler: no bank conflicts, cheap queue operations, no FIFO order
//   wf2_l2    2048 slots per CTA, the cold half of the state in global memory (L2), ring queues -- scenes with
//             deep mesh BVHs (768-thread CTAs, and the offload mode), which gain from more paths in flight
//             and from the rings' FIFO order (rays pushed together are traced together)
// so it has no include guard, and only defines macros whose text is the same every time.

#ifndef TB_WF2_THREADS
#define TB_WF2_THREADS 512
#endif
#ifndef TB_WF2_THREADS_WIDE
#define TB_WF2_THREADS_WIDE 768
#endif
#ifndef TB_WF2_CTAS_PER_SM
#define TB_WF2_CTAS_PER_SM 1
#endif
static_assert(TB_WF2_PATHS <= TB_WF2_SLOTS, "the host sizes the cold slot state and the offload queues with TB_WF2_SLOTS");
#define TB_WF2_MAX_PRIMS 40     // scene tables up to this size are staged in shared memory (the flat program: up to 16 primitives)
#define TB_WF2_MAX_PAIRS 40

enum { WF2_PH_EXT = 0, WF2_PH_NEE = 1 };
#define WF2_FLAG_EMPTY 1u   // the slot holds no sample (before its first camera ray)

// Cold half of a slot's state: eight 16-byte chunks (seven used) per slot in global memory, read with
// ld.global.cg / written with st.global.cg by the one lane that owns the slot at the time (ownership
// moves through the stage queues, whose pushes are preceded by a block-scope fence).
//   0: T.xyz, eta        1: L.xyz, bsdfPdf     2: absorption.xyz, sample index
//   3: rng1, rng2, NEE cursor, sampled light   4: shadow dist, light normal.xyz
//   5: sky pdf, NEE sum.xyz                     6: NEE radiance accumulator.xyz
// TB_WF2_COLD_SMEM: the same records in SHARED memory instead (stride 7 chunks), for builds with few enough
// slots per CTA: the cold half then costs 7 128-bit accesses per stage instead of 27 32-bit ones.
#ifdef TB_WF2_COLD_SMEM
enum { CC_T = 0, CC_L = 1, CC_A = 2, CC_RNG = 3, CC_SH = 4, CC_SUM = 5, CC_LA = 6, CC_CHUNKS = 7 };
TB_DEV float4 cold_ld(const float4* rec, int c) { return rec[c]; }
TB_DEV void cold_st(float4* rec, int c, float x, float y, float z, float w) { rec[c] = make_float4(x, y, z, w); }
#else
enum { CC_T = 0, CC_L = 1, CC_A = 2, CC_RNG = 3, CC_SH = 4, CC_SUM = 5, CC_LA = 6, CC_CHUNKS = 8 };
TB_DEV float4 cold_ld(const float4* rec, int c) { return __ldcg(rec + c); }
TB_DEV void cold_st(float4* rec, int c, float x, float y, float z, float w) { __stcg(rec + c, make_float4(x, y, z, w)); }
#endif
TB_DEV float u2f(uint32_t v) { return __uint_as_float(v); }
TB_DEV uint32_t f2u(float v) { return __float_as_uint(v); }

struct Wf2Shared {
    // HOT half of the slot state: what stage T (trace) reads and writes for every ray.  The cold half --
    // throughput, radiance, medium, RNG, the NEE bookkeeping: 27 words that only stages R, A and B touch,
    // once per ray -- lives in a 128-byte record per slot in global memory (Wf2Cold, L2-resident).  The
    // wavefront is bound by the number of paths an SM has in flight (halving the slots halves the
    // throughput on every scene, profiles/README.md round 2), and 18 words instead of 45 per slot double it.
    float ox[TB_WF2_PATHS], oy[TB_WF2_PATHS], oz[TB_WF2_PATHS];
    float dx[TB_WF2_PATHS], dy[TB_WF2_PATHS], dz[TB_WF2_PATHS];
    float time[TB_WF2_PATHS];
    uint32_t flags[TB_WF2_PATHS];     // bit 0 empty (no sample yet), bits 1-2 rayType, bit 3 phase, bits 8.. bounce
    // extension-ray hit
    float ht[TB_WF2_PATHS], hnx[TB_WF2_PATHS], hny[TB_WF2_PATHS], hnz[TB_WF2_PATHS];
    int hprim[TB_WF2_PATHS];
    // pending shadow ray (origin is recomputed from the surface point) and its result
    float sdx[TB_WF2_PATHS], sdy[TB_WF2_PATHS], sdz[TB_WF2_PATHS];
    float st[TB_WF2_PATHS];
    int sprim[TB_WF2_PATHS];
#ifdef TB_WF2_COLD_SMEM
    float4 cold[TB_WF2_PATHS * 7];
#endif
#ifdef TB_WF2_LANEQ
    // stage queues as bit sets (lane-owned slots, see wf2_push below): one word per queue and lane
    unsigned int lq[4][32];
#else
    // stage queues: rings of lap-tagged slot ids
    uint16_t ring[7][TB_WF2_PATHS];
    unsigned int head[7], tail[7];
    unsigned int snap[5];             // hard-phase mode: the tail each stage of the current phase runs with
#endif
    int pref;                         // stage the warps currently prefer (soft phases, see the main loop)
    int live;                         // slots that still hold (or may still receive) a path
    int exhausted;
    // offload mode: heads of this CTA's two answer rings (extension / shadow), answers still owed to it,
    // and whether the CTA has told the walkers that it is done
    unsigned int ansHead[2];
    int ansPending[2];
    int exitSignaled;
    // scene tables staged on chip by bulk copies (TMA) that complete on `stageBar`
    alignas(16) DPrim prims[TB_WF2_MAX_PRIMS + 1];   // +1: the copy length is rounded up to 16 bytes
    alignas(16) BvhPair pairs[TB_WF2_MAX_PAIRS];
    alignas(16) ProgOp flat[32];
    alignas(8) unsigned long long stageBar;
};

// dynamic shared memory of a CTA: the slot arrays; a build with fewer slots (experiments) still claims
// a whole SM, so that the slot count per SM is what changes
#ifdef TB_WF2_PAD_SMEM
#define TB_WF2_SMEM_BYTES (sizeof(Wf2Shared) > 204000 ? sizeof(Wf2Shared) : (size_t)204000)
#else
#define TB_WF2_SMEM_BYTES sizeof(Wf2Shared)
#endif

// stage queues.  T, A, B, R in the cyclic order of the free-running sweep; F0/F1 hold the freshly
// regenerated camera rays in hard-phase mode (double-buffered: R fills one while T drains the other)
// TM (free-running mode, scenes with DScene::splitValid): rays that enter the big mesh's box; same
// stage code as T, but a chunk taken from TM walks the mesh with all of its lanes and a chunk taken
// from T with none.
enum { WF2_Q_T = 0, WF2_Q_A = 1, WF2_Q_B = 2, WF2_Q_R = 3, WF2_Q_F0 = 4, WF2_Q_F1 = 5, WF2_Q_TM = 6 };
#define WF2_MASK (TB_WF2_PATHS - 1)
#define WF2_LOG2_PATHS (TB_WF2_PATHS == 2048 ? 11 : TB_WF2_PATHS == 1024 ? 10 : TB_WF2_PATHS == 512 ? 9 : TB_WF2_PATHS == 256 ? 8 : 7)
#define WF2_SLOT_BITS 11                      // queue cells: slot id in the low 11 bits, lap tag (1..31) above
#define WF2_SLOT_MASK ((1u << WF2_SLOT_BITS) - 1u)
static_assert(TB_WF2_PATHS <= (1 << WF2_SLOT_BITS) && (TB_WF2_PATHS & (TB_WF2_PATHS - 1)) == 0, "slot ids must fit the queue cells");

// cell value of ring index i holding slot s: the lap tag (1..31, never 0) makes "reserved but not
// yet written" and "left over from the previous lap" distinguishable from the expected entry
TB_DEV uint16_t wf2_cell(unsigned int index, int slot)
{
    const unsigned int lap = ((index >> WF2_LOG2_PATHS) % 31u) + 1u;
    return (uint16_t)((unsigned)slot | (lap << WF2_SLOT_BITS));
}

#ifdef TB_WF2_LANEQ
// Lane-owned slots (the third layout, wf2_lq: scenes held on chip, free-running warps).  Slot s is only ever
// handled by lane (s & 31) of whichever warp claims it, so every access to the slot arrays is free of bank
// conflicts (lane i touches bank i; with ring queues the slot ids of a chunk are a random permutation and 35 %
// of the kernel's shared-memory wavefronts were conflicts), and a stage queue needs no ring: it is one bit per
// slot, the 32 slots of a lane in one word.  A push is one atomicOr by the owning lane on its own word -- no
// ballot, no prefix sum, no reserved-but-unwritten cell for a consumer to trip over.  The price is the FIFO
// order (a chunk is no longer a run of rays that were pushed together), which costs scenes with deep BVHs
// their traversal coherence: those keep the ring queues.
// The caller has fenced its slot-state writes (__threadfence_block).
static_assert(TB_WF2_PATHS == 1024, "one word per lane and queue");
#ifndef WF2_STICK_LANES
#define WF2_STICK_LANES 16     // a warp stays with the preferred stage while that many lanes have a slot waiting there
#endif
TB_DEV void wf2_push(Wf2Shared& S, int q, bool flag, int slot)
{
    if (flag) atomicOr(&S.lq[q][slot & 31], 1u << (slot >> 5));
}
#else
// append `slot` to stage queue q for every lane with flag == true: ballot + prefix sum, one shared
// atomic per warp.  The caller has fenced its slot-state writes (__threadfence_block).
TB_DEV void wf2_push(Wf2Shared& S, int q, bool flag, int slot)
{
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(&S.tail[q], (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (flag) {
        const unsigned int idx = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        *(volatile uint16_t*)&S.ring[q][idx & WF2_MASK] = wf2_cell(idx, slot);
    }
}
#endif

// Push a slot whose pending ray is ready to be traced.  With `split` set the ray (extension ray, or
// the shadow ray from the hit point) is classified against the big mesh's bounds first; the test
// is a scheduling heuristic, so the unoffset hit point is good enough as the shadow-ray origin.
TB_DEV void wf2_push_trace(Wf2Shared& S, const DScene& sc, bool split, int q, bool flag, int s)
{
    if (!split) {
        wf2_push(S, q, flag, s);
        return;
    }
    bool big = false;
    if (flag) {
        V3 o = v3(S.ox[s], S.oy[s], S.oz[s]);
        V3 d = v3(S.dx[s], S.dy[s], S.dz[s]);
        if (((S.flags[s] >> 3) & 1u) != WF2_PH_EXT) {
            o = o + d * S.ht[s];
            d = v3(S.sdx[s], S.sdy[s], S.sdz[s]);
        }
        V3 rcp;
        rcp.x = 1.0f / d.x;
        rcp.y = 1.0f / d.y;
        rcp.z = 1.0f / d.z;
        float t;
        big = ray_aabb(o, rcp, sc.splitLo.x, sc.splitLo.y, sc.splitLo.z, sc.splitHi.x, sc.splitHi.y, sc.splitHi.z, t);
    }
    wf2_push(S, q, flag && !big, s);
    wf2_push(S, WF2_Q_TM, flag && big, s);
}

#ifdef TB_WF2_LANEQ
// Claim one waiting slot per lane.  All four queues (T, A, B, R) are looked at in one go -- four independent
// loads of the lane's own words and four ballots -- and the warp takes the CTA's preferred stage while at least
// `stick` lanes have something there, else the stage with the most lanes ready (which becomes the preference:
// the warps of a CTA herd through the stages, and share the instruction cache -- a per-warp preference instead
// of the CTA-wide one costs 35 %).  Every lane then picks a set bit of its word -- starting at a position that
// differs from warp to warp, so that two warps claiming at once rarely want the same one -- and clears it with
// an atomicAnd; whoever sees the bit set in the value returned owns the slot.  Returns the mask of lanes that
// got a slot (0: nothing waits anywhere) and the stage.
TB_DEV unsigned wf2_claim_fullest(Wf2Shared& S, int stick, int& slot, int& stage)
{
    const int lane = threadIdx.x & 31;
    const int rot = (int)(threadIdx.x >> 5) * 2 + 1;
    for (;;) {
        // the preference and the four queue words load side by side (no load depends on another)
        int pref = *(volatile int*)&S.pref;   // every lane loads (a broadcast, no branch); lane 0's copy is the one used
        const unsigned m0 = *(volatile unsigned int*)&S.lq[0][lane], m1 = *(volatile unsigned int*)&S.lq[1][lane];
        const unsigned m2 = *(volatile unsigned int*)&S.lq[2][lane], m3 = *(volatile unsigned int*)&S.lq[3][lane];
        pref = __shfl_sync(0xffffffffu, pref, 0);
        // lanes ready per queue, one byte each (selections by shifts, not branches)
        const unsigned cc = (unsigned)__popc(__ballot_sync(0xffffffffu, m0 != 0u)) | ((unsigned)__popc(__ballot_sync(0xffffffffu, m1 != 0u)) << 8) |
                            ((unsigned)__popc(__ballot_sync(0xffffffffu, m2 != 0u)) << 16) | ((unsigned)__popc(__ballot_sync(0xffffffffu, m3 != 0u)) << 24);
        int q = pref;
        unsigned bc = (cc >> (8 * pref)) & 0xffu;
        if (bc < (unsigned)stick) {
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const int qq = (pref + k) & 3;
                const unsigned cq = (cc >> (8 * qq)) & 0xffu;
                q = cq > bc ? qq : q;
                bc = cq > bc ? cq : bc;
            }
        }
        if (bc == 0u) return 0u;
        const unsigned lo = (q & 1) ? m1 : m0, hi = (q & 1) ? m3 : m2;
        const unsigned mine = (q & 2) ? hi : lo;
        bool got = false;
        if (mine != 0u) {
            const int k = (__ffs(__funnelshift_r(mine, mine, rot)) - 1 + rot) & 31;
            const unsigned bit = 1u << k;
            got = (atomicAnd(&S.lq[q][lane], ~bit) & bit) != 0u;
            slot = lane + 32 * k;
        }
        const unsigned act = __ballot_sync(0xffffffffu, got);
        if (act != 0u) {
            __threadfence_block();   // the producer's state writes precede its atomicOr
            if (q != pref && lane == 0) *(volatile int*)&S.pref = q;
            stage = q;
            return act;
        }
        // every lane lost its bit to another warp, which made progress: look again
    }
}
#else
// Claim up to 32 entries of queue q for this warp.  Returns the number claimed (warp-uniform);
// lane i < n receives its slot.  Lock-free: cells are read first, ownership is taken with a CAS
// on head, so a stalled warp can never read a cell that a producer has already recycled.
TB_DEV int wf2_claim(Wf2Shared& S, int q, int minCount, int& slot)
{
    const int lane = threadIdx.x & 31;
    int notReady = 0;
    for (;;) {
        unsigned int h = 0, t = 0;
        if (lane == 0) {
            h = *(volatile unsigned int*)&S.head[q];
            t = *(volatile unsigned int*)&S.tail[q];
        }
        h = __shfl_sync(0xffffffffu, h, 0);
        t = __shfl_sync(0xffffffffu, t, 0);
        int n = (int)(t - h);
        if (n < minCount) return 0;
        if (n > 32) n = 32;
        // optimistic read; entries reserved by a producer but not yet stored end the chunk early
        const unsigned int idx = h + (unsigned)lane;
        const uint16_t cell = *(volatile uint16_t*)&S.ring[q][idx & WF2_MASK];
        const bool ok = lane < n && (unsigned)(cell >> WF2_SLOT_BITS) == (((idx >> WF2_LOG2_PATHS) % 31u) + 1u);
        const unsigned good = __ballot_sync(0xffffffffu, ok);
        n = (good == 0xffffffffu) ? 32 : __ffs(~good) - 1;   // length of the valid prefix
        if (n <= 0) {
            if (++notReady > 8) return 0;   // a producer is between its reservation and its store
            continue;
        }
        unsigned int got = 0;
        if (lane == 0) got = atomicCAS(&S.head[q], h, h + (unsigned)n);
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got == h) {
            __threadfence_block();   // the producer's state writes precede its cell store
            slot = (int)(cell & WF2_SLOT_MASK);
            return n;
        }
        // lost the race to another warp, which made progress: try again
    }
}

// Hard-phase variant: while a stage runs nobody pushes into its queue, so `tail` is fixed and a
// ticket (atomicAdd on head) can never overshoot into entries that do not exist yet.  `head` is
// put back to `tail` by thread 0 after the barrier that ends the stage.
TB_DEV int wf2_claim_ticket(Wf2Shared& S, int q, unsigned int tail, int& slot)
{
    const int lane = threadIdx.x & 31;
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(&S.head[q], 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    const int avail = (int)(tail - base);
    if (avail <= 0) return 0;
    const int n = avail < 32 ? avail : 32;
    if (lane < n) slot = (int)(*(volatile uint16_t*)&S.ring[q][(base + (unsigned)lane) & WF2_MASK] & WF2_SLOT_MASK);
    return n;
}

// the same claims as lane masks (lanes 0..n-1), the form the kernel body works with
TB_DEV unsigned wf2_lanes(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }
#endif

TB_DEV Surface wf2_surface(const Wf2Shared& S, const DScene& sc, int s, float eta)
{
    // the quantities path_hit() derives from the path and the hit record (render.cpp:255-278)
    Surface sf;
    const V3 o = v3(S.ox[s], S.oy[s], S.oz[s]);
    const V3 d = v3(S.dx[s], S.dy[s], S.dz[s]);
    sf.prim = S.hprim[s];
    sf.p = o + d * S.ht[s];
    sf.n = v3(S.hnx[s], S.hny[s], S.hnz[s]);
    sf.wo = -d;
    sf.etaI = eta;
    const DPrim& prim = sc.prims[sf.prim];
    if (sf.etaI == 1.0f) {
        sf.etaO = prim.mat.ior;
        sf.outAbsorb = prim.mat.absorption;
    } else {
        sf.etaO = 1.0f;
        sf.outAbsorb = v3s(0.0f);
    }
    return sf;
}

// Streamed read-back bookkeeping (LaunchParams::bandFlags): called by a whole warp after its lanes
// with `have` set have retired sample `idx` (splatted it, or skipped it as tile padding).  Every
// lane fences its own framebuffer reductions, equal bands are counted with one atomic, and whoever
// completes a band publishes it to the host.
static __device__ __noinline__ void wf2_band_report(unsigned int* bandCount, volatile unsigned int* bandFlags, uint32_t bandSamples,
                                                    uint32_t samplesPerFrame, uint32_t bandTag, uint32_t idx, bool have)
{
    __threadfence();
    const unsigned act = __ballot_sync(0xffffffffu, have);
    if (!have) return;
    const uint32_t band = idx / bandSamples;
    const unsigned peers = __match_any_sync(act, band);
    if ((int)(threadIdx.x & 31u) != __ffs(peers) - 1) return;
    const uint32_t n = (uint32_t)__popc(peers);
    const uint32_t size = min(bandSamples, samplesPerFrame - band * bandSamples);
    if (atomicAdd(&bandCount[band], n) + n == size) {
        __threadfence_system();
        bandFlags[band] = bandTag;
    }
}

static __device__ __noinline__ void wf2_band_pad(unsigned int* bandCount, volatile unsigned int* bandFlags, uint32_t bandSamples,
                                                 uint32_t samplesPerFrame, uint32_t bandTag, uint32_t idx)
{
    const uint32_t band = idx / bandSamples;
    const uint32_t size = min(bandSamples, samplesPerFrame - band * bandSamples);
    if (atomicAdd(&bandCount[band], 1u) + 1u == size) {
        __threadfence_system();
        bandFlags[band] = bandTag;
    }
}

// the pending shadow ray and the NEE bookkeeping that waits for its result
TB_DEV void wf2_store_shadow(Wf2Shared& S, float4* rec, int s, const ShadowRay& sr, const NeeCursor& c, Rng rng)
{
    S.sdx[s] = sr.d.x; S.sdy[s] = sr.d.y; S.sdz[s] = sr.d.z;
    const uint32_t cursor = (uint32_t)c.slot | ((uint32_t)c.prim << 8) | ((uint32_t)c.sample << 20);
    cold_st(rec, CC_RNG, u2f(rng.s1), u2f(rng.s2), u2f(cursor), __int_as_float(sr.light));
    cold_st(rec, CC_SH, sr.dist, sr.lightN.x, sr.lightN.y, sr.lightN.z);
    cold_st(rec, CC_SUM, sr.skyPdf, c.sum.x, c.sum.y, c.sum.z);
    cold_st(rec, CC_LA, c.Lacc.x, c.Lacc.y, c.Lacc.z, 0.0f);
}

// finish the shading of a surface hit once all NEE samples are folded: BSDF sample, next ray
// (path_scatter) or termination.  Returns true when the slot continues with a new extension ray.
TB_DEV bool wf2_scatter(Wf2Shared& S, float4* rec, const DScene& sc, int s, const Surface& sf, V3 T, V3 L, Rng rng, float time, V3 neeSum,
                        int bounce, int maxDepth, V3 absorb, float bsdfPdf, uint32_t sample)
{
    PathState ps;
    ps.o = v3s(0.0f);
    ps.d = v3s(0.0f);
    ps.time = time;
    ps.T = T;
    ps.L = L;
    ps.eta = sf.etaI;
    ps.absorb = absorb;
    ps.rayType = (int)((S.flags[s] >> 1) & 3u);
    ps.bsdfPdf = bsdfPdf;
    ps.rng = rng;
    bool go;
    if (bounce + 1 >= maxDepth) {
        // the scattered ray of the final bounce is never traced (render.cpp:250): fold NEE only
        ps.L = ps.L + ps.T * neeSum;
        go = false;
    } else {
        go = path_scatter(sc, ps, sf, neeSum);
    }
    cold_st(rec, CC_L, ps.L.x, ps.L.y, ps.L.z, ps.bsdfPdf);
    if (!go) return false;
    S.ox[s] = ps.o.x; S.oy[s] = ps.o.y; S.oz[s] = ps.o.z;
    S.dx[s] = ps.d.x; S.dy[s] = ps.d.y; S.dz[s] = ps.d.z;
    cold_st(rec, CC_T, ps.T.x, ps.T.y, ps.T.z, ps.eta);
    cold_st(rec, CC_A, ps.absorb.x, ps.absorb.y, ps.absorb.z, u2f(sample));
    cold_st(rec, CC_RNG, u2f(ps.rng.s1), u2f(ps.rng.s2), 0.0f, 0.0f);
    S.flags[s] = ((uint32_t)ps.rayType << 1) | ((uint32_t)WF2_PH_EXT << 3) | ((uint32_t)(bounce + 1) << 8);
    return true;
}

#include "wavefront_walk.cuh"

// Claim up to 32 answers of the walkers for this CTA (offload mode): lane i < n receives the three
// chunks of its answer.  Cells are validated by their lap tags, ownership is taken with a CAS on the
// ring's head in shared memory (only this CTA's warps consume the ring).
TB_DEV int wf2_claim_answers(Wf2Shared& S, const WalkParams& W, int kind, int minCount, uint4& c0, uint4& c1, uint4& c2)
{
    const int lane = threadIdx.x & 31;
    unsigned int h = 0;
    int owed = 0;
    if (lane == 0) {
        owed = *(volatile int*)&S.ansPending[kind];
        h = *(volatile unsigned int*)&S.ansHead[kind];
    }
    owed = __shfl_sync(0xffffffffu, owed, 0);
    if (owed < minCount) return 0;   // nothing (or not enough) can have arrived: skip the round trip to L2
    h = __shfl_sync(0xffffffffu, h, 0);
    const unsigned int idx = h + (unsigned)lane;
    const unsigned int tag = (idx >> WF2_LOG2_PATHS) + 1u;
    const uint4* cell = W.ansRing + ((size_t)(blockIdx.x * 2 + kind) * TB_WF2_PATHS + (idx & WF2_MASK)) * 3;
    c0 = walk_ld(cell + 0);
    c1 = walk_ld(cell + 1);
    c2 = walk_ld(cell + 2);
    const bool ok = c0.w == tag && c1.w == tag && c2.w == tag;
    const unsigned good = __ballot_sync(0xffffffffu, ok);
    const int n = (good == 0xffffffffu) ? 32 : __ffs(~good) - 1;
    if (n <= 0 || n < minCount) return 0;
    unsigned int got = 0;
    if (lane == 0) got = atomicCAS(&S.ansHead[kind], h, h + (unsigned)n);
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got != h) return 0;   // another warp of the CTA took them
    if (lane == 0) atomicSub(&S.ansPending[kind], n);
    return n;
}

// Fold a walker's answer into the partial hit stage T left in the slot (offload mode).  Afterwards
// the slot holds exactly what trace_closest() would have produced for the pending ray.
TB_DEV void wf2_merge_answer(Wf2Shared& S, const DScene& sc, int s, bool isExt, uint4 c0, uint4 c1, uint4 c2)
{
    const uint32_t info = c2.z;
    const bool hit = ((info >> (WF2_SLOT_BITS + 8)) & 1u) != 0u, tie = ((info >> (WF2_SLOT_BITS + 9)) & 1u) != 0u;
    const int primM = (int)((info >> WF2_SLOT_BITS) & 0xffu);
    const float tM = __uint_as_float(c0.x);
    const float partialT = isExt ? S.ht[s] : S.st[s];
    const int partialPrim = isExt ? S.hprim[s] : S.sprim[s];
    V3 o = v3(S.ox[s], S.oy[s], S.oz[s]);
    V3 d = v3(S.dx[s], S.dy[s], S.dz[s]);
    const float time = S.time[s];
    if (!isExt) {
        // the shadow ray, as stage T built it (render.cpp:121,170)
        const V3 p = o + d * S.ht[s];
        const V3 nn = v3(S.hnx[s], S.hny[s], S.hnz[s]);
        d = v3(S.sdx[s], S.sdy[s], S.sdz[s]);
        o = p + face_forward(nn, d) * TB_RAY_EPS;
    }
    if (tie || (hit && partialPrim >= 0 && tM == partialT)) {
        // two primitives at exactly the same t: the reference's visit order decides -- redo in that order
        const Hit h = trace_ordered(sc, o, d, time, isExt);
        if (isExt) {
            S.ht[s] = h.t;
            S.hnx[s] = h.n.x; S.hny[s] = h.n.y; S.hnz[s] = h.n.z;
            S.hprim[s] = h.prim;
        } else {
            S.st[s] = h.t;
            S.sprim[s] = h.prim;
        }
        return;
    }
    if (!hit || !(tM < partialT)) return;   // the partial hit stands
    if (isExt) {
        PrimHit ph;
        ph.t = tM;
        ph.u = __uint_as_float(c0.y);
        ph.v = __uint_as_float(c0.z);
        ph.w = __uint_as_float(c1.x);
        ph.gn = v3(__uint_as_float(c1.y), __uint_as_float(c1.z), __uint_as_float(c2.x));
        ph.tri = (int)c2.y;
        const V3 n = face_forward(prim_normal(sc, sc.prims[primM], o, d, time, ph), -d);
        S.ht[s] = tM;
        S.hnx[s] = n.x; S.hny[s] = n.y; S.hnz[s] = n.z;
        S.hprim[s] = primM;
    } else {
        S.st[s] = tM;
        S.sprim[s] = primM;
    }
}

// THREADS: 512 (16 warps, up to 128 registers) for scenes held in shared memory, where more warps
// only add instruction-cache pressure; 768 (24 warps, 80 registers) for scenes with deep mesh BVHs,
// whose traversal is latency bound on L2 and pays for the extra warps (LaunchParams::wideCta).
// MODE selects the scheduler that is compiled in -- measured, not principled: the kernel's speed is
// sensitive to register allocation and code layout, so every variant that is not needed is kept out
// of the others' bodies (profiles/README.md, steps 12-14):
//   WF2_MODE_GENERIC  both schedulers, run-time flag, no split trace queue (the free-running default;
//                     its own specialisation measured 5 % slower on cornell)
//   WF2_MODE_HARD     hard phases only (veach +4 %)
//   WF2_MODE_SPLIT    free-running with the split trace queue TM (scenes with a big mesh)
//   WF2_MODE_OFFLOAD  free-running shader CTAs + walker CTAs that do the big meshes' BVH walks (wavefront_walk.cuh)
enum { WF2_MODE_GENERIC = 0, WF2_MODE_HARD = 1, WF2_MODE_SPLIT = 2, WF2_MODE_OFFLOAD = 3 };

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, TB_WF2_CTAS_PER_SM) k_wavefront2(LaunchParams P, unsigned long long total)
{
    extern __shared__ __align__(16) unsigned char wf_smem_raw[];
    Wf2Shared& S = *reinterpret_cast<Wf2Shared*>(wf_smem_raw);
    const int tid = threadIdx.x;
    constexpr bool offload = MODE == WF2_MODE_OFFLOAD;
    if (offload && (int)blockIdx.x >= P.walk.numShaders) {
        wf2_walker_role<THREADS>(P, wf_smem_raw, (int)TB_WF2_SMEM_BYTES);
        return;
    }

    // ---- prologue: stage the scene tables on chip ---------------------------------------------
    // Primitive records, the scene-level BVH (child-pair records) and the flat scene program go from
    // global to shared memory as TMA bulk copies (cp.async.bulk -> UBLKCP): thread 0 arms an mbarrier
    // with the byte count and issues up to three copies, everybody waits on the barrier's phase.  (The
    // walker CTAs of the offload mode stage the top of the big mesh's BVH the same way.)
    DScene sc = P.scene;
    const bool stagePrims = sc.numPrims <= TB_WF2_MAX_PRIMS;
    const bool stagePairs = sc.numPairs > 0 && sc.numPairs <= TB_WF2_MAX_PAIRS;
    const bool stageFlat = sc.numFlat > 0 && sc.numFlat <= 32;
    if (tid == 0) {
        const uint32_t bar = walk_smem_addr(&S.stageBar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // device arrays are padded by one element (api.cu: upload_image), so rounding up stays inside them
        const uint32_t primBytes = stagePrims ? (((uint32_t)sc.numPrims * (uint32_t)sizeof(DPrim) + 15u) & ~15u) : 0u;
        const uint32_t pairBytes = stagePairs ? (uint32_t)sc.numPairs * (uint32_t)sizeof(BvhPair) : 0u;
        const uint32_t flatBytes = stageFlat ? (uint32_t)sc.numFlat * (uint32_t)sizeof(ProgOp) : 0u;
        const bool stageTreelet = P.treeletBytes >= (int)sizeof(BvhPair) && P.scene.treeletMesh >= 0;
        const uint32_t treeBytes = stageTreelet ? (uint32_t)min(P.scene.treeletPairs, P.treeletBytes / (int)sizeof(BvhPair)) * (uint32_t)sizeof(BvhPair) : 0u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(primBytes + pairBytes + flatBytes + treeBytes) : "memory");
        if (treeBytes) {
            unsigned char* tl = wf_smem_raw + ((sizeof(Wf2Shared) + 127) & ~size_t(127));
            const unsigned char* src = reinterpret_cast<const unsigned char*>(P.scene.meshes[P.scene.treeletMesh].pairs);
            for (uint32_t off = 0; off < treeBytes; off += 32768u) {
                const uint32_t n = min(32768u, treeBytes - off);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(walk_smem_addr(tl + off)),
                             "l"(src + off), "r"(n), "r"(bar)
                             : "memory");
            }
        }
        if (primBytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(walk_smem_addr(S.prims)),
                         "l"(P.scene.prims), "r"(primBytes), "r"(bar)
                         : "memory");
        if (pairBytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(walk_smem_addr(S.pairs)),
                         "l"(P.scene.pairs), "r"(pairBytes), "r"(bar)
                         : "memory");
        if (flatBytes)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(walk_smem_addr(S.flat)),
                         "l"(P.scene.flat), "r"(flatBytes), "r"(bar)
                         : "memory");
    }
    if (stagePrims) sc.prims = S.prims;
    if (stagePairs) sc.pairs = S.pairs;
    if (stageFlat) sc.flat = S.flat;
    // the top of the biggest mesh's BVH goes into whatever shared memory the slot arrays leave (P.treeletBytes,
    // set by the launch from the device's limit): the first levels of every inline mesh walk are then shared-
    // memory reads instead of dependent L2 round trips
    sc.treelet = nullptr;
    if (P.treeletBytes >= (int)sizeof(BvhPair) && sc.treeletMesh >= 0) {
        unsigned char* tl = wf_smem_raw + ((sizeof(Wf2Shared) + 127) & ~size_t(127));
        sc.treeletPairs = min(sc.treeletPairs, P.treeletBytes / (int)sizeof(BvhPair));
        sc.treelet = walk_opaque(tl);
    } else {
        sc.treeletPairs = 0;
    }
    // every slot starts in the R queue "finished with nothing to splat": stage R fills it with a
    // camera sample
#ifdef TB_WF2_LANEQ
    for (int s = tid; s < TB_WF2_PATHS; s += THREADS) S.flags[s] = WF2_FLAG_EMPTY;
    if (tid < 4 * 32) S.lq[tid >> 5][tid & 31] = (tid >> 5) == WF2_Q_R ? 0xffffffffu : 0u;
    if (tid == 0) {
#else
    for (int s = tid; s < TB_WF2_PATHS; s += THREADS) {
        S.ring[WF2_Q_R][s] = wf2_cell((unsigned)s, s);
        S.ring[WF2_Q_T][s] = 0;
        S.ring[WF2_Q_A][s] = 0;
        S.ring[WF2_Q_B][s] = 0;
        S.ring[WF2_Q_F0][s] = 0;
        S.ring[WF2_Q_F1][s] = 0;
        S.ring[WF2_Q_TM][s] = 0;
        S.flags[s] = WF2_FLAG_EMPTY;
    }
    if (tid == 0) {
        for (int q = 0; q < 7; ++q) S.head[q] = S.tail[q] = 0u;
        S.tail[WF2_Q_R] = TB_WF2_PATHS;
        S.snap[0] = TB_WF2_PATHS;
        S.snap[1] = S.snap[2] = S.snap[3] = S.snap[4] = 0u;
#endif
        S.pref = WF2_Q_R;
        S.live = TB_WF2_PATHS;
        S.exhausted = 0;
        S.ansPending[0] = S.ansPending[1] = 0;
        S.exitSignaled = 0;
        if (offload) {
            // everything the walkers ever answered has been consumed: the rings' heads are their tails
            S.ansHead[0] = walk_ld1(P.walk.ansTail + blockIdx.x * 2 + 0);
            S.ansHead[1] = walk_ld1(P.walk.ansTail + blockIdx.x * 2 + 1);
        }
    }
    __syncthreads();   // the barrier word is initialised (thread 0, above) before anybody polls it
    {
        uint32_t landed = 0;
        while (!landed)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(landed)
                         : "r"(walk_smem_addr(&S.stageBar)), "r"(0u)
                         : "memory");
    }

    const int maxDepth = P.film.maxDepth;
    const int lane = tid & 31;
#ifdef TB_WF2_COLD_SMEM
    float4* const coldBase = S.cold;
#else
    float4* const coldBase = P.cold + (size_t)blockIdx.x * TB_WF2_PATHS * CC_CHUNKS;
#endif
#ifdef WALK_SHADER_WARPS
    // offload mode: with most slots parked at the walkers, fewer warps per shader CTA keep fuller chunks and
    // fewer distinct stages in flight (experiment knob)
    if (offload && (tid >> 5) >= WALK_SHADER_WARPS) return;
#endif

    // Two scheduling modes share the stage code below (P.hardPhases, uniform for the launch):
    //
    //  * hard phases: all warps of the CTA work through the same short list of stages, claiming
    //    chunks dynamically until each queue is empty, and meet at a block barrier twice per cycle.
    //    The kernel is far larger than the instruction cache, so warps that execute the same
    //    code at the same time share every instruction line; this is the faster mode when rays
    //    cost about the same (small scenes held in shared memory).
    //  * free running: no barriers at all; every warp picks the next stage itself, starting at a
    //    CTA-wide preferred stage and moving on cyclically when that queue has no full chunk left
    //    (dragging the preference along).  Nobody ever waits for a slow ray, which wins when ray
    //    cost varies wildly (deep mesh BVHs) and the hot loop is small enough to stay cached.
#ifdef TB_WF2_LANEQ
    static_assert(MODE == WF2_MODE_GENERIC, "the lane-owned layout serves the free-running scheduler only");
    constexpr bool hard = false, split = false;
    const int cycle = 0;
    for (;;) {
        int s = 0, stage = -1;
        const bool fromAnswer = false;
        const uint4 ans0 = make_uint4(0u, 0u, 0u, 0u), ans1 = ans0, ans2 = ans0;
        const unsigned act = wf2_claim_fullest(S, WF2_STICK_LANES, s, stage);
        if (act == 0u) {
            if (*(volatile int*)&S.live <= 0) break;
            __nanosleep(200);
            continue;
        }
        const bool active = ((act >> lane) & 1u) != 0u;
#else
    const bool hard = MODE == WF2_MODE_HARD ? true : (MODE == WF2_MODE_SPLIT || offload) ? false : (P.hardPhases != 0);
    constexpr bool split = MODE == WF2_MODE_SPLIT;
    // hard-phase schedule: each cycle is two block-synchronous phases,
    //   phase 0:  R (finished slots -> fresh camera rays into F[cycle&1]),  T,  F[~cycle&1]
    //   phase 1:  A,  B
    // Inside a phase the stages run back to back without a barrier: they consume queues that were
    // completed before the phase began (so their tails are fixed and tickets cannot overshoot into
    // missing entries) and touch disjoint slots.  step 0..4 = R, T, F, A, B.
    int step = 0, cycle = 0;
    unsigned int stepTail = TB_WF2_PATHS;   // every slot starts in the R queue
    int stepQueue = WF2_Q_R;

    unsigned int idleSpins = 0;
    for (;;) {
        int s = 0, n = 0, stage = -1;
        bool fromAnswer = false;      // offload mode: the chunk are walkers' answers, merged at the head of stage A / B
        uint4 ans0 = make_uint4(0u, 0u, 0u, 0u), ans1 = ans0, ans2 = ans0;
        if (hard) {
            n = wf2_claim_ticket(S, stepQueue, stepTail, s);
            if (n > 0) {
                stage = stepQueue;
            } else {
                // this stage's queue is drained; at the end of a phase meet the other warps, undo the
                // overshoot of the failed ticket requests (head back to the tail the stage ran with)
                const bool endOfPhase = (step == 2 || step == 4);
                if (endOfPhase) {
                    __syncthreads();
                    if (tid == 0) {
                        if (step == 2) {
                            S.head[WF2_Q_R] = S.snap[0];
                            S.head[WF2_Q_T] = S.snap[1];
                            S.head[WF2_Q_F0 + ((cycle + 1) & 1)] = S.snap[2];
                        } else {
                            S.head[WF2_Q_A] = S.snap[3];
                            S.head[WF2_Q_B] = S.snap[4];
                        }
                    }
                    if (step == 2 && *(volatile int*)&S.live <= 0) {
                        // `live` only changes during R; nothing is left anywhere once it reaches zero
                        break;
                    }
                }
                if (step == 4) {
                    step = 0;
                    ++cycle;
                } else {
                    ++step;
                }
                stepQueue = step == 0 ? WF2_Q_R : step == 1 ? WF2_Q_T : step == 2 ? WF2_Q_F0 + ((cycle + 1) & 1)
                          : step == 3 ? WF2_Q_A : WF2_Q_B;
                // the queue was completed before this phase began; thread 0 publishes the tail every
                // warp of the CTA uses (and that the head is reset to afterwards)
                stepTail = *(volatile unsigned int*)&S.tail[stepQueue];
                if (tid == 0) S.snap[step] = stepTail;
                continue;
            }
        } else {
            int p = 0;
            if (lane == 0) p = *(volatile int*)&S.pref | (*(volatile int*)&S.exhausted << 8);
            p = __shfl_sync(0xffffffffu, p, 0);
            // Once the sample counter has run dry the CTA is draining: nothing refills the queues any more, so
            // waiting for full 32-entry chunks only adds latency -- take whatever is there.  (A 1-spp Render() on
            // one of N GPUs is ALL drain: a single partial fill of the slot arrays.)
#ifdef TB_WF2_NO_DRAIN
            const int full = 32;
#else
            const int full = (p >> 8) ? 1 : 32;
#endif
            p &= 0xff;
            if (offload) {
                // sweep over six sources: T, answers(ext), A, answers(shadow), B, R
                for (int k = 0; k < WALK_SWEEP_PASSES * 6 && stage < 0; ++k) {
                    const int pos = (p + k) % 6;
                    const int minCount = k < 6 ? full : (k < 12 ? WALK_PARTIAL_MIN : 1);
                    if (pos == 1 || pos == 3) {
                        n = wf2_claim_answers(S, P.walk, pos == 1 ? WALK_KIND_EXT : WALK_KIND_SHADOW, minCount, ans0, ans1, ans2);
                        if (n > 0) {
                            stage = pos == 1 ? WF2_Q_A : WF2_Q_B;
                            fromAnswer = true;
                            s = (int)(ans2.z & WF2_SLOT_MASK);
                        }
                    } else {
                        const int q = pos == 0 ? WF2_Q_T : pos == 2 ? WF2_Q_A : pos == 4 ? WF2_Q_B : WF2_Q_R;
                        n = wf2_claim(S, q, minCount, s);
                        if (n > 0) stage = q;
                    }
                    if (n > 0 && pos != p && lane == 0) *(volatile int*)&S.pref = pos;
                }
            } else if (!split) {
                for (int k = 0; k < 8 && stage < 0; ++k) {
                    // k = 0..3: full chunks only; k = 4..7: whatever is left
                    const int q = (p + k) & 3;
                    n = wf2_claim(S, q, k < 4 ? full : 1, s);
                    if (n > 0) stage = q;
                }
            } else {
                // same sweep over five queues: T, TM, A, B, R (pref holds the position in this order)
                for (int k = 0; k < 10 && stage < 0; ++k) {
                    const int pos = (p + k) % 5;
                    const int q = pos == 0 ? WF2_Q_T : pos == 1 ? WF2_Q_TM : pos == 2 ? WF2_Q_A : pos == 3 ? WF2_Q_B : WF2_Q_R;
                    n = wf2_claim(S, q, k < 5 ? full : 1, s);
                    if (n > 0) {
                        stage = q;
                        if (pos != p && lane == 0) *(volatile int*)&S.pref = pos;
                    }
                }
            }
            if (!split && !offload && stage >= 0 && stage != p && lane == 0) *(volatile int*)&S.pref = stage;
            if (stage < 0) {
                if (*(volatile int*)&S.live <= 0) break;
                if (offload) {
                    // parked slots wait for the walkers: a watchdog instead of an endless spin if they never answer
                    if ((++idleSpins & 255u) == 0u) {
                        unsigned int bad = 0;
                        if (lane == 0) bad = walk_ld1(P.walk.abortFlag);
                        if (__shfl_sync(0xffffffffu, bad, 0)) break;
                        if (idleSpins > WALK_WATCHDOG_SPINS) {
                            if (lane == 0) atomicExch(P.walk.abortFlag, 1u);
                            break;
                        }
                    }
                }
                __nanosleep(200);
                continue;
            }
            idleSpins = 0;
        }
        const bool active = lane < n;
#endif

        if (stage == WF2_Q_R) {
            // ===================== R: splat the finished sample, regenerate =======================
            float4* rec = coldBase + (size_t)s * CC_CHUNKS;
            const bool finished = active && !(S.flags[s] & WF2_FLAG_EMPTY);
            uint32_t doneSample = 0xffffffffu;
            if (finished) {
                const float4 cl = cold_ld(rec, CC_L), ca = cold_ld(rec, CC_A);
                doneSample = f2u(ca.w);
                int px, py, frame;
                decode_sample(P, (unsigned long long)doneSample, px, py, frame);
                // raster position of the sample: its first two RNG draws (render.cpp:476,481-482)
                Rng rr = rng_seed(tb_sample_seed((uint32_t)(py * P.film.width + px), (uint32_t)frame));
                float rx = rng_float(rr);
                float ry = rng_float(rr);
                rx += px;
                ry += py;
                sample_end(P, px, py, rx, ry, v3(cl.x, cl.y, cl.z));
            }
            if (P.bandFlags)
                wf2_band_report(P.bandCount, P.bandFlags, P.bandSamples, (uint32_t)P.samplesPerFrame, P.bandTag, doneSample, finished);
            if (active) S.flags[s] = WF2_FLAG_EMPTY;
            // claim a new sample; indices on tile padding outside the image are skipped
            bool want = active && !*(volatile int*)&S.exhausted;
            bool fresh = false;
            for (int attempt = 0; attempt < 64; ++attempt) {
                if (!__any_sync(0xffffffffu, want)) break;
                const unsigned m = __ballot_sync(0xffffffffu, want);
                const int leader = __ffs(m) - 1;
                unsigned long long base = 0ull;
                if (lane == leader) {
                    base = atomicAdd(P.sampleCounter, (unsigned long long)__popc(m));
                    if (base + __popc(m) >= total) *(volatile int*)&S.exhausted = 1;
                }
                base = __shfl_sync(0xffffffffu, base, leader);
                const unsigned long long idx = base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
                if (want) {
                    if (idx >= total) {
                        want = false;
                    } else {
                        int px, py, frame;
                        if (decode_sample(P, idx, px, py, frame)) {
                            PathState ps;
                            float rx, ry;
                            sample_begin(P, px, py, frame, ps, rx, ry);
                            S.ox[s] = ps.o.x; S.oy[s] = ps.o.y; S.oz[s] = ps.o.z;
                            S.dx[s] = ps.d.x; S.dy[s] = ps.d.y; S.dz[s] = ps.d.z;
                            S.time[s] = ps.time;
                            cold_st(rec, CC_T, 1.0f, 1.0f, 1.0f, 1.0f);                     // throughput 1, eta 1
                            cold_st(rec, CC_L, 0.0f, 0.0f, 0.0f, 1.0f);                     // radiance 0, bsdfPdf 1
                            cold_st(rec, CC_A, 0.0f, 0.0f, 0.0f, u2f((uint32_t)idx));      // no absorption, the sample index
                            cold_st(rec, CC_RNG, u2f(ps.rng.s1), u2f(ps.rng.s2), 0.0f, 0.0f);
                            S.flags[s] = ((uint32_t)TB_REFLECTED << 1) | ((uint32_t)WF2_PH_EXT << 3);
                            fresh = true;
                            want = false;
                        } else if (P.bandFlags) {
                            // tile padding outside the image retires at once
                            wf2_band_pad(P.bandCount, P.bandFlags, P.bandSamples, (uint32_t)P.samplesPerFrame, P.bandTag, (uint32_t)idx);
                        }
                    }
                }
            }
            __threadfence_block();
            wf2_push_trace(S, sc, split, hard ? WF2_Q_F0 + (cycle & 1) : WF2_Q_T, fresh, s);
            // slots that could not be refilled die: the CTA exits when none is left
            const unsigned dead = __ballot_sync(0xffffffffu, active && !fresh);
            if (dead && lane == 0) atomicSub(&S.live, __popc(dead));
        } else if (stage == WF2_Q_T || stage >= WF2_Q_F0) {
            // ===================== T: trace the pending ray =======================================
            bool isExt = false, isNee = false;
            uint32_t parkMask = 0u;       // offload mode: big meshes the ray still has to be walked through
            V3 postO = v3s(0.0f), postD = v3s(0.0f);
            float postTime = 0.0f;
            if (active) {
                const uint32_t fl = S.flags[s];
                V3 o = v3(S.ox[s], S.oy[s], S.oz[s]);
                V3 d = v3(S.dx[s], S.dy[s], S.dz[s]);
                const float time = S.time[s];
                isExt = ((fl >> 3) & 1u) == WF2_PH_EXT;
                isNee = !isExt;
                if (isNee) {
                    // shadow ray from the surface point: origin = p + FaceForward(n, wi)*eps (render.cpp:121,170)
                    const V3 p = o + d * S.ht[s];
                    const V3 nn = v3(S.hnx[s], S.hny[s], S.hnz[s]);
                    d = v3(S.sdx[s], S.sdy[s], S.sdz[s]);
                    o = p + face_forward(nn, d) * TB_RAY_EPS;
                }
                if (isNee || maxDepth > 0) {
                    // one traversal instance serves both ray kinds (normals only for extension rays)
                    const Hit h = offload ? trace_partial(sc, o, d, time, isExt, parkMask) : trace_closest(sc, o, d, time, isExt);
                    if (offload) {
                        postO = o;
                        postD = d;
                        postTime = time;
                    }
                    if (isExt) {
                        S.ht[s] = h.t;
                        S.hnx[s] = h.n.x; S.hny[s] = h.n.y; S.hnz[s] = h.n.z;
                        S.hprim[s] = h.prim;
                    } else {
                        S.st[s] = h.t;
                        S.sprim[s] = h.prim;
                    }
                } else {
                    S.hprim[s] = -2;   // maxDepth == 0: no trace at all, radiance stays 0
                }
            }
            __threadfence_block();
            if (offload) {
                // rays that reach a big mesh park in their slot; the walkers' answer re-enters at stage A / B
                const bool park = parkMask != 0u;
                const unsigned pe = __ballot_sync(0xffffffffu, park && isExt), pn = __ballot_sync(0xffffffffu, park && isNee);
                if (lane == 0) {
                    if (pe) atomicAdd(&S.ansPending[WALK_KIND_EXT], __popc(pe));
                    if (pn) atomicAdd(&S.ansPending[WALK_KIND_SHADOW], __popc(pn));
                }
                walk_post(P.walk, park, postO, postD, postTime, parkMask,
                          (uint32_t)s | ((isExt ? WALK_KIND_EXT : WALK_KIND_SHADOW) << WF2_SLOT_BITS) | ((uint32_t)blockIdx.x << (WF2_SLOT_BITS + 1)));
                isExt = isExt && !park;
                isNee = isNee && !park;
            }
            wf2_push(S, WF2_Q_A, isExt, s);
            wf2_push(S, WF2_Q_B, isNee, s);
        } else if (stage == WF2_Q_A) {
            // ===================== A: extension-ray results =======================================
            bool cont = false, fin = false;
            if (offload && fromAnswer && active) wf2_merge_answer(S, sc, s, true, ans0, ans1, ans2);
            if (active) {
                float4* rec = coldBase + (size_t)s * CC_CHUNKS;
                const float4 ct = cold_ld(rec, CC_T), cl = cold_ld(rec, CC_L);
                const uint32_t fl = S.flags[s];
                const int bounce = (int)(fl >> 8);
                const int hprim = S.hprim[s];
                if (hprim < 0) {
                    if (hprim == -1) {
                        PathState ps;
                        ps.d = v3(S.dx[s], S.dy[s], S.dz[s]);
                        ps.T = v3(ct.x, ct.y, ct.z);
                        ps.L = v3(cl.x, cl.y, cl.z);
                        ps.rayType = (int)((fl >> 1) & 3u);
                        ps.bsdfPdf = cl.w;
                        path_miss(sc, ps, bounce);
                        cold_st(rec, CC_L, ps.L.x, ps.L.y, ps.L.z, cl.w);
                    }
                    fin = true;
                } else {
                    // hit prologue (render.cpp:255-310)
                    const float4 ca = cold_ld(rec, CC_A), cr = cold_ld(rec, CC_RNG);
                    PathState ps;
                    ps.o = v3(S.ox[s], S.oy[s], S.oz[s]);
                    ps.d = v3(S.dx[s], S.dy[s], S.dz[s]);
                    ps.time = S.time[s];
                    ps.T = v3(ct.x, ct.y, ct.z);
                    ps.L = v3(cl.x, cl.y, cl.z);
                    ps.eta = ct.w;
                    ps.absorb = v3(ca.x, ca.y, ca.z);
                    ps.rayType = (int)((fl >> 1) & 3u);
                    ps.bsdfPdf = cl.w;
                    ps.rng.s1 = f2u(cr.x);
                    ps.rng.s2 = f2u(cr.y);
                    Hit h;
                    h.t = S.ht[s];
                    h.n = v3(S.hnx[s], S.hny[s], S.hnz[s]);
                    h.prim = hprim;
                    Surface sf;
                    path_hit(sc, ps, h, bounce, sf);
                    NeeCursor c;
                    nee_begin(c);
                    ShadowRay sr;
                    if (nee_generate(sc, sf, ps.time, c, ps.rng, sr)) {
                        cold_st(rec, CC_T, ps.T.x, ps.T.y, ps.T.z, ct.w);
                        cold_st(rec, CC_L, ps.L.x, ps.L.y, ps.L.z, cl.w);
                        wf2_store_shadow(S, rec, s, sr, c, ps.rng);
                        S.flags[s] = (fl & ~(1u << 3)) | ((uint32_t)WF2_PH_NEE << 3);
                        cont = true;
                    } else {
                        // scene without lights or probe: straight to the BSDF
                        cont = wf2_scatter(S, rec, sc, s, sf, ps.T, ps.L, ps.rng, ps.time, c.sum, bounce, maxDepth, ps.absorb, ps.bsdfPdf, f2u(ca.w));
                        fin = !cont;
                    }
                }
            }
            __threadfence_block();
            wf2_push_trace(S, sc, split, WF2_Q_T, cont, s);
            wf2_push(S, WF2_Q_R, fin, s);
        } else {
            // ===================== B: shadow-ray results ==========================================
            bool cont = false, fin = false;
            if (offload && fromAnswer && active) wf2_merge_answer(S, sc, s, false, ans0, ans1, ans2);
            if (active) {
                float4* rec = coldBase + (size_t)s * CC_CHUNKS;
                const float4 ct = cold_ld(rec, CC_T), cl = cold_ld(rec, CC_L), ca = cold_ld(rec, CC_A), cr = cold_ld(rec, CC_RNG);
                const float4 csh = cold_ld(rec, CC_SH), csum = cold_ld(rec, CC_SUM), cla = cold_ld(rec, CC_LA);
                const uint32_t fl = S.flags[s];
                const int bounce = (int)(fl >> 8);
                const Surface sf = wf2_surface(S, sc, s, ct.w);
                ShadowRay sr;
                sr.o = v3s(0.0f);
                sr.d = v3(S.sdx[s], S.sdy[s], S.sdz[s]);
                sr.dist = csh.x;
                sr.lightN = v3(csh.y, csh.z, csh.w);
                sr.skyPdf = csum.x;
                sr.light = __float_as_int(cr.w);
                NeeCursor c;
                const uint32_t cw = f2u(cr.z);
                c.slot = (int)(cw & 0xffu);
                c.prim = (int)((cw >> 8) & 0xfffu);
                c.sample = (int)(cw >> 20);
                c.sum = v3(csum.y, csum.z, csum.w);
                c.Lacc = v3(cla.x, cla.y, cla.z);
                Hit sh;
                sh.t = S.st[s];
                sh.prim = S.sprim[s];
                sh.n = v3s(0.0f);
                nee_connect(sc, sf, sr, sh, c);

                Rng rng;
                rng.s1 = f2u(cr.x);
                rng.s2 = f2u(cr.y);
                const float time = S.time[s];
                if (nee_generate(sc, sf, time, c, rng, sr)) {
                    wf2_store_shadow(S, rec, s, sr, c, rng);
                    cont = true;
                } else {
                    cont = wf2_scatter(S, rec, sc, s, sf, v3(ct.x, ct.y, ct.z), v3(cl.x, cl.y, cl.z), rng, time, c.sum, bounce, maxDepth,
                                       v3(ca.x, ca.y, ca.z), cl.w, f2u(ca.w));
                    fin = !cont;
                }
            }
            __threadfence_block();
            wf2_push_trace(S, sc, split, WF2_Q_T, cont, s);
            wf2_push(S, WF2_Q_R, fin, s);
        }
    }
    if (offload) {
        // this CTA posts no more requests: tell the walkers (once)
        if (lane == 0 && atomicExch(&S.exitSignaled, 1) == 0) {
            __threadfence();
            atomicAdd(P.walk.shadersDone, 1u);
        }
    }
}

// The opt-in to > 48 KB of dynamic shared memory and the occupancy answer are PER DEVICE (and the
// attribute applies to the current device only), so they are kept per device ordinal and set up once
// per device under a lock: a host may own renderers on several GPUs, driven from several threads.
#define TB_WF2_MAX_DEVICES 64
template <int THREADS, int MODE>
static int wavefront2_ctas_per_sm()
{
    static std::mutex lock;
    static int ctasPerSM[TB_WF2_MAX_DEVICES] = {};   // 0 = not configured yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TB_WF2_MAX_DEVICES) dev = 0;
    std::lock_guard<std::mutex> guard(lock);
    if (ctasPerSM[dev] == 0) {
        int n = 0;
        cudaFuncSetAttribute(k_wavefront2<THREADS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)std::max(wavefront2_smem_limit(), (size_t)TB_WF2_SMEM_BYTES));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront2<THREADS, MODE>, THREADS, TB_WF2_SMEM_BYTES) != cudaSuccess || n < 1)
            n = 1;
        if (n > TB_WF2_MAX_CTAS_PER_SM) n = TB_WF2_MAX_CTAS_PER_SM;
        ctasPerSM[dev] = n;
    }
    return ctasPerSM[dev];
}

template <int THREADS, int MODE>
static void launch_wavefront2_t(const LaunchParams& p0, int numSMs, cudaStream_t stream, unsigned long long total)
{
    const int ctasPerSM = wavefront2_ctas_per_sm<THREADS, MODE>();
    // shared memory beyond the slot arrays holds the treelet (the shader role; walker CTAs lay out their own)
    LaunchParams p = p0;
    const size_t slotBytes = (TB_WF2_SMEM_BYTES + 127) & ~size_t(127);
    p.treeletBytes = 0;
    if (p.scene.treeletMesh >= 0 && p.scene.treeletPairs > 0 && MODE != WF2_MODE_OFFLOAD && wavefront2_smem_limit() > slotBytes) {
        const size_t room = (wavefront2_smem_limit() - slotBytes) / sizeof(BvhPair) * sizeof(BvhPair);
        p.treeletBytes = (int)std::min(room, (size_t)p.scene.treeletPairs * sizeof(BvhPair));
    }
    const size_t smemBytes = p.treeletBytes > 0 ? slotBytes + (size_t)p.treeletBytes : TB_WF2_SMEM_BYTES;
    // TB_WF2_CTAS_PER_SM resident CTAs per SM; small jobs use fewer so that every CTA has a full slot array
    const unsigned long long want = (total + TB_WF2_PATHS - 1) / TB_WF2_PATHS;
    int grid = (numSMs > 0 ? numSMs : 148) * ctasPerSM;
    if (MODE == WF2_MODE_OFFLOAD) {
        // shader CTAs [0, numShaders) + walker CTAs behind them, all resident at once (they wait for each other)
        LaunchParams q = p;
        int walkers = std::max(1, std::min(q.walk.numWalkers, grid - 1));
        int shaders = grid - walkers;
        if (want < (unsigned long long)shaders) shaders = (int)std::max<unsigned long long>(1ull, want);
        q.walk.numShaders = shaders;
        q.walk.numWalkers = walkers;
        k_wavefront2<THREADS, MODE><<<shaders + walkers, THREADS, smemBytes, stream>>>(q, total);
        return;
    }
    if (want < (unsigned long long)grid) grid = (int)want;
    if (grid < 1) grid = 1;
    k_wavefront2<THREADS, MODE><<<grid, THREADS, smemBytes, stream>>>(p, total);
}

// the launches of this layout (see the top of the file for which scheduler variants each layout serves)
void launch_layout(const LaunchParams& p, int numSMs, cudaStream_t stream, unsigned long long* launchCount)
{
    const unsigned long long total = p.samplesPerFrame * (unsigned long long)p.numFrames;
    if (total == 0ull) return;
    cudaMemsetAsync(p.sampleCounter, 0, sizeof(unsigned long long), stream);
#if !defined(TB_WF2_LANEQ)
    const bool split = !p.hardPhases && p.scene.splitValid;
#endif
#if defined(TB_WF2_LANEQ)
    launch_wavefront2_t<TB_WF2_THREADS, WF2_MODE_GENERIC>(p, numSMs, stream, total);
#elif defined(TB_WF2_COLD_SMEM)
    if (p.hardPhases) launch_wavefront2_t<TB_WF2_THREADS, WF2_MODE_HARD>(p, numSMs, stream, total);
    else if (split) launch_wavefront2_t<TB_WF2_THREADS, WF2_MODE_SPLIT>(p, numSMs, stream, total);
    else launch_wavefront2_t<TB_WF2_THREADS, WF2_MODE_GENERIC>(p, numSMs, stream, total);
#else
    if (wavefront2_wants_offload(p)) {
        if (p.wideCta) launch_wavefront2_t<TB_WF2_THREADS_WIDE, WF2_MODE_OFFLOAD>(p, numSMs, stream, total);
        else launch_wavefront2_t<TB_WF2_THREADS, WF2_MODE_OFFLOAD>(p, numSMs, stream, total);
    } else if (split) {
        launch_wavefront2_t<TB_WF2_THREADS_WIDE, WF2_MODE_SPLIT>(p, numSMs, stream, total);
    } else {
        launch_wavefront2_t<TB_WF2_THREADS_WIDE, WF2_MODE_GENERIC>(p, numSMs, stream, total);   // hard phases too (run-time flag)
    }
#endif
    if (launchCount) ++*launchCount;
}
