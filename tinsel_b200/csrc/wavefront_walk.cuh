// wavefront_walk.cuh -- the mesh-walk offload of the wavefront kernel (k_wavefront2, MODE_OFFLOAD).
//
// Scenes with a big triangle mesh spend most of their time in the mesh BVH walk
// (IntersectRayMesh, intersection.h:678-749): dependent 64-byte node fetches from L2, rays whose
// cost differs by 100x, a third of the lanes busy.  Run on the same SM as the shading stages it also
// fights them for the 32 KB instruction cache (profiles/README.md, steps 8 and 17).  In this mode
// the launch splits its CTAs -- one per SM -- into two roles:
//
//   shader CTAs   the usual stage machine (R, T, A, B) over 1024 path slots in shared memory.
//                 Stage T evaluates the scene program for everything but the big meshes; a ray
//                 that enters a big mesh's box is POSTED to the walk queue and its slot parks until
//                 the answer comes back.  Stages A / B merge the answer with the partial hit.
//   walker CTAs   nothing but the mesh walk: a few hundred instructions that own their SM's
//                 instruction cache, every lane refilled from the queue the moment its ray is done,
//                 the top of the mesh BVH staged once per CTA into shared memory with a bulk-copy
//                 (TMA, cp.async.bulk -> UBLKCP) so that only the lower levels go to L2.
//
// Queues live in global memory (L2): 48-byte records written as three 16-byte chunks that each
// carry the lap tag of their ring cell in their last word, so a reader validates every chunk by
// itself -- no fences, no flags, one round trip:
//   request ring   one for the launch, MPMC.  Producers (shader warps) reserve with one atomicAdd on
//                  `tail` per warp; consumers (walker warps) take tickets with one atomicAdd on
//                  `head` per refill and poll their cell until its tag matches.  The ring holds
//                  more cells than there are path slots on the device and every slot has at most
//                  one request in flight, so a producer can never lap an unconsumed cell.
//   answer rings   two per shader CTA (extension rays -> stage A, shadow rays -> stage B), 1024
//                  cells each.  Walker lanes reserve with an atomicAdd on the ring's tail; the
//                  CTA's warps claim prefixes of valid cells with a CAS on a head kept in shared memory.
// Counters run on across launches (tags are idx >> log2(capacity) + 1, never 0 = the cleared state).
//
// Exactness: the walker visits nodes and triangles in the reference's order (near child first,
// `t < tmax` culling, strict `t < closestT`), so a mesh's own result is the reference's bit for
// bit; across primitives the closest hit is order-independent except for exact ties in t, which
// are detected and redone with the reference-order walk (trace_ordered), as the scene program does.
// (no include guard: compiled once per slot-state layout, see wavefront2.cuh)

// second pass of a shader warp's sweep over its work sources takes chunks of at least this many entries
// (the first pass wants full chunks of 32, a third pass takes anything)
#ifndef WALK_PARTIAL_MIN
#define WALK_PARTIAL_MIN 1
#endif
#if WALK_PARTIAL_MIN > 1
#define WALK_SWEEP_PASSES 3
#else
#define WALK_SWEEP_PASSES 2
#endif
#ifndef WALK_BOX_MIN
#define WALK_BOX_MIN 20       // the BOX phase repeats while at least this many lanes have an interior node
#endif

#define WALK_KIND_EXT 0
#define WALK_KIND_SHADOW 1
// spin budget of the watchdogs (iterations of a ~0.5 us sleep): a launch that makes no progress for
// seconds is a bug, and must not take the GPU with it
#define WALK_WATCHDOG_SPINS (48u << 20)

TB_DEV uint4 walk_ld(const uint4* p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
TB_DEV void walk_st(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
TB_DEV unsigned int walk_ld1(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- producer side (shader CTAs, stage T) ---------------------------------------------------------
// `src` = slot | kind << WF2_SLOT_BITS | shader CTA << (WF2_SLOT_BITS + 1)
TB_DEV void walk_post(const WalkParams& W, bool flag, V3 o, V3 d, float time, uint32_t primMask, uint32_t src)
{
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(W.reqTail, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (flag) {
        const unsigned int idx = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        const unsigned int tag = (idx >> W.reqLog2) + 1u;
        uint4* cell = W.reqRing + (size_t)(idx & ((1u << W.reqLog2) - 1u)) * 3;
        walk_st(cell + 0, __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), tag);
        walk_st(cell + 1, __float_as_uint(d.x), __float_as_uint(d.y), __float_as_uint(d.z), tag);
        walk_st(cell + 2, __float_as_uint(time), primMask, src, tag);
    }
}

// ---- the walker role ------------------------------------------------------------------------------
struct WalkHit {
    float t, u, v, w;
    V3 gn;
    int tri, prim;
};

TB_DEV uint32_t walk_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// a generic address the compiler cannot trace back to shared memory: loads through it are generic loads
// (LD), so one instruction serves lanes that read the staged treelet and lanes that read global memory
TB_DEV const unsigned char* walk_opaque(const void* p)
{
    unsigned long long a = (unsigned long long)p;
    asm volatile("" : "+l"(a));
    return reinterpret_cast<const unsigned char*>(a);
}

#ifdef WALK_FAST_SLAB
// IntersectRayAABBFast (intersection.h:373-397) with hardware min/max (FMNMX) instead of the reference's
// `a < b ? a : b` selects.  The two differ only when a product is NaN (0 * inf: a ray parallel to a slab
// AND starting exactly on one of its planes) -- signed zeros compare equal everywhere the result is
// used -- so `exact` comes back false for those lanes and the caller redoes the test the slow way.
TB_DEV bool ray_aabb_fast(V3 pos, V3 rcp, float lx, float ly, float lz, float ux, float uy, float uz, float& t, bool& exact)
{
    const float a1 = (lx - pos.x) * rcp.x, a2 = (ux - pos.x) * rcp.x;
    const float b1 = (ly - pos.y) * rcp.y, b2 = (uy - pos.y) * rcp.y;
    const float c1 = (lz - pos.z) * rcp.z, c2 = (uz - pos.z) * rcp.z;
    const float lmin = fmaxf(fmaxf(fminf(a1, a2), fminf(b1, b2)), fminf(c1, c2));
    const float lmax = fminf(fminf(fmaxf(a1, a2), fmaxf(b1, b2)), fmaxf(c1, c2));
    const float sum = ((a1 + a2) + (b1 + b2)) + (c1 + c2);   // NaN if any product is (and for inf - inf: a harmless false alarm)
    exact = sum == sum;
    t = lmin;
    return (lmax >= 0.f) & (lmax >= lmin);
}
#endif

// One CTA of walkers.  Shared memory: [0,16) mbarrier, then the traversal stacks (32 entries per
// thread, entry-major so that a warp's accesses fall into 32 different banks), then the treelet.
template <int THREADS>
static __device__ void wf2_walker_role(const LaunchParams& P, unsigned char* smem, int smemBytes)
{
    const WalkParams& W = P.walk;
    const DScene& sc = P.scene;
    const int tid = threadIdx.x, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* stackBase = reinterpret_cast<uint32_t*>(smem + 16);
    BvhPair* treelet = reinterpret_cast<BvhPair*>(smem + 16 + (size_t)TB_STACK * THREADS * 4);
    const int treeletCap = (smemBytes - 16 - TB_STACK * THREADS * 4) / (int)sizeof(BvhPair);
    const int treeletPairs = W.treeletMesh >= 0 ? min(W.treeletPairs, treeletCap) : 0;
    const BvhPair* treeletSrc = nullptr;

    // ---- stage the top of the big mesh's BVH: one bulk copy per 32 KB, completion on an mbarrier --
    if (treeletPairs > 0) {
        treeletSrc = sc.meshes[W.treeletMesh].pairs;
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(walk_smem_addr(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)treeletPairs * (uint32_t)sizeof(BvhPair);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(walk_smem_addr(bar)), "r"(bytes) : "memory");
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = min(32768u, bytes - off);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 walk_smem_addr(reinterpret_cast<unsigned char*>(treelet) + off)),
                             "l"(reinterpret_cast<const unsigned char*>(treeletSrc) + off), "r"(n), "r"(walk_smem_addr(bar))
                             : "memory");
            }
        }
        __syncthreads();   // the barrier word is initialised before anybody polls it
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok)
                         : "r"(walk_smem_addr(bar)), "r"(0u)
                         : "memory");
    }
    uint32_t* stack = stackBase + tid;   // entry k of this thread: stack[k * THREADS]
    const unsigned char* treeletBytes = walk_opaque(treelet);

    // ---- per-lane state -------------------------------------------------------------------------------
    enum { PH_IDLE = 0, PH_TICKET = 1, PH_WALK = 2, PH_RETIRED = 3 };
    int phase = PH_IDLE;
    unsigned int ticket = 0;
    V3 o = v3s(0.0f), d = v3s(0.0f), rcp = v3s(0.0f);   // world ray until the first primitive is set up, then the mesh-space ray
    float rtime = 0.0f, tmax = FLT_MAX;
    uint32_t maskLeft = 0, src = 0;
    int prim = -1, meshId = -1;
    const BvhPair* pairs = nullptr;
    const float4* triVerts = nullptr;
    const uint32_t EMPTY = 0xffffffffu;
    uint32_t cur = EMPTY;         // node to visit next (TB_LEAF: a triangle); EMPTY: this mesh is finished / not begun
    int sp = 0;
    WalkHit best;                 // closest over the request's meshes so far
    best.t = FLT_MAX; best.prim = -1; best.tri = -1; best.u = best.v = best.w = 0.0f; best.gn = v3s(0.0f);
    WalkHit mh;                   // closest within the mesh being walked
    mh.t = FLT_MAX; mh.tri = -1; mh.prim = -1; mh.u = mh.v = mh.w = 0.0f; mh.gn = v3s(0.0f);
    bool tie = false;
    unsigned int idleSpins = 0;
    int fillWait = 0;             // passes until FILL may run again after a fruitless poll

    const unsigned int reqMask = (1u << W.reqLog2) - 1u;

    uint32_t pend = EMPTY;        // a triangle found but not yet tested (speculative traversal, see below)

    // The warp cycles through four phases; in each, the lanes that are ready for it work and the others
    // wait -- never two divergent branches in one pass, which is what a ray-per-lane loop with
    // data-dependent branches would execute:
    //   BOX   interior-node steps, repeated while most lanes have one.  A lane that reaches a triangle
    //         parks it in `pend` and keeps descending (speculatively: its later box tests use a stale,
    //         larger tmax, so it visits a superset of the reference's nodes; everything in the extra nodes
    //         lies at t >= the closest hit and fails the strict `t < closestT`, and triangles are still
    //         tested in the reference's order -- the result is the reference's, bit for bit)
    //   TRI   the parked triangles are tested
    //   NEXT  lanes whose mesh is finished fold its hit and set up the request's next mesh, or answer
    //   FILL  idle lanes take tickets (one atomic for the warp), ticket holders poll their request cell
    // NEXT and FILL run as soon as a handful of lanes want them: they are what puts lanes back to work.
    for (;;) {
        // ---- BOX -----------------------------------------------------------------------------------------
        int nBox;
        for (;;) {
            const bool wBox = phase == PH_WALK && cur != EMPTY && (cur & TB_LEAF) == 0u;
            nBox = __popc(__ballot_sync(0xffffffffu, wBox));
            if (nBox == 0) break;
            if (wBox) {
                // IntersectRayMesh interior step, intersection.h:702-727
                const unsigned char* recBase = (meshId == W.treeletMesh && cur < (uint32_t)treeletPairs)
                                                   ? treeletBytes : reinterpret_cast<const unsigned char*>(pairs);
                const BvhPair* pr = reinterpret_cast<const BvhPair*>(recBase + (size_t)cur * sizeof(BvhPair));
                const float4 a = pr->a, b = pr->b, c = pr->c;
                const uint2 kids = *reinterpret_cast<const uint2*>(&pr->left);
                float tLeft, tRight;
#ifdef WALK_FAST_SLAB
                bool exL, exR;
                bool hitLeft = ray_aabb_fast(o, rcp, a.x, a.y, a.z, a.w, b.x, b.y, tLeft, exL);
                bool hitRight = ray_aabb_fast(o, rcp, b.z, b.w, c.x, c.y, c.z, c.w, tRight, exR);
                if (!(exL && exR)) {
                    hitLeft = ray_aabb(o, rcp, a.x, a.y, a.z, a.w, b.x, b.y, tLeft);
                    hitRight = ray_aabb(o, rcp, b.z, b.w, c.x, c.y, c.z, c.w, tRight);
                }
                hitLeft = hitLeft && tLeft < tmax;
                hitRight = hitRight && tRight < tmax;
#else
                const bool hitLeft = ray_aabb(o, rcp, a.x, a.y, a.z, a.w, b.x, b.y, tLeft) && tLeft < tmax;
                const bool hitRight = ray_aabb(o, rcp, b.z, b.w, c.x, c.y, c.z, c.w, tRight) && tRight < tmax;
#endif
                // "traverse closest first": the reference pushes the far child, then the near one, and pops the near
                // one at once (intersection.h:716-727) -- the near child is `nxt`, the far one goes on the stack
                const bool both = hitLeft && hitRight;
                const bool swap = both && (tLeft < tRight);
                const uint32_t far = swap ? kids.y : kids.x;              // what the reference pushes first
                const uint32_t near = both ? (swap ? kids.x : kids.y) : (hitLeft ? kids.x : kids.y);
                const bool any = hitLeft || hitRight;
                // Stack traffic without divergent branches.  What this step can do: push `far` (both hit); pop
                // (nothing hit); park a triangle in `pend` and pop what the reference would visit after testing
                // it.  At most one push and two pops, of which the first pop can only return `far` itself when
                // there was a push -- so memory sees one store (the entry that survives) and two loads.
                const uint32_t top1 = sp > 0 ? stack[(sp - 1) * THREADS] : EMPTY;
                const uint32_t top2 = sp > 1 ? stack[(sp - 2) * THREADS] : EMPTY;
                // value sequence of pops after this step's push: p1, p2
                const uint32_t p1 = both ? far : top1;
                const uint32_t p2 = both ? top1 : top2;
                uint32_t nxt = any ? near : p1;          // nothing hit: pop
                int pops = any ? 0 : 1;
                const bool park = nxt != EMPTY && (nxt & TB_LEAF) != 0u && pend == EMPTY;
                if (park) {
                    // park the triangle, go on with what the reference would pop after testing it
                    pend = nxt;
                    nxt = pops == 0 ? p1 : p2;
                    pops += 1;
                }
                // net effect on the stack: +1 (push) - pops, never below zero (a pop from an empty stack returned EMPTY)
                const int depthAfterPush = sp + (both ? 1 : 0);
                const int newSp = depthAfterPush - pops < 0 ? 0 : depthAfterPush - pops;
                if (both && newSp > sp && sp < TB_STACK) stack[sp * THREADS] = far;   // the pushed entry survives this step
                sp = newSp;
                cur = nxt;
            }
            if (nBox < 20) break;   // give the other phases a turn (one step per turn keeps every lane progressing)
        }

        // ---- TRI -----------------------------------------------------------------------------------------
        {
            const int nPend = __popc(__ballot_sync(0xffffffffu, phase == PH_WALK && pend != EMPTY));
            if (nPend > 0 && (nPend >= 8 || nBox < 16)) {
                if (phase == PH_WALK && pend != EMPTY) {
                    // MeshQuery, intersection.h:629-674
                    const uint32_t i = pend & ~TB_LEAF;
                    const float4 q0 = __ldg(&triVerts[i * 3 + 0]);
                    const float4 q1 = __ldg(&triVerts[i * 3 + 1]);
                    const float4 q2 = __ldg(&triVerts[i * 3 + 2]);
                    float t, u, v, w, sign;
                    V3 n;
                    if (ray_tri(o, d, v3(q0.x, q0.y, q0.z), v3(q0.w, q1.x, q1.y), v3(q1.z, q1.w, q2.x), t, u, v, w, sign, n)) {
                        if (t > 0.0f && t < mh.t) {
                            mh.t = t;
                            mh.u = u;
                            mh.v = v;
                            mh.w = w;
                            mh.tri = (int)i;
                            mh.gn = n * sign;
                        }
                    }
                    tmax = mh.t;   // "truncate ray", intersection.h:700
                    pend = EMPTY;
                    if (cur != EMPTY && (cur & TB_LEAF) != 0u) {
                        // a second triangle was waiting behind the first
                        pend = cur;
                        cur = EMPTY;
                        if (sp > 0) {
                            sp -= 1;
                            cur = stack[sp * THREADS];
                        }
                    }
                }
            }
        }

        // ---- NEXT ----------------------------------------------------------------------------------------
        {
            const bool wNext = phase == PH_WALK && cur == EMPTY && pend == EMPTY;
            const int nNext = __popc(__ballot_sync(0xffffffffu, wNext));
            if (nNext > 0 && (nNext >= 4 || nBox < 8)) {
                if (wNext) {
                    // fold the finished mesh's hit: "t < minT && t > 0", first found wins (render.cpp:45) -- the
                    // reference meets the primitives in BVH order, this loop in index order: only an exact tie in
                    // t can tell the difference, and that is reported
                    if (prim >= 0 && mh.tri >= 0) {
                        if (mh.t < best.t) {
                            best = mh;
                            best.prim = prim;
                        } else if (mh.t == best.t) {
                            tie = true;
                        }
                    }
                    if (maskLeft != 0u) {
                        V3 ow = o, dw = d;
                        if (prim >= 0) {
                            // a second mesh in the same request: the world ray is still in its (unconsumable) ring cell
                            const uint4 c0 = walk_ld(W.reqRing + (size_t)(ticket & reqMask) * 3 + 0);
                            const uint4 c1 = walk_ld(W.reqRing + (size_t)(ticket & reqMask) * 3 + 1);
                            ow = v3(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z));
                            dw = v3(__uint_as_float(c1.x), __uint_as_float(c1.y), __uint_as_float(c1.z));
                        }
                        prim = __ffs(maskLeft) - 1;
                        maskLeft &= maskLeft - 1u;
                        const DPrim& p = sc.prims[prim];
                        // PrimitiveIntersect, mesh case (intersection.h:982-992): the ray in the mesh's space
                        const Xf xf = prim_transform(p, rtime);
                        o = inverse_transform_point(xf, ow);
                        d = inverse_transform_vector(xf, dw);
                        rcp.x = 1.0f / d.x;
                        rcp.y = 1.0f / d.y;
                        rcp.z = 1.0f / d.z;
                        meshId = p.mesh;
                        const DMesh& m = sc.meshes[meshId];
                        pairs = m.pairs;
                        triVerts = m.triVerts;
                        cur = m.rootRef;
                        if ((cur & TB_LEAF) != 0u) {   // a one-triangle mesh: its root is the triangle
                            pend = cur;
                            cur = EMPTY;
                        }
                        sp = 0;
                        tmax = FLT_MAX;
                        mh.t = FLT_MAX;
                        mh.tri = -1;
                    } else {
                        // the request is finished: answer it
                        const unsigned int cta = src >> (WF2_SLOT_BITS + 1), kind = (src >> WF2_SLOT_BITS) & 1u, slot = src & WF2_SLOT_MASK;
                        const unsigned int idx = atomicAdd(W.ansTail + cta * 2 + kind, 1u);
                        const unsigned int tag = (idx >> WF2_LOG2_PATHS) + 1u;
                        uint4* cell = W.ansRing + ((size_t)(cta * 2 + kind) * TB_WF2_PATHS + (idx & WF2_MASK)) * 3;
                        const uint32_t info = slot | ((uint32_t)(best.prim & 0xff) << WF2_SLOT_BITS) | ((best.prim >= 0 ? 1u : 0u) << (WF2_SLOT_BITS + 8)) |
                                              ((tie ? 1u : 0u) << (WF2_SLOT_BITS + 9));
                        walk_st(cell + 0, __float_as_uint(best.t), __float_as_uint(best.u), __float_as_uint(best.v), tag);
                        walk_st(cell + 1, __float_as_uint(best.w), __float_as_uint(best.gn.x), __float_as_uint(best.gn.y), tag);
                        walk_st(cell + 2, __float_as_uint(best.gn.z), (uint32_t)best.tri, info, tag);
                        phase = PH_IDLE;
                    }
                }
            }
        }

        // ---- FILL ----------------------------------------------------------------------------------------
        if (fillWait > 0) fillWait -= 1;
        const int nWalk = __popc(__ballot_sync(0xffffffffu, phase == PH_WALK));
        const int nFill = __popc(__ballot_sync(0xffffffffu, phase == PH_IDLE || phase == PH_TICKET));
        if (nFill > 0 && (nWalk == 0 || (fillWait == 0 && (nFill >= 4 || nBox < 8)))) {
            const unsigned idle = __ballot_sync(0xffffffffu, phase == PH_IDLE);
            if (idle != 0u) {
                const int leader = __ffs(idle) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(W.reqHead, (unsigned)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (phase == PH_IDLE) {
                    ticket = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                    phase = PH_TICKET;
                }
            }
            bool got = false;
            if (phase == PH_TICKET) {
                const uint4* cell = W.reqRing + (size_t)(ticket & reqMask) * 3;
                const unsigned int tag = (ticket >> W.reqLog2) + 1u;
                const uint4 c0 = walk_ld(cell + 0), c1 = walk_ld(cell + 1), c2 = walk_ld(cell + 2);
                if (c0.w == tag && c1.w == tag && c2.w == tag) {
                    o = v3(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z));   // world-space ray
                    d = v3(__uint_as_float(c1.x), __uint_as_float(c1.y), __uint_as_float(c1.z));
                    rtime = __uint_as_float(c2.x);
                    maskLeft = c2.y;
                    src = c2.z;
                    best.t = FLT_MAX; best.prim = -1; best.tri = -1;
                    tie = false;
                    prim = -1;
                    cur = EMPTY;          // NEXT sets up the first mesh of the mask
                    pend = EMPTY;
                    phase = PH_WALK;
                    got = true;
                }
            }
            if (__any_sync(0xffffffffu, got)) {
                fillWait = 0;
                idleSpins = 0;
            } else {
                fillWait = 16;
                if (nWalk == 0) {
                    // nothing to walk in this warp: wait for requests, or leave when the shaders are done.  The
                    // order matters: every ticket below the final tail was filled before the last shader left.
                    unsigned int done = 0, bad = 0;
                    if (lane == 0) {
                        done = walk_ld1(W.shadersDone);
                        bad = walk_ld1(W.abortFlag);
                    }
                    done = __shfl_sync(0xffffffffu, done, 0);
                    bad = __shfl_sync(0xffffffffu, bad, 0);
                    if (bad) break;
                    if (done >= (unsigned)W.numShaders) {
                        // re-poll once after seeing the flag; whoever still has no request will never get one
                        bool late = false;
                        if (phase == PH_TICKET) {
                            const uint4* cell = W.reqRing + (size_t)(ticket & reqMask) * 3;
                            const unsigned int tag = (ticket >> W.reqLog2) + 1u;
                            const uint4 c0 = walk_ld(cell + 0), c1 = walk_ld(cell + 1), c2 = walk_ld(cell + 2);
                            late = c0.w == tag && c1.w == tag && c2.w == tag;
                            if (!late) phase = PH_RETIRED;
                        }
                        if (!__any_sync(0xffffffffu, late)) break;
                        fillWait = 0;
                        continue;   // the next FILL picks the late request up
                    }
                    if (++idleSpins > WALK_WATCHDOG_SPINS) {
                        if (lane == 0) atomicExch(W.abortFlag, 2u);
                        break;
                    }
                    __nanosleep(400);
                }
            }
        } else if (nWalk == 0 && nFill == 0) {
            break;   // every lane of the warp has retired
        }
    }
    // the last walker out squares the request ring for the next launch: tickets taken beyond the final
    // tail were never filled, so head := tail; and the per-launch counters go back to zero
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(W.walkersDone, 1u) + 1u == (unsigned)W.numWalkers) {
            *(volatile unsigned int*)W.reqHead = walk_ld1(W.reqTail);
            *(volatile unsigned int*)W.shadersDone = 0u;
            *(volatile unsigned int*)W.walkersDone = 0u;
            __threadfence();
        }
    }
}
