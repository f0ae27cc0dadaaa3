// api.cu -- host side of the C ABI (include/tinsel_b200.h): scene translation into the GPU
// layout, accumulator management, frame sequencing, host read-back.
//
// Host arithmetic here restates reference *host-side* or *per-primitive-constant* expressions
// (cited inline) and is compiled without contraction (-Xcompiler -ffp-contract=off, no fast-math)
// so that it matches the reference's x86-64 build.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tb_kernels.cuh"

const char* tb200_snapshot_error();   // snapshot.cpp
unsigned long long tb200_snapshot_error_stamp();
unsigned long long tb200_error_tick();

namespace {

thread_local std::string g_error;
thread_local unsigned long long g_errorStamp = 0;

bool set_error(const std::string& what)
{
    g_error = what;
    g_errorStamp = tb200_error_tick();
    fprintf(stderr, "[tinsel_b200] %s\n", what.c_str());
    return false;
}

#define TB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                              \
            return false;                                                                               \
        }                                                                                               \
    } while (0)

template <typename T>
bool upload(const std::vector<T>& host, T** dev, uint64_t* h2d, size_t padElems = 0)
{
    *dev = nullptr;
    const size_t n = host.size() + padElems;
    if (n == 0) return true;
    TB_CUDA(cudaMalloc((void**)dev, n * sizeof(T)));
    if (padElems) TB_CUDA(cudaMemset(*dev, 0, n * sizeof(T)));
    if (!host.empty()) TB_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *h2d += host.size() * sizeof(T);
    return true;
}

V3 hv3(const float* f) { return v3(f[0], f[1], f[2]); }

Xf to_xf(const tb200_transform& t)
{
    Xf x;
    x.p = hv3(t.p);
    x.r = q4(t.r[0], t.r[1], t.r[2], t.r[3]);
    x.s = t.s;
    return x;
}

// Per-material constants.  The double intermediates are the reference's own:
//   GetIndexOfRefraction  scene.h:72-78    2.0f/(1.0f-sqrtf(0.08*specular)) - 1.0f
//   Cdlum                 disney.h:307     .3*r + .6*g + .1*b   (double, rounded to float)
//   Ctint, Cspec0         disney.h:309-310 specular*.08 is a double product converted to Real
//   clearcoat alpha       disney.h:387     Lerp(.1,.001,gloss) in double, converted to float
DMaterial make_material(const tb200_material& m)
{
    DMaterial d;
    d.emission = hv3(m.emission);
    d.color = hv3(m.color);
    d.absorption = hv3(m.absorption);
    d.metallic = m.metallic;
    d.subsurface = m.subsurface;
    d.specular = m.specular;
    d.roughness = m.roughness;
    d.specularTint = m.specularTint;
    d.clearcoat = m.clearcoat;
    d.clearcoatGloss = m.clearcoatGloss;
    d.transmission = m.transmission;

    if (m.eta == 0.0f)
        d.ior = 2.0f / (1.0f - sqrtf(float(0.08 * double(m.specular)))) - 1.0f;
    else
        d.ior = m.eta;

    const float Cdlum = float(.3 * double(m.color[0]) + .6 * double(m.color[1]) + .1 * double(m.color[2]));
    const V3 Cdlin = d.color;
    const V3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : v3s(1.0f);
    const float spec08 = float(double(m.specular) * .08);
    d.cspec0 = tb_lerp(spec08 * tb_lerp(v3s(1.0f), Ctint, m.specularTint), Cdlin, m.metallic);

    d.sqrtColor = v3(sqrtf(m.color[0]), sqrtf(m.color[1]), sqrtf(m.color[2]));
    d.alpha = tb_max(0.001f, m.roughness);

    const float a = float(.1 + (.001 - .1) * double(m.clearcoatGloss));
    d.gtr1Wide = (a >= 1) ? 1 : 0;
    const float a2 = a * a;
    d.gtr1A2m1 = a2 - 1;
    d.gtr1PiLogA2 = TB_PI * logf(a2);
    return d;
}

// BVHNode[] (bvh.h:9-19) -> BvhPair[]; see tb_scene.cuh.  Returns the root reference.  Pairs are laid out
// in breadth-first order from the root, so that the first K records are the top of the tree: the walker
// CTAs of the offload mode stage exactly that prefix in shared memory with one bulk copy
// (wavefront_walk.cuh), and the upper levels share cache lines for everybody else.  Record indices are
// private to this layout; topology, child order and therefore visit order are the reference's.
// *depth (optional) receives the number of levels of interior nodes.
uint32_t build_pairs(const tb200_bvh_node* nodes, int numNodes, std::vector<BvhPair>* out, int* depth = nullptr)
{
    out->clear();
    if (depth) *depth = 0;
    if (numNodes <= 0) return TB_LEAF;   // never traversed: callers skip empty trees
    if (nodes[0].right_leaf >> 31) return TB_LEAF | nodes[0].left;
    std::vector<uint32_t> pairIndex(numNodes, 0xffffffffu);
    std::vector<uint32_t> order;   // interior nodes in breadth-first order
    order.reserve(numNodes / 2 + 1);
    pairIndex[0] = 0;
    order.push_back(0);
    size_t levelEnd = 1;
    int levels = 1;
    for (size_t k = 0; k < order.size(); ++k) {
        if (k == levelEnd) {
            levelEnd = order.size();
            levels += 1;
        }
        const tb200_bvh_node& n = nodes[order[k]];
        const uint32_t kids[2] = {n.left, n.right_leaf & 0x7fffffffu};
        for (uint32_t c : kids) {
            if ((nodes[c].right_leaf >> 31) || pairIndex[c] != 0xffffffffu) continue;   // leaf, or already placed (malformed input)
            pairIndex[c] = (uint32_t)order.size();
            order.push_back(c);
        }
    }
    if (depth) *depth = levels;
    out->resize(order.size());
    auto ref_of = [&](uint32_t idx) -> uint32_t {
        const tb200_bvh_node& n = nodes[idx];
        return (n.right_leaf >> 31) ? (TB_LEAF | n.left) : pairIndex[idx];
    };
    for (size_t k = 0; k < order.size(); ++k) {
        const tb200_bvh_node& n = nodes[order[k]];
        const uint32_t li = n.left, ri = n.right_leaf & 0x7fffffffu;
        const tb200_bvh_node& L = nodes[li];
        const tb200_bvh_node& R = nodes[ri];
        BvhPair p;
        p.a = make_float4(L.lower[0], L.lower[1], L.lower[2], L.upper[0]);
        p.b = make_float4(L.upper[1], L.upper[2], R.lower[0], R.lower[1]);
        p.c = make_float4(R.lower[2], R.upper[0], R.upper[1], R.upper[2]);
        p.left = ref_of(li);
        p.right = ref_of(ri);
        p.pad0 = p.pad1 = 0;
        (*out)[k] = p;
    }
    return ref_of(0);
}

// Flat scene program for trace_closest() (tb_scene.cuh).  Walks the scene BVH once; infinite
// boxes (planes' +-1e8 bounds and every ancestor of one) are transparent, finite interior boxes get
// a bit in the per-ray mask, leaves become LEAF / LEAFBOX / PLANE ops guarded by the bit of their
// nearest finite ancestor.  Built only for scenes of at most 16 primitives.
void build_program(const tb200_scene* sc, std::vector<ProgOp>* out)
{
    out->clear();
    const tb200_bvh_node* nodes = sc->bvhNodes;
    if (sc->numBvhNodes <= 0 || sc->numPrimitives > 16) return;
    struct Item { uint32_t node; int guard; bool isRoot; };
    std::vector<Item> todo;
    todo.push_back({0u, TB_BIT_ALWAYS, true});
    int nextBit = 0;
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        const tb200_bvh_node& n = nodes[it.node];
        const bool leaf = (n.right_leaf >> 31) != 0;
        bool infinite = true;
        for (int k = 0; k < 3; ++k) infinite = infinite && n.lower[k] <= -5.0e7f && n.upper[k] >= 5.0e7f;
        if (it.isRoot) infinite = true;   // the root's own box is never tested (intersection.h:759-763)
        ProgOp op;
        memset(&op, 0, sizeof(op));
        op.a[0] = n.lower[0]; op.a[1] = n.lower[1]; op.a[2] = n.lower[2]; op.a[3] = n.upper[0];
        op.b[0] = n.upper[1]; op.b[1] = n.upper[2];
        if (leaf) {
            const tb200_primitive& p = sc->primitives[n.left];
            int kind = infinite ? TB_OP_LEAF : TB_OP_LEAFBOX;
            if (infinite && p.type == TB200_PLANE) {
                kind = TB_OP_PLANE;
                memcpy(op.a, p.plane, 16);
            }
            op.kindPrim = kind | (int(n.left) << 8);
            op.bits = it.guard;
            out->push_back(op);
        } else {
            int guard = it.guard;
            if (!infinite) {
                if (nextBit >= 31) { out->clear(); return; }
                op.kindPrim = TB_OP_BOX;
                op.bits = it.guard | (nextBit << 8);
                out->push_back(op);
                guard = nextBit++;
            }
            todo.push_back({n.right_leaf & 0x7fffffffu, guard, false});
            todo.push_back({n.left, guard, false});
        }
    }
    if (out->size() > 32) out->clear();
}

// CameraSampler constructor, util.h:49-71, with Mat44(Transform) (maths.h:841-849), Mat33(Quat)
// (maths.h:658-667) and MatrixMultiply<4,4,4> (maths.h:86-101: t = 0; t += a*b for k = 0..3).
struct M44 {
    float c[4][4];   // column major: c[col][row]
};

M44 mat_mul(const M44& a, const M44& b)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float t = 0.0f;
            for (int k = 0; k < 4; ++k) t += a.c[k][i] * b.c[j][k];
            r.c[j][i] = t;
        }
    return r;
}

M44 mat_rows(float m11, float m12, float m13, float m14, float m21, float m22, float m23, float m24, float m31, float m32,
             float m33, float m34, float m41, float m42, float m43, float m44)
{
    M44 m;
    m.c[0][0] = m11; m.c[0][1] = m21; m.c[0][2] = m31; m.c[0][3] = m41;
    m.c[1][0] = m12; m.c[1][1] = m22; m.c[1][2] = m32; m.c[1][3] = m42;
    m.c[2][0] = m13; m.c[2][1] = m23; m.c[2][2] = m33; m.c[2][3] = m43;
    m.c[3][0] = m14; m.c[3][1] = m24; m.c[3][2] = m34; m.c[3][3] = m44;
    return m;
}

void camera_setup(const tb200_camera& cam, int width, int height, DCamera* out)
{
    Xf t;
    t.p = hv3(cam.position);
    t.r = q4(cam.rotation[0], cam.rotation[1], cam.rotation[2], cam.rotation[3]);
    t.s = 1.0f;   // Transform(camera.position, camera.rotation), render.cpp:450-452

    M44 c2w;
    const V3 c0 = rotate(t.r, v3(1.0f, 0.0f, 0.0f)) * t.s;
    const V3 c1 = rotate(t.r, v3(0.0f, 1.0f, 0.0f)) * t.s;
    const V3 c2 = rotate(t.r, v3(0.0f, 0.0f, 1.0f)) * t.s;
    const V3 c3 = t.p * t.s;
    c2w.c[0][0] = c0.x; c2w.c[0][1] = c0.y; c2w.c[0][2] = c0.z; c2w.c[0][3] = 0.0f;
    c2w.c[1][0] = c1.x; c2w.c[1][1] = c1.y; c2w.c[1][2] = c1.z; c2w.c[1][3] = 0.0f;
    c2w.c[2][0] = c2.x; c2w.c[2][1] = c2.y; c2w.c[2][2] = c2.z; c2w.c[2][3] = 0.0f;
    c2w.c[3][0] = c3.x; c2w.c[3][1] = c3.y; c2w.c[3][2] = c3.z; c2w.c[3][3] = 1.0f;

    const M44 rasterToScreen = mat_rows(2.0f / width, 0.0f, 0.0f, -1.0f,
                                        0.0f, -2.0f / height, 0.0f, 1.0f,
                                        0.0f, 0.0f, 1.0f, 1.0f,
                                        0.0f, 0.0f, 0.0f, 1.0f);
    const float f = tanf(cam.fov * 0.5f);
    const float aspect = float(width) / height;
    const M44 screenToCamera = mat_rows(f * aspect, 0.0f, 0.0f, 0.0f,
                                        0.0f, f, 0.0f, 0.0f,
                                        0.0f, 0.0f, -1.0f, 0.0f,
                                        0.0f, 0.0f, 0.0f, 1.0f);
    const M44 r2w = mat_mul(mat_mul(c2w, screenToCamera), rasterToScreen);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out->r2w[c * 4 + r] = r2w.c[c][r];
    out->origin = v3(c2w.c[3][0], c2w.c[3][1], c2w.c[3][2]);
    out->shutterStart = cam.shutterStart;
    out->shutterEnd = cam.shutterEnd;
}

}  // namespace

struct tb200_renderer {
    int device = 0;
    int numSMs = 0;
    cudaStream_t stream = nullptr;        // the stream work is launched on
    cudaStream_t ownStream = nullptr;     // the renderer's private stream (default for `stream`)
    int shard = 0, numShards = 1;
    float4* boundAccum = nullptr;         // caller-owned accumulator (tb200_bind_accumulator)
    cudaEvent_t evStart = nullptr, evStop = nullptr;

    // device scene
    DScene scene;
    DPrim* dPrims = nullptr;
    BvhPair* dScenePairs = nullptr;
    ProgOp* dFlat = nullptr;
    DMesh* dMeshes = nullptr;
    std::vector<void*> owned;   // every other device allocation

    // film
    int width = 0, height = 0;
    float4* dAccum = nullptr;
    float* dRadiance = nullptr;   // tb200_trace_frame scratch
    float* dRaster = nullptr;
    unsigned long long* dCounter = nullptr;
    float4* dCold = nullptr;      // cold half of the wavefront's slot state (wavefront2.cuh): 128 B per slot
    int frame = 0;

    // display/finish step (tb200_finish)
    float4* dFiltered = nullptr;
    unsigned char* dRgb8 = nullptr;
    uint2* dDither = nullptr;
    float4* dMeans = nullptr;                // tb200_nlm scratch + output
    float4* dDenoised = nullptr;
    int finishWidth = 0, finishHeight = 0;   // size the buffers above were made for
    bool ditherReady = false;
    bool filteredValid = false;              // dFiltered holds the image of a tb200_finish since the last tb200_init

    // streamed read-back (tb200_render): band counters on the device, completion flags in mapped
    // host memory, a private copy stream
    unsigned int* dBandCount = nullptr;
    volatile unsigned int* hBandFlags = nullptr;
    unsigned int* dBandFlags = nullptr;   // device alias of hBandFlags
    cudaStream_t copyStream = nullptr;
    unsigned int bandTag = 0;
    int streamedReadback = 1;             // TINSEL_B200_READBACK=plain turns it off

    // host read-back: a caller-owned buffer pinned on request (tb200_pin_output), never implicitly
    void* registered = nullptr;
    size_t registeredBytes = 0;

    // owner-computes row slab (tb200_set_slab; slabRows < 0: the whole image)
    int slabRow0 = 0, slabRows = -1;

    // multi-device group (tb200_create_multi): the head owns one renderer per further device, each
    // driven by its own host thread
    std::vector<tb200_renderer*> peers;
    struct Worker* worker = nullptr;      // peers only
    std::string workerError;              // a peer's failure message, handed to the calling thread
    bool peerAccess = false;

    // mesh-walk offload (wavefront_walk.cuh): queues in device memory, walker CTAs per launch
    WalkParams walk;
    void* dWalkBlock = nullptr;   // one allocation behind every pointer of `walk`
    size_t walkBlockBytes = 0;
    int numWalkers = 0;           // 0: the offload mode is off

    int pipeline = 2;             // 0 = mega (validation), 2 = wavefront (product)
    int hardPhases = 1;           // wavefront scheduling mode (see wavefront2.cuh)
    int wideCta = 0;              // 768-thread CTAs for deep mesh BVHs (see wavefront2.cuh)
    int laneQueues = 0;           // lane-owned slots + bit-set queues (see wavefront2.cuh)
    // Split trace queue on/off is decided by measurement (every variant produces the same bits): the
    // second and third sizeable launches after tb200_create time one setting each, the faster stays.
    int tunePhase = 3;            // 0 warm-up (split), 1 measuring split, 2 measuring no split, 3 decided
    double tuneRate[2] = {0.0, 0.0};   // samples per ms with / without the split queue
    unsigned long long launchSamples = 0;   // samples of the launch whose time is in stats.gpuMs
    tb200_stats stats;
};

namespace {

void free_device(tb200_renderer* r)
{
    if (r->registered) {
        cudaHostUnregister(r->registered);
        r->registered = nullptr;
    }
    for (void* p : r->owned) cudaFree(p);
    r->owned.clear();
    cudaFree(r->dPrims);
    cudaFree(r->dScenePairs);
    cudaFree(r->dFlat);
    r->dFlat = nullptr;
    cudaFree(r->dMeshes);
    cudaFree(r->dAccum);
    cudaFree(r->dRadiance);
    cudaFree(r->dRaster);
    cudaFree(r->dCounter);
    r->dCounter = nullptr;
    cudaFree(r->dCold);
    r->dCold = nullptr;
    cudaFree(r->dBandCount);
    r->dBandCount = nullptr;
    cudaFree(r->dWalkBlock);
    r->dWalkBlock = nullptr;
    r->numWalkers = 0;
    cudaFree(r->dFiltered);
    cudaFree(r->dRgb8);
    cudaFree(r->dDither);
    cudaFree(r->dMeans);
    cudaFree(r->dDenoised);
    r->dFiltered = nullptr;
    r->dRgb8 = nullptr;
    r->dDither = nullptr;
    r->dMeans = r->dDenoised = nullptr;
    r->finishWidth = r->finishHeight = 0;
    r->ditherReady = false;
    r->filteredValid = false;
    if (r->hBandFlags) cudaFreeHost((void*)r->hBandFlags);
    r->hBandFlags = nullptr;
    r->dBandFlags = nullptr;
    r->dPrims = nullptr;
    r->dScenePairs = nullptr;
    r->dMeshes = nullptr;
    r->dAccum = nullptr;
    r->dRadiance = r->dRaster = nullptr;
}

// ---------------------------------------------------------------------------------------------------
// Scene image: everything build_scene() derives from a tb200_scene, in exactly the layouts the device
// holds -- DPrim records with their hoisted constants, child-pair BVH records in breadth-first order,
// pre-gathered triangles, the flat scene program, the probe with its four sampling tables.  It is
// built on the host (no GPU needed), uploaded array by array, and it is what the scene cache stores
// (tb200_scene_cache_save / tb200_create_cached): a cached scene goes from the file to the device
// without the loader, the BVH re-packing, the per-triangle gather or Probe::BuildCDF.
// ---------------------------------------------------------------------------------------------------
struct MeshImage {
    std::vector<BvhPair> pairs;
    std::vector<float4> verts, norms;
    std::vector<float> cdf;
    int numTris = 0;
    uint32_t rootRef = TB_LEAF;
    int depth = 0;
};

struct SceneImage {
    std::vector<DPrim> prims;
    std::vector<BvhPair> scenePairs;
    uint32_t sceneRoot = TB_LEAF;
    std::vector<ProgOp> flat;
    std::vector<MeshImage> meshes;
    V3 horizon = v3s(0.0f), zenith = v3s(0.0f);
    int numNee = 0;
    // scheduling hints: the largest mesh, and its primitive's world bounds (its leaf of the scene BVH)
    int maxTris = 0;
    int splitBoxValid = 0;
    V3 splitLo = v3s(0.0f), splitHi = v3s(0.0f);
    // probe (src/probe.h:9-86), each table with one element of slack (see below)
    int probeValid = 0, probeW = 0, probeH = 0;
    std::vector<float4> probeData;
    std::vector<float> pdfX, cdfX, pdfY, cdfY;
};

bool image_from_scene(const tb200_scene* s, SceneImage* img)
{
    // meshes: pairs + pre-gathered triangles
    img->meshes.resize(s->numMeshes);
    for (int m = 0; m < s->numMeshes; ++m) {
        const tb200_mesh& g = s->meshes[m];
        MeshImage& mi = img->meshes[m];
        mi.rootRef = build_pairs(g.nodes, g.numNodes, &mi.pairs, &mi.depth);
        const int numTris = g.numIndices / 3;
        mi.numTris = numTris;
        img->maxTris = std::max(img->maxTris, numTris);
        mi.verts.resize(size_t(numTris) * 3);
        mi.norms.resize(size_t(numTris) * 3);
        for (int t = 0; t < numTris; ++t) {
            const int i0 = g.indices[t * 3 + 0], i1 = g.indices[t * 3 + 1], i2 = g.indices[t * 3 + 2];
            const float* a = g.positions + size_t(i0) * 3;
            const float* b = g.positions + size_t(i1) * 3;
            const float* c = g.positions + size_t(i2) * 3;
            mi.verts[size_t(t) * 3 + 0] = make_float4(a[0], a[1], a[2], b[0]);
            mi.verts[size_t(t) * 3 + 1] = make_float4(b[1], b[2], c[0], c[1]);
            mi.verts[size_t(t) * 3 + 2] = make_float4(c[2], 0.0f, 0.0f, 0.0f);
            const float* n1 = g.normals + size_t(i0) * 3;
            const float* n2 = g.normals + size_t(i1) * 3;
            const float* n3 = g.normals + size_t(i2) * 3;
            mi.norms[size_t(t) * 3 + 0] = make_float4(n1[0], n1[1], n1[2], n2[0]);
            mi.norms[size_t(t) * 3 + 1] = make_float4(n2[1], n2[2], n3[0], n3[1]);
            mi.norms[size_t(t) * 3 + 2] = make_float4(n3[2], 0.0f, 0.0f, 0.0f);
        }
        mi.cdf.assign(size_t(numTris), 0.0f);   // only read when the mesh is sampled as a light
        if (g.cdf) memcpy(mi.cdf.data(), g.cdf, size_t(numTris) * sizeof(float));
    }
    // the largest mesh's primitive and its world bounds = that primitive's leaf in the scene BVH
    {
        int bigPrim = -1, bigTris = 0;
        for (int i = 0; i < s->numPrimitives; ++i) {
            const tb200_primitive& p = s->primitives[i];
            if (p.type == TB200_MESH && p.mesh >= 0 && p.mesh < s->numMeshes && s->meshes[p.mesh].numIndices / 3 > bigTris) {
                bigTris = s->meshes[p.mesh].numIndices / 3;
                bigPrim = i;
            }
        }
        for (int n = 0; n < s->numBvhNodes && bigPrim >= 0; ++n) {
            const tb200_bvh_node& node = s->bvhNodes[n];
            if ((node.right_leaf >> 31) != 0 && (int)node.left == bigPrim) {
                img->splitLo = v3(node.lower[0], node.lower[1], node.lower[2]);
                img->splitHi = v3(node.upper[0], node.upper[1], node.upper[2]);
                img->splitBoxValid = 1;
                break;
            }
        }
    }

    // primitives
    img->prims.resize(s->numPrimitives);
    img->numNee = s->sky.probeValid ? 1 : 0;
    for (int i = 0; i < s->numPrimitives; ++i) {
        const tb200_primitive& p = s->primitives[i];
        DPrim& d = img->prims[i];
        memset(&d, 0, sizeof(d));
        d.start = to_xf(p.start);
        d.end = to_xf(p.end);
        d.isStatic = memcmp(&p.start, &p.end, sizeof(tb200_transform)) == 0;
        // with start == end, Lerp(a,b,t) = a + (b-a)*t = a + 0*t is time independent
        d.fixed = interpolate_transform(d.start, d.end, 0.0f);
        d.type = p.type;
        d.radius = p.radius;
        memcpy(d.plane, p.plane, 16);
        d.mesh = p.mesh;
        d.lightSamples = p.lightSamples;
        if (p.type == TB200_MESH && (p.mesh < 0 || p.mesh >= s->numMeshes)) return set_error("primitive references a missing mesh");
        // PrimitiveArea, intersection.h:833-853
        if (p.type == TB200_SPHERE)
            d.area = 4.0f * TB_PI * p.radius * p.radius;
        else if (p.type == TB200_MESH)
            d.area = s->meshes[p.mesh].area * p.end.s;
        else
            d.area = 0.0f;
        d.mat = make_material(p.material);
        if (p.lightSamples > 0) img->numNee += p.lightSamples;
    }
    img->sceneRoot = build_pairs(s->bvhNodes, s->numBvhNodes, &img->scenePairs);
    build_program(s, &img->flat);
    img->horizon = hv3(s->sky.horizon);
    img->zenith = hv3(s->sky.zenith);
    if (s->sky.probeValid) {
        const size_t n = size_t(s->sky.probeWidth) * s->sky.probeHeight;
        img->probeValid = 1;
        img->probeW = s->sky.probeWidth;
        img->probeH = s->sky.probeHeight;
        // one element of slack after each table: ProbeSample's column search may land on
        // col == width (probe.h:217-220) and read one past the last row
        img->probeData.assign(n + 1, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        memcpy(img->probeData.data(), s->sky.probeData, n * 16);
        img->pdfX.assign(n + 1, 0.0f);
        img->cdfX.assign(n + 1, 0.0f);
        img->pdfY.assign(size_t(s->sky.probeHeight) + 1, 0.0f);
        img->cdfY.assign(size_t(s->sky.probeHeight) + 1, 0.0f);
        memcpy(img->pdfX.data(), s->sky.pdfValuesX, n * 4);
        memcpy(img->cdfX.data(), s->sky.cdfValuesX, n * 4);
        memcpy(img->pdfY.data(), s->sky.pdfValuesY, size_t(s->sky.probeHeight) * 4);
        memcpy(img->cdfY.data(), s->sky.cdfValuesY, size_t(s->sky.probeHeight) * 4);
    }
    return true;
}

// Mesh-walk offload (wavefront_walk.cuh): decided per scene.  Eligible: free-running scenes whose scene
// level runs as the flat program (<= 16 primitives) and that hold a mesh of more than 4096 triangles whose
// BVH is no deeper than the reference's own traversal stack (a deeper tree overflows `int stack[32]`,
// intersection.h:688, in the reference itself).  TINSEL_B200_OFFLOAD=0 turns it off, =1 offloads every
// mesh (tests); TINSEL_B200_WALKERS=n sets the number of walker CTAs.
bool setup_offload(tb200_renderer* r, const SceneImage& img)
{
    DScene& sc = r->scene;
    sc.deferMask = 0u;
    r->numWalkers = 0;
    memset(&r->walk, 0, sizeof(r->walk));
    r->walk.treeletMesh = -1;
    // Measured (profiles/README.md, round 2): with the walk offloaded ajax runs at 564 Msamples/s against 694
    // inline (env 1950 / 3220, meshlight 254 / 325), so the mode is OFF unless asked for:
    // TINSEL_B200_OFFLOAD=1 offloads every eligible mesh of more than 4096 triangles, =2 every mesh (tests).
    const char* env = getenv("TINSEL_B200_OFFLOAD");
    const int want = env ? atoi(env) : 0;
    const bool force = want == 2;
    if (want <= 0 || r->hardPhases || sc.numFlat <= 0 || img.prims.size() > 16) return true;
    int bigMesh = -1, bigTris = 0;
    for (size_t i = 0; i < img.prims.size(); ++i) {
        const DPrim& p = img.prims[i];
        if (p.type != TB200_MESH) continue;
        const MeshImage& mi = img.meshes[p.mesh];
        if ((mi.numTris > 4096 || force) && mi.depth <= TB_STACK) {
            sc.deferMask |= 1u << i;
            if (mi.numTris > bigTris) {
                bigTris = mi.numTris;
                bigMesh = p.mesh;
            }
        }
    }
    if (sc.deferMask == 0u) return true;

    // queues: a request ring with more cells than the device has path slots (one request per slot at most),
    // two answer rings of TB_WF2_SLOTS cells per shader CTA, a handful of counters
    const int ctas = std::max(1, r->numSMs);
    unsigned int log2 = 10;
    while ((1ull << log2) < (unsigned long long)ctas * TB_WF2_SLOTS) ++log2;
    const size_t reqBytes = (size_t(1) << log2) * 48, ansBytes = size_t(ctas) * 2 * TB_WF2_SLOTS * 48;
    const size_t ctrBytes = (size_t(ctas) * 2 + 16) * sizeof(unsigned int);
    r->walkBlockBytes = reqBytes + ansBytes + ctrBytes;
    TB_CUDA(cudaMalloc(&r->dWalkBlock, r->walkBlockBytes));
    TB_CUDA(cudaMemset(r->dWalkBlock, 0, r->walkBlockBytes));
    char* base = (char*)r->dWalkBlock;
    WalkParams& W = r->walk;
    W.reqRing = (uint4*)base;
    W.reqLog2 = log2;
    W.ansRing = (uint4*)(base + reqBytes);
    unsigned int* ctr = (unsigned int*)(base + reqBytes + ansBytes);
    W.reqTail = ctr + 0;
    W.reqHead = ctr + 1;
    W.shadersDone = ctr + 2;
    W.walkersDone = ctr + 3;
    W.abortFlag = ctr + 4;
    W.ansTail = ctr + 16;
    W.treeletMesh = bigMesh;
    W.treeletPairs = bigMesh >= 0 ? (int)img.meshes[bigMesh].pairs.size() : 0;
    const char* nw = getenv("TINSEL_B200_WALKERS");
    r->numWalkers = nw ? atoi(nw) : (ctas * 3) / 8;
    r->numWalkers = std::max(1, std::min(r->numWalkers, ctas - 1));
    r->tunePhase = 3;   // the split-queue tuner belongs to the inline walk
    return true;
}

// After a launch of the offload mode has been waited for: did a watchdog fire?  (It never should: the
// flag means shader and walker CTAs stopped waiting for each other after seconds without progress.)
bool check_offload(tb200_renderer* r)
{
    if (r->numWalkers <= 0 || !r->dWalkBlock) return true;
    unsigned int flag = 0;
    TB_CUDA(cudaMemcpyAsync(&flag, r->walk.abortFlag, sizeof(flag), cudaMemcpyDeviceToHost, r->stream));
    TB_CUDA(cudaStreamSynchronize(r->stream));
    if (flag == 0u) return true;
    cudaMemsetAsync(r->dWalkBlock, 0, r->walkBlockBytes, r->stream);   // queues and counters back to a clean state
    cudaStreamSynchronize(r->stream);
    return set_error(flag == 1u ? "mesh-walk offload: a shader CTA waited for the walkers without progress (watchdog); the frame is incomplete"
                                : "mesh-walk offload: a walker CTA waited for requests without progress (watchdog); the frame is incomplete");
}

// device upload of a scene image + the per-scene scheduling decisions (which read the environment)
bool upload_image(tb200_renderer* r, const SceneImage& img)
{
    uint64_t* h2d = &r->stats.h2dBytes;
    memset(&r->scene, 0, sizeof(r->scene));
    std::vector<DMesh> meshes(img.meshes.size());
    for (size_t m = 0; m < img.meshes.size(); ++m) {
        const MeshImage& mi = img.meshes[m];
        BvhPair* dPairs;
        float4 *dVerts, *dNorms;
        float* dCdf;
        if (!upload(mi.pairs, &dPairs, h2d) || !upload(mi.verts, &dVerts, h2d) || !upload(mi.norms, &dNorms, h2d) ||
            !upload(mi.cdf, &dCdf, h2d))
            return false;
        r->owned.push_back(dPairs);
        r->owned.push_back(dVerts);
        r->owned.push_back(dNorms);
        r->owned.push_back(dCdf);
        meshes[m].pairs = dPairs;
        meshes[m].triVerts = dVerts;
        meshes[m].triNormals = dNorms;
        meshes[m].cdf = dCdf;
        meshes[m].numTris = mi.numTris;
        meshes[m].rootRef = mi.rootRef;
    }
    if (!upload(meshes, &r->dMeshes, h2d)) return false;
    // Scheduling mode of the wavefront kernel (measured, profiles/README.md): block-synchronous
    // stages win when every surface hit spawns several shadow rays (the shade and trace stages then
    // alternate several times per bounce and the barrier keeps all warps on one stage's code);
    // free-running warps win otherwise, and by a wide margin when ray cost varies (deep mesh BVHs).
    // TINSEL_B200_SCHED=hard|free overrides.
    {
        r->hardPhases = (img.numNee > 1 && img.maxTris <= 4096) ? 1 : 0;
        const char* sched = getenv("TINSEL_B200_SCHED");
        if (sched && strcmp(sched, "hard") == 0) r->hardPhases = 1;
        if (sched && strcmp(sched, "free") == 0) r->hardPhases = 0;
        // CTA size: deep mesh BVHs make traversal latency bound (L2 hits), which more warps hide; on
        // scenes held in shared memory more warps only thrash the instruction cache.  TINSEL_B200_CTA=512|768.
        r->wideCta = (!r->hardPhases && img.maxTris > 4096) ? 1 : 0;
        // Split trace queue: rays entering the largest mesh's world bounds are queued apart from the
        // rest (scheduling only).
        r->scene.splitValid = 0;
        const char* split = getenv("TINSEL_B200_SPLIT");   // 0: never, 1: whenever the scene has a mesh (tests)
        if ((img.maxTris > 4096 || (split && atoi(split) == 1)) && img.splitBoxValid) {
            r->scene.splitLo = img.splitLo;
            r->scene.splitHi = img.splitHi;
            r->scene.splitValid = 1;
        }
        if (split && atoi(split) == 0) r->scene.splitValid = 0;
        r->tunePhase = (r->scene.splitValid && !split && !r->hardPhases) ? 0 : 3;   // an explicit TINSEL_B200_SPLIT is final
        const char* cta = getenv("TINSEL_B200_CTA");
        if (cta && atoi(cta) == 768) r->wideCta = 1;
        if (cta && atoi(cta) == 512) r->wideCta = 0;
        // Lane-owned slots with bit-set queues: scenes held on chip under the free-running scheduler, for launches
        // big enough to reach a steady state (kernels.cu: launch_wavefront2; every other combination runs the
        // ring queues).  TINSEL_B200_QUEUES=ring|lanes forces one or the other whatever the launch size.
        r->laneQueues = 1;
        const char* queues = getenv("TINSEL_B200_QUEUES");
        if (queues && strcmp(queues, "ring") == 0) r->laneQueues = 0;
        if (queues && strcmp(queues, "lanes") == 0) r->laneQueues = 2;
    }
    // one element of padding each: the kernels' bulk copies round their length up to 16 bytes
    if (!upload(img.prims, &r->dPrims, h2d, 1)) return false;
    if (!upload(img.scenePairs, &r->dScenePairs, h2d, 1)) return false;

    DScene& sc = r->scene;
    sc.prims = r->dPrims;
    sc.numPrims = (int)img.prims.size();
    sc.pairs = r->dScenePairs;
    sc.numPairs = (int)img.scenePairs.size();
    static const std::vector<ProgOp> noFlat;
    const std::vector<ProgOp>& flat = getenv("TINSEL_B200_NO_FLAT") ? noFlat : img.flat;
    if (!upload(flat, &r->dFlat, h2d, 1)) return false;
    sc.flat = r->dFlat;
    sc.numFlat = (int)flat.size();
    if (!setup_offload(r, img)) return false;
    // treelet: the top of the largest mesh's BVH staged in the shared memory the slot arrays leave, for the INLINE
    // mesh walk.  Measured (profiles/README.md round 2, step 29): slower on every scene -- ajax 629 vs 689, env
    // 2858 vs 3067, meshlight 371 vs 402 Msamples/s -- because the top levels already live in L1, and claiming the
    // last 45 KB of shared memory shrinks L1 from 60 to 28 KB.  Off unless TINSEL_B200_TREELET=1.  (The walker CTAs
    // of the offload mode always stage theirs: they have nothing else in shared memory.)
    sc.treelet = nullptr;
    sc.treeletMesh = -1;
    sc.treeletPairs = 0;
    {
        const char* tl = getenv("TINSEL_B200_TREELET");
        int big = -1;
        for (size_t m = 0; m < img.meshes.size(); ++m)
            if (img.meshes[m].numTris > 1024 && (big < 0 || img.meshes[m].numTris > img.meshes[big].numTris)) big = (int)m;
        if (big >= 0 && tl && atoi(tl) == 1) {
            sc.treeletMesh = big;
            sc.treeletPairs = (int)img.meshes[big].pairs.size();
        }
    }
    sc.rootRef = img.sceneRoot;
    sc.meshes = r->dMeshes;
    sc.numMeshes = (int)img.meshes.size();
    sc.horizon = img.horizon;
    sc.zenith = img.zenith;
    sc.numNee = img.numNee;
    if (img.probeValid) {
        float4* dData;
        float *dPdfX, *dCdfX, *dPdfY, *dCdfY;
        if (!upload(img.probeData, &dData, h2d) || !upload(img.pdfX, &dPdfX, h2d) || !upload(img.cdfX, &dCdfX, h2d) ||
            !upload(img.pdfY, &dPdfY, h2d) || !upload(img.cdfY, &dCdfY, h2d))
            return false;
        r->owned.push_back(dData);
        r->owned.push_back(dPdfX);
        r->owned.push_back(dCdfX);
        r->owned.push_back(dPdfY);
        r->owned.push_back(dCdfY);
        sc.probe.valid = 1;
        sc.probe.width = img.probeW;
        sc.probe.height = img.probeH;
        sc.probe.data = dData;
        sc.probe.pdfX = dPdfX;
        sc.probe.cdfX = dCdfX;
        sc.probe.pdfY = dPdfY;
        sc.probe.cdfY = dCdfY;
    }
    return true;
}

bool build_scene(tb200_renderer* r, const tb200_scene* s)
{
    SceneImage img;
    return image_from_scene(s, &img) && upload_image(r, img);
}

// ---- scene cache file ------------------------------------------------------------------------------
// "TB2CACHE", a layout stamp (the records are this build's device structs, not an interchange format:
// a stamp mismatch is refused and the caller falls back to tb200_create), then the SceneImage fields in
// declaration order; every array as a 64-bit count followed by its bytes.
const char kCacheMagic[8] = {'T', 'B', '2', 'C', 'A', 'C', 'H', 'E'};
const uint32_t kCacheVersion = 2;

struct CacheStamp {
    uint32_t version, sizeofPrim, sizeofPair, sizeofOp;
};

template <typename T>
bool put_vec(FILE* f, const std::vector<T>& v)
{
    const uint64_t n = v.size();
    return fwrite(&n, 8, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <typename T>
bool put_pod(FILE* f, const T& v) { return fwrite(&v, sizeof(T), 1, f) == 1; }

template <typename T>
bool get_vec(FILE* f, std::vector<T>* v, uint64_t fileBytes)
{
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1 || n > fileBytes / sizeof(T)) return false;   // a count can never exceed the file
    v->resize(size_t(n));
    return n == 0 || fread(v->data(), sizeof(T), size_t(n), f) == n;
}
template <typename T>
bool get_pod(FILE* f, T* v) { return fread(v, sizeof(T), 1, f) == 1; }

bool image_save(const SceneImage& img, const char* path)
{
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(std::string("tb200_scene_cache_save: cannot open ") + path);
    const CacheStamp stamp = {kCacheVersion, (uint32_t)sizeof(DPrim), (uint32_t)sizeof(BvhPair), (uint32_t)sizeof(ProgOp)};
    bool ok = fwrite(kCacheMagic, 1, 8, f) == 8 && put_pod(f, stamp);
    ok = ok && put_vec(f, img.prims) && put_vec(f, img.scenePairs) && put_pod(f, img.sceneRoot) && put_vec(f, img.flat);
    const uint64_t numMeshes = img.meshes.size();
    ok = ok && put_pod(f, numMeshes);
    for (const MeshImage& m : img.meshes)
        ok = ok && put_vec(f, m.pairs) && put_vec(f, m.verts) && put_vec(f, m.norms) && put_vec(f, m.cdf) && put_pod(f, m.numTris) &&
             put_pod(f, m.rootRef) && put_pod(f, m.depth);
    ok = ok && put_pod(f, img.horizon) && put_pod(f, img.zenith) && put_pod(f, img.numNee) && put_pod(f, img.maxTris) &&
         put_pod(f, img.splitBoxValid) && put_pod(f, img.splitLo) && put_pod(f, img.splitHi);
    ok = ok && put_pod(f, img.probeValid) && put_pod(f, img.probeW) && put_pod(f, img.probeH) && put_vec(f, img.probeData) &&
         put_vec(f, img.pdfX) && put_vec(f, img.cdfX) && put_vec(f, img.pdfY) && put_vec(f, img.cdfY);
    ok = (fclose(f) == 0) && ok;
    return ok ? true : set_error(std::string("tb200_scene_cache_save: write failed: ") + path);
}

bool image_load(const char* path, SceneImage* img)
{
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(std::string("tb200_create_cached: cannot open ") + path);
    fseek(f, 0, SEEK_END);
    const uint64_t fileBytes = (uint64_t)std::max(0L, ftell(f));
    fseek(f, 0, SEEK_SET);
    char magic[8];
    CacheStamp stamp;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kCacheMagic, 8) == 0 && get_pod(f, &stamp);
    if (ok && (stamp.version != kCacheVersion || stamp.sizeofPrim != sizeof(DPrim) || stamp.sizeofPair != sizeof(BvhPair) ||
               stamp.sizeofOp != sizeof(ProgOp))) {
        fclose(f);
        return set_error(std::string("tb200_create_cached: ") + path + " was written by another build of this library (layout stamp differs): recreate it");
    }
    ok = ok && get_vec(f, &img->prims, fileBytes) && get_vec(f, &img->scenePairs, fileBytes) && get_pod(f, &img->sceneRoot) &&
         get_vec(f, &img->flat, fileBytes);
    uint64_t numMeshes = 0;
    ok = ok && get_pod(f, &numMeshes) && numMeshes <= fileBytes;
    if (ok) img->meshes.resize(size_t(numMeshes));
    for (size_t m = 0; ok && m < img->meshes.size(); ++m) {
        MeshImage& mi = img->meshes[m];
        ok = get_vec(f, &mi.pairs, fileBytes) && get_vec(f, &mi.verts, fileBytes) && get_vec(f, &mi.norms, fileBytes) &&
             get_vec(f, &mi.cdf, fileBytes) && get_pod(f, &mi.numTris) && get_pod(f, &mi.rootRef) && get_pod(f, &mi.depth);
    }
    ok = ok && get_pod(f, &img->horizon) && get_pod(f, &img->zenith) && get_pod(f, &img->numNee) && get_pod(f, &img->maxTris) &&
         get_pod(f, &img->splitBoxValid) && get_pod(f, &img->splitLo) && get_pod(f, &img->splitHi);
    ok = ok && get_pod(f, &img->probeValid) && get_pod(f, &img->probeW) && get_pod(f, &img->probeH) && get_vec(f, &img->probeData, fileBytes) &&
         get_vec(f, &img->pdfX, fileBytes) && get_vec(f, &img->cdfX, fileBytes) && get_vec(f, &img->pdfY, fileBytes) &&
         get_vec(f, &img->cdfY, fileBytes);
    fclose(f);
    if (!ok) return set_error(std::string("tb200_create_cached: truncated or malformed cache file ") + path);
    return true;
}

// every index a kernel will follow stays inside its array (a cache file is as untrusted as a tb200_scene)
bool validate_image(const SceneImage& img, std::string* why)
{
    const size_t np = img.prims.size();
    if (np == 0 || np > 4095) return *why = "bad primitive count", false;
    auto ref_ok = [](uint32_t ref, size_t pairs, size_t items) { return (ref & TB_LEAF) ? (ref & ~TB_LEAF) < items : ref < pairs; };
    if (!ref_ok(img.sceneRoot, img.scenePairs.size(), np)) return *why = "scene root reference out of range", false;
    for (const BvhPair& p : img.scenePairs)
        if (!ref_ok(p.left, img.scenePairs.size(), np) || !ref_ok(p.right, img.scenePairs.size(), np)) return *why = "scene BVH reference out of range", false;
    for (const MeshImage& m : img.meshes) {
        const size_t nt = (size_t)std::max(0, m.numTris);
        if (nt == 0 || m.verts.size() != nt * 3 || m.norms.size() != nt * 3 || m.cdf.size() != nt) return *why = "mesh arrays do not match the triangle count", false;
        if (!ref_ok(m.rootRef, m.pairs.size(), nt)) return *why = "mesh root reference out of range", false;
        for (const BvhPair& p : m.pairs)
            if (!ref_ok(p.left, m.pairs.size(), nt) || !ref_ok(p.right, m.pairs.size(), nt)) return *why = "mesh BVH reference out of range", false;
    }
    for (const DPrim& p : img.prims) {
        if (p.type != TB200_SPHERE && p.type != TB200_PLANE && p.type != TB200_MESH) return *why = "unknown primitive type", false;
        if (p.type == TB200_MESH && (p.mesh < 0 || (size_t)p.mesh >= img.meshes.size())) return *why = "primitive references a missing mesh", false;
        if (p.lightSamples < 0 || p.lightSamples > 4095) return *why = "lightSamples out of range", false;
    }
    for (const ProgOp& op : img.flat)
        if ((size_t)(op.kindPrim >> 8) >= np && (op.kindPrim & 0xff) != TB_OP_BOX) return *why = "scene program references a missing primitive", false;
    if (img.flat.size() > 32) return *why = "scene program too long", false;
    if (img.probeValid) {
        const size_t n = (size_t)std::max(0, img.probeW) * (size_t)std::max(0, img.probeH);
        if (n == 0 || img.probeData.size() != n + 1 || img.pdfX.size() != n + 1 || img.cdfX.size() != n + 1 ||
            img.pdfY.size() != (size_t)img.probeH + 1 || img.cdfY.size() != (size_t)img.probeH + 1)
            return *why = "probe tables do not match the probe size", false;
    }
    return true;
}

// pixel rows a sample of row y can splat into: y - reach .. y + reach (render.cpp:404-407)
int filter_reach(float filterWidth) { return (int)ceilf(std::max(0.0f, filterWidth)) + 1; }

bool fill_params(tb200_renderer* r, const tb200_camera* camera, const tb200_options* o, LaunchParams* P)
{
    if (!r->dAccum) return set_error("tb200_init has not been called");
    if (o->width != r->width || o->height != r->height)
        return set_error("options.width/height differ from the last tb200_init");
    memset(P, 0, sizeof(*P));
    P->scene = r->scene;
    if (r->tunePhase == 2) P->scene.splitValid = 0;
    camera_setup(*camera, o->width, o->height, &P->camera);
    P->film.width = o->width;
    P->film.height = o->height;
    P->film.filterType = o->filterType;
    P->film.filterWidth = o->filterWidth;
    P->film.filterFalloff = o->filterFalloff;
    P->film.filterOffset = o->filterOffset;
    P->film.clamp = o->clamp;
    P->film.maxDepth = o->maxDepth;
    // Gaussian() is exactly +0 wherever expf(arg) < offset; keep a 1e-3 relative margin so the
    // shortcut never depends on the last bits of expf (tb_film.cuh: filter_gaussian)
    P->film.filterArgZero = -INFINITY;
    if (o->filterOffset > 0.0f && o->filterOffset < 1.0f) {
        const float a = (float)log((double)o->filterOffset * (1.0 - 1.0e-3));
        if (tbm_expf(a) < o->filterOffset) P->film.filterArgZero = a;
    }
    P->accum = r->boundAccum ? r->boundAccum : r->dAccum;
    P->sampleCounter = r->dCounter;
    P->cold = r->dCold;
    P->hardPhases = r->hardPhases;
    P->wideCta = r->wideCta;
    P->laneQueues = r->laneQueues;
    P->firstRow = 0;
    P->numRows = o->height;
    P->film.rowLo = 0;
    P->film.rowHi = o->height;
    if (r->slabRows >= 0) {
        // owner-computes slab: trace the owned rows plus every row whose samples can reach them
        // (render.cpp:404-407: a sample of row y splats into rows int(y' - w) .. int(y' + w), y' in [y, y+1]),
        // splat into the owned rows only
        const int reach = filter_reach(o->filterWidth);
        const int lo = std::max(0, std::min(r->slabRow0, o->height));
        const int hi = std::max(lo, std::min(r->slabRow0 + r->slabRows, o->height));
        P->film.rowLo = lo;
        P->film.rowHi = hi;
        P->firstRow = std::max(0, lo - reach);
        P->numRows = hi > lo ? std::min(o->height, hi + reach) - P->firstRow : 0;
    }
    P->shard = r->shard;
    P->numShards = r->numShards;
    P->walk = r->walk;
    P->walk.numWalkers = r->numWalkers;
    finalize_params(P);
    return true;
}

bool launch_frames(tb200_renderer* r, LaunchParams& P, bool recordStart = true)
{
    unsigned long long launches = 0;
    if (recordStart) TB_CUDA(cudaEventRecord(r->evStart, r->stream));
    if (P.samplesPerFrame > 0) {
        // slot records keep a 32-bit sample index: split very long jobs into several launches
        const int framesPerLaunch = (int)std::max<unsigned long long>(1ull, 0x7fffffffull / P.samplesPerFrame);
        const int frame0 = P.frame0, frames = P.numFrames;
        for (int f = 0; f < frames; f += framesPerLaunch) {
            P.frame0 = frame0 + f;
            P.numFrames = std::min(framesPerLaunch, frames - f);
            if (r->pipeline == 0)
                launch_mega(P, r->stream, &launches);
            else
                launch_wavefront2(P, r->numSMs, r->stream, &launches);
        }
        P.frame0 = frame0;
        P.numFrames = frames;
    }
    TB_CUDA(cudaEventRecord(r->evStop, r->stream));
    TB_CUDA(cudaGetLastError());
    r->stats.kernelLaunches += launches;
    {
        // samples inside the image (tile padding excluded; with a slab: the owned rows, not their halo)
        uint64_t rows = 0;
        if (r->slabRows >= 0)
            rows = (uint64_t)(P.film.rowHi - P.film.rowLo);
        else
            for (int t = P.shard; t * 4 < P.numRows; t += P.numShards) rows += (uint64_t)((P.numRows - t * 4) < 4 ? (P.numRows - t * 4) : 4);
        r->stats.samples += rows * (uint64_t)P.film.width * (uint64_t)P.numFrames;
        r->launchSamples = rows * (uint64_t)P.film.width * (uint64_t)P.numFrames;
    }
    return true;
}

// called once the time of a path-tracing launch is in stats.gpuMs
void tune_update(tb200_renderer* r)
{
    if (r->tunePhase >= 3 || r->launchSamples < (1ull << 18) || !(r->stats.gpuMs > 0.0)) return;
    const double rate = (double)r->launchSamples / r->stats.gpuMs;
    if (r->tunePhase == 1) r->tuneRate[0] = rate;
    if (r->tunePhase == 2) {
        r->tuneRate[1] = rate;
        r->scene.splitValid = r->tuneRate[0] >= r->tuneRate[1] ? 1 : 0;
    }
    r->tunePhase += 1;
}

bool finish_timing(tb200_renderer* r)
{
    TB_CUDA(cudaStreamSynchronize(r->stream));
    if (!check_offload(r)) return false;
    float ms = 0.0f;
    TB_CUDA(cudaEventElapsedTime(&ms, r->evStart, r->evStop));
    r->stats.gpuMs = ms;
    tune_update(r);
    return true;
}

// pixel rows [row0, row1) this renderer is responsible for delivering to the host
void owned_rows(const tb200_renderer* r, int* row0, int* row1)
{
    if (r->slabRows < 0) {
        *row0 = 0;
        *row1 = r->height;
    } else {
        *row0 = std::max(0, std::min(r->slabRow0, r->height));
        *row1 = std::max(*row0, std::min(r->slabRow0 + r->slabRows, r->height));
    }
}

// device -> host copy of the owned rows of the accumulator after the stream's work.  `output` is
// the caller's whole-image buffer; whether it is pinned (tb200_pin_output) is the caller's choice:
// a pageable buffer takes the driver's staged path.
bool read_back(tb200_renderer* r, float* output)
{
    int row0, row1;
    owned_rows(r, &row0, &row1);
    const size_t rowBytes = size_t(r->width) * sizeof(float4);
    const size_t bytes = size_t(row1 - row0) * rowBytes;
    const char* accum = (const char*)(r->boundAccum ? r->boundAccum : r->dAccum);
    if (bytes)
        TB_CUDA(cudaMemcpyAsync((char*)output + size_t(row0) * rowBytes, accum + size_t(row0) * rowBytes, bytes, cudaMemcpyDeviceToHost, r->stream));
    TB_CUDA(cudaStreamSynchronize(r->stream));
    r->stats.d2hBytes += bytes;
    return check_offload(r);
}

// One frame of the wavefront kernel with the read-back streamed underneath it: samples are handed
// out in tile-row order, the kernel flags each band of tile rows as its last sample retires
// (wavefront2.cuh, wf2_band_report), and this thread copies every owned pixel row that no
// unfinished sample can still touch while the kernel is tracing the rest.  What is left when the
// kernel ends is the last band or two instead of the whole slab.
bool render_streamed(tb200_renderer* r, LaunchParams& P, float* output, bool recordStart = true)
{
    const float4* accum = P.accum;
    const size_t rowBytes = size_t(r->width) * sizeof(float4);
    // bands of whole tile rows, at most TB_MAX_BANDS, about 256 KiB each: what is left to copy when the
    // kernel ends is one band plus the filter's reach, and consecutive finished bands go out in one copy
    int bandTileRows = std::max(1, (int)((size_t(1) << 18) / (rowBytes * 4)));
    bandTileRows = std::max(bandTileRows, (P.tileRows + TB_MAX_BANDS - 1) / TB_MAX_BANDS);
    const int numBands = (P.tileRows + bandTileRows - 1) / bandTileRows;
    const int reach = filter_reach(P.film.filterWidth);
    int own0, own1;
    owned_rows(r, &own0, &own1);

    r->bandTag += 1;
    if (r->bandTag == 0) r->bandTag = 1;
    P.bandCount = r->dBandCount;
    P.bandFlags = r->dBandFlags;
    P.bandSamples = (unsigned)bandTileRows * (unsigned)P.tilesX * 32u;
    P.bandTag = r->bandTag;
    if (recordStart) TB_CUDA(cudaEventRecord(r->evStart, r->stream));
    TB_CUDA(cudaMemsetAsync(r->dBandCount, 0, TB_MAX_BANDS * sizeof(unsigned int), r->stream));
    if (!launch_frames(r, P, false)) return false;

    int done = 0;          // bands 0..done-1 are complete
    int copied = own0;     // owned pixel rows [own0, copied) are already on their way to the host
    bool copyFailed = false;
    auto copy_rows_below = [&](int limit) {
        limit = std::min(limit, own1);
        if (limit <= copied) return;
        if (cudaMemcpyAsync((char*)output + size_t(copied) * rowBytes, (const char*)accum + size_t(copied) * rowBytes,
                            size_t(limit - copied) * rowBytes, cudaMemcpyDeviceToHost, r->copyStream) != cudaSuccess)
            copyFailed = true;
        copied = limit;
    };
    unsigned spins = 0;
    for (;;) {
        while (done < numBands && r->hBandFlags[done] == r->bandTag) ++done;
        if (done >= numBands) break;
        // unfinished samples sit in local tile rows >= done*bandTileRows, i.e. global pixel rows
        // >= firstRow + done*bandTileRows*numShards*4 (decode_sample)
        copy_rows_below(P.firstRow + done * bandTileRows * P.numShards * 4 - reach);
        // The band flags (plain reads of mapped host memory) are the signal; the event is only the safety net for a
        // launch that failed -- and a driver call per spin from every device's thread of a multi-device renderer
        // contends on the driver's locks, so it is asked rarely.
        ++spins;
        if ((spins & 1023u) == 0u && cudaEventQuery(r->evStop) != cudaErrorNotReady) break;   // finished (or failed): stop polling
        // back off: the calling thread may be the host application's UI thread (tinsel's GLUT loop)
        if (spins < 4096u)
            __builtin_ia32_pause();
        else
            std::this_thread::yield();
    }
    TB_CUDA(cudaStreamSynchronize(r->stream));
    copy_rows_below(own1);
    TB_CUDA(cudaStreamSynchronize(r->copyStream));
    if (copyFailed) TB_CUDA(cudaErrorUnknown);
    r->stats.d2hBytes += size_t(own1 - own0) * rowBytes;
    return check_offload(r);
}

}  // namespace

// One host thread per further device of a multi-device renderer.  Render() is called back to back
// (16 times per displayed frame, src/main.cpp:242-251) and a frame's kernel takes a fraction of a
// millisecond, so a worker spins briefly for its next job before it goes to sleep on the condition
// variable: a wake-up through the kernel would cost as much as the job.
struct Worker {
    std::thread thread;
    std::mutex lock;
    std::condition_variable wake;
    std::function<int(tb200_renderer*)> job;
    std::atomic<uint32_t> posted{0}, done{0};
    bool sleeping = false;
    std::atomic<bool> quit{false};
    int rc = 0;
};

namespace {

void worker_main(tb200_renderer* r)
{
    Worker* w = r->worker;
    cudaSetDevice(r->device);
    uint32_t seen = 0;
    for (;;) {
        int spins = 0;
        while (w->posted.load(std::memory_order_acquire) == seen) {
            if (++spins < 200000) {
                __builtin_ia32_pause();
                continue;
            }
            std::unique_lock<std::mutex> guard(w->lock);
            w->sleeping = true;
            w->wake.wait(guard, [&] { return w->posted.load(std::memory_order_acquire) != seen || w->quit.load(); });
            w->sleeping = false;
            spins = 0;
        }
        // tb200_destroy bumps `posted` to get a spinning worker out of its loop: that is not a job (the last one is
        // still in `job` -- running it again would render one more frame into a buffer the caller may have freed)
        if (w->quit.load(std::memory_order_acquire)) return;
        seen += 1;
        g_error.clear();
        w->rc = w->job(r);
        r->workerError = w->rc != 0 ? g_error : std::string();
        w->done.store(seen, std::memory_order_release);
    }
}

void worker_post(tb200_renderer* r, const std::function<int(tb200_renderer*)>& job)
{
    Worker* w = r->worker;
    w->job = job;
    w->posted.fetch_add(1, std::memory_order_release);
    std::lock_guard<std::mutex> guard(w->lock);
    if (w->sleeping) w->wake.notify_one();
}

int worker_wait(tb200_renderer* r)
{
    Worker* w = r->worker;
    const uint32_t want = w->posted.load(std::memory_order_relaxed);
    unsigned spins = 0;
    while (w->done.load(std::memory_order_acquire) != want) {
        if (++spins < 4096u)
            __builtin_ia32_pause();
        else
            std::this_thread::yield();
    }
    return w->rc;
}

// runs `job` on the head (calling thread) and on every peer (its worker thread) at the same time;
// returns 0 when all succeeded, else -1 with the first failure as this thread's last error
int group_run(tb200_renderer* head, const std::function<int(tb200_renderer*)>& job)
{
    for (tb200_renderer* p : head->peers) worker_post(p, job);
    cudaSetDevice(head->device);
    int rc = job(head);
    for (tb200_renderer* p : head->peers) {
        if (worker_wait(p) != 0 && rc == 0) {
            rc = -1;
            set_error("device " + std::to_string(p->device) + ": " + p->workerError);
        }
    }
    return rc;
}

// contiguous row slabs, cut at tile rows (4 pixel rows): member k of n owns [cut(k), cut(k+1))
int slab_cut(int height, int k, int n)
{
    if (k >= n) return height;
    const int tileRows = (height + 3) / 4;
    return std::min(height, (int)((long long)tileRows * k / n) * 4);
}

}  // namespace

extern "C" {

const char* tb200_last_error(void)
{
    // two sticky messages (this file's and snapshot.cpp's): the more recent one
    if (!g_error.empty() && g_errorStamp > tb200_snapshot_error_stamp()) return g_error.c_str();
    const char* snap = tb200_snapshot_error();
    if (snap && snap[0]) return snap;
    return g_error.c_str();
}

// Structural checks on the caller's scene before anything is uploaded: every index the kernels will
// follow must stay inside its array (the reference trusts its own loader; a C ABI cannot).
static bool validate_bvh(const tb200_bvh_node* nodes, int numNodes, int numItems, const char* what, std::string* why)
{
    if (numNodes <= 0 || !nodes) {
        *why = std::string(what) + ": no BVH nodes";
        return false;
    }
    for (int i = 0; i < numNodes; ++i) {
        const bool leaf = (nodes[i].right_leaf >> 31) != 0;
        const uint32_t left = nodes[i].left, right = nodes[i].right_leaf & 0x7fffffffu;
        if (leaf ? left >= (uint32_t)numItems : (left >= (uint32_t)numNodes || right >= (uint32_t)numNodes)) {
            *why = std::string(what) + ": BVH node " + std::to_string(i) + " points outside its array";
            return false;
        }
    }
    return true;
}

static bool validate_scene(const tb200_scene* s, std::string* why)
{
    if (s->numPrimitives <= 0 || !s->primitives) return *why = "scene has no primitives", false;
    if (s->numPrimitives > 4095) return *why = "more than 4095 primitives (the wavefront's NEE cursor holds 12-bit primitive indices)", false;
    if (s->numMeshes < 0 || (s->numMeshes > 0 && !s->meshes)) return *why = "bad mesh array", false;
    if (!validate_bvh(s->bvhNodes, s->numBvhNodes, s->numPrimitives, "scene", why)) return false;
    for (int m = 0; m < s->numMeshes; ++m) {
        const tb200_mesh& g = s->meshes[m];
        const std::string name = "mesh " + std::to_string(m);
        if (g.numVertices <= 0 || g.numIndices <= 0 || g.numIndices % 3 != 0 || !g.positions || !g.normals || !g.indices)
            return *why = name + ": empty or incomplete geometry", false;
        for (int i = 0; i < g.numIndices; ++i)
            if (g.indices[i] < 0 || g.indices[i] >= g.numVertices) return *why = name + ": vertex index out of range", false;
        if (!validate_bvh(g.nodes, g.numNodes, g.numIndices / 3, name.c_str(), why)) return false;
    }
    for (int i = 0; i < s->numPrimitives; ++i) {
        const tb200_primitive& p = s->primitives[i];
        if (p.type != TB200_SPHERE && p.type != TB200_PLANE && p.type != TB200_MESH)
            return *why = "primitive " + std::to_string(i) + ": unknown type", false;
        if (p.type == TB200_MESH && (p.mesh < 0 || p.mesh >= s->numMeshes))
            return *why = "primitive " + std::to_string(i) + ": mesh index out of range", false;
        if (p.type == TB200_MESH && p.lightSamples > 0 && !s->meshes[p.mesh].cdf)
            return *why = "primitive " + std::to_string(i) + ": mesh light without a sampling CDF", false;
        if (p.lightSamples < 0 || p.lightSamples > 4095) return *why = "primitive " + std::to_string(i) + ": lightSamples out of range", false;
    }
    if (s->sky.probeValid) {
        const tb200_sky& k = s->sky;
        if (k.probeWidth <= 0 || k.probeHeight <= 0 || !k.probeData || !k.pdfValuesX || !k.cdfValuesX || !k.pdfValuesY || !k.cdfValuesY)
            return *why = "sky probe: missing tables", false;
    }
    return true;
}

static tb200_renderer* create_with(int device, const std::function<bool(tb200_renderer*)>& build);

tb200_renderer* tb200_create(const tb200_scene* scene, int device)
{
    g_error.clear();
    if (!scene) {
        set_error("tb200_create: null scene");
        return nullptr;
    }
    {
        std::string why;
        if (!validate_scene(scene, &why)) {
            set_error("tb200_create: invalid scene: " + why);
            return nullptr;
        }
    }
    return create_with(device, [scene](tb200_renderer* r) { return build_scene(r, scene); });
}

int tb200_scene_cache_save(const tb200_scene* scene, const char* path)
{
    g_error.clear();
    if (!scene || !path) {
        set_error("tb200_scene_cache_save: null argument");
        return -1;
    }
    std::string why;
    if (!validate_scene(scene, &why)) {
        set_error("tb200_scene_cache_save: invalid scene: " + why);
        return -1;
    }
    SceneImage img;
    if (!image_from_scene(scene, &img)) return -1;
    return image_save(img, path) ? 0 : -1;
}

tb200_renderer* tb200_create_cached(const char* path, int device)
{
    g_error.clear();
    if (!path) {
        set_error("tb200_create_cached: null path");
        return nullptr;
    }
    // SceneImage is heavy (hundreds of MB for a big mesh + probe): shared with the build step, freed after it
    std::shared_ptr<SceneImage> img(new SceneImage());
    try {
        if (!image_load(path, img.get())) return nullptr;
    } catch (const std::exception& e) {   // bad_alloc from a corrupt count that passed the size check
        set_error(std::string("tb200_create_cached: ") + e.what());
        return nullptr;
    }
    std::string why;
    if (!validate_image(*img, &why)) {
        set_error("tb200_create_cached: invalid cache file: " + why);
        return nullptr;
    }
    return create_with(device, [img](tb200_renderer* r) { return upload_image(r, *img); });
}

static tb200_renderer* create_with(int device, const std::function<bool(tb200_renderer*)>& build)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        set_error("tb200_create: no CUDA device available (this library has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= count) {
        set_error("tb200_create: bad device ordinal");
        return nullptr;
    }
    tb200_renderer* r = new tb200_renderer();
    memset(&r->stats, 0, sizeof(r->stats));
    r->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    cudaDeviceProp prop;
    ok = ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok) r->numSMs = prop.multiProcessorCount;
    ok = ok && cudaStreamCreateWithFlags(&r->ownStream, cudaStreamNonBlocking) == cudaSuccess;
    r->stream = r->ownStream;
    ok = ok && cudaEventCreate(&r->evStart) == cudaSuccess && cudaEventCreate(&r->evStop) == cudaSuccess;
    if (!ok) {
        set_error(std::string("tb200_create: device setup failed: ") + cudaGetErrorString(cudaGetLastError()));
        delete r;
        return nullptr;
    }
    const char* pipe = getenv("TINSEL_B200_PIPELINE");
    if (pipe && strcmp(pipe, "mega") == 0) r->pipeline = 0;
    const char* rb = getenv("TINSEL_B200_READBACK");
    if (rb && strcmp(rb, "plain") == 0) r->streamedReadback = 0;
    unsigned int* hostFlags = nullptr;
    // cold slot state for every CTA a launch can have (one or two per SM, see launch_wavefront2)
    if (cudaMalloc((void**)&r->dCold, size_t(r->numSMs) * TB_WF2_MAX_CTAS_PER_SM * TB_WF2_SLOTS * 128) != cudaSuccess ||
        cudaMalloc((void**)&r->dCounter, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void**)&r->dBandCount, TB_MAX_BANDS * sizeof(unsigned int)) != cudaSuccess ||
        cudaHostAlloc((void**)&hostFlags, TB_MAX_BANDS * sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&r->dBandFlags, hostFlags, 0) != cudaSuccess ||
        cudaStreamCreateWithFlags(&r->copyStream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("tb200_create: counter allocation failed");
        if (hostFlags) cudaFreeHost(hostFlags);
        tb200_destroy(r);
        return nullptr;
    }
    memset(hostFlags, 0, TB_MAX_BANDS * sizeof(unsigned int));
    r->hBandFlags = hostFlags;
    if (!build(r)) {
        tb200_destroy(r);
        return nullptr;
    }
    return r;
}

tb200_renderer* tb200_create_multi(const tb200_scene* scene, const int* devices, int numDevices)
{
    if (!devices || numDevices < 1 || numDevices > TB200_MAX_DEVICES) {
        set_error("tb200_create_multi: bad device list");
        return nullptr;
    }
    for (int a = 0; a < numDevices; ++a)
        for (int b = 0; b < a; ++b)
            if (devices[a] == devices[b] && !getenv("TINSEL_B200_TEST_DUP_DEVICES")) {   // test hook: a one-GPU box
                set_error("tb200_create_multi: a device is listed twice");
                return nullptr;
            }
    // one complete renderer per device (the scene is replicated: SURVEY 8e, at most ~250 MB), built
    // side by side because most of the time is the host-side re-layout of the meshes
    std::vector<tb200_renderer*> members(numDevices, nullptr);
    std::vector<std::string> errors(numDevices);
    {
        std::vector<std::thread> builders;
        for (int k = 0; k < numDevices; ++k)
            builders.emplace_back([&, k] {
                members[k] = tb200_create(scene, devices[k]);
                if (!members[k]) errors[k] = g_error;
            });
        for (std::thread& t : builders) t.join();
    }
    for (int k = 0; k < numDevices; ++k)
        if (!members[k]) {
            set_error("tb200_create_multi: device " + std::to_string(devices[k]) + ": " + errors[k]);
            for (tb200_renderer* m : members) tb200_destroy(m);
            return nullptr;
        }
    tb200_renderer* head = members[0];
    for (int k = 1; k < numDevices; ++k) {
        tb200_renderer* p = members[k];
        head->peers.push_back(p);
        // device-to-device gathers go over NVLink when the pair allows peer access
        int can = 0;
        if (p->device != head->device && cudaDeviceCanAccessPeer(&can, p->device, head->device) == cudaSuccess && can) {
            cudaSetDevice(p->device);
            if (cudaDeviceEnablePeerAccess(head->device, 0) != cudaSuccess) cudaGetLastError();
        }
        p->worker = new Worker();
        p->worker->thread = std::thread(worker_main, p);
    }
    cudaSetDevice(head->device);
    return head;
}

static int init_one(tb200_renderer* r, int width, int height)
{
    cudaSetDevice(r->device);
    const size_t n = size_t(width) * height;
    if (width != r->width || height != r->height || !r->dAccum) {
        cudaFree(r->dAccum);
        cudaFree(r->dRadiance);
        cudaFree(r->dRaster);
        r->dAccum = nullptr;
        r->dRadiance = r->dRaster = nullptr;
        if (cudaMalloc((void**)&r->dAccum, n * sizeof(float4)) != cudaSuccess) {
            set_error("tb200_init: accumulator allocation failed");
            return -1;
        }
        r->width = width;
        r->height = height;
    }
    if (cudaMemsetAsync(r->dAccum, 0, n * sizeof(float4), r->stream) != cudaSuccess ||
        cudaStreamSynchronize(r->stream) != cudaSuccess) {
        set_error("tb200_init: accumulator clear failed");
        return -1;
    }
    r->frame = 0;
    r->stats.frames = 0;
    r->boundAccum = nullptr;
    r->filteredValid = false;
    return 0;
}

int tb200_init(tb200_renderer* r, int width, int height)
{
    if (!r || width <= 0 || height <= 0) {
        set_error("tb200_init: bad arguments");
        return -1;
    }
    if (r->peers.empty()) return init_one(r, width, height);
    // multi-device: contiguous row slabs cut at tile rows, one per device
    const int n = (int)r->peers.size() + 1;
    r->slabRow0 = 0;
    r->slabRows = slab_cut(height, 1, n);
    for (int k = 1; k < n; ++k) {
        r->peers[k - 1]->slabRow0 = slab_cut(height, k, n);
        r->peers[k - 1]->slabRows = slab_cut(height, k + 1, n) - slab_cut(height, k, n);
    }
    return group_run(r, [=](tb200_renderer* m) { return init_one(m, width, height); });
}

static int render_one(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, float* output)
{
    cudaSetDevice(r->device);
    LaunchParams P;
    if (!fill_params(r, camera, options, &P)) return -1;
    if (options->mode == TB200_MODE_NORMALS) {
        unsigned long long launches = 0;
        if (r->slabRows >= 0) {
            // eNormals writes each pixel from its own primary ray: the owned rows, no halo
            P.firstRow = P.film.rowLo;
            P.numRows = P.film.rowHi - P.film.rowLo;
            finalize_params(&P);
        }
        cudaEventRecord(r->evStart, r->stream);
        launch_normals(P, r->stream, &launches);
        cudaEventRecord(r->evStop, r->stream);
        r->stats.kernelLaunches += launches;
    } else {
        P.frame0 = r->frame;
        P.numFrames = 1;
        const bool streamed = r->streamedReadback && r->pipeline != 0 && P.samplesPerFrame > 0 && P.samplesPerFrame < 0x7fffffffull;
        if (streamed) {
            if (!render_streamed(r, P, output)) return -1;
        } else if (!launch_frames(r, P)) {
            return -1;
        }
        r->frame += 1;
        r->stats.frames += 1;
        if (streamed) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, r->evStart, r->evStop);
            r->stats.gpuMs = ms;
            tune_update(r);
            return 0;
        }
    }
    if (!read_back(r, output)) return -1;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, r->evStart, r->evStop);
    r->stats.gpuMs = ms;
    if (options->mode == TB200_MODE_PATHTRACE) tune_update(r);
    return 0;
}

int tb200_render(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, float* output)
{
    if (!r || !camera || !options || !output) {
        set_error("tb200_render: null argument");
        return -1;
    }
    if (options->mode == TB200_MODE_COMPLEXITY) return 0;   // render.cpp:516-519
    if (r->peers.empty()) return render_one(r, camera, options, output);
    // every device traces its slab and streams its own rows into `output` over its own PCIe link
    return group_run(r, [=](tb200_renderer* m) { return render_one(m, camera, options, output); });
}

static int render_device_one(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, int spp,
                             int firstRow, int numRows)
{
    cudaSetDevice(r->device);
    if (options->mode != TB200_MODE_PATHTRACE) {
        set_error("tb200_render_device: only ePathTrace is batched");
        return -1;
    }
    LaunchParams P;
    if (!fill_params(r, camera, options, &P)) return -1;
    if (numRows >= 0) {
        if (r->slabRows >= 0) {
            set_error("tb200_render_device: a row range cannot be combined with a row slab (tb200_set_slab / multi-device)");
            return -1;
        }
        if (firstRow < 0 || firstRow + numRows > options->height) {
            set_error("tb200_render_device: row range outside the image");
            return -1;
        }
        P.firstRow = firstRow;
        P.numRows = numRows;
        finalize_params(&P);
    }
    P.frame0 = r->frame;
    P.numFrames = spp;
    if (spp > 0 && P.numRows > 0) {
        if (!launch_frames(r, P)) return -1;
        if (!finish_timing(r)) return -1;
    }
    r->frame += spp;
    r->stats.frames += spp;
    return 0;
}

int tb200_render_device(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, int spp,
                        int firstRow, int numRows)
{
    if (!r || !camera || !options || spp < 0) {
        set_error("tb200_render_device: bad argument");
        return -1;
    }
    if (r->peers.empty()) return render_device_one(r, camera, options, spp, firstRow, numRows);
    return group_run(r, [=](tb200_renderer* m) { return render_device_one(m, camera, options, spp, firstRow, numRows); });
}

#define TB_SINGLE_DEVICE_ONLY(r, what)                                                              \
    if (!(r)->peers.empty()) {                                                                      \
        set_error(what ": not available on a multi-device renderer");                               \
        return -1;                                                                                  \
    }

int tb200_set_shard(tb200_renderer* r, int shard, int numShards)
{
    if (!r || numShards < 1 || shard < 0 || shard >= numShards) {
        set_error("tb200_set_shard: bad arguments");
        return -1;
    }
    TB_SINGLE_DEVICE_ONLY(r, "tb200_set_shard")
    r->shard = shard;
    r->numShards = numShards;
    return 0;
}

int tb200_set_stream(tb200_renderer* r, void* cudaStream, int external)
{
    if (!r) {
        set_error("tb200_set_stream: null renderer");
        return -1;
    }
    TB_SINGLE_DEVICE_ONLY(r, "tb200_set_stream")
    cudaSetDevice(r->device);
    cudaStreamSynchronize(r->stream);
    // handle 0 with external != 0 is CUDA's legacy default stream, a perfectly good caller stream
    r->stream = external ? (cudaStream_t)cudaStream : r->ownStream;
    return 0;
}

int tb200_set_slab(tb200_renderer* r, int firstRow, int numRows)
{
    if (!r || (numRows >= 0 && firstRow < 0)) {
        set_error("tb200_set_slab: bad arguments");
        return -1;
    }
    TB_SINGLE_DEVICE_ONLY(r, "tb200_set_slab")
    if (numRows >= 0 && r->numShards > 1) {
        set_error("tb200_set_slab: cannot be combined with tb200_set_shard");
        return -1;
    }
    r->slabRow0 = numRows >= 0 ? firstRow : 0;
    r->slabRows = numRows >= 0 ? numRows : -1;
    return 0;
}

int tb200_pin_output(tb200_renderer* r, float* output, size_t bytes)
{
    if (!r || !output || bytes == 0) {
        set_error("tb200_pin_output: bad arguments");
        return -1;
    }
    cudaSetDevice(r->device);
    if (r->registered == output && r->registeredBytes == bytes) return 0;
    if (r->registered) tb200_unpin_output(r);
    // portable: every device of a multi-device renderer copies its rows straight into this buffer
    if (cudaHostRegister(output, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        set_error(std::string("tb200_pin_output: cudaHostRegister: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    r->registered = output;
    r->registeredBytes = bytes;
    return 0;
}

int tb200_unpin_output(tb200_renderer* r)
{
    if (!r) return -1;
    if (r->registered) {
        cudaSetDevice(r->device);
        cudaHostUnregister(r->registered);
        cudaGetLastError();
        r->registered = nullptr;
        r->registeredBytes = 0;
    }
    return 0;
}

int tb200_bind_accumulator(tb200_renderer* r, float* deviceAccum)
{
    if (!r || !r->dAccum) {
        set_error("tb200_bind_accumulator: call tb200_init first");
        return -1;
    }
    TB_SINGLE_DEVICE_ONLY(r, "tb200_bind_accumulator")
    r->boundAccum = (float4*)deviceAccum;
    return 0;
}

float* tb200_device_accumulator(tb200_renderer* r)
{
    if (!r) return nullptr;
    return (float*)(r->boundAccum ? r->boundAccum : r->dAccum);
}

int tb200_read_accumulator(tb200_renderer* r, float* output)
{
    if (!r || !output || !r->dAccum) {
        set_error("tb200_read_accumulator: bad argument");
        return -1;
    }
    return group_run(r, [=](tb200_renderer* m) {
        cudaSetDevice(m->device);
        return read_back(m, output) ? 0 : -1;
    });
}

int tb200_gather_device(tb200_renderer* r)
{
    if (!r || !r->dAccum) {
        set_error("tb200_gather_device: call tb200_init first");
        return -1;
    }
    // every peer's slab -> the head's accumulator, device to device (NVLink when peer access is on)
    const size_t rowBytes = size_t(r->width) * sizeof(float4);
    for (tb200_renderer* p : r->peers) {
        int row0, row1;
        owned_rows(p, &row0, &row1);
        if (row1 <= row0) continue;
        cudaSetDevice(p->device);
        if (cudaMemcpyPeerAsync((char*)r->dAccum + size_t(row0) * rowBytes, r->device, (const char*)p->dAccum + size_t(row0) * rowBytes,
                                p->device, size_t(row1 - row0) * rowBytes, p->stream) != cudaSuccess) {
            set_error(std::string("tb200_gather_device: ") + cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
    }
    for (tb200_renderer* p : r->peers) {
        cudaSetDevice(p->device);
        if (cudaStreamSynchronize(p->stream) != cudaSuccess) {
            set_error(std::string("tb200_gather_device: ") + cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
    }
    cudaSetDevice(r->device);
    return 0;
}

int tb200_num_devices(const tb200_renderer* r) { return r ? (int)r->peers.size() + 1 : 0; }

void tb200_slab_rows(int height, int member, int numMembers, int* firstRow, int* numRows)
{
    if (numMembers < 1) numMembers = 1;
    member = std::max(0, std::min(member, numMembers - 1));
    const int a = slab_cut(std::max(0, height), member, numMembers), b = slab_cut(std::max(0, height), member + 1, numMembers);
    if (firstRow) *firstRow = a;
    if (numRows) *numRows = b - a;
}

void tb200_slab_traced_rows(int height, int firstRow, int numRows, float filterWidth, int* firstTraced, int* numTraced)
{
    const int reach = filter_reach(filterWidth);
    const int lo = std::max(0, std::min(firstRow, height)), hi = std::max(lo, std::min(firstRow + std::max(0, numRows), height));
    const int t0 = std::max(0, lo - reach);
    if (firstTraced) *firstTraced = t0;
    if (numTraced) *numTraced = hi > lo ? std::min(height, hi + reach) - t0 : 0;
}

static int render_n_one(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, int n, float* output)
{
    cudaSetDevice(r->device);
    LaunchParams P;
    if (!fill_params(r, camera, options, &P)) return -1;
    // frames k .. k+n-2 in one launch, the last one with the read-back streamed underneath it
    const bool streamed = r->streamedReadback && r->pipeline != 0 && P.samplesPerFrame > 0 && P.samplesPerFrame < 0x7fffffffull;
    const int head = streamed ? n - 1 : n;
    bool started = false;
    if (head > 0) {
        P.frame0 = r->frame;
        P.numFrames = head;
        if (!launch_frames(r, P)) return -1;
        started = true;
    }
    if (streamed) {
        P.frame0 = r->frame + head;
        P.numFrames = 1;
        if (!render_streamed(r, P, output, !started)) return -1;
    } else if (!read_back(r, output)) {
        return -1;
    }
    r->frame += n;
    r->stats.frames += n;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, r->evStart, r->evStop);
    r->stats.gpuMs = ms;
    return 0;
}

int tb200_render_n(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, int n, float* output)
{
    if (!r || !camera || !options || !output || n < 1) {
        set_error("tb200_render_n: bad argument");
        return -1;
    }
    if (options->mode != TB200_MODE_PATHTRACE) {
        set_error("tb200_render_n: only ePathTrace is batched");
        return -1;
    }
    if (r->peers.empty()) return render_n_one(r, camera, options, n, output);
    return group_run(r, [=](tb200_renderer* m) { return render_n_one(m, camera, options, n, output); });
}

int tb200_finish(tb200_renderer* r, float exposure, float limit, float* filtered, unsigned char* rgb8)
{
    (void)limit;   // ToneMap's filmic branch ignores it (util.h:25-42), kept for the call shape of main.cpp:269
    if (!r || !r->dAccum) {
        set_error("tb200_finish: tb200_init has not been called");
        return -1;
    }
    if (!r->peers.empty() && tb200_gather_device(r) != 0) return -1;   // the finish step runs on the head device
    cudaSetDevice(r->device);
    const int W = r->width, H = r->height;
    const size_t n = size_t(W) * H;
    const size_t rgbBytes = ((n + 3) / 4) * 12;
    if (r->finishWidth != W || r->finishHeight != H) {
        cudaFree(r->dFiltered);
        cudaFree(r->dRgb8);
        cudaFree(r->dDither);
        cudaFree(r->dMeans);
        cudaFree(r->dDenoised);
        r->dFiltered = nullptr;
        r->dRgb8 = nullptr;
        r->dDither = nullptr;
        r->dMeans = r->dDenoised = nullptr;
        r->ditherReady = false;
        r->filteredValid = false;
        r->finishWidth = r->finishHeight = 0;
        if (cudaMalloc((void**)&r->dFiltered, n * sizeof(float4)) != cudaSuccess ||
            cudaMalloc((void**)&r->dRgb8, rgbBytes) != cudaSuccess ||
            cudaMalloc((void**)&r->dDither, n * sizeof(uint2)) != cudaSuccess) {
            set_error("tb200_finish: buffer allocation failed");
            return -1;
        }
        r->finishWidth = W;
        r->finishHeight = H;
    }
    if (rgb8 && !r->ditherReady) {
        // WritePng (png.cpp:329-343) draws its dither from one sequential Random(): six draws per pixel
        // in pixel order.  The stream does not depend on the image, so the state in front of every
        // pixel is tabulated once per image size and the kernel continues from there.
        std::vector<uint2> states(n);
        uint32_t s1 = 315645664u, s2 = s1 ^ 0x13ab45feu;   // Random(0), maths.h:1041-1045
        for (size_t i = 0; i < n; ++i) {
            states[i] = make_uint2(s1, s2);
            for (int k = 0; k < 6; ++k) {
                const uint32_t a = s1, b = s2;
                s1 = (b ^ ((a << 5) | (a >> 27))) ^ (a * b);
                s2 = s1 ^ ((b << 12) | (b >> 20));
            }
        }
        if (cudaMemcpyAsync(r->dDither, states.data(), n * sizeof(uint2), cudaMemcpyHostToDevice, r->stream) != cudaSuccess ||
            cudaStreamSynchronize(r->stream) != cudaSuccess) {
            set_error("tb200_finish: dither table upload failed");
            return -1;
        }
        r->stats.h2dBytes += n * sizeof(uint2);
        r->ditherReady = true;
    }
    FinishParams F;
    F.accum = r->boundAccum ? r->boundAccum : r->dAccum;
    F.numPixels = (int)n;
    F.exposure = exposure;
    F.filtered = r->dFiltered;   // always kept on the device: tb200_nlm reads it
    F.rgb8 = rgb8 ? r->dRgb8 : nullptr;
    F.ditherState = r->dDither;
    unsigned long long launches = 0;
    cudaEventRecord(r->evStart, r->stream);
    launch_finish(F, r->stream, &launches);
    cudaEventRecord(r->evStop, r->stream);
    r->stats.kernelLaunches += launches;
    bool ok = cudaGetLastError() == cudaSuccess;
    if (ok && filtered) {
        ok = cudaMemcpyAsync(filtered, r->dFiltered, n * sizeof(float4), cudaMemcpyDeviceToHost, r->stream) == cudaSuccess;
        r->stats.d2hBytes += n * sizeof(float4);
    }
    if (ok && rgb8) {
        ok = cudaMemcpyAsync(rgb8, r->dRgb8, n * 3, cudaMemcpyDeviceToHost, r->stream) == cudaSuccess;
        r->stats.d2hBytes += n * 3;
    }
    ok = ok && cudaStreamSynchronize(r->stream) == cudaSuccess;
    if (!ok) {
        set_error(std::string("tb200_finish: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    r->filteredValid = true;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, r->evStart, r->evStop);
    r->stats.gpuMs = ms;
    return 0;
}

int tb200_nlm(tb200_renderer* r, float falloff, int radius, float* out)
{
    if (!r || !out || radius < 0) {
        set_error("tb200_nlm: bad argument");
        return -1;
    }
    if (!r->filteredValid || r->finishWidth != r->width || r->finishHeight != r->height) {
        set_error("tb200_nlm: tb200_finish has not produced an image since the last tb200_init");
        return -1;
    }
    cudaSetDevice(r->device);
    const size_t n = size_t(r->width) * r->height;
    if (!r->dMeans) {
        if (cudaMalloc((void**)&r->dMeans, n * sizeof(float4)) != cudaSuccess ||
            cudaMalloc((void**)&r->dDenoised, n * sizeof(float4)) != cudaSuccess) {
            set_error("tb200_nlm: buffer allocation failed");
            return -1;
        }
    }
    NlmParams N;
    N.in = r->dFiltered;
    N.means = r->dMeans;
    N.out = r->dDenoised;
    N.width = r->width;
    N.height = r->height;
    N.radius = radius;
    N.falloff = falloff;
    unsigned long long launches = 0;
    cudaEventRecord(r->evStart, r->stream);
    launch_nlm(N, r->stream, &launches);
    cudaEventRecord(r->evStop, r->stream);
    r->stats.kernelLaunches += launches;
    if (cudaGetLastError() != cudaSuccess ||
        cudaMemcpyAsync(out, r->dDenoised, n * sizeof(float4), cudaMemcpyDeviceToHost, r->stream) != cudaSuccess ||
        cudaStreamSynchronize(r->stream) != cudaSuccess) {
        set_error(std::string("tb200_nlm: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    r->stats.d2hBytes += n * sizeof(float4);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, r->evStart, r->evStop);
    r->stats.gpuMs = ms;
    return 0;
}

int tb200_trace_frame(tb200_renderer* r, const tb200_camera* camera, const tb200_options* options, int frame,
                      float* radiance, float* raster)
{
    if (!r || !camera || !options || !radiance || !raster) {
        set_error("tb200_trace_frame: null argument");
        return -1;
    }
    cudaSetDevice(r->device);
    LaunchParams P;
    {
        // a probe of the whole frame on this device, whatever slab it owns when rendering
        const int slabRows = r->slabRows;
        r->slabRows = -1;
        const bool ok = fill_params(r, camera, options, &P);
        r->slabRows = slabRows;
        if (!ok) return -1;
    }
    const size_t n = size_t(r->width) * r->height;
    if (!r->dRadiance) {
        if (cudaMalloc((void**)&r->dRadiance, n * 3 * sizeof(float)) != cudaSuccess ||
            cudaMalloc((void**)&r->dRaster, n * 2 * sizeof(float)) != cudaSuccess) {
            set_error("tb200_trace_frame: scratch allocation failed");
            return -1;
        }
    }
    P.frame0 = frame;
    P.numFrames = 1;
    P.outRadiance = r->dRadiance;
    P.outRaster = r->dRaster;
    // pixels of other shards (tb200_set_shard) are not traced by this renderer: they read back as zero
    if (cudaMemsetAsync(r->dRadiance, 0, n * 3 * sizeof(float), r->stream) != cudaSuccess ||
        cudaMemsetAsync(r->dRaster, 0, n * 2 * sizeof(float), r->stream) != cudaSuccess) {
        set_error("tb200_trace_frame: scratch clear failed");
        return -1;
    }
    const uint64_t samplesBefore = r->stats.samples;
    if (!launch_frames(r, P)) return -1;
    r->stats.samples = samplesBefore;
    if (cudaMemcpyAsync(radiance, r->dRadiance, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, r->stream) != cudaSuccess ||
        cudaMemcpyAsync(raster, r->dRaster, n * 2 * sizeof(float), cudaMemcpyDeviceToHost, r->stream) != cudaSuccess ||
        cudaStreamSynchronize(r->stream) != cudaSuccess) {
        set_error(std::string("tb200_trace_frame: ") + cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    r->stats.d2hBytes += n * 5 * sizeof(float);
    if (!check_offload(r)) return -1;
    return 0;
}

void tb200_set_frame(tb200_renderer* r, int frame)
{
    if (!r) return;
    r->frame = frame;
    for (tb200_renderer* p : r->peers) p->frame = frame;
}

void tb200_get_stats(tb200_renderer* r, tb200_stats* out)
{
    if (!r || !out) return;
    *out = r->stats;
    for (const tb200_renderer* p : r->peers) {
        out->samples += p->stats.samples;
        out->kernelLaunches += p->stats.kernelLaunches;
        out->d2hBytes += p->stats.d2hBytes;
        out->h2dBytes += p->stats.h2dBytes;
        out->gpuMs = std::max(out->gpuMs, p->stats.gpuMs);
    }
}

int tb200_get_member_stats(tb200_renderer* r, int member, tb200_stats* out, int* firstRow, int* numRows)
{
    if (!r || !out || member < 0 || member > (int)r->peers.size()) {
        set_error("tb200_get_member_stats: bad argument");
        return -1;
    }
    const tb200_renderer* m = member == 0 ? r : r->peers[member - 1];
    *out = m->stats;
    int row0, row1;
    owned_rows(m, &row0, &row1);
    if (firstRow) *firstRow = row0;
    if (numRows) *numRows = row1 - row0;
    return 0;
}

void tb200_destroy(tb200_renderer* r)
{
    if (!r) return;
    for (tb200_renderer* p : r->peers) tb200_destroy(p);
    r->peers.clear();
    if (r->worker) {
        {
            std::lock_guard<std::mutex> guard(r->worker->lock);
            r->worker->quit.store(true, std::memory_order_release);
        }
        r->worker->posted.fetch_add(1, std::memory_order_release);   // leave the spin, see `quit`
        r->worker->wake.notify_one();
        if (r->worker->thread.joinable()) r->worker->thread.join();
        delete r->worker;
        r->worker = nullptr;
    }
    tb200_unpin_output(r);
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    free_device(r);
    r->stream = r->ownStream;
    if (r->evStart) cudaEventDestroy(r->evStart);
    if (r->evStop) cudaEventDestroy(r->evStop);
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->copyStream) cudaStreamDestroy(r->copyStream);
    delete r;
}

}  // extern "C"
