// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" driver around the reference's own, unmodified sources.  It is compiled in
// the build container only (where /root/reference exists) into oracle/_ref/libtinsel_ref*.so
// by oracle/Makefile; the GPU box uses the prebuilt library.  Nothing here is copied from the
// reference: this TU #includes src/render.cpp where it lies, so CpuRenderer (render.cpp:390-525),
// PathTrace (render.cpp:230) and every inline header they use (intersection.h, disney.h,
// probe.h, ...) are the reference's compiled arithmetic.  loader/mesh/scene/pfm/bvh/platform
// .cpp are compiled as separate objects straight from /root/reference/src.
//
// Two flavours are built from this one file:
//   libtinsel_ref.so          literal: glibc libm, as any user would build tinsel (-O2, no fast-math)
//   libtinsel_ref_detmath.so  -include oracle/detmath_shim.h: sinf/cosf/expf/acosf/atan2 in the
//                             hot path are replaced by include/tb200_detmath.h, the deterministic
//                             versions the CUDA kernels use.  Still the reference's source text.
//
// Drivers exported:
//   ref_render_literal   CreateCpuRenderer + N x Render(): the reference exactly as shipped (one
//                        sequential RNG stream for the image, render.cpp:399,462-486).
//   ref_render_seeded    the same raster loop (render.cpp:462-490) but with Random re-seeded per
//                        (pixel, frame) by tb200_sample_seed(), calling the reference's PathTrace
//                        and CpuRenderer::AddSample.  This is the parity oracle ("oracle B").
//   ref_trace_frame      per-sample radiance + raster position for one frame, no filtering.
//   ref_finish           the display loop of src/main.cpp:262-271 over the reference's own ToneMap /
//                        LinearToSrgb (util.h, maths.h); ref_write_png = the reference's WritePng.
//   ref_*                known-answer hooks for single functions (Random, BSDF*, GenerateRay, ...).
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cfloat>
#include <climits>
#include <atomic>
#include <thread>
#include <vector>

#include "render.cpp"  // the reference's src/render.cpp (via -I/root/reference/src)
#include "loader.h"
#include "png.h"
#include "nlm.h"

#include "tinsel_b200.h"

namespace {

struct RefScene {
    Scene scene;
    Camera camera;
    Options options;
    // export views (tb200_* PODs pointing into the reference's own arrays)
    std::vector<tb200_primitive> xprims;
    std::vector<tb200_mesh> xmeshes;
    std::vector<unsigned long> meshIds;
    tb200_scene xscene;
    tb200_camera xcamera;
    tb200_options xoptions;
    tb200_snapshot* snapshot = nullptr;  // backing store when built from a .tsnap
};

void set_defaults(RefScene* rs)
{
    // src/main.cpp:181-193
    rs->options.width = 512;
    rs->options.height = 256;
    rs->options.filter = Filter(eFilterGaussian, 0.75f, 1.0f);
    rs->options.mode = ePathTrace;
    rs->options.exposure = 1.0f;
    rs->options.limit = 1.5f;
    rs->options.clamp = FLT_MAX;
    rs->options.maxDepth = 4;
    rs->options.maxSamples = INT_MAX;
    rs->camera.position = Vec3(0.0f, 1.0f, 5.0f);
    rs->camera.rotation = Quat();
    rs->camera.fov = DegToRad(35.0f);
}

void to_x(const Transform& t, tb200_transform* o)
{
    o->p[0] = t.p.x; o->p[1] = t.p.y; o->p[2] = t.p.z;
    o->r[0] = t.r.x; o->r[1] = t.r.y; o->r[2] = t.r.z; o->r[3] = t.r.w;
    o->s = t.s;
}

Transform from_x(const tb200_transform& t)
{
    return Transform(Vec3(t.p[0], t.p[1], t.p[2]), Quat(t.r[0], t.r[1], t.r[2], t.r[3]), t.s);
}

void sync_config_views(RefScene* rs)
{
    const Camera& c = rs->camera;
    tb200_camera& xc = rs->xcamera;
    xc.position[0] = c.position.x; xc.position[1] = c.position.y; xc.position[2] = c.position.z;
    xc.rotation[0] = c.rotation.x; xc.rotation[1] = c.rotation.y; xc.rotation[2] = c.rotation.z; xc.rotation[3] = c.rotation.w;
    xc.fov = c.fov; xc.shutterStart = c.shutterStart; xc.shutterEnd = c.shutterEnd;
    const Options& o = rs->options;
    tb200_options& xo = rs->xoptions;
    xo.mode = o.mode; xo.width = o.width; xo.height = o.height;
    xo.filterType = o.filter.type; xo.filterWidth = o.filter.width; xo.filterFalloff = o.filter.falloff;
    xo.filterOffset = o.filter.offset;
    xo.exposure = o.exposure; xo.limit = o.limit; xo.clamp = o.clamp;
    xo.maxDepth = o.maxDepth; xo.maxSamples = o.maxSamples;
}

// Scene -> tb200_scene view (what tinsel_plugin.cpp does for the product; kept separate on purpose)
void build_export(RefScene* rs)
{
    Scene& s = rs->scene;
    rs->xprims.clear();
    rs->xmeshes.clear();
    rs->meshIds.clear();
    for (size_t i = 0; i < s.primitives.size(); ++i) {
        const Primitive& p = s.primitives[i];
        tb200_primitive x;
        memset(&x, 0, sizeof(x));
        to_x(p.startTransform, &x.start);
        to_x(p.endTransform, &x.end);
        x.type = p.type;
        x.mesh = -1;
        if (p.type == eSphere) x.radius = p.sphere.radius;
        if (p.type == ePlane) memcpy(x.plane, p.plane.plane, 16);
        if (p.type == eMesh) {
            int found = -1;
            for (size_t m = 0; m < rs->meshIds.size(); ++m)
                if (rs->meshIds[m] == p.mesh.id) found = int(m);
            if (found < 0) {
                tb200_mesh g;
                g.positions = (const float*)p.mesh.positions;
                g.normals = (const float*)p.mesh.normals;
                g.indices = (const int32_t*)p.mesh.indices;
                g.nodes = (const tb200_bvh_node*)p.mesh.nodes;
                g.cdf = p.mesh.cdf;
                g.numVertices = p.mesh.numVertices;
                g.numIndices = p.mesh.numIndices;
                g.numNodes = p.mesh.numNodes;
                g.area = p.mesh.area;
                found = int(rs->xmeshes.size());
                rs->xmeshes.push_back(g);
                rs->meshIds.push_back(p.mesh.id);
            }
            x.mesh = found;
        }
        const Material& m = p.material;
        tb200_material& xm = x.material;
        memcpy(xm.emission, &m.emission, 12);
        memcpy(xm.color, &m.color, 12);
        memcpy(xm.absorption, &m.absorption, 12);
        xm.eta = m.eta; xm.metallic = m.metallic; xm.subsurface = m.subsurface; xm.specular = m.specular;
        xm.roughness = m.roughness; xm.specularTint = m.specularTint; xm.anisotropic = m.anisotropic;
        xm.sheen = m.sheen; xm.sheenTint = m.sheenTint; xm.clearcoat = m.clearcoat;
        xm.clearcoatGloss = m.clearcoatGloss; xm.transmission = m.transmission;
        x.lightSamples = p.lightSamples;
        rs->xprims.push_back(x);
    }
    tb200_scene& xs = rs->xscene;
    memset(&xs, 0, sizeof(xs));
    xs.primitives = rs->xprims.data();
    xs.numPrimitives = int(rs->xprims.size());
    xs.meshes = rs->xmeshes.data();
    xs.numMeshes = int(rs->xmeshes.size());
    xs.bvhNodes = (const tb200_bvh_node*)s.bvh.nodes;
    xs.numBvhNodes = s.bvh.numNodes;
    memcpy(xs.sky.horizon, &s.sky.horizon, 12);
    memcpy(xs.sky.zenith, &s.sky.zenith, 12);
    if (s.sky.probe.valid) {
        xs.sky.probeValid = 1;
        xs.sky.probeWidth = s.sky.probe.width;
        xs.sky.probeHeight = s.sky.probe.height;
        xs.sky.probeData = (const float*)s.sky.probe.data;
        xs.sky.pdfValuesX = s.sky.probe.pdfValuesX;
        xs.sky.cdfValuesX = s.sky.probe.cdfValuesX;
        xs.sky.pdfValuesY = s.sky.probe.pdfValuesY;
        xs.sky.cdfValuesY = s.sky.probe.cdfValuesY;
    }
    sync_config_views(rs);
}

}  // namespace

extern "C" {

static_assert(sizeof(BVHNode) == sizeof(tb200_bvh_node), "BVHNode layout");
static_assert(sizeof(Vec3) == 12 && sizeof(Color) == 16, "vector layout");

// LoadTin + Scene::Build exactly as src/main.cpp:174-199 does (width/height <= 0 keep the file's).
void* ref_load_tin(const char* path, int width, int height)
{
    RefScene* rs = new RefScene();
    set_defaults(rs);
    if (!LoadTin(path, &rs->scene, &rs->camera, &rs->options)) {
        delete rs;
        return nullptr;
    }
    if (width > 0) rs->options.width = width;    // -width=  src/main.cpp:146
    if (height > 0) rs->options.height = height;  // -height= src/main.cpp:147
    rs->scene.Build();
    build_export(rs);
    return rs;
}

// Rebuild a reference Scene from a snapshot (the GPU box has no .tin/.obj/.hdr files).
void* ref_from_snapshot(const char* path)
{
    tb200_snapshot* snap = tb200_snapshot_load(path);
    if (!snap) return nullptr;
    RefScene* rs = new RefScene();
    rs->snapshot = snap;
    set_defaults(rs);
    const tb200_scene* xs = tb200_snapshot_scene(snap);
    for (int i = 0; i < xs->numPrimitives; ++i) {
        const tb200_primitive& x = xs->primitives[i];
        Primitive p;
        memset(&p.mesh, 0, sizeof(p.mesh));
        p.startTransform = from_x(x.start);
        p.endTransform = from_x(x.end);
        p.type = (GeometryType)x.type;
        if (x.type == TB200_SPHERE) p.sphere.radius = x.radius;
        if (x.type == TB200_PLANE) memcpy(p.plane.plane, x.plane, 16);
        if (x.type == TB200_MESH) {
            const tb200_mesh& g = xs->meshes[x.mesh];
            p.mesh.positions = (const Vec3*)g.positions;
            p.mesh.normals = (const Vec3*)g.normals;
            p.mesh.indices = (const int*)g.indices;
            p.mesh.nodes = (const BVHNode*)g.nodes;
            p.mesh.cdf = g.cdf;
            p.mesh.numVertices = g.numVertices;
            p.mesh.numIndices = g.numIndices;
            p.mesh.numNodes = g.numNodes;
            p.mesh.area = g.area;
            p.mesh.id = (unsigned long)(x.mesh + 1);
        }
        Material m;
        const tb200_material& xm = x.material;
        memcpy(&m.emission, xm.emission, 12);
        memcpy(&m.color, xm.color, 12);
        memcpy(&m.absorption, xm.absorption, 12);
        m.eta = xm.eta; m.metallic = xm.metallic; m.subsurface = xm.subsurface; m.specular = xm.specular;
        m.roughness = xm.roughness; m.specularTint = xm.specularTint; m.anisotropic = xm.anisotropic;
        m.sheen = xm.sheen; m.sheenTint = xm.sheenTint; m.clearcoat = xm.clearcoat;
        m.clearcoatGloss = xm.clearcoatGloss; m.transmission = xm.transmission;
        p.material = m;
        p.lightSamples = x.lightSamples;
        rs->scene.primitives.push_back(p);
    }
    rs->scene.bvh.nodes = (BVHNode*)xs->bvhNodes;  // borrowed from the snapshot; never Clear()ed
    rs->scene.bvh.numNodes = xs->numBvhNodes;
    memcpy(&rs->scene.sky.horizon, xs->sky.horizon, 12);
    memcpy(&rs->scene.sky.zenith, xs->sky.zenith, 12);
    if (xs->sky.probeValid) {
        Probe& pr = rs->scene.sky.probe;
        pr.width = xs->sky.probeWidth;
        pr.height = xs->sky.probeHeight;
        pr.data = (Color*)xs->sky.probeData;
        pr.pdfValuesX = (float*)xs->sky.pdfValuesX;
        pr.cdfValuesX = (float*)xs->sky.cdfValuesX;
        pr.pdfValuesY = (float*)xs->sky.pdfValuesY;
        pr.cdfValuesY = (float*)xs->sky.cdfValuesY;
        pr.valid = true;
    }
    const tb200_camera* xc = tb200_snapshot_camera(snap);
    rs->camera.position = Vec3(xc->position[0], xc->position[1], xc->position[2]);
    rs->camera.rotation = Quat(xc->rotation[0], xc->rotation[1], xc->rotation[2], xc->rotation[3]);
    rs->camera.fov = xc->fov;
    rs->camera.shutterStart = xc->shutterStart;
    rs->camera.shutterEnd = xc->shutterEnd;
    const tb200_options* xo = tb200_snapshot_options(snap);
    rs->options.mode = (RenderMode)xo->mode;
    rs->options.width = xo->width;
    rs->options.height = xo->height;
    rs->options.filter.type = (FilterType)xo->filterType;
    rs->options.filter.width = xo->filterWidth;
    rs->options.filter.falloff = xo->filterFalloff;
    rs->options.filter.offset = xo->filterOffset;
    rs->options.exposure = xo->exposure;
    rs->options.limit = xo->limit;
    rs->options.clamp = xo->clamp;
    rs->options.maxDepth = xo->maxDepth;
    rs->options.maxSamples = xo->maxSamples;
    build_export(rs);
    return rs;
}

const tb200_scene* ref_scene(void* h) { return &((RefScene*)h)->xscene; }
const tb200_camera* ref_camera(void* h) { return &((RefScene*)h)->xcamera; }
const tb200_options* ref_options(void* h) { return &((RefScene*)h)->xoptions; }

// the reference's own objects, for tests that hand them to the C++ plugin (tinsel_plugin.cpp)
const void* ref_native_scene(void* h) { return &((RefScene*)h)->scene; }
const void* ref_native_camera(void* h) { return &((RefScene*)h)->camera; }
const void* ref_native_options(void* h) { return &((RefScene*)h)->options; }

void ref_set_size(void* h, int width, int height)
{
    RefScene* rs = (RefScene*)h;
    rs->options.width = width;
    rs->options.height = height;
    sync_config_views(rs);
}

void ref_set_mode(void* h, int mode)
{
    RefScene* rs = (RefScene*)h;
    rs->options.mode = (RenderMode)mode;
    sync_config_views(rs);
}

void ref_set_max_depth(void* h, int maxDepth)
{
    RefScene* rs = (RefScene*)h;
    rs->options.maxDepth = maxDepth;
    sync_config_views(rs);
}

int ref_save_snapshot(void* h, const char* path)
{
    RefScene* rs = (RefScene*)h;
    return tb200_snapshot_save(path, &rs->xscene, &rs->xcamera, &rs->xoptions);
}

// The reference exactly as shipped: CreateCpuRenderer, `spp` Render() calls into a zeroed buffer.
void ref_render_literal(void* h, int spp, float* out)
{
    RefScene* rs = (RefScene*)h;
    const int n = rs->options.width * rs->options.height;
    std::vector<Color> pixels(n);  // Vec4() zero-initialises (src/maths.h:292), as src/main.cpp:79
    Renderer* r = CreateCpuRenderer(&rs->scene);
    r->Init(rs->options.width, rs->options.height);
    for (int k = 0; k < spp; ++k) r->Render(rs->camera, rs->options, &pixels[0]);
    delete r;
    memcpy(out, &pixels[0], size_t(n) * sizeof(Color));
}

// Seeded restatement of CpuRenderer::Render's ePathTrace loop (src/render.cpp:447-492): same
// draw order (Sample2D x,y ; Sample1D t), same GenerateRay, the reference's PathTrace and
// AddSample; only `rand` is re-seeded per (pixel, frame).  Adds frames [frame0, frame0+nframes)
// into `out` (caller-zeroed running sums).  Rows are split into contiguous bands over `nthreads`
// host threads with private band buffers (+1 halo row each side) summed in band order, so the
// result is deterministic for a given thread count.
void ref_render_seeded(void* h, int frame0, int nframes, float* out, int nthreads)
{
    RefScene* rs = (RefScene*)h;
    const Options& options = rs->options;
    const Camera& camera = rs->camera;
    const int W = options.width, H = options.height;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > H) nthreads = H;

    auto band = [&](int tid, std::vector<Color>* buf) {
        const int r0 = int((long long)H * tid / nthreads), r1 = int((long long)H * (tid + 1) / nthreads);
        CpuRenderer cpu(&rs->scene);
        CameraSampler sampler(Transform(camera.position, camera.rotation), camera.fov, 0.001f, 1.0f, W, H);
        // full-height scratch would cost W*H*16 B per thread; a band with halo rows suffices
        // because the filter footprint is <= 1 pixel (width <= 1, src/render.cpp:426-429).
        const int halo = int(options.filter.width) + 1;
        const int b0 = Max(0, r0 - halo), b1 = Min(H, r1 + halo);
        buf->assign(size_t(W) * (b1 - b0), Color());
        Color* base = buf->data() - size_t(b0) * W;  // so that base[y*W+x] addresses band rows
        for (int k = frame0; k < frame0 + nframes; ++k) {
            for (int j = r0; j < r1; ++j) {
                for (int i = 0; i < W; ++i) {
                    cpu.rand = Random(int(tb200_sample_seed(uint32_t(j * W + i), uint32_t(k))));
                    float x, y, t;
                    Sample2D(cpu.rand, x, y);
                    Sample1D(cpu.rand, t);
                    float time = Lerp(camera.shutterStart, camera.shutterEnd, t);
                    x += i;
                    y += j;
                    Vec3 origin, dir;
                    sampler.GenerateRay(x, y, origin, dir);
                    Vec3 sample = PathTrace(rs->scene, origin, dir, time, options.maxDepth, cpu.rand);
                    cpu.AddSample(base, W, H, x, y, options.clamp, options.filter, sample);
                }
            }
        }
    };

    std::vector<std::vector<Color>> bufs(nthreads);
    std::vector<std::thread> threads;
    for (int t = 0; t < nthreads; ++t) threads.emplace_back(band, t, &bufs[t]);
    for (auto& th : threads) th.join();
    const int halo = int(options.filter.width) + 1;
    Color* o = (Color*)out;
    for (int t = 0; t < nthreads; ++t) {
        const int r0 = int((long long)H * t / nthreads), r1 = int((long long)H * (t + 1) / nthreads);
        const int b0 = Max(0, r0 - halo), b1 = Min(H, r1 + halo);
        for (int y = b0; y < b1; ++y)
            for (int x = 0; x < W; ++x) o[y * W + x] += bufs[t][size_t(y - b0) * W + x];
    }
}

// The CPU arm of the benchmark (bench.py --impl reference, cpu_baseline): the same seeded loop as
// ref_render_seeded -- the reference's own PathTrace and AddSample per sample -- with a work
// distribution that is fair to the CPU: `nthreads` threads pull strips of 4 pixel rows from a shared
// counter (cost per row is far from uniform: cornell's ceiling rows terminate at once), each strip is
// accumulated for all `nframes` frames into a private strip buffer with halo rows, and the strips are
// folded into `out` afterwards in strip order, two passes of non-overlapping strips in parallel.
// The result does not depend on the thread count.
void ref_render_pool(void* h, int frame0, int nframes, float* out, int nthreads)
{
    RefScene* rs = (RefScene*)h;
    const Options& options = rs->options;
    const Camera& camera = rs->camera;
    const int W = options.width, H = options.height;
    const int stripRows = 4;
    const int numStrips = (H + stripRows - 1) / stripRows;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > numStrips) nthreads = numStrips;
    const int halo = int(options.filter.width) + 1;

    std::vector<std::vector<Color>> strips(numStrips);
    std::atomic<int> next(0);
    auto work = [&]() {
        CpuRenderer cpu(&rs->scene);
        CameraSampler sampler(Transform(camera.position, camera.rotation), camera.fov, 0.001f, 1.0f, W, H);
        for (;;) {
            const int sidx = next.fetch_add(1);
            if (sidx >= numStrips) break;
            const int r0 = sidx * stripRows, r1 = Min(H, r0 + stripRows);
            const int b0 = Max(0, r0 - halo), b1 = Min(H, r1 + halo);
            std::vector<Color>& buf = strips[sidx];
            buf.assign(size_t(W) * (b1 - b0), Color());
            Color* base = buf.data() - size_t(b0) * W;
            for (int k = frame0; k < frame0 + nframes; ++k)
                for (int j = r0; j < r1; ++j)
                    for (int i = 0; i < W; ++i) {
                        cpu.rand = Random(int(tb200_sample_seed(uint32_t(j * W + i), uint32_t(k))));
                        float x, y, t;
                        Sample2D(cpu.rand, x, y);
                        Sample1D(cpu.rand, t);
                        float time = Lerp(camera.shutterStart, camera.shutterEnd, t);
                        x += i;
                        y += j;
                        Vec3 origin, dir;
                        sampler.GenerateRay(x, y, origin, dir);
                        Vec3 sample = PathTrace(rs->scene, origin, dir, time, options.maxDepth, cpu.rand);
                        cpu.AddSample(base, W, H, x, y, options.clamp, options.filter, sample);
                    }
        }
    };
    {
        std::vector<std::thread> threads;
        for (int t = 1; t < nthreads; ++t) threads.emplace_back(work);
        work();
        for (auto& th : threads) th.join();
    }
    // fold: strips whose halos overlap (neighbours, and next-but-one when halo > stripRows/2) must not
    // run together: `stride` interleaved passes, each over strips `stride` apart
    Color* o = (Color*)out;
    const int stride = 1 + (2 * halo + stripRows - 1) / stripRows;
    for (int pass = 0; pass < stride; ++pass) {
        std::atomic<int> cursor(pass);
        auto fold = [&]() {
            for (;;) {
                const int sidx = cursor.fetch_add(stride);
                if (sidx >= numStrips) break;
                const int r0 = sidx * stripRows, r1 = Min(H, r0 + stripRows);
                const int b0 = Max(0, r0 - halo), b1 = Min(H, r1 + halo);
                const std::vector<Color>& buf = strips[sidx];
                for (int y = b0; y < b1; ++y)
                    for (int x = 0; x < W; ++x) o[y * W + x] += buf[size_t(y - b0) * W + x];
            }
        };
        std::vector<std::thread> threads;
        const int nfold = Min(nthreads, 16);
        for (int t = 1; t < nfold; ++t) threads.emplace_back(fold);
        fold();
        for (auto& th : threads) th.join();
    }
}

// The finish loop of src/main.cpp:262-271, statement for statement, over caller-supplied sums.
void ref_finish(const float* pixels, int numPixels, float exposure, float limit, float* filtered)
{
    const Color* g_pixels = (const Color*)pixels;
    Color* g_filtered = (Color*)filtered;
    for (int i = 0; i < numPixels; ++i) {
        float s = exposure / g_pixels[i].w;
        g_filtered[i] = LinearToSrgb(ToneMap(g_pixels[i] * s, limit));
    }
}

// The reference's own PNG writer (src/png.cpp:329-371): dithered 8-bit quantisation + TinyPngOut.
void ref_write_png(const float* filtered, int width, int height, const char* path)
{
    WritePng((const Color*)filtered, width, height, path);
}

// The reference's own denoiser (src/nlm.cpp:36-73), as src/main.cpp:275 calls it.
void ref_nlm(const float* in, float* out, int width, int height, float falloff, int radius)
{
    NonLocalMeansFilter((const Color*)in, (Color*)out, width, height, falloff, radius);
}

// The reference's own binary mesh writer / reader (src/mesh.cpp:809-880) for the mesh cache tests.
int ref_mesh_bin_export(void* h, int meshIndex, const char* path)
{
    RefScene* rs = (RefScene*)h;
    if (meshIndex < 0 || meshIndex >= (int)rs->xmeshes.size()) return -1;
    const tb200_mesh& g = rs->xmeshes[meshIndex];
    Mesh m;   // ~Mesh deletes bvh.nodes: give it copies
    m.positions.assign((const Vec3*)g.positions, (const Vec3*)g.positions + g.numVertices);
    m.normals.assign((const Vec3*)g.normals, (const Vec3*)g.normals + g.numVertices);
    m.indices.assign(g.indices, g.indices + g.numIndices);
    m.cdf.assign(g.cdf, g.cdf + g.numIndices / 3);
    m.area = g.area;
    m.bvh.nodes = new BVHNode[g.numNodes];
    memcpy(m.bvh.nodes, g.nodes, sizeof(BVHNode) * g.numNodes);
    m.bvh.numNodes = g.numNodes;
    ExportMeshToBin(path, &m);
    return 0;
}
void* ref_mesh_bin_import(const char* path) { return ImportMeshFromBin(path); }
void ref_mesh_bin_info(void* mesh, int* counts, float* area)
{
    const Mesh* m = (const Mesh*)mesh;
    counts[0] = (int)m->positions.size();
    counts[1] = (int)m->indices.size();
    counts[2] = m->bvh.numNodes;
    *area = m->area;
}
void ref_mesh_bin_copy(void* mesh, float* positions, float* normals, int* indices, void* nodes, float* cdf)
{
    const Mesh* m = (const Mesh*)mesh;
    memcpy(positions, m->positions.data(), m->positions.size() * sizeof(Vec3));
    memcpy(normals, m->normals.data(), m->normals.size() * sizeof(Vec3));
    memcpy(indices, m->indices.data(), m->indices.size() * sizeof(int));
    memcpy(nodes, m->bvh.nodes, sizeof(BVHNode) * m->bvh.numNodes);
    memcpy(cdf, m->cdf.data(), m->cdf.size() * sizeof(float));
}
void ref_mesh_bin_free(void* mesh) { delete (Mesh*)mesh; }

// One frame, no filtering: radiance[3p..] = PathTrace result, raster[2p..] = (x,y) of pixel p.
void ref_trace_frame(void* h, int frame, float* radiance, float* raster, int nthreads)
{
    RefScene* rs = (RefScene*)h;
    const Options& options = rs->options;
    const Camera& camera = rs->camera;
    const int W = options.width, H = options.height;
    if (nthreads < 1) nthreads = 1;
    auto band = [&](int tid) {
        const int r0 = int((long long)H * tid / nthreads), r1 = int((long long)H * (tid + 1) / nthreads);
        CameraSampler sampler(Transform(camera.position, camera.rotation), camera.fov, 0.001f, 1.0f, W, H);
        for (int j = r0; j < r1; ++j)
            for (int i = 0; i < W; ++i) {
                Random rand(int(tb200_sample_seed(uint32_t(j * W + i), uint32_t(frame))));
                float x, y, t;
                Sample2D(rand, x, y);
                Sample1D(rand, t);
                float time = Lerp(camera.shutterStart, camera.shutterEnd, t);
                x += i;
                y += j;
                Vec3 origin, dir;
                sampler.GenerateRay(x, y, origin, dir);
                Vec3 sample = PathTrace(rs->scene, origin, dir, time, options.maxDepth, rand);
                const size_t p = size_t(j) * W + i;
                radiance[p * 3 + 0] = sample.x; radiance[p * 3 + 1] = sample.y; radiance[p * 3 + 2] = sample.z;
                raster[p * 2 + 0] = x; raster[p * 2 + 1] = y;
            }
    };
    std::vector<std::thread> threads;
    for (int t = 0; t < nthreads; ++t) threads.emplace_back(band, t);
    for (auto& th : threads) th.join();
}

void ref_destroy(void* h)
{
    RefScene* rs = (RefScene*)h;
    if (rs->snapshot) {
        rs->scene.bvh.nodes = nullptr;
        tb200_snapshot_free(rs->snapshot);
    }
    // meshes/probe of a LoadTin scene are intentionally leaked (process-lifetime test objects)
    delete rs;
}

// ---- known-answer hooks ------------------------------------------------------------------------

void ref_random_u32(int seed, int n, uint32_t* out)
{
    Random r(seed);
    for (int i = 0; i < n; ++i) out[i] = r.Rand();
}

void ref_random_f32(int seed, int n, float* out)
{
    Random r(seed);
    for (int i = 0; i < n; ++i) out[i] = r.Randf();
}

static Material material_from_x(const tb200_material* xm)
{
    Material m;
    memcpy(&m.emission, xm->emission, 12);
    memcpy(&m.color, xm->color, 12);
    memcpy(&m.absorption, xm->absorption, 12);
    m.eta = xm->eta; m.metallic = xm->metallic; m.subsurface = xm->subsurface; m.specular = xm->specular;
    m.roughness = xm->roughness; m.specularTint = xm->specularTint; m.anisotropic = xm->anisotropic;
    m.sheen = xm->sheen; m.sheenTint = xm->sheenTint; m.clearcoat = xm->clearcoat;
    m.clearcoatGloss = xm->clearcoatGloss; m.transmission = xm->transmission;
    return m;
}

float ref_material_ior(const tb200_material* xm) { return material_from_x(xm).GetIndexOfRefraction(); }

// BSDFEval + BSDFPdf (src/disney.h:296,125)
void ref_bsdf_eval(const tb200_material* xm, float etaI, float etaO, const float* n, const float* v, const float* l,
                   float* f, float* pdf)
{
    Material m = material_from_x(xm);
    Vec3 N(n[0], n[1], n[2]), V(v[0], v[1], v[2]), L(l[0], l[1], l[2]);
    Vec3 r = BSDFEval(m, etaI, etaO, Vec3(0.0f), N, V, L);
    f[0] = r.x; f[1] = r.y; f[2] = r.z;
    *pdf = BSDFPdf(m, etaI, etaO, Vec3(0.0f), N, V, L);
}

// BasisFromVector + BSDFSample (src/maths.h:1261, src/disney.h:170) with Random(seed)
void ref_bsdf_sample(const tb200_material* xm, float etaI, float etaO, const float* n, const float* v, int seed,
                     float* l, float* pdf, int* type, uint32_t* rngAfter)
{
    Material m = material_from_x(xm);
    Vec3 N(n[0], n[1], n[2]), V(v[0], v[1], v[2]);
    Vec3 U, W;
    BasisFromVector(N, &U, &W);
    Random rand(seed);
    Vec3 L;
    float p = -1.0f;
    BSDFType t = eReflected;
    BSDFSample(m, etaI, etaO, Vec3(0.0f), U, W, N, V, L, p, t, rand);
    l[0] = L.x; l[1] = L.y; l[2] = L.z;
    *pdf = p;
    *type = t;
    rngAfter[0] = rand.seed1;
    rngAfter[1] = rand.seed2;
}

// CameraSampler ctor + GenerateRay (src/util.h:49-79)
void ref_generate_ray(const tb200_camera* xc, int width, int height, float x, float y, float* origin, float* dir)
{
    Camera c;
    c.position = Vec3(xc->position[0], xc->position[1], xc->position[2]);
    c.rotation = Quat(xc->rotation[0], xc->rotation[1], xc->rotation[2], xc->rotation[3]);
    CameraSampler sampler(Transform(c.position, c.rotation), xc->fov, 0.001f, 1.0f, width, height);
    Vec3 o, d;
    sampler.GenerateRay(x, y, o, d);
    origin[0] = o.x; origin[1] = o.y; origin[2] = o.z;
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

// Filter::Eval (src/render.h:21-33) with an explicit offset (loader quirk)
float ref_filter_eval(int type, float width, float falloff, float offset, float x, float y)
{
    Filter f((FilterType)type, width, falloff);
    f.offset = offset;
    return f.Eval(x, y);
}

// Trace (src/render.cpp:17-62): closest hit; returns primitive index or -1
int ref_trace(void* h, const float* o, const float* d, float time, float* t, float* n)
{
    RefScene* rs = (RefScene*)h;
    const Primitive* prim = nullptr;
    Vec3 nn;
    float tt = 0.0f;
    bool hit = Trace(rs->scene, Ray(Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]), time), tt, nn, &prim);
    *t = tt;
    n[0] = nn.x; n[1] = nn.y; n[2] = nn.z;
    return hit ? int(prim - &rs->scene.primitives[0]) : -1;
}

// ProbeSample / ProbePdf / Sky::Eval (src/probe.h:205,136; src/scene.h:168)
void ref_probe_sample(void* h, int seed, float* dir, float* color, float* pdf)
{
    RefScene* rs = (RefScene*)h;
    Random rand(seed);
    Vec3 d, c;
    float p = 0.0f;
    ProbeSample(rs->scene.sky.probe, d, c, p, rand);
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    color[0] = c.x; color[1] = c.y; color[2] = c.z;
    *pdf = p;
}

void ref_sky_eval(void* h, const float* d, float* color, float* pdf)
{
    RefScene* rs = (RefScene*)h;
    Vec3 dir(d[0], d[1], d[2]);
    Vec3 c = rs->scene.sky.Eval(dir);
    color[0] = c.x; color[1] = c.y; color[2] = c.z;
    *pdf = rs->scene.sky.probe.valid ? ProbePdf(rs->scene.sky.probe, dir) : 0.0f;
}

// PrimitiveSample + PrimitiveArea (src/intersection.h:855,833)
void ref_primitive_sample(void* h, int prim, float time, int seed, float* pos, float* normal, float* area)
{
    RefScene* rs = (RefScene*)h;
    Random rand(seed);
    Vec3 p, n;
    PrimitiveSample(rs->scene.primitives[prim], time, p, n, rand);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    normal[0] = n.x; normal[1] = n.y; normal[2] = n.z;
    *area = PrimitiveArea(rs->scene.primitives[prim]);
}

}  // extern "C"
