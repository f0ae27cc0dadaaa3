// tinsel_plugin.cpp -- the reference-facing C++ adapter: defines tinsel's
//     Renderer* CreateGpuWavefrontRenderer(const Scene* s);            (src/render.h:78)
// on top of the C ABI (include/tinsel_b200.h).  The reference declares this factory but defines it
// only in src/wavefront.cu, which is in no build (tinsel.vcxproj:83-102,130; makefile:9) and does
// not compile with nvcc 12.9, so every shipped tinsel binary has this symbol undefined; linking
// this object (or libtinsel_b200_plugin.so) fills the hole without touching main.cpp or the loader.
//
// Compiled against the reference's OWN headers (-I<tinsel>/src, never copied into this repo):
// Scene/Primitive/Material/Camera/Options/Color layouts are theirs, the vtable is theirs.
// Semantics follow the reference GPU renderers (render.cu:1056-1103): the device keeps the running
// sums since Init() and Render() overwrites the caller's buffer with them -- equivalent to the CPU
// renderer's "+=" because the caller zeroes its buffer whenever it calls Init (main.cpp:73-88).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "render.h"   // the reference's src/render.h

#include "tinsel_b200.h"
#include "tinsel_b200_plugin.h"

static_assert(sizeof(BVHNode) == sizeof(tb200_bvh_node), "BVHNode layout (bvh.h:9-21)");
static_assert(sizeof(Color) == 16 && sizeof(Vec3) == 12, "vector layouts (maths.h)");
static_assert(sizeof(Camera) == sizeof(tb200_camera), "Camera layout (scene.h:11-31)");
static_assert(sizeof(Options) == sizeof(tb200_options), "Options layout (render.h:50-63)");

namespace {

void copy_transform(const Transform& t, tb200_transform* o)
{
    o->p[0] = t.p.x; o->p[1] = t.p.y; o->p[2] = t.p.z;
    o->r[0] = t.r.x; o->r[1] = t.r.y; o->r[2] = t.r.z; o->r[3] = t.r.w;
    o->s = t.s;
}

struct GpuWavefrontRenderer : public Renderer {
    tb200_renderer* impl = nullptr;
    std::vector<tb200_primitive> prims;
    std::vector<tb200_mesh> meshes;
    std::vector<unsigned long> meshIds;

    explicit GpuWavefrontRenderer(const Scene* s)
    {
        // Scene -> tb200_scene: primitives by value, meshes deduplicated by MeshGeometry::id
        // (util.h:20; the reference GPU renderer does the same, render.cu:1000-1011)
        for (size_t i = 0; i < s->primitives.size(); ++i) {
            const Primitive& p = s->primitives[i];
            tb200_primitive x;
            memset(&x, 0, sizeof(x));
            copy_transform(p.startTransform, &x.start);
            copy_transform(p.endTransform, &x.end);
            x.type = p.type;
            x.mesh = -1;
            if (p.type == eSphere) x.radius = p.sphere.radius;
            if (p.type == ePlane) memcpy(x.plane, p.plane.plane, sizeof(x.plane));
            if (p.type == eMesh) {
                int found = -1;
                for (size_t m = 0; m < meshIds.size(); ++m)
                    if (meshIds[m] == p.mesh.id) found = int(m);
                if (found < 0) {
                    tb200_mesh g;
                    g.positions = reinterpret_cast<const float*>(p.mesh.positions);
                    g.normals = reinterpret_cast<const float*>(p.mesh.normals);
                    g.indices = reinterpret_cast<const int32_t*>(p.mesh.indices);
                    g.nodes = reinterpret_cast<const tb200_bvh_node*>(p.mesh.nodes);
                    g.cdf = p.mesh.cdf;
                    g.numVertices = p.mesh.numVertices;
                    g.numIndices = p.mesh.numIndices;
                    g.numNodes = p.mesh.numNodes;
                    g.area = p.mesh.area;
                    found = int(meshes.size());
                    meshes.push_back(g);
                    meshIds.push_back(p.mesh.id);
                }
                x.mesh = found;
            }
            const Material& m = p.material;
            memcpy(x.material.emission, &m.emission, 12);
            memcpy(x.material.color, &m.color, 12);
            memcpy(x.material.absorption, &m.absorption, 12);
            x.material.eta = m.eta;
            x.material.metallic = m.metallic;
            x.material.subsurface = m.subsurface;
            x.material.specular = m.specular;
            x.material.roughness = m.roughness;
            x.material.specularTint = m.specularTint;
            x.material.anisotropic = m.anisotropic;
            x.material.sheen = m.sheen;
            x.material.sheenTint = m.sheenTint;
            x.material.clearcoat = m.clearcoat;
            x.material.clearcoatGloss = m.clearcoatGloss;
            x.material.transmission = m.transmission;
            x.lightSamples = p.lightSamples;
            prims.push_back(x);
        }
        tb200_scene xs;
        memset(&xs, 0, sizeof(xs));
        xs.primitives = prims.data();
        xs.numPrimitives = int(prims.size());
        xs.meshes = meshes.data();
        xs.numMeshes = int(meshes.size());
        xs.bvhNodes = reinterpret_cast<const tb200_bvh_node*>(s->bvh.nodes);
        xs.numBvhNodes = s->bvh.numNodes;
        memcpy(xs.sky.horizon, &s->sky.horizon, 12);
        memcpy(xs.sky.zenith, &s->sky.zenith, 12);
        if (s->sky.probe.valid) {
            const Probe& pr = s->sky.probe;
            xs.sky.probeValid = 1;
            xs.sky.probeWidth = pr.width;
            xs.sky.probeHeight = pr.height;
            xs.sky.probeData = reinterpret_cast<const float*>(pr.data);
            xs.sky.pdfValuesX = pr.pdfValuesX;
            xs.sky.cdfValuesX = pr.cdfValuesX;
            xs.sky.pdfValuesY = pr.pdfValuesY;
            xs.sky.cdfValuesY = pr.cdfValuesY;
        }
        // Options cannot change (render.h:50-63), so renderer-private knobs are environment variables:
        //   TINSEL_GPUS=N          spread the renderer over CUDA devices 0..N-1 of this box (row slabs,
        //                          every GPU copies its own rows to the host; tb200_create_multi)
        //   TINSEL_B200_DEVICE=k   single-device ordinal (default 0)
        //   TINSEL_B200_PIN=0      never page-lock the caller's frame buffer (see Render below)
        const char* dev = getenv("TINSEL_B200_DEVICE");
        const char* gpus = getenv("TINSEL_GPUS");
        const int n = gpus ? atoi(gpus) : 1;
        if (n > 1) {
            std::vector<int> devices;
            const bool oneGpuBox = getenv("TINSEL_B200_TEST_DUP_DEVICES") != nullptr;   // test hook: all members on device 0
            for (int k = 0; k < n; ++k) devices.push_back(oneGpuBox ? 0 : k);
            impl = tb200_create_multi(&xs, devices.data(), n);
        } else {
            impl = tb200_create(&xs, dev ? atoi(dev) : 0);
        }
        const char* pin = getenv("TINSEL_B200_PIN");
        allowPin = !(pin && atoi(pin) == 0);
        if (!impl) fprintf(stderr, "CreateGpuWavefrontRenderer: %s\n", tb200_last_error());
    }

    ~GpuWavefrontRenderer() override { tb200_destroy(impl); }   // unpins

    // Frame-buffer pinning.  The C ABI never page-locks caller memory by itself; this adapter does, on
    // the strength of tinsel's own contract for `output`: main.cpp owns ONE frame buffer, g_pixels,
    // passes it to every Render() (main.cpp:249) and reallocates it only in InitFrameBuffer(), which
    // calls Init() straight afterwards (main.cpp:73-88) and before any further Render().  So: a buffer
    // seen in two consecutive Render() calls is pinned; Init() -- the one notification a reallocation
    // produces -- unpins; a Render() with a different pointer unpins first.  Between the application's
    // delete[] and its Init() call the stale registration is never used for a copy.
    bool allowPin = true;
    Color* lastOutput = nullptr;
    Color* pinned = nullptr;

    void Init(int width, int height) override
    {
        if (!impl) return;
        tb200_unpin_output(impl);
        pinned = lastOutput = nullptr;
        if (tb200_init(impl, width, height) != 0) fprintf(stderr, "GpuWavefrontRenderer::Init: %s\n", tb200_last_error());
    }

    // The reference has no error channel (void, no exceptions): failures are logged, `output` is
    // left untouched and the sticky message stays readable through tb200_last_error().
    void Render(const Camera& c, const Options& options, Color* output) override
    {
        if (!impl) return;
        if (pinned && pinned != output) {
            tb200_unpin_output(impl);
            pinned = nullptr;
        }
        if (allowPin && !pinned && output == lastOutput &&
            tb200_pin_output(impl, reinterpret_cast<float*>(output), size_t(options.width) * options.height * sizeof(Color)) == 0)
            pinned = output;
        lastOutput = output;
        const tb200_camera* xc = reinterpret_cast<const tb200_camera*>(&c);      // identical layouts (static_assert above)
        const tb200_options* xo = reinterpret_cast<const tb200_options*>(&options);
        if (tb200_render(impl, xc, xo, reinterpret_cast<float*>(output)) != 0)
            fprintf(stderr, "GpuWavefrontRenderer::Render: %s\n", tb200_last_error());
    }
};

}  // namespace

Renderer* CreateGpuWavefrontRenderer(const Scene* s) { return new GpuWavefrontRenderer(s); }

// Optional fast paths (tinsel_b200_plugin.h) for a host that wants fewer read-backs than the
// Renderer interface allows; both return false -- and do nothing -- for any other Renderer.
bool TinselB200RenderN(Renderer* r, const Camera& c, const Options& options, int n, Color* output)
{
    GpuWavefrontRenderer* g = dynamic_cast<GpuWavefrontRenderer*>(r);
    if (!g || !g->impl) return false;
    if (tb200_render_n(g->impl, reinterpret_cast<const tb200_camera*>(&c), reinterpret_cast<const tb200_options*>(&options), n,
                       reinterpret_cast<float*>(output)) != 0) {
        fprintf(stderr, "TinselB200RenderN: %s\n", tb200_last_error());
        return false;
    }
    return true;
}

bool TinselB200Finish(Renderer* r, const Options& options, Color* filtered, unsigned char* rgb8)
{
    GpuWavefrontRenderer* g = dynamic_cast<GpuWavefrontRenderer*>(r);
    if (!g || !g->impl) return false;
    if (tb200_finish(g->impl, options.exposure, options.limit, reinterpret_cast<float*>(filtered), rgb8) != 0) {
        fprintf(stderr, "TinselB200Finish: %s\n", tb200_last_error());
        return false;
    }
    return true;
}

// C shim so that tests can drive the C++ factory through ctypes: Create -> Init -> spp x Render -> delete.
extern "C" int tb200_plugin_render(const void* scene, const void* camera, const void* options, int spp, float* output)
{
    const Options& o = *static_cast<const Options*>(options);
    Renderer* r = CreateGpuWavefrontRenderer(static_cast<const Scene*>(scene));
    r->Init(o.width, o.height);
    for (int k = 0; k < spp; ++k) r->Render(*static_cast<const Camera*>(camera), o, reinterpret_cast<Color*>(output));
    const char* err = tb200_last_error();
    const int rc = (err && err[0]) ? -1 : 0;
    delete r;
    return rc;
}

bool TinselB200Nlm(Renderer* r, float falloff, int radius, Color* out)
{
    GpuWavefrontRenderer* g = dynamic_cast<GpuWavefrontRenderer*>(r);
    if (!g || !g->impl) return false;
    if (tb200_nlm(g->impl, falloff, radius, reinterpret_cast<float*>(out)) != 0) {
        fprintf(stderr, "TinselB200Nlm: %s\n", tb200_last_error());
        return false;
    }
    return true;
}

// Same, through the fast paths: Create -> Init -> RenderN(spp) -> Finish -> delete.
extern "C" int tb200_plugin_present(const void* scene, const void* camera, const void* options, int spp, float* output,
                                    float* filtered, unsigned char* rgb8)
{
    const Options& o = *static_cast<const Options*>(options);
    Renderer* r = CreateGpuWavefrontRenderer(static_cast<const Scene*>(scene));
    r->Init(o.width, o.height);
    bool ok = TinselB200RenderN(r, *static_cast<const Camera*>(camera), o, spp, reinterpret_cast<Color*>(output));
    ok = ok && TinselB200Finish(r, o, reinterpret_cast<Color*>(filtered), rgb8);
    delete r;
    return ok ? 0 : -1;
}
