// Scene snapshot (".tsnap") reader/writer.  Plain host C++, no CUDA, no reference headers.
//
// A snapshot is the flat dump of what tinsel's loader hands the renderer: Scene::primitives,
// the scene BVH, every mesh's arrays + BVH + area CDF, the sky (probe pixels), the camera
// and the options.  It exists because the reference's .tin/.obj/.hdr assets and its loader
// are not present on the GPU box; snapshots are produced in the build container by
// oracle/ref_driver.cpp, which links the reference's own LoadTin/Scene::Build.
//
// Only the probe's importance tables are rebuilt here (they would double the file size):
// build_probe_tables() restates Probe::BuildCDF (src/probe.h:31-79) operation for operation
// (sequential fp32 sums, then scaling by the reciprocal row weight / division by the total).
// tests/test_snapshot.py checks the rebuilt tables bit-for-bit against the reference's.
#include "tinsel_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

namespace {

const char kMagic[8] = {'T', 'B', '2', 'S', 'N', 'A', 'P', '1'};

// g_snapError is assigned through snap_error() so that tb200_last_error() can tell which of the two
// sticky messages (this file's, api.cu's) is the more recent one
thread_local std::string g_snapError;
thread_local unsigned long long g_snapStamp = 0;
thread_local unsigned long long g_errorClock = 0;
static void snap_error(const std::string& what)
{
    g_snapError = what;
    g_snapStamp = ++g_errorClock;
}

struct MeshStore {
    std::vector<float> positions, normals, cdf;
    std::vector<int32_t> indices;
    std::vector<tb200_bvh_node> nodes;
};

}  // namespace

struct tb200_snapshot {
    tb200_scene scene;
    tb200_camera camera;
    tb200_options options;
    std::vector<tb200_primitive> primitives;
    std::vector<tb200_bvh_node> bvhNodes;
    std::vector<tb200_mesh> meshes;
    std::vector<MeshStore> meshData;
    std::vector<float> probeData, pdfX, cdfX, pdfY, cdfY;
};

// Probe::BuildCDF, src/probe.h:31-79.  Luminance() is src/maths.h:1571-1574.
static void build_probe_tables(tb200_snapshot* s, int width, int height)
{
    const float* data = s->probeData.data();
    // +1 element of slack: ProbeSample's column search can return col == width when
    // r2 exceeds the row's last cdf value (src/probe.h:217-220), reading one past the row.
    s->pdfX.assign(size_t(width) * height + 1, 0.0f);
    s->cdfX.assign(size_t(width) * height + 1, 0.0f);
    s->pdfY.assign(size_t(height) + 1, 0.0f);
    s->cdfY.assign(size_t(height) + 1, 0.0f);

    float totalWeightY = 0.0f;
    for (int j = 0; j < height; ++j) {
        float totalWeightX = 0.0f;
        for (int i = 0; i < width; ++i) {
            const float* c = data + (size_t(j) * width + i) * 4;
            float weight = c[0] * 0.3f + c[1] * 0.6f + c[2] * 0.1f;
            totalWeightX += weight;
            s->pdfX[size_t(j) * width + i] = weight;
            s->cdfX[size_t(j) * width + i] = totalWeightX;
        }
        float invTotalWeightX = 1.0f / totalWeightX;
        for (int i = 0; i < width; ++i) {
            s->pdfX[size_t(j) * width + i] *= invTotalWeightX;
            s->cdfX[size_t(j) * width + i] *= invTotalWeightX;
        }
        totalWeightY += totalWeightX;
        s->pdfY[j] = totalWeightX;
        s->cdfY[j] = totalWeightY;
    }
    for (int j = 0; j < height; ++j) {
        s->cdfY[j] /= float(totalWeightY);
        s->pdfY[j] /= float(totalWeightY);
    }
}

static bool read_exact(FILE* f, void* dst, size_t bytes)
{
    return bytes == 0 || fread(dst, 1, bytes, f) == bytes;
}

// bytes left between the read position and the end of the file (counts are checked against it before
// any array is sized from them, as tb200_mesh_bin_load does: a corrupt header must not turn into a
// bad_alloc thrown through the C ABI)
static uint64_t bytes_left(FILE* f)
{
    const long here = ftell(f);
    if (here < 0 || fseek(f, 0, SEEK_END) != 0) return 0;
    const long end = ftell(f);
    fseek(f, here, SEEK_SET);
    return end > here ? uint64_t(end - here) : 0;
}

static tb200_snapshot* snapshot_load_impl(FILE* f)
{
    std::unique_ptr<tb200_snapshot> s(new tb200_snapshot());   // freed if a resize throws
    bool ok = true;
    char magic[8];
    uint32_t hdr[6];
    ok = ok && read_exact(f, magic, 8) && memcmp(magic, kMagic, 8) == 0;
    ok = ok && read_exact(f, hdr, sizeof(hdr));
    ok = ok && read_exact(f, &s->camera, sizeof(tb200_camera));
    ok = ok && read_exact(f, &s->options, sizeof(tb200_options));
    float sky[6];
    ok = ok && read_exact(f, sky, sizeof(sky));
    if (ok) {
        const uint64_t left = bytes_left(f);
        ok = uint64_t(hdr[0]) * sizeof(tb200_primitive) + uint64_t(hdr[2]) * sizeof(tb200_bvh_node) +
                 uint64_t(hdr[1]) * 16u <= left;
    }
    if (ok) {
        s->primitives.resize(hdr[0]);
        s->bvhNodes.resize(hdr[2]);
        ok = ok && read_exact(f, s->primitives.data(), hdr[0] * sizeof(tb200_primitive));
        ok = ok && read_exact(f, s->bvhNodes.data(), hdr[2] * sizeof(tb200_bvh_node));
        s->meshes.resize(hdr[1]);
        s->meshData.resize(hdr[1]);
        for (uint32_t m = 0; ok && m < hdr[1]; ++m) {
            int32_t counts[3];
            float area;
            ok = ok && read_exact(f, counts, sizeof(counts)) && read_exact(f, &area, 4);
            if (!ok) break;
            if (counts[0] < 0 || counts[1] < 0 || counts[2] < 0) { ok = false; break; }
            const uint64_t need = uint64_t(counts[0]) * 24u + uint64_t(counts[1]) * 4u +
                                  uint64_t(counts[2]) * sizeof(tb200_bvh_node) + uint64_t(counts[1]) / 3u * 4u;
            if (need > bytes_left(f)) { ok = false; break; }
            MeshStore& d = s->meshData[m];
            d.positions.resize(size_t(counts[0]) * 3);
            d.normals.resize(size_t(counts[0]) * 3);
            d.indices.resize(size_t(counts[1]));
            d.nodes.resize(size_t(counts[2]));
            d.cdf.resize(size_t(counts[1]) / 3);
            ok = ok && read_exact(f, d.positions.data(), d.positions.size() * 4);
            ok = ok && read_exact(f, d.normals.data(), d.normals.size() * 4);
            ok = ok && read_exact(f, d.indices.data(), d.indices.size() * 4);
            ok = ok && read_exact(f, d.nodes.data(), d.nodes.size() * sizeof(tb200_bvh_node));
            ok = ok && read_exact(f, d.cdf.data(), d.cdf.size() * 4);
            tb200_mesh& g = s->meshes[m];
            g.positions = d.positions.data();
            g.normals = d.normals.data();
            g.indices = d.indices.data();
            g.nodes = d.nodes.data();
            g.cdf = d.cdf.data();
            g.numVertices = counts[0];
            g.numIndices = counts[1];
            g.numNodes = counts[2];
            g.area = area;
        }
    }
    memset(&s->scene, 0, sizeof(s->scene));
    if (ok && hdr[3]) {
        const uint64_t n64 = uint64_t(hdr[4]) * hdr[5];
        ok = n64 > 0 && n64 * 12u <= bytes_left(f);
        const size_t n = size_t(n64);
        std::vector<float> rgb(ok ? n * 3 : 0);
        ok = ok && read_exact(f, rgb.data(), rgb.size() * 4);
        if (ok) {
            s->probeData.assign((n + 1) * 4, 0.0f);
            for (size_t i = 0; i < n; ++i) {
                s->probeData[i * 4 + 0] = rgb[i * 3 + 0];
                s->probeData[i * 4 + 1] = rgb[i * 3 + 1];
                s->probeData[i * 4 + 2] = rgb[i * 3 + 2];
                s->probeData[i * 4 + 3] = 0.0f;  // Color(r,g,b): w defaults to 0 (src/maths.h:292)
            }
            build_probe_tables(s.get(), int(hdr[4]), int(hdr[5]));
            s->scene.sky.probeValid = 1;
            s->scene.sky.probeWidth = int(hdr[4]);
            s->scene.sky.probeHeight = int(hdr[5]);
            s->scene.sky.probeData = s->probeData.data();
            s->scene.sky.pdfValuesX = s->pdfX.data();
            s->scene.sky.cdfValuesX = s->cdfX.data();
            s->scene.sky.pdfValuesY = s->pdfY.data();
            s->scene.sky.cdfValuesY = s->cdfY.data();
        }
    }
    if (!ok) return nullptr;
    memcpy(s->scene.sky.horizon, sky, 12);
    memcpy(s->scene.sky.zenith, sky + 3, 12);
    s->scene.primitives = s->primitives.data();
    s->scene.numPrimitives = int32_t(s->primitives.size());
    s->scene.meshes = s->meshes.data();
    s->scene.numMeshes = int32_t(s->meshes.size());
    s->scene.bvhNodes = s->bvhNodes.data();
    s->scene.numBvhNodes = int32_t(s->bvhNodes.size());
    return s.release();
}

extern "C" tb200_snapshot* tb200_snapshot_load(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) {
        snap_error(std::string("cannot open snapshot ") + path);
        return nullptr;
    }
    tb200_snapshot* s = nullptr;
    try {
        s = snapshot_load_impl(f);   // no exception may cross the C ABI
    } catch (const std::exception& e) {
        fclose(f);
        snap_error(std::string("snapshot ") + path + ": " + e.what());
        return nullptr;
    }
    fclose(f);
    if (!s) snap_error(std::string("truncated or malformed snapshot ") + path);
    return s;
}

extern "C" int tb200_snapshot_save(const char* path, const tb200_scene* scene, const tb200_camera* camera,
                                   const tb200_options* options)
{
    FILE* f = fopen(path, "wb");
    if (!f) {
        snap_error(std::string("cannot open for writing ") + path);
        return -1;
    }
    uint32_t hdr[6] = {uint32_t(scene->numPrimitives), uint32_t(scene->numMeshes), uint32_t(scene->numBvhNodes),
                       uint32_t(scene->sky.probeValid ? 1 : 0), uint32_t(scene->sky.probeWidth),
                       uint32_t(scene->sky.probeHeight)};
    if (!scene->sky.probeValid) hdr[4] = hdr[5] = 0;
    fwrite(kMagic, 1, 8, f);
    fwrite(hdr, 1, sizeof(hdr), f);
    fwrite(camera, 1, sizeof(*camera), f);
    fwrite(options, 1, sizeof(*options), f);
    fwrite(scene->sky.horizon, 4, 3, f);
    fwrite(scene->sky.zenith, 4, 3, f);
    fwrite(scene->primitives, sizeof(tb200_primitive), scene->numPrimitives, f);
    fwrite(scene->bvhNodes, sizeof(tb200_bvh_node), scene->numBvhNodes, f);
    for (int m = 0; m < scene->numMeshes; ++m) {
        const tb200_mesh& g = scene->meshes[m];
        int32_t counts[3] = {g.numVertices, g.numIndices, g.numNodes};
        fwrite(counts, 4, 3, f);
        fwrite(&g.area, 4, 1, f);
        fwrite(g.positions, 4, size_t(g.numVertices) * 3, f);
        fwrite(g.normals, 4, size_t(g.numVertices) * 3, f);
        fwrite(g.indices, 4, size_t(g.numIndices), f);
        fwrite(g.nodes, sizeof(tb200_bvh_node), size_t(g.numNodes), f);
        fwrite(g.cdf, 4, size_t(g.numIndices) / 3, f);
    }
    if (scene->sky.probeValid) {
        const size_t n = size_t(scene->sky.probeWidth) * scene->sky.probeHeight;
        std::vector<float> rgb(n * 3);
        for (size_t i = 0; i < n; ++i) {
            rgb[i * 3 + 0] = scene->sky.probeData[i * 4 + 0];
            rgb[i * 3 + 1] = scene->sky.probeData[i * 4 + 1];
            rgb[i * 3 + 2] = scene->sky.probeData[i * 4 + 2];
        }
        fwrite(rgb.data(), 4, rgb.size(), f);
    }
    bool ok = ferror(f) == 0;
    fclose(f);
    if (!ok) snap_error(std::string("write failed for ") + path);
    return ok ? 0 : -1;
}

extern "C" const tb200_scene* tb200_snapshot_scene(const tb200_snapshot* s) { return &s->scene; }
extern "C" const tb200_camera* tb200_snapshot_camera(const tb200_snapshot* s) { return &s->camera; }
extern "C" const tb200_options* tb200_snapshot_options(const tb200_snapshot* s) { return &s->options; }
extern "C" void tb200_snapshot_free(tb200_snapshot* s) { delete s; }

// shared with api.cu through tb200_last_error()
const char* tb200_snapshot_error() { return g_snapError.c_str(); }
unsigned long long tb200_snapshot_error_stamp() { return g_snapStamp; }
unsigned long long tb200_error_tick() { return ++g_errorClock; }

// Per-(pixel, frame) seed for Random(seed) (src/maths.h:1040-1044).  Two rounds of a 32-bit
// bijective mixer: within a frame distinct pixels can never share a seed.  The device copy is
// tb_sample_seed() in tb_common.cuh; tests/test_parity_gpu.py checks they agree.
static inline uint32_t tb_mix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

extern "C" uint32_t tb200_sample_seed(uint32_t pixelIndex, uint32_t frame)
{
    uint32_t h = tb_mix32(pixelIndex + 0x9E3779B9u);
    return tb_mix32(h ^ (frame * 0x85EBCA6Bu + 0xC2B2AE35u));
}


// ---- tinsel binary meshes (src/mesh.cpp:809-880) ----------------------------------------------------
struct tb200_mesh_file {
    std::vector<float> positions, normals, cdf;
    std::vector<int32_t> indices;
    std::vector<tb200_bvh_node> nodes;
    tb200_mesh mesh;
};

extern "C" tb200_mesh_file* tb200_mesh_bin_load(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) {
        snap_error(std::string("cannot open ") + path);
        return nullptr;
    }
    int32_t counts[3] = {0, 0, 0};   // numVertices, numIndices, numNodes
    bool ok = fread(counts, 4, 3, f) == 3;
    // the format has no magic: reject counts the file cannot hold
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 12, SEEK_SET);
    ok = ok && counts[0] >= 0 && counts[1] >= 0 && counts[2] >= 0 && counts[1] % 3 == 0;
    const long long expect = 12ll + 24ll * counts[0] + 4ll * counts[1] + 32ll * counts[2] + 4ll + 4ll * (counts[1] / 3);
    ok = ok && expect == (long long)size;
    if (!ok) {
        fclose(f);
        snap_error(std::string("not a tinsel .bin mesh (size does not match its header): ") + path);
        return nullptr;
    }
    tb200_mesh_file* m = new tb200_mesh_file();
    m->positions.resize(size_t(counts[0]) * 3);
    m->normals.resize(size_t(counts[0]) * 3);
    m->indices.resize(size_t(counts[1]));
    m->nodes.resize(size_t(counts[2]));
    m->cdf.resize(size_t(counts[1]) / 3);
    float area = 0.0f;
    auto rd = [&](void* dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; };
    ok = rd(m->positions.data(), m->positions.size() * 4) && rd(m->normals.data(), m->normals.size() * 4) &&
         rd(m->indices.data(), m->indices.size() * 4) && rd(m->nodes.data(), m->nodes.size() * sizeof(tb200_bvh_node)) &&
         rd(&area, 4) && rd(m->cdf.data(), m->cdf.size() * 4);
    fclose(f);
    if (!ok) {
        delete m;
        snap_error(std::string("short read: ") + path);
        return nullptr;
    }
    m->mesh.positions = m->positions.data();
    m->mesh.normals = m->normals.data();
    m->mesh.indices = m->indices.data();
    m->mesh.nodes = m->nodes.data();
    m->mesh.cdf = m->cdf.data();
    m->mesh.numVertices = counts[0];
    m->mesh.numIndices = counts[1];
    m->mesh.numNodes = counts[2];
    m->mesh.area = area;
    return m;
}

extern "C" const tb200_mesh* tb200_mesh_bin_mesh(const tb200_mesh_file* f) { return f ? &f->mesh : nullptr; }

extern "C" int tb200_mesh_bin_save(const char* path, const tb200_mesh* g)
{
    if (!path || !g) {
        snap_error("tb200_mesh_bin_save: null argument");
        return -1;
    }
    FILE* f = fopen(path, "wb");
    if (!f) {
        snap_error(std::string("cannot open for writing ") + path);
        return -1;
    }
    const int32_t counts[3] = {g->numVertices, g->numIndices, g->numNodes};
    bool ok = fwrite(counts, 4, 3, f) == 3;
    auto wr = [&](const void* src, size_t bytes) { return bytes == 0 || fwrite(src, 1, bytes, f) == bytes; };
    ok = ok && wr(g->positions, size_t(g->numVertices) * 12) && wr(g->normals, size_t(g->numVertices) * 12) &&
         wr(g->indices, size_t(g->numIndices) * 4) && wr(g->nodes, size_t(g->numNodes) * sizeof(tb200_bvh_node)) &&
         wr(&g->area, 4) && wr(g->cdf, size_t(g->numIndices) / 3 * 4);
    ok = (fclose(f) == 0) && ok;
    if (!ok) snap_error(std::string("write failed: ") + path);
    return ok ? 0 : -1;
}

extern "C" void tb200_mesh_bin_free(tb200_mesh_file* f) { delete f; }
