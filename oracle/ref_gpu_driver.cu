// oracle/ref_gpu_driver.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Times the reference's OWN GPU renderer -- GpuRenderer / RenderGpu, src/render.cu:928-1110, the
// one-thread-per-pixel megakernel tinsel ships for Windows -- on the same B200 as the product, as the
// secondary bar SURVEY.md 2b / BASELINE.md 3.3 ask for (bench.py prints it as `gpu_baseline`).
// src/render.cu is compiled where it lies by oracle/Makefile (`make refgpu`) with the reference's
// Release flags (tinsel.vcxproj:134: -O3 -Xptxas -dlcm=cg -lineinfo -DNDEBUG -DCUDA -prec-div=false
// -prec-sqrt=false -ftz=true -use_fast_math) and -arch sm_100a instead of sm_50; this file only
// drives it through the reference's Renderer interface (render.h:66-79).
//
// It is NOT a parity arm: RenderGpu is a different estimator from the CPU renderer the product
// matches (kRayEpsilon 1e-3, its own raster jitter, light-list emission, termination before NEE,
// fast-math) -- only its speed is of interest.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "render.h"   // the reference's (via -I/root/reference/src)

namespace {
cudaEvent_t g_kernelEnd = nullptr;   // recorded in front of every device-to-host copy while timing
bool g_timing = false;
}

// hook for the cudaMemcpy calls of src/render.cu (oracle/ref_gpu_shim.h)
extern "C" cudaError_t tb_ref_memcpy(void* dst, const void* src, size_t n, cudaMemcpyKind kind)
{
    if (g_timing && kind == cudaMemcpyDeviceToHost && g_kernelEnd) cudaEventRecord(g_kernelEnd, 0);
    return cudaMemcpy(dst, src, n, kind);
}

// scene/camera/options: the reference's own objects (oracle/_ref/libtinsel_ref.so, ref_native_*).
// out[0] = wall ms per Render() call (kernel + blocking read-back = what a tinsel user sees),
// out[1] = RenderGpu kernel ms per call (CUDA events on the legacy default stream it launches on).
extern "C" int refgpu_bench(const void* scene, const void* camera, const void* options, int warmup, int calls, float* hostPixels,
                            double* out)
{
    const Options& o = *static_cast<const Options*>(options);
    const Camera& c = *static_cast<const Camera*>(camera);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return -1;
    cudaSetDevice(0);
    Renderer* r = CreateGpuRenderer(static_cast<const Scene*>(scene));
    r->Init(o.width, o.height);
    std::vector<Color> local;
    Color* pixels = reinterpret_cast<Color*>(hostPixels);
    if (!pixels) {
        local.resize(size_t(o.width) * o.height);
        pixels = local.data();
    }
    for (int k = 0; k < warmup; ++k) r->Render(c, o, pixels);
    cudaDeviceSynchronize();
    cudaEvent_t start;
    cudaEventCreate(&start);
    cudaEventCreate(&g_kernelEnd);
    double kernelMs = 0.0;
    g_timing = true;
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < calls; ++k) {
        cudaEventRecord(start, 0);
        r->Render(c, o, pixels);   // launch + blocking cudaMemcpy: synchronous on return
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, start, g_kernelEnd);
        kernelMs += ms;
    }
    const auto t1 = std::chrono::steady_clock::now();
    g_timing = false;
    const int err = cudaGetLastError() == cudaSuccess ? 0 : -2;
    out[0] = std::chrono::duration<double, std::milli>(t1 - t0).count() / calls;
    out[1] = kernelMs / calls;
    cudaEventDestroy(start);
    cudaEventDestroy(g_kernelEnd);
    g_kernelEnd = nullptr;
    delete r;
    return err;
}
