// oracle/ref_gpu_shim.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Force-included in front of the reference's src/render.cu (compiled where it lies, unmodified) when
// oracle/Makefile builds oracle/_ref/libtinsel_ref_gpu.so: it routes the file's cudaMemcpy calls
// through a hook so that the benchmark can separate RenderGpu's kernel time from the blocking
// device-to-host copy that ends every GpuRenderer::Render (render.cu:1099-1102).  Nothing else changes.
#pragma once
#include <cuda_runtime.h>
extern "C" cudaError_t tb_ref_memcpy(void* dst, const void* src, size_t n, cudaMemcpyKind kind);
#define cudaMemcpy tb_ref_memcpy
