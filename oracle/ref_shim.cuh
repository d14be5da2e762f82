// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Force-included (nvcc -include) when the reference's own src/odom/reduce.cu and
// src/odom/utils.cu are compiled, unmodified and where they lie under
// /root/reference, for sm_100a with CUDA 12.9 (see oracle/build_ref.py).  The
// reference targets CUDA 10.2 / Kepler-Maxwell, so exactly two removed-API gaps
// have to be bridged; nothing here changes any arithmetic:
//
//  1. `__shfl_down(v, offset)` (no `_sync`) is undeclared for sm_70+ in CUDA 12
//     (first use: src/odom/reduce.cu:94).  We declare overloads that forward to
//     `__shfl_down_sync(0xffffffff, ...)`; all 32 lanes are converged at every
//     call site (warpReduceSum is called from block-uniform code).
//  2. Legacy texture references were removed in CUDA 12
//     (src/odom/utils.cu:548-577: `texture<uchar4,2,...> inTex`, `tex2D(inTex,x,y)`,
//     `cudaBindTextureToArray`, `cudaUnbindTexture`).  We provide a dummy
//     `texture<>` type and route `tex2D(inTex, x, y)` to a plain linear uchar4
//     image whose pointer/width the harness "binds"; the harness passes its
//     linear device pointer disguised as the `cudaArray*` argument of
//     `imageBGRToIntensity`.  The reference's own arithmetic at utils.cu:560 runs.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <sstream>
#include <iostream>
#include <cstring>

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 700
__device__ __forceinline__ float __shfl_down(float v, int offset, int width = 32)
{
    return __shfl_down_sync(0xffffffffu, v, offset, width);
}
__device__ __forceinline__ int __shfl_down(int v, int offset, int width = 32)
{
    return __shfl_down_sync(0xffffffffu, v, offset, width);
}
#endif

// ---- legacy texture reference emulation (linear uchar4 image) ----
template <class T, int Dim, cudaTextureReadMode Mode>
struct texture
{
};

struct RefShimImage
{
    const uchar4 * data;
    int width;   // in pixels; rows are dense
};

// One definition per translation unit that includes the shim; only utils.cu uses it.
static __device__ RefShimImage g_ref_shim_image;
// Width of the image about to be bound; set by the harness (ref_shim_set_width).
extern int g_ref_shim_bind_width;

#define tex2D(texref, x, y) (g_ref_shim_image.data[(y) * g_ref_shim_image.width + (x)])

template <class Tex>
static inline cudaError_t cudaBindTextureToArray(const Tex &, cudaArray * disguisedLinearPtr)
{
    RefShimImage img;
    img.data = reinterpret_cast<const uchar4 *>(disguisedLinearPtr);
    img.width = g_ref_shim_bind_width;
    return cudaMemcpyToSymbol(g_ref_shim_image, &img, sizeof(img));
}

template <class Tex>
static inline cudaError_t cudaUnbindTexture(const Tex &)
{
    return cudaSuccess;
}
