// Shared host/device helpers for libslam_odom: error plumbing, launch geometry,
// single-pass grid reduction epilogue.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include "../../include/slam_odom.h"
#include "pixel_ops.cuh"

namespace slam {

void set_last_error(const std::string & msg);

#define SLAM_CUDA_TRY(expr)                                                                           \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t _e = (expr);                                                                      \
        if(_e != cudaSuccess)                                                                         \
        {                                                                                             \
            ::slam::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return SLAM_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while(0)

#define SLAM_ARG_CHECK(cond)                                                   \
    do                                                                         \
    {                                                                          \
        if(!(cond))                                                            \
        {                                                                      \
            ::slam::set_last_error(std::string("bad argument: ") + #cond);    \
            return SLAM_ERR_ARG;                                               \
        }                                                                      \
    } while(0)

__host__ __device__ static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Workspace of the single-launch reductions: partial sums of up to kMaxBlocks blocks,
// 32 words each, then the ticket counter.
constexpr int kMaxReduceBlocks = 2048;
constexpr int kPartialStride = 32;
constexpr size_t kWorkspaceBytes = (size_t)kMaxReduceBlocks * kPartialStride * 4 + 256;

__device__ __forceinline__ unsigned * workspace_ticket(void * ws)
{
    return reinterpret_cast<unsigned *>(reinterpret_cast<char *>(ws) + (size_t)kMaxReduceBlocks * kPartialStride * 4);
}

// Epilogue of every reduction kernel: block sum, publish the partial, and let the block that
// draws the last ticket fold all partials in block order.  One launch, deterministic result.
template <typename T, int NV>
__device__ __forceinline__ void grid_finish(T (&acc)[NV], void * workspace, T * out)
{
    static_assert(NV <= kPartialStride, "too many values");
    __shared__ T smem[32 * NV];
    __shared__ bool is_last;
    T * partials = reinterpret_cast<T *>(workspace);
    unsigned * ticket = workspace_ticket(workspace);

    const T tot = block_sum<T, NV>(acc, smem);
    if(threadIdx.x < NV) partials[blockIdx.x * kPartialStride + threadIdx.x] = tot;
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if(is_last)
    {
        __threadfence();
        const int lane = threadIdx.x & 31;
        const int wid = threadIdx.x >> 5;
        const int nw = blockDim.x >> 5;
        T s = 0;
        if(lane < NV)
            for(int b = wid; b < (int)gridDim.x; b += nw) s += __ldcg(partials + b * kPartialStride + lane);
        if(lane < NV) smem[wid * NV + lane] = s;
        __syncthreads();
        if(threadIdx.x < NV)
        {
            T total = 0;
            for(int w = 0; w < nw; w++) total += smem[w * NV + threadIdx.x];
            out[threadIdx.x] = total;
        }
        if(threadIdx.x == 0) *ticket = 0u;   // ready for the next launch
    }
}

}   // namespace slam
