// The four map-reduce operators of the tracker as single-launch sm_100a kernels
// (operator-level C ABI, include/slam_odom.h).  They replace the reference's
// kernel + reduceSum<<<1,512>>> + cudaDeviceSynchronize + D2H pairs:
//   icpStep            src/odom/reduce.cu:257-490
//   computeRgbResidual src/odom/reduce.cu:739-935
//   rgbStep            src/odom/reduce.cu:494-678
//   so3Step            src/odom/reduce.cu:937-1140
// Differences by design: vectorised (float4) loads of the planar maps, warp-shuffle +
// block reduction with a last-ticket fold in the same launch, results left in device
// memory, no cudaMalloc/sync/host copy per call.
#include "common.cuh"

namespace slam {

constexpr int kReduceThreads = 256;

// ---------------------------------------------------------------- ICP
template <int PX>
__global__ void __launch_bounds__(kReduceThreads) k_icp_step(const IcpArgs a, void * workspace, float * out29)
{
    float acc[29];
#pragma unroll
    for(int k = 0; k < 29; k++) acc[k] = 0.f;

    const int plane = a.rows * a.cols;
    const int nitems = plane / PX;

    for(int item = blockIdx.x * blockDim.x + threadIdx.x; item < nitems; item += gridDim.x * blockDim.x)
    {
        const int p = item * PX;
        float vx[PX], vy[PX], vz[PX], nx[PX], ny[PX], nz[PX];
        if constexpr(PX == 4)
        {
            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(a.vcurr + p));
            const float4 a1 = __ldg(reinterpret_cast<const float4 *>(a.vcurr + plane + p));
            const float4 a2 = __ldg(reinterpret_cast<const float4 *>(a.vcurr + 2 * plane + p));
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(a.ncurr + p));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(a.ncurr + plane + p));
            const float4 b2 = __ldg(reinterpret_cast<const float4 *>(a.ncurr + 2 * plane + p));
            vx[0] = a0.x; vx[1] = a0.y; vx[2] = a0.z; vx[3] = a0.w;
            vy[0] = a1.x; vy[1] = a1.y; vy[2] = a1.z; vy[3] = a1.w;
            vz[0] = a2.x; vz[1] = a2.y; vz[2] = a2.z; vz[3] = a2.w;
            nx[0] = b0.x; nx[1] = b0.y; nx[2] = b0.z; nx[3] = b0.w;
            ny[0] = b1.x; ny[1] = b1.y; ny[2] = b1.z; ny[3] = b1.w;
            nz[0] = b2.x; nz[1] = b2.y; nz[2] = b2.z; nz[3] = b2.w;
        }
        else
        {
            vx[0] = __ldg(a.vcurr + p);
            vy[0] = __ldg(a.vcurr + plane + p);
            vz[0] = __ldg(a.vcurr + 2 * plane + p);
            nx[0] = __ldg(a.ncurr + p);
            ny[0] = __ldg(a.ncurr + plane + p);
            nz[0] = __ldg(a.ncurr + 2 * plane + p);
        }
#pragma unroll
        for(int k = 0; k < PX; k++)
        {
            float row[7];
            const bool found = icp_pixel(a, make_float3(vx[k], vy[k], vz[k]), make_float3(nx[k], ny[k], nz[k]), row);
            accumulate_se3(acc, row, found);
        }
    }
    grid_finish<float, 29>(acc, workspace, out29);
}

// ---------------------------------------------------------------- RGB residual
__global__ void __launch_bounds__(kReduceThreads) k_rgb_residual(const ResidualArgs a, Corres * corres, void * workspace, int * out2)
{
    int acc[2] = {0, 0};
    const int N = a.rows * a.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int i = k / a.cols;
        const int j0 = k - i * a.cols;
        Corres c;
        c.zx = c.zy = c.ox = c.oy = 0;
        c.diff = 0.f;
        c.valid = 0;
        if(rgb_candidate(a, j0, i) && rgb_associate(a, j0, i, c))
        {
            acc[0] += 1;
            acc[1] += (int)(c.diff * c.diff);
        }
        reinterpret_cast<int4 *>(corres)[k] = *reinterpret_cast<const int4 *>(&c);
    }
    grid_finish<int, 2>(acc, workspace, out2);
}

// ---------------------------------------------------------------- RGB step
__global__ void __launch_bounds__(kReduceThreads) k_rgb_step(const RgbStepArgs a, const Corres * corres, void * workspace, float * out29)
{
    float acc[29];
#pragma unroll
    for(int k = 0; k < 29; k++) acc[k] = 0.f;
    const int N = a.rows * a.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(corres) + k);
        const Corres c = *reinterpret_cast<const Corres *>(&raw);
        if(c.valid & 0xff)
        {
            float row[7];
            rgb_row(a, c, row);
            accumulate_se3(acc, row, true);
        }
    }
    grid_finish<float, 29>(acc, workspace, out29);
}

// ---------------------------------------------------------------- SO3
__global__ void __launch_bounds__(kReduceThreads) k_so3_step(const So3Args a, void * workspace, float * out11)
{
    float acc[11];
#pragma unroll
    for(int k = 0; k < 11; k++) acc[k] = 0.f;
    const int N = a.rows * a.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int y = k / a.cols;
        const int x = k - y * a.cols;
        float row[4];
        const bool found = so3_pixel(a, x, y, row);
        accumulate_so3(acc, row, found);
    }
    grid_finish<float, 11>(acc, workspace, out11);
}

static inline int reduce_grid(int nitems)
{
    int g = div_up(nitems, kReduceThreads);
    if(g < 1) g = 1;
    if(g > kMaxReduceBlocks) g = kMaxReduceBlocks;
    return g;
}

// Host-side launchers used by both the operator C ABI and the host-stepped loop.
int launch_icp_step(const IcpArgs & a, void * workspace, float * out29, cudaStream_t s)
{
    const int plane = a.rows * a.cols;
    const bool vec = (plane % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.vcurr) | reinterpret_cast<uintptr_t>(a.ncurr)) % 16 == 0);
    if(vec)
        k_icp_step<4><<<reduce_grid(plane / 4), kReduceThreads, 0, s>>>(a, workspace, out29);
    else
        k_icp_step<1><<<reduce_grid(plane), kReduceThreads, 0, s>>>(a, workspace, out29);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

// ---------------------------------------------------------------- pose-hypothesis scoring (BASELINE.json configs[4])
// One launch scores n candidate poses of the current frame against the model prediction: for each hypothesis the ICP
// association of every pixel (reduce.cu:257-345, same arithmetic as icpStep) and the two sums the reference's acceptance
// test reads, residual = sum of squared point-to-plane distances (JtJJtrSE3::residual) and the inlier count.
// grid = (pixel blocks, hypotheses); the maps are shared by all hypotheses and stay in L2.  Per-hypothesis last-ticket fold
// in block order: deterministic.
constexpr int kScoreStride = 2;
// best_key != nullptr: the block that folds a hypothesis also turns its two sums into the reference's acceptance statistic
// lastICPError = sqrt(residual) / inliers (RGBDOdometryef.cpp:509; +inf below min_inliers, the guard of lc/Ferns.cpp:275-279) and
// folds (error bits << 32 | global hypothesis index) into *best_key with a 64-bit atomicMin: non-negative floats order like their
// bit patterns, so the smallest key is the smallest error, ties to the smaller index -- on one GPU and, after a min-all-reduce of
// that one word, on any number of them.
__global__ void __launch_bounds__(kReduceThreads) k_score_poses(const IcpArgs a0, const float * __restrict__ poses12, float * partials, unsigned * tickets,
                                                                float * out2, unsigned long long * best_key, const int index_base, const float min_inliers)
{
    __shared__ float s_red[2][kReduceThreads / 32];
    __shared__ int s_last;
    const int hyp = blockIdx.y;
    IcpArgs a = a0;
    const float * q = poses12 + 12 * hyp;
    a.Rcurr.r0 = make_float3(__ldg(q + 0), __ldg(q + 1), __ldg(q + 2));
    a.Rcurr.r1 = make_float3(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5));
    a.Rcurr.r2 = make_float3(__ldg(q + 6), __ldg(q + 7), __ldg(q + 8));
    a.tcurr = make_float3(__ldg(q + 9), __ldg(q + 10), __ldg(q + 11));
    const int plane = a.rows * a.cols;
    float res = 0.f, cnt = 0.f;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < plane; k += gridDim.x * blockDim.x)
    {
        float row[7];
        const bool found = icp_pixel(a, make_float3(__ldg(a.vcurr + k), __ldg(a.vcurr + plane + k), __ldg(a.vcurr + 2 * plane + k)),
                                     make_float3(__ldg(a.ncurr + k), __ldg(a.ncurr + plane + k), __ldg(a.ncurr + 2 * plane + k)), row);
        res += row[6] * row[6];
        cnt += found ? 1.f : 0.f;
    }
    res = warp_sum(res);
    cnt = warp_sum(cnt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if(lane == 0)
    {
        s_red[0][wid] = res;
        s_red[1][wid] = cnt;
    }
    __syncthreads();
    float * rows = partials + (size_t)hyp * gridDim.x * kScoreStride;
    if(threadIdx.x < 2)
    {
        float t = 0.f;
        for(int w = 0; w < kReduceThreads / 32; w++) t += s_red[threadIdx.x][w];
        rows[blockIdx.x * kScoreStride + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0) s_last = (atomicAdd(tickets + hyp, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if(!s_last) return;
    __threadfence();
    if(threadIdx.x < 32)
    {
        float t = 0.f;
        if(threadIdx.x < 2)
        {
            for(int b = 0; b < (int)gridDim.x; b++) t += __ldcg(rows + b * kScoreStride + threadIdx.x);
            out2[hyp * 2 + threadIdx.x] = t;
        }
        const float cnt = __shfl_sync(0xffffffffu, t, 1);
        if(best_key && threadIdx.x == 0)
        {
            float err = __fdiv_rn(__fsqrt_rn(t), cnt);   // float32 sqrt and divide, correctly rounded: what numpy computes from the same sums
            if(!(cnt >= min_inliers) || isnan(err)) err = __int_as_float(0x7f800000);
            const unsigned long long key = ((unsigned long long)__float_as_uint(err) << 32) | (unsigned long long)(unsigned)(index_base + hyp);
            atomicMin(best_key, key);
        }
    }
    if(threadIdx.x == 0) tickets[hyp] = 0u;
}

int score_blocks(int plane) { int g = div_up(plane, kReduceThreads * 4); return g < 1 ? 1 : (g > 64 ? 64 : g); }

size_t score_workspace_bytes(int n, int plane) { return (size_t)n * score_blocks(plane) * kScoreStride * 4 + (size_t)n * 4 + (size_t)n * 8 + 256; }

int launch_score_poses(const IcpArgs & a, const float * poses12, int n, void * workspace, float ** out2_dev, cudaStream_t s, unsigned long long * best_key,
                       int index_base, float min_inliers)
{
    const int plane = a.rows * a.cols;
    const int nb = score_blocks(plane);
    float * partials = reinterpret_cast<float *>(workspace);
    unsigned * tickets = reinterpret_cast<unsigned *>(partials + (size_t)n * nb * kScoreStride);
    float * out2 = reinterpret_cast<float *>(tickets + n);
    k_score_poses<<<dim3(nb, n), kReduceThreads, 0, s>>>(a, poses12, partials, tickets, out2, best_key, index_base, min_inliers);
    SLAM_CUDA_TRY(cudaGetLastError());
    *out2_dev = out2;
    return SLAM_OK;
}

int launch_rgb_residual(const ResidualArgs & a, Corres * corres, void * workspace, int * out2, cudaStream_t s)
{
    k_rgb_residual<<<reduce_grid(a.rows * a.cols), kReduceThreads, 0, s>>>(a, corres, workspace, out2);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_rgb_step(const RgbStepArgs & a, const Corres * corres, void * workspace, float * out29, cudaStream_t s)
{
    k_rgb_step<<<reduce_grid(a.rows * a.cols), kReduceThreads, 0, s>>>(a, corres, workspace, out29);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_so3_step(const So3Args & a, void * workspace, float * out11, cudaStream_t s)
{
    k_so3_step<<<reduce_grid(a.rows * a.cols), kReduceThreads, 0, s>>>(a, workspace, out11);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

}   // namespace slam

// ------------------------------------------------------------------ C ABI (operators)
using namespace slam;

extern "C" size_t slam_op_workspace_bytes(void) { return kWorkspaceBytes; }

extern "C" int slam_op_icp_step(const float * Rcurr9, const float * tcurr3, const float * vmap_curr, const float * nmap_curr,
                                const float * Rprev_inv9, const float * tprev3, float fx, float fy, float cx, float cy,
                                const float * vmap_g_prev, const float * nmap_g_prev, float dist_thresh, float angle_thresh,
                                int rows, int cols, void * workspace, float * out29, void * stream)
{
    SLAM_ARG_CHECK(Rcurr9 && tcurr3 && vmap_curr && nmap_curr && Rprev_inv9 && tprev3 && vmap_g_prev && nmap_g_prev && workspace && out29);
    SLAM_ARG_CHECK(rows > 0 && cols > 0);
    IcpArgs a;
    a.Rcurr = mat3_from(Rcurr9);
    a.tcurr = make_float3(tcurr3[0], tcurr3[1], tcurr3[2]);
    a.Rprev_inv = mat3_from(Rprev_inv9);
    a.tprev = make_float3(tprev3[0], tprev3[1], tprev3[2]);
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.distThres = dist_thresh;
    a.angleThres = angle_thresh;
    a.cols = cols; a.rows = rows;
    a.vcurr = vmap_curr; a.ncurr = nmap_curr; a.vprev = vmap_g_prev; a.nprev = nmap_g_prev;
    return launch_icp_step(a, workspace, out29, (cudaStream_t)stream);
}

extern "C" int slam_op_compute_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, const float * last_depth,
                                            const float * next_depth, const uint8_t * last_image, const uint8_t * next_image,
                                            void * corres_img16, float max_depth_delta, const float * kt3, const float * krkinv9,
                                            int rows, int cols, void * workspace, int * out_count_sigma2, void * stream)
{
    SLAM_ARG_CHECK(dIdx && dIdy && last_depth && next_depth && last_image && next_image && corres_img16 && kt3 && krkinv9 && workspace && out_count_sigma2);
    SLAM_ARG_CHECK(rows > 0 && cols > 0);
    ResidualArgs a;
    a.minScale = min_scale;
    a.dIdx = dIdx; a.dIdy = dIdy;
    a.lastDepth = last_depth; a.nextDepth = next_depth;
    a.lastImage = last_image; a.nextImage = next_image;
    a.maxDepthDelta = max_depth_delta;
    a.kt = make_float3(kt3[0], kt3[1], kt3[2]);
    a.krkinv = mat3_from(krkinv9);
    a.cols = cols; a.rows = rows;
    return launch_rgb_residual(a, reinterpret_cast<Corres *>(corres_img16), workspace, out_count_sigma2, (cudaStream_t)stream);
}

extern "C" int slam_op_rgb_step(const void * corres_img16, float sigma, const float * cloud3, float fx, float fy, const int16_t * dIdx,
                                const int16_t * dIdy, float sobel_scale, int rows, int cols, void * workspace, float * out29,
                                void * stream)
{
    SLAM_ARG_CHECK(corres_img16 && cloud3 && dIdx && dIdy && workspace && out29);
    SLAM_ARG_CHECK(rows > 0 && cols > 0);
    RgbStepArgs a;
    a.sigma = sigma;
    a.fx = fx; a.fy = fy;
    a.sobelScale = sobel_scale;
    a.cols = cols; a.rows = rows;
    a.dIdx = dIdx; a.dIdy = dIdy;
    a.lastDepth = nullptr;
    a.invFx = a.invFy = a.cx = a.cy = 0.f;
    a.cloud = cloud3;
    return launch_rgb_step(a, reinterpret_cast<const Corres *>(corres_img16), workspace, out29, (cudaStream_t)stream);
}

extern "C" int slam_op_so3_step(const uint8_t * last_image, const uint8_t * next_image, const float * image_basis9, const float * kinv9,
                                const float * krlr9, int rows, int cols, void * workspace, float * out11, void * stream)
{
    SLAM_ARG_CHECK(last_image && next_image && image_basis9 && kinv9 && krlr9 && workspace && out11);
    SLAM_ARG_CHECK(rows > 0 && cols > 0);
    So3Args a;
    a.lastImage = last_image; a.nextImage = next_image;
    a.imageBasis = mat3_from(image_basis9);
    a.kinv = mat3_from(kinv9);
    a.krlr = mat3_from(krlr9);
    a.cols = cols; a.rows = rows;
    return launch_so3_step(a, workspace, out11, (cudaStream_t)stream);
}
