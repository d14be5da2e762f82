// Randomised-fern relocalisation / loop-closure front end: the reference's `Ferns`
// (src/lc/Ferns.cpp) around the tracker's ICP path, with the key-frame database resident in HBM.
//
//   k_fern_encode   gl/Resize.cpp:70-154 (NEAREST fetch at the centre of each 8x8 block) for the colour, vertex and
//                   normal images + the per-fern 4-bit codes of Ferns.cpp:108-131 / 197-219, one launch
//   k_fern_search   the co-occurrence count of Ferns.cpp:121-124 / 213-216 (the reference walks inverted lists
//                   conservatory[i].ids[code]; counting equal valid codes per key frame is the same number), the
//                   dissimilarity of :136-147 / 227-239, blockHDAware's operands (:374-389) and the arg-min, as ONE
//                   streaming launch: one warp per key frame, 512 B per key frame, packed (dissimilarity, id) atomicMin
//   k_fern_photo    photometricCheck, Ferns.cpp:309-357, and the sampled vertices of the constraints (:281-299)
// The ICP refinement (Ferns.cpp:253-268) is a slam_odom handle at 1/8 resolution on the same stream.
#include <algorithm>
#include <array>
#include <cfloat>
#include <cstring>
#include <random>
#include <vector>
#include "../../include/slam_ferns.h"
#include "common.cuh"
#include "small_math.hpp"

namespace slam {

constexpr int kFernFactor = 8;       // Ferns::factor, Ferns.cpp:23
constexpr unsigned char kBadCode = 255;

struct FernDev
{
    short x, y;
    unsigned char r, g, b, pad;
    int d;
};

// ------------------------------------------------------------------ kernels
// blocks [0, nb_px): resize; block nb_px: codes
__global__ void __launch_bounds__(256) k_fern_encode(const uchar4 * __restrict__ rgba, const float4 * __restrict__ vert, const float4 * __restrict__ norm, int W,
                                                     int H, int sw, int sh, const FernDev * __restrict__ table, int num, int code_stride,
                                                     unsigned char * __restrict__ q_rgb, float4 * __restrict__ q_vert, float4 * __restrict__ q_norm,
                                                     unsigned char * __restrict__ q_codes, int * __restrict__ q_good, int nb_px)
{
    if((int)blockIdx.x < nb_px)
    {
        const int k = blockIdx.x * blockDim.x + threadIdx.x;
        if(k >= sw * sh) return;
        const int y = k / sw, x = k - y * sw;
        const int src = (kFernFactor * y + kFernFactor / 2) * W + (kFernFactor * x + kFernFactor / 2);
        const uchar4 c = rgba[src];
        q_rgb[3 * k + 0] = c.x;
        q_rgb[3 * k + 1] = c.y;
        q_rgb[3 * k + 2] = c.z;
        q_vert[k] = vert[src];
        q_norm[k] = norm[src];
        return;
    }
    __shared__ int warp_good[8];
    int good = 0;
    for(int i = threadIdx.x; i < code_stride; i += blockDim.x)
    {
        unsigned char code = kBadCode;
        if(i < num)
        {
            const FernDev f = table[i];
            const int src = (kFernFactor * f.y + kFernFactor / 2) * W + (kFernFactor * f.x + kFernFactor / 2);
            const float z = vert[src].z;
            if(z > 0)
            {
                const uchar4 pix = rgba[src];
                code = (unsigned char)(((pix.x > f.r) << 3) | ((pix.y > f.g) << 2) | ((pix.z > f.b) << 1) | ((int)__fmul_rn(z, 1000.0f) > f.d));
                good++;
            }
        }
        q_codes[i] = code;   // the padding bytes [num, code_stride) are bad codes: they never vote
    }
    for(int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
    if((threadIdx.x & 31) == 0) warp_good[threadIdx.x >> 5] = good;
    __syncthreads();
    if(threadIdx.x == 0)
    {
        int t = 0;
        for(int w = 0; w < (int)(blockDim.x >> 5); w++) t += warp_good[w];
        *q_good = t;
    }
}

// One warp per key frame.  best: packed (float bits of the dissimilarity << 32 | id), initialised to all ones.
__global__ void __launch_bounds__(256) k_fern_search(const unsigned char * __restrict__ q_codes, const int * __restrict__ q_good,
                                                     const unsigned char * __restrict__ db_codes, const int * __restrict__ db_good,
                                                     const int * __restrict__ db_time, int nframes, int code_stride, int time, int use_time,
                                                     float * __restrict__ dissim, int2 * __restrict__ co_both, unsigned long long * __restrict__ best)
{
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(f >= nframes) return;
    const uint4 * q = reinterpret_cast<const uint4 *>(q_codes);
    const uint4 * c = reinterpret_cast<const uint4 *>(db_codes + (size_t)f * code_stride);
    int co = 0, both = 0;
    for(int v = lane; v < code_stride / 16; v += 32)
    {
        const uint4 a = q[v];
        const uint4 b = __ldcs(c + v);   // streamed once per query
        const unsigned aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
            const unsigned qvalid = ~__vcmpeq4(aw[k], 0xffffffffu);
            const unsigned fvalid = ~__vcmpeq4(bw[k], 0xffffffffu);
            co += __popc(__vcmpeq4(aw[k], bw[k]) & qvalid) >> 3;
            both += __popc(qvalid & fvalid) >> 3;
        }
    }
    for(int o = 16; o > 0; o >>= 1)
    {
        co += __shfl_xor_sync(0xffffffffu, co, o);
        both += __shfl_xor_sync(0xffffffffu, both, o);
    }
    if(lane == 0)
    {
        const float maxCo = (float)min(*q_good, db_good[f]);                       // float maxCo = std::min(goodCodes, ...)
        const float d = __fdiv_rn(__fsub_rn(maxCo, (float)co), maxCo);           // (float)(maxCo - co) / (float)maxCo
        dissim[f] = d;
        co_both[f] = make_int2(co, both);
        const bool eligible = !use_time || (time - db_time[f] > 300);
        if(eligible && !isnan(d)) atomicMin(best, ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)f);
    }
}

struct PhotoArgs
{
    float diff[16];        // fernPose^-1 * estPose, row-major
    float cx, cy, fxr, fyr;   // cx, cy, 1/invfx, 1/invfy of the 1/8 camera
    int sw, sh, num, max_depth, cons_step;
};

// One block.  out[0] = photoSum (integer-valued), out[1] = photoCount; cons[j] = query vertex at fern j*cons_step, w = validity.
__global__ void __launch_bounds__(512) k_fern_photo(const PhotoArgs a, const FernDev * __restrict__ table, const float4 * __restrict__ q_vert,
                                                    const unsigned char * __restrict__ q_rgb, const unsigned char * __restrict__ f_rgb,
                                                    int * __restrict__ out, float4 * __restrict__ cons)
{
    __shared__ int s_sum[16], s_cnt[16];
    int sum = 0, cnt = 0;
    for(int i = threadIdx.x; i < a.num; i += blockDim.x)
    {
        const FernDev f = table[i];
        const int k = f.y * a.sw + f.x;
        const float4 v = q_vert[k];
        const bool usable = v.z > 0 && (int)__fmul_rn(v.z, 1000.0f) < a.max_depth;
        if(a.cons_step > 0 && i % a.cons_step == 0) cons[i / a.cons_step] = make_float4(v.x, v.y, v.z, usable ? 1.f : 0.f);
        if(!usable) continue;
        // diff * (x, y, z, 1): a linear combination of the columns, accumulated left to right
        float p[3];
#pragma unroll
        for(int r = 0; r < 3; r++)
            p[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.diff[r * 4 + 0], v.x), __fmul_rn(a.diff[r * 4 + 1], v.y)), __fmul_rn(a.diff[r * 4 + 2], v.z)),
                             a.diff[r * 4 + 3]);
        const int u = (int)__fadd_rn(__fdiv_rn(__fmul_rn(p[0], a.fxr), p[2]), a.cx);
        const int w = (int)__fadd_rn(__fdiv_rn(__fmul_rn(p[1], a.fyr), p[2]), a.cy);
        if(u >= 0 && w >= 0 && u < a.sw && w < a.sh)
        {
            const unsigned char * fp = f_rgb + 3 * (w * a.sw + u);
            if(fp[0] > 0 || fp[1] > 0 || fp[2] > 0)
            {
                const unsigned char * qp = q_rgb + 3 * k;
                sum += abs((int)fp[0] - (int)qp[0]) + abs((int)fp[1] - (int)qp[1]) + abs((int)fp[2] - (int)qp[2]);
                cnt++;
            }
        }
    }
    for(int o = 16; o > 0; o >>= 1)
    {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if((threadIdx.x & 31) == 0)
    {
        s_sum[threadIdx.x >> 5] = sum;
        s_cnt[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        int ts = 0, tc = 0;
        for(int w = 0; w < (int)(blockDim.x >> 5); w++)
        {
            ts += s_sum[w];
            tc += s_cnt[w];
        }
        out[0] = ts;
        out[1] = tc;
    }
}

}   // namespace slam

using namespace slam;

// ------------------------------------------------------------------ handle
struct slam_ferns
{
    slam_ferns_params p{};
    int num = 0, sw = 0, sh = 0, code_stride = 0, capacity = 0, nframes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    slam_odom_t rgbd = nullptr;          // RGBDOdometryef rgbd(w/8, h/8, cx/8, cy/8, fx/8, fy/8), Ferns.cpp:34-39
    std::vector<slam_fern> table;
    FernDev * d_table = nullptr;
    // query slot
    unsigned char * q_codes = nullptr, * q_rgb = nullptr;
    float4 * q_vert = nullptr, * q_norm = nullptr;
    int * q_good = nullptr;
    bool have_query = false;
    // database
    unsigned char * db_codes = nullptr, * db_rgb = nullptr;
    float4 * db_vert = nullptr, * db_norm = nullptr;
    int * db_good = nullptr, * db_time = nullptr;
    std::vector<std::array<float, 16>> poses;
    std::vector<int> times, goods;
    // results
    float * d_dissim = nullptr;
    int2 * d_co = nullptr;
    unsigned long long * d_best = nullptr;
    int * d_photo = nullptr;
    float4 * d_cons = nullptr;
    // pinned host mirror: [0..1] best key, then good, photo[2]
    unsigned long long * h_best = nullptr;
    int * h_ints = nullptr;
    float last_search_ms = 0.f;
};

static size_t px(const slam_ferns * h) { return (size_t)h->sw * h->sh; }

static int ferns_set_device(slam_ferns_t h)
{
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    return SLAM_OK;
}

extern "C" int slam_ferns_destroy(slam_ferns_t h)
{
    if(!h) return SLAM_ERR_ARG;
    cudaSetDevice(h->p.device);
    if(h->rgbd) slam_odom_destroy(h->rgbd);
    void * bufs[] = {h->d_table, h->q_codes, h->q_rgb, h->q_vert, h->q_norm, h->q_good, h->db_codes, h->db_rgb, h->db_vert, h->db_norm, h->db_good,
                     h->db_time, h->d_dissim, h->d_co, h->d_best, h->d_photo, h->d_cons};
    for(void * b : bufs)
        if(b) cudaFree(b);
    if(h->h_best) cudaFreeHost(h->h_best);
    if(h->h_ints) cudaFreeHost(h->h_ints);
    if(h->ev0) cudaEventDestroy(h->ev0);
    if(h->ev1) cudaEventDestroy(h->ev1);
    if(h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SLAM_OK;
}

static int ferns_alloc(slam_ferns_t h)
{
    const size_t n = px(h), cap = (size_t)h->capacity;
    SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    SLAM_CUDA_TRY(cudaEventCreate(&h->ev0));
    SLAM_CUDA_TRY(cudaEventCreate(&h->ev1));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_table, sizeof(FernDev) * h->num));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->q_codes, h->code_stride));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->q_rgb, n * 3));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->q_vert, n * 16));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->q_norm, n * 16));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->q_good, 4));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_codes, cap * h->code_stride));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_rgb, cap * n * 3));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_vert, cap * n * 16));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_norm, cap * n * 16));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_good, cap * 4));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->db_time, cap * 4));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_dissim, cap * 4));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_co, cap * 8));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_best, 8));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_photo, 8));
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_cons, sizeof(float4) * h->num));
    SLAM_CUDA_TRY(cudaMallocHost((void **)&h->h_best, 16));
    SLAM_CUDA_TRY(cudaMallocHost((void **)&h->h_ints, 64));
    std::vector<FernDev> t(h->num);
    for(int i = 0; i < h->num; i++)
    {
        const slam_fern & f = h->table[i];
        t[i].x = (short)f.x;
        t[i].y = (short)f.y;
        t[i].r = (unsigned char)f.r;
        t[i].g = (unsigned char)f.g;
        t[i].b = (unsigned char)f.b;
        t[i].pad = 0;
        t[i].d = f.d;
    }
    SLAM_CUDA_TRY(cudaMemcpy(h->d_table, t.data(), sizeof(FernDev) * h->num, cudaMemcpyHostToDevice));
    return SLAM_OK;
}

extern "C" int slam_ferns_create(const slam_ferns_params * params, const slam_fern * table, slam_ferns_t * out)
{
    SLAM_ARG_CHECK(params && out);
    SLAM_ARG_CHECK(params->width >= 8 * kFernFactor && params->height >= 8 * kFernFactor && params->max_depth_mm >= 400);
    int ndev = 0;
    SLAM_CUDA_TRY(cudaGetDeviceCount(&ndev));   // no CPU fallback: fails loudly without a CUDA device
    SLAM_ARG_CHECK(params->device >= 0 && params->device < ndev);
    slam_ferns * h = new slam_ferns();
    h->p = *params;
    h->num = params->num_ferns > 0 ? params->num_ferns : 500;
    h->sw = params->width / kFernFactor;
    h->sh = params->height / kFernFactor;
    h->code_stride = (h->num + 511) / 512 * 512;   // whole uint4 trips of a warp
    h->capacity = params->capacity > 0 ? params->capacity : 1024;
    h->table.resize(h->num);
    if(table)
    {
        for(int i = 0; i < h->num; i++)
        {
            const slam_fern & f = table[i];
            if(f.x < 0 || f.x >= h->sw || f.y < 0 || f.y >= h->sh || f.r < 0 || f.r > 255 || f.g < 0 || f.g > 255 || f.b < 0 || f.b > 255)
            {
                delete h;
                set_last_error("fern table entry out of range");
                return SLAM_ERR_ARG;
            }
            h->table[i] = f;
        }
    }
    else
    {
        // Ferns::generateFerns, Ferns.cpp:67-81: the same engine, distributions and draw order
        std::mt19937 random(params->seed);
        std::uniform_int_distribution<int32_t> widthDist(0, h->sw - 1), heightDist(0, h->sh - 1), rgbDist(0, 255), dDist(400, params->max_depth_mm);
        for(int i = 0; i < h->num; i++)
        {
            slam_fern f;
            f.x = widthDist(random);
            f.y = heightDist(random);
            f.r = rgbDist(random);
            f.g = rgbDist(random);
            f.b = rgbDist(random);
            f.d = dDist(random);
            h->table[i] = f;
        }
    }
    int rc = ferns_set_device(h);
    if(rc == SLAM_OK) rc = ferns_alloc(h);
    if(rc == SLAM_OK)
    {
        slam_odom_params op;
        memset(&op, 0, sizeof(op));
        op.width = h->sw;
        op.height = h->sh;
        op.cx = params->cx / kFernFactor;
        op.cy = params->cy / kFernFactor;
        op.fx = params->fx / kFernFactor;
        op.fy = params->fy / kFernFactor;
        op.device = params->device;
        op.stream = h->stream;
        rc = slam_odom_create(&op, &h->rgbd);
    }
    if(rc != SLAM_OK)
    {
        slam_ferns_destroy(h);
        return rc;
    }
    *out = h;
    return SLAM_OK;
}

extern "C" int slam_ferns_get_table(slam_ferns_t h, slam_fern * out)
{
    SLAM_ARG_CHECK(h && out);
    memcpy(out, h->table.data(), sizeof(slam_fern) * h->num);
    return SLAM_OK;
}

extern "C" int slam_ferns_num_frames(slam_ferns_t h) { return h ? h->nframes : SLAM_ERR_ARG; }

// resize + encode into the query slot (asynchronous on the handle's stream)
static int encode_query(slam_ferns_t h, const uint8_t * rgba, const float * vert, const float * norm)
{
    const int nb_px = div_up(h->sw * h->sh, 256);
    k_fern_encode<<<nb_px + 1, 256, 0, h->stream>>>((const uchar4 *)rgba, (const float4 *)vert, (const float4 *)norm, h->p.width, h->p.height, h->sw, h->sh,
                                                     h->d_table, h->num, h->code_stride, h->q_rgb, h->q_vert, h->q_norm, h->q_codes, h->q_good, nb_px);
    SLAM_CUDA_TRY(cudaGetLastError());
    h->have_query = true;
    return SLAM_OK;
}

// search of the query against the database (asynchronous); the best key and the query's good count land in pinned memory
static int search_query(slam_ferns_t h, int time, int use_time)
{
    SLAM_CUDA_TRY(cudaMemsetAsync(h->d_best, 0xff, 8, h->stream));
    if(h->nframes > 0)
    {
        SLAM_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
        k_fern_search<<<div_up(h->nframes, 8), 256, 0, h->stream>>>(h->q_codes, h->q_good, h->db_codes, h->db_good, h->db_time, h->nframes, h->code_stride,
                                                                      time, use_time, h->d_dissim, h->d_co, h->d_best);
        SLAM_CUDA_TRY(cudaGetLastError());
        SLAM_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    }
    SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_best, h->d_best, 8, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_ints, h->q_good, 4, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if(h->nframes > 0) SLAM_CUDA_TRY(cudaEventElapsedTime(&h->last_search_ms, h->ev0, h->ev1));
    return SLAM_OK;
}

static void unpack_best(const slam_ferns * h, int * min_id, float * minimum)
{
    const unsigned long long key = *h->h_best;
    if(key == ~0ull)
    {
        *min_id = -1;
        *minimum = FLT_MAX;   // std::numeric_limits<float>::max(), Ferns.cpp:134,225
        return;
    }
    const unsigned bits = (unsigned)(key >> 32);
    memcpy(minimum, &bits, 4);
    *min_id = (int)(key & 0xffffffffu);
}

extern "C" int slam_ferns_add_frame(slam_ferns_t h, const uint8_t * d_rgba, const float * d_vertices4, const float * d_normals4, const float * pose16,
                                    int src_time, float threshold, int * added)
{
    SLAM_ARG_CHECK(h && d_rgba && d_vertices4 && d_normals4 && pose16 && added);
    if(int rc = ferns_set_device(h)) return rc;
    if(int rc = encode_query(h, d_rgba, d_vertices4, d_normals4)) return rc;
    if(int rc = search_query(h, 0, 0)) return rc;
    int min_id;
    float minimum;
    unpack_best(h, &min_id, &minimum);
    const int good = h->h_ints[0];
    // Ferns.cpp:134-147: the minimum is only formed when goodCodes > 0
    *added = 0;
    if((minimum > threshold || h->nframes == 0) && good > 0)
    {
        if(h->nframes >= h->capacity)
        {
            set_last_error("fern database full (raise slam_ferns_params.capacity)");
            return SLAM_ERR_UNSUPPORTED;
        }
        const size_t n = px(h), f = (size_t)h->nframes;
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_codes + f * h->code_stride, h->q_codes, h->code_stride, cudaMemcpyDeviceToDevice, h->stream));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_rgb + f * n * 3, h->q_rgb, n * 3, cudaMemcpyDeviceToDevice, h->stream));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_vert + f * n, h->q_vert, n * 16, cudaMemcpyDeviceToDevice, h->stream));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_norm + f * n, h->q_norm, n * 16, cudaMemcpyDeviceToDevice, h->stream));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_good + f, h->q_good, 4, cudaMemcpyDeviceToDevice, h->stream));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->db_time + f, &src_time, 4, cudaMemcpyHostToDevice, h->stream));
        SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));   // src_time is a stack variable
        std::array<float, 16> P;
        memcpy(P.data(), pose16, 64);
        h->poses.push_back(P);
        h->times.push_back(src_time);
        h->goods.push_back(good);
        h->nframes++;
        *added = 1;
    }
    return SLAM_OK;
}

static int photometric(slam_ferns_t h, int id, const float * est16, const float * fern16, bool with_constraints, float * photo_error, int * photo_count)
{
    // Eigen::Matrix4f diff = fernPose.inverse() * estPose  (Ferns.cpp:332), fp32
    float inv[16], diff[16];
    smath::mat4_inverse(fern16, inv);
    smath::mat4_mul(inv, est16, diff);
    PhotoArgs a;
    memcpy(a.diff, diff, 64);
    a.cx = h->p.cx / kFernFactor;
    a.cy = h->p.cy / kFernFactor;
    const float invfx = 1.0f / float(h->p.fx / kFernFactor), invfy = 1.0f / float(h->p.fy / kFernFactor);
    a.fxr = 1 / invfx;
    a.fyr = 1 / invfy;
    a.sw = h->sw;
    a.sh = h->sh;
    a.num = h->num;
    a.max_depth = h->p.max_depth_mm;
    a.cons_step = with_constraints ? std::max(1, h->num / 50) : 0;
    k_fern_photo<<<1, 512, 0, h->stream>>>(a, h->d_table, h->q_vert, h->q_rgb, h->db_rgb + (size_t)id * px(h) * 3, h->d_photo, h->d_cons);
    SLAM_CUDA_TRY(cudaGetLastError());
    SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_ints + 2, h->d_photo, 8, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    const float photoSum = (float)h->h_ints[2];   // a sum of integers below 2^24: exact in the reference's float accumulator too
    *photo_count = h->h_ints[3];
    *photo_error = photoSum / float(h->h_ints[3]);
    return SLAM_OK;
}

extern "C" int slam_ferns_find_frame(slam_ferns_t h, const float * curr_pose16, const float * d_vertices4, const float * d_normals4, const uint8_t * d_rgba,
                                     int time, int lost, float * est_pose16, slam_ferns_match * match, slam_surface_constraint * constraints,
                                     int max_constraints, int * n_constraints)
{
    SLAM_ARG_CHECK(h && curr_pose16 && d_vertices4 && d_normals4 && d_rgba && est_pose16);
    if(int rc = ferns_set_device(h)) return rc;
    slam_ferns_match m;
    memset(&m, 0, sizeof(m));
    m.min_id = m.last_closest = -1;
    if(n_constraints) *n_constraints = 0;
    for(int k = 0; k < 16; k++) est_pose16[k] = (k % 5 == 0) ? 1.f : 0.f;

    if(int rc = encode_query(h, d_rgba, d_vertices4, d_normals4)) return rc;
    if(int rc = search_query(h, time, 1)) return rc;
    float minimum;
    unpack_best(h, &m.min_id, &minimum);
    m.dissimilarity = minimum;
    if(m.min_id >= 0)
    {
        int2 cb;
        SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_ints + 4, h->d_co + m.min_id, 8, cudaMemcpyDeviceToHost, h->stream));
        SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
        cb.x = h->h_ints[4];
        cb.y = h->h_ints[5];
        m.block_hd_aware = (float)cb.x / (float)cb.y;   // val / (float)count, Ferns.cpp:388
    }
    if(m.min_id != -1 && m.block_hd_aware > 0.3)
    {
        const size_t n = px(h);
        const float * fernPose = h->poses[m.min_id].data();
        const float cutoff = (float)h->p.max_depth_mm / 1000.0f;
        // WARNING initICP* must be called before initRGB*  (Ferns.cpp:253-268): model = the key frame, current = the query
        if(int rc = slam_odom_init_icp_model(h->rgbd, (const float *)(h->db_vert + (size_t)m.min_id * n), (const float *)(h->db_norm + (size_t)m.min_id * n),
                                             cutoff, fernPose))
            return rc;
        if(int rc = slam_odom_init_icp_maps(h->rgbd, (const float *)h->q_vert, (const float *)h->q_norm, cutoff)) return rc;
        float trans[3] = {fernPose[3], fernPose[7], fernPose[11]};
        float rot[9] = {fernPose[0], fernPose[1], fernPose[2], fernPose[4], fernPose[5], fernPose[6], fernPose[8], fernPose[9], fernPose[10]};
        if(int rc = slam_odom_get_incremental_transformation(h->rgbd, trans, rot, 0, 100.f, 0, 0, 0)) return rc;
        for(int r = 0; r < 3; r++)
        {
            for(int c = 0; c < 3; c++) est_pose16[4 * r + c] = rot[3 * r + c];
            est_pose16[4 * r + 3] = trans[r];
        }
        slam_odom_stats st;
        if(int rc = slam_odom_get_stats(h->rgbd, &st)) return rc;
        m.icp_ran = 1;
        m.icp_error = st.lastICPError;
        m.icp_count = st.lastICPCount;
        int pc = 0;
        if(int rc = photometric(h, m.min_id, est_pose16, fernPose, true, &m.photo_error, &pc)) return rc;
        const int icpCountThresh = lost ? 1400 : 2400;
        if(m.icp_error < 0.0003 && m.icp_count > icpCountThresh && m.photo_error < h->p.photo_thresh)
        {
            m.last_closest = m.min_id;
            const int step = std::max(1, h->num / 50);
            const int nc = (h->num + step - 1) / step;
            std::vector<float> cons(4 * (size_t)nc);
            SLAM_CUDA_TRY(cudaMemcpyAsync(cons.data(), h->d_cons, sizeof(float) * 4 * nc, cudaMemcpyDeviceToHost, h->stream));
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
            int out = 0;
            for(int j = 0; j < nc; j++)
            {
                const float * v = &cons[4 * j];
                if(v[3] == 0.f) continue;
                if(constraints && out < max_constraints)
                {
                    // currPose * (x, y, z, 1) and estPose * (x, y, z, 1), Ferns.cpp:284-296
                    for(int r = 0; r < 4; r++)
                    {
                        constraints[out].source[r] = ((curr_pose16[4 * r] * v[0] + curr_pose16[4 * r + 1] * v[1]) + curr_pose16[4 * r + 2] * v[2]) + curr_pose16[4 * r + 3];
                        constraints[out].target[r] = ((est_pose16[4 * r] * v[0] + est_pose16[4 * r + 1] * v[1]) + est_pose16[4 * r + 2] * v[2]) + est_pose16[4 * r + 3];
                    }
                }
                out++;
            }
            if(n_constraints) *n_constraints = out;
        }
    }
    if(match) *match = m;
    return SLAM_OK;
}

// ------------------------------------------------------------------ operator-level entry points
extern "C" int slam_ferns_encode(slam_ferns_t h, const uint8_t * d_rgba, const float * d_vertices4, const float * d_normals4, uint8_t * codes, int * good_codes,
                                 uint8_t * rgb_small, float * vert_small, float * norm_small)
{
    SLAM_ARG_CHECK(h && d_rgba && d_vertices4 && d_normals4);
    if(int rc = ferns_set_device(h)) return rc;
    if(int rc = encode_query(h, d_rgba, d_vertices4, d_normals4)) return rc;
    const size_t n = px(h);
    if(codes) SLAM_CUDA_TRY(cudaMemcpyAsync(codes, h->q_codes, h->num, cudaMemcpyDeviceToHost, h->stream));
    if(good_codes) SLAM_CUDA_TRY(cudaMemcpyAsync(good_codes, h->q_good, 4, cudaMemcpyDeviceToHost, h->stream));
    if(rgb_small) SLAM_CUDA_TRY(cudaMemcpyAsync(rgb_small, h->q_rgb, n * 3, cudaMemcpyDeviceToHost, h->stream));
    if(vert_small) SLAM_CUDA_TRY(cudaMemcpyAsync(vert_small, h->q_vert, n * 16, cudaMemcpyDeviceToHost, h->stream));
    if(norm_small) SLAM_CUDA_TRY(cudaMemcpyAsync(norm_small, h->q_norm, n * 16, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SLAM_OK;
}

extern "C" int slam_ferns_search(slam_ferns_t h, int time, int use_time, float * dissim, int * min_id, float * minimum, float * block_hd_aware)
{
    SLAM_ARG_CHECK(h);
    if(!h->have_query)
    {
        set_last_error("slam_ferns_search: no encoded query (call slam_ferns_encode / add_frame / find_frame first)");
        return SLAM_ERR_ORDER;
    }
    if(int rc = ferns_set_device(h)) return rc;
    if(int rc = search_query(h, time, use_time)) return rc;
    int id;
    float mn;
    unpack_best(h, &id, &mn);
    if(min_id) *min_id = id;
    if(minimum) *minimum = mn;
    if(dissim && h->nframes > 0) SLAM_CUDA_TRY(cudaMemcpy(dissim, h->d_dissim, sizeof(float) * h->nframes, cudaMemcpyDeviceToHost));
    if(block_hd_aware)
    {
        *block_hd_aware = 0.f;
        if(id >= 0)
        {
            int2 cb;
            SLAM_CUDA_TRY(cudaMemcpy(&cb, h->d_co + id, 8, cudaMemcpyDeviceToHost));
            *block_hd_aware = (float)cb.x / (float)cb.y;
        }
    }
    return SLAM_OK;
}

extern "C" int slam_ferns_photometric_check(slam_ferns_t h, int id, const float * est_pose16, const float * fern_pose16, float * photo_error, int * photo_count)
{
    SLAM_ARG_CHECK(h && est_pose16 && fern_pose16 && photo_error && id >= 0 && id < h->nframes);
    if(!h->have_query)
    {
        set_last_error("slam_ferns_photometric_check: no encoded query");
        return SLAM_ERR_ORDER;
    }
    if(int rc = ferns_set_device(h)) return rc;
    int pc = 0;
    if(int rc = photometric(h, id, est_pose16, fern_pose16, false, photo_error, &pc)) return rc;
    if(photo_count) *photo_count = pc;
    return SLAM_OK;
}

extern "C" int slam_ferns_get_frame(slam_ferns_t h, int id, uint8_t * codes, float * pose16, int * src_time, int * good_codes)
{
    SLAM_ARG_CHECK(h && id >= 0 && id < h->nframes);
    if(int rc = ferns_set_device(h)) return rc;
    if(codes) SLAM_CUDA_TRY(cudaMemcpy(codes, h->db_codes + (size_t)id * h->code_stride, h->num, cudaMemcpyDeviceToHost));
    if(pose16) memcpy(pose16, h->poses[id].data(), 64);
    if(src_time) *src_time = h->times[id];
    if(good_codes) *good_codes = h->goods[id];
    return SLAM_OK;
}

extern "C" int slam_ferns_last_search_ms(slam_ferns_t h, float * ms)
{
    SLAM_ARG_CHECK(h && ms);
    *ms = h->last_search_ms;
    return SLAM_OK;
}
