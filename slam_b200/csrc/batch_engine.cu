// Batched streaming engine: many independent sequences (or pose hypotheses) per GPU.
//
// The persistent kernel (gn_kernel.cu) is built for ONE sequence, where a frame is a chain of ~46
// latency-bound reductions.  With a batch of sequences the same chain is run in lock step for all of them
// and every link becomes a large streaming launch: grid = (pixel blocks, sequences), each block reduces its
// pixels of one sequence, the block that draws the last ticket of that sequence folds the partials
// (common.cuh: grid_finish), and a one-warp-per-sequence kernel does the fp64 bookkeeping between the map
// launches with the very routines the persistent kernel uses (gn_scalar.cuh).  Per Gauss-Newton iteration:
//     kb_phase_a  (ICP products + RGB association, 48 + ~14 B/px)      reduce.cu:257-416, 739-867
//     kb_phase_b  (RGB Jacobian products from the stored correspondences)  reduce.cu:494-624
//     kb_update   (combine, solve, pose, next parameters; 1 warp / sequence)  RGBDOdometryef.cpp:509-575
// and per SO3 iteration kb_so3_map + kb_so3_update.  All per-sequence state lives in device memory; the host
// enqueues a fixed launch sequence (sequences that finished a loop early skip their blocks) and reads back
// the GnResult array once.  Same per-pixel arithmetic (pixel_ops.cuh) => same masks as the single-sequence path.
#include "batch_engine.cuh"
#include "gn_scalar.cuh"

namespace slam {

constexpr int kBThreads = 256;

__device__ __forceinline__ const GnShared & bstate(const char * states, size_t stride, int seq)
{
    return *reinterpret_cast<const GnShared *>(states + (size_t)seq * stride);
}

// copy the persistent prefix of GnShared between global and shared memory (one warp)
__device__ __forceinline__ void state_load(GnShared & sh, const char * g)
{
    const int n = (int)(offsetof(GnShared, red) / 4);
    const unsigned * src = reinterpret_cast<const unsigned *>(g);
    unsigned * dst = reinterpret_cast<unsigned *>(&sh);
    for(int i = threadIdx.x; i < n; i += 32) dst[i] = src[i];
    __syncwarp();
}
__device__ __forceinline__ void state_store(const GnShared & sh, char * g)
{
    __syncwarp();
    const int n = (int)(offsetof(GnShared, red) / 4);
    unsigned * dst = reinterpret_cast<unsigned *>(g);
    const unsigned * src = reinterpret_cast<const unsigned *>(&sh);
    for(int i = threadIdx.x; i < n; i += 32) dst[i] = src[i];
}

// ------------------------------------------------------------------ one warp per sequence
__global__ void __launch_bounds__(32) kb_begin(const GnLaunch L, const GnSeqIn * seqs, char * states, size_t stride)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    if(threadIdx.x == 0)
    {
        seq_begin(sh, seqs[seq]);
        sh.stop_level = -1;
        sh.ntr = 0;
        sh.so3_done = L.so3 ? 0 : 1;
        for(int k = 0; k < 16; k++) sh.resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
        if(L.so3)
        {
            level_begin(sh, L.geom[2]);
            so3_prepare(sh);
        }
    }
    state_store(sh, states + (size_t)seq * stride);
}

__global__ void __launch_bounds__(32) kb_so3_update(const GnLaunch L, char * states, size_t stride, const float * sums, int it, slam_step_record * trace)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    state_load(sh, states + (size_t)seq * stride);
    if(sh.so3_done) return;
    if(threadIdx.x < 11) sh.total[threadIdx.x] = sums[seq * 64 + threadIdx.x];
    __syncwarp();
    if(threadIdx.x == 0)
    {
        slam_step_record * rec = (trace && sh.ntr < kGnMaxTrace) ? trace + (size_t)seq * kGnMaxTrace + sh.ntr : nullptr;
        if(rec) memset(rec, 0, sizeof(*rec));
        so3_update(sh, it, rec);
        if(rec) sh.ntr++;
        if(sh.stop || it == 9)
            sh.so3_done = 1;
        else
            so3_prepare(sh);
    }
    state_store(sh, states + (size_t)seq * stride);
}

__global__ void __launch_bounds__(32) kb_level_begin(const GnLaunch L, char * states, size_t stride, int lvl, int first)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    state_load(sh, states + (size_t)seq * stride);
    if(threadIdx.x == 0)
    {
        if(first && L.so3)
            for(int x = 0; x < 3; x++)
                for(int y = 0; y < 3; y++) sh.resultRt[x * 4 + y] = sh.resultR[x * 3 + y];
        sh.res.lastRGBError = FLT_MAX;
        sh.stop_level = -1;
        level_begin(sh, L.geom[lvl]);
    }
    __syncwarp();
    warp_prepare(sh, false);
    state_store(sh, states + (size_t)seq * stride);
}

__global__ void __launch_bounds__(32) kb_update(const GnLaunch L, char * states, size_t stride, const float * sums, int lvl, int j, slam_step_record * trace)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    state_load(sh, states + (size_t)seq * stride);
    if(sh.stop_level == lvl) return;
    for(int i = threadIdx.x; i < 64; i += 32) sh.total[i] = sums[seq * 64 + i];
    slam_step_record * rec = (trace && sh.ntr < kGnMaxTrace) ? trace + (size_t)seq * kGnMaxTrace + sh.ntr : nullptr;
    if(rec && threadIdx.x == 0)
    {
        memset(rec, 0, sizeof(*rec));
        rec->kind = 1;
        rec->level = lvl;
        rec->iteration = j;
        for(int k = 0; k < 9; k++)
        {
            rec->Rcurr_in[k] = sh.Rcurr[k];
            rec->krkinv_in[k] = sh.krk[k];
            rec->so3_in[k] = sh.Rprev_inv[k];
        }
        for(int k = 0; k < 3; k++)
        {
            rec->tcurr_in[k] = sh.tcurr[k];
            rec->kt_in[k] = sh.kt[k];
        }
    }
    __syncwarp();
    if(L.rgb)
    {
        if(threadIdx.x == 0)
        {
            gn_sigma(sh, L.rgb_only, rec);
            if(sh.stop) sh.stop_level = lvl;   // rgbOnly && rgbError > lastRGBError, RGBDOdometryef.cpp:460-463
        }
        __syncwarp();
    }
    if(sh.stop_level != lvl)
    {
        warp_update(sh, L.icp, L.rgb, L.icp_weight, threadIdx.x == 0 ? rec : nullptr, clock64());
        if(rec && threadIdx.x == 0) sh.ntr++;
        __syncwarp();
    }
    state_store(sh, states + (size_t)seq * stride);
}

__global__ void __launch_bounds__(32) kb_end(const GnLaunch L, char * states, size_t stride, GnResult * results, int * trace_count)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    state_load(sh, states + (size_t)seq * stride);
    if(threadIdx.x == 0)
    {
        seq_end(sh, L.rgb, L.rgb_only, &results[seq]);
        if(trace_count) trace_count[seq] = sh.ntr;
    }
}

// ------------------------------------------------------------------ streaming map-reduce launches
__global__ void __launch_bounds__(kBThreads) kb_so3_map(const GnLaunch L, const GnSeqIn * seqs, const char * states, size_t stride, char * ws, float * sums)
{
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.so3_done) return;
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[2];
    So3Args a;
    a.lastImage = in.lastNextImage[2];
    a.nextImage = in.nextImage[2];
    a.imageBasis = mat3_from(st.so3H);
    a.kinv = mat3_from(st.so3Kinv);
    a.krlr = mat3_from(st.so3KR);
    a.cols = g.cols;
    a.rows = g.rows;
    float acc[11];
#pragma unroll
    for(int k = 0; k < 11; k++) acc[k] = 0.f;
    const int N = g.rows * g.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int y = k / g.cols;
        const int x = k - y * g.cols;
        float row[4];
        const bool found = so3_pixel(a, x, y, row);
        accumulate_so3(acc, row, found);
    }
    grid_finish<float, 11>(acc, ws + (size_t)seq * kWorkspaceBytes, sums + seq * 64);
}

// Pose-independent half of the RGB association, once per level and frame (reduce.cu:780-807).
__global__ void __launch_bounds__(kBThreads) kb_candidates(const GnLaunch L, const GnSeqIn * seqs, unsigned char * cand0, size_t aux_stride, size_t cand_off, int lvl)
{
    const int seq = blockIdx.y;
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    ResidualArgs a;
    a.minScale = L.min_scale[lvl];
    a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
    a.nextDepth = in.nextDepth[lvl];
    a.nextImage = in.nextImage[lvl];
    a.cols = g.cols; a.rows = g.rows;
    unsigned char * cand = cand0 + (size_t)seq * aux_stride + cand_off;
    const int N = g.rows * g.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int i = k / g.cols;
        cand[k] = rgb_candidate(a, k - i * g.cols, i) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kBThreads) kb_phase_a(const GnLaunch L, const GnSeqIn * seqs, const char * states, size_t stride, char * ws_float, char * ws_int,
                                                        float * sums, unsigned char * cand0, size_t aux_stride, size_t cand_off, size_t vmask_off, int lvl)
{
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.stop_level == lvl) return;
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;
    const bool vec = (plane & 3) == 0;
    const int nitems = (plane + 3) >> 2;

    float acc[29];
#pragma unroll
    for(int k = 0; k < 29; k++) acc[k] = 0.f;
    int cnt[2] = {0, 0};

    IcpArgs ia;
    ia.Rcurr = mat3_from(st.Rcurr);
    ia.tcurr = make_float3(st.tcurr[0], st.tcurr[1], st.tcurr[2]);
    ia.Rprev_inv = mat3_from(st.Rprev_inv);
    ia.tprev = make_float3(st.tprev[0], st.tprev[1], st.tprev[2]);
    ia.fx = g.fx; ia.fy = g.fy; ia.cx = g.cx; ia.cy = g.cy;
    ia.distThres = L.dist_thresh;
    ia.angleThres = L.angle_thresh;
    ia.cols = g.cols; ia.rows = g.rows;
    ia.vcurr = in.vcurr[lvl]; ia.ncurr = in.ncurr[lvl]; ia.vprev = in.vprev[lvl]; ia.nprev = in.nprev[lvl];

    ResidualArgs ra;
    ra.minScale = L.min_scale[lvl];
    ra.dIdx = in.dIdx[lvl]; ra.dIdy = in.dIdy[lvl];
    ra.lastDepth = in.lastDepth[lvl]; ra.nextDepth = in.nextDepth[lvl];
    ra.lastImage = in.lastImage[lvl]; ra.nextImage = in.nextImage[lvl];
    ra.maxDepthDelta = L.max_depth_delta;
    ra.kt = make_float3(st.kt[0], st.kt[1], st.kt[2]);
    ra.krkinv = mat3_from(st.krk);
    ra.cols = g.cols; ra.rows = g.rows;
    const unsigned char * cand = cand0 + (size_t)seq * aux_stride + cand_off;
    unsigned char * vmask = cand0 + (size_t)seq * aux_stride + vmask_off;
    Corres * cimg = in.corres[lvl];

    for(int item = blockIdx.x * blockDim.x + threadIdx.x; item < nitems; item += gridDim.x * blockDim.x)
    {
        const int p = item << 2;
        if(L.icp)
        {
            float3 vg[4], nc[4], vp[4], np[4];
            int o[4];
            bool ok[4];
            if(vec)
            {
                const float4 a0 = __ldg(reinterpret_cast<const float4 *>(ia.vcurr + p));
                const float4 a1 = __ldg(reinterpret_cast<const float4 *>(ia.vcurr + plane + p));
                const float4 a2 = __ldg(reinterpret_cast<const float4 *>(ia.vcurr + 2 * plane + p));
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(ia.ncurr + p));
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(ia.ncurr + plane + p));
                const float4 b2 = __ldg(reinterpret_cast<const float4 *>(ia.ncurr + 2 * plane + p));
                vg[0] = make_float3(a0.x, a1.x, a2.x); vg[1] = make_float3(a0.y, a1.y, a2.y); vg[2] = make_float3(a0.z, a1.z, a2.z); vg[3] = make_float3(a0.w, a1.w, a2.w);
                nc[0] = make_float3(b0.x, b1.x, b2.x); nc[1] = make_float3(b0.y, b1.y, b2.y); nc[2] = make_float3(b0.z, b1.z, b2.z); nc[3] = make_float3(b0.w, b1.w, b2.w);
#pragma unroll
                for(int c = 0; c < 4; c++) ok[c] = true;
            }
            else
            {
#pragma unroll
                for(int c = 0; c < 4; c++)
                {
                    ok[c] = p + c < plane;
                    const int kk = ok[c] ? p + c : 0;
                    vg[c] = make_float3(__ldg(ia.vcurr + kk), __ldg(ia.vcurr + plane + kk), __ldg(ia.vcurr + 2 * plane + kk));
                    nc[c] = make_float3(__ldg(ia.ncurr + kk), __ldg(ia.ncurr + plane + kk), __ldg(ia.ncurr + 2 * plane + kk));
                }
            }
#pragma unroll
            for(int c = 0; c < 4; c++)
            {
                float3 g3;
                const bool inb = icp_project(ia, vg[c], g3, o[c]);
                vg[c] = g3;
                ok[c] = ok[c] && inb;
                if(!ok[c]) o[c] = 0;
            }
#pragma unroll
            for(int c = 0; c < 4; c++)
            {
                vp[c] = make_float3(__ldg(ia.vprev + o[c]), __ldg(ia.vprev + plane + o[c]), __ldg(ia.vprev + 2 * plane + o[c]));
                np[c] = make_float3(__ldg(ia.nprev + o[c]), __ldg(ia.nprev + plane + o[c]), __ldg(ia.nprev + 2 * plane + o[c]));
            }
#pragma unroll
            for(int c = 0; c < 4; c++)
            {
                float row[7];
                const bool found = icp_finish(ia, vg[c], nc[c], vp[c], np[c], row) && ok[c];
                if(found) accumulate_se3(acc, row, true);
            }
        }
        if(L.rgb)
        {
            // candidate flags of the four pixels, then the warp + gathers of the candidates
            unsigned cm = 0;
            if(vec)
                cm = __ldg(reinterpret_cast<const unsigned *>(cand + p));
            else
                for(int c = 0; c < 4; c++)
                    if(p + c < plane) cm |= (unsigned)cand[p + c] << (8 * c);
            unsigned vm = 0;
            if(cm)
            {
#pragma unroll
                for(int c = 0; c < 4; c++)
                    if((cm >> (8 * c)) & 0xff)
                    {
                        const int k = p + c;
                        const int i = k / g.cols;
                        Corres cc;
                        cc.zx = cc.zy = cc.ox = cc.oy = 0;
                        cc.diff = 0.f;
                        cc.valid = 0;
                        if(rgb_associate(ra, k - i * g.cols, i, cc))
                        {
                            cnt[0] += 1;
                            cnt[1] += (int)(cc.diff * cc.diff);
                            vm |= 1u << (8 * c);
                            reinterpret_cast<int4 *>(cimg)[k] = *reinterpret_cast<const int4 *>(&cc);
                        }
                    }
            }
            if(vec)
                *reinterpret_cast<unsigned *>(vmask + p) = vm;
            else
                for(int c = 0; c < 4; c++)
                    if(p + c < plane) vmask[p + c] = (vm >> (8 * c)) & 0xff;
        }
    }
    grid_finish<float, 29>(acc, ws_float + (size_t)seq * kWorkspaceBytes, sums + seq * 64);
    if(L.rgb) grid_finish<int, 2>(cnt, ws_int + (size_t)seq * kWorkspaceBytes, reinterpret_cast<int *>(sums + seq * 64 + 29));
}

__global__ void __launch_bounds__(kBThreads) kb_phase_b(const GnLaunch L, const GnSeqIn * seqs, const char * states, size_t stride, char * ws_float, float * sums,
                                                        const unsigned char * cand0, size_t aux_stride, size_t vmask_off, int lvl)
{
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.stop_level == lvl) return;
    // sigmaVal and the rgbOnly early exit are re-derived from the folded count / sigma by every block, identically
    // (RGBDOdometryef.cpp:457-471); kb_update does the bookkeeping once per sequence afterwards.
    const int rgbSize = __float_as_int(__ldcg(sums + seq * 64 + 29));
    const int sigma = __float_as_int(__ldcg(sums + seq * 64 + 30));
    const int sel = (rgbSize != 0 && sigma == 0) ? 1 : rgbSize;
    float sigmaVal = __fsqrt_rn((float)sel);
    if(L.rgb_only)
    {
        const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
        if(rgbError > st.res.lastRGBError) return;
        sigmaVal = -1;
    }
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;
    const bool vec = (plane & 3) == 0;
    const int nitems = (plane + 3) >> 2;
    RgbStepArgs a;
    a.sigma = sigmaVal;
    a.fx = g.fx; a.fy = g.fy;
    a.sobelScale = L.sobel_scale;
    a.cols = g.cols; a.rows = g.rows;
    a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
    a.lastDepth = in.lastDepth[lvl];
    a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
    a.cloud = nullptr;
    const unsigned char * vmask = cand0 + (size_t)seq * aux_stride + vmask_off;
    const Corres * cimg = in.corres[lvl];
    float acc[29];
#pragma unroll
    for(int k = 0; k < 29; k++) acc[k] = 0.f;
    for(int item = blockIdx.x * blockDim.x + threadIdx.x; item < nitems; item += gridDim.x * blockDim.x)
    {
        const int p = item << 2;
        unsigned vm = 0;
        if(vec)
            vm = __ldcg(reinterpret_cast<const unsigned *>(vmask + p));
        else
            for(int c = 0; c < 4; c++)
                if(p + c < plane) vm |= (unsigned)__ldcg(vmask + p + c) << (8 * c);
        if(!vm) continue;
#pragma unroll
        for(int c = 0; c < 4; c++)
            if((vm >> (8 * c)) & 0xff)
            {
                const int4 raw = __ldcg(reinterpret_cast<const int4 *>(cimg) + p + c);
                const Corres cc = *reinterpret_cast<const Corres *>(&raw);
                float row[7];
                rgb_row(a, cc, row);
                accumulate_se3(acc, row, true);
            }
    }
    grid_finish<float, 29>(acc, ws_float + (size_t)seq * kWorkspaceBytes, sums + seq * 64 + 32);
}

// ------------------------------------------------------------------ host side
static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

size_t batch_state_bytes(int batch, const LevelGeom * geom, int levels)
{
    size_t aux = 0;
    for(int l = 0; l < levels; l++) aux += 2 * up256((size_t)geom[l].rows * geom[l].cols);
    return up256(up256(offsetof(GnShared, red)) * batch) + up256((size_t)batch * 64 * 4) + aux * batch + 2 * kWorkspaceBytes * (size_t)batch + 4096;
}

void batch_bind_state(BatchDevice & d, char * base, int batch, const LevelGeom * geom, int levels, GnSeqIn * seq_in, GnResult * results)
{
    d.batch = batch;
    d.seq_in = seq_in;
    d.results = results;
    char * p = base;
    d.state_stride = up256(offsetof(GnShared, red));
    d.states = p;
    p += up256(d.state_stride * batch);
    d.sums = (float *)p;
    p += up256((size_t)batch * 64 * 4);
    size_t off = 0;
    for(int l = 0; l < levels; l++)
    {
        const size_t n = up256((size_t)geom[l].rows * geom[l].cols);
        d.cand_off[l] = off;
        off += n;
        d.vmask_off[l] = off;
        off += n;
    }
    d.aux_stride = off;
    d.cand0 = (unsigned char *)p;
    p += off * batch;
    d.ws_float = p;
    p += kWorkspaceBytes * (size_t)batch;
    d.ws_int = p;
}

static int blocks_for(int nitems)
{
    int g = (nitems + kBThreads - 1) / kBThreads;
    if(g < 1) g = 1;
    if(g > kMaxReduceBlocks) g = kMaxReduceBlocks;
    return g;
}

int batch_enqueue(BatchDevice & d, const GnLaunch & L, const GnSeqIn * h_seq_in_pinned, GnResult * h_results, slam_step_record * trace, int * trace_count,
                  cudaStream_t s)
{
    const int B = L.batch;
    if(!L.trace) trace = nullptr, trace_count = nullptr;
    SLAM_CUDA_TRY(cudaMemcpyAsync(d.seq_in, h_seq_in_pinned, sizeof(GnSeqIn) * B, cudaMemcpyHostToDevice, s));
    kb_begin<<<B, 32, 0, s>>>(L, d.seq_in, d.states, d.state_stride);
    d.launches++;
    if(L.so3)
    {
        const int nb = blocks_for(L.geom[2].rows * L.geom[2].cols);
        for(int it = 0; it < 10; it++)
        {
            kb_so3_map<<<dim3(nb, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws_float, d.sums);
            kb_so3_update<<<B, 32, 0, s>>>(L, d.states, d.state_stride, d.sums, it, trace);
            d.launches += 2;
        }
    }
    bool first = true;
    for(int lvl = L.levels - 1; lvl >= 0; lvl--)
    {
        const int plane = L.geom[lvl].rows * L.geom[lvl].cols;
        kb_level_begin<<<B, 32, 0, s>>>(L, d.states, d.state_stride, lvl, first ? 1 : 0);
        d.launches++;
        first = false;
        if(L.iterations[lvl] <= 0) continue;
        if(L.rgb)
        {
            kb_candidates<<<dim3(blocks_for(plane), B), kBThreads, 0, s>>>(L, d.seq_in, d.cand0, d.aux_stride, d.cand_off[lvl], lvl);
            d.launches++;
        }
        const int nb = blocks_for((plane + 3) / 4);
        for(int j = 0; j < L.iterations[lvl]; j++)
        {
            kb_phase_a<<<dim3(nb, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws_float, d.ws_int, d.sums, d.cand0, d.aux_stride, d.cand_off[lvl],
                                                         d.vmask_off[lvl], lvl);
            d.launches++;
            if(L.rgb)
            {
                kb_phase_b<<<dim3(nb, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws_float, d.sums, d.cand0, d.aux_stride, d.vmask_off[lvl], lvl);
                d.launches++;
            }
            kb_update<<<B, 32, 0, s>>>(L, d.states, d.state_stride, d.sums, lvl, j, trace);
            d.launches++;
        }
    }
    kb_end<<<B, 32, 0, s>>>(L, d.states, d.state_stride, d.results, trace_count);
    d.launches++;
    SLAM_CUDA_TRY(cudaGetLastError());
    SLAM_CUDA_TRY(cudaMemcpyAsync(h_results, d.results, sizeof(GnResult) * B, cudaMemcpyDeviceToHost, s));
    return SLAM_OK;
}

}   // namespace slam
