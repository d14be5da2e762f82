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
#include <cstdio>
#include <cstdlib>
#include "batch_engine.cuh"
#include "gn_scalar.cuh"

namespace slam {

constexpr int kBThreads = 256;

__device__ __forceinline__ const GnShared & bstate(const char * states, size_t stride, int seq)
{
    return *reinterpret_cast<const GnShared *>(states + (size_t)seq * stride);
}

template <int PX>
__device__ __forceinline__ void vec_load(const float * p, float (&v)[PX])
{
    if(PX == 4)
    {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2 % PX] = q.z; v[3 % PX] = q.w;
    }
    else if(PX == 2)
    {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = q.x; v[1 % PX] = q.y;
    }
    else
        v[0] = __ldg(p);
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

// ------------------------------------------------------------------ per-sequence reduction + bookkeeping tails
constexpr int kRowWords = 40;   // one partial row per block: 32 floats + 2 ints (+ pad)

// Block sum of 32 floats (+ optionally 2 ints) per thread, published as this block's partial row of its sequence;
// the block that draws the sequence's last ticket folds all rows in block order into sh.total[dst..dst+31] (ints as
// bit patterns in columns 29, 30 when with_ints) and returns true (block-uniform).  Deterministic for a given grid.
__device__ __forceinline__ bool seq_reduce(float (&acc)[32], int c0, int c1, const bool with_ints, GnShared & sh, char * ws_seq, const int dst)
{
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float s = warp_reduce_scatter32(acc);
    sh.red[wid * 32 + lane] = s;
    int * redi = reinterpret_cast<int *>(sh.red + 32 * 32);
    if(with_ints)
    {
        c0 = warp_sum(c0);
        c1 = warp_sum(c1);
        if(lane == 0)
        {
            redi[wid * 2] = c0;
            redi[wid * 2 + 1] = c1;
        }
    }
    __syncthreads();
    float * rows = reinterpret_cast<float *>(ws_seq);
    if(threadIdx.x < 32)
    {
        float total = 0.f;
        for(int w = 0; w < nw; w++) total += sh.red[w * 32 + threadIdx.x];
        rows[blockIdx.x * kRowWords + threadIdx.x] = total;
    }
    else if(threadIdx.x < 34 && with_ints)
    {
        int t = 0;
        for(int w = 0; w < nw; w++) t += redi[w * 2 + (threadIdx.x - 32)];
        reinterpret_cast<int *>(rows)[blockIdx.x * kRowWords + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    unsigned * ticket = workspace_ticket(ws_seq);
    if(threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if(!s_last) return false;
    __threadfence();
    const int nb = gridDim.x;
    if(threadIdx.x < 32)
    {
        float total = 0.f;
        int r = 0;
        for(; r + 8 <= nb; r += 8)
        {
            float v[8];
#pragma unroll
            for(int q = 0; q < 8; q++) v[q] = __ldcg(rows + (r + q) * kRowWords + threadIdx.x);
#pragma unroll
            for(int q = 0; q < 8; q++) total += v[q];
        }
        for(; r < nb; r++) total += __ldcg(rows + r * kRowWords + threadIdx.x);
        if(!(with_ints && threadIdx.x >= 29 && threadIdx.x <= 30)) sh.total[dst + threadIdx.x] = total;
    }
    else if(threadIdx.x < 34 && with_ints)
    {
        int t = 0;
        for(int r = 0; r < nb; r++) t += __ldcg(reinterpret_cast<const int *>(rows) + r * kRowWords + threadIdx.x);
        sh.total[dst + 29 + (threadIdx.x - 32)] = __int_as_float(t);
    }
    if(threadIdx.x == 0) *ticket = 0u;   // ready for the next launch
    __syncthreads();
    return true;
}

// One warp (warp 0 of the block that finished a sequence's reduction, sums in sh.total): RGBDOdometryef.cpp:457-575.
__device__ __forceinline__ void seq_update(GnShared & sh, const GnLaunch & L, char * state, int seq, int lvl, int j, slam_step_record * trace)
{
    state_load(sh, state);
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
        if(threadIdx.x == 0) gn_sigma(sh, L.rgb_only, rec);   // never a stop here: phase B returned early in that case
        __syncwarp();
    }
    warp_stats(sh, L.icp, L.rgb, L.icp_weight);
    warp_update(sh, L.icp, L.rgb, L.icp_weight, threadIdx.x == 0 ? rec : nullptr, clock64());
    if(rec && threadIdx.x == 0) sh.ntr++;
    __syncwarp();
    state_store(sh, state);
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
// SO3 pre-alignment iteration (reduce.cu:953-1054 + RGBDOdometryef.cpp:300-372): map over level 2, the sequence's last
// block updates the rotation estimate.
__global__ void __launch_bounds__(kBThreads) kb_so3(const GnLaunch L, const GnSeqIn * seqs, char * states, size_t stride, char * ws, int it, slam_step_record * trace)
{
    __shared__ GnShared sh;
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
    float acc[32];
#pragma unroll
    for(int k = 0; k < 32; k++) acc[k] = 0.f;
    const int N = g.rows * g.cols;
    for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const int y = k / g.cols;
        const int x = k - y * g.cols;
        float row[4];
        const bool found = so3_pixel(a, x, y, row);
        float a11[11];
#pragma unroll
        for(int q = 0; q < 11; q++) a11[q] = acc[q];
        accumulate_so3(a11, row, found);
#pragma unroll
        for(int q = 0; q < 11; q++) acc[q] = a11[q];
    }
    if(!seq_reduce(acc, 0, 0, false, sh, ws + (size_t)seq * kWorkspaceBytes, 0)) return;
    if(threadIdx.x >= 32) return;
    char * state = states + (size_t)seq * stride;
    state_load(sh, state);
    __syncwarp();
    {
        slam_step_record * rec = (threadIdx.x == 0 && trace && sh.ntr < kGnMaxTrace) ? trace + (size_t)seq * kGnMaxTrace + sh.ntr : nullptr;
        if(rec) memset(rec, 0, sizeof(*rec));
        warp_so3_update(sh, it, rec);   // also leaves the next iteration's H, K^-1, K R
        if(threadIdx.x == 0)
        {
            if(rec) sh.ntr++;
            if(sh.stop || it == 9) sh.so3_done = 1;
        }
        __syncwarp();
    }
    state_store(sh, state);
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

// Compact correspondence of the streaming engine (the 16 bytes of a Corres slot): everything phase B needs, so that
// phase B is a dense stream with no gathers.
struct __align__(16) BCorres
{
    int zxy;       // zx | zy << 16: the pixel of the last image
    float d0;      // lastDepth there (z of the reference's cloud point)
    float diff;    // next - last intensity
    int gxy;       // dIdx | dIdy << 16 at the pixel itself
};

constexpr int kCountsOffset = 1024 * kRowWords * 4;   // per-warp correspondence counts inside a sequence's workspace (after the partial rows)

// Pixel ownership of the streaming launches: warp w of a sequence's grid owns the contiguous pixels [w * chunk, (w + 1) * chunk),
// walked 32 * PX at a time.  Its valid correspondences are compacted, in (trip, c, lane) order, into the same range of the
// sequence's correspondence buffer; phase B (same grid) streams them back.
__device__ __forceinline__ int warp_chunk(int plane, int px)
{
    const int warps = gridDim.x * (kBThreads / 32);
    const int per = (plane + warps - 1) / warps;
    return (per + 32 * px - 1) / (32 * px) * (32 * px);
}

// Phase A of a Gauss-Newton iteration: ICP products (reduce.cu:257-416) and RGB association (reduce.cu:739-867) of
// PX consecutive pixels per thread and trip.  All coalesced loads of a trip are issued first, then all gathers.
template <int PX>
__global__ void __launch_bounds__(kBThreads, PX == 4 ? 1 : 2)
kb_phase_a(const GnLaunch L, const GnSeqIn * seqs, char * states, size_t stride, char * ws, float * sums, unsigned char * cand0, size_t aux_stride, size_t cand_off,
           int lvl)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.stop_level == lvl) return;
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;   // a multiple of PX (checked by the host)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    float acc[32];
#pragma unroll
    for(int k = 0; k < 32; k++) acc[k] = 0.f;
    int cnt0 = 0, cnt1 = 0;

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
    BCorres * cimg = reinterpret_cast<BCorres *>(in.corres[lvl]);

    const int chunk = warp_chunk(plane, PX);
    const int gw = blockIdx.x * (kBThreads / 32) + wid;
    const int w0 = min(gw * chunk, plane);
    const int w1 = min(w0 + chunk, plane);
    int wcount = 0;   // correspondences this warp has written (warp-uniform)

    for(int base = w0; base < w1; base += 32 * PX)
    {
        const int p = base + lane * PX;
        const bool inside = p < w1;
        // ---- stage 1: coalesced loads
        float vx[PX], vy[PX], vz[PX], nx[PX], ny[PX], nz[PX], d1[PX];
        unsigned cm = 0, im = 0;
        if(inside)
        {
            if(L.icp)
            {
                vec_load<PX>(ia.vcurr + p, vx);
                vec_load<PX>(ia.vcurr + plane + p, vy);
                vec_load<PX>(ia.vcurr + 2 * plane + p, vz);
                vec_load<PX>(ia.ncurr + p, nx);
                vec_load<PX>(ia.ncurr + plane + p, ny);
                vec_load<PX>(ia.ncurr + 2 * plane + p, nz);
            }
            if(L.rgb)
            {
                if(PX == 4) { cm = __ldg(reinterpret_cast<const unsigned *>(cand + p)); im = __ldg(reinterpret_cast<const unsigned *>(ra.nextImage + p)); }
                else if(PX == 2) { cm = __ldg(reinterpret_cast<const unsigned short *>(cand + p)); im = __ldg(reinterpret_cast<const unsigned short *>(ra.nextImage + p)); }
                else { cm = __ldg(cand + p); im = __ldg(ra.nextImage + p); }
                vec_load<PX>(ra.nextDepth + p, d1);
            }
        }
        // ---- stage 2: projections, then every gather of the trip
        float3 vg[PX], vp[PX], np[PX];
        int o[PX];
        bool ok[PX];
        float td1[PX], d0[PX];
        int zxy[PX];
        unsigned char li[PX];
        bool rok[PX];
#pragma unroll
        for(int c = 0; c < PX; c++)
        {
            ok[c] = false;
            o[c] = 0;
            if(L.icp && inside) ok[c] = icp_project(ia, make_float3(vx[c], vy[c], vz[c]), vg[c], o[c]);
            if(!ok[c]) o[c] = 0;
            rok[c] = false;
            zxy[c] = 0;
            if(L.rgb && ((cm >> (8 * c)) & 0xff))
            {
                const int k = p + c;
                const int y = k / g.cols;
                const int x = k - y * g.cols;
                int u0, v0;
                rok[c] = rgb_project(ra, x, y, d1[c], u0, v0, td1[c]);
                if(rok[c]) zxy[c] = u0 | (v0 << 16);
            }
        }
#pragma unroll
        for(int c = 0; c < PX; c++)
        {
            if(L.icp)
            {
                vp[c] = make_float3(__ldg(ia.vprev + o[c]), __ldg(ia.vprev + plane + o[c]), __ldg(ia.vprev + 2 * plane + o[c]));
                np[c] = make_float3(__ldg(ia.nprev + o[c]), __ldg(ia.nprev + plane + o[c]), __ldg(ia.nprev + 2 * plane + o[c]));
            }
            if(L.rgb)
            {
                const int r = (zxy[c] >> 16) * g.cols + (zxy[c] & 0xffff);
                d0[c] = __ldg(ra.lastDepth + r);
                li[c] = __ldg(ra.lastImage + r);
            }
        }
        // ---- stage 3: products
        if(L.icp)
        {
#pragma unroll
            for(int c = 0; c < PX; c++)
            {
                float row[7];
                const bool found = ok[c] && icp_finish(ia, vg[c], make_float3(nx[c], ny[c], nz[c]), vp[c], np[c], row);
                if(found)
                {
                    float a29[29];
#pragma unroll
                    for(int q = 0; q < 29; q++) a29[q] = acc[q];
                    accumulate_se3(a29, row, true);
#pragma unroll
                    for(int q = 0; q < 29; q++) acc[q] = a29[q];
                }
            }
        }
        if(L.rgb)
        {
#pragma unroll
            for(int c = 0; c < PX; c++)
            {
                const bool valid = rok[c] && rgb_accept(ra, td1[c], d0[c], li[c]);
                const unsigned m = __ballot_sync(0xffffffffu, valid);
                if(valid)
                {
                    BCorres cc;
                    cc.zxy = zxy[c];
                    cc.d0 = d0[c];
                    cc.diff = __fsub_rn(static_cast<float>((im >> (8 * c)) & 0xff), static_cast<float>(li[c]));
                    const int k = p + c;
                    cc.gxy = (int)(unsigned short)__ldg(ra.dIdx + k) | ((int)__ldg(ra.dIdy + k) << 16);
                    cnt0 += 1;
                    cnt1 += (int)(cc.diff * cc.diff);
                    reinterpret_cast<int4 *>(cimg)[w0 + wcount + __popc(m & ((1u << lane) - 1u))] = *reinterpret_cast<const int4 *>(&cc);
                }
                wcount += __popc(m);
            }
        }
    }
    char * ws_seq = ws + (size_t)seq * kWorkspaceBytes;
    if(L.rgb && lane == 0) reinterpret_cast<int *>(ws_seq + kCountsOffset)[gw] = wcount;
    if(!seq_reduce(acc, cnt0, cnt1, L.rgb, sh, ws_seq, 0)) return;
    if(threadIdx.x >= 32) return;
    // phase B (or, without RGB, kb_update) picks up the ICP sums, count and sigma here: the fp64 solve is kept out of
    // this kernel because its register footprint would cap the occupancy of the streaming part
    sums[seq * 64 + threadIdx.x] = sh.total[threadIdx.x];
}

// ---- staged variant: the coalesced operands of a trip reach shared memory through cp.async, kStages - 1 trips ahead of
// their use, so the only exposed latency of a trip is the gathers'.  One pipeline per warp (a trip = kTripPx contiguous
// pixels of the warp's chunk, consumed as two half trips of 2 pixels per lane); no block-level synchronisation.
constexpr int kTripPx = 128;
constexpr int kL2Ahead = 2;

__device__ __forceinline__ void cp_async16(void * smem, const void * gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void * smem, const void * gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void * smem, const void * gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
// TMA prefetch of a contiguous range into L2 (no destination): warms the lines the gathers of a later trip will hit
__device__ __forceinline__ void prefetch_l2_bulk(const void * gmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Roles: <true, true> does both halves of phase A in one pass; <true, false> and <false, true> are the ICP and the RGB
// association as separate launches with their own register budgets (80 / 48 registers, 3 / 5 blocks per SM instead of 128 / 2;
// stage counts and block targets tuned on a B200, overridable with -DSLAM_ICP_STAGES ... for experiments):
// run on two streams they fill each other's stalls (both are bound by gather round trips, not by issue slots or bandwidth).
template <bool kIcp, bool kRgb>
struct RoleCfg
{
#ifndef SLAM_ICP_STAGES
#define SLAM_ICP_STAGES 2
#endif
#ifndef SLAM_ICP_BLOCKS
#define SLAM_ICP_BLOCKS 3
#endif
#ifndef SLAM_RGB_STAGES
#define SLAM_RGB_STAGES 3
#endif
#ifndef SLAM_RGB_BLOCKS
#define SLAM_RGB_BLOCKS 5
#endif
    static constexpr int kStagesR = (kIcp && kRgb) ? 3 : (kIcp ? SLAM_ICP_STAGES : SLAM_RGB_STAGES);
    static constexpr int kStageBytes = (kIcp ? 6 * kTripPx * 4 : 0) + (kRgb ? kTripPx * 4 + 2 * kTripPx * 2 + 2 * kTripPx : 0);
    static constexpr int kMinBlocks = (kIcp && kRgb) ? 2 : (kIcp ? SLAM_ICP_BLOCKS : SLAM_RGB_BLOCKS);
    static constexpr int kSmem = (kBThreads / 32) * kStagesR * kStageBytes;
};

template <bool kIcp, bool kRgb>
__global__ void __launch_bounds__(kBThreads, (RoleCfg<kIcp, kRgb>::kMinBlocks))
kb_phase_a_staged(const GnLaunch L, const GnSeqIn * seqs, char * states, size_t stride, char * ws, float * sums, unsigned char * cand0, size_t aux_stride,
                  size_t cand_off, int lvl)
{
    using Cfg = RoleCfg<kIcp, kRgb>;
    constexpr int kStages = Cfg::kStagesR;
    extern __shared__ __align__(16) char dyn_smem[];
    static_assert(sizeof(GnShared) <= Cfg::kSmem, "the epilogue scratch aliases the staging buffers");
    GnShared & sh = *reinterpret_cast<GnShared *>(dyn_smem);   // epilogue only, after every warp is done with its stages
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.stop_level == lvl) return;
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;   // a multiple of 4 (checked by the host)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // stage layout of this role: [vcurr xyz | ncurr xyz] (ICP) [nextDepth | dIdx | dIdy | cand | nextImage] (RGB)
    char * stage0 = dyn_smem + (size_t)wid * kStages * Cfg::kStageBytes;
    struct StageView
    {
        char * base;
        __device__ float * f(int k) const { return reinterpret_cast<float *>(base) + k * kTripPx; }                                  // k < 6: ICP planes
        __device__ float * d1() const { return reinterpret_cast<float *>(base + (kIcp ? 6 * kTripPx * 4 : 0)); }
        __device__ short * gx() const { return reinterpret_cast<short *>(base + (kIcp ? 6 * kTripPx * 4 : 0) + kTripPx * 4); }
        __device__ short * gy() const { return gx() + kTripPx; }
        __device__ unsigned char * cand() const { return reinterpret_cast<unsigned char *>(gy() + kTripPx); }
        __device__ unsigned char * img() const { return cand() + kTripPx; }
    };
    auto stage_of = [&](int t) { return StageView{stage0 + (size_t)(t % kStages) * Cfg::kStageBytes}; };

    float acc[32];
#pragma unroll
    for(int k = 0; k < 32; k++) acc[k] = 0.f;
    int cnt0 = 0, cnt1 = 0;

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
    BCorres * cimg = reinterpret_cast<BCorres *>(in.corres[lvl]);

    const int chunk = warp_chunk(plane, 4);
    const int gw = blockIdx.x * (kBThreads / 32) + wid;
    const int w0 = min(gw * chunk, plane);
    const int w1 = min(w0 + chunk, plane);
    const int ntrips = (w1 - w0 + kTripPx - 1) / kTripPx;
    int wcount = 0;   // correspondences this warp has written (warp-uniform)

    // The projective gathers of a trip land near the same pixels of the model maps: TMA-prefetch those lines into L2
    // kL2Ahead trips before they are needed, one plane per lane.
    auto warm = [&](int t) {
        if(t >= ntrips) return;
        const int b0 = w0 + t * kTripPx;
        const unsigned npx = (unsigned)min(kTripPx, w1 - b0);
        if(kIcp && lane < 6) prefetch_l2_bulk((lane < 3 ? ia.vprev + lane * plane : ia.nprev + (lane - 3) * plane) + b0, npx * 4u);
        if(kRgb && lane == 6) prefetch_l2_bulk(ra.lastDepth + b0, npx * 4u);
        if(kRgb && lane == 7) prefetch_l2_bulk(ra.lastImage + b0, (npx + 15u) & ~15u);
    };
    auto issue = [&](int t) {
        if(t < ntrips)
        {
            const int p = w0 + t * kTripPx + lane * 4;
            if(p < w1)
            {
                const StageView S = stage_of(t);
                if(kIcp)
                {
#pragma unroll
                    for(int q = 0; q < 3; q++)
                    {
                        cp_async16(S.f(q) + lane * 4, ia.vcurr + q * plane + p);
                        cp_async16(S.f(3 + q) + lane * 4, ia.ncurr + q * plane + p);
                    }
                }
                if(kRgb)
                {
                    cp_async16(S.d1() + lane * 4, ra.nextDepth + p);
                    cp_async4(S.cand() + lane * 4, cand + p);
                    cp_async4(S.img() + lane * 4, ra.nextImage + p);
                    cp_async8(S.gx() + lane * 4, ra.dIdx + p);
                    cp_async8(S.gy() + lane * 4, ra.dIdy + p);
                }
            }
        }
        cp_async_commit();
    };
    for(int t = 0; t < kL2Ahead; t++) warm(t);
#pragma unroll
    for(int t = 0; t < kStages - 1; t++) issue(t);

    for(int t = 0; t < ntrips; t++)
    {
        warm(t + kL2Ahead);
        issue(t + kStages - 1);
        cp_async_wait<kStages - 1>();
        __syncwarp();
        const StageView S = stage_of(t);
#pragma unroll 1
        for(int half = 0; half < 2; half++)
        {
            constexpr int PX = 2;
            const int q0 = half * 64 + lane * 2;
            const int p = w0 + t * kTripPx + q0;
            const bool inside = p < w1;
            const int py0 = p / g.cols, px0 = p - py0 * g.cols;   // one division per pair of pixels
            float vx[PX], vy[PX], vz[PX], nx[PX], ny[PX], nz[PX], d1[PX];
            unsigned cm = 0, im = 0;
            if(inside)
            {
                if(kIcp)
                {
                    const float2 a0 = *reinterpret_cast<const float2 *>(S.f(0) + q0), a1 = *reinterpret_cast<const float2 *>(S.f(1) + q0),
                                 a2 = *reinterpret_cast<const float2 *>(S.f(2) + q0), b0 = *reinterpret_cast<const float2 *>(S.f(3) + q0),
                                 b1 = *reinterpret_cast<const float2 *>(S.f(4) + q0), b2 = *reinterpret_cast<const float2 *>(S.f(5) + q0);
                    vx[0] = a0.x; vx[1] = a0.y; vy[0] = a1.x; vy[1] = a1.y; vz[0] = a2.x; vz[1] = a2.y;
                    nx[0] = b0.x; nx[1] = b0.y; ny[0] = b1.x; ny[1] = b1.y; nz[0] = b2.x; nz[1] = b2.y;
                }
                if(kRgb)
                {
                    cm = *reinterpret_cast<const unsigned short *>(S.cand() + q0);
                    im = *reinterpret_cast<const unsigned short *>(S.img() + q0);
                    const float2 dd = *reinterpret_cast<const float2 *>(S.d1() + q0);
                    d1[0] = dd.x; d1[1] = dd.y;
                }
            }
            // a pair of pixels whose current vertex or normal is NaN can never associate (reduce.cu:282-283), and dead pixels come in
            // regions (beyond the depth cut-off, holes): when no lane of the warp has a live pixel or a candidate in this half trip,
            // the projection, the gathers and the products are skipped as a whole
            {
                bool live = false;
                if(kIcp && inside) live = (!isnan(vx[0]) && !isnan(nx[0])) || (!isnan(vx[1]) && !isnan(nx[1]));
                if(kRgb) live = live || cm != 0u;
                if(!__any_sync(0xffffffffu, live)) continue;
            }
            float3 vg[PX], vp[PX], np[PX];
            int o[PX];
            bool ok[PX];
            float td1[PX], d0[PX];
            int zxy[PX];
            unsigned char li[PX];
            bool rok[PX];
#pragma unroll
            for(int c = 0; c < PX; c++)
            {
                ok[c] = false;
                o[c] = 0;
                if(kIcp && inside) ok[c] = icp_project(ia, make_float3(vx[c], vy[c], vz[c]), vg[c], o[c]);
                if(!ok[c]) o[c] = 0;
                rok[c] = false;
                zxy[c] = 0;
                if(kRgb && ((cm >> (8 * c)) & 0xff))
                {
                    int x = px0 + c, y = py0;
                    if(x >= g.cols)
                    {
                        x -= g.cols;
                        y++;
                    }
                    int u0, v0;
                    rok[c] = rgb_project(ra, x, y, d1[c], u0, v0, td1[c]);
                    if(rok[c]) zxy[c] = u0 | (v0 << 16);
                }
            }
#pragma unroll
            for(int c = 0; c < PX; c++)
            {
                if(kIcp)
                {
                    vp[c] = make_float3(__ldg(ia.vprev + o[c]), __ldg(ia.vprev + plane + o[c]), __ldg(ia.vprev + 2 * plane + o[c]));
                    np[c] = make_float3(__ldg(ia.nprev + o[c]), __ldg(ia.nprev + plane + o[c]), __ldg(ia.nprev + 2 * plane + o[c]));
                }
                if(kRgb)
                {
                    const int r = (zxy[c] >> 16) * g.cols + (zxy[c] & 0xffff);
                    d0[c] = __ldg(ra.lastDepth + r);
                    li[c] = __ldg(ra.lastImage + r);
                }
            }
            if(kIcp)
            {
#pragma unroll
                for(int c = 0; c < PX; c++)
                {
                    float row[7];
                    bool found = icp_finish(ia, vg[c], make_float3(nx[c], ny[c], nz[c]), vp[c], np[c], row);
                    if(!ok[c])
                    {
                        found = false;
#pragma unroll
                        for(int q = 0; q < 7; q++) row[q] = 0.f;
                    }
                    if(__any_sync(0xffffffffu, found))   // warp-uniform: lanes that miss add zeros
                    {
                        float a29[29];
#pragma unroll
                        for(int q = 0; q < 29; q++) a29[q] = acc[q];
                        accumulate_se3(a29, row, found);
#pragma unroll
                        for(int q = 0; q < 29; q++) acc[q] = a29[q];
                    }
                }
            }
            if(kRgb)
            {
#pragma unroll
                for(int c = 0; c < PX; c++)
                {
                    const bool valid = rok[c] && rgb_accept(ra, td1[c], d0[c], li[c]);
                    const unsigned m = __ballot_sync(0xffffffffu, valid);
                    if(valid)
                    {
                        BCorres cc;
                        cc.zxy = zxy[c];
                        cc.d0 = d0[c];
                        cc.diff = __fsub_rn(static_cast<float>((im >> (8 * c)) & 0xff), static_cast<float>(li[c]));
                        cc.gxy = (int)(unsigned short)S.gx()[q0 + c] | ((int)S.gy()[q0 + c] << 16);
                        cnt0 += 1;
                        cnt1 += (int)(cc.diff * cc.diff);
                        reinterpret_cast<int4 *>(cimg)[w0 + wcount + __popc(m & ((1u << lane) - 1u))] = *reinterpret_cast<const int4 *>(&cc);
                    }
                    wcount += __popc(m);
                }
            }
        }
        __syncwarp();   // the stage is free for the issue of the next iteration
    }
    char * ws_seq = ws + (size_t)seq * kWorkspaceBytes;
    if(kRgb && lane == 0) reinterpret_cast<int *>(ws_seq + kCountsOffset)[gw] = wcount;
    __syncthreads();   // every warp is done with its staging buffers: the epilogue scratch may overwrite them
    if(!seq_reduce(acc, cnt0, cnt1, kRgb, sh, ws_seq, 0)) return;
    if(threadIdx.x >= 32) return;
    // phase B (or, without RGB, kb_update) picks up the ICP sums, count and sigma here: the fp64 solve is kept out of
    // this kernel because its register footprint would cap the occupancy of the streaming part.  Split roles write
    // their own columns of the row.
    if(kIcp && kRgb)
        sums[seq * 64 + threadIdx.x] = sh.total[threadIdx.x];
    else if(kIcp)
    {
        if(threadIdx.x < 29) sums[seq * 64 + threadIdx.x] = sh.total[threadIdx.x];
    }
    else if(threadIdx.x == 29 || threadIdx.x == 30)
        sums[seq * 64 + threadIdx.x] = sh.total[threadIdx.x];
}

// ICP-only runs: the update as its own one-warp-per-sequence launch.
__global__ void __launch_bounds__(32) kb_update(const GnLaunch L, char * states, size_t stride, const float * sums, int lvl, int j, slam_step_record * trace)
{
    __shared__ GnShared sh;
    const int seq = blockIdx.x;
    if(bstate(states, stride, seq).stop_level == lvl) return;
    sh.total[threadIdx.x] = sums[seq * 64 + threadIdx.x];
    __syncwarp();
    seq_update(sh, L, states + (size_t)seq * stride, seq, lvl, j, trace);
}

// Phase B: RGB Jacobian products (reduce.cu:494-624) from the compacted correspondences of phase A, then the update.
// The lists are short (a few hundred records per phase-A warp), so this launch is a single wave: each warp walks several
// lists as one flat index space (prefix sums of their counts in shared memory).  a_blocks / a_px describe phase A's grid.
__global__ void __launch_bounds__(kBThreads, 2)
kb_phase_b(const GnLaunch L, const GnSeqIn * seqs, char * states, size_t stride, char * ws, const char * ws_counts, const float * sums, int lvl, int j,
           slam_step_record * trace, int a_blocks, int a_px)
{
    __shared__ GnShared sh;
    __shared__ int s_pre[kBThreads / 32][33];
    const int seq = blockIdx.y;
    const GnShared & st = bstate(states, stride, seq);
    if(st.stop_level == lvl) return;
    // sigmaVal and the rgbOnly early exit are re-derived from the folded count / sigma by every block, identically
    // (RGBDOdometryef.cpp:457-471)
    const int rgbSize = __float_as_int(__ldcg(sums + seq * 64 + 29));
    const int sigma = __float_as_int(__ldcg(sums + seq * 64 + 30));
    const int sel = (rgbSize != 0 && sigma == 0) ? 1 : rgbSize;
    float sigmaVal = __fsqrt_rn((float)sel);
    if(L.rgb_only)
    {
        const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
        if(rgbError > st.res.lastRGBError)
        {
            // RGBDOdometryef.cpp:460-463: the level's remaining iterations are skipped (blocks that read the flag
            // early return one line above, the others here: same outcome)
            if(blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<GnShared *>(states + (size_t)seq * stride)->stop_level = lvl;
            return;
        }
        sigmaVal = -1;
    }
    const GnSeqIn & in = seqs[seq];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    RgbStepArgs a;
    a.sigma = sigmaVal;
    a.fx = g.fx; a.fy = g.fy;
    a.sobelScale = L.sobel_scale;
    a.cols = g.cols; a.rows = g.rows;
    a.dIdx = nullptr; a.dIdy = nullptr;
    a.lastDepth = nullptr;
    a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
    a.cloud = nullptr;
    char * ws_seq = ws + (size_t)seq * kWorkspaceBytes;
    // phase A's ownership: list l covers pixels [l * chunk, ...), its records start at the same offset
    const int a_warps = a_blocks * (kBThreads / 32);
    const int per = (plane + a_warps - 1) / a_warps;
    const int chunk = (per + 32 * a_px - 1) / (32 * a_px) * (32 * a_px);
    const int b_warps = gridDim.x * (kBThreads / 32);
    const int gw = blockIdx.x * (kBThreads / 32) + wid;
    // this warp's lists: gw, gw + b_warps, ... (at most 32, checked by the host); lane k holds the count of the k-th
    const int my_list = gw + lane * b_warps;
    const int cnt = my_list < a_warps ? __ldcg(reinterpret_cast<const int *>(ws_counts + (size_t)seq * kWorkspaceBytes + kCountsOffset) + my_list) : 0;
    int pre = cnt;   // inclusive prefix over the lanes
#pragma unroll
    for(int d = 1; d < 32; d <<= 1)
    {
        const int v = __shfl_up_sync(0xffffffffu, pre, d);
        if(lane >= d) pre += v;
    }
    s_pre[wid][lane + 1] = pre;
    if(lane == 0) s_pre[wid][0] = 0;
    __syncwarp();
    const int total = s_pre[wid][32];
    const int4 * cimg = reinterpret_cast<const int4 *>(in.corres[lvl]);
    float acc[32];
#pragma unroll
    for(int k = 0; k < 32; k++) acc[k] = 0.f;
    constexpr int kInFlight = 8;   // records a lane has in flight per trip
    for(int i0 = 0; i0 < total; i0 += 32 * kInFlight)
    {
        int4 raw[kInFlight];
        bool have[kInFlight];
#pragma unroll
        for(int q = 0; q < kInFlight; q++)
        {
            const int i = i0 + q * 32 + lane;
            have[q] = i < total;
            raw[q] = make_int4(0, 0, 0, 0);
            if(have[q])
            {
                // list k with s_pre[k] <= i < s_pre[k + 1]
                int lo = 0, hi = 31;
#pragma unroll
                for(int it = 0; it < 5; it++)
                {
                    const int mid = (lo + hi + 1) >> 1;
                    if(s_pre[wid][mid] <= i) lo = mid; else hi = mid - 1;
                }
                const int l = gw + lo * b_warps;
                raw[q] = __ldcg(cimg + min(l * chunk, plane) + (i - s_pre[wid][lo]));
            }
        }
#pragma unroll
        for(int q = 0; q < kInFlight; q++)
            if(have[q])
            {
                const BCorres cc = *reinterpret_cast<const BCorres *>(&raw[q]);
                float row[7];
                rgb_row_regs(a, cc.zxy & 0xffff, cc.zxy >> 16, cc.d0, (short)(cc.gxy & 0xffff), (short)(cc.gxy >> 16), cc.diff, row);
                float a29[29];
#pragma unroll
                for(int k = 0; k < 29; k++) a29[k] = acc[k];
                accumulate_se3(a29, row, true);
#pragma unroll
                for(int k = 0; k < 29; k++) acc[k] = a29[k];
            }
    }
    if(!seq_reduce(acc, 0, 0, false, sh, ws_seq, 32)) return;
    if(threadIdx.x >= 32) return;
    sh.total[threadIdx.x] = __ldcg(sums + seq * 64 + threadIdx.x);
    __syncwarp();
    seq_update(sh, L, states + (size_t)seq * stride, seq, lvl, j, trace);
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
    d.ws = p;
    p += kWorkspaceBytes * (size_t)batch;
    d.ws2 = p;
}

static int blocks_per_seq(int nitems, int batch, int num_sms)
{
    // ~16 blocks per SM over the whole batch (tail under 1/16 of a wave), at least one trip of 256 items per block
    int want = (num_sms * 16 + batch - 1) / batch;
    const int cap = (nitems + kBThreads - 1) / kBThreads;
    if(want > cap) want = cap;
    if(want < 1) want = 1;
    if(want > 1024) want = 1024;
    return want;
}

static int ensure_kernel_attributes()
{
    static bool attr_set = false;
    if(!attr_set)
    {
        SLAM_CUDA_TRY(cudaFuncSetAttribute(kb_phase_a_staged<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RoleCfg<true, true>::kSmem));
        SLAM_CUDA_TRY(cudaFuncSetAttribute(kb_phase_a_staged<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RoleCfg<true, false>::kSmem));
        SLAM_CUDA_TRY(cudaFuncSetAttribute(kb_phase_a_staged<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RoleCfg<false, true>::kSmem));
        attr_set = true;
    }
    return SLAM_OK;
}

static bool use_graphs()
{
    static int v = -1;
    if(v < 0)
    {
        const char * e = getenv("SLAM_BATCH_GRAPH");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

static bool split_roles()
{
    static int v = -1;
    if(v < 0)
    {
        const char * e = getenv("SLAM_BATCH_SPLIT");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

static int px_variant()
{
    static int v = -1;
    if(v < 0)
    {
        const char * e = getenv("SLAM_BATCH_PX");
        v = e ? atoi(e) : 0;   // 0: cp.async-staged phase A (default); 1 / 2 / 4: register-only variants with that many pixels per thread
        if(v != 0 && v != 1 && v != 2 && v != 4) v = 0;
    }
    return v;
}

// Development aid (SLAM_BATCH_DETAIL=1): CUDA-event pair around every launch, totals per kernel printed by batch_report().
namespace {
struct DetailProf
{
    bool on = false, init = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> tag;
    double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
} g_detail;
const char * kTagName[8] = {"begin/end/level", "so3", "candidates", "phase_a", "phase_b", "update", "", ""};
void detail_fold()
{
    for(size_t i = 0; i + 1 < g_detail.ev.size(); i += 2)
    {
        cudaEventSynchronize(g_detail.ev[i + 1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_detail.ev[i], g_detail.ev[i + 1]);
        g_detail.ms[g_detail.tag[i / 2]] += ms;
        g_detail.n[g_detail.tag[i / 2]]++;
        cudaEventDestroy(g_detail.ev[i]);
        cudaEventDestroy(g_detail.ev[i + 1]);
    }
    g_detail.ev.clear();
    g_detail.tag.clear();
}
struct DetailScope
{
    cudaStream_t s;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DetailScope(cudaStream_t s_, int tag) : s(s_)
    {
        if(!g_detail.on) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        g_detail.tag.push_back(tag);
    }
    ~DetailScope()
    {
        if(!g_detail.on) return;
        cudaEventRecord(e1, s);
        g_detail.ev.push_back(e0);
        g_detail.ev.push_back(e1);
    }
};
}   // namespace

void batch_report()
{
    if(!g_detail.on) return;
    detail_fold();
    for(int t = 0; t < 6; t++)
        if(g_detail.n[t]) fprintf(stderr, "[batch detail] %-16s %8lld launches %10.3f ms total %8.2f us/launch\n", kTagName[t], g_detail.n[t], g_detail.ms[t], 1e3 * g_detail.ms[t] / g_detail.n[t]);
    for(int t = 0; t < 8; t++) g_detail.ms[t] = 0, g_detail.n[t] = 0;
}

// The launch sequence of one group of sequences [s0, s0 + B) on stream s (all per-sequence arrays are indexed from s0).
static int enqueue_group(BatchDevice & d0, const GnLaunch & L, int s0, int B, slam_step_record * trace, int * trace_count, cudaStream_t s, cudaStream_t s_rgb,
                         cudaEvent_t ev_rgb_done, cudaEvent_t ev_updated)
{
    BatchDevice d = d0;
    d.seq_in += s0;
    d.states += (size_t)s0 * d.state_stride;
    d.sums += (size_t)s0 * 64;
    d.cand0 += (size_t)s0 * d.aux_stride;
    d.ws += (size_t)s0 * kWorkspaceBytes;
    d.ws2 += (size_t)s0 * kWorkspaceBytes;
    d.results += s0;
    if(trace) trace += (size_t)s0 * kGnMaxTrace;
    if(trace_count) trace_count += s0;
    d.launches = 0;
    kb_begin<<<B, 32, 0, s>>>(L, d.seq_in, d.states, d.state_stride);
    d.launches++;
    if(L.so3)
    {
        const int nb = blocks_per_seq(L.geom[2].rows * L.geom[2].cols, B, d.num_sms);
        for(int it = 0; it < 10; it++)
        {
            DetailScope ds(s, 1);
            kb_so3<<<dim3(nb, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, it, trace);
            d.launches++;
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
        if(L.rgb && !d.cand_ready)
        {
            DetailScope ds(s, 2);
            kb_candidates<<<dim3(blocks_per_seq(plane, B, d.num_sms), B), kBThreads, 0, s>>>(L, d.seq_in, d.cand0, d.aux_stride, d.cand_off[lvl], lvl);
            d.launches++;
        }
        int px = px_variant();
        const bool staged = px == 0 && plane % 4 == 0;
        if(px == 0) px = 4;
        while(plane % px) px >>= 1;
        const int nb = blocks_per_seq(plane / px, B, d.num_sms);
        // phase B: one wave (2 blocks per SM), but every warp may walk at most 32 of phase A's per-warp lists
        int nbB = (2 * d.num_sms) / B;
        if(nbB < (nb + 31) / 32) nbB = (nb + 31) / 32;
        if(nbB < 1) nbB = 1;
        if(nbB > nb) nbB = nb;
        // split roles: the ICP and the RGB association of an ICP+RGB iteration as two launches on two streams
        const bool split = staged && L.icp && L.rgb && s_rgb != nullptr && split_roles();
        if(split)
        {
            // the RGB stream joins here: the candidates of this level and the level-begin state are ready
            SLAM_CUDA_TRY(cudaEventRecord(ev_updated, s));
        }
        for(int j = 0; j < L.iterations[lvl]; j++)
        {
#define SLAM_PHASE_A(PXV) \
    kb_phase_a<PXV><<<dim3(nb, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, d.sums, d.cand0, d.aux_stride, d.cand_off[lvl], lvl)
            {
            DetailScope ds(s, 3);
            if(split)
            {
                SLAM_CUDA_TRY(cudaStreamWaitEvent(s_rgb, ev_updated, 0));
                kb_phase_a_staged<false, true><<<dim3(nb, B), kBThreads, RoleCfg<false, true>::kSmem, s_rgb>>>(L, d.seq_in, d.states, d.state_stride, d.ws2, d.sums, d.cand0,
                                                                                                          d.aux_stride, d.cand_off[lvl], lvl);
                SLAM_CUDA_TRY(cudaEventRecord(ev_rgb_done, s_rgb));
                kb_phase_a_staged<true, false><<<dim3(nb, B), kBThreads, RoleCfg<true, false>::kSmem, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, d.sums, d.cand0,
                                                                                                      d.aux_stride, d.cand_off[lvl], lvl);
                SLAM_CUDA_TRY(cudaStreamWaitEvent(s, ev_rgb_done, 0));
                d.launches++;
            }
            else if(staged && L.icp && L.rgb)
                kb_phase_a_staged<true, true><<<dim3(nb, B), kBThreads, RoleCfg<true, true>::kSmem, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, d.sums, d.cand0,
                                                                                                    d.aux_stride, d.cand_off[lvl], lvl);
            else if(staged && L.icp)
                kb_phase_a_staged<true, false><<<dim3(nb, B), kBThreads, RoleCfg<true, false>::kSmem, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, d.sums, d.cand0,
                                                                                                      d.aux_stride, d.cand_off[lvl], lvl);
            else if(staged)
                kb_phase_a_staged<false, true><<<dim3(nb, B), kBThreads, RoleCfg<false, true>::kSmem, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, d.sums, d.cand0,
                                                                                                      d.aux_stride, d.cand_off[lvl], lvl);
            else if(px == 4) SLAM_PHASE_A(4); else if(px == 2) SLAM_PHASE_A(2); else SLAM_PHASE_A(1);
            }
            d.launches++;
            if(L.rgb)
            {
                DetailScope ds(s, 4);
                kb_phase_b<<<dim3(nbB, B), kBThreads, 0, s>>>(L, d.seq_in, d.states, d.state_stride, d.ws, split ? d.ws2 : d.ws, d.sums, lvl, j, trace, nb, px);
                d.launches++;
                if(split) SLAM_CUDA_TRY(cudaEventRecord(ev_updated, s));   // the next iteration's RGB association needs the new pose
            }
            else
            {
                DetailScope ds(s, 5);
                kb_update<<<B, 32, 0, s>>>(L, d.states, d.state_stride, d.sums, lvl, j, trace);
                d.launches++;
            }
        }
    }
    kb_end<<<B, 32, 0, s>>>(L, d.states, d.state_stride, d.results, trace_count);
    d.launches++;
    SLAM_CUDA_TRY(cudaGetLastError());
    d0.launches += d.launches;
    return SLAM_OK;
}

// Whole batch.  With 24 or more sequences the batch runs as two (SLAM_BATCH_GROUPS: 1..4) independent groups on their own
// streams: while one group sits in a latency-bound stretch (phase B, the fp64 solves, the SO3 iterations of the small level)
// the other group's streaming launches keep the memory system busy.  Results do not depend on the grouping.
int batch_enqueue(BatchDevice & d, const GnLaunch & L, const GnSeqIn * h_seq_in_pinned, GnResult * h_results, slam_step_record * trace, int * trace_count,
                  cudaStream_t s, std::vector<cudaEvent_t> * prof_events)
{
    const int B = L.batch;
    if(!L.trace) trace = nullptr, trace_count = nullptr;
    if(!g_detail.init)
    {
        g_detail.init = true;
        const char * e = getenv("SLAM_BATCH_DETAIL");
        g_detail.on = e && atoi(e) != 0;
    }
    if(g_detail.on && g_detail.ev.size() > 8192) detail_fold();
    int groups = B >= 24 ? 2 : 1;   // measured on a B200: two groups pay off from a few dozen sequences on
    if(const char * e = getenv("SLAM_BATCH_GROUPS")) groups = atoi(e);
    if(groups < 1) groups = 1;
    if(groups > 4) groups = 4;
    if(groups > B) groups = B;
    if(g_detail.on) groups = 1;   // per-kernel timing wants the launches back to back
    while((int)d.side.size() < groups - 1)
    {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        SLAM_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        d.side.push_back(st);
        d.side_done.push_back(ev);
    }
    while((int)d.role.size() < groups)   // per group: the stream of the RGB association role + its two events
    {
        cudaStream_t st = nullptr;
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        SLAM_CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        SLAM_CUDA_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        d.role.push_back(st);
        d.role_done.push_back(e1);
        d.role_go.push_back(e2);
    }
    if(!d.fork) SLAM_CUDA_TRY(cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming));
    if(int rc = ensure_kernel_attributes()) return rc;

    // the launch sequence of a step: copy in, fork the groups, join, copy out
    auto body = [&]() -> int {
        SLAM_CUDA_TRY(cudaMemcpyAsync(d.seq_in, h_seq_in_pinned, sizeof(GnSeqIn) * B, cudaMemcpyHostToDevice, s));
        if(groups > 1) SLAM_CUDA_TRY(cudaEventRecord(d.fork, s));
        for(int gidx = 0; gidx < groups; gidx++)
        {
            const int s0 = (int)((long long)B * gidx / groups), s1 = (int)((long long)B * (gidx + 1) / groups);
            cudaStream_t st = gidx == 0 ? s : d.side[gidx - 1];
            if(gidx > 0) SLAM_CUDA_TRY(cudaStreamWaitEvent(st, d.fork, 0));
            if(int rc = enqueue_group(d, L, s0, s1 - s0, trace, trace_count, st, g_detail.on ? nullptr : d.role[gidx], d.role_done[gidx], d.role_go[gidx])) return rc;
            if(gidx > 0) SLAM_CUDA_TRY(cudaEventRecord(d.side_done[gidx - 1], st));
        }
        for(int gidx = 1; gidx < groups; gidx++) SLAM_CUDA_TRY(cudaStreamWaitEvent(s, d.side_done[gidx - 1], 0));
        SLAM_CUDA_TRY(cudaMemcpyAsync(h_results, d.results, sizeof(GnResult) * B, cudaMemcpyDeviceToHost, s));
        return SLAM_OK;
    };

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(prof_events)   // CUDA-event bracket of the whole engine section of this step
    {
        SLAM_CUDA_TRY(cudaEventCreate(&e0));
        SLAM_CUDA_TRY(cudaEventCreate(&e1));
        SLAM_CUDA_TRY(cudaEventRecord(e0, s));
    }
    if(use_graphs() && !g_detail.on && s != nullptr)   // (the legacy default stream cannot be captured)
    {
        // Every argument of every launch of a step is a constant of (mode, geometry, handle): the sequence is captured once
        // per such key as a CUDA graph and replayed, which removes ~75 stream launches and ~80 event operations per step
        // from the host's critical path (what small batches are bound by).
        std::vector<char> key(sizeof(GnLaunch) + 8 * sizeof(void *));
        memset(key.data(), 0, key.size());
        memcpy(key.data(), &L, sizeof(GnLaunch));
        const void * extra[8] = {trace, trace_count, h_seq_in_pinned, h_results, (const void *)s, (const void *)(size_t)groups, (const void *)(size_t)(d.cand_ready ? 1 : 0),
                                 (const void *)(size_t)B};
        memcpy(key.data() + sizeof(GnLaunch), extra, sizeof(extra));
        BatchDevice::GraphEntry * hit = nullptr;
        for(auto & g : d.graphs)
            if(g.key == key) hit = &g;
        if(!hit)
        {
            if(d.graphs.size() >= 16)   // modes come and go in tests: drop the oldest
            {
                cudaGraphExecDestroy(d.graphs.front().exec);
                d.graphs.erase(d.graphs.begin());
            }
            const long long before = d.launches;
            SLAM_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = body();
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            if(rc) { if(graph) cudaGraphDestroy(graph); return rc; }
            SLAM_CUDA_TRY(ce);
            BatchDevice::GraphEntry entry;
            entry.key = key;
            entry.launches = d.launches - before;
            d.launches = before;
            SLAM_CUDA_TRY(cudaGraphInstantiate(&entry.exec, graph, 0));
            cudaGraphDestroy(graph);
            d.graphs.push_back(entry);
            hit = &d.graphs.back();
        }
        SLAM_CUDA_TRY(cudaGraphLaunch(hit->exec, s));
        d.launches += hit->launches;
    }
    else if(int rc = body())
        return rc;
    if(prof_events)
    {
        SLAM_CUDA_TRY(cudaEventRecord(e1, s));
        prof_events->push_back(e0);
        prof_events->push_back(e1);
    }
    return SLAM_OK;
}

void batch_release(BatchDevice & d)
{
    for(auto & g : d.graphs) cudaGraphExecDestroy(g.exec);
    d.graphs.clear();
    for(auto st : d.side) cudaStreamDestroy(st);
    for(auto ev : d.side_done) cudaEventDestroy(ev);
    for(auto st : d.role) cudaStreamDestroy(st);
    for(auto ev : d.role_done) cudaEventDestroy(ev);
    for(auto ev : d.role_go) cudaEventDestroy(ev);
    d.role.clear();
    d.role_done.clear();
    d.role_go.clear();
    if(d.fork) cudaEventDestroy(d.fork);
    d.side.clear();
    d.side_done.clear();
    d.fork = nullptr;
}

}   // namespace slam
