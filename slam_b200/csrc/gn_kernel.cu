// Device-resident Gauss-Newton loop of the tracker: ONE persistent, cooperatively
// launched kernel runs the SO3 pre-alignment and the coarse-to-fine ICP+RGB iterations of
// RGBDOdometryef::getIncrementalTransformation (src/odom/RGBDOdometryef.cpp:267-595),
// including the 3x3 / 6x6 solves and the pose updates that the reference does on the host
// between ~40 kernel launches, ~25 cudaDeviceSynchronize and ~20 cudaMalloc/cudaFree pairs
// per frame.
//
// Layout of the computation
//   * the CTAs of the grid are split into groups of G CTAs; a group owns one sequence
//     (batch == 1: one group of all 148 CTAs; batch > 1: independent groups, so
//     independent sequences progress concurrently with no inter-group traffic);
//   * every thread owns a fixed set of pixels per level ("slots": k = gtid + m * gthreads),
//     so per-pixel state that does not depend on the pose (RGB candidate flag, depth,
//     intensity, gradients) is loaded ONCE per level into registers, and the photometric
//     correspondences found in phase A stay in registers for phase B;
//   * loads are issued for all of a thread's slots before any is consumed (the working set
//     lives in the 126 MB L2, so what matters is round trips, not bytes);
//   * a step is "map" (accumulate the 29/11 products in registers) + "publish" (transposed
//     warp reduction -> shared memory -> one 64-float partial row per CTA in global memory)
//     + a group barrier (release/acquire counter) + "fold" (every CTA re-reads the G partial
//     rows in a fixed order, so all CTAs hold bit-identical sums);
//   * warp 0 of EVERY CTA then solves the normal equations redundantly in fp64
//     (small_math.hpp): the 6x6 LDL^T on one lane, everything around it (combining the
//     systems, resultRt update, K R K^-1, K t, current pose) spread over the lanes in
//     shared-memory stages.  No host round trip, no second launch, no broadcast step;
//   * partial rows are double-buffered by step parity, which makes one barrier per
//     reduction sufficient.
// Per-pixel arithmetic is pixel_ops.cuh, shared with the single-launch operator kernels.
#include <cfloat>
#include <cstring>
#include "gn_kernel.cuh"
#include "gn_scalar.cuh"

namespace slam {

constexpr int kIcpChunk = 3;    // ICP gathers in flight per thread (register budget)

// ctr points at the 64-bit barrier word of the group; the arrival counter is its high half.
__device__ __forceinline__ void group_barrier(unsigned long long * ctr, unsigned & target, unsigned G)
{
    __syncthreads();
    if(G > 1 && threadIdx.x == 0)
    {
        unsigned * hi = reinterpret_cast<unsigned *>(ctr) + 1;
        target += G;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(hi), "r"(1u) : "memory");
        unsigned v;
        do
        {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(hi) : "memory");
        } while((int)(v - target) < 0);
    }
    __syncthreads();
}

// Barrier + all-reduce of one small unsigned per CTA in the same round trip: thread 0 adds `mine` to the low half of the
// barrier word (relaxed) before it arrives on the high half (release), and polls the whole word: the value that shows the
// last arrival also holds every CTA's contribution (a later contribution needs another barrier in between).  The low
// half is a running sum modulo 2^32; `running` carries the previous total.  Returns the sum over the group in sh_out
// (thread 0 writes it before the closing __syncthreads()).
__device__ __forceinline__ void group_barrier_sum(unsigned long long * ctr, unsigned & target, unsigned G, unsigned mine, unsigned & running, int * sh_out)
{
    __syncthreads();
    if(threadIdx.x == 0)
    {
        unsigned * lo = reinterpret_cast<unsigned *>(ctr);
        target += G;
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(lo), "r"(mine) : "memory");
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(lo + 1), "r"(1u) : "memory");
        unsigned long long v;
        do
        {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
        } while((int)((unsigned)(v >> 32) - target) < 0);
        const unsigned now = (unsigned)v;
        *sh_out = (int)(now - running);
        running = now;
    }
    __syncthreads();
}

// Block sum of up to 32 per-thread floats (v[NV..31] must be 0) -> dst[0..31] (global partial row
// of this CTA).  Transposed warp reduction, then 32 threads add the per-warp rows.
// The caller must reach a __syncthreads() (e.g. group_barrier) before sh.red is reused.
// With count_cols, slots 29 / 30 / 31 carry exact small integers as floats (RGB correspondence count, low
// 12 bits and high bits of the squared-residual sum); they are recombined into the two int32 columns
// 29 (count) and 30 (sigma) of the partial row.
// Returns (thread 0, with count_cols) the CTA's contribution to the mid-iteration all-reduce: count | (sigma != 0) << 24.
__device__ __forceinline__ unsigned cta_publish32(float (&v)[32], GnShared & sh, float * dst, const bool count_cols = false)
{
    unsigned word = 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float s = warp_reduce_scatter32(v);
    sh.red[wid * 32 + lane] = s;
    __syncthreads();
    if(threadIdx.x < 32)
    {
        float total = 0.f;
#pragma unroll 16
        for(int w = 0; w < nw; w++) total += sh.red[w * 32 + threadIdx.x];
        if(count_cols)
        {
            const float hi = __shfl_sync(0xffffffffu, total, 31);
            if(threadIdx.x == 29) total = __int_as_float((int)total);
            if(threadIdx.x == 30) total = __int_as_float((int)total + ((int)hi << 12));
            const unsigned cnt = (unsigned)__float_as_int(__shfl_sync(0xffffffffu, total, 29));
            const unsigned sig = (unsigned)__float_as_int(__shfl_sync(0xffffffffu, total, 30));
            word = cnt + (sig != 0u ? (1u << 24) : 0u);
        }
        dst[threadIdx.x] = total;
    }
    return word;
}

// Same for two per-thread ints (count, sigma) -> dst[0..1].
__device__ __forceinline__ void cta_publish_int2(int c0, int c1, GnShared & sh, int * dst)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    int * red = reinterpret_cast<int *>(sh.red);
    if(lane == 0)
    {
        red[wid * 2] = c0;
        red[wid * 2 + 1] = c1;
    }
    __syncthreads();
    if(threadIdx.x < 2)
    {
        int t = 0;
        for(int w = 0; w < nw; w++) t += red[w * 2 + threadIdx.x];
        dst[threadIdx.x] = t;
    }
    __syncthreads();
}

// Fold the G partial rows (64 columns each) of this group into sh.total, in a fixed order (so
// every CTA of the group gets bit-identical sums).  All of a thread's loads are issued before the
// first use: one L2 round trip for the whole fold.  Columns 29 and 30 hold the integer count /
// sigma of the RGB residual (bit patterns).
__device__ __forceinline__ void fold_partials(GnShared & sh, const float * rows, int G)
{
    constexpr int kVecPerRow = kGnPartialStride / 4;                       // 16 float4 per row
    constexpr int kMaxPasses = (kGnMaxCtas * kVecPerRow) / kGnThreads;     // 8
    const int nvec = G * kVecPerRow;
    const int c4 = threadIdx.x % kVecPerRow;    // which float4 column
    const int r0 = threadIdx.x / kVecPerRow;    // first row of this thread; stride 32 rows
    float4 v[kMaxPasses];
#pragma unroll
    for(int m = 0; m < kMaxPasses; m++)
    {
        const int q = threadIdx.x + m * kGnThreads;
        v[m] = (q < nvec) ? __ldcg(reinterpret_cast<const float4 *>(rows) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 acc = v[0];
    const bool has_int = (c4 == 7);   // columns 28..31 -> .y = count, .z = sigma
#pragma unroll
    for(int m = 1; m < kMaxPasses; m++)
    {
        acc.x += v[m].x;
        acc.w += v[m].w;
        if(has_int)
        {
            acc.y = __int_as_float(__float_as_int(acc.y) + __float_as_int(v[m].y));
            acc.z = __int_as_float(__float_as_int(acc.z) + __float_as_int(v[m].z));
        }
        else
        {
            acc.y += v[m].y;
            acc.z += v[m].z;
        }
    }
    reinterpret_cast<float4 *>(sh.red)[r0 * kVecPerRow + c4] = acc;
    __syncthreads();
    if(threadIdx.x < kGnPartialStride)
    {
        const bool is_int = (threadIdx.x == 29 || threadIdx.x == 30);
        float ft = 0.f;
        int it = 0;
#pragma unroll 8
        for(int k = 0; k < kGnThreads / kVecPerRow; k++)
        {
            const float x = sh.red[k * kGnPartialStride + threadIdx.x];
            if(is_int)
                it += __float_as_int(x);
            else
                ft += x;
        }
        sh.total[threadIdx.x] = is_int ? __int_as_float(it) : ft;
    }
    __syncthreads();
}

// count / sigma of the whole image right after the phase-A barrier (warp 0 only): all loads in flight at once.
__device__ __forceinline__ void fold_count_sigma(GnShared & sh, const float * rows, int G)
{
    constexpr int kMax = kGnMaxCtas / 32;
    int c0 = 0, c1 = 0;
    float a[kMax], b[kMax];
#pragma unroll
    for(int j = 0; j < kMax; j++)
    {
        const int r = (int)threadIdx.x + 32 * j;
        a[j] = (r < G) ? __ldcg(rows + r * kGnPartialStride + 29) : 0.f;
        b[j] = (r < G) ? __ldcg(rows + r * kGnPartialStride + 30) : 0.f;
    }
#pragma unroll
    for(int j = 0; j < kMax; j++)
    {
        c0 += __float_as_int(a[j]);
        c1 += __float_as_int(b[j]);
    }
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    if(threadIdx.x == 0)
    {
        sh.total[29] = __int_as_float(c0);
        sh.total[30] = __int_as_float(c1);
    }
}

// The buffers of one pyramid level of one sequence.  Read field by field from the kernel parameter (one sequence:
// constant-bank loads, no local copy of the block) or from the per-sequence array in global memory.
struct LevelPtrs
{
    const float * vcurr, * ncurr, * vprev, * nprev, * lastDepth, * nextDepth;
    const unsigned char * lastImage, * nextImage, * lastNextImage;
    const short * dIdx, * dIdy;
    Corres * corres;
};
#define GN_LEVEL_FIELDS(S) \
    p.vcurr = (S).vcurr[lvl]; p.ncurr = (S).ncurr[lvl]; p.vprev = (S).vprev[lvl]; p.nprev = (S).nprev[lvl]; \
    p.lastDepth = (S).lastDepth[lvl]; p.nextDepth = (S).nextDepth[lvl]; p.lastImage = (S).lastImage[lvl]; \
    p.nextImage = (S).nextImage[lvl]; p.lastNextImage = (S).lastNextImage[lvl]; p.dIdx = (S).dIdx[lvl]; \
    p.dIdy = (S).dIdy[lvl]; p.corres = (S).corres[lvl];
__device__ __forceinline__ LevelPtrs level_ptrs(const bool one, const GnSeqIn & seq0, const GnSeqIn * seqs, int seq, int lvl)
{
    LevelPtrs p;
    if(one)
    {
        GN_LEVEL_FIELDS(seq0)
    }
    else
    {
        GN_LEVEL_FIELDS(seqs[seq])
    }
    return p;
}

// Pose-independent half of the photometric association (reduce.cu:780-807) for NS slots of a thread at once, with every
// load of every slot issued before the first use (ONE round trip to L2 per batch instead of a chain of early-exit
// branches per pixel): the clipped 4x4 all-nonzero window of nextImage, the gradient pair (from dIdx/dIdy, or derived
// from the same window with the arithmetic of utils.cu:582-606 when no derivative images were made), nextDepth.
template <int C0, int NS>
__device__ __forceinline__ void candidate_batch(const ResidualArgs & a, const bool derive, const int gtid, const int gthreads, const int nslots, const int plane,
                                                unsigned & cand, float (&c_d1)[kSlotChunk], float (&c_img)[kSlotChunk], short (&c_gx)[kSlotChunk],
                                                short (&c_gy)[kSlotChunk])
{
    unsigned char w[NS][16];
    float d1[NS];
    short gxl[NS], gyl[NS];
    int px[NS], py[NS];
    bool live[NS];
#pragma unroll
    for(int s = 0; s < NS; s++)
    {
        const int c = C0 + s;
        const int k = gtid + c * gthreads;
        live[s] = (c < nslots) && (k < plane);
        const int kk = live[s] ? k : 0;
        py[s] = kk / a.cols;
        px[s] = kk - py[s] * a.cols;
#pragma unroll
        for(int r = 0; r < 4; r++)
#pragma unroll
            for(int q = 0; q < 4; q++)
            {
                const int u = min(max(py[s] - 2 + r, 0), a.rows - 1), v = min(max(px[s] - 2 + q, 0), a.cols - 1);
                w[s][r * 4 + q] = __ldg(a.nextImage + u * a.cols + v);
            }
        d1[s] = __ldg(a.nextDepth + kk);
        gxl[s] = derive ? (short)0 : __ldg(a.dIdx + kk);
        gyl[s] = derive ? (short)0 : __ldg(a.dIdy + kk);
    }
#pragma unroll
    for(int s = 0; s < NS; s++)
    {
        const int c = C0 + s;
        const int x = px[s], y = py[s];
        bool ok = live[s] && (x < a.cols - 5 && y < a.rows - 1);
        // window taps the reference's clipped loops never visit (rows / columns below 0) do not vote
#pragma unroll
        for(int r = 0; r < 4; r++)
#pragma unroll
            for(int q = 0; q < 4; q++)
            {
                const bool visited = (y - 2 + r >= 0) && (x - 2 + q >= 0);
                ok = ok && (!visited || w[s][r * 4 + q] > 0);
            }
        short gx = gxl[s], gy = gyl[s];
        if(derive)
        {
            if(x >= 1 && y >= 1 && x < a.cols - 1 && y < a.rows - 1)
            {
                // interior: the nine taps are window entries (r + 1, q + 1), accumulated in the reference's order
                const float fgx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
                const float fgy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
                float dxVal = 0, dyVal = 0;
#pragma unroll
                for(int t = 0; t < 9; t++)
                {
                    const float v = (float)w[s][(t / 3 + 1) * 4 + (t % 3 + 1)];
                    dxVal = __fmaf_rn(v, fgx[8 - t], dxVal);
                    dyVal = __fmaf_rn(v, fgy[8 - t], dyVal);
                }
                gx = (short)dxVal;
                gy = (short)dyVal;
            }
            else if(ok)
                derivative_pixel(a.nextImage, a.rows, a.cols, x, y, gx, gy);   // image border: the clipped loop itself
        }
        const int valx = gx, valy = gy;
        const float mTwo = (valx * valx) + (valy * valy);
        ok = ok && (mTwo >= a.minScale) && !isnan(d1[s]);
        if(ok)
        {
            cand |= 1u << c;
            c_d1[c] = d1[s];
            c_img[c] = static_cast<float>(w[s][2 * 4 + 2]);   // the pixel itself
            c_gx[c] = gx;
            c_gy[c] = gy;
        }
    }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kGnThreads, 1)
k_gn_persistent(const GnLaunch L, GnCtl * ctl, const GnSeqIn * seqs, const GnSeqIn seq0, float * partials, GnResult * results, slam_step_record * trace,
                int * trace_count, const int G, const int groups, GnResult * host_results, unsigned * host_flags, const unsigned host_seqno)
{
    __shared__ GnShared sh;

    const int group = blockIdx.x / G;
    const int rank = blockIdx.x - group * G;
    if(group >= groups) return;

    unsigned long long * bar = &ctl->barrier[group];
    unsigned sum_running = ctl->sum_base[group];   // thread 0's copy is the one that is used
    // the counter is never reset: every launch starts from the value the previous launch ended with, published in
    // ctl->base by the group leader (stable for the whole launch: it is rewritten only at the very end)
    unsigned target = ctl->base[group];
    unsigned step = 0;
    // partial rows of this group: [parity][rank][64]
    float * gpart = partials + (size_t)group * 2 * G * kGnPartialStride;

    const long long t_start = clock64();
#define GN_STAMP(rec, idx) do { if(rec) (rec)->t_cycles[idx] = (unsigned)(clock64() - t_start); } while(0)
    const int gtid = rank * blockDim.x + threadIdx.x;
    const int gthreads = G * blockDim.x;
    const bool leader = (rank == 0 && threadIdx.x == 0);
    const bool warp0 = threadIdx.x < 32;

    for(int seq = group; seq < L.batch; seq += groups)
    {
        slam_step_record * tr = (L.trace && leader) ? trace + (size_t)seq * kGnMaxTrace : nullptr;
        int ntr = 0;

        if(threadIdx.x == 0)
        {
            float Rp[9], tp[3];
            if(L.batch == 1)
            {
#pragma unroll
                for(int k = 0; k < 9; k++) Rp[k] = seq0.Rprev[k];
#pragma unroll
                for(int k = 0; k < 3; k++) tp[k] = seq0.tprev[k];
            }
            else
            {
#pragma unroll
                for(int k = 0; k < 9; k++) Rp[k] = seqs[seq].Rprev[k];
#pragma unroll
                for(int k = 0; k < 3; k++) tp[k] = seqs[seq].tprev[k];
            }
            seq_begin_pose(sh, Rp, tp);
        }
        __syncthreads();

        // ------------------------------------------------ SO3 pre-alignment, level 2
        if(L.so3)
        {
            const LevelGeom g = L.geom[2];
            const int N = g.rows * g.cols;
            const LevelPtrs P2 = level_ptrs(L.batch == 1, seq0, seqs, seq, 2);
            if(threadIdx.x == 0)
            {
                level_begin(sh, g);
                so3_prepare(sh);
            }
            __syncthreads();
            for(int it = 0; it < 10; it++)
            {
                So3Args a;
                a.lastImage = P2.lastNextImage;
                a.nextImage = P2.nextImage;
                a.imageBasis = mat3_from(sh.so3H);
                a.kinv = mat3_from(sh.so3Kinv);
                a.krlr = mat3_from(sh.so3KR);
                a.cols = g.cols;
                a.rows = g.rows;

                float acc[32];
#pragma unroll
                for(int k = 0; k < 32; k++) acc[k] = 0.f;
                for(int k = gtid; k < N; k += gthreads)
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
                float * myrow = gpart + ((step & 1) * G + rank) * kGnPartialStride;
                cta_publish32(acc, sh, myrow);
                group_barrier(bar, target, G);
                fold_partials(sh, gpart + (step & 1) * G * kGnPartialStride, G);
                step++;

                if(warp0)
                {
                    slam_step_record * rec = (threadIdx.x == 0 && tr && ntr < kGnMaxTrace) ? &tr[ntr] : nullptr;
                    if(rec) memset(rec, 0, sizeof(*rec));
                    warp_so3_update(sh, it, rec);   // also leaves the next iteration's H, K^-1, K R in shared memory
                    if(rec) ntr++;
                }
                __syncthreads();
                if(sh.stop) break;
            }
        }

        if(threadIdx.x == 0)
        {
            for(int k = 0; k < 16; k++) sh.resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
            if(L.so3)
                for(int x = 0; x < 3; x++)
                    for(int y = 0; y < 3; y++) sh.resultRt[x * 4 + y] = sh.resultR[x * 3 + y];
        }
        __syncthreads();

        // ------------------------------------------------ coarse-to-fine ICP + RGB
        for(int lvl = L.levels - 1; lvl >= 0; lvl--)
        {
            const LevelGeom g = L.geom[lvl];
            const int plane = g.rows * g.cols;
            const int nslots = (plane + gthreads - 1) / gthreads;
            // all of this thread's pixels fit one register-resident chunk (always true for one 640x480 sequence on a full GPU)
            const bool single = nslots <= kSlotChunk;
            const LevelPtrs P = level_ptrs(L.batch == 1, seq0, seqs, seq, lvl);
            if(warp0)
            {
                if(threadIdx.x == 0)
                {
                    sh.res.lastRGBError = FLT_MAX;
                    level_begin(sh, g);
                }
                __syncwarp();
                warp_prepare(sh, false);   // krk / kt of the first iteration of this level (Rcurr/tcurr carry over)
            }

            // ---- per-level, pose-independent pixel state (registers): RGB candidate test (reduce.cu:780-807) and its operands
            unsigned cand = 0;              // bit c: slot c is an RGB candidate
            float c_d1[kSlotChunk];         // nextDepth
            float c_img[kSlotChunk];        // nextImage as float
            short c_gx[kSlotChunk], c_gy[kSlotChunk];
#pragma unroll
            for(int c = 0; c < kSlotChunk; c++)
            {
                c_d1[c] = 0.f; c_img[c] = 0.f; c_gx[c] = 0; c_gy[c] = 0;
            }
            if(L.rgb && single)
            {
                ResidualArgs a;
                a.minScale = L.min_scale[lvl];
                a.dIdx = P.dIdx; a.dIdy = P.dIdy;
                a.nextDepth = P.nextDepth;
                a.nextImage = P.nextImage;
                a.cols = g.cols; a.rows = g.rows;
                candidate_batch<0, 3>(a, L.derive_gradients, gtid, gthreads, nslots, plane, cand, c_d1, c_img, c_gx, c_gy);
                if(nslots > 3) candidate_batch<3, kSlotChunk - 3>(a, L.derive_gradients, gtid, gthreads, nslots, plane, cand, c_d1, c_img, c_gx, c_gy);
            }
            __syncthreads();

            for(int j = 0; j < L.iterations[lvl]; j++)
            {
                slam_step_record * rec = nullptr;
                if(threadIdx.x == 0 && tr && ntr < kGnMaxTrace)
                {
                    rec = &tr[ntr];
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
                    GN_STAMP(rec, 0);
                    GN_STAMP(rec, 1);
                }

                float * rowsA = gpart + (step & 1) * G * kGnPartialStride;
                float * myrow = rowsA + rank * kGnPartialStride;

                // phase A -> B state of this thread's slots (registers, single-chunk case)
                unsigned valid_mask = 0;
                int r_zxy[kSlotChunk];      // (zy << 16) | zx : pixel in the last image
                float r_diff[kSlotChunk];   // next - last intensity
                float r_d0[kSlotChunk];     // lastDepth at that pixel
#pragma unroll
                for(int c = 0; c < kSlotChunk; c++)
                {
                    r_zxy[c] = 0; r_diff[c] = 0.f; r_d0[c] = 0.f;
                }

                unsigned mid_word = 0;
                float sigma_now = 0.f;
                // ---------------- phase A: ICP products + RGB association
                {
                    float acc[32];
#pragma unroll
                    for(int k = 0; k < 32; k++) acc[k] = 0.f;
                    if(L.icp)
                    {
                        IcpArgs a;
                        a.Rcurr = mat3_from(sh.Rcurr);
                        a.tcurr = make_float3(sh.tcurr[0], sh.tcurr[1], sh.tcurr[2]);
                        a.Rprev_inv = mat3_from(sh.Rprev_inv);
                        a.tprev = make_float3(sh.tprev[0], sh.tprev[1], sh.tprev[2]);
                        a.fx = g.fx; a.fy = g.fy; a.cx = g.cx; a.cy = g.cy;
                        a.distThres = L.dist_thresh;
                        a.angleThres = L.angle_thresh;
                        a.cols = g.cols;
                        a.rows = g.rows;
                        a.vcurr = P.vcurr; a.ncurr = P.ncurr; a.vprev = P.vprev; a.nprev = P.nprev;
                        for(int m0 = 0; m0 < nslots; m0 += kIcpChunk)
                        {
                            float3 vg[kIcpChunk], nc[kIcpChunk], vp[kIcpChunk], np[kIcpChunk];
                            int o[kIcpChunk];
                            bool ok[kIcpChunk];
                            // 1) current vertex + normal of every slot of the chunk (coalesced across the warp)
#pragma unroll
                            for(int c = 0; c < kIcpChunk; c++)
                            {
                                const int k = gtid + (m0 + c) * gthreads;
                                ok[c] = (m0 + c < nslots) && (k < plane);
                                const int kk = ok[c] ? k : 0;
                                vg[c] = make_float3(__ldg(a.vcurr + kk), __ldg(a.vcurr + plane + kk), __ldg(a.vcurr + 2 * plane + kk));
                                nc[c] = make_float3(__ldg(a.ncurr + kk), __ldg(a.ncurr + plane + kk), __ldg(a.ncurr + 2 * plane + kk));
                            }
                            // 2) project all, 3) issue all gathers
#pragma unroll
                            for(int c = 0; c < kIcpChunk; c++)
                            {
                                float3 g3;
                                const bool inb = icp_project(a, vg[c], g3, o[c]);
                                vg[c] = g3;
                                ok[c] = ok[c] && inb;
                                if(!ok[c]) o[c] = 0;
                            }
#pragma unroll
                            for(int c = 0; c < kIcpChunk; c++)
                            {
                                vp[c] = make_float3(__ldg(a.vprev + o[c]), __ldg(a.vprev + plane + o[c]), __ldg(a.vprev + 2 * plane + o[c]));
                                np[c] = make_float3(__ldg(a.nprev + o[c]), __ldg(a.nprev + plane + o[c]), __ldg(a.nprev + 2 * plane + o[c]));
                            }
                            // 4) gates, rows, products
#pragma unroll
                            for(int c = 0; c < kIcpChunk; c++)
                            {
                                float row[7];
                                const bool found = icp_finish(a, vg[c], nc[c], vp[c], np[c], row) && ok[c];
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
                    }
                    int cnt0 = 0, cnt1 = 0;
                    if(L.rgb)
                    {
                        ResidualArgs a;
                        a.minScale = L.min_scale[lvl];
                        a.dIdx = P.dIdx; a.dIdy = P.dIdy;
                        a.lastDepth = P.lastDepth; a.nextDepth = P.nextDepth;
                        a.lastImage = P.lastImage; a.nextImage = P.nextImage;
                        a.maxDepthDelta = L.max_depth_delta;
                        a.kt = make_float3(sh.kt[0], sh.kt[1], sh.kt[2]);
                        a.krkinv = mat3_from(sh.krk);
                        a.cols = g.cols; a.rows = g.rows;
                        Corres * cimg = P.corres;
                        if(single)
                        {
                            int o0[kSlotChunk];
                            float td1[kSlotChunk];
                            unsigned inb = 0;
                            // 1) warp every candidate, 2) issue the gathers, 3) gates
#pragma unroll
                            for(int c = 0; c < kSlotChunk; c++)
                            {
                                o0[c] = 0;
                                td1[c] = 0.f;
                                if((cand >> c) & 1u)
                                {
                                    const int k = gtid + c * gthreads;
                                    const int i = k / g.cols;
                                    int u0, v0;
                                    if(rgb_project(a, k - i * g.cols, i, c_d1[c], u0, v0, td1[c]))
                                    {
                                        inb |= 1u << c;
                                        o0[c] = v0 * g.cols + u0;
                                        r_zxy[c] = (v0 << 16) | u0;
                                    }
                                }
                            }
                            unsigned char lst[kSlotChunk];
#pragma unroll
                            for(int c = 0; c < kSlotChunk; c++)
                            {
                                r_d0[c] = __ldg(a.lastDepth + o0[c]);
                                lst[c] = __ldg(a.lastImage + o0[c]);
                            }
#pragma unroll
                            for(int c = 0; c < kSlotChunk; c++)
                            {
                                if(((inb >> c) & 1u) && rgb_accept(a, td1[c], r_d0[c], lst[c]))
                                {
                                    valid_mask |= 1u << c;
                                    r_diff[c] = __fsub_rn(c_img[c], static_cast<float>(lst[c]));
                                    cnt0 += 1;
                                    cnt1 += (int)(r_diff[c] * r_diff[c]);
                                }
                            }
                            if(L.full_corres)   // the reference writes a DataTerm for every pixel (reduce.cu:838): only when a test taps it
                            {
#pragma unroll
                                for(int c = 0; c < kSlotChunk; c++)
                                {
                                    const int k = gtid + c * gthreads;
                                    if(c < nslots && k < plane)
                                    {
                                        Corres cc;
                                        const bool v = (valid_mask >> c) & 1u;
                                        const int i = k / g.cols;
                                        cc.zx = v ? (short)(r_zxy[c] & 0xffff) : 0;
                                        cc.zy = v ? (short)(r_zxy[c] >> 16) : 0;
                                        cc.ox = v ? (short)(k - i * g.cols) : 0;
                                        cc.oy = v ? (short)i : 0;
                                        cc.diff = v ? r_diff[c] : 0.f;
                                        cc.valid = v ? 1 : 0;
                                        reinterpret_cast<int4 *>(cimg)[k] = *reinterpret_cast<const int4 *>(&cc);
                                    }
                                }
                            }
                        }
                        else
                        {
                            // general case (several sequences share the GPU, or a large image): correspondences go through memory
                            for(int m = 0; m < nslots; m++)
                            {
                                const int k = gtid + m * gthreads;
                                if(k >= plane) break;
                                const int i = k / g.cols;
                                const int j0 = k - i * g.cols;
                                Corres c;
                                c.zx = c.zy = c.ox = c.oy = 0;
                                c.diff = 0.f;
                                c.valid = 0;
                                if(rgb_candidate(a, j0, i) && rgb_associate(a, j0, i, c))
                                {
                                    cnt0 += 1;
                                    cnt1 += (int)(c.diff * c.diff);
                                }
                                reinterpret_cast<int4 *>(cimg)[k] = *reinterpret_cast<const int4 *>(&c);
                            }
                        }
                    }
                    GN_STAMP(rec, 2);
                    // per-thread counts are tiny: carried through the float reduction exactly (block sums < 2^24)
                    acc[29] = (float)cnt0;
                    acc[30] = (float)(cnt1 & 0xfff);
                    acc[31] = (float)(cnt1 >> 12);
                    mid_word = cta_publish32(acc, sh, myrow, L.rgb);
                }

                if(L.rgb)
                {
                    if(L.rgb_only)
                    {
                        group_barrier(bar, target, G);
                        GN_STAMP(rec, 3);
                        // count / sigma of the whole image -> rgbError decides the early exit (every CTA, identically)
                        if(warp0)
                        {
                            fold_count_sigma(sh, rowsA, G);
                            if(threadIdx.x == 0) gn_sigma(sh, true, rec);
                        }
                        __syncthreads();
                    }
                    else
                    {
                        // the barrier itself sums the correspondence counts: sigmaVal = sqrt(count) needs nothing else
                        // (RGBDOdometryef.cpp:457-471; the squared-residual sum is a statistic, folded after phase B)
                        group_barrier_sum(bar, target, G, mid_word, sum_running, &sh.mid_sum);
                        GN_STAMP(rec, 3);
                        const int word = sh.mid_sum;
                        const int rgbSize = word & 0xffffff;
                        const int sel = (rgbSize != 0 && (word >> 24) == 0) ? 1 : rgbSize;
                        sigma_now = __fsqrt_rn((float)sel);
                        if(rec) rec->sigma_in = sigma_now;
                    }
                    GN_STAMP(rec, 4);
                    if(L.rgb_only && sh.stop)
                    {
                        step++;
                        break;   // rgbOnly && rgbError > lastRGBError, RGBDOdometryef.cpp:460-463
                    }

                    // ---------------- phase B: RGB Jacobian products
                    float acc[32];
#pragma unroll
                    for(int k = 0; k < 32; k++) acc[k] = 0.f;
                    RgbStepArgs a;
                    a.sigma = L.rgb_only ? sh.sigmaVal : sigma_now;
                    a.fx = g.fx; a.fy = g.fy;
                    a.sobelScale = L.sobel_scale;
                    a.cols = g.cols; a.rows = g.rows;
                    a.dIdx = P.dIdx; a.dIdy = P.dIdy;
                    a.lastDepth = P.lastDepth;
                    a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
                    a.cloud = nullptr;
                    if(single)
                    {
#pragma unroll
                        for(int c = 0; c < kSlotChunk; c++)
                            if((valid_mask >> c) & 1u)
                            {
                                float row[7];
                                rgb_row_regs(a, r_zxy[c] & 0xffff, r_zxy[c] >> 16, r_d0[c], c_gx[c], c_gy[c], r_diff[c], row);
                                float a29[29];
#pragma unroll
                                for(int q = 0; q < 29; q++) a29[q] = acc[q];
                                accumulate_se3(a29, row, true);
#pragma unroll
                                for(int q = 0; q < 29; q++) acc[q] = a29[q];
                            }
                    }
                    else
                    {
                        const Corres * cimg = P.corres;
                        for(int m = 0; m < nslots; m++)
                        {
                            const int k = gtid + m * gthreads;
                            if(k >= plane) break;
                            const int4 raw = *(reinterpret_cast<const int4 *>(cimg) + k);   // written by this very thread in phase A
                            const Corres c = *reinterpret_cast<const Corres *>(&raw);
                            if(c.valid & 0xff)
                            {
                                float row[7];
                                rgb_row(a, c, row);
                                float a29[29];
#pragma unroll
                                for(int q = 0; q < 29; q++) a29[q] = acc[q];
                                accumulate_se3(a29, row, true);
#pragma unroll
                                for(int q = 0; q < 29; q++) acc[q] = a29[q];
                            }
                        }
                    }
                    GN_STAMP(rec, 5);
                    cta_publish32(acc, sh, myrow + 32);
                }
                group_barrier(bar, target, G);
                fold_partials(sh, rowsA, G);
                step++;
                GN_STAMP(rec, 6);
                if(L.rgb && !L.rgb_only && threadIdx.x == 32)
                {
                    // what gn_sigma records at the mid-iteration point, from the folded integer columns
                    const int rgbSize = __float_as_int(sh.total[29]), sigma = __float_as_int(sh.total[30]);
                    sh.rgb_sigma_last = sigma;
                    sh.rgb_count_last = rgbSize;
                    sh.res.lastRGBCount = (float)rgbSize;
                }
                if(rec && L.rgb && !L.rgb_only)
                {
                    rec->rgb_count = __float_as_int(sh.total[29]);
                    rec->rgb_sigma = __float_as_int(sh.total[30]);
                }

                if(warp0)
                    warp_update(sh, L.icp, L.rgb, L.icp_weight, rec, t_start);
                else if(threadIdx.x < 64)
                    warp_stats(sh, L.icp, L.rgb, L.icp_weight);   // lastA / lastb / ICP error: off the solving warp
                GN_STAMP(rec, 7);
                if(rec) ntr++;
                __syncthreads();
            }
        }

        if(threadIdx.x == 0) seq_end(sh, L.rgb, L.rgb_only, leader ? &results[seq] : nullptr);
        if(rank == 0 && host_results && warp0)   // `leader` is thread 0 of the group's first CTA: all of its warp 0 copies
        {
            // The result block also goes straight to mapped host memory, followed by a per-sequence flag the host polls: the caller
            // has its pose ~1 us after the last solve instead of after kernel retirement + a D2H copy + a stream synchronisation.
            __syncwarp();
            const uint2 * src = reinterpret_cast<const uint2 *>(&sh.res);
            uint2 * dst = reinterpret_cast<uint2 *>(host_results + seq);
            for(int k = threadIdx.x; k < (int)(sizeof(GnResult) / sizeof(uint2)); k += 32) dst[k] = src[k];
            __threadfence_system();
            __syncwarp();
            if(threadIdx.x == 0) *reinterpret_cast<volatile unsigned *>(host_flags + seq) = host_seqno;
        }
        if(leader && L.trace) trace_count[seq] = ntr;
        __syncthreads();
    }
    if(leader)
    {
        ctl->base[group] = target;   // every CTA of the group ends with the same target
        ctl->sum_base[group] = sum_running;
    }
}

// ------------------------------------------------------------------ host side
size_t gn_state_bytes(int batch)
{
    size_t b = 0;
    b += (sizeof(GnCtl) + 255) / 256 * 256;
    b += (sizeof(GnSeqIn) * batch + 255) / 256 * 256;
    b += (size_t)kGnMaxCtas * 2 * kGnPartialStride * 4;
    b += (sizeof(GnResult) * batch + 255) / 256 * 256;
    b += (sizeof(slam_step_record) * kGnMaxTrace * batch + 255) / 256 * 256;
    b += ((size_t)4 * batch + 255) / 256 * 256;
    return b;
}

void gn_bind_state(GnDevice & d, char * base, int batch)
{
    d.batch = batch;
    char * p = base;
    d.ctl = (GnCtl *)p;
    p += (sizeof(GnCtl) + 255) / 256 * 256;
    d.seq_in = (GnSeqIn *)p;
    p += (sizeof(GnSeqIn) * batch + 255) / 256 * 256;
    d.partials = (float *)p;
    p += (size_t)kGnMaxCtas * 2 * kGnPartialStride * 4;
    d.results = (GnResult *)p;
    p += (sizeof(GnResult) * batch + 255) / 256 * 256;
    d.trace = (slam_step_record *)p;
    p += (sizeof(slam_step_record) * kGnMaxTrace * batch + 255) / 256 * 256;
    d.trace_count = (int *)p;
    d.stage_bytes = sizeof(GnSeqIn) * batch;
}

// Fold finished event pairs into kernel_ms / kernel_launches (synchronises on them).
int gn_fold_profile(GnDevice & d)
{
    for(size_t i = 0; i + 1 < d.ev.size(); i += 2)
    {
        SLAM_CUDA_TRY(cudaEventSynchronize(d.ev[i + 1]));
        float ms = 0.f;
        SLAM_CUDA_TRY(cudaEventElapsedTime(&ms, d.ev[i], d.ev[i + 1]));
        d.kernel_ms += ms;
        d.kernel_launches++;
        cudaEventDestroy(d.ev[i]);
        cudaEventDestroy(d.ev[i + 1]);
    }
    d.ev.clear();
    return SLAM_OK;
}

void gn_release(GnDevice & d)
{
    for(auto e : d.ev) cudaEventDestroy(e);
    d.ev.clear();
    if(d.h_stage) cudaFreeHost(d.h_stage);
    d.h_stage = nullptr;
}

// Fill the pinned staging image of the per-sequence input blocks (pointers + prior pose).
int gn_stage_inputs(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnSeqIn ** out)
{
    if(!d.h_stage)
    {
        SLAM_CUDA_TRY(cudaMallocHost((void **)&d.h_stage, d.stage_bytes));
        int dev = 0;
        SLAM_CUDA_TRY(cudaGetDevice(&dev));
        SLAM_CUDA_TRY(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
        int coop = 0;
        SLAM_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if(!coop)
        {
            set_last_error("device does not support cooperative launch");
            return SLAM_ERR_UNSUPPORTED;
        }
        if(d.num_sms > kGnMaxCtas) d.num_sms = kGnMaxCtas;
    }
    GnSeqIn * in = reinterpret_cast<GnSeqIn *>(d.h_stage);
    memset(in, 0, sizeof(GnSeqIn) * L.batch);
    for(int b = 0; b < L.batch; b++)
    {
        const SeqBuffers & s = seqs[b];
        for(int l = 0; l < L.levels; l++)
        {
            in[b].vcurr[l] = s.vcurr[l]; in[b].ncurr[l] = s.ncurr[l];
            in[b].vprev[l] = s.vprev[l]; in[b].nprev[l] = s.nprev[l];
            in[b].lastDepth[l] = s.lastDepth[l]; in[b].nextDepth[l] = s.nextDepth[l];
            in[b].lastImage[l] = s.lastImage[l]; in[b].nextImage[l] = s.nextImage[l];
            in[b].lastNextImage[l] = s.lastNextImage[l];
            in[b].dIdx[l] = s.dIdx[l]; in[b].dIdy[l] = s.dIdy[l];
            in[b].corres[l] = s.corres[l];
        }
        memcpy(in[b].Rprev, rot + 9 * b, 36);
        memcpy(in[b].tprev, trans + 3 * b, 12);
    }
    *out = in;
    return SLAM_OK;
}

int gn_enqueue(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnResult * h_results,
               cudaStream_t stream, unsigned * h_flags, unsigned seqno)
{
    GnSeqIn * in = nullptr;
    if(int rc = gn_stage_inputs(d, L, seqs, trans, rot, &in)) return rc;
    // one sequence: its pointer / pose block travels as a kernel parameter; several: one H2D copy of the array
    if(L.batch > 1) SLAM_CUDA_TRY(cudaMemcpyAsync(d.seq_in, d.h_stage, sizeof(GnSeqIn) * L.batch, cudaMemcpyHostToDevice, stream));

    // group geometry: every sequence gets its own group of G CTAs while they fit
    int G = gn_group_size(d.num_sms, L.batch);
    int groups = L.batch >= d.num_sms ? d.num_sms : L.batch;
    GnLaunch Lc = L;
    GnCtl * ctl = d.ctl;
    const GnSeqIn * seq_in = d.seq_in;
    float * partials = d.partials;
    GnResult * results = d.results;
    slam_step_record * trace = d.trace;
    int * trace_count = d.trace_count;
    GnSeqIn seq0 = in[0];
    // h_flags != nullptr: h_results / h_flags are mapped pinned memory the kernel writes itself (device view == host pointer under UVA)
    GnResult * host_results = h_flags ? h_results : nullptr;
    void * args[] = {&Lc, &ctl, &seq_in, &seq0, &partials, &results, &trace, &trace_count, &G, &groups, &host_results, &h_flags, &seqno};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(d.profiling)
    {
        if(d.ev.size() >= 4096)
            if(int rc = gn_fold_profile(d)) return rc;
        SLAM_CUDA_TRY(cudaEventCreate(&e0));
        SLAM_CUDA_TRY(cudaEventCreate(&e1));
        SLAM_CUDA_TRY(cudaEventRecord(e0, stream));
    }
    SLAM_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_gn_persistent, dim3(G * groups), dim3(kGnThreads), args, 0, stream));
    if(d.profiling)
    {
        SLAM_CUDA_TRY(cudaEventRecord(e1, stream));
        d.ev.push_back(e0);
        d.ev.push_back(e1);
    }
    if(!h_flags) SLAM_CUDA_TRY(cudaMemcpyAsync(h_results, d.results, sizeof(GnResult) * L.batch, cudaMemcpyDeviceToHost, stream));
    d.so3_swapped = L.so3;
    return SLAM_OK;
}

int gn_read_trace(GnDevice & d, int seq, slam_step_record * out, int max_records, int * n_records, cudaStream_t stream)
{
    int n = 0;
    SLAM_CUDA_TRY(cudaMemcpyAsync(&n, d.trace_count + seq, 4, cudaMemcpyDeviceToHost, stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(stream));
    *n_records = n;
    const int m = n < max_records ? n : max_records;
    if(m > 0 && out)
    {
        SLAM_CUDA_TRY(cudaMemcpyAsync(out, d.trace + (size_t)seq * kGnMaxTrace, sizeof(slam_step_record) * m, cudaMemcpyDeviceToHost, stream));
        SLAM_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return SLAM_OK;
}

}   // namespace slam
