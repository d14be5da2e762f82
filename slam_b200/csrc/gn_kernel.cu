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
//   * a step is "map" (every thread accumulates its pixels' 29/11 products in registers)
//     + "publish" (warp shuffle -> shared memory -> one 64-float partial row per CTA in
//     global memory) + a group barrier (release/acquire counter) + "fold" (every CTA
//     re-reads the G partial rows in rank order, so all CTAs hold bit-identical sums);
//   * thread 0 of EVERY CTA then solves the normal equations redundantly in fp64
//     (small_math.hpp) and leaves the next iteration's parameters in shared memory;
//     no host round trip, no second launch, no broadcast step;
//   * partial rows are double-buffered by step parity, which makes one barrier per
//     reduction sufficient.
// Per-pixel arithmetic is pixel_ops.cuh, shared with the single-launch operator kernels.
#include <cfloat>
#include <cstring>
#include "gn_kernel.cuh"

namespace slam {

struct GnShared
{
    // parameters of the running iteration (thread 0 writes, everyone reads after a sync)
    float Rcurr[9], tcurr[3], Rprev[9], tprev[3], Rprev_inv[9];
    float krk[9], kt[3];
    float so3H[9], so3Kinv[9], so3KR[9];
    float sigmaVal;
    int stop;
    // solver state (thread 0)
    double resultRt[16];
    double resultR[9], lastResultR[9];
    double K[9], Kinv[9];   // intrinsics of the running level (and of level 2 during SO3)
    float R_lr[9];
    float lastError, lastCount;
    GnResult res;
    // reduction scratch
    alignas(16) float red[32 * kGnPartialStride];
    float total[kGnPartialStride];
};

__device__ __forceinline__ void group_barrier(unsigned * ctr, unsigned & target, unsigned G)
{
    __syncthreads();
    if(G > 1 && threadIdx.x == 0)
    {
        target += G;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
        unsigned v;
        do
        {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while((int)(v - target) < 0);
    }
    __syncthreads();
}

// Block sum of NV per-thread values -> dst[0..NV) (global partial row of this CTA).
template <typename T, int NV>
__device__ __forceinline__ void cta_publish(T (&acc)[NV], GnShared & sh, T * dst)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T * red = reinterpret_cast<T *>(sh.red);
#pragma unroll
    for(int k = 0; k < NV; k++)
    {
        const T s = warp_sum(acc[k]);
        if(lane == 0) red[wid * NV + k] = s;
    }
    __syncthreads();
    if(threadIdx.x < NV)
    {
        T total = 0;
        for(int w = 0; w < nw; w++) total += red[w * NV + threadIdx.x];
        dst[threadIdx.x] = total;
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

__device__ __forceinline__ void load4(const float * base, int p, int n, bool vec, float (&out)[4])
{
    if(vec)
    {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(base + p));
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    }
    else
    {
#pragma unroll
        for(int k = 0; k < 4; k++) out[k] = (p + k < n) ? __ldg(base + p + k) : SLAM_QNAN;
    }
}

__device__ __forceinline__ void k_matrix_d(const LevelGeom & g, double * K)
{
    for(int i = 0; i < 9; i++) K[i] = 0;
    K[0] = g.fx; K[4] = g.fy; K[2] = g.cx; K[5] = g.cy; K[8] = 1;
}

// ---- thread-0 scalar sections (kept out of line: they are fp64-heavy and must not
//      inflate the register allocation of the pixel loops) ---------------------------
__device__ __noinline__ void level_begin(GnShared & sh, const LevelGeom g)
{
    double K[9], Kinv[9];
    k_matrix_d(g, K);
    smath::mat3_inverse(K, Kinv);
    for(int k = 0; k < 9; k++)
    {
        sh.K[k] = K[k];
        sh.Kinv[k] = Kinv[k];
    }
}

__device__ __noinline__ void so3_prepare(GnShared & sh)
{
    double K[9], Kinv[9], R[9], KR[9], H[9];
    for(int k = 0; k < 9; k++)
    {
        K[k] = sh.K[k];
        Kinv[k] = sh.Kinv[k];
        R[k] = sh.resultR[k];
    }
    smath::mat3_mul(K, R, KR);
    smath::mat3_mul(KR, Kinv, H);
    for(int k = 0; k < 9; k++)
    {
        sh.so3H[k] = (float)H[k];
        sh.so3Kinv[k] = (float)Kinv[k];
        sh.so3KR[k] = (float)KR[k];
    }
}

// RGBDOdometryef.cpp:346-378
__device__ __noinline__ void so3_update(GnShared & sh, int it, slam_step_record * rec)
{
    const float * s = sh.total;
    float jtj[9], jtr[3];
    int shift = 0;
    for(int i = 0; i < 3; ++i)
        for(int j = i; j < 4; ++j)
        {
            const float value = s[shift++];
            if(j == 3)
                jtr[i] = value;
            else
                jtj[j * 3 + i] = jtj[i * 3 + j] = value;
        }
    const float residual0 = s[9], residual1 = s[10];
    sh.res.lastSO3Error = __fdiv_rn(__fsqrt_rn(residual0), residual1);
    sh.res.lastSO3Count = residual1;
    sh.res.so3_iterations++;

    if(rec)
    {
        rec->kind = 0;
        rec->level = 2;
        rec->iteration = it;
        for(int k = 0; k < 11; k++) rec->so3[k] = s[k];
        for(int k = 0; k < 9; k++)
        {
            rec->so3_in[k] = sh.so3H[k];
            rec->so3_in[9 + k] = sh.so3Kinv[k];
            rec->so3_in[18 + k] = sh.so3KR[k];
        }
    }

    bool stop = false;
    if(sh.res.lastSO3Error < sh.lastError && sh.lastCount == sh.res.lastSO3Count)
        stop = true;
    else if((double)sh.res.lastSO3Error > (double)sh.lastError + 0.001)
    {
        sh.res.lastSO3Error = sh.lastError;
        sh.res.lastSO3Count = sh.lastCount;
        for(int k = 0; k < 9; k++) sh.resultR[k] = sh.lastResultR[k];
        stop = true;
    }
    if(!stop)
    {
        sh.lastError = sh.res.lastSO3Error;
        sh.lastCount = sh.res.lastSO3Count;
        for(int k = 0; k < 9; k++) sh.lastResultR[k] = sh.resultR[k];
        float delta[3];
        smath::ldlt_solve<float, 3>(jtj, jtr, delta, FLT_EPSILON);
        const double dd[3] = {delta[0], delta[1], delta[2]};
        double rotUpdate[9];
        smath::rodrigues(dd, rotUpdate);
        float ru[9];
        for(int k = 0; k < 9; k++) ru[k] = (float)rotUpdate[k];
        smath::mat3_mul(ru, sh.R_lr, sh.R_lr);
        for(int k = 0; k < 9; k++) sh.resultR[k] = sh.R_lr[k];
        if(rec)
            for(int k = 0; k < 3; k++) rec->x[k] = delta[k];
    }
    if(rec)
        for(int k = 0; k < 9; k++) rec->Rcurr[k] = (float)sh.resultR[k];
    sh.stop = stop ? 1 : 0;
}

// RGBDOdometryef.cpp:422-432
__device__ __noinline__ void gn_prepare(GnShared & sh)
{
    double K[9], Kinv[9], M[16], Rt[16], R[9], KR[9], KRK[9];
    for(int k = 0; k < 9; k++)
    {
        K[k] = sh.K[k];
        Kinv[k] = sh.Kinv[k];
    }
    for(int k = 0; k < 16; k++) M[k] = sh.resultRt[k];
    smath::mat4_affine_inverse(M, Rt);
    for(int x = 0; x < 3; x++)
        for(int y = 0; y < 3; y++) R[x * 3 + y] = Rt[x * 4 + y];
    smath::mat3_mul(K, R, KR);
    smath::mat3_mul(KR, Kinv, KRK);
    for(int k = 0; k < 9; k++) sh.krk[k] = (float)KRK[k];
    for(int x = 0; x < 3; x++) sh.kt[x] = (float)smath::dot3(K[x * 3 + 0], Rt[3], K[x * 3 + 1], Rt[7], K[x * 3 + 2], Rt[11]);
}

// RGBDOdometryef.cpp:457-471; count/sigma are in sh.total[29], [30] (integer bit patterns)
__device__ __noinline__ void gn_sigma(GnShared & sh, const bool rgb_only, slam_step_record * rec)
{
    const int rgbSize = __float_as_int(sh.total[29]);
    const int sigma = __float_as_int(sh.total[30]);
    // sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize): the quotient is 0 only for sigma == 0 with rgbSize != 0
    const int sel = (rgbSize != 0 && sigma == 0) ? 1 : rgbSize;
    float sigmaVal = (float)sqrt((double)sel);
    const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
    sh.stop = (rgb_only && rgbError > sh.res.lastRGBError) ? 1 : 0;
    if(!sh.stop)
    {
        sh.res.lastRGBError = rgbError;
        sh.res.lastRGBCount = (float)rgbSize;
    }
    if(rgb_only) sigmaVal = -1;
    sh.sigmaVal = sigmaVal;
    if(rec)
    {
        rec->sigma_in = sigmaVal;
        rec->rgb_count = rgbSize;
        rec->rgb_sigma = sigma;
    }
}

// RGBDOdometryef.cpp:509-575: combine, solve, update pose.  icp sums = total[0..28],
// rgb sums = total[32..60].
__device__ __noinline__ void gn_update(GnShared & sh, const bool icp, const bool rgb, const float icpWeight, slam_step_record * rec)
{
    float A_icp[36], b_icp[6], A_rgb[36], b_rgb[6];
    for(int k = 0; k < 36; k++) A_icp[k] = A_rgb[k] = 0.f;
    for(int k = 0; k < 6; k++) b_icp[k] = b_rgb[k] = 0.f;
    int shift = 0;
    for(int i = 0; i < 6; ++i)
        for(int j = i; j < 7; ++j)
        {
            const float vi = sh.total[shift];
            const float vr = sh.total[32 + shift];
            shift++;
            if(j == 6)
            {
                b_icp[i] = vi;
                b_rgb[i] = vr;
            }
            else
            {
                A_icp[j * 6 + i] = A_icp[i * 6 + j] = vi;
                A_rgb[j * 6 + i] = A_rgb[i * 6 + j] = vr;
            }
        }
    if(icp)
    {
        sh.res.lastICPError = __fdiv_rn(__fsqrt_rn(sh.total[27]), sh.total[28]);
        sh.res.lastICPCount = sh.total[28];
    }
    double * A = sh.res.lastA;
    double * b = sh.res.lastb;
    if(icp && rgb)
    {
        const double w = icpWeight;
        const double ww = smath::mul(w, w);
        for(int k = 0; k < 36; k++) A[k] = smath::add((double)A_rgb[k], smath::mul(ww, (double)A_icp[k]));
        for(int k = 0; k < 6; k++) b[k] = smath::add((double)b_rgb[k], smath::mul(w, (double)b_icp[k]));
    }
    else if(icp)
    {
        for(int k = 0; k < 36; k++) A[k] = A_icp[k];
        for(int k = 0; k < 6; k++) b[k] = b_icp[k];
    }
    else
    {
        for(int k = 0; k < 36; k++) A[k] = A_rgb[k];
        for(int k = 0; k < 6; k++) b[k] = b_rgb[k];
    }
    double x[6];
    smath::ldlt_solve<double, 6>(A, b, x, DBL_EPSILON);
    smath::update_se3(sh.resultRt, x);
    smath::compose_current_pose(sh.Rprev, sh.tprev, sh.resultRt, sh.Rcurr, sh.tcurr);
    sh.res.gn_iterations++;
    if(rec)
    {
        for(int k = 0; k < 29; k++)
        {
            rec->icp[k] = icp ? sh.total[k] : 0.f;
            rec->rgb[k] = rgb ? sh.total[32 + k] : 0.f;
        }
        for(int k = 0; k < 6; k++) rec->x[k] = x[k];
        for(int k = 0; k < 9; k++) rec->Rcurr[k] = sh.Rcurr[k];
        for(int k = 0; k < 3; k++) rec->tcurr[k] = sh.tcurr[k];
    }
}

__device__ __noinline__ void seq_begin(GnShared & sh, const GnSeqIn & in)
{
    for(int k = 0; k < 9; k++) sh.Rprev[k] = sh.Rcurr[k] = in.Rprev[k];
    for(int k = 0; k < 3; k++) sh.tprev[k] = sh.tcurr[k] = in.tprev[k];
    smath::mat3_inverse(sh.Rprev, sh.Rprev_inv);
    for(int k = 0; k < 9; k++)
    {
        sh.resultR[k] = sh.lastResultR[k] = (k % 4 == 0) ? 1.0 : 0.0;
        sh.R_lr[k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    sh.lastError = FLT_MAX / 2;
    sh.lastCount = FLT_MAX / 2;
    memset(&sh.res, 0, sizeof(sh.res));
    sh.stop = 0;
}

__device__ __noinline__ void seq_end(GnShared & sh, const bool rgb, GnResult * out)
{
    if(rgb)
    {
        const float dx = smath::sub(sh.tcurr[0], sh.tprev[0]), dy = smath::sub(sh.tcurr[1], sh.tprev[1]), dz = smath::sub(sh.tcurr[2], sh.tprev[2]);
        const float n = __fsqrt_rn(smath::add(smath::add(smath::mul(dx, dx), smath::mul(dy, dy)), smath::mul(dz, dz)));
        if((double)n > 0.3)   // RGBDOdometryef.cpp:579-583
        {
            for(int k = 0; k < 9; k++) sh.Rcurr[k] = sh.Rprev[k];
            for(int k = 0; k < 3; k++) sh.tcurr[k] = sh.tprev[k];
        }
    }
    if(out)
    {
        for(int k = 0; k < 9; k++) sh.res.Rcurr[k] = sh.Rcurr[k];
        for(int k = 0; k < 3; k++) sh.res.tcurr[k] = sh.tcurr[k];
        *out = sh.res;
    }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kGnThreads, 1)
k_gn_persistent(const GnLaunch L, GnCtl * ctl, const GnSeqIn * seqs, float * partials, GnResult * results, slam_step_record * trace, int * trace_count,
                const int G, const int groups)
{
    __shared__ GnShared sh;

    const int group = blockIdx.x / G;
    const int rank = blockIdx.x - group * G;
    if(group >= groups) return;

    unsigned * bar = &ctl->barrier[group];
    unsigned target = 0;
    unsigned step = 0;
    // partial rows of this group: [parity][rank][64]
    float * gpart = partials + (size_t)group * 2 * G * kGnPartialStride;

    const long long t_start = clock64();
#define GN_STAMP(rec, idx) do { if(rec) (rec)->t_cycles[idx] = (unsigned)(clock64() - t_start); } while(0)
    const int gtid = rank * blockDim.x + threadIdx.x;
    const int gthreads = G * blockDim.x;
    const bool leader = (rank == 0 && threadIdx.x == 0);

    for(int seq = group; seq < L.batch; seq += groups)
    {
        const GnSeqIn & in = seqs[seq];
        slam_step_record * tr = (L.trace && leader) ? trace + (size_t)seq * kGnMaxTrace : nullptr;
        int ntr = 0;

        if(threadIdx.x == 0) seq_begin(sh, in);
        __syncthreads();

        // ------------------------------------------------ SO3 pre-alignment, level 2
        if(L.so3)
        {
            const LevelGeom g = L.geom[2];
            const int N = g.rows * g.cols;
            if(threadIdx.x == 0) level_begin(sh, g);
            for(int it = 0; it < 10; it++)
            {
                if(threadIdx.x == 0) so3_prepare(sh);
                __syncthreads();
                So3Args a;
                a.lastImage = in.lastNextImage[2];
                a.nextImage = in.nextImage[2];
                a.imageBasis = mat3_from(sh.so3H);
                a.kinv = mat3_from(sh.so3Kinv);
                a.krlr = mat3_from(sh.so3KR);
                a.cols = g.cols;
                a.rows = g.rows;

                float acc[11];
#pragma unroll
                for(int k = 0; k < 11; k++) acc[k] = 0.f;
                for(int k = gtid; k < N; k += gthreads)
                {
                    const int y = k / g.cols;
                    const int x = k - y * g.cols;
                    float row[4];
                    const bool found = so3_pixel(a, x, y, row);
                    accumulate_so3(acc, row, found);
                }
                float * myrow = gpart + ((step & 1) * G + rank) * kGnPartialStride;
                cta_publish<float, 11>(acc, sh, myrow);
                group_barrier(bar, target, G);
                fold_partials(sh, gpart + (step & 1) * G * kGnPartialStride, G);
                step++;

                if(threadIdx.x == 0)
                {
                    slam_step_record * rec = (tr && ntr < kGnMaxTrace) ? &tr[ntr] : nullptr;
                    if(rec) memset(rec, 0, sizeof(*rec));
                    so3_update(sh, it, rec);
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
            const bool vec = (plane & 3) == 0;
            const int nitems = (plane + 3) >> 2;
            if(threadIdx.x == 0)
            {
                sh.res.lastRGBError = FLT_MAX;
                level_begin(sh, g);
            }
            // Pose-independent half of the RGB association (border, 4x4 non-zero window, gradient magnitude,
            // finite depth; reduce.cu:780-807), evaluated once per level instead of once per iteration: bit m
            // of `cand` = this thread's m-th pixel (k = gtid + m * gthreads) is a candidate.
            const int nslots = (plane + gthreads - 1) / gthreads;
            const bool use_flags = L.rgb && nslots <= 32;
            unsigned cand = 0;
            if(use_flags)
            {
                ResidualArgs a;
                a.minScale = L.min_scale[lvl];
                a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                a.nextDepth = in.nextDepth[lvl];
                a.nextImage = in.nextImage[lvl];
                a.cols = g.cols; a.rows = g.rows;
                for(int m = 0; m < nslots; m++)
                {
                    const int k = gtid + m * gthreads;
                    if(k < plane)
                    {
                        const int i = k / g.cols;
                        if(rgb_candidate(a, k - i * g.cols, i)) cand |= 1u << m;
                    }
                }
            }

            for(int j = 0; j < L.iterations[lvl]; j++)
            {
                slam_step_record * rec = nullptr;
                const long long t_iter = clock64();
                if(threadIdx.x == 0)
                {
                    gn_prepare(sh);
                    if(tr && ntr < kGnMaxTrace)
                    {
                        rec = &tr[ntr];
                        memset(rec, 0, sizeof(*rec));
                        rec->kind = 1;
                        rec->level = lvl;
                        rec->iteration = j;
                        rec->t_cycles[0] = (unsigned)(t_iter - t_start);
                        GN_STAMP(rec, 1);
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
                }
                __syncthreads();

                float * rowsA = gpart + (step & 1) * G * kGnPartialStride;
                float * myrow = rowsA + rank * kGnPartialStride;

                unsigned valid_mask = 0;   // bit m: this thread's m-th pixel has an RGB correspondence this iteration
                // ---------------- phase A: ICP products + RGB association
                {
                    float acc[29];
#pragma unroll
                    for(int k = 0; k < 29; k++) acc[k] = 0.f;
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
                        a.vcurr = in.vcurr[lvl]; a.ncurr = in.ncurr[lvl]; a.vprev = in.vprev[lvl]; a.nprev = in.nprev[lvl];
                        for(int item = gtid; item < nitems; item += gthreads)
                        {
                            const int p = item << 2;
                            float vx[4], vy[4], vz[4], nx[4], ny[4], nz[4];
                            load4(a.vcurr, p, plane, vec, vx);
                            load4(a.vcurr + plane, p, plane, vec, vy);
                            load4(a.vcurr + 2 * plane, p, plane, vec, vz);
                            load4(a.ncurr, p, plane, vec, nx);
                            load4(a.ncurr + plane, p, plane, vec, ny);
                            load4(a.ncurr + 2 * plane, p, plane, vec, nz);
#pragma unroll
                            for(int k = 0; k < 4; k++)
                                if(p + k < plane)
                                {
                                    float row[7];
                                    const bool found = icp_pixel(a, make_float3(vx[k], vy[k], vz[k]), make_float3(nx[k], ny[k], nz[k]), row);
                                    accumulate_se3(acc, row, found);
                                }
                        }
                    }
                    int cnt[2] = {0, 0};
                    if(L.rgb)
                    {
                        ResidualArgs a;
                        a.minScale = L.min_scale[lvl];
                        a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                        a.lastDepth = in.lastDepth[lvl]; a.nextDepth = in.nextDepth[lvl];
                        a.lastImage = in.lastImage[lvl]; a.nextImage = in.nextImage[lvl];
                        a.maxDepthDelta = L.max_depth_delta;
                        a.kt = make_float3(sh.kt[0], sh.kt[1], sh.kt[2]);
                        a.krkinv = mat3_from(sh.krk);
                        a.cols = g.cols; a.rows = g.rows;
                        Corres * cimg = in.corres[lvl];
                        valid_mask = 0;
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
                            const bool is_cand = use_flags ? ((cand >> m) & 1u) != 0 : rgb_candidate(a, j0, i);
                            const bool ok = is_cand && rgb_associate(a, j0, i, c);
                            if(ok)
                            {
                                cnt[0] += 1;
                                cnt[1] += (int)(c.diff * c.diff);
                                if(m < 32) valid_mask |= 1u << m;
                            }
                            // the reference writes a DataTerm for every pixel; only the valid ones are ever read again,
                            // so the others are written only when a test wants to tap the whole image
                            if(ok || L.full_corres || !use_flags) reinterpret_cast<int4 *>(cimg)[k] = *reinterpret_cast<const int4 *>(&c);
                        }
                    }
                    GN_STAMP(rec, 2);
                    cta_publish<float, 29>(acc, sh, myrow);
                    if(L.rgb) cta_publish<int, 2>(cnt, sh, reinterpret_cast<int *>(myrow + 29));
                }

                if(L.rgb)
                {
                    group_barrier(bar, target, G);
                    GN_STAMP(rec, 3);
                    // count / sigma of the whole image -> sigmaVal (every CTA, identically)
                    if(threadIdx.x < 32)
                    {
                        int c0 = 0, c1 = 0;
                        for(int r = threadIdx.x; r < G; r += 32)
                        {
                            c0 += __float_as_int(__ldcg(rowsA + r * kGnPartialStride + 29));
                            c1 += __float_as_int(__ldcg(rowsA + r * kGnPartialStride + 30));
                        }
                        c0 = warp_sum(c0);
                        c1 = warp_sum(c1);
                        if(threadIdx.x == 0)
                        {
                            sh.total[29] = __int_as_float(c0);
                            sh.total[30] = __int_as_float(c1);
                            gn_sigma(sh, L.rgb_only, rec);
                        }
                    }
                    __syncthreads();
                    GN_STAMP(rec, 4);
                    if(sh.stop)
                    {
                        step++;
                        break;   // rgbOnly && rgbError > lastRGBError, RGBDOdometryef.cpp:460-463
                    }

                    // ---------------- phase B: RGB Jacobian products
                    float acc[29];
#pragma unroll
                    for(int k = 0; k < 29; k++) acc[k] = 0.f;
                    RgbStepArgs a;
                    a.sigma = sh.sigmaVal;
                    a.fx = g.fx; a.fy = g.fy;
                    a.sobelScale = L.sobel_scale;
                    a.cols = g.cols; a.rows = g.rows;
                    a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                    a.lastDepth = in.lastDepth[lvl];
                    a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
                    a.cloud = nullptr;
                    const Corres * cimg = in.corres[lvl];
                    for(int m = 0; m < nslots; m++)
                    {
                        const int k = gtid + m * gthreads;
                        if(k >= plane) break;
                        if(use_flags && !((valid_mask >> m) & 1u)) continue;
                        const int4 raw = *(reinterpret_cast<const int4 *>(cimg) + k);   // written by this very thread in phase A
                        const Corres c = *reinterpret_cast<const Corres *>(&raw);
                        if(c.valid & 0xff)
                        {
                            float row[7];
                            rgb_row(a, c, row);
                            accumulate_se3(acc, row, true);
                        }
                    }
                    GN_STAMP(rec, 5);
                    cta_publish<float, 29>(acc, sh, myrow + 32);
                }
                group_barrier(bar, target, G);
                fold_partials(sh, rowsA, G);
                step++;
                GN_STAMP(rec, 6);

                if(threadIdx.x == 0) gn_update(sh, L.icp, L.rgb, L.icp_weight, rec);
                GN_STAMP(rec, 7);
                if(rec) ntr++;
                __syncthreads();
            }
        }

        if(threadIdx.x == 0) seq_end(sh, L.rgb, leader ? &results[seq] : nullptr);
        if(leader && L.trace) trace_count[seq] = ntr;
        __syncthreads();
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
    d.seq_in = (GnSeqIn *)p;   // must directly follow ctl (one staging copy covers both)
    p += (sizeof(GnSeqIn) * batch + 255) / 256 * 256;
    d.partials = (float *)p;
    p += (size_t)kGnMaxCtas * 2 * kGnPartialStride * 4;
    d.results = (GnResult *)p;
    p += (sizeof(GnResult) * batch + 255) / 256 * 256;
    d.trace = (slam_step_record *)p;
    p += (sizeof(slam_step_record) * kGnMaxTrace * batch + 255) / 256 * 256;
    d.trace_count = (int *)p;
    d.stage_bytes = (sizeof(GnCtl) + 255) / 256 * 256 + sizeof(GnSeqIn) * batch;
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

int gn_enqueue(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnResult * h_results,
               cudaStream_t stream)
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
    memset(d.h_stage, 0, d.stage_bytes);
    GnSeqIn * in = reinterpret_cast<GnSeqIn *>(d.h_stage + (sizeof(GnCtl) + 255) / 256 * 256);
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
    SLAM_CUDA_TRY(cudaMemcpyAsync(d.ctl, d.h_stage, d.stage_bytes, cudaMemcpyHostToDevice, stream));

    // group geometry: every sequence gets its own group of G CTAs while they fit
    int G, groups;
    if(L.batch >= d.num_sms)
    {
        G = 1;
        groups = d.num_sms;
    }
    else
    {
        G = d.num_sms / L.batch;
        groups = L.batch;
    }
    GnLaunch Lc = L;
    GnCtl * ctl = d.ctl;
    const GnSeqIn * seq_in = d.seq_in;
    float * partials = d.partials;
    GnResult * results = d.results;
    slam_step_record * trace = d.trace;
    int * trace_count = d.trace_count;
    void * args[] = {&Lc, &ctl, &seq_in, &partials, &results, &trace, &trace_count, &G, &groups};
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
    SLAM_CUDA_TRY(cudaMemcpyAsync(h_results, d.results, sizeof(GnResult) * L.batch, cudaMemcpyDeviceToHost, stream));
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
