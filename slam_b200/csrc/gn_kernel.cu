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

namespace slam {

constexpr int kIcpChunk = 3;    // ICP gathers in flight per thread (register budget)

struct GnShared
{
    // parameters of the running iteration (warp 0 writes, everyone reads after a sync)
    float Rcurr[9], tcurr[3], Rprev[9], tprev[3], Rprev_inv[9];
    float krk[9], kt[3];
    float so3H[9], so3Kinv[9], so3KR[9];
    float sigmaVal;
    int stop;
    // solver state (warp 0)
    double resultRt[16];
    double resultR[9], lastResultR[9];
    double K[9], Kinv[9];   // intrinsics of the running level (and of level 2 during SO3)
    double A[36], b[6], x[6], Rinc[9], newRt[12], Mi[9], KR[9], tinv[3];
    double aug[2][42];      // ping-pong buffers of the 6x7 Gauss-Jordan elimination
    int solve_ok;
    int rgb_sigma_last, rgb_count_last;   // operands of lastRGBError (computed once, at the end)
    float tinvf[3];
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

// Block sum of up to 32 per-thread floats (v[NV..31] must be 0) -> dst[0..31] (global partial row
// of this CTA).  Transposed warp reduction, then 32 threads add the per-warp rows.
// The caller must reach a __syncthreads() (e.g. group_barrier) before sh.red is reused.
// With count_cols, slots 29 / 30 / 31 carry exact small integers as floats (RGB correspondence count, low
// 12 bits and high bits of the squared-residual sum); they are recombined into the two int32 columns
// 29 (count) and 30 (sigma) of the partial row.
__device__ __forceinline__ void cta_publish32(float (&v)[32], GnShared & sh, float * dst, const bool count_cols = false)
{
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
        }
        dst[threadIdx.x] = total;
    }
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

__device__ __forceinline__ void k_matrix_d(const LevelGeom & g, double * K)
{
    for(int i = 0; i < 9; i++) K[i] = 0;
    K[0] = g.fx; K[4] = g.fy; K[2] = g.cx; K[5] = g.cy; K[8] = 1;
}

// =====================================================================================
// fp64 bookkeeping on warp 0.  Every stage reads its inputs from shared memory, writes its
// outputs to shared memory and ends with __syncwarp(); each lane evaluates exactly the
// expression small_math.hpp evaluates for that entry, so the host-stepped loop (which runs
// the serial routines) and this code agree bit for bit.
// =====================================================================================
__device__ __noinline__ void level_begin(GnShared & sh, const LevelGeom g)   // lane 0
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

// cofactor index table of smath::mat3_inverse: r[k] = det2(m[a], m[b], m[c], m[d]) / det
__constant__ int kCof[9][4] = {{4, 8, 5, 7}, {2, 7, 1, 8}, {1, 5, 2, 4}, {5, 6, 3, 8}, {0, 8, 2, 6}, {2, 3, 0, 5}, {3, 7, 4, 6}, {1, 6, 0, 7}, {0, 4, 1, 3}};

// From sh.resultRt: krk = float(K R K^-1), kt = float(K t) with [R|t] = resultRt^-1 (RGBDOdometryef.cpp:422-432),
// and the current pose Rcurr/tcurr = [Rprev|tprev] * float(resultRt)^-1 (:563-575).  Warp 0, all lanes.
__device__ __forceinline__ void warp_prepare(GnShared & sh, const bool with_pose)
{
    const int lane = threadIdx.x & 31;
    const double * M = sh.resultRt;   // row-major 4x4, affine
    // ---- stage 1: Mi = (3x3 part)^-1 (lanes 0..8); float isometry inverse pieces (lanes 12..23)
    if(lane < 9)
    {
        auto m = [&](int i) { return M[(i / 3) * 4 + (i % 3)]; };
        const double c00 = smath::det2(m(4), m(8), m(5), m(7));
        const double c01 = smath::det2(m(5), m(6), m(3), m(8));
        const double c02 = smath::det2(m(3), m(7), m(4), m(6));
        const double det = smath::dot3(m(0), c00, m(1), c01, m(2), c02);
        const double id = smath::dvd(1.0, det);
        const double cof = smath::det2(m(kCof[lane][0]), m(kCof[lane][1]), m(kCof[lane][2]), m(kCof[lane][3]));
        sh.Mi[lane] = smath::mul(cof, id);
    }
    else if(with_pose && lane >= 12 && lane < 15)
    {
        // tinv[i] = -(Rinv[i][:] . to), Rinv = Ro^T, Ro/to = float(resultRt)
        const int i = lane - 12;
        sh.tinvf[i] = -smath::dot3((float)M[0 * 4 + i], (float)M[3], (float)M[1 * 4 + i], (float)M[7], (float)M[2 * 4 + i], (float)M[11]);
    }
    else if(with_pose && lane >= 15 && lane < 24)
    {
        // Rcurr = Rprev * Rinv
        const int i = (lane - 15) / 3, j = (lane - 15) % 3;
        sh.Rcurr[i * 3 + j] = smath::dot3(sh.Rprev[i * 3 + 0], (float)M[j * 4 + 0], sh.Rprev[i * 3 + 1], (float)M[j * 4 + 1], sh.Rprev[i * 3 + 2], (float)M[j * 4 + 2]);
    }
    __syncwarp();
    // ---- stage 2: KR = K * Mi (lanes 0..8), tinv = -Mi * t (lanes 9..11), tcurr (lanes 12..14)
    if(lane < 9)
    {
        const int i = lane / 3, j = lane % 3;
        sh.KR[lane] = smath::dot3(sh.K[i * 3 + 0], sh.Mi[0 * 3 + j], sh.K[i * 3 + 1], sh.Mi[1 * 3 + j], sh.K[i * 3 + 2], sh.Mi[2 * 3 + j]);
    }
    else if(lane < 12)
    {
        const int i = lane - 9;
        sh.tinv[i] = -smath::dot3(sh.Mi[i * 3 + 0], M[3], sh.Mi[i * 3 + 1], M[7], sh.Mi[i * 3 + 2], M[11]);
    }
    else if(with_pose && lane < 15)
    {
        const int i = lane - 12;
        sh.tcurr[i] = smath::add(smath::dot3(sh.Rprev[i * 3 + 0], sh.tinvf[0], sh.Rprev[i * 3 + 1], sh.tinvf[1], sh.Rprev[i * 3 + 2], sh.tinvf[2]), sh.tprev[i]);
    }
    __syncwarp();
    // ---- stage 3: KRK = KR * Kinv (lanes 0..8), kt = K * tinv (lanes 9..11)
    if(lane < 9)
    {
        const int i = lane / 3, j = lane % 3;
        sh.krk[lane] = (float)smath::dot3(sh.KR[i * 3 + 0], sh.Kinv[0 * 3 + j], sh.KR[i * 3 + 1], sh.Kinv[1 * 3 + j], sh.KR[i * 3 + 2], sh.Kinv[2 * 3 + j]);
    }
    else if(lane < 12)
    {
        const int i = lane - 9;
        sh.kt[i] = (float)smath::dot3(sh.K[i * 3 + 0], sh.tinv[0], sh.K[i * 3 + 1], sh.tinv[1], sh.K[i * 3 + 2], sh.tinv[2]);
    }
    __syncwarp();
}

// smath::gauss_jordan_solve<double, 6> with the 42 entries of [A | b] spread over the lanes of warp 0
// (same per-entry arithmetic, bit-identical result).  sh.aug[0] holds the system on entry; x lands in sh.x.
__device__ __forceinline__ void warp_gauss_jordan(GnShared & sh)
{
    const int lane = threadIdx.x & 31;
    double dmax = 0;
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
        const double d = sh.aug[0][i * 7 + i];
        dmax = d > dmax ? d : dmax;
    }
    const double floor_d = smath::mul(dmax, 1e-9);
    bool ok = dmax > 0.0;
#pragma unroll
    for(int k = 0; k < 6; k++)
    {
        const double * src = sh.aug[k & 1];
        double * dst = sh.aug[(k & 1) ^ 1];
        const double p = src[k * 7 + k];
        ok = ok && (p > floor_d);
        const double inv = smath::dvd(1.0, p);
#pragma unroll
        for(int pass = 0; pass < 2; pass++)
        {
            const int e = lane + 32 * pass;
            if(e < 42)
            {
                const int i = e / 7, j = e - i * 7;
                double v = src[e];
                if(j > k)
                {
                    const double rkj = smath::mul(src[k * 7 + j], inv);
                    v = (i == k) ? rkj : smath::sub(v, smath::mul(src[i * 7 + k], rkj));
                }
                dst[e] = v;
            }
        }
        __syncwarp();
    }
    // six steps: the result is back in aug[0]
    if(lane < 6) sh.x[lane] = sh.aug[0][lane * 7 + 6];
    if(lane == 0) sh.solve_ok = ok ? 1 : 0;
    __syncwarp();
}

// Degenerate system (a pivot not safely positive): the pivoted / pseudo-inverse LDL^T of small_math.hpp.  Lane 0.
__device__ __noinline__ void solve_fallback(GnShared & sh)
{
    double A[36], b[6], x[6];
    for(int k = 0; k < 36; k++) A[k] = sh.A[k];
    for(int k = 0; k < 6; k++) b[k] = sh.b[k];
    smath::ldlt_solve_pivoted<double, 6>(A, b, x, DBL_EPSILON);
    for(int k = 0; k < 6; k++) sh.x[k] = x[k];
}

// Incremental rotation of the step (odom/utils.h:16-52).  Lane 0.
__device__ __noinline__ void rodrigues_core(GnShared & sh)
{
    double r[3] = {sh.x[3], sh.x[4], sh.x[5]}, R[9];
    smath::rodrigues(r, R);
    for(int k = 0; k < 9; k++) sh.Rinc[k] = R[k];
}

// RGBDOdometryef.cpp:509-575 on warp 0: combine the two systems, solve, update resultRt, then the next
// iteration's parameters.  icp sums = total[0..28], rgb sums = total[32..60].
__device__ __forceinline__ void warp_update(GnShared & sh, const bool icp, const bool rgb, const float icpWeight, slam_step_record * rec, const long long t_start)
{
    const int lane = threadIdx.x & 31;
#define GN_SSTAMP(idx) do { if(rec) rec->t_solve[idx] = (unsigned)(clock64() - t_start); } while(0)
    // ---- stage 0: lastA / lastb (upper triangle + mirror), stats
    if(lane < 27)
    {
        // lane -> (i, j) of the row-major upper triangle of the 6x7 augmented system (reduce.cu:475-486)
        int i = 0, rem = lane;
        while(rem >= 7 - i)
        {
            rem -= 7 - i;
            i++;
        }
        const int j = i + rem;
        const float vi = sh.total[lane];
        const float vr = sh.total[32 + lane];
        double v;
        if(icp && rgb)
        {
            const double w = icpWeight;
            v = (j == 6) ? smath::add((double)vr, smath::mul(w, (double)vi)) : smath::add((double)vr, smath::mul(smath::mul(w, w), (double)vi));
        }
        else
            v = icp ? (double)vi : (double)vr;
        if(j == 6)
        {
            sh.b[i] = v;
            sh.aug[0][i * 7 + 6] = v;
        }
        else
        {
            sh.A[i * 6 + j] = v;
            sh.A[j * 6 + i] = v;
            sh.aug[0][i * 7 + j] = v;
            sh.aug[0][j * 7 + i] = v;
        }
    }
    else if(lane == 27 && icp)
    {
        sh.res.lastICPError = __fdiv_rn(__fsqrt_rn(sh.total[27]), sh.total[28]);
        sh.res.lastICPCount = sh.total[28];
    }
    __syncwarp();
    GN_SSTAMP(0);
    // ---- stage 1: x = A^-1 b (parallel elimination), then the incremental rotation
    warp_gauss_jordan(sh);
    GN_SSTAMP(1);
    if(lane == 0)
    {
        if(!sh.solve_ok) solve_fallback(sh);
        rodrigues_core(sh);
    }
    __syncwarp();
    GN_SSTAMP(2);
    // ---- stage 2: resultRt <- [Rinc | x[0:3]; 0 0 0 1] * resultRt (odom/utils.h:54-68), rows 0..2
    if(lane < 12)
    {
        const int i = lane / 4, j = lane % 4;
        double s = smath::mul(sh.Rinc[i * 3 + 0], sh.resultRt[0 * 4 + j]);   // add(0, x) == x
        s = smath::add(s, smath::mul(sh.Rinc[i * 3 + 1], sh.resultRt[1 * 4 + j]));
        s = smath::add(s, smath::mul(sh.Rinc[i * 3 + 2], sh.resultRt[2 * 4 + j]));
        s = smath::add(s, smath::mul(sh.x[i], sh.resultRt[3 * 4 + j]));
        sh.newRt[lane] = s;
    }
    __syncwarp();
    if(lane < 12) sh.resultRt[lane] = sh.newRt[lane];
    if(lane >= 12 && lane < 18) sh.res.lastb[lane - 12] = sh.b[lane - 12];
    for(int k = lane; k < 36; k += 32) sh.res.lastA[k] = sh.A[k];
    __syncwarp();
    GN_SSTAMP(3);
    // ---- stages 3..5: parameters of the next iteration
    warp_prepare(sh, true);
    GN_SSTAMP(4);
    if(lane == 0)
    {
        sh.res.gn_iterations++;
        if(rec)
        {
            for(int k = 0; k < 29; k++)
            {
                rec->icp[k] = icp ? sh.total[k] : 0.f;
                rec->rgb[k] = rgb ? sh.total[32 + k] : 0.f;
            }
            for(int k = 0; k < 6; k++) rec->x[k] = sh.x[k];
            for(int k = 0; k < 9; k++) rec->Rcurr[k] = sh.Rcurr[k];
            for(int k = 0; k < 3; k++) rec->tcurr[k] = sh.tcurr[k];
        }
    }
}

__device__ __noinline__ void so3_prepare(GnShared & sh)   // lane 0
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

// RGBDOdometryef.cpp:346-378 (lane 0)
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
        float ru[9], rl[9];
        for(int k = 0; k < 9; k++)
        {
            ru[k] = (float)rotUpdate[k];
            rl[k] = sh.R_lr[k];
        }
        smath::mat3_mul(ru, rl, rl);
        for(int k = 0; k < 9; k++)
        {
            sh.R_lr[k] = rl[k];
            sh.resultR[k] = rl[k];
        }
        if(rec)
            for(int k = 0; k < 3; k++) rec->x[k] = delta[k];
    }
    if(rec)
        for(int k = 0; k < 9; k++) rec->Rcurr[k] = (float)sh.resultR[k];
    sh.stop = stop ? 1 : 0;
}

// RGBDOdometryef.cpp:457-471; count/sigma are in sh.total[29], [30] (integer bit patterns).  Lane 0.
// sigmaVal = sqrt(rgbSize) (or 1, or -1): the fp32 square root of an integer below 2^24 equals the reference's
// float(sqrt(double)) exactly.  rgbError is only a statistic unless rgbOnly (where it decides the early exit), so
// outside that mode its fp64 arithmetic is deferred to the end of the sequence.
__device__ __forceinline__ void gn_sigma(GnShared & sh, const bool rgb_only, slam_step_record * rec)
{
    const int rgbSize = __float_as_int(sh.total[29]);
    const int sigma = __float_as_int(sh.total[30]);
    // sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize): the quotient is 0 only for sigma == 0 with rgbSize != 0
    const int sel = (rgbSize != 0 && sigma == 0) ? 1 : rgbSize;
    float sigmaVal = __fsqrt_rn((float)sel);
    sh.stop = 0;
    if(rgb_only)
    {
        const float rgbError = (float)(sqrt((double)sigma) / (double)(rgbSize == 0 ? 1 : rgbSize));
        sh.stop = (rgbError > sh.res.lastRGBError) ? 1 : 0;
        if(!sh.stop) sh.res.lastRGBError = rgbError;
        sigmaVal = -1;
    }
    if(!sh.stop)
    {
        sh.rgb_sigma_last = sigma;
        sh.rgb_count_last = rgbSize;
        sh.res.lastRGBCount = (float)rgbSize;
    }
    sh.sigmaVal = sigmaVal;
    if(rec)
    {
        rec->sigma_in = sigmaVal;
        rec->rgb_count = rgbSize;
        rec->rgb_sigma = sigma;
    }
}

__device__ __noinline__ void seq_begin(GnShared & sh, const GnSeqIn & in)   // lane 0
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
    sh.rgb_sigma_last = 0;
    sh.rgb_count_last = -1;
}

__device__ __noinline__ void seq_end(GnShared & sh, const bool rgb, const bool rgb_only, GnResult * out)   // lane 0
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
    if(rgb && !rgb_only && sh.rgb_count_last >= 0)   // RGBDOdometryef.cpp:458
        sh.res.lastRGBError = (float)(sqrt((double)sh.rgb_sigma_last) / (double)(sh.rgb_count_last == 0 ? 1 : sh.rgb_count_last));
    if(out)
    {
        for(int k = 0; k < 9; k++) sh.res.Rcurr[k] = sh.Rcurr[k];
        for(int k = 0; k < 3; k++) sh.res.tcurr[k] = sh.tcurr[k];
        *out = sh.res;
    }
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kGnThreads, 1)
k_gn_persistent(const GnLaunch L, GnCtl * ctl, const GnSeqIn * seqs, const GnSeqIn seq0, float * partials, GnResult * results, slam_step_record * trace,
                int * trace_count, const int G, const int groups)
{
    __shared__ GnShared sh;

    const int group = blockIdx.x / G;
    const int rank = blockIdx.x - group * G;
    if(group >= groups) return;

    unsigned * bar = &ctl->barrier[group];
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
        const GnSeqIn & in = (L.batch == 1) ? seq0 : seqs[seq];
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
            const int nslots = (plane + gthreads - 1) / gthreads;
            // all of this thread's pixels fit one register-resident chunk (always true for one 640x480 sequence on a full GPU)
            const bool single = nslots <= kSlotChunk;
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
                a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                a.nextDepth = in.nextDepth[lvl];
                a.nextImage = in.nextImage[lvl];
                a.cols = g.cols; a.rows = g.rows;
#pragma unroll
                for(int c = 0; c < kSlotChunk; c++)
                {
                    const int k = gtid + c * gthreads;
                    if(c < nslots && k < plane)
                    {
                        const int i = k / g.cols;
                        const int j0 = k - i * g.cols;
                        short gx, gy;
                        bool is_cand;
                        if(L.derive_gradients)
                            is_cand = rgb_candidate_derive(a, j0, i, gx, gy);
                        else
                        {
                            is_cand = rgb_candidate(a, j0, i);
                            gx = is_cand ? a.dIdx[k] : (short)0;
                            gy = is_cand ? a.dIdy[k] : (short)0;
                        }
                        if(is_cand)
                        {
                            cand |= 1u << c;
                            c_d1[c] = a.nextDepth[k];
                            c_img[c] = static_cast<float>(a.nextImage[k]);
                            c_gx[c] = gx;
                            c_gy[c] = gy;
                        }
                    }
                }
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
                        a.vcurr = in.vcurr[lvl]; a.ncurr = in.ncurr[lvl]; a.vprev = in.vprev[lvl]; a.nprev = in.nprev[lvl];
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
                        a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                        a.lastDepth = in.lastDepth[lvl]; a.nextDepth = in.nextDepth[lvl];
                        a.lastImage = in.lastImage[lvl]; a.nextImage = in.nextImage[lvl];
                        a.maxDepthDelta = L.max_depth_delta;
                        a.kt = make_float3(sh.kt[0], sh.kt[1], sh.kt[2]);
                        a.krkinv = mat3_from(sh.krk);
                        a.cols = g.cols; a.rows = g.rows;
                        Corres * cimg = in.corres[lvl];
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
                    cta_publish32(acc, sh, myrow, L.rgb);
                }

                if(L.rgb)
                {
                    group_barrier(bar, target, G);
                    GN_STAMP(rec, 3);
                    // count / sigma of the whole image -> sigmaVal (every CTA, identically)
                    if(warp0)
                    {
                        fold_count_sigma(sh, rowsA, G);
                        if(threadIdx.x == 0) gn_sigma(sh, L.rgb_only, rec);
                    }
                    __syncthreads();
                    GN_STAMP(rec, 4);
                    if(sh.stop)
                    {
                        step++;
                        break;   // rgbOnly && rgbError > lastRGBError, RGBDOdometryef.cpp:460-463
                    }

                    // ---------------- phase B: RGB Jacobian products
                    float acc[32];
#pragma unroll
                    for(int k = 0; k < 32; k++) acc[k] = 0.f;
                    RgbStepArgs a;
                    a.sigma = sh.sigmaVal;
                    a.fx = g.fx; a.fy = g.fy;
                    a.sobelScale = L.sobel_scale;
                    a.cols = g.cols; a.rows = g.rows;
                    a.dIdx = in.dIdx[lvl]; a.dIdy = in.dIdy[lvl];
                    a.lastDepth = in.lastDepth[lvl];
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
                        const Corres * cimg = in.corres[lvl];
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

                if(warp0) warp_update(sh, L.icp, L.rgb, L.icp_weight, rec, t_start);
                GN_STAMP(rec, 7);
                if(rec) ntr++;
                __syncthreads();
            }
        }

        if(threadIdx.x == 0) seq_end(sh, L.rgb, L.rgb_only, leader ? &results[seq] : nullptr);
        if(leader && L.trace) trace_count[seq] = ntr;
        __syncthreads();
    }
    if(leader) ctl->base[group] = target;   // every CTA of the group ends with the same target
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
    void * args[] = {&Lc, &ctl, &seq_in, &seq0, &partials, &results, &trace, &trace_count, &G, &groups};
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
