// Shared-memory state and the fp64 bookkeeping routines of the Gauss-Newton loop, used by the persistent
// kernel (gn_kernel.cu, one CTA group per sequence) and by the batched streaming engine (batch_engine.cu,
// one warp per sequence between streaming map-reduce launches).
#pragma once
#include <cfloat>
#include <cstring>
#include "gn_kernel.cuh"

namespace slam {

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
    int stop_level, so3_done;   // batched streaming engine only: level whose iterations were cut short (rgbOnly), SO3 loop finished
    int rgb_sigma_last, rgb_count_last;   // operands of lastRGBError (computed once, at the end)
    int ntr;                              // batched streaming engine only: step records written so far
    float tinvf[3];
    float R_lr[9];
    float lastError, lastCount;
    GnResult res;
    // reduction scratch
    alignas(16) float red[32 * kGnPartialStride];
    float total[kGnPartialStride];
};

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
static __device__ __noinline__ void level_begin(GnShared & sh, const LevelGeom g)   // lane 0
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
static __constant__ int kCof[9][4] = {{4, 8, 5, 7}, {2, 7, 1, 8}, {1, 5, 2, 4}, {5, 6, 3, 8}, {0, 8, 2, 6}, {2, 3, 0, 5}, {3, 7, 4, 6}, {1, 6, 0, 7}, {0, 4, 1, 3}};

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
static __device__ __noinline__ void solve_fallback(GnShared & sh)
{
    double A[36], b[6], x[6];
    for(int k = 0; k < 36; k++) A[k] = sh.A[k];
    for(int k = 0; k < 6; k++) b[k] = sh.b[k];
    smath::ldlt_solve_pivoted<double, 6>(A, b, x, DBL_EPSILON);
    for(int k = 0; k < 6; k++) sh.x[k] = x[k];
}

// Incremental rotation of the step (odom/utils.h:16-52).  Lane 0.
static __device__ __noinline__ void rodrigues_core(GnShared & sh)
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

static __device__ __noinline__ void so3_prepare(GnShared & sh)   // lane 0
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
static __device__ __noinline__ void so3_update(GnShared & sh, int it, slam_step_record * rec)
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

static __device__ __noinline__ void seq_begin(GnShared & sh, const GnSeqIn & in)   // lane 0
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

static __device__ __noinline__ void seq_end(GnShared & sh, const bool rgb, const bool rgb_only, GnResult * out)   // lane 0
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


}   // namespace slam
