// Shared-memory state and the fp64 bookkeeping routines of the Gauss-Newton loop, used by the persistent
// kernel (gn_kernel.cu, one CTA group per sequence) and by the batched streaming engine (batch_engine.cu,
// one warp per sequence between streaming map-reduce launches).
#pragma once
#include <cfloat>
#include <cstring>
#include "gn_kernel.cuh"

namespace slam {

constexpr int kGnPartialStride = 64;   // floats of the folded sums: [0..31] ICP / SO3, [32..63] RGB

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
    double Ab[28];          // combined system of the running step (27 entries, upper-triangle order of the 6x7 augmented system)
    // lane-parallel step update (gn_fast_math.cuh): staging areas between its stages
    double Wsys[48];        // the 6x7 combined system, row stride 8; later the eliminated columns
    double Wm[24];          // [0..11] affine rows of the new resultRt, [12..20] inverse of its linear part
    double Wk[12];          // [0..2] -M^-1 t, [3..11] K M^-1
    float Wf[4];            // float(resultRt)^-1 translation
    int solve_ok;
    int stop_level, so3_done;   // batched streaming engine only: level whose iterations were cut short (rgbOnly), SO3 loop finished
    int rgb_sigma_last, rgb_count_last;   // operands of lastRGBError (computed once, at the end)
    int ntr;                              // batched streaming engine only: step records written so far
    int mid_sum;                          // persistent kernel only: result of the mid-iteration barrier sum
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

// =====================================================================================
// Register-resident bookkeeping.  ALL lanes of warp 0 evaluate the same expressions on the same values
// (one instruction stream, no divergence, no shared-memory stage between the steps of the dependent chain):
// a warp-wide fp64 instruction costs the same issue slots whether one lane or 32 are active, and the chain
// solve -> Rodrigues -> resultRt -> inverse -> K R K^-1 is what bounds an iteration of a single sequence
// (ncu: the other 15 warps wait at the block barrier for it).  Every entry is evaluated with exactly the
// expression small_math.hpp uses for it, so the host-stepped loop and this code agree bit for bit.
// =====================================================================================

// From the affine rows M (3x4, row-major) of resultRt: krk = float(K R K^-1), kt = float(K t) with
// [R|t] = resultRt^-1 (RGBDOdometryef.cpp:422-432) and, with_pose, the current pose
// Rcurr/tcurr = [Rprev|tprev] * float(resultRt)^-1 (:563-575).  Warp 0, all lanes; lane 0 stores.
__device__ __forceinline__ void prepare_from_rows(GnShared & sh, const double (&M)[12], const bool with_pose)
{
    const int lane = threadIdx.x & 31;
    double K[9], Kinv[9];
#pragma unroll
    for(int k = 0; k < 9; k++)
    {
        K[k] = sh.K[k];
        Kinv[k] = sh.Kinv[k];
    }
    const double R3[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    double Mi[9], KR[9], KRK[9], tinv[3];
    smath::mat3_inverse(R3, Mi);
    smath::mat3_mul(K, Mi, KR);
#pragma unroll
    for(int i = 0; i < 3; i++) tinv[i] = -smath::dot3(Mi[i * 3 + 0], M[3], Mi[i * 3 + 1], M[7], Mi[i * 3 + 2], M[11]);
    smath::mat3_mul(KR, Kinv, KRK);
    float kt[3];
#pragma unroll
    for(int i = 0; i < 3; i++) kt[i] = (float)smath::dot3(K[i * 3 + 0], tinv[0], K[i * 3 + 1], tinv[1], K[i * 3 + 2], tinv[2]);
    if(lane == 0)
    {
#pragma unroll
        for(int k = 0; k < 9; k++) sh.krk[k] = (float)KRK[k];
#pragma unroll
        for(int k = 0; k < 3; k++) sh.kt[k] = kt[k];
    }
    if(with_pose)
    {
        float Rp[9], tp[3], Mf[12], tinvf[3];
#pragma unroll
        for(int k = 0; k < 9; k++) Rp[k] = sh.Rprev[k];
#pragma unroll
        for(int k = 0; k < 3; k++) tp[k] = sh.tprev[k];
#pragma unroll
        for(int k = 0; k < 12; k++) Mf[k] = (float)M[k];
        // tinv[i] = -(Rinv[i][:] . to), Rinv = Ro^T, Ro/to = float(resultRt)
#pragma unroll
        for(int i = 0; i < 3; i++) tinvf[i] = -smath::dot3(Mf[0 * 4 + i], Mf[3], Mf[1 * 4 + i], Mf[7], Mf[2 * 4 + i], Mf[11]);
        float Rc[9], tc[3];
        // Rcurr = Rprev * Rinv, tcurr = Rprev * tinv + tprev
#pragma unroll
        for(int i = 0; i < 3; i++)
        {
#pragma unroll
            for(int j = 0; j < 3; j++) Rc[i * 3 + j] = smath::dot3(Rp[i * 3 + 0], Mf[j * 4 + 0], Rp[i * 3 + 1], Mf[j * 4 + 1], Rp[i * 3 + 2], Mf[j * 4 + 2]);
            tc[i] = smath::add(smath::dot3(Rp[i * 3 + 0], tinvf[0], Rp[i * 3 + 1], tinvf[1], Rp[i * 3 + 2], tinvf[2]), tp[i]);
        }
        if(lane == 0)
        {
#pragma unroll
            for(int k = 0; k < 9; k++) sh.Rcurr[k] = Rc[k];
#pragma unroll
            for(int k = 0; k < 3; k++) sh.tcurr[k] = tc[k];
        }
    }
    __syncwarp();
}

// The parameters of the first iteration of a level (Rcurr/tcurr carry over): resultRt comes from shared memory.
__device__ __forceinline__ void warp_prepare(GnShared & sh, const bool with_pose)
{
    double M[12];
#pragma unroll
    for(int k = 0; k < 12; k++) M[k] = sh.resultRt[k];
    prepare_from_rows(sh, M, with_pose);
}

// Degenerate system (a pivot not safely positive): the pivoted / pseudo-inverse LDL^T of small_math.hpp.  Lane 0.
static __device__ __noinline__ void solve_fallback(GnShared & sh, const bool icp, const bool rgb, const float icpWeight)
{
    double A[36], b[6], x[6];
    int shift = 0;
    for(int i = 0; i < 6; i++)
        for(int j = i; j < 7; j++)
        {
            const float vi = sh.total[shift], vr = sh.total[32 + shift];
            shift++;
            double v;
            if(icp && rgb)
            {
                const double w = icpWeight;
                v = (j == 6) ? smath::add((double)vr, smath::mul(w, (double)vi)) : smath::add((double)vr, smath::mul(smath::mul(w, w), (double)vi));
            }
            else
                v = icp ? (double)vi : (double)vr;
            if(j == 6)
                b[i] = v;
            else
                A[i * 6 + j] = A[j * 6 + i] = v;
        }
    smath::ldlt_solve_pivoted<double, 6>(A, b, x, DBL_EPSILON);
    for(int k = 0; k < 6; k++) sh.x[k] = x[k];
}

// lastA / lastb / lastICPError / lastICPCount of the step (RGBDOdometryef.cpp:509-556) from the folded sums: same
// expressions as the solver's own combination below, evaluated by ANOTHER warp while warp 0 solves (nothing on the
// dependent chain reads them; the degenerate-pivot fallback does, after a block barrier in its caller's order --
// see solve_fallback_sync).  One full warp.
__device__ __forceinline__ void warp_stats(GnShared & sh, const bool icp, const bool rgb, const float icpWeight)
{
    const int lane = threadIdx.x & 31;
    if(lane < 27)
    {
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
            sh.res.lastb[i] = v;
        else
        {
            sh.res.lastA[i * 6 + j] = v;
            sh.res.lastA[j * 6 + i] = v;
        }
    }
    else if(lane == 27 && icp)
    {
        sh.res.lastICPError = __fdiv_rn(__fsqrt_rn(sh.total[27]), sh.total[28]);
        sh.res.lastICPCount = sh.total[28];
    }
}

// RGBDOdometryef.cpp:509-575 on warp 0: combine the two systems, solve, update resultRt, then the next
// iteration's parameters.  icp sums = total[0..28], rgb sums = total[32..60].  The caller runs warp_stats() as well
// (on another warp, or before this call).
//
// smath::gauss_jordan_solve<double, 6> with one COLUMN of [A | b] per lane (lane j < 7 holds column j, the other
// lanes shadow column 6): step k broadcasts column k with shuffles, every lane forms 1 / pivot itself and updates
// its own column -- the per-entry arithmetic of the serial routine, hence the same bits.
__device__ __forceinline__ void warp_update(GnShared & sh, const bool icp, const bool rgb, const float icpWeight, slam_step_record * rec, const long long t_start)
{
    const int lane = threadIdx.x & 31;
    const int j = lane < 7 ? lane : 6;
#define GN_SSTAMP(idx) do { if(rec) rec->t_solve[idx] = (unsigned)(clock64() - t_start); } while(0)
    // the affine rows of resultRt (needed after the solve: fetched now, off the dependent chain)
    double Rt[12];
#pragma unroll
    for(int k = 0; k < 12; k++) Rt[k] = sh.resultRt[k];
    // ---- column j of lastA | lastb: entry (i, j) is element (min, max) of the row-major upper triangle of the
    //      6x7 augmented system (reduce.cu:475-486)
    double c[6];
    {
        const double w = icpWeight;
        const double ww = (j == 6) ? w : smath::mul(w, w);
#pragma unroll
        for(int i = 0; i < 6; i++)
        {
            const int a = i < j ? i : j, b = i < j ? j : i;
            const int idx = 7 * a - (a * (a - 1)) / 2 + (b - a);
            const float vi = sh.total[idx];
            const float vr = sh.total[32 + idx];
            c[i] = (icp && rgb) ? smath::add((double)vr, smath::mul(ww, (double)vi)) : (icp ? (double)vi : (double)vr);
        }
    }
    GN_SSTAMP(0);
    // ---- x = A^-1 b
    double dmax = 0;
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
        const double d = __shfl_sync(0xffffffffu, c[i], i);
        dmax = d > dmax ? d : dmax;
    }
    const double floor_d = smath::mul(dmax, 1e-9);
    bool ok = dmax > 0.0;
#pragma unroll
    for(int k = 0; k < 6; k++)
    {
        double colk[6];
#pragma unroll
        for(int i = 0; i < 6; i++) colk[i] = __shfl_sync(0xffffffffu, c[i], k);
        const double p = colk[k];
        ok = ok && (p > floor_d);
        const double inv = smath::dvd(1.0, p);
        const double rkj = smath::mul(c[k], inv);
        const bool upd = j > k;
#pragma unroll
        for(int i = 0; i < 6; i++)
        {
            const double v = (i == k) ? rkj : smath::sub(c[i], smath::mul(colk[i], rkj));
            c[i] = upd ? v : c[i];
        }
    }
    double x[6];
#pragma unroll
    for(int i = 0; i < 6; i++) x[i] = __shfl_sync(0xffffffffu, c[i], 6);
    GN_SSTAMP(1);
    if(!ok)   // uniform: every lane saw the same pivots
    {
        __syncwarp();
        if(lane == 0) solve_fallback(sh, icp, rgb, icpWeight);
        __syncwarp();
#pragma unroll
        for(int i = 0; i < 6; i++) x[i] = sh.x[i];
    }
    // ---- incremental rotation of the step (odom/utils.h:16-52)
    double Rinc[9];
    smath::rodrigues(x + 3, Rinc);
    GN_SSTAMP(2);
    // ---- resultRt <- [Rinc | x[0:3]; 0 0 0 1] * resultRt (odom/utils.h:54-68), rows 0..2 (row 3 stays 0 0 0 1)
    double M[12];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
        for(int q = 0; q < 4; q++)
        {
            double s = smath::mul(Rinc[i * 3 + 0], Rt[0 * 4 + q]);   // add(0, x) == x
            s = smath::add(s, smath::mul(Rinc[i * 3 + 1], Rt[1 * 4 + q]));
            s = smath::add(s, smath::mul(Rinc[i * 3 + 2], Rt[2 * 4 + q]));
            s = smath::add(s, smath::mul(x[i], q == 3 ? 1.0 : 0.0));
            M[i * 4 + q] = s;
        }
    if(lane < 12)
    {
        // lane-indexed store of a register array: selected with predicated moves, no local memory
        double v = M[0];
#pragma unroll
        for(int k = 1; k < 12; k++) v = (lane == k) ? M[k] : v;
        sh.resultRt[lane] = v;
    }
    else if(lane < 18)
    {
        double v = x[0];
#pragma unroll
        for(int k = 1; k < 6; k++) v = (lane - 12 == k) ? x[k] : v;
        sh.x[lane - 12] = v;
    }
    GN_SSTAMP(3);
    // ---- parameters of the next iteration
    prepare_from_rows(sh, M, true);
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

// H = K R K^-1, K^-1, K R of the next SO3 iteration (RGBDOdometryef.cpp:313-323) from resultR, all in registers.
__device__ __forceinline__ void so3_params(GnShared & sh, const double (&R)[9], const bool store)
{
    double K[9], Kinv[9], KR[9], H[9];
#pragma unroll
    for(int k = 0; k < 9; k++)
    {
        K[k] = sh.K[k];
        Kinv[k] = sh.Kinv[k];
    }
    smath::mat3_mul(K, R, KR);
    smath::mat3_mul(KR, Kinv, H);
    if(store)
    {
#pragma unroll
        for(int k = 0; k < 9; k++)
        {
            sh.so3H[k] = (float)H[k];
            sh.so3Kinv[k] = (float)Kinv[k];
            sh.so3KR[k] = (float)KR[k];
        }
    }
}

static __device__ __noinline__ void so3_prepare(GnShared & sh)   // lane 0
{
    double R[9];
    for(int k = 0; k < 9; k++) R[k] = sh.resultR[k];
    so3_params(sh, R, true);
}

static __device__ __noinline__ void so3_solve_fallback(GnShared & sh, const float * jtj, const float * jtr)   // lane 0
{
    float A[9], b[3], x[3];
    for(int k = 0; k < 9; k++) A[k] = jtj[k];
    for(int k = 0; k < 3; k++) b[k] = jtr[k];
    smath::ldlt_solve_pivoted<float, 3>(A, b, x, FLT_EPSILON);
    for(int k = 0; k < 3; k++) sh.x[k] = x[k];
}

// RGBDOdometryef.cpp:346-378 followed by the parameters of the next SO3 iteration (:313-323).  One full warp: every
// lane evaluates the same expressions in registers (no local arrays, no shared-memory stages), lane 0 stores.
__device__ __forceinline__ void warp_so3_update(GnShared & sh, int it, slam_step_record * rec)
{
    const int lane = threadIdx.x & 31;
    float s[11];
#pragma unroll
    for(int k = 0; k < 11; k++) s[k] = sh.total[k];
    float jtj[9], jtr[3];
    {
        int shift = 0;
#pragma unroll
        for(int i = 0; i < 3; ++i)
#pragma unroll
            for(int j = i; j < 4; ++j)
            {
                const float value = s[shift++];
                if(j == 3)
                    jtr[i] = value;
                else
                    jtj[j * 3 + i] = jtj[i * 3 + j] = value;
            }
    }
    float so3Error = __fdiv_rn(__fsqrt_rn(s[9]), s[10]);
    float so3Count = s[10];
    const float lastError = sh.lastError, lastCount = sh.lastCount;
    double R[9];
#pragma unroll
    for(int k = 0; k < 9; k++) R[k] = sh.resultR[k];

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

    bool stop = false, restore = false;
    if(so3Error < lastError && lastCount == so3Count)
        stop = true;
    else if((double)so3Error > (double)lastError + 0.001)
    {
        so3Error = lastError;
        so3Count = lastCount;
        restore = true;
        stop = true;
    }
    __syncwarp();
    if(restore)
    {
        if(lane < 9) sh.resultR[lane] = sh.lastResultR[lane];
    }
    if(!stop)   // uniform
    {
        float delta[3];
        const bool ok = smath::ldlt_solve_nopivot<float, 3>(jtj, jtr, delta);
        if(!ok)
        {
            if(lane == 0) so3_solve_fallback(sh, jtj, jtr);
            __syncwarp();
#pragma unroll
            for(int k = 0; k < 3; k++) delta[k] = (float)sh.x[k];
            __syncwarp();
        }
        const double dd[3] = {delta[0], delta[1], delta[2]};
        double rotUpdate[9];
        smath::rodrigues(dd, rotUpdate);
        float ru[9], rl[9], rn[9];
#pragma unroll
        for(int k = 0; k < 9; k++)
        {
            ru[k] = (float)rotUpdate[k];
            rl[k] = sh.R_lr[k];
        }
        smath::mat3_mul(ru, rl, rn);
        double Rn[9];
#pragma unroll
        for(int k = 0; k < 9; k++) Rn[k] = rn[k];
        __syncwarp();
        if(lane == 0)
        {
            sh.lastError = so3Error;
            sh.lastCount = so3Count;
#pragma unroll
            for(int k = 0; k < 9; k++)
            {
                sh.lastResultR[k] = R[k];
                sh.R_lr[k] = rn[k];
                sh.resultR[k] = Rn[k];
            }
            if(rec)
                for(int k = 0; k < 3; k++) rec->x[k] = delta[k];
        }
        so3_params(sh, Rn, lane == 0);   // the next iteration's H, K^-1, K R
#pragma unroll
        for(int k = 0; k < 9; k++) R[k] = Rn[k];
    }
    if(lane == 0)
    {
        sh.res.lastSO3Error = so3Error;
        sh.res.lastSO3Count = so3Count;
        sh.res.so3_iterations++;
        sh.stop = stop ? 1 : 0;
    }
    __syncwarp();
    if(rec)
        for(int k = 0; k < 9; k++) rec->Rcurr[k] = (float)sh.resultR[k];
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

// lane 0.  The prior pose arrives by value (registers): the persistent kernel reads it from its parameter block.
// CLEAR_RES = false: the caller clears sh.res itself (the persistent kernel does it with another warp meanwhile).
template <bool CLEAR_RES = true>
__device__ __forceinline__ void seq_begin_pose(GnShared & sh, const float (&Rp)[9], const float (&tp)[3])
{
#pragma unroll
    for(int k = 0; k < 9; k++) sh.Rprev[k] = sh.Rcurr[k] = Rp[k];
#pragma unroll
    for(int k = 0; k < 3; k++) sh.tprev[k] = sh.tcurr[k] = tp[k];
    smath::mat3_inverse(sh.Rprev, sh.Rprev_inv);
    for(int k = 0; k < 9; k++)
    {
        sh.resultR[k] = sh.lastResultR[k] = (k % 4 == 0) ? 1.0 : 0.0;
        sh.R_lr[k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    sh.lastError = FLT_MAX / 2;
    sh.lastCount = FLT_MAX / 2;
    if(CLEAR_RES) memset(&sh.res, 0, sizeof(sh.res));
    sh.stop = 0;
    sh.rgb_sigma_last = 0;
    sh.rgb_count_last = -1;
}

static __device__ __noinline__ void seq_begin(GnShared & sh, const GnSeqIn & in)   // lane 0
{
    float Rp[9], tp[3];
    for(int k = 0; k < 9; k++) Rp[k] = in.Rprev[k];
    for(int k = 0; k < 3; k++) tp[k] = in.tprev[k];
    seq_begin_pose(sh, Rp, tp);
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
