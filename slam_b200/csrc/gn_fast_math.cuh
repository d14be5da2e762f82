// Latency-optimised fp64 bookkeeping of one Gauss-Newton step for the device-resident loops (warp 0 of a CTA).
//
// The step between two reductions is a single dependent chain (combine the systems -> 6x6 solve -> Rodrigues ->
// resultRt -> inverse -> K R K^-1 / K t / current pose, RGBDOdometryef.cpp:509-575 and :422-432); on B200 a dependent
// DFMA costs ~13 cycles, an fp64 division ~80, a 64-bit shuffle ~45 and sin + cos ~480 (tools/micro/sync_latency.cu), so
// the chain is shortened rather than parallelised:
//   * the 6x6 system is eliminated WITHOUT divisions on the chain (fraction-free Gauss-Jordan on a power-of-two
//     prescaled system: rows are multiplied by the pivot instead of the pivot row being divided; the six divisions that
//     remain are independent and happen once, at the end);
//   * the rotation increment uses the series of sin(t)/t and (1 - cos t)/t^2 in t^2 (|increment| <= 0.1 rad, far above any
//     Gauss-Newton step of a tracked frame; larger angles take the textbook route), no square root, no division;
//   * the products with K and K^-1 use their zero pattern;
//   * everything is FMA-contracted.
// The textbook forms (small_math.hpp: LDL^T / Gauss-Jordan with divisions, libm sin / cos, generic 3x3 products) remain the
// ones the host-stepped loop and the reference replay run; the two agree to ~1e-15 relative (tests compare them per step).
#pragma once
#include "gn_scalar.cuh"

namespace slam {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1 / d for a positive, finite, normal d: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps (error ~1e-24 relative
// before rounding).  Branch-free, so independent reciprocals overlap (the IEEE division carries a slow-path call that serialises them).
__device__ __forceinline__ double rcp_newton(const double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    y = __fma_rn(__fma_rn(-d, y, 1.0), y, y);
    y = __fma_rn(__fma_rn(-d, y, 1.0), y, y);
    y = __fma_rn(__fma_rn(-d, y, 1.0), y, y);
    return y;
}

// x = A^-1 b for the SPD system whose column min(lane, 6) of [A | b] is c[0..5] (lane j < 6: column j of A, lanes >= 6:
// b).  Fraction-free Gauss-Jordan: at step k every row i != k becomes row_i * p_k - a_ik * row_k (p_k = running pivot),
// so no division sits between two steps.  Returns false (uniformly) when a pivot is not safely positive -- the caller
// then takes the pivoted LDL^T of small_math.hpp.  All lanes return the same x.
// The step loop is deliberately NOT unrolled (the persistent kernel's per-iteration code has to stay inside the instruction
// cache): the rows are rotated by one position per step instead, so that the pivot row is always c[0] and every index in the
// body is a constant.  After six steps the rows are back in place.
__device__ __forceinline__ bool warp_solve6_fraction_free(double (&c)[6], double (&x)[6])
{
    const int lane = threadIdx.x & 31;
    const int j = lane < 7 ? lane : 6;
    // largest diagonal entry (the pivots are compared with 1e-9 of it, as smath::gauss_jordan_solve does)
    double diag = c[0];
#pragma unroll
    for(int q = 1; q < 6; q++) diag = (j == q) ? c[q] : diag;   // lane q < 6: a_qq
    double dmax = 0;
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
        const double d = shfl_d(diag, i);
        dmax = d > dmax ? d : dmax;
    }
    bool ok = dmax > 0.0 && dmax < 1.7e308;
    const double floor_d = dmax * 1e-9;
    double prod = 1.0;   // product of the (rescaled) pivots used so far: the LDL^T pivot d_k equals p_k / prod
#pragma unroll 1
    for(int k = 0; k < 6; k++)
    {
        // c[0] is row k of this lane's column; column k (lane k) in the same rotated order
        double colk[6];
#pragma unroll
        for(int i = 0; i < 6; i++) colk[i] = shfl_d(c[i], k);
        const double p = colk[0];
        ok = ok && (p > floor_d * prod);
        // multiplier rescaled into [1, 2) by an exact power of two, so that the rows neither grow nor decay step after step
        // (a non-positive or non-finite pivot makes this scale meaningless, but then ok is false and x is replaced)
        const double sk = __hiloint2double(0x7fe00000 - (__double2hiint(p) & 0x7ff00000), 0);
        const double ps = p * sk;
        // row k of a column that is already eliminated (j < k) is zero by construction; it is not stored, so it is forced here
        const double s2 = (j < k) ? 0.0 : c[0] * sk;
        const double keep = c[0];
#pragma unroll
        for(int i = 0; i < 5; i++) c[i] = __fma_rn(c[i + 1], ps, -(colk[i + 1] * s2));
        c[5] = keep;
        prod *= ps;
    }
    // the system is diagonal now: x_i = b_i / a_ii (a_ii sits in lane i, row i)
    diag = c[0];
#pragma unroll
    for(int q = 1; q < 6; q++) diag = (j == q) ? c[q] : diag;
    const double rd = rcp_newton(diag);   // every lane inverts its own entry (lanes < 6: the diagonal), then the six products
#pragma unroll
    for(int i = 0; i < 6; i++)
    {
        const double b = shfl_d(c[i], 6), d = shfl_d(diag, i), r = shfl_d(rd, i);
        const double q0 = b * r;
        x[i] = __fma_rn(__fma_rn(-d, q0, b), r, q0);   // one residual correction: within an ulp of the quotient
    }
    return ok;
}

static __device__ __noinline__ void rodrigues_textbook(const double wx, const double wy, const double wz, double (&R)[9])
{
    const double w[3] = {wx, wy, wz};
    double r[9];
    smath::rodrigues(w, r);
    for(int q = 0; q < 9; q++) R[q] = r[q];
}

// Rotation of the axis-angle vector w (odom/utils.h:16-52): R = I + A [w]x + B [w]x^2 with A = sin(t)/t, B = (1 - cos t)/t^2.
__device__ __forceinline__ void rodrigues_fast(const double wx, const double wy, const double wz, double (&R)[9])
{
    const double xx = wx * wx, yy = wy * wy, zz = wz * wz;
    const double t2 = xx + yy + zz;
    if(t2 > 0.01)   // uniform; 0.1 rad per Gauss-Newton step never happens on a tracked frame
    {
        rodrigues_textbook(wx, wy, wz, R);
        return;
    }
    // series in t2, 7 terms: truncation below 1e-25 for t2 <= 0.01
    double A = 1.0 / 6227020800.0;            // 1/13!
    double B = 1.0 / 87178291200.0;           // 1/14!
    A = __fma_rn(A, t2, -1.0 / 39916800.0);            // -1/11!
    B = __fma_rn(B, t2, -1.0 / 479001600.0);           // -1/12!
    A = __fma_rn(A, t2, 1.0 / 362880.0);               // 1/9!
    B = __fma_rn(B, t2, 1.0 / 3628800.0);              // 1/10!
    A = __fma_rn(A, t2, -1.0 / 5040.0);                // -1/7!
    B = __fma_rn(B, t2, -1.0 / 40320.0);               // -1/8!
    A = __fma_rn(A, t2, 1.0 / 120.0);                  // 1/5!
    B = __fma_rn(B, t2, 1.0 / 720.0);                  // 1/6!
    A = __fma_rn(A, t2, -1.0 / 6.0);                   // -1/3!
    B = __fma_rn(B, t2, -1.0 / 24.0);                  // -1/4!
    A = __fma_rn(A, t2, 1.0);
    B = __fma_rn(B, t2, 0.5);
    const double Bxy = B * wx * wy, Bxz = B * wx * wz, Byz = B * wy * wz;
    R[0] = __fma_rn(-B, yy + zz, 1.0);
    R[1] = __fma_rn(-A, wz, Bxy);
    R[2] = __fma_rn(A, wy, Bxz);
    R[3] = __fma_rn(A, wz, Bxy);
    R[4] = __fma_rn(-B, xx + zz, 1.0);
    R[5] = __fma_rn(-A, wx, Byz);
    R[6] = __fma_rn(-A, wy, Bxz);
    R[7] = __fma_rn(A, wx, Byz);
    R[8] = __fma_rn(-B, xx + yy, 1.0);
}

// The intrinsics of the running level as the scalars the structured products need (K = [fx 0 cx; 0 fy cy; 0 0 1]).
struct KParams
{
    double fx, fy, cx, cy;          // K
    double ifx, ify, icx, icy;      // K^-1 = [ifx 0 icx; 0 ify icy; 0 0 1]
};
__device__ __forceinline__ KParams k_params(const GnShared & sh)
{
    KParams k;
    k.fx = sh.K[0]; k.fy = sh.K[4]; k.cx = sh.K[2]; k.cy = sh.K[5];
    k.ifx = sh.Kinv[0]; k.ify = sh.Kinv[4]; k.icx = sh.Kinv[2]; k.icy = sh.Kinv[5];
    return k;
}

// H = K R K^-1 and KR = K R with the zero pattern of K (same values as the generic products: the skipped terms are exact zeros).
__device__ __forceinline__ void krk_structured(const KParams & k, const double (&R)[9], double (&KR)[9], double (&H)[9])
{
#pragma unroll
    for(int q = 0; q < 3; q++)
    {
        KR[0 + q] = __fma_rn(k.cx, R[6 + q], k.fx * R[0 + q]);
        KR[3 + q] = __fma_rn(k.cy, R[6 + q], k.fy * R[3 + q]);
        KR[6 + q] = R[6 + q];
    }
#pragma unroll
    for(int i = 0; i < 3; i++)
    {
        H[i * 3 + 0] = KR[i * 3 + 0] * k.ifx;
        H[i * 3 + 1] = KR[i * 3 + 1] * k.ify;
        H[i * 3 + 2] = __fma_rn(KR[i * 3 + 1], k.icy, __fma_rn(KR[i * 3 + 0], k.icx, KR[i * 3 + 2]));
    }
}

// Parameters of the next iteration from the affine rows M (3x4, row-major) of resultRt and Mi = the inverse of its linear part:
// krk = float(K R K^-1), kt = float(K t) with [R | t] = resultRt^-1 (RGBDOdometryef.cpp:422-432) and, with_pose, Rcurr / tcurr =
// [Rprev | tprev] * float(resultRt)^-1 (:563-575).  Warp 0, all lanes evaluate everything; lane 0 stores (also Mi, for the next step).
__device__ __forceinline__ void prepare_tail(GnShared & sh, const KParams & k, const double (&M)[12], const double (&Mi)[9], const bool with_pose)
{
    const int lane = threadIdx.x & 31;
    double tinv[3];
#pragma unroll
    for(int i = 0; i < 3; i++) tinv[i] = -__fma_rn(Mi[i * 3 + 2], M[11], __fma_rn(Mi[i * 3 + 1], M[7], Mi[i * 3 + 0] * M[3]));
    double KR[9], H[9];
    krk_structured(k, Mi, KR, H);
    const float kt0 = (float)__fma_rn(k.cx, tinv[2], k.fx * tinv[0]);
    const float kt1 = (float)__fma_rn(k.cy, tinv[2], k.fy * tinv[1]);
    const float kt2 = (float)tinv[2];
    if(lane == 0)
    {
#pragma unroll
        for(int q = 0; q < 9; q++)
        {
            sh.krk[q] = (float)H[q];
            sh.Mi[q] = Mi[q];
        }
        sh.kt[0] = kt0; sh.kt[1] = kt1; sh.kt[2] = kt2;
    }
    if(with_pose)
    {
        // odom/utils.h:70-73 + RGBDOdometryef.cpp:563-575 in fp32, the expressions of smath::compose_current_pose
        float Rp[9], tp[3], Mf[12], tinvf[3];
#pragma unroll
        for(int q = 0; q < 9; q++) Rp[q] = sh.Rprev[q];
#pragma unroll
        for(int q = 0; q < 3; q++) tp[q] = sh.tprev[q];
#pragma unroll
        for(int q = 0; q < 12; q++) Mf[q] = (float)M[q];
#pragma unroll
        for(int i = 0; i < 3; i++) tinvf[i] = -smath::dot3(Mf[0 * 4 + i], Mf[3], Mf[1 * 4 + i], Mf[7], Mf[2 * 4 + i], Mf[11]);
        float Rc[9], tc[3];
#pragma unroll
        for(int i = 0; i < 3; i++)
        {
#pragma unroll
            for(int q = 0; q < 3; q++) Rc[i * 3 + q] = smath::dot3(Rp[i * 3 + 0], Mf[q * 4 + 0], Rp[i * 3 + 1], Mf[q * 4 + 1], Rp[i * 3 + 2], Mf[q * 4 + 2]);
            tc[i] = smath::add(smath::dot3(Rp[i * 3 + 0], tinvf[0], Rp[i * 3 + 1], tinvf[1], Rp[i * 3 + 2], tinvf[2]), tp[i]);
        }
        if(lane == 0)
        {
#pragma unroll
            for(int q = 0; q < 9; q++) sh.Rcurr[q] = Rc[q];
#pragma unroll
            for(int q = 0; q < 3; q++) sh.tcurr[q] = tc[q];
        }
    }
    __syncwarp();
}

// The parameters of the first iteration of a level (Rcurr/tcurr carry over): resultRt comes from shared memory, and the inverse
// of its linear part is formed by cofactors (the closed form Eigen uses for 3x3) -- once per level; the steps update it.
__device__ __forceinline__ void warp_prepare_fast(GnShared & sh, const bool with_pose)
{
    double M[12];
#pragma unroll
    for(int q = 0; q < 12; q++) M[q] = sh.resultRt[q];
    const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[4], m4 = M[5], m5 = M[6], m6 = M[8], m7 = M[9], m8 = M[10];
    const double c00 = __fma_rn(m4, m8, -(m5 * m7));
    const double c01 = __fma_rn(m5, m6, -(m3 * m8));
    const double c02 = __fma_rn(m3, m7, -(m4 * m6));
    const double det = __fma_rn(m2, c02, __fma_rn(m1, c01, m0 * c00));
    const double id = __ddiv_rn(1.0, det);
    double Mi[9];
    Mi[0] = c00 * id;
    Mi[1] = __fma_rn(m2, m7, -(m1 * m8)) * id;
    Mi[2] = __fma_rn(m1, m5, -(m2 * m4)) * id;
    Mi[3] = c01 * id;
    Mi[4] = __fma_rn(m0, m8, -(m2 * m6)) * id;
    Mi[5] = __fma_rn(m2, m3, -(m0 * m5)) * id;
    Mi[6] = c02 * id;
    Mi[7] = __fma_rn(m1, m6, -(m0 * m7)) * id;
    Mi[8] = __fma_rn(m0, m4, -(m1 * m3)) * id;
    prepare_tail(sh, k_params(sh), M, Mi, with_pose);
}

// Entry (i, j) of the 6x7 augmented system <-> index in the row-major upper triangle the reductions produce (reduce.cu:475-486).
__device__ __forceinline__ int se3_index(const int i, const int j)
{
    const int a = i < j ? i : j, b = i < j ? j : i;
    return 7 * a - (a * (a - 1)) / 2 + (b - a);
}

// RGBDOdometryef.cpp:509-575 on warp 0: combine the two systems (lastA = A_rgb + w^2 A_icp, lastb = b_rgb + w b_icp), solve,
// update resultRt, derive the next iteration's parameters.  icp sums = total[0..28], rgb sums = total[32..60].
// The combined entries stay in sh.Ab (27 doubles, upper triangle order): lastA / lastb of the step for the statistics.
//
// One warp runs this between two reductions while the other fifteen wait, and it runs it from the instruction caches' slow
// side: straight-line code of a single warp is fetched at ~3 cycles per instruction from the 32 KB L1.5 and ~6.5 from L2
// (tools/micro/icache.cu), so what counts is the NUMBER of instructions.  Everything after the elimination is therefore laid
// out across the lanes (one output entry per lane, operands through shared memory) in three stages instead of being evaluated
// redundantly by all lanes: ~5x fewer instructions than the all-lanes form, at the price of three shared-memory round trips.
__device__ __forceinline__ void warp_update_fast(GnShared & sh, const bool icp, const bool rgb, const float icpWeight, slam_step_record * rec, const long long t_start)
{
    const int lane = threadIdx.x & 31;
    const int j = lane < 7 ? lane : 6;
#define GN_FSTAMP(idx) do { if(rec) rec->t_solve[idx] = (unsigned)(clock64() - t_start); } while(0)
    // ---- entry `lane` of the combined system -> 6x7 matrix (row stride 8) in shared memory, both triangles
    if(lane < 27)
    {
        // row i of the upper triangle that entry `lane` belongs to (rows start at 0, 7, 13, 18, 22, 25), branch-free
        const int i = (lane >= 7) + (lane >= 13) + (lane >= 18) + (lane >= 22) + (lane >= 25);
        const int jj = lane - (7 * i - (i * (i - 1)) / 2) + i;
        const float vi = sh.total[lane];
        const float vr = sh.total[32 + lane];
        const double w = icpWeight;
        const double ww = (jj == 6) ? w : w * w;
        const double v = (icp && rgb) ? smath::add((double)vr, smath::mul(ww, (double)vi)) : (icp ? (double)vi : (double)vr);
        sh.Ab[lane] = v;
        sh.Wsys[i * 8 + jj] = v;
        if(jj < 6) sh.Wsys[jj * 8 + i] = v;
    }
    __syncwarp();
    double c[6];
#pragma unroll
    for(int i = 0; i < 6; i++) c[i] = sh.Wsys[i * 8 + j];
    // largest diagonal entry, to 2^-20: positive doubles order like their high words (a negative or NaN entry makes the
    // maximum absurd and the test below fail, which is the right outcome)
    const unsigned dhi = __reduce_max_sync(0xffffffffu, lane < 6 ? (unsigned)__double2hiint(sh.Wsys[lane * 9]) : 0u);
    bool ok = dhi > 0u && dhi < 0x7ff00000u;
    const double floor_d = __hiloint2double((int)dhi, 0) * 1e-9;
    GN_FSTAMP(0);
    // ---- fraction-free Gauss-Jordan (see warp_solve6_fraction_free): rows rotate, the pivot row is always c[0]
    double prod = 1.0;
#pragma unroll 1
    for(int k = 0; k < 6; k++)
    {
        double colk[6];
#pragma unroll
        for(int i = 0; i < 6; i++) colk[i] = shfl_d(c[i], k);
        const double p = colk[0];
        ok = ok && (p > floor_d * prod);
        const double sk = __hiloint2double(0x7fe00000 - (__double2hiint(p) & 0x7ff00000), 0);
        const double ps = p * sk;
        const double s2 = (j < k) ? 0.0 : c[0] * sk;
        const double keep = c[0];
#pragma unroll
        for(int i = 0; i < 5; i++) c[i] = __fma_rn(c[i + 1], ps, -(colk[i + 1] * s2));
        c[5] = keep;
        prod *= ps;
    }
    // ---- x_i = b_i / a_ii: the columns go back to shared memory, lane i < 6 divides
    __syncwarp();   // every lane has read its column of Wsys
    if(lane < 7)
    {
#pragma unroll
        for(int i = 0; i < 6; i++) sh.Wsys[i * 8 + lane] = c[i];
    }
    __syncwarp();
    if(lane < 6)
    {
        const double d = sh.Wsys[lane * 9], b = sh.Wsys[lane * 8 + 6];
        const double r = rcp_newton(d);
        const double q0 = b * r;
        sh.x[lane] = __fma_rn(__fma_rn(-d, q0, b), r, q0);   // one residual correction: within an ulp of the quotient
    }
    __syncwarp();
    GN_FSTAMP(1);
    if(!ok)   // uniform: every lane saw the same pivots
    {
        if(lane == 0) solve_fallback(sh, icp, rgb, icpWeight);
        __syncwarp();
    }
    // ---- incremental rotation of the step (odom/utils.h:16-52), every lane; lane 0 publishes it
    {
        double Rinc[9];
        rodrigues_fast(sh.x[3], sh.x[4], sh.x[5], Rinc);
        if(lane == 0)
        {
#pragma unroll
            for(int q = 0; q < 9; q++) sh.Rinc[q] = Rinc[q];
        }
    }
    __syncwarp();
    GN_FSTAMP(2);
    // ---- stage 1.  lanes 0..11: resultRt <- [Rinc | x[0:3]; 0 0 0 1] * resultRt (odom/utils.h:54-68), entry (i, q) of rows 0..2;
    //      lanes 12..20: inverse of the new linear part, (Rinc R)^-1 = R^-1 Rinc^T (Rinc is a rotation to the last bit;
    //      resultRt.inverse(), RGBDOdometryef.cpp:422, without a determinant and a division on the chain)
    if(lane < 21)
    {
        const bool isM = lane < 12;
        const int r = isM ? lane : lane - 12;
        const int i = isM ? (r >> 2) : r / 3;
        const int q = isM ? (r & 3) : r - 3 * i;
        const double * pa = isM ? &sh.Rinc[i * 3] : &sh.Mi[i * 3];
        const double * pb = isM ? &sh.resultRt[q] : &sh.Rinc[q * 3];
        const int sb = isM ? 4 : 1;
        double s = __fma_rn(pa[2], pb[2 * sb], __fma_rn(pa[1], pb[sb], pa[0] * pb[0]));
        if(isM && q == 3) s += sh.x[i];
        sh.Wm[lane] = s;
    }
    __syncwarp();
    GN_FSTAMP(3);
    // ---- stage 2.  lanes 0..2: t' = -M^-1 t; 3..11: K M^-1; 12..14: translation of float(resultRt)^-1; 15..23: Rcurr
    //      (odom/utils.h:70-73 + RGBDOdometryef.cpp:563-575 in fp32, the expressions of smath::compose_current_pose);
    //      every lane < 21 also files its stage-1 entry in its permanent place
    if(lane < 12)
    {
        const double mine = sh.Wm[lane];
        const int i = lane < 3 ? lane : (lane - 3) / 3;
        const int q = lane < 3 ? 0 : (lane - 3) - 3 * i;
        const double * pa = lane < 3 ? &sh.Wm[12 + 3 * i] : &sh.K[i * 3];
        const double * pb = lane < 3 ? &sh.Wm[3] : &sh.Wm[12 + q];
        const int sb = lane < 3 ? 4 : 3;
        const double s = __fma_rn(pa[2], pb[2 * sb], __fma_rn(pa[1], pb[sb], pa[0] * pb[0]));
        sh.Wk[lane] = lane < 3 ? -s : s;
        sh.resultRt[lane] = mine;
    }
    else if(lane < 24)
    {
        if(lane < 21) sh.Mi[lane - 12] = sh.Wm[lane];
        const int r = lane < 15 ? lane - 12 : lane - 15;
        const int i = lane < 15 ? r : r / 3;
        const int q = lane < 15 ? 0 : r - 3 * i;
        if(lane < 15)
        {
            // tinv[i] = -(Rinv[i][:] . to), Rinv = Ro^T, Ro/to = float(resultRt)
            const float a0 = (float)sh.Wm[0 * 4 + i], a1 = (float)sh.Wm[1 * 4 + i], a2 = (float)sh.Wm[2 * 4 + i];
            sh.Wf[i] = -smath::dot3(a0, (float)sh.Wm[3], a1, (float)sh.Wm[7], a2, (float)sh.Wm[11]);
        }
        else
        {
            // Rcurr = Rprev * Rinv
            sh.Rcurr[r] = smath::dot3(sh.Rprev[i * 3 + 0], (float)sh.Wm[q * 4 + 0], sh.Rprev[i * 3 + 1], (float)sh.Wm[q * 4 + 1], sh.Rprev[i * 3 + 2], (float)sh.Wm[q * 4 + 2]);
        }
    }
    __syncwarp();
    // ---- stage 3.  lanes 0..8: krk = float(K M^-1 K^-1); 9..11: kt = float(K t') (RGBDOdometryef.cpp:422-432); 12..14: tcurr
    if(lane < 12)
    {
        const int i = lane < 9 ? lane / 3 : lane - 9;
        const int q = lane < 9 ? lane - 3 * i : 0;
        const double * pa = lane < 9 ? &sh.Wk[3 + i * 3] : &sh.K[i * 3];
        const double * pb = lane < 9 ? &sh.Kinv[q] : &sh.Wk[0];
        const int sb = lane < 9 ? 3 : 1;
        const float s = (float)__fma_rn(pa[2], pb[2 * sb], __fma_rn(pa[1], pb[sb], pa[0] * pb[0]));
        if(lane < 9)
            sh.krk[lane] = s;
        else
            sh.kt[lane - 9] = s;
    }
    else if(lane < 15)
    {
        const int i = lane - 12;
        // tcurr = Rprev * tinv + tprev
        sh.tcurr[i] = smath::add(smath::dot3(sh.Rprev[i * 3 + 0], sh.Wf[0], sh.Rprev[i * 3 + 1], sh.Wf[1], sh.Rprev[i * 3 + 2], sh.Wf[2]), sh.tprev[i]);
    }
    __syncwarp();
    GN_FSTAMP(4);
    if(lane == 0)
    {
        sh.res.gn_iterations++;
        if(rec)
        {
            for(int q = 0; q < 29; q++)
            {
                rec->icp[q] = icp ? sh.total[q] : 0.f;
                rec->rgb[q] = rgb ? sh.total[32 + q] : 0.f;
            }
            for(int q = 0; q < 6; q++) rec->x[q] = sh.x[q];
            for(int q = 0; q < 9; q++) rec->Rcurr[q] = sh.Rcurr[q];
            for(int q = 0; q < 3; q++) rec->tcurr[q] = sh.tcurr[q];
        }
    }
#undef GN_FSTAMP
}

// lastA / lastb / lastICPError / lastICPCount of the LAST step from sh.Ab and the folded sums (RGBDOdometryef.cpp:509-556).
// One warp, after the loop.
__device__ __forceinline__ void warp_stats_fast(GnShared & sh, const bool icp)
{
    const int lane = threadIdx.x & 31;
    if(lane < 27)
    {
        const int i = (lane >= 7) + (lane >= 13) + (lane >= 18) + (lane >= 22) + (lane >= 25);
        const int j = lane - (7 * i - (i * (i - 1)) / 2) + i;
        const double v = sh.Ab[lane];
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

// H = K R K^-1, K^-1, K R of an SO3 iteration (RGBDOdometryef.cpp:313-323) from resultR.
__device__ __forceinline__ void so3_params_fast(GnShared & sh, const KParams & k, const double (&R)[9], const bool store)
{
    double KR[9], H[9];
    krk_structured(k, R, KR, H);
    if(store)
    {
#pragma unroll
        for(int q = 0; q < 9; q++)
        {
            sh.so3H[q] = (float)H[q];
            sh.so3KR[q] = (float)KR[q];
        }
        sh.so3Kinv[0] = (float)k.ifx; sh.so3Kinv[1] = 0.f; sh.so3Kinv[2] = (float)k.icx;
        sh.so3Kinv[3] = 0.f; sh.so3Kinv[4] = (float)k.ify; sh.so3Kinv[5] = (float)k.icy;
        sh.so3Kinv[6] = 0.f; sh.so3Kinv[7] = 0.f; sh.so3Kinv[8] = 1.f;
    }
}

// RGBDOdometryef.cpp:346-378 followed by the parameters of the next SO3 iteration (:313-323).  One full warp, all lanes
// evaluate the same expressions in registers, lane 0 stores.
__device__ __forceinline__ void warp_so3_update_fast(GnShared & sh, int it, slam_step_record * rec)
{
    const int lane = threadIdx.x & 31;
    float s[11];
#pragma unroll
    for(int q = 0; q < 11; q++) s[q] = sh.total[q];
    float jtj[9], jtr[3];
    {
        int shift = 0;
#pragma unroll
        for(int i = 0; i < 3; ++i)
#pragma unroll
            for(int q = i; q < 4; ++q)
            {
                const float value = s[shift++];
                if(q == 3)
                    jtr[i] = value;
                else
                    jtj[q * 3 + i] = jtj[i * 3 + q] = value;
            }
    }
    float so3Error = __fdiv_rn(__fsqrt_rn(s[9]), s[10]);
    float so3Count = s[10];
    const float lastError = sh.lastError, lastCount = sh.lastCount;
    double R[9];
#pragma unroll
    for(int q = 0; q < 9; q++) R[q] = sh.resultR[q];
    const KParams kp = k_params(sh);

    if(rec)
    {
        rec->kind = 0;
        rec->level = 2;
        rec->iteration = it;
        for(int q = 0; q < 11; q++) rec->so3[q] = s[q];
        for(int q = 0; q < 9; q++)
        {
            rec->so3_in[q] = sh.so3H[q];
            rec->so3_in[9 + q] = sh.so3Kinv[q];
            rec->so3_in[18 + q] = sh.so3KR[q];
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
            for(int q = 0; q < 3; q++) delta[q] = (float)sh.x[q];
            __syncwarp();
        }
        double rotUpdate[9];
        rodrigues_fast((double)delta[0], (double)delta[1], (double)delta[2], rotUpdate);
        float ru[9], rl[9], rn[9];
#pragma unroll
        for(int q = 0; q < 9; q++)
        {
            ru[q] = (float)rotUpdate[q];
            rl[q] = sh.R_lr[q];
        }
        smath::mat3_mul(ru, rl, rn);
        double Rn[9];
#pragma unroll
        for(int q = 0; q < 9; q++) Rn[q] = rn[q];
        __syncwarp();
        if(lane == 0)
        {
            sh.lastError = so3Error;
            sh.lastCount = so3Count;
#pragma unroll
            for(int q = 0; q < 9; q++)
            {
                sh.lastResultR[q] = R[q];
                sh.R_lr[q] = rn[q];
                sh.resultR[q] = Rn[q];
            }
            if(rec)
                for(int q = 0; q < 3; q++) rec->x[q] = delta[q];
        }
        so3_params_fast(sh, kp, Rn, lane == 0);   // the next iteration's H, K^-1, K R
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
        for(int q = 0; q < 9; q++) rec->Rcurr[q] = (float)sh.resultR[q];
}

}   // namespace slam
