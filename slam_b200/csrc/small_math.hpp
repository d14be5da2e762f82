// Small fixed-size linear algebra for the Gauss-Newton bookkeeping of the tracker.
//
// Replaces the Eigen calls the reference makes on this path (Eigen itself is a system
// package that is not vendored in the reference; call sites cited per function):
//   - K.inverse(), Rprev.inverse()                 RGBDOdometryef.cpp:318,323,386,426
//   - resultRt.inverse() (4x4 double)              RGBDOdometryef.cpp:422
//   - jtj.ldlt().solve(jtr) (3x3 float)            RGBDOdometryef.cpp:366
//   - lastA.ldlt().solve(lastb) (6x6 double)       RGBDOdometryef.cpp:544,550,556
//   - lastA.lu().inverse()                         RGBDOdometryef.cpp:599
//   - OdometryProvider::rodrigues / computeUpdateSE3   odom/utils.h:16-74
//
// Every routine is __host__ __device__ so the device-resident Gauss-Newton loop and the
// host-stepped loop run the very same code.  All matrices are row-major.
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define SM_HD __host__ __device__ __forceinline__
#else
#define SM_HD inline
#endif

namespace smath {

// Arithmetic helpers that never contract into FMAs and always divide IEEE-exactly, so the
// device-resident loop (nvcc, -fmad=true, --prec-div=false) and the host-stepped loop (g++)
// produce the same bits from the same inputs.
#if defined(__CUDA_ARCH__)
SM_HD float mul(float a, float b) { return __fmul_rn(a, b); }
SM_HD float add(float a, float b) { return __fadd_rn(a, b); }
SM_HD float sub(float a, float b) { return __fsub_rn(a, b); }
SM_HD float dvd(float a, float b) { return __fdiv_rn(a, b); }
SM_HD double mul(double a, double b) { return __dmul_rn(a, b); }
SM_HD double add(double a, double b) { return __dadd_rn(a, b); }
SM_HD double sub(double a, double b) { return __dsub_rn(a, b); }
SM_HD double dvd(double a, double b) { return __ddiv_rn(a, b); }
#else
SM_HD float mul(float a, float b) { return a * b; }
SM_HD float add(float a, float b) { return a + b; }
SM_HD float sub(float a, float b) { return a - b; }
SM_HD float dvd(float a, float b) { return a / b; }
SM_HD double mul(double a, double b) { return a * b; }
SM_HD double add(double a, double b) { return a + b; }
SM_HD double sub(double a, double b) { return a - b; }
SM_HD double dvd(double a, double b) { return a / b; }
#endif
template <typename T> SM_HD T dot3(T a0, T b0, T a1, T b1, T a2, T b2) { return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2)); }
template <typename T> SM_HD T det2(T a, T b, T c, T d) { return sub(mul(a, b), mul(c, d)); }   // a*b - c*d

template <typename T>
SM_HD void mat3_mul(const T * a, const T * b, T * c)
{
    T r[9];
    for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++)
            r[i * 3 + j] = dot3(a[i * 3 + 0], b[0 * 3 + j], a[i * 3 + 1], b[1 * 3 + j], a[i * 3 + 2], b[2 * 3 + j]);
    for(int i = 0; i < 9; i++) c[i] = r[i];
}

template <typename T>
SM_HD void mat4_mul(const T * a, const T * b, T * c)
{
    T r[16];
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++)
        {
            T s = 0;
            for(int k = 0; k < 4; k++) s = add(s, mul(a[i * 4 + k], b[k * 4 + j]));
            r[i * 4 + j] = s;
        }
    for(int i = 0; i < 16; i++) c[i] = r[i];
}

// Cofactor inverse, the closed form Eigen uses for fixed 3x3.
template <typename T>
SM_HD void mat3_inverse(const T * m, T * out)
{
    const T c00 = det2(m[4], m[8], m[5], m[7]);
    const T c01 = det2(m[5], m[6], m[3], m[8]);
    const T c02 = det2(m[3], m[7], m[4], m[6]);
    const T det = dot3(m[0], c00, m[1], c01, m[2], c02);
    const T id = dvd(T(1), det);
    T r[9];
    r[0] = mul(c00, id);
    r[1] = mul(det2(m[2], m[7], m[1], m[8]), id);
    r[2] = mul(det2(m[1], m[5], m[2], m[4]), id);
    r[3] = mul(c01, id);
    r[4] = mul(det2(m[0], m[8], m[2], m[6]), id);
    r[5] = mul(det2(m[2], m[3], m[0], m[5]), id);
    r[6] = mul(c02, id);
    r[7] = mul(det2(m[1], m[6], m[0], m[7]), id);
    r[8] = mul(det2(m[0], m[4], m[1], m[3]), id);
    for(int i = 0; i < 9; i++) out[i] = r[i];
}

// General 4x4 inverse by cofactors (2x2 sub-determinants).
template <typename T>
SM_HD void mat4_inverse(const T * m, T * out)
{
    const T s0 = det2(m[0], m[5], m[4], m[1]);
    const T s1 = det2(m[0], m[6], m[4], m[2]);
    const T s2 = det2(m[0], m[7], m[4], m[3]);
    const T s3 = det2(m[1], m[6], m[5], m[2]);
    const T s4 = det2(m[1], m[7], m[5], m[3]);
    const T s5 = det2(m[2], m[7], m[6], m[3]);

    const T c5 = det2(m[10], m[15], m[14], m[11]);
    const T c4 = det2(m[9], m[15], m[13], m[11]);
    const T c3 = det2(m[9], m[14], m[13], m[10]);
    const T c2 = det2(m[8], m[15], m[12], m[11]);
    const T c1 = det2(m[8], m[14], m[12], m[10]);
    const T c0 = det2(m[8], m[13], m[12], m[9]);

    // a*x - b*y + c*z
    auto amb = [](T a, T x, T b, T y, T c, T z) { return add(sub(mul(a, x), mul(b, y)), mul(c, z)); };

    const T det = add(amb(s0, c5, s1, c4, s2, c3), amb(s3, c2, s4, c1, s5, c0));
    const T id = dvd(T(1), det);

    T r[16];
    r[0] = mul(amb(m[5], c5, m[6], c4, m[7], c3), id);
    r[1] = mul(-amb(m[1], c5, m[2], c4, m[3], c3), id);
    r[2] = mul(amb(m[13], s5, m[14], s4, m[15], s3), id);
    r[3] = mul(-amb(m[9], s5, m[10], s4, m[11], s3), id);

    r[4] = mul(-amb(m[4], c5, m[6], c2, m[7], c1), id);
    r[5] = mul(amb(m[0], c5, m[2], c2, m[3], c1), id);
    r[6] = mul(-amb(m[12], s5, m[14], s2, m[15], s1), id);
    r[7] = mul(amb(m[8], s5, m[10], s2, m[11], s1), id);

    r[8] = mul(amb(m[4], c4, m[5], c2, m[7], c0), id);
    r[9] = mul(-amb(m[0], c4, m[1], c2, m[3], c0), id);
    r[10] = mul(amb(m[12], s4, m[13], s2, m[15], s0), id);
    r[11] = mul(-amb(m[8], s4, m[9], s2, m[11], s0), id);

    r[12] = mul(-amb(m[4], c3, m[5], c1, m[6], c0), id);
    r[13] = mul(amb(m[0], c3, m[1], c1, m[2], c0), id);
    r[14] = mul(-amb(m[12], s3, m[13], s1, m[14], s0), id);
    r[15] = mul(amb(m[8], s3, m[9], s1, m[10], s0), id);
    for(int i = 0; i < 16; i++) out[i] = r[i];
}

// Robust Cholesky (LDL^T with symmetric diagonal pivoting), the decomposition behind
// Eigen's ldlt().solve(): pivot on the largest remaining |diagonal|, and in the solve
// treat pivots below max|D|*eps as zero (pseudo-inverse), so a rank-deficient system
// (e.g. no correspondences => A = 0) yields 0 rather than NaN.
template <typename T, int N>
SM_HD void ldlt_solve_pivoted(const T * A, const T * b, T * x, T eps)
{
    T m[N * N];
    int perm[N];
    for(int i = 0; i < N * N; i++) m[i] = A[i];
    for(int i = 0; i < N; i++) perm[i] = i;

    for(int k = 0; k < N; k++)
    {
        // largest remaining diagonal entry
        int p = k;
        T best = m[k * N + k] < 0 ? -m[k * N + k] : m[k * N + k];
        for(int i = k + 1; i < N; i++)
        {
            T v = m[i * N + i] < 0 ? -m[i * N + i] : m[i * N + i];
            if(v > best)
            {
                best = v;
                p = i;
            }
        }
        if(p != k)
        {
            for(int j = 0; j < N; j++)
            {
                T t = m[k * N + j];
                m[k * N + j] = m[p * N + j];
                m[p * N + j] = t;
            }
            for(int j = 0; j < N; j++)
            {
                T t = m[j * N + k];
                m[j * N + k] = m[j * N + p];
                m[j * N + p] = t;
            }
            int t = perm[k];
            perm[k] = perm[p];
            perm[p] = t;
        }
        const T d = m[k * N + k];
        if(best > T(0))
        {
            // trailing update with the un-normalised column, then normalise it into L
            for(int i = k + 1; i < N; i++)
            {
                const T l = dvd(m[i * N + k], d);
                for(int j = k + 1; j <= i; j++)
                {
                    m[i * N + j] = sub(m[i * N + j], mul(l, m[j * N + k]));
                    m[j * N + i] = m[i * N + j];
                }
            }
            for(int i = k + 1; i < N; i++) m[i * N + k] = dvd(m[i * N + k], d);   // L below the diagonal
        }
        else
        {
            for(int i = k + 1; i < N; i++) m[i * N + k] = 0;
        }
    }

    T y[N];
    for(int i = 0; i < N; i++) y[i] = b[perm[i]];
    // L y' = y
    for(int i = 0; i < N; i++)
        for(int j = 0; j < i; j++) y[i] = sub(y[i], mul(m[i * N + j], y[j]));
    // D
    T dmax = 0;
    for(int i = 0; i < N; i++)
    {
        T v = m[i * N + i] < 0 ? -m[i * N + i] : m[i * N + i];
        if(v > dmax) dmax = v;
    }
    const T tol = mul(dmax, eps);
    for(int i = 0; i < N; i++)
    {
        T v = m[i * N + i] < 0 ? -m[i * N + i] : m[i * N + i];
        y[i] = (v > tol) ? dvd(y[i], m[i * N + i]) : T(0);
    }
    // L^T z = y
    for(int i = N - 1; i >= 0; i--)
        for(int j = i + 1; j < N; j++) y[i] = sub(y[i], mul(m[j * N + i], y[j]));
    for(int i = 0; i < N; i++) x[perm[i]] = y[i];
}

// LDL^T without pivoting, written so that every index is a compile-time constant after
// unrolling (the whole factorisation lives in registers on the GPU; one reciprocal per
// pivot).  Returns false -- leaving x untouched -- when a pivot is not safely positive
// (d_k <= 1e-9 * largest diagonal entry): the caller then takes the pivoted route above.
// For the well-conditioned SPD normal equations of a tracked frame this path is always taken.
template <typename T, int N>
SM_HD bool ldlt_solve_nopivot(const T * A, const T * b, T * x)
{
    T L[N][N];
    T d[N], inv_d[N], y[N];
    T dmax = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < N; i++) dmax = A[i * N + i] > dmax ? A[i * N + i] : dmax;
    const T floor_d = mul(dmax, T(1e-9));
    bool ok = dmax > T(0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int j = 0; j < N; j++)
    {
        T dj = A[j * N + j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int k = 0; k < j; k++) dj = sub(dj, mul(mul(L[j][k], L[j][k]), d[k]));
        d[j] = dj;
        ok = ok && (dj > floor_d);
        inv_d[j] = dvd(T(1), dj);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int i = j + 1; i < N; i++)
        {
            T v = A[i * N + j];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int k = 0; k < j; k++) v = sub(v, mul(mul(L[i][k], L[j][k]), d[k]));
            L[i][j] = mul(v, inv_d[j]);
        }
    }
    if(!ok) return false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < N; i++)
    {
        T v = b[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int k = 0; k < i; k++) v = sub(v, mul(L[i][k], y[k]));
        y[i] = v;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < N; i++) y[i] = mul(y[i], inv_d[i]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = N - 1; i >= 0; i--)
    {
        T v = y[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int k = i + 1; k < N; k++) v = sub(v, mul(L[k][i], y[k]));
        y[i] = v;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int i = 0; i < N; i++) x[i] = y[i];
    return true;
}

// Gauss-Jordan elimination of the augmented system [A | b] without pivoting.  Every entry's update
// at step k reads only the previous step's values, so the 42 entries of a 6x7 system can be updated
// by 42 independent workers: the device-resident loop spreads them over the lanes of a warp
// (gn_kernel.cu: warp_gauss_jordan) and gets bit-identical results to this serial version.
// Returns false when a pivot is not safely positive (caller falls back to the pivoted LDL^T).
template <typename T, int N>
SM_HD bool gauss_jordan_solve(const T * A, const T * b, T * x)
{
    constexpr int W = N + 1;
    T cur[N * W], nxt[N * W];
    T dmax = 0;
    for(int i = 0; i < N; i++)
    {
        for(int j = 0; j < N; j++) cur[i * W + j] = A[i * N + j];
        cur[i * W + N] = b[i];
        dmax = A[i * N + i] > dmax ? A[i * N + i] : dmax;
    }
    const T floor_d = mul(dmax, T(1e-9));
    bool ok = dmax > T(0);
    for(int k = 0; k < N; k++)
    {
        const T p = cur[k * W + k];
        ok = ok && (p > floor_d);
        const T inv = dvd(T(1), p);
        for(int i = 0; i < N; i++)
            for(int j = 0; j < W; j++)
            {
                T v = cur[i * W + j];
                if(j > k)
                {
                    const T rkj = mul(cur[k * W + j], inv);
                    v = (i == k) ? rkj : sub(v, mul(cur[i * W + k], rkj));
                }
                nxt[i * W + j] = v;
            }
        for(int e = 0; e < N * W; e++) cur[e] = nxt[e];
    }
    if(!ok) return false;
    for(int i = 0; i < N; i++) x[i] = cur[i * W + N];
    return true;
}

// The 6x6 solve of every ICP / RGB step (`lastA.ldlt().solve(lastb)`, RGBDOdometryef.cpp:544-556): the
// parallelisable elimination for positive definite systems, the pivoted / pseudo-inverse LDL^T otherwise.
SM_HD void spd_solve6(const double * A, const double * b, double * x)
{
    if(!gauss_jordan_solve<double, 6>(A, b, x)) ldlt_solve_pivoted<double, 6>(A, b, x, DBL_EPSILON);
}

// The solver behind the remaining `A.ldlt().solve(b)` calls (3x3 float SO3 system): register-resident fast route for
// positive definite systems, pivoted / pseudo-inverse route for degenerate ones.
template <typename T, int N>
SM_HD void ldlt_solve(const T * A, const T * b, T * x, T eps)
{
    if(!ldlt_solve_nopivot<T, N>(A, b, x)) ldlt_solve_pivoted<T, N>(A, b, x, eps);
}

// Inverse of an affine 4x4 [M t; 0 0 0 1] (every resultRt is one: products of such matrices keep
// the last row exactly): [M^-1, -M^-1 t; 0 0 0 1].  Stands in for resultRt.inverse(),
// RGBDOdometryef.cpp:422.
template <typename T>
SM_HD void mat4_affine_inverse(const T * m, T * out)
{
    T M[9], Mi[9];
    for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++) M[i * 3 + j] = m[i * 4 + j];
    mat3_inverse(M, Mi);
    T r[16];
    for(int i = 0; i < 3; i++)
    {
        for(int j = 0; j < 3; j++) r[i * 4 + j] = Mi[i * 3 + j];
        r[i * 4 + 3] = -dot3(Mi[i * 3 + 0], m[3], Mi[i * 3 + 1], m[7], Mi[i * 3 + 2], m[11]);
    }
    r[12] = r[13] = r[14] = 0;
    r[15] = 1;
    for(int i = 0; i < 16; i++) out[i] = r[i];
}

// Axis-angle -> rotation (double), odom/utils.h:16-52; identity below DBL_EPSILON.
SM_HD void rodrigues(const double * r, double * R)
{
    for(int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    double rx = r[0], ry = r[1], rz = r[2];
    const double theta = sqrt(add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz)));
    if(theta >= DBL_EPSILON)
    {
        const double c = cos(theta);
        const double s = sin(theta);
        const double c1 = sub(1., c);
        const double itheta = dvd(1., theta);
        rx = mul(rx, itheta);
        ry = mul(ry, itheta);
        rz = mul(rz, itheta);
        const double rrt[9] = {mul(rx, rx), mul(rx, ry), mul(rx, rz), mul(rx, ry), mul(ry, ry), mul(ry, rz), mul(rx, rz), mul(ry, rz), mul(rz, rz)};
        const double rx_[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        for(int k = 0; k < 9; k++)
        {
            const double I = (k % 4 == 0) ? 1.0 : 0.0;
            R[k] = add(add(mul(c, I), mul(c1, rrt[k])), mul(s, rx_[k]));
        }
    }
}

// resultRt <- [rodrigues(x[3:6]) | x[0:3]] * resultRt   (odom/utils.h:54-68)
SM_HD void update_se3(double * resultRt, const double * x)
{
    double Rt[16];
    double R[9];
    rodrigues(x + 3, R);
    for(int i = 0; i < 3; i++)
    {
        for(int j = 0; j < 3; j++) Rt[i * 4 + j] = R[i * 3 + j];
        Rt[i * 4 + 3] = x[i];
    }
    Rt[12] = Rt[13] = Rt[14] = 0;
    Rt[15] = 1;
    mat4_mul(Rt, resultRt, resultRt);
}

// currentT = [Rprev|tprev] * rgbOdom^-1 with rgbOdom = float(resultRt) taken as an
// isometry (inverse = [R^T | -R^T t]); RGBDOdometryef.cpp:563-575, odom/utils.h:70-73.
// Rcurr is the linear part of the product (Eigen >= 3.4 Isometry::rotation(); Eigen 3.3
// would additionally polar-project it, a <= 1e-7 effect on an already orthonormal R).
SM_HD void compose_current_pose(const float * Rprev, const float * tprev, const double * resultRt, float * Rcurr, float * tcurr)
{
    float Ro[9], to[3];
    for(int i = 0; i < 3; i++)
    {
        for(int j = 0; j < 3; j++) Ro[i * 3 + j] = (float)resultRt[i * 4 + j];
        to[i] = (float)resultRt[i * 4 + 3];
    }
    float Rinv[9], tinv[3];
    for(int i = 0; i < 3; i++)
        for(int j = 0; j < 3; j++) Rinv[i * 3 + j] = Ro[j * 3 + i];
    for(int i = 0; i < 3; i++) tinv[i] = -dot3(Rinv[i * 3 + 0], to[0], Rinv[i * 3 + 1], to[1], Rinv[i * 3 + 2], to[2]);
    float Rc[9], tc[3];
    mat3_mul(Rprev, Rinv, Rc);
    for(int i = 0; i < 3; i++) tc[i] = add(dot3(Rprev[i * 3 + 0], tinv[0], Rprev[i * 3 + 1], tinv[1], Rprev[i * 3 + 2], tinv[2]), tprev[i]);
    for(int i = 0; i < 9; i++) Rcurr[i] = Rc[i];
    for(int i = 0; i < 3; i++) tcurr[i] = tc[i];
}

// General NxN inverse by LU with partial pivoting (getCovariance, RGBDOdometryef.cpp:597-600).
template <typename T, int N>
SM_HD bool lu_inverse(const T * A, T * out)
{
    T m[N * N];
    T inv[N * N];
    for(int i = 0; i < N * N; i++)
    {
        m[i] = A[i];
        inv[i] = ((i / N) == (i % N)) ? T(1) : T(0);
    }
    bool ok = true;
    for(int k = 0; k < N; k++)
    {
        int p = k;
        T best = m[k * N + k] < 0 ? -m[k * N + k] : m[k * N + k];
        for(int i = k + 1; i < N; i++)
        {
            T v = m[i * N + k] < 0 ? -m[i * N + k] : m[i * N + k];
            if(v > best)
            {
                best = v;
                p = i;
            }
        }
        if(best == T(0)) ok = false;
        if(p != k)
            for(int j = 0; j < N; j++)
            {
                T t = m[k * N + j];
                m[k * N + j] = m[p * N + j];
                m[p * N + j] = t;
                t = inv[k * N + j];
                inv[k * N + j] = inv[p * N + j];
                inv[p * N + j] = t;
            }
        const T d = T(1) / m[k * N + k];
        for(int j = 0; j < N; j++)
        {
            m[k * N + j] *= d;
            inv[k * N + j] *= d;
        }
        for(int i = 0; i < N; i++)
        {
            if(i == k) continue;
            const T f = m[i * N + k];
            for(int j = 0; j < N; j++)
            {
                m[i * N + j] -= f * m[k * N + j];
                inv[i * N + j] -= f * inv[k * N + j];
            }
        }
    }
    for(int i = 0; i < N * N; i++) out[i] = inv[i];
    return ok;
}

}   // namespace smath
