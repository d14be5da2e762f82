// Per-pixel arithmetic of the tracker, written once and used by every kernel (the
// single-launch operator kernels and the persistent Gauss-Newton kernel).
//
// Parity contract (BASELINE.json north_star): correspondence masks and pyramid images
// must match the reference bit for bit, so every value that feeds a rounding
// (__float2int_rn), a truncation (float->int/short/u8) or a threshold compare is
// computed with the same expression shape as the reference and this file is compiled
// with the reference's numeric flags (--ftz=true --prec-div=false --prec-sqrt=false,
// src/CMakeLists.txt:115-116).  Accumulation / reduction order is free (1e-4 relative).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slam {

#define SLAM_QNAN __int_as_float(0x7fffffff)   // the reference's NaN sentinel (utils.cu:130)

struct Mat3
{
    float3 r0, r1, r2;   // rows
};

__device__ __forceinline__ float3 operator-(const float3 & a, const float3 & b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(const float3 & a, const float3 & b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
// dot / cross are written with explicit roundings: exactly the instruction sequence nvcc
// (-fmad=true) emits for the reference's `a.x*b.x + a.y*b.y + a.z*b.z` and
// `a.y*b.z - a.z*b.y` (cuda/operators.cuh:67-75), read off the reference's sm_100a SASS
// (computeNmapKernel, resizeMapKernel, icpKernel): the middle product is rounded, the outer
// two are fused.  Pinning them keeps fused kernels bit-identical to the reference whatever
// contraction choices the compiler would make in a different inlining context.
__device__ __forceinline__ float dot3(const float3 & a, const float3 & b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
__device__ __forceinline__ float3 cross3(const float3 & a, const float3 & b)
{
    return make_float3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float norm3(const float3 & a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 unit3(const float3 & a)   // cuda/operators.cuh:82-86 (rsqrtf)
{
    const float rn = rsqrtf(dot3(a, a));
    return make_float3(a.x * rn, a.y * rn, a.z * rn);
}
__device__ __forceinline__ float3 operator*(const Mat3 & m, const float3 & a) { return make_float3(dot3(m.r0, a), dot3(m.r1, a), dot3(m.r2, a)); }

// `a / b` exactly as nvcc --prec-div=false lowers it on sm_100a (div.approx.ftz.f32, seen in the
// reference's SASS as MUFU.RCP + FMUL with both operands pre-scaled by 1/4 when |b| > 2^126),
// split in two so that the final multiply can be fused with a following add where ptxas does so
// in the reference (icpKernel: x*fx/z + cx  ->  FFMA(rcp, x*fx, cx)).
struct ApproxDivisor
{
    float rcp;
    bool big;
};
__device__ __forceinline__ ApproxDivisor approx_divisor(float b)
{
    ApproxDivisor d;
    d.big = fabsf(b) > 8.50705917302346158658e+37f;
    const float bs = d.big ? __fmul_rn(b, 0.25f) : b;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(d.rcp) : "f"(bs));
    return d;
}
__device__ __forceinline__ float approx_div(float a, const ApproxDivisor & d) { return __fmul_rn(d.rcp, d.big ? __fmul_rn(a, 0.25f) : a); }
__device__ __forceinline__ float approx_div_add(float a, const ApproxDivisor & d, float c) { return __fmaf_rn(d.rcp, d.big ? __fmul_rn(a, 0.25f) : a, c); }

__host__ __device__ inline Mat3 mat3_from(const float * p)
{
    Mat3 m;
    m.r0 = make_float3(p[0], p[1], p[2]);
    m.r1 = make_float3(p[3], p[4], p[5]);
    m.r2 = make_float3(p[6], p[7], p[8]);
    return m;
}

// ------------------------------------------------------------------------------------
// ICP: projective data association + point-to-plane row.   reduce.cu:282-348
// ------------------------------------------------------------------------------------
struct IcpArgs
{
    Mat3 Rcurr;
    float3 tcurr;
    Mat3 Rprev_inv;
    float3 tprev;
    float fx, fy, cx, cy;
    float distThres, angleThres;
    int cols, rows;
    const float * vcurr;   // planar [3][rows][cols], current frame, camera frame
    const float * ncurr;
    const float * vprev;   // planar, model prediction, global frame
    const float * nprev;
};

// The association is split in two so that a kernel can issue the gathers of several pixels before
// consuming any of them (memory-level parallelism): icp_project() needs only the current vertex,
// icp_finish() the gathered model vertex / normal.
//   icp_project: reduce.cu:285-299 -> global-frame vertex, linear index of the model pixel, in-bounds flag
__device__ __forceinline__ float3 icp_to_global(const IcpArgs & a, const float3 vcurr) { return a.Rcurr * vcurr + a.tcurr; }
__device__ __forceinline__ bool icp_project(const IcpArgs & a, const float3 vcurr, float3 & vcurr_g, int & o)
{
    vcurr_g = icp_to_global(a, vcurr);
    const float3 vcurr_cp = a.Rprev_inv * (vcurr_g - a.tprev);

    // vcurr_cp.x * fx / vcurr_cp.z + cx  (reduce.cu:295-296): one reciprocal shared by both components
    const ApproxDivisor dz = approx_divisor(vcurr_cp.z);
    const int ux = __float2int_rn(approx_div_add(__fmul_rn(vcurr_cp.x, a.fx), dz, a.cx));
    const int uy = __float2int_rn(approx_div_add(__fmul_rn(vcurr_cp.y, a.fy), dz, a.cy));

    o = uy * a.cols + ux;
    return !(ux < 0 || uy < 0 || ux >= a.cols || uy >= a.rows || vcurr_cp.z < 0);
}

//   icp_finish: reduce.cu:301-348 -> found_coresp and the row [n, s x n, n.(s-d)] (zeros when not found)
__device__ __forceinline__ bool icp_finish(const IcpArgs & a, const float3 vcurr_g, const float3 ncurr, const float3 vprev_g, const float3 nprev_g,
                                           float (&row)[7])
{
    const float3 ncurr_g = a.Rcurr * ncurr;

    const float dist = norm3(vprev_g - vcurr_g);
    const float sine = norm3(cross3(ncurr_g, nprev_g));

    const bool found = (sine < a.angleThres && dist <= a.distThres && !isnan(ncurr.x) && !isnan(nprev_g.x));
    // branch-free: the row is computed regardless and zeroed by selects (a divergent branch here costs more than the ~45
    // instructions it would skip in the lanes that miss)
    const float3 s_cp = a.Rprev_inv * (vcurr_g - a.tprev);
    const float3 d_cp = a.Rprev_inv * (vprev_g - a.tprev);
    const float3 n_cp = a.Rprev_inv * nprev_g;
    const float3 sxn = cross3(s_cp, n_cp);
    row[0] = found ? n_cp.x : 0.f;
    row[1] = found ? n_cp.y : 0.f;
    row[2] = found ? n_cp.z : 0.f;
    row[3] = found ? sxn.x : 0.f;
    row[4] = found ? sxn.y : 0.f;
    row[5] = found ? sxn.z : 0.f;
    row[6] = found ? dot3(n_cp, s_cp - d_cp) : 0.f;
    return found;
}

// vcurr / ncurr: this pixel's current-frame vertex and normal (already loaded).
// Returns found_coresp and fills row[7] (zeros when not found).
__device__ __forceinline__ bool icp_pixel(const IcpArgs & a, const float3 vcurr, const float3 ncurr, float (&row)[7])
{
#pragma unroll
    for(int k = 0; k < 7; k++) row[k] = 0.f;
    float3 vcurr_g;
    int o;
    if(!icp_project(a, vcurr, vcurr_g, o)) return false;
    const int plane = a.rows * a.cols;
    float3 vprev_g, nprev_g;
    vprev_g.x = __ldg(a.vprev + o);
    vprev_g.y = __ldg(a.vprev + o + plane);
    vprev_g.z = __ldg(a.vprev + o + 2 * plane);
    nprev_g.x = __ldg(a.nprev + o);
    nprev_g.y = __ldg(a.nprev + o + plane);
    nprev_g.z = __ldg(a.nprev + o + 2 * plane);
    return icp_finish(a, vcurr_g, ncurr, vprev_g, nprev_g, row);
}

// acc[0..26] += upper triangle of row^T row (7x7, row-major, without gg), acc[27] += gg
// (the residual), acc[28] += inlier.   Field order of JtJJtrSE3, cuda/types.cuh:79-92.
__device__ __forceinline__ void accumulate_se3(float (&acc)[29], const float (&row)[7], bool found)
{
    int k = 0;
#pragma unroll
    for(int i = 0; i < 7; i++)
#pragma unroll
        for(int j = i; j < 7; j++)
        {
            acc[k] += row[i] * row[j];
            k++;
        }
    acc[28] += found ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------
// RGB residual: photometric data association.   reduce.cu:768-841
// ------------------------------------------------------------------------------------
struct __align__(16) Corres   // byte-compatible with DataTerm (cuda/types.cuh:71-77): short2 zero, short2 one, float diff, bool valid
{
    short zx, zy;
    short ox, oy;
    float diff;
    int valid;
};

struct ResidualArgs
{
    float minScale;
    const short * dIdx;   // [rows][cols]
    const short * dIdy;
    const float * lastDepth;
    const float * nextDepth;
    const unsigned char * lastImage;
    const unsigned char * nextImage;
    float maxDepthDelta;
    float3 kt;
    Mat3 krkinv;
    int cols, rows;
};

// Pose-independent part of the test (reduce.cu:780-807): border, 4x4 all-nonzero window of
// nextImage, gradient magnitude, finite next depth.
__device__ __forceinline__ bool rgb_candidate(const ResidualArgs & a, int j0, int i)
{
    if(!(j0 < a.cols - 5 && i < a.rows - 1)) return false;
    bool valid = true;
    for(int u = max(i - 2, 0); u < min(i + 2, a.rows); u++)
        for(int v = max(j0 - 2, 0); v < min(j0 + 2, a.cols); v++) valid = valid && (a.nextImage[u * a.cols + v] > 0);
    if(!valid) return false;
    const int valx = a.dIdx[i * a.cols + j0];
    const int valy = a.dIdy[i * a.cols + j0];
    const float mTwo = (valx * valx) + (valy * valy);
    if(!(mTwo >= a.minScale)) return false;
    return !isnan(a.nextDepth[i * a.cols + j0]);
}

// Pose-dependent part (reduce.cu:809-831), again split around the gather:
//   rgb_project: warp pixel (x, y) with depth d1 into the last image -> (u0, v0), warped depth, in-bounds flag
__device__ __forceinline__ bool rgb_project(const ResidualArgs & a, int x, int y, float d1, int & u0, int & v0, float & transformed_d1)
{
    // d1 * (k.x * x + k.y * y + k.z) + kt, the reference's roundings (residualKernel SASS): k.y*y rounded, k.x*x fused,
    // + k.z added, then one FFMA with d1; the two quotients share one approximate reciprocal.
    const float xf = (float)x, yf = (float)y;
    const float s2 = __fadd_rn(__fmaf_rn(xf, a.krkinv.r2.x, __fmul_rn(yf, a.krkinv.r2.y)), a.krkinv.r2.z);
    const float s0 = __fadd_rn(__fmaf_rn(xf, a.krkinv.r0.x, __fmul_rn(yf, a.krkinv.r0.y)), a.krkinv.r0.z);
    const float s1 = __fadd_rn(__fmaf_rn(xf, a.krkinv.r1.x, __fmul_rn(yf, a.krkinv.r1.y)), a.krkinv.r1.z);
    transformed_d1 = __fmaf_rn(d1, s2, a.kt.z);
    const ApproxDivisor dz = approx_divisor(transformed_d1);
    u0 = __float2int_rn(approx_div(__fmaf_rn(d1, s0, a.kt.x), dz));
    v0 = __float2int_rn(approx_div(__fmaf_rn(d1, s1, a.kt.y), dz));
    return u0 >= 0 && v0 >= 0 && u0 < a.cols && v0 < a.rows;
}
//   rgb_accept: the gathered last depth / intensity pass the gates
__device__ __forceinline__ bool rgb_accept(const ResidualArgs & a, float transformed_d1, float d0, unsigned char l)
{
    return d0 > 0 && fabsf(transformed_d1 - d0) <= a.maxDepthDelta && l != 0;
}

// Returns validity; fills c (zero = pixel in the last image, one = this pixel, diff = next - last intensity).
__device__ __forceinline__ bool rgb_associate(const ResidualArgs & a, int x, int y, Corres & c)
{
    const float d1 = a.nextDepth[y * a.cols + x];
    int u0, v0;
    float transformed_d1;
    if(rgb_project(a, x, y, d1, u0, v0, transformed_d1))
    {
        const float d0 = __ldg(a.lastDepth + v0 * a.cols + u0);
        const unsigned char l = __ldg(a.lastImage + v0 * a.cols + u0);
        if(rgb_accept(a, transformed_d1, d0, l))
        {
            c.zx = (short)u0;
            c.zy = (short)v0;
            c.ox = (short)x;
            c.oy = (short)y;
            c.diff = __fsub_rn(static_cast<float>(a.nextImage[y * a.cols + x]), static_cast<float>(l));
            c.valid = 1;
            return true;
        }
    }
    return false;
}

// applyKernel, utils.cu:582-606: 3x3 derivative with a running kernelIndex from 8 downwards over the CLIPPED window
// (border taps misalign, on purpose), float -> short truncation.
__device__ __forceinline__ void derivative_pixel(const unsigned char * src, int rows, int cols, int x, int y, short & dx, short & dy)
{
    const float gx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float gy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    float dxVal = 0;
    float dyVal = 0;
    if(x >= 1 && y >= 1 && x < cols - 1 && y < rows - 1)
    {
        // interior: the nine taps are loaded first, then accumulated in the reference's order (kernelIndex 8 -> 0)
        const unsigned char * p = src + (y - 1) * cols + (x - 1);
        float v[9];
#pragma unroll
        for(int r = 0; r < 3; r++)
#pragma unroll
            for(int c = 0; c < 3; c++) v[r * 3 + c] = (float)p[r * cols + c];
        const float fgx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
        const float fgy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
#pragma unroll
        for(int t = 0; t < 9; t++)
        {
            dxVal = __fmaf_rn(v[t], fgx[8 - t], dxVal);
            dyVal = __fmaf_rn(v[t], fgy[8 - t], dyVal);
        }
        dx = (short)dxVal;
        dy = (short)dyVal;
        return;
    }
    int kernelIndex = 8;
    for(int j = max(y - 1, 0); j <= min(y + 1, rows - 1); j++)
        for(int i = max(x - 1, 0); i <= min(x + 1, cols - 1); i++)
        {
            const float s = (float)src[j * cols + i];
            dxVal = __fmaf_rn(s, gx[kernelIndex], dxVal);
            dyVal = __fmaf_rn(s, gy[kernelIndex], dyVal);
            --kernelIndex;
        }
    dx = (short)dxVal;
    dy = (short)dyVal;
}


// rgb_candidate with the derivatives computed on the spot from nextImage instead of read from dIdx/dIdy
// (bit-identical values; lets the persistent kernel run without the derivative images).
__device__ __forceinline__ bool rgb_candidate_derive(const ResidualArgs & a, int j0, int i, short & gx, short & gy)
{
    gx = gy = 0;
    if(!(j0 < a.cols - 5 && i < a.rows - 1)) return false;
    bool valid = true;
    for(int u = max(i - 2, 0); u < min(i + 2, a.rows); u++)
        for(int v = max(j0 - 2, 0); v < min(j0 + 2, a.cols); v++) valid = valid && (a.nextImage[u * a.cols + v] > 0);
    if(!valid) return false;
    derivative_pixel(a.nextImage, a.rows, a.cols, j0, i, gx, gy);
    const int valx = gx, valy = gy;
    const float mTwo = (valx * valx) + (valy * valy);
    if(!(mTwo >= a.minScale)) return false;
    return !isnan(a.nextDepth[i * a.cols + j0]);
}

// ------------------------------------------------------------------------------------
// RGB step: photometric Jacobian row.   reduce.cu:512-558
// ------------------------------------------------------------------------------------
struct RgbStepArgs
{
    float sigma;
    float fx, fy;          // intr(level)
    float sobelScale;
    int cols, rows;
    const short * dIdx;
    const short * dIdy;
    // the reference reads pointClouds[level] (projectToPointCloud, utils.cu:640-658); we
    // re-derive the same point from lastDepth with the same arithmetic instead of storing it
    const float * lastDepth;
    float invFx, invFy, cx, cy;   // 1.0f/intr(level).fx ... as computed at utils.cu:670
    const float * cloud;   // optional explicit cloud (operator API); nullptr => derive from lastDepth
};

__device__ __forceinline__ float3 cloud_point(const float * depth, int cols, int x, int y, float invFx, float invFy, float cx, float cy)
{
    const float z = __ldg(depth + y * cols + x);
    float3 p;
    p.x = __fmul_rn(__fmul_rn((x - cx), z), invFx);   // (float)((x - cx) * z * invFx), utils.cu:655
    p.y = __fmul_rn(__fmul_rn((y - cy), z), invFy);
    p.z = z;
    return p;
}

// `float invz = 1.0 / cloudPoint.z` (reduce.cu:541) is an fp64 division rounded to float in the reference.
// The correctly rounded fp32 reciprocal is the same number except when the fp64 quotient falls within
// 2^-29 (relative) of a float rounding boundary, and it only scales a Jacobian row (sums: 1e-4 contract),
// so the IEEE fp32 reciprocal is used: ~10x shorter dependency chain than the fp64 divide.
__device__ __forceinline__ float rgb_invz(float z) { return __frcp_rn(z); }

__device__ __forceinline__ void rgb_row(const RgbStepArgs & a, const Corres & c, float (&row)[7])
{
#define SLAM_FLT_EPSILON ((float)1.19209290E-07F)
    float w = a.sigma + fabsf(c.diff);
    w = w > SLAM_FLT_EPSILON ? 1.0f / w : 1.0f;
    if(a.sigma == -1) w = 1;

    row[6] = -w * c.diff;

    float3 cloudPoint;
    if(a.cloud)
    {
        const float * p = a.cloud + 3 * (c.zy * a.cols + c.zx);
        cloudPoint = make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
    else
        cloudPoint = cloud_point(a.lastDepth, a.cols, c.zx, c.zy, a.invFx, a.invFy, a.cx, a.cy);

    const float invz = rgb_invz(cloudPoint.z);
    const float dI_dx_val = w * a.sobelScale * __ldg(a.dIdx + c.oy * a.cols + c.ox);
    const float dI_dy_val = w * a.sobelScale * __ldg(a.dIdy + c.oy * a.cols + c.ox);
    const float v0 = dI_dx_val * a.fx * invz;
    const float v1 = dI_dy_val * a.fy * invz;
    const float v2 = -(v0 * cloudPoint.x + v1 * cloudPoint.y) * invz;

    row[0] = v0;
    row[1] = v1;
    row[2] = v2;
    row[3] = -cloudPoint.z * v1 + cloudPoint.y * v2;
    row[4] = cloudPoint.z * v0 - cloudPoint.x * v2;
    row[5] = -cloudPoint.y * v0 + cloudPoint.x * v1;
}

// Same row from values the caller already holds in registers: z = lastDepth at (zx, zy) (the reference's
// cloud point is (x - cx) * z * invFx, ..., utils.cu:655), the two gradients at the pixel itself, diff.
__device__ __forceinline__ void rgb_row_regs(const RgbStepArgs & a, int zx, int zy, float z, short gx, short gy, float diff, float (&row)[7])
{
    float w = a.sigma + fabsf(diff);
    w = w > SLAM_FLT_EPSILON ? 1.0f / w : 1.0f;
    if(a.sigma == -1) w = 1;
    row[6] = -w * diff;
    float3 cloudPoint;
    cloudPoint.x = __fmul_rn(__fmul_rn((zx - a.cx), z), a.invFx);
    cloudPoint.y = __fmul_rn(__fmul_rn((zy - a.cy), z), a.invFy);
    cloudPoint.z = z;
    const float invz = rgb_invz(cloudPoint.z);
    const float dI_dx_val = w * a.sobelScale * gx;
    const float dI_dy_val = w * a.sobelScale * gy;
    const float v0 = dI_dx_val * a.fx * invz;
    const float v1 = dI_dy_val * a.fy * invz;
    const float v2 = -(v0 * cloudPoint.x + v1 * cloudPoint.y) * invz;
    row[0] = v0;
    row[1] = v1;
    row[2] = v2;
    row[3] = -cloudPoint.z * v1 + cloudPoint.y * v2;
    row[4] = cloudPoint.z * v0 - cloudPoint.x * v2;
    row[5] = -cloudPoint.y * v0 + cloudPoint.x * v1;
}

// ------------------------------------------------------------------------------------
// SO3 pre-alignment row.   reduce.cu:953-1054
// ------------------------------------------------------------------------------------
struct So3Args
{
    const unsigned char * lastImage;
    const unsigned char * nextImage;
    Mat3 imageBasis, kinv, krlr;
    int cols, rows;
};

__device__ __forceinline__ float2 so3_gradient(const unsigned char * img, int cols, int x, int y)
{
    float2 g;
    const float actu = static_cast<float>(img[y * cols + x]);
    float back = static_cast<float>(img[y * cols + x - 1]);
    float fore = static_cast<float>(img[y * cols + x + 1]);
    g.x = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = static_cast<float>(img[(y - 1) * cols + x]);
    fore = static_cast<float>(img[(y + 1) * cols + x]);
    g.y = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    return g;
}

__device__ __forceinline__ bool so3_pixel(const So3Args & a, int x, int y, float (&row)[4])
{
    row[0] = row[1] = row[2] = row[3] = 0.f;

    const float3 unwarped = make_float3((float)x, (float)y, 1.0f);
    // imageBasis * (x, y, 1) with the roundings of the reference's so3Kernel SASS.  The z = 1 term folds to
    // an add; rows 1 and 2 round the y product and fuse the x product, row 0 is the other way round.
    float3 warped;
    warped.x = __fadd_rn(__fmaf_rn(unwarped.y, a.imageBasis.r0.y, __fmul_rn(unwarped.x, a.imageBasis.r0.x)), a.imageBasis.r0.z);
    warped.y = __fadd_rn(__fmaf_rn(unwarped.x, a.imageBasis.r1.x, __fmul_rn(unwarped.y, a.imageBasis.r1.y)), a.imageBasis.r1.z);
    warped.z = __fadd_rn(__fmaf_rn(unwarped.x, a.imageBasis.r2.x, __fmul_rn(unwarped.y, a.imageBasis.r2.y)), a.imageBasis.r2.z);
    const ApproxDivisor dz = approx_divisor(warped.z);
    const int wx = __float2int_rn(approx_div(warped.x, dz));
    const int wy = __float2int_rn(approx_div(warped.y, dz));

    const bool found = wx >= 1 && wx < a.cols - 1 && wy >= 1 && wy < a.rows - 1 && x >= 1 && x < a.cols - 1 && y >= 1 && y < a.rows - 1;
    if(!found) return false;

    const float2 gradNext = so3_gradient(a.nextImage, a.cols, wx, wy);
    const float2 gradLast = so3_gradient(a.lastImage, a.cols, x, y);

    const float gx = (gradNext.x + gradLast.x) / 2.0f;
    const float gy = (gradNext.y + gradLast.y) / 2.0f;

    const float3 point = a.kinv * unwarped;
    const float z2 = point.z * point.z;

    const float ka = a.krlr.r0.x, kb = a.krlr.r0.y, kc = a.krlr.r0.z;
    const float kd = a.krlr.r1.x, ke = a.krlr.r1.y, kf = a.krlr.r1.z;
    const float kg = a.krlr.r2.x, kh = a.krlr.r2.y, ki = a.krlr.r2.z;

    const float3 leftProduct = make_float3(((point.z * (kd * gy + ka * gx)) - (gy * kg * y) - (gx * kg * x)) / z2,
                                           ((point.z * (ke * gy + kb * gx)) - (gy * kh * y) - (gx * kh * x)) / z2,
                                           ((point.z * (kf * gy + kc * gx)) - (gy * ki * y) - (gx * ki * x)) / z2);
    const float3 jac = cross3(leftProduct, point);

    row[0] = jac.x;
    row[1] = jac.y;
    row[2] = jac.z;
    row[3] = -(static_cast<float>(a.nextImage[wy * a.cols + wx]) - static_cast<float>(a.lastImage[y * a.cols + x]));
    return true;
}

// Field order of JtJJtrSO3, cuda/types.cuh:138-147.
__device__ __forceinline__ void accumulate_so3(float (&acc)[11], const float (&row)[4], bool found)
{
    int k = 0;
#pragma unroll
    for(int i = 0; i < 4; i++)
#pragma unroll
        for(int j = i; j < 4; j++)
        {
            acc[k] += row[i] * row[j];
            k++;
        }
    acc[10] += found ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------
// Reductions: warp shuffle -> shared memory -> one partial per block -> the block that
// takes the last ticket folds all partials in a fixed order (deterministic), replacing
// the reference's second kernel (reduceSum<<<1,512>>>, reduce.cu:167-185).
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Transposed ("reduce-scatter") warp reduction: 32 values per lane in, lane L ends with the warp-wide
// sum of value L in v[0].  31 shuffles instead of 32 x 5, fixed summation pattern (deterministic).
template <int W>
__device__ __forceinline__ void warp_rs_step(float (&v)[32], const int lane)
{
    const bool upper = (lane & W) != 0;
#pragma unroll
    for(int i = 0; i < W; i++)
    {
        const float send = upper ? v[i] : v[i + W];
        const float keep = upper ? v[i + W] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, W);
    }
}
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32])
{
    const int lane = threadIdx.x & 31;
    warp_rs_step<16>(v, lane);
    warp_rs_step<8>(v, lane);
    warp_rs_step<4>(v, lane);
    warp_rs_step<2>(v, lane);
    warp_rs_step<1>(v, lane);
    return v[0];
}

// Block-wide sum of NV values per thread. blockDim.x must be a multiple of 32 and <= 1024.
// After the call threads 0..NV-1 of warp 0 hold total[t] in the return value.
template <typename T, int NV>
__device__ __forceinline__ T block_sum(T (&acc)[NV], T * smem /* [32][NV] */)
{
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
#pragma unroll
    for(int k = 0; k < NV; k++)
    {
        const T s = warp_sum(acc[k]);
        if(lane == 0) smem[wid * NV + k] = s;
    }
    __syncthreads();
    T total = 0;
    if(threadIdx.x < NV)
        for(int w = 0; w < nw; w++) total += smem[w * NV + threadIdx.x];
    __syncthreads();
    return total;
}

}   // namespace slam
