/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's RGB-D tracking path.
 *
 * Plain C + OpenMP restatement of siw-engineering/slam's RGBDOdometryef hot path, each function
 * citing the reference file:line it follows (paths relative to /root/reference/src).  It exists to
 *   (1) cross-check the CUDA path from an independent implementation (tests/, smoke()), and
 *   (2) be the host-cores baseline of bench.py (cpu_baseline, --impl reference): the reference has
 *       no CPU tracking path, so "a host-compiled build of the reference's reduction math" is this.
 * Only tests/, __graft_entry__.smoke() and bench.py may load it.  The product never does.
 *
 * Parity status: PINNED against the reference's own CUDA kernels run on a B200 -- golden vectors
 * under tests/golden/ (made by tests/golden/make_golden.py from oracle/_ref) -- with these caveats:
 * the reference is built with --prec-div=false --prec-sqrt=false --ftz=true and FMA contraction
 * (src/CMakeLists.txt:115-116), i.e. approximate reciprocal / rsqrt and fused multiply-adds whose
 * last-bit behaviour a CPU cannot reproduce.  This port uses IEEE fp32 (-ffp-contract=off), so
 * integer images agree bit-for-bit except where a quotient lands within 1 ulp of a truncation
 * boundary, float maps agree to ~1 ulp, masks agree except for boundary pixels, and reduced sums
 * agree to ~1e-5.  The bit-exact oracle for the GPU tests is oracle/_ref (the reference kernels).
 *
 * The 6x6 / 3x3 solves here are Gaussian elimination with partial pivoting -- deliberately NOT the
 * LDL^T of slam_b200/csrc/small_math.hpp -- so the two implementations check each other.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NUM_PYRS 3
static const uint32_t QNAN_BITS = 0x7fffffffu; /* utils.cu:130 */
static float qnan(void)
{
    float f;
    memcpy(&f, &QNAN_BITS, 4);
    return f;
}

typedef struct { float fx, fy, cx, cy; } intr_t;
static intr_t intr_level(intr_t k, int level) /* sensors/Camera.h:14-18 */
{
    const int div = 1 << level;
    intr_t r = {k.fx / div, k.fy / div, k.cx / div, k.cy / div};
    return r;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ prep (odom/utils.cu) */
/* utils.cu:57-94 pyrDownGaussKernel */
void oracle_pyr_down(const uint16_t * src, int srows, int scols, uint16_t * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const float sigma_color = 30;
    const float weights[] = {0.375f, 0.25f, 0.0625f};
#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            const int D = 5;
            const int center = src[(2 * y) * scols + 2 * x];
            const int x_mi = (0 > 2 * x - D / 2 ? 0 : 2 * x - D / 2) - 2 * x;
            const int y_mi = (0 > 2 * y - D / 2 ? 0 : 2 * y - D / 2) - 2 * y;
            const int x_ma = (scols < 2 * x - D / 2 + D ? scols : 2 * x - D / 2 + D) - 2 * x;
            const int y_ma = (srows < 2 * y - D / 2 + D ? srows : 2 * y - D / 2 + D) - 2 * y;
            float sum = 0, wall = 0;
            for(int yi = y_mi; yi < y_ma; ++yi)
                for(int xi = x_mi; xi < x_ma; ++xi)
                {
                    const int val = src[(2 * y + yi) * scols + 2 * x + xi];
                    if(abs(val - center) < 3 * sigma_color)
                    {
                        sum += val * weights[abs(xi)] * weights[abs(yi)];
                        wall += weights[abs(xi)] * weights[abs(yi)];
                    }
                }
            dst[y * dcols + x] = (uint16_t)(int)(sum / wall);
        }
}

/* utils.cu:109-133 computeVmapKernel (invalid: only the x plane becomes NaN) */
void oracle_create_vmap(intr_t k, const uint16_t * depth, int rows, int cols, float * vmap, float cutoff)
{
    const float fx_inv = 1.f / k.fx, fy_inv = 1.f / k.fy;
    const int plane = rows * cols;
#pragma omp parallel for schedule(static)
    for(int v = 0; v < rows; v++)
        for(int u = 0; u < cols; u++)
        {
            const float z = depth[v * cols + u] / 1000.f;
            if(z != 0 && z < cutoff)
            {
                vmap[v * cols + u] = z * (u - k.cx) * fx_inv;
                vmap[plane + v * cols + u] = z * (v - k.cy) * fy_inv;
                vmap[2 * plane + v * cols + u] = z;
            }
            else
                vmap[v * cols + u] = qnan();
        }
}

static void normalize3(float * x, float * y, float * z) /* cuda/operators.cuh:82-86 */
{
    const float rn = 1.0f / sqrtf(*x * *x + *y * *y + *z * *z);
    *x *= rn; *y *= rn; *z *= rn;
}

/* utils.cu:151-188 computeNmapKernel */
void oracle_create_nmap(const float * vmap, int rows, int cols, float * nmap)
{
    const int plane = rows * cols;
#pragma omp parallel for schedule(static)
    for(int v = 0; v < rows; v++)
        for(int u = 0; u < cols; u++)
        {
            const int o = v * cols + u;
            if(u == cols - 1 || v == rows - 1) { nmap[o] = qnan(); continue; }
            const float x00 = vmap[o], x01 = vmap[o + 1], x10 = vmap[o + cols];
            if(!isnan(x00) && !isnan(x01) && !isnan(x10))
            {
                const float ax = x01 - x00, ay = vmap[plane + o + 1] - vmap[plane + o], az = vmap[2 * plane + o + 1] - vmap[2 * plane + o];
                const float bx = x10 - x00, by = vmap[plane + o + cols] - vmap[plane + o], bz = vmap[2 * plane + o + cols] - vmap[2 * plane + o];
                float rx = ay * bz - az * by, ry = az * bx - ax * bz, rz = ax * by - ay * bx;
                normalize3(&rx, &ry, &rz);
                nmap[o] = rx; nmap[plane + o] = ry; nmap[2 * plane + o] = rz;
            }
            else
                nmap[o] = qnan();
        }
}

/* utils.cu:270-310 copyMapsKernel */
void oracle_copy_maps(const float * v4, const float * n4, int rows, int cols, float * vdst, float * ndst)
{
    const int plane = rows * cols;
#pragma omp parallel for schedule(static)
    for(int o = 0; o < plane; o++)
    {
        const float * vs = v4 + 4 * o;
        const float * ns = n4 + 4 * o;
        const int ok = !(vs[2] == 0);
        for(int c = 0; c < 3; c++)
        {
            vdst[c * plane + o] = ok ? vs[c] : qnan();
            ndst[c * plane + o] = ok ? ns[c] : qnan();
        }
    }
}

/* utils.cu:365-416 resizeMapKernel */
void oracle_resize_map(const float * src, int srows, int scols, float * dst, int normalize)
{
    const int drows = srows / 2, dcols = scols / 2, sp = srows * scols, dp = drows * dcols;
#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            const int s = (2 * y) * scols + 2 * x, o = y * dcols + x;
            const float x00 = src[s], x01 = src[s + 1], x10 = src[s + scols], x11 = src[s + scols + 1];
            if(isnan(x00) || isnan(x01) || isnan(x10) || isnan(x11)) { dst[o] = qnan(); continue; }
            float n[3];
            for(int c = 0; c < 3; c++) n[c] = (src[c * sp + s] + src[c * sp + s + 1] + src[c * sp + s + scols] + src[c * sp + s + scols + 1]) / 4;
            if(normalize) normalize3(&n[0], &n[1], &n[2]);
            for(int c = 0; c < 3; c++) dst[c * dp + o] = n[c];
        }
}

/* utils.cu:206-248 tranformMapsKernel (in place allowed) */
void oracle_transform_maps(float * vmap, float * nmap, int rows, int cols, const float * R, const float * t)
{
    const int plane = rows * cols;
#pragma omp parallel for schedule(static)
    for(int o = 0; o < plane; o++)
    {
        const float vx = vmap[o];
        if(!isnan(vx))
        {
            const float vy = vmap[plane + o], vz = vmap[2 * plane + o];
            vmap[o] = R[0] * vx + R[1] * vy + R[2] * vz + t[0];
            vmap[plane + o] = R[3] * vx + R[4] * vy + R[5] * vz + t[1];
            vmap[2 * plane + o] = R[6] * vx + R[7] * vy + R[8] * vz + t[2];
        }
        const float nx = nmap[o];
        if(!isnan(nx))
        {
            const float ny = nmap[plane + o], nz = nmap[2 * plane + o];
            nmap[o] = R[0] * nx + R[1] * ny + R[2] * nz;
            nmap[plane + o] = R[3] * nx + R[4] * ny + R[5] * nz;
            nmap[2 * plane + o] = R[6] * nx + R[7] * ny + R[8] * nz;
        }
    }
}

/* utils.cu:526-537 verticesToDepthKernel */
void oracle_vertices_to_depth(const float * v4, int n, float * dst, float cutoff)
{
#pragma omp parallel for schedule(static)
    for(int i = 0; i < n; i++)
    {
        const float z = v4[4 * i + 2];
        dst[i] = (z > cutoff || z <= 0) ? qnan() : z;
    }
}

static const float GAUSS5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

/* utils.cu:332-363 pyrDownKernelGaussF (window stops one short of the last row/col; int count) */
void oracle_pyr_down_gauss_f(const float * src, int srows, int scols, float * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            const int D = 5;
            const int tx = (2 * x - D / 2 + D < scols - 1) ? 2 * x - D / 2 + D : scols - 1;
            const int ty = (2 * y - D / 2 + D < srows - 1) ? 2 * y - D / 2 + D : srows - 1;
            float sum = 0;
            int count = 0;
            for(int cy = (0 > 2 * y - D / 2 ? 0 : 2 * y - D / 2); cy < ty; ++cy)
                for(int cx = (0 > 2 * x - D / 2 ? 0 : 2 * x - D / 2); cx < tx; ++cx)
                {
                    const float s = src[cy * scols + cx];
                    if(!isnan(s))
                    {
                        const float w = GAUSS5[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum += s * w;
                        count += w;
                    }
                }
            dst[y * dcols + x] = (float)(sum / (float)count);
        }
}

/* utils.cu:470-500 pyrDownKernelIntensityGauss */
void oracle_pyr_down_gauss_u8(const uint8_t * src, int srows, int scols, uint8_t * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
#pragma omp parallel for schedule(static)
    for(int y = 0; y < drows; y++)
        for(int x = 0; x < dcols; x++)
        {
            const int D = 5;
            const int tx = (2 * x - D / 2 + D < scols - 1) ? 2 * x - D / 2 + D : scols - 1;
            const int ty = (2 * y - D / 2 + D < srows - 1) ? 2 * y - D / 2 + D : srows - 1;
            float sum = 0;
            int count = 0;
            for(int cy = (0 > 2 * y - D / 2 ? 0 : 2 * y - D / 2); cy < ty; ++cy)
                for(int cx = (0 > 2 * x - D / 2 ? 0 : 2 * x - D / 2); cx < tx; ++cx)
                {
                    const uint8_t s = src[cy * scols + cx];
                    if(s > 0)
                    {
                        const float w = GAUSS5[(ty - cy - 1) * 5 + (tx - cx - 1)];
                        sum += s * w;
                        count += w;
                    }
                }
            const float q = sum / (float)count;
            dst[y * dcols + x] = isnan(q) ? 0 : (uint8_t)q; /* cvt of NaN is 0 on the GPU */
        }
}

/* utils.cu:550-563 bgr2IntensityKernel */
void oracle_bgr_to_intensity(const uint8_t * rgba, int n, uint8_t * dst)
{
#pragma omp parallel for schedule(static)
    for(int i = 0; i < n; i++)
    {
        const int value = (float)rgba[4 * i] * 0.114f + (float)rgba[4 * i + 1] * 0.299f + (float)rgba[4 * i + 2] * 0.587f;
        dst[i] = (uint8_t)value;
    }
}

/* utils.cu:582-606 applyKernel (running kernelIndex over the clipped window) */
void oracle_derivatives(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy)
{
    const float gx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float gy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            float dxVal = 0, dyVal = 0;
            int kernelIndex = 8;
            for(int j = (y - 1 > 0 ? y - 1 : 0); j <= (y + 1 < rows - 1 ? y + 1 : rows - 1); j++)
                for(int i = (x - 1 > 0 ? x - 1 : 0); i <= (x + 1 < cols - 1 ? x + 1 : cols - 1); i++)
                {
                    dxVal += (float)src[j * cols + i] * gx[kernelIndex];
                    dyVal += (float)src[j * cols + i] * gy[kernelIndex];
                    --kernelIndex;
                }
            dx[y * cols + x] = (int16_t)dxVal;
            dy[y * cols + x] = (int16_t)dyVal;
        }
}

/* utils.cu:640-658 projectPointsKernel */
void oracle_project_points(const float * depth, int rows, int cols, float * cloud3, intr_t k)
{
    const float invFx = 1.0f / k.fx, invFy = 1.0f / k.fy;
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            const float z = depth[y * cols + x];
            float * c = cloud3 + 3 * (y * cols + x);
            c[0] = (float)((x - k.cx) * z * invFx);
            c[1] = (float)((y - k.cy) * z * invFy);
            c[2] = z;
        }
}

/* ------------------------------------------------------------------ reductions (odom/reduce.cu) */
static void mat3_vec(const float * m, const float * v, float * o)
{
    o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    o[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    o[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
static void cross(const float * a, const float * b, float * o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static float norm(const float * a) { return sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

static void accumulate(float * acc, const float * row, int n) /* JtJJtrSE3 / JtJJtrSO3 field order, cuda/types.cuh:79-147 */
{
    int k = 0;
    for(int i = 0; i < n; i++)
        for(int j = i; j < n; j++) acc[k++] += row[i] * row[j];
}

/* reduce.cu:282-416 ICPReduction; out29 in JtJJtrSE3 order (27 products, residual, inliers) */
void oracle_icp_step(const float * Rcurr, const float * tcurr, const float * vcurr, const float * ncurr, const float * Rprev_inv, const float * tprev,
                     intr_t k, const float * vprev, const float * nprev, float distThres, float angleThres, int rows, int cols, float * out29,
                     uint8_t * mask_or_null)
{
    const int plane = rows * cols;
    int nt = oracle_num_threads();
    float * part = (float *)calloc((size_t)nt * 32, sizeof(float));
#pragma omp parallel
    {
#ifdef _OPENMP
        float * acc = part + 32 * omp_get_thread_num();
#else
        float * acc = part;
#endif
#pragma omp for schedule(static)
        for(int i = 0; i < plane; i++)
        {
            float row[7] = {0, 0, 0, 0, 0, 0, 0};
            int found = 0;
            const float vc[3] = {vcurr[i], vcurr[plane + i], vcurr[2 * plane + i]};
            float vg[3], d[3], vcp[3];
            mat3_vec(Rcurr, vc, vg);
            for(int c = 0; c < 3; c++) { vg[c] += tcurr[c]; d[c] = vg[c] - tprev[c]; }
            mat3_vec(Rprev_inv, d, vcp);
            /* __float2int_rn: round half to even = lrintf in the default rounding mode; NaN -> 0 */
            const float px = vcp[0] * k.fx / vcp[2] + k.cx, py = vcp[1] * k.fy / vcp[2] + k.cy;
            const long ux = isnan(px) ? 0 : (fabsf(px) < 2e9f ? lrintf(px) : (px > 0 ? 2147483647L : -2147483648L));
            const long uy = isnan(py) ? 0 : (fabsf(py) < 2e9f ? lrintf(py) : (py > 0 ? 2147483647L : -2147483648L));
            if(!(ux < 0 || uy < 0 || ux >= cols || uy >= rows || vcp[2] < 0))
            {
                const int o = (int)uy * cols + (int)ux;
                const float vp[3] = {vprev[o], vprev[plane + o], vprev[2 * plane + o]};
                const float np[3] = {nprev[o], nprev[plane + o], nprev[2 * plane + o]};
                const float nc[3] = {ncurr[i], ncurr[plane + i], ncurr[2 * plane + i]};
                float ng[3], dd[3], cr[3];
                mat3_vec(Rcurr, nc, ng);
                for(int c = 0; c < 3; c++) dd[c] = vp[c] - vg[c];
                cross(ng, np, cr);
                const float dist = norm(dd), sine = norm(cr);
                if(sine < angleThres && dist <= distThres && !isnan(nc[0]) && !isnan(np[0]))
                {
                    found = 1;
                    float s[3], dq[3], dcp[3], ncp[3], sxn[3], diff[3];
                    mat3_vec(Rprev_inv, d, s);
                    for(int c = 0; c < 3; c++) dq[c] = vp[c] - tprev[c];
                    mat3_vec(Rprev_inv, dq, dcp);
                    mat3_vec(Rprev_inv, np, ncp);
                    cross(s, ncp, sxn);
                    for(int c = 0; c < 3; c++) diff[c] = s[c] - dcp[c];
                    row[0] = ncp[0]; row[1] = ncp[1]; row[2] = ncp[2];
                    row[3] = sxn[0]; row[4] = sxn[1]; row[5] = sxn[2];
                    row[6] = ncp[0] * diff[0] + ncp[1] * diff[1] + ncp[2] * diff[2];
                }
            }
            if(found)
            {
                accumulate(acc, row, 7);
                acc[28] += 1.f;
            }
            if(mask_or_null) mask_or_null[i] = (uint8_t)found;
        }
    }
    for(int k2 = 0; k2 < 29; k2++)
    {
        float s = 0;
        for(int t = 0; t < nt; t++) s += part[32 * t + k2];
        out29[k2] = s;
    }
    free(part);
}

typedef struct { int16_t zx, zy, ox, oy; float diff; int32_t valid; } corres_t; /* DataTerm, cuda/types.cuh:71-77 */

/* reduce.cu:768-867 RGBResidual */
void oracle_rgb_residual(float minScale, const int16_t * dIdx, const int16_t * dIdy, const float * lastDepth, const float * nextDepth,
                         const uint8_t * lastImage, const uint8_t * nextImage, corres_t * corres, float maxDepthDelta, const float * kt, const float * krk,
                         int rows, int cols, int * count_sigma)
{
    long count = 0;
    int sigma = 0; /* int32 wrap-around semantics, as the reference */
    const int N = rows * cols;
    unsigned usig = 0;
#pragma omp parallel for schedule(static) reduction(+ : count, usig)
    for(int kk = 0; kk < N; kk++)
    {
        const int i = kk / cols, j0 = kk - i * cols;
        corres_t c;
        memset(&c, 0, sizeof(c));
        if(j0 < cols - 5 && i < rows - 1)
        {
            int valid = 1;
            for(int u = (i - 2 > 0 ? i - 2 : 0); u < (i + 2 < rows ? i + 2 : rows); u++)
                for(int v = (j0 - 2 > 0 ? j0 - 2 : 0); v < (j0 + 2 < cols ? j0 + 2 : cols); v++) valid = valid && (nextImage[u * cols + v] > 0);
            if(valid)
            {
                const int valx = dIdx[kk], valy = dIdy[kk];
                const float mTwo = (valx * valx) + (valy * valy);
                if(mTwo >= minScale)
                {
                    const int y = i, x = j0;
                    const float d1 = nextDepth[kk];
                    if(!isnan(d1))
                    {
                        const float td1 = (float)(d1 * (krk[6] * x + krk[7] * y + krk[8]) + kt[2]);
                        const float fu = (d1 * (krk[0] * x + krk[1] * y + krk[2]) + kt[0]) / td1;
                        const float fv = (d1 * (krk[3] * x + krk[4] * y + krk[5]) + kt[1]) / td1;
                        const long u0 = isnan(fu) ? 0 : (fabsf(fu) < 2e9f ? lrintf(fu) : -1);
                        const long v0 = isnan(fv) ? 0 : (fabsf(fv) < 2e9f ? lrintf(fv) : -1);
                        if(u0 >= 0 && v0 >= 0 && u0 < cols && v0 < rows)
                        {
                            const float d0 = lastDepth[v0 * cols + u0];
                            if(d0 > 0 && fabsf(td1 - d0) <= maxDepthDelta && lastImage[v0 * cols + u0] != 0)
                            {
                                c.zx = (int16_t)u0; c.zy = (int16_t)v0; c.ox = (int16_t)x; c.oy = (int16_t)y;
                                c.diff = (float)nextImage[kk] - (float)lastImage[v0 * cols + u0];
                                c.valid = 1;
                                count += 1;
                                usig += (unsigned)(int)(c.diff * c.diff);
                            }
                        }
                    }
                }
            }
        }
        corres[kk] = c;
    }
    sigma = (int)usig;
    count_sigma[0] = (int)count;
    count_sigma[1] = sigma;
}

/* reduce.cu:512-624 RGBReduction */
void oracle_rgb_step(const corres_t * corres, float sigma, const float * cloud3, float fx, float fy, const int16_t * dIdx, const int16_t * dIdy,
                     float sobelScale, int rows, int cols, float * out29)
{
    const int N = rows * cols;
    int nt = oracle_num_threads();
    float * part = (float *)calloc((size_t)nt * 32, sizeof(float));
#pragma omp parallel
    {
#ifdef _OPENMP
        float * acc = part + 32 * omp_get_thread_num();
#else
        float * acc = part;
#endif
#pragma omp for schedule(static)
        for(int i = 0; i < N; i++)
        {
            const corres_t * c = &corres[i];
            if(!c->valid) continue;
            float row[7];
            float w = sigma + fabsf(c->diff);
            w = w > FLT_EPSILON ? 1.0f / w : 1.0f;
            if(sigma == -1) w = 1;
            row[6] = -w * c->diff;
            const float * cp = cloud3 + 3 * (c->zy * cols + c->zx);
            const float invz = 1.0 / cp[2];
            const float dI_dx_val = w * sobelScale * dIdx[c->oy * cols + c->ox];
            const float dI_dy_val = w * sobelScale * dIdy[c->oy * cols + c->ox];
            const float v0 = dI_dx_val * fx * invz;
            const float v1 = dI_dy_val * fy * invz;
            const float v2 = -(v0 * cp[0] + v1 * cp[1]) * invz;
            row[0] = v0; row[1] = v1; row[2] = v2;
            row[3] = -cp[2] * v1 + cp[1] * v2;
            row[4] = cp[2] * v0 - cp[0] * v2;
            row[5] = -cp[1] * v0 + cp[0] * v1;
            accumulate(acc, row, 7);
            acc[28] += 1.f;
        }
    }
    for(int k2 = 0; k2 < 29; k2++)
    {
        float s = 0;
        for(int t = 0; t < nt; t++) s += part[32 * t + k2];
        out29[k2] = s;
    }
    free(part);
}

static void so3_gradient(const uint8_t * img, int cols, int x, int y, float * gx, float * gy) /* reduce.cu:953-969 */
{
    const float actu = img[y * cols + x];
    float back = img[y * cols + x - 1], fore = img[y * cols + x + 1];
    *gx = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
    back = img[(y - 1) * cols + x];
    fore = img[(y + 1) * cols + x];
    *gy = ((back + actu) / 2.0f) - ((fore + actu) / 2.0f);
}

/* reduce.cu:971-1080 SO3Reduction; out11 in JtJJtrSO3 order */
void oracle_so3_step(const uint8_t * lastImage, const uint8_t * nextImage, const float * H, const float * kinv, const float * krlr, int rows, int cols,
                     float * out11)
{
    const int N = rows * cols;
    int nt = oracle_num_threads();
    float * part = (float *)calloc((size_t)nt * 16, sizeof(float));
#pragma omp parallel
    {
#ifdef _OPENMP
        float * acc = part + 16 * omp_get_thread_num();
#else
        float * acc = part;
#endif
#pragma omp for schedule(static)
        for(int kk = 0; kk < N; kk++)
        {
            const int y = kk / cols, x = kk - y * cols;
            const float p[3] = {(float)x, (float)y, 1.0f};
            float wp[3];
            mat3_vec(H, p, wp);
            const float fx_ = wp[0] / wp[2], fy_ = wp[1] / wp[2];
            const long wx = isnan(fx_) ? 0 : (fabsf(fx_) < 2e9f ? lrintf(fx_) : -1), wy = isnan(fy_) ? 0 : (fabsf(fy_) < 2e9f ? lrintf(fy_) : -1);
            if(!(wx >= 1 && wx < cols - 1 && wy >= 1 && wy < rows - 1 && x >= 1 && x < cols - 1 && y >= 1 && y < rows - 1)) continue;
            float gnx, gny, glx, gly;
            so3_gradient(nextImage, cols, (int)wx, (int)wy, &gnx, &gny);
            so3_gradient(lastImage, cols, x, y, &glx, &gly);
            const float gx = (gnx + glx) / 2.0f, gy = (gny + gly) / 2.0f;
            float pt[3];
            mat3_vec(kinv, p, pt);
            const float z2 = pt[2] * pt[2];
            const float a = krlr[0], b = krlr[1], c = krlr[2], d = krlr[3], e = krlr[4], f = krlr[5], g = krlr[6], h = krlr[7], i = krlr[8];
            const float lp[3] = {((pt[2] * (d * gy + a * gx)) - (gy * g * y) - (gx * g * x)) / z2, ((pt[2] * (e * gy + b * gx)) - (gy * h * y) - (gx * h * x)) / z2,
                                 ((pt[2] * (f * gy + c * gx)) - (gy * i * y) - (gx * i * x)) / z2};
            float jr[3];
            cross(lp, pt, jr);
            const float row[4] = {jr[0], jr[1], jr[2], -((float)nextImage[wy * cols + wx] - (float)lastImage[y * cols + x])};
            accumulate(acc, row, 4);
            acc[10] += 1.f;
        }
    }
    for(int k2 = 0; k2 < 11; k2++)
    {
        float s = 0;
        for(int t = 0; t < nt; t++) s += part[16 * t + k2];
        out11[k2] = s;
    }
    free(part);
}

/* ------------------------------------------------------------------ small algebra (independent of small_math.hpp) */
static void mm(const double * a, const double * b, double * c, int n)
{
    double r[16];
    for(int i = 0; i < n; i++)
        for(int j = 0; j < n; j++)
        {
            double s = 0;
            for(int k = 0; k < n; k++) s += a[i * n + k] * b[k * n + j];
            r[i * n + j] = s;
        }
    memcpy(c, r, sizeof(double) * n * n);
}

/* Gauss-Jordan inverse with partial pivoting (n <= 6); returns 0 if singular */
static int invert(const double * A, double * out, int n)
{
    double m[36], inv[36];
    for(int i = 0; i < n * n; i++) { m[i] = A[i]; inv[i] = (i / n == i % n) ? 1.0 : 0.0; }
    for(int k = 0; k < n; k++)
    {
        int p = k;
        for(int i = k + 1; i < n; i++)
            if(fabs(m[i * n + k]) > fabs(m[p * n + k])) p = i;
        if(m[p * n + k] == 0.0) return 0;
        if(p != k)
            for(int j = 0; j < n; j++)
            {
                double t = m[k * n + j]; m[k * n + j] = m[p * n + j]; m[p * n + j] = t;
                t = inv[k * n + j]; inv[k * n + j] = inv[p * n + j]; inv[p * n + j] = t;
            }
        const double d = 1.0 / m[k * n + k];
        for(int j = 0; j < n; j++) { m[k * n + j] *= d; inv[k * n + j] *= d; }
        for(int i = 0; i < n; i++)
            if(i != k)
            {
                const double f = m[i * n + k];
                for(int j = 0; j < n; j++) { m[i * n + j] -= f * m[k * n + j]; inv[i * n + j] -= f * inv[k * n + j]; }
            }
    }
    memcpy(out, inv, sizeof(double) * n * n);
    return 1;
}

/* solve A x = b for symmetric positive semi-definite A (stands in for Eigen's ldlt().solve()); 0 when singular */
static void solve_sym(const double * A, const double * b, double * x, int n)
{
    double inv[36];
    if(!invert(A, inv, n)) { for(int i = 0; i < n; i++) x[i] = 0; return; }
    for(int i = 0; i < n; i++)
    {
        double s = 0;
        for(int j = 0; j < n; j++) s += inv[i * n + j] * b[j];
        x[i] = s;
    }
}

static void rodrigues(const double * r, double * R) /* odom/utils.h:16-52 */
{
    for(int i = 0; i < 9; i++) R[i] = (i % 4 == 0);
    double rx = r[0], ry = r[1], rz = r[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if(theta >= DBL_EPSILON)
    {
        const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
        rx *= it; ry *= it; rz *= it;
        const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        const double rx_[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        for(int k = 0; k < 9; k++) R[k] = c * (k % 4 == 0) + c1 * rrt[k] + s * rx_[k];
    }
}

static void unpack_se3(const float * s, double * A, double * b) /* reduce.cu:472-486 */
{
    int shift = 0;
    for(int i = 0; i < 6; ++i)
        for(int j = i; j < 7; ++j)
        {
            const float v = s[shift++];
            if(j == 6) b[i] = v; else A[j * 6 + i] = A[i * 6 + j] = v;
        }
}

/* ------------------------------------------------------------------ the tracker (odom/RGBDOdometryef.cpp) */
typedef struct
{
    int width, height;
    intr_t intr;
    float distThres, angleThres;
    uint16_t * depth[NUM_PYRS];
    float *vcurr[NUM_PYRS], *ncurr[NUM_PYRS], *vprev[NUM_PYRS], *nprev[NUM_PYRS];
    float *lastDepth[NUM_PYRS], *nextDepth[NUM_PYRS], *cloud[NUM_PYRS];
    uint8_t *lastImage[NUM_PYRS], *nextImage[NUM_PYRS], *lastNextImage[NUM_PYRS];
    int16_t *dIdx[NUM_PYRS], *dIdy[NUM_PYRS];
    corres_t * corres[NUM_PYRS];
    float *vtmp, *ntmp;
    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];
    int so3_iterations, gn_iterations;
} oracle_t;

void * oracle_create(int width, int height, float cx, float cy, float fx, float fy, float distThresh, float angleThresh)
{
    oracle_t * o = (oracle_t *)calloc(1, sizeof(oracle_t));
    o->width = width; o->height = height;
    o->intr.cx = cx; o->intr.cy = cy; o->intr.fx = fx; o->intr.fy = fy;
    o->distThres = distThresh ? distThresh : 0.10f;
    o->angleThres = angleThresh ? angleThresh : sinf(20.f * 3.14159254f / 180.f);
    for(int l = 0; l < NUM_PYRS; l++)
    {
        const size_t n = (size_t)(width >> l) * (height >> l);
        o->depth[l] = calloc(n, 2);
        o->vcurr[l] = calloc(n * 3, 4); o->ncurr[l] = calloc(n * 3, 4); o->vprev[l] = calloc(n * 3, 4); o->nprev[l] = calloc(n * 3, 4);
        o->lastDepth[l] = calloc(n, 4); o->nextDepth[l] = calloc(n, 4); o->cloud[l] = calloc(n * 3, 4);
        o->lastImage[l] = calloc(n, 1); o->nextImage[l] = calloc(n, 1); o->lastNextImage[l] = calloc(n, 1);
        o->dIdx[l] = calloc(n, 2); o->dIdy[l] = calloc(n, 2);
        o->corres[l] = calloc(n, sizeof(corres_t));
    }
    o->vtmp = calloc((size_t)width * height * 4, 4);
    o->ntmp = calloc((size_t)width * height * 4, 4);
    o->lastICPCount = o->lastRGBCount = o->lastSO3Count = (float)(width * height); /* RGBDOdometryef.cpp:26-31 */
    return o;
}

void oracle_destroy(void * h)
{
    oracle_t * o = (oracle_t *)h;
    for(int l = 0; l < NUM_PYRS; l++)
    {
        free(o->depth[l]); free(o->vcurr[l]); free(o->ncurr[l]); free(o->vprev[l]); free(o->nprev[l]);
        free(o->lastDepth[l]); free(o->nextDepth[l]); free(o->cloud[l]); free(o->lastImage[l]); free(o->nextImage[l]); free(o->lastNextImage[l]);
        free(o->dIdx[l]); free(o->dIdy[l]); free(o->corres[l]);
    }
    free(o->vtmp); free(o->ntmp);
    free(o);
}

/* RGBDOdometryef.cpp:118-142 */
void oracle_init_icp_depth(void * h, const uint16_t * depth, float cutoff)
{
    oracle_t * o = (oracle_t *)h;
    memcpy(o->depth[0], depth, (size_t)o->width * o->height * 2);
    for(int i = 1; i < NUM_PYRS; ++i) oracle_pyr_down(o->depth[i - 1], o->height >> (i - 1), o->width >> (i - 1), o->depth[i]);
    for(int i = 0; i < NUM_PYRS; ++i)
    {
        oracle_create_vmap(intr_level(o->intr, i), o->depth[i], o->height >> i, o->width >> i, o->vcurr[i], cutoff);
        oracle_create_nmap(o->vcurr[i], o->height >> i, o->width >> i, o->ncurr[i]);
    }
}

static void maps_from_textures(oracle_t * o, const float * v4, const float * n4, float ** vdst, float ** ndst)
{
    memcpy(o->vtmp, v4, (size_t)o->width * o->height * 16);
    memcpy(o->ntmp, n4, (size_t)o->width * o->height * 16);
    oracle_copy_maps(o->vtmp, o->ntmp, o->height, o->width, vdst[0], ndst[0]);
    for(int i = 1; i < NUM_PYRS; ++i)
    {
        oracle_resize_map(vdst[i - 1], o->height >> (i - 1), o->width >> (i - 1), vdst[i], 0);
        oracle_resize_map(ndst[i - 1], o->height >> (i - 1), o->width >> (i - 1), ndst[i], 1);
    }
}
/* RGBDOdometryef.cpp:144-167 */
void oracle_init_icp_maps(void * h, const float * v4, const float * n4, float cutoff)
{
    (void)cutoff;
    oracle_t * o = (oracle_t *)h;
    maps_from_textures(o, v4, n4, o->vcurr, o->ncurr);
}
/* RGBDOdometryef.cpp:169-206 */
void oracle_init_icp_model(void * h, const float * v4, const float * n4, float cutoff, const float * pose16)
{
    (void)cutoff;
    oracle_t * o = (oracle_t *)h;
    maps_from_textures(o, v4, n4, o->vprev, o->nprev);
    const float R[9] = {pose16[0], pose16[1], pose16[2], pose16[4], pose16[5], pose16[6], pose16[8], pose16[9], pose16[10]};
    const float t[3] = {pose16[3], pose16[7], pose16[11]};
    for(int i = 0; i < NUM_PYRS; ++i) oracle_transform_maps(o->vprev[i], o->nprev[i], o->height >> i, o->width >> i, R, t);
}
/* RGBDOdometryef.cpp:208-235 */
static void populate(oracle_t * o, const uint8_t * rgba, float ** dd, uint8_t ** di)
{
    if(dd)
    {
        oracle_vertices_to_depth(o->vtmp, o->width * o->height, dd[0], 6.0f);
        for(int i = 0; i + 1 < NUM_PYRS; i++) oracle_pyr_down_gauss_f(dd[i], o->height >> i, o->width >> i, dd[i + 1]);
    }
    oracle_bgr_to_intensity(rgba, o->width * o->height, di[0]);
    for(int i = 0; i + 1 < NUM_PYRS; i++) oracle_pyr_down_gauss_u8(di[i], o->height >> i, o->width >> i, di[i + 1]);
}
void oracle_init_rgb(void * h, const uint8_t * rgba) { oracle_t * o = (oracle_t *)h; populate(o, rgba, o->nextDepth, o->nextImage); }
void oracle_init_rgb_model(void * h, const uint8_t * rgba) { oracle_t * o = (oracle_t *)h; populate(o, rgba, o->lastDepth, o->lastImage); }
void oracle_init_first_rgb(void * h, const uint8_t * rgba) { oracle_t * o = (oracle_t *)h; populate(o, rgba, NULL, o->lastNextImage); }

static void kmat(intr_t k, double * K)
{
    memset(K, 0, 72);
    K[0] = k.fx; K[4] = k.fy; K[2] = k.cx; K[5] = k.cy; K[8] = 1;
}

/* RGBDOdometryef.cpp:267-595 */
void oracle_get_incremental_transformation(void * h, float * trans, float * rot, int rgbOnly, float icpWeight, int pyramid, int fastOdom, int so3)
{
    oracle_t * o = (oracle_t *)h;
    const int icp = !rgbOnly && icpWeight > 0;
    const int rgb = rgbOnly || icpWeight < 100;
    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, 36); memcpy(tprev, trans, 12); memcpy(Rcurr, rot, 36); memcpy(tcurr, trans, 12);
    o->so3_iterations = o->gn_iterations = 0;

    if(rgb)
        for(int i = 0; i < NUM_PYRS; i++) oracle_derivatives(o->nextImage[i], o->height >> i, o->width >> i, o->dIdx[i], o->dIdy[i]);

    double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if(so3)
    {
        const int L = 2;
        float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        double K[9], Kinv[9];
        kmat(intr_level(o->intr, L), K);
        invert(K, Kinv, 3);
        float lastError = FLT_MAX / 2, lastCount = FLT_MAX / 2;
        double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for(int i = 0; i < 10; i++)
        {
            double KR[9], H[9];
            mm(K, resultR, KR, 3);
            mm(KR, Kinv, H, 3);
            float Hf[9], Kif[9], KRf[9], s[11];
            for(int k = 0; k < 9; k++) { Hf[k] = (float)H[k]; Kif[k] = (float)Kinv[k]; KRf[k] = (float)KR[k]; }
            oracle_so3_step(o->lastNextImage[L], o->nextImage[L], Hf, Kif, KRf, o->height >> L, o->width >> L, s);
            o->so3_iterations++;
            o->lastSO3Error = sqrtf(s[9]) / s[10];
            o->lastSO3Count = s[10];
            if(o->lastSO3Error < lastError && lastCount == o->lastSO3Count) break;
            else if((double)o->lastSO3Error > (double)lastError + 0.001)
            {
                o->lastSO3Error = lastError; o->lastSO3Count = lastCount;
                memcpy(resultR, lastResultR, 72);
                break;
            }
            lastError = o->lastSO3Error; lastCount = o->lastSO3Count;
            memcpy(lastResultR, resultR, 72);
            /* 3x3 float system of the reference solved in double here, cast back to float */
            double A[9], b[3], x[3];
            int shift = 0;
            for(int r = 0; r < 3; ++r)
                for(int c = r; c < 4; ++c)
                {
                    const float v = s[shift++];
                    if(c == 3) b[r] = v; else A[c * 3 + r] = A[r * 3 + c] = v;
                }
            solve_sym(A, b, x, 3);
            const double dd[3] = {(float)x[0], (float)x[1], (float)x[2]};
            double ru[9];
            rodrigues(dd, ru);
            float nr[9];
            for(int r = 0; r < 3; r++)
                for(int c = 0; c < 3; c++) nr[r * 3 + c] = (float)ru[r * 3 + 0] * R_lr[0 * 3 + c] + (float)ru[r * 3 + 1] * R_lr[1 * 3 + c] + (float)ru[r * 3 + 2] * R_lr[2 * 3 + c];
            memcpy(R_lr, nr, 36);
            for(int k = 0; k < 9; k++) resultR[k] = R_lr[k];
        }
    }

    int iterations[NUM_PYRS] = {fastOdom ? 3 : 10, pyramid ? 5 : 0, pyramid ? 4 : 0};
    double Rp[9], Rpi[9];
    for(int k = 0; k < 9; k++) Rp[k] = Rprev[k];
    invert(Rp, Rpi, 3);
    float Rprev_inv[9];
    for(int k = 0; k < 9; k++) Rprev_inv[k] = (float)Rpi[k];

    double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if(so3)
        for(int x = 0; x < 3; x++)
            for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

    const float minGrad[NUM_PYRS] = {5, 3, 1};
    const float sobelScale = 1.0f / 8.0f;

    for(int i = NUM_PYRS - 1; i >= 0; i--)
    {
        const int rows = o->height >> i, cols = o->width >> i;
        const intr_t ki = intr_level(o->intr, i);
        if(rgb) oracle_project_points(o->lastDepth[i], rows, cols, o->cloud[i], ki);
        double K[9], Kinv[9];
        kmat(ki, K);
        invert(K, Kinv, 3);
        o->lastRGBError = FLT_MAX;
        for(int j = 0; j < iterations[i]; j++)
        {
            double Rt[16], R[9], KR[9], KRK[9];
            invert(resultRt, Rt, 4);
            for(int x = 0; x < 3; x++)
                for(int y = 0; y < 3; y++) R[x * 3 + y] = Rt[x * 4 + y];
            mm(K, R, KR, 3);
            mm(KR, Kinv, KRK, 3);
            float krk[9], kt[3];
            for(int k = 0; k < 9; k++) krk[k] = (float)KRK[k];
            for(int x = 0; x < 3; x++) kt[x] = (float)(K[x * 3] * Rt[3] + K[x * 3 + 1] * Rt[7] + K[x * 3 + 2] * Rt[11]);

            int cs[2] = {0, 0};
            if(rgb)
                oracle_rgb_residual((float)(pow(minGrad[i], 2.0) / pow(sobelScale, 2.0)), o->dIdx[i], o->dIdy[i], o->lastDepth[i], o->nextDepth[i], o->lastImage[i],
                                    o->nextImage[i], o->corres[i], 0.07f, kt, krk, rows, cols, cs);
            const int rgbSize = cs[0], sigma = cs[1];
            float sigmaVal = sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize);
            const float rgbError = sqrt(sigma) / (rgbSize == 0 ? 1 : rgbSize);
            if(rgbOnly && rgbError > o->lastRGBError) break;
            o->lastRGBError = rgbError;
            o->lastRGBCount = rgbSize;
            if(rgbOnly) sigmaVal = -1;

            float s_icp[29] = {0}, s_rgb[29] = {0};
            if(icp)
            {
                oracle_icp_step(Rcurr, tcurr, o->vcurr[i], o->ncurr[i], Rprev_inv, tprev, ki, o->vprev[i], o->nprev[i], o->distThres, o->angleThres, rows, cols,
                                s_icp, NULL);
                o->lastICPError = sqrtf(s_icp[27]) / s_icp[28];
                o->lastICPCount = s_icp[28];
            }
            if(rgb) oracle_rgb_step(o->corres[i], sigmaVal, o->cloud[i], ki.fx, ki.fy, o->dIdx[i], o->dIdy[i], sobelScale, rows, cols, s_rgb);
            o->gn_iterations++;

            double A_icp[36] = {0}, b_icp[6] = {0}, A_rgb[36] = {0}, b_rgb[6] = {0};
            unpack_se3(s_icp, A_icp, b_icp);
            unpack_se3(s_rgb, A_rgb, b_rgb);
            if(icp && rgb)
            {
                const double w = icpWeight;
                for(int k = 0; k < 36; k++) o->lastA[k] = A_rgb[k] + w * w * A_icp[k];
                for(int k = 0; k < 6; k++) o->lastb[k] = b_rgb[k] + w * b_icp[k];
            }
            else if(icp) { memcpy(o->lastA, A_icp, 288); memcpy(o->lastb, b_icp, 48); }
            else { memcpy(o->lastA, A_rgb, 288); memcpy(o->lastb, b_rgb, 48); }
            double x[6];
            solve_sym(o->lastA, o->lastb, x, 6);

            /* computeUpdateSE3, odom/utils.h:54-74 */
            double Rr[9], U[16];
            rodrigues(x + 3, Rr);
            memset(U, 0, sizeof(U));
            for(int r = 0; r < 3; r++) { for(int c = 0; c < 3; c++) U[r * 4 + c] = Rr[r * 3 + c]; U[r * 4 + 3] = x[r]; }
            U[15] = 1;
            mm(U, resultRt, resultRt, 4);
            /* currentT = [Rprev|tprev] * rgbOdom^-1 (float isometry), RGBDOdometryef.cpp:563-575 */
            float Ro[9], to[3], Ri[9], ti[3];
            for(int r = 0; r < 3; r++) { for(int c = 0; c < 3; c++) Ro[r * 3 + c] = (float)resultRt[r * 4 + c]; to[r] = (float)resultRt[r * 4 + 3]; }
            for(int r = 0; r < 3; r++)
                for(int c = 0; c < 3; c++) Ri[r * 3 + c] = Ro[c * 3 + r];
            for(int r = 0; r < 3; r++) ti[r] = -(Ri[r * 3] * to[0] + Ri[r * 3 + 1] * to[1] + Ri[r * 3 + 2] * to[2]);
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) Rcurr[r * 3 + c] = Rprev[r * 3] * Ri[c] + Rprev[r * 3 + 1] * Ri[3 + c] + Rprev[r * 3 + 2] * Ri[6 + c];
                tcurr[r] = Rprev[r * 3] * ti[0] + Rprev[r * 3 + 1] * ti[1] + Rprev[r * 3 + 2] * ti[2] + tprev[r];
            }
        }
    }
    if(rgb)
    {
        const float dx = tcurr[0] - tprev[0], dy = tcurr[1] - tprev[1], dz = tcurr[2] - tprev[2];
        if(sqrtf(dx * dx + dy * dy + dz * dz) > 0.3) { memcpy(Rcurr, Rprev, 36); memcpy(tcurr, tprev, 12); }
    }
    if(so3)
        for(int i = 0; i < NUM_PYRS; i++) { uint8_t * t = o->lastNextImage[i]; o->lastNextImage[i] = o->nextImage[i]; o->nextImage[i] = t; }
    memcpy(trans, tcurr, 12);
    memcpy(rot, Rcurr, 36);
}

void oracle_get_stats(void * h, float * six, double * lastA36, double * lastb6, int * iters2)
{
    oracle_t * o = (oracle_t *)h;
    six[0] = o->lastICPError; six[1] = o->lastICPCount; six[2] = o->lastRGBError; six[3] = o->lastRGBCount; six[4] = o->lastSO3Error; six[5] = o->lastSO3Count;
    memcpy(lastA36, o->lastA, 288);
    memcpy(lastb6, o->lastb, 48);
    iters2[0] = o->so3_iterations; iters2[1] = o->gn_iterations;
}

/* buffer access for the tests: same tap ids as include/slam_odom.h */
const void * oracle_buffer(void * h, int tap, int level)
{
    oracle_t * o = (oracle_t *)h;
    switch(tap)
    {
        case 0: return o->depth[level];
        case 1: return o->vcurr[level];
        case 2: return o->ncurr[level];
        case 3: return o->vprev[level];
        case 4: return o->nprev[level];
        case 5: return o->lastDepth[level];
        case 6: return o->nextDepth[level];
        case 7: return o->lastImage[level];
        case 8: return o->nextImage[level];
        case 9: return o->lastNextImage[level];
        case 10: return o->dIdx[level];
        case 11: return o->dIdy[level];
        case 12: return o->cloud[level];
        case 13: return o->corres[level];
    }
    return NULL;
}
