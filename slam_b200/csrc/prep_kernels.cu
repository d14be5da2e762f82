// Image / map preparation kernels of the tracker for sm_100a.
//
// Behavioural contract = the twelve prep kernels of the reference's src/odom/utils.cu
// (cited per kernel), including their border and NaN quirks (SURVEY.md appendix A,
// items 19-24).  The kernels here are organised differently from the reference:
// fused per pyramid level (depth -> vertex+normal map + next depth level in one
// launch; float4 model maps -> planar global-frame maps of all three levels in one
// launch; depth+intensity pyramids side by side), dense row-major buffers instead of
// pitched ones, and no per-call cudaMalloc/cudaMemcpy/cudaDeviceSynchronize.
// The un-fused operator entry points (slam_op_*) are kept for operator-level parity.
#include <vector>
#include <cstdio>
#include "odom_internal.hpp"

namespace slam {

// ------------------------------------------------------------------ per-pixel pieces
// The window operators read their source through an accessor: a dense image in global memory, or a tile (core + halo) of it that a
// block has staged in shared memory -- same taps, same order, same arithmetic.
template <class T>
struct ImageSrc
{
    const T * p;
    int cols;
    __device__ __forceinline__ T at(int r, int c) const { return p[r * cols + c]; }
};
template <class T, int PITCH>
struct TileSrc
{
    const T * p;          // element (0, 0) of the IMAGE as seen through the tile: tile base - (r0 * PITCH + c0)
    __device__ __forceinline__ TileSrc(const T * tile, int r0, int c0) : p(tile - (r0 * PITCH + c0)) {}
    __device__ __forceinline__ T at(int r, int c) const { return p[r * PITCH + c]; }
};

// utils.cu:57-94  pyrDownGaussKernel (u16 depth, bilateral-gated 5x5, sigma_color = 30)
// MODE (all three 5x5 window operators below): 0 = the pixel decides (interior / clipped window), 1 = the caller guarantees an interior
// pixel, 2 = the caller guarantees a border pixel.  The tiled frame preparation sorts its pixels by window_interior() first, so that
// no warp runs both paths for the sake of the one border pixel in each of its rows.
__device__ __forceinline__ bool window_interior_u16(int srows, int scols, int x, int y)
{
    return 2 * x - 2 >= 0 && 2 * y - 2 >= 0 && 2 * x + 3 <= scols && 2 * y + 3 <= srows;
}
__device__ __forceinline__ bool window_interior_gauss(int srows, int scols, int x, int y)
{
    return x >= 1 && y >= 1 && 2 * x + 3 <= scols - 1 && 2 * y + 3 <= srows - 1;
}
template <int MODE = 0, class Src>
__device__ __forceinline__ unsigned short pyr_down_u16_at(const Src & src, int srows, int scols, int x, int y)
{
    const int D = 5;
    const float sigma_color = 30;
    const int center = src.at(2 * y, 2 * x);

    const int x_mi = max(0, 2 * x - D / 2) - 2 * x;
    const int y_mi = max(0, 2 * y - D / 2) - 2 * y;
    const int x_ma = min(scols, 2 * x - D / 2 + D) - 2 * x;
    const int y_ma = min(srows, 2 * y - D / 2 + D) - 2 * y;

    float sum = 0;
    float wall = 0;

    if(MODE != 2 && (MODE == 1 || (x_mi == -2 && y_mi == -2 && x_ma == 3 && y_ma == 3)))
    {
        // interior: the full 5x5 window; all 25 loads are issued before the first use.  Every product and partial sum
        // is exact in fp32 (16-bit values x multiples of 1/256), so only the final quotient rounds, as in the reference.
        int val[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++) val[r][c] = src.at(2 * y - 2 + r, 2 * x - 2 + c);
        // The reference accumulates val * w[c] * w[r] and w[c] * w[r] in fp32 with w = {1, 4, 6, 4, 1} / 16: every term and every
        // partial sum is a multiple of 1/256 below 2^16, hence exact -- so the same two sums are taken in integers (one IMAD and one
        // IADD per tap instead of a conversion, two multiplies and two adds) and scaled by 1/256 exactly; the quotient, the only
        // rounding, gets the identical operands.
        const int w5i[5] = {1, 4, 6, 4, 1};
        int isum = 0, iwall = 0;
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(abs(val[r][c] - center) < 90)   // 3 * sigma_color
                {
                    isum += val[r][c] * (w5i[c] * w5i[r]);
                    iwall += w5i[c] * w5i[r];
                }
        sum = (float)isum * 0.00390625f;
        wall = (float)iwall * 0.00390625f;
        return static_cast<unsigned short>(static_cast<int>(sum / wall));
    }

    // border: the clipped window with every load issued first (the sums are exact, see above, so the order is free); the border
    // threads used to walk their taps one dependent round trip at a time and set the duration of the whole launch
    // The sums are exact for any subset of the taps (see above): the same integer form, a tap outside the clipped window gets a value
    // that fails the colour gate (the code is cold -- a handful of warps per frame run it -- so its length is its cost).
    {
        int val[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
            {
                const int yi = r - 2, xi = c - 2;
                const bool in = yi >= y_mi && yi < y_ma && xi >= x_mi && xi < x_ma;
                val[r][c] = in ? (int)src.at(2 * y + yi, 2 * x + xi) : (1 << 20);
            }
        const int w5i[5] = {1, 4, 6, 4, 1};   // the reference's weights[abs(xi)] = {0.375, 0.25, 0.0625}, xi = c - 2, times 16
        int isum = 0, iwall = 0;
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(abs(val[r][c] - center) < 90)   // 3 * sigma_color
                {
                    isum += val[r][c] * (w5i[c] * w5i[r]);
                    iwall += w5i[c] * w5i[r];
                }
        sum = (float)isum * 0.00390625f;
        wall = (float)iwall * 0.00390625f;
    }
    return static_cast<unsigned short>(static_cast<int>(sum / wall));
}

__device__ __forceinline__ unsigned short pyr_down_u16_pixel(const unsigned short * src, int srows, int scols, int x, int y)
{
    return pyr_down_u16_at(ImageSrc<unsigned short>{src, scols}, srows, scols, x, y);
}

// utils.cu:109-133 computeVmapKernel.  Returns false (vertex invalid) or the vertex.
__device__ __forceinline__ bool vertex_from_depth(unsigned short d, int u, int v, float fx_inv, float fy_inv, float cx, float cy,
                                                  float depthCutoff, float3 & p)
{
    // `depth / 1000.f` is a multiply by the rounded constant 0.001f in the reference's SASS
    // (computeVmapKernel: FMUL.FTZ R, R, 0.0010000000474974513); written as such, with an explicit
    // rounding, so it can never be fused into the subtractions of the normal computation below
    const float z = __fmul_rn((float)d, 0.001f);
    if(z != 0 && z < depthCutoff)
    {
        // explicit _rn multiplies: same roundings as the reference's stores, and they keep
        // the compiler from fusing them into the normal's subtractions below
        p.x = __fmul_rn(__fmul_rn(z, (u - cx)), fx_inv);
        p.y = __fmul_rn(__fmul_rn(z, (v - cy)), fy_inv);
        p.z = z;
        return true;
    }
    return false;
}

// 5x5 {1,4,6,4,1}^2 pyramid tap loop shared by the float-depth and u8-intensity versions
// (utils.cu:332-363 and 470-500): window [max(0,2x-2), min(2x+3, cols-1)), weight index
// (ty-cy-1)*5+(tx-cx-1), integer `count`.
// the same table as a product of two {1,4,6,4,1} entries picked without a memory lookup (a, b in [0, 4])
__device__ __forceinline__ float gauss5_weight_rc(int a, int b)
{
    const float wa = (a == 0 || a == 4) ? 1.f : (a == 2 ? 6.f : 4.f);
    const float wb = (b == 0 || b == 4) ? 1.f : (b == 2 ? 6.f : 4.f);
    return wa * wb;
}

template <int MODE = 0, class Src>
__device__ __forceinline__ float pyr_down_gauss_f_at(const Src & src, int srows, int scols, int x, int y)
{
    const int D = 5;
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int cy = max(0, 2 * y - D / 2);
    float sum = 0;
    int count = 0;
    if(MODE != 2 && (MODE == 1 || (x >= 1 && y >= 1 && tx == 2 * x + 3 && ty == 2 * y + 3)))
    {
        // interior: full window, weight index (4-r)*5 + (4-c) = the symmetric {1,4,6,4,1}^2 table; loads first,
        // then the reference's accumulation order (rows outer, columns inner), one FMA per finite tap.
        float v[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++) v[r][c] = src.at(2 * y - 2 + r, 2 * x - 2 + c);
        const float w5[5] = {1.f, 4.f, 6.f, 4.f, 1.f};
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(!isnan(v[r][c]))
                {
                    const float w = w5[r] * w5[c];
                    sum = __fmaf_rn(v[r][c], w, sum);
                    count += (int)w;
                }
        return (float)(sum / (float)count);
    }
    // border: the clipped window of the reference's loop (rows [cy, ty), columns [cx0, tx), at most 5 x 5), same taps in the same
    // order with the same misaligned weights -- but every load is issued before the first is consumed: the border threads
    // used to walk their taps one dependent L2 round trip at a time and set the duration of the whole launch
    {
        const int cx0 = max(0, 2 * x - D / 2);
        const int nr = ty - cy, nc = tx - cx0;
        float v[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++) v[r][c] = (r < nr && c < nc) ? src.at(cy + r, cx0 + c) : SLAM_QNAN;
        // weight of tap (r, c) = table entry (nr - 1 - r, nc - 1 - c), i.e. index (ty-cy-1)*5 + (tx-cx-1): one factor per row and per
        // column, formed once (the products of {1, 4, 6} are exact, and so is the count as a float sum of them)
        float wr[5], wc[5];
#pragma unroll
        for(int k = 0; k < 5; k++)
        {
            const int a = nr - 1 - k, b = nc - 1 - k;
            wr[k] = (a == 0 || a == 4) ? 1.f : (a == 2 ? 6.f : 4.f);
            wc[k] = (b == 0 || b == 4) ? 1.f : (b == 2 ? 6.f : 4.f);
        }
        float countf = 0.f;
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(!isnan(v[r][c]))   // taps outside the clipped window were given NaN
                {
                    const float w = wr[r] * wc[c];
                    sum = __fmaf_rn(v[r][c], w, sum);
                    countf += w;
                }
        return (float)(sum / countf);
    }
    return (float)(sum / (float)count);
}

__device__ __forceinline__ float pyr_down_gauss_f_pixel(const float * src, int srows, int scols, int x, int y)
{
    return pyr_down_gauss_f_at(ImageSrc<float>{src, scols}, srows, scols, x, y);
}

template <int MODE = 0, class Src>
__device__ __forceinline__ unsigned char pyr_down_gauss_u8_at(const Src & src, int srows, int scols, int x, int y)
{
    const int D = 5;
    const int tx = min(2 * x - D / 2 + D, scols - 1);
    const int ty = min(2 * y - D / 2 + D, srows - 1);
    int cy = max(0, 2 * y - D / 2);
    float sum = 0;
    int count = 0;
    if(MODE != 2 && (MODE == 1 || (x >= 1 && y >= 1 && tx == 2 * x + 3 && ty == 2 * y + 3)))
    {
        // interior: full window; integer arithmetic is exact here (<= 255 * 256), only the quotient rounds
        int v[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++) v[r][c] = src.at(2 * y - 2 + r, 2 * x - 2 + c);
        const int w5[5] = {1, 4, 6, 4, 1};
        int isum = 0;
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(v[r][c] > 0)
                {
                    isum += v[r][c] * w5[r] * w5[c];
                    count += w5[r] * w5[c];
                }
        return (unsigned char)((float)isum / (float)count);
    }
    // border: as in the float version, all loads first; the sums are small integers, exact in fp32 in any order
    {
        const int cx0 = max(0, 2 * x - D / 2);
        const int nr = ty - cy, nc = tx - cx0;
        int v[5][5];
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++) v[r][c] = (r < nr && c < nc) ? (int)src.at(cy + r, cx0 + c) : 0;
        // one weight factor per row and per column (see the float version); all sums are integers below 2^16, exact either way
        int wr[5], wc[5];
#pragma unroll
        for(int k = 0; k < 5; k++)
        {
            const int a = nr - 1 - k, b = nc - 1 - k;
            wr[k] = (a == 0 || a == 4) ? 1 : (a == 2 ? 6 : 4);
            wc[k] = (b == 0 || b == 4) ? 1 : (b == 2 ? 6 : 4);
        }
        int isum = 0;
#pragma unroll
        for(int r = 0; r < 5; r++)
#pragma unroll
            for(int c = 0; c < 5; c++)
                if(v[r][c] > 0)
                {
                    isum += v[r][c] * (wr[r] * wc[c]);
                    count += wr[r] * wc[c];
                }
        return (unsigned char)((float)isum / (float)count);
    }
    return (unsigned char)(sum / (float)count);
}

__device__ __forceinline__ unsigned char pyr_down_gauss_u8_pixel(const unsigned char * src, int srows, int scols, int x, int y)
{
    return pyr_down_gauss_u8_at(ImageSrc<unsigned char>{src, scols}, srows, scols, x, y);
}

// utils.cu:550-563 bgr2IntensityKernel (c0,c1,c2 = first three bytes of the RGBA8 texel)
__device__ __forceinline__ unsigned char intensity_pixel(uchar4 src)
{
    // (float)c0*0.114f + (float)c1*0.299f + (float)c2*0.587f with the reference's roundings (middle product rounded)
    const int value = __fmaf_rn((float)src.z, 0.587f, __fmaf_rn((float)src.x, 0.114f, __fmul_rn((float)src.y, 0.299f)));
    return (unsigned char)value;
}

// utils.cu:526-537 verticesToDepthKernel
__device__ __forceinline__ float depth_from_vertex_z(float z, float cutOff) { return z > cutOff || z <= 0 ? SLAM_QNAN : z; }

// Batched launches: gridDim.y = number of sequences; every per-sequence buffer sits at a constant byte stride from
// sequence 0's (odom_api.cu: layout_sequence), and so do the caller's dense input stacks.
template <class T>
__device__ __forceinline__ T * seq_shift(T * p, size_t stride_bytes)
{
    return p ? (T *)((const char *)p + (size_t)blockIdx.y * stride_bytes) : p;
}

// createVMap + createNMap (utils.cu:109-188) of one pixel, depth read through an accessor
template <class Src>
__device__ __forceinline__ void vertex_normal_pixel(const Src & depth, int u, int v, int rows, int cols, float fx_inv, float fy_inv, float cx, float cy,
                                                    float depthCutoff, float * __restrict__ vmap, float * __restrict__ nmap)
{
    const int plane = rows * cols;
    const int o = v * cols + u;

    float3 v00;
    const bool ok00 = vertex_from_depth(depth.at(v, u), u, v, fx_inv, fy_inv, cx, cy, depthCutoff, v00);
    if(ok00)
    {
        vmap[o] = v00.x;
        vmap[o + plane] = v00.y;
        vmap[o + 2 * plane] = v00.z;
    }
    else
        vmap[o] = SLAM_QNAN;   // y,z planes keep stale data, as in the reference

    if(u == cols - 1 || v == rows - 1)
    {
        nmap[o] = SLAM_QNAN;
        return;
    }
    float3 v01, v10;
    const bool ok01 = vertex_from_depth(depth.at(v, u + 1), u + 1, v, fx_inv, fy_inv, cx, cy, depthCutoff, v01);
    const bool ok10 = vertex_from_depth(depth.at(v + 1, u), u, v + 1, fx_inv, fy_inv, cx, cy, depthCutoff, v10);
    if(ok00 && ok01 && ok10)
    {
        const float3 r = unit3(cross3(v01 - v00, v10 - v00));
        nmap[o] = r.x;
        nmap[o + plane] = r.y;
        nmap[o + 2 * plane] = r.z;
    }
    else
        nmap[o] = SLAM_QNAN;
}

// ------------------------------------------------------------------ fused depth level
// One launch per pyramid level l of the CURRENT frame:
//   blocks [0, nb_map)  : depth_l -> vmap_l, nmap_l  (createVMap + createNMap, utils.cu:109-188)
//   blocks [nb_map, ..) : depth_l -> depth_{l+1}     (pyrDown, utils.cu:57-94), if dst != nullptr
__global__ void __launch_bounds__(256) k_depth_level(const unsigned short * __restrict__ depth, int rows, int cols, float fx_inv, float fy_inv,
                                                     float cx, float cy, float depthCutoff, float * __restrict__ vmap, float * __restrict__ nmap,
                                                     unsigned short * __restrict__ next_depth, int nb_map_x, int nb_map, size_t in_stride, size_t out_stride)
{
    depth = seq_shift(depth, in_stride);
    vmap = seq_shift(vmap, out_stride);
    nmap = seq_shift(nmap, out_stride);
    next_depth = seq_shift(next_depth, out_stride);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if((int)blockIdx.x < nb_map)
    {
        const int bx = blockIdx.x % nb_map_x, by = blockIdx.x / nb_map_x;
        const int u = bx * 32 + tx, v = by * 8 + ty;
        if(u >= cols || v >= rows) return;
        vertex_normal_pixel(ImageSrc<unsigned short>{depth, cols}, u, v, rows, cols, fx_inv, fy_inv, cx, cy, depthCutoff, vmap, nmap);
    }
    else
    {
        const int drows = rows / 2, dcols = cols / 2;
        const int nbx = div_up(dcols, 32);
        const int b = blockIdx.x - nb_map;
        const int x = (b % nbx) * 32 + tx, y = (b / nbx) * 8 + ty;
        if(x >= dcols || y >= drows) return;
        next_depth[y * dcols + x] = pyr_down_u16_pixel(depth, rows, cols, x, y);
    }
}

// ------------------------------------------------------------------ fused model maps
// initICPModel / initICP(maps): RGBA32F vertex+normal (camera frame) -> planar maps of all
// levels, optionally moved to the global frame.  One thread owns a 4x4 block of level-0
// pixels = 2x2 of level 1 = 1 of level 2 (copyMaps utils.cu:270-310, resizeMap :365-416,
// tranformMaps :206-248, verticesToDepth :526-537 for the RGB path's depth source).
struct ModelMapsArgs
{
    const float4 * vsrc;
    const float4 * nsrc;
    int rows, cols;        // level 0
    int levels;            // 1..3 handled here
    float * vdst[3];
    float * ndst[3];
    int transform;         // 0: keep camera frame (initICP maps overload), 1: apply R,t
    Mat3 R;
    float3 t;
    float * depth_tmp;     // [rows][cols] z of the vertex map with the maxDepthRGB cut, or nullptr
    float depth_cut;
    float * vcam2;         // optional camera-frame copies of level 2 (source of a 4th level), or nullptr
    float * ncam2;
    // batched launch (gridDim.y sequences): byte strides of the inputs / of the arena buffers, per-sequence [R | t] (12 floats each)
    size_t in_stride, out_stride;
    const float * poses12;
    // frame-level front end: level 0 of populateRGBDData for the model AND the current frame in the same pass
    // (RGBDOdometryef.cpp:208-235: depth[0] = z of the vertices after the maxDepthRGB cut, image[0] = intensity of the RGBA8 texel)
    float * lastDepth0;
    float * nextDepth0;
    const uchar4 * model_rgba;
    const uchar4 * rgba;
    unsigned char * lastImage0;
    unsigned char * nextImage0;
    size_t rgba_stride;
};

__device__ __forceinline__ void store_map_pixel(float * vdst, float * ndst, int plane, int o, float3 v, float3 n, bool transform, const Mat3 & R,
                                                const float3 & t)
{
    // tranformMapsKernel semantics: NaN in x => only the x plane is (re)written
    if(transform)
    {
        if(!isnan(v.x))
        {
            const float3 d = R * v + t;
            vdst[o] = d.x;
            vdst[o + plane] = d.y;
            vdst[o + 2 * plane] = d.z;
        }
        else
            vdst[o] = SLAM_QNAN;
        if(!isnan(n.x))
        {
            const float3 d = R * n;
            ndst[o] = d.x;
            ndst[o + plane] = d.y;
            ndst[o + 2 * plane] = d.z;
        }
        else
            ndst[o] = SLAM_QNAN;
    }
    else
    {
        vdst[o] = v.x;
        ndst[o] = n.x;
        if(!isnan(v.x))
        {
            vdst[o + plane] = v.y;
            vdst[o + 2 * plane] = v.z;
        }
        if(!isnan(n.x))
        {
            ndst[o + plane] = n.y;
            ndst[o + 2 * plane] = n.z;
        }
    }
}

// 2x2 average with the reference's NaN rule (any NaN x => result x NaN, y/z untouched).
template <bool normalize>
__device__ __forceinline__ float3 resize4(const float3 & a00, const float3 & a01, const float3 & a10, const float3 & a11)
{
    float3 n;
    if(isnan(a00.x) || isnan(a01.x) || isnan(a10.x) || isnan(a11.x))
    {
        n.x = SLAM_QNAN;
        n.y = n.z = 0.f;
        return n;
    }
    n.x = (a00.x + a01.x + a10.x + a11.x) / 4;
    n.y = (a00.y + a01.y + a10.y + a11.y) / 4;
    n.z = (a00.z + a01.z + a10.z + a11.z) / 4;
    if(normalize) n = unit3(n);
    return n;
}

__global__ void __launch_bounds__(128) k_model_maps(const ModelMapsArgs a0)
{
    ModelMapsArgs a = a0;
    if(gridDim.y > 1 || a.poses12)
    {
        a.vsrc = seq_shift(a.vsrc, a.in_stride);
        a.nsrc = seq_shift(a.nsrc, a.in_stride);
        for(int l = 0; l < 3; l++)
        {
            a.vdst[l] = seq_shift(a.vdst[l], a.out_stride);
            a.ndst[l] = seq_shift(a.ndst[l], a.out_stride);
        }
        a.depth_tmp = seq_shift(a.depth_tmp, a.out_stride);
        a.lastDepth0 = seq_shift(a.lastDepth0, a.out_stride);
        a.nextDepth0 = seq_shift(a.nextDepth0, a.out_stride);
        a.lastImage0 = seq_shift(a.lastImage0, a.out_stride);
        a.nextImage0 = seq_shift(a.nextImage0, a.out_stride);
        a.model_rgba = seq_shift(a.model_rgba, a.rgba_stride);
        a.rgba = seq_shift(a.rgba, a.rgba_stride);
        a.vcam2 = seq_shift(a.vcam2, a.out_stride);
        a.ncam2 = seq_shift(a.ncam2, a.out_stride);
        if(a.poses12)
        {
            const float * q = a.poses12 + 12 * blockIdx.y;
            a.R.r0 = make_float3(__ldg(q + 0), __ldg(q + 1), __ldg(q + 2));
            a.R.r1 = make_float3(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5));
            a.R.r2 = make_float3(__ldg(q + 6), __ldg(q + 7), __ldg(q + 8));
            a.t = make_float3(__ldg(q + 9), __ldg(q + 10), __ldg(q + 11));
        }
    }
    // Four consecutive lanes own one 4x4 pixel block (the unit that folds into one level-2 pixel), one 2x2 quarter each: quarter
    // q = lane & 3 sits at (2 (q & 1), 2 (q >> 1)) inside the block; consecutive blocks walk along x.  A lane loads and stores its
    // four level-0 pixels, folds them into its level-1 pixel, and the four level-1 pixels of a block meet through shuffles for the
    // level-2 pixel (same operands in the same order as one thread per 4x4 block, which this replaces: 4x the threads in flight).
    const int bcols = div_up(a.cols, 4), brows = div_up(a.rows, 4);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = tid >> 2, q = tid & 3;
    const bool live = b < bcols * brows;       // lanes past the end stay for the shuffles and touch no memory
    const int bx = live ? b % bcols : 0, by = live ? b / bcols : 0;
    const int qx = q & 1, qy = q >> 1;

    float3 v0[2][2], n0[2][2];
    const int plane0 = a.rows * a.cols;
#pragma unroll
    for(int j = 0; j < 2; j++)
#pragma unroll
        for(int i = 0; i < 2; i++)
        {
            const int x = bx * 4 + 2 * qx + i, y = by * 4 + 2 * qy + j;
            float3 v = make_float3(SLAM_QNAN, SLAM_QNAN, SLAM_QNAN), n = v;
            if(live && x < a.cols && y < a.rows)
            {
                const float4 vs = __ldg(a.vsrc + y * a.cols + x);
                const float4 ns = __ldg(a.nsrc + y * a.cols + x);
                if(!(vs.z == 0))   // copyMapsKernel: validity of BOTH maps keyed on the vertex z
                {
                    v = make_float3(vs.x, vs.y, vs.z);
                    n = make_float3(ns.x, ns.y, ns.z);
                }
                const float dz = depth_from_vertex_z(vs.z, a.depth_cut);
                if(a.depth_tmp) a.depth_tmp[y * a.cols + x] = dz;
                if(a.lastDepth0)
                {
                    const int o0 = y * a.cols + x;
                    a.lastDepth0[o0] = dz;
                    a.nextDepth0[o0] = dz;
                    a.lastImage0[o0] = intensity_pixel(__ldg(a.model_rgba + o0));
                    a.nextImage0[o0] = intensity_pixel(__ldg(a.rgba + o0));
                }
                // level-0 copyMaps writes all three planes, NaN included
                if(a.transform)
                    store_map_pixel(a.vdst[0], a.ndst[0], plane0, y * a.cols + x, v, n, true, a.R, a.t);
                else
                {
                    const int o = y * a.cols + x;
                    a.vdst[0][o] = v.x; a.vdst[0][o + plane0] = v.y; a.vdst[0][o + 2 * plane0] = v.z;
                    a.ndst[0][o] = n.x; a.ndst[0][o + plane0] = n.y; a.ndst[0][o + 2 * plane0] = n.z;
                }
            }
            v0[j][i] = v;
            n0[j][i] = n;
        }
    if(a.levels < 2) return;

    const int rows1 = a.rows / 2, cols1 = a.cols / 2, plane1 = rows1 * cols1;
    const float3 v1 = resize4<false>(v0[0][0], v0[0][1], v0[1][0], v0[1][1]);
    const float3 n1 = resize4<true>(n0[0][0], n0[0][1], n0[1][0], n0[1][1]);
    {
        const int x = bx * 2 + qx, y = by * 2 + qy;
        if(live && x < cols1 && y < rows1) store_map_pixel(a.vdst[1], a.ndst[1], plane1, y * cols1 + x, v1, n1, a.transform, a.R, a.t);
    }
    if(a.levels < 3) return;

    const int rows2 = rows1 / 2, cols2 = cols1 / 2, plane2 = rows2 * cols2;
    float3 qv[4], qn[4];   // the block's level-1 pixels in the order (0,0) (0,1) (1,0) (1,1) = quarters 0..3
    const int base = (threadIdx.x & 31) & ~3;
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
        qv[k] = make_float3(__shfl_sync(0xffffffffu, v1.x, base + k), __shfl_sync(0xffffffffu, v1.y, base + k), __shfl_sync(0xffffffffu, v1.z, base + k));
        qn[k] = make_float3(__shfl_sync(0xffffffffu, n1.x, base + k), __shfl_sync(0xffffffffu, n1.y, base + k), __shfl_sync(0xffffffffu, n1.z, base + k));
    }
    if(live && q == 0 && bx < cols2 && by < rows2)
    {
        const float3 v2 = resize4<false>(qv[0], qv[1], qv[2], qv[3]);
        const float3 n2 = resize4<true>(qn[0], qn[1], qn[2], qn[3]);
        store_map_pixel(a.vdst[2], a.ndst[2], plane2, by * cols2 + bx, v2, n2, a.transform, a.R, a.t);
        if(a.vcam2) store_map_pixel(a.vcam2, a.ncam2, plane2, by * cols2 + bx, v2, n2, false, a.R, a.t);
    }
}

// Levels beyond the third (the reference's NUM_PYRS is 3; BASELINE config 3 asks for 4):
// resize a camera-frame level and write it transformed, same per-level operators.
__global__ void __launch_bounds__(256) k_resize_transform(const float * vsrc, const float * nsrc, int srows, int scols, float * vdst, float * ndst,
                                                          int transform, const Mat3 R, const float3 t, float * vcam, float * ncam)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if(o >= drows * dcols) return;
    const int y = o / dcols, x = o - y * dcols;
    const int splane = srows * scols, dplane = drows * dcols;
    const int s = (2 * y) * scols + 2 * x;
    float3 v[4], n[4];
    const int offs[4] = {s, s + 1, s + scols, s + scols + 1};
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
        v[k] = make_float3(vsrc[offs[k]], vsrc[offs[k] + splane], vsrc[offs[k] + 2 * splane]);
        n[k] = make_float3(nsrc[offs[k]], nsrc[offs[k] + splane], nsrc[offs[k] + 2 * splane]);
    }
    const float3 vr = resize4<false>(v[0], v[1], v[2], v[3]);
    const float3 nr = resize4<true>(n[0], n[1], n[2], n[3]);
    store_map_pixel(vdst, ndst, dplane, o, vr, nr, transform, R, t);
    if(vcam) store_map_pixel(vcam, ncam, dplane, o, vr, nr, false, R, t);
}

// ------------------------------------------------------------------ RGB-D pyramids
// populateRGBDData (RGBDOdometryef.cpp:208-235) in three dependent launches instead of
// six kernels + four cudaMalloc/cudaFree:
//   k_rgbd_level0 : depth_tmp -> depth[0] (copy), rgba -> image[0]
//   k_rgbd_down   : depth[l] -> depth[l+1] and image[l] -> image[l+1]   (l = 0, 1)
__global__ void __launch_bounds__(256) k_rgbd_level0(const float * __restrict__ depth_tmp, float * __restrict__ depth0, const uchar4 * __restrict__ rgba,
                                                     unsigned char * __restrict__ image0, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(depth0) depth0[i] = depth_tmp[i];
    image0[i] = intensity_pixel(__ldg(rgba + i));
}

__global__ void __launch_bounds__(256) k_rgbd_down(const float * __restrict__ dsrc, float * __restrict__ ddst, const unsigned char * __restrict__ isrc,
                                                   unsigned char * __restrict__ idst, int srows, int scols)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int nbx = div_up(dcols, 32);
    const int x = (blockIdx.x % nbx) * 32 + (threadIdx.x & 31);
    const int y = (blockIdx.x / nbx) * 8 + (threadIdx.x >> 5);
    if(x >= dcols || y >= drows) return;
    if(ddst) ddst[y * dcols + x] = pyr_down_gauss_f_pixel(dsrc, srows, scols, x, y);
    if(idst) idst[y * dcols + x] = pyr_down_gauss_u8_pixel(isrc, srows, scols, x, y);
}

// Frame-level variants used by the one-call-per-frame front end (slam_odom_track_*): the "last" (model) and "next"
// (current frame) pyramids of populateRGBDData are built side by side; in frame-to-model tracking both depth pyramids
// derive from the same model vertices (RGBDOdometryef.cpp:239,245), so each depth value is computed once and stored twice.
__global__ void __launch_bounds__(256) k_rgbd_level0_dual(const float * __restrict__ depth_tmp, float * __restrict__ lastDepth0, float * __restrict__ nextDepth0,
                                                          const uchar4 * __restrict__ model_rgba, unsigned char * __restrict__ lastImage0,
                                                          const uchar4 * __restrict__ rgba, unsigned char * __restrict__ nextImage0, int n, size_t in_stride,
                                                          size_t arena_stride)
{
    depth_tmp = seq_shift(depth_tmp, arena_stride);
    lastDepth0 = seq_shift(lastDepth0, arena_stride);
    nextDepth0 = seq_shift(nextDepth0, arena_stride);
    lastImage0 = seq_shift(lastImage0, arena_stride);
    nextImage0 = seq_shift(nextImage0, arena_stride);
    model_rgba = seq_shift(model_rgba, in_stride);
    rgba = seq_shift(rgba, in_stride);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float d = depth_tmp[i];
    lastDepth0[i] = d;
    nextDepth0[i] = d;
    lastImage0[i] = intensity_pixel(__ldg(model_rgba + i));
    nextImage0[i] = intensity_pixel(__ldg(rgba + i));
}

__global__ void __launch_bounds__(256) k_rgbd_down_dual(const float * __restrict__ dsrc, float * __restrict__ ddstLast, float * __restrict__ ddstNext,
                                                        const unsigned char * __restrict__ isrcLast, unsigned char * __restrict__ idstLast,
                                                        const unsigned char * __restrict__ isrcNext, unsigned char * __restrict__ idstNext, int srows, int scols,
                                                        size_t arena_stride)
{
    dsrc = seq_shift(dsrc, arena_stride);
    ddstLast = seq_shift(ddstLast, arena_stride);
    ddstNext = seq_shift(ddstNext, arena_stride);
    isrcLast = seq_shift(isrcLast, arena_stride);
    idstLast = seq_shift(idstLast, arena_stride);
    isrcNext = seq_shift(isrcNext, arena_stride);
    idstNext = seq_shift(idstNext, arena_stride);
    const int drows = srows / 2, dcols = scols / 2;
    const int nbx = div_up(dcols, 32);
    const int x = (blockIdx.x % nbx) * 32 + (threadIdx.x & 31);
    const int y = (blockIdx.x / nbx) * 8 + (threadIdx.x >> 5);
    if(x >= dcols || y >= drows) return;
    const float d = pyr_down_gauss_f_pixel(dsrc, srows, scols, x, y);
    ddstLast[y * dcols + x] = d;
    ddstNext[y * dcols + x] = d;
    idstLast[y * dcols + x] = pyr_down_gauss_u8_pixel(isrcLast, srows, scols, x, y);
    idstNext[y * dcols + x] = pyr_down_gauss_u8_pixel(isrcNext, srows, scols, x, y);
}

// ------------------------------------------------------------------ the preparation of a whole frame in ONE launch
// Three-level pyramids (the reference's NUM_PYRS).  Every block owns a tile of 64 x 32 level-0 pixels = 32 x 16 of level 1 = 16 x 8 of
// level 2 and produces every output of its tile on all three levels: the level-0 inputs of the tile PLUS the halo that the 5x5
// windows of the coarser levels reach are staged in shared memory once, the level-1 tile (with its own halo) is computed from
// that into shared memory, and level 2 from level 1 -- no level waits for another launch.  The halo pixels of levels 0 / 1 are
// computed redundantly by neighbouring blocks with the same operands in the same order, so every output is bit-identical to the
// per-level launches above (which stay for four-level pyramids, odd sizes and the operator-level entry points).
//   blocks [0, tiles)         model + RGB-D role: k_model_maps + 2 x k_rgbd_down_dual
//   blocks [tiles, 2 tiles)   sensor depth role:  3 x k_depth_level
// window operators on tiles: same code as on images (the accessor is a template parameter)
#define tile_pyr_down_f pyr_down_gauss_f_at
#define tile_pyr_down_u8 pyr_down_gauss_u8_at
#define tile_pyr_down_u16 pyr_down_u16_at
constexpr int kBorderListMax = 512;   // border pixels of one pass of one tile (3 images x (row + column of the level-1 tile) at most)
#define tile_vertex_normal vertex_normal_pixel

// Tile = 64 x 36 level-0 pixels: a 640 x 480 frame is 10 x 14 = 140 tiles, at most one block of each role per SM of a 148-SM part
// (with 64 x 32 tiles = 150 blocks two SMs carried two of the heavy model blocks and the launch took 1.6x the average SM's time).
constexpr int kTileW = 64, kTileH = 36, kPrepThreads = (kTileW / 4) * (kTileH / 4) * 4;   // four lanes per 4x4 block: 576 threads
// model role: level-0 tile = core + 6 before / + 2 after (the 5x5 windows of level 1, which itself needs + 2 before / + 1 after for level 2)
constexpr int kM0W = kTileW + 9, kM0H = kTileH + 9, kM0P8 = (kM0W + 3) & ~3, kM1W = kTileW / 2 + 3, kM1H = kTileH / 2 + 3, kM1P8 = (kM1W + 3) & ~3;
// depth role: the normals reach one pixel further on every level
constexpr int kD0W = kTileW + 13, kD0H = kTileH + 13, kD0P = (kD0W + 1) & ~1, kD1W = kTileW / 2 + 5, kD1H = kTileH / 2 + 5, kD1P = (kD1W + 1) & ~1,
              kD2W = kTileW / 4 + 1, kD2H = kTileH / 4 + 1, kD2P = (kD2W + 1) & ~1;
static_assert(kTileW / 4 == 16, "the lane -> 4x4 block mapping below assumes 16 blocks per tile row");

__device__ __forceinline__ void cp_async16_prep(void * smem, const void * gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
template <class T>
__device__ __forceinline__ T * shifted(T * p, size_t bytes) { return (T *)((const char *)p + bytes); }

__global__ void __launch_bounds__(kPrepThreads, 2) k_prepare_frame(const PrepFrameArgs a0)
{
    __shared__ __align__(16) unsigned char smem[kM0W * kM0H * 4 + 2 * kM0P8 * kM0H + kM1W * kM1H * 4 + 2 * kM1P8 * kM1H];
    static_assert(sizeof(smem) >= (kD0P * kD0H + kD1P * kD1H + kD2P * kD2H) * 2, "the depth role's tiles alias the model role's");
    // Pixels whose 5x5 window is clipped by the image border (one column of a left / right tile, one row of a top / bottom tile) are
    // set aside by the pass that meets them and handled afterwards by a few dense warps: a warp that runs the interior AND the clipped
    // path for one border lane per row made the left / right tiles 1.7x as expensive as the others, and they set the launch's duration.
    __shared__ int border_n[2];
    __shared__ unsigned short border_list[2][kBorderListMax];
    const int t = threadIdx.x;
    if(t < 2) border_n[t] = 0;
    // the kernel behind this one (the SO3 cluster of the split Gauss-Newton launch) may take its SMs while this grid drains; it waits
    // for this grid's completion itself (griddepcontrol.wait) before it reads anything
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long dbg_t0 = 0ull;
    if(a0.dbg && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
    // batched launch: byte offsets of this sequence (zero for a single sequence)
    const size_t sh_map = (size_t)blockIdx.y * a0.map_in_stride, sh_rgba = (size_t)blockIdx.y * a0.rgba_stride, sh_depth = (size_t)blockIdx.y * a0.depth_in_stride,
                 sh_arena = (size_t)blockIdx.y * a0.arena_stride;
    const int rows = a0.rows, cols = a0.cols;
    const int rows1 = rows / 2, cols1 = cols / 2, rows2 = rows1 / 2, cols2 = cols1 / 2;
    if((int)blockIdx.x < a0.model_blocks)
    {
        // ------------------------------------------------ model maps + both RGB-D pyramids
        const int tbx = blockIdx.x % a0.tiles_x, tby = blockIdx.x / a0.tiles_x;
        const float4 * vsrc = shifted(a0.vsrc, sh_map);
        const float4 * nsrc = shifted(a0.nsrc, sh_map);
        const uchar4 * model_rgba = shifted(a0.model_rgba, sh_rgba);
        const uchar4 * rgba = shifted(a0.rgba, sh_rgba);
        Mat3 R = a0.R;
        float3 tr = a0.t;
        if(a0.poses12)
        {
            const float * q = a0.poses12 + 12 * blockIdx.y;
            R.r0 = make_float3(__ldg(q + 0), __ldg(q + 1), __ldg(q + 2));
            R.r1 = make_float3(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5));
            R.r2 = make_float3(__ldg(q + 6), __ldg(q + 7), __ldg(q + 8));
            tr = make_float3(__ldg(q + 9), __ldg(q + 10), __ldg(q + 11));
        }
        float * d0 = reinterpret_cast<float *>(smem);
        unsigned char * li0 = smem + kM0W * kM0H * 4;
        unsigned char * ni0 = li0 + kM0P8 * kM0H;
        float * d1 = reinterpret_cast<float *>(ni0 + kM0P8 * kM0H);
        unsigned char * li1 = reinterpret_cast<unsigned char *>(d1 + kM1W * kM1H);
        unsigned char * ni1 = li1 + kM1P8 * kM1H;
        const int X0 = tbx * kTileW, Y0 = tby * kTileH;   // core origin, level 0
        const int c0 = X0 - 6, r0 = Y0 - 6;               // tile origin, level 0
        const int X1 = X0 / 2, Y1 = Y0 / 2, c1o = X1 - 2, r1o = Y1 - 2;

        // halo of level 0: depth (z of the model vertex after the maxDepthRGB cut) and the two intensities, shared memory only.  The
        // 945 halo pixels (6 rows above, 3 below, 6 columns left, 3 right of the core) are enumerated densely, four per thread, and
        // every load of the block -- halo and core -- is issued before the first is consumed (the inputs come from DRAM).
        constexpr int kHaloTop = 6 * kM0W, kHaloBottom = 3 * kM0W, kHaloSides = kTileH * 9, kHalo = kHaloTop + kHaloBottom + kHaloSides;   // 6 + 3 rows, 6 + 3 columns
        constexpr int kHaloPer = (kHalo + kPrepThreads - 1) / kPrepThreads;
        float hz[kHaloPer];
        uchar4 hl[kHaloPer], hn[kHaloPer];
        int hs[kHaloPer];   // position in the tile, -1: nothing
#pragma unroll
        for(int k = 0; k < kHaloPer; k++)
        {
            const int hidx = t + k * kPrepThreads;
            int r, c;
            if(hidx < kHaloTop)
            {
                r = hidx / kM0W;
                c = hidx - r * kM0W;
            }
            else if(hidx < kHaloTop + kHaloBottom)
            {
                r = (hidx - kHaloTop) / kM0W;
                c = (hidx - kHaloTop) - r * kM0W;
                r += 6 + kTileH;
            }
            else
            {
                const int q = hidx - kHaloTop - kHaloBottom;
                r = q / 9;
                c = q - r * 9;
                r += 6;
                c = c < 6 ? c : c + kTileW;
            }
            const int gx = c0 + c, gy = r0 + r;
            const bool inb = hidx < kHalo && gx >= 0 && gx < cols && gy >= 0 && gy < rows;   // outside: never read, the windows are clipped to the image
            hs[k] = inb ? r * kM0W + c : -1;
            const int o = inb ? gy * cols + gx : 0;
            hz[k] = inb ? __ldg(reinterpret_cast<const float *>(vsrc + o) + 2) : 0.f;
            hl[k] = inb ? __ldg(model_rgba + o) : make_uchar4(0, 0, 0, 0);
            hn[k] = inb ? __ldg(rgba + o) : make_uchar4(0, 0, 0, 0);
        }
        // core of level 0 + the model maps of all three levels: four consecutive lanes own one 4x4 pixel block, one 2x2 quarter
        // each (k_model_maps): 128 blocks per tile, 512 lanes
        float * vdst[3], * ndst[3];
#pragma unroll
        for(int l = 0; l < 3; l++)
        {
            vdst[l] = shifted(a0.vprev[l], sh_arena);
            ndst[l] = shifted(a0.nprev[l], sh_arena);
        }
        float * depth_tmp = shifted(a0.depth_tmp, sh_arena);
        float * lastDepth0 = shifted(a0.lastDepth[0], sh_arena);
        float * nextDepth0 = shifted(a0.nextDepth[0], sh_arena);
        unsigned char * lastImage0 = shifted(a0.lastImage[0], sh_arena);
        unsigned char * nextImage0 = shifted(a0.nextImage[0], sh_arena);
        const int bcols = cols / 4, brows = rows / 4;   // the host takes this path only when both are multiples of 4
        const int plane0 = rows * cols, plane1 = rows1 * cols1, plane2 = rows2 * cols2;
        // the core's inputs (model vertex texels, both RGBA images: 24 B per pixel) go to shared memory by cp.async: all of them in flight
        // at once with no register held for them, next to the halo loads above and the normal texels below (registers)
        extern __shared__ __align__(16) unsigned char prep_dyn[];
        float4 * s_v = reinterpret_cast<float4 *>(prep_dyn);
        uchar4 * s_m = reinterpret_cast<uchar4 *>(s_v + kTileW * kTileH);
        uchar4 * s_c = s_m + kTileW * kTileH;
        {
            const bool full_w = X0 + kTileW <= cols;
            // 16-byte chunks: one texel of a float4 map, four texels of an RGBA8 image
            for(int ch = t; ch < kTileW * kTileH; ch += kPrepThreads)
            {
                const int ry = ch / kTileW, rx = ch - ry * kTileW;
                if(Y0 + ry < rows && X0 + rx < cols)
                {
                    const int o = (Y0 + ry) * cols + X0 + rx;
                    cp_async16_prep(s_v + ch, vsrc + o);
                }
            }
            for(int ch = t; ch < kTileW * kTileH / 4; ch += kPrepThreads)
            {
                const int ry = ch / (kTileW / 4), rx = 4 * (ch - ry * (kTileW / 4));
                if(Y0 + ry < rows && (full_w || X0 + rx + 3 < cols))   // cols is a multiple of 4: a chunk is inside or outside as a whole
                {
                    const int o = (Y0 + ry) * cols + X0 + rx;
                    cp_async16_prep(s_m + ry * kTileW + rx, model_rgba + o);
                    cp_async16_prep(s_c + ry * kTileW + rx, rgba + o);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        {
            const int b = t >> 2, q = t & 3;
            const int bx = tbx * (kTileW / 4) + (b & 15), by = tby * (kTileH / 4) + (b >> 4);
            float4 cns[4];
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                const bool lv = bx < bcols && by < brows;
                const int x = bx * 4 + 2 * (q & 1) + (k & 1), y = by * 4 + 2 * (q >> 1) + (k >> 1);
                cns[k] = lv ? __ldg(nsrc + y * cols + x) : make_float4(0, 0, 0, 0);
            }
            {
#pragma unroll
                for(int k = 0; k < kHaloPer; k++)
                    if(hs[k] >= 0)
                    {
                        const int r = hs[k] / kM0W, c = hs[k] - r * kM0W;
                        d0[hs[k]] = depth_from_vertex_z(hz[k], a0.depth_cut);
                        li0[r * kM0P8 + c] = intensity_pixel(hl[k]);
                        ni0[r * kM0P8 + c] = intensity_pixel(hn[k]);
                    }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const bool live = bx < bcols && by < brows;   // lanes outside the image stay for the shuffles and touch no memory
            const int qx = q & 1, qy = q >> 1;
            float3 v0[2][2], n0[2][2];
#pragma unroll
            for(int j = 0; j < 2; j++)
#pragma unroll
                for(int i = 0; i < 2; i++)
                {
                    const int x = bx * 4 + 2 * qx + i, y = by * 4 + 2 * qy + j;
                    float3 v = make_float3(SLAM_QNAN, SLAM_QNAN, SLAM_QNAN), n = v;
                    if(live)
                    {
                        const int o = y * cols + x;
                        const int so = (y - Y0) * kTileW + (x - X0);
                        const float4 vs = s_v[so];
                        const float4 ns = cns[j * 2 + i];
                        if(!(vs.z == 0))   // copyMapsKernel: validity of BOTH maps keyed on the vertex z
                        {
                            v = make_float3(vs.x, vs.y, vs.z);
                            n = make_float3(ns.x, ns.y, ns.z);
                        }
                        const float dz = depth_from_vertex_z(vs.z, a0.depth_cut);
                        const unsigned char li = intensity_pixel(s_m[so]);
                        const unsigned char ni = intensity_pixel(s_c[so]);
                        depth_tmp[o] = dz;
                        lastDepth0[o] = dz;
                        nextDepth0[o] = dz;
                        lastImage0[o] = li;
                        nextImage0[o] = ni;
                        const int sr = y - r0, sc = x - c0;
                        d0[sr * kM0W + sc] = dz;
                        li0[sr * kM0P8 + sc] = li;
                        ni0[sr * kM0P8 + sc] = ni;
                        store_map_pixel(vdst[0], ndst[0], plane0, o, v, n, true, R, tr);   // level-0 copyMaps + tranformMaps
                    }
                    v0[j][i] = v;
                    n0[j][i] = n;
                }
            const float3 v1 = resize4<false>(v0[0][0], v0[0][1], v0[1][0], v0[1][1]);
            const float3 n1 = resize4<true>(n0[0][0], n0[0][1], n0[1][0], n0[1][1]);
            if(live) store_map_pixel(vdst[1], ndst[1], plane1, (by * 2 + qy) * cols1 + bx * 2 + qx, v1, n1, true, R, tr);
            float3 qv[4], qn[4];   // the block's level-1 pixels in the order (0,0) (0,1) (1,0) (1,1) = quarters 0..3
            const int base = (t & 31) & ~3;
#pragma unroll
            for(int k = 0; k < 4; k++)
            {
                qv[k] = make_float3(__shfl_sync(0xffffffffu, v1.x, base + k), __shfl_sync(0xffffffffu, v1.y, base + k), __shfl_sync(0xffffffffu, v1.z, base + k));
                qn[k] = make_float3(__shfl_sync(0xffffffffu, n1.x, base + k), __shfl_sync(0xffffffffu, n1.y, base + k), __shfl_sync(0xffffffffu, n1.z, base + k));
            }
            if(live && q == 0)
            {
                const float3 v2 = resize4<false>(qv[0], qv[1], qv[2], qv[3]);
                const float3 n2 = resize4<true>(qn[0], qn[1], qn[2], qn[3]);
                store_map_pixel(vdst[2], ndst[2], plane2, by * cols2 + bx, v2, n2, true, R, tr);
            }
        }
        __syncthreads();
        // level 1 of the RGB-D pyramids (core + halo) from the level-0 tile.  The work items are (image, pixel) pairs -- depth, then the
        // model's intensity, then the frame's -- so that every warp of the block has the same share (a warp works on one image).
        {
            const TileSrc<float, kM0W> sd(d0, r0, c0);
            const TileSrc<unsigned char, kM0P8> sl(li0, r0, c0), sn(ni0, r0, c0);
            constexpr int kPx = kM1W * kM1H, kPxPad = (kPx + 31) & ~31;   // padded: an image starts at a warp boundary
            for(int item = t; item < 3 * kPxPad; item += kPrepThreads)
            {
                const int img = item / kPxPad, idx = item - img * kPxPad;
                if(idx >= kPx) continue;
                const int r = idx / kM1W, c = idx - r * kM1W;
                const int x = c1o + c, y = r1o + r;
                if(x < 0 || x >= cols1 || y < 0 || y >= rows1) continue;
                if(!window_interior_gauss(rows, cols, x, y))
                {
                    border_list[0][atomicAdd(&border_n[0], 1)] = (unsigned short)(img * kPx + idx);
                    continue;
                }
                const bool core = r >= 2 && r < 2 + kTileH / 2 && c >= 2 && c < 2 + kTileW / 2;
                const int o = y * cols1 + x;
                if(img == 0)
                {
                    const float d = tile_pyr_down_f<1>(sd, rows, cols, x, y);
                    d1[r * kM1W + c] = d;
                    if(core)
                    {
                        shifted(a0.lastDepth[1], sh_arena)[o] = d;
                        shifted(a0.nextDepth[1], sh_arena)[o] = d;
                    }
                }
                else
                {
                    const unsigned char v = tile_pyr_down_u8<1>(img == 1 ? sl : sn, rows, cols, x, y);
                    (img == 1 ? li1 : ni1)[r * kM1P8 + c] = v;
                    if(core) shifted(img == 1 ? a0.lastImage[1] : a0.nextImage[1], sh_arena)[o] = v;
                }
            }
            __syncthreads();
            // the border pixels of the pass (none in most tiles)
            const int nb = border_n[0];
            for(int k = t; k < nb; k += kPrepThreads)
            {
                const int item = border_list[0][k];
                const int img = item / kPx, idx = item - img * kPx;
                const int r = idx / kM1W, c = idx - r * kM1W;
                const int x = c1o + c, y = r1o + r;
                const bool core = r >= 2 && r < 2 + kTileH / 2 && c >= 2 && c < 2 + kTileW / 2;
                const int o = y * cols1 + x;
                if(img == 0)
                {
                    const float d = tile_pyr_down_f<2>(sd, rows, cols, x, y);
                    d1[r * kM1W + c] = d;
                    if(core)
                    {
                        shifted(a0.lastDepth[1], sh_arena)[o] = d;
                        shifted(a0.nextDepth[1], sh_arena)[o] = d;
                    }
                }
                else
                {
                    const unsigned char v = tile_pyr_down_u8<2>(img == 1 ? sl : sn, rows, cols, x, y);
                    (img == 1 ? li1 : ni1)[r * kM1P8 + c] = v;
                    if(core) shifted(img == 1 ? a0.lastImage[1] : a0.nextImage[1], sh_arena)[o] = v;
                }
            }
        }
        __syncthreads();
        // level 2 from the level-1 tile: one (image, pixel) item per thread
        constexpr int kPx2 = (kTileW / 4) * (kTileH / 4);
        static_assert(3 * kPx2 <= kPrepThreads, "one item per thread");
        if(t < 3 * kPx2)
        {
            const int img = t / kPx2, k = t - img * kPx2;
            const int x = X0 / 4 + (k & 15), y = Y0 / 4 + (k >> 4);
            if(x < cols2 && y < rows2)
            {
                const int o = y * cols2 + x;
                if(!window_interior_gauss(rows1, cols1, x, y))
                    border_list[1][atomicAdd(&border_n[1], 1)] = (unsigned short)t;
                else if(img == 0)
                {
                    const float d = tile_pyr_down_f<1>(TileSrc<float, kM1W>(d1, r1o, c1o), rows1, cols1, x, y);
                    shifted(a0.lastDepth[2], sh_arena)[o] = d;
                    shifted(a0.nextDepth[2], sh_arena)[o] = d;
                }
                else
                    shifted(img == 1 ? a0.lastImage[2] : a0.nextImage[2], sh_arena)[o] =
                        tile_pyr_down_u8<1>(TileSrc<unsigned char, kM1P8>(img == 1 ? li1 : ni1, r1o, c1o), rows1, cols1, x, y);
            }
        }
        __syncthreads();
        if(t < border_n[1])
        {
            const int item = border_list[1][t];
            const int img = item / kPx2, k = item - img * kPx2;
            const int x = X0 / 4 + (k & 15), y = Y0 / 4 + (k >> 4);
            const int o = y * cols2 + x;
            if(img == 0)
            {
                const float d = tile_pyr_down_f<2>(TileSrc<float, kM1W>(d1, r1o, c1o), rows1, cols1, x, y);
                shifted(a0.lastDepth[2], sh_arena)[o] = d;
                shifted(a0.nextDepth[2], sh_arena)[o] = d;
            }
            else
                shifted(img == 1 ? a0.lastImage[2] : a0.nextImage[2], sh_arena)[o] =
                    tile_pyr_down_u8<2>(TileSrc<unsigned char, kM1P8>(img == 1 ? li1 : ni1, r1o, c1o), rows1, cols1, x, y);
        }
    }
    else
    {
        // ------------------------------------------------ the current frame's depth: pyramid + vertex / normal maps
        const int tile = blockIdx.x - a0.model_blocks;
        const int tbx = tile % a0.tiles_x, tby = tile / a0.tiles_x;
        const unsigned short * depth = shifted(a0.depth, sh_depth);
        unsigned short * z0 = reinterpret_cast<unsigned short *>(smem);
        unsigned short * z1 = z0 + kD0P * kD0H;
        unsigned short * z2 = z1 + kD1P * kD1H;
        const int X0 = tbx * kTileW, Y0 = tby * kTileH;
        const int c0 = X0 - 6, r0 = Y0 - 6;
        const int X1 = X0 / 2, Y1 = Y0 / 2, c1o = X1 - 2, r1o = Y1 - 2;
        const int X2 = X0 / 4, Y2 = Y0 / 4;
        {
            // the tile as 32-bit words (it starts at an even column and the image has an even number of columns), every load of the
            // thread in flight before the first store
            constexpr int kWordsPerRow = kD0P / 2, kWords = kWordsPerRow * kD0H, kPer = (kWords + kPrepThreads - 1) / kPrepThreads;
            unsigned w[kPer];
            int ws[kPer];
#pragma unroll
            for(int k = 0; k < kPer; k++)
            {
                const int idx = t + k * kPrepThreads;
                const int r = idx / kWordsPerRow, c = 2 * (idx - r * kWordsPerRow);
                const int gx = c0 + c, gy = r0 + r;
                const bool inb = idx < kWords && gx >= 0 && gx + 1 < cols + 1 && gx < cols && gy >= 0 && gy < rows;
                ws[k] = inb ? r * kWordsPerRow + (c >> 1) : -1;
                w[k] = inb ? __ldg(reinterpret_cast<const unsigned *>(depth + gy * cols + gx)) : 0u;
            }
#pragma unroll
            for(int k = 0; k < kPer; k++)
                if(ws[k] >= 0) reinterpret_cast<unsigned *>(z0)[ws[k]] = w[k];
        }
        __syncthreads();
        {
            const TileSrc<unsigned short, kD0P> s0(z0, r0, c0);
            float * vmap = shifted(a0.vcurr[0], sh_arena);
            float * nmap = shifted(a0.ncurr[0], sh_arena);
#pragma unroll 1
            for(int idx = t; idx < kTileW * kTileH; idx += kPrepThreads)
            {
                const int u = X0 + (idx & (kTileW - 1)), v = Y0 + idx / kTileW;
                if(u < cols && v < rows) tile_vertex_normal(s0, u, v, rows, cols, a0.fx_inv[0], a0.fy_inv[0], a0.cx[0], a0.cy[0], a0.depthCutoff, vmap, nmap);
            }
            unsigned short * dst1 = shifted(a0.depth_l[1], sh_arena);
            for(int idx = t; idx < kD1W * kD1H; idx += kPrepThreads)
            {
                const int r = idx / kD1W, c = idx - r * kD1W;
                const int x = c1o + c, y = r1o + r;
                if(x < 0 || x >= cols1 || y < 0 || y >= rows1) continue;
                if(!window_interior_u16(rows, cols, x, y))
                {
                    border_list[0][atomicAdd(&border_n[0], 1)] = (unsigned short)idx;
                    continue;
                }
                const unsigned short d = tile_pyr_down_u16<1>(s0, rows, cols, x, y);
                z1[r * kD1P + c] = d;
                if(r >= 2 && r < 2 + kTileH / 2 && c >= 2 && c < 2 + kTileW / 2) dst1[y * cols1 + x] = d;
            }
            __syncthreads();
            const int nb = border_n[0];
            for(int k = t; k < nb; k += kPrepThreads)
            {
                const int idx = border_list[0][k];
                const int r = idx / kD1W, c = idx - r * kD1W;
                const int x = c1o + c, y = r1o + r;
                const unsigned short d = tile_pyr_down_u16<2>(s0, rows, cols, x, y);
                z1[r * kD1P + c] = d;
                if(r >= 2 && r < 2 + kTileH / 2 && c >= 2 && c < 2 + kTileW / 2) dst1[y * cols1 + x] = d;
            }
        }
        __syncthreads();
        {
            const TileSrc<unsigned short, kD1P> s1(z1, r1o, c1o);
            float * vmap = shifted(a0.vcurr[1], sh_arena);
            float * nmap = shifted(a0.ncurr[1], sh_arena);
#pragma unroll 1
            for(int idx = t; idx < (kTileW / 2) * (kTileH / 2); idx += kPrepThreads)
            {
                const int u = X1 + (idx & (kTileW / 2 - 1)), v = Y1 + idx / (kTileW / 2);
                if(u < cols1 && v < rows1) tile_vertex_normal(s1, u, v, rows1, cols1, a0.fx_inv[1], a0.fy_inv[1], a0.cx[1], a0.cy[1], a0.depthCutoff, vmap, nmap);
            }
            unsigned short * dst2 = shifted(a0.depth_l[2], sh_arena);
            if(t < kD2W * kD2H)
            {
                const int r = t / kD2W, c = t - r * kD2W;
                const int x = X2 + c, y = Y2 + r;
                if(x < cols2 && y < rows2)
                {
                    if(!window_interior_u16(rows1, cols1, x, y))
                        border_list[1][atomicAdd(&border_n[1], 1)] = (unsigned short)t;
                    else
                    {
                        const unsigned short d = tile_pyr_down_u16<1>(s1, rows1, cols1, x, y);
                        z2[r * kD2P + c] = d;
                        if(r < kTileH / 4 && c < kTileW / 4) dst2[y * cols2 + x] = d;
                    }
                }
            }
            __syncthreads();
            if(t < border_n[1])
            {
                const int q = border_list[1][t];
                const int r = q / kD2W, c = q - r * kD2W;
                const int x = X2 + c, y = Y2 + r;
                const unsigned short d = tile_pyr_down_u16<2>(s1, rows1, cols1, x, y);
                z2[r * kD2P + c] = d;
                if(r < kTileH / 4 && c < kTileW / 4) dst2[y * cols2 + x] = d;
            }
        }
        __syncthreads();
        if(t < (kTileW / 4) * (kTileH / 4))
        {
            const TileSrc<unsigned short, kD2P> s2(z2, Y2, X2);
            const int u = X2 + (t & 15), v = Y2 + (t >> 4);
            if(u < cols2 && v < rows2)
                tile_vertex_normal(s2, u, v, rows2, cols2, a0.fx_inv[2], a0.fy_inv[2], a0.cx[2], a0.cy[2], a0.depthCutoff, shifted(a0.vcurr[2], sh_arena),
                                    shifted(a0.ncurr[2], sh_arena));
        }
    }
    if(a0.dbg)
    {
        __syncthreads();
        if(t == 0)
        {
            unsigned long long t1;
            unsigned sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            a0.dbg[3 * blockIdx.x + 0] = dbg_t0;
            a0.dbg[3 * blockIdx.x + 1] = t1;
            a0.dbg[3 * blockIdx.x + 2] = sm;
        }
    }
}

// ------------------------------------------------------------------ image derivatives
// applyKernel, utils.cu:582-606: running kernelIndex from 8 downwards over the CLIPPED
// window (border taps misalign, on purpose), float -> short truncation.  All levels in
// one launch: level l owns blocks [first[l], first[l+1]).
struct DerivArgs
{
    const unsigned char * src[SLAM_MAX_LEVELS];
    short * dx[SLAM_MAX_LEVELS];
    short * dy[SLAM_MAX_LEVELS];
    int rows[SLAM_MAX_LEVELS], cols[SLAM_MAX_LEVELS];
    int first[SLAM_MAX_LEVELS + 1];
    int levels;
    size_t arena_stride;   // batched launch: gridDim.y sequences
};

__global__ void __launch_bounds__(256) k_derivatives(const DerivArgs a)
{
    int l = 0;
    while(l + 1 < a.levels && (int)blockIdx.x >= a.first[l + 1]) l++;
    const int rows = a.rows[l], cols = a.cols[l];
    const int nbx = div_up(cols, 32);
    const int b = blockIdx.x - a.first[l];
    const int x = (b % nbx) * 32 + (threadIdx.x & 31);
    const int y = (b / nbx) * 8 + (threadIdx.x >> 5);
    if(x >= cols || y >= rows) return;
    short dx, dy;
    derivative_pixel(seq_shift(a.src[l], a.arena_stride), rows, cols, x, y, dx, dy);
    seq_shift(a.dx[l], a.arena_stride)[y * cols + x] = dx;
    seq_shift(a.dy[l], a.arena_stride)[y * cols + x] = dy;
}

// Frame-level variant for the batched engine: derivative images AND the pose-independent half of the RGB association
// (rgb_candidate, reduce.cu:780-807) of all levels and sequences in one launch.  One thread = 4 consecutive pixels; in the
// interior the 4x4 / 3x3 windows come from twelve aligned 32-bit loads of nextImage instead of 25 byte loads per pixel.
struct DerivCandArgs
{
    const unsigned char * src[SLAM_MAX_LEVELS];   // nextImage
    const float * depth[SLAM_MAX_LEVELS];         // nextDepth
    short * dx[SLAM_MAX_LEVELS];
    short * dy[SLAM_MAX_LEVELS];
    unsigned char * cand[SLAM_MAX_LEVELS];        // sequence 0; other sequences at cand_stride
    float min_scale[SLAM_MAX_LEVELS];
    int rows[SLAM_MAX_LEVELS], cols[SLAM_MAX_LEVELS];
    int first[SLAM_MAX_LEVELS + 1];
    int levels;
    size_t arena_stride, cand_stride;
};

template <int K>
__device__ __forceinline__ unsigned byte12(unsigned w0, unsigned w1, unsigned w2)   // byte K of the 12-byte row segment
{
    return ((K < 4 ? w0 : (K < 8 ? w1 : w2)) >> (8 * (K & 3))) & 0xffu;
}

template <int C>
__device__ __forceinline__ void deriv_from_words(const unsigned (&w)[4][3], short & dx, short & dy)
{
    // taps of pixel C: rows 1..3 of the window (y-1..y+1), columns 3+C .. 5+C, accumulated in the reference's order
    const float fgx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
    const float fgy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
    float v[9];
#pragma unroll
    for(int r = 0; r < 3; r++)
    {
        v[r * 3 + 0] = (float)byte12<3 + C>(w[r + 1][0], w[r + 1][1], w[r + 1][2]);
        v[r * 3 + 1] = (float)byte12<4 + C>(w[r + 1][0], w[r + 1][1], w[r + 1][2]);
        v[r * 3 + 2] = (float)byte12<5 + C>(w[r + 1][0], w[r + 1][1], w[r + 1][2]);
    }
    float dxVal = 0.f, dyVal = 0.f;
#pragma unroll
    for(int t = 0; t < 9; t++)
    {
        dxVal = __fmaf_rn(v[t], fgx[8 - t], dxVal);
        dyVal = __fmaf_rn(v[t], fgy[8 - t], dyVal);
    }
    dx = (short)dxVal;
    dy = (short)dyVal;
}

__global__ void __launch_bounds__(256) k_deriv_cand(const DerivCandArgs a)
{
    int l = 0;
    while(l + 1 < a.levels && (int)blockIdx.x >= a.first[l + 1]) l++;
    const int rows = a.rows[l], cols = a.cols[l];
    const int qcols = cols >> 2;                 // cols is a multiple of 4 (checked by the host)
    const int nbx = div_up(qcols, 32);
    const int b = blockIdx.x - a.first[l];
    const int qx = (b % nbx) * 32 + (threadIdx.x & 31);
    const int y = (b / nbx) * 8 + (threadIdx.x >> 5);
    if(qx >= qcols || y >= rows) return;
    const int x0 = qx * 4;
    const unsigned char * src = seq_shift(a.src[l], a.arena_stride);
    const float * depth = seq_shift(a.depth[l], a.arena_stride);
    short * dxp = seq_shift(a.dx[l], a.arena_stride);
    short * dyp = seq_shift(a.dy[l], a.arena_stride);
    unsigned char * cand = a.cand[l] ? a.cand[l] + (size_t)blockIdx.y * a.cand_stride : nullptr;
    const int o = y * cols + x0;
    short dx[4], dy[4];
    unsigned char cd[4] = {0, 0, 0, 0};
    const float minScale = a.min_scale[l];
    if(y >= 2 && y + 1 < rows && x0 >= 4 && x0 + 8 <= cols)
    {
        unsigned w[4][3];
#pragma unroll
        for(int r = 0; r < 4; r++)
        {
            const unsigned * p = reinterpret_cast<const unsigned *>(src + (y - 2 + r) * cols + x0 - 4);
            w[r][0] = __ldg(p);
            w[r][1] = __ldg(p + 1);
            w[r][2] = __ldg(p + 2);
        }
        deriv_from_words<0>(w, dx[0], dy[0]);
        deriv_from_words<1>(w, dx[1], dy[1]);
        deriv_from_words<2>(w, dx[2], dy[2]);
        deriv_from_words<3>(w, dx[3], dy[3]);
        if(cand)
        {
            // zero bytes anywhere in rows y-2..y+1 (0xff per zero byte), then the 4-column window of each pixel
            const unsigned z0 = __vcmpeq4(w[0][0], 0u) | __vcmpeq4(w[1][0], 0u) | __vcmpeq4(w[2][0], 0u) | __vcmpeq4(w[3][0], 0u);
            const unsigned z1 = __vcmpeq4(w[0][1], 0u) | __vcmpeq4(w[1][1], 0u) | __vcmpeq4(w[2][1], 0u) | __vcmpeq4(w[3][1], 0u);
            const unsigned z2 = __vcmpeq4(w[0][2], 0u) | __vcmpeq4(w[1][2], 0u) | __vcmpeq4(w[2][2], 0u) | __vcmpeq4(w[3][2], 0u);
            const unsigned win[4] = {__funnelshift_r(z0, z1, 16), __funnelshift_r(z0, z1, 24), z1, __funnelshift_r(z1, z2, 8)};   // columns x-2 .. x+1
            const float4 d = __ldg(reinterpret_cast<const float4 *>(depth + o));
            const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for(int c = 0; c < 4; c++)
            {
                const int valx = dx[c], valy = dy[c];
                const float mTwo = (valx * valx) + (valy * valy);
                cd[c] = (x0 + c < cols - 5 && win[c] == 0u && mTwo >= minScale && !isnan(dd[c])) ? 1 : 0;
            }
        }
    }
    else
    {
#pragma unroll
        for(int c = 0; c < 4; c++)
        {
            const int x = x0 + c;
            derivative_pixel(src, rows, cols, x, y, dx[c], dy[c]);
            if(cand)
            {
                bool ok = x < cols - 5 && y < rows - 1;
                if(ok)
                    for(int u = max(y - 2, 0); u < min(y + 2, rows); u++)
                        for(int v = max(x - 2, 0); v < min(x + 2, cols); v++) ok = ok && (src[u * cols + v] > 0);
                if(ok)
                {
                    const int valx = dx[c], valy = dy[c];
                    const float mTwo = (valx * valx) + (valy * valy);
                    ok = mTwo >= minScale && !isnan(depth[o + c]);
                }
                cd[c] = ok ? 1 : 0;
            }
        }
    }
    *reinterpret_cast<short4 *>(dxp + o) = make_short4(dx[0], dx[1], dx[2], dx[3]);
    *reinterpret_cast<short4 *>(dyp + o) = make_short4(dy[0], dy[1], dy[2], dy[3]);
    if(cand) *reinterpret_cast<uchar4 *>(cand + o) = make_uchar4(cd[0], cd[1], cd[2], cd[3]);
}

// ------------------------------------------------------------------ depth pre-filter
// The 13x13 bilateral filter that produces DEPTH_FILTERED, the input of initICP (gl/shaders/depth_bilateral.frag:30-76 run by
// gl/ComputePack.cpp:41-73; caller apps/elastic_fusion_file.cpp:342-346).  Stencil kernel: a 32x8 output tile + 6-pixel halo
// is staged in shared memory, the 169 spatial exponents come from a constant table (same fp32 values as computed in place),
// every pixel walks its clipped window in the shader's order with the shader's operations (no FMA contraction: the sums
// decide a round-to-integer).  Compute-bound: 169 exp per valid pixel; HBM traffic is 4 B/px.
constexpr int kBilR = 6;
constexpr int kBilTileX = 32, kBilTileY = 8;
__constant__ float c_bil_space[13 * 13];

__global__ void __launch_bounds__(256) k_depth_bilateral(const unsigned short * __restrict__ src, int rows, int cols, unsigned cut, unsigned short * __restrict__ dst,
                                                         size_t image_stride)
{
    __shared__ unsigned short tile[kBilTileY + 2 * kBilR][kBilTileX + 2 * kBilR + 4];
    src = reinterpret_cast<const unsigned short *>(reinterpret_cast<const char *>(src) + (size_t)blockIdx.z * image_stride);
    dst = reinterpret_cast<unsigned short *>(reinterpret_cast<char *>(dst) + (size_t)blockIdx.z * image_stride);
    const int bx = blockIdx.x * kBilTileX, by = blockIdx.y * kBilTileY;
    for(int k = threadIdx.x; k < (kBilTileY + 2 * kBilR) * (kBilTileX + 2 * kBilR); k += blockDim.x)
    {
        const int ty = k / (kBilTileX + 2 * kBilR), tx = k - ty * (kBilTileX + 2 * kBilR);
        const int gx = bx + tx - kBilR, gy = by + ty - kBilR;
        tile[ty][tx] = (gx >= 0 && gy >= 0 && gx < cols && gy < rows) ? __ldg(src + gy * cols + gx) : (unsigned short)0;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = bx + lx, y = by + ly;
    if(x >= cols || y >= rows) return;
    const unsigned value = tile[ly + kBilR][lx + kBilR];
    if(value > cut || value < 300u)
    {
        dst[y * cols + x] = 0;
        return;
    }
    const float sigma_color2_inv_half = 0.000555556f;
    const float fv = (float)value;
    const int cy0 = max(y - kBilR, 0), cy1 = min(y + kBilR + 1, rows);
    const int cx0 = max(x - kBilR, 0), cx1 = min(x + kBilR + 1, cols);
    float sum1 = 0.f, sum2 = 0.f;
    for(int cy = cy0; cy < cy1; ++cy)
    {
        const int trow = cy - by + kBilR, tcol0 = kBilR - bx;
        const int srow = (cy - y + kBilR) * 13 + (kBilR - x);
#pragma unroll 13
        for(int cx = cx0; cx < cx1; ++cx)
        {
            const float ft = (float)tile[trow][cx + tcol0];
            const float dc = __fsub_rn(fv, ft);
            const float arg = __fadd_rn(c_bil_space[srow + cx], __fmul_rn(__fmul_rn(dc, dc), sigma_color2_inv_half));
            const float weight = expf(-arg);
            sum1 = __fadd_rn(sum1, __fmul_rn(ft, weight));
            sum2 = __fadd_rn(sum2, weight);
        }
    }
    dst[y * cols + x] = (unsigned short)(unsigned)roundf(__fdiv_rn(sum1, sum2));
}

int launch_depth_bilateral(const unsigned short * src, int rows, int cols, float max_depth_m, unsigned short * dst, int n_images, cudaStream_t s)
{
    static bool table_ready[64] = {};
    int dev = 0;
    SLAM_CUDA_TRY(cudaGetDevice(&dev));
    if(dev < 64 && !table_ready[dev])
    {
        float tab[13 * 13];
        for(int j = 0; j < 13; j++)
            for(int i = 0; i < 13; i++)
            {
                // (float(x) - float(cx))^2 + (float(y) - float(cy))^2, times sigma_space2_inv_half: small integers, exact products
                const float dx = (float)(kBilR - i), dy = (float)(kBilR - j);
                const float space2 = dx * dx + dy * dy;
                tab[j * 13 + i] = space2 * 0.024691358f;
            }
        SLAM_CUDA_TRY(cudaMemcpyToSymbol(c_bil_space, tab, sizeof(tab)));
        table_ready[dev] = true;
    }
    const unsigned cut = (unsigned)(max_depth_m * 1000.0f);
    k_depth_bilateral<<<dim3(div_up(cols, kBilTileX), div_up(rows, kBilTileY), n_images), 256, 0, s>>>(src, rows, cols, cut, dst, (size_t)rows * cols * 2);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

// ------------------------------------------------------------------ un-fused operator kernels
__global__ void __launch_bounds__(256) k_pyr_down_u16(const unsigned short * src, int srows, int scols, unsigned short * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int nbx = div_up(dcols, 32);
    const int x = (blockIdx.x % nbx) * 32 + (threadIdx.x & 31);
    const int y = (blockIdx.x / nbx) * 8 + (threadIdx.x >> 5);
    if(x >= dcols || y >= drows) return;
    dst[y * dcols + x] = pyr_down_u16_pixel(src, srows, scols, x, y);
}

__global__ void __launch_bounds__(256) k_create_vmap(const unsigned short * depth, int rows, int cols, float fx_inv, float fy_inv, float cx, float cy,
                                                     float depthCutoff, float * vmap)
{
    const int nbx = div_up(cols, 32);
    const int u = (blockIdx.x % nbx) * 32 + (threadIdx.x & 31);
    const int v = (blockIdx.x / nbx) * 8 + (threadIdx.x >> 5);
    if(u >= cols || v >= rows) return;
    const int plane = rows * cols, o = v * cols + u;
    float3 p;
    if(vertex_from_depth(depth[o], u, v, fx_inv, fy_inv, cx, cy, depthCutoff, p))
    {
        vmap[o] = p.x;
        vmap[o + plane] = p.y;
        vmap[o + 2 * plane] = p.z;
    }
    else
        vmap[o] = SLAM_QNAN;
}

// utils.cu:151-188 computeNmapKernel (reads an existing vertex map)
__global__ void __launch_bounds__(256) k_create_nmap(const float * vmap, int rows, int cols, float * nmap)
{
    const int nbx = div_up(cols, 32);
    const int u = (blockIdx.x % nbx) * 32 + (threadIdx.x & 31);
    const int v = (blockIdx.x / nbx) * 8 + (threadIdx.x >> 5);
    if(u >= cols || v >= rows) return;
    const int plane = rows * cols, o = v * cols + u;
    if(u == cols - 1 || v == rows - 1)
    {
        nmap[o] = SLAM_QNAN;
        return;
    }
    float3 v00, v01, v10;
    v00.x = vmap[o];
    v01.x = vmap[o + 1];
    v10.x = vmap[o + cols];
    if(!isnan(v00.x) && !isnan(v01.x) && !isnan(v10.x))
    {
        v00.y = vmap[o + plane];
        v01.y = vmap[o + 1 + plane];
        v10.y = vmap[o + cols + plane];
        v00.z = vmap[o + 2 * plane];
        v01.z = vmap[o + 1 + 2 * plane];
        v10.z = vmap[o + cols + 2 * plane];
        const float3 r = unit3(cross3(v01 - v00, v10 - v00));
        nmap[o] = r.x;
        nmap[o + plane] = r.y;
        nmap[o + 2 * plane] = r.z;
    }
    else
        nmap[o] = SLAM_QNAN;
}

__global__ void __launch_bounds__(256) k_transform_maps(const float * vsrc, const float * nsrc, int rows, int cols, const Mat3 R, const float3 t,
                                                        float * vdst, float * ndst)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = rows * cols;
    if(o >= plane) return;
    float3 v = make_float3(vsrc[o], 0.f, 0.f), n = make_float3(nsrc[o], 0.f, 0.f);
    if(!isnan(v.x))
    {
        v.y = vsrc[o + plane];
        v.z = vsrc[o + 2 * plane];
    }
    if(!isnan(n.x))
    {
        n.y = nsrc[o + plane];
        n.z = nsrc[o + 2 * plane];
    }
    store_map_pixel(vdst, ndst, plane, o, v, n, true, R, t);
}

template <bool normalize>
__global__ void __launch_bounds__(256) k_resize_map(const float * src, int srows, int scols, float * dst)
{
    const int drows = srows / 2, dcols = scols / 2;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if(o >= drows * dcols) return;
    const int y = o / dcols, x = o - y * dcols;
    const int splane = srows * scols, dplane = drows * dcols;
    const int s = (2 * y) * scols + 2 * x;
    float3 a00, a01, a10, a11;
    a00.x = src[s]; a01.x = src[s + 1]; a10.x = src[s + scols]; a11.x = src[s + scols + 1];
    if(isnan(a00.x) || isnan(a01.x) || isnan(a10.x) || isnan(a11.x))
    {
        dst[o] = SLAM_QNAN;
        return;
    }
    a00.y = src[s + splane]; a01.y = src[s + 1 + splane]; a10.y = src[s + scols + splane]; a11.y = src[s + scols + 1 + splane];
    a00.z = src[s + 2 * splane]; a01.z = src[s + 1 + 2 * splane]; a10.z = src[s + scols + 2 * splane]; a11.z = src[s + scols + 1 + 2 * splane];
    const float3 n = resize4<normalize>(a00, a01, a10, a11);
    dst[o] = n.x;
    dst[o + dplane] = n.y;
    dst[o + 2 * dplane] = n.z;
}

__global__ void __launch_bounds__(256) k_vertices_to_depth(const float4 * vsrc, int n, float * dst, float cutOff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    dst[i] = depth_from_vertex_z(__ldg(vsrc + i).z, cutOff);
}

// utils.cu:640-658 projectPointsKernel
__global__ void __launch_bounds__(256) k_project_points(const float * depth, int rows, int cols, float * cloud3, float invFx, float invFy, float cx,
                                                        float cy)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if(o >= rows * cols) return;
    const int y = o / cols, x = o - y * cols;
    const float3 p = cloud_point(depth, cols, x, y, invFx, invFy, cx, cy);
    cloud3[3 * o + 0] = p.x;
    cloud3[3 * o + 1] = p.y;
    cloud3[3 * o + 2] = p.z;
}

// ------------------------------------------------------------------ host launchers
static inline int tiles_32x8(int rows, int cols) { return div_up(cols, 32) * div_up(rows, 8); }

int launch_depth_level(const unsigned short * depth, int rows, int cols, float fx, float fy, float cx, float cy, float depthCutoff, float * vmap,
                       float * nmap, unsigned short * next_depth, cudaStream_t s, int nseq, size_t in_stride, size_t out_stride)
{
    const int nb_map_x = div_up(cols, 32);
    const int nb_map = tiles_32x8(rows, cols);
    const int nb_down = next_depth ? tiles_32x8(rows / 2, cols / 2) : 0;
    k_depth_level<<<dim3(nb_map + nb_down, nseq), 256, 0, s>>>(depth, rows, cols, 1.f / fx, 1.f / fy, cx, cy, depthCutoff, vmap, nmap, next_depth, nb_map_x, nb_map,
                                                               in_stride, out_stride);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_prepare_frame(PrepFrameArgs & a, cudaStream_t s, int nseq)
{
    a.tiles_x = div_up(a.cols, kTileW);
    a.tiles = a.tiles_x * div_up(a.rows, kTileH);
    a.model_blocks = a.tiles;
    int blocks = 2 * a.tiles;
    static const char * only = getenv("SLAM_PREP_ONLY_ROLE");   // development aid (timing of one role; the results are incomplete)
    if(only && only[0] == 'm') blocks = a.tiles;
    if(only && only[0] == 'd') a.model_blocks = 0, blocks = a.tiles;
    constexpr int kDyn = kTileW * kTileH * 24;   // the model role's input tile (float4 vertex, two RGBA8 texels per pixel)
    static bool attr_set = false;
    if(!attr_set)
    {
        SLAM_CUDA_TRY(cudaFuncSetAttribute((const void *)k_prepare_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, kDyn));
        attr_set = true;
    }
    static const bool debug = getenv("SLAM_PREP_DEBUG") != nullptr;
    static unsigned long long * dbg = nullptr;
    static int dbg_calls = 0;
    a.dbg = nullptr;
    if(debug && nseq == 1)
    {
        if(!dbg) SLAM_CUDA_TRY(cudaMalloc((void **)&dbg, sizeof(unsigned long long) * 3 * 1024));
        a.dbg = dbg;
    }
    k_prepare_frame<<<dim3(blocks, nseq), kPrepThreads, kDyn, s>>>(a);
    SLAM_CUDA_TRY(cudaGetLastError());
    if(a.dbg && ++dbg_calls == 40)
    {
        // per block: start and duration in ns relative to the first block, SM
        std::vector<unsigned long long> hbuf(3 * blocks);
        SLAM_CUDA_TRY(cudaStreamSynchronize(s));
        SLAM_CUDA_TRY(cudaMemcpy(hbuf.data(), dbg, sizeof(unsigned long long) * 3 * blocks, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for(int b = 0; b < blocks; b++) t0 = hbuf[3 * b] < t0 ? hbuf[3 * b] : t0;
        for(int b = 0; b < blocks; b++)
            fprintf(stderr, "prep block %3d role %c tile %3d sm %3d start %6llu dur %6llu\n", b, b < a.model_blocks ? 'm' : 'd', b < a.model_blocks ? b : b - a.model_blocks,
                    (int)hbuf[3 * b + 2], hbuf[3 * b] - t0, hbuf[3 * b + 1] - hbuf[3 * b]);
    }
    return SLAM_OK;
}

int launch_model_maps(const ModelMapsArgs & a, cudaStream_t s, int nseq = 1)
{
    const int nblk = div_up(a.cols, 4) * div_up(a.rows, 4);
    k_model_maps<<<dim3(div_up(nblk * 4, 128), nseq), 128, 0, s>>>(a);   // four lanes per 4x4 block
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_resize_transform(const float * vsrc, const float * nsrc, int srows, int scols, float * vdst, float * ndst, int transform, const Mat3 & R,
                            const float3 & t, float * vcam, float * ncam, cudaStream_t s)
{
    k_resize_transform<<<div_up((srows / 2) * (scols / 2), 256), 256, 0, s>>>(vsrc, nsrc, srows, scols, vdst, ndst, transform, R, t, vcam, ncam);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_rgbd_level0(const float * depth_tmp, float * depth0, const uchar4 * rgba, unsigned char * image0, int n, cudaStream_t s)
{
    k_rgbd_level0<<<div_up(n, 256), 256, 0, s>>>(depth_tmp, depth0, rgba, image0, n);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_rgbd_level0_dual(const float * depth_tmp, float * lastDepth0, float * nextDepth0, const uchar4 * model_rgba, unsigned char * lastImage0,
                            const uchar4 * rgba, unsigned char * nextImage0, int n, cudaStream_t s, int nseq, size_t in_stride, size_t arena_stride)
{
    k_rgbd_level0_dual<<<dim3(div_up(n, 256), nseq), 256, 0, s>>>(depth_tmp, lastDepth0, nextDepth0, model_rgba, lastImage0, rgba, nextImage0, n, in_stride,
                                                                  arena_stride);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_rgbd_down_dual(const float * dsrc, float * ddstLast, float * ddstNext, const unsigned char * isrcLast, unsigned char * idstLast,
                          const unsigned char * isrcNext, unsigned char * idstNext, int srows, int scols, cudaStream_t s, int nseq, size_t arena_stride)
{
    k_rgbd_down_dual<<<dim3(tiles_32x8(srows / 2, scols / 2), nseq), 256, 0, s>>>(dsrc, ddstLast, ddstNext, isrcLast, idstLast, isrcNext, idstNext, srows, scols,
                                                                                  arena_stride);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_rgbd_down(const float * dsrc, float * ddst, const unsigned char * isrc, unsigned char * idst, int srows, int scols, cudaStream_t s)
{
    k_rgbd_down<<<tiles_32x8(srows / 2, scols / 2), 256, 0, s>>>(dsrc, ddst, isrc, idst, srows, scols);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_derivatives(DerivArgs & a, cudaStream_t s, int nseq = 1)
{
    int total = 0;
    for(int l = 0; l < a.levels; l++)
    {
        a.first[l] = total;
        total += tiles_32x8(a.rows[l], a.cols[l]);
    }
    a.first[a.levels] = total;
    k_derivatives<<<dim3(total, nseq), 256, 0, s>>>(a);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_deriv_cand(int levels, const unsigned char * const * src, const float * const * depth, short * const * dx, short * const * dy,
                      unsigned char * const * cand, const float * min_scale, const int * rows, const int * cols, cudaStream_t s, int nseq, size_t arena_stride,
                      size_t cand_stride)
{
    DerivCandArgs a = {};
    a.levels = levels;
    a.arena_stride = arena_stride;
    a.cand_stride = cand_stride;
    int total = 0;
    for(int l = 0; l < levels; l++)
    {
        a.src[l] = src[l];
        a.depth[l] = depth[l];
        a.dx[l] = dx[l];
        a.dy[l] = dy[l];
        a.cand[l] = cand ? cand[l] : nullptr;
        a.min_scale[l] = min_scale ? min_scale[l] : 0.f;
        a.rows[l] = rows[l];
        a.cols[l] = cols[l];
        a.first[l] = total;
        total += div_up(cols[l] / 4, 32) * div_up(rows[l], 8);
    }
    a.first[levels] = total;
    k_deriv_cand<<<dim3(total, nseq), 256, 0, s>>>(a);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

int launch_model_maps_simple(const float4 * vsrc, const float4 * nsrc, int rows, int cols, int levels, float * const * vdst, float * const * ndst,
                             int transform, const Mat3 & R, const float3 & t, float * depth_tmp, float depth_cut, float * vcam2, float * ncam2,
                             cudaStream_t s, int nseq, size_t in_stride, size_t out_stride, const float * poses12, float * lastDepth0, float * nextDepth0,
                             const uchar4 * model_rgba, const uchar4 * rgba, unsigned char * lastImage0, unsigned char * nextImage0, size_t rgba_stride)
{
    ModelMapsArgs a = {};
    a.vsrc = vsrc;
    a.nsrc = nsrc;
    a.rows = rows;
    a.cols = cols;
    a.levels = levels < 3 ? levels : 3;
    for(int l = 0; l < a.levels; l++)
    {
        a.vdst[l] = vdst[l];
        a.ndst[l] = ndst[l];
    }
    a.transform = transform;
    a.R = R;
    a.t = t;
    a.depth_tmp = depth_tmp;
    a.depth_cut = depth_cut;
    a.vcam2 = vcam2;
    a.ncam2 = ncam2;
    a.in_stride = in_stride;
    a.out_stride = out_stride;
    a.poses12 = poses12;
    a.lastDepth0 = lastDepth0;
    a.nextDepth0 = nextDepth0;
    a.model_rgba = model_rgba;
    a.rgba = rgba;
    a.lastImage0 = lastImage0;
    a.nextImage0 = nextImage0;
    a.rgba_stride = rgba_stride;
    return launch_model_maps(a, s, nseq);
}

int launch_derivatives_simple(int levels, const unsigned char * const * src, short * const * dx, short * const * dy, const int * rows, const int * cols,
                              cudaStream_t s, int nseq, size_t arena_stride)
{
    DerivArgs a = {};
    a.arena_stride = arena_stride;
    a.levels = levels;
    for(int l = 0; l < levels; l++)
    {
        a.src[l] = src[l];
        a.dx[l] = dx[l];
        a.dy[l] = dy[l];
        a.rows[l] = rows[l];
        a.cols[l] = cols[l];
    }
    return launch_derivatives(a, s, nseq);
}

}   // namespace slam

// ------------------------------------------------------------------ C ABI (operators)
using namespace slam;

extern "C" int slam_op_depth_bilateral(const uint16_t * src, int rows, int cols, float max_depth_m, uint16_t * dst, int n_images, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src != dst && rows > 0 && cols > 0 && n_images > 0 && n_images <= 65535);
    return launch_depth_bilateral(src, rows, cols, max_depth_m, dst, n_images, (cudaStream_t)stream);
}

extern "C" int slam_op_pyr_down(const uint16_t * src, int src_rows, int src_cols, uint16_t * dst, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src_rows > 1 && src_cols > 1);
    k_pyr_down_u16<<<tiles_32x8(src_rows / 2, src_cols / 2), 256, 0, (cudaStream_t)stream>>>(src, src_rows, src_cols, dst);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_create_vmap(float fx, float fy, float cx, float cy, const uint16_t * depth, int rows, int cols, float * vmap,
                                   float depth_cutoff, void * stream)
{
    SLAM_ARG_CHECK(depth && vmap && rows > 0 && cols > 0);
    k_create_vmap<<<tiles_32x8(rows, cols), 256, 0, (cudaStream_t)stream>>>(depth, rows, cols, 1.f / fx, 1.f / fy, cx, cy, depth_cutoff, vmap);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_create_nmap(const float * vmap, int rows, int cols, float * nmap, void * stream)
{
    SLAM_ARG_CHECK(vmap && nmap && rows > 0 && cols > 0);
    k_create_nmap<<<tiles_32x8(rows, cols), 256, 0, (cudaStream_t)stream>>>(vmap, rows, cols, nmap);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_transform_maps(const float * vmap_src, const float * nmap_src, int rows, int cols, const float * R9, const float * t3,
                                      float * vmap_dst, float * nmap_dst, void * stream)
{
    SLAM_ARG_CHECK(vmap_src && nmap_src && R9 && t3 && vmap_dst && nmap_dst && rows > 0 && cols > 0);
    k_transform_maps<<<div_up(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(vmap_src, nmap_src, rows, cols, mat3_from(R9),
                                                                                 make_float3(t3[0], t3[1], t3[2]), vmap_dst, nmap_dst);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_copy_maps(const float * vertices4, const float * normals4, int rows, int cols, float * vmap_dst, float * nmap_dst,
                                 void * stream)
{
    SLAM_ARG_CHECK(vertices4 && normals4 && vmap_dst && nmap_dst && rows > 0 && cols > 0);
    ModelMapsArgs a = {};
    a.vsrc = reinterpret_cast<const float4 *>(vertices4);
    a.nsrc = reinterpret_cast<const float4 *>(normals4);
    a.rows = rows; a.cols = cols; a.levels = 1;
    a.vdst[0] = vmap_dst; a.ndst[0] = nmap_dst;
    a.transform = 0;
    a.depth_tmp = nullptr;
    return launch_model_maps(a, (cudaStream_t)stream);
}

extern "C" int slam_op_resize_vmap(const float * src, int src_rows, int src_cols, float * dst, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src_rows > 1 && src_cols > 1);
    k_resize_map<false><<<div_up((src_rows / 2) * (src_cols / 2), 256), 256, 0, (cudaStream_t)stream>>>(src, src_rows, src_cols, dst);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_resize_nmap(const float * src, int src_rows, int src_cols, float * dst, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src_rows > 1 && src_cols > 1);
    k_resize_map<true><<<div_up((src_rows / 2) * (src_cols / 2), 256), 256, 0, (cudaStream_t)stream>>>(src, src_rows, src_cols, dst);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_image_bgr_to_intensity(const uint8_t * rgba, int rows, int cols, uint8_t * dst, void * stream)
{
    SLAM_ARG_CHECK(rgba && dst && rows > 0 && cols > 0);
    return launch_rgbd_level0(nullptr, nullptr, reinterpret_cast<const uchar4 *>(rgba), dst, rows * cols, (cudaStream_t)stream);
}

extern "C" int slam_op_vertices_to_depth(const float * vertices4, int rows, int cols, float * dst, float cutoff, void * stream)
{
    SLAM_ARG_CHECK(vertices4 && dst && rows > 0 && cols > 0);
    k_vertices_to_depth<<<div_up(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(vertices4), rows * cols, dst, cutoff);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_project_to_point_cloud(const float * depth, int rows, int cols, float * cloud3, float fx, float fy, float cx, float cy,
                                              int level, void * stream)
{
    SLAM_ARG_CHECK(depth && cloud3 && rows > 0 && cols > 0 && level >= 0 && level < 16);
    const int div = 1 << level;   // CameraModel::operator()(level), sensors/Camera.h:14-18
    const float lfx = fx / div, lfy = fy / div, lcx = cx / div, lcy = cy / div;
    k_project_points<<<div_up(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(depth, rows, cols, cloud3, 1.0f / lfx, 1.0f / lfy, lcx, lcy);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_op_pyr_down_gauss_f(const float * src, int src_rows, int src_cols, float * dst, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src_rows > 1 && src_cols > 1);
    return launch_rgbd_down(src, dst, nullptr, nullptr, src_rows, src_cols, (cudaStream_t)stream);
}

extern "C" int slam_op_pyr_down_uchar_gauss(const uint8_t * src, int src_rows, int src_cols, uint8_t * dst, void * stream)
{
    SLAM_ARG_CHECK(src && dst && src_rows > 1 && src_cols > 1);
    return launch_rgbd_down(nullptr, nullptr, src, dst, src_rows, src_cols, (cudaStream_t)stream);
}

extern "C" int slam_op_compute_derivative_images(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy, void * stream)
{
    SLAM_ARG_CHECK(src && dx && dy && rows > 0 && cols > 0);
    DerivArgs a = {};
    a.levels = 1;
    a.src[0] = src; a.dx[0] = dx; a.dy[0] = dy;
    a.rows[0] = rows; a.cols[0] = cols;
    return launch_derivatives(a, (cudaStream_t)stream);
}
