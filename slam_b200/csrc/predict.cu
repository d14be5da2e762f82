// Model-prediction producer: IndexMap::combinedPredict (src/model/IndexMap.cpp:243-341, shaders
// model/shaders/splat.vert + combo_splat.frag) and the FillIn passes (src/gl/FillIn.cpp:68-198, shaders
// gl/shaders/fill_vertex.frag, fill_normal.frag, fill_rgb.frag) as two CUDA launches on linear HBM buffers.
//
//   k_splat     vertex stage per surfel (one lane each), then the fragments of the warp's 32 point sprites flattened over
//               the lanes: ray / surfel-disc intersection per covered pixel and GL's depth test as a 64-bit atomicMin on
//               (depth24 << 32 | surfel index).  Per-pixel view rays come from a table built once per handle.
//   k_resolve   one thread per pixel: the winning surfel's fragment outputs (recomputed with the same device functions),
//               optionally the three fill-in passes behind them, and the z-buffer word re-armed for the next frame.
//   k_fill      the FillIn passes on caller-supplied `existing` textures (operator-level entry points).
//
// This translation unit is compiled with IEEE division / square root and without FMA contraction (build.py), so that the
// CPU restatement the tests check against (compiled with -ffp-contract=off) reproduces it bit for bit; GLSL leaves the precision of
// these operations to the implementation.
//
// Rasterisation rules where GL is implementation defined or silent (the CPU restatement follows the same list):
//   R1  window position of the point = projectPointImage (fx x / z + cx, fy y / z + cy); the clip test is done on the
//       NDC coordinates of projectPoint, point discarded when its centre is outside [-1, 1] (GL clips points by centre);
//       the z clip never triggers after the shader's own cull (0 <= z / maxDepth <= 1).
//   R2  point size = clamp(gl_PointSize, 1, max_point_size); a fragment is generated for each pixel whose centre
//       (px + 0.5, py + 0.5) lies in [xw - s/2, xw + s/2) x [yw - s/2, yw + s/2), inside the viewport.
//   R3  depth buffer = 24-bit fixed point (Pangolin's GlRenderBuffer default GL_DEPTH_COMPONENT24), gl_FragDepth clamped
//       to [0, 1] and converted by round(d * (2^24 - 1)); GL_LESS against a buffer cleared to 1.0; equal depths keep the
//       fragment drawn first, i.e. the lower surfel index (glDrawTransformFeedback draws in buffer order).
//   R4  fragments whose depth is NaN (ray parallel to the surfel plane) are dropped.
//   R5  normalize(v) = v / sqrt(dot(v, v)); mat * vec sums left to right; decodeColor's /255 and the RGBA8 store's
//       round(c * 255) cancel exactly, so the image bytes are the three bytes of int(colour); timeTex keeps the low
//       16 bits of uint(colTime.z).
//   R6  texture fetches of the fill passes: NEAREST, CLAMP_TO_EDGE (Pangolin GlTexture defaults), texel = pixel.
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../include/slam_predict.h"
#include "common.cuh"
#include "small_math.hpp"

namespace slam {

constexpr unsigned long long kZClear = 0x00FFFFFFFFFFFFFFull;   // depth 1.0 (0xFFFFFF), no surfel
constexpr unsigned kDepthOne = 0x00FFFFFFu;

struct PredictCam
{
    float cx, cy, fx, fy;
    float cols, rows;
    float max_point;
    int W, H;
};

struct PredictCall
{
    float tinv[16];            // row-major
    float max_depth, conf_threshold;
    int time, max_time, time_delta;
};

struct SurfelView
{
    float px, py, pz, conf;    // position = vec4(vPosHome.xyz, vPosition.w)
    float nx, ny, nz, rad;     // normRad
};

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) { return (ax * bx + ay * by) + az * bz; }

// splat.vert:52-65: cull + transform.  Returns false for a culled surfel.
__device__ __forceinline__ bool surfel_view(const float4 pos, const float4 col, const float4 nr, const PredictCall & c, SurfelView & s)
{
    const float * m = c.tinv;
    s.px = ((m[0] * pos.x + m[1] * pos.y) + m[2] * pos.z) + m[3];
    s.py = ((m[4] * pos.x + m[5] * pos.y) + m[6] * pos.z) + m[7];
    s.pz = ((m[8] * pos.x + m[9] * pos.y) + m[10] * pos.z) + m[11];
    if(s.pz > c.max_depth || s.pz < 0.f || pos.w < c.conf_threshold || (float)c.time - col.w > (float)c.time_delta || col.w > (float)c.max_time) return false;
    s.conf = pos.w;
    const float rx = (m[0] * nr.x + m[1] * nr.y) + m[2] * nr.z;
    const float ry = (m[4] * nr.x + m[5] * nr.y) + m[6] * nr.z;
    const float rz = (m[8] * nr.x + m[9] * nr.y) + m[10] * nr.z;
    const float len = sqrtf(dot3(rx, ry, rz, rx, ry, rz));
    s.nx = rx / len;
    s.ny = ry / len;
    s.nz = rz / len;
    s.rad = nr.w;
    return true;
}

// splat.vert:56 (gl_Position, rule R1) and :67-86 (gl_PointSize, rule R2).  Returns false when the point is clipped.
// `quad` receives the bounding box (xmin, xmax, ymin, ymax) of the four projected points; quad_ok tells whether all four lie in
// front of the camera, in which case the projected disc is inside that box (the four points are the corners of the square
// circumscribing the disc, and a perspective projection maps the convex hull of points in front of the camera onto the hull
// of their projections).
__device__ __forceinline__ bool surfel_sprite(const SurfelView & s, const PredictCam & k, float & xw, float & yw, float & size, float4 & quad, bool & quad_ok)
{
    xw = (k.fx * s.px) / s.pz + k.cx;
    yw = (k.fy * s.py) / s.pz + k.cy;
    const float ndx = (xw - k.cols * 0.5f) / (k.cols * 0.5f);
    const float ndy = (yw - k.rows * 0.5f) / (k.rows * 0.5f);
    if(!(ndx >= -1.f && ndx <= 1.f && ndy >= -1.f && ndy <= 1.f)) return false;
    // x1 = normalize(vec3(n.y - n.z, -n.x, n.x)) * rad * 1.41421356;  y1 = cross(n, x1)
    const float ax = s.ny - s.nz, ay = -s.nx, az = s.nx;
    const float al = sqrtf(dot3(ax, ay, az, ax, ay, az));
    const float x1x = ((ax / al) * s.rad) * 1.41421356f, x1y = ((ay / al) * s.rad) * 1.41421356f, x1z = ((az / al) * s.rad) * 1.41421356f;
    const float y1x = s.ny * x1z - s.nz * x1y, y1y = s.nz * x1x - s.nx * x1z, y1z = s.nx * x1y - s.ny * x1x;
    float xmin, xmax, ymin, ymax, zmin;
    {
        const float qx = s.px + x1x, qy = s.py + x1y, qz = s.pz + x1z;
        xmin = xmax = (k.fx * qx) / qz + k.cx;
        ymin = ymax = (k.fy * qy) / qz + k.cy;
        zmin = qz;
    }
    {
        const float qx = s.px + y1x, qy = s.py + y1y, qz = s.pz + y1z;
        const float u = (k.fx * qx) / qz + k.cx, v = (k.fy * qy) / qz + k.cy;
        xmin = fminf(xmin, u), xmax = fmaxf(xmax, u), ymin = fminf(ymin, v), ymax = fmaxf(ymax, v), zmin = fminf(zmin, qz);
    }
    {
        const float qx = s.px - y1x, qy = s.py - y1y, qz = s.pz - y1z;
        const float u = (k.fx * qx) / qz + k.cx, v = (k.fy * qy) / qz + k.cy;
        xmin = fminf(xmin, u), xmax = fmaxf(xmax, u), ymin = fminf(ymin, v), ymax = fmaxf(ymax, v), zmin = fminf(zmin, qz);
    }
    {
        const float qx = s.px - x1x, qy = s.py - x1y, qz = s.pz - x1z;
        const float u = (k.fx * qx) / qz + k.cx, v = (k.fy * qy) / qz + k.cy;
        xmin = fminf(xmin, u), xmax = fmaxf(xmax, u), ymin = fminf(ymin, v), ymax = fmaxf(ymax, v), zmin = fminf(zmin, qz);
    }
    quad = make_float4(xmin, xmax, ymin, ymax);
    // all four finite and in front of the camera (a NaN anywhere fails the comparison chain)
    quad_ok = zmin > 0.f && (xmax - xmin) < 1e6f && (ymax - ymin) < 1e6f;
    const float xd = fabsf(xmax - xmin), yd = fabsf(ymax - ymin);
    size = fmaxf(0.f, fmaxf(xd, yd));
    size = fminf(fmaxf(size, 1.f), k.max_point);
    return true;
}

// combo_splat.frag:35-46,60 with the view ray l of the pixel; returns the 24-bit depth, kDepthOne + 1 for a discarded fragment.
__device__ __forceinline__ unsigned fragment_depth(const float4 l, float px, float py, float pz, float nx, float ny, float nz, float pn, float rad2, float two_max,
                                                   float & z)
{
    const float ln = dot3(l.x, l.y, l.z, nx, ny, nz);
    const float t = pn / ln;
    const float qx = t * l.x, qy = t * l.y, qz = t * l.z;
    const float dx = qx - px, dy = qy - py, dz = qz - pz;
    const float dd = dot3(dx, dy, dz, dx, dy, dz);
    z = qz;
    if(dd > rad2) return kDepthOne + 1u;
    float d = qz / two_max + 0.5f;
    if(!(d == d)) return kDepthOne + 1u;                       // R4
    d = fminf(fmaxf(d, 0.f), 1.f);
    return (unsigned)__float2uint_rn(d * 16777215.f);          // R3 (exact product: d has 24 significant bits at most ... rounded to nearest)
}

// vec3 l = normalize(vec3((gl_FragCoord.xy - cam.xy) / cam.zw, 1))   combo_splat.frag:35
__global__ void __launch_bounds__(256) k_ray_table(PredictCam k, float4 * __restrict__ rays)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= k.W * k.H) return;
    const int y = i / k.W, x = i - y * k.W;
    const float lx = (((float)x + 0.5f) - k.cx) / k.fx;
    const float ly = (((float)y + 0.5f) - k.cy) / k.fy;
    const float len = sqrtf(dot3(lx, ly, 1.f, lx, ly, 1.f));
    rays[i] = make_float4(lx / len, ly / len, 1.f / len, 0.f);
}

__global__ void __launch_bounds__(256) k_zclear(unsigned long long * __restrict__ z, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) z[i] = kZClear;
}

constexpr int kSplatWarps = 4;   // 128-thread blocks: finer tail than 256 (measured 40 vs 44 us on a 307 k-surfel model, equal on large ones)
constexpr int kRecWords = 12;
constexpr float kQuadPad = 0.01f;
constexpr int kFragsPerTrip = 2;   // measured on B200: 1..4 within 5 % of each other (the loop is issue bound), 2 and 3 best

// Pixel index of fragment `local` of a sprite record (row-major inside its bounding box).  The row comes from a float
// reciprocal multiply and is corrected by at most one, which is exact for boxes up to 2^26 pixels.
__device__ __forceinline__ int sprite_pixel(const float * r, int local, int W)
{
    const int w = __float_as_int(r[10]);
    int dy = (int)(((float)local + 0.5f) * r[11]);
    int dx = local - dy * w;
    if(dx < 0)
        dy--, dx += w;
    else if(dx >= w)
        dy++, dx -= w;
    return (__float_as_int(r[9]) + dy) * W + __float_as_int(r[8]) + dx;
}

// One warp per group of 32 consecutive surfels (grid-stride).  Shared record per sprite: view position, normal, dot(p, n),
// rad^2, bounding box origin / width; the inclusive prefix of the box areas drives the flattened fragment loop.
__global__ void __launch_bounds__(kSplatWarps * 32) k_splat(const float4 * __restrict__ surfels, int count, PredictCam k, PredictCall c,
                                                            const float4 * __restrict__ rays, unsigned long long * __restrict__ zbuf,
                                                            unsigned long long * __restrict__ frag_counter)
{
    __shared__ float rec[kSplatWarps][32][kRecWords];
    __shared__ int prefix[kSplatWarps][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ngroups = (count + 31) >> 5;
    const float two_max = 2.f * c.max_depth;
    unsigned long long my_frags = 0;
    for(int g = blockIdx.x * kSplatWarps + wid; g < ngroups; g += gridDim.x * kSplatWarps)
    {
        const int i = g * 32 + lane;
        int area = 0;
        if(i < count)
        {
            const float4 pos = __ldcs(surfels + 3 * (size_t)i), col = __ldcs(surfels + 3 * (size_t)i + 1), nr = __ldcs(surfels + 3 * (size_t)i + 2);
            SurfelView s;
            float xw, yw, size;
            float4 quad;
            bool quad_ok;
            if(surfel_view(pos, col, nr, c, s) && surfel_sprite(s, k, xw, yw, size, quad, quad_ok))
            {
                const float h = size * 0.5f;
                // pixels with xw - h <= px + 0.5 < xw + h
                int x0 = (int)ceilf((xw - h) - 0.5f), x1 = (int)ceilf((xw + h) - 0.5f) - 1;
                int y0 = (int)ceilf((yw - h) - 0.5f), y1 = (int)ceilf((yw + h) - 0.5f) - 1;
                x0 = max(x0, 0), y0 = max(y0, 0), x1 = min(x1, k.W - 1), y1 = min(y1, k.H - 1);
                if(quad_ok)
                {
                    // Fragments of the sprite square outside the projected quad can only fail the disc test: skip them.  The box is
                    // widened by kQuadPad pixels, two orders of magnitude above the fp32 rounding of the projections (~1e-4 px).
                    x0 = max(x0, (int)ceilf((quad.x - kQuadPad) - 0.5f)), x1 = min(x1, (int)ceilf((quad.y + kQuadPad) - 0.5f) - 1);
                    y0 = max(y0, (int)ceilf((quad.z - kQuadPad) - 0.5f)), y1 = min(y1, (int)ceilf((quad.w + kQuadPad) - 0.5f) - 1);
                }
                if(x1 >= x0 && y1 >= y0)
                {
                    const int w = x1 - x0 + 1;
                    area = w * (y1 - y0 + 1);
                    float * r = rec[wid][lane];
                    r[0] = s.px, r[1] = s.py, r[2] = s.pz;
                    r[3] = s.nx, r[4] = s.ny, r[5] = s.nz;
                    r[6] = dot3(s.px, s.py, s.pz, s.nx, s.ny, s.nz);
                    r[7] = s.rad * s.rad;
                    r[8] = __int_as_float(x0), r[9] = __int_as_float(y0), r[10] = __int_as_float(w), r[11] = 1.0f / (float)w;
                }
            }
        }
        int incl = area;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if(lane >= o) incl += v;
        }
        prefix[wid][lane + 1] = incl;
        if(lane == 0) prefix[wid][0] = 0;
        __syncwarp();
        const int total = prefix[wid][32];
        my_frags += (lane == 0) ? (unsigned long long)total : 0ull;
        // Flattened fragment list, kFragsPerTrip fragments per lane and trip (f, f + 32, ...) so that their ray / z-buffer loads
        // overlap.  Slots past the end repeat the last fragment: atomicMin with an equal key changes nothing.
        int j = 0;
        int jbeg = 0, jend = prefix[wid][1];
        for(int f0 = lane; f0 < total; f0 += 32 * kFragsPerTrip)
        {
            const float * r[kFragsPerTrip];
            int pix[kFragsPerTrip];
            unsigned id[kFragsPerTrip];
#pragma unroll
            for(int u = 0; u < kFragsPerTrip; u++)
            {
                const int f = min(f0 + 32 * u, total - 1);
                while(f >= jend)
                {
                    j++;
                    jbeg = jend;
                    jend = prefix[wid][j + 1];
                }
                r[u] = rec[wid][j];
                id[u] = (unsigned)(g * 32 + j);
                pix[u] = sprite_pixel(r[u], f - jbeg, k.W);
            }
            float4 ray[kFragsPerTrip];
            unsigned long long zcur[kFragsPerTrip];
#pragma unroll
            for(int u = 0; u < kFragsPerTrip; u++)
            {
                ray[u] = __ldg(rays + pix[u]);
                zcur[u] = __ldcg(zbuf + pix[u]);     // speculative: read before the disc test
            }
#pragma unroll
            for(int u = 0; u < kFragsPerTrip; u++)
            {
                float z;
                const float * q = r[u];
                const unsigned d24 = fragment_depth(ray[u], q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], two_max, z);
                const unsigned long long key = ((unsigned long long)d24 << 32) | id[u];
                if(d24 < kDepthOne && key < zcur[u]) atomicMin(zbuf + pix[u], key);
            }
        }
        __syncwarp();
    }
    if(frag_counter && lane == 0 && my_frags) atomicAdd(frag_counter, my_frags);
}

struct ResolveOut
{
    uchar4 * image;
    float4 * vertex;
    float4 * normal;
    unsigned short * time;
    uchar4 * fill_image;
    float4 * fill_vertex;
    float4 * fill_normal;
    unsigned long long * winners;   // copy of the resolved z-buffer (parity tap) or null
};

// fill_vertex.frag:36-39,51-52
__device__ __forceinline__ float4 raw_vertex(const unsigned short * __restrict__ depth, int x, int y, int tx, int ty, const PredictCam & k, float ifx, float ify)
{
    const float z = (float)depth[ty * k.W + tx] / 1000.0f;
    return make_float4((((float)x - k.cx) * z) * ifx, (((float)y - k.cy) * z) * ify, z, 1.f);
}

// geometry.glsl:43-61 (forward differences on the raw depth) as fill_normal.frag:48-49 uses them
__device__ __forceinline__ float4 raw_normal(const unsigned short * __restrict__ depth, int x, int y, const PredictCam & k, float ifx, float ify)
{
    const float4 v = raw_vertex(depth, x, y, x, y, k, ifx, ify);
    const float4 vx = raw_vertex(depth, x + 1, y, min(x + 1, k.W - 1), y, k, ifx, ify);
    const float4 vy = raw_vertex(depth, x, y + 1, x, min(y + 1, k.H - 1), k, ifx, ify);
    const float ax = vx.x - v.x, ay = vx.y - v.y, az = vx.z - v.z;
    const float bx = vy.x - v.x, by = vy.y - v.y, bz = vy.z - v.z;
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    const float len = sqrtf(dot3(cx, cy, cz, cx, cy, cz));
    return make_float4(cx / len, cy / len, cz / len, 1.f);
}

// mode bits: 1 = write the IndexMap textures, 2 = fill-in passes behind them (FillIn textures)
__global__ void __launch_bounds__(256) k_resolve(const float4 * __restrict__ surfels, PredictCam k, PredictCall c, unsigned long long * __restrict__ zbuf,
                                                 const float4 * __restrict__ rays, const unsigned short * __restrict__ raw_depth, const uchar4 * __restrict__ raw_rgba, ResolveOut o, int mode)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= k.W * k.H) return;
    const int y = p / k.W, x = p - y * k.W;
    const unsigned long long key = zbuf[p];
    zbuf[p] = kZClear;
    if(o.winners) o.winners[p] = key;
    uchar4 img = make_uchar4(0, 0, 0, 0);
    float4 vtx = make_float4(0.f, 0.f, 0.f, 0.f), nrm = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned short tm = 0;
    if((unsigned)(key >> 32) < kDepthOne)
    {
        const unsigned i = (unsigned)key;
        const float4 pos = surfels[3 * (size_t)i], col = surfels[3 * (size_t)i + 1], nr = surfels[3 * (size_t)i + 2];
        SurfelView s;
        surfel_view(pos, col, nr, c, s);
        const float fxc = (float)x + 0.5f, fyc = (float)y + 0.5f;
        const float4 l = rays[p];
        float z;
        fragment_depth(l, s.px, s.py, s.pz, s.nx, s.ny, s.nz, dot3(s.px, s.py, s.pz, s.nx, s.ny, s.nz), s.rad * s.rad, 2.f * c.max_depth, z);
        const int rgb = (int)col.x;
        img = make_uchar4((rgb >> 16) & 0xFF, (rgb >> 8) & 0xFF, rgb & 0xFF, 255);
        vtx = make_float4(((fxc - k.cx) * z) * (1.f / k.fx), ((fyc - k.cy) * z) * (1.f / k.fy), z, s.conf);
        nrm = make_float4(s.nx, s.ny, s.nz, s.rad);
        tm = (unsigned short)(unsigned)col.z;
    }
    if(mode & 1)
    {
        o.image[p] = img;
        o.vertex[p] = vtx;
        o.normal[p] = nrm;
        o.time[p] = tm;
    }
    if(mode & 2)
    {
        const float ifx = 1.0f / k.fx, ify = 1.0f / k.fy;
        if(vtx.z == 0.f) vtx = raw_vertex(raw_depth, x, y, x, y, k, ifx, ify);
        if(nrm.z == 0.f) nrm = raw_normal(raw_depth, x, y, k, ifx, ify);
        if((int)img.x + (int)img.y + (int)img.z == 0) img = raw_rgba[p];
        o.fill_vertex[p] = vtx;
        o.fill_normal[p] = nrm;
        o.fill_image[p] = img;
    }
}

// FillIn::vertex / normal / image on caller-supplied textures.  which bits: 1 vertex, 2 normal, 4 image.
__global__ void __launch_bounds__(256) k_fill(PredictCam k, const float4 * __restrict__ ex_vertex, const float4 * __restrict__ ex_normal,
                                              const uchar4 * __restrict__ ex_image, const unsigned short * __restrict__ raw_depth,
                                              const uchar4 * __restrict__ raw_rgba, float4 * __restrict__ out_vertex, float4 * __restrict__ out_normal,
                                              uchar4 * __restrict__ out_image, int which, int passthrough)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if(p >= k.W * k.H) return;
    const int y = p / k.W, x = p - y * k.W;
    const float ifx = 1.0f / k.fx, ify = 1.0f / k.fy;
    if(which & 1)
    {
        float4 v = ex_vertex[p];
        if(v.z == 0.f || passthrough) v = raw_vertex(raw_depth, x, y, x, y, k, ifx, ify);
        out_vertex[p] = v;
    }
    if(which & 2)
    {
        float4 n = ex_normal[p];
        if(n.z == 0.f || passthrough) n = raw_normal(raw_depth, x, y, k, ifx, ify);
        out_normal[p] = n;
    }
    if(which & 4)
    {
        uchar4 c = ex_image[p];
        if((int)c.x + (int)c.y + (int)c.z == 0 || passthrough) c = raw_rgba[p];
        out_image[p] = c;
    }
}

}   // namespace slam

using namespace slam;

struct slam_predict
{
    slam_predict_params p{};
    PredictCam cam{};
    PredictCall call{};
    bool have_call = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int sm_count = 148;
    float4 * rays = nullptr;
    unsigned long long * zbuf = nullptr, * winners = nullptr, * frag_counter = nullptr;
    uchar4 * image = nullptr, * fill_image = nullptr;
    float4 * vertex = nullptr, * normal = nullptr, * fill_vertex = nullptr, * fill_normal = nullptr;
    unsigned short * time = nullptr;
    // IndexMap's oldFrameBuffer attachments (INACTIVE prediction)
    uchar4 * old_image = nullptr;
    float4 * old_vertex = nullptr, * old_normal = nullptr;
    unsigned short * old_time = nullptr;
    bool timed = false;
    bool zdirty = false;   // a call failed between the splat and the resolve launch: the z-buffer may hold keys of that call
};

static int predict_check(slam_predict_t h)
{
    if(!h)
    {
        set_last_error("null handle");
        return SLAM_ERR_ARG;
    }
    return SLAM_OK;
}

// streams, events, buffers and the one-off tables of a new handle; on failure the caller destroys the partly built handle
static int predict_setup(slam_predict * h, const slam_predict_params * params)
{
    if(params->stream)
        h->stream = (cudaStream_t)params->stream;
    else
    {
        SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    for(auto & e : h->ev) SLAM_CUDA_TRY(cudaEventCreate(&e));
    const size_t n = (size_t)params->width * params->height;
    SLAM_CUDA_TRY(cudaMalloc(&h->rays, n * sizeof(float4)));
    SLAM_CUDA_TRY(cudaMalloc(&h->zbuf, n * 8));
    SLAM_CUDA_TRY(cudaMalloc(&h->winners, n * 8));
    SLAM_CUDA_TRY(cudaMalloc(&h->frag_counter, 8));
    SLAM_CUDA_TRY(cudaMalloc(&h->image, n * 4));
    SLAM_CUDA_TRY(cudaMalloc(&h->fill_image, n * 4));
    SLAM_CUDA_TRY(cudaMalloc(&h->vertex, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->normal, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->fill_vertex, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->fill_normal, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->time, n * 2));
    SLAM_CUDA_TRY(cudaMalloc(&h->old_image, n * 4));
    SLAM_CUDA_TRY(cudaMalloc(&h->old_vertex, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->old_normal, n * 16));
    SLAM_CUDA_TRY(cudaMalloc(&h->old_time, n * 2));
    const int nb = div_up((int)n, 256);
    k_ray_table<<<nb, 256, 0, h->stream>>>(h->cam, h->rays);
    k_zclear<<<nb, 256, 0, h->stream>>>(h->zbuf, (int)n);
    SLAM_CUDA_TRY(cudaMemsetAsync(h->frag_counter, 0, 8, h->stream));
    for(void * b : {(void *)h->image, (void *)h->fill_image, (void *)h->old_image}) SLAM_CUDA_TRY(cudaMemsetAsync(b, 0, n * 4, h->stream));
    for(void * b : {(void *)h->vertex, (void *)h->normal, (void *)h->fill_vertex, (void *)h->fill_normal, (void *)h->old_vertex, (void *)h->old_normal})
        SLAM_CUDA_TRY(cudaMemsetAsync(b, 0, n * 16, h->stream));
    SLAM_CUDA_TRY(cudaMemsetAsync(h->time, 0, n * 2, h->stream));
    SLAM_CUDA_TRY(cudaMemsetAsync(h->old_time, 0, n * 2, h->stream));
    SLAM_CUDA_TRY(cudaMemsetAsync(h->winners, 0xFF, n * 8, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_predict_create(const slam_predict_params * params, slam_predict_t * out)
{
    SLAM_ARG_CHECK(params && out);
    SLAM_ARG_CHECK(params->width > 0 && params->height > 0 && params->width <= 8192 && params->height <= 8192);
    SLAM_ARG_CHECK((size_t)params->width * params->height <= ((size_t)1 << 24));   // a warp's flattened fragment list (32 sprites) is indexed with an int
    SLAM_ARG_CHECK(params->fx != 0.f && params->fy != 0.f);
    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    {
        set_last_error("no CUDA device: libslam_odom has no CPU fallback");
        return SLAM_ERR_CUDA;
    }
    SLAM_ARG_CHECK(params->device >= 0 && params->device < ndev);
    SLAM_CUDA_TRY(cudaSetDevice(params->device));
    slam_predict * h = new slam_predict();
    h->p = *params;
    h->cam = PredictCam{params->cx, params->cy, params->fx, params->fy, (float)params->width, (float)params->height,
                        params->max_point_size > 0.f ? params->max_point_size : 2047.f, params->width, params->height};
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, params->device);
    if(int rc = predict_setup(h, params))
    {
        slam_predict_destroy(h);
        return rc;
    }
    *out = h;
    return SLAM_OK;
}

extern "C" int slam_predict_destroy(slam_predict_t h)
{
    if(!h) return SLAM_OK;
    cudaSetDevice(h->p.device);
    // a caller-supplied stream may already be gone (its owner was destroyed first): only our own stream is synchronised here,
    // cudaFree below waits for outstanding work on the buffers either way
    if(h->own_stream && h->stream) cudaStreamSynchronize(h->stream);
    for(void * b : {(void *)h->rays, (void *)h->zbuf, (void *)h->winners, (void *)h->frag_counter, (void *)h->image, (void *)h->fill_image, (void *)h->vertex,
                    (void *)h->normal, (void *)h->fill_vertex, (void *)h->fill_normal, (void *)h->time, (void *)h->old_image, (void *)h->old_vertex,
                    (void *)h->old_normal, (void *)h->old_time})
        cudaFree(b);
    for(auto & e : h->ev)
        if(e) cudaEventDestroy(e);
    if(h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SLAM_OK;
}

extern "C" int slam_predict_get_textures(slam_predict_t h, slam_predict_textures * out)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(out);
    out->image = reinterpret_cast<const uint8_t *>(h->image);
    out->vertex = reinterpret_cast<const float *>(h->vertex);
    out->normal = reinterpret_cast<const float *>(h->normal);
    out->time = h->time;
    out->fill_image = reinterpret_cast<const uint8_t *>(h->fill_image);
    out->fill_vertex = reinterpret_cast<const float *>(h->fill_vertex);
    out->fill_normal = reinterpret_cast<const float *>(h->fill_normal);
    out->old_image = reinterpret_cast<const uint8_t *>(h->old_image);
    out->old_vertex = reinterpret_cast<const float *>(h->old_vertex);
    out->old_normal = reinterpret_cast<const float *>(h->old_normal);
    out->old_time = h->old_time;
    return SLAM_OK;
}

// Eigen::Matrix4f t_inv = pose.inverse()   IndexMap.cpp:284
static void set_call(slam_predict * h, const float * pose16, float depth_cutoff, float conf_threshold, int time, int max_time, int time_delta)
{
    smath::mat4_inverse<float>(pose16, h->call.tinv);
    h->call.max_depth = depth_cutoff;
    h->call.conf_threshold = conf_threshold;
    h->call.time = time;
    h->call.max_time = max_time;
    h->call.time_delta = time_delta;
    h->have_call = true;
}

static int run_predict(slam_predict * h, const float * d_surfels, int count, const uint16_t * d_raw_depth, const uint8_t * d_raw_rgba, int mode, bool inactive = false)
{
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    const int n = h->cam.W * h->cam.H;
    const float4 * surfels = reinterpret_cast<const float4 *>(d_surfels);
    SLAM_CUDA_TRY(cudaMemsetAsync(h->frag_counter, 0, 8, h->stream));
    if(h->zdirty) k_zclear<<<div_up(n, 256), 256, 0, h->stream>>>(h->zbuf, n);
    h->zdirty = true;   // until the resolve launch below has re-armed the buffer
    SLAM_CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    if(count > 0)
    {
        const int ngroups = div_up(count, 32);
        const int blocks = std::min(div_up(ngroups, kSplatWarps), h->sm_count * 16);
        k_splat<<<blocks, kSplatWarps * 32, 0, h->stream>>>(surfels, count, h->cam, h->call, h->rays, h->zbuf, h->frag_counter);
    }
    SLAM_CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    ResolveOut o{h->image, h->vertex, h->normal, h->time, h->fill_image, h->fill_vertex, h->fill_normal, h->winners};
    if(inactive) o.image = h->old_image, o.vertex = h->old_vertex, o.normal = h->old_normal, o.time = h->old_time;
    k_resolve<<<div_up(n, 256), 256, 0, h->stream>>>(surfels, h->cam, h->call, h->zbuf, h->rays, d_raw_depth, reinterpret_cast<const uchar4 *>(d_raw_rgba), o, mode);
    SLAM_CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
    SLAM_CUDA_TRY(cudaGetLastError());
    h->zdirty = false;
    h->timed = true;
    return SLAM_OK;
}

extern "C" int slam_predict_combined(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold,
                                     int time, int max_time, int time_delta)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(pose16 && count >= 0 && (d_surfels || count == 0) && depth_cutoff > 0.f);
    set_call(h, pose16, depth_cutoff, conf_threshold, time, max_time, time_delta);
    return run_predict(h, d_surfels, count, nullptr, nullptr, 1);
}

extern "C" int slam_predict_combined_type(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold,
                                          int time, int max_time, int time_delta, int prediction_type)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(pose16 && count >= 0 && (d_surfels || count == 0) && depth_cutoff > 0.f);
    SLAM_ARG_CHECK(prediction_type == SLAM_PREDICT_ACTIVE || prediction_type == SLAM_PREDICT_INACTIVE);   // the reference asserts
    set_call(h, pose16, depth_cutoff, conf_threshold, time, max_time, time_delta);
    return run_predict(h, d_surfels, count, nullptr, nullptr, 1, prediction_type == SLAM_PREDICT_INACTIVE);
}

extern "C" int slam_predict_frame(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold,
                                  int time, int max_time, int time_delta, const uint16_t * d_raw_depth, const uint8_t * d_raw_rgba, int write_index_textures)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(pose16 && count >= 0 && (d_surfels || count == 0) && depth_cutoff > 0.f && d_raw_depth && d_raw_rgba);
    set_call(h, pose16, depth_cutoff, conf_threshold, time, max_time, time_delta);
    return run_predict(h, d_surfels, count, d_raw_depth, d_raw_rgba, 2 | (write_index_textures ? 1 : 0));
}

static int run_fill(slam_predict * h, const float * ex_v, const float * ex_n, const uint8_t * ex_i, const uint16_t * depth, const uint8_t * rgba, int which,
                    int passthrough)
{
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    const int n = h->cam.W * h->cam.H;
    k_fill<<<div_up(n, 256), 256, 0, h->stream>>>(h->cam, reinterpret_cast<const float4 *>(ex_v), reinterpret_cast<const float4 *>(ex_n),
                                                 reinterpret_cast<const uchar4 *>(ex_i), depth, reinterpret_cast<const uchar4 *>(rgba), h->fill_vertex,
                                                 h->fill_normal, h->fill_image, which, passthrough ? 1 : 0);
    SLAM_CUDA_TRY(cudaGetLastError());
    return SLAM_OK;
}

extern "C" int slam_predict_fill_vertex(slam_predict_t h, const float * d_existing_vertex4, const uint16_t * d_raw_depth, int passthrough)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(d_raw_depth);
    return run_fill(h, d_existing_vertex4 ? d_existing_vertex4 : reinterpret_cast<const float *>(h->vertex), nullptr, nullptr, d_raw_depth, nullptr, 1, passthrough);
}

extern "C" int slam_predict_fill_normal(slam_predict_t h, const float * d_existing_normal4, const uint16_t * d_raw_depth, int passthrough)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(d_raw_depth);
    return run_fill(h, nullptr, d_existing_normal4 ? d_existing_normal4 : reinterpret_cast<const float *>(h->normal), nullptr, d_raw_depth, nullptr, 2, passthrough);
}

extern "C" int slam_predict_fill_image(slam_predict_t h, const uint8_t * d_existing_rgba, const uint8_t * d_raw_rgba, int passthrough)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(d_raw_rgba);
    return run_fill(h, nullptr, nullptr, d_existing_rgba ? d_existing_rgba : reinterpret_cast<const uint8_t *>(h->image), nullptr, d_raw_rgba, 4, passthrough);
}

extern "C" int slam_predict_download(slam_predict_t h, int texture, void * host_out)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(host_out && texture >= 0 && texture <= 10);
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->cam.W * h->cam.H;
    const void * src[11] = {h->image, h->vertex, h->normal, h->time, h->fill_image, h->fill_vertex, h->fill_normal, h->old_image, h->old_vertex, h->old_normal, h->old_time};
    const size_t texel[11] = {4, 16, 16, 2, 4, 16, 16, 4, 16, 16, 2};
    SLAM_CUDA_TRY(cudaMemcpyAsync(host_out, src[texture], n * texel[texture], cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SLAM_OK;
}

extern "C" int slam_predict_get_tinv(slam_predict_t h, float * tinv16)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(tinv16 && h->have_call);
    memcpy(tinv16, h->call.tinv, sizeof(float) * 16);
    return SLAM_OK;
}

extern "C" int slam_predict_get_winners(slam_predict_t h, uint32_t * depth24, int32_t * surfel)
{
    if(int e = predict_check(h)) return e;
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->cam.W * h->cam.H;
    std::vector<unsigned long long> keys(n);
    SLAM_CUDA_TRY(cudaMemcpyAsync(keys.data(), h->winners, n * 8, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    for(size_t i = 0; i < n; i++)
    {
        const uint32_t d = (uint32_t)(keys[i] >> 32);
        if(depth24) depth24[i] = d >= kDepthOne ? kDepthOne : d;
        if(surfel) surfel[i] = d >= kDepthOne ? -1 : (int32_t)(uint32_t)keys[i];
    }
    return SLAM_OK;
}

extern "C" int slam_predict_last_ms(slam_predict_t h, float * splat_ms, float * resolve_ms)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(h->timed);
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    SLAM_CUDA_TRY(cudaEventSynchronize(h->ev[2]));
    if(splat_ms) SLAM_CUDA_TRY(cudaEventElapsedTime(splat_ms, h->ev[0], h->ev[1]));
    if(resolve_ms) SLAM_CUDA_TRY(cudaEventElapsedTime(resolve_ms, h->ev[1], h->ev[2]));
    return SLAM_OK;
}

extern "C" int slam_predict_last_fragments(slam_predict_t h, unsigned long long * fragments)
{
    if(int e = predict_check(h)) return e;
    SLAM_ARG_CHECK(fragments);
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    SLAM_CUDA_TRY(cudaMemcpyAsync(fragments, h->frag_counter, 8, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SLAM_OK;
}
