// Device-resident Gauss-Newton loop of the tracker: ONE persistent, cooperatively launched kernel runs the SO3
// pre-alignment and the coarse-to-fine ICP+RGB iterations of RGBDOdometryef::getIncrementalTransformation
// (src/odom/RGBDOdometryef.cpp:267-595), including the 3x3 / 6x6 solves and the pose updates that the reference does on
// the host between ~40 kernel launches, ~25 cudaDeviceSynchronize and ~20 cudaMalloc/cudaFree pairs per frame.
//
// A frame is a chain of ~30 reductions, each followed by a solve whose result the next map needs: what bounds it is the
// latency of that chain, not bytes (the 45 MB working set of a 640x480 frame sits in the 126 MB L2).  Layout of the computation:
//   * the CTAs of the grid are split into groups of G CTAs; a group owns one sequence (batch == 1: one group of all 148
//     CTAs; batch > 1: independent groups, no inter-group traffic);
//   * STAGING (once per frame and level): a CTA owns every P-th 32-pixel segment of a level and keeps the pose-independent
//     operands of its pixels in shared memory for the whole frame -- current vertex + normal for ICP (reduce.cu:282-283), and
//     for RGB the outcome of the pose-independent half of the association (reduce.cu:780-807: window test, gradient
//     threshold, finite depth) with its operands.  Pixels that cannot take part are dropped while staging (ballot + prefix
//     sum -> dense lists), so the iterations never spend a lane on a dead pixel and the per-CTA work is balanced;
//   * an ICP+RGB iteration is: RGB association (list -> gathers -> count) | post the count | ICP products (list -> gathers
//     -> 29 sums) | block reduction + post | read the global count (its trip through L2 overlapped the ICP map) -> sigma |
//     RGB products from the correspondences kept in shared memory | block reduction + post | read all sums | solve;
//   * INTER-CTA ALL-REDUCE: fixed-point words in L2 that carry their own arrival count (gn_kernel.cuh): one atomic add per word
//     and CTA, the readers poll the words; no fence, no fold, deterministic sums;
//   * warp 0 of EVERY CTA then solves the normal equations redundantly (gn_fast_math.cuh), so the next parameters are in
//     every CTA's shared memory without a broadcast;
//   * small levels (and the SO3 pre-alignment, whose two 160x120 images are staged in shared memory) run on the first P CTAs
//     only: fewer arrivals per reduction; the other CTAs just follow the sums.
// Levels whose lists do not fit shared memory (large images, small groups) stream their operands from L2 per iteration.
// Per-pixel arithmetic is pixel_ops.cuh, shared with the single-launch operator kernels.
#include <cfloat>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include "gn_kernel.cuh"
#include "gn_scalar.cuh"
#include "gn_fast_math.cuh"

namespace slam {

// Unroll factors are kept small on purpose: the per-iteration code of the whole CTA has to stay inside the 32 KB instruction
// cache (ncu: 18 % of the instruction-cache requests missed and every phase started with an instruction-fetch stall when it did not).
constexpr int kIcpChunk = 1;      // ICP entries in flight per thread
constexpr int kRgbChunk = 1;      // RGB entries in flight per thread
constexpr int kMaxStageSlots = 16;   // 32-pixel segments a warp stages per level at most (resident levels)
constexpr int kSpinCap = 1 << 21; // polls before a reader gives up (~1 s)
constexpr double kFracScale = 281474976710656.0;   // 2^48

// The buffers of one pyramid level of one sequence.
struct LevelPtrs
{
    const float * vcurr, * ncurr, * vprev, * nprev, * lastDepth, * nextDepth;
    const unsigned char * lastImage, * nextImage, * lastNextImage;
    const short * dIdx, * dIdy;
    Corres * corres;
};

struct GnWork
{
    // the running level: buffers and this CTA's lists.  Kept in shared memory and re-read by every phase, so that none of it
    // occupies registers across the phases of an iteration (the map loops need them all)
    LevelPtrs P;
    int lv_n_icp, lv_n_rgb;
    int n_icp[SLAM_MAX_LEVELS], n_rgb[SLAM_MAX_LEVELS];   // list lengths of this CTA
    int scan_cnt[2][kMaxStageSlots * kGnWarps];   // staging: survivors per (segment slot, warp), then their list offsets
    int cnt[2];                    // RGB correspondence count / squared-residual sum of this CTA (shared-memory atomics)
    unsigned long long mid;        // global correspondence count word
    long long sigma;               // global squared-residual sum (rgbOnly: read at the mid point)
    int timeouts;
    int so3_stop[2];               // outcome of the SO3 iteration, by iteration parity (read by every warp after the barrier while warp 0 may already be in the next one)
    unsigned ph[24];               // cycles per phase (leading CTA, thread 0)
    unsigned long long base[kRingSlots * kRingWords];   // value of every reduction word when this CTA last consumed it
};

// ------------------------------------------------------------------ fixed-point all-reduce
__device__ __forceinline__ void red_u64(unsigned long long * p, unsigned long long v)
{
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_u64_relaxed(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long * ring_word(unsigned long long * ring, unsigned step, int w)
{
    return ring + ((size_t)(step & (kRingSlots - 1)) * kRingWords + w) * kWordStride;
}
// One CTA's fp32 partial sum of column c: integer part and fraction (both exact) go to the two words of the column.
__device__ __forceinline__ void post_float(unsigned long long * ring, unsigned step, int wbase, int c, float v)
{
    const float vi = rintf(v);
    const float vf = v - vi;
    const long long qi = __float2ll_rn(vi);
    const long long qf = __double2ll_rn((double)vf * kFracScale);
    red_u64(ring_word(ring, step, wbase + 2 * c), ((unsigned long long)qi << 8) + 1ull);
    red_u64(ring_word(ring, step, wbase + 2 * c + 1), ((unsigned long long)qf << 8) + 1ull);
}
__device__ __forceinline__ void post_int(unsigned long long * ring, unsigned step, int w, long long v)
{
    red_u64(ring_word(ring, step, w), ((unsigned long long)v << 8) + 1ull);
}
// The words are never cleared: every CTA remembers the value a word had when it last consumed it (wk.base, loaded from memory
// at kernel start) and takes differences -- of the 8-bit arrival count modulo 256, of the 56-bit sum modulo 2^56.  Every CTA
// consumes every word that is posted in a step, so the bases stay in step without any fence or clearing pass on the chain.
__device__ __forceinline__ bool word_complete(const unsigned long long v, const unsigned long long base, const unsigned want)
{
    return (((unsigned)v - (unsigned)base) & 0xffu) == want;
}
__device__ __forceinline__ long long word_consume(const unsigned long long v, unsigned long long & base, const unsigned want)
{
    // the arrivals are added to the same 64-bit word: when the 8-bit count wraps, its carry lands in the sum's lowest bit
    const long long carry = (long long)((((unsigned)base & 0xffu) + want) >> 8);
    const long long d = ((long long)((v & ~0xffull) - (base & ~0xffull)) >> 8) - carry;
    base = v;
    return d;
}
// Wait until `want` CTAs have contributed to the word; returns the (signed) sum of their contributions.
__device__ __forceinline__ long long poll_word(unsigned long long * ring, unsigned step, int w, unsigned want, GnWork & wk)
{
    const unsigned long long * p = ring_word(ring, step, w);
    unsigned long long & base = wk.base[(step & (kRingSlots - 1)) * kRingWords + w];
    const unsigned long long b = base;
    unsigned long long v;
    int spin = 0;
    do
    {
        v = ld_u64_relaxed(p);
    } while(!word_complete(v, b, want) && ++spin < kSpinCap);
    if(spin >= kSpinCap) wk.timeouts = 1;
    return word_consume(v, base, want);
}
__device__ __forceinline__ void spin_cycles(int cycles)
{
    if(cycles > 0)
    {
        const long long t = clock64();
        while(clock64() - t < cycles) {}
    }
}

// ------------------------------------------------------------------ all-reduce inside one thread-block cluster (split launch)
// Every CTA stores its partial row into the ClArea of every peer (distributed shared memory), all threads of the cluster pass one
// barrier.cluster (release / acquire), every CTA folds the rows in rank order: one trip through the SM-to-SM network instead of
// one through L2 (~1.3 k against ~2.5 k cycles, profiles/r02_micro_sync_latency.log), deterministic.
__device__ __forceinline__ unsigned cl_map(const void * p, const unsigned rank)
{
    unsigned a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void cl_st_f32(const unsigned a, const float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void cl_st_v2(const unsigned a, const int x, const int y) { asm volatile("st.shared::cluster.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// lane `col` of one warp: this CTA's total of a column goes to every peer
__device__ __forceinline__ void cl_post_col(ClArea * cl, const int par, const int rank, const int csize, const int col, const float total)
{
#pragma unroll 4
    for(int r = 0; r < csize; r++) cl_st_f32(cl_map(&cl->rows[par][rank][col], (unsigned)r), total);
}
// after the barrier: the sum of column `col` over the first `want` CTAs
__device__ __forceinline__ float cl_fold_col(const ClArea * cl, const int par, const int want, const int col)
{
    double s = 0.0;
#pragma unroll 4
    for(int r = 0; r < want; r++) s += (double)cl->rows[par][r][col];
    return (float)s;
}

// Block sum of up to 32 per-thread floats (v[ncols..31] must be 0), posted to the ncols columns starting at word wbase.
// The caller must reach a __syncthreads() before sh.red is reused.
// Reduce-scatter of 16 values per lane (v[16..31] unused): lanes 2c and 2c + 1 end with the warp-wide sum of value c.  16 shuffles
// instead of the 31 of the 32-wide form; fixed pattern.
template <int W>
__device__ __forceinline__ void warp_rs16_step(float (&v)[32], const int lane)
{
    const bool upper = (lane & (2 * W)) != 0;
#pragma unroll
    for(int i = 0; i < W; i++)
    {
        const float send = upper ? v[i] : v[i + W];
        const float keep = upper ? v[i + W] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * W);
    }
}
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[32])
{
    const int lane = threadIdx.x & 31;
    warp_rs16_step<8>(v, lane);   // lanes with bit 4 keep values 8..15
    warp_rs16_step<4>(v, lane);
    warp_rs16_step<2>(v, lane);
    warp_rs16_step<1>(v, lane);   // value index = lane >> 1
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool CL>
__device__ __forceinline__ void cta_reduce_post(float (&v)[32], GnShared & sh, unsigned long long * ring, unsigned step, int wbase, int ncols, ClArea * cl, const int rank,
                                                const int csize)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if(CL)
    {
        // (the cluster kernel only reduces the 11 SO3 columns here)
        const float s = warp_reduce_scatter16(v);
        if((lane & 1) == 0) sh.red[wid * 32 + (lane >> 1)] = s;
    }
    else
    {
        const float s = warp_reduce_scatter32(v);
        sh.red[wid * 32 + lane] = s;
    }
    __syncthreads();
    if(wid == 0)
    {
        float total = 0.f;
#pragma unroll
        for(int w = 0; w < kGnWarps; w++) total += sh.red[w * 32 + lane];
        if(lane < ncols)
        {
            if(CL)
                cl_post_col(cl, (int)(step & 1u), rank, csize, lane, total);
            else
                post_float(ring, step, wbase, lane, total);
        }
    }
}

// Block-level half of an ICP + RGB reduction: every warp holds, per lane, its partial sum of ICP column `lane` (s_icp) and of RGB
// column `lane` (s_rgb), from two warp_reduce_scatter32 calls; one barrier, then warp 0 adds up and posts the ICP columns while
// warp 1 does the same for the RGB columns.  has_icp / has_rgb: which halves exist (CTA-uniform).
template <bool CL>
__device__ __forceinline__ void cta_post_pair(const float s_icp, const float s_rgb, const bool has_icp, const bool has_rgb, GnShared & sh, unsigned long long * ring,
                                              unsigned step, ClArea * cl, const int rank, const int csize)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    sh.red[wid * 64 + lane] = s_icp;
    sh.red[wid * 64 + 32 + lane] = s_rgb;
    __syncthreads();
    if(wid < 2 && (wid == 0 ? has_icp : has_rgb))
    {
        float total = 0.f;
#pragma unroll
        for(int w = 0; w < kGnWarps; w++) total += sh.red[w * 64 + wid * 32 + lane];
        if(lane < 29)
        {
            if(CL)
                cl_post_col(cl, (int)(step & 1u), rank, csize, wid * 32 + lane, total);
            else
                post_float(ring, step, wid == 0 ? kWIcp : kWRgb, lane, total);
        }
    }
}

// Threads 0..127: read the ncols columns at word wbase (two words each) into dst[0..ncols-1] (floats).  t = index of this
// thread within the reader set of the block of words, t < 2 * ncols active.  Whole warps must call (shuffle inside).
// extra_word >= 0: this thread (not one of the column readers) reads that single integer word in the same polling loop (so that
// its trip through L2 overlaps the others') and gets the sum back in *extra_out.
__device__ __forceinline__ void read_columns(unsigned long long * ring, unsigned step, int wbase, int ncols, int t, unsigned want, float * dst, GnWork & wk,
                                             const int extra_word = -1, long long * extra_out = nullptr)
{
    const bool active = t >= 0 && t < 2 * ncols;
    const bool extra = !active && extra_word >= 0;
    double d = 0.0;
    if(active || extra)
    {
        const long long v = poll_word(ring, step, extra ? extra_word : wbase + t, want, wk);
        if(extra)
            *extra_out = v;
        else
            d = (t & 1) ? (double)v * (1.0 / kFracScale) : (double)v;
    }
    const double o = __shfl_xor_sync(0xffffffffu, d, 1);
    if(active && !(t & 1)) dst[t >> 1] = (float)(d + o);
}

// Read field by field from the kernel parameter (one sequence: constant-bank loads, no local copy of the block) or from the
// per-sequence array in global memory.
#define GN_LEVEL_FIELDS(S) \
    p.vcurr = (S).vcurr[lvl]; p.ncurr = (S).ncurr[lvl]; p.vprev = (S).vprev[lvl]; p.nprev = (S).nprev[lvl]; \
    p.lastDepth = (S).lastDepth[lvl]; p.nextDepth = (S).nextDepth[lvl]; p.lastImage = (S).lastImage[lvl]; \
    p.nextImage = (S).nextImage[lvl]; p.lastNextImage = (S).lastNextImage[lvl]; p.dIdx = (S).dIdx[lvl]; \
    p.dIdy = (S).dIdy[lvl]; p.corres = (S).corres[lvl];
__device__ __forceinline__ LevelPtrs level_ptrs(const bool one, const GnSeqIn & seq0, const GnSeqIn * seqs, int seq, int lvl)
{
    LevelPtrs p;
    if(one)
    {
        GN_LEVEL_FIELDS(seq0)
    }
    else
    {
        GN_LEVEL_FIELDS(seqs[seq])
    }
    return p;
}

// Pose-independent half of the photometric association (reduce.cu:780-807) for NS pixels of a thread at once, with every
// load of every pixel issued before the first use: the clipped 4x4 all-nonzero window of nextImage, the gradient pair (from
// dIdx/dIdy, or derived from the same window with the arithmetic of utils.cu:582-606 when no derivative images were made),
// nextDepth.  cand[s] = the pixel passes; d1 / gxy (gx | gy << 16) / img are its operands.
template <int NS>
__device__ __forceinline__ void rgb_candidate_state(const ResidualArgs & a, const bool derive, const int (&k)[NS], const bool (&live)[NS], const int (&px)[NS],
                                                    const int (&py)[NS], bool (&cand)[NS], float (&d1o)[NS], unsigned (&gxy)[NS], unsigned (&img)[NS])
{
    unsigned char w[NS][16];
    float d1[NS];
    short gxl[NS], gyl[NS];
#pragma unroll
    for(int s = 0; s < NS; s++)
    {
        const int kk = live[s] ? k[s] : 0;
        const int x = live[s] ? px[s] : 0, y = live[s] ? py[s] : 0;
#pragma unroll
        for(int r = 0; r < 4; r++)
#pragma unroll
            for(int q = 0; q < 4; q++)
            {
                const int u = min(max(y - 2 + r, 0), a.rows - 1), v = min(max(x - 2 + q, 0), a.cols - 1);
                w[s][r * 4 + q] = live[s] ? __ldg(a.nextImage + u * a.cols + v) : (unsigned char)0;
            }
        d1[s] = live[s] ? __ldg(a.nextDepth + kk) : 0.f;
        gxl[s] = (derive || !live[s]) ? (short)0 : __ldg(a.dIdx + kk);
        gyl[s] = (derive || !live[s]) ? (short)0 : __ldg(a.dIdy + kk);
    }
#pragma unroll
    for(int s = 0; s < NS; s++)
    {
        const int x = px[s], y = py[s];
        bool ok = live[s] && (x < a.cols - 5 && y < a.rows - 1);
        // window taps the reference's clipped loops never visit (rows / columns below 0) do not vote
#pragma unroll
        for(int r = 0; r < 4; r++)
#pragma unroll
            for(int q = 0; q < 4; q++)
            {
                const bool visited = (y - 2 + r >= 0) && (x - 2 + q >= 0);
                ok = ok && (!visited || w[s][r * 4 + q] > 0);
            }
        short gx = gxl[s], gy = gyl[s];
        if(derive)
        {
            if(x >= 1 && y >= 1 && x < a.cols - 1 && y < a.rows - 1)
            {
                // interior: the nine taps are window entries (r + 1, q + 1), accumulated in the reference's order
                const float fgx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
                const float fgy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
                float dxVal = 0, dyVal = 0;
#pragma unroll
                for(int t = 0; t < 9; t++)
                {
                    const float v = (float)w[s][(t / 3 + 1) * 4 + (t % 3 + 1)];
                    dxVal = __fmaf_rn(v, fgx[8 - t], dxVal);
                    dyVal = __fmaf_rn(v, fgy[8 - t], dyVal);
                }
                gx = (short)dxVal;
                gy = (short)dyVal;
            }
            else if(ok)
                derivative_pixel(a.nextImage, a.rows, a.cols, x, y, gx, gy);   // image border: the clipped loop itself
        }
        const int valx = gx, valy = gy;
        const float mTwo = (valx * valx) + (valy * valy);
        ok = ok && (mTwo >= a.minScale) && !isnan(d1[s]);
        cand[s] = ok;
        d1o[s] = d1[s];
        gxy[s] = ((unsigned)(unsigned short)gx) | (((unsigned)(unsigned short)gy) << 16);
        img[s] = w[s][2 * 4 + 2];   // the pixel itself
    }
}

// The same test for one 32-pixel segment that lies in one image row and starts at a multiple of 32 columns (every level of a
// 640x480 or 1280x720 pyramid): ten lanes fetch the row strip [x0 - 4, x0 + 36) of each of the four window rows as aligned 32-bit
// words, every lane cuts its own four taps (columns x - 2 .. x + 1) out of two neighbouring words with two shuffles and a funnel
// shift, and "all sixteen taps non-zero" becomes four zero-byte tests.  Four loads per segment row instead of sixteen byte loads
// per pixel.  Taps outside the image are filled with 0xff: the reference's clipped loops never visit them, so they do not vote.
__device__ __forceinline__ void rgb_candidate_segment(const ResidualArgs & a, const bool derive, const bool live /* warp-uniform */, const int x0, const int y,
                                                      const int k, bool & cand, float & d1o, unsigned & gxy, unsigned & img)
{
    const int lane = threadIdx.x & 31;
    const int x = x0 + lane;
    unsigned w[4];
    const int wpr = a.cols >> 2;                 // words per image row
    const int wi = (x0 >> 2) - 1 + lane;         // lanes 0..9: word of the strip
    const int b = lane + 2;                      // byte offset of this lane's first tap inside the strip
#pragma unroll
    for(int r = 0; r < 4; r++)
    {
        const int yy = y - 2 + r;
        unsigned strip = 0xffffffffu;
        if(live && lane < 10 && yy >= 0 && yy < a.rows && wi >= 0 && wi < wpr) strip = __ldg(reinterpret_cast<const unsigned *>(a.nextImage + (size_t)yy * a.cols) + wi);
        const unsigned lo = __shfl_sync(0xffffffffu, strip, b >> 2);
        const unsigned hi = __shfl_sync(0xffffffffu, strip, (b >> 2) + 1);
        w[r] = __funnelshift_r(lo, hi, 8 * (b & 3));
    }
    const float d1 = live ? __ldg(a.nextDepth + k) : 0.f;
    short gx = (derive || !live) ? (short)0 : __ldg(a.dIdx + k);
    short gy = (derive || !live) ? (short)0 : __ldg(a.dIdy + k);
    bool ok = live && (x < a.cols - 5 && y < a.rows - 1);
#pragma unroll
    for(int r = 0; r < 4; r++) ok = ok && (((w[r] - 0x01010101u) & ~w[r] & 0x80808080u) == 0u);
    ok = ok && !isnan(d1);
    // the gradient is only needed where the other tests passed; pixels without depth come in regions (beyond the cut-off, holes), so
    // whole segments skip it
    if(derive && __any_sync(0xffffffffu, ok))
    {
        if(x >= 1 && y >= 1 && x < a.cols - 1 && y < a.rows - 1)
        {
            // interior: the nine taps are window entries (r + 1, q + 1), accumulated in the reference's order
            const float fgx[9] = {0.52201f, 0.00000f, -0.52201f, 0.79451f, -0.00000f, -0.79451f, 0.52201f, 0.00000f, -0.52201f};
            const float fgy[9] = {0.52201f, 0.79451f, 0.52201f, 0.00000f, 0.00000f, 0.00000f, -0.52201f, -0.79451f, -0.52201f};
            float dxVal = 0, dyVal = 0;
#pragma unroll
            for(int t = 0; t < 9; t++)
            {
                const float v = (float)((w[t / 3 + 1] >> (8 * (t % 3 + 1))) & 0xffu);
                dxVal = __fmaf_rn(v, fgx[8 - t], dxVal);
                dyVal = __fmaf_rn(v, fgy[8 - t], dyVal);
            }
            gx = (short)dxVal;
            gy = (short)dyVal;
        }
        else if(ok)
            derivative_pixel(a.nextImage, a.rows, a.cols, x, y, gx, gy);   // image border: the clipped loop itself
    }
    const int valx = gx, valy = gy;
    const float mTwo = (valx * valx) + (valy * valy);
    ok = ok && (mTwo >= a.minScale);
    cand = ok;
    d1o = d1;
    gxy = ((unsigned)(unsigned short)gx) | (((unsigned)(unsigned short)gy) << 16);
    img = (w[2] >> 16) & 0xffu;   // the pixel itself
}

// Stage one resident level of this CTA: lists of ICP and RGB entries in shared memory (see the header comment).
// Pass 1 loads the operands of every pixel of the CTA's segments ONCE (one round trip to L2 per segment slot), stores them at
// their uncompacted position and counts the survivors per (slot, warp); a block-wide prefix sum turns the counts into list
// positions; pass 2 compacts the lists in place, shared memory only (entries only move towards the front, slot by slot).
// Rolled on purpose: the kernel is bound by instruction fetch, and this code runs once per level.
__device__ __forceinline__ void stage_level(const GnLaunch & L, const bool icp, const bool rgb, const int lvl, const LevelPtrs & P, const int rank, GnWork & wk, char * dyn)
{
    const LevelPlan pl = L.plan[lvl];
    const LevelGeom g = L.geom[lvl];
    const int plane = g.rows * g.cols;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float * icp_list = reinterpret_cast<float *>(dyn + pl.off_icp);
    float * rgb_d1 = reinterpret_cast<float *>(dyn + pl.off_rgb);
    unsigned * rgb_gxy = reinterpret_cast<unsigned *>(rgb_d1 + pl.cap);
    unsigned * rgb_xyi = rgb_gxy + pl.cap;
    const int nslots = (pl.segs_per_cta + kGnWarps - 1) / kGnWarps;   // <= kMaxStageSlots (gn_make_plan); cap = segs_per_cta * 32
    const unsigned lt = (1u << lane) - 1u;
    const bool aligned = (g.cols & 31) == 0;   // every 32-pixel segment lies in one row and starts at a multiple of 32 columns
    // The ICP operands are plain copies (global -> their uncompacted place in the list): cp.async puts every one of them in flight up
    // front, with no register and no wait in the segment loop below, whose round trips to L2 they used to share; what the loop needs
    // from them (does the pixel survive?) is read back from shared memory afterwards.
    if(icp)
    {
#pragma unroll 1
        for(int m = 0; m < nslots; m++)
        {
            const int j = m * kGnWarps + wid;
            if(j >= pl.segs_per_cta) break;
            const int seg = j * pl.P + rank;
            const int u = m * kGnThreads + (int)threadIdx.x;
            const int k1 = seg * 32 + lane;
            if(seg < pl.nseg && k1 < plane)
            {
#pragma unroll
                for(int q = 0; q < 3; q++)
                {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(icp_list + q * pl.cap + u)), "l"(P.vcurr + q * plane + k1) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(icp_list + (3 + q) * pl.cap + u)), "l"(P.ncurr + q * plane + k1) : "memory");
                }
            }
            else
            {
#pragma unroll
                for(int q = 0; q < 6; q++) icp_list[q * pl.cap + u] = SLAM_QNAN;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {
        ResidualArgs ra;
        ra.minScale = L.min_scale[lvl];
        ra.dIdx = P.dIdx; ra.dIdy = P.dIdy;
        ra.nextDepth = P.nextDepth;
        ra.nextImage = P.nextImage;
        ra.cols = g.cols; ra.rows = g.rows;
#pragma unroll 1
        for(int m = 0; m < nslots; m++)
        {
            const int j = m * kGnWarps + wid;        // this warp's j-th segment of the CTA
            const int seg = j * pl.P + rank;
            const int u = m * kGnThreads + (int)threadIdx.x;
            const int k1[1] = {seg * 32 + lane};
            const bool l1[1] = {j < pl.segs_per_cta && seg < pl.nseg && k1[0] < plane};
            const int kk = l1[0] ? k1[0] : 0;
            const int y1[1] = {kk / g.cols};
            const int x1[1] = {kk - y1[0] * g.cols};
            bool c1[1] = {false};
            if(rgb)
            {
                float d1[1];
                unsigned g1[1], i1[1];
                if(aligned)
                    rgb_candidate_segment(ra, L.derive_gradients, j < pl.segs_per_cta && seg < pl.nseg, x1[0] - lane, y1[0], kk, c1[0], d1[0], g1[0], i1[0]);
                else
                    rgb_candidate_state<1>(ra, L.derive_gradients, k1, l1, x1, y1, c1, d1, g1, i1);
                // the reference writes a DataTerm for every pixel (reduce.cu:838): pixels that never become entries get their zero once
                if(L.full_corres && l1[0] && !c1[0]) reinterpret_cast<int4 *>(P.corres)[kk] = make_int4(0, 0, 0, 0);
                if(j < pl.segs_per_cta)
                {
                    rgb_d1[u] = d1[0];
                    rgb_gxy[u] = g1[0];
                    // a candidate's own intensity is non-zero (it is part of the window test): intensity 0 marks "not a candidate"
                    rgb_xyi[u] = c1[0] ? ((unsigned)x1[0] | ((unsigned)y1[0] << 11) | (i1[0] << 22)) : 0u;
                }
            }
            const unsigned br = __ballot_sync(0xffffffffu, c1[0]);
            if(lane == 0) wk.scan_cnt[1][m * kGnWarps + wid] = __popc(br);
        }
    }
    if(icp)
    {
        // the copies have landed: survivors per (slot, warp) from the staged vertex / normal x planes
        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll 1
        for(int m = 0; m < nslots; m++)
        {
            const int u = m * kGnThreads + (int)threadIdx.x;
            const bool exists = m * kGnWarps + wid < pl.segs_per_cta;
            const bool fi = exists && !isnan(icp_list[u]) && !isnan(icp_list[3 * pl.cap + u]);
            const unsigned bi = __ballot_sync(0xffffffffu, fi);
            if(lane == 0) wk.scan_cnt[0][m * kGnWarps + wid] = __popc(bi);
        }
    }
    else if(lane == 0)
    {
        for(int m = 0; m < nslots; m++) wk.scan_cnt[0][m * kGnWarps + wid] = 0;
    }
    __syncthreads();
    if(wid < 2)
    {
        // exclusive prefix of the counts in (slot, warp) order, in place; warp 0: ICP, warp 1: RGB
        constexpr int kPer = kMaxStageSlots * kGnWarps / 32;
        const int n = nslots * kGnWarps;
        int c[kPer], sum = 0;
#pragma unroll
        for(int q = 0; q < kPer; q++)
        {
            const int idx = lane * kPer + q;
            c[q] = idx < n ? wk.scan_cnt[wid][idx] : 0;
            sum += c[q];
        }
        int incl = sum;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if(lane >= o) incl += up;
        }
        int run = incl - sum;
#pragma unroll
        for(int q = 0; q < kPer; q++)
        {
            const int idx = lane * kPer + q;
            if(idx < n) wk.scan_cnt[wid][idx] = run;
            run += c[q];
        }
        if(lane == 31)
        {
            if(wid == 0) wk.n_icp[lvl] = run;
            else wk.n_rgb[lvl] = run;
        }
    }
    __syncthreads();
    // in-place compaction: slot m's entries are read, then (after a barrier) written at their final positions, which lie at or
    // before their old ones and behind everything earlier slots wrote; later slots' entries are not touched
#pragma unroll 1
    for(int m = 0; m < nslots; m++)
    {
        const int u = m * kGnThreads + (int)threadIdx.x;
        const bool exists = m * kGnWarps + wid < pl.segs_per_cta;
        float v[6];
        float d1 = 0.f;
        unsigned gq = 0u, xyi = 0u;
        bool fi = false;
        if(icp && exists)
        {
#pragma unroll
            for(int q = 0; q < 6; q++) v[q] = icp_list[q * pl.cap + u];
            fi = !isnan(v[0]) && !isnan(v[3]);
        }
        if(rgb && exists)
        {
            d1 = rgb_d1[u];
            gq = rgb_gxy[u];
            xyi = rgb_xyi[u];
        }
        const bool fr = rgb && (xyi >> 22) != 0u;
        const unsigned bi = __ballot_sync(0xffffffffu, fi);
        const unsigned br = __ballot_sync(0xffffffffu, fr);
        __syncthreads();
        if(fi)
        {
            const int pos = wk.scan_cnt[0][m * kGnWarps + wid] + __popc(bi & lt);
#pragma unroll
            for(int q = 0; q < 6; q++) icp_list[q * pl.cap + pos] = v[q];
        }
        if(fr)
        {
            const int pos = wk.scan_cnt[1][m * kGnWarps + wid] + __popc(br & lt);
            rgb_d1[pos] = d1;
            rgb_gxy[pos] = gq;
            rgb_xyi[pos] = xyi;
        }
    }
    __syncthreads();
}

// End of a sequence: the jump guard and lastRGBError (seq_end), then the result block of the sequence in DEVICE memory.  Only the
// fields this call computes are written -- the reference's lastICPError / lastRGBError / lastSO3Error are members that keep their
// value across calls which do not run that term -- so the block always holds the current statistics and the host fetches it when
// somebody asks (slam_odom_get_stats / _get_covariance), not once per frame.  Lane 0.
static __device__ __noinline__ void seq_end_merge(GnShared & sh, const bool icp, const bool rgb, const bool rgb_only, const bool so3, GnResult * out)
{
    seq_end(sh, rgb, rgb_only, nullptr);
    if(!out) return;
    for(int k = 0; k < 9; k++) out->Rcurr[k] = sh.Rcurr[k];
    for(int k = 0; k < 3; k++) out->tcurr[k] = sh.tcurr[k];
    if(icp)
    {
        out->lastICPError = sh.res.lastICPError;
        out->lastICPCount = sh.res.lastICPCount;
    }
    if(rgb)
    {
        out->lastRGBError = sh.res.lastRGBError;
        out->lastRGBCount = sh.res.lastRGBCount;
    }
    if(so3)
    {
        out->lastSO3Error = sh.res.lastSO3Error;
        out->lastSO3Count = sh.res.lastSO3Count;
    }
    for(int k = 0; k < 36; k++) out->lastA[k] = sh.res.lastA[k];
    for(int k = 0; k < 6; k++) out->lastb[k] = sh.res.lastb[k];
    out->so3_iterations = sh.res.so3_iterations;
    out->gn_iterations = sh.res.gn_iterations;
}

// The intrinsics of a level and their inverse, prepared by the host (same routine, same bits as level_begin of gn_scalar.cuh).  Lanes 0..8 of warp 0.
__device__ __forceinline__ void level_begin_from(GnShared & sh, const GnLaunch & L, const int lvl)
{
    const int lane = threadIdx.x & 31;
    if(lane < 9)
    {
        sh.K[lane] = L.K[lvl][lane];
        sh.Kinv[lane] = L.Kinv[lvl][lane];
    }
}

// ------------------------------------------------------------------ the kernel
// ICP / RGB / RGB_ONLY: the mode of the call (RGBDOdometryef.cpp:275-276), compile-time so that each variant carries only its
// own phases.  GEN = false is the product's common case and the one tuned for instruction footprint: every level resident,
// no step trace; GEN = true adds the streamed levels, the step trace / time stamps and the full DataTerm image.
// PH: per-phase cycle accounting of the leading CTA (slam_odom_get_phase_cycles).
// ROLE: 0 = the whole frame in one cooperative launch.  Split launch (one sequence, every level resident): 1 = the SO3 pre-alignment
// and the coarse levels on ONE thread-block cluster (grid = cluster = G CTAs, all-reduce through distributed shared memory), which
// hands the running state to 2 = the fine levels on the other SMs; that kernel is released (programmatic dependent launch) as soon
// as the cluster is resident, stages its levels next to the cluster's iterations and then waits for the hand-off.
template <bool ICP, bool RGB, bool RGB_ONLY, bool GEN, bool PH, int ROLE>
__global__ void __launch_bounds__(kGnThreads, 1)
k_gn_persistent(const GnLaunch L, GnCtl * ctl, const GnSeqIn * seqs, const GnSeqIn seq0, unsigned long long * rings, GnResult * results, slam_step_record * trace,
                int * trace_count, const int G, const int groups, GnResult * host_results, unsigned * host_flags, const unsigned host_seqno,
                const unsigned long long gate_target, const unsigned long long handoff_seq)
{
    __shared__ GnShared sh;
    __shared__ GnWork wk;
    extern __shared__ __align__(16) char dyn[];
    constexpr bool CL = ROLE == 1;
    // the fine-level kernel may start as soon as every CTA of the cluster is resident (it does not wait for this grid's memory:
    // what it needs from here arrives through GnCtl::handoff)
    // The cluster itself is launched as a programmatic dependent of the kernel in front of it (the frame preparation releases its
    // dependents when it starts): its CTAs take their SMs as that grid drains and wait here for its completion and memory.
    if(CL)
    {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }

    const int group = blockIdx.x / G;
    const int rank = blockIdx.x - group * G;   // CL: grid = one cluster, so this is %cluster_ctarank
    if(group >= groups) return;
    ClArea * cl = reinterpret_cast<ClArea *>(dyn + L.off_cl);

    unsigned long long * ring = rings + (size_t)group * (kRingBytes / 8);
    unsigned step = 0;   // reduction steps done so far: position in the ring of word sets

    const long long t_start = clock64();
#define GN_STAMP(rec, idx) do { if(GEN && rec) (rec)->t_cycles[idx] = (unsigned)(clock64() - t_start); } while(0)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool leader = (rank == 0 && threadIdx.x == 0);
    const bool warp0 = threadIdx.x < 32;
    if(threadIdx.x == 0)
    {
        wk.cnt[0] = wk.cnt[1] = 0;
        wk.timeouts = 0;
    }
    long long ph_t = t_start;
    if(PH && threadIdx.x < 24) wk.ph[threadIdx.x] = 0u;
    unsigned long long gt_start = 0ull;
    if(PH && ROLE != 0 && leader)
    {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_start));
        if(CL) *reinterpret_cast<volatile unsigned long long *>(&ctl->dbg_t[0]) = gt_start;
    }
#define GN_PHASE(idx) do { if(PH && leader) { const long long now_ = clock64(); wk.ph[idx] += (unsigned)(now_ - ph_t); ph_t = now_; } } while(0)

    // the reduction words as the previous launch left them (nobody posts before every CTA of the group has passed its first
    // wait, and that needs this CTA's own post)
    if(!CL)
        for(int w = threadIdx.x; w < kRingSlots * kRingWords; w += kGnThreads) wk.base[w] = ld_u64_relaxed(ring + (size_t)w * kWordStride);
    __syncthreads();
    // ... which has to hold for a CTA that starts late as well: every CTA checks in once its bases are loaded, and nobody posts
    // before all have (the check is made after the staging, when it has long been true)
    if(!CL && threadIdx.x == 0) red_u64(&ctl->arrived[group], 1ull);
    bool gate_open = CL;   // the cluster does not use the reduction words
    // nobody stores into a peer's shared memory before that peer is running: every CTA arrives here, and waits for the others only
    // right before the first all-reduce (by then they have long arrived)
    bool cl_start_pending = CL;
    if(CL) cl_arrive();
    // (iii) the SO3 images of the first sequence go on their way before anything else (cp.async, collected where they are needed)
    bool so3_images_hoisted = false;
    if(CL && L.so3 && L.so3_resident && rank < L.so3_P)
    {
        const LevelPtrs P2 = level_ptrs(L.batch == 1, seq0, seqs, (ROLE != 0 && L.batch > 1) ? L.batch - 1 : group, 2);
        const int N = L.geom[2].rows * L.geom[2].cols;
        if((N & 15) == 0)
        {
            unsigned char * s_last = reinterpret_cast<unsigned char *>(dyn + L.off_so3);
            unsigned char * s_next = s_last + ((N + 15) & ~15);
#pragma unroll 1
            for(int q = threadIdx.x; q < N / 16; q += kGnThreads)
            {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_last + 16 * q)), "l"(P2.lastNextImage + 16 * q) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_next + 16 * q)), "l"(P2.nextImage + 16 * q) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            so3_images_hoisted = true;
        }
    }
#define GN_GATE() do { if(!gate_open) { if(threadIdx.x == 0) { int spin_ = 0; while(ld_u64_relaxed(&ctl->arrived[group]) < gate_target && ++spin_ < kSpinCap) {} if(spin_ >= kSpinCap) wk.timeouts = 1; } __syncthreads(); gate_open = true; } } while(0)

    // A split pair that works through several sequences takes them LAST FIRST: the batched preparation in front of it wrote the
    // sequences in ascending order, so the last ones are what is still in L2 (8 x 45 MB do not fit)
    const int seq_first = (ROLE != 0 && L.batch > 1) ? L.batch - 1 : group;
    const int seq_step = (ROLE != 0 && L.batch > 1) ? -1 : groups;
    for(int seq = seq_first; seq >= 0 && seq < L.batch; seq += seq_step)
    {
        slam_step_record * tr = (GEN && L.trace && leader) ? trace + (size_t)seq * kGnMaxTrace : nullptr;
        int ntr = 0;

        if(threadIdx.x == 0)
        {
            float Rp[9], tp[3];
            if(L.batch == 1)
            {
#pragma unroll
                for(int k = 0; k < 9; k++) Rp[k] = seq0.Rprev[k];
#pragma unroll
                for(int k = 0; k < 3; k++) tp[k] = seq0.tprev[k];
            }
            else
            {
#pragma unroll
                for(int k = 0; k < 9; k++) Rp[k] = seqs[seq].Rprev[k];
#pragma unroll
                for(int k = 0; k < 3; k++) tp[k] = seqs[seq].tprev[k];
            }
            seq_begin_pose<false>(sh, Rp, tp);
        }
        else if(wid == 1)
        {
            for(int k = lane; k < (int)(sizeof(GnResult) / sizeof(int)); k += 32) reinterpret_cast<int *>(&sh.res)[k] = 0;
        }
        __syncthreads();

        // ------------------------------------------------ staging: every resident level, and the SO3 images
#pragma unroll 1
        for(int lvl = L.levels - 1; lvl >= 0; lvl--)
        {
            if(L.iterations[lvl] > 0 && L.plan[lvl].resident && rank < L.plan[lvl].P)
                stage_level(L, ICP, RGB, lvl, level_ptrs(L.batch == 1, seq0, seqs, seq, lvl), rank, wk, dyn);
            GN_PHASE(16 + (lvl < 3 ? lvl : 3));
        }
        if(L.so3 && L.so3_resident && rank < L.so3_P)
        {
            const LevelPtrs P2 = level_ptrs(L.batch == 1, seq0, seqs, seq, 2);
            const int N = L.geom[2].rows * L.geom[2].cols;
            unsigned char * s_last = reinterpret_cast<unsigned char *>(dyn + L.off_so3);
            unsigned char * s_next = s_last + ((N + 15) & ~15);
            if(so3_images_hoisted && seq == seq_first)
            {
                asm volatile("cp.async.wait_all;" ::: "memory");   // issued at the start of the kernel
            }
            else if((N & 15) == 0)
            {
#pragma unroll 1
                for(int q = threadIdx.x; q < N / 16; q += kGnThreads)
                {
                    reinterpret_cast<uint4 *>(s_last)[q] = __ldg(reinterpret_cast<const uint4 *>(P2.lastNextImage) + q);
                    reinterpret_cast<uint4 *>(s_next)[q] = __ldg(reinterpret_cast<const uint4 *>(P2.nextImage) + q);
                }
            }
            else
            {
#pragma unroll 1
                for(int q = threadIdx.x; q < N; q += kGnThreads)
                {
                    s_last[q] = __ldg(P2.lastNextImage + q);
                    s_next[q] = __ldg(P2.nextImage + q);
                }
            }
            __syncthreads();
        }
        GN_PHASE(20);
        GN_GATE();
        if(CL && cl_start_pending)
        {
            cl_wait();
            cl_start_pending = false;
        }
        GN_PHASE(0);

        // ------------------------------------------------ SO3 pre-alignment, level 2
        if(L.so3)
        {
            const LevelGeom g = L.geom[2];
            const int N = g.rows * g.cols;
            const int Pn = L.so3_P;
            if(warp0)
            {
                level_begin_from(sh, L, 2);
                __syncwarp();
                if(threadIdx.x == 0)
                {
                    so3_prepare(sh);
                    wk.P = level_ptrs(L.batch == 1, seq0, seqs, seq, 2);
                }
            }
            __syncthreads();
#pragma unroll 1
            for(int it = 0; it < 10; it++)
            {
                if(rank < Pn)
                {
                    So3Args a;
                    if(L.so3_resident)
                    {
                        a.lastImage = reinterpret_cast<const unsigned char *>(dyn + L.off_so3);
                        a.nextImage = a.lastImage + ((N + 15) & ~15);
                    }
                    else
                    {
                        a.lastImage = wk.P.lastNextImage;
                        a.nextImage = wk.P.nextImage;
                    }
                    a.imageBasis = mat3_from(sh.so3H);
                    a.kinv = mat3_from(sh.so3Kinv);
                    a.krlr = mat3_from(sh.so3KR);
                    a.cols = g.cols;
                    a.rows = g.rows;

                    float acc[32];
#pragma unroll
                    for(int k = 0; k < 32; k++) acc[k] = 0.f;
#pragma unroll 1
                    for(int k = rank * kGnThreads + threadIdx.x; k < N; k += Pn * kGnThreads)
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
                    GN_PHASE(11);
                    cta_reduce_post<CL>(acc, sh, ring, step, kWSo3, 11, cl, rank, G);
                }
                if(CL)
                {
                    cl_arrive();
                    cl_wait();
                }
                GN_PHASE(12);
                if(warp0)
                {
                    if(CL)
                    {
                        if(lane < 11) sh.total[lane] = cl_fold_col(cl, (int)(step & 1u), Pn, lane);
                    }
                    else
                    {
                        spin_cycles(L.poll_delay);
                        read_columns(ring, step, kWSo3, 11, threadIdx.x, (unsigned)Pn, sh.total, wk);
                    }
                    __syncwarp();
                    GN_PHASE(13);
                    slam_step_record * rec = (GEN && threadIdx.x == 0 && tr && ntr < kGnMaxTrace) ? &tr[ntr] : nullptr;
                    if(GEN && rec) memset(rec, 0, sizeof(*rec));
                    warp_so3_update_fast(sh, it, rec);   // also leaves the next iteration's H, K^-1, K R in shared memory
                    if(threadIdx.x == 0) wk.so3_stop[it & 1] = sh.stop;
                    if(GEN && rec) ntr++;
                    GN_PHASE(14);
                }
                step++;
                __syncthreads();
                if(wk.so3_stop[it & 1]) break;
            }
        }
        GN_PHASE(1);

        if(ROLE == 2)
        {
            // the running state after the cluster's levels (GnCtl::handoff, published with the number of this launch pair)
            if(warp0)
            {
                if(lane == 0)
                {
                    int spin = 0;
                    unsigned long long v;
                    do
                    {
                        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&ctl->handoff_seq[seq]) : "memory");
                    } while(v != handoff_seq && ++spin < kSpinCap);
                    if(spin >= kSpinCap) wk.timeouts = 1;
                }
                __syncwarp();
                const GnHandoff * ho = &ctl->handoff[seq];
                const uint2 * src = reinterpret_cast<const uint2 *>(&ho->res);
                uint2 * dst = reinterpret_cast<uint2 *>(&sh.res);
                for(int k = lane; k < (int)(sizeof(GnResult) / sizeof(uint2)); k += 32) dst[k] = __ldcg(src + k);
                if(lane < 16) sh.resultRt[lane] = __ldcg(&ho->resultRt[lane]);
                if(lane < 9) sh.Rcurr[lane] = __ldcg(&ho->Rcurr[lane]);
                if(lane < 3) sh.tcurr[lane] = __ldcg(&ho->tcurr[lane]);
                if(lane == 0)
                {
                    sh.rgb_sigma_last = __ldcg(&ho->rgb_sigma_last);
                    sh.rgb_count_last = __ldcg(&ho->rgb_count_last);
                    if(__ldcg(&ho->timeouts)) wk.timeouts = 1;
                }
            }
        }
        else if(threadIdx.x == 0)
        {
            for(int k = 0; k < 16; k++) sh.resultRt[k] = (k % 5 == 0) ? 1.0 : 0.0;
            if(L.so3)
                for(int x = 0; x < 3; x++)
                    for(int y = 0; y < 3; y++) sh.resultRt[x * 4 + y] = sh.resultR[x * 3 + y];
        }
        __syncthreads();
        GN_PHASE(21);
        if(PH && ROLE == 2 && leader)
        {
            // nanoseconds since the cluster kernel started: when this kernel started, when the hand-off arrived
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            const unsigned long long a0 = *reinterpret_cast<volatile unsigned long long *>(&ctl->dbg_t[0]);
            wk.ph[22] = (unsigned)(gt_start - a0);
            wk.ph[23] = (unsigned)(now - a0);
        }

        // ------------------------------------------------ coarse-to-fine ICP + RGB
#pragma unroll 1
        for(int lvl = L.levels - 1; lvl >= 0; lvl--)
        {
            if(L.iterations[lvl] <= 0) continue;
            const bool resident = !GEN || L.plan[lvl].resident != 0;
            const int Pn = L.plan[lvl].P;
            const bool part = rank < Pn;
            if(warp0)
            {
                level_begin_from(sh, L, lvl);
                if(threadIdx.x == 0)
                {
                    sh.res.lastRGBError = FLT_MAX;
                    wk.P = level_ptrs(L.batch == 1, seq0, seqs, seq, lvl);
                    wk.lv_n_icp = resident ? wk.n_icp[lvl] : 0;
                    wk.lv_n_rgb = resident ? wk.n_rgb[lvl] : 0;
                }
                __syncwarp();
                warp_prepare_fast(sh, false);   // krk / kt of the first iteration of this level (Rcurr/tcurr carry over)
            }
            __syncthreads();

#pragma unroll 1
            for(int j = 0; j < L.iterations[lvl]; j++)
            {
                slam_step_record * rec = nullptr;
                if(GEN && threadIdx.x == 0 && tr && ntr < kGnMaxTrace)
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
                GN_PHASE(2);

                // ---------------- RGB association (reduce.cu:809-841): correspondences + their count
                if(RGB && part)
                {
                    const LevelGeom g = L.geom[lvl];
                    const int cap = L.plan[lvl].cap;
                    ResidualArgs a;
                    a.minScale = L.min_scale[lvl];
                    a.dIdx = wk.P.dIdx; a.dIdy = wk.P.dIdy;
                    a.lastDepth = wk.P.lastDepth; a.nextDepth = wk.P.nextDepth;
                    a.lastImage = wk.P.lastImage; a.nextImage = wk.P.nextImage;
                    a.maxDepthDelta = L.max_depth_delta;
                    a.kt = make_float3(sh.kt[0], sh.kt[1], sh.kt[2]);
                    a.krkinv = mat3_from(sh.krk);
                    a.cols = g.cols; a.rows = g.rows;
                    int cnt0 = 0, cnt1 = 0;
                    if(resident)
                    {
                        const int n_rgb = wk.lv_n_rgb;
                        const float * rgb_d1 = reinterpret_cast<const float *>(dyn + L.plan[lvl].off_rgb);
                        const unsigned * rgb_xyi = reinterpret_cast<const unsigned *>(rgb_d1 + 2 * cap);
                        int * st_zxy = reinterpret_cast<int *>(dyn + L.off_state);   // (zy << 16) | zx of the correspondence, -1: none
                        float * st_diff = reinterpret_cast<float *>(st_zxy + cap);
                        float * st_d0 = st_diff + cap;
#pragma unroll 1
                        for(int e0 = threadIdx.x; e0 < n_rgb; e0 += kRgbChunk * kGnThreads)
                        {
                            int o0[kRgbChunk], zxy[kRgbChunk];
                            float td1[kRgbChunk], d0[kRgbChunk];
                            unsigned xyi[kRgbChunk];
                            unsigned char lst[kRgbChunk];
                            bool inb[kRgbChunk];
                            // 1) warp every entry, 2) issue the gathers, 3) gates
#pragma unroll
                            for(int c = 0; c < kRgbChunk; c++)
                            {
                                const int e = e0 + c * kGnThreads;
                                const bool live = e < n_rgb;
                                const int ee = live ? e : e0;
                                xyi[c] = rgb_xyi[ee];
                                int u0, v0;
                                inb[c] = rgb_project(a, (int)(xyi[c] & 2047u), (int)((xyi[c] >> 11) & 2047u), rgb_d1[ee], u0, v0, td1[c]) && live;
                                o0[c] = inb[c] ? v0 * g.cols + u0 : 0;
                                zxy[c] = (v0 << 16) | u0;
                            }
#pragma unroll
                            for(int c = 0; c < kRgbChunk; c++)
                            {
                                d0[c] = inb[c] ? __ldg(a.lastDepth + o0[c]) : 0.f;
                                lst[c] = inb[c] ? __ldg(a.lastImage + o0[c]) : (unsigned char)0;
                            }
#pragma unroll
                            for(int c = 0; c < kRgbChunk; c++)
                            {
                                const int e = e0 + c * kGnThreads;
                                if(e < n_rgb)
                                {
                                    const bool valid = inb[c] && rgb_accept(a, td1[c], d0[c], lst[c]);
                                    const float diff = __fsub_rn(static_cast<float>(xyi[c] >> 22), static_cast<float>(lst[c]));
                                    if(valid)
                                    {
                                        cnt0 += 1;
                                        cnt1 += (int)(diff * diff);
                                    }
                                    st_zxy[e] = valid ? zxy[c] : -1;
                                    st_diff[e] = diff;
                                    st_d0[e] = d0[c];
                                    if(GEN && L.full_corres)   // the reference writes a DataTerm for every pixel (reduce.cu:838): only when a test taps it
                                    {
                                        const int x = (int)(xyi[c] & 2047u), y = (int)((xyi[c] >> 11) & 2047u);
                                        Corres cc;
                                        cc.zx = valid ? (short)(zxy[c] & 0xffff) : 0;
                                        cc.zy = valid ? (short)(zxy[c] >> 16) : 0;
                                        cc.ox = valid ? (short)x : 0;
                                        cc.oy = valid ? (short)y : 0;
                                        cc.diff = valid ? diff : 0.f;
                                        cc.valid = valid ? 1 : 0;
                                        reinterpret_cast<int4 *>(wk.P.corres)[y * g.cols + x] = *reinterpret_cast<const int4 *>(&cc);
                                    }
                                }
                            }
                        }
                    }
                    else if(GEN)
                    {
                        // streamed level: correspondences go through memory (corresImg, as in the reference)
                        const int plane = g.rows * g.cols;
                        Corres * cimg = wk.P.corres;
#pragma unroll 1
                        for(int k = rank * kGnThreads + threadIdx.x; k < plane; k += G * kGnThreads)
                        {
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
                    const int w0 = __reduce_add_sync(0xffffffffu, cnt0);
                    const int w1 = __reduce_add_sync(0xffffffffu, cnt1);
                    if(lane == 0 && (w0 | w1))
                    {
                        atomicAdd(&wk.cnt[0], w0);
                        atomicAdd(&wk.cnt[1], w1);
                    }
                    __syncthreads();
                    if(threadIdx.x == 0)
                    {
                        const int c0 = wk.cnt[0], c1 = wk.cnt[1];
                        wk.cnt[0] = wk.cnt[1] = 0;
                        if(CL)
                        {
                            for(int r = 0; r < G; r++) cl_st_v2(cl_map(&cl->mid[step & 1u][rank], (unsigned)r), c0, c1);
                        }
                        else
                        {
                            post_int(ring, step, kWMid, (long long)c0 + (c1 != 0 ? (1ll << 32) : 0ll));
                            post_int(ring, step, kWSigma, (long long)c1);
                        }
                    }
                }
                if(CL && RGB) cl_arrive();   // the counts travel while the ICP products are formed
                GN_PHASE(3);

                // Two map + block-reduction passes share ONE copy of the reduction code: pass 0 = ICP products (reduce.cu:282-416;
                // the count's trip through L2 overlaps this map), pass 1 = RGB Jacobian products (reduce.cu:494-624).
                bool stop_now = false;
                unsigned long long mid_peek = 0ull;
                float s_icp = 0.f;   // this warp's partial sum of ICP column `lane` (kept in a register until the RGB products are in)
#pragma unroll 1
                for(int pass = ICP ? 0 : 1; pass < (RGB ? 2 : 1); pass++)
                {
                    float acc[32];
#pragma unroll
                    for(int k = 0; k < 32; k++) acc[k] = 0.f;
                    if(ICP && pass == 0)
                    {
                        if(part)
                        {
                            const LevelGeom g = L.geom[lvl];
                            const int plane = g.rows * g.cols;
                            const int cap = L.plan[lvl].cap;
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
                            a.vcurr = wk.P.vcurr; a.ncurr = wk.P.ncurr; a.vprev = wk.P.vprev; a.nprev = wk.P.nprev;
                            const float * icp_list = reinterpret_cast<const float *>(dyn + L.plan[lvl].off_icp);
                            const int n_items = resident ? wk.lv_n_icp : plane;
                            const int first = resident ? (int)threadIdx.x : rank * kGnThreads + (int)threadIdx.x;
                            const int stride = resident ? kGnThreads : G * kGnThreads;
#pragma unroll 1
                            for(int e0 = first; e0 < n_items; e0 += kIcpChunk * stride)
                            {
                                float3 vg[kIcpChunk], nc[kIcpChunk], vp[kIcpChunk], np[kIcpChunk];
                                int o[kIcpChunk];
                                bool ok[kIcpChunk];
                                // 1) current vertex + normal of every entry of the chunk, 2) project all, 3) issue all gathers, 4) gates, rows, products
#pragma unroll
                                for(int c = 0; c < kIcpChunk; c++)
                                {
                                    const int e = e0 + c * stride;
                                    ok[c] = e < n_items;
                                    const int ee = ok[c] ? e : e0;
                                    float3 v;
                                    if(resident)
                                    {
                                        v = make_float3(icp_list[ee], icp_list[cap + ee], icp_list[2 * cap + ee]);
                                        nc[c] = make_float3(icp_list[3 * cap + ee], icp_list[4 * cap + ee], icp_list[5 * cap + ee]);
                                    }
                                    else
                                    {
                                        v = make_float3(__ldg(a.vcurr + ee), __ldg(a.vcurr + plane + ee), __ldg(a.vcurr + 2 * plane + ee));
                                        nc[c] = make_float3(__ldg(a.ncurr + ee), __ldg(a.ncurr + plane + ee), __ldg(a.ncurr + 2 * plane + ee));
                                    }
                                    const bool inb = icp_project(a, v, vg[c], o[c]);
                                    ok[c] = ok[c] && inb;
                                    if(!ok[c]) o[c] = 0;
                                }
#pragma unroll
                                for(int c = 0; c < kIcpChunk; c++)
                                {
                                    vp[c] = make_float3(__ldg(a.vprev + o[c]), __ldg(a.vprev + plane + o[c]), __ldg(a.vprev + 2 * plane + o[c]));
                                    np[c] = make_float3(__ldg(a.nprev + o[c]), __ldg(a.nprev + plane + o[c]), __ldg(a.nprev + 2 * plane + o[c]));
                                }
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
                        GN_STAMP(rec, 2);
                        GN_PHASE(4);
                    }
                    else
                    {
                        // ---------------- the global correspondence count: every CTA (rgbOnly decides the early exit on it)
                        if(CL)
                        {
                            cl_wait();
                            if(threadIdx.x == 0)
                            {
                                long long c0 = 0, c1 = 0;
                                for(int r = 0; r < Pn; r++)
                                {
                                    const int2 m = cl->mid[step & 1u][r];
                                    c0 += m.x;
                                    c1 += m.y;
                                }
                                wk.mid = (unsigned long long)c0 + (c1 != 0 ? (1ull << 32) : 0ull);
                                wk.sigma = c1;
                            }
                        }
                        else if(threadIdx.x == 0)
                        {
                            unsigned long long & mbase = wk.base[(step & (kRingSlots - 1)) * kRingWords + kWMid];
                            if(ICP && word_complete(mid_peek, mbase, (unsigned)Pn))
                                wk.mid = (unsigned long long)word_consume(mid_peek, mbase, (unsigned)Pn);
                            else
                                wk.mid = (unsigned long long)poll_word(ring, step, kWMid, (unsigned)Pn, wk);
                            if(RGB_ONLY) wk.sigma = poll_word(ring, step, kWSigma, (unsigned)Pn, wk);
                        }
                        __syncthreads();
                        GN_STAMP(rec, 3);
                        GN_PHASE(5);
                        float sigma_now = 0.f;
                        if(RGB_ONLY)
                        {
                            // count / sigma of the whole image -> rgbError decides the early exit (every CTA, identically)
                            if(threadIdx.x == 0)
                            {
                                sh.total[29] = __int_as_float((int)(wk.mid & 0xffffffffull));
                                sh.total[30] = __int_as_float((int)wk.sigma);
                                gn_sigma(sh, true, rec);
                            }
                            __syncthreads();
                            if(sh.stop)
                            {
                                stop_now = true;   // rgbOnly && rgbError > lastRGBError, RGBDOdometryef.cpp:460-463
                                break;
                            }
                            sigma_now = sh.sigmaVal;
                        }
                        else
                        {
                            // sigmaVal = sqrt(count) needs nothing else (RGBDOdometryef.cpp:457-471; the squared-residual sum is a statistic)
                            const int rgbSize = (int)(wk.mid & 0xffffffffull);
                            const int sel = (rgbSize != 0 && (wk.mid >> 32) == 0ull) ? 1 : rgbSize;
                            sigma_now = __fsqrt_rn((float)sel);
                            if(GEN && rec) rec->sigma_in = sigma_now;
                        }
                        GN_STAMP(rec, 4);
                        if(part)
                        {
                            const LevelGeom g = L.geom[lvl];
                            const int cap = L.plan[lvl].cap;
                            RgbStepArgs a;
                            a.sigma = sigma_now;
                            a.fx = g.fx; a.fy = g.fy;
                            a.sobelScale = L.sobel_scale;
                            a.cols = g.cols; a.rows = g.rows;
                            a.dIdx = wk.P.dIdx; a.dIdy = wk.P.dIdy;
                            a.lastDepth = wk.P.lastDepth;
                            a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
                            a.cloud = nullptr;
                            if(resident)
                            {
                                const int n_rgb = wk.lv_n_rgb;
                                const unsigned * rgb_gxy = reinterpret_cast<const unsigned *>(dyn + L.plan[lvl].off_rgb) + cap;
                                const int * st_zxy = reinterpret_cast<const int *>(dyn + L.off_state);
                                const float * st_diff = reinterpret_cast<const float *>(st_zxy + cap);
                                const float * st_d0 = st_diff + cap;
#pragma unroll 1
                                for(int e = threadIdx.x; e < n_rgb; e += kGnThreads)
                                {
                                    const int zxy = st_zxy[e];
                                    if(zxy >= 0)
                                    {
                                        const unsigned gq = rgb_gxy[e];
                                        float row[7];
                                        rgb_row_regs(a, zxy & 0xffff, zxy >> 16, st_d0[e], (short)(gq & 0xffffu), (short)(gq >> 16), st_diff[e], row);
                                        float a29[29];
#pragma unroll
                                        for(int q = 0; q < 29; q++) a29[q] = acc[q];
                                        accumulate_se3(a29, row, true);
#pragma unroll
                                        for(int q = 0; q < 29; q++) acc[q] = a29[q];
                                    }
                                }
                            }
                            else if(GEN)
                            {
                                const Corres * cimg = wk.P.corres;
                                const int plane = g.rows * g.cols;
#pragma unroll 1
                                for(int k = rank * kGnThreads + threadIdx.x; k < plane; k += G * kGnThreads)
                                {
                                    const int4 raw = *(reinterpret_cast<const int4 *>(cimg) + k);   // written by this very thread above
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
                        }
                        GN_STAMP(rec, 5);
                    }
                    // the count word is fetched while the warp reduction runs: by now every CTA has posted to it long ago, so the
                    // value is normally complete when the reduction is done and nobody waits for this trip through L2
                    if(!CL && ICP && RGB && pass == 0 && threadIdx.x == 0) mid_peek = ld_u64_relaxed(ring_word(ring, step, kWMid));
                    if(part)
                    {
                        // warp level now (one copy of the code for both passes); block level + post once, after the last pass
                        const float s = warp_reduce_scatter32(acc);
                        if(ICP && RGB && pass == 0)
                            s_icp = s;
                        else
                            cta_post_pair<CL>((ICP && RGB) ? s_icp : (ICP ? s : 0.f), RGB ? s : 0.f, ICP, RGB, sh, ring, step, cl, rank, G);
                    }
                    if(pass == 1) GN_PHASE(6);
                }
                if(RGB_ONLY && stop_now)
                {
                    step++;
                    break;
                }

                // ---------------- all sums of the step: warps 0..3 read the words, warp 0 solves
                if(CL)
                {
                    cl_arrive();
                    cl_wait();
                }
                if(threadIdx.x < 128)
                {
                    const int t = (int)threadIdx.x;
                    long long sg = 0;
                    const int sigma_word = (RGB && !RGB_ONLY && t == 127) ? kWSigma : -1;   // thread 127 is never a column reader
                    if(CL)
                    {
                        if(t < 64 && (t & 31) < 29 && (t < 32 ? ICP : RGB)) sh.total[t] = cl_fold_col(cl, (int)(step & 1u), Pn, t);
                        sg = wk.sigma;
                    }
                    else
                    {
                        spin_cycles(L.poll_delay);
                        if(ICP && RGB)
                        {
                            const int kind = t >= 58 ? 1 : 0;
                            read_columns(ring, step, kind ? kWRgb : kWIcp, 29, t - 58 * kind, (unsigned)Pn, sh.total + 32 * kind, wk, sigma_word, &sg);
                        }
                        else
                            read_columns(ring, step, ICP ? kWIcp : kWRgb, 29, t, (unsigned)Pn, sh.total + (ICP ? 0 : 32), wk, sigma_word, &sg);
                    }
                    if(RGB && !RGB_ONLY && t == 127)
                    {
                        const int rgbSize = (int)(wk.mid & 0xffffffffull);
                        sh.total[29] = __int_as_float(rgbSize);
                        sh.total[30] = __int_as_float((int)sg);
                        // what gn_sigma records at the mid-iteration point, from the integer columns
                        sh.rgb_sigma_last = (int)sg;
                        sh.rgb_count_last = rgbSize;
                        sh.res.lastRGBCount = (float)rgbSize;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    GN_STAMP(rec, 6);
                    GN_PHASE(7);
                    if(GEN && rec && RGB && !RGB_ONLY)
                    {
                        rec->rgb_count = __float_as_int(sh.total[29]);
                        rec->rgb_sigma = __float_as_int(sh.total[30]);
                    }
                    if(warp0) warp_update_fast(sh, ICP, RGB, L.icp_weight, rec, t_start);
                    GN_STAMP(rec, 7);
                    GN_PHASE(8);
                    if(GEN && rec) ntr++;
                }
                step++;
                __syncthreads();
                GN_PHASE(9);
            }
        }

        if(CL)
        {
            // ---- hand the running state to the fine-level kernel
            if(rank == 0 && warp0)
            {
                GnHandoff * ho = &ctl->handoff[seq];
                const uint2 * src = reinterpret_cast<const uint2 *>(&sh.res);
                uint2 * dst = reinterpret_cast<uint2 *>(&ho->res);
                for(int k = lane; k < (int)(sizeof(GnResult) / sizeof(uint2)); k += 32) dst[k] = src[k];
                if(lane < 16) ho->resultRt[lane] = sh.resultRt[lane];
                if(lane < 9) ho->Rcurr[lane] = sh.Rcurr[lane];
                if(lane < 3) ho->tcurr[lane] = sh.tcurr[lane];
                if(lane == 0)
                {
                    ho->rgb_sigma_last = sh.rgb_sigma_last;
                    ho->rgb_count_last = sh.rgb_count_last;
                    ho->timeouts = wk.timeouts;
                }
                __threadfence();
                __syncwarp();
                if(lane == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&ctl->handoff_seq[seq]), "l"(handoff_seq) : "memory");
            }
            __syncthreads();   // the next sequence re-initialises the state the hand-off was copied from
            continue;
        }

        // ---- the pose goes out first: the caller is released ~4 us before the statistics are complete
        if(threadIdx.x == 0)
        {
            if(RGB)   // RGBDOdometryef.cpp:579-583 (seq_end repeats this test on the result, harmlessly)
            {
                const float dx = smath::sub(sh.tcurr[0], sh.tprev[0]), dy = smath::sub(sh.tcurr[1], sh.tprev[1]), dz = smath::sub(sh.tcurr[2], sh.tprev[2]);
                const float n = __fsqrt_rn(smath::add(smath::add(smath::mul(dx, dx), smath::mul(dy, dy)), smath::mul(dz, dz)));
                if((double)n > 0.3)
                {
                    for(int k = 0; k < 9; k++) sh.Rcurr[k] = sh.Rprev[k];
                    for(int k = 0; k < 3; k++) sh.tcurr[k] = sh.tprev[k];
                }
            }
            for(int k = 0; k < 9; k++) sh.res.Rcurr[k] = sh.Rcurr[k];
            for(int k = 0; k < 3; k++) sh.res.tcurr[k] = sh.tcurr[k];
        }
        if(rank == 0) __syncthreads();   // the guarded pose is final: warp 0 sends it out while warp 1 does the statistics
        if(rank == 0 && host_results && warp0)
        {
            // The result block goes straight to mapped host memory, followed by a per-sequence flag the host polls: the caller has
            // its pose ~1 us after the last solve instead of after kernel retirement + a D2H copy + a stream synchronisation.
            __syncwarp();
            const uint2 * src = reinterpret_cast<const uint2 *>(&sh.res);
            uint2 * dst = reinterpret_cast<uint2 *>(host_results + seq);
            if(threadIdx.x < 6) dst[threadIdx.x] = src[threadIdx.x];   // Rcurr[9] | tcurr[3]
            __threadfence_system();
            __syncwarp();
            // an inter-CTA wait that gave up poisons the flag: the host reports an error instead of this pose
            if(threadIdx.x == 0) *reinterpret_cast<volatile unsigned *>(host_flags + seq) = wk.timeouts ? (host_seqno | 0x80000000u) : host_seqno;
        }
        // ... the statistics (lastA, lastb, errors, counts) stay in device memory (results[seq]); the host fetches them on demand: no
        // second trip to host memory at the end of the kernel
        if(rank == 0 && wid == 1)
        {
            if(sh.res.gn_iterations > 0) warp_stats_fast(sh, ICP);   // lastA / lastb / ICP error of the last step
            __syncwarp();
            if(lane == 0)
            {
                if(wk.timeouts) sh.res.gn_iterations = -1;   // the host turns this into an error
                seq_end_merge(sh, ICP, RGB, RGB_ONLY, L.so3 || ROLE == 2, &results[seq]);
            }
        }
        if(GEN && leader && L.trace) trace_count[seq] = ntr;
        __syncthreads();
    }
    GN_PHASE(10);
    if(PH && blockIdx.x == 0 && threadIdx.x < 24 && (L.ph_role == 0 || L.ph_role == ROLE || ROLE == 0))
    {
        __syncwarp();
        if(threadIdx.x != 15) atomicAdd(&ctl->phase_cycles[threadIdx.x], (unsigned long long)wk.ph[threadIdx.x]);
        if(threadIdx.x == 0 && !CL) atomicAdd(&ctl->phase_cycles[15], 1ull);   // launches (a split pair counts once)
    }
    if(threadIdx.x == 0 && wk.timeouts) atomicAdd(&ctl->timeouts, 1u);
}

// the kernel variant of a launch
typedef void (*GnKernel)(const GnLaunch, GnCtl *, const GnSeqIn *, const GnSeqIn, unsigned long long *, GnResult *, slam_step_record *, int *, const int, const int,
                         GnResult *, unsigned *, const unsigned, const unsigned long long, const unsigned long long);
static GnKernel gn_pick_kernel(const GnLaunch & L, bool general, bool phases)
{
    if(general || phases)
    {
        // the general variants always account the phases (they are not the tuned path)
        if(L.rgb_only) return general ? k_gn_persistent<false, true, true, true, true, 0> : k_gn_persistent<false, true, true, false, true, 0>;
        if(L.icp && L.rgb) return general ? k_gn_persistent<true, true, false, true, true, 0> : k_gn_persistent<true, true, false, false, true, 0>;
        return general ? k_gn_persistent<true, false, false, true, true, 0> : k_gn_persistent<true, false, false, false, true, 0>;
    }
    if(L.rgb_only) return k_gn_persistent<false, true, true, false, false, 0>;
    if(L.icp && L.rgb) return k_gn_persistent<true, true, false, false, false, 0>;
    return k_gn_persistent<true, false, false, false, false, 0>;
}
// the pair of a split launch: role 1 = cluster (coarse levels), role 2 = fine levels
static GnKernel gn_pick_split_kernel(const GnLaunch & L, bool phases, int role)
{
    if(L.icp && L.rgb)
    {
        if(role == 1) return phases ? k_gn_persistent<true, true, false, false, true, 1> : k_gn_persistent<true, true, false, false, false, 1>;
        return phases ? k_gn_persistent<true, true, false, false, true, 2> : k_gn_persistent<true, true, false, false, false, 2>;
    }
    if(role == 1) return phases ? k_gn_persistent<true, false, false, false, true, 1> : k_gn_persistent<true, false, false, false, false, 1>;
    return phases ? k_gn_persistent<true, false, false, false, true, 2> : k_gn_persistent<true, false, false, false, false, 2>;
}
static const GnKernel kAllGnKernels[] = {
    k_gn_persistent<false, true, true, true, true, 0>,   k_gn_persistent<true, true, false, true, true, 0>,   k_gn_persistent<true, false, false, true, true, 0>,
    k_gn_persistent<false, true, true, false, true, 0>,  k_gn_persistent<true, true, false, false, true, 0>,  k_gn_persistent<true, false, false, false, true, 0>,
    k_gn_persistent<false, true, true, false, false, 0>, k_gn_persistent<true, true, false, false, false, 0>, k_gn_persistent<true, false, false, false, false, 0>,
    k_gn_persistent<true, true, false, false, true, 1>,  k_gn_persistent<true, true, false, false, false, 1>, k_gn_persistent<true, false, false, false, true, 1>,
    k_gn_persistent<true, false, false, false, false, 1>,
    k_gn_persistent<true, true, false, false, true, 2>,  k_gn_persistent<true, true, false, false, false, 2>, k_gn_persistent<true, false, false, false, true, 2>,
    k_gn_persistent<true, false, false, false, false, 2>,
};
static const GnKernel kClusterGnKernels[] = {
    k_gn_persistent<true, true, false, false, true, 1>,  k_gn_persistent<true, true, false, false, false, 1>, k_gn_persistent<true, false, false, false, true, 1>,
    k_gn_persistent<true, false, false, false, false, 1>,
};
// ------------------------------------------------------------------ host side
size_t gn_state_bytes(int batch, int num_sms)
{
    const int groups = batch >= num_sms ? num_sms : batch;
    size_t b = 0;
    b += (sizeof(GnCtl) + 255) / 256 * 256;
    b += (sizeof(GnSeqIn) * batch + 255) / 256 * 256;
    b += kRingBytes * groups;
    b += (sizeof(GnResult) * batch + 255) / 256 * 256;
    b += (sizeof(slam_step_record) * kGnMaxTrace * batch + 255) / 256 * 256;
    b += ((size_t)4 * batch + 255) / 256 * 256;
    return b;
}

void gn_bind_state(GnDevice & d, char * base, int batch, int num_sms)
{
    const int groups = batch >= num_sms ? num_sms : batch;
    d.batch = batch;
    char * p = base;
    d.ctl = (GnCtl *)p;
    p += (sizeof(GnCtl) + 255) / 256 * 256;
    d.seq_in = (GnSeqIn *)p;
    p += (sizeof(GnSeqIn) * batch + 255) / 256 * 256;
    d.ring = (unsigned long long *)p;
    p += kRingBytes * groups;
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
        for(cudaEvent_t e : {d.ev[i], d.ev[i + 1]})
        {
            if(d.ev_pool.size() < 8192)
                d.ev_pool.push_back(e);
            else
                cudaEventDestroy(e);   // (the batched engine brings its own events)
        }
    }
    d.ev.clear();
    return SLAM_OK;
}

// The fine-level kernel of a split pair is not a cooperative launch (a cooperative launch cannot be a programmatic dependent): its
// CTAs are placed one by one.  Two pairs of two HANDLES in flight on one GPU could therefore each hold a part of the SMs and wait
// for the other's (the waits are bounded, so that would end in two errors, not in a hang -- but it must not happen).  When more
// than one handle lives on a device, every pair is ordered behind the previous pair of any other handle with an event; a single
// handle (the normal case: several sequences belong in ONE batched handle) pays nothing.
struct PairGate
{
    std::mutex m;
    std::atomic<int> handles{0};
    cudaEvent_t last = nullptr;
    const GnDevice * owner = nullptr;
};
static PairGate g_pair_gate[64];

void gn_release(GnDevice & d)
{
    if(d.gate_member && d.device >= 0 && d.device < 64)
    {
        PairGate & g = g_pair_gate[d.device];
        std::lock_guard<std::mutex> lock(g.m);
        g.handles--;
        if(g.owner == &d)
        {
            g.owner = nullptr;
            g.last = nullptr;
        }
        d.gate_member = false;
    }
    if(d.pair_done) cudaEventDestroy(d.pair_done);
    d.pair_done = nullptr;
    for(auto e : d.ev) cudaEventDestroy(e);
    d.ev.clear();
    for(auto e : d.ev_pool) cudaEventDestroy(e);
    d.ev_pool.clear();
    if(d.h_stage) cudaFreeHost(d.h_stage);
    d.h_stage = nullptr;
}

static int gn_init_device(GnDevice & d)
{
    if(d.h_stage) return SLAM_OK;
    SLAM_CUDA_TRY(cudaMallocHost((void **)&d.h_stage, d.stage_bytes));
    int dev = 0;
    SLAM_CUDA_TRY(cudaGetDevice(&dev));
    d.device = dev;
    if(!d.gate_member && dev >= 0 && dev < 64)
    {
        g_pair_gate[dev].handles++;
        d.gate_member = true;
    }
    SLAM_CUDA_TRY(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    int coop = 0;
    SLAM_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if(!coop)
    {
        set_last_error("device does not support cooperative launch");
        return SLAM_ERR_UNSUPPORTED;
    }
    if(d.num_sms > kGnMaxCtas) d.num_sms = kGnMaxCtas;
    int optin = 0;
    SLAM_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t static_smem = 0;
    for(GnKernel k : kAllGnKernels)
    {
        cudaFuncAttributes fa;
        SLAM_CUDA_TRY(cudaFuncGetAttributes(&fa, (const void *)k));
        if(fa.sharedSizeBytes > static_smem) static_smem = fa.sharedSizeBytes;
    }
    d.smem_limit = optin - (int)static_smem - 1024;
    if(d.smem_limit < 0) d.smem_limit = 0;
    for(GnKernel k : kAllGnKernels) SLAM_CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_limit));
    const char * sp = getenv("SLAM_GN_SPLIT");
    d.split = sp ? atoi(sp) : 1;
    if(const char * sm = getenv("SLAM_GN_SEQ_MAX")) d.seq_max = atoi(sm);
    if(d.split)
        for(GnKernel k : kClusterGnKernels)
            if(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
            {
                cudaGetLastError();
                d.split = 0;
            }
    d.phases = getenv("SLAM_GN_PHASES") != nullptr;
    const char * pd = getenv("SLAM_GN_POLL_DELAY");
    d.poll_delay = pd ? atoi(pd) : 900;
    return SLAM_OK;
}

static bool gn_make_plan_for(GnDevice & d, GnLaunch & L, const int G, const bool cluster);
static bool gn_make_split(GnDevice & d, const GnLaunch & L, GnLaunch & La, GnLaunch & Lb);
int gn_configure(GnDevice & d) { return gn_init_device(d); }

bool gn_split_applies(GnDevice & d, const GnLaunch & L)
{
    if(gn_init_device(d) != SLAM_OK) return false;
    GnLaunch La, Lb;
    return !L.trace && !L.full_corres && gn_make_split(d, L, La, Lb);
}

bool gn_make_plan(GnDevice & d, GnLaunch & L)
{
    if(gn_init_device(d) != SLAM_OK) d.smem_limit = 0;
    return gn_make_plan_for(d, L, gn_group_size(d.num_sms, L.batch), false);
}

static bool gn_make_plan_for(GnDevice & d, GnLaunch & L, const int G, const bool cluster)
{
    static const bool no_resident = getenv("SLAM_GN_STREAMED") != nullptr;   // development aid: force the streamed path
    int off = 0;
    auto take = [&](int bytes) {
        const int o = off;
        off = (off + bytes + 15) & ~15;
        return o;
    };
    // candidates in order of benefit: level 0 carries most iterations and pixels
    int cap_state = 0;
    bool all_rgb_resident = true;
    for(int l = 0; l < L.levels; l++)
    {
        LevelPlan & pl = L.plan[l];
        const int plane = L.geom[l].rows * L.geom[l].cols;
        pl.nseg = (plane + 31) / 32;
        pl.P = G;
        pl.segs_per_cta = 0;
        pl.cap = 0;
        pl.off_icp = pl.off_rgb = 0;
        pl.resident = 0;
        if(L.iterations[l] <= 0) continue;
        int P = (pl.nseg + kGnWarps - 1) / kGnWarps;
        if(P > G) P = G;
        const int segs = (pl.nseg + P - 1) / P;
        const int nslots = (segs + kGnWarps - 1) / kGnWarps;
        const int cap = segs * 32;   // list capacity: every pixel of the CTA's segments
        const int need = (L.icp ? 24 * cap : 0) + (L.rgb ? 12 * cap : 0);
        const int state_now = L.rgb ? 12 * (cap > cap_state ? cap : cap_state) : 0;
        const int state_before = L.rgb ? 12 * cap_state : 0;
        const bool fits = !no_resident && nslots <= kMaxStageSlots && L.geom[l].cols <= 2048 && L.geom[l].rows <= 2048 && off + need + 64 + (state_now - state_before) + state_before <= d.smem_limit;
        if(fits)
        {
            pl.resident = 1;
            pl.P = P;
            pl.segs_per_cta = segs;
            pl.cap = cap;
            if(L.icp) pl.off_icp = take(24 * cap);
            if(L.rgb) pl.off_rgb = take(12 * cap);
            if(cap > cap_state) cap_state = cap;
        }
        else if(L.rgb)
            all_rgb_resident = false;
    }
    L.off_state = L.rgb ? take(12 * cap_state) : 0;
    L.so3_resident = 0;
    L.so3_P = G;
    L.off_so3 = 0;
    if(L.so3 && L.levels >= 3)
    {
        const int N = L.geom[2].rows * L.geom[2].cols;
        int P = (N + kGnThreads - 1) / kGnThreads;
        if(P > G) P = G;
        L.so3_P = P;
        const int bytes = 2 * ((N + 15) & ~15);
        if(!no_resident && off + bytes <= d.smem_limit)
        {
            L.so3_resident = 1;
            L.off_so3 = take(bytes);
        }
    }
    for(int l = 0; l < L.levels; l++)
    {
        double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        K[0] = L.geom[l].fx; K[4] = L.geom[l].fy; K[2] = L.geom[l].cx; K[5] = L.geom[l].cy; K[8] = 1;   // k_matrix_d
        smath::mat3_inverse(K, L.Kinv[l]);
        for(int k = 0; k < 9; k++) L.K[l][k] = K[k];
    }
    L.off_cl = 0;
    if(cluster)
    {
        L.off_cl = take((int)sizeof(ClArea));
        if(off > d.smem_limit) all_rgb_resident = false;   // (never with the levels a cluster takes)
    }
    L.dyn_bytes = off;
    L.poll_delay = d.poll_delay;
    {
        static const int ph_role = [] {
            const char * ph = getenv("SLAM_GN_PHASES");
            const int r = ph ? atoi(ph) - 1 : 0;
            return (r < 0 || r > 2) ? 0 : r;
        }();
        L.ph_role = ph_role;
    }
    return all_rgb_resident;
}

// Split launch: which levels the cluster takes (the coarse ones, from the top down to the first level that is too large) and the
// two shared-memory plans.  Returns false when the launch does not qualify (then the whole frame runs in one launch).
static bool gn_make_split(GnDevice & d, const GnLaunch & L, GnLaunch & La, GnLaunch & Lb)
{
    if(d.split != 1 || d.split_broken || L.batch > kSplitMaxSeqs || (L.batch > 1 && (L.batch < 3 || L.batch > d.seq_max)) || L.rgb_only || !L.icp || L.trace || L.full_corres) return false;
    if(d.num_sms < kClusterCtas + 32) return false;
    // Which levels go with the SO3 pre-alignment onto the cluster: by default none.  Measured (tools/iter_cost.py): an ICP+RGB iteration
    // of the 160x120 level costs 6.8 us on the 16 CTAs of a cluster against 5.2 us on 38 CTAs with the reduction words in L2 (the
    // map phases grow by more than the shorter all-reduce saves), an SO3 iteration 3.4 against 4.0 us.  SLAM_GN_SPLIT_LEVELS=1 moves
    // every level of at most kClusterMaxPixels pixels as well (development aid).
    static const bool coarse_too = getenv("SLAM_GN_SPLIT_LEVELS") && atoi(getenv("SLAM_GN_SPLIT_LEVELS")) > 0;
    int first_fine = L.levels - 1;   // levels [0, first_fine] stay with the fine kernel
    if(coarse_too)
        while(first_fine >= 0 && L.geom[first_fine].rows * L.geom[first_fine].cols <= kClusterMaxPixels) first_fine--;
    if(first_fine < 0) return false;                     // nothing for the fine kernel
    if(first_fine == L.levels - 1 && !L.so3) return false;   // nothing for the cluster
    if(L.so3 && first_fine >= 2 && first_fine != L.levels - 1) return false;   // a cluster that takes levels takes level 2 (SO3 runs there)
    int it_a = 0, it_b = 0;
    La = L;
    Lb = L;
    for(int l = 0; l < L.levels; l++)
    {
        if(l <= first_fine)
        {
            La.iterations[l] = 0;
            it_b += L.iterations[l];
        }
        else
        {
            Lb.iterations[l] = 0;
            it_a += L.iterations[l];
        }
    }
    if((it_a <= 0 && !L.so3) || it_b <= 0) return false;
    Lb.so3 = false;
    gn_make_plan_for(d, La, kClusterCtas, true);
    gn_make_plan_for(d, Lb, d.num_sms - kClusterCtas, false);
    for(int l = 0; l < L.levels; l++)
    {
        if(La.iterations[l] > 0 && !La.plan[l].resident) return false;
        if(Lb.iterations[l] > 0 && !Lb.plan[l].resident) return false;
    }
    if(La.so3 && !La.so3_resident) return false;
    return true;
}

// Fill the pinned staging image of the per-sequence input blocks (pointers + prior pose).
int gn_stage_inputs(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnSeqIn ** out)
{
    if(int rc = gn_init_device(d)) return rc;
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
    *out = in;
    return SLAM_OK;
}

int gn_enqueue(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnResult * h_results,
               cudaStream_t stream, unsigned * h_flags, unsigned seqno)
{
    GnSeqIn * in = nullptr;
    if(int rc = gn_stage_inputs(d, L, seqs, trans, rot, &in)) return rc;
    // one sequence: its pointer / pose block travels as a kernel parameter; several: one H2D copy of the array
    if(L.batch > 1) SLAM_CUDA_TRY(cudaMemcpyAsync(d.seq_in, d.h_stage, sizeof(GnSeqIn) * L.batch, cudaMemcpyHostToDevice, stream));

    // group geometry: every sequence gets its own group of G CTAs while they fit
    int G = gn_group_size(d.num_sms, L.batch);
    int groups = L.batch >= d.num_sms ? d.num_sms : L.batch;
    GnLaunch Lc = L;
    GnCtl * ctl = d.ctl;
    const GnSeqIn * seq_in = d.seq_in;
    unsigned long long * ring = d.ring;
    GnResult * results = d.results;
    slam_step_record * trace = d.trace;
    int * trace_count = d.trace_count;
    GnSeqIn seq0 = in[0];
    // h_flags != nullptr: h_results / h_flags are mapped pinned memory the kernel writes itself (device view == host pointer under UVA)
    GnResult * host_results = h_flags ? h_results : nullptr;
    d.launch_no++;
    unsigned long long handoff_seq = d.launch_no;
    // the tuned variant needs every level (and the SO3 images) resident and no step trace
    bool general = L.trace || L.full_corres || (L.so3 && !L.so3_resident);
    for(int l = 0; l < L.levels; l++) general = general || (L.iterations[l] > 0 && !L.plan[l].resident);
    // the split pair has its own shared-memory plans (132 CTAs for ALL sequences of a small batch, where the one-launch form would give
    // every sequence a fraction of the CTAs and stream its fine levels)
    GnLaunch La, Lb;
    const bool split = !L.trace && !L.full_corres && gn_make_split(d, L, La, Lb);
    if(split)
    {
        // one group for the whole batch: the pair works through the sequences one after the other (the cluster runs ahead with the
        // SO3 pre-alignments, the fine-level kernel finds every hand-off waiting but the first)
        G = d.num_sms - kClusterCtas;
        groups = 1;
        Lc = Lb;
    }
    // every launch raises the check-in counter of each group by G (GN_GATE)
    d.gate_total += (unsigned long long)G;
    unsigned long long gate_target = d.gate_total;
    void * args[] = {&Lc, &ctl, &seq_in, &seq0, &ring, &results, &trace, &trace_count, &G, &groups, &host_results, &h_flags, &seqno, &gate_target, &handoff_seq};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if(d.profiling)
    {
        if(d.ev.size() >= 4096)
            if(int rc = gn_fold_profile(d)) return rc;
        for(cudaEvent_t * e : {&e0, &e1})
        {
            if(d.ev_pool.empty())
                SLAM_CUDA_TRY(cudaEventCreate(e));
            else
            {
                *e = d.ev_pool.back();
                d.ev_pool.pop_back();
            }
        }
        SLAM_CUDA_TRY(cudaEventRecord(e0, stream));
    }
    const GnKernel kernel = split ? gn_pick_split_kernel(L, d.phases, 2) : gn_pick_kernel(L, general, d.phases);
    static const bool debug = getenv("SLAM_ODOM_DEBUG") != nullptr;
    if(debug && d.launch_no < 3)
    {
        fprintf(stderr, "gn_enqueue: general %d phases %d dyn %d of %d, so3 resident %d P %d, levels:", (int)general, (int)d.phases, L.dyn_bytes, d.smem_limit, L.so3_resident, L.so3_P);
        for(int l = 0; l < L.levels; l++) fprintf(stderr, " [%d: it %d res %d P %d cap %d]", l, L.iterations[l], L.plan[l].resident, L.plan[l].P, L.plan[l].cap);
        fprintf(stderr, "\n");
    }
    d.last_launches = 1;
    // several handles on this device: pairs of different handles run one after the other (see PairGate)
    static const bool gate_off = getenv("SLAM_GN_PAIR_GATE") && atoi(getenv("SLAM_GN_PAIR_GATE")) == 0;   // development aid
    PairGate * gate = (split && !gate_off && d.device >= 0 && d.device < 64 && g_pair_gate[d.device].handles.load() > 1) ? &g_pair_gate[d.device] : nullptr;
    std::unique_lock<std::mutex> gate_lock;
    if(gate)
    {
        gate_lock = std::unique_lock<std::mutex>(gate->m);
        if(gate->last && gate->owner != &d) SLAM_CUDA_TRY(cudaStreamWaitEvent(stream, gate->last, 0));
    }
    if(split)
    {
        d.last_launches = 2;
        // 1) the cluster: SO3 pre-alignment + coarse levels, one cluster of kClusterCtas CTAs
        int Ga = kClusterCtas, groups_a = 1;
        void * args_a[] = {&La, &ctl, &seq_in, &seq0, &ring, &results, &trace, &trace_count, &Ga, &groups_a, &host_results, &h_flags, &seqno, &gate_target, &handoff_seq};
        cudaLaunchConfig_t ca = {};
        ca.gridDim = dim3(kClusterCtas);
        ca.blockDim = dim3(kGnThreads);
        ca.dynamicSmemBytes = (size_t)La.dyn_bytes;
        ca.stream = stream;
        cudaLaunchAttribute aa[2];
        aa[0].id = cudaLaunchAttributeClusterDimension;
        aa[0].val.clusterDim.x = kClusterCtas;
        aa[0].val.clusterDim.y = 1;
        aa[0].val.clusterDim.z = 1;
        aa[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        aa[1].val.programmaticStreamSerializationAllowed = 1;
        ca.attrs = aa;
        static const bool chain = !(getenv("SLAM_GN_PDL_CHAIN") && atoi(getenv("SLAM_GN_PDL_CHAIN")) == 0);
        ca.numAttrs = chain ? 2 : 1;
        cudaError_t e = cudaLaunchKernelExC(&ca, (const void *)gn_pick_split_kernel(L, d.phases, 1), args_a);
        if(e == cudaSuccess)
        {
            // 2) the fine levels on the other SMs: released as soon as every CTA of the cluster has started (programmatic dependent
            //    launch), so that it can never take the SMs the cluster needs; not cooperative -- its CTAs are co-resident because
            //    the grid is the number of SMs the cluster leaves (if they are not, the late ones start when the cluster retires)
            cudaLaunchConfig_t cb = {};
            cb.gridDim = dim3(G);
            cb.blockDim = dim3(kGnThreads);
            cb.dynamicSmemBytes = (size_t)Lb.dyn_bytes;
            cb.stream = stream;
            cudaLaunchAttribute ab[1];
            ab[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            ab[0].val.programmaticStreamSerializationAllowed = 1;
            cb.attrs = ab;
            cb.numAttrs = 1;
            e = cudaLaunchKernelExC(&cb, (const void *)kernel, args);
            if(e != cudaSuccess)
            {
                // the cluster is already in the stream: its partner follows as an ordinary launch (it finds the hand-off waiting)
                cudaGetLastError();
                d.split_broken = 1;
                cb.numAttrs = 0;
                SLAM_CUDA_TRY(cudaLaunchKernelExC(&cb, (const void *)kernel, args));
            }
        }
        else
        {
            // no cluster on this device / in this context: one launch for the whole frame, from now on
            cudaGetLastError();
            d.split_broken = 1;
            if(general && L.derive_gradients)
            {
                // (the caller planned for resident levels and made no derivative images; its next call plans without the pair)
                set_last_error("the cluster launch of the split Gauss-Newton pair failed; call again");
                return SLAM_ERR_CUDA;
            }
            d.last_launches = 1;
            d.gate_total -= (unsigned long long)G;
            G = gn_group_size(d.num_sms, L.batch);
            groups = L.batch >= d.num_sms ? d.num_sms : L.batch;
            d.gate_total += (unsigned long long)G;
            gate_target = d.gate_total;
            Lc = L;
            SLAM_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)gn_pick_kernel(L, general, d.phases), dim3(G * groups), dim3(kGnThreads), args, (size_t)L.dyn_bytes, stream));
        }
    }
    else
        SLAM_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kernel, dim3(G * groups), dim3(kGnThreads), args, (size_t)L.dyn_bytes, stream));
    if(gate)
    {
        if(!d.pair_done) SLAM_CUDA_TRY(cudaEventCreateWithFlags(&d.pair_done, cudaEventDisableTiming));
        SLAM_CUDA_TRY(cudaEventRecord(d.pair_done, stream));
        gate->last = d.pair_done;
        gate->owner = &d;
    }
    if(d.profiling)
    {
        SLAM_CUDA_TRY(cudaEventRecord(e1, stream));
        d.ev.push_back(e0);
        d.ev.push_back(e1);
    }
    if(!h_flags) SLAM_CUDA_TRY(cudaMemcpyAsync(h_results, d.results, sizeof(GnResult) * L.batch, cudaMemcpyDeviceToHost, stream));
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
