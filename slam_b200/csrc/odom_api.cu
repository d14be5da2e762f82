// libslam_odom: handle, buffers and the C ABI of the tracker (include/slam_odom.h).
//
// Host side of the reference's RGBDOdometryef (src/odom/RGBDOdometryef.cpp) re-designed
// for B200: one device arena per handle, dense buffers, one stream, fused preparation
// kernels, and two ways of running the Gauss-Newton iterations:
//   * device-resident (default): the whole SO3 + coarse-to-fine ICP/RGB loop, including
//     the 3x3 / 6x6 solves and the pose update, runs inside ONE persistent kernel
//     (gn_kernel.cu); the host only enqueues it and reads back 48 bytes of pose;
//   * host-stepped (params.host_loop = 1): the reference's control flow, one reduction
//     launch + one sync per step, kept for step-by-step parity checks against the reference replay.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <limits>
#include "odom_internal.hpp"
#include "gn_kernel.cuh"
#include "batch_engine.cuh"

namespace slam {

static thread_local std::string g_last_error;
void set_last_error(const std::string & msg) { g_last_error = msg; }

}   // namespace slam

using namespace slam;

struct StagingSlot
{
    unsigned short * depth = nullptr;
    uchar4 * rgba = nullptr;
    float4 * mv = nullptr;
    float4 * mn = nullptr;
    uchar4 * mrgba = nullptr;
    float pose[16 * 1];
    std::vector<float> poses;
    const void * tag_depth = nullptr;   // host pointer this slot was prefetched from
    unsigned long long generation = 0;  // staging order (the older slot is the one a stale prefetch is dropped from)
    cudaEvent_t ready = nullptr;
    bool pending = false;
    float depth_cutoff = 0, model_depth_cutoff = 0;
};

// ---- cross-GPU minimum over peer memory (NVLink / NVSwitch), for the sharded hypothesis scoring
// Every rank owns an array of slots, one per rank and frame parity, in its own HBM; the arrays are exchanged as CUDA IPC handles
// once.  Per frame ONE warp per rank writes (key, frame number) into its slot on every peer -- the frame number with release
// semantics at system scope, after the key -- and then polls its own array until every peer's frame number has arrived: an
// all-gather of 16 bytes per peer and a min, in the launch that follows the scoring launch, with no collective library and no host
// round trip in between.  Two parities: a rank may already publish frame s + 1 while a slow peer still reads frame s (it cannot get
// to s + 2 without that peer's s + 1).
constexpr int kMaxPeers = 16;
struct PeerSlot
{
    unsigned long long key, seq;
};
struct PeerPtrs
{
    PeerSlot * p[kMaxPeers];
};

__global__ void __launch_bounds__(32) k_peer_min(const PeerPtrs peers, const int me, const int world, const unsigned long long * my_key, const unsigned long long seq,
                                                 PeerSlot * local, unsigned long long * result)
{
    const int r = threadIdx.x;
    unsigned long long k = ~0ull;
    bool gave_up = false;
    if(r < world)
    {
        const unsigned long long mine = *my_key;   // left by the scoring launch in front of this one
        PeerSlot * dst = peers.p[r] + (seq & 1ull) * kMaxPeers + me;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&dst->key), "l"(mine) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(seq) : "memory");
        const PeerSlot * src = local + (seq & 1ull) * kMaxPeers + r;
        unsigned long long got = 0ull, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for(;;)
        {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(&src->seq) : "memory");
            if(got == seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if(t1 - t0 > 10000000000ull)   // 10 s: a peer that never publishes must not hang this GPU
            {
                gave_up = true;
                break;
            }
        }
        if(!gave_up) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(k) : "l"(&src->key) : "memory");
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
    }
    const unsigned lost = __ballot_sync(0xffffffffu, gave_up);
    if(r == 0)
    {
        result[0] = k;
        result[2] = (unsigned long long)__popc(lost);
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(result + 1) = seq;
    }
}

struct slam_odom
{
    slam_odom_params p;
    int levels = 3;
    int batch = 1;
    LevelGeom geom[SLAM_MAX_LEVELS];
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // second branch of the per-frame preparation (depth pyramid / maps)
    cudaEvent_t compute_done = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    int num_sms = 0;

    char * arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<SeqBuffers> seq;
    size_t seq_stride = 0;            // bytes between the buffers of consecutive sequences in the arena
    unsigned short * filtered_depth = nullptr;   // [batch][H][W] output of the depth pre-filter (allocated on first use)
    char * score_ws = nullptr;        // pose-hypothesis scoring: poses | partials | tickets | results (grown on demand)
    size_t score_ws_bytes = 0;
    float * h_score_poses = nullptr;  // pinned staging of the hypotheses
    // hypothesis scoring across GPUs over peer memory (slam_odom_peer_export / _connect / _score_poses_best_peers)
    PeerSlot * peer_local = nullptr;              // [2][kMaxPeers] slots the other ranks write into (IPC-exported allocation)
    PeerSlot * peer_ptr[kMaxPeers] = {};          // rank r's slot array as mapped into this process (own rank: peer_local)
    int peer_rank = -1, peer_world = 0;
    unsigned long long peer_seq = 0;
    unsigned long long * d_peer_key = nullptr;    // this rank's best key of the frame being scored
    unsigned long long * h_peer_result = nullptr; // mapped pinned: [0] winner key, [1] sequence number, [2] polls that gave up, [3] INT64_MAX (source of the reset copy)
    unsigned long long * d_peer_result = nullptr;
    int score_ws_n = -1, score_ws_plane = -1;   // layout the workspace was last cleared for
    float * d_poses12 = nullptr;      // [batch][12] model poses (R row-major | t) of the batched preparation launches
    float * h_poses12 = nullptr;      // pinned staging of the same

    // device-resident loop
    GnDevice gn;                      // device pointers of the persistent kernel's state
    BatchDevice be;                   // batched streaming engine (batch >= be_min)
    int be_min = kBatchEngineMin;     // SLAM_BATCH_ENGINE_MIN overrides (development aid)
    GnResult * h_results = nullptr;   // pinned, mapped [batch]
    unsigned * h_flags = nullptr;     // pinned, mapped [2][batch]: sequence numbers written by the persistent kernel: [0] pose out, [1] statistics out
    bool stats_lazy = false;          // the statistics of the last track are still to be taken from h_results (second flag)
    unsigned zc_seqno = 0;            // sequence number of the launch in flight
    bool zero_copy = false;           // results + completion flag written by the kernel itself (single-launch loop only)
    bool zc_pending = false;
    float * h_sums = nullptr;         // pinned scratch for the host-stepped loop [96]

    std::vector<slam_odom_stats> stats;
    std::vector<float> last_pose;     // [batch][12]: rot9 | trans3 of the last collected track (slam_odom_wait after an implicit collection)
    std::vector<std::vector<slam_step_record>> trace;
    bool trace_on = false;
    int trace_level = 0;
    bool have_depth_tmp = false;
    bool pending_async = false;
    bool last_icp = false, last_rgb = false, last_so3 = false;
    bool pend_icp = false, pend_rgb = false, pend_so3 = false;   // terms computed by the launches whose statistics have not been merged into `stats` yet
    bool deriv_src_swapped = false;   // the frame the stale derivative images would belong to sits in lastNextImage (swap after an SO3 call)
    bool deriv_stale = true;          // dIdx / dIdy do not belong to the current nextImage pyramid (the persistent kernel derived its own gradients)
    long long launches = 0;

    StagingSlot slot[2];
    unsigned long long stage_generation = 0;
    int next_slot = 0;
    bool staging_ready = false;

    // constants of the reference ctor (RGBDOdometryef.cpp:34-37,108-110)
    float sobelScale = 1.0f / 8.0f;
    float maxDepthDeltaRGB = 0.07f;
    float maxDepthRGB = 6.0f;
    float minGrad[SLAM_MAX_LEVELS] = {5, 3, 1, 1};
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ArenaPlan
{
    size_t off = 0;
    size_t take(size_t bytes)
    {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

int check_handle(slam_odom_t h)
{
    if(!h)
    {
        set_last_error("null handle");
        return SLAM_ERR_ARG;
    }
    return SLAM_OK;
}

int set_device(slam_odom_t h)
{
    SLAM_CUDA_TRY(cudaSetDevice(h->p.device));
    return SLAM_OK;
}

// Lay out (plan == true: only measure) every per-sequence buffer inside the arena.
void layout_sequence(slam_odom * h, ArenaPlan & plan, SeqBuffers * sb)
{
    char * base = h->arena;
    auto P = [&](size_t bytes) -> char * {
        const size_t o = plan.take(bytes);
        return base ? base + o : nullptr;
    };
    SeqBuffers tmp;
    SeqBuffers & s = sb ? *sb : tmp;
    for(int l = 0; l < h->levels; l++)
    {
        const size_t n = (size_t)h->geom[l].rows * h->geom[l].cols;
        s.depth[l] = (unsigned short *)P(n * 2);
        s.vcurr[l] = (float *)P(n * 12);
        s.ncurr[l] = (float *)P(n * 12);
        s.vprev[l] = (float *)P(n * 12);
        s.nprev[l] = (float *)P(n * 12);
        s.lastDepth[l] = (float *)P(n * 4);
        s.nextDepth[l] = (float *)P(n * 4);
        s.lastImage[l] = (unsigned char *)P(n);
        s.nextImage[l] = (unsigned char *)P(n);
        s.lastNextImage[l] = (unsigned char *)P(n);
        s.dIdx[l] = (short *)P(n * 2);
        s.dIdy[l] = (short *)P(n * 2);
        s.corres[l] = (Corres *)P(n * sizeof(Corres));
    }
    const size_t n0 = (size_t)h->geom[0].rows * h->geom[0].cols;
    s.depth_tmp = (float *)P(n0 * 4);
    if(h->levels > 3)
    {
        const size_t n2 = (size_t)h->geom[2].rows * h->geom[2].cols;
        s.vcam = (float *)P(n2 * 12);
        s.ncam = (float *)P(n2 * 12);
    }
    else
        s.vcam = s.ncam = nullptr;
    s.workspace = P(kWorkspaceBytes);
    s.sums = (float *)P(128 * 4);
}

void default_iterations(const slam_odom * h, int pyramid, int fast_odom, int * it)
{
    bool user = false;
    for(int l = 0; l < SLAM_MAX_LEVELS; l++) user = user || h->p.iterations[l] != 0;
    for(int l = 0; l < SLAM_MAX_LEVELS; l++) it[l] = 0;
    if(user)
    {
        for(int l = 0; l < h->levels; l++) it[l] = h->p.iterations[l];
        return;
    }
    // RGBDOdometryef.cpp:382-384
    it[0] = fast_odom ? 3 : 10;
    if(h->levels > 1) it[1] = pyramid ? 5 : 0;
    if(h->levels > 2) it[2] = pyramid ? 4 : 0;
    if(h->levels > 3) it[3] = pyramid ? 4 : 0;
}

void unpack_se3(const float * s, float * A, float * b)   // reduce.cu:472-486
{
    int shift = 0;
    for(int i = 0; i < 6; ++i)
        for(int j = i; j < 7; ++j)
        {
            const float value = s[shift++];
            if(j == 6)
                b[i] = value;
            else
                A[j * 6 + i] = A[i * 6 + j] = value;
        }
}

void unpack_so3(const float * s, float * A, float * b)   // reduce.cu:1122-1136
{
    int shift = 0;
    for(int i = 0; i < 3; ++i)
        for(int j = i; j < 4; ++j)
        {
            const float value = s[shift++];
            if(j == 3)
                b[i] = value;
            else
                A[j * 3 + i] = A[i * 3 + j] = value;
        }
}

void k_matrix(const LevelGeom & g, double * K)
{
    for(int i = 0; i < 9; i++) K[i] = 0;
    K[0] = g.fx;
    K[4] = g.fy;
    K[2] = g.cx;
    K[5] = g.cy;
    K[8] = 1;
}

}   // namespace

// ------------------------------------------------------------------ frame preparation
namespace {

int enqueue_init_icp_depth(slam_odom * h, int b, const uint16_t * d_depth, size_t pitch, float cutoff)
{
    SeqBuffers & s = h->seq[b];
    const LevelGeom & g0 = h->geom[0];
    const size_t row_bytes = (size_t)g0.cols * 2;
    const unsigned short * src = d_depth;
    if(pitch != 0 && pitch != row_bytes)
    {
        SLAM_CUDA_TRY(cudaMemcpy2DAsync(s.depth[0], row_bytes, d_depth, pitch, row_bytes, g0.rows, cudaMemcpyDeviceToDevice, h->stream));
        src = s.depth[0];
    }
    for(int l = 0; l < h->levels; l++)
    {
        const LevelGeom & g = h->geom[l];
        const unsigned short * in = (l == 0) ? src : s.depth[l];
        unsigned short * next = (l + 1 < h->levels) ? s.depth[l + 1] : nullptr;
        int rc = launch_depth_level(in, g.rows, g.cols, g.fx, g.fy, g.cx, g.cy, cutoff, s.vcurr[l], s.ncurr[l], next, h->stream);
        if(rc) return rc;
        h->launches++;
    }
    if(src != s.depth[0] && h->trace_on)   // keep a copy for the depth tap
        SLAM_CUDA_TRY(cudaMemcpyAsync(s.depth[0], src, row_bytes * g0.rows, cudaMemcpyDeviceToDevice, h->stream));
    return SLAM_OK;
}

int enqueue_model_maps(slam_odom * h, int b, const float * v4, const float * n4, bool model, const float * pose16)
{
    SeqBuffers & s = h->seq[b];
    const LevelGeom & g0 = h->geom[0];
    Mat3 R = {};
    float3 t = make_float3(0, 0, 0);
    if(model)
    {
        R.r0 = make_float3(pose16[0], pose16[1], pose16[2]);
        R.r1 = make_float3(pose16[4], pose16[5], pose16[6]);
        R.r2 = make_float3(pose16[8], pose16[9], pose16[10]);
        t = make_float3(pose16[3], pose16[7], pose16[11]);
    }
    float ** vdst = model ? s.vprev : s.vcurr;
    float ** ndst = model ? s.nprev : s.ncurr;
    int rc = launch_model_maps_simple((const float4 *)v4, (const float4 *)n4, g0.rows, g0.cols, h->levels, vdst, ndst, model ? 1 : 0, R, t, s.depth_tmp,
                                      h->maxDepthRGB, h->levels > 3 ? s.vcam : nullptr, h->levels > 3 ? s.ncam : nullptr, h->stream);
    if(rc) return rc;
    h->launches++;
    if(h->levels > 3)
    {
        rc = launch_resize_transform(s.vcam, s.ncam, h->geom[2].rows, h->geom[2].cols, vdst[3], ndst[3], model ? 1 : 0, R, t, nullptr, nullptr, h->stream);
        if(rc) return rc;
        h->launches++;
    }
    return SLAM_OK;
}

// populateRGBDData, RGBDOdometryef.cpp:208-235 (destDepths may be null for initFirstRGB)
int enqueue_populate_rgbd(slam_odom * h, int b, const uint8_t * rgba, float ** destDepths, unsigned char ** destImages)
{
    SeqBuffers & s = h->seq[b];
    const LevelGeom & g0 = h->geom[0];
    int rc = launch_rgbd_level0(s.depth_tmp, destDepths ? destDepths[0] : nullptr, (const uchar4 *)rgba, destImages[0], g0.rows * g0.cols, h->stream);
    if(rc) return rc;
    h->launches++;
    for(int l = 0; l + 1 < h->levels; l++)
    {
        const LevelGeom & g = h->geom[l];
        rc = launch_rgbd_down(destDepths ? destDepths[l] : nullptr, destDepths ? destDepths[l + 1] : nullptr, destImages[l], destImages[l + 1], g.rows,
                              g.cols, h->stream);
        if(rc) return rc;
        h->launches++;
    }
    return SLAM_OK;
}

int enqueue_derivatives(slam_odom * h, int b)
{
    SeqBuffers & s = h->seq[b];
    int rows[SLAM_MAX_LEVELS], cols[SLAM_MAX_LEVELS];
    for(int l = 0; l < h->levels; l++)
    {
        rows[l] = h->geom[l].rows;
        cols[l] = h->geom[l].cols;
    }
    int rc = launch_derivatives_simple(h->levels, s.nextImage, s.dIdx, s.dIdy, rows, cols, h->stream);
    if(rc) return rc;
    h->launches++;
    if(b == h->batch - 1) h->deriv_stale = false;
    return SLAM_OK;
}

}   // namespace

// ------------------------------------------------------------------ host-stepped loop
namespace {

int host_loop_one(slam_odom * h, int b, float * trans, float * rot, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3)
{
    SeqBuffers & s = h->seq[b];
    slam_odom_stats & st = h->stats[b];
    std::vector<slam_step_record> * tr = h->trace_on ? &h->trace[b] : nullptr;
    if(tr) tr->clear();

    const bool icp = !rgbOnly && icpWeight > 0;
    const bool rgb = rgbOnly || icpWeight < 100;

    float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
    memcpy(Rprev, rot, sizeof(Rprev));
    memcpy(tprev, trans, sizeof(tprev));
    memcpy(Rcurr, Rprev, sizeof(Rprev));
    memcpy(tcurr, tprev, sizeof(tprev));

    if(rgb)
    {
        int rc = enqueue_derivatives(h, b);
        if(rc) return rc;
    }

    float * d_icp = s.sums;
    float * d_rgb = s.sums + 32;
    float * d_so3 = s.sums + 64;
    int * d_res = reinterpret_cast<int *>(s.sums + 80);

    double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    st.so3_iterations = 0;
    st.gn_iterations = 0;

    if(so3)
    {
        if(h->levels < 3)
        {
            set_last_error("so3 pre-alignment needs pyramid level 2 (num_levels >= 3)");
            return SLAM_ERR_UNSUPPORTED;
        }
        const int L = 2;   // RGBDOdometryef.cpp:296
        const LevelGeom & g = h->geom[L];
        float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        double K[9], Kinv[9];
        k_matrix(g, K);
        smath::mat3_inverse(K, Kinv);

        float lastError = std::numeric_limits<float>::max() / 2;
        float lastCount = std::numeric_limits<float>::max() / 2;
        double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        for(int i = 0; i < 10; i++)
        {
            double KR[9], H[9];
            smath::mat3_mul(K, resultR, KR);
            smath::mat3_mul(KR, Kinv, H);
            float Hf[9], Kinvf[9], KRf[9];
            for(int k = 0; k < 9; k++)
            {
                Hf[k] = (float)H[k];
                Kinvf[k] = (float)Kinv[k];
                KRf[k] = (float)KR[k];
            }
            So3Args a;
            a.lastImage = s.lastNextImage[L];
            a.nextImage = s.nextImage[L];
            a.imageBasis = mat3_from(Hf);
            a.kinv = mat3_from(Kinvf);
            a.krlr = mat3_from(KRf);
            a.cols = g.cols;
            a.rows = g.rows;
            int rc = launch_so3_step(a, s.workspace, d_so3, h->stream);
            if(rc) return rc;
            h->launches++;
            SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_sums, d_so3, 11 * 4, cudaMemcpyDeviceToHost, h->stream));
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
            st.so3_iterations++;

            float jtj[9], jtr[3];
            unpack_so3(h->h_sums, jtj, jtr);
            const float residual0 = h->h_sums[9], residual1 = h->h_sums[10];

            st.lastSO3Error = sqrtf(residual0) / residual1;
            st.lastSO3Count = residual1;

            slam_step_record rec = {};
            rec.kind = 0;
            rec.level = L;
            rec.iteration = i;
            memcpy(rec.so3, h->h_sums, 11 * 4);
            memcpy(rec.so3_in, Hf, 36);
            memcpy(rec.so3_in + 9, Kinvf, 36);
            memcpy(rec.so3_in + 18, KRf, 36);

            bool stop = false;
            if(st.lastSO3Error < lastError && lastCount == st.lastSO3Count)
                stop = true;   // converged
            else if((double)st.lastSO3Error > (double)lastError + 0.001)   // diverging
            {
                st.lastSO3Error = lastError;
                st.lastSO3Count = lastCount;
                memcpy(resultR, lastResultR, sizeof(resultR));
                stop = true;
            }
            if(!stop)
            {
                lastError = st.lastSO3Error;
                lastCount = st.lastSO3Count;
                memcpy(lastResultR, resultR, sizeof(resultR));

                float delta[3];
                smath::ldlt_solve<float, 3>(jtj, jtr, delta, FLT_EPSILON);
                const double dd[3] = {delta[0], delta[1], delta[2]};
                double rotUpdate[9];
                smath::rodrigues(dd, rotUpdate);
                float ru[9];
                for(int k = 0; k < 9; k++) ru[k] = (float)rotUpdate[k];
                smath::mat3_mul(ru, R_lr, R_lr);
                for(int k = 0; k < 9; k++) resultR[k] = R_lr[k];
                for(int k = 0; k < 3; k++) rec.x[k] = delta[k];
            }
            if(tr)
            {
                for(int k = 0; k < 9; k++) rec.Rcurr[k] = (float)resultR[k];
                tr->push_back(rec);
            }
            if(stop) break;
        }
    }

    int iterations[SLAM_MAX_LEVELS];
    default_iterations(h, pyramid, fastOdom, iterations);

    float Rprev_inv[9];
    smath::mat3_inverse(Rprev, Rprev_inv);

    double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if(so3)
        for(int x = 0; x < 3; x++)
            for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

    for(int i = h->levels - 1; i >= 0; i--)
    {
        const LevelGeom & g = h->geom[i];
        double K[9], Kinv[9];
        k_matrix(g, K);
        smath::mat3_inverse(K, Kinv);

        st.lastRGBError = std::numeric_limits<float>::max();

        for(int j = 0; j < iterations[i]; j++)
        {
            double Rt[16];
            smath::mat4_affine_inverse(resultRt, Rt);
            double R[9], KR[9], KRK_inv[9];
            for(int x = 0; x < 3; x++)
                for(int y = 0; y < 3; y++) R[x * 3 + y] = Rt[x * 4 + y];
            smath::mat3_mul(K, R, KR);
            smath::mat3_mul(KR, Kinv, KRK_inv);
            float krk[9];
            for(int k = 0; k < 9; k++) krk[k] = (float)KRK_inv[k];
            const double tv[3] = {Rt[3], Rt[7], Rt[11]};
            float kt[3];
            for(int x = 0; x < 3; x++) kt[x] = (float)(K[x * 3 + 0] * tv[0] + K[x * 3 + 1] * tv[1] + K[x * 3 + 2] * tv[2]);

            int sigma = 0;
            int rgbSize = 0;

            slam_step_record rec = {};
            rec.kind = 1;
            rec.level = i;
            rec.iteration = j;
            memcpy(rec.Rcurr_in, Rcurr, 36);
            memcpy(rec.tcurr_in, tcurr, 12);
            memcpy(rec.krkinv_in, krk, 36);
            memcpy(rec.so3_in, Rprev_inv, 36);
            memcpy(rec.kt_in, kt, 12);

            if(rgb)
            {
                ResidualArgs a;
                a.minScale = (float)(pow((double)h->minGrad[i], 2.0) / pow((double)h->sobelScale, 2.0));
                a.dIdx = s.dIdx[i];
                a.dIdy = s.dIdy[i];
                a.lastDepth = s.lastDepth[i];
                a.nextDepth = s.nextDepth[i];
                a.lastImage = s.lastImage[i];
                a.nextImage = s.nextImage[i];
                a.maxDepthDelta = h->maxDepthDeltaRGB;
                a.kt = make_float3(kt[0], kt[1], kt[2]);
                a.krkinv = mat3_from(krk);
                a.cols = g.cols;
                a.rows = g.rows;
                int rc = launch_rgb_residual(a, s.corres[i], s.workspace, d_res, h->stream);
                if(rc) return rc;
                h->launches++;
                SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_sums, d_res, 8, cudaMemcpyDeviceToHost, h->stream));
                SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
                rgbSize = reinterpret_cast<int *>(h->h_sums)[0];
                sigma = reinterpret_cast<int *>(h->h_sums)[1];
            }
            rec.rgb_count = rgbSize;
            rec.rgb_sigma = sigma;

            float sigmaVal = std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize);   // sic, RGBDOdometryef.cpp:457
            const float rgbError = std::sqrt(sigma) / (rgbSize == 0 ? 1 : rgbSize);

            if(rgbOnly && rgbError > st.lastRGBError) break;

            st.lastRGBError = rgbError;
            st.lastRGBCount = rgbSize;

            if(rgbOnly) sigmaVal = -1;
            rec.sigma_in = sigmaVal;

            float A_icp[36] = {0}, b_icp[6] = {0}, A_rgbd[36] = {0}, b_rgbd[6] = {0};

            if(icp)
            {
                IcpArgs a;
                a.Rcurr = mat3_from(Rcurr);
                a.tcurr = make_float3(tcurr[0], tcurr[1], tcurr[2]);
                a.Rprev_inv = mat3_from(Rprev_inv);
                a.tprev = make_float3(tprev[0], tprev[1], tprev[2]);
                a.fx = g.fx; a.fy = g.fy; a.cx = g.cx; a.cy = g.cy;
                a.distThres = h->p.dist_thresh;
                a.angleThres = h->p.angle_thresh;
                a.cols = g.cols;
                a.rows = g.rows;
                a.vcurr = s.vcurr[i]; a.ncurr = s.ncurr[i]; a.vprev = s.vprev[i]; a.nprev = s.nprev[i];
                int rc = launch_icp_step(a, s.workspace, d_icp, h->stream);
                if(rc) return rc;
                h->launches++;
            }
            if(rgb)
            {
                RgbStepArgs a;
                a.sigma = sigmaVal;
                a.fx = g.fx; a.fy = g.fy;
                a.sobelScale = h->sobelScale;
                a.cols = g.cols; a.rows = g.rows;
                a.dIdx = s.dIdx[i]; a.dIdy = s.dIdy[i];
                a.lastDepth = s.lastDepth[i];
                a.invFx = 1.0f / g.fx; a.invFy = 1.0f / g.fy; a.cx = g.cx; a.cy = g.cy;
                a.cloud = nullptr;
                int rc = launch_rgb_step(a, s.corres[i], s.workspace, d_rgb, h->stream);
                if(rc) return rc;
                h->launches++;
            }
            SLAM_CUDA_TRY(cudaMemcpyAsync(h->h_sums, s.sums, 64 * 4, cudaMemcpyDeviceToHost, h->stream));
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
            st.gn_iterations++;

            if(icp)
            {
                unpack_se3(h->h_sums, A_icp, b_icp);
                st.lastICPError = sqrtf(h->h_sums[27]) / h->h_sums[28];
                st.lastICPCount = h->h_sums[28];
                memcpy(rec.icp, h->h_sums, 29 * 4);
            }
            if(rgb)
            {
                unpack_se3(h->h_sums + 32, A_rgbd, b_rgbd);
                memcpy(rec.rgb, h->h_sums + 32, 29 * 4);
            }

            double result[6];
            if(icp && rgb)
            {
                const double w = icpWeight;
                for(int k = 0; k < 36; k++) st.lastA[k] = (double)A_rgbd[k] + w * w * (double)A_icp[k];
                for(int k = 0; k < 6; k++) st.lastb[k] = (double)b_rgbd[k] + w * (double)b_icp[k];
            }
            else if(icp)
            {
                for(int k = 0; k < 36; k++) st.lastA[k] = A_icp[k];
                for(int k = 0; k < 6; k++) st.lastb[k] = b_icp[k];
            }
            else
            {
                for(int k = 0; k < 36; k++) st.lastA[k] = A_rgbd[k];
                for(int k = 0; k < 6; k++) st.lastb[k] = b_rgbd[k];
            }
            smath::spd_solve6(st.lastA, st.lastb, result);

            smath::update_se3(resultRt, result);
            smath::compose_current_pose(Rprev, tprev, resultRt, Rcurr, tcurr);

            if(tr)
            {
                for(int k = 0; k < 6; k++) rec.x[k] = result[k];
                memcpy(rec.Rcurr, Rcurr, sizeof(Rcurr));
                memcpy(rec.tcurr, tcurr, sizeof(tcurr));
                tr->push_back(rec);
            }
        }
    }

    if(rgb)
    {
        const float dx = tcurr[0] - tprev[0], dy = tcurr[1] - tprev[1], dz = tcurr[2] - tprev[2];
        if(sqrtf(dx * dx + dy * dy + dz * dz) > 0.3)   // RGBDOdometryef.cpp:579-583
        {
            memcpy(Rcurr, Rprev, sizeof(Rprev));
            memcpy(tcurr, tprev, sizeof(tprev));
        }
    }
    if(so3)
        for(int l = 0; l < h->levels; l++) std::swap(s.lastNextImage[l], s.nextImage[l]);

    memcpy(trans, tcurr, sizeof(tcurr));
    memcpy(rot, Rcurr, sizeof(Rcurr));
    return SLAM_OK;
}

}   // namespace

static int finish_device_loop(slam_odom_t h, float * trans, float * rot);
static int resolve_stats(slam_odom_t h);
// The bookkeeping of an asynchronous track (the lastNextImage <-> nextImage swap of RGBDOdometryef.cpp:585-591, the statistics)
// happens when it is collected.  Every entry point that writes one of those buffers collects a pending track first, so a
// pipelined caller may prepare frame N + 1 before slam_odom_wait() without landing its images in the pre-swap buffers.
static int flush_pending(slam_odom_t h)
{
    if(h->pending_async) return finish_device_loop(h, nullptr, nullptr);
    return SLAM_OK;
}

// ------------------------------------------------------------------ C ABI
extern "C" const char * slam_odom_version(void) { return "slam_b200 0.1 (sm_100a)"; }
extern "C" const char * slam_odom_last_error(void) { return g_last_error.c_str(); }

extern "C" int slam_odom_destroy(slam_odom_t h);

static int create_body(slam_odom * h, const slam_odom_params * params)
{
    h->p = *params;
    if(h->p.dist_thresh == 0) h->p.dist_thresh = 0.10f;
    if(h->p.angle_thresh == 0) h->p.angle_thresh = sinf(20.f * 3.14159254f / 180.f);
    h->levels = params->num_levels ? params->num_levels : 3;
    h->batch = params->batch > 1 ? params->batch : 1;
    for(int l = 0; l < h->levels; l++)
    {
        const int div = 1 << l;
        h->geom[l].rows = params->height >> l;
        h->geom[l].cols = params->width >> l;
        h->geom[l].fx = params->fx / div;
        h->geom[l].fy = params->fy / div;
        h->geom[l].cx = params->cx / div;
        h->geom[l].cy = params->cy / div;
        if(h->geom[l].rows < 2 || h->geom[l].cols < 2)
        {
            set_last_error("image too small for the requested pyramid");
            return SLAM_ERR_ARG;
        }
    }
    if(params->stream)
        h->stream = (cudaStream_t)params->stream;
    else
    {
        SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    SLAM_CUDA_TRY(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    SLAM_CUDA_TRY(cudaEventCreateWithFlags(&h->compute_done, cudaEventDisableTiming));
    SLAM_CUDA_TRY(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    SLAM_CUDA_TRY(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    SLAM_CUDA_TRY(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, params->device));
    if(h->num_sms > kGnMaxCtas) h->num_sms = kGnMaxCtas;

    if(const char * e = getenv("SLAM_BATCH_ENGINE_MIN")) h->be_min = std::max(2, atoi(e));
    ArenaPlan plan;
    for(int b = 0; b < h->batch; b++) layout_sequence(h, plan, nullptr);
    const size_t pose_off = plan.take(sizeof(float) * 12 * h->batch);
    const size_t gn_off = plan.take(gn_state_bytes(h->batch, h->num_sms));
    const size_t be_off = h->batch >= h->be_min ? plan.take(batch_state_bytes(h->batch, h->geom, h->levels)) : 0;
    h->arena_bytes = plan.off;
    SLAM_CUDA_TRY(cudaMalloc((void **)&h->arena, h->arena_bytes));
    SLAM_CUDA_TRY(cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    h->seq.resize(h->batch);
    ArenaPlan place;
    for(int b = 0; b < h->batch; b++) layout_sequence(h, place, &h->seq[b]);
    h->seq_stride = h->batch > 1 ? (size_t)((char *)h->seq[1].depth[0] - (char *)h->seq[0].depth[0]) : 0;
    h->d_poses12 = (float *)(h->arena + pose_off);
    SLAM_CUDA_TRY(cudaMallocHost((void **)&h->h_poses12, sizeof(float) * 12 * h->batch));
    gn_bind_state(h->gn, h->arena + gn_off, h->batch, h->num_sms);
    if(h->batch >= h->be_min)
    {
        batch_bind_state(h->be, h->arena + be_off, h->batch, h->geom, h->levels, h->gn.seq_in, h->gn.results);
        h->be.num_sms = h->num_sms;
    }

    SLAM_CUDA_TRY(cudaHostAlloc((void **)&h->h_results, sizeof(GnResult) * h->batch, cudaHostAllocMapped));
    SLAM_CUDA_TRY(cudaHostAlloc((void **)&h->h_flags, sizeof(unsigned) * 2 * h->batch, cudaHostAllocMapped));
    memset(h->h_flags, 0, sizeof(unsigned) * 2 * h->batch);
    {
        // zero-copy completion needs the device view of the pinned block to be the host pointer (unified addressing)
        void * dv = nullptr;
        const char * off = getenv("SLAM_ODOM_ZERO_COPY");
        h->zero_copy = !(off && off[0] == '0') && cudaHostGetDevicePointer(&dv, h->h_results, 0) == cudaSuccess && dv == (void *)h->h_results &&
                       cudaHostGetDevicePointer(&dv, h->h_flags, 0) == cudaSuccess && dv == (void *)h->h_flags;
        cudaGetLastError();
        if(getenv("SLAM_ODOM_DEBUG")) fprintf(stderr, "slam_odom_create: zero-copy completion %s\n", h->zero_copy ? "on" : "off");
    }
    SLAM_CUDA_TRY(cudaMallocHost((void **)&h->h_sums, 128 * 4));

    h->stats.resize(h->batch);
    h->last_pose.assign((size_t)12 * h->batch, 0.f);
    h->trace.resize(h->batch);
    for(auto & st : h->stats)
    {
        memset(&st, 0, sizeof(st));
        // RGBDOdometryef.cpp:26-31
        st.lastICPCount = st.lastRGBCount = st.lastSO3Count = (float)(params->width * params->height);
    }
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SLAM_OK;
}

extern "C" int slam_odom_create(const slam_odom_params * params, slam_odom_t * out)
{
    SLAM_ARG_CHECK(params && out);
    SLAM_ARG_CHECK(params->width > 0 && params->height > 0);
    SLAM_ARG_CHECK(params->num_levels >= 0 && params->num_levels <= SLAM_MAX_LEVELS);
    int ndev = 0;
    SLAM_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if(ndev == 0)
    {
        set_last_error("no CUDA device: libslam_odom has no CPU fallback");
        return SLAM_ERR_CUDA;
    }
    SLAM_ARG_CHECK(params->device >= 0 && params->device < ndev);
    SLAM_CUDA_TRY(cudaSetDevice(params->device));

    slam_odom * h = new slam_odom();
    // any failure past this point releases whatever was created so far (streams, events, arena, pinned blocks)
    const int rc = create_body(h, params);
    if(rc != SLAM_OK)
    {
        const std::string keep = g_last_error;
        slam_odom_destroy(h);
        g_last_error = keep;
        return rc;
    }
    *out = h;
    return SLAM_OK;
}

extern "C" int slam_odom_destroy(slam_odom_t h)
{
    if(!h) return SLAM_OK;
    cudaSetDevice(h->p.device);
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->copy_stream);
    if(h->aux_stream) cudaStreamSynchronize(h->aux_stream);
    for(auto & sl : h->slot)
    {
        if(sl.depth) cudaFree(sl.depth);
        if(sl.ready) cudaEventDestroy(sl.ready);
    }
    if(h->batch >= h->be_min)
    {
        batch_report();
        for(auto st : h->be.side) cudaStreamSynchronize(st);
        for(auto st : h->be.role) cudaStreamSynchronize(st);
        batch_release(h->be);
    }
    gn_release(h->gn);
    if(h->arena) cudaFree(h->arena);
    if(h->h_results) cudaFreeHost(h->h_results);
    if(h->h_flags) cudaFreeHost(h->h_flags);
    if(h->h_sums) cudaFreeHost(h->h_sums);
    if(h->h_poses12) cudaFreeHost(h->h_poses12);
    if(h->score_ws) cudaFree(h->score_ws);
    if(h->h_score_poses) cudaFreeHost(h->h_score_poses);
    for(int r = 0; r < h->peer_world; r++)
        if(r != h->peer_rank && h->peer_ptr[r]) cudaIpcCloseMemHandle(h->peer_ptr[r]);
    if(h->peer_local) cudaFree(h->peer_local);
    if(h->d_peer_key) cudaFree(h->d_peer_key);
    if(h->h_peer_result) cudaFreeHost(h->h_peer_result);
    if(h->filtered_depth) cudaFree(h->filtered_depth);
    if(h->compute_done) cudaEventDestroy(h->compute_done);
    if(h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if(h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if(h->fork_ev) cudaEventDestroy(h->fork_ev);
    if(h->join_ev) cudaEventDestroy(h->join_ev);
    if(h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return SLAM_OK;
}

extern "C" int slam_odom_init_icp_depth(slam_odom_t h, const uint16_t * d_depth, size_t pitch_bytes, float depth_cutoff)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(d_depth);
    if(int rc = set_device(h)) return rc;
    if(int rc = flush_pending(h)) return rc;
    const size_t row = pitch_bytes ? pitch_bytes : (size_t)h->geom[0].cols * 2;
    for(int b = 0; b < h->batch; b++)
        if(int rc = enqueue_init_icp_depth(h, b, (const uint16_t *)((const char *)d_depth + (size_t)b * row * h->geom[0].rows), pitch_bytes, depth_cutoff))
            return rc;
    return SLAM_OK;
}

extern "C" int slam_odom_init_icp_maps(slam_odom_t h, const float * d_vertices4, const float * d_normals4, float /*depth_cutoff*/)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(d_vertices4 && d_normals4);
    if(int rc = set_device(h)) return rc;
    if(int rc = flush_pending(h)) return rc;
    const size_t n4 = (size_t)h->geom[0].rows * h->geom[0].cols * 4;
    for(int b = 0; b < h->batch; b++)
        if(int rc = enqueue_model_maps(h, b, d_vertices4 + b * n4, d_normals4 + b * n4, false, nullptr)) return rc;
    h->have_depth_tmp = true;
    return SLAM_OK;
}

extern "C" int slam_odom_init_icp_model(slam_odom_t h, const float * d_vertices4, const float * d_normals4, float /*depth_cutoff*/,
                                        const float * model_pose16)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(d_vertices4 && d_normals4 && model_pose16);
    if(int rc = set_device(h)) return rc;
    if(int rc = flush_pending(h)) return rc;
    const size_t n4 = (size_t)h->geom[0].rows * h->geom[0].cols * 4;
    for(int b = 0; b < h->batch; b++)
        if(int rc = enqueue_model_maps(h, b, d_vertices4 + b * n4, d_normals4 + b * n4, true, model_pose16 + 16 * b)) return rc;
    h->have_depth_tmp = true;
    return SLAM_OK;
}

static int init_rgb_common(slam_odom_t h, const uint8_t * d_rgba, int which)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(d_rgba);
    if(int rc = set_device(h)) return rc;
    if(int rc = flush_pending(h)) return rc;
    if(which != 2 && !h->have_depth_tmp)
    {
        // RGBDOdometryef.cpp:239,245: populateRGBDData reads vmaps_tmp written by initICPModel / initICP(maps)
        set_last_error("initRGB/initRGBModel called before initICPModel/initICP(maps)");
        return SLAM_ERR_ORDER;
    }
    const size_t n4 = (size_t)h->geom[0].rows * h->geom[0].cols * 4;
    for(int b = 0; b < h->batch; b++)
    {
        SeqBuffers & s = h->seq[b];
        int rc;
        if(which == 0)
        {
            h->deriv_stale = true;
            h->deriv_src_swapped = false;
            rc = enqueue_populate_rgbd(h, b, d_rgba + b * n4, s.nextDepth, s.nextImage);
        }
        else if(which == 1)
            rc = enqueue_populate_rgbd(h, b, d_rgba + b * n4, s.lastDepth, s.lastImage);
        else
            rc = enqueue_populate_rgbd(h, b, d_rgba + b * n4, nullptr, s.lastNextImage);
        if(rc) return rc;
    }
    return SLAM_OK;
}

extern "C" int slam_odom_init_rgb(slam_odom_t h, const uint8_t * d_rgba) { return init_rgb_common(h, d_rgba, 0); }
extern "C" int slam_odom_init_rgb_model(slam_odom_t h, const uint8_t * d_rgba) { return init_rgb_common(h, d_rgba, 1); }
extern "C" int slam_odom_init_first_rgb(slam_odom_t h, const uint8_t * d_rgba) { return init_rgb_common(h, d_rgba, 2); }

// Enqueue the device-resident loop for the whole batch.
static int enqueue_device_loop(slam_odom_t h, const float * trans, const float * rot, int rgb_only, float icp_weight, int pyramid, int fast_odom, int so3)
{
    const bool icp = !rgb_only && icp_weight > 0;
    const bool rgb = rgb_only || icp_weight < 100;
    if(so3 && h->levels < 3)
    {
        set_last_error("so3 pre-alignment needs pyramid level 2 (num_levels >= 3)");
        return SLAM_ERR_UNSUPPORTED;
    }
    // A few sequences with the SO3 step: ONE split launch pair works through them one after the other (gn_kernel.cu) -- at 8
    // sequences per GPU that beats both the streaming engine (too few sequences to fill its ~140 launches) and concurrent CTA groups.
    gn_configure(h->gn);
    const bool seq_split = h->batch >= 3 && h->batch <= h->gn.seq_max && h->batch <= kSplitMaxSeqs && so3 && icp && !rgb_only && !h->trace_on &&
                           h->gn.split == 1 && !h->gn.split_broken;
    const bool streaming = h->batch >= h->be_min && h->trace_level < 2 && !seq_split;   // many sequences: lock-step streaming launches
    GnLaunch L = {};
    L.levels = h->levels;
    L.batch = h->batch;
    for(int l = 0; l < h->levels; l++) L.geom[l] = h->geom[l];
    default_iterations(h, pyramid, fast_odom, L.iterations);
    L.icp = icp;
    L.rgb = rgb;
    L.rgb_only = rgb_only != 0;
    L.so3 = so3 != 0;
    // shared-memory plan of the persistent kernel: which levels keep their operands resident
    L.trace = h->trace_on;
    L.full_corres = h->trace_level >= 2;
    const bool all_resident = streaming ? false : (gn_make_plan(h->gn, L) || gn_split_applies(h->gn, L));
    // Derivative images are only materialised when something reads them from memory: a test tap / trace, or a level that
    // streams its operands.  Otherwise the persistent kernel derives the two gradients of its own pixels from nextImage
    // while staging them (same arithmetic, bit-identical).
    const bool derive = rgb && !h->trace_on && !streaming && all_resident;
    h->be.cand_ready = false;
    h->deriv_stale = !(rgb && !derive);
    if(rgb && !derive)
    {
        int rows[SLAM_MAX_LEVELS], cols[SLAM_MAX_LEVELS];
        for(int l = 0; l < h->levels; l++)
        {
            rows[l] = h->geom[l].rows;
            cols[l] = h->geom[l].cols;
        }
        SeqBuffers & s0 = h->seq[0];
        bool quads = true;
        for(int l = 0; l < h->levels; l++) quads = quads && cols[l] % 4 == 0;
        if(streaming && quads)
        {
            // the batched engine also wants the pose-independent half of the RGB association: same pass over nextImage
            unsigned char * cand[SLAM_MAX_LEVELS];
            float min_scale[SLAM_MAX_LEVELS];
            for(int l = 0; l < h->levels; l++)
            {
                cand[l] = h->be.cand0 + h->be.cand_off[l];
                min_scale[l] = (float)(pow((double)h->minGrad[l], 2.0) / pow((double)h->sobelScale, 2.0));
            }
            if(int rc = launch_deriv_cand(h->levels, s0.nextImage, s0.nextDepth, s0.dIdx, s0.dIdy, cand, min_scale, rows, cols, h->stream, h->batch, h->seq_stride,
                                          h->be.aux_stride))
                return rc;
            h->be.cand_ready = true;
        }
        else if(int rc = launch_derivatives_simple(h->levels, s0.nextImage, s0.dIdx, s0.dIdy, rows, cols, h->stream, h->batch, h->seq_stride))
            return rc;
        h->launches++;
    }

    L.derive_gradients = derive;
    L.icp_weight = icp_weight;
    L.dist_thresh = h->p.dist_thresh;
    L.angle_thresh = h->p.angle_thresh;
    L.sobel_scale = h->sobelScale;
    L.max_depth_delta = h->maxDepthDeltaRGB;
    for(int l = 0; l < h->levels; l++) L.min_scale[l] = (float)(pow((double)h->minGrad[l], 2.0) / pow((double)h->sobelScale, 2.0));
    L.trace = h->trace_on;
    L.full_corres = h->trace_level >= 2;
    int rc;
    if(streaming)
    {
        if(int rc2 = resolve_stats(h)) return rc2;   // (statistics of an earlier zero-copy launch: the D2H copy below replaces h_results)
        GnSeqIn * in = nullptr;
        rc = gn_stage_inputs(h->gn, L, h->seq.data(), trans, rot, &in);
        if(rc) return rc;
        const long long before = h->be.launches;
        if(h->gn.profiling && h->gn.ev.size() >= 4096)
            if(int rc2 = gn_fold_profile(h->gn)) return rc2;
        rc = batch_enqueue(h->be, L, in, h->h_results, h->gn.trace, h->gn.trace_count, h->stream, h->gn.profiling ? &h->gn.ev : nullptr);
        h->launches += h->be.launches - before;
        h->gn.so3_swapped = L.so3;
    }
    else
    {
        const bool zc = h->zero_copy && !h->trace_on;
        // zero-copy launches leave their statistics in the device-side result block, which merges from launch to launch (only the terms
        // a call computes are written): nothing to collect per frame.  A launch whose results follow by a D2H copy replaces h_results.
        if(!zc)
            if(int rc2 = resolve_stats(h)) return rc2;
        if(zc)
        {
            h->zc_seqno = (h->zc_seqno + 1) & 0x7fffffffu;   // bit 31 of the completion flag marks a failed launch
            if(!h->zc_seqno) h->zc_seqno = 1;
        }
        rc = gn_enqueue(h->gn, L, h->seq.data(), trans, rot, h->h_results, h->stream, zc ? h->h_flags : nullptr, h->zc_seqno);
        h->zc_pending = zc && rc == SLAM_OK;
        h->launches += h->gn.last_launches;
    }
    if(rc) return rc;
    h->pending_async = true;
    h->last_icp = icp;
    h->last_rgb = rgb;
    h->last_so3 = so3 != 0;
    h->pend_icp = h->pend_icp || icp;
    h->pend_rgb = h->pend_rgb || rgb;
    h->pend_so3 = h->pend_so3 || so3 != 0;
    return SLAM_OK;
}

// Wait for the completion flags the persistent kernel writes into mapped host memory after its last solve.  The stream is polled
// now and then so that a failed launch surfaces as an error instead of a hang.
static int wait_zero_copy(slam_odom_t h, int which = 0)
{
    volatile unsigned * flags = h->h_flags + (size_t)which * h->batch;
    for(unsigned spins = 1;; spins++)
    {
        bool all = true;
        for(int b = 0; b < h->batch; b++) all = all && flags[b] == h->zc_seqno;
        if(all) break;
        if((spins & 0x3fff) == 0)
        {
            const cudaError_t q = cudaStreamQuery(h->stream);
            if(q == cudaSuccess)
            {
                for(int b = 0; b < h->batch; b++)
                    if(flags[b] != h->zc_seqno)
                    {
                        set_last_error(flags[b] == (h->zc_seqno | 0x80000000u)
                                           ? "persistent kernel: a wait between thread blocks timed out; the results of this track are invalid"
                                           : "persistent kernel finished without publishing its results");
                        return SLAM_ERR_CUDA;
                    }
                break;
            }
            if(q != cudaErrorNotReady)
            {
                set_last_error(std::string("persistent kernel: ") + cudaGetErrorString(q));
                return SLAM_ERR_CUDA;
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return SLAM_OK;
}

// The statistics of a zero-copy track reach the host a few microseconds after its pose (second flag): taken when somebody
// asks for them, or before the next launch reuses the result block.
static int copy_stats(slam_odom_t h)
{
    for(int b = 0; b < h->batch; b++)
    {
        const GnResult & r = h->h_results[b];
        if(r.gn_iterations < 0)
        {
            set_last_error("persistent kernel: a wait between thread blocks timed out; the results of this track are invalid");
            return SLAM_ERR_CUDA;
        }
        slam_odom_stats & st = h->stats[b];
        if(h->pend_icp)
        {
            st.lastICPError = r.lastICPError;
            st.lastICPCount = r.lastICPCount;
        }
        if(h->pend_rgb)
        {
            st.lastRGBError = r.lastRGBError;
            st.lastRGBCount = r.lastRGBCount;
        }
        if(h->pend_so3)
        {
            st.lastSO3Error = r.lastSO3Error;
            st.lastSO3Count = r.lastSO3Count;
        }
        memcpy(st.lastA, r.lastA, sizeof(st.lastA));
        memcpy(st.lastb, r.lastb, sizeof(st.lastb));
        st.so3_iterations = r.so3_iterations;
        st.gn_iterations = r.gn_iterations;
    }
    h->pend_icp = h->pend_rgb = h->pend_so3 = false;
    return SLAM_OK;
}

static int resolve_stats(slam_odom_t h)
{
    if(!h->stats_lazy) return SLAM_OK;
    h->stats_lazy = false;
    // the persistent kernel keeps the result blocks in device memory: wait for its last instructions, fetch them
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    SLAM_CUDA_TRY(cudaMemcpy(h->h_results, h->gn.results, sizeof(GnResult) * h->batch, cudaMemcpyDeviceToHost));
    return copy_stats(h);
}

static int finish_device_loop(slam_odom_t h, float * trans, float * rot)
{
    bool lazy = false;
    if(h->pending_async && h->zc_pending)
    {
        h->zc_pending = false;
        if(int rc = wait_zero_copy(h, 0)) return rc;   // the pose is out; the statistics follow (resolve_stats)
        lazy = true;
    }
    else
        SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if(!h->pending_async) return SLAM_OK;
    h->pending_async = false;
    if(lazy)
        h->stats_lazy = true;
    else if(int rc = copy_stats(h))
        return rc;
    for(int b = 0; b < h->batch; b++)
    {
        const GnResult & r = h->h_results[b];
        if(trans) memcpy(trans + 3 * b, r.tcurr, 12);
        if(rot) memcpy(rot + 9 * b, r.Rcurr, 36);
        memcpy(&h->last_pose[12 * b], r.Rcurr, 36);
        memcpy(&h->last_pose[12 * b + 9], r.tcurr, 12);
        if(h->gn.so3_swapped)
        {
            for(int l = 0; l < h->levels; l++) std::swap(h->seq[b].lastNextImage[l], h->seq[b].nextImage[l]);
            h->deriv_src_swapped = true;
        }
    }
    h->gn.so3_swapped = false;
    return SLAM_OK;
}

extern "C" int slam_odom_get_incremental_transformation_async(slam_odom_t h, const float * trans, const float * rot, int rgb_only, float icp_weight,
                                                              int pyramid, int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(trans && rot);
    if(int rc = set_device(h)) return rc;
    if(h->p.host_loop)
    {
        set_last_error("async form needs the device-resident loop (host_loop = 0)");
        return SLAM_ERR_UNSUPPORTED;
    }
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    return enqueue_device_loop(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
}

extern "C" int slam_odom_wait(slam_odom_t h, float * trans, float * rot)
{
    if(int rc = check_handle(h)) return rc;
    if(int rc = set_device(h)) return rc;
    if(!h->pending_async)
    {
        // already collected (an init_* call for the next frame came first): hand out the pose that was kept
        for(int b = 0; b < h->batch; b++)
        {
            if(rot) memcpy(rot + 9 * b, &h->last_pose[12 * b], 36);
            if(trans) memcpy(trans + 3 * b, &h->last_pose[12 * b + 9], 12);
        }
        return SLAM_OK;
    }
    return finish_device_loop(h, trans, rot);
}

extern "C" int slam_odom_get_incremental_transformation(slam_odom_t h, float * trans, float * rot, int rgb_only, float icp_weight, int pyramid,
                                                        int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(trans && rot);
    if(int rc = set_device(h)) return rc;
    if(h->p.host_loop)
    {
        for(int b = 0; b < h->batch; b++)
            if(int rc = host_loop_one(h, b, trans + 3 * b, rot + 9 * b, rgb_only != 0, icp_weight, pyramid != 0, fast_odom != 0, so3 != 0)) return rc;
        return SLAM_OK;
    }
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    if(int rc = enqueue_device_loop(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3)) return rc;
    return finish_device_loop(h, trans, rot);
}

extern "C" int slam_odom_get_covariance(slam_odom_t h, double * out36)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(out36);
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    if(int rc = resolve_stats(h)) return rc;
    for(int b = 0; b < h->batch; b++) smath::lu_inverse<double, 6>(h->stats[b].lastA, out36 + 36 * b);   // RGBDOdometryef.cpp:597-600
    return SLAM_OK;
}

extern "C" int slam_odom_get_stats(slam_odom_t h, slam_odom_stats * stats)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(stats);
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    if(int rc = resolve_stats(h)) return rc;
    for(int b = 0; b < h->batch; b++) stats[b] = h->stats[b];
    return SLAM_OK;
}

extern "C" int slam_odom_init_icp_depth_raw(slam_odom_t h, const uint16_t * d_raw_depth, float filter_max_depth_m, float depth_cutoff)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(d_raw_depth);
    if(int rc = set_device(h)) return rc;
    const size_t n0 = (size_t)h->geom[0].rows * h->geom[0].cols;
    if(!h->filtered_depth) SLAM_CUDA_TRY(cudaMalloc((void **)&h->filtered_depth, n0 * 2 * h->batch));
    if(int rc = launch_depth_bilateral(d_raw_depth, h->geom[0].rows, h->geom[0].cols, filter_max_depth_m, h->filtered_depth, h->batch, h->stream)) return rc;
    h->launches++;
    return slam_odom_init_icp_depth(h, h->filtered_depth, 0, depth_cutoff);
}

// Common part of the two scoring entry points: upload the poses, launch.  best_key (device pointer, may be null): see k_score_poses.
static int enqueue_score_poses(slam_odom_t h, int seq, int level, int n, const float * prev_trans3, const float * prev_rot9, const float * trans3n,
                               const float * rot9n, float ** out2_dev, unsigned long long * d_best_key, int index_base, float min_inliers)
{
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    const LevelGeom & g = h->geom[level];
    const int plane = g.rows * g.cols;
    const size_t pose_bytes = align_up((size_t)n * 12 * 4, 256);
    const size_t need = pose_bytes + score_workspace_bytes(n, plane);
    if(need > h->score_ws_bytes)
    {
        SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
        if(h->score_ws) cudaFree(h->score_ws);
        if(h->h_score_poses) cudaFreeHost(h->h_score_poses);
        h->score_ws = nullptr;
        h->h_score_poses = nullptr;
        h->score_ws_bytes = 0;
        SLAM_CUDA_TRY(cudaMalloc((void **)&h->score_ws, need));
        SLAM_CUDA_TRY(cudaMallocHost((void **)&h->h_score_poses, pose_bytes));
        h->score_ws_bytes = need;
        h->score_ws_n = -1;
    }
    // the layout of the workspace depends on n (the tickets must start at zero; the kernel leaves them at zero)
    if(h->score_ws_n != n || h->score_ws_plane != plane)
    {
        SLAM_CUDA_TRY(cudaMemsetAsync(h->score_ws, 0, need, h->stream));
        h->score_ws_n = n;
        h->score_ws_plane = plane;
    }
    // the pinned staging block is reused: the previous upload must have been consumed
    SLAM_CUDA_TRY(cudaEventSynchronize(h->compute_done));
    float * poses = h->h_score_poses;
    for(int i = 0; i < n; i++)
    {
        memcpy(&poses[(size_t)i * 12], rot9n + (size_t)i * 9, 9 * sizeof(float));
        memcpy(&poses[(size_t)i * 12 + 9], trans3n + (size_t)i * 3, 3 * sizeof(float));
    }
    SLAM_CUDA_TRY(cudaMemcpyAsync(h->score_ws, poses, (size_t)n * 12 * 4, cudaMemcpyHostToDevice, h->stream));
    SLAM_CUDA_TRY(cudaEventRecord(h->compute_done, h->stream));
    SeqBuffers & s = h->seq[seq];
    float Rprev_inv[9];
    smath::mat3_inverse(prev_rot9, Rprev_inv);
    IcpArgs a;
    a.Rcurr = mat3_from(prev_rot9);
    a.tcurr = make_float3(prev_trans3[0], prev_trans3[1], prev_trans3[2]);
    a.Rprev_inv = mat3_from(Rprev_inv);
    a.tprev = make_float3(prev_trans3[0], prev_trans3[1], prev_trans3[2]);
    a.fx = g.fx; a.fy = g.fy; a.cx = g.cx; a.cy = g.cy;
    a.distThres = h->p.dist_thresh;
    a.angleThres = h->p.angle_thresh;
    a.cols = g.cols; a.rows = g.rows;
    a.vcurr = s.vcurr[level]; a.ncurr = s.ncurr[level]; a.vprev = s.vprev[level]; a.nprev = s.nprev[level];
    if(int rc = launch_score_poses(a, reinterpret_cast<const float *>(h->score_ws), n, h->score_ws + pose_bytes, out2_dev, h->stream, d_best_key, index_base, min_inliers))
        return rc;
    h->launches++;
    return SLAM_OK;
}

extern "C" int slam_odom_score_poses(slam_odom_t h, int seq, int level, int n, const float * prev_trans3, const float * prev_rot9, const float * trans3n,
                                     const float * rot9n, float * residual_n, float * count_n)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(seq >= 0 && seq < h->batch && level >= 0 && level < h->levels && n > 0 && n <= 65535);
    SLAM_ARG_CHECK(prev_trans3 && prev_rot9 && trans3n && rot9n && residual_n && count_n);
    if(int rc = set_device(h)) return rc;
    float * out2 = nullptr;
    if(int rc = enqueue_score_poses(h, seq, level, n, prev_trans3, prev_rot9, trans3n, rot9n, &out2, nullptr, 0, 1.f)) return rc;
    std::vector<float> out((size_t)n * 2);
    SLAM_CUDA_TRY(cudaMemcpyAsync(out.data(), out2, (size_t)n * 2 * 4, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    for(int i = 0; i < n; i++)
    {
        residual_n[i] = out[2 * i];
        count_n[i] = out[2 * i + 1];
    }
    return SLAM_OK;
}

extern "C" int slam_odom_score_poses_best(slam_odom_t h, int seq, int level, int n, int index_base, float min_inliers, const float * prev_trans3,
                                          const float * prev_rot9, const float * trans3n, const float * rot9n, unsigned long long * d_best_key)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(seq >= 0 && seq < h->batch && level >= 0 && level < h->levels && n > 0 && n <= 65535 && index_base >= 0);
    SLAM_ARG_CHECK(prev_trans3 && prev_rot9 && trans3n && rot9n && d_best_key);
    if(int rc = set_device(h)) return rc;
    float * out2 = nullptr;
    return enqueue_score_poses(h, seq, level, n, prev_trans3, prev_rot9, trans3n, rot9n, &out2, d_best_key, index_base, min_inliers);
}

extern "C" int slam_odom_peer_export(slam_odom_t h, void * handle64)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI passes the handle as 64 opaque bytes");
    if(int rc = set_device(h)) return rc;
    if(!h->peer_local)
    {
        SLAM_CUDA_TRY(cudaMalloc((void **)&h->peer_local, sizeof(PeerSlot) * 2 * kMaxPeers));
        SLAM_CUDA_TRY(cudaMemset(h->peer_local, 0, sizeof(PeerSlot) * 2 * kMaxPeers));
        SLAM_CUDA_TRY(cudaMalloc((void **)&h->d_peer_key, 8));
        SLAM_CUDA_TRY(cudaHostAlloc((void **)&h->h_peer_result, 4 * 8, cudaHostAllocMapped));
        memset(h->h_peer_result, 0, 4 * 8);
        h->h_peer_result[3] = 0x7fffffffffffffffull;
        SLAM_CUDA_TRY(cudaHostGetDevicePointer((void **)&h->d_peer_result, h->h_peer_result, 0));
    }
    cudaIpcMemHandle_t hd;
    SLAM_CUDA_TRY(cudaIpcGetMemHandle(&hd, h->peer_local));
    memcpy(handle64, &hd, 64);
    return SLAM_OK;
}

extern "C" int slam_odom_peer_connect(slam_odom_t h, int rank, int world, const void * handles64)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(handles64 && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world);
    if(!h->peer_local)
    {
        set_last_error("slam_odom_peer_connect: call slam_odom_peer_export first");
        return SLAM_ERR_ORDER;
    }
    if(h->peer_world)
    {
        set_last_error("slam_odom_peer_connect: already connected");
        return SLAM_ERR_ORDER;
    }
    if(int rc = set_device(h)) return rc;
    for(int r = 0; r < world; r++)
    {
        if(r == rank)
        {
            h->peer_ptr[r] = h->peer_local;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char *)handles64 + 64 * r, 64);
        void * p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
        if(e != cudaSuccess)
        {
            for(int q = 0; q < r; q++)
                if(q != rank && h->peer_ptr[q]) cudaIpcCloseMemHandle(h->peer_ptr[q]);
            for(int q = 0; q < kMaxPeers; q++) h->peer_ptr[q] = nullptr;
            set_last_error(std::string("slam_odom_peer_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            cudaGetLastError();
            return SLAM_ERR_CUDA;
        }
        h->peer_ptr[r] = (PeerSlot *)p;
    }
    h->peer_rank = rank;
    h->peer_world = world;
    return SLAM_OK;
}

extern "C" int slam_odom_score_poses_best_peers(slam_odom_t h, int seq, int level, int n, int index_base, float min_inliers, const float * prev_trans3,
                                                const float * prev_rot9, const float * trans3n, const float * rot9n, unsigned long long * best_key)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(seq >= 0 && seq < h->batch && level >= 0 && level < h->levels && n >= 0 && n <= 65535 && index_base >= 0 && best_key);
    SLAM_ARG_CHECK(n == 0 || (prev_trans3 && prev_rot9 && trans3n && rot9n));
    if(!h->peer_world)
    {
        set_last_error("slam_odom_score_poses_best_peers: not connected (slam_odom_peer_export / _connect)");
        return SLAM_ERR_ORDER;
    }
    if(int rc = set_device(h)) return rc;
    const unsigned long long fno = ++h->peer_seq;
    // this rank's key starts at INT64_MAX (a rank without hypotheses or without an acceptable one publishes that)
    SLAM_CUDA_TRY(cudaMemcpyAsync(h->d_peer_key, h->h_peer_result + 3, 8, cudaMemcpyHostToDevice, h->stream));
    if(n > 0)
    {
        float * out2 = nullptr;
        if(int rc = enqueue_score_poses(h, seq, level, n, prev_trans3, prev_rot9, trans3n, rot9n, &out2, h->d_peer_key, index_base, min_inliers)) return rc;
    }
    PeerPtrs pp = {};
    for(int r = 0; r < h->peer_world; r++) pp.p[r] = h->peer_ptr[r];
    k_peer_min<<<1, 32, 0, h->stream>>>(pp, h->peer_rank, h->peer_world, h->d_peer_key, fno, h->peer_local, h->d_peer_result);
    SLAM_CUDA_TRY(cudaGetLastError());
    h->launches++;
    volatile unsigned long long * res = h->h_peer_result;
    for(unsigned spins = 1; res[1] != fno; spins++)
    {
        if((spins & 0x3fff) == 0)
        {
            const cudaError_t q = cudaStreamQuery(h->stream);
            if(q == cudaSuccess) break;
            if(q != cudaErrorNotReady)
            {
                set_last_error(std::string("peer minimum: ") + cudaGetErrorString(q));
                return SLAM_ERR_CUDA;
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if(res[1] != fno || res[2] != 0)
    {
        set_last_error("peer minimum: a peer did not publish its key in time");
        return SLAM_ERR_CUDA;
    }
    *best_key = res[0];
    return SLAM_OK;
}

extern "C" long long slam_odom_launch_count(slam_odom_t h) { return h ? h->launches : 0; }

extern "C" int slam_odom_set_profiling(slam_odom_t h, int enable)
{
    if(int rc = check_handle(h)) return rc;
    h->gn.profiling = enable != 0;
    return SLAM_OK;
}

extern "C" int slam_odom_get_profile(slam_odom_t h, double * gn_kernel_ms, long long * gn_kernel_launches, int reset)
{
    if(int rc = check_handle(h)) return rc;
    if(int rc = set_device(h)) return rc;
    if(int rc = gn_fold_profile(h->gn)) return rc;
    if(gn_kernel_ms) *gn_kernel_ms = h->gn.kernel_ms;
    if(gn_kernel_launches) *gn_kernel_launches = h->gn.kernel_launches;
    if(reset)
    {
        h->gn.kernel_ms = 0.0;
        h->gn.kernel_launches = 0;
    }
    return SLAM_OK;
}

extern "C" int slam_odom_get_phase_cycles(slam_odom_t h, unsigned long long * out24, int reset)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(out24);
    if(int rc = set_device(h)) return rc;
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    SLAM_CUDA_TRY(cudaMemcpy(out24, h->gn.ctl->phase_cycles, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToHost));
    if(reset) SLAM_CUDA_TRY(cudaMemset(h->gn.ctl->phase_cycles, 0, sizeof(unsigned long long) * 24));
    return SLAM_OK;
}

extern "C" int slam_odom_set_split_launch(slam_odom_t h, int enable, int * previous)
{
    if(int rc = check_handle(h)) return rc;
    if(int rc = set_device(h)) return rc;
    if(int rc = gn_configure(h->gn)) return rc;
    if(previous) *previous = h->gn.split;
    h->gn.split = enable ? 1 : 0;
    return SLAM_OK;
}

extern "C" void * slam_odom_stream(slam_odom_t h) { return h ? (void *)h->stream : nullptr; }

extern "C" int slam_odom_set_trace(slam_odom_t h, int enable)
{
    if(int rc = check_handle(h)) return rc;
    h->trace_on = enable != 0;
    h->trace_level = enable;
    return SLAM_OK;
}

extern "C" int slam_odom_get_trace(slam_odom_t h, int seq, slam_step_record * out, int max_records, int * n_records)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(seq >= 0 && seq < h->batch && n_records);
    if(!h->p.host_loop)
    {
        if(h->pending_async)
            if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
        if(int rc = set_device(h)) return rc;
        return gn_read_trace(h->gn, seq, out, max_records, n_records, h->stream);
    }
    const auto & tr = h->trace[seq];
    *n_records = (int)tr.size();
    for(int i = 0; i < (int)tr.size() && i < max_records; i++) out[i] = tr[i];
    return SLAM_OK;
}

// ------------------------------------------------------------------ taps
extern "C" size_t slam_odom_tap_bytes(slam_odom_t h, int tap, int level)
{
    if(!h || level < 0 || level >= h->levels) return 0;
    const size_t n = (size_t)h->geom[level].rows * h->geom[level].cols;
    switch(tap)
    {
        case SLAM_TAP_DEPTH_U16: return n * 2;
        case SLAM_TAP_VMAP_CURR:
        case SLAM_TAP_NMAP_CURR:
        case SLAM_TAP_VMAP_PREV:
        case SLAM_TAP_NMAP_PREV: return n * 12;
        case SLAM_TAP_LAST_DEPTH:
        case SLAM_TAP_NEXT_DEPTH: return n * 4;
        case SLAM_TAP_LAST_IMAGE:
        case SLAM_TAP_NEXT_IMAGE:
        case SLAM_TAP_LASTNEXT_IMAGE: return n;
        case SLAM_TAP_DIDX:
        case SLAM_TAP_DIDY: return n * 2;
        case SLAM_TAP_CLOUD: return n * 12;
        case SLAM_TAP_CORRES: return n * 16;
    }
    return 0;
}

extern "C" int slam_odom_tap(slam_odom_t h, int tap, int level, int seq, void * host_dst, size_t bytes)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(host_dst && seq >= 0 && seq < h->batch && level >= 0 && level < h->levels);
    const size_t need = slam_odom_tap_bytes(h, tap, level);
    SLAM_ARG_CHECK(need != 0 && bytes >= need);
    if(int rc = set_device(h)) return rc;
    if(h->pending_async)
        if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
    SeqBuffers & s = h->seq[seq];
    const void * src = nullptr;
    switch(tap)
    {
        case SLAM_TAP_DEPTH_U16: src = s.depth[level]; break;
        case SLAM_TAP_VMAP_CURR: src = s.vcurr[level]; break;
        case SLAM_TAP_NMAP_CURR: src = s.ncurr[level]; break;
        case SLAM_TAP_VMAP_PREV: src = s.vprev[level]; break;
        case SLAM_TAP_NMAP_PREV: src = s.nprev[level]; break;
        case SLAM_TAP_LAST_DEPTH: src = s.lastDepth[level]; break;
        case SLAM_TAP_NEXT_DEPTH: src = s.nextDepth[level]; break;
        case SLAM_TAP_LAST_IMAGE: src = s.lastImage[level]; break;
        case SLAM_TAP_NEXT_IMAGE: src = s.nextImage[level]; break;
        case SLAM_TAP_LASTNEXT_IMAGE: src = s.lastNextImage[level]; break;
        case SLAM_TAP_DIDX:
        case SLAM_TAP_DIDY:
            if(h->deriv_stale)
            {
                // the last track derived its gradients in registers: produce the images now (computeDerivativeImages, utils.cu:579-638)
                int rows[SLAM_MAX_LEVELS], cols[SLAM_MAX_LEVELS];
                for(int l = 0; l < h->levels; l++)
                {
                    rows[l] = h->geom[l].rows;
                    cols[l] = h->geom[l].cols;
                }
                SeqBuffers & s0 = h->seq[0];
                if(int rc = launch_derivatives_simple(h->levels, h->deriv_src_swapped ? s0.lastNextImage : s0.nextImage, s0.dIdx, s0.dIdy, rows, cols, h->stream, h->batch, h->seq_stride))
                    return rc;
                h->launches++;
                h->deriv_stale = false;
            }
            src = tap == SLAM_TAP_DIDX ? (const void *)s.dIdx[level] : (const void *)s.dIdy[level];
            break;
        case SLAM_TAP_CORRES: src = s.corres[level]; break;
        case SLAM_TAP_CLOUD:
        {
            // pointClouds[level] is not materialised by the tracker (rgbStep re-derives the
            // point from lastDepth); build it on demand with the same operator.
            const LevelGeom & g = h->geom[level];
            float * tmp = nullptr;
            SLAM_CUDA_TRY(cudaMalloc((void **)&tmp, need));
            int rc = slam_op_project_to_point_cloud(s.lastDepth[level], g.rows, g.cols, tmp, h->p.fx, h->p.fy, h->p.cx, h->p.cy, level, h->stream);
            if(rc == SLAM_OK)
            {
                cudaError_t e = cudaMemcpyAsync(host_dst, tmp, need, cudaMemcpyDeviceToHost, h->stream);
                if(e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
                if(e != cudaSuccess)
                {
                    set_last_error(cudaGetErrorString(e));
                    rc = SLAM_ERR_CUDA;
                }
            }
            cudaFree(tmp);
            return rc;
        }
        default: return SLAM_ERR_ARG;
    }
    SLAM_CUDA_TRY(cudaMemcpyAsync(host_dst, src, need, cudaMemcpyDeviceToHost, h->stream));
    SLAM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return SLAM_OK;
}

// ------------------------------------------------------------------ per-frame front ends
// SLAM_ODOM_DEBUG_TIMING=1: host-side split of the per-frame call (enqueue of the preparation launches / of the persistent kernel / wait),
// printed every 256 frames -- a development aid for the launch-bound part of a frame.
struct FrameHostTiming
{
    double prep = 0, gn = 0, wait = 0;
    long long n = 0;
};
static thread_local FrameHostTiming g_frame_timing;
static inline double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int track_from_device_ptrs(slam_odom_t h, const unsigned short * depth, const uchar4 * rgba, const float4 * mv, const float4 * mn, const uchar4 * mrgba,
                                  const float * poses16, float depth_cutoff, float model_cutoff, float * trans, float * rot, int rgb_only,
                                  float icp_weight, int pyramid, int fast_odom, int so3, bool defer_wait = false)
{
    // apps/elastic_fusion_file.cpp:366-374: initICPModel -> initRGBModel -> initICP -> initRGB -> getIncrementalTransformation
    if(!h->trace_on && !h->p.host_loop)
    {
        // All five inputs are known up front: the current-frame depth branch (pyramid, vertex / normal maps) runs on a
        // second stream next to the model / RGB branch, and the "last" and "next" RGB-D pyramids are built together.
        static const bool debug_timing = getenv("SLAM_ODOM_DEBUG_TIMING") != nullptr;
        const double t_in = debug_timing ? now_us() : 0.0;
        if(int rc = set_device(h)) return rc;
        if(h->pending_async)
            if(int rc = finish_device_loop(h, nullptr, nullptr)) return rc;
        h->deriv_src_swapped = false;
        const size_t n0 = (size_t)h->geom[0].rows * h->geom[0].cols;
        // every launch covers all sequences of the batch (gridDim.y), see prep_kernels.cu: seq_shift
        const int B = h->batch;
        const size_t S = h->seq_stride;
        SeqBuffers & s = h->seq[0];
        int rc = SLAM_OK;
        // Three-level pyramids of a size that tiles evenly: the whole preparation is ONE launch (k_prepare_frame: shared-memory
        // tiles with halos, every level from the one above inside the block).  SLAM_ODOM_UNFUSED_PREP=1 keeps the per-level launches.
        static const bool unfused = getenv("SLAM_ODOM_UNFUSED_PREP") != nullptr;
        const uintptr_t align_all = reinterpret_cast<uintptr_t>(mv) | reinterpret_cast<uintptr_t>(mn) | reinterpret_cast<uintptr_t>(mrgba) | reinterpret_cast<uintptr_t>(rgba);
        const bool one_launch = !unfused && h->levels == 3 && h->geom[0].rows % 4 == 0 && h->geom[0].cols % 4 == 0 && (reinterpret_cast<uintptr_t>(depth) & 3) == 0 &&
                                (align_all & 15) == 0;   // 16-byte cp.async chunks, 32-bit depth words
        if(!one_launch)
        {
            SLAM_CUDA_TRY(cudaEventRecord(h->fork_ev, h->stream));
            SLAM_CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->fork_ev, 0));
        }
        for(int b = 0; b < B; b++)
        {
            const float * q = poses16 + 16 * b;
            float * o = h->h_poses12 + 12 * b;
            for(int r = 0; r < 3; r++)
            {
                for(int c = 0; c < 3; c++) o[3 * r + c] = q[4 * r + c];
                o[9 + r] = q[4 * r + 3];
            }
        }
        // one sequence: the pose travels in the kernel parameters (no copy in front of the first launch of the chain)
        if(B > 1) SLAM_CUDA_TRY(cudaMemcpyAsync(h->d_poses12, h->h_poses12, sizeof(float) * 12 * B, cudaMemcpyHostToDevice, h->stream));
        if(one_launch)
        {
            PrepFrameArgs a = {};
            a.rows = h->geom[0].rows;
            a.cols = h->geom[0].cols;
            a.vsrc = mv;
            a.nsrc = mn;
            a.model_rgba = mrgba;
            a.rgba = rgba;
            a.depth_tmp = s.depth_tmp;
            a.depth_cut = h->maxDepthRGB;
            a.depth = depth;
            a.depthCutoff = depth_cutoff;
            for(int l = 0; l < 3; l++)
            {
                a.vprev[l] = s.vprev[l];
                a.nprev[l] = s.nprev[l];
                a.lastDepth[l] = s.lastDepth[l];
                a.nextDepth[l] = s.nextDepth[l];
                a.lastImage[l] = s.lastImage[l];
                a.nextImage[l] = s.nextImage[l];
                a.depth_l[l] = s.depth[l];
                a.vcurr[l] = s.vcurr[l];
                a.ncurr[l] = s.ncurr[l];
                a.fx_inv[l] = 1.f / h->geom[l].fx;
                a.fy_inv[l] = 1.f / h->geom[l].fy;
                a.cx[l] = h->geom[l].cx;
                a.cy[l] = h->geom[l].cy;
            }
            if(B == 1)
            {
                const float * q = h->h_poses12;
                a.R.r0 = make_float3(q[0], q[1], q[2]);
                a.R.r1 = make_float3(q[3], q[4], q[5]);
                a.R.r2 = make_float3(q[6], q[7], q[8]);
                a.t = make_float3(q[9], q[10], q[11]);
            }
            else
                a.poses12 = h->d_poses12;
            a.map_in_stride = n0 * 16;
            a.rgba_stride = n0 * 4;
            a.depth_in_stride = n0 * 2;
            a.arena_stride = S;
            if(int rc2 = launch_prepare_frame(a, h->stream, B)) return rc2;
            h->launches++;
            h->have_depth_tmp = true;
        }
        else
        {
            Mat3 R0 = {};
            float3 t0 = make_float3(0, 0, 0);
            if(B == 1)
            {
                const float * q = h->h_poses12;
                R0.r0 = make_float3(q[0], q[1], q[2]);
                R0.r1 = make_float3(q[3], q[4], q[5]);
                R0.r2 = make_float3(q[6], q[7], q[8]);
                t0 = make_float3(q[9], q[10], q[11]);
            }
            rc = launch_model_maps_simple(mv, mn, h->geom[0].rows, h->geom[0].cols, h->levels, s.vprev, s.nprev, 1, R0, t0, s.depth_tmp,
                                          h->maxDepthRGB, h->levels > 3 ? s.vcam : nullptr, h->levels > 3 ? s.ncam : nullptr, h->stream, B, n0 * 16, S,
                                          B > 1 ? h->d_poses12 : nullptr, s.lastDepth[0], s.nextDepth[0], mrgba, rgba, s.lastImage[0], s.nextImage[0], n0 * 4);
            if(rc) return rc;
            h->launches++;
            if(h->levels > 3)
                for(int b = 0; b < B; b++)
                {
                    const float * q = h->h_poses12 + 12 * b;
                    Mat3 R;
                    R.r0 = make_float3(q[0], q[1], q[2]);
                    R.r1 = make_float3(q[3], q[4], q[5]);
                    R.r2 = make_float3(q[6], q[7], q[8]);
                    SeqBuffers & sb = h->seq[b];
                    rc = launch_resize_transform(sb.vcam, sb.ncam, h->geom[2].rows, h->geom[2].cols, sb.vprev[3], sb.nprev[3], 1, R, make_float3(q[9], q[10], q[11]),
                                                 nullptr, nullptr, h->stream);
                    if(rc) return rc;
                    h->launches++;
                }
            // level 0 of both RGB-D pyramids was written by the model-map launch
            for(int l = 0; l + 1 < h->levels; l++)
            {
                rc = launch_rgbd_down_dual(s.lastDepth[l], s.lastDepth[l + 1], s.nextDepth[l + 1], s.lastImage[l], s.lastImage[l + 1], s.nextImage[l],
                                           s.nextImage[l + 1], h->geom[l].rows, h->geom[l].cols, h->stream, B, S);
                if(rc) return rc;
                h->launches++;
            }
        }
        // the depth branch is enqueued after the model / RGB branch (it only waits for fork_ev); alternating the two chains' launches
        // was measured and makes no difference: the six preparation kernels (~45 us of device time) serialise on the GPU either way
        if(!one_launch)
        {
            for(int l = 0; l < h->levels && rc == SLAM_OK; l++)
            {
                const LevelGeom & g = h->geom[l];
                rc = launch_depth_level(l == 0 ? depth : s.depth[l], g.rows, g.cols, g.fx, g.fy, g.cx, g.cy, depth_cutoff, s.vcurr[l], s.ncurr[l],
                                        l + 1 < h->levels ? s.depth[l + 1] : nullptr, h->aux_stream, B, l == 0 ? n0 * 2 : S, S);
                h->launches++;
            }
            if(rc) return rc;
            SLAM_CUDA_TRY(cudaEventRecord(h->join_ev, h->aux_stream));
            h->have_depth_tmp = true;
            SLAM_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->join_ev, 0));
        }
        if(defer_wait) return slam_odom_get_incremental_transformation_async(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
        if(debug_timing)
        {
            const double t_prep = now_us();
            int rc2 = slam_odom_get_incremental_transformation_async(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
            const double t_gn = now_us();
            if(rc2 == SLAM_OK) rc2 = finish_device_loop(h, trans, rot);
            const double t_done = now_us();
            FrameHostTiming & ft = g_frame_timing;
            ft.prep += t_prep - t_in, ft.gn += t_gn - t_prep, ft.wait += t_done - t_gn;
            if(++ft.n % 256 == 0)
            {
                fprintf(stderr, "slam_odom frame host timing: enqueue prep %.1f us, enqueue persistent kernel %.1f us, wait %.1f us (mean of 256)\n", ft.prep / 256,
                        ft.gn / 256, ft.wait / 256);
                ft.prep = ft.gn = ft.wait = 0;
            }
            return rc2;
        }
        return slam_odom_get_incremental_transformation(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
    }
    if(int rc = slam_odom_init_icp_model(h, (const float *)mv, (const float *)mn, model_cutoff, poses16)) return rc;
    if(int rc = slam_odom_init_rgb_model(h, (const uint8_t *)mrgba)) return rc;
    if(int rc = slam_odom_init_icp_depth(h, depth, 0, depth_cutoff)) return rc;
    if(int rc = slam_odom_init_rgb(h, (const uint8_t *)rgba)) return rc;
    return slam_odom_get_incremental_transformation(h, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
}

extern "C" int slam_odom_track_device(slam_odom_t h, const slam_frame_host * f, float * trans, float * rot, int rgb_only, float icp_weight, int pyramid,
                                      int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(f && f->depth && f->rgba && f->model_vertices4 && f->model_normals4 && f->model_rgba && f->model_pose16 && trans && rot);
    return track_from_device_ptrs(h, f->depth, (const uchar4 *)f->rgba, (const float4 *)f->model_vertices4, (const float4 *)f->model_normals4,
                                  (const uchar4 *)f->model_rgba, f->model_pose16, f->depth_cutoff, f->model_depth_cutoff, trans, rot, rgb_only,
                                  icp_weight, pyramid, fast_odom, so3);
}

static int ensure_staging(slam_odom_t h)
{
    if(h->staging_ready) return SLAM_OK;
    const size_t n = (size_t)h->geom[0].rows * h->geom[0].cols * h->batch;
    for(auto & sl : h->slot)
    {
        // one allocation per slot: depth | rgba | mv | mn | mrgba
        const size_t bytes = align_up(n * 2, 256) + align_up(n * 4, 256) * 2 + align_up(n * 16, 256) * 2;
        char * base = nullptr;
        SLAM_CUDA_TRY(cudaMalloc((void **)&base, bytes));
        sl.depth = (unsigned short *)base;
        base += align_up(n * 2, 256);
        sl.rgba = (uchar4 *)base;
        base += align_up(n * 4, 256);
        sl.mrgba = (uchar4 *)base;
        base += align_up(n * 4, 256);
        sl.mv = (float4 *)base;
        base += align_up(n * 16, 256);
        sl.mn = (float4 *)base;
        sl.poses.resize(16 * h->batch);
        SLAM_CUDA_TRY(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
    }
    h->staging_ready = true;
    return SLAM_OK;
}

static int stage_frame(slam_odom_t h, StagingSlot & sl, const slam_frame_host * f, cudaStream_t cs)
{
    const size_t n = (size_t)h->geom[0].rows * h->geom[0].cols * h->batch;
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.mv, f->model_vertices4, n * 16, cudaMemcpyHostToDevice, cs));
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.mn, f->model_normals4, n * 16, cudaMemcpyHostToDevice, cs));
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.mrgba, f->model_rgba, n * 4, cudaMemcpyHostToDevice, cs));
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.depth, f->depth, n * 2, cudaMemcpyHostToDevice, cs));
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.rgba, f->rgba, n * 4, cudaMemcpyHostToDevice, cs));
    memcpy(sl.poses.data(), f->model_pose16, sizeof(float) * 16 * h->batch);
    sl.depth_cutoff = f->depth_cutoff;
    sl.model_depth_cutoff = f->model_depth_cutoff;
    sl.tag_depth = f->depth;
    sl.generation = ++h->stage_generation;
    SLAM_CUDA_TRY(cudaEventRecord(sl.ready, cs));
    sl.pending = true;
    return SLAM_OK;
}

// depth + rgba only (slam_odom_track_sensor): the model prediction stays where the caller rendered it
static int stage_sensor(slam_odom_t h, StagingSlot & sl, const slam_frame_host * f, cudaStream_t cs)
{
    const size_t n = (size_t)h->geom[0].rows * h->geom[0].cols * h->batch;
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.depth, f->depth, n * 2, cudaMemcpyHostToDevice, cs));
    SLAM_CUDA_TRY(cudaMemcpyAsync(sl.rgba, f->rgba, n * 4, cudaMemcpyHostToDevice, cs));
    sl.tag_depth = f->depth;
    sl.generation = ++h->stage_generation;
    SLAM_CUDA_TRY(cudaEventRecord(sl.ready, cs));
    sl.pending = true;
    return SLAM_OK;
}

static StagingSlot * find_staged(slam_odom_t h, const void * tag)
{
    for(auto & sl : h->slot)
        if(sl.pending && sl.tag_depth == tag) return &sl;
    return nullptr;
}

static StagingSlot * free_slot(slam_odom_t h)
{
    for(auto & sl : h->slot)
        if(!sl.pending) return &sl;
    return nullptr;
}

extern "C" int slam_odom_prefetch_host(slam_odom_t h, const slam_frame_host * f)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(f && f->depth && f->rgba && f->model_vertices4 && f->model_normals4 && f->model_rgba && f->model_pose16);
    if(int rc = set_device(h)) return rc;
    if(int rc = ensure_staging(h)) return rc;
    if(find_staged(h, f->depth)) return SLAM_OK;
    StagingSlot * sl = free_slot(h);
    if(!sl)
    {
        set_last_error("prefetch: both staging slots hold frames that were not tracked yet");
        return SLAM_ERR_ORDER;
    }
    // a free slot was last read by kernels of a frame whose track_host already returned (it synchronises),
    // so the copy stream may overwrite it now
    return stage_frame(h, *sl, f, h->copy_stream);
}

extern "C" int slam_odom_track_host(slam_odom_t h, const slam_frame_host * f, float * trans, float * rot, int rgb_only, float icp_weight, int pyramid,
                                    int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(f && f->depth && f->rgba && f->model_vertices4 && f->model_normals4 && f->model_rgba && f->model_pose16 && trans && rot);
    if(int rc = set_device(h)) return rc;
    if(int rc = ensure_staging(h)) return rc;
    StagingSlot * sl = find_staged(h, f->depth);
    if(!sl)
    {
        sl = free_slot(h);
        if(!sl)
        {
            sl = h->slot[0].generation <= h->slot[1].generation ? &h->slot[0] : &h->slot[1];   // drop the older stale prefetch
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
        }
        if(int rc = stage_frame(h, *sl, f, h->copy_stream)) return rc;
    }
    SLAM_CUDA_TRY(cudaStreamWaitEvent(h->stream, sl->ready, 0));
    const int rc = track_from_device_ptrs(h, sl->depth, sl->rgba, sl->mv, sl->mn, sl->mrgba, sl->poses.data(), sl->depth_cutoff, sl->model_depth_cutoff, trans,
                                          rot, rgb_only, icp_weight, pyramid, fast_odom, so3);
    sl->pending = false;
    return rc;
}

extern "C" int slam_odom_track_host_next(slam_odom_t h, const slam_frame_host * f, const slam_frame_host * next, float * trans, float * rot, int rgb_only,
                                         float icp_weight, int pyramid, int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(f && f->depth && f->rgba && f->model_vertices4 && f->model_normals4 && f->model_rgba && f->model_pose16 && trans && rot);
    if(int rc = set_device(h)) return rc;
    if(int rc = ensure_staging(h)) return rc;
    StagingSlot * sl = find_staged(h, f->depth);
    if(!sl)
    {
        sl = free_slot(h);
        if(!sl)
        {
            sl = h->slot[0].generation <= h->slot[1].generation ? &h->slot[0] : &h->slot[1];   // drop the older stale prefetch
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
        }
        if(int rc = stage_frame(h, *sl, f, h->copy_stream)) return rc;
    }
    SLAM_CUDA_TRY(cudaStreamWaitEvent(h->stream, sl->ready, 0));
    // enqueue this frame's work, and only then spend host time on the next frame's copies: they overlap the kernels
    int rc = track_from_device_ptrs(h, sl->depth, sl->rgba, sl->mv, sl->mn, sl->mrgba, sl->poses.data(), sl->depth_cutoff, sl->model_depth_cutoff, trans, rot,
                                    rgb_only, icp_weight, pyramid, fast_odom, so3, true);
    if(rc) return rc;
    if(next && next->depth && !find_staged(h, next->depth))
    {
        StagingSlot * other = nullptr;
        for(auto & cand : h->slot)
            if(&cand != sl && !cand.pending) other = &cand;
        if(other) rc = stage_frame(h, *other, next, h->copy_stream);   // the slot still in use by this frame's kernels is left alone
    }
    if(h->pending_async)
    {
        const int rc2 = finish_device_loop(h, trans, rot);
        if(!rc) rc = rc2;
    }
    sl->pending = false;
    return rc;
}

extern "C" int slam_odom_track_sensor(slam_odom_t h, const slam_frame_host * f, const slam_frame_host * next, float * trans, float * rot, int rgb_only,
                                      float icp_weight, int pyramid, int fast_odom, int so3)
{
    if(int rc = check_handle(h)) return rc;
    SLAM_ARG_CHECK(f && f->depth && f->rgba && f->model_vertices4 && f->model_normals4 && f->model_rgba && f->model_pose16 && trans && rot);
    if(int rc = set_device(h)) return rc;
    if(int rc = ensure_staging(h)) return rc;
    StagingSlot * sl = find_staged(h, f->depth);
    if(!sl)
    {
        sl = free_slot(h);
        if(!sl)
        {
            sl = h->slot[0].generation <= h->slot[1].generation ? &h->slot[0] : &h->slot[1];   // drop the older stale prefetch
            SLAM_CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
        }
        if(int rc = stage_sensor(h, *sl, f, h->copy_stream)) return rc;
    }
    SLAM_CUDA_TRY(cudaStreamWaitEvent(h->stream, sl->ready, 0));
    // enqueue this frame's work, and only then spend host time on the next frame's copies: they overlap the kernels
    int rc = track_from_device_ptrs(h, sl->depth, sl->rgba, (const float4 *)f->model_vertices4, (const float4 *)f->model_normals4, (const uchar4 *)f->model_rgba,
                                    f->model_pose16, f->depth_cutoff, f->model_depth_cutoff, trans, rot, rgb_only, icp_weight, pyramid, fast_odom, so3, true);
    if(rc) return rc;
    if(next && next->depth && next->rgba && !find_staged(h, next->depth))
    {
        StagingSlot * other = nullptr;
        for(auto & cand : h->slot)
            if(&cand != sl && !cand.pending) other = &cand;
        if(other) rc = stage_sensor(h, *other, next, h->copy_stream);   // the slot still in use by this frame's kernels is left alone
    }
    if(h->pending_async)
    {
        const int rc2 = finish_device_loop(h, trans, rot);
        if(!rc) rc = rc2;
    }
    sl->pending = false;
    return rc;
}
