// Device-resident Gauss-Newton loop: host-visible declarations (see gn_kernel.cu).
#pragma once
#include <vector>
#include "odom_internal.hpp"

namespace slam {

constexpr int kGnThreads = 512;
constexpr int kGnMaxCtas = 256;        // >= SM count of any target part (B200: 148)
constexpr int kGnPartialStride = 64;   // floats per CTA per buffer: [0..31] phase A, [32..63] phase B
constexpr int kGnMaxTrace = 64;        // step records kept per sequence
constexpr int kSlotChunk = 5;          // pixels per thread and level that stay register-resident (640x480 on 148 CTAs: 5)

// CTAs per sequence group for a batch on a device with num_sms SMs (must match gn_enqueue).
inline int gn_group_size(int num_sms, int batch) { return batch >= num_sms ? 1 : num_sms / batch; }

// Per-sequence inputs of one launch; refreshed by ONE host->device copy per launch (the
// image pointers change when the reference swaps lastNextImage/nextImage, and the prior
// pose comes from the caller).
struct GnSeqIn
{
    const float * vcurr[SLAM_MAX_LEVELS];
    const float * ncurr[SLAM_MAX_LEVELS];
    const float * vprev[SLAM_MAX_LEVELS];
    const float * nprev[SLAM_MAX_LEVELS];
    const float * lastDepth[SLAM_MAX_LEVELS];
    const float * nextDepth[SLAM_MAX_LEVELS];
    const unsigned char * lastImage[SLAM_MAX_LEVELS];
    const unsigned char * nextImage[SLAM_MAX_LEVELS];
    const unsigned char * lastNextImage[SLAM_MAX_LEVELS];
    const short * dIdx[SLAM_MAX_LEVELS];
    const short * dIdy[SLAM_MAX_LEVELS];
    Corres * corres[SLAM_MAX_LEVELS];
    float Rprev[9];
    float tprev[3];
};

struct GnCtl
{
    // one 64-bit word per CTA group, never reset: high half = arrival counter, low half = running sum of what the
    // arriving CTAs contribute at the mid-iteration barrier (RGB correspondence count, see group_barrier_sum)
    unsigned long long barrier[kGnMaxCtas];
    unsigned base[kGnMaxCtas];      // arrival counter at the end of the previous launch (written by the group leader)
    unsigned sum_base[kGnMaxCtas];  // running sum at the end of the previous launch
};

struct GnLaunch
{
    int levels, batch;
    LevelGeom geom[SLAM_MAX_LEVELS];
    int iterations[SLAM_MAX_LEVELS];
    bool icp, rgb, rgb_only, so3, trace, full_corres;
    bool derive_gradients;   // no derivative images were made: single-chunk groups derive them from nextImage in registers
    float icp_weight;
    float dist_thresh, angle_thresh;
    float sobel_scale, max_depth_delta;
    float min_scale[SLAM_MAX_LEVELS];
};

struct GnDevice
{
    // device
    GnCtl * ctl = nullptr;
    GnSeqIn * seq_in = nullptr;
    float * partials = nullptr;
    GnResult * results = nullptr;
    slam_step_record * trace = nullptr;
    int * trace_count = nullptr;
    // host
    char * h_stage = nullptr;       // pinned image of [GnCtl | GnSeqIn x batch]
    size_t stage_bytes = 0;
    int batch = 0;
    int num_sms = 0;
    bool so3_swapped = false;
    // optional CUDA-event timing of the persistent kernel (bench.py's roofline numerator)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;      // start/stop pairs of launches not yet folded into the totals
    double kernel_ms = 0.0;
    long long kernel_launches = 0;
};

int gn_fold_profile(GnDevice & d);

size_t gn_state_bytes(int batch);
void gn_bind_state(GnDevice & d, char * base, int batch);
int gn_stage_inputs(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnSeqIn ** out);
// h_flags != nullptr: zero-copy completion -- the kernel writes h_results (mapped pinned memory) itself and then stores seqno into
// h_flags[seq]; otherwise the results follow with a D2H copy on `stream`.
int gn_enqueue(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnResult * h_results,
               cudaStream_t stream, unsigned * h_flags = nullptr, unsigned seqno = 0);
int gn_read_trace(GnDevice & d, int seq, slam_step_record * out, int max_records, int * n_records, cudaStream_t stream);
void gn_release(GnDevice & d);

}   // namespace slam
