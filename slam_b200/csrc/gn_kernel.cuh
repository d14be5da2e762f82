// Device-resident Gauss-Newton loop: host-visible declarations (see gn_kernel.cu).
#pragma once
#include <vector>
#include "odom_internal.hpp"

namespace slam {

constexpr int kGnThreads = 512;
constexpr int kGnWarps = kGnThreads / 32;
constexpr int kGnMaxCtas = 255;        // >= SM count of any target part (B200: 148); the arrival count of a reduction word has 8 bits
constexpr int kGnMaxTrace = 64;        // step records kept per sequence

// ---- the inter-CTA all-reduce: fixed-point words in L2 -------------------------------------------------------------------
// Every reduced column is carried by two 64-bit words (integer part | fraction * 2^48 of each CTA's fp32 partial sum), each
// with the number of CTAs that have contributed in its low 8 bits: ONE relaxed atomic add per word publishes the data and
// the arrival together, and the readers poll the words themselves (one trip through L2 instead of store + fence + arrival +
// poll + fold; integer addition makes the sums independent of the arrival order, i.e. deterministic).  A ring of four word
// sets is used round-robin by consecutive steps; the words are never cleared -- the readers take differences against the value
// they consumed last (sums modulo 2^56, counts modulo 256).
constexpr int kRingSlots = 4;
constexpr int kWIcp = 0;               // 29 columns x 2 words: JtJJtrSE3 of icpStep
constexpr int kWRgb = 58;              // 29 columns x 2 words: JtJJtrSE3 of rgbStep
constexpr int kWSo3 = 116;             // 11 columns x 2 words: JtJJtrSO3
constexpr int kWMid = 138;             // RGB correspondence count (+ 2^32 per CTA whose squared-residual sum is non-zero)
constexpr int kWSigma = 139;           // squared-residual sum of the RGB association (int)
constexpr int kRingWords = 140;
constexpr int kWordStride = 32;        // in 64-bit units: the words are 256 B apart (different L2 slices)
constexpr size_t kRingBytes = (size_t)kRingSlots * kRingWords * kWordStride * 8;

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

// What the coarse-level kernel (one thread-block cluster: SO3 pre-alignment + the small pyramid levels) hands to the fine-level
// kernel that runs next to it (split launch, gn_kernel.cu): the state that carries over from one level to the next.
constexpr int kSplitMaxSeqs = 16;      // sequences one split launch pair works through, one after the other
struct GnHandoff
{
    GnResult res;
    double resultRt[16];
    float Rcurr[9], tcurr[3];
    int rgb_sigma_last, rgb_count_last;
    int timeouts, pad;
};

struct GnCtl
{
    unsigned long long arrived[kGnMaxCtas + 1];   // per CTA group: CTAs that have checked in, summed over all launches (GN_GATE)
    unsigned timeouts;               // polls that gave up (a lost arrival would otherwise hang the GPU): non-zero = results invalid
    unsigned long long phase_cycles[24];   // SM cycles the leading CTA spent per phase, accumulated over launches (slam_odom_get_phase_cycles)
    // split launch, per sequence: written by the leading CTA of the cluster, then handoff_seq = number of the launch (release); polled
    // by the fine-level kernel
    alignas(16) GnHandoff handoff[kSplitMaxSeqs];
    unsigned long long handoff_seq[kSplitMaxSeqs];
    unsigned long long dbg_t[4];     // phase accounting only: [0] = %globaltimer at the start of the cluster kernel
};

// Split launch: the cluster that runs the SO3 pre-alignment and the coarse levels.  Its all-reduce goes through distributed shared
// memory (every CTA stores its partial row into every peer, one barrier.cluster, every CTA folds the rows in rank order).
constexpr int kClusterCtas = 16;       // non-portable cluster size (needs cudaFuncAttributeNonPortableClusterSizeAllowed)
constexpr int kClusterMaxPixels = 24576;   // levels up to this many pixels run on the cluster
struct ClArea
{
    float rows[2][kClusterCtas][64];   // [parity][rank of the writer][column]: [0..31] ICP / SO3, [32..63] RGB
    int2 mid[2][kClusterCtas];         // RGB correspondence count, squared-residual sum of every CTA
};

// How one pyramid level is mapped onto the CTAs of a group.  Resident levels keep the pose-independent operands of their
// pixels in shared memory for the whole frame: the CTA owns every P-th 32-pixel segment of the image, and the pixels that can
// take part at all (valid current vertex + normal for ICP; the pose-independent half of the RGB association for RGB) are
// compacted into dense lists, so no lane idles on a dead pixel.  Non-resident levels (image too large for shared memory) stream
// their operands from L2 every iteration.
struct LevelPlan
{
    int resident;        // operands staged in shared memory
    int P;               // CTAs of the group that take part (small levels: fewer CTAs, fewer arrivals)
    int nseg;            // 32-pixel segments of the level
    int segs_per_cta;    // segments a participating CTA owns at most
    int cap;             // capacity of the CTA's lists (entries)
    int off_icp;         // byte offsets into the dynamic shared memory: ICP list, 6 float planes of cap entries
    int off_rgb;         // RGB list: nextDepth (float), gradients (2 x s16), x | y << 11 | intensity << 22 (u32), cap entries each
};

struct GnLaunch
{
    int levels, batch;
    LevelGeom geom[SLAM_MAX_LEVELS];
    int iterations[SLAM_MAX_LEVELS];
    bool icp, rgb, rgb_only, so3, trace, full_corres;
    bool derive_gradients;   // no derivative images were made: resident levels derive them from nextImage while staging
    float icp_weight;
    float dist_thresh, angle_thresh;
    float sobel_scale, max_depth_delta;
    float min_scale[SLAM_MAX_LEVELS];
    // shared-memory plan (gn_make_plan)
    LevelPlan plan[SLAM_MAX_LEVELS];
    int off_state;           // phase A -> phase B state of the RGB entries (3 words per entry), shared by all levels
    int so3_resident, so3_P; // SO3 pre-alignment: both level-2 images in shared memory; participating CTAs
    int off_so3;
    int off_cl;              // split launch, cluster kernel: the ClArea
    int dyn_bytes;
    int poll_delay;          // cycles between posting a contribution and the first poll
    double K[SLAM_MAX_LEVELS][9], Kinv[SLAM_MAX_LEVELS][9];   // intrinsics of every level and their inverse (gn_make_plan; the kernels used to
                                                              // form them with one thread at every level change, on the chain)
    int ph_role;             // phase accounting: 0 = both kernels of a split launch add their cycles, 1 / 2 = only that role
};

struct GnDevice
{
    // device
    GnCtl * ctl = nullptr;
    GnSeqIn * seq_in = nullptr;
    unsigned long long * ring = nullptr;   // [groups][kRingSlots][kRingWords] words, kWordStride apart
    GnResult * results = nullptr;
    slam_step_record * trace = nullptr;
    int * trace_count = nullptr;
    // host
    char * h_stage = nullptr;       // pinned image of [GnSeqIn x batch]
    size_t stage_bytes = 0;
    int batch = 0;
    int num_sms = 0;
    int smem_limit = 0;             // dynamic shared memory the persistent kernel may use
    int poll_delay = 0;
    unsigned long long launch_no = 0;   // launches so far
    unsigned long long gate_total = 0;  // check-ins every group's counter has seen so far (grows by the group size of each launch)
    int split = -1;                 // SLAM_GN_SPLIT: 1 = cluster + fine-level launch pair where it applies (default), 0 = one launch
    int last_launches = 1;          // kernels the last gn_enqueue put into the stream (2 for a split launch)
    int seq_max = 8;                // SLAM_GN_SEQ_MAX: batches of 3 .. seq_max sequences run as ONE split launch pair, sequence after sequence
                                    // (measured, frames/s at 2 / 3 / 4 / 8 / 12 / 16 sequences: pair 5430 / 5460 / 5470 / 5640 / 5720 / 5740, against 5710 / 3730
                                    // for concurrent CTA groups at 2 / 3 and 4040 / 5100 / 6450 / 7170 for the streaming engine at 4 / 8 / 12 / 16)
    int split_broken = 0;           // the device refused the cluster / programmatic launch once: stay with one launch
    bool phases = false;            // SLAM_GN_PHASES: launch the variants with per-phase cycle accounting
    bool so3_swapped = false;
    int device = -1;                // CUDA device of the handle (pair gate below)
    bool gate_member = false;
    cudaEvent_t pair_done = nullptr;   // recorded behind every split launch pair when several handles share the device
    // optional CUDA-event timing of the persistent kernel (bench.py's roofline numerator)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;      // start/stop pairs of launches not yet folded into the totals
    std::vector<cudaEvent_t> ev_pool; // folded events, reused (creating two events per frame showed up as host hiccups in short timed regions)
    double kernel_ms = 0.0;
    long long kernel_launches = 0;
};

int gn_fold_profile(GnDevice & d);
// One-time device set-up of the persistent kernel (attributes, limits, environment switches); idempotent.
int gn_configure(GnDevice & d);
// Would this launch (levels / geom / iterations / batch / mode / trace flags set) run as the split pair?
bool gn_split_applies(GnDevice & d, const GnLaunch & L);

size_t gn_state_bytes(int batch, int num_sms);
void gn_bind_state(GnDevice & d, char * base, int batch, int num_sms);
int gn_stage_inputs(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnSeqIn ** out);
// Fill the shared-memory plan of a launch (L.geom / levels / icp / rgb / so3 set): which levels are resident, list capacities,
// offsets.  Returns true when every RGB level is resident (no derivative images needed from memory).
bool gn_make_plan(GnDevice & d, GnLaunch & L);
// h_flags != nullptr: zero-copy completion -- the kernel writes h_results (mapped pinned memory) itself and then stores seqno into
// h_flags[seq]; otherwise the results follow with a D2H copy on `stream`.
int gn_enqueue(GnDevice & d, const GnLaunch & L, const SeqBuffers * seqs, const float * trans, const float * rot, GnResult * h_results,
               cudaStream_t stream, unsigned * h_flags = nullptr, unsigned seqno = 0);
int gn_read_trace(GnDevice & d, int seq, slam_step_record * out, int max_records, int * n_records, cudaStream_t stream);
void gn_release(GnDevice & d);

}   // namespace slam
