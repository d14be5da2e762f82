// Batched streaming engine: declarations (see batch_engine.cu).
#pragma once
#include <vector>
#include "gn_kernel.cuh"

namespace slam {

constexpr int kBatchEngineMin = 4;   // batches of at least this many sequences use the streaming engine

struct BatchDevice
{
    char * states = nullptr;        // [batch][state_stride]: persistent prefix of GnShared per sequence
    size_t state_stride = 0;
    float * sums = nullptr;         // [batch][64]: folded sums of the last phase A (0..31) / phase B (32..63)
    GnSeqIn * seq_in = nullptr;     // [batch] (shared with the persistent kernel's array)
    GnResult * results = nullptr;   // [batch]
    unsigned char * cand0 = nullptr;    // candidate masks, vmask: per sequence, per level
    size_t aux_stride = 0;              // bytes between consecutive sequences' aux blocks
    size_t cand_off[SLAM_MAX_LEVELS], vmask_off[SLAM_MAX_LEVELS];
    char * ws = nullptr;            // [batch][kWorkspaceBytes]: partial rows + ticket of each sequence's running reduction
    char * ws2 = nullptr;           // [batch][kWorkspaceBytes]: the same for the RGB association when it runs as its own launch
    int num_sms = 148;
    struct GraphEntry
    {
        std::vector<char> key;      // GnLaunch bytes + the pointers / counts the launch sequence depends on
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;
    };
    std::vector<GraphEntry> graphs;   // captured launch sequences of a step, one per key
    bool cand_ready = false;        // the candidate masks of this frame were written by the fused derivative launch
    std::vector<cudaStream_t> role;       // per sequence group: stream of the RGB association role
    std::vector<cudaEvent_t> role_done, role_go;
    std::vector<cudaStream_t> side;       // streams of the sequence groups beyond the first
    std::vector<cudaEvent_t> side_done;
    cudaEvent_t fork = nullptr;
    int batch = 0;
    long long launches = 0;
};

size_t batch_state_bytes(int batch, const LevelGeom * geom, int levels);
void batch_bind_state(BatchDevice & d, char * base, int batch, const LevelGeom * geom, int levels, GnSeqIn * seq_in, GnResult * results);
int batch_enqueue(BatchDevice & d, const GnLaunch & L, const GnSeqIn * h_seq_in_pinned, GnResult * h_results, slam_step_record * trace, int * trace_count,
                  cudaStream_t stream, std::vector<cudaEvent_t> * prof_events = nullptr);

void batch_release(BatchDevice & d);
void batch_report();   // SLAM_BATCH_DETAIL=1: print per-kernel event totals to stderr

}   // namespace slam
