/*
 * slam_odom.h -- C ABI of the B200-native RGB-D camera tracker (libslam_odom.so).
 *
 * Drop-in boundary for ONE path of siw-engineering/slam: the `RGBDOdometryef`
 * frame-to-model tracker (ICP + RGB + SO3 pre-alignment).  The reference exposes no
 * FFI layer; its boundary is the C++ class in src/odom/RGBDOdometryef.h:28-70 and,
 * beneath it, the 17 free host wrappers of src/odom/utils.cuh:62-175.  Each entry
 * point below names the reference interface it replaces.  A header-only C++ class
 * with the reference's method names sits on top (include/RGBDOdometryef.hpp).
 *
 * Conventions
 *  - plain pointers and sizes only; all image pointers are DEVICE pointers unless the
 *    name says `_host`; images are dense row-major (no pitch) unless a pitch is given;
 *  - where the reference takes a `GPUTexture*` (GL texture it immediately copies to
 *    linear memory, RGBDOdometryef.cpp:126,150,155,178,183) we take the linear device
 *    pointer of the same texel format: R16UI depth -> uint16 mm, RGBA8 -> 4 x uint8,
 *    RGBA32F vertex/normal -> 4 x float (x,y,z,conf / nx,ny,nz,radius), camera frame;
 *  - every function returns SLAM_OK (0) or a negative error code (the reference prints
 *    and exit(0)s on CUDA errors, cuda/convenience.cuh:64-71; we return instead);
 *  - a handle is bound to one GPU and one stream; it holds `batch` independent
 *    sequences.  For batch > 1 every image argument points at `batch` consecutive
 *    dense images and pose arguments at `batch` consecutive poses;
 *  - handles share no global state: distinct handles may be used from distinct host
 *    threads concurrently (the reference is not re-entrant, utils.cu:548,610).
 *  - there is NO CPU fallback: without a CUDA device slam_odom_create fails with
 *    SLAM_ERR_CUDA.
 */
#ifndef SLAM_ODOM_H_
#define SLAM_ODOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLAM_OK 0
#define SLAM_ERR_CUDA (-1)      /* a CUDA call failed; see slam_odom_last_error() */
#define SLAM_ERR_ARG (-2)       /* bad argument */
#define SLAM_ERR_ORDER (-3)     /* call-order contract violated (initICP* before initRGB*) */
#define SLAM_ERR_UNSUPPORTED (-4)

#define SLAM_MAX_LEVELS 4

typedef struct slam_odom * slam_odom_t;

/* ctor arguments of RGBDOdometryef (RGBDOdometryef.h:32-36) plus the knobs the
 * reference hard-codes (NUM_PYRS, iteration list) or lacks (device, stream, batch). */
typedef struct slam_odom_params
{
    int width, height;
    float cx, cy, fx, fy;
    float dist_thresh;          /* 0 => 0.10f                (RGBDOdometryef.h:35) */
    float angle_thresh;         /* 0 => sinf(20*3.14159254f/180)  (RGBDOdometryef.h:36) */
    int num_levels;             /* 0 => 3 (NUM_PYRS, RGBDOdometryef.h:104); 1..4 */
    int iterations[SLAM_MAX_LEVELS]; /* all 0 => reference rule {fastOdom?3:10, pyramid?5:0, pyramid?4:0}
                                        (RGBDOdometryef.cpp:382-384); otherwise used verbatim */
    int device;                 /* CUDA device ordinal */
    void * stream;              /* cudaStream_t to run on; NULL => the library creates one */
    int batch;                  /* 0/1 => one sequence; >1 => that many independent sequences */
    int host_loop;              /* 0 => device-resident Gauss-Newton loop (default);
                                   1 => reference-style host-stepped loop (one sync per step) */
} slam_odom_params;

/* Public result fields of RGBDOdometryef (RGBDOdometryef.h:62-70). */
typedef struct slam_odom_stats
{
    float lastICPError, lastICPCount;
    float lastRGBError, lastRGBCount;
    float lastSO3Error, lastSO3Count;
    double lastA[36];           /* row-major 6x6 */
    double lastb[6];
    int so3_iterations;         /* how many SO3 steps ran (not in the reference; diagnostic) */
    int gn_iterations;          /* how many ICP/RGB steps ran */
} slam_odom_stats;

const char * slam_odom_version(void);
/* Thread-local text of the last error. */
const char * slam_odom_last_error(void);

/* RGBDOdometryef::RGBDOdometryef / ~RGBDOdometryef   (RGBDOdometryef.cpp:21-116) */
int slam_odom_create(const slam_odom_params * params, slam_odom_t * out);
int slam_odom_destroy(slam_odom_t h);

/* initICP(GPUTexture* filteredDepth, depthCutoff)     (RGBDOdometryef.cpp:118-142)
 * d_depth: uint16 millimetres, 0 = invalid; pitch_bytes 0 => dense. */
int slam_odom_init_icp_depth(slam_odom_t h, const uint16_t * d_depth, size_t pitch_bytes, float depth_cutoff);
/* initICP(GPUTexture* predictedVertices, predictedNormals, depthCutoff)  (:144-167) */
int slam_odom_init_icp_maps(slam_odom_t h, const float * d_vertices4, const float * d_normals4, float depth_cutoff);
/* initICPModel(vertices, normals, depthCutoff, modelPose)   (:169-206); pose row-major 4x4 (host) */
int slam_odom_init_icp_model(slam_odom_t h, const float * d_vertices4, const float * d_normals4, float depth_cutoff,
                             const float * model_pose16);
/* initRGB / initRGBModel / initFirstRGB (GPUTexture* rgb)  (:237-265); d_rgba: 4 x uint8 per pixel */
int slam_odom_init_rgb(slam_odom_t h, const uint8_t * d_rgba);
int slam_odom_init_rgb_model(slam_odom_t h, const uint8_t * d_rgba);
int slam_odom_init_first_rgb(slam_odom_t h, const uint8_t * d_rgba);

/* getIncrementalTransformation(trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3)  (:267-595)
 * trans[3*batch], rot[9*batch] (row-major) are HOST arrays, in-out: prior pose in, new pose out. */
int slam_odom_get_incremental_transformation(slam_odom_t h, float * trans, float * rot, int rgb_only, float icp_weight,
                                             int pyramid, int fast_odom, int so3);
/* Asynchronous form for pipelines: enqueue only; the pose is read back by _wait(). */
int slam_odom_get_incremental_transformation_async(slam_odom_t h, const float * trans, const float * rot, int rgb_only,
                                                   float icp_weight, int pyramid, int fast_odom, int so3);
/* Collects the pending track.  The init_* entry points collect it themselves when they are called first (a pipelined caller
 * may prepare frame N + 1 before this call: the image swap of frame N has then already happened); the pose is kept for this call. */
int slam_odom_wait(slam_odom_t h, float * trans, float * rot);

/* getCovariance()   (:597-600); out[36*batch] row-major */
int slam_odom_get_covariance(slam_odom_t h, double * out36);
/* public fields lastICPError ... lastb   (RGBDOdometryef.h:62-70); stats[batch] */
int slam_odom_get_stats(slam_odom_t h, slam_odom_stats * stats);

/* ---- host-buffer front end (end-to-end path: inputs in pinned host memory) ------------
 * One frame of frame-to-model tracking in the reference's call order
 * (apps/elastic_fusion_file.cpp:366-374): initICPModel -> initRGBModel -> initICP(depth)
 * -> initRGB -> getIncrementalTransformation.  `prefetch` starts the H2D copies of a frame
 * into the idle staging slot on a copy stream, so they overlap the previous frame's solve;
 * `track_host` uses a prefetched frame if `frame` matches, else copies it itself. */
typedef struct slam_frame_host
{
    const uint16_t * depth;       /* current frame, uint16 mm            [batch][H][W]    */
    const uint8_t * rgba;         /* current frame RGBA8                 [batch][H][W][4] */
    const float * model_vertices4;/* model prediction, camera frame      [batch][H][W][4] */
    const float * model_normals4; /*                                     [batch][H][W][4] */
    const uint8_t * model_rgba;   /* model prediction colour             [batch][H][W][4] */
    const float * model_pose16;   /* pose of the model prediction        [batch][16]      */
    float depth_cutoff;           /* initICP depthCutoff                                    */
    float model_depth_cutoff;     /* initICPModel depthCutoff                               */
} slam_frame_host;

int slam_odom_prefetch_host(slam_odom_t h, const slam_frame_host * frame);
int slam_odom_track_host(slam_odom_t h, const slam_frame_host * frame, float * trans, float * rot, int rgb_only,
                         float icp_weight, int pyramid, int fast_odom, int so3);
/* Same, and the copies of the NEXT frame (may be NULL) are issued right after this frame's kernels, before waiting for them:
 * the host time of staging overlaps the device time of tracking (a pipelined reader loop in one call). */
int slam_odom_track_host_next(slam_odom_t h, const slam_frame_host * f, const slam_frame_host * next, float * trans, float * rot, int rgb_only,
                              float icp_weight, int pyramid, int fast_odom, int so3);
/* The reference's own data flow: the sensor frame (depth, rgba) arrives in HOST memory (apps/elastic_fusion_file.cpp:301-340
 * uploads it every frame), while the model prediction never leaves the device (it is rendered into GL textures,
 * :366-374) -- here model_vertices4 / model_normals4 / model_rgba are DEVICE pointers and depth / rgba are host pointers
 * (pinned for full PCIe speed).  1.8 MB cross PCIe per 640x480 frame instead of the 12.9 MB of slam_odom_track_host.
 * `next` (may be NULL): the following frame, whose depth / rgba copies are issued behind this frame's kernels. */
int slam_odom_track_sensor(slam_odom_t h, const slam_frame_host * f, const slam_frame_host * next, float * trans, float * rot, int rgb_only,
                           float icp_weight, int pyramid, int fast_odom, int so3);
/* Same sequence with all inputs already in device memory (one call per frame). */
int slam_odom_track_device(slam_odom_t h, const slam_frame_host * frame_dev, float * trans, float * rot, int rgb_only,
                           float icp_weight, int pyramid, int fast_odom, int so3);

/* ---- debug taps: copy an internal buffer of sequence `seq` to host (parity tests) ---- */
enum slam_tap
{
    SLAM_TAP_DEPTH_U16 = 0,     /* depth_tmp[level]        uint16 [h][w]                    */
    SLAM_TAP_VMAP_CURR = 1,     /* vmaps_curr_[level]      float  [3][h][w] (planar)        */
    SLAM_TAP_NMAP_CURR = 2,
    SLAM_TAP_VMAP_PREV = 3,     /* vmaps_g_prev_[level]                                     */
    SLAM_TAP_NMAP_PREV = 4,
    SLAM_TAP_LAST_DEPTH = 5,    /* lastDepth[level]        float  [h][w]                    */
    SLAM_TAP_NEXT_DEPTH = 6,
    SLAM_TAP_LAST_IMAGE = 7,    /* lastImage[level]        uint8  [h][w]                    */
    SLAM_TAP_NEXT_IMAGE = 8,
    SLAM_TAP_LASTNEXT_IMAGE = 9,
    SLAM_TAP_DIDX = 10,         /* nextdIdx[level]         int16  [h][w]                    */
    SLAM_TAP_DIDY = 11,
    SLAM_TAP_CLOUD = 12,        /* pointClouds[level]      float  [h][w][3]                 */
    SLAM_TAP_CORRES = 13        /* corresImg[level]        16-byte DataTerm [h][w]          */
};
size_t slam_odom_tap_bytes(slam_odom_t h, int tap, int level);
int slam_odom_tap(slam_odom_t h, int tap, int level, int seq, void * host_dst, size_t bytes);

/* Record of one Gauss-Newton step (host_loop = 1 or trace enabled), for step-by-step parity. */
typedef struct slam_step_record
{
    int kind;                   /* 0 = SO3 step, 1 = ICP/RGB step */
    int level, iteration;
    float so3[11];              /* JtJJtrSO3 sums (types.cuh:138-165 order) */
    float icp[29];              /* JtJJtrSE3 sums of icpStep (types.cuh:79-136 order) */
    float rgb[29];              /* JtJJtrSE3 sums of rgbStep */
    int rgb_count, rgb_sigma;   /* computeRgbResidual outputs */
    double x[6];                /* solved increment */
    float Rcurr[9], tcurr[3];   /* pose after the step */
    /* inputs of the step, exactly as handed to the kernels (for teacher-forced replays in the tests) */
    float Rcurr_in[9], tcurr_in[3];   /* icpStep: Rcurr, tcurr */
    float krkinv_in[9], kt_in[3];     /* computeRgbResidual: krkinv, kt */
    float sigma_in;                   /* rgbStep: sigma */
    float so3_in[27];                 /* kind 0: so3Step imageBasis, kinv, krlr; kind 1: [0..8] = icpStep Rprev_inv */
    /* device-resident loop only: SM-clock timestamps (cycles since the kernel started, CTA 0) at
     * [0] step begin, [1] parameters ready, [2] phase A mapped, [3] phase A published + barrier passed,
     * [4] sigma known, [5] phase B mapped, [6] phase B published + barrier + fold done, [7] solve done */
    unsigned int t_cycles[8];
    /* sub-stages of the solve: [0] systems combined, [1] elimination done, [2] rotation increment, [3] resultRt updated,
     * [4] next parameters ready (cycles since the kernel started) */
    unsigned int t_solve[8];
} slam_step_record;
/* enable: 0 = off, 1 = step records only, 2 = records + the full DataTerm image of every RGB residual
 * pass (what the reference writes; needed by SLAM_TAP_CORRES) */
int slam_odom_set_trace(slam_odom_t h, int enable);
int slam_odom_get_trace(slam_odom_t h, int seq, slam_step_record * out, int max_records, int * n_records);

/* initICP on a RAW depth frame: the reference app's pre-filter (slam_op_depth_bilateral with max depth filter_max_depth_m,
 * apps/elastic_fusion_file.cpp:342-346) followed by initICP(filtered, depth_cutoff) (:368), without leaving the device. */
int slam_odom_init_icp_depth_raw(slam_odom_t h, const uint16_t * d_raw_depth, float filter_max_depth_m, float depth_cutoff);

/* Relocalisation scoring (BASELINE.json configs[4]; the acceptance test of lc/Ferns.cpp:253-268 reads the same two numbers
 * after a full track): score n candidate poses of the CURRENT frame (row-major rot9, trans3 each) against the model
 * prediction prepared by init_icp_model + init_icp_depth/maps, at pyramid level `level`, in one launch.  prev_trans3 /
 * prev_rot9 = the pose the model maps were predicted at.  Per hypothesis: residual[i] = sum of squared point-to-plane
 * distances over the inliers of icpStep's association (JtJJtrSE3::residual, cuda/types.cuh:79-92), count[i] = inliers;
 * the reference's lastICPError is sqrt(residual) / count.  Host output arrays; synchronises the handle's stream. */
int slam_odom_score_poses(slam_odom_t h, int seq, int level, int n, const float * prev_trans3, const float * prev_rot9, const float * trans3n,
                          const float * rot9n, float * residual_n, float * count_n);

/* The same scoring without leaving the device: the n hypotheses (global indices index_base .. index_base + n - 1) are scored and
 * the launch folds (lastICPError bits << 32 | global index) of each into *d_best_key (device memory, uint64) with atomicMin;
 * lastICPError = sqrt(residual) / inliers in float32, +inf below min_inliers (the guard of lc/Ferns.cpp:275-279).  The caller sets
 * *d_best_key to UINT64_MAX (or INT64_MAX) beforehand and, with the hypotheses sharded over several GPUs, min-all-reduces that one
 * word (ncclMin): smallest error wins, ties go to the smaller index, whatever the number of GPUs.  Only enqueues (handle's stream). */
int slam_odom_score_poses_best(slam_odom_t h, int seq, int level, int n, int index_base, float min_inliers, const float * prev_trans3,
                               const float * prev_rot9, const float * trans3n, const float * rot9n, unsigned long long * d_best_key);

/* The sharded form of that decision over NVLink peer memory, one process per GPU on one node (no counterpart in the reference, which
 * scores one candidate at a time, lc/Ferns.cpp:253-279).  Set-up, once: every rank calls slam_odom_peer_export (64 opaque bytes: the CUDA
 * IPC handle of its slot array), the ranks exchange the handles by any means (bench.py: torch.distributed.all_gather_object) and call
 * slam_odom_peer_connect with all `world` handles in rank order (world <= 16).  Per frame every rank calls
 * slam_odom_score_poses_best_peers with ITS block of the hypotheses (n may be 0): the block is scored as by slam_odom_score_poses_best,
 * a one-warp launch writes the rank's key into its slot on every peer and takes the minimum over the slots the peers wrote here, and
 * the call returns the winning key of ALL ranks in *best_key (host memory): (lastICPError bits << 32 | global index), INT64_MAX if no
 * hypothesis on any rank was acceptable.  Collective: every connected rank has to make the call for every frame. */
int slam_odom_peer_export(slam_odom_t h, void * handle64);
int slam_odom_peer_connect(slam_odom_t h, int rank, int world, const void * handles64);
int slam_odom_score_poses_best_peers(slam_odom_t h, int seq, int level, int n, int index_base, float min_inliers, const float * prev_trans3,
                                     const float * prev_rot9, const float * trans3n, const float * rot9n, unsigned long long * best_key);

/* Kernel-launch counter (for bench.py's gpu_launches). */
long long slam_odom_launch_count(slam_odom_t h);
/* CUDA-event timing of the Gauss-Newton reduction kernels on the handle's stream: enable, run, then read the
 * accumulated device time and the number of timed brackets (reset = 1 clears the totals).  batch < 4: one bracket per
 * frame around the persistent kernel; batch >= 4 (streaming engine): one bracket per pyramid level around that level's
 * phase A / phase B launches. */
int slam_odom_set_profiling(slam_odom_t h, int enable);
int slam_odom_get_profile(slam_odom_t h, double * gn_kernel_ms, long long * gn_kernel_launches, int reset);
/* SM cycles the leading CTA of the persistent kernel spent per phase, summed over the launches since the handle was created
 * (or since the last reset): [0] staging, [1] rest of the SO3 loop, [2] step set-up, [3] RGB association + count post, [4] ICP
 * products, [5] their block reduction + post + wait for the global count, [6] RGB products + block reduction + post, [7] wait for
 * all sums, [8] solve, [9] end-of-step barrier, [10] tail, [11] SO3 map, [12] SO3 block reduction + post, [13] SO3 wait for the
 * sums, [14] SO3 update, [15] launches, [16..19] staging of level 0..3, [20] staging of the SO3 images ([0] then holds only the check-in gate), [21] split launch: wait for the cluster's hand-off.  Only the variants launched with SLAM_GN_PHASES=1 (or the step trace) count.
 * Synchronises the handle's stream. */
int slam_odom_get_phase_cycles(slam_odom_t h, unsigned long long out24[24], int reset);
/* Launch shape of the device-resident loop for one sequence (getIncrementalTransformation, RGBDOdometryef.cpp:267-595).  enable = 1
 * (the default; SLAM_GN_SPLIT=0 in the environment turns it off for the process): the SO3 pre-alignment runs on one thread-block
 * cluster of 16 SMs (all-reduce through distributed shared memory) next to a second kernel on the other SMs, which stages the pyramid
 * levels meanwhile, takes the rotation over and runs the ICP+RGB iterations; 0: one cooperative launch for the whole frame.  Calls that
 * do not qualify (no SO3 step, several sequences, rgb_only, step trace, levels that do not fit shared memory) always use one launch.
 * Returns the previous setting through *previous (may be NULL). */
int slam_odom_set_split_launch(slam_odom_t h, int enable, int * previous);
/* The stream the handle runs on (cudaStream_t), e.g. to record the caller's own events on it. */
void * slam_odom_stream(slam_odom_t h);

/* ---- operator-level API: the free host wrappers of src/odom/utils.cuh:62-175 ----------
 * Raw dense device pointers; `stream` is a cudaStream_t (NULL = default stream).  Each call
 * only enqueues work.  Planar maps are float [3][rows][cols]. */
/* Depth pre-filter that produces DEPTH_FILTERED, the input of initICP: 13x13 bilateral, values < 300 mm or > max_depth_m -> 0
 * (gl/shaders/depth_bilateral.frag:30-76, run by gl/ComputePack.cpp:41-73 from apps/elastic_fusion_file.cpp:342-346).
 * n_images dense u16 images back to back; src != dst. */
int slam_op_depth_bilateral(const uint16_t * src, int rows, int cols, float max_depth_m, uint16_t * dst, int n_images, void * stream);
int slam_op_pyr_down(const uint16_t * src, int src_rows, int src_cols, uint16_t * dst, void * stream);                   /* pyrDown            utils.cuh:155 */
int slam_op_create_vmap(float fx, float fy, float cx, float cy, const uint16_t * depth, int rows, int cols, float * vmap,
                        float depth_cutoff, void * stream);                                                               /* createVMap         :119 */
int slam_op_create_nmap(const float * vmap, int rows, int cols, float * nmap, void * stream);                            /* createNMap         :124 */
int slam_op_transform_maps(const float * vmap_src, const float * nmap_src, int rows, int cols, const float * R9,
                           const float * t3, float * vmap_dst, float * nmap_dst, void * stream);                          /* tranformMaps       :127 */
int slam_op_copy_maps(const float * vertices4, const float * normals4, int rows, int cols, float * vmap_dst, float * nmap_dst,
                      void * stream);                                                                                     /* copyMaps           :134 */
int slam_op_resize_vmap(const float * src, int src_rows, int src_cols, float * dst, void * stream);                      /* resizeVMap         :139 */
int slam_op_resize_nmap(const float * src, int src_rows, int src_cols, float * dst, void * stream);                      /* resizeNMap         :142 */
int slam_op_image_bgr_to_intensity(const uint8_t * rgba, int rows, int cols, uint8_t * dst, void * stream);              /* imageBGRToIntensity:145 */
int slam_op_vertices_to_depth(const float * vertices4, int rows, int cols, float * dst, float cutoff, void * stream);    /* verticesToDepth    :148 */
int slam_op_project_to_point_cloud(const float * depth, int rows, int cols, float * cloud3, float fx, float fy, float cx,
                                   float cy, int level, void * stream);                                                   /* projectToPointCloud:152 */
int slam_op_pyr_down_gauss_f(const float * src, int src_rows, int src_cols, float * dst, void * stream);                 /* pyrDownGaussF      :158 */
int slam_op_pyr_down_uchar_gauss(const uint8_t * src, int src_rows, int src_cols, uint8_t * dst, void * stream);         /* pyrDownUcharGauss  :161 */
int slam_op_compute_derivative_images(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy, void * stream); /* computeDerivativeImages :164 */

/* The four reductions.  `workspace` is device scratch of slam_op_workspace_bytes() bytes (zeroed once
 * by the caller before first use); results land in device memory `out` (29 / 11 floats, 2 ints) -- a
 * single launch each, no second-stage kernel, no sync, no host copy. */
size_t slam_op_workspace_bytes(void);
int slam_op_icp_step(const float * Rcurr9, const float * tcurr3, const float * vmap_curr, const float * nmap_curr,
                     const float * Rprev_inv9, const float * tprev3, float fx, float fy, float cx, float cy,
                     const float * vmap_g_prev, const float * nmap_g_prev, float dist_thresh, float angle_thresh, int rows,
                     int cols, void * workspace, float * out29, void * stream);                                           /* icpStep            :62 */
int slam_op_compute_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, const float * last_depth,
                                 const float * next_depth, const uint8_t * last_image, const uint8_t * next_image,
                                 void * corres_img16, float max_depth_delta, const float * kt3, const float * krkinv9,
                                 int rows, int cols, void * workspace, int * out_count_sigma2, void * stream);           /* computeRgbResidual :102 */
int slam_op_rgb_step(const void * corres_img16, float sigma, const float * cloud3, float fx, float fy, const int16_t * dIdx,
                     const int16_t * dIdy, float sobel_scale, int rows, int cols, void * workspace, float * out29,
                     void * stream);                                                                                      /* rgbStep            :81 */
int slam_op_so3_step(const uint8_t * last_image, const uint8_t * next_image, const float * image_basis9, const float * kinv9,
                     const float * krlr9, int rows, int cols, void * workspace, float * out11, void * stream);           /* so3Step            :89 */

#ifdef __cplusplus
}
#endif
#endif /* SLAM_ODOM_H_ */
