// Internal declarations shared by the translation units of libslam_odom.
#pragma once
#include <vector>
#include <string>
#include "common.cuh"
#include "small_math.hpp"

namespace slam {

struct LevelGeom
{
    int rows, cols;
    float fx, fy, cx, cy;   // CameraModel::operator()(level), sensors/Camera.h:14-18
};

// Device buffers of ONE sequence (all dense row-major; planar maps are [3][rows][cols]).
struct SeqBuffers
{
    unsigned short * depth[SLAM_MAX_LEVELS];      // depth_tmp[i]          RGBDOdometryef.h:80
    float * vcurr[SLAM_MAX_LEVELS];               // vmaps_curr_
    float * ncurr[SLAM_MAX_LEVELS];               // nmaps_curr_
    float * vprev[SLAM_MAX_LEVELS];               // vmaps_g_prev_
    float * nprev[SLAM_MAX_LEVELS];               // nmaps_g_prev_
    float * lastDepth[SLAM_MAX_LEVELS];
    float * nextDepth[SLAM_MAX_LEVELS];
    unsigned char * lastImage[SLAM_MAX_LEVELS];
    unsigned char * nextImage[SLAM_MAX_LEVELS];       // swapped with lastNextImage after an so3 call
    unsigned char * lastNextImage[SLAM_MAX_LEVELS];
    short * dIdx[SLAM_MAX_LEVELS];
    short * dIdy[SLAM_MAX_LEVELS];
    Corres * corres[SLAM_MAX_LEVELS];             // corresImg
    float * depth_tmp;                            // z of vmaps_tmp after the maxDepthRGB cut
    float * vcam;                                 // camera-frame level-2 maps (only when levels == 4)
    float * ncam;
    void * workspace;                             // reduction scratch
    float * sums;                                 // device: icp[32] rgb[32] so3[16] res(int)[2..]
};

// ---- device-resident Gauss-Newton state (one per sequence), see gn_kernel.cu ----
struct GnResult
{
    float Rcurr[9];
    float tcurr[3];
    float lastICPError, lastICPCount;
    float lastRGBError, lastRGBCount;
    float lastSO3Error, lastSO3Count;
    double lastA[36];
    double lastb[6];
    int so3_iterations;
    int gn_iterations;
};

// launchers implemented in prep_kernels.cu / reduce_kernels.cu
// Arguments of the one-launch frame preparation (three-level pyramids), see k_prepare_frame in prep_kernels.cu.
struct PrepFrameArgs
{
    int rows, cols;
    int tiles_x, tiles, model_blocks;
    // model + RGB-D role
    const float4 * vsrc;
    const float4 * nsrc;
    const uchar4 * model_rgba;
    const uchar4 * rgba;
    float * vprev[3];
    float * nprev[3];
    float * depth_tmp;
    float * lastDepth[3];
    float * nextDepth[3];
    unsigned char * lastImage[3];
    unsigned char * nextImage[3];
    Mat3 R;
    float3 t;
    const float * poses12;   // batched: per-sequence [R | t]
    float depth_cut;
    // sensor depth role
    const unsigned short * depth;
    unsigned short * depth_l[3];   // [1], [2]: the coarser depth levels (outputs)
    float * vcurr[3];
    float * ncurr[3];
    float fx_inv[3], fy_inv[3], cx[3], cy[3];
    float depthCutoff;
    // batched launch (gridDim.y sequences): byte strides of the caller's input stacks and of the arena buffers
    size_t map_in_stride, rgba_stride, depth_in_stride, arena_stride;
    unsigned long long * dbg;   // development aid (SLAM_PREP_DEBUG): per block {start, end (%globaltimer), SM}
};
int launch_prepare_frame(PrepFrameArgs & a, cudaStream_t s, int nseq);

struct ModelMapsArgs;
struct DerivArgs;
int launch_depth_level(const unsigned short * depth, int rows, int cols, float fx, float fy, float cx, float cy, float depthCutoff, float * vmap,
                       float * nmap, unsigned short * next_depth, cudaStream_t s, int nseq = 1, size_t in_stride = 0, size_t out_stride = 0);
int launch_rgbd_level0(const float * depth_tmp, float * depth0, const uchar4 * rgba, unsigned char * image0, int n, cudaStream_t s);
int launch_rgbd_level0_dual(const float * depth_tmp, float * lastDepth0, float * nextDepth0, const uchar4 * model_rgba, unsigned char * lastImage0,
                            const uchar4 * rgba, unsigned char * nextImage0, int n, cudaStream_t s, int nseq = 1, size_t in_stride = 0, size_t arena_stride = 0);
int launch_rgbd_down_dual(const float * dsrc, float * ddstLast, float * ddstNext, const unsigned char * isrcLast, unsigned char * idstLast,
                          const unsigned char * isrcNext, unsigned char * idstNext, int srows, int scols, cudaStream_t s, int nseq = 1,
                          size_t arena_stride = 0);
int launch_rgbd_down(const float * dsrc, float * ddst, const unsigned char * isrc, unsigned char * idst, int srows, int scols, cudaStream_t s);
int launch_resize_transform(const float * vsrc, const float * nsrc, int srows, int scols, float * vdst, float * ndst, int transform, const Mat3 & R,
                            const float3 & t, float * vcam, float * ncam, cudaStream_t s);
int launch_icp_step(const IcpArgs & a, void * workspace, float * out29, cudaStream_t s);
int launch_rgb_residual(const ResidualArgs & a, Corres * corres, void * workspace, int * out2, cudaStream_t s);
int launch_rgb_step(const RgbStepArgs & a, const Corres * corres, void * workspace, float * out29, cudaStream_t s);
int launch_so3_step(const So3Args & a, void * workspace, float * out11, cudaStream_t s);
size_t score_workspace_bytes(int n, int plane);
int launch_score_poses(const IcpArgs & a, const float * poses12, int n, void * workspace /* zero-initialised once */, float ** out2_dev, cudaStream_t s,
                       unsigned long long * best_key = nullptr, int index_base = 0, float min_inliers = 1.f);

// helpers exported by prep_kernels.cu that need the full argument structs
int launch_model_maps_simple(const float4 * vsrc, const float4 * nsrc, int rows, int cols, int levels, float * const * vdst, float * const * ndst,
                             int transform, const Mat3 & R, const float3 & t, float * depth_tmp, float depth_cut, float * vcam2, float * ncam2,
                             cudaStream_t s, int nseq = 1, size_t in_stride = 0, size_t out_stride = 0, const float * poses12 = nullptr, float * lastDepth0 = nullptr,
                             float * nextDepth0 = nullptr, const uchar4 * model_rgba = nullptr, const uchar4 * rgba = nullptr, unsigned char * lastImage0 = nullptr,
                             unsigned char * nextImage0 = nullptr, size_t rgba_stride = 0);
int launch_derivatives_simple(int levels, const unsigned char * const * src, short * const * dx, short * const * dy, const int * rows, const int * cols,
                              cudaStream_t s, int nseq = 1, size_t arena_stride = 0);

int launch_depth_bilateral(const unsigned short * src, int rows, int cols, float max_depth_m, unsigned short * dst, int n_images, cudaStream_t s);
// derivative images + RGB candidate masks of all levels and sequences in one launch (cols % 4 == 0 at every level)
int launch_deriv_cand(int levels, const unsigned char * const * src, const float * const * depth, short * const * dx, short * const * dy,
                      unsigned char * const * cand, const float * min_scale, const int * rows, const int * cols, cudaStream_t s, int nseq, size_t arena_stride,
                      size_t cand_stride);

}   // namespace slam
