/*
 * slam_ferns.h -- C ABI of the randomised-fern relocalisation / loop-closure front end
 * (SURVEY.md section 8f row 2), the direct consumer of the tracker's ICP path.
 *
 * Replaces the reference's C++ class `Ferns` (src/lc/Ferns.h:28-181, src/lc/Ferns.cpp), which has no FFI layer:
 *     Ferns(n, maxDepth, photoThresh, intr, w, h, shader_dir)          Ferns.cpp:21-57
 *     bool addFrame(image, vertex, normal, pose, srcTime, threshold)   Ferns.cpp:83-168
 *     Matrix4f findFrame(constraints, currPose, vertex, normal, image, time, lost)  Ferns.cpp:170-307
 *     (private) blockHDAware Ferns.cpp:374-389, photometricCheck Ferns.cpp:309-357
 * and the GL resize pass it calls first (src/gl/Resize.cpp:70-154: a NEAREST texture fetch at the centre of every
 * 8x8 block, i.e. source texel (8x+4, 8y+4)).
 *
 * B200-native layout: the key-frame database (codes, 80x60 colour / vertex / normal images, poses) lives in HBM;
 * encoding is one launch, the search over the whole database is one streaming launch (one warp per key frame,
 * 512 B per key frame, packed-key arg-min in the same launch), the photometric check one block.  The ICP refinement
 * is a slam_odom handle at 1/8 resolution (RGBDOdometryef rgbd(w/8, h/8, ...), Ferns.cpp:34-39) driven exactly as
 * Ferns.cpp:253-268 does.  No CPU fallback: without a CUDA device slam_ferns_create fails.
 *
 * Inputs are linear device pointers in the texel formats of the reference's textures: RGBA8 (the RGB upload of
 * gl/FillIn.cpp), RGBA32F vertex (x, y, z, conf) and normal (nx, ny, nz, radius) maps in the camera frame.
 */
#ifndef SLAM_FERNS_H_
#define SLAM_FERNS_H_

#include <stddef.h>
#include <stdint.h>
#include "slam_odom.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slam_ferns * slam_ferns_t;

/* Ferns::Fern (Ferns.h:62-72) without the inverted lists: position in the 1/8 image and the four thresholds. */
typedef struct slam_fern
{
    int32_t x, y;        /* pos(0) in [0, width/8), pos(1) in [0, height/8)        */
    int32_t r, g, b;     /* rgbd(0..2) in [0, 255]                                   */
    int32_t d;           /* rgbd(3) in [400, maxDepth] millimetres                   */
} slam_fern;

typedef struct slam_ferns_params
{
    int width, height;           /* full resolution (640 x 480); the fern images are width/8 x height/8 */
    float cx, cy, fx, fy;        /* CameraModel of the full-resolution camera                          */
    int num_ferns;               /* 0 => 500 (apps/elastic_fusion_file.cpp:271)                         */
    int max_depth_mm;            /* depthCutoff * 1000                                                  */
    float photo_thresh;          /* photoThresh                                                         */
    int capacity;                /* key frames the database can hold; 0 => 1024                         */
    uint32_t seed;               /* std::mt19937 seed of generateFerns (the reference uses time(0))     */
    int device;
} slam_ferns_params;

/* Ferns::SurfaceConstraint (Ferns.h:39-50): homogeneous source / target points. */
typedef struct slam_surface_constraint
{
    float source[4];
    float target[4];
} slam_surface_constraint;

/* What findFrame leaves behind besides its return value (lastClosest, the fields its acceptance test reads). */
typedef struct slam_ferns_match
{
    int min_id;                  /* closest key frame by code dissimilarity (-1: none older than 300 ticks) */
    float dissimilarity;         /* its dissimilarity                                                    */
    float block_hd_aware;        /* blockHDAware(query, frames[min_id])                                  */
    int icp_ran;                 /* 1 when block_hd_aware > 0.3 and the ICP refinement ran               */
    float icp_error, icp_count;  /* rgbd.lastICPError, rgbd.lastICPCount                                 */
    float photo_error;           /* photometricCheck                                                     */
    int last_closest;            /* Ferns::lastClosest: min_id when accepted, else -1                     */
} slam_ferns_match;

/* Ferns::Ferns.  `table` (num_ferns entries, host) overrides the generated conservatory when not NULL. */
int slam_ferns_create(const slam_ferns_params * params, const slam_fern * table, slam_ferns_t * out);
int slam_ferns_destroy(slam_ferns_t h);
/* the conservatory (host copy, num_ferns entries) */
int slam_ferns_get_table(slam_ferns_t h, slam_fern * out);
int slam_ferns_num_frames(slam_ferns_t h);

/* Ferns::addFrame.  pose16: row-major 4x4 (host).  *added = the reference's return value. */
int slam_ferns_add_frame(slam_ferns_t h, const uint8_t * d_rgba, const float * d_vertices4, const float * d_normals4, const float * pose16,
                         int src_time, float threshold, int * added);

/* Ferns::findFrame.  est_pose16 receives the returned pose (identity when nothing matched); constraints may be NULL. */
int slam_ferns_find_frame(slam_ferns_t h, const float * curr_pose16, const float * d_vertices4, const float * d_normals4, const uint8_t * d_rgba,
                          int time, int lost, float * est_pose16, slam_ferns_match * match, slam_surface_constraint * constraints,
                          int max_constraints, int * n_constraints);

/* ---- operator-level entry points (parity taps) ------------------------------------------------------------- */
/* Resize + encode one frame into the query slot; optional host outputs: codes[num_ferns], small images
 * rgb[h/8][w/8][3], vert[h/8][w/8][4], norm[h/8][w/8][4]. */
int slam_ferns_encode(slam_ferns_t h, const uint8_t * d_rgba, const float * d_vertices4, const float * d_normals4, uint8_t * codes, int * good_codes,
                      uint8_t * rgb_small, float * vert_small, float * norm_small);
/* Dissimilarity of the encoded query against every key frame (host array, num_frames entries) and the arg-min
 * under findFrame's rule (time - srcTime > 300) when use_time != 0, addFrame's rule (all frames) otherwise. */
int slam_ferns_search(slam_ferns_t h, int time, int use_time, float * dissim, int * min_id, float * minimum, float * block_hd_aware);
/* photometricCheck of the encoded query against key frame `id` for the given poses (host, row-major 4x4). */
int slam_ferns_photometric_check(slam_ferns_t h, int id, const float * est_pose16, const float * fern_pose16, float * photo_error, int * photo_count);
/* key frame `id`: codes[num_ferns], pose16, srcTime (any may be NULL) */
int slam_ferns_get_frame(slam_ferns_t h, int id, uint8_t * codes, float * pose16, int * src_time, int * good_codes);
/* kernel timing aid: CUDA-event time of the last search launch (ms) */
int slam_ferns_last_search_ms(slam_ferns_t h, float * ms);

#ifdef __cplusplus
}
#endif
#endif
