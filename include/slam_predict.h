/*
 * slam_predict.h -- C ABI of the model-prediction producer: the step on the other side of
 * the tracker's boundary, which turns the surfel map into the float4 vertex / normal and RGBA8 colour images that
 * initICPModel / initRGBModel consume.
 *
 * Replaces, for this path, two GL classes of the reference (neither has an FFI layer):
 *     IndexMap::combinedPredict(pose, model, depthCutoff, confThreshold, time, maxTime, timeDelta, type)
 *                                                  src/model/IndexMap.cpp:243-341
 *         vertex stage    src/model/shaders/splat.vert:18-87        (cull, transform, point-sprite size)
 *         fragment stage  src/model/shaders/combo_splat.frag:18-61  (ray / surfel-disc intersection, depth, outputs)
 *         textures        imageTex RGBA8, vertexTex RGBA32F, normalTex RGBA32F, timeTex R16UI (IndexMap.cpp:57-75)
 *     FillIn::vertex / ::normal / ::image(existing, raw, passthrough)       src/gl/FillIn.cpp:68-198
 *         shaders         src/gl/shaders/fill_vertex.frag, fill_normal.frag, fill_rgb.frag, geometry.glsl:43-61
 * as driven by predict() in src/apps/elastic_fusion_file.cpp:17-44 (and Model::performFillIn, src/model/Model.cpp:645-650).
 *
 * B200-native shape: the surfel buffer (the reference's VBO, 3 x vec4 per surfel, src/gl/Vertex.cpp) and all textures are
 * linear HBM buffers.  A frame's prediction is two launches: a splat launch (one warp per 32 surfels, the fragments of
 * the 32 point sprites flattened over the lanes, GL's depth test as a packed (depth24, draw order) 64-bit atomicMin) and
 * a resolve launch (one thread per pixel: winner's outputs, the three fill-in passes fused behind it, z-buffer re-armed).
 * No GL context, no glFinish, no texture -> linear copy.  No CPU fallback: without a CUDA device slam_predict_create fails.
 *
 * GL leaves point rasterisation partly implementation defined; the rules used here are stated at
 * slam_b200/csrc/predict.cu (top); the tests' CPU restatement follows the same list.
 */
#ifndef SLAM_PREDICT_H_
#define SLAM_PREDICT_H_

#include <stddef.h>
#include <stdint.h>
#include "slam_odom.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slam_predict * slam_predict_t;

#define SLAM_SURFEL_FLOATS 12   /* Vertex::SIZE / 4: position+confidence, colour+times, normal+radius (gl/Vertex.cpp) */

typedef struct slam_predict_params
{
    int width, height;          /* IndexMap(width, height, intr), FillIn(width, height, intr) */
    float cx, cy, fx, fy;       /* CameraModel */
    float max_point_size;       /* upper end of GL_POINT_SIZE_RANGE; 0 => 2047 */
    int device;
    void * stream;              /* cudaStream_t; NULL => the library creates one.  A supplied stream stays the caller's: it must
                                   outlive the handle's last launch; slam_predict_destroy neither synchronises nor destroys it */
} slam_predict_params;

/* Device pointers of the handle's textures (dense row-major, width x height texels). */
typedef struct slam_predict_textures
{
    const uint8_t * image;        /* IndexMap::imageTex()   RGBA8                      */
    const float * vertex;         /* IndexMap::vertexTex()  RGBA32F (x, y, z, conf)    */
    const float * normal;         /* IndexMap::normalTex()  RGBA32F (nx, ny, nz, rad)  */
    const uint16_t * time;        /* IndexMap::timeTex()    R16UI                      */
    const uint8_t * fill_image;   /* FillIn::imageTexture   RGBA8                      */
    const float * fill_vertex;    /* FillIn::vertexTexture  RGBA32F                    */
    const float * fill_normal;    /* FillIn::normalTexture  RGBA32F                    */
    const uint8_t * old_image;    /* IndexMap::oldImageTex()   (INACTIVE prediction)   */
    const float * old_vertex;     /* IndexMap::oldVertexTex()                          */
    const float * old_normal;     /* IndexMap::oldNormalTex()                          */
    const uint16_t * old_time;    /* IndexMap::oldTimeTex()                            */
} slam_predict_textures;

/* IndexMap::Prediction (model/IndexMap.h): which frame buffer combinedPredict renders into. */
#define SLAM_PREDICT_ACTIVE 0
#define SLAM_PREDICT_INACTIVE 1

int slam_predict_create(const slam_predict_params * params, slam_predict_t * out);
int slam_predict_destroy(slam_predict_t h);
int slam_predict_get_textures(slam_predict_t h, slam_predict_textures * out);

/* IndexMap::combinedPredict (ACTIVE).  d_surfels: count x 12 floats, the model VBO; pose16: row-major 4x4 (host). */
int slam_predict_combined(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold,
                          int time, int max_time, int time_delta);

/* The same with the reference's last argument: SLAM_PREDICT_INACTIVE renders the old part of the map (apps/elastic_fusion_file.cpp:450-457
 * calls it with time 0, maxTime tick - timeDelta) into the old* textures, which modelToModel.initICPModel / initRGBModel read
 * (apps/elastic_fusion_file.cpp:460-463); the ACTIVE textures are left untouched. */
int slam_predict_combined_type(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold,
                               int time, int max_time, int time_delta, int prediction_type);

/* FillIn::vertex / ::normal / ::image.  `existing` NULL => the handle's own IndexMap texture of that kind. */
int slam_predict_fill_vertex(slam_predict_t h, const float * d_existing_vertex4, const uint16_t * d_raw_depth, int passthrough);
int slam_predict_fill_normal(slam_predict_t h, const float * d_existing_normal4, const uint16_t * d_raw_depth, int passthrough);
int slam_predict_fill_image(slam_predict_t h, const uint8_t * d_existing_rgba, const uint8_t * d_raw_rgba, int passthrough);

/* predict() of apps/elastic_fusion_file.cpp:17-44 in one call: combinedPredict + the three FillIn passes (passthrough
 * false), two launches.  write_index_textures != 0 also writes the unfilled IndexMap textures (image, vertex, normal,
 * time); otherwise only the FillIn textures the tracker reads are produced. */
int slam_predict_frame(slam_predict_t h, const float * d_surfels, int count, const float * pose16, float depth_cutoff, float conf_threshold, int time,
                       int max_time, int time_delta, const uint16_t * d_raw_depth, const uint8_t * d_raw_rgba, int write_index_textures);

/* Copy one texture to the host (synchronises the handle's stream).  texture: index of the field in slam_predict_textures
 * (0 image, 1 vertex, 2 normal, 3 time, 4 fill_image, 5 fill_vertex, 6 fill_normal, 7 old_image, 8 old_vertex, 9 old_normal,
 * 10 old_time); host_out: width x height texels. */
int slam_predict_download(slam_predict_t h, int texture, void * host_out);

/* ---- parity taps / timing aids --------------------------------------------------------------------------------- */
/* t_inv = pose.inverse() as used by the last combined / frame call (row-major, host). */
int slam_predict_get_tinv(slam_predict_t h, float * tinv16);
/* the resolved z-buffer of the last combined / frame call: per pixel the 24-bit depth (0xFFFFFF = empty) and the index of
 * the winning surfel (-1 = empty); host arrays of width x height, either may be NULL.  Valid until the next call. */
int slam_predict_get_winners(slam_predict_t h, uint32_t * depth24, int32_t * surfel);
/* CUDA-event times of the last call's launches (ms) and the number of fragments its splat launch evaluated (GL's sprite
 * squares minus the parts outside the projected quad around each disc, which the fragment shader could only discard). */
int slam_predict_last_ms(slam_predict_t h, float * splat_ms, float * resolve_ms);
int slam_predict_last_fragments(slam_predict_t h, unsigned long long * fragments);

#ifdef __cplusplus
}
#endif
#endif
