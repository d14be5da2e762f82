// Header-only C++ classes with the reference's method names over the C ABI of the model-prediction producer
// (slam_predict.h) and of the fern relocaliser (slam_ferns.h) -- the producer and the caller on either side of the tracker.
//
//   IndexMap   the prediction half of src/model/IndexMap.h:28-215: combinedPredict(pose, model, depthCutoff, confThreshold, time,
//              maxTime, timeDelta, predictionType) and the imageTex / vertexTex / normalTex / timeTex / old*Tex accessors
//   FillIn     src/gl/FillIn.h: vertex / normal / image(existing, raw, passthrough) and the three result textures
//   Ferns      src/lc/Ferns.h:28-181: addFrame / findFrame / lastClosest
//
// `GPUTexture*` arguments become linear device pointers of the same texel format and the (vbo, count) pair of the model
// becomes (device pointer, count).  IndexMap and FillIn share one slam_predict handle, as the fused call
// (IndexMap::predictAndFill, = predict() of src/apps/elastic_fusion_file.cpp:17-44) needs both texture sets.  Poses are
// row-major 4x4 floats; with SLAM_ODOM_WITH_EIGEN defined Eigen::Matrix4f overloads with the reference's signatures exist.
// Errors throw std::runtime_error (the reference prints and exit(0)s).  There is no CPU fallback.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "slam_ferns.h"
#include "slam_predict.h"

#ifdef SLAM_ODOM_WITH_EIGEN
#include <Eigen/Core>
#endif

namespace slam_b200 {

inline void check(int rc)
{
    if(rc != SLAM_OK) throw std::runtime_error(std::string("libslam_odom: ") + slam_odom_last_error());
}

class FillIn;

class IndexMap
{
  public:
    enum Prediction { ACTIVE = SLAM_PREDICT_ACTIVE, INACTIVE = SLAM_PREDICT_INACTIVE };

    // IndexMap(width, height, intr, shader_dir) -- no shaders here
    IndexMap(int width, int height, float cx, float cy, float fx, float fy, int device = 0, void * cudaStream = nullptr) : handle_(nullptr)
    {
        slam_predict_params p = {};
        p.width = width; p.height = height;
        p.cx = cx; p.cy = cy; p.fx = fx; p.fy = fy;
        p.device = device;
        p.stream = cudaStream;
        check(slam_predict_create(&p, &handle_));
        check(slam_predict_get_textures(handle_, &tex_));
    }
    ~IndexMap() { slam_predict_destroy(handle_); }
    IndexMap(const IndexMap &) = delete;
    IndexMap & operator=(const IndexMap &) = delete;

    // model: (device pointer of the surfel buffer, number of surfels) -- the reference's std::pair<GLuint, GLuint> (vbo, count)
    void combinedPredict(const float * pose16, const std::pair<const float *, int> & model, const float depthCutoff, const float confThreshold,
                         const int time, const int maxTime, const int timeDelta, Prediction predictionType)
    {
        check(slam_predict_combined_type(handle_, model.first, model.second, pose16, depthCutoff, confThreshold, time, maxTime, timeDelta, (int)predictionType));
    }
    // predict(): combinedPredict(ACTIVE) + fillIn.vertex / normal / image(passthrough = false) in two launches
    void predictAndFill(const float * pose16, const std::pair<const float *, int> & model, const float depthCutoff, const float confThreshold, const int time,
                        const int maxTime, const int timeDelta, const uint16_t * rawDepth, const uint8_t * rawRgb, bool writeIndexTextures = true)
    {
        check(slam_predict_frame(handle_, model.first, model.second, pose16, depthCutoff, confThreshold, time, maxTime, timeDelta, rawDepth, rawRgb,
                                 writeIndexTextures ? 1 : 0));
    }
#ifdef SLAM_ODOM_WITH_EIGEN
    void combinedPredict(const Eigen::Matrix4f & pose, const std::pair<const float *, int> & model, const float depthCutoff, const float confThreshold,
                         const int time, const int maxTime, const int timeDelta, Prediction predictionType)
    {
        const Eigen::Matrix<float, 4, 4, Eigen::RowMajor> rm = pose;
        combinedPredict(rm.data(), model, depthCutoff, confThreshold, time, maxTime, timeDelta, predictionType);
    }
#endif

    const uint8_t * imageTex() const { return tex_.image; }
    const float * vertexTex() const { return tex_.vertex; }
    const float * normalTex() const { return tex_.normal; }
    const uint16_t * timeTex() const { return tex_.time; }
    const uint8_t * oldImageTex() const { return tex_.old_image; }
    const float * oldVertexTex() const { return tex_.old_vertex; }
    const float * oldNormalTex() const { return tex_.old_normal; }
    const uint16_t * oldTimeTex() const { return tex_.old_time; }

    slam_predict_t handle() const { return handle_; }
    const slam_predict_textures & textures() const { return tex_; }

  private:
    slam_predict_t handle_;
    slam_predict_textures tex_;
};

// FillIn works on the textures of the IndexMap it is constructed from (it does not own the handle).
class FillIn
{
  public:
    explicit FillIn(IndexMap & indexMap) : handle_(indexMap.handle()), tex_(indexMap.textures()) {}

    void vertex(const float * existingVertex, const uint16_t * rawDepth, bool passthrough) { check(slam_predict_fill_vertex(handle_, existingVertex, rawDepth, passthrough)); }
    void normal(const float * existingNormal, const uint16_t * rawDepth, bool passthrough) { check(slam_predict_fill_normal(handle_, existingNormal, rawDepth, passthrough)); }
    void image(const uint8_t * existingRgb, const uint8_t * rawRgb, bool passthrough) { check(slam_predict_fill_image(handle_, existingRgb, rawRgb, passthrough)); }

    const uint8_t * imageTexture() const { return tex_.fill_image; }
    const float * vertexTexture() const { return tex_.fill_vertex; }
    const float * normalTexture() const { return tex_.fill_normal; }

  private:
    slam_predict_t handle_;
    slam_predict_textures tex_;
};

class Ferns
{
  public:
    struct SurfaceConstraint
    {
        float sourcePoint[4];
        float targetPoint[4];
    };

    // Ferns(n, maxDepth, photoThresh, intr, w, h, shader_dir); `seed` replaces the reference's time(0)
    Ferns(int n, int maxDepth, float photoThresh, float cx, float cy, float fx, float fy, int width, int height, unsigned seed = 0, int device = 0)
     : lastClosest(-1), num_(n), handle_(nullptr)
    {
        slam_ferns_params p = {};
        p.width = width; p.height = height;
        p.cx = cx; p.cy = cy; p.fx = fx; p.fy = fy;
        p.num_ferns = n;
        p.max_depth_mm = maxDepth;
        p.photo_thresh = photoThresh;
        p.seed = seed;
        p.device = device;
        check(slam_ferns_create(&p, nullptr, &handle_));
    }
    ~Ferns() { slam_ferns_destroy(handle_); }
    Ferns(const Ferns &) = delete;
    Ferns & operator=(const Ferns &) = delete;

    bool addFrame(const uint8_t * imageTexture, const float * vertexTexture, const float * normalTexture, const float * pose16, int srcTime, float threshold)
    {
        int added = 0;
        check(slam_ferns_add_frame(handle_, imageTexture, vertexTexture, normalTexture, pose16, srcTime, threshold, &added));
        return added != 0;
    }

    // returns the estimated pose in estPose16 (identity when nothing matched), appends to constraints, sets lastClosest
    void findFrame(std::vector<SurfaceConstraint> & constraints, const float * currPose16, const float * vertexTexture, const float * normalTexture,
                   const uint8_t * imageTexture, int time, bool lost, float * estPose16)
    {
        std::vector<slam_surface_constraint> buf(num_);
        int n = 0;
        slam_ferns_match m;
        check(slam_ferns_find_frame(handle_, currPose16, vertexTexture, normalTexture, imageTexture, time, lost ? 1 : 0, estPose16, &m, buf.data(), num_, &n));
        for(int k = 0; k < n; k++)
        {
            SurfaceConstraint c;
            for(int j = 0; j < 4; j++)
            {
                c.sourcePoint[j] = buf[k].source[j];
                c.targetPoint[j] = buf[k].target[j];
            }
            constraints.push_back(c);
        }
        lastClosest = m.last_closest;
        lastMatch = m;
    }

    int numFrames() const { return slam_ferns_num_frames(handle_); }

    int lastClosest;
    slam_ferns_match lastMatch;

  private:
    int num_;
    slam_ferns_t handle_;
};

}   // namespace slam_b200
