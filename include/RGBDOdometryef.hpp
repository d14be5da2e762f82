// Header-only C++ class with the reference's method names over the C ABI (slam_odom.h).
//
// Drop-in for the reference's `RGBDOdometryef` (src/odom/RGBDOdometryef.h:28-70) for callers that can hand
// over linear device pointers instead of `GPUTexture*` -- which is what the reference itself turns its
// textures into on the first line of every init call (RGBDOdometryef.cpp:126,150,155,178,183).  Same method
// names, same argument meaning, same call-order contract (initICP* before initRGB*), same public result
// fields.  Eigen is optional: define SLAM_ODOM_WITH_EIGEN before including to get the Eigen overloads with
// the reference's exact signatures; without it the plain-array overloads below are used.
//
// Errors: the reference prints and exit(0)s on CUDA failures (cuda/convenience.cuh:64-71); this class throws
// std::runtime_error with the library's message instead.  There is no CPU fallback.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>

#include "slam_odom.h"

#ifdef SLAM_ODOM_WITH_EIGEN
#include <Eigen/Core>
#endif

class RGBDOdometryef
{
  public:
    RGBDOdometryef(int width, int height, float cx, float cy, float fx, float fy, float distThresh = 0.10f,
                   float angleThresh = std::sin(20.f * 3.14159254f / 180.f), int device = 0, void * cudaStream = nullptr)
     : lastICPError(0), lastICPCount((float)(width * height)), lastRGBError(0), lastRGBCount((float)(width * height)), lastSO3Error(0),
       lastSO3Count((float)(width * height)), handle_(nullptr)
    {
        for(double & v : lastA) v = 0;
        for(double & v : lastb) v = 0;
        slam_odom_params p = {};
        p.width = width; p.height = height;
        p.cx = cx; p.cy = cy; p.fx = fx; p.fy = fy;
        p.dist_thresh = distThresh; p.angle_thresh = angleThresh;
        p.num_levels = 3;   // NUM_PYRS
        p.device = device;
        p.stream = cudaStream;
        check(slam_odom_create(&p, &handle_));
    }
    virtual ~RGBDOdometryef() { slam_odom_destroy(handle_); }
    RGBDOdometryef(const RGBDOdometryef &) = delete;
    RGBDOdometryef & operator=(const RGBDOdometryef &) = delete;

    // initICP(GPUTexture * filteredDepth, depthCutoff): R16UI millimetres, linear device memory
    void initICP(const uint16_t * filteredDepth, const float depthCutoff, size_t pitchBytes = 0)
    {
        check(slam_odom_init_icp_depth(handle_, filteredDepth, pitchBytes, depthCutoff));
    }
    // initICP(GPUTexture * predictedVertices, GPUTexture * predictedNormals, depthCutoff): RGBA32F
    void initICP(const float * predictedVertices, const float * predictedNormals, const float depthCutoff)
    {
        check(slam_odom_init_icp_maps(handle_, predictedVertices, predictedNormals, depthCutoff));
    }
    // initICPModel(vertices, normals, depthCutoff, modelPose): pose row-major 4x4
    void initICPModel(const float * predictedVertices, const float * predictedNormals, const float depthCutoff, const float * modelPose16)
    {
        check(slam_odom_init_icp_model(handle_, predictedVertices, predictedNormals, depthCutoff, modelPose16));
    }
    void initRGB(const uint8_t * rgba) { check(slam_odom_init_rgb(handle_, rgba)); }
    void initRGBModel(const uint8_t * rgba) { check(slam_odom_init_rgb_model(handle_, rgba)); }
    void initFirstRGB(const uint8_t * rgba) { check(slam_odom_init_first_rgb(handle_, rgba)); }

    // trans[3], rot[9] row-major: prior pose in, new pose out
    void getIncrementalTransformation(float * trans, float * rot, const bool & rgbOnly, const float & icpWeight, const bool & pyramid, const bool & fastOdom,
                                      const bool & so3)
    {
        check(slam_odom_get_incremental_transformation(handle_, trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3));
        refreshStats();
    }

    void getCovariance(double * out36) { check(slam_odom_get_covariance(handle_, out36)); }

#ifdef SLAM_ODOM_WITH_EIGEN
    void initICPModel(const float * predictedVertices, const float * predictedNormals, const float depthCutoff, const Eigen::Matrix4f & modelPose)
    {
        Eigen::Matrix<float, 4, 4, Eigen::RowMajor> rm = modelPose;
        initICPModel(predictedVertices, predictedNormals, depthCutoff, rm.data());
    }
    void getIncrementalTransformation(Eigen::Vector3f & trans, Eigen::Matrix<float, 3, 3, Eigen::RowMajor> & rot, const bool & rgbOnly, const float & icpWeight,
                                      const bool & pyramid, const bool & fastOdom, const bool & so3)
    {
        getIncrementalTransformation(trans.data(), rot.data(), rgbOnly, icpWeight, pyramid, fastOdom, so3);
    }
    Eigen::MatrixXd getCovariance()
    {
        Eigen::Matrix<double, 6, 6, Eigen::RowMajor> c;
        getCovariance(c.data());
        return c;
    }
#endif

    // public result fields of the reference class (RGBDOdometryef.h:62-70)
    float lastICPError;
    float lastICPCount;
    float lastRGBError;
    float lastRGBCount;
    float lastSO3Error;
    float lastSO3Count;
    double lastA[36];   // row-major 6x6
    double lastb[6];

    slam_odom_t handle() const { return handle_; }

  private:
    void check(int rc)
    {
        if(rc != SLAM_OK) throw std::runtime_error(std::string("RGBDOdometryef: ") + slam_odom_last_error());
    }
    void refreshStats()
    {
        slam_odom_stats s;
        check(slam_odom_get_stats(handle_, &s));
        lastICPError = s.lastICPError; lastICPCount = s.lastICPCount;
        lastRGBError = s.lastRGBError; lastRGBCount = s.lastRGBCount;
        lastSO3Error = s.lastSO3Error; lastSO3Count = s.lastSO3Count;
        for(int i = 0; i < 36; i++) lastA[i] = s.lastA[i];
        for(int i = 0; i < 6; i++) lastb[i] = s.lastb[i];
    }
    slam_odom_t handle_;
};
