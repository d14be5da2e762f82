// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product.
//
// Replay harness around the REFERENCE's own CUDA code.  oracle/build (slam_b200/build.py:
// build_reference) compiles /root/reference/src/odom/reduce.cu, src/odom/utils.cu and
// src/cuda/containers/device_memory.cpp unmodified (where they lie, with oracle/ref_shim.cuh
// force-included) and links them with this file into oracle/_ref/libslam_ref.so.
//
// The reference's host class (src/odom/RGBDOdometryef.cpp) cannot be compiled here: it needs
// Eigen, Pangolin and OpenGL.  This harness therefore replays that class
// call-for-call on top of the reference's 17 free wrapper functions (src/odom/utils.cuh:62-175):
// same buffers (pitched DeviceArray2D), same GPUConfig launch shapes, same order of wrapper
// calls, same cudaDeviceSynchronize / cudaMalloc / cudaFree / tiny D2H pattern -- because those
// live inside the reference's wrappers -- and does the host algebra the reference does with
// Eigen using slam_b200/csrc/small_math.hpp.  GL texture inputs become linear device
// pointers that are copied device-to-device exactly where the reference copies its mapped
// cudaArrays (RGBDOdometryef.cpp:126,150,155,178,183).
//
// Roles: (1) parity oracle for the GPU tests (every intermediate can be tapped);
//        (2) "the reference's own CUDA path on B200" timing baseline for bench.py.
#include <vector>
#include <cmath>
#include <cfloat>
#include <limits>
#include <cstring>
#include <cstdint>

#include "odom/utils.cuh"      // reference: wrappers, DeviceArray, types, CameraModel
#include "GPUConfig.h"         // reference: launch-shape database
#include "small_math.hpp"      // ours: Eigen-free host algebra
#include "slam_odom.h"         // ours: tap ids and slam_step_record (shared with the tests)

int g_ref_shim_bind_width = 0;   // see ref_shim.cuh

namespace {

struct RefOdom
{
    // ---- members of RGBDOdometryef (RGBDOdometryef.h:72-131) ----
    // the class hard-codes NUM_PYRS = 3 (RGBDOdometryef.h:104); every wrapper it calls is level-agnostic (odom/utils.cuh:62-175), so the
    // replay takes the level count at run time (BASELINE configs[2]: four levels, iterations 10/5/4/4)
    static const int MAX_PYRS = 4;
    int NUM_PYRS = 3;
    std::vector<DeviceArray2D<unsigned short>> depth_tmp;
    DeviceArray<float> vmaps_tmp, nmaps_tmp;
    std::vector<DeviceArray2D<float>> vmaps_g_prev_, nmaps_g_prev_, vmaps_curr_, nmaps_curr_;
    CameraModel intr;
    DeviceArray<JtJJtrSE3> sumDataSE3, outDataSE3;
    DeviceArray<int2> sumResidualRGB;
    DeviceArray<JtJJtrSO3> sumDataSO3, outDataSO3;
    int sobelSize;
    float sobelScale, maxDepthDeltaRGB, maxDepthRGB;
    DeviceArray2D<float> lastDepth[MAX_PYRS], nextDepth[MAX_PYRS];
    DeviceArray2D<unsigned char> lastImage[MAX_PYRS], nextImage[MAX_PYRS], lastNextImage[MAX_PYRS];
    DeviceArray2D<short> nextdIdx[MAX_PYRS], nextdIdy[MAX_PYRS];
    DeviceArray2D<DataTerm> corresImg[MAX_PYRS];
    DeviceArray2D<float3> pointClouds[MAX_PYRS];
    std::vector<int> iterations;
    std::vector<float> minimumGradientMagnitudes;
    float distThres_, angleThres_;
    int width, height;

    float lastICPError, lastICPCount, lastRGBError, lastRGBCount, lastSO3Error, lastSO3Count;
    double lastA[36], lastb[6];

    bool trace_on = false;
    std::vector<slam_step_record> trace;
    int user_iterations[4] = {0, 0, 0, 0};

    // RGBDOdometryef.cpp:21-111
    RefOdom(int w, int h, float cx, float cy, float fx, float fy, float distThresh, float angleThresh, int levels = 3)
     : NUM_PYRS(levels), sobelSize(3), sobelScale(1.0 / pow(2.0, 3)), maxDepthDeltaRGB(0.07), maxDepthRGB(6.0), distThres_(distThresh), angleThres_(angleThresh),
       width(w), height(h)
    {
        lastICPError = 0; lastICPCount = w * h; lastRGBError = 0; lastRGBCount = w * h; lastSO3Error = 0; lastSO3Count = w * h;
        memset(lastA, 0, sizeof(lastA));
        memset(lastb, 0, sizeof(lastb));
        sumDataSE3.create(MAX_THREADS);
        outDataSE3.create(1);
        sumResidualRGB.create(MAX_THREADS);
        sumDataSO3.create(MAX_THREADS);
        outDataSO3.create(1);
        for(int i = 0; i < NUM_PYRS; i++)
        {
            const int r = h >> i, c = w >> i;
            lastDepth[i].create(r, c); lastImage[i].create(r, c);
            nextDepth[i].create(r, c); nextImage[i].create(r, c);
            lastNextImage[i].create(r, c);
            nextdIdx[i].create(r, c); nextdIdy[i].create(r, c);
            pointClouds[i].create(r, c);
            corresImg[i].create(r, c);
        }
        intr.cx = cx; intr.cy = cy; intr.fx = fx; intr.fy = fy;
        iterations.resize(NUM_PYRS);
        depth_tmp.resize(NUM_PYRS);
        vmaps_g_prev_.resize(NUM_PYRS); nmaps_g_prev_.resize(NUM_PYRS);
        vmaps_curr_.resize(NUM_PYRS); nmaps_curr_.resize(NUM_PYRS);
        for(int i = 0; i < NUM_PYRS; ++i)
        {
            const int r = h >> i, c = w >> i;
            depth_tmp[i].create(r, c);
            vmaps_g_prev_[i].create(r * 3, c); nmaps_g_prev_[i].create(r * 3, c);
            vmaps_curr_[i].create(r * 3, c); nmaps_curr_[i].create(r * 3, c);
        }
        vmaps_tmp.create(h * 4 * w);
        nmaps_tmp.create(h * 4 * w);
        minimumGradientMagnitudes.resize(NUM_PYRS);
        minimumGradientMagnitudes[0] = 5;
        minimumGradientMagnitudes[1] = 3;
        minimumGradientMagnitudes[2] = 1;
        if(NUM_PYRS > 3) minimumGradientMagnitudes[3] = 1;
    }

    // RGBDOdometryef.cpp:118-142
    void initICP(const unsigned short * d_depth, float depthCutoff)
    {
        cudaMemcpy2D(depth_tmp[0].ptr(0), depth_tmp[0].step(), d_depth, width * 2, depth_tmp[0].colsBytes(), depth_tmp[0].rows(), cudaMemcpyDeviceToDevice);
        for(int i = 1; i < NUM_PYRS; ++i) pyrDown(depth_tmp[i - 1], depth_tmp[i]);
        for(int i = 0; i < NUM_PYRS; ++i)
        {
            createVMap(intr(i), depth_tmp[i], vmaps_curr_[i], depthCutoff);
            createNMap(vmaps_curr_[i], nmaps_curr_[i]);
        }
        cudaDeviceSynchronize();
    }

    // RGBDOdometryef.cpp:144-167
    void initICP(const float * d_v4, const float * d_n4, float)
    {
        cudaMemcpy(vmaps_tmp.ptr(), d_v4, vmaps_tmp.sizeBytes(), cudaMemcpyDeviceToDevice);
        cudaMemcpy(nmaps_tmp.ptr(), d_n4, nmaps_tmp.sizeBytes(), cudaMemcpyDeviceToDevice);
        copyMaps(vmaps_tmp, nmaps_tmp, vmaps_curr_[0], nmaps_curr_[0]);
        for(int i = 1; i < NUM_PYRS; ++i)
        {
            resizeVMap(vmaps_curr_[i - 1], vmaps_curr_[i]);
            resizeNMap(nmaps_curr_[i - 1], nmaps_curr_[i]);
        }
        cudaDeviceSynchronize();
    }

    // RGBDOdometryef.cpp:169-206
    void initICPModel(const float * d_v4, const float * d_n4, float, const float * pose16)
    {
        cudaMemcpy(vmaps_tmp.ptr(), d_v4, vmaps_tmp.sizeBytes(), cudaMemcpyDeviceToDevice);
        cudaMemcpy(nmaps_tmp.ptr(), d_n4, nmaps_tmp.sizeBytes(), cudaMemcpyDeviceToDevice);
        copyMaps(vmaps_tmp, nmaps_tmp, vmaps_g_prev_[0], nmaps_g_prev_[0]);
        for(int i = 1; i < NUM_PYRS; ++i)
        {
            resizeVMap(vmaps_g_prev_[i - 1], vmaps_g_prev_[i]);
            resizeNMap(nmaps_g_prev_[i - 1], nmaps_g_prev_[i]);
        }
        mat33 device_Rcam;
        device_Rcam.data[0] = make_float3(pose16[0], pose16[1], pose16[2]);
        device_Rcam.data[1] = make_float3(pose16[4], pose16[5], pose16[6]);
        device_Rcam.data[2] = make_float3(pose16[8], pose16[9], pose16[10]);
        float3 device_tcam = make_float3(pose16[3], pose16[7], pose16[11]);
        for(int i = 0; i < NUM_PYRS; ++i)
            tranformMaps(vmaps_g_prev_[i], nmaps_g_prev_[i], device_Rcam, device_tcam, vmaps_g_prev_[i], nmaps_g_prev_[i]);
        cudaDeviceSynchronize();
    }

    // RGBDOdometryef.cpp:208-235
    void populateRGBDData(const unsigned char * d_rgba, DeviceArray2D<float> * destDepths, DeviceArray2D<unsigned char> * destImages)
    {
        verticesToDepth(vmaps_tmp, destDepths[0], maxDepthRGB);
        for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownGaussF(destDepths[i], destDepths[i + 1]);
        g_ref_shim_bind_width = width;
        imageBGRToIntensity(reinterpret_cast<cudaArray *>(const_cast<unsigned char *>(d_rgba)), destImages[0]);
        for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownUcharGauss(destImages[i], destImages[i + 1]);
        cudaDeviceSynchronize();
    }
    void initRGBModel(const unsigned char * d_rgba) { populateRGBDData(d_rgba, &lastDepth[0], &lastImage[0]); }
    void initRGB(const unsigned char * d_rgba) { populateRGBDData(d_rgba, &nextDepth[0], &nextImage[0]); }
    // RGBDOdometryef.cpp:249-265
    void initFirstRGB(const unsigned char * d_rgba)
    {
        g_ref_shim_bind_width = width;
        imageBGRToIntensity(reinterpret_cast<cudaArray *>(const_cast<unsigned char *>(d_rgba)), lastNextImage[0]);
        for(int i = 0; i + 1 < NUM_PYRS; i++) pyrDownUcharGauss(lastNextImage[i], lastNextImage[i + 1]);
    }

    static void kmat(const CameraModel & c, double * K)
    {
        for(int i = 0; i < 9; i++) K[i] = 0;
        K[0] = c.fx; K[4] = c.fy; K[2] = c.cx; K[5] = c.cy; K[8] = 1;
    }
    static mat33 to_mat33(const float * m)
    {
        mat33 r;
        r.data[0] = make_float3(m[0], m[1], m[2]);
        r.data[1] = make_float3(m[3], m[4], m[5]);
        r.data[2] = make_float3(m[6], m[7], m[8]);
        return r;
    }
    static void pack_se3(const float * A, const float * b, float * s)
    {
        int shift = 0;
        for(int i = 0; i < 6; ++i)
            for(int j = i; j < 7; ++j) s[shift++] = (j == 6) ? b[i] : A[i * 6 + j];
    }

    // RGBDOdometryef.cpp:267-595
    void getIncrementalTransformation(float * trans, float * rot, bool rgbOnly, float icpWeight, bool pyramid, bool fastOdom, bool so3)
    {
        const bool icp = !rgbOnly && icpWeight > 0;
        const bool rgb = rgbOnly || icpWeight < 100;
        trace.clear();

        float Rprev[9], tprev[3], Rcurr[9], tcurr[3];
        memcpy(Rprev, rot, 36); memcpy(tprev, trans, 12);
        memcpy(Rcurr, Rprev, 36); memcpy(tcurr, tprev, 12);

        if(rgb)
            for(int i = 0; i < NUM_PYRS; i++) computeDerivativeImages(nextImage[i], nextdIdx[i], nextdIdy[i]);

        double resultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        if(so3)
        {
            const int pyramidLevel = 2;
            float R_lr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            double K[9], Kinv[9];
            kmat(intr(pyramidLevel), K);
            float lastError = std::numeric_limits<float>::max() / 2;
            float lastCount = std::numeric_limits<float>::max() / 2;
            double lastResultR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            for(int i = 0; i < 10; i++)
            {
                float jtj[9], jtr[3];
                double KR[9], H[9];
                smath::mat3_inverse(K, Kinv);
                smath::mat3_mul(K, resultR, KR);
                smath::mat3_mul(KR, Kinv, H);
                float Hf[9], Kinvf[9], KRf[9];
                for(int k = 0; k < 9; k++) { Hf[k] = (float)H[k]; Kinvf[k] = (float)Kinv[k]; KRf[k] = (float)KR[k]; }
                float residual[2];
                so3Step(lastNextImage[pyramidLevel], nextImage[pyramidLevel], to_mat33(Hf), to_mat33(Kinvf), to_mat33(KRf), sumDataSO3, outDataSO3, jtj,
                        jtr, &residual[0], GPUConfig::getInstance().so3StepThreads, GPUConfig::getInstance().so3StepBlocks);
                lastSO3Error = sqrtf(residual[0]) / residual[1];
                lastSO3Count = residual[1];

                slam_step_record rec;
                memset(&rec, 0, sizeof(rec));
                if(trace_on)
                {
                    rec.kind = 0; rec.level = pyramidLevel; rec.iteration = i;
                    outDataSO3.download((JtJJtrSO3 *)rec.so3);
                    memcpy(rec.so3_in, Hf, 36); memcpy(rec.so3_in + 9, Kinvf, 36); memcpy(rec.so3_in + 18, KRf, 36);
                }
                bool stop = false;
                if(lastSO3Error < lastError && lastCount == lastSO3Count)
                    stop = true;
                else if((double)lastSO3Error > (double)lastError + 0.001)
                {
                    lastSO3Error = lastError;
                    lastSO3Count = lastCount;
                    memcpy(resultR, lastResultR, sizeof(resultR));
                    stop = true;
                }
                if(!stop)
                {
                    lastError = lastSO3Error;
                    lastCount = lastSO3Count;
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
                if(trace_on)
                {
                    for(int k = 0; k < 9; k++) rec.Rcurr[k] = (float)resultR[k];
                    trace.push_back(rec);
                }
                if(stop) break;
            }
        }

        iterations[0] = fastOdom ? 3 : 10;
        iterations[1] = pyramid ? 5 : 0;
        iterations[2] = pyramid ? 4 : 0;
        if(NUM_PYRS > 3) iterations[3] = pyramid ? 4 : 0;
        if(user_iterations[0] || user_iterations[1] || user_iterations[2] || user_iterations[3])
            for(int i = 0; i < NUM_PYRS; i++) iterations[i] = user_iterations[i];

        float Rprev_inv[9];
        smath::mat3_inverse(Rprev, Rprev_inv);
        mat33 device_Rprev_inv = to_mat33(Rprev_inv);
        float3 device_tprev = make_float3(tprev[0], tprev[1], tprev[2]);

        double resultRt[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        if(so3)
            for(int x = 0; x < 3; x++)
                for(int y = 0; y < 3; y++) resultRt[x * 4 + y] = resultR[x * 3 + y];

        for(int i = NUM_PYRS - 1; i >= 0; i--)
        {
            if(rgb) projectToPointCloud(lastDepth[i], pointClouds[i], intr, i);
            double K[9], Kinv[9];
            kmat(intr(i), K);
            smath::mat3_inverse(K, Kinv);
            lastRGBError = std::numeric_limits<float>::max();

            for(int j = 0; j < iterations[i]; j++)
            {
                double Rt[16], R[9], KR[9], KRK_inv[9];
                smath::mat4_affine_inverse(resultRt, Rt);
                for(int x = 0; x < 3; x++)
                    for(int y = 0; y < 3; y++) R[x * 3 + y] = Rt[x * 4 + y];
                smath::mat3_mul(K, R, KR);
                smath::mat3_mul(KR, Kinv, KRK_inv);
                float krk[9];
                for(int k = 0; k < 9; k++) krk[k] = (float)KRK_inv[k];
                mat33 krkInv = to_mat33(krk);
                float3 kt;
                kt.x = (float)smath::dot3(K[0], Rt[3], K[1], Rt[7], K[2], Rt[11]);
                kt.y = (float)smath::dot3(K[3], Rt[3], K[4], Rt[7], K[5], Rt[11]);
                kt.z = (float)smath::dot3(K[6], Rt[3], K[7], Rt[7], K[8], Rt[11]);

                int sigma = 0;
                int rgbSize = 0;
                slam_step_record rec;
                memset(&rec, 0, sizeof(rec));
                rec.kind = 1; rec.level = i; rec.iteration = j;
                memcpy(rec.Rcurr_in, Rcurr, 36); memcpy(rec.tcurr_in, tcurr, 12);
                memcpy(rec.so3_in, Rprev_inv, 36);
                memcpy(rec.krkinv_in, krk, 36); rec.kt_in[0] = kt.x; rec.kt_in[1] = kt.y; rec.kt_in[2] = kt.z;

                if(rgb)
                    computeRgbResidual(pow(minimumGradientMagnitudes[i], 2.0) / pow(sobelScale, 2.0), nextdIdx[i], nextdIdy[i], lastDepth[i], nextDepth[i],
                                       lastImage[i], nextImage[i], corresImg[i], sumResidualRGB, maxDepthDeltaRGB, kt, krkInv, sigma, rgbSize,
                                       GPUConfig::getInstance().rgbResThreads, GPUConfig::getInstance().rgbResBlocks);
                rec.rgb_count = rgbSize;
                rec.rgb_sigma = sigma;

                float sigmaVal = std::sqrt((float)sigma / rgbSize == 0 ? 1 : rgbSize);
                float rgbError = std::sqrt(sigma) / (rgbSize == 0 ? 1 : rgbSize);
                if(rgbOnly && rgbError > lastRGBError) break;
                lastRGBError = rgbError;
                lastRGBCount = rgbSize;
                if(rgbOnly) sigmaVal = -1;
                rec.sigma_in = sigmaVal;

                float A_icp[36] = {0}, b_icp[6] = {0};
                mat33 device_Rcurr = to_mat33(Rcurr);
                float3 device_tcurr = make_float3(tcurr[0], tcurr[1], tcurr[2]);
                float residual[2] = {0, 0};
                if(icp)
                {
                    icpStep(device_Rcurr, device_tcurr, vmaps_curr_[i], nmaps_curr_[i], device_Rprev_inv, device_tprev, intr(i), vmaps_g_prev_[i],
                            nmaps_g_prev_[i], distThres_, angleThres_, sumDataSE3, outDataSE3, A_icp, b_icp, &residual[0],
                            GPUConfig::getInstance().icpStepThreads, GPUConfig::getInstance().icpStepBlocks);
                    lastICPError = sqrtf(residual[0]) / residual[1];
                    lastICPCount = residual[1];
                    if(trace_on) outDataSE3.download((JtJJtrSE3 *)rec.icp);
                }
                float A_rgbd[36] = {0}, b_rgbd[6] = {0};
                if(rgb)
                {
                    rgbStep(corresImg[i], sigmaVal, pointClouds[i], intr(i).fx, intr(i).fy, nextdIdx[i], nextdIdy[i], sobelScale, sumDataSE3, outDataSE3,
                            A_rgbd, b_rgbd, GPUConfig::getInstance().rgbStepThreads, GPUConfig::getInstance().rgbStepBlocks);
                    if(trace_on) outDataSE3.download((JtJJtrSE3 *)rec.rgb);
                }

                double result[6];
                if(icp && rgb)
                {
                    const double w = icpWeight;
                    for(int k = 0; k < 36; k++) lastA[k] = (double)A_rgbd[k] + w * w * (double)A_icp[k];
                    for(int k = 0; k < 6; k++) lastb[k] = (double)b_rgbd[k] + w * (double)b_icp[k];
                }
                else if(icp)
                {
                    for(int k = 0; k < 36; k++) lastA[k] = A_icp[k];
                    for(int k = 0; k < 6; k++) lastb[k] = b_icp[k];
                }
                else
                {
                    for(int k = 0; k < 36; k++) lastA[k] = A_rgbd[k];
                    for(int k = 0; k < 6; k++) lastb[k] = b_rgbd[k];
                }
                smath::spd_solve6(lastA, lastb, result);
                smath::update_se3(resultRt, result);
                smath::compose_current_pose(Rprev, tprev, resultRt, Rcurr, tcurr);

                if(trace_on)
                {
                    for(int k = 0; k < 6; k++) rec.x[k] = result[k];
                    memcpy(rec.Rcurr, Rcurr, 36);
                    memcpy(rec.tcurr, tcurr, 12);
                    trace.push_back(rec);
                }
            }
        }
        if(rgb)
        {
            const float dx = tcurr[0] - tprev[0], dy = tcurr[1] - tprev[1], dz = tcurr[2] - tprev[2];
            if(sqrtf(dx * dx + dy * dy + dz * dz) > 0.3)
            {
                memcpy(Rcurr, Rprev, 36);
                memcpy(tcurr, tprev, 12);
            }
        }
        if(so3)
            for(int i = 0; i < NUM_PYRS; i++) std::swap(lastNextImage[i], nextImage[i]);
        memcpy(trans, tcurr, 12);
        memcpy(rot, Rcurr, 36);
    }
};

template <class T>
void upload2d(DeviceArray2D<T> & dst, const void * d_src, int rows, int cols)
{
    dst.create(rows, cols);
    cudaMemcpy2D(dst.ptr(0), dst.step(), d_src, cols * sizeof(T), cols * sizeof(T), rows, cudaMemcpyDeviceToDevice);
}
template <class T>
void download2d(const DeviceArray2D<T> & src, void * d_dst)
{
    cudaMemcpy2D(d_dst, src.cols() * sizeof(T), src.ptr(0), src.step(), src.cols() * sizeof(T), src.rows(), cudaMemcpyDeviceToDevice);
}

mat33 m33(const float * m)
{
    mat33 r;
    r.data[0] = make_float3(m[0], m[1], m[2]);
    r.data[1] = make_float3(m[3], m[4], m[5]);
    r.data[2] = make_float3(m[6], m[7], m[8]);
    return r;
}

}   // namespace

// ------------------------------------------------------------------ C ABI (mirrors slam_odom.h)
extern "C" {

void * ref_odom_create(int width, int height, float cx, float cy, float fx, float fy, float distThresh, float angleThresh)
{
    if(distThresh == 0) distThresh = 0.10f;
    if(angleThresh == 0) angleThresh = sinf(20.f * 3.14159254f / 180.f);
    return new RefOdom(width, height, cx, cy, fx, fy, distThresh, angleThresh);
}
void * ref_odom_create_levels(int width, int height, float cx, float cy, float fx, float fy, float distThresh, float angleThresh, int levels)
{
    if(levels < 3 || levels > RefOdom::MAX_PYRS) return nullptr;
    if(distThresh == 0) distThresh = 0.10f;
    if(angleThresh == 0) angleThresh = sinf(20.f * 3.14159254f / 180.f);
    return new RefOdom(width, height, cx, cy, fx, fy, distThresh, angleThresh, levels);
}
void ref_odom_destroy(void * h) { delete (RefOdom *)h; }
void ref_odom_set_iterations4(void * h, int i0, int i1, int i2, int i3)
{
    RefOdom * r = (RefOdom *)h;
    r->user_iterations[0] = i0; r->user_iterations[1] = i1; r->user_iterations[2] = i2; r->user_iterations[3] = i3;
}
void ref_odom_set_iterations(void * h, int i0, int i1, int i2)
{
    RefOdom * r = (RefOdom *)h;
    r->user_iterations[0] = i0; r->user_iterations[1] = i1; r->user_iterations[2] = i2;
}
void ref_odom_init_icp_depth(void * h, const uint16_t * d_depth, float cutoff) { ((RefOdom *)h)->initICP(d_depth, cutoff); }
void ref_odom_init_icp_maps(void * h, const float * v4, const float * n4, float cutoff) { ((RefOdom *)h)->initICP(v4, n4, cutoff); }
void ref_odom_init_icp_model(void * h, const float * v4, const float * n4, float cutoff, const float * pose16) { ((RefOdom *)h)->initICPModel(v4, n4, cutoff, pose16); }
void ref_odom_init_rgb(void * h, const uint8_t * rgba) { ((RefOdom *)h)->initRGB(rgba); }
void ref_odom_init_rgb_model(void * h, const uint8_t * rgba) { ((RefOdom *)h)->initRGBModel(rgba); }
void ref_odom_init_first_rgb(void * h, const uint8_t * rgba) { ((RefOdom *)h)->initFirstRGB(rgba); }
void ref_odom_get_incremental_transformation(void * h, float * trans, float * rot, int rgbOnly, float icpWeight, int pyramid, int fastOdom, int so3)
{
    ((RefOdom *)h)->getIncrementalTransformation(trans, rot, rgbOnly != 0, icpWeight, pyramid != 0, fastOdom != 0, so3 != 0);
}
void ref_odom_get_stats(void * h, slam_odom_stats * st)
{
    RefOdom * r = (RefOdom *)h;
    memset(st, 0, sizeof(*st));
    st->lastICPError = r->lastICPError; st->lastICPCount = r->lastICPCount;
    st->lastRGBError = r->lastRGBError; st->lastRGBCount = r->lastRGBCount;
    st->lastSO3Error = r->lastSO3Error; st->lastSO3Count = r->lastSO3Count;
    memcpy(st->lastA, r->lastA, sizeof(r->lastA));
    memcpy(st->lastb, r->lastb, sizeof(r->lastb));
}
void ref_odom_set_trace(void * h, int on) { ((RefOdom *)h)->trace_on = on != 0; }
int ref_odom_get_trace(void * h, slam_step_record * out, int max_records)
{
    RefOdom * r = (RefOdom *)h;
    const int n = (int)r->trace.size();
    for(int i = 0; i < n && i < max_records; i++) out[i] = r->trace[i];
    return n;
}

// Copy an internal buffer to a dense HOST buffer (same tap ids / layouts as slam_odom_tap).
int ref_odom_tap(void * h, int tap, int level, void * host_dst)
{
    RefOdom * r = (RefOdom *)h;
    cudaDeviceSynchronize();
    switch(tap)
    {
        case SLAM_TAP_DEPTH_U16: r->depth_tmp[level].download(host_dst, r->depth_tmp[level].cols() * 2); break;
        case SLAM_TAP_VMAP_CURR: r->vmaps_curr_[level].download(host_dst, r->vmaps_curr_[level].cols() * 4); break;
        case SLAM_TAP_NMAP_CURR: r->nmaps_curr_[level].download(host_dst, r->nmaps_curr_[level].cols() * 4); break;
        case SLAM_TAP_VMAP_PREV: r->vmaps_g_prev_[level].download(host_dst, r->vmaps_g_prev_[level].cols() * 4); break;
        case SLAM_TAP_NMAP_PREV: r->nmaps_g_prev_[level].download(host_dst, r->nmaps_g_prev_[level].cols() * 4); break;
        case SLAM_TAP_LAST_DEPTH: r->lastDepth[level].download(host_dst, r->lastDepth[level].cols() * 4); break;
        case SLAM_TAP_NEXT_DEPTH: r->nextDepth[level].download(host_dst, r->nextDepth[level].cols() * 4); break;
        case SLAM_TAP_LAST_IMAGE: r->lastImage[level].download(host_dst, r->lastImage[level].cols()); break;
        case SLAM_TAP_NEXT_IMAGE: r->nextImage[level].download(host_dst, r->nextImage[level].cols()); break;
        case SLAM_TAP_LASTNEXT_IMAGE: r->lastNextImage[level].download(host_dst, r->lastNextImage[level].cols()); break;
        case SLAM_TAP_DIDX: r->nextdIdx[level].download(host_dst, r->nextdIdx[level].cols() * 2); break;
        case SLAM_TAP_DIDY: r->nextdIdy[level].download(host_dst, r->nextdIdy[level].cols() * 2); break;
        case SLAM_TAP_CLOUD: r->pointClouds[level].download(host_dst, r->pointClouds[level].cols() * 12); break;
        case SLAM_TAP_CORRES:   // indexed linearly by the reference (reduce.cu:838)
            cudaMemcpy(host_dst, r->corresImg[level].ptr(0), (size_t)r->corresImg[level].rows() * r->corresImg[level].cols() * 16, cudaMemcpyDeviceToHost);
            break;
        default: return -1;
    }
    return 0;
}

// ---- operator-level entry points: dense device buffers in/out, reference wrapper inside ----
void ref_op_pyr_down(const uint16_t * src, int rows, int cols, uint16_t * dst)
{
    DeviceArray2D<unsigned short> s, d;
    upload2d(s, src, rows, cols);
    pyrDown(s, d);
    cudaDeviceSynchronize();
    download2d(d, dst);
}
void ref_op_create_vmap(float fx, float fy, float cx, float cy, const uint16_t * depth, int rows, int cols, float * vmap, float cutoff, int prefill_nan)
{
    DeviceArray2D<unsigned short> s;
    DeviceArray2D<float> v;
    upload2d(s, depth, rows, cols);
    v.create(rows * 3, cols);
    if(prefill_nan) cudaMemset2D(v.ptr(0), v.step(), 0xff, cols * 4, rows * 3);
    CameraModel c(fx, fy, cx, cy, cols, rows);
    createVMap(c, s, v, cutoff);
    cudaDeviceSynchronize();
    download2d(v, vmap);
}
void ref_op_create_nmap(const float * vmap, int rows, int cols, float * nmap, int prefill_nan)
{
    DeviceArray2D<float> v, n;
    upload2d(v, vmap, rows * 3, cols);
    n.create(rows * 3, cols);
    if(prefill_nan) cudaMemset2D(n.ptr(0), n.step(), 0xff, cols * 4, rows * 3);
    createNMap(v, n);
    cudaDeviceSynchronize();
    download2d(n, nmap);
}
void ref_op_transform_maps(const float * vsrc, const float * nsrc, int rows, int cols, const float * R9, const float * t3, float * vdst, float * ndst)
{
    DeviceArray2D<float> v, n;
    upload2d(v, vsrc, rows * 3, cols);
    upload2d(n, nsrc, rows * 3, cols);
    tranformMaps(v, n, m33(R9), make_float3(t3[0], t3[1], t3[2]), v, n);
    cudaDeviceSynchronize();
    download2d(v, vdst);
    download2d(n, ndst);
}
void ref_op_copy_maps(const float * v4, const float * n4, int rows, int cols, float * vdst, float * ndst)
{
    DeviceArray<float> vs, ns;
    vs.create(rows * cols * 4);
    ns.create(rows * cols * 4);
    cudaMemcpy(vs.ptr(), v4, vs.sizeBytes(), cudaMemcpyDeviceToDevice);
    cudaMemcpy(ns.ptr(), n4, ns.sizeBytes(), cudaMemcpyDeviceToDevice);
    DeviceArray2D<float> v, n;
    v.create(rows * 3, cols);
    n.create(rows * 3, cols);
    copyMaps(vs, ns, v, n);
    cudaDeviceSynchronize();
    download2d(v, vdst);
    download2d(n, ndst);
}
void ref_op_resize_map(const float * src, int rows, int cols, float * dst, int normalize)
{
    DeviceArray2D<float> s, d;
    upload2d(s, src, rows * 3, cols);
    d.create((rows / 2) * 3, cols / 2);
    cudaMemset2D(d.ptr(0), d.step(), 0xff, (cols / 2) * 4, (rows / 2) * 3);
    if(normalize)
        resizeNMap(s, d);
    else
        resizeVMap(s, d);
    download2d(d, dst);
}
void ref_op_image_bgr_to_intensity(const uint8_t * rgba, int rows, int cols, uint8_t * dst)
{
    DeviceArray2D<unsigned char> d;
    d.create(rows, cols);
    g_ref_shim_bind_width = cols;
    imageBGRToIntensity(reinterpret_cast<cudaArray *>(const_cast<uint8_t *>(rgba)), d);
    cudaDeviceSynchronize();
    download2d(d, dst);
}
void ref_op_vertices_to_depth(const float * v4, int rows, int cols, float * dst, float cutoff)
{
    DeviceArray<float> vs;
    vs.create(rows * cols * 4);
    cudaMemcpy(vs.ptr(), v4, vs.sizeBytes(), cudaMemcpyDeviceToDevice);
    DeviceArray2D<float> d;
    d.create(rows, cols);
    verticesToDepth(vs, d, cutoff);
    cudaDeviceSynchronize();
    download2d(d, dst);
}
void ref_op_project_to_point_cloud(const float * depth, int rows, int cols, float * cloud3, float fx, float fy, float cx, float cy, int level)
{
    DeviceArray2D<float> s;
    DeviceArray2D<float3> c;
    upload2d(s, depth, rows, cols);
    c.create(rows, cols);
    CameraModel intr(fx, fy, cx, cy, cols, rows);
    projectToPointCloud(s, c, intr, level);
    download2d(c, cloud3);
}
void ref_op_pyr_down_gauss_f(const float * src, int rows, int cols, float * dst)
{
    DeviceArray2D<float> s, d;
    upload2d(s, src, rows, cols);
    pyrDownGaussF(s, d);
    cudaDeviceSynchronize();
    download2d(d, dst);
}
void ref_op_pyr_down_uchar_gauss(const uint8_t * src, int rows, int cols, uint8_t * dst)
{
    DeviceArray2D<unsigned char> s, d;
    upload2d(s, src, rows, cols);
    pyrDownUcharGauss(s, d);
    cudaDeviceSynchronize();
    download2d(d, dst);
}
void ref_op_compute_derivative_images(const uint8_t * src, int rows, int cols, int16_t * dx, int16_t * dy)
{
    DeviceArray2D<unsigned char> s;
    DeviceArray2D<short> x, y;
    upload2d(s, src, rows, cols);
    x.create(rows, cols);
    y.create(rows, cols);
    computeDerivativeImages(s, x, y);
    download2d(x, dx);
    download2d(y, dy);
}
void ref_op_icp_step(const float * Rcurr9, const float * tcurr3, const float * vmap_curr, const float * nmap_curr, const float * Rprev_inv9,
                     const float * tprev3, float fx, float fy, float cx, float cy, const float * vmap_g_prev, const float * nmap_g_prev, float dist_thresh,
                     float angle_thresh, int rows, int cols, float * host_out29)
{
    DeviceArray2D<float> vc, nc, vp, np;
    upload2d(vc, vmap_curr, rows * 3, cols);
    upload2d(nc, nmap_curr, rows * 3, cols);
    upload2d(vp, vmap_g_prev, rows * 3, cols);
    upload2d(np, nmap_g_prev, rows * 3, cols);
    DeviceArray<JtJJtrSE3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    float A[36], b[6], residual[2];
    CameraModel intr(fx, fy, cx, cy, cols, rows);
    icpStep(m33(Rcurr9), make_float3(tcurr3[0], tcurr3[1], tcurr3[2]), vc, nc, m33(Rprev_inv9), make_float3(tprev3[0], tprev3[1], tprev3[2]), intr, vp, np,
            dist_thresh, angle_thresh, sum, out, A, b, residual, GPUConfig::getInstance().icpStepThreads, GPUConfig::getInstance().icpStepBlocks);
    out.download((JtJJtrSE3 *)host_out29);
}
void ref_op_compute_rgb_residual(float min_scale, const int16_t * dIdx, const int16_t * dIdy, const float * last_depth, const float * next_depth,
                                 const uint8_t * last_image, const uint8_t * next_image, void * corres_img16, float max_depth_delta, const float * kt3,
                                 const float * krkinv9, int rows, int cols, int * host_count_sigma)
{
    DeviceArray2D<short> dx, dy;
    DeviceArray2D<float> ld, nd;
    DeviceArray2D<unsigned char> li, ni;
    DeviceArray2D<DataTerm> cimg;
    upload2d(dx, dIdx, rows, cols);
    upload2d(dy, dIdy, rows, cols);
    upload2d(ld, last_depth, rows, cols);
    upload2d(nd, next_depth, rows, cols);
    upload2d(li, last_image, rows, cols);
    upload2d(ni, next_image, rows, cols);
    cimg.create(rows, cols);
    DeviceArray<int2> sumRes;
    sumRes.create(MAX_THREADS);
    int sigma = 0, count = 0;
    computeRgbResidual(min_scale, dx, dy, ld, nd, li, ni, cimg, sumRes, max_depth_delta, make_float3(kt3[0], kt3[1], kt3[2]), m33(krkinv9), sigma, count,
                       GPUConfig::getInstance().rgbResThreads, GPUConfig::getInstance().rgbResBlocks);
    host_count_sigma[0] = count;
    host_count_sigma[1] = sigma;
    // the reference indexes corresImg linearly (reduce.cu:838): hand back the first rows*cols entries
    cudaMemcpy(corres_img16, cimg.ptr(0), (size_t)rows * cols * 16, cudaMemcpyDeviceToDevice);
}
void ref_op_rgb_step(const void * corres_img16, float sigma, const float * cloud3, float fx, float fy, const int16_t * dIdx, const int16_t * dIdy,
                     float sobel_scale, int rows, int cols, float * host_out29)
{
    DeviceArray2D<DataTerm> cimg;
    cimg.create(rows, cols);
    cudaMemcpy(cimg.ptr(0), corres_img16, (size_t)rows * cols * 16, cudaMemcpyDeviceToDevice);
    DeviceArray2D<float3> cloud;
    DeviceArray2D<short> dx, dy;
    upload2d(cloud, cloud3, rows, cols);
    upload2d(dx, dIdx, rows, cols);
    upload2d(dy, dIdy, rows, cols);
    DeviceArray<JtJJtrSE3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    float A[36], b[6];
    rgbStep(cimg, sigma, cloud, fx, fy, dx, dy, sobel_scale, sum, out, A, b, GPUConfig::getInstance().rgbStepThreads,
            GPUConfig::getInstance().rgbStepBlocks);
    out.download((JtJJtrSE3 *)host_out29);
}
void ref_op_so3_step(const uint8_t * last_image, const uint8_t * next_image, const float * image_basis9, const float * kinv9, const float * krlr9, int rows,
                     int cols, float * host_out11)
{
    DeviceArray2D<unsigned char> li, ni;
    upload2d(li, last_image, rows, cols);
    upload2d(ni, next_image, rows, cols);
    DeviceArray<JtJJtrSO3> sum, out;
    sum.create(MAX_THREADS);
    out.create(1);
    float A[9], b[3], residual[2];
    so3Step(li, ni, m33(image_basis9), m33(kinv9), m33(krlr9), sum, out, A, b, residual, GPUConfig::getInstance().so3StepThreads,
            GPUConfig::getInstance().so3StepBlocks);
    out.download((JtJJtrSO3 *)host_out11);
}

}   // extern "C"
