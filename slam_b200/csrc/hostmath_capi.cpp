// Host-only C exports of small_math.hpp so the CPU test-suite can check the Eigen-free algebra
// against numpy (tests/test_host_math.py).  Built into slam_b200/libslam_hostmath.so with g++.
#include "small_math.hpp"

extern "C" {
void hm_mat3_inverse_f(const float * m, float * out) { smath::mat3_inverse(m, out); }
void hm_mat3_inverse_d(const double * m, double * out) { smath::mat3_inverse(m, out); }
void hm_mat4_inverse_d(const double * m, double * out) { smath::mat4_inverse(m, out); }
void hm_mat4_affine_inverse_d(const double * m, double * out) { smath::mat4_affine_inverse(m, out); }
void hm_ldlt6_d(const double * A, const double * b, double * x) { smath::spd_solve6(A, b, x); }
void hm_ldlt6_pivoted_d(const double * A, const double * b, double * x) { smath::ldlt_solve_pivoted<double, 6>(A, b, x, DBL_EPSILON); }
void hm_ldlt3_f(const float * A, const float * b, float * x) { smath::ldlt_solve<float, 3>(A, b, x, FLT_EPSILON); }
void hm_rodrigues(const double * r, double * R) { smath::rodrigues(r, R); }
void hm_update_se3(double * resultRt, const double * x) { smath::update_se3(resultRt, x); }
void hm_compose_current_pose(const float * Rprev, const float * tprev, const double * resultRt, float * Rcurr, float * tcurr)
{
    smath::compose_current_pose(Rprev, tprev, resultRt, Rcurr, tcurr);
}
int hm_lu_inverse6(const double * A, double * out) { return smath::lu_inverse<double, 6>(A, out) ? 1 : 0; }
}
