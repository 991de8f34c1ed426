// oracle/ref_cvo.cpp — drives the REFERENCE's own CVO class (thirdparty/cvo/src/cvo.cpp, LieGroup.cpp,
// pcd_generator.cpp, thirdparty/cvo/thirdparty/PixelSelector2.cpp, nanoflann.hpp — all compiled where they lie
// under /root/reference; `refcvo` target of oracle/Makefile) so that the oracle's restatement of SURVEY §8a
// rows I-Q can be pinned against outputs of the reference's source itself.
//
// TEST INFRASTRUCTURE.  Eigen, OpenCV, TBB and Boost are not in this image: the sources are compiled against
// the stand-in headers of oracle/shim/ (eager dense matrices in textbook order, sequential parallel_for,
// closed-form eigenvalues / matrix logarithm).  What this pins: every formula, constant, operand, cast and
// control-flow decision written in cvo.cpp / LieGroup.cpp (kernel values, thresholds, sparsification, flow,
// step-size polynomial, Exp_SEK3, stop tests, ell schedule, state persistence, inner products, Hessian).
// What it cannot pin: the rounding of Eigen's own vectorised reductions and iterative solvers.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iostream>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <unistd.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#include <opencv2/opencv.hpp>
#include <tbb/tbb.h>
#include <nanoflann.hpp>

#define private public   // this file reads and sets the private state of cvo::cvo (R, T, ell, clouds, A)
#include "cvo.hpp"
#undef private

extern "C" void oracle_gray_u8(const uint8_t *p, int n, uint8_t *out);   // (defined in ref_select.cpp, linked in)

struct RefRecord {   // == cvo_iter_record without B..E (locals of compute_step_size)
    float ell, omega[3], v[3], step;
    int nnz;
};

static void fill_cloud(cvo::point_cloud *pc, int n, const float *pos, const float *feat) {
    pc->num_points = n;
    pc->positions.resize(n);
    pc->features = Eigen::MatrixXf::Zero(n, NUM_FEATURES);
    for (int i = 0; i < n; i++) {
        pc->positions[i] = Eigen::Vector3f(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        for (int k = 0; k < NUM_FEATURES; k++) pc->features(i, k) = feat[5 * i + k];
    }
}

extern "C" {

void *refcvo_create(const float calib[5]) {
    char path[] = "/tmp/refcvo_calibXXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) return nullptr;
    FILE *f = fdopen(fd, "w");
    fprintf(f, "%%YAML:1.0\nCamera.fx: %.9g\nCamera.fy: %.9g\nCamera.cx: %.9g\nCamera.cy: %.9g\nDepthMapFactor: %.9g\n",
            calib[1], calib[2], calib[3], calib[4], calib[0]);
    fclose(f);
    cvo::cvo *c = new cvo::cvo(std::string(path));
    unlink(path);
    return c;
}
void refcvo_destroy(void *h) { delete static_cast<cvo::cvo *>(h); }

// the reference's own set_pcd (selection + features through pcd_generator): first call = fixed, later = moving
void refcvo_set_pcd(void *h, const uint8_t *bgr, const uint16_t *depth, int w, int hgt) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    cv::Mat img(hgt, w, CV_8UC3, (void *)bgr, (size_t)w * 3), dep(hgt, w, CV_16UC1, (void *)depth, (size_t)w * 2);
    cv::Mat img_own, dep_own;   // frame::image / depth alias the caller's buffers (pcd_generator.cpp:621-622): keep copies alive
    img.copyTo(img_own);
    dep.copyTo(dep_own);
    c->set_pcd(img_own, dep_own);
}
// clouds given directly: the tail of set_pcd (cvo.cpp:370-383) on caller-supplied points and features
void refcvo_set_clouds(void *h, int nf, const float *pf, const float *ff, int nm, const float *pm, const float *fm) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    c->ptr_fixed_pcd.reset(new cvo::point_cloud);
    c->ptr_moving_pcd.reset(new cvo::point_cloud);
    fill_cloud(c->ptr_fixed_pcd.get(), nf, pf, ff);
    fill_cloud(c->ptr_moving_pcd.get(), nm, pm, fm);
    c->init = true;
    c->num_fixed = nf;
    c->num_moving = nm;
    c->cloud_x = &(c->ptr_fixed_pcd->positions);
    c->cloud_y = new std::vector<Eigen::Vector3f>(c->ptr_moving_pcd->positions);
    c->A_trip_concur.reserve(nm * 20);
    c->A.resize(nf, nm);
    c->A.setZero();
    c->A_nonzero = 0;
}
int refcvo_sizes(void *h, int *nf, int *nm) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    *nf = c->ptr_fixed_pcd ? c->ptr_fixed_pcd->num_points : -1;
    *nm = c->ptr_moving_pcd ? c->ptr_moving_pcd->num_points : -1;
    return 0;
}
void refcvo_set_state(void *h, const float R[9], const float T[3], float ell) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) c->R(i, j) = R[i * 3 + j]; c->T(i) = T[i]; }
    c->ell = ell;
}
void refcvo_get_state(void *h, float R[9], float T[3], float *ell, float transform[16]) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i * 3 + j] = c->R(i, j); T[i] = c->T(i); }
    *ell = c->ell;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) transform[i * 4 + j] = c->transform.matrix()(i, j);
}
// one iteration body of cvo::align (cvo.cpp:770-779) at an injected state, without the update
int refcvo_iteration_at(void *h, const float R[9], const float T[3], float ell, RefRecord *out, int cap, int *ij, float *a) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    refcvo_set_state(h, R, T, ell);
    if (!c->cloud_y) c->cloud_y = new std::vector<Eigen::Vector3f>(c->ptr_moving_pcd->positions);
    c->update_tf();
    c->transform_pcd();
    c->compute_flow();
    c->compute_step_size();
    out->ell = ell;
    for (int k = 0; k < 3; k++) { out->omega[k] = c->omega(k); out->v[k] = c->v(k); }
    out->step = c->step;
    out->nnz = c->A_nonzero;
    int m = 0;
    for (int i = 0; i < c->num_fixed; i++)
        for (Eigen::SparseMatrix<float, Eigen::RowMajor>::InnerIterator it(c->A, i); it; ++it, ++m)
            if (m < cap) { ij[2 * m] = i; ij[2 * m + 1] = it.col(); a[m] = it.value(); }
    return m;
}
// cvo::align (cvo.cpp:763-821) from the object's current state
void refcvo_align(void *h, float transform[16], float last_iter_transform[16], int *iter, int *nnz, float *ell) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    if (!c->cloud_y) c->cloud_y = new std::vector<Eigen::Vector3f>(c->ptr_moving_pcd->positions);
    c->iter = -1;
    c->align();
    c->cloud_y = nullptr;   // (align() deleted it, cvo.cpp:820)
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        transform[i * 4 + j] = c->transform.matrix()(i, j);
        last_iter_transform[i * 4 + j] = c->prev_transform.matrix()(i, j);
    }
    *iter = c->iter;
    *nnz = c->A_nonzero;
    *ell = c->ell;
}
// cvo::compute_innerproduct (cvo.cpp:475-503): values / nums = {pre, post, fixed, moving}
void refcvo_compute_innerproduct(void *h, const float tran[16], float values[4], int nums[4], double H[36], int *inliers,
                                 float *cos_angle) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    cvo::inn_p pre, post, fx, mv;
    Eigen::Matrix<double, 6, 6> Hm;
    Eigen::Affine3f t;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t.matrix()(i, j) = tran[i * 4 + j];
    *inliers = 0;
    c->compute_innerproduct(pre, post, Hm, t, *inliers, fx, mv, *cos_angle);
    const cvo::inn_p *r[4] = {&pre, &post, &fx, &mv};
    for (int k = 0; k < 4; k++) { values[k] = r[k]->value; nums[k] = r[k]->num; }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[i * 6 + j] = Hm(i, j);
}
// cvo::compute_innerproduct_lc (cvo.cpp:505-561): values / nums = {prior, lc_prior, lc_pre, lc_post, fixed, moving};
// the four transforms are 4x4 row-major and are applied to the moving cloud
void refcvo_compute_innerproduct_lc(void *h, const float prior[16], const float lc_prior[16], const float lc_prior_2[16],
                                    const float lc[16], float values[6], int nums[6], double H[36], int *inliers_svd,
                                    int *inliers_pnpransac, float *cos_angle) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    cvo::inn_p r[6];
    Eigen::Matrix<double, 6, 6> Hm;
    Eigen::Affine3f t[4];
    const float *src[4] = {prior, lc_prior, lc_prior_2, lc};
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[k].matrix()(i, j) = src[k][i * 4 + j];
    *inliers_svd = 0;
    *inliers_pnpransac = 0;
    c->compute_innerproduct_lc(r[0], r[1], r[2], r[3], Hm, t[0], t[1], t[2], t[3], *inliers_svd, *inliers_pnpransac, r[4], r[5],
                               *cos_angle);
    for (int k = 0; k < 6; k++) { values[k] = r[k].value; nums[k] = r[k].num; }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[i * 6 + j] = Hm(i, j);
}
// the state shuffles (cvo.cpp:578-618)
void refcvo_update_fixed_pcd(void *h) { static_cast<cvo::cvo *>(h)->update_fixed_pcd(); }
void refcvo_reset_initial(void *h, const float odom[16], float back[16]) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    Eigen::Affine3f o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o.matrix()(i, j) = odom[i * 4 + j];
    Eigen::Affine3f b = c->reset_initial(o);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) back[i * 4 + j] = b.matrix()(i, j);
}
void refcvo_update_previous_pcd(void *h) { static_cast<cvo::cvo *>(h)->update_previous_pcd(); }
void refcvo_reset_keyframe(void *h, const float odom[16]) {
    Eigen::Affine3f o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o.matrix()(i, j) = odom[i * 4 + j];
    static_cast<cvo::cvo *>(h)->reset_keyframe(o);
}
void refcvo_reset_transform(void *h, const float odom[16]) {
    Eigen::Affine3f o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o.matrix()(i, j) = odom[i * 4 + j];
    static_cast<cvo::cvo *>(h)->reset_transform(o);
}
// points in the fixed / moving / previous slot (-1 = the slot is empty: its unique_ptr was moved from)
void refcvo_slot_sizes(void *h, int n[3]) {
    cvo::cvo *c = static_cast<cvo::cvo *>(h);
    n[0] = c->ptr_fixed_pcd ? c->ptr_fixed_pcd->num_points : -1;
    n[1] = c->ptr_moving_pcd ? c->ptr_moving_pcd->num_points : -1;
    n[2] = c->ptr_previous_pcd ? c->ptr_previous_pcd->num_points : -1;
}
void refcvo_set_max_iter(void *h, int n) { static_cast<cvo::cvo *>(h)->MAX_ITER = n; }

}  // extern "C"
