/*
 * cvo.hpp — drop-in replacement for thirdparty/cvo/include/cvo.hpp of bexilin/CVO-SLAM.
 *
 * Same namespace, class names, member names and signatures as the reference's `cvo::cvo` and
 * `cvo::inn_p` (thirdparty/cvo/include/cvo.hpp:52-282), so that src/local_tracker.cpp,
 * src/keyframe_graph.cpp and include/tracking_result.h compile unchanged; every numerical
 * method forwards to the C ABI of libcvo_b200.so (include/cvo_b200.h).  The state shuffles of
 * cvo.cpp:578-618 are host logic and live here.
 *
 * Types.  With Eigen and OpenCV on the include path (the reference's build) the class uses
 * Eigen::Affine3f / Eigen::Matrix<double,6,6> / cv::Mat / cv::Point2f and the reference's own
 * data_type.h.  Without them (this repository's CI image has neither) a minimal stand-in with
 * the same member spelling is used (namespace cvo::shim) so that the logic can be compiled and
 * tested; see tests/cpp/dropin_smoke.cpp.
 */
#ifndef RKHS_SE3_H
#define RKHS_SE3_H

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cvo_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Geometry>) && __has_include(<opencv2/core/mat.hpp>) && !defined(CVO_B200_FORCE_SHIM)
#define CVO_B200_HAVE_EIGEN_OPENCV 1
#endif
#endif

#ifdef CVO_B200_HAVE_EIGEN_OPENCV
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/core/mat.hpp>
#include "data_type.h" /* the reference's frame / point_cloud / camera_info */
#else
namespace cvo {
namespace shim {
struct Mat44f {
    float m[16];
    float &operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
};
struct Mat33f {
    float m[9];
    float &operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct Vec3f {
    float v[3];
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
};
/* The part of Eigen::Affine3f the callers of cvo::cvo use (src/local_tracker.cpp, src/keyframe_graph.cpp):
 * Identity(), matrix(), linear(), translation(), inverse(), operator*, cast<>(). */
struct Affine3f {
    Mat44f mat;
    Affine3f() { *this = Identity(); }
    static Affine3f Identity() {
        Affine3f a(0);
        for (int i = 0; i < 16; i++) a.mat.m[i] = (i % 5 == 0) ? 1.f : 0.f;
        return a;
    }
    Mat44f &matrix() { return mat; }
    const Mat44f &matrix() const { return mat; }
    Mat33f linear() const {
        Mat33f l;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) l(r, c) = mat(r, c);
        return l;
    }
    Mat33f rotation() const { return linear(); } /* (Eigen's rotation() re-orthogonalises: see reset_initial) */
    Vec3f translation() const { Vec3f t; for (int r = 0; r < 3; r++) t[r] = mat(r, 3); return t; }
    Affine3f operator*(const Affine3f &o) const {
        Affine3f out(0);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
            float s = 0;
            for (int k = 0; k < 4; k++) s += mat(r, k) * o.mat(k, c);
            out.mat(r, c) = s;
        }
        return out;
    }
    Affine3f inverse() const; /* cofactor inverse of the linear part (defined below, detail::inv_affine) */
    template <class T> Affine3f cast() const { return *this; }
private:
    explicit Affine3f(int) {}
};
typedef Affine3f Affine3d; /* the stand-in keeps float storage */
struct Matrix66d {
    double m[36];
    double &operator()(int r, int c) { return m[r * 6 + c]; }
    double operator()(int r, int c) const { return m[r * 6 + c]; }
};
struct Point2f {
    float x, y;
    Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {}
};
struct Mat { /* the subset of cv::Mat the path touches: 8UC3 or 16UC1, row stride in bytes */
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;
};
struct point_cloud { /* thirdparty/cvo/include/data_type.h:67-79 */
    int num_points = 0;
    std::vector<std::array<float, 3>> positions;
    std::vector<std::array<float, 5>> features;
};
}  // namespace shim
}  // namespace cvo
#endif

namespace cvo {

#ifdef CVO_B200_HAVE_EIGEN_OPENCV
typedef Eigen::Affine3f affine3f_t;
typedef Eigen::Affine3d affine3d_t;
typedef Eigen::Matrix<double, 6, 6> matrix66d_t;
typedef cv::Mat mat_t;
typedef cv::Point2f point2f_t;
typedef ::cvo::point_cloud point_cloud_t;
#else
typedef shim::Affine3f affine3f_t;
typedef shim::Affine3d affine3d_t;
typedef shim::Matrix66d matrix66d_t;
typedef shim::Mat mat_t;
typedef shim::Point2f point2f_t;
typedef shim::point_cloud point_cloud_t;
#endif

/* thirdparty/cvo/include/cvo.hpp:52-80 */
class inn_p {
public:
    float value;
    int num;
    int num_e;
    void copy(const inn_p &r) { value = r.value; num = r.num; num_e = r.num_e; }
    inn_p(const inn_p &r) : value(r.value), num(r.num), num_e(r.num_e) {}
    inn_p &operator=(const inn_p &r) { copy(r); return *this; }
    inn_p(float v, int n, int n_e) : value(v), num(n), num_e(n_e) {}
    inn_p() {}
};

namespace detail {
template <class A> inline void to_rows(const A &a, float out[16]) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[r * 4 + c] = (float)a.matrix()(r, c);
}
template <class A> inline void from_rows(const float in[16], A &a) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) a.matrix()(r, c) = in[r * 4 + c];
}
inline void mul44(const float a[16], const float b[16], float o[16]) {
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
        float s = 0;
        for (int k = 0; k < 4; k++) s += a[r * 4 + k] * b[k * 4 + c];
        o[r * 4 + c] = s;
    }
}
/* Eigen::Affine3f::inverse() as the reference's reset_initial uses it (cvo.cpp:613-617): the general
 * 3x3 inverse of the linear part from its cofactors (Eigen's compute_inverse_size3: determinant from
 * the first column's cofactors, one reciprocal), translation = -(inverse * t).  Not the transpose:
 * for a float rotation matrix the two differ in the last bits, and the alignment that starts from
 * this prior is sensitive to them.  Every operation in float, in this order. */
inline void inv_affine(const float a[16], float o[16]) {
    auto m = [&](int r, int c) { return a[r * 4 + c]; };
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const float det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
    const float invdet = 1.0f / det;
    float inv[9];
    inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
    inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
    inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) o[r * 4 + c] = inv[r * 3 + c];
        o[r * 4 + 3] = -((inv[r * 3] * a[3] + inv[r * 3 + 1] * a[7]) + inv[r * 3 + 2] * a[11]);
    }
    o[12] = o[13] = o[14] = 0.f;
    o[15] = 1.f;
}
}  // namespace detail
#ifndef CVO_B200_HAVE_EIGEN_OPENCV
inline shim::Affine3f shim::Affine3f::inverse() const {
    shim::Affine3f out = Identity();
    detail::inv_affine(mat.m, out.mat.m);
    return out;
}
#endif
namespace detail {
/* "Key: value" lines of an OpenCV FileStorage YAML (cvo.cpp:58-64 reads five scalars) */
inline bool yaml_scalar(const std::string &path, const std::string &key, float &out) {
    std::ifstream f(path.c_str());
    std::string line;
    while (std::getline(f, line)) {
        size_t p = line.find(key + ":");
        if (p == std::string::npos) continue;
        std::istringstream ss(line.substr(p + key.size() + 1));
        double v;
        if (ss >> v) { out = (float)v; return true; }
    }
    return false;
}
}  // namespace detail

/* thirdparty/cvo/include/cvo.hpp:82-282 */
class cvo {
private:
    cvo_handle *h_;
    cvo_handle *scratch_;  /* second handle, same parameters and device, for the public host-cloud entry points */
    cvo_params prm_;       /* the constructor's parameters (defaults of cvo.cpp:35-51 unless given)            */
    int device_;
    bool pre_pc_init;
    bool sizes_valid_;     /* num_fixed / num_moving are current (cvo.cpp:370-371 sets them in set_pcd)        */
    int num_fixed, num_moving;
    int A_nonzero;
    cvo_calib cam_info;

    cvo_handle *scratch_handle() {
        if (!scratch_ && cvo_create(&cam_info, &prm_, device_, &scratch_) != CVO_OK) scratch_ = nullptr;
        if (scratch_) {
            float ell = prm_.ell_init;
            cvo_get_ell(h_, &ell); /* the queries run at the member ell, whatever align left (cvo.cpp:391) */
            cvo_set_ell(scratch_, ell);
        }
        return scratch_;
    }

    void check(int rc, const char *what) const {
        if (rc != CVO_OK && rc != CVO_ERR_NOT_INIT)
            std::cout << "cvo_b200: " << what << " failed (" << rc << "): " << cvo_last_error() << "\n";
    }
    inn_p inner(int slot_a, const float *Ta, int slot_b) {
        float v = 0;
        int n = 0;
        check(cvo_inner_product(h_, slot_a, Ta, slot_b, &v, &n), "cvo_inner_product");
        return inn_p(v, n, 0);
    }
    matrix66d_t hess(int slot_a, const float *Ta, int slot_b, int &inliers) {
        double H[36];
        check(cvo_hessian(h_, slot_a, Ta, slot_b, H, &inliers), "cvo_hessian");
        matrix66d_t M;
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) M(r, c) = H[r * 6 + c];
        return M;
    }
    int upload(int slot, point_cloud_t *pc) {
        const int n = pc->num_points;
        std::vector<float> pos((size_t)3 * n), feat((size_t)5 * n);
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < 3; k++) pos[3 * i + k] = pc->positions[i][k];
#ifdef CVO_B200_HAVE_EIGEN_OPENCV
            for (int k = 0; k < 5; k++) feat[5 * i + k] = pc->features(i, k);
#else
            for (int k = 0; k < 5; k++) feat[5 * i + k] = pc->features[i][k];
#endif
        }
        return cvo_set_cloud(h_, slot, n, pos.data(), feat.data());
    }

public:
#ifdef CVO_B200_HAVE_EIGEN_OPENCV
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW /* reference cvo.hpp:146: `new cvo::cvo` (local_tracker.cpp:48-49) with Affine3f members */
#endif
    /* public variables (cvo.hpp:137-146) */
    bool first_frame;
    bool init;
    int iter;
    affine3f_t transform;
    affine3f_t prev_transform;
    affine3f_t accum_transform;

    /* cvo.cpp:18-71.  `device` is new (default 0); everything else as in the reference. */
    explicit cvo(const std::string &calib_file, int device = 0, const cvo_params *params = nullptr)
        : h_(nullptr), scratch_(nullptr), device_(device), pre_pc_init(false), sizes_valid_(false), num_fixed(0),
          num_moving(0), A_nonzero(0), first_frame(true), init(false), iter(0) {
        if (params) prm_ = *params;
        else cvo_default_params(&prm_);
        cam_info.fx = cam_info.fy = cam_info.cx = cam_info.cy = 0.f;
        cam_info.scaling_factor = 0.f;
        detail::yaml_scalar(calib_file, "Camera.fx", cam_info.fx);
        detail::yaml_scalar(calib_file, "Camera.fy", cam_info.fy);
        detail::yaml_scalar(calib_file, "Camera.cx", cam_info.cx);
        detail::yaml_scalar(calib_file, "Camera.cy", cam_info.cy);
        detail::yaml_scalar(calib_file, "DepthMapFactor", cam_info.scaling_factor);
        transform = affine3f_t::Identity();
        prev_transform = affine3f_t::Identity();
        accum_transform = affine3f_t::Identity();
        int rc = cvo_create(&cam_info, &prm_, device_, &h_);
        if (rc != CVO_OK) throw std::runtime_error(std::string("cvo_create: ") + cvo_last_error());
    }
    ~cvo() { cvo_destroy(scratch_); cvo_destroy(h_); }
    cvo(const cvo &) = delete;
    cvo &operator=(const cvo &) = delete;

    /* cvo.cpp:345-386 */
    void set_pcd(const mat_t &RGB_img, const mat_t &dep_img) {
        const int slot = init ? CVO_SLOT_MOVING : CVO_SLOT_FIXED;
        check(cvo_set_frame(h_, slot, (const uint8_t *)RGB_img.data, (size_t)RGB_img.step,
                            (const uint16_t *)dep_img.data, (size_t)dep_img.step, RGB_img.cols, RGB_img.rows),
              "cvo_set_frame");
        if (!init) { init = true; return; }
        sizes_valid_ = false; /* the reference refreshes num_fixed / num_moving here (cvo.cpp:370-371) */
        A_nonzero = 0;
    }

    /* cvo.cpp:763-821 */
    void align() {
        cvo_align_result r;
        int rc = cvo_align(h_, &r, nullptr, 0);
        check(rc, "cvo_align");
        if (rc != CVO_OK && rc != CVO_ERR_PAIR_OVERFLOW) return;
        if (r.iter >= 0) iter = r.iter; /* `iter` is written only on break (cvo.cpp:783,805) */
        A_nonzero = r.A_nonzero;
        num_fixed = r.num_fixed;   /* the sizes set_pcd would have cached, without a device sync */
        num_moving = r.num_moving;
        sizes_valid_ = true;
        /* cvo.cpp:815-816: prev_transform / accum_transform take `transform` as the LAST executed
         * iteration's update_tf() left it, not the final one */
        detail::from_rows(r.last_iter_transform, prev_transform);
        float a[16], c[16];
        detail::to_rows(accum_transform, a);
        detail::mul44(a, r.last_iter_transform, c);
        detail::from_rows(c, accum_transform);
        detail::from_rows(r.transform, transform);
    }

    /* Opt-in, not in the reference: LocalTracker hands the SAME image to its two cvo objects
     * (local_tracker.cpp:356,415), so the reference selects its points twice per frame.  A caller that knows this lets
     * the second object adopt the cloud the first one has just selected: set_pcd_from(src, slot) is set_pcd with the
     * device cloud of `src`'s slot (CVO_SLOT_MOVING before src.update_fixed_pcd(), CVO_SLOT_FIXED after it) in place
     * of the images; match_keyframe_from is match_keyframe on it.  Same clouds, same bits downstream. */
    void set_pcd_from(cvo &src, int src_slot) {
        const int slot = init ? CVO_SLOT_MOVING : CVO_SLOT_FIXED;
        check(cvo_copy_cloud(h_, slot, src.h_, src_slot), "cvo_copy_cloud");
        if (!init) { init = true; return; }
        sizes_valid_ = false;
        A_nonzero = 0;
    }
    void match_keyframe_from(cvo &src, int src_slot, affine3d_t &transformd) {
        if (init == false) { std::cout << "cvo not initialized !" << "\n"; return; }
        set_pcd_from(src, src_slot);
        align();
        transformd = transform.template cast<double>();
    }

    /* cvo.cpp:461-473 */
    void match_odometry(const mat_t &RGB_img, const mat_t &dep_img, affine3d_t &transformd) {
        if (init == false) { std::cout << "cvo not initialized !" << "\n"; return; }
        set_pcd(RGB_img, dep_img);
        align();
        transformd = transform.template cast<double>();
    }
    /* cvo.cpp:563-576 */
    void match_keyframe(const mat_t &RGB_img, const mat_t &dep_img, affine3d_t &transformd) {
        if (init == false) { std::cout << "cvo not initialized !" << "\n"; return; }
        set_pcd(RGB_img, dep_img);
        align();
        transformd = transform.template cast<double>();
    }

    /* cvo.cpp:388-459 and :620-759 on caller-supplied host clouds (upload path) */
    const inn_p function_inner_product(point_cloud_t *cloud_a, point_cloud_t *cloud_b) {
        cvo_handle *keep = h_;
        cvo_handle *tmp = scratch_handle(); /* same sigma / sp_thres / c_ell / c_sigma and device as this object */
        if (!tmp) return inn_p(0.f, 1, 0);
        h_ = tmp;
        upload(CVO_SLOT_MOVING, cloud_a);
        upload(CVO_SLOT_FIXED, cloud_b);
        inn_p r = inner(CVO_SLOT_MOVING, nullptr, CVO_SLOT_FIXED);
        h_ = keep;
        return r;
    }
    matrix66d_t se3_Hessian(point_cloud_t *cloud_a, point_cloud_t *cloud_b, int &inliers) {
        cvo_handle *keep = h_;
        matrix66d_t I;
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) I(r, c) = (r == c);
        cvo_handle *tmp = scratch_handle();
        if (!tmp) return I;
        h_ = tmp;
        upload(CVO_SLOT_MOVING, cloud_a);
        upload(CVO_SLOT_FIXED, cloud_b);
        matrix66d_t H = hess(CVO_SLOT_MOVING, nullptr, CVO_SLOT_FIXED, inliers);
        h_ = keep;
        return H;
    }

    /* cvo.cpp:475-503 */
    void compute_innerproduct(inn_p &inn_pre, inn_p &inn_post, matrix66d_t &post_hessian, affine3f_t &tran,
                              int &inliers, inn_p &inn_fixed_pcd, inn_p &inn_moving_pcd, float &cos_angle) {
        float T[16], v[4];
        int n[4];
        double H[36];
        detail::to_rows(tran, T);
        check(cvo_compute_innerproduct(h_, T, v, n, H, &inliers), "cvo_compute_innerproduct");
        inn_pre.copy(inn_p(v[0], n[0], 0));
        inn_post.copy(inn_p(v[1], n[1], 0));
        inn_fixed_pcd.copy(inn_p(v[2], n[2], 0));
        inn_moving_pcd.copy(inn_p(v[3], n[3], 0));
        cos_angle = inn_post.value / (std::sqrt(inn_fixed_pcd.value) * std::sqrt(inn_moving_pcd.value));
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) post_hessian(r, c) = H[r * 6 + c];
    }

    /* cvo.cpp:505-561 */
    void compute_innerproduct_lc(inn_p &inn_prior, inn_p &inn_lc_prior, inn_p &inn_lc_pre, inn_p &inn_lc_post,
                                 matrix66d_t &post_hessian, affine3f_t &prior_tran, affine3f_t &lc_prior_tran,
                                 affine3f_t &lc_prior_tran_2, affine3f_t &lc_tran, int &inliers_svd,
                                 int &inliers_pnpransac, inn_p &inn_fixed_pcd, inn_p &inn_moving_pcd,
                                 float &cos_angle) {
        float Tp[16], Tlp[16], Tlp2[16], Tl[16];
        detail::to_rows(prior_tran, Tp);
        detail::to_rows(lc_prior_tran, Tlp);
        detail::to_rows(lc_prior_tran_2, Tlp2);
        detail::to_rows(lc_tran, Tl);
        cvo_lc_result r;   /* the six inner products and two Hessians in one launch */
        check(cvo_compute_innerproduct_lc(h_, Tp, Tlp, Tlp2, Tl, &r), "cvo_compute_innerproduct_lc");
        inn_prior.copy(inn_p(r.value[0], r.num[0], 0));
        inn_lc_prior.copy(inn_p(r.value[1], r.num[1], 0));
        inn_lc_pre.copy(inn_p(r.value[2], r.num[2], 0));
        inn_lc_post.copy(inn_p(r.value[3], r.num[3], 0));
        inn_fixed_pcd.copy(inn_p(r.value[4], r.num[4], 0));
        inn_moving_pcd.copy(inn_p(r.value[5], r.num[5], 0));
        cos_angle = inn_lc_post.value / (std::sqrt(inn_fixed_pcd.value) * std::sqrt(inn_moving_pcd.value));
        inliers_svd = r.inliers_svd;
        inliers_pnpransac = r.inliers_pnpransac;
        for (int a = 0; a < 6; a++) for (int c = 0; c < 6; c++) post_hessian(a, c) = r.post_hessian[a * 6 + c];
    }

    /* cvo.cpp:578-618: the unique_ptr moves become slot moves on the device clouds */
    void update_fixed_pcd() { cvo_slot_move(h_, CVO_SLOT_FIXED, CVO_SLOT_MOVING); }
    void update_previous_pcd() {
        cvo_slot_move(h_, CVO_SLOT_PREVIOUS, CVO_SLOT_MOVING);
        pre_pc_init = true;
    }
    void reset_keyframe(affine3f_t &odometry) {
        if (!pre_pc_init) {
            cvo_slot_move(h_, CVO_SLOT_FIXED, CVO_SLOT_MOVING);
        } else {
            cvo_slot_move(h_, CVO_SLOT_FIXED, CVO_SLOT_PREVIOUS);
            update_previous_pcd();
        }
        reset_transform(odometry);
    }
    void reset_transform(affine3f_t &odometry) { transform = odometry; }
    affine3f_t reset_initial(affine3f_t &odometry) {
#ifdef CVO_B200_HAVE_EIGEN_OPENCV
        /* with the real Eigen this IS the reference's code (cvo.cpp:611-618), including rotation(), which
         * re-orthogonalises the linear part through an SVD */
        Eigen::Affine3f init_e = (transform * odometry).inverse();
        const Eigen::Matrix3f Re = init_e.rotation();
        const Eigen::Vector3f Te = init_e.translation();
        float Rr[9], Tr[3];
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) Rr[r * 3 + k] = Re(r, k); Tr[r] = Te(r); }
        cvo_set_RT(h_, Rr, Tr);
        return init_e.inverse();
#else
        /* stand-in types: the same product and cofactor inverse in float; R is the linear part itself where
         * Eigen's rotation() would re-orthogonalise it (a ~1e-7 difference in the prior; the oracle and the
         * Python mirror make the same choice, so the parity tests compare like with like) */
        float a[16], b[16], c[16], init_m[16], back[16];
        detail::to_rows(transform, a);
        detail::to_rows(odometry, b);
        detail::mul44(a, b, c);
        detail::inv_affine(c, init_m); /* init = (transform * odometry).inverse() */
        float R[9], T[3];
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) R[r * 3 + k] = init_m[r * 4 + k]; T[r] = init_m[r * 4 + 3]; }
        cvo_set_RT(h_, R, T);
        detail::inv_affine(init_m, back);
        affine3f_t out = affine3f_t::Identity();
        detail::from_rows(back, out);
        return out;
#endif
    }

    /* getters (cvo.hpp:268-276) */
    void get_fixed_and_moving_number(int &fixed_num, int &moving_num) {
        /* the reference returns what set_pcd cached (cvo.cpp:370-371), also after update_fixed_pcd has moved
         * the clouds; align() refreshes the cache from its result, so this costs a device sync only when it is
         * called between set_pcd and align */
        if (!sizes_valid_ && init) {
            int nf = 0, nm = 0;
            if (cvo_slot_size(h_, CVO_SLOT_FIXED, &nf) != CVO_ERR_NOT_INIT && cvo_slot_size(h_, CVO_SLOT_MOVING, &nm) != CVO_ERR_NOT_INIT) {
                num_fixed = nf;
                num_moving = nm;
                sizes_valid_ = true;
            }
        }
        fixed_num = num_fixed;
        moving_num = num_moving;
    }
    void get_iteration_number(int &iteration) { iteration = iter; }
    void get_A_nonzero(int &nonzero) { nonzero = A_nonzero; }
    void get_fixed_frame_selected_points(std::vector<point2f_t> &pts) { selected(CVO_SLOT_FIXED, pts); }
    void get_moving_frame_selected_points(std::vector<point2f_t> &pts) { selected(CVO_SLOT_MOVING, pts); }

    cvo_handle *native_handle() { return h_; }

private:
    void selected(int slot, std::vector<point2f_t> &pts) {
        int n = 0;
        pts.clear();
        if (cvo_slot_size(h_, slot, &n) != CVO_OK || n <= 0) return;
        std::vector<float> xy((size_t)2 * n);
        if (cvo_get_selected_points(h_, slot, xy.data(), n, &n) != CVO_OK) return;
        pts.reserve(n);
        for (int i = 0; i < n; i++) pts.push_back(point2f_t(xy[2 * i], xy[2 * i + 1]));
    }
};

}  // namespace cvo
#endif  // RKHS_SE3_H
