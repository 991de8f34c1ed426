// oracle/ref_select.cpp — drives the REFERENCE's own point selection (thirdparty/cvo/src/pcd_generator.cpp and
// thirdparty/cvo/thirdparty/PixelSelector2.cpp, compiled where they lie under /root/reference; see the `refsel`
// target of oracle/Makefile) so that the oracle's restatement of SURVEY §8a rows A-H can be pinned against
// outputs of the reference itself: status map, selected pixels, positions, features.
//
// TEST INFRASTRUCTURE.  Eigen, OpenCV and TBB are not in this image: the reference sources are compiled
// against the stand-in headers of oracle/shim/ (a dense matrix class, a plain-buffer cv::Mat, the two
// cvtColor conversions in OpenCV's 8-bit integer arithmetic).  What is pinned is therefore everything the
// reference's own files compute — pyramid, gradients, histograms, thresholds, the three-level selection walk,
// the recursion, the rand()-pattern sub-sampling, back-projection, features — with cvtColor pinned separately
// against cv2 (tests/test_oracle_pins.py).
#include <algorithm>
#include <cstdint>
#include <cstring>

// every standard / stand-in header first, so that the access trick below touches the reference's headers only
#include <complex>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
#include <tbb/concurrent_vector.h>

#define private public   // pcd_generator::map / num_want are private; this file only reads / sets them
#include "pcd_generator.hpp"
#undef private

static int g_gray_mode = 0;

extern "C" void oracle_gray_u8(const uint8_t *p, int n, uint8_t *out) {
    for (int i = 0; i < n; i++) {
        const int c0 = p[3 * i], c1 = p[3 * i + 1], c2 = p[3 * i + 2];
        out[i] = g_gray_mode == 1 ? (uint8_t)((c0 * 4899 + c1 * 9617 + c2 * 1868 + 8192) >> 14)     // OpenCV 3.x
                                  : (uint8_t)((c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15);  // OpenCV >= 4
    }
}
extern "C" void oracle_hsv_u8(const uint8_t *p, int n, uint8_t *out) {
    static int sdiv[256], hdiv[256], init = 0;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = (int)std::lrint((255 << 12) / (double)i);
            hdiv[i] = (int)std::lrint((180 << 12) / (6.0 * i));
        }
        init = 1;
    }
    for (int i = 0; i < n; i++) {
        const int r = p[3 * i], g = p[3 * i + 1], b = p[3 * i + 2];
        const int v = std::max(r, std::max(g, b)), vmin = std::min(r, std::min(g, b)), diff = v - vmin;
        const int vr = (v == r) ? -1 : 0, vg = (v == g) ? -1 : 0;
        const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
        int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
        h = (h * hdiv[diff] + (1 << 11)) >> 12;
        if (h < 0) h += 180;
        out[3 * i] = (uint8_t)h; out[3 * i + 1] = (uint8_t)s; out[3 * i + 2] = (uint8_t)v;
    }
}

extern "C" {

// calib = {scaling_factor, fx, fy, cx, cy}.  Returns the number of points, or -1 if cap is too small.
// map_out: w*h status map (0/1/2/4 after sub-sampling); pix: n x 2; pos: n x 3; feat: n x 5 (row-major).
int refsel_run(const uint8_t *bgr, const uint16_t *depth, int w, int h, const float calib[5], int num_want,
               int feature_type, int gray_mode, float *map_out, float *pix, float *pos, float *feat, int cap,
               float *gray_out /* w*h or null */) {
    g_gray_mode = gray_mode;
    cv::Mat img(h, w, CV_8UC3, (void *)bgr, (size_t)w * 3), dep(h, w, CV_16UC1, (void *)depth, (size_t)w * 2);
    cvo::camera_info cam{calib[0], calib[1], calib[2], calib[3], calib[4]};
    cvo::frame fr;
    fr.avg_abs_squared_grad = 0.f;   // (the reference accumulates into it uninitialised and never reads it)
    cvo::point_cloud pc;
    int n;
    {
        cvo::pcd_generator gen;
        gen.set_calib(cam);
        gen.num_want = num_want;
        gen.load_image(img, dep, &fr);
        gen.create_pointcloud(feature_type, &fr, &pc);
        n = pc.num_points;
        if (map_out) memcpy(map_out, gen.map, sizeof(float) * (size_t)w * h);
        if (gray_out)
            for (int i = 0; i < w * h; i++) gray_out[i] = fr.dI[i][0];
    }
    if (n > cap) return -1;
    for (int i = 0; i < n; i++) {
        pix[2 * i] = fr.selected_points[i].x; pix[2 * i + 1] = fr.selected_points[i].y;
        for (int k = 0; k < 3; k++) pos[3 * i + k] = pc.positions[i][k];
        for (int k = 0; k < 5; k++) feat[5 * i + k] = pc.features(i, k);
    }
    return n;
}

}  // extern "C"
