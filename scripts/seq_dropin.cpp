// C2 through the drop-in C++ class (include/cvo.hpp): the per-frame call pattern of LocalTracker
// (src/local_tracker.cpp:223-251, 349-431) with two cvo::cvo objects — consecutive-frame odometry and
// keyframe tracking — on a sequence of raw frames, timed on the host like the reference's own loop.
// usage: seq_dropin calib.yaml bgr.raw depth.raw N W H [passes] [dedup]
//   dedup = 1: the keyframe object adopts the odometry object's selection of the same image (cvo::set_pcd_from)
//   bgr.raw = N x H x W x 3 bytes, depth.raw = N x H x W x 2 bytes (uint16).
// prints one line:  frames N passes P ms_per_frame <p0> <p1> ... T_kf_last <16 floats>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cvo.hpp"

static std::vector<unsigned char> slurp(const char *p) {
    FILE *f = fopen(p, "rb");
    if (!f) { perror(p); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> b(n);
    if (fread(b.data(), 1, n, f) != (size_t)n) exit(3);
    fclose(f);
    return b;
}

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s calib.yaml bgr.raw depth.raw N W H [passes]\n", argv[0]); return 1; }
    const int N = atoi(argv[4]), W = atoi(argv[5]), H = atoi(argv[6]);
    const int passes = argc > 7 ? atoi(argv[7]) : 3;
    const bool dedup = argc > 8 && atoi(argv[8]) != 0;
    std::vector<unsigned char> bgr = slurp(argv[2]), dep = slurp(argv[3]);
    if (bgr.size() != (size_t)N * W * H * 3 || dep.size() != (size_t)N * W * H * 2 || N < 3) return 4;
    auto img = [&](int k, cvo::mat_t &c, cvo::mat_t &d) {
        c.rows = d.rows = H;
        c.cols = d.cols = W;
        c.data = bgr.data() + (size_t)k * W * H * 3; c.step = 3 * (size_t)W;
        d.data = dep.data() + (size_t)k * W * H * 2; d.step = 2 * (size_t)W;
    };
    std::vector<double> ms;
    cvo::affine3d_t T_last;
    for (int pass = 0; pass < passes; pass++) {
        const auto t0 = std::chrono::steady_clock::now();
        cvo::cvo odo(argv[1]), kf(argv[1]);
        cvo::mat_t c, d;
        cvo::inn_p pre, post, fx, mv;
        cvo::matrix66d_t Hm;
        int inliers = 0;
        float cosang = 0;
        // initNewLocalMap (local_tracker.cpp:223-345)
        img(0, c, d);
        odo.set_pcd(c, d);
        if (dedup) kf.set_pcd_from(odo, CVO_SLOT_FIXED);
        else kf.set_pcd(c, d);
        cvo::affine3d_t T;
        img(1, c, d);
        odo.match_odometry(c, d, T);
        cvo::affine3f_t Tf = T.cast<float>();
        odo.compute_innerproduct(pre, post, Hm, Tf, inliers, fx, mv, cosang);
        kf.first_frame = false;
        kf.reset_transform(Tf);
        odo.update_fixed_pcd();
        for (int k = 2; k < N; k++) {
            img(k, c, d);
            cvo::affine3d_t T_odo, T_kf;
            odo.match_odometry(c, d, T_odo);
            cvo::affine3f_t To = T_odo.cast<float>();
            odo.compute_innerproduct(pre, post, Hm, To, inliers, fx, mv, cosang);
            odo.update_fixed_pcd();
            kf.reset_initial(To);
            if (dedup) kf.match_keyframe_from(odo, CVO_SLOT_FIXED, T_kf);
            else kf.match_keyframe(c, d, T_kf);
            cvo::affine3f_t Tk = T_kf.cast<float>();
            kf.compute_innerproduct(pre, post, Hm, Tk, inliers, fx, mv, cosang);
            kf.update_previous_pcd();   // accepted frame (local_tracker.cpp:506)
            T_last = T_kf;
        }
        const auto t1 = std::chrono::steady_clock::now();
        ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count() / (N - 1));
    }
    printf("frames %d passes %d ms_per_frame", N, passes);
    for (double v : ms) printf(" %.4f", v);
    printf(" T_kf_last");
    for (int r = 0; r < 4; r++) for (int cc = 0; cc < 4; cc++) printf(" %.9g", (double)T_last.matrix()(r, cc));
    printf("\n");
    return 0;
}
