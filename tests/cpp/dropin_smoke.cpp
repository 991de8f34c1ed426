// Exercises the drop-in cvo::cvo class (include/cvo.hpp) the way LocalTracker does
// (src/local_tracker.cpp:228-251): set_pcd(keyframe), match_odometry(frame), compute_innerproduct.
// usage: dropin_smoke calib.yaml a_bgr.raw a_depth.raw b_bgr.raw b_depth.raw W H
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
    if (argc < 8) return 1;
    const int W = atoi(argv[6]), H = atoi(argv[7]);
    std::vector<unsigned char> a = slurp(argv[2]), ad = slurp(argv[3]), b = slurp(argv[4]), bd = slurp(argv[5]);
    cvo::mat_t A, AD, B, BD;
    A.rows = AD.rows = B.rows = BD.rows = H;
    A.cols = AD.cols = B.cols = BD.cols = W;
    A.data = a.data(); A.step = 3 * W;
    AD.data = ad.data(); AD.step = 2 * W;
    B.data = b.data(); B.step = 3 * W;
    BD.data = bd.data(); BD.step = 2 * W;

    cvo::cvo odo(argv[1]);
    cvo::affine3d_t T;
    odo.match_odometry(A, AD, T);   // prints "cvo not initialized !" and returns, like the reference
    odo.set_pcd(A, AD);
    odo.match_odometry(B, BD, T);
    int nf, nm, nnz, it;
    odo.get_fixed_and_moving_number(nf, nm);
    odo.get_A_nonzero(nnz);
    odo.get_iteration_number(it);
    cvo::inn_p pre, post, fx, mv;
    cvo::matrix66d_t Hm;
    int inliers = 0;
    float cosang = 0;
    cvo::affine3f_t tran = odo.transform;
    odo.compute_innerproduct(pre, post, Hm, tran, inliers, fx, mv, cosang);
    std::vector<cvo::point2f_t> pts;
    odo.get_fixed_frame_selected_points(pts);
    printf("N %d %d nnz %d iter %d inliers %d npts %zu\n", nf, nm, nnz, it, inliers, pts.size());
    printf("T");
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) printf(" %.9g", odo.transform.matrix()(r, c));
    printf("\ninn %.9g %.9g %.9g %.9g cos %.9g H00 %.9g\n", pre.value, post.value, fx.value, mv.value, cosang, Hm(0, 0));
    // the public upload-path wrappers (cvo.cpp:388-459, 620-759) on host clouds: same numbers as the
    // slot-based calls made by compute_innerproduct
    {
        int n = 0;
        cvo::point_cloud_t pcf, pcm;
        for (int slot = 0; slot < 2; slot++) {
            cvo_slot_size(odo.native_handle(), slot, &n);
            std::vector<float> pos(3 * n), feat(5 * n);
            cvo_get_cloud(odo.native_handle(), slot, pos.data(), feat.data(), n, &n);
            cvo::point_cloud_t &pc = slot == 0 ? pcf : pcm;
            pc.num_points = n;
            pc.positions.resize(n);
            pc.features.resize(n);
            for (int i = 0; i < n; i++) {
                for (int k = 0; k < 3; k++) pc.positions[i][k] = pos[3 * i + k];
                for (int k = 0; k < 5; k++) pc.features[i][k] = feat[5 * i + k];
            }
        }
        cvo::inn_p fip = odo.function_inner_product(&pcm, &pcf);
        int inl2 = 0;
        cvo::matrix66d_t H2 = odo.se3_Hessian(&pcm, &pcf, inl2);
        printf("fip %.9g %d pre %.9g %d hess_inliers %d H2_00 %.9g\n", fip.value, fip.num, pre.value, pre.num, inl2, H2(0, 0));
    }
    cvo::affine3f_t back = odo.reset_initial(tran);
    odo.update_fixed_pcd();
    odo.get_fixed_and_moving_number(nf, nm);
    printf("after update_fixed_pcd N %d %d back00 %.6f\n", nf, nm, back.matrix()(0, 0));
    return 0;
}
