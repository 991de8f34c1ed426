// oracle/cvo_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the CVO frame-pair alignment path of
// bexilin/CVO-SLAM (thirdparty/cvo).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product (libcvo_b200.so)
// never does.
//
// PARITY STATUS: "parity unpinned" by the reference itself — the reference ships no tests,
// golden vectors or fixtures and cannot be compiled here (needs Eigen, OpenCV C++,
// TBB, icpc).  What pins this oracle instead (tests/test_oracle_*.py):
//   * gray / HSV vs the in-container cv2 4.13 (the arithmetic the reference delegates to OpenCV)
//   * randomPattern vs this libc's srand/rand (the reference calls libc)
//   * in-cutoff pattern vs the reference's OWN vendored nanoflann.hpp (oracle/_ref build)
//   * cubic roots vs numpy.roots, Exp_SEK3 / dist_se3 vs scipy expm / logm
//   * convergence to an injected SE(3) on synthetic RGB-D
//
// Every function cites the reference lines it follows (paths relative to the reference root).
// Floating point: compiled with -ffp-contract=off so each float operation rounds once, in the
// order written in the reference source (the reference's own icpc build may contract or
// reassociate; that is outside what source can pin and is documented in DESIGN.md).

#include "../include/cvo_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_NANOFLANN
// The reference's own vendored KD-tree (STL-only), included from /root/reference at build
// time; only built into oracle/_ref/ (see oracle/Makefile).
#include "KDTreeVectorOfVectorsAdaptor.h"
#include "nanoflann.hpp"
#endif

namespace {

typedef std::array<float, 3> V3;
typedef std::array<float, 5> F5;

struct Cloud {
    int n = 0;
    std::vector<V3> pos;      // data_type.h:71
    std::vector<F5> feat;     // data_type.h:72 (N x 5)
    std::vector<float> pix;   // frame::selected_points (x,y) data_type.h:59
    // selection internals kept for the bit-exactness tests
    std::vector<uint8_t> map; // status map after sub-sampling (0/1/2/4)
    int info[5] = {0, 0, 0, 0, 0};
    bool valid = false;
};

// ---------------------------------------------------------------------------------------------
// Row A: cv::cvtColor RGB2GRAY / RGB2HSV as called from pcd_generator::load_image
// (thirdparty/cvo/src/pcd_generator.cpp:624-625).  The image is BGR (cv::imread,
// src/run_SLAM.cpp:137) but converted with the RGB code, so stored channel 0 takes the R weight.
// gray_mode 0: OpenCV >= 4 (15-bit fixed point), gray_mode 1: OpenCV 3.x (14-bit).
// ---------------------------------------------------------------------------------------------
inline uint8_t gray_px(int c0, int c1, int c2, int mode) {
    if (mode == 1) return (uint8_t)((c0 * 4899 + c1 * 9617 + c2 * 1868 + 8192) >> 14);
    return (uint8_t)((c0 * 9798 + c1 * 19235 + c2 * 3735 + 16384) >> 15);
}

struct HsvTables {
    int sdiv[256], hdiv[256];
    HsvTables() {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = (int)std::lrint((255 << 12) / (double)i);
            hdiv[i] = (int)std::lrint((180 << 12) / (6.0 * i));
        }
    }
};

// OpenCV 8-bit RGB2HSV (H in [0,180)); c0 plays "R" because of the BGR/RGB mix-up above.
inline void hsv_px(int r, int g, int b, uint8_t out[3]) {
    static const HsvTables t;
    int v = std::max(r, std::max(g, b));
    int vmin = std::min(r, std::min(g, b));
    int diff = v - vmin;
    int vr = (v == r) ? -1 : 0, vg = (v == g) ? -1 : 0;
    int s = (diff * t.sdiv[v] + (1 << 11)) >> 12;
    int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * t.hdiv[diff] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    out[0] = (uint8_t)h;
    out[1] = (uint8_t)s;
    out[2] = (uint8_t)v;
}

// ---------------------------------------------------------------------------------------------
// Row B: pcd_generator::make_pyramid (pcd_generator.cpp:50-143)
// ---------------------------------------------------------------------------------------------
struct Pyramid {
    int w[3], h[3];
    std::vector<float> I[3], dx[3], dy[3], g2[3];
};

void make_pyramid(const uint8_t *gray, int w, int h, Pyramid &p) {
    int wl = w, hl = h;
    for (int l = 0; l < 3; l++) {  // :60-71 (zero-initialised arrays)
        p.w[l] = wl;
        p.h[l] = hl;
        p.I[l].assign((size_t)wl * hl, 0.f);
        p.dx[l].assign((size_t)wl * hl, 0.f);
        p.dy[l].assign((size_t)wl * hl, 0.f);
        p.g2[l].assign((size_t)wl * hl, 0.f);
        wl /= 2;
        hl /= 2;
    }
    for (int i = 0; i < h; i++)  // :80-84
        for (int j = 0; j < w; j++) p.I[0][(size_t)i * w + j] = gray[(size_t)i * w + j];
    wl = w;
    hl = h;
    for (int lvl = 0; lvl < 3; lvl++) {
        std::vector<float> &I = p.I[lvl];
        if (lvl > 0) {
            // :100-115 — NOTE prev_wl = wl*2, not the true previous width (wrong for odd widths;
            // replicated, not fixed).
            int prev_wl = wl * 2;
            const std::vector<float> &P = p.I[lvl - 1];
            for (int y = 0; y < hl; y++)
                for (int x = 0; x < wl; x++) {
                    size_t b = (size_t)2 * x + (size_t)2 * y * prev_wl;
                    float s = P[b] + P[b + 1];  // left-to-right, as written at :109-112
                    s = s + P[b + prev_wl];
                    s = s + P[b + 1 + prev_wl];
                    I[(size_t)x + (size_t)y * wl] = 0.25f * s;
                }
        }
        for (int idx = wl; idx < wl * (hl - 1); idx++) {  // :119-135
            float dx = 0.5f * (I[idx + 1] - I[idx - 1]);
            float dy = 0.5f * (I[idx + wl] - I[idx - wl]);
            if (!std::isfinite(dx)) dx = 0;
            if (!std::isfinite(dy)) dy = 0;
            p.dx[lvl][idx] = dx;
            p.dy[lvl][idx] = dy;
            float a = dx * dx;
            float b = dy * dy;
            p.g2[lvl][idx] = a + b;
        }
        wl /= 2;
        hl /= 2;
    }
}

// ---------------------------------------------------------------------------------------------
// Rows C-F: dso::PixelSelector (thirdparty/cvo/thirdparty/PixelSelector2.cpp)
// ---------------------------------------------------------------------------------------------
struct PixelSelector {
    int w, h;
    std::vector<uint8_t> randomPattern;
    int currentPotential;
    std::vector<float> ths, thsSmoothed;
    int thsStep = 0;
    int last_n[3] = {0, 0, 0};
    int last_pot = 0;
    int passes = 0;

    // PixelSelector2.cpp:34-49
    PixelSelector(int w_, int h_) : w(w_), h(h_) {
        randomPattern.resize((size_t)w * h);
        std::srand(3141592);
        for (int i = 0; i < w * h; i++) randomPattern[i] = rand() & 0xFF;
        currentPotential = 3;
        ths.assign((size_t)(w / 32) * (h / 32) + 100, 0.f);
        thsSmoothed.assign((size_t)(w / 32) * (h / 32) + 100, 0.f);
    }

    // :59-68
    static int computeHistQuantil(const int *hist, float below) {
        int th = hist[0] * below + 0.5f;
        for (int i = 0; i < 90; i++) {
            th -= hist[i + 1];
            if (th < 0) return i;
        }
        return 90;
    }

    // :71-136
    void makeHists(const Pyramid &p) {
        const float *mapmax0 = p.g2[0].data();
        int w32 = w / 32, h32 = h / 32;
        thsStep = w32;
        int hist0[100];
        for (int y = 0; y < h32; y++)
            for (int x = 0; x < w32; x++) {
                const float *map0 = mapmax0 + 32 * x + 32 * y * w;
                memset(hist0, 0, sizeof(int) * 100);  // reference clears 50; bins >49 never hit
                for (int j = 0; j < 32; j++)
                    for (int i = 0; i < 32; i++) {
                        int it = i + 32 * x, jt = j + 32 * y;
                        if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
                        int g = sqrtf(map0[i + j * w]);
                        if (g > 48) g = 48;
                        hist0[g + 1]++;
                        hist0[0]++;
                    }
                ths[x + y * w32] = computeHistQuantil(hist0, 0.5f) + 7;
            }
        for (int y = 0; y < h32; y++)
            for (int x = 0; x < w32; x++) {
                float sum = 0, num = 0;
                if (x > 0) {
                    if (y > 0) { num++; sum += ths[x - 1 + (y - 1) * w32]; }
                    if (y < h32 - 1) { num++; sum += ths[x - 1 + (y + 1) * w32]; }
                    num++; sum += ths[x - 1 + (y)*w32];
                }
                if (x < w32 - 1) {
                    if (y > 0) { num++; sum += ths[x + 1 + (y - 1) * w32]; }
                    if (y < h32 - 1) { num++; sum += ths[x + 1 + (y + 1) * w32]; }
                    num++; sum += ths[x + 1 + (y)*w32];
                }
                if (y > 0) { num++; sum += ths[x + (y - 1) * w32]; }
                if (y < h32 - 1) { num++; sum += ths[x + (y + 1) * w32]; }
                num++; sum += ths[x + y * w32];
                thsSmoothed[x + y * w32] = (sum / num) * (sum / num);
            }
    }

    // :290-433, literal sequential restatement.  setting_selectDirectionDistribution is false
    // (PixelSelector2.h:31) so dirNorm is the gradient magnitude itself and the random
    // directions are dead.
    void select(const Pyramid &p, float *map_out, int pot, float thFactor, int n_out[3]) {
        const float *mapmax0 = p.g2[0].data();
        const float *mapmax1 = p.g2[1].data();
        const float *mapmax2 = p.g2[2].data();
        int w1 = w / 2, w2 = w / 4;
        memset(map_out, 0, (size_t)w * h * sizeof(float));
        float dw1 = 0.75;
        float dw2 = dw1 * dw1;
        int n3 = 0, n2 = 0, n4 = 0;
        for (int y4 = 0; y4 < h; y4 += (4 * pot))
            for (int x4 = 0; x4 < w; x4 += (4 * pot)) {
                int my3 = std::min((4 * pot), h - y4);
                int mx3 = std::min((4 * pot), w - x4);
                int bestIdx4 = -1;
                float bestVal4 = 0;
                for (int y3 = 0; y3 < my3; y3 += (2 * pot))
                    for (int x3 = 0; x3 < mx3; x3 += (2 * pot)) {
                        int x34 = x3 + x4, y34 = y3 + y4;
                        int my2 = std::min((2 * pot), h - y34);
                        int mx2 = std::min((2 * pot), w - x34);
                        int bestIdx3 = -1;
                        float bestVal3 = 0;
                        for (int y2 = 0; y2 < my2; y2 += pot)
                            for (int x2 = 0; x2 < mx2; x2 += pot) {
                                int x234 = x2 + x34, y234 = y2 + y34;
                                int my1 = std::min(pot, h - y234);
                                int mx1 = std::min(pot, w - x234);
                                int bestIdx2 = -1;
                                float bestVal2 = 0;
                                for (int y1 = 0; y1 < my1; y1 += 1)
                                    for (int x1 = 0; x1 < mx1; x1 += 1) {
                                        int idx = x1 + x234 + w * (y1 + y234);
                                        int xf = x1 + x234, yf = y1 + y234;
                                        if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                                        float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * thsStep];
                                        float pixelTH1 = pixelTH0 * dw1;
                                        float pixelTH2 = pixelTH1 * dw2;
                                        float ag0 = mapmax0[idx];
                                        if (ag0 > pixelTH0 * thFactor) {
                                            float dirNorm = ag0;
                                            if (dirNorm > bestVal2) {
                                                bestVal2 = dirNorm; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2;
                                            }
                                        }
                                        if (bestIdx3 == -2) continue;
                                        float ag1 = mapmax1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * w1];
                                        if (ag1 > pixelTH1 * thFactor) {
                                            float dirNorm = ag1;
                                            if (dirNorm > bestVal3) {
                                                bestVal3 = dirNorm; bestIdx3 = idx; bestIdx4 = -2;
                                            }
                                        }
                                        if (bestIdx4 == -2) continue;
                                        float ag2 = mapmax2[(int)(xf * 0.25f + 0.125) + (int)(yf * 0.25f + 0.125) * w2];
                                        if (ag2 > pixelTH2 * thFactor) {
                                            float dirNorm = ag2;
                                            if (dirNorm > bestVal4) { bestVal4 = dirNorm; bestIdx4 = idx; }
                                        }
                                    }
                                if (bestIdx2 > 0) { map_out[bestIdx2] = 1; bestVal3 = 1e10; n2++; }
                            }
                        if (bestIdx3 > 0) { map_out[bestIdx3] = 2; bestVal4 = 1e10; n3++; }
                    }
                if (bestIdx4 > 0) { map_out[bestIdx4] = 4; n4++; }
            }
        n_out[0] = n2; n_out[1] = n3; n_out[2] = n4;
        last_n[0] = n2; last_n[1] = n3; last_n[2] = n4;
        last_pot = pot;
        passes++;
    }

    // :137-286 (FAST branch commented out in the reference)
    int makeMaps(const Pyramid &p, float *map_out, float density, int recursionsLeft = 1,
                 float thFactor = 1) {
        float numHave = 0, numWant = density, quotia;
        int idealPotential = currentPotential;
        if (passes == 0) makeHists(p);  // `ptr_fr != gradHistFrame`: true once per selector
        int n[3];
        select(p, map_out, currentPotential, thFactor, n);
        numHave = n[0] + n[1] + n[2];
        quotia = numWant / numHave;
        float K = numHave * (currentPotential + 1) * (currentPotential + 1);
        idealPotential = sqrtf(K / numWant) - 1;
        if (idealPotential < 1) idealPotential = 1;
        if (recursionsLeft > 0 && quotia > 1.25 && currentPotential > 1) {
            if (idealPotential >= currentPotential) idealPotential = currentPotential - 1;
            currentPotential = idealPotential;
            return makeMaps(p, map_out, density, recursionsLeft - 1, thFactor);
        } else if (recursionsLeft > 0 && quotia < 0.25) {
            if (idealPotential <= currentPotential) idealPotential = currentPotential + 1;
            currentPotential = idealPotential;
            return makeMaps(p, map_out, density, recursionsLeft - 1, thFactor);
        }
        int numHaveSub = numHave;
        if (quotia < 0.95) {
            int wh = w * h;
            int rn = 0;
            unsigned char charTH = 255 * quotia;
            for (int i = 0; i < wh; i++) {
                if (map_out[i] != 0) {
                    if (randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; }
                    rn++;
                }
            }
        }
        currentPotential = idealPotential;
        return numHaveSub;
    }
};

// Stage dumps for tests.
struct StageDump {
    std::vector<uint8_t> gray;
    Pyramid pyr;
    std::vector<float> ths, thsSmoothed;
};

// Rows A-H: pcd_generator::load_image + create_pointcloud (pcd_generator.cpp:618-656)
void create_pointcloud(const uint8_t *bgr, size_t bgr_stride, const uint16_t *depth,
                       size_t depth_stride, int w, int h, const cvo_calib &cal,
                       const cvo_params &prm, Cloud &out, StageDump *dump) {
    std::vector<uint8_t> gray((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *row = bgr + (size_t)y * bgr_stride;
        for (int x = 0; x < w; x++)
            gray[(size_t)y * w + x] = gray_px(row[3 * x], row[3 * x + 1], row[3 * x + 2], prm.gray_mode);
    }
    Pyramid pyr;
    make_pyramid(gray.data(), w, h, pyr);                       // select_point :150
    std::vector<float> map((size_t)w * h);                      // :152
    PixelSelector sel(w, h);                                    // :154
    int num_selected = sel.makeMaps(pyr, map.data(), (float)prm.num_want);  // :155
    (void)num_selected;

    out = Cloud();
    out.map.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) out.map[i] = (uint8_t)map[i];
    out.info[0] = sel.last_n[0]; out.info[1] = sel.last_n[1]; out.info[2] = sel.last_n[2];
    out.info[3] = sel.last_pot;  out.info[4] = sel.passes;

    // get_points_from_pixels :456-499 and get_features :563-616 (same raster scan, same filter)
    for (int y = 0; y < h; y++) {
        const uint16_t *drow = (const uint16_t *)((const uint8_t *)depth + (size_t)y * depth_stride);
        const uint8_t *row = bgr + (size_t)y * bgr_stride;
        for (int x = 0; x < w; x++) {
            uint16_t dep = drow[x];
            if (map[(size_t)y * w + x] != 0 && dep != 0) {
                V3 p;
                p[2] = dep / cal.scaling_factor;            // :473
                p[0] = (x - cal.cx) * p[2] / cal.fx;        // :475
                p[1] = (y - cal.cy) * p[2] / cal.fy;        // :476
                out.pos.push_back(p);
                out.pix.push_back((float)x);
                out.pix.push_back((float)y);
                F5 f;
                size_t idx = (size_t)y * w + x;
                if (prm.feature_type == 0) {                // :570-592
                    uint8_t hsv[3];
                    hsv_px(row[3 * x], row[3 * x + 1], row[3 * x + 2], hsv);
                    f[0] = hsv[0] / 180.0;
                    f[1] = hsv[1] / 255.0;
                    f[2] = hsv[2] / 255.0;
                    f[3] = pyr.dx[0][idx] / 255.0 * 2;
                    f[4] = pyr.dy[0][idx] / 255.0 * 2;
                } else {                                    // :593-615
                    f[0] = row[3 * x];
                    f[1] = row[3 * x + 1];
                    f[2] = row[3 * x + 2];
                    f[3] = pyr.dx[0][idx];
                    f[4] = pyr.dy[0][idx];
                }
                out.feat.push_back(f);
            }
        }
    }
    out.n = (int)out.pos.size();
    out.valid = true;
    if (dump) {
        dump->gray = gray;
        dump->pyr = pyr;
        dump->ths = sel.ths;
        dump->thsSmoothed = sel.thsSmoothed;
    }
}

// ---------------------------------------------------------------------------------------------
// Order-independent summation.  The reference forms each row's flow contribution as an Eigen
// float dot product `(1/c*Ai)*cross_xy` (cvo.cpp:222-223) whose summation order is an
// implementation detail of Eigen's vectorised redux, and adds the row results — and the row sums
// Bi..Ei of compute_step_size — into doubles under a spin lock in TBB scheduling order
// (cvo.cpp:226-230, 309-314): its omega, v, B..E are only defined up to that order (~1e-7 and
// ~1e-16 relative).  The oracle takes the canonical representative of all admissible orders:
// the exact sum of the terms, rounded once.
//   * flow terms are products of two floats (exact in double); they are summed exactly in a
//     two-limb fixed-point accumulator (value = hi*2^-36 + lo*2^-84), an associative integer sum;
//   * B..E terms are doubles; they are summed in double-double (error ~1e-32 relative), whose
//     rounding to double is the correctly rounded exact sum for all practical purposes.
// Being order-free, both can be reproduced bit for bit by any parallel decomposition on the GPU.
// ---------------------------------------------------------------------------------------------
struct ExactAcc {
    long long hi = 0, lo = 0;
    inline void add(double t) {
        double h = std::nearbyint(t * 0x1p36);
        double r = t - h * 0x1p-36;          // exact
        hi += (long long)h;
        lo += (long long)std::nearbyint(r * 0x1p84);
    }
    inline void add(const ExactAcc &o) { hi += o.hi; lo += o.lo; }
    inline double value() const { return (double)hi * 0x1p-36 + (double)lo * 0x1p-84; }
};

struct DDAcc {   // double-double running sum (TwoSum + low-order accumulation)
    double hi = 0, lo = 0;
    inline void add(double t) {
        double a = hi;
        double sum = a + t;
        double bb = sum - a;
        double err = (a - (sum - bb)) + (t - bb);
        hi = sum;
        lo += err;
    }
    inline void add(const DDAcc &o) { add(o.hi); lo += o.lo; }
    inline double value() const { return hi + lo; }
};

// sin/cos of a float argument, correctly rounded to float (the reference calls std::sin(float)
// of whatever libm it links, LieGroup.cpp:174-175; the correctly rounded value is the canonical one)
inline float sin_cr(float x) { return (float)std::sin((double)x); }
inline float cos_cr(float x) { return (float)std::cos((double)x); }

// ---------------------------------------------------------------------------------------------
// small float 3x3 helpers (Eigen coefficient-wise products: ((a0*b0 + a1*b1) + a2*b2))
// ---------------------------------------------------------------------------------------------
struct M3 { float m[3][3]; };
inline M3 m3_identity() { M3 r{}; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.f; return r; }
inline M3 m3_mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = a.m[i][0] * b.m[0][j];
            s = s + a.m[i][1] * b.m[1][j];
            s = s + a.m[i][2] * b.m[2][j];
            r.m[i][j] = s;
        }
    return r;
}
inline V3 m3_vec(const M3 &a, const V3 &v) {
    V3 r;
    for (int i = 0; i < 3; i++) {
        float s = a.m[i][0] * v[0];
        s = s + a.m[i][1] * v[1];
        s = s + a.m[i][2] * v[2];
        r[i] = s;
    }
    return r;
}
inline M3 m3_T(const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
inline M3 m3_scale(const M3 &a, float s) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s; return r; }
inline M3 m3_add(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
// LieGroup.cpp:20-27
inline M3 skew(const V3 &v) {
    M3 r{};
    r.m[0][1] = -v[2]; r.m[0][2] = v[1];
    r.m[1][0] = v[2];  r.m[1][2] = -v[0];
    r.m[2][0] = -v[1]; r.m[2][1] = v[0];
    return r;
}
inline V3 cross(const V3 &a, const V3 &b) {
    return V3{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline float dot3(const V3 &a, const V3 &b) { float s = a[0] * b[0]; s = s + a[1] * b[1]; s = s + a[2] * b[2]; return s; }
inline float norm3(const V3 &a) { return std::sqrt(dot3(a, a)); }

// LieGroup.cpp:159-186 (K = 1).  NOTE theta < TOLERANCE gives Jl = I, not dt*I (replicated).
void Exp_SEK3(const V3 &w, const V3 &v, float dt, M3 &R, V3 &dT) {
    float theta = norm3(w);
    M3 I = m3_identity(), Jl;
    if (theta < 1e-6f) {
        R = I;
        Jl = I;
    } else {
        M3 A = skew(w);
        float theta2 = theta * theta;
        float stheta = sin_cr(dt * theta);
        float ctheta = cos_cr(dt * theta);
        float oneMinusCosTheta2 = (1 - ctheta) / (theta2);
        M3 A2 = m3_mul(A, A);
        R = m3_add(m3_add(I, m3_scale(A, stheta / theta)), m3_scale(A2, oneMinusCosTheta2));
        Jl = m3_add(m3_add(m3_scale(I, dt), m3_scale(A, oneMinusCosTheta2)),
                    m3_scale(A2, (dt * theta - stheta) / (theta2 * theta)));
    }
    dT = m3_vec(Jl, v);
}

// cvo.cpp:94-104: || logm([R T; 0 1]) ||_F.  The reference calls Eigen's unsupported
// MatrixFunctions (absent here); restated with the closed-form SE(3) logarithm, evaluated in
// double on the float inputs.
float dist_se3(const M3 &R, const V3 &T) {
    double r[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = R.m[i][j];
    double ax = 0.5 * (r[2][1] - r[1][2]), ay = 0.5 * (r[0][2] - r[2][0]), az = 0.5 * (r[1][0] - r[0][1]);
    double s = std::sqrt(ax * ax + ay * ay + az * az);
    double c = 0.5 * (r[0][0] + r[1][1] + r[2][2] - 1.0);
    double theta = std::atan2(s, c);
    double wx, wy, wz;
    if (s < 1e-12) { wx = ax; wy = ay; wz = az; }
    else { double k = theta / s; wx = ax * k; wy = ay * k; wz = az * k; }
    double t[3] = {T[0], T[1], T[2]};
    // u = V^-1 t,  V^-1 = I - 1/2 W + coef W^2
    double coef;
    if (theta < 1e-4) coef = 1.0 / 12.0;
    else coef = (1.0 - theta * std::sin(theta) / (2.0 * (1.0 - std::cos(theta)))) / (theta * theta);
    double wt[3] = {wy * t[2] - wz * t[1], wz * t[0] - wx * t[2], wx * t[1] - wy * t[0]};
    double wwt[3] = {wy * wt[2] - wz * wt[1], wz * wt[0] - wx * wt[2], wx * wt[1] - wy * wt[0]};
    double u[3];
    for (int i = 0; i < 3; i++) u[i] = t[i] - 0.5 * wt[i] + coef * wwt[i];
    double f2 = 2.0 * theta * theta + u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    return (float)std::sqrt(f2);
}

// cvo.cpp:76-92 + :317-333: roots of p0 t^3 + p1 t^2 + p2 t + p3 via the eigenvalues of the
// companion matrix (Eigen general eigen-solver, float).  Restated with the closed-form cubic in
// double + Newton polish.  Returns the number of REAL roots written to `re` (complex pairs of the
// eigen-solver carry imag != 0 and are rejected at :326).  A zero / non-finite leading
// coefficient makes the companion matrix inf/NaN -> no admissible root.
int cubic_real_roots(double a, double b, double c, double d, double re[3]) {
    if (!(std::isfinite(a) && std::isfinite(b) && std::isfinite(c) && std::isfinite(d)) || a == 0.0)
        return 0;
    const double A = b / a, B = c / a, C = d / a;  // monic t^3 + A t^2 + B t + C
    if (!(std::isfinite(A) && std::isfinite(B) && std::isfinite(C))) return 0;
    auto polish = [&](double t) {
        for (int it = 0; it < 4; it++) {
            double f = ((t + A) * t + B) * t + C;
            double fp = (3.0 * t + 2.0 * A) * t + B;
            if (fp == 0.0 || !std::isfinite(fp)) break;
            double tn = t - f / fp;
            if (!std::isfinite(tn)) break;
            t = tn;
        }
        return t;
    };
    // 1) one real root from the closed form (the candidate of largest magnitude), polished
    double sq = A * A;
    double p = (3.0 * B - sq) / 3.0;
    double q = (2.0 * A * sq - 9.0 * A * B + 27.0 * C) / 27.0;
    double disc = q * q / 4.0 + p * p * p / 27.0;
    double r;
    if (disc > 0) {
        double sd = std::sqrt(disc);
        r = std::cbrt(-q / 2.0 + sd) + std::cbrt(-q / 2.0 - sd) - A / 3.0;
    } else if (p == 0.0) {
        r = -A / 3.0;
    } else {
        double m = 2.0 * std::sqrt(-p / 3.0);
        double arg = std::max(-1.0, std::min(1.0, 3.0 * q / (p * m)));
        double th = std::acos(arg) / 3.0;
        const double two_pi_3 = 2.0943951023931954923;
        r = 0;
        for (int k = 0; k < 3; k++) {
            double cand = m * std::cos(th - two_pi_3 * k) - A / 3.0;
            if (std::fabs(cand) >= std::fabs(r)) r = cand;
        }
    }
    r = polish(r);
    // 2) deflate (t - r)(t^2 + b1 t + b0): backward when r is a large root, forward otherwise
    double b1, b0;
    if (r != 0.0 && std::fabs(r * r * r) >= std::fabs(C)) {
        b0 = -C / r;
        b1 = (b0 - B) / r;
    } else {
        b1 = A + r;
        b0 = B + r * b1;
    }
    // 3) the remaining pair: real iff the quadratic's discriminant is non-negative
    int n = 0;
    re[n++] = r;
    double d2 = b1 * b1 - 4.0 * b0;
    if (d2 >= 0) {
        double qq = -0.5 * (b1 + (b1 >= 0 ? 1.0 : -1.0) * std::sqrt(d2));
        double r2 = qq, r3 = (qq != 0.0) ? b0 / qq : 0.0;
        re[n++] = polish(r2);
        re[n++] = polish(r3);
    }
    return n;
}

// Symmetric eigenvalues (cyclic Jacobi, double).  Stands in for Eigen's general
// `.eigenvalues()` real parts at cvo.cpp:728 on a matrix that is symmetric by construction.
void sym_eigenvalues6(const double Hin[36], double ev[6]) {
    double a[6][6];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) a[i][j] = 0.5 * (Hin[i * 6 + j] + Hin[j * 6 + i]);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < 6; i++) for (int j = i + 1; j < 6; j++) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 6; p++)
            for (int q = p + 1; q < 6; q++) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 6; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 6; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
            }
    }
    for (int i = 0; i < 6; i++) ev[i] = a[i][i];
}

// ---------------------------------------------------------------------------------------------
// Neighbour search: the reference uses nanoflann (cvo.cpp:133-148): exact radius search on the
// squared distance ((dx*dx + dy*dy) + dz*dz) (nanoflann.hpp:403-406 with dim 3), strict `<`.
// Three interchangeable back ends producing the same (idx, d2) sets:
//   0 = uniform cell list (default), 1 = brute force, 2 = the reference's nanoflann (only in the
//   oracle/_ref build).
// ---------------------------------------------------------------------------------------------
inline float dist2(const V3 &q, const V3 &p) {
    float d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
    float r = d0 * d0;
    r = r + d1 * d1;
    r = r + d2 * d2;
    return r;
}

struct Searcher {
    const std::vector<V3> *pts = nullptr;
    int mode = 0;
    // cell list
    float h = 0;
    V3 lo{};
    int nx = 0, ny = 0, nz = 0;
    std::vector<int> cell_start, order;
#ifdef ORACLE_NANOFLANN
    typedef KDTreeVectorOfVectorsAdaptor<std::vector<V3>, float> kd_tree_t;
    std::unique_ptr<kd_tree_t> kd;
#endif

    void build(const std::vector<V3> &p, float radius2, int mode_) {
        pts = &p;
        mode = mode_;
        if (mode == 2) {
#ifdef ORACLE_NANOFLANN
            // cvo.cpp:135-136: the adaptor ctor builds the index and buildIndex() builds it again
            kd.reset(new kd_tree_t(3, p, 10));
            kd->index->buildIndex();
            return;
#else
            mode = 0;
#endif
        }
        if (mode == 1 || p.empty()) return;
        float r = std::sqrt(radius2) * 1.0001f + 1e-6f;
        V3 hi;
        lo = hi = p[0];
        for (const V3 &q : p)
            for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], q[k]); hi[k] = std::max(hi[k], q[k]); }
        h = r;
        for (;;) {
            double cx = std::floor((hi[0] - lo[0]) / h) + 1, cy = std::floor((hi[1] - lo[1]) / h) + 1,
                   cz = std::floor((hi[2] - lo[2]) / h) + 1;
            if (cx * cy * cz <= 4.0e6) { nx = (int)cx; ny = (int)cy; nz = (int)cz; break; }
            h *= 1.5f;
        }
        cell_start.assign((size_t)nx * ny * nz + 1, 0);
        std::vector<int> cid(p.size());
        for (size_t i = 0; i < p.size(); i++) {
            int cx = cell(p[i][0], 0), cy = cell(p[i][1], 1), cz = cell(p[i][2], 2);
            cid[i] = (cz * ny + cy) * nx + cx;
            cell_start[cid[i] + 1]++;
        }
        for (size_t c = 0; c + 1 < cell_start.size(); c++) cell_start[c + 1] += cell_start[c];
        order.resize(p.size());
        std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
        for (size_t i = 0; i < p.size(); i++) order[fill[cid[i]]++] = (int)i;
    }
    inline int cell(float v, int k) const { return (int)std::floor((v - lo[k]) / h); }

    // appends (idx, d2) with d2 < radius2, ascending idx
    void radius(const V3 &q, float radius2, std::vector<std::pair<int, float>> &out) const {
        out.clear();
        const std::vector<V3> &p = *pts;
        if (mode == 2) {
#ifdef ORACLE_NANOFLANN
            std::vector<std::pair<size_t, float>> ret;
            nanoflann::SearchParams params;
            kd->index->radiusSearch(q.data(), radius2, ret, params);
            for (auto &m : ret) out.emplace_back((int)m.first, m.second);
            std::sort(out.begin(), out.end());
#endif
            return;
        }
        if (mode == 1) {
            for (size_t j = 0; j < p.size(); j++) {
                float d2 = dist2(q, p[j]);
                if (d2 < radius2) out.emplace_back((int)j, d2);
            }
            return;
        }
        if (p.empty()) return;
        int cx = cell(q[0], 0), cy = cell(q[1], 1), cz = cell(q[2], 2);
        for (int z = std::max(cz - 1, 0); z <= std::min(cz + 1, nz - 1); z++)
            for (int y = std::max(cy - 1, 0); y <= std::min(cy + 1, ny - 1); y++)
                for (int x = std::max(cx - 1, 0); x <= std::min(cx + 1, nx - 1); x++) {
                    int c = (z * ny + y) * nx + x;
                    for (int s = cell_start[c]; s < cell_start[c + 1]; s++) {
                        int j = order[s];
                        float d2 = dist2(q, p[j]);
                        if (d2 < radius2) out.emplace_back(j, d2);
                    }
                }
        std::sort(out.begin(), out.end());
    }
};

struct Triplet { int i, j; float a; };

// ---------------------------------------------------------------------------------------------
// class cvo::cvo (thirdparty/cvo/include/cvo.hpp:82-282, src/cvo.cpp)
// ---------------------------------------------------------------------------------------------
struct OracleCvo {
    cvo_calib cal;
    cvo_params prm;
    int search_mode = 0;
    Cloud slot[3];  // fixed / moving / previous
    float ell;
    M3 R;
    V3 T;
    M3 tf_lin;  // transform.linear()
    V3 tf_tr;   // transform.translation()
    int iter = -1;
    int A_nonzero = 0;
    int64_t n_evals = 0, n_iters = 0;

    // per-iteration working state
    std::vector<V3> cloud_y;
    std::vector<Triplet> trips;      // A in CSR order (row asc, col asc)
    std::vector<int> row_ptr;
    V3 omega{}, v{};
    double cB = 0, cC = 0, cD = 0, cE = 0;
    float step = 0;

    OracleCvo(const cvo_calib &c, const cvo_params &p) : cal(c), prm(p) {
        ell = p.ell_init;           // cvo.cpp:35
        R = m3_identity();          // :66
        T = V3{0, 0, 0};            // :67
        tf_lin = m3_identity();     // :68
        tf_tr = V3{0, 0, 0};
    }

    // cvo.cpp:106-110
    void update_tf() {
        tf_lin = m3_T(R);
        M3 neg = m3_scale(tf_lin, -1.f);
        tf_tr = m3_vec(neg, T);
    }

    // cvo.cpp:336-341
    void transform_pcd() {
        const Cloud &mv = slot[CVO_SLOT_MOVING];
        cloud_y.resize(mv.n);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < mv.n; j++) {
            V3 r = m3_vec(tf_lin, mv.pos[j]);
            cloud_y[j] = V3{r[0] + tf_tr[0], r[1] + tf_tr[1], r[2] + tf_tr[2]};
        }
    }

    float d2_threshold(float l) const {  // cvo.cpp:125 (double arithmetic, rounded to float)
        float s2 = prm.sigma * prm.sigma;
        return (float)(-2.0 * l * l * std::log(prm.sp_thres / s2));
    }
    float d2c_threshold() const {  // cvo.cpp:126
        return (float)(-2.0 * prm.c_ell * prm.c_ell * std::log(prm.sp_thres / prm.c_sigma / prm.c_sigma));
    }

    // cvo.cpp:122-184
    void se_kernel(float l, float s2) {
        const Cloud &fx = slot[CVO_SLOT_FIXED];
        const Cloud &mv = slot[CVO_SLOT_MOVING];
        float d2_thres = d2_threshold(l);
        float d2_c_thres = d2c_threshold();
        Searcher idx;
        idx.build(cloud_y, d2_thres, search_mode);
        std::vector<std::vector<Triplet>> rows(fx.n);
        int64_t evals = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : evals)
        for (int i = 0; i < fx.n; i++) {
            std::vector<std::pair<int, float>> ret;
            idx.radius(fx.pos[i], d2_thres, ret);
            const F5 &feature_x = fx.feat[i];
            for (auto &m : ret) {
                int id = m.first;
                float d2 = m.second;
                if (d2 < d2_thres) {
                    const F5 &feature_y = mv.feat[id];
                    float d2_color = 0;
                    for (int k = 0; k < 5; k++) {
                        float df = feature_x[k] - feature_y[k];
                        d2_color = d2_color + df * df;
                    }
                    evals++;
                    if (d2_color < d2_c_thres) {
                        float k = s2 * std::exp(-d2 / (2.0 * l * l));                                  // :172
                        float ck = prm.c_sigma * prm.c_sigma * std::exp(-d2_color / (2.0 * prm.c_ell * prm.c_ell));  // :173
                        float a = ck * k;
                        if (a > prm.sp_thres) rows[i].push_back(Triplet{i, id, a});
                    }
                }
            }
        }
        n_evals += evals;
        // A.setFromTriplets + makeCompressed (:182-183): CSR, columns ascending
        trips.clear();
        row_ptr.assign(fx.n + 1, 0);
        for (int i = 0; i < fx.n; i++) {
            row_ptr[i] = (int)trips.size();
            trips.insert(trips.end(), rows[i].begin(), rows[i].end());
        }
        row_ptr[fx.n] = (int)trips.size();
    }

    // cvo.cpp:187-236
    void compute_flow() {
        se_kernel(ell, prm.sigma * prm.sigma);
        const Cloud &fx = slot[CVO_SLOT_FIXED];
        float inv_c = 1 / prm.c, inv_d = 1 / prm.d;
        ExactAcc tw[3], tv[3];
#pragma omp parallel
        {
            ExactAcc lw[3], lv[3];
#pragma omp for schedule(static) nowait
            for (int i = 0; i < fx.n; i++) {
                for (int t = row_ptr[i]; t < row_ptr[i + 1]; t++) {
                    const V3 &x = fx.pos[i];
                    const V3 &y = cloud_y[trips[t].j];
                    V3 cr = cross(x, y);
                    float wa = inv_c * trips[t].a;  // (1/c*Ai)
                    float va = inv_d * trips[t].a;
                    for (int k = 0; k < 3; k++) {
                        float df = y[k] - x[k];
                        lw[k].add((double)wa * (double)cr[k]);   // products of two floats: exact in double
                        lv[k].add((double)va * (double)df);
                    }
                }
            }
#pragma omp critical
            for (int k = 0; k < 3; k++) { tw[k].add(lw[k]); tv[k].add(lv[k]); }
        }
        A_nonzero = (int)trips.size();
        for (int k = 0; k < 3; k++) { omega[k] = (float)tw[k].value(); v[k] = (float)tv[k].value(); }
    }

    // cvo.cpp:239-334
    void compute_step_size() {
        const Cloud &fx = slot[CVO_SLOT_FIXED];
        int num_moving = (int)cloud_y.size();
        M3 oh = skew(omega);
        M3 oh2 = m3_mul(oh, oh), oh3 = m3_mul(oh2, oh), oh4 = m3_mul(oh3, oh);
        V3 ohv = m3_vec(oh, v), oh2v = m3_vec(oh2, v), oh3v = m3_vec(oh3, v);
        std::vector<V3> xiz(num_moving), xi2z(num_moving), xi3z(num_moving), xi4z(num_moving);
        std::vector<float> normxiz2(num_moving), xiz_dot_xi2z(num_moving), epsil_const(num_moving);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < num_moving; j++) {  // :252-264
            const V3 &y = cloud_y[j];
            V3 c1 = cross(omega, y);
            xiz[j] = V3{c1[0] + v[0], c1[1] + v[1], c1[2] + v[2]};
            V3 a2 = m3_vec(oh2, y), a3 = m3_vec(oh3, y), a4 = m3_vec(oh4, y);
            for (int k = 0; k < 3; k++) {
                xi2z[j][k] = a2[k] + ohv[k];
                xi3z[j][k] = a3[k] + oh2v[k];
                xi4z[j][k] = a4[k] + oh3v[k];
            }
            normxiz2[j] = dot3(xiz[j], xiz[j]);
            xiz_dot_xi2z[j] = -dot3(xiz[j], xi2z[j]);
            epsil_const[j] = dot3(xi2z[j], xi2z[j]) + 2 * dot3(xiz[j], xi3z[j]);
        }
        float temp_coef = 1 / (2.0 * ell * ell);  // :267
        float m2tc = (float)(-2.0 * temp_coef);
        float p2tc = (float)(2.0 * temp_coef);
        float mtc = -temp_coef;
        // terms of cvo.cpp:301-305, summed in double-double (order-free, see DDAcc)
        std::vector<DDAcc> rB(fx.n), rC(fx.n), rD(fx.n), rE(fx.n);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < fx.n; i++) {  // :275-315
            DDAcc Bi, Ci, Di, Ei;
            for (int t = row_ptr[i]; t < row_ptr[i + 1]; t++) {
                int idx = trips[t].j;
                const V3 &x = fx.pos[i];
                const V3 &y = cloud_y[idx];
                V3 diff{x[0] - y[0], x[1] - y[1], x[2] - y[2]};
                // (-2.0*temp_coef * xiz.row(idx)) * diff_xy: Eigen scales the row, then dots
                V3 sx{m2tc * xiz[idx][0], m2tc * xiz[idx][1], m2tc * xiz[idx][2]};
                float beta_ij = dot3(sx, diff);
                float gamma_ij = mtc * (normxiz2[idx] + 2.0f * dot3(xi2z[idx], diff));
                float delta_ij = p2tc * (xiz_dot_xi2z[idx] + (-dot3(xi3z[idx], diff)));
                float epsil_ij = mtc * (epsil_const[idx] + 2.0f * dot3(xi4z[idx], diff));
                float A_ij = trips[t].a;
                // the RHS below is evaluated in double where the reference's literals are double
                Bi.add(double(A_ij * beta_ij));
                Ci.add(double(A_ij * (gamma_ij + beta_ij * beta_ij / 2.0)));
                Di.add(double(A_ij * (delta_ij + beta_ij * gamma_ij + beta_ij * beta_ij * beta_ij / 6.0)));
                Ei.add(double(A_ij * (epsil_ij + beta_ij * delta_ij + 1 / 2.0 * beta_ij * beta_ij * gamma_ij +
                                      1 / 2.0 * gamma_ij * gamma_ij + 1 / 24.0 * beta_ij * beta_ij * beta_ij * beta_ij)));
            }
            rB[i] = Bi; rC[i] = Ci; rD[i] = Di; rE[i] = Ei;
        }
        DDAcc sB, sC, sD, sE;
        for (int i = 0; i < fx.n; i++) { sB.add(rB[i]); sC.add(rC[i]); sD.add(rD[i]); sE.add(rE[i]); }
        double B = sB.value(), C = sC.value(), D = sD.value(), E = sE.value();
        cB = B; cC = C; cD = D; cE = E;
        // :317-333
        float p0 = 4.0 * float(E), p1 = 3.0 * float(D), p2 = 2.0 * float(C), p3 = float(B);
        double re[3];
        // the companion matrix holds the FLOAT quotients -(coef/coef(0)) (cvo.cpp:86)
        int nr = cubic_real_roots(1.0, p1 / p0, p2 / p0, p3 / p0, re);
        float temp_step = std::numeric_limits<float>::max();
        for (int i = 0; i < nr; i++) {
            float r = (float)re[i];
            if (r > 0 && r < temp_step) temp_step = r;
        }
        step = temp_step == std::numeric_limits<float>::max() ? prm.min_step : temp_step;
        step = step > prm.max_step ? prm.max_step : step;
    }

    void record(cvo_iter_record *r) const {
        r->ell = ell;
        for (int k = 0; k < 3; k++) { r->omega[k] = omega[k]; r->v[k] = v[k]; }
        r->B = cB; r->C = cC; r->D = cD; r->E = cE;
        r->step = step;
        r->nnz = A_nonzero;
    }

    // cvo.cpp:763-821
    int align(cvo_align_result *out, cvo_iter_record *trace, int trace_cap) {
        if (!slot[CVO_SLOT_FIXED].valid || !slot[CVO_SLOT_MOVING].valid) return CVO_ERR_NOT_INIT;
        int iterations = prm.max_iter;
        iter = -1;
        for (int k = 0; k < prm.max_iter; k++) {
            update_tf();
            transform_pcd();
            compute_flow();
            compute_step_size();
            n_iters++;
            if (trace && k < trace_cap) record(&trace[k]);
            if (norm3(omega) < prm.eps && norm3(v) < prm.eps) { iter = k; iterations = k + 1; break; }
            M3 dR; V3 dT;
            Exp_SEK3(omega, v, step, dR, dT);
            V3 RdT = m3_vec(R, dT);
            T = V3{RdT[0] + T[0], RdT[1] + T[1], RdT[2] + T[2]};
            R = m3_mul(R, dR);
            if (dist_se3(dR, dT) < prm.eps_2) { iter = k; iterations = k + 1; break; }
            ell = (k > 2) ? prm.ell_after_k2 : ell;
            ell = (k > 9) ? prm.ell_after_k9 : ell;
            ell = (k > 19) ? prm.ell_after_k19 : ell;
        }
        // cvo.cpp:815-816: prev_transform / accum_transform take `transform` as the last executed
        // iteration's update_tf() left it, BEFORE the final update_tf()
        float last_tf[16] = {0};
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) last_tf[i * 4 + j] = tf_lin.m[i][j];
            last_tf[i * 4 + 3] = tf_tr[i];
        }
        last_tf[15] = 1.f;
        update_tf();
        if (out) {
            memset(out, 0, sizeof(*out));
            memcpy(out->last_iter_transform, last_tf, sizeof(last_tf));
            out->num_fixed = slot[CVO_SLOT_FIXED].n;
            out->num_moving = slot[CVO_SLOT_MOVING].n;
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) { out->transform[i * 4 + j] = tf_lin.m[i][j]; out->R[i * 3 + j] = R.m[i][j]; }
                out->transform[i * 4 + 3] = tf_tr[i];
                out->T[i] = T[i];
            }
            out->transform[15] = 1.f;
            out->ell = ell;
            out->iterations = iterations;
            out->iter = iter;
            out->A_nonzero = A_nonzero;
            out->status = CVO_OK;
        }
        return CVO_OK;
    }

    int iteration_at(const float Rin[9], const float Tin[3], float l, cvo_iter_record *rec) {
        if (!slot[CVO_SLOT_FIXED].valid || !slot[CVO_SLOT_MOVING].valid) return CVO_ERR_NOT_INIT;
        M3 Rs = R; V3 Ts = T; float ls = ell; M3 tl = tf_lin; V3 tt = tf_tr;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R.m[i][j] = Rin[i * 3 + j]; T[i] = Tin[i]; }
        ell = l;
        update_tf();
        transform_pcd();
        compute_flow();
        compute_step_size();
        if (rec) record(rec);
        R = Rs; T = Ts; ell = ls; tf_lin = tl; tf_tr = tt;
        return CVO_OK;
    }

    static void apply(const float *Ta, const Cloud &c, std::vector<V3> &out) {
        out.resize(c.n);
        for (int i = 0; i < c.n; i++) {
            if (!Ta) { out[i] = c.pos[i]; continue; }
            const V3 &p = c.pos[i];
            for (int r = 0; r < 3; r++) {
                float s = Ta[r * 4 + 0] * p[0];
                s = s + Ta[r * 4 + 1] * p[1];
                s = s + Ta[r * 4 + 2] * p[2];
                out[i][r] = s + Ta[r * 4 + 3];
            }
        }
    }

    // cvo.cpp:388-459
    int inner_product(int sa, const float *Ta, int sb, float *value, int *num) {
        const Cloud &a = slot[sa];
        const Cloud &b = slot[sb];
        if (!a.valid || !b.valid) return CVO_ERR_NOT_INIT;
        std::vector<V3> pa;
        apply(Ta, a, pa);
        double sum_A = 0, sum = 0;
        float d2_thres = (float)(-2.0 * ell * ell * std::log(prm.sp_thres / prm.sigma / prm.sigma));
        float d2_c_thres = d2c_threshold();
        Searcher idx;
        idx.build(b.pos, d2_thres, search_mode);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : sum_A, sum)
        for (int i = 0; i < a.n; i++) {
            std::vector<std::pair<int, float>> ret;
            idx.radius(pa[i], d2_thres, ret);
            const F5 &fa = a.feat[i];
            for (auto &m : ret) {
                float d2 = m.second;
                if (d2 < d2_thres) {
                    const F5 &fb = b.feat[m.first];
                    float d2_color = 0;
                    for (int k = 0; k < 5; k++) { float df = fa[k] - fb[k]; d2_color = d2_color + df * df; }
                    if (d2_color < d2_c_thres) {
                        float k = prm.sigma * prm.sigma * std::exp(-d2 / (2.0 * ell * ell));
                        float ck = prm.c_sigma * prm.c_sigma * std::exp(-d2_color / (2.0 * prm.c_ell * prm.c_ell));
                        float av = ck * k;
                        sum_A += av;
                        sum += 1;
                    }
                }
            }
        }
        if (sum == 0) sum = 1;
        *value = (float)sum_A;  // inn_p(float v, int n, int n_e), cvo.hpp:71
        *num = (int)sum;
        return CVO_OK;
    }

    // cvo.cpp:620-759
    int hessian(int sa, const float *Ta, int sb, double Hout[36], int *inliers_out) {
        const Cloud &a = slot[sa];
        const Cloud &b = slot[sb];
        if (!a.valid || !b.valid) return CVO_ERR_NOT_INIT;
        std::vector<V3> pa;
        apply(Ta, a, pa);
        float H[6][6];
        memset(H, 0, sizeof(H));
        int inliers = 0;
        float d2_thres = (float)(-2.0 * ell * ell * std::log(prm.sp_thres / prm.sigma / prm.sigma));
        float d2_c_thres = d2c_threshold();
        Searcher idx;
        idx.build(b.pos, d2_thres, search_mode);
        float iell2 = 1 / (ell * ell);
        std::vector<std::pair<int, float>> ret;
        for (int i = 0; i < a.n; i++) {  // sequential: float accumulation order = (i asc, j asc)
            idx.radius(pa[i], d2_thres, ret);
            const V3 &A_ = pa[i];
            const F5 &fa = a.feat[i];
            for (auto &m : ret) {
                float d2 = m.second;
                if (!(d2 < d2_thres)) continue;
                const F5 &fb = b.feat[m.first];
                float d2_color = 0;
                for (int k = 0; k < 5; k++) { float df = fa[k] - fb[k]; d2_color = d2_color + df * df; }
                if (!(d2_color < d2_c_thres)) continue;
                const V3 &B_ = b.pos[m.first];
                float k = prm.sigma * prm.sigma * std::exp(-d2 / (2.0 * ell * ell));
                float cdot = 0;
                for (int q = 0; q < 5; q++) cdot = cdot + fa[q] * fb[q];
                V3 cr = cross(A_, B_);
                float Bl[6][6];
                float dot1 = A_[1] * B_[1] + A_[2] * B_[2];
                float dot2 = A_[0] * B_[0] + A_[2] * B_[2];
                float dot3_ = A_[0] * B_[0] + A_[1] * B_[1];
                Bl[0][0] = iell2 * cr[0] * cr[0] - dot1;
                Bl[1][1] = iell2 * cr[1] * cr[1] - dot2;
                Bl[2][2] = iell2 * cr[2] * cr[2] - dot3_;
                Bl[0][1] = Bl[1][0] = iell2 * cr[0] * cr[1] + 0.5 * (A_[0] * B_[1] + A_[1] * B_[0]);
                Bl[0][2] = Bl[2][0] = iell2 * cr[0] * cr[2] + 0.5 * (A_[0] * B_[2] + A_[2] * B_[0]);
                Bl[1][2] = Bl[2][1] = iell2 * cr[1] * cr[2] + 0.5 * (A_[1] * B_[2] + A_[2] * B_[1]);
                V3 df{B_[0] - A_[0], B_[1] - A_[1], B_[2] - A_[2]};
                float Cm[3][3];
                Cm[0][0] = iell2 * cr[0] * df[0];
                Cm[1][1] = iell2 * cr[1] * df[1];
                Cm[2][2] = iell2 * cr[2] * df[2];
                Cm[1][0] = A_[2] + iell2 * df[1] * cr[0];
                Cm[2][0] = -A_[1] + iell2 * df[2] * cr[0];
                Cm[0][1] = -A_[2] + iell2 * df[0] * cr[1];
                Cm[2][1] = A_[0] + iell2 * df[2] * cr[1];
                Cm[0][2] = A_[1] + iell2 * df[0] * cr[2];
                Cm[1][2] = -A_[0] + iell2 * df[1] * cr[2];
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) {
                        Bl[3 + r][c] = Cm[r][c];        // Blocks(3,0) = C
                        Bl[c][3 + r] = Cm[r][c];        // Blocks(0,3) = C^T
                    }
                Bl[3][3] = iell2 * df[0] * df[0] - 1;
                Bl[4][4] = iell2 * df[1] * df[1] - 1;
                Bl[5][5] = iell2 * df[2] * df[2] - 1;
                Bl[3][4] = Bl[4][3] = iell2 * df[0] * df[1];
                Bl[3][5] = Bl[5][3] = iell2 * df[0] * df[2];
                Bl[4][5] = Bl[5][4] = iell2 * df[1] * df[2];
                float wgt = iell2 * cdot * k;
                for (int r = 0; r < 6; r++)
                    for (int c = 0; c < 6; c++) H[r][c] = H[r][c] + wgt * Bl[r][c];
                inliers++;
            }
        }
        *inliers_out = inliers;
        finish_hessian(&H[0][0], inliers, Hout);
        return CVO_OK;
    }

    // cvo.cpp:726-758: scale, eigenvalue shift until min |lambda| >= 1
    static void finish_hessian(const float *Hf, int inliers, double Hout[36]) {
        float H[36];
        if (inliers) {
            for (int i = 0; i < 36; i++) H[i] = Hf[i] * (float)(-1.0 / 100000);
            double Hd[36], evd[6];
            for (int i = 0; i < 36; i++) Hd[i] = H[i];
            sym_eigenvalues6(Hd, evd);
            float ev[6];
            for (int i = 0; i < 6; i++) ev[i] = (float)evd[i];
            float sufficient_scale = 0.0;
            auto argmin_abs = [&]() { int m = 0; for (int i = 1; i < 6; i++) if (std::fabs(ev[i]) < std::fabs(ev[m])) m = i; return m; };
            float min_eigen = ev[argmin_abs()];
            int guard = 0;
            while (std::fabs(min_eigen) < 1.0 && guard++ < 64) {
                sufficient_scale += (1.0 - min_eigen);
                float add = (1.0 - min_eigen);
                for (int i = 0; i < 6; i++) ev[i] += add * 1.0f;
                min_eigen = ev[argmin_abs()];
            }
            for (int i = 0; i < 6; i++) H[i * 6 + i] += sufficient_scale;
        } else {
            for (int i = 0; i < 36; i++) H[i] = 0;
            for (int i = 0; i < 6; i++) H[i * 6 + i] = 1;
        }
        for (int i = 0; i < 36; i++) Hout[i] = H[i];
    }
};

}  // namespace

// =============================================================================================
// C interface for ctypes (mirrors include/cvo_b200.h with the prefix oracle_)
// =============================================================================================
extern "C" {

struct oracle_handle { OracleCvo *o; };

int oracle_has_nanoflann(void) {
#ifdef ORACLE_NANOFLANN
    return 1;
#else
    return 0;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void oracle_default_params(cvo_params *p) {
    p->ell_init = 0.15f; p->sigma = 0.1f; p->sp_thres = 8e-3f; p->c = 7.0f; p->d = 7.0f;
    p->c_ell = 200.f; p->c_sigma = 1.f; p->max_iter = 2000; p->min_step = 2 * 1.0e-1f;
    p->max_step = 0.8f; p->eps = 5 * 1.0e-5f; p->eps_2 = 1.0e-5f;
    p->ell_after_k2 = 0.10f; p->ell_after_k9 = 0.06f; p->ell_after_k19 = 0.03f;
    p->num_want = 3000; p->feature_type = 1; p->gray_mode = 0; p->exp_mode = 0;
}

int oracle_create(const cvo_calib *c, const cvo_params *p, oracle_handle **out) {
    if (!c || !p || !out) return CVO_ERR_INVALID;
    *out = new oracle_handle{new OracleCvo(*c, *p)};
    return CVO_OK;
}
int oracle_destroy(oracle_handle *h) { if (h) { delete h->o; delete h; } return CVO_OK; }
int oracle_set_search(oracle_handle *h, int mode) { h->o->search_mode = mode; return CVO_OK; }

int oracle_set_frame(oracle_handle *h, int slot, const uint8_t *bgr, size_t bgr_stride,
                     const uint16_t *depth, size_t depth_stride, int w, int hgt) {
    if (!h || slot < 0 || slot > 2 || !bgr || !depth) return CVO_ERR_INVALID;
    create_pointcloud(bgr, bgr_stride, depth, depth_stride, w, hgt, h->o->cal, h->o->prm, h->o->slot[slot], nullptr);
    return CVO_OK;
}

int oracle_set_cloud(oracle_handle *h, int slot, int n, const float *pos, const float *feat) {
    if (!h || slot < 0 || slot > 2 || n < 0) return CVO_ERR_INVALID;
    Cloud c;
    c.n = n;
    c.pos.resize(n);
    c.feat.resize(n);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) c.pos[i][k] = pos[i * 3 + k];
        for (int k = 0; k < 5; k++) c.feat[i][k] = feat[i * 5 + k];
    }
    c.pix.assign((size_t)2 * n, 0.f);
    c.valid = true;
    h->o->slot[slot] = std::move(c);
    return CVO_OK;
}

int oracle_slot_move(oracle_handle *h, int dst, int src) {
    if (!h || dst < 0 || dst > 2 || src < 0 || src > 2) return CVO_ERR_INVALID;
    if (dst == src) return CVO_OK;
    h->o->slot[dst] = std::move(h->o->slot[src]);
    h->o->slot[src] = Cloud();
    return CVO_OK;
}
int oracle_slot_size(oracle_handle *h, int slot, int *n) {
    if (!h->o->slot[slot].valid) { *n = 0; return CVO_ERR_NOT_INIT; }
    *n = h->o->slot[slot].n;
    return CVO_OK;
}
int oracle_set_RT(oracle_handle *h, const float R[9], const float T[3]) {
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) h->o->R.m[i][j] = R[i * 3 + j]; h->o->T[i] = T[i]; }
    return CVO_OK;
}
int oracle_get_RT(oracle_handle *h, float R[9], float T[3]) {
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i * 3 + j] = h->o->R.m[i][j]; T[i] = h->o->T[i]; }
    return CVO_OK;
}
int oracle_set_ell(oracle_handle *h, float ell) { h->o->ell = ell; return CVO_OK; }
int oracle_get_ell(oracle_handle *h, float *ell) { *ell = h->o->ell; return CVO_OK; }

int oracle_align(oracle_handle *h, cvo_align_result *out, cvo_iter_record *trace, int trace_cap) {
    return h->o->align(out, trace, trace_cap);
}
int oracle_iteration_at(oracle_handle *h, const float R[9], const float T[3], float ell, cvo_iter_record *rec) {
    return h->o->iteration_at(R, T, ell, rec);
}
// in-cutoff pattern of the last iteration (after oracle_iteration_at / oracle_align): (i, j, a)
int oracle_last_pattern(oracle_handle *h, int32_t *ij, float *a, int cap, int *n) {
    int m = (int)h->o->trips.size();
    *n = m;
    for (int t = 0; t < std::min(m, cap); t++) {
        ij[2 * t] = h->o->trips[t].i;
        ij[2 * t + 1] = h->o->trips[t].j;
        a[t] = h->o->trips[t].a;
    }
    return CVO_OK;
}
int oracle_inner_product(oracle_handle *h, int sa, const float *Ta, int sb, float *value, int *num) {
    return h->o->inner_product(sa, Ta, sb, value, num);
}
int oracle_hessian(oracle_handle *h, int sa, const float *Ta, int sb, double H[36], int *inl) {
    return h->o->hessian(sa, Ta, sb, H, inl);
}
int oracle_get_selected_points(oracle_handle *h, int slot, float *xy, int cap, int *n) {
    const Cloud &c = h->o->slot[slot];
    if (!c.valid) return CVO_ERR_NOT_INIT;
    *n = c.n;
    memcpy(xy, c.pix.data(), sizeof(float) * 2 * std::min(cap, c.n));
    return CVO_OK;
}
int oracle_get_cloud(oracle_handle *h, int slot, float *pos, float *feat, int cap, int *n) {
    const Cloud &c = h->o->slot[slot];
    if (!c.valid) return CVO_ERR_NOT_INIT;
    *n = c.n;
    for (int i = 0; i < std::min(cap, c.n); i++) {
        for (int k = 0; k < 3; k++) pos[i * 3 + k] = c.pos[i][k];
        for (int k = 0; k < 5; k++) feat[i * 5 + k] = c.feat[i][k];
    }
    return CVO_OK;
}
int oracle_get_selection_debug(oracle_handle *h, int slot, uint8_t *map, int32_t info[5]) {
    const Cloud &c = h->o->slot[slot];
    if (!c.valid || c.map.empty()) return CVO_ERR_NOT_INIT;
    if (map) memcpy(map, c.map.data(), c.map.size());
    for (int i = 0; i < 5; i++) info[i] = c.info[i];
    return CVO_OK;
}
int oracle_stats(oracle_handle *h, int64_t s[3]) {
    s[0] = 0; s[1] = h->o->n_evals; s[2] = h->o->n_iters;
    return CVO_OK;
}

// ---- stage-level entry points for the unit tests ------------------------------------------------
int oracle_gray(const uint8_t *bgr, int n_px, int mode, uint8_t *out) {
    for (int i = 0; i < n_px; i++) out[i] = gray_px(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2], mode);
    return 0;
}
int oracle_hsv(const uint8_t *bgr, int n_px, uint8_t *out) {
    for (int i = 0; i < n_px; i++) hsv_px(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2], out + 3 * i);
    return 0;
}
// pyramid + thresholds: g2 levels concatenated (w*h + (w/2)*(h/2) + (w/4)*(h/4)), ths and
// thsSmoothed ((w/32)*(h/32) each), level-0 dx/dy (w*h each)
int oracle_stages(const uint8_t *bgr, int w, int h, int gray_mode, uint8_t *gray, float *g2,
                  float *ths, float *ths_smoothed, float *dx0, float *dy0) {
    cvo_calib cal{5000.f, 500.f, 500.f, w * 0.5f, h * 0.5f};
    cvo_params prm;
    oracle_default_params(&prm);
    prm.gray_mode = gray_mode;
    std::vector<uint16_t> depth((size_t)w * h, 1);
    Cloud c;
    StageDump d;
    create_pointcloud(bgr, (size_t)3 * w, depth.data(), (size_t)2 * w, w, h, cal, prm, c, &d);
    if (gray) memcpy(gray, d.gray.data(), d.gray.size());
    size_t off = 0;
    for (int l = 0; l < 3; l++) {
        if (g2) memcpy(g2 + off, d.pyr.g2[l].data(), sizeof(float) * d.pyr.g2[l].size());
        off += d.pyr.g2[l].size();
    }
    size_t nb = (size_t)(w / 32) * (h / 32);
    if (ths) memcpy(ths, d.ths.data(), sizeof(float) * nb);
    if (ths_smoothed) memcpy(ths_smoothed, d.thsSmoothed.data(), sizeof(float) * nb);
    if (dx0) memcpy(dx0, d.pyr.dx[0].data(), sizeof(float) * (size_t)w * h);
    if (dy0) memcpy(dy0, d.pyr.dy[0].data(), sizeof(float) * (size_t)w * h);
    return 0;
}
int oracle_random_pattern(uint8_t *out, int n) {
    std::srand(3141592);
    for (int i = 0; i < n; i++) out[i] = rand() & 0xFF;
    return 0;
}
int oracle_cubic_real_roots(double a, double b, double c, double d, double re[3]) {
    return cubic_real_roots(a, b, c, d, re);
}
// step selection exactly as compute_step_size does from (B, C, D, E)
float oracle_step_from_coeffs(double B, double C, double D, double E, float min_step, float max_step) {
    float p0 = 4.0 * float(E), p1 = 3.0 * float(D), p2 = 2.0 * float(C), p3 = float(B);
    double re[3];
    int nr = cubic_real_roots(1.0, p1 / p0, p2 / p0, p3 / p0, re);
    float temp_step = std::numeric_limits<float>::max();
    for (int i = 0; i < nr; i++) { float r = (float)re[i]; if (r > 0 && r < temp_step) temp_step = r; }
    float step = temp_step == std::numeric_limits<float>::max() ? min_step : temp_step;
    return step > max_step ? max_step : step;
}
int oracle_exp_sek3(const float w[3], const float v[3], float dt, float R[9], float dT[3]) {
    M3 Rm; V3 t;
    Exp_SEK3(V3{w[0], w[1], w[2]}, V3{v[0], v[1], v[2]}, dt, Rm, t);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i * 3 + j] = Rm.m[i][j]; dT[i] = t[i]; }
    return 0;
}
float oracle_dist_se3(const float R[9], const float T[3]) {
    M3 Rm;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rm.m[i][j] = R[i * 3 + j];
    return dist_se3(Rm, V3{T[0], T[1], T[2]});
}
int oracle_finish_hessian(const float Hf[36], int inliers, double Hout[36]) {
    OracleCvo::finish_hessian(Hf, inliers, Hout);
    return 0;
}
// radius search with a chosen back end: returns count, writes ascending (j, d2)
int oracle_radius_search(const float *pts, int n, const float *q, float radius2, int mode, int32_t *idx,
                         float *d2, int cap) {
    std::vector<V3> p(n);
    for (int i = 0; i < n; i++) p[i] = V3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    Searcher s;
    s.build(p, radius2, mode);
    std::vector<std::pair<int, float>> out;
    s.radius(V3{q[0], q[1], q[2]}, radius2, out);
    for (int i = 0; i < (int)out.size() && i < cap; i++) { idx[i] = out[i].first; d2[i] = out[i].second; }
    return (int)out.size();
}

}  // extern "C"
