// select.cu — point selection + feature construction on the device.
//
// Replaces, per frame: cv::cvtColor RGB2GRAY (pcd_generator.cpp:624), make_pyramid (:50-143),
// dso::PixelSelector::makeHists / select / makeMaps (PixelSelector2.cpp:71-433),
// get_points_from_pixels (:456-499) and get_features (:563-616).  Bit-exact by construction:
// every float value on this path is either an exactly representable small rational or is
// produced by one IEEE-rounded operation written with explicit _rn intrinsics (no FMA
// contraction, IEEE sqrt and division).
//
// The sequential tile walk of PixelSelector::select reduces to three order-independent rules
// (derived in DESIGN.md §selection): per pot-block the first arg-max of g2_0 above threshold;
// per 2pot-tile without any level-0 hit the first arg-max of g2_1; per 4pot-tile without any
// level-0 or level-1 hit the first arg-max of g2_2 — "first" in the reference's traversal
// order.  One lane handles one pot-block, 16 lanes one 4pot-tile, combined with shuffles.
//
// All kernels take a frame index in blockIdx.y so that a chunk of frames is one launch.

#include "common.cuh"

#include <cooperative_groups.h>

#include <math.h>
#include <stdarg.h>
#include <string.h>

namespace cvo_b200 {

// ------------------------------------------------------------------------------------------------
struct SelState {       // per chunk-local frame, device
    int n_a[3];         // n2, n3, n4 of select pass 1
    int n_b[3];         // ... of pass 2 (if any)
    int need2, pot2;    // recursion decision (PixelSelector2.cpp:193-223)
    int use_b;          // final map is map_b
    int subsample, charTH;
    int n_selected;     // after sub-sampling (numHaveSub)
    int n_out;          // after the depth != 0 filter
    int overflow;       // n_out exceeded the cloud capacity
    int pad[2];
};

struct SelDev {         // kernel parameter block
    int w, h, w1, h1, w2, h2, w32, h32, nb;
    int npx, npx_pad;   // pixels per frame, padded to 16
    size_t n1, n2;      // pixels at level 1, 2
    uint8_t *gray;      // [chunk][npx_pad]
    float *I1, *I2;     // [chunk][n1], [chunk][n2]
    float *g0, *g1, *g2;
    float *ths, *thsS;  // [chunk][nb]
    uint8_t *map_a, *map_b;  // [chunk][npx_pad]
    SelState *st;       // [chunk]
    const uint8_t *rnd; // [npx] rand() & 0xFF table
    float num_want;
    int gray_mode, feature_type;
    float scaling_factor, fx, fy, cx, cy;
};

struct SelWorkspace {
    SelDev d;
    int chunk;
    uint8_t *bgr = nullptr;      // staging [chunk][npx*3 padded]
    uint16_t *depth = nullptr;   // staging [chunk][npx padded]
    size_t bgr_stride, depth_stride;
    uint8_t *rnd = nullptr;
    void *blob = nullptr;
};

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

// ------------------------------------------------------------------------------------------------
// glibc random_r TYPE_3 (r[i] = r[i-3] + r[i-31]) seeded with srand(3141592), as used at
// PixelSelector2.cpp:37-38.  Restated here so that the library neither touches nor depends on
// the process-global libc RNG state; pinned against libc's rand() by tests/test_host_logic.py.
void host_random_pattern(uint8_t *out, int n) {
    const uint32_t seed = 3141592u;
    int32_t r[34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    // state as a ring of 31 words, front at 3, rear at 0; discard 310 outputs
    uint32_t st[31];
    for (int i = 0; i < 31; i++) st[i] = (uint32_t)r[i];
    int f = 3, b = 0;
    auto next = [&]() -> uint32_t {
        st[f] += st[b];
        uint32_t res = st[f] >> 1;
        f = (f + 1) % 31;
        b = (b + 1) % 31;
        return res;
    };
    for (int i = 0; i < 310; i++) next();
    for (int i = 0; i < n; i++) out[i] = (uint8_t)(next() & 0xFF);
}

// ------------------------------------------------------------------------------------------------
// K1a: BGR8 -> gray8 (row A).  The image is BGR but converted with the RGB code, so stored
// channel 0 takes the R weight.  4 pixels / thread, 3 x 32-bit loads, 1 x 32-bit store.
__device__ __forceinline__ uint32_t gray_px(uint32_t c0, uint32_t c1, uint32_t c2, int mode) {
    return mode == 1 ? (c0 * 4899u + c1 * 9617u + c2 * 1868u + 8192u) >> 14
                     : (c0 * 9798u + c1 * 19235u + c2 * 3735u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) k_gray(SelDev d, const uint8_t *__restrict__ bgr, size_t bgr_stride) {
    const int f = blockIdx.y;
    const uint8_t *src = bgr + (size_t)f * bgr_stride;
    uint8_t *dst = d.gray + (size_t)f * d.npx_pad;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = q * 4;
    if (p0 >= d.npx) return;
    if (p0 + 4 <= d.npx && (((uintptr_t)src) & 3) == 0) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src) + 3 * (size_t)q;
        uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
        // a: B0 G0 R0 B1 | b: G1 R1 B2 G2 | c: R2 B3 G3 R3  (little endian)
        uint32_t g0 = gray_px(a & 255, (a >> 8) & 255, (a >> 16) & 255, d.gray_mode);
        uint32_t g1 = gray_px(a >> 24, b & 255, (b >> 8) & 255, d.gray_mode);
        uint32_t g2 = gray_px((b >> 16) & 255, b >> 24, c & 255, d.gray_mode);
        uint32_t g3 = gray_px((c >> 8) & 255, (c >> 16) & 255, c >> 24, d.gray_mode);
        *reinterpret_cast<uint32_t *>(dst + p0) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (int p = p0; p < min(p0 + 4, d.npx); p++)
            dst[p] = (uint8_t)gray_px(src[3 * (size_t)p], src[3 * (size_t)p + 1], src[3 * (size_t)p + 2], d.gray_mode);
    }
}

// K1b/c: 2x2 box mean with the reference's stride prev_wl = 2*wl (pcd_generator.cpp:100-115;
// wrong for odd widths — replicated).  Sums are exact in fp32.
__global__ void __launch_bounds__(256) k_pyr(SelDev d, int level) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (level == 1) {
        if (i >= d.w1 * d.h1) return;
        const uint8_t *P = d.gray + (size_t)f * d.npx_pad;
        int x = i % d.w1, y = i / d.w1, pw = d.w1 * 2;
        size_t b = (size_t)2 * x + (size_t)2 * y * pw;
        float s = __fadd_rn((float)P[b], (float)P[b + 1]);
        s = __fadd_rn(s, (float)P[b + pw]);
        s = __fadd_rn(s, (float)P[b + 1 + pw]);
        d.I1[(size_t)f * d.n1 + i] = __fmul_rn(0.25f, s);
    } else {
        if (i >= d.w2 * d.h2) return;
        const float *P = d.I1 + (size_t)f * d.n1;
        int x = i % d.w2, y = i / d.w2, pw = d.w2 * 2;
        size_t b = (size_t)2 * x + (size_t)2 * y * pw;
        float s = __fadd_rn(P[b], P[b + 1]);
        s = __fadd_rn(s, P[b + pw]);
        s = __fadd_rn(s, P[b + 1 + pw]);
        d.I2[(size_t)f * d.n2 + i] = __fmul_rn(0.25f, s);
    }
}

// K1d: squared gradient magnitude at the three levels (pcd_generator.cpp:119-135); indices
// outside [wl, wl*(hl-1)) stay zero.  One flat launch over n0 + n1 + n2 pixels.
__device__ __forceinline__ float g2_from(float l, float r, float u, float dn) {
    float dx = __fmul_rn(0.5f, __fsub_rn(r, l));
    float dy = __fmul_rn(0.5f, __fsub_rn(dn, u));
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__global__ void __launch_bounds__(256) k_grad(SelDev d) {
    const int f = blockIdx.y;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)d.npx) {
        const uint8_t *I = d.gray + (size_t)f * d.npx_pad;
        int idx = (int)i, wl = d.w, hl = d.h;
        float v = 0.f;
        if (idx >= wl && idx < wl * (hl - 1))
            v = g2_from((float)I[idx - 1], (float)I[idx + 1], (float)I[idx - wl], (float)I[idx + wl]);
        d.g0[(size_t)f * d.npx_pad + idx] = v;
        return;
    }
    i -= d.npx;
    if (i < d.n1) {
        const float *I = d.I1 + (size_t)f * d.n1;
        int idx = (int)i, wl = d.w1, hl = d.h1;
        float v = 0.f;
        if (idx >= wl && idx < wl * (hl - 1)) v = g2_from(I[idx - 1], I[idx + 1], I[idx - wl], I[idx + wl]);
        d.g1[(size_t)f * d.n1 + idx] = v;
        return;
    }
    i -= d.n1;
    if (i < d.n2) {
        const float *I = d.I2 + (size_t)f * d.n2;
        int idx = (int)i, wl = d.w2, hl = d.h2;
        float v = 0.f;
        if (idx >= wl && idx < wl * (hl - 1)) v = g2_from(I[idx - 1], I[idx + 1], I[idx - wl], I[idx + wl]);
        d.g2[(size_t)f * d.n2 + idx] = v;
    }
}

// K2: per 32x32 block histogram of min(int(sqrtf(g2_0)), 48) and its 50 % quantile + 7
// (PixelSelector2.cpp:83-105, :59-68).  One CTA per block.
__global__ void __launch_bounds__(256) k_hist(SelDev d) {
    __shared__ int hist[92];
    const int f = blockIdx.y;
    const int bx = blockIdx.x % d.w32, by = blockIdx.x / d.w32;
    for (int i = threadIdx.x; i < 92; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const float *g0 = d.g0 + (size_t)f * d.npx_pad;
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
        int it = (k & 31) + 32 * bx, jt = (k >> 5) + 32 * by;
        if (it > d.w - 2 || jt > d.h - 2 || it < 1 || jt < 1) continue;
        int g = (int)__fsqrt_rn(g0[(size_t)jt * d.w + it]);
        if (g > 48) g = 48;
        atomicAdd(&hist[g + 1], 1);
        atomicAdd(&hist[0], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int th = (int)__fadd_rn(__fmul_rn((float)hist[0], 0.5f), 0.5f);
        int q = 90;
        for (int i = 0; i < 90; i++) {
            th -= hist[i + 1];
            if (th < 0) { q = i; break; }
        }
        d.ths[(size_t)f * d.nb + bx + by * d.w32] = (float)(q + 7);
    }
}

// K2b: 3x3 neighbourhood mean, squared (PixelSelector2.cpp:107-131).  The sums are small
// integers (exact in any order); division and product are single IEEE operations.
__global__ void __launch_bounds__(128) k_smooth(SelDev d) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.nb) return;
    const float *ths = d.ths + (size_t)f * d.nb;
    const int x = i % d.w32, y = i / d.w32, w32 = d.w32, h32 = d.h32;
    float sum = 0, num = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            int xx = x + dx, yy = y + dy;
            if (xx < 0 || xx >= w32 || yy < 0 || yy >= h32) continue;
            num += 1.f;
            sum += ths[xx + yy * w32];
        }
    float m = __fdiv_rn(sum, num);
    d.thsS[(size_t)f * d.nb + i] = __fmul_rn(m, m);
}

// makeMaps' recursion decision after the first select pass (PixelSelector2.cpp:184-223),
// evaluated redundantly by whoever needs it (cheap) instead of a separate launch.
__device__ __forceinline__ void decide_pass2(const int n_a[3], float num_want, int pot, int &need2,
                                             int &pot2) {
    float numHave = (float)(n_a[0] + n_a[1] + n_a[2]);
    float quotia = __fdiv_rn(num_want, numHave);
    float K = __fmul_rn(__fmul_rn(numHave, (float)(pot + 1)), (float)(pot + 1));
    int ideal = (int)__fsub_rn(__fsqrt_rn(__fdiv_rn(K, num_want)), 1.0f);
    if (ideal < 1) ideal = 1;
    need2 = 0;
    pot2 = pot;
    if ((double)quotia > 1.25 && pot > 1) {
        if (ideal >= pot) ideal = pot - 1;
        need2 = 1;
        pot2 = ideal;
    } else if ((double)quotia < 0.25) {
        if (ideal <= pot) ideal = pot + 1;
        need2 = 1;
        pot2 = ideal;
    }
}

// K3: hierarchical selection (PixelSelector2.cpp:290-433).  lane s of a 16-lane group =
// pot-block (y3idx, x3idx, y2idx, x2idx) of one 4pot-tile, i.e. the reference's visit order.
__device__ __forceinline__ void take_better(float &v, int &idx, unsigned lane, int off) {
    float ov = __shfl_xor_sync(0xffffffffu, v, off);
    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    unsigned ol = lane ^ off;
    // larger value wins; on a tie the earlier pot-block (lower lane) wins ("dirNorm > bestVal")
    if (ov > v || (ov == v && ol < lane)) { v = ov; idx = oi; }
}

__global__ void __launch_bounds__(256) k_select(SelDev d, int pass) {
    const int f = blockIdx.y;
    SelState *st = d.st + f;
    int pot = 3;   // PixelSelector2.cpp:40, a fresh selector per frame (pcd_generator.cpp:154)
    uint8_t *map = d.map_a + (size_t)f * d.npx_pad;
    int *n_out = st->n_a;
    if (pass == 1) {
        int need2, pot2;
        decide_pass2(st->n_a, d.num_want, 3, need2, pot2);
        if (!need2) return;
        pot = pot2;
        map = d.map_b + (size_t)f * d.npx_pad;
        n_out = st->n_b;
    }
    const int w = d.w, h = d.h;
    const int tiles_x = (w + 4 * pot - 1) / (4 * pot), tiles_y = (h + 4 * pot - 1) / (4 * pot);
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const int tile = gt >> 4, s = gt & 15;
    if ((gt & ~31) >= tiles_x * tiles_y * 16) return;   // whole warp out of range
    const float *g0 = d.g0 + (size_t)f * d.npx_pad;
    const float *g1 = d.g1 + (size_t)f * d.n1;
    const float *g2 = d.g2 + (size_t)f * d.n2;
    const float *thsS = d.thsS + (size_t)f * d.nb;

    float b0 = -1.f, b1 = -1.f, b2 = -1.f;
    int i0 = -1, i1 = -1, i2 = -1;
    bool any0 = false, any1 = false;
    if (tile < tiles_x * tiles_y) {
        const int ty = tile / tiles_x, tx = tile % tiles_x;
        const int x0 = tx * 4 * pot + ((s >> 2) & 1) * 2 * pot + (s & 1) * pot;
        const int y0 = ty * 4 * pot + ((s >> 3) & 1) * 2 * pot + ((s >> 1) & 1) * pot;
        if (x0 < w && y0 < h) {
            const int mx = min(pot, w - x0), my = min(pot, h - y0);
            for (int y1 = 0; y1 < my; y1++)
                for (int x1 = 0; x1 < mx; x1++) {
                    const int xf = x0 + x1, yf = y0 + y1;
                    if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;   // :364
                    const int idx = xf + w * yf;
                    const float TH0 = thsS[(xf >> 5) + (yf >> 5) * d.w32];
                    const float TH1 = __fmul_rn(TH0, 0.75f);
                    const float TH2 = __fmul_rn(TH1, 0.5625f);
                    const float ag0 = g0[idx];
                    if (ag0 > TH0) { any0 = true; if (ag0 > b0) { b0 = ag0; i0 = idx; } }
                    const float ag1 = g1[(xf >> 1) + (yf >> 1) * d.w1];   // (int)(xf*0.5f+0.25f)
                    if (ag1 > TH1) { any1 = true; if (ag1 > b1) { b1 = ag1; i1 = idx; } }
                    const float ag2 = g2[(xf >> 2) + (yf >> 2) * d.w2];   // (int)(xf*0.25f+0.125)
                    if (ag2 > TH2) { if (ag2 > b2) { b2 = ag2; i2 = idx; } }
                }
        }
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, any0);
    const unsigned m1 = __ballot_sync(0xffffffffu, any1);
    const unsigned quad = 0xFu << (lane & ~3u), hexm = 0xFFFFu << (lane & 16u);
    // level 0: every pot-block keeps its own winner
    const bool sel0 = i0 >= 0;
    if (sel0) map[i0] = 1;
    // level 1: the 2pot-tile (4 lanes) keeps one winner iff no level-0 hit inside it
    take_better(b1, i1, lane, 1);
    take_better(b1, i1, lane, 2);
    const bool sel1 = ((lane & 3) == 0) && !(m0 & quad) && i1 >= 0;
    if (sel1) map[i1] = 2;
    // level 2: the 4pot-tile (16 lanes) keeps one winner iff no level-0/1 hit inside it
    take_better(b2, i2, lane, 1);
    take_better(b2, i2, lane, 2);
    take_better(b2, i2, lane, 4);
    take_better(b2, i2, lane, 8);
    const bool sel2 = ((lane & 15) == 0) && !((m0 | m1) & hexm) && i2 >= 0;
    if (sel2) map[i2] = 4;
    const int c0 = __popc(__ballot_sync(0xffffffffu, sel0));
    const int c1 = __popc(__ballot_sync(0xffffffffu, sel1));
    const int c2 = __popc(__ballot_sync(0xffffffffu, sel2));
    if (lane == 0) {
        if (c0) atomicAdd(&n_out[0], c0);
        if (c1) atomicAdd(&n_out[1], c1);
        if (c2) atomicAdd(&n_out[2], c2);
    }
}

// ------------------------------------------------------------------------------------------------
// K4: sub-sampling + depth filter + raster-order compaction + back-projection + features
// (PixelSelector2.cpp:226-244, pcd_generator.cpp:456-499, :563-616).  One CTA per frame; each
// thread owns a contiguous run of pixels, so raster order needs only two block scans.
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_tot, int *total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int wv = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0;
        int winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        warp_tot[lane] = winc - wv;          // exclusive prefix of warp totals
        if (lane == 31) warp_tot[32] = winc;  // grand total
    }
    __syncthreads();
    int res = warp_tot[wid] + inc - v;
    *total = warp_tot[32];
    __syncthreads();
    return res;
}

// The same scan across the kCompactCtas CTAs of a thread-block cluster (one cluster per frame): block scan, the
// CTA totals are exchanged through distributed shared memory.  `xch` is a per-CTA shared slot (one per call site,
// so that a slot is never rewritten while a slower peer still reads it).
constexpr int kCompactCtas = 8;
__device__ __forceinline__ int cluster_exclusive_scan(int v, int *warp_tot, int *xch, int *total) {
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    int cta_total;
    const int local = block_exclusive_scan(v, warp_tot, &cta_total);
    if (threadIdx.x == 0) *xch = cta_total;
    cl.sync();
    int before = 0, all = 0;
    const unsigned me = cl.block_rank();
    for (unsigned r = 0; r < cl.num_blocks(); r++) {
        const int t = *cl.map_shared_rank(xch, r);
        if (r < me) before += t;
        all += t;
    }
    *total = all;
    return before + local;
}

struct ArenaDev {
    float4 *pos;
    float4 *f03;
    float *f4;
    float2 *pix;
    int *n;
    int *ovf;
    int cap;
    int first;
};

__global__ void __launch_bounds__(1024) k_compact(SelDev d, ArenaDev A, const uint8_t *__restrict__ bgr,
                                                  size_t bgr_stride, const uint16_t *__restrict__ depth,
                                                  size_t depth_stride) {
    __shared__ int warp_tot[33];
    __shared__ int s_use_b, s_sub, s_charTH;
    __shared__ int s_xch[3];   // CTA totals of the three scans, read by the peers of the cluster
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int crank = (int)cl.block_rank(), csize = (int)cl.num_blocks();
    const int f = blockIdx.x / csize;   // one cluster per frame: 8 x 1024 threads share the raster scan
    SelState *st = d.st + f;
    if (threadIdx.x == 0) {
        int need2, pot2;
        decide_pass2(st->n_a, d.num_want, 3, need2, pot2);
        const int *n = need2 ? st->n_b : st->n_a;
        float numHave = (float)(n[0] + n[1] + n[2]);
        float quotia = __fdiv_rn(d.num_want, numHave);
        int sub = (double)quotia < 0.95;
        int charTH = sub ? (int)(unsigned char)__fmul_rn(255.f, quotia) : 255;
        if (crank == 0) {   // (every CTA of the cluster derives the same decision; one stores it)
            st->need2 = need2;
            st->pot2 = pot2;
            st->use_b = need2;
            st->subsample = sub;
            st->charTH = charTH;
        }
        s_use_b = need2;
        s_sub = sub;
        s_charTH = charTH;
    }
    __syncthreads();
    uint8_t *map = (s_use_b ? d.map_b : d.map_a) + (size_t)f * d.npx_pad;
    const uint32_t *map32 = reinterpret_cast<const uint32_t *>(map);
    const int sub = s_sub, charTH = s_charTH;
    const int nwords = d.npx_pad / 4;
    const int nthr = (int)blockDim.x * csize, gthr = crank * (int)blockDim.x + (int)threadIdx.x;
    const int per = (nwords + nthr - 1) / nthr;
    const int w0 = min(gthr * per, nwords), w1 = min(w0 + per, nwords);
    const uint16_t *dep = depth + (size_t)f * depth_stride;
    const uint8_t *img = bgr + (size_t)f * bgr_stride;
    const uint8_t *gray = d.gray + (size_t)f * d.npx_pad;

    // pass 1: selected pixels in my run -> running rank `rn` base
    int c = 0;
    for (int k = w0; k < w1; k++) c += __popc(__vcmpne4(map32[k], 0u)) >> 3;
    int total_sel;
    const int rn0 = cluster_exclusive_scan(c, warp_tot, &s_xch[0], &total_sel);

    // pass 2: apply randomPattern[rn] > charTH removal and the depth filter; count survivors
    int rn = rn0, kept = 0, kept_sub = 0;
    for (int k = w0; k < w1; k++) {
        uint32_t v = map32[k];
        if (!v) continue;
        for (int b = 0; b < 4; b++) {
            if (!((v >> (8 * b)) & 255u)) continue;
            const int idx = 4 * k + b;
            bool keep = !(sub && (int)d.rnd[rn] > charTH);
            rn++;
            if (!keep) { map[idx] = 0; continue; }   // map after sub-sampling (debug parity)
            kept_sub++;
            if (dep[idx] != 0) kept++;
        }
    }
    __syncthreads();
    int total_sub, total_out;
    cluster_exclusive_scan(kept_sub, warp_tot, &s_xch[1], &total_sub);
    const int out0 = cluster_exclusive_scan(kept, warp_tot, &s_xch[2], &total_out);

    // pass 3: survivors in raster order -> position, pixel, features
    const size_t base = (size_t)(A.first + f) * A.cap;
    int o = out0;
    for (int k = w0; k < w1; k++) {
        uint32_t v = map32[k];
        if (!v) continue;
        for (int b = 0; b < 4; b++) {
            if (!((v >> (8 * b)) & 255u)) continue;
            const int idx = 4 * k + b;
            const uint16_t dp = dep[idx];
            if (dp == 0) continue;
            if (o < A.cap) {
                const int x = idx % d.w, y = idx / d.w;
                // pcd_generator.cpp:473-476, same operation order, IEEE division
                const float z = __fdiv_rn((float)dp, d.scaling_factor);
                const float X = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, d.cx), z), d.fx);
                const float Y = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, d.cy), z), d.fy);
                A.pos[base + o] = make_float4(X, Y, z, 0.f);
                A.pix[base + o] = make_float2((float)x, (float)y);
                float dx = 0.f, dy = 0.f;   // level-0 gradient (pcd_generator.cpp:121-122)
                if (idx >= d.w && idx < d.w * (d.h - 1)) {
                    dx = __fmul_rn(0.5f, __fsub_rn((float)gray[idx + 1], (float)gray[idx - 1]));
                    dy = __fmul_rn(0.5f, __fsub_rn((float)gray[idx + d.w], (float)gray[idx - d.w]));
                }
                const uint32_t c0 = img[3 * (size_t)idx], c1 = img[3 * (size_t)idx + 1], c2 = img[3 * (size_t)idx + 2];
                float4 fa;
                float f4;
                if (d.feature_type == 0) {   // HSV + gradient, normalised (:570-592)
                    int r = c0, g = c1, bl = c2;   // "R" is stored channel 0 (RGB code on BGR data)
                    int vmax = max(r, max(g, bl)), vmin = min(r, min(g, bl));
                    int diff = vmax - vmin;
                    int vr = (vmax == r) ? -1 : 0, vg = (vmax == g) ? -1 : 0;
                    int s = (diff * c_sdiv[vmax] + (1 << 11)) >> 12;
                    int hh = (vr & (g - bl)) + (~vr & ((vg & (bl - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
                    hh = (hh * c_hdiv[diff] + (1 << 11)) >> 12;
                    if (hh < 0) hh += 180;
                    fa.x = (float)__ddiv_rn((double)hh, 180.0);
                    fa.y = (float)__ddiv_rn((double)s, 255.0);
                    fa.z = (float)__ddiv_rn((double)vmax, 255.0);
                    fa.w = (float)__dmul_rn(__ddiv_rn((double)dx, 255.0), 2.0);
                    f4 = (float)__dmul_rn(__ddiv_rn((double)dy, 255.0), 2.0);
                } else {                      // raw channel bytes + raw gradient (:593-615)
                    fa = make_float4((float)c0, (float)c1, (float)c2, dx);
                    f4 = dy;
                }
                A.f03[base + o] = fa;
                A.f4[base + o] = f4;
            }
            o++;
        }
    }
    if (threadIdx.x == 0 && crank == 0) {
        st->n_selected = total_sub;
        st->n_out = total_out;
        st->overflow = total_out > A.cap;
        A.n[A.first + f] = min(total_out, A.cap);
        A.ovf[A.first + f] = total_out > A.cap ? 1 : 0;   // surfaces as CVO_ERR_CAPACITY / CVO_ERR_PAIR_OVERFLOW downstream
    }
    cl.sync();   // no CTA leaves while a peer may still read its scan totals
}

// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sel_create(SelWorkspace **out, int w, int h, int chunk) {
    if (w < 64 || h < 64 || chunk < 1) return CVO_ERR_INVALID;
    SelWorkspace *ws = new SelWorkspace();
    ws->chunk = chunk;
    SelDev &d = ws->d;
    memset(&d, 0, sizeof(d));
    d.w = w; d.h = h; d.w1 = w / 2; d.h1 = h / 2; d.w2 = d.w1 / 2; d.h2 = d.h1 / 2;
    d.w32 = w / 32; d.h32 = h / 32; d.nb = d.w32 * d.h32;
    d.npx = w * h;
    d.npx_pad = (int)align_up((size_t)d.npx, 16);
    d.n1 = (size_t)d.w1 * d.h1;
    d.n2 = (size_t)d.w2 * d.h2;
    ws->bgr_stride = align_up((size_t)d.npx * 3, 16);
    ws->depth_stride = align_up((size_t)d.npx, 8);   // in elements
    // one allocation, carved
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_bgr = carve(ws->bgr_stride * chunk);
    size_t o_dep = carve(ws->depth_stride * 2 * chunk);
    size_t o_gray = carve((size_t)d.npx_pad * chunk);
    size_t o_I1 = carve(d.n1 * 4 * chunk), o_I2 = carve(d.n2 * 4 * chunk);
    size_t o_g0 = carve((size_t)d.npx_pad * 4 * chunk), o_g1 = carve(d.n1 * 4 * chunk), o_g2 = carve(d.n2 * 4 * chunk);
    size_t o_ths = carve((size_t)d.nb * 4 * chunk), o_thsS = carve((size_t)d.nb * 4 * chunk);
    size_t o_map = carve((size_t)d.npx_pad * 2 * chunk);
    size_t o_st = carve(sizeof(SelState) * chunk);
    size_t o_rnd = carve((size_t)d.npx);
    char *blob = nullptr;
    if (cudaMalloc(&blob, off) != cudaSuccess) {
        set_last_error("sel_create: cudaMalloc(%zu) failed", off);
        delete ws;
        return CVO_ERR_CUDA;
    }
    ws->blob = blob;
    ws->bgr = (uint8_t *)(blob + o_bgr);
    ws->depth = (uint16_t *)(blob + o_dep);
    d.gray = (uint8_t *)(blob + o_gray);
    d.I1 = (float *)(blob + o_I1); d.I2 = (float *)(blob + o_I2);
    d.g0 = (float *)(blob + o_g0); d.g1 = (float *)(blob + o_g1); d.g2 = (float *)(blob + o_g2);
    d.ths = (float *)(blob + o_ths); d.thsS = (float *)(blob + o_thsS);
    d.map_a = (uint8_t *)(blob + o_map);
    d.map_b = d.map_a + (size_t)d.npx_pad * chunk;
    d.st = (SelState *)(blob + o_st);
    ws->rnd = (uint8_t *)(blob + o_rnd);
    d.rnd = ws->rnd;
    {
        uint8_t *tab = new uint8_t[d.npx];
        host_random_pattern(tab, d.npx);
        cudaError_t e = cudaMemcpy(ws->rnd, tab, d.npx, cudaMemcpyHostToDevice);
        delete[] tab;
        int sdiv[256], hdiv[256];
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) {
            sdiv[i] = (int)lrint((255 << 12) / (double)i);
            hdiv[i] = (int)lrint((180 << 12) / (6.0 * i));
        }
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv));
        if (e != cudaSuccess) {
            set_last_error("sel_create: table upload failed: %s", cudaGetErrorString(e));
            cudaFree(blob);
            delete ws;
            return CVO_ERR_CUDA;
        }
    }
    *out = ws;
    return CVO_OK;
}

void sel_destroy(SelWorkspace *ws) {
    if (!ws) return;
    cudaFree(ws->blob);
    delete ws;
}

size_t sel_frame_bytes_bgr(const SelWorkspace *ws) { return ws->bgr_stride; }
uint8_t *sel_bgr_ptr(SelWorkspace *ws, int k) { return ws->bgr + ws->bgr_stride * k; }
uint16_t *sel_depth_ptr(SelWorkspace *ws, int k) { return ws->depth + ws->depth_stride * k; }

int sel_run(SelWorkspace *ws, int n, const uint8_t *bgr_dev, const uint16_t *depth_dev,
            const cvo_calib &cal, const cvo_params &prm, const CloudArena &arena, int first,
            cudaStream_t stream, int64_t *launches) {
    if (n < 1 || n > ws->chunk || first < 0 || first + n > arena.frames) return CVO_ERR_INVALID;
    SelDev d = ws->d;
    d.num_want = (float)prm.num_want;
    d.gray_mode = prm.gray_mode;
    d.feature_type = prm.feature_type;
    d.scaling_factor = cal.scaling_factor; d.fx = cal.fx; d.fy = cal.fy; d.cx = cal.cx; d.cy = cal.cy;
    size_t bgr_stride, depth_stride;
    if (bgr_dev == nullptr) {   // frames staged in the workspace
        bgr_dev = ws->bgr; depth_dev = ws->depth;
        bgr_stride = ws->bgr_stride; depth_stride = ws->depth_stride;
    } else {                    // tightly packed external device images
        bgr_stride = (size_t)d.npx * 3; depth_stride = (size_t)d.npx;
    }
    CVO_CUDA_TRY(cudaMemsetAsync(d.map_a, 0, (size_t)d.npx_pad * 2 * ws->chunk, stream));
    CVO_CUDA_TRY(cudaMemsetAsync(d.st, 0, sizeof(SelState) * n, stream));
    const unsigned fn = (unsigned)n;
    k_gray<<<dim3((d.npx / 4 + 1 + 255) / 256, fn), 256, 0, stream>>>(d, bgr_dev, bgr_stride);
    k_pyr<<<dim3((unsigned)((d.n1 + 255) / 256), fn), 256, 0, stream>>>(d, 1);
    k_pyr<<<dim3((unsigned)((d.n2 + 255) / 256), fn), 256, 0, stream>>>(d, 2);
    const size_t ng = (size_t)d.npx + d.n1 + d.n2;
    k_grad<<<dim3((unsigned)((ng + 255) / 256), fn), 256, 0, stream>>>(d);
    k_hist<<<dim3(d.nb, fn), 256, 0, stream>>>(d);
    k_smooth<<<dim3((d.nb + 127) / 128, fn), 128, 0, stream>>>(d);
    {   // pass 1 at pot = 3; pass 2 sized for the smallest pot (1) and exits early if unused
        int tiles = ((d.w + 11) / 12) * ((d.h + 11) / 12);
        k_select<<<dim3((tiles * 16 + 255) / 256, fn), 256, 0, stream>>>(d, 0);
        tiles = ((d.w + 3) / 4) * ((d.h + 3) / 4);
        k_select<<<dim3((tiles * 16 + 255) / 256, fn), 256, 0, stream>>>(d, 1);
    }
    ArenaDev A{arena.pos, arena.f03, arena.f4, arena.pix, arena.n, arena.ovf, arena.cap, first};
    {   // one cluster of kCompactCtas CTAs per frame
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(fn * kCompactCtas));
        cfg.blockDim = dim3(1024);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCompactCtas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CVO_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_compact, d, A, bgr_dev, bgr_stride, depth_dev, depth_stride));
    }
    if (launches) *launches += 9;
    CVO_CUDA_TRY(cudaGetLastError());
    return CVO_OK;
}

// the 8-bit gray image (cv::cvtColor RGB2GRAY, row A) of chunk-local frame k as the selection left it on the device
const uint8_t *sel_gray_ptr(const SelWorkspace *ws, int k) {
    if (!ws || k < 0 || k >= ws->chunk) return nullptr;
    return ws->d.gray + (size_t)k * ws->d.npx_pad;
}

int sel_debug(SelWorkspace *ws, int k, uint8_t *map_host, int32_t info[5], cudaStream_t stream) {
    if (k < 0 || k >= ws->chunk) return CVO_ERR_INVALID;
    SelState st;
    CVO_CUDA_TRY(cudaMemcpyAsync(&st, ws->d.st + k, sizeof(st), cudaMemcpyDeviceToHost, stream));
    CVO_CUDA_TRY(cudaStreamSynchronize(stream));
    const uint8_t *map = (st.use_b ? ws->d.map_b : ws->d.map_a) + (size_t)k * ws->d.npx_pad;
    if (map_host) {
        CVO_CUDA_TRY(cudaMemcpyAsync(map_host, map, ws->d.npx, cudaMemcpyDeviceToHost, stream));
        CVO_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    const int *n = st.use_b ? st.n_b : st.n_a;
    info[0] = n[0]; info[1] = n[1]; info[2] = n[2];
    info[3] = st.use_b ? st.pot2 : 3;
    info[4] = st.use_b ? 2 : 1;
    return CVO_OK;
}

}  // namespace cvo_b200
