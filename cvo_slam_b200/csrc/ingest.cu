// ingest.cu — image ingest for the alignment path: the reference loads every frame with
// cv::imread (src/run_SLAM.cpp:134-143: colour PNG -> 8-bit BGR, depth PNG with ANYDEPTH -> 16-bit) and hands
// host cv::Mats to cvo::set_pcd.  With the alignment at ~1.5 ms a 5-10 ms single-threaded decode is the
// largest cost of a frame, so the library offers
//   * cvo_png_decode_*: a PNG decoder (container + the five scan-line filters here, DEFLATE by the system
//     zlib) that produces exactly the buffers cv::imread produces for the TUM / ETH3D files: BGR8 and u16;
//   * cvo_set_frame_png: colour and depth decoded on two threads straight into pinned staging memory, then
//     the normal cvo_set_frame path (H2D + selection on the handle's stream);
//   * cvo_ingest_*: a prefetcher with its own worker threads and a ring of pinned frames — the caller submits
//     frame k+1's PNG bytes and goes on aligning frame k; the decoded frame is waiting (in pinned memory, so
//     the H2D copy is asynchronous) when it is needed.
// Host-only code except for the pinned allocations.
#include "common.cuh"

#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>
#include <zlib.h>

using namespace cvo_b200;

namespace {

struct PngHeader {
    int w = 0, h = 0, depth = 0, color = 0, channels = 0;
};

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int png_header(const uint8_t *png, size_t n, PngHeader &hd) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (!png || n < 33 || memcmp(png, sig, 8) != 0 || be32(png + 8) != 13 || memcmp(png + 12, "IHDR", 4) != 0) {
        set_last_error("png: not a PNG stream");
        return CVO_ERR_INVALID;
    }
    hd.w = (int)be32(png + 16);
    hd.h = (int)be32(png + 20);
    hd.depth = png[24];
    hd.color = png[25];
    if (png[26] != 0 || png[27] != 0 || png[28] != 0) { set_last_error("png: interlaced / unknown compression or filter method"); return CVO_ERR_INVALID; }
    switch (hd.color) {
        case 0: hd.channels = 1; break;
        case 2: hd.channels = 3; break;
        case 4: hd.channels = 2; break;
        case 6: hd.channels = 4; break;
        default: set_last_error("png: colour type %d (palette) is not supported", hd.color); return CVO_ERR_INVALID;
    }
    if ((hd.depth != 8 && hd.depth != 16) || hd.w < 1 || hd.h < 1 || hd.w > 16384 || hd.h > 16384) {
        set_last_error("png: %d x %d, bit depth %d is not supported", hd.w, hd.h, hd.depth);
        return CVO_ERR_INVALID;
    }
    return CVO_OK;
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// -> raw samples, rows of w * channels * depth / 8 bytes (PNG byte order: 16-bit samples big endian)
int png_scanlines(const uint8_t *png, size_t n, const PngHeader &hd, std::vector<uint8_t> &rows) {
    std::vector<uint8_t> z;
    size_t off = 8;
    bool end = false;
    while (off + 12 <= n && !end) {
        const uint32_t len = be32(png + off);
        const uint8_t *type = png + off + 4, *data = png + off + 8;
        if (off + 12 + (size_t)len > n) { set_last_error("png: truncated chunk"); return CVO_ERR_INVALID; }
        if (crc32(crc32(0L, Z_NULL, 0), type, len + 4) != be32(data + len)) { set_last_error("png: chunk CRC mismatch"); return CVO_ERR_INVALID; }
        if (memcmp(type, "IDAT", 4) == 0) z.insert(z.end(), data, data + len);
        else if (memcmp(type, "IEND", 4) == 0) end = true;
        off += 12 + (size_t)len;
    }
    const int bpp = hd.channels * hd.depth / 8;
    const size_t stride = (size_t)hd.w * bpp;
    std::vector<uint8_t> raw((stride + 1) * hd.h);
    uLongf out_len = (uLongf)raw.size();
    if (z.empty() || uncompress(raw.data(), &out_len, z.data(), (uLong)z.size()) != Z_OK || out_len != raw.size()) {
        set_last_error("png: inflate failed (%zu compressed bytes)", z.size());
        return CVO_ERR_INVALID;
    }
    rows.resize(stride * hd.h);
    std::vector<uint8_t> zero(stride, 0);
    for (int y = 0; y < hd.h; y++) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *in = raw.data() + (stride + 1) * y + 1;
        uint8_t *cur = rows.data() + stride * y;
        const uint8_t *up = y ? cur - stride : zero.data();
        switch (ft) {
            case 0: memcpy(cur, in, stride); break;
            case 1: for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(in[i] + (i >= (size_t)bpp ? cur[i - bpp] : 0)); break;
            case 2: for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(in[i] + up[i]); break;
            case 3: for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(in[i] + (((i >= (size_t)bpp ? cur[i - bpp] : 0) + up[i]) >> 1)); break;
            case 4:
                for (size_t i = 0; i < stride; i++) {
                    const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
                    cur[i] = (uint8_t)(in[i] + paeth(a, b, c));
                }
                break;
            default: set_last_error("png: scan-line filter %d", ft); return CVO_ERR_INVALID;
        }
    }
    return CVO_OK;
}

// what cv::imread(path) returns: 8-bit, 3 channels, B G R (16-bit samples keep their high byte; alpha is dropped;
// gray is replicated)
int decode_bgr8(const uint8_t *png, size_t n, int *w, int *h, uint8_t *out, size_t cap) {
    PngHeader hd;
    int rc = png_header(png, n, hd);
    if (rc != CVO_OK) return rc;
    if (w) *w = hd.w;
    if (h) *h = hd.h;
    if (!out) return CVO_OK;
    if (cap < (size_t)hd.w * hd.h * 3) { set_last_error("png: output buffer too small"); return CVO_ERR_CAPACITY; }
    std::vector<uint8_t> rows;
    rc = png_scanlines(png, n, hd, rows);
    if (rc != CVO_OK) return rc;
    const int sb = hd.depth / 8, px = hd.channels * sb;
    const size_t np = (size_t)hd.w * hd.h;
    for (size_t i = 0; i < np; i++) {
        const uint8_t *s = rows.data() + i * px;
        if (hd.channels >= 3) { out[3 * i] = s[2 * sb]; out[3 * i + 1] = s[sb]; out[3 * i + 2] = s[0]; }
        else { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = s[0]; }
    }
    return CVO_OK;
}

// what cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH) returns for the 16-bit gray depth maps of TUM / ETH3D: u16, host order
int decode_gray16(const uint8_t *png, size_t n, int *w, int *h, uint16_t *out, size_t cap) {
    PngHeader hd;
    int rc = png_header(png, n, hd);
    if (rc != CVO_OK) return rc;
    if (w) *w = hd.w;
    if (h) *h = hd.h;
    if (!out) return CVO_OK;
    if (hd.depth != 16 || hd.channels != 1) { set_last_error("png: the depth image must be 16-bit gray (is %d bit, %d channels)", hd.depth, hd.channels); return CVO_ERR_INVALID; }
    if (cap < (size_t)hd.w * hd.h) { set_last_error("png: output buffer too small"); return CVO_ERR_CAPACITY; }
    std::vector<uint8_t> rows;
    rc = png_scanlines(png, n, hd, rows);
    if (rc != CVO_OK) return rc;
    const size_t np = (size_t)hd.w * hd.h;
    for (size_t i = 0; i < np; i++) out[i] = (uint16_t)((rows[2 * i] << 8) | rows[2 * i + 1]);
    return CVO_OK;
}

}  // namespace

struct cvo_ingest {
    struct Slot {
        uint8_t *bgr = nullptr;      // pinned
        uint16_t *depth = nullptr;   // pinned
        std::vector<uint8_t> png_rgb, png_depth;
        int w = 0, h = 0, rc = CVO_OK, state = 0;   // 0 free, 1 queued, 2 decoding, 3 ready
        char err[256] = "";
    };
    int w = 0, h = 0, n_slots = 0;
    std::vector<Slot> slots;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    long long submitted = 0, consumed = 0;   // frames are consumed in submission order
    bool stop = false;

    void run() {
        for (;;) {
            int k = -1;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] {
                    if (stop) return true;
                    for (int i = 0; i < n_slots; i++) if (slots[i].state == 1) return true;
                    return false;
                });
                if (stop) return;
                for (int i = 0; i < n_slots; i++) if (slots[i].state == 1) { k = i; break; }
                slots[k].state = 2;
            }
            Slot &s = slots[k];
            int w1 = 0, h1 = 0, w2 = 0, h2 = 0;
            int rc = decode_bgr8(s.png_rgb.data(), s.png_rgb.size(), &w1, &h1, s.bgr, (size_t)w * h * 3);
            if (rc == CVO_OK) rc = decode_gray16(s.png_depth.data(), s.png_depth.size(), &w2, &h2, s.depth, (size_t)w * h);
            if (rc == CVO_OK && (w1 != w || h1 != h || w2 != w || h2 != h)) {
                set_last_error("ingest: frame is %d x %d / %d x %d, the ring was created for %d x %d", w1, h1, w2, h2, w, h);
                rc = CVO_ERR_INVALID;
            }
            if (rc != CVO_OK) { strncpy(s.err, cvo_last_error(), sizeof(s.err) - 1); s.err[sizeof(s.err) - 1] = 0; }
            {
                std::lock_guard<std::mutex> lk(mu);
                s.rc = rc; s.w = w1; s.h = h1;
                s.state = 3;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" {

int cvo_png_info(const uint8_t *png, size_t n, int *width, int *height, int *channels, int *bit_depth) {
    PngHeader hd;
    int rc = png_header(png, n, hd);
    if (rc != CVO_OK) return rc;
    if (width) *width = hd.w;
    if (height) *height = hd.h;
    if (channels) *channels = hd.channels;
    if (bit_depth) *bit_depth = hd.depth;
    return CVO_OK;
}

int cvo_png_decode_bgr8(const uint8_t *png, size_t n, uint8_t *out, size_t out_bytes, int *width, int *height) {
    return decode_bgr8(png, n, width, height, out, out_bytes);
}

int cvo_png_decode_depth16(const uint8_t *png, size_t n, uint16_t *out, size_t out_elems, int *width, int *height) {
    return decode_gray16(png, n, width, height, out, out_elems);
}

int cvo_ingest_create(int width, int height, int n_slots, int n_threads, cvo_ingest **out) {
    if (!out || width < 64 || height < 64 || n_slots < 1 || n_slots > 64 || n_threads < 1 || n_threads > 64) return CVO_ERR_INVALID;
    cvo_ingest *g = new cvo_ingest();
    g->w = width; g->h = height; g->n_slots = n_slots;
    g->slots.resize(n_slots);
    for (int i = 0; i < n_slots; i++) {
        cudaError_t e = cudaMallocHost(&g->slots[i].bgr, (size_t)width * height * 3);
        if (e == cudaSuccess) e = cudaMallocHost(&g->slots[i].depth, (size_t)width * height * 2);
        if (e != cudaSuccess) {   // no CUDA device (host-only use): pageable memory still works, the copies just are not asynchronous
            cudaGetLastError();
            if (!g->slots[i].bgr) g->slots[i].bgr = (uint8_t *)malloc((size_t)width * height * 3);
            if (!g->slots[i].depth) g->slots[i].depth = (uint16_t *)malloc((size_t)width * height * 2);
            g->slots[i].err[255] = 1;   // marks malloc'ed memory
        }
    }
    for (int t = 0; t < n_threads; t++) g->workers.emplace_back([g] { g->run(); });
    *out = g;
    return CVO_OK;
}

int cvo_ingest_destroy(cvo_ingest *g) {
    if (!g) return CVO_OK;
    { std::lock_guard<std::mutex> lk(g->mu); g->stop = true; }
    g->cv_work.notify_all();
    for (std::thread &t : g->workers) t.join();
    for (auto &s : g->slots) {
        if (s.err[255] == 1) { free(s.bgr); free(s.depth); }
        else { if (s.bgr) cudaFreeHost(s.bgr); if (s.depth) cudaFreeHost(s.depth); }
    }
    delete g;
    return CVO_OK;
}

// queues one frame (the PNG bytes are copied); blocks only while all slots of the ring are occupied
int cvo_ingest_submit(cvo_ingest *g, const uint8_t *rgb_png, size_t rgb_bytes, const uint8_t *depth_png, size_t depth_bytes) {
    if (!g || !rgb_png || !depth_png || !rgb_bytes || !depth_bytes) return CVO_ERR_INVALID;
    std::unique_lock<std::mutex> lk(g->mu);
    const int k = (int)(g->submitted % g->n_slots);
    g->cv_done.wait(lk, [&] { return g->slots[k].state == 0; });
    cvo_ingest::Slot &s = g->slots[k];
    s.png_rgb.assign(rgb_png, rgb_png + rgb_bytes);
    s.png_depth.assign(depth_png, depth_png + depth_bytes);
    s.state = 1;
    g->submitted++;
    lk.unlock();
    g->cv_work.notify_one();
    return CVO_OK;
}

// the oldest submitted frame, decoded: pinned BGR8 / u16 buffers that stay valid until cvo_ingest_release
int cvo_ingest_wait(cvo_ingest *g, const uint8_t **bgr, const uint16_t **depth, int *width, int *height) {
    if (!g || !bgr || !depth) return CVO_ERR_INVALID;
    std::unique_lock<std::mutex> lk(g->mu);
    if (g->consumed >= g->submitted) { set_last_error("ingest: nothing submitted"); return CVO_ERR_NOT_INIT; }
    const int k = (int)(g->consumed % g->n_slots);
    g->cv_done.wait(lk, [&] { return g->slots[k].state == 3; });
    cvo_ingest::Slot &s = g->slots[k];
    *bgr = s.bgr;
    *depth = s.depth;
    if (width) *width = s.w;
    if (height) *height = s.h;
    if (s.rc != CVO_OK) set_last_error("%s", s.err);
    return s.rc;
}

int cvo_ingest_release(cvo_ingest *g) {
    if (!g) return CVO_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        if (g->consumed >= g->submitted) return CVO_ERR_NOT_INIT;
        g->slots[g->consumed % g->n_slots].state = 0;
        g->consumed++;
    }
    g->cv_done.notify_all();
    return CVO_OK;
}

}  // extern "C"
