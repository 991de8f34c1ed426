// multi.cu — one list of frames and frame pairs over several GPUs of one process (SURVEY §8b:
// cvo_align_batch(ctx, n_pairs, descs, results, n_devices); §8e: pairs shard, no collective).
//
// The caller of the loop-closure candidate loop (src/keyframe_graph.cpp:622-731) holds host images
// and a list of independent (fixed, moving) pairs.  Pairs are split into n_devices contiguous
// blocks: a block of consecutive pairs touches a compact set of frames (in the verification pattern
// consecutive pairs share their fixed keyframe), so each device selects points only for the frames
// its own block needs — a round-robin split would make every device select (almost) every frame.
// Iteration-count variance inside a block is absorbed by the per-GPU pair queue of k_align_batch.
// One host thread per device drives that device's cvo_batch (its own streams); nothing is exchanged
// between devices, results are written straight into the caller's array at the pairs' positions.
#include "common.cuh"

#include <algorithm>
#include <string.h>
#include <thread>
#include <vector>

using namespace cvo_b200;

struct cvo_multi {
    int n_devices = 0;
    int w = 0, h = 0, max_frames = 0, max_pairs = 0;
    std::vector<int> devices;
    std::vector<cvo_batch *> batch;
    std::vector<int> last_frames, last_pairs;   // per device: frames selected / pairs aligned by the last call
    std::vector<float> last_ms;                 // per device: device time of the last call's align kernel
};

extern "C" {

int cvo_multi_create(const cvo_calib *calib, const cvo_params *params, int n_devices, const int *devices,
                     int max_frames, int max_pairs, int width, int height, cvo_multi **out) {
    if (!calib || !out || n_devices < 1 || max_frames < 1 || max_pairs < 1) return CVO_ERR_INVALID;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { set_last_error("cvo_multi_create: no CUDA device"); return CVO_ERR_CUDA; }
    cvo_multi *m = new cvo_multi();
    m->n_devices = n_devices;
    m->w = width; m->h = height; m->max_frames = max_frames; m->max_pairs = max_pairs;
    for (int d = 0; d < n_devices; d++) {
        const int dev = devices ? devices[d] : d;
        if (dev < 0 || dev >= have) { set_last_error("cvo_multi_create: device %d of %d", dev, have); delete m; return CVO_ERR_INVALID; }
        m->devices.push_back(dev);
    }
    m->batch.assign(n_devices, nullptr);
    m->last_frames.assign(n_devices, 0);
    m->last_pairs.assign(n_devices, 0);
    m->last_ms.assign(n_devices, 0.f);
    // a device's block holds ceil(max_pairs / n) pairs and at most min(max_frames, 2 * pairs) frames
    const int pairs_per = (max_pairs + n_devices - 1) / n_devices;
    const int frames_per = std::min(max_frames, 2 * pairs_per);
    for (int d = 0; d < n_devices; d++) {
        int rc = cvo_batch_create(calib, params, m->devices[d], frames_per, pairs_per, width, height, &m->batch[d]);
        if (rc != CVO_OK) {
            for (cvo_batch *b : m->batch) cvo_batch_destroy(b);
            delete m;
            return rc;
        }
    }
    *out = m;
    return CVO_OK;
}

int cvo_multi_destroy(cvo_multi *m) {
    if (!m) return CVO_OK;
    for (cvo_batch *b : m->batch) cvo_batch_destroy(b);
    delete m;
    return CVO_OK;
}

// block of pairs [lo, hi) that device d of n owns
static inline void block_of(int n_pairs, int n, int d, int &lo, int &hi) {
    const long per = (n_pairs + n - 1) / n;
    lo = (int)std::min<long>((long)n_pairs, per * d);
    hi = (int)std::min<long>((long)n_pairs, per * (d + 1));
}

int cvo_multi_align(cvo_multi *m, int n_frames, const uint8_t *bgr, const uint16_t *depth, int n_pairs,
                    const cvo_pair_desc *pairs, cvo_align_result *results, float *values, int *nums) {
    if (!m || !bgr || !depth || !pairs || !results || n_frames < 1 || n_frames > m->max_frames || n_pairs < 0 ||
        n_pairs > m->max_pairs || (values && !nums) || (nums && !values))
        return CVO_ERR_INVALID;
    for (int i = 0; i < n_pairs; i++)
        if (pairs[i].fixed_frame < 0 || pairs[i].fixed_frame >= n_frames || pairs[i].moving_frame < 0 ||
            pairs[i].moving_frame >= n_frames)
            return CVO_ERR_INVALID;
    const size_t fb = (size_t)m->w * m->h * 3, fd = (size_t)m->w * m->h;
    std::vector<int> rcs(m->n_devices, CVO_OK);
    auto work = [&](int d) {
        int lo, hi;
        block_of(n_pairs, m->n_devices, d, lo, hi);
        m->last_pairs[d] = hi - lo;
        m->last_frames[d] = 0;
        m->last_ms[d] = 0.f;
        if (hi <= lo) return;
        cvo_batch *b = m->batch[d];
        // frames this block touches, sorted; local index = rank in that list
        std::vector<int> need;
        need.reserve(2 * (size_t)(hi - lo));
        for (int i = lo; i < hi; i++) { need.push_back(pairs[i].fixed_frame); need.push_back(pairs[i].moving_frame); }
        std::sort(need.begin(), need.end());
        need.erase(std::unique(need.begin(), need.end()), need.end());
        m->last_frames[d] = (int)need.size();
        std::vector<int> local(n_frames, -1);
        for (size_t k = 0; k < need.size(); k++) local[need[k]] = (int)k;
        // upload + select: one call per run of consecutive frame ids (the images are contiguous there)
        int rc = CVO_OK;
        for (size_t k = 0; k < need.size() && rc == CVO_OK;) {
            size_t e = k + 1;
            while (e < need.size() && need[e] == need[e - 1] + 1) e++;
            rc = cvo_batch_set_frames(b, (int)k, (int)(e - k), bgr + (size_t)need[k] * fb, depth + (size_t)need[k] * fd);
            k = e;
        }
        std::vector<cvo_pair_desc> mine(pairs + lo, pairs + hi);
        for (cvo_pair_desc &p : mine) { p.fixed_frame = local[p.fixed_frame]; p.moving_frame = local[p.moving_frame]; }
        if (rc == CVO_OK) rc = cvo_batch_align(b, hi - lo, mine.data(), results + lo);
        if (rc == CVO_OK) cvo_batch_last_align_ms(b, &m->last_ms[d]);
        if (rc == CVO_OK && values) rc = cvo_batch_inner_product(b, hi - lo, mine.data(), results + lo, values + lo, nums + lo);
        rcs[d] = rc;
    };
    if (m->n_devices == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < m->n_devices; d++) th.emplace_back(work, d);
        for (std::thread &t : th) t.join();
    }
    for (int rc : rcs)
        if (rc != CVO_OK) return rc;
    return CVO_OK;
}

int cvo_multi_last_shares(cvo_multi *m, int *frames, int *pairs, float *align_ms) {
    if (!m) return CVO_ERR_INVALID;
    for (int d = 0; d < m->n_devices; d++) {
        if (frames) frames[d] = m->last_frames[d];
        if (pairs) pairs[d] = m->last_pairs[d];
        if (align_ms) align_ms[d] = m->last_ms[d];
    }
    return CVO_OK;
}

}  // extern "C"
