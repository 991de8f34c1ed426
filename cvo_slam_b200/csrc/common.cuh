// common.cuh — shared host/device definitions of libcvo_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cvo_b200.h"

namespace cvo_b200 {

// ---- error plumbing -----------------------------------------------------------------------------
void set_last_error(const char *fmt, ...);

#define CVO_CUDA_TRY(expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cvo_b200::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                  \
                                     cudaGetErrorString(_e));                                      \
            return CVO_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

// ---- device clouds ------------------------------------------------------------------------------
// One cloud = SoA arrays of capacity `cap`.  pos.w is unused (0).  Features are 5 floats
// (data_type.h:72): f03 = {f0,f1,f2,f3}, f4 separately, so that a candidate costs one 16-byte
// load for the distance test and 20 more bytes only when it is inside the cutoff.
struct CloudView {
    const float4 *pos;
    const float4 *f03;
    const float *f4;
    const float2 *pix;   // selected pixel (x, y)  (frame::selected_points)
    const int *n;        // device-resident point count
    const int *ovf;      // != 0: the selection found more points than the arena holds (the cloud is truncated)
};

// F clouds of equal capacity in one allocation (frame k at offset k*cap in every array).
// A cvo handle owns an arena of 3 (fixed / moving / previous map onto arena indices);
// a batch owns one of max_frames.
struct CloudArena {
    float4 *pos = nullptr;
    float4 *f03 = nullptr;
    float *f4 = nullptr;
    float2 *pix = nullptr;
    int *n = nullptr;       // [frames]
    int *ovf = nullptr;     // [frames] selection overflow flags
    int cap = 0;
    int frames = 0;
    CloudView view(int k) const {
        size_t o = (size_t)k * cap;
        return CloudView{pos + o, f03 + o, f4 + o, pix + o, n + k, ovf + k};
    }
};

int arena_alloc(CloudArena &a, int frames, int cap);
void arena_free(CloudArena &a);

inline int cloud_capacity_for(const cvo_params &p, int w, int h) {
    // makeMaps keeps <= numWant/0.95 points without sub-sampling and ~numWant (+1/256 of the
    // selected set, + noise) with it (PixelSelector2.cpp:226-244)
    long cap = (long)(p.num_want * 1.25) + 512;
    long px = (long)w * h;
    if (cap > px) cap = px;
    if (cap < 1024) cap = 1024;
    return (int)((cap + 31) / 32 * 32);
}

// ---- selection (select.cu) ----------------------------------------------------------------------
struct SelWorkspace;   // opaque; per-chunk scratch for `frames_per_chunk` frames
int sel_create(SelWorkspace **ws, int width, int height, int frames_per_chunk);
void sel_destroy(SelWorkspace *ws);
size_t sel_frame_bytes_bgr(const SelWorkspace *ws);
// device staging buffers for chunk-local frame `k`
uint8_t *sel_bgr_ptr(SelWorkspace *ws, int k);
uint16_t *sel_depth_ptr(SelWorkspace *ws, int k);
// Runs point selection + features for `n` (<= frames_per_chunk) frames given as tightly packed
// device images (the workspace's own staging buffers or external pointers) and writes arena
// clouds first..first+n-1.  All asynchronous on `stream`.
int sel_run(SelWorkspace *ws, int n, const uint8_t *bgr_dev, const uint16_t *depth_dev,
            const cvo_calib &cal, const cvo_params &prm, const CloudArena &arena, int first,
            cudaStream_t stream, int64_t *launch_counter);
// debug: status map (after sub-sampling) and {n2,n3,n4,pot,passes} of chunk-local frame k
int sel_debug(SelWorkspace *ws, int k, uint8_t *map_host, int32_t info[5], cudaStream_t stream);
const uint8_t *sel_gray_ptr(const SelWorkspace *ws, int k);   // device gray image of chunk-local frame k (tightly packed rows)
void host_random_pattern(uint8_t *out, int n);

// ---- alignment (align.cu) -----------------------------------------------------------------------
struct AlignTask {          // one frame pair
    CloudView fixed, moving;
    float R[9], T[3];
    float ell;
};

struct AlignWorkspace;      // opaque scratch for `n_workgroups` concurrent pairs
// max_workgroups > 0 bounds the number of resident CTAs the scratch is sized for (a handle only
// ever runs one pair, i.e. one cluster); 0 = fill the GPU (batches)
int align_ws_create(AlignWorkspace **ws, int max_points, int device, int max_workgroups, int coop_ctas = 0);
void align_ws_destroy(AlignWorkspace *ws);
int align_ws_max_points(const AlignWorkspace *ws);
// Runs tasks[0..n) (device array) -> results[0..n) (device array).  trace may be null.
// single_iteration: evaluate exactly one compute_flow + compute_step_size, do not update.
int align_run(AlignWorkspace *ws, const cvo_params &prm, int n_tasks, const AlignTask *tasks_dev,
              cvo_align_result *results_dev, cvo_iter_record *trace_dev, int trace_cap,
              bool single_iteration, cudaStream_t stream, int64_t *launch_counter);
// debug: in-cutoff pattern left in workgroup 0's scratch by the last run (host arrays)
int align_last_pattern(AlignWorkspace *ws, int nnz, int32_t *ij, float *a, int cap, int *n,
                       cudaStream_t stream);
// cumulative {kernel evals, iterations, sum of nnz over iterations}
void align_ws_stats(AlignWorkspace *ws, cudaStream_t stream, int64_t out[3]);
void align_ws_phase_cycles(AlignWorkspace *ws, cudaStream_t stream, int64_t out[8]);

struct QueryTask {          // <Ta * a, b> at `ell`
    CloudView a, b;
    float Ta[12];           // 3x4 row-major
    float ell;
    int kind;               // 0 = inner product, 1 = Hessian
};
struct QueryOut {
    double sum;             // sum of a_ij           (kind 0)
    double H[21];           // upper triangle blocks (kind 1): A(6) C(9) D(6)
    int count;
};
int query_run(AlignWorkspace *ws, const cvo_params &prm, int n, const QueryTask *tasks_dev,
              QueryOut *out_dev, cudaStream_t stream, int64_t *launch_counter);
// host: cvo.cpp:726-758 on the accumulated (unscaled) Hessian
void finish_hessian_host(const QueryOut &q, double H[36]);

// compute_innerproduct_lc (cvo.cpp:505-561) for one pair: six queries <T_k a, b> that share the
// grid over b.  k = 0..3 inner products (prior, lc_prior, identity, lc), k = 4, 5 Hessians (lc,
// lc_prior_2); only the first Hessian's matrix is used (cvo.cpp:555,558).
struct LcTask {
    CloudView a, b;         // a = moving, b = fixed
    float T[6][12];         // 3x4 row-major each
    float ell;
};
struct LcOut {
    int truncated;          // a cloud of the pair was truncated (selection overflow / larger than the scratch)
    double sum[4];
    double H[36];           // post_hessian, scaled and eigenvalue-shifted (cvo.cpp:726-758)
    int count[4];
    int inliers[2];
};
int lc_run(AlignWorkspace *ws, const cvo_params &prm, int n, const LcTask *tasks_dev, LcOut *out_dev,
           cudaStream_t stream, int64_t *launch_counter);

}  // namespace cvo_b200
