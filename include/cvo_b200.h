/*
 * cvo_b200.h — C ABI of libcvo_b200.so, the B200 (sm_100a) implementation of the
 * CVO (RKHS) RGB-D frame-pair alignment hot path of bexilin/CVO-SLAM.
 *
 * The reference has no FFI layer: its trackers link the C++ class `cvo::cvo`
 * (thirdparty/cvo/include/cvo.hpp:82-282).  This header is the boundary a
 * drop-in `cvo::cvo` (include/cvo.hpp in this repo) binds to; every entry point
 * names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain pointers and sizes only, no C++/torch types.  All entry points
 * return 0 on success and a negative CVO_ERR_* code otherwise; none throws or aborts.
 * Host pointers unless the name ends in `_device`.  Matrices are row-major.
 */
#ifndef CVO_B200_H
#define CVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVO_NUM_FEATURES 5 /* thirdparty/cvo/include/data_type.h:26 */

enum {
    CVO_OK = 0,
    CVO_ERR_INVALID = -1,      /* bad argument / null pointer / bad slot          */
    CVO_ERR_CUDA = -2,         /* a CUDA runtime call failed (see cvo_last_error) */
    CVO_ERR_NOT_INIT = -3,     /* slot empty ("cvo not initialized !", cvo.cpp:463) */
    CVO_ERR_CAPACITY = -4,     /* more points than the configured capacity         */
    CVO_ERR_PAIR_OVERFLOW = -5 /* in-cutoff pair list exceeded its scratch          */
};

enum { CVO_SLOT_FIXED = 0, CVO_SLOT_MOVING = 1, CVO_SLOT_PREVIOUS = 2, CVO_NUM_SLOTS = 3 };

/* thirdparty/cvo/include/data_type.h:32-38 (camera_info), filled from
 * Camera.fx/fy/cx/cy + DepthMapFactor in cvo.cpp:58-64. */
typedef struct cvo_calib {
    float scaling_factor;
    float fx, fy, cx, cy;
} cvo_calib;

/* Every constant on the path.  Defaults = the reference's hard-coded
 * constructor constants, set by cvo_default_params(). */
typedef struct cvo_params {
    float ell_init;      /* 0.15   cvo.cpp:35  (per object, never reset)          */
    float sigma;         /* 0.1    cvo.cpp:36  (s2 = sigma*sigma in float)        */
    float sp_thres;      /* 8e-3   cvo.cpp:37                                      */
    float c;             /* 7.0    cvo.cpp:38                                      */
    float d;             /* 7.0    cvo.cpp:39                                      */
    float c_ell;         /* 200    cvo.cpp:41                                      */
    float c_sigma;       /* 1      cvo.cpp:42                                      */
    int32_t max_iter;    /* 2000   cvo.cpp:48                                      */
    float min_step;      /* 0.2    cvo.cpp:49                                      */
    float max_step;      /* 0.8    cvo.cpp:333                                     */
    float eps;           /* 5e-5   cvo.cpp:50,782                                  */
    float eps_2;         /* 1e-5   cvo.cpp:51,804                                  */
    float ell_after_k2;  /* 0.10   cvo.cpp:810  applied at END of iteration k>2   */
    float ell_after_k9;  /* 0.06   cvo.cpp:811                                     */
    float ell_after_k19; /* 0.03   cvo.cpp:812                                     */
    int32_t num_want;    /* 3000   pcd_generator.cpp:22                            */
    int32_t feature_type;/* 1      cvo.cpp:355,366 (0 = HSV+grad normalised)      */
    int32_t gray_mode;   /* 0 = OpenCV>=4 15-bit RGB2GRAY, 1 = OpenCV 3.x 14-bit  */
    int32_t exp_mode;    /* 0 = k, ck as the reference: exp in double rounded to float
                            (cvo.cpp:172-173; bit-faithful trajectory), 1 = MUFU ex2 (fast) */
} cvo_params;

/* What align() leaves behind (cvo.cpp:763-821) */
typedef struct cvo_align_result {
    float transform[16]; /* 4x4 row-major = [R' | -R'T] (cvo.cpp:106-110): moving -> fixed */
    float R[9];          /* internal state R (row-major)                                   */
    float T[3];          /* internal state T                                               */
    float ell;           /* ell after the loop (persists into the next align)              */
    int32_t iterations;  /* loop iterations executed (k+1 on break, max_iter otherwise)    */
    int32_t iter;        /* the reference's `iter` member: k at break; -1 if no break      */
    int32_t A_nonzero;   /* nnz of A at the last compute_flow (cvo.cpp:197-229)            */
    int32_t status;      /* CVO_OK or CVO_ERR_PAIR_OVERFLOW                                 */
    int32_t num_fixed;   /* points of the fixed / moving cloud this alignment ran on: what set_pcd  */
    int32_t num_moving;  /* caches in num_fixed / num_moving (cvo.cpp:370-371), without a sync     */
    float last_iter_transform[16]; /* `transform` as update_tf() left it at the top of the LAST executed
                          * iteration: what cvo.cpp:815-816 store in prev_transform and multiply into
                          * accum_transform before the final update_tf()                   */
} cvo_align_result;

/* One iteration's observable scalars; used by the parity tests. */
typedef struct cvo_iter_record {
    float ell;
    float omega[3], v[3];
    double B, C, D, E;
    float step;
    int32_t nnz;
} cvo_iter_record;

typedef struct cvo_handle cvo_handle;
typedef struct cvo_batch cvo_batch;

void cvo_default_params(cvo_params *p);
const char *cvo_last_error(void);
/* table of `rand() & 0xFF` after srand(3141592) (PixelSelector2.cpp:36-38), n entries */
int cvo_random_pattern(uint8_t *out, int n);

/* ---- one cvo::cvo object == one handle == one CUDA stream ----------------------- */

/* replaces cvo::cvo(const string& calib_file), cvo.cpp:18-71 */
int cvo_create(const cvo_calib *calib, const cvo_params *params, int device, cvo_handle **out);
int cvo_destroy(cvo_handle *h);

/* replaces pcd_generator::load_image + create_pointcloud as called by cvo::set_pcd
 * (cvo.cpp:345-367; pcd_generator.cpp:618-656): BGR8 + depth u16 -> cloud in `slot`.
 * Strides in bytes.  Images are copied; no host pointer is retained. */
int cvo_set_frame(cvo_handle *h, int slot, const uint8_t *bgr, size_t bgr_stride,
                  const uint16_t *depth, size_t depth_stride, int width, int height);
/* same, inputs already in device memory (tight rows: 3*width, 2*width bytes) */
int cvo_set_frame_device(cvo_handle *h, int slot, const uint8_t *bgr_dev,
                         const uint16_t *depth_dev, int width, int height);
/* upload path for a host point_cloud (data_type.h:67-79): positions n x 3,
 * features n x 5 row-major.  Used by function_inner_product/se3_Hessian wrappers. */
int cvo_set_cloud(cvo_handle *h, int slot, int n, const float *positions, const float *features);
/* replaces the unique_ptr moves of update_fixed_pcd / update_previous_pcd /
 * reset_keyframe (cvo.cpp:578-604): dst <- src, src becomes empty. */
int cvo_slot_move(cvo_handle *h, int dst, int src);
int cvo_slot_size(cvo_handle *h, int slot, int *n);
/* copies the device cloud of (src, src_slot) into (dst, dst_slot): lets a caller that feeds one image to
 * two cvo objects (src/local_tracker.cpp:356,415) run the point selection once */
int cvo_copy_cloud(cvo_handle *dst, int dst_slot, cvo_handle *src, int src_slot);

/* state that persists between align() calls (cvo.hpp:103,122-123) */
int cvo_set_RT(cvo_handle *h, const float R[9], const float T[3]);
int cvo_get_RT(cvo_handle *h, float R[9], float T[3]);
int cvo_set_ell(cvo_handle *h, float ell);
int cvo_get_ell(cvo_handle *h, float *ell);

/* replaces cvo::align (cvo.cpp:763-821): FIXED x MOVING from the current R,T,ell.
 * `trace` (optional, capacity trace_cap) receives one record per iteration. */
int cvo_align(cvo_handle *h, cvo_align_result *out, cvo_iter_record *trace, int trace_cap);
/* one compute_flow + compute_step_size (cvo.cpp:187-334) at an injected state;
 * does not touch the handle's R,T,ell.  Test hook. */
int cvo_iteration_at(cvo_handle *h, const float R[9], const float T[3], float ell,
                     cvo_iter_record *out);

/* in-cutoff pattern (i = fixed index, j = moving index, a_ij) of the last cvo_iteration_at /
 * cvo_align iteration, unordered; n receives the total count.  Test hook. */
int cvo_last_pattern(cvo_handle *h, int32_t *ij, float *a, int cap, int *n);

/* replaces cvo::function_inner_product (cvo.cpp:388-459) for <Ta*slot_a, slot_b> at the
 * handle's current ell.  Ta = 3x4 row-major [R|t] applied to slot_a, or NULL. */
int cvo_inner_product(cvo_handle *h, int slot_a, const float *Ta, int slot_b, float *value,
                      int *num);
/* replaces cvo::se3_Hessian (cvo.cpp:620-759), including the eigenvalue shift. */
int cvo_hessian(cvo_handle *h, int slot_a, const float *Ta, int slot_b, double H[36],
                int *inliers);

/* replaces the body of cvo::compute_innerproduct (cvo.cpp:475-503) in one launch:
 * values/nums = {inn_pre, inn_post, inn_fixed_pcd, inn_moving_pcd}, H = post_hessian; tran is the
 * 4x4 row-major transform applied to the moving cloud for inn_post and the Hessian. */
int cvo_compute_innerproduct(cvo_handle *h, const float tran[16], float values[4], int nums[4],
                             double H[36], int *inliers);

/* Loop-closure verification record: everything cvo::compute_innerproduct_lc (cvo.cpp:505-561)
 * hands back for one candidate, plus the accept rule its only caller applies
 * (src/keyframe_graph.cpp:711-712). */
typedef struct cvo_lc_result {
    /* {inn_prior, inn_lc_prior, inn_lc_pre, inn_lc_post, inn_fixed_pcd, inn_moving_pcd} */
    float value[6];
    int32_t num[6];
    double post_hessian[36]; /* se3_Hessian(lc_tran * moving, fixed), eigenvalue-shifted   */
    int32_t inliers_svd;     /* pairs of that Hessian                       (cvo.cpp:555) */
    int32_t inliers_pnpransac; /* pairs under lc_prior_tran_2               (cvo.cpp:558) */
    float cos_angle;         /* inn_lc_post / (sqrt(inn_fixed) sqrt(inn_moving))           */
    int32_t accept;          /* inn_lc_post > {inn_lc_pre, inn_lc_prior, inn_prior} and cos_angle >= 0.1 */
} cvo_lc_result;

/* replaces the body of cvo::compute_innerproduct_lc (cvo.cpp:505-561) in one launch (the
 * reference runs 6 KD-tree inner products and 2 Hessians one after the other).  All four
 * transforms are 4x4 row-major and are applied to the moving cloud. */
int cvo_compute_innerproduct_lc(cvo_handle *h, const float prior_tran[16], const float lc_prior_tran[16],
                                const float lc_prior_tran_2[16], const float lc_tran[16],
                                cvo_lc_result *out);

/* replaces get_{fixed,moving}_frame_selected_points (cvo.hpp:275-276): xy pairs */
int cvo_get_selected_points(cvo_handle *h, int slot, float *xy, int cap, int *n);
/* the same selected pixels without the round trip through host vectors (the
 * ORB side filters its keypoints against them, src/ORBextractor.cpp:1114-1145): device pointer to
 * n (x, y) float pairs in raster order, valid until the slot is overwritten or moved; work queued on
 * the handle's stream has completed when the call returns. */
int cvo_get_selected_points_device(cvo_handle *h, int slot, const float **xy_dev, int *n);
/* the 8-bit gray image of the frame most recently set on `slot` (RGB2GRAY as pcd_generator::load_image computes it,
 * pcd_generator.cpp:624), width x height bytes, tightly packed, on the device: the ORB side (include/keyframe.h:34-55)
 * converts the same image again on the host.  Valid until the next cvo_set_frame* on this handle; CVO_ERR_NOT_INIT
 * when another slot was set last. */
int cvo_get_gray_device(cvo_handle *h, int slot, const uint8_t **gray_dev, int *width, int *height);
/* positions n x 3, features n x 5 row-major (tests, and host point_cloud mirrors) */
int cvo_get_cloud(cvo_handle *h, int slot, float *positions, float *features, int cap, int *n);
/* selector internals for the bit-exactness tests: status map (w*h bytes, 0/1/2/4 after
 * sub-sampling) and {n2,n3,n4 of the last select pass, pot of that pass, passes run} */
int cvo_get_selection_debug(cvo_handle *h, int slot, uint8_t *map, int32_t info[5]);

/* ---- batches of independent frame pairs (keyframe_graph.cpp:622-731) -------------- */

typedef struct cvo_pair_desc {
    int32_t fixed_frame;  /* index into the batch's frame store                        */
    int32_t moving_frame;
    float R[9];           /* initial state, e.g. from reset_initial (cvo.cpp:611-618)  */
    float T[3];
    float ell;            /* initial ell (ell_init for a fresh object)                 */
} cvo_pair_desc;

int cvo_batch_create(const cvo_calib *calib, const cvo_params *params, int device,
                     int max_frames, int max_pairs, int width, int height, cvo_batch **out);
int cvo_batch_destroy(cvo_batch *b);
/* point selection + features for n frames starting at frame index `first`;
 * images are n tightly packed BGR8 / depth u16 planes.
 * The host variant only ENQUEUES (H2D copies on a copy stream, selection on a selection stream) and returns:
 * the images must stay valid until a call that consumes these frames (align / inner_product / verify_lc /
 * frame_size) has returned.  It is not ordered behind work on OTHER frames, so a caller that creates the batch
 * with room for two sets of frames can upload the frames of step k+1 while step k is being aligned; a call
 * that overwrites frames still in use waits for their consumers. */
int cvo_batch_set_frames(cvo_batch *b, int first, int n, const uint8_t *bgr,
                         const uint16_t *depth);
int cvo_batch_set_frames_device(cvo_batch *b, int first, int n, const uint8_t *bgr_dev,
                                const uint16_t *depth_dev);
int cvo_batch_frame_size(cvo_batch *b, int frame, int *n);
/* set_pcd'd frames x pairs -> one cvo_align_result per pair (fresh cvo object each). */
int cvo_batch_align(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs,
                    cvo_align_result *results);
/* per-pair <T*moving, fixed> at each pair's final ell (compute_innerproduct_lc, cvo.cpp:545) */
int cvo_batch_inner_product(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs,
                            const cvo_align_result *results, float *values, int *nums);
/* compute_innerproduct_lc for every pair of a batch in one launch (the candidate loop of
 * detectLoopClousure_top10, src/keyframe_graph.cpp:693-731, after cvo_batch_align):
 * lc_tran = results[i].transform; prior / lc_prior / lc_prior_2 are n_pairs x 16 floats
 * (4x4 row-major each).  The self inner products <fixed,fixed>, <moving,moving> are evaluated
 * once per distinct (frame, ell). */
int cvo_batch_verify_lc(cvo_batch *b, int n_pairs, const cvo_pair_desc *pairs,
                        const cvo_align_result *results, const float *prior_tran,
                        const float *lc_prior_tran, const float *lc_prior_tran_2,
                        cvo_lc_result *out);
/* counters since creation: {kernel launches, in-cutoff kernel evaluations (d2 < d2_thres),
 * align iterations, stored non-zeros summed over iterations} */
int cvo_batch_stats(cvo_batch *b, int64_t stats[4]);
int cvo_handle_stats(cvo_handle *h, int64_t stats[4]);
/* cumulative SM cycles (thread 0 of every CTA) per phase of the align kernel:
 * {grid build, P0 bookkeeping, P1a neighbour-list construction (search or filter, colour kernel, tiling),
 * P1b kernel values + flow, P2 step coefficients, P3 scalar update}, then two counters: neighbour lists derived
 * by filtering the current one, and neighbour lists built by a grid search */
int cvo_handle_phase_cycles(cvo_handle *h, int64_t cycles[8]);
int cvo_batch_phase_cycles(cvo_batch *b, int64_t cycles[8]);
/* device-time of the last cvo_batch_align's kernel in ms (CUDA events on its stream) */
int cvo_batch_last_align_ms(cvo_batch *b, float *ms);
/* CUDA events on the batch's own stream (the stream every kernel of the batch is launched on):
 * mark(0) before and mark(1) after a region, elapsed_ms waits for mark 1 and returns the time. */
int cvo_batch_mark(cvo_batch *b, int which);
int cvo_batch_elapsed_ms(cvo_batch *b, float *ms);

/* ---- image ingest (src/run_SLAM.cpp:134-143: cv::imread of the colour and the depth PNG) ------------------------
 * cvo_png_decode_bgr8 produces the buffer cv::imread(path) returns for an 8/16-bit gray / RGB / RGBA PNG (8-bit,
 * 3 channels, B G R); cvo_png_decode_depth16 the buffer cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH) returns for a 16-bit
 * gray PNG (u16, host byte order).  out may be NULL to query the size only.  Non-interlaced, non-palette files. */
int cvo_png_info(const uint8_t *png, size_t n, int *width, int *height, int *channels, int *bit_depth);
int cvo_png_decode_bgr8(const uint8_t *png, size_t n, uint8_t *out, size_t out_bytes, int *width, int *height);
int cvo_png_decode_depth16(const uint8_t *png, size_t n, uint16_t *out, size_t out_elems, int *width, int *height);
/* imread x 2 + cvo::set_pcd in one call: both PNGs decoded in parallel into pinned staging, then cvo_set_frame */
int cvo_set_frame_png(cvo_handle *h, int slot, const uint8_t *rgb_png, size_t rgb_bytes,
                      const uint8_t *depth_png, size_t depth_bytes);
/* A decode prefetcher: n_threads workers decode submitted frames, in order, into a ring of n_slots pinned frames while
 * the caller aligns the previous ones.  submit copies the PNG bytes and returns (blocks only when the ring is full);
 * wait returns the oldest frame's pinned BGR8 / u16 buffers (ready for cvo_set_frame: the H2D copy from pinned memory
 * is asynchronous); release hands the slot back. */
typedef struct cvo_ingest cvo_ingest;
int cvo_ingest_create(int width, int height, int n_slots, int n_threads, cvo_ingest **out);
int cvo_ingest_destroy(cvo_ingest *g);
int cvo_ingest_submit(cvo_ingest *g, const uint8_t *rgb_png, size_t rgb_bytes, const uint8_t *depth_png,
                      size_t depth_bytes);
int cvo_ingest_wait(cvo_ingest *g, const uint8_t **bgr, const uint16_t **depth, int *width, int *height);
int cvo_ingest_release(cvo_ingest *g);

/* ---- one list of frames and pairs over several GPUs of this process ----------------------------
 * SURVEY section 8b lists a batch entry with an n_devices argument: the candidate loop of
 * src/keyframe_graph.cpp:622-731 with the pairs split into n_devices contiguous blocks (section 8e: pairs
 * shard, no collective).  Each device uploads and selects only the frames its block touches; one host
 * thread per device; results land in the caller's array in pair order. */
typedef struct cvo_multi cvo_multi;
/* devices: n_devices CUDA ordinals, or NULL for 0..n_devices-1 */
int cvo_multi_create(const cvo_calib *calib, const cvo_params *params, int n_devices, const int *devices,
                     int max_frames, int max_pairs, int width, int height, cvo_multi **out);
int cvo_multi_destroy(cvo_multi *m);
/* host images of n_frames frames (tightly packed BGR8 / depth u16 planes) and n_pairs pairs indexing
 * them -> results[n_pairs]; values / nums (both or neither may be NULL) receive the post-alignment
 * inner product <T*moving, fixed> of every pair (cvo.cpp:545). */
int cvo_multi_align(cvo_multi *m, int n_frames, const uint8_t *bgr, const uint16_t *depth, int n_pairs,
                    const cvo_pair_desc *pairs, cvo_align_result *results, float *values, int *nums);
/* per device, for the last cvo_multi_align: frames selected, pairs aligned, device time of the align kernel */
int cvo_multi_last_shares(cvo_multi *m, int *frames, int *pairs, float *align_ms);

#ifdef __cplusplus
}
#endif
#endif /* CVO_B200_H */
