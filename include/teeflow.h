/*
 * teeflow.h -- C ABI of libteeflow.so, the B200-native (sm_100a) TV-L1 dense optical-flow engine.
 *
 * Drop-in boundary for ONE path of nquach/TEE_optical_flow: the per-frame-pair TV-L1 call that the reference
 * makes into OpenCV (file:line relative to /root/reference):
 *
 *   reference interface                                             replaced by
 *   -------------------------------------------------------------   ---------------------------------------
 *   cv2.optflow.createOptFlow_DualTVL1()                            teeflow_create
 *        optical_flow/calculate_optical_flow.py:577
 *   OF_model.setLambda(config.lambda_value)      (:578)             teeflow_set_param(h, "lambda", v)
 *   OF_model.calc(saliency_1, saliency_2, None)  (:642)             teeflow_calc_pair_host
 *   the serial pair loop of process_video        (:584-600)         teeflow_calc_clip / teeflow_calc_clip_host
 *        (flow_list.append(flow_list[-1]); np.stack(..) * conversion_factor; .astype(float16) at :403)
 *   cv2.cuda.OpticalFlowDual_TVL1 upload/calc/download (:633-639)   teeflow_calc_clip (device pointers)
 *
 * Plain pointers and sizes only; no C++ or torch types.  All entry points return 0 on success or a negative
 * teeflow_status; teeflow_last_error() gives the text.  A handle is bound to one CUDA device, owns its
 * workspace, and is not thread-safe (use one handle per thread / per GPU).  There is no CPU fallback: without
 * a CUDA device teeflow_create fails with TEEFLOW_ERR_CUDA.
 */
#ifndef TEEFLOW_H_
#define TEEFLOW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TEEFLOW_API __attribute__((visibility("default")))
#else
#define TEEFLOW_API
#endif

#define TEEFLOW_ABI_VERSION 1
#define TEEFLOW_MAX_LEVELS 16

typedef enum {
    TEEFLOW_OK = 0,
    TEEFLOW_ERR_BAD_ARG = -1,   /* NULL pointer, unknown key, value out of range */
    TEEFLOW_ERR_BAD_SHAPE = -2, /* H/W/n_frames not supported or larger than the handle's capacity */
    TEEFLOW_ERR_CUDA = -3,      /* a CUDA runtime call failed (text in teeflow_last_error) */
    TEEFLOW_ERR_NCCL = -4,      /* reserved for the multi-GPU gather */
    TEEFLOW_ERR_STATE = -5      /* internal scheduler did not converge (should never happen) */
} teeflow_status;

typedef enum { TEEFLOW_U8 = 0, TEEFLOW_F32 = 1 } teeflow_dtype;

/* OpenCV DualTVL1OpticalFlow parameters (defaults of createOptFlow_DualTVL1).
 * gamma and useInitialFlow are not exposed: the reference never sets them (gamma = 0, no initial flow). */
typedef struct {
    double tau;              /* 0.25 */
    double lambda;           /* 0.15  == OpticalFlowCalculationConfig.lambda_value, config.py:177 */
    double theta;            /* 0.3 */
    double epsilon;          /* 0.01 */
    double scale_step;       /* 0.8 */
    int32_t nscales;         /* 5 */
    int32_t warps;           /* 5 */
    int32_t inner_iterations; /* 30 */
    int32_t outer_iterations; /* 10 */
    int32_t median_filtering; /* 5 (<=1 off, 3, 5) */
    int32_t max_slots;       /* frame pairs solved concurrently (0 = default 64) */
} teeflow_params;

typedef struct teeflow_engine* teeflow_handle;

TEEFLOW_API void teeflow_default_params(teeflow_params* p);
TEEFLOW_API int teeflow_abi_version(void);

TEEFLOW_API int teeflow_create(const teeflow_params* p, int device, teeflow_handle* out);
TEEFLOW_API int teeflow_destroy(teeflow_handle h);
/* keys: tau lambda theta epsilon scale_step nscales warps inner_iterations outer_iterations median_filtering */
TEEFLOW_API int teeflow_set_param(teeflow_handle h, const char* key, double value);
TEEFLOW_API int teeflow_get_param(teeflow_handle h, const char* key, double* value);
/* h may be NULL: returns the text of the last error raised without a handle (e.g. by teeflow_create) */
TEEFLOW_API const char* teeflow_last_error(teeflow_handle h);

/*
 * Flow of every consecutive frame pair (i -> i+1) of one clip, all pairs in flight together.
 *   frames_dev      device pointer, n_frames x H x W, dtype u8 (x1) or f32 in [0,1] (x255, like OpenCV)
 *   frame_stride    elements between consecutive frames (>= H*W)
 *   flow_f32_dev    device, [n_out, H, W, 2] float32, may be NULL;   n_out = n_frames-1 (+1 if duplicate_last)
 *   flow_f16_dev    device, [n_out, H, W, 2] IEEE half (RNE),  may be NULL
 *   out_scale       conversion_factor = pixel_spacing * frame_rate (calculate_optical_flow.py:538-541,600)
 *   duplicate_last  append a copy of the last flow so that n_out == n_frames (:599)
 *   stream          cudaStream_t (as void*) the caller's inputs are ready on; the call returns after the
 *                   results are complete on that stream (it synchronises the stream).
 */
TEEFLOW_API int teeflow_calc_clip(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                      int64_t frame_stride, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                      int duplicate_last, void* stream);

/* Asynchronous form of teeflow_calc_clip ("asynchronous on the caller's stream"): enqueues the image
 * pyramid and the ONE dataflow launch that solves every pair on `stream` and returns; work the caller enqueues on the
 * same stream afterwards runs behind it.  teeflow_finish() waits for the run, checks the scheduler's verdict and
 * completes the statistics; teeflow_get_counters / teeflow_get_stats / teeflow_destroy call it implicitly, and a new
 * calc on the handle is refused until then.  Not available in stepped mode. */
TEEFLOW_API int teeflow_calc_clip_async(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                            int64_t frame_stride, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                            int duplicate_last, void* stream);
TEEFLOW_API int teeflow_finish(teeflow_handle h);

/* Generic form: n_pairs arbitrary (frame a -> frame b) pairs over a frame array (used for batches of clips and
 * for sharding a clip by pair range).  out_index[p] / dup_index[p] (host arrays) give the output slot of pair p
 * and an optional second slot to copy it to (-1 = none). */
TEEFLOW_API int teeflow_calc_pairs(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                       int64_t frame_stride, const int32_t* pair_a, const int32_t* pair_b,
                       const int32_t* out_index, const int32_t* dup_index, int n_pairs, float* flow_f32_dev,
                       void* flow_f16_dev, float out_scale, void* stream);

/* Same as teeflow_calc_clip with HOST buffers: copies the frames in, the flow out (the end-to-end path). */
TEEFLOW_API int teeflow_calc_clip_host(teeflow_handle h, const void* frames_host, int dtype, int n_frames, int H, int W,
                           float* flow_f32_host, void* flow_f16_host, float out_scale, int duplicate_last);

/* OF_model.calc(I0, I1, None): two H x W host images -> H x W x 2 float32 host flow. */
TEEFLOW_API int teeflow_calc_pair_host(teeflow_handle h, const void* I0_host, const void* I1_host, int dtype, int H, int W,
                           float* flow_host);

/* Timing / launch record of the last calc (device times from CUDA events on the caller's stream). */
typedef struct {
    int32_t n_pairs;          /* frame pairs solved */
    int32_t n_levels;         /* pyramid levels used */
    int32_t n_slots;          /* pairs in flight */
    int32_t grid_ctas;        /* persistent CTAs per solver launch */
    int64_t solver_launches;  /* tvl1_step_kernel launches (scheduler super-steps, incl. over-issued ones) */
    int64_t kernel_launches;  /* all kernels launched by the call (pyramid + solver) */
    float device_ms;          /* whole call on the device */
    float pyramid_ms;         /* pyramid + gradient pack kernels */
    float solver_ms;          /* first to last solver launch */
    float reserved;
    int32_t double_steps;            /* two-iteration passes of the inner loop that were applied ... */
    int32_t double_steps_discarded;  /* ... and discarded (the first iteration already met the exit test) */
} teeflow_stats;

/* Work actually executed by the last calc: counters[p][level][3] = inner iterations, median passes, warps
 * (int32, level 0 = finest, TEEFLOW_MAX_LEVELS levels per pair) -- needed for the roofline accounting because
 * the inner loop exits early.  n_pairs_cap = capacity of `counters` in pairs (counters may be NULL).
 * Returns the number of pairs of the last calc. */
TEEFLOW_API int teeflow_get_counters(teeflow_handle h, int32_t* counters, int n_pairs_cap);
TEEFLOW_API int teeflow_get_stats(teeflow_handle h, teeflow_stats* out);
/* Diagnostics of the dataflow scheduler (libraries built with -DTEEFLOW_FLOW_STATS=1 only; returns 0 and zeros
 * otherwise): out32[i] = warp cycles summed over all warps of the last calc, i = solver phase (1 level-init, 2 warp,
 * 3 median, 4 inner, 5 final, 6 WASE, 7 two-iteration pass), 9 = waiting for a task to be published, 10 = scheduling
 * (ticket, descriptor probe, fences, arrival, hand-over); out32[16 + i] = number of such intervals. */
TEEFLOW_API int teeflow_get_flow_stats(teeflow_handle h, uint64_t* out32);
/* Diagnostics for the roofline accounting of the individual phases: time the first n (<= 64) solver launches of
 * every following calc with CUDA events on the launching stream (n = 0 switches it off), and fetch the durations
 * (ms) of the last calc.  With one slot group (environment TEEFLOW_GROUPS=1 at teeflow_create) and as many slots as
 * pairs, all pairs step in lockstep at first, so launch 0 is a pure level-init, 1 a pure warp, 2 a pure median,
 * 3.. pure inner iterations over all pairs.  teeflow_get_launch_times returns the number of launches timed. */
TEEFLOW_API int teeflow_time_launches(teeflow_handle h, int n);
TEEFLOW_API int teeflow_get_launch_times(teeflow_handle h, float* ms, int cap);

/* Pyramid geometry the handle would use for an H x W image: level sizes (finest first). Returns the level count. */
TEEFLOW_API int teeflow_level_sizes(teeflow_handle h, int H, int W, int32_t* Hs, int32_t* Ws);

/* ---- frame prep: the `no_saliency=True` input stage, img2uint8(rgb2gray(frame)) per frame
 * (calculate_optical_flow.py:588; optical_flow_utils.py:30-31).  rgb_dev: (n_frames,H,W,3) uint8, gray_dev:
 * (n_frames,H,W) uint8, both device pointers. */
TEEFLOW_API int teeflow_prepare_frames(teeflow_handle h, const uint8_t* rgb_dev, int n_frames, int H, int W,
                                       uint8_t* gray_dev, void* stream);

/* ---- saliency input stage: the `no_saliency=False` branch, cv2.saliency.StaticSaliencyFineGrained_create()
 * .computeSaliency(frame) per frame (calculate_optical_flow.py:560, :586).  rgb_dev: (n_frames,H,W,3) uint8;
 * saliency_dev: (n_frames,H,W) float32 in [0,1] (what computeSaliency returns and the solver takes as its image),
 * intensity_dev: the uint8 conspicuity map before the 1/255 scaling; either may be NULL.  Device pointers,
 * asynchronous on `stream`.  Parity status: csrc/saliency_kernels.cuh. */
TEEFLOW_API int teeflow_saliency_fine_grained(teeflow_handle h, const uint8_t* rgb_dev, int n_frames, int H, int W,
                                              float* saliency_dev, uint8_t* intensity_dev, void* stream);

/* ---- mask post-processing and AV centroid (connected components on the GPU)
 * teeflow_clean_masks: the per-label pipeline of clean_mask (calculate_optical_flow.py:113-182) on a class map
 * (n_frames,H,W) uint8 (SAM argmax): (class == class_id) -> moving_avg_mask(window, threshold) over frames (:91-111)
 * -> scipy.ndimage.binary_fill_holes -> skimage remove_small_objects(min_size) per frame.  mask_dev: (n_frames,H,W)
 * bool (0/1 bytes), device pointers.
 * teeflow_av_centroids: calc_AV_centroid's per-frame step (analysis.py:58-63): 8-connected components of
 * mask[..., 0] (mask_dev is (nframes,H,W,channels) bool), centroid (row, col) of the largest component (first in
 * label order among ties); NaN when the frame is empty (the caller applies the copy-previous / image-centre
 * fallback and the Savitzky-Golay filter). */
TEEFLOW_API int teeflow_clean_masks(teeflow_handle h, const uint8_t* classmap_dev, int n_frames, int H, int W,
                                    int class_id, int window, double threshold, int min_size, uint8_t* mask_dev,
                                    void* stream);
TEEFLOW_API int teeflow_av_centroids(teeflow_handle h, const uint8_t* mask_dev, int channels, int nframes, int H,
                                     int W, double* centroids_host, int32_t* n_components_host, void* stream);

/* ---- WASE background compensation (calculate_optical_flow.py:649-660).
 * teeflow_wase_weights: w[y,x,c] = sum_n bkgd[n,y,x,c] from the (n_frames,H,W,2) bool mask `mask_dict['bkgd']`
 * (device pointers).  teeflow_set_wase(h, w_dev, H, W): every following calc_* subtracts, per pair, the scalar
 *   background = mean of the non-zero entries of flow * bkgd  ==  sum(w f [f != 0]) / sum(w [f != 0])
 * before the out_scale multiplication; w_dev stays owned by the caller; NULL switches it off (bkgd_comp='none').
 * teeflow_get_backgrounds returns the scalars of the last calc (one per pair). */
TEEFLOW_API int teeflow_wase_weights(teeflow_handle h, const uint8_t* bkgd_dev, int n_frames, int H, int W,
                                     float* w_dev, void* stream);
TEEFLOW_API int teeflow_set_wase(teeflow_handle h, const float* w_dev, int H, int W);
TEEFLOW_API int teeflow_get_backgrounds(teeflow_handle h, float* bg_host, int n_pairs_cap);

/* ---- masked radial / longitudinal decomposition with per-frame reductions
 * (optical_flow/analysis.py:89-327: radial_vecgrid, calculate_comp_magnitude, calc_bidirectional_hist,
 *  calculate_3dhist; optical_flow/cardiac_cycle_detection.py:100-116: per-frame angle mode).
 * Inputs: the stored flow (N,H,W,2) fp16 (device), the (N,H,W,2) bool mask of the analysed label (device) --
 * masked_arr = flow.astype(f32) * mask -- and the per-frame AV centroids (row, col) float64 (host).  The first
 * `nframes` frames are analysed.  Output arrays (host, caller-allocated, any may be NULL) hold one value per
 * frame; a frame without non-zero entries yields NaN (the caller applies the reference's carry-forward rule). */
typedef struct {
    float* mag_hi;      /* [nframes] np.percentile(|v| != 0, perc_hi), float32 arithmetic like numpy */
    float* ang_mode;    /* [nframes] scipy.stats.mode(np.round(angle, 2) != 0) */
    double* rad_hi;     /* [nframes] np.percentile(radial != 0, perc_hi) */
    double* rad_lo;     /* [nframes]                              perc_lo */
    double* long_hi;    /* [nframes] longitudinal */
    double* long_lo;
    int64_t* counts;    /* [nframes][4] non-zero entries: magnitude, angle, radial, longitudinal */
    float mag_min, mag_max, ang_min, ang_max;          /* np.min / np.max over the analysed arrays */
    double rad_min, rad_max, long_min, long_max;
} teeflow_analysis;

TEEFLOW_API int teeflow_analyze_clip(teeflow_handle h, const void* flow_f16_dev, const uint8_t* mask_dev,
                                     const double* centroids_host, int nframes, int H, int W, double perc_lo,
                                     double perc_hi, teeflow_analysis* out, void* stream);
/* np.histogram(non-zero entries, bins=nbins, range=(edges[0], edges[nbins])) per frame for one quantity of the
 * last teeflow_analyze_clip: 0 magnitude, 1 angle (float32 edges), 2 radial, 3 longitudinal (float64 edges);
 * edges = np.linspace(first, last, nbins + 1) computed by the caller.  freq_host: [nframes][nbins] int64. */
TEEFLOW_API int teeflow_analysis_histogram(teeflow_handle h, int quantity, const void* edges_host, int nbins,
                                           int64_t* freq_host, void* stream);

/* Diagnostics: compares the engine's shared-reciprocal exact division with IEEE division (__fdiv_rn) on n
 * pseudo-random operand pairs (mode 0: dual-update operand ranges, 1: thresholding ranges, 2: all exponents) and
 * returns the number of bit mismatches (must be 0). */
TEEFLOW_API int teeflow_selftest_division(teeflow_handle h, int mode, int64_t n, uint64_t seed, int64_t* mismatches);
/* Diagnostics: compares the inner iteration's packed float-float hypot (fast form) with
 * (float)sqrt((double)a*a + (double)b*b) on n pseudo-random operand pairs (mode 0: flow-gradient magnitudes
 * 2^-30..2^4, both operands of similar size; 1: very different magnitudes, zeros and exact squares mixed in;
 * 2: all exponents incl. denormals / huge).  mismatches = pairs the fast form ACCEPTED with a wrong result (must
 * be 0); rejected = pairs it handed to the exact form. */
TEEFLOW_API int teeflow_selftest_hypot(teeflow_handle h, int mode, int64_t n, uint64_t seed, int64_t* mismatches,
                                       int64_t* rejected);

#ifdef __cplusplus
}
#endif
#endif /* TEEFLOW_H_ */
