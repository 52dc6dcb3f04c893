// teeflow.cu -- host side of libteeflow.so: the C ABI declared in include/teeflow.h.
//
// Replaces, for the reference's TV-L1 path only (optical_flow/calculate_optical_flow.py:564-600,627-660):
// the cv2.optflow DualTVL1 object and the serial per-pair loop that drives it.  One handle == one GPU.
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <thread>
#include <string>
#include <vector>

#include "../../include/teeflow.h"
#include "tvl1_kernels.cuh"
#include "finalize_kernels.cuh"
#include "saliency_kernels.cuh"

using namespace teeflow;

static_assert(TEEFLOW_MAX_LEVELS == kMaxLevels, "ABI level count");

static thread_local std::string g_last_error;
static const int kMaxGroups = 4;        // slot groups (streams) per handle
static const int kMinGroupSlots = 4;    // do not split fewer than 2 * this many slots
static const int kCtlInts = 2 + 2 * kMaxGroups + 2;   // next_pair, pairs_done, per-group item counters [parity], two-iteration stats
#ifndef TEEFLOW_PITCH0
#define TEEFLOW_PITCH0 1024
#endif
static const int kPitches[] = {TEEFLOW_PITCH0, 2048, 4096};  // instantiated plane pitches (float2 elements): W <= pitch
static const size_t kPlaneAlign = (size_t)kPlanes * 4096 * sizeof(float2);   // Lay<4096>::ROWB, a multiple of the others

typedef void (*step_kernel_t)(const EngineParams, const int);
static step_kernel_t step_kernel_for(int pitch) {
    switch (pitch) {
        case TEEFLOW_PITCH0: return tvl1_step_kernel<TEEFLOW_PITCH0>;
        case 2048: return tvl1_step_kernel<2048>;
        case 4096: return tvl1_step_kernel<4096>;
        default: return nullptr;
    }
}

typedef void (*flow_kernel_t)(const EngineParams);
static flow_kernel_t flow_kernel_for(int pitch) {
    switch (pitch) {
        case TEEFLOW_PITCH0: return tvl1_flow_kernel<TEEFLOW_PITCH0>;
        case 2048: return tvl1_flow_kernel<2048>;
        case 4096: return tvl1_flow_kernel<4096>;
        default: return nullptr;
    }
}

struct teeflow_engine {
    teeflow_params p;
    int device = 0;
    int num_sms = 0;
    int ctas_per_sm[3] = {1, 1, 1};   // per instantiated pitch (stepped kernel)
    int ctas_per_sm_flow[3] = {1, 1, 1};   // ... (dataflow kernel)
    int stepped = 0;                  // 1: one launch per phase step (profiling / A-B); 0: one dataflow launch per run
    int async_mode = 0;               // set by teeflow_calc_clip_async for the duration of its run_pairs call
    int pending = 0;                  // an asynchronous dataflow run is in flight: teeflow_finish() completes it
    int pending_pairs = 0, pending_grid = 0;
    cudaStream_t pending_stream = nullptr;
    CUtensorMap* tmaps = nullptr;     // [kMaxLevels][3] device copies of the tensor maps over the slot planes (TMA staging)
    Task* tasks = nullptr;            // [kTaskRing] task descriptors of the dataflow scheduler
    FlowCtl* flow_ctl = nullptr;
    unsigned long long* flow_stats = nullptr;   // [32] device, diagnostic builds
    unsigned long long flow_stats_host[32] = {0};
    int* h_flow_order = nullptr;      // mapped pinned host memory: completion order written by the running kernel
    int* d_flow_order = nullptr;      // its device address
    size_t cap_flow_order = 0;
    std::string err;
    // workspace (grown on demand)
    size_t cap_frames = 0, cap_pyr_stride = 0;
    size_t cap_plane_elems = 0;   // float2 elements available behind `planes`
    size_t cap_tiles = 0, cap_pairs = 0;
    float* pyrI = nullptr;
    float4* pyrG = nullptr;
    void* planes_raw = nullptr;   // allocation behind `planes`
    float2* planes = nullptr;     // [S][H0][kPlanes][PITCH], aligned to kPlaneAlign (struct Lay)
    Slot* slots = nullptr;     // [2][S]
    unsigned* arrive = nullptr;
    double* partial = nullptr;
    int* ctl = nullptr;        // [0] next_pair, [1] pairs_done
    int* pair_lists = nullptr; // [5][cap_pairs]: pair_a, pair_b, out_index, dup_index, done_order
    int* h_order = nullptr;    // pinned [2][cap_pairs]: snapshots of done_order
    cudaStream_t copy_stream = nullptr;   // early device-to-host copies of finished flows (host-buffer entry points)
    int* counters = nullptr;   // [cap_pairs][kMaxLevels][3]
    float* bg = nullptr;       // [cap_pairs] WASE background scalars
    const float* wase_w = nullptr;  // caller-owned [H][W][2] weight map (nullptr: no background compensation)
    int wase_H = 0, wase_W = 0;
    // analysis scratch (teeflow_analyze_clip)
    size_t an_cap = 0, an_frames_cap = 0; int an_frames = 0, an_H = 0, an_W = 0;
    float *an_mag = nullptr, *an_ang = nullptr;
    double *an_rad = nullptr, *an_long = nullptr, *an_cent = nullptr;
    FrameStats* an_stats = nullptr;
    unsigned* an_anghist = nullptr;
    long long* an_ranks = nullptr;
    unsigned long long* an_keys = nullptr;
    unsigned long long* prep_mm = nullptr; size_t prep_cap = 0;
    // saliency scratch (teeflow_saliency_fine_grained), sized for kSalChunk frames
    size_t sal_cap_px = 0, sal_cap_int = 0;   // capacities in pixels / integral-image entries per frame
    uint8_t *sal_g0 = nullptr, *sal_g1 = nullptr, *sal_on = nullptr, *sal_off = nullptr;
    float *sal_prefix = nullptr, *sal_integ = nullptr;
    uint16_t *sal_son = nullptr, *sal_soff = nullptr;
    SalMax* sal_max = nullptr;
    // connected-component scratch (teeflow_clean_masks / teeflow_av_centroids), sized for kCclChunk frames
    size_t ccl_cap = 0;
    int *ccl_L = nullptr, *ccl_area = nullptr, *ccl_touch = nullptr, *ccl_ncomp = nullptr;
    unsigned long long *ccl_sr = nullptr, *ccl_sc = nullptr, *ccl_best = nullptr;
    void* an_edges = nullptr; unsigned long long* an_freq = nullptr; size_t an_freq_cap = 0;
    int* h_done = nullptr;     // pinned
    void* stage_in = nullptr; size_t stage_in_bytes = 0;
    void* stage_f32 = nullptr; size_t stage_f32_bytes = 0;
    void* stage_f16 = nullptr; size_t stage_f16_bytes = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_tp = nullptr;
    cudaStream_t own_stream = nullptr;
    int groups = 2;                                   // slot groups stepping on separate streams
    float spec_factor = 0.0f;                         // two-iteration passes while error > spec_factor * threshold (0: off)
    cudaStream_t group_stream[kMaxGroups - 1] = {};
    cudaEvent_t ev_group[2][kMaxGroups - 1] = {};
    // per-launch timing diagnostics (teeflow_time_launches)
    static const int kMaxLaunchEvents = 64;
    int n_launch_events = 0, n_launches_timed = 0;
    cudaEvent_t ev_launch[kMaxLaunchEvents + 1] = {};
    float launch_ms[kMaxLaunchEvents] = {};
    // last-call record
    teeflow_stats st{};
};

static int fail(teeflow_engine* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_last_error = buf;
    return code;
}

#define CU_TRY(h, call)                                                                             \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(h, TEEFLOW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

static int validate_params(teeflow_engine* h, const teeflow_params& p) {
    if (!(p.tau > 0) || !(p.lambda > 0) || !(p.theta > 0) || !(p.epsilon >= 0))
        return fail(h, TEEFLOW_ERR_BAD_ARG, "tau, lambda, theta must be > 0 and epsilon >= 0");
    if (!(p.scale_step > 0 && p.scale_step < 1)) return fail(h, TEEFLOW_ERR_BAD_ARG, "scale_step must be in (0,1)");
    if (p.nscales < 1 || p.nscales > kMaxLevels) return fail(h, TEEFLOW_ERR_BAD_ARG, "nscales must be in [1,%d]", kMaxLevels);
    if (p.warps < 1 || p.inner_iterations < 1 || p.outer_iterations < 1)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "warps, inner_iterations, outer_iterations must be >= 1");
    if (p.median_filtering > 1 && p.median_filtering != 3 && p.median_filtering != 5)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "median_filtering must be <=1 (off), 3 or 5");
    if (p.max_slots < 0 || p.max_slots > kMaxSlots) return fail(h, TEEFLOW_ERR_BAD_ARG, "max_slots must be in [0,%d]", kMaxSlots);
    return TEEFLOW_OK;
}

// pyramid geometry: dsize = cvRound(size * scale_step); stop before a level with a side < 16 px (tvl1flow.cpp calc)
static int level_geometry(const teeflow_params& p, int H, int W, int* Hs, int* Ws) {
    Hs[0] = H; Ws[0] = W;
    int L = 1;
    for (int s = 1; s < p.nscales; ++s) {
        const int w = (int)std::lrint(Ws[s - 1] * p.scale_step);
        const int hh = (int)std::lrint(Hs[s - 1] * p.scale_step);
        if (w < 16 || hh < 16) break;
        Hs[s] = hh; Ws[s] = w; L = s + 1;
    }
    return L;
}

static int query_occupancy(teeflow_engine* h) {
    for (int i = 0; i < 3; ++i) {
        int occ = 0;
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel_for(kPitches[i]), kThreads, 0));
        if (occ < 1) return fail(h, TEEFLOW_ERR_CUDA, "tvl1_step_kernel cannot be resident on this device");
        h->ctas_per_sm[i] = occ;
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, flow_kernel_for(kPitches[i]), kThreads, 0));
        if (occ < 1) return fail(h, TEEFLOW_ERR_CUDA, "tvl1_flow_kernel cannot be resident on this device");
        h->ctas_per_sm_flow[i] = occ;
    }
    int coop = 0;
    CU_TRY(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) h->stepped = 1;        // no residency guarantee for the dataflow kernel: one launch per phase step instead
    return TEEFLOW_OK;
}

extern "C" {

void teeflow_default_params(teeflow_params* p) {
    if (!p) return;
    p->tau = 0.25; p->lambda = 0.15; p->theta = 0.3; p->epsilon = 0.01; p->scale_step = 0.8;
    p->nscales = 5; p->warps = 5; p->inner_iterations = 30; p->outer_iterations = 10; p->median_filtering = 5;
    p->max_slots = 0;
}

int teeflow_abi_version(void) { return TEEFLOW_ABI_VERSION; }

const char* teeflow_last_error(teeflow_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

// everything of teeflow_create that can fail after the handle exists; the caller destroys the handle on error
static int create_resources(teeflow_engine* h) {
    CU_TRY(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CU_TRY(h, cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    int rc = query_occupancy(h);
    if (rc) return rc;
    CU_TRY(h, cudaMallocHost(&h->h_done, sizeof(int) * 4));
    CU_TRY(h, cudaMalloc(&h->ctl, sizeof(int) * kCtlInts));
    for (auto& ev : h->ev) CU_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU_TRY(h, cudaEventCreate(&h->ev_t0));
    CU_TRY(h, cudaEventCreate(&h->ev_t1));
    CU_TRY(h, cudaEventCreate(&h->ev_tp));
    CU_TRY(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CU_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int g = 0; g < kMaxGroups - 1; ++g) {
        CU_TRY(h, cudaStreamCreateWithFlags(&h->group_stream[g], cudaStreamNonBlocking));
        for (int w = 0; w < 2; ++w) CU_TRY(h, cudaEventCreateWithFlags(&h->ev_group[w][g], cudaEventDisableTiming));
    }
    return TEEFLOW_OK;
}

int teeflow_create(const teeflow_params* p, int device, teeflow_handle* out) {
    if (!out) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "out handle pointer is NULL");
    *out = nullptr;
    teeflow_params pp;
    if (p) pp = *p; else teeflow_default_params(&pp);
    int rc = validate_params(nullptr, pp);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TEEFLOW_ERR_CUDA, "no CUDA device available (%s); teeflow has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "device %d out of range [0,%d)", device, ndev);
    teeflow_engine* h = new (std::nothrow) teeflow_engine();
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "out of host memory");
    h->p = pp;
    h->device = device;
    rc = create_resources(h);
    if (rc) {                          // streams, events, pinned memory created so far go with the handle
        const std::string msg = h->err;
        teeflow_destroy(h);
        g_last_error = msg;
        return rc;
    }
    if (const char* e = getenv("TEEFLOW_GROUPS")) h->groups = std::max(1, std::min(atoi(e), kMaxGroups));
    if (const char* e = getenv("TEEFLOW_SPEC")) h->spec_factor = std::max(0.0f, (float)atof(e));
    if (const char* e = getenv("TEEFLOW_STEPPED")) h->stepped = atoi(e) != 0 || h->stepped;
    *out = h;
    return TEEFLOW_OK;
}

int teeflow_destroy(teeflow_handle h) {
    if (!h) return TEEFLOW_OK;
    cudaSetDevice(h->device);
    if (h->pending) { cudaStreamSynchronize(h->pending_stream); h->pending = 0; }
    cudaFree(h->pyrI); cudaFree(h->pyrG);
    cudaFree(h->planes_raw); cudaFree(h->slots); cudaFree(h->arrive); cudaFree(h->partial); cudaFree(h->ctl);
    cudaFree(h->pair_lists); cudaFree(h->counters); cudaFree(h->bg); cudaFree(h->tasks); cudaFree(h->flow_ctl); cudaFree(h->flow_stats); cudaFree(h->tmaps);
    if (h->h_flow_order) cudaFreeHost(h->h_flow_order);
    cudaFree(h->an_mag); cudaFree(h->an_ang); cudaFree(h->an_rad); cudaFree(h->an_long); cudaFree(h->an_cent);
    cudaFree(h->an_stats); cudaFree(h->an_anghist); cudaFree(h->an_ranks); cudaFree(h->an_keys);
    cudaFree(h->an_edges); cudaFree(h->an_freq); cudaFree(h->prep_mm);
    cudaFree(h->sal_g0); cudaFree(h->sal_g1); cudaFree(h->sal_on); cudaFree(h->sal_off); cudaFree(h->sal_prefix);
    cudaFree(h->sal_integ); cudaFree(h->sal_son); cudaFree(h->sal_soff); cudaFree(h->sal_max);
    cudaFree(h->ccl_L); cudaFree(h->ccl_area); cudaFree(h->ccl_touch); cudaFree(h->ccl_ncomp);
    cudaFree(h->ccl_sr); cudaFree(h->ccl_sc); cudaFree(h->ccl_best);
    cudaFree(h->stage_in); cudaFree(h->stage_f32); cudaFree(h->stage_f16);
    if (h->h_done) cudaFreeHost(h->h_done);
    if (h->h_order) cudaFreeHost(h->h_order);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->ev_tp) cudaEventDestroy(h->ev_tp);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (auto& ev : h->ev_launch) if (ev) cudaEventDestroy(ev);
    for (int g = 0; g < kMaxGroups - 1; ++g) {
        if (h->group_stream[g]) cudaStreamDestroy(h->group_stream[g]);
        for (int w = 0; w < 2; ++w) if (h->ev_group[w][g]) cudaEventDestroy(h->ev_group[w][g]);
    }
    delete h;
    return TEEFLOW_OK;
}

static double* param_slot_d(teeflow_params& p, const char* key) {
    if (!strcmp(key, "spec_factor") || !strcmp(key, "stepped")) return nullptr;   // handle-level knobs, see teeflow_set_param
    if (!strcmp(key, "tau")) return &p.tau;
    if (!strcmp(key, "lambda")) return &p.lambda;
    if (!strcmp(key, "theta")) return &p.theta;
    if (!strcmp(key, "epsilon")) return &p.epsilon;
    if (!strcmp(key, "scale_step")) return &p.scale_step;
    return nullptr;
}
static int32_t* param_slot_i(teeflow_params& p, const char* key) {
    if (!strcmp(key, "nscales")) return &p.nscales;
    if (!strcmp(key, "warps")) return &p.warps;
    if (!strcmp(key, "inner_iterations")) return &p.inner_iterations;
    if (!strcmp(key, "outer_iterations")) return &p.outer_iterations;
    if (!strcmp(key, "median_filtering")) return &p.median_filtering;
    if (!strcmp(key, "max_slots")) return &p.max_slots;
    return nullptr;
}

int teeflow_set_param(teeflow_handle h, const char* key, double value) {
    if (!h || !key) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL handle or key");
    if (!strcmp(key, "spec_factor")) {   // two-iteration passes of the inner loop (0: off); results never depend on it
        if (!(value >= 0)) return fail(h, TEEFLOW_ERR_BAD_ARG, "spec_factor must be >= 0");
        h->spec_factor = (float)value;
        return TEEFLOW_OK;
    }
    if (!strcmp(key, "stepped")) {       // 1: one launch per phase step (profiling); results never depend on it
        h->stepped = value != 0;
        return TEEFLOW_OK;
    }
    teeflow_params np = h->p;
    if (double* d = param_slot_d(np, key)) *d = value;
    else if (int32_t* i = param_slot_i(np, key)) {
        if (value != std::floor(value)) return fail(h, TEEFLOW_ERR_BAD_ARG, "parameter %s must be an integer", key);
        *i = (int32_t)value;
    } else return fail(h, TEEFLOW_ERR_BAD_ARG, "unknown parameter '%s'", key);
    int rc = validate_params(h, np);
    if (rc) return rc;
    h->p = np;
    return TEEFLOW_OK;
}

int teeflow_get_param(teeflow_handle h, const char* key, double* value) {
    if (!h || !key || !value) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL argument");
    if (!strcmp(key, "spec_factor")) { *value = h->spec_factor; return TEEFLOW_OK; }
    if (!strcmp(key, "stepped")) { *value = h->stepped; return TEEFLOW_OK; }
    if (double* d = param_slot_d(h->p, key)) { *value = *d; return TEEFLOW_OK; }
    if (int32_t* i = param_slot_i(h->p, key)) { *value = (double)*i; return TEEFLOW_OK; }
    return fail(h, TEEFLOW_ERR_BAD_ARG, "unknown parameter '%s'", key);
}

int teeflow_time_launches(teeflow_handle h, int n) {
    if (!h || n < 0 || n > teeflow_engine::kMaxLaunchEvents) return fail(h, TEEFLOW_ERR_BAD_ARG, "n must be in [0,%d]", teeflow_engine::kMaxLaunchEvents);
    CU_TRY(h, cudaSetDevice(h->device));
    for (int i = 0; i <= n && n > 0; ++i)
        if (!h->ev_launch[i]) CU_TRY(h, cudaEventCreate(&h->ev_launch[i]));
    h->n_launch_events = n;
    h->n_launches_timed = 0;
    return TEEFLOW_OK;
}

int teeflow_get_launch_times(teeflow_handle h, float* ms, int cap) {
    if (!h || (!ms && cap > 0) || cap < 0) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    for (int i = 0; i < h->n_launches_timed && i < cap; ++i) ms[i] = h->launch_ms[i];
    return h->n_launches_timed;
}

int teeflow_level_sizes(teeflow_handle h, int H, int W, int32_t* Hs, int32_t* Ws) {
    if (!h || !Hs || !Ws || H <= 0 || W <= 0) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    int hs[kMaxLevels], ws[kMaxLevels];
    const int L = level_geometry(h->p, H, W, hs, ws);
    for (int i = 0; i < L; ++i) { Hs[i] = hs[i]; Ws[i] = ws[i]; }
    return L;
}

}  // extern "C"

// ---- self test of the exact-division fast path (diagnostics entry point, used by the GPU tests)
__global__ void selftest_division_kernel(unsigned long long n, unsigned long long seed, int mode,
                                         unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        // splitmix64 -> two floats with controlled exponents
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const unsigned ma = (unsigned)z & 0x7FFFFFu, mb = (unsigned)(z >> 23) & 0x7FFFFFu;
        const unsigned sa = (unsigned)(z >> 46) & 1u;
        int ea, eb;
        if (mode == 0) {          // dual update: b = 1 + taut*|grad u| in [1, 2^24), |a| in [2^-110, 2^110)
            ea = 127 - 110 + (int)((z >> 47) % 220); eb = 127 + (int)((z >> 55) % 24);
        } else if (mode == 1) {   // thresholding: b = grad in [2^-23, 2^24), |a| <= l_t * b
            eb = 127 - 23 + (int)((z >> 53) % 47); ea = eb - 5 - (int)((z >> 47) % 40);
        } else {                  // wide: everything the guard may let through or reject
            ea = 1 + (int)((z >> 47) % 253); eb = 1 + (int)((z >> 55) % 253);
        }
        float a = __uint_as_float((sa << 31) | ((unsigned)ea << 23) | ma);
        const float b = __uint_as_float(((unsigned)eb << 23) | mb);
        if ((z >> 62) == 3ull && mode != 2) a = sa ? -0.0f : 0.0f;     // zero numerators are common
        float q;
        if (mode == 0) {          // the dual update's guard + shared-reciprocal path
            q = dual_ok(dual_num_tiny(a), fabsf(a), b) ? div_with_rcp(a, b, refined_rcp(b)) : __fdiv_rn(a, b);
        } else {
            q = div_exact(a, b);
        }
        const float w = __fdiv_rn(a, b);
        bad += (__float_as_uint(q) != __float_as_uint(w));
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" TEEFLOW_API int teeflow_selftest_division(teeflow_handle h, int mode, int64_t n, uint64_t seed,
                                                     int64_t* mismatches) {
    if (!h || !mismatches || n < 0 || mode < 0 || mode > 2) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    unsigned long long* d = nullptr;
    CU_TRY(h, cudaMalloc(&d, sizeof(*d)));
    CU_TRY(h, cudaMemset(d, 0, sizeof(*d)));
    selftest_division_kernel<<<h->num_sms * 8, 256>>>((unsigned long long)n, seed, mode, d);
    unsigned long long out = 0;
    cudaError_t e = cudaMemcpy(&out, d, sizeof(out), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, TEEFLOW_ERR_CUDA, "selftest failed: %s", cudaGetErrorString(e));
    *mismatches = (int64_t)out;
    return TEEFLOW_OK;
}

// ---- self test of the packed float-float hypot (fast form of the dual update)
__global__ void selftest_hypot_kernel(unsigned long long n, unsigned long long seed, int mode, float negzero,
                                      unsigned long long* out) {
    unsigned long long bad = 0, rej = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float v[4];
        for (int k = 0; k < 2; ++k) {     // two operand pairs per call: the fast form works on both channels at once
            unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (2 * i + k + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            const unsigned ma = (unsigned)z & 0x7FFFFFu, mb = (unsigned)(z >> 23) & 0x7FFFFFu;
            const unsigned sa = (unsigned)(z >> 46) & 1u, sb = (unsigned)(z >> 47) & 1u;
            int ea, eb;
            if (mode == 0) { ea = 127 - 30 + (int)((z >> 48) % 35); eb = ea - 3 + (int)((z >> 56) % 7); }
            else if (mode == 1) { ea = 127 - 40 + (int)((z >> 48) % 60); eb = 127 - 60 + (int)((z >> 56) % 80); }
            else { ea = (int)((z >> 48) % 255); eb = (int)((z >> 56) % 255); }
            float a = __uint_as_float((sa << 31) | ((unsigned)ea << 23) | ma);
            float b = __uint_as_float((sb << 31) | ((unsigned)eb << 23) | mb);
            if (mode == 1) {
                const unsigned sel = (unsigned)(z >> 60) & 7u;
                if (sel == 0) a = 0.0f;
                if (sel == 1) b = -0.0f;
                if (sel == 2) { a = 0.0f; b = 0.0f; }
                if (sel == 3) { a = (float)(int)(ma >> 12); b = 0.0f; }                            // exact results
                if (sel == 4) { a = 3.0f * (float)(1 + (ma >> 14)); b = 4.0f * (float)(1 + (ma >> 14)); }  // 3-4-5
            }
            v[2 * k] = a; v[2 * k + 1] = b;
        }
        bool ok;
        const float2 g = hypot2_fast(make_float2(v[0], v[2]), make_float2(v[1], v[3]), negzero, ok);
        if (!ok) { ++rej; continue; }
        const float w0 = hypot_f(v[0], v[1]), w1 = hypot_f(v[2], v[3]);
        bad += (__float_as_uint(g.x) != __float_as_uint(w0)) || (__float_as_uint(g.y) != __float_as_uint(w1));
    }
    if (bad) atomicAdd(out, bad);
    if (rej) atomicAdd(out + 1, rej);
}

extern "C" TEEFLOW_API int teeflow_selftest_hypot(teeflow_handle h, int mode, int64_t n, uint64_t seed,
                                                  int64_t* mismatches, int64_t* rejected) {
    if (!h || !mismatches || !rejected || n < 0 || mode < 0 || mode > 2) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    unsigned long long* d = nullptr;
    CU_TRY(h, cudaMalloc(&d, 2 * sizeof(*d)));
    CU_TRY(h, cudaMemset(d, 0, 2 * sizeof(*d)));
    selftest_hypot_kernel<<<h->num_sms * 8, 256>>>((unsigned long long)n, seed, mode, -0.0f, d);
    unsigned long long out[2] = {0, 0};
    cudaError_t e = cudaMemcpy(out, d, sizeof(out), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, TEEFLOW_ERR_CUDA, "selftest failed: %s", cudaGetErrorString(e));
    *mismatches = (int64_t)out[0];
    *rejected = (int64_t)out[1];
    return TEEFLOW_OK;
}

template <typename T>
static cudaError_t regrow(T*& ptr, size_t count) {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    return cudaMalloc((void**)&ptr, count * sizeof(T));
}

static int ensure_workspace(teeflow_engine* h, size_t n_frames, size_t pyr_stride, int S, size_t slot_elems,
                            size_t max_tiles, size_t n_pairs) {
    // a capacity is zeroed before its buffers are reallocated, so a failed cudaMalloc cannot leave a stale size
    if (n_frames * pyr_stride > h->cap_frames * h->cap_pyr_stride) {
        h->cap_frames = 0; h->cap_pyr_stride = 0;
        CU_TRY(h, regrow(h->pyrI, n_frames * pyr_stride));
        CU_TRY(h, regrow(h->pyrG, n_frames * pyr_stride));
        h->cap_frames = n_frames; h->cap_pyr_stride = pyr_stride;
    }
    if ((size_t)S * slot_elems > h->cap_plane_elems) {
        h->cap_plane_elems = 0; h->planes = nullptr;
        if (h->planes_raw) cudaFree(h->planes_raw);
        h->planes_raw = nullptr;
        CU_TRY(h, cudaMalloc(&h->planes_raw, (size_t)S * slot_elems * sizeof(float2) + kPlaneAlign + 4096));   // + slack past the last row
        h->planes = (float2*)(((uintptr_t)h->planes_raw + kPlaneAlign - 1) / kPlaneAlign * kPlaneAlign);
        h->cap_plane_elems = (size_t)S * slot_elems;
    }
    if ((size_t)S * max_tiles > h->cap_tiles) {
        h->cap_tiles = 0;
        CU_TRY(h, regrow(h->partial, (size_t)S * max_tiles));
        h->cap_tiles = (size_t)S * max_tiles;
    }
    if (!h->slots) {
        CU_TRY(h, regrow(h->slots, (size_t)2 * kMaxSlots));
        CU_TRY(h, regrow(h->arrive, (size_t)kMaxSlots));
    }
    if (n_pairs > h->cap_pairs) {
        h->cap_pairs = 0;
        CU_TRY(h, regrow(h->pair_lists, 5 * n_pairs));
        if (h->h_order) cudaFreeHost(h->h_order);
        h->h_order = nullptr;
        CU_TRY(h, cudaMallocHost(&h->h_order, sizeof(int) * 2 * n_pairs));
        CU_TRY(h, regrow(h->counters, n_pairs * kMaxLevels * 3));
        CU_TRY(h, regrow(h->bg, n_pairs));
        h->cap_pairs = n_pairs;
    }
    return TEEFLOW_OK;
}

// host_f32 / host_f16 (optional): host mirrors of the output buffers; the flow of a finished pair is copied out while
// the other pairs are still being solved (a frame pair is an independent unit, its result is final once written)
// ---- tensor maps of the TMA-staged inner iteration (op_inner_tma): per pyramid level three 4-D maps over the slot
// planes, dimensions (column, plane, row, slot) in 8-byte elements; elements outside a level's W x H read as zero.
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)p;
    }
    return fn;
}

static int build_tensor_maps(teeflow_engine* h, const EngineParams& P, int pitch, int n_slots, CUtensorMap* maps /* [L][3] */) {
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) return fail(h, TEEFLOW_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t PB = (cuuint64_t)pitch * 8u, ROWB = PB * kPlanes, SLOTB = (cuuint64_t)P.slot_stride * 8u;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    for (int l = 0; l < P.L; ++l) {
        const cuuint64_t Wl = (cuuint64_t)P.lv[l].W, Hl = (cuuint64_t)P.lv[l].H;
        CUresult r;
        {   // [0] one plane, 34 columns x kTR rows (U[ucur], CA)
            const cuuint64_t dims[4] = {Wl, (cuuint64_t)kPlanes, Hl, (cuuint64_t)n_slots};
            const cuuint64_t strides[3] = {PB, ROWB, SLOTB};
            const cuuint32_t box[4] = {(cuuint32_t)kTBW, 1, (cuuint32_t)kTR, 1};
            r = enc(&maps[l * 3 + 0], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)P.planes, dims, strides, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(h, TEEFLOW_ERR_CUDA, "cuTensorMapEncodeTiled (plane box, level %d) failed: %d", l, (int)r);
            // [1] PX[pcur] and PY[pcur] (two planes apart: plane step 2), 34 columns, kTR rows
            const cuuint32_t boxp[4] = {(cuuint32_t)kTBW, 4, (cuuint32_t)kTR, 1};
            const cuuint32_t stepp[4] = {1, 2, 1, 1};
            r = enc(&maps[l * 3 + 1], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)P.planes, dims, strides, boxp, stepp,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(h, TEEFLOW_ERR_CUDA, "cuTensorMapEncodeTiled (dual box, level %d) failed: %d", l, (int)r);
        }
        {   // [2] the CB plane of the even image rows: (rho_c(y), rho_c(y + 1)) per element, 34 columns x kTR/2 row pairs
            const cuuint64_t dims[4] = {Wl, 1, (Hl + 1) / 2, (cuuint64_t)n_slots};
            const cuuint64_t strides[3] = {PB, 2 * ROWB, SLOTB};
            const cuuint32_t box[4] = {(cuuint32_t)kTBW, 1, (cuuint32_t)(kTR / 2), 1};
            r = enc(&maps[l * 3 + 2], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)(P.planes + (size_t)PL_CB * pitch), dims, strides, box,
                    ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(h, TEEFLOW_ERR_CUDA, "cuTensorMapEncodeTiled (rho_c box, level %d) failed: %d", l, (int)r);
        }
    }
    return TEEFLOW_OK;
}

static int run_pairs_impl(teeflow_engine* h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                          int64_t frame_stride, const int32_t* pair_a, const int32_t* pair_b, const int32_t* out_index,
                          const int32_t* dup_index, int n_pairs, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                          cudaStream_t stream, float* host_f32, void* host_f16);

// An error inside the launch loop must not return while step kernels are still queued on the group streams or
// early device-to-host copies are still writing into the caller's buffers: drain everything first.
static int run_pairs(teeflow_engine* h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                     int64_t frame_stride, const int32_t* pair_a, const int32_t* pair_b, const int32_t* out_index,
                     const int32_t* dup_index, int n_pairs, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                     cudaStream_t stream, float* host_f32 = nullptr, void* host_f16 = nullptr) {
    const int rc = run_pairs_impl(h, frames_dev, dtype, n_frames, H, W, frame_stride, pair_a, pair_b, out_index, dup_index,
                                  n_pairs, flow_f32_dev, flow_f16_dev, out_scale, stream, host_f32, host_f16);
    if (rc != TEEFLOW_OK && h) {
        const std::string msg = h->err;
        cudaStreamSynchronize(stream);
        for (int g = 0; g < kMaxGroups - 1; ++g) if (h->group_stream[g]) cudaStreamSynchronize(h->group_stream[g]);
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        cudaGetLastError();
        h->err = msg; g_last_error = msg;
    }
    return rc;
}

static int run_pairs_impl(teeflow_engine* h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                          int64_t frame_stride, const int32_t* pair_a, const int32_t* pair_b, const int32_t* out_index,
                          const int32_t* dup_index, int n_pairs, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                          cudaStream_t stream, float* host_f32, void* host_f16) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (!frames_dev || !pair_a || !pair_b || !out_index || !dup_index)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL frames or pair list");
    if (dtype != TEEFLOW_U8 && dtype != TEEFLOW_F32) return fail(h, TEEFLOW_ERR_BAD_ARG, "dtype must be TEEFLOW_U8 or TEEFLOW_F32");
    if (H < 1 || W < 1 || n_frames < 1 || (int64_t)H * W > (1 << 28)) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad frame shape %dx%d", H, W);
    int pitch_i = 0;
    while (pitch_i < 3 && W + kXMargin > kPitches[pitch_i]) ++pitch_i;
    if (pitch_i == 3 || (int64_t)(H + 2) * kPlanes * kPitches[pitch_i] >= (1ll << 31))
        return fail(h, TEEFLOW_ERR_BAD_SHAPE, "frame shape %dx%d exceeds the engine's plane layout (W <= %d)", H, W, kPitches[2] - kXMargin);
    const int pitch = kPitches[pitch_i];
    if (frame_stride < (int64_t)H * W) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "frame_stride smaller than H*W");
    if (n_pairs < 0) return fail(h, TEEFLOW_ERR_BAD_ARG, "negative pair count");
    if (!flow_f32_dev && !flow_f16_dev) return fail(h, TEEFLOW_ERR_BAD_ARG, "no output buffer");
    for (int i = 0; i < n_pairs; ++i)
        if (pair_a[i] < 0 || pair_a[i] >= n_frames || pair_b[i] < 0 || pair_b[i] >= n_frames || out_index[i] < 0)
            return fail(h, TEEFLOW_ERR_BAD_ARG, "pair %d references a frame outside [0,%d)", i, n_frames);
    CU_TRY(h, cudaSetDevice(h->device));
    if (h->pending) return fail(h, TEEFLOW_ERR_STATE, "an asynchronous run is in flight: call teeflow_finish() first");
    memset(&h->st, 0, sizeof(h->st));
    h->st.n_pairs = n_pairs;
    if (n_pairs == 0) return TEEFLOW_OK;

    EngineParams P;
    memset(&P, 0, sizeof(P));
    int Hs[kMaxLevels], Ws[kMaxLevels];
    const int L = level_geometry(h->p, H, W, Hs, Ws);
    h->st.n_levels = L;
    long long off = 0;
    for (int l = 0; l < L; ++l) {
        LevelGeom& g = P.lv[l];
        g.H = Hs[l]; g.W = Ws[l];
        g.in_sx = (g.W + kIW - 1) / kIW; g.in_items = g.in_sx * ((g.H + kIR - 1) / kIR);
        g.pw_sx = (g.W + 31) / 32; g.pw_items = g.pw_sx * ((g.H + kPR - 1) / kPR);
        g.in2_sx = (g.W + kIW2 - 1) / kIW2; g.in2_items = g.in2_sx * ((g.H + kIR2 - 1) / kIR2);
        g.pyr_off = off;
        off += ((long long)g.H * g.W + 63) / 64 * 64;
        g.scaled_eps = (float)(h->p.epsilon * h->p.epsilon * (double)(g.H * g.W));
        if (l + 1 < L) {
            // resize(u(l+1), u(l), size(l)): inv_scale = dsize/ssize; scale = 1./inv_scale (imgproc/resize.cpp)
            g.up_sx = 1.0 / ((double)Ws[l] / (double)Ws[l + 1]);
            g.up_sy = 1.0 / ((double)Hs[l] / (double)Hs[l + 1]);
        }
    }
    const int S = std::min(h->p.max_slots > 0 ? h->p.max_slots : 64, std::min(n_pairs, kMaxSlots));
    P.L = L; P.S = S; P.n_pairs = n_pairs;
    P.warps = h->p.warps; P.inner = h->p.inner_iterations; P.outer = h->p.outer_iterations; P.median = h->p.median_filtering;
    P.l_t = (float)(h->p.lambda * h->p.theta);
    P.theta = (float)h->p.theta;
    P.taut = (float)(h->p.tau / h->p.theta);
    P.up_mul = (float)(1.0 / h->p.scale_step);
    P.out_scale = out_scale;
    P.negzero = -0.0f;
    P.frame_pyr_stride = off;
    P.pitch = pitch;
    P.slot_stride = (long long)(H + 2) * kPlanes * pitch;   // two pad rows: the inner iteration's look-ahead loads
    P.max_tiles = std::max(std::max(P.lv[0].in_items, 2 * P.lv[0].in2_items), 2 * P.lv[0].pw_items);
    P.spec_factor = h->p.inner_iterations >= 2 ? h->spec_factor : 0.0f;
    if (h->wase_w && (h->wase_H != H || h->wase_W != W))
        return fail(h, TEEFLOW_ERR_BAD_SHAPE, "WASE weight map is %dx%d but the frames are %dx%d", h->wase_H, h->wase_W, H, W);

    int rc = ensure_workspace(h, (size_t)n_frames, (size_t)off, S, (size_t)P.slot_stride, (size_t)P.max_tiles, (size_t)n_pairs);
    if (rc) return rc;
    P.pyrI = h->pyrI; P.pyrG = h->pyrG;
    for (int i = 0; i < 2; ++i) P.slots[i] = h->slots + (size_t)i * kMaxSlots;
    P.planes = h->planes;
    P.arrive = h->arrive; P.partial = h->partial;
    P.next_pair = h->ctl; P.pairs_done = h->ctl + 1; P.item_counter = h->ctl + 2;
    P.spec_stats = h->ctl + 2 + 2 * kMaxGroups;
    P.pair_a = h->pair_lists; P.pair_b = h->pair_lists + h->cap_pairs;
    P.out_index = h->pair_lists + 2 * h->cap_pairs; P.dup_index = h->pair_lists + 3 * h->cap_pairs;
    P.done_order = h->pair_lists + 4 * h->cap_pairs;
    P.counters_out = h->counters;
    P.wase_w = h->wase_w;
    P.bg_out = h->bg;
    P.flow_f32 = (float2*)flow_f32_dev;
    P.flow_f16 = (uint32_t*)flow_f16_dev;
    alignas(64) CUtensorMap tmaps_host[kMaxLevels * 3];   // host temporary: copied before the first stream synchronisation below
    if (kTma) {
        rc = build_tensor_maps(h, P, pitch, S, tmaps_host);
        if (rc) return rc;
        if (!h->tmaps) CU_TRY(h, cudaMalloc(&h->tmaps, sizeof(CUtensorMap) * kMaxLevels * 3));
        CU_TRY(h, cudaMemcpyAsync(h->tmaps, tmaps_host, sizeof(CUtensorMap) * 3 * L, cudaMemcpyHostToDevice, stream));
        P.tmaps = h->tmaps;
    }

    CU_TRY(h, cudaEventRecord(h->ev_t0, stream));
    CU_TRY(h, cudaMemsetAsync(h->bg, 0, sizeof(float) * n_pairs, stream));
    CU_TRY(h, cudaMemsetAsync(P.done_order, 0xFF, sizeof(int) * n_pairs, stream));
    // pair lists (small) -- pageable host memory: the copies are staged by the runtime before the call returns
    CU_TRY(h, cudaMemcpyAsync((void*)P.pair_a, pair_a, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream));
    CU_TRY(h, cudaMemcpyAsync((void*)P.pair_b, pair_b, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream));
    CU_TRY(h, cudaMemcpyAsync((void*)P.out_index, out_index, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream));
    CU_TRY(h, cudaMemcpyAsync((void*)P.dup_index, dup_index, sizeof(int) * n_pairs, cudaMemcpyHostToDevice, stream));

    // ---- image pyramid + packed gradients, once per frame
    {
        const int npx = H * W;
        dim3 g0((unsigned)std::min((npx + 255) / 256, 1024), (unsigned)n_frames);
        pyr_level0_kernel<<<g0, 256, 0, stream>>>(frames_dev, dtype, (long long)frame_stride, n_frames, npx, h->pyrI, off);
        const dim3 blk(32, 8);
        for (int l = 0; l < L; ++l) {
            const LevelGeom& g = P.lv[l];
            if (l > 0) {
                const LevelGeom& s = P.lv[l - 1];
                dim3 grd((g.W + 31) / 32, (g.H + 7) / 8, (unsigned)n_frames);
                pyr_down_kernel<<<grd, blk, 0, stream>>>(h->pyrI, off, n_frames, s.pyr_off, s.H, s.W, g.pyr_off, g.H, g.W,
                                                         1.0 / h->p.scale_step);
            }
            dim3 grd((g.W + 31) / 32, (g.H + 7) / 8, (unsigned)n_frames);
            pyr_pack_kernel<<<grd, blk, 0, stream>>>(h->pyrI, h->pyrG, off, n_frames, g.pyr_off, g.H, g.W);
        }
        CU_TRY(h, cudaGetLastError());
        h->st.kernel_launches = 1 + 2 * (long long)L - 1;
    }

    // ---- launch geometry: G slot groups on G streams, persistent grid of resident CTAs
    const int G = (S >= 2 * kMinGroupSlots && h->groups > 1) ? std::min(h->groups, kMaxGroups) : 1;
    // every launch fills all resident CTA slots; giving each group 1/G of them, so that the groups' launches are
    // co-resident on every SM all the time, was measured slower (1037 vs 1134 pairs/s): few slots at a time sweeping
    // the same image rows keeps DRAM pages open
    const int grid = h->num_sms * h->ctas_per_sm[pitch_i];

    // ---- slot table: the first S pairs start at the coarsest level, parity 0
    {
        std::vector<Slot> init((size_t)kMaxSlots);
        memset(init.data(), 0, sizeof(Slot) * init.size());
        for (int s = 0; s < kMaxSlots; ++s) {
            Slot& sl = init[s];
            sl.pair = -1; sl.phase = PH_IDLE;
            if (s < S) { sl.pair = s; sl.phase = PH_LEVEL_INIT; sl.level = L - 1; sl.error = FLT_MAX; }
        }
        CU_TRY(h, cudaMemcpyAsync(h->slots, init.data(), sizeof(Slot) * kMaxSlots, cudaMemcpyHostToDevice, stream));
        CU_TRY(h, cudaMemsetAsync(h->arrive, 0, sizeof(unsigned) * kMaxSlots, stream));
        int ctl0[kCtlInts] = {0};
        ctl0[0] = S;
        // strip counters of launch parity 0: the first strip of a warp is its own index (tvl1_step_kernel)
        for (int g = 0; g < kMaxGroups; ++g) ctl0[2 + 2 * g] = grid * kWarpsPerCta;
        CU_TRY(h, cudaMemcpyAsync(h->ctl, ctl0, sizeof(ctl0), cudaMemcpyHostToDevice, stream));
        CU_TRY(h, cudaStreamSynchronize(stream));  // `init` / ctl0 / pair lists are host temporaries
        CU_TRY(h, cudaEventRecord(h->ev_tp, stream));
    }

    // early copy-out of finished pairs (host-buffer entry points)
    const bool copy_out = host_f32 || host_f16;
    const size_t npx_out = (size_t)H * W;
    int n_copied = 0;
    auto copy_pair = [&](int pair) -> cudaError_t {
        const int idx[2] = {out_index[pair], dup_index[pair]};
        for (int k = 0; k < 2; ++k) {
            if (idx[k] < 0) continue;
            cudaError_t e = cudaSuccess;
            if (host_f32) e = cudaMemcpyAsync(host_f32 + (size_t)idx[k] * npx_out * 2, flow_f32_dev + (size_t)idx[k] * npx_out * 2,
                                              npx_out * 8, cudaMemcpyDeviceToHost, h->copy_stream);
            if (e == cudaSuccess && host_f16)
                e = cudaMemcpyAsync((char*)host_f16 + (size_t)idx[k] * npx_out * 4, (char*)flow_f16_dev + (size_t)idx[k] * npx_out * 4,
                                    npx_out * 4, cudaMemcpyDeviceToHost, h->copy_stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    auto finish_stats = [&](long long launches, int grid_ctas, int applied, int discarded) -> int {
        CU_TRY(h, cudaEventElapsedTime(&h->st.device_ms, h->ev_t0, h->ev_t1));
        CU_TRY(h, cudaEventElapsedTime(&h->st.pyramid_ms, h->ev_t0, h->ev_tp));
        CU_TRY(h, cudaEventElapsedTime(&h->st.solver_ms, h->ev_tp, h->ev_t1));
        h->st.double_steps = applied; h->st.double_steps_discarded = discarded;
        h->st.solver_launches = launches;
        h->st.kernel_launches += launches;
        h->st.n_slots = S;
        h->st.grid_ctas = grid_ctas;
        return TEEFLOW_OK;
    };

    // ---- dataflow scheduler: ONE cooperative launch solves every pair (tvl1_flow_kernel); the host only watches the
    // completion list the kernel writes into mapped pinned memory and copies finished flows out meanwhile
    if (!h->stepped && h->n_launch_events == 0) {
        if ((unsigned)P.max_tiles > kTaskItemsMax || n_pairs > kTaskPairsMax)
            return fail(h, TEEFLOW_ERR_BAD_SHAPE, "run too large for the dataflow scheduler's task descriptors");
        if (!h->tasks) CU_TRY(h, cudaMalloc(&h->tasks, sizeof(Task) * kTaskRing));
        if (!h->flow_ctl) CU_TRY(h, cudaMalloc(&h->flow_ctl, sizeof(FlowCtl)));
        if (!h->flow_stats) CU_TRY(h, cudaMalloc(&h->flow_stats, sizeof(unsigned long long) * 32));
        CU_TRY(h, cudaMemsetAsync(h->flow_stats, 0, sizeof(unsigned long long) * 32, stream));
        if ((size_t)n_pairs > h->cap_flow_order) {
            h->cap_flow_order = 0;
            if (h->h_flow_order) cudaFreeHost(h->h_flow_order);
            h->h_flow_order = nullptr;
            CU_TRY(h, cudaHostAlloc(&h->h_flow_order, sizeof(int) * n_pairs, cudaHostAllocMapped));
            CU_TRY(h, cudaHostGetDevicePointer(&h->d_flow_order, h->h_flow_order, 0));
            h->cap_flow_order = (size_t)n_pairs;
        }
        for (int i = 0; i < n_pairs; ++i) h->h_flow_order[i] = -1;
        const unsigned items0 = (unsigned)P.lv[L - 1].pw_items;      // strips of a level-init task at the coarsest level
        std::vector<Task> t0((size_t)S);
        memset(t0.data(), 0, sizeof(Task) * t0.size());
        for (int s = 0; s < S; ++s) {
            Task& t = t0[s];
            t.seq = (unsigned)s + 1u; t.first = (unsigned)s * items0;
            t.what = items0 | ((unsigned)PH_LEVEL_INIT << 20) | ((unsigned)(L - 1) << 24);
            t.who = (unsigned)s | ((unsigned)s << 9);
        }
        FlowCtl c0;
        memset(&c0, 0, sizeof(c0));
        c0.alloc = ((unsigned long long)S << 32) | ((unsigned long long)S * items0);
        c0.next_pair = S;
        CU_TRY(h, cudaMemsetAsync(h->tasks, 0, sizeof(Task) * kTaskRing, stream));
        CU_TRY(h, cudaMemcpyAsync(h->tasks, t0.data(), sizeof(Task) * S, cudaMemcpyHostToDevice, stream));
        CU_TRY(h, cudaMemcpyAsync(h->flow_ctl, &c0, sizeof(c0), cudaMemcpyHostToDevice, stream));
        CU_TRY(h, cudaStreamSynchronize(stream));                     // t0 / c0 are host temporaries
        CU_TRY(h, cudaEventRecord(h->ev_tp, stream));
        EngineParams Pf = P;
        Pf.tasks = h->tasks; Pf.flow = h->flow_ctl;
        Pf.spec_stats = &h->flow_ctl->spec_applied;
        Pf.host_done = copy_out ? h->d_flow_order : nullptr;
        Pf.flow_stats = h->flow_stats;
        Pf.watchdog_cycles = 20000000000ll;                           // ~10 s at 2 GHz: a hang becomes an error code
        const int grid_f = h->num_sms * h->ctas_per_sm_flow[pitch_i];
        void* args[] = {(void*)&Pf};
        CU_TRY(h, cudaLaunchCooperativeKernel((const void*)flow_kernel_for(pitch), dim3((unsigned)grid_f), dim3(kThreads), args, 0, stream));
        CU_TRY(h, cudaEventRecord(h->ev_t1, stream));
        if (h->async_mode && !copy_out) {               // the caller fetches the verdict with teeflow_finish()
            h->pending = 1; h->pending_pairs = n_pairs; h->pending_grid = grid_f; h->pending_stream = stream;
            h->st.n_slots = S;
            return TEEFLOW_OK;
        }
        if (copy_out) {
            volatile int* order = h->h_flow_order;
            for (;;) {
                while (n_copied < n_pairs && order[n_copied] >= 0) CU_TRY(h, copy_pair(order[n_copied++]));
                if (n_copied == n_pairs) break;
                const cudaError_t q = cudaEventQuery(h->ev_t1);
                if (q == cudaSuccess) break;
                if (q != cudaErrorNotReady) return fail(h, TEEFLOW_ERR_CUDA, "dataflow kernel failed: %s", cudaGetErrorString(q));
                std::this_thread::sleep_for(std::chrono::microseconds(50));
            }
        }
        CU_TRY(h, cudaStreamSynchronize(stream));
        FlowCtl c1;
        CU_TRY(h, cudaMemcpy(&c1, h->flow_ctl, sizeof(c1), cudaMemcpyDeviceToHost));
        CU_TRY(h, cudaMemcpy(h->flow_stats_host, h->flow_stats, sizeof(h->flow_stats_host), cudaMemcpyDeviceToHost));
        if (c1.abort) return fail(h, TEEFLOW_ERR_STATE, c1.abort == 1 ? "dataflow scheduler watchdog: a warp waited too long for a task"
                                                       : c1.abort == 3 ? "a warp waited too long for a TMA box of the inner iteration"
                                                                       : "dataflow scheduler lost the task ring");
        if (c1.pairs_done < n_pairs) return fail(h, TEEFLOW_ERR_STATE, "scheduler stopped with %d of %d pairs done", c1.pairs_done, n_pairs);
        if (copy_out) {
            for (; n_copied < n_pairs; ++n_copied) {
                const int pair = h->h_flow_order[n_copied];
                if (pair < 0 || pair >= n_pairs) return fail(h, TEEFLOW_ERR_STATE, "completion list is corrupt at %d", n_copied);
                CU_TRY(h, copy_pair(pair));
            }
            CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
        }
        return finish_stats(1, grid_f, c1.spec_applied, c1.spec_discarded);
    }

    // ---- super-steps (stepped scheduler).  The slots are split into groups that step on separate streams: while one group's launch
    const step_kernel_t step_kernel = step_kernel_for(pitch);
    const int chunk = 16;
    EngineParams Pg[kMaxGroups];
    cudaStream_t gs[kMaxGroups];
    for (int g = 0; g < G; ++g) {
        Pg[g] = P;
        Pg[g].slot0 = (int)((long long)S * g / G);
        Pg[g].S = (int)((long long)S * (g + 1) / G) - Pg[g].slot0;
        Pg[g].item_counter = h->ctl + 2 + 2 * g;
        gs[g] = g == 0 ? stream : h->group_stream[g - 1];
        if (g > 0) CU_TRY(h, cudaStreamWaitEvent(gs[g], h->ev_tp, 0));      // pyramid + slot table are ready
    }
    // upper bound on steps per pair: per level 1 init + warps * (1 + outer * (1 + inner)), + wase + final
    const long long steps_per_pair = (long long)L * (1 + (long long)h->p.warps * (1 + (long long)h->p.outer_iterations * (1 + h->p.inner_iterations))) + 2;
    const long long max_steps = steps_per_pair * ((n_pairs + S - 1) / S + 1) * G + 2 * chunk;
    long long step = 0;
    int n_chunks = 0;
    h->n_launches_timed = 0;
    bool done = false;
    volatile int* hd = h->h_done;
    hd[0] = hd[1] = 0;
    while (!done) {
        if (step > max_steps) return fail(h, TEEFLOW_ERR_STATE, "scheduler exceeded %lld steps", max_steps);
        for (int k = 0; k < chunk; ++k, ++step)
            for (int g = 0; g < G; ++g) {
                const bool timed = g == 0 && step < h->n_launch_events;
                if (timed && step == 0) CU_TRY(h, cudaEventRecord(h->ev_launch[0], gs[0]));
                step_kernel<<<grid, kThreads, 0, gs[g]>>>(Pg[g], (int)(step & 1));
                if (timed) { CU_TRY(h, cudaEventRecord(h->ev_launch[step + 1], gs[0])); h->n_launches_timed = (int)step + 1; }
            }
        CU_TRY(h, cudaGetLastError());
        const int which = n_chunks & 1;
        // join the group streams into the caller's stream at the chunk boundary, then sample the done counter
        for (int g = 1; g < G; ++g) {
            CU_TRY(h, cudaEventRecord(h->ev_group[which][g - 1], gs[g]));
            CU_TRY(h, cudaStreamWaitEvent(stream, h->ev_group[which][g - 1], 0));
        }
        CU_TRY(h, cudaMemcpyAsync((void*)(hd + which), h->ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (copy_out)
            CU_TRY(h, cudaMemcpyAsync(h->h_order + (size_t)which * h->cap_pairs, P.done_order, sizeof(int) * n_pairs,
                                      cudaMemcpyDeviceToHost, stream));
        CU_TRY(h, cudaEventRecord(h->ev[which], stream));
        for (int g = 1; g < G; ++g) CU_TRY(h, cudaStreamWaitEvent(gs[g], h->ev[which], 0));
        if (++n_chunks >= 2) {
            const int prev = which ^ 1;
            CU_TRY(h, cudaEventSynchronize(h->ev[prev]));
            if (hd[prev] >= n_pairs) done = true;
            if (copy_out) {
                // every launch up to that chunk has completed: the flows of the pairs it lists are final
                const int* order = h->h_order + (size_t)prev * h->cap_pairs;
                while (n_copied < n_pairs && n_copied < hd[prev] && order[n_copied] >= 0)
                    CU_TRY(h, copy_pair(order[n_copied++]));
            }
        }
    }
    CU_TRY(h, cudaEventRecord(h->ev_t1, stream));
    CU_TRY(h, cudaStreamSynchronize(stream));
    for (int g = 1; g < G; ++g) CU_TRY(h, cudaStreamSynchronize(gs[g]));
    if (hd[0] < n_pairs && hd[1] < n_pairs) {
        // the last issued chunk may be the one that finished the work
        int d = 0;
        CU_TRY(h, cudaMemcpy(&d, h->ctl + 1, sizeof(int), cudaMemcpyDeviceToHost));
        if (d < n_pairs) return fail(h, TEEFLOW_ERR_STATE, "scheduler stopped with %d of %d pairs done", d, n_pairs);
    }
    if (copy_out) {
        // the rest: everything is complete now
        CU_TRY(h, cudaMemcpy(h->h_order, P.done_order, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost));
        for (; n_copied < n_pairs; ++n_copied) {
            const int pair = h->h_order[n_copied];
            if (pair < 0 || pair >= n_pairs) return fail(h, TEEFLOW_ERR_STATE, "completion list is corrupt at %d", n_copied);
            CU_TRY(h, copy_pair(pair));
        }
        CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
    }
    step *= G;
    for (int i = 0; i < h->n_launches_timed; ++i)
        CU_TRY(h, cudaEventElapsedTime(&h->launch_ms[i], h->ev_launch[i], h->ev_launch[i + 1]));
    int sp[2] = {0, 0};
    CU_TRY(h, cudaMemcpy(sp, h->ctl + 2 + 2 * kMaxGroups, sizeof(sp), cudaMemcpyDeviceToHost));
    return finish_stats(step, grid, sp[0], sp[1]);
}

extern "C" {

int teeflow_calc_pairs(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                       int64_t frame_stride, const int32_t* pair_a, const int32_t* pair_b, const int32_t* out_index,
                       const int32_t* dup_index, int n_pairs, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                       void* stream) {
    return run_pairs(h, frames_dev, dtype, n_frames, H, W, frame_stride, pair_a, pair_b, out_index, dup_index, n_pairs,
                     flow_f32_dev, flow_f16_dev, out_scale, (cudaStream_t)stream);
}

int teeflow_calc_clip(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                      int64_t frame_stride, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                      int duplicate_last, void* stream) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (n_frames < 2) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "a clip needs at least 2 frames, got %d", n_frames);
    const int n_pairs = n_frames - 1;
    std::vector<int32_t> a(n_pairs), b(n_pairs), o(n_pairs), d(n_pairs, -1);
    for (int i = 0; i < n_pairs; ++i) { a[i] = i; b[i] = i + 1; o[i] = i; }
    if (duplicate_last) d[n_pairs - 1] = n_pairs;
    return run_pairs(h, frames_dev, dtype, n_frames, H, W, frame_stride, a.data(), b.data(), o.data(), d.data(), n_pairs,
                     flow_f32_dev, flow_f16_dev, out_scale, (cudaStream_t)stream);
}

int teeflow_calc_clip_async(teeflow_handle h, const void* frames_dev, int dtype, int n_frames, int H, int W,
                            int64_t frame_stride, float* flow_f32_dev, void* flow_f16_dev, float out_scale,
                            int duplicate_last, void* stream) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (h->stepped || h->n_launch_events > 0)
        return fail(h, TEEFLOW_ERR_STATE, "asynchronous runs need the dataflow scheduler (stepped mode / launch timing is on)");
    h->async_mode = 1;
    const int rc = teeflow_calc_clip(h, frames_dev, dtype, n_frames, H, W, frame_stride, flow_f32_dev, flow_f16_dev, out_scale,
                                     duplicate_last, stream);
    h->async_mode = 0;
    return rc;
}

int teeflow_finish(teeflow_handle h) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (!h->pending) return TEEFLOW_OK;
    h->pending = 0;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->pending_stream));
    FlowCtl c1;
    CU_TRY(h, cudaMemcpy(&c1, h->flow_ctl, sizeof(c1), cudaMemcpyDeviceToHost));
    if (c1.abort) return fail(h, TEEFLOW_ERR_STATE, c1.abort == 1 ? "dataflow scheduler watchdog: a warp waited too long for a task"
                                                                   : "dataflow scheduler lost the task ring");
    if (c1.pairs_done < h->pending_pairs)
        return fail(h, TEEFLOW_ERR_STATE, "scheduler stopped with %d of %d pairs done", c1.pairs_done, h->pending_pairs);
    CU_TRY(h, cudaEventElapsedTime(&h->st.device_ms, h->ev_t0, h->ev_t1));
    CU_TRY(h, cudaEventElapsedTime(&h->st.pyramid_ms, h->ev_t0, h->ev_tp));
    CU_TRY(h, cudaEventElapsedTime(&h->st.solver_ms, h->ev_tp, h->ev_t1));
    h->st.double_steps = c1.spec_applied; h->st.double_steps_discarded = c1.spec_discarded;
    h->st.solver_launches = 1;
    h->st.kernel_launches += 1;
    h->st.grid_ctas = h->pending_grid;
    return TEEFLOW_OK;
}

static int ensure_stage(teeflow_engine* h, void*& ptr, size_t& cap, size_t bytes) {
    if (bytes > cap) {
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        CU_TRY(h, cudaMalloc(&ptr, bytes));
        cap = bytes;
    }
    return TEEFLOW_OK;
}

int teeflow_calc_clip_host(teeflow_handle h, const void* frames_host, int dtype, int n_frames, int H, int W,
                           float* flow_f32_host, void* flow_f16_host, float out_scale, int duplicate_last) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (!frames_host) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL frames");
    if (n_frames < 2 || H < 1 || W < 1) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad clip shape");
    if (dtype != TEEFLOW_U8 && dtype != TEEFLOW_F32) return fail(h, TEEFLOW_ERR_BAD_ARG, "dtype must be TEEFLOW_U8 or TEEFLOW_F32");
    if (!flow_f32_host && !flow_f16_host) return fail(h, TEEFLOW_ERR_BAD_ARG, "no output buffer");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t npx = (size_t)H * W;
    const size_t esz = dtype == TEEFLOW_U8 ? 1 : 4;
    const size_t n_out = (size_t)(n_frames - 1) + (duplicate_last ? 1 : 0);
    int rc;
    if ((rc = ensure_stage(h, h->stage_in, h->stage_in_bytes, npx * esz * n_frames))) return rc;
    if (flow_f32_host && (rc = ensure_stage(h, h->stage_f32, h->stage_f32_bytes, n_out * npx * 8))) return rc;
    if (flow_f16_host && (rc = ensure_stage(h, h->stage_f16, h->stage_f16_bytes, n_out * npx * 4))) return rc;
    cudaStream_t st = h->own_stream;
    CU_TRY(h, cudaMemcpyAsync(h->stage_in, frames_host, npx * esz * n_frames, cudaMemcpyHostToDevice, st));
    const int n_pairs = n_frames - 1;
    std::vector<int32_t> a(n_pairs), b(n_pairs), o(n_pairs), d(n_pairs, -1);
    for (int i = 0; i < n_pairs; ++i) { a[i] = i; b[i] = i + 1; o[i] = i; }
    if (duplicate_last) d[n_pairs - 1] = n_pairs;
    // finished pairs are copied to the host buffers while the others are still being solved
    rc = run_pairs(h, h->stage_in, dtype, n_frames, H, W, (int64_t)npx, a.data(), b.data(), o.data(), d.data(), n_pairs,
                   flow_f32_host ? (float*)h->stage_f32 : nullptr, flow_f16_host ? h->stage_f16 : nullptr, out_scale, st,
                   flow_f32_host, flow_f16_host);
    if (rc) return rc;
    CU_TRY(h, cudaStreamSynchronize(st));
    return TEEFLOW_OK;
}

int teeflow_calc_pair_host(teeflow_handle h, const void* I0_host, const void* I1_host, int dtype, int H, int W,
                           float* flow_host) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (!I0_host || !I1_host || !flow_host) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL image or output");
    if (H < 1 || W < 1) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad image shape");
    if (dtype != TEEFLOW_U8 && dtype != TEEFLOW_F32) return fail(h, TEEFLOW_ERR_BAD_ARG, "dtype must be TEEFLOW_U8 or TEEFLOW_F32");
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t npx = (size_t)H * W;
    const size_t esz = dtype == TEEFLOW_U8 ? 1 : 4;
    int rc;
    if ((rc = ensure_stage(h, h->stage_in, h->stage_in_bytes, npx * esz * 2))) return rc;
    if ((rc = ensure_stage(h, h->stage_f32, h->stage_f32_bytes, npx * 8))) return rc;
    cudaStream_t st = h->own_stream;
    CU_TRY(h, cudaMemcpyAsync(h->stage_in, I0_host, npx * esz, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync((char*)h->stage_in + npx * esz, I1_host, npx * esz, cudaMemcpyHostToDevice, st));
    rc = teeflow_calc_clip(h, h->stage_in, dtype, 2, H, W, (int64_t)npx, (float*)h->stage_f32, nullptr, 1.0f, 0, (void*)st);
    if (rc) return rc;
    CU_TRY(h, cudaMemcpyAsync(flow_host, h->stage_f32, npx * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    return TEEFLOW_OK;
}

int teeflow_get_counters(teeflow_handle h, int32_t* counters, int n_pairs_cap) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (h->pending) { const int rc = teeflow_finish(h); if (rc) return rc; }
    const int n = h->st.n_pairs;
    if (counters) {
        if (n_pairs_cap < n) return fail(h, TEEFLOW_ERR_BAD_ARG, "counter buffer holds %d pairs, need %d", n_pairs_cap, n);
        CU_TRY(h, cudaSetDevice(h->device));
        if (n > 0)
            CU_TRY(h, cudaMemcpy(counters, h->counters, sizeof(int) * (size_t)n * kMaxLevels * 3, cudaMemcpyDeviceToHost));
    }
    return n;
}

int teeflow_get_flow_stats(teeflow_handle h, uint64_t* out32) {
    if (!h || !out32) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL argument");
    for (int i = 0; i < 32; ++i) out32[i] = h->flow_stats_host[i];
    return TEEFLOW_FLOW_STATS;
}

int teeflow_get_stats(teeflow_handle h, teeflow_stats* out) {
    if (!h || !out) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL argument");
    if (h->pending) { const int rc = teeflow_finish(h); if (rc) return rc; }
    *out = h->st;
    return TEEFLOW_OK;
}

}  // extern "C"

extern "C" {

// ---- WASE background compensation (calculate_optical_flow.py:649-660)
int teeflow_wase_weights(teeflow_handle h, const uint8_t* bkgd_dev, int n_frames, int H, int W, float* w_dev,
                         void* stream) {
    if (!h || !bkgd_dev || !w_dev || n_frames < 1 || H < 1 || W < 1) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    const int n_elem = H * W * 2;
    wase_weights_kernel<<<std::min((n_elem + 255) / 256, h->num_sms * 16), 256, 0, (cudaStream_t)stream>>>(
        bkgd_dev, n_frames, n_elem, w_dev);
    CU_TRY(h, cudaGetLastError());
    return TEEFLOW_OK;
}

int teeflow_set_wase(teeflow_handle h, const float* w_dev, int H, int W) {
    if (!h) return fail(nullptr, TEEFLOW_ERR_BAD_ARG, "NULL handle");
    if (w_dev && (H < 1 || W < 1)) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad weight map shape");
    h->wase_w = w_dev; h->wase_H = H; h->wase_W = W;
    return TEEFLOW_OK;
}

int teeflow_get_backgrounds(teeflow_handle h, float* bg_host, int n_pairs_cap) {
    if (!h || !bg_host) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL argument");
    const int n = h->st.n_pairs;
    if (n_pairs_cap < n) return fail(h, TEEFLOW_ERR_BAD_ARG, "buffer holds %d pairs, need %d", n_pairs_cap, n);
    CU_TRY(h, cudaSetDevice(h->device));
    if (n > 0) CU_TRY(h, cudaMemcpy(bg_host, h->bg, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return n;
}

// ---- frame prep: img2uint8(rgb2gray(frame)) per frame (calculate_optical_flow.py:588, optical_flow_utils.py:30-31)
int teeflow_prepare_frames(teeflow_handle h, const uint8_t* rgb_dev, int n_frames, int H, int W, uint8_t* gray_dev,
                           void* stream_v) {
    if (!h || !rgb_dev || !gray_dev || n_frames < 1 || H < 1 || W < 1) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    if ((size_t)n_frames * 2 > h->prep_cap) {
        h->prep_cap = 0;
        CU_TRY(h, regrow(h->prep_mm, (size_t)n_frames * 2));
        h->prep_cap = (size_t)n_frames * 2;
    }
    std::vector<unsigned long long> init((size_t)n_frames * 2);
    for (int f = 0; f < n_frames; ++f) { init[2 * f] = 0x7ff0000000000000ull; init[2 * f + 1] = 0ull; }   // +inf, +0
    CU_TRY(h, cudaMemcpyAsync(h->prep_mm, init.data(), sizeof(unsigned long long) * init.size(), cudaMemcpyHostToDevice, stream));
    const int npx = H * W;
    const dim3 grid(std::max(1, std::min((npx + 256 * 4 - 1) / (256 * 4), 128)), n_frames);
    prep_minmax_kernel<<<grid, 256, 0, stream>>>(rgb_dev, npx, h->prep_mm);
    prep_quantize_kernel<<<grid, 256, 0, stream>>>(rgb_dev, npx, h->prep_mm, gray_dev);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(stream));     // `init` is a host temporary
    return TEEFLOW_OK;
}

// ---- saliency input stage: cv2.saliency.StaticSaliencyFineGrained.computeSaliency per frame
// (calculate_optical_flow.py:560, :586); kernels and the parity status in saliency_kernels.cuh
static const int kSalChunk = 16;

int teeflow_saliency_fine_grained(teeflow_handle h, const uint8_t* rgb_dev, int n_frames, int H, int W,
                                  float* saliency_dev, uint8_t* intensity_dev, void* stream_v) {
    if (!h || !rgb_dev || (!saliency_dev && !intensity_dev) || n_frames < 1) return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    if (H < 3 || W < 3 || (int64_t)(H + 1) * (W + 1) > (1 << 28)) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad frame shape %dx%d", H, W);
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t npx = (size_t)H * W, nint = (size_t)(H + 1) * (W + 1);
    if (npx > h->sal_cap_px || nint > h->sal_cap_int) {
        // grow to the maximum of both measures seen so far: a later frame shape with fewer integral entries but more
        // pixels (3 x 3333 then 100 x 100) must not run on the smaller allocation
        const size_t cpx = std::max(npx, h->sal_cap_px), cint = std::max(nint, h->sal_cap_int);
        h->sal_cap_px = 0; h->sal_cap_int = 0;
        CU_TRY(h, regrow(h->sal_g0, cpx * kSalChunk)); CU_TRY(h, regrow(h->sal_g1, cpx * kSalChunk));
        CU_TRY(h, regrow(h->sal_on, cpx * kSalChunk)); CU_TRY(h, regrow(h->sal_off, cpx * kSalChunk));
        CU_TRY(h, regrow(h->sal_prefix, cpx * kSalChunk)); CU_TRY(h, regrow(h->sal_integ, cint * kSalChunk));
        CU_TRY(h, regrow(h->sal_son, cpx * kSalChunk)); CU_TRY(h, regrow(h->sal_soff, cpx * kSalChunk));
        if (!h->sal_max) CU_TRY(h, regrow(h->sal_max, (size_t)kSalChunk));
        h->sal_cap_px = cpx; h->sal_cap_int = cint;
    }
    const dim3 blk(32, 8);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(kSalChunk, (size_t)INT_MAX / npx));   // 32-bit pixel index per chunk
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        const int nf = std::min(chunk, n_frames - f0);
        const int tot = (int)(npx * nf);
        const dim3 g2((W + 31) / 32, (H + 7) / 8, nf);
        const dim3 g1(std::max(1, std::min((int)((npx + 1023) / 1024), 256)), nf);
        CU_TRY(h, cudaMemsetAsync(h->sal_max, 0, sizeof(SalMax) * kSalChunk, stream));
        sal_gray_kernel<<<std::max(1, std::min((tot + 255) / 256, 4096)), 256, 0, stream>>>(rgb_dev + (size_t)f0 * npx * 3, tot, h->sal_g0);
        sal_blur5_kernel<<<g2, blk, 0, stream>>>(h->sal_g0, h->sal_g1, nf, H, W);
        sal_blur5_kernel<<<g2, blk, 0, stream>>>(h->sal_g1, h->sal_g0, nf, H, W);
        sal_rowprefix_kernel<<<(nf * H + 7) / 8, 256, 0, stream>>>(h->sal_g0, h->sal_prefix, nf * H, W);
        sal_integral_kernel<<<dim3((W + 1 + 127) / 128, nf), 128, 0, stream>>>(h->sal_prefix, h->sal_integ, nf, H, W);
        sal_scales_kernel<<<g2, blk, 0, stream>>>(h->sal_g0, h->sal_integ, nf, H, W, h->sal_son, h->sal_soff, h->sal_max);
        sal_mix_kernel<<<g1, 256, 0, stream>>>(h->sal_son, h->sal_soff, nf, (int)npx, h->sal_on, h->sal_off, h->sal_max);
        sal_final_kernel<<<g1, 256, 0, stream>>>(h->sal_on, h->sal_off, nf, (int)npx, h->sal_max,
                                                 saliency_dev ? saliency_dev + (size_t)f0 * npx : nullptr,
                                                 intensity_dev ? intensity_dev + (size_t)f0 * npx : nullptr);
        CU_TRY(h, cudaGetLastError());
    }
    return TEEFLOW_OK;
}

// ---- connected components: clean_mask (calculate_optical_flow.py:91-182) and calc_AV_centroid (analysis.py:39-86)
static const int kCclChunk = 32;

static int ccl_reserve(teeflow_engine* h, size_t npx) {
    const size_t need = npx * kCclChunk;
    if (need > h->ccl_cap) {
        h->ccl_cap = 0;
        CU_TRY(h, regrow(h->ccl_L, need)); CU_TRY(h, regrow(h->ccl_area, need)); CU_TRY(h, regrow(h->ccl_touch, need));
        CU_TRY(h, regrow(h->ccl_sr, need)); CU_TRY(h, regrow(h->ccl_sc, need));
        h->ccl_cap = need;
    }
    if (!h->ccl_best) { CU_TRY(h, regrow(h->ccl_best, (size_t)kCclChunk * 4)); CU_TRY(h, regrow(h->ccl_ncomp, (size_t)kCclChunk)); }
    return TEEFLOW_OK;
}

// labels + per-root area (and optionally coordinate sums / border contact) of `nf` frames starting at img
static int ccl_run(teeflow_engine* h, const uint8_t* img, int nf, int H, int W, int ch, int fg, int conn8, bool sums,
                   bool border, cudaStream_t stream) {
    const int npx = H * W;
    const dim3 grid(std::max(1, std::min((npx + 255) / 256, 256)), nf);
    ccl_init_kernel<<<grid, 256, 0, stream>>>(img, npx, ch, fg, h->ccl_L);
    ccl_merge_kernel<<<grid, 256, 0, stream>>>(h->ccl_L, H, W, conn8);
    CU_TRY(h, cudaMemsetAsync(h->ccl_area, 0, sizeof(int) * (size_t)nf * npx, stream));
    if (sums) {
        CU_TRY(h, cudaMemsetAsync(h->ccl_sr, 0, sizeof(unsigned long long) * (size_t)nf * npx, stream));
        CU_TRY(h, cudaMemsetAsync(h->ccl_sc, 0, sizeof(unsigned long long) * (size_t)nf * npx, stream));
    }
    if (border) CU_TRY(h, cudaMemsetAsync(h->ccl_touch, 0, sizeof(int) * (size_t)nf * npx, stream));
    ccl_stats_kernel<<<grid, 256, 0, stream>>>(h->ccl_L, H, W, h->ccl_area, sums ? h->ccl_sr : nullptr,
                                               sums ? h->ccl_sc : nullptr, border ? h->ccl_touch : nullptr);
    CU_TRY(h, cudaGetLastError());
    return TEEFLOW_OK;
}

int teeflow_clean_masks(teeflow_handle h, const uint8_t* classmap_dev, int n_frames, int H, int W, int class_id,
                        int window, double threshold, int min_size, uint8_t* mask_dev, void* stream_v) {
    if (!h || !classmap_dev || !mask_dev || n_frames < 1 || H < 1 || W < 1 || window < 1)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    const int npx = H * W;
    int rc = ccl_reserve(h, (size_t)npx);
    if (rc) return rc;
    const dim3 gall(std::max(1, std::min((npx + 255) / 256, 256)), n_frames);
    mask_vote_kernel<<<gall, 256, 0, stream>>>(classmap_dev, n_frames, npx, class_id, window, threshold, mask_dev);
    for (int f0 = 0; f0 < n_frames; f0 += kCclChunk) {
        const int nf = std::min(kCclChunk, n_frames - f0);
        uint8_t* m = mask_dev + (size_t)f0 * npx;
        const dim3 grid(gall.x, nf);
        // binary_fill_holes: 4-connected background components that do not reach the image border
        if ((rc = ccl_run(h, m, nf, H, W, 1, /*fg=*/0, /*conn8=*/0, false, true, stream))) return rc;
        fill_holes_kernel<<<grid, 256, 0, stream>>>(m, h->ccl_L, h->ccl_touch, npx);
        // remove_small_objects(min_size), connectivity 1
        if ((rc = ccl_run(h, m, nf, H, W, 1, /*fg=*/1, /*conn8=*/0, false, false, stream))) return rc;
        remove_small_kernel<<<grid, 256, 0, stream>>>(m, h->ccl_L, h->ccl_area, npx, min_size);
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(stream));
    return TEEFLOW_OK;
}

int teeflow_av_centroids(teeflow_handle h, const uint8_t* mask_dev, int channels, int nframes, int H, int W,
                         double* centroids_host, int32_t* n_components_host, void* stream_v) {
    if (!h || !mask_dev || !centroids_host || nframes < 1 || H < 1 || W < 1 || channels < 1)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    const int npx = H * W;
    int rc = ccl_reserve(h, (size_t)npx);
    if (rc) return rc;
    std::vector<unsigned long long> best((size_t)kCclChunk * 3);
    std::vector<int> ncomp(kCclChunk);
    for (int f0 = 0; f0 < nframes; f0 += kCclChunk) {
        const int nf = std::min(kCclChunk, nframes - f0);
        const uint8_t* m = mask_dev + (size_t)f0 * npx * channels;
        // skimage.measure.label default connectivity for 2-D = 2 (8-connected); channel 0 of the mask (analysis.py:59)
        if ((rc = ccl_run(h, m, nf, H, W, channels, 1, 1, true, false, stream))) return rc;
        CU_TRY(h, cudaMemsetAsync(h->ccl_best, 0, sizeof(unsigned long long) * nf, stream));
        CU_TRY(h, cudaMemsetAsync(h->ccl_ncomp, 0, sizeof(int) * nf, stream));
        const dim3 grid(std::max(1, std::min((npx + 255) / 256, 256)), nf);
        largest_component_kernel<<<grid, 256, 0, stream>>>(h->ccl_L, h->ccl_area, npx, h->ccl_best, h->ccl_ncomp);
        CU_TRY(h, cudaGetLastError());
        largest_component_gather_kernel<<<1, kCclChunk, 0, stream>>>(h->ccl_best, h->ccl_ncomp, h->ccl_sr, h->ccl_sc, npx,
                                                                     nf, h->ccl_best + kCclChunk);
        CU_TRY(h, cudaGetLastError());
        CU_TRY(h, cudaMemcpyAsync(best.data(), h->ccl_best + kCclChunk, sizeof(unsigned long long) * 3 * nf,
                                  cudaMemcpyDeviceToHost, stream));
        CU_TRY(h, cudaMemcpyAsync(ncomp.data(), h->ccl_ncomp, sizeof(int) * nf, cudaMemcpyDeviceToHost, stream));
        CU_TRY(h, cudaStreamSynchronize(stream));
        for (int f = 0; f < nf; ++f) {
            if (n_components_host) n_components_host[f0 + f] = ncomp[f];
            if (ncomp[f] == 0) { centroids_host[2 * (f0 + f)] = centroids_host[2 * (f0 + f) + 1] = std::nan(""); continue; }
            const double area = (double)best[3 * f];
            centroids_host[2 * (f0 + f)] = (double)best[3 * f + 1] / area;      // regionprops.centroid: mean (row, col)
            centroids_host[2 * (f0 + f) + 1] = (double)best[3 * f + 2] / area;
        }
    }
    return TEEFLOW_OK;
}

// ---- masked radial / longitudinal decomposition + per-frame reductions (analysis.py, cardiac_cycle_detection.py)
static float pct_f32(const float a, const float b, int n, double q) {
    // np.percentile on a float32 array: the whole computation runs in float32 (numpy >= 2)
    if (n <= 1) return a;
    const float qf = (float)q / 100.0f;
    const float virt = (float)(n - 1) * qf;
    const float g = virt - floorf(virt);
    const float d = b - a;
    return g >= 0.5f ? b - d * (1.0f - g) : a + d * g;
}
static double pct_f64(const double a, const double b, long long n, double q) {
    if (n <= 1) return a;
    const double virt = (double)(n - 1) * (q / 100.0);
    const double g = virt - floor(virt);
    const double d = b - a;
    return g >= 0.5 ? b - d * (1.0 - g) : a + d * g;
}
static void pct_ranks(long long n, double q, bool f32, long long* lo, long long* hi) {
    if (n <= 0) { *lo = *hi = -1; return; }
    double prev;
    if (f32) { const float virt = (float)(n - 1) * ((float)q / 100.0f); prev = floorf(virt); }
    else { const double virt = (double)(n - 1) * (q / 100.0); prev = floor(virt); }
    long long l = (long long)prev;
    if (l >= n - 1) { *lo = *hi = n - 1; return; }   // numpy: index above bounds -> last element twice
    if (l < 0) l = 0;
    *lo = l; *hi = l + 1;
}

static float host_key_f32(unsigned long long k64) {
    unsigned k = (unsigned)k64; unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; float v; memcpy(&v, &b, 4); return v;
}
static double host_key_f64(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k; double v; memcpy(&v, &b, 8); return v;
}

int teeflow_analyze_clip(teeflow_handle h, const void* flow_f16_dev, const uint8_t* mask_dev,
                         const double* centroids_host, int nframes, int H, int W, double perc_lo, double perc_hi,
                         teeflow_analysis* out, void* stream_v) {
    if (!h || !flow_f16_dev || !mask_dev || !centroids_host || !out) return fail(h, TEEFLOW_ERR_BAD_ARG, "NULL argument");
    if (nframes < 1 || H < 1 || W < 1) return fail(h, TEEFLOW_ERR_BAD_SHAPE, "bad shape");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    const size_t npx = (size_t)H * W, need = npx * nframes;
    if (need > h->an_cap) {
        h->an_cap = 0;
        CU_TRY(h, regrow(h->an_mag, need)); CU_TRY(h, regrow(h->an_ang, need));
        CU_TRY(h, regrow(h->an_rad, need)); CU_TRY(h, regrow(h->an_long, need));
        h->an_cap = need;
    }
    if ((size_t)nframes > h->an_frames_cap) {
        h->an_frames_cap = 0;
        CU_TRY(h, regrow(h->an_stats, (size_t)nframes)); CU_TRY(h, regrow(h->an_anghist, (size_t)nframes * kAngBins));
        CU_TRY(h, regrow(h->an_cent, (size_t)nframes * 2)); CU_TRY(h, regrow(h->an_ranks, (size_t)nframes * 12));
        CU_TRY(h, regrow(h->an_keys, (size_t)nframes * 12));
        h->an_frames_cap = (size_t)nframes;
    }
    h->an_frames = nframes; h->an_H = H; h->an_W = W;
    CU_TRY(h, cudaMemcpyAsync(h->an_cent, centroids_host, sizeof(double) * 2 * nframes, cudaMemcpyHostToDevice, stream));
    analysis_init_kernel<<<(nframes * kAngBins + 255) / 256, 256, 0, stream>>>(h->an_stats, h->an_anghist, nframes);
    const int chunks = std::max(1, std::min((int)((npx + 256 * 8 - 1) / (256 * 8)), 64));
    analysis_values_kernel<<<dim3(chunks, nframes), 256, 0, stream>>>(
        (const __half2*)flow_f16_dev, mask_dev, h->an_cent, H, W, h->an_mag, h->an_ang, h->an_rad, h->an_long,
        h->an_stats, h->an_anghist);
    CU_TRY(h, cudaGetLastError());
    std::vector<FrameStats> st(nframes);
    std::vector<unsigned> ah((size_t)nframes * kAngBins);
    CU_TRY(h, cudaMemcpyAsync(st.data(), h->an_stats, sizeof(FrameStats) * nframes, cudaMemcpyDeviceToHost, stream));
    CU_TRY(h, cudaMemcpyAsync(ah.data(), h->an_anghist, sizeof(unsigned) * ah.size(), cudaMemcpyDeviceToHost, stream));
    CU_TRY(h, cudaStreamSynchronize(stream));

    // global ranges (np.min / np.max over the whole array, zeros included)
    unsigned mmin = 0xffffffffu, mmax = 0, amin = 0xffffffffu, amax = 0;
    unsigned long long rmin = ~0ull, rmax = 0, lmin = ~0ull, lmax = 0;
    for (int f = 0; f < nframes; ++f) {
        mmin = std::min(mmin, st[f].mag_min); mmax = std::max(mmax, st[f].mag_max);
        amin = std::min(amin, st[f].ang_min); amax = std::max(amax, st[f].ang_max);
        rmin = std::min(rmin, st[f].rad_min); rmax = std::max(rmax, st[f].rad_max);
        lmin = std::min(lmin, st[f].long_min); lmax = std::max(lmax, st[f].long_max);
    }
    out->mag_min = host_key_f32(mmin); out->mag_max = host_key_f32(mmax);
    out->ang_min = host_key_f32(amin); out->ang_max = host_key_f32(amax);
    out->rad_min = host_key_f64(rmin); out->rad_max = host_key_f64(rmax);
    out->long_min = host_key_f64(lmin); out->long_max = host_key_f64(lmax);

    // target ranks per frame: [q][4] = mag: (hi.lo, hi.hi, -, -); rad/long: (lo.lo, lo.hi, hi.lo, hi.hi)
    std::vector<long long> ranks((size_t)nframes * 12, -1);
    for (int f = 0; f < nframes; ++f) {
        long long* r = ranks.data() + (size_t)f * 12;
        pct_ranks((long long)st[f].cnt[0], perc_hi, true, &r[0], &r[1]);
        pct_ranks((long long)st[f].cnt[2], perc_lo, false, &r[4], &r[5]);
        pct_ranks((long long)st[f].cnt[2], perc_hi, false, &r[6], &r[7]);
        pct_ranks((long long)st[f].cnt[3], perc_lo, false, &r[8], &r[9]);
        pct_ranks((long long)st[f].cnt[3], perc_hi, false, &r[10], &r[11]);
    }
    CU_TRY(h, cudaMemcpyAsync(h->an_ranks, ranks.data(), sizeof(long long) * ranks.size(), cudaMemcpyHostToDevice, stream));
    radix_select_kernel<<<dim3(nframes, 3), 1024, 0, stream>>>(h->an_mag, h->an_rad, h->an_long, (int)npx, h->an_stats, h->an_ranks, h->an_keys);
    CU_TRY(h, cudaGetLastError());
    std::vector<unsigned long long> keys((size_t)nframes * 12);
    CU_TRY(h, cudaMemcpyAsync(keys.data(), h->an_keys, sizeof(unsigned long long) * keys.size(), cudaMemcpyDeviceToHost, stream));
    CU_TRY(h, cudaStreamSynchronize(stream));

    const double nan = std::nan("");
    for (int f = 0; f < nframes; ++f) {
        const unsigned long long* k = keys.data() + (size_t)f * 12;
        const long long nm = (long long)st[f].cnt[0], nr = (long long)st[f].cnt[2], nl = (long long)st[f].cnt[3];
        if (out->counts) for (int q = 0; q < 4; ++q) out->counts[(size_t)f * 4 + q] = (int64_t)st[f].cnt[q];
        if (out->mag_hi) out->mag_hi[f] = nm > 0 ? pct_f32(host_key_f32(k[0]), host_key_f32(k[1]), (int)nm, perc_hi) : (float)nan;
        if (out->rad_lo) out->rad_lo[f] = nr > 0 ? pct_f64(host_key_f64(k[4]), host_key_f64(k[5]), nr, perc_lo) : nan;
        if (out->rad_hi) out->rad_hi[f] = nr > 0 ? pct_f64(host_key_f64(k[6]), host_key_f64(k[7]), nr, perc_hi) : nan;
        if (out->long_lo) out->long_lo[f] = nl > 0 ? pct_f64(host_key_f64(k[8]), host_key_f64(k[9]), nl, perc_lo) : nan;
        if (out->long_hi) out->long_hi[f] = nl > 0 ? pct_f64(host_key_f64(k[10]), host_key_f64(k[11]), nl, perc_hi) : nan;
        if (out->ang_mode) {
            // scipy.stats.mode of np.round(ang, 2) over the non-zero entries: smallest value among ties
            const unsigned* hst = ah.data() + (size_t)f * kAngBins;
            unsigned best = 0; int bk = -1;
            for (int b = 1; b < kAngBins; ++b) if (hst[b] > best) { best = hst[b]; bk = b; }
            out->ang_mode[f] = bk < 0 ? (float)nan : (float)bk / 100.0f;
        }
    }
    return TEEFLOW_OK;
}

// np.histogram of the non-zero entries of one analysed quantity (0 mag, 1 ang: float32 edges; 2 rad, 3 long:
// float64 edges) with caller-provided numpy edges (np.linspace(first, last, nbins + 1)); freq_host[nframes][nbins]
int teeflow_analysis_histogram(teeflow_handle h, int quantity, const void* edges_host, int nbins, int64_t* freq_host,
                               void* stream_v) {
    if (!h || !edges_host || !freq_host || quantity < 0 || quantity > 3 || nbins < 1 || nbins > 8192)
        return fail(h, TEEFLOW_ERR_BAD_ARG, "bad argument");
    if (h->an_frames < 1) return fail(h, TEEFLOW_ERR_STATE, "teeflow_analyze_clip has not been called");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CU_TRY(h, cudaSetDevice(h->device));
    const int nframes = h->an_frames;
    const size_t npx = (size_t)h->an_H * h->an_W;
    const bool f64 = quantity >= 2;
    const size_t esz = f64 ? 8 : 4;
    if (!h->an_edges) CU_TRY(h, cudaMalloc(&h->an_edges, 8 * 8193));
    if ((size_t)nframes * nbins > h->an_freq_cap) {
        h->an_freq_cap = 0;
        CU_TRY(h, regrow(h->an_freq, (size_t)nframes * nbins));
        h->an_freq_cap = (size_t)nframes * nbins;
    }
    CU_TRY(h, cudaMemcpyAsync(h->an_edges, edges_host, esz * (nbins + 1), cudaMemcpyHostToDevice, stream));
    CU_TRY(h, cudaMemsetAsync(h->an_freq, 0, sizeof(unsigned long long) * nframes * nbins, stream));
    const int chunks = std::max(1, std::min((int)((npx + 256 * 8 - 1) / (256 * 8)), 64));
    const dim3 grid(chunks, nframes);
    const size_t smem = sizeof(unsigned) * nbins;
    if (!f64) {
        const float* e = (const float*)edges_host;
        np_histogram_kernel<float><<<grid, 256, smem, stream>>>(quantity == 0 ? h->an_mag : h->an_ang, (int)npx, h->an_stats, quantity,
                                                                (const float*)h->an_edges, nbins, e[0], e[nbins], h->an_freq);
    } else {
        const double* e = (const double*)edges_host;
        np_histogram_kernel<double><<<grid, 256, smem, stream>>>(quantity == 2 ? h->an_rad : h->an_long, (int)npx, h->an_stats, quantity,
                                                                 (const double*)h->an_edges, nbins, e[0], e[nbins], h->an_freq);
    }
    CU_TRY(h, cudaGetLastError());
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "freq layout");
    CU_TRY(h, cudaMemcpyAsync(freq_host, h->an_freq, sizeof(int64_t) * nframes * nbins, cudaMemcpyDeviceToHost, stream));
    CU_TRY(h, cudaStreamSynchronize(stream));
    return TEEFLOW_OK;
}

}  // extern "C"
