// tvl1_kernels.cuh -- the kernels of the TV-L1 engine.
//
//  * pyramid kernels (once per FRAME, not per pair): u8/f32 -> f32, resize x0.8, centred gradient + pack.
//  * tvl1_step_kernel: one persistent launch = one "super-step" of the device-side scheduler.  Every slot
//    (a frame pair in flight) is in some phase (level-init / warp / median / inner / final); the work of all
//    slots is flattened into warp-sized strip items, persistent warps grid-stride over them, and the warp that
//    finishes the last strip of a slot reduces the slot's error partials (fixed order, float64) and advances
//    its state machine exactly like OpenCV's procOneScale control flow.  No host round trip
//    per iteration, per-pair exact early exit, continuous refill of finished slots.
//  * The inner iteration is a warp-autonomous register-rolling stencil: a warp owns 31 output columns (+1 halo
//    column) and walks down 32 rows; vertical neighbours stay in registers, horizontal ones come by warp
//    shuffle, loads run one row ahead.  No shared memory, no block barrier.
#pragma once
#include "tvl1_device.cuh"

namespace teeflow {

#ifndef TEEFLOW_THREADS
#define TEEFLOW_THREADS 128
#endif
constexpr int kThreads = TEEFLOW_THREADS;  // threads per CTA
constexpr int kWarpsPerCta = kThreads / 32;
#ifndef TEEFLOW_IW
#define TEEFLOW_IW 31
#endif
constexpr int kIW = TEEFLOW_IW;        // inner strip: output columns per warp (lane 31 = right halo column)
#ifndef TEEFLOW_STRIP_ROWS
#define TEEFLOW_STRIP_ROWS 32
#endif
constexpr int kIR = TEEFLOW_STRIP_ROWS;   // inner strip: rows per warp (measured 16: 1215, 20: 1245, 24: 1241, 32: 1233, 40: 1200 pairs/s)
constexpr int kIW2 = 29;       // two-iteration strip: output columns per warp (lanes 1..29; lane 0 / 30 / 31 halo columns)
#ifndef TEEFLOW_STRIP2_ROWS
#define TEEFLOW_STRIP2_ROWS 32
#endif
constexpr int kIR2 = TEEFLOW_STRIP2_ROWS;   // two-iteration strip: rows per warp (3 extra row steps per strip)
#ifndef TEEFLOW_POINT_ROWS
#define TEEFLOW_POINT_ROWS 16
#endif
constexpr int kPR = TEEFLOW_POINT_ROWS;   // pointwise strip: rows per warp, 32 columns (measured 8: 1243, 12: 1250, 16: 1264, 24: 1244 pairs/s)
#ifndef TEEFLOW_INNER_UNROLL
#define TEEFLOW_INNER_UNROLL 1
#endif
constexpr int kInnerUnroll = TEEFLOW_INNER_UNROLL;   // rows per trip of the single-iteration loop (1 halves its code)
#ifndef TEEFLOW_WARP_PF
#define TEEFLOW_WARP_PF 0
#endif
constexpr int kWarpPF = TEEFLOW_WARP_PF;   // warp op: tap rows pulled into L2 ahead of the gather window (0: off)
#ifndef TEEFLOW_WARP_PF1
#define TEEFLOW_WARP_PF1 2
#endif
constexpr int kWarpPF1 = TEEFLOW_WARP_PF1; // warp op: tap rows that enter the window kWarpPF1 pixel rows from now are pulled into L1 (0: off)
#ifndef TEEFLOW_LATE_HANDOVER
#define TEEFLOW_LATE_HANDOVER 1
#endif
#ifndef TEEFLOW_PF_ROWS
#define TEEFLOW_PF_ROWS 4
#endif
constexpr int kPF = TEEFLOW_PF_ROWS;   // inner iteration: rows ahead of the current one that are pulled into L2 (0: off)
#ifndef TEEFLOW_DYNAMIC_ITEMS
#define TEEFLOW_DYNAMIC_ITEMS 1
#endif
// 1: the CB plane holds only rho_c, two image rows per float2 -- element x of the CB row of an EVEN image row y is
// (rho_c(x, y), rho_c(x, y + 1)) -- and the inner iteration recomputes grad = I1wx^2 + I1wy^2 from the CA plane (the
// same two products and one sum the warp op rounded): 60 instead of 64 bytes per pixel and iteration.
// 0: CB = (grad, rho_c) per pixel.
#ifndef TEEFLOW_RHO_PACK
#define TEEFLOW_RHO_PACK 1
#endif
constexpr bool kRhoPack = TEEFLOW_RHO_PACK != 0;
static_assert(!kRhoPack || (kIR % 2 == 0 && kIR2 % 2 == 0 && kPR % 2 == 0), "row-pair packing of rho_c needs even strip heights");

// ---- TMA staging of the inner iteration's input rows (op_inner_tma)
// 0 (shipped default): the inner iteration reads its rows through registers one row ahead + L2 prefetches (op_inner).
// 1 (libteeflow_tma.so, built and parity-tested next to the default): TMA-staged rows (op_inner_tma); bit-identical,
// measured 10-12 % slower on the clip in every ring geometry (profiles/r2_summary.md).
#ifndef TEEFLOW_TMA_INNER
#define TEEFLOW_TMA_INNER 0
#endif
#ifndef TEEFLOW_TMA_ROWS
#define TEEFLOW_TMA_ROWS 2
#endif
#ifndef TEEFLOW_TMA_STAGES
#define TEEFLOW_TMA_STAGES 2
#endif
constexpr bool kTma = TEEFLOW_TMA_INNER != 0;
constexpr int kTR = TEEFLOW_TMA_ROWS;      // image rows per box (one chunk of a strip)
constexpr int kTS = TEEFLOW_TMA_STAGES;    // chunks in flight per warp (ring depth)
static_assert(!kTma || (kRhoPack && kTR % 2 == 0 && kTR >= 2 && kTS >= 2 && kTS <= 8), "TMA ring geometry");
// The first coordinate of a box must be a multiple of 16 bytes = two 8-byte elements (an odd one faults: measured,
// tools/tma_probe2.cu), and strips start every 31 columns: every box is 34 columns wide and starts on the even column
// at or below the first column it needs; the lanes read at an offset of 0 or 1 elements.
constexpr int kTBW = 34;
constexpr unsigned round128(unsigned v) { return (v + 127u) & ~127u; }
constexpr unsigned kTRowB = (unsigned)kTBW * 8u;            // bytes of one staged plane row
constexpr unsigned kTOffU = 0u;
constexpr unsigned kTOffCA = round128(kTOffU + kTR * kTRowB);
constexpr unsigned kTOffP = round128(kTOffCA + kTR * kTRowB);    // box layout: [row][PX, PY][34 columns]
constexpr unsigned kTOffCB = round128(kTOffP + kTR * 2u * kTRowB);
constexpr unsigned kTStageB = round128(kTOffCB + (kTR / 2) * kTRowB);
constexpr unsigned kTTxBytes = 2u * kTR * kTRowB + kTR * 2u * kTRowB + (kTR / 2) * kTRowB;   // bytes one chunk delivers
constexpr unsigned kTWarpB = kTS * kTStageB;

struct TmaRing {              // one warp's staging ring (shared memory)
    uint32_t ring;            // shared-memory address of its kTS stages
    uint32_t mbar;            // ... of its kTS mbarriers (8 bytes each)
    unsigned* phase;          // parity bits of the mbarriers: persist from strip to strip
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// One chunk of a strip: arm the stage's mbarrier and ask for its four boxes.  Executed by the WHOLE warp with identical
// operands; one elected lane issues (in a divergent `if (lane == 0)` ptxas wraps every UTMALDG in a waterfall loop
// of eight R2UR broadcasts -- 78 instructions per chunk; here the warp-uniform operands are moved once).
__device__ __forceinline__ void tma_issue_chunk(uint32_t dst, uint32_t bar, const CUtensorMap* tm, int xe, int xp, int plane_u,
                                                int plane_px, int row, int slot) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 d1, d2, d3, r2;\n\t"
        ".reg .b64 t1, t2;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "add.u32 d1, %0, %9;\n\t"
        "add.u32 d2, %0, %10;\n\t"
        "add.u32 d3, %0, %11;\n\t"
        "add.u64 t1, %2, 128;\n\t"
        "add.u64 t2, %2, 256;\n\t"
        "shr.s32 r2, %7, 1;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %12;\n\t"
        "@p cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%2, {%3, %5, %7, %8}], [%1];\n\t"
        "@p cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [d1], [%2, {%3, %13, %7, %8}], [%1];\n\t"
        "@p cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [d2], [t1, {%4, %6, %7, %8}], [%1];\n\t"
        "@p cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [d3], [t2, {%3, %14, r2, %8}], [%1];\n\t"
        "}"
        ::"r"(dst), "r"(bar), "l"(tm), "r"(xe), "r"(xp), "r"(plane_u), "r"(plane_px), "r"(row), "r"(slot), "n"(kTOffCA),
          "n"(kTOffP), "n"(kTOffCB), "n"(kTTxBytes), "n"((int)PL_CA), "n"(0)
        : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------ pyramid
// level 0: convertTo(CV_32F, 1) for u8, x255 for f32 (tvl1flow.cpp: I0mult / I1mult)
__global__ void pyr_level0_kernel(const void* __restrict__ frames, int dtype, long long frame_stride, int n_frames,
                                  int npx, float* __restrict__ pyrI, long long pyr_stride) {
    const int f = blockIdx.y;
    if (f >= n_frames) return;
    float* dst = pyrI + (size_t)f * pyr_stride;
    if (dtype == 0) {
        const uint8_t* src = (const uint8_t*)frames + (size_t)f * frame_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
            dst[i] = (float)src[i];
    } else {
        const float* src = (const float*)frames + (size_t)f * frame_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
            dst[i] = src[i] * 255.0f;
    }
}

// resize(src, dst, Size(), scaleStep, scaleStep, INTER_LINEAR): dsize = cvRound(size*f), scale = 1/f
__global__ void pyr_down_kernel(float* __restrict__ pyrI, long long pyr_stride, int n_frames, long long src_off,
                                int sH, int sW, long long dst_off, int dH, int dW, double scale) {
    const int f = blockIdx.z;
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (f >= n_frames || dx >= dW || dy >= dH) return;
    const float* S = pyrI + (size_t)f * pyr_stride + src_off;
    float* D = pyrI + (size_t)f * pyr_stride + dst_off;
    int x0, x1, y0, y1; float a0, a1, b0, b1;
    lin_coeff_x(dx, scale, sW, x0, x1, a0, a1);
    lin_coeff_y(dy, scale, sH, y0, y1, b0, b1);
    const float r0 = S[(size_t)y0 * sW + x0] * a0 + S[(size_t)y0 * sW + x1] * a1;
    const float r1 = S[(size_t)y1 * sW + x0] * a0 + S[(size_t)y1 * sW + x1] * a1;
    D[(size_t)dy * dW + dx] = r0 * b0 + r1 * b1;
}

// centeredGradient + pack (I, Ix, Iy, 0) so that the bicubic gather needs one 16-byte load per tap
__global__ void pyr_pack_kernel(const float* __restrict__ pyrI, float4* __restrict__ pyrG, long long pyr_stride,
                                int n_frames, long long off, int H, int W) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (f >= n_frames || x >= W || y >= H) return;
    const float* I = pyrI + (size_t)f * pyr_stride + off;
    const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
    float4 g;
    g.x = I[(size_t)y * W + x];
    g.y = 0.5f * (I[(size_t)y * W + xp] - I[(size_t)y * W + xm]);
    g.z = 0.5f * (I[(size_t)yp * W + x] - I[(size_t)ym * W + x]);
    g.w = 0.f;
    pyrG[(size_t)f * pyr_stride + off + (size_t)y * W + x] = g;
}

// ------------------------------------------------------------------------------------------- state machine
// Control flow of OpticalFlowDual_TVL1::procOneScale / ::calc, one transition per finished phase.
// Two iterations per pass (PH_INNER2) are speculative: OpenCV tests the exit condition after EVERY iteration.  The
// pass returns both error sums; if the first iteration already ends the loop, the pass is discarded (its inputs
// are intact: it wrote the ping-pong partners) and that iteration is redone alone -- results never depend on the
// choice, only the time does.  A pass is tried from the second iteration of a loop on, while the last error is
// above spec_factor x the exit threshold (the error shrinks by less than that per iteration).
__device__ inline void advance_slot(const EngineParams& P, Slot& s, double err_sum, double err_sum2) {
    const float eps = P.lv[s.level].scaled_eps;
    enum { NONE, CHECK_INNER, CHECK_OUTER, NEXT_WARP } todo = NONE;
    switch (s.phase) {
        case PH_LEVEL_INIT:
            // the coarsest level wrote U[0]; finer levels up-sampled U[ucur] into U[ucur ^ 1]; p lives in P[0]
            if (s.level != P.L - 1) s.ucur ^= 1;
            s.pcur = 0; s.warp = 0; s.phase = PH_WARP; return;
        case PH_WARP:
            s.cnt[s.level][2]++;
            s.error = FLT_MAX; s.n_outer = 0; todo = CHECK_OUTER; break;
        case PH_MEDIAN:
            s.cnt[s.level][1]++;
            s.ucur ^= 1; s.n_inner = 0; todo = CHECK_INNER; break;
        case PH_INNER:
            s.cnt[s.level][0]++;
            s.ucur ^= 1; s.pcur ^= 1;
            s.error = (float)err_sum; s.n_inner++; todo = CHECK_INNER; break;
        case PH_INNER2: {
            const float e1 = (float)err_sum;
            if (!(e1 > eps && s.n_inner + 1 < P.inner)) {        // the loop ends after the first iteration
                atomicAdd(P.spec_stats + 1, 1);
                s.force_single = 1; s.phase = PH_INNER; return;
            }
            atomicAdd(P.spec_stats, 1);
            s.cnt[s.level][0] += 2;
            s.ucur ^= 1; s.pcur ^= 1;                            // the results sit in the partner planes
            s.error = (float)err_sum2; s.n_inner += 2; todo = CHECK_INNER; break;
        }
        default: return;
    }
    for (;;) {
        if (todo == CHECK_INNER) {
            if (s.error > eps && s.n_inner < P.inner) {
                const bool two = !s.force_single && s.n_inner >= 1 && s.n_inner + 2 <= P.inner &&
                                 P.spec_factor > 0.f && s.error > P.spec_factor * eps;
                s.force_single = 0;
                s.phase = two ? PH_INNER2 : PH_INNER; return;
            }
            s.n_outer++; todo = CHECK_OUTER;
        }
        if (todo == CHECK_OUTER) {
            if (s.error > eps && s.n_outer < P.outer) {
                if (P.median > 1) { s.phase = PH_MEDIAN; return; }
                s.n_inner = 0; todo = CHECK_INNER; continue;
            }
            todo = NEXT_WARP;
        }
        if (todo == NEXT_WARP) {
            s.warp++;
            if (s.warp < P.warps) { s.phase = PH_WARP; return; }
            if (s.level == 0) { s.phase = P.wase_w ? PH_WASE : PH_FINAL; return; }
            s.level--; s.phase = PH_LEVEL_INIT; return;
        }
    }
}

__device__ inline void start_pair(const EngineParams& P, Slot& s, int pair) {
    s.pair = pair; s.phase = PH_LEVEL_INIT; s.level = P.L - 1; s.warp = 0; s.n_outer = 0; s.n_inner = 0;
    s.ucur = 0; s.pcur = 0; s.error = FLT_MAX; s.bg = 0.f; s.force_single = 0;
    for (int l = 0; l < kMaxLevels; ++l) { s.cnt[l][0] = 0; s.cnt[l][1] = 0; s.cnt[l][2] = 0; }
}

__device__ __forceinline__ int items_of(const EngineParams& P, int phase, int pair, int level) {
    if (phase == PH_IDLE || pair < 0) return 0;
    if (phase == PH_FINAL || phase == PH_WASE) return P.lv[0].pw_items;
    if (phase == PH_INNER2) return P.lv[level].in2_items;
    return phase == PH_INNER ? P.lv[level].in_items : P.lv[level].pw_items;
}

#ifndef TEEFLOW_CG_NEIGHBOURS
#define TEEFLOW_CG_NEIGHBOURS 1   // 1: median / level-init read their taps with ld.global.cg too and take no acquire fence
#endif
constexpr bool kCgNeighbours = TEEFLOW_CG_NEIGHBOURS != 0;
__device__ __forceinline__ bool needs_l1_acquire(int phase) {
    return !kCgNeighbours && (phase == PH_MEDIAN || phase == PH_LEVEL_INIT);
}
template <typename T>
__device__ __forceinline__ T ld_tap(const T* p) { return kCgNeighbours ? __ldcg(p) : *p; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------------ strip ops
// Slot planes live in the row-interleaved constant-pitch layout of struct Lay (tvl1_device.cuh).
__device__ __forceinline__ float2* slot_base(const EngineParams& P, int slot) {
    return P.planes + (size_t)slot * (size_t)P.slot_stride;
}

// PH_LEVEL_INIT: u = 0 (coarsest) or u = resize(u_coarse, INTER_LINEAR) * (1/scaleStep); p = 0
template <int PITCH>
__device__ __forceinline__ void op_level_init(const EngineParams& P, int level, int ucur, int slot, int strip, int lane) {
    using L = Lay<PITCH>;
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const bool coarsest = (level == P.L - 1);
    const unsigned pUd = PL_U + (coarsest ? 0u : (unsigned)(ucur ^ 1));
    const unsigned pUs = PL_U + (unsigned)ucur;
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    int x0 = 0, x1 = 0; float a0 = 0.f, a1 = 0.f;
    int cH = 0;
    if (!coarsest) {
        cH = P.lv[level + 1].H;
        lin_coeff_x(x, g.up_sx, P.lv[level + 1].W, x0, x1, a0, a1);
    }
    for (int y = y0; y < y1; ++y) {
        float2 u = make_float2(0.f, 0.f);
        if (!coarsest) {
            int ya, yb; float b0, b1;
            lin_coeff_y(y, g.up_sy, cH, ya, yb, b0, b1);
            const float2 s00 = ld_tap(SB + L::at(pUs, ya, x0)), s01 = ld_tap(SB + L::at(pUs, ya, x1));
            const float2 s10 = ld_tap(SB + L::at(pUs, yb, x0)), s11 = ld_tap(SB + L::at(pUs, yb, x1));
            const float r0x = s00.x * a0 + s01.x * a1, r1x = s10.x * a0 + s11.x * a1;
            const float r0y = s00.y * a0 + s01.y * a1, r1y = s10.y * a0 + s11.y * a1;
            u.x = (r0x * b0 + r1x * b1) * P.up_mul;
            u.y = (r0y * b0 + r1y * b1) * P.up_mul;
        }
        SB[L::at(pUd, y, x)] = u;
        SB[L::at(PL_PX, y, x)] = make_float2(0.f, 0.f);
        SB[L::at(PL_PY, y, x)] = make_float2(0.f, 0.f);
    }
}

// PH_WARP: buildFlowMap + remap(I1, I1x, I1y; INTER_CUBIC) + calcGradRho
template <int PITCH>
__device__ __forceinline__ void op_warp(const EngineParams& P, int level, int ucur, int pair, int slot, int strip,
                                        int lane, const float4* s_cubic) {
    using L = Lay<PITCH>;
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const int fa = P.pair_a[pair], fb = P.pair_b[pair];
    const float* I0 = P.pyrI + (size_t)fa * P.frame_pyr_stride + g.pyr_off;
    const float4* G1 = P.pyrG + (size_t)fb * P.frame_pyr_stride + g.pyr_off;
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    float2* row = SB + L::at(0, y0, x);          // plane 0 of row y0; planes / rows at constant offsets
    const unsigned oU = (PL_U + (unsigned)ucur) * (unsigned)PITCH;
    // the flow / I0 of the next row are fetched while the current row's gather runs (one row ahead)
    float2 u_n = __ldcg(row + oU);
    float i0_n = __ldg(I0 + (unsigned)(y0 * g.W + x));
    float rho_even = 0.f;                        // kRhoPack: rho_c of the even row above
    for (int y = y0; y < y1; ++y) {
        const unsigned q = (unsigned)(y * g.W + x);
        const float2 u = u_n;
        const float i0 = i0_n;
        if (y + 1 < y1) { u_n = __ldcg(row + oU + L::ROW); i0_n = __ldg(I0 + q + (unsigned)g.W); }
        const float mx = (float)x + u.x, my = (float)y + u.y;
        if (kWarpPF > 0) {
            // the flow is smooth: the window of the next pixel rows sits (almost) straight below this one, so the tap
            // row that enters it kWarpPF rows from now is pulled into L2 already (4 taps x 16 bytes per lane)
            const int pr = min(max((int)my + 2 + kWarpPF, 0), g.H - 1), pcx = min(max((int)mx - 1, 0), g.W - 1);
            prefetch_l2(G1 + (unsigned)(pr * g.W + pcx));
        }
        if (kWarpPF1 > 0) {
            // consecutive pixel rows share three of their four tap rows; the one that enters (four taps = 64 bytes per
            // lane) is requested ahead, so that its gather finds it in L1 instead of waiting for L2
            const int pr = min(max((int)my + 2 + kWarpPF1, 0), g.H - 1), pcx = min(max((int)mx - 1, 0), max(g.W - 4, 0));
            prefetch_l1(G1 + (unsigned)(pr * g.W + pcx));
        }
        const float3 w = remap_cubic3(G1, g.H, g.W, mx, my, s_cubic, P.negzero);
        const float Ix2 = w.y * w.y, Iy2 = w.z * w.z;
        row[PL_CA * PITCH] = make_float2(w.y, w.z);
        const float rho_c = w.x - w.y * u.x - w.z * u.y - i0;
        if (kRhoPack) {                          // strips start on even rows: the pair (y - 1, y) is written from row y
            if (y & 1) row[(int)(PL_CB * PITCH) - (int)L::ROW] = make_float2(rho_even, rho_c);
            else if (y + 1 == g.H) row[PL_CB * PITCH] = make_float2(rho_c, 0.f);
            rho_even = rho_c;
        } else {
            row[PL_CB * PITCH] = make_float2(Ix2 + Iy2, rho_c);
        }
        row += L::ROW;
    }
}

// PH_MEDIAN: medianBlur(u1, ksize), medianBlur(u2, ksize) with BORDER_REPLICATE
template <int PITCH>
__device__ __forceinline__ void op_median(const EngineParams& P, int level, int ucur, int slot, int strip, int lane) {
    using L = Lay<PITCH>;
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const unsigned pUs = PL_U + (unsigned)ucur, pUd = PL_U + (unsigned)(ucur ^ 1);
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    if (P.median == 5) {
        // sorted-row walk (tvl1_device.cuh): two output rows per step, one flow channel at a time (the rolling
        // state of a channel is 25 registers)
        const float* Uf = reinterpret_cast<const float*>(SB);
        float* Uw = reinterpret_cast<float*>(SB);
        unsigned xs[5];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) xs[dx] = 2u * (unsigned)clampi(x + dx - 2, 0, g.W - 1);
        const bool interior = x >= 2 && x + 2 < g.W;
#pragma unroll 1
        for (unsigned ch = 0; ch < 2; ++ch) {
            auto load_sorted = [&](int yy, float* v) {   // sorted 5-tuple of row yy (BORDER_REPLICATE) around x
                const float* row = Uf + (2u * L::at(pUs, clampi(yy, 0, g.H - 1), 0) + ch);
                if (interior) {                          // x-2 .. x+2 inside the image: one pointer + immediates
                    const float* q = row + xs[0];
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) v[dx] = ld_tap(q + 2 * dx);
                } else {
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) v[dx] = ld_tap(row + xs[dx]);
                }
                TF_MED_SORT5(v)
            };
            float T[5], S0[5], N1[5], Pp[10];
            {
                float a[5];
                load_sorted(y0 - 2, T);
                load_sorted(y0 - 1, a);
                load_sorted(y0, S0);
                med_merge55(a, S0, Pp);
                load_sorted(y0 + 1, N1);
            }
            for (int y = y0; y < y1; y += 2) {
                float N2[5], B[5], Pn[10], mid[6];
                load_sorted(y + 2, N2);
                load_sorted(y + 3, B);
                med_merge55(N1, N2, Pn);
                med_mid6(Pp, Pn, mid);
                Uw[2u * L::at(pUd, y, x) + ch] = med_finish(mid, T);
                if (y + 1 < y1) Uw[2u * L::at(pUd, y + 1, x) + ch] = med_finish(mid, B);
#pragma unroll
                for (int i = 0; i < 5; ++i) { T[i] = S0[i]; S0[i] = N2[i]; N1[i] = B[i]; }
#pragma unroll
                for (int i = 0; i < 10; ++i) Pp[i] = Pn[i];
            }
        }
        return;
    }
    for (int y = y0; y < y1; ++y) {
        float2 out;
        {
            float v[9], w[9];
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = clampi(y + dy, 0, g.H - 1);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = clampi(x + dx, 0, g.W - 1);
                    const float2 t = ld_tap(SB + L::at(pUs, yy, xx));
                    v[(dy + 1) * 3 + dx + 1] = t.x;
                    w[(dy + 1) * 3 + dx + 1] = t.y;
                }
            }
            out.x = median9(v);
            out.y = median9(w);
        }
        SB[L::at(pUd, y, x)] = out;
    }
}

// ---- inner iteration pieces -------------------------------------------------------------------------------
// The two flow channels are packed in float2 values and go through f32x2 instructions; every half rounds like
// the scalar operation of the C++ source.  Each piece has a branch-free FAST form that also reports whether its
// result can be trusted (operands inside the domain on which the fast sequence is proven exact) and an EXACT form
// (IEEE division, double-precision hypot); a row takes one merged branch to the exact forms when any lane needs it.
// ca = (I1wx, I1wy), cb = (grad, rho_c); kRhoPack: cb = (rho_c of the even row of the pair, rho_c of the odd row)
struct InnerRow { float2 u, ca, cb, px, py, pxl; };
struct InnerConst { float l_t, theta, taut, negzero; };

// estimateV: d = v - u.  c3_bad: the thresholding division ran outside the fast path's domain (needs_exact).
struct VStep { float2 d; float nrho, grad; bool bad; };
__device__ __forceinline__ VStep estimate_v_fast(const InnerRow& r, const InnerConst& K, bool odd_row) {
    VStep o;
    const float2 cu = mul2(r.ca, r.u);
    float rho_c = r.cb.y, grad = r.cb.x;
    if (kRhoPack) {                              // I1wx^2 + I1wy^2, rounded like the warp op's Ix2 + Iy2
        rho_c = odd_row ? r.cb.y : r.cb.x;
        const float2 sq = mul2_nofuse(r.ca, r.ca, K.negzero);
        grad = sq.x + sq.y;
    }
    const float rho = rho_c + (cu.x + cu.y);
    const float lg = K.l_t * grad;
    const bool c1 = rho < -lg;
    const bool c2 = !c1 && rho > lg;
    const bool c3 = !c1 && !c2 && grad > FLT_EPSILON;
    o.nrho = -rho; o.grad = grad;
    const float fi = div_with_rcp(o.nrho, grad, refined_rcp(grad));
    // in case 3, 2^-23 < grad and |rho| <= l_t * grad: of div_den_ok / div_fast_ok only these bounds can fail
    o.bad = c3 && (grad > 1.0995116e12f || (fabsf(rho) < 8.6736174e-19f && rho != 0.0f));
    // d = (l_t | -l_t | fi) * (I1wx, I1wy); (-l_t) * c == -(l_t * c) exactly
    const float k = c1 ? K.l_t : (c2 ? -K.l_t : fi);
    o.d = mul2(r.ca, splat2(k));
    if (!(c1 || c2 || c3)) o.d = make_float2(0.f, 0.f);
    return o;
}
__device__ __forceinline__ float2 estimate_v_exact(const InnerRow& r, const VStep& v) {   // only reached in case 3
    return mul2(r.ca, splat2(__fdiv_rn(v.nrho, v.grad)));
}

// theta * divergence(p) for one pixel
__device__ __forceinline__ float2 theta_div_px(const InnerRow& r, float2 pxl, float2 pyu, bool strip_at_x0,
                                               bool first_col_not_first_row, const InnerConst& K) {
    float2 dv = add2(sub2(r.px, pxl), sub2(r.py, pyu));      // interior; first row / corner: missing terms == 0
    if (strip_at_x0) {                                       // warp-uniform
        const float2 alt = sub2(add2(r.px, r.py), pyu);      // first column: v1 + v2 - v2(y-1)
        if (first_col_not_first_row) dv = alt;
    }
    return mul2_nofuse(dv, splat2(K.theta), K.negzero);
}

// estimateV + divergence + estimateU for one pixel (tvl1flow.cpp order of operations), self-contained form
__device__ __forceinline__ float2 estimate_u_px(const InnerRow& r, float2 pxl, float2 pyu, bool strip_at_x0,
                                                bool first_col_not_first_row, const InnerConst& K, bool odd_row) {
    VStep v = estimate_v_fast(r, K, odd_row);
    if (v.bad) v.d = estimate_v_exact(r, v);
    return add2(add2(r.u, v.d), theta_div_px(r, pxl, pyu, strip_at_x0, first_col_not_first_row, K));
}

__device__ __forceinline__ float hypot_f(float a, float b) {
    // static_cast<float>(hypot(a, b)) == (float)sqrt((double)a*a + (double)b*b)
    if (a == 0.f && b == 0.f) return 0.f;
    return (float)sqrt((double)a * (double)a + (double)b * (double)b);
}

// (hypot_f(a.x, b.x), hypot_f(a.y, b.y)) without double precision.
// s = a^2 + b^2 is held exactly as an unevaluated float sum (S + low: products split by fma, sum by two-sum);
// h = S * rsqrt(S) is corrected by one Newton step on the exact residual: h + c is within 2^-42 (relative) of the
// double-precision value the reference rounds to float.  Rounding h + c + delta and h + c - delta, with
// delta = 2^-37 h + 2^-15 |c| far above that error, gives the same float unless a rounding boundary is that close
// (probability ~2^-12 per value) -- by monotonicity of rounding that float is then the reference's result.
// ok = false also when s overflows (NaN: the comparison fails) or the operands are non-zero but below 2^-45.
__device__ __forceinline__ float2 hypot2_fast(float2 a, float2 b, float negzero, bool& ok) {
    const float2 A = mul2_nofuse(a, a, negzero), B = mul2_nofuse(b, b, negzero);   // they feed S = A + B
    const float2 Ae = fma2(a, a, neg2(A)), Be = fma2(b, b, neg2(B));
    const float2 S = add2(A, B);
    const float2 Bv = sub2(S, A), Av = sub2(S, Bv);
    const float2 Se = add2(sub2(A, Av), sub2(B, Bv));
    const float2 low = add2(Se, add2(Ae, Be));
    float2 y;   // s == 0 -> y = 2^50, h = c = 0, result 0
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(fmaxf(S.x, 7.888609e-31f)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(fmaxf(S.y, 7.888609e-31f)));
    const float2 h = mul2(S, y);
    const float2 rr = add2(fma2(neg2(h), h, S), low);
    const float2 c = mul2(rr, mul2(y, splat2(0.5f)));
    const float2 cabs = make_float2(fabsf(c.x), fabsf(c.y));
    const float2 delta = fma2(cabs, splat2(3.0517578e-5f), mul2(h, splat2(7.2759576e-12f)));   // 2^-15, 2^-37
    const float2 g1 = add2(h, add2(c, delta)), g2 = add2(h, sub2(c, delta));
    // a == b == 0 is exact; otherwise max(|a|, |b|) < 2^-45 is refused (the squares and their split parts underflow)
    const unsigned mx = __float_as_uint(fmaxf(fabsf(a.x), fabsf(b.x))), my = __float_as_uint(fmaxf(fabsf(a.y), fabsf(b.y)));
    const bool in_range = (mx - 1u) >= (0x29000000u - 1u) && (my - 1u) >= (0x29000000u - 1u);
    ok = in_range && (g1.x == g2.x) && (g1.y == g2.y);
    return g1;
}

// forwardGradient(u_new) + estimateDualVariables for one pixel: ux = (u1x, u2x), uy = (u1y, u2y),
// px = (p11, p21), py = (p12, p22); pxn / pyn receive the updated dual variables.  Returns false when the fast
// sequences were outside their domain (then pxn / pyn are not valid and dual_update_exact must be used).
__device__ __forceinline__ bool dual_update_fast(float2 ux, float2 uy, float2 px, float2 py, const InnerConst& K,
                                                 float2& pxn, float2& pyn) {
    bool hyp_ok;
    const float2 gg = hypot2_fast(ux, uy, K.negzero, hyp_ok);
    const float2 taut2 = splat2(K.taut);
    const float2 ng = add2(splat2(1.0f), mul2_nofuse(taut2, gg, K.negzero));
    const float2 ax = add2(px, mul2_nofuse(taut2, ux, K.negzero));    // (a11, a21)
    const float2 ay = add2(py, mul2_nofuse(taut2, uy, K.negzero));    // (a12, a22)
    // one guard for the four numerators: every |a| is 0 or >= 2^-100, the largest <= 2^100; ng in [1, 2^20]
    const float amax = fmaxf(fmaxf(fabsf(ax.x), fabsf(ay.x)), fmaxf(fabsf(ax.y), fabsf(ay.y)));
    const bool tiny = dual_num_tiny(ax.x) || dual_num_tiny(ay.x) || dual_num_tiny(ax.y) || dual_num_tiny(ay.y);
    // refined_rcp + div_with_rcp of tvl1_device.cuh, two channels per instruction
    float2 r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.x) : "f"(ng.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0.y) : "f"(ng.y));
    const float2 mng = neg2(ng);
    const float2 r = fma2(r0, fma2(mng, r0, splat2(1.0f)), r0);
    const float2 zero = splat2(0.0f);
    const float2 qx0 = fma2(ax, r, zero), qy0 = fma2(ay, r, zero);
    const float2 qx = fma2(r, fma2(mng, qx0, ax), qx0), qy = fma2(r, fma2(mng, qy0, ay), qy0);
    pxn = make_float2(or_sign(qx.x, ax.x), or_sign(qx.y, ax.y));
    pyn = make_float2(or_sign(qy.x, ay.x), or_sign(qy.y, ay.y));
    return hyp_ok && dual_ok(tiny, amax, fmaxf(ng.x, ng.y));
}
__device__ __forceinline__ void dual_update_exact(float2 ux, float2 uy, float2 px, float2 py, const InnerConst& K,
                                                  float2& pxn, float2& pyn) {
    const float g1 = hypot_f(ux.x, uy.x), g2 = hypot_f(ux.y, uy.y);
    const float ng1 = 1.0f + K.taut * g1, ng2 = 1.0f + K.taut * g2;
    pxn.x = __fdiv_rn(px.x + K.taut * ux.x, ng1); pyn.x = __fdiv_rn(py.x + K.taut * uy.x, ng1);
    pxn.y = __fdiv_rn(px.y + K.taut * ux.y, ng2); pyn.y = __fdiv_rn(py.y + K.taut * uy.y, ng2);
}

__device__ __forceinline__ float and_mask(float v, unsigned m) { return __uint_as_float(__float_as_uint(v) & m); }
__device__ __forceinline__ float and_or(float v, unsigned m, float o) {   // (v & m) | o : one LOP3
    return __uint_as_float((__float_as_uint(v) & m) | __float_as_uint(o));
}

// PH_INNER: one primal-dual iteration, warp-autonomous register-rolling strip.
// A warp owns kIW output columns (+1 halo column in lane 31) and walks down kIR rows; vertical neighbours stay in
// registers, horizontal ones come by warp shuffle, the loads of row y+2 are in flight while row y+1 is computed.
// Addressing: three row pointers (U, P, coefficient planes) advanced by one image row per iteration; every plane,
// the row look-ahead and the ping-pong partners are immediate offsets / one XOR from them.  Slots carry two pad
// rows below the image, so the look-ahead loads need no bounds test.  (A shared-memory ring filled by 16-byte
// cp.async copies, 2-4 rows deep, was measured: not faster -- the pure inner launch already runs at ~80 % of the
// HBM peak with this one-row register look-ahead -- and its 46 KB per CTA cost the other phases their L1.)
template <int PITCH>
__device__ __forceinline__ double op_inner(const EngineParams& P, int level, int ucur, int pcur, int slot, int strip,
                                           int lane) {
    using L = Lay<PITCH>;
    constexpr int PB = (int)L::PB, ROWB = (int)L::ROWB;
    const LevelGeom& g = P.lv[level];
    const int W = g.W, H = g.H;
    const InnerConst K = {P.l_t, P.theta, P.taut, P.negzero};

    const int x0 = (strip % g.in_sx) * kIW;
    const int y0 = (strip / g.in_sx) * kIR, y1 = min(y0 + kIR, H);
    const int x = x0 + lane;
    const bool valid = x < W;                    // lane computes u_new
    const bool owner = valid && lane < kIW;      // lane owns the outputs of its column
    const unsigned right_mask = (x + 1 < W) ? 0xffffffffu : 0u;   // forward x-difference exists
    const unsigned not_lane0 = lane == 0 ? 0u : 0xffffffffu;
    const bool strip_at_x0 = (x0 == 0);          // warp-uniform
    const bool first_col = (x == 0);
    const bool lane0_left = (lane == 0 && x0 > 0);
    const int xc = valid ? x : W - 1;            // clamp: idle lanes read a legal address
    const char* base = reinterpret_cast<const char*>(slot_base(P, slot)) + ((size_t)y0 * (size_t)ROWB + (size_t)(xc + kXMargin) * 8u);
    const char* pu = base + ucur * PB;                  // U[ucur] of row y        (partner: ^ PB)
    const char* pp = base + ((int)PL_PX + pcur) * PB;   // PX[pcur]; PY[pcur] at + 2 PB (partners: ^ PB)
    const char* pc = base + (int)PL_CA * PB;            // CA; CB at + PB

    auto ld = [](const char* p, int off) { return __ldcg(reinterpret_cast<const float2*>(p + off)); };
    // ping-pong partner plane: one XOR when PB is a power of two (address bits of row / plane / column are disjoint),
    // else an add of +-PB
    const int du = ucur ? -PB : PB, dp = pcur ? -PB : PB;
    auto partner_u = [&](const char* p) {
        if constexpr ((PB & (PB - 1)) == 0) return reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(p) ^ (uintptr_t)PB);
        else return const_cast<char*>(p) + du;
    };
    auto partner_p = [&](const char* p) {
        if constexpr ((PB & (PB - 1)) == 0) return reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(p) ^ (uintptr_t)PB);
        else return const_cast<char*>(p) + dp;
    };
    auto st = [](char* p, int off, float2 v) { *reinterpret_cast<float2*>(p + off) = v; };
    // kRhoPack: an odd image row shares the CB element of the even row above it (`above`): no load
    auto load_row = [&](int rows_ahead, bool odd, float2 above) {   // odd: parity of the loaded image row (warp-uniform)
        const int d = rows_ahead * ROWB;
        InnerRow r;
        r.u = ld(pu, d);
        r.ca = ld(pc, d);
        r.cb = above;
        if (!kRhoPack || !odd) r.cb = ld(pc, d + PB);
        r.px = ld(pp, d);
        r.py = ld(pp, d + 2 * PB);
        r.pxl = make_float2(0.f, 0.f);
        if (lane0_left) r.pxl = ld(pp, d - 8);   // only lane 0 of a strip that does not start at x = 0
        return r;
    };
    // The register look-ahead is one row (~1.5 us of work per warp); under load a DRAM access takes about as long, and
    // with the other phases' warps on the SM too few bytes are in flight (the first use of the next row's loads was
    // the kernel's top stall).  Rows further ahead are therefore pulled into L2 -- no registers, five instructions.
    auto prefetch_row = [&](int rows_ahead, bool odd) {
        const int d = rows_ahead * ROWB;
        prefetch_l2(pu + d); prefetch_l2(pc + d); prefetch_l2(pp + d); prefetch_l2(pp + d + 2 * PB);
        if (!kRhoPack || !odd) prefetch_l2(pc + d + PB);
    };
    // left neighbour's px: by shuffle, lane 0 takes the value it loaded itself (zero at the image border)
    auto left_px = [&](const InnerRow& r) {
        return make_float2(and_or(__shfl_up_sync(0xffffffffu, r.px.x, 1), not_lane0, r.pxl.x),
                           and_or(__shfl_up_sync(0xffffffffu, r.px.y, 1), not_lane0, r.pxl.y));
    };
    // forward x-difference of u_new (zero in the last image column)
    auto diff_x = [&](float2 un) {
        const float2 d = sub2(make_float2(__shfl_down_sync(0xffffffffu, un.x, 1), __shfl_down_sync(0xffffffffu, un.y, 1)), un);
        return make_float2(and_mask(d.x, right_mask), and_mask(d.y, right_mask));
    };

    double err = 0.0;
    float2 pyu = make_float2(0.f, 0.f);
    if (y0 > 0) pyu = ld(pp, 2 * PB - ROWB);
    const InnerRow cur = load_row(0, false, make_float2(0.f, 0.f));     // y0 is even (kIR is)
    InnerRow row = load_row(1, true, cur.cb);    // row y0 + 1 <= H: inside the image or the first pad row
    if (kPF > 0) {
#pragma unroll
        for (int k = 2; k < kPF; ++k) if (y0 + k < y1) prefetch_row(k, (k & 1) != 0);
    }

    float2 un = estimate_u_px(cur, left_px(cur), pyu, strip_at_x0, first_col && y0 > 0, K, false);
    if (owner) {
        st(partner_u(pu), 0, un);
        const float2 du = sub2(un, cur.u);
        const float2 sq = mul2(du, du);
        err += (double)(sq.x + sq.y);
    }
    float2 px_c = cur.px, py_c = cur.py;


    // rows y0 .. y1-2: row y+1 belongs to this strip (u_new stored, error counted); row y+2 <= H is loaded meanwhile.
    // `row` (y+1) is complete when a trip starts; the loads of row y+2 are issued first thing and are first touched
    // by the hand-over `row = nxt` at the END of the trip -- a whole row of work later.  (With the hand-over at the top
    // of the trip ptxas hoisted the new loads above it into scratch registers and copied them home at once: that copy
    // waited for the very load it was meant to hide -- 22 % of all stall samples of the run on one MOV.)
    // (A loop body of two trips with swapped row buffers -- no hand-over moves, compile-time row parity -- was
    // measured in round 2: 5.8 % slower on the clip, its larger live set spills inside the loop at 80 registers.)
#pragma unroll kInnerUnroll
    for (int y = y0; y < y1 - 1; ++y) {
        const bool y_odd = (y & 1) != 0;         // row y + 2 has the parity of y, row y + 1 the other one
        const InnerRow nxt = load_row(2, y_odd, row.cb);
        if (kPF > 0 && y + kPF < y1) prefetch_row(kPF, ((y + kPF) & 1) != 0);
        // u_new of row y+1, then forwardGradient(u_new) + estimateDualVariables of row y -- fast forms
        VStep v = estimate_v_fast(row, K, !y_odd);
        const float2 tdv = theta_div_px(row, left_px(row), py_c, strip_at_x0, first_col, K);
        float2 un_n = add2(add2(row.u, v.d), tdv);
        const float2 ux = diff_x(un);
        float2 pxn, pyn;
        const bool ok = dual_update_fast(ux, sub2(un_n, un), px_c, py_c, K, pxn, pyn);
        if (v.bad || !ok) {                      // rare: redo this lane's row with the exact forms
            if (v.bad) un_n = add2(add2(row.u, estimate_v_exact(row, v)), tdv);
            dual_update_exact(ux, sub2(un_n, un), px_c, py_c, K, pxn, pyn);
        }
        if (owner) {
            char* puw = partner_u(pu);
            char* ppw = partner_p(pp);
            st(puw, ROWB, un_n);
            st(ppw, 0, pxn);
            st(ppw, 2 * PB, pyn);
            const float2 du = sub2(un_n, row.u);
            const float2 sq = mul2(du, du);
            err += (double)(sq.x + sq.y);
        }
        un = un_n; px_c = row.px; py_c = row.py;
        row = nxt;
#if TEEFLOW_LATE_HANDOVER
        // ptxas otherwise hoists the hand-over of the dual variables to right behind their loads (to reuse the load's
        // registers as scratch), where it waits for them; tying it to a value that exists only at the end of the trip
        // (one LOP3 each, like the MOV it replaces) keeps the whole row of work between load and first use
        {
            const unsigned late = __float_as_uint(pxn.x) & P.zero_mask;
            row.px = make_float2(__uint_as_float(__float_as_uint(nxt.px.x) | late), __uint_as_float(__float_as_uint(nxt.px.y) | late));
            row.py = make_float2(__uint_as_float(__float_as_uint(nxt.py.x) | late), __uint_as_float(__float_as_uint(nxt.py.y) | late));
        }
        pu += ROWB; pp += ROWB; pc += ROWB;
    }
    // last row of the strip (y = y1-1): u_new of row y1 is only needed for the y-difference (the next strip owns it)
    {
        float2 uy = make_float2(0.f, 0.f);
        if (y1 < H) {                            // warp-uniform
            const float2 un_n = estimate_u_px(row, left_px(row), py_c, strip_at_x0, first_col, K, (y1 & 1) != 0);
            uy = sub2(un_n, un);
        }
        const float2 ux = diff_x(un);
        float2 pxn, pyn;
        if (!dual_update_fast(ux, uy, px_c, py_c, K, pxn, pyn)) dual_update_exact(ux, uy, px_c, py_c, K, pxn, pyn);
        char* ppw = partner_p(pp);
        if (owner) { st(ppw, 0, pxn); st(ppw, 2 * PB, pyn); }
    }
#endif
    // fixed-order warp reduction of the float64 error partial
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err += __shfl_down_sync(0xffffffffu, err, o);
    return err;
}

// PH_INNER, TMA-staged form (the shipped one when TEEFLOW_TMA_INNER): the same iteration, the same strip geometry and
// the same arithmetic in the same order as op_inner, but the input rows do not travel through registers one row
// ahead: lane 0 asks the TMA unit for boxes of kTR image rows (cp.async.bulk.tensor through the tensor maps of the
// level: U, CA, PX + PY in one box (from column x0-1, so every lane finds its left neighbour's px next to its own, and
// the column left of the image reads as zero) and the row-pair packed rho_c, all 34 columns wide), kTS boxes deep into the warp's shared-memory ring, each signalled by an mbarrier; the lanes pick their
// values up with 8-byte shared loads when the trip needs them.  Up to kTS * kTR rows per warp are in flight without a
// register, which is what the register form could not afford (its first use of the next row's loads held 9 % of the
// kernel's stall samples).  Columns >= W and the row below the image are out of bounds of the tensor map and read as
// zero: they take no exact-form branch and are never stored.  Stores go straight from registers to the partner planes.
template <int PITCH>
__device__ __forceinline__ double op_inner_tma(const EngineParams& P, int level, int ucur, int pcur, int slot, int strip,
                                               int lane, const TmaRing& T) {
    using L = Lay<PITCH>;
    constexpr int PB = (int)L::PB, ROWB = (int)L::ROWB;
    const LevelGeom& g = P.lv[level];
    const int W = g.W, H = g.H;
    const InnerConst K = {P.l_t, P.theta, P.taut, P.negzero};

    const int x0 = (strip % g.in_sx) * kIW;
    const int y0 = (strip / g.in_sx) * kIR, y1 = min(y0 + kIR, H);
    const int x = x0 + lane;
    const bool valid = x < W;
    const bool owner = valid && lane < kIW;      // lane owns the outputs of its column
    const unsigned right_mask = (x + 1 < W) ? 0xffffffffu : 0u;   // forward x-difference exists
    const bool strip_at_x0 = (x0 == 0);          // warp-uniform
    const bool first_col = (x == 0);
    const int xc = valid ? x : W - 1;            // idle lanes: a legal address (they never store)
    char* gbase = reinterpret_cast<char*>(slot_base(P, slot)) + ((size_t)y0 * (size_t)ROWB + (size_t)(xc + kXMargin) * 8u);
    char* gu = gbase + (ucur ^ 1) * PB;                     // partner U of row y0 (results)
    char* gp = gbase + ((int)PL_PX + (pcur ^ 1)) * PB;      // partner PX; partner PY at + 2 PB
    auto st = [](char* p, int off, float2 v) { *reinterpret_cast<float2*>(p + off) = v; };

    // input rows y0 .. y1 (row y1 only for the y-difference of the last row; none below the image)
    const int n_in = (y1 - y0) + (y1 < H ? 1 : 0);
    const int n_chunks = (n_in + kTR - 1) / kTR;
    const CUtensorMap* tm = P.tmaps + level * 3;
    const int xe = x0 & ~1, xp = (x0 - 1) & ~1;              // even first columns of the boxes (xp = -2 for the first strip)
    const unsigned lo = (unsigned)(x0 - xe + lane) * 8u;     // byte offset of this lane's column in a staged U / CA / rho_c row
    const unsigned lp = (unsigned)(x0 - 1 - xp + lane) * 8u; // ... of its LEFT neighbour's column in a staged PX / PY row
    unsigned ph = *T.phase;
    auto issue = [&](int c, int stage) {         // whole warp, convergent: chunk c = rows y0 + c kTR ... into `stage`
        tma_issue_chunk(T.ring + (unsigned)stage * kTStageB, T.mbar + (unsigned)stage * 8u, tm, xe, xp, ucur, (int)PL_PX + pcur,
                        y0 + c * kTR, slot);
    };
    // the planes were written by other SMs' ordinary stores (made visible before the task was published): order them
    // before the reads through the async proxy
    asm volatile("fence.proxy.async;" ::: "memory");
#pragma unroll
    for (int c = 0; c < kTS; ++c) if (c < n_chunks) issue(c, c);
    float2 pyu = make_float2(0.f, 0.f);          // py of the row above the strip: one direct load
    if (y0 > 0) pyu = __ldcg(reinterpret_cast<const float2*>(gbase + ((int)PL_PY + pcur) * PB - ROWB));

    int k = 0, rr = 0, stage = 0, c_next = kTS;
    bool timed_out = false;
    auto fetch = [&]() {                         // the next input row of the strip, from the ring
        if (rr == 0) {
            const uint32_t bar = T.mbar + (unsigned)stage * 8u;
            const uint32_t parity = (ph >> stage) & 1u;
            if (!mbar_try_wait(bar, parity)) {
                const long long t0 = clock64();
                unsigned spins = 0;
                while (!mbar_try_wait(bar, parity)) {
                    if ((++spins & 0x3ffu) == 0u) {
                        const bool dead = P.flow && *reinterpret_cast<volatile int*>(&P.flow->abort) != 0;
                        if (dead || clock64() - t0 > P.watchdog_cycles) {
                            if (P.flow && !dead) atomicExch(&P.flow->abort, 3);
                            timed_out = true;
                            break;
                        }
                    }
                }
            }
            ph ^= 1u << stage;
        }
        const uint32_t sb = T.ring + (unsigned)stage * kTStageB;
        const uint32_t a = sb + (unsigned)rr * kTRowB + lo;
        const uint32_t ap = sb + kTOffP + (unsigned)rr * (2u * kTRowB) + lp;
        InnerRow r;
        r.u = lds64(a + kTOffU);
        r.ca = lds64(a + kTOffCA);
        r.pxl = lds64(ap);                       // px of column x - 1 (zero left of the image)
        r.px = lds64(ap + 8u);
        r.py = lds64(ap + kTRowB + 8u);
        r.cb = lds64(sb + kTOffCB + (unsigned)(rr >> 1) * kTRowB + lo);   // (rho_c even row, rho_c odd row)
        ++k; ++rr;
        if (rr == kTR || k == n_in) {            // every lane has its values of this stage: hand it back to the TMA unit
            __syncwarp();
            if (c_next < n_chunks) issue(c_next, stage);
            ++c_next; rr = 0;
            stage = (stage + 1 == kTS) ? 0 : stage + 1;
        }
        return r;
    };
    auto diff_x = [&](float2 un) {               // forward x-difference of u_new (zero in the last image column)
        const float2 d = sub2(make_float2(__shfl_down_sync(0xffffffffu, un.x, 1), __shfl_down_sync(0xffffffffu, un.y, 1)), un);
        return make_float2(and_mask(d.x, right_mask), and_mask(d.y, right_mask));
    };

    double err = 0.0;
    const InnerRow cur = fetch();                // row y0 (even: kIR is)
    float2 un = estimate_u_px(cur, cur.pxl, pyu, strip_at_x0, first_col && y0 > 0, K, false);
    if (owner) {
        st(gu, 0, un);
        const float2 du = sub2(un, cur.u);
        const float2 sq = mul2(du, du);
        err += (double)(sq.x + sq.y);
    }
    float2 px_c = cur.px, py_c = cur.py;
#pragma unroll 1
    for (int y = y0; y < y1 - 1; ++y) {
        const InnerRow row = fetch();            // row y + 1
        // u_new of row y+1, then forwardGradient(u_new) + estimateDualVariables of row y -- fast forms
        VStep v = estimate_v_fast(row, K, (y & 1) == 0);
        const float2 tdv = theta_div_px(row, row.pxl, py_c, strip_at_x0, first_col, K);
        float2 un_n = add2(add2(row.u, v.d), tdv);
        const float2 ux = diff_x(un);
        float2 pxn, pyn;
        const bool ok = dual_update_fast(ux, sub2(un_n, un), px_c, py_c, K, pxn, pyn);
        if (v.bad || !ok) {                      // rare: redo this lane's row with the exact forms
            if (v.bad) un_n = add2(add2(row.u, estimate_v_exact(row, v)), tdv);
            dual_update_exact(ux, sub2(un_n, un), px_c, py_c, K, pxn, pyn);
        }
        if (owner) {
            st(gu, ROWB, un_n);
            st(gp, 0, pxn);
            st(gp, 2 * PB, pyn);
            const float2 du = sub2(un_n, row.u);
            const float2 sq = mul2(du, du);
            err += (double)(sq.x + sq.y);
        }
        un = un_n; px_c = row.px; py_c = row.py;
        gu += ROWB; gp += ROWB;
    }
    // last row of the strip (y = y1-1): u_new of row y1 is only needed for the y-difference (the next strip owns it)
    {
        float2 uy = make_float2(0.f, 0.f);
        if (y1 < H) {                            // warp-uniform
            const InnerRow row = fetch();
            const float2 un_n = estimate_u_px(row, row.pxl, py_c, strip_at_x0, first_col, K, (y1 & 1) != 0);
            uy = sub2(un_n, un);
        }
        const float2 ux = diff_x(un);
        float2 pxn, pyn;
        if (!dual_update_fast(ux, uy, px_c, py_c, K, pxn, pyn)) dual_update_exact(ux, uy, px_c, py_c, K, pxn, pyn);
        if (owner) { st(gp, 0, pxn); st(gp, 2 * PB, pyn); }
    }
    if (lane == 0) *T.phase = ph;
    __syncwarp();
    if (timed_out) err = __longlong_as_double(0x7ff8000000000000ll);   // the run is aborted; make the damage visible
    // fixed-order warp reduction of the float64 error partial
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err += __shfl_down_sync(0xffffffffu, err, o);
    return err;
}

// PH_INNER2: TWO primal-dual iterations in one pass over the state (temporal blocking): 40 bytes read and 24
// written per pixel for two iterations instead of one.  Register-only: a warp owns kIW2 = 29 output columns and kIR2
// rows; its 32 lanes sit on columns x0-1 .. x0+30 and it walks rows y0-1 .. y1+1, a four-stage software pipeline
// per row step r:
//   A(r)   u'  = first-iteration  u of row r        (all lanes)          needs p(r), p(r-1), u(r), coefficients(r)
//   B(r-1) p'  = first-iteration  p of row r-1      (lanes 0..30)        needs u'(r-1), u'(r), p(r-1)
//   C(r-1) u'' = second-iteration u of row r-1      (lanes 1..30)        needs u'(r-1), p'(r-1), p'(r-2), coeff.(r-1)
//   D(r-2) p'' = second-iteration p of row r-2      (lanes 1..29)        needs u''(r-2), u''(r-1), p'(r-2)
// Carried in registers between row steps: u'(r-1), p'(r-2), u''(r-2), the inputs p / coefficients of row r-1, and
// the loads of row r+1 (issued one step ahead, like the single-iteration op) -- no shared memory, no barrier, so the
// pass costs the other phases nothing (the round-1 version staged rows in a shared-memory ring whose carve-out took
// the L1 the other phases live in).  Horizontal neighbours come by warp shuffle.  Image borders follow the
// single-iteration rules in BOTH iterations; halo lanes outside the image read a clamped column, compute on
// whatever it holds and are masked where a neighbour reads them.  Error sums of both iterations are returned
// (err1: |u' - u|^2, err2: |u'' - u'|^2 over the owned pixels).
template <int PITCH>
__device__ __forceinline__ void op_inner2(const EngineParams& P, int level, int ucur, int pcur, int slot, int strip,
                                          int lane, double& err1, double& err2) {
    using L = Lay<PITCH>;
    constexpr int PB = (int)L::PB, ROWB = (int)L::ROWB;
    const LevelGeom& g = P.lv[level];
    const int W = g.W, H = g.H;
    const InnerConst K = {P.l_t, P.theta, P.taut, P.negzero};

    const int x0 = (strip % g.in2_sx) * kIW2;
    const int y0 = (strip / g.in2_sx) * kIR2, y1 = min(y0 + kIR2, H);
    const int c = x0 - 1 + lane;                              // this lane's column (-1 or >= W: dead halo lane)
    const bool owner = lane >= 1 && lane <= kIW2 && c < W;    // lane owns the outputs of its column
    const unsigned right_mask = (c + 1 < W) ? 0xffffffffu : 0u;   // forward x-difference exists
    const unsigned not_col0 = (c > 0) ? 0xffffffffu : 0u;         // a left neighbour exists
    const unsigned shfl_left = (lane != 0 && c > 0) ? 0xffffffffu : 0u;   // ... and it sits in lane - 1
    const bool lane0_left = (lane == 0 && c > 0);             // lane 0 loads the input px of column c - 1 itself
    const bool strip_at_x0 = (x0 == 0);                       // warp-uniform
    const bool first_col = (c == 0);
    const int ra = max(y0 - 1, 0);                            // first row of stage A
    const int re = min(y1 + 1, H - 1);                        // last row of stage A
    const int qb = min(y1, H - 1);                            // last row of stages B and C
    const int cc = clampi(c, 0, W - 1);                       // dead lanes read a legal address

    const char* base = reinterpret_cast<const char*>(slot_base(P, slot)) + ((size_t)ra * (size_t)ROWB + (size_t)(cc + kXMargin) * 8u);
    const char* pu = base + ucur * PB;                  // U[ucur] of row r
    const char* pp = base + ((int)PL_PX + pcur) * PB;   // PX[pcur]; PY[pcur] at + 2 PB
    const char* pc = base + (int)PL_CA * PB;            // CA; CB at + PB
    const int du = ucur ? -PB : PB, dp = pcur ? -PB : PB;     // ping-pong partner planes (the results go there)

    auto ld = [](const char* p, int off) { return __ldcg(reinterpret_cast<const float2*>(p + off)); };
    auto st = [](const char* p, int off, float2 v) { *reinterpret_cast<float2*>(const_cast<char*>(p) + off) = v; };
    auto load_row = [&](int rows_ahead, bool odd, float2 above) {   // see op_inner
        const int d = rows_ahead * ROWB;
        InnerRow r;
        r.u = ld(pu, d);
        r.ca = ld(pc, d);
        r.cb = above;
        if (!kRhoPack || !odd) r.cb = ld(pc, d + PB);
        r.px = ld(pp, d);
        r.py = ld(pp, d + 2 * PB);
        r.pxl = make_float2(0.f, 0.f);
        if (lane0_left) r.pxl = ld(pp, d - 8);
        return r;
    };
    auto prefetch_row = [&](int rows_ahead, bool odd) {    // into L2, see op_inner
        const int d = rows_ahead * ROWB;
        prefetch_l2(pu + d); prefetch_l2(pc + d); prefetch_l2(pp + d); prefetch_l2(pp + d + 2 * PB);
        if (!kRhoPack || !odd) prefetch_l2(pc + d + PB);
    };
    auto diff_x = [&](float2 v) {                // forward x-difference (zero in the last image column)
        const float2 d = sub2(make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1)), v);
        return make_float2(and_mask(d.x, right_mask), and_mask(d.y, right_mask));
    };
    auto left_of = [&](float2 v) {               // the left neighbour's value (zero at the image border)
        return make_float2(and_mask(__shfl_up_sync(0xffffffffu, v.x, 1), not_col0),
                           and_mask(__shfl_up_sync(0xffffffffu, v.y, 1), not_col0));
    };
    auto sq_norm = [](float2 a, float2 b) { const float2 d = sub2(a, b); const float2 q = mul2(d, d); return (double)(q.x + q.y); };

    double e1 = 0.0, e2 = 0.0;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 u1p = zero2;                                       // u'  of row r-1
    float2 p1x = zero2, p1y = zero2;                          // p'  of row r-2 (when a step starts)
    float2 u2p = zero2;                                       // u'' of row r-2 (when a step starts)
    float2 ca_p = zero2, cb_p = zero2, px_p = zero2;          // inputs of row r-1
    float2 py_p = ra >= 1 ? ld(pp, 2 * PB - ROWB) : zero2;
    // row ra; kRhoPack: an odd first row takes the CB element of the (even) row above it
    InnerRow row = load_row(0, (ra & 1) != 0, (kRhoPack && (ra & 1)) ? ld(pc, PB - ROWB) : zero2);

#pragma unroll 1
    for (int r = ra; r <= y1 + 1; ++r) {
        const bool has_a = r <= re;                            // row r exists (warp-uniform)
        InnerRow nxt = row;
        const bool r_odd = (r & 1) != 0;
        if (r < re) nxt = load_row(1, !r_odd, row.cb);         // row r + 1 is in flight while this step computes
        if (kPF > 0 && r + kPF <= re) prefetch_row(kPF, ((r + kPF) & 1) != 0);
        float2 u1 = zero2;
        if (has_a) {
            // A(r): estimateV + divergence(p) + estimateU, first iteration
            InnerRow in = row;
            in.pxl = make_float2(and_or(__shfl_up_sync(0xffffffffu, row.px.x, 1), shfl_left, row.pxl.x),
                                 and_or(__shfl_up_sync(0xffffffffu, row.px.y, 1), shfl_left, row.pxl.y));
            const float2 pyu = r >= 1 ? py_p : zero2;
            VStep v = estimate_v_fast(in, K, r_odd);
            if (v.bad) v.d = estimate_v_exact(in, v);
            u1 = add2(add2(in.u, v.d), theta_div_px(in, in.pxl, pyu, strip_at_x0, first_col && r > 0, K));
            if (owner && r >= y0 && r < y1) e1 += sq_norm(u1, in.u);
        }
        const int q = r - 1;
        float2 n1x = zero2, n1y = zero2, u2 = zero2;          // p' and u'' of row q
        if (q >= ra && q <= qb) {
            // B(q): forwardGradient(u') + estimateDualVariables, first iteration
            const float2 ux = diff_x(u1p);
            const float2 uy = has_a ? sub2(u1, u1p) : zero2;
            if (!dual_update_fast(ux, uy, px_p, py_p, K, n1x, n1y)) dual_update_exact(ux, uy, px_p, py_p, K, n1x, n1y);
            if (q >= y0) {
                // C(q): estimateV + divergence(p') + estimateU, second iteration
                InnerRow in;
                in.u = u1p; in.ca = ca_p; in.cb = cb_p; in.px = n1x; in.py = n1y;
                in.pxl = left_of(n1x);
                const float2 pyu = q >= 1 ? p1y : zero2;
                VStep v = estimate_v_fast(in, K, !r_odd);     // row q = r - 1
                if (v.bad) v.d = estimate_v_exact(in, v);
                u2 = add2(add2(u1p, v.d), theta_div_px(in, in.pxl, pyu, strip_at_x0, first_col && q > 0, K));
                if (owner && q < y1) { st(pu, du - ROWB, u2); e2 += sq_norm(u2, u1p); }
            }
        }
        const int q2 = r - 2;
        if (q2 >= y0 && q2 < y1) {
            // D(q2): forwardGradient(u'') + estimateDualVariables, second iteration
            const float2 ux = diff_x(u2p);
            const float2 uy = (q2 + 1 <= H - 1) ? sub2(u2, u2p) : zero2;
            float2 pxn, pyn;
            if (!dual_update_fast(ux, uy, p1x, p1y, K, pxn, pyn)) dual_update_exact(ux, uy, p1x, p1y, K, pxn, pyn);
            if (owner) { st(pp, dp - 2 * ROWB, pxn); st(pp, dp - 2 * ROWB + 2 * PB, pyn); }
        }
        ca_p = row.ca; cb_p = row.cb; px_p = row.px; py_p = row.py;
        row = nxt;                                             // first use of the loads: a whole row step after their issue
#if TEEFLOW_LATE_HANDOVER
        {   // see op_inner: keeps ptxas from hoisting the hand-over onto the loads
            const unsigned late = (__float_as_uint(u2.x) | __float_as_uint(n1y.y)) & P.zero_mask;
            row.px = make_float2(__uint_as_float(__float_as_uint(nxt.px.x) | late), __uint_as_float(__float_as_uint(nxt.px.y) | late));
            row.py = make_float2(__uint_as_float(__float_as_uint(nxt.py.x) | late), __uint_as_float(__float_as_uint(nxt.py.y) | late));
            row.ca = make_float2(__uint_as_float(__float_as_uint(nxt.ca.x) | late), __uint_as_float(__float_as_uint(nxt.ca.y) | late));
            row.cb = make_float2(__uint_as_float(__float_as_uint(nxt.cb.x) | late), __uint_as_float(__float_as_uint(nxt.cb.y) | late));
        }
#endif
        u1p = u1; p1x = n1x; p1y = n1y; u2p = u2;
        pu += ROWB; pp += ROWB; pc += ROWB;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e1 += __shfl_down_sync(0xffffffffu, e1, o);
        e2 += __shfl_down_sync(0xffffffffu, e2, o);
    }
    err1 = e1; err2 = e2;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const unsigned short lo = __half_as_ushort(__float2half_rn(a));
    const unsigned short hi = __half_as_ushort(__float2half_rn(b));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

// PH_WASE: background = mean(masked_flow[masked_flow != 0]) with masked_flow = flow * bkgd[all N frames]
// (calculate_optical_flow.py:649-652) == sum(w f [f != 0]) / sum(w [f != 0]) with w = sum_n bkgd[n]  (float64 sums)
template <int PITCH>
__device__ __forceinline__ void op_wase(const EngineParams& P, int ucur, int slot, int strip, int lane, double& sum,
                                        double& cnt) {
    using L = Lay<PITCH>;
    const LevelGeom& g = P.lv[0];
    const float2* SB = slot_base(P, slot);
    const float2* Wt = reinterpret_cast<const float2*>(P.wase_w);
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    sum = 0.0; cnt = 0.0;
    if (x < g.W) {
        for (int y = y0; y < y1; ++y) {
            const unsigned q = (unsigned)(y * g.W + x);
            const float2 u = __ldcg(SB + L::at(PL_U + (unsigned)ucur, y, x));
            const float2 w = __ldg(Wt + q);
            if (u.x != 0.f) { sum += (double)w.x * (double)u.x; cnt += (double)w.x; }
            if (u.y != 0.f) { sum += (double)w.y * (double)u.y; cnt += (double)w.y; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
}

// PH_FINAL: (merge(u1,u2) - background) * conversion_factor -> (H,W,2) f32 and/or f16
// (calculate_optical_flow.py:659, 600, 403)
template <int PITCH>
__device__ __forceinline__ void op_final(const EngineParams& P, int ucur, int pair, float bg, int slot, int strip,
                                         int lane) {
    using L = Lay<PITCH>;
    const LevelGeom& g = P.lv[0];
    const float2* SB = slot_base(P, slot);
    const size_t npx = (size_t)g.H * g.W;
    const int o0 = P.out_index[pair], o1 = P.dup_index[pair];
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    for (int y = y0; y < y1; ++y) {
        const unsigned q = (unsigned)(y * g.W + x);
        float2 u = __ldcg(SB + L::at(PL_U + (unsigned)ucur, y, x));
        u.x = (u.x - bg) * P.out_scale;
        u.y = (u.y - bg) * P.out_scale;
        if (P.flow_f32) {
            P.flow_f32[(size_t)o0 * npx + q] = u;
            if (o1 >= 0) P.flow_f32[(size_t)o1 * npx + q] = u;
        }
        if (P.flow_f16) {
            const uint32_t h = pack_half2(u.x, u.y);
            P.flow_f16[(size_t)o0 * npx + q] = h;
            if (o1 >= 0) P.flow_f16[(size_t)o1 * npx + q] = h;
        }
    }
}

// ------------------------------------------------------------------------------------------- strip dispatch
// How the strip ops read the slot planes (the dataflow kernel re-reads, within ONE launch, planes that other SMs
// rewrote in the meantime, so the non-coherent ld.global.nc path is never used for them):
//   * every op reads the planes with ld.global.cg -- served by L2, the point of coherence, so no stale L1 line can be
//     hit and no op depends on a fence dropping L1 lines.  That includes the ops that re-read neighbours (median: five
//     taps per row; level-init: the four bilinear taps; ld_tap()): their re-reads hit L2 instead of L1, which costs
//     the ALU-bound median 2 % phase-pure and wins 0.7 % on the clip, because the alternative -- plain loads behind an
//     acquire fence per strip (TEEFLOW_CG_NEIGHBOURS=0, needs_l1_acquire()) -- throws the SM's whole L1 away every
//     few microseconds under the warp op's gathers;
//   * a finished strip arrives at its slot's counter with a release atomic, not __threadfence() + atomicAdd (which is
//     MEMBAR.SC + CCTL.IVALL: again the whole L1, once per strip; +1.2 % on the clip).
// Read-only inputs (pyramid, WASE weights) keep ld.global.nc.


#ifndef TEEFLOW_FLOW_STATS
#define TEEFLOW_FLOW_STATS 0
#endif
#ifndef TEEFLOW_EARLY_PROBE
#define TEEFLOW_EARLY_PROBE 1    // dataflow scheduler: read the task descriptors at the cursor before the release fence
#endif
#ifndef TEEFLOW_PHASE_MASK
#define TEEFLOW_PHASE_MASK 0xffu
#endif
// TEEFLOW_PHASE_MASK (analysis builds only): compile a subset of the strip ops, to read one op's register need and
// SASS in isolation (tools/op_resources.sh); the shipped library has all of them
__device__ __forceinline__ constexpr bool has_op(int phase) { return ((TEEFLOW_PHASE_MASK >> phase) & 1u) != 0; }

template <int PITCH>
__device__ __forceinline__ void run_strip(const EngineParams& P, int phase, int level, int ucur, int pcur, int pair,
                                          float bg, int slot, int strip, int lane, const float4* s_cubic, const TmaRing& T,
                                          double& err, double& aux) {
    if (has_op(PH_LEVEL_INIT) && phase == PH_LEVEL_INIT) op_level_init<PITCH>(P, level, ucur, slot, strip, lane);
    else if (has_op(PH_WARP) && phase == PH_WARP) op_warp<PITCH>(P, level, ucur, pair, slot, strip, lane, s_cubic);
    else if (has_op(PH_MEDIAN) && phase == PH_MEDIAN) op_median<PITCH>(P, level, ucur, slot, strip, lane);
    else if (has_op(PH_INNER) && phase == PH_INNER) {
        if constexpr (kTma) err = op_inner_tma<PITCH>(P, level, ucur, pcur, slot, strip, lane, T);
        else err = op_inner<PITCH>(P, level, ucur, pcur, slot, strip, lane);
    }
    else if (has_op(PH_INNER2) && phase == PH_INNER2) op_inner2<PITCH>(P, level, ucur, pcur, slot, strip, lane, err, aux);
    else if (has_op(PH_WASE) && phase == PH_WASE) op_wase<PITCH>(P, ucur, slot, strip, lane, err, aux);
    else if (has_op(PH_FINAL) && phase == PH_FINAL) op_final<PITCH>(P, ucur, pair, bg, slot, strip, lane);
}

// Arrival of a finished strip at its slot's counter: a RELEASE atomic (the strip's stores -- all lanes', ordered before
// lane 0 by the __syncwarp that precedes it -- are visible before the count).  __threadfence() + atomicAdd does the
// same but compiles to MEMBAR.SC + CCTL.IVALL: every finished strip would throw away the whole L1 of its SM, about
// once per microsecond, and the warp op's bicubic gathers and the median's row re-reads live there.
#ifndef TEEFLOW_RELEASE_ARRIVE
#define TEEFLOW_RELEASE_ARRIVE 1
#endif
__device__ __forceinline__ unsigned arrive_release(unsigned* counter) {
#if TEEFLOW_RELEASE_ARRIVE
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
#else
    __threadfence();
    return atomicAdd(counter, 1u);
#endif
}

// lane 0 records the strip's error partial(s) and takes an arrival ticket of the slot; true for the warp that
// delivered the last strip of the task
__device__ __forceinline__ bool strip_arrive(const EngineParams& P, int phase, int slot, int strip, int n_items, int lane,
                                             double err, double aux) {
    int last = 0;
    if (lane == 0) {
        if (phase == PH_INNER) P.partial[(size_t)slot * P.max_tiles + strip] = err;
        if (phase == PH_WASE || phase == PH_INNER2) {
            P.partial[(size_t)slot * P.max_tiles + 2 * strip] = err;
            P.partial[(size_t)slot * P.max_tiles + 2 * strip + 1] = aux;
        }
        const unsigned ticket = arrive_release(P.arrive + slot);
        last = (ticket == (unsigned)n_items - 1u);
    }
    return __shfl_sync(0xffffffffu, last, 0) != 0;
}

// fixed-order (strip order, then a shuffle tree) float64 reduction of a finished task's partials
__device__ __forceinline__ void reduce_partials(const EngineParams& P, int phase, int slot, int n_items, int lane, double& e,
                                                double& e2) {
    e = 0.0; e2 = 0.0;
    const double* part = P.partial + (size_t)slot * P.max_tiles;
    if (phase == PH_INNER) {
        for (int t = lane; t < n_items; t += 32) e += __ldcg(part + t);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    } else if (phase == PH_WASE || phase == PH_INNER2) {
        for (int t = lane; t < n_items; t += 32) { e += __ldcg(part + 2 * t); e2 += __ldcg(part + 2 * t + 1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_down_sync(0xffffffffu, e, o);
            e2 += __shfl_down_sync(0xffffffffu, e2, o);
        }
    }
}

// One finished task -> the slot's next state (lane 0 of the finishing warp).  Returns the pair that was completed by
// this task (PH_FINAL), or -1.  `spec` points at the two-iteration statistics (applied, discarded).
__device__ __forceinline__ int next_state(const EngineParams& P, Slot& n, double e, double e2, int* next_pair) {
    if (n.phase == PH_WASE) {
        n.bg = (float)(e / e2);          // 0/0 -> NaN, like np.mean of an empty selection
        P.bg_out[n.pair] = n.bg;
        n.phase = PH_FINAL;
        return -1;
    }
    if (n.phase == PH_FINAL) {
        const int finished = n.pair;
        int* co = P.counters_out + (size_t)n.pair * kMaxLevels * 3;
        for (int l = 0; l < kMaxLevels; ++l) { co[l * 3] = n.cnt[l][0]; co[l * 3 + 1] = n.cnt[l][1]; co[l * 3 + 2] = n.cnt[l][2]; }
        const int next = atomicAdd(next_pair, 1);
        if (next < P.n_pairs) start_pair(P, n, next);
        else { n.pair = -1; n.phase = PH_IDLE; }
        return finished;
    }
    advance_slot(P, n, e, e2);
    return -1;
}

// ------------------------------------------------------------------------------------------- the super-step
// (stepped scheduler: one launch = one phase of every slot of a group; kept for per-phase timing / profiling and as
// the A/B reference of the dataflow kernel below -- teeflow_set_param("stepped", 1))
#ifndef TEEFLOW_MIN_CTAS
#define TEEFLOW_MIN_CTAS 6
#endif
template <int PITCH>
__global__ void __launch_bounds__(kThreads, TEEFLOW_MIN_CTAS)
tvl1_step_kernel(const __grid_constant__ EngineParams P, const int parity) {
    __shared__ int s_prefix[kMaxSlots + 1];
    __shared__ float4 s_cubic[32];
    // TMA staging ring of the inner iteration: kTS stages and mbarriers per warp (op_inner_tma)
    __shared__ __align__(128) unsigned char s_ring[kTma ? kWarpsPerCta * kTWarpB : 16];
    __shared__ __align__(8) unsigned long long s_mbar[kWarpsPerCta * kTS];
    __shared__ unsigned s_tphase[kWarpsPerCta];
    if (kTma) {
        if (threadIdx.x < kWarpsPerCta * kTS) mbar_init(smem_u32(&s_mbar[threadIdx.x]), 1u);
        if (threadIdx.x < kWarpsPerCta) s_tphase[threadIdx.x] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const TmaRing T = {smem_u32(s_ring) + (threadIdx.x >> 5) * kTWarpB, smem_u32(s_mbar) + (threadIdx.x >> 5) * (kTS * 8u),
                       &s_tphase[threadIdx.x >> 5]};

    // this launch serves the slot group [slot0, slot0 + S): groups run on separate streams so that the tail and
    // the launch gap of one group's step are filled by the other group's strips
    const Slot* __restrict__ cur = P.slots[parity] + P.slot0;
    Slot* __restrict__ nxt = P.slots[parity ^ 1] + P.slot0;
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid < 32) s_cubic[tid] = cubic_coeffs(tid);
    for (int s = tid; s < P.S; s += kThreads) s_prefix[s + 1] = items_of(P, cur[s].phase, cur[s].pair, cur[s].level);
    __syncthreads();
    if (tid == 0) {
        s_prefix[0] = 0;
        for (int s = 0; s < P.S; ++s) s_prefix[s + 1] += s_prefix[s];
    }
    __syncthreads();
    const int total = s_prefix[P.S];

    // slots without work this step: carry their state over to the other parity unchanged
    if (blockIdx.x == 0)
        for (int s = tid; s < P.S; s += kThreads)
            if (s_prefix[s + 1] == s_prefix[s]) nxt[s] = cur[s];

    // dynamic distribution: every warp pulls the next strip from a per-launch counter, so strips of unequal cost
    // (inner / median / warp phases mix in one launch) balance out; the counter of the other parity is re-armed.
    // Strips are handed out slot by slot in raster order: at any moment the running warps work on neighbouring
    // strips of very few slots and sweep the same image rows together, which is what keeps DRAM pages open.
    // The first strip of every warp is its global warp index (no burst of same-address atomics at launch); the
    // counter therefore starts at the number of warps of the grid.
    if (blockIdx.x == 0 && tid == 0) P.item_counter[parity ^ 1] = (int)gridDim.x * kWarpsPerCta;
    for (bool first = true;; first = false) {
        int item = (int)blockIdx.x * kWarpsPerCta + (tid >> 5);
        if (!first) {
            if (lane == 0) item = atomicAdd(P.item_counter + parity, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
        }
        if (item >= total) break;
        int lo = 0, hi = P.S;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] <= item) lo = mid; else hi = mid; }
        const int strip = item - s_prefix[lo];
        const int slot = P.slot0 + lo;            // absolute slot: planes, arrival counter, partials
        const Slot* sp = cur + lo;
        const int pair = sp->pair, phase = sp->phase, level = sp->level, ucur = sp->ucur, pcur = sp->pcur;
        const int n_items = s_prefix[lo + 1] - s_prefix[lo];

        double err = 0.0, aux = 0.0;
        run_strip<PITCH>(P, phase, level, ucur, pcur, pair, sp->bg, slot, strip, lane, s_cubic, T, err, aux);
        __syncwarp();
        if (strip_arrive(P, phase, slot, strip, n_items, lane, err, aux)) {
            // last strip of this slot for this step: reduce the partials in strip order and advance the slot
            __threadfence();
            double e, e2;
            reduce_partials(P, phase, slot, n_items, lane, e, e2);
            if (lane == 0) {
                Slot n = *sp;
                const int finished = next_state(P, n, e, e2, P.next_pair);
                if (finished >= 0) {
                    __threadfence();
                    P.done_order[atomicAdd(P.pairs_done, 1)] = finished;   // the host copies finished flows out early
                }
                nxt[lo] = n;
                P.arrive[slot] = 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- dataflow scheduler
// ONE launch per run: every phase of every slot is a task in a device-side FIFO; a task is published by the warp
// that retires the last strip of the slot's previous task, so slots advance independently of each other -- no grid
// barrier, no launch boundary, no host round trip.  Warps take strip tickets from one global counter; tickets map to
// tasks in publication order (tasks are handed out oldest first: the running warps still sweep neighbouring strips
// of few slots together, which keeps DRAM pages open).  A warp whose ticket belongs to a task that is not published
// yet waits for it; that can never deadlock: a task's predecessor (same slot) has smaller tickets, all of which are
// held by warps that run them, and no strip waits for anything.  The grid is launched cooperatively with exactly the
// resident CTA count, so every warp is running.  A watchdog (clock64) turns a would-be hang into an error code.
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v4(void* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// lane 0 of a finishing warp: append the task that runs slot state `n` next (or the terminal task)
__device__ __forceinline__ void publish_task(const EngineParams& P, int slot, const Slot& n, int phase, unsigned n_items) {
    const unsigned long long old = atomicAdd(&P.flow->alloc, (1ull << 32) | (unsigned long long)n_items);
    const unsigned k = (unsigned)(old >> 32), first = (unsigned)old;
    Task* t = P.tasks + (k & (kTaskRing - 1u));
    const unsigned what = (phase == PH_EXIT ? 0u : n_items) | ((unsigned)phase << 20) | ((unsigned)n.level << 24) |
                          ((unsigned)n.ucur << 28) | ((unsigned)n.pcur << 29);
    const unsigned who = (unsigned)slot | ((unsigned)(n.pair < 0 ? 0 : n.pair) << 9);
    __threadfence();                                   // everything the task reads is visible before its descriptor
    st_volatile_v4(t, make_uint4(k + 1u, first, what, who));
}

constexpr unsigned kExitItems = 0x40000000u;   // ticket range of the terminal task: every warp takes one more ticket

// lane 0 of the warp that retired the last strip of a task: advance the slot and hand its next task over
__device__ __forceinline__ void retire_task(const EngineParams& P, int slot, double e, double e2) {
    FlowCtl* const F = P.flow;
    Slot& n = P.slots[0][slot];                        // only the warp that retires a slot's task touches its state
    const int finished = next_state(P, n, e, e2, &F->next_pair);
    P.arrive[slot] = 0u;
    bool all_done = false;
    if (finished >= 0) {
        __threadfence();                               // the flow of the finished pair is visible before its index
        const int pos = atomicAdd(&F->pairs_done, 1);
        P.done_order[pos] = finished;
        if (P.host_done) {                             // the host copies finished flows out while the rest is solved
            __threadfence_system();
            P.host_done[pos] = finished;
        }
        all_done = (pos + 1 == P.n_pairs);
    }
    if (n.phase != PH_IDLE) publish_task(P, slot, n, n.phase, (unsigned)items_of(P, n.phase, n.pair, n.level));
    if (all_done) publish_task(P, slot, n, PH_EXIT, kExitItems);   // swallows every further ticket
}

template <int PITCH>
__global__ void __launch_bounds__(kThreads, TEEFLOW_MIN_CTAS)
tvl1_flow_kernel(const __grid_constant__ EngineParams P) {
    __shared__ float4 s_cubic[32];
    // TMA staging ring of the inner iteration: kTS stages and mbarriers per warp (op_inner_tma)
    __shared__ __align__(128) unsigned char s_ring[kTma ? kWarpsPerCta * kTWarpB : 16];
    __shared__ __align__(8) unsigned long long s_mbar[kWarpsPerCta * kTS];
    __shared__ unsigned s_tphase[kWarpsPerCta];
    if (kTma) {
        if (threadIdx.x < kWarpsPerCta * kTS) mbar_init(smem_u32(&s_mbar[threadIdx.x]), 1u);
        if (threadIdx.x < kWarpsPerCta) s_tphase[threadIdx.x] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const TmaRing T = {smem_u32(s_ring) + (threadIdx.x >> 5) * kTWarpB, smem_u32(s_mbar) + (threadIdx.x >> 5) * (kTS * 8u),
                       &s_tphase[threadIdx.x >> 5]};
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) s_cubic[tid] = cubic_coeffs(tid);
    __syncthreads();
    FlowCtl* const F = P.flow;
    unsigned kc = 0;                                   // task cursor of this warp: only ever moves forward
#if TEEFLOW_FLOW_STATS
    // diagnostic build: where do the warps spend their cycles?  [phase] strip time, [9] waiting for an unpublished
    // task, [10] everything else (ticket, descriptor probe, fences, arrival, task hand-over); counts in [16 + i]
    unsigned long long acc[11] = {0}, cnt[11] = {0};
    long long c_prev = clock64();
#define TF_STAT(i) { const long long c_now = clock64(); acc[i] += (unsigned long long)(c_now - c_prev); cnt[i] += 1; c_prev = c_now; }
#define TF_STAT_FLUSH() { if (lane == 0) for (int i = 0; i < 11; ++i) { atomicAdd(P.flow_stats + i, acc[i]); atomicAdd(P.flow_stats + 16 + i, cnt[i]); } }
#else
#define TF_STAT(i)
#define TF_STAT_FLUSH()
#endif

    // The scheduling round trips of a strip overlap each other instead of queueing up behind one another: the next
    // ticket is requested BEFORE the release fence of the strip just finished (the fence drains the strip's stores
    // while the ticket travels), and the arrival ticket is only looked at after the first descriptor probe of the
    // next strip.  `pend_*` describe the strip whose arrival is still in flight (lane 0 holds its arrival ticket).
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(&F->ticket, 1u);
    unsigned pend_arrive = 0, pend_n = 0;              // pend_n == 0: nothing pending
    int pend_slot = 0, pend_phase = 0;
#if TEEFLOW_EARLY_PROBE
    uint4 d_early = make_uint4(0u, 0u, 0u, 0u);        // descriptors at the cursor, read before the release fence
    bool have_early = false;
#endif

    for (;;) {
        t = __shfl_sync(0xffffffffu, t, 0);
        // ---- which task owns ticket t?  32 descriptors per probe, one per lane
        uint4 desc = make_uint4(0u, 0u, 0u, 0u);
        for (bool first_probe = true;; first_probe = false) {
            const Task* q = P.tasks + ((kc + (unsigned)lane) & (kTaskRing - 1u));
            // the whole descriptor: one 16-byte store of the publisher.  A published descriptor never changes, so the
            // copy read before the previous strip's fence serves the first probe (kc has not moved since)
#if TEEFLOW_EARLY_PROBE
            const uint4 d = (first_probe && have_early) ? d_early : ld_volatile_v4(q);
#else
            const uint4 d = ld_volatile_v4(q);
#endif
            const unsigned want = kc + (unsigned)lane + 1u;
            const bool pub = d.x == want;
            const bool mine = pub && (((d.z >> 20) & 0xfu) == (unsigned)PH_EXIT || (t - d.y) < (d.z & kTaskItemsMax));
            const unsigned hit = __ballot_sync(0xffffffffu, mine);
            const unsigned unp = __ballot_sync(0xffffffffu, !pub);
            const bool lost = __any_sync(0xffffffffu, (int)(d.x - want) > 0);
            if (first_probe && pend_n) {
                // the arrival of the previous strip has landed meanwhile: was it the last one of its task?
                const bool last = __shfl_sync(0xffffffffu, (int)(pend_arrive == pend_n - 1u), 0) != 0;
                if (last) {
                    __threadfence();
                    double e, e2;
                    reduce_partials(P, pend_phase, pend_slot, (int)pend_n, lane, e, e2);
                    if (lane == 0) retire_task(P, pend_slot, e, e2);
                }
                pend_n = 0;
            }
            if (hit) {
                const int src = __ffs(hit) - 1;
                kc += (unsigned)src;
                desc.x = __shfl_sync(0xffffffffu, d.x, src); desc.y = __shfl_sync(0xffffffffu, d.y, src);
                desc.z = __shfl_sync(0xffffffffu, d.z, src); desc.w = __shfl_sync(0xffffffffu, d.w, src);
                break;
            }
            // a descriptor of a later lap of the ring: this warp lagged kTaskRing tasks behind (cannot happen in practice)
            if (lost) { if (lane == 0) atomicExch(&F->abort, 2); TF_STAT_FLUSH() return; }
            if (unp == 0u) { kc += 32u; continue; }    // 32 published tasks, all of them before ticket t
            kc += (unsigned)(__ffs(unp) - 1);          // the first unpublished one: wait until it appears (one lane polls)
            TF_STAT(10)
            int give_up = 0;
            if (lane == 0) {
                const unsigned* seq = &P.tasks[kc & (kTaskRing - 1u)].seq;
                const long long t0 = clock64();
                unsigned ns = 128;
                while (ld_volatile_u32(seq) != kc + 1u) {
                    if (*reinterpret_cast<volatile int*>(&F->abort)) { give_up = 1; break; }
                    if (clock64() - t0 > P.watchdog_cycles) { atomicExch(&F->abort, 1); give_up = 1; break; }
                    __nanosleep(ns);
                    if (ns < 2048) ns *= 2;
                }
            }
            TF_STAT(9)
            if (__shfl_sync(0xffffffffu, give_up, 0)) { TF_STAT_FLUSH() return; }
        }
        const int phase = (int)((desc.z >> 20) & 0xfu);
        if (phase == PH_EXIT) { TF_STAT(10) TF_STAT_FLUSH() return; }
        const unsigned n_items = desc.z & kTaskItemsMax;
        const int level = (int)((desc.z >> 24) & 0xfu), ucur = (int)((desc.z >> 28) & 1u), pcur = (int)((desc.z >> 29) & 1u);
        const int slot = (int)(desc.w & 0x1ffu), pair = (int)(desc.w >> 9);
        const int strip = (int)(t - desc.y);
        if (needs_l1_acquire(phase)) __threadfence();  // plain (L1) plane reads ahead: drop the SM's stale lines
        const float bg = (phase == PH_FINAL && P.wase_w) ? __ldcg(P.bg_out + pair) : 0.f;
        TF_STAT(10)

        double err = 0.0, aux = 0.0;
        run_strip<PITCH>(P, phase, level, ucur, pcur, pair, bg, slot, strip, lane, s_cubic, T, err, aux);
        __syncwarp();
        TF_STAT(phase)
#if TEEFLOW_EARLY_PROBE
        // the probe of the next ticket's descriptor is in flight while lane 0's fence drains this strip's stores
        d_early = ld_volatile_v4(P.tasks + ((kc + (unsigned)lane) & (kTaskRing - 1u)));
        have_early = true;
#endif
        if (lane == 0) {
            if (phase == PH_INNER) P.partial[(size_t)slot * P.max_tiles + strip] = err;
            if (phase == PH_WASE || phase == PH_INNER2) {
                P.partial[(size_t)slot * P.max_tiles + 2 * strip] = err;
                P.partial[(size_t)slot * P.max_tiles + 2 * strip + 1] = aux;
            }
            // travels while the release below drains this strip's stores.  (Requesting it already when the strip STARTS
            // was measured: 9 % slower -- a strip that is claimed but not started delays the completion of its task.)
            t = atomicAdd(&F->ticket, 1u);
            pend_arrive = arrive_release(P.arrive + slot);
        }
        pend_n = n_items; pend_slot = slot; pend_phase = phase;
    }
}

}  // namespace teeflow
