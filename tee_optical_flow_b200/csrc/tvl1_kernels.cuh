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

constexpr int kThreads = 256;  // threads per CTA
constexpr int kWarpsPerCta = kThreads / 32;
constexpr int kIW = 31;        // inner strip: output columns per warp (lane 31 = right halo column)
constexpr int kIR = 16;        // inner strip: rows per warp (16 beats 32 / 24 / 12 / 8 once strips are handed out dynamically)
constexpr int kPR = 8;         // pointwise strip: rows per warp (32 columns)
#ifndef TEEFLOW_DYNAMIC_ITEMS
#define TEEFLOW_DYNAMIC_ITEMS 1
#endif

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
__device__ inline void advance_slot(const EngineParams& P, Slot& s, double err_sum) {
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
        default: return;
    }
    for (;;) {
        if (todo == CHECK_INNER) {
            if (s.error > eps && s.n_inner < P.inner) { s.phase = PH_INNER; return; }
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
    s.ucur = 0; s.pcur = 0; s.error = FLT_MAX; s.bg = 0.f;
    for (int l = 0; l < kMaxLevels; ++l) { s.cnt[l][0] = 0; s.cnt[l][1] = 0; s.cnt[l][2] = 0; }
}

__device__ __forceinline__ int items_of(const EngineParams& P, int phase, int pair, int level) {
    if (phase == PH_IDLE || pair < 0) return 0;
    if (phase == PH_FINAL || phase == PH_WASE) return P.lv[0].pw_items;
    return phase == PH_INNER ? P.lv[level].in_items : P.lv[level].pw_items;
}

// ------------------------------------------------------------------------------------------------ strip ops
// All slot planes are addressed as  slot base (one 64-bit pointer) + 32-bit float2 element index.
__device__ __forceinline__ float2* slot_base(const EngineParams& P, int slot) {
    return P.planes + (size_t)slot * (size_t)P.slot_stride;
}
__device__ __forceinline__ unsigned plane_at(const EngineParams& P, unsigned plane) { return plane * (unsigned)P.slot_px; }

// PH_LEVEL_INIT: u = 0 (coarsest) or u = resize(u_coarse, INTER_LINEAR) * (1/scaleStep); p = 0
__device__ __forceinline__ void op_level_init(const EngineParams& P, int level, int ucur, int slot, int strip, int lane) {
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const bool coarsest = (level == P.L - 1);
    const unsigned oUd = plane_at(P, PL_U + (coarsest ? 0u : (unsigned)(ucur ^ 1)));
    const unsigned oUs = plane_at(P, PL_U + (unsigned)ucur);
    const unsigned oPX = plane_at(P, PL_PX), oPY = plane_at(P, PL_PY);
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    int x0 = 0, x1 = 0; float a0 = 0.f, a1 = 0.f;
    int cH = 0, cW = 0;
    if (!coarsest) {
        cH = P.lv[level + 1].H; cW = P.lv[level + 1].W;
        lin_coeff_x(x, g.up_sx, cW, x0, x1, a0, a1);
    }
    for (int y = y0; y < y1; ++y) {
        const unsigned q = (unsigned)(y * g.W + x);
        float2 u = make_float2(0.f, 0.f);
        if (!coarsest) {
            int ya, yb; float b0, b1;
            lin_coeff_y(y, g.up_sy, cH, ya, yb, b0, b1);
            const float2 s00 = SB[oUs + (unsigned)(ya * cW + x0)], s01 = SB[oUs + (unsigned)(ya * cW + x1)];
            const float2 s10 = SB[oUs + (unsigned)(yb * cW + x0)], s11 = SB[oUs + (unsigned)(yb * cW + x1)];
            const float r0x = s00.x * a0 + s01.x * a1, r1x = s10.x * a0 + s11.x * a1;
            const float r0y = s00.y * a0 + s01.y * a1, r1y = s10.y * a0 + s11.y * a1;
            u.x = (r0x * b0 + r1x * b1) * P.up_mul;
            u.y = (r0y * b0 + r1y * b1) * P.up_mul;
        }
        SB[oUd + q] = u;
        SB[oPX + q] = make_float2(0.f, 0.f);
        SB[oPY + q] = make_float2(0.f, 0.f);
    }
}

// PH_WARP: buildFlowMap + remap(I1, I1x, I1y; INTER_CUBIC) + calcGradRho
__device__ __forceinline__ void op_warp(const EngineParams& P, int level, int ucur, int pair, int slot, int strip,
                                        int lane, const float4* s_cubic) {
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const float2* U = SB + plane_at(P, PL_U + (unsigned)ucur);
    float4* COEF = reinterpret_cast<float4*>(SB + plane_at(P, PL_COEF));
    const int fa = P.pair_a[pair], fb = P.pair_b[pair];
    const float* I0 = P.pyrI + (size_t)fa * P.frame_pyr_stride + g.pyr_off;
    const float4* G1 = P.pyrG + (size_t)fb * P.frame_pyr_stride + g.pyr_off;
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    // the flow / I0 of the next row are fetched while the current row's gather runs (one row ahead)
    float2 u_n = __ldg(U + (unsigned)(y0 * g.W + x));
    float i0_n = __ldg(I0 + (unsigned)(y0 * g.W + x));
    for (int y = y0; y < y1; ++y) {
        const unsigned q = (unsigned)(y * g.W + x);
        const float2 u = u_n;
        const float i0 = i0_n;
        if (y + 1 < y1) { u_n = __ldg(U + q + (unsigned)g.W); i0_n = __ldg(I0 + q + (unsigned)g.W); }
        const float mx = (float)x + u.x, my = (float)y + u.y;
        const float3 w = remap_cubic3(G1, g.H, g.W, mx, my, s_cubic);
        const float Ix2 = w.y * w.y, Iy2 = w.z * w.z;
        float4 c;
        c.x = w.y; c.y = w.z;
        c.z = Ix2 + Iy2;
        c.w = (w.x - w.y * u.x - w.z * u.y - i0);
        COEF[q] = c;
    }
}

// PH_MEDIAN: medianBlur(u1, ksize), medianBlur(u2, ksize) with BORDER_REPLICATE
__device__ __forceinline__ void op_median(const EngineParams& P, int level, int ucur, int slot, int strip, int lane) {
    const LevelGeom& g = P.lv[level];
    float2* SB = slot_base(P, slot);
    const float2* __restrict__ Us = SB + plane_at(P, PL_U + (unsigned)ucur);
    float2* __restrict__ Ud = SB + plane_at(P, PL_U + (unsigned)(ucur ^ 1));
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    for (int y = y0; y < y1; ++y) {
        float2 out;
        if (P.median == 5) {
            float v[25], w[25];
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy) {
                const int yy = clampi(y + dy, 0, g.H - 1);
#pragma unroll
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xx = clampi(x + dx, 0, g.W - 1);
                    const float2 t = __ldg(Us + (unsigned)(yy * g.W + xx));
                    v[(dy + 2) * 5 + dx + 2] = t.x;
                    w[(dy + 2) * 5 + dx + 2] = t.y;
                }
            }
            out.x = median25(v);
            out.y = median25(w);
        } else {
            float v[9], w[9];
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = clampi(y + dy, 0, g.H - 1);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = clampi(x + dx, 0, g.W - 1);
                    const float2 t = __ldg(Us + (unsigned)(yy * g.W + xx));
                    v[(dy + 1) * 3 + dx + 1] = t.x;
                    w[(dy + 1) * 3 + dx + 1] = t.y;
                }
            }
            out.x = median9(v);
            out.y = median9(w);
        }
        Ud[(unsigned)(y * g.W + x)] = out;
    }
}

// ---- inner iteration pieces -------------------------------------------------------------------------------
struct InnerRow { float2 u; float4 c; float2 px, py, pxl; };

// estimateV + divergence + estimateU for one pixel (tvl1flow.cpp order of operations), branch-free: the
// thresholding division runs on every lane through the shared fast path and is selected where it applies.
__device__ __forceinline__ float2 estimate_u_px(const InnerRow& r, float2 pxl, float2 pyu, bool x_is_0, bool y_is_0,
                                                float l_t, float theta) {
    const float4 c = r.c;
    const float rho = c.w + (c.x * r.u.x + c.y * r.u.y);
    const float lg = l_t * c.z;
    const bool c1 = rho < -lg;
    const bool c2 = !c1 && rho > lg;
    const bool c3 = !c1 && !c2 && c.z > FLT_EPSILON;
    const float nrho = -rho;
    float fi = div_with_rcp(nrho, c.z, refined_rcp(c.z));
    if (c3 && !(div_den_ok(c.z) && div_fast_ok(nrho))) fi = __fdiv_rn(nrho, c.z);   // rare: IEEE slow path
    const float a1 = l_t * c.x, a2 = l_t * c.y;
    const float d1 = c1 ? a1 : (c2 ? -a1 : (c3 ? fi * c.x : 0.f));
    const float d2 = c1 ? a2 : (c2 ? -a2 : (c3 ? fi * c.y : 0.f));
    const float v1 = r.u.x + d1, v2 = r.u.y + d2;
    float div1, div2;
    if (x_is_0 && !y_is_0) {          // first column: v1 + v2 - v2(y-1)
        div1 = r.px.x + r.py.x - pyu.x;
        div2 = r.px.y + r.py.y - pyu.y;
    } else {                          // interior; first row / corner follow with the missing terms == 0
        div1 = (r.px.x - pxl.x) + (r.py.x - pyu.x);
        div2 = (r.px.y - pxl.y) + (r.py.y - pyu.y);
    }
    return make_float2(v1 + theta * div1, v2 + theta * div2);
}

__device__ __forceinline__ float hypot_f(float a, float b) {
    // static_cast<float>(hypot(a, b)) == (float)sqrt((double)a*a + (double)b*b)
    if (a == 0.f && b == 0.f) return 0.f;
    return (float)sqrt((double)a * (double)a + (double)b * (double)b);
}

// PH_INNER: one primal-dual iteration, warp-autonomous register-rolling strip.
// Addressing: one slot base pointer + seven 32-bit element indices that advance by W per row.
__device__ __forceinline__ double op_inner(const EngineParams& P, int level, int ucur, int pcur, int slot, int strip,
                                           int lane) {
    const LevelGeom& g = P.lv[level];
    const int W = g.W, H = g.H;
    float2* __restrict__ SB = slot_base(P, slot);
    const float4* __restrict__ SB4 = reinterpret_cast<const float4*>(SB);
    const float l_t = P.l_t, theta = P.theta, taut = P.taut;

    const int x0 = (strip % g.in_sx) * kIW;
    const int y0 = (strip / g.in_sx) * kIR, y1 = min(y0 + kIR, H);
    const int x = x0 + lane;
    const bool valid = x < W;                    // lane computes u_new
    const bool owner = valid && lane < kIW;      // lane owns the outputs of its column
    const bool has_right = x + 1 < W;
    const bool x_is_0 = (x == 0);
    const bool lane0_left = (lane == 0 && x0 > 0);
    const int xc = valid ? x : W - 1;            // clamp: idle lanes read a legal address
    const unsigned q0 = (unsigned)(y0 * W + xc); // pixel offset of the strip's first row
    // element indices (float2 units; COEF in float4 units) of row y, advanced by W per row
    unsigned iU = plane_at(P, PL_U + (unsigned)ucur) + q0, iUn = plane_at(P, PL_U + (unsigned)(ucur ^ 1)) + q0;
    unsigned iPX = plane_at(P, PL_PX + (unsigned)pcur) + q0, iPXn = plane_at(P, PL_PX + (unsigned)(pcur ^ 1)) + q0;
    unsigned iPY = plane_at(P, PL_PY + (unsigned)pcur) + q0, iPYn = plane_at(P, PL_PY + (unsigned)(pcur ^ 1)) + q0;
    unsigned iC = (plane_at(P, PL_COEF) >> 1) + q0;
    const unsigned uW = (unsigned)W;

    auto load_row = [&](unsigned rows_ahead) {
        const unsigned d = rows_ahead * uW;
        InnerRow r;
        r.u = __ldg(SB + iU + d);
        r.c = __ldg(SB4 + iC + d);
        r.px = __ldg(SB + iPX + d);
        r.py = __ldg(SB + iPY + d);
        r.pxl = make_float2(0.f, 0.f);
        if (lane0_left) r.pxl = __ldg(SB + iPX + d - 1);   // only lane 0 of a strip that does not start at x = 0
        return r;
    };

    double err = 0.0;
    float2 pyu = make_float2(0.f, 0.f);
    if (y0 > 0) pyu = __ldg(SB + iPY - uW);
    InnerRow cur = load_row(0);
    InnerRow nxt = cur;
    if (y0 + 1 < H) nxt = load_row(1);

    float2 pxl = make_float2(__shfl_up_sync(0xffffffffu, cur.px.x, 1), __shfl_up_sync(0xffffffffu, cur.px.y, 1));
    if (lane == 0) pxl = cur.pxl;
    float2 un = estimate_u_px(cur, pxl, pyu, x_is_0, y0 == 0, l_t, theta);
    if (owner) {
        SB[iUn] = un;
        const float t = (un.x - cur.u.x) * (un.x - cur.u.x) + (un.y - cur.u.y) * (un.y - cur.u.y);
        err += (double)t;
    }
    float2 px_c = cur.px, py_c = cur.py;

#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
        const bool has_next = (y + 1 < H);       // warp-uniform
        float2 un_n = make_float2(0.f, 0.f);
        InnerRow row = nxt;                      // row y+1 (already in flight)
        if (y + 2 < H && y + 1 < y1) nxt = load_row(2);
        if (has_next) {
            float2 pl = make_float2(__shfl_up_sync(0xffffffffu, row.px.x, 1), __shfl_up_sync(0xffffffffu, row.px.y, 1));
            if (lane == 0) pl = row.pxl;
            un_n = estimate_u_px(row, pl, py_c, x_is_0, false, l_t, theta);
            if (owner && y + 1 < y1) {
                SB[iUn + uW] = un_n;
                const float t = (un_n.x - row.u.x) * (un_n.x - row.u.x) + (un_n.y - row.u.y) * (un_n.y - row.u.y);
                err += (double)t;
            }
        }
        // forwardGradient(u_new) + estimateDualVariables for row y
        const float unr_x = __shfl_down_sync(0xffffffffu, un.x, 1), unr_y = __shfl_down_sync(0xffffffffu, un.y, 1);
        const float u1x = has_right ? unr_x - un.x : 0.f, u2x = has_right ? unr_y - un.y : 0.f;
        const float u1y = has_next ? un_n.x - un.x : 0.f, u2y = has_next ? un_n.y - un.y : 0.f;
        const float g1 = hypot_f(u1x, u1y), g2 = hypot_f(u2x, u2y);
        const float ng1 = 1.0f + taut * g1, ng2 = 1.0f + taut * g2;
        const float a11 = px_c.x + taut * u1x, a12 = py_c.x + taut * u1y;
        const float a21 = px_c.y + taut * u2x, a22 = py_c.y + taut * u2y;
        float2 pxn, pyn;
        // one guard for the four numerators: every |a| is 0 or >= 2^-100, the largest <= 2^100; ng in [1, 2^20]
        const float amax = fmaxf(fmaxf(fabsf(a11), fabsf(a12)), fmaxf(fabsf(a21), fabsf(a22)));
        const bool tiny = dual_num_tiny(a11) || dual_num_tiny(a12) || dual_num_tiny(a21) || dual_num_tiny(a22);
        if (dual_ok(tiny, amax, fmaxf(ng1, ng2))) {
            const float r1 = refined_rcp(ng1), r2 = refined_rcp(ng2);
            pxn.x = div_with_rcp(a11, ng1, r1); pyn.x = div_with_rcp(a12, ng1, r1);
            pxn.y = div_with_rcp(a21, ng2, r2); pyn.y = div_with_rcp(a22, ng2, r2);
        } else {
            pxn.x = __fdiv_rn(a11, ng1); pyn.x = __fdiv_rn(a12, ng1);
            pxn.y = __fdiv_rn(a21, ng2); pyn.y = __fdiv_rn(a22, ng2);
        }
        if (owner) { SB[iPXn] = pxn; SB[iPYn] = pyn; }
        un = un_n; px_c = row.px; py_c = row.py;
        iU += uW; iUn += uW; iPX += uW; iPXn += uW; iPY += uW; iPYn += uW; iC += uW;
    }
    // fixed-order warp reduction of the float64 error partial
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err += __shfl_down_sync(0xffffffffu, err, o);
    return err;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const unsigned short lo = __half_as_ushort(__float2half_rn(a));
    const unsigned short hi = __half_as_ushort(__float2half_rn(b));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

// PH_WASE: background = mean(masked_flow[masked_flow != 0]) with masked_flow = flow * bkgd[all N frames]
// (calculate_optical_flow.py:649-652) == sum(w f [f != 0]) / sum(w [f != 0]) with w = sum_n bkgd[n]  (float64 sums)
__device__ __forceinline__ void op_wase(const EngineParams& P, int ucur, int slot, int strip, int lane, double& sum,
                                        double& cnt) {
    const LevelGeom& g = P.lv[0];
    const float2* U = slot_base(P, slot) + plane_at(P, PL_U + (unsigned)ucur);
    const float2* Wt = reinterpret_cast<const float2*>(P.wase_w);
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    sum = 0.0; cnt = 0.0;
    if (x < g.W) {
        for (int y = y0; y < y1; ++y) {
            const unsigned q = (unsigned)(y * g.W + x);
            const float2 u = U[q];
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
__device__ __forceinline__ void op_final(const EngineParams& P, int ucur, int pair, float bg, int slot, int strip,
                                         int lane) {
    const LevelGeom& g = P.lv[0];
    const float2* U = slot_base(P, slot) + plane_at(P, PL_U + (unsigned)ucur);
    const size_t npx = (size_t)g.H * g.W;
    const int o0 = P.out_index[pair], o1 = P.dup_index[pair];
    const int x = (strip % g.pw_sx) * 32 + lane;
    const int y0 = (strip / g.pw_sx) * kPR, y1 = min(y0 + kPR, g.H);
    if (x >= g.W) return;
    for (int y = y0; y < y1; ++y) {
        const unsigned q = (unsigned)(y * g.W + x);
        float2 u = U[q];
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

// ------------------------------------------------------------------------------------------- the super-step
#ifndef TEEFLOW_MIN_CTAS
#define TEEFLOW_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kThreads, TEEFLOW_MIN_CTAS)
tvl1_step_kernel(const __grid_constant__ EngineParams P, const int parity) {
    __shared__ int s_prefix[kMaxSlots + 1];
    __shared__ float4 s_cubic[32];

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

#if TEEFLOW_DYNAMIC_ITEMS
    // dynamic distribution: every warp pulls the next strip from a per-launch counter, so strips of unequal cost
    // (inner / median / warp phases mix in one launch) balance out; the counter of the other parity is re-armed
    if (blockIdx.x == 0 && tid == 0) P.item_counter[parity ^ 1] = 0;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(P.item_counter + parity, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
#else
    const int n_warps = gridDim.x * kWarpsPerCta;
    // CTA-interleaved item order: the 8 warps of a CTA take 8 consecutive strips (shared cache lines)
    for (int item = blockIdx.x * kWarpsPerCta + (tid >> 5); item < total; item += n_warps) {
#endif
        int lo = 0, hi = P.S;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] <= item) lo = mid; else hi = mid; }
        const int slot = P.slot0 + lo;            // absolute slot: planes, arrival counter, partials
        const int strip = item - s_prefix[lo];
        const Slot* sp = cur + lo;
        const int pair = sp->pair, phase = sp->phase, level = sp->level, ucur = sp->ucur, pcur = sp->pcur;
        const int n_items = s_prefix[lo + 1] - s_prefix[lo];

        double err = 0.0, aux = 0.0;
        switch (phase) {
            case PH_LEVEL_INIT: op_level_init(P, level, ucur, slot, strip, lane); break;
            case PH_WARP: op_warp(P, level, ucur, pair, slot, strip, lane, s_cubic); break;
            case PH_MEDIAN: op_median(P, level, ucur, slot, strip, lane); break;
            case PH_INNER: err = op_inner(P, level, ucur, pcur, slot, strip, lane); break;
            case PH_WASE: op_wase(P, ucur, slot, strip, lane, err, aux); break;
            case PH_FINAL: op_final(P, ucur, pair, sp->bg, slot, strip, lane); break;
            default: break;
        }
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            if (phase == PH_INNER) P.partial[(size_t)slot * P.max_tiles + strip] = err;
            if (phase == PH_WASE) {
                P.partial[(size_t)slot * P.max_tiles + 2 * strip] = err;
                P.partial[(size_t)slot * P.max_tiles + 2 * strip + 1] = aux;
            }
            __threadfence();
            const unsigned ticket = atomicAdd(P.arrive + slot, 1u);
            last = (ticket == (unsigned)n_items - 1u);
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            // last strip of this slot for this step: reduce the partials in strip order and advance the slot
            __threadfence();
            double e = 0.0, e2 = 0.0;
            if (phase == PH_INNER) {
                const double* part = P.partial + (size_t)slot * P.max_tiles;
                for (int t = lane; t < n_items; t += 32) e += __ldcg(part + t);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
            } else if (phase == PH_WASE) {
                const double* part = P.partial + (size_t)slot * P.max_tiles;
                for (int t = lane; t < n_items; t += 32) { e += __ldcg(part + 2 * t); e2 += __ldcg(part + 2 * t + 1); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    e += __shfl_down_sync(0xffffffffu, e, o);
                    e2 += __shfl_down_sync(0xffffffffu, e2, o);
                }
            }
            if (lane == 0) {
                Slot n = *sp;
                if (n.phase == PH_WASE) {
                    n.bg = (float)(e / e2);          // 0/0 -> NaN, like np.mean of an empty selection
                    P.bg_out[n.pair] = n.bg;
                    n.phase = PH_FINAL;
                } else if (n.phase == PH_FINAL) {
                    int* co = P.counters_out + (size_t)n.pair * kMaxLevels * 3;
                    for (int l = 0; l < kMaxLevels; ++l) { co[l * 3] = n.cnt[l][0]; co[l * 3 + 1] = n.cnt[l][1]; co[l * 3 + 2] = n.cnt[l][2]; }
                    const int next = atomicAdd(P.next_pair, 1);
                    if (next < P.n_pairs) start_pair(P, n, next);
                    else { n.pair = -1; n.phase = PH_IDLE; }
                    __threadfence();
                    atomicAdd(P.pairs_done, 1);
                } else {
                    advance_slot(P, n, e);
                }
                nxt[lo] = n;
                P.arrive[slot] = 0u;
            }
        }
    }
}

}  // namespace teeflow
