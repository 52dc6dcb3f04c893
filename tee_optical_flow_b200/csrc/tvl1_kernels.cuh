// tvl1_kernels.cuh -- the kernels of the TV-L1 engine.
//
//  * pyramid kernels (once per FRAME, not per pair): u8/f32 -> f32, resize x0.8, centred gradient + pack.
//  * tvl1_step_kernel: one persistent launch = one "super-step" of the device-side scheduler.  Every slot
//    (a frame pair in flight) is in some phase (level-init / warp / median / inner / final); the work of all
//    slots is flattened into tile items, persistent CTAs grid-stride over them, and the CTA that finishes the
//    last tile of a slot reduces the slot's error partials (fixed order, float64) and advances its state
//    machine exactly like OpenCV's procOneScale control flow.  No host round trip per
//    iteration, per-pair exact early exit, continuous refill of finished slots.
#pragma once
#include "tvl1_device.cuh"

namespace teeflow {

constexpr int kTW = 64;        // tile width  (pixels)
constexpr int kTH = 16;        // tile height (pixels)
constexpr int kThreads = 256;  // threads per CTA

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
            if (s.level == 0) { s.phase = PH_FINAL; return; }
            s.level--; s.phase = PH_LEVEL_INIT; return;
        }
    }
}

__device__ inline void start_pair(const EngineParams& P, Slot& s, int pair) {
    s.pair = pair; s.phase = PH_LEVEL_INIT; s.level = P.L - 1; s.warp = 0; s.n_outer = 0; s.n_inner = 0;
    s.ucur = 0; s.pcur = 0; s.error = FLT_MAX;
    for (int l = 0; l < kMaxLevels; ++l) { s.cnt[l][0] = 0; s.cnt[l][1] = 0; s.cnt[l][2] = 0; }
}

__device__ __forceinline__ int tiles_of(const EngineParams& P, const Slot& s) {
    return (s.phase == PH_IDLE || s.pair < 0) ? 0 : P.lv[s.level].ntiles;
}

// ------------------------------------------------------------------------------------------------ tile ops
// PH_LEVEL_INIT
__device__ __forceinline__ void op_level_init(const EngineParams& P, const Slot& st, int slot, int tx0, int ty0) {
    const LevelGeom g = P.lv[st.level];
    const size_t base = (size_t)slot * P.slot_px;
    const bool coarsest = (st.level == P.L - 1);
    float2* Ud = P.U[coarsest ? 0 : (st.ucur ^ 1)] + base;
    const float2* Us = P.U[st.ucur] + base;
    float2* PXd = P.PX[0] + base;
    float2* PYd = P.PY[0] + base;
    const int cH = coarsest ? 0 : P.lv[st.level + 1].H, cW = coarsest ? 0 : P.lv[st.level + 1].W;
    for (int i = threadIdx.x; i < kTW * kTH; i += kThreads) {
        const int x = tx0 + (i % kTW), y = ty0 + (i / kTW);
        if (x >= g.W || y >= g.H) continue;
        const size_t q = (size_t)y * g.W + x;
        float2 u = make_float2(0.f, 0.f);
        if (!coarsest) {
            int x0, x1, y0, y1; float a0, a1, b0, b1;
            lin_coeff_x(x, g.up_sx, cW, x0, x1, a0, a1);
            lin_coeff_y(y, g.up_sy, cH, y0, y1, b0, b1);
            const float2 s00 = Us[(size_t)y0 * cW + x0], s01 = Us[(size_t)y0 * cW + x1];
            const float2 s10 = Us[(size_t)y1 * cW + x0], s11 = Us[(size_t)y1 * cW + x1];
            const float r0x = s00.x * a0 + s01.x * a1, r1x = s10.x * a0 + s11.x * a1;
            const float r0y = s00.y * a0 + s01.y * a1, r1y = s10.y * a0 + s11.y * a1;
            u.x = (r0x * b0 + r1x * b1) * P.up_mul;
            u.y = (r0y * b0 + r1y * b1) * P.up_mul;
        }
        Ud[q] = u;
        PXd[q] = make_float2(0.f, 0.f);
        PYd[q] = make_float2(0.f, 0.f);
    }
}

// PH_WARP: buildFlowMap + remap(I1, I1x, I1y; INTER_CUBIC) + calcGradRho
__device__ __forceinline__ void op_warp(const EngineParams& P, const Slot& st, int slot, int tx0, int ty0,
                                        const float4* s_cubic) {
    const LevelGeom g = P.lv[st.level];
    const size_t base = (size_t)slot * P.slot_px;
    const float2* U = P.U[st.ucur] + base;
    float4* COEF = P.COEF + base;
    const int fa = P.pair_a[st.pair], fb = P.pair_b[st.pair];
    const float* I0 = P.pyrI + (size_t)fa * P.frame_pyr_stride + g.pyr_off;
    const float4* G1 = P.pyrG + (size_t)fb * P.frame_pyr_stride + g.pyr_off;
    for (int i = threadIdx.x; i < kTW * kTH; i += kThreads) {
        const int x = tx0 + (i % kTW), y = ty0 + (i / kTW);
        if (x >= g.W || y >= g.H) continue;
        const size_t q = (size_t)y * g.W + x;
        const float2 u = U[q];
        const float mx = (float)x + u.x, my = (float)y + u.y;
        const float3 w = remap_cubic3(G1, g.H, g.W, mx, my, s_cubic);
        const float Ix2 = w.y * w.y, Iy2 = w.z * w.z;
        float4 c;
        c.x = w.y; c.y = w.z;
        c.z = Ix2 + Iy2;
        c.w = (w.x - w.y * u.x - w.z * u.y - I0[q]);
        COEF[q] = c;
    }
}

// PH_MEDIAN: medianBlur(u1, ksize), medianBlur(u2, ksize) with BORDER_REPLICATE
__device__ __forceinline__ void op_median(const EngineParams& P, const Slot& st, int slot, int tx0, int ty0) {
    const LevelGeom g = P.lv[st.level];
    const size_t base = (size_t)slot * P.slot_px;
    const float2* Us = P.U[st.ucur] + base;
    float2* Ud = P.U[st.ucur ^ 1] + base;
    for (int i = threadIdx.x; i < kTW * kTH; i += kThreads) {
        const int x = tx0 + (i % kTW), y = ty0 + (i / kTW);
        if (x >= g.W || y >= g.H) continue;
        float2 out;
        if (P.median == 5) {
            float v[25], w[25];
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy) {
                const int yy = clampi(y + dy, 0, g.H - 1);
#pragma unroll
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xx = clampi(x + dx, 0, g.W - 1);
                    const float2 t = Us[(size_t)yy * g.W + xx];
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
                    const float2 t = Us[(size_t)yy * g.W + xx];
                    v[(dy + 1) * 3 + dx + 1] = t.x;
                    w[(dy + 1) * 3 + dx + 1] = t.y;
                }
            }
            out.x = median9(v);
            out.y = median9(w);
        }
        Ud[(size_t)y * g.W + x] = out;
    }
}

// PH_INNER: one primal-dual iteration (estimateV, divergence, estimateU, forwardGradient, estimateDualVariables)
// fused in one pass.  u_new is needed at (x,y), (x+1,y), (x,y+1) for the dual update, so it is computed on the tile
// plus a one-pixel right/bottom halo into shared memory; p_old is read with a one-pixel left/top halo.
__device__ __forceinline__ double op_inner(const EngineParams& P, const Slot& st, int slot, int tx0, int ty0,
                                           float2* s_un) {
    const LevelGeom g = P.lv[st.level];
    const size_t base = (size_t)slot * P.slot_px;
    const float2* __restrict__ U = P.U[st.ucur] + base;
    float2* __restrict__ Un = P.U[st.ucur ^ 1] + base;
    const float2* __restrict__ PX = P.PX[st.pcur] + base;
    const float2* __restrict__ PY = P.PY[st.pcur] + base;
    float2* __restrict__ PXn = P.PX[st.pcur ^ 1] + base;
    float2* __restrict__ PYn = P.PY[st.pcur ^ 1] + base;
    const float4* __restrict__ COEF = P.COEF + base;
    const float l_t = P.l_t, theta = P.theta, taut = P.taut;
    constexpr int RW = kTW + 1, RH = kTH + 1;
    double err = 0.0;
    // phase 1: u_new on the extended region
    for (int i = threadIdx.x; i < RW * RH; i += kThreads) {
        const int lx = i % RW, ly = i / RW;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= g.W || y >= g.H) continue;
        const size_t q = (size_t)y * g.W + x;
        const float2 u = U[q];
        const float4 c = COEF[q];
        // estimateV
        const float rho = c.w + (c.x * u.x + c.y * u.y);
        float d1 = 0.f, d2 = 0.f;
        if (rho < -l_t * c.z) { d1 = l_t * c.x; d2 = l_t * c.y; }
        else if (rho > l_t * c.z) { d1 = -l_t * c.x; d2 = -l_t * c.y; }
        else if (c.z > FLT_EPSILON) { const float fi = -rho / c.z; d1 = fi * c.x; d2 = fi * c.y; }
        const float v1 = u.x + d1, v2 = u.y + d2;
        // divergence of (p11,p12) and (p21,p22), backward differences
        const float2 px = PX[q], py = PY[q];
        float div1, div2;
        if (x > 0 && y > 0) {
            const float2 pxl = PX[q - 1], pyu = PY[q - g.W];
            div1 = (px.x - pxl.x) + (py.x - pyu.x);
            div2 = (px.y - pxl.y) + (py.y - pyu.y);
        } else if (y == 0 && x > 0) {
            const float2 pxl = PX[q - 1];
            div1 = px.x - pxl.x + py.x;
            div2 = px.y - pxl.y + py.y;
        } else if (x == 0 && y > 0) {
            const float2 pyu = PY[q - g.W];
            div1 = px.x + py.x - pyu.x;
            div2 = px.y + py.y - pyu.y;
        } else {
            div1 = px.x + py.x;
            div2 = px.y + py.y;
        }
        // estimateU
        float2 un;
        un.x = v1 + theta * div1;
        un.y = v2 + theta * div2;
        s_un[ly * RW + lx] = un;
        if (lx < kTW && ly < kTH) {
            Un[q] = un;
            const float t = (un.x - u.x) * (un.x - u.x) + (un.y - u.y) * (un.y - u.y);
            err += (double)t;
        }
    }
    __syncthreads();
    // phase 2: forwardGradient(u_new) + estimateDualVariables
    for (int i = threadIdx.x; i < kTW * kTH; i += kThreads) {
        const int lx = i % kTW, ly = i / kTW;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= g.W || y >= g.H) continue;
        const size_t q = (size_t)y * g.W + x;
        const float2 un = s_un[ly * RW + lx];
        float u1x = 0.f, u2x = 0.f, u1y = 0.f, u2y = 0.f;
        if (x < g.W - 1) { const float2 r = s_un[ly * RW + lx + 1]; u1x = r.x - un.x; u2x = r.y - un.y; }
        if (y < g.H - 1) { const float2 b = s_un[(ly + 1) * RW + lx]; u1y = b.x - un.x; u2y = b.y - un.y; }
        const float g1 = (float)sqrt((double)u1x * (double)u1x + (double)u1y * (double)u1y);
        const float g2 = (float)sqrt((double)u2x * (double)u2x + (double)u2y * (double)u2y);
        const float ng1 = 1.0f + taut * g1;
        const float ng2 = 1.0f + taut * g2;
        const float2 px = PX[q], py = PY[q];
        float2 pxn, pyn;
        pxn.x = (px.x + taut * u1x) / ng1;   // p11
        pyn.x = (py.x + taut * u1y) / ng1;   // p12
        pxn.y = (px.y + taut * u2x) / ng2;   // p21
        pyn.y = (py.y + taut * u2y) / ng2;   // p22
        PXn[q] = pxn;
        PYn[q] = pyn;
    }
    return err;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const unsigned short lo = __half_as_ushort(__float2half_rn(a));
    const unsigned short hi = __half_as_ushort(__float2half_rn(b));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}

// PH_FINAL: merge(u1,u2) * conversion_factor -> (H,W,2) f32 and/or f16 (calculate_optical_flow.py:600,403)
__device__ __forceinline__ void op_final(const EngineParams& P, const Slot& st, int slot, int tx0, int ty0) {
    const LevelGeom g = P.lv[0];
    const size_t base = (size_t)slot * P.slot_px;
    const float2* U = P.U[st.ucur] + base;
    const size_t npx = (size_t)g.H * g.W;
    const int o0 = P.out_index[st.pair], o1 = P.dup_index[st.pair];
    for (int i = threadIdx.x; i < kTW * kTH; i += kThreads) {
        const int x = tx0 + (i % kTW), y = ty0 + (i / kTW);
        if (x >= g.W || y >= g.H) continue;
        const size_t q = (size_t)y * g.W + x;
        float2 u = U[q];
        u.x = u.x * P.out_scale;
        u.y = u.y * P.out_scale;
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
__global__ void __launch_bounds__(kThreads)
tvl1_step_kernel(const __grid_constant__ EngineParams P, const int parity) {
    __shared__ int s_prefix[kMaxSlots + 1];
    __shared__ float4 s_cubic[32];
    __shared__ float2 s_un[(kTW + 1) * (kTH + 1)];
    __shared__ double s_red[kThreads / 32];
    __shared__ int s_last;

    const Slot* __restrict__ cur = P.slots[parity];
    Slot* __restrict__ nxt = P.slots[parity ^ 1];
    const int tid = threadIdx.x;

    if (tid < 32) s_cubic[tid] = cubic_coeffs(tid);
    for (int s = tid; s < P.S; s += kThreads) s_prefix[s + 1] = tiles_of(P, cur[s]);
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

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        // slot of this item: largest s with prefix[s] <= item  (uniform across the CTA)
        int lo = 0, hi = P.S;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] <= item) lo = mid; else hi = mid; }
        const int slot = lo;
        const int tile = item - s_prefix[slot];
        const Slot* sp = cur + slot;
        Slot st;
        st.pair = sp->pair; st.phase = sp->phase; st.level = sp->level; st.ucur = sp->ucur; st.pcur = sp->pcur;
        const LevelGeom& g = P.lv[st.level];
        const int tx0 = (tile % g.tiles_x) * kTW, ty0 = (tile / g.tiles_x) * kTH;

        double err = 0.0;
        switch (st.phase) {
            case PH_LEVEL_INIT: op_level_init(P, st, slot, tx0, ty0); break;
            case PH_WARP: op_warp(P, st, slot, tx0, ty0, s_cubic); break;
            case PH_MEDIAN: op_median(P, st, slot, tx0, ty0); break;
            case PH_INNER: err = op_inner(P, st, slot, tx0, ty0, s_un); break;
            case PH_FINAL: op_final(P, st, slot, tx0, ty0); break;
            default: break;
        }

        if (st.phase == PH_INNER) {
            // deterministic CTA reduction of the float64 error partial (fixed shuffle tree, fixed warp order)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) err += __shfl_down_sync(0xffffffffu, err, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = err;
        }
        __syncthreads();   // all tile work of this CTA is issued; s_red complete
        if (tid == 0) {
            if (st.phase == PH_INNER) {
                double e = 0.0;
                for (int w = 0; w < kThreads / 32; ++w) e += s_red[w];
                P.partial[(size_t)slot * P.max_tiles + tile] = e;
            }
            __threadfence();
            const unsigned ticket = atomicAdd(P.arrive + slot, 1u);
            s_last = (ticket == (unsigned)g.ntiles - 1u);
        }
        __syncthreads();
        if (s_last) {
            // last tile of this slot for this step: reduce the partials in tile order and advance the slot
            __threadfence();
            double e = 0.0;
            if (st.phase == PH_INNER) {
                const double* part = P.partial + (size_t)slot * P.max_tiles;
                for (int t = tid; t < g.ntiles; t += kThreads) e += __ldcg(part + t);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
                if ((tid & 31) == 0) s_red[tid >> 5] = e;
                __syncthreads();
                e = 0.0;
                if (tid == 0) for (int w = 0; w < kThreads / 32; ++w) e += s_red[w];
            }
            if (tid == 0) {
                Slot n = *sp;
                if (n.phase == PH_FINAL) {
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
                nxt[slot] = n;
                P.arrive[slot] = 0u;
            }
        }
        __syncthreads();   // s_last / s_red / s_un are reused by the next item
    }
}

}  // namespace teeflow
