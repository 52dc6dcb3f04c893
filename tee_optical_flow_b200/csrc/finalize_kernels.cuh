// finalize_kernels.cuh -- kernels around the solver: WASE weight map and the masked radial / longitudinal
// decomposition with its per-frame reductions (reference: optical_flow/analysis.py:89-327,
// cardiac_cycle_detection.py:100-116, calculate_optical_flow.py:649-652).
//
// Exactness: order statistics are found by radix select on order-preserving keys (no sorting library, no
// approximation), the percentile interpolation and the histogram binning replay numpy's arithmetic, and
// cartToPolar replays OpenCV's FMA polynomial (oracle/probe_polar_percentile.py pins all three).
#pragma once
#include <cuda_fp16.h>
#include "tvl1_device.cuh"

namespace teeflow {

// ---------------------------------------------------------------------------------------------- WASE weights
// w[y,x,c] = sum_n bkgd[n,y,x,c]  (calculate_optical_flow.py:650: mask_dict['bkgd'] holds ALL frames)
__global__ void wase_weights_kernel(const uint8_t* __restrict__ bkgd, int n_frames, int n_elem /* H*W*2 */,
                                    float* __restrict__ w) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += gridDim.x * blockDim.x) {
        int c = 0;
        for (int n = 0; n < n_frames; ++n) c += bkgd[(size_t)n * n_elem + i] != 0;
        w[i] = (float)c;
    }
}

// ------------------------------------------------------------------------------------------------ frame prep
// The `no_saliency=True` input stage of the pair loop (calculate_optical_flow.py:588): img2uint8(rgb2gray(frame))
//   gray = R/255*0.2125 + G/255*0.7154 + B/255*0.0721 (float64, skimage.color.rgb2gray on img_as_float input)
//   out  = img_as_ubyte((gray - min) / max) = clip(rint(((gray - min) / max) * 255))   (optical_flow_utils.py:30-31;
//          sic: divides by the frame maximum, not by the range)
__device__ __forceinline__ double rgb_to_gray(uchar3 p) {
    const double r = (double)p.x / 255.0, g = (double)p.y / 255.0, b = (double)p.z / 255.0;
    return (r * 0.2125 + g * 0.7154) + b * 0.0721;
}
__global__ void prep_minmax_kernel(const uint8_t* __restrict__ rgb, int npx, unsigned long long* __restrict__ mm /* [N][2] keys */) {
    const int f = blockIdx.y;
    const uint8_t* src = rgb + (size_t)f * npx * 3;
    double mn = 1e300, mx = -1e300;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        const double g = rgb_to_gray(make_uchar3(src[3 * i], src[3 * i + 1], src[3 * i + 2]));
        mn = fmin(mn, g); mx = fmax(mx, g);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        // gray >= 0: the IEEE bit pattern of a non-negative double is order-preserving as an unsigned integer
        atomicMin(&mm[2 * f], (unsigned long long)__double_as_longlong(mn));
        atomicMax(&mm[2 * f + 1], (unsigned long long)__double_as_longlong(mx));
    }
}
__global__ void prep_quantize_kernel(const uint8_t* __restrict__ rgb, int npx, const unsigned long long* __restrict__ mm,
                                     uint8_t* __restrict__ out) {
    const int f = blockIdx.y;
    const uint8_t* src = rgb + (size_t)f * npx * 3;
    const double mn = __longlong_as_double((long long)mm[2 * f]), mx = __longlong_as_double((long long)mm[2 * f + 1]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        const double g = rgb_to_gray(make_uchar3(src[3 * i], src[3 * i + 1], src[3 * i + 2]));
        double v = rint(((g - mn) / mx) * 255.0);
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);      // NaN (all-black frame: 0/0) -> 0 like astype(uint8) of nan -> 0
        out[(size_t)f * npx + i] = (uint8_t)(v == v ? v : 0.0);
    }
}

// ------------------------------------------------------------------------------------- order-preserving keys
__device__ __forceinline__ unsigned f32_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long f64_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

// cv::cartToPolar (radians) as built in OpenCV 4.x with FMA SIMD: magnitude sqrt(fma(x,x,y*y)); fastAtan2
// polynomial in degrees evaluated with FMAs, then * (float)(pi/180)
__device__ __forceinline__ void cart_to_polar(float x, float y, float& mag, float& ang) {
    mag = __fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(y, y)));
    const float p1 = 0.9997878412794807f * 57.29577951308232f, p3 = -0.3258083974640975f * 57.29577951308232f;
    const float p5 = 0.1555786518463281f * 57.29577951308232f, p7 = -0.04432655554792128f * 57.29577951308232f;
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    const float c = __fdiv_rn(mn, __fadd_rn(mx, 2.220446049250313e-16f));
    const float cc = __fmul_rn(c, c);
    float a = __fmaf_rn(p7, cc, p5);
    a = __fmaf_rn(a, cc, p3);
    a = __fmaf_rn(a, cc, p1);
    a = __fmul_rn(a, c);
    if (ay > ax) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    ang = __fmul_rn(a, 0.017453292519943295f);
}

struct FrameStats {
    // non-zero counts: mag, ang, rad, long == lengths of the frame's compacted value lists.  (cnt[0], cnt[1]) and
    // (cnt[2], cnt[3]) are each bumped by ONE 64-bit atomic (low word = first count; a frame has < 2^28 pixels)
    unsigned cnt[4];
    unsigned mag_min, mag_max, ang_min, ang_max;          // keys over ALL pixels (np.min / np.max of the arrays)
    unsigned long long rad_min, rad_max, long_min, long_max;
};
constexpr int kAngBins = 640;        // rint(ang * 100) in [0, 628]

__global__ void analysis_init_kernel(FrameStats* st, unsigned* ang_hist, int nframes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nframes) {
        FrameStats s;
        for (int k = 0; k < 4; ++k) s.cnt[k] = 0;
        s.mag_min = s.ang_min = 0xffffffffu; s.mag_max = s.ang_max = 0u;
        s.rad_min = s.long_min = ~0ull; s.rad_max = s.long_max = 0ull;
        st[i] = s;
    }
    for (int k = i; k < nframes * kAngBins; k += gridDim.x * blockDim.x) ang_hist[k] = 0u;
}

// masked_arr = vel_array * mask (optical_flow_dataset.py:189-197); mag/ang (analysis.py:232-236);
// radial unit grid + projections in float64 (analysis.py:89-163).  grid = (chunks, nframes)
// Every consumer (percentiles, histograms, the angle mode) works on flat[flat != 0] of a frame, so only the NON-ZERO
// values are kept: frame f's values of a quantity are appended, in no particular order (all consumers are order
// independent), to the list that starts at element f * npx of its array; stats[f].cnt[] holds the list lengths.
// One warp-aggregated 64-bit atomic reserves the places of two quantities at once, and warps that see only zeros
// (outside the mask) touch nothing: the pass writes ~24 bytes per MASKED pixel instead of per pixel.
__global__ void __launch_bounds__(256)
analysis_values_kernel(const __half2* __restrict__ flow16, const uint8_t* __restrict__ mask,
                       const double* __restrict__ centroids, int H, int W, float* __restrict__ mag_out,
                       float* __restrict__ ang_out, double* __restrict__ rad_out, double* __restrict__ long_out,
                       FrameStats* __restrict__ stats, unsigned* __restrict__ ang_hist) {
    __shared__ unsigned s_hist[kAngBins];
    __shared__ unsigned s_k32[4];
    __shared__ unsigned long long s_k64[4];
    const int f = blockIdx.y;
    const int npx = H * W;
    for (int k = threadIdx.x; k < kAngBins; k += blockDim.x) s_hist[k] = 0u;
    if (threadIdx.x == 0) { s_k32[0] = 0xffffffffu; s_k32[1] = 0u; s_k32[2] = 0xffffffffu; s_k32[3] = 0u;
                            s_k64[0] = ~0ull; s_k64[1] = 0ull; s_k64[2] = ~0ull; s_k64[3] = 0ull; }
    __syncthreads();
    const double cH = centroids[2 * f], cW = centroids[2 * f + 1];
    const size_t fo = (size_t)f * npx;
    FrameStats* st = stats + f;
    unsigned long long* cnt_ma = reinterpret_cast<unsigned long long*>(&st->cnt[0]);
    unsigned long long* cnt_rl = reinterpret_cast<unsigned long long*>(&st->cnt[2]);
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    unsigned mag_mn = 0xffffffffu, mag_mx = 0u, ang_mn = 0xffffffffu, ang_mx = 0u;
    unsigned long long rad_mn = ~0ull, rad_mx = 0ull, long_mn = ~0ull, long_mx = 0ull;
    for (int i0 = blockIdx.x * blockDim.x; i0 < npx; i0 += gridDim.x * blockDim.x) {    // block-uniform trip count
        const int i = i0 + (int)threadIdx.x;
        const bool in = i < npx;
        float mg = 0.f, an = 0.f;
        double rad = 0.0, lng = 0.0;
        if (in) {
            const int r = i / W, c = i - r * W;
            const float2 v = __half22float2(flow16[fo + i]);
            const uchar2 m = reinterpret_cast<const uchar2*>(mask)[fo + i];
            const float fx = v.x * (m.x ? 1.f : 0.f), fy = v.y * (m.y ? 1.f : 0.f);
            cart_to_polar(fx, fy, mg, an);
            const double v0 = cH - (double)r, v1 = cW - (double)c;
            const double nrm = sqrt(v0 * v0 + v1 * v1);
            double u0 = v0 / nrm, u1 = v1 / nrm;
            if (u0 != u0) u0 = 0.0;                 // np.nan_to_num(vec / norm, nan=0): 0/0 at the centroid
            if (u1 != u1) u1 = 0.0;
            rad = (double)fx * u0 + (double)fy * u1;
            lng = (double)fx * u1 + (double)fy * (-u0);
            if (an != 0.f) {
                const int key = (int)rintf(__fmul_rn(an, 100.f));     // np.round(ang, 2) == rint(ang*100)/100
                if (key != 0) atomicAdd(&s_hist[min(key, kAngBins - 1)], 1u);
            }
            const unsigned km = f32_key(mg), ka = f32_key(an);
            mag_mn = min(mag_mn, km); mag_mx = max(mag_mx, km); ang_mn = min(ang_mn, ka); ang_mx = max(ang_mx, ka);
            const unsigned long long kr = f64_key(rad), kl = f64_key(lng);
            rad_mn = min(rad_mn, kr); rad_mx = max(rad_mx, kr); long_mn = min(long_mn, kl); long_mx = max(long_mx, kl);
        }
        // append the non-zero values to the frame's lists (NaN != 0 is true, like numpy's flat != 0)
        const bool nz_m = in && mg != 0.f, nz_a = in && an != 0.f, nz_r = in && rad != 0.0, nz_l = in && lng != 0.0;
        const unsigned bm = __ballot_sync(0xffffffffu, nz_m), ba = __ballot_sync(0xffffffffu, nz_a);
        const unsigned br = __ballot_sync(0xffffffffu, nz_r), bl = __ballot_sync(0xffffffffu, nz_l);
        if (bm | ba) {
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(cnt_ma, ((unsigned long long)__popc(ba) << 32) | (unsigned long long)__popc(bm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (nz_m) mag_out[fo + (unsigned)base + __popc(bm & lt)] = mg;
            if (nz_a) ang_out[fo + (unsigned)(base >> 32) + __popc(ba & lt)] = an;
        }
        if (br | bl) {
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(cnt_rl, ((unsigned long long)__popc(bl) << 32) | (unsigned long long)__popc(br));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (nz_r) rad_out[fo + (unsigned)base + __popc(br & lt)] = rad;
            if (nz_l) long_out[fo + (unsigned)(base >> 32) + __popc(bl & lt)] = lng;
        }
    }
    atomicMin(&s_k32[0], mag_mn); atomicMax(&s_k32[1], mag_mx); atomicMin(&s_k32[2], ang_mn); atomicMax(&s_k32[3], ang_mx);
    atomicMin(&s_k64[0], rad_mn); atomicMax(&s_k64[1], rad_mx); atomicMin(&s_k64[2], long_mn); atomicMax(&s_k64[3], long_mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMin(&st->mag_min, s_k32[0]); atomicMax(&st->mag_max, s_k32[1]);
        atomicMin(&st->ang_min, s_k32[2]); atomicMax(&st->ang_max, s_k32[3]);
        atomicMin(&st->rad_min, s_k64[0]); atomicMax(&st->rad_max, s_k64[1]);
        atomicMin(&st->long_min, s_k64[2]); atomicMax(&st->long_max, s_k64[3]);
    }
    for (int k = threadIdx.x; k < kAngBins; k += blockDim.x)
        if (s_hist[k]) atomicAdd(&ang_hist[(size_t)f * kAngBins + k], s_hist[k]);
}

// Exact order statistics of the NON-ZERO values of one frame (its compacted list) by MSB-first radix select (8-bit digits).
// One CTA per (frame, target group): quantity 0 = mag (float32), 1 = rad, 2 = long (float64); up to 4 target
// ranks (0-based, among the non-zero values in ascending order; rank < 0 = unused).
constexpr int kSelTargets = 4;
__global__ void __launch_bounds__(1024)
radix_select_kernel(const float* __restrict__ mag, const double* __restrict__ rad, const double* __restrict__ lng,
                    int npx, const FrameStats* __restrict__ stats, const long long* __restrict__ ranks /* [nframes][3][4] */,
                    unsigned long long* __restrict__ out_keys /* [nframes][3][4] */) {
    __shared__ unsigned s_hist[kSelTargets][256];
    __shared__ unsigned long long s_prefix[kSelTargets];
    __shared__ long long s_rank[kSelTargets];
    const int f = blockIdx.x, qn = blockIdx.y;
    const bool is64 = qn != 0;
    const int nbits = is64 ? 64 : 32;
    const size_t fo = (size_t)f * npx;
    const int n = (int)stats[f].cnt[qn == 0 ? 0 : qn + 1];                                     // list length: mag, rad, long
    const unsigned long long zero_pos = is64 ? 0x8000000000000000ull : 0x80000000ull;          // key of +0
    const unsigned long long zero_neg = is64 ? 0x7fffffffffffffffull : 0x7fffffffull;          // key of -0
    if (threadIdx.x < kSelTargets) {
        s_prefix[threadIdx.x] = 0ull;
        s_rank[threadIdx.x] = ranks[((size_t)f * 3 + qn) * kSelTargets + threadIdx.x];
    }
    __syncthreads();
    for (int shift = nbits - 8; shift >= 0; shift -= 8) {
        for (int k = threadIdx.x; k < kSelTargets * 256; k += blockDim.x) (&s_hist[0][0])[k] = 0u;
        __syncthreads();
        unsigned long long pre[kSelTargets];
        bool act[kSelTargets];
#pragma unroll
        for (int t = 0; t < kSelTargets; ++t) { pre[t] = s_prefix[t]; act[t] = s_rank[t] >= 0; }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned long long key;
            if (qn == 0) key = (unsigned long long)f32_key(mag[fo + i]);
            else key = f64_key(qn == 1 ? rad[fo + i] : lng[fo + i]);
            if (key == zero_pos || key == zero_neg) continue;       // flat[flat != 0]
            const unsigned digit = (unsigned)(key >> shift) & 0xffu;
            const unsigned long long hi = (shift + 8 >= 64) ? 0ull : (key >> (shift + 8));
#pragma unroll
            for (int t = 0; t < kSelTargets; ++t) {
                const unsigned long long phi = (shift + 8 >= 64) ? 0ull : (pre[t] >> (shift + 8));
                if (act[t] && hi == phi) atomicAdd(&s_hist[t][digit], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < kSelTargets && s_rank[threadIdx.x] >= 0) {
            const int t = threadIdx.x;
            long long r = s_rank[t];
            int d = 0;
            for (; d < 255; ++d) {
                const long long c = (long long)s_hist[t][d];
                if (r < c) break;
                r -= c;
            }
            s_rank[t] = r;
            s_prefix[t] |= ((unsigned long long)d) << shift;
        }
        __syncthreads();
    }
    if (threadIdx.x < kSelTargets)
        out_keys[((size_t)f * 3 + qn) * kSelTargets + threadIdx.x] = s_prefix[threadIdx.x];
}

// np.histogram(flat_nonzero, bins=nbins, range=(first, last)) with numpy's uniform-bin arithmetic
// (lib/_histograms_impl.py): index = int(((v - first) / (last - first)) * nbins), fixed up against the edges.
// T = float (mag, ang) or double (rad, long).  grid = (chunks, nframes)
template <typename T>
__global__ void __launch_bounds__(256)
np_histogram_kernel(const T* __restrict__ vals, int npx, const FrameStats* __restrict__ stats, int quantity,
                    const T* __restrict__ edges, int nbins, T first, T last,
                    unsigned long long* __restrict__ freq /* [nframes][nbins] */) {
    extern __shared__ unsigned s_bins[];
    const int f = blockIdx.y;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) s_bins[k] = 0u;
    __syncthreads();
    const T denom = last - first;
    const size_t fo = (size_t)f * npx;
    const int n = (int)stats[f].cnt[quantity];       // the frame's list of non-zero values
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const T v = vals[fo + i];
        if (v == (T)0 || !(v >= first) || !(v <= last)) continue;
        const T fi = ((v - first) / denom) * (T)nbins;
        int idx = (int)fi;
        if (idx == nbins) idx -= 1;
        if (v < edges[idx]) idx -= 1;
        if (idx != nbins - 1 && v >= edges[idx + 1]) idx += 1;
        atomicAdd(&s_bins[idx], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nbins; k += blockDim.x)
        if (s_bins[k]) atomicAdd(&freq[(size_t)f * nbins + k], (unsigned long long)s_bins[k]);
}

}  // namespace teeflow

// ============================================================================================================
// Connected components (union-find with atomicMin) and the mask post-processing built on it
// (calculate_optical_flow.py:91-182 clean_mask / moving_avg_mask; analysis.py:39-86 calc_AV_centroid).
// Labels are the smallest pixel index of each component, i.e. components are ordered like skimage.measure.label /
// scipy.ndimage.label (raster order of their first pixel).
// ============================================================================================================
namespace teeflow {

// parents only ever decrease; reads are volatile because other threads hook roots concurrently
__device__ __forceinline__ int uf_find(int* Lp, int i) {
    volatile int* L = Lp;
    int p = L[i];
    while (p != i) {
        const int gp = L[p];
        if (gp != p) atomicMin(&Lp[i], gp);   // path halving; atomicMin keeps parents monotonically decreasing
        i = p; p = gp;
    }
    return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a); b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);     // hook the larger root under the smaller one
        if (old == b) return;
        b = old;
    }
}

// moving_avg_mask (calculate_optical_flow.py:91-111): frames padded [first, 0..N-1, last, last]; window mean > thr
__global__ void mask_vote_kernel(const uint8_t* __restrict__ cls, int N, int npx, int class_id, int window,
                                 double threshold, uint8_t* __restrict__ out) {
    const int f = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        int cnt = 0;
        for (int k = 0; k < window; ++k) {
            int t = f + k - 1;                                   // index into the un-padded sequence
            t = t < 0 ? 0 : (t > N - 1 ? N - 1 : t);
            cnt += cls[(size_t)t * npx + i] == class_id;
        }
        out[(size_t)f * npx + i] = ((double)cnt / (double)window) > threshold;
    }
}

// L[i] = i for pixels whose value == fg, else -1.  img may be a multi-channel bool array (pixel stride `ch`).
__global__ void ccl_init_kernel(const uint8_t* __restrict__ img, int npx, int ch, int fg, int* __restrict__ L) {
    const int f = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
        L[(size_t)f * npx + i] = ((img[((size_t)f * npx + i) * ch] != 0) == (fg != 0)) ? i : -1;
}
__global__ void ccl_merge_kernel(int* Lall, int H, int W, int conn8) {
    const int f = blockIdx.y;
    int* L = Lall + (size_t)f * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        if (L[i] < 0) continue;
        const int y = i / W, x = i - y * W;
        if (x > 0 && L[i - 1] >= 0) uf_union(L, i, i - 1);
        if (y > 0 && L[i - W] >= 0) uf_union(L, i, i - W);
        if (conn8 && y > 0) {
            if (x > 0 && L[i - W - 1] >= 0) uf_union(L, i, i - W - 1);
            if (x < W - 1 && L[i - W + 1] >= 0) uf_union(L, i, i - W + 1);
        }
    }
}
// flatten + per-root statistics: area, (optionally) coordinate sums and border contact
__global__ void ccl_stats_kernel(int* Lall, int H, int W, int* __restrict__ area,
                                 unsigned long long* __restrict__ sum_r, unsigned long long* __restrict__ sum_c,
                                 int* __restrict__ touches_border) {
    const int f = blockIdx.y;
    const size_t fo = (size_t)f * H * W;
    int* L = Lall + fo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        if (L[i] < 0) continue;
        const int r = uf_find(L, i);
        L[i] = r;
        atomicAdd(&area[fo + r], 1);
        const int y = i / W, x = i - y * W;
        if (sum_r) { atomicAdd(&sum_r[fo + r], (unsigned long long)y); atomicAdd(&sum_c[fo + r], (unsigned long long)x); }
        if (touches_border && (y == 0 || x == 0 || y == H - 1 || x == W - 1)) touches_border[fo + r] = 1;
    }
}
// binary_fill_holes: background components (4-connectivity) that do not touch the image border become foreground
__global__ void fill_holes_kernel(uint8_t* __restrict__ m, const int* __restrict__ Lbg, const int* __restrict__ touches,
                                  int npx) {
    const int f = blockIdx.y;
    const size_t fo = (size_t)f * npx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        const int r = Lbg[fo + i];
        if (r >= 0 && !touches[fo + r]) m[fo + i] = 1;
    }
}
// remove_small_objects(min_size): drop components with area < min_size
__global__ void remove_small_kernel(uint8_t* __restrict__ m, const int* __restrict__ L, const int* __restrict__ area,
                                    int npx, int min_size) {
    const int f = blockIdx.y;
    const size_t fo = (size_t)f * npx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        const int r = L[fo + i];
        if (r >= 0 && area[fo + r] < min_size) m[fo + i] = 0;
    }
}
// largest component per frame (first one in label order among ties): packed key (area << 32) | ~root, atomicMax
__global__ void largest_component_kernel(const int* __restrict__ L, const int* __restrict__ area, int npx,
                                         unsigned long long* __restrict__ best, int* __restrict__ n_comp) {
    const int f = blockIdx.y;
    const size_t fo = (size_t)f * npx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) {
        if (L[fo + i] == i) {                       // a root
            atomicAdd(&n_comp[f], 1);
            const unsigned long long key = ((unsigned long long)(unsigned)area[fo + i] << 32) | (unsigned)(~(unsigned)i);
            atomicMax(&best[f], key);
        }
    }
}

// (area, sum_r, sum_c) of each frame's winning root, so that the host reads one small array
__global__ void largest_component_gather_kernel(const unsigned long long* __restrict__ best, const int* __restrict__ n_comp,
                                                const unsigned long long* __restrict__ sum_r,
                                                const unsigned long long* __restrict__ sum_c, int npx, int nf,
                                                unsigned long long* __restrict__ out /* [nf][3] */) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    if (n_comp[f] == 0) { out[3 * f] = out[3 * f + 1] = out[3 * f + 2] = 0ull; return; }
    const unsigned root = ~(unsigned)(best[f] & 0xffffffffull);
    out[3 * f] = best[f] >> 32;
    out[3 * f + 1] = sum_r[(size_t)f * npx + root];
    out[3 * f + 2] = sum_c[(size_t)f * npx + root];
}

}  // namespace teeflow
