// DRAM-efficiency probe for the inner iteration's access pattern (nvcc -O3 -arch=sm_100a tools/stream_probe.cu).
// Warps walk strips of the row-interleaved slot layout ([slot][row][8 planes][PITCH] float2) exactly like op_inner:
// per row 5 plane segments are read and 3 written, one row ahead in registers.  VEC = 1: 8 bytes per lane (256-byte
// requests, the engine's pattern); VEC = 2: 16 bytes per lane (512-byte requests, two pixels per lane).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int PITCH = 1024, PLANES = 8, H = 600, W = 800, ROWS = 20;
template <int VEC> struct V;
template <> struct V<1> { using T = float2; };
template <> struct V<2> { using T = float4; };
template <int VEC>
__global__ void __launch_bounds__(256, 4) probe(float2* base, int slots, int* counter) {
    using T = typename V<VEC>::T;
    const int lane = threadIdx.x & 31;
    const int cols = 32 * VEC, sx = (W + cols - 1) / cols, sy = (H + ROWS - 1) / ROWS;
    const int per_slot = sx * sy, total = per_slot * slots;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int slot = item / per_slot, s = item % per_slot;
        const int x = (s % sx) * cols + lane * VEC, y0 = (s / sx) * ROWS, y1 = min(y0 + ROWS, H);
        if (x >= W) continue;
        float2* row = base + ((size_t)slot * (H + 2) * PLANES + (size_t)y0 * PLANES) * PITCH + x;
        T a = *(const T*)(row + 0 * PITCH), b = *(const T*)(row + 2 * PITCH), c = *(const T*)(row + 4 * PITCH),
          d = *(const T*)(row + 6 * PITCH), e = *(const T*)(row + 7 * PITCH);
        for (int y = y0; y < y1; ++y) {
            float2* nx = row + PLANES * PITCH;
            T a2 = *(const T*)(nx + 0 * PITCH), b2 = *(const T*)(nx + 2 * PITCH), c2 = *(const T*)(nx + 4 * PITCH),
              d2 = *(const T*)(nx + 6 * PITCH), e2 = *(const T*)(nx + 7 * PITCH);
            T o1 = a, o2 = b, o3 = c;
            o1.x += d.x * e.x; o2.x += d.y * e.y; o3.x += a.y;
            *(T*)(row + 1 * PITCH) = o1; *(T*)(row + 3 * PITCH) = o2; *(T*)(row + 5 * PITCH) = o3;
            a = a2; b = b2; c = c2; d = d2; e = e2; row = nx;
        }
    }
}
// the engine's exact geometry: 31 owned columns per warp (lane 31 reads a halo column, stores nothing), 248-byte
// row segments at 248-byte offsets (partial sectors at both ends), one extra row read per strip; WORK dependent
// dummy FMAs per row stand in for the arithmetic between the loads (0: none, ~180: the real kernel)
template <int WORK>
__global__ void __launch_bounds__(256, 4) probe31(float2* base, int slots, int* counter) {
    const int lane = threadIdx.x & 31;
    const int sx = (W + 30) / 31, sy = (H + ROWS - 1) / ROWS;
    const int per_slot = sx * sy, total = per_slot * slots;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int slot = item / per_slot, s = item % per_slot;
        const int x = min((s % sx) * 31 + lane, W - 1), y0 = (s / sx) * ROWS, y1 = min(y0 + ROWS, H);
        const bool owner = lane < 31 && (s % sx) * 31 + lane < W;
        float2* row = base + ((size_t)slot * (H + 2) * PLANES + (size_t)y0 * PLANES) * PITCH + x;
        float2 a = row[0], b = row[2 * PITCH], c = row[4 * PITCH], d = row[6 * PITCH], e = row[7 * PITCH];
        for (int y = y0; y < y1; ++y) {
            float2* nx = row + PLANES * PITCH;
            const float2 a2 = nx[0], b2 = nx[2 * PITCH], c2 = nx[4 * PITCH], d2 = nx[6 * PITCH], e2 = nx[7 * PITCH];
            float2 o1 = a, o2 = b, o3 = c;
            float t = d.x;
#pragma unroll
            for (int k = 0; k < WORK; ++k) t = fmaf(t, e.x, d.y);
            o1.x += t; o2.x += d.y * e.y; o3.x += a.y;
            if (owner) { row[1 * PITCH] = o1; row[3 * PITCH] = o2; row[5 * PITCH] = o3; }
            a = a2; b = b2; c = c2; d = d2; e = e2; row = nx;
        }
    }
}
template <int WORK>
static void run31(float2* buf, int slots, int* counter, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(counter, 0, 4);
        cudaEventRecord(e0);
        probe31<WORK><<<148 * 4, 256>>>(buf, slots, counter);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double bytes = (double)slots * H * W * 64.0;
    printf("%s: %.1f us, %.0f GB/s algorithmic (64 B/px)\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9);
}
template <int VEC>
static void run(float2* buf, int slots, int* counter, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(counter, 0, 4);
        cudaEventRecord(e0);
        probe<VEC><<<148 * 4, 256>>>(buf, slots, counter);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double bytes = (double)slots * H * W * 64.0;
    printf("%s: %.1f us, %.0f GB/s algorithmic (64 B/px), %s\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int slots = 63;
    float2* buf; int* counter;
    const size_t n = (size_t)slots * (H + 2) * PLANES * PITCH;
    cudaMalloc(&buf, n * sizeof(float2)); cudaMemset(buf, 0, n * sizeof(float2)); cudaMalloc(&counter, 4);
    run<1>(buf, slots, counter, "8 B per lane (256 B requests)");
    run<2>(buf, slots, counter, "16 B per lane (512 B requests)");
    run31<0>(buf, slots, counter, "engine geometry (31 columns + halo lane, +1 row), no arithmetic");
    run31<60>(buf, slots, counter, "engine geometry, 60 dependent FMAs per row");
    run31<120>(buf, slots, counter, "engine geometry, 120 dependent FMAs per row");
    run31<180>(buf, slots, counter, "engine geometry, 180 dependent FMAs per row");
    return 0;
}
