// tma_probe.cu -- standalone check of the tensor-map boxes op_inner_tma uses (one warp, one box per launch).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int MODE>   // 0: map in global memory, 1: + fence.proxy.tensormap acquire, 2: map passed as a __grid_constant__ parameter
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int c0, int c1, int c2, int c3, unsigned bytes, unsigned long long* out, int n_out) {
    __shared__ __align__(128) unsigned long long buf[1024];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(buf);
    for (int i = threadIdx.x; i < 1024; i += 32) buf[i] = 0xdeadbeefull;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const CUtensorMap* tm = MODE == 2 ? &pmap : gmap;
    if (threadIdx.x == 0) {
        if (MODE == 1) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(d), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b) : "memory");
    }
    unsigned ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 2000000000ll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    __syncwarp();
    for (int i = threadIdx.x; i < n_out; i += 32) out[i] = ok ? buf[i] : 0xffffffffffffffffull;
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc_t enc = (enc_t)p;
    const int PITCH = 1024, PLANES = 8, H = 40, W = 100, S = 3;
    const size_t slot_elems = (size_t)(H + 2) * PLANES * PITCH;
    std::vector<unsigned long long> host(slot_elems * S);
    for (int s = 0; s < S; ++s) for (int y = 0; y < H + 2; ++y) for (int pl = 0; pl < PLANES; ++pl) for (int x = 0; x < PITCH; ++x)
        host[(size_t)s * slot_elems + ((size_t)y * PLANES + pl) * PITCH + x] = ((unsigned long long)s << 48) | ((unsigned long long)pl << 32) | ((unsigned long long)y << 16) | x;
    unsigned long long* dev; cudaMalloc(&dev, host.size() * 8); cudaMemcpy(dev, host.data(), host.size() * 8, cudaMemcpyHostToDevice);
    const cuuint64_t PB = PITCH * 8ull, ROWB = PB * PLANES, SLOTB = slot_elems * 8ull;
    alignas(64) CUtensorMap maps[3];
    const cuuint64_t dims[4] = {W, PLANES, H, S}, strides[3] = {PB, ROWB, SLOTB};
    const cuuint32_t ones[4] = {1, 1, 1, 1}, box0[4] = {32, 1, 2, 1}, box1[4] = {34, 4, 2, 1}, step1[4] = {1, 2, 1, 1};
    CUresult r0 = enc(&maps[0], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, dev, dims, strides, box0, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r1 = enc(&maps[1], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, dev, dims, strides, box1, step1, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint64_t dims2[4] = {W, 1, (H + 1) / 2, S}, strides2[3] = {PB, 2 * ROWB, SLOTB};
    const cuuint32_t box2[4] = {32, 1, 1, 1};
    CUresult r2 = enc(&maps[2], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, dev + 7 * PITCH, dims2, strides2, box2, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d %d\n", (int)r0, (int)r1, (int)r2);
    CUtensorMap* dmaps; cudaMalloc(&dmaps, sizeof(maps)); cudaMemcpy(dmaps, maps, sizeof(maps), cudaMemcpyHostToDevice);
    unsigned long long* out; cudaMalloc(&out, 8 * 1024); std::vector<unsigned long long> ho(1024);
    struct T { int m, c0, c1, c2, c3; unsigned bytes; int n; } tests[] = {
        {0, 31, 1, 4, 2, 512, 64}, {0, 93, 6, 38, 0, 512, 64}, {1, 30, 3, 4, 1, 1088, 136}, {1, -1, 2, 38, 2, 1088, 136}, {2, 62, 0, 3, 1, 256, 32}};
    const int mode = getenv("TMA_MODE") ? atoi(getenv("TMA_MODE")) : 0;
    for (auto& t : tests) {
        if (mode == 0) probe<0><<<1, 32>>>(maps[t.m], dmaps + t.m, t.c0, t.c1, t.c2, t.c3, t.bytes, out, t.n);
        if (mode == 1) probe<1><<<1, 32>>>(maps[t.m], dmaps + t.m, t.c0, t.c1, t.c2, t.c3, t.bytes, out, t.n);
        if (mode == 2) probe<2><<<1, 32>>>(maps[t.m], dmaps + t.m, t.c0, t.c1, t.c2, t.c3, t.bytes, out, t.n);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d map %d at (%d,%d,%d,%d): %s\n", mode, t.m, t.c0, t.c1, t.c2, t.c3, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(ho.data(), out, 8 * t.n, cudaMemcpyDeviceToHost);
        for (int i = 0; i < t.n; i += (t.m == 1 ? 34 : 32)) {
            const unsigned long long a = ho[i], b = ho[i + 1], z = ho[i + (t.m == 1 ? 33 : 31)];
            printf("  [%3d] slot %llu plane %llu row %llu col %llu | next col %llu | last: plane %llu row %llu col %llu raw %llx\n", i, a >> 48, (a >> 32) & 0xffff,
                   (a >> 16) & 0xffff, a & 0xffff, b & 0xffff, (z >> 32) & 0xffff, (z >> 16) & 0xffff, z & 0xffff, z);
        }
    }
    return 0;
}
