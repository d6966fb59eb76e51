// stream_read.cu — calibration only (not part of the product): read-only streaming rate of a 512 MiB region on
// B200 with (a) LDG.128 grid-stride, (b) 1-D bulk TMA 16 KiB tiles, (c) 2-D TMA 4 x [32 x 128 B] swizzled boxes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_read stream_read.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_ldg(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
    uint32_t acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t ph) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(ph) : "memory");
}
// per-warp pipelines, STAGES x 16 KiB; MODE 0: 1-D bulk, 1: 2-D 4 boxes
template <int WARPS, int STAGES, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) k_tma(const __grid_constant__ CUtensorMap tm, const uint8_t* base, uint32_t tiles, uint32_t* ctr, uint32_t* out) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bars[WARPS * STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sb = (s32(raw) + 1023u) & ~1023u;
    if (threadIdx.x < WARPS * STAGES) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[threadIdx.x])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t my = sb + warp * STAGES * 16384, mb = s32(&bars[warp * STAGES]);
    uint32_t tile_of[STAGES];
    auto issue = [&](int s, uint32_t t) {
        const uint32_t bar = mb + s * 8, dst = my + s * 16384;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384u) : "memory");
        if (MODE == 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(base + (size_t)t * 16384), "r"(16384u), "r"(bar) : "memory");
        } else {
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst + cb * 4096), "l"(&tm), "r"(bar), "r"(cb * 128), "r"((int)(t * 32)) : "memory");
        }
    };
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
        uint32_t t = 0;
        if (lane == 0) { t = atomicAdd(ctr, 1u); if (t < tiles) issue(s, t); }
        tile_of[s] = __shfl_sync(0xffffffffu, t, 0);
    }
    uint32_t ph = 0, acc = 0;
    for (uint32_t it = 0;; ++it) {
        const int s = it % STAGES;
        uint32_t tile = tile_of[0];
#pragma unroll
        for (int q = 1; q < STAGES; ++q) if (s == q) tile = tile_of[q];
        if (tile >= tiles) break;
        mwait(mb + s * 8, ph);
        uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(my + s * 16384 + lane * 512) : "memory");
        acc += v.x;
        __syncwarp();
        uint32_t t = 0;
        if (lane == 0) { asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ctr) : "memory"); if (t < tiles) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(s, t); } }
        t = __shfl_sync(0xffffffffu, t, 0);
#pragma unroll
        for (int q = 0; q < STAGES; ++q) if (s == q) tile_of[q] = t;
        if (s == STAGES - 1) ph ^= 1;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int S, int M>
int run(const char* name, CUtensorMap tm, uint8_t* buf, size_t bytes, uint32_t* ctr, uint32_t* out, int sms) {
    size_t smem = (size_t)W * S * 16384 + 1024;
    CK(cudaFuncSetAttribute(k_tma<W, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
        uint8_t* b = buf + (size_t)(r % 3) * bytes;
        CK(cudaMemset(ctr, 0, 4));
        cudaEventRecord(e0);
        k_tma<W, S, M><<<sms, W * 32, smem>>>(tm, b, (uint32_t)(bytes / 16384), ctr, out);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 2 && ms < best) best = ms;
    }
    printf("%-34s %8.4f ms  %7.1f GB/s\n", name, best, bytes / best / 1e6);
    return 0;
}
int main() {
    const size_t bytes = 512ull << 20;
    uint8_t* buf; CK(cudaMalloc(&buf, 3 * bytes)); CK(cudaMemset(buf, 1, 3 * bytes));
    uint32_t *ctr, *out; CK(cudaMalloc(&ctr, 8)); out = ctr + 1;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bpsm : {4, 8, 16}) {
        float best = 1e9;
        for (int r = 0; r < 6; ++r) {
            cudaEventRecord(e0);
            k_ldg<<<sms * bpsm, 256>>>((const uint4*)(buf + (size_t)(r % 3) * bytes), bytes / 16, out);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 2 && ms < best) best = ms;
        }
        printf("ldg.128 x8 unroll, %2d CTA/SM         %8.4f ms  %7.1f GB/s\n", bpsm, best, bytes / best / 1e6);
    }
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    enc_fn enc = (enc_fn)p;
    for (int promo = 0; promo < 4; ++promo) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {512, 3 * bytes / 512}; cuuint64_t str[1] = {512}; cuuint32_t box[2] = {128, 32}, es[2] = {1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        char nm[64];
        snprintf(nm, 64, "tma2d 14x1 l2promo=%d", promo); if (run<14, 1, 1>(nm, tm, buf, bytes, ctr, out, sms)) return 1;
        snprintf(nm, 64, "tma2d 7x2  l2promo=%d", promo); if (run<7, 2, 1>(nm, tm, buf, bytes, ctr, out, sms)) return 1;
        if (promo == 0) {
            if (run<14, 1, 0>("bulk1d 14x1", tm, buf, bytes, ctr, out, sms)) return 1;
            if (run<7, 2, 0>("bulk1d 7x2", tm, buf, bytes, ctr, out, sms)) return 1;
            if (run<4, 3, 0>("bulk1d 4x3", tm, buf, bytes, ctr, out, sms)) return 1;
            if (run<2, 7, 0>("bulk1d 2x7", tm, buf, bytes, ctr, out, sms)) return 1;
        }
    }
    return 0;
}
