// atomics.cu — calibration only: cost of 1 M random index operations on a large table (B200).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned long long mix(unsigned long long k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
typedef unsigned __int128 u128;
// MODE 0: cas128+red.max  1: cas128 only  2: cas64 only  3: cas64 + red.max(u32 next word)  4: plain load 16B  5: plain store 16B
// 6: red.max.u64 only  7: prefetch then cas128  8: atom.exch.b64
template <int MODE, int SLOT>
__global__ void k(uint8_t* table, unsigned long long mask, uint32_t n, unsigned long long salt, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = mix(i + salt), idx = h & mask;
    uint8_t* s = table + idx * SLOT;
    unsigned long long klo = h | 1, khi = mix(h);
    uint32_t acc = 0;
    if (MODE == 0 || MODE == 1 || MODE == 7) {
        if (MODE == 7) asm volatile("prefetch.global.L2 [%0];" ::"l"(s));
        u128 key = ((u128)khi << 64) | klo, old;
        asm volatile("atom.relaxed.gpu.global.cas.b128 %0, [%1], %2, %3;" : "=q"(old) : "l"(s), "q"((u128)0), "q"(key) : "memory");
        acc = (uint32_t)old;
        if (MODE == 0) asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(s + 16), "r"(~i) : "memory");
    } else if (MODE == 2 || MODE == 3) {
        unsigned long long old = atomicCAS((unsigned long long*)s, 0ULL, klo);
        acc = (uint32_t)old;
        if (MODE == 3) asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(s + 8), "r"(~i) : "memory");
    } else if (MODE == 4) {
        uint4 v = __ldcg((const uint4*)s); acc = v.x ^ v.w;
    } else if (MODE == 5) {
        __stcg((uint4*)s, make_uint4((uint32_t)klo, (uint32_t)(klo >> 32), (uint32_t)khi, i));
    } else if (MODE == 6) {
        asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(s), "l"(klo) : "memory");
    } else if (MODE == 8) {
        unsigned long long old = atomicExch((unsigned long long*)s, klo); acc = (uint32_t)old;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int MODE, int SLOT>
int run(const char* name, uint8_t* table, size_t slots, uint32_t n, uint32_t* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE, SLOT><<<(n + 255) / 256, 256>>>(table, slots - 1, n, 1234567ULL * (r + 1) + MODE * 977, out);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 1 && ms < best) best = ms;
    }
    printf("%-44s %8.4f ms  %7.2f G ops/s\n", name, best, n / best / 1e6);
    return 0;
}
int main(int argc, char** argv) {
    const uint32_t n = 1 << 20;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    if (argc > 1) {   // size sweep: random 16 B loads and cas.b128 over tables of growing size
        for (int lg = 21; lg <= 29; ++lg) {
            size_t slots = (size_t)1 << lg;
            uint8_t* t; CK(cudaMalloc(&t, slots * 32)); CK(cudaMemset(t, 0, slots * 32));
            char nm[64]; snprintf(nm, 64, "table %6.0f MB: ld.cg 16 B", slots * 32 / 1048576.0);
            run<4, 32>(nm, t, slots, n, out);
            snprintf(nm, 64, "table %6.0f MB: cas.b128", slots * 32 / 1048576.0);
            run<1, 32>(nm, t, slots, n, out);
            cudaFree(t);
        }
        return 0;
    }
    for (size_t slots : {(size_t)1 << 21, (size_t)1 << 26}) {
        uint8_t* t; CK(cudaMalloc(&t, slots * 32)); CK(cudaMemset(t, 0, slots * 32));
        printf("-- table %zu slots (%.0f MB at 32 B)\n", slots, slots * 32 / 1e6);
        run<0, 32>("cas.b128 + red.max.u32 (32 B slot)", t, slots, n, out);
        run<1, 32>("cas.b128 only", t, slots, n, out);
        run<7, 32>("prefetch.L2 + cas.b128", t, slots, n, out);
        run<2, 16>("cas.b64 only (16 B slot)", t, slots, n, out);
        run<3, 16>("cas.b64 + red.max.u32 (16 B slot)", t, slots, n, out);
        run<8, 16>("atom.exch.b64 (16 B slot)", t, slots, n, out);
        run<6, 16>("red.max.u64 only", t, slots, n, out);
        run<4, 32>("ld.cg 16 B", t, slots, n, out);
        run<5, 32>("st.cg 16 B", t, slots, n, out);
        cudaFree(t);
    }
    return 0;
}
