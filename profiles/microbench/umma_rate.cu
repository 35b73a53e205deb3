// How long does one small tcgen05.mma take?  kind::tf32, K = 8, cta_group::1; one issuing thread per CTA streams R dependent-free
// MMAs of one shape (alternating between two accumulators) and commits; cycles = clock64 from first issue to the commit's arrival.
// Variants: A from shared memory (K-major, no swizzle) or from tensor memory; 1 or 2 CTAs per SM, every SM busy.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

template <bool TS, bool F16>
__global__ void __launch_bounds__(128) rate(uint32_t idesc, int reps, int ncols, long long *out) {
    extern __shared__ unsigned char raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 24 * 1024 / 4; i += 128) reinterpret_cast<float *>(raw)[i] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (tid == 0) {
        const uint64_t ad = make_desc(smem_u32(raw), 128, 256), bd = make_desc(smem_u32(raw + 8192), 128, 256);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (r & 1) * ncols;     // two accumulators: no accumulate dependency between neighbours
            if (TS) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(tmem + 240u), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            } else if (F16) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    if (tid == 0) { t2 = clock64(); if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; } }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int reps = 512;
    struct Case { const char *name; int ts, f16, m, n; };
    const Case cases[] = {
        {"SS tf32 M=128 N=16", 0, 0, 128, 16}, {"SS tf32 M=128 N=32", 0, 0, 128, 32}, {"SS tf32 M=128 N=64", 0, 0, 128, 64},
        {"SS tf32 M=128 N=96", 0, 0, 128, 96}, {"SS tf32 M=128 N=128", 0, 0, 128, 128}, {"SS tf32 M=128 N=256", 0, 0, 128, 256},
        {"SS tf32 M=64  N=32", 0, 0, 64, 32}, {"SS tf32 M=64  N=64", 0, 0, 64, 64}, {"SS tf32 M=64  N=128", 0, 0, 64, 128}, {"SS tf32 M=64  N=256", 0, 0, 64, 256},
        {"TS tf32 M=128 N=32", 1, 0, 128, 32}, {"TS tf32 M=128 N=64", 1, 0, 128, 64}, {"TS tf32 M=128 N=128", 1, 0, 128, 128},
        {"SS f16 (K=16) M=128 N=32", 0, 1, 128, 32}, {"SS f16 (K=16) M=128 N=64", 0, 1, 128, 64}, {"SS f16 (K=16) M=64 N=256", 0, 1, 64, 256},
    };
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
        printf("== %d CTA(s) per SM, all 148 SMs busy; %d MMAs per CTA\n", ctas_per_sm, reps);
        for (const Case &c : cases) {
            const uint32_t fmt = c.f16 ? 0u : 2u;      // a/b format: f16 = 0, tf32 = 2
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
            const int ncols = c.n > 120 ? 0 : c.n;      // N = 128 / 256: one accumulator (dependent accumulates)
            const size_t smem = ctas_per_sm == 1 ? 120 * 1024 : 100 * 1024;
            for (int it = 0; it < 2; ++it) {
                if (c.ts) { cudaFuncSetAttribute(rate<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate<true, false><<<148 * ctas_per_sm, 128, smem>>>(idesc, reps, ncols, d); }
                else if (c.f16) { cudaFuncSetAttribute(rate<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate<false, true><<<148 * ctas_per_sm, 128, smem>>>(idesc, reps, ncols, d); }
                else { cudaFuncSetAttribute(rate<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); rate<false, false><<<148 * ctas_per_sm, 128, smem>>>(idesc, reps, ncols, d); }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-28s issue %6.1f cycles/MMA, complete %6.1f cycles/MMA (per CTA)\n", c.name, (double)h[0] / reps, (double)h[1] / reps);
        }
    }
    return 0;
}
