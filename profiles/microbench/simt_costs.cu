// Cost of the non-arithmetic instructions the tensor-core edge-conv epilogues lean on, per warp, with W warps per SM doing the same:
// redux.sync.max.f32 (independent / dependent), fence.proxy.async after shared stores, mbarrier.try_wait on a completed phase,
// tcgen05.ld 32x32b.x16 + wait::ld, mbarrier.arrive.  build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o simt_costs simt_costs.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024) costs(int mode, int reps, float *sink, long long *out) {
    __shared__ __align__(16) float buf[1024 * 4];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1 << 20) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");   // phase 0 of bar[0] completes
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    float v[12], acc = 0.f;
    for (int i = 0; i < 12; ++i) v[i] = (float)(tid * 13 % 17 + i);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (mode == 0) {            // 12 independent redux
#pragma unroll
            for (int i = 0; i < 12; ++i) { float m; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v[i])); if (lane == i) acc += m; }
        } else if (mode == 1) {     // 12 dependent redux
            float m = v[0];
#pragma unroll
            for (int i = 0; i < 12; ++i) { asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(m + v[i])); }
            acc += m;
        } else if (mode == 2) {     // 6 STS.128 + fence.proxy.async
#pragma unroll
            for (int i = 0; i < 6; ++i) *reinterpret_cast<float4 *>(&buf[tid * 4]) = make_float4(v[i], v[i + 1], acc, 1.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else if (mode == 3) {     // 6 STS.128 only
#pragma unroll
            for (int i = 0; i < 6; ++i) *reinterpret_cast<float4 *>(&buf[tid * 4]) = make_float4(v[i], v[i + 1], acc, 1.f);
            asm volatile("" ::: "memory");
        } else if (mode == 4) {     // try_wait on a completed phase
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
            acc += (float)done;
        } else if (mode == 5) {     // two tcgen05.ld x16 + wait
            uint32_t a[16], b[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]),
                           "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]) : "r"(tmem) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]), "=r"(b[10]),
                           "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]) : "r"(tmem + 16) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(a[3]) + __uint_as_float(b[5]);
        } else if (mode == 6) {     // mbarrier.arrive
            uint32_t done;             // test_wait instead of try_wait
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
            acc += (float)done;
        } else if (mode == 7) {     // tcgen05 fences
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        } else if (mode == 9) {     // 12-value max over the warp by a shuffle butterfly that halves the values per lane
            float w[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) w[i] = v[i] + acc;
            float a6[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { const bool up = lane & 16; const float send = up ? w[i] : w[6 + i], keep = up ? w[6 + i] : w[i]; a6[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16)); }
            float a3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { const bool up = lane & 8; const float send = up ? a6[i] : a6[3 + i], keep = up ? a6[3 + i] : a6[i]; a3[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8)); }
            float a2[2];
            { const bool up = lane & 4; const float send = up ? a3[0] : a3[1], keep = up ? a3[1] : a3[0]; a2[0] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4)); a2[1] = fmaxf(a3[2], __shfl_xor_sync(0xffffffffu, a3[2], 4)); }
            float a1;
            { const bool up = lane & 2; const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0]; a1 = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 2)); }
            a1 = fmaxf(a1, __shfl_xor_sync(0xffffffffu, a1, 1));
            acc += a1 * 1e-30f;
        } else if (mode == 8) {     // 12 cvt.rna.tf32 + 12 sub
#pragma unroll
            for (int i = 0; i < 12; ++i) { uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v[i] + acc)); acc += v[i] - __uint_as_float(h); }
        }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) sink[tid] = acc;
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(64u) : "memory");
}

int main() {
    long long *d, h; float *sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
    const char *names[] = {"12 independent redux.sync.max.f32 (+ select)", "12 dependent redux.sync.max.f32", "6 STS.128 + fence.proxy.async", "6 STS.128",
                           "mbarrier.try_wait (phase complete)", "2 x tcgen05.ld x16 + wait::ld", "mbarrier.test_wait (phase complete)", "tcgen05.fence before + after", "12 x (cvt.rna.tf32 + sub)", "12-value warp max by shuffle butterfly (13 SHFL)"};
    const int reps = 256;
    for (int threads = 128; threads <= 1024; threads *= 2) {
        printf("== %d warps per SM\n", threads / 32);
        for (int mode = 0; mode < 10; ++mode) {
            for (int it = 0; it < 2; ++it) costs<<<148, threads>>>(mode, reps, sink, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", names[mode], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%-48s %7.1f cycles per repetition (as seen by one warp)\n", names[mode], (double)h / reps);
        }
    }
    return 0;
}
