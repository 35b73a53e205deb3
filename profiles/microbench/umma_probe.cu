// Stand-alone bring-up probe for tcgen05.mma kind::tf32 (one CTA, one MMA of M=128, N=64, K=8):
// checks shared-memory descriptor conventions against a CPU product.
//   mode 0: A K-major no swizzle, B K-major no swizzle
//   mode 1: A K-major SW128,      B K-major SW128
//   mode 2: A MN-major SW128,     B K-major SW128   (yields zeros: not a valid tf32 layout)
//   mode 3: A MN-major SW128 with 32-byte atoms (layout type 1), B K-major SW128   (the layout conv_tc.cu uses)
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, int mode, uint32_t idesc,
                                             uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int *info) {
    extern __shared__ unsigned char raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t s0 = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char *sm = raw + (s0 - smem_u32(raw));
    float *sa = reinterpret_cast<float *>(sm);            // 16 KB region for A
    float *sb = reinterpret_cast<float *>(sm + 16384);    // 16 KB region for B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) reinterpret_cast<float *>(sm)[i] = 0.f;
    __syncthreads();
    // A is (128 x 8) row-major in global, B is (64 x 8) row-major in global
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        int off;   // in bytes
        if (mode == 0) off = (r / 8) * 256 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4;
        else if (mode == 1) off = r * 128 + (((k / 4) ^ (r & 7)) << 4) + (k % 4) * 4;
        else if (mode == 2) off = (r / 32) * 4096 + k * 128 + ((((r % 32) / 4) ^ (k & 7)) << 4) + (r % 4) * 4;
        else off = (r / 32) * 4096 + k * 128 + ((((r % 32) / 8) ^ (k & 3)) << 5) + (r % 8) * 4;   // 128B swizzle, 32B atoms
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sa) + off) = A[i];
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i / 8, k = i % 8;
        int off;
        if (mode == 0) off = (r / 8) * 256 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4;
        else off = r * 128 + (((k / 4) ^ (r & 7)) << 4) + (k % 4) * 4;
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sb) + off) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t layout = mode == 0 ? 0u : 2u;
        const uint64_t ad = make_desc(smem_u32(sa), a_lbo, a_sbo, mode == 3 ? 1u : layout);
        const uint64_t bd = make_desc(smem_u32(sb), b_lbo, b_sbo, layout);
        info[0] = (int)tmem; info[1] = (int)smem_u32(sa); info[2] = (int)(ad >> 32); info[3] = (int)ad;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait phase 0
    {
        uint32_t done = 0; int spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (++spins > 2000000) { if (tid == 0) info[4] = -1; break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int ch = 0; ch < 2; ++ch) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ch * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * 64 + ch * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
    float hA[128 * 8], hB[64 * 8], hD[128 * 64];
    for (int i = 0; i < 128 * 8; ++i) hA[i] = (float)((i * 7) % 13 - 6);
    for (int i = 0; i < 64 * 8; ++i) hB[i] = (float)((i * 5) % 11 - 5);
    float *dA, *dB, *dD; int *dinfo;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD); cudaMalloc(&dinfo, 64);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
    struct Case { const char *name; int mode; int a_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; };
    const Case cases[] = {
        {"K/K no-swizzle  LBO=128 SBO=256", 0, 0, 128, 256, 128, 256},
        {"K/K no-swizzle  LBO=256 SBO=128", 0, 0, 256, 128, 256, 128},
        {"K/K SW128       LBO=16  SBO=1024", 1, 0, 16, 1024, 16, 1024},
        {"MN/K SW128      A LBO=4096 SBO=1024", 2, 1, 4096, 1024, 16, 1024},
        {"MN/K SW128      A LBO=1024 SBO=4096", 2, 1, 1024, 4096, 16, 1024},
        {"MN(32B atoms)/K A LBO=4096 SBO=512", 3, 1, 4096, 512, 16, 1024},
        {"MN(32B atoms)/K A LBO=512 SBO=4096", 3, 1, 512, 4096, 16, 1024},
    };
    for (const Case &c : cases) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        cudaMemset(dD, 0xff, sizeof hD); cudaMemset(dinfo, 0, 64);
        probe<<<1, 128, 34 * 1024>>>(dA, dB, dD, c.mode, idesc, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, dinfo);
        cudaError_t e = cudaDeviceSynchronize();
        int info[8]; cudaMemcpy(info, dinfo, 32, cudaMemcpyDeviceToHost);
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0; double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < 8; ++k) ref += (double)hA[m * 8 + k] * hB[n * 8 + k];
                const double err = fabs(ref - hD[m * 64 + n]);
                if (err > 1e-3) ++bad;
                if (hD[m * 64 + n] == 0.f) ++zeros;
                if (err > maxerr) maxerr = err;
            }
        printf("%-40s idesc=%08x: %s  bad=%d/8192 zeros=%d maxerr=%g  cuda=%s tmem=%x smemA=%x desc=%08x%08x timeout=%d  D[0][0..3]=%g %g %g %g\n",
               c.name, idesc, bad ? "MISMATCH" : "OK", bad, zeros, maxerr, cudaGetErrorString(e), info[0], info[1], info[2], info[3], info[4],
               hD[0], hD[1], hD[2], hD[3]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
