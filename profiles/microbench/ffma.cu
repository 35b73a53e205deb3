// FP32 FMA issue-rate microbenchmark for sm_100a: scalar FFMA (3 register operands) vs packed fma.rn.f32x2.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma ffma.cu && ./ffma
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_scalar(float *out, float a, float b, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_packed(float *out, float a, float b, int iters) {
    unsigned long long acc[ILP];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    unsigned long long a2 = *reinterpret_cast<unsigned long long *>(&av), b2 = *reinterpret_cast<unsigned long long *>(&bv);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); acc[i] = *reinterpret_cast<unsigned long long *>(&v); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 v = *reinterpret_cast<float2 *>(&acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA with one operand from shared memory via broadcast LDS.128 (4 weights) vs registers: issue mix test
template <int NW>
__global__ void k_lds_mix(float *out, int iters) {
    __shared__ float4 w[64];
    if (threadIdx.x < 64) w[threadIdx.x] = make_float4(0.5f, 0.25f, 0.125f, 1.5f);
    __syncthreads();
    float acc[NW * 4];
#pragma unroll
    for (int i = 0; i < NW * 4; ++i) acc[i] = threadIdx.x + i;
    float x = threadIdx.x * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const float4 ww = w[(it + j) & 63];
            acc[j * 4 + 0] = fmaf(ww.x, x, acc[j * 4 + 0]);
            acc[j * 4 + 1] = fmaf(ww.y, x, acc[j * 4 + 1]);
            acc[j * 4 + 2] = fmaf(ww.z, x, acc[j * 4 + 2]);
            acc[j * 4 + 3] = fmaf(ww.w, x, acc[j * 4 + 3]);
        }
        x += 1e-6f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NW * 4; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 20000, blocks = 148 * 2, threads = 512;
    double n;
    float ms;
    ms = timeit([&] { k_scalar<16><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    n = (double)blocks * threads * iters * 16; printf("scalar FFMA ILP16 : %.1f TFLOP/s  (%.3f ms)\n", 2 * n / ms / 1e9, ms);
    ms = timeit([&] { k_packed<8><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    n = (double)blocks * threads * iters * 8 * 2; printf("packed f32x2 ILP8  : %.1f TFLOP/s  (%.3f ms)\n", 2 * n / ms / 1e9, ms);
    ms = timeit([&] { k_packed<16><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
    n = (double)blocks * threads * iters * 16 * 2; printf("packed f32x2 ILP16 : %.1f TFLOP/s  (%.3f ms)\n", 2 * n / ms / 1e9, ms);
    ms = timeit([&] { k_lds_mix<3><<<blocks, threads>>>(out, iters); });
    n = (double)blocks * threads * iters * 12; printf("LDS.128 + 4 FFMA x3: %.1f TFLOP/s  (%.3f ms)\n", 2 * n / ms / 1e9, ms);
    ms = timeit([&] { k_lds_mix<6><<<blocks, threads>>>(out, iters); });
    n = (double)blocks * threads * iters * 24; printf("LDS.128 + 4 FFMA x6: %.1f TFLOP/s  (%.3f ms)\n", 2 * n / ms / 1e9, ms);
    return 0;
}
