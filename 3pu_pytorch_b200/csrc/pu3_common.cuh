// Shared device/host helpers of libpu3_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pu3_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpu3_b200 is written for sm_100a (B200) only"
#endif

namespace pu3 {

// ---- host: status + error text -------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_status(cudaError_t e, const char *what);  // 0 or positive cudaError_t with text recorded

#define PU3_ARG_CHECK(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::pu3::set_error(__VA_ARGS__);  \
            return PU3_E_ARG;               \
        }                                   \
    } while (0)

#define PU3_LAUNCH_CHECK(what)                                   \
    do {                                                         \
        int _st = ::pu3::cuda_status(cudaGetLastError(), what);  \
        if (_st) return _st;                                     \
    } while (0)

struct DeviceInfo {
    int sm_count;
    int smem_optin;
    int cc;
};
const DeviceInfo &device_info();  // cached per device

static inline cudaStream_t as_stream(pu3_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// per-family timing inside composite entry points (no-ops unless pu3_prof_enable(1))
enum ProfTag { PROF_CONV = 0, PROF_KNN_FEAT = 1, PROF_EDGECONV = 2, PROF_KNN_SKIP = 3, PROF_SKIP_FUSE = 4, PROF_EXPAND = 5, PROF_MISC = 6, PROF_CONV_TC = 7, PROF_CONV_TC_PREP = 8, PROF_NTAGS = 9 };
int prof_begin(int tag, cudaStream_t s);
void prof_end(int id, cudaStream_t s);

// ---- device ---------------------------------------------------------------------------------
// The reference's squared distance as its SASS evaluates it (nmdistance_cuda.cu:27-30,
// sampling_cuda.cu:143): FMUL(dy,dy) -> FFMA(dx,dx,.) -> FFMA(dz,dz,.).  Spelled with
// intrinsics so no compiler version can reassociate it.
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Monotone map float -> uint32 (total order, -0 < +0, NaNs at the ends).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    // negative: ~u, else u | 0x80000000 -- both are u ^ (sign-extension | 0x80000000): a shift and one LOP3, no select
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// cluster helpers (inline PTX; cooperative_groups would do, this keeps the SASS visible)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local_smem_ptr` in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_cta(const void *local_smem_ptr, uint32_t rank) {
    uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem_ptr)), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// ---- packed fp32 (sm_100a FFMA2): two IEEE round-to-nearest FMAs per issued instruction ---------------
// B200 issues scalar FFMA at full rate, but kernels that also have to issue the shared-memory loads of their
// operands are issue-bound (profiles/r1_microbench_ffma.md); FFMA2 halves the FMA issue cost.  Results are
// bit-identical to two scalar fmaf().
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ---- mbarrier + async DSMEM stores (leader-to-leader exchange without a cluster-wide barrier) ----------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init_cluster() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {   // try_wait suspends the thread in hardware for a bounded time, so this is not a hot spin
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
// TMA bulk copy (the copy engine, SASS UBLKCP): `bytes` (multiple of 16) from 16-byte aligned global memory into 16-byte aligned
// shared memory of this CTA, completing on `bar` (complete_tx::bytes)
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// 16-byte store into a peer CTA's shared memory that signals `bytes written` on that peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t remote_bar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar) : "memory");
}

}  // namespace pu3
