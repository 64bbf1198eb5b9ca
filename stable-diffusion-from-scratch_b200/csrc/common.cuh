// sdb200 — shared device/host helpers for every kernel translation unit.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sdb200.h"

namespace sdb {

// ---- error plumbing (C-ABI: never throw; return code + last-error string) -------------------
void set_last_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaGetLastError -> sdb code

#define SDB_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::sdb::set_last_error(__VA_ARGS__);                  \
            return SDB_ERR_INVALID;                              \
        }                                                        \
    } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the library is launched with programmatic stream serialization: it may become resident
// while its predecessor in the stream is still draining, runs its private prologue (barrier init, TMEM
// allocation, descriptor prefetch) and then blocks in pdl_wait() until the predecessor has completed and
// its writes are visible.  Rules every kernel follows: pdl_trigger() first (lets the successor in), no
// global-memory access before pdl_wait(), and pdl_wait() is executed by every thread that touches memory.
// SDB200_PDL=0 turns the launch attribute off (the device-side instructions are then no-ops).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through check_launch()
}

// same, for a kernel launched as thread-block clusters of `cluster_x` CTAs along x (runtime cluster size)
template <typename... KArgs, typename... Args>
static inline void launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// exact variants for the fp32 parity mode (expf, not the fast intrinsic)
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU(erf) for the bf16 tensor-core path: erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// bf16 rounding of the stored product), 2 MUFU + ~12 FMA-pipe instructions and no divergent branches; the
// fp32 parity mode keeps erff (gelu_erf above).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = poly * t * __expf(-z * z);          // 1 - erf(z), z >= 0
    const float erf_abs = 1.0f - e;
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// streaming 128-bit accesses (bandwidth kernels: data is touched once, keep it out of L1)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" :: "l"(p), "r"(a), "r"(b) : "memory");
}

}  // namespace sdb
