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

// GELU(erf) for the bf16 tensor-core path.  erfc(z) = 2^(-z P(z)) with a quartic P fitted on [0, 4.2] (|erf error| <= 6e-7;
// beyond 4.2 both sides are 0 in fp32), z = |x| / sqrt(2) folded into the coefficients, so
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2) = max(x, 0) - 0.5 |x 2^(|x| Q(|x|))|.
// ONE MUFU (ex2) and 8 FMA/ALU-pipe instructions, no branches; max |error| 1.2e-6 over [-12, 12] (checked against
// float64 erf), far below the bf16 rounding of the stored product.  The previous form (Abramowitz-Stegun 7.1.26:
// rcp + exp + ~17 FMA-pipe instructions) made the GEGLU epilogue, not the MMA, pace the K <= 640 feed-forward layers
// (profiles/r01_pair_timeline.txt).  The fp32 parity mode keeps erff (gelu_erf above).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float a = fabsf(x);
    float q = fmaf(-0.0005204587f, a, 0.007397511f);
    q = fmaf(q, a, -0.05256124f);
    q = fmaf(q, a, -0.45925468f);
    q = fmaf(q, a, -1.1510913f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * q));
    return fmaf(-0.5f, fabsf(x * e), fmaxf(x, 0.0f));
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
