// Hardware question (measurement tool): can a K-major SWIZZLE_128B UMMA descriptor start at an arbitrary 128-byte row of a larger
// shared-memory tile (start address NOT 1024-byte aligned) and step between its 8-row groups with SBO != 1024?  That is what a
// shared-memory-resident halo for 3x3 convolutions needs: the (tw + 2) x (th + 2) input patch of a 64-channel slab is loaded ONCE
// (row = pixel, 128 B = 64 bf16 channels, 16-byte chunk c of row r stored at chunk c ^ (r & 7), the pattern TMA writes) and the nine
// taps are issued from the same bytes with start = base + (r * (tw + 2) + s) * 128 and SBO = (tw + 2) * 128.
// One CTA, M = 128 (16 groups of 8 rows), N = 64, K = 64, B = identity: D[m][n] must equal A[row(m)][n].
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../stable-diffusion-from-scratch_b200/csrc/ptx.cuh"
using namespace sdb::ptx;

constexpr int TW = 8, TH = 16, HW_ = TW + 2, HH = TH + 2, HROWS = HW_ * HH;      // 10 x 18 = 180 halo rows

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) k(float* out, int tap_r, int tap_s, int use_base_off) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                            // 180 rows x 128 B (rounded up to 24 KB)
    uint8_t* sB = smem + 24 * 1024;                // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // A[row][k] = (row + 3 * k) % 251 (exact in bf16), swizzled like TMA writes it (pattern keyed on the ABSOLUTE row index)
    for (int i = tid; i < HROWS * 8; i += 128) {
        const int row = i >> 3, c = i & 7;
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((float)((row + 3 * (c * 8 + j)) % 251));
        *reinterpret_cast<uint4*>(sA + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int row = i >> 3, c = i & 7;
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((c * 8 + j) == row ? 1.0f : 0.0f);
        *reinterpret_cast<uint4*>(sB + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t off_rows = tap_r * HW_ + tap_s;
        const uint32_t a_addr = smem_u32(sA) + off_rows * 128;
        const uint64_t adesc = desc_sw128(a_addr, HW_ * 128, use_base_off ? (off_rows & 7) : 0);
        const uint64_t bdesc = desc_sw128(smem_u32(sB), 1024, 0);
        const uint32_t idesc = umma_idesc_bf16(64, false, false);
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
    float* d; cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    for (int bo = 0; bo < 2; ++bo)
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) {
                k<<<1, 128, 40 * 1024>>>(d, r, s, bo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("tap (%d,%d) base_off=%d: CUDA error %s\n", r, s, bo, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                int bad = 0, first = -1;
                for (int m = 0; m < 128; ++m) {
                    const int row = (m / 8) * HW_ + (m % 8) + r * HW_ + s;
                    for (int n = 0; n < 64; ++n)
                        if (h[m * 64 + n] != (float)((row + 3 * n) % 251)) { if (first < 0) first = m * 64 + n; ++bad; }
                }
                printf("tap (%d,%d) start row %2d base_offset field %s: %s (%d wrong of 8192%s)\n", r, s, r * HW_ + s, bo ? "set" : "0  ",
                       bad ? "MISMATCH" : "exact", bad, bad ? "" : "");
                if (bad && first >= 0) printf("    first wrong: m=%d n=%d got %.0f\n", first / 64, first % 64, h[first]);
            }
    return 0;
}
