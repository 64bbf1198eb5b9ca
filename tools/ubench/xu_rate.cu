// Microbenchmark (measurement tool): issue cost in clocks per warp instruction of MUFU.EX2, F2FP (bf16x2 pack), FFMA, FADD, 3-input FMNMX
// and of the softmax inner chunk of tc_attention.cu, with 1, 2 or 4 warps resident per SM sub-partition.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
    unsigned acc = 0; float s0 = 0.f, s1 = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { for (int i = 0; i < 16; ++i) v[i] = ex2(v[i]); }
        if (MODE == 1) { for (int i = 0; i < 16; i += 2) acc += pack(v[i], v[i + 1]); for (int i = 0; i < 16; ++i) v[i] += 1.0f; }
        if (MODE == 2) { for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], 0.999f, -0.5f); }
        if (MODE == 3) {   // the attention chunk: 8 x (FFMA, EX2), 6 FADD, 4 pack
            for (int h = 0; h < 2; ++h) {
                float e[8];
                for (int j = 0; j < 8; ++j) e[j] = ex2(fmaf(v[8 * h + j], 0.125f, -1.0f));
                s0 += (e[0] + e[1]) + (e[2] + e[3]); s1 += (e[4] + e[5]) + (e[6] + e[7]);
                acc += pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
                for (int j = 0; j < 8; ++j) v[8 * h + j] += s0;
            }
        }
        if (MODE == 4) {   // same without the packs
            for (int h = 0; h < 2; ++h) {
                float e[8];
                for (int j = 0; j < 8; ++j) e[j] = ex2(fmaf(v[8 * h + j], 0.125f, -1.0f));
                s0 += (e[0] + e[1]) + (e[2] + e[3]); s1 += (e[4] + e[5]) + (e[6] + e[7]);
                for (int j = 0; j < 8; ++j) v[8 * h + j] += s0;
            }
        }
    }
    long long t1 = clock64();
    float r = s0 + s1; for (int i = 0; i < 16; ++i) r += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 8);
    const char* names[] = {"16 x MUFU.EX2", "8 x F2FP + 16 FADD", "16 x FFMA", "2 x attention chunk (16 FFMA, 16 EX2, 12 FADD, 8 F2FP, 16 FADD)", "same without F2FP"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int iters = 2000; long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k<0><<<148, warps * 32>>>(out, clk, iters); break;
                    case 1: k<1><<<148, warps * 32>>>(out, clk, iters); break;
                    case 2: k<2><<<148, warps * 32>>>(out, clk, iters); break;
                    case 3: k<3><<<148, warps * 32>>>(out, clk, iters); break;
                    default: k<4><<<148, warps * 32>>>(out, clk, iters); break;
                }
                cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            }
            printf("%-70s warps/SMSP %d: %.2f clk per iteration per SMSP\n", names[mode], warps / 4, (double)h / iters);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
