// sdb200 — fp32 SIMT contraction: the fp32 parity mode (eps rel-L2 <= 1e-5 needs true fp32
// products, which a single bf16/tf32 tensor-core pass cannot give) and the layers whose K is
// too small for a 64-wide tensor-core slab (conv_in C_in=4, DDPM initial conv C_in=3).
//
// out[m,n] = alpha * sum_k A(m,k) B(n,k) + bias[n] + rowvec[img(m),n] + residual[m,n]
//   conv mode : A(m,k) gathers NHWC x, k = (tap, c); B packed [tap][Cout][Cin]
//               (nn.Conv2d: openai_model/model.py:88-90,117,181,207,218,365,531)
//   gemm mode : A [M,K], B [N,K] or [K,N]  (nn.Linear / torch.bmm)
//
// 128x64x16 tiles, 256 threads, 8x4 register tile per thread, smem stored k-major so the inner
// product reads are conflict-free float4s.  Roofline: fp32 FFMA pipe (not tensor) — this path is
// for correctness, the throughput path is tc_contract.cu.
#include "common.cuh"

namespace sdb {

constexpr int SBM = 128, SBN = 64, SBK = 16;
constexpr int SIMT_FLUSH = 16;      // k-tiles per first-level partial sum

struct SimtP {
    const float* A; const float* B; float* out;
    const float* bias; const float* rowvec; const float* residual;
    long long lda, ldb, ldc, ldr, ldv;
    int M, N, K;
    float alpha;
    int b_kn;
    int nb2;
    long long sa1, sa2, sb1, sb2, sc1, sc2;
    int kh, kw, stride, pad, up;
    int NB, IH, IW, Cin, OH, OW;
    int out_bf16;
    int vecA, vecB;   // 128-bit loads legal
};

template <bool CONV>
__global__ void __launch_bounds__(256) simt_contract_kernel(const SimtP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int z = blockIdx.z;
    const int b1 = z / p.nb2, b2 = z % p.nb2;
    const float* A = p.A + b1 * p.sa1 + b2 * p.sa2;
    const float* B = p.B + b1 * p.sb1 + b2 * p.sb2;
    const long long coff = b1 * p.sc1 + b2 * p.sc2;

    // A loader: thread -> (row, 8 consecutive k)
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    const int gm = m0 + a_row;
    int img = 0, oh = 0, ow = 0;
    if (CONV && gm < p.M) {
        img = gm / (p.OH * p.OW);
        int rem = gm % (p.OH * p.OW);
        oh = rem / p.OW;
        ow = rem % p.OW;
    }
    // B loader
    const int b_row = p.b_kn ? (tid >> 4) : (tid >> 2);          // k (kn) or n (nk)
    const int b_col = p.b_kn ? (tid & 15) * 4 : (tid & 3) * 4;   // n (kn) or k (nk)

    const int ty = tid >> 4, tx = tid & 15;
    // Two-level accumulation: `acc` sums at most SIMT_FLUSH k-tiles (256 products) and is then folded into `tot`.  A single
    // running fp32 sum over K = 9 * 512 = 4608 products drifts by ~sqrt(K) ulp; the VAE decoder (39 such convs in a row) then
    // sat at 8.2e-6 of the 1e-5 fp32-mode budget at 512x512.  Blocked summation keeps the growth at ~sqrt(256) + sqrt(K / 256).
    float acc[8][4], tot[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }

    const int IHU = p.IH * p.up, IWU = p.IW * p.up;

    for (int k0 = 0; k0 < p.K; k0 += SBK) {
        // ---- load A tile ----
        float av[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) av[j] = 0.f;
        if (gm < p.M) {
            if (CONV) {
                if (p.vecA) {   // Cin % 16 == 0: the 16-wide k chunk sits inside one tap
                    int k = k0 + a_k;
                    int tap = k / p.Cin, c = k % p.Cin;
                    int r = tap / p.kw, s = tap % p.kw;
                    int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
                    if (ih >= 0 && ih < IHU && iw >= 0 && iw < IWU && k < p.K) {
                        const float* src = A + (((long long)img * p.IH + ih / p.up) * p.IW + iw / p.up) * p.lda + c;
                        float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
                        float4 v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
                        av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int k = k0 + a_k + j;
                        if (k < p.K) {
                            int tap = k / p.Cin, c = k % p.Cin;
                            int r = tap / p.kw, s = tap % p.kw;
                            int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
                            if (ih >= 0 && ih < IHU && iw >= 0 && iw < IWU)
                                av[j] = __ldg(A + (((long long)img * p.IH + ih / p.up) * p.IW + iw / p.up) * p.lda + c);
                        }
                    }
                }
            } else {
                const float* src = A + (long long)gm * p.lda + k0 + a_k;
                if (p.vecA && k0 + a_k + 8 <= p.K) {
                    float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
                    float4 v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                    av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
                    av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (k0 + a_k + j < p.K) av[j] = __ldg(src + j);
                }
            }
        }
        // ---- load B tile ----
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.b_kn) {
            int k = k0 + b_row, n = n0 + b_col;
            if (k < p.K) {
                const float* src = B + (long long)k * p.ldb + n;
                if (p.vecB && n + 4 <= p.N) {
                    float4 v = __ldg(reinterpret_cast<const float4*>(src));
                    bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < p.N) bv[j] = __ldg(src + j);
                }
            }
        } else {
            int n = n0 + b_row, k = k0 + b_col;
            if (n < p.N) {
                const float* src;
                if (CONV) {
                    int tap = k / p.Cin, c = k % p.Cin;      // valid for the vec path (chunk in one tap)
                    src = B + ((long long)tap * p.N + n) * p.Cin + c;
                    if (p.vecB && k + 4 <= p.K) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(src));
                        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            int kk = k + j;
                            if (kk < p.K) {
                                int t2 = kk / p.Cin, c2 = kk % p.Cin;
                                bv[j] = __ldg(B + ((long long)t2 * p.N + n) * p.Cin + c2);
                            }
                        }
                    }
                } else {
                    src = B + (long long)n * p.ldb + k;
                    if (p.vecB && k + 4 <= p.K) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(src));
                        bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (k + j < p.K) bv[j] = __ldg(src + j);
                    }
                }
            }
        }
        __syncthreads();   // previous tile fully consumed
#pragma unroll
        for (int j = 0; j < 8; ++j) As[a_k + j][a_row] = av[j];
        if (p.b_kn) {
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[b_row][b_col + j] = bv[j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[b_col + j][b_row] = bv[j];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if ((((k0 / SBK) + 1) % SIMT_FLUSH) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += tot[i][j];

    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
        int im = CONV ? m / (p.OH * p.OW) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = p.alpha * acc[i][j];
            if (p.bias) v += p.bias[n];
            if (CONV && p.rowvec) v += p.rowvec[(long long)im * p.ldv + n];
            if (p.residual) v += p.residual[coff + (long long)m * p.ldr + n];
            long long o = coff + (long long)m * p.ldc + n;
            if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v);
            else p.out[o] = v;
        }
    }
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_simt_contract(const sdb_simt_args* a, void* stream) {
    SDB_REQUIRE(a && a->A && a->B && a->out, "simt_contract: null pointer");
    SDB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "simt_contract: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
    SimtP p;
    p.A = a->A; p.B = a->B; p.out = a->out;
    p.bias = a->bias; p.rowvec = a->rowvec; p.residual = a->residual;
    p.lda = a->lda; p.ldb = a->ldb; p.ldc = a->ldc; p.ldr = a->ldr; p.ldv = a->ldv;
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.alpha = a->alpha;
    p.b_kn = a->b_kn;
    int nb1 = a->nb1 > 0 ? a->nb1 : 1;
    p.nb2 = a->nb2 > 0 ? a->nb2 : 1;
    p.sa1 = a->sa1; p.sa2 = a->sa2; p.sb1 = a->sb1; p.sb2 = a->sb2; p.sc1 = a->sc1; p.sc2 = a->sc2;
    p.kh = a->kh; p.kw = a->kw; p.stride = a->stride; p.pad = a->pad; p.up = a->up > 0 ? a->up : 1;
    p.NB = a->NB; p.IH = a->IH; p.IW = a->IW; p.Cin = a->Cin; p.OH = a->OH; p.OW = a->OW;
    p.out_bf16 = a->out_dtype == SDB_BF16;
    const bool conv = a->kh > 0;
    if (conv) {
        SDB_REQUIRE(a->kw > 0 && a->stride > 0 && a->Cin > 0 && a->NB > 0, "simt_contract: bad conv geometry");
        SDB_REQUIRE(a->K == a->kh * a->kw * a->Cin, "simt_contract: K != kh*kw*Cin");
        SDB_REQUIRE(a->M == a->NB * a->OH * a->OW, "simt_contract: M != NB*OH*OW");
        SDB_REQUIRE(!a->b_kn && nb1 * p.nb2 == 1, "simt_contract: conv mode is unbatched, B [N,K]");
        p.vecA = (a->Cin % 16 == 0) && (a->lda % 4 == 0) && (((uintptr_t)a->A & 15) == 0);
        p.vecB = (a->Cin % 16 == 0) && (((uintptr_t)a->B & 15) == 0);
    } else {
        p.vecA = (a->lda % 4 == 0) && (((uintptr_t)a->A & 15) == 0) && (a->sa1 % 4 == 0) && (a->sa2 % 4 == 0);
        p.vecB = (a->ldb % 4 == 0) && (((uintptr_t)a->B & 15) == 0) && (a->sb1 % 4 == 0) && (a->sb2 % 4 == 0);
    }
    dim3 grid(ceil_div(a->M, SBM), ceil_div(a->N, SBN), nb1 * p.nb2);
    SDB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "simt_contract: grid too large");
    if (conv) launch_pdl(simt_contract_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
    else launch_pdl(simt_contract_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
    return check_launch("simt_contract_kernel");
}
