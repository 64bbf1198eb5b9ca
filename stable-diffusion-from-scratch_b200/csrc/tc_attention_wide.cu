// sdb200 — fused attention forward for WIDE single heads (d = 256 or 512) on tcgen05: the VAE AttnBlock.
//
// Replaces AttnBlock.forward's  w_ = bmm(q, k) * c**-0.5; softmax; h_ = bmm(v, w_)  (ldm/modules/diffusionmodules/
// model.py:180-204: one head, d = C = 512 channels, 4096 tokens at the 64x64 mid block), which the reference — and this
// repo until round 2 — executes as three launches per image over a materialised 4096 x 4096 score matrix.
//
// A 128-row query tile with d = 512 needs 512 TMEM columns for O alone, so the head is processed as d/256 independent
// OUTPUT halves: CTA (q tile, half, image) recomputes S = Q K^T over the full d and accumulates
// O[:, half] += P V[:, half] (256 TMEM columns) — 1.5x the minimal MMA work, no score matrix in memory, one launch.
//   warp 0      TMA producer : Q tile once (d/64 blocks of [128 x 64], resident), then per 64-key tile the d/64 K chunks
//                              [64 keys x 64 ch] through a 4-deep ring and the V half tile [64 keys x 256 ch]
//   warp 1      MMA issuer   : S (64 TMEM columns) = sum over chunks Q_c K_c^T, 4 x K16 each; O (256 columns) += P V
//   warp 2      TMEM allocator
//   warps 4..7  softmax      : thread == query row; S row -> registers (S is then free for the next Q K^T), row max,
//                              p = exp2(s*c - m), row sum, P -> smem (bf16, 128B-swizzled K-major), lazy (2^8) rescale of O
// Smem: Q 128 KB + K ring 32 KB + V 32 KB + P 16 KB = 208 KB.  Tensor-bound: per key tile 1024 clk of Q K^T + 512 clk of P V
// against ~550 clk of exponentials (128 x 64 / 16 per clk), so the MUFU work hides under the MMAs.
// Algorithmic FLOP = 4 * B * Sq * Sk * d (executed: 6 * B * Sq * Sk * d for d = 512).
#include "common.cuh"
#include "ptx.cuh"
#include <string.h>

namespace sdb {

using namespace ptx;

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr);

struct AttnWP {
    void* out;
    long long o_bs, o_ss;
    int Sq, Sk, d;
    float scale_log2;
};

constexpr int AW_KT = 64;                 // keys per tile
constexpr int AW_DV = 256;                // output channels per CTA
constexpr int AW_KRING = 4;               // K chunk ring depth
constexpr int AW_QBLK = 128 * 64 * 2;     // one [128 x 64] bf16 block
constexpr int AW_KBLK = AW_KT * 64 * 2;   // one [64 keys x 64 ch] block
constexpr int AW_SMEM = 1024 + 8 * AW_QBLK + AW_KRING * AW_KBLK + 4 * AW_KBLK + AW_QBLK + 256;
constexpr int AW_S_COL = 0, AW_O_COL = 64;

__device__ __forceinline__ float aw_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(256, 1)
tc_attention_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const AttnWP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                               // [8] blocks of [128 rows x 64 ch]
    uint8_t* sK = sQ + 8 * AW_QBLK;                   // [AW_KRING] chunks [64 keys x 64 ch]
    uint8_t* sV = sK + AW_KRING * AW_KBLK;            // [4] blocks [64 keys x 64 ch] = the 256-channel half
    uint8_t* sP = sV + 4 * AW_KBLK;                   // [128 rows x 64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AW_QBLK);
    uint64_t* q_full = bars;                // 1
    uint64_t* k_full = bars + 1;            // AW_KRING
    uint64_t* k_empty = k_full + AW_KRING;  // AW_KRING
    uint64_t* v_full = k_empty + AW_KRING;  // 1
    uint64_t* v_empty = v_full + 1;         // 1
    uint64_t* s_full = v_empty + 1;         // 1  S written by the MMA
    uint64_t* s_empty = s_full + 1;         // 1  S copied to registers by the 128 softmax threads
    uint64_t* p_full = s_empty + 1;         // 1  P written to smem
    uint64_t* pv_done = p_full + 1;         // 1  O += P V retired (P and O are free)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, half = blockIdx.y, b = blockIdx.z;
    const int nq = p.d / 64;                          // 64-channel chunks of the contraction
    const int ntiles = (p.Sk + AW_KT - 1) / AW_KT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < AW_KRING; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        mbar_init(v_full, 1); mbar_init(v_empty, 1);
        mbar_init(s_full, 1); mbar_init(s_empty, 128);
        mbar_init(p_full, 128); mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, (uint32_t)nq * AW_QBLK);
            for (int j = 0; j < nq; ++j) tma_load_4d(sQ + j * AW_QBLK, &tmQ, q_full, j * 64, 0, q0, b);
            int s = 0; uint32_t ph = 0;
            for (int t = 0; t < ntiles; ++t) {
                for (int c = 0; c < nq; ++c) {
                    mbar_wait(&k_empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&k_full[s], AW_KBLK);
                    tma_load_4d(sK + s * AW_KBLK, &tmK, &k_full[s], c * 64, 0, t * AW_KT, b);
                    if (++s == AW_KRING) { s = 0; ph ^= 1; }
                }
                mbar_wait(v_empty, (t & 1) ^ 1);
                mbar_arrive_expect_tx(v_full, 4 * AW_KBLK);
                for (int j = 0; j < 4; ++j) tma_load_4d(sV + j * AW_KBLK, &tmV, v_full, half * AW_DV + j * 64, 0, t * AW_KT, b);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
        const uint32_t idesc_s = umma_idesc_bf16(AW_KT, false, false);        // S = Q K^T : N = 64 keys
        const uint32_t idesc_o = umma_idesc_bf16(AW_DV, false, true);         // O = P V   : N = 256, V is MN-major
        const uint64_t qdesc0 = umma_desc_kmajor_sw128(smem_u32(sQ));
        const uint64_t kdesc0 = umma_desc_kmajor_sw128(smem_u32(sK));
        const uint64_t vdesc0 = umma_desc_mnmajor_sw128(smem_u32(sV), AW_KBLK);
        const uint64_t pdesc0 = umma_desc_kmajor_sw128(smem_u32(sP));
        mbar_wait(q_full, 0);
        int s = 0; uint32_t ph = 0;
        for (int t = 0; t <= ntiles; ++t) {
            if (t < ntiles) {
                mbar_wait(s_empty, (t & 1) ^ 1);                 // the softmax holds S(t-1) in registers
                tcgen05_fence_after();
                for (int c = 0; c < nq; ++c) {
                    mbar_wait(&k_full[s], ph);
                    tcgen05_fence_after();
                    const uint64_t qd = qdesc0 + (uint64_t)(c * (AW_QBLK >> 4));
                    const uint64_t kd = kdesc0 + (uint64_t)(s * (AW_KBLK >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (elect_one()) umma_bf16_ss(tmem + AW_S_COL, qd + 2 * k, kd + 2 * k, idesc_s, (c > 0 || k > 0) ? 1u : 0u);
                    if (elect_one()) umma_commit(&k_empty[s]);
                    if (++s == AW_KRING) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(s_full);
            }
            if (t >= 1) {
                const int tp = t - 1;
                mbar_wait(v_full, tp & 1);
                mbar_wait(p_full, tp & 1);
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < AW_KT / 16; ++k)             // 64 keys = 4 x K16
                    if (elect_one()) umma_bf16_ss(tmem + AW_O_COL, pdesc0 + (uint64_t)(2 * k), vdesc0 + (uint64_t)(k * 128), idesc_o,
                                                  (tp > 0 || k > 0) ? 1u : 0u);
                if (elect_one()) umma_commit(pv_done);
                if (elect_one()) umma_commit(v_empty);
            }
        }
    } else if (warp >= 4) {
        // ================= softmax / correction / epilogue =================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
        const uint32_t s_addr = lane_addr + AW_S_COL, o_addr = lane_addr + AW_O_COL;
        const uint32_t prow = smem_u32(sP) + row * 128;
        float m_used = -INFINITY, l = 0.f;
        const float c = p.scale_log2;
        for (int t = 0; t < ntiles; ++t) {
            const int kvalid = p.Sk - t * AW_KT;
            mbar_wait(s_full, t & 1);
            tcgen05_fence_after();
            uint32_t r[AW_KT];
            tmem_ld_x32(s_addr, r);
            tmem_ld_x32(s_addr + 32, r + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(s_empty);
            if (kvalid < AW_KT) {
#pragma unroll
                for (int j = 0; j < AW_KT; ++j) if (j >= kvalid) r[j] = 0xff800000u;   // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < AW_KT; j += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(r[j]));
                mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
            }
            const float mx = fmaxf(mx0, mx1) * c;
            const bool need = mx > m_used + 8.0f;
            float alpha = 1.0f;
            if (need) { alpha = aw_ex2(m_used - mx); m_used = mx; }
            const bool any_need = __any_sync(0xffffffffu, need);
            // the single P buffer and O are free once P V of the previous tile has retired
            if (t > 0) {
                mbar_wait(pv_done, (t - 1) & 1);
                tcgen05_fence_after();
                if (any_need) {
#pragma unroll 1
                    for (int c0 = 0; c0 < AW_DV; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld_x16(o_addr + c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                        tmem_st_x16(o_addr + c0, o);
                    }
                    tmem_st_wait();
                }
            }
            l *= alpha;
            float sum0 = 0.f, sum1 = 0.f;
            const float nm = -m_used;
#pragma unroll
            for (int c0 = 0; c0 < AW_KT; c0 += 8) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = aw_ex2(fmaf(__uint_as_float(r[c0 + j]), c, nm));
                sum0 += (e[0] + e[1]) + (e[2] + e[3]);
                sum1 += (e[4] + e[5]) + (e[6] + e[7]);
                sts128(prow + ((((uint32_t)(c0 >> 3)) ^ ((uint32_t)row & 7u)) << 4),
                       pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            }
            l += sum0 + sum1;
            tcgen05_fence_before();
            fence_proxy_async_smem();               // P visible to the tensor core (async proxy)
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> out[b, q, half*256 .. +256) ----
        mbar_wait(pv_done, (ntiles - 1) & 1);
        tcgen05_fence_after();
        const float inv = 1.0f / l;
        const int q = q0 + row;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)b * p.o_bs + (long long)q * p.o_ss + half * AW_DV;
#pragma unroll 1
        for (int c0 = 0; c0 < AW_DV; c0 += 16) {
            uint32_t o[16];
            tmem_ld_x16(o_addr + c0, o);
            tmem_ld_wait();
            if (q < p.Sq) {
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[8 * gg + 0]) * inv, __uint_as_float(o[8 * gg + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(o[8 * gg + 2]) * inv, __uint_as_float(o[8 * gg + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(o[8 * gg + 4]) * inv, __uint_as_float(o[8 * gg + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(o[8 * gg + 6]) * inv, __uint_as_float(o[8 * gg + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c0 + 8 * gg) = u;
                }
            }
        }
    }

    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace sdb

using namespace sdb;

// q / k / v: bf16 [B, S, d] views with explicit batch / sequence strides (channels contiguous), one head of d = 256 or 512
// channels; out [B, Sq, d] bf16.  scale = d**-0.5 for the VAE AttnBlock.
extern "C" int sdb_attention_wide_fwd(const void* q, const void* k, const void* v, void* out, long long q_bs, long long q_ss,
                                      long long k_bs, long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss,
                                      int B, int Sq, int Sk, int d, float scale, void* stream) {
    SDB_REQUIRE(q && k && v && out, "attention_wide: null pointer");
    SDB_REQUIRE(B > 0 && Sq > 0 && Sk > 0, "attention_wide: empty problem");
    SDB_REQUIRE(d == 256 || d == 512, "attention_wide: d=%d unsupported (256 or 512)", d);
    SDB_REQUIRE(B <= 65535, "attention_wide: grid too large");
    SDB_REQUIRE(o_ss % 8 == 0 && o_bs % 8 == 0 && ((uintptr_t)out & 15) == 0, "attention_wide: output rows must be 16-byte aligned");
    CUtensorMap tmQ, tmK, tmV;
    int es[4] = {1, 1, 1, 1};
    {
        long long dims[4] = {d, 1, Sq, B};
        long long str[3] = {d, q_ss, q_bs};
        int box[4] = {64, 1, 128, 1};
        int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, es);
        if (rc) return rc;
    }
    {
        long long dims[4] = {d, 1, Sk, B};
        int box[4] = {64, 1, AW_KT, 1};
        long long strk[3] = {d, k_ss, k_bs};
        int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, es);
        if (rc) return rc;
        long long strv[3] = {d, v_ss, v_bs};
        rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, es);
        if (rc) return rc;
    }
    static bool attr_set_dev[64] = {false};
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    if (!attr_set_dev[cur_dev]) {
        cudaError_t e = cudaFuncSetAttribute(tc_attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM);
        if (e != cudaSuccess) { set_last_error("attention_wide: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set_dev[cur_dev] = true;
    }
    AttnWP p;
    p.out = out; p.o_bs = o_bs; p.o_ss = o_ss;
    p.Sq = Sq; p.Sk = Sk; p.d = d;
    p.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((unsigned)ceil_div(Sq, 128), (unsigned)(d / AW_DV), (unsigned)B);
    launch_pdl(tc_attention_wide_kernel, dim3(grid), dim3(256), AW_SMEM, (cudaStream_t)stream, tmQ, tmK, tmV, p);
    return check_launch("tc_attention_wide_kernel");
}
