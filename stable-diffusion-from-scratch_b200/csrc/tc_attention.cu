// sdb200 — fused attention forward on tcgen05 (bf16 operands, fp32 softmax statistics).
//
// Replaces flash_attn_func(q, k, v, softmax_scale, causal=False) as called by CrossAttention
// (openai_model/attention.py:99-112): out = softmax(q k^T * scale) v per (batch, head).
//
// One CTA = 128 query rows of one (batch, head); 192 threads:
//   warp 0      TMA producer : Q once, then K_j / V_j tiles (128 keys) into a smem ring
//   warp 1      MMA issuer   : S_j = Q K_j^T (TMEM, double buffered), O += P_j V_j (TMEM)
//   warps 2..5  softmax      : thread == query row (TMEM lane): two passes over S_j via tcgen05.ld
//                              (row max, then exp2 + row sum), P_j -> smem as bf16 in the 128B-swizzled
//                              K-major layout the PV MMA reads; O rescaled in TMEM only when the
//                              running max moved by more than 2^8 (lazy rescale, warp-uniform vote).
// Heads are zero-padded in memory from d to DPAD (multiple of 64) so every operand tile is a whole
// number of 128-byte swizzle rows.  V is consumed MN-major (d contiguous) straight from its TMA tile.
// Roofline: tensor pipe for d >= 80; for d = 40 the MUFU.EX2 rate (128 exp per row-tile) bounds it.
#include "common.cuh"
#include "ptx.cuh"
#include <string.h>

namespace sdb {

using namespace ptx;

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr);

struct AttnP {
    void* out;
    long long o_bs, o_ss, o_hs;
    int Sq, Sk, d, H;
    float scale_log2;
};

template <int DPAD>
struct AttnCfg {
    static constexpr int NBLK = DPAD / 64;                 // 64-wide column blocks per operand tile
    static constexpr int TILE_BYTES = 128 * DPAD * 2;      // Q / K / V tile
    static constexpr int STAGES = DPAD == 64 ? 3 : (DPAD == 128 ? 2 : 1);
    static constexpr int P_BYTES = 128 * 128 * 2;
    static constexpr int SMEM_BYTES = 1024 + TILE_BYTES * (1 + 2 * STAGES) + P_BYTES + 512;
    static constexpr int TMEM_COLS = 512;
    static constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int DPAD>
__global__ void __launch_bounds__(192, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnP p) {
    using Cfg = AttnCfg<DPAD>;
    constexpr int ST = Cfg::STAGES;
    constexpr int NBLK = Cfg::NBLK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Cfg::TILE_BYTES;
    uint8_t* sV = sK + ST * Cfg::TILE_BYTES;
    uint8_t* sP = sV + ST * Cfg::TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // ST
    uint64_t* k_empty = k_full + ST;    // ST
    uint64_t* v_full = k_empty + ST;    // ST
    uint64_t* v_empty = v_full + ST;    // ST
    uint64_t* s_full = v_empty + ST;    // 2
    uint64_t* s_empty = s_full + 2;     // 2
    uint64_t* p_full = s_empty + 2;     // 1
    uint64_t* pv_done = p_full + 1;     // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int ntiles = (p.Sk + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ST; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); }
        mbar_init(p_full, 128);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, Cfg::TILE_BYTES);
#pragma unroll
            for (int j = 0; j < NBLK; ++j) tma_load_4d(sQ + j * 16384, &tmQ, q_full, j * 64, h, q0, b);
            int s = 0; uint32_t ph = 0;
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[s], Cfg::TILE_BYTES);
#pragma unroll
                for (int j = 0; j < NBLK; ++j) tma_load_4d(sK + s * Cfg::TILE_BYTES + j * 16384, &tmK, &k_full[s], j * 64, h, t * 128, b);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[s], Cfg::TILE_BYTES);
#pragma unroll
                for (int j = 0; j < NBLK; ++j) tma_load_4d(sV + s * Cfg::TILE_BYTES + j * 16384, &tmV, &v_full[s], j * 64, h, t * 128, b);
                if (++s == ST) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, false, false);     // S = Q K^T : N = 128 keys
            const uint32_t idesc_o = umma_idesc_bf16(DPAD, false, true);     // O = P V   : B (V) is MN-major
            const int ksteps_qk = (p.d + 15) / 16;
            mbar_wait(q_full, 0);
            int s = 0; uint32_t ph = 0;        // K ring
            int sv = 0; uint32_t phv = 0;      // V ring (lags by one tile)
            for (int t = 0; t <= ntiles; ++t) {
                if (t < ntiles) {
                    const int sb = t & 1;
                    mbar_wait(&k_full[s], ph);
                    mbar_wait(&s_empty[sb], ((t >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + s * Cfg::TILE_BYTES);
                    for (int k = 0; k < ksteps_qk; ++k) {
                        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
                        umma_bf16_ss(tmem + (sb ? Cfg::S_COL1 : Cfg::S_COL0), umma_desc_kmajor_sw128(qa + off),
                                     umma_desc_kmajor_sw128(ka + off), idesc_s, k > 0 ? 1u : 0u);
                    }
                    umma_commit(&s_full[sb]);
                    umma_commit(&k_empty[s]);
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
                if (t >= 1) {
                    const int tp = t - 1;
                    mbar_wait(p_full, tp & 1);
                    mbar_wait(&v_full[sv], phv);
                    tcgen05_fence_after();
                    const uint32_t pa = smem_u32(sP), va = smem_u32(sV + sv * Cfg::TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {      // 128 keys = 8 x K16
                        const uint64_t adesc = umma_desc_kmajor_sw128(pa + (k >> 2) * 16384 + (k & 3) * 32);
                        const uint64_t bdesc = umma_desc_mnmajor_sw128(va + k * 2048, 16384);
                        umma_bf16_ss(tmem + Cfg::O_COL, adesc, bdesc, idesc_o, (tp > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(pv_done);
                    umma_commit(&v_empty[sv]);
                    if (++sv == ST) { sv = 0; phv ^= 1; }
                }
            }
        }
    } else {
        // ================= softmax / correction / epilogue =================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
        float m_used = -INFINITY, l = 0.f;
        const float c = p.scale_log2;
        for (int t = 0; t < ntiles; ++t) {
            const int sb = t & 1;
            const uint32_t s_addr = lane_addr + (sb ? Cfg::S_COL1 : Cfg::S_COL0);
            const int kvalid = p.Sk - t * 128;         // keys >= kvalid are padding
            mbar_wait(&s_full[sb], (t >> 1) & 1);
            tcgen05_fence_after();
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(s_addr + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]);
                    if (c0 + j < kvalid) mx = fmaxf(mx, v);
                }
            }
            mx *= c;
            const bool need = mx > m_used + 8.0f;
            float alpha = 1.0f;
            if (need) { alpha = ex2_approx(m_used - mx); m_used = mx; }
            const bool any_need = __any_sync(0xffffffffu, need);
            // O and the P buffer are free once PV_{t-1} has retired
            if (t > 0) {
                mbar_wait(pv_done, (t - 1) & 1);
                tcgen05_fence_after();
                if (any_need) {
#pragma unroll 1
                    for (int c0 = 0; c0 < DPAD; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld_x16(lane_addr + Cfg::O_COL + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
                        tmem_st_x16(lane_addr + Cfg::O_COL + c0, r);
                    }
                    tmem_st_wait();
                }
            }
            l *= alpha;
            // pass 2: p = exp2(s*c - m), row sum, write P (bf16, K-major SW128: two [128 x 64] blocks)
            float sum = 0.f;
            uint8_t* prow = sP + row * 128;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(s_addr + c0, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float p0 = (c0 + j < kvalid) ? ex2_approx(fmaf(__uint_as_float(r[j]), c, -m_used)) : 0.f;
                    float p1 = (c0 + j + 1 < kvalid) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), c, -m_used)) : 0.f;
                    sum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                uint8_t* blk = prow + (c0 >> 6) * 16384;
#pragma unroll
                for (int g = 0; g < 4; ++g) {       // 4 x 16-byte chunks (8 keys each)
                    int chunk = ((c0 & 63) >> 3) + g;
                    uint4 u = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    *reinterpret_cast<uint4*>(blk + ((chunk ^ (row & 7)) << 4)) = u;
                }
            }
            l += sum;
            tcgen05_fence_before();
            mbar_arrive(&s_empty[sb]);              // S_t fully read
            fence_proxy_async_smem();               // P visible to the tensor core (async proxy)
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> out[b, q, h, 0..d) ----
        mbar_wait(pv_done, (ntiles - 1) & 1);
        tcgen05_fence_after();
        const float inv = 1.0f / l;
        const int q = q0 + row;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)b * p.o_bs + (long long)q * p.o_ss + (long long)h * p.o_hs;
#pragma unroll 1
        for (int c0 = 0; c0 < DPAD; c0 += 16) {
            uint32_t r[16];
            tmem_ld_x16(lane_addr + Cfg::O_COL + c0, r);
            tmem_ld_wait();
            if (q < p.Sq && c0 < p.d) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (c0 + 8 * g < p.d) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c0 + 8 * g) = u;
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, Cfg::TMEM_COLS);
    }
}

template <int DPAD>
static int launch_attn(const sdb_attn_args* a, cudaStream_t st) {
    using Cfg = AttnCfg<DPAD>;
    CUtensorMap tmQ, tmK, tmV;
    int box[4] = {64, 1, 128, 1};
    int es[4] = {1, 1, 1, 1};
    {
        long long dims[4] = {DPAD, a->H, a->Sq, a->B};
        long long str[3] = {a->q_hs, a->q_ss, a->q_bs};
        int rc = make_tmap_bf16(&tmQ, a->q, 4, dims, str, box, es);
        if (rc) return rc;
    }
    {
        long long dims[4] = {DPAD, a->H, a->Sk, a->B};
        long long str[3] = {a->k_hs, a->k_ss, a->k_bs};
        int rc = make_tmap_bf16(&tmK, a->k, 4, dims, str, box, es);
        if (rc) return rc;
        long long strv[3] = {a->v_hs, a->v_ss, a->v_bs};
        rc = make_tmap_bf16(&tmV, a->v, 4, dims, strv, box, es);
        if (rc) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_attention_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    AttnP p;
    p.out = a->out; p.o_bs = a->o_bs; p.o_ss = a->o_ss; p.o_hs = a->o_hs;
    p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d; p.H = a->H;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    dim3 grid((unsigned)ceil_div(a->Sq, 128), (unsigned)a->H, (unsigned)a->B);
    tc_attention_kernel<DPAD><<<grid, 192, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    return check_launch("tc_attention_kernel");
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention_fwd(const sdb_attn_args* a, void* stream) {
    SDB_REQUIRE(a && a->q && a->k && a->v && a->out, "attention: null pointer");
    SDB_REQUIRE(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Sk > 0 && a->d > 0, "attention: empty problem");
    SDB_REQUIRE(a->d % 8 == 0 && a->dpad % 64 == 0 && a->dpad >= a->d && a->dpad <= 192, "attention: d=%d dpad=%d unsupported", a->d, a->dpad);
    SDB_REQUIRE(a->B <= 65535 && a->H <= 65535, "attention: grid too large");
    SDB_REQUIRE(a->o_hs % 8 == 0 && a->o_ss % 8 == 0 && a->o_bs % 8 == 0 && ((uintptr_t)a->out & 15) == 0, "attention: output must be 16-byte aligned per head row");
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->dpad) {
        case 64: return launch_attn<64>(a, st);
        case 128: return launch_attn<128>(a, st);
        default: return launch_attn<192>(a, st);
    }
}
