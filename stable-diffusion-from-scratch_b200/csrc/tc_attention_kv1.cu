// sdb200 — cross-attention forward for a SHORT key sequence (Sk <= 128: the 77 CLIP tokens) on tcgen05.
//
// Replaces flash_attn_func(q, k, v) of CrossAttention.forward when `context` is the text conditioning
// (openai_model/attention.py:99-112 with k, v = to_k(context), to_v(context)): out = softmax(q k^T * scale) v per
// (batch, head), one key tile only.
//
// tc_attention_kernel walks KEY tiles and keeps one (query tile pair) per CTA.  With a single key tile that is one exposed
// latency chain per CTA — TMEM allocation, Q/K/V loads, Q K^T, 128 exponentials per row, P V, store — of ~6 us, seven waves of
// them per 64x64 call (profiles/r02_ncu_full_summary.txt: 41-50 us, tensor pipe 8 % busy, 5 % of the HBM floor).  Here a CTA
// keeps K and V of its (batch, head) resident and walks QUERY items (2 tiles x 128 rows) instead:
//   warp 0          TMA producer : K, V once; the Q tiles of item i + 1 as soon as Q K^T of item i has retired
//   warps 1, 3      MMA issuers  : one per query tile: S_g = Q_g K^T of item i + 1 is issued BEFORE O_g = P_g V of item i,
//                                  so the next scores are in TMEM when the softmax warps come back from their stores
//   warp 2          TMEM allocator
//   warps 4..11     softmax      : thread == query row; score row (NK = 80 or 128 columns: the 77 tokens are not padded to
//                                  128 exponentials) -> registers, exact row max (one key tile: no running max, no rescale),
//                                  p = exp2(s*c - m), P_g -> swizzled smem, then O_g / l -> bf16 -> global.
// No further barriers are needed between items: a thread arrives on p_full(i + 1) only after it has read O_g(i), so
// P_g V of item i + 1 cannot overwrite an accumulator that is still being read, and it writes P_g(i + 1) only after pv_done(i).
// Roofline: MUFU.EX2 (NK exponentials per row) / HBM (q in, out back: 4 B per query channel); algorithmic FLOP 4*B*H*Sq*Sk*d.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace sdb {

using namespace ptx;

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr);

#ifdef SDB_XATTN_TRACE
static long long* g_xattn_trace = nullptr;
extern "C" void sdb_xattn_set_trace(long long* ptr) { g_xattn_trace = ptr; }
// [10 rows: softmax warps 4..11 by (warp - 4), issuer g=0 -> 8, issuer g=1 -> 9][32 items][8 stamps], CTA (0,0,0) only
#define XA_TRACE(rowi, it_, slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (it_) < 32) \
    p.trace[((rowi) * 32 + (it_)) * 8 + (slot)] = clock64(); } while (0)
#else
#define XA_TRACE(rowi, it_, slot) do { } while (0)
#endif

struct XAttnP {
#ifdef SDB_XATTN_TRACE
    long long* trace;
#endif
    void* out;
    long long o_bs, o_ss, o_hs;
    int Sq, Sk, d, H;
    int items;              // query items (256 rows) per (batch, head)
    int chunk;              // items per CTA
    float scale_log2;
};

template <int DPAD, int NK>
struct XAttnCfg {
    static constexpr int NBLK = DPAD / 64;
    static constexpr int QTILE_BYTES = 128 * DPAD * 2;
    static constexpr int KTILE_BYTES = NK * DPAD * 2;
    static constexpr int QST = DPAD == 64 ? 2 : 1;              // Q stages (one item = two tiles)
    static constexpr int P_BYTES = 128 * 128 * 2;               // two [128 x 64] K-major blocks per query tile
    static constexpr int PBUF = DPAD == 64 ? 2 : 1;             // P buffers per query tile (shared memory allows two at DPAD = 64)
    static constexpr int SMEM_BYTES = 1024 + QST * 2 * QTILE_BYTES + 2 * KTILE_BYTES + 2 * PBUF * P_BYTES + 256;
    static constexpr int THREADS = 384;
    __host__ __device__ static constexpr int s_col(int g) { return g * 128; }
    __host__ __device__ static constexpr int o_col(int g) { return 256 + g * DPAD; }
};

#ifndef SDB_XATTN_DEBUG
#define SDB_XATTN_DEBUG 0       // measurement builds (tools/gpu_xattn_variants.sh): 1 no output stores, 2 no MUFU, 3 Q loaded once, 4 no P stores
#endif
__device__ __forceinline__ float xa_ex2(float x) {
#if SDB_XATTN_DEBUG == 2
    return x * 0.001f + 1.0f;
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
__device__ __forceinline__ float xa_max3(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}

template <int DPAD, int NK>
__global__ void __launch_bounds__(384, 1)
tc_attention_kv1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const XAttnP p) {
    using Cfg = XAttnCfg<DPAD, NK>;
    constexpr int NBLK = Cfg::NBLK;
    constexpr int QST = Cfg::QST;
    constexpr int PBUF = Cfg::PBUF;
    constexpr uint32_t KBLK16 = (NK * 128) >> 4;               // one 64-channel block of the K / V tile, in 16-byte units
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                         // [QST][2] tiles of 128 rows
    uint8_t* sK = sQ + QST * 2 * Cfg::QTILE_BYTES;              // NK rows
    uint8_t* sV = sK + Cfg::KTILE_BYTES;
    uint8_t* sP = sV + Cfg::KTILE_BYTES;                        // [2][PBUF] 128 x 128 bf16
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * PBUF * Cfg::P_BYTES);
    uint64_t* q_full = bars;                // QST
    uint64_t* q_empty = q_full + QST;       // QST  (both issuers)
    uint64_t* kv_full = q_empty + QST;      // 1
    uint64_t* s_full = kv_full + 1;         // 2
    uint64_t* s_empty = s_full + 2;         // 2
    uint64_t* p_full = s_empty + 2;         // 2
    uint64_t* pv_done = p_full + 2;         // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int item0 = blockIdx.x * p.chunk;
    const int nit = min(p.chunk, p.items - item0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        for (int s = 0; s < QST; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 2); }
        mbar_init(kv_full, 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 128);
            mbar_init(&p_full[g], 128); mbar_init(&pv_done[g], 1);
        }
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
    pdl_wait();                                      // q (and k / v) written by the preceding projections are visible

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0 && nit > 0) {
            mbar_arrive_expect_tx(kv_full, 2 * Cfg::KTILE_BYTES);
#pragma unroll
            for (int j = 0; j < NBLK; ++j) {
                tma_load_4d(sK + j * (NK * 128), &tmK, kv_full, j * 64, h, 0, b);
                tma_load_4d(sV + j * (NK * 128), &tmV, kv_full, j * 64, h, 0, b);
            }
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < nit; ++it) {
                mbar_wait(&q_empty[s], ph ^ 1);
#if SDB_XATTN_DEBUG == 3
                if (it >= QST) { mbar_arrive(&q_full[s]); if (++s == QST) { s = 0; ph ^= 1; } continue; }
#endif
                mbar_arrive_expect_tx(&q_full[s], 2 * Cfg::QTILE_BYTES);
                const int q0 = (item0 + it) * 256;
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int j = 0; j < NBLK; ++j)
                        tma_load_4d(sQ + (s * 2 + g) * Cfg::QTILE_BYTES + j * 16384, &tmQ, &q_full[s], j * 64, h, q0 + g * 128, b);
                if (++s == QST) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ================= MMA issuers (whole warp in the loop, one elected lane issues) =================
        const int g = warp == 1 ? 0 : 1;
        const uint32_t idesc_s = umma_idesc_bf16(NK, false, false);       // S = Q K^T : N = NK keys
        const uint32_t idesc_o = umma_idesc_bf16(DPAD, false, true);      // O = P V   : B (V) is MN-major
        const int ksteps_qk = (p.d + 15) / 16;
        const uint64_t qdesc0 = umma_desc_kmajor_sw128(smem_u32(sQ));
        const uint64_t kdesc0 = umma_desc_kmajor_sw128(smem_u32(sK));
        const uint64_t vdesc0 = umma_desc_mnmajor_sw128(smem_u32(sV), NK * 128);
        const uint64_t pdesc0 = umma_desc_kmajor_sw128(smem_u32(sP + g * PBUF * Cfg::P_BYTES));
        auto issue_qk = [&](int it) {
            const int s = it % QST;
            XA_TRACE(8 + g, it, 0);
            mbar_wait(&q_full[s], (uint32_t)(it / QST) & 1u);
            XA_TRACE(8 + g, it, 1);
            mbar_wait(&s_empty[g], (uint32_t)(it & 1) ^ 1u);              // the softmax warps hold S_g(it - 1) in registers
            tcgen05_fence_after();
            XA_TRACE(8 + g, it, 2);
            const uint64_t qdesc = qdesc0 + (uint64_t)((s * 2 + g) * (Cfg::QTILE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < DPAD / 16; ++k) {
                if (k < ksteps_qk) {
                    const uint64_t offq = (uint64_t)((k >> 2) * 1024 + (k & 3) * 2);
                    const uint64_t offk = (uint64_t)((k >> 2) * KBLK16 + (k & 3) * 2);
                    if (elect_one()) umma_bf16_ss(tmem + Cfg::s_col(g), qdesc + offq, kdesc0 + offk, idesc_s, k > 0 ? 1u : 0u);
                }
            }
            if (elect_one()) { umma_commit(&s_full[g]); umma_commit(&q_empty[s]); }
            XA_TRACE(8 + g, it, 3);
        };
        if (nit > 0) {
            mbar_wait(kv_full, 0);
            issue_qk(0);
            for (int it = 0; it < nit; ++it) {
                if (it + 1 < nit) issue_qk(it + 1);
                XA_TRACE(8 + g, it, 4);
                mbar_wait(&p_full[g], (uint32_t)it & 1u);
                tcgen05_fence_after();
                XA_TRACE(8 + g, it, 5);
                const uint64_t pdesc = pdesc0 + (uint64_t)((it % PBUF) * (Cfg::P_BYTES >> 4));
#pragma unroll
                for (int k = 0; k < NK / 16; ++k) {
                    if (elect_one())
                        umma_bf16_ss(tmem + Cfg::o_col(g), pdesc + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), vdesc0 + (uint64_t)(k * 128),
                                     idesc_o, k > 0 ? 1u : 0u);
                }
                if (elect_one()) umma_commit(&pv_done[g]);
                XA_TRACE(8 + g, it, 6);
            }
        }
    } else if (warp >= 4) {
        // ================= softmax + output =================
        const int g = (warp - 4) >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
        const uint32_t s_addr = lane_addr + Cfg::s_col(g);
        const uint32_t o_addr = lane_addr + Cfg::o_col(g);
        const uint32_t prow = smem_u32(sP + g * PBUF * Cfg::P_BYTES) + row * 128;
        const float c = p.scale_log2;
        // scores of item `it` -> P_g (buffer it % PBUF) in shared memory; returns 1 / row sum.  Does NOT publish P.
        auto softmax_item = [&](int it) -> float {
            XA_TRACE(warp - 4, it, 0);
            mbar_wait(&s_full[g], (uint32_t)it & 1u);
            tcgen05_fence_after();
            XA_TRACE(warp - 4, it, 1);
            uint32_t r[NK];
#pragma unroll
            for (int c0 = 0; c0 + 32 <= NK; c0 += 32) tmem_ld_x32(s_addr + c0, r + c0);
            if (NK % 32 == 16) tmem_ld_x16(s_addr + (NK - 16), r + (NK - 16));
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&s_empty[g]);
            XA_TRACE(warp - 4, it, 2);
            // keys >= Sk are padding (their K rows are TMA zero fill): -inf.  Usually only the last 16 columns can be affected.
            if (p.Sk < NK) {
                if (p.Sk >= NK - 16) {
#pragma unroll
                    for (int j = NK - 16; j < NK; ++j) if (j >= p.Sk) r[j] = 0xff800000u;
                } else {
#pragma unroll
                    for (int j = 0; j < NK; ++j) if (j >= p.Sk) r[j] = 0xff800000u;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < NK; j += 4) {
                mx0 = xa_max3(mx0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                mx1 = xa_max3(mx1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
            const float nm = -fmaxf(mx0, mx1) * c;
            XA_TRACE(warp - 4, it, 3);
            const uint32_t pr = prow + (uint32_t)(it % PBUF) * Cfg::P_BYTES;
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < NK; c0 += 8) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = xa_ex2(fmaf(__uint_as_float(r[c0 + j]), c, nm));
                sum0 += (e[0] + e[1]) + (e[2] + e[3]);
                sum1 += (e[4] + e[5]) + (e[6] + e[7]);
                const int chunk = (c0 & 63) >> 3;
#if SDB_XATTN_DEBUG == 4
                if (e[0] == 123.f)
#endif
                sts128(pr + (c0 >> 6) * 16384 + ((chunk ^ (row & 7)) << 4),
                       pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            }
            XA_TRACE(warp - 4, it, 4);
            return 1.0f / (sum0 + sum1);
        };
        // P_g of the item just written -> visible to the tensor core; also orders this thread's reads of O_g before P_g V of that item
        auto publish_p = [&]() {
            tcgen05_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(&p_full[g]);
        };
        // O_g(it) / l -> bf16 -> out[b, q, h, 0..d)
        auto output_item = [&](int it, float inv) {
            XA_TRACE(warp - 4, it, 5);
            mbar_wait(&pv_done[g], (uint32_t)it & 1u);
            tcgen05_fence_after();
            XA_TRACE(warp - 4, it, 6);
            const int q = (item0 + it) * 256 + g * 128 + row;
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)b * p.o_bs + (long long)q * p.o_ss + (long long)h * p.o_hs;
            // two 16-column loads in flight per wait
#pragma unroll
            for (int c0 = 0; c0 < DPAD; c0 += 32) {
                if (c0 < p.d) {                                        // warp-uniform
                    uint32_t o[32];
                    tmem_ld_x16(o_addr + c0, o);
                    if (c0 + 16 < p.d) tmem_ld_x16(o_addr + c0 + 16, o + 16);
                    tmem_ld_wait();
                    if (q < p.Sq) {
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            if (c0 + 8 * gg < p.d) {
                                uint4 u;
                                u.x = pack_bf16x2(__uint_as_float(o[8 * gg + 0]) * inv, __uint_as_float(o[8 * gg + 1]) * inv);
                                u.y = pack_bf16x2(__uint_as_float(o[8 * gg + 2]) * inv, __uint_as_float(o[8 * gg + 3]) * inv);
                                u.z = pack_bf16x2(__uint_as_float(o[8 * gg + 4]) * inv, __uint_as_float(o[8 * gg + 5]) * inv);
                                u.w = pack_bf16x2(__uint_as_float(o[8 * gg + 6]) * inv, __uint_as_float(o[8 * gg + 7]) * inv);
#if SDB_XATTN_DEBUG == 1
                                if (u.x == 0x12345678u)
#endif
                                *reinterpret_cast<uint4*>(orow + c0 + 8 * gg) = u;
                            }
                        }
                    }
                }
            }
            XA_TRACE(warp - 4, it, 7);
        };
        if (PBUF == 2) {
            // two P buffers: the softmax of item it + 1 runs while P_g V of item it is in flight; its P is published only after
            // O_g(it) has been read (the accumulator is single), so the chain per item is softmax + output, not + the MMA round trip
            float inv = 0.f;
            if (nit > 0) { inv = softmax_item(0); publish_p(); }
            for (int it = 0; it < nit; ++it) {
                float inv_next = 0.f;
                if (it + 1 < nit) inv_next = softmax_item(it + 1);     // buffer (it + 1) & 1: last read by P V of item it - 1 (retired)
                output_item(it, inv);
                if (it + 1 < nit) publish_p();
                inv = inv_next;
            }
        } else {
            for (int it = 0; it < nit; ++it) {
                // P_g is free: this thread waited for pv_done(it - 1) before it read O_g(it - 1)
                const float inv = softmax_item(it);
                publish_p();
                output_item(it, inv);
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

static int xattn_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

template <int DPAD, int NK>
static int launch_xattn(const sdb_attn_args* a, cudaStream_t st) {
    using Cfg = XAttnCfg<DPAD, NK>;
    CUtensorMap tmQ, tmK, tmV;
    int es[4] = {1, 1, 1, 1};
    {
        int box[4] = {64, 1, 128, 1};
        long long dims[4] = {a->dense ? a->d : DPAD, a->H, a->Sq, a->B};
        long long str[3] = {a->q_hs, a->q_ss, a->q_bs};
        int rc = make_tmap_bf16(&tmQ, a->q, 4, dims, str, box, es);
        if (rc) return rc;
    }
    {
        int box[4] = {64, 1, NK, 1};
        long long dims[4] = {a->dense ? a->d : DPAD, a->H, a->Sk, a->B};
        long long str[3] = {a->k_hs, a->k_ss, a->k_bs};
        int rc = make_tmap_bf16(&tmK, a->k, 4, dims, str, box, es);
        if (rc) return rc;
        long long strv[3] = {a->v_hs, a->v_ss, a->v_bs};
        rc = make_tmap_bf16(&tmV, a->v, 4, dims, strv, box, es);
        if (rc) return rc;
    }
    static bool attr_set_dev[64] = {false};
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    bool& attr_set = attr_set_dev[cur_dev];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_attention_kv1_kernel<DPAD, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("attention(kv1): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    XAttnP p;
#ifdef SDB_XATTN_TRACE
    p.trace = g_xattn_trace;
#endif
    p.out = a->out; p.o_bs = a->o_bs; p.o_ss = a->o_ss; p.o_hs = a->o_hs;
    p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d; p.H = a->H;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.items = ceil_div(a->Sq, 256);
    // CTAs per (batch, head): as many as fit in ONE wave (a second wave would pay the whole start-up chain again)
    const long long bh = (long long)a->B * a->H;
    int per_bh = (int)(xattn_sm_count() / bh);
    if (per_bh < 1) per_bh = 1;
    if (per_bh > p.items) per_bh = p.items;
    p.chunk = ceil_div(p.items, per_bh);
    dim3 grid((unsigned)ceil_div(p.items, p.chunk), (unsigned)a->H, (unsigned)a->B);
    launch_pdl(tc_attention_kv1_kernel<DPAD, NK>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmQ, tmK, tmV, p);
    return check_launch("tc_attention_kv1_kernel");
}

// Entry from sdb_attention_fwd (tc_attention.cu): returns 1 when this kernel does not cover the problem.
static int g_kv1_enabled = -1;
static int kv1_enabled() {
    if (g_kv1_enabled < 0) {
        const char* e = getenv("SDB200_XATTN");
        g_kv1_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    return g_kv1_enabled;
}

int attention_kv1_dispatch(const sdb_attn_args* a, cudaStream_t st) {
    if (!kv1_enabled() || a->causal || a->Sk > 128 || a->dpad > 128) return 1;
    const bool nk80 = a->Sk <= 80;
    if (a->dpad == 64) return nk80 ? launch_xattn<64, 80>(a, st) : launch_xattn<64, 128>(a, st);
    return nk80 ? launch_xattn<128, 80>(a, st) : launch_xattn<128, 128>(a, st);
}

}  // namespace sdb

extern "C" int sdb_attention_set_short_key_kernel(int enable) {
    const int prev = sdb::kv1_enabled();
    sdb::g_kv1_enabled = enable ? 1 : 0;
    return prev;
}
