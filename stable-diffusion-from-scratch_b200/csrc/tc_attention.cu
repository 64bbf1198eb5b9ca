// sdb200 — fused attention forward on tcgen05 (bf16 operands, fp32 softmax statistics).
//
// Replaces flash_attn_func(q, k, v, softmax_scale, causal=False) as called by CrossAttention
// (openai_model/attention.py:99-112): out = softmax(q k^T * scale) v per (batch, head).
//
// One CTA = G query tiles of 128 rows of one (batch, head) sharing every K/V tile (G = 2 when TMEM
// allows: 2 x 128 S columns + 2 x DPAD O columns <= 512), 128 + 128*G threads:
//   warp 0           TMA producer : Q tiles once, then K_j / V_j tiles (128 keys) into a smem ring
//   warp 1           MMA issuer   : S_g = Q_g K_j^T (TMEM), O_g += P_g V_j (TMEM); the tensor pipe
//                                   alternates between the query tiles
//   warp 2           TMEM allocator
//   warps 4..4+4G-1  softmax      : thread == query row (TMEM lane).  The whole 128-key score row is read
//                                   from TMEM once into registers (S_g is then free for the next Q K^T),
//                                   row max with 3-input max, p = exp2(s*c - m), row sum, P_g -> smem as
//                                   bf16 in the 128B-swizzled K-major layout the PV MMA reads; O_g is
//                                   rescaled in TMEM only when the running max moved by more than 2^8
//                                   (lazy rescale, warp-uniform vote).
// With two query tiles every SM sub-partition holds two softmax warps, so one warp's TMEM / barrier /
// MUFU latencies are hidden by the other's issue slots and K/V smem traffic per MMA halves.
// Heads are zero-padded in memory from d to DPAD (multiple of 64) so every operand tile is a whole
// number of 128-byte swizzle rows.  V is consumed MN-major (d contiguous) straight from its TMA tile.
// Roofline: tensor pipe for d >= 80; for d = 40 the MUFU.EX2 rate (16 exp/clk/SM, 128 x 128 exp per
// tile = 1024 clk against 448 clk of MMA) bounds it.
#include "common.cuh"
#include "ptx.cuh"
#include <string.h>

namespace sdb {

using namespace ptx;

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr);
int attention_kv1_dispatch(const sdb_attn_args* a, cudaStream_t st);   // tc_attention_kv1.cu: one key tile, CTA walks query items

#ifdef SDB_ATTN_TRACE
static long long* g_attn_trace = nullptr;
extern "C" void sdb_attn_set_trace(long long* ptr) { g_attn_trace = ptr; }
#define AT_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && t < 32) \
    p.trace[((warp - 4) * 32 + t) * 8 + (slot)] = clock64(); } while (0)
#define AT_TRACE_MMA(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t < 32 && lane == 0) \
    p.trace[(8 * 32 + t) * 8 + (slot)] = clock64(); } while (0)
#define AT_TRACE_ISS(slot) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t < 32 && g_lo == 0 && lane == 0) \
    p.trace[(9 * 32 + t) * 8 + (slot)] = clock64(); } while (0)
#else
#define AT_TRACE(slot) do { } while (0)
#define AT_TRACE_MMA(slot) do { } while (0)
#define AT_TRACE_ISS(slot) do { } while (0)
#endif

struct AttnP {
#ifdef SDB_ATTN_TRACE
    long long* trace;
#endif
    void* out;
    long long o_bs, o_ss, o_hs;
    int Sq, Sk, d, H;
    float scale_log2;
    int causal;             // query q sees keys 0 .. q (CLIP text tower); 0 = full attention
};

template <int DPAD>
struct AttnCfg {
    static constexpr int G = (256 + 2 * DPAD <= 512) ? 2 : 1;   // query tiles per CTA (TMEM: G * (128 + DPAD) columns)
    static constexpr int NBLK = DPAD / 64;                      // 64-wide column blocks per operand tile
    static constexpr int TILE_BYTES = 128 * DPAD * 2;           // one Q / K / V tile
    static constexpr int STAGES = DPAD == 64 ? 2 : 1;
    static constexpr int P_BYTES = 128 * 128 * 2;
    // d <= 64: two P buffers per query tile, so the softmax of key tile t never waits for P_{t-1} V to retire (that
    // wait left the MUFU pipe idle ~40 % of the time, profiles/r01_ncu_full_summary.txt); larger heads have no smem for it
    static constexpr int PBUF = DPAD == 64 ? 2 : 1;
    static constexpr int SMEM_BYTES = 1024 + TILE_BYTES * (G + 2 * STAGES) + G * PBUF * P_BYTES + 512;
    static constexpr int TMEM_COLS = 512;
    static constexpr int THREADS = 128 + 128 * G;
    __host__ __device__ static constexpr int s_col(int g) { return g * 128; }
    __host__ __device__ static constexpr int o_col(int g) { return G * 128 + g * DPAD; }
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifndef SDB_ATTN_VARIANT
#define SDB_ATTN_VARIANT 0      // measurement builds: bit 0 = 2-input max, bit 1 = integer bf16 pack (tools/gpu_attn_variants.sh)
#endif
__device__ __forceinline__ float fmax3(float a, float b, float c) {
#if SDB_ATTN_VARIANT & 1
    return fmaxf(fmaxf(a, b), c);
#else
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
#endif
}
// Chunks (of 8 exponentials, 16 per key tile) after which a softmax warp hands the MUFU phase to its partner on the same
// scheduler.  16 = strict alternation.  One warp alone sustains only ~12 clk per exponential (dependency-bound) against the
// pipe's 8, two concurrent warps 9.4 combined (tools/ubench/xu_rate.cu), so a partial overlap fills the idle MUFU slots.
#ifndef SDB_ATTN_HANDOVER
#define SDB_ATTN_HANDOVER 11
#endif
#ifndef SDB_ATTN_MAXCHAINS
#define SDB_ATTN_MAXCHAINS 2
#endif
#ifndef SDB_ATTN_POLY
#define SDB_ATTN_POLY 0         // exponentials out of every 8 evaluated on the FMA pipe instead of MUFU.EX2
#endif
// 2^x on the FMA pipe: x = n + f with n = round(x), f in [-0.5, 0.5]; cubic for 2^f (relative error 1.0e-4, 20x below the
// bf16 rounding of P), exponent added with integer arithmetic.  x <= 8 here (lazy rescale), clamped at -125 from below.
__device__ __forceinline__ float ex2_fma(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
    const float f = x - (t - 12582912.0f);
    float q = fmaf(0.05583828315138817f, f, 0.2426394820213318f);
    q = fmaf(q, f, 0.6931367516517639f);
    q = fmaf(q, f, 0.9999245405197144f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
template <int J>
__device__ __forceinline__ float ex2_sel(float x) {
    constexpr bool poly = (SDB_ATTN_POLY == 1 && J == 7) || (SDB_ATTN_POLY == 2 && (J == 3 || J == 7)) ||
                          (SDB_ATTN_POLY == 3 && (J == 2 || J == 5 || J == 7)) || (SDB_ATTN_POLY == 4 && (J & 1));
    return poly ? ex2_fma(x) : ex2_approx(x);
}
#ifndef SDB_ATTN_ONES
#define SDB_ATTN_ONES 1         // row sums from the tensor core: pad channel d of every V tile is set to 1, so O[:, d] = sum_j P_ij
#endif
#ifndef SDB_ATTN_PACKED
#define SDB_ATTN_PACKED 1       // FFMA2 / FADD2 in the exponential phase (0 = scalar fp32, the round-1 form)
#endif
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// two non-negative fp32 -> packed bf16x2 (lo in the low half), round-to-nearest
__device__ __forceinline__ uint32_t pack_p_bf16x2(float lo, float hi) {
#if SDB_ATTN_VARIANT & 2
    // integer rounding (half away from zero on the magnitude; the inputs are probabilities in [0, 1]): 2 IADD + 1 PRMT
    const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
    return __byte_perm(a, b, 0x7632);
#else
    return pack_bf16x2(lo, hi);
#endif
}

// ONES: head_dim < DPAD, so channel d of every V tile can be set to 1 and the row sums come out of the PV MMA as O[:, d]
template <int DPAD, bool ONES>
__global__ void __launch_bounds__(AttnCfg<DPAD>::THREADS, 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnP p) {
    using Cfg = AttnCfg<DPAD>;
    constexpr int ST = Cfg::STAGES;
    constexpr int NBLK = Cfg::NBLK;
    constexpr int G = Cfg::G;
    constexpr int PB = Cfg::PBUF;
    constexpr int NI = (G == 2 && DPAD == 64) ? 2 : 1;      // MMA-issuing threads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                   // [G] tiles
    uint8_t* sK = sQ + G * Cfg::TILE_BYTES;               // [ST]
    uint8_t* sV = sK + ST * Cfg::TILE_BYTES;              // [ST]
    uint8_t* sP = sV + ST * Cfg::TILE_BYTES;              // [G][PB] 128 x 128 bf16
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + G * PB * Cfg::P_BYTES);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // ST
    uint64_t* k_empty = k_full + ST;    // ST
    uint64_t* v_full = k_empty + ST;    // ST
    uint64_t* v_empty = v_full + ST;    // ST
    uint64_t* s_full = v_empty + ST;    // G   S_g written by the MMA
    uint64_t* s_empty = s_full + G;     // G   S_g copied to registers by its 128 softmax threads
    uint64_t* p_full = s_empty + G;     // G*PB  P_g (buffer b) written to smem
    uint64_t* pv_done = p_full + G * PB;   // G*PB  O_g += P_g[b] V retired (that P buffer is free; with PB = 1 also O_g)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + G * PB);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (128 * G), h = blockIdx.y, b = blockIdx.z;
    const int ntiles = (p.Sk + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ST; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], NI);   // every MMA issuer releases the slot
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], NI);
        }
        for (int g = 0; g < G; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 128); }
        for (int i = 0; i < G * PB; ++i) { mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                      // q / k / v written by the preceding projection are visible

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, G * Cfg::TILE_BYTES);
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int j = 0; j < NBLK; ++j)
                    tma_load_4d(sQ + g * Cfg::TILE_BYTES + j * 16384, &tmQ, q_full, j * 64, h, q0 + g * 128, b);
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
    } else if (warp == 1 || (NI == 2 && warp == 3)) {
        // ================= MMA issuers: one elected thread per issuer =================
        // A single thread issuing both tiles' 22 small MMAs per key tile (3 x QK^T, 8 x PV each, d = 40) took ~3000 cycles
        // per key tile (profiles/r01_attn_timeline.txt) — longer than the exponentials; the tensor pipe sat idle behind
        // the issue thread.  For d <= 64 two issuers (warp 1 -> tile 0, warp 3 -> tile 1) halve that chain and decouple the
        // tiles; descriptors are built once and every MMA just adds an offset.
        {   // the whole warp runs this loop (uniform control flow); one elected lane issues
            const int g_lo = (NI == 2 && warp == 3) ? 1 : 0;
            const int g_hi = NI == 2 ? g_lo + 1 : G;
            const uint32_t idesc_s = umma_idesc_bf16(128, false, false);     // S = Q K^T : N = 128 keys
            const uint32_t idesc_o = umma_idesc_bf16(DPAD, false, true);     // O = P V   : B (V) is MN-major
            const int ksteps_qk = (p.d + 15) / 16;
            constexpr uint32_t TILE16 = Cfg::TILE_BYTES >> 4;                // descriptor address units are 16 bytes
            const uint64_t qdesc0 = umma_desc_kmajor_sw128(smem_u32(sQ));
            const uint64_t kdesc0 = umma_desc_kmajor_sw128(smem_u32(sK));
            const uint64_t vdesc0 = umma_desc_mnmajor_sw128(smem_u32(sV), 16384);
            const uint64_t pdesc0 = umma_desc_kmajor_sw128(smem_u32(sP));
            constexpr bool ones = ONES && SDB_ATTN_ONES;
            mbar_wait(q_full, 0);
            int s = 0; uint32_t ph = 0;        // K ring
            int sv = 0; uint32_t phv = 0;      // V ring (lags by one tile)
            for (int t = 0; t <= ntiles; ++t) {
                if (t < ntiles) {
                    if (g_lo == 0) AT_TRACE_MMA(0);
                    mbar_wait(&k_full[s], ph);
                    if (g_lo == 0) AT_TRACE_MMA(1);
                    const uint64_t kdesc = kdesc0 + (uint64_t)(s * TILE16);
                    for (int g = g_lo; g < g_hi; ++g) {
                        mbar_wait(&s_empty[g], (t & 1) ^ 1);       // softmax g holds S_g(t-1) in registers
                        tcgen05_fence_after();
                        if (g == 0) AT_TRACE_MMA(2);
                        const uint64_t qdesc = qdesc0 + (uint64_t)(g * TILE16);
#pragma unroll
                        for (int k = 0; k < DPAD / 16; ++k) {
                            if (k < ksteps_qk) {                   // (k >> 2) * 16 KB column block + (k & 3) * 32 B inside the swizzle row
                                const uint64_t off = (uint64_t)((k >> 2) * 1024 + (k & 3) * 2);
                                if (elect_one()) umma_bf16_ss(tmem + Cfg::s_col(g), qdesc + off, kdesc + off, idesc_s, k > 0 ? 1u : 0u);
                            }
                        }
                        if (elect_one()) umma_commit(&s_full[g]);
                    }
                    if (elect_one()) umma_commit(&k_empty[s]);     // one arrival per issuer
                    if (g_lo == 0) AT_TRACE_MMA(3);
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
                if (t >= 1) {
                    const int tp = t - 1;
                    if (g_lo == 0) AT_TRACE_MMA(4);
                    mbar_wait(&v_full[sv], phv);
                    if (ones) {                      // (with two issuers both patch: same values, each fences before its own MMAs)
                        // channel d (a pad channel: TMA zero fill or zeros in memory) of all 128 key rows := 1.0 (bf16), in the
                        // swizzled layout the MMA reads: row r, 16-byte chunk (d / 8) ^ (r & 7) of 64-channel block d / 64
                        const uint32_t vb = smem_u32(sV) + (uint32_t)sv * Cfg::TILE_BYTES + (uint32_t)(p.d >> 6) * 16384u;
                        const uint32_t ch = (uint32_t)((p.d & 63) >> 3), within = (uint32_t)(p.d & 7) * 2u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t r = (uint32_t)(lane + 32 * i);
                            asm volatile("st.shared.b16 [%0], %1;" ::"r"(vb + r * 128u + ((ch ^ (r & 7u)) << 4) + within), "h"((unsigned short)0x3F80) : "memory");
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                    }
                    if (g_lo == 0) AT_TRACE_MMA(5);
                    const int pb = PB == 2 ? (tp & 1) : 0;         // P buffer of key tile tp
                    const uint64_t vdesc = vdesc0 + (uint64_t)(sv * TILE16);
                    for (int g = g_lo; g < g_hi; ++g) {
                        mbar_wait(&p_full[g * PB + pb], PB == 2 ? ((tp >> 1) & 1) : (tp & 1));
                        tcgen05_fence_after();
                        if (g == 0) AT_TRACE_MMA(6);
                        const uint64_t pdesc = pdesc0 + (uint64_t)((g * PB + pb) * (Cfg::P_BYTES >> 4));
                        AT_TRACE_ISS(0);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {              // 128 keys = 8 x K16
                            if (elect_one())
                                umma_bf16_ss(tmem + Cfg::o_col(g), pdesc + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), vdesc + (uint64_t)(k * 128),
                                             idesc_o, (tp > 0 || k > 0) ? 1u : 0u);
                            if (k == 0) AT_TRACE_ISS(1);
                            if (k == 3) AT_TRACE_ISS(2);
                        }
                        AT_TRACE_ISS(3);
                        if (elect_one()) umma_commit(&pv_done[g * PB + pb]);
                        AT_TRACE_ISS(4);
                    }
                    if (elect_one()) umma_commit(&v_empty[sv]);    // one arrival per issuer
                    AT_TRACE_ISS(5);
                    if (g_lo == 0) AT_TRACE_MMA(7);
                    if (++sv == ST) { sv = 0; phv ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================= softmax / correction / epilogue =================
        const int g = (warp - 4) >> 2;                  // query tile of this warp
        const int lg = warp & 3;                        // TMEM lane quarter
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
        const uint32_t s_addr = lane_addr + Cfg::s_col(g);
        const uint32_t o_addr = lane_addr + Cfg::o_col(g);
        const uint32_t prow0 = smem_u32(sP + g * PB * Cfg::P_BYTES) + row * 128;
        float m_used = -INFINITY, l = 0.f;
        constexpr bool ones = ONES && SDB_ATTN_ONES;         // the row sum comes out of the PV MMA (channel d of V is 1)
        const float c = p.scale_log2;
        // Exponential phase ping-pong (G == 2): the two softmax warps of one SM sub-partition share its MUFU unit and its
        // TMEM read port.  Left alone they run in lockstep — both reading S, then both in ex2 at half rate — and the MUFU
        // pipe idles ~40 % of the time.  A pair of named barriers per lane quarter hands the ex2 phase back and forth, so
        // one warp's TMEM read + row max always runs under the other's exponentials.
        const uint32_t bar_mine = 2 + 4 * g + lg, bar_other = 2 + 4 * (g ^ 1) + lg;
        if (G == 2 && g == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");   // tile 0 goes first
        for (int t = 0; t < ntiles; ++t) {
            const int kvalid = p.Sk - t * 128;          // keys >= kvalid are padding
            AT_TRACE(0);
            mbar_wait(&s_full[g], t & 1);
            tcgen05_fence_after();
            AT_TRACE(1);
            // the whole score row -> registers, then S_g is free for the next Q K^T
            uint32_t r[128];
            tmem_ld_x32(s_addr, r);
            tmem_ld_x32(s_addr + 32, r + 32);
            tmem_ld_x32(s_addr + 64, r + 64);
            tmem_ld_x32(s_addr + 96, r + 96);
            tmem_ld_wait();
            AT_TRACE(2);
            tcgen05_fence_before();
            mbar_arrive(&s_empty[g]);
            if (kvalid < 128) {
#pragma unroll
                for (int j = 0; j < 128; ++j) if (j >= kvalid) r[j] = 0xff800000u;   // -inf
            }
            if (p.causal) {
                // keys after this thread's query are masked; key 0 is always visible, so no row is ever empty
                const int jmax = q0 + g * 128 + row - t * 128;                       // last visible key of this tile
#pragma unroll
                for (int j = 0; j < 128; ++j) if (j > jmax) r[j] = 0xff800000u;
            }
#if SDB_ATTN_MAXCHAINS == 4
            // four independent 3-input max chains of 16 (a chain link costs the ALU latency, not an issue slot)
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 128; j += 8) {
                mx0 = fmax3(mx0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                mx1 = fmax3(mx1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                mx2 = fmax3(mx2, __uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
                mx3 = fmax3(mx3, __uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
#else
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 128; j += 4) {
                mx0 = fmax3(mx0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                mx1 = fmax3(mx1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
            const float mx = fmaxf(mx0, mx1) * c;
#endif
            const bool need = mx > m_used + 8.0f;
            float alpha = 1.0f;
            if (need) { alpha = ex2_approx(m_used - mx); m_used = mx; }
            const bool any_need = __any_sync(0xffffffffu, need);
            AT_TRACE(3);
            const int pb = PB == 2 ? (t & 1) : 0;
            const uint32_t prow = prow0 + pb * Cfg::P_BYTES;
            if (PB == 2) {
                // this P buffer was last read by P_{t-2} V: almost always retired long ago
                if (t >= 2) mbar_wait(&pv_done[g * PB + pb], ((t >> 1) - 1) & 1);
            }
            // PB == 1: O_g and the single P_g buffer are free once PV_g(t-1) has retired;
            // PB == 2: only a rescale of O_g (rare) has to wait for P_{t-1} V
            if (t > 0 && (PB == 1 || any_need)) {
                if (PB == 2) mbar_wait(&pv_done[g * PB + ((t - 1) & 1)], ((t - 1) >> 1) & 1);
                else mbar_wait(&pv_done[g], (t - 1) & 1);
                tcgen05_fence_after();
                if (any_need) {
#pragma unroll 1
                    for (int c0 = 0; c0 < DPAD; c0 += 16) {
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
            AT_TRACE(4);
            if (G == 2) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");              // my turn on the MUFU
            AT_TRACE(5);
            // p = exp2(s*c - m), row sum, P_g (bf16, K-major SW128: two [128 x 64] blocks)
#if SDB_ATTN_PACKED
            // Packed fp32x2 arithmetic (FFMA2 / FADD2): s*c - m for two keys per instruction and pairwise row-sum accumulation,
            // 2.6 instead of 3.6 issue slots per exponential.  A softmax warp that is alone in its exponential phase is
            // issue- / dependency-bound (~12 clk per exponential against the MUFU pipe's 8, tools/ubench/xu_rate.cu).
            uint64_t acc_a = 0ull, acc_b = 0ull;                 // two (lo, hi) fp32 pair accumulators
            const uint64_t c2 = pack_f32x2(c, c), nm2 = pack_f32x2(-m_used, -m_used);
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 8) {
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(r[c0 + j]), __uint_as_float(r[c0 + j + 1])), c2, nm2);
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    e[j] = ex2_approx(x0);
                    e[j + 1] = ex2_approx(x1);
                }
                if (!ones) {
                    acc_a = fadd2(acc_a, pack_f32x2(e[0], e[1]));
                    acc_b = fadd2(acc_b, pack_f32x2(e[2], e[3]));
                    acc_a = fadd2(acc_a, pack_f32x2(e[4], e[5]));
                    acc_b = fadd2(acc_b, pack_f32x2(e[6], e[7]));
                }
                const int chunk = (c0 & 63) >> 3;
                sts128(prow + (c0 >> 6) * 16384 + ((chunk ^ (row & 7)) << 4),
                       pack_p_bf16x2(e[0], e[1]), pack_p_bf16x2(e[2], e[3]), pack_p_bf16x2(e[4], e[5]), pack_p_bf16x2(e[6], e[7]));
                if (G == 2 && SDB_ATTN_HANDOVER < 16 && c0 == 8 * (SDB_ATTN_HANDOVER - 1))
                    asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
            }
            float sum0, sum1;
            {
                float a0, a1, b0, b1;
                unpack_f32x2(acc_a, a0, a1);
                unpack_f32x2(acc_b, b0, b1);
                sum0 = a0 + a1;
                sum1 = b0 + b1;
            }
#else
            float sum0 = 0.f, sum1 = 0.f;
            const float nm = -m_used;
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 8) {
                float e[8];
#define SDB_E(j) e[j] = ex2_sel<j>(fmaf(__uint_as_float(r[c0 + j]), c, nm))
                SDB_E(0); SDB_E(1); SDB_E(2); SDB_E(3); SDB_E(4); SDB_E(5); SDB_E(6); SDB_E(7);
#undef SDB_E
                sum0 += (e[0] + e[1]) + (e[2] + e[3]);
                sum1 += (e[4] + e[5]) + (e[6] + e[7]);
                const int chunk = (c0 & 63) >> 3;
                sts128(prow + (c0 >> 6) * 16384 + ((chunk ^ (row & 7)) << 4),
                       pack_p_bf16x2(e[0], e[1]), pack_p_bf16x2(e[2], e[3]), pack_p_bf16x2(e[4], e[5]), pack_p_bf16x2(e[6], e[7]));
                // early hand-over: the other warp's exponentials start under the tail of this warp's (see SDB_ATTN_HANDOVER)
                if (G == 2 && SDB_ATTN_HANDOVER < 16 && c0 == 8 * (SDB_ATTN_HANDOVER - 1))
                    asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
            }
#endif
            l += sum0 + sum1;
            AT_TRACE(6);
            if (G == 2 && SDB_ATTN_HANDOVER >= 16) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");   // hand the MUFU over
            tcgen05_fence_before();
            fence_proxy_async_smem();               // P visible to the tensor core (async proxy)
            mbar_arrive(&p_full[g * PB + pb]);
            AT_TRACE(7);
        }
        // ---- epilogue: O / l -> bf16 -> out[b, q, h, 0..d) ----
        if (PB == 2) mbar_wait(&pv_done[g * PB + ((ntiles - 1) & 1)], ((ntiles - 1) >> 1) & 1);   // in-order pipe: the last PV retires last
        else mbar_wait(&pv_done[g], (ntiles - 1) & 1);
        tcgen05_fence_after();
        if (ones) {
            l = __uint_as_float(tmem_ld_x1(o_addr + (uint32_t)p.d));      // sum_j P_ij, accumulated (and rescaled) with O
            tmem_ld_wait();
        }
        const float inv = 1.0f / l;
        const int q = q0 + g * 128 + row;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)b * p.o_bs + (long long)q * p.o_ss + (long long)h * p.o_hs;
#pragma unroll 1
        for (int c0 = 0; c0 < DPAD; c0 += 16) {
            uint32_t o[16];
            tmem_ld_x16(o_addr + c0, o);
            tmem_ld_wait();
            if (q < p.Sq && c0 < p.d) {
#pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    if (c0 + 8 * gg < p.d) {
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
    }

    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, Cfg::TMEM_COLS);
    }
}

template <int DPAD, bool ONES>
static int launch_attn(const sdb_attn_args* a, cudaStream_t st) {
    using Cfg = AttnCfg<DPAD>;
    CUtensorMap tmQ, tmK, tmV;
    int box[4] = {64, 1, 128, 1};
    int es[4] = {1, 1, 1, 1};
    {
        // dense heads: the map is only d channels wide, the rest of each 64-channel box is out of bounds = zero-filled
        long long dims[4] = {a->dense ? a->d : DPAD, a->H, a->Sq, a->B};
        long long str[3] = {a->q_hs, a->q_ss, a->q_bs};
        int rc = make_tmap_bf16(&tmQ, a->q, 4, dims, str, box, es);
        if (rc) return rc;
    }
    {
        long long dims[4] = {a->dense ? a->d : DPAD, a->H, a->Sk, a->B};
        long long str[3] = {a->k_hs, a->k_ss, a->k_bs};
        int rc = make_tmap_bf16(&tmK, a->k, 4, dims, str, box, es);
        if (rc) return rc;
        long long strv[3] = {a->v_hs, a->v_ss, a->v_bs};
        rc = make_tmap_bf16(&tmV, a->v, 4, dims, strv, box, es);
        if (rc) return rc;
    }
    static bool attr_set_dev[64] = {false};          // the attribute is per device
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    bool& attr_set = attr_set_dev[cur_dev];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_attention_kernel<DPAD, ONES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    AttnP p;
#ifdef SDB_ATTN_TRACE
    p.trace = g_attn_trace;
#endif
    p.out = a->out; p.o_bs = a->o_bs; p.o_ss = a->o_ss; p.o_hs = a->o_hs;
    p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d; p.H = a->H;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.causal = a->causal ? 1 : 0;
    dim3 grid((unsigned)ceil_div(a->Sq, 128 * Cfg::G), (unsigned)a->H, (unsigned)a->B);
    launch_pdl(tc_attention_kernel<DPAD, ONES>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmQ, tmK, tmV, p);
    return check_launch("tc_attention_kernel");
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_attention_fwd(const sdb_attn_args* a, void* stream) {
    SDB_REQUIRE(a && a->q && a->k && a->v && a->out, "attention: null pointer");
    SDB_REQUIRE(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Sk > 0 && a->d > 0, "attention: empty problem");
    SDB_REQUIRE(a->d % 8 == 0 && a->dpad % 64 == 0 && a->dpad >= a->d && a->dpad <= 192, "attention: d=%d dpad=%d unsupported", a->d, a->dpad);
    SDB_REQUIRE(a->B <= 65535 && a->H <= 65535, "attention: grid too large");
    SDB_REQUIRE(!a->causal || a->Sq == a->Sk, "attention: the causal mask needs Sq == Sk (got %d, %d)", a->Sq, a->Sk);
    SDB_REQUIRE(a->o_hs % 8 == 0 && a->o_ss % 8 == 0 && a->o_bs % 8 == 0 && ((uintptr_t)a->out & 15) == 0, "attention: output must be 16-byte aligned per head row");
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int rc = attention_kv1_dispatch(a, st);     // Sk <= 128 (the text conditioning): 1 = not covered
        if (rc != 1) return rc;
    }
    // a pad channel exists and the kernel is exponential-bound (64-channel heads): row sums from the PV MMA.  Wider heads are
    // MMA-paced and lose 0.5-2 % to the V-tile patch in front of every PV issue (profiles/r02_attn_notes.txt)
    const bool spare = a->d < a->dpad && a->dpad == 64;
    switch (a->dpad) {
        case 64: return spare ? launch_attn<64, true>(a, st) : launch_attn<64, false>(a, st);
        case 128: return spare ? launch_attn<128, true>(a, st) : launch_attn<128, false>(a, st);
        default: return spare ? launch_attn<192, true>(a, st) : launch_attn<192, false>(a, st);
    }
}
