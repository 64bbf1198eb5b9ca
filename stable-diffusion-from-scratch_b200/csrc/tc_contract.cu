// sdb200 — tcgen05 contraction kernel: GEMM and implicit-GEMM convolution on the 5th-gen tensor
// cores (bf16 operands, fp32 accumulation in TMEM), operands staged by TMA into a 128B-swizzled
// shared-memory ring.
//
// Replaces (bf16 mode): nn.Conv2d 3x3 s1/s2 and 1x1 (openai_model/model.py:88-90,117,181,207,218,
// 365,531; openai_model/attention.py:319-334; ldm/modules/diffusionmodules/model.py:49,94,104,
// 158-177,493,535), nn.Linear (openai_model/attention.py:40-47,133,159-167) with the GEGLU product
// (attention.py:140-141) fused into the epilogue, and the ResBlock epilogue adds
// (h + emb_out[..., None, None], model.py:241-250; skip_connection(x) + h, model.py:252).
//
// Structure (one CTA = one 128 x BN output tile, 192 threads):
//   warp 0      TMA producer   : A tile [128 rows x 64 k] + B tile [BN rows x 64 k] per k-block
//   warp 1      MMA issuer     : 4 x tcgen05.mma (M=128, N=BN, K=16) per k-block, commit -> empty
//   warps 2..5  epilogue       : tcgen05.ld 32x32b (thread == output row) -> bias / time-emb /
//                                residual / GEGLU -> global store (or red.add for split-K)
// Implicit conv: the A tile for tap (r,s) of a tw x th x tn block of output pixels is one 4-D TMA
// box of the NHWC activation starting at (c0, ow0*stride+s-pad, oh0*stride+r-pad, n0); TMA
// zero-fills out-of-bounds (= the conv padding), traversal strides give stride-2.
// Two CTAs are resident per SM (3-stage rings each), so one CTA's epilogue overlaps the other's
// mainloop.  Roofline: tensor pipe; algorithmic FLOP = 2*M*N*K*taps.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace sdb {

using namespace ptx;

struct TcP {
    void* out;
    const float* bias; const float* rowvec; const float* residual;
    float* ws; long long ws_split_stride;   // split-K partial sums: [split_k][M_out][N] fp32
    float* colstats; long long colstats_sq; // per-(32-row slot, column) sum / sum of squares of the stored values: [2][slots][N]
    long long ldc, ldr, ldv;
    int M, N;
    int kblocks;            // total k-blocks = taps * kpt
    int kpt;                // k-blocks per tap
    int out_bf16, geglu;
    int col_group, col_group_stride;
    int split_k;
    int tiles_n;
    int m_pairs;            // pair kernel: number of 256-row tile pairs
    int off32;              // every element offset into out / residual / workspace fits 32 bits (the fast epilogue's addressing)
    int stages;             // pair kernel: depth of the operand ring (what is left of shared memory after the epilogue staging)
    int epi_tma;            // pair kernel: 1 = the epilogue stages 32 x 32 boxes in swizzled smem and moves them with TMA
    int epi_nbuf;           //   staging boxes per epilogue warp (2; 3 when a residual tile is prefetched two chunks ahead)
    int epi_bw, epi_bh;     //   conv: the 32 rows of a lane quarter are the pixel box bw x bh x 32/(bw*bh) of the tile
    int b_const;            // B (weights) is not written by the preceding launch: its first tiles are fetched before pdl_wait()
    int full_tiles;         // one-CTA kernel: tiles [0, full_tiles) are whole 128 x BN tiles; every later tile is computed by TWO CTAs as a
                            // 96-column and a (BN - 96)-column sub-tile (tail split against wave quantisation, see launch_tc)
    // conv
    int conv;
    int kw, stride, pad_h, pad_w;
    int NB, OH, OW;
    int tw, th, tn, tiles_w, tiles_h;
    int cout_pad;
    int out_sh, out_sw, out_oh, out_ow, OHF, OWF;
#ifdef SDB_TC_TRACE
    long long* trace;       // [pair][64 units][8] SM clock stamps (measurement build only, tools/trace_pair.py)
#endif
};

#ifdef SDB_TC_TRACE
#define TC_TRACE(slot, it_)                                                                              \
    do {                                                                                                 \
        if (p.trace && rank == 0 && (it_) < 64) p.trace[((long long)pair * 64 + (it_)) * 8 + (slot)] = clock64(); \
    } while (0)
static long long* g_trace_ptr = nullptr;
extern "C" void sdb_tc_set_trace(long long* ptr) { g_trace_ptr = ptr; }
#define TC1_TRACE(slot)                                                                                  \
    do {                                                                                                 \
        if (p.trace && blockIdx.x < 2048 && blockIdx.y == 0) {                                           \
            if ((slot) == 7) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.trace[(long long)blockIdx.x * 8 + 7] = sm; } \
            else p.trace[(long long)blockIdx.x * 8 + (slot)] = (long long)globaltimer_ns();               \
        }                                                                                                \
    } while (0)
#else
#define TC_TRACE(slot, it_) do { } while (0)
#define TC1_TRACE(slot) do { } while (0)
#endif

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB

template <int BN>
struct TcCfg {
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 2 : (BN >= 128 ? 3 : (BN >= 64 ? 4 : 5));
    static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256 + BN * 4;
};

constexpr int EPI_LD = 36;                        // floats per staged row: 16-B aligned, conflict-free for 128-bit access
constexpr int TEPI_BOX_BYTES = 4096;              // TMA epilogue: one staged box = 32 rows x 128 B (fp32) or 32 rows x 64 B (bf16, half used)
constexpr int EPI_WARP_BYTES = 32 * EPI_LD * 4;   // one 32 x 32 fp32 chunk per epilogue warp

__device__ __forceinline__ float4 ldg_f4_or_zero(const float* p, bool pred) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pred) r = __ldg(reinterpret_cast<const float4*>(p));
    return r;
}

// L2 prefetch of one output row's residual segment [nb, nb + cols): issued long before the epilogue reads it (at kernel
// start in the one-CTA kernel, one work unit ahead in the persistent pair kernel), so that the epilogue's residual
// loads hit L2 (~700 cycles) instead of DRAM (~2500 under load) — with one 32-column chunk of loads in flight per warp
// the epilogue is latency-bound otherwise.
__device__ __forceinline__ void prefetch_residual_row(const TcP& p, long long pix, int nb, int cols) {
    if (p.residual == nullptr || p.split_k > 1 || p.geglu || pix < 0) return;
    int w = p.N - nb;
    if (w > cols) w = cols;
    if (w <= 0) return;
    const float* a = p.residual + pix * p.ldr + nb;
    const uint32_t bytes = (uint32_t)w * 4u;
    if ((reinterpret_cast<uintptr_t>(a) & 15) != 0 || (bytes & 15) != 0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

// Epilogue of one warp = 32 rows (its TMEM lane quarter) x (a subset of) the BN columns of an accumulator.
// tcgen05.ld gives each lane one ROW (32 consecutive columns per chunk); stores in that layout would
// touch 32 different 128-B lines per instruction.  The chunk is therefore transposed through a padded
// smem tile so that 8 lanes cover 32 consecutive columns (128 B) of one row and one warp instruction
// covers 4 whole rows: bias / time-emb row / residual loads and the output stores are all coalesced,
// and the residual / time-emb loads of a chunk are issued before its TMEM load so their latency overlaps.
// `row_pix` = output row index of THIS lane's tile row (or -1 when the row is outside the problem),
// `row_img` = its image index (conv mode; selects the time-embedding row).  The warp handles the
// 32-column chunks ch0, ch0 + chstep, ... (two warps can share one lane quarter).
template <int BN>
__device__ __noinline__ void epilogue_generic(const TcP p /* by value: a reference would force the kernel's parameter block into local memory */, uint32_t taddr, const float* s_bias, float* stage, int lane,
                                              int n0, int nt, long long row_pix, int row_img, int split, int ch0, int chstep,
                                              uint64_t* acc_ready, uint32_t acc_parity, int slot, int ncols) {
    // rows this lane stores after the transpose: r_i = 4*i + (lane >> 3), i = 0..7
    long long rpix[8];
    int rimg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int src = 4 * i + (lane >> 3);
        rpix[i] = __shfl_sync(0xffffffffu, row_pix, src);
        rimg[i] = __shfl_sync(0xffffffffu, row_img, src);
    }
    const int cq = 4 * (lane & 7);                // this lane's 4 columns inside a 32-column chunk
    const uint32_t stage_a = smem_u32(stage), sbias_a = smem_u32(s_bias);
    const bool partial = p.split_k > 1;           // write raw partial sums to the split-K workspace
    float* const ws = partial ? p.ws + (long long)split * p.ws_split_stride : nullptr;
    const float* const resid = (p.residual && !partial) ? p.residual : nullptr;
    const float* const rowv = (p.rowvec && !partial && p.conv) ? p.rowvec : nullptr;
    constexpr int CH = 32;
    constexpr int NCHUNK = (BN + CH - 1) / CH;
    const int n_out = p.geglu ? p.N / 2 : p.N;    // stored columns
    const int nb = p.geglu ? nt * (BN / 2) : n0;  // first stored column of this tile
    const int cols = p.geglu ? BN / 2 : BN;
    // vector fast path: whole 4-column groups, every row pitch and base 16-byte aligned (warp-uniform test)
    const bool vec_ok = (n_out % 4 == 0) && (p.ldc % 4 == 0) && (p.ldr % 4 == 0) && (p.ldv % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.residual) |
                          reinterpret_cast<uintptr_t>(p.rowvec)) & 15) == 0 &&
                        (!p.col_group || (p.col_group % 4 == 0 && p.col_group_stride % 4 == 0));
    // residual / time-emb addends of a chunk, fetched one chunk ahead (the first one before the accumulator is even
    // complete) so that their DRAM latency overlaps the MMA tail, the TMEM read and the previous chunk's stores
    float4 addn[8];
    auto fetch_addends = [&](int ch, float4 (&dst)[8]) {
        const int cn_ = nb + ch * CH + cq;
        const bool live = vec_ok && ch < NCHUNK && ch * CH < cols && cn_ < n_out;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = live && rpix[i] >= 0;
            dst[i] = ldg_f4_or_zero(resid + rpix[i] * p.ldr + cn_, ok && resid != nullptr);
            if (rowv) {
                float4 q = ldg_f4_or_zero(rowv + (long long)rimg[i] * p.ldv + cn_, ok);
                dst[i].x += q.x; dst[i].y += q.y; dst[i].z += q.z; dst[i].w += q.w;
            }
        }
    };
    fetch_addends(ch0, addn);
    mbar_wait(acc_ready, acc_parity);
    tcgen05_fence_after();
#pragma unroll 1
    for (int ch = ch0; ch < NCHUNK; ch += chstep) {
        const int c0 = ch * CH;
        if (c0 >= cols || c0 >= ncols || nb + c0 >= n_out) break;      // warp-uniform
        const int cn = nb + c0 + cq;                                    // first of this lane's 4 output columns
        const bool col_ok = cn < n_out;
        float4 addv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) addv[i] = addn[i];
        fetch_addends(ch + chstep, addn);
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        if (p.geglu) {
            // value columns [0, BN/2) | gate columns [BN/2, BN) of the same outputs (weights packed so by the host)
            uint32_t g[32];
            tmem_ld_x32(taddr + BN / 2 + c0, g);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float a = __uint_as_float(r[j]) + lds32(sbias_a + 4 * (c0 + j));
                float gg = __uint_as_float(g[j]) + lds32(sbias_a + 4 * (BN / 2 + c0 + j));
                r[j] = __float_as_uint(a * gelu_erf_fast(gg));
            }
        } else {
            tmem_ld_wait();
        }
        __syncwarp();                                                   // previous chunk's readers are done
#pragma unroll
        for (int q = 0; q < 8; ++q)
            sts128(stage_a + 4 * (lane * EPI_LD + 4 * q), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.geglu && !partial) bias4 = lds128(sbias_a + 4 * (c0 + cq));
        if (vec_ok) {
            // column statistics of the stored tile (GroupNorm of the consumer, see sdb_tc_args.colstats)
            const bool want_cs = p.colstats != nullptr && !partial;
            float4 cs_s = make_float4(0.f, 0.f, 0.f, 0.f), cs_q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long pix = rpix[i];
                float4 v = lds128(stage_a + 4 * ((4 * i + (lane >> 3)) * EPI_LD + cq));
                v.x += bias4.x + addv[i].x; v.y += bias4.y + addv[i].y; v.z += bias4.z + addv[i].z; v.w += bias4.w + addv[i].w;
                if (!(col_ok && pix >= 0)) continue;
                if (want_cs) {
                    cs_s.x += v.x; cs_s.y += v.y; cs_s.z += v.z; cs_s.w += v.w;
                    cs_q.x = fmaf(v.x, v.x, cs_q.x); cs_q.y = fmaf(v.y, v.y, cs_q.y);
                    cs_q.z = fmaf(v.z, v.z, cs_q.z); cs_q.w = fmaf(v.w, v.w, cs_q.w);
                }
                if (partial) {
                    *reinterpret_cast<float4*>(ws + pix * p.N + cn) = v;                      // dense [rows_out, N]
                } else if (p.out_bf16) {
                    const int dn = p.col_group ? (cn / p.col_group) * p.col_group_stride + cn % p.col_group : cn;
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldc + dn) =
                        make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                } else {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.ldc + cn) = v;
                }
            }
            if (want_cs) {
                // rows 4i + (lane >> 3) live in this lane: fold the 4 lanes that share a column quad (fixed order)
#pragma unroll
                for (int o = 8; o <= 16; o <<= 1) {
                    cs_s.x += __shfl_xor_sync(0xffffffffu, cs_s.x, o); cs_s.y += __shfl_xor_sync(0xffffffffu, cs_s.y, o);
                    cs_s.z += __shfl_xor_sync(0xffffffffu, cs_s.z, o); cs_s.w += __shfl_xor_sync(0xffffffffu, cs_s.w, o);
                    cs_q.x += __shfl_xor_sync(0xffffffffu, cs_q.x, o); cs_q.y += __shfl_xor_sync(0xffffffffu, cs_q.y, o);
                    cs_q.z += __shfl_xor_sync(0xffffffffu, cs_q.z, o); cs_q.w += __shfl_xor_sync(0xffffffffu, cs_q.w, o);
                }
                if (lane < 8 && col_ok) {
                    float* dst = p.colstats + (long long)slot * p.N + cn;
                    *reinterpret_cast<float4*>(dst) = cs_s;
                    *reinterpret_cast<float4*>(dst + p.colstats_sq) = cs_q;
                }
            }
        } else {
            // generic path: any N / pitch / alignment, element by element
            const int nvalid = n_out - cn < 4 ? n_out - cn : 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long pix = rpix[i];
                if (pix < 0 || nvalid <= 0) continue;
                const float4 v4 = lds128(stage_a + 4 * ((4 * i + (lane >> 3)) * EPI_LD + cq));
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (t >= nvalid) break;
                    const int n2 = cn + t;
                    float v = t == 0 ? v4.x + bias4.x : (t == 1 ? v4.y + bias4.y : (t == 2 ? v4.z + bias4.z : v4.w + bias4.w));
                    if (rowv) v += __ldg(rowv + (long long)rimg[i] * p.ldv + n2);
                    if (resid) v += __ldg(resid + pix * p.ldr + n2);
                    if (partial) {
                        ws[pix * p.N + n2] = v;
                    } else {
                        const int d2 = p.col_group ? (n2 / p.col_group) * p.col_group_stride + n2 % p.col_group : n2;
                        if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[pix * p.ldc + d2] = __float2bfloat16_rn(v);
                        else reinterpret_cast<float*>(p.out)[pix * p.ldc + d2] = v;
                    }
                }
            }
        }
    }
    __syncwarp();
}

// ---- fast epilogue -------------------------------------------------------------------------------------
// Same data flow as epilogue_generic (TMEM row-per-lane -> padded smem transpose -> 4 full 128-B row segments per
// store instruction) for the layouts the models actually use (16-byte aligned, N % 4 == 0, residual pitch == output
// pitch), with everything that does not change inside a work unit hoisted out of the chunk loop: the eight row
// offsets and their validity mask, the column-group remap, the uniform time-embedding row; stores are predicated,
// not branched around, and the output kind is a template parameter.  The generic version spent ~900 warp
// instructions per 32x32 chunk on index arithmetic and divergence bookkeeping (profiles/r01_pair_timeline.txt: the
// epilogue, not the MMA, paced every K <= 1280 layer); this one needs ~200.
enum { EPI_F32 = 0, EPI_BF16 = 1, EPI_PARTIAL = 2, EPI_GEGLU = 3 };

template <int BN, int MODE, bool HAS_ADD>
__device__ __forceinline__ void epilogue_fast(const TcP& p, uint32_t taddr, const float* s_bias, float* stage, int lane,
                                              int n0, int nt, long long row_pix, int row_img, int split, int ch0, int chstep,
                                              uint64_t* acc_ready, uint32_t acc_parity, int slot, int ncols) {
    constexpr int CH = 32;
    constexpr bool GEGLU = MODE == EPI_GEGLU;
    constexpr int COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCHUNK = (COLS + CH - 1) / CH;
    const int n_out = GEGLU ? p.N / 2 : p.N;
    const int nb = GEGLU ? nt * (BN / 2) : n0;
    const int cq = 4 * (lane & 7);
    const int rsel = lane >> 3;
    const uint32_t stage_a = smem_u32(stage), sbias_a = smem_u32(s_bias);
    const long long ldo = MODE == EPI_PARTIAL ? (long long)p.N : p.ldc;
    // rows this lane stores after the transpose: 4*i + rsel.  32-bit element offsets (the launcher checked that they fit):
    // with 64-bit ones ptxas kept the pixel indices and re-derived px * ld (two 64-bit IMADs + carries) and px >= 0 in front
    // of every load and store, ~1/4 of the epilogue's instructions (profiles/r01_ncu_epilogue.txt)
    uint32_t roff[8];
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long px = __shfl_sync(0xffffffffu, row_pix, 4 * i + rsel);
        if (px >= 0) mask |= 1u << i;
        roff[i] = (uint32_t)px * (uint32_t)ldo;
    }
    const float* const resid = (HAS_ADD && p.residual) ? p.residual : nullptr;
    const float* rowv = (HAS_ADD && p.rowvec && p.conv) ? p.rowvec : nullptr;
    const int img0 = __shfl_sync(0xffffffffu, row_img, 0);
    const bool rv_uniform = __all_sync(0xffffffffu, row_img == img0);
    int rimg[8];
    if (HAS_ADD && rowv && !rv_uniform) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rimg[i] = __shfl_sync(0xffffffffu, row_img, 4 * i + rsel);
    }
    float* const ws = MODE == EPI_PARTIAL ? p.ws + (long long)split * p.ws_split_stride : nullptr;
    const bool want_cs = MODE == EPI_F32 && p.colstats != nullptr;

    float4 addn[8];
    auto fetch_residual = [&](int ch, float4 (&dst)[8]) {
        const int cn_ = nb + ch * CH + cq;
        const bool live = resid != nullptr && ch < NCHUNK && cn_ < n_out;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = ldg_f4_or_zero(resid + (roff[i] + (uint32_t)cn_), live && ((mask >> i) & 1u));
    };
    if (HAS_ADD) fetch_residual(ch0, addn);
    mbar_wait(acc_ready, acc_parity);
    tcgen05_fence_after();
#pragma unroll 1
    for (int ch = ch0; ch < NCHUNK; ch += chstep) {
        const int c0 = ch * CH;
        if (c0 >= ncols || nb + c0 >= n_out) break;                     // warp-uniform
        const int cn = nb + c0 + cq;
        const bool col_ok = cn < n_out;
        float4 addv[8];
        if (HAS_ADD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) addv[i] = addn[i];
            fetch_residual(ch + chstep, addn);
        }
        float4 cadd = make_float4(0.f, 0.f, 0.f, 0.f);                  // per-column addend: bias
        // time-embedding row: one load per chunk when the whole tile belongs to one image, else one per row; it is added
        // LAST in both cases, so a sample's bits do not depend on whether its tile is shared with another sample
        float4 rv4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // (mask == 0: every row of this warp lies outside the problem — the phantom m-tile of an odd pair, or a tile past
        // the last image — and img0 may be >= NB: nothing is stored, so nothing may be loaded either)
        if (HAS_ADD && rowv && rv_uniform) rv4 = ldg_f4_or_zero(rowv + (long long)img0 * p.ldv + cn, col_ok && mask != 0u && img0 < p.NB);
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        if (GEGLU) {
            uint32_t g[32];
            tmem_ld_x32(taddr + BN / 2 + c0, g);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 bv = lds128(sbias_a + 4 * (c0 + 4 * q));
                const float4 bg = lds128(sbias_a + 4 * (BN / 2 + c0 + 4 * q));
                r[4 * q + 0] = __float_as_uint((__uint_as_float(r[4 * q + 0]) + bv.x) * gelu_erf_fast(__uint_as_float(g[4 * q + 0]) + bg.x));
                r[4 * q + 1] = __float_as_uint((__uint_as_float(r[4 * q + 1]) + bv.y) * gelu_erf_fast(__uint_as_float(g[4 * q + 1]) + bg.y));
                r[4 * q + 2] = __float_as_uint((__uint_as_float(r[4 * q + 2]) + bv.z) * gelu_erf_fast(__uint_as_float(g[4 * q + 2]) + bg.z));
                r[4 * q + 3] = __float_as_uint((__uint_as_float(r[4 * q + 3]) + bv.w) * gelu_erf_fast(__uint_as_float(g[4 * q + 3]) + bg.w));
            }
        } else {
            tmem_ld_wait();
        }
        __syncwarp();                                                   // previous chunk's readers are done
#pragma unroll
        for (int q = 0; q < 8; ++q)
            sts128(stage_a + 4 * (lane * EPI_LD + 4 * q), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        if (MODE == EPI_F32 || MODE == EPI_BF16) {
            const float4 b4 = lds128(sbias_a + 4 * (c0 + cq));
            cadd.x += b4.x; cadd.y += b4.y; cadd.z += b4.z; cadd.w += b4.w;
        }
        int dn = cn;                                                    // destination column (q/k/v head padding remap)
        if (MODE == EPI_BF16 && p.col_group) dn = (cn / p.col_group) * p.col_group_stride + cn % p.col_group;
        float4 cs_s = make_float4(0.f, 0.f, 0.f, 0.f), cs_q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = lds128(stage_a + 4 * ((4 * i + rsel) * EPI_LD + cq));
            if (MODE == EPI_F32 || MODE == EPI_BF16) { v.x += cadd.x; v.y += cadd.y; v.z += cadd.z; v.w += cadd.w; }
            if (HAS_ADD) {
                v.x += addv[i].x; v.y += addv[i].y; v.z += addv[i].z; v.w += addv[i].w;
                float4 q4 = rv4;
                if (rowv && !rv_uniform) q4 = ldg_f4_or_zero(rowv + (long long)rimg[i] * p.ldv + cn, col_ok && ((mask >> i) & 1u));
                v.x += q4.x; v.y += q4.y; v.z += q4.z; v.w += q4.w;
            }
            const bool ok = col_ok && ((mask >> i) & 1u);
            if (MODE == EPI_F32) {
                if (want_cs && ok) {
                    cs_s.x += v.x; cs_s.y += v.y; cs_s.z += v.z; cs_s.w += v.w;
                    cs_q.x = fmaf(v.x, v.x, cs_q.x); cs_q.y = fmaf(v.y, v.y, cs_q.y);
                    cs_q.z = fmaf(v.z, v.z, cs_q.z); cs_q.w = fmaf(v.w, v.w, cs_q.w);
                }
                if (ok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (roff[i] + (uint32_t)cn)) = v;
            } else if (MODE == EPI_PARTIAL) {
                if (ok) *reinterpret_cast<float4*>(ws + (roff[i] + (uint32_t)cn)) = v;
            } else {      // EPI_BF16, EPI_GEGLU (the fast path takes GEGLU only with a bf16 output, see epilogue_warp)
                if (ok) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (roff[i] + (uint32_t)dn)) =
                            make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
            }
        }
        if (want_cs) {
            // rows 4i + rsel live in this lane: fold the 4 lanes that share a column quad (fixed order)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
                cs_s.x += __shfl_xor_sync(0xffffffffu, cs_s.x, o); cs_s.y += __shfl_xor_sync(0xffffffffu, cs_s.y, o);
                cs_s.z += __shfl_xor_sync(0xffffffffu, cs_s.z, o); cs_s.w += __shfl_xor_sync(0xffffffffu, cs_s.w, o);
                cs_q.x += __shfl_xor_sync(0xffffffffu, cs_q.x, o); cs_q.y += __shfl_xor_sync(0xffffffffu, cs_q.y, o);
                cs_q.z += __shfl_xor_sync(0xffffffffu, cs_q.z, o); cs_q.w += __shfl_xor_sync(0xffffffffu, cs_q.w, o);
            }
            if (lane < 8 && col_ok) {
                float* dst = p.colstats + (long long)slot * p.N + cn;
                *reinterpret_cast<float4*>(dst) = cs_s;
                *reinterpret_cast<float4*>(dst + p.colstats_sq) = cs_q;
            }
        }
    }
    __syncwarp();
}

// picks the epilogue for this launch (warp-uniform): the fast one whenever the layout allows
template <int BN>
__device__ __forceinline__ void epilogue_warp(const TcP& p, uint32_t taddr, const float* s_bias, float* stage, int lane,
                                              int n0, int nt, long long row_pix, int row_img, int split, int ch0, int chstep,
                                              uint64_t* acc_ready, uint32_t acc_parity, int slot, int ncols = 1 << 30) {
    const int n_out = p.geglu ? p.N / 2 : p.N;
    const bool partial = p.split_k > 1;
    const bool fast_ok = p.off32 && (n_out % 4 == 0) && (p.ldc % 4 == 0) && (p.ldv % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.residual) |
                           reinterpret_cast<uintptr_t>(p.rowvec) | reinterpret_cast<uintptr_t>(p.ws)) & 15) == 0 &&
                         (!p.col_group || (p.col_group % 4 == 0 && p.col_group_stride % 4 == 0)) &&
                         (p.residual == nullptr || partial || p.ldr == p.ldc) && !(p.geglu && (partial || !p.out_bf16));
#define SDB_EPI(MODE, ADD) epilogue_fast<BN, MODE, ADD>(p, taddr, s_bias, stage, lane, n0, nt, row_pix, row_img, split, ch0, chstep, acc_ready, acc_parity, slot, ncols)
    if (!fast_ok) {
        epilogue_generic<BN>(p, taddr, s_bias, stage, lane, n0, nt, row_pix, row_img, split, ch0, chstep, acc_ready, acc_parity, slot, ncols);
    } else if (p.geglu) {
        SDB_EPI(EPI_GEGLU, false);
    } else if (partial) {
        SDB_EPI(EPI_PARTIAL, false);
    } else {
        const bool has_add = p.residual != nullptr || (p.rowvec != nullptr && p.conv);
        if (p.out_bf16) { if (has_add) SDB_EPI(EPI_BF16, true); else SDB_EPI(EPI_BF16, false); }
        else            { if (has_add) SDB_EPI(EPI_F32, true);  else SDB_EPI(EPI_F32, false); }
    }
#undef SDB_EPI
}

// ---- TMA epilogue of the pair kernel ---------------------------------------------------------------------------
// One epilogue warp owns the 32 rows of a TMEM lane quarter and every other 32-column chunk of the accumulator (two warps per
// quarter).  tcgen05.ld hands each lane one ROW of a chunk; the lane adds bias / time-embedding row / residual (or forms the
// GEGLU product), and writes its row into a 32-row staging box in shared memory, 16-byte pieces XOR-swizzled by the row index
// exactly as the tensor map's swizzle mode expects (conflict-free: the 8 lanes of a store phase hit 8 different bank groups).
// ONE lane then issues cp.async.bulk.tensor (shared -> global): the TMA unit writes whole 128-byte (fp32) / 64-byte (bf16) row
// segments and clips rows / columns outside the tensor, so the warp spends no instructions on addresses, predicates or
// per-row stores, and nothing is transposed.  An fp32 residual tile arrives the same way in the other direction — a TMA load
// into the box that will hold the result, issued `nbuf - 1` chunks ahead (across work units), completion on a per-box mbarrier —
// and is updated in place.  The next chunk's tcgen05.ld is in flight while this chunk is processed, and the accumulator is
// handed back to the MMA issuer as soon as its last chunk sits in registers, before the stores have drained.
// Replaces epilogue_fast for: fp32 out (+bias, +time-embedding row, +fp32 residual, +GroupNorm column statistics, or split-K
// partials), bf16 out (+bias), GEGLU -> bf16.  Layouts it does not cover (sub-pixel phase remap, padded-head column remap,
// bf16 out with residual) keep the register-store epilogue.
struct UnitBox { int split, nt, mt, col0, w, h, n; };

// One 32-row x 32-column chunk of an accumulator (this lane's row in `cur`, GEGLU gate in `g`) -> bias / residual (already in
// the staging box `sbuf`) / time-embedding row / GEGLU -> swizzled staging box -> ONE TMA store (+ GroupNorm column statistics).
// Shared by the persistent CTA-pair kernel and the one-CTA kernel.
template <int BN, int MODE, bool HAS_RES>
__device__ __forceinline__ void tepi_chunk(const TcP& p, const CUtensorMap* tmC, const uint32_t (&cur)[32], const uint32_t* g,
                                           uint32_t sb_a, uint32_t sbuf, int c0, int col, const UnitBox& ub, int lane, int lg,
                                           uint32_t rowmask, int img_w, bool rv_ok, const float* rowv, bool want_cs, bool partial,
                                           int n_out) {
    constexpr bool GEGLU = MODE == EPI_GEGLU;
    constexpr bool OUT16 = MODE != EPI_F32;
    if (OUT16) {
        // bf16 rows of 64 B: four 16-byte pieces, piece j at (j ^ ((row >> 1) & 3)) (SWIZZLE_64B)
        const uint32_t rowa = sbuf + lane * 64;
        const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const float4 b4 = lds128(sb_a + 4 * (c0 + 8 * j + 4 * h2));
                float x0 = __uint_as_float(cur[8 * j + 4 * h2 + 0]) + b4.x, x1 = __uint_as_float(cur[8 * j + 4 * h2 + 1]) + b4.y;
                float x2 = __uint_as_float(cur[8 * j + 4 * h2 + 2]) + b4.z, x3 = __uint_as_float(cur[8 * j + 4 * h2 + 3]) + b4.w;
                if (GEGLU) {
                    const float4 bg = lds128(sb_a + 4 * (BN / 2 + c0 + 8 * j + 4 * h2));
                    x0 *= gelu_erf_fast(__uint_as_float(g[(8 * j + 4 * h2 + 0) % (GEGLU ? 32 : 1)]) + bg.x);
                    x1 *= gelu_erf_fast(__uint_as_float(g[(8 * j + 4 * h2 + 1) % (GEGLU ? 32 : 1)]) + bg.y);
                    x2 *= gelu_erf_fast(__uint_as_float(g[(8 * j + 4 * h2 + 2) % (GEGLU ? 32 : 1)]) + bg.z);
                    x3 *= gelu_erf_fast(__uint_as_float(g[(8 * j + 4 * h2 + 3) % (GEGLU ? 32 : 1)]) + bg.w);
                }
                v[4 * h2 + 0] = x0; v[4 * h2 + 1] = x1; v[4 * h2 + 2] = x2; v[4 * h2 + 3] = x3;
            }
            sts128(rowa + ((((uint32_t)j) ^ sw) << 4), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                   pack_bf16x2(v[6], v[7]));
        }
    } else {
        // fp32 rows of 128 B: eight 16-byte pieces, piece j at (j ^ (row & 7)) (SWIZZLE_128B); residual updated in place
        const uint32_t rowa = sbuf + lane * 128;
        const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t pa = rowa + ((((uint32_t)j) ^ sw) << 4);
            float4 v = make_float4(__uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]), __uint_as_float(cur[4 * j + 2]),
                                   __uint_as_float(cur[4 * j + 3]));
            if (!partial) {
                const float4 b4 = lds128(sb_a + 4 * (c0 + 4 * j));
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            }
            if (HAS_RES) {
                const float4 r4 = lds128(pa);
                v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
            }
            if (rowv) {
                // time-embedding row of this quarter's image (same address in every lane: one broadcast transaction), added last
                const float4 t4 = ldg_f4_or_zero(rowv + (long long)img_w * p.ldv + col + 4 * j, rv_ok && col + 4 * j < p.N);
                v.x += t4.x; v.y += t4.y; v.z += t4.z; v.w += t4.w;
            }
            sts128f(pa, v.x, v.y, v.z, v.w);
        }
    }
    fence_proxy_async_smem();                     // generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
        tma_store_5d(tmC, sbuf, col, ub.w, ub.h, ub.n, ub.split);
        tma_store_commit();
    }
    if (want_cs) {
        // GroupNorm column statistics of the stored values, read back from the staged box 128 bits at a time: this lane owns the
        // column quad 4 * (lane & 7) of rows 4 i + (lane >> 3) (one row = the 8 pieces of a quarter-warp: conflict-free), then the
        // 4 lanes that share a quad are folded in a fixed order — the same summation order as the register-store epilogue
        const int rsel = lane >> 3;
        const uint32_t qd = (uint32_t)lane & 7u;
        float4 cs_s = make_float4(0.f, 0.f, 0.f, 0.f), cs_q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsel;
            const float4 v = lds128(sbuf + r * 128 + ((qd ^ ((uint32_t)r & 7u)) << 4));
            if ((rowmask >> r) & 1u) {
                cs_s.x += v.x; cs_s.y += v.y; cs_s.z += v.z; cs_s.w += v.w;
                cs_q.x = fmaf(v.x, v.x, cs_q.x); cs_q.y = fmaf(v.y, v.y, cs_q.y);
                cs_q.z = fmaf(v.z, v.z, cs_q.z); cs_q.w = fmaf(v.w, v.w, cs_q.w);
            }
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            cs_s.x += __shfl_xor_sync(0xffffffffu, cs_s.x, o); cs_s.y += __shfl_xor_sync(0xffffffffu, cs_s.y, o);
            cs_s.z += __shfl_xor_sync(0xffffffffu, cs_s.z, o); cs_s.w += __shfl_xor_sync(0xffffffffu, cs_s.w, o);
            cs_q.x += __shfl_xor_sync(0xffffffffu, cs_q.x, o); cs_q.y += __shfl_xor_sync(0xffffffffu, cs_q.y, o);
            cs_q.z += __shfl_xor_sync(0xffffffffu, cs_q.z, o); cs_q.w += __shfl_xor_sync(0xffffffffu, cs_q.w, o);
        }
        const int cn = col + 4 * (int)qd;
        if (lane < 8 && cn < n_out) {
            float* dst = p.colstats + (long long)(ub.mt * 4 + lg) * p.N + cn;
            *reinterpret_cast<float4*>(dst) = cs_s;
            *reinterpret_cast<float4*>(dst + p.colstats_sq) = cs_q;
        }
    }
}

template <int BN, int MODE, bool HAS_RES>
__device__ __forceinline__ void epilogue_tma_units(const TcP& p, const CUtensorMap* tmC, const CUtensorMap* tmR, uint32_t tmem_d,
                                                   float* s_bias2, uint32_t stg_a, uint64_t* rbar, uint64_t* tfull_bar,
                                                   uint32_t tempty_leader0, uint32_t tempty_leader1, int warp, int lane, uint32_t rank,
                                                   int pair, int npairs, int units) {
    constexpr bool GEGLU = MODE == EPI_GEGLU;
    constexpr bool OUT16 = MODE != EPI_F32;
    constexpr int COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCHUNK = (COLS + 31) / 32;
    constexpr int CPW_MAX = (NCHUNK + 1) / 2;
    constexpr int ACC_STRIDE = 256;
    const int lg = warp & 3, half = (warp - 4) >> 2;
    const int et = threadIdx.x - 128;
    const int cpw = (NCHUNK - half + 1) / 2;             // chunks of this warp per work unit: half, half + 2, ...
    const int nbuf = p.epi_nbuf, lead = nbuf - 1;
    const bool partial = p.split_k > 1;
    const float* const biasp = partial ? nullptr : p.bias;
    const float* const rowv = (!partial && p.rowvec && p.conv && MODE == EPI_F32) ? p.rowvec : nullptr;
    const bool want_cs = MODE == EPI_F32 && p.colstats != nullptr && !partial;
    const int n_out = GEGLU ? p.N / 2 : p.N;
    const int r0 = lg * 32;
    int iw0 = 0, ih0 = 0, in0 = 0;                        // origin of this quarter's pixel box inside a conv tile
    if (p.conv) { iw0 = r0 % p.tw; ih0 = (r0 / p.tw) % p.th; in0 = r0 / (p.tw * p.th); }

    auto unit_box = [&](int u) -> UnitBox {
        UnitBox b;
        b.split = u % p.split_k;
        const int t = u / p.split_k;
        b.nt = t % p.tiles_n;
        b.mt = (t / p.tiles_n) * 2 + (int)rank;
        b.col0 = GEGLU ? b.nt * (BN / 2) : b.nt * BN;
        if (p.conv) {
            const int tww = b.mt % p.tiles_w, thh = (b.mt / p.tiles_w) % p.tiles_h, tnb = b.mt / (p.tiles_w * p.tiles_h);
            b.w = tww * p.tw + iw0; b.h = thh * p.th + ih0; b.n = tnb * p.tn + in0;
        } else {
            b.w = b.mt * TC_BM + r0; b.h = 0; b.n = 0;
        }
        return b;
    };
    // residual tile of this warp's chunk number `qq` (counted over all its work units) -> staging box qq % nbuf
    auto issue_residual = [&](uint32_t qq) {
        if (!HAS_RES || lane != 0) return;
        const int itq = (int)(qq / (uint32_t)cpw), k = (int)(qq % (uint32_t)cpw);
        const long long u = (long long)pair + (long long)itq * npairs;
        if (u >= units) return;
        const UnitBox b = unit_box((int)u);
        const uint32_t bi = qq % (uint32_t)nbuf;
        const uint32_t bar = smem_u32(rbar + bi);
        mbar_arrive_expect_tx_a(bar, 32 * 128);
        tma_load_5d(stg_a + bi * TEPI_BOX_BYTES, tmR, bar, b.col0 + (half + 2 * k) * 32, b.w, b.h, b.n, 0);
    };

    // the residual boxes of a whole work unit -> L2, one unit ahead of the TMA loads that bring them into the staging boxes:
    // those loads are issued only `lead` chunks ahead (the staging is small) and must not each pay a DRAM round trip
    auto prefetch_residual_unit = [&](long long u) {
        if (!HAS_RES || lane != 0 || u >= units) return;
        const UnitBox b = unit_box((int)u);
        for (int k = 0; k < cpw; ++k) tma_prefetch_l2_5d(tmR, b.col0 + (half + 2 * k) * 32, b.w, b.h, b.n, 0);
    };
    prefetch_residual_unit(pair);
    uint32_t q = 0;                                       // chunks this warp has processed
    for (int j = 0; j < lead; ++j) issue_residual((uint32_t)j);
    // bias row of a unit: fetched one unit ahead into a register, published into the unit's half of s_bias2
    auto bias_of = [&](int u) -> float {
        if (u >= units || et >= BN || biasp == nullptr) return 0.f;
        const int n = ((u / p.split_k) % p.tiles_n) * BN + et;
        return n < p.N ? __ldg(biasp + n) : 0.f;
    };
    float bias_next = bias_of(pair);
    int it = 0;
    for (int u = pair; u < units; u += npairs, ++it) {
        const UnitBox ub = unit_box(u);
        const int buf = it & 1;
        float* const sb = s_bias2 + buf * BN;
        if (et < BN) sb[et] = bias_next;
        bias_next = bias_of(u + npairs);
        prefetch_residual_unit((long long)u + npairs);
        named_bar_sync(1, 32 * 8);                        // bias row visible; every warp is done with unit it - 1 (and so with sb of it - 2)
        const uint32_t sb_a = smem_u32(sb);
        // rows of this quarter that exist (conv tiles may hang over the image / batch edge; gemm rows past M)
        uint32_t rowmask = 0xffffffffu;
        int img_w = 0;
        if (want_cs || rowv) {
            bool valid;
            int img = 0;
            if (p.conv) {
                const int row = r0 + lane;
                const int in_ = row / (p.th * p.tw), rem = row - in_ * (p.th * p.tw);
                const int ih = rem / p.tw, iw = rem - ih * p.tw;
                const int tww = ub.mt % p.tiles_w, thh = (ub.mt / p.tiles_w) % p.tiles_h, tnb = ub.mt / (p.tiles_w * p.tiles_h);
                img = tnb * p.tn + in_;
                valid = img < p.NB && thh * p.th + ih < p.OH && tww * p.tw + iw < p.OW;
            } else {
                valid = ub.mt * TC_BM + r0 + lane < p.M;
            }
            rowmask = __ballot_sync(0xffffffffu, valid);
            img_w = __shfl_sync(0xffffffffu, img, 0);     // the launcher admits a time-embedding row only when a quarter lies in one image
        }
        const bool rv_ok = rowv != nullptr && rowmask != 0u && img_w < p.NB;
        const uint32_t taddr = tmem_d + ((uint32_t)r0 << 16) + buf * ACC_STRIDE;
        mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
        tcgen05_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld_x32(taddr + half * 32, ra);
#pragma unroll
        for (int k = 0; k < CPW_MAX; ++k) {
            const int ch = half + 2 * k;
            if (ch >= NCHUNK) break;                      // warp-uniform
            uint32_t (&cur)[32] = (k & 1) ? rb : ra;
            uint32_t (&nxt)[32] = (k & 1) ? ra : rb;
            const int c0 = ch * 32;
            const int col = ub.col0 + c0;                 // first output column of the chunk
            const uint32_t bi = q % (uint32_t)nbuf;
            const uint32_t sbuf = stg_a + bi * TEPI_BOX_BYTES;
            uint32_t g[GEGLU ? 32 : 1];
            if (GEGLU) tmem_ld_x32(taddr + BN / 2 + c0, g);
            tmem_ld_wait();                               // chunk k (and its gate half) is in registers
            const bool last = ch + 2 >= NCHUNK;
            if (!last) {
                tmem_ld_x32(taddr + c0 + 64, nxt);        // next chunk's read overlaps this chunk's processing
            } else {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);   // accumulator drained: back to the MMA issuer
            }
            if (HAS_RES) {
                mbar_wait(rbar + bi, (q / (uint32_t)nbuf) & 1u);  // this chunk's residual tile has landed in its box
            } else {
                if (lane == 0) tma_store_wait_read<1>();  // the store of chunk q - 2 has finished reading this box
                __syncwarp();
            }
            tepi_chunk<BN, MODE, HAS_RES>(p, tmC, cur, g, sb_a, sbuf, c0, col, ub, lane, lg, rowmask, img_w, rv_ok, rowv, want_cs, partial, n_out);
            if (HAS_RES) {
                if (lane == 0) tma_store_wait_read<1>();  // store of chunk q - 1 done reading -> its box takes the residual of chunk q + lead
                issue_residual(q + (uint32_t)lead);
            }
            ++q;
        }
    }
    // the staging boxes must outlive the TMA unit's READS of them; the global writes themselves complete asynchronously and are
    // visible at grid completion like any other store (waiting for them here kept every CTA resident for a DRAM round trip)
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
}

// The same epilogue for the one-CTA kernel (one 128 x BN tile per CTA, two CTAs per SM).  Once the accumulator is complete every
// operand load has been consumed, so the A/B ring is the staging area: each epilogue warp owns one box per 32-column chunk.
// ALL residual boxes of the warp are requested at once the moment the ring is free (one mbarrier, one L2 round trip — they were
// prefetched into L2 when the CTA started) instead of one exposed round trip per chunk, which is what kept the K <= 640
// "+ residual" layers at 2x their memory bound (profiles/r02_ncu_full_summary.txt: long_scoreboard on the residual adds).
template <int BN, int MODE, bool HAS_RES>
__device__ __forceinline__ void epilogue_tma_single(const TcP& p, const CUtensorMap* tmC, const CUtensorMap* tmR, uint32_t taddr,
                                                    uint32_t sb_a, uint32_t stg_a, uint64_t* rbar, uint64_t* accum_bar, const UnitBox& ub,
                                                    int lane, int lg, uint32_t rowmask, int img_w, int ncols) {
    constexpr bool GEGLU = MODE == EPI_GEGLU;
    constexpr int COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCHUNK = (COLS + 31) / 32;
    const int nch = ncols >= COLS ? NCHUNK : (ncols + 31) / 32;         // chunks of this CTA's (sub-)tile
    const bool partial = p.split_k > 1;
    const float* const rowv = (!partial && p.rowvec && p.conv && MODE == EPI_F32) ? p.rowvec : nullptr;
    const bool want_cs = MODE == EPI_F32 && p.colstats != nullptr && !partial;
    const int n_out = GEGLU ? p.N / 2 : p.N;
    const bool rv_ok = rowv != nullptr && rowmask != 0u && img_w < p.NB;
    if (HAS_RES && lane == 0) {
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch)
            if (ch < nch) tma_prefetch_l2_5d(tmR, ub.col0 + ch * 32, ub.w, ub.h, ub.n, 0);
    }
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    if (HAS_RES && lane == 0) {
        mbar_arrive_expect_tx_a(smem_u32(rbar), nch * 32 * 128);
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch)
            if (ch < nch) tma_load_5d(stg_a + ch * TEPI_BOX_BYTES, tmR, smem_u32(rbar), ub.col0 + ch * 32, ub.w, ub.h, ub.n, 0);
    }
    uint32_t ra[32], rb[32];
    tmem_ld_x32(taddr, ra);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        if (ch >= nch) break;                                           // warp-uniform
        uint32_t (&cur)[32] = (ch & 1) ? rb : ra;
        uint32_t (&nxt)[32] = (ch & 1) ? ra : rb;
        const int c0 = ch * 32;
        uint32_t g[GEGLU ? 32 : 1];
        if (GEGLU) tmem_ld_x32(taddr + BN / 2 + c0, g);
        tmem_ld_wait();
        if (ch + 1 < nch) tmem_ld_x32(taddr + c0 + 32, nxt);
        if (HAS_RES && ch == 0) mbar_wait(rbar, 0);
        tepi_chunk<BN, MODE, HAS_RES>(p, tmC, cur, g, sb_a, stg_a + ch * TEPI_BOX_BYTES, c0, ub.col0 + c0, ub, lane, lg, rowmask, img_w, rv_ok,
                                      rowv, want_cs, partial, n_out);
    }
    if (lane == 0) tma_store_wait_read<0>();               // smem may be released once the TMA unit has read the boxes
    __syncwarp();
}

template <int BN>
__global__ void __launch_bounds__(192, 2)
tc_contract_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcP p) {
    pdl_trigger();
    if (threadIdx.x == 0) { TC1_TRACE(0); TC1_TRACE(7); }
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * TC_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    uint64_t* res_bar = accum_bar + 2;               // [4] TMA epilogue: the residual boxes of epilogue warp w have landed
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    int tile = blockIdx.x;
    int ncols = BN, n_off = 0;                       // columns of the 128 x BN tile this CTA computes
    if (tile >= p.full_tiles) {
        // tail split: the tiles past full_tiles are computed as a 96-column and a (BN - 96)-column sub-tile by two CTAs each, so the
        // last, partial wave is spread over all SMs in half-size pieces (same K order per output element: identical bits)
        const int t2 = tile - p.full_tiles;
        tile = p.full_tiles + (t2 >> 1);
        if (t2 & 1) { n_off = 96; ncols = BN - 96; } else { ncols = 96; }
    }
    const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
    const int n0 = nt * BN + n_off;
    const int split = blockIdx.y;
    const int kb0 = (int)((long long)p.kblocks * split / p.split_k);
    const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.split_k);

    int m0 = mt * TC_BM;            // gemm mode
    int ow0 = 0, oh0 = 0, img0 = 0; // conv mode
    if (p.conv) {
        int tww = mt % p.tiles_w;
        int thh = (mt / p.tiles_w) % p.tiles_h;
        int tnb = mt / (p.tiles_w * p.tiles_h);
        ow0 = tww * p.tw; oh0 = thh * p.th; img0 = tnb * p.tn;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        for (int w = 0; w < 4; ++w) mbar_init(&res_bar[w], 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    if (threadIdx.x == 0) TC1_TRACE(1);
    // weights do not depend on the preceding launch: arm the first stages and fetch their B tiles while it is still draining
    int b_pre = 0;
    if (p.b_const && warp == 0 && lane == 0) {
        b_pre = kb1 - kb0 < STAGES ? kb1 - kb0 : STAGES;
        for (int i = 0; i < b_pre; ++i) {
            const int kb = kb0 + i;
            const int tap = kb / p.kpt, cs = kb - tap * p.kpt;
            mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES);
            tma_load_2d(sB + i * Cfg::B_BYTES, &tmB, &full_bar[i], cs * TC_BK, tap * p.cout_pad + n0);
        }
    }
    pdl_wait();                                      // predecessor's outputs (A, residual, ...) are complete and visible
    if (threadIdx.x == 0) TC1_TRACE(2);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const bool pre = kb - kb0 < b_pre;             // stage armed and its B tile requested before pdl_wait()
                if (!pre) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                }
                const int tap = kb / p.kpt, cs = kb - tap * p.kpt;
                if (p.conv) {
                    const int r = tap / p.kw, sx = tap - r * p.kw;
                    tma_load_4d(sA + s * TC_A_BYTES, &tmA, &full_bar[s], cs * TC_BK,
                                ow0 * p.stride + sx - p.pad_w, oh0 * p.stride + r - p.pad_h, img0);
                } else {
                    tma_load_2d(sA + s * TC_A_BYTES, &tmA, &full_bar[s], cs * TC_BK, m0);
                }
                if (!pre) tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], cs * TC_BK, tap * p.cout_pad + n0);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16((uint32_t)ncols, false, false);
            int s = 0; uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sA + s * TC_A_BYTES));
                const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sB + s * Cfg::B_BYTES));
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                    // advance 16 bf16 = 32 B along K inside the 128-B swizzle atom: +2 (16-B units)
                    umma_bf16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // smem slot reusable once these MMAs retire
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(accum_bar);                  // accumulator complete
            TC1_TRACE(3);
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1) =================
        const int et = threadIdx.x - 64;             // 0..127
        for (int i = et; i < BN; i += 128) {
            int n = n0 + i;
            s_bias[i] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
        }
        named_bar_sync(1, 128);

        const int lg = warp & 3;                     // TMEM lane group this warp may access
        const int row = lg * 32 + lane;              // row of the 128-row tile
        long long pix;                               // output row index (pixel / token), -1 = outside the problem
        int img = 0;
        if (p.conv) {
            int in_ = row / (p.th * p.tw);
            int rem = row - in_ * (p.th * p.tw);
            int ih = rem / p.tw, iw = rem - ih * p.tw;
            img = img0 + in_;
            int oh = oh0 + ih, ow = ow0 + iw;
            bool valid = (img < p.NB) && (oh < p.OH) && (ow < p.OW);
            pix = valid ? ((long long)img * p.OHF + (oh * p.out_sh + p.out_oh)) * p.OWF + (ow * p.out_sw + p.out_ow) : -1;
        } else {
            pix = (m0 + row) < p.M ? m0 + row : -1;
        }
        const uint32_t taddr = tmem_d + ((uint32_t)(lg * 32) << 16);
        if (p.epi_tma) {
            // ---- TMA epilogue (epilogue_tma_single): the ring is the staging area, one box per (warp, chunk) ----
            constexpr int NCH = BN / 32 > 0 ? BN / 32 : 1;
            UnitBox ub;
            ub.split = split; ub.nt = nt; ub.mt = mt; ub.col0 = p.geglu ? nt * (BN / 2) : n0;
            const int r0 = lg * 32;
            if (p.conv) {
                ub.w = ow0 + r0 % p.tw; ub.h = oh0 + (r0 / p.tw) % p.th; ub.n = img0 + r0 / (p.tw * p.th);
            } else {
                ub.w = m0 + r0; ub.h = 0; ub.n = 0;
            }
            const uint32_t rowmask = __ballot_sync(0xffffffffu, pix >= 0);
            const int img_w = __shfl_sync(0xffffffffu, img, 0);
            const uint32_t stg_a = smem_u32(sA) + (uint32_t)(warp - 2) * (uint32_t)(NCH * TEPI_BOX_BYTES);
            uint64_t* rbar = &res_bar[warp - 2];
#define SDB_TEPI1(MODE, RES) epilogue_tma_single<BN, MODE, RES>(p, &tmC, &tmR, taddr, smem_u32(s_bias), stg_a, rbar, accum_bar, ub, lane, lg, rowmask, img_w, ncols)
            if (p.geglu) {
                if constexpr (BN % 64 == 0) SDB_TEPI1(EPI_GEGLU, false);
            } else if (p.split_k > 1) SDB_TEPI1(EPI_F32, false);
            else if (p.out_bf16) SDB_TEPI1(EPI_BF16, false);
            else if (p.residual != nullptr) SDB_TEPI1(EPI_F32, true);
            else SDB_TEPI1(EPI_F32, false);
#undef SDB_TEPI1
        } else {
            prefetch_residual_row(p, pix, n0, BN);
            // once the accumulator is ready every TMA load has been consumed: the A ring doubles as the transpose staging area
            float* stage = reinterpret_cast<float*>(sA) + (warp - 2) * (EPI_WARP_BYTES / 4);
            if (warp == 2 && lane == 0) TC1_TRACE(4);
            epilogue_warp<BN>(p, taddr, s_bias, stage, lane, n0, nt, pix, img, split, 0, 1, accum_bar, 0, mt * 4 + lg, ncols);
            if (warp == 2 && lane == 0) TC1_TRACE(5);
        }
    }

    // ---- teardown ----
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_d, Cfg::TMEM_COLS);
    }
}

// ---- CTA-pair persistent kernel ------------------------------------------------------------------
// A cluster of two CTAs (one per SM of a TPC) owns a 256 x BN output tile: tcgen05.mma.cta_group::2
// with M = 256, each CTA staging its own 128 rows of A and HALF of the B tile (BN/2 rows), so the
// bytes pulled from L2 per FLOP drop by 1/3 .. 1/2 against the one-CTA 128 x BN kernel.  The pair is
// persistent over a strided list of work units (m-pair, n-tile, k-split); the 512-column TMEM holds
// two accumulators so the epilogue of unit i overlaps the mainloop of unit i+1.
//   warp 0 : TMA producer for A, warp 3 : TMA producer for B (both CTAs; completion bytes are
//            signalled on the LEADER's full barrier, which the leader's A-producer arms)
//   warp 1 : MMA issuer (leader CTA only); tcgen05.commit multicasts "slot free" / "accumulator
//            ready" to the barriers of both CTAs
//   warp 2 : TMEM allocator
//   warps 4..4+EW-1 : epilogue, EW/4 warps per TMEM lane quarter (warp & 3), interleaved 32-column chunks.
// EW = 8 in the product; 16 exists for the measurement build only (see launch_tc_pair).
constexpr int TC2_MAX_STAGES = 8;
constexpr int TC2_SMEM_LIMIT = 232448;            // 227 KB: the per-CTA dynamic shared memory limit of sm_100

template <int BN, int EW>
struct Tc2Cfg {
    static constexpr int THREADS = 128 + 32 * EW;
    static constexpr int B_HALF_BYTES = (BN / 2) * TC_BK * 2;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_HALF_BYTES;       // per CTA
    static constexpr int ACC_STRIDE = 256;                                // TMEM columns between the two accumulators
    // layout: [<= 1023 alignment slack][ring: stages x (A | B half)][512 B barriers][2 x BN bias floats][pad to 1024][epilogue staging]
    static constexpr int FIXED_BYTES = 1024 + 512 + 2 * BN * 4 + 1024;
    __host__ __device__ static constexpr int staging_bytes(int epi_tma, int nbuf) { return epi_tma ? EW * nbuf * TEPI_BOX_BYTES : EW * EPI_WARP_BYTES; }
    __host__ static int stages_for(int staging) {
        int st = (TC2_SMEM_LIMIT - FIXED_BYTES - staging) / STAGE_BYTES;
        return st > TC2_MAX_STAGES ? TC2_MAX_STAGES : st;
    }
    __host__ static int smem_bytes(int stages, int staging) { return FIXED_BYTES + stages * STAGE_BYTES + staging; }
};

template <int BN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
tc_contract_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcP p) {
    pdl_trigger();
    using Cfg = Tc2Cfg<BN, EW>;
    constexpr int TC2_EPI_WARPS = EW;
    const int STAGES = p.stages;                     // ring depth: what the epilogue staging leaves of the 227 KB
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * TC_A_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + TC2_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + TC2_MAX_STAGES;   // [2] accumulator ready   (arrives: MMA commit, multicast)
    uint64_t* tempty_bar = tfull_bar + 2;            // [2] accumulator drained (leader's copy is the one waited on)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint64_t* res_bar = tempty_bar + 4;              // [EW][3] TMA epilogue: residual tile landed in staging box b of warp w
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 512);      // [2][BN] (the register-store epilogue uses half 0)
    uint8_t* staging = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_bias + 2 * BN) + 1023) & ~(uintptr_t)1023);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int units = p.m_pairs * p.tiles_n * p.split_k;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * TC2_EPI_WARPS); }
        for (int b = 0; b < 3 * EW && EW == 8; ++b) mbar_init(&res_bar[b], 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, 512);
        tmem_relinquish_pair();
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    // the B producer (warp 3) reads nothing but the weights: with b_const it does not wait for the preceding launch, so the first
    // `stages` B tiles are in flight (from DRAM) while that launch is still draining; everybody else waits
    if (!(warp == 3 && p.b_const)) pdl_wait();       // predecessor's outputs (A, residual, ...) are complete and visible

    if (warp == 0 || warp == 3) {
        // ================= TMA producers (both CTAs): warp 0 loads A, warp 3 loads B =================
        if (lane == 0) {
            const bool load_a = warp == 0;
            const uint32_t fb0 = mapa_u32(&full_bar[0], 0);           // leader's full barriers (8 bytes apart)
            int s = 0; uint32_t ph = 0;
            for (int u = pair; u < units; u += npairs) {
                const int split = u % p.split_k, t = u / p.split_k;
                const int nt = t % p.tiles_n, mt = (t / p.tiles_n) * 2 + (int)rank;
                const int brow0 = nt * BN + (int)rank * (BN / 2);
                const int kb0 = (int)((long long)p.kblocks * split / p.split_k);
                const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.split_k);
                const int m0 = mt * TC_BM;
                int cw = 0, chh = 0, img0 = 0;
                if (p.conv) {
                    int tww = mt % p.tiles_w;
                    int thh = (mt / p.tiles_w) % p.tiles_h;
                    int tnb = mt / (p.tiles_w * p.tiles_h);
                    cw = tww * p.tw * p.stride - p.pad_w; chh = thh * p.th * p.stride - p.pad_h; img0 = tnb * p.tn;
                }
                int tap = kb0 / p.kpt, cs = kb0 - tap * p.kpt;
                int r = p.conv ? tap / p.kw : 0, sx = p.conv ? tap - r * p.kw : 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (load_a && kb == kb0) TC_TRACE(6, (u - pair) / npairs);
                    if (load_a && kb == kb1 - 1) TC_TRACE(7, (u - pair) / npairs);
                    const uint32_t fb = fb0 + 8u * (uint32_t)s;
                    if (load_a) {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
                        if (p.conv) tma_load_4d_pair(sA + s * TC_A_BYTES, &tmA, fb, cs * TC_BK, cw + sx, chh + r, img0);
                        else tma_load_2d_pair(sA + s * TC_A_BYTES, &tmA, fb, cs * TC_BK, m0);
                    } else {
                        tma_load_2d_pair(sB + s * Cfg::B_HALF_BYTES, &tmB, fb, cs * TC_BK, tap * p.cout_pad + brow0);
                    }
                    if (++cs == p.kpt) { cs = 0; ++tap; if (++sx == p.kw) { sx = 0; ++r; } }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA) =================
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_m256(BN, false, false);
            const uint64_t adesc0 = umma_desc_kmajor_sw128(smem_u32(sA));
            const uint64_t bdesc0 = umma_desc_kmajor_sw128(smem_u32(sB));
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int u = pair; u < units; u += npairs, ++it) {
                const int split = u % p.split_k;
                const int kb0 = (int)((long long)p.kblocks * split / p.split_k);
                const int kb1 = (int)((long long)p.kblocks * (split + 1) / p.split_k);
                const int buf = it & 1;
                TC_TRACE(5, it);
                mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);     // both CTAs' epilogues drained this accumulator
                tcgen05_fence_after();
                TC_TRACE(0, it);
                const uint32_t acc = tmem_d + buf * Cfg::ACC_STRIDE;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    if (kb == kb0) TC_TRACE(1, it);
                    const uint64_t adesc = adesc0 + (uint64_t)(s * (TC_A_BYTES >> 4));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(s * (Cfg::B_HALF_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma_bf16_ss_pair(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(&empty_bar[s], 3);              // slot free in both CTAs
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_pair(&tfull_bar[buf], 3);                // accumulator ready in both CTAs
                TC_TRACE(2, it);
            }
        }
    } else if (warp >= 4 && EW == 8 && p.epi_tma) {
        // ================= epilogue through TMA (see epilogue_tma_units) =================
        const uint32_t tl0 = mapa_u32(&tempty_bar[0], 0), tl1 = mapa_u32(&tempty_bar[1], 0);
        const uint32_t stg_a = smem_u32(staging) + (uint32_t)(warp - 4) * (uint32_t)(p.epi_nbuf * TEPI_BOX_BYTES);
        uint64_t* rbar = res_bar + 3 * (warp - 4);
#define SDB_TEPI(MODE, RES) epilogue_tma_units<BN, MODE, RES>(p, &tmC, &tmR, tmem_d, s_bias, stg_a, rbar, tfull_bar, tl0, tl1, warp, lane, rank, pair, npairs, units)
        if (p.geglu) {
            if constexpr (BN % 64 == 0) SDB_TEPI(EPI_GEGLU, false);
        } else if (p.split_k > 1) SDB_TEPI(EPI_F32, false);
        else if (p.out_bf16) SDB_TEPI(EPI_BF16, false);
        else if (p.residual != nullptr) SDB_TEPI(EPI_F32, true);
        else SDB_TEPI(EPI_F32, false);
#undef SDB_TEPI
    } else if (warp >= 4) {
        // ================= epilogue (warps 4.. -> TMEM lane quarter warp & 3, chunk phase (warp - 4) >> 2 of EW/4) ====
        const int et = threadIdx.x - 128;
        const int lg = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = lg * 32 + lane;
        const uint32_t tempty_leader0 = mapa_u32(&tempty_bar[0], 0);
        const uint32_t tempty_leader1 = mapa_u32(&tempty_bar[1], 0);
        float* stage = reinterpret_cast<float*>(staging) + (warp - 4) * (EPI_WARP_BYTES / 4);
        // output row (pixel / token) of this lane's tile row in m-tile `mt`, -1 = outside the problem
        auto row_of = [&](int mt, int& img) -> long long {
            img = 0;
            if (p.conv) {
                int tww = mt % p.tiles_w;
                int thh = (mt / p.tiles_w) % p.tiles_h;
                int tnb = mt / (p.tiles_w * p.tiles_h);
                int in_ = row / (p.th * p.tw);
                int rem = row - in_ * (p.th * p.tw);
                int ih = rem / p.tw, iw = rem - ih * p.tw;
                img = tnb * p.tn + in_;
                int oh = thh * p.th + ih, ow = tww * p.tw + iw;
                bool valid = (img < p.NB) && (oh < p.OH) && (ow < p.OW);
                return valid ? ((long long)img * p.OHF + (oh * p.out_sh + p.out_oh)) * p.OWF + (ow * p.out_sw + p.out_ow) : -1;
            }
            return (mt * TC_BM + row) < p.M ? (long long)mt * TC_BM + row : -1;
        };
        auto prefetch_unit = [&](int u) {             // residual rows of work unit u (one warp per lane quarter issues)
            if (half != 0 || u >= units || p.residual == nullptr) return;
            const int t = u / p.split_k;
            const int nt = t % p.tiles_n, mt = (t / p.tiles_n) * 2 + (int)rank;
            int img_;
            prefetch_residual_row(p, row_of(mt, img_), nt * BN, BN);
        };
        prefetch_unit(pair);
        int it = 0;
        for (int u = pair; u < units; u += npairs, ++it) {
            const int split = u % p.split_k, t = u / p.split_k;
            const int nt = t % p.tiles_n, mt = (t / p.tiles_n) * 2 + (int)rank;
            const int n0 = nt * BN;
            const int buf = it & 1;
            prefetch_unit(u + npairs);               // one unit ahead: in L2 by the time its epilogue starts
            named_bar_sync(1, 32 * TC2_EPI_WARPS);   // previous unit's readers of s_bias are done
            for (int i = et; i < BN; i += 32 * TC2_EPI_WARPS) {
                int n = n0 + i;
                s_bias[i] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
            }
            named_bar_sync(1, 32 * TC2_EPI_WARPS);
            int img = 0;
            const long long pix = row_of(mt, img);
            const uint32_t taddr = tmem_d + ((uint32_t)(lg * 32) << 16) + buf * Cfg::ACC_STRIDE;
            if (warp == 4 && lane == 0) TC_TRACE(3, it);
            epilogue_warp<BN>(p, taddr, s_bias, stage, lane, n0, nt, pix, img, split, half, EW / 4, &tfull_bar[buf], (it >> 1) & 1,
                              mt * 4 + lg);
            if (warp == 4 && lane == 0) TC_TRACE(4, it);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
        }
    }

    // ---- teardown: both CTAs must be done with TMEM and with each other's barriers ----
    __syncwarp();
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_pair(tmem_d, 512);
    }
}

// ---- host side: tensor-map encoding --------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// rank-R tensor map, zero OOB fill. dims/strides innermost first; strides in elements for dims 1..R-1.
// elem_bytes 2 = bf16, 4 = fp32; swizzle_bytes 128 / 64 (shared-memory swizzle mode of the box).
int make_tmap(CUtensorMap* tm, const void* base, int elem_bytes, int swizzle_bytes, int rank, const long long* dims,
              const long long* strides_elems, const int* box, const int* estr) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable"); return SDB_ERR_NOTMA; }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bdim[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)estr[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = (cuuint64_t)strides_elems[i] * (cuuint64_t)elem_bytes;
    if (((uintptr_t)base & 15) != 0) { set_last_error("tensor map base %p not 16-byte aligned", base); return SDB_ERR_INVALID; }
    for (int i = 0; i + 1 < rank; ++i)
        if (gstr[i] % 16) { set_last_error("tensor map stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i]); return SDB_ERR_INVALID; }
    CUresult r = enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                     const_cast<void*>(base), gdim, gstr, bdim, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %lld %lld box %d %d)", (int)r, rank,
                       dims[0], rank > 1 ? dims[1] : 0, box[0], rank > 1 ? box[1] : 0);
        return SDB_ERR_CUDA;
    }
    return SDB_OK;
}

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr) {
    return make_tmap(tm, base, 2, 128, rank, dims, strides_elems, box, estr);
}

// choose the pixel-block decomposition tw x th x tn = 128 with the fewest tiles
static void pick_tile(int OW, int OH, int NB, int stride, int* tw, int* th, int* tn) {
    long long best = -1;
    for (int w = 128; w >= 1; w >>= 1) {
        if (w * stride > 256) continue;
        for (int h = 128 / w; h >= 1; h >>= 1) {
            if (h * stride > 256) continue;
            int n = 128 / (w * h);
            if (n > 256) continue;
            long long tiles = (long long)ceil_div(OW, w) * ceil_div(OH, h) * ceil_div(NB, n);
            if (best < 0 || tiles < best) { best = tiles; *tw = w; *th = h; *tn = n; }
        }
    }
}

static int sm_count_cached();
static bool tail_split_enabled();

template <int BN>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, TcP& p, int m_tiles,
                     cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool attr_set_dev[64] = {false};          // the attribute is per device
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    bool& attr_set = attr_set_dev[cur_dev];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_contract_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("tc_contract: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    // TMA epilogue: one 4 KB box per (epilogue warp, 32-column chunk) must fit in the operand ring it re-uses
    constexpr int NCH = BN / 32 > 0 ? BN / 32 : 1;
    if (4 * NCH * TEPI_BOX_BYTES > Cfg::STAGES * Cfg::STAGE_BYTES) p.epi_tma = 0;
    // Tail split (BN = 160): with T tiles on S SMs the slowest SM computes ceil(T / S) of them — 512 tiles of a 64x64, N = 320 layer
    // at batch 8 are 3.46 per SM, i.e. 4 on some (86 %).  The T mod S tiles of the last, partial wave are issued as two sub-tiles
    // (96 | 64 columns) each, at the end of the grid, so that the tail is handed out in half-size pieces.  Results are bit-identical
    // (every output element sums over K in the same order whatever the tile width).
    const int total = m_tiles * p.tiles_n;
    p.full_tiles = total;
    unsigned gx = (unsigned)total;
    if (BN == 160 && tail_split_enabled() && p.split_k == 1 && !p.geglu) {
        const int S = sm_count_cached();
        const int r = total % S;
        if (total >= S && 100 * r >= 12 * S && 100 * r <= 65 * S) { p.full_tiles = total - r; gx = (unsigned)(total + r); }
    }
    dim3 grid(gx, (unsigned)p.split_k);
    launch_pdl(tc_contract_kernel<BN>, dim3(grid), dim3(192), Cfg::SMEM_BYTES, st, tmA, tmB, tmC, tmR, p);
    return check_launch("tc_contract_kernel");
}

static int sm_count_cached() {
    static int n[64] = {0};          // per device: the launch goes to the CURRENT device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

template <int BN, int EW>
static int launch_tc_pair_ew(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, TcP& p,
                             int m_tiles, cudaStream_t st) {
    using Cfg = Tc2Cfg<BN, EW>;
    static bool attr_set_dev[64] = {false};          // the attribute is per device
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    bool& attr_set = attr_set_dev[cur_dev];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_contract_pair_kernel<BN, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_LIMIT);
        if (e != cudaSuccess) { set_last_error("tc_contract(pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    if (EW != 8) p.epi_tma = 0;
    const int staging = Cfg::staging_bytes(p.epi_tma, p.epi_nbuf);
    p.stages = Cfg::stages_for(staging);
    if (p.stages < 2) { set_last_error("tc_contract(pair): no room for an operand ring (BN %d)", BN); return SDB_ERR_INVALID; }
    p.m_pairs = (m_tiles + 1) / 2;
    long long units = (long long)p.m_pairs * p.tiles_n * p.split_k;
    int pairs = sm_count_cached() / 2;
    if (units < pairs) pairs = (int)units;
    launch_pdl(tc_contract_pair_kernel<BN, EW>, dim3(dim3(2 * pairs)), dim3(Cfg::THREADS), Cfg::smem_bytes(p.stages, staging), st,
               tmA, tmB, tmC, tmR, p);
    return check_launch("tc_contract_pair_kernel");
}

// Epilogue warps of the pair kernel: 8.  A 16-warp build (-DSDB_TC_EW16, then SDB200_TC_EW=16 at run time) was measured on
// B200 and rejected: K = 320 GEGLU 77.7 -> 92.6 us, K = 320 bf16-out 44.8 -> 50.4 us, K = 320 fp32 + residual 40.1 -> 54.6 us,
// whole UNet call 12.69 -> 13.01 ms (profiles/r01_epilogue_warps.txt).  The epilogue of the short-K layers is bound by its
// instruction count (~24 warp instructions per output element, ~3.1 k issue slots per scheduler and unit against 2.6 k clocks
// of MMA at K = 320), not by latency that more warps could hide; 16 warps also spill (548 B) under the 96-register cap.
template <int BN>
static int launch_tc_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, TcP& p,
                          int m_tiles, cudaStream_t st) {
#ifdef SDB_TC_EW16
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("SDB200_TC_EW"); forced = e ? atoi(e) : 0; }
    if (forced == 16) return launch_tc_pair_ew<BN, 16>(tmA, tmB, tmC, tmR, p, m_tiles, st);
#endif
    return launch_tc_pair_ew<BN, 8>(tmA, tmB, tmC, tmR, p, m_tiles, st);
}

// Kernel selection: 1 = CTA-pair persistent kernel for BN >= 128 (default), 0 = one-CTA kernel everywhere.
// Initialised from SDB200_TC_KERNEL=single|pair, changeable at run time through sdb_tc_set_pair_kernel.
static int g_pair_kernel = -1;
static bool pair_kernel_enabled() {
    if (g_pair_kernel < 0) {
        const char* e = getenv("SDB200_TC_KERNEL");
        g_pair_kernel = (e && strcmp(e, "single") == 0) ? 0 : 1;
    }
    return g_pair_kernel == 1;
}

// Tail split of the one-CTA kernel (see launch_tc): SDB200_TC_TAILSPLIT=0 turns it off, sdb_tc_set_tail_split at run time
static int g_tail_split = -1;
static bool tail_split_enabled() {
    if (g_tail_split < 0) { const char* e = getenv("SDB200_TC_TAILSPLIT"); g_tail_split = (e && e[0] == '0') ? 0 : 1; }
    return g_tail_split == 1;
}

// TMA epilogue switches: SDB200_TC_EPI=regs forces the register-store epilogue everywhere (A/B measurements);
// SDB200_TC_EPI_MAXKB = largest number of 64-wide k-blocks per work unit for which the TMA epilogue is used (longer
// contractions hide any epilogue behind their mainloop and keep the deeper operand ring instead).
static int g_tma_epilogue = -1;
static bool tma_epilogue_enabled() {
    if (g_tma_epilogue < 0) { const char* e = getenv("SDB200_TC_EPI"); g_tma_epilogue = (e && strcmp(e, "regs") == 0) ? 0 : 1; }
    return g_tma_epilogue == 1;
}
static int tma_epilogue_max_kblocks() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SDB200_TC_EPI_MAXKB"); v = e ? atoi(e) : 40; }
    return v;
}

// ---- split-K: deterministic reduction of the per-split partial tiles --------------------------------
// out[m, n] = sum_s ws[s][m][n] + bias[n] + rowvec[m / rows_per_img][n] + residual[m][n]   (4 columns per thread)
// A conv whose output is remapped into a sub-lattice of a larger tensor (one sub-pixel phase of an upsampling conv) reduces ONLY
// its own pixels: `rows` then counts the conv's dense output rows (image, oh, ow) and `rm` maps them to rows of the workspace /
// output lattice.  (Reducing the whole lattice — as this kernel once did — rewrites the other phases' pixels from whatever their
// workspace entries hold, which is only right while every phase happens to get the same recycled workspace block.)
struct ReduceRemap { int on, OH, OW, sh, sw, oh0, ow0, OHF, OWF; };

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, long long split_stride, int splits, long long rows, int N,
                                     const float* __restrict__ bias, const float* __restrict__ rowvec, long long ldv,
                                     long long rows_per_img, const float* __restrict__ residual, long long ldr,
                                     void* __restrict__ out, long long ldc, int out_bf16, const ReduceRemap rm) {
    pdl_trigger();
    pdl_wait();
    const int nq = N >> 2;
    const long long total = rows * nq;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long m = i / nq;
        const int n = (int)(i - m * nq) << 2;
        if (rm.on) {
            const long long img = m / ((long long)rm.OH * rm.OW);
            const int rem = (int)(m - img * rm.OH * rm.OW);
            const int oh = rem / rm.OW, ow = rem - oh * rm.OW;
            m = (img * rm.OHF + (oh * rm.sh + rm.oh0)) * rm.OWF + (ow * rm.sw + rm.ow0);
        }
        float4 acc = *reinterpret_cast<const float4*>(ws + m * N + n);
        for (int s = 1; s < splits; ++s) {
            float4 q = *reinterpret_cast<const float4*>(ws + s * split_stride + m * N + n);
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
        if (bias) { acc.x += bias[n]; acc.y += bias[n + 1]; acc.z += bias[n + 2]; acc.w += bias[n + 3]; }
        if (rowvec) {
            const float* rp = rowvec + (m / rows_per_img) * ldv + n;
            acc.x += rp[0]; acc.y += rp[1]; acc.z += rp[2]; acc.w += rp[3];
        }
        if (residual) {
            const float* rp = residual + m * ldr + n;
            acc.x += rp[0]; acc.y += rp[1]; acc.z += rp[2]; acc.w += rp[3];
        }
        if (out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + m * ldc + n;
            o[0] = __float2bfloat16_rn(acc.x); o[1] = __float2bfloat16_rn(acc.y);
            o[2] = __float2bfloat16_rn(acc.z); o[3] = __float2bfloat16_rn(acc.w);
        } else {
            float* o = reinterpret_cast<float*>(out) + m * ldc + n;
            o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
        }
    }
}

// Same reduction for an fp32 output that feeds a GroupNorm: one CTA = one 32-row slot x 128 columns, and the per-(slot, column)
// sum / sum of squares of the stored values go to `colstats` ([2][slots][N], the layout gn_colstats_finalize_kernel folds), so the
// consumer needs no statistics pass of its own.  Thread = (row group of 4 rows, column quad); the 8 row groups of a slot are
// folded in a fixed order.  rows % 32 == 0 (the launcher checks that a slot never straddles two samples).
__global__ void __launch_bounds__(256)
splitk_reduce_stats_kernel(const float* __restrict__ ws, long long split_stride, int splits, long long rows, int N,
                           const float* __restrict__ bias, const float* __restrict__ rowvec, long long ldv, long long rows_per_img,
                           const float* __restrict__ residual, long long ldr, float* __restrict__ out, long long ldc,
                           float* __restrict__ colstats, long long colstats_sq) {
    pdl_trigger();
    pdl_wait();
    __shared__ float part[8][32][8];
    const int rg = threadIdx.x >> 5, qi = threadIdx.x & 31;
    const long long slot = blockIdx.x;
    const int n = (blockIdx.y * 32 + qi) << 2;
    const bool col_ok = n < N;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const long long m = slot * 32 + rg * 4 + r;
            float4 acc = __ldcg(reinterpret_cast<const float4*>(ws + m * N + n));
            for (int sp = 1; sp < splits; ++sp) {
                const float4 q = __ldcg(reinterpret_cast<const float4*>(ws + sp * split_stride + m * N + n));
                acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
            if (rowvec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rowvec + (m / rows_per_img) * ldv + n));
                acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            if (residual) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(residual + m * ldr + n));
                acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            *reinterpret_cast<float4*>(out + m * ldc + n) = acc;
            cs.x += acc.x; cs.y += acc.y; cs.z += acc.z; cs.w += acc.w;
            cq.x = fmaf(acc.x, acc.x, cq.x); cq.y = fmaf(acc.y, acc.y, cq.y); cq.z = fmaf(acc.z, acc.z, cq.z); cq.w = fmaf(acc.w, acc.w, cq.w);
        }
    }
    float* pp = part[rg][qi];
    pp[0] = cs.x; pp[1] = cs.y; pp[2] = cs.z; pp[3] = cs.w; pp[4] = cq.x; pp[5] = cq.y; pp[6] = cq.z; pp[7] = cq.w;
    __syncthreads();
    if (rg == 0 && col_ok) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = part[0][qi][j];
#pragma unroll
        for (int g = 1; g < 8; ++g)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] += part[g][qi][j];
        float* dst = colstats + slot * N + n;
        *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
        *reinterpret_cast<float4*>(dst + colstats_sq) = make_float4(t[4], t[5], t[6], t[7]);
    }
}

// ---- launch plan: tile width, kernel variant, pixel-block decomposition, split-K ---------------------
struct TcPlan {
    int bn, use_pair, taps, Kt, kpt, kblocks, tiles_n, m_tiles, split_k;
    int tw, th, tn, tiles_w, tiles_h;
    long long rows_out;          // rows of the output / workspace (M, or NB*OHF*OWF for a remapped conv)
    long long ws_bytes;
};

static int make_plan(const sdb_tc_args* a, TcPlan* pl) {
    SDB_REQUIRE(a && a->M > 0 && a->N > 0 && a->K > 0, "tc_contract: empty problem");
    memset(pl, 0, sizeof(*pl));
    const bool conv = a->taps > 0;
    pl->taps = conv ? a->taps : 1;
    pl->Kt = conv ? a->Cin : a->K;                        // contraction length per tap
    SDB_REQUIRE(pl->Kt % 8 == 0, "tc_contract: K per tap (%d) must be a multiple of 8", pl->Kt);
    SDB_REQUIRE(a->variant >= 0 && a->variant <= 2, "tc_contract: variant %d unknown", a->variant);
    const bool want_pair = a->variant ? a->variant == 2 : pair_kernel_enabled();
    int bn = a->block_n;
    if (bn == 0) {
        if (a->N <= 32) bn = 32;
        else if (a->N <= 64) bn = 64;
        else if (want_pair) {
            // CTA-pair kernel: widest tile that divides N (bytes pulled from L2 per FLOP fall with BN)
            if (a->N % 256 == 0) bn = 256;
            else if (a->N % 160 == 0) bn = 160;
            else bn = a->N > 128 ? 256 : 128;
        }
        else if (a->N % 160 == 0 && a->N % 128 != 0) bn = 160;
        else if (a->N % 256 == 0 && (long long)a->M * a->N >= 148LL * 2 * 128 * 256) bn = 256;
        else bn = 128;
    }
    SDB_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256, "tc_contract: block_n %d unsupported", bn);
    if (a->geglu) SDB_REQUIRE(a->N % bn == 0 && (bn / 2) % 32 == 0 && !conv, "tc_contract: geglu needs N %% block_n == 0, block_n %% 64 == 0");
    pl->bn = bn;
    pl->use_pair = want_pair && bn >= 128;
    pl->kpt = ceil_div(pl->Kt, TC_BK);
    pl->kblocks = pl->taps * pl->kpt;
    pl->tiles_n = ceil_div(a->N, bn);
    if (conv) {
        SDB_REQUIRE(a->kw > 0 && pl->taps % a->kw == 0 && a->stride >= 1 && a->stride <= 2, "tc_contract: bad conv geometry");
        SDB_REQUIRE(a->NB > 0 && a->IH > 0 && a->IW > 0 && a->OH > 0 && a->OW > 0, "tc_contract: bad conv dims");
        SDB_REQUIRE(a->M == a->NB * a->OH * a->OW, "tc_contract: M != NB*OH*OW");
        SDB_REQUIRE(a->cout_pad >= a->N, "tc_contract: cout_pad < N");
        // The pixel-block shape is chosen for a NOMINAL batch of 8 images, not the actual one: it decides whether a tile holds
        // one image (then the epilogue can emit GroupNorm column statistics) or several, and in bf16 mode two statistics
        // paths that differ by 1e-7 decorrelate to the bf16 noise level within a few layers (measured 4.4e-3 at an 80x80
        // latent) — sample i of a batch must come out bit-identical to the same sample run alone.
        pick_tile(a->OW, a->OH, 8, a->stride, &pl->tw, &pl->th, &pl->tn);
        pl->tiles_w = ceil_div(a->OW, pl->tw); pl->tiles_h = ceil_div(a->OH, pl->th);
        pl->m_tiles = pl->tiles_w * pl->tiles_h * ceil_div(a->NB, pl->tn);
        const long long OHF = a->OHF > 0 ? a->OHF : a->OH, OWF = a->OWF > 0 ? a->OWF : a->OW;
        pl->rows_out = (long long)a->NB * OHF * OWF;
    } else {
        pl->m_tiles = ceil_div(a->M, TC_BM);
        pl->rows_out = a->M;
    }
    // split-K: only when the tile grid leaves most of the machine idle and K is long
    int sk = a->split_k;
    const bool splittable = !a->geglu && !a->col_group && a->N % 4 == 0;
    if (sk == 0) {
        sk = 1;
        if (splittable && a->ws) {
            // The choice must not depend on how many independent samples share the call (sample i of a batch has to
            // come out bit-identical to the same sample run alone), so the tile count is taken at a nominal batch
            // of 8 samples whenever the caller says how many rows one sample owns.
            const long long rpi = conv ? (long long)a->OH * a->OW : (long long)a->rows_per_item;
            const long long m_ref = rpi > 0 ? 8 * rpi : (long long)a->M;
            const long long mt_ref = (m_ref + TC_BM - 1) / TC_BM;
            const long long units = (pl->use_pair ? (mt_ref + 1) / 2 : mt_ref) * pl->tiles_n;
            const long long slots = pl->use_pair ? sm_count_cached() / 2 : 2LL * sm_count_cached();
            if (units * 2 <= slots && pl->kblocks >= 16) {
                long long want = slots / units;
                if (want > pl->kblocks / 8) want = pl->kblocks / 8;
                if (want > 16) want = 16;
                if (want > 1) sk = (int)want;
            }
        }
    }
    if (sk < 1) sk = 1;
    if (sk > pl->kblocks) sk = pl->kblocks;
    SDB_REQUIRE(sk == 1 || splittable, "tc_contract: split_k needs a plain (no GEGLU / column remap) output with N %% 4 == 0");
    pl->split_k = sk;
    pl->ws_bytes = sk > 1 ? (long long)sk * pl->rows_out * a->N * 4 : 0;
    return SDB_OK;
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_tc_set_pair_kernel(int enable) {
    int prev = pair_kernel_enabled() ? 1 : 0;
    g_pair_kernel = enable ? 1 : 0;
    return prev;
}

extern "C" int sdb_tc_set_tail_split(int enable) {
    int prev = tail_split_enabled() ? 1 : 0;
    g_tail_split = enable ? 1 : 0;
    return prev;
}

extern "C" int sdb_tc_set_tma_epilogue(int enable) {
    int prev = tma_epilogue_enabled() ? 1 : 0;
    g_tma_epilogue = enable ? 1 : 0;
    return prev;
}

extern "C" long long sdb_tc_workspace_bytes(const sdb_tc_args* a) {
    if (!a) return -1;
    // what the automatic split choice would need if a workspace were supplied
    sdb_tc_args t = *a;
    if (t.split_k == 0 && !t.ws) t.ws = reinterpret_cast<void*>(16);
    TcPlan pl;
    if (make_plan(&t, &pl) != SDB_OK) return -1;
    return pl.ws_bytes;
}

// Column-statistics layout of a plan.  Without split-K the epilogue writes them: 4 slots (32-row lane quarters) per 128-row
// m-tile; the slots of a sample must be consecutive: one sample per tile (tn == 1), or tn whole images of >= 32 pixels in one
// tile; a sample then owns 4 * tiles_w * tiles_h / tn consecutive slots.  With split-K the fixed-order reduction kernel writes them: one slot per 32
// consecutive output rows, OH * OW / 32 per sample.  Either way only an fp32 output that stores its final values qualifies.
static bool colstats_supported(const sdb_tc_args* a, const TcPlan& pl) {
    // (an output remap — sub-pixel phase of an upsampling conv — is fine without split-K: the caller gives every phase its own slot region)
    if (!(a->taps > 0 && !a->geglu && !a->col_group && a->out_dtype == SDB_F32 && a->N % 4 == 0 && a->ldc % 4 == 0)) return false;
    if (pl.split_k > 1) return a->out_sh <= 1 && ((long long)a->OH * a->OW) % 32 == 0 && (a->OHF <= 0 || a->OHF == a->OH) && (a->OWF <= 0 || a->OWF == a->OW);
    // tn > 1: the tile holds tn whole images (tiles_w == tiles_h == 1), image i in rows [i*tw*th, (i+1)*tw*th): its slots are
    // consecutive as long as an image covers whole 32-row quarters
    return pl.tn == 1 || (pl.tiles_w * pl.tiles_h == 1 && pl.tw * pl.th >= 32);
}

static void colstats_layout(const sdb_tc_args* a, const TcPlan& pl, long long* slots, long long* slots_per_item) {
    if (pl.split_k > 1) {
        *slots = pl.rows_out / 32;
        *slots_per_item = ((long long)a->OH * a->OW) / 32;
    } else {
        const long long mt = pl.use_pair ? 2LL * ((pl.m_tiles + 1) / 2) : pl.m_tiles;
        *slots = 4 * mt;
        *slots_per_item = 4LL * pl.tiles_w * pl.tiles_h / pl.tn;
    }
}

extern "C" int sdb_tc_colstats_layout(const sdb_tc_args* a, long long* slots, long long* slots_per_item) {
    if (!a || !slots || !slots_per_item) return SDB_ERR_INVALID;
    *slots = 0; *slots_per_item = 0;
    sdb_tc_args t = *a;
    if (t.split_k == 0 && !t.ws) t.ws = reinterpret_cast<void*>(16);
    TcPlan pl;
    int rc = make_plan(&t, &pl);
    if (rc) return rc;
    if (!colstats_supported(a, pl)) return SDB_OK;
    colstats_layout(a, pl, slots, slots_per_item);
    return SDB_OK;
}

extern "C" int sdb_tc_contract(const sdb_tc_args* a, void* stream) {
    SDB_REQUIRE(a && a->A && a->B && a->out, "tc_contract: null pointer");
    TcPlan pl;
    int rc = make_plan(a, &pl);
    if (rc) return rc;
    if (a->colstats) {
        SDB_REQUIRE(colstats_supported(a, pl), "tc_contract: column statistics need an fp32 conv output whose 32-row slots lie inside one sample");
        long long need_slots = 0, spi_ = 0;
        colstats_layout(a, pl, &need_slots, &spi_);
        SDB_REQUIRE(a->colstats_slots >= need_slots, "tc_contract: colstats buffer has %lld slots, needs %lld", a->colstats_slots, need_slots);
        SDB_REQUIRE(((uintptr_t)a->colstats & 15) == 0 && ((uintptr_t)a->out & 15) == 0 && (!a->residual || (((uintptr_t)a->residual & 15) == 0 && a->ldr % 4 == 0)) &&
                    (!a->rowvec || (((uintptr_t)a->rowvec & 15) == 0 && a->ldv % 4 == 0)), "tc_contract: column statistics need 16-byte aligned operands");
    }
    const bool conv = a->taps > 0;
    const int bn = pl.bn;
    const bool use_pair = pl.use_pair;

    TcP p;
    memset(&p, 0, sizeof(p));
    p.out = a->out; p.bias = a->bias; p.rowvec = a->rowvec; p.residual = a->residual;
    p.ldc = a->ldc; p.ldr = a->ldr; p.ldv = a->ldv;
    p.M = a->M; p.N = a->N;
    p.kpt = pl.kpt;
    p.b_const = a->b_const ? 1 : 0;
    p.kblocks = pl.kblocks;
    p.out_bf16 = a->out_dtype == SDB_BF16;
    p.geglu = a->geglu;
    p.col_group = a->col_group; p.col_group_stride = a->col_group_stride;
    p.split_k = pl.split_k;
    if (p.split_k > 1) {
        SDB_REQUIRE(a->ws && a->ws_bytes >= pl.ws_bytes, "tc_contract: split_k=%d needs a %lld-byte workspace (got %lld)",
                    p.split_k, pl.ws_bytes, a->ws ? a->ws_bytes : 0LL);
        SDB_REQUIRE(((uintptr_t)a->ws & 15) == 0, "tc_contract: workspace must be 16-byte aligned");
        p.ws = reinterpret_cast<float*>(a->ws);
        p.ws_split_stride = pl.rows_out * a->N;
    }
    {
        const long long ld_max = a->ldc > a->N ? a->ldc : a->N;
        p.off32 = (pl.rows_out + 1) * ld_max + a->N < (1LL << 32) ? 1 : 0;
    }
    p.colstats = pl.split_k > 1 ? nullptr : a->colstats;      // split-K: the reduction kernel writes them
    p.colstats_sq = a->colstats ? a->colstats_slots * (long long)a->N : 0;
#ifdef SDB_TC_TRACE
    p.trace = g_trace_ptr;
#endif
    p.tiles_n = pl.tiles_n;
    p.conv = conv;
    p.cout_pad = conv ? a->cout_pad : 0;

    CUtensorMap tmA, tmB;
    const int m_tiles = pl.m_tiles;
    const int taps = pl.taps;
    if (conv) {
        p.tw = pl.tw; p.th = pl.th; p.tn = pl.tn;
        p.kw = a->kw; p.stride = a->stride; p.pad_h = a->pad_h; p.pad_w = a->pad_w;
        p.NB = a->NB; p.OH = a->OH; p.OW = a->OW;
        p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
        p.out_sh = a->out_sh > 0 ? a->out_sh : 1; p.out_sw = a->out_sw > 0 ? a->out_sw : 1;
        p.out_oh = a->out_oh; p.out_ow = a->out_ow;
        p.OHF = a->OHF > 0 ? a->OHF : a->OH; p.OWF = a->OWF > 0 ? a->OWF : a->OW;
        long long ldx = a->lda > 0 ? a->lda : a->Cin;
        long long dims[4] = {a->Cin, a->IW, a->IH, a->NB};
        long long strides[3] = {ldx, ldx * a->IW, ldx * a->IW * a->IH};
        int box[4] = {TC_BK, p.tw * a->stride, p.th * a->stride, p.tn};
        int es[4] = {1, a->stride, a->stride, 1};
        rc = make_tmap_bf16(&tmA, a->A, 4, dims, strides, box, es);
        if (rc) return rc;
        long long bdims[2] = {a->Cin, (long long)taps * a->cout_pad};
        long long bstr[1] = {a->ldb > 0 ? a->ldb : a->Cin};
        int bbox[2] = {TC_BK, use_pair ? bn / 2 : bn};
        int bes[2] = {1, 1};
        rc = make_tmap_bf16(&tmB, a->B, 2, bdims, bstr, bbox, bes);
        if (rc) return rc;
    } else {
        long long dims[2] = {a->K, a->M};
        long long strides[1] = {a->lda};
        int box[2] = {TC_BK, TC_BM};
        int es[2] = {1, 1};
        rc = make_tmap_bf16(&tmA, a->A, 2, dims, strides, box, es);
        if (rc) return rc;
        long long bdims[2] = {a->K, a->N};
        long long bstr[1] = {a->ldb};
        int bbox[2] = {TC_BK, use_pair ? bn / 2 : bn};
        rc = make_tmap_bf16(&tmB, a->B, 2, bdims, bstr, bbox, es);
        if (rc) return rc;
    }
    SDB_REQUIRE((long long)m_tiles * p.tiles_n < (1LL << 31), "tc_contract: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tmC, tmR;
    memset(&tmC, 0, sizeof(tmC));
    memset(&tmR, 0, sizeof(tmR));
    if (tma_epilogue_enabled()) {
        // ---- TMA epilogue (epilogue_tma_units): short-K launches whose output layout is a plain [rows, N] / NHWC tensor ----
        const bool partial = p.split_k > 1;
        const bool out16 = !partial && p.out_bf16;
        const int kb_split = pl.kblocks / p.split_k;
        const int ebytes = out16 ? 2 : 4;
        bool ok = !a->col_group && kb_split <= tma_epilogue_max_kblocks();
        void* obase = partial ? (void*)p.ws : a->out;
        const long long ldo = partial ? (long long)a->N : a->ldc;
        ok = ok && ((uintptr_t)obase & 15) == 0 && (ldo * ebytes) % 16 == 0;
        if (conv) ok = ok && p.out_sh == 1 && p.out_sw == 1 && p.out_oh == 0 && p.out_ow == 0 && p.OHF == p.OH && p.OWF == p.OW;
        const bool has_res = a->residual != nullptr && !partial;
        if (has_res) ok = ok && !out16 && ((uintptr_t)a->residual & 15) == 0 && a->ldr % 4 == 0;
        if (a->geglu) ok = ok && out16 && (bn / 2) % 32 == 0;
        if (a->rowvec && conv && !partial)
            ok = ok && !out16 && p.tw * p.th >= 32 && a->N % 4 == 0 && a->ldv % 4 == 0 && ((uintptr_t)a->rowvec & 15) == 0;
        if (ok) {
            int bw = 32, bh = 1, bnn = 1;
            if (conv) { bw = p.tw < 32 ? p.tw : 32; bh = p.th < 32 / bw ? p.th : 32 / bw; bnn = 32 / (bw * bh); }
            const long long S = p.split_k;
            long long dims[5], str[4];
            int box[5] = {32, bw, bh, bnn, 1}, es5[5] = {1, 1, 1, 1, 1};
            if (conv) {
                dims[0] = a->N; dims[1] = p.OW; dims[2] = p.OH; dims[3] = p.NB; dims[4] = S;
                str[0] = ldo; str[1] = ldo * p.OW; str[2] = ldo * p.OW * p.OH; str[3] = partial ? p.ws_split_stride : ldo * p.OW * p.OH * p.NB;
            } else {
                dims[0] = a->N; dims[1] = a->M; dims[2] = 1; dims[3] = 1; dims[4] = S;
                str[0] = ldo; str[1] = ldo * a->M; str[2] = ldo * a->M; str[3] = partial ? p.ws_split_stride : ldo * a->M;
            }
            if (a->geglu) dims[0] = a->N / 2;
            rc = make_tmap(&tmC, obase, ebytes, out16 ? 64 : 128, 5, dims, str, box, es5);
            if (rc == SDB_OK && has_res) {
                const long long ldr = a->ldr;
                dims[4] = 1;
                if (conv) { str[0] = ldr; str[1] = ldr * p.OW; str[2] = ldr * p.OW * p.OH; str[3] = ldr * p.OW * p.OH * p.NB; }
                else { str[0] = ldr; str[1] = ldr * a->M; str[2] = ldr * a->M; str[3] = ldr * a->M; }
                rc = make_tmap(&tmR, a->residual, 4, 128, 5, dims, str, box, es5);
            }
            if (rc == SDB_OK) {
                p.epi_tma = 1;
                p.epi_bw = bw; p.epi_bh = bh;
                // a residual tile is prefetched nbuf - 1 chunks ahead: two ahead while the ring can spare the smem (short K)
                p.epi_nbuf = has_res ? (kb_split <= 24 ? 3 : 2) : 2;
            } else {
                rc = SDB_OK;              // odd strides etc.: the register-store epilogue takes any layout
            }
        }
    }
    if (!p.epi_tma) p.epi_nbuf = 0;
    if (use_pair) {
        switch (bn) {
            case 128: rc = launch_tc_pair<128>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            case 160: rc = launch_tc_pair<160>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            default: rc = launch_tc_pair<256>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
        }
    } else {
        switch (bn) {
            case 32: rc = launch_tc<32>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            case 64: rc = launch_tc<64>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            case 128: rc = launch_tc<128>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            case 160: rc = launch_tc<160>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
            default: rc = launch_tc<256>(tmA, tmB, tmC, tmR, p, m_tiles, st); break;
        }
    }
    if (rc || p.split_k == 1) return rc;
    // remapped output (sub-pixel phase): the reduction visits this conv's own pixels only
    ReduceRemap rm;
    memset(&rm, 0, sizeof(rm));
    long long red_rows = pl.rows_out;
    if (conv && (p.out_sh != 1 || p.out_sw != 1 || p.out_oh != 0 || p.out_ow != 0 || p.OHF != p.OH || p.OWF != p.OW)) {
        rm.on = 1; rm.OH = p.OH; rm.OW = p.OW; rm.sh = p.out_sh; rm.sw = p.out_sw; rm.oh0 = p.out_oh; rm.ow0 = p.out_ow;
        rm.OHF = p.OHF; rm.OWF = p.OWF;
        red_rows = (long long)p.NB * p.OH * p.OW;
    }
    const long long total = red_rows * (a->N / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const long long rows_per_img = conv ? (long long)p.OHF * p.OWF : 1;
    if (a->colstats) {
        SDB_REQUIRE(!p.out_bf16 && ((uintptr_t)a->out & 15) == 0 && a->ldc % 4 == 0 && (!a->bias || ((uintptr_t)a->bias & 15) == 0),
                    "tc_contract: split-K column statistics need 16-byte aligned fp32 operands");
        dim3 g((unsigned)(pl.rows_out / 32), (unsigned)((a->N / 4 + 31) / 32));
        launch_pdl(splitk_reduce_stats_kernel, dim3(g), dim3(256), 0, st, p.ws, p.ws_split_stride, p.split_k, pl.rows_out, a->N, a->bias,
                   conv ? a->rowvec : nullptr, a->ldv, rows_per_img, a->residual, a->ldr, reinterpret_cast<float*>(a->out), a->ldc,
                   a->colstats, p.colstats_sq);
        return check_launch("splitk_reduce_stats_kernel");
    }
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, st, p.ws, p.ws_split_stride, p.split_k, red_rows, a->N, a->bias,
                                                 conv ? a->rowvec : nullptr, a->ldv, rows_per_img, a->residual, a->ldr,
                                                 a->out, a->ldc, p.out_bf16, rm);
    return check_launch("splitk_reduce_kernel");
}
