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

namespace sdb {

using namespace ptx;

struct TcP {
    void* out;
    const float* bias; const float* rowvec; const float* residual;
    long long ldc, ldr, ldv;
    int M, N;
    int kblocks;            // total k-blocks = taps * kpt
    int kpt;                // k-blocks per tap
    int out_bf16, geglu;
    int col_group, col_group_stride;
    int split_k;
    int tiles_n;
    // conv
    int conv;
    int kw, stride, pad_h, pad_w;
    int NB, OH, OW;
    int tw, th, tn, tiles_w, tiles_h;
    int cout_pad;
    int out_sh, out_sw, out_oh, out_ow, OHF, OWF;
};

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

__device__ __forceinline__ void store_row_chunk(const TcP& p, long long row_off, int n, const float* v, int cnt, bool atomic) {
    // v[0..cnt) are consecutive output columns n..n+cnt-1 (cnt multiple of 4 unless at the N edge)
    if (p.out_bf16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out);
        for (int j = 0; j < cnt; j += 8) {
            int nn = n + j;
            if (nn >= p.N) break;
            int dn = p.col_group ? (nn / p.col_group) * p.col_group_stride + nn % p.col_group : nn;
            if (nn + 8 <= p.N && j + 8 <= cnt && ((row_off + dn) & 7) == 0) {
                uint4 u;
                u.x = pack_bf16x2(v[j], v[j + 1]); u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                u.z = pack_bf16x2(v[j + 4], v[j + 5]); u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(o + row_off + dn) = u;
            } else {
                for (int t = 0; t < 8 && j + t < cnt && nn + t < p.N; ++t) {
                    int n2 = nn + t;
                    int d2 = p.col_group ? (n2 / p.col_group) * p.col_group_stride + n2 % p.col_group : n2;
                    o[row_off + d2] = __float2bfloat16_rn(v[j + t]);
                }
            }
        }
    } else {
        float* o = reinterpret_cast<float*>(p.out);
        for (int j = 0; j < cnt; j += 4) {
            int nn = n + j;
            if (nn >= p.N) break;
            if (atomic) {
                for (int t = 0; t < 4 && j + t < cnt && nn + t < p.N; ++t) atomicAdd(o + row_off + nn + t, v[j + t]);
            } else if (nn + 4 <= p.N && j + 4 <= cnt && ((row_off + nn) & 3) == 0) {
                *reinterpret_cast<float4*>(o + row_off + nn) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
                for (int t = 0; t < 4 && j + t < cnt && nn + t < p.N; ++t) o[row_off + nn + t] = v[j + t];
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(192, 2)
tc_contract_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcP p) {
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
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    const int tile = blockIdx.x;
    const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
    const int n0 = nt * BN;
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

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                const int tap = kb / p.kpt, cs = kb - tap * p.kpt;
                if (p.conv) {
                    const int r = tap / p.kw, sx = tap - r * p.kw;
                    tma_load_4d(sA + s * TC_A_BYTES, &tmA, &full_bar[s], cs * TC_BK,
                                ow0 * p.stride + sx - p.pad_w, oh0 * p.stride + r - p.pad_h, img0);
                } else {
                    tma_load_2d(sA + s * TC_A_BYTES, &tmA, &full_bar[s], cs * TC_BK, m0);
                }
                tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], cs * TC_BK, tap * p.cout_pad + n0);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(BN, false, false);
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
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1) =================
        const int et = threadIdx.x - 64;             // 0..127
        for (int i = et; i < BN; i += 128) {
            int n = n0 + i;
            s_bias[i] = (p.bias && n < p.N && split == 0) ? p.bias[n] : 0.f;
        }
        named_bar_sync(1, 128);

        const int lg = warp & 3;                     // TMEM lane group this warp may access
        const int row = lg * 32 + lane;              // row of the 128-row tile
        bool valid;
        long long pix;                               // output row index (pixel / token)
        int img = 0;
        if (p.conv) {
            int in_ = row / (p.th * p.tw);
            int rem = row - in_ * (p.th * p.tw);
            int ih = rem / p.tw, iw = rem - ih * p.tw;
            img = img0 + in_;
            int oh = oh0 + ih, ow = ow0 + iw;
            valid = (img < p.NB) && (oh < p.OH) && (ow < p.OW);
            pix = ((long long)img * p.OHF + (oh * p.out_sh + p.out_oh)) * p.OWF + (ow * p.out_sw + p.out_ow);
        } else {
            valid = (m0 + row) < p.M;
            pix = m0 + row;
        }
        const long long out_off = pix * p.ldc;
        const float* res = (p.residual && split == 0) ? p.residual + pix * p.ldr : nullptr;
        const float* rv = (p.rowvec && split == 0 && p.conv) ? p.rowvec + (long long)img * p.ldv : nullptr;
        const bool atomic = p.split_k > 1;

        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_d + ((uint32_t)(lg * 32) << 16);

        if (!p.geglu) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(taddr + c0, r);
                tmem_ld_wait();
                if (valid && n0 + c0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
                    if (rv) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n0 + c0 + j < p.N) v[j] += __ldg(rv + n0 + c0 + j);
                    }
                    if (res) {
                        if (n0 + c0 + 32 <= p.N && ((pix * p.ldr + n0 + c0) & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 q = __ldg(reinterpret_cast<const float4*>(res + n0 + c0 + j));
                                v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j) if (n0 + c0 + j < p.N) v[j] += res[n0 + c0 + j];
                        }
                    }
                    store_row_chunk(p, out_off, n0 + c0, v, 32, atomic);
                }
            }
        } else {
            // GEGLU: tile columns [0, BN/2) are the value half, [BN/2, BN) the gate half of the same
            // output columns nt*BN/2 + j  (weights packed that way by the host).
            constexpr int HALF = BN / 2;
            const int on0 = nt * HALF;
            const int NO = p.N / 2;                  // output columns
#pragma unroll 1
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                uint32_t ra[16], rg[16];
                tmem_ld_x16(taddr + c0, ra);
                tmem_ld_x16(taddr + HALF + c0, rg);
                tmem_ld_wait();
                if (valid && on0 + c0 < NO) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(ra[j]) + s_bias[c0 + j];
                        float g = __uint_as_float(rg[j]) + s_bias[HALF + c0 + j];
                        v[j] = a * gelu_erf(g);
                    }
                    TcP q = p;
                    q.N = NO;
                    store_row_chunk(q, out_off, on0 + c0, v, 16, false);
                }
            }
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

// rank-R bf16 tensor map, 128B swizzle, zero OOB fill. dims/strides innermost first; strides in
// elements for dims 1..R-1.
int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
                   const int* box, const int* estr) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable"); return SDB_ERR_NOTMA; }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bdim[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)estr[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = (cuuint64_t)strides_elems[i] * 2;
    if (((uintptr_t)base & 15) != 0) { set_last_error("tensor map base %p not 16-byte aligned", base); return SDB_ERR_INVALID; }
    for (int i = 0; i + 1 < rank; ++i)
        if (gstr[i] % 16) { set_last_error("tensor map stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i]); return SDB_ERR_INVALID; }
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %lld %lld box %d %d)", (int)r, rank,
                       dims[0], rank > 1 ? dims[1] : 0, box[0], rank > 1 ? box[1] : 0);
        return SDB_ERR_CUDA;
    }
    return SDB_OK;
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

template <int BN>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcP& p, int m_tiles, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_contract_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("tc_contract: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
        attr_set = true;
    }
    dim3 grid((unsigned)(m_tiles * p.tiles_n), (unsigned)p.split_k);
    tc_contract_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p);
    return check_launch("tc_contract_kernel");
}

}  // namespace sdb

using namespace sdb;

extern "C" int sdb_tc_contract(const sdb_tc_args* a, void* stream) {
    SDB_REQUIRE(a && a->A && a->B && a->out, "tc_contract: null pointer");
    SDB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "tc_contract: empty problem");
    const bool conv = a->taps > 0;
    const int taps = conv ? a->taps : 1;
    const int Kt = conv ? a->Cin : a->K;                 // contraction length per tap
    SDB_REQUIRE(Kt % 8 == 0, "tc_contract: K per tap (%d) must be a multiple of 8", Kt);

    int bn = a->block_n;
    if (bn == 0) {
        if (a->N <= 32) bn = 32;
        else if (a->N <= 64) bn = 64;
        else if (a->N % 160 == 0 && a->N % 128 != 0) bn = 160;
        else if (a->N % 256 == 0 && (long long)a->M * a->N >= 148LL * 2 * 128 * 256) bn = 256;
        else bn = 128;
    }
    SDB_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 256, "tc_contract: block_n %d unsupported", bn);
    if (a->geglu) SDB_REQUIRE(a->N % bn == 0 && (bn / 2) % 16 == 0 && !conv, "tc_contract: geglu needs N %% block_n == 0");

    TcP p;
    memset(&p, 0, sizeof(p));
    p.out = a->out; p.bias = a->bias; p.rowvec = a->rowvec; p.residual = a->residual;
    p.ldc = a->ldc; p.ldr = a->ldr; p.ldv = a->ldv;
    p.M = a->M; p.N = a->N;
    p.kpt = ceil_div(Kt, TC_BK);
    p.kblocks = taps * p.kpt;
    p.out_bf16 = a->out_dtype == SDB_BF16;
    p.geglu = a->geglu;
    p.col_group = a->col_group; p.col_group_stride = a->col_group_stride;
    p.split_k = a->split_k > 1 ? a->split_k : 1;
    if (p.split_k > p.kblocks) p.split_k = p.kblocks;
    SDB_REQUIRE(p.split_k == 1 || (!p.out_bf16 && !a->geglu), "tc_contract: split_k needs fp32 plain output");
    p.tiles_n = ceil_div(a->N, bn);
    p.conv = conv;
    p.cout_pad = conv ? a->cout_pad : 0;

    CUtensorMap tmA, tmB;
    int m_tiles;
    int rc;
    if (conv) {
        SDB_REQUIRE(a->kw > 0 && taps % a->kw == 0 && a->stride >= 1 && a->stride <= 2, "tc_contract: bad conv geometry");
        SDB_REQUIRE(a->NB > 0 && a->IH > 0 && a->IW > 0 && a->OH > 0 && a->OW > 0, "tc_contract: bad conv dims");
        SDB_REQUIRE(a->M == a->NB * a->OH * a->OW, "tc_contract: M != NB*OH*OW");
        SDB_REQUIRE(a->cout_pad >= a->N, "tc_contract: cout_pad < N");
        pick_tile(a->OW, a->OH, a->NB, a->stride, &p.tw, &p.th, &p.tn);
        p.kw = a->kw; p.stride = a->stride; p.pad_h = a->pad_h; p.pad_w = a->pad_w;
        p.NB = a->NB; p.OH = a->OH; p.OW = a->OW;
        p.tiles_w = ceil_div(a->OW, p.tw); p.tiles_h = ceil_div(a->OH, p.th);
        m_tiles = p.tiles_w * p.tiles_h * ceil_div(a->NB, p.tn);
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
        int bbox[2] = {TC_BK, bn};
        int bes[2] = {1, 1};
        rc = make_tmap_bf16(&tmB, a->B, 2, bdims, bstr, bbox, bes);
        if (rc) return rc;
    } else {
        m_tiles = ceil_div(a->M, TC_BM);
        long long dims[2] = {a->K, a->M};
        long long strides[1] = {a->lda};
        int box[2] = {TC_BK, TC_BM};
        int es[2] = {1, 1};
        rc = make_tmap_bf16(&tmA, a->A, 2, dims, strides, box, es);
        if (rc) return rc;
        long long bdims[2] = {a->K, a->N};
        long long bstr[1] = {a->ldb};
        int bbox[2] = {TC_BK, bn};
        rc = make_tmap_bf16(&tmB, a->B, 2, bdims, bstr, bbox, es);
        if (rc) return rc;
    }
    SDB_REQUIRE((long long)m_tiles * p.tiles_n < (1LL << 31), "tc_contract: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 32: return launch_tc<32>(tmA, tmB, p, m_tiles, st);
        case 64: return launch_tc<64>(tmA, tmB, p, m_tiles, st);
        case 128: return launch_tc<128>(tmA, tmB, p, m_tiles, st);
        case 160: return launch_tc<160>(tmA, tmB, p, m_tiles, st);
        default: return launch_tc<256>(tmA, tmB, p, m_tiles, st);
    }
}
