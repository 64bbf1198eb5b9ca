// sdb200 — bandwidth-bound normalisation kernels (NHWC / row-major, fp32 residual stream in,
// bf16 tensor-core operand or fp32 out).
//
//   GroupNorm(32)+SiLU : reference openai_model/utils.py:15-22 + model.py:178-181,202-205,528-531;
//                        Normalize(eps 1e-6) openai_model/attention.py:10-11, ldm/.../model.py:40-41
//   LayerNorm          : reference openai_model/attention.py:216-218
//
// Roofline: HBM. Algorithmic bytes/element = 4 (fp32 read) + 2 (bf16 write) [bf16 mode] or 4+4
// [fp32 mode]; the statistics pass re-reads the tensor (L2-resident for the UNet's <=126 MB maps).
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace sdb {

// ---- GroupNorm geometry shared by ws sizing and launches ----------------------------------------
struct GnGeom {
    int V;        // float4 vectors per row (C/4)
    int R;        // rows processed concurrently by one CTA
    int threads;  // V*R
    int chunks;   // CTAs along HW per sample
    int rows_per_chunk;
};

static GnGeom gn_geom(int N, int HW, int C) {
    GnGeom g;
    g.V = C / 4;
    g.R = 512 / g.V;
    if (g.R < 1) g.R = 1;
    if (g.R > 32) g.R = 32;
    if (g.R > HW) g.R = HW;
    g.threads = g.V * g.R;
    // The split depends on (HW, C) only, never on N: the summation order of a sample's statistics is
    // then independent of the batch it is in (sample i of a batch == the same sample run alone, bit for bit).
    (void)N;
    int by_rows = ceil_div(HW, g.R);                 // at least one row-slot per CTA
    int want = 148;                                  // one CTA per SM per sample
    int cap = ceil_div(HW, g.R * 8);                 // <= 8 rows per thread
    int chunks = want > cap ? want : cap;
    if (chunks > by_rows) chunks = by_rows;
    if (chunks < 1) chunks = 1;
    g.chunks = chunks;
    g.rows_per_chunk = ceil_div(HW, chunks);
    return g;
}

// SiLU for the bf16 tensor-core path: ex2.approx + rcp.approx (2 MUFU + 3 FMA-pipe instructions; the IEEE division of
// x / (1 + e^-x) costs ~10 and made the apply pass instruction-bound)
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// Statistics from the producer: the tcgen05 conv that wrote the tensor also wrote, per 32-row slot and channel, the sum
// and sum of squares of what it stored (sdb_tc_args.colstats).  One warp per (sample, group) folds its slots x channels
// in fp64 in a fixed order -> (mean, rstd); the tensor itself is not read.  Two sources = the channel concat.
// layout of one source's statistics: cs = fp32 [2][slots][C]; a sample owns `spi` consecutive slots in each of `regions`
// regions that start `rstride` slots apart (regions > 1: the sub-pixel phases of an upsampling conv, one region each)
struct CsSrc {
    const float* cs;
    int C;
    long long slots, spi, rstride;
    int regions;
};

constexpr int GNF_THREADS = 256, GNF_UNROLL = 8;

__global__ void __launch_bounds__(GNF_THREADS)
gn_colstats_finalize_kernel(const CsSrc s0, const CsSrc s1, int groups, double count, float eps, float2* __restrict__ stats) {
    pdl_trigger();
    pdl_wait();
    __shared__ double redS[GNF_THREADS], redQ[GNF_THREADS];
    const int w = blockIdx.x;                       // (sample, group)
    const int n = w / groups, g = w - n * groups;
    const int C = s0.C + s1.C, cpg = C / groups;
    // channels of this group inside each source
    const int c_lo = g * cpg, c_hi = c_lo + cpg;
    const int a_lo = c_lo < s0.C ? c_lo : s0.C, a_hi = c_hi < s0.C ? c_hi : s0.C;              // [a_lo, a_hi) in source 0
    const int b_lo = (c_lo > s0.C ? c_lo : s0.C) - s0.C, b_hi = (c_hi > s0.C ? c_hi : s0.C) - s0.C;   // [b_lo, b_hi) in source 1
    const int na = a_hi - a_lo, nb = b_hi - b_lo;
    // item = (source, region, slot, channel), 32-bit indices (the host checks the counts); GNF_UNROLL items (2 loads each) in flight per
    // thread, summed in index order
    const uint32_t items0 = (uint32_t)(s0.regions * s0.spi * na);
    const uint32_t items1 = s1.cs ? (uint32_t)(s1.regions * s1.spi * nb) : 0u;
    const uint32_t items = items0 + items1;
    const uint32_t spi0 = (uint32_t)s0.spi, spi1 = (uint32_t)s1.spi;
    double S = 0.0, Q = 0.0;
    for (uint32_t i0 = threadIdx.x; i0 < items; i0 += GNF_UNROLL * GNF_THREADS) {
        float a[GNF_UNROLL], q[GNF_UNROLL];
#pragma unroll
        for (int u = 0; u < GNF_UNROLL; ++u) {
            const uint32_t i = i0 + u * GNF_THREADS;
            a[u] = 0.f; q[u] = 0.f;
            if (i < items) {
                const bool first = i < items0;
                const CsSrc& s = first ? s0 : s1;
                const uint32_t k = first ? i : i - items0;
                const uint32_t nch = first ? (uint32_t)na : (uint32_t)nb, spi = first ? spi0 : spi1;
                const uint32_t qd = k / nch;
                const int c = (first ? a_lo : b_lo) + (int)(k - qd * nch);
                const uint32_t region = s.regions > 1 ? qd / spi : 0u, sl = qd - region * spi;
                const long long slot = (long long)region * s.rstride + (long long)n * s.spi + sl;
                a[u] = __ldcg(s.cs + slot * s.C + c);
                q[u] = __ldcg(s.cs + (s.slots + slot) * s.C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < GNF_UNROLL; ++u) { S += (double)a[u]; Q += (double)q[u]; }
    }
    redS[threadIdx.x] = S; redQ[threadIdx.x] = Q;
    __syncthreads();
    for (int o = GNF_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { redS[threadIdx.x] += redS[threadIdx.x + o]; redQ[threadIdx.x] += redQ[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double mean = redS[0] / count;
        double var = redQ[0] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[w] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}

// Pass 1: per-(sample, chunk, group) partial sum / sum of squares, then the LAST CTA of a sample to finish
// (atomic ticket on `counters[n]`, which it leaves at zero again) combines the chunk partials of that
// sample in a fixed order into (mean, rstd) per group: no separate finalize launch, bit-reproducible.
// Thread t owns vector column v = t % V for rows rr, rr+R, ...: per-channel fp32 partials in
// registers, reduced over R in smem, then per-group in fp64.
__global__ void gn_stats_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                                int HW, int groups, int V, int R, int rows_per_chunk,
                                double* __restrict__ partial, int* __restrict__ counters, double count, float eps,
                                float2* __restrict__ stats) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float red[];   // [2][R][C] floats; reused as [slots][groups] double2 by the finalizing CTA
    __shared__ int s_last;
    const int C = C0 + C1;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % V, rr = threadIdx.x / V;
    const int c = v * 4;
    const float* src;
    long long ld;
    int cc;
    if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
    const int row0 = chunk * rows_per_chunk;
    int row1 = row0 + rows_per_chunk;
    if (row1 > HW) row1 = HW;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    const float* p = src + ((long long)n * HW) * ld + cc;
#pragma unroll 8
    for (int row = row0 + rr; row < row1; row += R) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p + (long long)row * ld));
        s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
        q0 += a.x * a.x; q1 += a.y * a.y; q2 += a.z * a.z; q3 += a.w * a.w;
    }
    float* rs = red + (long long)rr * C + c;
    float* rq = red + (long long)R * C + (long long)rr * C + c;
    rs[0] = s0; rs[1] = s1; rs[2] = s2; rs[3] = s3;
    rq[0] = q0; rq[1] = q1; rq[2] = q2; rq[3] = q3;
    __syncthreads();
    const int cpg = C / groups;
    const int chunks = gridDim.x;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double S = 0.0, Q = 0.0;
        for (int r = 0; r < R; ++r) {
            const float* a = red + (long long)r * C + g * cpg;
            const float* b = red + (long long)R * C + (long long)r * C + g * cpg;
            for (int j = 0; j < cpg; ++j) { S += (double)a[j]; Q += (double)b[j]; }
        }
        double* o = partial + (((long long)n * chunks + chunk) * groups + g) * 2;
        o[0] = S; o[1] = Q;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&counters[n], 1) == chunks - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- finalize sample n: slot-strided fp64 sums over the chunks, then a fixed-order sum over the slots ----
    double2* dred = reinterpret_cast<double2*>(red);
    int nslots = (int)blockDim.x / groups;
    if (nslots < 1) nslots = 1;
    if (nslots > chunks) nslots = chunks;
    for (int item = threadIdx.x; item < groups * nslots; item += blockDim.x) {
        const int g = item % groups, slot = item / groups;
        double S = 0.0, Q = 0.0;
        for (int ch = slot; ch < chunks; ch += nslots) {
            const double* q = partial + (((long long)n * chunks + ch) * groups + g) * 2;
            S += __ldcg(q); Q += __ldcg(q + 1);
        }
        dred[item] = make_double2(S, Q);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double S = 0.0, Q = 0.0;
        for (int slot = 0; slot < nslots; ++slot) { double2 t = dred[slot * groups + g]; S += t.x; Q += t.y; }
        double mean = S / count;
        double var = Q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        double rstd = 1.0 / sqrt(var + (double)eps);
        stats[n * groups + g] = make_float2((float)mean, (float)rstd);
    }
    if (threadIdx.x == 0) counters[n] = 0;
}

// Pass 2: normalise + affine (+ SiLU), emit bf16 (tensor-core operand) or fp32; optionally also the raw
// (un-normalised) bf16 copy of the concatenated input, which is the operand of a ResBlock's 1x1 skip conv.
// One wave of CTAs (2 per SM); a thread streams its rows in blocks of 4 with the next block's loads issued before
// the current block is processed, so every SM keeps ~100 KB of loads in flight from the first to the last row.
constexpr int GNA_UB = 4;

// FUSED: the (mean, rstd) of the sample's groups are folded from the producer's column statistics in the prologue of every
// CTA (warp per group, fp64, fixed order) instead of by a separate gn_colstats_finalize launch — for small maps, where the
// statistics of one sample are a few tens of KB and the finalize launch costs more than the whole normalisation.
constexpr int GNA_MAX_GROUPS = 64;

template <bool OUT_BF16, bool EXACT, bool RAW, bool FUSED>
__global__ void __launch_bounds__(1024, 1)      // <= 64 registers: two ~480-thread CTAs per SM; C up to 4096 in one CTA row
gn_apply_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                int HW, int groups, int V, int R, int rows_per_chunk,
                const float2* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, long long gb_stride, int act, void* __restrict__ out,
                __nv_bfloat16* __restrict__ raw_out, const CsSrc s0, const CsSrc s1, double count, float eps) {
    pdl_trigger();
    __shared__ float2 s_stats[FUSED ? GNA_MAX_GROUPS : 1];
    const int C = C0 + C1;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int v = threadIdx.x % V, rr = threadIdx.x / V;
    const int c = v * 4;
    const float* src;
    long long ld;
    int cc;
    if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
    const int cpg = C / groups;
    const int row0 = chunk * rows_per_chunk;
    int row1 = row0 + rows_per_chunk;
    if (row1 > HW) row1 = HW;
    const float* p = src + ((long long)n * HW) * ld + cc;
    const long long rstep = (long long)R * ld;
    constexpr int UB = GNA_UB;
    pdl_wait();
    float4 cur[UB], nxt[UB];
    int rb = row0 + rr;
#pragma unroll
    for (int i = 0; i < UB; ++i) {
        cur[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rb + i * R < row1) cur[i] = ld_stream_f4(p + (long long)rb * ld + i * rstep);
    }
    if (FUSED) {
        // (1) thread per channel: sum of the sample's slots (independent loads, four slots in flight), fp64, slot order;
        // (2) warp per group: fold its channels from shared memory in channel order.
        extern __shared__ double2 s_ch[];                  // [C] (sum, sum of squares) per channel
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            const bool first = ch < s0.C;
            const CsSrc& sr = first ? s0 : s1;
            const int cch = first ? ch : ch - s0.C;
            double S = 0.0, Q = 0.0;
            for (int region = 0; region < sr.regions; ++region) {
                const long long base = (long long)region * sr.rstride + (long long)n * sr.spi;
                for (long long sl0 = 0; sl0 < sr.spi; sl0 += 4) {
                    float a[4], q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        a[u] = 0.f; q[u] = 0.f;
                        if (sl0 + u < sr.spi) {
                            a[u] = __ldcg(sr.cs + (base + sl0 + u) * sr.C + cch);
                            q[u] = __ldcg(sr.cs + (sr.slots + base + sl0 + u) * sr.C + cch);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { S += (double)a[u]; Q += (double)q[u]; }
                }
            }
            s_ch[ch] = make_double2(S, Q);
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
        for (int g = warp; g < groups; g += nwarps) {
            double S = 0.0, Q = 0.0;
            for (int j = lane; j < cpg; j += 32) { const double2 v = s_ch[g * cpg + j]; S += v.x; Q += v.y; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                S += __shfl_xor_sync(0xffffffffu, S, o);
                Q += __shfl_xor_sync(0xffffffffu, Q, o);
            }
            if (lane == 0) {
                const double mean = S / count;
                double var = Q / count - mean * mean;
                if (var < 0.0) var = 0.0;
                s_stats[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
            }
        }
        __syncthreads();
    }
    float sc[4], sh[4];
    {
        // gb_stride != 0: per-sample affine rows (the scale-shift ResBlock folds (1 + scale), shift into gamma / beta)
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + (long long)n * gb_stride + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + (long long)n * gb_stride + c));
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 st = FUSED ? s_stats[(c + j) / cpg] : __ldcg(stats + n * groups + (c + j) / cpg);
            sc[j] = st.y * gg[j];
            sh[j] = bb[j] - st.x * st.y * gg[j];
        }
    }
    const long long obase = ((long long)n * HW) * C + c;
    for (; rb < row1; rb += UB * R) {
        const int rn = rb + UB * R;
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rn + i * R < row1) nxt[i] = ld_stream_f4(p + (long long)rn * ld + i * rstep);
        }
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            const int row = rb + i * R;
            if (row < row1) {
                float y0 = fmaf(cur[i].x, sc[0], sh[0]), y1 = fmaf(cur[i].y, sc[1], sh[1]);
                float y2 = fmaf(cur[i].z, sc[2], sh[2]), y3 = fmaf(cur[i].w, sc[3], sh[3]);
                if (act == 1) {
                    if (EXACT) { y0 = silu_exact(y0); y1 = silu_exact(y1); y2 = silu_exact(y2); y3 = silu_exact(y3); }
                    else       { y0 = silu_fast(y0);  y1 = silu_fast(y1);  y2 = silu_fast(y2);  y3 = silu_fast(y3); }
                }
                const long long o = obase + (long long)row * C;
                if (OUT_BF16) st_stream_u2(reinterpret_cast<__nv_bfloat16*>(out) + o, pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
                else st_stream_f4(reinterpret_cast<float*>(out) + o, make_float4(y0, y1, y2, y3));
                if (RAW) st_stream_u2(raw_out + o, pack_bf16x2(cur[i].x, cur[i].y), pack_bf16x2(cur[i].z, cur[i].w));
            }
        }
#pragma unroll
        for (int i = 0; i < UB; ++i) cur[i] = nxt[i];
    }
}

// apply-pass geometry: elementwise given the statistics, so (unlike the statistics pass) it may depend on the batch:
// ~2 CTAs per SM in one wave
static GnGeom gn_apply_geom(int N, int HW, int C) {
    GnGeom g;
    g.V = C / 4;
    g.R = 512 / g.V;
    if (g.R < 1) g.R = 1;
    if (g.R > 32) g.R = 32;
    if (g.R > HW) g.R = HW;
    g.threads = g.V * g.R;
    int want = ceil_div(2 * 148, N);
    int by_rows = ceil_div(HW, g.R);
    if (want > by_rows) want = by_rows;
    if (want < 1) want = 1;
    g.rows_per_chunk = ceil_div(HW, want);
    g.chunks = ceil_div(HW, g.rows_per_chunk);
    return g;
}

// ---- GroupNorm as ONE kernel: a thread-block cluster per sample ------------------------------------
// The CTAs of a cluster (8 or 16, one per SM) split the rows of one sample.  Pass 1 accumulates per-channel
// sums (fp32 over blocks of <= 8 rows, fp64 across blocks), folds them to per-group partials and all-gathers
// those through distributed shared memory (every CTA stores its 32 x (S, Q) into every peer's smem, then one
// cluster barrier); each CTA then sums the ranks' partials in rank order -> (mean, rstd): deterministic, no
// atomics, no second launch, and a sample's bits do not depend on the batch it is in.  Pass 2 re-reads the
// CTA's slab (L2-resident: it was read a few microseconds earlier), normalises, applies SiLU and writes the
// bf16 / fp32 result (and optionally the raw bf16 copy).  Traffic: 4 B/element from HBM + 4 B/element from
// L2 + 2 B/element written (bf16 out).
constexpr int GNC_MAX_CS = 16;

template <bool OUT_BF16, bool EXACT, bool RAW>
__global__ void __launch_bounds__(1024, 1)
gn_cluster_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                  int HW, int groups, int V, int R, int rows_per_cta, int cs, float eps,
                  const float* __restrict__ gamma, const float* __restrict__ beta, long long gb_stride, int act,
                  void* __restrict__ out, __nv_bfloat16* __restrict__ raw_out) {
    extern __shared__ double gsm[];
    pdl_trigger();
    // phase 0 of the cluster barrier: "this CTA is running" (a peer's shared memory may only be written once it is)
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    const int C = C0 + C1;
    double* redS = gsm;                                   // [R][C]
    double* redQ = gsm + (size_t)R * C;                   // [R][C]
    double2* gpart = reinterpret_cast<double2*>(redQ + (size_t)R * C);          // [cs][groups] (written by every rank)
    double2* gquart = gpart + (size_t)GNC_MAX_CS * groups;                      // [4][groups]
    float2* gstat = reinterpret_cast<float2*>(gquart + 4 * (size_t)groups);     // [groups] (mean, rstd)
    const uint32_t rank = ptx::cluster_ctarank();
    const int n = blockIdx.y;
    const int v = threadIdx.x % V, rr = threadIdx.x / V;
    const int c = v * 4;
    const float* src;
    long long ld;
    int cc;
    if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
    const int cpg = C / groups;
    const int row0 = (int)rank * rows_per_cta;
    int row1 = row0 + rows_per_cta;
    if (row1 > HW) row1 = HW;
    const float* p = src + ((long long)n * HW) * ld + cc;
    const long long rstep = (long long)R * ld;
    constexpr int UB = 8;                                 // rows per thread in flight
    pdl_wait();

    // ---- pass 1: statistics ----
    double S0 = 0, S1 = 0, S2 = 0, S3 = 0, Q0 = 0, Q1 = 0, Q2 = 0, Q3 = 0;
    for (int rb = row0 + rr; rb < row1; rb += UB * R) {
        float4 a[UB];
        const float* q = p + (long long)rb * ld;
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rb + i * R < row1) a[i] = __ldg(reinterpret_cast<const float4*>(q + i * rstep));
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            s0 += a[i].x; s1 += a[i].y; s2 += a[i].z; s3 += a[i].w;
            q0 = fmaf(a[i].x, a[i].x, q0); q1 = fmaf(a[i].y, a[i].y, q1);
            q2 = fmaf(a[i].z, a[i].z, q2); q3 = fmaf(a[i].w, a[i].w, q3);
        }
        S0 += s0; S1 += s1; S2 += s2; S3 += s3;
        Q0 += q0; Q1 += q1; Q2 += q2; Q3 += q3;
    }
    {
        double* ps = redS + (size_t)rr * C + c;
        double* pq = redQ + (size_t)rr * C + c;
        ps[0] = S0; ps[1] = S1; ps[2] = S2; ps[3] = S3;
        pq[0] = Q0; pq[1] = Q1; pq[2] = Q2; pq[3] = Q3;
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {             // fold the R row-slots of every channel (fixed order)
        double S = 0.0, Q = 0.0;
        for (int r = 0; r < R; ++r) { S += redS[(size_t)r * C + ch]; Q += redQ[(size_t)r * C + ch]; }
        redS[ch] = S; redQ[ch] = Q;
    }
    __syncthreads();
    for (int item = threadIdx.x; item < 4 * groups; item += blockDim.x) {   // quarter-group sums
        const int g = item % groups, part = item / groups;
        const int lo = g * cpg + (part * cpg) / 4, hi = g * cpg + ((part + 1) * cpg) / 4;
        double S = 0.0, Q = 0.0;
        for (int ch = lo; ch < hi; ++ch) { S += redS[ch]; Q += redQ[ch]; }
        gquart[part * groups + g] = make_double2(S, Q);
    }
    __syncthreads();
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");       // every CTA of the cluster has started
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double S = 0.0, Q = 0.0;
        for (int part = 0; part < 4; ++part) { double2 t = gquart[part * groups + g]; S += t.x; Q += t.y; }
        const uint32_t local = ptx::smem_u32(&gpart[(size_t)rank * groups + g]);
        for (int k = 0; k < cs; ++k) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(k));
            asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(remote), "d"(S), "d"(Q) : "memory");
        }
    }
    ptx::cluster_sync_all();                                            // release / acquire: every rank's partials have landed
    const double count = (double)HW * cpg;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double S = 0.0, Q = 0.0;
        for (int k = 0; k < cs; ++k) { double2 t = gpart[(size_t)k * groups + g]; S += t.x; Q += t.y; }
        double mean = S / count;
        double var = Q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        gstat[g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
    __syncthreads();

    // ---- pass 2: normalise + affine (+ SiLU) ----
    float sc[4], sh[4];
    {
        // gb_stride != 0: per-sample affine rows (the scale-shift ResBlock folds (1 + scale), shift into gamma / beta)
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + (long long)n * gb_stride + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + (long long)n * gb_stride + c));
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 st = gstat[(c + j) / cpg];
            sc[j] = st.y * gg[j];
            sh[j] = bb[j] - st.x * st.y * gg[j];
        }
    }
    const long long obase = ((long long)n * HW) * C + c;
    for (int rb = row0 + rr; rb < row1; rb += UB * R) {
        float4 a[UB];
        const float* q = p + (long long)rb * ld;
#pragma unroll
        for (int i = 0; i < UB; ++i)
            if (rb + i * R < row1) a[i] = ld_stream_f4(q + i * rstep);
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            const int row = rb + i * R;
            if (row >= row1) break;
            float y0 = fmaf(a[i].x, sc[0], sh[0]), y1 = fmaf(a[i].y, sc[1], sh[1]);
            float y2 = fmaf(a[i].z, sc[2], sh[2]), y3 = fmaf(a[i].w, sc[3], sh[3]);
            if (act == 1) {
                if (EXACT) { y0 = silu_exact(y0); y1 = silu_exact(y1); y2 = silu_exact(y2); y3 = silu_exact(y3); }
                else       { y0 = silu_fast(y0);  y1 = silu_fast(y1);  y2 = silu_fast(y2);  y3 = silu_fast(y3); }
            }
            const long long o = obase + (long long)row * C;
            if (OUT_BF16) st_stream_u2(reinterpret_cast<__nv_bfloat16*>(out) + o, pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            else st_stream_f4(reinterpret_cast<float*>(out) + o, make_float4(y0, y1, y2, y3));
            if (RAW) st_stream_u2(raw_out + o, pack_bf16x2(a[i].x, a[i].y), pack_bf16x2(a[i].z, a[i].w));
        }
    }
}

// geometry of the cluster kernel: V float4 columns, R row slots (V * R threads <= 1024)
struct GncGeom { int V, R, threads, cs, rows_per_cta; size_t smem; };

static bool gnc_geom(int HW, int C, int groups, int max_cs, GncGeom* g) {
    g->V = C / 4;
    if (g->V < 1 || g->V > 1024 || max_cs < 2) return false;
    g->R = 1024 / g->V;
    if (g->R > 32) g->R = 32;
    if (g->R > HW) g->R = HW;
    g->threads = g->V * g->R;
    int cs = max_cs;
    while (cs > 2 && HW < cs * g->R) cs >>= 1;            // at least one row per thread slot and CTA
    g->cs = cs;
    g->rows_per_cta = ceil_div(HW, cs);
    g->smem = (size_t)2 * g->R * C * sizeof(double) + (size_t)(GNC_MAX_CS + 4) * groups * sizeof(double2) +
              (size_t)groups * sizeof(float2) + 16;
    return g->smem <= 200 * 1024;
}

template <bool BF, bool EX, bool RW>
static int gnc_max_cluster() {
    // largest cluster size (16, 8, ...) the device can co-schedule for this instantiation; 0 = cluster path unusable
    static int cached = -1;
    if (cached >= 0) return cached;
    cached = 0;
    auto k = gn_cluster_kernel<BF, EX, RW>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); }
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) { cudaGetLastError(); return cached; }
    // 16-CTA clusters fit only 7 at a time on a B200 (launch__cluster_max_active), i.e. two waves for 8 samples: the
    // portable size 8 (one wave) is faster at every UNet shape (profiles/r01_gn_cluster_sizes.txt)
    int cs_cap = 8;
    if (const char* e = getenv("SDB200_GN_CS")) { int v = atoi(e); if (v >= 2 && v <= GNC_MAX_CS) cs_cap = v; }   // measurement only
    for (int cs = cs_cap; cs >= 2; cs >>= 1) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cs, 1, 1); cfg.blockDim = dim3(1024, 1, 1); cfg.dynamicSmemBytes = 72 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, k, &cfg) == cudaSuccess && nclusters >= 1) { cached = cs; break; }
        cudaGetLastError();
    }
    return cached;
}

template <bool BF, bool EX, bool RW>
static int launch_gn_cluster(const float* x0, int C0, const float* x1, int C1, int N, int HW, int groups, float eps,
                             const float* gamma, const float* beta, long long gb_stride, int act, void* out, __nv_bfloat16* raw,
                             cudaStream_t st, bool* launched) {
    *launched = false;
    const int max_cs = gnc_max_cluster<BF, EX, RW>();
    GncGeom g;
    if (max_cs < 2 || !gnc_geom(HW, C0 + C1, groups, max_cs, &g)) return SDB_OK;
    launch_pdl_cluster(gn_cluster_kernel<BF, EX, RW>, dim3(g.cs, N), dim3(g.threads), g.smem, st, g.cs,
                       x0, C0, x1, C1, HW, groups, g.V, g.R, g.rows_per_cta, g.cs, eps, gamma, beta, gb_stride, act, out, raw);
    *launched = true;
    return check_launch("gn_cluster_kernel");
}

// SDB200_GN=split forces the two-kernel (statistics + apply) path; measurement only
static bool gn_cluster_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("SDB200_GN");
        on = (e && strcmp(e, "split") == 0) ? 0 : 1;
    }
    return on == 1;
}

// ---- LayerNorm: one warp per row, row held in registers (two-pass mean/variance) ----------------
// One row per warp measured fastest: several rows per warp (more loads in flight per lane, fewer resident warps) and a
// bulk-async shared-memory ring with 8-24 consumer warps were both 20-60 % slower at the UNet shapes — the kernel is
// bound by its two dependent shuffle reductions per row, which only more resident warps hide.

template <bool OUT_BF16, int MAXV>   // MAXV float4 vectors per lane: C <= 128 * MAXV
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int rows, int C, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int V = C >> 2;
    const float* p = x + (long long)row * C;
    float4 r[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i;
        r[i] = make_float4(0.f, 0.f, 0.f, 0.f);            // padding lanes hold zeros
        if (v < V) r[i] = ld_stream_f4(p + 4 * v);
    }
    const float invC = 1.0f / (float)C;
    float mean, rstd;
    if (OUT_BF16) {
        // tensor-core operand: sum and sum of squares reduced together (ONE dependent shuffle chain per row instead of
        // two); E[x^2] - mean^2 in fp32 is far inside the bf16 rounding of the result
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            s += (r[i].x + r[i].y) + (r[i].z + r[i].w);
            q += (r[i].x * r[i].x + r[i].y * r[i].y) + (r[i].z * r[i].z + r[i].w * r[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        mean = s * invC;
        rstd = rsqrtf(fmaxf(q * invC - mean * mean, 0.f) + eps);
    } else {
        // fp32 parity mode: centred two-pass variance
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) s += (r[i].x + r[i].y) + (r[i].z + r[i].w);
        s = warp_sum(s);
        mean = s * invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = lane + 32 * i;
            if (v < V) {
                float a = r[i].x - mean, bb = r[i].y - mean, c = r[i].z - mean, d = r[i].w - mean;
                q += (a * a + bb * bb) + (c * c + d * d);
            }
        }
        q = warp_sum(q);
        rstd = rsqrtf(q * invC + eps);
    }
    const long long obase = (long long)row * C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i;
        if (v < V) {
            const float4 gi = __ldg(reinterpret_cast<const float4*>(gamma) + v);
            const float4 bi = __ldg(reinterpret_cast<const float4*>(beta) + v);
            float y0 = (r[i].x - mean) * rstd * gi.x + bi.x;
            float y1 = (r[i].y - mean) * rstd * gi.y + bi.y;
            float y2 = (r[i].z - mean) * rstd * gi.z + bi.z;
            float y3 = (r[i].w - mean) * rstd * gi.w + bi.w;
            const long long o = obase + 4 * v;
            if (OUT_BF16) st_stream_u2(reinterpret_cast<__nv_bfloat16*>(out) + o, pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            else st_stream_f4(reinterpret_cast<float*>(out) + o, make_float4(y0, y1, y2, y3));
        }
    }
}

static inline int ln_rows_per_warp(int nv) { (void)nv; return 1; }

// largest per-sample statistics block (bytes) that the apply CTAs fold themselves (SDB200_GN_FUSED_MAX, measurement switch)
static long long gn_fused_stats_max_bytes() {
    static long long v = -1;
    // Default 0 = off.  Measured on B200 (profiles/r02_gn_fused_stats.txt): folding the statistics in every apply CTA costs more
    // than the finalize launch it saves — 22.5 vs 12.7 us at N=8, HW=256, C=1280 and 14.5 vs 10.7 us at HW=64 — because each of
    // the ~300 CTAs serialises load -> barrier -> fold -> barrier before its first row, while the finalize kernel overlaps the
    // producer's tail through programmatic dependent launch.
    if (v < 0) { const char* e = getenv("SDB200_GN_FUSED_MAX"); v = e ? atoll(e) : 0; }
    return v;
}

}  // namespace sdb

using namespace sdb;

extern "C" {

long long sdb_groupnorm_ws_bytes(int N, int HW, int C, int groups) {
    if (N <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % 4) return -1;
    GnGeom g = gn_geom(N, HW, C);
    return (long long)N * g.chunks * groups * 2 * sizeof(double) + (long long)N * groups * sizeof(float2) + 256;
}

int sdb_groupnorm_nhwc(const float* x0, int C0, const float* x1, int C1, int N, int HW, int groups,
                       float eps, const float* gamma, const float* beta, long long gb_stride, int act, int exact,
                       void* out, int out_dtype, void* raw_out, void* ws, int* counters, void* stream) {
    const int C = C0 + C1;
    SDB_REQUIRE(x0 && out && ws && gamma && beta && counters, "groupnorm: null pointer");
    SDB_REQUIRE(N > 0 && HW > 0 && C > 0, "groupnorm: empty tensor N=%d HW=%d C=%d", N, HW, C);
    SDB_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0, "groupnorm: C0=%d C1=%d must be multiples of 4", C0, C1);
    SDB_REQUIRE((C1 == 0) == (x1 == nullptr), "groupnorm: x1/C1 mismatch");
    SDB_REQUIRE(groups > 0 && C % groups == 0, "groupnorm: C=%d not divisible by groups=%d", C, groups);
    SDB_REQUIRE(C / 4 <= 1024, "groupnorm: C=%d too wide", C);
    SDB_REQUIRE(out_dtype == SDB_F32 || out_dtype == SDB_BF16, "groupnorm: bad out_dtype");
    SDB_REQUIRE((((uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 && gb_stride % 4 == 0 && gb_stride >= 0,
                "groupnorm: gamma/beta (and their per-sample stride) must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* raw = reinterpret_cast<__nv_bfloat16*>(raw_out);
    if (gn_cluster_enabled() && N <= 65535) {
        bool launched = false;
        int rc = SDB_OK;
#define TRY_CLUSTER(BF, EX, RW) rc = launch_gn_cluster<BF, EX, RW>(x0, C0, x1, C1, N, HW, groups, eps, gamma, beta, gb_stride, act, out, raw, st, &launched)
        if (raw) {
            if (out_dtype == SDB_BF16) { if (exact) TRY_CLUSTER(true, true, true); else TRY_CLUSTER(true, false, true); }
            else                       { if (exact) TRY_CLUSTER(false, true, true); else TRY_CLUSTER(false, false, true); }
        } else {
            if (out_dtype == SDB_BF16) { if (exact) TRY_CLUSTER(true, true, false); else TRY_CLUSTER(true, false, false); }
            else                       { if (exact) TRY_CLUSTER(false, true, false); else TRY_CLUSTER(false, false, false); }
        }
#undef TRY_CLUSTER
        if (rc || launched) return rc;
    }
    GnGeom g = gn_geom(N, HW, C);
    double* partial = reinterpret_cast<double*>(ws);
    float2* stats = reinterpret_cast<float2*>(reinterpret_cast<char*>(ws) +
                                              (long long)N * g.chunks * groups * 2 * sizeof(double));
    dim3 grid(g.chunks, N);
    size_t smem = (size_t)2 * g.R * C * sizeof(float);
    if (smem < (size_t)groups * sizeof(double2)) smem = (size_t)groups * sizeof(double2);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    launch_pdl(gn_stats_kernel, dim3(grid), dim3(g.threads), smem, st, x0, C0, x1, C1, HW, groups, g.V, g.R, g.rows_per_chunk, partial,
                                                   counters, (double)HW * (C / groups), eps, stats);
    int rc = check_launch("gn_stats_kernel");
    if (rc) return rc;
    const GnGeom ga = gn_apply_geom(N, HW, C);
    const dim3 grid_a(ga.chunks, N);
    CsSrc nos;
    memset(&nos, 0, sizeof(nos));
#define LAUNCH_APPLY(BF, EX, RW)                                                                         \
    launch_pdl(gn_apply_kernel<BF, EX, RW, false>, dim3(grid_a), dim3(ga.threads), 0, st, x0, C0, x1, C1, HW, groups, ga.V, ga.R, \
                                                             ga.rows_per_chunk, stats, gamma, beta, gb_stride, act, out, raw, nos, nos, 0.0, eps)
    if (raw) {
        if (out_dtype == SDB_BF16) { if (exact) LAUNCH_APPLY(true, true, true); else LAUNCH_APPLY(true, false, true); }
        else                       { if (exact) LAUNCH_APPLY(false, true, true); else LAUNCH_APPLY(false, false, true); }
    } else {
        if (out_dtype == SDB_BF16) { if (exact) LAUNCH_APPLY(true, true, false); else LAUNCH_APPLY(true, false, false); }
        else                       { if (exact) LAUNCH_APPLY(false, true, false); else LAUNCH_APPLY(false, false, false); }
    }
#undef LAUNCH_APPLY
    return check_launch("gn_apply_kernel");
}

int sdb_groupnorm_from_colstats(const float* x0, int C0, const float* cs0, const long long* layout0,
                                const float* x1, int C1, const float* cs1, const long long* layout1,
                                int N, int HW, int groups, float eps,
                                const float* gamma, const float* beta, long long gb_stride, int act, int exact,
                                void* out, int out_dtype, void* raw_out, void* ws, void* stream) {
    const int C = C0 + C1;
    SDB_REQUIRE(x0 && cs0 && out && ws && gamma && beta, "groupnorm_from_colstats: null pointer");
    SDB_REQUIRE(N > 0 && HW > 0 && C > 0 && layout0, "groupnorm_from_colstats: empty tensor");
    SDB_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0 && (C1 == 0) == (x1 == nullptr) && (C1 == 0) == (cs1 == nullptr),
                "groupnorm_from_colstats: bad channel split %d + %d", C0, C1);
    SDB_REQUIRE(groups > 0 && C % groups == 0 && C / 4 <= 1024, "groupnorm_from_colstats: C=%d groups=%d unsupported", C, groups);
    // layout = {slots, slots per sample (per region), regions, region stride}
    CsSrc s0, s1;
    memset(&s1, 0, sizeof(s1));
    s0.cs = cs0; s0.C = C0; s0.slots = layout0[0]; s0.spi = layout0[1]; s0.regions = (int)layout0[2]; s0.rstride = layout0[3];
    if (cs1) {
        SDB_REQUIRE(layout1, "groupnorm_from_colstats: second source has no layout");
        s1.cs = cs1; s1.C = C1; s1.slots = layout1[0]; s1.spi = layout1[1]; s1.regions = (int)layout1[2]; s1.rstride = layout1[3];
    }
    SDB_REQUIRE((long long)s0.regions * s0.spi * C + (long long)s1.regions * s1.spi * C < (1LL << 31),
                "groupnorm_from_colstats: too many statistics slots per sample");
    SDB_REQUIRE(s0.spi > 0 && s0.regions >= 1 && (s0.regions - 1) * s0.rstride + (long long)N * s0.spi <= s0.slots,
                "groupnorm_from_colstats: statistics buffer of source 0 too small");
    SDB_REQUIRE(!cs1 || (s1.spi > 0 && s1.regions >= 1 && (s1.regions - 1) * s1.rstride + (long long)N * s1.spi <= s1.slots),
                "groupnorm_from_colstats: statistics buffer of source 1 too small");
    SDB_REQUIRE(out_dtype == SDB_F32 || out_dtype == SDB_BF16, "groupnorm_from_colstats: bad out_dtype");
    SDB_REQUIRE((((uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 && gb_stride % 4 == 0 && gb_stride >= 0,
                "groupnorm_from_colstats: gamma/beta (and their per-sample stride) must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float2* stats = reinterpret_cast<float2*>(ws);                 // [N][groups]
    const int total = N * groups;
    const double count = (double)HW * (C / groups);
    // One sample's statistics small enough (a property of the per-sample geometry, never of the batch) -> every apply CTA folds
    // them itself and the finalize launch disappears; otherwise that would multiply the L2 traffic of the pass.
    const long long stat_bytes = ((long long)s0.regions * s0.spi * C0 + (long long)s1.regions * s1.spi * C1) * 8;
    const bool fused = stat_bytes <= gn_fused_stats_max_bytes() && groups <= GNA_MAX_GROUPS && (long long)C * 16 <= 40 * 1024;
    int rc = SDB_OK;
    if (!fused) {
        launch_pdl(gn_colstats_finalize_kernel, dim3(total), dim3(GNF_THREADS), 0, st, s0, s1, groups, count, eps, stats);
        rc = check_launch("gn_colstats_finalize_kernel");
        if (rc) return rc;
    }
    GnGeom g = gn_apply_geom(N, HW, C);
    dim3 grid(g.chunks, N);
    __nv_bfloat16* raw = reinterpret_cast<__nv_bfloat16*>(raw_out);
#define LAUNCH_APPLY2(BF, EX, RW, FU)                                                                    \
    launch_pdl(gn_apply_kernel<BF, EX, RW, FU>, dim3(grid), dim3(g.threads), FU ? (size_t)C * 16 : 0, st, x0, C0, x1, C1, HW, groups, g.V, g.R, \
               g.rows_per_chunk, stats, gamma, beta, gb_stride, act, out, raw, s0, s1, count, eps)
#define LAUNCH_APPLY(BF, EX, RW) do { if (fused) LAUNCH_APPLY2(BF, EX, RW, true); else LAUNCH_APPLY2(BF, EX, RW, false); } while (0)
    if (raw) {
        if (out_dtype == SDB_BF16) { if (exact) LAUNCH_APPLY(true, true, true); else LAUNCH_APPLY(true, false, true); }
        else                       { if (exact) LAUNCH_APPLY(false, true, true); else LAUNCH_APPLY(false, false, true); }
    } else {
        if (out_dtype == SDB_BF16) { if (exact) LAUNCH_APPLY(true, true, false); else LAUNCH_APPLY(true, false, false); }
        else                       { if (exact) LAUNCH_APPLY(false, true, false); else LAUNCH_APPLY(false, false, false); }
    }
#undef LAUNCH_APPLY
#undef LAUNCH_APPLY2
    return check_launch("gn_apply_kernel");
}

int sdb_layernorm(const float* x, int rows, int C, float eps, const float* gamma, const float* beta,
                  void* out, int out_dtype, void* stream) {
    SDB_REQUIRE(x && out && gamma && beta, "layernorm: null pointer");
    SDB_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, "layernorm: bad shape rows=%d C=%d", rows, C);
    SDB_REQUIRE(C <= 2048, "layernorm: C=%d too wide (max 2048)", C);
    cudaStream_t st = (cudaStream_t)stream;
    SDB_REQUIRE(out_dtype == SDB_BF16 || out_dtype == SDB_F32, "layernorm: bad out_dtype");
    const int threads = 256;
    const int nv = ceil_div(C / 4, 32);      // float4 vectors per lane; the rows stay in registers
    const int nvt = nv <= 1 ? 1 : (nv <= 3 ? 3 : (nv <= 5 ? 5 : (nv <= 10 ? 10 : 16)));   // template instance used below
    const int blocks = ceil_div(ceil_div(rows, ln_rows_per_warp(nvt)), threads / 32);
#define LAUNCH_LN(NV)                                                                                              \
    do {                                                                                                           \
        if (out_dtype == SDB_BF16) launch_pdl(layernorm_kernel<true, NV>, dim3(blocks), dim3(threads), 0, st, x, rows, C, eps, gamma, beta, out); \
        else launch_pdl(layernorm_kernel<false, NV>, dim3(blocks), dim3(threads), 0, st, x, rows, C, eps, gamma, beta, out);           \
    } while (0)
    if (nv <= 1) LAUNCH_LN(1);
    else if (nv <= 3) LAUNCH_LN(3);
    else if (nv <= 5) LAUNCH_LN(5);
    else if (nv <= 10) LAUNCH_LN(10);
    else LAUNCH_LN(16);
#undef LAUNCH_LN
    return check_launch("layernorm_kernel");
}

}  // extern "C"
