// Persistent warp-specialised bf16 GEMM for sm_100a:  out[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
//   A, W     bf16, K contiguous ("K-major"), loaded by TMA (128B swizzle) into a STAGES-deep smem ring
//   MMA      tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BN x K=16, fp32 accumulators in TMEM,
//            two accumulator stages (2*BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   roles    warp 0 lane 0: TMA producer | warp 1 lane 0: MMA issuer | warp 2: TMEM alloc/dealloc
//            warps 4..11: epilogue (warp%4 = TMEM lane quarter, (warp-4)/4 = column half), one thread per output row
//   tiles    static round-robin over (m_blk, n_blk) with n fastest, so CTAs resident at the same time share A rows
//            in L2 and the weight matrix (<= 8 MB) stays L2-resident for the whole launch.
//
// This is the kernel behind every Linear of the reference's LFQBert (bert.py:26-31,84,411-417).
//
// LayerNorm never runs as a kernel of its own.  The residual stream is kept as the PRE-norm sum y (bf16) plus per-row
// partial (sum, sum of squares) statistics, and the reference's post-norm structure  x = LN(y) ; out = Linear(x)  is
// evaluated as
//     Linear(LN(y))[m,n] = rstd_m * (sum_k y[m,k] W'[n,k]  -  mean_m * u[n]) + c[n]
//     W' = W * gamma (folded, bf16),  u[n] = sum_k W'[n,k],  c[n] = sum_k W[n,k] beta[k] + b[n]           ("LN-in" epilogues)
// and wherever the reference adds the residual x, the epilogue rebuilds x = (y - mean) * rstd * gamma + beta from y and the
// row statistics ("residual" epilogue), writes the new pre-norm sum as bf16 and emits its partial statistics.
// eps = 1e-12 as in the reference (bert.py:33,86,394,414).
#pragma once
#include "ptx.cuh"

#ifndef GEMM_TIMING_NO_STORE
#define GEMM_TIMING_NO_STORE 0   // timing experiments only: skip the bf16 output stores
#endif
#ifndef GEMM_TIMING_NO_GELU
#define GEMM_TIMING_NO_GELU 0    // timing experiments only
#endif

#ifndef GEMM_TIMING_NO_STATS
#define GEMM_TIMING_NO_STATS 0   // timing experiments only: skip the row-statistics stores of the STATS epilogues
#endif
#ifndef GEMM_GELU_MODE
#define GEMM_GELU_MODE 2         // 0: MUFU-erf form and polynomial form alternating pair by pair | 1: polynomial only | 2: MUFU form only
#endif
#ifndef GEMM_TIMING_NO_RES_LOAD
#define GEMM_TIMING_NO_RES_LOAD 0   // timing experiments only: residual epilogue runs on a zero slab (no residual loads)
#endif
#ifndef GEMM_TIMING_NO_MMA
#define GEMM_TIMING_NO_MMA 0     // timing experiments only (CTA-pair kernel): issue no MMAs, keep loads + barriers + epilogue
#endif
#ifndef GEMM_TIMING_NO_EPI
#define GEMM_TIMING_NO_EPI 0     // timing experiments only (CTA-pair kernel): epilogue warps only hand the accumulator back
#endif
#ifndef GEMM2_ROLES_HIGH
#define GEMM2_ROLES_HIGH 0       // CTA-pair kernel: 1 = epilogue warps 0..7, TMA / MMA / TMEM-alloc warps 8 / 9 / 10
#endif
#ifndef GEMM2_STAGES
#define GEMM2_STAGES 6
#endif
#ifndef GEMM_TRACE
#define GEMM_TRACE 0             // 1: the leader CTA of pair 0 records clock64 stamps per tile into GemmParams::trace
#endif
#if GEMM_TRACE
#define GEMM_EV(role, ev, it, val)                                                                                 \
    do {                                                                                                           \
        if (blockIdx.x == 0 && p.trace && (it) < 32u) p.trace[((role) * 4 + (ev)) * 32 + (it)] = (val);            \
    } while (0)
#else
#define GEMM_EV(role, ev, it, val) do {} while (0)
#endif
#ifndef GEMM_RES_PREFETCH
#define GEMM_RES_PREFETCH 1      // residual epilogue: load the thread's residual-row slab before the accumulator wait
#endif

namespace mb {

enum EpiMode : int {
    EPI_BIAS_BF16 = 0,            // out bf16 = acc + bias
    EPI_BIAS_GELU_BF16 = 1,       // out bf16 = gelu_erf(acc + bias)
    EPI_BIAS_RES_F32 = 2,         // out fp32 = acc + bias + residual(bf16)
    EPI_BIAS_F32_SEQ = 3,         // out fp32 = acc + bias, rows remapped: drop row seq_in-1 of every sequence (bert.py:503)
    EPI_BIAS_GELU_F32 = 4,        // out fp32 = gelu_erf(acc + bias)
    EPI_LNIN_BF16 = 5,            // out bf16 = rstd*(acc - mean*u) + c                               (QKV in-proj)
    EPI_LNIN_GELU_BF16 = 6,       // out bf16 = gelu_erf(rstd*(acc - mean*u) + c)                     (MLP up)
    EPI_RES_LN_BF16_STATS = 7,    // out bf16 = acc + bias + LN(y_res)  (+ partial row stats)         (attention out-proj, MLP down)
    EPI_LNIN_GELU_BF16_STATS = 8, // out bf16 = gelu_erf(rstd*(acc - mean*u) + c) (+ partial stats)   (head last_layer.0)
    EPI_LNIN_F32_SEQ = 9,         // out fp32 = rstd*(acc - mean*u) + c, class row dropped            (prediction layer)
};
constexpr int GEMM_NUM_EPI = 10;
constexpr int LN_PARTIALS = 8;    // partial (sum, sumsq) slots per row: one per 128-column epilogue-warp slab of a 1024-wide row

struct GemmParams {
    int M, N, K;
    const float* bias;               // [N]: bias | c (LN-in) | bias + beta (residual)
    const float* vec2;               // [N]: u (LN-in) | gamma (residual)
    const __nv_bfloat16* residual;   // [M, ldr]: residual (mode 2) | pre-norm y of the residual stream (mode 7)
    int ldr;
    const float2* stats_in;          // [M][LN_PARTIALS] partial (sum, sumsq) of the LayerNorm input rows (A rows or y_res rows)
    float2* stats_out;               // [M][LN_PARTIALS] partials of the rows written by this GEMM (STATS modes; N == 1024, BN == 256)
    float inv_d, eps;                // 1 / normalised width, LayerNorm eps
    void* out;                       // bf16 or fp32, row stride ldo elements
    int ldo;
    int seq_in, seq_out;             // *_SEQ: rows per sequence in A / kept rows per sequence in out
    long long* trace = nullptr;      // GEMM_TRACE builds: [roles 4][events 4][tiles 32] clock64 stamps, else unused
    // CTA-pair kernel, small M: the launch still fills the GPU and the pairs WITHOUT a tile pull the NEXT GEMM's weight matrix into
    // L2 (cp.async.bulk.prefetch.L2), before griddepcontrol.wait -- weights do not depend on the predecessor.  At batch 1 a forward
    // is a weight stream (610 MB per forward, no reuse), and a 12-tile GEMM reading its 8 MB from HBM through 24 SMs is bound by
    // their few loads in flight (~120 GB/s per SM); prefetched by ~120 otherwise idle CTAs the same bytes arrive at HBM speed and the
    // GEMM's own loads hit L2.
    const void* prefetch = nullptr;
    unsigned long long prefetch_bytes = 0;
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // 128 / 256 / 512: powers of two
    static constexpr int NUM_THREADS = 384;
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*statistics exchange*/;
};

// Exact-erf GELU (torch.nn.GELU() default, bert.py:29,413) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, the
// accuracy class of erff itself) on two MUFU ops (rcp, ex2) + 11 FMA-pipe ops instead of erff's ~30: the GELU epilogue of the
// MLP up-projection is issue-bound, not MMA-bound, with erff.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    q = fmaf(q, t, 0.5f * 1.421413741f);
    q = fmaf(q, t, 0.5f * -0.284496736f);
    q = fmaf(q, t, 0.5f * 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * -0.72134752044448170f));   // exp(-z^2)
    const float h = (q * t) * e;                   // 0.5 * erfc(z) = 1 - Phi(|x|)
    return fmaf(-fabsf(x), h, fmaxf(x, 0.f));      // x > 0: x - x h ; x < 0: x h
}

// Two GELUs at once on packed fp32 pairs (FFMA2 / FMUL2): 10 FMA-pipe + 4 ALU + 4 MUFU issue slots per pair instead of 2 x 16.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    float d0, d1, t0, t1, e0, e1, q0, q1, s0, s1, h0, h1;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(0.3275911f * 0.70710678118654752f), "f"(1.0f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
    // q = ((((a5 t + a4) t + a3) t + a2) t + a1) * 0.5, Horner on pairs
    asm("{\n\t.reg .b64 rt, rq, rk;\n\t"
        "mov.b64 rt, {%2, %3};\n\t"
        "mov.b64 rq, {%4, %4};\n\tmov.b64 rk, {%5, %5};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%6, %6};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%7, %7};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%8, %8};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mul.rn.f32x2 rq, rq, rt;\n\t"
        "mov.b64 {%0, %1}, rq;\n\t}"
        : "=f"(q0), "=f"(q1)
        : "f"(t0), "f"(t1), "f"(0.5f * 1.061405429f), "f"(0.5f * -1.453152027f), "f"(0.5f * 1.421413741f),
          "f"(0.5f * -0.284496736f), "f"(0.5f * 0.254829592f));
    asm("{\n\t.reg .b64 rx, rc, rd;\n\t"
        "mov.b64 rx, {%2, %3};\n\tmov.b64 rc, {%4, %4};\n\t"
        "mul.rn.f32x2 rd, rx, rx;\n\tmul.rn.f32x2 rd, rd, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(s0), "=f"(s1) : "f"(x0), "f"(x1), "f"(-0.72134752044448170f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
    // hm = 0.5 - q * e  (= Phi(|x|) - 0.5) ; out = 0.5 x + |x| * hm      (x Phi(x) = x/2 + |x| (Phi(|x|) - 1/2))
    asm("{\n\t.reg .b64 rq, re, ra, rx, rh, rd;\n\t"
        "mov.b64 rq, {%2, %3};\n\tmov.b64 re, {%4, %5};\n\tmov.b64 ra, {%6, %7};\n\tmov.b64 rx, {%8, %9};\n\tmov.b64 rh, {%10, %10};\n\t"
        "fma.rn.f32x2 rd, rq, re, rh;\n\tmul.rn.f32x2 rd, ra, rd;\n\tfma.rn.f32x2 rd, rx, rh, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(h0), "=f"(h1) : "f"(-q0), "f"(-q1), "f"(e0), "f"(e1), "f"(a0), "f"(a1), "f"(x0), "f"(x1), "f"(0.5f));
    x0 = h0; x1 = h1;
}

// The same two GELUs without the MUFU pipe: Phi(x) - 0.5 = x * Q(x^2), Q a degree-9 least-squares polynomial on x^2 <= 4.5^2
// (Phi saturated to 0 / 1 beyond: Phi(4.5) = 1 - 3.4e-6); |gelu error| <= 5e-5 at |x| ~ 4.5 and ~1e-5 elsewhere, far below the bf16
// rounding of the stored value.  GEMM_GELU_MODE picks the mix.  When the epilogue was the up-projection's critical path the two
// forms alternated pair by pair (MUFU form alone: the special-function pipe, 16 ops / clk / SM and two per element, was as busy
// as the tensor pipe over a 128 x 256 x 1024 tile).  With the epilogue overlapped, the job is energy-bound and the MUFU form
// alone measures 1 % less energy per launch than either the mix or the polynomial alone (tools/kpower.py), and is the more
// accurate of the two: it is the default.
__device__ __forceinline__ void gelu_poly2(float& x0, float& x1) {
    // t = min(x^2, 4.5^2): beyond the fitted range x Q(t) is linear in x with slope Q(4.5^2) = (0.5 - 3.4e-6) / 4.5, so
    // 0.5 + x Q(t) leaves [0, 1] there and the saturating FMA returns Phi = 0 / 1 -- no clamp of x itself.
    float t0, t1;
    fmul2(t0, t1, x0, x1, x0, x1);
    t0 = fminf(t0, 20.25f); t1 = fminf(t1, 20.25f);
    float q0, q1;
    asm("{\n\t.reg .b64 rt, rq, rk;\n\t"
        "mov.b64 rt, {%2, %3};\n\t"
        "mov.b64 rq, {%4, %4};\n\tmov.b64 rk, {%5, %5};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%6, %6};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%7, %7};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%8, %8};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%9, %9};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%10, %10};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%11, %11};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%12, %12};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 rk, {%13, %13};\n\tfma.rn.f32x2 rq, rq, rt, rk;\n\t"
        "mov.b64 {%0, %1}, rq;\n\t}"
        : "=f"(q0), "=f"(q1)
        : "f"(t0), "f"(t1), "f"(-1.334576828e-12f), "f"(1.631205285e-10f), "f"(-8.910449133e-09f), "f"(2.891752189e-07f),
          "f"(-6.269251990e-06f), "f"(9.702424500e-05f), "f"(-1.118151080e-03f), "f"(9.819429864e-03f), "f"(-6.631637927e-02f),
          "f"(3.988728748e-01f));
    fmul2(x0, x1, x0, x1, __saturatef(fmaf(x0, q0, 0.5f)), __saturatef(fmaf(x1, q1, 0.5f)));     // x * Phi(x)
}

// mean and rstd of a row from its LN_PARTIALS partial sums
__device__ __forceinline__ void ln_row_stats(const float2* __restrict__ st, float inv_d, float eps, float& mean, float& rstd) {
    const float4* q = reinterpret_cast<const float4*>(st);
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < LN_PARTIALS / 2; ++i) {
        const float4 v = __ldg(q + i);
        s += v.x + v.z; ss += v.y + v.w;
    }
    mean = s * inv_d;
    const float var = fmaxf(fmaf(-mean, mean, ss * inv_d), 0.f);
    rstd = rsqrtf(var + eps);
}

// One output row per thread.  epi_prepare runs before the accumulator wait (its loads overlap the tile's MMAs), epi_run after.
struct EpiRow { int row; bool row_ok, store_ok; long long out_row; float rs, nmr; };

// The thread's slab of the residual-stream row (COLS bf16 columns = COLS/8 16-byte words) for the residual epilogue, loaded
// BEFORE the accumulator wait: y comes from HBM (the residual stream is larger than L2), and four dependent
// load -> use round trips per tile inside the epilogue were what held the K=1024 out-projection at 47 % tensor-pipe time.
template <int COLS>
struct ResSlab { uint4 v[COLS / 8]; };
template <int COLS>
__device__ __forceinline__ void epi_load_residual(const GemmParams& p, const EpiRow& er, int col0, ResSlab<COLS>& s) {
    // 256-bit loads: every lane takes whole 32-byte sectors of its own row (a 128-bit load per lane fetches each sector twice)
    const __nv_bfloat16* rp = p.residual + (size_t)(er.row_ok ? er.row : 0) * p.ldr + col0;
#pragma unroll
    for (int i = 0; i < COLS / 16; ++i)
        if (GEMM_TIMING_NO_RES_LOAD) { s.v[2 * i] = make_uint4(0, 0, 0, 0); s.v[2 * i + 1] = make_uint4(0, 0, 0, 0); } else
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(s.v[2 * i].x), "=r"(s.v[2 * i].y), "=r"(s.v[2 * i].z), "=r"(s.v[2 * i].w), "=r"(s.v[2 * i + 1].x),
                       "=r"(s.v[2 * i + 1].y), "=r"(s.v[2 * i + 1].z), "=r"(s.v[2 * i + 1].w)
                     : "l"(rp + 16 * i));
}

template <int EPI>
__device__ __forceinline__ EpiRow epi_prepare(const GemmParams& p, int row) {
    constexpr bool kLnIn = EPI == EPI_LNIN_BF16 || EPI == EPI_LNIN_GELU_BF16 || EPI == EPI_LNIN_GELU_BF16_STATS || EPI == EPI_LNIN_F32_SEQ;
    constexpr bool kSeq = EPI == EPI_BIAS_F32_SEQ || EPI == EPI_LNIN_F32_SEQ;
    EpiRow r;
    r.row = row; r.row_ok = row < p.M; r.out_row = row; r.store_ok = r.row_ok;
    if (kSeq) {
        const int sq = row / p.seq_in, rr = row - sq * p.seq_in;
        r.store_ok = r.row_ok && rr < p.seq_out;
        r.out_row = (long long)sq * p.seq_out + rr;
    }
    // LayerNorm row statistics: of the A row for LN-in modes, of the residual-stream row for the residual mode
    r.rs = 1.f; r.nmr = 0.f;                // rstd, -mean * rstd
    // (residual mode with stats_in == nullptr: the residual is added as stored, rs = 1, nmr = 0 -- the pre-norm trunk)
    if ((kLnIn || EPI == EPI_RES_LN_BF16_STATS) && r.row_ok && p.stats_in != nullptr) {
        float mean, rstd;
        ln_row_stats(p.stats_in + (size_t)row * LN_PARTIALS, p.inv_d, p.eps, mean, rstd);
        r.rs = rstd; r.nmr = -mean * rstd;
    }
    return r;
}

// Output staging for the TMA-store path: the warp's 32 rows x 64 bf16 columns as 128-byte rows with the 128B swizzle
// (what a box {64, 32} store with CU_TENSOR_MAP_SWIZZLE_128B reads).  Direct per-thread stores write 16 B of a different
// row per lane and instruction -- half-sector L2 writes that cap the K=1024 GEMMs at ~70 % of their store-free speed.
struct EpiStage {
    uint8_t* buf;             // 4 KB, 1024-byte aligned, private to the warp
    const CUtensorMap* tm;    // output tensor map, box {64 columns, 32 rows}
    int row0;                 // first output row of the warp's 32-row slab
};

// taddr: TMEM address of this warp's lane quarter at column 0 of the accumulator tile; the warp handles columns
// [part * BN/NSPLIT, (part+1) * BN/NSPLIT) of the BN-wide tile n_blk.  kTma: bf16 output through smem + cp.async.bulk.tensor store.
// svec: the tile's per-column vectors staged in shared memory ([BN] bias | [BN] vec2), or nullptr to read them from global.
// With ~225 KB of the SM's 228 KB carved out as shared memory there is no L1 left: every __ldg of bias / u / gamma is an L2
// round trip issued after the accumulator wait, and those round trips (not arithmetic) were most of the epilogue's time.
// kResS: the residual rows were loaded by TMA into the staging tiles (stg.buf + 4096 * (64-column pair), 128B swizzle); each
// thread reads its row's 64 bytes of a chunk and later writes the chunk's results over them, then the pair is TMA-stored.
template <int BN, int EPI, bool kTma = false, int NSPLIT = 2, bool kPre = false, bool kSV = false, bool kResS = false>
__device__ __forceinline__ void epi_run(const GemmParams& p, const EpiRow& er, uint32_t taddr, int n_blk, int half,
                                        EpiStage stg = EpiStage{nullptr, nullptr, 0}, const uint4* res = nullptr,
                                        const float* svec = nullptr, float2* stats_xchg = nullptr) {
    constexpr bool kLnIn = EPI == EPI_LNIN_BF16 || EPI == EPI_LNIN_GELU_BF16 || EPI == EPI_LNIN_GELU_BF16_STATS || EPI == EPI_LNIN_F32_SEQ;
    constexpr bool kGelu = EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_F32 || EPI == EPI_LNIN_GELU_BF16 || EPI == EPI_LNIN_GELU_BF16_STATS;
    constexpr bool kStats = EPI == EPI_RES_LN_BF16_STATS || EPI == EPI_LNIN_GELU_BF16_STATS;
    constexpr bool kOutBf16 = EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_LNIN_BF16 || EPI == EPI_LNIN_GELU_BF16 ||
                              EPI == EPI_RES_LN_BF16_STATS || EPI == EPI_LNIN_GELU_BF16_STATS;
    constexpr int COLS_PER_WARP = BN / NSPLIT;
    const int row = er.row;
    const bool row_ok = er.row_ok, store_ok = er.store_ok;
    const long long out_row = er.out_row;
    const float rs = er.rs, nmr = er.nmr;
    float st_sum = 0.f, st_sq = 0.f, st_sum1 = 0.f, st_sq1 = 0.f;   // even / odd column partial sums (packed pairs)
    auto chunk_v = [&](const int c, const uint32_t (&v)[32]) {
        const int col0 = half * COLS_PER_WARP + c;
        const int n0 = n_blk * BN + col0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b4 = kSV ? *reinterpret_cast<const float4*>(svec + col0 + j) : __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
            if (kLnIn) {   // rstd*acc + (-mean*rstd)*u + c, on packed pairs
                const float4 u4 = kSV ? *reinterpret_cast<const float4*>(svec + BN + col0 + j)
                                      : __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j));
                float t0, t1, t2, t3;
                ffma2(t0, t1, nmr, nmr, u4.x, u4.y, b4.x, b4.y);
                ffma2(t2, t3, nmr, nmr, u4.z, u4.w, b4.z, b4.w);
                ffma2(f[j + 0], f[j + 1], rs, rs, __uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]), t0, t1);
                ffma2(f[j + 2], f[j + 3], rs, rs, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]), t2, t3);
            } else {
                fadd2(f[j + 0], f[j + 1], __uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]), b4.x, b4.y);
                fadd2(f[j + 2], f[j + 3], __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]), b4.z, b4.w);
            }
        }
        if (kGelu && !GEMM_TIMING_NO_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (GEMM_GELU_MODE == 1) { gelu_poly2(f[j], f[j + 1]); gelu_poly2(f[j + 2], f[j + 3]); }
                else if (GEMM_GELU_MODE == 2) { gelu_erf2(f[j], f[j + 1]); gelu_erf2(f[j + 2], f[j + 3]); }
                else { gelu_erf2(f[j], f[j + 1]); gelu_poly2(f[j + 2], f[j + 3]); }
            }
        }
        if (EPI == EPI_BIAS_RES_F32 || EPI == EPI_RES_LN_BF16_STATS) {
            if (row_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 r4;
                    if (kResS) r4 = *reinterpret_cast<const uint4*>(stg.buf + (c >> 6) * 4096 + (threadIdx.x & 31) * 128 +
                                                                   (((((c >> 5) & 1) * 4 + j) ^ (threadIdx.x & 7)) << 4));
                    else r4 = kPre ? res[(c >> 3) + j] : __ldg(rp + j);
                    const uint32_t w[4] = {r4.x, r4.y, r4.z, r4.w};
                    if (EPI == EPI_RES_LN_BF16_STATS) {   // + ((y - mean) * rstd) * gamma   (beta is folded into p.bias)
                        const float4 g0 = kSV ? *reinterpret_cast<const float4*>(svec + BN + col0 + j * 8)
                                              : __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j * 8));
                        const float4 g1 = kSV ? *reinterpret_cast<const float4*>(svec + BN + col0 + j * 8 + 4)
                                              : __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j * 8 + 4));
                        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            float r0, r1;
                            ffma2(r0, r1, __uint_as_float(w[t] << 16), __uint_as_float(w[t] & 0xffff0000u), rs, rs, nmr, nmr);
                            ffma2(f[j * 8 + 2 * t], f[j * 8 + 2 * t + 1], r0, r1, g[2 * t], g[2 * t + 1], f[j * 8 + 2 * t], f[j * 8 + 2 * t + 1]);
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            f[j * 8 + 2 * t + 0] += __uint_as_float(w[t] << 16);
                            f[j * 8 + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
                        }
                    }
                }
            }
        }
        if (kOutBf16) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h);
                if (kStats) {   // statistics of the values as stored (bf16-rounded): the consumer normalises those
                    const float a = __uint_as_float(w[j] << 16), b = __uint_as_float(w[j] & 0xffff0000u);
                    fadd2(st_sum, st_sum1, st_sum, st_sum1, a, b);
                    ffma2(st_sq, st_sq1, a, b, a, b, st_sq, st_sq1);
                }
            }
            if (kTma) {
                const int lane = threadIdx.x & 31;
                const int sub = (c >> 5) & 1;                 // which 32-column half of the 64-column staging row
                if (sub == 0 && !kResS) {                     // the previous store must have finished reading the buffer
                    if (elect_one()) tma_store_wait_read<0>();   // bulk groups are per thread: the lane that issued the store
                    __syncwarp();
                }
                uint8_t* rowp = stg.buf + (kResS ? (c >> 6) * 4096 : 0) + lane * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(rowp + (((sub * 4 + j) ^ (lane & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                if (sub == 1) {
                    fence_async_proxy();                      // generic-proxy smem writes -> visible to the bulk copy
                    __syncwarp();
                    if (elect_one()) {
                        tma_store_2d(stg.tm, stg.buf + (kResS ? (c >> 6) * 4096 : 0), n0 - 32, stg.row0);
                        tma_store_commit();
                    }
                }
            } else if (store_ok && !GEMM_TIMING_NO_STORE) {
                // 256-bit stores: every lane writes whole 32-byte sectors of its own row
                __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out) + (size_t)out_row * p.ldo + n0;
                st_global_v8(op, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
                st_global_v8(op + 16, w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]);
            }
        } else if (store_ok) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)out_row * p.ldo + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
    };
    auto chunk = [&](const int c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + half * COLS_PER_WARP + c, v);
        tmem_ld_wait();
        chunk_v(c, v);
    };
    // (Loading the accumulator chunk c+1 out of TMEM while chunk c is processed was tried for the residual epilogue of the K = 1024
    //  out-projection and changed nothing, isolated or sustained: profiles/r02_epi7_pipe_neutral.txt.  That GEMM runs at 81 % tensor-pipe
    //  time next to 50-60 % of the HBM rate -- A, the residual stream and the output are each as large as the weights are small.)
    if constexpr (kPre) {   // prefetched slab: indices must be compile-time constants to stay in registers
#pragma unroll
        for (int c = 0; c < COLS_PER_WARP; c += 32) chunk(c);
    } else {
#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP; c += 32) chunk(c);
    }
    if constexpr (kStats && COLS_PER_WARP == 64) {
        // 128-column tiles (the 1-CTA kernel at small M): the two warps of a lane quarter hold half a statistics slot each; the upper
        // half hands its sums over through shared memory and the lower half stores lower + upper (a fixed order)
        const int quarter = (threadIdx.x >> 5) & 3, lane = threadIdx.x & 31;
        if (half == 1) stats_xchg[quarter * 32 + lane] = make_float2(st_sum + st_sum1, st_sq + st_sq1);
        named_bar_sync(4 + quarter, 64);
        if (half == 0 && row_ok && !GEMM_TIMING_NO_STATS) {
            const float2 o = stats_xchg[quarter * 32 + lane];
            p.stats_out[(size_t)row * LN_PARTIALS + (n_blk * BN) / 128] = make_float2((st_sum + st_sum1) + o.x, (st_sq + st_sq1) + o.y);
        }
        named_bar_sync(4 + quarter, 64);            // the slot is free for the next tile
    } else if (kStats && row_ok && !GEMM_TIMING_NO_STATS) {   // slot = index of the warp's 128-column slab in the 1024-wide row
        // (other slab widths are rejected on the host: launch_gemm)
        float2* so = p.stats_out + (size_t)row * LN_PARTIALS + (n_blk * BN + half * COLS_PER_WARP) / 128;
        so[0] = make_float2(st_sum + st_sum1, st_sq + st_sq1);
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmParams p) {
    using C = GemmCfg<BN>;
    constexpr int BM = C::BM, BK = C::BK, STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzle atoms are 1024 B: align the ring manually (dynamic smem base is only guaranteed 16 B aligned)
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* base = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* smem_a = base;
    uint8_t* smem_b = base + STAGES * C::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * (C::A_BYTES + C::B_BYTES));
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float2* stats_xchg = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [4 lane quarters][32 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (p.M + BM - 1) / BM, num_n = p.N / BN;
    const int num_tiles = num_m * num_n, num_k = p.K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {  // ---------------- TMA producer (elect.sync region: operands stay in uniform registers, no R2UR waterfall per issue)
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], C::A_BYTES + C::B_BYTES);
                    tma_load_2d(smem_a + stage * C::A_BYTES, &tm_a, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(smem_b + stage * C::B_BYTES, &tm_b, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {  // ---------------- MMA issuer
            constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, BM, BN);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = make_sdesc_k128(smem_u32(smem_a + stage * C::A_BYTES));
                    const uint64_t b_desc = make_sdesc_k128(smem_u32(smem_b + stage * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)   // +32 B along K inside the swizzle atom = +2 in (addr >> 4)
                        umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);      // smem slot reusable once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[as]);         // accumulator complete
            }
        }
    } else if (warp >= 4) {  // ---------------- epilogue: TMEM -> registers -> global
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const EpiRow er = epi_prepare<EPI>(p, m_blk * BM + quarter * 32 + lane);
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            epi_run<BN, EPI>(p, er, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN, n_blk, half,
                             EpiStage{nullptr, nullptr, 0}, nullptr, nullptr, stats_xchg);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair variant
// Same pipeline with tcgen05.mma.cta_group::2: a cluster of two CTAs (one TPC) computes a 256 x 256 tile, each CTA holding
// the accumulators of its 128 rows in its own TMEM and staging only HALF of the B tile (128 of the 256 weight rows); the
// tensor cores of both SMs read both halves.  Per CTA and k-block that is 16 KB of A + 16 KB of B instead of 16 + 32 KB:
// a third less L2 -> SM traffic and shared-memory fill per MMA (the 1-CTA kernel is bound there, not in the tensor pipe),
// which also makes room for 6 pipeline stages instead of 4.
//   leader CTA (cluster rank 0): arms full[s] for the bytes of BOTH CTAs and issues every MMA;
//   both CTAs: TMA producer for their own halves (completion signalled on the leader's full[s]), epilogue of their own rows
//   by 8 warps;
//   tcgen05.commit multicasts the "stage consumed" / "accumulator ready" arrivals to both CTAs; the epilogue warps of both
//   CTAs release the accumulator stage on the leader's tmem_empty barrier.
// kResTma (residual epilogue): the residual rows arrive by TMA in the epilogue's own staging tiles (two 64-column tiles per warp,
// results written over them in place), paid for with one pipeline stage.
template <bool kResTma>
struct Gemm2CfgT {
    static constexpr int BM = 256, BN = 256, BK = 64, STAGES = kResTma ? 5 : GEMM2_STAGES;
    // 4 TMEM lane quarters x 2 column halves.  16 warps (4 x 4, 5 stages to make room for their staging) measured SLOWER on
    // every GEMM (QKV 1435 -> 1249, up 1272 -> 1146 TFLOP/s): the extra warps and the lost stage cost more than the added
    // epilogue parallelism buys.
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
    static constexpr int A_BYTES = 128 * BK * 2;      // per CTA
    static constexpr int B_BYTES = 128 * BK * 2;      // per CTA (half of the 256 weight rows)
    static constexpr int STG_WARP = kResTma ? 8192 : 4096;   // per epilogue warp: 32 rows x 128 B (x 2 column pairs with kResTma)
    static constexpr int STG_BYTES = EPI_WARPS * STG_WARP;   // output staging
    static constexpr int VEC_BYTES = 2 * BN * 4;       // the tile's per-column epilogue vectors: [BN] bias | [BN] vec2
    static constexpr int BAR_BYTES = 256;              // 2*STAGES + 4 (+ EPI_WARPS residual) mbarriers + the TMEM base slot
    // The dynamic shared-memory window is 1024-byte aligned (declared so, and checked at kernel entry): no alignment slack,
    // which is what leaves room for VEC_BYTES next to six pipeline stages (232 448 B limit).
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + STG_BYTES + VEC_BYTES + BAR_BYTES;
    static_assert(SMEM_BYTES <= 232448, "CTA-pair GEMM shared memory over the 227 KB limit");
    static_assert((2 * STAGES + 4 + EPI_WARPS) * 8 + 4 <= BAR_BYTES, "barrier area too small");
};
using Gemm2Cfg = Gemm2CfgT<false>;
#ifndef GEMM2_RES_TMA
#define GEMM2_RES_TMA 1          // residual epilogue of the CTA-pair kernel: residual rows by TMA into the staging tiles (else LDG.256)
#endif
__host__ __device__ constexpr bool gemm2_res_tma(int epi) { return GEMM2_RES_TMA && epi == EPI_RES_LN_BF16_STATS; }
__host__ __device__ constexpr int gemm2_smem_bytes(int epi) { return gemm2_res_tma(epi) ? Gemm2CfgT<true>::SMEM_BYTES : Gemm2CfgT<false>::SMEM_BYTES; }
// bf16-output epilogues of the CTA-pair kernel go through shared memory and TMA stores
#ifndef GEMM2_DIRECT256
#define GEMM2_DIRECT256 0   // 1: CTA-pair kernel stores straight from registers with 256-bit stores instead of smem + TMA
#endif
__host__ __device__ constexpr bool gemm2_tma_store(int epi) {
    return !GEMM2_DIRECT256 && (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_LNIN_BF16 || epi == EPI_LNIN_GELU_BF16 ||
           epi == EPI_RES_LN_BF16_STATS || epi == EPI_LNIN_GELU_BF16_STATS);
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg::THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                          const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_r, GemmParams p) {
    constexpr bool kResTma = gemm2_res_tma(EPI);     // tm_r: the residual tensor, box {64 columns, 32 rows} like tm_c
    using C = Gemm2CfgT<kResTma>;
    constexpr int BN = C::BN, BK = C::BK, STAGES = C::STAGES;
    constexpr bool kTma = gemm2_tma_store(EPI);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    if (smem_u32(base) & 1023u) asm volatile("trap;");   // 128B-swizzle atoms need the 1024-byte alignment declared above
    uint8_t* smem_a = base;
    uint8_t* smem_b = base + STAGES * C::A_BYTES;
    uint8_t* smem_stg = base + STAGES * (C::A_BYTES + C::B_BYTES);
    float* smem_vec = reinterpret_cast<float*>(smem_stg + C::STG_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + C::STG_BYTES + C::VEC_BYTES);
    uint64_t* full = bars;                          // used in the leader CTA only
    uint64_t* empty = bars + STAGES;                // per CTA, arrived by the leader's multicast commit
    uint64_t* tmem_full = bars + 2 * STAGES;        // per CTA, arrived by the leader's multicast commit
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // leader only: 8 epilogue warps of each CTA
    uint64_t* res_full = bars + 2 * STAGES + 4;     // kResTma: per epilogue warp, the tile's residual boxes have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + C::EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) GEMM_EV(3, 0, 0u, clock64());   // role 3 = kernel lifecycle: entry | set up | predecessor complete | work done
    // Warp roles.  The SM's warp schedulers favour the highest warp id among eligible warps, so with GEMM2_ROLES_HIGH the
    // single-thread TMA / MMA issuers sit above the eight epilogue warps instead of below them.
    constexpr int W_EPI0 = GEMM2_ROLES_HIGH ? 0 : 4, W_PROD = GEMM2_ROLES_HIGH ? 8 : 0, W_MMA = GEMM2_ROLES_HIGH ? 9 : 1,
                  W_ALLOC = GEMM2_ROLES_HIGH ? 10 : 2;
    const bool is_epi = warp >= W_EPI0 && warp < W_EPI0 + C::EPI_WARPS;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_m = (p.M + C::BM - 1) / C::BM, num_n = p.N / BN;
    const int num_tiles = num_m * num_n, num_k = p.K / BK;

    if (warp == W_PROD && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        if (kTma) tma_prefetch_desc(&tm_c);
        if (kResTma) tma_prefetch_desc(&tm_r);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * C::EPI_WARPS); }
        for (int s = 0; s < C::EPI_WARPS; ++s) mbar_init(&res_full[s], 1);
        fence_mbar_init();
    }
    if (warp == W_ALLOC) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                              // barriers of both CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) GEMM_EV(3, 1, 0u, clock64());
    if (p.prefetch != nullptr && pair >= num_tiles && warp == W_PROD) {
        if (elect_one()) {  // ---------------- idle pair: its share of the next GEMM's weights -> L2, 4 KB per instruction
            const unsigned long long n_idle = 2ull * (unsigned)(num_pairs - num_tiles), idx = blockIdx.x - 2u * (unsigned)num_tiles;
            const unsigned long long share = ((p.prefetch_bytes + n_idle - 1) / n_idle + 4095ull) & ~4095ull;
            const unsigned long long b0 = idx * share, b1 = b0 + share < p.prefetch_bytes ? b0 + share : p.prefetch_bytes;
            const char* src = static_cast<const char*>(p.prefetch);
            for (unsigned long long off = b0; off < b1; off += 4096ull) {
                const uint32_t n = b1 - off < 4096ull ? (uint32_t)((b1 - off) & ~15ull) : 4096u;
                if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(n) : "memory");
            }
        }
    }
    pdl_launch_dependents();
    pdl_wait();                                      // the predecessor's outputs (A rows, statistics, residual) are complete from here on
    if (threadIdx.x == 0) GEMM_EV(3, 2, 0u, clock64());

    if (warp == W_PROD) {
        if (elect_one()) {  // ---------------- TMA producer (each CTA: its 128 A rows, its 128 of the 256 B rows)
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
#if GEMM_TRACE
                long long empty_wait = 0;
#endif
                for (int kb = 0; kb < num_k; ++kb) {
#if GEMM_TRACE
                    const long long t0 = clock64();
#endif
                    mbar_wait(&empty[stage], phase ^ 1);
#if GEMM_TRACE
                    empty_wait += clock64() - t0;
                    if (kb == num_k - 1) { GEMM_EV(2, 0, it, empty_wait); GEMM_EV(2, 1, it, clock64()); }
#endif
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (C::A_BYTES + C::B_BYTES));
                    const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
                    tma_load_2d_2sm(smem_a + stage * C::A_BYTES, &tm_a, bar, kb * BK, m_blk * C::BM + (int)rank * 128);
                    tma_load_2d_2sm(smem_b + stage * C::B_BYTES, &tm_b, bar, kb * BK, n_blk * BN + (int)rank * 128);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        if (leader && elect_one()) {  // ---------------- MMA issuer (leader CTA only)
            constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 256, BN);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                GEMM_EV(0, 0, it, clock64());
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                GEMM_EV(0, 1, it, clock64());
                const uint32_t d_tmem = tmem_base + as * BN;
#if GEMM_TRACE
                long long full_wait = 0;
#endif
                for (int kb = 0; kb < num_k; ++kb) {
#if GEMM_TRACE
                    const long long t0 = clock64();
#endif
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
#if GEMM_TRACE
                    full_wait += clock64() - t0;
#endif
                    const uint64_t a_desc = make_sdesc_k128(smem_u32(smem_a + stage * C::A_BYTES));
                    const uint64_t b_desc = make_sdesc_k128(smem_u32(smem_b + stage * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        if (!GEMM_TIMING_NO_MMA) umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2sm(&empty[stage], 3);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[as], 3);
#if GEMM_TRACE
                GEMM_EV(0, 2, it, full_wait);
                GEMM_EV(0, 3, it, clock64());
#endif
            }
        }
    } else if (is_epi) {  // ---------------- epilogue: each CTA its own 128 rows; warp = (lane quarter, 128-column half)
        const int ew = warp - W_EPI0, quarter = warp & 3, half = ew >> 2;
        // Per-column vectors of the tile -> shared memory, one float2 per epilogue thread (threads 0..127 bias, 128..255 vec2).
        // The global load for tile i+1 is issued before tile i's epilogue runs and parked in a register pair.
        const int e = threadIdx.x - W_EPI0 * 32;
        const float* vsrc = e < BN / 2 ? p.bias : p.vec2;
        const int vcol = 2 * (e & (BN / 2 - 1));
        float2 vreg = make_float2(0.f, 0.f);
        if (pair < num_tiles && vsrc) vreg = __ldg(reinterpret_cast<const float2*>(vsrc + (pair % num_n) * BN + vcol));
        uint32_t it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            named_bar_sync(1, C::EPI_WARPS * 32);                       // every warp has finished reading the previous tile's vectors
            reinterpret_cast<float2*>(smem_vec)[e] = vreg;
            named_bar_sync(2, C::EPI_WARPS * 32);
            if (tile + num_pairs < num_tiles && vsrc)
                vreg = __ldg(reinterpret_cast<const float2*>(vsrc + ((tile + num_pairs) % num_n) * BN + vcol));
            const int row0 = m_blk * C::BM + (int)rank * 128 + quarter * 32;
            const EpiRow er = epi_prepare<EPI>(p, row0 + lane);
            constexpr int CPW = BN / (C::EPI_WARPS / 4);
            constexpr bool kPre = GEMM_RES_PREFETCH && EPI == EPI_RES_LN_BF16_STATS && !kResTma;
            ResSlab<kPre ? CPW : 8> slab;
            if constexpr (kPre) epi_load_residual<CPW>(p, er, n_blk * BN + half * CPW, slab);
            uint8_t* stg_w = smem_stg + ew * C::STG_WARP;
            if constexpr (kResTma) {
                // the previous tile's two stores have read the staging tiles -> refill them with this tile's residual rows
                if (elect_one()) {
                    tma_store_wait_read<0>();
                    mbar_arrive_expect_tx(&res_full[ew], 2 * 4096);
                    tma_load_2d(stg_w, &tm_r, &res_full[ew], n_blk * BN + half * CPW, row0);
                    tma_load_2d(stg_w + 4096, &tm_r, &res_full[ew], n_blk * BN + half * CPW + 64, row0);
                }
                __syncwarp();
            }
            if (ew == 0 && lane == 0) GEMM_EV(1, 0, it, clock64());
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            if (ew == 0 && lane == 0) GEMM_EV(1, 1, it, clock64());
            if constexpr (kResTma) mbar_wait(&res_full[ew], it & 1);
            if (!GEMM_TIMING_NO_EPI)
            epi_run<BN, EPI, kTma, C::EPI_WARPS / 4, kPre, true, kResTma>(p, er, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN,
                                                                          n_blk, half, EpiStage{stg_w, &tm_c, row0}, slab.v, smem_vec);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[as]), 0));
            if (ew == 0 && lane == 0) GEMM_EV(1, 2, it, clock64());
        }
        if (kTma && elect_one()) tma_store_wait_all<0>();   // bulk stores complete before the CTA retires its smem
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();                              // the leader's MMAs write the peer's TMEM: both done before dealloc
    if (threadIdx.x == 0) GEMM_EV(3, 3, 0u, clock64());
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem_base);
    }
}

}  // namespace mb
