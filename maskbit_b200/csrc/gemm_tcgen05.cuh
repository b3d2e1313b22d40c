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
constexpr int LN_PARTIALS = 8;    // partial (sum, sumsq) slots per row: one per (256-wide n-block, 128-column half) of a 1024-wide row

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
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128, BK = 64;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // 128 / 256 / 512: powers of two
    static constexpr int NUM_THREADS = 384;
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
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

template <int BN, int EPI>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmParams p) {
    using C = GemmCfg<BN>;
    constexpr int BM = C::BM, BK = C::BK, STAGES = C::STAGES;
    constexpr bool kLnIn = EPI == EPI_LNIN_BF16 || EPI == EPI_LNIN_GELU_BF16 || EPI == EPI_LNIN_GELU_BF16_STATS || EPI == EPI_LNIN_F32_SEQ;
    constexpr bool kGelu = EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_F32 || EPI == EPI_LNIN_GELU_BF16 || EPI == EPI_LNIN_GELU_BF16_STATS;
    constexpr bool kStats = EPI == EPI_RES_LN_BF16_STATS || EPI == EPI_LNIN_GELU_BF16_STATS;
    constexpr bool kSeq = EPI == EPI_BIAS_F32_SEQ || EPI == EPI_LNIN_F32_SEQ;
    constexpr bool kOutBf16 = EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_LNIN_BF16 || EPI == EPI_LNIN_GELU_BF16 ||
                              EPI == EPI_RES_LN_BF16_STATS || EPI == EPI_LNIN_GELU_BF16_STATS;
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

    if (warp == 0) {
        if (lane == 0) {  // ---------------- TMA producer
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
        if (lane == 0) {  // ---------------- MMA issuer
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
        constexpr int COLS_PER_WARP = BN / 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const int row = m_blk * BM + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            long long out_row = row;
            bool store_ok = row_ok;
            if (kSeq) {
                const int sq = row / p.seq_in, r = row - sq * p.seq_in;
                store_ok = row_ok && r < p.seq_out;
                out_row = (long long)sq * p.seq_out + r;
            }
            // LayerNorm row statistics (of the A row for LN-in modes, of the residual-stream row for the residual mode):
            // fetched before the accumulator wait so the loads overlap the tile's MMAs
            float rs = 1.f, nmr = 0.f;              // rstd, -mean * rstd
            if ((kLnIn || EPI == EPI_RES_LN_BF16_STATS) && row_ok) {
                float mean, rstd;
                ln_row_stats(p.stats_in + (size_t)row * LN_PARTIALS, p.inv_d, p.eps, mean, rstd);
                rs = rstd; nmr = -mean * rstd;
            }
            float st_sum = 0.f, st_sq = 0.f;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP; c += 32) {
                const int col0 = half * COLS_PER_WARP + c;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + col0, v);
                tmem_ld_wait();
                const int n0 = n_blk * BN + col0;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                    if (kLnIn) {   // rstd*acc + (-mean*rstd)*u + c
                        const float4 u4 = __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j));
                        f[j + 0] = fmaf(rs, __uint_as_float(v[j + 0]), fmaf(nmr, u4.x, b4.x));
                        f[j + 1] = fmaf(rs, __uint_as_float(v[j + 1]), fmaf(nmr, u4.y, b4.y));
                        f[j + 2] = fmaf(rs, __uint_as_float(v[j + 2]), fmaf(nmr, u4.z, b4.z));
                        f[j + 3] = fmaf(rs, __uint_as_float(v[j + 3]), fmaf(nmr, u4.w, b4.w));
                    } else {
                        f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
                        f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                        f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                        f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                    }
                }
                if (kGelu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                }
                if (EPI == EPI_BIAS_RES_F32 || EPI == EPI_RES_LN_BF16_STATS) {
                    if (row_ok) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + n0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 r4 = __ldg(rp + j);
                            const uint32_t w[4] = {r4.x, r4.y, r4.z, r4.w};
                            if (EPI == EPI_RES_LN_BF16_STATS) {   // + ((y - mean) * rstd) * gamma   (beta is folded into p.bias)
                                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j * 8));
                                const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.vec2 + n0 + j * 8 + 4));
                                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    f[j * 8 + 2 * t + 0] = fmaf(fmaf(__uint_as_float(w[t] << 16), rs, nmr), g[2 * t], f[j * 8 + 2 * t + 0]);
                                    f[j * 8 + 2 * t + 1] = fmaf(fmaf(__uint_as_float(w[t] & 0xffff0000u), rs, nmr), g[2 * t + 1], f[j * 8 + 2 * t + 1]);
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
                            st_sum += a + b;
                            st_sq = fmaf(a, a, fmaf(b, b, st_sq));
                        }
                    }
                    if (store_ok) {
                        uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (size_t)out_row * p.ldo + n0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                    }
                } else if (store_ok) {
                    float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (size_t)out_row * p.ldo + n0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                }
            }
            if (kStats && row_ok) p.stats_out[(size_t)row * LN_PARTIALS + n_blk * 2 + half] = make_float2(st_sum, st_sq);
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

}  // namespace mb
