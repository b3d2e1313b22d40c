// Input stage of the generator trunk (one warp per 1024-wide row, fp32 math):
//   embed_kernel : LFQBert.preprocess_tokens + input_proj + class token + pos_emb   (reference bert.py:440-454,482-493)
// It writes the PRE-LayerNorm sum y0 as bf16 together with the row's (sum, sum of squares) statistics; the first LayerNorm
// (bert.py:496) is folded into the consumers (layer 0's QKV GEMM and out-projection residual, see gemm_tcgen05.cuh).
// HBM-bound: per row it writes 2 KB + 64 B and reads <= 12 weight rows that stay in L1/L2.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// tokens   int64 [n_token_rows, seq_len, splits]; sequence n reads tokens[n % n_token_rows] (the sampler's CFG double batch
//          shares one token tensor between the conditional and unconditional halves, sampling.py:84-88)
// labels   int64 [n_label_rows]; sequence n uses labels[n % n_label_rows], replaced by nclass when drop[n] != 0
//          (drop == nullptr: every label dropped -- the reference's drop_label_mask=None quirk, bert.py:484)
// w_in_t   fp32 [bits, D] = input_proj.weight transposed;  pos fp32 [seq_len+1, D];  class_emb fp32 [nclass+1, D]
// y        bf16 [rows, D] pre-LayerNorm sum;  stats float2 [rows][n_partials]: slot 0 = (sum, sumsq) of the stored row, rest 0
// tok_tables  nullptr for LFQBert.  Bert (embedding-table generator, bert.py:313-315): fp32 [splits][V+1][D], the row of token t of
//          split g is added (t = V is the mask token's own learned row); w_in_t / b_in are unused.
// ln_g/ln_b  nullptr for the post-norm trunk.  Pre-norm trunk (use_prenorm, bert.py:496 with :106-123): the residual stream carries
//          first_layer's LayerNorm OUTPUT, so it is applied here (fp32 statistics of the fp32 row, eps 1e-12) before the store.
template <int D>
__global__ void __launch_bounds__(256)
embed_kernel(const int64_t* __restrict__ tokens, int n_token_rows, const int64_t* __restrict__ labels, int n_label_rows,
             const uint8_t* __restrict__ drop, int n_seq, int seq_len, int splits, int eff_bits, int nclass,
             const float* __restrict__ w_in_t, const float* __restrict__ b_in, const float* __restrict__ class_emb,
             const float* __restrict__ pos, __nv_bfloat16* __restrict__ y, float2* __restrict__ stats, int n_partials,
             const float* __restrict__ ln_g = nullptr, const float* __restrict__ ln_b = nullptr,
             const float* __restrict__ tok_tables = nullptr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long rows = (long long)n_seq * (seq_len + 1);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // programmatic dependent launch (ptx.cuh): the tokens come from
    asm volatile("griddepcontrol.wait;" ::: "memory");                 // the previous step's select kernel
    if (row >= rows) return;
    const int n = (int)(row / (seq_len + 1)), s = (int)(row - (long long)n * (seq_len + 1));
    float x[D / 32];   // row elements e = 4*(lane + 32*j) + i  (j = 0..D/128-1, i = 0..3)
    const float* prow = pos + (size_t)s * D;
    if (s < seq_len && tok_tables != nullptr) {
        const int64_t* trow = tokens + ((size_t)(n % n_token_rows) * seq_len + s) * splits;
        const int64_t rows_per_table = ((int64_t)1 << eff_bits) + 1;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) x[i] = 0.f;
        for (int g = 0; g < splits; ++g) {
            int64_t tok = trow[g];
            tok = tok < 0 ? 0 : (tok >= rows_per_table ? rows_per_table - 1 : tok);     // stay inside the table
            const float* w = tok_tables + ((size_t)g * rows_per_table + tok) * D;
#pragma unroll
            for (int j = 0; j < D / 128; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(w + 4 * (lane + 32 * j)));
                x[4 * j] += v.x; x[4 * j + 1] += v.y; x[4 * j + 2] += v.z; x[4 * j + 3] += v.w;
            }
        }
    } else if (s < seq_len) {
#pragma unroll
        for (int j = 0; j < D / 128; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(b_in + 4 * (lane + 32 * j)));
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
        const int64_t* trow = tokens + ((size_t)(n % n_token_rows) * seq_len + s) * splits;
        const int64_t mask_token = (int64_t)1 << eff_bits;
        for (int g = 0; g < splits; ++g) {
            const int64_t tok = trow[g];
            if (tok == mask_token) continue;                        // masked group contributes 0 (bert.py:452)
            for (int k = 0; k < eff_bits; ++k) {
                const float sgn = ((tok >> k) & 1) ? 1.0f : -1.0f;  // bit k <-> 2^k, coded +-1 (bert.py:450-451)
                const float* w = w_in_t + (size_t)(g * eff_bits + k) * D;
#pragma unroll
                for (int j = 0; j < D / 128; ++j) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(w + 4 * (lane + 32 * j)));
                    x[4 * j] += sgn * v.x; x[4 * j + 1] += sgn * v.y; x[4 * j + 2] += sgn * v.z; x[4 * j + 3] += sgn * v.w;
                }
            }
        }
    } else {
        int64_t cls = labels[n % n_label_rows];
        if (drop == nullptr || drop[n]) cls = nclass;
        const float* c = class_emb + (size_t)cls * D;
#pragma unroll
        for (int j = 0; j < D / 128; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(c + 4 * (lane + 32 * j)));
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
    }
#pragma unroll
    for (int j = 0; j < D / 128; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(prow + 4 * (lane + 32 * j)));
        x[4 * j] += v.x; x[4 * j + 1] += v.y; x[4 * j + 2] += v.z; x[4 * j + 3] += v.w;
    }
    if (ln_g != nullptr) {   // two-pass LayerNorm in registers
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) s1 += x[i];
        const float mean = warp_sum(s1) * (1.0f / D);
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) { const float d = x[i] - mean; s2 = fmaf(d, d, s2); }
        const float rstd = rsqrtf(warp_sum(s2) * (1.0f / D) + 1e-12f);
#pragma unroll
        for (int j = 0; j < D / 128; ++j) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g + 4 * (lane + 32 * j)));
            const float4 b = __ldg(reinterpret_cast<const float4*>(ln_b + 4 * (lane + 32 * j)));
            x[4 * j] = fmaf((x[4 * j] - mean) * rstd, g.x, b.x);
            x[4 * j + 1] = fmaf((x[4 * j + 1] - mean) * rstd, g.y, b.y);
            x[4 * j + 2] = fmaf((x[4 * j + 2] - mean) * rstd, g.z, b.z);
            x[4 * j + 3] = fmaf((x[4 * j + 3] - mean) * rstd, g.w, b.w);
        }
    }
    float sum = 0.f, sq = 0.f;
    __nv_bfloat16* orow = y + (size_t)row * D;
#pragma unroll
    for (int j = 0; j < D / 128; ++j) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(x[4 * j], x[4 * j + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(x[4 * j + 2], x[4 * j + 3]);
        const float a = __low2float(h0), b = __high2float(h0), c = __low2float(h1), d = __high2float(h1);
        sum += (a + b) + (c + d);                                   // statistics of the values as stored
        sq = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, sq))));
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&h0);
        w.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(orow + 4 * (lane + 32 * j)) = w;
    }
    sum = warp_sum(sum); sq = warp_sum(sq);
    if (lane < n_partials) stats[(size_t)row * n_partials + lane] = lane == 0 ? make_float2(sum, sq) : make_float2(0.f, 0.f);
}

// Bert's per-position logit bias (bert.py:333): logits fp32 [n_seq, seq_len, splits, V] += bias[splits][seq_len][V]
__global__ void add_pos_bias_kernel(float* __restrict__ logits, const float* __restrict__ bias, long long total, int seq_len, int splits, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int v = (int)(i % V);
    const int g = (int)((i / V) % splits);
    const int s = (int)((i / ((long long)V * splits)) % seq_len);
    logits[i] += __ldg(bias + ((size_t)g * seq_len + s) * V + v);
}

}  // namespace mb
