// Row-wise kernels of the generator trunk (one warp per 1024-wide row, fp32 math, bf16 activations out):
//   embed_ln_kernel : LFQBert.preprocess_tokens + input_proj + class token + pos_emb + first LayerNorm
//                     (reference bert.py:440-454,482-496)
//   layernorm_kernel: torch.nn.LayerNorm(eps=1e-12) on the fp32 pre-norm sum written by the GEMM epilogue
//                     (reference bert.py:70,139,500)
// Both are HBM-bound: per row they read <= 4 KB and write 2 KB.
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

// x[32] holds the row elements e = 4*(lane + 32*j) + i  (j = 0..7, i = 0..3) for D = 1024
template <int D>
__device__ __forceinline__ void ln_store_row(float (&x)[D / 32], const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out,
                                             int lane) {
    constexpr int NV = D / 128;  // float4 per lane
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) s += x[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { const float d = x[i] - mean; ss += d * d; }
    const float var = warp_sum(ss) * (1.0f / D);
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int e = 4 * (lane + 32 * j);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + e));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + e));
        const float y0 = (x[4 * j + 0] - mean) * rstd * g.x + b.x;
        const float y1 = (x[4 * j + 1] - mean) * rstd * g.y + b.y;
        const float y2 = (x[4 * j + 2] - mean) * rstd * g.z + b.z;
        const float y3 = (x[4 * j + 3] - mean) * rstd * g.w + b.w;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(y2, y3);
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&h0);
        w.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(out + e) = w;
    }
}

template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       __nv_bfloat16* __restrict__ out, int rows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* r = in + (size_t)row * D;
    float x[D / 32];
#pragma unroll
    for (int j = 0; j < D / 128; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(r + 4 * (lane + 32 * j)));
        x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
    ln_store_row<D>(x, gamma, beta, eps, out + (size_t)row * D, lane);
}

// tokens   int64 [n_token_rows, seq_len, splits]; sequence n reads tokens[n % n_token_rows] (the sampler's CFG double batch
//          shares one token tensor between the conditional and unconditional halves, sampling.py:84-88)
// labels   int64 [n_label_rows]; sequence n uses labels[n % n_label_rows], replaced by nclass when drop[n] != 0
//          (drop == nullptr: every label dropped -- the reference's drop_label_mask=None quirk, bert.py:484)
// w_in_t   fp32 [bits, D] = input_proj.weight transposed;  pos fp32 [seq_len+1, D];  class_emb fp32 [nclass+1, D]
template <int D>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int64_t* __restrict__ tokens, int n_token_rows, const int64_t* __restrict__ labels, int n_label_rows,
                const uint8_t* __restrict__ drop, int n_seq, int seq_len, int splits, int eff_bits, int nclass,
                const float* __restrict__ w_in_t, const float* __restrict__ b_in, const float* __restrict__ class_emb,
                const float* __restrict__ pos, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                __nv_bfloat16* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long rows = (long long)n_seq * (seq_len + 1);
    if (row >= rows) return;
    const int n = (int)(row / (seq_len + 1)), s = (int)(row - (long long)n * (seq_len + 1));
    float x[D / 32];
    const float* prow = pos + (size_t)s * D;
    if (s < seq_len) {
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
    ln_store_row<D>(x, gamma, beta, eps, out + (size_t)row * D, lane);
}

}  // namespace mb
