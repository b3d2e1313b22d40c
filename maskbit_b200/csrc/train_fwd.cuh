// Forward half of the generator's training step (reference scripts/train_maskbit.py:362-380):
//   split_factorized_tokens  modeling/modules/factorization.py:27-46   full index -> per-group tokens          (integer)
//   get_mask_tokens          modeling/modules/masking.py:7-38          per-sample masking ratio, random re-mask (integer)
//   MLMLoss.forward          modeling/modules/losses.py:289-339        label-smoothed cross entropy + accuracies over all and
//                                                                      over the masked slots                    (fp32)
// HBM-bound kernels: tokens are read once; the loss reads the logits once (one warp per row of V logits).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mb {

// tokens [n] -> out [n, splits]: out[i, g] = (tokens[i] >> (g * shift)) & (2^shift - 1)
__global__ void split_tokens_kernel(const int64_t* __restrict__ tokens, int64_t* __restrict__ out, size_t n, int splits, int shift) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t t = tokens[i], bm = ((int64_t)1 << shift) - 1;
    for (int g = 0; g < splits; ++g) out[i * splits + g] = (t & (bm << (g * shift))) >> (g * shift);   // factorization.py:42-45
}

// masked[b, s] = u[b, s] < val_to_mask[b] ? mask_token : tokens[b, s];  mask[b, s] = that predicate  (masking.py:34-37)
__global__ void mask_tokens_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ u, const float* __restrict__ val_to_mask,
                                   int64_t mask_token, int64_t* __restrict__ masked, uint8_t* __restrict__ mask, int B, int slots) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * slots) return;
    const bool m = u[i] < val_to_mask[i / slots];
    masked[i] = m ? mask_token : tokens[i];
    mask[i] = m ? 1 : 0;
}

constexpr int MLM_PARTIALS = 7;      // nll, smooth, correct | masked: nll, smooth, correct, count
constexpr int MLM_BLOCKS = 592;      // 4 x 148 SMs

// One warp per row.  log-softmax the way torch.nn.CrossEntropyLoss evaluates it in fp32: x - max - log(sum exp(x - max));
//   nll = -logp[target],  smooth = -mean_j logp[j]  (label smoothing term),  hit = argmax (first maximum) == target
// Row results are accumulated per warp in double, in row order; block and grid sums are taken in fixed order -> deterministic.
__global__ void __launch_bounds__(256)
mlm_loss_partial_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, const uint8_t* __restrict__ masks,
                        long long rows, int V, double* __restrict__ partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double acc[MLM_PARTIALS];
#pragma unroll
    for (int k = 0; k < MLM_PARTIALS; ++k) acc[k] = 0.0;
    for (long long r = (long long)blockIdx.x * wpb + warp; r < rows; r += (long long)gridDim.x * wpb) {
        const float* x = logits + r * V;
        float mx = -CUDART_INF_F; int am = 0;
        for (int j = lane; j < V; j += 32) { const float v = __ldg(x + j); if (v > mx) { mx = v; am = j; } }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oi = __shfl_xor_sync(0xffffffffu, am, o);
            if (ov > mx || (ov == mx && oi < am)) { mx = ov; am = oi; }
        }
        float se = 0.f, sx = 0.f;
        for (int j = lane; j < V; j += 32) { const float v = __ldg(x + j); se += expf(v - mx); sx += v; }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); }
        if (lane == 0) {
            const int64_t t = targets[r];
            const float lse = mx + logf(se);
            const float nll = lse - __ldg(x + t);
            const float smooth = lse - sx / (float)V;
            const double hit = am == (int)t ? 1.0 : 0.0;
            acc[0] += nll; acc[1] += smooth; acc[2] += hit;
            if (masks[r]) { acc[3] += nll; acc[4] += smooth; acc[5] += hit; acc[6] += 1.0; }
        }
    }
    __shared__ double red[8][MLM_PARTIALS];
    if (lane == 0)
        for (int k = 0; k < MLM_PARTIALS; ++k) red[warp][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < MLM_PARTIALS) {
        double s = 0.0;
        for (int w = 0; w < wpb; ++w) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * MLM_PARTIALS + threadIdx.x] = s;
    }
}

// out[0] = mlm_loss, out[1] = correct_tokens, out[2] = masked_token_loss, out[3] = masked_correct_tokens  (losses.py:318-337)
__global__ void mlm_loss_final_kernel(const double* __restrict__ partial, int n_blocks, long long rows, int splits, float eps,
                                      int sum_splits, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[MLM_PARTIALS];
    for (int k = 0; k < MLM_PARTIALS; ++k) s[k] = 0.0;
    for (int b = 0; b < n_blocks; ++b)
        for (int k = 0; k < MLM_PARTIALS; ++k) s[k] += partial[(size_t)b * MLM_PARTIALS + k];
    const double scale = sum_splits ? (double)splits : 1.0;
    const double n = (double)rows, nm = s[6];
    out[0] = (float)(((1.0 - eps) * s[0] + eps * s[1]) / n * scale);
    out[1] = (float)pow(s[2] / n, (double)splits);
    out[2] = (float)(((1.0 - eps) * s[3] + eps * s[4]) / nm * scale);      // no masked slot: 0 / 0 = NaN, like torch's mean of nothing
    out[3] = (float)pow(s[5] / nm, (double)splits);
}

}  // namespace mb
