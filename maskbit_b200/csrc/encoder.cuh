// Tokenizer encode path, the pieces that are not tensor-core convolutions
// (reference autoencoder.py:230-286 ConvEncoder, lookup_free.py:46-94 LookupFreeQuantizer.forward):
//   enc_conv_in_kernel   conv_in 3x3 (3 -> C0, no bias) from the fp32 NCHW image to fp32 NHWC
//   enc_conv_out_kernel  norm_out-apply + SiLU + conv_out 1x1 (C -> bits, + bias) -> latents z (fp32 NCHW) and the LFQ
//                        token = sum_k [z_k > 0] << k  (convert_bits_to_indices, bit k <-> 2^k)
// Everything between the two (ResidualBlocks, stride-2 down convs) runs on conv_tcgen05_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

// img fp32 [B,3,H,W]; w_t fp32 [27][C0] (tap-major: (cin*9 + ky*3 + kx) -> row); out fp32 NHWC [B,H,W,C0]
__global__ void __launch_bounds__(256)
enc_conv_in_kernel(const float* __restrict__ img, const float* __restrict__ w_t, float* __restrict__ out, int B, int H, int W, int C0) {
    extern __shared__ float ws[];                       // [27][C0]
    for (int i = threadIdx.x; i < 27 * C0; i += blockDim.x) ws[i] = w_t[i];
    __syncthreads();
    const int c4n = C0 >> 2;
    const long long total = (long long)B * H * W * c4n;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4 = (int)(idx % c4n);
    const long long pix = idx / c4n;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((long long)W * H));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const size_t hw = (size_t)H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
        const float* plane = img + ((size_t)n * 3 + ci) * hw;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = x + kx - 1;
                if (ix < 0 || ix >= W) continue;
                const float v = __ldg(plane + (size_t)iy * W + ix);
                const float4 wv = *reinterpret_cast<const float4*>(ws + (ci * 9 + ky * 3 + kx) * C0 + c4 * 4);
                a0 = fmaf(v, wv.x, a0); a1 = fmaf(v, wv.y, a1); a2 = fmaf(v, wv.z, a2); a3 = fmaf(v, wv.w, a3);
            }
        }
    }
    *reinterpret_cast<float4*>(out + pix * C0 + c4 * 4) = make_float4(a0, a1, a2, a3);
}

// in fp32 NHWC [B, HW, C]; scale/shift [B][C] (GroupNorm-apply); w fp32 [bits][C]; bias [bits]
// z fp32 NCHW [B, bits, HW] (or nullptr); indices int64 [B, HW] (or nullptr).  One warp per pixel.
__global__ void __launch_bounds__(256)
enc_conv_out_kernel(const float* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ z, int64_t* __restrict__ indices,
                    int B, int HW, int C, int bits) {
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= (long long)B * HW) return;
    const int n = (int)(pix / HW), r = (int)(pix - (long long)n * HW);
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
        float4 v = __ldg(reinterpret_cast<const float4*>(in + pix * C + c));
        const float4 a = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * C + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * C + c));
        float t;
        t = fmaf(v.x, a.x, b.x); v.x = t / (1.0f + __expf(-t));
        t = fmaf(v.y, a.y, b.y); v.y = t / (1.0f + __expf(-t));
        t = fmaf(v.z, a.z, b.z); v.z = t / (1.0f + __expf(-t));
        t = fmaf(v.w, a.w, b.w); v.w = t / (1.0f + __expf(-t));
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k < bits) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C + c));
                acc[k] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[k]))));
            }
        }
    }
    int64_t tok = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k < bits) {
            float s = acc[k];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            s += bias[k];
            if (lane == 0 && z) z[((size_t)n * bits + k) * HW + r] = s;
            if (s > 0.0f) tok |= (int64_t)1 << k;                       // lookup_free.py:60,126-127
        }
    }
    if (lane == 0 && indices) indices[pix] = tok;
}

}  // namespace mb
