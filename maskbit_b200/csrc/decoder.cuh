// ConvVQModel.decode_tokens kernels (reference conv_vqgan.py:98-112, autoencoder.py:7-96,187-227,358-423) other than the
// tensor-core convolutions (conv_tcgen05.cuh).  Activations are fp32 NHWC in HBM.
//
//   conv_in_tokens_kernel : LFQ unpack (token -> +-1 bits, lookup_free.py:108-111) fused into conv_in 3x3 (bits -> C)
//   gn_partial / gn_finalize : GroupNorm(32, eps 1e-6) statistics -> per-(image, channel) scale/shift
//   conv_out_kernel       : norm_out + SiLU + conv_out 3x3 (C -> 3), writes fp32 NCHW like the reference
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

// ------------------------------------------------------------------------------------------------ conv_in
// tokens int64 [B, P*P] full codebook indices; w_t fp32 [9][bits][C] (tap-major), bias [C]; out fp32 NHWC [B,P,P,C]
__global__ void __launch_bounds__(256)
conv_in_tokens_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ w_t, const float* __restrict__ bias,
                      float* __restrict__ out, int B, int P, int bits, int C) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * P * P * C;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const long long pix = idx / C;
    const int x = (int)(pix % P), y = (int)((pix / P) % P), n = (int)(pix / ((long long)P * P));
    float acc = bias[c];
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = y + dy - 1;
        if (iy < 0 || iy >= P) continue;
        for (int dx = 0; dx < 3; ++dx) {
            const int ix = x + dx - 1;
            if (ix < 0 || ix >= P) continue;
            const int64_t tok = tokens[((size_t)n * P + iy) * P + ix];
            const float* w = w_t + (size_t)((dy * 3 + dx) * bits) * C + c;
            for (int k = 0; k < bits; ++k) {
                const float wv = __ldg(w + (size_t)k * C);
                acc += ((tok >> k) & 1) ? wv : -wv;
            }
        }
    }
    out[idx] = acc;
}

// ------------------------------------------------------------------------------------------------ GroupNorm stats
// x fp32 NHWC [B, HW, C]; partial[(n*chunks + chunk)*32 + g] = (sum, sumsq) over the chunk's pixels, as double2
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, double2* __restrict__ partial, int HW, int C, int chunks) {
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int cpg = C / 32;
    const int c4 = C / 4;                                   // float4 per pixel
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    __shared__ double s_sum[32], s_sq[32];
    if (threadIdx.x < 32) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
    __syncthreads();
    // thread owns a fixed float4 column (so a fixed group) when 256 % c4 == 0 or c4 % 256 == 0; general path otherwise
    const float4* base = reinterpret_cast<const float4*>(x + (size_t)n * HW * C);
    const long long total4 = (long long)(p1 - p0) * c4;
    float s = 0.f, q = 0.f;
    int cur_g = -1;
    for (long long i = threadIdx.x; i < total4; i += blockDim.x) {
        const int col4 = (int)(i % c4);
        const int g = (col4 * 4) / cpg;
        if (g != cur_g) {
            if (cur_g >= 0) { atomicAdd(&s_sum[cur_g], (double)s); atomicAdd(&s_sq[cur_g], (double)q); }
            cur_g = g; s = 0.f; q = 0.f;
        }
        const float4 v = __ldg(base + (size_t)p0 * c4 + i);
        if (cpg >= 4) {
            s += (v.x + v.y) + (v.z + v.w);
            q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        } else {  // cpg == 2 or 1: a float4 spans several groups (not used by the shipped configs, C >= 128)
            const float vv[4] = {v.x, v.y, v.z, v.w};
            for (int t = 0; t < 4; ++t) {
                const int gg = (col4 * 4 + t) / cpg;
                atomicAdd(&s_sum[gg], (double)vv[t]); atomicAdd(&s_sq[gg], (double)vv[t] * vv[t]);
            }
            cur_g = -1;
        }
    }
    if (cur_g >= 0) { atomicAdd(&s_sum[cur_g], (double)s); atomicAdd(&s_sq[cur_g], (double)q); }
    __syncthreads();
    if (threadIdx.x < 32) partial[((size_t)n * chunks + chunk) * 32 + threadIdx.x] = make_double2(s_sum[threadIdx.x], s_sq[threadIdx.x]);
}

// The same partials from the per-box sums the convolution epilogue left behind (conv_tcgen05.cuh, ConvTcParams::gn_part):
// part float2 [B * HW / 32][32]; one block per image, thread = (group, 1 of 8 interleaved box subsets), fp64 accumulation
// in a fixed order.  Reads HW/32 * 256 B per image instead of the HW * C * 4 B activation.
__global__ void __launch_bounds__(256)
gn_reduce_kernel(const float2* __restrict__ part, double2* __restrict__ partial, int boxes_per_img) {
    const int n = blockIdx.x, g = threadIdx.x & 31, sub = threadIdx.x >> 5;
    const float2* base = part + (size_t)n * boxes_per_img * 32 + g;
    double s = 0.0, q = 0.0;
    for (int b = sub; b < boxes_per_img; b += 8) {
        const float2 v = __ldg(base + (size_t)b * 32);
        s += (double)v.x; q += (double)v.y;
    }
    __shared__ double sh_s[8][32], sh_q[8][32];
    sh_s[sub][g] = s; sh_q[sub][g] = q;
    __syncthreads();
    if (threadIdx.x < 32) {
        double ts = 0.0, tq = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ts += sh_s[k][g]; tq += sh_q[k][g]; }
        partial[(size_t)n * 32 + g] = make_double2(ts, tq);
    }
}

// scale[n][c] = rstd * gamma[c]; shift[n][c] = beta[c] - mean * rstd * gamma[c]
__global__ void gn_finalize_kernel(const double2* __restrict__ partial, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift,
                                   int HW, int C, int chunks, float eps) {
    const int n = blockIdx.x;
    const int cpg = C / 32;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        double s = 0.0, q = 0.0;
        for (int k = 0; k < chunks; ++k) {
            const double2 v = partial[((size_t)n * chunks + k) * 32 + g];
            s += v.x; q += v.y;
        }
        const double cnt = (double)HW * cpg;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float a = rstd * gamma[c];
        scale[(size_t)n * C + c] = a;
        shift[(size_t)n * C + c] = beta[c] - (float)mean * a;
    }
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------ conv_out
// in fp32 NHWC [B,H,W,C] -> GroupNorm-apply + SiLU -> 3x3 conv to 3 channels (+bias) -> fp32 NCHW [B,3,H,W].
// w fp32 [9][C][4] (4th lane zero).  One CTA = 16 rows x 64 columns of outputs, one thread = 4 consecutive pixels of a row
// (12 accumulators); channels stream through shared memory 8 at a time as PLANES [c][row][x] (x contiguous), so a thread reads
// its 6 inputs of a tap row with one 128-bit + one 64-bit load and every weight float4 is reused for 4 pixels:
// 15 shared-memory loads per 108 FMAs (one pixel per thread with [pixel][c] tiles was 2 loads per 3 FMAs and LDS-bound,
// 1.36 ms per 32 images against 0.16 ms of HBM time: profiles/r02_launches_B256_T1.txt).
constexpr int CO_TY = 16, CO_TX = 64, CO_CC = 8;
constexpr int CO_PITCH = CO_TX + 4;                        // 66 columns with the halo, padded to a multiple of 4 floats
constexpr int CO_PLANE = (CO_TY + 2) * CO_PITCH + 4;       // +4: planes 4 apart land 16 banks apart (conflict-free fill stores)
__global__ void __launch_bounds__(256)
conv_out_kernel(const float* __restrict__ in, const float* __restrict__ gn_scale, const float* __restrict__ gn_shift,
                const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int H, int W, int C) {
    __shared__ __align__(16) float tile[CO_CC * CO_PLANE];
    __shared__ float4 wsm[9][CO_CC];
    const int n = blockIdx.z, ty0 = blockIdx.y * CO_TY, tx0 = blockIdx.x * CO_TX;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    float acc[4][3];
#pragma unroll
    for (int px = 0; px < 4; ++px) { acc[px][0] = bias[0]; acc[px][1] = bias[1]; acc[px][2] = bias[2]; }
    for (int c0 = 0; c0 < C; c0 += CO_CC) {
        __syncthreads();
        // fill: lanes (pixel, half) -> the pixel's 32-byte sector of 8 channels is read by two adjacent lanes
        for (int i = threadIdx.x; i < (CO_TY + 2) * (CO_TX + 2) * 2; i += 256) {
            const int c4 = i & 1, pix = i >> 1;
            const int row = pix / (CO_TX + 2), sx = pix - row * (CO_TX + 2);
            const int yy = ty0 + row - 1, xx = tx0 + sx - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)n * H + yy) * W + xx) * C + c0 + c4 * 4));
                const float4 a = __ldg(reinterpret_cast<const float4*>(gn_scale + (size_t)n * C + c0 + c4 * 4));
                const float4 b = __ldg(reinterpret_cast<const float4*>(gn_shift + (size_t)n * C + c0 + c4 * 4));
                v.x = silu(fmaf(v.x, a.x, b.x)); v.y = silu(fmaf(v.y, a.y, b.y));
                v.z = silu(fmaf(v.z, a.z, b.z)); v.w = silu(fmaf(v.w, a.w, b.w));
            }
            float* dst = tile + (c4 * 4) * CO_PLANE + row * CO_PITCH + sx;
            dst[0] = v.x; dst[CO_PLANE] = v.y; dst[2 * CO_PLANE] = v.z; dst[3 * CO_PLANE] = v.w;
        }
        for (int i = threadIdx.x; i < 9 * CO_CC; i += 256) {
            const int tap = i / CO_CC, c = i % CO_CC;
            wsm[tap][c] = __ldg(reinterpret_cast<const float4*>(w + ((size_t)tap * C + c0 + c) * 4));
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CO_CC; ++c) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const float* rp = tile + c * CO_PLANE + (ly + dy) * CO_PITCH + 4 * lx;
                const float4 v0 = *reinterpret_cast<const float4*>(rp);
                const float2 v1 = *reinterpret_cast<const float2*>(rp + 4);
                const float v[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 ww = wsm[dy * 3 + dx][c];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        acc[px][0] = fmaf(v[px + dx], ww.x, acc[px][0]);
                        acc[px][1] = fmaf(v[px + dx], ww.y, acc[px][1]);
                        acc[px][2] = fmaf(v[px + dx], ww.z, acc[px][2]);
                    }
                }
            }
        }
    }
    const int y = ty0 + ly, x = tx0 + 4 * lx;
    if (y < H && x < W) {
        const size_t hw = (size_t)H * W;
        float* o = out + (size_t)n * 3 * hw + (size_t)y * W + x;
        if (x + 3 < W && (W & 3) == 0) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                *reinterpret_cast<float4*>(o + ch * hw) = make_float4(acc[0][ch], acc[1][ch], acc[2][ch], acc[3][ch]);
        } else {
            for (int px = 0; px < 4 && x + px < W; ++px) { o[px] = acc[px][0]; o[hw + px] = acc[px][1]; o[2 * hw + px] = acc[px][2]; }
        }
    }
}

// clamp(0,1) * 255 -> uint8 NHWC (truncating cast), reference scripts/eval_maskbit.py:134-135
__global__ void postprocess_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int B, int H, int W) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * H * W * 3;
    if (idx >= total) return;
    const int c = (int)(idx % 3);
    const long long pix = idx / 3;
    const long long hw = (long long)H * W;
    const long long n = pix / hw, r = pix - n * hw;
    float v = img[(n * 3 + c) * hw + r];
    v = fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f;
    out[idx] = (uint8_t)v;
}

}  // namespace mb
