// ConvVQModel.decode_tokens kernels (reference conv_vqgan.py:98-112, autoencoder.py:7-96,187-227,358-423).
// Activations are fp32 NHWC in HBM.  The 1e-3-abs pixel tolerance rules out plain bf16 tensor-core convs (8e-2
// measured), so the implicit-GEMM conv splits both operands into bf16 hi + lo and issues three
// tensor-core MMAs per tile (hi*hi + hi*lo + lo*hi, fp32 accumulate): ~2^-16 relative error per product.
//
//   conv_in_tokens_kernel : LFQ unpack (token -> +-1 bits, lookup_free.py:108-111) fused into conv_in 3x3 (bits -> C)
//   gn_partial / gn_finalize : GroupNorm(32, eps 1e-6) statistics -> per-(image, channel) scale/shift
//   conv_igemm_kernel     : 3x3 / 1x1 conv as implicit GEMM; fused on load: GroupNorm-apply + SiLU, nearest x2
//                           upsample (index >> 1); fused on store: bias, residual add
//   conv_out_kernel       : norm_out + SiLU + conv_out 3x3 (C -> 3), writes fp32 NCHW like the reference
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "attention.cuh"  // ldsm_x4, mma_bf16_16816, pack_bf16

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

// ------------------------------------------------------------------------------------------------ implicit-GEMM conv
struct ConvParams {
    const float* in;        // fp32 NHWC [B, Hin, Win, Cin]   (Hin = H >> up)
    float* out;             // fp32 NHWC [B, H, W, Cout]
    const __nv_bfloat16* w_hi;  // [Cout][taps*Cin]  (k = tap*Cin + c)
    const __nv_bfloat16* w_lo;
    const float* bias;      // [Cout] or nullptr
    const float* residual;  // fp32 NHWC [B,H,W,Cout] or nullptr
    const float* gn_scale;  // [B][Cin] or nullptr (no GroupNorm+SiLU on load)
    const float* gn_shift;
    int B, H, W, Cin, Cout; // H, W = output size
    int taps;               // 9 (3x3, SAME pad 1) or 1 (1x1)
    int up;                 // 1: input is half resolution, nearest-upsampled on load (autoencoder.py:224)
    int logW, logH;
};

constexpr int CV_BM = 128, CV_BN = 128, CV_BK = 32, CV_LDS = 40;   // 40 bf16 = 80 B rows: conflict-free ldmatrix
constexpr int CV_THREADS = 256;
constexpr int CV_TILE_ELEMS = CV_BM * CV_LDS;                      // per array (A_hi, A_lo, B_hi, B_lo)
constexpr int CV_SMEM_BYTES = 2 /*stages*/ * 4 * CV_TILE_ELEMS * 2;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// grid: (ceil(B*H*W / 128), Cout / 128); 8 warps as 4 (pixels) x 2 (channels); warp tile 32 x 64
__global__ void __launch_bounds__(CV_THREADS, 2) conv_igemm_kernel(ConvParams p) {
    extern __shared__ __align__(16) uint8_t cv_smem[];
    __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(cv_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const long long npix = (long long)p.B * p.H * p.W;
    const long long pix0 = (long long)blockIdx.x * CV_BM;
    const int co0 = blockIdx.y * CV_BN;
    const int kchunks = p.Cin / CV_BK;
    const int nk = p.taps * kchunks;
    const int Hin = p.H >> p.up, Win = p.W >> p.up;
    const int Ktot = p.taps * p.Cin;

    // A loader mapping: thread -> (pixel row pr + 32*i, float4 column cq) ; 8 threads cover one pixel's 32 channels
    const int cq = tid & 7, pr = tid >> 3;
    int py[4], px[4], pn[4]; bool pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long pp = pix0 + pr + 32 * i;
        pv[i] = pp < npix;
        const long long q = pv[i] ? pp : 0;
        px[i] = (int)(q & (p.W - 1));
        py[i] = (int)((q >> p.logW) & (p.H - 1));
        pn[i] = (int)(q >> (p.logW + p.logH));
    }
    float4 areg[4];
    uint32_t aok = 0;
    // raw loads only (kept in flight across the MMAs of the current chunk); GroupNorm+SiLU is applied in store_a
    auto load_a = [&](int kc) {
        const int tap = kc / kchunks, c0 = (kc - tap * kchunks) * CV_BK;
        const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
        aok = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int iy = py[i] + dy, ix = px[i] + dx;
            const bool ok = pv[i] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                const size_t src = (((size_t)pn[i] * Hin + (iy >> p.up)) * Win + (ix >> p.up)) * p.Cin + c0 + cq * 4;
                v = __ldg(reinterpret_cast<const float4*>(p.in + src));
                aok |= 1u << i;
            }
            areg[i] = v;
        }
    };
    auto store_a = [&](int kc, int stage) {
        __nv_bfloat16* a_hi = sm + stage * 4 * CV_TILE_ELEMS;
        __nv_bfloat16* a_lo = a_hi + CV_TILE_ELEMS;
        const int tap = kc / kchunks, c0 = (kc - tap * kchunks) * CV_BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v = areg[i];
            if (p.gn_scale && ((aok >> i) & 1)) {   // zero padding applies to the activated tensor (Conv2dSame pads its input)
                const float4 a = __ldg(reinterpret_cast<const float4*>(p.gn_scale + (size_t)pn[i] * p.Cin + c0 + cq * 4));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.gn_shift + (size_t)pn[i] * p.Cin + c0 + cq * 4));
                v.x = silu(fmaf(v.x, a.x, b.x)); v.y = silu(fmaf(v.y, a.y, b.y));
                v.z = silu(fmaf(v.z, a.z, b.z)); v.w = silu(fmaf(v.w, a.w, b.w));
            }
            const float f[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                h[t] = __float2bfloat16_rn(f[t]);
                l[t] = __float2bfloat16_rn(f[t] - __bfloat162float(h[t]));
            }
            const int off = (pr + 32 * i) * CV_LDS + cq * 4;
            *reinterpret_cast<uint2*>(a_hi + off) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(a_lo + off) = *reinterpret_cast<const uint2*>(l);
        }
    };
    auto load_b = [&](int kc, int stage) {   // 128 couts x 32 k x {hi, lo}: 4 x 16 B chunks per row per array
        __nv_bfloat16* b_hi = sm + stage * 4 * CV_TILE_ELEMS + 2 * CV_TILE_ELEMS;
        __nv_bfloat16* b_lo = b_hi + CV_TILE_ELEMS;
        const int tap = kc / kchunks, c0 = (kc - tap * kchunks) * CV_BK;
        const size_t kof = (size_t)tap * p.Cin + c0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + 256 * i;          // 0..511
            const int row = idx >> 2, ch = idx & 3;
            const size_t src = (size_t)(co0 + row) * Ktot + kof + ch * 8;
            const uint32_t d_hi = static_cast<uint32_t>(__cvta_generic_to_shared(b_hi + row * CV_LDS + ch * 8));
            const uint32_t d_lo = static_cast<uint32_t>(__cvta_generic_to_shared(b_lo + row * CV_LDS + ch * 8));
            cp_async16(d_hi, p.w_hi + src);
            cp_async16(d_lo, p.w_lo + src);
        }
        cp_async_commit();
    };

    float acc[2][8][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    load_a(0);
    load_b(0, 0);
    store_a(0, 0);
    cp_async_wait0();
    __syncthreads();

    const uint32_t sm_addr = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
    for (int kc = 0; kc < nk; ++kc) {
        const int st = kc & 1;
        if (kc + 1 < nk) { load_a(kc + 1); load_b(kc + 1, st ^ 1); }
        const uint32_t a_hi = sm_addr + (st * 4 * CV_TILE_ELEMS) * 2, a_lo = a_hi + CV_TILE_ELEMS * 2;
        const uint32_t b_hi = a_hi + 2 * CV_TILE_ELEMS * 2, b_lo = b_hi + CV_TILE_ELEMS * 2;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {        // two k16 steps per 32-wide chunk
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                // A 16x16 tile: matrices (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
                const int row = wm * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int col = ks * 16 + (lane >> 4) * 8;
                ldsm_x4(ah[mt], a_hi + (row * CV_LDS + col) * 2);
                ldsm_x4(al[mt], a_lo + (row * CV_LDS + col) * 2);
            }
#pragma unroll
            for (int np = 0; np < 4; ++np) {     // pairs of n-tiles (16 couts)
                // B: matrices (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
                const int row = wn * 64 + np * 16 + (lane & 7) + (lane >> 4) * 8;
                const int col = ks * 16 + ((lane >> 3) & 1) * 8;
                uint32_t bh[4], bl[4];
                ldsm_x4(bh, b_hi + (row * CV_LDS + col) * 2);
                ldsm_x4(bl, b_lo + (row * CV_LDS + col) * 2);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16_16816(acc[mt][2 * np], al[mt], bh[0], bh[1]);       // small terms first
                    mma_bf16_16816(acc[mt][2 * np], ah[mt], bl[0], bl[1]);
                    mma_bf16_16816(acc[mt][2 * np], ah[mt], bh[0], bh[1]);
                    mma_bf16_16816(acc[mt][2 * np + 1], al[mt], bh[2], bh[3]);
                    mma_bf16_16816(acc[mt][2 * np + 1], ah[mt], bl[2], bl[3]);
                    mma_bf16_16816(acc[mt][2 * np + 1], ah[mt], bh[2], bh[3]);
                }
            }
        }
        if (kc + 1 < nk) {
            store_a(kc + 1, st ^ 1);
            cp_async_wait0();
        }
        __syncthreads();
    }

    // epilogue: C fragment rows g / g+8 -> pixels, cols 2t,2t+1 -> couts
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const long long pp = pix0 + wm * 32 + mt * 16 + g + hh * 8;
            if (pp >= npix) continue;
            float* orow = p.out + (size_t)pp * p.Cout + co0 + wn * 64 + 2 * t;
            const float* rrow = p.residual ? p.residual + (size_t)pp * p.Cout + co0 + wn * 64 + 2 * t : nullptr;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float v0 = acc[mt][nt][2 * hh], v1 = acc[mt][nt][2 * hh + 1];
                if (p.bias) { v0 += __ldg(p.bias + co0 + wn * 64 + nt * 8 + 2 * t); v1 += __ldg(p.bias + co0 + wn * 64 + nt * 8 + 2 * t + 1); }
                if (rrow) { const float2 r = __ldg(reinterpret_cast<const float2*>(rrow + nt * 8)); v0 += r.x; v1 += r.y; }
                *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(v0, v1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv_out
// in fp32 NHWC [B,H,W,C] -> GroupNorm-apply + SiLU -> 3x3 conv to 3 channels (+bias) -> fp32 NCHW [B,3,H,W].
// w fp32 [9][C][4] (4th lane zero), 16x16 output tile per CTA, channels streamed through smem 16 at a time.
constexpr int CO_T = 16, CO_CC = 16;
__global__ void __launch_bounds__(256)
conv_out_kernel(const float* __restrict__ in, const float* __restrict__ gn_scale, const float* __restrict__ gn_shift,
                const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int H, int W, int C) {
    __shared__ float tile[(CO_T + 2) * (CO_T + 2)][CO_CC + 1];
    __shared__ float4 wsm[9][CO_CC];
    const int n = blockIdx.z, ty0 = blockIdx.y * CO_T, tx0 = blockIdx.x * CO_T;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    float a0 = bias[0], a1 = bias[1], a2 = bias[2];
    for (int c0 = 0; c0 < C; c0 += CO_CC) {
        __syncthreads();
        for (int i = threadIdx.x; i < (CO_T + 2) * (CO_T + 2) * (CO_CC / 4); i += 256) {
            const int c4 = i % (CO_CC / 4), pix = i / (CO_CC / 4);
            const int yy = ty0 + pix / (CO_T + 2) - 1, xx = tx0 + pix % (CO_T + 2) - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)n * H + yy) * W + xx) * C + c0 + c4 * 4));
                const float4 a = __ldg(reinterpret_cast<const float4*>(gn_scale + (size_t)n * C + c0 + c4 * 4));
                const float4 b = __ldg(reinterpret_cast<const float4*>(gn_shift + (size_t)n * C + c0 + c4 * 4));
                v.x = silu(fmaf(v.x, a.x, b.x)); v.y = silu(fmaf(v.y, a.y, b.y));
                v.z = silu(fmaf(v.z, a.z, b.z)); v.w = silu(fmaf(v.w, a.w, b.w));
            }
            tile[pix][c4 * 4 + 0] = v.x; tile[pix][c4 * 4 + 1] = v.y; tile[pix][c4 * 4 + 2] = v.z; tile[pix][c4 * 4 + 3] = v.w;
        }
        for (int i = threadIdx.x; i < 9 * CO_CC; i += 256) {
            const int tap = i / CO_CC, c = i % CO_CC;
            wsm[tap][c] = __ldg(reinterpret_cast<const float4*>(w + ((size_t)tap * C + c0 + c) * 4));
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float* tp = tile[(ly + tap / 3) * (CO_T + 2) + lx + tap % 3];
#pragma unroll
            for (int c = 0; c < CO_CC; ++c) {
                const float v = tp[c];
                const float4 ww = wsm[tap][c];
                a0 = fmaf(v, ww.x, a0); a1 = fmaf(v, ww.y, a1); a2 = fmaf(v, ww.z, a2);
            }
        }
    }
    const int y = ty0 + ly, x = tx0 + lx;
    if (y < H && x < W) {
        const size_t hw = (size_t)H * W;
        float* o = out + (size_t)n * 3 * hw + (size_t)y * W + x;
        o[0] = a0; o[hw] = a1; o[2 * hw] = a2;
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
