// Decoder convolutions on tcgen05 (reference autoencoder.py:7-36 Conv2dSame, :84-96 ResidualBlock, :224-225 upsample conv).
//
// A 3x3 (or 1x1) stride-1 SAME convolution over NHWC activations is a GEMM  out[pixel, cout] = sum_{tap, cin} A * W  with
//   M = pixels (tile: 128 consecutive pixels = bh full rows, or half a row at W = 256), N = 128 output channels,
//   K = taps x Cin consumed in chunks of 64 channels of one tap.
// Stride-2 convolutions of the tokenizer encoder (Conv2dSame: pad 0 top/left, 1 bottom/right) read the same zero-bordered
// input stored as four parity planes (space-to-depth of the padded image), so every tap is again a contiguous box.
// The 1e-3-abs pixel bar against the fp32 reference rules out single-pass bf16 (8e-2 measured), so both operands are split
// into bf16 hi + lo and every K-step issues three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM): ~2^-16 relative.
//
//   act_split_kernel      fp32 NHWC -> GroupNorm-apply + SiLU (optional) -> nearest x2 upsample (optional) -> bf16 hi / lo
//                         written into a zero-bordered [N, H+2, W+2, C] tensor, so that every tap of every tile is a plain
//                         TMA box (no im2col gather, no boundary predicates in the conv kernel)
//   conv_tcgen05_kernel   persistent, warp-specialised like the GEMM: TMA producer (A_hi, A_lo boxes of the tap-shifted
//                         padded input + W_hi, W_lo tiles) | one MMA thread (12 tcgen05.mma per stage) | 8 epilogue warps
//                         (bias, fp32 residual, fp32 NHWC output through smem + TMA store)
#pragma once
#include "ptx.cuh"

namespace mb {

// ------------------------------------------------------------------------------------------------ activation transform
// in    fp32 NHWC [N, Hin, Win, C]  (Hin = H >> up)
// scale / shift  [N][C] GroupNorm-apply coefficients or nullptr (no norm, no SiLU)
// hi/lo bf16 [N, H+2, W+2, C], border = 0;  planes = 1
//       or, planes = 4 (input of a stride-2 conv): [4][N, (H+2)/2, (W+2)/2, C], plane = (yp & 1) * 2 + (xp & 1) of padded (yp, xp)
// grid = (ceil((W+2) * C/8 / 256), N * (H+2)): one padded row per blockIdx.y, so the only per-thread index arithmetic is one
// 32-bit divide (the 64-bit div / mod chain of a flat index held this HBM-bound kernel at 43-60 % of the copy bandwidth).
__global__ void __launch_bounds__(256)
act_split_kernel(const float* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int N, int H, int W, int C, int up, int planes) {
    const int c8 = C >> 3;                                      // 8-channel groups per pixel
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // position inside the padded row
    if (i >= (W + 2) * c8) return;
    const int xp = i / c8, cg = i - xp * c8;
    const int n = blockIdx.y / (H + 2), yp = blockIdx.y - n * (H + 2);
    const long long idx = ((long long)blockIdx.y * (W + 2) + xp) * c8 + cg;
    uint4 oh = make_uint4(0, 0, 0, 0), ol = make_uint4(0, 0, 0, 0);
    if (xp >= 1 && xp <= W && yp >= 1 && yp <= H) {
        const int Hin = H >> up, Win = W >> up;
        const int ys = (yp - 1) >> up, xs = (xp - 1) >> up;
        const float4* src = reinterpret_cast<const float4*>(in + (((size_t)n * Hin + ys) * Win + xs) * C + cg * 8);
        float4 a = __ldg(src), b = __ldg(src + 1);
        float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (scale) {
            const float4* sp = reinterpret_cast<const float4*>(scale + (size_t)n * C + cg * 8);
            const float4* hp = reinterpret_cast<const float4*>(shift + (size_t)n * C + cg * 8);
            const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1), h0 = __ldg(hp), h1 = __ldg(hp + 1);
            const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float t = fmaf(v[i], s[i], h[i]);
                v[i] = __fdividef(t, 1.0f + __expf(-t));          // SiLU
            }
        }
        uint32_t wh[4], wl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
            wh[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            wl[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        oh = make_uint4(wh[0], wh[1], wh[2], wh[3]);
        ol = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
    long long o = idx;
    if (planes == 4) {
        const int Hp = (H + 2) >> 1, Wp = (W + 2) >> 1;
        const int pl = (yp & 1) * 2 + (xp & 1);
        o = ((((long long)pl * N + n) * Hp + (yp >> 1)) * Wp + (xp >> 1)) * c8 + cg;
    }
    reinterpret_cast<uint4*>(hi)[o] = oh;
    reinterpret_cast<uint4*>(lo)[o] = ol;
}

// ------------------------------------------------------------------------------------------------ conv kernel
struct ConvTcParams {
    int n_img, H, W, Cin, Cout, taps;   // H, W = OUTPUT size; taps = 9 (3x3) or 1 (1x1)
    int stride;                         // 1, or 2 (3x3 only; input stored as 4 parity planes)
    int phases;                         // 1, or 4: nearest-x2 upsample folded into the conv (autoencoder.py:224-225).  H, W are then
                                        // the LOW-resolution size; output pixel (2y+py, 2x+px) of phase (py, px) is a 2x2-tap conv
                                        // over the low-resolution input with pre-summed weights (taps = 4, tap (a, b) reads padded
                                        // pixel (y + a + py, x + b + px)): 16 instead of 36 tap-products per 2x2 output pixels.
                                        // Weights [phase][Cout][4*Cin]; tm_out is the 5-D view {C, px, W, py, n*H + y} of the output
    int bw, bh;                         // pixel tile = bh rows x bw columns, bw * bh = 128
    const float* bias;                  // [Cout] or nullptr
    const float* residual;              // fp32 NHWC [n_img*H*W, Cout] or nullptr
    float2* gn_part;                    // [n_img*H*W / 32][32] (sum, sumsq) of every 32-pixel x GroupNorm-group box of the OUTPUT, or
                                        // nullptr: the statistics pass of the GroupNorm that consumes this conv's output
                                        // (autoencoder.py:39-43, 32 groups) is fused here instead of re-reading the activation
    int gn_cpg_log2;                    // log2(channels per group) = log2(Cout / 32): 2, 3 or 4
};

struct ConvTcCfg {
    static constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
    static constexpr int TILE_BYTES = 128 * BK * 2;                    // every operand tile: 128 rows x 64 bf16
    static constexpr int STAGE_BYTES = 4 * TILE_BYTES;                 // A_hi, A_lo, W_hi, W_lo
    static constexpr int STG_BYTES = 8 * 4096;                         // output staging: 8 warps x (32 rows x 32 fp32)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 + 256;
};

// tm_ahi / tm_alo : 5D maps of the padded bf16 inputs {C, Wp, Hp, N, planes}, box {64, bw, bh, 1, 1}
//                   (stride 1: Wp = W+2, planes = 1; stride 2: Wp = (Win+2)/2 = W+1, planes = 4)
// tm_whi / tm_wlo : 2D maps of the packed weights [Cout][taps*Cin], box {64, 128}
// tm_out          : 2D map of the fp32 output [n_img*H*W, Cout], box {32, 32}
__global__ void __launch_bounds__(384, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                    const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo,
                    const __grid_constant__ CUtensorMap tm_out, ConvTcParams p) {
    using C = ConvTcCfg;
    constexpr int BN = C::BN, STAGES = C::STAGES;
    extern __shared__ uint8_t cv_smem_raw[];
    const uint32_t raw = smem_u32(cv_smem_raw);
    uint8_t* base = cv_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* smem_stg = base + STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + C::STG_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = p.W / p.bw, tiles_y = p.H / p.bh;
    const int pix_tiles = p.n_img * tiles_y * tiles_x;
    const int num_n = p.Cout / BN;
    const int num_np = num_n * p.phases;                 // tiles sharing one input pixel tile are consecutive
    const int num_tiles = pix_tiles * num_np;
    const int kchunks = p.Cin / C::BK;
    const int num_k = p.taps * kchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_ahi); tma_prefetch_desc(&tm_alo); tma_prefetch_desc(&tm_whi); tma_prefetch_desc(&tm_wlo);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 8); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {  // ---------------- TMA producer
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int pt = tile / num_np, rem = tile - pt * num_np;
                const int uph = rem / num_n, n_blk = rem - uph * num_n;
                const int xb = pt % tiles_x, yb = (pt / tiles_x) % tiles_y, img = pt / (tiles_x * tiles_y);
                const int x0 = xb * p.bw, y0 = yb * p.bh;
                const int wrow = uph * p.Cout + n_blk * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    const int tap = kb / kchunks, c0 = (kb - tap * kchunks) * C::BK;
                    // stride 1, 3x3: padded input pixel of output (y, x) under tap (dy, dx) is (y + dy, x + dx), dy, dx in 0..2;
                    //           1x1: the centre (y + 1, x + 1)
                    // stride 2 (SAME pad 0 / 1): padded pixel (2y + dy + 1, 2x + dx + 1) = plane ((dy+1)&1, (dx+1)&1), cell
                    //           (y + (dy+1)/2, x + (dx+1)/2)
                    int dy = p.taps == 9 ? tap / 3 : 1, dx = p.taps == 9 ? tap % 3 : 1, plane = 0;
                    if (p.stride == 2) { plane = ((dy + 1) & 1) * 2 + ((dx + 1) & 1); dy = (dy + 1) >> 1; dx = (dx + 1) >> 1; }
                    if (p.phases == 4) { dy = (tap >> 1) + (uph >> 1); dx = (tap & 1) + (uph & 1); }
                    uint8_t* sb = base + stage * C::STAGE_BYTES;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    tma_load_5d(sb, &tm_ahi, &full[stage], c0, x0 + dx, y0 + dy, img, plane);
                    tma_load_5d(sb + C::TILE_BYTES, &tm_alo, &full[stage], c0, x0 + dx, y0 + dy, img, plane);
                    tma_load_2d(sb + 2 * C::TILE_BYTES, &tm_whi, &full[stage], tap * p.Cin + c0, wrow);
                    tma_load_2d(sb + 3 * C::TILE_BYTES, &tm_wlo, &full[stage], tap * p.Cin + c0, wrow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {  // ---------------- MMA issuer: D += A_lo W_hi + A_hi W_lo + A_hi W_hi (small terms first)
            constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 128, BN);
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(base + stage * C::STAGE_BYTES);
                    const uint64_t ahi = make_sdesc_k128(sb), alo = make_sdesc_k128(sb + C::TILE_BYTES);
                    const uint64_t whi = make_sdesc_k128(sb + 2 * C::TILE_BYTES), wlo = make_sdesc_k128(sb + 3 * C::TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < C::BK / 16; ++k) {
                        umma_f16(d_tmem, alo + 2 * k, whi + 2 * k, idesc, (kb | k) != 0);
                        umma_f16(d_tmem, ahi + 2 * k, wlo + 2 * k, idesc, 1);
                        umma_f16(d_tmem, ahi + 2 * k, whi + 2 * k, idesc, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else if (warp >= 4) {  // ---------------- epilogue: + bias + residual -> fp32 NHWC via smem + TMA store
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        uint8_t* stg = smem_stg + (warp - 4) * 4096;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int pt = tile / num_np, rem = tile - pt * num_np;
            const int uph = rem / num_n, n_blk = rem - uph * num_n;
            const uint32_t as = it & 1, aphase = (it >> 1) & 1;
            const long long row0 = (long long)pt * 128 + quarter * 32;     // tiles enumerate pixels in NHWC order
            const long long row = row0 + lane;
            // phases == 4: this warp's 32 low-resolution pixels (one row segment, or two 16-pixel rows) and its statistics slot
            const int tpi = tiles_x * tiles_y, img = pt / tpi, pti = pt - img * tpi;
            const int px0 = (pti % tiles_x) * p.bw + (quarter * 32) % p.bw, py0 = (pti / tiles_x) * p.bh + (quarter * 32) / p.bw;
            const long long gn_slot = p.phases == 4 ? ((long long)(img * 4 + uph) * tpi + pti) * 4 + quarter : (row0 >> 5);
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 2; c += 32) {
                const int col0 = half * (BN / 2) + c;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + col0, v);
                tmem_ld_wait();
                const int n0 = n_blk * BN + col0;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                if (p.residual) {
                    const float4* rp = reinterpret_cast<const float4*>(p.residual + (size_t)row * p.Cout + n0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 r4 = __ldg(rp + j);
                        f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
                    }
                }
                if (elect_one()) tma_store_wait_read<0>();   // bulk groups are per thread: the lane that issued the store            // previous store finished reading the staging buffer
                __syncwarp();
                uint8_t* rowp = stg + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                fence_async_proxy();
                __syncwarp();
                if (elect_one()) {
                    if (p.phases == 4) tma_store_5d(&tm_out, stg, n0, uph & 1, px0, uph >> 1, img * p.H + py0);
                    else tma_store_2d(&tm_out, stg, n0, (int)row0);
                    tma_store_commit();
                }
                if (p.gn_part) {
                    // GroupNorm partials of the values as stored: lane = channel n0 + lane, summed over the box's 32 pixels from
                    // the staging tile (a row's 16-byte chunks are XOR-swizzled by the row: one bank per lane, no conflicts),
                    // then over the lanes of a group.  Fixed order -> run-to-run identical statistics.
                    float s1 = 0.f, s2 = 0.f;
                    const uint8_t* colp = stg + (lane & 3) * 4;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const float v = *reinterpret_cast<const float*>(colp + r * 128 + (((lane >> 2) ^ (r & 7)) << 4));
                        s1 += v; s2 = fmaf(v, v, s2);
                    }
                    for (int o = 1; o < (1 << p.gn_cpg_log2); o <<= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    }
                    if ((lane & ((1 << p.gn_cpg_log2) - 1)) == 0)
                        p.gn_part[(size_t)gn_slot * 32 + ((n0 + lane) >> p.gn_cpg_log2)] = make_float2(s1, s2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        if (elect_one()) tma_store_wait_all<0>();
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

}  // namespace mb
