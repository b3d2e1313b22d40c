// Persistent tcgen05 full-attention kernel for the fixed 16x16 (+1 class) token grid: S = 257, head_dim = 64.
// (reference: nn.MultiheadAttention inside BertAttention, bert.py:84,137 -- softmax(Q K^T / 8) V per head, no mask, no cache)
//
// grid = #SMs, one CTA per SM looping over (sequence, head) work items; 12 warps:
//   warp 0      TMA producer: Q / K / V of the item (3 x 256 rows x 64 dims, 128B swizzle) + the class-token rows
//               (row 256 of Q, K, V) into a 2-stage shared-memory ring
//   warp 1      MMA issuer (one thread): per 128-query tile t:  S_t = Q_t K^T  (tcgen05.mma M128 N256 K64, smem x smem)
//                                                               O_t = P_t V    (A = P from TMEM, B = V MN-major from smem)
//   warps 2,3   class-token query row (q = 256) on CUDA cores, alternating items: 257 dot products, softmax, 257-term
//               weighted V sum; warp 3 also allocates TMEM (512 columns)
//   warps 4-11  softmax: warpgroup t owns query tile t, thread = one query row.  Two passes over S_t in TMEM
//               (max, then exp2 -> bf16 P written back over S), then (O_t + p256 v256) / l -> bf16 -> global.
//               The class-token KEY (k = 256) is a rank-1 side path: its score column comes from a 16-wide MMA
//               (Q_t x K[256..271]^T, column 0 used) issued with P V, its P*V is added in the epilogue.  Row sums l are
//               also produced by the tensor core (P x ones) so that the exp loop is FFMA2 + MUFU + F2FP only.
// TMEM per query tile (256 columns): [0,128) P as packed bf16x2 (aliases S columns already consumed), [128,192) O,
//               [192,208) class-key scores, [208,224) row sums.
// 257 = 2*128 + 1: tensor tiles cover the 256x256 block exactly; the odd row and column never touch a padded MMA tile.
//
// Input  qkv  bf16 [rows, 3*D] through two TMA maps (box 64x256 and box 64x16); head h at columns h*64 of each third
// Output out  bf16 [n_seq*257, D]
#pragma once
#include "ptx.cuh"

namespace mb {

constexpr int ATC_THREADS = 384;
constexpr int ATC_TILE_BYTES = 256 * 128;             // 256 rows x 64 bf16
constexpr int ATC_ROW_BYTES = 16 * 128;               // 16-row box holding the class-token row in its first 128 B
constexpr int ATC_STAGE_BYTES = 3 * ATC_TILE_BYTES + 3 * ATC_ROW_BYTES;
constexpr int ATC_ONES_BYTES = 16 * 128;              // B operand of the row-sum MMA: 16 x 64 bf16 ones
constexpr int ATC_SMEM_BYTES = 2 * ATC_STAGE_BYTES + ATC_ONES_BYTES + 1024 /*align*/ + 2304 /*p_cls*/ + 256 /*barriers*/;

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand tile stored as rows (K index) of 128 bytes (64 contiguous MN elements) with the 128-byte swizzle;
// 8 K-rows per 1024 B atom (stride-dim byte offset), one 64-element MN block (leading-dim offset unused).
__device__ __forceinline__ uint64_t make_sdesc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1024 >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; arguments here are <= 0, ftz is harmless
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32 pairs (FFMA2 / FMUL2 on sm_100): one issue slot for two lanes' worth of work
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 16-byte chunk `ch` of row `r` inside a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128(uint32_t tile, int r, int ch) { return tile + r * 128 + ((ch ^ (r & 7)) << 4); }

__device__ __forceinline__ float dot8(uint4 a, uint4 b, float acc) {
    acc = fmaf(bf_lo(a.x), bf_lo(b.x), acc); acc = fmaf(bf_hi(a.x), bf_hi(b.x), acc);
    acc = fmaf(bf_lo(a.y), bf_lo(b.y), acc); acc = fmaf(bf_hi(a.y), bf_hi(b.y), acc);
    acc = fmaf(bf_lo(a.z), bf_lo(b.z), acc); acc = fmaf(bf_hi(a.z), bf_hi(b.z), acc);
    acc = fmaf(bf_lo(a.w), bf_lo(b.w), acc); acc = fmaf(bf_hi(a.w), bf_hi(b.w), acc);
    return acc;
}

struct AttnTcParams {
    __nv_bfloat16* out;   // [n_seq*257, D]
    int n_items;          // n_seq * H
    int H, D;
    float sl2;            // log2(e) / sqrt(64)
};

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_big, const __grid_constant__ CUtensorMap tm_row, AttnTcParams p) {
    constexpr int S = 257;
    extern __shared__ uint8_t atc_smem_raw[];
    const uint32_t raw = smem_u32(atc_smem_raw);
    uint8_t* base = atc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint32_t* ones = reinterpret_cast<uint32_t*>(base + 2 * ATC_STAGE_BYTES);        // 2 KB of bf16 1.0
    float* p_cls = reinterpret_cast<float*>(base + 2 * ATC_STAGE_BYTES + ATC_ONES_BYTES);   // [2][288] class-row probabilities
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * ATC_STAGE_BYTES + ATC_ONES_BYTES + 2304);
    uint64_t* full = bars;          // [2] TMA landed
    uint64_t* empty = bars + 2;     // [2] stage consumed (MMA commit + 8 softmax warps + class warp)
    uint64_t* s_full = bars + 4;    // [2] S_t complete in TMEM
    uint64_t* p_full = bars + 6;    // [2] P_t written to TMEM (4 warps)
    uint64_t* o_full = bars + 8;    // [2] O_t complete
    uint64_t* t_free = bars + 10;   // [2] TMEM region t drained by the epilogue (4 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_big); tma_prefetch_desc(&tm_row); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], 10);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&o_full[s], 1); mbar_init(&t_free[s], 4);
        }
        fence_mbar_init();
    }
    if (warp == 3) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < ATC_ONES_BYTES / 4; i += ATC_THREADS) ones[i] = 0x3f803f80u;
    fence_async_proxy();                                    // generic-proxy smem writes -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(base);

    if (warp == 0) {
        if (lane == 0) {  // ---------------------------------------------------------------- TMA producer
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                const int seq = item / p.H, head = item - seq * p.H;
                const int row0 = seq * S, col = head * 64;
                uint8_t* sb = base + st * ATC_STAGE_BYTES;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&full[st], ATC_STAGE_BYTES);
                tma_load_2d(sb, &tm_big, &full[st], col, row0);                                          // Q rows 0..255
                tma_load_2d(sb + ATC_TILE_BYTES, &tm_big, &full[st], p.D + col, row0);                   // K
                tma_load_2d(sb + 2 * ATC_TILE_BYTES, &tm_big, &full[st], 2 * p.D + col, row0);           // V
                tma_load_2d(sb + 3 * ATC_TILE_BYTES, &tm_row, &full[st], col, row0 + 256);               // Q[256]
                tma_load_2d(sb + 3 * ATC_TILE_BYTES + ATC_ROW_BYTES, &tm_row, &full[st], p.D + col, row0 + 256);      // K[256]
                tma_load_2d(sb + 3 * ATC_TILE_BYTES + 2 * ATC_ROW_BYTES, &tm_row, &full[st], 2 * p.D + col, row0 + 256);  // V[256]
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_s = make_idesc(1, 128, 256);
            constexpr uint32_t idesc_o = make_idesc(1, 128, 64) | (1u << 16);   // B (= V) is MN-major
            constexpr uint32_t idesc_16 = make_idesc(1, 128, 16);
            const uint64_t ones_desc = make_sdesc_k128(smem0 + 2 * ATC_STAGE_BYTES);
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1, ip = it & 1;
                const uint32_t sq = smem0 + st * ATC_STAGE_BYTES, sk = sq + ATC_TILE_BYTES, sv = sk + ATC_TILE_BYTES;
                mbar_wait(&full[st], ph);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 2; ++t) {                       // S_t = Q_t K^T
                    mbar_wait(&t_free[t], ip ^ 1);
                    tc_fence_after();
                    const uint64_t a = make_sdesc_k128(sq + t * 128 * 128), b = make_sdesc_k128(sk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_base + t * 256, a + 2 * k, b + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[t]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {                       // O_t = P_t V
                    mbar_wait(&p_full[t], ip);
                    tc_fence_after();
                    const uint64_t b = make_sdesc_mn128(sv);
                    const uint32_t tr = tmem_base + t * 256;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        umma_f16_ts(tr + 128, tr + 8 * j, b + (uint64_t)(j * 128), idesc_o, j != 0);      // O += P_j V_j
                        umma_f16_ts(tr + 208, tr + 8 * j, ones_desc, idesc_16, j != 0);                   // l += P_j 1
                    }
                    const uint64_t a = make_sdesc_k128(sq + t * 128 * 128), kc16 = make_sdesc_k128(sq + 3 * ATC_TILE_BYTES + ATC_ROW_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tr + 192, a + 2 * k, kc16 + 2 * k, idesc_16, k != 0);  // Q_t K[256..]^T
                    umma_commit(&o_full[t]);
                }
                umma_commit(&empty[st]);
            }
        }
    } else if (warp == 2 || warp == 3) {  // -------------------------------------------- class-token query row
        // warps 2 and 3 alternate work items (each item's class row costs ~2.5k instructions of one warp)
        float* pc = p_cls + (warp - 2) * 288;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            if ((it & 1) != (uint32_t)(warp - 2)) continue;
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int seq = item / p.H, head = item - seq * p.H;
            const uint32_t sq = smem0 + st * ATC_STAGE_BYTES, sk = sq + ATC_TILE_BYTES, sv = sk + ATC_TILE_BYTES;
            const uint32_t qc = sq + 3 * ATC_TILE_BYTES, kc = qc + ATC_ROW_BYTES, vc = kc + ATC_ROW_BYTES;
            mbar_wait(&full[st], ph);
            float q[64];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 w = lds128(qc + c * 16);
                q[8 * c + 0] = bf_lo(w.x); q[8 * c + 1] = bf_hi(w.x); q[8 * c + 2] = bf_lo(w.y); q[8 * c + 3] = bf_hi(w.y);
                q[8 * c + 4] = bf_lo(w.z); q[8 * c + 5] = bf_hi(w.z); q[8 * c + 6] = bf_lo(w.w); q[8 * c + 7] = bf_hi(w.w);
            }
            auto qdot = [&](uint32_t row_addr, int rsw) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 w = lds128(row_addr + ((c ^ rsw) << 4));
                    a0 = fmaf(q[8 * c + 0], bf_lo(w.x), a0); a1 = fmaf(q[8 * c + 1], bf_hi(w.x), a1);
                    a0 = fmaf(q[8 * c + 2], bf_lo(w.y), a0); a1 = fmaf(q[8 * c + 3], bf_hi(w.y), a1);
                    a0 = fmaf(q[8 * c + 4], bf_lo(w.z), a0); a1 = fmaf(q[8 * c + 5], bf_hi(w.z), a1);
                    a0 = fmaf(q[8 * c + 6], bf_lo(w.w), a0); a1 = fmaf(q[8 * c + 7], bf_hi(w.w), a1);
                }
                return a0 + a1;
            };
            float sc[9];
#pragma unroll
            for (int r = 0; r < 8; ++r) sc[r] = qdot(sk + (lane + 32 * r) * 128, lane & 7);
            sc[8] = qdot(kc, 0);                            // key 256 (same value in every lane)
            float mx = sc[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) mx = fmaxf(mx, sc[r]);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float ms = mx * p.sl2;
            float sum = 0.f;
            __syncwarp();                                   // previous readers of pc are done
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float e = fast_exp2(fmaf(sc[r], p.sl2, -ms));
                sum += e;
                pc[lane + 32 * r] = e;
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float e256 = fast_exp2(fmaf(sc[8], p.sl2, -ms));
            sum += e256;
            __syncwarp();
            // P V: lane = (key group kg = lane >> 3: keys kg, kg+4, ...; dim chunk dc = lane & 7: dims 8dc..8dc+7)
            const int kg = lane >> 3, dc = lane & 7;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll 4
            for (int i = 0; i < 64; ++i) {
                const int key = kg + 4 * i;
                const float pk = pc[key];
                const uint4 w = lds128(sw128(sv, key, dc));
                o[0] = fmaf(pk, bf_lo(w.x), o[0]); o[1] = fmaf(pk, bf_hi(w.x), o[1]);
                o[2] = fmaf(pk, bf_lo(w.y), o[2]); o[3] = fmaf(pk, bf_hi(w.y), o[3]);
                o[4] = fmaf(pk, bf_lo(w.z), o[4]); o[5] = fmaf(pk, bf_hi(w.z), o[5]);
                o[6] = fmaf(pk, bf_lo(w.w), o[6]); o[7] = fmaf(pk, bf_hi(w.w), o[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
                o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
            }
            if (kg == 0) {
                const uint4 w = lds128(vc + dc * 16);
                const float inv = 1.0f / sum;
                uint4 r;
                r.x = pack2_bf16(fmaf(e256, bf_lo(w.x), o[0]) * inv, fmaf(e256, bf_hi(w.x), o[1]) * inv);
                r.y = pack2_bf16(fmaf(e256, bf_lo(w.y), o[2]) * inv, fmaf(e256, bf_hi(w.y), o[3]) * inv);
                r.z = pack2_bf16(fmaf(e256, bf_lo(w.z), o[4]) * inv, fmaf(e256, bf_hi(w.z), o[5]) * inv);
                r.w = pack2_bf16(fmaf(e256, bf_lo(w.w), o[6]) * inv, fmaf(e256, bf_hi(w.w), o[7]) * inv);
                __nv_bfloat16* orow = p.out + ((size_t)seq * S + 256) * p.D + head * 64;
                reinterpret_cast<uint4*>(orow)[dc] = r;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    } else if (warp >= 4) {  // ------------------------------------------------------------ softmax + epilogue
        const int t = (warp - 4) >> 2, quarter = warp & 3;
        const int row = t * 128 + quarter * 32 + lane;                 // query row inside the sequence (0..255)
        const uint32_t treg = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1, ip = it & 1;
            const int seq = item / p.H, head = item - seq * p.H;
            const uint32_t vc = smem0 + st * ATC_STAGE_BYTES + 3 * ATC_TILE_BYTES + 2 * ATC_ROW_BYTES;
            mbar_wait(&s_full[t], ip);
            tc_fence_after();
            float mx = -INFINITY;
            uint32_t va[32], vb[32];
            // pass 1: row max over keys 0..255.  The load of chunk c+1 is in flight while chunk c is reduced.
            tmem_ld_32x32(treg, va);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_32x32(treg + (c + 1) * 32, vb);
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
                tmem_ld_wait();
                tmem_ld_32x32(treg + ((c + 2) & 7) * 32, va);      // wraps to chunk 0: first chunk of pass 2
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
                tmem_ld_wait();
            }
            const float nms = -mx * p.sl2;
            // pass 2: p = exp2(s * sl2 - max * sl2) -> bf16 pairs written over S columns already consumed
            auto exp_store = [&](const uint32_t (&v)[32], int c) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x0, x1;
                    ffma2(x0, x1, __uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), p.sl2, p.sl2, nms, nms);
                    pk[j] = pack2_bf16(fast_exp2(x0), fast_exp2(x1));
                }
                tmem_st_32x32_x16(treg + c * 16, pk);
            };
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_32x32(treg + (c + 1) * 32, vb);
                exp_store(va, c);
                tmem_ld_wait();
                if (c + 2 < 8) tmem_ld_32x32(treg + (c + 2) * 32, va);
                exp_store(vb, c + 1);
                if (c + 2 < 8) tmem_ld_wait();
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            // epilogue: (O_t + p256 * V[256]) / (l + p256) -> bf16 row
            mbar_wait(&full[st], ph);                                  // (long complete) makes the TMA-written V[256] row visible
            uint4 vrow[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) vrow[c] = lds128(vc + c * 16);
            mbar_wait(&o_full[t], ip);
            tc_fence_after();
            tmem_ld_32x32(treg + 192, va);                             // [0] class-key score, [16] row sum
            tmem_ld_32x32(treg + 128, vb);                             // O columns 0..31
            tmem_ld_wait();
            // exponent clamped: if the class key dominates by more than 2^100 the result is V[256] to fp32 precision anyway
            const float e256 = fast_exp2(fminf(fmaf(__uint_as_float(va[0]), p.sl2, nms), 100.0f));
            const float inv = 1.0f / (__uint_as_float(va[16]) + e256);
            const float ei = e256 * inv;
            __nv_bfloat16* orow = p.out + ((size_t)seq * S + row) * p.D + head * 64;
            tmem_ld_32x32(treg + 160, va);                             // O columns 32..63 (in flight during the first half)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t (&v)[32] = hh ? va : vb;
                if (hh) tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 vv = vrow[hh * 4 + c];
                    const uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a, b, a2, b2;
                        fmul2(a, b, bf_lo(w[j]), bf_hi(w[j]), ei, ei);
                        ffma2(a2, b2, __uint_as_float(v[c * 8 + 2 * j]), __uint_as_float(v[c * 8 + 2 * j + 1]), inv, inv, a, b);
                        o[j] = pack2_bf16(a2, b2);
                    }
                    reinterpret_cast<uint4*>(orow)[hh * 4 + c] = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&t_free[t]); mbar_arrive(&empty[st]); }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace mb
