// Persistent tcgen05 full-attention kernel for the fixed 16x16 (+1 class) token grid: S = 257, head_dim = 64.
// (reference: nn.MultiheadAttention inside BertAttention, bert.py:84,137 -- softmax(Q K^T / 8) V per head, no mask, no cache)
//
// grid = #SMs, one CTA per SM looping over (sequence, head) work items; 12 warps:
//   warp 0      TMA producer: Q / K / V of the item (3 x 256 rows x 64 dims, 128B swizzle) + 16-row boxes holding the
//               class-token rows (row 256 of Q, K, V) into a 2-stage shared-memory ring
//   warps 1,2   MMA issuers (one thread each, one per 128-query tile t):
//                   S_t = Q_t K^T            tcgen05.mma M128 N256 K16 x4, smem x smem
//                   O_t = P_t V              A = P from TMEM, B = V MN-major from smem, 16 K-steps into one accumulator
//   warp 3      the class-token QUERY row q = 256: softmax over its 257 scores (256 of them computed by the softmax warps,
//               below), then O = p V as 68 warp-level mma.sync tiles (A = V^T via ldmatrix.trans, B = p in column 0) -> global
//               (also allocates TMEM, 512 columns)
//   warps 4-11  softmax: warpgroup t owns query tile t, thread = one query row.  First the 257th row and column for the
//               warp's own 32 rows / keys on mma.sync tiles (16 ldmatrix + 16 mma per warp, while the S MMAs run):
//                   s256[r] = Q[r] . K[256]  (class-token KEY column, stays in a register of the row's thread)
//                   scls[r] = Q[256] . K[r]  (class-token QUERY row) -> smem for warp 3
//               (one warp doing all of this cost 10 k clk per item and set the kernel's pace: profiles/r02a_attn_trace.txt).
//               Pass 1 row max over S_t in TMEM,
//               pass 2 exp2 -> bf16 P written back over the S columns already consumed (FFMA2 + MUFU + F2FP + FADD2),
//               epilogue (O + p256 V[256]) / l -> bf16 -> global.  The two warpgroups take turns in pass 2
//               (named-barrier ping-pong) so each has the MUFU pipe to itself while the other waits on its MMAs.
// TMEM per query tile (256 columns): S fp32 [0,256) -> P packed bf16x2 [0,128), O [128,192).
// 257 = 2*128 + 1: tensor tiles cover the 256x256 block exactly; the odd row and column never touch a padded tcgen05 tile.
//
// Input  qkv  bf16 [rows, 3*D] through two TMA maps (box 64x256 and box 64x16); head h at columns h*64 of each third
// Output out  bf16 [n_seq*257, D] (one 128-byte row segment per thread, 256-bit stores)
#pragma once
#include <type_traits>
#include "attention.cuh"   // ldsm_x4, ldsm_x4_t, mma_bf16_16816
#include "ptx.cuh"

#ifndef ATC_TRACE
#define ATC_TRACE 0            // 1: block 0 records (role, event, item, clock) tuples into AttnTcParams::trace
#endif
#if ATC_TRACE
#define ATC_EV(role, ev, it)                                                                                       \
    do {                                                                                                           \
        if (blockIdx.x == 0 && p.trace && (it) < 12) {                                                             \
            long long* _t = p.trace + (((role) * 8 + (ev)) * 12 + (it));                                           \
            *_t = clock64();                                                                                       \
        }                                                                                                          \
    } while (0)
#else
#define ATC_EV(role, ev, it) do {} while (0)
#endif
#ifndef ATC_PINGPONG
#define ATC_PINGPONG 1      // named-barrier turn-taking of the two softmax warpgroups in pass 2 (first tried: slower, 0.417 vs 0.365 ms; with the
                            // 3-input-max pass 1 it is 4 % faster, 0.374 vs 0.389 ms, and 1.3 % less energy per launch: tools/kpower.py)
#endif

#ifndef ATC_POLY
#define ATC_POLY 0          // of every 16 exponential pairs in pass 2, this many are evaluated on the FMA pipe (round-to-nearest range
                            // reduction + degree-3 polynomial + exponent insert) instead of MUFU.EX2: the MUFU pipe (16 / clk / SM) is
                            // the floor of pass 2 (4.1 k clk per item); P is rounded to bf16 afterwards, the polynomial's 7.5e-5 is invisible
#endif
#ifndef ATC_SPLIT_S
#define ATC_SPLIT_S 0       // 1: S_t as two N = 128 halves (keys 0..127 -> TMEM columns [0,128), keys 128..255 -> [128,256)).  The first
                            // half of the NEXT item is issued right behind this item's P V MMAs (same thread: in order, and by then P in
                            // [0,128) is consumed), i.e. while the epilogue still drains O from [128,192); only the second half waits for
                            // t_free.  Takes half of the S latency off the warpgroup's serial chain.
#endif
#ifndef ATC_FASTMAX
#define ATC_FASTMAX 0       // 1: no row-max pass.  Softmax is shift-invariant and P (bf16) / the row sum / O (fp32) have the fp32 exponent
                            // range, so any shift within ~2^100 of the true maximum gives the same result: rows are shifted by their
                            // class-key score s256 (already in a register).  A row whose sum overflows (maximum more than ~88 nats above
                            // its class-key logit) flags its item; flagged items are redone by the same CTA with the exact row maximum
                            // after its main loop (second round below), so the result never depends on the estimate.
#endif

namespace mb {

constexpr int ATC_THREADS = 384;
constexpr int ATC_TILE_BYTES = 256 * 128;             // 256 rows x 64 bf16
constexpr int ATC_ROW_BYTES = 16 * 128;               // 16-row box holding the class-token row in its first 128 B
constexpr int ATC_STAGE_BYTES = 3 * ATC_TILE_BYTES + 3 * ATC_ROW_BYTES;
constexpr int ATC_SCLS_BYTES = 2 * 256 * 4;           // class-query scores against the 256 tile keys, per stage
constexpr int ATC_PCLS_BYTES = 576;                   // class-query probabilities, bf16 [272] (keys 257.. = 0)
constexpr int ATC_RETRY_CAP = 1024;                   // local items per CTA that the fast-max round can flag (more: exact from the start)
constexpr int ATC_RETRY_BYTES = ATC_RETRY_CAP / 8 + ATC_RETRY_CAP * 2 + 16;   // bit mask + compacted list (uint16) + count
constexpr int ATC_SMEM_BYTES = 2 * ATC_STAGE_BYTES + ATC_SCLS_BYTES + ATC_PCLS_BYTES + 1024 /*align*/ + 256 /*barriers*/ + ATC_RETRY_BYTES;

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
// exp2 of a pair on the FMA pipe: x = n + f with n = round(x) (magic-number add), f in [-0.5, 0.5]; 2^f by a degree-3 minimax
// polynomial (7.5e-5 relative); 2^n inserted by adding n to the exponent field.  Valid for x in [-126, 127].
__device__ __forceinline__ void exp2_poly2(float& x0, float& x1) {
    constexpr float MAGIC = 12582912.0f;                 // 1.5 * 2^23: the integer lands in the low mantissa bits
    x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
#if ATC_FASTMAX
    x0 = fminf(x0, 128.0f); x1 = fminf(x1, 128.0f);      // 2^128 -> exponent field 255: inf / NaN, caught by the row-sum check
#endif
    float t0, t1, n0, n1, f0, f1, p0, p1;
    fadd2(t0, t1, x0, x1, MAGIC, MAGIC);
    fadd2(n0, n1, t0, t1, -MAGIC, -MAGIC);
    ffma2(f0, f1, n0, n1, -1.0f, -1.0f, x0, x1);
    ffma2(p0, p1, f0, f1, 0.0551716685f, 0.0551716685f, 0.242611125f, 0.242611125f);
    ffma2(p0, p1, p0, p1, f0, f1, 0.693260968f, 0.693260968f);
    ffma2(p0, p1, p0, p1, f0, f1, 0.999928057f, 0.999928057f);
    x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// three-input maximum (FMNMX3 on sm_100): the row-max pass needs one instruction per two scores
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 16-byte chunk `ch` of row `r` inside a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128(uint32_t tile, int r, int ch) { return tile + r * 128 + ((ch ^ (r & 7)) << 4); }

struct AttnTcParams {
    __nv_bfloat16* out;   // [n_seq*257, D]
    int n_items;          // n_seq * H
    int H, D;
    float sl2;            // log2(e) / sqrt(64)
    long long* trace;     // ATC_TRACE builds only: [roles 8][events 8][items 12] clock64 stamps of block 0
};

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_big, const __grid_constant__ CUtensorMap tm_row, AttnTcParams p) {
    constexpr int S = 257;
    extern __shared__ uint8_t atc_smem_raw[];
    const uint32_t raw = smem_u32(atc_smem_raw);
    uint8_t* base = atc_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    float* scls = reinterpret_cast<float*>(base + 2 * ATC_STAGE_BYTES);              // [2 stages][256 keys]
    __nv_bfloat16* pcls = reinterpret_cast<__nv_bfloat16*>(base + 2 * ATC_STAGE_BYTES + ATC_SCLS_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * ATC_STAGE_BYTES + ATC_SCLS_BYTES + ATC_PCLS_BYTES);
    // Each stage is two independently recycled halves: Q | K | class rows of Q, K (needed until S and the class scores are done,
    // early in an item) and V | class row of V (needed until P V is done, at its end).  With one barrier pair per stage the next
    // item's load could only start when the LATER warpgroup had finished the previous item, and arrived ~2 k clk after the
    // earlier warpgroup wanted it (profiles/r02c_attn_trace.txt).
    uint64_t* fullqk = bars;         // [2] Q, K (+ class rows) landed
    uint64_t* fullv = bars + 2;      // [2] V (+ class row) landed
    uint64_t* emptyqk = bars + 4;    // [2] consumed: 2 S commits + 8 softmax warps (class pass) + class warp
    uint64_t* emptyv = bars + 6;     // [2] consumed: 2 P V commits + 8 softmax warps (epilogue) + class warp
    uint64_t* s_full = bars + 8;     // [2] S_t complete in TMEM
    uint64_t* p_full = bars + 10;    // [2] P_t written to TMEM (4 warps)
    uint64_t* o_full = bars + 12;    // [2] O_t complete
    uint64_t* t_free = bars + 14;    // [2] TMEM region t drained by the epilogue (4 warps)
    uint64_t* scls_ready = bars + 16;// [2] scls[stage] written by the 8 softmax warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    uint32_t* retry_mask = reinterpret_cast<uint32_t*>(bars + 32);                               // [ATC_RETRY_CAP / 32]
    uint16_t* retry_list = reinterpret_cast<uint16_t*>(retry_mask + ATC_RETRY_CAP / 32);         // [ATC_RETRY_CAP]
#if ATC_FASTMAX
    int* retry_count = reinterpret_cast<int*>(retry_list + ATC_RETRY_CAP);
#endif

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_big); tma_prefetch_desc(&tm_row); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&fullqk[s], 1); mbar_init(&fullv[s], 1); mbar_init(&emptyqk[s], 11); mbar_init(&emptyv[s], 11);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&o_full[s], 1); mbar_init(&t_free[s], 4);
            mbar_init(&scls_ready[s], 8);
        }
        fence_mbar_init();
    }
    if (warp == 2 && lane < ATC_RETRY_CAP / 32) retry_mask[lane] = 0;
    if (warp == 3) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(base);
    pdl_launch_dependents();
    pdl_wait();                                      // qkv is complete from here on
    const int n_local = p.n_items > (int)blockIdx.x ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // Processed-item sequence of this CTA: k in [0, n_local) is local item k; with ATC_FASTMAX a second round
    // k in [n_local, n_local + n_retry) redoes the flagged local items retry_list[k - n_local] with the exact row maximum.
    // Every role runs the same sequence; k is also the running index all barrier parities and stage slots derive from.
    const bool fast_round = ATC_FASTMAX && n_local <= ATC_RETRY_CAP;
    int n_retry = 0;
    auto item_of = [&](int k) -> int { return (int)blockIdx.x + (k < n_local ? k : (int)retry_list[k - n_local]) * (int)gridDim.x; };
#pragma unroll 1
  for (int round = 0; round < (ATC_FASTMAX ? 2 : 1); ++round) {
    const int k0 = round ? n_local : 0, k1 = round ? n_local + n_retry : n_local;
    const bool exact = round != 0 || !fast_round;
    (void)exact;

    if (warp == 0) {
        if (elect_one()) {  // -------------------------------------------------------------- TMA producer
            // (elect.sync region, here and in the MMA issuers: the compiler keeps descriptors and addresses in uniform registers;
            //  under `lane == 0` every UTMALDG / UTCHMMA was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, ~150 clk each)
            for (int it = k0; it < k1; ++it) {
                const int item = item_of(it);
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                const int seq = item / p.H, head = item - seq * p.H;
                const int row0 = seq * S, col = head * 64;
                uint8_t* sb = base + st * ATC_STAGE_BYTES;
                // (the V half of a slot is released before the Q/K half of the other slot: waiting in this order never holds back
                //  a load whose buffer is already free)
                mbar_wait(&emptyqk[st], ph ^ 1);
                ATC_EV(0, 0, it);
                mbar_arrive_expect_tx(&fullqk[st], 2 * ATC_TILE_BYTES + 2 * ATC_ROW_BYTES);
                tma_load_2d(sb + ATC_TILE_BYTES, &tm_big, &fullqk[st], p.D + col, row0);                 // K rows 0..255
                tma_load_2d(sb, &tm_big, &fullqk[st], col, row0);                                        // Q
                tma_load_2d(sb + 3 * ATC_TILE_BYTES, &tm_row, &fullqk[st], col, row0 + 256);             // Q[256..]
                tma_load_2d(sb + 3 * ATC_TILE_BYTES + ATC_ROW_BYTES, &tm_row, &fullqk[st], p.D + col, row0 + 256);    // K[256..]
                mbar_wait(&emptyv[st], ph ^ 1);
                mbar_arrive_expect_tx(&fullv[st], ATC_TILE_BYTES + ATC_ROW_BYTES);
                tma_load_2d(sb + 2 * ATC_TILE_BYTES, &tm_big, &fullv[st], 2 * p.D + col, row0);          // V
                tma_load_2d(sb + 3 * ATC_TILE_BYTES + 2 * ATC_ROW_BYTES, &tm_row, &fullv[st], 2 * p.D + col, row0 + 256);  // V[256..]
            }
        }
    } else if (warp == 1 || warp == 2) {
        if (elect_one()) {  // -------------------------------------------------------------- MMA issuers: warp 1 -> tile 0, warp 2 -> tile 1
            constexpr uint32_t idesc_s = make_idesc(1, 128, 256); (void)idesc_s;
            constexpr uint32_t idesc_o = make_idesc(1, 128, 64) | (1u << 16);   // B (= V) is MN-major
            const int t = warp - 1;
            const uint32_t tr = tmem_base + t * 256;
#if ATC_SPLIT_S
            constexpr uint32_t idesc_sh = make_idesc(1, 128, 128);
            bool sa_done = false;                                               // first half of S for item `it` already issued
#endif
            for (int it = k0; it < k1; ++it) {
                const int st = it & 1;
                const uint32_t sq = smem0 + st * ATC_STAGE_BYTES, sk = sq + ATC_TILE_BYTES, sv = sk + ATC_TILE_BYTES;
                mbar_wait(&fullqk[st], (it >> 1) & 1);
                ATC_EV(5 + t, 0, it);
#if ATC_SPLIT_S
                {                                                               // S_t = Q_t K^T in two key halves
                    const uint64_t a = make_sdesc_k128(sq + t * 128 * 128);
                    if (!sa_done) {                                             // first item of a round: the whole tile must be free
                        mbar_wait(&t_free[t], (it & 1) ^ 1);
                        tc_fence_after();
                        const uint64_t b0 = make_sdesc_k128(sk);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tr, a + 2 * k, b0 + 2 * k, idesc_sh, k != 0);
                    } else {
                        mbar_wait(&t_free[t], (it & 1) ^ 1);                    // O of the previous item is out of columns [128,192)
                        tc_fence_after();
                    }
                    ATC_EV(5 + t, 1, it);
                    const uint64_t b1 = make_sdesc_k128(sk + 128 * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tr + 128, a + 2 * k, b1 + 2 * k, idesc_sh, k != 0);
                    umma_commit(&s_full[t]);
                    umma_commit(&emptyqk[st]);                                  // this tile's S MMAs have read Q and K
                    ATC_EV(1, t, it);
                }
#else
                mbar_wait(&t_free[t], (it & 1) ^ 1);
                ATC_EV(5 + t, 1, it);
                tc_fence_after();
                {                                                               // S_t = Q_t K^T
                    const uint64_t a = make_sdesc_k128(sq + t * 128 * 128), b = make_sdesc_k128(sk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tr, a + 2 * k, b + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[t]);
                    umma_commit(&emptyqk[st]);                                  // this tile's S MMAs have read Q and K
                    ATC_EV(1, t, it);
                }
#endif
                mbar_wait(&p_full[t], it & 1);
                ATC_EV(5 + t, 2, it);
                mbar_wait(&fullv[st], (it >> 1) & 1);
                ATC_EV(5 + t, 3, it);
                tc_fence_after();
                {                                                               // O_t = P_t V
                    const uint64_t b = make_sdesc_mn128(sv);
#pragma unroll
                    for (int j = 0; j < 16; ++j) umma_f16_ts(tr + 128, tr + 8 * j, b + (uint64_t)(j * 128), idesc_o, j != 0);
                    umma_commit(&o_full[t]);
                    umma_commit(&emptyv[st]);                                   // this tile's P V MMAs have read V
                    ATC_EV(1, 2 + t, it);
                }
#if ATC_SPLIT_S
                sa_done = false;
                if (it + 1 < k1) {                                              // next item's first S half, behind the P V MMAs
                    const int sn = (it + 1) & 1;
                    const uint32_t sqn = smem0 + sn * ATC_STAGE_BYTES;
                    mbar_wait(&fullqk[sn], ((it + 1) >> 1) & 1);
                    tc_fence_after();
                    const uint64_t a = make_sdesc_k128(sqn + t * 128 * 128), b0 = make_sdesc_k128(sqn + ATC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tr, a + 2 * k, b0 + 2 * k, idesc_sh, k != 0);
                    sa_done = true;
                }
#endif
            }
        }
    } else if (warp == 3) {  // -------------------------------------------------------------- the class-token query row
        const int g = lane >> 2, t4 = lane & 3;
        const uint32_t pcls_a = smem_u32(pcls);
        for (int it = k0; it < k1; ++it) {
            const int item = item_of(it);
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int seq = item / p.H, head = item - seq * p.H;
            const uint32_t sq = smem0 + st * ATC_STAGE_BYTES, sv = sq + 2 * ATC_TILE_BYTES;
            const uint32_t qc = sq + 3 * ATC_TILE_BYTES, kc = qc + ATC_ROW_BYTES, vc = kc + ATC_ROW_BYTES;
            mbar_wait(&fullqk[st], ph);
            if (lane == 0) ATC_EV(2, 0, it);
            // its score against its own key: lane l holds dims 2l, 2l+1
            const uint32_t qw = lds32(qc + lane * 4), kw = lds32(kc + lane * 4);
            float sd = fmaf(bf_lo(qw), bf_lo(kw), bf_hi(qw) * bf_hi(kw));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
            // the other 256 scores come from the softmax warps (lane l takes keys l, l + 32, ...)
            mbar_wait(&scls_ready[st], ph);
            float sc[8];
            float mx = sd;
#pragma unroll
            for (int i = 0; i < 8; ++i) { sc[i] = scls[st * 256 + lane + 32 * i]; mx = fmaxf(mx, sc[i]); }
            // Only now hand the Q/K half back: the softmax warps refill scls[st] (and re-arrive on scls_ready[st]) as soon as the
            // NEXT item using this slot has landed, which must not happen before this warp has read the current scores.
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyqk[st]);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float nms = -mx * p.sl2;
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float e = fast_exp2(fmaf(sc[i], p.sl2, nms));
                sum += e;
                pcls[lane + 32 * i] = __float2bfloat16_rn(e);
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float e256 = fast_exp2(fmaf(sd, p.sl2, nms));
            if (lane < 16) pcls[256 + lane] = __float2bfloat16_rn(lane == 0 ? e256 : 0.f);   // keys 257.. do not exist
            __syncwarp();
            mbar_wait(&fullv[st], ph);
            // O^T[d][0] = sum_key V^T[d][key] p[key]: 4 d-tiles x 17 key-tiles of m16n8k16 (tile 16 = keys 256..271 from the
            // class-row box); A = V^T by ldmatrix.trans of the swizzled V tile, B = p in column n = 0
            float o[4][4];
#pragma unroll
            for (int dt = 0; dt < 4; ++dt) { o[dt][0] = 0.f; o[dt][1] = 0.f; o[dt][2] = 0.f; o[dt][3] = 0.f; }
#pragma unroll
            for (int kt = 0; kt < 17; ++kt) {
                const uint32_t b0 = g == 0 ? lds32(pcls_a + (kt * 16 + 2 * t4) * 2) : 0u;
                const uint32_t b1 = g == 0 ? lds32(pcls_a + (kt * 16 + 8 + 2 * t4) * 2) : 0u;
                const uint32_t tile = kt < 16 ? sv : vc;
                const int r = (kt < 16 ? kt * 16 : 0) + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
                for (int dt = 0; dt < 4; ++dt) {
                    uint32_t a[4];
                    ldsm_x4_t(a, sw128(tile, r, 2 * dt + ((lane >> 3) & 1)));
                    mma_bf16_16816(o[dt], a, b0, b1);
                }
            }
            if (t4 == 0) {
                const float inv = 1.0f / (sum + e256);
                __nv_bfloat16* orow = p.out + ((size_t)seq * S + 256) * p.D + head * 64 + g;
#pragma unroll
                for (int dt = 0; dt < 4; ++dt) {
                    orow[dt * 16] = __float2bfloat16_rn(o[dt][0] * inv);
                    orow[dt * 16 + 8] = __float2bfloat16_rn(o[dt][2] * inv);
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&emptyv[st]); ATC_EV(2, 1, it); }
        }
    } else if (warp >= 4) {  // ------------------------------------------------------------ softmax + epilogue
        const int t = (warp - 4) >> 2, quarter = warp & 3;
        const int row = t * 128 + quarter * 32 + lane;                 // query row inside the sequence (0..255)
        const uint32_t treg = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
#if ATC_PINGPONG
        if (t == 1 && round == 0) named_bar_arrive(1, 256);            // warpgroup 0 takes the first turn in pass 2 (the hand-over
                                                                       // after a round's last item is the first turn of the next round)
#endif
        // ---- the 257th column and row for this warp's 32 query rows / 32 keys (rows base .. base+31 of the Q and K tiles of local
        // item `j`): m16n8k16 tiles with the class-token vector as the single live column of B.  Returns s256 of this thread's row.
        // Runs one item AHEAD (for item j+1 while item j's P V MMAs are in flight): on the warpgroup's serial chain between two
        // items it cost ~1.8 k clk per item (profiles/r02d_attn_trace.txt).
        auto class_pass = [&](uint32_t j) -> float {
            const int st = j & 1;
            const uint32_t sq = smem0 + st * ATC_STAGE_BYTES, sk = sq + ATC_TILE_BYTES;
            const uint32_t qc = sq + 3 * ATC_TILE_BYTES, kc = qc + ATC_ROW_BYTES;
            mbar_wait(&fullqk[st], (j >> 1) & 1);
            const int g = lane >> 2, t4 = lane & 3, base_row = row - lane;
            float ccol[2][4], crow[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) { ccol[mt][i] = 0.f; crow[mt][i] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                // B fragments: b0 = dims 16kk+2t4,+1 ; b1 = dims 16kk+8+2t4,+1 (column n = g = 0 only)
                const uint32_t kb0 = g == 0 ? lds32(kc + (kk * 8 + t4) * 4) : 0u, kb1 = g == 0 ? lds32(kc + (kk * 8 + 4 + t4) * 4) : 0u;
                const uint32_t qb0 = g == 0 ? lds32(qc + (kk * 8 + t4) * 4) : 0u, qb1 = g == 0 ? lds32(qc + (kk * 8 + 4 + t4) * 4) : 0u;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int r = base_row + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    uint32_t a[4];
                    ldsm_x4(a, sw128(sq, r, kk * 2 + (lane >> 4)));
                    mma_bf16_16816(ccol[mt], a, kb0, kb1);                 // Q[r] . K[256]
                    ldsm_x4(a, sw128(sk, r, kk * 2 + (lane >> 4)));
                    mma_bf16_16816(crow[mt], a, qb0, qb1);                 // K[r] . Q[256]
                }
            }
            if (t4 == 0) {
                float* dst = scls + st * 256 + base_row + g;
                dst[0] = crow[0][0]; dst[8] = crow[0][2]; dst[16] = crow[1][0]; dst[24] = crow[1][2];
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&scls_ready[st]); mbar_arrive(&emptyqk[st]); }
            // row base + lane = m-tile lane >> 4, fragment row lane & 15: held by lane 4 * (lane & 7) as c[0] (rows 0..7) or c[2]
            const int src = (lane & 7) * 4;
            const float v00 = __shfl_sync(0xffffffffu, ccol[0][0], src), v02 = __shfl_sync(0xffffffffu, ccol[0][2], src);
            const float v10 = __shfl_sync(0xffffffffu, ccol[1][0], src), v12 = __shfl_sync(0xffffffffu, ccol[1][2], src);
            return (lane & 16) ? ((lane & 8) ? v12 : v10) : ((lane & 8) ? v02 : v00);
        };
        // The item loop exists once per mode (exact row maximum / class-key shift): as a run-time flag the exact pass was
        // if-converted into ~300 predicated-off instructions per item and the "fast" round ran 38 % SLOWER than the exact kernel.
        auto item_loop = [&](auto exact_c) {
        constexpr bool kExact = decltype(exact_c)::value;
        float s256 = k1 > k0 ? class_pass(k0) : 0.f;
        for (int it = k0; it < k1; ++it) {
            const int item = item_of(it);
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1, ip = it & 1;
            const int seq = item / p.H, head = item - seq * p.H;
            const uint32_t vc = smem0 + st * ATC_STAGE_BYTES + 3 * ATC_TILE_BYTES + 2 * ATC_ROW_BYTES;
            mbar_wait(&s_full[t], ip);
            tc_fence_after();
            if (quarter == 0 && lane == 0) ATC_EV(3 + t, 0, it);
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            uint32_t va[32], vb[32];
            tmem_ld_32x32(treg, va);
            tmem_ld_wait();
            if constexpr (kExact) {
                // pass 1: row max over keys 0..255.  The load of chunk c+1 is in flight while chunk c is reduced.
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_32x32(treg + (c + 1) * 32, vb);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) m4[(j >> 1) & 3] = fmax3(m4[(j >> 1) & 3], __uint_as_float(va[j]), __uint_as_float(va[j + 1]));
                    tmem_ld_wait();
                    tmem_ld_32x32(treg + ((c + 2) & 7) * 32, va);      // wraps to chunk 0: first chunk of pass 2
#pragma unroll
                    for (int j = 0; j < 32; j += 2) m4[(j >> 1) & 3] = fmax3(m4[(j >> 1) & 3], __uint_as_float(vb[j]), __uint_as_float(vb[j + 1]));
                    tmem_ld_wait();
                }
            }
            // (fast round: the shift is the class-key score alone -- see ATC_FASTMAX)
            const float nms = -fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), s256) * p.sl2;
            if (quarter == 0 && lane == 0) ATC_EV(3 + t, 1, it);

#if ATC_PINGPONG
            named_bar_sync(1 + t, 256);                                // my turn on the MUFU pipe
#endif
            // pass 2: p = exp2(s * sl2 - max * sl2) -> bf16 pairs written over S columns already consumed
            float sum0 = 0.f, sum1 = 0.f;
            auto exp_store = [&](const uint32_t (&v)[32], int c) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x0, x1;
                    ffma2(x0, x1, __uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), p.sl2, p.sl2, nms, nms);
                    if (ATC_POLY > 0 && ((j * ATC_POLY) & 15) < ATC_POLY) exp2_poly2(x0, x1);
                    else { x0 = fast_exp2(x0); x1 = fast_exp2(x1); }
                    fadd2(sum0, sum1, sum0, sum1, x0, x1);
                    pk[j] = pack2_bf16(x0, x1);
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
#if ATC_PINGPONG
            named_bar_arrive(1 + (t ^ 1), 256);                        // hand the MUFU pipe to the other warpgroup
#endif
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            if (quarter == 0 && lane == 0) ATC_EV(3 + t, 2, it);
#if ATC_FASTMAX
            if constexpr (!kExact) {   // a row whose sum left the fp32 range (or met an inf / NaN) was shifted too little: redo the item exactly
                const bool bad = !(sum0 + sum1 < 1e30f);
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&retry_mask[it >> 5], 1u << (it & 31));
            }
#endif
            const float s256_next = it + 1 < k1 ? class_pass(it + 1) : 0.f;   // while this item's P V MMAs run
            // epilogue: (O + p256 * V[256]) / (l + p256) -> bf16 row
            const float e256 = fast_exp2(fmaf(s256, p.sl2, nms));
            const float inv = 1.0f / (sum0 + sum1 + e256);
            const float ei = e256 * inv;
            // Output: each thread owns one 128-byte row of the head's 64 columns and writes it with four 256-bit stores (whole
            // 32-byte sectors): no shared-memory staging, so nothing of the stage outlives the item.
            mbar_wait(&fullv[st], ph);                                 // (long complete) makes the TMA-written V[256] row visible
            uint32_t vcw[32];                                          // V[256][0..63] as bf16 pairs, same for every row
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 vv = lds128(vc + c * 16);
                vcw[4 * c] = vv.x; vcw[4 * c + 1] = vv.y; vcw[4 * c + 2] = vv.z; vcw[4 * c + 3] = vv.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyv[st]);                   // this warp's last read of the stage
            mbar_wait(&o_full[t], ip);
            tc_fence_after();
            if (quarter == 0 && lane == 0) ATC_EV(3 + t, 3, it);
            tmem_ld_32x32(treg + 128, va);                             // O columns 0..31
            tmem_ld_32x32(treg + 160, vb);                             // O columns 32..63
            tmem_ld_wait();
            tc_fence_before();                                         // O is in registers: the TMEM tile can take the next item's S now
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_free[t]);
            __nv_bfloat16* orow = p.out + ((size_t)seq * S + row) * p.D + head * 64;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t (&vo)[32] = hh ? vb : va;
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t w = vcw[hh * 16 + j];
                    const float a = fmaf(__uint_as_float(vo[2 * j]), inv, bf_lo(w) * ei);
                    const float b = fmaf(__uint_as_float(vo[2 * j + 1]), inv, bf_hi(w) * ei);
                    o[j] = pack2_bf16(a, b);
                }
                st_global_v8(orow + hh * 32, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                st_global_v8(orow + hh * 32 + 16, o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]);
            }
            if (quarter == 0 && lane == 0) ATC_EV(3 + t, 4, it);
            s256 = s256_next;
        }
        };
#if ATC_FASTMAX
        if (exact) item_loop(std::true_type{}); else item_loop(std::false_type{});
#else
        item_loop(std::true_type{});
#endif
    }
#if ATC_FASTMAX
    if (round == 0) {       // compact the flagged items (normally none) into the second round's sequence
        __syncwarp();
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            for (int w = 0; w < ATC_RETRY_CAP / 32; ++w)
                for (uint32_t m = retry_mask[w]; m; m &= m - 1) retry_list[n++] = (uint16_t)(w * 32 + __ffs(m) - 1);
            *retry_count = n;
        }
        __syncthreads();
        n_retry = fast_round ? *retry_count : 0;
    }
#endif
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
