/* ORACLE -- test infrastructure, not product code.
 *
 * Plain-C restatement of one step of the reference's token select path
 * (modeling/modules/sampling.py:90-131 of the reference): classifier-free-guidance combine, softmax over the
 * per-group vocabulary, Categorical sample (== argmax(p_hat / q), q ~ Exp(1), torch.multinomial n=1 fast path),
 * confidence = log p[tok] + gumbel * mult, k-th-smallest threshold over the 512 (position, group) slots of a
 * sample with the batch-global k taken from sample 0 (sampling.py:109,123-126), and the re-mask.
 *
 * Floating-point contract (shared with the CUDA kernel, stated in DESIGN.md "select arithmetic"): every
 * operation is an IEEE-754 binary32 round-to-nearest-even add / mul / div / fma in a FIXED order; exp and log
 * are the polynomial kernels below built only from those operations.  Build with -ffp-contract=off so the
 * compiler neither fuses nor splits anything.  Under that contract the CUDA kernel and this file agree
 * bit-for-bit on every output; against the reference (which uses torch's vectorised exp/log) the token outputs
 * agree except on measure-zero near-ties -- tests/test_oracle.py checks exact token equality on the golden
 * traces recorded from the reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits_from_f(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* exp(x) for x <= 0.  x < -87 flushes to 0 (result would be below the normal range). */
float mbo_expf(float x) {
    if (!(x >= -87.0f)) return 0.0f;
    float t = x * 1.44269504088896341f;
    float n = rintf(t);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    int ni = (int)n;                                   /* in [-126, 0] */
    float s = f_from_bits((uint32_t)(ni + 127) << 23);
    return y * s;
}

/* log(x) for finite x >= 0 (x == 0 -> -inf); handles subnormals. */
float mbo_logf(float x) {
    if (x == 0.0f) return -INFINITY;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t u = bits_from_f(x);
    e += (int)(u >> 23) - 126;                         /* x = m * 2^e, m in [0.5, 1) */
    float m = f_from_bits((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m; }
    m = m - 1.0f;
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = (p * m) * z;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* sum in the kernel's order: lane l (0..31) adds its strided elements l, l+32, ... sequentially, then a
 * 5-level xor butterfly (offsets 16,8,4,2,1).  fp add is commutative, so every lane ends with the same value. */
static float warp_order_sum(const float* e, int v) {
    float lane[32];
    for (int l = 0; l < 32; ++l) {
        float s = 0.0f;
        int first = 1;
        for (int j = l; j < v; j += 32) { s = first ? e[j] : s + e[j]; first = 0; }
        lane[l] = s;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof(lane));
    }
    return lane[0];
}

static int gt_nanmax(float b, float a) { /* "b beats a" with NaN treated as the maximum (torch.argmax) */
    if (isnan(a)) return 0;
    if (isnan(b)) return 1;
    return b > a;
}

/* total order key for the k-th-smallest threshold (torch.sort ascending; NaN last) */
static uint32_t sort_key(float f) {
    if (isnan(f)) return 0xffffffffu;
    uint32_t u = bits_from_f(f);
    if (u == 0x80000000u) u = 0;                       /* -0 == +0 */
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

/* One select step for a batch.
 *   logits_c / logits_u : [B, rows_per_seq_stride, m*V] fp32 rows; row (b, pos) at ((b*seq_stride + pos) * m*V);
 *                         logits_u may be NULL (no guidance, sampling.py:100-101)
 *   q      : [B*n*m, V] Exp(1) draws;  gumbel : [B, n, m] raw Gumbel(0,1) draws
 *   tokens_in [B,n,m] int64 -> predicted [B,n,m], tokens_out [B,n,m]
 * Returns the k used.
 */
int mbo_select_step(const float* logits_c, const float* logits_u, float scale, float temperature,
                    const float* q, const float* gumbel, float randomize_temperature, float one_minus_progress,
                    float mask_len,
                    const int64_t* tokens_in, int64_t* predicted, int64_t* tokens_out,
                    int B, int n, int m, int V, int seq_stride, int64_t mask_token) {
    const int slots = n * m;
    /* k from sample 0 (sampling.py:109,123-124): clamp(mask_len, 1, num_masked-1) with torch.clamp semantics
     * (min applied first, then max wins when min > max) */
    int num_masked0 = 0;
    for (int s = 0; s < slots; ++s) num_masked0 += (tokens_in[s] == mask_token);
    float kf = mask_len < 1.0f ? 1.0f : mask_len;
    float hi = (float)(num_masked0 - 1);
    if (kf > hi) kf = hi;
    int k = (int)kf;
    int kth = k - 1;
    if (kth < 0) kth += slots;                         /* python negative index on sorted[:, k-1] */

    float conf[4096];
    float e[1024];
    for (int b = 0; b < B; ++b) {
        for (int s = 0; s < slots; ++s) {
            int pos = s / m, g = s % m;
            const float* lc = logits_c + ((size_t)(b * seq_stride + pos) * m + g) * V;
            const float* lu = logits_u ? logits_u + ((size_t)(b * seq_stride + pos) * m + g) * V : 0;
            const float* qr = q + ((size_t)b * slots + s) * V;
            int64_t tin = tokens_in[(size_t)b * slots + s];
            int masked = (tin == mask_token);
            float mx = -INFINITY;
            for (int j = 0; j < V; ++j) {
                float x = lc[j];
                if (lu) { float d = lc[j] - lu[j]; float t = scale * d; x = lc[j] + t; }
                x = x / temperature;
                e[j] = x;
                if (x > mx) mx = x;
            }
            for (int j = 0; j < V; ++j) e[j] = mbo_expf(e[j] - mx);
            float sum = warp_order_sum(e, V);
            for (int j = 0; j < V; ++j) e[j] = e[j] / sum;         /* probabilities */
            float sum2 = warp_order_sum(e, V);                     /* Categorical renormalisation */
            int best = 0; float bestv = 0.0f;
            for (int j = 0; j < V; ++j) {
                float r = (e[j] / sum2) / qr[j];
                if (j == 0 || gt_nanmax(r, bestv)) { best = j; bestv = r; }
            }
            int64_t tok = masked ? (int64_t)best : tin;
            predicted[(size_t)b * slots + s] = tok;
            float c;
            if (masked) {
                /* sampling.py:117-118: noise = (g * rt) * (1 - progress), two fp32 multiplies, then log + noise */
                float nz = gumbel[(size_t)b * slots + s] * randomize_temperature;
                nz = nz * one_minus_progress;
                c = mbo_logf(e[tok]) + nz;
            }
            else c = INFINITY;
            conf[s] = c;
        }
        /* k-th smallest (0-based index kth) under the total order, ties broken by slot index */
        float thr = 0.0f;
        for (int i = 0; i < slots; ++i) {
            uint32_t ki = sort_key(conf[i]);
            int rank = 0;
            for (int j = 0; j < slots; ++j) {
                uint32_t kj = sort_key(conf[j]);
                rank += (kj < ki) || (kj == ki && j < i);
            }
            if (rank == kth) thr = conf[i];
        }
        for (int s = 0; s < slots; ++s) {
            int should_mask = conf[s] <= thr;
            tokens_out[(size_t)b * slots + s] = should_mask ? mask_token : predicted[(size_t)b * slots + s];
        }
    }
    return k;
}
