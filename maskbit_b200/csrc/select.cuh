// Per-step token select of the sampler (reference modeling/modules/sampling.py:90-131), one CTA per sample:
//   CFG combine -> softmax over the per-group vocabulary -> Categorical sample (argmax(p_hat / q), q ~ Exp(1))
//   -> confidence = log p[tok] + (gumbel * rt) * (1 - progress) -> k-th smallest of the sample's n*m confidences
//   (k = clamp(mask_len, 1, num_masked(sample 0) - 1), sampling.py:109,123-124) -> re-mask.
// Warp-shuffle integer/fp32 kernel; HBM-bound (reads the step's logits once: 2 * n*m*V*4 bytes per sample).
//
// Arithmetic contract ("select arithmetic"): IEEE binary32 RN add/mul/div/fma in a fixed order, exp/log as
// the polynomial kernels below, sums as lane-strided sequential adds followed by a 5-level xor butterfly.  The plain-C
// oracle (oracle/select_oracle.c) states the same contract independently; outputs agree bit-for-bit.
#pragma once
#include "ptx.cuh"
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mb {

__device__ __forceinline__ float sel_expf(float x) {   // x <= 0
    if (!(x >= -87.0f)) return 0.0f;
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float n = rintf(t);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    float y = __fmaf_rn(p, r2, r);
    y = __fadd_rn(y, 1.0f);
    const int ni = (int)n;
    const float s = __uint_as_float((uint32_t)(ni + 127) << 23);
    return __fmul_rn(y, s);
}

__device__ __forceinline__ float sel_logf(float x) {   // finite x >= 0
    if (x == 0.0f) return -CUDART_INF_F;
    int e = 0;
    if (x < 1.17549435e-38f) { x = __fmul_rn(x, 8388608.0f); e = -23; }
    const uint32_t u = __float_as_uint(x);
    e += (int)(u >> 23) - 126;
    float m = __uint_as_float((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = __fadd_rn(m, m); }
    m = __fadd_rn(m, -1.0f);
    const float z = __fmul_rn(m, m);
    float p = 7.0376836292e-2f;
    p = __fmaf_rn(p, m, -1.1514610310e-1f);
    p = __fmaf_rn(p, m, 1.1676998740e-1f);
    p = __fmaf_rn(p, m, -1.2420140846e-1f);
    p = __fmaf_rn(p, m, 1.4249322787e-1f);
    p = __fmaf_rn(p, m, -1.6668057665e-1f);
    p = __fmaf_rn(p, m, 2.0000714765e-1f);
    p = __fmaf_rn(p, m, -2.4999993993e-1f);
    p = __fmaf_rn(p, m, 3.3333331174e-1f);
    float y = __fmul_rn(__fmul_rn(p, m), z);
    const float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = __fadd_rn(m, y);
    r = __fmaf_rn(fe, 0.693359375f, r);
    return r;
}

__device__ __forceinline__ uint32_t sel_sort_key(float f) {   // total order of torch.sort: ascending, NaN last
    if (f != f) return 0xffffffffu;
    uint32_t u = __float_as_uint(f);
    if (u == 0x80000000u) u = 0;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Philox4x32-10 (production-mode noise; parity mode injects q / gumbel instead)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
// Production-mode noise transforms.  The uniform is built from 23 random bits as (k + 0.5) * 2^-23: every value is exactly
// representable and lies in [2^-24, 1 - 2^-24], so neither log below can see 0 or 1 (a 24-bit k * 2^-24 + 2^-25 rounds to
// 1.0f for k = 2^24 - 1: -log(1) = -0 made that token unselectable, and its Gumbel value +inf).  logf, not __logf: the fast
// intrinsic has ~2^-21 ABSOLUTE error near 1, exactly where the small Exp(1) draws that decide argmax(p / q) come from.
__device__ __forceinline__ float u01_open(uint32_t r) { return ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f); }
__device__ __forceinline__ float sel_exp1(uint32_t r) { return -logf(u01_open(r)); }                  // Exp(1), in (0, 16.7)
__device__ __forceinline__ float sel_gumbel(uint32_t r) { return -logf(-logf(u01_open(r))); }          // Gumbel(0, 1), finite

struct SelectParams {
    const float* logits_c;      // [B, seq_stride, m, V]
    const float* logits_u;      // same, or nullptr (no guidance)
    const float* q;             // [B, n*m, V] Exp(1) draws, or nullptr -> device Philox
    const float* gumbel;        // [B, n*m] raw Gumbel(0,1) draws, or nullptr -> device Philox
    const int64_t* tokens_in;   // [B, n*m]
    int64_t* predicted;         // [B, n*m]
    int64_t* tokens_out;        // [B, n*m]
    float scale, temperature, randomize_temperature, one_minus_progress, mask_len;
    int n, m, V, seq_stride;
    int64_t mask_token;
    uint64_t seed;              // Philox key (production mode)
    uint32_t step;
    int csize;                  // CTAs per sample: 1, or a thread-block cluster of 8 for small batches (one CTA per sample leaves a
                                // batch-1 step on a single SM for ~100 us): the slots are dealt over the cluster's warps, confidences
                                // and predictions land in rank 0's shared memory (DSMEM stores), rank 0 ranks and writes out
};

// VPL = V / 32 values per lane.  blockDim = 512; dynamic smem: slots * (4 + 8) bytes.
template <int VPL>
__global__ void __launch_bounds__(512) select_step_kernel(SelectParams p) {
    extern __shared__ __align__(16) uint8_t sel_smem[];
    const int slots = p.n * p.m;
    float* conf = reinterpret_cast<float*>(sel_smem);
    int64_t* pred_s = reinterpret_cast<int64_t*>(sel_smem + ((slots * 4 + 15) & ~15));
    __shared__ int s_count;
    __shared__ float s_thr;
    const int cs = p.csize;
    const int b = blockIdx.x / cs;
    const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int V = VPL * 32;
    // Cluster form: a CTA's shared memory may only be written remotely once that CTA is known to have started.  Arrive here, wait
    // just before the slot loop (the first DSMEM store) -- the barrier's latency hides behind the mask count.
    if (cs > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // programmatic dependent launch (ptx.cuh): the logits come from
    asm volatile("griddepcontrol.wait;" ::: "memory");                 // the prediction-layer GEMM

    // num_masked of sample 0 (sampling.py:109): every CTA recounts it from the step's input tokens
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int cnt = 0;
    for (int s = threadIdx.x; s < slots; s += blockDim.x) cnt += (p.tokens_in[s] == p.mask_token);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(&s_count, cnt);

    if (cs > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");    // every CTA of the cluster is running
    for (int s = (int)crank * nwarps + warp; s < slots; s += nwarps * cs) {
        const int pos = s / p.m, g = s - pos * p.m;
        const size_t off = (((size_t)b * p.seq_stride + pos) * p.m + g) * V;
        const int64_t tin = p.tokens_in[(size_t)b * slots + s];
        const bool masked = tin == p.mask_token;
        float x[VPL];
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float lc = __ldg(p.logits_c + off + lane + 32 * i);
            float v = lc;
            if (p.logits_u) {
                const float lu = __ldg(p.logits_u + off + lane + 32 * i);
                v = __fadd_rn(lc, __fmul_rn(p.scale, __fadd_rn(lc, -lu)));   // lc + scale*(lc - lu), sampling.py:99
            }
            v = __fdiv_rn(v, p.temperature);
            x[i] = v;
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            x[i] = sel_expf(__fadd_rn(x[i], -mx));
            sum = i == 0 ? x[0] : __fadd_rn(sum, x[i]);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
        float sum2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            x[i] = __fdiv_rn(x[i], sum);                      // probabilities (torch.softmax)
            sum2 = i == 0 ? x[0] : __fadd_rn(sum2, x[i]);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) sum2 = __fadd_rn(sum2, __shfl_xor_sync(0xffffffffu, sum2, o));
        // Categorical: argmax(p_hat / q), first index wins ties, NaN is the maximum
        float bestv = 0.f; int besti = 0;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int j = lane + 32 * i;
            float qv;
            if (p.q) {
                qv = __ldg(p.q + ((size_t)b * slots + s) * V + j);
            } else {
                const uint4 r = philox4x32(make_uint4((uint32_t)j, (uint32_t)s, (uint32_t)b, p.step),
                                           make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
                qv = sel_exp1(r.x);
            }
            const float r = __fdiv_rn(__fdiv_rn(x[i], sum2), qv);
            const bool beats = i == 0 || (!(bestv != bestv) && ((r != r) || r > bestv));
            if (beats) { bestv = r; besti = j; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            const bool a_nan = bestv != bestv, o_nan = ov != ov;
            bool take;
            if (a_nan && o_nan) take = oi < besti;
            else if (a_nan) take = false;
            else if (o_nan) take = true;
            else take = (ov > bestv) || (ov == bestv && oi < besti);
            if (take) { bestv = ov; besti = oi; }
        }
        const int tok = masked ? besti : (int)tin;
        // p[tok] lives in lane tok & 31, register tok >> 5
        float mine = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) if ((tok >> 5) == i) mine = x[i];
        const float ptok = __shfl_sync(0xffffffffu, mine, tok & 31);
        if (lane == 0) {
            float c = CUDART_INF_F;
            if (masked) {
                float gz;
                if (p.gumbel) gz = p.gumbel[(size_t)b * slots + s];
                else {
                    const uint4 r = philox4x32(make_uint4(0xffffffffu, (uint32_t)s, (uint32_t)b, p.step),
                                               make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
                    gz = sel_gumbel(r.x);
                }
                const float nz = __fmul_rn(__fmul_rn(gz, p.randomize_temperature), p.one_minus_progress);
                c = __fadd_rn(sel_logf(ptok), nz);
            }
            const int64_t pv = masked ? (int64_t)besti : tin;
            if (cs > 1) {   // into rank 0's copy of conf / pred_s
                asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(mapa_u32(smem_u32(conf + s), 0)), "f"(c) : "memory");
                asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(mapa_u32(smem_u32(pred_s + s), 0)), "l"(pv) : "memory");
            } else {
                conf[s] = c;
                pred_s[s] = pv;
            }
        }
    }
    if (cs > 1) {
        cluster_sync_all();          // release / acquire at cluster scope: every rank's DSMEM stores are visible to rank 0
        if (crank != 0) return;
    }
    __syncthreads();
    // k (torch.clamp semantics: lower bound first, then the upper bound wins), python index k-1 may wrap to the last
    float kf = p.mask_len < 1.0f ? 1.0f : p.mask_len;
    const float hi = (float)(s_count - 1);
    if (kf > hi) kf = hi;
    int kth = (int)kf - 1;
    if (kth < 0) kth += slots;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        const uint32_t ki = sel_sort_key(conf[i]);
        int rank = 0;
        for (int j = 0; j < slots; ++j) {
            const uint32_t kj = sel_sort_key(conf[j]);
            rank += (kj < ki) || (kj == ki && j < i);
        }
        if (rank == kth) s_thr = conf[i];
    }
    __syncthreads();
    const float thr = s_thr;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) {
        const int64_t pr = pred_s[i];
        p.predicted[(size_t)b * slots + i] = pr;
        p.tokens_out[(size_t)b * slots + i] = (conf[i] <= thr) ? p.mask_token : pr;
    }
}

}  // namespace mb
