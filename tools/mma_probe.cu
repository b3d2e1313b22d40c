// Micro-probe: issue-to-completion time (SM clocks) of chains of tcgen05.mma of the shapes an attention item is built from, issued
// the way the kernels issue them (elect.sync region, compile-time shapes: operands in uniform registers).  Data is zeros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu && gpurun -- tools/mma_probe
#include <cstdio>
#include "../maskbit_b200/csrc/ptx.cuh"

using namespace mb;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t sdesc_mn128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)(1024 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// TS: A from TMEM.  NI instructions round-robin over NACC accumulators.  BMN: B MN-major (V), else K-major (K).
template <bool TS, int M, int N, int NI, int NACC, bool BMN>
__device__ __forceinline__ void run_case(uint32_t tm, uint32_t sa, uint32_t sb, uint64_t* bar, uint32_t& phase, long long* out, int reps) {
    constexpr uint32_t idesc = make_idesc(1, M, N) | (BMN ? (1u << 16) : 0u);
    long long best = 1ll << 60, best_issue = 0;
    for (int r = 0; r < reps; ++r) {
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t d = tm + 256 + (i % NACC) * 64;
            const uint64_t bd = BMN ? sdesc_mn128(sb) + (uint64_t)((i & 15) * 128) : make_sdesc_k128(sb) + 2 * (i & 3);
            if (TS) umma_ts(d, tm + 8 * (i & 15), bd, idesc, i >= NACC);
            else umma_f16(d, make_sdesc_k128(sa) + 2 * (i & 3), bd, idesc, i >= NACC);
        }
        umma_commit(bar);
        const long long t1 = clock64();
        mbar_wait(bar, phase); phase ^= 1;
        const long long t2 = clock64();
        if (t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
    }
    out[0] = best; out[1] = best_issue;
}

__global__ void __launch_bounds__(256, 1) probe_kernel(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 160 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(slot);
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    const uint32_t sa = smem_u32(base), sb = sa + 64 * 1024;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            uint32_t phase = 0;
            run_case<false, 128, 256, 4, 1, false>(tm, sa, sb, bar, phase, out + 0, reps);    // S full
            run_case<false, 128, 256, 8, 1, false>(tm, sa, sb, bar, phase, out + 2, reps);
            run_case<false, 128, 256, 16, 1, false>(tm, sa, sb, bar, phase, out + 4, reps);
            run_case<false, 128, 128, 4, 1, false>(tm, sa, sb, bar, phase, out + 6, reps);    // S unit
            run_case<false, 128, 128, 16, 1, false>(tm, sa, sb, bar, phase, out + 8, reps);
            run_case<true, 128, 64, 16, 1, true>(tm, sa, sb, bar, phase, out + 10, reps);     // PV now
            run_case<true, 128, 64, 16, 2, true>(tm, sa, sb, bar, phase, out + 12, reps);
            run_case<true, 128, 64, 8, 1, true>(tm, sa, sb, bar, phase, out + 14, reps);
            run_case<true, 128, 64, 1, 1, true>(tm, sa, sb, bar, phase, out + 16, reps);
            run_case<false, 128, 256, 1, 1, false>(tm, sa, sb, bar, phase, out + 18, reps);
            run_case<false, 128, 64, 16, 1, true>(tm, sa, sb, bar, phase, out + 20, reps);    // PV with P in smem
            run_case<true, 128, 256, 4, 1, false>(tm, sa, sb, bar, phase, out + 22, reps);    // S with Q in TMEM
            run_case<true, 128, 256, 16, 1, false>(tm, sa, sb, bar, phase, out + 24, reps);
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
    const char* names[] = {
        "SS M128 N256 x4   (S of one tile)                ideal 512",
        "SS M128 N256 x8                                  ideal 1024",
        "SS M128 N256 x16                                 ideal 2048",
        "SS M128 N128 x4   (S, 128-key half)              ideal 256",
        "SS M128 N128 x16                                 ideal 1024",
        "TS M128 N64 x16 1 acc, B mn-major (PV)           ideal 512",
        "TS M128 N64 x16 2 acc                            ideal 512",
        "TS M128 N64 x8  1 acc                            ideal 256",
        "TS M128 N64 x1  (latency of one)",
        "SS M128 N256 x1 (latency of one)",
        "SS M128 N64 x16, B mn-major (P from smem)        ideal 512",
        "TS M128 N256 x4  (S with Q in TMEM)              ideal 512",
        "TS M128 N256 x16                                 ideal 2048",
    };
    long long* dout;
    cudaMalloc(&dout, 64 * sizeof(long long));
    cudaMemset(dout, 0, 64 * sizeof(long long));
    const int smem = 160 * 1024 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 256, smem>>>(dout, 20);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[64];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-62s %10s %10s\n", "case (one CTA, nothing else running)", "clk total", "clk issue");
    for (int c = 0; c < 13; ++c) printf("%-62s %10lld %10lld\n", names[c], h[2 * c], h[2 * c + 1]);
    return 0;
}
