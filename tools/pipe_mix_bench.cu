// Which instructions of the DP cell body share the ALU pipe, and which can run beside it?
// Each kernel issues an independent-chain mix of two instruction kinds (A : B = 1 : 1, or A alone) from 8 warps
// per SMSP and reports warp-instructions per clock per SMSP (CUDA-event time x the maximum SM clock; clock64 was found to tick at
// a different rate than the SM issue clock on this part). Two kinds that share a pipe of rate 1/2 per clock together reach 0.5; kinds on
// different pipes together approach 1.0 (the issue limit).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix_bench pipe_mix_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ILP = 8, ITERS = 32768;

enum Kind { K_NONE, K_VIADDMNMX, K_VIADD16, K_VIMNMX3, K_LOP3, K_PRMT, K_IMAD, K_HMNMX2, K_HSET2, K_HADD2, K_IADD3, K_SEL, K_MOV, K_VIMNMX3U };

template <int K>
__device__ __forceinline__ void op(unsigned& x, unsigned y, unsigned z, unsigned one) {
    if (K == K_VIADDMNMX)      asm volatile("{ .reg .b32 t; add.s16x2 t, %0, %1; max.s16x2 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_VIADD16)   asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    else if (K == K_VIMNMX3)   asm volatile("{ .reg .b32 t; max.s16x2 t, %0, %1; max.s16x2 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_VIMNMX3U)  asm volatile("{ .reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_LOP3)      asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_PRMT)      asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_IMAD)      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(one), "r"(y));
    else if (K == K_HMNMX2)    asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    else if (K == K_HSET2)     asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    else if (K == K_HADD2)     asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    else if (K == K_IADD3)     asm volatile("{ .reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_SEL)       asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p; }" : "+r"(x) : "r"(y), "r"(z));
    else if (K == K_MOV)       asm volatile("mad.lo.u32 %0, %1, %2, 0;" : "+r"(x) : "r"(x), "r"(one));
}

template <int A, int B>
__global__ void __launch_bounds__(256) k_mix(unsigned* out, unsigned a0, unsigned b0, unsigned c0, unsigned one, long long* cycles) {
    unsigned x[ILP], w[ILP], y = b0 + threadIdx.x, z = c0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a0 + i * 0x00010001u + threadIdx.x; w[i] = a0 ^ (i * 0x00030005u); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            op<A>(x[i], y, z, one);
            if (B != K_NONE) op<B>(w[i], y, z, one);
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Row { const char* name; double ipc; };

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount, nthr = 256, bps = 4, nblk = nsm * bps;   // 32 warps per SM = 8 per SMSP
    unsigned* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, (size_t)nblk * nthr * 4)); CK(cudaMalloc(&d_cyc, nblk * 8));
    std::vector<Row> rows;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double clk_hz = (double)p.clockRate * 1e3;      /* max SM clock; bench.py samples the real clock under load */
#define RUN(A, B, NAME) { \
        for (int r = 0; r < 2; ++r) k_mix<A, B><<<nblk, nthr>>>(d_out, 0x00010002u, 0x00030001u, 0x00000005u, 1u, d_cyc); \
        CK(cudaDeviceSynchronize()); \
        CK(cudaEventRecord(e0)); \
        for (int r = 0; r < 3; ++r) k_mix<A, B><<<nblk, nthr>>>(d_out, 0x00010002u, 0x00030001u, 0x00000005u, 1u, d_cyc); \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3; \
        const double n_instr = (double)ILP * ITERS * ((B) == K_NONE ? 1 : 2) * (nthr / 32) * bps / 4.0;  /* warp-instr per SMSP */ \
        rows.push_back({NAME, n_instr / (ms * 1e-3 * clk_hz)}); }
    RUN(K_VIADDMNMX, K_NONE, "VIADDMNMX.S16x2");
    RUN(K_VIADD16, K_NONE, "VIADD.16x2");
    RUN(K_VIMNMX3, K_NONE, "VIMNMX3.S16x2");
    RUN(K_VIMNMX3U, K_NONE, "VIMNMX3.U16x2");
    RUN(K_LOP3, K_NONE, "LOP3");
    RUN(K_PRMT, K_NONE, "PRMT");
    RUN(K_IMAD, K_NONE, "IMAD");
    RUN(K_HMNMX2, K_NONE, "HMNMX2");
    RUN(K_HSET2, K_NONE, "HSET2");
    RUN(K_HADD2, K_NONE, "HADD2");
    RUN(K_IADD3, K_NONE, "IADD3");
    RUN(K_SEL, K_NONE, "SEL(+ISETP)");
    RUN(K_VIADDMNMX, K_VIADD16, "VIADDMNMX + VIADD.16x2");
    RUN(K_VIADDMNMX, K_VIMNMX3, "VIADDMNMX + VIMNMX3");
    RUN(K_VIADDMNMX, K_LOP3, "VIADDMNMX + LOP3");
    RUN(K_VIADDMNMX, K_PRMT, "VIADDMNMX + PRMT");
    RUN(K_VIADDMNMX, K_IMAD, "VIADDMNMX + IMAD");
    RUN(K_VIADDMNMX, K_HMNMX2, "VIADDMNMX + HMNMX2");
    RUN(K_VIADDMNMX, K_HSET2, "VIADDMNMX + HSET2");
    RUN(K_VIADDMNMX, K_HADD2, "VIADDMNMX + HADD2");
    RUN(K_VIADDMNMX, K_IADD3, "VIADDMNMX + IADD3");
    RUN(K_IMAD, K_HMNMX2, "IMAD + HMNMX2");
    RUN(K_IMAD, K_HSET2, "IMAD + HSET2");
    RUN(K_HMNMX2, K_HSET2, "HMNMX2 + HSET2");
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"warps_per_smsp\": %d, \"ilp\": %d, \"clock_khz\": %d, \"unit\": \"warp-instructions per clock per SMSP at the max SM clock (CUDA-event timed)\", \"results\": [\n", p.name, nsm, nthr / 32 * bps / 4, ILP, p.clockRate);
    for (size_t i = 0; i < rows.size(); ++i) printf("  {\"mix\": \"%s\", \"ipc_per_smsp\": %.3f}%s\n", rows[i].name, rows[i].ipc, i + 1 < rows.size() ? "," : "");
    printf("]}\n");
    return 0;
}
