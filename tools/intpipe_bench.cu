// Integer-pipe microbenchmark for the DP roofline denominator.
// Measures sustained warp-instruction throughput (ops/clk/SM) of the packed s16x2
// DPX / video instructions the alignment kernel is built from, on the GPU it runs on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o intpipe_bench intpipe_bench.cu
// Output: one JSON object on stdout (committed as profiles/intpipe_peaks_rNN.json).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 4096;

enum Op { OP_VIADD16X2, OP_VIMNMX16X2, OP_VIMNMX3, OP_VIMNMX3_RELU, OP_VIADDMNMX, OP_HSET2, OP_LOP3, OP_PRMT, OP_IMAD, OP_IADD3, OP_SHFL,
          OP_VIADDMNMX32, OP_VIMNMX3_32, OP_CELL2P, OP_CELL1P, OP_COUNT };

static __device__ __forceinline__ unsigned heqmask(unsigned a, unsigned b) {
    __half2 ha = *reinterpret_cast<__half2*>(&a), hb = *reinterpret_cast<__half2*>(&b);
    return __heq2_mask(ha, hb);
}

template <int OP>
__global__ void __launch_bounds__(256) k_bench(unsigned* out, unsigned a0, unsigned b0, unsigned c0, long long* cycles) {
    unsigned x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a0 + i * 0x00010001u + threadIdx.x; y[i] = b0 ^ (i * 0x00030005u); z[i] = c0 + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == OP_VIADD16X2)        x[i] = __vadd2(x[i], y[i]);
            else if (OP == OP_VIMNMX16X2)  x[i] = __vmaxs2(x[i], y[i]) ^ 0; 
            else if (OP == OP_VIMNMX3)     x[i] = __vimax3_s16x2(x[i], y[i], z[i]);
            else if (OP == OP_VIMNMX3_RELU) x[i] = __vimax3_s16x2_relu(x[i], y[i], z[i]);
            else if (OP == OP_VIADDMNMX)   x[i] = __viaddmax_s16x2(x[i], y[i], z[i]);
            else if (OP == OP_HSET2)       x[i] = heqmask(x[i], y[i]);
            else if (OP == OP_LOP3)        x[i] = (x[i] & y[i]) ^ z[i];
            else if (OP == OP_PRMT)        x[i] = __byte_perm(x[i], y[i], 0x5432);
            else if (OP == OP_IMAD)        x[i] = x[i] * y[i] + z[i];
            else if (OP == OP_IADD3)       x[i] = x[i] + y[i] + z[i];
            else if (OP == OP_SHFL)        x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
            else if (OP == OP_VIADDMNMX32) x[i] = (unsigned)__viaddmax_s32((int)x[i], (int)y[i], (int)z[i]);
            else if (OP == OP_VIMNMX3_32)  x[i] = (unsigned)__vimax3_s32((int)x[i], (int)y[i], (int)z[i]);
        }
        if (OP == OP_VIMNMX16X2) {  // keep the max chain from collapsing
#pragma unroll
            for (int i = 0; i < ILP; ++i) y[i] += 0x00010001u;
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// One packed DP cell-pair body, exactly the op mix of the alignment kernel's inner slot
// (two-piece affine, local): used to measure the attainable cell rate of the body alone.
template <bool TWO_PIECE>
__global__ void __launch_bounds__(256) k_cell(unsigned* out, unsigned a0, unsigned b0, unsigned c0, long long* cycles) {
    unsigned H[ILP], E1[ILP], E2[ILP], F1[ILP], F2[ILP], P1[ILP], P2[ILP], Q[ILP], T[ILP], best = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { H[i] = a0 + i; E1[i] = b0 + i; E2[i] = b0 ^ i; F1[i] = c0 + i; F2[i] = c0 ^ i; P1[i] = a0 ^ i; P2[i] = a0 - i; Q[i] = 0x3c003d00u + (threadIdx.x & 3) * 0x100u; T[i] = 0x3c003d00u + i * 0x100u; }
    const unsigned NE1 = 0xfffefffeu, NE2 = 0xffffffffu, NOE1 = 0xfffafffau, NOE2 = 0xffe7ffe7u, MAT = 0x00020002u, MIS = 0xfffcfffcu;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            unsigned m = heqmask(Q[i], T[i]);
            unsigned S = (m & MAT) | (~m & MIS);
            unsigned e1 = __viaddmax_s16x2(E1[i], NE1, P1[i]);
            unsigned f1 = __viaddmax_s16x2(F1[i], NE1, P1[(i + 1) % ILP]);
            unsigned h = __vadd2(H[i], S);
            unsigned hh;
            if (TWO_PIECE) {
                unsigned e2 = __viaddmax_s16x2(E2[i], NE2, P2[i]);
                unsigned f2 = __viaddmax_s16x2(F2[i], NE2, P2[(i + 1) % ILP]);
                hh = __vimax3_s16x2(h, e1, f1);
                hh = __vimax3_s16x2_relu(hh, e2, f2);
                E2[i] = e2; F2[i] = f2;
                P2[i] = __vadd2(hh, NOE2);
            } else {
                hh = __vimax3_s16x2_relu(h, e1, f1);
            }
            E1[i] = e1; F1[i] = f1;
            P1[i] = __vadd2(hh, NOE1);
            H[i] = hh;
            best = __vmaxs2(best, hh);
            T[i] ^= (hh & 0x100u);
        }
    }
    long long t1 = clock64();
    unsigned acc = best;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= H[i] ^ E1[i] ^ F1[i] ^ E2[i] ^ F2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Variant: H is kept with a +0x1000 bias per half so that H - const never borrows across the halves and the
// three H-derived adds become plain 32-bit IMADs (FMA pipe) instead of VIADD.16x2 (ALU pipe); the zero floor
// becomes a max with the bias (one extra VIMNMX).
__device__ __forceinline__ unsigned imad_add(unsigned a, unsigned one, unsigned b) {
    unsigned d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d;
}
__global__ void __launch_bounds__(256) k_cell_biased(unsigned* out, unsigned a0, unsigned b0, unsigned c0, unsigned one, long long* cycles) {
    unsigned H[ILP], E1[ILP], E2[ILP], F1[ILP], F2[ILP], P1[ILP], P2[ILP], Q[ILP], T[ILP], best = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { H[i] = a0 + i; E1[i] = b0 + i; E2[i] = b0 ^ i; F1[i] = c0 + i; F2[i] = c0 ^ i; P1[i] = a0 ^ i; P2[i] = a0 - i; Q[i] = 0x3c003d00u + (threadIdx.x & 3) * 0x100u; T[i] = 0x3c003d00u + i * 0x100u; }
    const unsigned NE1 = 0xfffefffeu, NE2 = 0xffffffffu, NOE1 = 0xfffafffau, NOE2 = 0xffe7ffe7u, MAT = 0x00020002u, AB = 0x00060006u, BIAS = 0x10001000u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            unsigned m = heqmask(Q[i], T[i]);
            unsigned S = MAT - (~m & AB);                       // +2 / -4 per half as a carry-free 32-bit value relative to the bias
            unsigned e1 = __viaddmax_s16x2(E1[i], NE1, P1[i]);
            unsigned f1 = __viaddmax_s16x2(F1[i], NE1, P1[(i + 1) % ILP]);
            unsigned e2 = __viaddmax_s16x2(E2[i], NE2, P2[i]);
            unsigned f2 = __viaddmax_s16x2(F2[i], NE2, P2[(i + 1) % ILP]);
            unsigned h = imad_add(H[i], one, S);
            unsigned hh = __vimax3_s16x2(h, e1, f1);
            hh = __vimax3_s16x2(hh, e2, f2);
            hh = __vmaxs2(hh, BIAS);
            E1[i] = e1; F1[i] = f1; E2[i] = e2; F2[i] = f2;
            P1[i] = imad_add(hh, one, NOE1);
            P2[i] = imad_add(hh, one, NOE2);
            H[i] = hh;
            best = __vmaxs2(best, hh);
            T[i] ^= (hh & 0x100u);
        }
    }
    long long t1 = clock64();
    unsigned acc = best;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= H[i] ^ E1[i] ^ F1[i] ^ E2[i] ^ F2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Res { const char* name; double ops_per_clk_sm; double tops; double ms; };

template <typename F>
static int run(const char* name, F launch, int nblk, int nthr, double ops_per_thread, int nsm, std::vector<Res>& res, long long* d_cycles) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    std::vector<long long> cyc(nblk);
    CK(cudaMemcpy(cyc.data(), d_cycles, nblk * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : cyc) avg += (double)c; avg /= nblk;
    // per-SM rate: blocks resident on one SM run concurrently; all blocks are co-resident (nblk = nsm * blocks_per_sm)
    double blocks_per_sm = (double)nblk / nsm;
    double warp_ops_per_block = ops_per_thread * nthr;        // lane-ops
    double ops_clk_sm = warp_ops_per_block * blocks_per_sm / avg;  // lane-ops per clk per SM
    double tops = ops_per_thread * nthr * (double)nblk / (ms * 1e-3) / 1e12;
    res.push_back({name, ops_clk_sm, tops, ms});
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    const int nthr = 256, bps = 4;   // 1024 threads = 32 warps per SM, 8 per SMSP
    const int nblk = nsm * bps;
    unsigned* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, (size_t)nblk * nthr * 4)); CK(cudaMalloc(&d_cyc, nblk * 8));
    std::vector<Res> res;
    double opt = (double)ILP * ITERS;
#define RUN(OPN, NAME) if (run(NAME, [&]{ k_bench<OPN><<<nblk, nthr>>>(d_out, 0x00010002u, 0x00030001u, 0x00000005u, d_cyc); }, nblk, nthr, opt, nsm, res, d_cyc)) return 1;
    RUN(OP_VIADD16X2, "VIADD.16x2");
    RUN(OP_VIMNMX16X2, "VIMNMX.S16x2(+VIADD)");
    RUN(OP_VIMNMX3, "VIMNMX3.S16x2");
    RUN(OP_VIMNMX3_RELU, "VIMNMX3.S16x2.RELU");
    RUN(OP_VIADDMNMX, "VIADDMNMX.S16x2");
    RUN(OP_HSET2, "HSET2.BM.EQ");
    RUN(OP_LOP3, "LOP3");
    RUN(OP_PRMT, "PRMT");
    RUN(OP_IMAD, "IMAD");
    RUN(OP_IADD3, "IADD3");
    RUN(OP_SHFL, "SHFL.UP");
    RUN(OP_VIADDMNMX32, "VIADDMNMX.S32");
    RUN(OP_VIMNMX3_32, "VIMNMX3.S32");
    // cell bodies: ops_per_thread counts CELLS (2 per packed slot)
    if (run("cell_two_piece(cells)", [&]{ k_cell<true><<<nblk, nthr>>>(d_out, 0x00010002u, 0x00030001u, 0x00000005u, d_cyc); }, nblk, nthr, 2.0 * opt, nsm, res, d_cyc)) return 1;
    if (run("cell_one_piece(cells)", [&]{ k_cell<false><<<nblk, nthr>>>(d_out, 0x00010002u, 0x00030001u, 0x00000005u, d_cyc); }, nblk, nthr, 2.0 * opt, nsm, res, d_cyc)) return 1;
    if (run("cell_two_piece_biased_imad(cells)", [&]{ k_cell_biased<<<nblk, nthr>>>(d_out, 0x10011002u, 0x00030001u, 0x00000005u, 1u, d_cyc); }, nblk, nthr, 2.0 * opt, nsm, res, d_cyc)) return 1;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"threads_per_sm\": %d, \"ilp\": %d, \"results\": [\n", p.name, nsm, p.clockRate, nthr * bps, ILP);
    for (size_t i = 0; i < res.size(); ++i)
        printf("  {\"op\": \"%s\", \"lane_ops_per_clk_per_sm\": %.2f, \"tera_lane_ops_per_s\": %.3f, \"ms\": %.4f}%s\n", res[i].name, res[i].ops_per_clk_sm, res[i].tops, res[i].ms, i + 1 < res.size() ? "," : "");
    printf("]}\n");
    return 0;
}
