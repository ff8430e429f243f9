// Issue-rate microbenchmark for the integer instructions the unsqueeze kernels are built from (sm_100a).
// Development aid, not product: answers "which pipe, how many cycles per warp instruction per SM sub-partition" for
// VIADD.16x2 / VIMNMX(3).S16x2 / VIADDMNMX / PRMT / LOP3 / IADD3 / SHF / IMAD / IMAD.HI and for a few mixes, so the
// instruction budget of DESIGN.md section 3.2 rests on measured numbers.   Build: nvcc -arch=sm_100a -O3 -o ubench_pipes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define NACC 8
template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
    if (OP == 0) return a + b;                                  // IADD3
    if (OP == 1) return (a ^ b) & (c | a);                      // LOP3 (one 3-input function)
    if (OP == 2) return __funnelshift_l(a, b, 7);               // SHF
    if (OP == 3) return __byte_perm(a, b, c);                   // PRMT (register selector)
    if (OP == 4) return a * b + c;                              // IMAD
    if (OP == 5) return __umulhi(a, b);                         // IMAD.HI
    if (OP == 6) return __vadd2(a, b);                          // VIADD.16x2
    if (OP == 7) return __vmins2(a, b);                         // VIMNMX.S16x2
    if (OP == 8) return __vimin3_s16x2(a, b, c);                // VIMNMX3.S16x2
    if (OP == 9) return __viaddmin_s16x2(a, b, c);              // VIADDMNMX.S16x2
    if (OP == 10) return __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));   // FFMA
    if (OP == 11) return (unsigned)min((int)a, (int)b);         // VIMNMX (32 bit)
    if (OP == 12) return a + b + c;                             // IADD3, three inputs
    if (OP == 13) return (unsigned)(((int)a) >> 3) + b;         // SHF.R.S32 + IADD (2 instr)
    return a;
}

// MIX: op A on even accumulators, op B on odd ones
template <int OPA, int OPB>
__global__ void __launch_bounds__(1024) k_bench(unsigned *out, long long *clk, int iters, unsigned seed) {
    unsigned acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                const unsigned b = acc[(i + 1) % NACC], c = acc[(i + 3) % NACC];
                acc[i] = (i & 1) ? op<OPB>(acc[i], b, c) : op<OPA>(acc[i], b, c);
            }
        }
    }
    const long long t1 = clock64();
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) x ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB>
static void run(const char *name, int instr_per_op_a, int instr_per_op_b) {
    const int blocks = 148, threads = 1024, iters = 2000;
    unsigned *out; long long *clk;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&clk, blocks * 8);
    k_bench<OPA, OPB><<<blocks, threads>>>(out, clk, 10, 12345u);
    k_bench<OPA, OPB><<<blocks, threads>>>(out, clk, iters, 12345u);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += (double)h[i];
    mean /= blocks;
    // warp instructions issued per SM sub-partition: 8 warps x iters x 4 x NACC source-level ops
    const double ops = 8.0 * iters * 4 * NACC;
    const double sass = 8.0 * iters * 4 * (NACC / 2) * (instr_per_op_a + instr_per_op_b);
    printf("%-34s cycles/source-op/SMSP %.3f   (assumed SASS instr: cycles/instr %.3f)\n", name, mean / ops, mean / sass);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0, 0>("IADD3", 1, 1);
    run<12, 12>("IADD3 3-input", 1, 1);
    run<1, 1>("LOP3", 1, 1);
    run<2, 2>("SHF", 1, 1);
    run<3, 3>("PRMT", 1, 1);
    run<4, 4>("IMAD", 1, 1);
    run<5, 5>("IMAD.HI", 1, 1);
    run<6, 6>("VIADD.16x2", 1, 1);
    run<7, 7>("VIMNMX.S16x2", 1, 1);
    run<8, 8>("VIMNMX3.S16x2", 1, 1);
    run<9, 9>("VIADDMNMX.S16x2", 1, 1);
    run<11, 11>("VIMNMX (32 bit)", 1, 1);
    run<10, 10>("FFMA", 1, 1);
    run<1, 4>("LOP3 + IMAD", 1, 1);
    run<6, 4>("VIADD.16x2 + IMAD", 1, 1);
    run<6, 1>("VIADD.16x2 + LOP3", 1, 1);
    run<6, 8>("VIADD.16x2 + VIMNMX3", 1, 1);
    run<8, 4>("VIMNMX3 + IMAD", 1, 1);
    run<3, 4>("PRMT + IMAD", 1, 1);
    run<6, 10>("VIADD.16x2 + FFMA", 1, 1);
    run<5, 1>("IMAD.HI + LOP3", 1, 1);
    run<0, 4>("IADD3 + IMAD", 1, 1);
    return 0;
}
