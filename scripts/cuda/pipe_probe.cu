// pipe_probe.cu -- how much multiplier throughput does a B200 SM have besides the IMAD ("fmaheavy") pipe that the 256-bit
// Montgomery product lives on?  Three kernels over the whole GPU, timed with CUDA events:
//   imad : dependent IMAD.WIDE chains (what ff.cuh issues: 32 x 32 -> 64 bit, 4 cycles per warp instruction and sub-partition)
//   dfma : dependent DFMA chains (the FP64 pipe: a 52-bit-limb Montgomery product would run there)
//   both : the two interleaved in one warp (do the pipes overlap?)
// Independent of the library: no field code, no uzkge header -- bench.py runs it (`pipe_probe --json`) for the PEAK of its integer
// roofline, so that the peak does not come from the product's own multiplier loop.
// Build (also done by __graft_entry__.build()):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_probe.cu -o pipe_probe
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(unsigned long long* sink, double* dsink, int iters) {
    unsigned lo[8], hi[8], xa[8], xb[8];
    double d[8];
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        lo[i] = t * 8 + i + 1;
        hi[i] = i;
        xa[i] = t + i;
        xb[i] = t ^ i;
        d[i] = 1.0 + 1e-9 * (t + i);
    }
    const unsigned m = 0x9E3779B1u + t;
    const double dm = 1.0000001;
#pragma unroll 1
    for (int it = 0; it < iters; it += 16) {
        // inline PTX so that the loop body is 128 multiplier instructions (8 independent accumulator chains x 16): left to
        // itself nvcc re-materialises the 64-bit addend with MOV / IMAD.MOV pairs, which share the pipe and halve the figure
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                // what ff.cuh issues: a mad.lo.cc / madc.hi pair on one 64-bit accumulator, which ptxas fuses into ONE
                // IMAD.WIDE.U32 with the addend inside the instruction (a plain mad.wide.u32 is split into IMAD.WIDE + IADD3 +
                // IADD3.X, and every extra instruction takes an issue slot away from the multiplier).  The multiplicand is the
                // low word of the NEIGHBOURING chain: nothing is loop-invariant for ptxas to hoist
                if (MODE != 1 && MODE != 4)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                                 : "+r"(lo[i]), "+r"(hi[i]) : "r"(lo[(i + 1) & 7]), "r"(m));
                if (MODE == 1 || MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dm), "d"(1e-12));   // DFMA
                // MODE 3: one add-with-carry pair (IADD3 + IADD3.X, the ALU pipe) beside every multiplier instruction; 4: those alone
                if (MODE >= 3) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %2;" : "+r"(xa[i]), "+r"(xb[i]) : "r"(m));
            }
        }
    }
    unsigned long long s = 0;
    double ds = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s += (((unsigned long long)hi[i] << 32) | lo[i]) + xa[i] + xb[i];
        ds += d[i];
    }
    if (s == 0x1234567ull) sink[t] = s;
    if (ds == 1.2345) dsink[t] = ds;
}

template <int MODE>
static double run(int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned long long* sink;
    double* dsink;
    const int blocks = sms * 8, threads = 256;
    cudaMalloc(&sink, sizeof(unsigned long long) * blocks * threads);
    cudaMalloc(&dsink, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<blocks, threads>>>(sink, dsink, iters);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(sink, dsink, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(sink);
    cudaFree(dsink);
    return (double)blocks * threads * 8.0 * iters / (ms * 1e-3);   // operations of EACH kind per second
}

int main(int argc, char** argv) {
    const int iters = 20000;
    if (argc > 1 && std::string(argv[1]) == "--json") {
        if (argc > 2) cudaSetDevice(atoi(argv[2]));
        int dev = 0, sms = 0, khz = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        // a fresh process finds the GPU at idle clocks: keep the multiplier pipe busy for ~0.4 s first, then best of 5
        double best = 0;
        for (int rep = 0; rep < 40; rep++) run<0>(iters);
        for (int rep = 0; rep < 5; rep++) {
            const double v = run<0>(iters);
            if (v > best) best = v;
        }
        // issue model: IMAD.WIDE runs on the fmaheavy half of the FMA pipe, 8 lanes per SM sub-partition and clock
        printf("{\"imad_wide_per_s\": %.6e, \"sms\": %d, \"max_clock_khz\": %d, \"model_per_s\": %.6e, "
               "\"how\": \"8 dependent IMAD.WIDE chains per thread, 8 x 256 threads per SM, CUDA events, 0.4 s of warm-up, best of 5\"}\n",
               best, sms, khz, (double)sms * 4 * 8 * khz * 1e3);
        return 0;
    }
    const double imad = run<0>(iters), dfma = run<1>(iters), both = run<2>(iters);
    printf("IMAD.WIDE alone : %8.2f T/s\n", imad / 1e12);
    printf("DFMA alone      : %8.2f T/s\n", dfma / 1e12);
    printf("interleaved     : %8.2f T/s of each (IMAD.WIDE + DFMA issued by the same warps)\n", both / 1e12);
    const double mix = run<3>(iters), adds = run<4>(iters);
    printf("IMAD.WIDE + 2 IADD3 per multiplier instruction: %8.2f T IMAD.WIDE/s;  the IADD3 pairs alone: %8.2f T pairs/s\n", mix / 1e12, adds / 1e12);
    return 0;
}
