// alu_peak.cu -- measures the integer pipe rates that bound cand31_kernel on this GPU:
// LOP3 and SHF (ALU pipe), IMAD (FMA pipe), and a 3:1 LOP3+IMAD mix (both pipes).  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/alu_peak tools/alu_peak.cu && tools/alu_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters)
{
    uint32_t x[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; j++) x[j] = threadIdx.x * 2654435761u + j * 40503u + a;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int j = 0; j < CHAINS; j++) {
                if (MODE == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]), "r"(a));
                if (MODE == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]), "r"(b));
                if (MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(x[(j + 1) % CHAINS]));
                if (MODE == 4) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]));                    // two register sources
                if (MODE == 5) asm volatile("lop3.b32 %0, %0, %1, 0x5bd1e995, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]));  // two registers + immediate
                if (MODE == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]));           // immediate shift
                if (MODE == 7) asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]));               // immediate multiplier
                if (MODE == 3) {
                    if (j % 3 == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(x[(j + 1) % CHAINS]));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % CHAINS]), "r"(a));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; j++) s ^= x[j];
    if (s == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // keeps the chains alive
}

template <int MODE>
static void run(const char* name, int sms, double clock_ghz)
{
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<sms * 8, 256>>>(out, 0x9e3779b9u, 7, 100);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        rate_kernel<MODE><<<sms * 8, 256>>>(out, 0x9e3779b9u, 7, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)sms * 8 * 256 * iters * UNROLL * CHAINS;
    const double tops = ops / (best * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"tera_thread_ops_per_s\": %.2f, \"lanes_per_clk_per_sm_at_%.3f_ghz\": %.1f}\n",
           name, best, tops, clock_ghz, tops * 1e12 / (sms * clock_ghz * 1e9));
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    run<0>("LOP3 (ALU pipe)", p.multiProcessorCount, ghz);
    run<1>("SHF funnel (ALU pipe)", p.multiProcessorCount, ghz);
    run<2>("IMAD (FMA pipe)", p.multiProcessorCount, ghz);
    run<3>("3 LOP3 : 1 IMAD (both pipes)", p.multiProcessorCount, ghz);
    // operand-bandwidth question left open by the four lines above (all three-register forms): do forms with two
    // register sources issue faster?  (Not yet run on a GPU.  ptxas fuses pairs of the two-source XORs into LOP3s --
    // check the instruction count with cuobjdump before trusting that line.)
    run<4>("XOR, two register sources", p.multiProcessorCount, ghz);
    run<5>("LOP3, two registers + immediate", p.multiProcessorCount, ghz);
    run<6>("SHF funnel, immediate shift", p.multiProcessorCount, ghz);
    run<7>("IMAD, immediate multiplier", p.multiProcessorCount, ghz);
    return 0;
}
