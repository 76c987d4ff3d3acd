// pipe_mix.cu -- which integer instruction mixes issue above 64 thread-ops/clk/SM on this GPU?
// tools/alu_peak.cu measured LOP3, SHF and IMAD at 64 lanes/clk/SM each but its 3:1 mix used ONE dependent chain
// set for both instruction kinds; here every kind has its own independent chains, so a mix is limited only by the
// pipes and the issue slots.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_mix tools/pipe_mix.cu && tools/pipe_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define UNROLL 8

// NA chains of kind A and NB chains of kind B per thread, interleaved
template <int KA, int KB, int NA, int NB>
__global__ void __launch_bounds__(256) mix_kernel(uint32_t* out, uint32_t a, uint32_t b, int iters)
{
    uint32_t x[NA > 0 ? NA : 1], y[NB > 0 ? NB : 1];
#pragma unroll
    for (int j = 0; j < NA; j++) x[j] = threadIdx.x * 2654435761u + j * 40503u + a;
#pragma unroll
    for (int j = 0; j < NB; j++) y[j] = threadIdx.x * 40503u + j * 2654435761u + b;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int j = 0; j < (NA > NB ? NA : NB); j++) {
                if (j < NA) {
                    if (KA == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[(j + 1) % NA]), "r"(a));
                    if (KA == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(x[(j + 1) % NA]));
                    if (KA == 2) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x[j]) : "r"(x[(j + 1) % NA]));
                    if (KA == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[(j + 1) % NA]));
                }
                if (j < NB) {
                    if (KB == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[j]) : "r"(a), "r"(y[(j + 1) % NB]));
                    if (KB == 1) asm volatile("mad.lo.u32 %0, %0, 16, %1;" : "+r"(y[j]) : "r"(y[(j + 1) % NB]));      // shift-left-add as IMAD
                    if (KB == 2) asm volatile("popc.b32 %0, %0;" : "+r"(y[j]));                                      // XU/other pipe
                    if (KB == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&y[j]) : "f"(1.0001f), "f"(*(float*)&y[(j + 1) % NB]));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < NA; j++) s ^= x[j];
#pragma unroll
    for (int j = 0; j < NB; j++) s ^= y[j];
    if (s == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KA, int KB, int NA, int NB>
static void run(const char* name, int sms, double ghz)
{
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<KA, KB, NA, NB><<<sms * 8, 256>>>(out, 0x9e3779b9u, 7, 100);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        mix_kernel<KA, KB, NA, NB><<<sms * 8, 256>>>(out, 0x9e3779b9u, 7, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)sms * 8 * 256 * iters * UNROLL * (NA + NB);
    const double tops = ops / (best * 1e-3) / 1e12;
    printf("{\"mix\": \"%s\", \"ms\": %.3f, \"tera_thread_ops_per_s\": %.2f, \"thread_ops_per_clk_per_sm_at_%.3f_ghz\": %.1f}\n",
           name, best, tops, ghz, tops * 1e12 / (sms * ghz * 1e9));
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    const int s = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_ghz\": %.3f}\n", p.name, s, ghz);
    run<0, 0, 8, 0>("LOP3 only (8 chains)", s, ghz);
    run<0, 0, 0, 8>("IMAD only (8 chains)", s, ghz);
    run<0, 0, 4, 4>("LOP3 : IMAD 1:1", s, ghz);
    run<0, 0, 6, 2>("LOP3 : IMAD 3:1", s, ghz);
    run<0, 1, 4, 4>("LOP3 : IMAD(imm shift-add) 1:1", s, ghz);
    run<1, 0, 4, 4>("SHF : IMAD 1:1", s, ghz);
    run<2, 0, 8, 0>("PRMT only", s, ghz);
    run<2, 0, 4, 4>("PRMT : IMAD 1:1", s, ghz);
    run<3, 0, 8, 0>("IADD only", s, ghz);
    run<0, 2, 6, 2>("LOP3 : POPC 3:1", s, ghz);
    run<0, 3, 4, 4>("LOP3 : FFMA 1:1", s, ghz);
    run<0, 3, 0, 8>("FFMA only", s, ghz);
    return 0;
}
