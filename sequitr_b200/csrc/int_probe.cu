// Issue-rate probe for the integer instructions the weight-map column passes are made of (sm_100a):
// VIADDMNMX (DPX fused add-min, 16x2 and 32-bit), plain IADD3, VIMNMX (two-input min) and the select.
// Each kernel runs 8 independent dependency chains per thread of ONE instruction kind; the figure printed is
// warp-instructions per clock per SM.  Build: make ../_build/int_probe.  Used for DESIGN.md section 4.
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) probe(unsigned *out, unsigned seed, int iters)
{
    unsigned acc[8], g = seed ^ threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = seed * (j + 3) + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (MODE == 0) acc[j] = __viaddmin_u16x2(acc[j], 0x00010001u * (u + 1), g);
                if (MODE == 1) acc[j] = __viaddmin_u32(acc[j], 7u * (u + 1), g);
                if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[j]) : "r"(g));   // IADD (kept from folding)
                if (MODE == 3) acc[j] = min(acc[j] ^ (unsigned)(u + 1), g);        // LOP3 + VIMNMX (2 instructions)
                if (MODE == 4) acc[j] = (acc[j] != g) ? acc[j] + 1u : g;           // ISETP + SEL-ish
                if (MODE == 5) acc[j] = acc[j] ^ (g + u);                          // LOP3 only (reference rate)
            }
        g = g * 1664525u + 1013904223u;
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    unsigned *out;
    const int blocks = sms * 8, iters = 2048;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    probe<MODE><<<blocks, 256>>>(out, 1u, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, 7u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double steps = (double)blocks * 8 * iters * 64.0;                    // chain steps per warp, summed over warps
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-44s %8.3f ms  %6.2f chain steps / clk / SM (nominal %d MHz)\n", name, ms, steps / clk / sms, khz / 1000);
    cudaFree(out);
}

int main()
{
    run<5>("LOP3 (reference: full-rate ALU op)");
    run<2>("IADD3");
    run<0>("VIADDMNMX.U16x2 (DPX add-min, two columns)");
    run<1>("VIADDMNMX.U32   (DPX add-min)");
    run<3>("LOP3 + VIMNMX.U32 (min, 2 instructions)");
    run<4>("ISETP + SEL/IADD  (compare-select step)");
    return 0;
}
