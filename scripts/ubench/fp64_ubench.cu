// fp64_ubench.cu -- micro-benchmarks of the B200 FP64 pipe that the pair kernels live on.
// Questions: (1) issue rate of DFMA by number of DISTINCT 64-bit register operands (RF-bank read limit),
// (2) does an operand shared by consecutive DFMAs (.reuse) lift that limit, (3) DFMA dependent latency,
// (4) MUFU.RSQ64H rate, (5) does DMMA (FP64 tensor) overlap with the vector FP64 pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <mma.h>

#define ITERS 4096
#define NACC 8

__global__ void k_3reg(double *out, const double *in) { // x_i = fma(a_i, b_i, x_i): 3 distinct registers
    double x[NACC], a[NACC], b[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; a[i] = in[8 + i + threadIdx.x % 2]; b[i] = in[20 + i + threadIdx.x % 3]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(a[i], b[i], x[i]);
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_3reg_shared(double *out, const double *in) { // x_i = fma(a, b_i, x_i): operand a shared by neighbours
    double x[NACC], b[NACC]; double a = in[40 + threadIdx.x % 2];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; b[i] = in[20 + i + threadIdx.x % 3]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(a, b[i], x[i]);
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_2reg(double *out, const double *in) { // x_i = fma(a_i, a_i, x_i): 2 distinct registers
    double x[NACC], a[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; a[i] = in[8 + i + threadIdx.x % 2]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(a[i], a[i], x[i]);
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_2reg_imm(double *out, const double *in) { // x_i = fma(x_i, a_i, 0.5): 2 registers + immediate
    double x[NACC], a[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; a[i] = in[8 + i + threadIdx.x % 2]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(x[i], a[i], 0.5);
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_1reg_const(double *out, const double *in, double ca, double cb) { // x = fma(x, c[..], c[..])
    double x[NACC];
    for (int i = 0; i < NACC; ++i) x[i] = in[i];
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = fma(x[i], ca, cb);
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dadd2(double *out, const double *in) { // x_i = x_i + a_i
    double x[NACC], a[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; a[i] = in[8 + i + threadIdx.x % 2]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) x[i] = x[i] + a[i];
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_latency(double *out, const double *in, double ca, double cb) { // one dependent chain
    double x = in[threadIdx.x % 4];
    for (int it = 0; it < ITERS * NACC; ++it) x = fma(x, ca, cb);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_mufu(double *out, const double *in) { // MUFU.RSQ64H throughput
    double x[NACC];
    for (int i = 0; i < NACC; ++i) x[i] = in[i] + 1.0;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = y; }
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per 3-reg DFMA one FFMA (does FP32 work ride along for free?)
__global__ void k_3reg_plus_ffma(double *out, const double *in) {
    double x[NACC], a[NACC], b[NACC]; float f[NACC];
    for (int i = 0; i < NACC; ++i) { x[i] = in[i]; a[i] = in[8 + i + threadIdx.x % 2]; b[i] = in[20 + i + threadIdx.x % 3]; f[i] = (float)in[i]; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) { x[i] = fma(a[i], b[i], x[i]); f[i] = fmaf(f[i], 0.999f, 0.001f); }
    double s = 0; for (int i = 0; i < NACC; ++i) s += x[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA m8n8k4 alone and mixed with DFMA
__global__ void k_dmma(double *out, const double *in, int with_dfma) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = in[threadIdx.x % 8], b = in[8 + threadIdx.x % 8];
    double x[4] = {in[0], in[1], in[2], in[3]}, p = in[5], q = in[6];
    for (int it = 0; it < ITERS; ++it) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
        if (with_dfma) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = fma(x[i], p, q);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1] + x[0] + x[1] + x[2] + x[3];
}
__global__ void k_dfma16(double *out, const double *in) { // same DFMA load as the mixed kernel, without DMMA
    double x[4] = {in[0], in[1], in[2], in[3]}, p = in[5], q = in[6];
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = fma(x[i], p, q);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1e-3 * (i + 1); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("device %s, %d SMs, max clock %.0f MHz\n", p.name, sms, clk_khz / 1e3);
    const int blocks = sms * 2, threads = 512; // 32 warps / SM = 8 / SMSP
    const double winst = (double)blocks * (threads / 32) * ITERS * NACC; // warp-instructions of the measured op
    auto report = [&](const char *name, float ms, double wi) {
        double per_smsp_per_s = wi / (sms * 4) / (ms * 1e-3);
        printf("%-28s %8.3f ms  -> %.3f cycles/warp-inst/SMSP at %.0f MHz (%.2f Tinst-lanes/s)\n", name, ms,
               (clk_khz * 1e3) / per_smsp_per_s, clk_khz / 1e3, wi * 32 / (ms * 1e-3) / 1e12);
    };
    report("DFMA 3 distinct regs", timeit([&] { k_3reg<<<blocks, threads>>>(out, in); }), winst);
    report("DFMA 3 regs, 1 shared", timeit([&] { k_3reg_shared<<<blocks, threads>>>(out, in); }), winst);
    report("DFMA 2 distinct regs", timeit([&] { k_2reg<<<blocks, threads>>>(out, in); }), winst);
    report("DFMA 2 regs + imm", timeit([&] { k_2reg_imm<<<blocks, threads>>>(out, in); }), winst);
    report("DFMA 1 reg + 2 const", timeit([&] { k_1reg_const<<<blocks, threads>>>(out, in, 0.999, 1e-3); }), winst);
    report("DADD 2 regs", timeit([&] { k_dadd2<<<blocks, threads>>>(out, in); }), winst);
    report("MUFU.RSQ64H", timeit([&] { k_mufu<<<blocks, threads>>>(out, in); }), winst);
    report("DFMA 3 regs + FFMA each", timeit([&] { k_3reg_plus_ffma<<<blocks, threads>>>(out, in); }), winst);
    { // latency: 1 warp per SMSP
        float ms = timeit([&] { k_latency<<<sms, 128>>>(out, in, 0.999, 1e-3); });
        printf("%-28s %8.3f ms  -> %.2f cycles dependent-issue latency\n", "DFMA dependent chain", ms,
               ms * 1e-3 * clk_khz * 1e3 / ((double)ITERS * NACC));
    }
    {
        double wi_mma = (double)blocks * (threads / 32) * ITERS * 4;
        double wi_fma = (double)blocks * (threads / 32) * ITERS * 16;
        float t_mma = timeit([&] { k_dmma<<<blocks, threads>>>(out, in, 0); });
        float t_fma = timeit([&] { k_dfma16<<<blocks, threads>>>(out, in); });
        float t_mix = timeit([&] { k_dmma<<<blocks, threads>>>(out, in, 1); });
        report("DMMA m8n8k4 alone", t_mma, wi_mma);
        report("16 DFMA alone", t_fma, wi_fma);
        printf("DMMA(4)+DFMA(16) mixed       %8.3f ms  (sum of parts %.3f ms; max of parts %.3f ms)\n", t_mix, t_mma + t_fma,
               t_mma > t_fma ? t_mma : t_fma);
        printf("DMMA: %.2f TFLOP/s (2*8*8*4 flop per warp-inst)\n", wi_mma * 512 / (t_mma * 1e-3) / 1e12);
    }
    return 0;
}
