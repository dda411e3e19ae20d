// Probe: issue rate of the instructions the tensor-core epilogues are made of, per SM sub-partition (one CTA of 512 threads =
// 4 warps per scheduler, 8 independent chains per thread).  Prints cycles per warp-instruction per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates tools/probes/pipe_rates.cu && /tmp/pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
#define CHAINS 8

template <int OP>
__device__ __forceinline__ void body(float (&x)[CHAINS], uint32_t (&h)[CHAINS], unsigned long long (&p)[CHAINS / 2]) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
        if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(0.999f), "f"(0.001f));
        if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 3) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[j]), "f"(x[(j + 1) % CHAINS]));
        if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[j]), "f"(x[(j + 1) % CHAINS]));
        if (OP == 5) asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %1;\n\tcvt.f32.f16 %0, a;\n\t}" : "=f"(x[j]) : "r"(h[j]));
        if (OP == 6) asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %1;\n\tneg.f16 a, a;\n\tadd.f32.f16 %0, a, %0;\n\t}" : "+f"(x[j]) : "r"(h[j]));
        if (OP == 10) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    }
#pragma unroll
    for (int j = 0; j < CHAINS / 2; ++j) {
        if (OP == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]), "l"(p[(j + 2) % (CHAINS / 2)]));
        if (OP == 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]));
        if (OP == 9) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]));
    }
    if (OP == 11) {   // the epilogue's mix per 8 activations: 4 FADD2, 4 FFMA2, 8 EX2, 4 FFMA2, 8 RCP, 4 FMUL2
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]));
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]), "l"(p[(j + 2) % (CHAINS / 2)]));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]), "l"(p[(j + 2) % (CHAINS / 2)]));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(p[(j + 1) % (CHAINS / 2)]));
    }
    if (OP == 12) {   // the scalar mix of round 1: 8 FADD, 8 FFMA, 8 FMUL, 8 EX2, 8 FADD, 8 RCP, 8 FMUL
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(0.5f));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(0.999f), "f"(0.001f));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(0.5f));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(1.0f));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(0.75f));
    }
    if (OP == 13) {   // the hi / lo split of 8 values: 4 F2FP, 8 cvt.f32.f16, 8 FADD, 4 F2FP
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[2 * j + 1]), "f"(x[2 * j]));
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) {
            float a, b;
            asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(a), "=f"(b) : "r"(h[j]));
            asm volatile("sub.rn.f32 %0, %0, %1;" : "+f"(x[2 * j]) : "f"(a));
            asm volatile("sub.rn.f32 %0, %0, %1;" : "+f"(x[2 * j + 1]) : "f"(b));
        }
#pragma unroll
        for (int j = 0; j < CHAINS / 2; ++j) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[4 + j]) : "f"(x[2 * j + 1]), "f"(x[2 * j]));
    }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int warps_active) {
    float x[CHAINS];
    uint32_t h[CHAINS];
    unsigned long long p[CHAINS / 2];
    for (int j = 0; j < CHAINS; ++j) { x[j] = 1.0f + 0.001f * (threadIdx.x + j); h[j] = 0x3c003c00u + j; }
    for (int j = 0; j < CHAINS / 2; ++j) p[j] = 0x3f8000003f800000ull + j;
    __syncthreads();
    const long long t0 = clock64();
    if ((int)(threadIdx.x >> 5) < warps_active) {
        for (int r = 0; r < REP; ++r) body<OP>(x, h, p);
    }
    const long long t1 = clock64();
    float acc = 0.f;
    for (int j = 0; j < CHAINS; ++j) acc += x[j] + __uint_as_float(h[j]);
    for (int j = 0; j < CHAINS / 2; ++j) acc += (float)p[j];
    if (acc == 12345.678f) sink[threadIdx.x] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_body, long long* d, float* sink) {
    for (int wa : {4, 16}) {
        k<OP><<<1, 512>>>(d, sink, wa);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, sizeof(c), cudaMemcpyDeviceToHost);
        const double per_sched = (double)wa / 4.0 * REP * instr_per_body;      // warp-instructions per scheduler
        printf("%-44s warps=%2d cycles=%8lld  cycles per warp-instruction per scheduler = %6.2f\n", name, wa, c, c / per_sched);
    }
}

int main() {
    long long* d;
    float* sink;
    cudaMalloc(&d, 64);
    cudaMalloc(&sink, 4096);
    run<0>("FFMA", 8, d, sink);
    run<7>("FFMA2", 4, d, sink);
    run<8>("FADD2", 4, d, sink);
    run<9>("FMUL2", 4, d, sink);
    run<1>("MUFU.EX2", 8, d, sink);
    run<2>("MUFU.RCP", 8, d, sink);
    run<10>("sqrt.approx (MUFU.RSQ/SQRT)", 8, d, sink);
    run<3>("F2FP.SATFINITE.F16.F32.PACK_AB", 8, d, sink);
    run<4>("F2FP.F16.F32.PACK_AB", 8, d, sink);
    run<5>("HADD2.F32 (cvt.f32.f16)", 8, d, sink);
    run<6>("FHADD (add.f32.f16, negated)", 8, d, sink);
    run<11>("epilogue mix, packed (32 instr / 8 act)", 32, d, sink);
    run<12>("epilogue mix, scalar (56 instr / 8 act)", 56, d, sink);
    run<13>("hi/lo split of 8 values (24 instr)", 24, d, sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
