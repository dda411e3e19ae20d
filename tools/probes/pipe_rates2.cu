// Probe (round 2): issue / pipe cost of the instructions of the hidden-layer epilogue, measured with LOOP-CARRIED dependencies so
// that ptxas cannot hoist or fold them (the first probe's conversion and RCP numbers were folded away).  One CTA of 16 warps (4 per
// scheduler) or 4 warps (1 per scheduler); every thread runs 8 independent chains; prints cycles per warp-instruction per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/pipe_rates2 tools/probes/pipe_rates2.cu && /tmp/pipe_rates2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 8;

template <int OP>
__device__ __forceinline__ void body(float (&x)[CH], uint32_t (&h)[CH]) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(0.999f), "f"(0.001f));
        if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 3) {   // f32 pair -> f16x2 (saturating); the result feeds the next conversion as one of the two inputs
            asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[j]), "f"(__uint_as_float(h[j])));
        }
        if (OP == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[j]), "f"(__uint_as_float(h[j])));
        if (OP == 5) {   // f16 -> f32 of the low half
            asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %1;\n\tcvt.f32.f16 %0, a;\n\t}" : "=f"(x[j]) : "r"(__float_as_uint(x[j])));
        }
        if (OP == 6) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 7) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x[j]), "f"(__uint_as_float(h[j])));
        if (OP == 9) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(0.5f));
        if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[j]) : "r"(0x12345u), "r"(h[(j + 1) % CH]));
        if (OP == 11) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(__uint_as_float(h[j])));
        if (OP == 12) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[j]));
        if (OP == 13) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[j]) : "r"(0x3c003c00u), "r"(h[(j + 1) % CH]));
    }
    if (OP == 20 || OP == 21 || OP == 22 || OP == 23) {
        // the epilogue's arithmetic on 8 values: u = a * s + b; t = 2^u; d = (1 + t) / 8; r = 1 / d; y = u * r; (hi, lo) = split(y)
        //   20: as the kernel does it (EX2 + RCP on the MUFU, F2FP split)      21: without the split
        //   22: RCP by Newton iterations on the FMA pipe (magic-constant seed, 3 steps), F2FP split     23: 22 without the split
        float u[CH], y[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) u[j] = fmaf(x[j], 0.7f, 0.1f);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            float t, d, r;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(u[j]));
            d = fmaf(t, 0.125f, 0.125f);
            if (OP == 20 || OP == 21) {
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
            } else {
                d = fminf(d, 1e30f);
                r = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
                float e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
                e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
                e = fmaf(-d, r, 1.0f); r = fmaf(r, e, r);
            }
            y[j] = u[j] * r;
        }
        if (OP == 20 || OP == 22) {
#pragma unroll
            for (int j = 0; j < CH; j += 2) {
                uint32_t hh, ll;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(y[j + 1]), "f"(y[j]));
                float h0, h1;
                asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hh));
                const float r0 = y[j] - h0, r1 = y[j + 1] - h1;
                asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ll) : "f"(r1), "f"(r0));
                h[j] ^= hh; h[j + 1] ^= ll;
                x[j] = y[j] + __uint_as_float(ll & 1u); x[j + 1] = y[j + 1];
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) x[j] = y[j];
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int warps_active, int iters) {
    float x[CH];
    uint32_t h[CH];
    for (int j = 0; j < CH; ++j) { x[j] = 0.3f + 0.001f * (threadIdx.x + j); h[j] = 0x3c003c00u + j + threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
    if ((int)(threadIdx.x >> 5) < warps_active)
        for (int i = 0; i < iters; ++i) body<OP>(x, h);
    const long long t1 = clock64();
    __syncthreads();
    float acc = 0.f;
    for (int j = 0; j < CH; ++j) acc += x[j] + __uint_as_float(h[j]);
    if (acc == 123.456f) sink[0] = acc;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, long long* d, float* sink) {
    for (int warps : {4, 16}) {
        const int iters = 256;
        k<OP><<<1, 512>>>(d, sink, warps, iters);
        cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%-58s warps=%2d  cycles per warp-instruction (or per 8 activations) per scheduler = %7.2f\n", name, warps,
               (double)c / ((double)iters * per_iter * (warps / 4)));
    }
}

int main() {
    long long* d;
    float* sink;
    cudaMalloc(&d, 64);
    cudaMalloc(&sink, 64);
    run<0>("FFMA", CH, d, sink);
    run<9>("FADD", CH, d, sink);
    run<10>("LOP3", CH, d, sink);
    run<11>("FMNMX", CH, d, sink);
    run<1>("MUFU.EX2", CH, d, sink);
    run<2>("MUFU.RCP", CH, d, sink);
    run<6>("MUFU.SQRT (sqrt.approx)", CH, d, sink);
    run<7>("MUFU.TANH", CH, d, sink);
    run<12>("MUFU.EX2.F16x2 (ex2.approx.f16x2)", CH, d, sink);
    run<3>("F2FP.SATFINITE.F16.F32.PACK_AB (cvt.rn.satfinite.f16x2.f32)", CH, d, sink);
    run<4>("F2FP.F16.F32.PACK_AB (cvt.rn.f16x2.f32)", CH, d, sink);
    run<8>("F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32)", CH, d, sink);
    run<5>("HADD2.F32 (cvt.f32.f16)", CH, d, sink);
    run<13>("HFMA2 (fma.rn.f16x2)", CH, d, sink);
    run<20>("8 activations: EX2 + RCP + F2FP split (kernel)", 1, d, sink);
    run<21>("8 activations: EX2 + RCP, no split", 1, d, sink);
    run<22>("8 activations: EX2 + Newton RCP + F2FP split", 1, d, sink);
    run<23>("8 activations: EX2 + Newton RCP, no split", 1, d, sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
