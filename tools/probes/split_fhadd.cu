// Probe (compile only): the fp32 -> (fp16 hi, fp16 lo) split with the mixed-precision add of PTX ISA 8.6 (sm_100+).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin -o /tmp/p.cubin tools/probes/split_fhadd.cu && cuobjdump -sass /tmp/p.cubin
// k_neg:  neg.f16 + add.f32.f16 fold into ONE `FHADD R, -R.H0, R` per value: the split of a pair is 2 F2FP + 2 FHADD = 4
//         instructions instead of the 6 of csrc/ptx.cuh split2 (F2FP, 2 HADD2.F32, 2 FADD, F2FP); same value bit for bit
//         (x - float(h) in fp32 with one rounding either way).  One instruction per activation less in every epilogue of
//         the tensor-core kernels.  Not yet in the kernels: found after this round's GPU
//         budget was spent, and a kernel change does not ship unmeasured.
// k_fma:  fma.rn.f32.f16 with a -1 constant: also one FHFMA per value, plus a constant register.
#include <cstdint>
__global__ void k_neg(const float* x, uint32_t* out) {
    float x0 = x[threadIdx.x * 2], x1 = x[threadIdx.x * 2 + 1];
    uint32_t h, lo;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    float r0, r1;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tneg.f16 a, a;\n\tneg.f16 b, b;\n\tadd.f32.f16 %0, a, %3;\n\tadd.f32.f16 %1, b, %4;\n\t}" : "=f"(r0), "=f"(r1) : "r"(h), "f"(x0), "f"(x1));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
    out[threadIdx.x * 2] = h; out[threadIdx.x * 2 + 1] = lo;
}
__global__ void k_fma(const float* x, uint32_t* out) {
    float x0 = x[threadIdx.x * 2], x1 = x[threadIdx.x * 2 + 1];
    uint32_t h, lo;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    float r0, r1;
    asm("{\n\t.reg .b16 a, b, m;\n\tmov.b32 {a, b}, %2;\n\tmov.b16 m, 0xBC00;\n\tfma.rn.f32.f16 %0, a, m, %3;\n\tfma.rn.f32.f16 %1, b, m, %4;\n\t}" : "=f"(r0), "=f"(r1) : "r"(h), "f"(x0), "f"(x1));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
    out[threadIdx.x * 2] = h; out[threadIdx.x * 2 + 1] = lo;
}
