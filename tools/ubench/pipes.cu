// Pipe-rate microbenchmark (developer tool): cycles per warp-instruction per SM sub-partition for the instructions the fused
// softmax kernels are built from.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  unsigned h0 = __float_as_uint(x0), h1 = __float_as_uint(x1), h2 = __float_as_uint(x2), h3 = __float_as_uint(x3);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {  // MUFU.EX2 f32
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x7));
      } else if (MODE == 1) {  // ex2.f16x2 (two MUFU.EX2.F16 per instruction)
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
      } else if (MODE == 2) {  // FFMA, three register operands
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x0) : "f"(a), "f"(b)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x1) : "f"(a), "f"(b));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x2) : "f"(a), "f"(b)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x3) : "f"(a), "f"(b));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x4) : "f"(a), "f"(b)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x5) : "f"(a), "f"(b));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x6) : "f"(a), "f"(b)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x7) : "f"(a), "f"(b));
      } else if (MODE == 3) {  // FFMA with an immediate addend
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x0) : "f"(a)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x1) : "f"(a));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x2) : "f"(a)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x3) : "f"(a));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x4) : "f"(a)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x5) : "f"(a));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x6) : "f"(a)); asm volatile("fma.rn.ftz.f32 %0, %0, %1, 0f3F000000;" : "+f"(x7) : "f"(a));
      } else if (MODE == 4) {  // FADD
        asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x0) : "f"(a)); asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x1) : "f"(a));
        asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x2) : "f"(a)); asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x3) : "f"(a));
        asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x4) : "f"(a)); asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x5) : "f"(a));
        asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x6) : "f"(a)); asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x7) : "f"(a));
      } else if (MODE == 5) {  // 3-input max (ALU pipe)
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x0) : "f"(a), "f"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x1) : "f"(a), "f"(b));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x2) : "f"(a), "f"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x3) : "f"(a), "f"(b));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x4) : "f"(a), "f"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x5) : "f"(a), "f"(b));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x6) : "f"(a), "f"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x7) : "f"(a), "f"(b));
      } else if (MODE == 6) {  // HFMA2 (packed half FMA)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h0) : "r"(h1), "r"(h2)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h3) : "r"(h1), "r"(h2));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h0) : "r"(h1), "r"(h2)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h3) : "r"(h1), "r"(h2));
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int warps_per_smsp) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  int threads = 32 * 4 * warps_per_smsp;
  k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.5f);
  k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double instr = 256.0 * 8 * per_iter * warps_per_smsp;    // warp-instructions per sub-partition
  printf("%-34s warps/SMSP %d : %.2f cycles per warp-instruction per sub-partition\n", name, warps_per_smsp, c / instr);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("MUFU.EX2 f32", 8, w);
    run<1>("ex2.f16x2 (2 elements / instr)", 4, w);
    run<2>("FFMA reg,reg,reg", 8, w);
    run<3>("FFMA reg,reg,imm", 8, w);
    run<4>("FADD", 8, w);
    run<5>("FMNMX3", 8, w);
    run<6>("HFMA2", 4, w);
  }
  return 0;
}
