// Issue-rate microbenchmark for the sm_100a pipes the per-frame kernels lean on (tuning aid, not product code).
// Each kernel runs ILP independent dependency chains of one instruction kind per thread; reported: warp-instructions
// per clock per SM at full occupancy (148 x 8 CTAs x 256 threads).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int KIND> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int n) {
  float x[ILP], y[ILP];
  unsigned u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = x[i] * 0.5f; u[i] = threadIdx.x + i; }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b)); }
      if (KIND == 1) { // f32x2: two FMAs per instruction
        unsigned long long v, A, B;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(y[i]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(A), "l"(B));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(y[i]) : "l"(v));
      }
      if (KIND == 2) { asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP])); asm volatile("xor.b32 %0, %0, 0x5bd1e995;" : "+r"(u[i])); }
      if (KIND == 3) { // FFMA + LOP3 interleaved: do the two pipes dual-issue?
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        asm volatile("xor.b32 %0, %0, 0x5bd1e995;" : "+r"(u[i]));
      }
      if (KIND == 4) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a)); }
      if (KIND == 5) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a)); }
      if (KIND == 6) { asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a)); }
      if (KIND == 7) { int t; asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(t) : "f"(x[i])); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(x[i]) : "r"(t)); }
      if (KIND == 8) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]), "r"(it)); }
      if (KIND == 9) { // half -> float unpack
        unsigned short h = (unsigned short)u[i];
        asm volatile("cvt.f32.f16 %0, %1;" : "=f"(x[i]) : "h"(h));
        u[i] = __float_as_uint(x[i]);
      }
      if (KIND == 10) { asm volatile("fma.rn.f32 %0, %0, %1, 0f3f000000;" : "+f"(x[i]) : "f"(a)); } // immediate addend
      if (KIND == 11) { // FFMA + FADD interleaved (same pipe?)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(a));
      }
      if (KIND == 12) { // FFMA + FMNMX interleaved
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(a));
      }
      if (KIND == 13) { // FFMA2 + LOP3 interleaved
        unsigned long long v, A, B;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(y[i]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(A), "l"(B));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(y[i]) : "l"(v));
        asm volatile("xor.b32 %0, %0, 0x5bd1e995;" : "+r"(u[i]));
      }
      if (KIND == 14) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND> static void run(const char* name, int instrPerStep, float* out, int sms, double clkGHz) {
  const int ctas = sms * 8;
  k<KIND><<<ctas, 256>>>(out, 1.0001f, 0.5f, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<KIND><<<ctas, 256>>>(out, 1.0001f, 0.5f, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warpInstr = (double)ctas * 8.0 * ITERS * ILP * instrPerStep;
  const double perClkPerSm = warpInstr / (ms * 1e-3 * clkGHz * 1e9) / sms;
  printf("{\"kind\": \"%s\", \"ms\": %.3f, \"warp_instr_per_clk_per_sm\": %.3f}\n", name, ms, perClkPerSm);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ghz = clk * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  float* out;
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(float));
  run<0>("FFMA", 1, out, p.multiProcessorCount, ghz);
  run<10>("FFMA_imm", 1, out, p.multiProcessorCount, ghz);
  run<1>("FFMA2", 1, out, p.multiProcessorCount, ghz);
  run<4>("FADD", 1, out, p.multiProcessorCount, ghz);
  run<5>("FMUL", 1, out, p.multiProcessorCount, ghz);
  run<6>("FMNMX", 1, out, p.multiProcessorCount, ghz);
  run<2>("IADD+LOP3", 2, out, p.multiProcessorCount, ghz);
  run<8>("IMAD", 1, out, p.multiProcessorCount, ghz);
  run<3>("FFMA+LOP3", 2, out, p.multiProcessorCount, ghz);
  run<11>("FFMA+FADD", 2, out, p.multiProcessorCount, ghz);
  run<12>("FFMA+FMNMX", 2, out, p.multiProcessorCount, ghz);
  run<13>("FFMA2+LOP3", 2, out, p.multiProcessorCount, ghz);
  run<7>("F2I+I2F", 2, out, p.multiProcessorCount, ghz);
  run<9>("HADD2.F32 unpack", 1, out, p.multiProcessorCount, ghz);
  run<14>("MUFU.RCP", 1, out, p.multiProcessorCount, ghz);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
