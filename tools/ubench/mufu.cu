// Micro-benchmark: throughput of the XU (MUFU) operations the conv epilogue could use for SiLU, per SM per clock.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/mufu tools/ubench/mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 3) { float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e + 1.0f)); }
  else if (OP == 4) { unsigned u, v; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(x)); asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v << 16); }
  else if (OP == 5) { y = fmaf(x, 1.0009765f, 0.5f); }   // FMA-pipe reference
  else { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(x)); y = __uint_as_float(u << 16); }
  return y;
}

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.001f * (threadIdx.x + j);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = op<OP>(a[j]);
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == 123.456f) out[threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms, float ghz) {
  float* d;
  cudaMalloc(&d, 4096);
  const int iters = 4096;
  k<OP><<<sms * 8, 256>>>(d, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<sms * 8, 256>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)sms * 8 * 256 * iters * 8;
  printf("{\"op\": \"%s\", \"ms\": %.3f, \"ops_per_clk_per_sm\": %.2f}\n", name, ms, ops / (ms * 1e-3) / (ghz * 1e9) / sms);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz / 1e6f;
  printf("{\"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.multiProcessorCount, ghz);
  run<5>("fma(ref)", p.multiProcessorCount, ghz);
  run<0>("tanh.approx.f32", p.multiProcessorCount, ghz);
  run<1>("ex2.approx.f32", p.multiProcessorCount, ghz);
  run<2>("rcp.approx.f32", p.multiProcessorCount, ghz);
  run<3>("ex2+add+rcp (sigmoid)", p.multiProcessorCount, ghz);
  run<4>("cvt+tanh.approx.bf16x2 (per instr)", p.multiProcessorCount, ghz);
  run<6>("cvt.rn.bf16x2.f32", p.multiProcessorCount, ghz);
  return 0;
}
