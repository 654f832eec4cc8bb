// Microbenchmark: issue / pipe throughput of packed fp32x2 (FFMA2 / FADD2) versus scalar FFMA / FADD on
// sm_100a, alone and mixed with integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ unsigned iop(unsigned a, unsigned b) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE>
__global__ void k(float* out, int iters) {
  float s[8]; u64 v[8]; unsigned q[8];
  for (int i = 0; i < 8; ++i) { s[i] = threadIdx.x * 0.001f + i; v[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 1.f); q[i] = threadIdx.x + i; }
  const float c = 1.0001f; const u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) s[i] = fma1(s[i], c, c);                       // scalar FFMA
      if (MODE == 1) v[i] = fma2(v[i], c2, c2);                     // FFMA2
      if (MODE == 2) s[i] = add1(s[i], c);                          // scalar FADD
      if (MODE == 3) v[i] = add2(v[i], c2);                         // FADD2
      if (MODE == 4) { s[i] = fma1(s[i], c, c); q[i] = iop(q[i], 0x9e3779b9u); }          // FFMA + LOP3
      if (MODE == 5) { v[i] = fma2(v[i], c2, c2); q[i] = iop(q[i], 0x9e3779b9u); }        // FFMA2 + LOP3
      if (MODE == 6) { v[i] = fma2(v[i], c2, c2); q[i] = iop(q[i], 0x9e3779b9u); q[i] = iop(q[i], 0x7f4a7c15u); }  // FFMA2 + 2 LOP3
      if (MODE == 7) q[i] = iop(q[i], 0x9e3779b9u);                 // LOP3 alone
    }
  }
  float acc = 0.f;
  for (int i = 0; i < 8; ++i) acc += s[i] + __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32)) + (float)q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int ops_per_iter_elem, float* out) {
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  k<MODE><<<blocks, threads>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-instructions per SM per cycle assuming 1.965 GHz
  const double winst = (double)blocks * threads / 32 * iters * 8.0 * ops_per_iter_elem;
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-22s %8.3f ms  warp-instr/cycle/SM = %.2f\n", name, ms, winst / cyc / 148.0);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  run<0>("FFMA", 1, out);
  run<1>("FFMA2", 1, out);
  run<2>("FADD", 1, out);
  run<3>("FADD2", 1, out);
  run<7>("LOP3", 1, out);
  run<4>("FFMA + LOP3", 2, out);
  run<5>("FFMA2 + LOP3", 2, out);
  run<6>("FFMA2 + 2 LOP3", 3, out);
  return 0;
}
