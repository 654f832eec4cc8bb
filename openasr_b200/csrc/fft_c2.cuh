// Packed-complex arithmetic on Blackwell's fp32x2 pipe (sm_100a: add/sub/mul/fma.f32x2 -> SASS
// FADD2 / FMUL2 / FFMA2) and the in-register DIF FFT network of fft_regs.cuh restated on it.
//
// A complex value lives in one 64-bit register pair (re = low word, im = high word).  One packed
// instruction does the work of two scalar ones in ONE issue slot; the operand modifiers of the
// packed instructions (half swap .LO_HI, per-half negation, fp32 broadcast) make the complex
// twiddle products cost two instructions instead of four.  tools/micro/f32x2_bench.cu: FFMA2 runs at
// half the FFMA warp rate (same flops) but leaves the other issue slot to the integer / LDS / SHFL
// work that shares the kernel -- which is issue-bound (profiles/r1_summary.md), not pipe-bound.
#pragma once
#include "fft_regs.cuh"

namespace spl {

struct c2 {
  unsigned long long v;
};

__device__ __forceinline__ c2 c2_make(float re, float im) {
  c2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im));
  return r;
}
__device__ __forceinline__ float c2_re(c2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return lo;
}
__device__ __forceinline__ float c2_im(c2 a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
  return hi;
}
__device__ __forceinline__ c2 c2_swap(c2 a) { return c2_make(c2_im(a), c2_re(a)); }
__device__ __forceinline__ c2 operator+(c2 a, c2 b) {
  c2 r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ c2 operator-(c2 a, c2 b) {
  c2 r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
// element-wise (NOT complex) product / fused multiply-add of the two halves
__device__ __forceinline__ c2 c2_mul(c2 a, c2 b) {
  c2 r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ c2 c2_fma(c2 a, c2 b, c2 c) {
  c2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ c2 c2_splat(float s) { return c2_make(s, s); }

// t * (c - i s) = (tr c + ti s, ti c - tr s)
__device__ __forceinline__ c2 c2_mulw(c2 t, float c, float s) {
  return c2_fma(c2_swap(t), c2_make(s, -s), c2_mul(t, c2_splat(c)));
}
// t * (-i) = (ti, -tr)
__device__ __forceinline__ c2 c2_mul_mi(c2 t) { return c2_make(c2_im(t), -c2_re(t)); }

// One DIF butterfly: (a, b) -> (a + b, (a - b) * W_M^j),  W_M = e^{-2 pi i / M}.
template <int M, int J>
__device__ __forceinline__ void dif_bfly_c(c2& a, c2& b) {
  const c2 t = a - b;
  a = a + b;
  if constexpr (J == 0) {
    b = t;
  } else if constexpr (4 * J == M) {
    b = c2_mul_mi(t);
  } else {
    constexpr float c = cos64(J * (64 / M));
    constexpr float s = sin64(J * (64 / M));
    b = c2_mulw(t, c, s);
  }
}
// second input known to be zero: (a, 0) -> (a, a * W_M^j)
template <int M, int J>
__device__ __forceinline__ void dif_bfly_bzero_c(c2 a, c2& b) {
  if constexpr (J == 0) {
    b = a;
  } else if constexpr (4 * J == M) {
    b = c2_mul_mi(a);
  } else {
    constexpr float c = cos64(J * (64 / M));
    constexpr float s = sin64(J * (64 / M));
    b = c2_mulw(a, c, s);
  }
}

template <int N, int SPAN, int BASE, int J, int NZ>
struct DifStageJc {
  static __device__ __forceinline__ void run(c2 (&z)[N]) {
    if constexpr (2 * SPAN == N && BASE + J + SPAN >= NZ) {
      static_assert(BASE + J < NZ || NZ == 0, "at least half of the inputs must be populated");
      dif_bfly_bzero_c<2 * SPAN, J>(z[BASE + J], z[BASE + J + SPAN]);
    } else {
      dif_bfly_c<2 * SPAN, J>(z[BASE + J], z[BASE + J + SPAN]);
    }
    if constexpr (J + 1 < SPAN) DifStageJc<N, SPAN, BASE, J + 1, NZ>::run(z);
  }
};
template <int N, int SPAN, int BASE, int NZ>
struct DifStageBc {
  static __device__ __forceinline__ void run(c2 (&z)[N]) {
    DifStageJc<N, SPAN, BASE, 0, NZ>::run(z);
    if constexpr (BASE + 2 * SPAN < N) DifStageBc<N, SPAN, BASE + 2 * SPAN, NZ>::run(z);
  }
};
template <int N, int SPAN, int NZ>
struct DifAllc {
  static __device__ __forceinline__ void run(c2 (&z)[N]) {
    DifStageBc<N, SPAN, 0, NZ>::run(z);
    if constexpr (SPAN > 1) DifAllc<N, SPAN / 2, NZ>::run(z);
  }
};

// In-place forward DFT of N packed complex values held in registers; X[k] = z[bitrev<N>(k)].
// NZ: inputs z[NZ..N) are known zeros (need not be initialised; the first stage writes them).
template <int N, int NZ = N>
__device__ __forceinline__ void fft_dif_c(c2 (&z)[N]) {
  static_assert(N >= 2 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two <= 32");
  static_assert(NZ > N / 2 && NZ <= N, "pruning supports up to N/2 trailing zeros");
  DifAllc<N, N / 2, NZ>::run(z);
}

}  // namespace spl
