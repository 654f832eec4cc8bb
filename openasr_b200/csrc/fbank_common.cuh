// Pieces shared by the fbank kernels: FFT geometry / shared-memory bank layout, RNG, reductions.
#pragma once
#include "fft_regs.cuh"
#include "spl_internal.cuh"

namespace spl {

// ---------------------------------------------------------------------------------------------
template <int NFFT>
struct Geo {
  static constexpr int R1 = 16;
  static constexpr int R2 = NFFT / 16;   // 32 or 16
  static constexpr int H = R2 / 2;       // useful k2 per lane (k < NFFT/2)
  static constexpr int NBIN = NFFT / 2;  // bins 0 .. NFFT/2-1 carry mel weight (Nyquist has none)
  static constexpr int EP = 17;          // exchange pitch (odd: conflict-free transposes)
  static constexpr int PL = ((R2 * EP + 15) / 32) * 32 + 16;  // per-pair plane stride, == 16 (mod 32)
  static constexpr int PLANE = 2 * PL;     // one plane (re or im) for both pairs
  static constexpr int PP = NBIN + 8;      // power-row pitch, == 8 (mod 32)
  static constexpr int RW = 2 * PLANE + 1; // per-warp region, == 1 (mod 32)
  static_assert(PL % 32 == 16 && PL >= R2 * EP, "pair offset must map to the other half of the banks");
  static_assert(PP % 32 == 8, "power pitch");
  static_assert(RW % 32 == 1, "warp region stride");
  static_assert(4 * PP <= 2 * PLANE, "power rows alias the exchange buffer");
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based; one call -> 4 x 32 random bits)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}

// The reference's one-uniform pseudo Box-Muller (kaldi_signal.py:176-177):
//   x = max(eps, u),  g = sqrt(-2 ln x) * cos(2 pi x),  u ~ U[0,1) with 24-bit resolution.
__device__ __forceinline__ float dither_from_bits(uint32_t r) {
  const float u = (float)(r >> 8) * 5.9604644775390625e-08f;  // 2^-24, same grid as torch.rand
  const float x = fmaxf(u, kEps);
  return sqrtf(-2.0f * __logf(x)) * __cosf(6.283185307179586f * x);
}

// lg2 / ln of a NORMAL positive float on the MUFU pipe (flush-to-zero form: no denormal pre-scaling code;
// every argument here is >= 2^-23, far from the denormal range)
__device__ __forceinline__ float fast_log2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_log(float x) { return fast_log2(x) * 0.6931471805599453f; }

__device__ __forceinline__ float group_sum(float v, int width) {
  // butterfly reduction inside aligned groups of `width` lanes (16 or 32)
  if (width == 32) v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

}  // namespace spl
