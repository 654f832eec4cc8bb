// Diagnostics that are not on the hot path: the device dither generator written out sample by sample, so that
// tests can check the noise ITSELF (moments, range, independence across overlapping frames) against the
// reference's rand_gauss (kaldi_signal.py:174-178) and feed it back through the host-noise mode.
#include "fbank_frame.cuh"

namespace spl {

// Same streams as load_frame_pair's device-RNG branches (fbank_frame.cuh) for (utterance b, frame t, sample j = R2 n1 + n2).
//  dtab != NULL (16-warp FFT engine): the f-th 12-bit field of Philox4x32-7(ctr = (c R2 + n2, tk, b, 0x5eed), key = seed)
//    indexes the table of d g(u_i); rows n1 < 10: c = 0, tk = t, f = n1; later rows: c = 1 and, when the frame has
//    at most 15 rows, tk = t & ~1 (the pair shares the call), f = n1 - 10 + 5 (t & 1), else tk = t, f = n1 - 10.
//  dtab == NULL (8-warp variant): the (n1 % 8)-th 16-bit word of ctr = ((n1 / 8) R2 + n2, t, b, 0x5eed) through the formula.
__global__ void dither_noise_kernel(float* __restrict__ out, int B, int T, int Nw, int R2, uint32_t seed_lo, uint32_t seed_hi,
                                    const float* __restrict__ dtab, float inv_d) {
  const size_t n = (size_t)B * T * Nw;
  // rows per lane as the engine's template sees them: compiled-in window (400 @ 512, 200 @ 256) or the generic 16
  const bool fixed = (R2 == 32 && Nw == 400) || (R2 == 16 && Nw == 200);
  const int nrow = fixed ? (Nw + R2 - 1) / R2 : 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Nw);
    const int t = (int)((i / Nw) % T), b = (int)(i / ((size_t)Nw * T));
    const int n1 = j / R2, n2 = j - n1 * R2;
    if (dtab != nullptr) {
      int c = 0, tk = t, f = n1;
      if (n1 >= kDithPerCall) {
        c = 1;
        f = n1 - kDithPerCall;
        if (nrow <= kDithPerCall + kDithPerCall / 2) {
          tk = t & ~1;
          f += (kDithPerCall / 2) * (t & 1);
        }
      }
      const uint4 r = philox4x32_7(make_uint4((uint32_t)(c * R2 + n2), (uint32_t)tk, (uint32_t)b, 0x5eedu), seed_lo, seed_hi);
      uint32_t off = 0;
#pragma unroll
      for (int ff = 0; ff < kDithPerCall; ++ff)
        if (ff == f) off = dith_off(r, ff);
      out[i] = dtab[off >> 2] * inv_d;
      continue;
    }
    const uint4 r = philox4x32_7(make_uint4((uint32_t)((n1 >> 3) * R2 + n2), (uint32_t)t, (uint32_t)b, 0x5eedu), seed_lo, seed_hi);
    const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
    const int k = n1 & 7;
    const float u = (float)((k & 1) ? (wd[k >> 1] >> 16) : (wd[k >> 1] & 0xffffu)) + 0.5f;
    const float a = fmaf(fast_log2(u), -1.3862943611198906f, 16.0001f * 1.3862943611198906f);
    float cs;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(u * 9.587379924285257e-05f));
    out[i] = fast_sqrt(a) * cs;
  }
}

cudaError_t launch_dither_noise(float* out, int B, int T, int Nw, int R2, uint64_t seed, const float* dtab, float dither,
                                cudaStream_t st) {
  if (dither == 0.f) dtab = nullptr;
  dither_noise_kernel<<<592, 256, 0, st>>>(out, B, T, Nw, R2, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), dtab,
                                           dither != 0.f ? 1.0f / dither : 1.0f);
  return cudaGetLastError();
}

}  // namespace spl
