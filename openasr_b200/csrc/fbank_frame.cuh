// Device helpers shared by the persistent and the warp-pipelined fbank kernels:
// mbarrier / TMA bulk-copy PTX wrappers, the dither generator, and the per-frame pre-processing
// (kaldi_signal.py:174-199) in the stage-1 register layout.
#pragma once
#include "fbank_common.cuh"
#include "fft_c2.cuh"

namespace spl {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Philox4x32-7 (Salmon et al. 2011: 7 rounds is the Crush-resistant minimum; 10 is the default)
__device__ __forceinline__ uint4 philox4x32_7(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}

// Table-driven dither: byte offset (entry * 4) of the f-th 12-bit field of one Philox result, ten fields per call --
// two per word (bits 0..11, 12..23) and two more from the top bytes of word pairs (0, 1) and (2, 3).
constexpr int kDithPerCall = 10;
__device__ __forceinline__ uint32_t dith_off(const uint4& r, int f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  constexpr uint32_t kMask = (uint32_t)(kDitherTab - 1) << 2;
  if (f < 8) return ((f & 1) ? (w[f >> 1] >> 10) : (w[f >> 1] << 2)) & kMask;
  return (__byte_perm(w[2 * (f - 8)], w[2 * (f - 8) + 1], 0x0073) << 2) & kMask;
}

// ---------------------------------------------------------------------------------------------
// Compile-time frame geometry: lane n2 of the stage-1 layout holds samples j = R2*n1 + n2.
template <int NFFT, int NW>
struct FG {
  static constexpr int R2 = NFFT / 16;
  static constexpr bool kStatic = NW > 0;
  static constexpr int NROW = kStatic ? (NW + R2 - 1) / R2 : 16;  // rows that can be non-zero
  static constexpr int FULL = kStatic ? NW / R2 : 0;              // rows valid for every lane
  static constexpr int REM = kStatic ? NW - FULL * R2 : 0;        // lanes valid in row FULL
  static_assert(!kStatic || NROW > 8, "window must exceed half the padded size");
};

template <int NFFT, int NW>
__device__ __forceinline__ bool row_valid(int n1, int n2, int Nw) {
  using F = FG<NFFT, NW>;
  if constexpr (F::kStatic) return n1 < F::FULL || n2 < F::REM;
  return F::R2 * n1 + n2 < Nw;
}

// ---------------------------------------------------------------------------------------------
// Two frames at once in the packed stage-1 layout: z[n1] = (frame A sample, frame B sample) of
// j = R2 n1 + n2 after dither -> DC removal -> (raw log-energy) -> pre-emphasis -> window
// (kaldi_signal.py:174-199).  Frame A becomes the real part and frame B the imaginary part of the
// complex FFT input, so every arithmetic step is one fp32x2 instruction for both frames.
// ST: element type of the staged samples (float, or int16_t PCM converted at the first register load).
// DTAB: the device-RNG noise comes from the shared-memory table `dtab` of d g(u_i) (12-bit uniforms: two shifts/masks,
// one LDS and half a packed add per sample instead of three MUFU ops and a dozen ALU/FMA instructions).
template <int NFFT, int NW, bool NOISE, typename ST, bool DTAB = false>
__device__ __forceinline__ void load_frame_pair(c2 (&z)[16], const FbankParams& p, const ST* frA, const ST* frB,
                                                const float* win, float* energy_slots /*[2]*/, int n2, int b, int tA,
                                                int tB, bool validA, bool validB, const float* nz_utt /*[T, Nw] or NULL*/,
                                                const float* dtab = nullptr) {
  using G = Geo<NFFT>;
  using F = FG<NFFT, NW>;
  const int Nw = F::kStatic ? NW : p.Nw;
  c2 x[F::NROW];
#pragma unroll
  for (int n1 = 0; n1 < F::NROW; ++n1) {
    const bool rv = row_valid<NFFT, NW>(n1, n2, Nw);
    x[n1] = c2_make(rv ? (float)frA[G::R2 * n1 + n2] : 0.f, rv ? (float)frB[G::R2 * n1 + n2] : 0.f);
  }
  if constexpr (NOISE) {
    if (nz_utt != nullptr) {  // parity mode: host-drawn rand_gauss of this utterance, [T, Nw]
      const float* nzA = nz_utt + (size_t)tA * Nw;
      const float* nzB = nz_utt + (size_t)tB * Nw;
#pragma unroll
      for (int n1 = 0; n1 < F::NROW; ++n1)
        if (row_valid<NFFT, NW>(n1, n2, Nw)) {
          const int j = G::R2 * n1 + n2;
          x[n1] = c2_fma(c2_make(validA ? __ldg(nzA + j) : 0.f, validB ? __ldg(nzB + j) : 0.f), c2_splat(p.dither), x[n1]);
        }
    } else if constexpr (DTAB) {
      // counter-based stream keyed by (seed; b, t, n2, call): ten 12-bit table draws per Philox call.  Rows 0..9 of a
      // frame come from its own call; with NROW <= 15 the remaining rows of BOTH frames of the pair (tA even, tB = tA + 1)
      // share a third call keyed by tA (fields 0..4 -> frame A, 5..9 -> frame B): 3 calls per pair instead of 4.
      constexpr int kHead = F::NROW < kDithPerCall ? F::NROW : kDithPerCall;
      const char* tb = reinterpret_cast<const char*>(dtab);
      auto draw = [&](const uint4& rA, int fA, const uint4& rB, int fB) {
        return c2_make(*reinterpret_cast<const float*>(tb + dith_off(rA, fA)),
                       *reinterpret_cast<const float*>(tb + dith_off(rB, fB)));
      };
      {
        const uint4 rA = philox4x32_7(make_uint4((uint32_t)n2, (uint32_t)tA, (uint32_t)b, 0x5eedu), p.seed_lo, p.seed_hi);
        const uint4 rB = philox4x32_7(make_uint4((uint32_t)n2, (uint32_t)tB, (uint32_t)b, 0x5eedu), p.seed_lo, p.seed_hi);
#pragma unroll
        for (int n1 = 0; n1 < kHead; ++n1) {
          const c2 g = draw(rA, n1, rB, n1);
          if (row_valid<NFFT, NW>(n1, n2, Nw)) x[n1] = x[n1] + g;
        }
      }
      if constexpr (F::NROW > kDithPerCall && F::NROW <= kDithPerCall + kDithPerCall / 2) {
        const uint4 r = philox4x32_7(make_uint4((uint32_t)(G::R2 + n2), (uint32_t)tA, (uint32_t)b, 0x5eedu), p.seed_lo, p.seed_hi);
#pragma unroll
        for (int n1 = kDithPerCall; n1 < F::NROW; ++n1) {
          const c2 g = draw(r, n1 - kDithPerCall, r, n1 - kDithPerCall + kDithPerCall / 2);
          if (row_valid<NFFT, NW>(n1, n2, Nw)) x[n1] = x[n1] + g;
        }
      } else if constexpr (F::NROW > kDithPerCall) {
        const uint4 rA = philox4x32_7(make_uint4((uint32_t)(G::R2 + n2), (uint32_t)tA, (uint32_t)b, 0x5eedu), p.seed_lo, p.seed_hi);
        const uint4 rB = philox4x32_7(make_uint4((uint32_t)(G::R2 + n2), (uint32_t)tB, (uint32_t)b, 0x5eedu), p.seed_lo, p.seed_hi);
#pragma unroll
        for (int n1 = kDithPerCall; n1 < F::NROW; ++n1) {
          const c2 g = draw(rA, n1 - kDithPerCall, rB, n1 - kDithPerCall);
          if (row_valid<NFFT, NW>(n1, n2, Nw)) x[n1] = x[n1] + g;
        }
      }
    } else {  // 8-warp variant (no room for the table): the formula on a 16-bit grid, eight uniforms per call
      const float nk = 1.3862943611198906f * p.dither * p.dither;  // 2 ln 2 d^2
      const float sgn = p.dither < 0.f ? -1.f : 1.f;
#pragma unroll
      for (int c8 = 0; c8 * 8 < F::NROW; ++c8) {
        const uint4 rA = philox4x32_7(make_uint4((uint32_t)(c8 * G::R2 + n2), (uint32_t)tA, (uint32_t)b, 0x5eedu),
                                      p.seed_lo, p.seed_hi);
        const uint4 rB = philox4x32_7(make_uint4((uint32_t)(c8 * G::R2 + n2), (uint32_t)tB, (uint32_t)b, 0x5eedu),
                                      p.seed_lo, p.seed_hi);
        const uint32_t wA[4] = {rA.x, rA.y, rA.z, rA.w}, wB[4] = {rB.x, rB.y, rB.z, rB.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n1 = 8 * c8 + i;
          if (n1 < F::NROW) {
            // d sqrt(-2 ln u) cos(2 pi u), u = (v + 1/2) 2^-16 (kaldi_signal.py:176-177 on a 16-bit grid) for both
            // frames; the MUFU ops are scalar, the rest packed
            const float uA = (float)((i & 1) ? (wA[i >> 1] >> 16) : (wA[i >> 1] & 0xffffu)) + 0.5f;
            const float uB = (float)((i & 1) ? (wB[i >> 1] >> 16) : (wB[i >> 1] & 0xffffu)) + 0.5f;
            const c2 a = c2_fma(c2_make(fast_log2(uA), fast_log2(uB)), c2_splat(-nk), c2_splat(16.0001f * nk));  // > 0
            const c2 ang = c2_mul(c2_make(uA, uB), c2_splat(9.587379924285257e-05f));  // 2 pi 2^-16
            float csA, csB;
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(csA) : "f"(c2_re(ang)));
            asm("cos.approx.ftz.f32 %0, %1;" : "=f"(csB) : "f"(c2_im(ang)));
            const c2 g = c2_mul(c2_make(fast_sqrt(c2_re(a)), fast_sqrt(c2_im(a))), c2_make(csA, csB));
            if (row_valid<NFFT, NW>(n1, n2, Nw)) x[n1] = c2_fma(g, c2_splat(sgn), x[n1]);  // sign of `dither`
          }
        }
      }
    }
  }
  c2 sum = x[0];
#pragma unroll
  for (int n1 = 1; n1 < F::NROW; ++n1) sum = sum + x[n1];
  c2 mean = c2_splat(0.f);
  if (p.remove_dc)
    mean = c2_mul(c2_make(group_sum(c2_re(sum), G::R2), group_sum(c2_im(sum), G::R2)), c2_splat(1.0f / (float)Nw));
  if (p.use_energy) {
    c2 e = c2_splat(0.f);
#pragma unroll
    for (int n1 = 0; n1 < F::NROW; ++n1) {
      const c2 d = row_valid<NFFT, NW>(n1, n2, Nw) ? x[n1] - mean : c2_splat(0.f);
      e = c2_fma(d, d, e);
    }
    const float eA = group_sum(c2_re(e), G::R2), eB = group_sum(c2_im(e), G::R2);
    if (n2 == 0) {
      energy_slots[0] = fast_log(fmaxf(eA, kEps));
      energy_slots[1] = fast_log(fmaxf(eB, kEps));
    }
  }
  const float c = p.preemph;
  const c2 mu = c2_mul(mean, c2_splat(1.0f - c));
  const c2 mc = c2_splat(-c);
  [[maybe_unused]] c2 rot_prev = c2_splat(0.f);
#pragma unroll
  for (int n1 = 0; n1 < F::NROW; ++n1) {
    c2 prev;
    if constexpr (NOISE) {
      // previous sample of the noisy frame: the neighbouring lane's x[n1]; lane 0 takes the last lane's
      // x[n1 - 1], i.e. what it received from the previous row's rotation (j == 0: replicate padding)
      const int src = (n2 + G::R2 - 1) & (G::R2 - 1);
      const c2 rot = c2_make(__shfl_sync(0xffffffffu, c2_re(x[n1]), src, G::R2),
                             __shfl_sync(0xffffffffu, c2_im(x[n1]), src, G::R2));
      const c2 wrap = n1 == 0 ? x[0] : rot_prev;
      prev = (n2 == 0) ? wrap : rot;
      rot_prev = rot;
    } else {
      const int j = G::R2 * n1 + n2;
      const int jp = (row_valid<NFFT, NW>(n1, n2, Nw) ? j : 1) - 1;
      prev = (n1 == 0 && n2 == 0) ? x[0] : c2_make((float)frA[jp], (float)frB[jp]);
    }
    const bool rv = row_valid<NFFT, NW>(n1, n2, Nw);
    const float wj = rv ? win[G::R2 * n1 + n2] : 0.f;
    z[n1] = c2_mul(c2_fma(prev, mc, x[n1]) - mu, c2_splat(wj));
  }
  (void)validA;
  (void)validB;
}

}  // namespace spl
