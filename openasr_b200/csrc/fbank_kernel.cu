// Kernel A: fused  frame -> (dither) -> DC removal -> pre-emphasis -> window -> real FFT ->
// power -> sparse mel -> log  for one tile of 32 consecutive frames of one utterance.
//
// Replaces, per utterance, the ~100 ATen dispatches of
//   src/third_party/kaldi_signal.py:163-211 (_get_window) and :510-552 (fbank)
// and the pad/stack of src/blocks/sp_layers.py:87-91.
//
// Data flow inside a CTA (256 threads = 8 warps, 32 frames):
//   1. the tile's contiguous sample span (31*S + Nw samples) is staged ONCE in shared memory
//      (overlapping frames re-read shared memory, never HBM);
//   2. FFT phase, one warp per 4 frames: two real frames are packed into one complex
//      Nfft-point FFT (re = frame A, im = frame B).  Nfft = 16 x R2 (R2 = 32 @16 kHz, 16 @8 kHz):
//      stage 1 = radix-16 butterflies in registers (lane = n2), twiddle, transpose through a
//      padded, conflict-free shared-memory exchange, stage 2 = radix-R2 butterflies in registers
//      (lane = (pair, k1)); the Hermitian partner k <-> Nfft-k lives in lane (pair, 16-k1) and is
//      fetched with warp shuffles; power spectra of both frames are written to shared memory;
//   3. mel phase, lane = frame, warp = group of mel filters: sparse triangular filters
//      (<= 2 filters per bin), log(max(.,eps)), staged as a [32][D_out] tile;
//   4. coalesced store of the tile (rows past the utterance's frame count are written as 0),
//      per-utterance sum x / sum x^2 (CMVN + SpecAug time means) via fp64 atomics.
#include "fbank_common.cuh"

namespace spl {

struct SmemLayout {
  int span, off_win, off_melw, off_melidx, off_energy, off_warp, total;
};

__host__ __device__ inline SmemLayout make_layout(int nfft, int S, int Nw, int D, int D_out, int nnz) {
  SmemLayout L;
  L.span = (kTileFrames - 1) * S + Nw;
  int out_tile = kTileFrames * (D_out | 1);
  int samp = L.span > out_tile ? L.span : out_tile;  // the output tile aliases the sample span
  samp = (samp + 3) & ~3;
  L.off_win = samp;
  L.off_melw = L.off_win + ((Nw + 3) & ~3);
  L.off_melidx = L.off_melw + ((nnz + 3) & ~3);
  L.off_energy = L.off_melidx + 3 * D;
  L.off_warp = (L.off_energy + kTileFrames + 31) & ~31;
  int rw = nfft == 512 ? Geo<512>::RW : Geo<256>::RW;
  L.total = L.off_warp + kWarps * rw;
  return L;
}

size_t fbank_smem_bytes(int nfft, int S, int Nw, int D, int D_out, int nnz) {
  return sizeof(float) * (size_t)make_layout(nfft, S, Nw, D, D_out, nnz).total;
}

// Pre-processing of ONE frame in the stage-1 register layout: lane n2 holds j = R2*n1 + n2.
// Restates kaldi_signal.py:174-199 (dither, DC removal, raw log-energy, pre-emphasis, window).
template <int NFFT, bool NOISE>
__device__ __forceinline__ void load_frame(float (&z)[16], const FbankParams& p, const float* samp,
                                           const float* win, float* energy_slot, int fbase /*local sample offset*/,
                                           int n2, int b, int t /*global frame*/, bool frame_valid) {
  using G = Geo<NFFT>;
  const int Nw = p.Nw;
  float x[16];
  float sum = 0.f;
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int j = G::R2 * n1 + n2;
    x[n1] = (j < Nw) ? samp[fbase + j] : 0.f;
  }
  if constexpr (NOISE) {
    if (frame_valid) {
      if (p.noise != nullptr) {  // parity mode: host-drawn rand_gauss, [B, T, Nw]
        const float* nz = p.noise + ((size_t)b * p.T + t) * Nw;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int j = G::R2 * n1 + n2;
          if (j < Nw) x[n1] = fmaf(__ldg(nz + j), p.dither, x[n1]);
        }
      } else {  // throughput mode: counter-based stream keyed by (seed; b, t, j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (G::R2 * 4 * q < Nw) {
            const uint4 r = philox4x32_10(make_uint4(q * G::R2 + n2, (uint32_t)t, (uint32_t)b, 0x5eedu),
                                          p.seed_lo, p.seed_hi);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n1 = 4 * q + i;
              const int j = G::R2 * n1 + n2;
              if (j < Nw) x[n1] = fmaf(dither_from_bits(rr[i]), p.dither, x[n1]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) sum += x[n1];
  float mean = 0.f;
  if (p.remove_dc) mean = group_sum(sum, G::R2) / (float)Nw;
  if (p.use_energy) {
    float e = 0.f;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = G::R2 * n1 + n2;
      const float d = (j < Nw) ? x[n1] - mean : 0.f;
      e = fmaf(d, d, e);
    }
    e = group_sum(e, G::R2);
    if (n2 == 0) *energy_slot = __logf(fmaxf(e, kEps));
  }
  const float c = p.preemph;
  if constexpr (NOISE) {
    // previous sample of the *noisy* frame lives in the neighbouring lane
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const float up = __shfl_up_sync(0xffffffffu, x[n1], 1, G::R2);
      float wrap = x[0];  // j == 0: replicate padding (kaldi_signal.py:192-193)
      if (n1 > 0) wrap = __shfl_sync(0xffffffffu, x[n1 - 1], G::R2 - 1, G::R2);
      const float prev = (n2 == 0) ? wrap : up;
      const int j = G::R2 * n1 + n2;
      z[n1] = (j < Nw) ? ((x[n1] - mean) - c * (prev - mean)) * win[j] : 0.f;
    }
  } else {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = G::R2 * n1 + n2;
      const float prev = (j == 0) ? x[n1] : samp[fbase + (j < Nw ? j : 1) - 1];
      z[n1] = (j < Nw) ? ((x[n1] - mean) - c * (prev - mean)) * win[j] : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int NFFT, bool NOISE>
__global__ void __launch_bounds__(kThreads, 2) fbank_kernel(const FbankParams p) {
  using G = Geo<NFFT>;
  extern __shared__ __align__(16) float smem[];
  const SmemLayout L = make_layout(NFFT, p.S, p.Nw, p.D, p.D_out, p.tab.mel_nnz);
  float* samp = smem;
  float* out_tile = smem;  // aliases samp (used only after the FFT phase)
  float* win = smem + L.off_win;
  float* melw = smem + L.off_melw;
  int* mel_lo = reinterpret_cast<int*>(smem + L.off_melidx);
  int* mel_cnt = mel_lo + p.D;
  int* mel_off = mel_cnt + p.D;
  float* energy = smem + L.off_energy;
  float* warp_base = smem + L.off_warp;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kTileFrames;
  const int S = p.S, Nw = p.Nw, D = p.D, D_out = p.D_out;
  const int Dp = D_out | 1;

  const long long n_b = p.wav_len[b];
  const int m_b = n_b >= Nw ? (int)(1 + (n_b - Nw) / S) : 0;  // kaldi_signal.py:90
  if (blockIdx.x == 0 && tid == 0 && p.feat_len) p.feat_len[b] = m_b;
  int nvalid = m_b - t0;
  nvalid = nvalid < 0 ? 0 : (nvalid > kTileFrames ? kTileFrames : nvalid);
  const int rows = (p.T - t0) < kTileFrames ? (p.T - t0) : kTileFrames;  // rows of this tile inside [0,T)
  float* out_g = p.feats + ((size_t)b * p.T + t0) * D_out;

  if (nvalid == 0) {  // pure padding tile: exact zeros (sp_layers.py:88)
    for (int i = tid; i < rows * D_out; i += kThreads) out_g[i] = 0.f;
    return;
  }

  // ---- 1. stage samples + tables -------------------------------------------------------------
  {
    const int need = (nvalid - 1) * S + Nw;  // <= n_b - t0*S by construction
    const size_t g0 = (size_t)b * p.wav_pitch + (size_t)t0 * S;
    if (p.sample_format == SPL_SAMPLES_F32) {
      const float* src = static_cast<const float*>(p.wav) + g0;
      for (int i = tid; i < L.span; i += kThreads) samp[i] = i < need ? __ldg(src + i) : 0.f;
    } else {
      const int16_t* src = static_cast<const int16_t*>(p.wav) + g0;
      for (int i = tid; i < L.span; i += kThreads) samp[i] = i < need ? (float)__ldg(src + i) : 0.f;
    }
    for (int i = tid; i < Nw; i += kThreads) win[i] = __ldg(p.tab.window + i);
    for (int i = tid; i < p.tab.mel_nnz; i += kThreads) melw[i] = __ldg(p.tab.mel_w + i);
    for (int i = tid; i < D; i += kThreads) {
      mel_lo[i] = __ldg(p.tab.mel_lo + i);
      mel_cnt[i] = __ldg(p.tab.mel_cnt + i);
      mel_off[i] = __ldg(p.tab.mel_off + i);
    }
  }
  __syncthreads();

  // ---- 2. FFT phase: warp w owns local frames 4w .. 4w+3 ----------------------------------------
  float* wr = warp_base + w * G::RW;  // exchange: re plane at [0, PLANE), im plane at [PLANE, 2*PLANE)
  if (4 * w < nvalid) {
    if constexpr (NFFT == 512) {
      const int n2 = lane;
      float twr[16], twi[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        twr[k1] = __ldg(p.tab.tw_re + n2 * 16 + k1);
        twi[k1] = __ldg(p.tab.tw_im + n2 * 16 + k1);
      }
#pragma unroll 1
      for (int pr = 0; pr < 2; ++pr) {
        const int fa = 4 * w + 2 * pr;
        float re[16], im[16];
        load_frame<NFFT, NOISE>(re, p, samp, win, energy + fa, fa * S, n2, b, t0 + fa, fa < nvalid);
        load_frame<NFFT, NOISE>(im, p, samp, win, energy + fa + 1, (fa + 1) * S, n2, b, t0 + fa + 1,
                                fa + 1 < nvalid);
        fft_dif<16>(re, im);
        float* er = wr + pr * G::PL + n2 * G::EP;
        float* ei = er + G::PLANE;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          const float vr = re[bitrev<16>(k1)], vi = im[bitrev<16>(k1)];
          er[k1] = vr * twr[k1] + vi * twi[k1];  // * (c - i s)
          ei[k1] = vi * twr[k1] - vr * twi[k1];
        }
      }
    } else {
      const int pr = lane >> 4, n2 = lane & 15;
      const int fa = 4 * w + 2 * pr;
      float re[16], im[16];
      load_frame<NFFT, NOISE>(re, p, samp, win, energy + fa, fa * S, n2, b, t0 + fa, fa < nvalid);
      load_frame<NFFT, NOISE>(im, p, samp, win, energy + fa + 1, (fa + 1) * S, n2, b, t0 + fa + 1,
                              fa + 1 < nvalid);
      fft_dif<16>(re, im);
      float* er = wr + pr * G::PL + n2 * G::EP;
      float* ei = er + G::PLANE;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const float c = __ldg(p.tab.tw_re + n2 * 16 + k1), s = __ldg(p.tab.tw_im + n2 * 16 + k1);
        const float vr = re[bitrev<16>(k1)], vi = im[bitrev<16>(k1)];
        er[k1] = vr * c + vi * s;
        ei[k1] = vi * c - vr * s;
      }
    }
    __syncwarp();

    // stage 2: lane = (pair, k1), registers = n2
    const int pr = lane >> 4, k1 = lane & 15;
    float xr[G::R2], xi[G::R2];
    {
      const float* er = wr + pr * G::PL + k1;
      const float* ei = er + G::PLANE;
#pragma unroll
      for (int n2 = 0; n2 < G::R2; ++n2) {
        xr[n2] = er[n2 * G::EP];
        xi[n2] = ei[n2 * G::EP];
      }
    }
    __syncwarp();  // exchange buffer is dead from here on; the power rows alias it
    fft_dif<G::R2>(xr, xi);

    // Hermitian partner + power.  Z[k1 + 16 k2] sits at register bitrev(k2).
    const int partner = (lane & 16) | ((16 - k1) & 15);
    float* pa = wr + (2 * pr) * G::PP + k1;  // row of frame A (re part), frame B is the next row
    float* pb = pa + G::PP;
#pragma unroll
    for (int k2 = 0; k2 < G::H; ++k2) {
      const float zr = xr[bitrev<G::R2>(k2)], zi = xi[bitrev<G::R2>(k2)];
      float qr = __shfl_sync(0xffffffffu, xr[bitrev<G::R2>(G::R2 - 1 - k2)], partner);
      float qi = __shfl_sync(0xffffffffu, xi[bitrev<G::R2>(G::R2 - 1 - k2)], partner);
      if (k1 == 0) {  // own lane: N - k = 16 * (R2 - k2)
        qr = xr[bitrev<G::R2>((G::R2 - k2) & (G::R2 - 1))];
        qi = xi[bitrev<G::R2>((G::R2 - k2) & (G::R2 - 1))];
      }
      const float ar = zr + qr, ai = zi - qi;  // 2 * X_A[k]
      const float br = zi + qi, bi = qr - zr;  // 2 * X_B[k]
      pa[16 * k2] = ar * ar + ai * ai;         // 4 |X_A|^2  (the 1/4 is folded into the mel weights)
      pb[16 * k2] = br * br + bi * bi;
    }
  }
  __syncthreads();

  // ---- 3. mel phase: lane = local frame, warp = filter group -----------------------------------
  {
    const float* prow = warp_base + (lane >> 2) * G::RW + (lane & 3) * G::PP;
    const int m_beg = p.tab.grp_beg[w], m_end = p.tab.grp_beg[w + 1];
    const int col0 = p.use_energy ? 1 : 0;
    for (int m = m_beg; m < m_end; ++m) {
      const int lo = mel_lo[m], cnt = mel_cnt[m];
      const float* wv = melw + mel_off[m];
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(prow[lo + i], wv[i], acc);
      // kaldi_signal.py:540  log(max(E, eps))
      if (lane < nvalid) out_tile[lane * Dp + col0 + m] = __logf(fmaxf(acc, kEps));
    }
    if (p.use_energy && w == 0 && lane < nvalid) out_tile[lane * Dp] = energy[lane];
  }
  __syncthreads();

  // ---- 4. store + statistics ---------------------------------------------------------------------
  for (int i = tid; i < rows * D_out; i += kThreads) {
    const int r = i / D_out, c = i - r * D_out;
    out_g[i] = r < nvalid ? out_tile[r * Dp + c] : 0.f;
  }
  if (p.utt_stats != nullptr || p.global_stats != nullptr) {
    if (tid < D_out) {
      double s1 = 0.0, s2 = 0.0;  // fp64: sum x^2 - mean^2 must survive std << mean
      for (int r = 0; r < nvalid; ++r) {
        const double v = (double)out_tile[r * Dp + tid];
        s1 += v;
        s2 = fma(v, v, s2);
      }
      if (p.utt_stats) {
        atomicAdd(p.utt_stats + ((size_t)b * 2 + 0) * D_out + tid, s1);
        atomicAdd(p.utt_stats + ((size_t)b * 2 + 1) * D_out + tid, s2);
      }
      if (p.global_stats) {
        atomicAdd(p.global_stats + tid, s1);
        atomicAdd(p.global_stats + D_out + tid, s2);
      }
    }
    if (p.global_stats && tid == 0) atomicAdd(p.global_stats + 2 * D_out, (double)nvalid);
  }
}

// ---------------------------------------------------------------------------------------------
template <int NFFT, bool NOISE>
static cudaError_t launch_t(const FbankParams& p, cudaStream_t st) {
  const size_t smem = fbank_smem_bytes(NFFT, p.S, p.Nw, p.D, p.D_out, p.tab.mel_nnz);
  static thread_local size_t configured[8] = {0};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 8 && configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(fbank_kernel<NFFT, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = smem;
  } else if (dev >= 8) {
    cudaFuncSetAttribute(fbank_kernel<NFFT, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 grid((p.T + kTileFrames - 1) / kTileFrames, p.B);
  fbank_kernel<NFFT, NOISE><<<grid, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fbank(const FbankParams& p, int nfft, bool with_noise, cudaStream_t st) {
  if (nfft == 512) return with_noise ? launch_t<512, true>(p, st) : launch_t<512, false>(p, st);
  return with_noise ? launch_t<256, true>(p, st) : launch_t<256, false>(p, st);
}

}  // namespace spl
