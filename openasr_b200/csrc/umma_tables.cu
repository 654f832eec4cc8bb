// Host-side tables of the tcgen05 DFT-as-GEMM engine (fbank_umma.cu): pre-swizzled FP16 hi/lo twiddle
// images (the B operand, streamed by TMA), the DC / Nyquist correction chunk, the delayed windows and the
// streaming mel program derived from the reference's dense bank (kaldi_signal.py:389-455).
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "spl_internal.cuh"

namespace spl {

namespace {

inline uint16_t h16(double v) {
  const __half h = __float2half_rn((float)v);
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}
inline double h16_back(uint16_t u) {
  __half h;
  std::memcpy(&h, &u, 2);
  return (double)__half2float(h);
}
// byte offset of element (row n, K column e in [0, 16)) inside a SWIZZLE_32B K-major FP16 tile
inline size_t sw32_h(int n, int e) { return (size_t)n * 32 + (size_t)((((e >> 3) ^ (n >> 2)) & 1) << 4) + (size_t)(e & 7) * 2; }

}  // namespace

void build_umma_tables(int nfft, int Nw, int D, const float* window, const float* mel_dense, UmmaHostTables& out) {
  out.ok = false;
  const int half = nfft / 4, nb = nfft / 2, nch = nb / 32;
  if ((nfft != 256 && nfft != 512) || Nw <= nb || Nw + 3 > nfft) return;

  // ---- streaming mel program over the steps: step s < half -> bin s + 1 ; s == half -> none ; s > half -> bin s ----
  std::vector<float> melw(2 * (size_t)nb, 0.f);
  std::vector<uint32_t> melc(nb / 16, 0u);
  int fA = 0;
  for (int m = 0; m < D; ++m)
    if (mel_dense[(size_t)m * nb] != 0.f) return;  // bin 0 must carry no weight (low_freq > 0)
  for (int s = 0; s < nb; ++s) {
    const int bin = s < half ? s + 1 : (s == half ? -1 : s);
    if (bin < 0) continue;
    int fmin = -1, fmax = -1;
    for (int m = 0; m < D; ++m)
      if (mel_dense[(size_t)m * nb + bin] != 0.f) {
        if (fmin < 0) fmin = m;
        fmax = m;
      }
    if (fmin < 0) continue;
    if (fmax - fmin > 1) return;  // more than two overlapping filters: not a triangular bank
    int ns = 0;
    while (fmax > fA + 1) {
      ++ns;
      ++fA;
    }
    if (fmin < fA || ns > 3) return;
    melc[s >> 4] |= (uint32_t)ns << (2 * (s & 15));
    melw[2 * s + 0] = mel_dense[(size_t)fA * nb + bin];
    melw[2 * s + 1] = fA + 1 < D ? mel_dense[(size_t)(fA + 1) * nb + bin] : 0.f;
  }
  out.nflush = D - fA;
  for (int s0 = 0; s0 < nb; s0 += 8) {  // the epilogue's staging ring holds 24 columns and is drained per 8 steps
    int e = 0;
    for (int s = s0; s < s0 + 8; ++s) e += (int)((melc[s >> 4] >> (2 * (s & 15))) & 3u);
    if (e > 16) return;
  }
  {  // the program must reproduce the dense bank exactly (same weights, same filters)
    std::vector<double> P(nb), E(D, 0.0), ref(D, 0.0);
    for (int b = 0; b < nb; ++b) P[b] = 1.0 + 0.37 * b + (double)((b * 2654435761u) % 97u);
    for (int m = 0; m < D; ++m)
      for (int b = 1; b < nb; ++b) ref[m] += (double)mel_dense[(size_t)m * nb + b] * P[b];
    double a = 0.0, bq = 0.0;
    int col = 0;
    auto emit = [&]() {
      if (col < D) E[col] = a;
      ++col;
      a = bq;
      bq = 0.0;
    };
    for (int s = 0; s < nb; ++s) {
      const int bin = s < half ? s + 1 : (s == half ? -1 : s);
      for (uint32_t ns = (melc[s >> 4] >> (2 * (s & 15))) & 3u; ns; --ns) emit();
      if (bin < 0) continue;
      a += (double)melw[2 * s] * P[bin];
      bq += (double)melw[2 * s + 1] * P[bin];
    }
    for (int i = 0; i < out.nflush; ++i) emit();
    if (col != D) return;
    for (int m = 0; m < D; ++m)
      if (std::fabs(E[m] - ref[m]) > 1e-9 * (1.0 + std::fabs(ref[m]))) return;
  }

  // epilogue parts: the steps are split over 4 (or 2, or 1) warps per TMEM lane quarter at multiples of 8, balanced
  // by cost (an 8-step iteration ~ 1 unit, 8 finished filters ~ 1.5 units: log, store, column sums); a part above
  // the first hands its first two emitted filters to the part below, so it needs at least two emits of its own
  out.nparts = 1;
  out.part_s0[0] = 0;
  out.part_s0[1] = nb;
  {
    const int ng = nb / 8;
    std::vector<int> ge(ng, 0);
    std::vector<double> cum(ng + 1, 0.0);
    for (int g = 0; g < ng; ++g) {
      for (int s = 8 * g; s < 8 * g + 8; ++s) ge[g] += (int)((melc[s >> 4] >> (2 * (s & 15))) & 3u);
      cum[g + 1] = cum[g] + 1.0 + 1.5 * ge[g] / 8.0;
    }
    for (int np = 4; np >= 2 && out.nparts == 1; np >>= 1) {
      int g0[5] = {0, 0, 0, 0, 0};
      g0[np] = ng;
      for (int pt = 1; pt < np; ++pt) {
        int g = g0[pt - 1] + 1;
        while (g < ng - (np - pt) && cum[g] < cum[ng] * pt / np) ++g;
        g0[pt] = g;
      }
      bool ok = true;
      int f = 0, f0[4] = {0, 0, 0, 0};
      for (int pt = 0; pt < np && ok; ++pt) {
        f0[pt] = f;
        int emits = 0;
        for (int g = g0[pt]; g < g0[pt + 1]; ++g) emits += ge[g];
        f += emits;
        if (g0[pt + 1] <= g0[pt] || (pt > 0 && emits < 2)) ok = false;
      }
      if (ok) {
        out.nparts = np;
        for (int pt = 0; pt < 4; ++pt) out.part_f0[pt] = f0[pt];
        for (int pt = 0; pt <= np; ++pt) out.part_s0[pt] = 8 * g0[pt];
      }
    }
  }

  const size_t b_tile = (size_t)half * 32, b_stage = 4 * b_tile;
  for (int fmt = 0; fmt < 2; ++fmt) {
    const int nshift = fmt == 0 ? 4 : 8;
    if (Nw + nshift - 1 > nfft) continue;  // this sample format stays on the FFT engine
    const int kx_n = (3 * nshift + 2 + 15) / 16;
    // ---- table blob: delayed windows | mel weights | shift codes ----
    std::vector<float>& tab = out.tab[fmt];
    tab.assign((size_t)nshift * nfft + 2 * (size_t)nb + (size_t)(nb / 16), 0.f);
    for (int h = 0; h < nshift; ++h)
      for (int j = 0; j < Nw; ++j) tab[(size_t)h * nfft + h + j] = window[j];
    out.off_melw[fmt] = nshift * nfft;
    out.off_melc[fmt] = nshift * nfft + 2 * nb;
    std::memcpy(tab.data() + (size_t)nshift * nfft, melw.data(), melw.size() * 4);
    std::memcpy(tab.data() + (size_t)nshift * nfft + 2 * (size_t)nb, melc.data(), melc.size() * 4);
    while (tab.size() % 4) tab.push_back(0.f);  // 16-byte multiple for the bulk copy

    // ---- twiddle images ----
    std::vector<uint8_t>& tw = out.twiddles[fmt];
    tw.assign((size_t)(nch + 1) * 2 * b_stage, 0);
    auto put = [&](size_t tile_off, int n, int e, uint16_t v) { std::memcpy(tw.data() + tile_off + sw32_h(n, e), &v, 2); };
    for (int c = 0; c < nch; ++c)
      for (int hf = 0; hf < 2; ++hf)
        for (int pp = 0; pp < 2; ++pp) {
          const size_t t_hi = ((size_t)c * 2 + hf) * b_stage + (size_t)(2 * pp) * b_tile, t_lo = t_hi + b_tile;
          for (int n = 0; n < half; ++n)
            for (int e = 0; e < 16; ++e) {
              const long j = 32 * c + 2 * e + pp, k = n + 1;
              const double ang = 2.0 * M_PI * (double)((j * k) % nfft) / (double)nfft;
              const double v = hf == 0 ? std::cos(ang) : std::sin(ang);
              const uint16_t hi = h16(v);
              put(t_hi, n, e, hi);
              put(t_lo, n, e, h16(v - h16_back(hi)));
            }
        }
    // correction chunk: DFT of the delayed windows at bins k and nb - k, combined for the four accumulators
    std::vector<double> Wc((size_t)nshift * (nb + 1)), Ws((size_t)nshift * (nb + 1));
    for (int h = 0; h < nshift; ++h)
      for (int b = 0; b <= nb; ++b) {
        double sc_ = 0.0, ss_ = 0.0;
        for (int j = 0; j < Nw; ++j) {
          const double ang = 2.0 * M_PI * (double)(((long)(j + h) * b) % nfft) / (double)nfft;
          sc_ += (double)window[j] * std::cos(ang);
          ss_ += (double)window[j] * std::sin(ang);
        }
        Wc[(size_t)h * (nb + 1) + b] = sc_;
        Ws[(size_t)h * (nb + 1) + b] = ss_;
      }
    for (int hf = 0; hf < 2; ++hf)
      for (int kx = 0; kx < kx_n; ++kx)
        for (int pp = 0; pp < 2; ++pp) {
          const size_t t0 = ((size_t)nch * 2 + hf) * b_stage + (size_t)(kx * 2 + pp) * b_tile;
          for (int n = 0; n < half; ++n) {
            const int k = n + 1;
            for (int e = 0; e < 16; ++e) {
              const int s = 16 * kx + e;
              double v = 0.0;
              bool lo_part = false;
              if (s < 3 * nshift) {
                const int h = s / 3;
                const double* C = &Wc[(size_t)h * (nb + 1)];
                const double* Sn = &Ws[(size_t)h * (nb + 1)];
                if (hf == 0) v = pp == 0 ? 0.5 * (C[k] + C[nb - k]) : 0.5 * (C[k] - C[nb - k]);
                else v = pp == 0 ? 0.5 * (Sn[k] - Sn[nb - k]) : 0.5 * (Sn[k] + Sn[nb - k]);
                lo_part = (s % 3) == 2;
              } else if ((s == 3 * nshift || s == 3 * nshift + 1) && hf == 0 && pp == 0) {
                v = (k & 1) ? -1.0 : 1.0;  // (-1)^k z'_{NB}
              }
              const uint16_t hi = h16(v);
              put(t0, n, e, lo_part ? h16(v - h16_back(hi)) : hi);
            }
          }
        }
  }
  out.ok = !out.tab[0].empty() || !out.tab[1].empty();
}

}  // namespace spl
