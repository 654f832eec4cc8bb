// Kernel A (persistent, load-balanced): fused frame -> (dither) -> DC removal -> pre-emphasis ->
// window -> real FFT -> power -> sparse mel -> log.
//
// Replaces src/third_party/kaldi_signal.py:163-211 + :510-552 and the pad/stack loop of
// src/blocks/sp_layers.py:81-91.  Same arithmetic as fbank_kernel.cu (kept as the simple
// reference/fallback), re-organised for the B200:
//   * grid = 2 CTAs per SM, every CTA owns an equal share of the batch's flattened frame list
//     (sum_i m_i / gridDim frames), walked in chunks of <= 32 frames that never cross an
//     utterance -- no wave quantisation, no tail;
//   * the chunk's sample span arrives by ONE cp.async.bulk (TMA, 1-D) completing on an mbarrier,
//     issued right after the previous chunk's FFT phase so the HBM latency hides behind the mel
//     and store phases; every sample is read from HBM once per chunk;
//   * compile-time window length (400 @16 kHz / 200 @8 kHz): no per-element bounds tests, the
//     zero-padded rows of the radix-16 stage are pruned;
//   * Philox4x32-7 with five 24-bit uniforms per call, MUFU lg2/sqrt/cos for the dither;
//   * mel filters padded to groups of 4 bins: one broadcast LDS.128 of weights per 4 FMAs.
#include "fbank_frame.cuh"

namespace spl {

// ---------------------------------------------------------------------------------------------
struct PLayout {
  int span;        // sample slots (floats): 32*S + Nw + 8 (incl. the zeroed tail of an odd chunk)
  int op;          // output-row pitch inside a warp region, == 1 (mod 32)
  int off_tab, off_energy, off_pre, off_bar, off_warp, total;
};

// Per-warp region of the persistent kernel: [exchange planes | power rows alias them][output rows].
// RW == 4 (mod 32) keeps every power row 16-byte aligned and makes the mel phase's LDS.128
// (lane = frame) conflict-free per 8-lane phase; the output pitch == 1 (mod 32) does the same
// for its scalar writes.
template <int NFFT>
struct PGeo {
  static constexpr int RW = 2 * Geo<NFFT>::PLANE + 4;
  static_assert(RW % 32 == 4 && Geo<NFFT>::PP % 4 == 0, "aligned, conflict-free power rows");
};

__host__ __device__ inline int out_pitch(int D_out) { return (D_out + 30) / 32 * 32 + 1; }

__host__ __device__ inline PLayout make_playout(int nfft, int S, int Nw, int D_out, int ptab_words) {
  PLayout L;
  L.span = ((kTileFrames * S + Nw + 8) + 3) & ~3;
  L.op = out_pitch(D_out);
  L.off_tab = L.span;  // table block, bulk-copied verbatim from Tables::ptab (16-byte aligned)
  L.off_energy = L.off_tab + ptab_words;
  L.off_pre = L.off_energy + kTileFrames;
  L.off_bar = (L.off_pre + kMaxPersistentB + 1 + 1) & ~1;  // two 8-byte mbarriers
  L.off_warp = (L.off_bar + 4 + 31) & ~31;
  const int rw = nfft == 512 ? PGeo<512>::RW : PGeo<256>::RW;
  L.total = L.off_warp + kWarps * rw;
  return L;
}

size_t fbank_persistent_smem_bytes(int nfft, int S, int Nw, int D_out, int ptab_words) {
  return sizeof(float) * (size_t)make_playout(nfft, S, Nw, D_out, ptab_words).total;
}

// ---------------------------------------------------------------------------------------------
struct Chunk {
  int b, t0, nf;
};

template <int NFFT, int NW, bool NOISE>
__global__ void __launch_bounds__(kThreads, 2) fbank_persistent_kernel(const FbankParams p) {
  using G = Geo<NFFT>;
  using F = FG<NFFT, NW>;
  extern __shared__ __align__(128) float smem[];
  const int S = p.S, Nw = F::kStatic ? NW : p.Nw, D = p.D, D_out = p.D_out;
  const PLayout L = make_playout(NFFT, S, Nw, D_out, p.tab.ptab_words);
  float* samp = smem;
  float* tab = smem + L.off_tab;
  const float4* melw8 = reinterpret_cast<const float4*>(tab);  // pair weights {wa[4], wb[4]} per group
  const uint32_t* pdesc = reinterpret_cast<const uint32_t*>(tab + p.tab.pt_off_desc);
  const float* win = tab + p.tab.pt_off_win;
  const float* tws = tab + p.tab.pt_off_tw;  // tws[k1 * R2 + n2] = cos, tws[NFFT + k1 * R2 + n2] = sin
  float* energy = smem + L.off_energy;
  int* pre = reinterpret_cast<int*>(smem + L.off_pre);  // pre[b] = frames of utterances < b ; pre[B] = total
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + L.off_bar);  // [0] samples, [1] tables
  float* warp_base = smem + L.off_warp;
  const int OP = L.op;
  constexpr int OUT_OFF = 4 * G::PP;  // output rows live above the power rows of the owning warp region
  constexpr int RW = PGeo<NFFT>::RW;
  (void)D;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int B = p.B, T = p.T;

  // ---- 0. tables (one TMA bulk copy), frame prefix (warp 0) ------------------------------------
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(mbar + 1, (uint32_t)p.tab.ptab_words * 4u);
    bulk_g2s(tab, p.tab.ptab, (uint32_t)p.tab.ptab_words * 4u, mbar + 1);
  }
  if (w == 0) {
    int carry = 0;
    for (int base = 0; base < B; base += 32) {
      const int bb = base + lane;
      int m = 0;
      if (bb < B) {
        const long long n = p.wav_len[bb];
        m = n >= Nw ? (int)(1 + (n - Nw) / S) : 0;  // kaldi_signal.py:90
        m = m > T ? T : m;
        if (blockIdx.x == 0 && p.feat_len) p.feat_len[bb] = m;
      }
      int incl = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (bb < B) pre[bb] = carry + incl - m;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) pre[B] = carry;
  }
  __syncthreads();

  const long long total = pre[B];
  const long long nct = gridDim.x;
  const int r0 = (int)(total * blockIdx.x / nct), r1 = (int)(total * (blockIdx.x + 1) / nct);

  auto find_utt = [&](int pos) {  // largest b with pre[b] <= pos  (pos < total)
    int lo = 0, hi = B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
  };
  auto chunk_at = [&](int pos, int bhint) {
    Chunk c;
    int b = bhint;
    while (pre[b + 1] <= pos) ++b;
    c.b = b;
    c.t0 = pos - pre[b];
    int nf = pre[b + 1] - pos;
    nf = nf > r1 - pos ? r1 - pos : nf;
    c.nf = nf > kTileFrames ? kTileFrames : nf;
    return c;
  };
  // Stage the chunk's samples; returns the head offset of the first sample inside `samp`.
  // Bulk path (fp32, 16-byte aligned window inside the wav buffer): one elected thread arms the
  // mbarrier and issues the copy; everybody waits on the barrier before the FFT phase.
  const char* wav_lo = static_cast<const char*>(p.wav);
  const size_t esz = p.sample_format == SPL_SAMPLES_F32 ? 4 : 2;
  const char* wav_hi = wav_lo + ((size_t)(B - 1) * p.wav_pitch + (size_t)p.wav_cols) * esz;
  auto bulk_plan = [&](const Chunk& c, const char*& a0, uint32_t& bytes, int& head) {
    if (p.sample_format != SPL_SAMPLES_F32) return false;
    const int need = (c.nf - 1) * S + Nw;
    const char* src = wav_lo + ((size_t)c.b * p.wav_pitch + (size_t)c.t0 * S) * 4;
    a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
    head = (int)((src - a0) >> 2);
    bytes = (uint32_t)(((head + need) * 4 + 15) & ~15);
    return a0 >= wav_lo && a0 + bytes <= wav_hi;
  };

  uint32_t parity = 0;
  Chunk cur{0, 0, 0};
  bool cur_bulk = false;
  int cur_head = 0;
  if (r0 < r1) {
    cur = chunk_at(r0, find_utt(r0));
    const char* a0;
    uint32_t bytes;
    cur_bulk = bulk_plan(cur, a0, bytes, cur_head);
    if (cur_bulk && tid == 0) {
      mbar_expect_tx(mbar, bytes);
      bulk_g2s(samp, a0, bytes, mbar);
    }
  }

  // ---- 0b. zero padding rows: an equal share of the B*T - total padded rows per CTA (sp_layers.py:88)
  {
    const long long total_pad = (long long)B * T - total;
    int q = (int)(total_pad * blockIdx.x / nct);
    const int q1 = (int)(total_pad * (blockIdx.x + 1) / nct);
    if (q < q1) {
      int lo = 0, hi = B - 1;  // largest b with ppre[b] = b*T - pre[b] <= q
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)mid * T - pre[mid] <= q) lo = mid; else hi = mid - 1;
      }
      int b = lo;
      while (q < q1) {
        while ((long long)(b + 1) * T - pre[b + 1] <= q) ++b;
        const int m_b = pre[b + 1] - pre[b];
        const int ofs = q - (int)((long long)b * T - pre[b]);
        int nrows = (T - m_b) - ofs;
        nrows = nrows > q1 - q ? q1 - q : nrows;
        float* dst = p.feats + ((size_t)b * T + m_b + ofs) * D_out;
        for (int i = tid; i < nrows * D_out; i += kThreads) dst[i] = 0.f;
        q += nrows;
      }
    }
  }

  mbar_wait(mbar + 1, 0);  // tables have landed

  // ---- main loop over this CTA's chunks ----------------------------------------------------------
  int pos = r0;
  while (pos < r1) {
    const Chunk c = cur;
    const int nf = c.nf;
    const int need = (nf - 1) * S + Nw;
    int head = cur_head;
    if (cur_bulk) {
      mbar_wait(mbar, parity);
      parity ^= 1;
    } else {  // scalar staging (int16 ingest, unaligned or boundary windows)
      head = 0;
      const size_t g0 = (size_t)c.b * p.wav_pitch + (size_t)c.t0 * S;
      if (p.sample_format == SPL_SAMPLES_F32) {
        const float* src = static_cast<const float*>(p.wav) + g0;
        for (int i = tid; i < need; i += kThreads) samp[i] = __ldg(src + i);
      } else {
        const int16_t* src = static_cast<const int16_t*>(p.wav) + g0;
        for (int i = tid; i < need; i += kThreads) samp[i] = (float)__ldg(src + i);
      }
      __syncthreads();
    }
    const float* sbase = samp + head;

    // ---- FFT phase: warp w owns local frames 4w .. 4w+3 ----
    float* wr = warp_base + w * RW;
    if (4 * w < nf) {
      if ((nf & 1) && (nf >> 2) == w) {  // odd chunk: the invalid partner frame must be finite -> zero its tail
        for (int i = lane; i < S; i += 32) samp[head + need + i] = 0.f;
        __syncwarp();
      }
      if constexpr (NFFT == 512) {
        const int n2 = lane;
        float twr[16], twi[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          twr[k1] = tws[k1 * G::R2 + n2];
          twi[k1] = tws[NFFT + k1 * G::R2 + n2];
        }
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          const int fa = 4 * w + 2 * pr;
          float re[16], im[16];
          load_frame_p<NFFT, NW, NOISE>(re, p, sbase + fa * S, win, energy + fa, n2, c.b, c.t0 + fa, fa < nf);
          load_frame_p<NFFT, NW, NOISE>(im, p, sbase + (fa + 1) * S, win, energy + fa + 1, n2, c.b, c.t0 + fa + 1,
                                        fa + 1 < nf);
          fft_dif<16, F::NROW>(re, im);
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
        load_frame_p<NFFT, NW, NOISE>(re, p, sbase + fa * S, win, energy + fa, n2, c.b, c.t0 + fa, fa < nf);
        load_frame_p<NFFT, NW, NOISE>(im, p, sbase + (fa + 1) * S, win, energy + fa + 1, n2, c.b, c.t0 + fa + 1,
                                      fa + 1 < nf);
        fft_dif<16, F::NROW>(re, im);
        float* er = wr + pr * G::PL + n2 * G::EP;
        float* ei = er + G::PLANE;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          const float cs = tws[k1 * G::R2 + n2], sn = tws[NFFT + k1 * G::R2 + n2];
          const float vr = re[bitrev<16>(k1)], vi = im[bitrev<16>(k1)];
          er[k1] = vr * cs + vi * sn;
          ei[k1] = vi * cs - vr * sn;
        }
      }
      __syncwarp();

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
      __syncwarp();  // exchange buffer dead; the power rows alias it
      fft_dif<G::R2>(xr, xi);

      const int partner = (lane & 16) | ((16 - k1) & 15);
      float* pa = wr + (2 * pr) * G::PP + k1;
      float* pb = pa + G::PP;
#pragma unroll
      for (int k2 = 0; k2 < G::H; ++k2) {
        const float zr = xr[bitrev<G::R2>(k2)], zi = xi[bitrev<G::R2>(k2)];
        float qr = __shfl_sync(0xffffffffu, xr[bitrev<G::R2>(G::R2 - 1 - k2)], partner);
        float qi = __shfl_sync(0xffffffffu, xi[bitrev<G::R2>(G::R2 - 1 - k2)], partner);
        if (k1 == 0) {
          qr = xr[bitrev<G::R2>((G::R2 - k2) & (G::R2 - 1))];
          qi = xi[bitrev<G::R2>((G::R2 - k2) & (G::R2 - 1))];
        }
        const float ar = zr + qr, ai = zi - qi;
        const float br = zi + qi, bi = qr - zr;
        pa[16 * k2] = ar * ar + ai * ai;  // 4 |X_A|^2 (the 1/4 is folded into the mel weights)
        pb[16 * k2] = br * br + bi * bi;
      }
    }
    __syncthreads();  // power rows complete; samples dead

    // ---- prefetch the next chunk's samples behind the mel / store phases ----
    const int npos = pos + nf;
    Chunk nxt{c.b, 0, 0};
    bool nxt_bulk = false;
    int nxt_head = 0;
    if (npos < r1) {
      nxt = chunk_at(npos, c.b);
      const char* a0;
      uint32_t bytes;
      nxt_bulk = bulk_plan(nxt, a0, bytes, nxt_head);
      if (nxt_bulk && tid == 0) {
        fence_proxy_async();  // generic-proxy reads of `samp` above happen-before the async-proxy writes
        mbar_expect_tx(mbar, bytes);
        bulk_g2s(samp, a0, bytes, mbar);
      }
    }

    // ---- mel phase: lane = local frame, warp = group of filter PAIRS (two independent FMA chains).
    //      For chunks of <= 16 frames the half-warps take alternate pairs so that no lane idles. ----
    {
      const bool split = nf <= 16;
      const int fr = split ? (lane & 15) : lane;
      const int hsel = split ? (lane >> 4) : 0;
      const int istep = split ? 2 : 1;
      const float4* prow4 = reinterpret_cast<const float4*>(warp_base + (fr >> 2) * RW + (fr & 3) * G::PP);
      float* orow = warp_base + (fr >> 2) * RW + OUT_OFF + (fr & 3) * OP + (p.use_energy ? 1 : 0);
      const int i_beg = p.tab.pgrp_beg[w], i_end = p.tab.pgrp_beg[w + 1];
      for (int i = i_beg + hsel; i < i_end; i += istep) {
        const uint32_t dsc = pdesc[i];
        const float4* pa4 = prow4 + (dsc & 63u);
        const float4* pb4 = prow4 + ((dsc >> 6) & 63u);
        const float4* wv = melw8 + 2 * ((dsc >> 18) & 8191u);
        const int n4 = (dsc >> 12) & 63u;
        float accA = 0.f, accB = 0.f;
#pragma unroll 2
        for (int g = 0; g < n4; ++g) {
          const float4 pa = pa4[g], pb = pb4[g];
          const float4 wa = wv[2 * g], wb = wv[2 * g + 1];
          accA = fmaf(pa.x, wa.x, accA);
          accB = fmaf(pb.x, wb.x, accB);
          accA = fmaf(pa.y, wa.y, accA);
          accB = fmaf(pb.y, wb.y, accB);
          accA = fmaf(pa.z, wa.z, accA);
          accB = fmaf(pb.z, wb.z, accB);
          accA = fmaf(pa.w, wa.w, accA);
          accB = fmaf(pb.w, wb.w, accB);
        }
        orow[2 * i] = fast_log(fmaxf(accA, kEps));  // kaldi_signal.py:540
        if (dsc >> 31) orow[2 * i + 1] = fast_log(fmaxf(accB, kEps));
      }
      if (p.use_energy && w == 0 && hsel == 0) orow[-1] = energy[fr];
    }
    __syncthreads();

    // ---- store + statistics ----
    {
      float* out_g = p.feats + ((size_t)c.b * T + c.t0 + 4 * w) * D_out;  // rows 4w..4w+3 live in this warp's region
      const float* obase = warp_base + w * RW + OUT_OFF;
      const int nrow = nf - 4 * w < 4 ? nf - 4 * w : 4;
      if ((D_out & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.feats) & 15) == 0)) {
        const int q = D_out >> 2;
        for (int rr = 0; rr < nrow; ++rr) {
          const float* orow = obase + rr * OP;
          float4* dst = reinterpret_cast<float4*>(out_g + (size_t)rr * D_out);
          for (int cc = lane; cc < q; cc += 32)
            dst[cc] = make_float4(orow[4 * cc], orow[4 * cc + 1], orow[4 * cc + 2], orow[4 * cc + 3]);
        }
      } else {
        for (int rr = 0; rr < nrow; ++rr) {
          const float* orow = obase + rr * OP;
          for (int cc = lane; cc < D_out; cc += 32) out_g[(size_t)rr * D_out + cc] = orow[cc];
        }
      }
      if (p.utt_stats != nullptr || p.global_stats != nullptr) {
        if (tid < D_out) {
          double s1 = 0.0, s2 = 0.0;  // fp64: sum x^2 - mean^2 must survive std << mean
          for (int r = 0; r < nf; ++r) {
            const double v = (double)warp_base[(r >> 2) * RW + OUT_OFF + (r & 3) * OP + tid];
            s1 += v;
            s2 = fma(v, v, s2);
          }
          if (p.utt_stats) {
            atomicAdd(p.utt_stats + ((size_t)c.b * 2 + 0) * D_out + tid, s1);
            atomicAdd(p.utt_stats + ((size_t)c.b * 2 + 1) * D_out + tid, s2);
          }
          if (p.global_stats) {
            atomicAdd(p.global_stats + tid, s1);
            atomicAdd(p.global_stats + D_out + tid, s2);
          }
        }
        if (p.global_stats && tid == 0) atomicAdd(p.global_stats + 2 * D_out, (double)nf);
      }
    }
    __syncthreads();  // output rows alias the exchange buffers of the next chunk

    pos = npos;
    cur = nxt;
    cur_bulk = nxt_bulk;
    cur_head = nxt_head;
  }
}

// ---------------------------------------------------------------------------------------------
template <int NFFT, int NW, bool NOISE>
static cudaError_t launch_p(const FbankParams& p, int num_ctas, cudaStream_t st) {
  const size_t smem = fbank_persistent_smem_bytes(NFFT, p.S, NW > 0 ? NW : p.Nw, p.D_out, p.tab.ptab_words);
  static thread_local size_t configured[16] = {0};  // per device, per instantiation
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(fbank_persistent_kernel<NFFT, NW, NOISE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = smem;
  }
  fbank_persistent_kernel<NFFT, NW, NOISE><<<num_ctas, kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fbank_persistent(const FbankParams& p, int nfft, bool with_noise, int num_ctas, cudaStream_t st) {
  if (nfft == 512) {
    if (p.Nw == 400) return with_noise ? launch_p<512, 400, true>(p, num_ctas, st) : launch_p<512, 400, false>(p, num_ctas, st);
    return with_noise ? launch_p<512, 0, true>(p, num_ctas, st) : launch_p<512, 0, false>(p, num_ctas, st);
  }
  if (p.Nw == 200) return with_noise ? launch_p<256, 200, true>(p, num_ctas, st) : launch_p<256, 200, false>(p, num_ctas, st);
  return with_noise ? launch_p<256, 0, true>(p, num_ctas, st) : launch_p<256, 0, false>(p, num_ctas, st);
}

}  // namespace spl
