// Kernel A (pair-pipelined, 16 kHz / Nfft = 512): the warp-pipelined design of fbank_warp.cu
// re-cut so that 24 warps are resident per SM instead of 16.
//
// Replaces src/third_party/kaldi_signal.py:163-211 + :510-552 and the pad/stack loop of
// src/blocks/sp_layers.py:81-91; identical arithmetic up to fp32 summation order.
//
// ncu on fbank_warp.cu (profiles/r1_summary.md) shows the kernel bound by issue slots left idle
// on dependent shared-memory / shuffle latencies with only 4 warps per scheduler: its radix-32
// stage keeps 64 live registers per lane and its 4-frame group needs 11.6 KB of shared memory per
// warp.  Here the unit of work is ONE packed pair (2 frames = one complex FFT):
//   * stage 1: lane = n2 (32 lanes), radix-16 over n1 in registers, twiddles read from shared
//     memory instead of being pinned in 32 registers;
//   * stage 2: the radix-32 over n2 = 4 b + a is split over TWO lanes per k1: lane type A owns the
//     quarters a in {0, 2}, type B a in {1, 3}.  Each lane runs two radix-8 networks, applies
//     W_32^(a k2'), and the final radix-4 across quarters needs one complex exchange with the
//     sibling lane (lane ^ 16): A ends with bins k2'' in {0, 2}, B with {1, 3} (k2 = k2' + 8 k2'');
//   * Hermitian split: a lane keeps bins k2'' = typ (k < 256) and receives the mirrored bins
//     N - k, which are exactly the k2'' = 3 - typ values of lane (16 - k1, other type);
//   * mel: lane = (frame, 16 slices), two independent filter streams per slice, flat entry
//     lists balanced on the host, so every lane executes the same E iterations.
// <= 80 registers and 6.4 KB of shared memory per warp: 2 CTAs x 12 warps per SM.
#include "fbank_frame.cuh"

namespace spl {

constexpr int kPWarps = 12;
constexpr int kPThreads = kPWarps * 32;
constexpr int kPStatUtts = 4;

// exchange planes for one pair, Nfft = 512:  element (n2 = 4 b + a, k1) of a plane sits at
//   66 b + 16 a + (a >> 1) + k1
// stage-1 writes (lane = n2, fixed k1) hit banks 2 b + {0, 16, 1, 17}[a]  -> all 32 distinct;
// stage-2 reads (lane = (k1, typ), fixed b, a = typ or typ + 2)           -> k1 + 16 typ (+ c).
constexpr int kQB = 66;             // stride of b
constexpr int kQPlane = 8 * kQB;    // 528 floats per plane (re, im)
constexpr int kQPP = 272;           // power-row pitch == 16 (mod 32): A / B lanes write opposite bank halves
constexpr int kQSamp = 2 * kQPP;    // sample buffer starts behind the two power rows
constexpr int kQRegion = kQSamp + 564;  // head (<= 3) + S + Nw = 560 samples, rounded to 16 bytes
static_assert(kQRegion >= 2 * kQPlane && kQRegion % 4 == 0, "region");
__host__ __device__ constexpr int qaddr(int n2) { return kQB * (n2 >> 2) + 16 * (n2 & 3) + ((n2 >> 1) & 1); }

struct PLayout {
  int op, out_off, en_off, st_off, rw;
  int off_gpre, off_fpre, off_stat, off_ctl, off_bar, off_warp, total;
};

__host__ __device__ inline PLayout make_playout(int D_out, int qtab_words) {
  PLayout L;
  L.op = (D_out + 3) & ~3;
  L.out_off = kQRegion;
  L.en_off = L.out_off + 2 * L.op;
  L.st_off = L.en_off + 4;  // fp64 rows, 8-byte aligned (everything before is a multiple of 4)
  L.rw = L.st_off + 4 * L.op;
  L.off_gpre = qtab_words;
  L.off_fpre = L.off_gpre + kMaxPersistentB + 1;
  L.off_stat = (L.off_fpre + kMaxPersistentB + 1 + 3) & ~3;
  L.off_ctl = L.off_stat + kPStatUtts * 4 * L.op;
  L.off_bar = (L.off_ctl + 8 + 1) & ~1;
  L.off_warp = (L.off_bar + 2 * (kPWarps + 1) + 31) & ~31;
  L.total = L.off_warp + kPWarps * L.rw;
  return L;
}

size_t fbank_pair_smem_bytes(int D_out, int qtab_words) {
  return sizeof(float) * (size_t)make_playout(D_out, qtab_words).total;
}

// ---------------------------------------------------------------------------------------------
template <int NW, bool NOISE>
__global__ void __launch_bounds__(kPThreads, 2) fbank_pair_kernel(const FbankParams p) {
  constexpr int NFFT = 512;
  using G = Geo<NFFT>;
  using F = FG<NFFT, NW>;
  extern __shared__ __align__(128) float smem[];
  const int S = p.S, Nw = F::kStatic ? NW : p.Nw, D_out = p.D_out;
  const PLayout L = make_playout(D_out, p.tab.qtab_words);
  float* tab = smem;
  const float4* melw = reinterpret_cast<const float4*>(tab);
  const uint2* mdesc = reinterpret_cast<const uint2*>(tab + p.tab.qt_off_desc);
  const float* win = tab + p.tab.qt_off_win;
  const float* tws = tab + p.tab.qt_off_tw;
  const float4* tw2 = reinterpret_cast<const float4*>(tab + p.tab.qt_off_tw2);
  int* gpre = reinterpret_cast<int*>(smem + L.off_gpre);  // gpre[b] = pairs of utterances < b
  int* fpre = reinterpret_cast<int*>(smem + L.off_fpre);  // fpre[b] = frames of utterances < b
  double* cstat = reinterpret_cast<double*>(smem + L.off_stat);
  int* ctl = reinterpret_cast<int*>(smem + L.off_ctl);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int B = p.B, T = p.T, OP = L.op;
  float* wr = smem + L.off_warp + w * L.rw;
  float* e0 = wr;             // exchange planes (re at 0, im at kQPlane); the power rows alias [0, 2 kQPP)
  float* samp = wr + kQSamp;  // sample buffer: aliases the tail of the exchange planes
  float* orows = wr + L.out_off;
  float* energy = wr + L.en_off;
  double* wstat = reinterpret_cast<double*>(wr + L.st_off);
  uint64_t* mybar = bars + 1 + w;

  // ---- 0. tables (one TMA bulk copy), pair / frame prefix sums (warp 0), barriers ----------------
  if (tid == 0) {
    mbar_init(bars, 1);
    for (int i = 0; i < kPWarps; ++i) mbar_init(bars + 1 + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bars, (uint32_t)p.tab.qtab_words * 4u);
    bulk_g2s(tab, p.tab.qtab, (uint32_t)p.tab.qtab_words * 4u, bars);
  }
  for (int i = tid; i < kPStatUtts * 2 * OP; i += kPThreads) cstat[i] = 0.0;
  if (w == 0) {
    int carry = 0, fcarry = 0;
    for (int base = 0; base < B; base += 32) {
      const int bb = base + lane;
      int m = 0;
      if (bb < B) {
        const long long n = p.wav_len[bb];
        m = n >= Nw ? (int)(1 + (n - Nw) / S) : 0;  // kaldi_signal.py:90
        m = m > T ? T : m;
        if (blockIdx.x == 0 && p.feat_len) p.feat_len[bb] = m;
      }
      const int gcount = (m + 1) >> 1;
      int incl = gcount, finc = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        const int u = __shfl_up_sync(0xffffffffu, finc, o);
        if (lane >= o) {
          incl += v;
          finc += u;
        }
      }
      if (bb < B) {
        gpre[bb] = carry + incl - gcount;
        fpre[bb] = fcarry + finc - m;
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
      fcarry += __shfl_sync(0xffffffffu, finc, 31);
    }
    if (lane == 0) {
      gpre[B] = carry;
      fpre[B] = fcarry;
    }
  }
  for (int i = lane; i < 2 * OP; i += 32) wstat[i] = 0.0;
  __syncthreads();
  if (tid == 0) {  // this CTA's contiguous share of the pair list and of the zero-padding rows
    const long long NG = gpre[B];
    const int g0 = (int)(NG * blockIdx.x / gridDim.x), g1 = (int)(NG * (blockIdx.x + 1) / gridDim.x);
    int lo = 0, hi = B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (gpre[mid] <= g0) lo = mid; else hi = mid - 1;
    }
    ctl[0] = g0;
    ctl[1] = g1;
    ctl[2] = lo;
    const long long total_pad = (long long)B * T - fpre[B];
    ctl[3] = (int)(total_pad * blockIdx.x / gridDim.x);
    ctl[4] = (int)(total_pad * (blockIdx.x + 1) / gridDim.x);
  }
  __syncthreads();
  const int g1 = ctl[1], b_first = ctl[2];

  // ---- 0b. zero padding rows (sp_layers.py:88): an equal share per CTA ---------------------------
  {
    int q = ctl[3];
    const int q1 = ctl[4];
    if (q < q1) {
      int lo = 0, hi = B - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (mid * T - fpre[mid] <= q) lo = mid; else hi = mid - 1;
      }
      int b = lo;
      const bool vec = (D_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.feats) & 15) == 0;
      while (q < q1) {
        while ((b + 1) * T - fpre[b + 1] <= q) ++b;
        const int m_b = fpre[b + 1] - fpre[b];
        const int ofs = q - (b * T - fpre[b]);
        int nrows = (T - m_b) - ofs;
        nrows = nrows > q1 - q ? q1 - q : nrows;
        float* dst = p.feats + ((size_t)b * T + m_b + ofs) * D_out;
        if (vec) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          const int n4 = nrows * (D_out >> 2);
          for (int i = tid; i < n4; i += kPThreads) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (int i = tid; i < nrows * D_out; i += kPThreads) dst[i] = 0.f;
        }
        q += nrows;
      }
    }
  }

  // ---- per-warp helpers ---------------------------------------------------------------------------
  const char* wav_lo = static_cast<const char*>(p.wav);
  const size_t esz = p.sample_format == SPL_SAMPLES_F32 ? 4 : 2;
  const char* wav_hi = wav_lo + ((size_t)(B - 1) * p.wav_pitch + (size_t)p.wav_cols) * esz;
  int b_hint = b_first;
  auto fetch_group = [&]() {
    int g = 0;
    if (lane == 0) g = atomicAdd(ctl, 1);
    return __shfl_sync(0xffffffffu, g, 0);
  };
  struct Grp {
    int b, t0, n, head;
    bool bulk;
  };
  auto stage_group = [&](int g, Grp& q) {
    int b = b_hint;
    while (gpre[b + 1] <= g) ++b;
    b_hint = b;
    q.b = b;
    q.t0 = 2 * (g - gpre[b]);
    const int left = (fpre[b + 1] - fpre[b]) - q.t0;
    q.n = left < 2 ? left : 2;
    const int need = (q.n - 1) * S + Nw;
    q.bulk = false;
    q.head = 0;
    if (p.sample_format == SPL_SAMPLES_F32) {
      const char* src = wav_lo + ((size_t)b * p.wav_pitch + (size_t)q.t0 * S) * 4;
      const char* a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
      const int head = (int)((src - a0) >> 2);
      const uint32_t bytes = (uint32_t)(((head + need) * 4 + 15) & ~15);
      if (a0 >= wav_lo && a0 + bytes <= wav_hi) {
        q.bulk = true;
        q.head = head;
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(mybar, bytes);
          bulk_g2s(samp, a0, bytes, mybar);
        }
      }
    }
  };

  mbar_wait(bars, 0);  // tables have landed

  int stat_b = -1, stat_rows = 0;
  const bool want_stats = p.utt_stats != nullptr || p.global_stats != nullptr;
  auto flush_stats = [&]() {
    if (stat_b < 0) return;
    const int slot = stat_b - b_first;
    for (int c = lane; c < D_out; c += 32) {
      const double v1 = wstat[c], v2 = wstat[OP + c];
      wstat[c] = 0.0;
      wstat[OP + c] = 0.0;
      if (slot < kPStatUtts) {
        atomicAdd(cstat + (slot * 2 + 0) * OP + c, v1);
        atomicAdd(cstat + (slot * 2 + 1) * OP + c, v2);
      } else {
        if (p.utt_stats) {
          atomicAdd(p.utt_stats + ((size_t)stat_b * 2 + 0) * D_out + c, v1);
          atomicAdd(p.utt_stats + ((size_t)stat_b * 2 + 1) * D_out + c, v2);
        }
        if (p.global_stats) {
          atomicAdd(p.global_stats + c, v1);
          atomicAdd(p.global_stats + D_out + c, v2);
        }
      }
    }
    if (p.global_stats && lane == 0) atomicAdd(p.global_stats + 2 * D_out, (double)stat_rows);
    stat_rows = 0;
  };

  // lane roles
  const int k1 = lane & 15, typ = lane >> 4;
  const int partner = ((16 - k1) & 15) | ((typ ^ 1) << 4);
  const int st1_base = qaddr(lane);  // stage-1 write base of lane n2 = lane

  // ---- main loop: one pair (<= 2 frames of one utterance) per iteration ---------------------------
  uint32_t parity = 0;
  Grp cur, nxt;
  int g_cur = fetch_group();
  if (g_cur < g1) stage_group(g_cur, cur);
  while (g_cur < g1) {
    const int g_nxt = fetch_group();
    const int n = cur.n;
    const int need = (n - 1) * S + Nw;
    if (cur.bulk) {
      mbar_wait(mybar, parity);
      parity ^= 1;
    } else {  // scalar staging (int16 ingest, unaligned or boundary windows)
      const size_t gofs = (size_t)cur.b * p.wav_pitch + (size_t)cur.t0 * S;
      if (p.sample_format == SPL_SAMPLES_F32) {
        const float* src = static_cast<const float*>(p.wav) + gofs;
        for (int i = lane; i < need; i += 32) samp[i] = __ldg(src + i);
      } else {
        const int16_t* src = static_cast<const int16_t*>(p.wav) + gofs;
        for (int i = lane; i < need; i += 32) samp[i] = (float)__ldg(src + i);
      }
      __syncwarp();
    }
    const float* sbase = samp + cur.head;
    if (n < 2) {  // the absent second frame must read finite data
      for (int i = need + lane; i < S + Nw; i += 32) samp[cur.head + i] = 0.f;
      __syncwarp();
    }

    // ---- stage 1: radix-16 over n1 (lane = n2), twiddle, transpose through the exchange planes ----
    {
      float re[16], im[16];
      load_frame_p<NFFT, NW, NOISE>(re, p, sbase, win, energy + 0, lane, cur.b, cur.t0, true);
      load_frame_p<NFFT, NW, NOISE>(im, p, sbase + S, win, energy + 1, lane, cur.b, cur.t0 + 1, n > 1);
      __syncwarp();  // every lane has its samples in registers: the planes may overwrite the buffer
      fft_dif<16, F::NROW>(re, im);
      float* er = e0 + st1_base;
      float* ei = er + kQPlane;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float cs = tws[q * G::R2 + lane], sn = tws[NFFT + q * G::R2 + lane];
        const float vr = re[bitrev<16>(q)], vi = im[bitrev<16>(q)];
        er[q] = vr * cs + vi * sn;  // * (c - i s)
        ei[q] = vi * cs - vr * sn;
      }
    }
    __syncwarp();

    // ---- stage 2: lane = (k1, typ); two radix-8 networks over b for the quarters typ and typ + 2 ----
    float ar[8], ai[8], br[8], bi[8];
    {
      const float* ea = e0 + 16 * typ + k1;
      const float* eb = ea + 33;  // a + 2: 16 (a + 2) + 1
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        ar[b] = ea[kQB * b];
        ai[b] = ea[kQPlane + kQB * b];
        br[b] = eb[kQB * b];
        bi[b] = eb[kQPlane + kQB * b];
      }
    }
    __syncwarp();  // exchange planes dead: power rows / next samples may land
    if (g_nxt < g1) stage_group(g_nxt, nxt);
    fft_dif<8>(ar, ai);
    fft_dif<8>(br, bi);
    // twiddle by W_32^(a k2'), radix-4 across the quarters with the sibling lane:
    //   u = Y0 + Y2, v = Y0 - Y2 (type A);  s = Y1 + Y3, t = -i (Y1 - Y3) (type B)
    //   A: X[k2' + 0] = u + s, X[k2' + 16] = u - s ;  B: X[k2' + 8] = v + t, X[k2' + 24] = v - t
    // afterwards (ar, ai)[k2'] = bin k2'' = typ ("low"), (br, bi)[k2'] = bin k2'' = typ + 2 ("high")
    {
      const float4* t2 = tw2 + typ * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        constexpr int dummy = 0;
        (void)dummy;
        const int r = bitrev<8>(k);
        const float4 tw = t2[k];
        const float y0r = ar[r] * tw.x + ai[r] * tw.y, y0i = ai[r] * tw.x - ar[r] * tw.y;
        const float y1r = br[r] * tw.z + bi[r] * tw.w, y1i = bi[r] * tw.z - br[r] * tw.w;
        const float ur = y0r + y1r, ui = y0i + y1i, dr = y0r - y1r, di = y0i - y1i;
        const float sr = typ ? ur : dr, si = typ ? ui : di;
        const float rr = __shfl_xor_sync(0xffffffffu, sr, 16), ri = __shfl_xor_sync(0xffffffffu, si, 16);
        const float xr = typ ? rr : ur, xi = typ ? ri : ui;
        const float yr = typ ? di : rr, yi = typ ? -dr : ri;
        ar[r] = xr + yr;
        ai[r] = xi + yi;
        br[r] = xr - yr;
        bi[r] = xi - yi;
      }
    }
    // Hermitian partner + power.  Bin k = k1 + 16 k2' + 128 typ; its mirror N - k is the "high" value
    // index 7 - k2' of lane (16 - k1, other type); for k1 == 0 it is index 8 - k2' of the sibling
    // (the value shuffled one iteration earlier) and, for k2' == 0, the lane's own X[0] / X[384].
    {
      float* p0 = e0 + typ * kQPP + 128 * typ + k1;        // A lanes: row a first; B lanes: row b first
      float* p1 = e0 + (typ ^ 1) * kQPP + 128 * typ + k1;
      float pvr = typ ? br[0] : ar[0], pvi = typ ? bi[0] : ai[0];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float zr = ar[bitrev<8>(k)], zi = ai[bitrev<8>(k)];
        const float hr = __shfl_sync(0xffffffffu, br[bitrev<8>(7 - k)], partner);
        const float hi = __shfl_sync(0xffffffffu, bi[bitrev<8>(7 - k)], partner);
        const float qr = k1 == 0 ? pvr : hr, qi = k1 == 0 ? pvi : hi;
        pvr = hr;
        pvi = hi;
        const float fr = zr + qr, fi = zi - qi;  // 2 X_a[k]
        const float gr = zi + qi, gi = qr - zr;  // 2 X_b[k]
        const float pa = fr * fr + fi * fi, pb = gr * gr + gi * gi;  // the 1/4 is folded into the mel weights
        p0[16 * k] = typ ? pb : pa;
        p1[16 * k] = typ ? pa : pb;
      }
    }
    __syncwarp();

    // ---- mel: lane = (frame f, slice s), two filter streams per slice, E flat entries each ----
    {
      const int f = lane & 1, sl = lane >> 1;
      const float4* prow4 = reinterpret_cast<const float4*>(e0 + f * kQPP);
      float* orow = orows + f * OP + (p.use_energy ? 1 : 0);
      float acc0 = 0.f, acc1 = 0.f;
      const int E = p.tab.qE;
#pragma unroll 1
      for (int e = 0; e < E; ++e) {
        const uint2 d = mdesc[e * 16 + sl];
        const float4 w0 = melw[(2 * e) * 16 + sl], w1 = melw[(2 * e + 1) * 16 + sl];
        const float4 q0 = prow4[d.x & 63u], q1 = prow4[d.y & 63u];
        acc0 = fmaf(q0.x, w0.x, acc0);
        acc1 = fmaf(q1.x, w1.x, acc1);
        acc0 = fmaf(q0.y, w0.y, acc0);
        acc1 = fmaf(q1.y, w1.y, acc1);
        acc0 = fmaf(q0.z, w0.z, acc0);
        acc1 = fmaf(q1.z, w1.z, acc1);
        acc0 = fmaf(q0.w, w0.w, acc0);
        acc1 = fmaf(q1.w, w1.w, acc1);
        if (d.x & 0x10000u) {
          orow[(d.x >> 8) & 127u] = fast_log(fmaxf(acc0, kEps));  // kaldi_signal.py:540
          acc0 = 0.f;
        }
        if (d.y & 0x10000u) {
          orow[(d.y >> 8) & 127u] = fast_log(fmaxf(acc1, kEps));
          acc1 = 0.f;
        }
      }
      if (p.use_energy && lane < 2) orows[lane * OP] = energy[lane];
    }
    __syncwarp();

    // ---- store the pair's rows (contiguous in global memory) + column sums ----
    {
      float* out_g = p.feats + ((size_t)cur.b * T + cur.t0) * D_out;
      if ((D_out & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.feats) & 15) == 0)) {
        const int q = D_out >> 2;  // OP == D_out here
        const float4* src = reinterpret_cast<const float4*>(orows);
        float4* dst = reinterpret_cast<float4*>(out_g);
        for (int i = lane; i < n * q; i += 32) dst[i] = src[i];
      } else {
        for (int r = 0; r < n; ++r)
          for (int c = lane; c < D_out; c += 32) out_g[(size_t)r * D_out + c] = orows[r * OP + c];
      }
      if (want_stats) {
        if (cur.b != stat_b) {
          flush_stats();
          stat_b = cur.b;
        }
        stat_rows += n;
        for (int c = lane; c < D_out; c += 32) {
          double a1 = wstat[c], a2 = wstat[OP + c];
          for (int r = 0; r < n; ++r) {
            const double v = (double)orows[r * OP + c];
            a1 += v;
            a2 = fma(v, v, a2);
          }
          wstat[c] = a1;
          wstat[OP + c] = a2;
        }
      }
    }
    __syncwarp();
    g_cur = g_nxt;
    cur = nxt;
  }

  // ---- epilogue: merge the CTA's column sums ----
  if (want_stats) {
    flush_stats();
    __syncthreads();
    for (int sw = 0; sw < 2 * kPStatUtts; ++sw) {
      const int slot = sw >> 1, which = sw & 1, b = b_first + slot;
      if (b >= B) break;
      for (int c = tid; c < D_out; c += kPThreads) {
        const double v = cstat[sw * OP + c];
        if (v != 0.0) {
          if (p.utt_stats) atomicAdd(p.utt_stats + ((size_t)b * 2 + which) * D_out + c, v);
          if (p.global_stats) atomicAdd(p.global_stats + which * D_out + c, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int NW, bool NOISE>
static cudaError_t launch_q(const FbankParams& p, int num_ctas, cudaStream_t st) {
  const size_t smem = fbank_pair_smem_bytes(p.D_out, p.tab.qtab_words);
  static thread_local size_t configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || configured[dev] < smem) {
    cudaError_t e =
        cudaFuncSetAttribute(fbank_pair_kernel<NW, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = smem;
  }
  fbank_pair_kernel<NW, NOISE><<<num_ctas, kPThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fbank_pair(const FbankParams& p, bool with_noise, int num_ctas, cudaStream_t st) {
  if (p.Nw == 400) return with_noise ? launch_q<400, true>(p, num_ctas, st) : launch_q<400, false>(p, num_ctas, st);
  return with_noise ? launch_q<0, true>(p, num_ctas, st) : launch_q<0, false>(p, num_ctas, st);
}

}  // namespace spl
