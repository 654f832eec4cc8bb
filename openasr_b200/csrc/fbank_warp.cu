// Kernel A (warp-pipelined, default): fused frame -> (dither) -> DC removal -> pre-emphasis -> window ->
// real FFT -> power -> sparse mel -> log, two CTAs of 8 *independent* warps per SM.
//
// Replaces src/third_party/kaldi_signal.py:163-211 + :510-552 and the pad/stack loop of
// src/blocks/sp_layers.py:81-91.  Execution structure, driven by the ncu profiles and the clock64
// timelines under profiles/:
//   * the unit of work is a GROUP of <= 4 consecutive frames of one utterance, processed end to
//     end by ONE warp: TMA bulk copy of the group's samples (fp32 or raw int16 PCM) -> two packed complex FFTs ->
//     power rows -> mel (lane = frame x filter slice) -> log -> coalesced store.  Inside the main
//     loop there is no __syncthreads: warps never wait for each other, only for their own TMA;
//   * all floating-point work runs on Blackwell's packed fp32 pipe (fft_c2.cuh: FADD2 / FMUL2 /
//     FFMA2): frames (t, t+1) are the (re, im) halves of one 64-bit register pair from the first
//     sample load on, so framing, butterflies, twiddles and power cost one instruction per pair;
//   * every CTA owns an equal, contiguous share of the batch's group list (the two CTAs that share an
//     SM get consecutive shares); warp w starts on group g0 + w, later groups come from a
//     shared-memory counter, so the load balance is at 4-frame granularity;
//   * the sample buffer aliases the second pair's exchange planes (dead until the stage-1 output
//     of that pair is written), so the next group's samples land behind FFT stage 2 / mel / store;
//   * per-utterance CMVN sums are kept per warp in fp64 shared-memory rows and merged per CTA after
//     the loop, one fp64 global atomic per (utterance, column) per CTA; the batch's zero-padding
//     rows are written by each warp after its last group.
#include "fbank_frame.cuh"

namespace spl {

#ifdef SPL_TRACE  // per-warp clock64 timeline (tools/trace_warp.py); never defined in the shipped library
__device__ unsigned long long g_trace[296 * 8 * 32];  // [CTA][warp][slot], CTA stride = warps per CTA
#define TR(slot)                                                                                   \
  do {                                                                                             \
    if (lane == 0 && (slot) < 32) g_trace[((size_t)blockIdx.x * (blockDim.x >> 5) + w) * 32 + (slot)] = clock64(); \
  } while (0)
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#else
#define TR(slot) do { } while (0)
#endif

// Resident warps per SM are register-file bound (128 regs x 16 warps).  Default: ONE CTA of 16 warps per SM,
// so that all 16 warps pull groups from one counter (two 8-warp CTAs per SM finish 20 % apart: the warps of the
// CTA that was placed first win the issue arbitration, profiles/r1_summary.md).  The 8-warp variant is the
// throughput mode (SPL_CTAS_PER_SM=1: two different launches co-resident on every SM).
constexpr int kWMaxWarps = 16;

struct WLayout {
  int pair;      // floats of one pair's exchange area (re plane + im plane)
  int p1;        // offset of pair 1's area (== 16 mod 32 away from pair 0: conflict-free stage-2 reads)
  int out_off;   // output rows (4 x op)
  int op;        // output-row pitch
  int en_off;    // 4 energy slots
  int st_off;    // this warp's running column sums of the current utterance: [2][op]
  int rw;        // region stride (multiple of 4)
  int off_tab, off_gpre, off_fpre, off_ppre, off_ctot, off_ctl, off_bar, off_warp, total;
};

__host__ __device__ inline WLayout make_wlayout(int nfft, int S, int Nw, int D_out, int wtab_words, int warps) {
  WLayout L;
  const int pl = nfft == 512 ? Geo<512>::PL : Geo<256>::PL;
  L.pair = 2 * pl;
  L.p1 = L.pair + 16;
  L.op = (D_out + 3) & ~3;
  L.out_off = L.p1 + L.pair;
  L.en_off = L.out_off + 4 * L.op;
  L.st_off = L.en_off + 4;
  L.st_off = (L.st_off + 1) & ~1;  // fp64 rows, 8-byte aligned
  L.rw = (L.st_off + 4 * L.op + 3) & ~3;
  L.off_tab = 0;
  L.off_gpre = wtab_words;
  L.off_fpre = L.off_gpre + kMaxPersistentB + 1;
  L.off_ppre = L.off_fpre + kMaxPersistentB + 1;
  L.off_ctot = L.off_ppre + kMaxPersistentB + 1;
  L.off_ctl = (L.off_ctot + 3 * (kMaxPersistentB / 32) + 3) & ~3;
  L.off_bar = (L.off_ctl + 8 + warps + 1) & ~1;  // ctl[8] + last utterance of every warp
  L.off_warp = (L.off_bar + 2 * (warps + 1) + 31) & ~31;
  L.total = L.off_warp + warps * L.rw;
  (void)S;
  (void)Nw;
  return L;
}

size_t fbank_warp_smem_bytes(int nfft, int S, int Nw, int D_out, int wtab_words, int warps) {
  return sizeof(float) * (size_t)make_wlayout(nfft, S, Nw, D_out, wtab_words, warps).total;
}

// ---------------------------------------------------------------------------------------------
// ST = element type of the input samples: float (int16-scaled) or int16_t PCM (row f1); both are staged
// by TMA as raw bytes and converted when the frames are loaded into registers.
template <int NFFT, int NW, bool NOISE, typename ST, int kWWarps>
__global__ void __launch_bounds__(kWWarps * 32, 16 / kWWarps) fbank_warp_kernel(const FbankParams p) {
  constexpr int kWThreads = kWWarps * 32;
  constexpr bool kDTab = NOISE && kWWarps == 16;  // the 16-warp CTA has room for the dither table
  using G = Geo<NFFT>;
  using F = FG<NFFT, NW>;
  extern __shared__ __align__(128) float smem[];
  const int S = p.S, Nw = F::kStatic ? NW : p.Nw, D_out = p.D_out;
  const int tab_words = p.tab.wtab_words + (kDTab ? kDitherTab : 0);
  const WLayout L = make_wlayout(NFFT, S, Nw, D_out, tab_words, kWWarps);
  static_assert(2 * G::PL >= 4 * G::PP, "power rows must fit in pair 0's exchange area");
  float* tab = smem + L.off_tab;
  const float4* melw = reinterpret_cast<const float4*>(tab);
  const uint32_t* pdesc = reinterpret_cast<const uint32_t*>(tab + p.tab.wt_off_desc);
  const uint32_t* jinfo = reinterpret_cast<const uint32_t*>(tab + p.tab.wt_off_jinfo);
  const float* win = tab + p.tab.wt_off_win;
  const float* tws = tab + p.tab.wt_off_tw;
  [[maybe_unused]] const float* dtab = tab + p.tab.wt_off_dith;
  int* gpre = reinterpret_cast<int*>(smem + L.off_gpre);  // gpre[b] = groups of utterances < b
  int* fpre = reinterpret_cast<int*>(smem + L.off_fpre);  // fpre[b] = frames of utterances < b
  int* ppre = reinterpret_cast<int*>(smem + L.off_ppre);  // ppre[b] = zero-padding rows of utterances < b
  int* ctl = reinterpret_cast<int*>(smem + L.off_ctl);    // [0] next group ... [8 + w] last utterance of warp w
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);  // [0] tables, [1 + w] samples of warp w

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int B = p.total_utts, OP = L.op;  // B: flattened utterances of all batches of the call
  TR(0);
#ifdef SPL_TRACE
  if (lane == 0) {
    g_trace[((size_t)blockIdx.x * (blockDim.x >> 5) + w) * 32 + 1] = gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_trace[((size_t)blockIdx.x * (blockDim.x >> 5) + w) * 32 + 26] = smid;
  }
  int tr_it = 0;
#endif
  float* wr = smem + L.off_warp + w * L.rw;  // this warp's region
  float* e0 = wr;                            // pair 0 exchange: re plane, im plane at + PL ; power rows alias it
  float* e1 = wr + L.p1;                     // pair 1 exchange ; the sample buffer aliases it
  ST* samp = reinterpret_cast<ST*>(e1);
  float* orows = wr + L.out_off;
  float* energy = wr + L.en_off;
  double* wstat = reinterpret_cast<double*>(wr + L.st_off);  // fp64: sum x^2 - mean^2 must survive std << mean
  uint64_t* mybar = bars + 1 + w;

  auto batch_of = [&](int u) -> int {  // batch of flattened utterance u (uniform loop over <= kMaxBatches descriptors)
    int k = 0;
#pragma unroll 1
    for (int i = 1; i < p.nb; ++i)
      if (u >= p.bd[i].u0) k = i;
    return k;
  };
  // ---- 0. tables (one TMA bulk copy), group prefix (warp 0), barriers -----------------------------
  if (tid == 0) {
    mbar_init(bars, 1);
    for (int i = 0; i < kWWarps; ++i) mbar_init(bars + 1 + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bars, (uint32_t)tab_words * 4u);
    bulk_g2s(tab, p.tab.wtab, (uint32_t)tab_words * 4u, bars);
  }
  // Prefix tables over the flattened utterances (groups, frames, zero-padding rows), in parallel: warp w scans the
  // 32-utterance chunks w, w + warps, ... (one round of global loads for the whole CTA), the chunk totals are scanned
  // by every warp after the barrier and added to its own entries.
  constexpr int kChunks = kMaxPersistentB / 32;
  static_assert(kChunks <= 32, "chunk totals are scanned by one warp");
  int* ctot = reinterpret_cast<int*>(smem + L.off_ctot);  // [3][kChunks]
  const int nchunk = (B + 31) >> 5;
  for (int c = w; c < nchunk; c += kWWarps) {
    const int bb = 32 * c + lane;
    int m = 0, pad = 0;
    if (bb < B) {
      const UBatch& bd = p.bd[batch_of(bb)];
      const long long n = bd.wav_len[bb - bd.u0];
      m = n >= Nw ? (int)(1 + (n - Nw) / S) : 0;  // kaldi_signal.py:90
      m = m > bd.T ? bd.T : m;
      pad = bd.T - m;
      if (blockIdx.x == 0 && bd.feat_len) bd.feat_len[bb - bd.u0] = m;
    }
    const int gcount = (m + 3) >> 2;
    int ginc = gcount, finc = m, pinc = pad;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vg = __shfl_up_sync(0xffffffffu, ginc, o);
      const int vf = __shfl_up_sync(0xffffffffu, finc, o);
      const int vp = __shfl_up_sync(0xffffffffu, pinc, o);
      if (lane >= o) {
        ginc += vg;
        finc += vf;
        pinc += vp;
      }
    }
    if (bb < B) {  // chunk-local exclusive prefixes
      gpre[bb] = ginc - gcount;
      fpre[bb] = finc - m;
      ppre[bb] = pinc - pad;
    }
    if (lane == 31) {
      ctot[c] = ginc;
      ctot[kChunks + c] = finc;
      ctot[2 * kChunks + c] = pinc;
    }
  }
  for (int i = lane; i < 2 * OP; i += 32) wstat[i] = 0.0;
  TR(22);
  __syncthreads();
  {
    const int og = lane < nchunk ? ctot[lane] : 0, of = lane < nchunk ? ctot[kChunks + lane] : 0,
              op = lane < nchunk ? ctot[2 * kChunks + lane] : 0;
    int ginc = og, finc = of, pinc = op;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vg = __shfl_up_sync(0xffffffffu, ginc, o);
      const int vf = __shfl_up_sync(0xffffffffu, finc, o);
      const int vp = __shfl_up_sync(0xffffffffu, pinc, o);
      if (lane >= o) {
        ginc += vg;
        finc += vf;
        pinc += vp;
      }
    }
    for (int c = w; c < nchunk; c += kWWarps) {  // exclusive chunk offsets of this warp's chunks
      const int ag = __shfl_sync(0xffffffffu, ginc - og, c), af = __shfl_sync(0xffffffffu, finc - of, c),
                ap = __shfl_sync(0xffffffffu, pinc - op, c);
      const int bb = 32 * c + lane;
      if (bb < B) {
        gpre[bb] += ag;
        fpre[bb] += af;
        ppre[bb] += ap;
      }
    }
    const int tg = __shfl_sync(0xffffffffu, ginc, nchunk - 1), tf = __shfl_sync(0xffffffffu, finc, nchunk - 1),
              tp = __shfl_sync(0xffffffffu, pinc, nchunk - 1);
    if (tid == 0) {
      gpre[B] = tg;
      fpre[B] = tf;
      ppre[B] = tp;
    }
  }
  __syncthreads();
  TR(23);
  // This CTA's contiguous share of the group list (thread 0) and of the zero-padding rows (thread 32).
  // floor(total * i / G) = q i + floor(r i / G) with total = q G + r: 32-bit divisions only (a 64-bit
  // division costs ~500 cycles of a single thread while the whole CTA waits).
  if (tid == 0 || tid == 32) {
    // CTAs c and c + G/2 are co-resident on one SM (two CTAs per SM, breadth-first placement): give
    // them consecutive shares so that every SM gets floor or ceil of the same 2/G of the work
    const unsigned G = gridDim.x;
    const unsigned bid = (G & 1u) ? blockIdx.x : 2u * (blockIdx.x % (G >> 1)) + blockIdx.x / (G >> 1);
    auto share = [&](long long total, unsigned i) -> int {
      if (total < 0x7fffffffLL && G < 0x10000u) {
        const unsigned t32 = (unsigned)total, q = t32 / G, r = t32 - q * G;
        return (int)(q * i + (r * i) / G);
      }
      return (int)(total * i / G);
    };
    if (tid == 0) {
      const int g0 = share(gpre[B], bid), g1 = share(gpre[B], bid + 1);
      int lo = 0, hi = B - 1;  // utterance of the first group
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (gpre[mid] <= g0) lo = mid; else hi = mid - 1;
      }
      ctl[0] = g0 + kWWarps;  // next group to hand out: warp w starts on group g0 + w
      ctl[1] = g1;
      ctl[2] = lo;
      ctl[5] = g0;
      ctl[6] = 0;  // valid rows this CTA contributed to the global statistics
    } else {
      const long long total_pad = ppre[B];
      ctl[3] = share(total_pad, bid);
      ctl[4] = share(total_pad, bid + 1);
    }
  }
  __syncthreads();
  TR(24);
  const int g1 = ctl[1], b_first = ctl[2];

  // ---- per-warp helpers ---------------------------------------------------------------------------
  constexpr int ES = (int)sizeof(ST);
  int b_hint = b_first, k_hint = 0;  // groups (hence utterances and batches) only move forward within a warp
  auto fetch_group = [&]() {  // warp-uniform: next group id of this CTA (or >= g1)
    int g = 0;
    if (lane == 0) g = atomicAdd(ctl, 1);
    return __shfl_sync(0xffffffffu, g, 0);
  };
  struct Grp {
    int b, t0, n, head, k;  // flattened utterance, first frame, frames, staged head elements, batch
    bool bulk;
  };
  // Locate group g and start staging its samples into `samp`.
  auto stage_group = [&](int g, Grp& q) {
    int b = b_hint;
    while (gpre[b + 1] <= g) ++b;
    b_hint = b;
    q.b = b;
    while (k_hint + 1 < p.nb && b >= p.bd[k_hint + 1].u0) ++k_hint;
    q.k = k_hint;
    const UBatch& bd = p.bd[q.k];
    const char* wav_lo = static_cast<const char*>(bd.wav);
    const char* wav_hi = wav_lo + ((size_t)(bd.B - 1) * bd.wav_pitch + (size_t)bd.wav_cols) * ES;
    q.t0 = 4 * (g - gpre[b]);
    const int left = (fpre[b + 1] - fpre[b]) - q.t0;
    q.n = left < 4 ? left : 4;
    const int need = (q.n - 1) * S + Nw;
    q.bulk = false;
    q.head = 0;
    {
      const char* src = wav_lo + ((size_t)(b - bd.u0) * bd.wav_pitch + (size_t)q.t0 * S) * ES;
      const char* a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
      const int head = (int)((src - a0) / ES);  // 0..3 floats or 0..7 int16 before the first sample
      const uint32_t bytes = (uint32_t)(((head + need) * ES + 15) & ~15);
      if (a0 >= wav_lo && a0 + bytes <= wav_hi) {
        q.bulk = true;
        q.head = head;
        if (lane == 0) {
          fence_proxy_async();  // earlier generic-proxy accesses of this area precede the async writes
          mbar_expect_tx(mybar, bytes);
          bulk_g2s(samp, a0, bytes, mybar);
        }
      }
    }
  };

  // per-utterance column sums (CMVN / SpecAug time means): running fp64 sums of this warp's current
  // utterance live in its own shared-memory rows; the CTA merges them once, after the loop
  int stat_b = -1, stat_rows = 0;
  bool want_stats = p.global_stats != nullptr;
  for (int k = 0; k < p.nb; ++k) want_stats = want_stats || p.bd[k].utt_stats != nullptr;
  auto flush_stats = [&]() {  // utterance change inside the loop (rare: a CTA's share spans 1-2 utterances)
    if (stat_b < 0) return;
    for (int c = lane; c < D_out; c += 32) {
      const double v1 = wstat[c], v2 = wstat[OP + c];
      wstat[c] = 0.0;
      wstat[OP + c] = 0.0;
      const UBatch& sbd = p.bd[batch_of(stat_b)];
      if (sbd.utt_stats) {
        atomicAdd(sbd.utt_stats + ((size_t)(stat_b - sbd.u0) * 2 + 0) * D_out + c, v1);
        atomicAdd(sbd.utt_stats + ((size_t)(stat_b - sbd.u0) * 2 + 1) * D_out + c, v2);
      }
      if (p.global_stats) {
        atomicAdd(p.global_stats + c, v1);
        atomicAdd(p.global_stats + D_out + c, v2);
      }
    }
    if (p.global_stats && lane == 0) atomicAdd(ctl + 6, stat_rows);
    stat_rows = 0;
  };

  // ---- main loop: one group (<= 4 frames of one utterance) per iteration, no block-wide barriers ----
  uint32_t parity = 0;
  Grp cur, nxt;
  int g_cur = ctl[5] + w;
  if (g_cur < g1) stage_group(g_cur, cur);
  TR(25);

  TR(2);
  mbar_wait(bars, 0);  // tables have landed (the first group's samples are already in flight)
  TR(3);
  while (g_cur < g1) {
    const int g_nxt = fetch_group();
    const int n = cur.n;
    const int need = (n - 1) * S + Nw;
    TR(4 + 6 * tr_it);
    if (cur.bulk) {
      mbar_wait(mybar, parity);
      parity ^= 1;
      TR(5 + 6 * tr_it);
    } else {  // warp-load staging (windows whose 16-byte envelope would leave the batch buffer)
      const UBatch& cbd = p.bd[cur.k];
      const ST* src = static_cast<const ST*>(cbd.wav) + (size_t)(cur.b - cbd.u0) * cbd.wav_pitch + (size_t)cur.t0 * S;
      for (int i = lane; i < need; i += 32) samp[i] = __ldg(src + i);
      __syncwarp();
    }
    const ST* sbase = samp + cur.head;
    const UBatch& gbd = p.bd[cur.k];
    [[maybe_unused]] const float* nz_utt = gbd.noise ? gbd.noise + (size_t)(cur.b - gbd.u0) * gbd.T * Nw : nullptr;
    if (n < 4) {  // invalid frames must read finite data: zero everything past the valid span
      for (int i = need + lane; i < 3 * S + Nw; i += 32) samp[cur.head + i] = (ST)0;
      __syncwarp();
    }

    // ---- stage 1: radix-16 over n1 (lane = n2), twiddle, transpose through the exchange planes.
    // Frames (t0, t0 + 1) and (t0 + 2, t0 + 3) are the (re, im) halves of two packed complex FFTs;
    // all arithmetic runs on the fp32x2 pipe (fft_c2.cuh) ----
    if constexpr (NFFT == 512) {
      const int n2 = lane;
      float twr[16], twi[16];
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        twr[k1] = tws[k1 * G::R2 + n2];
        twi[k1] = tws[NFFT + k1 * G::R2 + n2];
      }
      c2 z0[16], z1[16];
      load_frame_pair<NFFT, NW, NOISE, ST, kDTab>(z0, p, sbase, sbase + S, win, energy + 0, n2, cur.b, cur.t0,
                                                  cur.t0 + 1, true, n > 1, nz_utt, dtab);
      fft_dif_c<16, F::NROW>(z0);
      {
        float* er = e0 + n2 * G::EP;
        float* ei = er + G::PL;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          const c2 o = c2_mulw(z0[bitrev<16>(k1)], twr[k1], twi[k1]);  // * (c - i s)
          er[k1] = c2_re(o);
          ei[k1] = c2_im(o);
        }
      }
      load_frame_pair<NFFT, NW, NOISE, ST, kDTab>(z1, p, sbase + 2 * S, sbase + 3 * S, win, energy + 2, n2, cur.b,
                                                  cur.t0 + 2, cur.t0 + 3, n > 2, n > 3, nz_utt, dtab);
      __syncwarp();  // every lane has its samples in registers: pair 1's planes may overwrite the buffer
      fft_dif_c<16, F::NROW>(z1);
      {
        float* er = e1 + n2 * G::EP;
        float* ei = er + G::PL;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
          const c2 o = c2_mulw(z1[bitrev<16>(k1)], twr[k1], twi[k1]);
          er[k1] = c2_re(o);
          ei[k1] = c2_im(o);
        }
      }
    } else {
      const int pr = lane >> 4, n2 = lane & 15;
      const int fa = 2 * pr;
      c2 z[16];
      load_frame_pair<NFFT, NW, NOISE, ST, kDTab>(z, p, sbase + fa * S, sbase + (fa + 1) * S, win, energy + fa, n2, cur.b,
                                                  cur.t0 + fa, cur.t0 + fa + 1, fa < n, fa + 1 < n, nz_utt, dtab);
      __syncwarp();  // samples are in registers; pair 1's planes alias the buffer
      fft_dif_c<16, F::NROW>(z);
      float* er = (pr ? e1 : e0) + n2 * G::EP;
      float* ei = er + G::PL;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const c2 o = c2_mulw(z[bitrev<16>(k1)], tws[k1 * G::R2 + n2], tws[NFFT + k1 * G::R2 + n2]);
        er[k1] = c2_re(o);
        ei[k1] = c2_im(o);
      }
    }
    __syncwarp();
    TR(6 + 6 * tr_it);

    // ---- stage 2: lane = (pair, k1), registers = n2 ----
    const int pr = lane >> 4, k1 = lane & 15;
    c2 x[G::R2];
    {
      const float* er = (pr ? e1 : e0) + k1;
      const float* ei = er + G::PL;
#pragma unroll
      for (int n2 = 0; n2 < G::R2; ++n2) x[n2] = c2_make(er[n2 * G::EP], ei[n2 * G::EP]);
    }
    __syncwarp();  // exchange planes dead: pair 0's area becomes the power rows, pair 1's the next samples
    if (g_nxt < g1) stage_group(g_nxt, nxt);
    fft_dif_c<G::R2>(x);

    // Hermitian partner + power: Z[k1 + 16 k2] sits at register bitrev(k2).  With q = Z[N - k]:
    //   2 X_a = (zr + qr, zi - qi),  2 X_b = (zi + qi, qr - zr)   (frames a / b of the pair)
    //   S = Z + q = (ar, br),  D = Z - q = (-bi, ai)  ->  (4|X_a|^2, 4|X_b|^2) = S*S + swap(D)*swap(D)
    {
      const int partner = (lane & 16) | ((16 - k1) & 15);
      float* pa = e0 + (2 * pr) * G::PP + k1;
      float* pb = pa + G::PP;
#pragma unroll
      for (int k2 = 0; k2 < G::H; ++k2) {
        const c2 zv = x[bitrev<G::R2>(k2)];
        const c2 qs = x[bitrev<G::R2>(G::R2 - 1 - k2)];
        c2 q = c2_make(__shfl_sync(0xffffffffu, c2_re(qs), partner), __shfl_sync(0xffffffffu, c2_im(qs), partner));
        if (k1 == 0) q = x[bitrev<G::R2>((G::R2 - k2) & (G::R2 - 1))];
        const c2 sv = zv + q, dv = c2_swap(zv - q);
        const c2 pw = c2_fma(dv, dv, c2_mul(sv, sv));  // the 1/4 is folded into the mel weights
        pa[16 * k2] = c2_re(pw);
        pb[16 * k2] = c2_im(pw);
      }
    }
    __syncwarp();
    TR(7 + 6 * tr_it);

    // ---- mel: lane = (frame f, slice s); iteration j handles filter pairs 8 j + s ----
    {
      const int f = lane & 3, sl = lane >> 2;
      const float4* prow4 = reinterpret_cast<const float4*>(e0 + f * G::PP);
      float* orow = orows + f * OP + (p.use_energy ? 1 : 0);
      for (int j = 0; j < p.tab.nj; ++j) {
        const uint32_t ji = jinfo[j];
        const int n4 = ji & 255u;
        const float4* wv = melw + (size_t)(ji >> 8) * 16 + sl;
        const uint32_t dsc = pdesc[8 * j + sl];
        const float4* pa4 = prow4 + (dsc & 63u);
        const float4* pb4 = prow4 + ((dsc >> 6) & 63u);
        c2 accA = c2_splat(0.f), accB = c2_splat(0.f);  // (x + z, y + w) partial sums of the 4-bin groups
        auto mac = [&](int g) {
          const ulonglong2 pa = reinterpret_cast<const ulonglong2*>(pa4)[g], pb = reinterpret_cast<const ulonglong2*>(pb4)[g];
          const ulonglong2 wa = reinterpret_cast<const ulonglong2*>(wv)[16 * g],
                           wb = reinterpret_cast<const ulonglong2*>(wv)[16 * g + 8];
          accA = c2_fma(c2{pa.x}, c2{wa.x}, accA);
          accB = c2_fma(c2{pb.x}, c2{wb.x}, accB);
          accA = c2_fma(c2{pa.y}, c2{wa.y}, accA);
          accB = c2_fma(c2{pb.y}, c2{wb.y}, accB);
        };
        int g = 0;
#pragma unroll 1
        for (; g + 2 <= n4; g += 2) {  // two 4-bin groups per trip: half the pointer / branch overhead
          mac(g);
          mac(g + 1);
        }
        if (g < n4) mac(g);
        const int m0 = 2 * (8 * j + sl);
        if (dsc & 0x40000000u) orow[m0] = fast_log(fmaxf(c2_re(accA) + c2_im(accA), kEps));  // kaldi_signal.py:540
        if (dsc & 0x80000000u) orow[m0 + 1] = fast_log(fmaxf(c2_re(accB) + c2_im(accB), kEps));
      }
      if (p.use_energy && lane < 4) orows[lane * OP] = energy[lane];
    }
    __syncwarp();
    TR(8 + 6 * tr_it);

    // ---- store the group's rows (contiguous in global memory) + column sums ----
    {
      float* out_g = gbd.feats + ((size_t)(cur.b - gbd.u0) * gbd.T + cur.t0) * D_out;
      if ((D_out & 3) == 0 && ((reinterpret_cast<uintptr_t>(gbd.feats) & 15) == 0)) {
        const int q = D_out >> 2;  // OP == D_out here: the rows are contiguous in shared memory too
        const float4* src = reinterpret_cast<const float4*>(orows);
        float4* dst = reinterpret_cast<float4*>(out_g);
        const int tot = n * q;
        for (int i0 = lane; i0 < tot; i0 += 128) {  // four independent 16-byte copies in flight per lane
          float4 tmp[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < tot) tmp[u] = src[i0 + 32 * u];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + 32 * u < tot) dst[i0 + 32 * u] = tmp[u];
        }
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
        for (int c0 = lane; c0 < D_out; c0 += 96) {  // three columns per lane: independent fp64 chains
          double a1[3], a2[3];
          float v[3][4];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int c = c0 + 32 * k;
            const bool cv = c < D_out;
            a1[k] = cv ? wstat[c] : 0.0;
            a2[k] = cv ? wstat[OP + c] : 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) v[k][r] = (cv && r < n) ? orows[r * OP + c] : 0.f;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const double d = (double)v[k][r];
              a1[k] += d;
              a2[k] = fma(d, d, a2[k]);
            }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int c = c0 + 32 * k;
            if (c < D_out) {
              wstat[c] = a1[k];
              wstat[OP + c] = a2[k];
            }
          }
        }
      }
    }
    __syncwarp();  // output rows / power rows are rewritten by the next iteration
    TR(9 + 6 * tr_it);
#ifdef SPL_TRACE
    ++tr_it;
#endif
    g_cur = g_nxt;
    cur = nxt;
  }

  TR(30);
  // ---- zero padding rows (sp_layers.py:88): the CTA's equal share of the batch's padded rows, one
  // eighth per warp, written after the warp's last group -- off the critical path, since the warps
  // that ran out of groups early would otherwise idle at the final barrier ----
  {
    const int q0 = ctl[3], qn = ctl[4] - q0;
    int q = q0 + (int)((long long)qn * w / kWWarps);
    const int q1 = q0 + (int)((long long)qn * (w + 1) / kWWarps);
    if (q < q1) {
      int lo = 0, hi = B - 1;  // largest b with ppre[b] <= q
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ppre[mid] <= q) lo = mid; else hi = mid - 1;
      }
      int b = lo;
      while (q < q1) {
        while (b + 1 < B && ppre[b + 1] <= q) ++b;
        const UBatch& zbd = p.bd[batch_of(b)];
        const int T = zbd.T;
        const bool vec = (D_out & 3) == 0 && (reinterpret_cast<uintptr_t>(zbd.feats) & 15) == 0;
        const int m_b = fpre[b + 1] - fpre[b];
        const int ofs = q - ppre[b];
        int nrows = (T - m_b) - ofs;
        nrows = nrows > q1 - q ? q1 - q : nrows;
        float* dst = zbd.feats + ((size_t)(b - zbd.u0) * T + m_b + ofs) * D_out;
        if (vec) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          const int n4 = nrows * (D_out >> 2);
          for (int i = lane; i < n4; i += 32) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (int i = lane; i < nrows * D_out; i += 32) dst[i] = 0.f;
        }
        q += nrows;
      }
    }
  }
  // ---- epilogue: the warps' running sums (one utterance each) are merged by column, without
  // shared-memory atomics: thread i owns entry i of the [2][OP] rows and walks the 8 warps ----
  if (want_stats) {
    if (lane == 0) {
      ctl[8 + w] = stat_b;
      if (p.global_stats && stat_rows) atomicAdd(ctl + 6, stat_rows);
    }
    __syncthreads();
    int b_lo = 0x7fffffff, b_hi = -1;
#pragma unroll
    for (int ww = 0; ww < kWWarps; ++ww) {
      const int bw = ctl[8 + ww];
      if (bw >= 0) {
        b_lo = bw < b_lo ? bw : b_lo;
        b_hi = bw > b_hi ? bw : b_hi;
      }
    }
    for (int i = tid; i < 2 * OP; i += kWThreads) {
      const int which = i >= OP ? 1 : 0, c = i - which * OP;
      if (c >= D_out) continue;
      for (int b = b_lo; b <= b_hi; ++b) {
        double v = 0.0;
        bool any = false;
#pragma unroll
        for (int ww = 0; ww < kWWarps; ++ww)
          if (ctl[8 + ww] == b) {
            v += reinterpret_cast<const double*>(smem + L.off_warp + ww * L.rw + L.st_off)[i];
            any = true;
          }
        if (any) {
          const UBatch& mbd = p.bd[batch_of(b)];
          if (mbd.utt_stats) atomicAdd(mbd.utt_stats + ((size_t)(b - mbd.u0) * 2 + which) * D_out + c, v);
          if (p.global_stats) atomicAdd(p.global_stats + which * D_out + c, v);
        }
      }
    }
    if (p.global_stats && tid == 0 && ctl[6]) atomicAdd(p.global_stats + 2 * D_out, (double)ctl[6]);
  }
  TR(31);
}

#ifdef SPL_TRACE
extern "C" __attribute__((visibility("default"))) int spl_debug_trace(unsigned long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * (size_t)n);
}
#endif

// ---------------------------------------------------------------------------------------------
template <int NFFT, int NW, bool NOISE, typename ST, int WARPS>
static cudaError_t launch_w(const FbankParams& p, int num_ctas, cudaStream_t st) {
  const size_t smem = fbank_warp_smem_bytes(NFFT, p.S, NW > 0 ? NW : p.Nw, p.D_out,
                                            p.tab.wtab_words + (NOISE && WARPS == 16 ? kDitherTab : 0), WARPS);
  static thread_local size_t configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(fbank_warp_kernel<NFFT, NW, NOISE, ST, WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = smem;
  }
  fbank_warp_kernel<NFFT, NW, NOISE, ST, WARPS><<<num_ctas, WARPS * 32, smem, st>>>(p);
  return cudaGetLastError();
}

template <int NFFT, int NW, bool NOISE>
static cudaError_t launch_fmt(const FbankParams& p, int warps, int num_ctas, cudaStream_t st) {
  if (warps == 16)
    return p.sample_format == SPL_SAMPLES_F32 ? launch_w<NFFT, NW, NOISE, float, 16>(p, num_ctas, st)
                                              : launch_w<NFFT, NW, NOISE, int16_t, 16>(p, num_ctas, st);
  return p.sample_format == SPL_SAMPLES_F32 ? launch_w<NFFT, NW, NOISE, float, 8>(p, num_ctas, st)
                                            : launch_w<NFFT, NW, NOISE, int16_t, 8>(p, num_ctas, st);
}

// warps: 16 (one CTA per SM, default) or 8 (throughput mode / shared-memory fallback)
cudaError_t launch_fbank_warp(const FbankParams& p, int nfft, bool with_noise, int warps, int num_ctas, cudaStream_t st) {
  if (nfft == 512) {
    if (p.Nw == 400)
      return with_noise ? launch_fmt<512, 400, true>(p, warps, num_ctas, st) : launch_fmt<512, 400, false>(p, warps, num_ctas, st);
    return with_noise ? launch_fmt<512, 0, true>(p, warps, num_ctas, st) : launch_fmt<512, 0, false>(p, warps, num_ctas, st);
  }
  if (p.Nw == 200)
    return with_noise ? launch_fmt<256, 200, true>(p, warps, num_ctas, st) : launch_fmt<256, 200, false>(p, warps, num_ctas, st);
  return with_noise ? launch_fmt<256, 0, true>(p, warps, num_ctas, st) : launch_fmt<256, 0, false>(p, warps, num_ctas, st);
}

}  // namespace spl
