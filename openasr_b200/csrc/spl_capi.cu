// C-ABI of the B200 speech front-end: handle management, argument validation, launches.
// See include/spl_capi.h for the contract and the reference interfaces each entry replaces.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "spl_internal.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return SPL_ERR_CUDA;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace

struct spl_handle {
  spl_config cfg;
  int device;
  int D_out;
  int num_sms;
  int kernel;  // 0 warp-pipelined (default), 1 simple one-tile-per-CTA (SPL_LEGACY_KERNEL=1), 2 persistent CTA tiles (=2)
  size_t smem_warp;
  size_t smem_warp16;
  size_t smem_pair;
  int ctas_per_sm;
  void* blob;  // single device allocation holding every table
  spl::Tables tab;
  size_t smem_bytes;
};

extern "C" {

int spl_abi_version(void) { return SPL_ABI_VERSION; }
const char* spl_last_error(void) { return g_err.c_str(); }
uint64_t spl_launch_count(void) { return g_launches.load(); }
int spl_feature_dim(const spl_handle* h) { return h ? h->D_out : 0; }

int spl_create(const spl_config* cfg, const float* window, const float* mel_dense, int device, spl_handle** out) {
  if (!cfg || !window || !mel_dense || !out) return fail(SPL_ERR_INVALID_ARG, "spl_create: null argument");
  if (cfg->abi_version != SPL_ABI_VERSION) return fail(SPL_ERR_INVALID_ARG, "spl_create: ABI version mismatch");
  const int S = cfg->window_shift, Nw = cfg->window_size, nfft = cfg->padded_size, D = cfg->num_mel_bins;
  if (nfft != 256 && nfft != 512)
    return fail(SPL_ERR_UNSUPPORTED,
                "spl_create: padded window must be 256 or 512 samples (sample rates ~5.2-20.4 kHz at 25 ms)");
  if (Nw < 2 || Nw > nfft || 2 * Nw <= nfft) return fail(SPL_ERR_INVALID_ARG, "spl_create: window_size/padded_size");
  if (S < 1 || S > Nw) return fail(SPL_ERR_INVALID_ARG, "spl_create: window_shift must be in [1, window_size]");
  if (D < 4 || D > spl::kMaxMel) return fail(SPL_ERR_INVALID_ARG, "spl_create: num_mel_bins must be in [4, 128]");
  if (!(cfg->preemph >= 0.f && cfg->preemph <= 1.f)) return fail(SPL_ERR_INVALID_ARG, "spl_create: preemph");

  // ---- sparse mel bank (weights pre-scaled by 1/4: the kernel accumulates 4*|X|^2) ----
  const int nb = nfft / 2;
  std::vector<int32_t> lo(D), cnt(D), off(D);
  std::vector<float> wts;
  for (int m = 0; m < D; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < nb; ++k)
      if (mel_dense[(size_t)m * nb + k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    lo[m] = first < 0 ? 0 : first;
    cnt[m] = first < 0 ? 0 : last - first + 1;
    off[m] = (int32_t)wts.size();
    for (int i = 0; i < cnt[m]; ++i) wts.push_back(0.25f * mel_dense[(size_t)m * nb + lo[m] + i]);
  }
  const int nnz = (int)wts.size();
  // filter groups for the 8 warps of the mel phase, balanced by (weights + per-filter overhead)
  int32_t grp[spl::kWarps + 1];
  {
    std::vector<int> cost(D);
    long total = 0;
    for (int m = 0; m < D; ++m) total += (cost[m] = cnt[m] + 6);
    int m = 0;
    long acc = 0;
    grp[0] = 0;
    for (int w = 1; w < spl::kWarps; ++w) {
      const long target = total * w / spl::kWarps;
      while (m < D && acc + cost[m] / 2 < target) acc += cost[m++];
      grp[w] = m;
    }
    grp[spl::kWarps] = D;
  }
  // ---- persistent-kernel table block (see spl_internal.cuh) ----
  const int npairs = (D + 1) / 2;
  std::vector<float> pw;          // pair weights
  std::vector<uint32_t> pdesc(npairs);
  std::vector<int> pcost(npairs);
  for (int i = 0; i < npairs; ++i) {
    int lo_[2], n4_[2];
    for (int s2 = 0; s2 < 2; ++s2) {
      const int m = 2 * i + s2;
      if (m < D && cnt[m] > 0) {
        lo_[s2] = lo[m] & ~3;
        n4_[s2] = (lo[m] + cnt[m] - lo_[s2] + 3) / 4;
      } else {
        lo_[s2] = 0;
        n4_[s2] = 0;
      }
    }
    const int n4p = n4_[0] > n4_[1] ? n4_[0] : n4_[1];
    for (int s2 = 0; s2 < 2; ++s2)
      if (lo_[s2] + 4 * n4p > nb) lo_[s2] = nb - 4 * n4p;  // keep the padded run inside the power row
    const uint32_t off8 = (uint32_t)(pw.size() / 8);
    for (int g = 0; g < n4p; ++g)
      for (int s2 = 0; s2 < 2; ++s2)
        for (int j = 0; j < 4; ++j) {
          const int m = 2 * i + s2, k = lo_[s2] + 4 * g + j;
          const bool in = m < D && k >= lo[m] && k < lo[m] + cnt[m];
          pw.push_back(in ? 0.25f * mel_dense[(size_t)m * nb + k] : 0.f);
        }
    pdesc[i] = (uint32_t)(lo_[0] >> 2) | ((uint32_t)(lo_[1] >> 2) << 6) | ((uint32_t)n4p << 12) | (off8 << 18) |
               ((2 * i + 1 < D) ? 0x80000000u : 0u);
    pcost[i] = 14 + 15 * n4p;
    if (n4p > 63 || off8 > 8191) return fail(SPL_ERR_UNSUPPORTED, "spl_create: mel bank too large for the descriptor");
  }
  int32_t pgrp[spl::kWarps + 1];
  {
    long total = 0;
    for (int i = 0; i < npairs; ++i) total += pcost[i];
    int i = 0;
    long acc = 0;
    pgrp[0] = 0;
    for (int w = 1; w < spl::kWarps; ++w) {
      const long target = total * w / spl::kWarps;
      while (i < npairs && acc + pcost[i] / 2 < target) acc += pcost[i++];
      pgrp[w] = i;
    }
    pgrp[spl::kWarps] = npairs;
  }
  // ---- warp-pipelined kernel table block (see spl_internal.cuh) ----
  const int nj = (npairs + 7) / 8;
  std::vector<float> ww;
  std::vector<uint32_t> wdesc(nj * 8, 0u), jinfo(nj, 0u);
  {
    int goff = 0;
    for (int j = 0; j < nj; ++j) {
      int lo_[8][2], n4j = 0;
      for (int s2 = 0; s2 < 8; ++s2)
        for (int hh = 0; hh < 2; ++hh) {
          const int m = 2 * (8 * j + s2) + hh;
          lo_[s2][hh] = 0;
          if (m < D && cnt[m] > 0) {
            lo_[s2][hh] = lo[m] & ~3;
            const int n4m = (lo[m] + cnt[m] - lo_[s2][hh] + 3) / 4;
            n4j = n4m > n4j ? n4m : n4j;
          }
        }
      for (int s2 = 0; s2 < 8; ++s2)
        for (int hh = 0; hh < 2; ++hh)
          if (lo_[s2][hh] + 4 * n4j > nb) lo_[s2][hh] = nb - 4 * n4j;
      for (int g = 0; g < n4j; ++g)
        for (int hh = 0; hh < 2; ++hh)
          for (int s2 = 0; s2 < 8; ++s2)
            for (int q = 0; q < 4; ++q) {
              const int m = 2 * (8 * j + s2) + hh, k = lo_[s2][hh] + 4 * g + q;
              const bool in = m < D && k >= lo[m] && k < lo[m] + cnt[m];
              ww.push_back(in ? 0.25f * mel_dense[(size_t)m * nb + k] : 0.f);
            }
      for (int s2 = 0; s2 < 8; ++s2) {
        const int m0 = 2 * (8 * j + s2);
        wdesc[8 * j + s2] = (uint32_t)(lo_[s2][0] >> 2) | ((uint32_t)(lo_[s2][1] >> 2) << 6) |
                            (m0 < D ? 0x40000000u : 0u) | (m0 + 1 < D ? 0x80000000u : 0u);
      }
      if (n4j > 255) return fail(SPL_ERR_UNSUPPORTED, "spl_create: mel bank too wide");
      jinfo[j] = (uint32_t)n4j | ((uint32_t)goff << 8);
      goff += n4j;
    }
  }
  // ---- pair-pipelined kernel table block (see spl_internal.cuh / fbank_pair.cu) ----
  std::vector<float> qw;
  std::vector<uint32_t> qdesc;
  int qE = 0;
  if (nfft == 512) {
    std::vector<int> lo4(D), n4(D), order(D);
    for (int m = 0; m < D; ++m) {
      lo4[m] = cnt[m] > 0 ? (lo[m] & ~3) : 0;
      n4[m] = cnt[m] > 0 ? (lo[m] + cnt[m] - lo4[m] + 3) / 4 : 1;  // empty filters still emit log(eps)
      if (lo4[m] + 4 * n4[m] > nb) lo4[m] = nb - 4 * n4[m];
      order[m] = m;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n4[a] > n4[b]; });
    std::vector<std::vector<int>> streams(32);
    std::vector<int> load(32, 0);
    for (int m : order) {  // longest first onto the least-loaded stream
      int best = 0;
      for (int s2 = 1; s2 < 32; ++s2)
        if (load[s2] < load[best]) best = s2;
      streams[best].push_back(m);
      load[best] += n4[m];
    }
    for (int s2 = 0; s2 < 32; ++s2) qE = load[s2] > qE ? load[s2] : qE;
    qw.assign((size_t)qE * 32 * 4, 0.f);
    qdesc.assign((size_t)qE * 32, 0u);
    for (int st2 = 0; st2 < 32; ++st2) {
      const int sl = st2 & 15, t = st2 >> 4;
      int e = 0;
      for (int m : streams[st2])
        for (int g = 0; g < n4[m]; ++g, ++e) {
          for (int q = 0; q < 4; ++q) {
            const int k = lo4[m] + 4 * g + q;
            const bool in = k >= lo[m] && k < lo[m] + cnt[m];
            qw[((size_t)(e * 2 + t) * 16 + sl) * 4 + q] = in ? 0.25f * mel_dense[(size_t)m * nb + k] : 0.f;
          }
          qdesc[((size_t)e * 16 + sl) * 2 + t] =
              (uint32_t)(lo4[m] / 4 + g) | ((uint32_t)m << 8) | (g == n4[m] - 1 ? 0x10000u : 0u);
        }
    }
  }
  std::vector<float> qtw2(64, 0.f);
  for (int typ = 0; typ < 2; ++typ)
    for (int k = 0; k < 8; ++k) {
      const double a0 = 2.0 * M_PI * (double)(typ * k) / 32.0, a1 = 2.0 * M_PI * (double)((typ + 2) * k) / 32.0;
      qtw2[(typ * 8 + k) * 4 + 0] = (float)std::cos(a0);
      qtw2[(typ * 8 + k) * 4 + 1] = (float)std::sin(a0);
      qtw2[(typ * 8 + k) * 4 + 2] = (float)std::cos(a1);
      qtw2[(typ * 8 + k) * 4 + 3] = (float)std::sin(a1);
    }
  // ---- tcgen05 DFT-as-GEMM tables (see spl_internal.cuh / fbank_tc.cu) ----
  const int half = nfft / 4, units = nfft / 32;
  auto tf32_rn = [](double v) {  // nearest TF32 (10 explicit mantissa bits), returned as float
    float f = (float)v;
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0xFFFu + ((u >> 13) & 1u);  // round to nearest even on the 13 dropped bits
    u &= 0xFFFFE000u;
    std::memcpy(&f, &u, 4);
    return f;
  };
  std::vector<float> tcb((size_t)units * 8 * half * 8);
  for (int u = 0; u < units; ++u)
    for (int blk = 0; blk < 4; ++blk)
      for (int n = 0; n < half; ++n)
        for (int e = 0; e < 8; ++e) {
          const int jp = 8 * u + e, j = 2 * jp + (blk & 1), k = n + 1;
          const double ang = 2.0 * M_PI * (double)((long)j * k % nfft) / (double)nfft;
          const double v = blk < 2 ? std::cos(ang) : std::sin(ang);
          const float hi = tf32_rn(v), lo = tf32_rn(v - (double)hi);
          // SWIZZLE_32B K-major tile of [half x 8]: byte offset r*32 + ((c>>2 ^ r>>2)&1)*16 + (c&3)*4
          const size_t off = (size_t)(n * 32 + ((((e >> 2) ^ (n >> 2)) & 1) << 4) + ((e & 3) << 2)) / 4;
          const size_t tile = (size_t)half * 8;
          tcb[((size_t)u * 8 + blk * 2 + 0) * tile + off] = hi;
          tcb[((size_t)u * 8 + blk * 2 + 1) * tile + off] = lo;
        }
  // mel segments over the split power layout: array 0 index n <-> bin n+1, array 1 index n <-> bin nb-1-n
  std::vector<float> sw;
  std::vector<uint32_t> sdesc;  // pairs (x, y)
  std::vector<int> scost, sfilt;
  for (int m = 0; m < D; ++m) {
    struct Run { int arr, i0, i1; };
    Run runs[2];
    int nr = 0;
    if (cnt[m] > 0) {
      const int b0 = lo[m], b1 = lo[m] + cnt[m] - 1;  // bins (>= 1, <= nb-1)
      if (b0 <= half) runs[nr++] = {0, (b0 < 1 ? 1 : b0) - 1, (b1 < half ? b1 : half) - 1};
      if (b1 > half) runs[nr++] = {1, nb - 1 - b1, nb - 1 - (b0 > half + 1 ? b0 : half + 1)};
    }
    if (nr == 0) runs[nr++] = {0, 0, -1};  // empty filter: one zero-length segment -> log(eps)
    for (int r = 0; r < nr; ++r) {
      int start4 = runs[r].i0 & ~3;
      int n4 = runs[r].i1 >= runs[r].i0 ? (runs[r].i1 - start4 + 4) / 4 : 0;
      if (start4 + 4 * n4 > half) start4 = half - 4 * n4;
      const uint32_t woff4 = (uint32_t)(sw.size() / 4);
      for (int i = 0; i < 4 * n4; ++i) {
        const int idx = start4 + i;
        const int bin = runs[r].arr == 0 ? idx + 1 : nb - 1 - idx;
        const bool in = idx >= runs[r].i0 && idx <= runs[r].i1 && bin >= lo[m] && bin < lo[m] + cnt[m];
        sw.push_back(in ? mel_dense[(size_t)m * nb + bin] : 0.f);
      }
      sdesc.push_back((uint32_t)(start4 >> 2) | ((uint32_t)n4 << 6) | ((uint32_t)runs[r].arr << 12) |
                      ((r == 0 ? 1u : 0u) << 13) | ((r == nr - 1 ? 1u : 0u) << 14) | ((uint32_t)m << 16));
      sdesc.push_back(woff4);
      scost.push_back(10 + 6 * n4 + (r == nr - 1 ? 8 : 0));
      sfilt.push_back(m);
    }
  }
  const int nseg = (int)scost.size();
  int32_t sgrp[5];
  {
    long total = 0;
    for (int i = 0; i < nseg; ++i) total += scost[i];
    int i = 0;
    long acc = 0;
    sgrp[0] = 0;
    for (int g = 1; g < 4; ++g) {
      const long target = total * g / 4;
      while (i < nseg && (acc + scost[i] / 2 < target || (i > 0 && sfilt[i] == sfilt[i - 1]))) acc += scost[i++];
      sgrp[g] = i;
    }
    sgrp[4] = nseg;
  }
  // DFT of the window (dither-mean correction): Wc[k] = sum w_j cos, Ws[k] = sum w_j sin, k = 0..nb-1
  std::vector<float> wc(nb), wsn(nb);
  for (int k = 0; k < nb; ++k) {
    double c = 0, s_ = 0;
    for (int j = 0; j < Nw; ++j) {
      const double ang = 2.0 * M_PI * (double)((long)j * k % nfft) / (double)nfft;
      c += (double)window[j] * std::cos(ang);
      s_ += (double)window[j] * std::sin(ang);
    }
    wc[k] = (float)c;
    wsn[k] = (float)s_;
  }

  // stage-1 twiddles W_N^{n2 k1}
  const int R2 = nfft / 16;
  std::vector<float> twr(R2 * 16), twi(R2 * 16);
  for (int n2 = 0; n2 < R2; ++n2)
    for (int k1 = 0; k1 < 16; ++k1) {
      const double a = 2.0 * M_PI * (double)(n2 * k1) / (double)nfft;
      twr[n2 * 16 + k1] = (float)std::cos(a);
      twi[n2 * 16 + k1] = (float)std::sin(a);
    }

  DeviceGuard guard(device);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_create: cudaSetDevice failed");
  // one blob: persistent table block (16-byte aligned, first) | window | tw_re | tw_im | mel_w | lo | cnt | off
  auto pad4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  const size_t pt_desc = pad4(pw.size()), pt_win = pt_desc + pad4(npairs), pt_tw = pt_win + pad4(Nw);
  const size_t pt_words = pt_tw + 2 * (size_t)nfft;
  const size_t wt_desc = pad4(ww.size()), wt_jinfo = wt_desc + pad4(wdesc.size()), wt_win = wt_jinfo + pad4(jinfo.size());
  const size_t wt_tw = wt_win + pad4(Nw), wt_words = wt_tw + 2 * (size_t)nfft;
  const size_t tt_desc = pad4(sw.size()), tt_win = tt_desc + pad4(sdesc.size()), tt_wc = tt_win + pad4(Nw);
  const size_t tt_ws = tt_wc + pad4(nb), tt_words = tt_ws + pad4(nb);
  const size_t tcb_words = tcb.size();  // multiple of 4
  const size_t qt_desc = pad4(qw.size()), qt_win = qt_desc + pad4(qdesc.size()), qt_tw = qt_win + pad4(Nw);
  const size_t qt_tw2 = qt_tw + 2 * (size_t)nfft, qt_words = qt_tw2 + 64;
  const size_t n_items = pt_words + wt_words + tt_words + tcb_words + qt_words + (size_t)Nw + 2 * (size_t)R2 * 16 +
                         (size_t)(nnz > 0 ? nnz : 1) + 3 * (size_t)D;
  std::vector<uint32_t> host(n_items, 0u);
  std::memcpy(host.data(), pw.data(), pw.size() * 4);
  std::memcpy(host.data() + pt_desc, pdesc.data(), npairs * 4);
  std::memcpy(host.data() + pt_win, window, Nw * 4);
  for (int n2 = 0; n2 < R2; ++n2)
    for (int k1 = 0; k1 < 16; ++k1) {  // transposed: conflict-free for lane = n2
      std::memcpy(host.data() + pt_tw + k1 * R2 + n2, &twr[n2 * 16 + k1], 4);
      std::memcpy(host.data() + pt_tw + nfft + k1 * R2 + n2, &twi[n2 * 16 + k1], 4);
    }
  {
    uint32_t* wt = host.data() + pt_words;  // pt_words is a multiple of 4: the block stays 16-byte aligned
    std::memcpy(wt, ww.data(), ww.size() * 4);
    std::memcpy(wt + wt_desc, wdesc.data(), wdesc.size() * 4);
    std::memcpy(wt + wt_jinfo, jinfo.data(), jinfo.size() * 4);
    std::memcpy(wt + wt_win, window, Nw * 4);
    std::memcpy(wt + wt_tw, host.data() + pt_tw, 2 * (size_t)nfft * 4);
  }
  {
    uint32_t* tt = host.data() + pt_words + wt_words;
    std::memcpy(tt, sw.data(), sw.size() * 4);
    std::memcpy(tt + tt_desc, sdesc.data(), sdesc.size() * 4);
    std::memcpy(tt + tt_win, window, Nw * 4);
    std::memcpy(tt + tt_wc, wc.data(), nb * 4);
    std::memcpy(tt + tt_ws, wsn.data(), nb * 4);
    std::memcpy(tt + tt_words, tcb.data(), tcb_words * 4);
  }
  {
    uint32_t* qt = host.data() + pt_words + wt_words + tt_words + tcb_words;  // all multiples of 4: 16-byte aligned
    if (!qw.empty()) std::memcpy(qt, qw.data(), qw.size() * 4);
    if (!qdesc.empty()) std::memcpy(qt + qt_desc, qdesc.data(), qdesc.size() * 4);
    std::memcpy(qt + qt_win, window, Nw * 4);
    std::memcpy(qt + qt_tw, host.data() + pt_tw, 2 * (size_t)nfft * 4);
    std::memcpy(qt + qt_tw2, qtw2.data(), 64 * 4);
  }
  size_t o = pt_words + wt_words + tt_words + tcb_words + qt_words;
  auto put = [&](const void* src, size_t n) {
    std::memcpy(host.data() + o, src, n * 4);
    size_t at = o;
    o += n;
    return at;
  };
  const size_t o_win = put(window, Nw), o_twr = put(twr.data(), twr.size()), o_twi = put(twi.data(), twi.size());
  const float zero = 0.f;
  const size_t o_w = nnz > 0 ? put(wts.data(), nnz) : put(&zero, 1);
  const size_t o_lo = put(lo.data(), D), o_cnt = put(cnt.data(), D), o_off = put(off.data(), D);

  void* blob = nullptr;
  cudaError_t e = cudaMalloc(&blob, n_items * 4);
  if (e != cudaSuccess) return fail_cuda(e, "spl_create: cudaMalloc");
  e = cudaMemcpy(blob, host.data(), n_items * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(blob);
    return fail_cuda(e, "spl_create: cudaMemcpy");
  }
  spl_handle* h = new spl_handle();
  h->cfg = *cfg;
  h->device = device;
  h->D_out = D + (cfg->use_energy ? 1 : 0);
  h->blob = blob;
  const float* fb = static_cast<const float*>(blob);
  const int32_t* ib = static_cast<const int32_t*>(blob);
  h->tab.window = fb + o_win;
  h->tab.tw_re = fb + o_twr;
  h->tab.tw_im = fb + o_twi;
  h->tab.mel_w = fb + o_w;
  h->tab.mel_lo = ib + o_lo;
  h->tab.mel_cnt = ib + o_cnt;
  h->tab.mel_off = ib + o_off;
  h->tab.mel_nnz = nnz;
  for (int w = 0; w <= spl::kWarps; ++w) h->tab.grp_beg[w] = grp[w];
  h->tab.ptab = fb;
  h->tab.ptab_words = (int32_t)pt_words;
  h->tab.pt_off_desc = (int32_t)pt_desc;
  h->tab.pt_off_win = (int32_t)pt_win;
  h->tab.pt_off_tw = (int32_t)pt_tw;
  h->tab.npairs = npairs;
  h->tab.wtab = fb + pt_words;
  h->tab.wtab_words = (int32_t)wt_words;
  h->tab.wt_off_desc = (int32_t)wt_desc;
  h->tab.wt_off_jinfo = (int32_t)wt_jinfo;
  h->tab.wt_off_win = (int32_t)wt_win;
  h->tab.wt_off_tw = (int32_t)wt_tw;
  h->tab.nj = nj;
  h->tab.tc_tab = fb + pt_words + wt_words;
  h->tab.tc_b = fb + pt_words + wt_words + tt_words;
  h->tab.tc_tab_words = (int32_t)tt_words;
  h->tab.tc_off_desc = (int32_t)tt_desc;
  h->tab.tc_off_win = (int32_t)tt_win;
  h->tab.tc_off_wc = (int32_t)tt_wc;
  h->tab.tc_off_ws = (int32_t)tt_ws;
  h->tab.tc_nseg = nseg;
  for (int g = 0; g < 5; ++g) h->tab.tc_sgrp_beg[g] = sgrp[g];
  h->tab.qtab = fb + pt_words + wt_words + tt_words + tcb_words;
  h->tab.qtab_words = (int32_t)qt_words;
  h->tab.qt_off_desc = (int32_t)qt_desc;
  h->tab.qt_off_win = (int32_t)qt_win;
  h->tab.qt_off_tw = (int32_t)qt_tw;
  h->tab.qt_off_tw2 = (int32_t)qt_tw2;
  h->tab.qE = qE;
  for (int w = 0; w <= spl::kWarps; ++w) h->tab.pgrp_beg[w] = pgrp[w];
  h->num_sms = 148;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  const char* leg = std::getenv("SPL_LEGACY_KERNEL");
  // 3: tcgen05 DFT-as-GEMM (experimental), 4: pair-pipelined (Nfft = 512)
  h->kernel = (leg && leg[0] >= '0' && leg[0] <= '4') ? leg[0] - '0' : 0;
  const char* cps = std::getenv("SPL_CTAS_PER_SM");  // experiment knob: persistent CTAs per SM (1 or 2)
  h->ctas_per_sm = (cps && cps[0] == '1') ? 1 : 2;
  h->smem_pair = nfft == 512 ? spl::fbank_pair_smem_bytes(h->D_out, (int)qt_words) : 0;
  if (h->kernel == 4 && (nfft != 512 || h->smem_pair > 113 * 1024 || S + Nw + 4 > 564)) h->kernel = 0;
  h->smem_warp = spl::fbank_warp_smem_bytes(nfft, S, Nw, h->D_out, (int)wt_words, 8);
  h->smem_warp16 = spl::fbank_warp_smem_bytes(nfft, S, Nw, h->D_out, (int)wt_words, 16);
  {  // the warp kernel stages a group's samples inside one pair's exchange planes
    const int pl = ((nfft / 16 * 17 + 15) / 32) * 32 + 16;
    if (h->kernel == 0 && (h->smem_warp > 113 * 1024 || 3 * S + Nw + 4 > 2 * pl)) h->kernel = 2;
  }
  h->smem_bytes = spl::fbank_smem_bytes(nfft, S, Nw, D, h->D_out, nnz);
  const size_t smem_p = spl::fbank_persistent_smem_bytes(nfft, S, Nw, h->D_out, (int)pt_words);
  if (smem_p > h->smem_bytes) h->smem_bytes = smem_p;
  if (h->smem_bytes > 113 * 1024) {  // two CTAs per SM must fit in 227 KB
    cudaFree(blob);
    delete h;
    return fail(SPL_ERR_UNSUPPORTED, "spl_create: configuration needs too much shared memory");
  }
  *out = h;
  return SPL_OK;
}

void spl_destroy(spl_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaFree(h->blob);
  delete h;
}

int spl_fbank_forward(spl_handle* h, const spl_fbank_args* a, void* stream) {
  if (!h || !a) return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: null argument");
  if (!a->wav || !a->wav_len || !a->feats) return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: null buffer");
  if (a->B < 1 || a->B > 65535 || a->T < 1) return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: B/T out of range");
  if (a->wav_cols < h->cfg.window_size || a->wav_pitch < a->wav_cols)
    return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: need window_size <= wav_cols <= wav_pitch");
  if (a->sample_format != SPL_SAMPLES_F32 && a->sample_format != SPL_SAMPLES_I16)
    return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: sample_format");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_fbank_forward: cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  spl::FbankParams p;
  p.S = h->cfg.window_shift;
  p.Nw = h->cfg.window_size;
  p.D = h->cfg.num_mel_bins;
  p.D_out = h->D_out;
  p.use_energy = h->cfg.use_energy;
  p.remove_dc = h->cfg.remove_dc;
  p.preemph = h->cfg.preemph;
  p.dither = h->cfg.dither;
  p.wav = a->wav;
  p.wav_pitch = a->wav_pitch;
  p.wav_cols = a->wav_cols;
  p.sample_format = a->sample_format;
  p.wav_len = a->wav_len;
  p.B = a->B;
  p.T = a->T;
  p.feats = a->feats;
  p.feat_len = a->feat_len;
  p.noise = a->noise;
  p.seed_lo = (uint32_t)(a->dither_seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(a->dither_seed >> 32);
  p.utt_stats = a->utt_stats;
  p.global_stats = a->global_stats;
  p.tab = h->tab;

  if (a->utt_stats) {
    cudaError_t e = cudaMemsetAsync(a->utt_stats, 0, sizeof(double) * 2 * (size_t)a->B * h->D_out, st);
    if (e != cudaSuccess) return fail_cuda(e, "spl_fbank_forward: cudaMemsetAsync");
  }
  const bool with_noise = h->cfg.dither != 0.f;
  cudaError_t e;
  if (h->kernel == 3 && a->B <= spl::kMaxPersistentB && a->sample_format == SPL_SAMPLES_F32 && h->cfg.dither == 0.f &&
      h->cfg.window_size * 2 > h->cfg.padded_size)
    e = spl::launch_fbank_tc(p, h->cfg.padded_size, with_noise, h->num_sms, st);
  else if (h->kernel == 4 && a->B <= spl::kMaxPersistentB)
    e = spl::launch_fbank_pair(p, with_noise, h->ctas_per_sm * h->num_sms, st);
  else if ((h->kernel == 0 || h->kernel == 3 || h->kernel == 4) && a->B <= spl::kMaxPersistentB)
    {
    // default: one 16-warp CTA per SM; SPL_CTAS_PER_SM=1 (throughput mode) or a table block too large for
    // 227 KB: 8-warp CTAs
    const bool wide = h->ctas_per_sm == 2 && h->smem_warp16 <= 227 * 1024;
    e = spl::launch_fbank_warp(p, h->cfg.padded_size, with_noise, wide ? 16 : 8,
                               wide ? h->num_sms : h->ctas_per_sm * h->num_sms, st);
  }
  else if (h->kernel == 2 && a->B <= spl::kMaxPersistentB)
    e = spl::launch_fbank_persistent(p, h->cfg.padded_size, with_noise, 2 * h->num_sms, st);
  else
    e = spl::launch_fbank(p, h->cfg.padded_size, with_noise, st);
  if (e != cudaSuccess) return fail_cuda(e, "spl_fbank_forward: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

int spl_post_inplace(spl_handle* h, const spl_post_args* a, void* stream) {
  if (!a || !a->feats || !a->feat_len) return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: null argument");
  if (a->B < 1 || a->B > 65535 || a->T < 1 || a->Dm < 1 || a->Dm > spl::kMaxDm)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: B/T/Dm out of range (Dm <= 160)");
  const int nmask = a->mask_params ? a->n_freq_masks + a->n_time_masks : 0;
  if (a->n_freq_masks < 0 || a->n_time_masks < 0 || nmask > spl::kMaxMasks)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: at most 32 masks per utterance");
  if (a->cmvn_mode == SPL_CMVN_UTTERANCE && !a->utt_stats)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: utterance CMVN needs utt_stats");
  if (a->cmvn_mode == SPL_CMVN_GLOBAL && (!a->global_mean || (a->norm_vars && !a->global_istd)))
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: global CMVN needs global_mean/global_istd");
  if (a->mask_params && a->n_time_masks > 0 && !a->utt_stats)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: time masks need utt_stats (time means)");
  if (a->cmvn_mode == SPL_CMVN_NONE && nmask == 0) return SPL_OK;  // nothing to do
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  DeviceGuard guard(h ? h->device : cur_dev);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_post_inplace: cudaSetDevice failed");
  spl::PostParams p;
  p.feats = a->feats;
  p.feat_len = a->feat_len;
  p.B = a->B;
  p.T = a->T;
  p.Dm = a->Dm;
  p.cmvn_mode = a->cmvn_mode;
  p.norm_vars = a->norm_vars;
  p.utt_stats = a->utt_stats;
  p.global_mean = a->global_mean;
  p.global_istd = a->global_istd;
  p.n_freq = a->mask_params ? a->n_freq_masks : 0;
  p.n_time = a->mask_params ? a->n_time_masks : 0;
  p.mask_params = a->mask_params;
  cudaError_t e = spl::launch_post(p, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "spl_post_inplace: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

int spl_column_stats(spl_handle* h, const float* feats, const int64_t* feat_len, int32_t B, int32_t T, int32_t Dm,
                     double* utt_stats, void* stream) {
  if (!feats || !feat_len || !utt_stats) return fail(SPL_ERR_INVALID_ARG, "spl_column_stats: null argument");
  if (B < 1 || B > 65535 || T < 1 || Dm < 1) return fail(SPL_ERR_INVALID_ARG, "spl_column_stats: B/T/Dm out of range");
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  DeviceGuard guard(h ? h->device : cur_dev);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_column_stats: cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(utt_stats, 0, sizeof(double) * 2 * (size_t)B * Dm, st);
  if (e != cudaSuccess) return fail_cuda(e, "spl_column_stats: cudaMemsetAsync");
  e = spl::launch_column_stats(feats, feat_len, B, T, Dm, utt_stats, st);
  if (e != cudaSuccess) return fail_cuda(e, "spl_column_stats: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

int spl_conv0_relu(spl_handle* h, const float* feats, int32_t B, int32_t T, int32_t D, const float* weight,
                   const float* bias, int32_t C, float* out, void* stream) {
  if (!feats || !weight || !out) return fail(SPL_ERR_INVALID_ARG, "spl_conv0_relu: null argument");
  if (B < 1 || B > 65535 || T < 3 || D < 3 || C < 1 || C > 64)
    return fail(SPL_ERR_INVALID_ARG, "spl_conv0_relu: need 1 <= B <= 65535, T >= 3, D >= 3, 1 <= C <= 64");
  if ((long long)((T - 3) / 2 + 1) * (D - 2) > 0x7fffffffLL - 4096)
    return fail(SPL_ERR_INVALID_ARG, "spl_conv0_relu: output plane too large");
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  DeviceGuard guard(h ? h->device : cur_dev);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_conv0_relu: cudaSetDevice failed");
  cudaError_t e = spl::launch_conv0_relu(feats, weight, bias, out, B, T, D, C, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "spl_conv0_relu: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

// Host helper: SpecAug rectangles from the uniforms, same float32 arithmetic and truncation as
// sp_layers.py:59-62 / :68-71 and Python slice semantics of x[b, s:s+w] (:64, :73).
int spl_specaug_rects(const float* uniforms, const int64_t* frames, int32_t B, int32_t T, int32_t V, int32_t n_freq,
                      float freq_width, int32_t n_time, float time_width, int32_t* out) {
  if (!uniforms || !frames || !out || B < 1 || n_freq < 0 || n_time < 0)
    return fail(SPL_ERR_INVALID_ARG, "spl_specaug_rects: bad argument");
  const int M = n_freq + n_time;
  for (int j = 0; j < M; ++j) {
    const float* uw = uniforms + (size_t)(2 * j) * B;
    const float* us = uw + B;
    const bool is_f = j < n_freq;
    const float wmax = is_f ? freq_width : time_width;
    const int64_t size = is_f ? V : T;
    for (int b = 0; b < B; ++b) {
      const int64_t width = (int64_t)(wmax * uw[b]);                              // (W * rand).long()
      const int64_t limit = is_f ? (int64_t)V : frames[b];
      const int64_t start = (int64_t)((float)(limit - width) * us[b]);            // ((V - fs).float() * rand).long()
      const int64_t end = start + width;
      int64_t s_ = start < 0 ? start + size : start, e_ = end < 0 ? end + size : end;
      s_ = s_ < 0 ? 0 : (s_ > size ? size : s_);
      e_ = e_ < 0 ? 0 : (e_ > size ? size : e_);
      out[((size_t)b * M + j) * 2 + 0] = (int32_t)s_;
      out[((size_t)b * M + j) * 2 + 1] = (int32_t)(e_ > s_ ? e_ : s_);
    }
  }
  return SPL_OK;
}

int spl_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t* status, void* stream) {
  if (!A || !B || !D || !status) return fail(SPL_ERR_INVALID_ARG, "spl_tc_selftest: null argument");
  if (N < 16 || N > 256 || (N & 15) || K < 8 || (K & 7) || (size_t)K * (128 + N) * 4 > 200 * 1024)
    return fail(SPL_ERR_INVALID_ARG, "spl_tc_selftest: need 16 <= N <= 256 (multiple of 16), K multiple of 8, tiles <= 200 KB");
  // K not a multiple of 32 selects the SWIZZLE_32B variant (one 8-wide tile per K step)
  cudaError_t e = (K & 31) ? spl::launch_tc_selftest_sw32(A, B, D, N, K, status, static_cast<cudaStream_t>(stream))
                           : spl::launch_tc_selftest(A, B, D, N, K, status, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "spl_tc_selftest: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

}  // extern "C"
