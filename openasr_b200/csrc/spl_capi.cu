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

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler injects itself

#include "spl_internal.cuh"

namespace {

struct NvtxRange {  // shows kernel A / kernel B as named ranges in nsys / ncu timelines
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

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

enum Engine { ENGINE_UMMA = 0, ENGINE_FFT = 1, ENGINE_SIMPLE = 2 };

struct spl_handle {
  spl_config cfg;
  int device;
  int D_out;
  int num_sms;
  int engine;        // preferred kernel-A engine (SPL_ENGINE=fft|umma|simple; default fft)
  bool fft_ok;       // the warp-pipelined FFT kernel fits this configuration
  size_t smem_warp;
  size_t smem_warp16;
  int ctas_per_sm;
  void* blob;        // single device allocation holding every table
  int32_t* status;   // device word set by a kernel whose barrier wait timed out (spl_debug_status)
  float* debug_acc;  // SPL_UMMA_DEBUG=1: raw accumulators of the first tile (spl_debug_umma_acc)
  spl::Tables tab;
  size_t smem_bytes;
  // tcgen05 engine, per sample format (0 fp32, 1 int16): tables present, device pointers
  bool umma_ok[2];
  const uint8_t* umma_tw[2];
  const float* umma_tab[2];
  int umma_tab_bytes[2], umma_off_melw[2], umma_off_melc[2], umma_nflush, umma_nparts, umma_part_f0[4], umma_part_s0[5];
  int umma_ctas;     // grid of the tcgen05 engine (SM count; SPL_UMMA_CTAS overrides, diagnostics)
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
  const int npairs = (D + 1) / 2;
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
  // one blob: warp-kernel table block (16-byte aligned, first) | window | tw_re | tw_im | mel_w | lo | cnt | off |
  //           umma tables and twiddle images of both sample formats (16-byte aligned)
  auto pad4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  const size_t wt_desc = pad4(ww.size()), wt_jinfo = wt_desc + pad4(wdesc.size()), wt_win = wt_jinfo + pad4(jinfo.size());
  const size_t wt_tw = wt_win + pad4(Nw), wt_words = wt_tw + 2 * (size_t)nfft;
  const size_t wt_dith = wt_words, wt_end = wt_dith + spl::kDitherTab;  // dither table: staged only by the 16-warp variant
  spl::UmmaHostTables ut;
  if (!cfg->use_energy) spl::build_umma_tables(nfft, Nw, D, window, mel_dense, ut);
  size_t n_items = wt_end + (size_t)Nw + 2 * (size_t)R2 * 16 + (size_t)(nnz > 0 ? nnz : 1) + 3 * (size_t)D;
  n_items = pad4(n_items);
  size_t o_utab[2] = {0, 0}, o_utw[2] = {0, 0};
  for (int f = 0; f < 2; ++f) {
    o_utab[f] = n_items;
    n_items += pad4(ut.tab[f].size());
    o_utw[f] = n_items;
    n_items += pad4(ut.twiddles[f].size() / 4);
  }
  std::vector<uint32_t> host(n_items, 0u);
  std::vector<float> twt(2 * (size_t)nfft);
  for (int n2 = 0; n2 < R2; ++n2)
    for (int k1 = 0; k1 < 16; ++k1) {  // transposed: conflict-free for lane = n2
      twt[k1 * R2 + n2] = twr[n2 * 16 + k1];
      twt[nfft + k1 * R2 + n2] = twi[n2 * 16 + k1];
    }
  {
    uint32_t* wt = host.data();
    std::memcpy(wt, ww.data(), ww.size() * 4);
    std::memcpy(wt + wt_desc, wdesc.data(), wdesc.size() * 4);
    std::memcpy(wt + wt_jinfo, jinfo.data(), jinfo.size() * 4);
    std::memcpy(wt + wt_win, window, Nw * 4);
    std::memcpy(wt + wt_tw, twt.data(), 2 * (size_t)nfft * 4);
    float* dt = reinterpret_cast<float*>(wt + wt_dith);
    for (int i = 0; i < spl::kDitherTab; ++i) {
      const double u = ((double)i + 0.5) / (double)spl::kDitherTab;
      dt[i] = (float)((double)cfg->dither * std::sqrt(-2.0 * std::log(u)) * std::cos(2.0 * M_PI * u));
    }
  }
  for (int f = 0; f < 2; ++f) {
    if (!ut.tab[f].empty()) std::memcpy(host.data() + o_utab[f], ut.tab[f].data(), ut.tab[f].size() * 4);
    if (!ut.twiddles[f].empty()) std::memcpy(host.data() + o_utw[f], ut.twiddles[f].data(), ut.twiddles[f].size());
  }
  size_t o = wt_end;
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
  h->tab.wtab = fb;
  h->tab.wtab_words = (int32_t)wt_words;
  h->tab.wt_off_desc = (int32_t)wt_desc;
  h->tab.wt_off_jinfo = (int32_t)wt_jinfo;
  h->tab.wt_off_win = (int32_t)wt_win;
  h->tab.wt_off_tw = (int32_t)wt_tw;
  h->tab.wt_off_dith = (int32_t)wt_dith;
  h->tab.nj = nj;
  h->umma_nflush = ut.nflush;
  h->umma_nparts = ut.nparts;
  for (int i = 0; i < 4; ++i) h->umma_part_f0[i] = ut.part_f0[i];
  for (int i = 0; i < 5; ++i) h->umma_part_s0[i] = ut.part_s0[i];
  for (int f = 0; f < 2; ++f) {
    h->umma_ok[f] = ut.ok && !ut.tab[f].empty();
    h->umma_tab[f] = fb + o_utab[f];
    h->umma_tw[f] = reinterpret_cast<const uint8_t*>(fb + o_utw[f]);
    h->umma_tab_bytes[f] = (int)(ut.tab[f].size() * 4);
    h->umma_off_melw[f] = ut.off_melw[f];
    h->umma_off_melc[f] = ut.off_melc[f];
    if (h->umma_ok[f] &&
        spl::fbank_umma_smem_bytes(nfft, f == 0 ? 4 : 2, h->umma_tab_bytes[f], h->D_out) > 227 * 1024)
      h->umma_ok[f] = false;
  }
  h->num_sms = 148;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  // SPL_ENGINE: fft (default: warp-pipelined FFT on the fp32x2 pipe -- the faster engine by measurement,
  // profiles/r2_summary.md), umma (tcgen05 DFT-as-GEMM), simple (one tile per CTA, cross-check)
  const char* eng = std::getenv("SPL_ENGINE");
  h->engine = ENGINE_FFT;
  if (eng && !std::strcmp(eng, "umma")) h->engine = ENGINE_UMMA;
  if (eng && !std::strcmp(eng, "simple")) h->engine = ENGINE_SIMPLE;
  const char* cps = std::getenv("SPL_CTAS_PER_SM");  // FFT engine knob: 1 = 8-warp CTAs (two launches co-resident per SM)
  h->ctas_per_sm = (cps && cps[0] == '1') ? 1 : 2;
  h->smem_warp = spl::fbank_warp_smem_bytes(nfft, S, Nw, h->D_out, (int)wt_words, 8);
  h->smem_warp16 = spl::fbank_warp_smem_bytes(nfft, S, Nw, h->D_out, (int)wt_end, 16);
  {  // the warp kernel stages a group's samples inside one pair's exchange planes
    const int pl = ((nfft / 16 * 17 + 15) / 32) * 32 + 16;
    h->fft_ok = h->smem_warp <= 113 * 1024 && 3 * S + Nw + 4 <= 2 * pl;
  }
  h->smem_bytes = spl::fbank_smem_bytes(nfft, S, Nw, D, h->D_out, nnz);
  if (h->smem_bytes > 113 * 1024) {  // two CTAs per SM must fit in 227 KB
    cudaFree(blob);
    delete h;
    return fail(SPL_ERR_UNSUPPORTED, "spl_create: configuration needs too much shared memory");
  }
  h->status = nullptr;
  e = cudaMalloc(reinterpret_cast<void**>(&h->status), 256);
  if (e == cudaSuccess) e = cudaMemset(h->status, 0, 256);
  if (e != cudaSuccess) {
    cudaFree(blob);
    delete h;
    return fail_cuda(e, "spl_create: status word");
  }
  h->umma_ctas = h->num_sms;
  if (const char* uc = std::getenv("SPL_UMMA_CTAS"); uc && std::atoi(uc) > 0) h->umma_ctas = std::atoi(uc);
  h->debug_acc = nullptr;
  if (const char* dbg = std::getenv("SPL_UMMA_DEBUG"); dbg && dbg[0] == '1') {
    if (cudaMalloc(reinterpret_cast<void**>(&h->debug_acc), 128 * 513 * sizeof(float)) == cudaSuccess)
      cudaMemset(h->debug_acc, 0, 128 * 513 * sizeof(float));
    else
      h->debug_acc = nullptr;
  }
  *out = h;
  return SPL_OK;
}

void spl_destroy(spl_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaFree(h->blob);
  cudaFree(h->status);
  cudaFree(h->debug_acc);
  delete h;
}

// ---- kernel A dispatch --------------------------------------------------------------------------------
namespace {

int check_fbank_args(const spl_handle* h, const spl_fbank_args* a, const char* who) {
  if (!a->wav || !a->wav_len || !a->feats) return fail(SPL_ERR_INVALID_ARG, std::string(who) + ": null buffer");
  if (a->B < 1 || a->B > (1 << 22) || a->T < 1) return fail(SPL_ERR_INVALID_ARG, std::string(who) + ": B/T out of range");
  if (a->wav_cols < h->cfg.window_size || a->wav_pitch < a->wav_cols)
    return fail(SPL_ERR_INVALID_ARG, std::string(who) + ": need window_size <= wav_cols <= wav_pitch");
  if (a->sample_format != SPL_SAMPLES_F32 && a->sample_format != SPL_SAMPLES_I16)
    return fail(SPL_ERR_INVALID_ARG, std::string(who) + ": sample_format");
  return SPL_OK;
}

// utterances [b0, b0 + nB) of a batch as a batch of their own (utterances are independent)
spl_fbank_args slice_batch(const spl_handle* h, const spl_fbank_args& a, int b0, int nB) {
  spl_fbank_args s = a;
  const size_t es = a.sample_format == SPL_SAMPLES_F32 ? 4 : 2;
  s.wav = static_cast<const char*>(a.wav) + (size_t)b0 * a.wav_pitch * es;
  s.wav_len = a.wav_len + b0;
  s.B = nB;
  s.feats = a.feats + (size_t)b0 * a.T * h->D_out;
  s.feat_len = a.feat_len ? a.feat_len + b0 : nullptr;
  s.noise = a.noise ? a.noise + (size_t)b0 * a.T * h->cfg.window_size : nullptr;
  s.utt_stats = a.utt_stats ? a.utt_stats + (size_t)b0 * 2 * h->D_out : nullptr;
  return s;
}

// FFT engines: `v[0..n)` share sample format / noise mode / global_stats; the warp-pipelined kernel takes them in
// one launch (<= kMaxBatches batches, <= kMaxPersistentB utterances), the simple kernel one batch per launch
cudaError_t launch_fft(spl_handle* h, const spl_fbank_args* v, int n, cudaStream_t st, bool simple) {
  const spl_fbank_args& a = v[0];
  spl::FbankParams p;
  std::memset(&p, 0, sizeof(p));
  p.S = h->cfg.window_shift;
  p.Nw = h->cfg.window_size;
  p.D = h->cfg.num_mel_bins;
  p.D_out = h->D_out;
  p.use_energy = h->cfg.use_energy;
  p.remove_dc = h->cfg.remove_dc;
  p.preemph = h->cfg.preemph;
  p.dither = h->cfg.dither;
  p.wav = a.wav;
  p.wav_pitch = a.wav_pitch;
  p.wav_cols = a.wav_cols;
  p.sample_format = a.sample_format;
  p.wav_len = a.wav_len;
  p.B = a.B;
  p.T = a.T;
  p.feats = a.feats;
  p.feat_len = a.feat_len;
  p.noise = a.noise;
  p.seed_lo = (uint32_t)(a.dither_seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(a.dither_seed >> 32);
  p.utt_stats = a.utt_stats;
  p.global_stats = a.global_stats;
  p.tab = h->tab;
  p.nb = n;
  int u0 = 0;
  for (int k = 0; k < n; ++k) {
    spl::UBatch& bd = p.bd[k];
    bd.wav = v[k].wav;
    bd.wav_len = v[k].wav_len;
    bd.feats = v[k].feats;
    bd.feat_len = v[k].feat_len;
    bd.noise = v[k].noise;
    bd.utt_stats = v[k].utt_stats;
    bd.wav_pitch = v[k].wav_pitch;
    bd.wav_cols = v[k].wav_cols;
    bd.B = v[k].B;
    bd.T = v[k].T;
    bd.u0 = u0;
    u0 += v[k].B;
  }
  p.total_utts = u0;
  const bool with_noise = h->cfg.dither != 0.f;
  if (simple) return spl::launch_fbank(p, h->cfg.padded_size, with_noise, st);
  // one 16-warp CTA per SM; SPL_CTAS_PER_SM=1 (throughput mode) or a table block too large for 227 KB: 8-warp CTAs
  const bool wide = h->ctas_per_sm == 2 && h->smem_warp16 <= 227 * 1024;
  return spl::launch_fbank_warp(p, h->cfg.padded_size, with_noise, wide ? 16 : 8,
                                wide ? h->num_sms : h->ctas_per_sm * h->num_sms, st);
}

// every batch of `v` shares sample format / noise mode; <= kMaxBatches batches, <= kMaxUmmaUtts utterances
cudaError_t launch_umma(spl_handle* h, const spl_fbank_args* v, int n, cudaStream_t st) {
  const int f = v[0].sample_format == SPL_SAMPLES_F32 ? 0 : 1;
  spl::UmmaParams p;
  std::memset(&p, 0, sizeof(p));
  p.S = h->cfg.window_shift;
  p.Nw = h->cfg.window_size;
  p.D_out = h->D_out;
  p.remove_dc = h->cfg.remove_dc;
  p.preemph = h->cfg.preemph;
  p.dither = h->cfg.dither;
  p.seed_lo = (uint32_t)(v[0].dither_seed & 0xffffffffu);
  p.seed_hi = (uint32_t)(v[0].dither_seed >> 32);
  p.nb = n;
  p.nflush = h->umma_nflush;
  p.nparts = h->umma_nparts;
  for (int i = 0; i < 4; ++i) p.part_f0[i] = h->umma_part_f0[i];
  for (int i = 0; i < 5; ++i) p.part_s0[i] = h->umma_part_s0[i];
  p.global_stats = v[0].global_stats;
  p.status = h->status;
  p.debug_acc = h->debug_acc;
  p.twiddles = h->umma_tw[f];
  p.tab = h->umma_tab[f];
  p.tab_bytes = h->umma_tab_bytes[f];
  p.off_melw = h->umma_off_melw[f];
  p.off_melc = h->umma_off_melc[f];
  int u0 = 0;
  for (int k = 0; k < n; ++k) {
    spl::UBatch& b = p.bd[k];
    b.wav = v[k].wav;
    b.wav_len = v[k].wav_len;
    b.feats = v[k].feats;
    b.feat_len = v[k].feat_len;
    b.noise = v[k].noise;
    b.utt_stats = v[k].utt_stats;
    b.wav_pitch = v[k].wav_pitch;
    b.wav_cols = v[k].wav_cols;
    b.B = v[k].B;
    b.T = v[k].T;
    b.u0 = u0;
    u0 += v[k].B;
    if (v[k].utt_stats) p.want_utt_stats = 1;
  }
  p.total_utts = u0;
  const int noise_mode = h->cfg.dither == 0.f ? 0 : (v[0].noise ? 2 : 1);
  return spl::launch_fbank_umma(p, h->cfg.padded_size, v[0].sample_format, noise_mode, h->umma_ctas, st);
}

}  // namespace

int spl_fbank_forward_multi(spl_handle* h, const spl_fbank_args* args, int32_t n, void* stream) {
  if (!h || !args || n < 1) return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward_multi: null argument");
  for (int i = 0; i < n; ++i) {
    const int rc = check_fbank_args(h, args + i, "spl_fbank_forward");
    if (rc != SPL_OK) return rc;
  }
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_fbank_forward: cudaSetDevice failed");
  NvtxRange range("spl:fbank (kernel A)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // zero the per-utterance sums (contiguous buffers of consecutive batches share one memset)
  for (int i = 0; i < n;) {
    if (!args[i].utt_stats) {
      ++i;
      continue;
    }
    double* beg = args[i].utt_stats;
    double* end = beg + 2 * (size_t)args[i].B * h->D_out;
    int j = i + 1;
    while (j < n && args[j].utt_stats == end) {
      end += 2 * (size_t)args[j].B * h->D_out;
      ++j;
    }
    cudaError_t e = cudaMemsetAsync(beg, 0, (size_t)(end - beg) * sizeof(double), st);
    if (e != cudaSuccess) return fail_cuda(e, "spl_fbank_forward: cudaMemsetAsync");
    i = j;
  }

  // split into launches: pieces of <= kMaxUmmaUtts utterances; the tcgen05 engine takes up to kMaxBatches pieces
  // (same sample format / noise mode / global_stats target) per launch, the FFT engines one piece per launch
  std::vector<spl_fbank_args> pieces;
  for (int i = 0; i < n; ++i)
    for (int b0 = 0; b0 < args[i].B; b0 += spl::kMaxUmmaUtts)
      pieces.push_back(slice_batch(h, args[i], b0, std::min<int>(spl::kMaxUmmaUtts, args[i].B - b0)));
  size_t i = 0;
  while (i < pieces.size()) {
    const spl_fbank_args& a = pieces[i];
    const int f = a.sample_format == SPL_SAMPLES_F32 ? 0 : 1;
    const bool umma = h->engine == ENGINE_UMMA && h->umma_ok[f];
    const bool simple = !umma && (h->engine == ENGINE_SIMPLE || !h->fft_ok);
    cudaError_t e;
    size_t j = i + 1;
    int utts = a.B;
    long long rows = (long long)a.B * a.T;  // row prefixes of one launch are 32-bit
    if (rows > 0x7fffffffLL) return fail(SPL_ERR_UNSUPPORTED, "spl_fbank_forward: more than 2^31 - 1 output rows in 512 utterances");
    while (!simple && j < pieces.size() && j - i < (size_t)spl::kMaxBatches && utts + pieces[j].B <= spl::kMaxUmmaUtts &&
           (rows += (long long)pieces[j].B * pieces[j].T) <= 0x7fffffffLL &&
           pieces[j].sample_format == a.sample_format && (pieces[j].noise != nullptr) == (a.noise != nullptr) &&
           pieces[j].global_stats == a.global_stats) {
      utts += pieces[j].B;
      ++j;
    }
    e = umma ? launch_umma(h, pieces.data() + i, (int)(j - i), st) : launch_fft(h, pieces.data() + i, (int)(j - i), st, simple);
    i = j;
    if (e != cudaSuccess) return fail_cuda(e, "spl_fbank_forward: launch");
    g_launches.fetch_add(1);
  }
  return SPL_OK;
}

int spl_fbank_forward(spl_handle* h, const spl_fbank_args* a, void* stream) {
  if (!h || !a) return fail(SPL_ERR_INVALID_ARG, "spl_fbank_forward: null argument");
  return spl_fbank_forward_multi(h, a, 1, stream);
}

int spl_debug_status(spl_handle* h) {
  if (!h) return -1;
  DeviceGuard guard(h->device);
  int32_t v = -1;
  if (cudaMemcpy(&v, h->status, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

int spl_debug_umma_tables(int32_t nfft, int32_t Nw, int32_t D, const float* window, const float* mel_dense, int32_t fmt,
                          void* twiddles, size_t twiddle_cap, float* tab, size_t tab_cap, int32_t* info) {
  if (!window || !mel_dense || !info || fmt < 0 || fmt > 1) return fail(SPL_ERR_INVALID_ARG, "spl_debug_umma_tables: bad argument");
  spl::UmmaHostTables ut;
  spl::build_umma_tables(nfft, Nw, D, window, mel_dense, ut);
  info[0] = ut.ok && !ut.tab[fmt].empty();
  info[1] = (int32_t)ut.twiddles[fmt].size();
  info[2] = (int32_t)ut.tab[fmt].size();
  info[3] = ut.off_melw[fmt];
  info[4] = ut.off_melc[fmt];
  info[5] = ut.nflush;
  info[6] = ut.nparts;
  for (int i = 0; i < 4; ++i) info[7 + i] = ut.part_f0[i];
  for (int i = 0; i < 5; ++i) info[11 + i] = ut.part_s0[i];
  if (!info[0]) return SPL_OK;
  if (twiddles && twiddle_cap >= ut.twiddles[fmt].size()) std::memcpy(twiddles, ut.twiddles[fmt].data(), ut.twiddles[fmt].size());
  if (tab && tab_cap >= ut.tab[fmt].size()) std::memcpy(tab, ut.tab[fmt].data(), ut.tab[fmt].size() * 4);
  return SPL_OK;
}

int spl_debug_umma_acc(spl_handle* h, float* host_out, size_t n_floats) {
  if (!h || !h->debug_acc || !host_out) return fail(SPL_ERR_INVALID_ARG, "spl_debug_umma_acc: needs SPL_UMMA_DEBUG=1 at spl_create");
  DeviceGuard guard(h->device);
  const size_t n = std::min<size_t>(n_floats, 128 * 513);
  cudaError_t e = cudaMemcpy(host_out, h->debug_acc, n * sizeof(float), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? SPL_OK : fail_cuda(e, "spl_debug_umma_acc");
}

int spl_debug_dither_noise(spl_handle* h, float* out, int32_t B, int32_t T, uint64_t seed, void* stream) {
  if (!h || !out || B < 1 || T < 1) return fail(SPL_ERR_INVALID_ARG, "spl_debug_dither_noise: bad argument");
  DeviceGuard guard(h->device);
  // the stream of the default configuration: the 16-warp FFT engine draws from the table, the 8-warp variant
  // (SPL_CTAS_PER_SM=1) evaluates the formula on a 16-bit grid
  const bool wide = h->ctas_per_sm == 2 && h->smem_warp16 <= 227 * 1024;
  cudaError_t e = spl::launch_dither_noise(out, B, T, h->cfg.window_size, h->cfg.padded_size / 16, seed,
                                           wide ? h->tab.wtab + h->tab.wt_off_dith : nullptr, h->cfg.dither,
                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, "spl_debug_dither_noise: launch");
  g_launches.fetch_add(1);
  return SPL_OK;
}

const char* spl_engine_name(const spl_handle* h, int32_t sample_format) {
  if (!h) return "";
  const int f = sample_format == SPL_SAMPLES_F32 ? 0 : 1;
  if (h->engine == ENGINE_UMMA && h->umma_ok[f]) return "umma";
  return (h->engine == ENGINE_SIMPLE || !h->fft_ok) ? "simple" : "fft";
}

namespace {

int check_post_args(const spl_post_args* a, const spl_post_args* first) {
  if (!a->feats || !a->feat_len) return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: null argument");
  if (a->B < 1 || a->B > (1 << 22) || a->T < 1 || a->Dm < 1 || a->Dm > spl::kMaxDm)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: B/T/Dm out of range (Dm <= 160)");
  const bool masks = a->mask_params || a->mask_uniforms;
  const int nmask = masks ? a->n_freq_masks + a->n_time_masks : 0;
  if (a->n_freq_masks < 0 || a->n_time_masks < 0 || nmask > spl::kMaxMasks)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: at most 32 masks per utterance");
  if (a->cmvn_mode == SPL_CMVN_UTTERANCE && !a->utt_stats)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: utterance CMVN needs utt_stats");
  if (a->cmvn_mode == SPL_CMVN_GLOBAL && (!a->global_mean || (a->norm_vars && !a->global_istd)))
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: global CMVN needs global_mean/global_istd");
  if (masks && a->n_time_masks > 0 && !a->utt_stats)
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: time masks need utt_stats (time means)");
  if (a->Dm != first->Dm || a->cmvn_mode != first->cmvn_mode || a->norm_vars != first->norm_vars ||
      a->n_freq_masks != first->n_freq_masks || a->n_time_masks != first->n_time_masks ||
      a->global_mean != first->global_mean || a->global_istd != first->global_istd ||
      a->freq_mask_width != first->freq_mask_width || a->time_mask_width != first->time_mask_width ||
      masks != (first->mask_params || first->mask_uniforms))
    return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace_multi: the batches of one call must share Dm, CMVN and mask settings");
  return SPL_OK;
}

}  // namespace

int spl_post_inplace_multi(spl_handle* h, const spl_post_args* args, int32_t n, void* stream) {
  if (!args || n < 1) return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: null argument");
  for (int i = 0; i < n; ++i) {
    const int rc = check_post_args(args + i, args);
    if (rc != SPL_OK) return rc;
  }
  const bool masks = args[0].mask_params || args[0].mask_uniforms;
  if (args[0].cmvn_mode == SPL_CMVN_NONE && (!masks || args[0].n_freq_masks + args[0].n_time_masks == 0)) return SPL_OK;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  DeviceGuard guard(h ? h->device : cur_dev);
  if (!guard.ok) return fail(SPL_ERR_CUDA, "spl_post_inplace: cudaSetDevice failed");
  NvtxRange range("spl:cmvn+specaug (kernel B)");
  for (int i0 = 0; i0 < n; i0 += spl::kMaxPostBatches) {
    spl::PostParams p;
    std::memset(&p, 0, sizeof(p));
    const spl_post_args& a0 = args[0];
    p.Dm = a0.Dm;
    p.cmvn_mode = a0.cmvn_mode;
    p.norm_vars = a0.norm_vars;
    p.n_freq = masks ? a0.n_freq_masks : 0;
    p.n_time = masks ? a0.n_time_masks : 0;
    p.freq_width = a0.freq_mask_width;
    p.time_width = a0.time_mask_width;
    p.global_mean = a0.global_mean;
    p.global_istd = a0.global_istd;
    p.nb = std::min<int>(spl::kMaxPostBatches, n - i0);
    int u0 = 0;
    for (int k = 0; k < p.nb; ++k) {
      const spl_post_args& a = args[i0 + k];
      spl::PostBatch& b = p.bd[k];
      b.feats = a.feats;
      b.feat_len = a.feat_len;
      b.utt_stats = a.utt_stats;
      b.mask_params = a.mask_params;
      b.mask_uniforms = a.mask_params ? nullptr : a.mask_uniforms;
      b.B = a.B;
      b.T = a.T;
      b.u0 = u0;
      u0 += a.B;
    }
    if (u0 > 65535) return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: more than 65535 utterances in one launch");
    cudaError_t e = spl::launch_post(p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "spl_post_inplace: launch");
    g_launches.fetch_add(1);
  }
  return SPL_OK;
}

int spl_post_inplace(spl_handle* h, const spl_post_args* a, void* stream) {
  if (!a) return fail(SPL_ERR_INVALID_ARG, "spl_post_inplace: null argument");
  return spl_post_inplace_multi(h, a, 1, stream);
}

int spl_forward_multi(spl_handle* h, const spl_fbank_args* fbank, const spl_post_args* post, int32_t n, void* stream) {
  if (!h || !fbank || n < 1) return fail(SPL_ERR_INVALID_ARG, "spl_forward_multi: null argument");
  int rc = spl_fbank_forward_multi(h, fbank, n, stream);
  if (rc != SPL_OK || !post) return rc;
  std::vector<spl_post_args> pa(post, post + n);
  for (int i = 0; i < n; ++i) {
    if (!pa[i].feats) pa[i].feats = fbank[i].feats;
    if (!pa[i].feat_len) pa[i].feat_len = fbank[i].feat_len;
    if (!pa[i].utt_stats) pa[i].utt_stats = fbank[i].utt_stats;
    if (!pa[i].B) pa[i].B = fbank[i].B;
    if (!pa[i].T) pa[i].T = fbank[i].T;
    if (!pa[i].Dm) pa[i].Dm = h->D_out;
  }
  return spl_post_inplace_multi(h, pa.data(), n, stream);
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
