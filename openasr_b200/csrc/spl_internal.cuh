// Internal declarations shared by the C-ABI translation unit and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spl_capi.h"

namespace spl {

constexpr int kTileFrames = 32;   // frames per CTA tile (8 warps x 4 frames)
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxMel = 128;
constexpr int kMaxDm = 160;     // widest feature row kernel B handles (offline features included)
constexpr int kMaxMasks = 32;   // freq + time masks per utterance
constexpr float kEps = 1.1920928955078125e-07f;  // kaldi_signal.py:48

// Device-resident constant tables of one handle.
struct Tables {
  const float* window;     // [Nw]
  const float* tw_re;      // [R2 * 16] stage-1 twiddles cos(2 pi n2 k1 / Nfft)
  const float* tw_im;      // [R2 * 16] sin(2 pi n2 k1 / Nfft)
  const float* mel_w;      // [nnz] packed non-zero mel weights (already scaled by 1/4, see kernel)
  const int32_t* mel_lo;   // [D] first FFT bin of filter
  const int32_t* mel_cnt;  // [D] number of bins
  const int32_t* mel_off;  // [D] offset into mel_w
  int32_t mel_nnz;
  int32_t grp_beg[kWarps + 1];  // filters [grp_beg[w], grp_beg[w+1]) handled by warp w in the mel phase
  // persistent kernel: ONE contiguous block that is bulk-copied (TMA) into shared memory as is:
  //   [pair weights | pair descriptors | window | stage-1 twiddles, transposed]
  // Filters are processed two at a time; every pair is padded (zero weights) to a common number of
  // 16-byte aligned 4-bin groups kept inside [0, Nfft/2).
  //   weights     : for pair i, group g: 8 floats {wa[4], wb[4]}
  //   descriptors : loA/4 | loB/4 << 6 | n4 << 12 | (weight offset / 8) << 18 | validB << 31
  const float* ptab;
  int32_t ptab_words;           // total 4-byte words, multiple of 4
  int32_t pt_off_desc, pt_off_win, pt_off_tw;  // word offsets inside the block
  int32_t npairs;
  int32_t pgrp_beg[kWarps + 1]; // pairs [pgrp_beg[w], pgrp_beg[w+1]) handled by warp w in the mel phase
  // warp-pipelined kernel: mel phase with lane = (frame f = lane & 3, slice s = lane >> 2); in
  // iteration j slice s handles filter pair 8 j + s, all eight pairs padded to n4j groups.
  //   [weights: float4 index ((goff_j + g) * 2 + half) * 8 + s | pair descriptors (8 per j):
  //    loA/4 | loB/4 << 6 | validA << 30 | validB << 31 | jinfo: n4j | goff_j << 8 | window | twiddles^T]
  const float* wtab;
  int32_t wtab_words;
  int32_t wt_off_desc, wt_off_jinfo, wt_off_win, wt_off_tw;
  int32_t nj;
  // tcgen05 DFT-as-GEMM kernel (fbank_tc.cu).  HALF = Nfft/4 outputs / inputs per folded block.
  //   tc_b   : twiddle operand images, [Nfft/32 units][4 blocks: cos-even, cos-odd, sin-even, sin-odd]
  //            [hi, lo][HALF n x 8 k tf32, K-major SWIZZLE_32B] -- copied to shared memory verbatim
  //   tc_tab : [segment weights | segment descriptors (uint2) | window | Wc | Ws], one TMA bulk copy
  //   segment descriptor .x: start/4 | n4 << 6 | array (0: bins 1..HALF, 1: mirrored bins) << 12 |
  //            first << 13 | last << 14 | filter << 16 ;  .y: weight offset / 4
  const float* tc_b;
  const float* tc_tab;
  int32_t tc_tab_words, tc_off_desc, tc_off_win, tc_off_wc, tc_off_ws;
  int32_t tc_nseg;
  int32_t tc_sgrp_beg[5];
  // pair-pipelined kernel (fbank_pair.cu, Nfft = 512): mel phase with lane = (frame f = lane & 1,
  // slice s = lane >> 1) and two filter streams t per slice; stream (s, t) walks qE flat entries.
  //   [weights: float4 index (e * 2 + t) * 16 + s | entry descriptors: uint2 (t = 0, t = 1) at e * 16 + s,
  //    each  bin/4 | filter << 8 | last-of-filter << 16 | window | stage-1 twiddles^T |
  //    stage-2 twiddles: float4 {cos, sin (a = typ), cos, sin (a = typ + 2)} of 2 pi a k / 32 at typ * 8 + k]
  const float* qtab;
  int32_t qtab_words, qt_off_desc, qt_off_win, qt_off_tw, qt_off_tw2, qE;
};

struct FbankParams {
  // config
  int32_t S, Nw, D, D_out, use_energy, remove_dc;
  float preemph, dither;
  // call
  const void* wav;
  int64_t wav_pitch, wav_cols;
  int32_t sample_format;
  const int64_t* wav_len;
  int32_t B, T;
  float* feats;
  int64_t* feat_len;
  const float* noise;
  uint32_t seed_lo, seed_hi;
  double* utt_stats;
  double* global_stats;
  Tables tab;
};

struct PostParams {
  float* feats;
  const int64_t* feat_len;
  int32_t B, T, Dm;
  int32_t cmvn_mode, norm_vars;
  const double* utt_stats;
  const float* global_mean;
  const float* global_istd;
  int32_t n_freq, n_time;
  const int32_t* mask_params;
};

// host-side launchers (defined in the .cu files); return cudaError_t of the launch
cudaError_t launch_fbank(const FbankParams& p, int nfft, bool with_noise, cudaStream_t st);
size_t fbank_smem_bytes(int nfft, int S, int Nw, int D, int D_out, int nnz);
// persistent, load-balanced kernel (B <= kMaxPersistentB); num_ctas = 2 * SM count
constexpr int kMaxPersistentB = 512;
cudaError_t launch_fbank_persistent(const FbankParams& p, int nfft, bool with_noise, int num_ctas, cudaStream_t st);
size_t fbank_persistent_smem_bytes(int nfft, int S, int Nw, int D_out, int ptab_words);
// warp-pipelined kernel: one CTA of 16 independent warps per SM, no block-wide barriers in the loop
cudaError_t launch_fbank_warp(const FbankParams& p, int nfft, bool with_noise, int warps, int num_ctas, cudaStream_t st);
size_t fbank_warp_smem_bytes(int nfft, int S, int Nw, int D_out, int wtab_words, int warps);
// pair-pipelined kernel (Nfft = 512 only): 2 CTAs x 12 warps per SM, one packed pair per warp iteration
cudaError_t launch_fbank_pair(const FbankParams& p, bool with_noise, int num_ctas, cudaStream_t st);
size_t fbank_pair_smem_bytes(int D_out, int qtab_words);
cudaError_t launch_post(const PostParams& p, cudaStream_t st);
cudaError_t launch_column_stats(const float* feats, const int64_t* feat_len, int B, int T, int Dm,
                                double* utt_stats, cudaStream_t st);

// Conv2d(1 -> C, 3x3, stride (2, 1)) + ReLU on the [B, T, D] features (conv0_kernel.cu); C <= 64
cudaError_t launch_conv0_relu(const float* x, const float* w, const float* bias, float* out, int B, int T, int D, int C,
                              cudaStream_t st);

// tcgen05 DFT-as-GEMM fbank kernel (fp32 samples, B <= kMaxPersistentB); one CTA per SM
cudaError_t launch_fbank_tc(const FbankParams& p, int nfft, bool with_noise, int num_ctas, cudaStream_t st);
size_t fbank_tc_smem_bytes(int nfft, int D_out, int tc_tab_words);

// tcgen05 building-block self-test (tc_selftest.cu)
cudaError_t launch_tc_selftest_sw32(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st);
cudaError_t launch_tc_selftest(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st);

}  // namespace spl
