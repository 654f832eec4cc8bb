// Internal declarations shared by the C-ABI translation unit and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

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
  // warp-pipelined kernel: mel phase with lane = (frame f = lane & 3, slice s = lane >> 2); in
  // iteration j slice s handles filter pair 8 j + s, all eight pairs padded to n4j groups.
  //   [weights: float4 index ((goff_j + g) * 2 + half) * 8 + s | pair descriptors (8 per j):
  //    loA/4 | loB/4 << 6 | validA << 30 | validB << 31 | jinfo: n4j | goff_j << 8 | window | twiddles^T]
  const float* wtab;
  int32_t wtab_words;
  int32_t wt_off_desc, wt_off_jinfo, wt_off_win, wt_off_tw;
  // dither table (16-warp variant): kDitherTab values d * g(u_i), u_i = (i + 1/2) / kDitherTab, of the reference's
  // one-uniform transform g(u) = sqrt(-2 ln u) cos(2 pi u) (kaldi_signal.py:176-177), right after the block above
  int32_t wt_off_dith;
  int32_t nj;
};

constexpr int kDitherTab = 4096;   // entries of the dither table (12-bit uniforms)
constexpr int kMaxBatches = 16;     // batches per launch
constexpr int kMaxUmmaUtts = 512;   // flattened utterances per launch (frame prefix table in shared memory)

struct UBatch {  // one padded batch of a (multi-)call; device pointers
  const void* wav;
  const int64_t* wav_len;
  float* feats;
  int64_t* feat_len;
  const float* noise;
  double* utt_stats;
  int64_t wav_pitch, wav_cols;
  int32_t B, T;
  int32_t u0;    // index of the batch's first utterance in the flattened list
  int32_t pad_;
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
  // warp-pipelined kernel: the call as a list of batches (nb >= 1; the single-batch fields above describe bd[0])
  int32_t nb, total_utts;
  UBatch bd[kMaxBatches];
};

constexpr int kMaxPostBatches = 8;

struct PostBatch {  // one batch of a (multi-)call of kernel B
  float* feats;
  const int64_t* feat_len;
  const double* utt_stats;
  const int32_t* mask_params;   // [B, F+T, 2] resolved rectangles, or NULL
  const float* mask_uniforms;   // [2 (F+T), B] uniforms in the reference's draw order (resolved in the kernel), or NULL
  int32_t B, T;
  int32_t u0;                   // first flattened utterance of the batch
  int32_t pad_;
};

struct PostParams {
  int32_t Dm, cmvn_mode, norm_vars, n_freq, n_time, nb;
  float freq_width, time_width;
  const float* global_mean;
  const float* global_istd;
  PostBatch bd[kMaxPostBatches];
};

// host-side launchers (defined in the .cu files); return cudaError_t of the launch
cudaError_t launch_fbank(const FbankParams& p, int nfft, bool with_noise, cudaStream_t st);
size_t fbank_smem_bytes(int nfft, int S, int Nw, int D, int D_out, int nnz);
constexpr int kMaxPersistentB = 512;  // utterances per launch of the warp-pipelined kernel (prefix tables in shared memory)
// warp-pipelined kernel: one CTA of 16 independent warps per SM, no block-wide barriers in the loop
cudaError_t launch_fbank_warp(const FbankParams& p, int nfft, bool with_noise, int warps, int num_ctas, cudaStream_t st);
size_t fbank_warp_smem_bytes(int nfft, int S, int Nw, int D_out, int wtab_words, int warps);
cudaError_t launch_post(const PostParams& p, cudaStream_t st);
cudaError_t launch_column_stats(const float* feats, const int64_t* feat_len, int B, int T, int Dm,
                                double* utt_stats, cudaStream_t st);

// Conv2d(1 -> C, 3x3, stride (2, 1)) + ReLU on the [B, T, D] features (conv0_kernel.cu); C <= 64
cudaError_t launch_conv0_relu(const float* x, const float* w, const float* bias, float* out, int B, int T, int D, int C,
                              cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// tcgen05 DFT-as-GEMM engine (fbank_umma.cu): persistent, multi-batch (UBatch, kMaxBatches above)
struct UmmaParams {
  int32_t S, Nw, D_out, remove_dc;
  float preemph, dither;
  uint32_t seed_lo, seed_hi;
  int32_t nb, total_utts, nflush, want_utt_stats;
  int32_t nparts, part_f0[4], part_s0[5];  // epilogue: bin parts (1, 2 or 4): first step (multiple of 8) and the filter
                                           // the running sums start on in each part
  double* global_stats;
  int32_t* status;          // device word: 0x10000 | barrier | warp << 8 of the first barrier wait that timed out
  float* debug_acc;         // diagnostics: raw accumulators of tile 0 of CTA 0, [128][4 HALF + 1], or NULL
  const uint8_t* twiddles;  // [(Nfft/64 + 1) chunks][cos | sin half][4 tiles of HALF x 16 FP16, SWIZZLE_32B]
  const float* tab;         // [delayed windows: shifts x Nfft | mel steps: float2 (w_a, w_b) x Nfft/2 | shift codes]
  int32_t tab_bytes, off_melw, off_melc, pad_;
  UBatch bd[kMaxBatches];
};

cudaError_t launch_fbank_umma(const UmmaParams& p, int nfft, int sample_format, int noise_mode, int num_ctas, cudaStream_t st);
size_t fbank_umma_smem_bytes(int nfft, int es, int tab_bytes, int D_out);

// Host-built tables of the engine (umma_tables.cu); `ok == false`: the configuration is outside the engine's domain
struct UmmaHostTables {
  bool ok = false;
  int nflush = 0;
  int nparts = 1, part_f0[4] = {0, 0, 0, 0}, part_s0[5] = {0, 0, 0, 0, 0};
  int off_melw[2] = {0, 0}, off_melc[2] = {0, 0};  // float offsets inside tab[fmt]
  std::vector<uint8_t> twiddles[2];  // [0] fp32 samples (4 shifts), [1] int16 samples (8 shifts)
  std::vector<float> tab[2];
};
void build_umma_tables(int nfft, int Nw, int D, const float* window, const float* mel_dense, UmmaHostTables& out);

// device dither generator, sample by sample (debug_kernels.cu)
cudaError_t launch_dither_noise(float* out, int B, int T, int Nw, int R2, uint64_t seed, const float* dtab, float dither,
                                cudaStream_t st);

// tcgen05 building-block self-test (tc_selftest.cu)
cudaError_t launch_tc_selftest_sw32(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st);
cudaError_t launch_tc_selftest(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st);

}  // namespace spl
