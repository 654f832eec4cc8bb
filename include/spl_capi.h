/*
 * spl_capi.h -- C ABI of the B200-native OpenASR speech front-end (SPLayer hot path).
 *
 * Drop-in boundary: this library replaces the arithmetic behind the reference's
 *   src/blocks/sp_layers.py:76-101   SPLayer.forward   (per-utterance fbank loop, pad/stack)
 *   src/blocks/sp_layers.py:51-74    SPLayer.spec_aug  (frequency/time masking)
 *   src/third_party/kaldi_signal.py:458-552  fbank     (+ :163-211 _get_window, :67-106 _get_strided)
 * The reference has no FFI (it is pure Python/PyTorch); the binding a maintainer adds is
 * the ctypes stub in openasr_b200/_capi.py (shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain C types and raw device pointers only; no torch/ATen types.
 *  - every entry point returns 0 on success, a negative spl_status otherwise;
 *    spl_last_error() returns a thread-local message for the last failure.
 *  - all launches are asynchronous on the caller's stream; nothing here synchronises.
 *  - the library never allocates outputs or workspaces: the caller owns all buffers
 *    (PyTorch's caching allocator in the Python host).  Constant tables (window,
 *    sparse mel bank) live in the handle, created once per (device, config).
 *  - thread-safe: a handle may be used concurrently from several host threads
 *    (DataParallel worker threads, src/train.py:134) as long as each call uses
 *    its own buffers.
 */
#ifndef SPL_CAPI_H_
#define SPL_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_ABI_VERSION 2

#if defined(__GNUC__)
#define SPL_API __attribute__((visibility("default")))
#else
#define SPL_API
#endif

typedef enum spl_status {
  SPL_OK = 0,
  SPL_ERR_INVALID_ARG = -1,
  SPL_ERR_UNSUPPORTED = -2, /* e.g. a sample rate whose padded window is not 256/512 */
  SPL_ERR_CUDA = -3,
  SPL_ERR_SHORT_UTTERANCE = -4 /* host-side length check: n_i < window (kaldi_signal.py:154) */
} spl_status;

typedef enum spl_window { SPL_WINDOW_POVEY = 0, SPL_WINDOW_HAMMING = 1, SPL_WINDOW_OTHER = 2 } spl_window;
typedef enum spl_cmvn { SPL_CMVN_NONE = 0, SPL_CMVN_UTTERANCE = 1, SPL_CMVN_GLOBAL = 2 } spl_cmvn;
typedef enum spl_sample_format { SPL_SAMPLES_F32 = 0, SPL_SAMPLES_I16 = 1 } spl_sample_format;

/* Front-end configuration; mirrors the keys SPLayer.__init__ reads (sp_layers.py:27-46)
 * plus the fbank defaults it leaves in place (kaldi_signal.py:458-464). */
typedef struct spl_config {
  int32_t abi_version;   /* SPL_ABI_VERSION */
  int32_t window_shift;  /* S    = int(sr * 10 * 0.001)   kaldi_signal.py:150 */
  int32_t window_size;   /* Nw   = int(sr * 25 * 0.001)   kaldi_signal.py:151 */
  int32_t padded_size;   /* Nfft = next pow2(Nw)          kaldi_signal.py:152 ; 256 or 512 */
  int32_t num_mel_bins;  /* D    (4..128) */
  int32_t use_energy;    /* 1: prepend raw log-energy column (kaldi_signal.py:543-549) */
  int32_t remove_dc;     /* 1 in the reference */
  float preemph;         /* 0.97 in the reference */
  float dither;          /* 1.0 in the reference (always on: sp_layers.py:41-46 never passes it) */
} spl_config;

typedef struct spl_handle spl_handle;

/* window[Nw] and mel_dense[D * Nfft/2] are HOST arrays built by the caller with the
 * reference's own formulas (kaldi_signal.py:109-128 and :389-455) so the tables are
 * bit-identical to the reference's; the handle uploads them (mel as a sparse bank). */
SPL_API int spl_create(const spl_config* cfg, const float* window, const float* mel_dense, int device,
               spl_handle** out);
SPL_API void spl_destroy(spl_handle* h);

/* Per-call arguments of the fused fbank kernel (kernel A).  Device pointers unless noted. */
typedef struct spl_fbank_args {
  const void* wav;         /* [B, wav_pitch] samples, int16-scaled; f32 or i16 per sample_format */
  int64_t wav_pitch;       /* elements between utterance rows (>= wav_cols) */
  int64_t wav_cols;        /* addressable elements per row (>= max n_i); the library never touches
                              memory outside [wav, wav + (B-1)*wav_pitch + wav_cols) */
  int32_t sample_format;   /* spl_sample_format */
  const int64_t* wav_len;  /* [B] valid samples per utterance */
  int32_t B;
  int32_t T;               /* output frames per row = max_i m_i (host-computed) */
  float* feats;            /* [B, T, D_out] fp32, D_out = D + use_energy; fully overwritten
                              (rows t >= m_i are written as exactly 0.0) */
  int64_t* feat_len;       /* [B] m_i = 1 + (n_i - Nw) / S, or NULL */
  const float* noise;      /* parity mode: [B, T, Nw] host-drawn rand_gauss (kaldi_signal.py:177), or NULL */
  uint64_t dither_seed;    /* throughput mode (noise == NULL, dither != 0): Philox key */
  double* utt_stats;       /* [B, 2, D_out] fp64 sum x / sum x^2 over valid frames; zeroed by the
                              library (memset on `stream`) before the kernel accumulates into it;
                              NULL to skip (no CMVN / SpecAug time means needed) */
  double* global_stats;    /* [2*D_out + 1] fp64 running sum x, sum x^2, frame count (accumulated,
                              never reset by the library); NULL to skip */
} spl_fbank_args;

SPL_API int spl_fbank_forward(spl_handle* h, const spl_fbank_args* a, void* stream /* cudaStream_t */);

/* Several padded batches in ONE persistent launch of kernel A (the per-utterance loop of sp_layers.py:81-91
 * over a queue of batches): `args[0..n)` as for spl_fbank_forward; every batch has its own buffers, lengths
 * and T.  Launch, prologue and tail are paid once, and the frames of all batches are balanced over the SMs
 * together.  Up to 16 batches / 512 utterances share a launch; larger calls are split transparently.
 * dither_seed and global_stats are taken from the first batch of each launch. */
SPL_API int spl_fbank_forward_multi(spl_handle* h, const spl_fbank_args* args, int32_t n, void* stream);

/* CMVN + SpecAug in place on [B, T, Dm] (kernel B).  Closed form of sp_layers.py:51-74:
 * time-masked -> time mean of the (normalised, un-masked) features; else freq-masked ->
 * per-frame mean over Dm; else the (normalised) value.  Rows t >= feat_len stay 0 except
 * where the reference itself spills (frequency masks write the row mean, 0, there). */
typedef struct spl_post_args {
  float* feats;              /* [B, T, Dm] in place */
  const int64_t* feat_len;   /* [B] */
  int32_t B, T, Dm;
  int32_t cmvn_mode;         /* spl_cmvn */
  int32_t norm_vars;         /* 1: divide by std-dev */
  const double* utt_stats;   /* [B, 2, Dm] from kernel A (utterance CMVN, and time means) or NULL */
  const float* global_mean;  /* [Dm] (global CMVN) or NULL */
  const float* global_istd;  /* [Dm] 1/std (global CMVN) or NULL */
  int32_t n_freq_masks, n_time_masks;
  const int32_t* mask_params; /* [B, n_freq+n_time, 2] (start, end) half-open, host-drawn; NULL = no SpecAug
                                 unless mask_uniforms is given */
  /* Sync-free alternative to mask_params: the 2*(n_freq+n_time) x B uniforms in the reference's draw order
   * (sp_layers.py:58-71; row 2j = width draw, row 2j+1 = start draw of mask j).  The kernel resolves them against
   * the DEVICE feat_len with the reference's float32 arithmetic and Python slice semantics -- bit-identical to
   * spl_specaug_rects -- so the host never needs the frame counts. */
  const float* mask_uniforms;
  float freq_mask_width, time_mask_width; /* W_f, W_t (used with mask_uniforms) */
} spl_post_args;

/* `h` may be NULL for spl_post_inplace / spl_column_stats (offline-feature mode has no fbank
 * handle): the current CUDA device is used. */
SPL_API int spl_post_inplace(spl_handle* h, const spl_post_args* a, void* stream);
/* Kernel B over several batches in one launch (same Dm, CMVN mode, mask counts / widths and global tables, taken
 * from args[0]; per-batch buffers, B, T).  Up to 8 batches per launch; larger calls are split. */
SPL_API int spl_post_inplace_multi(spl_handle* h, const spl_post_args* args, int32_t n, void* stream);
/* The whole SPLayer.forward device work in ONE call: kernel A over `n` batches (spl_fbank_forward_multi) followed by
 * kernel B over the same batches (post[i].feats / feat_len / utt_stats / B / T / Dm default to fbank[i]'s when 0).
 * post == NULL: kernel A only. */
SPL_API int spl_forward_multi(spl_handle* h, const spl_fbank_args* fbank, const spl_post_args* post, int32_t n, void* stream);

/* Per-utterance column sums for the offline (pre-computed feature) SpecAug path
 * (sp_layers.py:92-99): utt_stats[B,2,Dm] += sum_t x, sum_t x^2 over t < feat_len. */
SPL_API int spl_column_stats(spl_handle* h, const float* feats, const int64_t* feat_len, int32_t B, int32_t T,
                     int32_t Dm, double* utt_stats, void* stream);

/* Row f2: first layer of the conv-subsampling block on the [B, T, D] features, replacing
 * Conv2d(1, C, 3, (2, 1)) + ReLU of src/blocks/conv_layers.py:125-126 ("subsample/conv0",
 * "subsample/relu0") together with the unsqueeze(1) of :139:
 *   out[b, c, t1, d1] = max(0, bias[c] + sum_{i,j<3} weight[c, 0, i, j] * feats[b, 2 t1 + i, d1 + j])
 * feats [B, T, D] fp32 contiguous; weight [C, 1, 3, 3] and bias [C] (NULL = no bias) in torch's
 * parameter layout; out [B, C, (T-3)/2+1, D-2] fp32 contiguous.  T >= 3, D >= 3, 1 <= C <= 64.
 * `h` may be NULL (current device).  Rows t >= feat_len are convolved like any other row, exactly
 * as the reference does on its zero-padded batch. */
SPL_API int spl_conv0_relu(spl_handle* h, const float* feats, int32_t B, int32_t T, int32_t D, const float* weight,
                           const float* bias, int32_t C, float* out, void* stream);

/* Host-side helper (no device work): turn the 2*(F+T) x B uniforms drawn in the reference's order
 * (sp_layers.py:58-71) into [B, F+T, 2] half-open mask rectangles with the reference's float32
 * arithmetic and Python slice semantics.  `frames` = valid frames per utterance (host). */
SPL_API int spl_specaug_rects(const float* uniforms, const int64_t* frames, int32_t B, int32_t T, int32_t V,
                              int32_t n_freq, float freq_width, int32_t n_time, float time_width, int32_t* out);

/* tcgen05 building-block self-test: D[128,N] = A[128,K] * B[N,K]^T in kind::tf32 (operands are used
 * as TF32, i.e. the low 13 mantissa bits are ignored), accumulator in TMEM.  Device pointers;
 * *status (device int) becomes non-zero if the MMA completion barrier timed out. */
SPL_API int spl_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t* status,
                            void* stream);

/* Introspection */
/* kernel-A engine a call with this sample format runs on: "umma" (tcgen05 DFT-as-GEMM), "fft", "simple" */
SPL_API const char* spl_engine_name(const spl_handle* h, int32_t sample_format);
/* Host-only (no device needed): the tcgen05 engine's tables for a configuration, for the CPU model of the kernel
 * in tests/ (tools/emulate_umma.py).  fmt 0 = fp32 samples, 1 = int16.  info[16] = {supported, twiddle bytes, table
 * floats, offset of the mel weights, offset of the shift codes, trailing emits, epilogue parts, first filter of part 0..3,
 * first step of part 0..4}.  Buffers may be NULL to query sizes. */
SPL_API int spl_debug_umma_tables(int32_t nfft, int32_t Nw, int32_t D, const float* window, const float* mel_dense,
                                  int32_t fmt, void* twiddles, size_t twiddle_cap, float* tab, size_t tab_cap,
                                  int32_t* info);
/* Diagnostics: the unit-dither noise g[b, t, j] (B x T x window_size floats, device) the FFT engine's device RNG adds
 * for `seed` -- the same Philox stream, sample by sample -- so tests can check the noise itself against the
 * reference's rand_gauss (kaldi_signal.py:174-178) and replay it through the host-noise mode. */
SPL_API int spl_debug_dither_noise(spl_handle* h, float* out, int32_t B, int32_t T, uint64_t seed, void* stream);
/* SYNCHRONOUS diagnostics: raw tcgen05 accumulators [128][4*Nfft/4 + 1] of the first tile (handle created with
 * SPL_UMMA_DEBUG=1 in the environment) */
SPL_API int spl_debug_umma_acc(spl_handle* h, float* host_out, size_t n_floats);
/* SYNCHRONOUS (tests only): 0 = no kernel of this handle ever timed out on an internal barrier */
SPL_API int spl_debug_status(spl_handle* h);
SPL_API int spl_feature_dim(const spl_handle* h);       /* D_out */
SPL_API int spl_abi_version(void);
SPL_API const char* spl_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
SPL_API uint64_t spl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPL_CAPI_H_ */
