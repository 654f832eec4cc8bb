// Kernel B: CMVN (extension) + SpecAug (src/blocks/sp_layers.py:51-74) in place
// on the [B, T, Dm] feature tensor kernel A just wrote (5-8 MB: L2 resident on B200).
//
// SpecAug closed form.  The reference computes freq_means[b,t] = mean_d x and
// time_means[b,d] = sum_t x / len ONCE from the un-masked input (:52-54), applies all
// frequency masks (:58-64) and then all time masks (:67-73), so the final value is
//     time-masked(t)  ->  time_means[b,d]
//     freq-masked(d)  ->  freq_means[b,t]
//     otherwise       ->  x[b,t,d]
// Mask rectangles arrive as host-drawn integers (start, end) already resolved with Python
// slice semantics (negative starts, spill into padding rows), so placement is bit-exact.
#include "spl_internal.cuh"

namespace spl {

constexpr int kPostRows = 64;  // rows of one utterance per CTA
constexpr int kPostThreads = 256;

// VEC4: rows are float4-addressable; NG: float4 groups per lane (1: Dm <= 128, 2: Dm <= 256)
template <bool VEC4, int NG>
__global__ void __launch_bounds__(kPostThreads, 5) post_kernel(const PostParams p) {
  __shared__ __align__(16) float s_mean[kMaxDm];
  __shared__ __align__(16) float s_istd[kMaxDm];
  __shared__ __align__(16) float s_tm[kMaxDm];
  __shared__ int s_mask[2 * kMaxMasks];
  __shared__ unsigned char s_rowtm[kPostRows];  // row t0 + r lies inside a time mask
  // blockIdx.z = batch of the call, blockIdx.y = utterance within it (the grid covers the largest batch)
  const PostBatch& bd = p.bd[blockIdx.z];
  const int b = blockIdx.y, t0 = blockIdx.x * kPostRows;
  if (b >= bd.B) return;
  const int T = bd.T;
  if (t0 >= T) return;
  float* const feats = bd.feats;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Dm = p.Dm;
  const bool have_masks = bd.mask_params != nullptr || bd.mask_uniforms != nullptr;
  const int nmask = have_masks ? (p.n_freq + p.n_time) : 0;
  // every global load of the prologue is issued before the first use (one L2 round trip, not three)
  const long long len64 = bd.feat_len[b];
  double s1 = 0.0, s2 = 0.0;
  if (bd.utt_stats && tid < Dm) {
    s1 = bd.utt_stats[((size_t)b * 2 + 0) * Dm + tid];
    s2 = bd.utt_stats[((size_t)b * 2 + 1) * Dm + tid];
  }
  int my_mask = 0;
  float uw = 0.f, us = 0.f;
  if (bd.mask_params) {
    if (tid < 2 * nmask) my_mask = bd.mask_params[(size_t)b * 2 * nmask + tid];
  } else if (tid < nmask) {  // uniforms in the reference's draw order: row 2j = width draw, row 2j+1 = start draw
    uw = bd.mask_uniforms[(size_t)(2 * tid) * bd.B + b];
    us = bd.mask_uniforms[(size_t)(2 * tid + 1) * bd.B + b];
  }
  const int len = (int)len64;

  // rows this CTA has to touch: valid rows, plus padding rows hit by a (spilled) time mask
  if (bd.mask_params) {
    if (tid < 2 * nmask) s_mask[tid] = my_mask;
  } else if (tid < nmask) {
    // sp_layers.py:59-62 / :68-71 in the reference's float32 arithmetic ((W * rand).long(),
    // ((limit - width).float() * rand).long()), then the Python slice semantics of x[b, s:s+w] (:64, :73);
    // same code as the host helper spl_specaug_rects
    const bool is_f = tid < p.n_freq;
    const long long size = is_f ? Dm : T;
    const long long width = (long long)__fmul_rn(is_f ? p.freq_width : p.time_width, uw);
    const long long limit = is_f ? (long long)Dm : len64;
    const long long start = (long long)__fmul_rn(__ll2float_rn(limit - width), us);
    const long long end = start + width;
    long long s_ = start < 0 ? start + size : start, e_ = end < 0 ? end + size : end;
    s_ = s_ < 0 ? 0 : (s_ > size ? size : s_);
    e_ = e_ < 0 ? 0 : (e_ > size ? size : e_);
    s_mask[2 * tid] = (int)s_;
    s_mask[2 * tid + 1] = (int)(e_ > s_ ? e_ : s_);
  }
  __syncthreads();
  int t_hi = len;
  for (int j = p.n_freq; j < nmask; ++j) {
    const int e = s_mask[2 * j + 1];
    t_hi = e > t_hi ? e : t_hi;
  }
  if (t0 >= t_hi) return;
  if (tid < kPostRows) {  // time-mask membership once per row (not once per lane and row)
    bool tmask = false;
    for (int j = p.n_freq; j < nmask; ++j) tmask |= (t0 + tid >= s_mask[2 * j] && t0 + tid < s_mask[2 * j + 1]);
    s_rowtm[tid] = tmask ? 1 : 0;
  }

  if (tid < Dm) {
    const int d = tid;
    float mean = 0.f, istd = 1.f;
    const double inv_len = len > 0 ? 1.0 / (double)len : 0.0;
    const double umean = s1 * inv_len;
    if (p.cmvn_mode == SPL_CMVN_UTTERANCE) {
      double var = s2 * inv_len - umean * umean;
      var = var < 1e-20 ? 1e-20 : var;
      mean = (float)umean;
      istd = p.norm_vars ? (float)rsqrt(var) : 1.f;
    } else if (p.cmvn_mode == SPL_CMVN_GLOBAL) {
      mean = p.global_mean[d];
      istd = p.norm_vars ? p.global_istd[d] : 1.f;
    }
    s_mean[d] = mean;
    s_istd[d] = istd;
    s_tm[d] = (float)((umean - (double)mean) * (double)istd);  // time mean of the normalised features
  }
  __syncthreads();

  const int tend = min(min(t0 + kPostRows, T), t_hi);
  const float inv_d = 1.0f / (float)Dm;
  const bool need_fm = p.n_freq > 0 && nmask > 0;
  constexpr int kWarpsB = kPostThreads / 32;
  constexpr int RB = 4;  // rows in flight per warp: their loads are issued together (one L2 round trip, not four)
  float* const cta_rows = feats + ((size_t)b * T + t0) * Dm;  // row t of this CTA: cta_rows + (t - t0) * Dm
  // frequency masks of this lane's columns, resolved once: bit 4 i + e <-> element e of float4 group lane + 32 i
  [[maybe_unused]] uint32_t fbits = 0;
  if (VEC4 && need_fm) {
    for (int j = 0; j < p.n_freq; ++j) {
      const int f0 = s_mask[2 * j], f1 = s_mask[2 * j + 1];
#pragma unroll
      for (int i = 0; i < NG; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = 4 * (lane + 32 * i) + e;
          if (d >= f0 && d < f1) fbits |= 1u << (4 * i + e);
        }
    }
  }
  for (int tb = t0 + w; tb < tend; tb += RB * kWarpsB) {
    if (VEC4) {
      // Dm % 4 == 0, rows 16-byte aligned: lane handles float4 groups lane (+ 32 when NG == 2)
      const int q = Dm >> 2;
      float4 x[RB][NG];
      bool tm[RB], live[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int t = tb + r * kWarpsB;
        const bool in = t < tend;
        const bool tmask = in && s_rowtm[in ? t - t0 : 0] != 0;
        tm[r] = tmask;
        live[r] = in && !tmask && t < len;  // padding stays exactly 0 (freq means of a zero row are 0)
        const float4* row4 = reinterpret_cast<const float4*>(cta_rows + (t - t0) * Dm);
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          const int g = lane + 32 * i;
          x[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live[r] && g < q) x[r][i] = row4[g];
        }
      }
      float sum[RB];
#pragma unroll
      for (int i = 0; i < NG; ++i) {
        const int g = lane + 32 * i;
        if (g < q) {
          const float4 m = reinterpret_cast<const float4*>(s_mean)[g], sd = reinterpret_cast<const float4*>(s_istd)[g];
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            float4& v = x[r][i];
            v = make_float4((v.x - m.x) * sd.x, (v.y - m.y) * sd.y, (v.z - m.z) * sd.z, (v.w - m.w) * sd.w);
            const float part = (v.x + v.y) + (v.z + v.w);
            sum[r] = i == 0 ? part : sum[r] + part;
          }
        } else {
#pragma unroll
          for (int r = 0; r < RB; ++r)
            if (i == 0) sum[r] = 0.f;
        }
      }
      if (need_fm) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int r = 0; r < RB; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int t = tb + r * kWarpsB;
        float4* row4 = reinterpret_cast<float4*>(cta_rows + (t - t0) * Dm);
        if (tm[r]) {  // may legitimately touch padding rows (reference quirk for len < width)
          for (int g = lane; g < q; g += 32) row4[g] = reinterpret_cast<const float4*>(s_tm)[g];
          continue;
        }
        if (!live[r]) continue;
        const float fm = need_fm ? sum[r] * inv_d : 0.f;
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          const int g = lane + 32 * i;
          if (g < q) {
            float4 v = x[r][i];
            const uint32_t fb = fbits >> (4 * i);
            if (fb & 1u) v.x = fm;
            if (fb & 2u) v.y = fm;
            if (fb & 4u) v.z = fm;
            if (fb & 8u) v.w = fm;
            row4[g] = v;
          }
        }
      }
    } else {
      for (int r = 0; r < RB; ++r) {
        const int t = tb + r * kWarpsB;
        if (t >= tend) break;
        float* row = cta_rows + (t - t0) * Dm;
        const bool tmask = s_rowtm[t - t0] != 0;
        if (tmask) {
          for (int d = lane; d < Dm; d += 32) row[d] = s_tm[d];
          continue;
        }
        if (t >= len) continue;
        float y[kMaxDm / 32];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxDm / 32; ++i) {
          const int d = lane + 32 * i;
          y[i] = 0.f;
          if (d < Dm) {
            y[i] = (row[d] - s_mean[d]) * s_istd[d];
            sum += y[i];
          }
        }
        float fm = 0.f;
        if (need_fm) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          fm = sum * inv_d;
        }
#pragma unroll
        for (int i = 0; i < kMaxDm / 32; ++i) {
          const int d = lane + 32 * i;
          if (d < Dm) {
            float v = y[i];
            for (int j = 0; j < p.n_freq && j < nmask; ++j)
              if (d >= s_mask[2 * j] && d < s_mask[2 * j + 1]) v = fm;
            row[d] = v;
          }
        }
      }
    }
  }
}

cudaError_t launch_post(const PostParams& p, cudaStream_t st) {
  int tmax = 1, utts = 1;
  bool aligned = true;
  for (int k = 0; k < p.nb; ++k) {
    tmax = p.bd[k].T > tmax ? p.bd[k].T : tmax;
    utts = p.bd[k].B > utts ? p.bd[k].B : utts;
    aligned = aligned && (reinterpret_cast<uintptr_t>(p.bd[k].feats) & 15) == 0;
  }
  dim3 grid((tmax + kPostRows - 1) / kPostRows, utts, p.nb);
  const bool vec = (p.Dm & 3) == 0 && p.Dm <= 256 && aligned;
  if (vec && p.Dm <= 128)
    post_kernel<true, 1><<<grid, kPostThreads, 0, st>>>(p);
  else if (vec)
    post_kernel<true, 2><<<grid, kPostThreads, 0, st>>>(p);
  else
    post_kernel<false, 1><<<grid, kPostThreads, 0, st>>>(p);
  return cudaGetLastError();
}

// Column statistics over valid frames for the offline-feature path (sp_layers.py:92-99).
__global__ void __launch_bounds__(256) column_stats_kernel(const float* __restrict__ feats,
                                                           const int64_t* __restrict__ feat_len, int T, int Dm,
                                                           double* __restrict__ utt_stats) {
  const int b = blockIdx.y, t0 = blockIdx.x * kPostRows;
  const int len = (int)feat_len[b];
  const int tend = min(min(t0 + kPostRows, T), len);
  for (int d = threadIdx.x; d < Dm; d += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int t = t0; t < tend; ++t) {
      const double v = (double)feats[((size_t)b * T + t) * Dm + d];
      s1 += v;
      s2 = fma(v, v, s2);
    }
    if (tend > t0) {
      atomicAdd(utt_stats + ((size_t)b * 2 + 0) * Dm + d, s1);
      atomicAdd(utt_stats + ((size_t)b * 2 + 1) * Dm + d, s2);
    }
  }
}

cudaError_t launch_column_stats(const float* feats, const int64_t* feat_len, int B, int T, int Dm,
                                double* utt_stats, cudaStream_t st) {
  dim3 grid((T + kPostRows - 1) / kPostRows, B);
  column_stats_kernel<<<grid, 256, 0, st>>>(feats, feat_len, T, Dm, utt_stats);
  return cudaGetLastError();
}

}  // namespace spl
