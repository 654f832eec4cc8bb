// Kernel A, tcgen05 variant: the spectral stage as a DFT-as-GEMM on the 5th-generation tensor cores.
//
// Same path as fbank_warp.cu (kaldi_signal.py:163-211 + :510-552), different spectral engine:
//   X_k = sum_j z_j e^{-2 pi i jk/N}  for a tile of 128 frames at once, as four real GEMMs.
// Folding the zero-padded frame with the symmetries j <-> N-j and k <-> N/2-k leaves four
// HALF x HALF blocks (HALF = N/4):
//   a_j = z_j + z_{N-j},  b_j = z_j - z_{N-j}   (j = 0 .. N/2-1;  z_j = 0 for j >= Nw)
//   Ce_k = sum_{j even} a_j cos(2 pi jk/N)   Co_k = sum_{j odd} a_j cos(..)    k = 1 .. HALF
//   Se_k = sum_{j even} b_j sin(2 pi jk/N)   So_k = sum_{j odd} b_j sin(..)
//   Re X_k = Ce + Co + (-1)^k z_{N/2},  Re X_{N/2-k} = Ce - Co + (-1)^k z_{N/2}
//   Im X_k = -(Se + So),                Im X_{N/2-k} = -(So - Se)
// Precision: both operands are split into TF32 hi + lo (3 products hi*hi + lo*hi + hi*lo, fp32
// accumulation in TMEM); a single TF32/BF16 pass is ~1000x outside the parity tolerance.
//
// One CTA (16 warps) per SM, tile = 128 rows (frames, any utterance mix) of the CTA's share of the
// flattened frame list.  K loop over 8-wide chunks of the folded index: all warps build the A
// operand (pre-processing + fold + hi/lo split) of chunk u into one of two 64 KB stages while the
// tensor core consumes chunk u-1; the twiddle operand arrives per chunk by one TMA bulk copy of a
// pre-swizzled image.  Epilogue: TMEM -> registers -> power -> shared memory -> mel -> log -> store.
#include "fbank_frame.cuh"
#include "tc_common.cuh"

namespace spl {

constexpr int kTcWarps = 16;                       // producer / epilogue warps
constexpr int kTcThreads = (kTcWarps + 1) * 32;    // + one warp whose lane 0 issues TMA and MMAs
constexpr int kTcRows = 128;

struct TLayout {  // byte offsets
  int stage_bytes, a_tile, b_tile, region;  // region = max(2 stages, power arrays)
  int pp;                                   // power-row pitch (floats), == 4 (mod 32)
  int off_out, op, off_tab, off_rows, off_fpre, off_bar, total;
};

__host__ __device__ inline TLayout make_tlayout(int nfft, int D_out, int tab_words) {
  TLayout L;
  const int half = nfft / 4;
  L.a_tile = kTcRows * 32;  // 128 rows x 8 tf32
  L.b_tile = half * 32;
  L.stage_bytes = 8 * L.a_tile + 8 * L.b_tile;
  L.pp = half + 4;
  const int pbytes = 2 * kTcRows * L.pp * 4;
  L.region = 2 * L.stage_bytes > pbytes ? 2 * L.stage_bytes : pbytes;
  L.region = (L.region + 1023) & ~1023;
  L.off_out = L.region;
  L.op = D_out | 1;
  L.off_tab = (L.off_out + kTcRows * L.op * 4 + 15) & ~15;
  L.off_rows = L.off_tab + tab_words * 4;
  // per-row arrays: mu'[128] z_half[128] energy[128] sumg[128] (float) | src offset[128] (int64) | b[128] t[128] (int)
  L.off_fpre = L.off_rows + 128 * 4 * 4 + 128 * 8 + 128 * 4 * 2;
  L.off_bar = (L.off_fpre + (kMaxPersistentB + 1) * 4 + 7) & ~7;
  L.total = L.off_bar + 10 * 8;
  return L;
}

size_t fbank_tc_smem_bytes(int nfft, int D_out, int tc_tab_words) {
  return (size_t)make_tlayout(nfft, D_out, tc_tab_words).total + 1024;  // + alignment slack
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
template <int NFFT, bool NOISE>
__global__ void __launch_bounds__(kTcThreads, 1) fbank_tc_kernel(const FbankParams p, int* __restrict__ dbg) {
  constexpr int HALF = NFFT / 4, UNITS = NFFT / 32, NB = NFFT / 2;
  constexpr int TMEM_COLS = 4 * HALF;  // 512 (16 kHz) or 256 (8 kHz)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.S, Nw = p.Nw, D_out = p.D_out, B = p.B, T = p.T;
  const TLayout L = make_tlayout(NFFT, D_out, p.tab.tc_tab_words);
  float* parr = reinterpret_cast<float*>(sm);  // power arrays alias the stages: [2][128][pp]
  float* out_stage = reinterpret_cast<float*>(sm + L.off_out);
  const float* tab = reinterpret_cast<const float*>(sm + L.off_tab);
  const float4* segw = reinterpret_cast<const float4*>(tab);
  const uint2* segd = reinterpret_cast<const uint2*>(tab + p.tab.tc_off_desc);
  const float* win = tab + p.tab.tc_off_win;
  const float* wct = tab + p.tab.tc_off_wc;
  const float* wst = tab + p.tab.tc_off_ws;
  float* r_mu = reinterpret_cast<float*>(sm + L.off_rows);
  float* r_zh = r_mu + 128;
  float* r_en = r_zh + 128;
  float* r_sg = r_en + 128;
  long long* r_src = reinterpret_cast<long long*>(r_sg + 128);
  int* r_b = reinterpret_cast<int*>(r_src + 128);
  int* r_t = r_b + 128;
  int* fpre = reinterpret_cast<int*>(sm + L.off_fpre);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.off_bar);  // [0] tables, [1..2] B landed, [3..4] MMAs done, [5..6] A written
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* wav = static_cast<const float*>(p.wav);
  const float c = p.preemph;

  // ---- 0. tables (TMA), frame prefix (warp 0), TMEM (warp 1), barriers --------------------------
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
    mbar_init(bars + 5, kTcWarps);
    mbar_init(bars + 6, kTcWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bars, (uint32_t)p.tab.tc_tab_words * 4u);
    bulk_g2s(sm + L.off_tab, p.tab.tc_tab, (uint32_t)p.tab.tc_tab_words * 4u, bars);
  }
  if (w == 1) tc::tmem_alloc<TMEM_COLS>(tslot);
  if (w == 0) {
    int fcarry = 0;
    for (int base = 0; base < B; base += 32) {
      const int bb = base + lane;
      int m = 0;
      if (bb < B) {
        const long long n = p.wav_len[bb];
        m = n >= Nw ? (int)(1 + (n - Nw) / S) : 0;
        m = m > T ? T : m;
        if (blockIdx.x == 0 && p.feat_len) p.feat_len[bb] = m;
      }
      int finc = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, finc, o);
        if (lane >= o) finc += u;
      }
      if (bb < B) fpre[bb] = fcarry + finc - m;
      fcarry += __shfl_sync(0xffffffffu, finc, 31);
    }
    if (lane == 0) fpre[B] = fcarry;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  const long long total = fpre[B];
  const int r0 = (int)(total * blockIdx.x / gridDim.x), r1 = (int)(total * (blockIdx.x + 1) / gridDim.x);

  // ---- 0b. zero padding rows (sp_layers.py:88): equal share per CTA ------------------------------
  {
    const long long total_pad = (long long)B * T - total;
    long long q = total_pad * blockIdx.x / gridDim.x;
    const long long q1 = total_pad * (blockIdx.x + 1) / gridDim.x;
    if (q < q1) {
      int lo = 0, hi = B - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)mid * T - fpre[mid] <= q) lo = mid; else hi = mid - 1;
      }
      int b = lo;
      while (q < q1) {
        while ((long long)(b + 1) * T - fpre[b + 1] <= q) ++b;
        const int m_b = fpre[b + 1] - fpre[b];
        const long long ofs = q - ((long long)b * T - fpre[b]);
        long long nrows = (T - m_b) - ofs;
        nrows = nrows > q1 - q ? q1 - q : nrows;
        float* dst = p.feats + ((size_t)b * T + m_b + (size_t)ofs) * D_out;
        for (long long i = tid; i < nrows * D_out; i += kTcThreads) dst[i] = 0.f;
        q += nrows;
      }
    }
  }
  mbar_wait(bars, 0);  // tables have landed

  uint32_t use_cnt[2] = {0, 0};  // completed uses of each stage (parity tracking)
  bool failed = false;

  for (int pos = r0; pos < r1; pos += kTcRows) {
    const int nrows = r1 - pos < kTcRows ? r1 - pos : kTcRows;

    // ---- 1. row table + per-row mean / energy / z_{N/2} (warp per row).  Pass A locates the rows and
    //      prefetches their samples into L2 (first touch comes from HBM); pass B reduces them. ----------
    if (w < kTcWarps) {
      int bh = 0;
      {
        const int g0 = pos + (w < nrows ? w : 0);
        int lo = 0, hi = B - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (fpre[mid] <= g0) lo = mid; else hi = mid - 1;
        }
        bh = lo;
      }
      for (int r = w; r < kTcRows; r += kTcWarps) {
        const bool valid = r < nrows;
        const int g = pos + (valid ? r : 0);  // invalid rows mirror row 0 (finite data, never stored)
        int b = valid ? bh : 0;
        if (!valid) {
          int lo = 0, hi = B - 1;
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (fpre[mid] <= g) lo = mid; else hi = mid - 1;
          }
          b = lo;
        } else {
          while (fpre[b + 1] <= g) ++b;
          bh = b;
        }
        const int t = g - fpre[b];
        const long long src = (long long)b * p.wav_pitch + (long long)t * S;
        if (32 * lane < Nw) asm volatile("prefetch.global.L2 [%0];" ::"l"(wav + src + 32 * lane));
        if (lane == 0) {
          r_src[r] = src;
          r_b[r] = b;
          r_t[r] = t;
        }
      }
      __syncwarp();
      for (int r = w; r < kTcRows; r += kTcWarps) {
        const float* x = wav + r_src[r];
        float v[16];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = lane + 32 * i;
          v[i] = j < Nw ? __ldg(x + j) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += v[i];
        float mean = 0.f;
        if (p.remove_dc) mean = group_sum(sum, 32) * (1.0f / (float)Nw);
        float e = 0.f;
        if (p.use_energy) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float d = (lane + 32 * i < Nw) ? v[i] - mean : 0.f;
            e = fmaf(d, d, e);
          }
          e = group_sum(e, 32);
        }
        // z_{N/2}: x[N/2] sits in lane 0 (i = N/64), x[N/2 - 1] in lane 31 (i = N/64 - 1)
        const float xh = v[NB / 32], xhm = __shfl_sync(0xffffffffu, v[NB / 32 - 1], 31);
        if (lane == 0) {
          const float mu = (1.0f - c) * mean;
          r_mu[r] = mu;
          r_en[r] = fast_log(fmaxf(e, kEps));
          r_sg[r] = 0.f;
          r_zh[r] = NB < Nw ? win[NB] * (xh - c * xhm - mu) : 0.f;
        }
      }
    }
    __syncthreads();

    // ---- 2. K loop, warp specialised: 16 producer warps build the A operand of chunk u (fold + hi/lo
    //      split, samples prefetched one chunk ahead in registers) and arrive on full[s]; lane 0 of the
    //      17th warp streams the twiddle images (TMA) and issues the MMAs, whose commit frees the stage.
    if (w < kTcWarps) {
      const int i8 = tid & 7;
      const float* xr[2];
      float mur[2];
      uint32_t offr[2];
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int r = (tid >> 3) + it * 64;
        xr[it] = wav + r_src[r];
        mur[it] = r_mu[r];
        offr[it] = tc::sw32_offset(r, i8);
      }
      float nx[2][6];
      auto load_unit = [&](int u, float (&dst)[2][6]) {
        const int j = 16 * u + 2 * i8, m = NFFT - j;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const float* x = xr[it];
          dst[it][0] = __ldg(x + (j > 0 ? j - 1 : 0));
          dst[it][1] = __ldg(x + j);
          dst[it][2] = __ldg(x + j + 1);              // j + 1 <= N/2 - 1 < Nw
          dst[it][3] = (m - 1 < Nw) ? __ldg(x + m - 2) : 0.f;
          dst[it][4] = (m - 1 < Nw) ? __ldg(x + m - 1) : 0.f;
          dst[it][5] = (m < Nw) ? __ldg(x + m) : 0.f;
        }
      };
      load_unit(0, nx);
#pragma unroll 1
      for (int u = 0; u < UNITS; ++u) {
        const int s = u & 1;
        uint8_t* stage = sm + (size_t)s * L.stage_bytes;
        float cur[2][6];
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
          for (int q = 0; q < 6; ++q) cur[it][q] = nx[it][q];
        if (u + 1 < UNITS) load_unit(u + 1, nx);
        const int j = 16 * u + 2 * i8, m = NFFT - j;
        const float wj0 = win[j], wj1 = win[j + 1];
        const float wm0 = m < Nw ? win[m] : 0.f, wm1 = m - 1 < Nw ? win[m - 1] : 0.f;
        if (use_cnt[s] > 0 && !tc::mbar_wait_bounded(bars + 3 + s, (use_cnt[s] - 1) & 1)) failed = true;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const float mu = mur[it];
          // z_j = w_j (x_j - c x_{j-1} - (1-c) mean), x_{-1} := x_0 ; zero beyond the window
          const float zj0 = wj0 * (cur[it][1] - c * cur[it][0] - mu);
          const float zj1 = wj1 * (cur[it][2] - c * cur[it][1] - mu);
          const float zm1 = wm1 * (cur[it][4] - c * cur[it][3] - mu);
          const float zm0 = wm0 * (cur[it][5] - c * cur[it][4] - mu);
          const float v0 = zj0 + zm0, v1 = zj1 + zm1, v2 = zj0 - zm0, v3 = zj1 - zm1;  // ce, co, se, so entries
          uint8_t* dstp = stage + offr[it];
          const float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u);
          const float h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
          const float h2 = __uint_as_float(__float_as_uint(v2) & 0xFFFFE000u);
          const float h3 = __uint_as_float(__float_as_uint(v3) & 0xFFFFE000u);
          *reinterpret_cast<float*>(dstp + 0 * L.a_tile) = h0;
          *reinterpret_cast<float*>(dstp + 1 * L.a_tile) = v0 - h0;
          *reinterpret_cast<float*>(dstp + 2 * L.a_tile) = h1;
          *reinterpret_cast<float*>(dstp + 3 * L.a_tile) = v1 - h1;
          *reinterpret_cast<float*>(dstp + 4 * L.a_tile) = h2;
          *reinterpret_cast<float*>(dstp + 5 * L.a_tile) = v2 - h2;
          *reinterpret_cast<float*>(dstp + 6 * L.a_tile) = h3;
          *reinterpret_cast<float*>(dstp + 7 * L.a_tile) = v3 - h3;
        }
        fence_proxy_async();  // generic-proxy writes of the A tiles -> visible to the tensor core
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_addr(bars + 5 + s)) : "memory");
        use_cnt[s] += 1;
      }
    } else {
      if (lane == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(HALF);
#pragma unroll 1
        for (int u = 0; u < UNITS; ++u) {
          const int s = u & 1;
          uint8_t* stage = sm + (size_t)s * L.stage_bytes;
          if (use_cnt[s] > 0 && !tc::mbar_wait_bounded(bars + 3 + s, (use_cnt[s] - 1) & 1)) failed = true;
          mbar_expect_tx(bars + 1 + s, 8u * (uint32_t)L.b_tile);
          bulk_g2s(stage + 8 * L.a_tile, reinterpret_cast<const uint8_t*>(p.tab.tc_b) + (size_t)u * 8 * L.b_tile,
                   8u * (uint32_t)L.b_tile, bars + 1 + s);
          if (!tc::mbar_wait_bounded(bars + 5 + s, use_cnt[s] & 1)) failed = true;
          if (!tc::mbar_wait_bounded(bars + 1 + s, use_cnt[s] & 1)) failed = true;
          tc::fence_after_sync();
          const uint32_t abase = tc::smem_addr(stage), bbase = abase + 8 * L.a_tile;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const uint64_t ahi = tc::make_desc_sw32(abase + (blk * 2 + 0) * L.a_tile);
            const uint64_t alo = tc::make_desc_sw32(abase + (blk * 2 + 1) * L.a_tile);
            const uint64_t bhi = tc::make_desc_sw32(bbase + (blk * 2 + 0) * L.b_tile);
            const uint64_t blo = tc::make_desc_sw32(bbase + (blk * 2 + 1) * L.b_tile);
            const uint32_t d = tmem + blk * HALF;
            tc::mma_tf32(d, ahi, bhi, idesc, u > 0);
            tc::mma_tf32(d, alo, bhi, idesc, 1);
            tc::mma_tf32(d, ahi, blo, idesc, 1);
          }
          tc::commit(bars + 3 + s);
          use_cnt[s] += 1;
        }
      } else {
        use_cnt[0] += (UNITS + 1) / 2;
        use_cnt[1] += UNITS / 2;
      }
      __syncwarp();
    }
    // all MMAs of the tile complete (commits arrive in order; wait for the last use of both stages)
    for (int s = 0; s < 2; ++s)
      if (use_cnt[s] > 0 && !tc::mbar_wait_bounded(bars + 3 + s, (use_cnt[s] - 1) & 1)) failed = true;
    tc::fence_after_sync();
    __syncthreads();  // every thread is past its last stage access: the power arrays may overwrite them

    // ---- 3. epilogue A: TMEM -> power spectrum -> shared memory -------------------------------------
    if (w < kTcWarps) {
      const int q = w & 3, cpart = w >> 2;  // TMEM lane quarter, column part
      const int row = 32 * q + lane;
      const float zh = r_zh[row];
      float* p1 = parr + (size_t)row * L.pp;
      float* p2 = parr + (size_t)(kTcRows + row) * L.pp;
      constexpr int CPW = HALF / 4;  // columns per warp
#pragma unroll 1
      for (int c0 = cpart * CPW; c0 < (cpart + 1) * CPW; c0 += 16) {
        float ce[16], co[16], se[16], so[16];
        const uint32_t ta = tmem + ((uint32_t)(32 * q) << 16) + c0;
        tmem_ld16(ta, ce);
        tmem_ld16(ta + HALF, co);
        tmem_ld16(ta + 2 * HALF, se);
        tmem_ld16(ta + 3 * HALF, so);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float a[4], bq[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = 4 * g4 + e;
            const float cz = ((c0 + n) & 1) ? zh : -zh;  // (-1)^k z_{N/2}, k = c0 + n + 1
            const float cp = ce[n] + co[n] + cz, cm = ce[n] - co[n] + cz;
            const float sp = se[n] + so[n], smn = so[n] - se[n];
            a[e] = cp * cp + sp * sp;     // bin k
            bq[e] = cm * cm + smn * smn;  // bin N/2 - k
          }
          *reinterpret_cast<float4*>(p1 + c0 + 4 * g4) = make_float4(a[0], a[1], a[2], a[3]);
          *reinterpret_cast<float4*>(p2 + c0 + 4 * g4) = make_float4(bq[0], bq[1], bq[2], bq[3]);
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    // ---- 4. epilogue B: mel segments, lane = row, warp = (row group, segment group) -------------------
    if (w < kTcWarps) {
      const int rg = w & 3, cg = w >> 2;
      const int row = 32 * rg + lane;
      const float4* pr0 = reinterpret_cast<const float4*>(parr + (size_t)row * L.pp);
      const float4* pr1 = reinterpret_cast<const float4*>(parr + (size_t)(kTcRows + row) * L.pp);
      float* orow = out_stage + row * L.op + (p.use_energy ? 1 : 0);
      float acc0 = 0.f, acc1 = 0.f;
      for (int sidx = p.tab.tc_sgrp_beg[cg]; sidx < p.tab.tc_sgrp_beg[cg + 1]; ++sidx) {
        const uint2 d = segd[sidx];
        const float4* pp4 = ((d.x >> 12) & 1 ? pr1 : pr0) + (d.x & 63u);
        const float4* wv = segw + d.y;
        const int n4 = (d.x >> 6) & 63u;
        if (d.x & (1u << 13)) acc0 = acc1 = 0.f;
#pragma unroll 1
        for (int g = 0; g < n4; ++g) {
          const float4 pa = pp4[g], wa = wv[g];
          acc0 = fmaf(pa.x, wa.x, acc0);
          acc1 = fmaf(pa.y, wa.y, acc1);
          acc0 = fmaf(pa.z, wa.z, acc0);
          acc1 = fmaf(pa.w, wa.w, acc1);
        }
        if (d.x & (1u << 14)) orow[d.x >> 16] = fast_log(fmaxf(acc0 + acc1, kEps));  // kaldi_signal.py:540
      }
      if (p.use_energy && cg == 0) out_stage[row * L.op] = r_en[row];
    }
    __syncthreads();

    // ---- 5. store rows + per-utterance column sums ---------------------------------------------------
    for (int r = w; r < nrows && w < kTcWarps; r += kTcWarps) {
      float* dst = p.feats + ((size_t)r_b[r] * T + r_t[r]) * D_out;
      const float* src = out_stage + r * L.op;
      for (int cc = lane; cc < D_out; cc += 32) dst[cc] = src[cc];
    }
    if ((p.utt_stats != nullptr || p.global_stats != nullptr) && tid < 4 * D_out) {
      const int col = tid % D_out, q = tid / D_out;  // column, 32-row group
      const int rbeg = 32 * q, rend = min(32 * q + 32, nrows);
      if (rbeg < rend) {
        double s1 = 0.0, s2 = 0.0;
        int cur_b = r_b[rbeg];
        for (int r = rbeg; r <= rend; ++r) {
          const int b = r < rend ? r_b[r] : -1;
          if (b != cur_b) {
            if (p.utt_stats) {
              atomicAdd(p.utt_stats + ((size_t)cur_b * 2 + 0) * D_out + col, s1);
              atomicAdd(p.utt_stats + ((size_t)cur_b * 2 + 1) * D_out + col, s2);
            }
            if (p.global_stats) {
              atomicAdd(p.global_stats + col, s1);
              atomicAdd(p.global_stats + D_out + col, s2);
            }
            s1 = s2 = 0.0;
            cur_b = b;
          }
          if (r < rend) {
            const double v = (double)out_stage[r * L.op + col];
            s1 += v;
            s2 = fma(v, v, s2);
          }
        }
      }
    }
    if (p.global_stats && tid == 0) atomicAdd(p.global_stats + 2 * D_out, (double)nrows);
    __syncthreads();  // row tables / output stage are rewritten by the next tile
  }

  if (failed && dbg) atomicExch(dbg, 1);
  tc::fence_before_sync();
  __syncthreads();
  if (w == 1) tc::tmem_dealloc<TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------
template <int NFFT, bool NOISE>
static cudaError_t launch_tcT(const FbankParams& p, int num_ctas, cudaStream_t st) {
  const size_t smem = fbank_tc_smem_bytes(NFFT, p.D_out, p.tab.tc_tab_words);
  static thread_local size_t configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(fbank_tc_kernel<NFFT, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = smem;
  }
  fbank_tc_kernel<NFFT, NOISE><<<num_ctas, kTcThreads, smem, st>>>(p, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_fbank_tc(const FbankParams& p, int nfft, bool with_noise, int num_ctas, cudaStream_t st) {
  (void)with_noise;  // dither is not implemented in this variant yet: the caller routes dither != 0 elsewhere
  if (nfft == 512) return launch_tcT<512, false>(p, num_ctas, st);
  return launch_tcT<256, false>(p, num_ctas, st);
}

}  // namespace spl
