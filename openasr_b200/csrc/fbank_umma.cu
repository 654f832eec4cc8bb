// Kernel A, default engine: the spectral stage as a folded real DFT on the 5th-generation tensor cores
// (tcgen05.mma kind::f16, accumulators in TMEM), warp specialised, persistent, multi-batch.
//
// Path: src/third_party/kaldi_signal.py:163-211 (dither -> DC removal -> pre-emphasis -> window -> pad) and
// :510-552 (rfft -> power -> mel -> log) + the pad/stack loop of src/blocks/sp_layers.py:81-91, for up to
// kMaxBatches padded batches per launch.
//
// Arithmetic (validated on the CPU by tools/emulate_umma.py against the oracle, max |d| 5e-4 on log-mel):
//   * a frame's origin is moved to the 16-byte floor of its first sample (shift h = 0..3 floats / 0..7 int16): the
//     window table of shift h is the window delayed by h, the power spectrum does not change, and every
//     shared-memory load of the samples is an aligned 128-bit load;
//   * x~ = s (x - mu_c) + s d g : pivot mu_c (mean of the warp's sample span) and power-of-two scale s per
//     producer warp (8 frames), noise g (kaldi_signal.py:176-177) from Philox4x32-7, eight 16-bit uniforms per call;
//   * z'_j = w_j (x~_j - c x~_{j-1})   (x~_{-1} := x~_0), WITHOUT the mean: with N = padded size, NB = N/2, HALF = N/4
//       a_j = z'_j + z'_{N-j},  b_j = z'_j - z'_{N-j}   (j = 0 .. NB-1)
//       Ce_k = sum_{j even} a_j cos(2 pi jk/N)   Co_k = sum_{j odd} a_j cos(..)     k = 1 .. HALF
//       Se_k = sum_{j even} b_j sin(2 pi jk/N)   So_k = sum_{j odd} b_j sin(..)
//       X_k = (Ce + Co) - i (Se + So),   X_{NB-k} = (Ce - Co) - i (So - Se)
//     four [128 frames x HALF] accumulators = all 512 TMEM columns at 16 kHz;
//   * both operands are split into FP16 hi + lo (3 products, fp32 accumulation): 22 mantissa bits, K = 16 per MMA;
//   * the DC removal x - mean(x) (mean of the NOISY frame) and the Nyquist-position sample z'_{NB} are folded
//     into the GEMM as one extra K step: the A row holds -(1-c) mean(x~) (hi, lo) in the slots of its shift and
//     z'_{NB} (hi, lo); the B image holds the DFT of the delayed window and (-1)^k.  No pre-pass over the frame is
//     needed: the mean is accumulated while the operand is produced;
//   * epilogue, thread per frame: TMEM -> power of bins 1..NB-1 in ascending order (two passes over the
//     accumulators) -> streaming triangular mel bank (two running filters, weights from the reference's dense
//     bank) -> log -> 16-column pieces through shared memory -> coalesced stores + fp64 column sums.
//
// Roles (one CTA of 19 warps per SM): warps 0-15 are the workers -- they produce the A operand of a tile (warp w
// owns rows 8w..8w+7, lane = (row, octet of 8 samples)) and then run its epilogue together: warp w reads TMEM lane
// quarter w % 4 and takes part w / 4 of the bins (the two filters a part boundary cuts are handed to the part
// below through a small side buffer); the accumulators are single-buffered (4 x HALF = all 512 TMEM columns), so
// an epilogue on few warps would serialise with the MMAs -- measured 47 k of 65 k cycles per tile with four
// dedicated epilogue warps (profiles/r2_summary.md).  Warp 16 issues the MMAs; warp 17 streams the pre-swizzled
// twiddle images (TMA bulk copies, 3-deep ring); warp 18 builds the next tile (row table, sample staging by TMA)
// while the workers are in the epilogue.
#include <cuda_fp16.h>

#include "fbank_frame.cuh"
#include "tc_common.cuh"

namespace spl {

#ifdef SPL_TRACE  // per-warp clock64 timeline of CTA 0 (tools/trace_umma.py); never defined in the shipped library
__device__ unsigned long long g_utrace[23 * 512];
#define UTR(tag)                                                                                                  \
  do {                                                                                                            \
    if (blockIdx.x == 0 && lane == 0 && tr_n < 510)                                                               \
      g_utrace[w * 512 + tr_n++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)clock64() & 0xffffffffffffULL); \
  } while (0)
#else
#define UTR(tag) do { } while (0)
#endif

namespace {

constexpr int kURows = 128;
constexpr int kUProd = 16;                  // producer warps
constexpr int kUWarps = kUProd + 3;         // + MMA issuer, twiddle TMA, tile setup
constexpr int kUThreads = kUWarps * 32;
constexpr int kBStages = 3;                 // twiddle half-stages in flight
constexpr int kMaxSeg = 16;                 // utterance segments per tile (more: the tile is cut short)
constexpr unsigned long long kWaitCycles = 1ull << 31;  // ~1 s: a wait that long is a protocol bug, not a stall

// ---------------------------------------------------------------------------------------------
// shared-memory layout (byte offsets from the 1024-byte aligned base)
struct ULayout {
  int a_sub;       // one A sub-tile: 128 rows x 32 B
  int a_stage;     // 8 sub-tiles: ce_hi ce_lo co_hi co_lo se_hi se_lo so_hi so_lo
  int b_tile;      // one twiddle tile: HALF rows x 32 B
  int b_stage;     // half-stage: 4 tiles
  int off_a, off_b, off_samp, samp_bytes, off_tab, off_sb, off_rt, off_seg, off_fpre, off_gst, off_bar, total;
};

// row table of one tile (double buffered)
struct RowTab {
  float* out[kURows];    // destination of the row's features (nullptr: row not valid)
  int off[kURows];       // element index (sample units) of the row's 16-byte floor inside the staging buffer
  int ut[kURows];        // flattened utterance | shift h << 24 ; -1: invalid row
  int t[kURows];         // frame index inside the utterance
  float inv2[kURows];    // 1 / s^2 of the row (written by its producer warp)
  int nrows;             // 0: no more tiles
  int pad_[3];
};

__host__ __device__ inline ULayout make_ulayout(int nfft, int es, int tab_bytes, int D_out) {
  ULayout L;
  const int half = nfft / 4;
  L.a_sub = kURows * 32;
  L.a_stage = 8 * L.a_sub;
  L.b_tile = half * 32;
  L.b_stage = 4 * L.b_tile;
  L.off_a = 0;
  L.off_b = L.off_a + 2 * L.a_stage;
  L.off_samp = L.off_b + kBStages * L.b_stage;
  // staging: a full tile of one utterance ((127 S + Nw) samples at 16 kHz / fp32 = 82 880 B) + alignment heads
  L.samp_bytes = es == 4 ? 86016 : 45056;
  L.off_tab = L.off_samp + L.samp_bytes;
  L.off_sb = (L.off_tab + tab_bytes + 15) & ~15;   // boundary partials [3][128 rows][2]
  L.off_rt = L.off_sb + 3 * kURows * 2 * 4;
  L.off_seg = L.off_rt + 2 * (int)sizeof(RowTab);
  L.off_fpre = L.off_seg + kMaxSeg * 16;
  L.off_gst = (L.off_fpre + (kMaxUmmaUtts + 1) * 4 + 7) & ~7;
  L.off_bar = L.off_gst + 2 * D_out * 8;
  L.total = L.off_bar + 40 * 8;
  return L;
}

// barrier indices
enum : int {
  BAR_TAB = 0,       // tables landed (TMA tx)
  BAR_AFULL = 1,     // [2] producers -> MMA (16 arrivals)
  BAR_AEMPTY = 3,    // [2] MMA commit -> producers
  BAR_BFULL = 5,     // [3] twiddle TMA tx -> MMA
  BAR_BEMPTY = 8,    // [3] MMA commit -> twiddle TMA
  BAR_TFULL = 11,    // accumulators complete (commit) -> workers (epilogue)
  BAR_TEMPTY = 12,   // workers drained TMEM (16 arrivals) -> MMA
  BAR_SFULL = 13,    // samples landed (TMA tx) -> producers
  BAR_SEMPTY = 14,   // producers done with the staging buffer (16 arrivals) -> setup warp
  BAR_READY = 15,    // [2] tile published (1 arrival) -> everyone
  BAR_SB = 17,       // [3 boundaries x 4 quarters] boundary partials written (1 arrival) -> the part below
  BAR_COUNT = 29
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Bounded wait: a protocol error must not hang the GPU.  Returns false on timeout or when another role aborted.
// The first polls are back to back (the common case: the phase is complete or about to be); after that the warp
// sleeps between polls so that waiting roles do not take issue slots from the working ones.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t addr, uint32_t parity, volatile int* abort_flag, int code) {
  unsigned long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_try(addr, parity)) return true;
    __nanosleep(it < 16 ? 32 : 128);
    if ((it & 255u) == 255u) {
      if (*abort_flag) return false;
      const unsigned long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitCycles) {
        if (*abort_flag == 0) *abort_flag = 0x10000 | code;  // which wait gave up first: barrier index | warp << 8
        return false;
      }
    }
  }
}
__device__ __forceinline__ bool mbar_wait_or_abort(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int code = 0) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try(addr, parity)) return true;
  if (mbar_try(addr, parity)) return true;
  return mbar_wait_slow(addr, parity, abort_flag, code);
}

// A wait every worker warp needs: warp 0 polls the mbarrier, the other fifteen block on a hardware named barrier
// (no issue slots) until warp 0 joins it.  16 polling warps cost 37 % of all issued instructions (ncu, profiles/).
__device__ __forceinline__ bool workers_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int code, int w) {
  if (w == 0) mbar_wait_or_abort(bar, parity, abort_flag, code);
  asm volatile("bar.sync 1, 512;" ::: "memory");
  return *abort_flag == 0;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16, FP16 operands, fp32 accumulate, A and B K-major, M = 128, N = n
__device__ __host__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// fp64 add to global memory without a return value (SASS RED): the warp does not wait for the L2 round trip
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {  // (lo -> bits 0..15, hi -> bits 16..31)
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// v[0..3] -> FP16 hi parts (2 words) and FP16 lo parts (2 words) of the hi/lo split
__device__ __forceinline__ void split4(const float (&v)[4], uint2& hi, uint2& lo) {
  const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// eight consecutive staged samples (aligned) as floats
template <typename ST>
__device__ __forceinline__ void load8(const ST* p, float (&x)[8]) {
  if constexpr (sizeof(ST) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = (float)(short)(w[i] & 0xffffu);
      x[2 * i + 1] = (float)((int)w[i] >> 16);
    }
  }
}

// batch of flattened utterance u (uniform loop over <= kMaxBatches descriptors)
__device__ __forceinline__ int batch_of(const UmmaParams& p, int u) {
  int k = 0;
#pragma unroll 1
  for (int i = 1; i < p.nb; ++i)
    if (u >= p.bd[i].u0) k = i;
  return k;
}

}  // namespace

size_t fbank_umma_smem_bytes(int nfft, int es, int tab_bytes, int D_out) {
  return (size_t)make_ulayout(nfft, es, tab_bytes, D_out).total + 1024;  // + alignment slack
}

// ---------------------------------------------------------------------------------------------
// NOISE: 0 none, 1 device RNG (Philox), 2 host stream (parity mode: noise[b][t][Nw] of every batch)
template <int NFFT, typename ST, int NOISE>
__global__ void __launch_bounds__(kUThreads, 1) fbank_umma_kernel(const __grid_constant__ UmmaParams p) {
  constexpr int HALF = NFFT / 4, NB = NFFT / 2, NCH = NB / 32;
  constexpr int TMEM_COLS = 4 * HALF;
  constexpr int ES = (int)sizeof(ST);
  constexpr int NSHIFT = 16 / ES;            // 4 (fp32) or 8 (int16)
  constexpr int KX = (3 * NSHIFT + 2 + 15) / 16;  // K steps of the correction chunk
  constexpr int HS_PER_TILE = 2 * (NCH + 1);
  // the dynamic shared memory window starts 1024-byte aligned; pointers derived from `sm` stay in the shared
  // state space (LDS / STS with 32-bit addresses -- an integer round-up of the base would turn every access
  // into a generic LD / ST with 64-bit address arithmetic)
  extern __shared__ __align__(1024) uint8_t sm[];
  if ((smem_u32(sm) & 255u) != 0u) {  // SWIZZLE_32B tiles need 256-byte alignment
    if (threadIdx.x == 0 && p.status) atomicCAS(p.status, 0, 0x20000);
    return;
  }
  const ULayout L = make_ulayout(NFFT, ES, p.tab_bytes, p.D_out);
  ST* samp = reinterpret_cast<ST*>(sm + L.off_samp);
  const float* tab = reinterpret_cast<const float*>(sm + L.off_tab);
  const float* wtab = tab;                                              // [NSHIFT][NFFT] delayed windows
  const float2* melw = reinterpret_cast<const float2*>(tab + p.off_melw);   // [2 HALF] (w_a, w_b) per step
  const uint32_t* melc = reinterpret_cast<const uint32_t*>(tab + p.off_melc);  // 2 bits per step: shifts before it
  RowTab* rtab = reinterpret_cast<RowTab*>(sm + L.off_rt);
  int4* segs = reinterpret_cast<int4*>(sm + L.off_seg);
  int* fpre = reinterpret_cast<int*>(sm + L.off_fpre);
  double* gst = reinterpret_cast<double*>(sm + L.off_gst);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.off_bar);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tslot + 1);
  int* gcount = reinterpret_cast<int*>(tslot + 2);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int S = p.S, Nw = p.Nw, D_out = p.D_out, U = p.total_utts;
  [[maybe_unused]] int tr_n = 0;
  UTR(0);

  // ---- 0. barriers, TMEM, tables, frame prefix ------------------------------------------------------
  if (tid == 0) {
    mbar_init(bars + BAR_TAB, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + BAR_AFULL + i, kUProd);
      mbar_init(bars + BAR_AEMPTY + i, 1);
      mbar_init(bars + BAR_READY + i, 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(bars + BAR_BFULL + i, 1);
      mbar_init(bars + BAR_BEMPTY + i, 1);
    }
    mbar_init(bars + BAR_TFULL, 1);
    mbar_init(bars + BAR_TEMPTY, kUProd);
    for (int i = 0; i < 12; ++i) mbar_init(bars + BAR_SB + i, 1);
    mbar_init(bars + BAR_SFULL, 1);
    mbar_init(bars + BAR_SEMPTY, kUProd);
    *abort_flag = 0;
    *gcount = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bars + BAR_TAB, (uint32_t)p.tab_bytes);
    bulk_g2s(sm + L.off_tab, p.tab, (uint32_t)p.tab_bytes, bars + BAR_TAB);
  }
  if (w == kUProd) tc::tmem_alloc<TMEM_COLS>(tslot);
  if (w == 0) {  // frames of every flattened utterance, exclusive prefix; all loads issued before the scan
    int carry = 0;
    for (int base = 0; base < U; base += 32) {
      const int u = base + lane;
      int m = 0;
      if (u < U) {
        const int k = batch_of(p, u);
        const UBatch& bd = p.bd[k];
        const long long n = bd.wav_len[u - bd.u0];
        m = n >= Nw ? (int)(1 + (n - Nw) / S) : 0;  // kaldi_signal.py:90
        m = m > bd.T ? bd.T : m;
        if (blockIdx.x == 0 && bd.feat_len) bd.feat_len[u - bd.u0] = m;
      }
      int inc = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (u < U) fpre[u] = carry + inc - m;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) fpre[U] = carry;
  }
  for (int i = tid; i < 2 * D_out; i += kUThreads) gst[i] = 0.0;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tslot;
  const long long total = fpre[U];
  const int r0 = (int)(total * blockIdx.x / gridDim.x), r1 = (int)(total * (blockIdx.x + 1) / gridDim.x);

  // =================================================================================================
  if (w < kUProd) {
    // ---- producers ------------------------------------------------------------------------------
    // zero padding rows (sp_layers.py:88): this CTA's equal share, one sixteenth per warp, written while the
    // first tile's samples are in flight
    {
      long long tot_rows = 0;
      for (int k = 0; k < p.nb; ++k) tot_rows += (long long)p.bd[k].B * p.bd[k].T;
      const long long total_pad = tot_rows - total;
      const long long c0 = total_pad * blockIdx.x / gridDim.x, c1 = total_pad * (blockIdx.x + 1) / gridDim.x;
      long long q = c0 + (c1 - c0) * w / kUProd;
      const long long q1 = c0 + (c1 - c0) * (w + 1) / kUProd;
      if (q < q1) {
        auto rows_before = [&](int u) -> long long {  // output rows of all utterances < u
          long long rb = 0;
          for (int k = 0; k < p.nb; ++k) {
            const int cnt = u - p.bd[k].u0;
            if (cnt > 0) rb += (long long)(cnt < p.bd[k].B ? cnt : p.bd[k].B) * p.bd[k].T;
          }
          return rb;
        };
        int lo = 0, hi = U - 1;  // largest u with pad rows before it <= q
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (rows_before(mid) - fpre[mid] <= q) lo = mid; else hi = mid - 1;
        }
        int u = lo;
        while (q < q1) {
          while (u + 1 < U && rows_before(u + 1) - fpre[u + 1] <= q) ++u;
          const int k = batch_of(p, u);
          const UBatch& bd = p.bd[k];
          const int m_u = fpre[u + 1] - fpre[u];
          const long long ofs = q - (rows_before(u) - fpre[u]);
          long long nrows = (bd.T - m_u) - ofs;
          nrows = nrows > q1 - q ? q1 - q : nrows;
          float* dst = bd.feats + ((size_t)(u - bd.u0) * bd.T + m_u + (size_t)ofs) * D_out;
          const long long nfl = nrows * D_out;
          if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (nfl & 3) == 0) {
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (long long i = lane; i < (nfl >> 2); i += 32) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            for (long long i = lane; i < nfl; i += 32) dst[i] = 0.f;
          }
          q += nrows;
        }
      }
    }
    if (!workers_wait(bars + BAR_TAB, 0, abort_flag, BAR_TAB | (w << 8), w)) goto done;
    {
      const int rl = lane >> 2, q = lane & 3;   // row inside the warp, octet inside the chunk
      const int row = 8 * w + rl;
      const float cpre = p.preemph;
      const float dith = p.dither;
      uint32_t chunk_seq = 0;                   // A stages used so far
      for (int tile = 0;; ++tile) {
        const int slot = tile & 1;
        if (w == 0) mbar_wait_or_abort(bars + BAR_READY + slot, (tile >> 1) & 1, abort_flag, BAR_READY | (w << 8));
        RowTab& rt = rtab[slot];
        if (w == 0 && !*abort_flag && rt.nrows != 0)  // the tile's samples have landed
          mbar_wait_or_abort(bars + BAR_SFULL, tile & 1, abort_flag, BAR_SFULL | (w << 8));
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (*abort_flag) goto done;
        const int nrows = rt.nrows;
        if (nrows == 0) break;
        UTR(1);
        UTR(2);
        const bool rvalid = row < nrows;
        // rows past the tile's end mirror the warp's first row (finite data, results never stored)
        const int rsrc = rvalid ? row : (8 * w < nrows ? 8 * w : 0);
        const int roff = rt.off[rsrc];
        const int ut = rt.ut[rsrc];
        const int h = ut >> 24, uflat = ut & 0xffffff;
        const int tfr = rt.t[rsrc];
        const ST* xrow = samp + roff;
        const float* wrow = wtab + h * NFFT;
        const int hNw = h + Nw;

        // pass 1: max |x| and mean over the warp's sample span -> power-of-two scale, pivot
        float sc, piv;
        {
          const int lo_off = __shfl_sync(0xffffffffu, roff, 0);
          const int wrows = nrows - 8 * w >= 8 ? 8 : (nrows - 8 * w > 0 ? nrows - 8 * w : 1);
          const int hi_off = __shfl_sync(0xffffffffu, roff + hNw, 4 * (wrows - 1));
          const int first = lo_off + __shfl_sync(0xffffffffu, h, 0);  // first real sample of the span
          const int n8 = (hi_off - lo_off + 7) >> 3;  // octets; elements outside [first, hi_off) may be foreign memory
          float mx = 0.f, sm_ = 0.f;
          for (int i = lane; i < n8; i += 32) {
            float v[8];
            load8<ST>(samp + lo_off + 8 * i, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int idx = lo_off + 8 * i + e;
              const float ve = (idx >= first && idx < hi_off) ? v[e] : 0.f;
              mx = fmaxf(mx, fabsf(ve));
              sm_ += ve;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sm_ += __shfl_xor_sync(0xffffffffu, sm_, o);
          }
          piv = sm_ / (float)(hi_off - first);
          if (!(fabsf(piv) <= 3.0e38f) || !p.remove_dc) piv = 0.f;  // NaN / Inf samples: no pivot
          // |a| <= 4 max|x~|, |x~| <= 2 max|x| + 6 |dither|  ->  scaled operand below 2^13
          const float bound = 4.f * (2.f * mx + 6.f * fabsf(dith));
          int e = 13 - (bound > 0.f && bound < 3.0e38f ? (int)ceilf(log2f(bound)) : 0);
          e = e > 60 ? 60 : (e < -60 ? -60 : e);
          sc = __int_as_float((127 + e) << 23);
          if (q == 0 && rvalid) rt.inv2[row] = __int_as_float((127 - 2 * e) << 23);
        }
        UTR(3);
        const float so = -piv * sc;
        float rowsum = 0.f, zhalf = 0.f;
        float carry_j = 0.f, carry_m = 0.f;   // previous chunk's last x~ (lane q = 3) / lowest mirrored x~
        [[maybe_unused]] const float nk = 1.3862943611198906f * sc * sc * dith * dith;   // 2 ln 2 (s d)^2
        [[maybe_unused]] const float nsgn = dith < 0.f ? -1.f : 1.f;
        [[maybe_unused]] const float* nzrow = nullptr;
        if constexpr (NOISE == 2) {
          const UBatch& bd = p.bd[batch_of(p, uflat)];
          nzrow = bd.noise + ((size_t)(uflat - bd.u0) * bd.T + tfr) * Nw - h;  // indexed by shifted position
        }
        // noise for the eight positions pos0 .. pos0+7 (valid: h <= pos < h + Nw), added to x
        auto add_noise = [&](float (&x)[8], int pos0, uint32_t block_id) {
          if constexpr (NOISE == 1) {
            const uint4 r = philox4x32_7(make_uint4(block_id, (uint32_t)tfr, (uint32_t)uflat, 0x5eedu), p.seed_lo, p.seed_hi);
            const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float vf = (float)((i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xffffu)) + 0.5f;  // u = vf 2^-16
              const float a = fmaf(fast_log2(vf), -nk, 16.0001f * nk);  // -2 ln u (s d)^2 > 0
              float cs;
              asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(vf * 9.587379924285257e-05f));  // 2 pi 2^-16
              x[i] = fmaf(fast_sqrt(a) * cs, nsgn, x[i]);
            }
          } else if constexpr (NOISE == 2) {
            const float sd = sc * dith;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int pos = pos0 + i;
              if (pos >= h && pos < hNw) x[i] = fmaf(__ldg(nzrow + pos), sd, x[i]);
            }
          }
          (void)pos0;
          (void)block_id;
        };

#pragma unroll 1
        for (int c = 0; c <= NCH; ++c, ++chunk_seq) {
          const int st = chunk_seq & 1;
          uint8_t* stage = sm + L.off_a + st * L.a_stage;
          if (chunk_seq >= 2 && !workers_wait(bars + BAR_AEMPTY + st, ((chunk_seq >> 1) - 1) & 1, abort_flag, BAR_AEMPTY | (w << 8), w)) goto done;
          UTR(10 + c);
          // byte offset of this lane's 8-byte slice (K columns 4q..4q+3) inside a sub-tile (SWIZZLE_32B)
          const uint32_t aoff = (uint32_t)(row * 32 + ((((q >> 1) ^ (row >> 2)) & 1) << 4) + ((q & 1) << 3));
          if (c < NCH) {
            const int j0 = 32 * c + 8 * q;
            // ---- direct part: positions j0 .. j0+7 ----
            float x[8];
            load8<ST>(xrow + j0, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], sc, so);
            add_noise(x, j0, (uint32_t)(4 * c + q));
            if (c == 0) {  // positions before the frame's first sample: excluded from the mean, replicate x~_h
              float xh = x[0];
#pragma unroll
              for (int i = 1; i < NSHIFT; ++i) xh = (i == h) ? x[i] : xh;
#pragma unroll
              for (int i = 0; i < NSHIFT - 1; ++i)
                if (q == 0 && i < h) {  // replaced by x~_h and taken out of the sum that follows
                  x[i] = xh;
                  rowsum -= xh;
                }
            }
            rowsum += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
            float pv = __shfl_up_sync(0xffffffffu, x[7], 1, 4);
            if (q == 0) pv = (c == 0) ? x[0] : carry_j;
            carry_j = __shfl_sync(0xffffffffu, x[7], 3, 4);
            float zj[8];
            {
              const float4 w0 = *reinterpret_cast<const float4*>(wrow + j0), w1 = *reinterpret_cast<const float4*>(wrow + j0 + 4);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) zj[i] = wv[i] * fmaf(-cpre, i == 0 ? pv : x[i - 1], x[i]);
            }
            // ---- mirrored part: positions N-8-j0 .. N-1-j0 (+ the one above, from the neighbour octet) ----
            const bool mirror = NFFT - 32 * c - 32 < Nw + NSHIFT - 1;  // uniform: some octet of the chunk is live
            float zm[8];  // zm[i] pairs with zj[i]: position N - (j0 + i)
            if (mirror) {
              const int p0 = NFFT - 8 - j0;
              float y[8];
              load8<ST>(xrow + p0, y);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], sc, so);
              add_noise(y, p0, (uint32_t)(NFFT / 8 + 4 * c + q));
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = (p0 + i < hNw) ? y[i] : 0.f;
              rowsum += ((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]));
              float top = __shfl_up_sync(0xffffffffu, y[0], 1, 4);  // position N - j0: lowest of the octet above
              if (q == 0) top = carry_m;
              carry_m = __shfl_sync(0xffffffffu, y[0], 3, 4);
              const float4 w0 = *reinterpret_cast<const float4*>(wrow + p0), w1 = *reinterpret_cast<const float4*>(wrow + p0 + 4);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              // position p0 + e (e = 1..8; e = 8 is `top`) pairs with the direct position j0 + 8 - e
              const float wtop = (j0 > 0) ? wrow[NFFT - j0] : 0.f;
#pragma unroll
              for (int i = 0; i < 7; ++i) zm[7 - i] = wv[i + 1] * fmaf(-cpre, y[i], y[i + 1]);  // position p0 + 1 + i
              zm[0] = (j0 > 0) ? wtop * fmaf(-cpre, y[7], top) : 0.f;                          // position N - j0
              if (c == NCH - 1 && q == 3) zhalf = wv[0] * fmaf(-cpre, x[7], y[0]);  // position NB: prev is x~_{NB-1}
            } else {
              carry_m = 0.f;
            }
            // ---- fold, split, store ----
            float ae[4], ao[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ae[i] = mirror ? zj[2 * i] + zm[2 * i] : zj[2 * i];
              ao[i] = mirror ? zj[2 * i + 1] + zm[2 * i + 1] : zj[2 * i + 1];
            }
            uint2 hi, lo;
            split4(ae, hi, lo);
            *reinterpret_cast<uint2*>(stage + 0 * L.a_sub + aoff) = hi;
            *reinterpret_cast<uint2*>(stage + 1 * L.a_sub + aoff) = lo;
            split4(ao, hi, lo);
            *reinterpret_cast<uint2*>(stage + 2 * L.a_sub + aoff) = hi;
            *reinterpret_cast<uint2*>(stage + 3 * L.a_sub + aoff) = lo;
            if (mirror) {
              float be[4], bo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                be[i] = zj[2 * i] - zm[2 * i];
                bo[i] = zj[2 * i + 1] - zm[2 * i + 1];
              }
              split4(be, hi, lo);
              *reinterpret_cast<uint2*>(stage + 4 * L.a_sub + aoff) = hi;
              *reinterpret_cast<uint2*>(stage + 5 * L.a_sub + aoff) = lo;
              split4(bo, hi, lo);
              *reinterpret_cast<uint2*>(stage + 6 * L.a_sub + aoff) = hi;
              *reinterpret_cast<uint2*>(stage + 7 * L.a_sub + aoff) = lo;
            }
          } else {
            // ---- correction chunk: -(1-c) mean(x~) in the slots of shift h, z'_{NB} in the last two ----
            float rs = rowsum;
            rs += __shfl_xor_sync(0xffffffffu, rs, 1);
            rs += __shfl_xor_sync(0xffffffffu, rs, 2);
            const float g = p.remove_dc ? -(1.f - cpre) * rs / (float)Nw : 0.f;
            const float zh = __shfl_sync(0xffffffffu, zhalf, 3, 4);
            const __half gh = __float2half_rn(g), zhh = __float2half_rn(zh);
            const __half gl = __float2half_rn(g - __half2float(gh)), zhl = __float2half_rn(zh - __half2float(zhh));
#pragma unroll
            for (int kx = 0; kx < KX; ++kx) {
              __half v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int s_ = 16 * kx + 4 * q + i;  // slot
                __half val = __float2half_rn(0.f);
                if (s_ < 3 * NSHIFT) {
                  if (s_ / 3 == h) val = (s_ % 3 == 1) ? gl : gh;
                } else if (s_ == 3 * NSHIFT) val = zhh;
                else if (s_ == 3 * NSHIFT + 1) val = zhl;
                v[i] = val;
              }
              uint2 wv;
              wv.x = (uint32_t)__half_as_ushort(v[0]) | ((uint32_t)__half_as_ushort(v[1]) << 16);
              wv.y = (uint32_t)__half_as_ushort(v[2]) | ((uint32_t)__half_as_ushort(v[3]) << 16);
              *reinterpret_cast<uint2*>(stage + kx * L.a_sub + aoff) = wv;
            }
          }
          fence_proxy_async();  // generic-proxy writes of the A tiles -> visible to the tensor core
          __syncwarp();
          UTR(20 + c);
          if (lane == 0) {
            mbar_arrive(bars + BAR_AFULL + st);
            if (c == NCH - 1) mbar_arrive(bars + BAR_SEMPTY);  // every lane's last read of the staging buffer is done
          }
        }

        // ---- epilogue of this tile, all workers: warp = (TMEM lane quarter eq, bin part pt) --------------------
        if (!workers_wait(bars + BAR_TFULL, tile & 1, abort_flag, BAR_TFULL | (w << 8), w)) goto done;
        tc::fence_after_sync();
        UTR(111);
        {
          const int eq = w & 3, pt = w >> 2;
          const int erow = 32 * eq + lane;
          if (pt < p.nparts) {
            const float inv2 = erow < nrows ? rt.inv2[erow] : 1.f;
            const int my_ut = rt.ut[erow];
            // all 32 rows of this warp valid and of one utterance (the common case): column sums without per-row checks
            const int ut0 = __shfl_sync(0xffffffffu, my_ut, 0);
            const bool one_utt = __all_sync(0xffffffffu, my_ut >= 0 && ((my_ut ^ ut0) & 0xffffff) == 0);
            const bool want_utt = p.want_utt_stats != 0, want_g = p.global_stats != nullptr;
            // staging ring of raw mel energies, 32 column positions x 32 rows: this warp's own 256-byte slices of the
            // (idle) A ring -- rows 8w..8w+7 of the 16 sub-tiles, which no other warp writes and the tensor core only
            // reads after this warp's next arrival.  Two positions per slice, row index XOR-skewed by the position so
            // that both the emit (lane = row, one position) and the flush (lane = 4 rows x 8 positions) are conflict-free
            uint8_t* const stg_base = sm + L.off_a + (w << 8);
            auto stg = [&](int pos, int r) -> float& {
              return *reinterpret_cast<float*>(stg_base + ((pos >> 1) << 12) + ((pos & 1) << 7) + ((r ^ ((pos & 7) << 2)) << 2));
            };
            float* const sbuf = reinterpret_cast<float*>(sm + L.off_sb);
            const int s_beg = p.part_s0[pt], s_end = p.part_s0[pt + 1];  // multiples of 8, balanced by cost on the host
            int outc = pt == 0 ? 0 : p.part_f0[pt] + 2;  // next output column this part finalises
            int pc0 = outc;                              // first column still staged
            int wpos = 0, rpos = 0;                      // ring positions of outc / pc0
            int emitted = 0;
            float accA = 0.f, accB = 0.f;
            // log + store + column sums of the n <= 8 oldest staged columns [pc0, pc0 + n): lane = (4-row group rr, column c)
            auto flush8 = [&](int n) {
              UTR(123);
              __syncwarp();
              const int c = lane & 7, rr = lane >> 3;
              const int cp = (rpos + c) & 31;
              double s1 = 0.0, s2 = 0.0;
              float f1 = 0.f, f2 = 0.f;  // fp32 partial sums over this warp's 32 rows, accumulated in fp64 across warps
              if (c < n) {
                if (one_utt) {
#pragma unroll
                  for (int it = 0; it < 8; ++it) {
                    const int r = 4 * it + rr;
                    const float v = fast_log(fmaxf(stg(cp, r), kEps));  // kaldi_signal.py:540
                    rt.out[32 * eq + r][pc0 + c] = v;
                    f1 += v;
                    f2 = fmaf(v, v, f2);
                  }
                } else {  // utterance boundaries / rows past the tile's end inside this warp: per-row bookkeeping
                  int cu = -1;
                  auto push = [&]() {
                    if (cu >= 0 && want_utt) {
                      const UBatch& bd = p.bd[batch_of(p, cu)];
                      if (bd.utt_stats) {
                        red_add_f64(bd.utt_stats + ((size_t)(cu - bd.u0) * 2 + 0) * D_out + pc0 + c, s1);
                        red_add_f64(bd.utt_stats + ((size_t)(cu - bd.u0) * 2 + 1) * D_out + pc0 + c, s2);
                      }
                    }
                    if (cu >= 0 && want_g) {
                      atomicAdd(gst + pc0 + c, s1);
                      atomicAdd(gst + D_out + pc0 + c, s2);
                    }
                    s1 = s2 = 0.0;
                  };
                  for (int it = 0; it < 8; ++it) {
                    const int r = 4 * it + rr;
                    float* o = rt.out[32 * eq + r];
                    if (o == nullptr) continue;
                    const float v = fast_log(fmaxf(stg(cp, r), kEps));
                    o[pc0 + c] = v;
                    const int u = rt.ut[32 * eq + r] & 0xffffff;
                    if (u != cu) {
                      push();
                      cu = u;
                    }
                    const double d = (double)v;
                    s1 += d;
                    s2 = fma(d, d, s2);
                  }
                  push();
                }
              }
              UTR(124);
              if (one_utt && (want_utt || want_g)) {  // reduce the four row groups, one atomic per column
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                  f1 += __shfl_xor_sync(0xffffffffu, f1, o);
                  f2 += __shfl_xor_sync(0xffffffffu, f2, o);
                }
                s1 = (double)f1;
                s2 = (double)f2;
                if (rr == 0 && c < n) {
                  if (want_utt) {
                    const int cu = ut0 & 0xffffff;
                    const UBatch& bd = p.bd[batch_of(p, cu)];
                    if (bd.utt_stats) {
                      red_add_f64(bd.utt_stats + ((size_t)(cu - bd.u0) * 2 + 0) * D_out + pc0 + c, s1);
                      red_add_f64(bd.utt_stats + ((size_t)(cu - bd.u0) * 2 + 1) * D_out + pc0 + c, s2);
                    }
                  }
                  if (want_g) {
                    atomicAdd(gst + pc0 + c, s1);
                    atomicAdd(gst + D_out + pc0 + c, s2);
                  }
                }
              }
              pc0 += n;
              rpos = (rpos + n) & 31;
              __syncwarp();
              UTR(125);
            };
            auto finalise = [&](float e) {  // stage the raw energy; the log is taken lane-parallel in flush8
              stg(wpos, lane) = e * inv2;
              ++outc;
              wpos = (wpos + 1) & 31;
            };
            // A filter is complete: parts above the first hand their first two (cut by the part boundary) to the
            // part below as raw partial sums; everything else is finalised here
            auto emit = [&]() {
              if (pt > 0 && emitted < 2) {
                sbuf[((pt - 1) * kURows + erow) * 2 + emitted] = accA;
                if (emitted == 1) {
                  __syncwarp();
                  if (lane == 0) mbar_arrive(bars + BAR_SB + (pt - 1) * 4 + eq);
                }
              } else {
                finalise(accA);
              }
              ++emitted;
              accA = accB;
              accB = 0.f;
            };
            const uint32_t tbase = tmem + ((uint32_t)(32 * eq) << 16);
            if (p.debug_acc != nullptr && blockIdx.x == 0 && tile == 0 && pt == 0) {  // diagnostic: raw accumulators + 1/s^2
              for (int c8 = 0; c8 < 4 * HALF; c8 += 8) {
                float v[8];
                tmem_ld8(tbase + c8, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) p.debug_acc[(size_t)erow * (4 * HALF + 1) + c8 + i] = v[i];
              }
              p.debug_acc[(size_t)erow * (4 * HALF + 1) + 4 * HALF] = inv2;
            }
#pragma unroll 1
            for (int step0 = s_beg; step0 < s_end; step0 += 8) {
              UTR(120);
              // pass 1 (step < HALF): bin = step + 1 = column + 1;  pass 2: bin = HALF + i, column = HALF - 1 - i
              const bool second = step0 >= HALF;
              const int c0 = second ? (2 * HALF - 8 - step0) : step0;  // first TMEM column of the 8 read here
              float ce[8], co[8], se[8], so2[8];
              tmem_ld8(tbase + c0, ce);
              tmem_ld8(tbase + HALF + c0, co);
              tmem_ld8(tbase + 2 * HALF + c0, se);
              tmem_ld8(tbase + 3 * HALF + c0, so2);
              // mel weights of the eight steps (broadcast loads, issued while the TMEM loads are in flight)
              float2 wv[8];
              {
                const float4* m4 = reinterpret_cast<const float4*>(melw + step0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 t4 = m4[i];
                  wv[2 * i] = make_float2(t4.x, t4.y);
                  wv[2 * i + 1] = make_float2(t4.z, t4.w);
                }
              }
              const uint32_t ctl = (melc[step0 >> 4] >> ((step0 & 8) << 1)) & 0xffffu;
              tmem_ld_wait();
              UTR(121);
              float pw[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int ci = second ? 7 - i : i;  // column inside the 8 (descending in pass 2)
                const float re = second ? ce[ci] - co[ci] : ce[ci] + co[ci];
                const float im = second ? so2[ci] - se[ci] : se[ci] + so2[ci];
                pw[i] = fmaf(re, re, im * im);
              }
              if (ctl == 0u) {  // no filter ends inside these eight bins
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  accA = fmaf(wv[i].x, pw[i], accA);
                  accB = fmaf(wv[i].y, pw[i], accB);
                }
              } else if ((ctl & 0xaaaau) == 0u && (pt == 0 || emitted >= 2)) {
                // at most one filter ends per bin and none of them is handed down: straight-line, predicated
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const bool pe = ((ctl >> (2 * i)) & 1u) != 0u;  // uniform
                  stg(wpos, lane) = accA * inv2;  // unconditional: the slot is rewritten until a filter really ends
                  wpos = (wpos + (pe ? 1 : 0)) & 31;
                  outc += pe ? 1 : 0;
                  emitted += pe ? 1 : 0;
                  accA = pe ? accB : accA;
                  accB = pe ? 0.f : accB;
                  accA = fmaf(wv[i].x, pw[i], accA);
                  accB = fmaf(wv[i].y, pw[i], accB);
                }
                UTR(122);
#pragma unroll 1
                while (outc - pc0 >= 8) flush8(8);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  uint32_t ns = (ctl >> (2 * i)) & 3u;
                  while (ns) {  // uniform
                    emit();
                    --ns;
                  }
                  accA = fmaf(wv[i].x, pw[i], accA);
                  accB = fmaf(wv[i].y, pw[i], accB);
                }
                UTR(122);
#pragma unroll 1
                while (outc - pc0 >= 8) flush8(8);  // at most 16 emits per iteration (host-checked): the ring holds 24
              }
            }
            // this warp's share of the accumulators is drained: the next tile's MMAs may overwrite them
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + BAR_TEMPTY);
            UTR(130);
            if (pt + 1 < p.nparts) {
              // the two filters cut by the upper boundary: this part's tail + the head sums of the part above
              // (on a timeout the abort flag is set and every worker leaves together at its next common wait)
              mbar_wait_or_abort(bars + BAR_SB + pt * 4 + eq, tile & 1, abort_flag, BAR_SB | (w << 8));
              const float v0 = accA + sbuf[(pt * kURows + erow) * 2 + 0], v1 = accB + sbuf[(pt * kURows + erow) * 2 + 1];
              finalise(v0);
              finalise(v1);
            } else {
              for (int i = 0; i < p.nflush; ++i) emit();
            }
#pragma unroll 1
            while (outc > pc0) flush8(outc - pc0 < 8 ? outc - pc0 : 8);
            if (want_g && w == 0 && lane == 0) atomicAdd(gcount, nrows);
          } else {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + BAR_TEMPTY);
          }
        }
        UTR(131);
      }
    }
  } else if (w == kUProd) {
    // ---- MMA issuer ---------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(HALF);
      const uint32_t a_base = tc::smem_addr(sm + L.off_a), b_base = tc::smem_addr(sm + L.off_b);
      uint32_t chunk_seq = 0, hs_seq = 0;
      for (int tile = 0;; ++tile) {
        const int slot = tile & 1;
        if (!mbar_wait_or_abort(bars + BAR_READY + slot, (tile >> 1) & 1, abort_flag, BAR_READY | (w << 8))) break;
        if (rtab[slot].nrows == 0) break;
        if (tile > 0 && !mbar_wait_or_abort(bars + BAR_TEMPTY, (tile - 1) & 1, abort_flag, BAR_TEMPTY | (w << 8))) break;
        tc::fence_after_sync();
        UTR(40);
        bool ok = true;
        for (int c = 0; c <= NCH && ok; ++c, ++chunk_seq) {
          const int st = chunk_seq & 1;
          if (!mbar_wait_or_abort(bars + BAR_AFULL + st, (chunk_seq >> 1) & 1, abort_flag, BAR_AFULL | (w << 8))) { ok = false; break; }
          UTR(50 + c);
          const uint32_t a_st = a_base + st * L.a_stage;
          const bool mirror = c < NCH && (NFFT - 32 * c - 32 < Nw + NSHIFT - 1);
          for (int hf = 0; hf < 2; ++hf, ++hs_seq) {
            const int bs = hs_seq % kBStages;
            if (!mbar_wait_or_abort(bars + BAR_BFULL + bs, (hs_seq / kBStages) & 1, abort_flag, BAR_BFULL | (w << 8))) { ok = false; break; }
            tc::fence_after_sync();
            const uint32_t b_st = b_base + bs * L.b_stage;
            // descriptors: constant fields | (address >> 4); sub-tiles are reached by integer adds (no carries:
            // shared-memory addresses are below 2^18)
            const uint64_t ad0 = tc::make_desc_sw32(a_st), bd0 = tc::make_desc_sw32(b_st);
            const uint64_t a_step = (uint64_t)(L.a_sub >> 4), b_step = (uint64_t)(L.b_tile >> 4);
            if (c < NCH) {
#pragma unroll
              for (int pp = 0; pp < 2; ++pp) {  // even / odd block of this half (cos: 0, 1; sin: 2, 3)
                const int blk = 2 * hf + pp;
                const int asub = (hf == 1 && !mirror) ? 2 * pp : 2 * blk;  // no mirror: b == a, reuse the cos operand
                const uint64_t ahi = ad0 + a_step * asub, alo = ahi + a_step;
                const uint64_t bhi = bd0 + b_step * (2 * pp), blo = bhi + b_step;
                const uint32_t d = tmem + blk * HALF;
                mma_f16(d, ahi, bhi, idesc, c > 0);
                mma_f16(d, alo, bhi, idesc, 1);
                mma_f16(d, ahi, blo, idesc, 1);
              }
            } else {
#pragma unroll
              for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int kx = 0; kx < KX; ++kx)
                  mma_f16(tmem + (2 * hf + pp) * HALF, ad0 + a_step * kx, bd0 + b_step * (kx * 2 + pp), idesc, 1);
            }
            tc::commit(bars + BAR_BEMPTY + bs);
          }
          if (!ok) break;
          tc::commit(bars + BAR_AEMPTY + st);
          UTR(70 + c);
        }
        if (!ok) break;
        tc::commit(bars + BAR_TFULL);
      }
    }
    __syncwarp();
  } else if (w == kUProd + 1) {
    // ---- twiddle stream: half-stage i of a tile = image i, the same for every tile ------------------
    if (lane == 0) {
      uint32_t hs_seq = 0;
      for (int tile = 0;; ++tile) {
        const int slot = tile & 1;
        if (!mbar_wait_or_abort(bars + BAR_READY + slot, (tile >> 1) & 1, abort_flag, BAR_READY | (w << 8))) break;
        if (rtab[slot].nrows == 0) break;
        bool ok = true;
        for (int i = 0; i < HS_PER_TILE; ++i, ++hs_seq) {
          const int bs = hs_seq % kBStages;
          if (hs_seq >= kBStages && !mbar_wait_or_abort(bars + BAR_BEMPTY + bs, (hs_seq / kBStages - 1) & 1, abort_flag, BAR_BEMPTY | (w << 8))) { ok = false; break; }
          UTR(90);
          mbar_expect_tx(bars + BAR_BFULL + bs, (uint32_t)L.b_stage);
          bulk_g2s(sm + L.off_b + bs * L.b_stage, p.twiddles + (size_t)i * L.b_stage, (uint32_t)L.b_stage, bars + BAR_BFULL + bs);
        }
        if (!ok) break;
      }
    }
    __syncwarp();
  } else if (w == kUProd + 2) {
    // ---- tile setup: row table, sample staging ---------------------------------------------------------
    int pos = r0;
    int u_hint = 0;
    {
      int lo = 0, hi = U - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (fpre[mid] <= pos) lo = mid; else hi = mid - 1;
      }
      u_hint = lo;
    }
    const int cap = L.samp_bytes / ES - 16;  // elements; the pass over a warp's span reads up to 7 past its end
    for (int tile = 0;; ++tile) {
      const int slot = tile & 1;
      RowTab& rt = rtab[slot];
      int nrows = 0, nseg = 0;
      if (pos < r1) {
        // lane 0 walks the utterance segments of the tile
        if (lane == 0) {
          int u = u_hint, used = 0, left = r1 - pos < kURows ? r1 - pos : kURows;
          int g = pos;
          while (left > 0 && nseg < kMaxSeg) {
            while (fpre[u + 1] <= g) ++u;
            const int k = batch_of(p, u);
            const UBatch& bd = p.bd[k];
            const int t0 = g - fpre[u];
            int n = fpre[u + 1] - g;
            n = n < left ? n : left;
            const char* src = static_cast<const char*>(bd.wav) + ((size_t)(u - bd.u0) * bd.wav_pitch + (size_t)t0 * S) * ES;
            const int head = (int)((reinterpret_cast<uintptr_t>(src) & 15) / ES);
            // rows that fit in what is left of the staging buffer
            const int room = cap - used - head - Nw;
            if (room < 0) break;
            const int fit = room / S + 1;
            if (n > fit) n = fit;
            const int span = head + (n - 1) * S + Nw;
            const int span16 = ((span * ES + 15) & ~15) / ES;
            segs[nseg] = make_int4(u, t0, n, used);
            used += span16;
            nrows += n;
            left -= n;
            g += n;
            ++nseg;
            if (n == fit && left > 0 && fpre[u + 1] > g) break;  // the buffer is full inside this utterance
          }
          u_hint = u;
        }
        nrows = __shfl_sync(0xffffffffu, nrows, 0);
        nseg = __shfl_sync(0xffffffffu, nseg, 0);
        u_hint = __shfl_sync(0xffffffffu, u_hint, 0);
      }
      UTR(100);
      if (tile > 0 && !mbar_wait_or_abort(bars + BAR_SEMPTY, (tile - 1) & 1, abort_flag, BAR_SEMPTY | (w << 8))) break;
      __syncwarp();
      UTR(101);
      if (nrows > 0) {
        // stage the segments: TMA bulk copy from the 16-byte floor, or warp loads when the envelope would leave the batch
        uint32_t tx = 0;
        for (int s_ = 0; s_ < nseg; ++s_) {
          const int4 sg = segs[s_];
          const UBatch& bd = p.bd[batch_of(p, sg.x)];
          const char* wav_lo = static_cast<const char*>(bd.wav);
          const char* wav_hi = wav_lo + ((size_t)(bd.B - 1) * bd.wav_pitch + (size_t)bd.wav_cols) * ES;
          const char* src = wav_lo + ((size_t)(sg.x - bd.u0) * bd.wav_pitch + (size_t)sg.y * S) * ES;
          const char* a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15);
          const int head = (int)((src - a0) / ES);
          const uint32_t bytes = (uint32_t)(((head + (sg.z - 1) * S + Nw) * ES + 15) & ~15);
          ST* dst = samp + sg.w;
          if (a0 >= wav_lo && a0 + bytes <= wav_hi) {
            if (lane == 0) {
              fence_proxy_async();
              bulk_g2s(dst, a0, bytes, bars + BAR_SFULL);
            }
            tx += bytes;
          } else {
            const ST* s0 = reinterpret_cast<const ST*>(src);
            const int need = (sg.z - 1) * S + Nw;
            for (int i = lane; i < head; i += 32) dst[i] = (ST)0;
            for (int i = lane; i < need; i += 32) dst[head + i] = __ldg(s0 + i);
            for (int i = head + need + lane; i < (int)(bytes / ES); i += 32) dst[i] = (ST)0;
          }
        }
        if (lane == 0) mbar_expect_tx(bars + BAR_SFULL, tx);  // tx == 0: plain arrival
        UTR(102);
        // row table while the copies are in flight
        for (int r = lane; r < kURows; r += 32) {
          int ut = -1, off = 0, tt = 0;
          float* out = nullptr;
          if (r < nrows) {
            int acc = 0, s_ = 0;
            while (acc + segs[s_].z <= r) acc += segs[s_++].z;
            const int4 sg = segs[s_];
            const UBatch& bd = p.bd[batch_of(p, sg.x)];
            const char* src = static_cast<const char*>(bd.wav) + ((size_t)(sg.x - bd.u0) * bd.wav_pitch + (size_t)sg.y * S) * ES;
            const int head = (int)((reinterpret_cast<uintptr_t>(src) & 15) / ES);
            const int e0 = sg.w + head + (r - acc) * S;  // first sample of the row
            off = e0 & ~(16 / ES - 1);
            ut = sg.x | ((e0 - off) << 24);
            tt = sg.y + (r - acc);
            out = bd.feats + ((size_t)(sg.x - bd.u0) * bd.T + tt) * D_out;
          }
          rt.out[r] = out;
          rt.off[r] = off;
          rt.ut[r] = ut;
          rt.t[r] = tt;
        }
        __syncwarp();
      }
      if (lane == 0) {
        rt.nrows = nrows;
        __threadfence_block();
        mbar_arrive(bars + BAR_READY + slot);
      }
      UTR(103);
      __syncwarp();
      if (nrows == 0) break;
      pos += nrows;
    }
  }

done:
  UTR(200);
#ifdef SPL_TRACE
  if (blockIdx.x == 0 && lane == 0) g_utrace[w * 512 + 511] = (unsigned long long)tr_n;
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (p.global_stats != nullptr && !*abort_flag) {
    for (int i = tid; i < 2 * D_out; i += kUThreads)
      if (gst[i] != 0.0) atomicAdd(p.global_stats + i, gst[i]);
    if (tid == 0 && *gcount) atomicAdd(p.global_stats + 2 * D_out, (double)*gcount);
  }
  if (*abort_flag && tid == 0 && p.status) atomicCAS(p.status, 0, *abort_flag);
  if (w == kUProd) tc::tmem_dealloc<TMEM_COLS>(tmem);
}

#ifdef SPL_TRACE
extern "C" __attribute__((visibility("default"))) int spl_debug_utrace(unsigned long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_utrace, sizeof(unsigned long long) * (size_t)n);
}
#endif

// ---------------------------------------------------------------------------------------------
template <int NFFT, typename ST, int NOISE>
static cudaError_t launch_uT(const UmmaParams& p, int num_ctas, cudaStream_t st) {
  const size_t smem = fbank_umma_smem_bytes(NFFT, (int)sizeof(ST), p.tab_bytes, p.D_out);
  static thread_local size_t configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || configured[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(fbank_umma_kernel<NFFT, ST, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = smem;
  }
  fbank_umma_kernel<NFFT, ST, NOISE><<<num_ctas, kUThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int NFFT, typename ST>
static cudaError_t launch_uN(const UmmaParams& p, int noise_mode, int num_ctas, cudaStream_t st) {
  if (noise_mode == 0) return launch_uT<NFFT, ST, 0>(p, num_ctas, st);
  if (noise_mode == 1) return launch_uT<NFFT, ST, 1>(p, num_ctas, st);
  return launch_uT<NFFT, ST, 2>(p, num_ctas, st);
}

cudaError_t launch_fbank_umma(const UmmaParams& p, int nfft, int sample_format, int noise_mode, int num_ctas, cudaStream_t st) {
  if (nfft == 512)
    return sample_format == SPL_SAMPLES_F32 ? launch_uN<512, float>(p, noise_mode, num_ctas, st)
                                            : launch_uN<512, int16_t>(p, noise_mode, num_ctas, st);
  return sample_format == SPL_SAMPLES_F32 ? launch_uN<256, float>(p, noise_mode, num_ctas, st)
                                          : launch_uN<256, int16_t>(p, noise_mode, num_ctas, st);
}

}  // namespace spl
