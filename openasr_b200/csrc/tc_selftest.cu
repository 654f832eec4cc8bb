// tcgen05 building-block self-test: D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32,
// A and B staged K-major in the canonical SWIZZLE_128B shared-memory layout, accumulator in TMEM.
// Used by tests/test_gpu_tcgen05.py to pin the descriptor encodings the DFT-as-GEMM spectral
// stage relies on.  One CTA of 128 threads.
#include "fbank_frame.cuh"
#include "tc_common.cuh"

namespace spl {

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int N, int K, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // [A atoms: K/32 x (128 rows x 128 B)] [B atoms: K/32 x (N rows x 128 B)] [mbarrier] [tmem slot]
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int atoms = K / 32;
  uint8_t* sA = base;
  uint8_t* sB = sA + (size_t)atoms * 128 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)atoms * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, w = tid >> 5;

  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<float*>(sA + (size_t)(k >> 5) * 128 * 128 + tc::sw128_offset(r, k & 31)) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<float*>(sB + (size_t)(k >> 5) * N * 128 + tc::sw128_offset(r, k & 31)) = B[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 0) tc::tmem_alloc<256>(slot);
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;

  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(N);
    uint32_t acc = 0;
    for (int a = 0; a < atoms; ++a)
      for (int ks = 0; ks < 4; ++ks) {  // K = 8 tf32 = 32 bytes per MMA
        const uint64_t ad = tc::make_desc_sw128(tc::smem_addr(sA + (size_t)a * 128 * 128) + ks * 32);
        const uint64_t bd = tc::make_desc_sw128(tc::smem_addr(sB + (size_t)a * N * 128) + ks * 32);
        tc::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    tc::commit(bar);
  }
  const bool ok = tc::mbar_wait_bounded(bar, 0);
  tc::fence_after_sync();
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(32 * w) << 16) + c0, v);
      const int row = 32 * w + (tid & 31);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc<256>(tmem);
}

// Same GEMM with SWIZZLE_32B operand tiles: one [rows x 8 tf32] tile per K step (the layout the
// fused DFT kernel uses for its 8-wide K chunks).
__global__ void __launch_bounds__(128, 1) tc_selftest_sw32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                  float* __restrict__ D, int N, int K,
                                                                  int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int steps = K / 8;
  uint8_t* sA = base;                              // steps x (128 rows x 32 B)
  uint8_t* sB = sA + (size_t)steps * 128 * 32;     // steps x (N rows x 32 B)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)steps * N * 32);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, w = tid >> 5;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<float*>(sA + (size_t)(k >> 3) * 128 * 32 + tc::sw32_offset(r, k & 7)) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i - r * K;
    *reinterpret_cast<float*>(sB + (size_t)(k >> 3) * N * 32 + tc::sw32_offset(r, k & 7)) = B[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 0) tc::tmem_alloc<256>(slot);
  fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(N);
    for (int s = 0; s < steps; ++s)
      tc::mma_tf32(tmem, tc::make_desc_sw32(tc::smem_addr(sA + (size_t)s * 128 * 32)),
                   tc::make_desc_sw32(tc::smem_addr(sB + (size_t)s * N * 32)), idesc, s > 0);
    tc::commit(bar);
  }
  const bool ok = tc::mbar_wait_bounded(bar, 0);
  tc::fence_after_sync();
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(32 * w) << 16) + c0, v);
      const int row = 32 * w + (tid & 31);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc<256>(tmem);
}

cudaError_t launch_tc_selftest_sw32(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)(K / 8) * (128 + N) * 32 + 64;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_sw32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_selftest_sw32_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, status);
  return cudaGetLastError();
}

cudaError_t launch_tc_selftest(const float* A, const float* B, float* D, int N, int K, int* status, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)(K / 32) * (128 + N) * 128 + 64;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_selftest_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, status);
  return cudaGetLastError();
}

}  // namespace spl
