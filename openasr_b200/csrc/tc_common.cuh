// tcgen05 / TMEM helpers (sm_100a): descriptors, MMA issue, TMEM alloc / load, commit.
// Bit layouts follow the PTX ISA "Matrix Descriptor" / "Instruction descriptor" tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS headers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spl {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 bytes, 8-row groups of
// 1024 bytes (stride byte offset), 16-byte chunks XOR-swizzled by (row & 7).  `addr` must be the
// (1024-byte aligned) tile start plus, for a K step inside the 128-byte row, the byte offset of
// that step (32 bytes per K=8 tf32 step).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16)  // leading byte offset (unused for swizzled K-major)
         | ((uint64_t)(1024 >> 4) << 32)                          // stride byte offset between 8-row groups
         | ((uint64_t)1 << 46)                                    // descriptor version (Blackwell)
         | ((uint64_t)2 << 61);                                   // SWIZZLE_128B
}

// K-major SWIZZLE_32B tile: rows of 32 bytes (8 tf32 = one kind::tf32 K step), 8-row groups of
// 256 bytes, the two 16-byte chunks of a row swapped when bit 7 of the byte address is set
// (Swizzle<1,4,3>), i.e. for rows 4..7 of every group.  Tile start 256-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);  // SWIZZLE_32B
}
__device__ __host__ __forceinline__ uint32_t sw32_offset(int r, int c /* 0..7 */) {
  return (uint32_t)(r * 32 + ((((c >> 2) ^ (r >> 2)) & 1) << 4) + ((c & 3) << 2));
}

// byte offset of element (row r, 32-bit column c in [0, 32)) inside a SWIZZLE_128B K-major tile
__device__ __host__ __forceinline__ uint32_t sw128_offset(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ (r & 7)) & 7) << 4) + ((c & 3) << 2));
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n
__device__ __host__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// executed by ONE full warp; writes the TMEM base address to *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "power of two >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// bounded mbarrier wait: returns false on timeout instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t max_spins = 1u << 22) {
  const uint32_t addr = smem_addr(bar);
  for (uint32_t i = 0; i < max_spins; ++i) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}

}  // namespace tc
}  // namespace spl
