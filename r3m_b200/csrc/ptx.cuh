// Thin inline-PTX layer for sm_100a: mbarrier, TMA (tiled + im2col), tcgen05 (alloc / mma / commit /
// ld / fences) and the shared-memory / instruction descriptor encoders.  Everything in csrc/ builds on
// these; nothing here is shape specific.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3m {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Programmatic dependent launch (launch.h): wait until the preceding grid has completed and its memory is visible, then
// let the following grid's CTAs be scheduled behind this one.  First statement of every kernel (the tcgen05 kernels
// run their barrier / TMEM prologue first: it touches no global memory).
#ifndef R3M_PDL_TRIGGER
#define R3M_PDL_TRIGGER 0  // 0: never (dependents start at completion; measured best: -0.8 ms per RN50 step), 1: right after the wait (+0.5 ms), 2: at pdl_done() (+-0)
#endif
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if R3M_PDL_TRIGGER == 1
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// after the kernel's main loop: the CTA has issued all its bulk work
__device__ __forceinline__ void pdl_done() {
#if R3M_PDL_TRIGGER == 2
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in the EVEN CTA of this CTA's pair (cluster scope)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & 0xFEFFFFFFu)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must never hang a (remote, shared) GPU.  On time-out the caller raises the
// kernel's error flag and leaves its role loop; the host turns the flag into an error code.
#ifndef R3M_WAIT_TIMEOUT_NS
#define R3M_WAIT_TIMEOUT_NS 2000000000ull
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > R3M_WAIT_TIMEOUT_NS) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// tiled mode, rank-4 tensor (C, W, H, N): a box starting at signed coordinates (elements outside the tensor read as zero)
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
// bulk tensor store of a rank-4 box (elements outside the tensor are dropped); completes in the caller's bulk group
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_smem, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}
// im2col mode, rank-4 tensor (C, W, H, N): coordinates name the *base pixel* inside the bounding box,
// (off_w, off_h) the filter-tap offset added to it.
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05
// ---------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tf32 operands (fp32 storage), fp32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE MMA of M = 256.  Each CTA holds its
// own 128 rows of A and HALF of the N columns of B in its shared memory and receives its 128 rows of D in its tensor
// memory, so a CTA fills its shared memory with half the B bytes per FLOP.  Only the even CTA (rank 0) issues the MMA;
// both issue TMA loads, whose transaction bytes land on the LEADER's mbarrier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {  // the same warp of BOTH CTAs, same smem offset
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in every CTA of cta_mask once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
constexpr uint64_t kTmaCacheDefault = 0x1000000000000000ull;
// TMA loads of a CTA pair: data into THIS CTA's shared memory, transaction bytes onto the leader's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "l"(kTmaCacheDefault)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                        int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w),
      "r"(h), "r"(n), "h"(off_w), "h"(off_h), "l"(kTmaCacheDefault)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 format): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | base_offset [49,52) | layout [61,64) (2 = SWIZZLE_128B).
__host__ __device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                                  uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// The same with a matrix base offset (bits [49,52)): for operands that do NOT start on a 1024-byte boundary of the 128-byte
// swizzle pattern, base_offset = (start address >> 7) & 7 (the row of the 8-row pattern the matrix starts in).
__host__ __device__ __forceinline__ uint64_t make_smem_desc_sw128_bo(uint32_t smem_addr, uint32_t lbo_bytes,
                                                                     uint32_t sbo_bytes, uint32_t base_offset) {
  return make_smem_desc_sw128(smem_addr, lbo_bytes, sbo_bytes) | (static_cast<uint64_t>(base_offset & 7u) << 49);
}
// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
//   fmt: 0 = f16, 1 = bf16, 2 = tf32; a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t m, uint32_t n, uint32_t a_mn,
                                                  uint32_t b_mn) {
  return (1u << 4)            // c_format = F32
         | (fmt << 7)         // a_format
         | (fmt << 10)        // b_format
         | (a_mn << 15)       // a_major
         | (b_mn << 16)       // b_major
         | ((n >> 3) << 17)   // n_dim
         | ((m >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------------------------------------
// small numeric helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// ---------------------------------------------------------------------------------------------
// order-independent (hence run-to-run deterministic) accumulation of floats
// ---------------------------------------------------------------------------------------------
// A 128-bit fixed-point accumulator (Q64.64: resolution 2^-64 ~ 5e-20, range +-9e18) kept as FOUR signed 32-bit limbs,
// each in its own 64-bit word (kFxWords = 4 words per value): value = sum_k limb[k] * 2^(32 k - 64).  A float converts
// to it exactly (24-bit mantissa: it touches at most two neighbouring limbs; bits below 2^-64 are truncated), limbs are
// added with fire-and-forget 64-bit reductions — no carries at add time (a limb word absorbs 2^31 contributions before
// it could overflow) and no returned values to wait for — and integer addition is associative, so the total does not
// depend on the order in which blocks arrive, unlike fp32 atomics.  The reader folds the limbs (fx_to_float).
constexpr int kFxWords = 4;
__device__ __forceinline__ void fx_red(unsigned long long* p, long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fx_add(unsigned long long* acc, float v) {
  if (v == 0.f || !(fabsf(v) < 9.0e18f)) return;  // zeros, and inf / nan / out of range (never produced by sane inputs)
  int e;
  const float m = frexpf(fabsf(v), &e);                  // |v| = m * 2^e, m in [0.5, 1)
  const long long mant = (long long)(m * 16777216.f);    // 24-bit integer, exact
  const int shift = e - 24 + 64;                         // |v| * 2^64 = mant << shift, shift in (-inf, 103]
  if (shift <= -24) return;                              // below the accumulator's resolution
  const long long sgn = v < 0.f ? -1ll : 1ll;
  if (shift < 0) {
    fx_red(acc, sgn * (mant >> (-shift)));
    return;
  }
  const int k = shift >> 5, r = shift & 31;              // mant << shift = (mant << r) << (32 k): limbs k and k + 1
  const long long wide = mant << r;                      // < 2^55
  const long long low = wide & 0xFFFFFFFFll, high = wide >> 32;
  if (low) fx_red(acc + k, sgn * low);
  if (high) fx_red(acc + k + 1, sgn * high);             // k + 1 <= 3 because |v| < 2^63
}
// exact value of the accumulator rounded to nearest-even fp32; integer arithmetic only (fp64 is slow on this part)
__device__ __forceinline__ float fx_to_float(const unsigned long long* acc) {
  __int128 t = 0;
#pragma unroll
  for (int k = kFxWords - 1; k >= 0; --k) t = (t << 32) + (__int128)(long long)__ldcg(acc + k);
  const bool neg = t < 0;
  const unsigned __int128 mag = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  const unsigned long long hi = (unsigned long long)(mag >> 64), lo = (unsigned long long)mag;
  if ((hi | lo) == 0ull) return 0.f;
  const int lz = hi ? __clzll((long long)hi) : 64 + __clzll((long long)lo);  // leading zeros of the 128-bit magnitude
  unsigned long long top, rest;  // magnitude << lz: top = bits 127..64, rest = bits 63..0
  if (lz == 0) {
    top = hi;
    rest = lo;
  } else if (lz < 64) {
    top = (hi << lz) | (lo >> (64 - lz));
    rest = lo << lz;
  } else {
    top = lz == 64 ? lo : lo << (lz - 64);
    rest = 0ull;
  }
  unsigned int mant = (unsigned int)(top >> 40);  // 24 significant bits
  const bool round_bit = ((top >> 39) & 1ull) != 0ull;
  const bool sticky = ((top & ((1ull << 39) - 1ull)) | rest) != 0ull;
  if (round_bit && (sticky || (mant & 1u))) ++mant;  // may reach 2^24: still exact in fp32
  // bit 127 of the shifted magnitude weighs 2^(63 - lz); it is bit 23 of mant
  const float v = (float)mant * __int_as_float((40 - lz + 127) << 23);
  return neg ? -v : v;
}
// the accumulator as fp64: the top 64 significant bits of the magnitude rounded to 53 (relative error < 2^-52; a
// function of the accumulator's bits only, so as reproducible as fx_to_float).  For the few places where an fp32
// rounding of the sum would be amplified by a cancellation downstream (BatchNorm variance, elementwise.cu).
__device__ __forceinline__ double fx_to_double(const unsigned long long* acc) {
  __int128 t = 0;
#pragma unroll
  for (int k = kFxWords - 1; k >= 0; --k) t = (t << 32) + (__int128)(long long)__ldcg(acc + k);
  const bool neg = t < 0;
  const unsigned __int128 mag = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  const unsigned long long hi = (unsigned long long)(mag >> 64), lo = (unsigned long long)mag;
  if ((hi | lo) == 0ull) return 0.0;
  const int lz = hi ? __clzll((long long)hi) : 64 + __clzll((long long)lo);
  unsigned long long top;  // bits 127..64 of magnitude << lz
  if (lz == 0) top = hi;
  else if (lz < 64) top = (hi << lz) | (lo >> (64 - lz));
  else top = lz == 64 ? lo : lo << (lz - 64);
  // magnitude = top * 2^(64 - lz) accumulator units of 2^-64  ->  value = top * 2^-lz, lz in [0, 127]
  const double v = __ull2double_rn(top) * __longlong_as_double((long long)(1023 - lz) << 52);
  return neg ? -v : v;
}
// mean and biased variance of M values from the exact accumulators of their sum and sum of squares: the division by M
// and the E[x^2] - mean^2 subtraction in fp64, so that |mean| >> std costs ~1e-16 * (mean / std)^2 of the variance
// instead of fp32's 6e-8 * (mean / std)^2 (BatchNorm statistics: elementwise.cu bn_batch_moments, conv_igemm.cu tail)
__device__ __forceinline__ void fx_moments(const unsigned long long* acc_sum, const unsigned long long* acc_sq, int M,
                                           float& mean, float& var) {
  const double inv_m = 1.0 / (double)M;
  const double m = fx_to_double(acc_sum) * inv_m;
  mean = (float)m;
  var = fmaxf((float)(fx_to_double(acc_sq) * inv_m - m * m), 0.f);
}
__device__ __forceinline__ void fx_clear(unsigned long long* acc) {
#pragma unroll
  for (int k = 0; k < kFxWords; ++k) acc[k] = 0ull;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace r3m
