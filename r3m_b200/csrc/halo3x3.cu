// 3x3 stride-1 convolution of 64 -> 64 channels (ResNet-50 layer1 conv2: forward and data gradient) with TAP REUSE — and
// the stem (7x7 / stride 2 re-expressed as FOUR VERTICAL taps over the 64-"channel" space-to-depth operand):
// the same kernel with a 19 x 10 halo whose origin is two rows above the tile and taps (0..3, 0).
//
// conv_igemm fetches one 128-pixel x 64-channel im2col tile PER TAP, so every input pixel crosses L2 -> shared memory nine
// times, and at N = 64 a tile carries only ~128 tensor cycles per 16 KB: those launches are bound by the operand fill rate
// of the SM (ncu: 11.7 TB/s of L2 traffic, 28 % tensor pipe).  Here an M tile is a 16 x 8 patch of output pixels:
//   * ONE tiled (not im2col) rank-4 TMA load brings the 18 x 10 x 64-channel input halo (23 KB, zero filled outside the
//     image) as 180 rows of 128 bytes in SWIZZLE_128B order, row index = 10 * patch_row + patch_col;
//   * tap (r, s) is the SAME shared memory seen through a K-major descriptor that starts at row 10 r + s, with one 8-row
//     core group per output row (the 8 output columns: 8 consecutive 128-byte rows) and a group stride of 10 rows
//     (SBO = 1280 bytes): nine shifted views, no further loads;
//   * the whole 64 x 576 filter (72 KB) stays resident in shared memory for the CTA's lifetime.
// L2 -> shared-memory traffic per tile: 23 KB instead of 9 x 16 + 72 KB.  16 x 8 tiles: W must be a multiple of 8 and
// H of 4 (rows past the image are computed and dropped: 56 = 3.5 x 16 wastes 12.5 % of the MMA rows).
// Epilogue as in conv_igemm: TMEM -> bf16 -> swizzled staging -> BatchNorm statistics (registers) + one rank-4 bulk
// tensor store of 4 rows x 8 columns x 64 channels per warp and tile; two warp sets drain alternate tiles.
#include "conv_igemm.cuh"
#include "launch.h"
#include "ptx.cuh"

namespace r3m {

namespace {

constexpr int kTileRows = 16, kTileCols = 8;          // output pixels of an M tile (16 x 8 = 128 GEMM rows)
constexpr int kHaloCols = 10;                              // halo row pitch in 128-byte rows (the descriptors' group stride)
constexpr int kMaxHaloRows = 19;                           // 18 for the 3x3 convs, 19 for the stem's four vertical taps
constexpr int kPatchStride = 24 * 1024;                    // 1024-byte aligned slot
static_assert(kMaxHaloRows * kHaloCols * 128 <= kPatchStride, "halo slot");
constexpr int kStages = 3;
constexpr int kFilterBytes = 9 * 64 * 128;                 // nine K blocks of 64 filters x 64 channels
constexpr int kBufs = 2;                                   // staging units per epilogue warp
constexpr int kUnit = 32 * 128;                            // 32 rows x 64 channels bf16
// Two epilogue warp sets (4 warps each, one per TMEM lane quadrant): set e drains the tiles whose accumulator lives in
// TMEM stage e, i.e. every other tile of the CTA.  With one set the kernel was bound by the epilogue's latency chain
// (TMEM load -> pack -> staging -> store -> statistics, ~2 us per tile against 0.6 us of MMAs).
// MMAs that accumulate into the SAME tensor-memory accumulator are a dependent chain, and at N = 64 an instruction is 32
// cycles of tensor work behind ~90 cycles of accumulate latency: one chain of 36 MMAs per tile kept the tensor pipe 34 %
// busy (ncu).  So a tile's taps are dealt to kChains independent accumulators, issued round-robin, and added in the
// epilogue (fp32).
constexpr int kChains = 1;
// Accumulator stages in tensor memory (64 columns x kChains each): tile l of a CTA uses stage l % kAccStages.  The round
// trip MMA -> commit -> epilogue wake-up -> TMEM read -> release takes several microseconds whatever the tile does, so the
// number of tiles in flight, not any unit's throughput, bounded the kernel with two stages.
constexpr int kAccStages = 8;  // upper bound (512 TMEM columns / (64 kChains)); the launch uses acc_stages of them
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kSmem = kFilterBytes + kStages * kPatchStride + kEpiWarps * kBufs * kUnit + kEpiWarps * 64 * 2 * 4 + 256 + 1024;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(kThreads, 1)
halo3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmY, const HaloKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_filter = smem;
  uint8_t* s_patch = s_filter + kFilterBytes;
  uint8_t* staging = s_patch + kStages * kPatchStride;
  float* s_part = reinterpret_cast<float*>(staging + kEpiWarps * kBufs * kUnit);  // [8 warps][64 channels][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_part + kEpiWarps * 64 * 2);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + kAccStages;
  uint64_t* filter_bar = tempty_bar + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(filter_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(filter_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);  // 2 stages x kChains accumulators of 64 columns (384 used)
  pdl_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_image = p.tiles_h * p.tiles_w;
  const int total_tiles = p.N * tiles_per_image;
  auto tile_coords = [&](int t, int* n, int* p0, int* q0) {
    if (p.rev) t = total_tiles - 1 - t;
    *n = t / tiles_per_image;
    const int rem = t - *n * tiles_per_image;
    const int th = rem / p.tiles_w;
    *p0 = th * kTileRows;
    *q0 = (rem - th * p.tiles_w) * kTileCols;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(filter_bar, static_cast<uint32_t>(p.num_taps * 64 * 128));
      for (int t = 0; t < p.num_taps; ++t) tma_load_2d(&tmW, filter_bar, s_filter + t * 8192, t * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) {
          atomicExch(p.error_flag, 41);
          break;
        }
        int n, p0, q0;
        tile_coords(tile, &n, &p0, &q0);
        if (p.debug & 4) {
          mbar_arrive(&full_bar[stage]);
        } else {
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.patch_bytes));
          tma_load_4d(&tmX, &full_bar[stage], s_patch + stage * kPatchStride, 0, q0 + p.org_w, p0 + p.org_h, n);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      pdl_done();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, 128, 64, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      bool ok = mbar_wait(filter_bar, 0);
      if (!ok) atomicExch(p.error_flag, 42);
      const uint32_t filter_addr = smem_u32(s_filter);
      for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++local) {
        const int acc = local % p.acc_stages;
        const uint32_t acc_phase = static_cast<uint32_t>(local / p.acc_stages) & 1u;
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1u) || !mbar_wait(&full_bar[stage], phase)) {
          atomicExch(p.error_flag, 43);
          break;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kChains * 64);
        const uint32_t patch_addr = smem_u32(s_patch + stage * kPatchStride);
        // chain c accumulates taps c, c + kChains, ...; consecutive instructions belong to different chains
        for (int t0 = 0; t0 < p.num_taps; t0 += kChains) {
          uint64_t da[kChains], db[kChains];
#pragma unroll
          for (int c = 0; c < kChains; ++c) {
            const int t = min(t0 + c, p.num_taps - 1);
            // tap (r, s): rows 10 r + s ... of the halo; output row g of the tile = core group g, 10 halo rows further on
            const uint32_t row0 = static_cast<uint32_t>(p.tap_h[t]) * kHaloCols + p.tap_w[t];
            da[c] = make_smem_desc_sw128_bo(patch_addr + row0 * 128, 16, kHaloCols * 128,
                                            p.base_offset_mode ? (row0 & 7u) : 0u);
            db[c] = make_smem_desc_sw128(filter_addr + t * 8192, 16, 1024);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < kChains; ++c)
              if (t0 + c < p.num_taps && (!(p.debug & 2) || (t0 | k | c) == 0))
                umma_bf16(d_tmem + c * 64, da[c] + static_cast<uint64_t>(k * 2), db[c] + static_cast<uint64_t>(k * 2), idesc,
                          (t0 | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: warp (set e, quadrant q) owns output rows
    // 4q .. 4q+3 of every other tile of the CTA
    const int q = warp & 3;
    const int e = (warp - 4) >> 2;
    const uint32_t stg_base = smem_u32(staging + (warp - 4) * (kBufs * kUnit));
    int buf = 0;
    int local = 0;  // index of the tile among this CTA's tiles
    float r_s0 = 0.f, r_s1 = 0.f, r_q0 = 0.f, r_q1 = 0.f;  // BatchNorm statistics of columns (2 lane, 2 lane + 1)
    const bool do_stats = p.stat_acc != nullptr;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      if ((local & 1) != e) continue;
      const int acc = local % p.acc_stages;
      const uint32_t acc_phase = static_cast<uint32_t>(local / p.acc_stages) & 1u;
      int n, p0, q0;
      tile_coords(tile, &n, &p0, &q0);
      if (!mbar_wait(&tfull_bar[acc], acc_phase)) {
        if (lane == 0) atomicExch(p.error_flag, 44);
        break;
      }
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kChains * 64);
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(t_row, v0);
      tmem_ld_32x32(t_row + 32, v1);
      tc_wait_ld();
      const int used_chains = p.num_taps < kChains ? p.num_taps : kChains;
#pragma unroll 1
      for (int c = 1; c < used_chains; ++c) {  // the other chains' partial sums
        uint32_t u0[32], u1[32];
        tmem_ld_32x32(t_row + c * 64, u0);
        tmem_ld_32x32(t_row + c * 64 + 32, u1);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v0[j] = __float_as_uint(__uint_as_float(v0[j]) + __uint_as_float(u0[j]));
          v1[j] = __float_as_uint(__uint_as_float(v1[j]) + __uint_as_float(u1[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);  // the accumulator is in registers: hand the stage back
      const bool valid = p0 + 4 * q < p.H && !(p.debug & 1);  // rows past the image (H % 4 == 0: the same for the whole warp)
      if (valid) {
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = pack_bf16x2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
          pk[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        }
        const uint32_t stg = stg_base + buf * kUnit;
        if (++buf == kBufs) buf = 0;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBufs - 1) : "memory");
        __syncwarp();
        const uint32_t rbase = stg + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = rbase + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                       "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmY, stg, 0, q0, p0 + 4 * q, n);  // 64 channels x 8 columns x 4 rows
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (do_stats) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            uint32_t w;
            const uint32_t addr = stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2);
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr));
            const float a = bf16lo(w), b = bf16hi(w);
            r_s0 += a;
            r_s1 += b;
            r_q0 = fmaf(a, a, r_q0);
            r_q1 = fmaf(b, b, r_q1);
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (do_stats) {
      float* dst = s_part + ((warp - 4) * 64 + 2 * lane) * 2;
      dst[0] = r_s0;
      dst[1] = r_q0;
      dst[2] = r_s1;
      dst[3] = r_q1;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      const int et = threadIdx.x - 128;
      if (et < 64) {
        float sm = 0.f, sq = 0.f;
#pragma unroll
        for (int qq = 0; qq < kEpiWarps; ++qq) {
          sm += s_part[(qq * 64 + et) * 2];
          sq += s_part[(qq * 64 + et) * 2 + 1];
        }
        fx_add(p.stat_acc + 2 * kFxWords * et, sm);
        fx_add(p.stat_acc + 2 * kFxWords * et + kFxWords, sq);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

cudaError_t halo3x3_launch(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY,
                           const HaloKernelParams& p, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(halo3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  launch_kernel(halo3x3_kernel, dim3(grid), dim3(kThreads), kSmem, stream, tmX, tmW, tmY, p);
  return cudaGetLastError();
}

}  // namespace r3m
