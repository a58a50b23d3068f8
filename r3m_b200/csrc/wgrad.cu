// Filter-gradient (wgrad) implicit GEMM on tcgen05 (sm_100a).
//
//   dW[k, tap, c] = sum_m dY[m, k] * X[pixel(m) + tap, c]            m = flattened (n, p, q)
//
// The reduction runs over pixels, which is the OUTER dimension of both NHWC operands, so both are fed to the
// tensor core as MN-major tiles (channels contiguous, 128B-swizzled rows of 64 channels):
//   A' = dY tile   [64 pixels][128 out-channels]  -> two tiled-TMA boxes of 64 channels
//   B' = X  tiles  [64 pixels][64 in-channels] for up to 8 (tap, channel-block) items -> im2col TMA each
// One CTA owns a 128 x (64*group) slab of dW for a contiguous range of pixel blocks (split-K); the fp32 partial
// sits in TMEM for the whole CTA lifetime (up to all 512 columns) and is added to global memory once, with
// vectorised red.global.add.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue.
#include <algorithm>
#include <cstdlib>

#include "conv_igemm.cuh"
#include "launch.h"
#include "ptx.cuh"

namespace r3m {

namespace {

constexpr int kTmemCols = 512;
// One "atom" = [pix_block pixels][64 channels] bf16, one TMA op.  The TMA unit retires roughly one op per ~300 cycles
// per SM regardless of its size (measured), so the reduction block per stage is as tall as shared memory allows.

// Epilogue of one epilogue warp: its 32 rows of the CTA's 128 x (64 * g_count) slab, TMEM -> global.  Item t of the
// group sits in TMEM columns [64 t, 64 t + 64).
__device__ __forceinline__ void wgrad_store_slab(const WgradKernelParams& p, uint32_t tmem_base, int item0, int g_count,
                                                 int k0, int ew, int row, bool row_ok, int split) {
  int tap = item0 / p.cblocks;
  int cb = item0 - tap * p.cblocks;
  for (int g = 0; g < g_count; ++g) {
    // split-K partials go to this split's private copy of dW (summed in split order by wgrad_reduce_kernel: no
    // floating-point atomics, bit-identical from run to run); a single split adds straight into dW
    float* base = p.scratch ? p.scratch + static_cast<size_t>(split) * p.dw_elems : p.dW;
    float* dst_row = base + static_cast<size_t>(k0 + row) * p.ldw + static_cast<size_t>(tap) * p.Cin + cb * 64;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + g * 64 + half * 32, v);
      tc_wait_ld();
      if (row_ok) {
        if (p.scratch) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst_row + half * 32 + j * 4) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float* d = dst_row + half * 32 + j * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(__uint_as_float(v[4 * j])),
                         "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])),
                         "f"(__uint_as_float(v[4 * j + 3]))
                         : "memory");
          }
        }
      }
    }
    if (++cb == p.cblocks) {
      cb = 0;
      ++tap;
    }
  }
}

__global__ void __launch_bounds__(256, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
             const WgradKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kPixBlock = p.pix_block;
  const int kAtomBytes = kPixBlock * 128;
  const int stage_bytes = (2 + p.group) * kAtomBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int item0 = blockIdx.y * p.group;
  const int g_count = min(p.group, p.num_items - item0);
  const int k0 = blockIdx.z * 128;
  const int a_atoms = (p.Cout - k0 >= 128) ? 2 : 1;
  const int blk_begin = blockIdx.x * p.mblocks_per_split;
  const int blk_end = min(p.mblocks_total, blk_begin + p.mblocks_per_split);
  const int nblk = blk_end - blk_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_sync();  // everything above is CTA-local; global memory is first touched below
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = static_cast<uint32_t>((a_atoms + g_count) * kAtomBytes);
        for (int blk = blk_begin; blk < blk_end; ++blk) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) {
            atomicExch(p.error_flag, 11);
            break;
          }
          const int m0 = blk * kPixBlock;
          const int n_img = m0 / p.PQ;
          const int rem = m0 - n_img * p.PQ;
          const int pp = rem / p.Q;
          const int qq = rem - pp * p.Q;
          const int cw = p.base_w + qq * p.stride;
          const int ch = p.base_h + pp * p.stride;
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + 2 * kAtomBytes;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          for (int a = 0; a < a_atoms; ++a) tma_load_2d(&tmDy, &full_bar[stage], sa + a * kAtomBytes, k0 + a * 64, m0);
          int tap = item0 / p.cblocks;
          int cb = item0 - tap * p.cblocks;
          for (int g = 0; g < g_count; ++g) {
            tma_load_im2col_4d(&tmX, &full_bar[stage], sb + g * kAtomBytes, cb * 64, cw, ch, n_img, p.tap_w[tap],
                               p.tap_h[tap]);
            if (++cb == p.cblocks) {
              cb = 0;
              ++tap;
            }
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        pdl_done();  // all loads of this CTA are issued
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // MN-major SW128 canonical layout: 64-channel atoms LBO apart along M/N, 8-pixel groups SBO apart along K.
        // With a single 64-channel atom of dY (Cout == 64) the second M atom aliases the first (LBO = 0); its
        // 64 duplicate accumulator rows are never read back.
        const uint32_t lbo_a = (a_atoms == 2) ? kAtomBytes : 0;
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int blk = 0; blk < nblk; ++blk) {
          if (!mbar_wait(&full_bar[stage], phase)) {
            atomicExch(p.error_flag, 12);
            ok = false;
            break;
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + 2 * kAtomBytes;
          for (int ks = 0; ks < kPixBlock / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 2048, lbo_a, 1024);
            // the items of a group are consecutive 64-wide N atoms (LBO apart), so up to four of them go into one
            // N = 256 instruction: A is then read from shared memory once per four items
            for (int g = 0; g < g_count; g += 4) {
              const int n = min(4, g_count - g) * 64;
              const uint32_t idesc = make_idesc(/*bf16*/ 1, 128, n, /*a MN-major*/ 1, /*b MN-major*/ 1);
              const uint64_t db = make_smem_desc_sw128(b_addr + g * kAtomBytes + ks * 2048, kAtomBytes, 1024);
              umma_bf16(tmem_base + g * 64, da, db, idesc, (blk | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (ok) umma_commit(done_bar);
      }
    } else if (warp >= 4) {
      const int ew = warp - 4;
      const int row = ew * 32 + lane;  // dW row inside the 128-row slab
      const bool row_ok = (k0 + row < p.Cout) && (a_atoms == 2 || row < 64);
      if (!mbar_wait(done_bar, 0)) {
        if (lane == 0) atomicExch(p.error_flag, 13);
      } else {
        tc_fence_after();
        wgrad_store_slab(p, tmem_base, item0, g_count, k0, ew, row, row_ok, blockIdx.x);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): the two CTAs of a cluster own two adjacent 128-row k-slabs of dW over the same pixel
// range and the same group of (tap, 64-channel) items.  ONE tcgen05.mma of M = 256 serves both: each CTA stages its own
// dY tile (A) and only HALF of the group's activation items (B) — the kernel is bound by the bytes an SM can pull into
// shared memory (~40 B/clk: ncu, profiles/r2_ncu_per_kernel.md), and a pair needs (2 + g/2) atoms per CTA and stage for
// the FLOPs the single-CTA kernel pays (2 + g) for.  Item t of the group lands in TMEM columns [64 t, 64 t + 64) of BOTH
// CTAs: instruction i covers items 4i .. 4i+3, the first half from the even CTA's shared memory, the second half from
// the odd CTA's.  Only the even CTA issues MMAs; TMA transaction bytes of both CTAs land on its full barriers, and its
// commits arrive on the empty / done barriers of both.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const WgradKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kPixBlock = p.pix_block;
  const int kAtomBytes = kPixBlock * 128;
  const int b_atoms = p.group / 2;  // activation items this CTA stages per pipeline stage
  const int stage_bytes = (2 + b_atoms) * kAtomBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.num_stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* done_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0: leader (even CTA of the pair)

  const int item0 = blockIdx.y * p.group;
  const int g_count = min(p.group, p.num_items - item0);  // even (host)
  const int k0 = blockIdx.x * 128;  // grid: (k-tiles [the pair axis], item groups, splits)
  const int blk_begin = blockIdx.z * p.mblocks_per_split;
  const int blk_end = min(p.mblocks_total, blk_begin + p.mblocks_per_split);
  const int nblk = blk_end - blk_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_pair<kTmemCols>(tmem_slot);
  pdl_sync();  // everything above is CTA-local; global memory is first touched below
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nblk > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = static_cast<uint32_t>(2 * (2 + g_count / 2) * kAtomBytes);  // both CTAs' loads
        for (int blk = blk_begin; blk < blk_end; ++blk) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) {
            atomicExch(p.error_flag, 14);
            break;
          }
          const int m0 = blk * kPixBlock;
          const int n_img = m0 / p.PQ;
          const int rem = m0 - n_img * p.PQ;
          const int pp = rem / p.Q;
          const int qq = rem - pp * p.Q;
          const int cw = p.base_w + qq * p.stride;
          const int ch = p.base_h + pp * p.stride;
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + 2 * kAtomBytes;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_bytes);
          for (int a = 0; a < 2; ++a) tma_load_2d_pair(&tmDy, &full_bar[stage], sa + a * kAtomBytes, k0 + a * 64, m0);
          for (int i = 0; 4 * i < g_count; ++i) {
            const int half = min(4, g_count - 4 * i) / 2;
            for (int j = 0; j < half; ++j) {
              const int item = item0 + 4 * i + static_cast<int>(rank) * half + j;
              const int tap = item / p.cblocks;
              const int cb = item - tap * p.cblocks;
              tma_load_im2col_4d_pair(&tmX, &full_bar[stage], sb + (2 * i + j) * kAtomBytes, cb * 64, cw, ch, n_img,
                                      p.tap_w[tap], p.tap_h[tap]);
            }
          }
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        pdl_done();  // all loads of this CTA are issued
      }
    } else if (warp == 1) {
      if (lane == 0 && rank == 0) {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int blk = 0; blk < nblk; ++blk) {
          if (!mbar_wait(&full_bar[stage], phase)) {
            atomicExch(p.error_flag, 15);
            ok = false;
            break;
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + 2 * kAtomBytes;
          for (int ks = 0; ks < kPixBlock / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 2048, kAtomBytes, 1024);
            for (int i = 0; 4 * i < g_count; ++i) {
              const int n = min(4, g_count - 4 * i) * 64;  // N of the pair: half of it from each CTA's shared memory
              const uint32_t idesc = make_idesc(/*bf16*/ 1, 256, n, /*a MN-major*/ 1, /*b MN-major*/ 1);
              const uint64_t db = make_smem_desc_sw128(b_addr + 2 * i * kAtomBytes + ks * 2048, kAtomBytes, 1024);
              umma_bf16_pair(tmem_base + i * 256, da, db, idesc, (blk | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (ok) umma_commit_pair(done_bar, 3);
      }
    } else if (warp >= 4) {
      const int ew = warp - 4;
      const int row = ew * 32 + lane;  // dW row inside this CTA's 128-row slab
      if (!mbar_wait(done_bar, 0)) {
        if (lane == 0) atomicExch(p.error_flag, 16);
      } else {
        tc_fence_after();
        wgrad_store_slab(p, tmem_base, item0, g_count, k0, ew, row, k0 + row < p.Cout, blockIdx.z);
      }
    }
  }

  // neither CTA may retire (its shared memory is an MMA operand of the pair) before both are done
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

// dW[i] += sum over splits of scratch[s][i], in a FIXED association (deterministic): the splits are cut into G
// contiguous groups, thread (x, g) adds group g's copies of position x in split order, and the G group sums are added
// in group order.  One thread walking all (up to 148) copies of its position was a chain of dependent L2 round trips:
// 8-16 us per launch for 10-28 MB (ncu), 53 launches per ResNet-50 step.
template <int G>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float4* __restrict__ scratch, float4* __restrict__ dW,
                                                           size_t n4, int splits) {
  pdl_sync();
  constexpr int kPos = 256 / G;
  __shared__ float4 part[G][kPos];
  const int x = threadIdx.x % kPos, g = threadIdx.x / kPos;
  const int per = (splits + G - 1) / G;
  const int s0 = g * per, s1 = min(splits, s0 + per);
  for (size_t base = (size_t)blockIdx.x * kPos; base < n4; base += (size_t)gridDim.x * kPos) {
    const size_t i = base + x;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      int s = s0;
      for (; s + 4 <= s1; s += 4) {  // four copies in flight
        const float4 v0 = __ldcg(scratch + (size_t)s * n4 + i);
        const float4 v1 = __ldcg(scratch + (size_t)(s + 1) * n4 + i);
        const float4 v2 = __ldcg(scratch + (size_t)(s + 2) * n4 + i);
        const float4 v3 = __ldcg(scratch + (size_t)(s + 3) * n4 + i);
        a.x = ((a.x + v0.x) + v1.x) + v2.x + v3.x;
        a.y = ((a.y + v0.y) + v1.y) + v2.y + v3.y;
        a.z = ((a.z + v0.z) + v1.z) + v2.z + v3.z;
        a.w = ((a.w + v0.w) + v1.w) + v2.w + v3.w;
      }
      for (; s < s1; ++s) {
        const float4 v = __ldcg(scratch + (size_t)s * n4 + i);
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
      }
    }
    if (G > 1) {
      part[g][x] = a;
      __syncthreads();
      if (g == 0 && i < n4) {
        float4 t = dW[i];
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const float4 v = part[k][x];
          t.x += v.x;
          t.y += v.y;
          t.z += v.z;
          t.w += v.w;
        }
        dW[i] = t;
      }
      __syncthreads();
    } else if (i < n4) {
      float4 t = dW[i];
      t.x += a.x;
      t.y += a.y;
      t.z += a.z;
      t.w += a.w;
      dW[i] = t;
    }
  }
}

}  // namespace

int wgrad_smem_bytes(int group, int num_stages, int pix_block) {
  return num_stages * (2 + group) * pix_block * 128 + 256 + 1024;
}

cudaError_t wgrad_launch(const CUtensorMap& tmDy, const CUtensorMap& tmX, const WgradKernelParams& p, int splits,
                         int groups, int ktiles, cudaStream_t stream) {
  static int configured_bytes = 0, configured_pair_bytes = 0;
  const int bytes = wgrad_smem_bytes(p.pair ? p.group / 2 : p.group, p.num_stages, p.pix_block);
  if (p.pair && bytes > configured_pair_bytes) {
    cudaError_t ce = cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (ce != cudaSuccess) return ce;
    configured_pair_bytes = bytes;
  }
  if (!p.pair && bytes > configured_bytes) {
    cudaError_t ce = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (ce != cudaSuccess) return ce;
    configured_bytes = bytes;
  }
  cudaError_t e;
  if (p.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ktiles, groups, splits);  // the pair axis must be x (cluster 2 x 1 x 1)
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = (pdl_enabled() && g_pdl_suppress == 0) ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, wgrad_pair_kernel, tmDy, tmX, p);
    if (e == cudaSuccess) e = cudaGetLastError();
  } else {
    launch_kernel(wgrad_kernel, dim3(splits, groups, ktiles), 256, bytes, stream, tmDy, tmX, p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess || p.scratch == nullptr) return e;
  const size_t n4 = p.dw_elems / 4;
  const float4* src = reinterpret_cast<const float4*>(p.scratch);
  float4* dst = reinterpret_cast<float4*>(p.dW);
  // G split groups per position: enough threads to cover the copies of small filters (few positions, many splits)
  auto blocks_for = [&](int pos_per_block) { return (int)std::min<size_t>((n4 + pos_per_block - 1) / pos_per_block, 148 * 8); };
  static const int ab = std::getenv("R3M_AB") ? atoi(std::getenv("R3M_AB")) : 0;  // A/B aid: bit 0 = one group
  if (ab & 1)
    launch_kernel(wgrad_reduce_kernel<1>, blocks_for(256), 256, 0, stream, src, dst, n4, splits);
  else if (splits >= 16)
    launch_kernel(wgrad_reduce_kernel<8>, blocks_for(32), 256, 0, stream, src, dst, n4, splits);
  else if (splits >= 4)
    launch_kernel(wgrad_reduce_kernel<4>, blocks_for(64), 256, 0, stream, src, dst, n4, splits);
  else
    launch_kernel(wgrad_reduce_kernel<1>, blocks_for(256), 256, 0, stream, src, dst, n4, splits);
  return cudaGetLastError();
}

}  // namespace r3m
