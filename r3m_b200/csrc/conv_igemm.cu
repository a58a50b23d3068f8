// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 in / bf16 out, fp32 accumulate.
//
//   Y[m, k] = sum_{tap, c} X[pixel(m) + tap, c] * Wp[k, tap, c]          m = flattened (n, p, q)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    TMA producer : A tile = 128 output pixels x 64 channels of ONE filter tap, fetched by a single
//                            im2col-mode TMA (the hardware walks W->H->N inside the padded bounding box and
//                            zero-fills the halo); B tile = BN x 64 slice of the packed filter, tiled TMA.
//                            Both land in 128B-swizzled K-major shared memory.
//   warp 1    MMA issuer   : one thread issues 4 x tcgen05.mma (M=128, N=BN, K=16) per stage; accumulators
//                            live in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps
//                            the main loop of tile i+1.
//   warp 2    TMEM allocator.
//   warps 4-11 epilogue    : two warps per TMEM lane quadrant, alternating 64-column units:
//                            tcgen05.ld (next unit in flight) -> bf16 -> 128B-swizzled smem staging ->
//                            (a) per-channel sum / sum-of-squares for train-mode BatchNorm, accumulated per CTA in
//                            smem and flushed once with atomics, (b) dense outputs leave through TMA bulk tensor
//                            stores (cp.reduce...add when accumulating into an existing gradient), strided scatter
//                            outputs (stride-2 dgrad parity classes) through 64-byte row segments.
//
// The same kernel serves: every forward conv of ResNet-18/34/50 (ref: torchvision resnet.py conv1/conv2/conv3/
// downsample and the 7x7 stem re-expressed as a 4x1-tap conv over a space-to-depth view), and every dgrad
// (stride-1: rotated/transposed filter; stride-2: one launch per output-parity class with scatter stores).
#include "conv_igemm.cuh"

#include "launch.h"
#include "ptx.cuh"

namespace r3m {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kUnitCols = 64;  // one epilogue unit: 32 rows x 64 columns bf16 = 128-byte rows (TMA stores are paced per row)
constexpr int kUnitBytes = 32 * kUnitCols * 2;
template <int BN>
struct StatC {  // per-CTA statistics / affine table capacity (channels): narrow tiles serve Cout <= 1024
  static constexpr int value = (BN == 256) ? 2048 : 1024;
};

// STAGES: depth of the TMA->MMA operand ring.  BUFS: staging buffers per epilogue warp (bulk stores in flight).
// Long-K tiles (3x3 convs) want the deep ring, short-K tiles (1x1 convs into wide outputs) are epilogue bound and want
// several stores in flight; both fit the 227 KB budget only one at a time for BN = 256.
// EPI: epilogue warps (4 or 8 = one or two per TMEM lane quadrant).
// KPS: 64-wide K blocks per pipeline stage.  The TMA-producer and MMA-issuer roles are single threads whose per-stage
// bookkeeping (mbarrier try_wait ~90 cycles, expect_tx, descriptor math, op issue) costs ~450-600 cycles (measured:
// per-stage time is flat in the number of MMAs and TMA ops), so a stage must carry at least that much tensor work:
// 4 MMAs of N=256 (512 cycles) do, 4 MMAs of N=64 (128 cycles) do not -> narrow tiles put several K blocks in a stage.
// PAIR: the CTA is one half of a cta_group::2 pair (two M tiles against one filter tile): it stages only HALF of the
// filter tile's rows (the N columns of the pair's B operand are split between the two shared memories).
template <int BN, int STAGES, int BUFS, int EPI, int KPS, bool PAIR = false>
struct Cfg {
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBlockK * 2;
  static constexpr int kKbBytes = kABytes + kBBytes;   // one K block: A sub-tile then B sub-tile
  static constexpr int kStageBytes = KPS * kKbBytes;
  static constexpr int kStages = STAGES;
  static constexpr int kStagingBytes = EPI * BUFS * kUnitBytes;
  static constexpr int kThreads = 128 + EPI * 32;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + 2 * StatC<BN>::value * 4 + 256 /*barriers*/ + 1024 /*align slack*/;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// MODE (compile time, because the three epilogues' register needs do not fit together in the 170-register budget of
// the 8-epilogue-warp configurations: compiled into one kernel they spilled and the stride-2 dgrads lost 30 %):
//   kModeDense   dense output (TMA bulk stores / reduce-add), optional folded-BN epilogue
//   kModeStats   dense output + BatchNorm statistics (training forward convs)
//   kModeScatter strided scatter output with optional read-modify-write (stride-2 dgrad parity classes)
enum { kModeDense = 0, kModeStats = 1, kModeScatter = 2 };
// PREC: 0 = bf16 operands / bf16 output (training and the fast inference tier); 1 = fp32 storage with kind::tf32 tensor
// core arithmetic and fp32 (tf32-rounded) output — the parity tier of the inference path (dense mode only).  The
// 128-byte swizzle row then holds 32 channels instead of 64 and an MMA consumes K = 8 of them; tile BYTES, the operand
// ring and the descriptor stepping are unchanged.
// PAIR (bf16, BN = 256, dense / statistics modes): clusters of two CTAs; the pair computes two adjacent M tiles against
// one filter tile with ONE tcgen05.mma.cta_group::2 of M = 256 per K step.  Each CTA loads its own activation tile and
// half of the filter tile (32 KB of operands per K block instead of 48 KB: the kernel is bound by the bytes an SM can
// pull into shared memory, profiles/r2_ncu_per_kernel.md), TMA bytes of both CTAs complete on the even CTA's full
// barriers, only the even CTA issues MMAs, its commits arrive on the barriers of both, and the odd CTA's epilogue warps
// release the accumulator stage on the even CTA's barrier.  An odd number of M tiles: the last pair's second CTA repeats
// the first one's tile and discards the result.
template <int BN, int STAGES, int BUFS, int EPI, int KPS, int MODE, int PREC = 0, bool PAIR = false>
__global__ void __launch_bounds__(128 + EPI * 32, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const ConvKernelParams p) {
  using C = Cfg<BN, STAGES, BUFS, EPI, KPS, PAIR>;
  static_assert(!PAIR || (PREC == 0 && MODE != kModeScatter), "pairs: bf16, dense outputs");
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + C::kStages * C::kStageBytes;
  float* s_sum = reinterpret_cast<float*>(staging + C::kStagingBytes);
  float* s_sq = s_sum + StatC<BN>::value;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_sq + StatC<BN>::value);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr bool STATS = (MODE == kModeStats);
  constexpr bool do_stats = STATS;
  static_assert(PREC == 0 || MODE == kModeDense, "the tf32 tier serves dense (inference) outputs only");
  constexpr int kElemsK = PREC ? 32 : 64;  // operand elements per 128-byte K block
  // Chunked accumulation (tf32 tier, 64-wide tiles: the split-tf32 GEMMs of the sentence encoder and the language head,
  // which must stay in the fp32 parity band): the tensor core adds each MMA's partial sum into the fp32 accumulator with
  // truncation, a bias that grows with the number of MMAs of a reduction (2e-5 relative after 384 of them: ten times the
  // round-off of an fp32 FMA chain, i.e. ten times as many ReLU pre-activations on the wrong side of zero).  So a
  // reduction is cut into CH chunks that accumulate from zero in separate TMEM column blocks and are added in the
  // epilogue with rounded fp32 adds.
  constexpr int CH = (PREC == 1 && BN == 64) ? 4 : 1;
  constexpr int kTmemCols = 2 * BN * CH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.out_mode == 0) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * EPI : EPI);  // pairs: the epilogue warps of BOTH CTAs release the leader's stage
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR)
      tmem_alloc_pair<kTmemCols>(tmem_slot);
    else
      tmem_alloc<kTmemCols>(tmem_slot);
  }
  pdl_sync();  // everything above is CTA-local; global memory is first touched below
  const bool do_affine = !STATS && (p.ep_scale != nullptr);
  if (do_affine) {
    // the statistics arrays double as the per-channel scale / shift table of the fused inference epilogue
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      s_sum[i] = p.ep_scale[i];
      s_sq[i] = p.ep_shift[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tiles of this CTA: tile0, tile0 + tile_step, ...; a pair walks pair-tiles (two M tiles x one N tile)
  const int total_tiles = (PAIR ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;
  const int tile0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_kb = p.num_taps * p.cblocks;
  const int num_steps = (num_kb + KPS - 1) / KPS;       // pipeline stages per tile
  const int nch = CH < num_steps ? CH : num_steps;      // accumulation chunks per tile (every chunk gets >= 1 stage)
  // M tile of this CTA inside tile `t` (+ whether it is the discarded duplicate that completes an odd last pair)
  auto m_tile_of = [&](int t_m_idx, bool* dup) {
    int m_idx = t_m_idx;
    *dup = false;
    if constexpr (PAIR) {
      m_idx = 2 * t_m_idx + static_cast<int>(rank);
      if (m_idx >= p.num_m_tiles) {
        m_idx = p.num_m_tiles - 1;
        *dup = true;
      }
    }
    return p.rev_m ? p.num_m_tiles - 1 - m_idx : m_idx;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = tile0; tile < total_tiles && ok; tile += tile_step) {
        const int m_idx = tile / p.num_n_tiles;
        const int n_tile = tile - m_idx * p.num_n_tiles;
        bool dup;
        const int m_tile = m_tile_of(m_idx, &dup);
        const int m0 = m_tile * kBlockM;
        const int n_img = m0 / p.PQ;
        const int rem = m0 - n_img * p.PQ;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        const int cw = p.base_w + qq * p.stride;
        const int ch = p.base_h + pp * p.stride;
        int tap = 0, cb = 0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += KPS) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) {
            atomicExch(p.error_flag, 1);
            ok = false;
            break;
          }
          const int nk = min(KPS, num_kb - kb0);
          uint8_t* st = smem + stage * C::kStageBytes;
          if constexpr (PAIR) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(2 * nk * C::kKbBytes));  // both CTAs' loads
          } else {
            mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(nk * C::kKbBytes));
          }
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              uint8_t* sa = st + j * C::kKbBytes;
              if constexpr (PAIR) {
                tma_load_im2col_4d_pair(&tmA, &full_bar[stage], sa, cb * kElemsK, cw, ch, n_img, p.tap_w[tap],
                                        p.tap_h[tap]);
                tma_load_2d_pair(&tmB, &full_bar[stage], sa + kABytes, (kb0 + j) * kElemsK,
                                 n_tile * BN + static_cast<int>(rank) * (BN / 2));
              } else {
              tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * kElemsK, cw, ch, n_img, p.tap_w[tap], p.tap_h[tap]);
              tma_load_2d(&tmB, &full_bar[stage], sa + kABytes, (kb0 + j) * kElemsK, n_tile * BN);
              }
              if (++cb == p.cblocks) {
                cb = 0;
                ++tap;
              }
            }
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      pdl_done();  // all loads of this CTA are issued
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(PREC ? /*tf32*/ 2 : /*bf16*/ 1, PAIR ? 2 * kBlockM : kBlockM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ok = true;
      for (int tile = tile0; tile < total_tiles && ok; tile += tile_step) {
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1u)) {
          atomicExch(p.error_flag, 2);
          break;
        }
        tc_fence_after();
        const uint32_t d_tile = tmem_base + static_cast<uint32_t>(acc * BN * CH);
        int cur_chunk = -1;
        for (int kb0 = 0; kb0 < num_kb; kb0 += KPS) {
          const int chunk = CH > 1 ? ((kb0 / KPS) * nch) / num_steps : 0;
          const bool fresh = chunk != cur_chunk;  // first stage of an accumulation chunk: its first MMA overwrites
          cur_chunk = chunk;
          const uint32_t d_tmem = d_tile + static_cast<uint32_t>(chunk * BN);
          if (!mbar_wait(&full_bar[stage], phase)) {
            atomicExch(p.error_flag, 3);
            ok = false;
            break;
          }
          tc_fence_after();
          const int nk = min(KPS, num_kb - kb0);
          const uint32_t st_addr = smem_u32(smem + stage * C::kStageBytes);
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              const uint32_t a_addr = st_addr + j * C::kKbBytes;
              const uint64_t da = make_smem_desc_sw128(a_addr, 16, 1024);
              const uint64_t db = make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // +32 bytes per MMA (K = 16 bf16 or 8 tf32) inside the 128-byte swizzle row (start-address field is >>4)
                if constexpr (PAIR)
                  umma_bf16_pair(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                                 (!fresh || (j | k) != 0) ? 1u : 0u);
                else if constexpr (PREC == 0)
                  umma_bf16(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                            (!fresh || (j | k) != 0) ? 1u : 0u);
                else
                  umma_tf32(d_tmem, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                            (!fresh || (j | k) != 0) ? 1u : 0u);
              }
            }
          }
          if constexpr (PAIR)
            umma_commit_pair(&empty_bar[stage], 3);
          else
            umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (!ok) break;
        if constexpr (PAIR)
          umma_commit_pair(&tfull_bar[acc], 3);
        else
          umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // warp (q, h): TMEM lane quadrant q = warp % 4 (rows 32q..32q+31 of the tile), 32-column units u = h, h+2, ...
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    auto release_acc = [&](uint64_t* bar) {  // the accumulator stage is free again: tell the (pair's) MMA issuer
      if constexpr (PAIR)
        mbar_arrive_leader(bar);
      else
        mbar_arrive(bar);
    };
    const uint32_t stg_base = smem_u32(staging + (warp - 4) * (BUFS * kUnitBytes));
    int buf = 0;
    constexpr bool dense = (MODE != kModeScatter);
    int acc = 0;
    uint32_t acc_phase = 0;
    __nv_bfloat16* __restrict__ outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    constexpr int kUnits = BN / kUnitCols;
    // BatchNorm statistics: a lane owns two columns of each unit its warp drains.  The partial sums stay in registers
    // for as long as the CTA's tiles keep the same column block (always, when gridDim is a multiple of num_n_tiles)
    // and reach the CTA's shared-memory table once; shared-memory fp32 atomics are CAS loops, and four of them per
    // unit were ~20 % of the epilogue's issue slots (ncu source view of the 64 -> 256 1x1 conv).
    constexpr int kUPW = STATS ? (kUnits + EPI / 4 - 1) / (EPI / 4) : 1;  // units per warp per tile
    float racc[kUPW][4];
#pragma unroll
    for (int i = 0; i < kUPW; ++i) racc[i][0] = racc[i][1] = racc[i][2] = racc[i][3] = 0.f;
    int acc_ntile = -1;
    // Deterministic statistics: the host sizes the grid as a multiple of num_n_tiles, so a CTA keeps ONE column block
    // (n_tile = blockIdx.x % num_n_tiles) for all its tiles and the per-lane partials never leave registers before the
    // end.  Then: per-quadrant partials -> shared memory, summed over the four quadrants in fixed order -> added to the
    // column block's fixed-point accumulators (fx_add: exact, order independent) -> the LAST CTA of the column block
    // (ticket) converts the totals to fp32.  No floating-point atomics: the sums are bit-identical from run to run.
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int m_idx = tile / p.num_n_tiles;
      const int n_tile = tile - m_idx * p.num_n_tiles;
      bool dup;
      const int m_tile = m_tile_of(m_idx, &dup);
      // a duplicate tile (odd last pair) is drained like any other but leaves no trace: its rows are placed beyond M
      const int m0 = (dup ? p.num_m_tiles : m_tile) * kBlockM + q * 32;
      const int n0 = n_tile * BN;
      if constexpr (STATS) acc_ntile = n_tile;
      long long row_off[dense ? 1 : 8];
      if constexpr (!dense) {
        // element offsets of the 8 rows this lane stores (row = 4*i + lane/8 of the warp's 32 rows)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + 4 * i + (lane >> 3);
          if (m >= p.M_total) {
            row_off[i] = -1;
          } else {
            const int n_img = m / p.PQ;
            const int rem = m - n_img * p.PQ;
            const int pp = rem / p.Q;
            const int qq = rem - pp * p.Q;
            row_off[i] = ((static_cast<long long>(n_img) * p.oH + (pp * p.o_stride + p.o_h0)) * p.oW +
                          (qq * p.o_stride + p.o_w0)) * p.ldo;
          }
        }
      }
      if (!mbar_wait(&tfull_bar[acc], acc_phase)) {
        if (lane == 0) atomicExch(p.error_flag, 4);
        break;
      }
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN * CH);
      if constexpr (PREC == 1) {
        // fp32 output: units of 32 rows x 32 columns (128-byte rows); folded BatchNorm (+ fp32 residual) (+ ReLU) on the
        // accumulators, results rounded to tf32 (round-to-nearest: the next conv's tensor cores would truncate)
        constexpr int kUnits32 = BN / 32;
        const float* __restrict__ resp = reinterpret_cast<const float*>(p.ep_res);
        uint32_t v[32];
        if (h < kUnits32) tmem_ld_32x32(t_row + h * 32, v);
#pragma unroll 1
        for (int u = h; u < kUnits32; u += EPI / 4) {
          tc_wait_ld();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if constexpr (CH > 1) {
#pragma unroll 1
            for (int c = 1; c < nch; ++c) {  // the other accumulation chunks: rounded fp32 adds
              tmem_ld_32x32(t_row + c * BN + u * 32, v);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
            }
          }
          if (u + EPI / 4 < kUnits32) {
            tmem_ld_32x32(t_row + (u + EPI / 4) * 32, v);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(&tempty_bar[acc]);
          }
          if (do_affine) {
            const float* sc = s_sum + n0 + u * 32;
            const float* sh = s_sq + n0 + u * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
          }
          if (resp != nullptr && m0 + lane < p.M_total) {
            const float4* rrow = reinterpret_cast<const float4*>(resp + static_cast<long long>(m0 + lane) * p.ldo + n0 + u * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 r = __ldg(rrow + c);
              f[4 * c] += r.x;
              f[4 * c + 1] += r.y;
              f[4 * c + 2] += r.z;
              f[4 * c + 3] += r.w;
            }
          }
          if (p.ep_relu == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (p.ep_relu == 2) {  // exact (erf) GELU: the sentence encoder's feed-forward activation
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
          }
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (p.ep_exact)
              pk[j] = __float_as_uint(f[j]);
            else
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(pk[j]) : "f"(f[j]));
          }
          const uint32_t stg_u32 = stg_base + buf * kUnitBytes;
          if (++buf == BUFS) buf = 0;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BUFS - 1) : "memory");
          __syncwarp();
          const uint32_t rbase = stg_u32 + lane * 128;
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
            const int c0 = n0 + u * 32;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(stg_u32), "r"(c0), "r"(m0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (h >= kUnits32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(&tempty_bar[acc]);
        }
      } else if (h >= kUnits) {
        // BN == 64: a single unit per quadrant; the second warp of the pair has nothing to drain
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(&tempty_bar[acc]);
      } else {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(t_row + h * kUnitCols, v0);
        tmem_ld_32x32(t_row + h * kUnitCols + 32, v1);
        int ui = 0;
#pragma unroll 1
        for (int u = h; u < kUnits; u += EPI / 4, ++ui) {
          tc_wait_ld();
          if (do_affine) {
            // folded BatchNorm (+ residual) (+ ReLU) on the fp32 accumulators; thread = output row m0 + lane
            const float* sc = s_sum + n0 + u * kUnitCols;
            const float* sh = s_sq + n0 + u * kUnitCols;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v0[j] = __float_as_uint(fmaf(__uint_as_float(v0[j]), sc[j], sh[j]));
              v1[j] = __float_as_uint(fmaf(__uint_as_float(v1[j]), sc[32 + j], sh[32 + j]));
            }
            if (p.ep_res != nullptr && m0 + lane < p.M_total) {
              const uint4* rrow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.ep_res) +
                                                                 static_cast<long long>(m0 + lane) * p.ldo + n0 +
                                                                 u * kUnitCols);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 r = __ldg(rrow + c);
                uint32_t* dst = (c < 4) ? &v0[8 * c] : &v1[8 * (c - 4)];
                dst[0] = __float_as_uint(__uint_as_float(dst[0]) + bf16lo(r.x));
                dst[1] = __float_as_uint(__uint_as_float(dst[1]) + bf16hi(r.x));
                dst[2] = __float_as_uint(__uint_as_float(dst[2]) + bf16lo(r.y));
                dst[3] = __float_as_uint(__uint_as_float(dst[3]) + bf16hi(r.y));
                dst[4] = __float_as_uint(__uint_as_float(dst[4]) + bf16lo(r.z));
                dst[5] = __float_as_uint(__uint_as_float(dst[5]) + bf16hi(r.z));
                dst[6] = __float_as_uint(__uint_as_float(dst[6]) + bf16lo(r.w));
                dst[7] = __float_as_uint(__uint_as_float(dst[7]) + bf16hi(r.w));
              }
            }
            if (p.ep_relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v0[j] = __float_as_uint(fmaxf(__uint_as_float(v0[j]), 0.f));
                v1[j] = __float_as_uint(fmaxf(__uint_as_float(v1[j]), 0.f));
              }
            }
          }
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pk[j] = pack_bf16x2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
            pk[16 + j] = pack_bf16x2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
          }
          if (u + EPI / 4 < kUnits) {
            tmem_ld_32x32(t_row + (u + EPI / 4) * kUnitCols, v0);  // in flight while this unit is written out
            tmem_ld_32x32(t_row + (u + EPI / 4) * kUnitCols + 32, v1);
          } else {
            // the accumulator stage is fully in registers: hand it back to the MMA warp early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(&tempty_bar[acc]);
          }
          // the staging buffer about to be overwritten must no longer be read by the bulk store issued BUFS units ago
          const uint32_t stg_u32 = stg_base + buf * kUnitBytes;
          if (++buf == BUFS) buf = 0;
          if (dense && lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BUFS - 1) : "memory");
          __syncwarp();
          // thread = row `lane`: eight 16-byte chunks, XOR-swizzled by row (== TMA SWIZZLE_128B) so that the row-wise
          // writes, the column-pair reads of the statistics and the bulk store agree and are bank-conflict free
          const uint32_t rbase = stg_u32 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = rbase + ((j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                         "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
          }
          if (dense) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (dense && lane == 0) {
            const int c0 = n0 + u * kUnitCols;
            if (p.accumulate) {
              asm volatile(
                  "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                      reinterpret_cast<uint64_t>(&tmC)),
                  "r"(stg_u32), "r"(c0), "r"(m0)
                  : "memory");
            } else {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                               reinterpret_cast<uint64_t>(&tmC)),
                           "r"(stg_u32), "r"(c0), "r"(m0)
                           : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (do_stats && !dup) {
            // lane owns columns (2*lane, 2*lane+1) of the unit; rows beyond M_total are exact zeros (TMA zero fill)
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              uint32_t w;
              const uint32_t addr = stg_u32 + r * 128 + ((((lane >> 2)) ^ (r & 7)) << 4) + ((lane & 3) << 2);
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(addr));
              const float a = bf16lo(w), b = bf16hi(w);
              s0 += a;
              s1 += b;
              q0 = fmaf(a, a, q0);
              q1 = fmaf(b, b, q1);
            }
#pragma unroll
            for (int i = 0; i < kUPW; ++i)
              if (i == ui) {
                racc[i][0] += s0;
                racc[i][1] += s1;
                racc[i][2] += q0;
                racc[i][3] += q1;
              }
          }
          if constexpr (!dense) {
            // strided scatter (stride-2 dgrad parity classes): 8 lanes x 16 B = one 128-byte row segment.  When
            // accumulating, the eight read-modify-write loads of the lane are issued together (one global round trip per
            // unit instead of eight dependent ones).
            uint4 prev[8];
            if (p.accumulate) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                prev[i] = make_uint4(0u, 0u, 0u, 0u);
                if (row_off[i] >= 0)
                  prev[i] = *reinterpret_cast<const uint4*>(outp + row_off[i] + n0 + u * kUnitCols + (lane & 7) * 8);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3);
              const int c = lane & 7;
              uint32_t x0, x1, x2, x3;
              const uint32_t addr = stg_u32 + r * 128 + ((c ^ (r & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                           : "r"(addr));
              if (row_off[i] >= 0) {
                uint4* dst = reinterpret_cast<uint4*>(outp + row_off[i] + n0 + u * kUnitCols + c * 8);
                if (p.accumulate) {
                  const uint4 o = prev[i];
                  x0 = pack_bf16x2(bf16lo(x0) + bf16lo(o.x), bf16hi(x0) + bf16hi(o.x));
                  x1 = pack_bf16x2(bf16lo(x1) + bf16lo(o.y), bf16hi(x1) + bf16hi(o.y));
                  x2 = pack_bf16x2(bf16lo(x2) + bf16lo(o.z), bf16hi(x2) + bf16hi(o.z));
                  x3 = pack_bf16x2(bf16lo(x3) + bf16lo(o.w), bf16hi(x3) + bf16hi(o.w));
                }
                *dst = make_uint4(x0, x1, x2, x3);
              }
            }
            __syncwarp();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (dense && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (do_stats) {
      // s_sum doubles as the quadrant table [4][BN][2] (4 * 256 * 2 floats <= the 2 * StatC floats reserved)
      float* s_part = s_sum;
#pragma unroll
      for (int i = 0; i < kUPW; ++i) {
        const int u = h + i * (EPI / 4);
        if (u < kUnits) {
          float* dst = s_part + (q * BN + u * kUnitCols + 2 * lane) * 2;
          dst[0] = racc[i][0];
          dst[1] = racc[i][2];
          dst[2] = racc[i][1];
          dst[3] = racc[i][3];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
      const int et = threadIdx.x - 128;  // epilogue thread index
      const int my_ntile = tile0 % p.num_n_tiles;
      // this CTA's column sums -> fixed-point accumulators (exact integer adds: independent of the CTAs' arrival order)
      // raw mode (engine): the layer's own accumulators, converted by the consuming BatchNorm kernel — no tail at all
      unsigned long long* acc = (p.stat_raw ? reinterpret_cast<unsigned long long*>(p.stat_sum)
                                            : reinterpret_cast<unsigned long long*>(p.stat_scratch)) +
                                static_cast<size_t>(my_ntile) * BN * 2 * kFxWords;
      if (acc_ntile >= 0) {
        for (int c = et; c < BN; c += EPI * 32) {
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            sm += s_part[(qq * BN + c) * 2];
            sq += s_part[(qq * BN + c) * 2 + 1];
          }
          fx_add(acc + 2 * kFxWords * c, sm);
          fx_add(acc + 2 * kFxWords * c + kFxWords, sq);
        }
      }
      if (!p.stat_raw) {
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
      int* s_flag = reinterpret_cast<int*>(tmem_slot) + 2;
      if (et == 0) {
        const int contributors = (PAIR ? 2 : 1) * ((tile_step - my_ntile + p.num_n_tiles - 1) / p.num_n_tiles);
        const int t = atomicAdd(&p.stat_ticket[my_ntile], 1);
        *s_flag = (t == contributors - 1) ? 1 : 0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI * 32) : "memory");
      if (*s_flag) {
        // last CTA of the column block: totals -> fp32, accumulators and ticket back to zero for the next launch
        __threadfence();
        for (int c = et; c < BN; c += EPI * 32) {
          if (p.stat_rows > 0) {
            fx_moments(acc + 2 * kFxWords * c, acc + 2 * kFxWords * c + kFxWords, p.stat_rows,
                       p.stat_sum[my_ntile * BN + c], p.stat_sq[my_ntile * BN + c]);
          } else {
            p.stat_sum[my_ntile * BN + c] = fx_to_float(acc + 2 * kFxWords * c);
            p.stat_sq[my_ntile * BN + c] = fx_to_float(acc + 2 * kFxWords * c + kFxWords);
          }
          fx_clear(acc + 2 * kFxWords * c);
          fx_clear(acc + 2 * kFxWords * c + kFxWords);
        }
        if (et == 0) p.stat_ticket[my_ntile] = 0;
      }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's shared memory is an operand of the pair's MMAs until both are done
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR)
      tmem_dealloc_pair<kTmemCols>(tmem_base);
    else
      tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BN, int STAGES, int BUFS, int EPI, int KPS, int MODE>
cudaError_t launch_pair_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                            const ConvKernelParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN, STAGES, BUFS, EPI, KPS, true>;
  auto kernel = conv_igemm_kernel<BN, STAGES, BUFS, EPI, KPS, MODE, 0, true>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (grid < 2 || (grid & 1)) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, p);
  return e != cudaSuccess ? e : cudaGetLastError();
}

template <int BN, int STAGES, int BUFS, int EPI, int KPS, int MODE, int PREC = 0>
cudaError_t launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       const ConvKernelParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN, STAGES, BUFS, EPI, KPS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES, BUFS, EPI, KPS, MODE, PREC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  launch_kernel(conv_igemm_kernel<BN, STAGES, BUFS, EPI, KPS, MODE, PREC>, grid, C::kThreads, C::kSmemBytes, stream, tmA,
                tmB, tmC, p);
  return cudaGetLastError();
}

template <int BN, int STAGES, int BUFS, int EPI, int KPS>
cudaError_t launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                      const ConvKernelParams& p, int grid, cudaStream_t stream) {
  if (p.tf32) {
    if (p.stat_sum != nullptr || p.out_mode != 0 || p.accumulate) return cudaErrorInvalidValue;
    return launch_cfg<BN, STAGES, BUFS, EPI, KPS, kModeDense, 1>(tmA, tmB, tmC, p, grid, stream);
  }
  if (p.stat_sum != nullptr) {
    if (p.out_mode != 0 || p.ep_scale != nullptr || grid % p.num_n_tiles != 0 ||
        (!p.stat_raw && (p.stat_sq == nullptr || p.stat_scratch == nullptr || p.stat_ticket == nullptr)))
      return cudaErrorInvalidValue;
    return launch_cfg<BN, STAGES, BUFS, EPI, KPS, kModeStats>(tmA, tmB, tmC, p, grid, stream);
  }
  if (p.out_mode != 0) {
    // the scatter epilogue keeps 8 row offsets + 8 prefetched 16-byte words per lane: 4 epilogue warps (254 registers)
    if constexpr (EPI == 4) {
      if (p.ep_scale != nullptr) return cudaErrorInvalidValue;
      return launch_cfg<BN, STAGES, BUFS, EPI, KPS, kModeScatter>(tmA, tmB, tmC, p, grid, stream);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  return launch_cfg<BN, STAGES, BUFS, EPI, KPS, kModeDense>(tmA, tmB, tmC, p, grid, stream);
}

// CTA pairs: the filter half-tile shrinks a stage from 48 to 32 KB, which buys the ring one (epilogue-bound
// configuration) or two (MMA-bound configuration) more stages
template <int BN, int STAGES, int BUFS, int EPI, int KPS>
cudaError_t launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                        const ConvKernelParams& p, int grid, cudaStream_t stream) {
  if (p.tf32 || p.out_mode != 0) return cudaErrorInvalidValue;
  if (p.stat_sum != nullptr) {
    if (p.ep_scale != nullptr || (grid / 2) % p.num_n_tiles != 0 ||
        (!p.stat_raw && (p.stat_sq == nullptr || p.stat_scratch == nullptr || p.stat_ticket == nullptr)))
      return cudaErrorInvalidValue;
    return launch_pair_cfg<BN, STAGES, BUFS, EPI, KPS, kModeStats>(tmA, tmB, tmC, p, grid, stream);
  }
  return launch_pair_cfg<BN, STAGES, BUFS, EPI, KPS, kModeDense>(tmA, tmB, tmC, p, grid, stream);
}

}  // namespace

cudaError_t conv_igemm_launch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                              const ConvKernelParams& p, int grid, cudaStream_t stream) {
  const int num_kb = p.num_taps * p.cblocks;
  const bool scatter = p.out_mode != 0;
  switch (bn) {
    case 64:  // one unit per quadrant
      if (p.Cout > StatC<64>::value) return cudaErrorInvalidValue;
      if (p.pair) {  // experiment (R3M_CONV_PAIR64=1): does an M = 256 instruction cost what an M = 128 one does at N = 64?
        if (num_kb >= 9) return launch_pair<64, 4, 1, 4, 2>(tmA, tmB, tmC, p, grid, stream);
        return launch_pair<64, 6, 4, 4, 1>(tmA, tmB, tmC, p, grid, stream);
      }
      if (num_kb >= 9) return launch_bn<64, 4, 1, 4, 2>(tmA, tmB, tmC, p, grid, stream);  // 3x3: two K blocks per stage
      return launch_bn<64, 6, 4, 4, 1>(tmA, tmB, tmC, p, grid, stream);
    case 128:
      if (p.Cout > StatC<128>::value) return cudaErrorInvalidValue;
      if (p.pair) return launch_pair<128, 4, 1, 4, 2>(tmA, tmB, tmC, p, grid, stream);  // 3x3 (host: >= 18 K blocks)
      if (num_kb >= 18 || scatter) return launch_bn<128, 3, 1, 4, 2>(tmA, tmB, tmC, p, grid, stream);  // 3x3
      return launch_bn<128, 4, 2, 8, 1>(tmA, tmB, tmC, p, grid, stream);
    case 256:
      if (p.pair) {
        if (num_kb >= 12) return launch_pair<256, 6, 1, 4, 1>(tmA, tmB, tmC, p, grid, stream);
        return launch_pair<256, 4, 2, 8, 1>(tmA, tmB, tmC, p, grid, stream);
      }
      if (num_kb >= 12 || scatter) return launch_bn<256, 4, 1, 4, 1>(tmA, tmB, tmC, p, grid, stream);  // MMA bound: deep ring
      return launch_bn<256, 3, 2, 8, 1>(tmA, tmB, tmC, p, grid, stream);                    // epilogue bound
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace r3m
