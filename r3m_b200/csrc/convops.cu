#include "convops.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "launch.h"
#include "tmap.h"

namespace r3m {

thread_local int g_pdl_suppress = 0;

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = std::getenv("R3M_PDL");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int* device_error_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

float* device_stat_scratch() {
  static float* buf = nullptr;
  if (!buf) {
    const size_t bytes = kStatScratchFloats * sizeof(float) + kStatTickets * sizeof(int);
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return nullptr;
    cudaMemset(buf, 0, bytes);
  }
  return buf;
}

float* device_wgrad_scratch() {
  static float* buf = nullptr;
  if (!buf && cudaMalloc(&buf, kWgradScratchBytes) != cudaSuccess) return nullptr;
  return buf;
}

static int pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  return 64;
}

std::string plan_conv(const GatherConv& g, ConvPlan* plan) {
  const int eb = g.tf32 ? 4 : 2;          // operand / output element bytes
  const int kelems = g.tf32 ? 32 : 64;    // elements per 128-byte K block
  if (g.C % kelems != 0) return "gather conv: source channels must be a multiple of one 128-byte K block";
  if (g.tf32 && (g.stat_sum || g.out_mode != 0 || g.accumulate))
    return "gather conv: the tf32 tier needs a dense, non-accumulating output without statistics";
  if (g.Cout % 64 != 0) return "gather conv: output channels must be a multiple of 64";
  if (g.ntaps < 1 || g.ntaps > kMaxTaps) return "gather conv: tap count out of range";
  if (g.stride < 1 || g.stride > 8) return "gather conv: stride out of range";
  if (g.ldo % 8 != 0) return "gather conv: output row pitch must be a multiple of 8 elements";
  *plan = {};
  {
    // Tap reuse (halo3x3.cu): 3x3-footprint, stride-1, pad-1 gather convs of 64 -> 64 channels with a dense bf16 output,
    // no accumulation / affine epilogue, statistics (if any) straight into raw accumulators.  R3M_HALO=0 disables.
    static const bool halo_env = !(std::getenv("R3M_HALO") && std::getenv("R3M_HALO")[0] == '0');
    // measured on B200: the shifted views are correct with base offset 0 (TMA and UMMA both derive the swizzle phase from
    // the absolute shared-memory address); setting the descriptor's base-offset field to the start row's phase is WRONG
    static const int halo_bo = std::getenv("R3M_HALO_BO") ? atoi(std::getenv("R3M_HALO_BO")) : 0;
    // two footprints: the 3x3 / pad 1 convs (halo origin (-1, -1), taps (0..2, 0..2)) and the stem's four vertical taps
    // (origin (-2, 0), taps (0..3, 0): one more halo row, no halo columns; R3M_HALO_STEM=0 keeps it on conv_igemm)
    static const bool halo_stem_env = !(std::getenv("R3M_HALO_STEM") && std::getenv("R3M_HALO_STEM")[0] == '0');
    const bool k3 = g.base_h == -1 && g.base_w == -1;
    const bool stem4 = halo_stem_env && g.base_h == -2 && g.base_w == 0;
    const int max_th = k3 ? 2 : 3, max_tw = k3 ? 2 : 0;
    bool ok = halo_env && !g.tf32 && g.C == 64 && g.Cout == 64 && g.stride == 1 && (k3 || stem4) &&
              g.P == g.H && g.Q == g.W && g.W % 8 == 0 && g.H % 4 == 0 && g.ntaps >= 1 && g.ntaps <= 9 &&
              g.out_mode == 0 && g.ldo == 64 && !g.accumulate && g.ep_scale == nullptr &&
              (g.stat_sum == nullptr || g.stat_raw) && g.bn == 0;
    for (int t = 0; ok && t < g.ntaps; ++t)
      ok = g.tap_h[t] >= 0 && g.tap_h[t] <= max_th && g.tap_w[t] >= 0 && g.tap_w[t] <= max_tw;
    if (ok) {
      const int halo_rows = 16 + max_th;
      std::string e = encode_tiled_4d_map(&plan->tmA, g.src, 64, g.W, g.H, g.N, 64, 10, halo_rows);
      if (e.empty())
        e = encode_tiled_2d_map(&plan->tmB, g.wpk, (uint64_t)g.ntaps * 64, 64, (uint64_t)g.ntaps * 64 * 2, 64, 64, 128, 2);
      if (e.empty()) e = encode_tiled_4d_map(&plan->tmC, g.out, 64, g.W, g.H, g.N, 64, 8, 4);
      if (!e.empty()) return e;
      HaloKernelParams& hp = plan->hp;
      hp.N = g.N;
      hp.H = g.H;
      hp.W = g.W;
      hp.tiles_h = (g.H + 15) / 16;
      hp.tiles_w = g.W / 8;
      hp.num_taps = g.ntaps;
      for (int t = 0; t < g.ntaps; ++t) {
        hp.tap_w[t] = (uint16_t)g.tap_w[t];
        hp.tap_h[t] = (uint16_t)g.tap_h[t];
      }
      hp.org_h = g.base_h;
      hp.org_w = g.base_w;
      hp.patch_bytes = halo_rows * 10 * 128;
      hp.rev = g.rev_m;
      hp.stat_acc = reinterpret_cast<unsigned long long*>(g.stat_sum);
      hp.base_offset_mode = halo_bo;
      hp.acc_stages = std::getenv("R3M_HALO_ACC") ? std::max(2, std::min(8, atoi(std::getenv("R3M_HALO_ACC")))) : 8;
      hp.debug = std::getenv("R3M_HALO_DEBUG") ? atoi(std::getenv("R3M_HALO_DEBUG")) : 0;
      hp.error_flag = device_error_flag();
      if (!hp.error_flag) return "could not allocate the device error flag";
      plan->halo = 1;
      plan->grid = std::min(hp.N * hp.tiles_h * hp.tiles_w, device_sm_count());
      return std::string();
    }
  }
  const int upper_w = (g.Q - 1) * g.stride + 1 + g.base_w - g.W;
  const int upper_h = (g.P - 1) * g.stride + 1 + g.base_h - g.H;
  std::string err = encode_im2col_map(&plan->tmA, g.src, g.C, g.W, g.H, g.N, g.base_w, g.base_h, upper_w, upper_h,
                                      kelems, 128, g.stride, eb);
  if (!err.empty()) return err;
  int bn = pick_bn(g.Cout);
  if (g.bn != 0) {
    if ((g.bn != 64 && g.bn != 128 && g.bn != 256) || g.Cout % g.bn != 0) return "gather conv: bad N tile override";
    bn = g.bn;
  }
  if (bn < 256 && g.Cout > 1024) return "gather conv: N tiles of 64 / 128 serve at most 1024 output channels per launch";
  const uint64_t kdim = (uint64_t)g.ntaps * g.C;
  ConvKernelParams& p = plan->p;
  p.M_total = g.N * g.P * g.Q;
  // CTA pairs (cta_group::2): bf16, dense outputs, 256-wide tiles, at least one pair of M tiles per filter tile
  // (R3M_CONV_PAIR=0: single-CTA kernel everywhere)
  static const bool pair_env = !(std::getenv("R3M_CONV_PAIR") && std::getenv("R3M_CONV_PAIR")[0] == '0');
  // ... and a reduction axis of at least 8 K blocks: measured per launch in the ResNet-50 step, pairs win 6-12 % on the
  // fill-bound shapes (3x3 convs and 4C -> C 1x1 convs of layers 3-4, K >= 512) and LOSE 10-35 % on the short-K,
  // wide-output 1x1 convs (C -> 4C), which are bound by their epilogue and only pay for the lock-step of two CTAs
  static const int pair_min_kb = std::getenv("R3M_CONV_PAIR_MIN_KB") ? atoi(std::getenv("R3M_CONV_PAIR_MIN_KB")) : 8;
  // 128-wide tiles (the 128-channel layers): only the 3x3 configuration (>= 18 K blocks) has a pair variant
  static const bool pair128 = !(std::getenv("R3M_CONV_PAIR128") && std::getenv("R3M_CONV_PAIR128")[0] == '0');
  static const bool pair64 = std::getenv("R3M_CONV_PAIR64") && std::getenv("R3M_CONV_PAIR64")[0] == '1';  // experiment
  const int nkb = g.ntaps * (g.C / kelems);
  p.pair = (pair_env && !g.tf32 && g.out_mode == 0 && p.M_total > 128 &&
            ((bn == 256 && nkb >= pair_min_kb) || (bn == 128 && pair128 && nkb >= 18) || (bn == 64 && pair64))) ? 1 : 0;
  err = encode_tiled_2d_map(&plan->tmB, g.wpk, kdim, (uint64_t)g.Cout, kdim * eb, kelems, p.pair ? bn / 2 : bn, 128, eb);
  if (!err.empty()) return err;
  if (g.out_mode == 0) {
    err = encode_tiled_2d_map(&plan->tmC, g.out, (uint64_t)g.Cout, (uint64_t)p.M_total, (uint64_t)g.ldo * eb, kelems, 32,
                              /*swizzle_bytes=*/128, eb);
    if (!err.empty()) return err;
  } else {
    plan->tmC = plan->tmB;
  }
  p.PQ = g.P * g.Q;
  p.Q = g.Q;
  p.stride = g.stride;
  p.base_w = g.base_w;
  p.base_h = g.base_h;
  p.num_taps = g.ntaps;
  p.cblocks = g.C / kelems;
  p.tf32 = g.tf32;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.tap_w[t] < 0 || g.tap_h[t] < 0) return "gather conv: tap offsets must be non-negative";
    p.tap_w[t] = (uint16_t)g.tap_w[t];
    p.tap_h[t] = (uint16_t)g.tap_h[t];
  }
  p.Cout = g.Cout;
  p.num_m_tiles = (p.M_total + 127) / 128;
  p.num_n_tiles = g.Cout / bn;
  p.out = g.out;
  p.out_mode = g.out_mode;
  p.ldo = g.ldo;
  p.oH = g.oH;
  p.oW = g.oW;
  p.o_stride = g.o_stride;
  p.o_h0 = g.o_h0;
  p.o_w0 = g.o_w0;
  p.accumulate = g.accumulate;
  p.rev_m = g.rev_m;
  p.stat_sum = g.stat_sum;
  p.stat_sq = g.stat_sq;
  p.stat_scratch = g.stat_scratch;
  p.stat_ticket = g.stat_ticket;
  p.stat_raw = g.stat_raw;
  p.stat_rows = g.stat_raw ? 0 : g.stat_rows;
  if (g.stat_sum && !g.stat_scratch && !g.stat_raw) {
    float* shared = device_stat_scratch();
    if (!shared) return "could not allocate the statistics scratch";
    p.stat_scratch = shared;
    p.stat_ticket = reinterpret_cast<int*>(shared + kStatScratchFloats);
  }
  if (g.ep_scale && (g.stat_sum || g.out_mode != 0 || g.accumulate))
    return "gather conv: the fused affine epilogue needs a dense, non-accumulating output without statistics";
  p.ep_scale = g.ep_scale;
  p.ep_shift = g.ep_shift;
  p.ep_res = g.ep_res;
  p.ep_relu = g.ep_relu;
  p.ep_exact = g.ep_exact;
  if ((g.ep_relu == 2 || g.ep_exact) && !g.tf32) return "gather conv: GELU / exact outputs exist in the tf32 tier only";
  p.error_flag = device_error_flag();
  if (!p.error_flag) return "could not allocate the device error flag";
  plan->bn = bn;
  plan->grid = std::min(p.num_m_tiles * p.num_n_tiles, device_sm_count());
  if (p.pair) {
    int pairs = std::min(((p.num_m_tiles + 1) / 2) * p.num_n_tiles, device_sm_count() / 2);
    if (g.stat_sum) pairs -= pairs % p.num_n_tiles;  // every pair keeps one column block
    if (pairs < 1) return "gather conv: no room for a CTA pair per column block";
    plan->grid = 2 * pairs;
  } else if (g.stat_sum) {
    // deterministic statistics: every CTA keeps one column block (see conv_igemm.cu)
    plan->grid -= plan->grid % p.num_n_tiles;
  }
  if (g.stat_sum && ((size_t)g.Cout * 16 > kStatScratchFloats || p.num_n_tiles > (int)kStatTickets))
    return "gather conv: statistics scratch too small for this grid";
  return std::string();
}

cudaError_t run_conv(const ConvPlan& plan, cudaStream_t stream) {
  if (plan.halo) return halo3x3_launch(plan.tmA, plan.tmB, plan.tmC, plan.hp, plan.grid, stream);
  return conv_igemm_launch(plan.bn, plan.tmA, plan.tmB, plan.tmC, plan.p, plan.grid, stream);
}

std::string plan_wgrad(const WgradDesc& d, WgradPlan* plan) {
  if (d.C % 64 != 0 || d.Cout % 64 != 0) return "wgrad: channel counts must be multiples of 64";
  if (d.ntaps < 1 || d.ntaps > kMaxTaps) return "wgrad: tap count out of range";
  *plan = {};
  const int M = d.N * d.P * d.Q;
  int pix = 128;
  if (const char* e = getenv("R3M_WGRAD_PIX")) pix = atoi(e);  // tuning aid
  // CTA pairs (cta_group::2) wherever two adjacent 128-row k-slabs exist and the items split evenly: each CTA then
  // stages half of the group's activation items (R3M_WGRAD_PAIR=0: single-CTA kernel everywhere)
  static const bool pair_env = !(getenv("R3M_WGRAD_PAIR") && getenv("R3M_WGRAD_PAIR")[0] == '0');
  const bool pair = pair_env && d.Cout % 256 == 0 && (d.ntaps * (d.C / 64)) % 2 == 0;
  if (pair)
    if (const char* e = getenv("R3M_WGRAD_PAIR_PIX")) pix = atoi(e);  // tuning aid
  if (pix != 64 && pix != 128) return "wgrad: pix_block must be 64 or 128";
  std::string err =
      encode_tiled_2d_map(&plan->tmDy, d.dy, (uint64_t)d.Cout, (uint64_t)M, (uint64_t)d.Cout * 2, 64, pix);
  if (!err.empty()) return err;
  const int upper_w = (d.Q - 1) * d.stride + 1 + d.base_w - d.W;
  const int upper_h = (d.P - 1) * d.stride + 1 + d.base_h - d.H;
  err = encode_im2col_map(&plan->tmX, d.x, d.C, d.W, d.H, d.N, d.base_w, d.base_h, upper_w, upper_h, 64, pix, d.stride);
  if (!err.empty()) return err;
  WgradKernelParams& p = plan->p;
  p.M_total = M;
  p.PQ = d.P * d.Q;
  p.Q = d.Q;
  p.stride = d.stride;
  p.base_w = d.base_w;
  p.base_h = d.base_h;
  p.cblocks = d.C / 64;
  p.num_items = d.ntaps * p.cblocks;
  // 5 is the largest group whose two pipeline stages of 128-pixel atoms fit in shared memory; the items are spread
  // evenly over ceil(items / 5) groups (16 items -> 4 x 4, not 5 + 5 + 5 + 1: the CTAs of a short group idle while the
  // others finish; measured 4.49 -> 4.37 ms over the 53 wgrads of ResNet-50).
  int group_pref = -5;
  if (const char* e = getenv("R3M_WGRAD_GROUP")) group_pref = atoi(e);  // tuning aid (positive: fixed group size)
  if (group_pref < 0) {
    const int groups = (p.num_items - group_pref - 1) / (-group_pref);
    p.group = (p.num_items + groups - 1) / groups;
  } else {
    p.group = std::min(group_pref, p.num_items);
  }
  p.pair = pair ? 1 : 0;
  if (p.pair) {
    int max_items = 8;  // per pair and group: 4 atoms per CTA, one N = 256 instruction per 4 items
    if (const char* e = getenv("R3M_WGRAD_PAIR_ITEMS")) max_items = std::max(2, std::min(8, atoi(e) & ~1));  // tuning aid
    const int groups = (p.num_items + max_items - 1) / max_items;
    p.group = (p.num_items + groups - 1) / groups;
    p.group += p.group & 1;  // even: the pair splits every group in two
  }
  for (int t = 0; t < d.ntaps; ++t) {
    p.tap_w[t] = (uint16_t)d.tap_w[t];
    p.tap_h[t] = (uint16_t)d.tap_h[t];
  }
  p.Cout = d.Cout;
  p.Cin = d.C;
  p.ldw = d.ntaps * d.C;
  p.pix_block = pix;
  p.mblocks_total = (M + pix - 1) / pix;
  p.num_stages = std::min(6, (227 * 1024 - 2048) / ((2 + (p.pair ? p.group / 2 : p.group)) * pix * 128));
  if (p.num_stages < 2) return "wgrad: group too large for the shared-memory budget";
  if (const char* e = getenv("R3M_WGRAD_STAGES")) p.num_stages = std::min(p.num_stages, atoi(e));
  p.dW = d.dw;
  p.error_flag = device_error_flag();
  if (!p.error_flag) return "could not allocate the device error flag";
  plan->groups = (p.num_items + p.group - 1) / p.group;
  plan->ktiles = (d.Cout + 127) / 128;
  const int slabs = plan->groups * plan->ktiles;
  // split-K so that the grid is at most `waves` full waves of CTAs (rounding DOWN: a grid slightly above a multiple of
  // the SM count pays a whole extra wave)
  int waves = 1;
  if (const char* e = getenv("R3M_WGRAD_WAVES")) waves = atoi(e);
  int splits = (waves * device_sm_count()) / slabs;
  splits = std::max(1, std::min(splits, p.mblocks_total));
  p.mblocks_per_split = (p.mblocks_total + splits - 1) / splits;
  plan->splits = (p.mblocks_total + p.mblocks_per_split - 1) / p.mblocks_per_split;
  p.dw_elems = (size_t)d.Cout * p.ldw;
  p.scratch = nullptr;
  if (plan->splits > 1) {
    // deterministic split-K: shrink the split count until the partial copies of dW fit the scratch
    while (plan->splits > 1 && (size_t)plan->splits * p.dw_elems * 4 > kWgradScratchBytes) {
      p.mblocks_per_split = (p.mblocks_total + plan->splits - 2) / (plan->splits - 1);
      plan->splits = (p.mblocks_total + p.mblocks_per_split - 1) / p.mblocks_per_split;
    }
    if (plan->splits > 1) {
      p.scratch = d.scratch ? d.scratch : device_wgrad_scratch();
      if (!p.scratch) return "wgrad: could not allocate the split-K scratch";
      if (p.dw_elems % 4 != 0) return "wgrad: |dW| must be a multiple of 4";
    }
  }
  return std::string();
}

cudaError_t run_wgrad(const WgradPlan& plan, cudaStream_t stream) {
  return wgrad_launch(plan.tmDy, plan.tmX, plan.p, plan.splits, plan.groups, plan.ktiles, stream);
}

void fill_fwd_geometry(GatherConv* g, int R, int S, int stride, int pad) {
  g->P = (g->H + 2 * pad - R) / stride + 1;
  g->Q = (g->W + 2 * pad - S) / stride + 1;
  g->stride = stride;
  g->base_h = -pad;
  g->base_w = -pad;
  g->ntaps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      g->tap_h[r * S + s] = r;
      g->tap_w[r * S + s] = s;
    }
}

void fill_fwd_geometry(WgradDesc* d, int R, int S, int stride, int pad) {
  d->P = (d->H + 2 * pad - R) / stride + 1;
  d->Q = (d->W + 2 * pad - S) / stride + 1;
  d->stride = stride;
  d->base_h = -pad;
  d->base_w = -pad;
  d->ntaps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      d->tap_h[r * S + s] = r;
      d->tap_w[r * S + s] = s;
    }
}

namespace {
struct AxisTap {
  int src;  // forward filter index
  int d;    // dY offset relative to the class index i
};
// taps along one axis for output parity `par`: forward index r contributes iff (par + pad - r) is a multiple of stride
std::vector<AxisTap> axis_taps(int par, int R, int stride, int pad) {
  std::vector<AxisTap> v;
  for (int r = 0; r < R; ++r) {
    const int num = par + pad - r;
    int q = num / stride;
    if (q * stride != num) continue;
    v.push_back({r, q});
  }
  return v;
}
}  // namespace

std::vector<DgradClass> dgrad_classes(int H, int W, int R, int S, int stride, int pad) {
  std::vector<DgradClass> out;
  for (int ph = 0; ph < stride; ++ph)
    for (int pw = 0; pw < stride; ++pw) {
      DgradClass c;
      c.ph = ph;
      c.pw = pw;
      c.Pc = (H - ph + stride - 1) / stride;
      c.Qc = (W - pw + stride - 1) / stride;
      const std::vector<AxisTap> th = axis_taps(ph, R, stride, pad);
      const std::vector<AxisTap> tw = axis_taps(pw, S, stride, pad);
      c.ntaps = 0;
      if (!th.empty() && !tw.empty()) {
        int bh = th[0].d, bw = tw[0].d;
        for (const AxisTap& a : th) bh = std::min(bh, a.d);
        for (const AxisTap& a : tw) bw = std::min(bw, a.d);
        c.base_h = bh;
        c.base_w = bw;
        for (const AxisTap& a : th)
          for (const AxisTap& b : tw) {
            c.tap_h[c.ntaps] = a.d - bh;
            c.tap_w[c.ntaps] = b.d - bw;
            c.src_r[c.ntaps] = a.src;
            c.src_s[c.ntaps] = b.src;
            ++c.ntaps;
          }
      }
      out.push_back(c);
    }
  return out;
}

}  // namespace r3m
