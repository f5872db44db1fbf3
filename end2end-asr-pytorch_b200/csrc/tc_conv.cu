// tcgen05 implicit-GEMM 3x3 convolution on channels-last [B,T,F,C] activations: forward / data gradient and weight
// gradient, as policies of the persistent tile engine (tc_engine.cuh).
//
//   out[p][n] = act( sum_{tap} sum_c in[p + off(tap)][c] * wk[tap][n][c] + bias[n] ) (.* mask > 0)
//
// One tile = an 8 (time) x 16 (freq) patch = 128 output pixels x BN output channels.  The A operand of k-block
// (tap, 32-channel slice) is ONE 4-D TMA box {32 ch, 16 freq, 8 time, 1 utt} fetched at the tap-shifted coordinate:
// the zero padding of the convolution is TMA's out-of-bounds fill, so there is no im2col buffer and no halo logic.
// The box lands in shared memory as 128 rows x 128 B (K-major, 128B swizzle) -- exactly the UMMA canonical tile.
// The B operand is the repacked weight wk[tap][n][c] (K-major, one 2-D box), pre-split into hi | lo for 3xTF32.
//
// Weight gradient:  dwr[tap][ci][co] += sum_p x[p + off(tap)][ci] * dy[p][co]
// GEMM with the PIXEL axis as the contraction: both operands are MN-major (channels contiguous, pixels strided), i.e.
// exactly the channels-last activations as they lie in HBM -- no transpose.  One k-block = a 2 (time) x 16 (freq) patch
// of 32 pixels; the A tile is four 32-channel chunks {32 ch, 16, 2, 1} fetched at the tap-shifted coordinate (zero fill
// = padding), the B tile Co/32 chunks of dy at the unshifted coordinate.  M = 128 rows: one tap when Cin = 128, a PAIR of
// taps when Cin = 64 (rows 0-63 tap a, 64-127 tap b; they share the dy tile).  Each tile reduces a contiguous range of
// pixel blocks and adds its 128 x Co partial into dwr with fp32 atomics (dwr is zeroed by the caller).  The bias gradient
// (column sums of dy over all pixels) is accumulated by the split warps of the tap-group-0 tiles while the dy tile is in
// shared memory (3xTF32 only), instead of a separate pass that re-reads dy from HBM.
//
// (An earlier non-persistent kernel pair and a "halo" variant -- one 18 x 16 patch fetch shared by the nine taps through
// shifted UMMA descriptor views, base_offset 0 -- were measured and removed: the engine policies below supersede the
// former, and the latter gained only ~8% because the 3xTF32 mainloop is bound by shared-memory bandwidth on the B side.)
#include <stdlib.h>

#include "../../include/b200asr.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_engine.cuh"

namespace b200asr {
namespace tc {

constexpr int CT_T = 8, CT_F = 16;                    // forward pixel patch: 8 x 16 = 128 rows of the MMA
constexpr int A_TILE = 128 * 128;                     // bytes per A k-block
constexpr int WG_PF = 16, WG_PT = 2;                  // weight-gradient pixel patch of one k-block (32 pixels)

struct ConvEpi {
  float* out;
  const float* bias;
  const float* mask;
  int relu, B, T, F, Cin, Cout;
};

struct WgP {
  float* dwr;
  float* dbias;        // [Co] accumulated (caller zeroes), or nullptr
  int B, T, F, Ci, Co, nft, ntt, blocks_per_cta, total_blocks;
};

template <int BN_, int NSPLIT>
struct ConvPolicy {
  static constexpr int BN = BN_, kABytes = A_TILE, kBBytes = BN_ * 128;
  static constexpr bool kSplitA = true, kSplitB = false, kAMN = false, kBMN = false;     // weights arrive pre-split
  static constexpr bool kSumA = false, kSumB = false;
  struct Params { ConvEpi e; int nft, ntt, cch; };
  static __device__ __forceinline__ int num_tiles(const Params& p) { return p.nft * p.ntt * p.e.B; }
  static __device__ __forceinline__ int num_kb(const Params& p, int) { return 9 * p.cch; }
  struct Tile { int f0, t0, b, tap, df, dt, cb; };      // tile origin + k cursor (tap, 32-channel block)
  static __device__ __forceinline__ Tile tile(const Params& p, int tile) {
    Tile t;
    t.f0 = (tile % p.nft) * CT_F;
    const int r = tile / p.nft;
    t.t0 = (r % p.ntt) * CT_T;
    t.b = r / p.ntt;
    t.tap = 0; t.df = -1; t.dt = -1; t.cb = 0;
    return t;
  }
  // k order: tap-major (tap = 3 * (df + 1) + (dt + 1)), 32-channel blocks inside a tap
  static __device__ __forceinline__ void load(const Params& p, Tile& t, const CUtensorMap* mapA, const CUtensorMap* mapB,
                                              uint32_t sa, uint32_t sb, uint32_t sb_lo, uint32_t bar, bool leader) {
    const int c0 = t.cb * 32;
    if (leader) {
      tma_load_4d(sa, mapA, bar, c0, t.f0 + t.df, t.t0 + t.dt, t.b);
      tma_load_2d(sb, mapB, bar, c0, t.tap * BN);
      if (NSPLIT == 3) tma_load_2d(sb_lo, mapB, bar, c0, (9 + t.tap) * BN);
    }
    if (++t.cb == p.cch) {
      t.cb = 0; t.tap++;
      if (++t.dt == 2) { t.dt = -1; t.df++; }
    }
  }
  static __device__ __forceinline__ uint64_t a_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 32, 16, 1024); }
  static __device__ __forceinline__ uint64_t b_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 32, 16, 1024); }
  static __device__ __forceinline__ void store(const Params& p, const Tile& t, int r, int c0, const float (&v)[32]) {
    const ConvEpi& e = p.e;
    const int tt = t.t0 + r / CT_F, ff = t.f0 + r % CT_F;
    if (tt >= e.T || ff >= e.F) return;
    const size_t pix = ((size_t)t.b * e.T + tt) * e.F + ff;
    float* orow = e.out + pix * e.Cout;
    const float* mrow = e.mask ? e.mask + pix * e.Cout : nullptr;
    // all loads of a 16-channel half before its first store (see GemmPolicy::store)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      float4 bb[4], mm[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int col = c0 + half * 16 + q * 4;
        bb[q] = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        mm[q] = mrow ? __ldg(reinterpret_cast<const float4*>(mrow + col)) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int col = c0 + half * 16 + q * 4, j = half * 16 + q * 4;
        float o[4] = {v[j] + bb[q].x, v[j + 1] + bb[q].y, v[j + 2] + bb[q].z, v[j + 3] + bb[q].w};
        if (e.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
        o[0] = mm[q].x > 0.f ? o[0] : 0.f; o[1] = mm[q].y > 0.f ? o[1] : 0.f;
        o[2] = mm[q].z > 0.f ? o[2] : 0.f; o[3] = mm[q].w > 0.f ? o[3] : 0.f;
        *reinterpret_cast<float4*>(orow + col) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
};

template <int CI, int BN_>
struct WgradPolicy {
  static constexpr int BN = BN_, kABytes = 4 * 4096, kBBytes = (BN_ / 32) * 4096;
  static constexpr bool kSplitA = true, kSplitB = true, kAMN = true, kBMN = true;
  static constexpr bool kSumA = false, kSumB = true;       // bias gradient = column sums of the dy tiles
#ifndef WGRAD_CAT
#define WGRAD_CAT 0
#endif
  // Co = 64: [hi*hi | hi*lo] as one N = 128 MMA (tc_engine.cuh).  Measured at cfg2's conv1_2: 1.91 ms with, 1.76 ms without --
  // the wider accumulator leaves 4 instead of 6 stages of A in tensor memory, which costs more than the N = 64 MMAs do.
  static constexpr bool kCat = WGRAD_CAT != 0;
  static constexpr int kGroups = CI == 64 ? 5 : 9;
  struct Params { WgP e; int splits; };
  static __device__ __forceinline__ int num_tiles(const Params& p) { return kGroups * p.splits; }
  static __device__ __forceinline__ int num_kb(const Params& p, int tile) {
    const int blk0 = (tile / kGroups) * p.e.blocks_per_cta;
    return max(0, min(p.e.blocks_per_cta, p.e.total_blocks - blk0));
  }
  static __device__ __forceinline__ void taps(int tile, int& tap_a, int& tap_b) {
    tap_a = (tile % kGroups) * (128 / CI);
    tap_b = (CI == 64 && tap_a + 1 < 9) ? tap_a + 1 : tap_a;
  }
  struct Tile { int tap_a, tap_b, ft, tt, b, nkb; };     // tap pair + pixel-block cursor (ft fastest, then tt, then b)
  static __device__ __forceinline__ Tile tile(const Params& p, int tile) {
    Tile t;
    taps(tile, t.tap_a, t.tap_b);
    const int blk = (tile / kGroups) * p.e.blocks_per_cta;
    t.ft = blk % p.e.nft; t.tt = (blk / p.e.nft) % p.e.ntt; t.b = blk / (p.e.nft * p.e.ntt);
    t.nkb = num_kb(p, tile);
    return t;
  }
  static __device__ __forceinline__ void load(const Params& p, Tile& t, const CUtensorMap* mapX, const CUtensorMap* mapDy,
                                              uint32_t sa, uint32_t sb, uint32_t, uint32_t bar, bool leader) {
    const int f0 = t.ft * WG_PF, t0 = t.tt * WG_PT;
    if (leader) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int tap = (c * 32) / CI == 0 ? t.tap_a : t.tap_b;
        tma_load_4d(sa + c * 4096, mapX, bar, (c * 32) % CI, f0 + tap / 3 - 1, t0 + tap % 3 - 1, t.b);
      }
#pragma unroll
      for (int c = 0; c < BN / 32; c++) tma_load_4d(sb + c * 4096, mapDy, bar, c * 32, f0, t0, t.b);
    }
    if (++t.ft == p.e.nft) {
      t.ft = 0;
      if (++t.tt == p.e.ntt) { t.tt = 0; t.b++; }
    }
  }
  // every pixel block is visited once by the tap-group-0 tiles (over all splits)
  static __device__ __forceinline__ bool want_sums(const Params& p, const Tile& t) { return p.e.dbias != nullptr && t.tap_a == 0 && t.nkb > 0; }
  static __device__ __forceinline__ void sum_b_store(const Params& p, const Tile&, int col, const float4& v) {
    atomicAdd(p.e.dbias + col, v.x); atomicAdd(p.e.dbias + col + 1, v.y);
    atomicAdd(p.e.dbias + col + 2, v.z); atomicAdd(p.e.dbias + col + 3, v.w);
  }
  static __device__ __forceinline__ uint64_t a_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 1024, 4096, 512, kLayoutSW128Base32B); }
  static __device__ __forceinline__ uint64_t b_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 1024, 4096, 512, kLayoutSW128Base32B); }
  // bf16 modes: the dy tile converted by the split warps into bf16 MN-major 64-column chunks (tc_engine.cuh)
  static __device__ __forceinline__ uint64_t b_desc16(uint32_t s, int ks) { return make_smem_desc(s + ks * 2048, 4096, 1024, kLayoutSW128); }
  static __device__ __forceinline__ void store(const Params& p, const Tile& t, int r, int c0, const float (&v)[32]) {
    const int slot = r / CI, ci = r % CI;
    if (slot == 1 && t.tap_b == t.tap_a) return;
    if (t.nkb == 0) return;
    float* orow = p.e.dwr + ((size_t)(slot == 0 ? t.tap_a : t.tap_b) * p.e.Ci + ci) * p.e.Co + c0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) red_add_v4(orow + j, v[j], v[j + 1], v[j + 2], v[j + 3]);     // dwr: 16-byte aligned rows of Co floats
  }
};

// The same weight gradient with the dy operand arriving as bf16 hi | lo "pairs" that its producer (max-pool backward, the
// convolution data gradient) wrote next to the fp32 tensor: [2][B,T,F,Co] bf16, hi first.  The tile is the bf16 MN-major
// image the tensor core reads (64-channel chunks of 4 KB, one 128-byte k-line per pixel, 128B swizzle) straight from TMA
// -- no conversion of B in the kernel, which is what bound the bf16 variant of WgradPolicy (tc_engine.cuh).
template <int CI, int BN_>
struct WgradPairPolicy : WgradPolicy<CI, BN_> {
  using Base = WgradPolicy<CI, BN_>;
  using typename Base::Params;
  using typename Base::Tile;
  static constexpr bool kSplitB = false, kSumB = true;
  static __device__ __forceinline__ void load16(const Params& p, Tile& t, const CUtensorMap* mapX, const CUtensorMap* mapDy16,
                                                uint32_t sa, uint32_t sb, uint32_t sb_lo, uint32_t bar, bool leader, int halves) {
    const int f0 = t.ft * WG_PF, t0 = t.tt * WG_PT;
    if (leader) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int tap = (c * 32) / CI == 0 ? t.tap_a : t.tap_b;
        tma_load_4d(sa + c * 4096, mapX, bar, (c * 32) % CI, f0 + tap / 3 - 1, t0 + tap % 3 - 1, t.b);
      }
#pragma unroll
      for (int c = 0; c < BN_ / 64; c++) {
        tma_load_4d(sb + c * 4096, mapDy16, bar, c * 64, f0, t0, t.b);
        if (halves == 2) tma_load_4d(sb_lo + c * 4096, mapDy16, bar, c * 64, f0, t0, t.b + p.e.B);     // lo = second half of the batch axis
      }
    }
    if (++t.ft == p.e.nft) {
      t.ft = 0;
      if (++t.tt == p.e.ntt) { t.tt = 0; t.b++; }
    }
  }
  static __device__ __forceinline__ void sum_b_store1(const Params& p, const Tile&, int col, float v) { atomicAdd(p.e.dbias + col, v); }
};

template <int BN, int NSPLIT>
static int launch_conv_persistent(const CUtensorMap& ma, const CUtensorMap& mb, const ConvEpi& e, cudaStream_t st) {
  typename ConvPolicy<BN, NSPLIT>::Params p{e, ceil_div(e.F, CT_F), ceil_div(e.T, CT_T), e.Cin / 32};
  const long long tiles = (long long)p.nft * p.ntt * e.B;
  if (tiles >= (1LL << 31)) { set_error("conv3x3_tc: too many tiles"); return B200ASR_BAD_SHAPE; }
  return launch_engine<ConvPolicy<BN, NSPLIT>, NSPLIT>(ma, mb, p, (int)tiles, st, "tc_conv3x3");
}

template <int CI, int BN, int NSPLIT>
static int launch_wgrad_persistent(const CUtensorMap& mx, const CUtensorMap& mdy, WgP e, cudaStream_t st) {
  using Pol = WgradPolicy<CI, BN>;
  const int sms = device_sm_count();
  int splits = max(1, (2 * sms) / Pol::kGroups);
  e.blocks_per_cta = max(64, ceil_div(e.total_blocks, splits));
  splits = ceil_div(e.total_blocks, e.blocks_per_cta);
  typename Pol::Params p{e, splits};
  return launch_engine<Pol, NSPLIT>(mx, mdy, p, Pol::kGroups * splits, st, "tc_conv3x3_wgrad");
}

template <int CI, int BN, int NSPLIT>
static int launch_wgrad_pairs(const CUtensorMap& mx, const CUtensorMap& mdy, WgP e, cudaStream_t st) {
  using Pol = WgradPairPolicy<CI, BN>;
  const int sms = device_sm_count();
  int splits = max(1, (2 * sms) / Pol::kGroups);
  e.blocks_per_cta = max(64, ceil_div(e.total_blocks, splits));
  splits = ceil_div(e.total_blocks, e.blocks_per_cta);
  typename Pol::Params p{e, splits};
  return launch_engine<Pol, NSPLIT>(mx, mdy, p, Pol::kGroups * splits, st, "tc_conv3x3_wgrad_pairs");
}

}  // namespace tc

// wk: [9][Cout][Cin] K-major weights (conv_repack_k_kernel); for precision 3 the buffer holds [2][9][Cout][Cin]:
// hi = rn_tf32(w) followed by lo = w - hi, so the kernel only has to split the activation tiles.
int conv3x3_tc(const float* in, const float* wk, const float* bias, const float* mask, float* out, int B, int T, int F,
               int Cin, int Cout, int relu, int precision, cudaStream_t st) {
  using namespace tc;
  B200_REQUIRE(precision == 1 || precision == 3, B200ASR_BAD_ARG, "conv3x3_tc: precision must be 1 or 3");
  B200_REQUIRE(Cin % 32 == 0 && (Cout == 64 || Cout == 128), B200ASR_BAD_SHAPE,
               "conv3x3_tc: needs Cin %% 32 == 0 and Cout in {64,128} (Cin=%d Cout=%d)", Cin, Cout);
  B200_REQUIRE(aligned16(in) && aligned16(wk) && aligned16(out) && (!bias || aligned16(bias)) && (!mask || aligned16(mask)),
               B200ASR_BAD_ALIGN, "conv3x3_tc: pointers must be 16-byte aligned");
  if (precision == 3) {
    static const int halo = [] { const char* v = getenv("B200ASR_CONV_HALO"); return v ? atoi(v) : 1; }();
    if (halo) return conv3x3_tc_halo(in, wk, bias, mask, out, B, T, F, Cin, Cout, relu, 3, st);
  }
  CUtensorMap ma, mb;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin, (uint64_t)F * Cin, (uint64_t)T * F * Cin};
    uint32_t box[4] = {32, CT_F, CT_T, 1};
    int rc = make_tensor_map_f32(&ma, in, 4, dims, strides, box, false, precision == 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)(precision == 3 ? 18 : 9) * Cout};     // [hi | lo] halves for 3xTF32
    uint64_t strides[1] = {(uint64_t)Cin};
    uint32_t box[2] = {32, (uint32_t)Cout};
    int rc = make_tensor_map_f32(&mb, wk, 2, dims, strides, box, false, precision == 1);
    if (rc) return rc;
  }
  ConvEpi e{out, bias, mask, relu, B, T, F, Cin, Cout};
  if (Cout == 64) return precision == 1 ? launch_conv_persistent<64, 1>(ma, mb, e, st) : launch_conv_persistent<64, 3>(ma, mb, e, st);
  return precision == 1 ? launch_conv_persistent<128, 1>(ma, mb, e, st) : launch_conv_persistent<128, 3>(ma, mb, e, st);
}

// dbias (optional, 3xTF32 only): [Co], zeroed here and accumulated by the kernel; returns with *dbias_done = 1 when the
// kernel took care of it (the caller runs the separate column-sum pass otherwise)
int conv3x3_wgrad_tc(const float* x, const float* dy, float* dwr, int B, int T, int F, int Ci, int Co, int precision,
                     cudaStream_t st, float* dbias, int* dbias_done, const void* dy16) {
  using namespace tc;
  if (dbias_done) *dbias_done = 0;
  if (dy16 && (precision == 6 || precision == 2)) {
    // dy as bf16 pairs [2][B,T,F,Co] written by its producer: B tiles by TMA, no conversion in the kernel
    B200_REQUIRE((Ci == 64 || Ci == 128) && (Co == 64 || Co == 128), B200ASR_BAD_SHAPE, "conv3x3_wgrad_tc: needs Ci, Co in {64,128} (Ci=%d Co=%d)", Ci, Co);
    B200_REQUIRE(aligned16(x) && aligned16(dy16) && aligned16(dwr), B200ASR_BAD_ALIGN, "conv3x3_wgrad_tc: alignment");
    CUtensorMap mx, mdy;
    {
      uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)F, (uint64_t)T, (uint64_t)B};
      uint64_t strides[3] = {(uint64_t)Ci, (uint64_t)F * Ci, (uint64_t)T * F * Ci};
      uint32_t box[4] = {32, WG_PF, WG_PT, 1};
      if (int rc = make_tensor_map_f32(&mx, x, 4, dims, strides, box, true, false)) return rc;
    }
    {
      uint64_t dims[4] = {(uint64_t)Co, (uint64_t)F, (uint64_t)T, (uint64_t)2 * B};
      uint64_t strides[3] = {(uint64_t)Co, (uint64_t)F * Co, (uint64_t)T * F * Co};
      uint32_t box[4] = {64, WG_PF, WG_PT, 1};
      if (int rc = make_tensor_map_bf16(&mdy, dy16, 4, dims, strides, box)) return rc;
    }
    if (dbias) {
      cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)Co, st);
      if (dbias_done) *dbias_done = 1;
    }
    WgP e{dwr, dbias, B, T, F, Ci, Co, ceil_div(F, WG_PF), ceil_div(T, WG_PT), 0, 0};
    const long long total = (long long)B * e.nft * e.ntt;
    B200_REQUIRE(total < (1LL << 31), B200ASR_BAD_SHAPE, "conv3x3_wgrad_tc: too many pixel blocks");
    e.total_blocks = (int)total;
#define WGPP(CIv, BNv) return precision == 6 ? launch_wgrad_pairs<CIv, BNv, 6>(mx, mdy, e, st) : launch_wgrad_pairs<CIv, BNv, 2>(mx, mdy, e, st)
    if (Ci == 64 && Co == 64) WGPP(64, 64);
    if (Ci == 64 && Co == 128) WGPP(64, 128);
    if (Ci == 128 && Co == 64) WGPP(128, 64);
    WGPP(128, 128);
#undef WGPP
  }
  B200_REQUIRE(precision == 1 || precision == 3 || precision == 2 || precision == 6, B200ASR_BAD_ARG, "conv3x3_wgrad_tc: precision must be 1, 3 (tf32) or 2, 6 (bf16)");
  B200_REQUIRE((Ci == 64 || Ci == 128) && (Co == 64 || Co == 128), B200ASR_BAD_SHAPE,
               "conv3x3_wgrad_tc: needs Ci, Co in {64,128} (Ci=%d Co=%d)", Ci, Co);
  B200_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dwr), B200ASR_BAD_ALIGN, "conv3x3_wgrad_tc: alignment");
  CUtensorMap mx, mdy;
  const bool tf32 = precision == 1;
  {
    uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Ci, (uint64_t)F * Ci, (uint64_t)T * F * Ci};
    uint32_t box[4] = {32, WG_PF, WG_PT, 1};
    int rc = make_tensor_map_f32(&mx, x, 4, dims, strides, box, true, tf32);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Co, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Co, (uint64_t)F * Co, (uint64_t)T * F * Co};
    uint32_t box[4] = {32, WG_PF, WG_PT, 1};
    int rc = make_tensor_map_f32(&mdy, dy, 4, dims, strides, box, true, tf32);
    if (rc) return rc;
  }
  const bool fuse_bias = dbias != nullptr && precision != 1;
  if (fuse_bias) {
    cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)Co, st);
    if (dbias_done) *dbias_done = 1;
  }
  WgP e{dwr, fuse_bias ? dbias : nullptr, B, T, F, Ci, Co, ceil_div(F, WG_PF), ceil_div(T, WG_PT), 0, 0};
  const long long total = (long long)B * e.nft * e.ntt;
  B200_REQUIRE(total < (1LL << 31), B200ASR_BAD_SHAPE, "conv3x3_wgrad_tc: too many pixel blocks");
  e.total_blocks = (int)total;
#define WGP(CIv, BNv)                                                                     \
  switch (precision) {                                                                    \
    case 1: return launch_wgrad_persistent<CIv, BNv, 1>(mx, mdy, e, st);                  \
    case 3: return launch_wgrad_persistent<CIv, BNv, 3>(mx, mdy, e, st);                  \
    case 6: return launch_wgrad_persistent<CIv, BNv, 6>(mx, mdy, e, st);                  \
    default: return launch_wgrad_persistent<CIv, BNv, 2>(mx, mdy, e, st);                 \
  }
  if (Ci == 64 && Co == 64) WGP(64, 64);
  if (Ci == 64 && Co == 128) WGP(64, 128);
  if (Ci == 128 && Co == 64) WGP(128, 64);
  WGP(128, 128);
#undef WGP
}

}  // namespace b200asr
