// tcgen05 implicit-GEMM 3x3 convolution on channels-last [B,T,F,C] activations (forward and data gradient).
//
//   out[p][n] = act( sum_{tap} sum_c in[p + off(tap)][c] * wk[tap][n][c] + bias[n] ) (.* mask > 0)
//
// One CTA computes an 8 (time) x 16 (freq) patch = 128 output pixels x BN output channels.  The A operand of k-block
// (tap, 32-channel slice) is ONE 4-D TMA box {32 ch, 16 freq, 8 time, 1 utt} fetched at the tap-shifted coordinate:
// the zero padding of the convolution is TMA's out-of-bounds fill, so there is no im2col buffer and no halo logic.
// The box lands in shared memory as 128 rows x 128 B (K-major, 128B swizzle) -- exactly the UMMA canonical tile.
// The B operand is the repacked weight wk[tap][n][c] (K-major, one 2-D box).  Mainloop, 3xTF32 operand split and
// TMEM epilogue are the same design as tc_gemm.cu.
#include <stdlib.h>

#include "../../include/b200asr.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace b200asr {
namespace tc {

constexpr int CT_T = 8, CT_F = 16;                    // pixel patch: 8 x 16 = 128 rows of the MMA
constexpr int A_TILE = 128 * 128;                     // bytes per A k-block
constexpr int CONV_THREADS = 320;                     // warp 0 TMA, warp 1 MMA, warps 2-9 split + epilogue
constexpr int CONV_WORKERS = 256;

template <int BN, int NSPLIT> struct ConvCfg {
  static constexpr int kBTile = BN * 128;
  static constexpr int kStageBytes = NSPLIT * 0 + (NSPLIT == 1 ? (A_TILE + kBTile) : 2 * (A_TILE + kBTile));
  static constexpr int kStages = NSPLIT == 1 ? 6 : (BN == 64 ? 4 : 3);
  static constexpr int kOffAlo = A_TILE;                                   // (x3) A_hi | A_lo | B_hi | B_lo
  static constexpr int kOffBhi = NSPLIT == 1 ? A_TILE : 2 * A_TILE;
  static constexpr int kOffBlo = 2 * A_TILE + kBTile;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct ConvEpi;
struct WgP;
int conv_engine_enabled();
int conv_fwd_persistent(const CUtensorMap& ma, const CUtensorMap& mb, const ConvEpi& e, int precision, cudaStream_t st);
int conv_wgrad_persistent(const CUtensorMap& mx, const CUtensorMap& mdy, const WgP& e, int precision, cudaStream_t st);

struct ConvEpi {
  float* out;
  const float* bias;
  const float* mask;
  int relu, B, T, F, Cin, Cout;
  int halo_bo;   // bring-up switch: 0 = descriptor base_offset 0 for shifted views, 1 = base_offset = freq shift
};

template <int BN, int NSPLIT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvEpi e) {
  using Cfg = ConvCfg<BN, NSPLIT>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto xfm_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * S);
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 1);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * CT_F, t0 = blockIdx.y * CT_T, b = blockIdx.z;
  const int cch = e.Cin / 32;
  const int nkb = 9 * cch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(xfm_bar(s), CONV_WORKERS); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kOffBhi;
        const int tap = kb / cch, c0 = (kb - tap * cch) * 32;
        const int df = tap / 3 - 1, dt = tap % 3 - 1;          // tap = kf*3 + kt of w[Co][Ci][kf(freq)][kt(time)]
        mbar_expect_tx(full_bar(s), A_TILE + (NSPLIT == 1 ? 1 : 2) * Cfg::kBTile);
        tma_load_4d(sa, &mapA, full_bar(s), c0, f0 + df, t0 + dt, b);   // halo / image border = TMA zero fill
        tma_load_2d(sb, &mapB, full_bar(s), c0, tap * BN);
        if (NSPLIT == 3) tma_load_2d(sa + Cfg::kOffBlo, &mapB, full_bar(s), c0, (9 + tap) * BN);   // pre-split lo half
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, false, false);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(NSPLIT == 1 ? full_bar(s) : xfm_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kOffBhi;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t a_hi = make_smem_desc(sa + ks * 32, 16, 1024);
          const uint64_t b_hi = make_smem_desc(sb + ks * 32, 16, 1024);
          const uint32_t acc0 = (kb | ks) != 0 ? 1u : 0u;
          if (NSPLIT == 1) {
            umma_tf32(tmem_base, a_hi, b_hi, idesc, acc0);
          } else {
            const uint64_t a_lo = make_smem_desc(sa + Cfg::kOffAlo + ks * 32, 16, 1024);
            const uint64_t b_lo = make_smem_desc(sa + Cfg::kOffBlo + ks * 32, 16, 1024);
            umma_tf32(tmem_base, a_lo, b_hi, idesc, acc0);
            umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
            umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
  } else {
    const int t = threadIdx.x - 64;
    if (NSPLIT == 3) {
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(full_bar(s), ph);
        float4* stage = reinterpret_cast<float4*>(gen_base + s * Cfg::kStageBytes);
        split_tf32_inplace(stage, stage + Cfg::kOffAlo / 16, A_TILE / 16, t, CONV_WORKERS);   // weights arrive pre-split
        fence_proxy_async_smem();
        mbar_arrive(xfm_bar(s));
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;               // MMA row = pixel (t0 + r/16, f0 + r%16)
    const int tt = t0 + r / CT_F, ff = f0 + r % CT_F;
    const bool ok = tt < e.T && ff < e.F;
    const size_t pix = ((size_t)b * e.T + tt) * e.F + ff;
    float* orow = e.out + pix * e.Cout;
    const float* mrow = e.mask ? e.mask + pix * e.Cout : nullptr;
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const int col = c * 32 + j4 * 4;
        float o[4] = {v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]};
        if (e.bias) {
          const float4 bb = *reinterpret_cast<const float4*>(e.bias + col);
          o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
        }
        if (e.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
        if (mrow) {
          const float4 m = *reinterpret_cast<const float4*>(mrow + col);
          o[0] = m.x > 0.f ? o[0] : 0.f; o[1] = m.y > 0.f ? o[1] : 0.f; o[2] = m.z > 0.f ? o[2] : 0.f; o[3] = m.w > 0.f ? o[3] : 0.f;
        }
        *reinterpret_cast<float4*>(orow + col) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int NSPLIT>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const ConvEpi& e, cudaStream_t st) {
  using Cfg = ConvCfg<BN, NSPLIT>;
  auto* kern = tc_conv3x3_kernel<BN, NSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (r != cudaSuccess) { set_error("tc_conv: cannot reserve %d bytes of shared memory: %s", Cfg::kSmemBytes, cudaGetErrorString(r)); return B200ASR_CUDA_ERROR; }
    attr_set = true;
  }
  dim3 grid(ceil_div(e.F, CT_F), ceil_div(e.T, CT_T), e.B);
  kern<<<grid, CONV_THREADS, Cfg::kSmemBytes, st>>>(ma, mb, e);
  return check_launch("tc_conv3x3");
}

}  // namespace tc

int conv3x3_tc_halo(const float* in, const float* wk, const float* bias, const float* mask, float* out, int B, int T, int F,
                    int Cin, int Cout, int relu, int precision, cudaStream_t st);

// wk: [9][Cout][Cin] K-major weights (conv_repack_k_kernel); for precision 3 the buffer holds [2][9][Cout][Cin]:
// hi = rna_tf32(w) followed by lo = rna_tf32(w - hi), so the kernel only has to split the activation tiles.
int conv3x3_tc(const float* in, const float* wk, const float* bias, const float* mask, float* out, int B, int T, int F,
               int Cin, int Cout, int relu, int precision, cudaStream_t st) {
  using namespace tc;
  B200_REQUIRE(precision == 1 || precision == 3, B200ASR_BAD_ARG, "conv3x3_tc: precision must be 1 or 3");
  B200_REQUIRE(Cin % 32 == 0 && (Cout == 64 || Cout == 128), B200ASR_BAD_SHAPE,
               "conv3x3_tc: needs Cin %% 32 == 0 and Cout in {64,128} (Cin=%d Cout=%d)", Cin, Cout);
  B200_REQUIRE(aligned16(in) && aligned16(wk) && aligned16(out) && (!bias || aligned16(bias)) && (!mask || aligned16(mask)),
               B200ASR_BAD_ALIGN, "conv3x3_tc: pointers must be 16-byte aligned");
  B200_REQUIRE(B <= 65535 && ceil_div(T, CT_T) <= 65535, B200ASR_BAD_SHAPE, "conv3x3_tc: grid too large");
  {
    static const int halo = [] { const char* v = getenv("B200ASR_CONV_HALO"); return v ? atoi(v) : 0; }();
    if (halo) return conv3x3_tc_halo(in, wk, bias, mask, out, B, T, F, Cin, Cout, relu, precision, st);
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
  ConvEpi e{out, bias, mask, relu, B, T, F, Cin, Cout, 0};
  if (conv_engine_enabled()) return conv_fwd_persistent(ma, mb, e, precision, st);
  if (Cout == 64) return precision == 1 ? launch_conv<64, 1>(ma, mb, e, st) : launch_conv<64, 3>(ma, mb, e, st);
  return precision == 1 ? launch_conv<128, 1>(ma, mb, e, st) : launch_conv<128, 3>(ma, mb, e, st);
}

int conv3x3_tc_halo(const float* in, const float* wk, const float* bias, const float* mask, float* out, int B, int T, int F,
                    int Cin, int Cout, int relu, int precision, cudaStream_t st);

}  // namespace b200asr

// =====================================================================================================================
// Weight gradient:  dwr[tap][ci][co] += sum_p x[p + off(tap)][ci] * dy[p][co]
//
// GEMM with the PIXEL axis as the contraction: both operands are MN-major (channels contiguous, pixels strided), i.e.
// exactly the channels-last activations as they lie in HBM -- no transpose.  One k-block = a 2 (time) x 16 (freq) patch
// of 32 pixels; the A tile is four 32-channel chunks {32 ch, 16, 2, 1} fetched at the tap-shifted coordinate (zero fill
// = padding), the B tile Co/32 chunks of dy at the unshifted coordinate.  M = 128 rows: one tap when Cin = 128, a PAIR of
// taps when Cin = 64 (rows 0-63 tap a, 64-127 tap b; they share the dy tile).  Each CTA reduces a contiguous range of
// pixel blocks and adds its 128 x Co partial into dwr with fp32 atomics (dwr is zeroed by the caller).
namespace b200asr {
namespace tc {

constexpr int WG_PF = 16, WG_PT = 2;                 // pixel patch of one k-block (32 pixels)

template <int BN, int NSPLIT> struct WgCfg {
  static constexpr int kATile = 4 * 4096;
  static constexpr int kBTile = (BN / 32) * 4096;
  static constexpr int kStageBytes = (NSPLIT == 1 ? 1 : 2) * (kATile + kBTile);
  static constexpr int kStages = NSPLIT == 1 ? 6 : (BN == 64 ? 4 : 3);
  static constexpr int kOffAlo = kATile;
  static constexpr int kOffBhi = NSPLIT == 1 ? kATile : 2 * kATile;
  static constexpr int kOffBlo = 2 * kATile + kBTile;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct WgP {
  float* dwr;
  int B, T, F, Ci, Co, nft, ntt, blocks_per_cta, total_blocks;
};

template <int CI, int BN, int NSPLIT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
tc_conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDy, const WgP e) {
  using Cfg = WgCfg<BN, NSPLIT>;
  constexpr int S = Cfg::kStages;
  constexpr int TAPS_PER_CTA = 128 / CI;             // 2 (Cin = 64) or 1 (Cin = 128)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto xfm_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * S);
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 1);
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap_a = blockIdx.x * TAPS_PER_CTA;
  const int tap_b = (TAPS_PER_CTA == 2 && tap_a + 1 < 9) ? tap_a + 1 : tap_a;    // odd tap count: slot b repeats tap a, unused
  const int blk0 = blockIdx.y * e.blocks_per_cta;
  const int nkb = max(0, min(e.blocks_per_cta, e.total_blocks - blk0));

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(xfm_bar(s), CONV_WORKERS); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapDy);
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kOffBhi;
        const int blk = blk0 + kb;
        const int ft = blk % e.nft, tt = (blk / e.nft) % e.ntt, b = blk / (e.nft * e.ntt);
        const int f0 = ft * WG_PF, t0 = tt * WG_PT;
        mbar_expect_tx(full_bar(s), Cfg::kATile + Cfg::kBTile);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int tap = (c * 32) / CI == 0 ? tap_a : tap_b;
          const int df = tap / 3 - 1, dt = tap % 3 - 1;
          tma_load_4d(sa + c * 4096, &mapX, full_bar(s), (c * 32) % CI, f0 + df, t0 + dt, b);
        }
#pragma unroll
        for (int c = 0; c < BN / 32; c++) tma_load_4d(sb + c * 4096, &mapDy, full_bar(s), c * 32, f0, t0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, true, true);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(NSPLIT == 1 ? full_bar(s) : xfm_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        const uint32_t sb = sa + Cfg::kOffBhi;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t a_hi = make_smem_desc(sa + ks * 1024, 4096, 512, kLayoutSW128Base32B);
          const uint64_t b_hi = make_smem_desc(sb + ks * 1024, 4096, 512, kLayoutSW128Base32B);
          const uint32_t acc0 = (kb | ks) != 0 ? 1u : 0u;
          if (NSPLIT == 1) {
            umma_tf32(tmem_base, a_hi, b_hi, idesc, acc0);
          } else {
            const uint64_t a_lo = make_smem_desc(sa + Cfg::kOffAlo + ks * 1024, 4096, 512, kLayoutSW128Base32B);
            const uint64_t b_lo = make_smem_desc(sa + Cfg::kOffBlo + ks * 1024, 4096, 512, kLayoutSW128Base32B);
            umma_tf32(tmem_base, a_lo, b_hi, idesc, acc0);
            umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
            umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
  } else {
    const int t = threadIdx.x - 64;
    if (NSPLIT == 3) {
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(full_bar(s), ph);
        float4* stage = reinterpret_cast<float4*>(gen_base + s * Cfg::kStageBytes);
        split_tf32_inplace(stage, stage + Cfg::kOffAlo / 16, Cfg::kATile / 16, t, CONV_WORKERS);
        split_tf32_inplace(stage + Cfg::kOffBhi / 16, stage + Cfg::kOffBlo / 16, Cfg::kBTile / 16, t, CONV_WORKERS);
        fence_proxy_async_smem();
        mbar_arrive(xfm_bar(s));
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;                 // MMA row = (tap slot, input channel)
    const int slot = r / CI, ci = r % CI;
    const int tap = slot == 0 ? tap_a : tap_b;
    const bool ok = nkb > 0 && (slot == 0 || tap_b != tap_a);
    float* orow = e.dwr + ((size_t)tap * e.Ci + ci) * e.Co;
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      float v[32];
      if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
      if (!ok) continue;
#pragma unroll
      for (int j = 0; j < 32; j++) atomicAdd(orow + c * 32 + j, v[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int CI, int BN, int NSPLIT>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdy, WgP e, cudaStream_t st) {
  using Cfg = WgCfg<BN, NSPLIT>;
  auto* kern = tc_conv3x3_wgrad_kernel<CI, BN, NSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (r != cudaSuccess) { set_error("tc_conv_wgrad: cannot reserve %d bytes of shared memory: %s", Cfg::kSmemBytes, cudaGetErrorString(r)); return B200ASR_CUDA_ERROR; }
    attr_set = true;
  }
  const int groups = CI == 64 ? 5 : 9;
  const int sms = device_sm_count();
  int splits = max(1, (4 * sms) / groups);
  e.blocks_per_cta = max(64, ceil_div(e.total_blocks, splits));
  splits = ceil_div(e.total_blocks, e.blocks_per_cta);
  dim3 grid(groups, splits);
  kern<<<grid, CONV_THREADS, Cfg::kSmemBytes, st>>>(mx, mdy, e);
  return check_launch("tc_conv3x3_wgrad");
}

}  // namespace tc

int conv3x3_wgrad_tc(const float* x, const float* dy, float* dwr, int B, int T, int F, int Ci, int Co, int precision,
                     cudaStream_t st) {
  using namespace tc;
  B200_REQUIRE(precision == 1 || precision == 3, B200ASR_BAD_ARG, "conv3x3_wgrad_tc: precision must be 1 or 3");
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
  WgP e{dwr, B, T, F, Ci, Co, ceil_div(F, WG_PF), ceil_div(T, WG_PT), 0, 0};
  const long long total = (long long)B * e.nft * e.ntt;
  B200_REQUIRE(total < (1LL << 31), B200ASR_BAD_SHAPE, "conv3x3_wgrad_tc: too many pixel blocks");
  e.total_blocks = (int)total;
  if (conv_engine_enabled()) return conv_wgrad_persistent(mx, mdy, e, precision, st);
#define WG(CIv, BNv) return precision == 1 ? launch_wgrad<CIv, BNv, 1>(mx, mdy, e, st) : launch_wgrad<CIv, BNv, 3>(mx, mdy, e, st)
  if (Ci == 64 && Co == 64) WG(64, 64);
  if (Ci == 64 && Co == 128) WG(64, 128);
  if (Ci == 128 && Co == 64) WG(128, 64);
  WG(128, 128);
#undef WG
}

}  // namespace b200asr

// =====================================================================================================================
// "Halo" variant of the forward / data-gradient convolution.  The input patch of a 16 (time) x 8 (freq) output tile,
// including its one-pixel halo, is fetched ONCE per 32-channel slice as a {32 ch, 16 freq, 18 time} TMA box
// (288 rows x 128 B; rows are patch pixels in (time, freq) order, 16 per time step so that one time step = 2048 B).
// The A operand of tap (kf, kt) is then just a shifted VIEW of that patch: UMMA descriptor start = patch + (kt*16 + kf)
// rows, 8-row groups 2048 B apart, swizzle phase (base_offset) = kf.  Nine taps reuse one fetch (and, for 3xTF32, one
// operand split) instead of nine shifted fetches: ~4.4x less activation traffic through L2 and 4x less split work.
namespace b200asr {
namespace tc {

constexpr int HT = 16, HF = 8;                        // output tile: 16 time steps x 8 freq bins = 128 MMA rows
constexpr int HPF = 16, HPT = HT + 2;                 // patch box: 16 freq columns (10 used) x 18 time steps
constexpr int PATCH_BYTES = HPF * HPT * 128;          // 36,864 B per 32-channel slice

template <int BN, int NSPLIT> struct HaloCfg {
  static constexpr int kBTile = BN * 128;
  static constexpr int kPatchStages = 2;
  static constexpr int kPatchStageBytes = (NSPLIT == 1 ? 1 : 2) * PATCH_BYTES;      // [hi | lo]
  static constexpr int kBStageBytes = (NSPLIT == 1 ? 1 : 2) * kBTile;               // [hi | lo] (pre-split weights)
  static constexpr int kBStages = NSPLIT == 1 ? 6 : (BN == 64 ? 3 : 2);
  static constexpr int kOffB = kPatchStages * kPatchStageBytes;
  static constexpr int kOffBar = kOffB + kBStages * kBStageBytes;
  static constexpr int kSmemBytes = kOffBar + 256 + 1024;
};

template <int BN, int NSPLIT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
tc_conv3x3_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvEpi e) {
  using Cfg = HaloCfg<BN, NSPLIT>;
  constexpr int PS = Cfg::kPatchStages, BS = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::kOffBar;
  auto p_full = [&](int s) { return bar_base + 8u * s; };
  auto p_xfm = [&](int s) { return bar_base + 8u * (PS + s); };
  auto p_empty = [&](int s) { return bar_base + 8u * (2 * PS + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (3 * PS + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (3 * PS + BS + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * PS + 2 * BS);
  const uint32_t tmem_slot = bar_base + 8u * (3 * PS + 2 * BS + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * HF, t0 = blockIdx.y * HT, b = blockIdx.z;
  const int cch = e.Cin / 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PS; s++) { mbar_init(p_full(s), 1); mbar_init(p_xfm(s), CONV_WORKERS); mbar_init(p_empty(s), 1); }
    for (int s = 0; s < BS; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < cch; c++) {
        const int ps = c % PS;
        mbar_wait(p_empty(ps), ((c / PS) & 1) ^ 1);
        mbar_expect_tx(p_full(ps), PATCH_BYTES);
        tma_load_4d(smem_base + ps * Cfg::kPatchStageBytes, &mapA, p_full(ps), c * 32, f0 - 1, t0 - 1, b);
        for (int tap = 0; tap < 9; tap++, it++) {
          const int bs = it % BS;
          mbar_wait(b_empty(bs), ((it / BS) & 1) ^ 1);
          const uint32_t sb = smem_base + Cfg::kOffB + bs * Cfg::kBStageBytes;
          mbar_expect_tx(b_full(bs), Cfg::kBStageBytes);
          tma_load_2d(sb, &mapB, b_full(bs), c * 32, tap * BN);
          if (NSPLIT == 3) tma_load_2d(sb + Cfg::kBTile, &mapB, b_full(bs), c * 32, (9 + tap) * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, BN, false, false);
      int it = 0;
      for (int c = 0; c < cch; c++) {
        const int ps = c % PS;
        mbar_wait(NSPLIT == 1 ? p_full(ps) : p_xfm(ps), (c / PS) & 1);
        tc_fence_after();
        const uint32_t pa = smem_base + ps * Cfg::kPatchStageBytes;
        for (int tap = 0; tap < 9; tap++, it++) {
          const int bs = it % BS;
          mbar_wait(b_full(bs), (it / BS) & 1);
          tc_fence_after();
          const uint32_t sb = smem_base + Cfg::kOffB + bs * Cfg::kBStageBytes;
          const int kf = tap / 3, kt = tap % 3;                     // tap = kf*3 + kt (freq, time offsets)
          const uint32_t arow = (uint32_t)(kt * HPF + kf) * 128u;    // shifted view into the haloed patch
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint64_t a_hi = make_smem_desc(pa + arow + ks * 32, 16, HPF * 128, kLayoutSW128, e.halo_bo ? (uint32_t)kf : 0u);
            const uint64_t b_hi = make_smem_desc(sb + ks * 32, 16, 1024);
            const uint32_t acc0 = (it | ks) != 0 ? 1u : 0u;
            if (NSPLIT == 1) {
              umma_tf32(tmem_base, a_hi, b_hi, idesc, acc0);
            } else {
              const uint64_t a_lo = make_smem_desc(pa + PATCH_BYTES + arow + ks * 32, 16, HPF * 128, kLayoutSW128, e.halo_bo ? (uint32_t)kf : 0u);
              const uint64_t b_lo = make_smem_desc(sb + Cfg::kBTile + ks * 32, 16, 1024);
              umma_tf32(tmem_base, a_lo, b_hi, idesc, acc0);
              umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
              umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
            }
          }
          umma_commit(b_empty(bs));
        }
        umma_commit(p_empty(ps));
      }
      umma_commit(accum_bar);
    }
  } else {
    const int t = threadIdx.x - 64;
    if (NSPLIT == 3) {
      for (int c = 0; c < cch; c++) {
        const int ps = c % PS;
        mbar_wait(p_full(ps), (c / PS) & 1);
        float4* hi = reinterpret_cast<float4*>(gen_base + ps * Cfg::kPatchStageBytes);
        split_tf32_inplace(hi, hi + PATCH_BYTES / 16, PATCH_BYTES / 16, t, CONV_WORKERS);
        fence_proxy_async_smem();
        mbar_arrive(p_xfm(ps));
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;               // MMA row = pixel (t0 + r/8, f0 + r%8)
    const int tt = t0 + r / HF, ff = f0 + r % HF;
    const bool ok = tt < e.T && ff < e.F;
    const size_t pix = ((size_t)b * e.T + tt) * e.F + ff;
    float* orow = e.out + pix * e.Cout;
    const float* mrow = e.mask ? e.mask + pix * e.Cout : nullptr;
#pragma unroll 1
    for (int c = half; c < BN / 32; c += 2) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++) {
        const int col = c * 32 + j4 * 4;
        float o[4] = {v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]};
        if (e.bias) {
          const float4 bb = *reinterpret_cast<const float4*>(e.bias + col);
          o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
        }
        if (e.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
        if (mrow) {
          const float4 m = *reinterpret_cast<const float4*>(mrow + col);
          o[0] = m.x > 0.f ? o[0] : 0.f; o[1] = m.y > 0.f ? o[1] : 0.f; o[2] = m.z > 0.f ? o[2] : 0.f; o[3] = m.w > 0.f ? o[3] : 0.f;
        }
        *reinterpret_cast<float4*>(orow + col) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int NSPLIT>
static int launch_conv_halo(const CUtensorMap& ma, const CUtensorMap& mb, const ConvEpi& e, cudaStream_t st) {
  using Cfg = HaloCfg<BN, NSPLIT>;
  auto* kern = tc_conv3x3_halo_kernel<BN, NSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (r != cudaSuccess) { set_error("tc_conv_halo: cannot reserve %d bytes of shared memory: %s", Cfg::kSmemBytes, cudaGetErrorString(r)); return B200ASR_CUDA_ERROR; }
    attr_set = true;
  }
  dim3 grid(ceil_div(e.F, HF), ceil_div(e.T, HT), e.B);
  kern<<<grid, CONV_THREADS, Cfg::kSmemBytes, st>>>(ma, mb, e);
  return check_launch("tc_conv3x3_halo");
}

}  // namespace tc

int conv3x3_tc_halo(const float* in, const float* wk, const float* bias, const float* mask, float* out, int B, int T, int F,
                    int Cin, int Cout, int relu, int precision, cudaStream_t st) {
  using namespace tc;
  CUtensorMap ma, mb;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)F, (uint64_t)T, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin, (uint64_t)F * Cin, (uint64_t)T * F * Cin};
    uint32_t box[4] = {32, HPF, HPT, 1};
    int rc = make_tensor_map_f32(&ma, in, 4, dims, strides, box, false, precision == 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)(precision == 3 ? 18 : 9) * Cout};
    uint64_t strides[1] = {(uint64_t)Cin};
    uint32_t box[2] = {32, (uint32_t)Cout};
    int rc = make_tensor_map_f32(&mb, wk, 2, dims, strides, box, false, precision == 1);
    if (rc) return rc;
  }
  static const int halo_mode = [] { const char* v = getenv("B200ASR_CONV_HALO"); return v ? atoi(v) : 0; }();
  ConvEpi e{out, bias, mask, relu, B, T, F, Cin, Cout, halo_mode == 2 ? 1 : 0};
  if (Cout == 64) return precision == 1 ? launch_conv_halo<64, 1>(ma, mb, e, st) : launch_conv_halo<64, 3>(ma, mb, e, st);
  return precision == 1 ? launch_conv_halo<128, 1>(ma, mb, e, st) : launch_conv_halo<128, 3>(ma, mb, e, st);
}

}  // namespace b200asr

// =====================================================================================================================
// Policies for the persistent engine (tc_engine.cuh): same tiles and operands as the kernels above, but one CTA per SM
// walks the tile list with the epilogue of tile i overlapped with the mainloop of tile i+1.
#include "tc_engine.cuh"

namespace b200asr {
namespace tc {

template <int BN_, int NSPLIT>
struct ConvPolicy {
  static constexpr int BN = BN_, kABytes = A_TILE, kBBytes = BN_ * 128;
  static constexpr bool kSplitA = true, kSplitB = false, kAMN = false, kBMN = false;     // weights arrive pre-split
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
  static __device__ __forceinline__ uint64_t a_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 1024, 4096, 512, kLayoutSW128Base32B); }
  static __device__ __forceinline__ uint64_t b_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 1024, 4096, 512, kLayoutSW128Base32B); }
  static __device__ __forceinline__ void store(const Params& p, const Tile& t, int r, int c0, const float (&v)[32]) {
    const int slot = r / CI, ci = r % CI;
    if (slot == 1 && t.tap_b == t.tap_a) return;
    if (t.nkb == 0) return;
    float* orow = p.e.dwr + ((size_t)(slot == 0 ? t.tap_a : t.tap_b) * p.e.Ci + ci) * p.e.Co + c0;
#pragma unroll
    for (int j = 0; j < 32; j++) atomicAdd(orow + j, v[j]);
  }
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

int conv_engine_enabled() {
  static const int engine = [] { const char* v = getenv("B200ASR_ENGINE"); return v ? atoi(v) : 1; }();
  return engine;
}

int conv_fwd_persistent(const CUtensorMap& ma, const CUtensorMap& mb, const ConvEpi& e, int precision, cudaStream_t st) {
  if (e.Cout == 64) return precision == 1 ? launch_conv_persistent<64, 1>(ma, mb, e, st) : launch_conv_persistent<64, 3>(ma, mb, e, st);
  return precision == 1 ? launch_conv_persistent<128, 1>(ma, mb, e, st) : launch_conv_persistent<128, 3>(ma, mb, e, st);
}

int conv_wgrad_persistent(const CUtensorMap& mx, const CUtensorMap& mdy, const WgP& e, int precision, cudaStream_t st) {
#define WGP(CIv, BNv) return precision == 1 ? launch_wgrad_persistent<CIv, BNv, 1>(mx, mdy, e, st) : launch_wgrad_persistent<CIv, BNv, 3>(mx, mdy, e, st)
  if (e.Ci == 64 && e.Co == 64) WGP(64, 64);
  if (e.Ci == 64 && e.Co == 128) WGP(64, 128);
  if (e.Ci == 128 && e.Co == 64) WGP(128, 64);
  WGP(128, 128);
#undef WGP
}

}  // namespace tc
}  // namespace b200asr
