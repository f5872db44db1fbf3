// Fused attention at fp32 grade on tcgen05 (kind::f16, 2-term bf16 split "bf16x3"): the default attention path.
//
// models/common_layers.py:211-225 in ONE forward kernel and TWO backward kernels per call: Q K^T and P V are tensor-core
// tiles fed by TMA into shared memory; scale, key-padding / causal / dense masks, softmax and dropout run on the accumulator
// in TENSOR MEMORY, so neither the scores nor the probabilities ever touch shared or global memory.  The backward recomputes
// the probabilities from Q, K and the forward's log-sum-exp (nothing of size Tq x Tk is stored between the passes).
//
// Precision: every product is hi*hi + hi*lo + lo*hi over x ~ hi + lo (bf16 + bf16 = 16 significant bits): product error
// ~2^-17, measured ~4e-6 on the outputs -- the grade of the 3xTF32 path at twice its MMA rate and half its operand bytes.
//   * Q, K, V (and dO in the backward) are converted ONCE per call by an elementwise pre-pass into compact bf16 [B,H,T,64]
//     hi / lo arrays (the forward's copies are kept for the backward).  A bf16 tile of 64 features is 128 bytes per row:
//     with the 128B swizzle the shared-memory image of a K-major tile (row = token, contraction = feature) and of an
//     MN-major tile (k-line = token, contraction = token) is THE SAME, so one TMA load serves both uses of a tile (the tf32
//     kernels needed two loads with two different swizzles).
//   * P (forward) and P^T / dS^T / dS (backward) are produced by the softmax threads in registers, split into bf16 hi / lo,
//     packed two keys per 32-bit column and written back IN PLACE over the fp32 scores they came from -- chunk-local: the 32
//     score columns of a 32-key chunk become 16 columns of hi pairs followed by 16 columns of lo pairs -- from where they
//     feed the second GEMM as the TMEM-resident A operand (TS-MMA).
//
// Shape rules: dk = dv = 64; forward keeps the whole score row block in TMEM (Tk <= 448).  Other shapes run the materialised
// 3xTF32 path (attention_mat.cu) or the fp32 CUDA-core kernels.
#include <math.h>

#include "../../include/b200asr.h"
#include "attention.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {
namespace tc {

constexpr int F16_D = 64;                 // head dim (dk = dv)
constexpr int F16_TILE = 128 * 128;       // 128 rows x 128 B
constexpr int F16_THREADS = 192;
constexpr int F16_MAX_TK = 448;
constexpr int F16_O_COL = 448;
constexpr int F16_VSTAGES = 3;

// K-major tile (row = token): 8-row groups 1 KB apart; a 16-deep k-step = 32 B along the row
__device__ __forceinline__ uint64_t kmaj_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 32, 16, 1024, kLayoutSW128); }
// the same bytes read as an MN-major tile (k-line = token, 64 features = one 128-byte chunk): a k-step = 16 k-lines = 2 KB
__device__ __forceinline__ uint64_t mnmaj_desc(uint32_t s, int ks) { return make_smem_desc(s + ks * 2048, 4096, 1024, kLayoutSW128); }

// D[128 x N] = A[128 x 64] B[N x 64]^T with both operands K-major bf16 (hi | lo tiles): 4 k-steps x 3 MMAs
__device__ __forceinline__ void score_mma16(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int N) {
  const uint32_t idesc = make_idesc_bf16(128, N, false, false);
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    umma_bf16(d_tmem, kmaj_desc(a_lo, ks), kmaj_desc(b_hi, ks), idesc, ks != 0 ? 1u : 0u);
    umma_bf16(d_tmem, kmaj_desc(a_hi, ks), kmaj_desc(b_lo, ks), idesc, 1u);
    umma_bf16(d_tmem, kmaj_desc(a_hi, ks), kmaj_desc(b_hi, ks), idesc, 1u);
  }
}
// D[128 x 64] (+)= A[tmem, chunk-local packed hi | lo] B[nk k-lines x 64] (MN-major bf16 hi | lo tiles).  A covers key
// chunks starting at TMEM column a_tmem (32 columns per 32 keys: 16 of hi pairs, 16 of lo pairs); nsteps 16-key steps.
__device__ __forceinline__ void grad_mma16(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int nsteps, bool accumulate) {
  const uint32_t idesc = make_idesc_bf16(128, F16_D, false, true);
  for (int ks = 0; ks < nsteps; ks++) {
    const uint32_t a_h = a_tmem + (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8), a_l = a_h + 16;
    umma_bf16_ts(d_tmem, a_l, mnmaj_desc(b_hi, ks), idesc, (accumulate || ks != 0) ? 1u : 0u);
    umma_bf16_ts(d_tmem, a_h, mnmaj_desc(b_lo, ks), idesc, 1u);
    umma_bf16_ts(d_tmem, a_h, mnmaj_desc(b_hi, ks), idesc, 1u);
  }
}
// 32 fp32 values of one row -> 16 packed hi words + 16 packed lo words, written over the 32 columns they were read from
__device__ __forceinline__ void store_split32(uint32_t taddr, const float* v) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; i++) split_bf16_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  tmem_st16u(taddr, hi);
  tmem_st16u(taddr + 16, lo);
  tmem_wait_st();
}

// ---------------------------------------------------------------------------------------------------------------------
// elementwise pre-pass: (B,H,T,64) fp32 view (any batch / head / row strides) -> compact bf16 hi and lo [B,H,T,64]
__global__ void split_bhtd_kernel(const float* __restrict__ src, long long bs, long long hs, long long rs, uint16_t* __restrict__ hi,
                                  uint16_t* __restrict__ lo, int B, int H, int T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 4 features
  const long long total = (long long)B * H * T * (F16_D / 4);
  if (i >= total) return;
  const int c4 = (int)(i % (F16_D / 4));
  const long long row = i / (F16_D / 4);
  const int t = (int)(row % T);
  const int h = (int)((row / T) % H);
  const int b = (int)(row / ((long long)T * H));
  const float4 x = *reinterpret_cast<const float4*>(src + b * bs + h * hs + (long long)t * rs + c4 * 4);
  uint2 ph, pl;
  split_bf16_pair(x.x, x.y, ph.x, pl.x);
  split_bf16_pair(x.z, x.w, ph.y, pl.y);
  *reinterpret_cast<uint2*>(hi + row * F16_D + c4 * 4) = ph;
  *reinterpret_cast<uint2*>(lo + row * F16_D + c4 * 4) = pl;
}

static int split_bhtd(const float* src, long long bs, long long hs, long long rs, uint16_t* hi, uint16_t* lo, int B, int H, int T,
                      cudaStream_t st) {
  const long long total = (long long)B * H * T * (F16_D / 4);
  split_bhtd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(src, bs, hs, rs, hi, lo, B, H, T);
  return check_launch("sdpa_fused_split");
}

static int map16(CUtensorMap* m, const uint16_t* base, int T, int H, int B, int box_rows) {
  uint64_t dims[4] = {(uint64_t)F16_D, (uint64_t)T, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)F16_D, (uint64_t)T * F16_D, (uint64_t)H * T * F16_D};
  uint32_t box[4] = {F16_D, (uint32_t)box_rows, 1, 1};
  return make_tensor_map_bf16(m, base, 4, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: CTA = (utterance, head, 128-query tile)
struct FwdSmem {
  static constexpr int kQ = 2 * F16_TILE;                    // hi | lo
  static constexpr int kKBlock = 2 * F16_TILE;               // 128 keys, hi | lo
  static constexpr int kOffK = kQ;
  static constexpr int kVStage = 2 * 64 * 128;               // 64 keys, hi | lo
  static constexpr int kOffV = kOffK + 4 * kKBlock;
  static constexpr int kOffPad = kOffV + F16_VSTAGES * kVStage;
  static constexpr int kOffBar = kOffPad + 512;
  static constexpr int kBytes = kOffBar + 256 + 1024;
};

__global__ void __launch_bounds__(F16_THREADS, 1)
sdpa_fused_fwd_kernel(const __grid_constant__ CUtensorMap mQh, const __grid_constant__ CUtensorMap mQl,
                      const __grid_constant__ CUtensorMap mKh, const __grid_constant__ CUtensorMap mKl,
                      const __grid_constant__ CUtensorMap mVh, const __grid_constant__ CUtensorMap mVl, const AttnP p) {
  using L = FwdSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base, sK = smem_base + L::kOffK, sV = smem_base + L::kOffV;
  uint8_t* pad_smem = gen_base + L::kOffPad;
  const uint32_t bar_base = smem_base + L::kOffBar;
  auto qk_full = [&](int j) { return bar_base + 8u * j; };
  const uint32_t s_full = bar_base + 8u * 4;
  auto p_ready = [&](int j) { return bar_base + 8u * (5 + j); };
  auto v_full = [&](int s) { return bar_base + 8u * (12 + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (15 + s); };
  const uint32_t o_full = bar_base + 8u * 18;
  const uint32_t tmem_slot = bar_base + 8u * 19;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int Tk = p.Tk;
  const int tk32 = (Tk + 31) & ~31;
  const int nkblk = (tk32 + 127) / 128;
  const int nvblk = (tk32 + 63) / 64;

  if (threadIdx.x == 0) {
    for (int j = 0; j < 4; j++) mbar_init(qk_full(j), 1);
    mbar_init(s_full, 1);
    for (int j = 0; j < 7; j++) mbar_init(p_ready(j), 128);
    for (int s = 0; s < F16_VSTAGES; s++) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    mbar_init(o_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mQh); tma_prefetch_desc(&mKh); tma_prefetch_desc(&mVh);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < nkblk; j++) {
        mbar_expect_tx(qk_full(j), L::kKBlock + (j == 0 ? L::kQ : 0));
        if (j == 0) {
          tma_load_4d(sQ, &mQh, qk_full(0), 0, q0, h, b);
          tma_load_4d(sQ + F16_TILE, &mQl, qk_full(0), 0, q0, h, b);
        }
        tma_load_4d(sK + j * L::kKBlock, &mKh, qk_full(j), 0, j * 128, h, b);
        tma_load_4d(sK + j * L::kKBlock + F16_TILE, &mKl, qk_full(j), 0, j * 128, h, b);
      }
      for (int vb = 0; vb < nvblk; vb++) {
        const int s = vb % F16_VSTAGES;
        const uint32_t ph = (vb / F16_VSTAGES) & 1;
        mbar_wait(v_empty(s), ph ^ 1);
        mbar_expect_tx(v_full(s), L::kVStage);
        tma_load_4d(sV + s * L::kVStage, &mVh, v_full(s), 0, vb * 64, h, b);
        tma_load_4d(sV + s * L::kVStage + 8192, &mVl, v_full(s), 0, vb * 64, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int j = 0; j < nkblk; j++) {
        const int nj = min(128, tk32 - j * 128);
        mbar_wait(qk_full(j), 0);
        tc_fence_after();
        score_mma16(tmem_base + (uint32_t)(j * 128), sQ, sQ + F16_TILE, sK + j * L::kKBlock, sK + j * L::kKBlock + F16_TILE, nj);
      }
      umma_commit(s_full);
      for (int vb = 0; vb < nvblk; vb++) {
        const int s = vb % F16_VSTAGES;
        const uint32_t ph = (vb / F16_VSTAGES) & 1;
        mbar_wait(p_ready(vb), 0);
        mbar_wait(v_full(s), ph);
        tc_fence_after();
        grad_mma16(tmem_base + (uint32_t)F16_O_COL, tmem_base + (uint32_t)(vb * 64), sV + s * L::kVStage, sV + s * L::kVStage + 8192,
                   min(64, tk32 - vb * 64) / 16, vb != 0);
        umma_commit(v_empty(s));
      }
      umma_commit(o_full);
    }
  } else {
    const int t = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int i = t; i < 512; i += 128) pad_smem[i] = (i < Tk && p.key_pad) ? p.key_pad[(size_t)b * Tk + i] : (i < Tk ? 0 : 1);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint8_t* drow = (p.dense && q < p.Tq) ? p.dense + ((size_t)b * p.Tq + q) * Tk : nullptr;
    const int kcausal = p.causal ? q : 0x7fffffff;
    const float scale = p.scale;
    mbar_wait(s_full, 0);
    tc_fence_after();
    const int nchunk = tk32 / 32;
    float m = -INFINITY;
    for (int c = 0; c < nchunk; c++) {
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int key = c * 32 + j;
        const bool masked = pad_smem[key] || key > kcausal || (drow && drow[key]);
        if (!masked) m = fmaxf(m, v[j] * scale);
      }
    }
    float l = 0.f;
    const size_t drop_row = (((size_t)b * p.H + h) * p.Tq + q) * (size_t)Tk;
    for (int c = 0; c < nchunk; c++) {
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)(c * 32), v);
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int key = c * 32 + j;
        const bool masked = pad_smem[key] || key > kcausal || (drow && drow[key]);
        const float s = masked ? -INFINITY : v[j] * scale;
        float e = __expf(s - m);                       // fully masked row: -inf - -inf = NaN, as the reference
        l += e;
        if (p.thresh && key < Tk) e = dropout_keep(p.key, drop_row + key, p.thresh) ? e * p.inv_keep : 0.f;
        v[j] = e;
      }
      store_split32(lane_addr + (uint32_t)(c * 32), v);
      if ((c & 1) || c == nchunk - 1) {
        tc_fence_before();
        mbar_arrive(p_ready(c >> 1));
      }
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const bool ok = q < p.Tq;
    float* orow = p.o + b * p.o_bs + h * p.o_hs + (long long)q * p.o_rs;
#pragma unroll 1
    for (int c = 0; c < F16_D / 32; c++) {
      float v[32];
      tmem_ld32(lane_addr + (uint32_t)(F16_O_COL + c * 32), v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++)
        *reinterpret_cast<float4*>(orow + c * 32 + j4 * 4) =
            make_float4(v[j4 * 4] * inv, v[j4 * 4 + 1] * inv, v[j4 * 4 + 2] * inv, v[j4 * 4 + 3] * inv);
    }
    if (ok && p.lse) p.lse[((size_t)b * p.H + h) * p.Tq + q] = m + logf(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward 1: CTA = (utterance, head, 128-key block); loops over 128-query tiles.  TMEM lanes = keys.
//   S^T = K Q^T, dP^T = V dO^T;  Pd^T and dS^T written back in place (packed hi | lo);  dV += Pd^T dO,  dK += dS^T Q.
struct BwdSmem {
  static constexpr int kPair = 2 * F16_TILE;                 // one 128-row operand, hi | lo
  // dkdv: K, V resident; Q, dO per query tile.   dq: Q, dO resident; K, V per key block.
  static constexpr int kOff1 = kPair, kOff2 = 2 * kPair, kOff3 = 3 * kPair;
  static constexpr int kOffVec = 4 * kPair;                  // lse[128], delta[128] (dkdv) / key_pad bytes (dq)
  static constexpr int kOffBar = kOffVec + 1024;
  static constexpr int kBytes = kOffBar + 128 + 1024;
};

__global__ void __launch_bounds__(F16_THREADS, 1)
sdpa_fused_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap mQh, const __grid_constant__ CUtensorMap mQl,
                           const __grid_constant__ CUtensorMap mKh, const __grid_constant__ CUtensorMap mKl,
                           const __grid_constant__ CUtensorMap mVh, const __grid_constant__ CUtensorMap mVl,
                           const __grid_constant__ CUtensorMap mOh, const __grid_constant__ CUtensorMap mOl, const AttnP p,
                           const float* __restrict__ delta, float* __restrict__ dk_out, float* __restrict__ dv_out) {
  using L = BwdSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base, sV = smem_base + L::kOff1, sQ = smem_base + L::kOff2, sO = smem_base + L::kOff3;
  float* vec = reinterpret_cast<float*>(gen_base + L::kOffVec);
  const uint32_t bar_base = smem_base + L::kOffBar;
  const uint32_t kv_full = bar_base, q_full = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24,
                 tile_done = bar_base + 32, acc_full = bar_base + 40, tmem_slot = bar_base + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nqt = (p.Tq + 127) / 128;
  const int qt0 = p.causal ? min(nqt, k0 / 128) : 0;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1); mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(tile_done, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DV = tmem_base + 256, T_DK = tmem_base + 320;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * L::kPair);
      tma_load_4d(sK, &mKh, kv_full, 0, k0, h, b); tma_load_4d(sK + F16_TILE, &mKl, kv_full, 0, k0, h, b);
      tma_load_4d(sV, &mVh, kv_full, 0, k0, h, b); tma_load_4d(sV + F16_TILE, &mVl, kv_full, 0, k0, h, b);
      for (int qt = qt0, it = 0; qt < nqt; qt++, it++) {
        mbar_wait(tile_done, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * L::kPair);
        tma_load_4d(sQ, &mQh, q_full, 0, qt * 128, h, b); tma_load_4d(sQ + F16_TILE, &mQl, q_full, 0, qt * 128, h, b);
        tma_load_4d(sO, &mOh, q_full, 0, qt * 128, h, b); tma_load_4d(sO + F16_TILE, &mOl, q_full, 0, qt * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(kv_full, 0);
      for (int qt = qt0, it = 0; qt < nqt; qt++, it++) {
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        score_mma16(T_S, sK, sK + F16_TILE, sQ, sQ + F16_TILE, 128);      // S^T  [keys x queries]
        score_mma16(T_DP, sV, sV + F16_TILE, sO, sO + F16_TILE, 128);     // dP^T [keys x queries]
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        grad_mma16(T_DV, T_S, sO, sO + F16_TILE, 8, it != 0);             // dV += Pd^T dO   (dO read as an MN-major tile)
        grad_mma16(T_DK, T_DP, sQ, sQ + F16_TILE, 8, it != 0);            // dK += dS^T Q
        umma_commit(tile_done);
      }
      umma_commit(acc_full);
    }
  } else {
    const int t = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int key = k0 + row;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const bool key_masked = key >= p.Tk || (p.key_pad && p.key_pad[(size_t)b * p.Tk + key]);
    const size_t bh = (size_t)b * p.H + h;
    for (int qt = qt0, it = 0; qt < nqt; qt++, it++) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      {
        const int q = qt * 128 + t;
        vec[t] = q < p.Tq ? p.lse[bh * p.Tq + q] : 0.f;
        vec[128 + t] = q < p.Tq ? delta[bh * p.Tq + q] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        float s[32], dp[32];
        tmem_ld32(T_S + lane_addr + c * 32, s);
        tmem_ld32(T_DP + lane_addr + c * 32, dp);
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int q = qt * 128 + c * 32 + j;
          bool masked = key_masked || q >= p.Tq || (p.causal && key > q);
          if (!masked && p.dense) masked = p.dense[((size_t)b * p.Tq + q) * p.Tk + key] != 0;
          float pr = 0.f, ds = 0.f;
          if (!masked) {
            const float pe = __expf(s[j] * p.scale - vec[c * 32 + j]);
            float keep = 1.f;
            if (p.thresh) keep = dropout_keep(p.key, (bh * p.Tq + q) * (size_t)p.Tk + key, p.thresh) ? p.inv_keep : 0.f;
            pr = pe * keep;
            ds = pe * (dp[j] * keep - vec[128 + c * 32 + j]) * p.scale;
          }
          s[j] = pr;
          dp[j] = ds;
        }
        store_split32(T_S + lane_addr + c * 32, s);
        store_split32(T_DP + lane_addr + c * 32, dp);
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool any = qt0 < nqt;
    const bool ok = key < p.Tk;
    float* dkrow = dk_out + b * p.k_bs + h * p.k_hs + (long long)key * p.k_rs;
    float* dvrow = dv_out + b * p.v_bs + h * p.v_hs + (long long)key * p.v_rs;
#pragma unroll 1
    for (int c = 0; c < F16_D / 32; c++) {
      float v[32];
      if (any) tmem_ld32(T_DV + lane_addr + c * 32, v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++)
        *reinterpret_cast<float4*>(dvrow + c * 32 + j4 * 4) =
            any ? make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int c = 0; c < F16_D / 32; c++) {
      float v[32];
      if (any) tmem_ld32(T_DK + lane_addr + c * 32, v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++)
        *reinterpret_cast<float4*>(dkrow + c * 32 + j4 * 4) =
            any ? make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// backward 2: CTA = (utterance, head, 128-query tile); loops over 128-key blocks.  TMEM lanes = queries.
//   S = Q K^T, dP = dO V^T;  dS written back in place (packed hi | lo);  dQ += dS K.
__global__ void __launch_bounds__(F16_THREADS, 1)
sdpa_fused_bwd_dq_kernel(const __grid_constant__ CUtensorMap mQh, const __grid_constant__ CUtensorMap mQl,
                         const __grid_constant__ CUtensorMap mKh, const __grid_constant__ CUtensorMap mKl,
                         const __grid_constant__ CUtensorMap mVh, const __grid_constant__ CUtensorMap mVl,
                         const __grid_constant__ CUtensorMap mOh, const __grid_constant__ CUtensorMap mOl, const AttnP p,
                         const float* __restrict__ delta, float* __restrict__ dq_out) {
  using L = BwdSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base, sO = smem_base + L::kOff1, sK = smem_base + L::kOff2, sV = smem_base + L::kOff3;
  uint8_t* pad_smem = gen_base + L::kOffVec;
  const uint32_t bar_base = smem_base + L::kOffBar;
  const uint32_t q_full = bar_base, k_full = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24,
                 tile_done = bar_base + 32, acc_full = bar_base + 40, tmem_slot = bar_base + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb_all = (p.Tk + 127) / 128;
  const int nkb = p.causal ? min(nkb_all, (q0 + 127) / 128 + 1) : nkb_all;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(tile_done, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DQ = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * L::kPair);
      tma_load_4d(sQ, &mQh, q_full, 0, q0, h, b); tma_load_4d(sQ + F16_TILE, &mQl, q_full, 0, q0, h, b);
      tma_load_4d(sO, &mOh, q_full, 0, q0, h, b); tma_load_4d(sO + F16_TILE, &mOl, q_full, 0, q0, h, b);
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(tile_done, (kb & 1) ^ 1);
        mbar_expect_tx(k_full, 2 * L::kPair);
        tma_load_4d(sK, &mKh, k_full, 0, kb * 128, h, b); tma_load_4d(sK + F16_TILE, &mKl, k_full, 0, kb * 128, h, b);
        tma_load_4d(sV, &mVh, k_full, 0, kb * 128, h, b); tma_load_4d(sV + F16_TILE, &mVl, k_full, 0, kb * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(k_full, kb & 1);
        tc_fence_after();
        score_mma16(T_S, sQ, sQ + F16_TILE, sK, sK + F16_TILE, 128);      // S  [queries x keys]
        score_mma16(T_DP, sO, sO + F16_TILE, sV, sV + F16_TILE, 128);     // dP [queries x keys]
        umma_commit(s_full);
        mbar_wait(p_ready, kb & 1);
        tc_fence_after();
        grad_mma16(T_DQ, T_DP, sK, sK + F16_TILE, 8, kb != 0);            // dQ += dS K   (K read as an MN-major tile)
        umma_commit(tile_done);
      }
      umma_commit(acc_full);
    }
  } else {
    const int t = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const size_t bh = (size_t)b * p.H + h;
    const bool qok = q < p.Tq;
    const float lse = qok ? p.lse[bh * p.Tq + q] : 0.f;
    const float dl = qok ? delta[bh * p.Tq + q] : 0.f;
    const uint8_t* drow = (p.dense && qok) ? p.dense + ((size_t)b * p.Tq + q) * p.Tk : nullptr;
    for (int kb = 0; kb < nkb; kb++) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      {
        const int key = kb * 128 + t;
        pad_smem[t] = (key >= p.Tk || (p.key_pad && p.key_pad[(size_t)b * p.Tk + key])) ? 1 : 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, kb & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        float s[32], dp[32];
        tmem_ld32(T_S + lane_addr + c * 32, s);
        tmem_ld32(T_DP + lane_addr + c * 32, dp);
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int key = kb * 128 + c * 32 + j;
          const bool masked = !qok || pad_smem[c * 32 + j] || (p.causal && key > q) || (drow && drow[key]);
          float ds = 0.f;
          if (!masked) {
            const float pe = __expf(s[j] * p.scale - lse);
            float keep = 1.f;
            if (p.thresh) keep = dropout_keep(p.key, (bh * p.Tq + q) * (size_t)p.Tk + key, p.thresh) ? p.inv_keep : 0.f;
            ds = pe * (dp[j] * keep - dl) * p.scale;
          }
          dp[j] = ds;
        }
        store_split32(T_DP + lane_addr + c * 32, dp);
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* dqrow = dq_out + b * p.q_bs + h * p.q_hs + (long long)q * p.q_rs;
#pragma unroll 1
    for (int c = 0; c < F16_D / 32; c++) {
      float v[32];
      if (nkb > 0) tmem_ld32(T_DQ + lane_addr + c * 32, v);
      if (!qok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++)
        *reinterpret_cast<float4*>(dqrow + c * 32 + j4 * 4) =
            nkb > 0 ? make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// delta[b,h,q] = sum_c dO * O   (one warp per row)
__global__ void sdpa_fused_delta_kernel(AttnP p, const float* __restrict__ dout, float* __restrict__ delta) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.H * p.Tq;
  if (w >= total) return;
  const int q = w % p.Tq, bh = w / p.Tq, h = bh % p.H, b = bh / p.H;
  const float* og = p.o + b * p.o_bs + h * p.o_hs + (long long)q * p.o_rs;
  const float* dg = dout + b * p.o_bs + h * p.o_hs + (long long)q * p.o_rs;
  float s = 0.f;
  for (int c = lane; c < F16_D; c += 32) s += og[c] * dg[c];
  s = warp_sum(s);
  if (lane == 0) delta[w] = s;
}

struct Ws16 {                  // compact bf16 copies inside the forward workspace
  uint16_t *qh, *ql, *kh, *kl, *vh, *vl;
};
static size_t ws16_elems(int B, int H, int T) { return ((size_t)B * H * T * F16_D + 63) / 64 * 64; }
static Ws16 carve(void* ws, int B, int H, int Tq, int Tk) {
  uint16_t* base = (uint16_t*)ws;
  const size_t nq = ws16_elems(B, H, Tq), nk = ws16_elems(B, H, Tk);
  Ws16 w;
  w.qh = base; w.ql = base + nq; w.kh = base + 2 * nq; w.kl = w.kh + nk; w.vh = w.kl + nk; w.vl = w.vh + nk;
  return w;
}

}  // namespace tc

using namespace tc;

static int fill16(AttnP& p, const float* q, const float* k, const float* v, float* out, float* lse, long long q_bs, long long q_hs,
                  long long q_rs, long long k_bs, long long k_hs, long long k_rs, long long v_bs, long long v_hs, long long v_rs,
                  long long o_bs, long long o_hs, long long o_rs, const uint8_t* key_pad, const uint8_t* dense_mask, int causal, int B,
                  int H, int Tq, int Tk, int dk, int dv, float scale, float p_drop, uint64_t seed, uint64_t offset) {
  B200_REQUIRE(q && k && v && out && lse, B200ASR_BAD_ARG, "sdpa_fused: null pointer");
  B200_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && H <= 65535 && B <= 65535, B200ASR_BAD_SHAPE, "sdpa_fused: bad problem size");
  B200_REQUIRE(dk == F16_D && dv == F16_D, B200ASR_BAD_SHAPE, "sdpa_fused: needs dk = dv = 64 (dk=%d dv=%d); use the materialised path", dk, dv);
  B200_REQUIRE(q_rs % 4 == 0 && k_rs % 4 == 0 && v_rs % 4 == 0 && o_rs % 4 == 0 && q_bs % 4 == 0 && q_hs % 4 == 0 && k_bs % 4 == 0 &&
                   k_hs % 4 == 0 && v_bs % 4 == 0 && v_hs % 4 == 0 && o_bs % 4 == 0 && o_hs % 4 == 0,
               B200ASR_BAD_ALIGN, "sdpa_fused: strides must be multiples of 4 elements");
  B200_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), B200ASR_BAD_ALIGN, "sdpa_fused: pointers must be 16-byte aligned");
  B200_REQUIRE(p_drop >= 0.f && p_drop < 1.f, B200ASR_BAD_ARG, "sdpa_fused: p_drop=%f", p_drop);
  p.q = q; p.k = k; p.v = v; p.o = out; p.lse = lse;
  p.q_bs = q_bs; p.q_hs = q_hs; p.q_rs = q_rs; p.k_bs = k_bs; p.k_hs = k_hs; p.k_rs = k_rs;
  p.v_bs = v_bs; p.v_hs = v_hs; p.v_rs = v_rs; p.o_bs = o_bs; p.o_hs = o_hs; p.o_rs = o_rs;
  p.key_pad = key_pad; p.dense = dense_mask; p.causal = causal;
  p.B = B; p.H = H; p.Tq = Tq; p.Tk = Tk; p.dk = dk; p.dv = dv; p.scale = scale;
  p.thresh = p_drop > 0.f ? dropout_thresh16(p_drop) : 0u;
  p.inv_keep = dropout_inv_keep(p_drop);
  p.key = dropout_key(seed, offset);
  return B200ASR_OK;
}

}  // namespace b200asr

using namespace b200asr;
using namespace b200asr::tc;

extern "C" {

size_t b200asr_sdpa_fused_ws_bytes(int B, int H, int Tq, int Tk) {
  return 2 * (2 * ws16_elems(B, H, Tq) + 4 * ws16_elems(B, H, Tk));
}

size_t b200asr_sdpa_fused_bwd_ws_bytes(int B, int H, int Tq) {
  return 2 * 2 * ws16_elems(B, H, Tq) + sizeof(float) * (((size_t)B * H * Tq + 3) / 4 * 4);
}

int b200asr_sdpa_fused_fwd(const float* q, const float* k, const float* v, long long q_bs, long long q_hs, long long q_rs,
                           long long k_bs, long long k_hs, long long k_rs, long long v_bs, long long v_hs, long long v_rs,
                           const uint8_t* key_pad, const uint8_t* dense_mask, int causal, float* out, long long o_bs,
                           long long o_hs, long long o_rs, float* lse, void* ws16, int B, int H, int Tq, int Tk, int dk, int dv,
                           float scale, float p_drop, uint64_t seed, uint64_t offset, b200asr_stream_t stream) {
  AttnP p;
  int rc = fill16(p, q, k, v, out, lse, q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs, o_bs, o_hs, o_rs, key_pad, dense_mask,
                  causal, B, H, Tq, Tk, dk, dv, scale, p_drop, seed, offset);
  if (rc) return rc;
  B200_REQUIRE(ws16 && aligned16(ws16), B200ASR_BAD_ARG, "sdpa_fused_fwd: workspace");
  B200_REQUIRE(Tk <= F16_MAX_TK, B200ASR_BAD_SHAPE, "sdpa_fused_fwd: Tk=%d exceeds %d resident score columns; use the materialised path", Tk, F16_MAX_TK);
  cudaStream_t st = (cudaStream_t)stream;
  const Ws16 w = carve(ws16, B, H, Tq, Tk);
  if ((rc = split_bhtd(q, q_bs, q_hs, q_rs, w.qh, w.ql, B, H, Tq, st))) return rc;
  if ((rc = split_bhtd(k, k_bs, k_hs, k_rs, w.kh, w.kl, B, H, Tk, st))) return rc;
  if ((rc = split_bhtd(v, v_bs, v_hs, v_rs, w.vh, w.vl, B, H, Tk, st))) return rc;
  CUtensorMap mQh, mQl, mKh, mKl, mVh, mVl;
  if ((rc = map16(&mQh, w.qh, Tq, H, B, 128)) || (rc = map16(&mQl, w.ql, Tq, H, B, 128)) || (rc = map16(&mKh, w.kh, Tk, H, B, 128)) ||
      (rc = map16(&mKl, w.kl, Tk, H, B, 128)) || (rc = map16(&mVh, w.vh, Tk, H, B, 64)) || (rc = map16(&mVl, w.vl, Tk, H, B, 64)))
    return rc;
  static bool attr_set[kMaxDevices] = {};
  if ((rc = ensure_dynamic_smem((const void*)sdpa_fused_fwd_kernel, FwdSmem::kBytes, attr_set, "sdpa_fused_fwd"))) return rc;
  dim3 grid(ceil_div(Tq, 128), H, B);
  sdpa_fused_fwd_kernel<<<grid, F16_THREADS, FwdSmem::kBytes, st>>>(mQh, mQl, mKh, mKl, mVh, mVl, p);
  return check_launch("sdpa_fused_fwd");
}

int b200asr_sdpa_fused_bwd(const float* dout, const float* q, const float* k, const float* v, const float* out, const float* lse,
                           long long q_bs, long long q_hs, long long q_rs, long long k_bs, long long k_hs, long long k_rs,
                           long long v_bs, long long v_hs, long long v_rs, long long o_bs, long long o_hs, long long o_rs,
                           const uint8_t* key_pad, const uint8_t* dense_mask, int causal, float* dq, float* dk_out, float* dv_out,
                           const void* ws16, void* ws_bwd, int B, int H, int Tq, int Tk, int dk, int dv, float scale, float p_drop,
                           uint64_t seed, uint64_t offset, b200asr_stream_t stream) {
  B200_REQUIRE(dout && dq && dk_out && dv_out && ws16 && ws_bwd, B200ASR_BAD_ARG, "sdpa_fused_bwd: null pointer");
  B200_REQUIRE(aligned16(dout) && aligned16(dq) && aligned16(dk_out) && aligned16(dv_out) && aligned16(ws_bwd), B200ASR_BAD_ALIGN, "sdpa_fused_bwd: alignment");
  AttnP p;
  int rc = fill16(p, q, k, v, const_cast<float*>(out), const_cast<float*>(lse), q_bs, q_hs, q_rs, k_bs, k_hs, k_rs, v_bs, v_hs, v_rs,
                  o_bs, o_hs, o_rs, key_pad, dense_mask, causal, B, H, Tq, Tk, dk, dv, scale, p_drop, seed, offset);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const Ws16 w = carve(const_cast<void*>(ws16), B, H, Tq, Tk);
  uint16_t* oh = (uint16_t*)ws_bwd;
  uint16_t* ol = oh + ws16_elems(B, H, Tq);
  float* delta = (float*)(ol + ws16_elems(B, H, Tq));
  if ((rc = split_bhtd(dout, o_bs, o_hs, o_rs, oh, ol, B, H, Tq, st))) return rc;
  const int total = B * H * Tq;
  sdpa_fused_delta_kernel<<<ceil_div(total, 8), 256, 0, st>>>(p, dout, delta);
  if ((rc = check_launch("sdpa_fused_delta"))) return rc;
  CUtensorMap mQh, mQl, mKh, mKl, mVh, mVl, mOh, mOl;
  if ((rc = map16(&mQh, w.qh, Tq, H, B, 128)) || (rc = map16(&mQl, w.ql, Tq, H, B, 128)) || (rc = map16(&mKh, w.kh, Tk, H, B, 128)) ||
      (rc = map16(&mKl, w.kl, Tk, H, B, 128)) || (rc = map16(&mVh, w.vh, Tk, H, B, 128)) || (rc = map16(&mVl, w.vl, Tk, H, B, 128)) ||
      (rc = map16(&mOh, oh, Tq, H, B, 128)) || (rc = map16(&mOl, ol, Tq, H, B, 128)))
    return rc;
  static bool attr_a[kMaxDevices] = {}, attr_b[kMaxDevices] = {};
  if ((rc = ensure_dynamic_smem((const void*)sdpa_fused_bwd_dkdv_kernel, BwdSmem::kBytes, attr_a, "sdpa_fused_bwd_dkdv"))) return rc;
  if ((rc = ensure_dynamic_smem((const void*)sdpa_fused_bwd_dq_kernel, BwdSmem::kBytes, attr_b, "sdpa_fused_bwd_dq"))) return rc;
  sdpa_fused_bwd_dkdv_kernel<<<dim3(ceil_div(Tk, 128), H, B), F16_THREADS, BwdSmem::kBytes, st>>>(mQh, mQl, mKh, mKl, mVh, mVl, mOh, mOl, p,
                                                                                                  delta, dk_out, dv_out);
  if ((rc = check_launch("sdpa_fused_bwd_dkdv"))) return rc;
  sdpa_fused_bwd_dq_kernel<<<dim3(ceil_div(Tq, 128), H, B), F16_THREADS, BwdSmem::kBytes, st>>>(mQh, mQl, mKh, mKl, mVh, mVl, mOh, mOl, p, delta, dq);
  return check_launch("sdpa_fused_bwd_dq");
}

}  // extern "C"
