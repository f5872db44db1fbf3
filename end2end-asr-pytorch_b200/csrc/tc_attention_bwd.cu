// Fused attention backward on tcgen05 (precision 1).  Two kernels, both recompute the probabilities from Q, K and the
// forward's LSE so nothing of size Tq x Tk is ever stored:
//
//   tc_sdpa_bwd_dkdv : CTA = (utterance, head, 128-key block); loops over 128-query tiles.
//        S^T = K Q^T, dP^T = V dO^T (SS-MMA, TMEM lanes = keys); per-thread (= key row) softmax-backward in TMEM:
//        Pd^T = P^T o keep -> TMEM (in place), dS^T = P^T o (dP^T o keep - delta) * scale -> TMEM (in place);
//        dV += Pd^T dO,  dK += dS^T Q   (TS-MMA: A straight from TMEM, B = dO / Q as MN-major tiles).
//   tc_sdpa_bwd_dq   : CTA = (utterance, head, 128-query tile); loops over 128-key blocks.
//        S = Q K^T, dP = dO V^T (lanes = queries); dS -> TMEM; dQ += dS K (TS-MMA, B = K as an MN-major tile).
//
// Operands arrive through TFLOAT32 tensor maps (rounded to nearest in flight), the same rounding the forward kernel
// uses, so delta = rowsum(dO o O) computed from the forward's O is consistent with the recomputed P.  Tiles that are
// needed both K-major (as the N operand of a score GEMM) and MN-major (as the B operand of a gradient GEMM) are fetched
// twice with the two swizzles tf32 requires (128B / 128B-with-32B-atoms).
#include <math.h>

#include "../../include/b200asr.h"
#include "attention.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {
namespace tc {

constexpr int AB_THREADS = 192;

__device__ __forceinline__ void load_kmajor(uint32_t dst, const CUtensorMap* m, uint32_t bar, int D, int row0, int h, int b) {
  for (int sub = 0; sub < D / 32; sub++) tma_load_4d(dst + sub * 16384, m, bar, sub * 32, row0, h, b);
}
__device__ __forceinline__ void load_mnmajor(uint32_t dst, const CUtensorMap* m, uint32_t bar, int D, int row0, int h, int b) {
  for (int ch = 0; ch < D / 32; ch++) tma_load_4d(dst + ch * 16384, m, bar, ch * 32, row0, h, b);
}
// SS score GEMM: D[128 x 128] = A[128 x Dk] * B[128 x Dk]^T, both K-major
__device__ __forceinline__ void score_mma(uint32_t d_tmem, uint32_t sa, uint32_t sb, int D) {
  const uint32_t idesc = make_idesc_tf32(128, 128, false, false);
  for (int ks = 0; ks < D / 8; ks++) {
    const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
    umma_tf32(d_tmem, make_smem_desc(sa + off, 16, 1024), make_smem_desc(sb + off, 16, 1024), idesc, ks != 0 ? 1u : 0u);
  }
}
// TS gradient GEMM: D[128 x N] (+)= A[tmem: 128 lanes x 128 cols] * B[128 k-lines x N] (MN-major chunks of 16 KB)
__device__ __forceinline__ void grad_mma(uint32_t d_tmem, uint32_t a_tmem, uint32_t sb, int N, bool accumulate) {
  const uint32_t idesc = make_idesc_tf32(128, N, false, true);
  for (int ks = 0; ks < 16; ks++)
    umma_tf32_ts(d_tmem, a_tmem + ks * 8, make_smem_desc(sb + ks * 1024, 16384, 512, kLayoutSW128Base32B), idesc,
                 (accumulate || ks != 0) ? 1u : 0u);
}

template <int DK, int DV>
struct BwdCfg {
  static constexpr int kK = (DK / 32) * 16384, kV = (DV / 32) * 16384;
  // dkdv kernel: K (kmaj), V (kmaj) resident; Qk, Qm, Ok, Om per query tile
  static constexpr int kA_offV = kK, kA_offQk = kK + kV, kA_offQm = kA_offQk + kK, kA_offOk = kA_offQm + kK, kA_offOm = kA_offOk + kV;
  static constexpr int kA_offVec = kA_offOm + kV;                 // lse[128], delta[128]
  static constexpr int kA_offBar = kA_offVec + 1024;
  static constexpr int kA_smem = kA_offBar + 128 + 1024;
  // dq kernel: Q (kmaj), dO (kmaj) resident; Kk, Km, Vk per key block
  static constexpr int kB_offO = kK, kB_offKk = kK + kV, kB_offKm = kB_offKk + kK, kB_offVk = kB_offKm + kK;
  static constexpr int kB_offPad = kB_offVk + kV;                 // key_pad bytes for the current key block (128)
  static constexpr int kB_offBar = kB_offPad + 128;
  static constexpr int kB_smem = kB_offBar + 128 + 1024;
};

// ---------------------------------------------------------------------------------------------------------------------
template <int DK, int DV>
__global__ void __launch_bounds__(AB_THREADS, 1)
tc_sdpa_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap mapQk, const __grid_constant__ CUtensorMap mapQm,
                        const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                        const __grid_constant__ CUtensorMap mapOk, const __grid_constant__ CUtensorMap mapOm, const AttnP p,
                        const float* __restrict__ delta, float* __restrict__ dk_out, float* __restrict__ dv_out) {
  using Cfg = BwdCfg<DK, DV>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base, sV = smem_base + Cfg::kA_offV, sQk = smem_base + Cfg::kA_offQk, sQm = smem_base + Cfg::kA_offQm,
                 sOk = smem_base + Cfg::kA_offOk, sOm = smem_base + Cfg::kA_offOm;
  float* vec = reinterpret_cast<float*>(gen_base + Cfg::kA_offVec);     // [0,128) lse, [128,256) delta
  const uint32_t bar_base = smem_base + Cfg::kA_offBar;
  const uint32_t kv_full = bar_base, q_full = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24,
                 tile_done = bar_base + 32, acc_full = bar_base + 40, tmem_slot = bar_base + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nqt = (p.Tq + 127) / 128;
  const int qt0 = p.causal ? min(nqt, k0 / 128) : 0;           // query tiles entirely before the key block see none of it

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
      mbar_expect_tx(kv_full, Cfg::kK + Cfg::kV);
      load_kmajor(sK, &mapK, kv_full, DK, k0, h, b);
      load_kmajor(sV, &mapV, kv_full, DV, k0, h, b);
      for (int qt = qt0, it = 0; qt < nqt; qt++, it++) {
        mbar_wait(tile_done, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * Cfg::kK + 2 * Cfg::kV);
        load_kmajor(sQk, &mapQk, q_full, DK, qt * 128, h, b);
        load_mnmajor(sQm, &mapQm, q_full, DK, qt * 128, h, b);
        load_kmajor(sOk, &mapOk, q_full, DV, qt * 128, h, b);
        load_mnmajor(sOm, &mapOm, q_full, DV, qt * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(kv_full, 0);
      for (int qt = qt0, it = 0; qt < nqt; qt++, it++) {
        mbar_wait(q_full, it & 1);
        tc_fence_after();
        score_mma(T_S, sK, sQk, DK);        // S^T  [keys x queries]
        score_mma(T_DP, sV, sOk, DV);       // dP^T [keys x queries]
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        grad_mma(T_DV, T_S, sOm, DV, it != 0);     // dV += Pd^T dO
        grad_mma(T_DK, T_DP, sQm, DK, it != 0);    // dK += dS^T Q
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
      // lse / delta of this query tile (the previous tile's readers are past their p_ready arrive)
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
          s[j] = tf32_rn(pr);
          dp[j] = tf32_rn(ds);
        }
        tmem_st32(T_S + lane_addr + c * 32, s);
        tmem_st32(T_DP + lane_addr + c * 32, dp);
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // epilogue: dK, dV rows of this key block
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool any = qt0 < nqt;
    const bool ok = key < p.Tk;
    float* dkrow = dk_out + b * p.k_bs + h * p.k_hs + (long long)key * p.k_rs;
    float* dvrow = dv_out + b * p.v_bs + h * p.v_hs + (long long)key * p.v_rs;
#pragma unroll 1
    for (int c = 0; c < DV / 32; c++) {
      float v[32];
      if (any) tmem_ld32(T_DV + lane_addr + c * 32, v);
      if (!ok) continue;
#pragma unroll
      for (int j4 = 0; j4 < 8; j4++)
        *reinterpret_cast<float4*>(dvrow + c * 32 + j4 * 4) =
            any ? make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int c = 0; c < DK / 32; c++) {
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

// ---------------------------------------------------------------------------------------------------------------------
template <int DK, int DV>
__global__ void __launch_bounds__(AB_THREADS, 1)
tc_sdpa_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapO,
                      const __grid_constant__ CUtensorMap mapKk, const __grid_constant__ CUtensorMap mapKm,
                      const __grid_constant__ CUtensorMap mapV, const AttnP p, const float* __restrict__ delta,
                      float* __restrict__ dq_out) {
  using Cfg = BwdCfg<DK, DV>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base, sO = smem_base + Cfg::kB_offO, sKk = smem_base + Cfg::kB_offKk, sKm = smem_base + Cfg::kB_offKm,
                 sVk = smem_base + Cfg::kB_offVk;
  uint8_t* pad_smem = gen_base + Cfg::kB_offPad;
  const uint32_t bar_base = smem_base + Cfg::kB_offBar;
  const uint32_t q_full = bar_base, k_full = bar_base + 8, s_full = bar_base + 16, p_ready = bar_base + 24,
                 tile_done = bar_base + 32, acc_full = bar_base + 40, tmem_slot = bar_base + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb_all = (p.Tk + 127) / 128;
  const int nkb = p.causal ? min(nkb_all, (q0 + 127) / 128 + 1) : nkb_all;     // key blocks beyond the last query see nothing

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
      mbar_expect_tx(q_full, Cfg::kK + Cfg::kV);
      load_kmajor(sQ, &mapQ, q_full, DK, q0, h, b);
      load_kmajor(sO, &mapO, q_full, DV, q0, h, b);
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(tile_done, (kb & 1) ^ 1);
        mbar_expect_tx(k_full, 2 * Cfg::kK + Cfg::kV);
        load_kmajor(sKk, &mapKk, k_full, DK, kb * 128, h, b);
        load_mnmajor(sKm, &mapKm, k_full, DK, kb * 128, h, b);
        load_kmajor(sVk, &mapV, k_full, DV, kb * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(k_full, kb & 1);
        tc_fence_after();
        score_mma(T_S, sQ, sKk, DK);        // S  [queries x keys]
        score_mma(T_DP, sO, sVk, DV);       // dP [queries x keys]
        umma_commit(s_full);
        mbar_wait(p_ready, kb & 1);
        tc_fence_after();
        grad_mma(T_DQ, T_DP, sKm, DK, kb != 0);    // dQ += dS K
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
          dp[j] = tf32_rn(ds);
        }
        tmem_st32(T_DP + lane_addr + c * 32, dp);
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* dqrow = dq_out + b * p.q_bs + h * p.q_hs + (long long)q * p.q_rs;
#pragma unroll 1
    for (int c = 0; c < DK / 32; c++) {
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
template <int DV>
__global__ void sdpa_delta_kernel(AttnP p, const float* __restrict__ dout, float* __restrict__ delta) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.H * p.Tq;
  if (w >= total) return;
  const int q = w % p.Tq, bh = w / p.Tq, h = bh % p.H, b = bh / p.H;
  const float* og = p.o + b * p.o_bs + h * p.o_hs + (long long)q * p.o_rs;
  const float* dg = dout + b * p.o_bs + h * p.o_hs + (long long)q * p.o_rs;
  float s = 0.f;
  for (int c = lane; c < DV; c += 32) s += og[c] * dg[c];
  s = warp_sum(s);
  if (lane == 0) delta[w] = s;
}

static int bhtd_map(CUtensorMap* m, const float* base, int d, int T, int H, int B, long long rs, long long hs, long long bs, bool mn) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)T, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)rs, (uint64_t)hs, (uint64_t)bs};
  uint32_t box[4] = {32, 128, 1, 1};
  return make_tensor_map_f32(m, base, 4, dims, strides, box, mn, true);
}

template <int DK, int DV>
static int launch_bwd(const AttnP& p, const float* dout, float* dq, float* dk, float* dv, float* delta, cudaStream_t st) {
  using Cfg = BwdCfg<DK, DV>;
  const int total = p.B * p.H * p.Tq;
  sdpa_delta_kernel<DV><<<ceil_div(total, 8), 256, 0, st>>>(p, dout, delta);
  int rc = check_launch("sdpa_delta");
  if (rc) return rc;
  CUtensorMap mQk, mQm, mKk, mKm, mV, mOk, mOm;
  if ((rc = bhtd_map(&mQk, p.q, DK, p.Tq, p.H, p.B, p.q_rs, p.q_hs, p.q_bs, false))) return rc;
  if ((rc = bhtd_map(&mQm, p.q, DK, p.Tq, p.H, p.B, p.q_rs, p.q_hs, p.q_bs, true))) return rc;
  if ((rc = bhtd_map(&mKk, p.k, DK, p.Tk, p.H, p.B, p.k_rs, p.k_hs, p.k_bs, false))) return rc;
  if ((rc = bhtd_map(&mKm, p.k, DK, p.Tk, p.H, p.B, p.k_rs, p.k_hs, p.k_bs, true))) return rc;
  if ((rc = bhtd_map(&mV, p.v, DV, p.Tk, p.H, p.B, p.v_rs, p.v_hs, p.v_bs, false))) return rc;
  if ((rc = bhtd_map(&mOk, dout, DV, p.Tq, p.H, p.B, p.o_rs, p.o_hs, p.o_bs, false))) return rc;
  if ((rc = bhtd_map(&mOm, dout, DV, p.Tq, p.H, p.B, p.o_rs, p.o_hs, p.o_bs, true))) return rc;
  auto* ka = tc_sdpa_bwd_dkdv_kernel<DK, DV>;
  auto* kb = tc_sdpa_bwd_dq_kernel<DK, DV>;
  static bool attr_a[kMaxDevices] = {}, attr_b[kMaxDevices] = {};
  if ((rc = ensure_dynamic_smem((const void*)ka, Cfg::kA_smem, attr_a, "tc_sdpa_bwd_dkdv"))) return rc;
  if ((rc = ensure_dynamic_smem((const void*)kb, Cfg::kB_smem, attr_b, "tc_sdpa_bwd_dq"))) return rc;
  ka<<<dim3(ceil_div(p.Tk, 128), p.H, p.B), AB_THREADS, Cfg::kA_smem, st>>>(mQk, mQm, mKk, mV, mOk, mOm, p, delta, dk, dv);
  rc = check_launch("tc_sdpa_bwd_dkdv");
  if (rc) return rc;
  kb<<<dim3(ceil_div(p.Tq, 128), p.H, p.B), AB_THREADS, Cfg::kB_smem, st>>>(mQk, mOk, mKk, mKm, mV, p, delta, dq);
  return check_launch("tc_sdpa_bwd_dq");
}

}  // namespace tc

int sdpa_bwd_tc(const AttnP& p, const float* dout, float* dq, float* dk, float* dv, float* delta, cudaStream_t st) {
  using namespace tc;
  B200_REQUIRE(p.H <= 65535 && p.B <= 65535, B200ASR_BAD_SHAPE, "sdpa_bwd (tcgen05): grid too large");
  if (p.dk == 64 && p.dv == 64) return launch_bwd<64, 64>(p, dout, dq, dk, dv, delta, st);
  if (p.dk == 32 && p.dv == 32) return launch_bwd<32, 32>(p, dout, dq, dk, dv, delta, st);
  if (p.dk == 64 && p.dv == 32) return launch_bwd<64, 32>(p, dout, dq, dk, dv, delta, st);
  if (p.dk == 32 && p.dv == 64) return launch_bwd<32, 64>(p, dout, dq, dk, dv, delta, st);
  set_error("sdpa_bwd (tcgen05): (dk=%d, dv=%d) unsupported, needs dk, dv in {32, 64}; use precision 0", p.dk, p.dv);
  return B200ASR_BAD_SHAPE;
}

}  // namespace b200asr
