// Persistent tcgen05 tile engine shared by the GEMM and the implicit-GEMM convolutions.
//
// One CTA per SM loops over output tiles (static round-robin).  Warp roles:
//   warp 0      TMA producer            -- runs ahead across tiles through a ring of smem stages
//   warp 1      MMA issuer + TMEM owner -- accumulates tile i into TMEM buffer (i & 1)
//   warps 2-5   epilogue                -- drain buffer (i & 1) with tcgen05.ld while tile i+1 is being multiplied
//   warps 6-13  3xTF32 operand split    -- A: smem -> (hi, lo) in TMEM; B: hi in place / lo in smem (only when NSPLIT == 3);
//               two groups of 4 warps take alternating k-blocks
// so per-tile prologue/epilogue latency is hidden behind the next tile's mainloop instead of being paid 200+ times per SM.
//
// A Policy supplies the problem-specific parts:
//   static constexpr int  BN, kABytes, kBBytes;  static constexpr bool kSplitA, kSplitB (operand needs the in-kernel split);
//   static constexpr bool kAMN, kBMN (operand majors);  struct Params;
//   static __device__ int  num_tiles(const Params&);          static __device__ int num_kb(const Params&, int tile);
//   struct Tile;  static __device__ Tile tile(p, tile);       // tile coordinates, decoded ONCE per tile (integer divisions)
//   static __device__ void load(p, Tile&, mapA, mapB, sa, sb, sb_lo, bar, leader);   // TMA for the next k-block of the tile
//                              // (fixed tx bytes, issued by the leader lane only); every lane advances the k cursor in Tile
//   static __device__ uint64_t a_desc(uint32_t saddr, int ks); static __device__ uint64_t b_desc(uint32_t saddr, int ks);
//   static __device__ void store(p, const Tile&, row, col0, const float (&v)[32]);     // 32 accumulator columns of one row
//   static constexpr bool kSumA, kSumB;   // fused reductions over the contraction index done by the split warps (3xTF32
//       only): kSumA = row sums of an MN-major A tile, kSumB = column sums of an MN-major B tile (bias gradients of the
//       weight-gradient GEMMs -- the operand is in shared memory anyway, a separate column-sum pass re-reads it from HBM)
//   static __device__ bool want_sums(p, Tile);  void sum_a_store(p, Tile, row, v);  void sum_b_store(p, Tile, col, float4 v);
// The producer is ONE thread: anything it does per k-block is on the critical path of the whole SM (the first version
// re-derived (b, t, f, tap) with integer divisions for every k-block -- ~1000 clk each -- and that, not the tensor pipe,
// set the pace; see profiles/engine_ConvPolicy_r1.md).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200asr {
namespace tc {

// NSPLIT selects the arithmetic: 1 = TF32 (SS-MMA), 3 = 3xTF32 (A split into TMEM, TS-MMA), and the kind::f16 modes
// 2 = bf16 (A converted to bf16 into TMEM, one TS-MMA) and 6 = bf16x3 (2-term bf16 split, three TS-MMAs at twice the tf32
// rate).  In the bf16 modes the fp32 A tile (K- or MN-major) is converted by the split warps on its way into TMEM, and B is
//   * either a PRE-CONVERTED K-major bf16 operand ([hi] or [hi | lo], 64-byte rows, SWIZZLE_64B; Policy::load16) -- weights,
//   * or (Policy::kSplitB, both operands MN-major: the weight-gradient GEMMs, whose operands are activations) the raw fp32
//     MN-major tile, which the split warps convert into bf16 MN-major tiles (64 columns = 128 bytes per k-line, SWIZZLE_128B)
//     in a second shared-memory region of the stage.
//   * or (Policy::kBMN without kSplitB) MN-major bf16 hi | lo tiles that the PRODUCER of the activation already wrote next to
//     the fp32 tensor ("pairs"): loaded by TMA like the weights, no conversion at all (convolution weight gradient).
// Policy::b_desc16 describes the bf16 B tile in every case.
constexpr bool eng_is_bf16(int nsplit) { return nsplit == 2 || nsplit == 6; }
constexpr bool eng_has_split_warps(int nsplit) { return nsplit != 1; }

constexpr int ENG_THREADS_X1 = 192;     // TMA, MMA, 4 epilogue warps
constexpr int ENG_THREADS_X3 = 448;     // + 8 split warps
constexpr int ENG_SPLIT_THREADS = 256;
#ifndef ENG_SPLIT_GROUPS
#define ENG_SPLIT_GROUPS 2              // the 8 split warps work as this many groups on alternating k-blocks
#endif

// 3xTF32 keeps the A operand in TENSOR MEMORY: the split warps read the raw fp32 tile from smem once and write
// hi / lo straight into TMEM (tcgen05.st), and the MMAs run in the TS form (A from TMEM, B from smem).  Shared memory then
// holds only [A_raw | B_hi | B_lo] per stage (48 KB instead of 64 KB for a 128x128 tile -> one more stage in flight) and the
// tensor core no longer competes with the split for smem bandwidth on the A side (an SS 128x128x8 MMA reads 8 KB of smem
// per 64 cycles = the full 128 B/clk of the SM).
// A policy may opt in (static constexpr bool kCat = true) to the "concatenated" form of the 3xTF32 products at BN = 64: the
// tensor core is markedly less efficient at N = 64 than at N = 128 (see tc_conv_halo.cu), so hi*hi and hi*lo become ONE N = 128
// MMA against the adjacent [B_hi ; B_lo] tiles of the stage into a 128-column accumulator, lo*hi an N = 64 MMA into its first
// half, and the epilogue adds the halves.  Costs accumulator columns: 4 stages of A in tensor memory instead of 6.
template <class P, class = void> struct policy_cat : std::false_type {};
template <class P> struct policy_cat<P, std::void_t<decltype(P::kCat)>> : std::bool_constant<P::kCat> {};

template <class Policy, int NSPLIT> struct EngineCfg {
  static constexpr bool kBf16 = eng_is_bf16(NSPLIT);
  static constexpr bool kCat = policy_cat<Policy>::value && Policy::BN == 64 && NSPLIT == 3;
  static constexpr int kAccW = kCat ? 128 : Policy::BN;                                      // accumulator columns per buffer
  static constexpr int kBHalves = (NSPLIT == 3 || NSPLIT == 6) ? 2 : 1;                     // B tiles per stage (hi | lo)
  static constexpr int kBTile = kBf16 ? Policy::BN * 64 : Policy::kBBytes;                   // bf16: 32 k x 2 B = 64-byte rows
  static constexpr int kACols = NSPLIT == 3 ? 64 : (NSPLIT == 6 ? 32 : (NSPLIT == 2 ? 16 : 0));   // TMEM columns of A per stage
  static constexpr int kBRaw = (kBf16 && Policy::kSplitB) ? Policy::kBBytes : 0;             // bf16 + in-kernel B conversion: raw fp32 tile
  static constexpr int kStageBytes = Policy::kABytes + kBRaw + kBHalves * kBTile;
  static constexpr int kMaxByTmem = NSPLIT == 1 ? 8 : (512 - 2 * kAccW) / kACols;
  static constexpr int kBySmem = (200 * 1024) / kStageBytes;
#ifndef ENG_MAX_STAGES
#define ENG_MAX_STAGES 8
#endif
  static constexpr int kStagesRaw = kBySmem < kMaxByTmem ? kBySmem : kMaxByTmem;
  static constexpr int kStagesCap = kStagesRaw > ENG_MAX_STAGES ? ENG_MAX_STAGES : kStagesRaw;
  static constexpr int kStages = NSPLIT == 1 ? kStagesCap : kStagesCap - kStagesCap % ENG_SPLIT_GROUPS;
  static constexpr int kOffBraw = Policy::kABytes;
  static constexpr int kOffBhi = Policy::kABytes + kBRaw;
  static constexpr int kOffBlo = kOffBhi + kBTile;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kSmemBytes = kBarOff + 512 + 1024;
  static constexpr int kAccCols = 2 * kAccW;
  static constexpr int kTmemCols = NSPLIT != 1 ? 512 : (kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : 256)));
  static constexpr int kTxBytes = kBf16 ? Policy::kABytes + (Policy::kSplitB ? Policy::kBBytes : kBHalves * kBTile)
                                        : Policy::kABytes + Policy::kBBytes +
                                          (NSPLIT == 3 && !Policy::kSplitB ? Policy::kBBytes : 0);   // pre-split B arrives as hi+lo
  static_assert(!kBf16 || !Policy::kSplitB || (Policy::kAMN && Policy::kBMN),
                "bf16 modes: a pre-converted bf16 B (K-major weights, or MN-major activation pairs), or both operands MN-major fp32 "
                "with B converted in the kernel");
  static_assert(kStages >= 2, "stage too large");
  // Each split group must see EVERY phase of the full barriers it waits on: mbarrier parity waits only distinguish the
  // current phase from the one before, so a group that skipped a phase of a stage could take a stale completion for the
  // one it is waiting for.  With the stage count a multiple of the group count a stage always belongs to one group.
  static_assert(NSPLIT == 1 || kStages % ENG_SPLIT_GROUPS == 0, "stage count must be a multiple of the split-group count");
};

template <class Policy, int NSPLIT>
__global__ void __launch_bounds__(eng_has_split_warps(NSPLIT) ? ENG_THREADS_X3 : ENG_THREADS_X1, 1)
tc_engine_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const typename Policy::Params p) {
  using Cfg = EngineCfg<Policy, NSPLIT>;
  constexpr int S = Cfg::kStages;
  constexpr int BN = Policy::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto xfm_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (3 * S + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (3 * S + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = Policy::num_tiles(p);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(xfm_bar(s), ENG_SPLIT_THREADS / ENG_SPLIT_GROUPS); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; b++) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  griddep_launch();      // the next kernel in the stream may start its own prologue as our CTAs retire ...
  griddep_wait();        // ... and ours ran under the predecessor's tail: from here on its results are visible (common.cuh)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    const bool leader = elect_one();
    uint32_t kbg = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nkb = Policy::num_kb(p, tile);
      typename Policy::Tile tc = Policy::tile(p, tile);
      for (int kb = 0; kb < nkb; kb++, kbg++) {
        const int s = kbg % S;
        mbar_wait(empty_bar(s), ((kbg / S) & 1) ^ 1);
        const uint32_t sa = smem_base + s * Cfg::kStageBytes;
        if (leader) mbar_expect_tx(full_bar(s), Cfg::kTxBytes);
        if constexpr (Cfg::kBf16 && !Policy::kSplitB) Policy::load16(p, tc, &mapA, &mapB, sa, sa + Cfg::kOffBhi, sa + Cfg::kOffBlo, full_bar(s), leader, Cfg::kBHalves);
        else if constexpr (Cfg::kBf16) Policy::load(p, tc, &mapA, &mapB, sa, sa + Cfg::kOffBraw, 0u, full_bar(s), leader);     // raw fp32 A and B tiles
        else Policy::load(p, tc, &mapA, &mapB, sa, sa + Cfg::kOffBhi, sa + Cfg::kOffBlo, full_bar(s), leader);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    // The issuing thread's own instruction stream is on the critical path: the tensor pipe's queue holds only a few MMAs,
    // so whatever it executes between the last MMA of one k-block and the first of the next (barrier poll, descriptor
    // arithmetic) must be shorter than the time that queue takes to drain -- ~130 clk with N = 64.  Descriptors are
    // therefore built ONCE per CTA and advanced by adding to their 14-bit address field (shared memory is < 256 KB, so
    // the field never carries into the stride fields above it); stage index and phase are running counters.
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_tf32(128, BN, Policy::kAMN, Policy::kBMN);
    constexpr uint32_t idesc_ts = make_idesc_tf32(128, BN, false, Policy::kBMN);
    constexpr uint32_t idesc_cat = make_idesc_tf32(128, 128, false, Policy::kBMN);
    constexpr bool kNeedXfm = NSPLIT != 1;
    constexpr uint32_t idesc_bf = make_idesc_bf16(128, BN, false, Policy::kBMN);
    constexpr uint32_t kStageStep = Cfg::kStageBytes >> 4, kLoStep = Cfg::kBTile >> 4;
    uint64_t bd0, ad0;
    uint32_t b_ks, a_ks;
    if constexpr (Cfg::kBf16) {
      bd0 = Policy::b_desc16(smem_base + Cfg::kOffBhi, 0); ad0 = 0;
      b_ks = (uint32_t)(Policy::b_desc16(0, 1) - Policy::b_desc16(0, 0)); a_ks = 0;   // address-field step per 16-deep k-step
    } else {
      bd0 = Policy::b_desc(smem_base + Cfg::kOffBhi, 0); ad0 = Policy::a_desc(smem_base, 0);
      b_ks = (uint32_t)(Policy::b_desc(0, 1) - Policy::b_desc(0, 0));     // address-field step per 8-deep k-step
      a_ks = (uint32_t)(Policy::a_desc(0, 1) - Policy::a_desc(0, 0));
    }
    const uint32_t bd_hi = (uint32_t)(bd0 >> 32), ad_hi = (uint32_t)(ad0 >> 32);
    auto mk = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    uint32_t s = 0, ph = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
      const int nkb = Policy::num_kb(p, tile);
      const uint32_t buf = tcount & 1;
      mbar_wait(tempty_bar(buf), ((tcount >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * Cfg::kAccW;
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(kNeedXfm ? xfm_bar(s) : full_bar(s), ph);
        tc_fence_after();
        if (leader) {
          const uint32_t b_lo32 = (uint32_t)bd0 + s * kStageStep, a_lo32 = (uint32_t)ad0 + s * kStageStep;
          const uint32_t a_t = tmem_base + Cfg::kAccCols + s * Cfg::kACols;
#pragma unroll
          for (int ks = 0; ks < (Cfg::kBf16 ? 2 : 4); ks++) {
            const uint64_t b_hi = mk(bd_hi, b_lo32 + ks * b_ks);
            const uint32_t acc0 = (kb | ks) != 0 ? 1u : 0u;
            if constexpr (NSPLIT == 1) {
              umma_tf32(d_tmem, mk(ad_hi, a_lo32 + ks * a_ks), b_hi, idesc, acc0);
            } else if constexpr (NSPLIT == 2) {
              umma_bf16_ts(d_tmem, a_t + ks * 8, b_hi, idesc_bf, acc0);          // 16 k = 8 packed columns per step
            } else if constexpr (NSPLIT == 6) {
              const uint64_t b_lo = mk(bd_hi, b_lo32 + kLoStep + ks * b_ks);
              const uint32_t a_hi = a_t + ks * 8, a_lo = a_hi + 16;
              umma_bf16_ts(d_tmem, a_lo, b_hi, idesc_bf, acc0);
              umma_bf16_ts(d_tmem, a_hi, b_lo, idesc_bf, 1u);
              umma_bf16_ts(d_tmem, a_hi, b_hi, idesc_bf, 1u);
            } else {
              const uint64_t b_lo = mk(bd_hi, b_lo32 + kLoStep + ks * b_ks);
              const uint32_t a_hi = a_t + ks * 8, a_lo = a_hi + 32;
              if constexpr (Cfg::kCat) {
                umma_tf32_ts(d_tmem, a_hi, b_hi, idesc_cat, acc0);             // [hi*hi | hi*lo]: B_lo follows B_hi in the stage
                umma_tf32_ts(d_tmem, a_lo, b_hi, idesc_ts, 1u);                // + lo*hi into the first 64 columns
              } else {
                umma_tf32_ts(d_tmem, a_lo, b_hi, idesc_ts, acc0);
                umma_tf32_ts(d_tmem, a_hi, b_lo, idesc_ts, 1u);
                umma_tf32_ts(d_tmem, a_hi, b_hi, idesc_ts, 1u);
              }
            }
          }
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (leader) umma_commit(tfull_bar(buf));
      __syncwarp();
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (warps 2-5: TMEM lane quarters 2,3,0,1)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
      const uint32_t buf = tcount & 1;
      const int nkb = Policy::num_kb(p, tile);
      const typename Policy::Tile tc = Policy::tile(p, tile);
      mbar_wait(tfull_bar(buf), (tcount >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; c++) {
        float v[32];
        if (nkb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * Cfg::kAccW + (uint32_t)(c * 32), v);
          if constexpr (Cfg::kCat) {
            float w[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * Cfg::kAccW + (uint32_t)(64 + c * 32), w);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] += w[j];
          }
        } else
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0.f;
        if (c == BN / 32 - 1) {            // all TMEM reads of this buffer are done: hand it back before the global stores
          tc_fence_before();
          mbar_arrive(tempty_bar(buf));
        }
        Policy::store(p, tc, row, c * 32, v);
      }
    }
  } else {
    // ------------------------------------------------------------------ operand split warps (3xTF32 only)
    if constexpr (Cfg::kBf16) {
      // bf16 modes: the thread that owns A-tile row r reads its 32 fp32 k-values (K-major: one 128B-swizzled row; MN-major:
      // one element per k-line), converts them to bf16 (NSPLIT == 6: hi and lo of the 2-term split) and stores them PACKED,
      // two k per 32-bit column, into TMEM.  With Policy::kSplitB the group also converts the raw fp32 MN-major B tile into
      // the bf16 MN-major tile(s) the tensor core reads.
      constexpr int G = ENG_SPLIT_GROUPS, WPG = 8 / G, NT = ENG_SPLIT_THREADS / G;
      static_assert(WPG == 4, "bf16 split: one thread per A row and group");
      const int group = (warp - 6) / WPG;
      const int t = threadIdx.x - 192 - group * NT;
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      uint32_t kbg = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nkb = Policy::num_kb(p, tile);
        constexpr bool kSums = Policy::kSumA || Policy::kSumB;
        typename Policy::Tile stc;
        bool want = false;
        if constexpr (kSums) { stc = Policy::tile(p, tile); want = Policy::want_sums(p, stc); }
        float asum = 0.f;
        float4 bsum[BN / 32];
#pragma unroll
        for (int c = 0; c < BN / 32; c++) bsum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = 0; kb < nkb; kb++, kbg++) {
          if ((int)(kbg % G) != group) continue;
          const int s = kbg % S;
          mbar_wait(full_bar(s), (kbg / S) & 1);
          uint8_t* stage = gen_base + s * Cfg::kStageBytes;
          const uint8_t* araw = stage;
          uint32_t hi[16], lo[16];
          if constexpr (!Policy::kAMN) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float4 x = *reinterpret_cast<const float4*>(araw + row * 128 + ((j ^ (row & 7)) << 4));
              if constexpr (NSPLIT == 6) {
                split_bf16_pair(x.x, x.y, hi[2 * j], lo[2 * j]);
                split_bf16_pair(x.z, x.w, hi[2 * j + 1], lo[2 * j + 1]);
              } else {
                hi[2 * j] = pack_bf16_pair(x.x, x.y);
                hi[2 * j + 1] = pack_bf16_pair(x.z, x.w);
              }
            }
          } else {
            // MN-major tile: chunk (row/32) of 32 k-lines x 128 B; 32-byte atom a = (row%32)/8 stored at atom (a ^ (k & 3))
            const uint8_t* cbase = araw + (row >> 5) * 4096 + (row & 7) * 4;
            const int atom = (row & 31) >> 3;
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float x0 = *reinterpret_cast<const float*>(cbase + (2 * j) * 128 + ((atom ^ ((2 * j) & 3)) << 5));
              const float x1 = *reinterpret_cast<const float*>(cbase + (2 * j + 1) * 128 + ((atom ^ ((2 * j + 1) & 3)) << 5));
              if constexpr (NSPLIT == 6) split_bf16_pair(x0, x1, hi[j], lo[j]);
              else hi[j] = pack_bf16_pair(x0, x1);
              if constexpr (Policy::kSumA) asum += x0 + x1;
            }
          }
          const uint32_t acol = tmem_base + ((uint32_t)(quarter * 32) << 16) + Cfg::kAccCols + s * Cfg::kACols;
          tmem_st16u(acol, hi);
          if constexpr (NSPLIT == 6) tmem_st16u(acol + 16, lo);
          if constexpr (Policy::kSumB && !Policy::kSplitB) {
            // column sums of a bf16 MN-major B tile that arrived by TMA (hi + lo = the 16-bit value the MMAs use): thread t
            // owns column t -- 64-column chunks of 4 KB, k-line k at k * 128, 16-byte unit u stored at u ^ (k & 7)
            if (want && t < BN) {
              const uint8_t* bh = stage + Cfg::kOffBhi + ((t >> 6) << 12) + ((t & 7) << 1);
              const int unit = (t & 63) >> 3;
              float sum = 0.f;
#pragma unroll 8
              for (int k = 0; k < 32; k++) {
                const uint8_t* q = bh + (k << 7) + ((unit ^ (k & 7)) << 4);
                sum += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(q)) << 16);
                if constexpr (NSPLIT == 6) sum += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(q + Cfg::kBTile)) << 16);
              }
              bsum[0].x += sum;
            }
          }
          if constexpr (Policy::kSplitB) {
            // raw tile: float4 i sits in 4 KB chunk i / 256 (32 columns), k-line (i % 256) / 8, 16-byte slot i % 8, and the
            // logical 32-byte atom is (slot / 2) ^ (k & 3) (SWIZZLE_128B_ATOM_32B).  bf16 tile: 64-column chunks of 4 KB, k-line
            // k at k * 128, 16-byte unit (8 columns) u stored at unit u ^ (k & 7) (SWIZZLE_128B).
            const float4* braw = reinterpret_cast<const float4*>(stage + Cfg::kOffBraw);
            uint8_t* bhi = stage + Cfg::kOffBhi;
            uint8_t* blo = stage + Cfg::kOffBlo;
            constexpr int NF4 = Policy::kBBytes / 16;
#pragma unroll
            for (int j = 0; j < NF4 / NT; j++) {
              const int idx = t + j * NT;
              const float4 x = braw[idx];
              const int k = (idx & 255) >> 3, slot = idx & 7;
              const int n = ((idx >> 8) << 5) + ((((slot >> 1) ^ (k & 3))) << 3) + ((slot & 1) << 2);
              const int off = ((n >> 6) << 12) + (k << 7) + (((((n & 63) >> 3) ^ (k & 7))) << 4) + ((n & 7) << 1);
              uint2 h, l;
              if constexpr (NSPLIT == 6) {
                split_bf16_pair(x.x, x.y, h.x, l.x);
                split_bf16_pair(x.z, x.w, h.y, l.y);
                *reinterpret_cast<uint2*>(blo + off) = l;
              } else {
                h.x = pack_bf16_pair(x.x, x.y);
                h.y = pack_bf16_pair(x.z, x.w);
              }
              *reinterpret_cast<uint2*>(bhi + off) = h;
              if constexpr (Policy::kSumB) {
                float4& acc = bsum[(j * NT) / 256];
                acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
              }
            }
            fence_proxy_async_smem();
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(xfm_bar(s));
        }
        if constexpr (kSums) {
          if (want) {
            if constexpr (Policy::kSumA) Policy::sum_a_store(p, stc, row, asum);
            if constexpr (Policy::kSumB && Policy::kSplitB) {
              const int slot = t & 7, kph = (t >> 3) & 3;
              const int col_in_chunk = (((slot >> 1) ^ kph) << 3) + ((slot & 1) << 2);
#pragma unroll
              for (int c = 0; c < BN / 32; c++) Policy::sum_b_store(p, stc, c * 32 + col_in_chunk, bsum[c]);
            }
            if constexpr (Policy::kSumB && !Policy::kSplitB) {
              if (t < BN) Policy::sum_b_store1(p, stc, t, bsum[0].x);
            }
          }
        }
      }
    } else if (NSPLIT == 3) {
      // The smem -> registers -> TMEM chain of one k-block is a serial latency (LDS, tcgen05.st, wait::st, arrive), so the
      // warps are divided into groups that take alternating k-blocks: each thread handles a wider slice less often and the
      // chains of consecutive k-blocks overlap.
      constexpr int G = ENG_SPLIT_GROUPS, WPG = 8 / G, KPT = 128 / WPG;   // warps per group, k columns per thread
      const int group = (warp - 6) / WPG;
      const int kpart = ((warp - 6) % WPG) >> 2;    // which KPT-wide slice of the 32-deep block
      const int t = threadIdx.x - 192 - group * WPG * 32;
      const int quarter = warp & 3;                 // TMEM lanes this warp may touch
      const int row = quarter * 32 + lane;          // A-tile row owned by this thread
      uint32_t kbg = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nkb = Policy::num_kb(p, tile);
        constexpr bool kSums = Policy::kSumA || Policy::kSumB;
        static_assert(!Policy::kSumA || Policy::kAMN, "row sums need an MN-major A tile (thread = row)");
        static_assert(!Policy::kSumB || (Policy::kBMN && Policy::kSplitB), "column sums need an MN-major B tile split in the kernel");
        typename Policy::Tile stc;
        bool want = false;
        if constexpr (kSums) { stc = Policy::tile(p, tile); want = Policy::want_sums(p, stc); }
        float asum = 0.f;
        float4 bsum[BN / 32];
#pragma unroll
        for (int c = 0; c < BN / 32; c++) bsum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = 0; kb < nkb; kb++, kbg++) {
          if ((int)(kbg % G) != group) continue;
          const int s = kbg % S;
          mbar_wait(full_bar(s), (kbg / S) & 1);
          const uint8_t* araw = gen_base + s * Cfg::kStageBytes;
          float hi[KPT], lo[KPT];
          if (!Policy::kAMN) {
            // K-major tile: row = 128 B, 16-byte chunk j stored at chunk (j ^ (row & 7)) (128B swizzle)
#pragma unroll
            for (int j = 0; j < KPT / 4; j++) {
              const int chunk = (kpart * (KPT / 4) + j) ^ (row & 7);
              const float4 x = *reinterpret_cast<const float4*>(araw + row * 128 + chunk * 16);
              hi[j * 4 + 0] = tf32_rn(x.x); lo[j * 4 + 0] = x.x - hi[j * 4 + 0];
              hi[j * 4 + 1] = tf32_rn(x.y); lo[j * 4 + 1] = x.y - hi[j * 4 + 1];
              hi[j * 4 + 2] = tf32_rn(x.z); lo[j * 4 + 2] = x.z - hi[j * 4 + 2];
              hi[j * 4 + 3] = tf32_rn(x.w); lo[j * 4 + 3] = x.w - hi[j * 4 + 3];
            }
          } else {
            // MN-major tile: chunk (row/32) of 32 k-lines x 128 B; 32-byte atom a = (row%32)/8 stored at atom (a ^ (k & 3))
            const uint8_t* cbase = araw + (row >> 5) * 4096 + (row & 7) * 4;
            const int atom = (row & 31) >> 3;
#pragma unroll
            for (int j = 0; j < KPT; j++) {
              const int k = kpart * KPT + j;
              const float x = *reinterpret_cast<const float*>(cbase + k * 128 + ((atom ^ (k & 3)) << 5));
              hi[j] = tf32_rn(x); lo[j] = x - hi[j];
              if constexpr (Policy::kSumA) asum += x;
            }
          }
          const uint32_t acol = tmem_base + ((uint32_t)(quarter * 32) << 16) + Cfg::kAccCols + s * 64 + kpart * KPT;
#pragma unroll
          for (int i = 0; i < KPT / 16; i++) {
            tmem_st16(acol + i * 16, hi + i * 16);
            tmem_st16(acol + 32 + i * 16, lo + i * 16);
          }
          if (Policy::kSplitB) {
            float4* stage = reinterpret_cast<float4*>(gen_base + s * Cfg::kStageBytes);
            if constexpr (Policy::kSumB) {
              // same elementwise split, plus per-thread partial column sums: float4 i of the tile sits in 4 KB chunk i / 256
              // (32 columns), k-line (i % 256) / 8, 16-byte slot i % 8; with NT threads striding by NT (a divisor of 256)
              // a thread keeps its slot and its k & 3, hence its 4 columns within every chunk (see sum_b_col below)
              constexpr int NT = ENG_SPLIT_THREADS / G, NF4 = Policy::kBBytes / 16;
              float4* bh = stage + Cfg::kOffBhi / 16;
              float4* bl = stage + Cfg::kOffBlo / 16;
#pragma unroll
              for (int j = 0; j < NF4 / NT; j++) {
                const int idx = t + j * NT;
                const float4 x = bh[idx];
                float4 h, l;
                h.x = tf32_rn(x.x); l.x = x.x - h.x;
                h.y = tf32_rn(x.y); l.y = x.y - h.y;
                h.z = tf32_rn(x.z); l.z = x.z - h.z;
                h.w = tf32_rn(x.w); l.w = x.w - h.w;
                bh[idx] = h;
                bl[idx] = l;
                float4& acc = bsum[(j * NT) / 256];
                acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
              }
            } else {
              split_tf32_inplace(stage + Cfg::kOffBhi / 16, stage + Cfg::kOffBlo / 16, Policy::kBBytes / 16, t, ENG_SPLIT_THREADS / G);
            }
            fence_proxy_async_smem();
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(xfm_bar(s));
        }
        if constexpr (kSums) {
          if (want) {
            if constexpr (Policy::kSumA) Policy::sum_a_store(p, stc, row, asum);
            if constexpr (Policy::kSumB) {
              // logical 32-byte atom = physical atom ^ (k & 3) (SWIZZLE_128B_ATOM_32B); k & 3 = (t / 8) & 3 for every j
              const int slot = t & 7, kph = (t >> 3) & 3;
              const int col_in_chunk = (((slot >> 1) ^ kph) << 3) + ((slot & 1) << 2);
#pragma unroll
              for (int c = 0; c < BN / 32; c++) Policy::sum_b_store(p, stc, c * 32 + col_in_chunk, bsum[c]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <class Policy, int NSPLIT>
int launch_engine(const CUtensorMap& ma, const CUtensorMap& mb, const typename Policy::Params& p, int ntiles, cudaStream_t st,
                  const char* what) {
  using Cfg = EngineCfg<Policy, NSPLIT>;
  auto* kern = tc_engine_kernel<Policy, NSPLIT>;
  static bool attr_set[kMaxDevices] = {};
  if (int rc = ensure_dynamic_smem((const void*)kern, Cfg::kSmemBytes, attr_set, what)) return rc;
  if (ntiles <= 0) return B200ASR_OK;
  const int grid = min(ntiles, device_sm_count());
  launch_pdl(kern, dim3(grid), dim3(eng_has_split_warps(NSPLIT) ? ENG_THREADS_X3 : ENG_THREADS_X1), Cfg::kSmemBytes, st, ma, mb, p);
  return check_launch(what);
}

}  // namespace tc
}  // namespace b200asr
