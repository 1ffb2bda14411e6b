// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 / TMEM), fed by TMA.
//
//   fprop : y[m, co]  = sum_{tap, ci} x[m (+) tap, ci] * w[tap][co][ci]  (+ bias, + residual, + BN statistics)
//   dgrad : dx[m, ci] = sum_{tap, co} dy[m (-) tap, co] * w[tap][ci][co] (+ addend)
//
// Both are the same kernel: a [128 pixels] x [BN channels] output tile per CTA, K = taps * Cin walked in
// 64-channel slices.  The A operand (activations, NHWC bf16) is fetched with a 4-D TMA box
// {64 ch, bw, bh, bn} whose (w, h) origin is shifted by the filter tap: TMA's out-of-bounds zero fill IS the
// convolution padding, so im2col never exists in memory.  The B operand (weights, [tap][Cout][Cin] bf16)
// is a 3-D TMA box {64, BN, 1}.  Both land in shared memory in the 128-byte-swizzled K-major layout that
// tcgen05.mma reads through a shared-memory descriptor; accumulators live in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> +bias/+residual -> bf16 -> swizzled smem -> TMA store; optional
// per-channel sum / sum-of-squares for the BatchNorm that consumes this tensor).
//
// Replaces the cuDNN convolutions behind nn.Conv2d in ResidualBlock/lin/creatModel
// (reference try_with_torch.py:186-193,199-207,248,253,271-273,291-297).
#include "hg_common.cuh"

// -DHG_DBG_TS=1 compiles per-phase clock64() stamps of CTA 0 into conv_gemm_kernel (hg_set_option("dbg_ts", 1|2));
// even an untaken `if (p.ts)` per phase costs ~8 % on the 64x64 kernels, so the production build has none.
#ifndef HG_DBG_TS
#define HG_DBG_TS 0
#endif
#define HG_TS (HG_DBG_TS && p.ts)
#ifndef HG_EPI_ONEWAVE
#define HG_EPI_ONEWAVE 256
#endif

namespace hg {

int g_short_1stage = 0;        // 1: multi-wave 1x1 kernels without residual: ONE stage (32 KB), 4 CTAs/SM
int g_long_k_3cta = 1;         // 1: multi-wave 3x3 kernels with 2 stages and 3 CTAs/SM instead of 3 stages and 2 CTAs/SM
int g_short_alias = 1;         // 1: multi-wave 1x1 kernels with a residual / raw BN input load it AFTER the (short) main loop
                               // into the aliased staging tile: 64 KB per CTA, 3 CTAs/SM instead of 2
int g_small_n_tiles = 0;       // 1: one-wave grids use 64-channel N tiles
int g_mid_n_tiles = 0;         // 1: grids of 1..2 waves of 128-wide tiles use 64-channel N tiles
int g_wgrad_big_n_panels = -1;  // larger maps: input-channel panels per CTA (< 0 = all (default), 0 = policy: 1x1 -> one panel,
                                // 3x3 -> all: isolated 1x1 launches -5 .. -30 %, step unchanged, 978 vs 980 images/s; > 0 = fixed)
int g_wgrad_small_n_panels = 0; // small maps (see g_wgrad_t1_max_kb): input-channel panels (of 64) per CTA; 0 = policy
                                // (1x1 or <= 32 K blocks: 1 panel, else 2), -1 = never split the input channels
int g_wgrad_t1_max_kb = 128;    // "small map": at most this many 64-pixel K blocks (16x16 at batch 32); 3x3: one tap per CTA     // 3x3 wgrad: one tap per CTA when the map has at most this many 64-pixel K blocks
int g_wgrad_bulk_reduce = 0;    // 1: wgrad epilogue through shared memory + cp.reduce.async.bulk (measured: no faster)
int g_wgrad_dbg = 0;            // HG_DBG_TS builds only: 1 = wgrad epilogue without the atomics, 2 = no epilogue at all
long long* g_dbg_ts = nullptr;  // debug: per-phase clock64 stamps of CTA 0 (hg_set_option dbg_ts)
int g_single_wave_deep = 1;   // 1: one-wave grids use the deep (6/4-stage, ~190 KB) pipelines
int g_wgrad_smem_kb = 196;    // shared-memory budget of the wgrad pipeline
int g_wgrad_kpx = 128;        // pixels per K block on the large maps (64 or 128; 128: 3x3 @32x32 29.1 -> 20.7 us, @64x64 71 -> 60 us)
int g_wgrad_halo = 1;          // 1: 3x3 weight gradients on the large maps take one filter column per CTA (one x box + halo rows)
int g_wgrad_halo_min_kb = 512; // ... from this many 128-pixel K blocks on (64x64 at batch 32: 57.2 -> 52.8 us; at 32x32 three
                               // accumulators per CTA and 43 instead of 16 K splits triple the atomics: 18.9 -> 24.1 us)
int g_onewave_cluster = 0;    // experiment: cluster size for one-wave grids of <= 32 CTAs (0 = plain launch)
int g_wgrad_fused_bias = 1;   // 1: the bias gradient is an extra all-ones N slab of the wgrad GEMM (no column-sum kernel)

// a = [relu](scale * x + shift) on 8 consecutive channels held in one 16-byte register quad
__device__ __forceinline__ uint4 bn_relu_chunk(uint4 u, const float (&sc)[8], const float (&sh)[8], bool relu) {
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __bfloat1622float2(h2[e]);
    f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
    f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
    if (relu) {
      f.x = fmaxf(f.x, 0.f);
      f.y = fmaxf(f.y, 0.f);
    }
    h2[e] = __floats2bfloat162_rn(f.x, f.y);
  }
  return u;
}

struct ConvGemmParams {
  int M_total;        // output pixels of this GEMM = N*H*W
  int H, W;           // spatial size of the GEMM's output grid (stride-2 data gradient: the half-resolution grid of one
                      // output parity class)
  int stride;         // A-operand origin = (h0, w0) * stride + tap offset (2: fprop of a stride-2 convolution)
  int Hin, Win;       // spatial size of the A-operand tensor (kFold: which tile rows are convolution padding)
  int ntaps;          // filter taps walked by this launch
  signed char tap_dh[12], tap_dw[12];  // A-operand offset of each tap (fprop: r*dil - pad; dgrad: pad - r*dil; ...)
  signed char tap_w[12];               // index of the tap's weight matrix inside the packed weight
  int kchunks;        // padded Cin / 64
  int n_total;        // padded Cout (row pitch of bias / stats)
  int c_real;         // real Cout (for the NCHW fp32 side output)
  const float* bias;  // [n_total] or null
  float* stats;       // [2*n_total] or null;  kMask: the BatchNorm-backward sums {sum g, sum g*xhat}
  float* out_nchw;    // optional fp32 NCHW copy of the first c_real channels (heatmap heads)
  int has_res;        // kPlain/kFold: residual added;  kMask: tmR is the raw BatchNorm input
  int n_tiles;        // padded Cout / BN
  int parity;         // 1: tmC / tmR are 5-D parity views {C, pw, W, ph, N*H} of a tensor of twice the resolution: the
  int par_h, par_w;   //    tile is stored to (read from) the pixels (2h + par_h, 2w + par_w)  (stride-2 data gradient)
  BnFoldDev fold;     // kFold: BatchNorm of the INPUT channels;  kMask: BatchNorm of the OUTPUT channels
  long long* ts;      // debug timestamps or null
};

// Shared-memory plan.  ALIAS = the epilogue staging buffers (output tile C; kMask: + the fp32 g*xhat tile Q) reuse the
// pipeline stages, which are idle once the last MMA has retired: a 3-stage 3x3 kernel then needs ~100 KB and TWO CTAs
// share an SM (one CTA's epilogue overlaps the other's main loop), a 2-stage 1x1 kernel ~70 KB (three CTAs).  A
// residual / raw-BatchNorm-input tile that must be in C before the epilogue is then loaded after the main loop;
// kernels with a residual and a short K keep a dedicated C buffer instead (ALIAS = false) and load it up front.
template <int BN, int STAGES, int MODE, bool ALIAS>
struct ConvGemmSmem {
  static constexpr int kABytes = 128 * 128;      // 128 pixels x 64 ch x 2 B
  static constexpr int kBBytes = BN * 128;       // BN out-channels x 64 ch x 2 B
  static constexpr int kCPanels = BN / 64;       // output staging: panels of 128 rows x 64 ch
  static constexpr int kCBytes = kCPanels * 128 * 128;
  static constexpr int kStageBytes = STAGES * (kABytes + kBBytes);
  static constexpr int kQBytes = MODE == kMask ? kCBytes : 0;   // kMask: masked gradient tile G (C holds the raw input)
  // ALIAS: C (and Q) inside the stages.  !ALIAS: dedicated C behind the stages (loaded up front); Q still aliases
  // the stages, which are idle by the time the epilogue writes it.
  static constexpr int kMainBytes =
      ALIAS ? (kStageBytes > kCBytes + kQBytes ? kStageBytes : kCBytes + kQBytes)
            : (kStageBytes > kQBytes ? kStageBytes : kQBytes) + kCBytes;
  static constexpr int kCOffset = ALIAS ? 0 : (kStageBytes > kQBytes ? kStageBytes : kQBytes);
  static constexpr int kQOffset = ALIAS ? kCBytes : 0;
  static constexpr int kBarBytes = 256;          // mbarriers + TMEM slot
  static constexpr int kBiasBytes = BN * 4;
  static constexpr int kCoefBytes = 2048;        // kFold: scale/shift[256];  kMask: scale/shift/A/B[BN]
  // + per-channel column sums of the epilogue: one {sum, sum of squares}[2*BN] slice per row group (kAccBytes below)
  static constexpr int kTotal = kMainBytes + kBarBytes + kBiasBytes + kCoefBytes + 1024 /*align slack*/;
};

// Epilogue warps.  One-wave grids (MINB == 1: one CTA per SM, pure latency): two warps per TMEM lane quarter, each takes
// half of the tile's columns -- a single warp per SM sub-partition has nothing to hide its tcgen05.ld / shared-memory
// latencies behind (3x3 @4x4 9.2 -> 8.8 us, 1x1 5.2 -> 5.0 us).  Multi-wave grids keep four warps: the co-resident
// CTAs already interleave, and the extra threads only cost (1x1 @64x64 24.6 -> 25.7 us with eight).
template <int MINB, int BN>
struct ConvThreads {
  static constexpr int kEpi = MINB == 1 ? (BN >= 128 ? HG_EPI_ONEWAVE : 256) : 128;
  static constexpr int kAll = 64 + kEpi;
  static constexpr int kAccBytes = (kEpi / (BN / 4)) * 2 * BN * 4;   // column-sum slices of the epilogue's row groups
};

template <int BN, int STAGES, int MINB, int MODE, bool ALIAS>
__global__ void __launch_bounds__(ConvThreads<MINB, BN>::kAll, MINB)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const ConvGemmParams p) {
  using L = ConvGemmSmem<BN, STAGES, MODE, ALIAS>;
  constexpr int kEpiThreads = ConvThreads<MINB, BN>::kEpi;
  static_assert(3 * STAGES + 2 <= 30, "barrier region too small");
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the shared array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * L::kABytes;
  uint8_t* sC = smem + L::kCOffset;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kMainBytes);
  uint64_t* full_bar = bars;                  // [STAGES] TMA landed
  uint64_t* empty_bar = bars + STAGES;        // [STAGES] MMAs done reading the slot
  uint64_t* tmem_full = bars + 2 * STAGES;    // accumulator ready
  uint64_t* res_full = bars + 2 * STAGES + 1;
  uint64_t* ready_bar = bars + 2 * STAGES + 2;  // [STAGES] kFold: A tile rewritten, MMA may read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + L::kBarBytes);  // [BN]
  float* coef_s = bias_s + BN;                                                                // 512 floats
  float* acc_s = coef_s + 512;                                       // [row groups][2 * BN] (ConvThreads::kAccBytes)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (HG_TS && blockIdx.x == 0 && threadIdx.x == 0) p.ts[0] = clock64();
  // 1-D grid, N tile fastest: the CTAs that share an activation tile run back to back (second read hits L2)
  const int m0 = (blockIdx.x / p.n_tiles) * 128;
  const int n_off = (blockIdx.x % p.n_tiles) * BN;
  const int num_kb = p.ntaps * p.kchunks;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&ready_bar[s], kEpiThreads);
    }
    mbar_init(tmem_full, 1);
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (HG_TS && blockIdx.x == 0 && threadIdx.x == 0) p.ts[1] = clock64();
  // tile origin (n0, h0, w0): three integer divisions that must not sit between the dependency wait and the first load
  const int hw_t = p.H * p.W;
  const int n0_t = m0 / hw_t;
  const int rem_t = m0 - n0_t * hw_t;
  const int h0_t = rem_t / p.W;
  const int w0_t = rem_t - h0_t * p.W;
  // PDL: everything above overlapped the previous kernel's tail; global memory is ours from here.  The producer warp
  // starts its TMA loads at once; the per-channel coefficients are fetched by the epilogue warps meanwhile.
  pdl_wait();
  if (HG_TS && blockIdx.x == 0 && threadIdx.x == 0) p.ts[2] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // warp-uniform loop, one elected lane issues (tensor-map and barrier operands stay in uniform registers)
    {
      const int n0 = n0_t, h0 = h0_t, w0 = w0_t;
      auto load_res = [&]() {
        mbar_expect_tx(res_full, L::kCBytes);
        for (int pnl = 0; pnl < L::kCPanels; ++pnl) {
          if (p.parity) tma_load_5d(sC + pnl * 16384, &tmR, res_full, n_off + pnl * 64, p.par_w, w0, p.par_h, n0 * p.H + h0);
          else tma_load_2d(sC + pnl * 16384, &tmR, res_full, n_off + pnl * 64, m0);
        }
      };
      if (!ALIAS && p.has_res && elect_one()) load_res();
      __syncwarp();
      if (HG_TS && blockIdx.x == 0 && lane == 0) p.ts[15] = clock64();
      const int ws = w0 * p.stride, hs = h0 * p.stride;
      int kb = 0, st = 0;
      uint32_t ph = 1;                      // empty-barrier parity: a fresh barrier passes a parity-1 wait
      for (int t = 0; t < p.ntaps; ++t) {
        const int dh = p.tap_dh[t], dw = p.tap_dw[t], wt = p.tap_w[t];
        for (int kc = 0; kc < p.kchunks; ++kc, ++kb) {
          mbar_wait(&empty_bar[st], ph);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[st], L::kABytes + L::kBBytes);
            tma_load_4d(sA + st * L::kABytes, &tmA, &full_bar[st], kc * 64, ws + dw, hs + dh, n0);
            tma_load_3d(sB + st * L::kBBytes, &tmB, &full_bar[st], kc * 64, n_off, wt);
            if (HG_TS && blockIdx.x == 0 && kb == 0) p.ts[3] = clock64();
            if (HG_TS && blockIdx.x == 0 && kb < 8) p.ts[16 + kb] = clock64();
          }
          __syncwarp();
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
      }
      if (ALIAS && p.has_res) {
        // C aliases the pipeline stages: the residual / raw BatchNorm input may only land once every MMA has read them
        mbar_wait(tmem_full, 0);
        if (elect_one()) load_res();
      }
    }
    __syncwarp();
    pdl_trigger();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ONE thread runs the loop: it is a chain of dependent scalar instructions whose length per K block is what the
    // tensor core idles between K blocks (four N = 128 MMAs need 256 cycles): slot / phase are counters, the
    // descriptors of a slot are one add away from those of slot 0.
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, 1024);
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if constexpr (MODE == kFold) mbar_wait(&ready_bar[st], ph);
        else mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        if (HG_TS && blockIdx.x == 0 && lane == 0 && kb == 0) p.ts[4] = clock64();
        if (HG_TS && blockIdx.x == 0 && lane == 0 && kb < 8) p.ts[24 + kb] = clock64();
        const uint64_t adesc = adesc0 + (uint64_t)(uint32_t)(st * (L::kABytes >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)(uint32_t)(st * (L::kBBytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 16 K-elements = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[st]);  // frees the smem slot once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(tmem_full);
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
    if (HG_TS && blockIdx.x == 0 && lane == 0) p.ts[5] = clock64();
    pdl_trigger();
  } else {
    // ===================== transform (kFold) + epilogue (warps 2..5) =====================
    const int sub = warp & 3;           // TMEM sub-partition this warp may read
    const int row = sub * 32 + lane;    // accumulator row == pixel within the tile
    const int m = m0 + row;
    const bool row_ok = m < p.M_total;
    {
      const int et = threadIdx.x - 64;
    for (int c = et; c < BN; c += kEpiThreads) bias_s[c] = p.bias ? p.bias[n_off + c] : 0.f;
    if constexpr (MODE == kPlain) {
      // pivots of the output statistics (sums of y - pivot): fetched now, under the main loop -- a global load in the
      // middle of the epilogue is a whole L2 round trip on the critical path of every one-wave kernel
      if (p.stats != nullptr)
        for (int c = et; c < BN; c += kEpiThreads) coef_s[c] = p.stats[2 * p.n_total + n_off + c];
    }
    if constexpr (MODE == kPlainBnOut) {   // y = scale * acc + (scale * bias + shift)
      for (int c = et; c < BN; c += kEpiThreads) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, n_off + c, mu, is, sc, sh);
        coef_s[c] = sc;
        bias_s[c] = fmaf(bias_s[c], sc, sh);
      }
    }
    if constexpr (MODE == kFold) {
      // scale / shift of every INPUT channel (<= 256)
      for (int c = et; c < p.kchunks * 64; c += kEpiThreads) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
        coef_s[c] = sc;
        coef_s[256 + c] = sh;
      }
    }
    if constexpr (MODE == kMask) {
      for (int c = et; c < BN; c += kEpiThreads) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, n_off + c, mu, is, sc, sh);
        coef_s[c] = sc;                 // ReLU mask: scale * y + shift > 0 (the forward's own expression)
        coef_s[BN + c] = sh;
        coef_s[2 * BN + c] = is;        // xhat = y * A + B  ->  sum g*xhat = A * sum(g*y) + B * sum(g)
        coef_s[3 * BN + c] = -mu * is;
      }
    }
    }
    const int et = threadIdx.x - 64;    // 0..kEpiThreads-1
    named_bar_sync(1, kEpiThreads);     // coefficients visible to the epilogue warps

    if constexpr (MODE == kFold) {
      // Rewrite every A tile in place: a = [relu](scale * x + shift).  Thread = one 16-byte channel chunk (8
      // channels: its 16 coefficients live in registers) x kR rows, so the shared-memory loads of a tile are
      // issued back to back.  Rows that fall into the convolution padding (TMA zero fill) or past the end of the
      // tensor must be zero AFTER the transform.
      const int jch = et & 7;
      constexpr int kRB = kEpiThreads / 8;       // row bases (32)
      constexpr int kR = 128 / kRB;              // rows per thread (4)
      const int rbase = et >> 3;                 // rows rbase + kRB * i
      const int swz = (jch ^ (rbase & 7)) << 4;  // (row & 7) == (rbase & 7) for every row of this thread
      const int hw = p.H * p.W;
      int hrow[kR], wrow[kR];
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        const int mm = m0 + rbase + kRB * i;
        const int rem = mm % hw;
        hrow[i] = mm < p.M_total ? rem / p.W : -0x100000;  // out-of-range rows never pass the bounds test
        wrow[i] = rem % p.W;
      }
      const bool relu = p.fold.relu != 0;
      int kb = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int dh = p.tap_dh[t], dw = p.tap_dw[t];
        uint32_t vmask = 0;
#pragma unroll
        for (int i = 0; i < kR; ++i)
          if ((unsigned)(hrow[i] * p.stride + dh) < (unsigned)p.Hin && (unsigned)(wrow[i] * p.stride + dw) < (unsigned)p.Win)
            vmask |= 1u << i;
        for (int kc = 0; kc < p.kchunks; ++kc, ++kb) {
          const int st = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          float sc[8], sh[8];
          load_coef8(coef_s + kc * 64 + jch * 8, sc);
          load_coef8(coef_s + 256 + kc * 64 + jch * 8, sh);
          uint8_t* base = sA + st * L::kABytes + rbase * 128 + swz;
          mbar_wait(&full_bar[st], ph);
          uint4 u[kR];
#pragma unroll
          for (int i = 0; i < kR; ++i) u[i] = *reinterpret_cast<const uint4*>(base + i * kRB * 128);
#pragma unroll
          for (int i = 0; i < kR; ++i)
            u[i] = (vmask >> i) & 1u ? bn_relu_chunk(u[i], sc, sh, relu) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int i = 0; i < kR; ++i) *reinterpret_cast<uint4*>(base + i * kRB * 128) = u[i];
          fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
          mbar_arrive(&ready_bar[st]);
        }
      }
    }

    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (HG_TS && blockIdx.x == 0 && et == 0) p.ts[6] = clock64();
    pdl_trigger();  // main loop done: the next kernel may start its prologue under our epilogue
    if (p.has_res) mbar_wait(res_full, 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    float* nchw_row = nullptr;
    int plane = 0;
    if (p.out_nchw != nullptr && row_ok) {
      plane = p.H * p.W;
      const int n = m / plane;
      nchw_row = p.out_nchw + (size_t)n * p.c_real * plane + (m - n * plane);
    }
    // kMask: second staging buffer (masked gradient G, bf16) inside the idle pipeline stages
    uint8_t* sQ = smem + L::kQOffset;
    constexpr int kColSplit = kEpiThreads / 128;          // warps per TMEM lane quarter (1 or 2)
    const int chalf = (warp - 2) >> 2;                    // which part of the tile's columns this warp stages
#pragma unroll 1
    for (int j = chalf * (BN / 32 / kColSplit); j < (chalf + 1) * (BN / 32 / kColSplit); ++j) {
      float v[32];
      tmem_ld32(taddr + j * 32, v);
      tmem_ld_wait();
      const int pnl = (j * 32) / 64;
      const int chunk0 = ((j * 32) % 64) / 8;
      uint8_t* rowp = sC + pnl * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int swz = ((chunk0 + q) ^ (row & 7)) << 4;
        uint4* cp = reinterpret_cast<uint4*>(rowp + swz);
        float o[8];
        if constexpr (MODE == kMask) {
          // raw BatchNorm input of these 8 channels (stays in C) -> ReLU mask; the masked gradient goes to G
          const uint4 u = *cp;
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
          float cS[8], cT[8];
          const int c0 = j * 32 + q * 8;
          load_coef8(coef_s + c0, cS);
          load_coef8(coef_s + BN + c0, cT);
          const bool relu = p.fold.relu != 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = f2fma(__bfloat1622float2(h[e]), make_float2(cS[2 * e], cS[2 * e + 1]),
                                   make_float2(cT[2 * e], cT[2 * e + 1]));
            const bool k0 = !relu || t.x > 0.f;
            const bool k1 = !relu || t.y > 0.f;
            o[2 * e] = k0 ? v[q * 8 + 2 * e] : 0.f;
            o[2 * e + 1] = k1 ? v[q * 8 + 2 * e + 1] : 0.f;
          }
          cp = reinterpret_cast<uint4*>(sQ + pnl * 16384 + row * 128 + swz);
        } else {
          if constexpr (MODE == kPlainBnOut) {
            const bool relu = p.fold.relu != 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = fmaf(v[q * 8 + e], coef_s[j * 32 + q * 8 + e], bias_s[j * 32 + q * 8 + e]);
              o[e] = relu ? fmaxf(t, 0.f) : t;
            }
          } else {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + j * 32 + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + j * 32 + q * 8 + 4);
            const float2 o0 = f2add(make_float2(v[q * 8 + 0], v[q * 8 + 1]), make_float2(b0.x, b0.y));
            const float2 o1 = f2add(make_float2(v[q * 8 + 2], v[q * 8 + 3]), make_float2(b0.z, b0.w));
            const float2 o2 = f2add(make_float2(v[q * 8 + 4], v[q * 8 + 5]), make_float2(b1.x, b1.y));
            const float2 o3 = f2add(make_float2(v[q * 8 + 6], v[q * 8 + 7]), make_float2(b1.z, b1.w));
            o[0] = o0.x; o[1] = o0.y; o[2] = o1.x; o[3] = o1.y; o[4] = o2.x; o[5] = o2.y; o[6] = o3.x; o[7] = o3.y;
          }
          if (p.has_res) {
            const uint4 u = *cp;
            const float2 r0 = f2add(make_float2(o[0], o[1]), bf2_to_f2(u.x));
            const float2 r1 = f2add(make_float2(o[2], o[3]), bf2_to_f2(u.y));
            const float2 r2 = f2add(make_float2(o[4], o[5]), bf2_to_f2(u.z));
            const float2 r3 = f2add(make_float2(o[6], o[7]), bf2_to_f2(u.w));
            o[0] = r0.x; o[1] = r0.y; o[2] = r1.x; o[3] = r1.y; o[4] = r2.x; o[5] = r2.y; o[6] = r3.x; o[7] = r3.y;
          }
          if (nchw_row != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = n_off + j * 32 + q * 8 + e;
              if (c < p.c_real) nchw_row[(size_t)c * plane] = o[e];
            }
          }
        }
        uint4 w;
        __nv_bfloat162* hw2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) hw2[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        *cp = w;
      }
    }
    if (HG_TS && blockIdx.x == 0 && et == 0) p.ts[7] = clock64();
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, kEpiThreads);
    if (et == 0) {
      const uint8_t* src = MODE == kMask ? sQ : sC;
      if (p.parity) {
        const int hw = p.H * p.W;
        const int n0 = m0 / hw;
        const int rem = m0 - n0 * hw;
        const int h0 = rem / p.W;
        const int w0 = rem - h0 * p.W;
        for (int pnl = 0; pnl < L::kCPanels; ++pnl)
          tma_store_5d(&tmC, src + pnl * 16384, n_off + pnl * 64, p.par_w, w0, p.par_h, n0 * p.H + h0);
      } else {
        for (int pnl = 0; pnl < L::kCPanels; ++pnl) tma_store_2d(&tmC, src + pnl * 16384, n_off + pnl * 64, m0);
      }
      tma_store_commit();
    }
    if (p.stats != nullptr) {
      // Per-channel column sums of the values just staged (bf16, exactly what the consumers will read):
      //   kPlain / kFold: sum y, sum y^2 (statistics for the BatchNorm that consumes this tensor)
      //   kMask         : sum g, sum g*xhat (BatchNorm backward)
      // thread = 4 adjacent channels x a slice of the rows; every slice leaves its partial sums in shared memory, 2*BN/4
      // threads add the slices up in a fixed order (shared-memory float atomics are compare-and-swap loops: with 8 - 16
      // slices per address they were ~0.5 us of every launch) and
      // ONE vector reduction per 4 channels leaves the CTA (every CTA of the grid adds into the same 2*Cout floats:
      // the L2 serialises per cache line, so the number of global atomics is what this costs).
      int valid = p.M_total - m0;
      valid = valid > 128 ? 128 : valid;
      constexpr int kQuads = BN / 4;            // 32 or 16
      constexpr int kGroups = kEpiThreads / kQuads;   // row slices: 4 .. 32
      constexpr int kRows = 128 / kGroups;            // 32 .. 4
      {
        const int quad = et % kQuads, grp = et / kQuads;
        const int c = quad * 4;
        const int coff = (c >> 6) * 16384 + (c & 7) * 2;
        const int chunk = (c & 63) >> 3;
        const uint8_t* vcol = (MODE == kMask ? sQ : sC) + coff;   // the values whose sum is taken
        const uint8_t* ycol = sC + coff;                          // kMask: raw BatchNorm input
        float2 s01 = make_float2(0.f, 0.f), s23 = s01, q01 = s01, q23 = s01;
        // BatchNorm statistics are sums of (y - pivot): the pivot of these four channels (0 for the backward sums)
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MODE == kPlain) pv = *reinterpret_cast<const float4*>(coef_s + c);
        else if constexpr (MODE != kMask) pv = *reinterpret_cast<const float4*>(p.stats + 2 * p.n_total + n_off + c);
        const float2 npv0 = make_float2(-pv.x, -pv.y), npv1 = make_float2(-pv.z, -pv.w);
#pragma unroll 8
        for (int i = 0; i < kRows; ++i) {
          const int r = grp * kRows + i;
          const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
          const uint2 u = *reinterpret_cast<const uint2*>(vcol + off);
          float2 f0 = bf2_to_f2(u.x), f1 = bf2_to_f2(u.y);
          if constexpr (MODE != kMask) {
            f0 = f2add(f0, npv0);      // x + (-p) == x - p
            f1 = f2add(f1, npv1);
          }
          float2 y0 = f0, y1 = f1;
          if constexpr (MODE == kMask) {
            const uint2 uy = *reinterpret_cast<const uint2*>(ycol + off);
            y0 = bf2_to_f2(uy.x);
            y1 = bf2_to_f2(uy.y);
          }
          if (r < valid) {
            s01 = f2add(s01, f0);
            s23 = f2add(s23, f1);
            q01 = f2fma(f0, y0, q01);
            q23 = f2fma(f1, y1, q23);
          }
        }
        *reinterpret_cast<float4*>(acc_s + grp * 2 * BN + c) = make_float4(s01.x, s01.y, s23.x, s23.y);
        *reinterpret_cast<float4*>(acc_s + grp * 2 * BN + BN + c) = make_float4(q01.x, q01.y, q23.x, q23.y);
      }
      named_bar_sync(1, kEpiThreads);
      if (et < 2 * kQuads) {
        const int which = et / kQuads, quad = et % kQuads;
        auto slices = [&](int w) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < kGroups; ++g) {
            const float4 a = *reinterpret_cast<const float4*>(acc_s + g * 2 * BN + w * BN + quad * 4);
            t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
          }
          return t;
        };
        float4 v4 = slices(which);
        if (MODE == kMask && which == 1) {
          // sum g*xhat = A * sum(g*y) + B * sum(g)
          const float4 sg = slices(0);
          const float4 cA = *reinterpret_cast<const float4*>(coef_s + 2 * BN + quad * 4);
          const float4 cB = *reinterpret_cast<const float4*>(coef_s + 3 * BN + quad * 4);
          v4 = make_float4(fmaf(cA.x, v4.x, cB.x * sg.x), fmaf(cA.y, v4.y, cB.y * sg.y),
                           fmaf(cA.z, v4.z, cB.z * sg.z), fmaf(cA.w, v4.w, cB.w * sg.w));
        }
        float* dst = p.stats + which * p.n_total + n_off + quad * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v4.x), "f"(v4.y), "f"(v4.z),
                     "f"(v4.w)
                     : "memory");
      }
    }
    if (HG_TS && blockIdx.x == 0 && et == 0) p.ts[8] = clock64();
    if (et == 0) tma_store_wait_read();
    if (HG_TS && blockIdx.x == 0 && et == 0) p.ts[9] = clock64();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (HG_TS && blockIdx.x == 0 && threadIdx.x == 0) p.ts[10] = clock64();
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
static inline int pad64(int c) { return (c + 63) & ~63; }

BnFoldDev make_fold(const HgBnFold* f, int C, long long count) {
  BnFoldDev d;
  d.stats = f->stats;
  d.gamma = f->gamma;
  d.beta = f->beta;
  d.rmean = f->running_mean;
  d.rvar = f->running_var;
  d.count = (float)count;
  d.eps = f->eps;
  d.relu = f->relu;
  d.use_running = f->use_running;
  d.C = C;
  d.Cp = pad64(C);
  return d;
}

template <int BN, int STAGES, int MINB, int MODE, bool ALIAS>
static int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                            const CUtensorMap& tmR, const ConvGemmParams& p, cudaStream_t st) {
  using L = ConvGemmSmem<BN, STAGES, MODE, ALIAS>;
  static bool attr_set = false;
  if (!attr_set) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, MINB, MODE, ALIAS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal + ConvThreads<MINB, BN>::kAccBytes));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.M_total, 128) * p.n_tiles);
  int cl = 0;   // experiment (hg_set_option "onewave_cluster"): launch the small one-wave grids as thread-block clusters
  if (MINB == 1 && g_onewave_cluster > 1 && (int)grid.x <= 32) {
    cl = g_onewave_cluster;
    while (cl > 1 && grid.x % cl) cl >>= 1;
  }
  if (cl > 1) {
    static bool np_set = false;
    if (!np_set) {
      HG_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, MINB, MODE, ALIAS>,
                                      cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      np_set = true;
    }
    launch_k_cluster(conv_gemm_kernel<BN, STAGES, MINB, MODE, ALIAS>, dim3(grid), dim3(ConvThreads<MINB, BN>::kAll),
                     L::kTotal + ConvThreads<MINB, BN>::kAccBytes, st, cl, tmA, tmB, tmC, tmR, p);
  } else {
    launch_k(conv_gemm_kernel<BN, STAGES, MINB, MODE, ALIAS>, dim3(grid), dim3(ConvThreads<MINB, BN>::kAll),
             L::kTotal + ConvThreads<MINB, BN>::kAccBytes, st, tmA, tmB, tmC, tmR, p);
  }
  HG_LAUNCH_OK("conv_gemm_kernel");
  count_launch();
  return HG_OK;
}

// Tile configuration: multi-wave grids run 2 pipeline stages with the staging tile aliased onto them = 64 KB per CTA =
// 3 CTAs/SM, for 1x1 AND 3x3 kernels (3 stages x 2 CTAs/SM for the 3x3: 53.7 us, 2 x 3: 49.8 us @64x64 -- more CTAs in
// different phases hide the long epilogue better than a deeper pipeline).  Grids of at most one
// wave (the 4x4 .. 16x16 levels of the hourglass) are pure latency: one CTA per SM with every K block's TMA load in
// flight at once (6 / 4 stages).
template <int BN, int MODE>
static int dispatch_conv_gemm(bool long_k, bool has_res, const CUtensorMap& tmA, const CUtensorMap& tmB,
                              const CUtensorMap& tmC, const CUtensorMap& tmR, const ConvGemmParams& p,
                              cudaStream_t st) {
  const bool single_wave = g_single_wave_deep && ceil_div(p.M_total, 128) * p.n_tiles <= kNumSMs;
  // the 2-stage / 3-CTAs-per-SM configuration of the 3x3 kernels only pays when a third CTA per SM exists: at
  // 32x32 (256 CTAs) every CTA is resident with two per SM and the deeper pipeline wins (12.0 vs 15.4 us)
  const bool three_cta = g_long_k_3cta && ceil_div(p.M_total, 128) * p.n_tiles > 2 * kNumSMs;
  if constexpr (MODE == kMask) {
    // the raw BatchNorm input tile is prefetched into a dedicated C buffer wherever shared memory allows; only the
    // multi-wave 3x3 kernel (two CTAs per SM) loads it after its (long) main loop
    if (single_wave) return long_k ? launch_conv_gemm<BN, 5, 1, kMask, false>(tmA, tmB, tmC, tmR, p, st)
                                   : launch_conv_gemm<BN, 4, 1, kMask, false>(tmA, tmB, tmC, tmR, p, st);
    if ((!long_k && g_short_alias) || (long_k && three_cta))
      return launch_conv_gemm<BN, 2, 3, kMask, true>(tmA, tmB, tmC, tmR, p, st);
    return long_k ? launch_conv_gemm<BN, 3, 2, kMask, true>(tmA, tmB, tmC, tmR, p, st)
                  : launch_conv_gemm<BN, 2, 2, kMask, false>(tmA, tmB, tmC, tmR, p, st);
  } else {
    constexpr int kShortMinB = MODE == kFold ? 2 : 3;  // the transform needs > 113 registers
    if constexpr (MODE != kPlainBnOut) if (has_res) {
      if (long_k) return launch_conv_gemm<BN, 3, 1, MODE, false>(tmA, tmB, tmC, tmR, p, st);
      if (!single_wave && g_short_alias) return launch_conv_gemm<BN, 2, kShortMinB, MODE, true>(tmA, tmB, tmC, tmR, p, st);
      return single_wave ? launch_conv_gemm<BN, 4, 1, MODE, false>(tmA, tmB, tmC, tmR, p, st)
                         : launch_conv_gemm<BN, 2, 2, MODE, false>(tmA, tmB, tmC, tmR, p, st);
    }
    if (long_k && !single_wave && three_cta)
      return launch_conv_gemm<BN, 2, kShortMinB, MODE, true>(tmA, tmB, tmC, tmR, p, st);
    if (long_k) return single_wave ? launch_conv_gemm<BN, 6, 1, MODE, true>(tmA, tmB, tmC, tmR, p, st)
                                   : launch_conv_gemm<BN, 3, 2, MODE, true>(tmA, tmB, tmC, tmR, p, st);
    if constexpr (MODE == kPlain) {
      if (!single_wave && g_short_1stage) return launch_conv_gemm<BN, 1, 4, MODE, true>(tmA, tmB, tmC, tmR, p, st);
    }
    return single_wave ? launch_conv_gemm<BN, 4, 1, MODE, true>(tmA, tmB, tmC, tmR, p, st)
                       : launch_conv_gemm<BN, 2, kShortMinB, MODE, true>(tmA, tmB, tmC, tmR, p, st);
  }
}

// persistent large-map kernel (conv_persist.cu)
bool conv_persist_eligible(int N, int H, int W, int Kp, int Np, int ntaps, const signed char* dh, const signed char* dw,
                           int stride, int parity, int mode, const float* out_nchw, bool has_res);
int conv_persist_launch(int N, int H, int W, int Kp, int Np, int mode, int ntaps, const signed char* dh,
                        const signed char* dw, const signed char* wt, const void* act, const void* wpk, int wtaps,
                        const float* bias, const void* res, void* out, float* stats, const BnFoldDev* fold,
                        cudaStream_t st);

// Geometry of one GEMM launch: the output grid [N,H,W] (128-pixel tiles), the A-operand tensor it gathers from and
// the list of filter taps (A offset + weight index) it walks.
struct GemmGeom {
  int N, H, W;      // output grid of this launch
  int Ha, Wa;       // spatial size of the A-operand tensor
  int stride;       // A origin = output (h, w) * stride + tap offset; also the element stride of the A box
  int ntaps;
  signed char dh[12], dw[12], wt[12];
  int wtaps;        // filter taps in the packed weight (R*S)
  int parity, par_h, par_w;   // out / res are parity views of a [N,2H,2W,Np] tensor
};

// act: [N,Ha,Wa,Kp] bf16 (A operand); wpk: [wtaps][Np][Kp] bf16; out/res: [N,H,W,Np] bf16 (or the parity view).
// mode kFold: `fold` describes the BatchNorm of act's channels (act is the raw tensor).
// mode kMask: `fold` describes the BatchNorm of out's channels, `res` is its raw input, `stats` receives the sums.
static int conv_gemm_bf16(const GemmGeom& g, int Kp, int Np, const void* act, const void* wpk, const float* bias,
                          const void* res, void* out, float* stats, float* out_nchw, int c_real, int mode,
                          const BnFoldDev* fold, cudaStream_t st) {
  const int N = g.N, H = g.H, W = g.W;
  if (!is_pow2(H) || !is_pow2(W) || W > 128) {
    set_error("conv_gemm_bf16: output H and W must be powers of two with W <= 128 (got %dx%d)", H, W);
    return HG_ERR_UNSUPPORTED;
  }
  if (Kp % 64 || Np % 64 || Np > 256 || Kp > 256) {
    set_error("conv_gemm_bf16: padded channels must be multiples of 64 and <= 256 (got %d -> %d)", Kp, Np);
    return HG_ERR_UNSUPPORTED;
  }
  const int bw = W < 128 ? W : 128;
  int bh = 128 / bw;
  if (bh > H) bh = H;
  const int bn = 128 / (bw * bh);
  const long long M = (long long)N * H * W;
  if (g.Ha == H && g.Wa == W &&
      conv_persist_eligible(N, H, W, Kp, Np, g.ntaps, g.dh, g.dw, g.stride, g.parity, mode, out_nchw, res != nullptr))
    // large-map 1x1 / 3x3 convolution: one resident CTA per SM (conv_persist.cu)
    return conv_persist_launch(N, H, W, Kp, Np, mode, g.ntaps, g.dh, g.dw, g.wt, act, wpk, g.wtaps, bias, res, out, stats,
                               fold, st);
  // N tile of at most 128 channels: short-K (1x1) kernels then fit two CTAs per SM (one CTA's epilogue overlaps
  // the other's TMA/MMA phase); 256 output channels = two N tiles that share the activation tile through L2.
  // One-wave grids (4x4 .. 16x16 levels) are pure latency: 64-channel N tiles double the CTA count, which halves
  // the per-CTA epilogue and weight-load time (the activation tile is then read by two SMs in parallel).
  int BN = Np > 128 ? 128 : Np;
  if (g_small_n_tiles && BN == 128 && ceil_div(M, 128) * (Np / 64) <= kNumSMs) BN = 64;
  // grids between one and two waves of 128-wide tiles (the 32x32 level: 256 CTAs on 148 SMs leave 40 SMs with one CTA
  // and 108 with two): 64-wide tiles double the CTAs (option, measured in DESIGN 8)
  if (g_mid_n_tiles && BN == 128) {
    const long long ctas = (long long)ceil_div(M, 128) * (Np / 128);
    if (ctas > kNumSMs && ctas <= 2 * kNumSMs) BN = 64;
  }
  if (BN != 64 && BN != 128) {
    set_error("conv_gemm_bf16: unsupported padded Cout %d", Np);
    return HG_ERR_UNSUPPORTED;
  }
  CUtensorMap tmA, tmB, tmC, tmR;
  {
    // a box of bw x bh x bn OUTPUT pixels walks the A tensor with element stride `stride` (TMA traverses
    // ceil(box / elementStride) elements per dimension; out-of-bounds elements are zero = convolution padding)
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)g.Wa, (uint64_t)g.Ha, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Kp * 2, (uint64_t)g.Wa * Kp * 2, (uint64_t)g.Ha * g.Wa * Kp * 2};
    uint32_t box[4] = {64, (uint32_t)(bw * g.stride), (uint32_t)(bh * g.stride), (uint32_t)bn};
    uint32_t es[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
    int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)Np, (uint64_t)g.wtaps};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Np * Kp * 2};
    uint32_t box[3] = {64, (uint32_t)BN, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wpk, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (!g.parity) {
    uint64_t dims[2] = {(uint64_t)Np, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)Np * 2};
    uint32_t box[2] = {64, 128};
    uint32_t es[2] = {1, 1};
    int rc = encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, res ? res : out, dims, str, box, es,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    // [N, 2H, 2W, Np] seen as {Np, pw (2), W, ph (2), N*H}: pixel (2h + ph, 2w + pw) of image n sits at
    // (c, pw, w, ph, n*H + h); a tile of 128 grid pixels is the box {64, 1, bw, 1, bh*bn}
    const uint64_t row = (uint64_t)2 * W * Np * 2;   // bytes of one full-resolution image row
    uint64_t dims[5] = {(uint64_t)Np, 2, (uint64_t)W, 2, (uint64_t)N * H};
    uint64_t str[4] = {(uint64_t)Np * 2, (uint64_t)Np * 4, row, row * 2};
    uint32_t box[5] = {64, 1, (uint32_t)bw, 1, (uint32_t)(bh * bn)};
    uint32_t es[5] = {1, 1, 1, 1, 1};
    int rc = encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, res ? res : out, dims, str, box, es,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M_total = (int)M;
  p.H = H;
  p.W = W;
  p.stride = g.stride;
  p.Hin = g.Ha;
  p.Win = g.Wa;
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) {
    p.tap_dh[t] = g.dh[t];
    p.tap_dw[t] = g.dw[t];
    p.tap_w[t] = g.wt[t];
  }
  p.parity = g.parity;
  p.par_h = g.par_h;
  p.par_w = g.par_w;
  p.kchunks = Kp / 64;
  p.n_total = Np;
  p.c_real = c_real;
  p.bias = bias;
  p.stats = stats;
  p.out_nchw = out_nchw;
  p.has_res = res != nullptr;
  p.n_tiles = Np / BN;
  if (fold) p.fold = *fold;
  p.ts = g_dbg_ts;
  if (mode == kPlainBnOut) {
    if (!fold || !fold->use_running || res || stats) {
      set_error("conv_gemm_bf16: output BatchNorm needs running statistics and takes no residual / statistics");
      return HG_ERR_BAD_ARG;
    }
  }
  if (g.parity && (mode != kPlain || stats || out_nchw)) {
    set_error("conv_gemm_bf16: a parity-view output takes the plain epilogue only");
    return HG_ERR_BAD_ARG;
  }
  const bool long_k = g.ntaps * (Kp / 64) > 4;
  const bool has_res = res != nullptr;
  if (mode == kMask && (!res || !stats)) {
    set_error("conv_gemm_bf16: mask mode needs the raw BatchNorm input and the reduction buffer");
    return HG_ERR_BAD_ARG;
  }
  if (BN == 64) {
    if (mode == kFold) return dispatch_conv_gemm<64, kFold>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
    if (mode == kMask) return dispatch_conv_gemm<64, kMask>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
    if (mode == kPlainBnOut) return dispatch_conv_gemm<64, kPlainBnOut>(long_k, false, tmA, tmB, tmC, tmR, p, st);
    return dispatch_conv_gemm<64, kPlain>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
  }
  if (mode == kFold) return dispatch_conv_gemm<128, kFold>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
  if (mode == kMask) return dispatch_conv_gemm<128, kMask>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
  if (mode == kPlainBnOut) return dispatch_conv_gemm<128, kPlainBnOut>(long_k, false, tmA, tmB, tmC, tmR, p, st);
  return dispatch_conv_gemm<128, kPlain>(long_k, has_res, tmA, tmB, tmC, tmR, p, st);
}

static inline int out_size(int in, int k, int stride, int pad, int dil) {
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}

// y[N,Ho,Wo,Cout_p] = conv(x[N,H,W,Cin_p]) (+ epilogue of `mode`): stride 1 or 2.
int conv_tc_fprop(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias, const void* res, void* y,
                  float* stats, float* out_nchw, int mode, const BnFoldDev* fold, cudaStream_t st) {
  GemmGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N;
  g.H = out_size(d->H, d->R, d->stride, d->pad, d->dil);
  g.W = out_size(d->W, d->S, d->stride, d->pad, d->dil);
  g.Ha = d->H;
  g.Wa = d->W;
  g.stride = d->stride;
  g.wtaps = d->R * d->S;
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s) {
      g.dh[g.ntaps] = (signed char)(r * d->dil - d->pad);
      g.dw[g.ntaps] = (signed char)(s * d->dil - d->pad);
      g.wt[g.ntaps] = (signed char)(r * d->S + s);
      ++g.ntaps;
    }
  return conv_gemm_bf16(g, pad64(d->Cin), pad64(d->Cout), x, w_fprop, bias, res, y, stats, out_nchw, d->Cout, mode,
                        fold, st);
}

// dx[N,H,W,Cin_p] = conv_transpose(dy[N,Ho,Wo,Cout_p]) [+ addend] (+ epilogue of `mode`, stride 1 only).
// Stride 2: hi = 2*ho - pad + r*dil, so the input pixels of one parity class (hi & 1, wi & 1) receive contributions
// from a fixed subset of the taps, each a plain shifted read of dy: one launch per class over the half-resolution
// grid, stored through a parity view of dx (no scatter, no zero-stuffed dy).
int conv_tc_dgrad(const HgConvDesc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx, float* red,
                  int mode, const BnFoldDev* fold, cudaStream_t st) {
  const int Ho = out_size(d->H, d->R, d->stride, d->pad, d->dil);
  const int Wo = out_size(d->W, d->S, d->stride, d->pad, d->dil);
  const int Kp = pad64(d->Cout), Np = pad64(d->Cin);
  GemmGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N;
  g.Ha = Ho;
  g.Wa = Wo;
  g.stride = 1;
  g.wtaps = d->R * d->S;
  if (d->stride == 1) {
    g.H = d->H;
    g.W = d->W;
    for (int r = 0; r < d->R; ++r)
      for (int s = 0; s < d->S; ++s) {
        g.dh[g.ntaps] = (signed char)(d->pad - r * d->dil);
        g.dw[g.ntaps] = (signed char)(d->pad - s * d->dil);
        g.wt[g.ntaps] = (signed char)(r * d->S + s);
        ++g.ntaps;
      }
    return conv_gemm_bf16(g, Kp, Np, dy, w_dgrad, nullptr, addend, dx, red, nullptr, 0, mode, fold, st);
  }
  if (mode != kPlain) {
    set_error("conv_tc_dgrad: the BatchNorm-fused data gradient takes stride-1 convolutions only");
    return HG_ERR_UNSUPPORTED;
  }
  g.H = d->H / 2;
  g.W = d->W / 2;
  g.parity = 1;
  GemmGeom cls[4];
  bool any_empty = false;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      GemmGeom& c = cls[ph * 2 + pw];
      c = g;
      c.par_h = ph;
      c.par_w = pw;
      for (int r = 0; r < d->R; ++r) {
        const int th = ph + d->pad - r * d->dil;
        if (th & 1) continue;
        for (int s = 0; s < d->S; ++s) {
          const int tw = pw + d->pad - s * d->dil;
          if (tw & 1) continue;
          c.dh[c.ntaps] = (signed char)(th / 2);
          c.dw[c.ntaps] = (signed char)(tw / 2);
          c.wt[c.ntaps] = (signed char)(r * d->S + s);
          ++c.ntaps;
        }
      }
      any_empty = any_empty || c.ntaps == 0;
    }
  if (any_empty) {
    // no tap reaches some parity class (1x1 stride 2): those pixels are the addend alone / zero
    const size_t bytes = (size_t)d->N * d->H * d->W * Np * 2;
    if (addend) HG_CUDA_OK(cudaMemcpyAsync(dx, addend, bytes, cudaMemcpyDeviceToDevice, st));
    else HG_CUDA_OK(cudaMemsetAsync(dx, 0, bytes, st));
  }
  for (int i = 0; i < 4; ++i) {
    if (cls[i].ntaps == 0) continue;
    int rc = conv_gemm_bf16(cls[i], Kp, Np, dy, w_dgrad, nullptr, addend, dx, nullptr, nullptr, 0, kPlain, nullptr, st);
    if (rc) return rc;
  }
  return HG_OK;
}

// ======================================================================================================
// wgrad:  dw[tap][co][ci] += sum_m dy[m, co] * x[m (+) tap, ci]
//
// GEMM with the PIXELS as the reduction dimension: both operands are read "MN-major" straight from the NHWC
// tensors (a TMA box of {64 channels, 64 pixels} is exactly the 128B-swizzled MN-major tile tcgen05.mma
// reads), so nothing is transposed in memory.  M = 128 out-channels per CTA, N = Cin (<= 256), up to three
// filter taps per CTA (their accumulators sit side by side in TMEM).  The pixel range is split across CTAs
// and partial sums are reduced with vector fp32 atomics into the GEMM-layout gradient buffer (the same
// buffer accumulates every call site of a shared weight, reference try_with_torch.py:217,224-237).
//
// FOLD = true: x is the RAW input of a BatchNorm(+ReLU) whose output the convolution consumed; the x tiles
// are rewritten in shared memory as [relu](scale*x + shift) (zero in the padding) before the MMA reads them.
// ======================================================================================================
struct WgradParams {
  int H, W, N;       // grid of dy (the reduction dimension): H, W are the OUTPUT size of the convolution
  int stride;        // x origin = (h, w) * stride + tap offset
  int Hin, Win;      // spatial size of x
  int taps_s;        // filter width S
  int dil, pad;
  int kpx;           // pixels per K block (64, or 128: half as many TMA operations per byte)
  int tap_rows;      // taps handled per CTA (T)
  int n_panels;      // 64-channel input panels per CTA (Cin_p / 64 / n_groups)
  int n_groups;      // CTAs that split the input channels
  int Cin_p, Cout_p;
  int total_kb;      // ceil(M / 64)
  int kb_per_cta;
  int stages;
  int dbg;
  int bulk_reduce;   // epilogue: partial tile -> shared memory -> cp.reduce.async.bulk (else per-thread red.v4)
  int stage_bytes;
  int halo;          // 1: the T = 3 taps of a CTA are ONE filter column (dh = -1, 0, +1): one x box with a halo row above
                     //    and below serves all three (the B operand of tap dh starts dh * W pixel rows further down)
  int xp_bytes;      // bytes of one 64-channel x panel of a stage (halo: (rows + 2) * W * 128)
  float* dw;         // [taps][Cout_p][Cin_p] fp32, accumulated
  float* dbias;      // optional [Cout]: += sum over pixels of dy (bias gradient), from an extra all-ones N slab
  int Cout;          // real output channels (rows of dbias)
  BnFoldDev fold;
};

template <bool FOLD>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the shared array keeps the address space: LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tmem_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  uint64_t* ready_bar = bars + 18;                             // [8]
  float* coef_s = reinterpret_cast<float*>(bars + 32);         // scale[256], shift[256]
  // Bias gradient on the tensor core: dbias[co] = sum_m dy[m, co] * 1 -- one more N slab (16 columns) whose B operand is
  // a constant all-ones tile, so the dy tiles already in shared memory are reused and no separate column-sum pass over
  // dy exists (it was a second kernel: 67 MB re-read per 128->256 convolution at 64x64).
  uint8_t* ones_s = reinterpret_cast<uint8_t*>(bars) + 4096;   // [kpx K rows][128 B], 1024-byte aligned
  const int pb = p.kpx * 128;                                  // bytes of one 64-channel operand panel

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = p.n_panels * 64;
  const int T = p.tap_rows;
  const int tap0 = p.halo ? (int)(blockIdx.y / p.n_groups) : (int)(blockIdx.y / p.n_groups) * T;   // halo: filter column
  const int tap_step = p.halo ? p.taps_s : 1;                                                    // tap of accumulator t = tap0 + t * tap_step
  const int pn0 = (blockIdx.y % p.n_groups) * p.n_panels;   // first input-channel panel of this CTA
  const int co_off = blockIdx.z * 128;
  const int kb_beg = blockIdx.x * p.kb_per_cta;
  int kb_end = kb_beg + p.kb_per_cta;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int nkb = kb_end - kb_beg;
  const bool with_bias = p.dbias != nullptr && blockIdx.y == 0;   // tap group 0, input-channel group 0 only
  uint32_t cols = 32;
  while ((int)cols < T * N + (p.dbias != nullptr ? 16 : 0)) cols <<= 1;
  if (p.dbias != nullptr) {
    for (int i = threadIdx.x; i < pb / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(ones_s)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int a_bytes = 2 * pb;
  const int ksteps = p.kpx / 16;

  if (nkb > 0) {
    if (warp == 0) {
      // warp-uniform loop, one elected lane issues; H and W are powers of two (tc_eligible): shifts, no divisions
      {
        const int lw = 31 - __clz(p.W), lhw = 31 - __clz(p.H * p.W);
        const int nst = p.stages;
        int st = 0;
        uint32_t ph = 1;                    // empty-barrier parity: a fresh barrier passes a parity-1 wait
        for (int i = 0; i < nkb; ++i) {
          uint8_t* sA = smem + st * p.stage_bytes;
          uint8_t* sB = sA + a_bytes;
          const int m0 = (kb_beg + i) * p.kpx;
          const int n0 = m0 >> lhw;
          const int rem = m0 & ((1 << lhw) - 1);
          const int h0 = rem >> lw;
          const int w0 = rem & (p.W - 1);
          mbar_wait(&empty_bar[st], ph);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[st], p.stage_bytes);
            tma_load_2d(sA, &tmDy, &full_bar[st], co_off, m0);
            tma_load_2d(sA + pb, &tmDy, &full_bar[st], co_off + 64, m0);
            if (p.halo) {
              // one box of (rows + 2) image rows per input-channel panel: column tap0, rows h0 - 1 .. h0 + rows
              for (int pn = 0; pn < p.n_panels; ++pn)
                tma_load_4d(sB + pn * p.xp_bytes, &tmX, &full_bar[st], (pn0 + pn) * 64, w0 + tap0 - p.pad, h0 - p.pad, n0);
            } else {
              int r = tap0 / p.taps_s, sx = tap0 - r * p.taps_s;
              for (int t = 0; t < T; ++t) {
                const int dh = r * p.dil - p.pad, dw = sx * p.dil - p.pad;
                for (int pn = 0; pn < p.n_panels; ++pn)
                  tma_load_4d(sB + (t * p.n_panels + pn) * pb, &tmX, &full_bar[st], (pn0 + pn) * 64, w0 * p.stride + dw,
                              h0 * p.stride + dh, n0);
                if (++sx == p.taps_s) { sx = 0; ++r; }
              }
            }
          }
          __syncwarp();
          if (++st == nst) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
      pdl_trigger();
    } else if (warp == 1) {
      // one thread, counters instead of divisions, descriptors by addition (see conv_gemm_kernel)
      {
        const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
        const uint32_t idesc1 = make_idesc_bf16(128, 16, 1, 1);
        const uint32_t s0 = smem_u32(smem);
        const uint64_t adesc0 = make_smem_desc(s0, pb, 1024);
        const uint64_t bdesc0 = make_smem_desc(s0 + a_bytes, p.halo ? p.xp_bytes : pb, 1024);
        const uint64_t odesc0 = make_smem_desc(smem_u32(ones_s), pb, 1024);
        const uint32_t st_step = (uint32_t)p.stage_bytes >> 4;
        const uint32_t t_step = (uint32_t)(p.halo ? p.W * 128 : p.n_panels * pb) >> 4;   // B operand of the next tap
        const int nst = p.stages;
        int st = 0;
        uint32_t ph = 0;
        uint32_t accum = 0;
        for (int i = 0; i < nkb; ++i) {
          if constexpr (FOLD) mbar_wait(&ready_bar[st], ph);
          else mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)st * st_step);
          uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)st * st_step);
          if (elect_one()) {
            for (int t = 0; t < T; ++t, bdesc += t_step) {
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k)   // 16 pixels further = 2048 B in both MN-major operands
                umma_bf16(tmem_base + t * N, adesc + 128 * k, bdesc + 128 * k, idesc, (accum | (uint32_t)k) ? 1u : 0u);
            }
            if (with_bias) {
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(tmem_base + T * N, adesc + 128 * k, odesc0 + 128 * k, idesc1, (accum | (uint32_t)k) ? 1u : 0u);
            }
            umma_commit(&empty_bar[st]);
            if (i == nkb - 1) umma_commit(tmem_full);
          }
          __syncwarp();
          accum = 1;
          if (++st == nst) { st = 0; ph ^= 1; }
        }
      }
      pdl_trigger();
    } else {
      const int sub = warp & 3;
      const int et = threadIdx.x - 64;
      (void)et;
      if constexpr (FOLD) {
        for (int c = et; c < p.Cin_p; c += 128) {
          float mu, is, sc, sh;
          bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
          coef_s[c] = sc;
          coef_s[256 + c] = sh;
        }
        named_bar_sync(1, 128);
        // rewrite the x tiles of every stage (see conv_gemm_kernel): thread = one 16-byte channel chunk x 4 rows of
        // each 64-row tile
        const int jch = et & 7;
        const int rbase = et >> 3;                 // rows rbase + 16 * i, i < 4
        const int swz = (jch ^ (rbase & 7)) << 4;
        const int hw = p.H * p.W;
        const bool relu = p.fold.relu != 0;
        for (int i = 0; i < nkb; ++i) {
          const int st = i % p.stages;
          const uint32_t ph = (i / p.stages) & 1;
          int hrow[4], wrow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int mm = (kb_beg + i) * 64 + rbase + 16 * k;
            const int rem = mm % hw;
            hrow[k] = mm / hw < p.N ? rem / p.W : -0x100000;
            wrow[k] = rem % p.W;
          }
          uint8_t* sB = smem + st * p.stage_bytes + a_bytes + rbase * 128 + swz;
          mbar_wait(&full_bar[st], ph);
          for (int t = 0; t < T; ++t) {
            const int tap = tap0 + t;
            const int r = tap / p.taps_s, s = tap - r * p.taps_s;
            const int dh = r * p.dil - p.pad, dw = s * p.dil - p.pad;
            uint32_t vmask = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((unsigned)(hrow[k] * p.stride + dh) < (unsigned)p.Hin &&
                  (unsigned)(wrow[k] * p.stride + dw) < (unsigned)p.Win)
                vmask |= 1u << k;
            for (int pn = 0; pn < p.n_panels; ++pn) {
              float sc[8], sh[8];
              load_coef8(coef_s + (pn0 + pn) * 64 + jch * 8, sc);
              load_coef8(coef_s + 256 + (pn0 + pn) * 64 + jch * 8, sh);
              uint8_t* base = sB + (t * p.n_panels + pn) * 8192;
              uint4 u[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(base + k * 2048);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                u[k] = (vmask >> k) & 1u ? bn_relu_chunk(u[k], sc, sh, relu) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
              for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(base + k * 2048) = u[k];
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&ready_bar[st]);
        }
      }
      const int co = co_off + sub * 32 + lane;
      const bool row_ok = co < p.Cout_p;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      pdl_trigger();
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
      if (with_bias) {
        float v[16];
        tmem_ld16(taddr + T * N, v);
        tmem_ld_wait();
        if (co < p.Cout) atomicAdd(p.dbias + co, v[0]);
      }
      if (p.bulk_reduce) {
        // The partial tile goes through shared memory (the pipeline stages are idle: every MMA has completed) as a
        // linear [128 co][N ci] fp32 image per tap and is added to the gradient by ONE bulk reduce per tap: the
        // per-thread red.global.add.v4 of a TMEM row touch 32 different 512-byte-apart rows per warp instruction
        // (12288 scattered 16-byte atomics per CTA; measured 5-9 us of every wgrad launch).
        float* stage = reinterpret_cast<float*>(smem);
        const int rloc = sub * 32 + lane;
        const int nchunk = N / 32;
        for (int t = 0; t < ((HG_DBG_TS && p.dbg == 2) ? 0 : T); ++t) {
          float* rowp = stage + ((size_t)t * 128 + rloc) * N;
          for (int j = 0; j < nchunk; ++j) {
            float v[32];
            tmem_ld32(taddr + t * N + j * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int qq = (q + lane) & 7;   // a quarter warp stores 8 different 16-byte columns: no bank conflict
              float a = v[0], b = v[1], c = v[2], d = v[3];
#pragma unroll
              for (int z = 1; z < 8; ++z)
                if (qq == z) { a = v[z * 4]; b = v[z * 4 + 1]; c = v[z * 4 + 2]; d = v[z * 4 + 3]; }
              *reinterpret_cast<float4*>(rowp + j * 32 + qq * 4) = make_float4(a, b, c, d);
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0 && !(HG_DBG_TS && p.dbg)) {
          int rows = p.Cout_p - co_off;
          if (rows > 128) rows = 128;
          for (int t = 0; t < T; ++t)
            bulk_reduce_add_f32(p.dw + ((size_t)(tap0 + t * tap_step) * p.Cout_p + co_off) * p.Cin_p, stage + (size_t)t * 128 * N,
                                (uint32_t)rows * N * 4);
          tma_store_commit();
          tma_store_wait_read();
        }
      } else {
        // All CTAs of a split finish together and add into the SAME tile: start each CTA at a different column so
        // that concurrent atomics hit different addresses (the L2 atomic unit serialises per address).
        const int nchunk = N / 32;
        for (int tt = 0; tt < ((HG_DBG_TS && p.dbg == 2) ? 0 : T); ++tt) {
          const int t = (tt + blockIdx.x) % T;
          float* dst = p.dw + ((size_t)(tap0 + t * tap_step) * p.Cout_p + co) * p.Cin_p + pn0 * 64;
          for (int jj = 0; jj < nchunk; ++jj) {
            const int j = (jj + blockIdx.x) % nchunk;
            float v[32];
            tmem_ld32(taddr + t * N + j * 32, v);
            tmem_ld_wait();
            if (row_ok && !(HG_DBG_TS && p.dbg == 1)) {
              const int rot = (blockIdx.x / nchunk) & 7;
  #pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int qq = (q + rot) & 7;
                float a = v[0], b = v[1], c = v[2], d = v[3];
  #pragma unroll
                for (int z = 1; z < 8; ++z)
                  if (qq == z) { a = v[z * 4]; b = v[z * 4 + 1]; c = v[z * 4 + 2]; d = v[z * 4 + 3]; }
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j * 32 + qq * 4), "f"(a),
                             "f"(b), "f"(c), "f"(d)
                             : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, cols);
  }
}

int colsum_launch(int dtype, const void* dy, long long M, int Cp, int C, float* out, cudaStream_t st);

int conv_wgrad_bf16(const HgConvDesc* d, const void* x, const void* dy, float* dw, float* dbias,
                    const BnFoldDev* fold, cudaStream_t st) {
  const int Cin_p = pad64(d->Cin), Cout_p = pad64(d->Cout);
  // the reduction runs over the pixels of dy: H, W = OUTPUT size of the convolution; x is read at stride `stride`
  const int H = out_size(d->H, d->R, d->stride, d->pad, d->dil), W = out_size(d->W, d->S, d->stride, d->pad, d->dil);
  const long long M = (long long)d->N * H * W;
  if (dw) {
    const int taps = d->R * d->S;
    // Three taps per CTA share one dy tile (fewer loads), but the CTA then pushes 3 x 64 KB of fp32 adds through its
    // SM's L2 port (~20 B/clk: the epilogue atomics are 5-9 us of a launch).  Small maps cannot fill the SMs with
    // K slices anyway: there one tap per CTA triples the CTAs that share the atomics.
    int T = 1;
    if (taps == 9 && Cin_p <= 128 && (M + 63) / 64 > g_wgrad_t1_max_kb) T = 3;
    // K block = 64 pixels, or 128 on the large maps (option wgrad_kpx): the TMA unit of an SM retires about one bulk
    // tensor operation per ~320 cycles whatever its size (16 KB boxes: ~50 B/clk, the forward kernels; 8 KB boxes:
    // ~26 B/clk, this kernel), so twice the pixels per box is half the operations per byte
    int kpx = 64;
    if (g_wgrad_kpx == 128 && !fold && (M + 63) / 64 > g_wgrad_t1_max_kb && M % 128 == 0 && W <= 128 &&
        (long long)H * W >= 128) {
      const int sb = 2 * 16384 + T * (Cin_p / 64) * 16384;
      if (2 * sb <= g_wgrad_smem_kb * 1024) kpx = 128;
      else if (T == 3 && 2 * (2 * 16384 + (Cin_p / 64) * 16384) <= g_wgrad_smem_kb * 1024) {
        kpx = 128;   // one tap per CTA so that two 128-pixel stages fit
        T = 1;
      }
    }
    const int bw = W < kpx ? W : kpx;
    int bh = kpx / bw;
    if (bh > H) bh = H;
    const int bn = kpx / (bw * bh);
    // 3x3 on the large maps: a CTA takes one filter COLUMN (three taps dh = -1, 0, +1) and loads ONE x box with a halo
    // row above and below per K block instead of three shifted boxes: 96 KB instead of 192 KB of operands per 128
    // pixels and three taps at 64x64 (the kernel runs at the per-SM operand rate, ~50 B/clk)
    bool halo = false;
    if (g_wgrad_halo && taps == 9 && d->R == 3 && d->S == 3 && d->stride == 1 && d->dil == 1 && d->pad == 1 && !fold &&
        kpx == 128 && bn == 1 && bh * bw == kpx && H % bh == 0 && Cin_p <= 128 && M / kpx >= g_wgrad_halo_min_kb) {
      const int sbh = 2 * 16384 + (Cin_p / 64) * (bh + 2) * bw * 128;
      if (2 * sbh <= g_wgrad_smem_kb * 1024) {
        halo = true;
        T = 3;
      }
    }
    CUtensorMap tmDy, tmX;
    {
      uint64_t dims[2] = {(uint64_t)Cout_p, (uint64_t)M};
      uint64_t str[1] = {(uint64_t)Cout_p * 2};
      uint32_t box[2] = {64, (uint32_t)kpx};
      uint32_t es[2] = {1, 1};
      int rc = encode_tmap(&tmDy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {(uint64_t)Cin_p, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
      uint64_t str[3] = {(uint64_t)Cin_p * 2, (uint64_t)d->W * Cin_p * 2, (uint64_t)d->H * d->W * Cin_p * 2};
      uint32_t box[4] = {64, (uint32_t)(bw * d->stride), (uint32_t)((halo ? bh + 2 : bh) * d->stride), (uint32_t)bn};
      uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
      int rc = encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.H = H;
    p.W = W;
    p.N = d->N;
    p.stride = d->stride;
    p.Hin = d->H;
    p.Win = d->W;
    p.taps_s = d->S;
    p.dil = d->dil;
    p.pad = d->pad;
    p.tap_rows = T;
    p.kpx = kpx;
    // small maps: split the input channels over CTAs as well (same reason as one tap per CTA: the fp32 adds a CTA
    // pushes through its SM's L2 port are what a small wgrad launch costs)
    // (B200, batch 32, us per launch: 3x3 128->128 @4x4 11.5 -> 4.5, 1x1 256->128 8.1 -> 4.5, 3x3 @16x16 15.8 -> 10.7)
    p.n_groups = 1;
    // larger maps, 1x1: ONE 64-channel input panel per CTA as well (two for 256 -> 256 at 64x64).  The CTAs then split K
    // less finely (148 / n_groups slices) and each adds a quarter of the columns: a quarter of the fp32 atomics, which
    // outweighs the N = 64 MMAs and the dy tile re-read through L2 (B200, batch 32, us per launch: 256->128 @64x64
    // 29.8 -> 21.2, @32x32 14.2 -> 10.1; 256->256 @32x32 17.4 -> 12.3; 128->256 @32x32 12.6 -> 10.2; 256->16 @64x64
    // 23.7 -> 17.6).  3x3: slower (42.7 -> 49.3 at 64x64), stays whole.
    int big_panels = g_wgrad_big_n_panels;   // 0 = this policy, > 0 = fixed, < 0 = never
    if (big_panels == 0) big_panels = taps == 1 ? ((Cin_p == 256 && Cout_p == 256 && M >= 131072) ? 2 : 1) : -1;
    if ((M + 63) / 64 > g_wgrad_t1_max_kb && big_panels > 0 && Cin_p / 64 > big_panels && (Cin_p / 64) % big_panels == 0)
      p.n_groups = Cin_p / 64 / big_panels;
    if ((M + 63) / 64 <= g_wgrad_t1_max_kb && g_wgrad_small_n_panels >= 0) {
      int per_cta = g_wgrad_small_n_panels;
      if (per_cta == 0) per_cta = (taps == 1 || (M + 63) / 64 <= 32) ? 1 : 2;
      if (Cin_p / 64 > per_cta && (Cin_p / 64) % per_cta == 0) p.n_groups = Cin_p / 64 / per_cta;
    }
    p.n_panels = Cin_p / 64 / p.n_groups;
    p.Cin_p = Cin_p;
    p.Cout_p = Cout_p;
    p.total_kb = (int)((M + kpx - 1) / kpx);
    p.halo = halo ? 1 : 0;
    p.xp_bytes = halo ? (bh + 2) * bw * 128 : kpx * 128;
    p.stage_bytes = halo ? 2 * kpx * 128 + p.n_panels * p.xp_bytes : (2 + T * p.n_panels) * kpx * 128;
    p.stages = (g_wgrad_smem_kb * 1024) / p.stage_bytes;
    if (p.stages < 2) p.stages = 2;
    if (p.stages > 6) p.stages = 6;
    const int tap_groups = taps / T;
    const int mgroups = (Cout_p + 127) / 128;
    // one wave of CTAs, and at least 8 K-blocks (512 pixels) of work per CTA: the split-K partial sums are
    // reduced with atomics, so small problems must not be cut into many slices
    // (small maps are latency-bound: there, two K blocks per CTA and more atomics beat a long serial loop)
    int min_kb = (M + 63) / 64 <= 256 ? 2 : 8;   // in 64-pixel blocks
    min_kb = min_kb / (kpx / 64) > 0 ? min_kb / (kpx / 64) : 1;
    int nsplit = kNumSMs / (tap_groups * mgroups * p.n_groups);
    if (nsplit > p.total_kb / min_kb) nsplit = p.total_kb / min_kb;
    if (nsplit < 1) nsplit = 1;
    p.kb_per_cta = (p.total_kb + nsplit - 1) / nsplit;
    nsplit = (p.total_kb + p.kb_per_cta - 1) / p.kb_per_cta;
    p.dw = dw;
    p.dbg = g_wgrad_dbg;
    p.bulk_reduce = g_wgrad_bulk_reduce;
    if (fold) p.fold = *fold;
    if (p.n_groups > 1) p.bulk_reduce = 0;   // the CTA's rows are not contiguous in dw
    if (p.bulk_reduce && p.stages * p.stage_bytes < T * Cin_p * 512) {
      // room for the fp32 staging image [T][128][Cin_p]: more stages if they fit, else the per-thread atomics
      const int need = (T * Cin_p * 512 + p.stage_bytes - 1) / p.stage_bytes;
      if (need <= 8 && need * p.stage_bytes <= 220 * 1024) p.stages = need;
      else p.bulk_reduce = 0;
    }
    // dbias rides on the weight-gradient GEMM (all-ones N slab) whenever both are wanted
    p.dbias = (dbias && g_wgrad_fused_bias) ? dbias : nullptr;
    p.Cout = d->Cout;
    if (p.dbias) dbias = nullptr;
    const int smem_bytes = p.stages * p.stage_bytes + 4096 + kpx * 128 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
      HG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
      HG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
      attr_set = true;
    }
    dim3 grid(nsplit, tap_groups * p.n_groups, mgroups);
    if (fold) launch_k(conv_wgrad_kernel<true>, dim3(grid), dim3(192), smem_bytes, st, tmDy, tmX, p);
    else launch_k(conv_wgrad_kernel<false>, dim3(grid), dim3(192), smem_bytes, st, tmDy, tmX, p);
    HG_LAUNCH_OK("conv_wgrad_kernel");
    count_launch();
  }
  if (dbias) return colsum_launch(HG_BF16, dy, M, Cout_p, d->Cout, dbias, st);
  return HG_OK;
}

}  // namespace hg
