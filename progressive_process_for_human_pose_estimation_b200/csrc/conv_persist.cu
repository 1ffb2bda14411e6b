// Persistent convolution kernel for the large maps (64x64 / 32x32 levels at batch 32): 1x1 and 3x3 (stride 1,
// dilation 1), fprop and dgrad, y[m, co] = sum_{dh, dw, ci} a[m + (dh, dw), ci] * w[tap(dh, dw)][co][ci].
// (conv1 / conv2 / conv3 of every ResidualBlock, lin, conv3 / conv4 of creatModel: reference try_with_torch.py:186-193,
// 199-207,248,272-273, and their data gradients.)
//
// Why: the tile-per-CTA kernel of conv_tc.cu pays barrier / TMEM / descriptor set-up, a cold two-stage pipeline and a
// serial epilogue for every 128 pixels.  Measured on B200 (tools/gpu_l2_probe.py): its 1x1 256->128 launch at 64x64
// takes 29.5 us with every operand L2-resident and 30.5 us streaming from HBM -- the kernel is bound by the life of
// its CTAs, not by memory (the HBM time of its 100 MB is 15 us).  Here ONE CTA per SM stays resident and walks a
// contiguous range of 128-pixel units:
//   * warp 0 streams activation boxes (and weight tiles) through TMA rings, running ahead across tiles;
//   * warp 1 issues tcgen05.mma into one of TWO accumulator sets in TMEM (2 x 256 columns);
//   * warps 2..17 drain the other set meanwhile: TMEM -> registers -> (+bias, +residual | ReLU mask) -> bf16 -> swizzled
//     staging -> TMA store, and keep the per-channel sums (BatchNorm statistics of the output / the two
//     BatchNorm-backward sums) in REGISTERS across all tiles of the CTA: one shared-memory reduction and one vector
//     atomic per 4 channels per CTA at the end of the kernel.
// An accumulator set holds two 128 x 128 parts: two 128-pixel sub-tiles of a 256-pixel tile (Np <= 128: both share
// every weight tile), or the two 128-column halves of a 128-pixel tile (Np = 256: one MMA with N = 256).
// 1x1: the whole weight matrix stays resident in shared memory when it fits next to the rings.
// 3x3: ONE activation box of R+2 image rows (the tile plus a halo row above and below), shifted by dw, serves the three
// taps dh = -1, 0, +1 -- the A operand of tap dh is the same shared-memory image read (dh+1)*W pixel rows further
// down (whole image rows: every start stays 1024-byte aligned for the 128-byte swizzle).  That halves the operand bytes
// per FLOP twice over (96 KB per 12.6 MFLOP at 64x64 instead of 32 KB per 2.1 MFLOP).  What paces it (measured,
// profiles/r02g_mma_issue_probe.txt): an SS-mode tcgen05.mma reads its shared-memory operands at ~64 B/clk -- M=128, N=128,
// K=16 is 8 KB = 125 cycles (64 if MAC-bound) -- and the chip-wide L2 -> SM rate (340 MB per launch = 28.7 us with the
// MMAs switched off).  With 128 output channels (TR) the accumulators are therefore TRANSPOSED: the weight tile is the
// M = 128 operand, the 256 pixel rows of the box ONE N = 256 operand (12 KB per 2 x 128x128x16 MACs instead of 16 KB),
// D[channel][pixel] in TMEM, per-channel sums in registers: 42.9 us at 64x64 vs 51.0 us on the tile kernel; on by default
// at >= persist_min_units (persist_3x3 = 1; 2 = also the pixel-major 3x3 shapes, which only draw level).
#include "hg_common.cuh"

// -DHG_DBG_TS=1 (make DBG=1): CTA 0 accumulates the cycles each role spends waiting on each barrier / in each epilogue
// phase into the dbg_ts buffer (hg_set_option("dbg_ts", 1), printed by ("dbg_ts", 3)); ps_dbg 1 = no epilogue work,
// 2 = no MMAs issued, 3 = both (what the TMA stream alone takes)
#ifndef HG_DBG_TS
#define HG_DBG_TS 0
#endif
#if HG_DBG_TS
#define PS_TIC long long _t0 = clock64()
#define PS_TOC(var) var += clock64() - _t0
#else
#define PS_TIC
#define PS_TOC(var)
#endif

namespace hg {

extern long long* g_dbg_ts;
int g_ps_dbg = 0;
int g_persist_1x1 = 1;
int g_persist_3x3 = 1;               // 128-channel 3x3 convolutions at >= persist_min_units: transposed persistent kernel (43 vs 51 us @64x64)
int g_persist_transposed = 1;     // 3x3 with 128 output channels: accumulators as [channel][pixel] (one N = 256 MMA per
                                  // 256-pixel tile and K step instead of two N = 128 ones; no column pass in the epilogue)
int g_persist_dynamic = 0;        // 1: tiles handed out by a grid-wide atomic counter (see PsQueue); measured: isolated launches 2 - 8 % slower, step 970 -> 961 images/s
static int* g_ps_ctr_pool = nullptr;      // kPsCtrSlots x {next tile, CTAs done}; every launch (graph node) takes the next slot
static int g_ps_ctr_next = 0;
constexpr int kPsCtrSlots = 8192;
int g_persist_min_units = 512;    // at least this many 128-pixel units: two 256-pixel tiles per SM and more (64x64 at
                                  // batch 32; at 32x32 every CTA has ONE tile, nothing to pipeline: step 883 vs 894 images/s)

struct PsParams {
  int M_total;        // N*H*W (a multiple of 128)
  int H, W;
  int units;          // M_total / 128
  int upi;            // 128-pixel units per image (3x3: a tile never crosses an image; 1x1: = units)
  int tile_units;     // 2 (Np <= 128) or 1 (Np = 256)
  int kchunks;        // Kp / 64
  int tw, th;         // tap columns / tap rows: 1 x 1 or 3 x 3
  int nA, nB;         // ring depths
  int nY;             // residual / raw-input buffers (2: the next part's tile is requested one part ahead)
  int a_bytes;        // one activation box
  int b_resident;     // 1: every weight tile has its own slot and is loaded once
  int has_res;        // kPlain: residual added;  kMask: tmR is the raw BatchNorm input (always)
  const float* bias;  // [Np] or null
  float* stats;       // kPlain: {S1, S2, pivot}[3*Np] or null;  kMask: {sum g, sum g*xhat}[2*Np]
  BnFoldDev fold;     // kMask: BatchNorm of the OUTPUT channels
  int* ctr;           // dynamic tile scheduling: {next tile, CTAs done} (self-resetting) or null = static unit ranges
  int num_tiles;      // dynamic: tiles of tile_units units
  signed char wt[3][3];   // weight matrix of the tap whose A offset is (dh, dw) = (i - th/2, j - tw/2)
  long long* ts;      // debug counters or null
  int dbg;
  int offB, offC, offY, offBar;
};

constexpr int kPsThreads = 576;   // warp 0 producer, warp 1 MMA, warps 2..17 epilogue
constexpr int kPsEpi = 512;       // four epilogue warps per SM sub-partition: the row / column passes are chains of
                                  // dependent shared-memory round trips, two warps per scheduler left them latency-bound
constexpr int kPsMisc = 8192;     // barriers (512 B) + bias [256] + coefficients [4][256] + column sums [2][256]

// 128-pixel units [u, u1) of this CTA, cut into tiles of `tile_units` units (one where the range or the image ends)
struct PsTiles {
  int u, u1, upi, tu;
  __device__ __forceinline__ PsTiles(const PsParams& p) {
    const int base = p.units / (int)gridDim.x, rem = p.units % (int)gridDim.x;
    const int b = (int)blockIdx.x;
    u = b * base + (b < rem ? b : rem);
    u1 = u + base + (b < rem ? 1 : 0);
    upi = p.upi;
    tu = p.tile_units;
  }
  __device__ __forceinline__ bool next(int& m0, int& mt) {
    if (u >= u1) return false;
    mt = (tu == 2 && u1 - u >= 2 && (u % upi) != upi - 1) ? 2 : 1;
    m0 = u * 128;
    u += mt;
    return true;
  }
};

// Tile queue: the producer warp decides which tile comes next -- its static unit range, or (p.ctr != null) the next
// tile of a grid-wide atomic counter -- and publishes {first unit, units} through a small shared-memory ring; the MMA
// issuer and the epilogue warps read the same sequence.  Dynamic order matters in the training step: the persistent
// kernels share the GPU with the one-wave kernels of the low-resolution hourglass levels (other stream lanes, higher
// priority), so some CTAs start several microseconds late; with static ranges the launch lasts until the LAST CTA has
// worked through its whole range, with the counter a late CTA simply finds less left to do.
constexpr int kPsQ = 8;
struct PsQueue {
  int* tq;
  uint64_t* full;
  uint64_t* empty;
  __device__ __forceinline__ bool get(int idx, int& m0, int& mt) const {
    const int q = idx & (kPsQ - 1);
    mbar_wait(&full[q], (uint32_t)(idx / kPsQ) & 1u);
    const int v = *reinterpret_cast<volatile int*>(tq + q);
    if (v < 0) return false;
    m0 = (v >> 2) * 128;
    mt = v & 3;
    return true;
  }
  __device__ __forceinline__ void release(int idx) const { mbar_arrive(&empty[idx & (kPsQ - 1)]); }
};

template <int MODE, int NP, bool TR>
__global__ void __launch_bounds__(kPsThreads, 1)
conv_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const PsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  static_assert(!TR || NP == 128, "transposed accumulators: the weight tile is the M = 128 operand");
  constexpr int kBBytes = NP * 128;                  // one weight tile: NP out-channels x 64 in-channels
  constexpr int kPartCols = NP > 128 ? 128 : NP;     // columns of one accumulator part
  constexpr int kPanels = kPartCols / 64;
  constexpr int kCBytes = kPanels * 16384;           // one part of the output: 128 pixels x kPartCols channels
  constexpr uint32_t kTmemCols = 4 * kPartCols;      // two accumulator sets x two parts
  uint8_t* sA = smem;                          // [nA][a_bytes]
  uint8_t* sB = smem + p.offB;                 // [nB][NP rows x 128 B]
  uint8_t* sC = smem + p.offC;                 // [kPanels][128 rows x 128 B]: output staging (TMA store source)
  uint8_t* sY = smem + p.offY;                 // [nY] residual (kPlain) / raw BatchNorm input (kMask) of a part
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
  uint64_t* a_full = bars;                     // [8]
  uint64_t* a_empty = bars + 8;                // [8]
  uint64_t* b_full = bars + 16;                // [8]
  uint64_t* b_empty = bars + 24;               // [8]
  uint64_t* tmem_full = bars + 32;             // [2]
  uint64_t* tmem_empty = bars + 34;            // [2]
  uint64_t* y_full = bars + 36;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);
  PsQueue tqu;
  tqu.full = bars + 40;                        // [kPsQ]
  tqu.empty = bars + 48;                       // [kPsQ]
  tqu.tq = reinterpret_cast<int*>(bars + 56);  // [kPsQ]
  float* bias_s = reinterpret_cast<float*>(bars + 64);   // [256]
  float* coef_s = bias_s + 256;                            // kMask: scale / shift / A / B [4][256];  kPlain: pivots
  float* acc_s = coef_s + 1024;                            // [2][256] per-CTA column sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    prefetch_tmap(&tmR);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 1);
    }
    mbar_init(&y_full[0], 1);
    mbar_init(&y_full[1], 1);
    for (int s = 0; s < kPsQ; ++s) {
      mbar_init(&tqu.full[s], 1);
      mbar_init(&tqu.empty[s], 2);             // MMA issuer + epilogue
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // warp-uniform loop; one elected lane issues (see the MMA issuer)
    {
      const int hw = p.H * p.W;
      PsTiles tiles(p);
      int m0, mt;
      int sa = 0, sb = 0;
      uint32_t pha = 1, phb = 1;          // empty-barrier parities (a fresh barrier passes a parity-1 wait)
      bool first_tile = true;
      long long w_ae = 0, w_be = 0;
      const long long tstart = clock64();
      const int nA = p.nA, nB = p.nB, th = p.th, tw = p.tw;
      const bool nodata = HG_DBG_TS && (p.dbg & 4);   // debug: no data movement at all (what the MMA stream alone takes)
      // tile sequence: static range, or first tile = blockIdx.x and then whatever the grid-wide counter hands out
      const bool dyn = p.ctr != nullptr;
      int qi = 0;
      auto publish = [&](int v) {
        const int q = qi & (kPsQ - 1);
        mbar_wait(&tqu.empty[q], ((uint32_t)(qi / kPsQ) & 1u) ^ 1u);
        if (elect_one()) {
          tqu.tq[q] = v;
          mbar_arrive(&tqu.full[q]);
        }
        __syncwarp();
        ++qi;
      };
      auto dyn_tile = [&](int t) {              // tile index -> {first unit * 4 + units} or -1
        if (t >= p.num_tiles) return -1;
        const int u0 = t * p.tile_units;
        const int n = p.units - u0 < p.tile_units ? p.units - u0 : p.tile_units;
        return u0 * 4 + n;
      };
      int cur;
      if (dyn) cur = dyn_tile((int)blockIdx.x);
      else cur = tiles.next(m0, mt) ? (m0 / 128) * 4 + mt : -1;
      publish(cur);
      while (cur >= 0) {
        m0 = (cur >> 2) * 128;
        mt = cur & 3;
        // the next tile is decided (and published) before this one's loads are issued: the consumers can look ahead
        int nxt;
        if (dyn) {
          int t = 0;
          if (lane == 0) t = atomicAdd(p.ctr, 1) + (int)gridDim.x;
          t = __shfl_sync(0xffffffffu, t, 0);
          nxt = dyn_tile(t);
        } else {
          int m0n, mtn;
          nxt = tiles.next(m0n, mtn) ? (m0n / 128) * 4 + mtn : -1;
        }
        publish(nxt);
        const int n = m0 / hw;
        const int h0 = (m0 - n * hw) / p.W;
        int sbr = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int j = 0; j < tw; ++j) {
            {
              PS_TIC;
              mbar_wait(&a_empty[sa], pha);
              PS_TOC(w_ae);
            }
            if (elect_one()) {
              if (nodata) {
                mbar_arrive(&a_full[sa]);
              } else {
                mbar_expect_tx(&a_full[sa], (uint32_t)p.a_bytes);
                if (th == 1) tma_load_2d(sA + sa * p.a_bytes, &tmA, &a_full[sa], kc * 64, m0);
                else tma_load_4d(sA + sa * p.a_bytes, &tmA, &a_full[sa], kc * 64, j - 1, h0 - 1, n);
              }
            }
            __syncwarp();
            if (++sa == nA) { sa = 0; pha ^= 1; }
            for (int i = 0; i < th; ++i) {
              if (p.b_resident) {
                if (first_tile && elect_one()) {
                  mbar_expect_tx(&b_full[sbr], kBBytes);
                  tma_load_3d(sB + sbr * kBBytes, &tmB, &b_full[sbr], kc * 64, 0, p.wt[i][j]);
                }
                __syncwarp();
                ++sbr;
              } else {
                {
                  PS_TIC;
                  mbar_wait(&b_empty[sb], phb);
                  PS_TOC(w_be);
                }
                if (elect_one()) {
                  if (nodata) {
                    mbar_arrive(&b_full[sb]);
                  } else {
                    mbar_expect_tx(&b_full[sb], kBBytes);
                    tma_load_3d(sB + sb * kBBytes, &tmB, &b_full[sb], kc * 64, 0, p.wt[i][j]);
                  }
                }
                __syncwarp();
                if (++sb == nB) { sb = 0; phb ^= 1; }
              }
            }
          }
        }
        first_tile = false;
        cur = nxt;
      }
      if (HG_DBG_TS && p.ts && blockIdx.x == 0 && lane == 0) {
        p.ts[0] = clock64() - tstart;
        p.ts[1] = w_ae;
        p.ts[2] = w_be;
      }
    }
    __syncwarp();
    pdl_trigger();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ONE thread runs the whole loop.  It is a chain of dependent scalar instructions, and what it costs per weight tile is
    // what the tensor core idles: with ring slots found by integer division, descriptors rebuilt per tile and the warp
    // re-converged around every instruction group it took ~1000 cycles per weight tile (four N = 256 MMAs need 512) --
    // measured with every data movement switched off (ps_dbg 5).  Slots and phases are counters, descriptors are adds.
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, NP, 0, 0);
      int m0, mt;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      long long w_te = 0, w_af = 0, w_bf = 0;
      const long long tstart = clock64();
      const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 16, 1024);
      const uint32_t a_step = (uint32_t)p.a_bytes >> 4;       // descriptor address field counts 16-byte units
      const uint32_t row_step = (uint32_t)p.W * 128u >> 4;    // one image row of pixels (tap row i)
      const int nA = p.nA, nB = p.nB, th = p.th, per_kc = p.tw;
      const bool resident = p.b_resident != 0;
      const bool no_mma = HG_DBG_TS && (p.dbg & 2);
      for (int it = 0; tqu.get(it, m0, mt); ++it) {
        const int acc = it & 1;
        __syncwarp();                                  // every lane has read the queue slot
        if (elect_one()) tqu.release(it);
        {
          PS_TIC;
          mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
          PS_TOC(w_te);
        }
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * 2 * kPartCols;
        const uint32_t idt = TR ? (mt == 2 ? make_idesc_bf16(128, 256, 0, 0) : make_idesc_bf16(128, 128, 0, 0)) : idesc;
        uint32_t accum = 0;
        int sbr = 0;                       // resident weights: slot = running tile index
        for (int kj = p.kchunks * per_kc; kj > 0; --kj) {
          {
            PS_TIC;
            mbar_wait(&a_full[sa], pha);
            PS_TOC(w_af);
          }
          const uint64_t a_desc = a_desc0 + (uint64_t)((uint32_t)sa * a_step);
          for (int i = 0; i < th; ++i) {
            const int slot = resident ? sbr : sb;
            {
              PS_TIC;
              mbar_wait(&b_full[slot], resident ? 0u : phb);   // resident: completed once, parity 0 stays satisfied
              PS_TOC(w_bf);
            }
            tc_fence_after();
            const uint64_t bdesc = b_desc0 + (uint64_t)((uint32_t)slot * (uint32_t)(kBBytes >> 4));
            const uint64_t xdesc = a_desc + (uint64_t)((uint32_t)i * row_step);
            const bool leader = elect_one();
            if (leader && !no_mma) {
              if constexpr (TR) {
                // transposed: D[channel][pixel] -- the weight tile is the M = 128 operand, the mt * 128 pixel rows of the
                // box (tap row i: i image rows further down) are ONE N = 128 / 256 operand
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tacc, bdesc + 2 * k, xdesc + 2 * k, idt, (accum | (uint32_t)k) ? 1u : 0u);
              } else {
                // sub-tile t: 128 pixel rows (16 KB) further down
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tacc, xdesc + 2 * k, bdesc + 2 * k, idt, (accum | (uint32_t)k) ? 1u : 0u);
                if (mt == 2) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(tacc + kPartCols, xdesc + 1024 + 2 * k, bdesc + 2 * k, idt, (accum | (uint32_t)k) ? 1u : 0u);
                }
              }
            }
            accum = 1;
            if (resident) {
              ++sbr;
            } else {
              if (leader) umma_commit(&b_empty[sb]);
              if (++sb == nB) { sb = 0; phb ^= 1; }
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
          if (++sa == nA) { sa = 0; pha ^= 1; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
      }
      if (HG_DBG_TS && p.ts && blockIdx.x == 0 && lane == 0) {
        p.ts[4] = clock64() - tstart;
        p.ts[5] = w_te;
        p.ts[6] = w_af;
        p.ts[7] = w_bf;
      }
    }
    __syncwarp();
    pdl_trigger();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int et = threadIdx.x - 64;            // 0..255
    const int sub = warp & 3;                   // TMEM lane quarter of this warp
    const int colq = (warp - 2) >> 2;           // which 32-column chunk of the part this warp stages
    const int row = sub * 32 + lane;
    const bool row_warp = colq * 32 < kPartCols;   // 64-column parts: half of the warps only take part in the column pass
    const bool need_y = MODE == kMask || p.has_res;
    for (int c = et; c < NP; c += kPsEpi) bias_s[c] = p.bias ? p.bias[c] : 0.f;
    for (int c = et; c < 2 * 256; c += kPsEpi) acc_s[c] = 0.f;
    if constexpr (MODE == kMask) {
      for (int c = et; c < NP; c += kPsEpi) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
        coef_s[c] = sc;                 // ReLU mask: scale * y + shift > 0 (the forward's own expression)
        coef_s[256 + c] = sh;
        coef_s[512 + c] = is;           // xhat = y * A + B  ->  sum g*xhat = A * sum(g*y) + B * sum(g)
        coef_s[768 + c] = -mu * is;
      }
    } else if constexpr (MODE == kPlainBnOut) {
      // inference: y = [relu](scale * (acc + bias) + shift) with the running statistics of the OUTPUT channels
      for (int c = et; c < NP; c += kPsEpi) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
        coef_s[c] = sc;
        bias_s[c] = fmaf(p.bias ? p.bias[c] : 0.f, sc, sh);
      }
    } else {
      // statistics are sums of (y - pivot) (bn.cu): the pivots of the output channels
      if (p.stats != nullptr)
        for (int c = et; c < NP; c += kPsEpi) coef_s[c] = p.stats[2 * NP + c];
    }
    const bool relu = p.fold.relu != 0;
    // column pass: thread = 4 adjacent channels (quad) x a slice of the rows (rg); its sums stay in registers over
    // every tile of the CTA.  Np = 256: the two parts of a tile are different channels -> two register sets.
    constexpr int kQuads = kPartCols / 4;          // 16 or 32 column quads
    constexpr int kGroups = kPsEpi / kQuads;       // 16 or 8 row slices
    constexpr int kRows = 128 / kGroups;           // 8 or 16 rows each
    constexpr int kSets = NP > 128 ? 2 : 1;
    const int quad = et % kQuads, rg = et / kQuads;
    float cs[kSets][4], cq[kSets][4];
#pragma unroll
    for (int s = 0; s < kSets; ++s)
#pragma unroll
      for (int e = 0; e < 4; ++e) cs[s][e] = cq[s][e] = 0.f;

    auto load_y = [&](int buf, int m, int ccol) {
      mbar_expect_tx(&y_full[buf], kCBytes);
      for (int pnl = 0; pnl < kPanels; ++pnl)
        tma_load_2d(sY + buf * kCBytes + pnl * 16384, &tmR, &y_full[buf], ccol + pnl * 64, m);
    };
    int m0, mt;
    bool have = tqu.get(0, m0, mt);
    if (need_y && have && et == 0) load_y(0, m0, 0);
    named_bar_sync(1, kPsEpi);
    int g = 0;                                  // parts drained so far
    long long w_tf = 0, w_y = 0, t_row = 0, t_col = 0, t_sw = 0;
    const long long tstart = clock64();
    for (int it = 0; have; ++it) {
      const int acc = it & 1;
      int m0n = 0, mtn = 0;
      const bool have_next = tqu.get(it + 1, m0n, mtn);
      {
        PS_TIC;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        PS_TOC(w_tf);
      }
      tc_fence_after();
      if (HG_DBG_TS && (p.dbg & 1)) {
        tc_fence_before();
        named_bar_sync(1, kPsEpi);
        if (et == 0) mbar_arrive(&tmem_empty[acc]);
        if (et == 0) tqu.release(it);
        have = have_next;
        m0 = m0n;
        mt = mtn;
        continue;
      }
      const int nparts = NP > 128 ? 2 : mt;
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        if (pp >= nparts) break;
        const int ybuf = p.nY == 2 ? (g & 1) : 0;
        const int mp = NP > 128 ? m0 : m0 + pp * 128;      // first pixel of the part
        const int ccol = NP > 128 ? pp * 128 : 0;          // first output channel of the part
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 2 * kPartCols + pp * kPartCols + colq * 32;
        float v[32];
        if (row_warp) tmem_ld32(taddr, v);
        {
          PS_TIC;
          if (need_y) mbar_wait(&y_full[ybuf], (g / p.nY) & 1);
          if (row_warp) tmem_ld_wait();
          PS_TOC(w_y);
        }
        const uint8_t* sYp = sY + ybuf * kCBytes;
        if (need_y && p.nY == 2 && et == 0) {
          // the other buffer is free (its part was finished before this one started): request the next part now
          if (pp + 1 < nparts) load_y(ybuf ^ 1, NP > 128 ? m0 : m0 + (pp + 1) * 128, NP > 128 ? (pp + 1) * 128 : 0);
          else if (have_next) load_y(ybuf ^ 1, m0n, 0);
        }
#if HG_DBG_TS
        const long long _tr = clock64();
#endif
        // ---- row pass: registers -> (+bias, +residual | mask) -> bf16 -> swizzled staging ----
        if constexpr (TR) {
          // transposed accumulator: this thread owns ONE channel (TMEM lane) and 32 pixels (columns) of the part.  Bias,
          // mask coefficients and pivot are per-thread constants, the per-channel sums are plain register adds (no
          // column pass); the staging tile is written (and the residual / raw-input tile read) 2 bytes at a time --
          // the 32 lanes of a warp cover 64 contiguous bytes of one pixel row, so neither access conflicts.
          const int ch = row;                                   // output channel (NP == 128: one part = all channels)
          const int boff = (ch >> 6) * 16384 + (ch & 7) * 2;
          const int chunk = (ch & 63) >> 3;
          const int r0 = colq * 32;                             // first pixel row of this warp inside the part
          if constexpr (MODE == kMask) {
            const float cS = coef_s[ch], cT = coef_s[256 + ch];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int r = r0 + i;
              const int off = boff + r * 128 + ((chunk ^ (r & 7)) << 4);
              const __nv_bfloat16 yb = *reinterpret_cast<const __nv_bfloat16*>(sYp + off);
              const float y = __bfloat162float(yb);
              const bool keep = !relu || fmaf(y, cS, cT) > 0.f;
              const __nv_bfloat16 gb = __float2bfloat16_rn(keep ? v[i] : 0.f);
              *reinterpret_cast<__nv_bfloat16*>(sC + off) = gb;
              const float gq = __bfloat162float(gb);
              cs[0][0] += gq;
              cq[0][0] = fmaf(gq, y, cq[0][0]);
            }
          } else if constexpr (MODE == kPlainBnOut) {
            const float sc = coef_s[ch], bi = bias_s[ch];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int r = r0 + i;
              const int off = boff + r * 128 + ((chunk ^ (r & 7)) << 4);
              const float t = fmaf(v[i], sc, bi);
              *reinterpret_cast<__nv_bfloat16*>(sC + off) = __float2bfloat16_rn(relu ? fmaxf(t, 0.f) : t);
            }
          } else {
            const float bch = bias_s[ch];
            const float pv = p.stats != nullptr ? coef_s[ch] : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int r = r0 + i;
              const int off = boff + r * 128 + ((chunk ^ (r & 7)) << 4);
              float o = v[i] + bch;
              if (p.has_res) o += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sYp + off));
              const __nv_bfloat16 ob = __float2bfloat16_rn(o);
              *reinterpret_cast<__nv_bfloat16*>(sC + off) = ob;
              const float f = __bfloat162float(ob) - pv;
              cs[0][0] += f;
              cq[0][0] = fmaf(f, f, cq[0][0]);
            }
          }
        } else if (row_warp) {
          const int col0 = colq * 32;                           // first column of the chunk inside the part
          const int pnl = col0 / 64;
          const int chunk0 = (col0 % 64) / 8;
          const int rowoff = pnl * 16384 + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int swz = ((chunk0 + q) ^ (row & 7)) << 4;
            const int ch = ccol + col0 + q * 8;                 // output channel of o[0]
            float o[8];
            if constexpr (MODE == kMask) {
              const uint4 u = *reinterpret_cast<const uint4*>(sYp + rowoff + swz);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
              float cS[8], cT[8];
              load_coef8(coef_s + ch, cS);
              load_coef8(coef_s + 256 + ch, cT);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = f2fma(__bfloat1622float2(h[e]), make_float2(cS[2 * e], cS[2 * e + 1]),
                                       make_float2(cT[2 * e], cT[2 * e + 1]));
                const bool k0 = !relu || t.x > 0.f;
                const bool k1 = !relu || t.y > 0.f;
                o[2 * e] = k0 ? v[q * 8 + 2 * e] : 0.f;
                o[2 * e + 1] = k1 ? v[q * 8 + 2 * e + 1] : 0.f;
              }
            } else if constexpr (MODE == kPlainBnOut) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float t = fmaf(v[q * 8 + e], coef_s[ch + e], bias_s[ch + e]);
                o[e] = relu ? fmaxf(t, 0.f) : t;
              }
            } else {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch), b1 = *reinterpret_cast<const float4*>(bias_s + ch + 4);
              float2 o2[4];
              o2[0] = f2add(make_float2(v[q * 8 + 0], v[q * 8 + 1]), make_float2(b0.x, b0.y));
              o2[1] = f2add(make_float2(v[q * 8 + 2], v[q * 8 + 3]), make_float2(b0.z, b0.w));
              o2[2] = f2add(make_float2(v[q * 8 + 4], v[q * 8 + 5]), make_float2(b1.x, b1.y));
              o2[3] = f2add(make_float2(v[q * 8 + 6], v[q * 8 + 7]), make_float2(b1.z, b1.w));
              if (p.has_res) {
                const uint4 u = *reinterpret_cast<const uint4*>(sYp + rowoff + swz);
                o2[0] = f2add(o2[0], bf2_to_f2(u.x));
                o2[1] = f2add(o2[1], bf2_to_f2(u.y));
                o2[2] = f2add(o2[2], bf2_to_f2(u.z));
                o2[3] = f2add(o2[3], bf2_to_f2(u.w));
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) { o[2 * e] = o2[e].x; o[2 * e + 1] = o2[e].y; }
            }
            uint4 w;
            __nv_bfloat162* hw2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) hw2[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
            *reinterpret_cast<uint4*>(sC + rowoff + swz) = w;
          }
        }
        // every TMEM read of this accumulator set is done after its last part: hand it back to the MMA warp
        if (pp == nparts - 1) tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1, kPsEpi);
#if HG_DBG_TS
        const long long _tc = clock64();
        t_row += _tc - _tr;
#endif
        if (et == 0) {
          if (pp == nparts - 1) mbar_arrive(&tmem_empty[acc]);
          for (int pnl = 0; pnl < kPanels; ++pnl) tma_store_2d(&tmC, sC + pnl * 16384, ccol + pnl * 64, mp);
          tma_store_commit();
        }
        // ---- column pass: per-channel sums of what was just staged (bf16, exactly what the consumers read) ----
        if (!TR && p.stats != nullptr) {
          constexpr int kS = NP > 128 ? 1 : 0;
          const int si = kS * pp;                        // register set (compile-time: pp is unrolled)
          const int c = quad * 4;
          const int coff = (c >> 6) * 16384 + (c & 7) * 2;
          const int chunk = (c & 63) >> 3;
          const uint8_t* vcol = sC + coff;
          const uint8_t* ycol = (MODE == kMask ? sYp : sC) + coff;
          float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);   // statistics are sums of (y - pivot)
          if constexpr (MODE != kMask) pv = *reinterpret_cast<const float4*>(coef_s + ccol + c);
          const float2 npv0 = make_float2(-pv.x, -pv.y), npv1 = make_float2(-pv.z, -pv.w);
          float2 s01 = make_float2(cs[si][0], cs[si][1]), s23 = make_float2(cs[si][2], cs[si][3]);
          float2 q01 = make_float2(cq[si][0], cq[si][1]), q23 = make_float2(cq[si][2], cq[si][3]);
#pragma unroll 8
          for (int k = 0; k < kRows; ++k) {
            const int r = rg * kRows + k;
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
            s01 = f2add(s01, f0);
            s23 = f2add(s23, f1);
            q01 = f2fma(f0, y0, q01);
            q23 = f2fma(f1, y1, q23);
          }
          cs[si][0] = s01.x; cs[si][1] = s01.y; cs[si][2] = s23.x; cs[si][3] = s23.y;
          cq[si][0] = q01.x; cq[si][1] = q01.y; cq[si][2] = q23.x; cq[si][3] = q23.y;
        }
#if HG_DBG_TS
        const long long _ts = clock64();
        t_col += _ts - _tc;
#endif
        // the TMA store must have read the staging tile before the next row pass rewrites it
        if (et == 0) tma_store_wait_read();
        named_bar_sync(1, kPsEpi);
#if HG_DBG_TS
        t_sw += clock64() - _ts;
#endif
        if (need_y && p.nY == 1 && et == 0) {
          if (pp + 1 < nparts) load_y(0, NP > 128 ? m0 : m0 + (pp + 1) * 128, NP > 128 ? (pp + 1) * 128 : 0);
          else if (have_next) load_y(0, m0n, 0);
        }
        ++g;
      }
      // every epilogue thread read this tile's queue slot before the barriers above
      if (et == 0) tqu.release(it);
      have = have_next;
      m0 = m0n;
      mt = mtn;
    }
    if (p.stats != nullptr) {
      // the row slices add up in shared memory (once per kernel), then one vector atomic per 4 channels per CTA
      if constexpr (TR) {
        // four warps (pixel-column groups) per channel
        atomicAdd(acc_s + row, cs[0][0]);
        atomicAdd(acc_s + 256 + row, cq[0][0]);
      } else {
#pragma unroll
      for (int s = 0; s < kSets; ++s)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicAdd(acc_s + s * 128 + quad * 4 + e, cs[s][e]);
          atomicAdd(acc_s + 256 + s * 128 + quad * 4 + e, cq[s][e]);
        }
      }
      named_bar_sync(1, kPsEpi);
      for (int q = et; q < 2 * (NP / 4); q += kPsEpi) {
        const int which = q / (NP / 4), qd = q % (NP / 4);
        float4 v4 = *reinterpret_cast<const float4*>(acc_s + which * 256 + qd * 4);
        if (MODE == kMask && which == 1) {
          const float4 sg = *reinterpret_cast<const float4*>(acc_s + qd * 4);
          const float4 cA = *reinterpret_cast<const float4*>(coef_s + 512 + qd * 4);
          const float4 cB = *reinterpret_cast<const float4*>(coef_s + 768 + qd * 4);
          v4 = make_float4(fmaf(cA.x, v4.x, cB.x * sg.x), fmaf(cA.y, v4.y, cB.y * sg.y),
                           fmaf(cA.z, v4.z, cB.z * sg.z), fmaf(cA.w, v4.w, cB.w * sg.w));
        }
        float* dst = p.stats + which * NP + qd * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v4.x), "f"(v4.y), "f"(v4.z),
                     "f"(v4.w)
                     : "memory");
      }
    }
    if (et == 0) tma_store_wait_all();
    if (HG_DBG_TS && p.ts && blockIdx.x == 0 && et == 0) {
      p.ts[8] = clock64() - tstart;
      p.ts[9] = w_tf;
      p.ts[10] = w_y;
      p.ts[11] = t_row;
      p.ts[12] = t_col;
      p.ts[13] = t_sw;
    }
    pdl_trigger();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.ctr != nullptr && threadIdx.x == 0) {
    // the last CTA to get here (every CTA has made its final fetch) re-arms the counter for the next launch of this node
    __threadfence();
    if (atomicAdd(p.ctr + 1, 1) == (int)gridDim.x - 1) {
      p.ctr[0] = 0;
      p.ctr[1] = 0;
      __threadfence();
    }
  }
}

// Shared-memory plan; returns the dynamic shared-memory size or 0 when the shape does not fit.
static int ps_plan(int W, int Kp, int Np, int taps, bool need_y, PsParams& p) {
  const int tile_units = Np > 128 ? 1 : 2;
  const int tile_px = tile_units * 128;
  const int th = taps == 9 ? 3 : 1;
  const int part_cols = Np > 128 ? 128 : Np;
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - kPsMisc;
  p.tile_units = tile_units;
  p.tw = p.th = th;
  p.kchunks = Kp / 64;
  p.a_bytes = th == 3 ? (tile_px / W + 2) * W * 128 : tile_px * 128;
  const int bbytes = Np * 128;
  const int cbytes = (part_cols / 64) * 16384;
  const int btot = taps * p.kchunks * bbytes;
  int nA = 0, nB = 0, nY = 0;
  // preference: two residual / raw-input buffers (the next part's tile is in flight while this one is consumed) as long
  // as the rings keep their depth; resident weights next to >= 2 activation boxes; else rings for both operands (two
  // activation boxes are the minimum; a third one only when a box's worth of weight tiles still fits next to it)
  for (int ny = need_y ? 2 : 0; ny >= (need_y ? 1 : 0) && nA == 0; --ny) {
    const int ring = budget - cbytes - ny * cbytes;
    if (taps * p.kchunks <= 8 && ring - btot >= (ny == 2 ? 3 : 2) * p.a_bytes) {
      p.b_resident = 1;
      nB = taps * p.kchunks;
      nA = (ring - btot) / p.a_bytes;
      nY = ny;
    } else if (ny <= 1 || ring - 3 * p.a_bytes >= 3 * bbytes) {
      p.b_resident = 0;
      nA = 3;
      if (ring - 3 * p.a_bytes < 3 * bbytes) nA = 2;
      nB = (ring - nA * p.a_bytes) / bbytes;
      nY = ny;
      if (nB < (th == 3 ? 3 : 2)) return 0;
    }
  }
  const int ybytes = nY * cbytes;
  if (nA > 6) nA = 6;
  if (nB > 8) nB = 8;
  if (nA < 2) return 0;
  p.nA = nA;
  p.nB = nB;
  p.nY = nY;
  p.offB = nA * p.a_bytes;
  p.offC = p.offB + nB * bbytes;
  p.offY = p.offC + cbytes;
  p.offBar = p.offY + ybytes;
  return p.offBar + kPsMisc + 1024;
}

// taps of the launch as (dh, dw): eligible = one tap (0, 0), or the nine offsets {-1,0,1}^2, each once
bool conv_persist_eligible(int N, int H, int W, int Kp, int Np, int ntaps, const signed char* dh, const signed char* dw,
                           int stride, int parity, int mode, const float* out_nchw, bool has_res) {
  if (stride != 1 || parity || out_nchw != nullptr) return false;
  if (mode != kPlain && mode != kMask && mode != kPlainBnOut) return false;
  if (!(Np == 64 || Np == 128 || Np == 256) || Kp % 64 || Kp > 256) return false;
  const long long M = (long long)N * H * W;
  if (M % 128 != 0 || M / 128 < g_persist_min_units || M > 0x7fffffffLL) return false;
  if (ntaps == 1) {
    if (!g_persist_1x1 || dh[0] != 0 || dw[0] != 0) return false;
  } else if (ntaps == 9) {
    if (!g_persist_3x3) return false;
    // 1 (default): only the shape with transposed accumulators (128 output channels), which beats the tile kernel;
    // 2: every 3x3 shape the kernel supports (tests / probes)
    if (g_persist_3x3 == 1 && !(Np == 128 && g_persist_transposed)) return false;
    const int tile_px = Np > 128 ? 128 : 256;
    if (!is_pow2(W) || !is_pow2(H) || W > 128 || W < 16 || tile_px / W > H || tile_px % W || (H * W) % tile_px) return false;
    unsigned seen = 0;
    for (int t = 0; t < 9; ++t) {
      if (dh[t] < -1 || dh[t] > 1 || dw[t] < -1 || dw[t] > 1) return false;
      seen |= 1u << ((dh[t] + 1) * 3 + dw[t] + 1);
    }
    if (seen != 0x1FFu) return false;
  } else {
    return false;
  }
  PsParams p;
  return ps_plan(W, Kp, Np, ntaps, mode == kMask || has_res, p) > 0;
}

template <int MODE, int NP, bool TR = false>
static int ps_launch(int grid, int smem, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                     const CUtensorMap& tmR, const PsParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_persist_kernel<MODE, NP, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  launch_k(conv_persist_kernel<MODE, NP, TR>, dim3(grid), dim3(kPsThreads), (size_t)smem, st, tmA, tmB, tmC, tmR, p);
  HG_LAUNCH_OK("conv_persist_kernel");
  count_launch();
  return HG_OK;
}

int conv_persist_launch(int N, int H, int W, int Kp, int Np, int mode, int ntaps, const signed char* dh,
                        const signed char* dw, const signed char* wt, const void* act, const void* wpk, int wtaps,
                        const float* bias, const void* res, void* out, float* stats, const BnFoldDev* fold,
                        cudaStream_t st) {
  PsParams p;
  memset(&p, 0, sizeof(p));
  const bool need_y = mode == kMask || res != nullptr;
  const int smem = ps_plan(W, Kp, Np, ntaps, need_y, p);
  if (smem <= 0) {
    set_error("conv_persist_launch: %d -> %d (%d taps) @%dx%d does not fit the persistent kernel", Kp, Np, ntaps, H, W);
    return HG_ERR_UNSUPPORTED;
  }
  if (mode == kMask && (!res || !stats)) {
    set_error("conv_persist_launch: mask mode needs the raw BatchNorm input and the reduction buffer");
    return HG_ERR_BAD_ARG;
  }
  if (mode == kPlainBnOut && (!fold || !fold->use_running || res || stats)) {
    set_error("conv_persist_launch: output BatchNorm needs running statistics and takes no residual / statistics");
    return HG_ERR_BAD_ARG;
  }
  const long long M = (long long)N * H * W;
  const int tile_px = p.tile_units * 128;
  CUtensorMap tmA, tmB, tmC, tmR;
  if (ntaps == 9) {
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Kp * 2, (uint64_t)W * Kp * 2, (uint64_t)H * W * Kp * 2};
    uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(tile_px / W + 2), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)Kp * 2};
    uint32_t box[2] = {64, (uint32_t)tile_px};
    uint32_t es[2] = {1, 1};
    int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, act, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)Np, (uint64_t)wtaps};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Np * Kp * 2};
    uint32_t box[3] = {64, (uint32_t)Np, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wpk, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Np, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)Np * 2};
    uint32_t box[2] = {64, 128};
    uint32_t es[2] = {1, 1};
    int rc = encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, res ? res : out, dims, str, box, es,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  p.M_total = (int)M;
  p.H = H;
  p.W = W;
  p.units = (int)(M / 128);
  p.upi = ntaps == 9 ? H * W / 128 : p.units;
  p.has_res = res != nullptr ? 1 : 0;
  p.bias = bias;
  p.stats = stats;
  if (fold) p.fold = *fold;
  p.ts = g_dbg_ts;
  p.dbg = g_ps_dbg;
  if (ntaps == 9) {
    for (int t = 0; t < 9; ++t) p.wt[dh[t] + 1][dw[t] + 1] = wt[t];
  } else {
    p.wt[0][0] = wt[0];
  }
  // one CTA per SM; every CTA gets at least one full tile
  int grid = (p.units + p.tile_units - 1) / p.tile_units;
  p.num_tiles = grid;
  if (grid > kNumSMs) grid = kNumSMs;
  // dynamic order needs whole tiles that never straddle an image (3x3): units per image a multiple of the tile
  if (g_persist_dynamic && p.num_tiles > grid && (ntaps == 1 || p.upi % p.tile_units == 0)) {
    if (!g_ps_ctr_pool) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cs);
      if (cs == cudaStreamCaptureStatusNone) {
        HG_CUDA_OK(cudaMalloc(&g_ps_ctr_pool, kPsCtrSlots * 2 * sizeof(int)));
        HG_CUDA_OK(cudaMemset(g_ps_ctr_pool, 0, kPsCtrSlots * 2 * sizeof(int)));
      }
    }
    if (g_ps_ctr_pool) {
      p.ctr = g_ps_ctr_pool + 2 * g_ps_ctr_next;
      g_ps_ctr_next = (g_ps_ctr_next + 1) % kPsCtrSlots;
    }
  }
  if (mode == kPlainBnOut) {
    if (Np == 128 && ntaps == 9 && g_persist_transposed) return ps_launch<kPlainBnOut, 128, true>(grid, smem, tmA, tmB, tmC, tmR, p, st);
    if (Np == 256) return ps_launch<kPlainBnOut, 256>(grid, smem, tmA, tmB, tmC, tmR, p, st);
    if (Np == 128) return ps_launch<kPlainBnOut, 128>(grid, smem, tmA, tmB, tmC, tmR, p, st);
    return ps_launch<kPlainBnOut, 64>(grid, smem, tmA, tmB, tmC, tmR, p, st);
  }
  if (Np == 256)
    return mode == kMask ? ps_launch<kMask, 256>(grid, smem, tmA, tmB, tmC, tmR, p, st)
                         : ps_launch<kPlain, 256>(grid, smem, tmA, tmB, tmC, tmR, p, st);
  if (Np == 128 && ntaps == 9 && g_persist_transposed)
    return mode == kMask ? ps_launch<kMask, 128, true>(grid, smem, tmA, tmB, tmC, tmR, p, st)
                         : ps_launch<kPlain, 128, true>(grid, smem, tmA, tmB, tmC, tmR, p, st);
  if (Np == 128)
    return mode == kMask ? ps_launch<kMask, 128>(grid, smem, tmA, tmB, tmC, tmR, p, st)
                         : ps_launch<kPlain, 128>(grid, smem, tmA, tmB, tmC, tmR, p, st);
  return mode == kMask ? ps_launch<kMask, 64>(grid, smem, tmA, tmB, tmC, tmR, p, st)
                       : ps_launch<kPlain, 64>(grid, smem, tmA, tmB, tmC, tmR, p, st);
}

}  // namespace hg
