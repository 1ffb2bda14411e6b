// Persistent 3x3 convolution (stride 1, dilation 1, pad 1) for the large maps, fprop and dgrad:
//   y[m, co] = sum_{dh, dw, ci} a[m + (dh, dw), ci] * w[tap(dh, dw)][co][ci]        dh, dw in {-1, 0, 1}
// (conv2 of every ResidualBlock, reference try_with_torch.py:189,202-204, and its data gradient.)
//
// The tile-per-CTA kernel of conv_tc.cu fetches 32 KB of operands (a 128-pixel activation tile + a 128-channel weight
// tile) per 2.1 MFLOP K block: at 64x64 it runs at the L2 -> SM operand rate (~45-60 B/clk/SM), not at the tensor
// core's.  This kernel halves the bytes per FLOP twice over:
//   * a CTA tile is 256 pixels (R = 256/W whole image rows) x Np channels: TWO 128-row accumulators share every
//     weight tile;
//   * ONE activation box of R+2 image rows (the tile plus one halo row above and below), shifted by dw, serves the
//     three taps dh = -1, 0, +1: the A operand of tap dh is the same shared-memory image read from row offset
//     (dh+1)*W -- whole image rows, so every shifted start stays 1024-byte aligned for the 128-byte swizzle.
//   Per (64-channel slice, dw): (R+2)*W*128 B of activations + 3 weight tiles for 6 MMA blocks -> 96 KB per
//   12.6 MFLOP at 64x64 (131 FLOP/B instead of 64).
// ONE CTA per SM stays resident and walks a contiguous range of 128-pixel units (neighbouring tiles share their halo
// rows in L2; the range is cut into 256-pixel tiles, plus a 128-pixel one where a range or an image ends on an odd
// unit), with two accumulator sets in TMEM (4 x Np columns): the eight epilogue warps drain tile i while the MMA
// warp is already issuing tile i+1.  Activations and weights run through two separate rings (a weight tile is
// consumed three times as often as an activation box).
// Epilogues: kPlain (+bias, +residual, shifted BatchNorm statistics of the output) and kMask (ReLU mask + the two
// BatchNorm-backward sums), exactly those of conv_gemm_kernel; the per-channel sums of the whole CTA stay in shared
// memory and leave as one vector atomic per 4 channels per CTA.
#include "hg_common.cuh"

// -DHG_DBG_TS=1 (make DBG=1): CTA 0 accumulates the cycles each role spends waiting on each barrier / in each epilogue
// phase into the dbg_ts buffer (hg_set_option("dbg_ts", 1), printed by ("dbg_ts", 3)); p3_dbg 1 = no epilogue work,
// 2 = no MMAs issued, 3 = both (what the TMA stream alone takes)
#ifndef HG_DBG_TS
#define HG_DBG_TS 0
#endif
#if HG_DBG_TS
#define P3_TIC long long _t0 = clock64()
#define P3_TOC(var) var += clock64() - _t0
#else
#define P3_TIC
#define P3_TOC(var)
#endif

namespace hg {

extern long long* g_dbg_ts;
int g_p3_dbg = 0;

int g_persist_3x3 = 0;   // measured: equal in isolation (51 vs 52 us @64x64), step 36.7 -> 37.6 ms (owns every SM: other lanes cannot interleave)
int g_persist3_min_units = 256;   // at least this many 128-pixel units (32x32 at batch 32)

struct P3Params {
  int M_total;        // N*H*W (a multiple of 128)
  int H, W;
  int units;          // M_total / 128
  int upi;            // 128-pixel units per image
  int kchunks;        // Kp / 64
  int nA, nB;         // ring depths
  int a_bytes;        // (R+2) * W * 128
  int has_res;        // kPlain: residual added;  kMask: tmR is the raw BatchNorm input (always)
  const float* bias;  // [Np] or null
  float* stats;       // kPlain: {S1, S2, pivot}[3*Np] or null;  kMask: {sum g, sum g*xhat}[2*Np]
  BnFoldDev fold;     // kMask: BatchNorm of the OUTPUT channels
  signed char wt[3][3];   // weight matrix of the tap whose A offset is (dh, dw) = (i-1, j-1)
  long long* ts;      // debug counters or null
  int dbg;
  int offB, offC, offY, offBar;
};

constexpr int kP3Threads = 320;   // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int kP3Epi = 256;

// 128-pixel units [u, u1) of this CTA, cut into tiles of two units (one where the range or the image ends)
struct P3Tiles {
  int u, u1, upi;
  __device__ __forceinline__ P3Tiles(const P3Params& p) {
    const int base = p.units / (int)gridDim.x, rem = p.units % (int)gridDim.x;
    const int b = (int)blockIdx.x;
    u = b * base + (b < rem ? b : rem);
    u1 = u + base + (b < rem ? 1 : 0);
    upi = p.upi;
  }
  __device__ __forceinline__ bool next(int& m0, int& mt) {
    if (u >= u1) return false;
    mt = (u1 - u >= 2 && (u % upi) != upi - 1) ? 2 : 1;
    m0 = u * 128;
    u += mt;
    return true;
  }
};

template <int MODE, int NP>
__global__ void __launch_bounds__(kP3Threads, 1)
conv3x3_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                       const P3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kBBytes = NP * 128;            // one weight tile: NP out-channels x 64 in-channels
  constexpr int kPanels = NP / 64;
  constexpr int kCBytes = kPanels * 16384;     // one 128-pixel sub-tile of the output
  uint8_t* sA = smem;                          // [nA][(R+2)*W rows x 128 B]
  uint8_t* sB = smem + p.offB;                 // [nB][NP rows x 128 B]
  uint8_t* sC = smem + p.offC;                 // [kPanels][128 rows x 128 B]: output staging (TMA store source)
  uint8_t* sY = smem + p.offY;                 // residual (kPlain) / raw BatchNorm input (kMask) of the sub-tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
  uint64_t* a_full = bars;                     // [4]
  uint64_t* a_empty = bars + 4;                // [4]
  uint64_t* b_full = bars + 8;                 // [8]
  uint64_t* b_empty = bars + 16;               // [8]
  uint64_t* tmem_full = bars + 24;             // [2]
  uint64_t* tmem_empty = bars + 26;            // [2]
  uint64_t* y_full = bars + 28;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  float* bias_s = reinterpret_cast<float*>(bars + 32);   // [128]
  float* coef_s = bias_s + 128;                            // kMask: scale / shift / A / B [4][128]
  float* acc_s = coef_s + 512;                             // [2][128] per-CTA column sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 4 * NP;       // two accumulator sets x two sub-tiles

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    prefetch_tmap(&tmR);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 1);
    }
    mbar_init(y_full, 1);
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
    if (lane == 0) {
      const int hw = p.H * p.W;
      P3Tiles tiles(p);
      int m0, mt;
      int ia = 0, ib = 0;
      long long w_ae = 0, w_be = 0;
      const long long tstart = clock64();
      while (tiles.next(m0, mt)) {
        const int n = m0 / hw;
        const int h0 = (m0 - n * hw) / p.W;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int j = 0; j < 3; ++j) {
            const int sa = ia % p.nA;
            {
              P3_TIC;
              mbar_wait(&a_empty[sa], ((ia / p.nA) & 1) ^ 1);
              P3_TOC(w_ae);
            }
            mbar_expect_tx(&a_full[sa], (uint32_t)p.a_bytes);
            tma_load_4d(sA + sa * p.a_bytes, &tmA, &a_full[sa], kc * 64, j - 1, h0 - 1, n);
            ++ia;
            for (int i = 0; i < 3; ++i) {
              const int sb = ib % p.nB;
              {
                P3_TIC;
                mbar_wait(&b_empty[sb], ((ib / p.nB) & 1) ^ 1);
                P3_TOC(w_be);
              }
              mbar_expect_tx(&b_full[sb], kBBytes);
              tma_load_3d(sB + sb * kBBytes, &tmB, &b_full[sb], kc * 64, 0, p.wt[i][j]);
              ++ib;
            }
          }
        }
      }
      if (HG_DBG_TS && p.ts && blockIdx.x == 0) {
        p.ts[0] = clock64() - tstart;
        p.ts[1] = w_ae;
        p.ts[2] = w_be;
      }
    }
    __syncwarp();
    pdl_trigger();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, NP, 0, 0);
    P3Tiles tiles(p);
    int m0, mt;
    int ia = 0, ib = 0;
    long long w_te = 0, w_af = 0, w_bf = 0;
    const long long tstart = clock64();
    for (int it = 0; tiles.next(m0, mt); ++it) {
      const int acc = it & 1;
      {
        P3_TIC;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        P3_TOC(w_te);
      }
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * 2 * NP;
      uint32_t accum = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        for (int j = 0; j < 3; ++j) {
          const int sa = ia % p.nA;
          {
            P3_TIC;
            mbar_wait(&a_full[sa], (ia / p.nA) & 1);
            P3_TOC(w_af);
          }
          ++ia;
          const uint32_t a_addr = smem_u32(sA + sa * p.a_bytes);
          for (int i = 0; i < 3; ++i) {
            const int sb = ib % p.nB;
            {
              P3_TIC;
              mbar_wait(&b_full[sb], (ib / p.nB) & 1);
              P3_TOC(w_bf);
            }
            ++ib;
            tc_fence_after();
            if (lane == 0) {
              const uint64_t bdesc = make_smem_desc(smem_u32(sB + sb * kBBytes), 16, 1024);
              for (int t = 0; t < ((HG_DBG_TS && (p.dbg & 2)) ? 0 : mt); ++t) {
                // tap row i: the same box, (i * W) pixel rows further down; sub-tile t: 128 pixel rows further
                const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)(i * p.W + t * 128) * 128u, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tacc + t * NP, adesc + 2 * k, bdesc + 2 * k, idesc, (accum | (uint32_t)k) ? 1u : 0u);
              }
              umma_commit(&b_empty[sb]);
              if (i == 2) umma_commit(&a_empty[sa]);
            }
            accum = 1;
            __syncwarp();
          }
        }
      }
      if (lane == 0) umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
    if (HG_DBG_TS && p.ts && blockIdx.x == 0 && lane == 0) {
      p.ts[4] = clock64() - tstart;
      p.ts[5] = w_te;
      p.ts[6] = w_af;
      p.ts[7] = w_bf;
    }
    pdl_trigger();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int et = threadIdx.x - 64;            // 0..255
    const int sub = warp & 3;                   // TMEM lane quarter of this warp
    const int chalf = (warp - 2) >> 2;          // which half of the sub-tile's columns this warp stages
    const int row = sub * 32 + lane;
    constexpr int kChunks = NP / 64;            // 32-column chunks per warp
    const bool need_y = MODE == kMask || p.has_res;
    for (int c = et; c < NP; c += kP3Epi) bias_s[c] = p.bias ? p.bias[c] : 0.f;
    for (int c = et; c < 2 * 128; c += kP3Epi) acc_s[c] = 0.f;
    if constexpr (MODE == kMask) {
      for (int c = et; c < NP; c += kP3Epi) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
        coef_s[c] = sc;                 // ReLU mask: scale * y + shift > 0 (the forward's own expression)
        coef_s[128 + c] = sh;
        coef_s[256 + c] = is;           // xhat = y * A + B  ->  sum g*xhat = A * sum(g*y) + B * sum(g)
        coef_s[384 + c] = -mu * is;
      }
    }
    if constexpr (MODE != kMask) {
      // statistics are sums of (y - pivot) (bn.cu): the pivots of the output channels
      if (p.stats != nullptr)
        for (int c = et; c < NP; c += kP3Epi) coef_s[c] = p.stats[2 * NP + c];
    }
    const bool relu = p.fold.relu != 0;
    float acc_sum[kChunks], acc_sq[kChunks];   // lane l: channel chalf * NP/2 + jj * 32 + l over this warp's rows
#pragma unroll
    for (int jj = 0; jj < kChunks; ++jj) acc_sum[jj] = acc_sq[jj] = 0.f;
    auto load_y = [&](int m) {
      mbar_expect_tx(y_full, kCBytes);
      for (int pnl = 0; pnl < kPanels; ++pnl) tma_load_2d(sY + pnl * 16384, &tmR, y_full, pnl * 64, m);
    };
    P3Tiles tiles(p);
    int m0, mt;
    bool have = tiles.next(m0, mt);
    if (need_y && have && et == 0) load_y(m0);
    named_bar_sync(1, kP3Epi);
    int g = 0;                                  // sub-tiles drained so far
    long long w_tf = 0, w_y = 0, t_row = 0, t_col = 0, t_sw = 0;
    const long long tstart = clock64();
    for (int it = 0; have; ++it) {
      const int acc = it & 1;
      int m0n = 0, mtn = 0;
      const bool have_next = tiles.next(m0n, mtn);
      {
        P3_TIC;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        P3_TOC(w_tf);
      }
      tc_fence_after();
      if (HG_DBG_TS && (p.dbg & 1)) {
        tc_fence_before();
        named_bar_sync(1, kP3Epi);
        if (et == 0) mbar_arrive(&tmem_empty[acc]);
        have = have_next;
        m0 = m0n;
        mt = mtn;
        continue;
      }
      for (int t = 0; t < mt; ++t, ++g) {
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * 2 * NP + t * NP + chalf * (NP / 2);
        float v[kChunks][32];
#pragma unroll
        for (int jj = 0; jj < kChunks; ++jj) tmem_ld32(taddr + jj * 32, v[jj]);
        {
          P3_TIC;
          if (need_y) mbar_wait(y_full, g & 1);
          tmem_ld_wait();
          P3_TOC(w_y);
        }
        P3_TIC;
        // ---- row pass: registers -> (+bias, +residual | mask) -> bf16 -> swizzled staging ----
#pragma unroll
        for (int jj = 0; jj < kChunks; ++jj) {
          float sq[32];
          const int col0 = chalf * (NP / 2) + jj * 32;          // first output channel of the chunk
          const int pnl = col0 / 64;
          const int chunk0 = (col0 % 64) / 8;
          const int rowoff = pnl * 16384 + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int swz = ((chunk0 + q) ^ (row & 7)) << 4;
            float o[8];
            float yv[8];   // kMask: raw BatchNorm input of these 8 channels
            if constexpr (MODE == kMask) {
              const uint4 u = *reinterpret_cast<const uint4*>(sY + rowoff + swz);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
              float cS[8], cT[8];
              load_coef8(coef_s + col0 + q * 8, cS);
              load_coef8(coef_s + 128 + col0 + q * 8, cT);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                yv[2 * e] = f.x;
                yv[2 * e + 1] = f.y;
                const bool k0 = !relu || fmaf(f.x, cS[2 * e], cT[2 * e]) > 0.f;
                const bool k1 = !relu || fmaf(f.y, cS[2 * e + 1], cT[2 * e + 1]) > 0.f;
                o[2 * e] = k0 ? v[jj][q * 8 + 2 * e] : 0.f;
                o[2 * e + 1] = k1 ? v[jj][q * 8 + 2 * e + 1] : 0.f;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = v[jj][q * 8 + e] + bias_s[col0 + q * 8 + e];
              if (p.has_res) {
                const uint4 u = *reinterpret_cast<const uint4*>(sY + rowoff + swz);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  o[2 * e] += f.x;
                  o[2 * e + 1] += f.y;
                }
              }
            }
            uint4 w;
            __nv_bfloat162* hw2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) hw2[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
            *reinterpret_cast<uint4*>(sC + rowoff + swz) = w;
            if (p.stats != nullptr) {
              // per-channel sums of the values just staged (bf16, exactly what the consumers read): the terms of this
              // row replace the accumulator values in v[] (kPlain: y - pivot and its square; kMask: g and g * y)
              float pv[8];
              if constexpr (MODE != kMask) load_coef8(coef_s + col0 + q * 8, pv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 r = __bfloat1622float2(hw2[e]);
                if constexpr (MODE == kMask) {
                  v[jj][q * 8 + 2 * e] = r.x;
                  v[jj][q * 8 + 2 * e + 1] = r.y;
                  sq[q * 8 + 2 * e] = r.x * yv[2 * e];
                  sq[q * 8 + 2 * e + 1] = r.y * yv[2 * e + 1];
                } else {
                  const float a = r.x - pv[2 * e], b = r.y - pv[2 * e + 1];
                  v[jj][q * 8 + 2 * e] = a;
                  v[jj][q * 8 + 2 * e + 1] = b;
                  sq[q * 8 + 2 * e] = a * a;
                  sq[q * 8 + 2 * e + 1] = b * b;
                }
              }
            }
          }
          if (p.stats != nullptr) {
            // sum over the 32 rows of this warp: afterwards lane l holds the totals of channel col0 + l
            acc_sum[jj] += warp_transpose_sum32(v[jj], lane);
            acc_sq[jj] += warp_transpose_sum32(sq, lane);
          }
        }
        // every TMEM read of this accumulator set is done after its last sub-tile: hand it back to the MMA warp
        if (t == mt - 1) tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1, kP3Epi);
        P3_TOC(t_row);
        if (et == 0) {
          if (t == mt - 1) mbar_arrive(&tmem_empty[acc]);
          for (int pnl = 0; pnl < kPanels; ++pnl) tma_store_2d(&tmC, sC + pnl * 16384, pnl * 64, m0 + t * 128);
          tma_store_commit();
        }
#if HG_DBG_TS
        const long long _t1 = clock64();
#endif
        // the TMA store must have read the staging tile before the next row pass rewrites it
#if HG_DBG_TS
        const long long _t2 = clock64();
        t_col += _t2 - _t1;
#endif
        if (et == 0) tma_store_wait_read();
        named_bar_sync(1, kP3Epi);
#if HG_DBG_TS
        t_sw += clock64() - _t2;
#endif
        if (need_y && et == 0) {
          if (t + 1 < mt) load_y(m0 + (t + 1) * 128);
          else if (have_next) load_y(m0n);
        }
      }
      have = have_next;
      m0 = m0n;
      mt = mtn;
    }
    if (p.stats != nullptr) {
      // the four row-quarter warps of a column half add up in shared memory, then
      // one vector atomic per 4 channels per CTA for the whole kernel
#pragma unroll
      for (int jj = 0; jj < kChunks; ++jj) {
        const int c = chalf * (NP / 2) + jj * 32 + lane;
        atomicAdd(acc_s + c, acc_sum[jj]);
        atomicAdd(acc_s + 128 + c, acc_sq[jj]);
      }
      named_bar_sync(1, kP3Epi);
      for (int q = et; q < 2 * (NP / 4); q += kP3Epi) {
        const int which = q / (NP / 4), quad = q % (NP / 4);
        float4 v4 = *reinterpret_cast<const float4*>(acc_s + which * 128 + quad * 4);
        if (MODE == kMask && which == 1) {
          const float4 sg = *reinterpret_cast<const float4*>(acc_s + quad * 4);
          const float4 cA = *reinterpret_cast<const float4*>(coef_s + 256 + quad * 4);
          const float4 cB = *reinterpret_cast<const float4*>(coef_s + 384 + quad * 4);
          v4 = make_float4(fmaf(cA.x, v4.x, cB.x * sg.x), fmaf(cA.y, v4.y, cB.y * sg.y),
                           fmaf(cA.z, v4.z, cB.z * sg.z), fmaf(cA.w, v4.w, cB.w * sg.w));
        }
        float* dst = p.stats + which * NP + quad * 4;
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
}

// Shared-memory plan; returns the dynamic shared-memory size or 0 when the shape does not fit.
static int p3_plan(int W, int Np, bool need_y, P3Params& p) {
  const int R = 256 / W;
  const int misc = 4096;   // barriers (256 B) + bias + coefficients + column sums
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - misc;
  p.a_bytes = (R + 2) * W * 128;
  const int bbytes = Np * 128;
  const int cbytes = (Np / 64) * 16384;
  const int ybytes = need_y ? cbytes : 0;
  const int ring = budget - cbytes - ybytes;
  // two activation boxes are the minimum (one in flight while one is read); a third one is taken only when at least
  // three weight tiles (one box's worth) still fit next to it
  int nA = 3;
  if (ring - 3 * p.a_bytes < 3 * bbytes) nA = 2;
  int nB = (ring - nA * p.a_bytes) / bbytes;
  if (nB > 8) nB = 8;
  if (nB < 3) return 0;
  p.nA = nA;
  p.nB = nB;
  p.offB = nA * p.a_bytes;
  p.offC = p.offB + nB * bbytes;
  p.offY = p.offC + cbytes;
  p.offBar = p.offY + ybytes;
  return p.offBar + misc + 1024;
}

// g_* : the taps of the launch as (dh, dw, weight index); eligible = the nine offsets {-1,0,1}^2, each once
bool conv_p3_eligible(int N, int H, int W, int Kp, int Np, int ntaps, const signed char* dh, const signed char* dw,
                      int stride, int parity, int mode, const float* out_nchw, bool has_res) {
  if (!g_persist_3x3 || ntaps != 9 || stride != 1 || parity || out_nchw != nullptr) return false;
  if (mode != kPlain && mode != kMask) return false;
  if (!(Np == 64 || Np == 128) || Kp % 64 || Kp > 256) return false;
  if (!is_pow2(W) || !is_pow2(H) || W > 128 || W < 16 || 256 / W > H) return false;
  const long long M = (long long)N * H * W;
  if (M / 128 < g_persist3_min_units) return false;
  unsigned seen = 0;
  for (int t = 0; t < 9; ++t) {
    if (dh[t] < -1 || dh[t] > 1 || dw[t] < -1 || dw[t] > 1) return false;
    seen |= 1u << ((dh[t] + 1) * 3 + dw[t] + 1);
  }
  if (seen != 0x1FFu) return false;
  P3Params p;
  return p3_plan(W, Np, mode == kMask || has_res, p) > 0;
}

int conv_p3_launch(int N, int H, int W, int Kp, int Np, int mode, const signed char* dh, const signed char* dw,
                   const signed char* wt, const void* act, const void* wpk, int wtaps, const float* bias,
                   const void* res, void* out, float* stats, const BnFoldDev* fold, cudaStream_t st) {
  P3Params p;
  memset(&p, 0, sizeof(p));
  const bool need_y = mode == kMask || res != nullptr;
  const int smem = p3_plan(W, Np, need_y, p);
  if (smem <= 0) {
    set_error("conv_p3_launch: %d -> %d @%dx%d does not fit the persistent kernel", Kp, Np, H, W);
    return HG_ERR_UNSUPPORTED;
  }
  const long long M = (long long)N * H * W;
  const int R = 256 / W;
  CUtensorMap tmA, tmB, tmC, tmR;
  {
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Kp * 2, (uint64_t)W * Kp * 2, (uint64_t)H * W * Kp * 2};
    uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(R + 2), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
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
  p.upi = H * W / 128;
  p.kchunks = Kp / 64;
  p.has_res = res != nullptr ? 1 : 0;
  p.bias = bias;
  p.stats = stats;
  if (fold) p.fold = *fold;
  p.ts = g_dbg_ts;
  p.dbg = g_p3_dbg;
  for (int t = 0; t < 9; ++t) p.wt[dh[t] + 1][dw[t] + 1] = wt[t];
  if (mode == kMask && (!res || !stats)) {
    set_error("conv_p3_launch: mask mode needs the raw BatchNorm input and the reduction buffer");
    return HG_ERR_BAD_ARG;
  }
  static bool attr_set = false;
  if (!attr_set) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_persist_kernel<kPlain, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_persist_kernel<kMask, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_persist_kernel<kPlain, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_persist_kernel<kMask, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // one CTA per SM; every CTA gets at least one 256-pixel tile
  int grid = (p.units + 1) / 2;
  if (grid > kNumSMs) grid = kNumSMs;
  const dim3 g(grid), b(kP3Threads);
  if (Np == 128) {
    if (mode == kMask) launch_k(conv3x3_persist_kernel<kMask, 128>, g, b, (size_t)smem, st, tmA, tmB, tmC, tmR, p);
    else launch_k(conv3x3_persist_kernel<kPlain, 128>, g, b, (size_t)smem, st, tmA, tmB, tmC, tmR, p);
  } else {
    if (mode == kMask) launch_k(conv3x3_persist_kernel<kMask, 64>, g, b, (size_t)smem, st, tmA, tmB, tmC, tmR, p);
    else launch_k(conv3x3_persist_kernel<kPlain, 64>, g, b, (size_t)smem, st, tmA, tmB, tmC, tmR, p);
  }
  HG_LAUNCH_OK("conv3x3_persist_kernel");
  count_launch();
  return HG_OK;
}

}  // namespace hg
