// Persistent 1x1 convolution (pointwise GEMM) for the large maps: y[m, co] = sum_ci a[m, ci] * w[co][ci].
//
// The 1x1 convolutions of the residual blocks (reference try_with_torch.py:186,192,193; lin :248; conv3/conv4 :272-273)
// at 64x64 / 128x128 are HBM-bound, and the tile-per-CTA kernel of conv_tc.cu leaves half of the bandwidth unused: every
// CTA re-fetches the weights (as many bytes as its activation tile), pays barrier / TMEM / descriptor set-up per 128
// pixels, and its epilogue only overlaps other CTAs' loads by luck of co-residency.  Here ONE CTA per SM stays resident:
//   * the whole weight matrix [Np x Kp] (<= 128 KB) is loaded once and stays in shared memory;
//   * a producer warp streams 128-pixel activation tiles (64-channel chunks) through a ring, running ahead across tiles;
//   * the MMA warp issues tcgen05.mma with N = Np (up to 256) into one of TWO TMEM accumulators;
//   * eight epilogue warps drain the other accumulator meanwhile (bias / residual / ReLU-mask epilogues of conv_tc.cu),
//     stage 128-column groups in shared memory for the TMA store, and keep the per-channel statistics of the whole CTA
//     in shared memory: one vector atomic per 4 channels per CTA at the end instead of one per tile;
//   * the residual / raw-BatchNorm-input rows are prefetched into REGISTERS one column group ahead (each epilogue
//     thread owns one pixel row: 64 contiguous bytes per 32-column chunk), so their latency hides behind the previous
//     group's store and column pass instead of sitting in front of every group (first version, TMA into the staging
//     buffer with a two-buffer hand-shake: 64 us for 128->256 + residual @64x64; the tile-per-CTA kernel takes 47).
// Modes: kPlain (fprop / dgrad: + bias, + residual, + BatchNorm statistics of the output) and kMask (dgrad whose result
// feeds a BatchNorm backward: ReLU mask + the two BatchNorm-backward sums), stride 1, Kp, Np in {64, 128, 256}.
#include "hg_common.cuh"

namespace hg {

// Measured on B200 (batch 32, 64x64; tools/gpu_top_kernels.py, us per launch, tile-per-CTA kernel -> this kernel):
//   fprop 256->128 29.1 -> 26.5;  fprop 128->256 + residual 47.1 -> 63.9 (residual by TMA) / 71.5 (residual rows by
//   per-thread loads);  masked dgrad 128->256 36.4 -> 34.1 / 42.5;  masked dgrad 256->128 46.6 -> 53.8 / 72.7;  whole
//   training step 882 -> 851 / 830 images/s.  The epilogue of a 256-column tile (two staged column groups, three named
//   barriers each) is as long as the tile's HBM time, and 148 resident CTAs that own every SM for the whole launch stop
//   the other stream lanes from interleaving.  So the kernel is OFF by default (hg_set_option("persist_1x1", 1) turns it
//   on); it stays in the library, parity-tested (tests/test_gpu_ops.py::test_persistent_pointwise_kernel*).
int g_persist_1x1 = 0;
int g_persist_min_tiles = 296;    // used for at least this many 128-pixel tiles (2 per SM)

struct P1Params {
  int M_total;
  int num_tiles;
  int kchunks;        // Kp / 64
  int Np;             // padded output channels (MMA N)
  int nst;            // activation ring stages (16 KB each)
  int groups;         // 128-column groups per tile (Np / 128, at least 1)
  int gpanels;        // 64-column panels per group (1 or 2)
  int has_res;        // kPlain: residual added;  kMask: tmR is the raw BatchNorm input (always)
  const float* bias;  // [Np] or null
  const void* res;    // [M][Np] bf16: residual (kPlain, may alias the output) / raw BatchNorm input (kMask), or null
  float* stats;       // [2*Np] or null
  BnFoldDev fold;     // kMask: BatchNorm of the OUTPUT channels
  // shared-memory offsets (bytes from the 1024-aligned base)
  int offA, offG, offY, offBar;
};

constexpr int kP1Threads = 320;   // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr int kP1Epi = 256;
constexpr int kCBufs = 2;

template <int MODE>
__global__ void __launch_bounds__(kP1Threads, 1)
conv1x1_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                       const P1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                       // [kchunks][Np rows x 128 B]
  uint8_t* sA = smem + p.offA;              // [nst][128 rows x 128 B]
  uint8_t* sG = smem + p.offG;              // [kCBufs][gpanels][128 rows x 128 B]: output staging (TMA store source)
  uint8_t* sY = smem + p.offY;              // kMask: [gpanels][128 rows x 128 B]: raw BatchNorm input of the group
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
  uint64_t* full_bar = bars;                // [8]  activation chunk landed
  uint64_t* empty_bar = bars + 8;           // [8]  MMAs done with the slot
  uint64_t* tmem_full = bars + 16;          // [2]  accumulator ready
  uint64_t* tmem_empty = bars + 18;         // [2]  accumulator drained by the epilogue
  uint64_t* b_full = bars + 24;             //      weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  float* bias_s = reinterpret_cast<float*>(bars + 32);   // [256]
  float* coef_s = bias_s + 256;                            // kMask: scale / shift / A / B [4][256]
  float* acc_s = coef_s + 1024;                            // [2][256] per-CTA column sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Np = p.Np;
  const int cbuf_bytes = p.gpanels * 16384;
  const uint32_t tmem_cols = 2 * Np < 32 ? 32 : 2 * Np;   // Np in {64,128,256}: 128 / 256 / 512 columns

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 1);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(b_full, (uint32_t)(p.kchunks * Np * 128));
      for (int kc = 0; kc < p.kchunks; ++kc) tma_load_3d(sB + kc * Np * 128, &tmB, b_full, kc * 64, 0, 0);
      int kb = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
        for (int kc = 0; kc < p.kchunks; ++kc, ++kb) {
          const int st = kb % p.nst;
          mbar_wait(&empty_bar[st], ((kb / p.nst) & 1) ^ 1);
          mbar_expect_tx(&full_bar[st], 16384);
          tma_load_2d(sA + st * 16384, &tmA, &full_bar[st], kc * 64, m0);
        }
      }
    }
    __syncwarp();
    pdl_trigger();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, Np, 0, 0);
    mbar_wait(b_full, 0);
    int kb = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      mbar_wait(&tmem_empty[acc], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < p.kchunks; ++kc, ++kb) {
        const int st = kb % p.nst;
        mbar_wait(&full_bar[st], (kb / p.nst) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_smem_desc(smem_u32(sA + st * 16384), 16, 1024);
          const uint64_t bdesc = make_smem_desc(smem_u32(sB + kc * Np * 128), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + acc * Np, adesc + 2 * k, bdesc + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[st]);
          if (kc == p.kchunks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
    pdl_trigger();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int et = threadIdx.x - 64;            // 0..255
    const int sub = warp & 3;                   // TMEM lane quarter of this warp
    const int chalf = (warp - 2) >> 2;          // which half of a group's 32-column chunks this warp stages
    const int row = sub * 32 + lane;
    const int gcols = p.gpanels * 64;           // columns per group (64 or 128)
    const int chunks_per_warp = gcols / 64;     // 32-column chunks per warp per group (1 or 2)
    for (int c = et; c < Np; c += kP1Epi) bias_s[c] = p.bias ? p.bias[c] : 0.f;
    for (int c = et; c < 2 * 256; c += kP1Epi) acc_s[c] = 0.f;
    if constexpr (MODE == kMask) {
      for (int c = et; c < Np; c += kP1Epi) {
        float mu, is, sc, sh;
        bn_fold_coeffs(p.fold, c, mu, is, sc, sh);
        coef_s[c] = sc;                 // ReLU mask: scale * y + shift > 0 (the forward's own expression)
        coef_s[256 + c] = sh;
        coef_s[512 + c] = is;           // xhat = y * A + B  ->  sum g*xhat = A * sum(g*y) + B * sum(g)
        coef_s[768 + c] = -mu * is;
      }
    }
    named_bar_sync(1, kP1Epi);
    const bool relu = p.fold.relu != 0;
    // residual / raw-input chunk(s) of the NEXT column group of this thread's row, kept unconverted
    uint4 rr[2][4];
    auto prefetch_rows = [&](int it, int grp) {
      const int m = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + row;
      const bool ok = it < my_tiles && m < p.M_total;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj < chunks_per_warp) {
          const int col0 = grp * 128 + (chalf * chunks_per_warp + jj) * 32;
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) +
                                                            (size_t)m * Np + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q) rr[jj][q] = ok ? src[q] : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    if (p.has_res) prefetch_rows(0, 0);
    int g = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
      int valid = p.M_total - m0;
      valid = valid > 128 ? 128 : valid;
      mbar_wait(&tmem_full[acc], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * Np;
      for (int grp = 0; grp < p.groups; ++grp, ++g) {
        // staging buffer g % 2: the store issued from it two groups ago has read it (thread 0 waited for that before the
        // previous group's closing barrier)
        uint8_t* gbuf = sG + (g % kCBufs) * cbuf_bytes;
        // ---- row pass: TMEM -> registers -> (+bias, +residual | mask) -> bf16 -> swizzled staging ----
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (jj < chunks_per_warp) {
            const int j = chalf * chunks_per_warp + jj;          // 32-column chunk inside the group
            const int col0 = grp * 128 + j * 32;                 // first output channel of the chunk
            float v[32];
            tmem_ld32(taddr + col0, v);
            tmem_ld_wait();
            const int pnl = (j * 32) / 64;
            const int chunk0 = ((j * 32) % 64) / 8;
            const int rowoff = pnl * 16384 + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int swz = ((chunk0 + q) ^ (row & 7)) << 4;
              float o[8];
              if constexpr (MODE == kMask) {
                const uint4 u = rr[jj][q];
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
                const float* cS = coef_s + col0 + q * 8;
                const float* cT = coef_s + 256 + col0 + q * 8;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  const bool k0 = !relu || fmaf(f.x, cS[2 * e], cT[2 * e]) > 0.f;
                  const bool k1 = !relu || fmaf(f.y, cS[2 * e + 1], cT[2 * e + 1]) > 0.f;
                  o[2 * e] = k0 ? v[q * 8 + 2 * e] : 0.f;
                  o[2 * e + 1] = k1 ? v[q * 8 + 2 * e + 1] : 0.f;
                }
                *reinterpret_cast<uint4*>(sY + rowoff + swz) = u;   // the column pass needs sum g * y
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = v[q * 8 + e] + bias_s[col0 + q * 8 + e];
                if (p.has_res) {
                  const uint4 u = rr[jj][q];
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
              *reinterpret_cast<uint4*>(gbuf + rowoff + swz) = w;
            }
          }
        }
        // next group's rows: requested now, consumed after this group's store / column pass (and the next accumulator wait)
        if (p.has_res) {
          if (grp + 1 < p.groups) prefetch_rows(i, grp + 1);
          else prefetch_rows(i + 1, 0);
        }
        if (grp == p.groups - 1) {
          // every TMEM read of this accumulator is done (tcgen05.wait::ld above): hand it back to the MMA warp
          tc_fence_before();
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kP1Epi);
        if (et == 0) {
          if (grp == p.groups - 1) mbar_arrive(&tmem_empty[acc]);
          for (int pnl = 0; pnl < p.gpanels; ++pnl) tma_store_2d(&tmC, gbuf + pnl * 16384, grp * 128 + pnl * 64, m0);
          tma_store_commit();
        }
        // ---- column pass: per-channel sums of what was just staged (bf16, exactly what the consumers read) ----
        if (p.stats != nullptr) {
          const int kQuads = gcols / 4;                  // 16 or 32 column quads
          const int kGroups = kP1Epi / kQuads;           // 16 or 8 row slices
          const int kRows = 128 / kGroups;               // 8 or 16 rows each
          const int quad = et % kQuads, rg = et / kQuads;
          const int c = quad * 4;
          const int coff = (c >> 6) * 16384 + (c & 7) * 2;
          const int chunk = (c & 63) >> 3;
          const uint8_t* vcol = gbuf + coff;
          const uint8_t* ycol = (MODE == kMask ? sY : gbuf) + coff;
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
          float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);   // statistics are sums of (y - pivot) (bn.cu)
          if constexpr (MODE != kMask) pv = *reinterpret_cast<const float4*>(p.stats + 2 * Np + grp * 128 + c);
          for (int k = 0; k < kRows; ++k) {
            const int r = rg * kRows + k;
            const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
            const uint2 u = *reinterpret_cast<const uint2*>(vcol + off);
            float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            if constexpr (MODE != kMask) {
              f0.x -= pv.x; f0.y -= pv.y; f1.x -= pv.z; f1.y -= pv.w;
            }
            float2 y0 = f0, y1 = f1;
            if constexpr (MODE == kMask) {
              const uint2 uy = *reinterpret_cast<const uint2*>(ycol + off);
              y0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uy.x));
              y1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uy.y));
            }
            if (r < valid) {
              s4[0] += f0.x; s4[1] += f0.y; s4[2] += f1.x; s4[3] += f1.y;
              q4[0] = fmaf(f0.x, y0.x, q4[0]); q4[1] = fmaf(f0.y, y0.y, q4[1]);
              q4[2] = fmaf(f1.x, y1.x, q4[2]); q4[3] = fmaf(f1.y, y1.y, q4[3]);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            atomicAdd(acc_s + grp * 128 + c + e, s4[e]);
            atomicAdd(acc_s + 256 + grp * 128 + c + e, q4[e]);
          }
        }
        // the store just issued may stay in flight; the one issued from the OTHER staging buffer must have read it
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1, kP1Epi);   // column pass done with gbuf / sY; thread 0's wait is visible to everybody
      }
    }
    if (p.stats != nullptr) {
      // one vector atomic per 4 channels per CTA for the whole kernel
      for (int q = et; q < 2 * (Np / 4); q += kP1Epi) {
        const int which = q / (Np / 4), quad = q % (Np / 4);
        float4 v4 = *reinterpret_cast<const float4*>(acc_s + which * 256 + quad * 4);
        if (MODE == kMask && which == 1) {
          const float4 sg = *reinterpret_cast<const float4*>(acc_s + quad * 4);
          const float4 cA = *reinterpret_cast<const float4*>(coef_s + 512 + quad * 4);
          const float4 cB = *reinterpret_cast<const float4*>(coef_s + 768 + quad * 4);
          v4 = make_float4(fmaf(cA.x, v4.x, cB.x * sg.x), fmaf(cA.y, v4.y, cB.y * sg.y),
                           fmaf(cA.z, v4.z, cB.z * sg.z), fmaf(cA.w, v4.w, cB.w * sg.w));
        }
        float* dst = p.stats + which * Np + quad * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v4.x), "f"(v4.y), "f"(v4.z),
                     "f"(v4.w)
                     : "memory");
      }
    }
    if (et == 0) tma_store_wait_all();
    pdl_trigger();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// Shared-memory plan for (Kp, Np, mode); returns the dynamic shared-memory size or 0 when the shape does not fit.
static int p1_plan(int Kp, int Np, int mode, P1Params& p) {
  const int budget = 227 * 1024 - 1024 /*alignment slack*/;
  const int misc = 8192;   // barriers + bias + coefficients + column sums (256 + 4*(256 + 1024 + 512) B)
  p.kchunks = Kp / 64;
  p.Np = Np;
  p.groups = Np > 128 ? Np / 128 : 1;
  p.gpanels = Np >= 128 ? 2 : 1;
  const int bbytes = Kp * Np * 2;
  const int gbytes = kCBufs * p.gpanels * 16384;
  const int ybytes = mode == kMask ? p.gpanels * 16384 : 0;
  int nst = (budget - misc - bbytes - gbytes - ybytes) / 16384;
  if (nst > 8) nst = 8;
  if (nst < 2) return 0;
  p.nst = nst;
  p.offA = bbytes;
  p.offG = p.offA + nst * 16384;
  p.offY = p.offG + gbytes;
  p.offBar = p.offY + ybytes;
  return p.offBar + misc + 1024;
}

bool conv_p1_eligible(long long M, int Kp, int Np, int ntaps, int stride, int parity, int mode, const float* out_nchw) {
  if (!g_persist_1x1 || ntaps != 1 || stride != 1 || parity || out_nchw != nullptr) return false;
  if (mode != kPlain && mode != kMask) return false;
  if (!(Kp == 64 || Kp == 128 || Kp == 256) || !(Np == 64 || Np == 128 || Np == 256)) return false;
  if ((M + 127) / 128 < g_persist_min_tiles) return false;
  P1Params p;
  return p1_plan(Kp, Np, mode, p) > 0;
}

// tmA: {Kp, M} box {64, 128};  tmB: {Kp, Np, 1} box {64, Np, 1};  tmC / tmR: {Np, M} box {64, 128}
int conv_p1_launch(long long M, int Kp, int Np, int mode, const CUtensorMap& tmA, const CUtensorMap& tmB,
                   const CUtensorMap& tmC, const CUtensorMap& tmR, const float* bias, float* stats, const void* res,
                   const BnFoldDev* fold, cudaStream_t st) {
  P1Params p;
  memset(&p, 0, sizeof(p));
  const int smem = p1_plan(Kp, Np, mode, p);
  if (smem <= 0) {
    set_error("conv_p1_launch: shape %d -> %d does not fit the persistent kernel", Kp, Np);
    return HG_ERR_UNSUPPORTED;
  }
  p.M_total = (int)M;
  p.num_tiles = (int)((M + 127) / 128);
  p.has_res = res != nullptr ? 1 : 0;
  p.res = res;
  p.bias = bias;
  p.stats = stats;
  if (fold) p.fold = *fold;
  static bool attr_set = false;
  if (!attr_set) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv1x1_persist_kernel<kPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    HG_CUDA_OK(cudaFuncSetAttribute(conv1x1_persist_kernel<kMask>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    attr_set = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  if (mode == kMask)
    launch_k(conv1x1_persist_kernel<kMask>, dim3(grid), dim3(kP1Threads), (size_t)smem, st, tmA, tmB, tmC, tmR, p);
  else
    launch_k(conv1x1_persist_kernel<kPlain>, dim3(grid), dim3(kP1Threads), (size_t)smem, st, tmA, tmB, tmC, tmR, p);
  HG_LAUNCH_OK("conv1x1_persist_kernel");
  count_launch();
  return HG_OK;
}

}  // namespace hg
