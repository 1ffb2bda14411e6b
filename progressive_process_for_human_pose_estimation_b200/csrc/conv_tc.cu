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

namespace hg {

struct ConvGemmParams {
  int M_total;        // output pixels = N*H*W
  int H, W;           // spatial size (stride-1 "same" convolutions: in == out)
  int taps_r, taps_s; // filter size
  int dil, pad;       // dilation, padding
  int sign;           // +1 fprop (offset = r*dil - pad), -1 dgrad (offset = pad - r*dil)
  int kchunks;        // padded Cin / 64
  int n_total;        // padded Cout (row pitch of bias / stats)
  int c_real;         // real Cout (for the NCHW fp32 side output)
  const float* bias;  // [n_total] or null
  float* stats;       // [2*n_total] or null
  float* out_nchw;    // optional fp32 NCHW copy of the first c_real channels (heatmap heads)
  int has_res;
  int n_tiles;        // padded Cout / BN
};

template <int BN, int STAGES>
struct ConvGemmSmem {
  static constexpr int kABytes = 128 * 128;      // 128 pixels x 64 ch x 2 B
  static constexpr int kBBytes = BN * 128;       // BN out-channels x 64 ch x 2 B
  static constexpr int kCPanels = BN / 64;       // output staging: panels of 128 rows x 64 ch
  static constexpr int kCBytes = kCPanels * 128 * 128;
  static constexpr int kBarBytes = 2048;
  static constexpr int kTotal = STAGES * (kABytes + kBBytes) + kCBytes + kBarBytes + 1024 /*align slack*/;
};

template <int BN, int STAGES, int MINB>
__global__ void __launch_bounds__(192, MINB)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 const ConvGemmParams p) {
  using L = ConvGemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * L::kABytes;
  uint8_t* sC = sB + STAGES * L::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + L::kCBytes);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // accumulator ready
  uint64_t* res_full = bars + 2 * STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
  float* bias_s = reinterpret_cast<float*>(bars + 2 * STAGES + 4);  // [BN] (<= 256 floats = 1 KB - 64 B)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1-D grid, N tile fastest: the CTAs that share an activation tile run back to back (second read hits L2)
  const int m0 = (blockIdx.x / p.n_tiles) * 128;
  const int n_off = (blockIdx.x % p.n_tiles) * BN;
  const int num_kb = p.taps_r * p.taps_s * p.kchunks;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  if (warp >= 2) {
    for (int c = threadIdx.x - 64; c < BN; c += 128) bias_s[c] = p.bias ? p.bias[n_off + c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int hw = p.H * p.W;
      const int n0 = m0 / hw;
      const int rem = m0 - n0 * hw;
      const int h0 = rem / p.W;
      const int w0 = rem - h0 * p.W;
      if (p.has_res) {
        mbar_expect_tx(res_full, L::kCBytes);
        for (int pnl = 0; pnl < L::kCPanels; ++pnl)
          tma_load_2d(sC + pnl * 16384, &tmR, res_full, n_off + pnl * 64, m0);
      }
      int kb = 0;
      for (int r = 0; r < p.taps_r; ++r) {
        for (int s = 0; s < p.taps_s; ++s) {
          const int dh = p.sign * (r * p.dil - p.pad);
          const int dw = p.sign * (s * p.dil - p.pad);
          for (int kc = 0; kc < p.kchunks; ++kc, ++kb) {
            const int st = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty_bar[st], ph ^ 1);
            mbar_expect_tx(&full_bar[st], L::kABytes + L::kBBytes);
            tma_load_4d(sA + st * L::kABytes, &tmA, &full_bar[st], kc * 64, w0 + dw, h0 + dh, n0);
            tma_load_3d(sB + st * L::kBBytes, &tmB, &full_bar[st], kc * 64, n_off, r * p.taps_s + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc(smem_u32(sA + st * L::kABytes), 16, 1024);
        const uint64_t bdesc = make_smem_desc(smem_u32(sB + st * L::kBBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 K-elements = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);  // frees the smem slot once these MMAs have read it
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int sub = warp & 3;           // TMEM sub-partition this warp may read
    const int row = sub * 32 + lane;    // accumulator row == pixel within the tile
    const int et = threadIdx.x - 64;    // 0..127
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (p.has_res) mbar_wait(res_full, 0);
    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
    const int m = m0 + row;
    const bool row_ok = m < p.M_total;
    float* nchw_row = nullptr;
    int plane = 0;
    if (p.out_nchw != nullptr && row_ok) {
      plane = p.H * p.W;
      const int n = m / plane;
      nchw_row = p.out_nchw + (size_t)n * p.c_real * plane + (m - n * plane);
    }
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      float v[32];
      tmem_ld32(taddr + j * 32, v);
      tmem_ld_wait();
      const int pnl = (j * 32) / 64;
      const int chunk0 = ((j * 32) % 64) / 8;
      uint8_t* rowp = sC + pnl * 16384 + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4* cp = reinterpret_cast<uint4*>(rowp + (((chunk0 + q) ^ (row & 7)) << 4));
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = v[q * 8 + e] + bias_s[j * 32 + q * 8 + e];
        if (p.has_res) {
          uint4 u = *cp;
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h[e]);
            o[2 * e] += f.x;
            o[2 * e + 1] += f.y;
          }
        }
        if (nchw_row != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = n_off + j * 32 + q * 8 + e;
            if (c < p.c_real) nchw_row[(size_t)c * plane] = o[e];
          }
        }
        uint4 w;
        __nv_bfloat162* hw2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) hw2[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        *cp = w;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (et == 0) {
      for (int pnl = 0; pnl < L::kCPanels; ++pnl) tma_store_2d(&tmC, sC + pnl * 16384, n_off + pnl * 64, m0);
      tma_store_commit();
    }
    if (p.stats != nullptr) {
      // per-channel sum / sum of squares of the bf16 values just staged: thread = 2 adjacent channels x a slice
      // of the 128 rows (fixed trip count so the loop unrolls and the LDS latencies overlap)
      int valid = p.M_total - m0;
      valid = valid > 128 ? 128 : valid;
      constexpr int kPairs = BN / 2;
      constexpr int kGroups = 128 / kPairs > 0 ? 128 / kPairs : 1;
      constexpr int kRows = 128 / kGroups;
      for (int pi = et; pi < kPairs * kGroups; pi += 128) {
        const int cpair = pi % kPairs, grp = pi / kPairs;
        const int c = cpair * 2;
        const uint8_t* colp = sC + (c >> 6) * 16384 + (c & 7) * 2;
        const int chunk = (c & 63) >> 3;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int i = 0; i < kRows; ++i) {
          const int r = grp * kRows + i;
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(colp + r * 128 + ((chunk ^ (r & 7)) << 4));
          float2 f = __bfloat1622float2(h);
          if (r >= valid) f = make_float2(0.f, 0.f);
          s0 += f.x;
          s1 += f.y;
          q0 = fmaf(f.x, f.x, q0);
          q1 = fmaf(f.y, f.y, q1);
        }
        atomicAdd(p.stats + n_off + c, s0);
        atomicAdd(p.stats + n_off + c + 1, s1);
        atomicAdd(p.stats + p.n_total + n_off + c, q0);
        atomicAdd(p.stats + p.n_total + n_off + c + 1, q1);
      }
    }
    if (et == 0) tma_store_wait_read();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
static inline int pad64(int c) { return (c + 63) & ~63; }

template <int BN, int STAGES, int MINB>
static int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                            const CUtensorMap& tmR, const ConvGemmParams& p, cudaStream_t st) {
  using L = ConvGemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    HG_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::kTotal));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.M_total, 128) * p.n_tiles);
  conv_gemm_kernel<BN, STAGES, MINB><<<grid, 192, L::kTotal, st>>>(tmA, tmB, tmC, tmR, p);
  HG_LAUNCH_OK("conv_gemm_kernel");
  count_launch();
  return HG_OK;
}

// act: [N,H,W,Kp] bf16 (A operand); wpk: [taps][Np][Kp] bf16; out/res: [N,H,W,Np] bf16.
int conv_gemm_bf16(int N, int H, int W, int Kp, int Np, int R, int S, int pad, int dil, int sign,
                   const void* act, const void* wpk, const float* bias, const void* res, void* out,
                   float* stats, float* out_nchw, int c_real, cudaStream_t st) {
  if (!is_pow2(H) || !is_pow2(W) || W > 128) {
    set_error("conv_gemm_bf16: H and W must be powers of two with W <= 128 (got %dx%d)", H, W);
    return HG_ERR_UNSUPPORTED;
  }
  if (Kp % 64 || Np % 64 || Np > 256) {
    set_error("conv_gemm_bf16: padded channels must be multiples of 64, Cout <= 256 (got %d -> %d)", Kp, Np);
    return HG_ERR_UNSUPPORTED;
  }
  const int bw = W < 128 ? W : 128;
  int bh = 128 / bw;
  if (bh > H) bh = H;
  const int bn = 128 / (bw * bh);
  const long long M = (long long)N * H * W;
  // N tile of at most 128 channels: short-K (1x1) kernels then fit two CTAs per SM (one CTA's epilogue overlaps
  // the other's TMA/MMA phase); 256 output channels = two N tiles that share the activation tile through L2.
  const int BN = Np > 128 ? 128 : Np;
  if (BN != 64 && BN != 128) {
    set_error("conv_gemm_bf16: unsupported padded Cout %d", Np);
    return HG_ERR_UNSUPPORTED;
  }
  CUtensorMap tmA, tmB, tmC, tmR;
  {
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Kp * 2, (uint64_t)W * Kp * 2, (uint64_t)H * W * Kp * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    uint32_t es[4] = {1, 1, 1, 1};
    int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)Np, (uint64_t)(R * S)};
    uint64_t str[2] = {(uint64_t)Kp * 2, (uint64_t)Np * Kp * 2};
    uint32_t box[3] = {64, (uint32_t)BN, 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wpk, dims, str, box, es,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
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
  }
  ConvGemmParams p;
  p.M_total = (int)M;
  p.H = H;
  p.W = W;
  p.taps_r = R;
  p.taps_s = S;
  p.dil = dil;
  p.pad = pad;
  p.sign = sign;
  p.kchunks = Kp / 64;
  p.n_total = Np;
  p.c_real = c_real;
  p.bias = bias;
  p.stats = stats;
  p.out_nchw = out_nchw;
  p.has_res = res != nullptr;
  p.n_tiles = Np / BN;
  const bool long_k = R * S * (Kp / 64) > 4;
  if (BN == 64) return long_k ? launch_conv_gemm<64, 4, 2>(tmA, tmB, tmC, tmR, p, st)
                              : launch_conv_gemm<64, 2, 3>(tmA, tmB, tmC, tmR, p, st);
  return long_k ? launch_conv_gemm<128, 4, 1>(tmA, tmB, tmC, tmR, p, st)
                : launch_conv_gemm<128, 2, 2>(tmA, tmB, tmC, tmR, p, st);
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
// ======================================================================================================
struct WgradParams {
  int H, W;
  int taps_s;        // filter width S
  int dil, pad;
  int tap_rows;      // taps handled per CTA (T)
  int n_panels;      // Cin_p / 64
  int Cin_p, Cout_p;
  int total_kb;      // ceil(M / 64)
  int kb_per_cta;
  int stages;
  int stage_bytes;
  float* dw;         // [taps][Cout_p][Cin_p] fp32, accumulated
};

__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tmem_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = p.n_panels * 64;
  const int T = p.tap_rows;
  const int tap0 = blockIdx.y * T;
  const int co_off = blockIdx.z * 128;
  const int kb_beg = blockIdx.x * p.kb_per_cta;
  int kb_end = kb_beg + p.kb_per_cta;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int nkb = kb_end - kb_beg;
  uint32_t cols = 32;
  while ((int)cols < T * N) cols <<= 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDy);
    prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int a_bytes = 2 * 8192;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int hw = p.H * p.W;
        for (int i = 0; i < nkb; ++i) {
          const int st = i % p.stages;
          const uint32_t ph = (i / p.stages) & 1;
          uint8_t* sA = smem + st * p.stage_bytes;
          uint8_t* sB = sA + a_bytes;
          const int m0 = (kb_beg + i) * 64;
          const int n0 = m0 / hw;
          const int rem = m0 - n0 * hw;
          const int h0 = rem / p.W;
          const int w0 = rem - h0 * p.W;
          mbar_wait(&empty_bar[st], ph ^ 1);
          mbar_expect_tx(&full_bar[st], p.stage_bytes);
          tma_load_2d(sA, &tmDy, &full_bar[st], co_off, m0);
          tma_load_2d(sA + 8192, &tmDy, &full_bar[st], co_off + 64, m0);
          for (int t = 0; t < T; ++t) {
            const int tap = tap0 + t;
            const int r = tap / p.taps_s, s = tap - r * p.taps_s;
            const int dh = r * p.dil - p.pad, dw = s * p.dil - p.pad;
            for (int pn = 0; pn < p.n_panels; ++pn)
              tma_load_4d(sB + (t * p.n_panels + pn) * 8192, &tmX, &full_bar[st], pn * 64, w0 + dw, h0 + dh, n0);
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const int st = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + st * p.stage_bytes);
          const uint32_t sB = sA + a_bytes;
          for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(sA + k * 2048, 8192, 1024);
              const uint64_t bdesc = make_smem_desc(sB + t * p.n_panels * 8192 + k * 2048, 8192, 1024);
              umma_bf16(tmem_base + t * N, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[st]);
          if (i == nkb - 1) umma_commit(tmem_full);
        }
        __syncwarp();
      }
    } else {
      const int sub = warp & 3;
      const int co = co_off + sub * 32 + lane;
      const bool row_ok = co < p.Cout_p;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
      // All CTAs of a split finish together and add into the SAME tile: start each CTA at a different column so
      // that concurrent atomics hit different addresses (the L2 atomic unit serialises per address).
      const int nchunk = N / 32;
      for (int tt = 0; tt < T; ++tt) {
        const int t = (tt + blockIdx.x) % T;
        float* dst = p.dw + ((size_t)(tap0 + t) * p.Cout_p + co) * p.Cin_p;
        for (int jj = 0; jj < nchunk; ++jj) {
          const int j = (jj + blockIdx.x) % nchunk;
          float v[32];
          tmem_ld32(taddr + t * N + j * 32, v);
          tmem_ld_wait();
          if (row_ok) {
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

int conv_wgrad_bf16(const HgConvDesc* d, const void* x, const void* dy, float* dw, float* dbias, cudaStream_t st) {
  const int Cin_p = pad64(d->Cin), Cout_p = pad64(d->Cout);
  const int H = d->H, W = d->W;
  const long long M = (long long)d->N * H * W;
  if (dw) {
    const int taps = d->R * d->S;
    int T = 1;
    if (taps == 9 && Cin_p <= 128) T = 3;
    const int bw = W < 64 ? W : 64;
    int bh = 64 / bw;
    if (bh > H) bh = H;
    const int bn = 64 / (bw * bh);
    CUtensorMap tmDy, tmX;
    {
      uint64_t dims[2] = {(uint64_t)Cout_p, (uint64_t)M};
      uint64_t str[1] = {(uint64_t)Cout_p * 2};
      uint32_t box[2] = {64, 64};
      uint32_t es[2] = {1, 1};
      int rc = encode_tmap(&tmDy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {(uint64_t)Cin_p, (uint64_t)W, (uint64_t)H, (uint64_t)d->N};
      uint64_t str[3] = {(uint64_t)Cin_p * 2, (uint64_t)W * Cin_p * 2, (uint64_t)H * W * Cin_p * 2};
      uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
      uint32_t es[4] = {1, 1, 1, 1};
      int rc = encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    WgradParams p;
    p.H = H;
    p.W = W;
    p.taps_s = d->S;
    p.dil = d->dil;
    p.pad = d->pad;
    p.tap_rows = T;
    p.n_panels = Cin_p / 64;
    p.Cin_p = Cin_p;
    p.Cout_p = Cout_p;
    p.total_kb = (int)((M + 63) / 64);
    p.stage_bytes = 2 * 8192 + T * p.n_panels * 8192;
    p.stages = (200 * 1024) / p.stage_bytes;
    if (p.stages > 6) p.stages = 6;
    const int tap_groups = taps / T;
    const int mgroups = (Cout_p + 127) / 128;
    // one wave of CTAs, and at least 8 K-blocks (512 pixels) of work per CTA: the split-K partial sums are
    // reduced with atomics, so small problems must not be cut into many slices
    int nsplit = kNumSMs / (tap_groups * mgroups);
    if (nsplit > p.total_kb / 8) nsplit = p.total_kb / 8;
    if (nsplit < 1) nsplit = 1;
    p.kb_per_cta = (p.total_kb + nsplit - 1) / nsplit;
    nsplit = (p.total_kb + p.kb_per_cta - 1) / p.kb_per_cta;
    p.dw = dw;
    const int smem_bytes = p.stages * p.stage_bytes + 1024 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
      HG_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set = true;
    }
    dim3 grid(nsplit, tap_groups, mgroups);
    conv_wgrad_kernel<<<grid, 192, smem_bytes, st>>>(tmDy, tmX, p);
    HG_LAUNCH_OK("conv_wgrad_kernel");
    count_launch();
  }
  if (dbias) return colsum_launch(HG_BF16, dy, M, Cout_p, d->Cout, dbias, st);
  return HG_OK;
}

}  // namespace hg
