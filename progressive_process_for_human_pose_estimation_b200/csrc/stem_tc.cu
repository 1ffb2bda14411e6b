// Stem of creatModel on the tensor cores (bf16 path): 7x7 stride-2 pad-3 convolution 3 -> 64 + bias + ReLU on the fp32 NCHW
// image batch, and its weight / bias gradient (reference try_with_torch.py:262,276-277).
//
// K = 3*7*7 = 147 has no TMA form (3 channels innermost), so the im2col tile is BUILT in shared memory: a block loads the
// fp32 input patch of an 8 x 16 output tile (21 x 37 x 3 values) and its threads write the [128 pixels][192 k] bf16 tile
// (k = ci*49 + r*7 + s, zero-padded to three 64-element panels) in the 128-byte-swizzled layout tcgen05.mma reads.  That
// one tile is the K-major A operand of the forward GEMM (y[px, co] = tile[px, k] * w[co, k]) and, read MN-major, the A
// operand of the weight-gradient GEMM (dw[k, co] += tile[px, k] * g[px, co], pixels as the reduction dimension; slot
// k = 147 holds the constant 1, so the bias gradient is row 147).  The stem is the serial head of the forward graph and the
// serial tail of the backward graph: on the CUDA cores (stem.cu, the fp32 path) it cost 273 + 455 us of every step.
// Forward precision: the image enters as bf16(x) + bf16(x - bf16(x)) (two tiles, two passes of MMAs): with the image
// rounded to bf16 alone one sampled gradient of the three-stage `train` family moved 8.9 % from the fp32 reference
// (limit: twice the reference's own autocast divergence); the weights enter as bf16, as everywhere else.
#include "hg_common.cuh"

namespace hg {

int g_stem_tc = 1;   // bf16 path: tensor-core stem kernels (hg_set_option "stem_tc")

namespace {

constexpr int kTH = 8, kTW = 16;                 // output pixels per tile (128)
constexpr int kPH = 2 * kTH + 5;                 // 21 input rows
constexpr int kPW = 2 * kTW + 5;                 // 37 input columns
constexpr int kPWS = 38;                         // shared-memory row stride of the patch
constexpr int kPatchFloats = 3 * kPH * kPWS;     // 2394
constexpr int kPatchBytes = (kPatchFloats * 4 + 127) & ~127;
constexpr int kABytes = 3 * 16384;               // [3 panels][128 pixels][64 k] bf16
constexpr int kWBytes = 3 * 8192;                // forward: [3 panels][64 co][64 k] bf16
constexpr int kGBytes = 16384;                   // backward: [128 pixels][64 co] bf16
constexpr int kFwdSmem = 2 * kABytes + kWBytes + kPatchBytes + 512 + 1024;   // hi + lo image tiles
constexpr int kBwdSmem = kABytes + kGBytes + kPatchBytes + 512 + 1024;

__device__ __forceinline__ void load_patch(float* patch, const float* __restrict__ x, int n, int H, int W, int oy0, int ox0) {
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  for (int i = threadIdx.x; i < 3 * kPH * kPW; i += blockDim.x) {
    const int px = i % kPW;
    const int py = (i / kPW) % kPH;
    const int ci = i / (kPW * kPH);
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * 3 + ci) * H + iy) * W + ix];
    patch[(ci * kPH + py) * kPWS + px] = v;
  }
}

// Builder thread t < 192 owns the 16-byte chunk kc = t % 24 (k = 8*kc .. 8*kc+7) of the rows t/24 + 8*i: the patch
// offsets of its eight k are loop invariants.  one_slot: k == 147 is the constant 1 (backward) instead of padding.
struct TileBuilder {
  int koff[8];
  int kc, row0;
  __device__ __forceinline__ TileBuilder(bool one_slot) {
    kc = threadIdx.x % 24;
    row0 = threadIdx.x / 24;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc * 8 + j;
      if (k < 147) {
        const int ci = k / 49, r = (k % 49) / 7, s = k % 7;
        koff[j] = (ci * kPH + r) * kPWS + s;
      } else {
        koff[j] = (one_slot && k == 147) ? -2 : -1;
      }
    }
  }
  // sLo != null: a second tile with the bf16 residuals x - bf16(x) (forward: image precision 2^-17 instead of 2^-9)
  __device__ __forceinline__ void build(uint8_t* sA, const float* patch, uint8_t* sLo = nullptr) const {
    uint8_t* panel = sA + (kc >> 3) * 16384;
    const int chunk = kc & 7;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int row = row0 + 8 * i;
      const float* base = patch + (2 * (row >> 4)) * kPWS + 2 * (row & 15);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = koff[j] >= 0 ? base[koff[j]] : (koff[j] == -2 ? 1.f : 0.f);
      uint4 u;
      u.x = f2_to_bf2(make_float2(v[0], v[1]));
      u.y = f2_to_bf2(make_float2(v[2], v[3]));
      u.z = f2_to_bf2(make_float2(v[4], v[5]));
      u.w = f2_to_bf2(make_float2(v[6], v[7]));
      *reinterpret_cast<uint4*>(panel + row * 128 + ((chunk ^ (row & 7)) << 4)) = u;
      if (sLo != nullptr) {
        const float2 h0 = bf2_to_f2(u.x), h1 = bf2_to_f2(u.y), h2 = bf2_to_f2(u.z), h3 = bf2_to_f2(u.w);
        uint4 l;
        l.x = f2_to_bf2(make_float2(v[0] - h0.x, v[1] - h0.y));
        l.y = f2_to_bf2(make_float2(v[2] - h1.x, v[3] - h1.y));
        l.z = f2_to_bf2(make_float2(v[4] - h2.x, v[5] - h2.y));
        l.w = f2_to_bf2(make_float2(v[6] - h3.x, v[7] - h3.y));
        *reinterpret_cast<uint4*>(sLo + (kc >> 3) * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = l;
      }
    }
  }
};

__global__ void __launch_bounds__(256, 1)
stem_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ y, int N, int H, int W, int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sLo = smem + kABytes;
  uint8_t* sW = smem + 2 * kABytes;
  float* patch = reinterpret_cast<float*>(smem + 2 * kABytes + kWBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kABytes + kWBytes + kPatchBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* bias_s = reinterpret_cast<float*>(bar + 2);   // [64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = Wo / kTW, tiles_y = Ho / kTH;
  const int ntiles = N * tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  pdl_wait();
  // weights once per block: [co][k] bf16, K-major, three 64-k panels (k >= 147: zero)
  for (int i = threadIdx.x; i < 64 * 24; i += 256) {
    const int co = i / 24, kc = i % 24;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc * 8 + j;
      v[j] = k < 147 ? w[co * 147 + k] : 0.f;
    }
    uint4 u;
    u.x = f2_to_bf2(make_float2(v[0], v[1]));
    u.y = f2_to_bf2(make_float2(v[2], v[3]));
    u.z = f2_to_bf2(make_float2(v[4], v[5]));
    u.w = f2_to_bf2(make_float2(v[6], v[7]));
    *reinterpret_cast<uint4*>(sW + (kc >> 3) * 8192 + co * 128 + (((kc & 7) ^ (co & 7)) << 4)) = u;
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();

  const TileBuilder tb(false);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, 1024);
  const uint64_t wdesc0 = make_smem_desc(smem_u32(sW), 16, 1024);
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int oy0 = ty * kTH, ox0 = tx * kTW;
    load_patch(patch, x, n, H, W, oy0, ox0);
    __syncthreads();
    if (threadIdx.x < 192) tb.build(sA, patch, sLo);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        // y = bf16(x) * w + (x - bf16(x)) * w: the image enters with 16 mantissa bits, the weights as bf16
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base, adesc0 + (uint64_t)(t * 3072 + p * 1024 + ks * 2), wdesc0 + (uint64_t)(p * 512 + ks * 2),
                        idesc, (t | p | ks) ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);      // accumulator complete; the tile and the patch may be rebuilt
    phase ^= 1;
    tc_fence_after();
    {
      // epilogue: warp = (TMEM lane quarter, column half); thread = one pixel, 32 channels
      const int row = (warp & 3) * 32 + lane, half = warp >> 2;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + half * 32, v);
      tmem_ld_wait();
      const int oy = oy0 + (row >> 4), ox = ox0 + (row & 15);
      __nv_bfloat16* dst = y + (((long long)n * Ho + oy) * Wo + ox) * 64 + half * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = v[q * 8 + e] + bias_s[half * 32 + q * 8 + e];
          o[e] = relu ? fmaxf(t, 0.f) : t;
        }
        store8(dst + q * 8, o);
      }
    }
    tc_fence_before();
    __syncthreads();            // every TMEM read is done before the next tile's first MMA overwrites the accumulator
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

__global__ void __launch_bounds__(256, 2)
stem_tc_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                   float* __restrict__ dw, float* __restrict__ db, int N, int H, int W, int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // [3 panels][128 pixels][64 k]: MN-major A (M = k, K = pixels)
  uint8_t* sG = smem + kABytes;             // [128 pixels][64 co]: MN-major B (N = co, K = pixels)
  float* patch = reinterpret_cast<float*>(smem + kABytes + kGBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kABytes + kGBytes + kPatchBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = Wo / kTW, tiles_y = Ho / kTH;
  const int ntiles = N * tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const TileBuilder tb(true);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
  // D1 = rows k 0..127 (panels 0, 1), D2 = rows k 64..191 (panels 1, 2): rows 64..127 of D2 are k = 128..191
  const uint64_t a1desc0 = make_smem_desc(smem_u32(sA), 16384, 1024);
  const uint64_t a2desc0 = make_smem_desc(smem_u32(sA) + 16384, 16384, 1024);
  const uint64_t gdesc0 = make_smem_desc(smem_u32(sG), 16384, 1024);
  uint32_t phase = 0;
  uint32_t accum = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int oy0 = ty * kTH, ox0 = tx * kTW;
    load_patch(patch, x, n, H, W, oy0, ox0);
    // g = dy * [y > 0] (the stored activation is the ReLU mask)
    for (int i = threadIdx.x; i < 128 * 8; i += 256) {
      const int row = i >> 3, vch = i & 7;
      const int oy = oy0 + (row >> 4), ox = ox0 + (row & 15);
      const long long off = (((long long)n * Ho + oy) * Wo + ox) * 64 + vch * 8;
      float yy[8], gg[8];
      load8(y + off, yy);
      load8(dy + off, gg);
#pragma unroll
      for (int e = 0; e < 8; ++e) gg[e] = (!relu || yy[e] > 0.f) ? gg[e] : 0.f;
      uint4 u;
      u.x = f2_to_bf2(make_float2(gg[0], gg[1]));
      u.y = f2_to_bf2(make_float2(gg[2], gg[3]));
      u.z = f2_to_bf2(make_float2(gg[4], gg[5]));
      u.w = f2_to_bf2(make_float2(gg[6], gg[7]));
      *reinterpret_cast<uint4*>(sG + row * 128 + ((vch ^ (row & 7)) << 4)) = u;
    }
    __syncthreads();
    if (threadIdx.x < 192) tb.build(sA, patch);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 pixels further = 2048 B in both MN-major operands
          umma_bf16(tmem_base, a1desc0 + (uint64_t)(ks * 128), gdesc0 + (uint64_t)(ks * 128), idesc, (accum | (uint32_t)ks) ? 1u : 0u);
          umma_bf16(tmem_base + 64, a2desc0 + (uint64_t)(ks * 128), gdesc0 + (uint64_t)(ks * 128), idesc,
                    (accum | (uint32_t)ks) ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    accum = 1;
    mbar_wait(bar, phase);      // the MMAs have read the tiles: they may be rebuilt
    phase ^= 1;
  }
  tc_fence_after();
  if (warp < 4 && (int)blockIdx.x < ntiles) {
    // lane L of the accumulators = k: D1 -> k = L, D2 -> k = 64 + L (only L >= 64 is new)
    const int L = warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
      const int k = part == 0 ? L : 64 + L;
      const bool use = part == 0 || L >= 64;     // (warp-uniform: L >= 64 <=> warp >= 2)
      if (!use) continue;
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        float v[32];
        tmem_ld32(taddr + part * 64 + j * 32, v);
        tmem_ld_wait();
        if (k < 147) {
          if (dw) {
#pragma unroll
            for (int c = 0; c < 32; ++c) atomicAdd(dw + (j * 32 + c) * 147 + k, v[c]);
          }
        } else if (k == 147 && db) {
#pragma unroll
          for (int c = 0; c < 32; ++c) atomicAdd(db + j * 32 + c, v[c]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

int stem_tc_fwd_launch(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int relu,
                       cudaStream_t st) {
  static bool set = false;
  if (!set) {
    HG_CUDA_OK(cudaFuncSetAttribute(stem_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    set = true;
  }
  const long long ntiles = (long long)N * (W / 2 / kTW) * (H / 2 / kTH);
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);   // 131 KB of shared memory: one block per SM
  launch_k(stem_tc_fwd_kernel, dim3(grid), dim3(256), (size_t)kFwdSmem, st, x, w, bias, (__nv_bfloat16*)y, N, H, W, relu);
  HG_LAUNCH_OK("stem_tc_fwd_kernel");
  count_launch();
  return HG_OK;
}

int stem_tc_bwd_launch(const float* x, const void* y, const void* dy, float* dw, float* db, int N, int H, int W, int relu,
                       cudaStream_t st) {
  static bool set = false;
  if (!set) {
    HG_CUDA_OK(cudaFuncSetAttribute(stem_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    set = true;
  }
  const long long ntiles = (long long)N * (W / 2 / kTW) * (H / 2 / kTH);
  const int grid = (int)(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
  launch_k(stem_tc_bwd_kernel, dim3(grid), dim3(256), (size_t)kBwdSmem, st, x, (const __nv_bfloat16*)y,
           (const __nv_bfloat16*)dy, dw, db, N, H, W, relu);
  HG_LAUNCH_OK("stem_tc_bwd_kernel");
  count_launch();
  return HG_OK;
}

}  // namespace hg
