// Stem of creatModel: 7x7 stride-2 pad-3 convolution 3 -> 64 + bias + ReLU, reading the fp32 NCHW image batch
// the training loop provides and writing NHWC activations (reference try_with_torch.py:262,276-277), and its
// backward (weight / bias gradient; the image needs no gradient).  K = 3*7*7 = 147 is a poor tensor-core shape
// and the layer is 0.3 % of the FLOPs, so this is a direct FFMA kernel with the input patch and the weights in
// shared memory.
#include "hg_common.cuh"

namespace hg {

int g_stem_bwd_blocks_per_sm = 2;                  // persistent blocks of the stem weight gradient (hg_set_option)
constexpr int kTileH = 8, kTileW = 16;             // output pixels per block
constexpr int kPatchH = 2 * kTileH + 5;            // 21
constexpr int kPatchW = 2 * kTileW + 5;            // 37
constexpr int kPatchWS = 38;                       // smem row stride
constexpr int kPatch = 3 * kPatchH * kPatchWS;     // floats
constexpr int kPatchPad = (kPatch + 1 + 3) & ~3;  // + the constant-one slot, 16 B aligned

__device__ __forceinline__ void load_patch(float* patch, const float* __restrict__ x, int n, int H, int W, int oy0,
                                           int ox0, int tid, int nthreads) {
  const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
  for (int i = tid; i < 3 * kPatchH * kPatchW; i += nthreads) {
    const int px = i % kPatchW;
    const int py = (i / kPatchW) % kPatchH;
    const int ci = i / (kPatchW * kPatchH);
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * 3 + ci) * H + iy) * W + ix];
    patch[(ci * kPatchH + py) * kPatchWS + px] = v;
  }
}

// grid = (Wo/16, Ho/8, N); 128 threads; thread = 4 consecutive x pixels x 16 channels
template <typename T>
__global__ void __launch_bounds__(128) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, T* __restrict__ y, int N, int H,
                                                       int W, int relu) {
  extern __shared__ float sm[];
  float* ws = sm;               // [147][64]
  float* patch = sm + 147 * 64; // [3][21][38]
  const int Ho = H / 2, Wo = W / 2;
  // The weights are transposed into shared memory ONCE per block (coalesced global reads: consecutive threads read
  // consecutive k of one output channel); the block then walks tiles (a block per tile re-read the 37 KB with a
  // 147-float stride between neighbouring threads for every 128 output pixels).
  for (int i = threadIdx.x; i < 147 * 64; i += 128) {
    const int co = i / 147, k = i - co * 147;  // w is OIHW: [co][ci][r][s] -> k = ci*49 + r*7 + s
    ws[k * 64 + co] = w[i];
  }
  const int tiles_x = Wo / kTileW, tiles_y = Ho / kTileH;
  const int ntiles = N * tiles_x * tiles_y;
  const int cg = threadIdx.x >> 5;       // 16-channel group (warp-uniform -> weight reads broadcast)
  const int q = threadIdx.x & 31;
  const int prow = q >> 2, pcol = (q & 3) * 4;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
  const int oy0 = ty * kTileH, ox0 = tx * kTileW;
  __syncthreads();                        // the previous tile's patch has been consumed
  load_patch(patch, x, n, H, W, oy0, ox0, threadIdx.x, 128);
  __syncthreads();
  // accumulators as packed fp32 pairs (FFMA2: two channels per issue slot; same IEEE results as scalar FFMA)
  float2 acc2[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc2[j][c] = make_float2(0.f, 0.f);
  for (int ci = 0; ci < 3; ++ci) {
    for (int r = 0; r < 7; ++r) {
      const float* prow_p = patch + (ci * kPatchH + 2 * prow + r) * kPatchWS + 2 * pcol;
      float in[13];
#pragma unroll
      for (int i = 0; i < 13; ++i) in[i] = prow_p[i];
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float4* wp = reinterpret_cast<const float4*>(ws + (ci * 49 + r * 7 + s) * 64 + cg * 16);
        float2 wv2[8];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 f = wp[v];
          wv2[2 * v] = make_float2(f.x, f.y);
          wv2[2 * v + 1] = make_float2(f.z, f.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 iv = make_float2(in[2 * j + s], in[2 * j + s]);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc2[j][c] = f2fma(iv, wv2[c], acc2[j][c]);
        }
      }
    }
  }
  float acc[4][16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc[j][2 * c] = acc2[j][c].x;
      acc[j][2 * c + 1] = acc2[j][c].y;
    }
  const int oy = oy0 + prow;
  if (oy < Ho) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ox = ox0 + pcol + j;
      if (ox >= Wo) continue;
      T* dst = y + (((long long)n * Ho + oy) * Wo + ox) * 64 + cg * 16;
      float o[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = acc[j][h * 8 + e] + (bias ? bias[cg * 16 + h * 8 + e] : 0.f);
          o[e] = relu ? fmaxf(v, 0.f) : v;
        }
        store8(dst + h * 8, o);
      }
    }
  }
  }
}

// Backward: dw[co][k] += sum_p g[p][co] * patch[p (+) k], db[co] += sum_p g[p][co], with g = dy * [y > 0].
// The bias is handled as a 148th "tap" whose input is the constant 1.  148 = 37 x 4: thread = 4 taps x 16
// channels (148 threads), persistent over tiles, accumulators in registers, one atomic per output at the end.
template <typename T>
__global__ void __launch_bounds__(160) stem_bwd_kernel(const float* __restrict__ x, const T* __restrict__ y,
                                                       const T* __restrict__ dy, float* __restrict__ dw,
                                                       float* __restrict__ db, int N, int H, int W, int relu) {
  extern __shared__ float sm[];
  float* patch = sm;            // [3][21][38] + 1 (constant one)
  float* gs = sm + kPatchPad;    // [128][64]
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = Wo / kTileW, tiles_y = Ho / kTileH;
  const long long ntiles = (long long)N * tiles_x * tiles_y;
  const int t = threadIdx.x;
  const bool worker = t < 148;
  const int cg = t & 3, kq = t >> 2;  // kq 0..36
  int koff[4], kmul[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = kq * 4 + j;
    if (k < 147) {
      const int ci = k / 49, r = (k % 49) / 7, s = k % 7;
      koff[j] = (ci * kPatchH + r) * kPatchWS + s;
      kmul[j] = 1;
    } else {
      koff[j] = kPatch;  // the constant one
      kmul[j] = 0;
    }
  }
  float2 acc2[4][8];   // packed fp32 pairs (FFMA2): two channels per issue slot
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc2[j][c] = make_float2(0.f, 0.f);
  if (t == 0) patch[kPatch] = 1.f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_x);
    const int ty = (int)((tile / tiles_x) % tiles_y);
    const int n = (int)(tile / ((long long)tiles_x * tiles_y));
    const int oy0 = ty * kTileH, ox0 = tx * kTileW;
    __syncthreads();
    load_patch(patch, x, n, H, W, oy0, ox0, t, 160);
    for (int i = t; i < 128 * 8; i += 160) {
      const int p = i >> 3, v = i & 7;
      const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
      const long long off = (((long long)n * Ho + oy) * Wo + ox) * 64 + v * 8;
      float yy[8], gg[8];
      load8(y + off, yy);
      load8(dy + off, gg);
#pragma unroll
      for (int e = 0; e < 8; ++e) gs[p * 64 + v * 8 + e] = (!relu || yy[e] > 0.f) ? gg[e] : 0.f;
    }
    __syncthreads();
    if (worker) {
      for (int p = 0; p < 128; ++p) {
        const int poff = (2 * (p >> 4)) * kPatchWS + 2 * (p & 15);
        const float4* gp = reinterpret_cast<const float4*>(gs + p * 64 + cg * 16);
        float2 g2[8];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 f = gp[v];
          g2[2 * v] = make_float2(f.x, f.y);
          g2[2 * v + 1] = make_float2(f.z, f.w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xv = patch[koff[j] + kmul[j] * poff];
          const float2 xv2 = make_float2(xv, xv);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc2[j][c] = f2fma(xv2, g2[c], acc2[j][c]);
        }
      }
    }
  }
  if (worker) {
    float acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        acc[j][2 * c] = acc2[j][c].x;
        acc[j][2 * c + 1] = acc2[j][c].y;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kq * 4 + j;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int co = cg * 16 + c;
        if (k < 147) {
          if (dw) atomicAdd(dw + co * 147 + k, acc[j][c]);
        } else if (db) {
          atomicAdd(db + co, acc[j][c]);
        }
      }
    }
  }
}

}  // namespace hg

namespace hg {
extern int g_stem_tc;
int stem_tc_fwd_launch(const float* x, const float* w, const float* bias, void* y, int N, int H, int W, int relu,
                       cudaStream_t st);
int stem_tc_bwd_launch(const float* x, const void* y, const void* dy, float* dw, float* db, int N, int H, int W, int relu,
                       cudaStream_t st);
}  // namespace hg

using namespace hg;

extern "C" {

int hg_stem_fwd(int dtype, const float* x_nchw, const float* w_oihw, const float* bias, int N, int H, int W, int relu,
                void* y, void* stream) {
  HG_REQUIRE(dtype == HG_BF16 || dtype == HG_F32, "hg_stem_fwd: bad dtype");
  HG_REQUIRE(x_nchw && w_oihw && y, "hg_stem_fwd: NULL pointer");
  HG_REQUIRE(N > 0 && H > 0 && W > 0, "hg_stem_fwd: non-positive size");
  if (H % 16 != 0 || W % 32 != 0) {
    set_error("hg_stem_fwd: image size %dx%d must be a multiple of 16x32", H, W);
    return HG_ERR_UNSUPPORTED;
  }
  if (dtype == HG_BF16 && g_stem_tc)   // bf16 path: im2col tile built in shared memory, tcgen05 GEMM (stem_tc.cu)
    return stem_tc_fwd_launch(x_nchw, w_oihw, bias, y, N, H, W, relu, (cudaStream_t)stream);
  const int smem = (147 * 64 + kPatch) * sizeof(float);
  long long ntiles = (long long)(W / 2 / kTileW) * (H / 2 / kTileH) * N;
  dim3 grid((unsigned)(ntiles < 4 * kNumSMs ? ntiles : 4 * kNumSMs));   // four resident blocks per SM walk the tiles
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HG_BF16) {
    static bool set = false;
    if (!set) {
      HG_CUDA_OK(cudaFuncSetAttribute(stem_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    stem_fwd_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias, (__nv_bfloat16*)y, N, H, W, relu);
  } else {
    static bool set = false;
    if (!set) {
      HG_CUDA_OK(cudaFuncSetAttribute(stem_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    stem_fwd_kernel<float><<<grid, 128, smem, st>>>(x_nchw, w_oihw, bias, (float*)y, N, H, W, relu);
  }
  HG_LAUNCH_OK("stem_fwd_kernel");
  count_launch();
  return HG_OK;
}

int hg_stem_bwd(int dtype, const float* x_nchw, const void* y, const void* dy, int N, int H, int W, int relu,
                float* dw_oihw, float* dbias, void* stream) {
  HG_REQUIRE(dtype == HG_BF16 || dtype == HG_F32, "hg_stem_bwd: bad dtype");
  HG_REQUIRE(x_nchw && y && dy, "hg_stem_bwd: NULL pointer");
  if (H % 16 != 0 || W % 32 != 0) {
    set_error("hg_stem_bwd: image size %dx%d must be a multiple of 16x32", H, W);
    return HG_ERR_UNSUPPORTED;
  }
  if (dtype == HG_BF16 && g_stem_tc)
    return stem_tc_bwd_launch(x_nchw, y, dy, dw_oihw, dbias, N, H, W, relu, (cudaStream_t)stream);
  const int smem = (kPatchPad + 128 * 64) * sizeof(float);
  const long long ntiles = (long long)N * (W / 2 / kTileW) * (H / 2 / kTileH);
  int blocks = g_stem_bwd_blocks_per_sm * kNumSMs;
  if (blocks > ntiles) blocks = (int)ntiles;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HG_BF16) {
    static bool set = false;
    if (!set) {
      HG_CUDA_OK(cudaFuncSetAttribute(stem_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    stem_bwd_kernel<__nv_bfloat16><<<blocks, 160, smem, st>>>(x_nchw, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy,
                                                             dw_oihw, dbias, N, H, W, relu);
  } else {
    static bool set = false;
    if (!set) {
      HG_CUDA_OK(cudaFuncSetAttribute(stem_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = true;
    }
    stem_bwd_kernel<float><<<blocks, 160, smem, st>>>(x_nchw, (const float*)y, (const float*)dy, dw_oihw, dbias, N, H, W,
                                                       relu);
  }
  HG_LAUNCH_OK("stem_bwd_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
