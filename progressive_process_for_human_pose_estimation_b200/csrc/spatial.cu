// Memory-bound spatial kernels of the hourglass on NHWC activations: 2x2 max-pool, x2 up-sampling fused with
// the skip add (bilinear align_corners=True and nearest), their adjoints, tensor add, and the NCHW<->NHWC
// conversions at the module boundary.  One thread = 8 consecutive channels of one output pixel (128-bit
// accesses for bf16), grid-stride over pixels.
//
// Replaces nn.MaxPool2d(2), F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) + `up1 + up2`
// (reference try_with_torch.py:220,226,238-239,265) and the nearest variant (hourglass_compare.py:532-542).
#include "hg_common.cuh"

namespace hg {

int g_upsample_fwd_cap = 0;   // blocks per SM of the forward kernel (0 = policy)
int g_upsample_sep = 1;   // 1: separable up-sampling kernels through shared memory (hg_set_option "upsample_sep")


// ------------------------------------------------------------------------------------------------------
// max-pool 2x2 stride 2
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H,
                                                           int W, int Cp, float* __restrict__ stats) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * vecs;
  float st_s[8] = {}, st_q[8] = {}, pv[8] = {};
  if (stats != nullptr) {   // shifted sums (bn.cu): the pivots of this thread's channels
#pragma unroll
    for (int e = 0; e < 8; ++e) pv[e] = stats[2 * Cp + (threadIdx.x % vecs) * 8 + e];
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vc = (int)(i % vecs);
    long long pix = i / vecs;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const T* p = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * Cp + vc * 8;
    float a[8], b[8], c[8], d[8], o[8];
    load8(p, a);
    load8(p + Cp, b);
    load8(p + (long long)W * Cp, c);
    load8(p + (long long)W * Cp + Cp, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = fmaxf(fmaxf(a[e], b[e]), fmaxf(c[e], d[e]));
      const float r = to_f(from_f<T>(o[e])) - pv[e];  // statistics of the values as stored
      st_s[e] += r;
      st_q[e] = fmaf(r, r, st_q[e]);
    }
    store8(y + i * 8, o);
  }
  if (stats != nullptr) block_channel_stats(st_s, st_q, Cp, stats);  // (vc is the same for every row of a thread)
}

// dx[window] = dy routed to the FIRST row-major maximum of the window (PyTorch's tie-break), [+ addend]
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           const T* __restrict__ addend, T* __restrict__ dx, int N,
                                                           int H, int W, int Cp) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vc = (int)(i % vecs);
    long long pix = i / vecs;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * Cp + vc * 8;
    const long long off[4] = {0, Cp, (long long)W * Cp, (long long)W * Cp + Cp};
    float v[4][8], g[8], o[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) load8(x + base + off[q], v[q]);
    load8(dy + i * 8, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int best = 0;
      float bv = v[0][e];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q][e] > bv) {
          bv = v[q][e];
          best = q;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][e] = (q == best) ? g[e] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (addend) {
        float ad[8];
        load8(addend + base + off[q], ad);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[q][e] += ad[e];
      }
      store8(dx + base + off[q], o[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// x2 up-sampling (+ skip add).  mode 0 = bilinear, align_corners=True; mode 1 = nearest.
// PyTorch's index arithmetic is reproduced in fp32: scale = (in-1)/(out-1), src = scale*dst, i0 = (int)src,
// i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_coord(int o, int in_size, float scale, int& i0, int& i1, float& l0, float& l1) {
  const float src = scale * (float)o;
  i0 = (int)src;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// Work decomposition of both kernels: a block walks whole image ROWS (row = (n, y)), its threads the (x, channel vector)
// pairs of the row.  All index arithmetic is 32-bit and the row's vertical coordinates / weights are computed once per
// row: the element-wise form (three 64-bit divisions and two coordinate evaluations per 16-byte vector) was
// instruction-bound at 2.5 TB/s (forward) / 1.4 TB/s (backward) on the 64x64 maps.
template <typename T>
__global__ void __launch_bounds__(256) upsample2_add_fwd_kernel(const T* __restrict__ low, const T* __restrict__ skip,
                                                                T* __restrict__ out, int N, int h, int w, int Cp,
                                                                int mode, float* __restrict__ stats) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, H = 2 * h, W = 2 * w;
  const float sh = h > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sw = w > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int rows = N * H;
  const int row_vecs = W * vecs;
  // 256 % vecs == 0 (vecs = 8, 16 or 32): a thread keeps its channel vector across iterations (per-thread statistics)
  const int vc = threadIdx.x % vecs;
  float st_s[8] = {}, st_q[8] = {}, pv[8] = {};
  if (stats != nullptr) {   // shifted sums (bn.cu): the pivots of this thread's channels
#pragma unroll
    for (int e = 0; e < 8; ++e) pv[e] = stats[2 * Cp + vc * 8 + e];
  }
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / H, y = r - n * H;
    int y0, y1;
    float ly0, ly1;
    bilin_coord(y, h, sh, y0, y1, ly0, ly1);
    if (mode == 1) y0 = y >> 1;
    const T* lrow0 = low + ((long long)n * h + y0) * w * Cp + vc * 8;
    const T* lrow1 = low + ((long long)n * h + y1) * w * Cp + vc * 8;
    const long long obase = (long long)r * row_vecs * 8;
    for (int t = threadIdx.x; t < row_vecs; t += 256) {
      const int x = t / vecs;
      float o[8];
      if (mode == 1) {
        load8(lrow0 + (long long)(x >> 1) * Cp, o);
      } else {
        int x0, x1;
        float lx0, lx1;
        bilin_coord(x, w, sw, x0, x1, lx0, lx1);
        float a[8], b[8], c[8], d[8];
        load8(lrow0 + (long long)x0 * Cp, a);
        load8(lrow0 + (long long)x1 * Cp, b);
        load8(lrow1 + (long long)x0 * Cp, c);
        load8(lrow1 + (long long)x1 * Cp, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ly0 * (lx0 * a[e] + lx1 * b[e]) + ly1 * (lx0 * c[e] + lx1 * d[e]);
      }
      if (skip) {
        float s[8];
        load8(skip + obase + (long long)t * 8, s);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += s[e];
      }
      if (stats != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float rr = to_f(from_f<T>(o[e])) - pv[e];
          st_s[e] += rr;
          st_q[e] = fmaf(rr, rr, st_q[e]);
        }
      }
      store8(out + obase + (long long)t * 8, o);
    }
  }
  if (stats != nullptr) block_channel_stats(st_s, st_q, Cp, stats);
}

// adjoint as a GATHER over the low-resolution pixels (deterministic, no atomics):
// dlow[yi, xi] = sum over outputs (y, x) that read (yi, xi) of weight * dout[y, x]   [+ addend]
// The bilinear weights are separable: wy[j] of the (at most six) candidate output rows once per low-resolution row,
// wx[k] of the candidate columns once per pixel.
template <typename T>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ addend,
                                                            T* __restrict__ dlow, int N, int h, int w, int Cp,
                                                            int mode) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, H = 2 * h, W = 2 * w;
  const float sh = h > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sw = w > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int rows = N * h;
  const int row_vecs = w * vecs;
  const int vc = threadIdx.x % vecs;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / h, yi = r - n * h;
    const T* db = dout + (long long)n * H * W * Cp + vc * 8;
    // candidate outputs: src in (yi-1, yi+1)  ->  y in [2*yi-2, 2*yi+3] is a safe superset for scale ~ 1/2
    const int ylo = max(0, 2 * yi - 2), yhi = min(H - 1, 2 * yi + 3);
    float wy[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int y = ylo + j;
      float v = 0.f;
      if (mode == 1) {
        v = (y <= yhi && (y >> 1) == yi) ? 1.f : 0.f;
      } else if (y <= yhi) {
        int y0, y1;
        float ly0, ly1;
        bilin_coord(y, h, sh, y0, y1, ly0, ly1);
        if (y0 == yi) v += ly0;
        if (y1 == yi) v += ly1;
      }
      wy[j] = v;
    }
    const long long obase = (long long)r * row_vecs * 8;
    for (int t = threadIdx.x; t < row_vecs; t += 256) {
      const int xi = t / vecs;
      const int xlo = max(0, 2 * xi - 2), xhi = min(W - 1, 2 * xi + 3);
      float wx[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int x = xlo + k;
        float v = 0.f;
        if (mode == 1) {
          v = (x <= xhi && (x >> 1) == xi) ? 1.f : 0.f;
        } else if (x <= xhi) {
          int x0, x1;
          float lx0, lx1;
          bilin_coord(x, w, sw, x0, x1, lx0, lx1);
          if (x0 == xi) v += lx0;
          if (x1 == xi) v += lx1;
        }
        wx[k] = v;
      }
      float acc[8] = {};
      // same visiting order (rows, then columns) and the same weight product wy * wx as the element-wise form
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (wy[j] == 0.f) continue;
        const T* drow = db + (long long)(ylo + j) * W * Cp;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (wx[k] == 0.f) continue;
          float g[8];
          load8(drow + (long long)(xlo + k) * Cp, g);
          const float wgt = wy[j] * wx[k];
          if (mode == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += g[e];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, g[e], acc[e]);
          }
        }
      }
      if (addend) {
        float ad[8];
        load8(addend + obase + (long long)t * 8, ad);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += ad[e];
      }
      store8(dlow + obase + (long long)t * 8, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Separable forms of the two kernels above through shared memory (used whenever the row buffer fits).  The bilinear
// weights factor into a vertical and a horizontal pair, so an output row is the horizontal blend of ONE vertically
// blended low-resolution row (forward: 2 instead of 4 gathers per output vector), and a low-resolution gradient row is
// the horizontal reduction of ONE vertically reduced full-resolution row (backward: 6 coalesced row loads per
// full-resolution vector instead of up to 36 dependent gathers per low-resolution vector).  Measured on B200 at
// 32x32 -> 64x64, 256 channels, batch 32 (tools/gpu_upsample_probe.py): forward 68.4 -> 49.3 us, backward 58.8 -> 34.0 us.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upsample2_add_fwd_sep_kernel(const T* __restrict__ low, const T* __restrict__ skip,
                                                                    T* __restrict__ out, int N, int h, int w, int Cp,
                                                                    int mode, float* __restrict__ stats) {
  extern __shared__ float4 up_smem4[];
  float* m = reinterpret_cast<float*>(up_smem4);   // [w][Cp]: vertically blended low-resolution row
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, H = 2 * h, W = 2 * w;
  const float sh = h > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sw = w > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int rows = N * H;
  const int vc = threadIdx.x % vecs, x00 = threadIdx.x / vecs, rlanes = 256 / vecs;
  float st_s[8] = {}, st_q[8] = {}, pv[8] = {};
  if (stats != nullptr) {   // shifted sums (bn.cu): the pivots of this thread's channels
#pragma unroll
    for (int e = 0; e < 8; ++e) pv[e] = stats[2 * Cp + vc * 8 + e];
  }
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / H, y = r - n * H;
    int y0, y1;
    float ly0, ly1;
    bilin_coord(y, h, sh, y0, y1, ly0, ly1);
    if (mode == 1) y0 = y >> 1;
    const T* lrow0 = low + ((long long)n * h + y0) * w * Cp + vc * 8;
    const T* lrow1 = low + ((long long)n * h + y1) * w * Cp + vc * 8;
    constexpr int kI1 = 2;                        // low-resolution vectors per thread in flight (x 2 rows)
    constexpr int kI = sizeof(T) == 2 ? 4 : 2;   // output vectors per thread in flight (measured: a third block per SM
                                                 // at 80 registers spills and is slower, 69.7 vs 49.3 us)
    // phase 1: m[xl] = ly0 * low[y0][xl] + ly1 * low[y1][xl]  (nearest: the row itself)
    for (int xb = x00; xb < w; xb += kI1 * rlanes) {
      Raw8<T> ra[kI1], rc[kI1];
#pragma unroll
      for (int i = 0; i < kI1; ++i) {
        const int xl = xb + i * rlanes;
        if (xl < w) {
          load_raw(lrow0 + (long long)xl * Cp, ra[i]);
          if (mode == 0) load_raw(lrow1 + (long long)xl * Cp, rc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < kI1; ++i) {
        const int xl = xb + i * rlanes;
        if (xl < w) {
          float a[8], c[8];
          unpack(ra[i], a);
          if (mode == 0) {
            unpack(rc[i], c);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = ly0 * a[e] + ly1 * c[e];
          }
          float4* dst = reinterpret_cast<float4*>(m + xl * Cp + vc * 8);
          dst[0] = make_float4(a[0], a[1], a[2], a[3]);
          dst[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
      }
    }
    __syncthreads();
    // phase 2: out[x] = lx0 * m[x0] + lx1 * m[x1] (+ skip)
    const long long obase = (long long)r * W * Cp + vc * 8;
    for (int xb = x00; xb < W; xb += kI * rlanes) {
      Raw8<T> rs[kI];
#pragma unroll
      for (int i = 0; i < kI; ++i) {
        const int x = xb + i * rlanes;
        if (skip && x < W) load_raw(skip + obase + (long long)x * Cp, rs[i]);
      }
#pragma unroll
      for (int i = 0; i < kI; ++i) {
        const int x = xb + i * rlanes;
        if (x >= W) continue;
        float o[8];
        if (mode == 1) {
          const float4* src = reinterpret_cast<const float4*>(m + (x >> 1) * Cp + vc * 8);
          const float4 a = src[0], b = src[1];
          o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
        } else {
          int x0, x1;
          float lx0, lx1;
          bilin_coord(x, w, sw, x0, x1, lx0, lx1);
          const float4* s0 = reinterpret_cast<const float4*>(m + x0 * Cp + vc * 8);
          const float4* s1 = reinterpret_cast<const float4*>(m + x1 * Cp + vc * 8);
          const float4 a = s0[0], b = s0[1], c = s1[0], d = s1[1];
          o[0] = lx0 * a.x + lx1 * c.x; o[1] = lx0 * a.y + lx1 * c.y; o[2] = lx0 * a.z + lx1 * c.z;
          o[3] = lx0 * a.w + lx1 * c.w; o[4] = lx0 * b.x + lx1 * d.x; o[5] = lx0 * b.y + lx1 * d.y;
          o[6] = lx0 * b.z + lx1 * d.z; o[7] = lx0 * b.w + lx1 * d.w;
        }
        if (skip) {
          float sv[8];
          unpack(rs[i], sv);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += sv[e];
        }
        if (stats != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float rr = to_f(from_f<T>(o[e])) - pv[e];
            st_s[e] += rr;
            st_q[e] = fmaf(rr, rr, st_q[e]);
          }
        }
        store8(out + obase + (long long)x * Cp, o);
      }
    }
    __syncthreads();
  }
  if (stats != nullptr) block_channel_stats(st_s, st_q, Cp, stats);
}

template <typename T>
__global__ void __launch_bounds__(256, 3) upsample2_bwd_sep_kernel(const T* __restrict__ dout, const T* __restrict__ addend,
                                                                T* __restrict__ dlow, int N, int h, int w, int Cp,
                                                                int mode) {
  extern __shared__ float4 up_smem4[];
  float* v = reinterpret_cast<float*>(up_smem4);   // [W][Cp]: sum_j wy[j] * dout[ylo + j][x]
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3, H = 2 * h, W = 2 * w;
  const float sh = h > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float sw = w > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int rows = N * h;
  const int vc = threadIdx.x % vecs, x00 = threadIdx.x / vecs, rlanes = 256 / vecs;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / h, yi = r - n * h;
    const T* db = dout + (long long)n * H * W * Cp + vc * 8;
    // candidate outputs: src in (yi-1, yi+1)  ->  y in [2*yi-2, 2*yi+3] is a safe superset for scale ~ 1/2
    const int ylo = max(0, 2 * yi - 2), yhi = min(H - 1, 2 * yi + 3);
    float wy[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int y = ylo + j;
      float wv = 0.f;
      if (mode == 1) {
        wv = (y <= yhi && (y >> 1) == yi) ? 1.f : 0.f;
      } else if (y <= yhi) {
        int y0, y1;
        float ly0, ly1;
        bilin_coord(y, h, sh, y0, y1, ly0, ly1);
        if (y0 == yi) wv += ly0;
        if (y1 == yi) wv += ly1;
      }
      wy[j] = wv;
    }
    // phase 1: vertical reduction of the candidate rows.  The non-zero weights are consecutive (at most five rows):
    // all five row loads of two full-resolution vectors are issued before the first is consumed (ten 16-byte loads
    // in flight per thread: a loop over the rows with a load -> FMA dependency each was one L2 round trip per row)
    int first = 0;
#pragma unroll
    for (int j = 5; j >= 0; --j)
      if (wy[j] != 0.f) first = j;
    float rw[5];
    const T* rp[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float wv = 0.f;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j == first + q) wv = wy[j];
      rw[q] = wv;
      const int y = min(ylo + first + q, yhi);   // (weight 0 beyond the candidates: any valid row)
      rp[q] = db + (long long)y * W * Cp;
    }
    constexpr int kI = sizeof(T) == 2 ? 2 : 1;   // full-resolution vectors per thread in flight (x 5 rows)
    for (int xb = x00; xb < W; xb += kI * rlanes) {
      Raw8<T> raw[5][kI];
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int i = 0; i < kI; ++i) {
          const int x = xb + i * rlanes;
          if (x < W) load_raw(rp[q] + (long long)x * Cp, raw[q][i]);
        }
#pragma unroll
      for (int i = 0; i < kI; ++i) {
        const int x = xb + i * rlanes;
        if (x < W) {
          float acc[8] = {};
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            float gq[8];
            unpack(raw[q][i], gq);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(rw[q], gq[e], acc[e]);
          }
          float4* dst = reinterpret_cast<float4*>(v + x * Cp + vc * 8);
          dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
    __syncthreads();
    // phase 2: horizontal reduction
    const long long obase = (long long)r * w * Cp + vc * 8;
    for (int xi = x00; xi < w; xi += rlanes) {
      const int xlo = max(0, 2 * xi - 2), xhi = min(W - 1, 2 * xi + 3);
      float acc[8] = {};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int x = xlo + k;
        if (x > xhi) continue;
        float wv = 0.f;
        if (mode == 1) {
          wv = (x >> 1) == xi ? 1.f : 0.f;
        } else {
          int x0, x1;
          float lx0, lx1;
          bilin_coord(x, w, sw, x0, x1, lx0, lx1);
          if (x0 == xi) wv += lx0;
          if (x1 == xi) wv += lx1;
        }
        if (wv == 0.f) continue;
        const float4* src = reinterpret_cast<const float4*>(v + x * Cp + vc * 8);
        const float4 a = src[0], b = src[1];
        acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]);
        acc[3] = fmaf(wv, a.w, acc[3]); acc[4] = fmaf(wv, b.x, acc[4]); acc[5] = fmaf(wv, b.y, acc[5]);
        acc[6] = fmaf(wv, b.z, acc[6]); acc[7] = fmaf(wv, b.w, acc[7]);
      }
      if (addend) {
        float ad[8];
        load8(addend + obase + (long long)xi * Cp, ad);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += ad[e];
      }
      store8(dlow + obase + (long long)xi * Cp, acc);
    }
    __syncthreads();
  }
}

// out = a + b
template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                  long long nvec) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    load8(a + i * 8, x);
    load8(b + i * 8, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    store8(out + i * 8, x);
  }
}

// fp32 NCHW [N,C,H,W] -> NHWC T [N,H,W,Cp] (zero padded channels) [+ addend NHWC]: gradients of the heatmaps
// coming back from the loss.  One thread = one pixel x 8 channels; reads are coalesced across pixels.
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, const T* __restrict__ addend,
                                                           T* __restrict__ dst, int N, int C, int HW, int Cp) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const int vecs = Cp >> 3;
  const long long total = (long long)N * HW * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // pixel fastest so that a warp reads consecutive pixels of one channel plane
    const int p = (int)(i % HW);
    long long r = i / HW;
    const int vc = (int)(r % vecs);
    const int n = (int)(r / vecs);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = vc * 8 + e;
      o[e] = (src != nullptr && c < C) ? src[((long long)n * C + c) * HW + p] : 0.f;
    }
    const long long di = ((long long)n * HW + p) * Cp + vc * 8;
    if (addend) {
      float ad[8];
      load8(addend + di, ad);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += ad[e];
    }
    store8(dst + di, o);
  }
}

// NHWC T [N,H,W,Cp] -> fp32 NCHW [N,C,H,W]
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N,
                                                           int C, int HW, int Cp) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const long long total = (long long)N * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    long long r = i / HW;
    const int c = (int)(r % C);
    const int n = (int)(r / C);
    dst[i] = to_f(src[((long long)n * HW + p) * Cp + c]);
  }
}

// with_stats: every block ends with 2*Cp/4 vector atomics into the same cache lines -> fewer, fatter blocks
static inline int grid_for(long long total, bool with_stats = false) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * (with_stats ? 4 : 16);
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}


// ------------------------------------------------------------------------------------------------------
// Global average pooling and its transpose, the image-level branch of ASPP (train.py:476-479,488-489:
// nn.AdaptiveAvgPool2d((1, 1)) ... F.interpolate(x5, size, mode='bilinear', align_corners=True) of a 1x1 map, i.e. a
// broadcast).  y[n, c] = scale * sum_hw x[n, hw, c] [+ addend[n, c]];   x[n, hw, c] = scale * y[n, c] [+ addend].
// Each is the other's backward.  Block = (n, 32 channel vectors) x 8 row lanes.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) spatial_sum_kernel(const T* __restrict__ x, const T* __restrict__ addend,
                                                          T* __restrict__ y, int HW, int Cp, float scale) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][32][9];
  const int vecs = Cp >> 3;
  const int vl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int vc = blockIdx.x * 32 + vl;
  const long long n = blockIdx.y;
  float acc[8] = {};
  if (vc < vecs) {
    for (int r = rl; r < HW; r += 8) {
      float v[8];
      load8(x + (n * HW + r) * Cp + vc * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][vl][e] = acc[e];
  __syncthreads();
  if (rl == 0 && vc < vecs) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[k][vl][e];
      o[e] = t * scale;
    }
    if (addend) {
      float a[8];
      load8(addend + n * Cp + vc * 8, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += a[e];
    }
    store8(y + n * Cp + vc * 8, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) spatial_bcast_kernel(const T* __restrict__ y, const T* __restrict__ addend,
                                                            T* __restrict__ x, long long total, int HW, int Cp,
                                                            float scale) {
  pdl_wait();
  pdl_trigger();
  const int vecs = Cp >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vc = (int)(i % vecs);
    const long long n = i / vecs / HW;
    float v[8];
    load8(y + n * Cp + vc * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= scale;
    if (addend) {
      float a[8];
      load8(addend + i * 8, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += a[e];
    }
    store8(x + i * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------------
// Channel-window copy: dst[r, dst_c0 + c] = src[r, src_c0 + c] [+ addend[r, dst_c0 + c]] for c < 8 * vecs.
// torch.cat(xs, dim=1) whose result feeds a BatchNorm (train.py:528-538,570-583) is one call per input; its backward
// (a channel slice of the gradient, accumulated into the input's gradient) is the same kernel with the roles swapped.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) channel_copy_kernel(const T* __restrict__ src, int src_Cp, int src_c0,
                                                           const T* __restrict__ addend, T* __restrict__ dst, int dst_Cp,
                                                           int dst_c0, int vecs, long long total) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vecs;
    const int v = (int)(i - r * vecs);
    float x[8];
    load8(src + r * src_Cp + src_c0 + v * 8, x);
    if (addend) {
      float a[8];
      load8(addend + r * dst_Cp + dst_c0 + v * 8, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] += a[e];
    }
    store8(dst + r * dst_Cp + dst_c0 + v * 8, x);
  }
}

// ------------------------------------------------------------------------------------------------------
// Input pipeline tail (next row N1): transforms.ToTensor() + transforms.Normalize(mean, std) of the reference's
// datasets (try_with_torch.py:310-313: mean = std = 0.5) on the GPU -- uint8 HWC pixels in, fp32 NCHW planes out,
// the exact fp32 operations of torchvision: t = u / 255;  y = (t - mean[c]) / std[c].  The host then ships one byte
// per sample instead of four.
// ------------------------------------------------------------------------------------------------------
struct U8Norm {
  float mean[4], stdv[4];
};

__global__ void __launch_bounds__(256) image_u8_to_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                               long long npix_total, int HW, int C, U8Norm nm) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix_total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const uint8_t* s = src + i * C;
    float* d = dst + n * (long long)C * HW + p;
    for (int c = 0; c < C; ++c) {
      const float t = __fdiv_rn((float)s[c], 255.f);
      d[(long long)c * HW] = __fdiv_rn(__fsub_rn(t, nm.mean[c]), nm.stdv[c]);
    }
  }
}

}  // namespace hg

using namespace hg;

#define HG_DISPATCH_T(dtype, CALL)                    \
  do {                                                \
    if ((dtype) == HG_BF16) {                         \
      typedef __nv_bfloat16 T;                        \
      CALL;                                           \
    } else {                                          \
      typedef float T;                                \
      CALL;                                           \
    }                                                 \
  } while (0)

extern "C" {

static int check_spatial(int dtype, int N, int H, int W, int C, const char* who) {
  HG_REQUIRE(dtype == HG_BF16 || dtype == HG_F32, "%s: bad dtype", who);
  HG_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "%s: non-positive size", who);
  return HG_OK;
}

int hg_maxpool2_fwd(int dtype, const void* x, int N, int H, int W, int C, void* y, float* stats, void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_maxpool2_fwd");
  if (rc) return rc;
  HG_REQUIRE(x && y, "hg_maxpool2_fwd: NULL pointer");
  HG_REQUIRE(H % 2 == 0 && W % 2 == 0, "hg_maxpool2_fwd: odd spatial size %dx%d", H, W);
  const int Cp = (C + 63) & ~63;
  const long long total = (long long)N * (H / 2) * (W / 2) * (Cp / 8);
  HG_DISPATCH_T(dtype, (launch_k(maxpool2_fwd_kernel<T>, dim3(grid_for(total, stats != nullptr)), dim3(256), 0,
                                 (cudaStream_t)stream, (const T*)x, (T*)y, N, H, W, Cp, stats)));
  HG_LAUNCH_OK("maxpool2_fwd_kernel");
  count_launch();
  return HG_OK;
}

int hg_maxpool2_bwd(int dtype, const void* x, const void* dy, const void* addend, int N, int H, int W, int C,
                    void* dx, void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_maxpool2_bwd");
  if (rc) return rc;
  HG_REQUIRE(x && dy && dx, "hg_maxpool2_bwd: NULL pointer");
  HG_REQUIRE(H % 2 == 0 && W % 2 == 0, "hg_maxpool2_bwd: odd spatial size %dx%d", H, W);
  const int Cp = (C + 63) & ~63;
  const long long total = (long long)N * (H / 2) * (W / 2) * (Cp / 8);
  HG_DISPATCH_T(dtype, (launch_k(maxpool2_bwd_kernel<T>, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, 
                           (const T*)x, (const T*)dy, (const T*)addend, (T*)dx, N, H, W, Cp)));
  HG_LAUNCH_OK("maxpool2_bwd_kernel");
  count_launch();
  return HG_OK;
}

constexpr size_t kUpSepMaxBytes = 96 * 1024;
static bool up_sep_attr_set = false;
static cudaError_t up_sep_set_attrs() {
  cudaError_t e = cudaSuccess;
#define HG_UP_ATTR(K)                                                                                              \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpSepMaxBytes)
  HG_UP_ATTR(upsample2_add_fwd_sep_kernel<__nv_bfloat16>);
  HG_UP_ATTR(upsample2_add_fwd_sep_kernel<float>);
  HG_UP_ATTR(upsample2_bwd_sep_kernel<__nv_bfloat16>);
  HG_UP_ATTR(upsample2_bwd_sep_kernel<float>);
#undef HG_UP_ATTR
  return e;
}

int hg_upsample2x_add_fwd(int dtype, int mode, const void* low, const void* skip, int N, int h, int w, int C,
                          void* out, float* stats, void* stream) {
  int rc = check_spatial(dtype, N, h, w, C, "hg_upsample2x_add_fwd");
  if (rc) return rc;
  HG_REQUIRE(low && out, "hg_upsample2x_add_fwd: NULL pointer");
  HG_REQUIRE(mode == 0 || mode == 1, "hg_upsample2x_add_fwd: mode must be 0 (bilinear_ac) or 1 (nearest)");
  const int Cp = (C + 63) & ~63;
  HG_REQUIRE(256 % (Cp / 8) == 0, "hg_upsample2x_add_fwd: padded channel count %d unsupported", Cp);
  // one block per output row, capped (statistics: every block ends with 2*Cp/4 vector atomics into the same lines)
  int rows_grid = N * 2 * h;
  const int cap = kNumSMs * (g_upsample_fwd_cap > 0 ? g_upsample_fwd_cap : (stats != nullptr ? 4 : 8));
  if (rows_grid > cap) rows_grid = cap;
  const size_t sep_bytes = (size_t)w * Cp * sizeof(float);
  if (g_upsample_sep && sep_bytes <= kUpSepMaxBytes) {
    // separable form: one vertically blended low-resolution row in shared memory per output row
    if (!up_sep_attr_set) {
      HG_CUDA_OK(up_sep_set_attrs());
      up_sep_attr_set = true;
    }
    HG_DISPATCH_T(dtype, (launch_k(upsample2_add_fwd_sep_kernel<T>, dim3(rows_grid), dim3(256), sep_bytes,
                                   (cudaStream_t)stream, (const T*)low, (const T*)skip, (T*)out, N, h, w, Cp, mode,
                                   stats)));
  } else {
    HG_DISPATCH_T(dtype, (launch_k(upsample2_add_fwd_kernel<T>, dim3(rows_grid), dim3(256), 0,
                                   (cudaStream_t)stream, (const T*)low, (const T*)skip, (T*)out, N, h, w, Cp, mode,
                                   stats)));
  }
  HG_LAUNCH_OK("upsample2_add_fwd_kernel");
  count_launch();
  return HG_OK;
}

int hg_upsample2x_bwd(int dtype, int mode, const void* dout, const void* addend, int N, int h, int w, int C,
                      void* dlow, void* stream) {
  int rc = check_spatial(dtype, N, h, w, C, "hg_upsample2x_bwd");
  if (rc) return rc;
  HG_REQUIRE(dout && dlow, "hg_upsample2x_bwd: NULL pointer");
  HG_REQUIRE(mode == 0 || mode == 1, "hg_upsample2x_bwd: mode must be 0 (bilinear_ac) or 1 (nearest)");
  const int Cp = (C + 63) & ~63;
  HG_REQUIRE(256 % (Cp / 8) == 0, "hg_upsample2x_bwd: padded channel count %d unsupported", Cp);
  int rows_grid = N * h;
  if (rows_grid > kNumSMs * 8) rows_grid = kNumSMs * 8;
  const size_t sep_bytes = (size_t)2 * w * Cp * sizeof(float);
  if (g_upsample_sep && sep_bytes <= kUpSepMaxBytes) {
    if (!up_sep_attr_set) {
      HG_CUDA_OK(up_sep_set_attrs());
      up_sep_attr_set = true;
    }
    if (rows_grid > kNumSMs * 3) rows_grid = kNumSMs * 3;   // 64 KB of shared memory per block at 64 x 256
    HG_DISPATCH_T(dtype, (launch_k(upsample2_bwd_sep_kernel<T>, dim3(rows_grid), dim3(256), sep_bytes,
                                   (cudaStream_t)stream, (const T*)dout, (const T*)addend, (T*)dlow, N, h, w, Cp, mode)));
  } else {
    HG_DISPATCH_T(dtype, (launch_k(upsample2_bwd_kernel<T>, dim3(rows_grid), dim3(256), 0, (cudaStream_t)stream,
                                   (const T*)dout, (const T*)addend, (T*)dlow, N, h, w, Cp, mode)));
  }
  HG_LAUNCH_OK("upsample2_bwd_kernel");
  count_launch();
  return HG_OK;
}

int hg_add(int dtype, const void* a, const void* b, void* out, long long n_elems, void* stream) {
  HG_REQUIRE(dtype == HG_BF16 || dtype == HG_F32, "hg_add: bad dtype");
  HG_REQUIRE(a && b && out && n_elems > 0 && n_elems % 8 == 0, "hg_add: bad arguments");
  const long long nvec = n_elems / 8;
  HG_DISPATCH_T(dtype, (launch_k(add_kernel<T>, dim3(grid_for(nvec)), dim3(256), 0, (cudaStream_t)stream, (const T*)a, (const T*)b,
                                                                                        (T*)out, nvec)));
  HG_LAUNCH_OK("add_kernel");
  count_launch();
  return HG_OK;
}

int hg_nchw_f32_to_nhwc(int dtype, const float* src_nchw, const void* addend, int N, int C, int H, int W, void* dst,
                        void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_nchw_f32_to_nhwc");
  if (rc) return rc;
  HG_REQUIRE(dst != nullptr, "hg_nchw_f32_to_nhwc: NULL dst");
  const int Cp = (C + 63) & ~63;
  const long long total = (long long)N * H * W * (Cp / 8);
  HG_DISPATCH_T(dtype, (launch_k(nchw_to_nhwc_kernel<T>, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, 
                           src_nchw, (const T*)addend, (T*)dst, N, C, H * W, Cp)));
  HG_LAUNCH_OK("nchw_to_nhwc_kernel");
  count_launch();
  return HG_OK;
}

int hg_nhwc_to_nchw_f32(int dtype, const void* src, int N, int C, int H, int W, float* dst_nchw, void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_nhwc_to_nchw_f32");
  if (rc) return rc;
  HG_REQUIRE(src && dst_nchw, "hg_nhwc_to_nchw_f32: NULL pointer");
  const int Cp = (C + 63) & ~63;
  const long long total = (long long)N * C * H * W;
  HG_DISPATCH_T(dtype, (launch_k(nhwc_to_nchw_kernel<T>, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, 
                           (const T*)src, dst_nchw, N, C, H * W, Cp)));
  HG_LAUNCH_OK("nhwc_to_nchw_kernel");
  count_launch();
  return HG_OK;
}

int hg_image_u8_to_nchw_f32(const uint8_t* src_nhwc, int N, int H, int W, int C, const float* mean_host,
                            const float* std_host, float* dst_nchw, void* stream) {
  HG_REQUIRE(src_nhwc && dst_nchw && mean_host && std_host, "hg_image_u8_to_nchw_f32: NULL pointer");
  HG_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 4, "hg_image_u8_to_nchw_f32: 1..4 channels, positive sizes");
  U8Norm nm;
  for (int c = 0; c < 4; ++c) {
    nm.mean[c] = c < C ? mean_host[c] : 0.f;
    nm.stdv[c] = c < C ? std_host[c] : 1.f;
    HG_REQUIRE(nm.stdv[c] != 0.f, "hg_image_u8_to_nchw_f32: std must be non-zero");
  }
  const long long npix = (long long)N * H * W;
  long long blocks = (npix + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  launch_k(image_u8_to_nchw_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src_nhwc, dst_nchw, npix,
           H * W, C, nm);
  HG_LAUNCH_OK("image_u8_to_nchw_kernel");
  count_launch();
  return HG_OK;
}

int hg_spatial_mean(int dtype, const void* x, int N, int H, int W, int C, float scale, const void* addend, void* y,
                    void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_spatial_mean");
  if (rc) return rc;
  HG_REQUIRE(x && y, "hg_spatial_mean: NULL pointer");
  const int Cp = (C + 63) & ~63;
  dim3 grid((Cp / 8 + 31) / 32, N);
  HG_DISPATCH_T(dtype, (launch_k(spatial_sum_kernel<T>, grid, dim3(256), 0, (cudaStream_t)stream, (const T*)x,
                                 (const T*)addend, (T*)y, H * W, Cp, scale)));
  HG_LAUNCH_OK("spatial_sum_kernel");
  count_launch();
  return HG_OK;
}

int hg_spatial_broadcast(int dtype, const void* y, int N, int H, int W, int C, float scale, const void* addend,
                         void* x, void* stream) {
  int rc = check_spatial(dtype, N, H, W, C, "hg_spatial_broadcast");
  if (rc) return rc;
  HG_REQUIRE(x && y, "hg_spatial_broadcast: NULL pointer");
  const int Cp = (C + 63) & ~63;
  const long long total = (long long)N * H * W * (Cp / 8);
  HG_DISPATCH_T(dtype, (launch_k(spatial_bcast_kernel<T>, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream,
                                 (const T*)y, (const T*)addend, (T*)x, total, H * W, Cp, scale)));
  HG_LAUNCH_OK("spatial_bcast_kernel");
  count_launch();
  return HG_OK;
}

int hg_channel_copy(int dtype, const void* src, int src_channels, int src_c0, const void* addend, void* dst,
                    int dst_channels, int dst_c0, int channels, long long rows, void* stream) {
  HG_REQUIRE(dtype == HG_BF16 || dtype == HG_F32, "hg_channel_copy: bad dtype");
  HG_REQUIRE(src && dst && rows > 0 && channels > 0, "hg_channel_copy: bad arguments");
  const int sCp = (src_channels + 63) & ~63, dCp = (dst_channels + 63) & ~63;
  const int vecs = (channels + 7) / 8;
  HG_REQUIRE(src_c0 % 8 == 0 && dst_c0 % 8 == 0, "hg_channel_copy: channel offsets must be multiples of 8");
  HG_REQUIRE(src_c0 + vecs * 8 <= sCp && dst_c0 + vecs * 8 <= dCp, "hg_channel_copy: window exceeds the padded tensor");
  const long long total = rows * vecs;
  HG_DISPATCH_T(dtype, (launch_k(channel_copy_kernel<T>, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream,
                                 (const T*)src, sCp, src_c0, (const T*)addend, (T*)dst, dCp, dst_c0, vecs, total)));
  HG_LAUNCH_OK("channel_copy_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
