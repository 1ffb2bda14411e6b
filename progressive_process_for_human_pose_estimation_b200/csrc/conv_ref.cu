// CUDA-core (FFMA) convolution kernels: the fp32 path of the library (rtol 1e-5 against the reference) and
// the on-device cross-check for the tcgen05 kernels.  General stride / padding / dilation, NHWC, channels
// padded to a multiple of 64 like everywhere else.  Plain smem-tiled GEMMs, fp32 accumulation.
//
// Replaces nn.Conv2d forward/backward (reference try_with_torch.py:186-193,199-207) when activations are
// fp32, and every geometry the tensor-core kernels do not take.
#include "hg_common.cuh"

namespace hg {

struct ConvRefParams {
  int N, H, W;        // input spatial (of x)
  int Ho, Wo;         // output spatial (of y)
  int Cin_p, Cout_p;  // padded channels
  int R, S, stride, pad, dil;
  int c_real;
};

// ------------------------------------------------------------------------------------------------------
// fprop (MODE 0): out[m=(n,ho,wo), co] = sum x[n, ho*s - pad + r*dil, wo*s - pad + s*dil, ci] * w[tap][co][ci]
// dgrad (MODE 1): out[m=(n,hi,wi), ci] = sum dy[n, (hi + pad - r*dil)/s, (wi + pad - s*dil)/s, co] * w[tap][ci][co]
//   (w is the matching packed layout: [tap][Nout][Kin])
// ------------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) conv_ref_kernel(ConvRefParams p, const T* __restrict__ act,
                                                       const T* __restrict__ w, const float* __restrict__ bias,
                                                       const T* __restrict__ res, T* __restrict__ out,
                                                       float* __restrict__ out_nchw) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  // geometry of the gathered (A) tensor and of the output tensor
  const int aH = MODE == 0 ? p.H : p.Ho, aW = MODE == 0 ? p.W : p.Wo;
  const int oH = MODE == 0 ? p.Ho : p.H, oW = MODE == 0 ? p.Wo : p.W;
  const int Kp = MODE == 0 ? p.Cin_p : p.Cout_p;
  const int Np = MODE == 0 ? p.Cout_p : p.Cin_p;
  const long long M = (long long)p.N * oH * oW;
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int lp = t >> 2;          // pixel / out-channel row this thread loads
  const int lk = (t & 3) * 4;     // 4 consecutive K elements
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};

  const long long lm = m0 + lp;
  int ln = 0, lh = 0, lw = 0;
  const bool lm_ok = lm < M;
  if (lm_ok) {
    ln = (int)(lm / (oH * oW));
    const int rem = (int)(lm - (long long)ln * oH * oW);
    lh = rem / oW;
    lw = rem - lh * oW;
  }
  for (int r = 0; r < p.R; ++r) {
    for (int s = 0; s < p.S; ++s) {
      int ah, aw;
      bool ok = lm_ok;
      if (MODE == 0) {
        ah = lh * p.stride - p.pad + r * p.dil;
        aw = lw * p.stride - p.pad + s * p.dil;
      } else {
        const int th = lh + p.pad - r * p.dil, tw = lw + p.pad - s * p.dil;
        ok = ok && (th % p.stride == 0) && (tw % p.stride == 0) && th >= 0 && tw >= 0;
        ah = th / p.stride;
        aw = tw / p.stride;
      }
      ok = ok && ah >= 0 && ah < aH && aw >= 0 && aw < aW;
      const T* arow = act + (((long long)ln * aH + ah) * aW + aw) * Kp;
      const T* wrow = w + ((long long)(r * p.S + s) * Np + (n0 + lp)) * Kp;
      for (int k0 = 0; k0 < Kp; k0 += 16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          As[lk + e][lp] = ok ? to_f(arow[k0 + lk + e]) : 0.f;
          Bs[lk + e][lp] = (n0 + lp < Np) ? to_f(wrow[k0 + lk + e]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= Np) continue;
      float v = acc[i][j] + (bias ? bias[c] : 0.f);
      if (res) v += to_f(res[m * Np + c]);
      out[m * Np + c] = from_f<T>(v);
      if (out_nchw && c < p.c_real) {
        const int plane = oH * oW;
        const long long n = m / plane;
        out_nchw[(n * p.c_real + c) * plane + (m - n * plane)] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// wgrad: dw[tap][co][ci] (GEMM layout, padded) += sum_m dy[m, co] * x[m (+) tap, ci];  split over pixels, fp32 atomics.
// grid = (ksplit, taps, (Cout_p/64)*(Cin_p/64))
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_ref_kernel(ConvRefParams p, const T* __restrict__ x,
                                                             const T* __restrict__ dy, float* __restrict__ dw,
                                                             int Cin, int Cout, int pix_per_block) {
  __shared__ float Ds[16][68];  // dy  [pixel][co]
  __shared__ float Xs[16][68];  // x   [pixel][ci]
  const int t = threadIdx.x;
  const int tap = blockIdx.y, r = tap / p.S, s = tap - r * p.S;
  const int nci = p.Cin_p / 64;
  const int co0 = (blockIdx.z / nci) * 64, ci0 = (blockIdx.z % nci) * 64;
  const long long M = (long long)p.N * p.Ho * p.Wo;
  const long long mbeg = (long long)blockIdx.x * pix_per_block;
  long long mend = mbeg + pix_per_block;
  if (mend > M) mend = M;
  const int lp = t >> 4;        // 0..15 pixel within chunk
  const int lc = (t & 15) * 4;  // 4 consecutive channels
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (long long mb = mbeg; mb < mend; mb += 16) {
    const long long m = mb + lp;
    bool ok = m < mend;
    int n = 0, ho = 0, wo = 0;
    if (ok) {
      n = (int)(m / (p.Ho * p.Wo));
      const int rem = (int)(m - (long long)n * p.Ho * p.Wo);
      ho = rem / p.Wo;
      wo = rem - ho * p.Wo;
    }
    const int hi = ho * p.stride - p.pad + r * p.dil, wi = wo * p.stride - p.pad + s * p.dil;
    const bool xok = ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
    const T* dyp = dy + m * p.Cout_p + co0 + lc;
    const T* xp = x + (((long long)n * p.H + hi) * p.W + wi) * p.Cin_p + ci0 + lc;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      Ds[lp][lc + e] = ok ? to_f(dyp[e]) : 0.f;
      Xs[lp][lc + e] = xok ? to_f(xp[e]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ds[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci >= Cin) continue;
      atomicAdd(dw + ((long long)tap * p.Cout_p + co) * p.Cin_p + ci, acc[i][j]);
    }
  }
}

int colsum_launch(int dtype, const void* dy, long long M, int Cp, int C, float* out, cudaStream_t st);

static inline int pad64(int c) { return (c + 63) & ~63; }

static ConvRefParams make_params(const HgConvDesc* d) {
  ConvRefParams p;
  p.N = d->N;
  p.H = d->H;
  p.W = d->W;
  p.Ho = (d->H + 2 * d->pad - d->dil * (d->R - 1) - 1) / d->stride + 1;
  p.Wo = (d->W + 2 * d->pad - d->dil * (d->S - 1) - 1) / d->stride + 1;
  p.Cin_p = pad64(d->Cin);
  p.Cout_p = pad64(d->Cout);
  p.R = d->R;
  p.S = d->S;
  p.stride = d->stride;
  p.pad = d->pad;
  p.dil = d->dil;
  p.c_real = d->Cout;
  return p;
}

template <typename T>
int conv_ref_fprop(const HgConvDesc* d, const void* x, const void* w, const float* bias, const void* res,
                   void* y, float* out_nchw, cudaStream_t st) {
  ConvRefParams p = make_params(d);
  const long long M = (long long)p.N * p.Ho * p.Wo;
  dim3 grid(ceil_div(M, 64), p.Cout_p / 64);
  conv_ref_kernel<T, 0><<<grid, 256, 0, st>>>(p, (const T*)x, (const T*)w, bias, (const T*)res, (T*)y, out_nchw);
  HG_LAUNCH_OK("conv_ref_kernel<fprop>");
  count_launch();
  return HG_OK;
}
template <typename T>
int conv_ref_dgrad(const HgConvDesc* d, const void* dy, const void* w, const void* addend, void* dx,
                   cudaStream_t st) {
  ConvRefParams p = make_params(d);
  const long long M = (long long)p.N * p.H * p.W;
  dim3 grid(ceil_div(M, 64), p.Cin_p / 64);
  conv_ref_kernel<T, 1><<<grid, 256, 0, st>>>(p, (const T*)dy, (const T*)w, nullptr, (const T*)addend, (T*)dx,
                                               nullptr);
  HG_LAUNCH_OK("conv_ref_kernel<dgrad>");
  count_launch();
  return HG_OK;
}
template <typename T>
int conv_ref_wgrad(const HgConvDesc* d, const void* x, const void* dy, float* dw, float* dbias, cudaStream_t st) {
  ConvRefParams p = make_params(d);
  const long long M = (long long)p.N * p.Ho * p.Wo;
  if (dw) {
    int ksplit = (int)((M + 2047) / 2048);
    if (ksplit > 512) ksplit = 512;
    if (ksplit < 1) ksplit = 1;
    int ppb = (int)((M + ksplit - 1) / ksplit);
    ppb = (ppb + 15) & ~15;
    ksplit = (int)((M + ppb - 1) / ppb);
    dim3 grid(ksplit, p.R * p.S, (p.Cout_p / 64) * (p.Cin_p / 64));
    conv_wgrad_ref_kernel<T><<<grid, 256, 0, st>>>(p, (const T*)x, (const T*)dy, dw, d->Cin, d->Cout, ppb);
    HG_LAUNCH_OK("conv_wgrad_ref_kernel");
    count_launch();
  }
  if (dbias) return colsum_launch(sizeof(T) == 2 ? HG_BF16 : HG_F32, dy, M, p.Cout_p, d->Cout, dbias, st);
  return HG_OK;
}

template int conv_ref_fprop<float>(const HgConvDesc*, const void*, const void*, const float*, const void*, void*,
                                   float*, cudaStream_t);
template int conv_ref_fprop<__nv_bfloat16>(const HgConvDesc*, const void*, const void*, const float*, const void*,
                                           void*, float*, cudaStream_t);
template int conv_ref_dgrad<float>(const HgConvDesc*, const void*, const void*, const void*, void*, cudaStream_t);
template int conv_ref_dgrad<__nv_bfloat16>(const HgConvDesc*, const void*, const void*, const void*, void*,
                                           cudaStream_t);
template int conv_ref_wgrad<float>(const HgConvDesc*, const void*, const void*, float*, float*, cudaStream_t);
template int conv_ref_wgrad<__nv_bfloat16>(const HgConvDesc*, const void*, const void*, float*, float*,
                                           cudaStream_t);

// ------------------------------------------------------------------------------------------------------
// weight repack: fp32 OIHW -> [tap][Cout_p][Cin_p] and [tap][Cin_p][Cout_p] in T (zero padded)
// ------------------------------------------------------------------------------------------------------
// (the source may be a slice [cin_off, cin_off + Cin) of a wider weight with cin_total input channels: the
//  "virtual concat" convolutions of try_different_stack.py:316-328 / try_with_aspp_remove_max_pool.py:239-240)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout,
                                   int Cin, int taps, int Cout_p, int Cin_p, int cin_total, int cin_off) {
  const long long total = (long long)taps * Cout_p * Cin_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int co = (int)((i / Cin_p) % Cout_p);
    const int tap = (int)(i / ((long long)Cin_p * Cout_p));
    const float v = (ci < Cin && co < Cout) ? w[((long long)co * cin_total + cin_off + ci) * taps + tap] : 0.f;
    if (wf) wf[i] = from_f<T>(v);
    if (wd) wd[((long long)tap * Cin_p + ci) * Cout_p + co] = from_f<T>(v);
  }
}

// GEMM-layout fp32 gradient [tap][Cout_p][Cin_p] -> OIHW [Cout][Cin][R][S]  (assign or accumulate)
__global__ void unpack_wgrad_kernel(const float* __restrict__ g, float* __restrict__ dw, int Cout, int Cin, int taps,
                                    int Cout_p, int Cin_p, int accumulate, int cin_total, int cin_off) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const float v = g[((long long)tap * Cout_p + co) * Cin_p + ci];
    const long long o = ((long long)co * cin_total + cin_off + ci) * taps + tap;
    dw[o] = accumulate ? dw[o] + v : v;
  }
}
int unpack_wgrad(const HgConvDesc* d, const float* g, float* dw, int accumulate, int cin_total, int cin_off,
                 cudaStream_t st) {
  const int taps = d->R * d->S;
  const long long total = (long long)d->Cout * d->Cin * taps;
  int blocks = ceil_div(total, 256);
  if (blocks > 1184) blocks = 1184;
  unpack_wgrad_kernel<<<blocks, 256, 0, st>>>(g, dw, d->Cout, d->Cin, taps, pad64(d->Cout), pad64(d->Cin), accumulate,
                                              cin_total, cin_off);
  HG_LAUNCH_OK("unpack_wgrad_kernel");
  count_launch();
  return HG_OK;
}

template <typename T>
int pack_weight(const HgConvDesc* d, const float* w, void* wf, void* wd, int cin_total, int cin_off,
                cudaStream_t st) {
  const int taps = d->R * d->S, Cout_p = pad64(d->Cout), Cin_p = pad64(d->Cin);
  const long long total = (long long)taps * Cout_p * Cin_p;
  int blocks = ceil_div(total, 256);
  if (blocks > 1184) blocks = 1184;
  pack_weight_kernel<T><<<blocks, 256, 0, st>>>(w, (T*)wf, (T*)wd, d->Cout, d->Cin, taps, Cout_p, Cin_p, cin_total,
                                                cin_off);
  HG_LAUNCH_OK("pack_weight_kernel");
  count_launch();
  return HG_OK;
}
template int pack_weight<float>(const HgConvDesc*, const float*, void*, void*, int, int, cudaStream_t);
template int pack_weight<__nv_bfloat16>(const HgConvDesc*, const float*, void*, void*, int, int, cudaStream_t);

// out[Ro, cols] (=|+=) T[Ro, Ri] * in[Ri, cols]   or, transposed,   out[Ri, cols] (=|+=) T^T * in[Ro, cols]:
// recombination of head channels (limb mix of try_skeleton_and_keypoints.py:279-298, gather-add limb maps of
// try_skeleton_from_keypoints_merge.py:296-298) folded into the head's weights and un-folded from their gradients
__global__ void mix_rows_kernel(const float* __restrict__ T, const float* __restrict__ in, float* __restrict__ out,
                                int Ro, int Ri, int cols, int transpose, int accumulate) {
  const int rows_out = transpose ? Ri : Ro, rows_in = transpose ? Ro : Ri;
  const long long total = (long long)rows_out * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols), r = (int)(i / cols);
    float acc = 0.f;
    for (int k = 0; k < rows_in; ++k) {
      const float t = transpose ? T[k * Ri + r] : T[r * Ri + k];
      if (t != 0.f) acc = fmaf(t, in[(long long)k * cols + c], acc);
    }
    out[i] = accumulate ? out[i] + acc : acc;
  }
}
int mix_rows(const float* T, const float* in, float* out, int Ro, int Ri, int cols, int transpose, int accumulate,
             cudaStream_t st) {
  const long long total = (long long)(transpose ? Ri : Ro) * cols;
  int blocks = ceil_div(total, 256);
  if (blocks > 1184) blocks = 1184;
  mix_rows_kernel<<<blocks, 256, 0, st>>>(T, in, out, Ro, Ri, cols, transpose, accumulate);
  HG_LAUNCH_OK("mix_rows_kernel");
  count_launch();
  return HG_OK;
}

}  // namespace hg
