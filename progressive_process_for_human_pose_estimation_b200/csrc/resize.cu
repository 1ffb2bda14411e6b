// Input pipeline (next row N1): PIL's `image.resize([256, 256])` of the reference datasets (try_with_torch.py:99,
// default filter = BICUBIC) for a whole batch of variable-size RGB images on the GPU, bit-exact with Pillow's
// libImaging/Resample.c for 8-bit images:
//   * two passes, horizontal first, the intermediate image rounded to uint8;
//   * per output pixel  clip8((2^21 + sum_x in[xmin + x] * k[x]) >> 22)  with the 22-bit fixed-point coefficients of
//     normalize_coeffs_8bpc.  The coefficient / bounds tables depend only on (input size, output size); the host
//     computes them in float64 in Pillow's operation order (progressive_..._b200/preprocess.py) and caches them.
// Integer arithmetic only on the device: exactness does not depend on floating-point contraction rules.
#include "hg_common.cuh"

namespace hg {

__device__ __forceinline__ int clip8(int v) {
  v >>= 22;   // arithmetic shift, like Pillow's clip8_lookups[in >> PRECISION_BITS]
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// tmp[b][y][xo][c] = horizontal pass of src[b][y][*][c]
__global__ void __launch_bounds__(256) resize_h_kernel(const HgResizeImage* __restrict__ imgs, const int* __restrict__ coef,
                                                       const int* __restrict__ bounds, uint8_t* __restrict__ tmp,
                                                       int out_w, int max_h) {
  pdl_wait();
  pdl_trigger();
  const HgResizeImage im = imgs[blockIdx.y];
  const long long total = (long long)im.h * out_w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / out_w), xo = (int)(i - (long long)y * out_w);
    const int xmin = bounds[im.bx_off + 2 * xo], xmax = bounds[im.bx_off + 2 * xo + 1];
    const int* k = coef + im.kx_off + (long long)xo * im.ksize_x;
    const uint8_t* row = im.src + ((long long)y * im.w + xmin) * 3;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int x = 0; x < xmax; ++x) {
      const int kk = k[x];
      s0 += row[3 * x] * kk;
      s1 += row[3 * x + 1] * kk;
      s2 += row[3 * x + 2] * kk;
    }
    uint8_t* o = tmp + im.tmp_off + ((long long)y * out_w + xo) * 3;
    o[0] = (uint8_t)clip8(s0);
    o[1] = (uint8_t)clip8(s1);
    o[2] = (uint8_t)clip8(s2);
  }
}

// out[b][yo][xo][c] = vertical pass of tmp[b][*][xo][c];  optionally also ToTensor + Normalize into fp32 NCHW
__global__ void __launch_bounds__(256) resize_v_kernel(const HgResizeImage* __restrict__ imgs, const int* __restrict__ coef,
                                                       const int* __restrict__ bounds, const uint8_t* __restrict__ tmp,
                                                       uint8_t* __restrict__ out, float* __restrict__ out_norm, int out_w,
                                                       int out_h, float m0, float m1, float m2, float d0, float d1,
                                                       float d2) {
  pdl_wait();
  pdl_trigger();
  const HgResizeImage im = imgs[blockIdx.y];
  const int total = out_h * out_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int yo = i / out_w, xo = i - yo * out_w;
    const int ymin = bounds[im.by_off + 2 * yo], ymax = bounds[im.by_off + 2 * yo + 1];
    const int* k = coef + im.ky_off + (long long)yo * im.ksize_y;
    const uint8_t* col = tmp + im.tmp_off + ((long long)ymin * out_w + xo) * 3;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int y = 0; y < ymax; ++y) {
      const int kk = k[y];
      const uint8_t* p = col + (long long)y * out_w * 3;
      s0 += p[0] * kk;
      s1 += p[1] * kk;
      s2 += p[2] * kk;
    }
    const int r = clip8(s0), g = clip8(s1), b = clip8(s2);
    const long long pix = (long long)blockIdx.y * total + i;
    if (out) {
      out[pix * 3] = (uint8_t)r;
      out[pix * 3 + 1] = (uint8_t)g;
      out[pix * 3 + 2] = (uint8_t)b;
    }
    if (out_norm) {   // torchvision: t = u / 255; (t - mean) / std, fp32
      float* q = out_norm + (long long)blockIdx.y * 3 * total + i;
      q[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)r, 255.f), m0), d0);
      q[total] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)g, 255.f), m1), d1);
      q[2 * (long long)total] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)b, 255.f), m2), d2);
    }
  }
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_resize_bicubic_u8(const HgResizeImage* images_dev, int num_images, int max_h, int out_w, int out_h,
                         const int32_t* coef_dev, const int32_t* bounds_dev, uint8_t* tmp, uint8_t* out_nhwc,
                         float* out_nchw_norm, const float* mean_host, const float* std_host, void* stream) {
  HG_REQUIRE(images_dev && coef_dev && bounds_dev && tmp, "hg_resize_bicubic_u8: NULL pointer");
  HG_REQUIRE(out_nhwc || out_nchw_norm, "hg_resize_bicubic_u8: no output requested");
  HG_REQUIRE(num_images > 0 && max_h > 0 && out_w > 0 && out_h > 0, "hg_resize_bicubic_u8: non-positive size");
  HG_REQUIRE(!out_nchw_norm || (mean_host && std_host), "hg_resize_bicubic_u8: mean / std missing");
  float m[3] = {0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
  if (out_nchw_norm)
    for (int c = 0; c < 3; ++c) {
      m[c] = mean_host[c];
      s[c] = std_host[c];
      HG_REQUIRE(s[c] != 0.f, "hg_resize_bicubic_u8: std must be non-zero");
    }
  long long bh = ((long long)max_h * out_w + 255) / 256;
  if (bh > 4 * kNumSMs) bh = 4 * kNumSMs;
  launch_k(resize_h_kernel, dim3((unsigned)bh, (unsigned)num_images), dim3(256), 0, (cudaStream_t)stream, images_dev,
           (const int*)coef_dev, (const int*)bounds_dev, tmp, out_w, max_h);
  HG_LAUNCH_OK("resize_h_kernel");
  count_launch();
  int bv = (out_h * out_w + 255) / 256;
  launch_k(resize_v_kernel, dim3((unsigned)bv, (unsigned)num_images), dim3(256), 0, (cudaStream_t)stream, images_dev,
           (const int*)coef_dev, (const int*)bounds_dev, (const uint8_t*)tmp, out_nhwc, out_nchw_norm, out_w, out_h, m[0],
           m[1], m[2], s[0], s[1], s[2]);
  HG_LAUNCH_OK("resize_v_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
