// Heatmap decoding (argmax with lowest-index tie-break) and PCKh counting, fused: the heatmaps are read once
// from HBM, one block per (image, joint), warp-shuffle arg-reductions, integer counters.
//
// Replaces the per-joint torch.max / torch.nonzero / python loops of
//   PCKh "A"  only_one_hourgless.py:285-313 (= try_with_torch_100.py:283-311)      -> hg_pckh_a
//   PCKh "C"  hourglass_compare.py:812-844, performance_compare.py:581-615          -> hg_pckh_sweep (chan_offset 0)
//   PCKh "B"  train.py:759-791, performance_compare.py:544-578                      -> hg_pckh_sweep (chan_offset 1)
//   decode    hourglass_compare.py:831,1092; read_mscoco.py:81                      -> hg_decode_argmax
// Float arithmetic that decides a count (sqrtf, division, float32 thresholds) is IEEE-exact, no fast math.
#include <cuda_fp16.h>

#include "hg_common.cuh"

namespace hg {

__device__ __forceinline__ float ld_val(const void* p, int dtype, long long i) {
  if (dtype == HG_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == HG_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// first row-major index of the maximum of n values (block-wide). NaN is never "greater", like torch.max's
// comparison on well-formed heatmaps. Result valid in every thread.
template <class F>
__device__ int block_argmax_f(F value, int n, float* out_max) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = value(i);
    if (v > bv || (v == bv && i < bi)) {
      bv = v;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) {
    s_val[warp] = bv;
    s_idx[warp] = bi;
  }
  __syncthreads();
  bv = s_val[0];
  bi = s_idx[0];
  for (int w = 1; w < nw; ++w) {
    const float ov = s_val[w];
    const int oi = s_idx[w];
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  if (out_max) *out_max = bv;
  if (bi == 0x7fffffff) bi = 0;  // all-NaN map
  return bi;
}

__device__ int block_argmax(const void* base, int dtype, long long off, int n, float* out_max) {
  return block_argmax_f([=](int i) { return ld_val(base, dtype, off + i); }, n, out_max);
}

// Channel softmax evaluated on the fly (fused `softmax(result[2])` -> PCKh of hourglass_compare.py:1160 /
// performance_compare.py:646-647): ms[pixel] = {max_c x, sum_c exp(x - max)} from softmax_stats_kernel; the value is
// exp(x - max) / sum, the expression (and channel order of the sum) of PyTorch's softmax kernels.
__device__ int block_argmax_softmax(const float* x, long long off, const float2* ms, int n, float* out_max) {
  return block_argmax_f(
      [=](int i) {
        const float2 c = ms[i];
        return __fdiv_rn(expf(x[off + i] - c.x), c.y);
      },
      n, out_max);
}

__global__ void __launch_bounds__(256) softmax_stats_kernel(const float* __restrict__ x, int C, int HW, long long npix,
                                                            float2* __restrict__ ms) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW;
    const float* p = x + b * (long long)C * HW + (i - b * HW);
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(p + (long long)c * HW));
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(__ldg(p + (long long)c * HW) - m);
    ms[i] = make_float2(m, sum);
  }
}

// first row-major index i with lab[i] == value, or -1
__device__ int block_first_equal(const long long* lab, int n, long long value) {
  __shared__ int s_first;
  if (threadIdx.x == 0) s_first = 0x7fffffff;
  __syncthreads();
  int f = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (lab[i] == value && i < f) f = i;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
  if ((threadIdx.x & 31) == 0 && f != 0x7fffffff) atomicMin(&s_first, f);
  __syncthreads();
  const int r = s_first;
  __syncthreads();
  return r == 0x7fffffff ? -1 : r;
}

__global__ void __launch_bounds__(256) decode_argmax_kernel(const void* __restrict__ hm, int dtype, int HW, int W,
                                                            int* __restrict__ out_yx, float* __restrict__ out_max) {
  const long long map = blockIdx.x;
  float mx;
  const int idx = block_argmax(hm, dtype, map * HW, HW, &mx);
  if (threadIdx.x == 0) {
    out_yx[2 * map] = idx / W;
    out_yx[2 * map + 1] = idx % W;
    if (out_max) out_max[map] = mx;
  }
}

// grid = (njoints, B)
__global__ void __launch_bounds__(256) pckh_sweep_kernel(const void* __restrict__ x, int dtype, int C, int H, int W,
                                                         const long long* __restrict__ target,
                                                         const float* __restrict__ rect, int chan_offset, int njoints,
                                                         const float* __restrict__ thr, int nthr, int rule,
                                                         int* __restrict__ correct, int* __restrict__ total,
                                                         int* __restrict__ predict_xy, int* __restrict__ label_xy,
                                                         int* __restrict__ found, float* __restrict__ standard_out,
                                                         const float2* __restrict__ ms) {
  const int j = blockIdx.x, b = blockIdx.y, HW = H * W;
  const int li = block_first_equal(target + (long long)b * HW, HW, (long long)(j + 1));
  // standard = sqrt((x1-x2)^2 + (y1-y2)^2) * 0.6, all in float32 (rect is a float32 tensor)
  const float rx = rect[4 * b] - rect[4 * b + 2], ry = rect[4 * b + 1] - rect[4 * b + 3];
  const float standard = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry))), 0.6f);
  if (j == 0 && threadIdx.x == 0 && standard_out) standard_out[b] = standard;
  if (li < 0) {
    if (threadIdx.x == 0) found[b * njoints + j] = 0;
    return;
  }
  const long long xoff = ((long long)b * C + j + chan_offset) * HW;
  const int pi = ms ? block_argmax_softmax(reinterpret_cast<const float*>(x), xoff, ms + (long long)b * HW, HW, nullptr)
                    : block_argmax(x, dtype, xoff, HW, nullptr);
  if (threadIdx.x == 0) {
    const int ly = li / W, lx = li % W, py = pi / W, px = pi % W;
    const int dy = ly - py, dx = lx - px;
    // rule 0 (PCKh B/C): sqrt(d2) / standard < k;  rule 1 (PCKh D, calculate_parameters.py:927-929):
    // sqrt(d2) < standard * k -- the two roundings differ, so each evaluator keeps its own float32 expression
    const float root = __fsqrt_rn((float)(dy * dy + dx * dx));
    const float dist = __fdiv_rn(root, standard);
    for (int s = 0; s < nthr; ++s) {
      const bool hit = rule == 0 ? dist < thr[s] : root < __fmul_rn(standard, thr[s]);
      if (hit) atomicAdd(correct + b * nthr + s, 1);
      atomicAdd(total + b * nthr + s, 1);
    }
    predict_xy[(b * njoints + j) * 2] = px;
    predict_xy[(b * njoints + j) * 2 + 1] = py;
    label_xy[(b * njoints + j) * 2] = lx;
    label_xy[(b * njoints + j) * 2 + 1] = ly;
    found[b * njoints + j] = 1;
  }
}

// first index of the maximum of one row
__device__ int row_argmax(const void* base, int dtype, long long off, int W) {
  float bv = ld_val(base, dtype, off);
  int bi = 0;
  for (int i = 1; i < W; ++i) {
    const float v = ld_val(base, dtype, off + i);
    if (v > bv) {
      bv = v;
      bi = i;
    }
  }
  return bi;
}

// PCKh "A" (quirk Q8 kept: both x coordinates come from the LABEL map at row head_ys, so the x error is 0).
// grid = B images; counts[0] += correct, counts[1] += total
__global__ void __launch_bounds__(256) pckh_a_kernel(const void* __restrict__ x, int xdtype,
                                                     const void* __restrict__ target, int tdtype, int Cx, int Ct, int H,
                                                     int W, int njoints, int head_ch, int neck_ch,
                                                     int* __restrict__ counts) {
  const int b = blockIdx.x, HW = H * W;
  const int hi = block_argmax(target, tdtype, ((long long)b * Ct + head_ch) * HW, HW, nullptr);
  const int ni = block_argmax(target, tdtype, ((long long)b * Ct + neck_ch) * HW, HW, nullptr);
  const int head_ys = hi / W, head_xs = hi % W, neck_ys = ni / W, neck_xs = ni % W;
  const int d2 = (head_ys - neck_ys) * (head_ys - neck_ys) + (head_xs - neck_xs) * (head_xs - neck_xs);
  const float standard = __fdiv_rn(__fsqrt_rn((float)d2), 2.0f);
  int correct = 0, total = 0;
  for (int j = 0; j < njoints; ++j) {
    float lmax;
    const int li = block_argmax(target, tdtype, ((long long)b * Ct + j) * HW, HW, &lmax);
    if (lmax == 0.f) continue;
    const int pi = block_argmax(x, xdtype, ((long long)b * Cx + j) * HW, HW, nullptr);
    if (threadIdx.x == 0) {
      const int label_ys = li / W, predict_ys = pi / W;
      const int label_xs = row_argmax(target, tdtype, ((long long)b * Ct + j) * HW + (long long)head_ys * W, W);
      const int predict_xs = label_xs;
      const int e2 = (label_ys - predict_ys) * (label_ys - predict_ys) + (label_xs - predict_xs) * (label_xs - predict_xs);
      if (__fsqrt_rn((float)e2) < standard) ++correct;
      ++total;
    }
  }
  if (threadIdx.x == 0) {
    atomicAdd(counts, correct);
    atomicAdd(counts + 1, total);
  }
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_decode_argmax(const void* heatmaps, int dtype, int num_maps, int H, int W, int32_t* out_yx, float* out_max,
                     void* stream) {
  HG_REQUIRE(heatmaps && out_yx, "hg_decode_argmax: NULL pointer");
  HG_REQUIRE(dtype >= 0 && dtype <= 2, "hg_decode_argmax: dtype must be HG_BF16, HG_F32 or HG_F16");
  HG_REQUIRE(num_maps > 0 && H > 0 && W > 0, "hg_decode_argmax: non-positive size");
  decode_argmax_kernel<<<num_maps, 256, 0, (cudaStream_t)stream>>>(heatmaps, dtype, H * W, W, out_yx, out_max);
  HG_LAUNCH_OK("decode_argmax_kernel");
  count_launch();
  return HG_OK;
}

static int pckh_launch(const char* who, int rule, const float2* ms, const void* x, int dtype, int B, int C, int H, int W,
                       const int64_t* target, const float* rect, int chan_offset, int njoints, const float* thresholds,
                       int nthr, int32_t* correct, int32_t* total, int32_t* predict_xy, int32_t* label_xy,
                       int32_t* found, float* standard, void* stream) {
  HG_REQUIRE(x && target && rect && thresholds && correct && total && predict_xy && label_xy && found,
             "%s: NULL pointer", who);
  HG_REQUIRE(dtype >= 0 && dtype <= 2, "%s: bad dtype", who);
  HG_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && njoints > 0 && nthr > 0, "%s: non-positive size", who);
  HG_REQUIRE(chan_offset >= 0 && njoints + chan_offset <= C, "%s: joints exceed channels", who);
  dim3 grid(njoints, B);
  pckh_sweep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, dtype, C, H, W, (const long long*)target, rect,
                                                             chan_offset, njoints, thresholds, nthr, rule, correct,
                                                             total, predict_xy, label_xy, found, standard, ms);
  HG_LAUNCH_OK("pckh_sweep_kernel");
  count_launch();
  return HG_OK;
}

int hg_pckh_sweep(const void* x, int dtype, int B, int C, int H, int W, const int64_t* target, const float* rect,
                  int chan_offset, int njoints, const float* thresholds, int nthr, int32_t* correct, int32_t* total,
                  int32_t* predict_xy, int32_t* label_xy, int32_t* found, float* standard, void* stream) {
  return pckh_launch("hg_pckh_sweep", 0, nullptr, x, dtype, B, C, H, W, target, rect, chan_offset, njoints, thresholds, nthr,
                     correct, total, predict_xy, label_xy, found, standard, stream);
}

int hg_pckh_abs(const void* x, int dtype, int B, int C, int H, int W, const int64_t* target, const float* rect,
                int chan_offset, int njoints, const float* factors, int nfac, int32_t* correct, int32_t* total,
                int32_t* predict_xy, int32_t* label_xy, int32_t* found, float* standard, void* stream) {
  return pckh_launch("hg_pckh_abs", 1, nullptr, x, dtype, B, C, H, W, target, rect, chan_offset, njoints, factors, nfac, correct,
                     total, predict_xy, label_xy, found, standard, stream);
}

int hg_softmax_stats(const float* logits, int B, int C, int H, int W, float* max_sum, void* stream) {
  HG_REQUIRE(logits && max_sum, "hg_softmax_stats: NULL pointer");
  HG_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "hg_softmax_stats: non-positive size");
  HG_REQUIRE((reinterpret_cast<uintptr_t>(max_sum) & 7) == 0, "hg_softmax_stats: max_sum must be 8-byte aligned");
  const long long npix = (long long)B * H * W;
  long long blocks = (npix + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  softmax_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(logits, C, H * W, npix,
                                                                            reinterpret_cast<float2*>(max_sum));
  HG_LAUNCH_OK("softmax_stats_kernel");
  count_launch();
  return HG_OK;
}

int hg_pckh_logits(const float* logits, const float* max_sum, int absolute, int B, int C, int H, int W,
                   const int64_t* target, const float* rect, int chan_offset, int njoints, const float* thresholds,
                   int nthr, int32_t* correct, int32_t* total, int32_t* predict_xy, int32_t* label_xy, int32_t* found,
                   float* standard, void* stream) {
  HG_REQUIRE(max_sum != nullptr, "hg_pckh_logits: max_sum (hg_softmax_stats) is NULL");
  return pckh_launch("hg_pckh_logits", absolute ? 1 : 0, reinterpret_cast<const float2*>(max_sum), logits, HG_F32, B, C,
                     H, W, target, rect, chan_offset, njoints, thresholds, nthr, correct, total, predict_xy, label_xy,
                     found, standard, stream);
}

int hg_pckh_a(const void* x, int x_dtype, const void* target, int t_dtype, int B, int Cx, int Ct, int H, int W,
              int njoints, int head_ch, int neck_ch, int32_t* counts, void* stream) {
  HG_REQUIRE(x && target && counts, "hg_pckh_a: NULL pointer");
  HG_REQUIRE(x_dtype >= 0 && x_dtype <= 2 && t_dtype >= 0 && t_dtype <= 2, "hg_pckh_a: bad dtype");
  HG_REQUIRE(B > 0 && H > 0 && W > 0 && njoints > 0, "hg_pckh_a: non-positive size");
  HG_REQUIRE(njoints <= Cx && njoints <= Ct && head_ch < Ct && neck_ch < Ct, "hg_pckh_a: channel out of range");
  pckh_a_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, x_dtype, target, t_dtype, Cx, Ct, H, W, njoints, head_ch,
                                                      neck_ch, counts);
  HG_LAUNCH_OK("pckh_a_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
