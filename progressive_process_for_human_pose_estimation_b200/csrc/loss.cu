// Intermediate-supervision MSE heatmap loss of the stacked hourglass, all stacks in ONE pass:
//   loss[s] = mean((pred_s - target)^2),   dpred_s = 2 * gscale * (pred_s - target) / numel
// The reference evaluates nStack separate nn.MSELoss modules and sums them (try_with_torch.py:305-308,333-341):
// per stack a subtraction, a square and a mean kernel forward and as many again backward, each re-reading the same
// target.  Here the target is read once, every stack's prediction once, and the gradient the backward pass needs is
// written in the same sweep (HBM-bound: (2 S + 1) * 4 bytes per element).
//
// The loss cannot move further up into the head convolution's epilogue behind the reference API: creatModel.forward
// (try_with_torch.py:275-298) returns the heatmaps before the training loop shows it the target.
#include "hg_common.cuh"

namespace hg {

struct MseArgs {
  const float* pred[HG_MSE_MAX_STACKS];
  float* dpred[HG_MSE_MAX_STACKS];
  const float* target;
  float* loss;
  long long numel;   // per stack
  int S;
  float gscale;      // upstream gradient of every per-stack loss (1 for `sum of losses`.backward())
};

__global__ void __launch_bounds__(256) mse_multi_kernel(const MseArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[HG_MSE_MAX_STACKS][8];
  float acc[HG_MSE_MAX_STACKS];
#pragma unroll
  for (int s = 0; s < HG_MSE_MAX_STACKS; ++s) acc[s] = 0.f;
  const long long nvec = a.numel >> 2;
  const float k = 2.f * a.gscale / (float)a.numel;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 t = reinterpret_cast<const float4*>(a.target)[i];
#pragma unroll
    for (int s = 0; s < HG_MSE_MAX_STACKS; ++s) {
      if (s < a.S) {
        const float4 p = reinterpret_cast<const float4*>(a.pred[s])[i];
        const float4 d = make_float4(p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w);
        acc[s] += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (a.dpred[s]) reinterpret_cast<float4*>(a.dpred[s])[i] = make_float4(k * d.x, k * d.y, k * d.z, k * d.w);
      }
    }
  }
  // scalar tail (numel not a multiple of 4)
  for (long long i = (nvec << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.numel;
       i += (long long)gridDim.x * blockDim.x) {
    const float t = a.target[i];
#pragma unroll
    for (int s = 0; s < HG_MSE_MAX_STACKS; ++s) {
      if (s < a.S) {
        const float d = a.pred[s][i] - t;
        acc[s] += d * d;
        if (a.dpred[s]) a.dpred[s][i] = k * d;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < HG_MSE_MAX_STACKS; ++s) {
    if (s < a.S) {
      const float v = warp_sum(acc[s]);
      if (lane == 0) red[s][warp] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < a.S) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(a.loss + threadIdx.x, v / (float)a.numel);
  }
}


// ------------------------------------------------------------------------------------------------------
// Per-pixel class cross-entropy heads (nn.CrossEntropyLoss on [B,C,64,64] logits with [B,64,64] int64 labels:
// only_one_hourgless.py:348,370; try_different_stack.py:360-361,388-389; and on the channel slices [:, :18] /
// [:, 18:] of every stack's output, try_skeleton_and_keypoints.py:390-397,423-435).  The reference runs one
// log-softmax + NLL + mean (and their backward kernels, plus a slice-gradient zero-fill) per term; here ALL terms
// of a step go through one launch: blockIdx.y = term, one thread per pixel, channels strided by H*W (coalesced
// across the warp), max / sum-exp / gradient in three sweeps of which only the first misses L1.
// mean reduction over the labels != ignore_index, like the module's defaults.
struct CeArgs {
  HgCeTerm term[HG_CE_MAX_TERMS];
  float* loss;         // [T]  += sum nll / count
  int* count;          // [T]  valid labels (written by ce_count_kernel)
  int* bad;            // != NULL: set to 1 when a label is outside [0, C) and not ignore_index
  int B, HW, ignore_index;
  float gscale;
};

__global__ void __launch_bounds__(256) ce_count_kernel(const CeArgs a) {
  pdl_wait();
  pdl_trigger();
  const HgCeTerm& t = a.term[blockIdx.y];
  const long long n = (long long)a.B * a.HW;
  int c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long y = t.target[i];
    if (y >= 0 && y < t.channels) ++c;
    else if (y != a.ignore_index && a.bad) *a.bad = 1;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(a.count + blockIdx.y, c);
}

__global__ void __launch_bounds__(256) ce_multi_kernel(const CeArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  const HgCeTerm& t = a.term[blockIdx.y];
  const long long n = (long long)a.B * a.HW;
  // normaliser: the valid-label count (nn.CrossEntropyLoss, mean) or the caller's (bootstrapped / masked variants)
  const float inv = 1.f / (t.norm > 0.f ? t.norm : (float)a.count[blockIdx.y]);  // inf when every label is ignored
  const float k = a.gscale * inv;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / a.HW), p = (int)(i % a.HW);
    const float* x = t.logits + (long long)b * t.logits_bstride + p;
    const long long y = t.target[i];
    const bool valid = y >= 0 && y < t.channels;
    float m = -INFINITY;
    for (int c = 0; c < t.channels; ++c) m = fmaxf(m, __ldg(x + (long long)c * a.HW));
    float s = 0.f;
    for (int c = 0; c < t.channels; ++c) s += expf(__ldg(x + (long long)c * a.HW) - m);
    const float lse = m + logf(s);
    const float nll = valid ? lse - __ldg(x + y * a.HW) : 0.f;
    const float w = t.pixel_weight ? t.pixel_weight[i] : 1.f;
    acc += w * nll;
    if (t.nll_out) t.nll_out[i] = nll;
    if (t.dlogits) {
      float* g = t.dlogits + (long long)b * t.dlogits_bstride + p;
      const float kw = k * w;
      for (int c = 0; c < t.channels; ++c) {
        const float sm = expf(__ldg(x + (long long)c * a.HW) - lse);
        g[(long long)c * a.HW] = valid ? kw * (sm - (c == y ? 1.f : 0.f)) : 0.f;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float v = warp_sum(acc);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    atomicAdd(a.loss + blockIdx.y, tot * inv);
  }
}

// ------------------------------------------------------------------------------------------------------
// Bootstrapping (train.py:343-362 Costomer_CrossEntropyLoss, :394-408 Costomer_MSELoss): the mean of the k largest
// per-pixel losses of every image, torch.topk(loss.view(B, -1), k).  One block per image: 4-pass radix select (8-bit
// digits, MSB first) of the k-th largest value, then mask[i] = 1 for the k selected elements (values above the
// threshold, and the lowest-index ties at the threshold) -- the mask is the per-pixel weight of the loss kernels.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int order_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // larger float <=> larger key
}

__global__ void __launch_bounds__(1024) topk_mask_kernel(const float* __restrict__ v, int n, int k,
                                                         float* __restrict__ mask, float* __restrict__ kth) {
  pdl_wait();
  pdl_trigger();
  __shared__ int hist[256];
  __shared__ unsigned int s_prefix;
  __shared__ int s_need;
  __shared__ int s_scan[1024];
  const float* row = v + (long long)blockIdx.x * n;
  float* mrow = mask + (long long)blockIdx.x * n;
  if (threadIdx.x == 0) {
    s_prefix = 0u;
    s_need = k;
  }
  for (int d = 3; d >= 0; --d) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int himask = d == 3 ? 0u : (0xffffffffu << (8 * (d + 1)));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int key = order_key(row[i]);
      if ((key & himask) == prefix) atomicAdd(&hist[(key >> (8 * d)) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int need = s_need, b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= need) break;
        need -= hist[b];
      }
      s_need = need;
      s_prefix = prefix | ((unsigned int)b << (8 * d));
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix;   // key of the k-th largest value
  const int need = s_need;           // how many elements equal to it belong to the top k
  // ties at the threshold: the lowest indices win (contiguous chunk per thread + block scan of the tie counts)
  const int chunk = (n + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * chunk, hi = min(n, lo + chunk);
  int eq = 0;
  for (int i = lo; i < hi; ++i) eq += order_key(row[i]) == T;
  s_scan[threadIdx.x] = eq;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) {
      const int c = s_scan[t];
      s_scan[t] = run;
      run += c;
    }
  }
  __syncthreads();
  int before = s_scan[threadIdx.x];
  for (int i = lo; i < hi; ++i) {
    const unsigned int key = order_key(row[i]);
    float m = key > T ? 1.f : 0.f;
    if (key == T) {
      m = before < need ? 1.f : 0.f;
      ++before;
    }
    mrow[i] = m;
  }
  if (kth && threadIdx.x == 0) {
    const unsigned int u = (T & 0x80000000u) ? (T & 0x7fffffffu) : ~T;
    kth[blockIdx.x] = __uint_as_float(u);
  }
}

// ------------------------------------------------------------------------------------------------------
// Weighted / bootstrapped MSE (train.py:379-408 Costomer_MSELoss_with_mask, Costomer_MSELoss): one element-wise pass
//   sq_out[i] = (p - t)^2                      (feeds hg_topk_mask)            and / or
//   loss += sum w * (p - t)^2 / norm,  dpred[i] = 2 * gscale * w * (p - t) / norm
// with w per element (top-k mask over C*H*W) or per pixel, broadcast over the channels (mask input [B,H,W]).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_weighted_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                           const float* __restrict__ w, int w_per_pixel, int C, int HW,
                                                           long long numel, float norm, float gscale,
                                                           float* __restrict__ sq_out, float* __restrict__ dpred,
                                                           float* __restrict__ loss) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  float acc = 0.f;
  const float k = 2.f * gscale / norm;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    const float d = pred[i] - tgt[i];
    const float sq = d * d;
    if (sq_out) sq_out[i] = sq;
    float wi = 1.f;
    if (w) wi = w_per_pixel ? w[(i / ((long long)C * HW)) * HW + i % HW] : w[i];
    acc += wi * sq;
    if (dpred) dpred[i] = k * wi * d;
  }
  const float v = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(loss, tot / norm);
  }
}

}  // namespace hg

using namespace hg;

struct ScaleArgs {
  float* t[HG_MSE_MAX_STACKS];
  const float* scales;
  long long n4;
};

__global__ void __launch_bounds__(256) scale_multi_kernel(const ScaleArgs a) {
  pdl_wait();
  pdl_trigger();
  float* t = a.t[blockIdx.y];
  if (t == nullptr) return;
  const float sc = a.scales[blockIdx.y];
  float4* p = reinterpret_cast<float4*>(t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = p[i];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    p[i] = v;
  }
}

extern "C" {

int hg_zero_async(void* p, int64_t bytes, void* stream) {
  HG_REQUIRE(p != nullptr && bytes >= 0, "hg_zero_async: bad arguments");
  if (bytes > 0) HG_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream));
  return HG_OK;
}

int hg_scale_multi(int64_t numel, int32_t num_tensors, float* const* tensors_host, const float* scales, void* stream) {
  HG_REQUIRE(tensors_host && scales, "hg_scale_multi: NULL pointer");
  HG_REQUIRE(num_tensors > 0 && num_tensors <= HG_MSE_MAX_STACKS, "hg_scale_multi: 1..%d tensors supported", HG_MSE_MAX_STACKS);
  HG_REQUIRE(numel > 0 && numel % 4 == 0, "hg_scale_multi: numel must be a positive multiple of 4");
  ScaleArgs a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < num_tensors; ++s) {
    HG_REQUIRE((reinterpret_cast<uintptr_t>(tensors_host[s]) & 15) == 0, "hg_scale_multi: tensors must be 16-byte aligned");
    a.t[s] = tensors_host[s];
  }
  a.scales = scales;
  a.n4 = numel / 4;
  long long blocks = (a.n4 + 255) / 256;
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  launch_k(scale_multi_kernel, dim3((unsigned)blocks, (unsigned)num_tensors), dim3(256), 0, (cudaStream_t)stream, a);
  HG_LAUNCH_OK("scale_multi_kernel");
  count_launch();
  return HG_OK;
}

int hg_mse_multi(const HgMseDesc* d, const float* const* preds_host, const float* target, float* const* dpreds_host,
                 float* loss, void* stream) {
  HG_REQUIRE(d && preds_host && target && loss, "hg_mse_multi: NULL pointer");
  HG_REQUIRE(d->num_stacks > 0 && d->num_stacks <= HG_MSE_MAX_STACKS, "hg_mse_multi: 1..%d stacks supported",
             HG_MSE_MAX_STACKS);
  HG_REQUIRE(d->numel > 0, "hg_mse_multi: empty tensors");
  MseArgs a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < d->num_stacks; ++s) {
    HG_REQUIRE(preds_host[s] != nullptr, "hg_mse_multi: prediction %d is NULL", s);
    HG_REQUIRE((reinterpret_cast<uintptr_t>(preds_host[s]) & 15) == 0, "hg_mse_multi: tensors must be 16-byte aligned");
    a.pred[s] = preds_host[s];
    a.dpred[s] = dpreds_host ? dpreds_host[s] : nullptr;
  }
  HG_REQUIRE((reinterpret_cast<uintptr_t>(target) & 15) == 0, "hg_mse_multi: tensors must be 16-byte aligned");
  a.target = target;
  a.loss = loss;
  a.numel = d->numel;
  a.S = d->num_stacks;
  a.gscale = d->grad_scale;
  long long blocks = (d->numel / 4 + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  launch_k(mse_multi_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
  HG_LAUNCH_OK("mse_multi_kernel");
  count_launch();
  return HG_OK;
}

int hg_ce_multi(const HgCeDesc* d, const HgCeTerm* terms_host, float* loss, int32_t* count, int32_t* bad_label,
                void* stream) {
  HG_REQUIRE(d && terms_host && loss && count, "hg_ce_multi: NULL pointer");
  HG_REQUIRE(d->num_terms > 0 && d->num_terms <= HG_CE_MAX_TERMS, "hg_ce_multi: 1..%d terms supported",
             HG_CE_MAX_TERMS);
  HG_REQUIRE(d->B > 0 && d->HW > 0, "hg_ce_multi: empty tensors");
  CeArgs a;
  memset(&a, 0, sizeof(a));
  for (int t = 0; t < d->num_terms; ++t) {
    const HgCeTerm& h = terms_host[t];
    HG_REQUIRE(h.logits && h.target, "hg_ce_multi: term %d has a NULL tensor", t);
    HG_REQUIRE(h.channels > 0, "hg_ce_multi: term %d has no channels", t);
    HG_REQUIRE(h.logits_bstride >= (long long)h.channels * d->HW &&
                   (!h.dlogits || h.dlogits_bstride >= (long long)h.channels * d->HW),
               "hg_ce_multi: term %d: batch stride smaller than channels * H * W", t);
    a.term[t] = h;
  }
  a.loss = loss;
  a.count = count;
  a.bad = bad_label;
  a.B = d->B;
  a.HW = d->HW;
  a.ignore_index = d->ignore_index;
  a.gscale = d->grad_scale;
  const long long n = (long long)d->B * d->HW;
  long long blocks = (n + 255) / 256;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  dim3 grid((unsigned)blocks, (unsigned)d->num_terms);
  launch_k(ce_count_kernel, grid, dim3(256), 0, (cudaStream_t)stream, a);
  HG_LAUNCH_OK("ce_count_kernel");
  count_launch();
  launch_k(ce_multi_kernel, grid, dim3(256), 0, (cudaStream_t)stream, a);
  HG_LAUNCH_OK("ce_multi_kernel");
  count_launch();
  return HG_OK;
}

int hg_topk_mask(const float* values, int rows, int n, int k, float* mask, float* kth_value, void* stream) {
  HG_REQUIRE(values && mask, "hg_topk_mask: NULL pointer");
  HG_REQUIRE(rows > 0 && n > 0 && k > 0 && k <= n, "hg_topk_mask: need 0 < k <= n");
  launch_k(topk_mask_kernel, dim3((unsigned)rows), dim3(1024), 0, (cudaStream_t)stream, values, n, k, mask, kth_value);
  HG_LAUNCH_OK("topk_mask_kernel");
  count_launch();
  return HG_OK;
}

int hg_mse_weighted(const float* pred, const float* target, const float* weight, int weight_per_pixel, int B, int C,
                    int HW, float norm, float grad_scale, float* sq_out, float* dpred, float* loss, void* stream) {
  HG_REQUIRE(pred && target, "hg_mse_weighted: NULL pointer");
  HG_REQUIRE(B > 0 && C > 0 && HW > 0 && norm > 0.f, "hg_mse_weighted: bad sizes / normaliser");
  const long long numel = (long long)B * C * HW;
  long long blocks = (numel + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  launch_k(mse_weighted_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, pred, target, weight,
           weight_per_pixel, C, HW, numel, norm, grad_scale, sq_out, dpred, loss);
  HG_LAUNCH_OK("mse_weighted_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
