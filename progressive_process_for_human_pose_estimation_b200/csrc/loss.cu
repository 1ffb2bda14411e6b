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

}  // namespace hg

using namespace hg;

extern "C" {

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

}  // extern "C"
