// Adam update of every parameter tensor of the model in ONE launch (torch.optim.Adam(model.parameters(), lr=1e-4):
// try_with_torch.py:317,342-344; eps=1e-4 under amp, hourglass_compare.py:885).  The 1.9 M parameters of the
// weight-shared hourglass live in 199 tensors: the stock optimizer is launch-bound (one or several kernels per
// tensor or per foreach-group); here a device-resident chunk table maps every thread block to a slice of one tensor.
//
// Arithmetic follows torch.optim.Adam's single-tensor path (amsgrad=False, maximize=False), all in fp32:
//   g   = grad [+ weight_decay * p]
//   m   = m + (g - m) * (1 - beta1)                      (Tensor.lerp_)
//   v   = v * beta2 + (1 - beta2) * g * g                (mul_ + addcmul_)
//   p   = p - step_size * m / (sqrt(v) / sqrt(bc2) + eps),   step_size = lr / bc1,  bcK = 1 - betaK^step
#include "hg_common.cuh"

namespace hg {

struct AdamArgs {
  const HgAdamChunk* chunks;
  float one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, step_size, bc2_sqrt;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
  m = m + (g - m) * a.one_minus_beta1;
  v = v * a.beta2 + a.one_minus_beta2 * g * g;
  const float denom = __fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt) + a.eps;
  p = p - a.step_size * __fdiv_rn(m, denom);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamArgs a) {
  pdl_wait();
  pdl_trigger();
  const HgAdamChunk c = a.chunks[blockIdx.x];
  const bool vec = ((reinterpret_cast<uintptr_t>(c.param) | reinterpret_cast<uintptr_t>(c.grad) |
                     reinterpret_cast<uintptr_t>(c.exp_avg) | reinterpret_cast<uintptr_t>(c.exp_avg_sq)) & 15) == 0;
  long long i0 = 0;
  if (vec) {
    const long long nv = c.n >> 2;
    for (long long i = threadIdx.x; i < nv; i += blockDim.x) {
      float4 p = reinterpret_cast<float4*>(c.param)[i];
      const float4 g = reinterpret_cast<const float4*>(c.grad)[i];
      float4 m = reinterpret_cast<float4*>(c.exp_avg)[i];
      float4 v = reinterpret_cast<float4*>(c.exp_avg_sq)[i];
      adam_one(p.x, g.x, m.x, v.x, a);
      adam_one(p.y, g.y, m.y, v.y, a);
      adam_one(p.z, g.z, m.z, v.z, a);
      adam_one(p.w, g.w, m.w, v.w, a);
      reinterpret_cast<float4*>(c.param)[i] = p;
      reinterpret_cast<float4*>(c.exp_avg)[i] = m;
      reinterpret_cast<float4*>(c.exp_avg_sq)[i] = v;
    }
    i0 = nv << 2;
  }
  for (long long i = i0 + threadIdx.x; i < c.n; i += blockDim.x) {
    float p = c.param[i], m = c.exp_avg[i], v = c.exp_avg_sq[i];
    adam_one(p, c.grad[i], m, v, a);
    c.param[i] = p;
    c.exp_avg[i] = m;
    c.exp_avg_sq[i] = v;
  }
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_adam_multi(const HgAdamDesc* d, const HgAdamChunk* chunks_dev, void* stream) {
  HG_REQUIRE(d && chunks_dev, "hg_adam_multi: NULL pointer");
  HG_REQUIRE(d->num_chunks > 0, "hg_adam_multi: no chunks");
  HG_REQUIRE(d->step >= 1, "hg_adam_multi: step counts from 1");
  HG_REQUIRE(d->beta1 >= 0.f && d->beta1 < 1.f && d->beta2 >= 0.f && d->beta2 < 1.f && d->eps >= 0.f && d->lr >= 0.f,
             "hg_adam_multi: invalid hyper-parameters");
  AdamArgs a;
  a.chunks = chunks_dev;
  // python-side torch.optim.Adam evaluates the bias corrections and step_size in double precision
  const double bc1 = 1.0 - pow((double)d->beta1_d, (double)d->step);
  const double bc2 = 1.0 - pow((double)d->beta2_d, (double)d->step);
  a.one_minus_beta1 = (float)(1.0 - d->beta1_d);
  a.beta2 = (float)d->beta2_d;
  a.one_minus_beta2 = (float)(1.0 - d->beta2_d);
  a.eps = d->eps;
  a.weight_decay = d->weight_decay;
  a.step_size = (float)((double)d->lr_d / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  launch_k(adam_multi_kernel, dim3((unsigned)d->num_chunks), dim3(256), 0, (cudaStream_t)stream, a);
  HG_LAUNCH_OK("adam_multi_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
