// BatchNorm (training-mode batch statistics and eval-mode running statistics) + ReLU, forward and backward,
// on NHWC activations.  HBM-bound streaming kernels: 128-bit accesses, one thread = 8 consecutive channels,
// per-channel reductions finish with a handful of fp32 atomics per block.
//
// The saved state of a BatchNorm call site is just its raw statistics (sum, sum of squares over N*H*W); mean,
// 1/std, scale and shift are re-derived in registers wherever they are needed, so no "finalize" pass exists.
//
// Replaces nn.BatchNorm2d + nn.ReLU(True) in ResidualBlock / lin
// (reference try_with_torch.py:184-192,196-204,249-250,254-255).
#include "hg_common.cuh"

#ifndef HG_BN_BWD_MINB
#define HG_BN_BWD_MINB 2   // resident blocks per SM the register allocation of bn_bwd_apply is held to
#endif

namespace hg {

// Batch statistics are SHIFTED sums: stats = {S1[Cp], S2[Cp], pivot[Cp]} with S1 = sum(x - pivot), S2 = sum((x - pivot)^2);
// mean = pivot + S1/n, var = S2/n - (S1/n)^2.  The cancellation in the variance is governed by (mean - pivot)^2 / var
// instead of mean^2 / var; the plan sets pivot = running_mean of the consuming BatchNorm (0 for a fresh module = plain
// sums), so a channel whose mean is large against its spread keeps its variance once the running mean has found it.
struct BnArgs {
  const float* stats;   // [3*Cp]: shifted sum, shifted sum of squares, pivot  (training mode)
  const float* gamma;   // [C]
  const float* beta;    // [C]
  const float* rmean;   // [C] running mean (eval mode)
  const float* rvar;    // [C] running var  (eval mode)
  float count;          // N*H*W
  float eps;
  int use_running;
  int relu;
  int C, Cp;
};

// scale/shift/mean/invstd of 8 consecutive channels starting at c0
__device__ __forceinline__ void bn_coeffs(const BnArgs& a, int c0, float (&mean)[8], float (&invstd)[8],
                                          float (&scale)[8], float (&shift)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    if (c < a.C) {
      float mu, var;
      if (a.use_running) {
        mu = a.rmean[c];
        var = a.rvar[c];
      } else {
        const float m1 = a.stats[c] / a.count;
        mu = a.stats[2 * a.Cp + c] + m1;
        var = fmaxf(a.stats[a.Cp + c] / a.count - m1 * m1, 0.f);
      }
      const float is = rsqrtf(var + a.eps);
      mean[e] = mu;
      invstd[e] = is;
      scale[e] = a.gamma[c] * is;
      shift[e] = a.beta[c] - mu * scale[e];
    } else {
      mean[e] = 0.f;
      invstd[e] = 0.f;
      scale[e] = 0.f;
      shift[e] = 0.f;
    }
  }
}

// Block-wide version: thread c (< Cp <= 256) derives channel c once, the block shares the result through shared
// memory (the per-thread version above costs 32 scalar global loads per thread, which dominates small tensors).
// coef layout: [4][256] = mean, invstd, scale, shift.  Contains a __syncthreads().
__device__ __forceinline__ void bn_coeffs_block(const BnArgs& a, float (*coef)[256], int vc, float (&mean)[8],
                                                float (&invstd)[8], float (&scale)[8], float (&shift)[8]) {
  for (int c = threadIdx.x; c < a.Cp; c += blockDim.x) {
    float mu = 0.f, is = 0.f, sc = 0.f, sh = 0.f;
    if (c < a.C) {
      float var;
      if (a.use_running) {
        mu = a.rmean[c];
        var = a.rvar[c];
      } else {
        const float m1 = a.stats[c] / a.count;
        mu = a.stats[2 * a.Cp + c] + m1;
        var = fmaxf(a.stats[a.Cp + c] / a.count - m1 * m1, 0.f);
      }
      is = rsqrtf(var + a.eps);
      sc = a.gamma[c] * is;
      sh = a.beta[c] - mu * sc;
    }
    coef[0][c] = mu;
    coef[1][c] = is;
    coef[2][c] = sc;
    coef[3][c] = sh;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mean[e] = coef[0][vc * 8 + e];
    invstd[e] = coef[1][vc * 8 + e];
    scale[e] = coef[2][vc * 8 + e];
    shift[e] = coef[3][vc * 8 + e];
  }
}

// ------------------------------------------------------------------------------------------------------
// statistics: stats[c] += sum_m (x[m,c] - pivot[c]); stats[Cp+c] += sum_m (x[m,c] - pivot[c])^2; pivot = stats[2Cp+c]
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long long M, int Cp,
                                                       float* __restrict__ stats, int rows_per_block) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  __shared__ float red[2][256][9];
  const int vecs = Cp >> 3;            // 8-channel vectors per row (8, 16 or 32)
  const int vc = threadIdx.x % vecs;   // vector column
  const int rl = threadIdx.x / vecs;   // row lane
  const int rlanes = 256 / vecs;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float s[8] = {}, ss[8] = {}, pv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) pv[e] = stats[2 * Cp + vc * 8 + e];
#pragma unroll 4
  for (long long m = m0 + rl; m < m1; m += rlanes) {
    float v[8];
    load8(x + m * Cp + vc * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[e] - pv[e];
      s[e] += d;
      ss[e] = fmaf(d, d, ss[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][threadIdx.x][e] = s[e];
    red[1][threadIdx.x][e] = ss[e];
  }
  __syncthreads();
  // thread t < Cp reduces channel t over the row lanes
  for (int c = threadIdx.x; c < Cp; c += 256) {
    const int v = c >> 3, e = c & 7;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rlanes; ++r) {
      a += red[0][r * vecs + v][e];
      b += red[1][r * vecs + v][e];
    }
    atomicAdd(stats + c, a);
    atomicAdd(stats + Cp + c, b);
  }
}

// rows per block: enough blocks to fill the machine, few enough that the per-channel atomics stay cheap
int g_bn_blocks_per_sm = 6;       // grid cap of bn_apply (no per-block atomics) (hg_set_option "bn_blocks_per_sm")
// Kernels whose blocks end with per-channel atomics into the SAME 2*C floats (bn_bwd_apply's bias-gradient column
// sums, bn_stats, colsum, bn_bwd_reduce): the L2 serialises those per cache line, so their grid is ONE wave.  Measured
// (B200, batch 32, bn_bwd_apply C128): 6 blocks/SM 24.4 us @64x64, 12.4 us @32x32;  2 blocks/SM 18.5 / 7.0 us;
// C256+addend @16x16: 9.7 us (6/SM) -> 6.0 us (1/SM).  0 = policy below, > 0 = fixed blocks per SM.
int g_bn_bwd_blocks_per_sm = 0;
int g_bn_apply_u4 = 1;            // bn_apply on large bf16 tensors: 4 rows in flight per thread (hg_set_option "bn_apply_u4")

static inline int atomic_grid_per_sm(long long M) {
  if (g_bn_bwd_blocks_per_sm > 0) return g_bn_bwd_blocks_per_sm;
  return M >= 32768 ? 2 : 1;
}

static inline int rows_per_block_for(long long M, int rlanes, int per_sm = 0) {
  long long blocks = M / (rlanes * 2);  // small tensors are latency-bound: at most two dependent loads per thread
  if (per_sm <= 0) per_sm = g_bn_blocks_per_sm;
  if (blocks > (long long)per_sm * kNumSMs) blocks = (long long)per_sm * kNumSMs;
  if (blocks < 1) blocks = 1;
  return (int)((M + blocks - 1) / blocks);
}

int bn_stats_launch(int dtype, const void* x, long long M, int Cp, float* stats, cudaStream_t st) {
  if (Cp % 64 != 0 || Cp > 2048) {
    set_error("bn_stats: padded channel count %d unsupported", Cp);
    return HG_ERR_UNSUPPORTED;
  }
  if (Cp > 256) {  // wide tensors (virtual cat inputs): one launch per 256-channel slab is not needed today
    set_error("bn_stats: Cp > 256 unsupported");
    return HG_ERR_UNSUPPORTED;
  }
  const int rpb = rows_per_block_for(M, 256 / (Cp >> 3), atomic_grid_per_sm(M));
  const int blocks = ceil_div(M, rpb);
  if (dtype == HG_BF16)
    launch_k(bn_stats_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, M, Cp, stats, rpb);
  else
    launch_k(bn_stats_kernel<float>, dim3(blocks), dim3(256), 0, st, (const float*)x, M, Cp, stats, rpb);
  HG_LAUNCH_OK("bn_stats_kernel");
  count_launch();
  return HG_OK;
}

// column sums only (bias gradient of a convolution: dbias[c] += sum_m dy[m, c])
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long M, int Cp, int C,
                                                     float* __restrict__ out, int rows_per_block) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  __shared__ float red[256][9];
  const int vecs = Cp >> 3;
  const int vc = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const int rlanes = 256 / vecs;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float s[8] = {};
#pragma unroll 4
  for (long long m = m0 + rl; m < m1; m += rlanes) {
    float v[8];
    load8(x + m * Cp + vc * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = s[e];
  __syncthreads();
  __shared__ float tot[256];
  for (int c = threadIdx.x; c < Cp; c += 256) {
    const int v = c >> 3, e = c & 7;
    float a = 0.f;
    for (int r = 0; r < rlanes; ++r) a += red[r * vecs + v][e];
    tot[c] = a;
  }
  __syncthreads();
  red_add_channels(out, tot, C);
}

int colsum_launch(int dtype, const void* dy, long long M, int Cp, int C, float* out, cudaStream_t st) {
  if (Cp % 64 != 0 || Cp > 256) {
    set_error("colsum: padded channel count %d unsupported", Cp);
    return HG_ERR_UNSUPPORTED;
  }
  const int rpb = rows_per_block_for(M, 256 / (Cp >> 3), atomic_grid_per_sm(M));
  const int blocks = ceil_div(M, rpb);
  if (dtype == HG_BF16)
    launch_k(colsum_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)dy, M, Cp, C, out, rpb);
  else
    launch_k(colsum_kernel<float>, dim3(blocks), dim3(256), 0, st, (const float*)dy, M, Cp, C, out, rpb);
  HG_LAUNCH_OK("colsum_kernel");
  count_launch();
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------------
// apply: y = [relu](gamma * (x - mean) * invstd + beta)
// ------------------------------------------------------------------------------------------------------
template <typename T, int U>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long M,
                                                       BnArgs a) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  // U rows in flight per thread, kept unconverted (a bf16 row = 4 registers): at full occupancy (8 blocks/SM) U = 2 is
  // 64 KB of loads in flight per SM -- short of what HBM3e needs at ~1 us latency (4.7 TB/s measured); U = 4 doubles it
  const int vecs = a.Cp >> 3;
  const int vc = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const int rlanes = 256 / vecs;
  const long long stride = (long long)gridDim.x * rlanes;
  long long m = (long long)blockIdx.x * rlanes + rl;
  // the first batch of rows is requested BEFORE the per-channel coefficients: one memory round trip instead of two
  // (small tensors are pure latency)
  Raw8<T> r[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (m + u * stride < M) load_raw(x + (m + u * stride) * a.Cp + vc * 8, r[u]);
  __shared__ float coef[4][256];
  float mean[8], invstd[8], scale[8], shift[8];
  bn_coeffs_block(a, coef, vc, mean, invstd, scale, shift);
  while (m < M) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (m + u * stride < M) {
        float v[8];
        unpack(r[u], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float o = fmaf(v[e], scale[e], shift[e]);
          if (a.relu) o = fmaxf(o, 0.f);
          v[e] = o;
        }
        store8(y + (m + u * stride) * a.Cp + vc * 8, v);
      }
    }
    m += U * stride;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (m + u * stride < M) load_raw(x + (m + u * stride) * a.Cp + vc * 8, r[u]);
  }
}

// ------------------------------------------------------------------------------------------------------
// backward, pass 1: red[c] += sum_m g, red[Cp+c] += sum_m g*xhat, with g = da * [bn(x) > 0]
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ x,
                                                            long long M, BnArgs a, float* __restrict__ redout,
                                                            int rows_per_block) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  __shared__ float red[2][256][9];
  const int vecs = a.Cp >> 3;
  const int vc = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const int rlanes = 256 / vecs;
  float mean[8], invstd[8], scale[8], shift[8];
  bn_coeffs(a, vc * 8, mean, invstd, scale, shift);
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float s[8] = {}, sx[8] = {};
#pragma unroll 2
  for (long long m = m0 + rl; m < m1; m += rlanes) {
    float xv[8], gv[8];
    load8(x + m * a.Cp + vc * 8, xv);
    load8(da + m * a.Cp + vc * 8, gv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float g = gv[e];
      if (a.relu && !(fmaf(xv[e], scale[e], shift[e]) > 0.f)) g = 0.f;
      s[e] += g;
      sx[e] = fmaf(g, (xv[e] - mean[e]) * invstd[e], sx[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][threadIdx.x][e] = s[e];
    red[1][threadIdx.x][e] = sx[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.Cp; c += 256) {
    const int v = c >> 3, e = c & 7;
    float p = 0.f, q = 0.f;
    for (int r = 0; r < rlanes; ++r) {
      p += red[0][r * vecs + v][e];
      q += red[1][r * vecs + v][e];
    }
    atomicAdd(redout + c, p);
    atomicAdd(redout + a.Cp + c, q);
  }
}

// ------------------------------------------------------------------------------------------------------
// backward, pass 2: dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) [+ addend]
//   training mode; in eval mode (running statistics are constants) dx = gamma*invstd*g [+ addend].
//   Block 0 adds dgamma += sum g*xhat, dbeta += sum g.  colsum (optional) += sum_m dx-without-addend,
//   i.e. the bias gradient of the convolution that produced x.
// ------------------------------------------------------------------------------------------------------
template <typename T, bool ADD, bool CS>
__global__ void __launch_bounds__(256, ADD ? 1 : HG_BN_BWD_MINB) bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ x,
                                                           const T* __restrict__ addend, T* __restrict__ dx,
                                                           long long M, BnArgs a, const float* __restrict__ redin,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           float* __restrict__ colsum, int rows_per_block) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  __shared__ float red[256][9];
  const int vecs = a.Cp >> 3;
  const int vc = threadIdx.x % vecs;
  const int rl = threadIdx.x / vecs;
  const int rlanes = 256 / vecs;
  // rows in flight per thread (three / four tensors each, kept unconverted).  With the addend (four streams) ONE block
  // per SM with eight rows in flight per thread (196 registers) runs at 0.84 of the copy bandwidth where two blocks with
  // four rows each reach 0.67 (C256 @64x64, batch 32: 60.0 -> 49.2 us); the three-stream kernel is the other way round
  // (18.4 vs 20.6 us at C128).
  constexpr int U = sizeof(T) == 2 ? (ADD ? 8 : 4) : 2;
  const long long m0 = (long long)blockIdx.x * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  long long m = m0 + rl;
  // request the first rows before the coefficient / reduction loads (latency of small tensors)
  Raw8<T> xr[U], gr[U], ar[ADD ? U : 1];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long mm = m + (long long)u * rlanes;
    if (mm < m1) {
      load_raw(x + mm * a.Cp + vc * 8, xr[u]);
      load_raw(da + mm * a.Cp + vc * 8, gr[u]);
      if (ADD) load_raw(addend + mm * a.Cp + vc * 8, ar[u]);
    }
  }
  __shared__ float coef[4][256];
  __shared__ float mred[2][256];
  for (int c = threadIdx.x; c < a.Cp; c += blockDim.x) {
    float r0 = 0.f, r1 = 0.f;
    if (c < a.C) {
      r0 = redin[c];
      r1 = redin[a.Cp + c];
      if (blockIdx.x == 0) {
        if (dgamma) atomicAdd(dgamma + c, r1);
        if (dbeta) atomicAdd(dbeta + c, r0);
      }
    }
    const bool stat = !a.use_running && c < a.C;
    mred[0][c] = stat ? r0 / a.count : 0.f;
    mred[1][c] = stat ? r1 / a.count : 0.f;
  }
  float mean[8], invstd[8], scale[8], shift[8];
  bn_coeffs_block(a, coef, vc, mean, invstd, scale, shift);  // (its __syncthreads also publishes mred)
  // dx = scale*(g - mg - xhat*mgx) = scale*g + cB*x + cC   (xhat = (x - mean)*invstd)
  float cB[8], cC[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float mg = mred[0][vc * 8 + e], mgx = mred[1][vc * 8 + e];
    cB[e] = -scale[e] * mgx * invstd[e];
    cC[e] = -scale[e] * mg - cB[e] * mean[e];
  }
  float cs[8] = {};
  while (m < m1) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long mm = m + (long long)u * rlanes;
      if (mm < m1) {
        float xv[8], gv[8], o[8];
        unpack(xr[u], xv);
        unpack(gr[u], gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float g = gv[e];
          if (a.relu && !(fmaf(xv[e], scale[e], shift[e]) > 0.f)) g = 0.f;
          o[e] = fmaf(scale[e], g, fmaf(cB[e], xv[e], cC[e]));
          if (CS) cs[e] += o[e];
        }
        if (ADD) {
          float ad[8];
          unpack(ar[ADD ? u : 0], ad);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += ad[e];
        }
        store8(dx + mm * a.Cp + vc * 8, o);
      }
    }
    m += (long long)U * rlanes;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long mm = m + (long long)u * rlanes;
      if (mm < m1) {
        load_raw(x + mm * a.Cp + vc * 8, xr[u]);
        load_raw(da + mm * a.Cp + vc * 8, gr[u]);
        if (ADD) load_raw(addend + mm * a.Cp + vc * 8, ar[u]);
      }
    }
  }
  if (CS) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = cs[e];
    __syncthreads();
    float* tot = &coef[0][0];  // coefficients are in registers by now
    for (int c = threadIdx.x; c < a.Cp; c += 256) {
      const int v = c >> 3, e = c & 7;
      float p = 0.f;
      for (int r = 0; r < rlanes; ++r) p += red[r * vecs + v][e];
      tot[c] = p;
    }
    __syncthreads();
    red_add_channels(colsum, tot, a.C);
  }
}

// ------------------------------------------------------------------------------------------------------
// running statistics: one block per BatchNorm module applies the EMA updates of all its call sites in the
// order the reference's forward executes them (a shared module is called 6-8 x nStack times per forward,
// reference try_with_torch.py:224-237; num_batches_tracked counts every call).
// ------------------------------------------------------------------------------------------------------
struct BnRunningSite {
  const float* stats;
  float count;
  int pad;
};
struct BnRunningModule {
  float* rmean;
  float* rvar;
  long long* nbt;
  int C, Cp;
  int first_site, num_sites;
  float momentum;
  int pad;
};

__global__ void bn_running_kernel(const BnRunningModule* __restrict__ mods, const BnRunningSite* __restrict__ sites) {
  pdl_wait();     // PDL: nothing below may touch global memory before the previous kernel has drained
  pdl_trigger();  // elementwise / streaming kernel: let the next kernel's CTAs queue up behind ours

  const BnRunningModule md = mods[blockIdx.x];
  for (int c = threadIdx.x; c < md.C; c += blockDim.x) {
    float rm = md.rmean[c], rv = md.rvar[c];
    // The EMA is a sequential recurrence over the call sites, but its inputs are not: the statistics of eight sites are
    // requested together (a site at a time was two dependent L2 round trips per site: 80 us for the 64 calls of a shared
    // module at the end of every forward; 54 us now).
    for (int i0 = 0; i0 < md.num_sites; i0 += 8) {
      float mu[8], unb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = i0 + j < md.num_sites ? i0 + j : md.num_sites - 1;
        const BnRunningSite s = sites[md.first_site + i];
        const float s1 = s.stats[c], s2 = s.stats[md.Cp + c], pv = s.stats[2 * md.Cp + c];
        const float m1 = s1 / s.count;
        mu[j] = pv + m1;
        const float var = fmaxf(s2 / s.count - m1 * m1, 0.f);
        unb[j] = s.count > 1.f ? var * s.count / (s.count - 1.f) : var;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (i0 + j < md.num_sites) {
          rm = (1.f - md.momentum) * rm + md.momentum * mu[j];
          rv = (1.f - md.momentum) * rv + md.momentum * unb[j];
        }
      }
    }
    md.rmean[c] = rm;
    md.rvar[c] = rv;
  }
  if (threadIdx.x == 0 && md.nbt) *md.nbt += md.num_sites;
}

// One block per statistics slot: S1 = S2 = 0, pivot = pivot_src (running mean of the consuming BatchNorm) or 0.
__global__ void bn_prepare_stats_kernel(const HgBnStatsSlot* __restrict__ slots) {
  pdl_wait();
  pdl_trigger();
  const HgBnStatsSlot sl = slots[blockIdx.x];
  for (int c = threadIdx.x; c < sl.Cp; c += blockDim.x) {
    sl.stats[c] = 0.f;
    sl.stats[sl.Cp + c] = 0.f;
    sl.stats[2 * sl.Cp + c] = (sl.pivot_src != nullptr && c < sl.C) ? sl.pivot_src[c] : 0.f;
  }
}

static int check_bn(const HgBnDesc* d) {
  HG_REQUIRE(d != nullptr, "HgBnDesc is NULL");
  HG_REQUIRE(d->M > 0 && d->C > 0, "HgBnDesc: non-positive size");
  HG_REQUIRE(d->dtype == HG_BF16 || d->dtype == HG_F32, "HgBnDesc: bad dtype");
  const int Cp = (d->C + 63) & ~63;
  if (Cp > 256) {
    set_error("BatchNorm over %d channels unsupported (max 256)", d->C);
    return HG_ERR_UNSUPPORTED;
  }
  return HG_OK;
}

static BnArgs make_args(const HgBnDesc* d, const float* stats, const float* gamma, const float* beta,
                        const float* rmean, const float* rvar) {
  BnArgs a;
  a.stats = stats;
  a.gamma = gamma;
  a.beta = beta;
  a.rmean = rmean;
  a.rvar = rvar;
  a.count = (float)d->M;
  a.eps = d->eps;
  a.use_running = d->use_running;
  a.relu = d->relu;
  a.C = d->C;
  a.Cp = (d->C + 63) & ~63;
  return a;
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_bn_stats(const HgBnDesc* d, const void* x, float* stats, void* stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  HG_REQUIRE(x && stats, "hg_bn_stats: NULL pointer");
  return bn_stats_launch(d->dtype, x, d->M, (d->C + 63) & ~63, stats, (cudaStream_t)stream);
}

int hg_bn_apply(const HgBnDesc* d, const void* x, const float* stats, const float* gamma, const float* beta,
                const float* running_mean, const float* running_var, void* y, void* stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  HG_REQUIRE(x && y && gamma && beta, "hg_bn_apply: NULL pointer");
  HG_REQUIRE(d->use_running ? (running_mean && running_var) : (stats != nullptr),
             "hg_bn_apply: statistics missing for the selected mode");
  BnArgs a = make_args(d, stats, gamma, beta, running_mean, running_var);
  const int rlanes = 256 / (a.Cp >> 3);
  int blocks = ceil_div(d->M, rlanes * 2);  // two rows in flight per thread; large tensors: 8 blocks per SM
  const bool deep = g_bn_apply_u4 && blocks > kNumSMs * 8 * 2;   // >= 4 rows per thread at the grid cap
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == HG_BF16) {
    if (deep) launch_k(bn_apply_kernel<__nv_bfloat16, 4>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, d->M, a);
    else launch_k(bn_apply_kernel<__nv_bfloat16, 2>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, d->M, a);
  } else {
    launch_k(bn_apply_kernel<float, 2>, dim3(blocks), dim3(256), 0, st, (const float*)x, (float*)y, d->M, a);
  }
  HG_LAUNCH_OK("bn_apply_kernel");
  count_launch();
  return HG_OK;
}

int hg_bn_bwd_reduce(const HgBnDesc* d, const void* da, const void* x, const float* stats, const float* gamma,
                     const float* beta, const float* running_mean, const float* running_var, float* red,
                     void* stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  HG_REQUIRE(da && x && gamma && beta && red, "hg_bn_bwd_reduce: NULL pointer");
  HG_REQUIRE(d->use_running ? (running_mean && running_var) : (stats != nullptr),
             "hg_bn_bwd_reduce: statistics missing for the selected mode");
  BnArgs a = make_args(d, stats, gamma, beta, running_mean, running_var);
  const int rpb = rows_per_block_for(d->M, 256 / (a.Cp >> 3), atomic_grid_per_sm(d->M));
  const int blocks = ceil_div(d->M, rpb);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == HG_BF16)
    launch_k(bn_bwd_reduce_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)da, (const __nv_bfloat16*)x, d->M, a, red, rpb);
  else
    launch_k(bn_bwd_reduce_kernel<float>, dim3(blocks), dim3(256), 0, st, (const float*)da, (const float*)x, d->M, a, red, rpb);
  HG_LAUNCH_OK("bn_bwd_reduce_kernel");
  count_launch();
  return HG_OK;
}

int hg_bn_bwd_apply(const HgBnDesc* d, const void* da, const void* x, const float* stats, const float* gamma,
                    const float* beta, const float* running_mean, const float* running_var, const float* red,
                    const void* addend, void* dx, float* dgamma, float* dbeta, float* colsum, void* stream) {
  int rc = check_bn(d);
  if (rc) return rc;
  HG_REQUIRE(da && x && gamma && beta && dx, "hg_bn_bwd_apply: NULL pointer");
  HG_REQUIRE(d->use_running ? (running_mean && running_var) : (stats && red),
             "hg_bn_bwd_apply: statistics missing for the selected mode");
  BnArgs a = make_args(d, stats, gamma, beta, running_mean, running_var);
  const bool wide = addend != nullptr && d->dtype == HG_BF16;   // U = 8, one block per SM (see the kernel)
  const int rpb = rows_per_block_for(d->M, 256 / (a.Cp >> 3), wide && g_bn_bwd_blocks_per_sm <= 0 ? 1 : atomic_grid_per_sm(d->M));
  const int blocks = ceil_div(d->M, rpb);
  cudaStream_t st = (cudaStream_t)stream;
#define HG_BWD_APPLY(T, ADD, CS)                                                                                 \
  launch_k(bn_bwd_apply_kernel<T, ADD, CS>, dim3(blocks), dim3(256), 0, st, (const T*)da, (const T*)x,            \
           (const T*)addend, (T*)dx, d->M, a, red, dgamma, dbeta, colsum, rpb)
#define HG_BWD_APPLY_T(T)                                   \
  do {                                                      \
    if (addend && colsum) HG_BWD_APPLY(T, true, true);      \
    else if (addend) HG_BWD_APPLY(T, true, false);          \
    else if (colsum) HG_BWD_APPLY(T, false, true);          \
    else HG_BWD_APPLY(T, false, false);                     \
  } while (0)
  if (d->dtype == HG_BF16) HG_BWD_APPLY_T(__nv_bfloat16);
  else HG_BWD_APPLY_T(float);
#undef HG_BWD_APPLY_T
#undef HG_BWD_APPLY
  HG_LAUNCH_OK("bn_bwd_apply_kernel");
  count_launch();
  return HG_OK;
}

int hg_bn_prepare_stats(const HgBnStatsSlot* slots_dev, int num_slots, void* stream) {
  HG_REQUIRE(slots_dev && num_slots > 0, "hg_bn_prepare_stats: bad arguments");
  launch_k(bn_prepare_stats_kernel, dim3(num_slots), dim3(64), 0, (cudaStream_t)stream, slots_dev);
  HG_LAUNCH_OK("bn_prepare_stats_kernel");
  count_launch();
  return HG_OK;
}

int hg_bn_update_running(const void* modules_dev, const void* sites_dev, int num_modules, void* stream) {
  HG_REQUIRE(modules_dev && sites_dev && num_modules > 0, "hg_bn_update_running: bad arguments");
  launch_k(bn_running_kernel, dim3(num_modules), dim3(256), 0, (cudaStream_t)stream, (const BnRunningModule*)modules_dev,
                                                                   (const BnRunningSite*)sites_dev);
  HG_LAUNCH_OK("bn_running_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
