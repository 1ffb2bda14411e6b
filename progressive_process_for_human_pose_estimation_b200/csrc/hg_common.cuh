// Shared device/host helpers for the hourglass sm_100a kernels.
// Everything here is written for sm_100a only (tcgen05 / TMEM / TMA / mbarrier).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hg_sm100a.h"

namespace hg {

// ------------------------------------------------------------------------------------------
// Host-side error plumbing: the C ABI never throws; it returns a negative HgStatus and keeps
// the message for hg_last_error_string().
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define HG_CUDA_OK(expr)                                                  \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) return ::hg::cuda_fail(_e, #expr);             \
  } while (0)

#define HG_REQUIRE(cond, ...)                                             \
  do {                                                                    \
    if (!(cond)) {                                                        \
      ::hg::set_error(__VA_ARGS__);                                       \
      return HG_ERR_BAD_ARG;                                              \
    }                                                                     \
  } while (0)

#define HG_LAUNCH_OK(name)                                                \
  do {                                                                    \
    cudaError_t _e = cudaGetLastError();                                  \
    if (_e != cudaSuccess) return ::hg::cuda_fail(_e, name);              \
  } while (0)

// Number of launches issued through the library since load (read by hg_launch_count()).
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

constexpr int kNumSMs = 148;

// Programmatic dependent launch (PDL): consecutive kernels of a stream are chained so that the next kernel's CTAs are
// resident and through their prologue when the previous kernel drains.  Every kernel launched through launch_k()
// executes pdl_wait() before its first access to global memory that an earlier kernel may produce (or still read),
// and pdl_trigger() once its main work is issued.  hg_set_option("pdl", 0) falls back to plain stream order.
extern int g_use_pdl;

// ------------------------------------------------------------------------------------------
// TMA descriptor encoding (driver entry point fetched at run time: no libcuda link needed).
// ------------------------------------------------------------------------------------------
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */,
                const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swz);


// Epilogue / prologue fusion modes of the tensor-core convolution kernels (conv_tc.cu, conv_p1.cu, conv_p3.cu).
//   kPlain : y = conv(x) [+bias][+residual][+stats]                                   (fprop, dgrad)
//   kFold  : the A operand is a RAW tensor x that a BatchNorm(+ReLU) normalises: every A tile is rewritten in
//            shared memory as a = [relu](scale_c * x + shift_c) before the tensor core reads it, so the normalised
//            activation never exists in HBM (fprop of BN -> ReLU -> conv, reference try_with_torch.py:196-205)
//   kMask  : dgrad whose result is the gradient of a BatchNorm(+ReLU) output: the epilogue applies the ReLU mask
//            g = da * [bn(x) > 0], stores g and accumulates the two BatchNorm-backward sums (sum g, sum g*xhat)
//   kPlainBnOut : kPlain whose epilogue also applies an inference-mode BatchNorm(+ReLU) to the OUTPUT channels,
//            y = [relu](scale_c * (conv + bias_c) + shift_c) with running statistics (no residual, no statistics)
enum { kPlain = 0, kFold = 1, kMask = 2, kPlainBnOut = 3 };

// BatchNorm folded into a convolution (device view of HgBnFold)
struct BnFoldDev {
  const float* stats;   // shifted sums {S1, S2, pivot}[3*Cp] of the raw tensor (training mode, bn.cu)
  const float* gamma;
  const float* beta;
  const float* rmean;
  const float* rvar;
  float count;
  float eps;
  int relu;
  int use_running;
  int C, Cp;
};

#ifdef __CUDACC__
__device__ __forceinline__ void bn_fold_coeffs(const BnFoldDev& f, int c, float& mean, float& invstd, float& scale,
                                               float& shift) {
  if (c < f.C) {
    float mu, var;
    if (f.use_running) {
      mu = f.rmean[c];
      var = f.rvar[c];
    } else {
      const float m1 = f.stats[c] / f.count;     // shifted sums: {S1, S2, pivot} (bn.cu)
      mu = f.stats[2 * f.Cp + c] + m1;
      var = fmaxf(f.stats[f.Cp + c] / f.count - m1 * m1, 0.f);
    }
    invstd = rsqrtf(var + f.eps);
    mean = mu;
    scale = f.gamma[c] * invstd;
    shift = f.beta[c] - mu * scale;
  } else {
    mean = invstd = scale = shift = 0.f;
  }
}

__device__ __forceinline__ void load_coef8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch_k with a thread-block cluster of `cluster_x` CTAs along x (the grid must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------
// Storage-type helpers: activations are bf16 (tensor-core path) or fp32 (CUDA-core path).
// ------------------------------------------------------------------------------------------
template <typename T> struct Vec8;  // 8 consecutive channels

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 8-channel vector load/store (16 B for bf16, 32 B for fp32). Pointers must be aligned.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = reinterpret_cast<const float4*>(p)[0];
  float4 b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// Raw (unconverted) 8-channel vectors: streaming kernels keep several rows in flight per thread, and a bf16 row
// costs 4 registers raw but 8 once converted to float.
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 v; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void load_raw(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) {
  r.v = *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void load_raw(const float* p, Raw8<float>& r) {
  r.a = reinterpret_cast<const float4*>(p)[0];
  r.b = reinterpret_cast<const float4*>(p)[1];
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void unpack(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
  v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

// Every block of a reduction kernel adds C per-channel partial sums into the same few cache lines, and the L2
// serialises atomics per line: 4 channels per red.global.add.v4.f32 (dst 16-byte aligned), scalar tail.
__device__ __forceinline__ void red_add_channels(float* dst, const float* src_smem, int C) {
  const int c4n = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (C >> 2) : 0;
  for (int q = threadIdx.x; q < c4n; q += blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(src_smem + q * 4);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q * 4), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
  }
  for (int c = c4n * 4 + threadIdx.x; c < C; c += blockDim.x) atomicAdd(dst + c, src_smem[c]);
}

// BatchNorm statistics fused into a streaming producer (256 threads, thread = 8 channels `vc*8..` of some rows):
// stats[c] += sum, stats[Cp + c] += sum of squares over everything this block wrote.  Contains __syncthreads().
__device__ __forceinline__ void block_channel_stats(const float (&s)[8], const float (&ss)[8], int Cp,
                                                    float* __restrict__ stats) {
  __shared__ float red[2][256][9];
  __shared__ __align__(16) float tot[2][256];
  const int vecs = Cp >> 3;
  const int rlanes = 256 / vecs;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][threadIdx.x][e] = s[e];
    red[1][threadIdx.x][e] = ss[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cp; c += 256) {
    const int v = c >> 3, e = c & 7;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rlanes; ++r) {
      a += red[0][r * vecs + v][e];
      b += red[1][r * vecs + v][e];
    }
    tot[0][c] = a;
    tot[1][c] = b;
  }
  __syncthreads();
  red_add_channels(stats, tot[0], Cp);
  red_add_channels(stats + Cp, tot[1], Cp);
}

// Packed fp32 pairs (sm_100: FADD2 / FFMA2 retire two IEEE round-to-nearest operations per issue slot; results are
// bit-identical to the scalar instructions).  The convolution epilogues are instruction-issue bound.
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums of a 32 x 32 register tile (lane = row, v[i] = column i): afterwards lane l holds the sum of column l.
// Each step sends one half of the remaining columns to the partner lane and keeps the other: 31 shuffles in all
// (a butterfly per column would be 160).  v[] is destroyed.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int off = 16 >> step;          // partner distance == number of columns kept
    const bool up = (lane & off) != 0;   // lanes with this bit set keep the upper half of the columns
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("hg: mbarrier wait timed out (block %d,%d thread %d)\n", (int)blockIdx.x,
             (int)blockIdx.y, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// dst[0..bytes) += src[0..bytes) as fp32, reduced in L2 by the bulk-copy engine (both 16-byte aligned, bytes % 16 == 0)
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst_global, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                   reinterpret_cast<uint64_t>(dst_global)),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// --- tcgen05 ---
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- TMEM lane i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// One lane of a converged warp (the MMA issuer runs its loop warp-uniformly and issues under this predicate: inside a
// divergent `if (lane == 0)` the compiler cannot keep the tcgen05 operands in uniform registers and wraps every
// instruction in a per-lane election loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   K-major : rows of 64 bf16 (128 B); 8-row groups 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major: 64 MN-elements (128 B) per K-row; 8-K-row groups SBO apart; 64-element MN panels
//             LBO apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                      // D format: F32
         | (1u << 7)                    // A format: BF16
         | (1u << 10)                   // B format: BF16
         | ((uint32_t)a_mn_major << 15) // A major
         | ((uint32_t)b_mn_major << 16) // B major
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}
#endif  // __CUDACC__

}  // namespace hg
