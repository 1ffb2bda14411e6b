// C ABI entry points shared by all kernels: error reporting, TMA descriptor encoding, dispatch between the
// tensor-core (bf16) and CUDA-core (fp32 / odd geometry) convolution kernels.
#include <stdarg.h>

#include "hg_common.cuh"

namespace hg {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
static int g_force_ref_conv = 0;
static int g_allow_ref_conv = 0;
int g_use_pdl = 1;
extern int g_stem_tc, g_stem_bwd_blocks_per_sm, g_persist_dynamic, g_persist_transposed, g_persist_3x3, g_persist_1x1, g_persist_min_units, g_ps_dbg, g_upsample_sep, g_upsample_fwd_cap, g_onewave_cluster, g_wgrad_halo;
extern int g_wgrad_fused_bias, g_mid_n_tiles, g_wgrad_kpx;
extern int g_single_wave_deep, g_wgrad_smem_kb, g_small_n_tiles, g_wgrad_dbg, g_wgrad_bulk_reduce, g_wgrad_t1_max_kb, g_wgrad_small_n_panels, g_wgrad_big_n_panels, g_short_alias, g_long_k_3cta, g_short_1stage;
extern int g_bn_blocks_per_sm, g_bn_bwd_blocks_per_sm, g_bn_apply_u4;
extern long long* g_dbg_ts;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return HG_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                CUtensorMapSwizzle swz) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || p == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return HG_ERR_CUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, base %p, dims %llu %llu %llu %llu)", (int)r,
              rank, base, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
    return HG_ERR_CUDA;
  }
  return HG_OK;
}

// implemented in conv_tc.cu / conv_ref.cu
BnFoldDev make_fold(const HgBnFold* f, int C, long long count);
int conv_tc_fprop(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias, const void* res, void* y,
                  float* stats, float* out_nchw, int mode, const BnFoldDev* fold, cudaStream_t st);
int conv_tc_dgrad(const HgConvDesc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx, float* red,
                  int mode, const BnFoldDev* fold, cudaStream_t st);
int conv_wgrad_bf16(const HgConvDesc* d, const void* x, const void* dy, float* dw, float* dbias,
                    const BnFoldDev* fold, cudaStream_t st);
template <typename T>
int conv_ref_fprop(const HgConvDesc*, const void*, const void*, const float*, const void*, void*, float*,
                   cudaStream_t);
template <typename T>
int conv_ref_dgrad(const HgConvDesc*, const void*, const void*, const void*, void*, cudaStream_t);
template <typename T>
int conv_ref_wgrad(const HgConvDesc*, const void*, const void*, float*, float*, cudaStream_t);
template <typename T>
int pack_weight(const HgConvDesc*, const float*, void*, void*, int, int, cudaStream_t);
int unpack_wgrad(const HgConvDesc* d, const float* g, float* dw, int accumulate, int cin_total, int cin_off,
                 cudaStream_t st);
int mix_rows(const float* T, const float* in, float* out, int Ro, int Ri, int cols, int transpose, int accumulate,
             cudaStream_t st);
int bn_stats_launch(int dtype, const void* x, long long M, int Cp, float* stats, cudaStream_t st);

static inline int pad64(int c) { return (c + 63) & ~63; }

static int check_desc(const HgConvDesc* d) {
  HG_REQUIRE(d != nullptr, "HgConvDesc is NULL");
  HG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "HgConvDesc: non-positive size");
  HG_REQUIRE(d->R > 0 && d->S > 0 && d->stride > 0 && d->dil > 0 && d->pad >= 0, "HgConvDesc: bad filter geometry");
  HG_REQUIRE(d->dtype == HG_BF16 || d->dtype == HG_F32, "HgConvDesc: dtype must be HG_BF16 or HG_F32");
  return HG_OK;
}

// The tensor-core kernels take stride-1 "same" convolutions and their stride-2 counterparts (output = input / 2:
// 3x3 pad 1, 1x1 pad 0) whose OUTPUT map is a power of two no wider than 128, with at most 256 padded channels.
static bool tc_eligible(const HgConvDesc* d) {
  if (g_force_ref_conv || d->dtype != HG_BF16) return false;
  if ((d->stride != 1 && d->stride != 2) || d->R != d->S) return false;
  int Ho = d->H, Wo = d->W;
  if (d->stride == 1) {
    if (2 * d->pad != d->dil * (d->R - 1)) return false;
  } else {
    if ((d->H & 1) || (d->W & 1) || d->dil != 1) return false;
    if (!((d->R == 3 && d->pad == 1) || (d->R == 1 && d->pad == 0))) return false;
    Ho = d->H / 2;
    Wo = d->W / 2;
  }
  if (!is_pow2(Ho) || !is_pow2(Wo) || Wo > 128) return false;
  if (d->R * d->S > 12 || d->dil * (d->R - 1) > 127) return false;
  const int ci = pad64(d->Cin), co = pad64(d->Cout);
  if (ci > 256 || co > 256 || co == 192 || ci == 192) return false;
  return true;
}

// bf16 convolutions the tensor-core kernels do not take: a hard error (north_star: no silent slow path) unless the
// caller opted into the CUDA-core kernels with hg_set_option("allow_ref_conv", 1).
static int ref_conv_allowed(const HgConvDesc* d, const char* who) {
  if (d->dtype != HG_BF16 || g_allow_ref_conv || g_force_ref_conv) return HG_OK;
  set_error("%s: bf16 convolution %dx%d k%d stride %d pad %d dil %d, %d -> %d channels is outside the tensor-core "
            "kernels' geometry (stride 1 'same' or stride 2 halving, power-of-two output map <= 128 wide, <= 256 "
            "padded channels); hg_set_option(\"allow_ref_conv\", 1) runs it on the 20-50x slower CUDA-core kernel",
            who, d->H, d->W, d->R, d->stride, d->pad, d->dil, d->Cin, d->Cout);
  return HG_ERR_UNSUPPORTED;
}

}  // namespace hg

using namespace hg;

extern "C" {

const char* hg_last_error_string(void) { return g_err; }
unsigned long long hg_launch_count(void) { return g_launches; }

int hg_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

int hg_set_option(const char* name, int value) {
  if (strcmp(name, "force_ref_conv") == 0) {
    g_force_ref_conv = value;
    return HG_OK;
  }
  if (strcmp(name, "allow_ref_conv") == 0) {
    g_allow_ref_conv = value;
    return HG_OK;
  }
  if (strcmp(name, "dbg_ts") == 0) {   // 1: start stamping conv_gemm CTA 0 phases; 2: print the last kernel's stamps
    if (value == 1 && !g_dbg_ts) {
      if (cudaMalloc(&g_dbg_ts, 32 * sizeof(long long)) != cudaSuccess) return HG_ERR_CUDA;
      cudaMemset(g_dbg_ts, 0, 32 * sizeof(long long));
    } else if (value == 2 && g_dbg_ts) {
      long long h[32];
      cudaDeviceSynchronize();
      cudaMemcpy(h, g_dbg_ts, sizeof(h), cudaMemcpyDeviceToHost);
      static const char* nm[11] = {"entry", "prologue done", "pdl_wait done", "first TMA issued", "first tile landed",
                                   "last MMA committed", "accumulator ready", "row pass done", "stats+store issued",
                                   "store read done", "exit"};
      for (int i = 1; i < 11; ++i)
        fprintf(stderr, "  %-20s +%6lld cycles (%.2f us)\n", nm[i], h[i] - h[0], (h[i] - h[0]) / 1965.0);
      fprintf(stderr, "  producer loop entry +%lld; TMA pair issued at:", h[15] - h[0]);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld", h[16 + i] - h[0]);
      fprintf(stderr, "\n  MMA warp saw tile k at:");
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %lld", h[24 + i] - h[0]);
      fprintf(stderr, "\n");
    } else if (value == 3 && g_dbg_ts) {   // conv_persist_kernel (DBG build): per-role wait / phase cycles of CTA 0
      long long h[32];
      cudaDeviceSynchronize();
      cudaMemcpy(h, g_dbg_ts, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "  producer: total %lld, a_empty %lld, b_empty %lld\n", h[0], h[1], h[2]);
      fprintf(stderr, "  mma     : total %lld, tmem_empty %lld, a_full %lld, b_full %lld\n", h[4], h[5], h[6], h[7]);
      fprintf(stderr, "  epilogue: total %lld, tmem_full %lld, y+ld %lld, row %lld, col %lld, store wait %lld\n", h[8], h[9],
              h[10], h[11], h[12], h[13]);
    } else if (value == 0) {
      g_dbg_ts = nullptr;
    }
    return HG_OK;
  }
  if (strcmp(name, "wgrad_bulk_reduce") == 0) {
    g_wgrad_bulk_reduce = value;
    return HG_OK;
  }
  if (strcmp(name, "bn_blocks_per_sm") == 0 && value > 0) {
    g_bn_blocks_per_sm = value;
    return HG_OK;
  }
  if (strcmp(name, "bn_bwd_blocks_per_sm") == 0 && value >= 0) {
    g_bn_bwd_blocks_per_sm = value;
    return HG_OK;
  }
  if (strcmp(name, "persist_1x1") == 0) {   // large-map 1x1 convolutions through conv_persist_kernel (default 1)
    g_persist_1x1 = value;
    return HG_OK;
  }
  if (strcmp(name, "persist_3x3") == 0) {   // large-map 3x3 convolutions through conv_persist_kernel (default 0)
    g_persist_3x3 = value;
    return HG_OK;
  }
  if (strcmp(name, "stem_tc") == 0) {   // bf16 stem on the tensor cores (default 1); 0: the CUDA-core kernels
    g_stem_tc = value;
    return HG_OK;
  }
  if (strcmp(name, "stem_bwd_blocks_per_sm") == 0 && value > 0) {
    g_stem_bwd_blocks_per_sm = value;
    return HG_OK;
  }
  if (strcmp(name, "persist_dynamic") == 0) {   // persistent kernels: tiles from a grid-wide counter instead of static ranges
    g_persist_dynamic = value;
    return HG_OK;
  }
  if (strcmp(name, "persist_transposed") == 0) {   // persistent 3x3 kernel, 128 output channels: [channel][pixel] accumulators
    g_persist_transposed = value;
    return HG_OK;
  }
  if (strcmp(name, "persist_min_units") == 0 && value > 0) {   // smallest launch (in 128-pixel units) it is used for
    g_persist_min_units = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_halo") == 0) {
    g_wgrad_halo = value;
    return HG_OK;
  }
  if (strcmp(name, "onewave_cluster") == 0) {
    g_onewave_cluster = value;
    return HG_OK;
  }
  if (strcmp(name, "upsample_fwd_cap") == 0) {
    g_upsample_fwd_cap = value;
    return HG_OK;
  }
  if (strcmp(name, "upsample_sep") == 0) {
    g_upsample_sep = value;
    return HG_OK;
  }
  if (strcmp(name, "ps_dbg") == 0) {
    g_ps_dbg = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_fused_bias") == 0) {
    g_wgrad_fused_bias = value;
    return HG_OK;
  }
  if (strcmp(name, "bn_apply_u4") == 0) {
    g_bn_apply_u4 = value;
    return HG_OK;
  }
  if (strcmp(name, "short_1stage") == 0) {
    g_short_1stage = value;
    return HG_OK;
  }
  if (strcmp(name, "long_k_3cta") == 0) {
    g_long_k_3cta = value;
    return HG_OK;
  }
  if (strcmp(name, "short_alias") == 0) {
    g_short_alias = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_big_n_panels") == 0) {
    g_wgrad_big_n_panels = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_small_n_panels") == 0) {
    g_wgrad_small_n_panels = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_t1_max_kb") == 0) {
    g_wgrad_t1_max_kb = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_dbg") == 0) {
    g_wgrad_dbg = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_kpx") == 0 && (value == 64 || value == 128)) {
    g_wgrad_kpx = value;
    return HG_OK;
  }
  if (strcmp(name, "mid_n_tiles") == 0) {
    g_mid_n_tiles = value;
    return HG_OK;
  }
  if (strcmp(name, "small_n_tiles") == 0) {
    g_small_n_tiles = value;
    return HG_OK;
  }
  if (strcmp(name, "single_wave_deep") == 0) {
    g_single_wave_deep = value;
    return HG_OK;
  }
  if (strcmp(name, "wgrad_smem_kb") == 0) {
    g_wgrad_smem_kb = value;
    return HG_OK;
  }
  if (strcmp(name, "pdl") == 0) {
    g_use_pdl = value;
    return HG_OK;
  }
  set_error("hg_set_option: unknown option '%s'", name);
  return HG_ERR_BAD_ARG;
}

int hg_pack_conv_weight_slice(const HgConvDesc* d, const float* w_oihw, int cin_total, int cin_offset, void* w_fprop,
                              void* w_dgrad, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(w_oihw != nullptr, "hg_pack_conv_weight: w_oihw is NULL");
  HG_REQUIRE(cin_offset >= 0 && cin_offset + d->Cin <= cin_total, "hg_pack_conv_weight_slice: slice out of range");
  cudaStream_t st = (cudaStream_t)stream;
  return d->dtype == HG_BF16 ? pack_weight<__nv_bfloat16>(d, w_oihw, w_fprop, w_dgrad, cin_total, cin_offset, st)
                             : pack_weight<float>(d, w_oihw, w_fprop, w_dgrad, cin_total, cin_offset, st);
}

int hg_pack_conv_weight(const HgConvDesc* d, const float* w_oihw, void* w_fprop, void* w_dgrad, void* stream) {
  return hg_pack_conv_weight_slice(d, w_oihw, d ? d->Cin : 0, 0, w_fprop, w_dgrad, stream);
}

int hg_mix_rows(const float* T, const float* in, float* out, int R, int cols, int transpose, int accumulate,
                void* stream) {
  HG_REQUIRE(T && in && out && R > 0 && cols > 0, "hg_mix_rows: bad arguments");
  HG_REQUIRE(in != out, "hg_mix_rows: in-place recombination is not supported");
  return mix_rows(T, in, out, R, R, cols, transpose, accumulate, (cudaStream_t)stream);
}

int hg_mix_rows_rect(const float* T, const float* in, float* out, int rows_out, int rows_in, int cols, int transpose,
                     int accumulate, void* stream) {
  HG_REQUIRE(T && in && out && rows_out > 0 && rows_in > 0 && cols > 0, "hg_mix_rows_rect: bad arguments");
  HG_REQUIRE(in != out, "hg_mix_rows_rect: in-place recombination is not supported");
  return mix_rows(T, in, out, rows_out, rows_in, cols, transpose, accumulate, (cudaStream_t)stream);
}

int hg_conv_fprop_ex(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias,
                     const void* residual, void* y, float* stats, float* out_nchw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(x && w_fprop && y, "hg_conv_fprop: x, w_fprop and y must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_eligible(d)) return conv_tc_fprop(d, x, w_fprop, bias, residual, y, stats, out_nchw, 0, nullptr, st);
  rc = ref_conv_allowed(d, "hg_conv_fprop");
  if (rc) return rc;
  rc = d->dtype == HG_BF16 ? conv_ref_fprop<__nv_bfloat16>(d, x, w_fprop, bias, residual, y, out_nchw, st)
                           : conv_ref_fprop<float>(d, x, w_fprop, bias, residual, y, out_nchw, st);
  if (rc) return rc;
  if (stats) {
    const int Ho = (d->H + 2 * d->pad - d->dil * (d->R - 1) - 1) / d->stride + 1;
    const int Wo = (d->W + 2 * d->pad - d->dil * (d->S - 1) - 1) / d->stride + 1;
    return bn_stats_launch(d->dtype, y, (long long)d->N * Ho * Wo, pad64(d->Cout), stats, st);
  }
  return HG_OK;
}

int hg_conv_fprop(const HgConvDesc* d, const void* x, const void* w_fprop, const float* bias, const void* residual,
                  void* y, float* stats, void* stream) {
  return hg_conv_fprop_ex(d, x, w_fprop, bias, residual, y, stats, nullptr, stream);
}

int hg_conv_dgrad(const HgConvDesc* d, const void* dy, const void* w_dgrad, const void* addend, void* dx,
                  void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(dy && w_dgrad && dx, "hg_conv_dgrad: dy, w_dgrad and dx must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_eligible(d)) return conv_tc_dgrad(d, dy, w_dgrad, addend, dx, nullptr, 0, nullptr, st);
  rc = ref_conv_allowed(d, "hg_conv_dgrad");
  if (rc) return rc;
  return d->dtype == HG_BF16 ? conv_ref_dgrad<__nv_bfloat16>(d, dy, w_dgrad, addend, dx, st)
                             : conv_ref_dgrad<float>(d, dy, w_dgrad, addend, dx, st);
}

static int check_fold(const HgConvDesc* d, const HgBnFold* bn, const char* who) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(bn != nullptr && bn->gamma && bn->beta, "%s: HgBnFold / gamma / beta is NULL", who);
  HG_REQUIRE(bn->use_running ? (bn->running_mean && bn->running_var) : (bn->stats != nullptr),
             "%s: statistics missing for the selected BatchNorm mode", who);
  if (!tc_eligible(d) || d->stride != 1) {
    set_error("%s: geometry not taken by the BatchNorm-fused tensor-core kernels (hg_conv_fold_eligible() == 0); run "
              "hg_bn_apply and the plain convolution instead", who);
    return HG_ERR_UNSUPPORTED;
  }
  return HG_OK;
}

int hg_conv_tc_eligible(const HgConvDesc* d) { return (d && check_desc(d) == HG_OK && tc_eligible(d)) ? 1 : 0; }
int hg_conv_fold_eligible(const HgConvDesc* d) {
  return (d && check_desc(d) == HG_OK && tc_eligible(d) && d->stride == 1) ? 1 : 0;
}

int hg_conv_fprop_bn(const HgConvDesc* d, const HgBnFold* bn, const void* x_raw, const void* w_fprop,
                     const float* bias, const void* residual, void* y, float* stats, float* out_nchw, void* stream) {
  int rc = check_fold(d, bn, "hg_conv_fprop_bn");
  if (rc) return rc;
  HG_REQUIRE(x_raw && w_fprop && y, "hg_conv_fprop_bn: x_raw, w_fprop and y must be non-NULL");
  const BnFoldDev f = make_fold(bn, d->Cin, (long long)d->N * d->H * d->W);
  return conv_tc_fprop(d, x_raw, w_fprop, bias, residual, y, stats, out_nchw, 1, &f, (cudaStream_t)stream);
}

int hg_conv_fprop_bnout(const HgConvDesc* d, const HgBnFold* bn_out, const void* x, const void* w_fprop,
                        const float* bias, void* y, float* out_nchw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(bn_out && bn_out->gamma && bn_out->beta && bn_out->use_running && bn_out->running_mean &&
                 bn_out->running_var,
             "hg_conv_fprop_bnout: an inference-mode BatchNorm (use_running = 1, running_mean / running_var / gamma / "
             "beta) is required");
  HG_REQUIRE(x && w_fprop && y, "hg_conv_fprop_bnout: x, w_fprop and y must be non-NULL");
  if (!tc_eligible(d)) {
    set_error("hg_conv_fprop_bnout: geometry not taken by the tensor-core kernels (hg_conv_tc_eligible() == 0); run "
              "the plain convolution and hg_bn_apply instead");
    return HG_ERR_UNSUPPORTED;
  }
  const BnFoldDev f = make_fold(bn_out, d->Cout, (long long)d->N * d->H * d->W);
  return conv_tc_fprop(d, x, w_fprop, bias, nullptr, y, nullptr, out_nchw, 3, &f, (cudaStream_t)stream);
}

int hg_conv_wgrad_bn(const HgConvDesc* d, const HgBnFold* bn, const void* x_raw, const void* dy, float* dw_packed,
                     float* dbias, void* stream) {
  int rc = check_fold(d, bn, "hg_conv_wgrad_bn");
  if (rc) return rc;
  HG_REQUIRE(x_raw && dy, "hg_conv_wgrad_bn: x_raw and dy must be non-NULL");
  const BnFoldDev f = make_fold(bn, d->Cin, (long long)d->N * d->H * d->W);
  return conv_wgrad_bf16(d, x_raw, dy, dw_packed, dbias, &f, (cudaStream_t)stream);
}

int hg_conv_dgrad_bn(const HgConvDesc* d, const HgBnFold* bn, const void* dy, const void* w_dgrad, const void* x_raw,
                     void* g, float* red, void* stream) {
  int rc = check_fold(d, bn, "hg_conv_dgrad_bn");
  if (rc) return rc;
  HG_REQUIRE(dy && w_dgrad && x_raw && g && red, "hg_conv_dgrad_bn: NULL pointer");
  const BnFoldDev f = make_fold(bn, d->Cin, (long long)d->N * d->H * d->W);
  return conv_tc_dgrad(d, dy, w_dgrad, x_raw, g, red, 2, &f, (cudaStream_t)stream);
}

int hg_unpack_conv_wgrad_slice(const HgConvDesc* d, const float* dw_packed, float* dw_oihw, int cin_total,
                               int cin_offset, int accumulate, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(dw_packed && dw_oihw, "hg_unpack_conv_wgrad: NULL pointer");
  HG_REQUIRE(cin_offset >= 0 && cin_offset + d->Cin <= cin_total, "hg_unpack_conv_wgrad_slice: slice out of range");
  return unpack_wgrad(d, dw_packed, dw_oihw, accumulate, cin_total, cin_offset, (cudaStream_t)stream);
}

int hg_unpack_conv_wgrad(const HgConvDesc* d, const float* dw_packed, float* dw_oihw, int accumulate, void* stream) {
  return hg_unpack_conv_wgrad_slice(d, dw_packed, dw_oihw, d ? d->Cin : 0, 0, accumulate, stream);
}

int hg_conv_wgrad(const HgConvDesc* d, const void* x, const void* dy, float* dw_oihw, float* dbias, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  HG_REQUIRE(x && dy, "hg_conv_wgrad: x and dy must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_eligible(d)) return conv_wgrad_bf16(d, x, dy, dw_oihw, dbias, nullptr, st);
  rc = ref_conv_allowed(d, "hg_conv_wgrad");
  if (rc) return rc;
  return d->dtype == HG_BF16 ? conv_ref_wgrad<__nv_bfloat16>(d, x, dy, dw_oihw, dbias, st)
                             : conv_ref_wgrad<float>(d, x, dy, dw_oihw, dbias, st);
}

}  // extern "C"
