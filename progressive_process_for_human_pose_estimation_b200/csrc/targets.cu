// Target rendering on the GPU: Gaussian keypoint heatmaps (variants G1-G6 of the reference datasets) and the
// integer limb / keypoint / background label maps PIL's ImageDraw produces.
//
// Gaussian: one thread per output element, evaluated in float64 exactly like the numpy code it replaces
// (centre = kp / w * 64, optional truncation, exp(-scale * d2 / (2 sigma^2)) in double, cast to float32 at the
// end: reference try_with_torch.py:107-132, try_with_torch_100.py:64-85, only_one_hourgless.py:112-132,
// hourglass_compare.py:286-313,713-734).
//
// Label maps: the draw order matters (later points / lines overwrite earlier ones), so one thread per image
// replays the draw list into a shared-memory canvas with PIL's own Bresenham stepping and the block writes
// the int64 map out (reference try_different_stack.py:114-155, try_skeleton_and_keypoints.py:93-114).
#include "hg_common.cuh"

namespace hg {

__device__ __forceinline__ double centre_of(double kp, double size, double grid, int center_mode, int truncate) {
  double c = center_mode == 0 ? kp / size * grid : kp * 256.0 / size / 4.0;
  if (truncate) c = trunc(c);
  return c;
}

__global__ void __launch_bounds__(256) render_gauss_kernel(HgGaussDesc d, const double* __restrict__ kp,
                                                           const int* __restrict__ num_persons,
                                                           const double* __restrict__ img_wh, float* __restrict__ out) {
  const long long total = (long long)d.B * d.J * d.H * d.W;
  const double denom = 2.0 * d.sigma * d.sigma;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % d.W);
    long long r = i / d.W;
    const int y = (int)(r % d.H);
    r /= d.H;
    const int j = (int)(r % d.J);
    const int b = (int)(r / d.J);
    const int np = num_persons ? num_persons[b] : d.P;
    const double iw = img_wh[2 * b], ih = img_wh[2 * b + 1];
    double acc = 0.0;
    const int p_begin = d.accumulate ? 0 : (np > 0 ? np - 1 : 0);
    for (int p = p_begin; p < np; ++p) {
      const double* k = kp + (((long long)b * d.P + p) * d.J + j) * 3;
      if (!(k[2] > 0.0)) continue;
      const double cx = centre_of(k[0], iw, (double)d.W, d.center_mode, d.truncate);
      const double cy = centre_of(k[1], ih, (double)d.H, d.center_mode, d.truncate);
      const double dx = (double)x - cx, dy = (double)y - cy;
      double t = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // no FMA contraction: numpy rounds each op
      if (d.pre_scale != 1.0) t = d.pre_scale * t;
      t = t / denom;
      const double g = d.amplitude == 1.0 ? exp(-t) : d.amplitude * exp(-t);
      acc = d.accumulate ? acc + g : g;
    }
    out[i] = (float)acc;
  }
}

// PIL ImageDraw point / line on an 8-bit canvas (Pillow src/libImaging/Draw.c: point8, line8; draw_lines adds
// the final end point).  Coordinates are truncated toward zero by the caller.
__device__ __forceinline__ void put8(unsigned char* cv, int W, int H, int x, int y, int ink) {
  if (x >= 0 && x < W && y >= 0 && y < H) cv[y * W + x] = (unsigned char)ink;
}
__device__ void line8(unsigned char* cv, int W, int H, int x0, int y0, int x1, int y1, int ink) {
  int dx = x1 - x0, dy = y1 - y0, xs = 1, ys = 1;
  if (dx < 0) { dx = -dx; xs = -1; }
  if (dy < 0) { dy = -dy; ys = -1; }
  if (dx == 0) {
    for (int i = 0; i < dy; ++i) { put8(cv, W, H, x0, y0, ink); y0 += ys; }
  } else if (dy == 0) {
    for (int i = 0; i < dx; ++i) { put8(cv, W, H, x0, y0, ink); x0 += xs; }
  } else if (dx > dy) {
    const int n = dx;
    dy += dy;
    int e = dy - dx;
    dx += dx;
    for (int i = 0; i < n; ++i) {
      put8(cv, W, H, x0, y0, ink);
      if (e >= 0) { y0 += ys; e -= dx; }
      e += dy;
      x0 += xs;
    }
  } else {
    const int n = dy;
    dx += dx;
    int e = dx - dy;
    dy += dy;
    for (int i = 0; i < n; ++i) {
      put8(cv, W, H, x0, y0, ink);
      if (e >= 0) { x0 += xs; e -= dy; }
      e += dx;
      y0 += ys;
    }
  }
  put8(cv, W, H, x1, y1, ink);  // ImageDraw.line draws the last point explicitly
}

__global__ void __launch_bounds__(128) render_labels_kernel(HgLabelDesc d, const double* __restrict__ kp,
                                                            const int* __restrict__ num_persons,
                                                            const double* __restrict__ img_wh,
                                                            const int* __restrict__ limbs, long long* __restrict__ out) {
  extern __shared__ unsigned char canvas[];
  const int b = blockIdx.x;
  const int npx = d.H * d.W;
  for (int i = threadIdx.x; i < npx; i += blockDim.x) canvas[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int np = num_persons ? num_persons[b] : d.P;
    const double iw = img_wh[2 * b], ih = img_wh[2 * b + 1];
    for (int p = 0; p < np; ++p) {
      const double* k = kp + ((long long)b * d.P + p) * d.J * 3;
      if (d.draw_points) {
        for (int j = 0; j < d.J; ++j) {
          if (!(k[3 * j + 2] > 0.0)) continue;
          if (d.draw_points == 2) {
            // MPII keypoint map (train.py:681-686): ImageDraw.ellipse((x-.5, y-.5, x+.5, y+.5)) of the FLOAT centre.
            // Pillow truncates the four box coordinates toward zero; a 1x1 / 1x2 / 2x1 / 2x2 box is filled
            // completely, a box that collapses to one point (both extents 0) draws nothing (Pillow 12.2 Draw.c).
            const double fx = centre_of(k[3 * j], iw, (double)d.W, d.center_mode, 0);
            const double fy = centre_of(k[3 * j + 1], ih, (double)d.H, d.center_mode, 0);
            const int x0 = (int)(fx - 0.5), y0 = (int)(fy - 0.5), x1 = (int)(fx + 0.5), y1 = (int)(fy + 0.5);
            if (x1 > x0 || y1 > y0)
              for (int yy = y0; yy <= y1; ++yy)
                for (int xx = x0; xx <= x1; ++xx) put8(canvas, d.W, d.H, xx, yy, j + 1);
            continue;
          }
          const int x = (int)centre_of(k[3 * j], iw, (double)d.W, d.center_mode, 1);
          const int y = (int)centre_of(k[3 * j + 1], ih, (double)d.H, d.center_mode, 1);
          put8(canvas, d.W, d.H, x, y, j + 1);
        }
      }
      if (d.draw_lines) {
        for (int l = 0; l < d.L; ++l) {
          const int a = limbs[2 * l], c = limbs[2 * l + 1];
          if (!(k[3 * a + 2] > 0.0) || !(k[3 * c + 2] > 0.0)) continue;
          const int x0 = (int)centre_of(k[3 * a], iw, (double)d.W, d.center_mode, 1);
          const int y0 = (int)centre_of(k[3 * a + 1], ih, (double)d.H, d.center_mode, 1);
          const int x1 = (int)centre_of(k[3 * c], iw, (double)d.W, d.center_mode, 1);
          const int y1 = (int)centre_of(k[3 * c + 1], ih, (double)d.H, d.center_mode, 1);
          line8(canvas, d.W, d.H, x0, y0, x1, y1, d.line_value > 0 ? d.line_value : (d.line_value < 0 ? l : l + 1));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npx; i += blockDim.x) out[(long long)b * npx + i] = (long long)canvas[i];
}

// ------------------------------------------------------------------------------------------------------
// Annotation -> keypoint tensor: the dataset's annotations live in HBM once; a batch of sample indices
// becomes the dense [B, P, J, 3] (x, y, v) tensor + num_persons[B] + img_wh[B, 2] the render kernels consume.
//   mode 0 (COCO, try_with_torch.py:103-113): table row = one person, J*3 doubles exactly as the JSON `keypoints` list;
//           CSR offsets per image; the LAST P persons are kept when an image has more (quirk Q7: last person wins).
//   mode 1 (MPII, hourglass_compare.py:691-703): table row = one annotated point (id, x, y, is_visible): scattered into
//           ONE person row, points_rect[id] = [x, y, is_visible != 0], later records of the same id overwrite earlier ones.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gather_annotations_kernel(HgAnnotDesc d, const double* __restrict__ table,
                                                                 const int* __restrict__ offset,
                                                                 const double* __restrict__ wh_all,
                                                                 const long long* __restrict__ index,
                                                                 double* __restrict__ kp, int* __restrict__ num_persons,
                                                                 double* __restrict__ img_wh) {
  const int b = blockIdx.x;
  const long long idx = index[b];
  const int beg = offset[idx], end = offset[idx + 1];
  double* dst = kp + (long long)b * d.P * d.J * 3;
  const int row = d.J * 3;
  if (threadIdx.x == 0) {
    img_wh[2 * b] = wh_all[2 * idx];
    img_wh[2 * b + 1] = wh_all[2 * idx + 1];
  }
  if (d.mode == 0) {
    const int n = end - beg;
    const int np = n < d.P ? n : d.P;
    const int first = end - np;
    for (int i = threadIdx.x; i < d.P * row; i += blockDim.x)
      dst[i] = i < np * row ? table[(long long)first * row + i] : 0.0;
    if (threadIdx.x == 0) num_persons[b] = np;
  } else {
    for (int i = threadIdx.x; i < d.P * row; i += blockDim.x) dst[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int r = beg; r < end; ++r) {
        const double* rec = table + (long long)r * 4;
        const int id = (int)rec[0];
        if (id >= 0 && id < d.J) {
          dst[id * 3] = rec[1];
          dst[id * 3 + 1] = rec[2];
          dst[id * 3 + 2] = rec[3] != 0.0 ? 1.0 : 0.0;
        }
      }
      num_persons[b] = 1;
    }
  }
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_render_gauss(const HgGaussDesc* d, const double* keypoints, const int32_t* num_persons, const double* img_wh,
                    float* out, void* stream) {
  HG_REQUIRE(d && keypoints && img_wh && out, "hg_render_gauss: NULL pointer");
  HG_REQUIRE(d->B > 0 && d->P > 0 && d->J > 0 && d->H > 0 && d->W > 0, "hg_render_gauss: non-positive size");
  HG_REQUIRE(d->center_mode == 0 || d->center_mode == 1, "hg_render_gauss: bad center_mode");
  HG_REQUIRE(d->sigma > 0.0, "hg_render_gauss: sigma must be positive");
  const long long total = (long long)d->B * d->J * d->H * d->W;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  render_gauss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*d, keypoints, num_persons, img_wh, out);
  HG_LAUNCH_OK("render_gauss_kernel");
  count_launch();
  return HG_OK;
}

int hg_render_labels(const HgLabelDesc* d, const double* keypoints, const int32_t* num_persons, const double* img_wh,
                     const int32_t* limbs, int64_t* out, void* stream) {
  HG_REQUIRE(d && keypoints && img_wh && out, "hg_render_labels: NULL pointer");
  HG_REQUIRE(d->B > 0 && d->P > 0 && d->J > 0 && d->H > 0 && d->W > 0 && d->L >= 0,
             "hg_render_labels: non-positive size");
  HG_REQUIRE(!d->draw_lines || limbs != nullptr, "hg_render_labels: limbs missing");
  HG_REQUIRE(d->H * d->W <= 48 * 1024, "hg_render_labels: canvas too large");
  HG_REQUIRE(d->J <= 255 && d->L <= 255, "hg_render_labels: 8-bit canvas holds at most 255 classes");
  render_labels_kernel<<<d->B, 128, d->H * d->W, (cudaStream_t)stream>>>(*d, keypoints, num_persons, img_wh, limbs,
                                                                          (long long*)out);
  HG_LAUNCH_OK("render_labels_kernel");
  count_launch();
  return HG_OK;
}

int hg_gather_annotations(const HgAnnotDesc* d, const double* table, const int32_t* offset, const double* wh_all,
                          const int64_t* sample_index, double* keypoints, int32_t* num_persons, double* img_wh,
                          void* stream) {
  HG_REQUIRE(d && d->B > 0 && d->P > 0 && d->J > 0 && (d->mode == 0 || d->mode == 1), "hg_gather_annotations: bad desc");
  HG_REQUIRE(table && offset && wh_all && sample_index && keypoints && num_persons && img_wh,
             "hg_gather_annotations: NULL pointer");
  gather_annotations_kernel<<<d->B, 128, 0, (cudaStream_t)stream>>>(*d, table, offset, wh_all,
                                                                    (const long long*)sample_index, keypoints,
                                                                    num_persons, img_wh);
  HG_LAUNCH_OK("gather_annotations_kernel");
  count_launch();
  return HG_OK;
}

}  // extern "C"
