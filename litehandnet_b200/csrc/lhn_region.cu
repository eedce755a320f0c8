// Region-map bbox decode and the point / window restricted keypoint refinements of the legacy parsers
// (SURVEY §8f rank 4): utils/result_parser.py:50-59,131-229,288-306, utils/SPheatmapParser.py:32-138,
// utils/evaluation.py:94-212, utils/HeatmapParser.py:197-223.
//
// These are low-volume ops (three planes per image, or a handful of points) that sit right after the keypoint
// decode; the reference runs them as Python loops over .tolist()-ed tensors with a torchvision NMS per image.
// Here one CTA per image (or per plane) keeps the plane in shared memory and does every step on the device.
#include <math_constants.h>

#include "lhn_common.cuh"
#include "lhn_heatmap.cuh"

namespace lhn {
namespace {

constexpr int kNT = 256;                 // threads per CTA
constexpr int kNW = kNT / 32;

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// running maximum of torch's max_pool: (val > max) || isnan(val) -> max = val  (NaN wins and sticks)
__device__ __forceinline__ float pool_max(float m, float v) { return (v > m || v != v) ? v : m; }

// ---- block-wide reductions -------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  unsigned long long r = red[0];
  const int nw = blockDim.x >> 5;
  for (int i = 1; i < nw; ++i) r = red[i] > r ? red[i] : r;
  __syncthreads();
  return r;
}

__device__ __forceinline__ float block_nanmax(float v, float* red) {   // np.max: NaN propagates
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  const int nw = blockDim.x >> 5;
  for (int i = 1; i < nw; ++i) r = nanmax(r, red[i]);
  __syncthreads();
  return r;
}

// (order key, lowest index first) as one comparable word
__device__ __forceinline__ unsigned long long composite(float v, uint32_t idx) {
  return ((unsigned long long)order_key(v) << 32) | (unsigned long long)(0xffffffffu - idx);
}

// ---- heatmap_nms on a plane resident in shared memory ----------------------------------------------------
// sP: the plane; sR: scratch (row maxima); sN: result = sP * eq(maxpool_k(sP), sP).  sN may be sP.
__device__ void nms_plane(const float* sP, float* sR, float* sN, int H, int W, int k) {
  const int p = (k - 1) >> 1, HW = H * W;
  for (int e = threadIdx.x; e < HW; e += blockDim.x) {
    const int y = e / W, x = e - y * W;
    float m = -CUDART_INF_F;
    const int lo = max(x - p, 0), hi = min(x + p, W - 1);
    for (int xx = lo; xx <= hi; ++xx) m = pool_max(m, sP[y * W + xx]);
    sR[e] = m;
  }
  __syncthreads();
  // sN may alias sP: the column pass reads sR, and sP[e] only through the thread that then writes sN[e]
  for (int e = threadIdx.x; e < HW; e += blockDim.x) {
    const int y = e / W, x = e - y * W;
    float m = -CUDART_INF_F;
    const int lo = max(y - p, 0), hi = min(y + p, H - 1);
    for (int yy = lo; yy <= hi; ++yy) m = pool_max(m, sR[yy * W + x]);
    const float v = sP[e];
    sN[e] = __fmul_rn(v, (m == v) ? 1.0f : 0.0f);        // NaN * 0 = NaN, inf * 0 = NaN: as heatmaps *= mask
  }
  __syncthreads();
}

// ---- legacy DARK (heatmap_post_processing.py:35-91) at a given point of a shared-memory plane ------------
// Row pass of the zero-padded f64 blur over the whole plane (cv2's order: sequential FMA over the taps).
__device__ void blur_rows_f64(const float* sN, double* sD, int H, int W, int ks, const double* taps) {
  const int HW = H * W;
  for (int e = threadIdx.x; e < HW; e += blockDim.x) {
    const int y = e / W, x = e - y * W;
    sD[e] = blur_row<double, false>(sN, H, W, y, x, ks, taps);
  }
  __syncthreads();
}
// Column pass at (x, y): centre tap, then symmetric pairs fused-added; the result is cast to f32 as the
// reference stores it back into its f32 heatmap.
__device__ __forceinline__ float blur_col_f64(const double* sD, int H, int W, int x, int y, int ks, const double* taps) {
  const int b = (ks - 1) >> 1;
  double acc = __dmul_rn(taps[b], sD[y * W + x]);
  for (int j = 1; j <= b; ++j) {
    const double up = (y - j >= 0) ? sD[(y - j) * W + x] : 0.0;
    const double dn = (y + j < H) ? sD[(y + j) * W + x] : 0.0;
    acc = __fma_rn(taps[b + j], __dadd_rn(dn, up), acc);
  }
  return (float)acc;
}
__device__ float blurred_plane_max(const double* sD, int H, int W, int ks, const double* taps, float* red) {
  const int HW = H * W;
  float m = -CUDART_INF_F;
  bool first = true;
  for (int e = threadIdx.x; e < HW; e += blockDim.x) {
    const int y = e / W, x = e - y * W;
    const float v = blur_col_f64(sD, H, W, x, y, ks, taps);
    m = first ? v : nanmax(m, v);
    first = false;
  }
  // threads without an element hold -inf: harmless unless every value is NaN-free and below -inf (impossible)
  return block_nanmax(m, red);
}
// One warp: Taylor step at (px, py) (guard already checked) with hm = log(max(blur * sc, 1e-10)).
__device__ void dark_legacy_at(const double* sD, int H, int W, int ks, const double* taps, int px, int py, float sc,
                               float* win /* 25 floats of this warp */, float& rx, float& ry) {
  const int lane = threadIdx.x & 31;
  if (lane < 25) {
    const int dy = lane / 5 - 2, dx = lane % 5 - 2;
    float v = __fmul_rn(blur_col_f64(sD, H, W, px + dx, py + dy, ks, taps), sc);
    v = (v != v) ? v : fmaxf(v, 1e-10f);                 // np.maximum propagates NaN
    win[lane] = logf(v);
  }
  __syncwarp();
  if (lane == 0) {
#define HH(dy, dx) win[((dy) + 2) * 5 + (dx) + 2]
    const float ddx = __fmul_rn(0.5f, __fsub_rn(HH(0, 1), HH(0, -1)));
    const float ddy = __fmul_rn(0.5f, __fsub_rn(HH(1, 0), HH(-1, 0)));
    const float dxx = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(HH(0, 2), __fmul_rn(2.f, HH(0, 0))), HH(0, -2)));
    const float dxy = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(__fsub_rn(HH(1, 1), HH(-1, 1)), HH(1, -1)), HH(-1, -1)));
    const float dyy = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(HH(2, 0), __fmul_rn(2.f, HH(0, 0))), HH(-2, 0)));
#undef HH
    const float det = __fsub_rn(__fmul_rn(dxx, dyy), __fmul_rn(dxy, dxy));
    if (det != 0.f) {                                    // true for NaN, as in numpy
      const float ox = -__fdiv_rn(__fsub_rn(__fmul_rn(dyy, ddx), __fmul_rn(dxy, ddy)), det);
      const float oy = -__fdiv_rn(__fsub_rn(__fmul_rn(dxx, ddy), __fmul_rn(dxy, ddx)), det);
      rx = __fadd_rn(rx, ox); ry = __fadd_rn(ry, oy);
    }
  }
  __syncwarp();
}

// ---- non_max_suppression (result_parser.py:177-215 == SPheatmapParser.py:101-138 == evaluation.py:170-212) --
// filter (confidence > det_thr, min_wh < w, h < max_wh), xywh -> xyxy (bbox_metric.py:33-40), torchvision.ops.nms
// (stable sort by score descending, greedy suppression of IoU > iou_thr; f32 arithmetic, the IoU compared with the
// double threshold), keep the first max_num.  Sequential: run by ONE thread.
struct BoxScratch {
  float bx[LHN_MAX_CANDIDATES][5];        // x1, y1, x2, y2, area
  int sel[LHN_MAX_CANDIDATES];
  unsigned char sup[LHN_MAX_CANDIDATES];
};
__device__ void box_filter_nms(const float* cand, int N, float det_thr, float min_wh, float max_wh, double iou_thr,
                               int max_num, BoxScratch* s, float* out, int32_t* count) {
  int n = 0;
  for (int k = 0; k < N; ++k) {
    const float* c = cand + 5 * k;
    if (c[4] > det_thr && c[2] > min_wh && c[2] < max_wh && c[3] > min_wh && c[3] < max_wh) s->sel[n++] = k;
  }
  for (int i = 1; i < n; ++i) {                           // stable insertion sort, score descending
    const int k = s->sel[i];
    const float sc = cand[5 * k + 4];
    int j = i - 1;
    while (j >= 0 && cand[5 * s->sel[j] + 4] < sc) { s->sel[j + 1] = s->sel[j]; --j; }
    s->sel[j + 1] = k;
  }
  for (int i = 0; i < n; ++i) {
    const float* c = cand + 5 * s->sel[i];
    float* q = s->bx[i];
    q[0] = __fsub_rn(c[0], __fdiv_rn(c[2], 2.f)); q[1] = __fsub_rn(c[1], __fdiv_rn(c[3], 2.f));
    q[2] = __fadd_rn(c[0], __fdiv_rn(c[2], 2.f)); q[3] = __fadd_rn(c[1], __fdiv_rn(c[3], 2.f));
    q[4] = __fmul_rn(__fsub_rn(q[2], q[0]), __fsub_rn(q[3], q[1]));
    s->sup[i] = 0;
  }
  int kept = 0;
  for (int i = 0; i < n && kept < max_num; ++i) {         // boxes past the first max_num cannot change the result
    if (s->sup[i]) continue;
    const float* c = cand + 5 * s->sel[i];
    for (int j = 0; j < 5; ++j) out[kept * 5 + j] = c[j];
    ++kept;
    const float* qi = s->bx[i];
    for (int j = i + 1; j < n; ++j) {
      if (s->sup[j]) continue;
      const float* qj = s->bx[j];
      const float w = fmaxf(0.f, __fsub_rn(fminf(qi[2], qj[2]), fmaxf(qi[0], qj[0])));
      const float h = fmaxf(0.f, __fsub_rn(fminf(qi[3], qj[3]), fmaxf(qi[1], qj[1])));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(qi[4], qj[4]), inter));
      if ((double)ovr > iou_thr) s->sup[j] = 1;
    }
  }
  for (int i = kept; i < max_num; ++i)
    for (int j = 0; j < 5; ++j) out[i * 5 + j] = 0.f;
  *count = kept;
}

// ==========================================================================================================
// lhn_region_bbox_decode: one CTA per image
// ==========================================================================================================
struct RegionArgs {
  const void* center; const void* size; void* nms_out;
  float* candidates; float* boxes; int32_t* counts;
  int64_t center_stride_b, size_stride_b, size_stride_c;
  int H, W;
  lhn_region_params rp;
};

struct RegionSmem {
  unsigned long long red64[kNW];
  float redf[kNW];
  uint32_t top_idx[LHN_MAX_CANDIDATES];
  float cand[LHN_MAX_CANDIDATES][5];
  float win[kNW][25];
  BoxScratch box;
};

template <typename T>
__global__ void __launch_bounds__(kNT) region_bbox_kernel(const __grid_constant__ RegionArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const lhn_region_params& rp = a.rp;
  const int H = a.H, W = a.W, HW = H * W, N = rp.num_candidates;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x;
  RegionSmem* sh = reinterpret_cast<RegionSmem*>(smem_raw);
  float* sP = reinterpret_cast<float*>(smem_raw + align_up(sizeof(RegionSmem), 16));
  float* sN = sP + HW;
  double* sD = reinterpret_cast<double*>(sN + HW);       // row maxima (as float) / f64 blur rows

  const T* cplane = reinterpret_cast<const T*>(a.center) + b * a.center_stride_b;
  for (int e = tid; e < HW; e += kNT) sP[e] = Elem<T>::to_f32(cplane[e]);
  __syncthreads();

  // ---- heatmap_nms ----
  if (rp.nms_kernel > 0) nms_plane(sP, reinterpret_cast<float*>(sD), sN, H, W, rp.nms_kernel);
  else sN = sP;
  if (a.nms_out) {
    T* o = reinterpret_cast<T*>(a.nms_out) + b * a.center_stride_b;
    for (int e = tid; e < HW; e += kNT) o[e] = from_f32<T>(sN[e]);
  }

  // ---- torch.topk(k = N): descending, NaN greatest, equal values lowest index first ----
  unsigned long long last = 0;
  for (int r = 0; r < N; ++r) {
    unsigned long long best = 0;
    for (int e = tid; e < HW; e += kNT) {
      const unsigned long long c = composite(sN[e], (uint32_t)e);
      if ((r == 0 || c < last) && c > best) best = c;
    }
    best = block_max_u64(best, sh->red64);
    last = best;
    if (tid == 0) sh->top_idx[r] = 0xffffffffu - (uint32_t)(best & 0xffffffffull);
  }
  __syncthreads();

  // ---- legacy DARK of the candidate centres (ResultParser, cfg['DARK']) ----
  const bool dark = rp.mode == LHN_REGION_RP && rp.refine == LHN_REFINE_DARK_LEGACY;
  float sc = 1.0f;
  if (dark) {
    blur_rows_f64(sN, sD, H, W, rp.blur_ksize, rp.taps);
    const float bmax = blurred_plane_max(sD, H, W, rp.blur_ksize, rp.taps, sh->redf);
    const float origin_max = sN[sh->top_idx[0]];          // np.max of the plane (NaN included) = the top-1 value
    sc = __fdiv_rn(origin_max, __fadd_rn(bmax, 1e-6f));
  }

  // ---- candidates: one warp each ----
  const T* s0 = reinterpret_cast<const T*>(a.size) + b * a.size_stride_b;
  const T* s1 = s0 + a.size_stride_c;
  for (int k = warp; k < N; k += kNW) {
    const int idx = (int)sh->top_idx[k];
    const int x = idx % W, y = idx / W;
    const float conf = sN[idx];
    float cx = (float)x, cy = (float)y, cw = 0.f, ch = 0.f;
    if (rp.mode == LHN_REGION_CS) {
      if (conf > rp.cand_thr) {
        // _get_wh: the window [c - 6, c + 7) clipped to [0, hs - 1] on both ends (the clip drops the last row/column)
        const int x1 = min(max(x - 6, 0), W - 1), x2 = min(max(x + 7, 0), W - 1);
        const int y1 = min(max(y - 6, 0), W - 1), y2 = min(min(max(y + 7, 0), W - 1), H);
        const int ww = x2 - x1, hh = y2 - y1, n = (ww > 0 && hh > 0) ? ww * hh : 0;
        double sx = 0.0, sy = 0.0;
        for (int e = lane; e < n; e += 32) {
          const int yy = y1 + e / ww, xx = x1 + e % ww;
          sx += (double)Elem<T>::to_f32(s0[yy * W + xx]);
          sy += (double)Elem<T>::to_f32(s1[yy * W + xx]);
        }
        sx = warp_sum(sx); sy = warp_sum(sy);
        const float gx = n ? (float)(sx / (double)n) : CUDART_NAN_F;
        const float gy = n ? (float)(sy / (double)n) : CUDART_NAN_F;
        cw = __fdiv_rn(__fmul_rn(gx, rp.image_w), (float)W);
        ch = __fdiv_rn(__fmul_rn(gy, rp.image_w), (float)W);
        const float up = (float)((double)rp.image_w / (double)W);
        cx = __fmul_rn(cx, up); cy = __fmul_rn(cy, up);
      } else {
        cx = 0.f; cy = 0.f;
      }
    } else {
      // AvgPool2d(k, 1, (k-1)/2), count_include_pad: f32 sum of the in-bounds window in row-major order / k^2
      const int p = (rp.avg_kernel - 1) >> 1;
      float sw = 0.f, shh = 0.f;
      for (int yy = max(y - p, 0); yy < min(y - p + rp.avg_kernel, H); ++yy)
        for (int xx = max(x - p, 0); xx < min(x - p + rp.avg_kernel, W); ++xx) {
          sw = __fadd_rn(sw, Elem<T>::to_f32(s0[yy * W + xx]));
          shh = __fadd_rn(shh, Elem<T>::to_f32(s1[yy * W + xx]));
        }
      const float kk = (float)(rp.avg_kernel * rp.avg_kernel);
      cw = __fdiv_rn(sw, kk); ch = __fdiv_rn(shh, kk);
      if (rp.mode == LHN_REGION_SH) {
        cw = (cw != cw) ? cw : fminf(fmaxf(cw, 0.f), 0.99f);      // torch.clip propagates NaN
        ch = (ch != ch) ? ch : fminf(fmaxf(ch, 0.f), 0.99f);
        cx = __fmul_rn(cx, __fdiv_rn(rp.image_w, (float)W));
        cy = __fmul_rn(cy, __fdiv_rn(rp.image_h, (float)H));
        cw = __fmul_rn(cw, rp.image_w); ch = __fmul_rn(ch, rp.image_h);
      } else {
        if (dark && 1 < x && x < W - 2 && 1 < y && y < H - 2)
          dark_legacy_at(sD, H, W, rp.blur_ksize, rp.taps, x, y, sc, sh->win[warp], cx, cy);
        cx = __fmul_rn(cx, rp.stride_x); cy = __fmul_rn(cy, rp.stride_y);
        cw = __fmul_rn(cw, rp.stride_x); ch = __fmul_rn(ch, rp.stride_y);
      }
    }
    if (lane == 0) {
      float* c = sh->cand[k];
      c[0] = cx; c[1] = cy; c[2] = cw; c[3] = ch; c[4] = conf;
    }
  }
  __syncthreads();
  if (a.candidates)
    for (int e = tid; e < N * 5; e += kNT) a.candidates[b * N * 5 + e] = sh->cand[e / 5][e % 5];

  // ---- non_max_suppression: filter, xywh -> xyxy, torchvision.ops.nms, keep the first max_num_bbox ----
  if (tid == 0)
    box_filter_nms(&sh->cand[0][0], N, rp.det_thr, rp.min_wh, rp.max_wh, rp.iou_thr, rp.max_num_bbox, &sh->box,
                   a.boxes + b * rp.max_num_bbox * 5, a.counts + b);
}

// lhn_box_nms: one thread per image (N <= 32 candidates: a few hundred scalar operations)
__global__ void box_nms_kernel(const float* cand, int64_t B, int N, float det_thr, float min_wh, float max_wh,
                               double iou_thr, int max_num, float* boxes, int32_t* counts) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  BoxScratch s;
  box_filter_nms(cand + b * N * 5, N, det_thr, min_wh, max_wh, iou_thr, max_num, &s, boxes + b * max_num * 5, counts + b);
}

// ==========================================================================================================
// lhn_heatmap_nms: one CTA per plane
// ==========================================================================================================
template <typename T>
__global__ void __launch_bounds__(kNT) heatmap_nms_kernel(const T* hm, T* out, int C, int H, int W, int64_t stride_b,
                                                          int64_t stride_c, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int HW = H * W;
  float* sP = reinterpret_cast<float*>(smem_raw);
  float* sR = sP + HW;
  const int64_t p = blockIdx.x, b = p / C, c = p - b * C;
  const T* src = hm + b * stride_b + c * stride_c;
  T* dst = out + b * stride_b + c * stride_c;
  for (int e = threadIdx.x; e < HW; e += kNT) sP[e] = Elem<T>::to_f32(src[e]);
  __syncthreads();
  nms_plane(sP, sR, sP, H, W, k);
  for (int e = threadIdx.x; e < HW; e += kNT) dst[e] = from_f32<T>(sP[e]);
}

// ==========================================================================================================
// lhn_vector_nms: one thread per element
// ==========================================================================================================
template <typename T>
__global__ void vector_nms_kernel(const T* v, T* out, int64_t n, int L) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int i = (int)(e % L);
  const float c = Elem<T>::to_f32(v[e]);
  float m = -CUDART_INF_F;
  if (i > 0) m = pool_max(m, Elem<T>::to_f32(v[e - 1]));
  m = pool_max(m, c);
  if (i + 1 < L) m = pool_max(m, Elem<T>::to_f32(v[e + 1]));
  out[e] = from_f32<T>(__fmul_rn(c, (m == c) ? 1.0f : 0.0f));
}

// ==========================================================================================================
// lhn_refine_points: one thread per point
// ==========================================================================================================
template <typename T>
__global__ void refine_points_kernel(const T* hm, int64_t B, int C, int H, int W, int64_t stride_b, int64_t stride_c,
                                     const int32_t* bc, float* xy, int xy_stride, int64_t n, int refine) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = bc[2 * i], c = bc[2 * i + 1];
  if (b < 0 || b >= B || c < 0 || c >= C) return;
  const T* P = hm + (int64_t)b * stride_b + (int64_t)c * stride_c;
  float x = xy[i * xy_stride], y = xy[i * xy_stride + 1];
  // int() truncates; the reference would raise IndexError outside the plane — clamp instead of faulting
  const int xx = min(max((int)x, 0), W - 1), yy = min(max((int)y, 0), H - 1);
  auto at = [&](int r, int q) { return Elem<T>::to_f32(P[r * W + q]); };
  x += (at(yy, min(xx + 1, W - 1)) > at(yy, max(xx - 1, 0))) ? 0.25f : -0.25f;
  y += (at(min(yy + 1, H - 1), xx) > at(max(yy - 1, 0), xx)) ? 0.25f : -0.25f;
  if (refine == LHN_REFINE_OFFSET_HALF) { x += 0.5f; y += 0.5f; }
  xy[i * xy_stride] = x; xy[i * xy_stride + 1] = y;
}

// ==========================================================================================================
// lhn_decode_heatmap_roi: one CTA per plane, the crop dense in shared memory
// ==========================================================================================================
struct RoiArgs {
  const void* hm; const int32_t* roi; float* out; int32_t* out_idx;
  int64_t stride_b, stride_c;
  int K, H, W, refine, ksize;
  float scale_x, scale_y;
  double taps[LHN_MAX_TAPS];
};

template <typename T>
__global__ void __launch_bounds__(kNT) decode_roi_kernel(const __grid_constant__ RoiArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long red64[kNW];
  __shared__ float redf[kNW];
  __shared__ float win[25];
  const int tid = threadIdx.x;
  const int64_t p = blockIdx.x, b = p / a.K, k = p - b * a.K;
  int x0 = a.roi[4 * b], y0 = a.roi[4 * b + 1], x1 = a.roi[4 * b + 2], y1 = a.roi[4 * b + 3];
  x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, a.W); y1 = min(y1, a.H);
  if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = a.W; y1 = a.H; }   // empty window: the whole plane (:301-303)
  const int w = x1 - x0, h = y1 - y0, hw = w * h;
  float* sP = reinterpret_cast<float*>(smem_raw);
  double* sD = reinterpret_cast<double*>(smem_raw + align_up((size_t)a.H * a.W * 4, 16));
  const T* src = reinterpret_cast<const T*>(a.hm) + b * a.stride_b + k * a.stride_c;
  for (int e = tid; e < hw; e += kNT) {
    const int y = e / w, x = e - y * w;
    sP[e] = Elem<T>::to_f32(src[(y0 + y) * a.W + x0 + x]);
  }
  __syncthreads();
  // first-index argmax of the crop in its own row-major order (topk(k=1) of the reshaped crop)
  unsigned long long best = 0;
  for (int e = tid; e < hw; e += kNT) {
    const unsigned long long c = composite(sP[e], (uint32_t)e);
    best = c > best ? c : best;
  }
  best = block_max_u64(best, red64);
  const int idx = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
  const int px = idx % w, py = idx / w;
  const float maxval = sP[idx];
  float rx = (float)px, ry = (float)py;
  if (a.refine == LHN_REFINE_DARK_LEGACY && 1 < px && px < w - 2 && 1 < py && py < h - 2) {
    blur_rows_f64(sP, sD, h, w, a.ksize, a.taps);
    const float bmax = blurred_plane_max(sD, h, w, a.ksize, a.taps, redf);
    const float sc = __fdiv_rn(maxval, __fadd_rn(bmax, 1e-6f));
    if (tid < 32) dark_legacy_at(sD, h, w, a.ksize, a.taps, px, py, sc, win, rx, ry);
  }
  if (tid != 0) return;
  if (a.refine == LHN_REFINE_OFFSET || a.refine == LHN_REFINE_OFFSET_HALF) {
    rx += (sP[py * w + min(px + 1, w - 1)] > sP[py * w + max(px - 1, 0)]) ? 0.25f : -0.25f;
    ry += (sP[min(py + 1, h - 1) * w + px] > sP[max(py - 1, 0) * w + px]) ? 0.25f : -0.25f;
    if (a.refine == LHN_REFINE_OFFSET_HALF) { rx += 0.5f; ry += 0.5f; }
  }
  float* o = a.out + 3 * p;
  o[0] = __fmul_rn(__fadd_rn(rx, (float)x0), a.scale_x);
  o[1] = __fmul_rn(__fadd_rn(ry, (float)y0), a.scale_y);
  o[2] = maxval;
  if (a.out_idx) a.out_idx[p] = idx;
}

// ==========================================================================================================
// lhn_render_region_wh: the width/height planes of SRHandNet's region map (generateTarget.py:350-365)
// ==========================================================================================================
__global__ void render_region_wh_kernel(const int32_t* __restrict__ rect, const float* __restrict__ gamma, int H, int W,
                                        float* __restrict__ out, int64_t out_stride_b, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int HW = H * W;
  const int64_t b = e / (2 * HW);
  const int r = (int)(e - b * 2 * HW), c = r / HW, q = r - c * HW, y = q / W, x = q - y * W;
  const int32_t* rc = rect + 4 * b;
  const bool in = x >= rc[0] && x < rc[1] && y >= rc[2] && y < rc[3];
  out[b * out_stride_b + (int64_t)c * HW + q] = in ? gamma[2 * b + c] : 0.f;
}

// ==========================================================================================================
// lhn_dark_refine_points: the legacy DARK at GIVEN positions, one CTA per point
// ==========================================================================================================
struct DarkPointsArgs {
  const void* hm; const int32_t* bc; float* xy;
  int64_t stride_b, stride_c, B;
  int C, H, W, xy_stride, ksize;
  double taps[LHN_MAX_TAPS];
};

template <typename T>
__global__ void __launch_bounds__(kNT) dark_points_kernel(const __grid_constant__ DarkPointsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float redf[kNW];
  __shared__ float win[25];
  const int tid = threadIdx.x, HW = a.H * a.W;
  const int64_t i = blockIdx.x;
  const int b = a.bc[2 * i], c = a.bc[2 * i + 1];
  if (b < 0 || b >= a.B || c < 0 || c >= a.C) return;
  float* xyp = a.xy + i * a.xy_stride;
  float rx = xyp[0], ry = xyp[1];
  const int px = (int)rx, py = (int)ry;                      // int() truncates
  if (!(1 < px && px < a.W - 2 && 1 < py && py < a.H - 2)) return;   // taylor()'s guard: coordinates unchanged
  float* sP = reinterpret_cast<float*>(smem_raw);
  double* sD = reinterpret_cast<double*>(smem_raw + align_up((size_t)HW * 4, 16));
  const T* src = reinterpret_cast<const T*>(a.hm) + (int64_t)b * a.stride_b + (int64_t)c * a.stride_c;
  float m = -CUDART_INF_F;
  bool first = true;
  for (int e = tid; e < HW; e += kNT) {
    const float v = Elem<T>::to_f32(src[e]);
    sP[e] = v;
    m = first ? v : nanmax(m, v);
    first = false;
  }
  const float origin_max = block_nanmax(m, redf);            // np.max of the plane (ends with a CTA barrier)
  blur_rows_f64(sP, sD, a.H, a.W, a.ksize, a.taps);
  const float bmax = blurred_plane_max(sD, a.H, a.W, a.ksize, a.taps, redf);
  const float sc = __fdiv_rn(origin_max, __fadd_rn(bmax, 1e-6f));
  if (tid < 32) {
    dark_legacy_at(sD, a.H, a.W, a.ksize, a.taps, px, py, sc, win, rx, ry);
    if (tid == 0) { xyp[0] = rx; xyp[1] = ry; }
  }
}

template <typename K> int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return LHN_EINVAL;
  if (bytes > 40 * 1024) {                                  // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_last_error(e); return LHN_ECUDA; }
  }
  return LHN_OK;
}

}  // namespace
}  // namespace lhn

using namespace lhn;

#define LHN_DISPATCH_DTYPE(dtype, ...)                                  \
  switch (dtype) {                                                      \
    case LHN_F32: { using T = float; __VA_ARGS__; break; }              \
    case LHN_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }     \
    case LHN_F16: { using T = __half; __VA_ARGS__; break; }             \
    default: return LHN_EDTYPE;                                         \
  }

extern "C" int lhn_region_bbox_decode(const void* center, const void* size, int dtype, int64_t B, int H, int W,
                                      int64_t center_stride_b, int64_t size_stride_b, int64_t size_stride_c,
                                      const lhn_region_params* rp, void* nms_out, float* candidates,
                                      float* boxes, int32_t* counts, lhn_stream_t stream) {
  if (!center || !size || !rp || !boxes || !counts || B < 0 || H < 1 || W < 1) return LHN_EINVAL;
  if (rp->mode < LHN_REGION_SH || rp->mode > LHN_REGION_CS) return LHN_EINVAL;
  if (rp->num_candidates < 1 || rp->num_candidates > LHN_MAX_CANDIDATES || rp->num_candidates > H * W) return LHN_EINVAL;
  if (rp->max_num_bbox < 1 || rp->max_num_bbox > rp->num_candidates) return LHN_EINVAL;
  if (rp->nms_kernel < 0 || (rp->nms_kernel > 0 && (rp->nms_kernel & 1) == 0)) return LHN_EINVAL;
  if (rp->mode != LHN_REGION_CS && (rp->avg_kernel < 1 || (rp->avg_kernel & 1) == 0)) return LHN_EINVAL;
  if (rp->mode == LHN_REGION_CS && H != W) return LHN_EINVAL;     // cs_from_region_map uses shape[-1] for both axes
  const bool dark = rp->mode == LHN_REGION_RP && rp->refine == LHN_REFINE_DARK_LEGACY;
  if (rp->mode == LHN_REGION_RP && rp->refine != LHN_REFINE_NONE && !dark) return LHN_EINVAL;
  if (dark && (rp->blur_ksize < 3 || rp->blur_ksize > LHN_MAX_TAPS || (rp->blur_ksize & 1) == 0)) return LHN_EINVAL;
  if (B == 0) return LHN_OK;
  RegionArgs a;
  a.center = center; a.size = size; a.nms_out = nms_out; a.candidates = candidates; a.boxes = boxes; a.counts = counts;
  a.center_stride_b = center_stride_b; a.size_stride_b = size_stride_b; a.size_stride_c = size_stride_c;
  a.H = H; a.W = W; a.rp = *rp;
  const size_t HW = (size_t)H * W;
  const size_t smem = align_up(sizeof(RegionSmem), 16) + 2 * HW * 4 + (dark ? HW * 8 : HW * 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    int rc = set_smem(region_bbox_kernel<T>, smem);
    if (rc) return rc;
    region_bbox_kernel<T><<<(unsigned)B, kNT, smem, st>>>(a);
  });
  return check_launch();
}

extern "C" int lhn_heatmap_nms(const void* hm, void* out, int dtype, int64_t B, int C, int H, int W,
                               int64_t stride_b, int64_t stride_c, int nms_kernel, lhn_stream_t stream) {
  if (!hm || !out || B < 0 || C < 1 || H < 1 || W < 1 || nms_kernel < 1 || (nms_kernel & 1) == 0) return LHN_EINVAL;
  if (B == 0) return LHN_OK;
  const size_t smem = 2 * (size_t)H * W * 4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    int rc = set_smem(heatmap_nms_kernel<T>, smem);
    if (rc) return rc;
    heatmap_nms_kernel<T><<<(unsigned)(B * C), kNT, smem, st>>>(reinterpret_cast<const T*>(hm), reinterpret_cast<T*>(out),
                                                                C, H, W, stride_b, stride_c, nms_kernel);
  });
  return check_launch();
}

extern "C" int lhn_vector_nms(const void* v, void* out, int dtype, int64_t n_rows, int L, lhn_stream_t stream) {
  if (!v || !out || v == out || n_rows < 0 || L < 1) return LHN_EINVAL;
  const int64_t n = n_rows * L;
  if (n == 0) return LHN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    vector_nms_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const T*>(v),
                                                                      reinterpret_cast<T*>(out), n, L);
  });
  return check_launch();
}

extern "C" int lhn_refine_points(const void* hm, int dtype, int64_t B, int C, int H, int W, int64_t stride_b,
                                 int64_t stride_c, const int32_t* bc, float* xy, int xy_stride, int64_t n,
                                 int refine, lhn_stream_t stream) {
  if (!hm || !bc || !xy || B < 1 || C < 1 || H < 1 || W < 1 || xy_stride < 2 || n < 0) return LHN_EINVAL;
  if (refine != LHN_REFINE_OFFSET && refine != LHN_REFINE_OFFSET_HALF) return LHN_EINVAL;
  if (n == 0) return LHN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    refine_points_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(reinterpret_cast<const T*>(hm), B, C, H, W,
                                                                         stride_b, stride_c, bc, xy, xy_stride, n, refine);
  });
  return check_launch();
}

extern "C" int lhn_decode_heatmap_roi(const void* hm, int dtype, int64_t B, int K, int H, int W, int64_t stride_b,
                                      int64_t stride_c, const int32_t* roi, const lhn_decode_params* dp,
                                      float* out, int32_t* out_idx, lhn_stream_t stream) {
  if (!hm || !roi || !dp || !out || B < 0 || K < 1 || H < 1 || W < 1) return LHN_EINVAL;
  const int r = dp->refine;
  if (r != LHN_REFINE_NONE && r != LHN_REFINE_OFFSET && r != LHN_REFINE_OFFSET_HALF && r != LHN_REFINE_DARK_LEGACY)
    return LHN_EINVAL;
  const bool dark = r == LHN_REFINE_DARK_LEGACY;
  if (dark && (dp->blur_ksize < 3 || dp->blur_ksize > LHN_MAX_TAPS || (dp->blur_ksize & 1) == 0)) return LHN_EINVAL;
  if (B == 0) return LHN_OK;
  RoiArgs a;
  a.hm = hm; a.roi = roi; a.out = out; a.out_idx = out_idx; a.stride_b = stride_b; a.stride_c = stride_c;
  a.K = K; a.H = H; a.W = W; a.refine = r; a.ksize = dp->blur_ksize;
  a.scale_x = dp->transform == LHN_XFORM_SCALE ? dp->scale_x : 1.0f;
  a.scale_y = dp->transform == LHN_XFORM_SCALE ? dp->scale_y : 1.0f;
  for (int i = 0; i < LHN_MAX_TAPS; ++i) a.taps[i] = dp->taps[i];
  const size_t HW = (size_t)H * W;
  const size_t smem = align_up(HW * 4, 16) + (dark ? HW * 8 : 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    int rc = set_smem(decode_roi_kernel<T>, smem);
    if (rc) return rc;
    decode_roi_kernel<T><<<(unsigned)(B * K), kNT, smem, st>>>(a);
  });
  return check_launch();
}

extern "C" int lhn_box_nms(const float* candidates, int64_t B, int N, float det_thr, float min_wh, float max_wh,
                           double iou_thr, int max_num, float* boxes, int32_t* counts, lhn_stream_t stream) {
  if (!candidates || !boxes || !counts || B < 0 || N < 1 || N > LHN_MAX_CANDIDATES || max_num < 1 || max_num > N)
    return LHN_EINVAL;
  if (B == 0) return LHN_OK;
  box_nms_kernel<<<(unsigned)((B + 31) / 32), 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      candidates, B, N, det_thr, min_wh, max_wh, iou_thr, max_num, boxes, counts);
  return check_launch();
}

extern "C" int lhn_render_region_wh(const int32_t* rect, const float* gamma, int64_t B, int H, int W, float* out,
                                    int64_t out_stride_b, lhn_stream_t stream) {
  if (!rect || !gamma || !out || B < 0 || H < 1 || W < 1 || out_stride_b < (int64_t)2 * H * W) return LHN_EINVAL;
  const int64_t n = B * 2 * H * W;
  if (n == 0) return LHN_OK;
  render_region_wh_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      rect, gamma, H, W, out, out_stride_b, n);
  return check_launch();
}

extern "C" int lhn_dark_refine_points(const void* hm, int dtype, int64_t B, int C, int H, int W, int64_t stride_b,
                                      int64_t stride_c, const int32_t* bc, float* xy, int xy_stride, int64_t n,
                                      const lhn_decode_params* dp, lhn_stream_t stream) {
  if (!hm || !bc || !xy || !dp || B < 1 || C < 1 || H < 1 || W < 1 || xy_stride < 2 || n < 0) return LHN_EINVAL;
  if (dp->refine != LHN_REFINE_DARK_LEGACY) return LHN_EINVAL;
  if (dp->blur_ksize < 3 || dp->blur_ksize > LHN_MAX_TAPS || (dp->blur_ksize & 1) == 0) return LHN_EINVAL;
  if (n == 0) return LHN_OK;
  DarkPointsArgs a;
  a.hm = hm; a.bc = bc; a.xy = xy; a.stride_b = stride_b; a.stride_c = stride_c; a.B = B; a.C = C; a.H = H; a.W = W;
  a.xy_stride = xy_stride; a.ksize = dp->blur_ksize;
  for (int i = 0; i < LHN_MAX_TAPS; ++i) a.taps[i] = dp->taps[i];
  const size_t HW = (size_t)H * W;
  const size_t smem = align_up(HW * 4, 16) + HW * 8;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LHN_DISPATCH_DTYPE(dtype, {
    int rc = set_smem(dark_points_kernel<T>, smem);
    if (rc) return rc;
    dark_points_kernel<T><<<(unsigned)n, kNT, smem, st>>>(a);
  });
  return check_launch();
}
