// Shared declarations of the heatmap kernels (v1 CTA-per-plane fallback and the persistent
// warp-per-plane kernel): argument block, DARK blur helpers.
#pragma once
#include <math_constants.h>

#include <type_traits>

#include "lhn_common.cuh"
#include "lhn_exchange.cuh"

namespace lhn {

constexpr int kMaxWarps = 8;
constexpr int kMaxTeamsPerCta = 12;   // persistent team kernel: teams per CTA (sizes the one-launch workspace)

struct HmArgs {
  const void* hm;
  const void* hm_flip;
  const int32_t* flip_index;
  int64_t stride_b, stride_c, fstride_b, fstride_c;
  int64_t n_planes;
  int C, K, H, W, HW;
  int use_tma;
  // decode
  const float* center;
  const float* scale;
  int mask_mode, refine, transform, use_udp, ksize;
  float scale_x, scale_y;
  float tapsf[LHN_MAX_TAPS];
  double tapsd[LHN_MAX_TAPS];
  float* out_hm;
  float* out_kpts;
  int32_t* out_idx;
  // render + loss
  int loss_mode, unbiased;
  double feat_x, feat_y;
  float pos_value;
  float sigma[LHN_MAX_STACKS];
  const float* joints;
  int joints_stride;
  const float* vis;
  int vis_stride;
  float* out_weight;
  double* partials;
  // fused metrics
  const float* gt;
  const uint8_t* mask;
  const float* bbox_wh;
  float pck_thr, auc_nor;
  int auc_steps;
  float auc_thr[64];               // (float)(t / auc_steps) in double, t < auc_steps (top_down_eval.py:190-193)
  int64_t* counters;
  // persistent warp-per-plane kernel
  double inv2s2[LHN_MAX_STACKS];   // 1 / (2 sigma^2), host-computed in double
  float pos_radius[LHN_MAX_STACKS];// radius (px) inside which target > pos_value can hold
  int warp_smem;                   // bytes of shared memory owned by one team (stage + aux)
  int team_warps;                  // warps per team
  int stage_bytes;                 // bytes of one TMA stage (plane0 [+ plane1])
  int stagger_ns;                  // delay between the first loads of consecutive teams of a CTA (0 = none)
  int sweeper_tables;              // the sweepers (not the epilogue warp) evaluate each plane's Gaussian tables
  int stages;                      // stages per team (1, 2 or 4), at the start of the team's shared memory
  int tile_dim;                    // DARK window side = blur_ksize + 4 (0 when DARK is off)
  int force_cta_kernel;            // testing: bypass the warp kernel (env LHN_FORCE_CTA_KERNEL=1)
  // one-launch loss (team kernel only; all optional): per-team f64 sums + a self-resetting ticket
  double* team_sums;               // [teams, 4] workspace
  unsigned int* ticket;            // zero before the first launch; the kernel leaves it zero
  double* sums_out;                // [4] (S_pos, S_neg, N_pos, numel) of the whole launch
  float* loss_out;                 // [1] finalised loss (as lhn_loss_finalize)
  float loss_scale;
  int sum_reduction;
  double* fallback_partials;       // per-plane sums for the CTA-per-plane fallback of the one-launch step
  int loss_accumulate;             // LHN_FLAG_ACCUMULATE_LOSS
  int spare_sms;                   // LHN_FLAG_SPARE_SMS: SMs left free for concurrent kernels
  int overlap_previous;            // LHN_FLAG_OVERLAP_PREVIOUS: launch with programmatic stream serialization
  int feat_pow2;                   // feat_x, feat_y are powers of two: joint / feat == joint * inv_feat exactly
  double inv_feat_x, inv_feat_y;
  // in-kernel cross-GPU exchange (lhn_exchange; team kernel only)
  XchCtx xch;                      // xch.world == 0: no exchange
  unsigned int xch_seq;            // this launch's step number
  long long* xch_totals;           // fused metrics: running totals that receive += sum over ranks of a step's block
  unsigned long long* xch_prev_block;   // fused metrics: the PREVIOUS launch's per-step block (exchanged by this launch)
  unsigned int xch_prev_seq;            // ... and its step number
  unsigned long long* xch_prev2_block;  // the block of two launches back (published by the previous launch, consumed by this one)
  unsigned int xch_prev2_seq;
  int trigger_halfway;             // overlapped launch: let the successor grid in half-way (else after plane `trigger_plane`)
  int trigger_plane;               // default 2 (the third plane)
  int xch_courier;                 // 1: the grid's last CTA carries no planes and runs the exchange (set by the launcher)
};

// ---- shared-memory layout --------------------------------------------------------------------
struct SmemHeader {
  uint64_t bar;                // TMA completion barrier
  uint64_t pad;
  uint32_t red_key[kMaxWarps];
  uint32_t red_idx[kMaxWarps];
  double red_s[kMaxWarps];
  double red_b[kMaxWarps];     // slow-path blur max partials (as double; NaN propagates)
  // final per-plane values published by warp 0
  int px, py;
  uint32_t idx;
  float maxval;
  int need_slow;
  int nan_found;
  float bmax;
  float pad2;
  double hbuf[(5 + LHN_MAX_TAPS - 1) * 5];  // row-pass values around the peak (f32 or f64)
  float hout[32];                           // 5x5 blurred values around the peak
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
__host__ __device__ inline size_t smem_bytes_for(int H, int W, bool flip) {
  size_t off = align_up(sizeof(SmemHeader), 16);
  off += align_up((size_t)(W + H) * sizeof(float), 16);          // ex, ey tables
  off = align_up(off, 128);
  if (!std::is_same<T, float>::value) off += align_up((size_t)H * W * sizeof(float), 128);
  off += align_up((size_t)H * W * sizeof(T), 128);
  if (flip) off += align_up((size_t)H * W * sizeof(T), 128);
  return off;
}

// ---- DARK helpers ------------------------------------------------------------------------------
template <typename A> __device__ __forceinline__ A fma_rn(A a, A b, A c);
template <> __device__ __forceinline__ float fma_rn<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fma_rn<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <typename A> __device__ __forceinline__ A mul_rn(A a, A b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename A> __device__ __forceinline__ A add_rn(A a, A b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// Row pass of the zero-padded separable blur at (y, x): sequential FMA over taps 0..k-1
// (the order cv2's generic RowFilter uses for CV_32F / CV_64F).
// MIRROR: the decoded plane is stored reversed along W (see the flip-average note in the kernel).
template <typename A, bool MIRROR>
__device__ __forceinline__ A blur_row(const float* plane, int H, int W, int y, int x, int ksize,
                                      const A* taps) {
  A acc = 0;
  if (y < 0 || y >= H) return acc;
  const float* row = plane + (size_t)y * W;
  const int b = (ksize - 1) >> 1;
  for (int j = 0; j < ksize; ++j) {
    int xx = x + j - b;
    A v = (xx >= 0 && xx < W) ? (A)row[MIRROR ? (W - 1 - xx) : xx] : (A)0;
    acc = fma_rn<A>(taps[j], v, acc);
  }
  return acc;
}

// Blurred values on the 5x5 neighbourhood of (px, py), computed by one warp.
template <typename A, bool MIRROR>
__device__ void dark_window(const float* plane, int H, int W, int px, int py, int ksize,
                            const A* taps, A* hbuf, float* hout, int lane) {
  const int b = (ksize - 1) >> 1;
  const int nrows = 5 + 2 * b;
  for (int e = lane; e < nrows * 5; e += kWarp) {
    int r = e / 5, c = e - r * 5;
    hbuf[e] = blur_row<A, MIRROR>(plane, H, W, py - 2 - b + r, px - 2 + c, ksize, taps);
  }
  __syncwarp();
  if (lane < 25) {
    int dr = lane / 5, c = lane - dr * 5;
    A acc = mul_rn<A>(taps[b], hbuf[(dr + b) * 5 + c]);
    for (int j = 1; j <= b; ++j)
      acc = fma_rn<A>(taps[b + j], add_rn<A>(hbuf[(dr + b + j) * 5 + c], hbuf[(dr + b - j) * 5 + c]), acc);
    hout[lane] = (float)acc;
  }
  __syncwarp();
}

// Exact blurred value at one pixel (slow path: full-plane max of the blurred map).
template <typename A, bool MIRROR>
__device__ float blur_at(const float* plane, int H, int W, int x, int y, int ksize, const A* taps) {
  const int b = (ksize - 1) >> 1;
  A acc = mul_rn<A>(taps[b], blur_row<A, MIRROR>(plane, H, W, y, x, ksize, taps));
  for (int j = 1; j <= b; ++j)
    acc = fma_rn<A>(taps[b + j],
                    add_rn<A>(blur_row<A, MIRROR>(plane, H, W, y + j, x, ksize, taps),
                              blur_row<A, MIRROR>(plane, H, W, y - j, x, ksize, taps)), acc);
  return (float)acc;
}

__device__ __forceinline__ float nanmax(float a, float b) {  // np.max: NaN propagates
  return (a != a || a > b) ? a : ((b != b) ? b : (a >= b ? a : b));
}


// persistent team kernel (lhn_heatmap_team.cuh, launched from lhn_heatmap_team_launch.cu): returns LHN_OK after launching, or
// +1 when the shape/alignment is outside its envelope (caller falls back to the CTA-per-plane kernel)
int launch_heatmap_warp_kernel(HmArgs& a, int dtype, cudaStream_t st);

}  // namespace lhn
