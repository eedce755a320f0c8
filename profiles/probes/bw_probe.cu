// Read-bandwidth calibration for the B200 roofline discussion in DESIGN.md (not part of liblhn.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu && ./bw_probe
// A: plain LDG.128 grid-stride read+sum.  B: one warp per CTA streaming cp.async.bulk (TMA 1-D) chunks
// through an N-stage shared-memory ring, touching one word per chunk.  C: B with every byte read by
// LDS.128 from a team of warps (the access structure of the fused heatmap kernel, no math).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_1d(void* dst, const void* src, uint32_t bytes, uint64_t* b, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t pol_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }

template <int U>
__global__ void read_ldg(const float4* __restrict__ src, size_t n4, float* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

// B: one producer/consumer thread per CTA
__global__ void read_tma(const char* __restrict__ src, size_t bytes, int chunk, int stages, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  unsigned char* ring = sm + 1024;
  if (threadIdx.x != 0) return;
  for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint64_t pol = pol_evict_first();
  const size_t nchunks = bytes / chunk;
  size_t c = blockIdx.x;
  size_t issued = c;
  for (int s = 0; s < stages && issued < nchunks; ++s, issued += gridDim.x) {
    mbar_expect(bars + s, chunk);
    tma_1d(ring + (size_t)s * chunk, src + issued * chunk, chunk, bars + s, pol);
  }
  float acc = 0.f;
  int s = 0; uint32_t ph = 0;
  for (; c < nchunks; c += gridDim.x) {
    mbar_wait(bars + s, ph);
    acc += *reinterpret_cast<volatile float*>(ring + (size_t)s * chunk);
    if (issued < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(bars + s, chunk);
      tma_1d(ring + (size_t)s * chunk, src + issued * chunk, chunk, bars + s, pol);
      issued += gridDim.x;
    }
    if (++s == stages) { s = 0; ph ^= 1; }
  }
  if (acc == 123.456f) out[0] = acc;
}

// C: teams of TW warps, one stage per team, every byte read with LDS.128 (sum), stage re-armed after the sweep
__global__ void read_tma_teams(const char* __restrict__ src, size_t bytes, int chunk, int tw, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / tw, wt = warp % tw, tl = wt * 32 + lane, TT = tw * 32;
  const int nteams = (blockDim.x >> 5) / tw;
  unsigned char* stage = sm + (size_t)team * (chunk + 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + chunk);
  const size_t nchunks = bytes / chunk;
  const size_t total = (size_t)gridDim.x * nteams;
  size_t c = (size_t)blockIdx.x * nteams + team;
  uint64_t pol = 0;
  if (tl == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pol = pol_evict_first();
    if (c < nchunks) { mbar_expect(bar, chunk); tma_1d(stage, src + c * chunk, chunk, bar, pol); }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TT) : "memory");
  float acc = 0.f; uint32_t ph = 0;
  for (; c < nchunks; c += total) {
    mbar_wait(bar, ph); ph ^= 1;
    const float4* p = reinterpret_cast<const float4*>(stage);
    for (int q = tl; q < chunk / 16; q += TT) { float4 v = p[q]; acc += v.x + v.y + v.z + v.w; }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TT) : "memory");
    if (tl == 0 && c + total < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(bar, chunk); tma_1d(stage, src + (c + total) * chunk, chunk, bar, pol);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}


// D: the fused heatmap kernel's exact traffic shape: each team stage = 16 KB from tensor A + 16 KB from tensor B
// (plane + flipped plane), 21504 plane pairs (705 MB), teams walk planes g, g+888, ...
__global__ void read_tma_teams2(const char* __restrict__ srcA, const char* __restrict__ srcB, size_t nplanes, int chunk, int tw, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / tw, wt = warp % tw, tl = wt * 32 + lane, TT = tw * 32;
  const int nteams = (blockDim.x >> 5) / tw;
  unsigned char* stage = sm + (size_t)team * (2 * chunk + 128);
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + 2 * chunk);
  const size_t total = (size_t)gridDim.x * nteams;
  size_t c = (size_t)blockIdx.x * nteams + team;
  uint64_t pol = 0;
  if (tl == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pol = pol_evict_first();
    if (c < nplanes) { mbar_expect(bar, 2 * chunk); tma_1d(stage, srcA + c * chunk, chunk, bar, pol); tma_1d(stage + chunk, srcB + c * chunk, chunk, bar, pol); }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TT) : "memory");
  float acc = 0.f; uint32_t ph = 0;
  for (; c < nplanes; c += total) {
    mbar_wait(bar, ph); ph ^= 1;
    const float4* p = reinterpret_cast<const float4*>(stage);
    for (int q = tl; q < 2 * chunk / 16; q += TT) { float4 v = p[q]; acc += v.x + v.y + v.z + v.w; }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TT) : "memory");
    if (tl == 0 && c + total < nplanes) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(bar, 2 * chunk); tma_1d(stage, srcA + (c + total) * chunk, chunk, bar, pol); tma_1d(stage + chunk, srcB + (c + total) * chunk, chunk, bar, pol);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static float best_ms(F f, int reps = 8) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (i > 1 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

__global__ void copy_k(const float4* __restrict__ s, float4* __restrict__ d, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) d[i] = s[i];
}

int main() {
  const size_t bytes = (size_t)1408 << 20;   // 1.4 GiB
  char *src, *dst; float* out;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(src, 1, bytes)); CK(cudaMemset(dst, 0, bytes));
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  { float ms = best_ms([&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D          : %8.1f GB/s (read+write)\n", 2.0 * bytes / ms / 1e6); }
  { float ms = best_ms([&] { copy_k<<<sms * 8, 512>>>((const float4*)src, (float4*)dst, bytes / 16); });
    printf("copy kernel             : %8.1f GB/s (read+write)\n", 2.0 * bytes / ms / 1e6); }
  for (int ctas : {2, 4, 8}) for (int thr : {256, 512}) {
    float ms = best_ms([&] { read_ldg<8><<<sms * ctas, thr>>>((const float4*)src, bytes / 16, out); });
    printf("LDG.128 U8 %d CTA/SM x%4d : %8.1f GB/s\n", ctas, thr, bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(read_tma_teams, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  for (int chunk : {4096, 8192, 16384, 32768}) for (int stages : {2, 4, 6, 12, 24, 48}) {
    if ((size_t)chunk * stages > 200 * 1024) continue;
    float ms = best_ms([&] { read_tma<<<sms, 32, 1024 + chunk * stages>>>(src, bytes, chunk, stages, out); });
    printf("TMA ring chunk %5d x %2d stages (%3d KB/SM) : %8.1f GB/s\n", chunk, stages, chunk * stages / 1024, bytes / ms / 1e6);
  }
  for (int chunk : {16384, 32768}) for (int tw : {2, 4}) {
    int nteams = (200 * 1024) / (chunk + 128); if (nteams * tw > 32) nteams = 32 / tw;
    float ms = best_ms([&] { read_tma_teams<<<sms, nteams * tw * 32, nteams * (chunk + 128)>>>(src, bytes, chunk, tw, out); });
    printf("TMA teams chunk %5d, %2d teams x %d warps : %8.1f GB/s\n", chunk, nteams, tw, bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(read_tma_teams2, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  {
    const size_t np = 21504; const int chunk = 16384;
    for (int tw : {2, 4}) {
      int nteams = 6;
      float ms = best_ms([&] { read_tma_teams2<<<sms, nteams * tw * 32, nteams * (2 * chunk + 128)>>>(src, src + (np * chunk), np, chunk, tw, out); }, 12);
      printf("K1 shape: 2 x 16 KB per stage, 21504 plane pairs (705 MB), 6 teams x %d warps : %8.1f GB/s (%.1f us)\n", tw, 2.0 * np * chunk / ms / 1e6, ms * 1e3);
    }
    // same bytes as one contiguous tensor with 32 KB chunks
    float ms = best_ms([&] { read_tma_teams<<<sms, 6 * 4 * 32, 6 * (32768 + 128)>>>(src, np * 2 * chunk, 32768, 4, out); }, 12);
    printf("same 705 MB, one tensor, 32 KB chunks, 6 teams x 4 warps : %8.1f GB/s (%.1f us)\n", 2.0 * np * chunk / ms / 1e6, ms * 1e3);
    ms = best_ms([&] { read_ldg<8><<<sms * 4, 512>>>((const float4*)src, np * 2 * chunk / 16, out); }, 12);
    printf("same 705 MB, LDG.128 U8 4 CTA/SM x 512 : %8.1f GB/s (%.1f us)\n", 2.0 * np * chunk / ms / 1e6, ms * 1e3);
  }
  return 0;
}
