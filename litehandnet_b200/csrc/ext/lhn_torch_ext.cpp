// Torch extension shim over the C ABI (include/lhn.h): checks the tensors, allocates the outputs, takes torch's current
// CUDA stream and hands RAW DEVICE POINTERS to liblhn.so.  No arithmetic lives here.  It exists because the eager
// Python/ctypes wrapper costs 20-30 us of host time per call, which is more than the kernel at the reference's batch
// sizes (test.py:114-126 decodes 32-64 samples per call); this path costs a few microseconds.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../../include/lhn.h"

namespace {

int dtype_code(const at::Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return LHN_F32;
    case at::kBFloat16: return LHN_BF16;
    case at::kHalf: return LHN_F16;
    default: TORCH_CHECK(false, "lhn: unsupported dtype ", t.scalar_type(), " (f32, bf16 or f16)");
  }
}

// [B,C,H,W] with contiguous planes -> strides in elements
void plane_strides(const at::Tensor& hm, const char* name, int64_t& sb, int64_t& sc) {
  TORCH_CHECK(hm.is_cuda(), "lhn: ", name, " must be a CUDA tensor (this path has no CPU fallback)");
  TORCH_CHECK(hm.dim() == 4, "lhn: ", name, " must be [B,K,H,W]");
  TORCH_CHECK(hm.stride(3) == 1 && hm.stride(2) == hm.size(3), "lhn: ", name, " planes must be contiguous");
  sb = hm.stride(0);
  sc = hm.size(1) > 1 ? hm.stride(1) : hm.size(2) * hm.size(3);
}

const float* f32_ptr(const c10::optional<at::Tensor>& t, const at::Tensor& like, const char* name) {
  if (!t.has_value()) return nullptr;
  TORCH_CHECK(t->is_cuda() && t->device() == like.device(), "lhn: ", name, " must be on the heatmaps' device");
  TORCH_CHECK(t->scalar_type() == at::kFloat && t->is_contiguous(), "lhn: ", name, " must be contiguous f32");
  return t->data_ptr<float>();
}

void check_rc(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, ": error ", rc, rc == LHN_ECUDA ? std::string(" ") + lhn_last_cuda_error() : std::string());
}

}  // namespace

// lhn_decode_heatmap (decode only).  dp_addr: address of a (cached) lhn_decode_params built by the Python layer.
// Returns (hm_kpts [B,C,3], kpts [B,C,3], idx [B,C] int32 or an empty tensor).
std::vector<at::Tensor> decode_heatmap(const at::Tensor& hm, const c10::optional<at::Tensor>& hm_flip,
                                       const c10::optional<at::Tensor>& flip_index, const c10::optional<at::Tensor>& center,
                                       const c10::optional<at::Tensor>& scale, int64_t dp_addr, bool want_idx) {
  int64_t sb, sc, fb = 0, fc = 0;
  plane_strides(hm, "heatmaps", sb, sc);
  const void* flip = nullptr;
  if (hm_flip.has_value()) {
    TORCH_CHECK(hm_flip->sizes() == hm.sizes() && hm_flip->scalar_type() == hm.scalar_type() && hm_flip->device() == hm.device(),
                "lhn: flipped heatmaps must match heatmaps in shape, dtype and device");
    plane_strides(*hm_flip, "flipped heatmaps", fb, fc);
    flip = hm_flip->data_ptr();
  }
  const int32_t* fi = nullptr;
  if (flip_index.has_value()) {
    TORCH_CHECK(flip_index->is_cuda() && flip_index->scalar_type() == at::kInt && flip_index->is_contiguous() &&
                flip_index->numel() == hm.size(1), "lhn: flip_index must be a contiguous int32 CUDA tensor of K entries");
    fi = flip_index->data_ptr<int32_t>();
  }
  const c10::cuda::CUDAGuard guard(hm.device());
  const int64_t B = hm.size(0), C = hm.size(1);
  auto f32 = hm.options().dtype(at::kFloat);
  at::Tensor out_hm = at::empty({B, C, 3}, f32), out_k = at::empty({B, C, 3}, f32);
  at::Tensor out_idx = want_idx ? at::empty({B, C}, hm.options().dtype(at::kInt)) : at::Tensor();
  const int rc = lhn_decode_heatmap(hm.data_ptr(), flip, fi, dtype_code(hm), B, (int)C, (int)hm.size(2), (int)hm.size(3), sb, sc, fb, fc,
                                    f32_ptr(center, hm, "center"), f32_ptr(scale, hm, "scale"),
                                    reinterpret_cast<const lhn_decode_params*>(dp_addr), out_hm.data_ptr<float>(),
                                    out_k.data_ptr<float>(), want_idx ? out_idx.data_ptr<int32_t>() : nullptr, nullptr, nullptr,
                                    0, nullptr, 0, nullptr, nullptr, at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "lhn_decode_heatmap");
  return {out_hm, out_k, out_idx};
}

// lhn_decode_simdr_flags.  Returns out [B,K,3].
at::Tensor decode_simdr(const at::Tensor& xv, const at::Tensor& yv, int64_t k, const c10::optional<at::Tensor>& center,
                        const c10::optional<at::Tensor>& scale, int64_t flags) {
  TORCH_CHECK(xv.is_cuda() && yv.is_cuda() && xv.device() == yv.device(), "lhn: SimDR vectors must be CUDA tensors on one device");
  TORCH_CHECK(xv.dim() == 3 && yv.dim() == 3 && xv.size(0) == yv.size(0) && xv.size(1) == yv.size(1) && xv.is_contiguous() &&
              yv.is_contiguous() && xv.scalar_type() == yv.scalar_type(), "lhn: x/y vectors must be contiguous [B,K,L] of one dtype");
  const c10::cuda::CUDAGuard guard(xv.device());
  at::Tensor out = at::empty({xv.size(0), xv.size(1), 3}, xv.options().dtype(at::kFloat));
  const int rc = lhn_decode_simdr_flags(xv.data_ptr(), yv.data_ptr(), dtype_code(xv), xv.size(0), (int)xv.size(1), (int)xv.size(2),
                                        (int)yv.size(2), (int)k, f32_ptr(center, xv, "center"), f32_ptr(scale, xv, "scale"), 0,
                                        nullptr, out.data_ptr<float>(), nullptr, (int)flags,
                                        at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "lhn_decode_simdr");
  return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("decode_heatmap", &decode_heatmap, "lhn_decode_heatmap (decode only) on torch's current stream");
  m.def("decode_simdr", &decode_simdr, "lhn_decode_simdr_flags on torch's current stream");
}
