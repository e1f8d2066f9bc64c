// Thin PyTorch binding above the C ABI (include/ddh.h) for the batch-1 latency path: one C++ call
// validates the tensors, allocates the four outputs (one caching-allocator block), fetches the
// current stream and calls ddh_forward through the function pointer the Python side took from the
// ctypes handle of _ddh.so.  The arithmetic is entirely behind the C ABI; this file only removes
// ~10 us of Python per call (tensor views, ctypes argument marshalling).  Anything it does not
// recognise (host tensors, other dtypes / devices, non-contiguous inputs) returns None and the
// general Python path (trajectory_head.py) handles it, including the error messages.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAFunctions.h>

#include <cstdint>

namespace {

typedef int (*ForwardFn)(void*, const float*, const float*, const void*, int, int, const float*, float*,
                         float*, float*, int64_t*, int, void*);

inline bool plain_f32(const at::Tensor& t, int dev) {
  return t.is_cuda() && t.get_device() == dev && t.scalar_type() == at::kFloat && t.is_contiguous();
}

// bev_layout: 0 NCHW, 1 NHWC (ddh.h).  Returns a dict of outputs, an int status on failure of the C
// call, or None when the fast path does not apply.
py::object forward_fast(std::uintptr_t fn, std::uintptr_t handle, const at::Tensor& ego, const at::Tensor& agents,
                        const at::Tensor& bev, const py::object& noise_obj, int bev_layout, int A, int P, int Na,
                        int C, int H, int W, int dev) {
  if (c10::cuda::current_device() != dev) return py::none();
  if (!plain_f32(ego, dev) || !plain_f32(agents, dev) || ego.dim() < 1) return py::none();
  const int64_t B = ego.size(0);
  if (agents.dim() != 3 || agents.size(0) != B || agents.size(1) != Na || agents.size(2) != 256 ||
      ego.numel() != B * 256)
    return py::none();
  if (!bev.is_cuda() || bev.get_device() != dev || !bev.is_contiguous() || bev.dim() != 4 || bev.size(0) != B)
    return py::none();
  const auto bt = bev.scalar_type();
  if (bt != at::kFloat && bt != at::kBFloat16) return py::none();
  if (bev_layout == 0) {
    if (bev.size(1) != C || bev.size(2) != H || bev.size(3) != W) return py::none();
  } else {
    if (bev.size(1) != H || bev.size(2) != W || bev.size(3) != C) return py::none();
  }
  at::Tensor noise;
  if (noise_obj.is_none()) {
    noise = at::randn({B, A, P, 2}, ego.options());   // transfuser_model_v2.py:593
  } else {
    noise = noise_obj.cast<at::Tensor>();
    if (!plain_f32(noise, dev) || noise.numel() != B * A * P * 2) return py::none();
  }
  const int64_t n_t = B * P * 3, n_m = B * A * P * 3, n_s = B * A;
  const int64_t off_i = (n_t + n_m + n_s + 1) & ~int64_t(1);   // int64 mode_idx: 8-byte aligned tail
  const int64_t total = off_i + 2 * B;
  // Everything before the launch is on the latency path (the GPU idles until the kernel arrives), so
  // the output block of THIS call was allocated right after the previous launch (one spare per
  // thread, same size and device), and the views are made after the launch.  Every call still
  // returns fresh tensors: a spare is handed out once.
  static thread_local at::Tensor spare;
  static thread_local cudaStream_t spare_stream = nullptr;   // (the caching allocator ties a block to its stream)
  cudaStream_t st = c10::cuda::getCurrentCUDAStream(dev).stream();
  at::Tensor flat;
  if (spare.defined() && spare.numel() == total && spare.get_device() == dev && spare_stream == st) {
    flat = std::move(spare);
    spare = at::Tensor();
  } else {
    flat = at::empty({total}, ego.options());
  }
  float* base = flat.data_ptr<float>();
  const int rc = reinterpret_cast<ForwardFn>(fn)(
      reinterpret_cast<void*>(handle), ego.data_ptr<float>(), agents.data_ptr<float>(), bev.data_ptr(),
      bt == at::kBFloat16 ? 1 : 0, bev_layout, noise.data_ptr<float>(), base, base + n_t, base + n_t + n_m,
      reinterpret_cast<int64_t*>(base + off_i), (int)B, reinterpret_cast<void*>(st));
  if (rc != 0) return py::int_(rc);
  at::Tensor traj = flat.as_strided({B, P, 3}, {P * 3, 3, 1}, 0);
  at::Tensor modes = flat.as_strided({B, A, P, 3}, {A * P * 3, P * 3, 3, 1}, n_t);
  at::Tensor scores = flat.as_strided({B, A}, {A, 1}, n_t + n_m);
  at::Tensor idx = flat.as_strided({2 * B}, {1}, off_i).view(at::kLong);
  if (B <= 64) {   // for the next call (small batches only)
    spare = at::empty({total}, ego.options());
    spare_stream = st;
  }
  py::dict out;
  out["trajectory"] = traj;
  out["trajectory_modes"] = modes;
  out["trajectory_scores"] = scores;
  out["mode_idx"] = idx;
  return std::move(out);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward_fast", &forward_fast, "ddh_forward with output allocation and stream lookup in C++");
}
