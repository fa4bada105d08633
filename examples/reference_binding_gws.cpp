// reference_binding_gws.cpp -- the binding a GeoT maintainer adds in THEIR tree, as a file
// that compiles: the body of csrc/gather_weight_scatter.cpp:22-34 (gather_weight_scatter_cuda_fwd_impl) after the swap
// from `gather_weight_scatter_cuda(...)` (csrc/cuda/header_cuda.h:12-17) to the C ABI of libgeot_b200.so.
// Compile check (tests/test_abi_host.py): g++ -std=c++17 -c with the torch include paths; link with -lgeot_b200.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "geot_b200.h"

at::Tensor gather_weight_scatter_cuda_fwd_impl(at::Tensor src_index, at::Tensor dst_index, at::Tensor weight, at::Tensor src) {
  c10::cuda::CUDAGuard guard(src.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  src = src.contiguous();
  weight = weight.contiguous();
  src_index = src_index.contiguous();
  dst_index = dst_index.contiguous();
  const int64_t E = dst_index.numel(), F = src.size(1);

  int64_t last = 0;                                      // replaces dst_index[-1].item()   (gather_weight_scatter.cpp:24)
  TORCH_CHECK(geot_b200_index_last(dst_index.data_ptr<int64_t>(), E, &last, stream) == GEOT_OK, "index is empty");
  const int64_t S = last + 1;
  auto out = at::empty({S, F}, src.options());           // replaces torch::zeros            (:28)
  const size_t ws_bytes = geot_b200_workspace_bytes(E, F, GEOT_F32, /*sorted=*/1);
  auto ws = at::empty({(int64_t)ws_bytes}, src.options().dtype(at::kByte));
  const int st = geot_b200_gather_weight_scatter(
      src.data_ptr(), src_index.data_ptr<int64_t>(), dst_index.data_ptr<int64_t>(), weight.data_ptr(), out.data_ptr(), E, S, F,
      GEOT_F32, GEOT_SUM, /*plan=*/nullptr, ws.data_ptr(), ws_bytes, stream);
  TORCH_CHECK(st == GEOT_OK, "gather_weight_scatter: ", geot_b200_status_string(st), " ", geot_b200_last_cuda_error());
  return out;
}
