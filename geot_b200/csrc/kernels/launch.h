// launch.h -- host-side interface between the C ABI (abi.cu) and the per-(dtype, reduce) kernel
// translation units (inst.cu compiled once per pair so the build parallelises).
#pragma once
#include <cuda_runtime.h>
#include "segment_reduce.cuh"

namespace geot {

struct Shape {
  int vecw;       // elements per lane vector (16 bytes worth, or 1 for the unaligned fallback)
  int lpr;        // lanes per row (group size): power of two <= 32
  int vpl;        // vectors per lane: 1, 2 or 4 (only with lpr == 32)
  int col_tiles;  // ceil(W / (lpr*vpl*vecw)); grid.x = n_tiles * col_tiles, column tile is the slow index
  int wm;         // WM_NONE / WM_EDGE / WM_GENERIC (sum kernels; the other reduce ops are built WM_GENERIC only)
  int pf;         // 0: gathered rows go straight to registers; > 0: cp.async shared-memory ring, pf sub-batches in flight
};

// ev0 / ev1 (optional): recorded on `stream` right before / after the main kernel (profiling hook)
typedef cudaError_t (*launch_fn)(const Params &, const Shape &, cudaStream_t, cudaEvent_t, cudaEvent_t);

// defined in inst.cu, one per (dtype, reduce op); index [dtype][red] with red in {sum,max,min,prod}
#define GEOT_DECL(TN, R) cudaError_t launch_##TN##_##R(const Params &, const Shape &, cudaStream_t, cudaEvent_t, cudaEvent_t);
GEOT_DECL(f32, 0) GEOT_DECL(f32, 2) GEOT_DECL(f32, 3) GEOT_DECL(f32, 4)
GEOT_DECL(f64, 0) GEOT_DECL(f64, 2) GEOT_DECL(f64, 3) GEOT_DECL(f64, 4)
GEOT_DECL(bf16, 0) GEOT_DECL(bf16, 2) GEOT_DECL(bf16, 3) GEOT_DECL(bf16, 4)
GEOT_DECL(f16, 0) GEOT_DECL(f16, 2) GEOT_DECL(f16, 3) GEOT_DECL(f16, 4)
#undef GEOT_DECL

}  // namespace geot
