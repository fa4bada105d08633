// inst.cu -- instantiates segment_reduce_kernel / segment_fixup_kernel for ONE (dtype, reduce op)
// pair, selected with -DGEOT_T=<type> -DGEOT_TN=<name> -DGEOT_RED=<0|2|3|4>.  mean shares the sum
// kernels (Params::mean).
#include "launch.h"

#ifndef GEOT_T
#error "compile with -DGEOT_T=... -DGEOT_TN=... -DGEOT_RED=..."
#endif

namespace geot {
namespace {

template <typename T, int VECW, int LPR, int VPL, int RED, int WM>
cudaError_t launch_wm(const Params &p, const Shape &sh, cudaStream_t stream) {
  using A = typename AccOf<T>::type;
  constexpr int NG = kThreads / LPR;
  constexpr int CW = LPR * VPL * VECW;
  constexpr size_t smem = 2 * (size_t)NG * CW * sizeof(A) + (size_t)NG * (4 * 8 + 4);
  auto kern = segment_reduce_kernel<T, VECW, LPR, VPL, RED, WM>;
  if (smem > 48 * 1024) {
    static bool configured = false;   // per instantiation
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured = true;
    }
  }
  const long long blocks = (long long)p.n_tiles * sh.col_tiles;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// sum kernels exist per weight mode; max / min / prod are built for the generic weight mode only
template <typename T, int VECW, int LPR, int VPL, int RED>
cudaError_t launch_one(const Params &p, const Shape &sh, cudaStream_t stream) {
  if constexpr (RED == RED_SUM) {
    if (sh.wm == WM_NONE) return launch_wm<T, VECW, LPR, VPL, RED, WM_NONE>(p, sh, stream);
    if (sh.wm == WM_EDGE) return launch_wm<T, VECW, LPR, VPL, RED, WM_EDGE>(p, sh, stream);
  }
  return launch_wm<T, VECW, LPR, VPL, RED, WM_GENERIC>(p, sh, stream);
}

template <typename T, int VECW, int RED>
cudaError_t launch_shape(const Params &p, const Shape &sh, cudaStream_t stream) {
  if (sh.vpl == 1) {
    switch (sh.lpr) {
      case 1: return launch_one<T, VECW, 1, 1, RED>(p, sh, stream);
      case 2: return launch_one<T, VECW, 2, 1, RED>(p, sh, stream);
      case 4: return launch_one<T, VECW, 4, 1, RED>(p, sh, stream);
      case 8: return launch_one<T, VECW, 8, 1, RED>(p, sh, stream);
      case 16: return launch_one<T, VECW, 16, 1, RED>(p, sh, stream);
      case 32: return launch_one<T, VECW, 32, 1, RED>(p, sh, stream);
    }
  } else if (sh.lpr == 32) {
    if (sh.vpl == 2) return launch_one<T, VECW, 32, 2, RED>(p, sh, stream);
    if constexpr (VECW <= 4) {   // 8-element vectors stop at 2 per lane (accumulator registers)
      if (sh.vpl == 4) return launch_one<T, VECW, 32, 4, RED>(p, sh, stream);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace

#define GEOT_CAT_(a, b, c) a##b##_##c
#define GEOT_CAT(a, b, c) GEOT_CAT_(a, b, c)

cudaError_t GEOT_CAT(launch_, GEOT_TN, GEOT_RED)(const Params &p, const Shape &sh, cudaStream_t stream,
                                                  cudaEvent_t ev0, cudaEvent_t ev1) {
  using T = GEOT_T;
  constexpr int FULL = 16 / (int)sizeof(T);
  cudaError_t e;
  if (ev0) cudaEventRecord(ev0, stream);
  if (sh.vecw == FULL) e = launch_shape<T, FULL, GEOT_RED>(p, sh, stream);
  else if (sh.vecw == 1) e = launch_shape<T, 1, GEOT_RED>(p, sh, stream);
  else return cudaErrorInvalidValue;
  if (ev1) cudaEventRecord(ev1, stream);
  if (e != cudaSuccess) return e;
  // second pass: segments cut by tile boundaries
  const unsigned blocks = (unsigned)((p.n_tiles + (kThreads / 32) - 1) / (kThreads / 32));
  segment_fixup_kernel<T, GEOT_RED><<<blocks, kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace geot
