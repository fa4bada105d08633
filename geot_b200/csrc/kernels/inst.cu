// inst.cu -- instantiates segment_reduce_kernel / segment_fixup_kernel for ONE (dtype, reduce op)
// pair, selected with -DGEOT_T=<type> -DGEOT_TN=<name> -DGEOT_RED=<0|2|3|4>.  mean shares the sum
// kernels (Params::mean).
#include "launch.h"

#ifndef GEOT_T
#error "compile with -DGEOT_T=... -DGEOT_TN=... -DGEOT_RED=..."
#endif

namespace geot {
namespace {

template <typename T, int VECW, int LPR, int VPL, int RED, int WM, int PF>
cudaError_t launch_pf(const Params &p, const Shape &sh, cudaStream_t stream) {
  constexpr size_t smem = ShapeOf<T, VECW, LPR, VPL, PF>::smem_bytes;
  auto kern = segment_reduce_kernel<T, VECW, LPR, VPL, RED, WM, PF>;
  if (smem > 48 * 1024) {
    // the attribute is per (function, device): one flag per instantiation and device ordinal.  A racing second
    // thread at worst sets the same value again.
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  const long long blocks = (long long)p.n_tiles * sh.col_tiles;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// Ring variants built.  Production: the lean ring (kLeanFlag + depth; depths GEOT_LEAN_A / _B, clamped to what the
// batch allows) for the sum kernels with fp32 accumulators and at most one weight per edge, and the first-generation
// ring (GEOT_PF_A / _B = depths 2 and 3) for what the lean ring does not serve (fp64, per-head weights) and as the
// A/B reference (GEOT_B200_RING=2).  Tuning builds override the lists.
#ifndef GEOT_PF_A
#define GEOT_PF_A 2
#define GEOT_PF_B 3
#endif
#ifndef GEOT_LEAN_A
#define GEOT_LEAN_A 3
#define GEOT_LEAN_B 7
#endif
// largest lean depth d <= want with (d + 1) | SB
constexpr int lean_depth(int want, int sb) {
  int d = want < sb ? want : sb - 1;
  while (d > 0 && sb % (d + 1) != 0) --d;
  return d;
}
template <typename T, int VECW, int LPR, int VPL, int RED, int WM>
cudaError_t launch_wm(const Params &p, const Shape &sh, cudaStream_t stream) {
  if constexpr (RED == RED_SUM && VECW * sizeof(T) == 16 && LPR >= 8) {
    int pf = sh.pf;
    if (pf & kLeanFlag) {
      if constexpr (WM != WM_GENERIC && sizeof(typename AccOf<T>::type) == 4) {
        if (p.chunk_edges % LPR == 0) {
          constexpr int SB = ShapeOf<T, VECW, LPR, VPL, kLeanFlag | 1>::SB;
          constexpr int DA = lean_depth(GEOT_LEAN_A, SB), DB = lean_depth(GEOT_LEAN_B, SB);
          static_assert(DA >= 1 && DB >= 1, "a batch has at least two sub-batches");
          const int want = pf & kDepthMask;
          // the instantiation with the segment_reduce_ex options only when the call uses one of them
          const bool ext = p.accumulate || p.zero_gaps || p.edge_perm != nullptr || p.mean_rowptr != nullptr;
          if (want > DA && DB != DA && ShapeOf<T, VECW, LPR, VPL, kLeanFlag | DB>::max_blocks >= 1)
            return ext ? launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | kExtFlag | DB>(p, sh, stream)
                       : launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | DB>(p, sh, stream);
          if (ShapeOf<T, VECW, LPR, VPL, kLeanFlag | DA>::max_blocks >= 1)
            return ext ? launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | kExtFlag | DA>(p, sh, stream)
                       : launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | DA>(p, sh, stream);
        }
      }
      if constexpr (WM == WM_GENERIC && sizeof(typename AccOf<T>::type) == 4) {
        // per-head weights (mh_spmm): the lean ring with the weights parked [head][edge]; up to kHeadMax heads
        if (p.chunk_edges % LPR == 0 && p.weight != nullptr && p.W / p.F <= kHeadMax && p.W % p.F == 0 && p.edge_perm == nullptr) {
          constexpr int SB = ShapeOf<T, VECW, LPR, VPL, kLeanFlag | kHeadFlag | 1>::SB;
          constexpr int DA = lean_depth(GEOT_LEAN_A, SB);
          const bool ext = p.accumulate || p.zero_gaps || p.mean_rowptr != nullptr;
          if (ShapeOf<T, VECW, LPR, VPL, kLeanFlag | kHeadFlag | DA>::max_blocks >= 1)
            return ext ? launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | kHeadFlag | kExtFlag | DA>(p, sh, stream)
                       : launch_pf<T, VECW, LPR, VPL, RED, WM, kLeanFlag | kHeadFlag | DA>(p, sh, stream);
        }
      }
      pf = (VPL == 1) ? GEOT_PF_B : GEOT_PF_A;    // not served by the lean ring: first-generation ring (depth 3; 2 for rows >= 1 KB)
    }
    constexpr int U = ShapeOf<T, VECW, LPR, VPL, 1>::U;   // ring sub-batch
    constexpr int PFMAX = LPR / U;
    // a list entry is a depth, clamped to what the batch allows
#define GEOT_CLAMP_PF(X) (((X) & kDepthMask) < PFMAX ? ((X) & kDepthMask) : PFMAX)
    constexpr int PFA = GEOT_CLAMP_PF(GEOT_PF_A);
    if (pf == GEOT_PF_A && ShapeOf<T, VECW, LPR, VPL, PFA>::max_blocks >= 1)
      return launch_pf<T, VECW, LPR, VPL, RED, WM, PFA>(p, sh, stream);
#ifdef GEOT_PF_B
    constexpr int PFB = GEOT_CLAMP_PF(GEOT_PF_B);
    if (pf == GEOT_PF_B && ShapeOf<T, VECW, LPR, VPL, PFB>::max_blocks >= 1)
      return launch_pf<T, VECW, LPR, VPL, RED, WM, PFB>(p, sh, stream);
#endif
#ifdef GEOT_PF_C
    constexpr int PFC = GEOT_CLAMP_PF(GEOT_PF_C);
    if (pf == GEOT_PF_C && ShapeOf<T, VECW, LPR, VPL, PFC>::max_blocks >= 1)
      return launch_pf<T, VECW, LPR, VPL, RED, WM, PFC>(p, sh, stream);
#endif
  }
  return launch_pf<T, VECW, LPR, VPL, RED, WM, 0>(p, sh, stream);
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
  // second pass: segments cut by tile boundaries.  Programmatic dependent launch: the fixup grid is set up while the
  // main kernel is still running and its CTAs start at griddepcontrol.wait, which returns when the main grid has
  // completed and its writes are visible -- the launch gap between the two kernels (a third of the step on the small
  // BASELINE shapes) overlaps the main kernel's tail.
  const unsigned blocks = (unsigned)((p.n_tiles + (kThreads / 32) - 1) / (kThreads / 32));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, segment_fixup_kernel<T, GEOT_RED>, p);
}

}  // namespace geot
