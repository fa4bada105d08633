// sddmm.cuh -- sampled dense-dense product on a COO edge list, sm_100a:
//
//     out[e] = < mat1[row_index[e], :] , mat2[col_index[e], :] >          e = 0 .. E-1
//
// Replaces the reference's sddmm_coo_ebalance_{vec4,vec2,scalar} (csrc/cuda/sddmm_coo_kernel.cuh:3-215,
// launcher csrc/cuda/gather_weight_scatter_cuda.cu:41-62), which serve the weight gradient of
// gather_weight_scatter (geot/gather_weight_scatter.py:47).  The reference is fp32 + int32 only and
// launches nnz/16 blocks of 32 threads; this one is built like segment_reduce_kernel:
//
//   * edge-count partition: a GROUP of LPR lanes owns `chunk_edges` consecutive edges and keeps whole
//     rows in registers (128-bit loads); a 256-thread CTA owns 256/LPR chunks;
//   * index streams are read once per edge, coalesced (lane l loads edge l of the batch), with an L2
//     evict-first policy; row offsets reach the other lanes by shuffle;
//   * row_index is the sorted one in the backward pass (dst-sorted edge list): the mat1 row stays in
//     registers while consecutive edges share it and is re-loaded only at a segment head -- half the
//     row traffic of the reference.  Sortedness is not required for correctness;
//   * indices are loaded two batches ahead; the gathered mat2 rows are staged through a per-group shared-memory
//     ring by cp.async (PF sub-batches of U rows in flight, as in segment_reduce_kernel), or -- narrow rows and
//     unaligned data -- double-buffered in registers;
//   * the U per-lane partial dot products are reduced across the group with a transposing butterfly
//     (U - 1 + log2(LPR/U) shuffles per U edges instead of U * log2(LPR)), after which lane k*(LPR/U)
//     holds edge k's result: U consecutive outputs = one 32-byte sector.
#pragma once
#include "segment_reduce.cuh"

namespace geot {

struct SddmmParams {
  const void *mat1;             // [N1, W], indexed by row_index
  const void *mat2;             // [N2, W], indexed by col_index
  const int64_t *row_index;     // [E]
  const int64_t *col_index;     // [E]
  void *out;                    // [E]
  int64_t E;
  int64_t W;
  int chunk_edges;
};

// Transposing butterfly over a group of LPR lanes: p[0..U) per lane in, and on return p[0] of lane gl
// holds the group-wide sum of value (gl / (LPR / U)).
template <int U, int LPR, typename A>
__device__ __forceinline__ void group_transpose_reduce(A (&p)[U], unsigned gmask, int gl) {
  static_assert(U <= LPR, "U values need at least U lanes");
  int n = U;
#pragma unroll
  for (int s = LPR / 2; s >= 1; s >>= 1) {
    if (n > 1) {
      const bool up = (gl & s) != 0;
      const int h = n / 2;
#pragma unroll
      for (int i = 0; i < U / 2; ++i) {
        if (i < h) {
          const A send = up ? p[i] : p[i + h];
          const A keep = up ? p[i + h] : p[i];
          p[i] = keep + __shfl_xor_sync(gmask, send, s, 32);
        }
      }
      n = h;
    } else {
      p[0] += __shfl_xor_sync(gmask, p[0], s, 32);
    }
  }
}

// PF == 0: gathered mat2 rows go global -> registers, double buffered.  PF > 0: they are staged through a per-group
// shared-memory ring by cp.async exactly as in segment_reduce_kernel (PF sub-batches of U rows in flight).
template <typename T, int VECW, int LPR, int VPL, int PF>
struct SddmmShape {
  static constexpr int NG = kThreads / LPR;
  static constexpr int CW = LPR * VPL * VECW;
  static constexpr int U0 = PF > 0 ? (VPL >= 4 ? 2 : 4) : ((VPL >= 4) ? 2 : (VPL == 2 ? 4 : 8));
  static constexpr int U = (LPR < U0) ? LPR : U0;
  static constexpr int NS = PF + 1;
  static constexpr size_t smem_bytes = PF > 0 ? (size_t)NG * NS * U * CW * sizeof(T) : 0;
  static constexpr int max_blocks = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int min_blocks = PF == 0 ? 2 : (max_blocks >= 3 ? 3 : (max_blocks >= 2 ? 2 : 1));
  static_assert(PF == 0 || PF * U <= LPR, "the prefetch distance must stay within one batch ahead");
};

template <typename T, int VECW, int LPR, int VPL, int PF>
__global__ void __launch_bounds__(kThreads, (SddmmShape<T, VECW, LPR, VPL, PF>::min_blocks)) sddmm_coo_kernel(const SddmmParams p) {
  using A = typename AccOf<T>::type;
  using VecT = Vec<T, VECW>;
  using SH = SddmmShape<T, VECW, LPR, VPL, PF>;
  constexpr int NG = SH::NG, CW = SH::CW, U = SH::U, NS = SH::NS;
  extern __shared__ __align__(16) unsigned char sddmm_smem[];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = tid / LPR;
  const int gl = tid % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const int gshift = lane & ~(LPR - 1);
  const int64_t E = p.E, W = p.W;
  const int C = p.chunk_edges;
  const int64_t e_begin = ((int64_t)blockIdx.x * NG + g) * (int64_t)C;
  const int64_t e_end = min(e_begin + (int64_t)C, E);
  if (e_begin >= e_end) return;     // whole groups leave together: no block-level barrier below

  const T *__restrict__ mat1 = static_cast<const T *>(p.mat1);
  const T *__restrict__ mat2 = static_cast<const T *>(p.mat2);
  T *__restrict__ out = static_cast<T *>(p.out);
  const uint64_t pol = policy_evict_first();
  const int64_t row_bytes = W * (int64_t)sizeof(T);

  // This lane's columns; lanes past the row end read column 0 and contribute nothing.
  bool col_ok[VPL];
  int64_t col_b[VPL];   // byte offset of this lane's vector j inside a row
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int64_t c = (int64_t)(j * LPR + gl) * VECW;
    col_ok[j] = c < W;
    col_b[j] = (col_ok[j] ? c : 0) * (int64_t)sizeof(T);
  }
  const char *m1 = reinterpret_cast<const char *>(mat1);
  const char *m2 = reinterpret_cast<const char *>(mat2);

  T *ring = nullptr;      // this group's ring; lane piece j of row r at r*CW + (j*LPR + gl)*VECW
  uint32_t ring_s = 0;
  if constexpr (PF > 0) {
    ring = reinterpret_cast<T *>(sddmm_smem) + (size_t)g * (NS * U * CW) + gl * VECW;
    ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  }

  A a[VPL][VECW];   // the current mat1 row, as accumulator type, zero outside the row
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < VECW; ++i) a[j][i] = A(0);

  auto load_a = [&](int64_t off1) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const VecT v = *reinterpret_cast<const VecT *>(m1 + off1 + col_b[j]);
#pragma unroll
      for (int i = 0; i < VECW; ++i) a[j][i] = col_ok[j] ? to_acc<T>(v.v[i]) : A(0);
    }
  };

  // batch operands of this lane: edge b + gl.  Edges past the chunk end alias edge e_end-1: valid
  // addresses, no extra segment heads, results never stored.
  auto load_idx = [&](int64_t b, int64_t &row, int64_t &off1, int64_t &off2) {
    const int64_t le = min(b + gl, e_end - 1);
    row = ld_stream(p.row_index + le, pol);
    off1 = row * row_bytes;
    off2 = ld_stream(p.col_index + le, pol) * row_bytes;
  };
  // U gathered mat2 rows of sub-batch [k0, k0+U) of a batch whose offsets are in `offs`: to registers ...
  auto load_v = [&](VecT(&v)[U][VPL], int64_t offs, int k0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off2 = __shfl_sync(gmask, offs, k0 + u, LPR);
#pragma unroll
      for (int j = 0; j < VPL; ++j) v[u][j] = *reinterpret_cast<const VecT *>(m2 + off2 + col_b[j]);
    }
  };
  // ... or into ring stage `stage`
  auto ring_issue = [&](int64_t offs, int k0, int stage) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off2 = __shfl_sync(gmask, offs, k0 + u, LPR);
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        cp_async_16(ring_s + (uint32_t)(((stage * U + u) * CW + j * LPR * VECW) * sizeof(T)), m2 + off2 + col_b[j]);
    }
  };

  // operands of the current batch (my_*) and the next (n_*); the one after is loaded at the top of each iteration
  int64_t my_row, my_off1, my_off2, n_row = 0, n_off1 = 0, n_off2 = 0;
  load_idx(e_begin, my_row, my_off1, my_off2);
  if (e_begin + LPR < e_end) load_idx(e_begin + LPR, n_row, n_off1, n_off2);
  VecT vc[U][VPL];                 // the sub-batch being consumed
  int stage = 0;
  if constexpr (PF > 0) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      ring_issue(my_off2, q * U, q);
      cp_async_commit();
    }
  } else {
    load_v(vc, my_off2, 0);
  }
  int64_t last_row = -1;           // row_index of the edge left of the current batch (none yet: forces a load)

  for (int64_t b = e_begin; b < e_end; b += LPR) {
    const bool has_next = b + LPR < e_end;
    int64_t nn_row = 0, nn_off1 = 0, nn_off2 = 0;
    if (b + 2 * LPR < e_end) load_idx(b + 2 * LPR, nn_row, nn_off1, nn_off2);

    int64_t left = __shfl_up_sync(gmask, my_row, 1, LPR);
    if (gl == 0) left = last_row;
    const unsigned bmask = (__ballot_sync(gmask, my_row != left) >> gshift) & low_bits<LPR>();
    last_row = __shfl_sync(gmask, my_row, LPR - 1, LPR);

    A mine = A(0);   // result of edge b + gl, collected from the sub-batches below
#pragma unroll
    for (int k0 = 0; k0 < LPR; k0 += U) {
      VecT vn[U][VPL];
      if constexpr (PF > 0) {
        // keep PF sub-batches in flight, then wait for the oldest and read this lane's own pieces back
        const int kk = k0 + PF * U;
        int st = stage + PF;
        if (st >= NS) st -= NS;
        if (kk < LPR) ring_issue(my_off2, kk, st);
        else if (has_next) ring_issue(n_off2, kk - LPR, st);
        cp_async_commit();
        cp_async_wait<PF>();
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            vc[u][j] = *reinterpret_cast<const VecT *>(ring + ((stage * U + u) * CW + j * LPR * VECW));
        stage = (stage + 1 == NS) ? 0 : stage + 1;
      } else {
        // next sub-batch's rows leave before this one is reduced (this batch, else the next batch's first)
        if (k0 + U < LPR) load_v(vn, my_off2, k0 + U);
        else if (has_next) load_v(vn, n_off2, 0);
      }

      A part[U];
      const unsigned sub = (bmask >> k0) & low_bits<U>();
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (sub != 0 && ((sub >> u) & 1u)) load_a(__shfl_sync(gmask, my_off1, k0 + u, LPR));
        A s = A(0);
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < VECW; ++i) s += a[j][i] * to_acc<T>(vc[u][j].v[i]);
        part[u] = s;
      }
      if constexpr (LPR > 1) {
        group_transpose_reduce<U, LPR, A>(part, gmask, gl);
        // value k sits in lanes [k*(LPR/U), (k+1)*(LPR/U)); lane gl wants edge gl = sub-batch gl/U, value gl%U
        const A r = __shfl_sync(gmask, part[0], (gl % U) * (LPR / U), LPR);
        if (gl / U == k0 / U) mine = r;
      } else {
        mine = part[0];
      }
      if constexpr (PF == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VPL; ++j) vc[u][j] = vn[u][j];
      }
    }
    if (b + gl < e_end) out[b + gl] = from_acc<T>(mine);
    my_row = n_row; my_off1 = n_off1; my_off2 = n_off2;
    n_row = nn_row; n_off1 = nn_off1; n_off2 = nn_off2;
  }
}

// Rows wider than a group can hold in registers (W > 32 lanes * 4 vectors): one warp per edge, columns in a loop.
template <typename T>
__global__ void __launch_bounds__(kThreads) sddmm_coo_wide_kernel(const SddmmParams p) {
  using A = typename AccOf<T>::type;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * kThreads) >> 5;
  const T *__restrict__ mat1 = static_cast<const T *>(p.mat1);
  const T *__restrict__ mat2 = static_cast<const T *>(p.mat2);
  T *__restrict__ out = static_cast<T *>(p.out);
  for (int64_t e = warp; e < p.E; e += n_warps) {
    const T *x = mat1 + p.row_index[e] * p.W;
    const T *y = mat2 + p.col_index[e] * p.W;
    A s = A(0);
    for (int64_t c = lane; c < p.W; c += 32) s += to_acc<T>(x[c]) * to_acc<T>(y[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[e] = from_acc<T>(s);
  }
}

}  // namespace geot
