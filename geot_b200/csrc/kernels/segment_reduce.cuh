// segment_reduce.cuh -- the one kernel family behind index_scatter / gather_scatter /
// gather_weight_scatter / mh_spmm on sm_100a.
//
// Replaces the reference's SR and PR kernel families (csrc/cuda/index_scatter_kernel.cuh:48-201,
// gather_scatter_kernel.cuh:22-186, gather_weight_scatter_kernel.cuh:22-185,
// mh_spmm_kernel.cuh:30-213) and their decision-tree dispatch (wrapper/*_rule.h).  Not a port:
//
//   * Edge-count partition.  The sorted edge list is cut into fixed-size CHUNKS of `chunk_edges`
//     edges; a GROUP of LPR lanes owns one chunk and a 256-thread CTA owns a TILE of 256/LPR
//     consecutive chunks.  Every CTA does the same number of edges whatever the degree skew.
//   * A group keeps the whole feature row in registers: lane gl holds VPL vectors of VECW elements
//     (128-bit loads when the row allows it), so one row = one coalesced 16*LPR-byte request.
//   * Segment detection from the sorted index: each lane loads one dst index of the batch, compares
//     it with its left neighbour (shfl_up) and a ballot gives the group a bitmask of segment heads;
//     a batch with no head runs the branch-free accumulate loop.
//   * No atomics and no pre-zeroed dst.  A segment that lies inside a chunk is stored once with a
//     plain vector store.  A segment cut by a chunk boundary leaves a partial ("carry") in shared
//     memory; after one __syncthreads the owner of the segment's first partial adds the later ones
//     in order.  Only segments cut by a TILE boundary leave the CTA: their partials go to the
//     workspace and segment_fixup_kernel finishes them (one chain per cut segment).  The summation
//     tree is fixed by (E, chunk_edges) alone, so results are bit-reproducible run to run.
//   * U independent row loads are issued before the first is consumed (memory-level parallelism
//     for the L2-resident gather); index / weight streams are read once per edge, coalesced.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace geot {

enum : int { RED_SUM = 0, RED_MEAN = 1, RED_MAX = 2, RED_MIN = 3, RED_PROD = 4 };
enum : int { FLAG_HEAD = 1, FLAG_THROUGH = 2, FLAG_TAIL = 4 };

constexpr int kThreads = 256;

struct Params {
  const void *src;
  const int64_t *src_index;  // null: src row = edge id
  const int64_t *dst_index;
  const void *weight;        // null: no weight
  void *dst;
  int64_t E;
  int64_t W;                 // row width in elements (H*F)
  int64_t F;                 // per-head width
  int64_t ws_e, ws_h;        // weight element (e,h) at weight[e*ws_e + h*ws_h]
  int per_head_weight;       // 1: weight differs per head (mh_spmm); 0: one weight per edge
  int mean;                  // 1: divide by the segment length at the end
  int chunk_edges;           // edges per group chunk
  int64_t n_tiles;
  // carries of segments cut by tile boundaries (workspace)
  void *carry_head;          // [n_tiles][W] accumulator type
  void *carry_tail;          // [n_tiles][W]
  long long *head_cnt;       // [n_tiles]
  long long *tail_cnt;       // [n_tiles]
  int64_t *tail_row;         // [n_tiles]
  unsigned char *flags;      // [n_tiles]
};

// ---- scalar type helpers -----------------------------------------------------------------------
template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename AccOf<T>::type to_acc(T v) { return v; }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_acc(typename AccOf<T>::type v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }

template <typename T, int N> struct alignas(sizeof(T) * N) Vec { T v[N]; };

template <int RED, typename A> __device__ __forceinline__ A red_identity() {
  if (RED == RED_MAX) return -INFINITY;
  if (RED == RED_MIN) return INFINITY;
  if (RED == RED_PROD) return A(1);
  return A(0);
}
// NaN-propagating max/min, as torch amax/amin (ATen ReduceUtils.h _max/_min).
template <int RED, typename A> __device__ __forceinline__ A red_op(A a, A x) {
  if (RED == RED_MAX) return (x > a || x != x) ? x : a;
  if (RED == RED_MIN) return (x < a || x != x) ? x : a;
  if (RED == RED_PROD) return a * x;
  return a + x;
}

template <int LPR> __device__ __forceinline__ unsigned group_mask(int lane) {
  if constexpr (LPR == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << LPR) - 1u) << (lane & ~(LPR - 1));
  }
}

// ---- main kernel -------------------------------------------------------------------------------
template <typename T, int VECW, int LPR, int VPL, int RED>
__global__ void __launch_bounds__(kThreads)
segment_reduce_kernel(const Params p) {
  using A = typename AccOf<T>::type;
  using VecT = Vec<T, VECW>;
  constexpr int NG = kThreads / LPR;      // chunks per tile
  constexpr int CW = LPR * VPL * VECW;    // columns per CTA (grid.y tiles wider rows)
  constexpr int U = (VPL >= 4) ? 2 : (VPL == 2 ? 4 : 8);  // row loads in flight per group

  extern __shared__ __align__(16) unsigned char smem_raw[];
  A *s_head = reinterpret_cast<A *>(smem_raw);              // [NG][CW]
  A *s_tail = s_head + NG * CW;                             // [NG][CW]
  long long *s_head_cnt = reinterpret_cast<long long *>(s_tail + NG * CW);  // [NG]
  long long *s_tail_cnt = s_head_cnt + NG;                  // [NG]
  int64_t *s_head_row = reinterpret_cast<int64_t *>(s_tail_cnt + NG);       // [NG]
  int64_t *s_tail_row = s_head_row + NG;                    // [NG]
  int *s_flags = reinterpret_cast<int *>(s_tail_row + NG);  // [NG]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = tid / LPR;
  const int gl = tid % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const int gshift = lane & ~(LPR - 1);

  const int64_t tile = blockIdx.x;
  const int64_t col0 = (int64_t)blockIdx.y * CW;
  const int64_t E = p.E, W = p.W;
  const int C = p.chunk_edges;
  const int64_t e_begin = (tile * NG + g) * (int64_t)C;
  const int64_t e_end = min(e_begin + (int64_t)C, E);

  const T *__restrict__ src = static_cast<const T *>(p.src);
  const T *__restrict__ weight = static_cast<const T *>(p.weight);
  const int64_t *__restrict__ dst_index = p.dst_index;
  const int64_t *__restrict__ src_index = p.src_index;
  T *__restrict__ dst = static_cast<T *>(p.dst);

  // this lane's columns and (mh_spmm) the weight offset of the head each vector belongs to
  int64_t col[VPL];
  bool col_ok[VPL];
  int64_t hoff[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    col[j] = col0 + (int64_t)(j * LPR + gl) * VECW;
    col_ok[j] = col[j] < W;
    hoff[j] = (p.per_head_weight && col_ok[j]) ? (col[j] / p.F) * p.ws_h : 0;
  }

  A acc[VPL][VECW];
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < VECW; ++i) acc[j][i] = red_identity<RED, A>();
  long long cnt = 0;
  int flags = 0;

  auto finalize_store = [&](int64_t row, A(&a)[VPL][VECW], long long n) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      if (!col_ok[j]) continue;
      VecT out;
#pragma unroll
      for (int i = 0; i < VECW; ++i) {
        A v = a[j][i];
        if (p.mean) v = v / static_cast<A>(n);
        out.v[i] = from_acc<T>(v);
      }
      *reinterpret_cast<VecT *>(dst + row * W + col[j]) = out;
    }
  };

  if (e_begin < e_end) {
    const int64_t prev_row = (e_begin > 0) ? dst_index[e_begin - 1] : -1;
    const int64_t next_row = (e_end < E) ? dst_index[e_end] : -1;
    int64_t cur_row = dst_index[e_begin];
    int64_t last_dst = cur_row;          // dst of the edge left of the current batch
    bool is_head = (cur_row == prev_row);

    // ends the current run: a complete segment is stored, a cut one is parked in shared memory
    auto flush = [&](bool continues) {
      if (is_head) {
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < VECW; ++i) s_head[g * CW + (j * LPR + gl) * VECW + i] = acc[j][i];
        if (gl == 0) { s_head_cnt[g] = cnt; s_head_row[g] = cur_row; }
        flags |= FLAG_HEAD | (continues ? FLAG_THROUGH : 0);
      } else if (continues) {
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < VECW; ++i) s_tail[g * CW + (j * LPR + gl) * VECW + i] = acc[j][i];
        if (gl == 0) { s_tail_cnt[g] = cnt; s_tail_row[g] = cur_row; }
        flags |= FLAG_TAIL;
      } else {
        finalize_store(cur_row, acc, cnt);
      }
#pragma unroll
      for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int i = 0; i < VECW; ++i) acc[j][i] = red_identity<RED, A>();
      cnt = 0;
      is_head = false;
    };

    for (int64_t b = e_begin; b < e_end; b += LPR) {
      const int nb = (int)min((int64_t)LPR, e_end - b);
      const int64_t my_e = b + gl;
      const bool valid = gl < nb;
      const int64_t my_dst = valid ? dst_index[my_e] : last_dst;
      const int64_t my_src = valid ? (src_index ? src_index[my_e] : my_e) : 0;
      A my_w = A(1);
      if (weight != nullptr && !p.per_head_weight && valid) my_w = to_acc<T>(weight[my_e * p.ws_e]);

      // segment heads of this batch as a bitmask (bit k: edge b+k starts a new dst row)
      int64_t left = __shfl_up_sync(gmask, my_dst, 1, LPR);
      if (gl == 0) left = last_dst;
      const unsigned bmask = (__ballot_sync(gmask, valid && my_dst != left) >> gshift);
      last_dst = __shfl_sync(gmask, my_dst, nb - 1, LPR);

      for (int k0 = 0; k0 < nb; k0 += U) {
        VecT v[U][VPL];
        A w[U][VPL];
        // ---- issue U row loads --------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = k0 + u;
          const int ks = (k < nb) ? k : (nb - 1);   // keep shuffles convergent; result unused if k>=nb
          const int64_t s = __shfl_sync(gmask, my_src, ks, LPR);
          const A we = __shfl_sync(gmask, my_w, ks, LPR);
          if (k < nb) {
            const T *rowp = src + s * W;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if (col_ok[j]) v[u][j] = *reinterpret_cast<const VecT *>(rowp + col[j]);
              w[u][j] = we;
              if (p.per_head_weight && col_ok[j])
                w[u][j] = to_acc<T>(weight[(b + k) * p.ws_e + hoff[j]]);
            }
          }
        }
        // ---- consume in edge order ------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = k0 + u;
          if (k < nb) {
            if ((bmask >> k) & 1u) {
              flush(false);
              cur_row = __shfl_sync(gmask, my_dst, k, LPR);
            }
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if (!col_ok[j]) continue;
#pragma unroll
              for (int i = 0; i < VECW; ++i) {
                A x = to_acc<T>(v[u][j].v[i]);
                if (weight != nullptr) x = x * w[u][j];
                acc[j][i] = red_op<RED, A>(acc[j][i], x);
              }
            }
            ++cnt;
          }
        }
      }
    }
    flush(cur_row == next_row);
  }
  if (gl == 0) s_flags[g] = flags;
  __syncthreads();

  // ---- join the partials of segments cut by chunk boundaries, inside the tile -------------------
  // A chain starts at a TAIL partial (or at the tile's own HEAD, chunk 0) and runs over the HEAD
  // partials of the following chunks while they are THROUGH.  Its owner sums it left to right.
  auto run_chain = [&](A *first, long long n0, int64_t row, int j_start, bool from_tile_head) {
    A a[VPL][VECW];
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int i = 0; i < VECW; ++i) a[j][i] = first[(j * LPR + gl) * VECW + i];
    long long n = n0;
    bool open = true;   // still waiting for the partial that ends the segment
    for (int c = j_start; c < NG && open; ++c) {
      const int f = s_flags[c];
#pragma unroll
      for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int i = 0; i < VECW; ++i)
          a[j][i] = red_op<RED, A>(a[j][i], s_head[c * CW + (j * LPR + gl) * VECW + i]);
      n += s_head_cnt[c];
      open = (f & FLAG_THROUGH) != 0;
    }
    if (!open && !from_tile_head) {
      finalize_store(row, a, n);   // the segment starts and ends inside this tile
      return;
    }
    // the segment crosses a tile boundary: park the partial in the workspace
    A *out = static_cast<A *>(from_tile_head ? p.carry_head : p.carry_tail) + tile * W;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      if (!col_ok[j]) continue;
#pragma unroll
      for (int i = 0; i < VECW; ++i) out[col[j] + i] = a[j][i];
    }
    if (gl == 0 && blockIdx.y == 0) {
      if (from_tile_head) p.head_cnt[tile] = n;
      else { p.tail_cnt[tile] = n; p.tail_row[tile] = row; }
    }
  };

  if (g == 0 && (flags & FLAG_HEAD)) {
    // chain that entered the tile from the left; it continues while chunks are THROUGH
    if (flags & FLAG_THROUGH) run_chain(s_head, s_head_cnt[0], s_head_row[0], 1, true);
    else run_chain(s_head, s_head_cnt[0], s_head_row[0], NG, true);
  }
  if (flags & FLAG_TAIL) run_chain(s_tail + g * CW, s_tail_cnt[g], s_tail_row[g], g + 1, false);

  // tile-level flags follow from the index alone
  if (tid == 0 && blockIdx.y == 0) {
    const int64_t t_begin = tile * NG * (int64_t)C;
    const int64_t t_end = min(t_begin + (int64_t)NG * C, E);
    const int64_t first_row = dst_index[t_begin], last_row = dst_index[t_end - 1];
    const bool head = t_begin > 0 && dst_index[t_begin - 1] == first_row;
    const bool cont = t_end < E && dst_index[t_end] == last_row;
    const bool through = head && cont && first_row == last_row;
    p.flags[tile] = (unsigned char)((head ? FLAG_HEAD : 0) | (through ? FLAG_THROUGH : 0) |
                                    ((cont && !through) ? FLAG_TAIL : 0));
  }
}

// ---- fixup: finish the segments cut by tile boundaries -------------------------------------------
// One warp per tile whose last segment continues into the next tile: tail[t] + head[t+1] + ... up to
// and including the first head that is not THROUGH.  Lanes stride over the W columns.
template <typename T, int RED>
__global__ void __launch_bounds__(kThreads)
segment_fixup_kernel(const Params p) {
  using A = typename AccOf<T>::type;
  const int64_t t = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= p.n_tiles) return;
  if (!(p.flags[t] & FLAG_TAIL)) return;
  const int64_t W = p.W;
  const A *tail = static_cast<const A *>(p.carry_tail) + t * W;
  const A *head = static_cast<const A *>(p.carry_head);
  T *dst = static_cast<T *>(p.dst) + p.tail_row[t] * W;
  // chain length first (flags are bytes: cheap), then the sums column by column
  int64_t last = t + 1;
  while (p.flags[last] & FLAG_THROUGH) ++last;
  long long n = p.tail_cnt[t];
  for (int64_t j = t + 1; j <= last; ++j) n += p.head_cnt[j];
  for (int64_t c = lane; c < W; c += 32) {
    A a = tail[c];
    for (int64_t j = t + 1; j <= last; ++j) a = red_op<RED, A>(a, head[j * W + c]);
    if (p.mean) a = a / static_cast<A>(n);
    dst[c] = from_acc<T>(a);
  }
}

}  // namespace geot
