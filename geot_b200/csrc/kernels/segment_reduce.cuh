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
//   * Index / weight streams are read once per edge, coalesced (lane l of the group loads edge l of
//     the batch and turns its src index into a row byte offset); offset, weight and the segment-head
//     flags reach the other lanes by shuffle / ballot.  These read-once streams carry an L2
//     evict-first policy so that they do not push the re-used src rows out of L2.
//   * Segment detection from the sorted index: each lane compares its dst index with its left
//     neighbour (shfl_up); the ballot is the batch's bitmask of segment heads.  A run of edges
//     without a head takes a branch-free accumulate loop.
//   * No atomics and no pre-zeroed dst.  A segment that lies inside a chunk is stored once with a
//     plain vector store.  A segment cut by a chunk boundary leaves a partial in shared memory;
//     after one __syncthreads the owner of the segment's first partial adds the later ones in
//     order.  Only segments cut by a TILE boundary leave the CTA: their partials go to the workspace
//     and segment_fixup_kernel finishes them.  The summation tree depends on (E, chunk_edges)
//     alone, so results are bit-reproducible run to run.
//   * U independent row loads are issued before the first is consumed, and the next batch's
//     index / weight loads are issued before the current batch is processed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace geot {

enum : int { RED_SUM = 0, RED_MEAN = 1, RED_MAX = 2, RED_MIN = 3, RED_PROD = 4 };
enum : int { FLAG_HEAD = 1, FLAG_THROUGH = 2, FLAG_TAIL = 4 };
// weight handling inside the kernel: none / one weight per edge (shuffled) / generic (per-lane loads:
// the per-head weights of mh_spmm; also serves the rarely used reduce ops for every weight layout)
enum : int { WM_NONE = 0, WM_EDGE = 1, WM_GENERIC = 2 };

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
  int mean;                  // 1: divide by the segment length at the end
  int accumulate;            // 1: dst[row] += result instead of dst[row] = result (sum / mean; later passes of a bucketed
                             //    reduction: dist.py two-bucket exchange, src-blocked passes)
  int zero_gaps;             // 1: rows of [fill_lo, fill_hi) that receive no edge are zero-filled by the group that sees
                             //    the jump in dst_index (replaces a memset of the whole dst)
  int64_t fill_lo, fill_hi;  // zero_gaps: the rows this call owns (rows left of the first edge's / right of the last
                             //    edge's row included)
  const int64_t *mean_rowptr;  // mean over a bucketed reduction: divide by rowptr[row+1] - rowptr[row] (the row's degree
                             //    in the COMPLETE edge list) instead of by this pass's run length; null: run length
  const int32_t *edge_perm;  // weight of edge e is weight[edge_perm[e]] (bucketed edge lists keep the caller's weight
                             //    order); null: weight[e].  One weight per edge only (WM_EDGE)
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
template <int N> __device__ __forceinline__ unsigned low_bits() {
  if constexpr (N >= 32) return 0xffffffffu;
  else return (1u << N) - 1u;
}

// Read-once streams (indices, per-edge weights): bypass L1, first in line for L2 eviction.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int64_t ld_stream(const int64_t *p, uint64_t pol) {
  int64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t ld_stream32(const int32_t *p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <typename T> __device__ __forceinline__ T ld_stream_t(const T *p, uint64_t) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_stream_t<float>(const float *p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

// ---- main kernel -------------------------------------------------------------------------------
// tuning knobs (compile time): row loads in flight per group, minimum resident CTAs per SM
#ifndef GEOT_U0
#define GEOT_U0 8
#endif
#ifndef GEOT_MINB
#define GEOT_MINB 2
#endif
#ifndef GEOT_RING_U
#define GEOT_RING_U 0      // tuning builds: 16-byte pieces per lane per ring sub-batch (rows = GEOT_RING_U / VPL); 0 = built-in choice
#endif

// Shape constants shared by the kernel and its launcher.
//   PF == 0: gathered rows go global -> registers (U loads in flight per group).
//   PF  > 0: gathered rows are staged through a per-group shared-memory ring by cp.async (LDGSTS.128, L1
//            bypassed): PF sub-batches of U rows are in flight while one is consumed, so the bytes in flight
//            are bounded by shared memory (up to ~128 KB per SM) instead of by registers.  Every lane reads
//            back only the 16-byte pieces it copied itself: cp.async.wait_group is the only synchronisation.
//   (Two TMA-filled rings were built and measured, then removed: one 1-D bulk copy per row (cp.async.bulk, UBLKCP) was
//   15-40 % slower than cp.async at 256 B - 1 KB rows (profiles/r01_ring_sweep_run3_tma.txt), and Blackwell's gather4
//   (cp.async.bulk.tensor.2d...tile::gather4, UTMALDG.2D.GATHER4: 4 rows per instruction) 3x slower on Reddit gws
//   (11.5 vs 3.9 ms, profiles/r02e_sass_gather4_ring.txt): the TMA unit serialises the row fetches of a gather.)
constexpr int kDepthMask = 15;
//   PF & 32 (kLeanFlag): the LEAN ring -- the same cp.async ring with the per-edge bookkeeping moved out of the
//            instruction stream.  The batch's src row ids (32 bit: a ring row is >= 128 bytes, so any valid row id
//            of a 180 GB device fits) and weights are parked in a small per-group shared-memory buffer and come
//            back U at a time with one broadcast LDS.128 each, instead of 3 SHFL + a 64-bit add per edge; the ring
//            has NS = depth + 1 stages with NS | (LPR / U), so every stage address is an immediate of the unrolled
//            batch; the run length is a difference of edge positions instead of a per-edge counter.  About 10 warp
//            instructions per 512-byte row instead of 32 (profiles/r01c_*).  sum kernels with fp32 accumulators,
//            no per-head weights, chunks that are whole batches.
constexpr int kLeanFlag = 32;
//   PF & 64 (kExtFlag, lean ring only): the instantiation that honours the options of geot_b200_segment_reduce_ex
//            (accumulate, zero_gaps, edge_perm, mean_rowptr).  The plain instantiation has none of that code: measured
//            5 % (Reddit gws) to 10 % (products gs64) faster than carrying the unused branches
//            (profiles/r02e_tune_ext_ab_gather4_ab.txt), so the launcher picks per call.
constexpr int kExtFlag = 64;
//   PF & 128 (kHeadFlag, lean ring only): per-head weights (mh_spmm: weight[e, h], H <= kHeadMax heads).  The batch's
//            weights are parked TRANSPOSED in the operand buffer, [head][edge], so that a lane reads the U weights of its
//            own head for a sub-batch with one LDS.128 -- instead of one __ldg per edge and lane on the first-generation
//            ring (arxiv-shape mh_spmm was 4x slower per edge than Reddit gws at the same row size).
constexpr int kHeadFlag = 128;
constexpr int kHeadMax = 8;
//   (A "lean register path" -- the same bookkeeping with the rows loaded straight into double-buffered registers, so
//   that a gathered byte crosses the L1TEX data pipe once -- was built and measured in round 2: 15-25 % SLOWER than the
//   ring on every gather workload, profiles/r02a_ring96_ab.txt; too few bytes in flight per SM.  Removed.  So was a
//   hybrid -- one or two rows of every ring sub-batch loaded straight into registers, the rest through cp.async: 18 % /
//   42 % slower on Reddit gws, profiles/r02z_mix_ring_ab.txt.)
template <typename T, int VECW, int LPR, int VPL, int PF_>
struct ShapeOf {
  using A = typename AccOf<T>::type;
  static constexpr bool LEAN = (PF_ & kLeanFlag) != 0;   // the lean ring
  static constexpr bool HEADW = LEAN && (PF_ & kHeadFlag) != 0;   // ... with per-head weights
  static constexpr int PF = PF_ & kDepthMask;           // ring depth
  static constexpr int NG = kThreads / LPR;      // chunks (groups) per tile
  static constexpr int CW = LPR * VPL * VECW;    // columns per CTA
  // ring: 4 rows per sub-batch (2 for the widest rows) measured best on B200 (profiles/r01_ring_sweep.md)
  static constexpr int RU = GEOT_RING_U > 0 ? (GEOT_RING_U / VPL > 0 ? GEOT_RING_U / VPL : 1) : (VPL >= 4 ? 2 : 4);
  // lean ring: 2 KB of rows per warp and stage, at least 4 sub-batches per batch
  static constexpr int LU = (VPL >= 4) ? 1 : (VPL == 2 ? 2 : (LPR >= 16 ? 4 : (LPR >= 8 ? 2 : 1)));
  static constexpr int U0 = LEAN ? LU : (PF > 0 ? RU : ((VPL >= 4) ? 2 : (VPL == 2 ? 4 : GEOT_U0)));
  static constexpr int U = (LPR < U0) ? LPR : U0;   // rows per sub-batch
  static constexpr int NS = PF + 1;              // ring stages
  static constexpr int SB = LPR / U;             // sub-batches per batch
  // carries: head (and tail, unless it is parked in the group's own drained ring) + per-chunk scalars
  static constexpr size_t tail_off = (size_t)NG * CW * sizeof(A);
  static constexpr size_t scalars_off = (LEAN ? 1 : 2) * (size_t)NG * CW * sizeof(A);
  static constexpr size_t carry_bytes = ((scalars_off + (size_t)NG * (4 * 8 + 4)) + 127) & ~(size_t)127;
  // lean: per group two operand buffers of LPR src row ids + LPR weights
  // (per-head weights: two batches of kHeadMax x LPR weights instead of LPR)
  static constexpr int ops_words = HEADW ? (2 * LPR + 2 * kHeadMax * LPR) : 4 * LPR;   // per group
  static constexpr size_t ops_bytes = LEAN ? (size_t)NG * ops_words * 4 : 0;
  static constexpr size_t ring_off = carry_bytes + ops_bytes;
  static constexpr size_t ring_bytes = PF > 0 ? (size_t)NG * NS * U * CW * sizeof(T) : 0;
  static constexpr size_t smem_bytes = ring_off + ring_bytes;
  static constexpr int max_blocks = (int)((227 * 1024) / (smem_bytes + 1024));
  static constexpr int min_blocks_direct = (VPL == 1 ? GEOT_MINB : (VPL == 2 && VECW * sizeof(T) <= 16 ? 2 : 1));
  static constexpr int min_blocks = PF == 0 ? min_blocks_direct : (max_blocks >= 3 ? 3 : (max_blocks >= 2 ? 2 : 1));
  static_assert(PF == 0 || LEAN || PF * U <= LPR, "the prefetch distance must stay within one batch ahead");
  static_assert(!LEAN || (PF > 0 && SB % NS == 0 && sizeof(A) == 4 && U * sizeof(T) >= sizeof(A)),
                "lean ring: the stages must divide the batch; fp32 accumulators; a stage holds a carry row");
};

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// base + row * row_bytes: 32 x 32 -> 64 bit multiply-add (IMAD.WIDE.U32)
__device__ __forceinline__ const char *row_addr(const char *base, uint32_t row, uint32_t row_bytes) {
  return base + (uint64_t)row * row_bytes;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Rows [lo, hi) receive no edge: they read 0.  A cold path (the jump in dst_index is seen by one group), kept out of
// line so that the close-a-row code of the main kernel stays small.
template <typename T, int VECW, int LPR, int VPL>
__device__ __noinline__ void fill_zero_rows(T *dst, int64_t W, int64_t lo, int64_t hi, int64_t col0, int gl) {
  Vec<T, VECW> z;
#pragma unroll
  for (int i = 0; i < VECW; ++i) z.v[i] = from_acc<T>(typename AccOf<T>::type(0));
#pragma unroll 1
  for (int64_t r = lo; r < hi; ++r) {
#pragma unroll 1
    for (int j = 0; j < VPL; ++j) {
      const int64_t c = col0 + (int64_t)(j * LPR + gl) * VECW;
      if (c < W) *reinterpret_cast<Vec<T, VECW> *>(dst + r * W + c) = z;
    }
  }
}

// v / n for an integer count n, n_inv = RN(1 / n): one multiply and two FMAs (Markstein's correction step: the residual
// v - q*n is exact in an FMA) instead of a division sequence per element -- equal to IEEE division on every tested
// input (tests/test_abi_host.py::test_mean_division_identity), and 10x fewer instructions at every store site.
template <typename A> __device__ __forceinline__ A div_by_count(A v, A n, A n_inv) {
  const A q = v * n_inv;
  const A r = fma(-q, n, v);
  return fma(r, n_inv, q);
}

template <typename T, int VECW, int LPR, int VPL, int RED, int WM, int PF_>
__global__ void __launch_bounds__(kThreads, (ShapeOf<T, VECW, LPR, VPL, PF_>::min_blocks))
segment_reduce_kernel(const __grid_constant__ Params p) {
  // the options of geot_b200_segment_reduce_ex: always compiled in, except in the plain lean-ring instantiation
  constexpr bool kExt = !ShapeOf<T, VECW, LPR, VPL, PF_>::LEAN || (PF_ & kExtFlag) != 0;
  asm volatile("griddepcontrol.launch_dependents;");      // the fixup grid may be set up now; it waits for this grid to finish
  const bool o_accumulate = kExt && p.accumulate != 0;
  const bool o_zero_gaps = kExt && p.zero_gaps != 0;
  const int64_t *const o_mean_rowptr = kExt ? p.mean_rowptr : nullptr;
  const int32_t *const o_edge_perm = kExt ? p.edge_perm : nullptr;
  using A = typename AccOf<T>::type;
  using VecT = Vec<T, VECW>;
  using SH = ShapeOf<T, VECW, LPR, VPL, PF_>;
  constexpr int PF = SH::PF;
  constexpr bool LEAN = SH::LEAN;
  constexpr int NG = SH::NG;      // chunks per tile
  constexpr int CW = SH::CW;      // columns per CTA
  constexpr int U = SH::U;        // row loads in flight per group (PF == 0) / rows per ring sub-batch
  constexpr int NS = SH::NS;
  static_assert(PF == 0 || VECW * sizeof(T) == 16, "the ring moves 16-byte pieces");
  static_assert(!LEAN || (RED == RED_SUM && (WM == WM_GENERIC) == SH::HEADW), "lean ring: sum; per-head weights in their own instantiation");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  A *s_head = reinterpret_cast<A *>(smem_raw);              // [NG][CW]
  // tail partial of chunk c: its own slot, or (lean ring) the start of group c's ring, which the group has
  // drained by the time it parks a tail
  auto tail_slot = [&](int c) -> A * {
    if constexpr (LEAN) return reinterpret_cast<A *>(smem_raw + SH::ring_off + (size_t)c * (NS * U * CW * sizeof(T)));
    else return reinterpret_cast<A *>(smem_raw + SH::tail_off) + c * CW;
  };
  long long *s_head_cnt = reinterpret_cast<long long *>(smem_raw + SH::scalars_off);  // [NG]
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

  // Column tiles are the slow index of the grid: every tile of column tile 0 is scheduled before
  // column tile 1, so a wide row is swept in column slabs.
  const int64_t tile = (int64_t)blockIdx.x % p.n_tiles;
  const int64_t col_tile = (int64_t)blockIdx.x / p.n_tiles;
  const int64_t col0 = col_tile * CW;
  const bool first_col_tile = (col_tile == 0);
  const int64_t E = p.E, W = p.W;
  const int C = p.chunk_edges;
  const int64_t e_begin = (tile * NG + g) * (int64_t)C;
  const int64_t e_end = min(e_begin + (int64_t)C, E);

  const T *__restrict__ src = static_cast<const T *>(p.src);
  const T *__restrict__ weight = static_cast<const T *>(p.weight);
  const int64_t *__restrict__ dst_index = p.dst_index;
  const int64_t *__restrict__ src_index = p.src_index;
  T *__restrict__ dst = static_cast<T *>(p.dst);
  const uint64_t pol = policy_evict_first();

  // This lane's columns.  Lanes past the row end re-read a valid column (same sectors as another
  // lane, no extra traffic) and never store: the hot loop carries no column predicate.
  int64_t col[VPL];
  bool col_ok[VPL];
  const char *lane_src[VPL];    // src + clamped column, as bytes
  const T *lane_w[VPL];         // WM_GENERIC: weight + head offset of this vector (null: no weight)
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    col[j] = col0 + (int64_t)(j * LPR + gl) * VECW;
    col_ok[j] = col[j] < W;
    const int64_t c = col_ok[j] ? col[j] : col0;
    lane_src[j] = reinterpret_cast<const char *>(src + c);
    lane_w[j] = (WM == WM_GENERIC && weight != nullptr) ? weight + (c / p.F) * p.ws_h : nullptr;
  }
  const int64_t row_bytes = W * (int64_t)sizeof(T);

  // this group's ring: NS stages of U rows; lane piece j of row r at r*CW + (j*LPR + gl)*VECW
  T *ring = nullptr;
  uint32_t ring_s = 0;
  if constexpr (PF > 0) {
    ring = reinterpret_cast<T *>(smem_raw + SH::ring_off) + (size_t)g * (NS * U * CW) + gl * VECW;
    ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  }
  A acc[VPL][VECW];    // the open run
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < VECW; ++i) acc[j][i] = red_identity<RED, A>();
  long long cnt = 0;
  int flags = 0;

  // parks a partial in this group's shared-memory slot (private to the group until the barrier)
  auto park = [&](A *slot, long long *slot_cnt, int64_t *slot_row, int64_t row) {   // slot: this group's [CW]
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      Vec<A, VECW> t;
#pragma unroll
      for (int i = 0; i < VECW; ++i) t.v[i] = acc[j][i];
      *reinterpret_cast<Vec<A, VECW> *>(slot + (j * LPR + gl) * VECW) = t;
    }
    if (gl == 0) { slot_cnt[g] = cnt; slot_row[g] = row; }
  };

  auto finalize_store = [&](int64_t row, A(&a)[VPL][VECW], long long n) {
    A nA = A(1), n_inv = A(1);
    if (p.mean) {
      if (o_mean_rowptr != nullptr) n = o_mean_rowptr[row + 1] - o_mean_rowptr[row];
      nA = static_cast<A>(n);
      n_inv = A(1) / nA;
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      if (!col_ok[j]) continue;
      VecT *q = reinterpret_cast<VecT *>(dst + row * W + col[j]);
      VecT out;
      if (o_accumulate) out = *q;
#pragma unroll
      for (int i = 0; i < VECW; ++i) {
        A v = a[j][i];
        if (p.mean) v = div_by_count<A>(v, nA, n_inv);
        if (o_accumulate) v = to_acc<T>(out.v[i]) + v;
        out.v[i] = from_acc<T>(v);
      }
      *q = out;
    }
  };
  // rows [lo, hi) receive no edge (zero_gaps)
  auto fill_gap = [&](int64_t lo, int64_t hi) { fill_zero_rows<T, VECW, LPR, VPL>(dst, W, lo, hi, col0, gl); };

  if (e_begin < e_end) {
    const int64_t prev_row = (e_begin > 0) ? dst_index[e_begin - 1] : -1;
    const int64_t next_row = (e_end < E) ? dst_index[e_end] : -1;
    int64_t last_dst = dst_index[e_begin];   // dst of the edge left of the current batch
    bool is_head = (last_dst == prev_row);   // the open run entered the chunk from the left
    if (o_zero_gaps) {
      const int64_t left_row = (e_begin > 0) ? prev_row : p.fill_lo - 1;
      if (last_dst > left_row + 1) fill_gap(left_row + 1, last_dst);
    }

    // closes the open run, whose row is `row`; the run that starts has row `next`
    auto close_run = [&](int64_t row, int64_t next) {
      if (next > row + 1) fill_gap(row + 1, next);         // (callers pass next = row + 1 unless zero_gaps)
      if (is_head) {
        park(s_head + g * CW, s_head_cnt, s_head_row, row);
        flags |= FLAG_HEAD;
        is_head = false;
      } else {
        finalize_store(row, acc, cnt);
      }
#pragma unroll
      for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int i = 0; i < VECW; ++i) acc[j][i] = red_identity<RED, A>();
      cnt = 0;
    };

    auto accumulate = [&](const VecT(&v)[VPL], const A(&w)[VPL]) {
#pragma unroll
      for (int j = 0; j < VPL; ++j)
#pragma unroll
        for (int i = 0; i < VECW; ++i) {
          A x = to_acc<T>(v[j].v[i]);
          if (WM != WM_NONE) x = x * w[j];
          acc[j][i] = red_op<RED, A>(acc[j][i], x);
        }
    };

    if constexpr (LEAN) {
      // ---- lean ring (see ShapeOf) ---------------------------------------------------------------------
      constexpr int SB = SH::SB;                   // sub-batches per batch; NS | SB
      constexpr int RING_WORDS = 2 * LPR;          // operand buffers: two batches, circular
      const int n_edges = (int)(e_end - e_begin);
      const int nfull = n_edges / LPR;             // whole batches; a remainder only in the edge list's last chunk
      const int n_ring = nfull * LPR;              // edges that go through the ring
      constexpr bool HEADW = SH::HEADW;            // per-head weights, parked [batch][head][edge]
      constexpr int NW = HEADW ? kHeadMax : 1;     // weights a lane loads per edge
      uint32_t *ids = reinterpret_cast<uint32_t *>(smem_raw + SH::carry_bytes) + g * SH::ops_words;      // src row ids
      float *wts = reinterpret_cast<float *>(ids + RING_WORDS);                                          // weights
      const uint32_t row_bytes32 = (uint32_t)p.W * (uint32_t)sizeof(T);
      uint32_t ld32 = (uint32_t)last_dst;          // dst row of the edge left of the current batch
      const int Hn = HEADW ? (int)(p.W / p.F) : 1; // heads (<= kHeadMax: the launcher checked)
      int hj[VPL];                                 // head of this lane's vector j
#pragma unroll
      for (int j = 0; j < VPL; ++j) hj[j] = HEADW ? (int)((col_ok[j] ? col[j] : col0) / p.F) : 0;
      // all H weights of an edge as one 16-byte load when they are contiguous and fill it ([E, H] layout)
      const bool w_vec16 = HEADW && p.ws_h == 1 && p.ws_e == Hn && Hn * (int)sizeof(T) == 16 &&
                           (reinterpret_cast<uintptr_t>(weight) & 15) == 0;

      // this lane's operands of batch bi: dst row, src row id, weight(s)
      auto ld_ops = [&](int bi, uint32_t &d, uint32_t &sid, float(&wv)[NW]) {
        const int64_t e = e_begin + (int64_t)bi * LPR + gl;
        d = (uint32_t)ld_stream(dst_index + e, pol);
        sid = src_index ? (uint32_t)ld_stream(src_index + e, pol) : (uint32_t)e;
#pragma unroll
        for (int h = 0; h < NW; ++h) wv[h] = 1.f;
        if constexpr (HEADW) {
          if (w_vec16) {
            constexpr int PER = 16 / (int)sizeof(T);
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(weight + e * PER));
            const Vec<T, PER> t = *reinterpret_cast<const Vec<T, PER> *>(&raw);
#pragma unroll
            for (int h = 0; h < PER && h < NW; ++h) wv[h] = to_acc<T>(t.v[h]);
          } else {
#pragma unroll
            for (int h = 0; h < NW; ++h)
              if (h < Hn) wv[h] = to_acc<T>(ld_stream_t<T>(weight + e * p.ws_e + h * p.ws_h, pol));
          }
        } else if (WM == WM_EDGE) {
          wv[0] = to_acc<T>(ld_stream_t<T>(weight + (o_edge_perm ? (int64_t)ld_stream32(o_edge_perm + e, pol) : e), pol));
        }
      };
      // parks this lane's operands of a batch in operand buffer half b (0 / 1)
      auto park_ops = [&](int b, uint32_t sid, const float(&wv)[NW]) {
        ids[b * LPR + gl] = sid;
        if constexpr (HEADW) {
#pragma unroll
          for (int h = 0; h < NW; ++h)
            if (h < Hn) wts[(b * kHeadMax + h) * LPR + gl] = wv[h];
        } else {
          wts[b * LPR + gl] = wv[0];
        }
      };
      // word offset in `wts` of the weights of (buffer half b, position k in the batch) for this lane's vector j
      auto w_at = [&](int b, int k, int j) -> int { return HEADW ? (b * kHeadMax + hj[j]) * LPR + k : b * LPR + k; };
      // copies the U rows whose ids are at o[0..U) into ring stage st (compile-time)
      auto issue = [&](const uint32_t *o, int st) {
        const Vec<uint32_t, U> r = *reinterpret_cast<const Vec<uint32_t, U> *>(o);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            cp_async_16(ring_s + (uint32_t)(((st * U + u) * CW + j * LPR * VECW) * sizeof(T)),
                        row_addr(lane_src[j], r.v[u], row_bytes32));
        }
      };
      auto add_edge = [&](const VecT(&v)[VPL], const float(&we)[VPL]) {
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
          for (int i = 0; i < VECW; ++i) {
            float x = to_acc<T>(v[j].v[i]);
            if (WM != WM_NONE) x = x * we[j];
            acc[j][i] = acc[j][i] + x;
          }
      };

      uint32_t d_cur = 0, d_nxt = 0, l_d = 0, l_s = 0;
      float l_w[NW];
      if (nfull > 0) {
        ld_ops(0, d_cur, l_s, l_w);
        park_ops(0, l_s, l_w);
      }
      if (nfull > 1) {
        ld_ops(1, d_nxt, l_s, l_w);
        park_ops(1, l_s, l_w);
      }
      __syncwarp(gmask);
      if (nfull > 0) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {      // PF < NS <= SB: all inside batch 0
          issue(ids + q * U, q);
          cp_async_commit();
        }
      }
      int pos = 0;          // chunk-relative position of the current sub-batch block
      int run_start = 0;    // chunk-relative position where the open run began
      int slot = 0;         // word offset of the current batch in the operand buffers: 0 or LPR
#pragma unroll 1
      for (int bi = 0; bi < nfull; ++bi) {
        const bool has_nn = bi + 2 < nfull;
        if (has_nn) ld_ops(bi + 2, l_d, l_s, l_w);     // parked at the end of this batch, used from the next one on
        // segment heads of this batch as a bitmask (bit k: edge k of the batch starts a new dst row)
        uint32_t left = __shfl_up_sync(gmask, d_cur, 1, LPR);
        if (gl == 0) left = ld32;
        unsigned bmask = (__ballot_sync(gmask, d_cur != left) >> gshift) & low_bits<LPR>();
        const uint32_t batch_left = ld32;
        ld32 = __shfl_sync(gmask, d_cur, LPR - 1, LPR);
        const int batch_pos = pos;
#pragma unroll 1
        for (int s0 = 0; s0 < SB; s0 += NS) {
          const int blk = slot + s0 * U;                              // word offset of this block of NS sub-batches
          const int blk_next = (blk + NS * U) & (RING_WORDS - 1);     // ... and of the one after (may be the next batch)
#pragma unroll
          for (int t = 0; t < NS; ++t) {
            // keep PF sub-batches in flight: issue the one PF ahead (stage (t + PF) % NS, the one consumed last)
            const int c = (t + PF) * U;
            if (pos + c < n_ring) issue(c < NS * U ? ids + blk + c : ids + blk_next + (c - NS * U), (t + PF) % NS);
            cp_async_commit();
            cp_async_wait<PF>();
            // the U weights of this sub-batch (per vector of the lane when the weights are per head)
            Vec<float, U> wv[VPL];
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
#pragma unroll
              for (int u = 0; u < U; ++u) wv[j].v[u] = 1.f;
              if (WM != WM_NONE && (HEADW || j == 0))
                wv[j] = *reinterpret_cast<const Vec<float, U> *>(wts + w_at(slot / LPR, (s0 + t) * U, j));
              else if (WM != WM_NONE)
                wv[j] = wv[0];
            }
            VecT v[U][VPL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < VPL; ++j)
                v[u][j] = *reinterpret_cast<const VecT *>(ring + ((t * U + u) * CW + j * LPR * VECW));
            const unsigned sub = bmask & low_bits<U>();
            bmask >>= U;
            if (sub == 0) {
#pragma unroll
              for (int u = 0; u < U; ++u) {
                float we[VPL];
#pragma unroll
                for (int j = 0; j < VPL; ++j) we[j] = wv[j].v[u];
                add_edge(v[u], we);
              }
            } else if constexpr (!kExt && !HEADW && VPL == 1 && LPR >= 16) {
              // a dst row starts inside these U edges: the same adds, straight from the registers, with the open run
              // closed in front of every edge that starts a row.  Only the close is a branch (a warp of two or four
              // groups diverges for the few instructions of the store, not for the edges), and nothing is read twice --
              // on short-row graphs (products: 25 edges per row, arxiv: 7) up to half of the sub-batches come here.
              // products gs64 1.52 -> 1.46 ms, its 8 shards 0.221-0.257 -> 0.212-0.245 (profiles/r02s_tune.txt, r02t_tune.txt).
              // (The instantiation with the segment_reduce_ex options, rows wider than 512 bytes and 128-byte rows keep
              // the rolled loop below: unrolled, their larger close paths spill; with per-head weights the unrolled form
              // needs 128 registers and ran arxiv mh_spmm 30 % SLOWER, r02t_tune.txt.)
#pragma unroll
              for (int u = 0; u < U; ++u) {
                if ((sub >> u) & 1u) {
                  const int k = pos + t * U + u;          // chunk-relative position of this edge
                  const int kb = k - batch_pos;           // position inside the batch
                  const uint32_t row = __shfl_sync(gmask, d_cur, kb > 0 ? kb - 1 : 0, LPR);
                  const int64_t prev = (int64_t)(kb > 0 ? row : batch_left);
                  cnt = k - run_start;
                  close_run(prev, prev + 1);
                  run_start = k;
                }
                float we[VPL];
#pragma unroll
                for (int j = 0; j < VPL; ++j) we[j] = wv[j].v[u];
                add_edge(v[u], we);
              }
            } else {
              // ... one edge at a time, operands re-read from shared memory
#pragma unroll 1
              for (int u = 0; u < U; ++u) {
                const int k = pos + t * U + u;            // chunk-relative position of this edge
                if ((sub >> u) & 1u) {
                  const int kb = k - batch_pos;           // position inside the batch
                  const uint32_t row = __shfl_sync(gmask, d_cur, kb > 0 ? kb - 1 : 0, LPR);
                  const int64_t prev = (int64_t)(kb > 0 ? row : batch_left);
                  const int64_t nxt = o_zero_gaps ? (int64_t)__shfl_sync(gmask, d_cur, kb, LPR) : prev + 1;
                  cnt = k - run_start;
                  close_run(prev, nxt);
                  run_start = k;
                }
                VecT vv[VPL];
#pragma unroll
                for (int j = 0; j < VPL; ++j)
                  vv[j] = *reinterpret_cast<const VecT *>(ring + ((t * U + u) * CW + j * LPR * VECW));
                float we[VPL];
#pragma unroll
                for (int j = 0; j < VPL; ++j) we[j] = (WM != WM_NONE) ? wts[w_at(slot / LPR, (s0 + t) * U + u, j)] : 1.f;
                add_edge(vv, we);
              }
            }
          }
          pos += NS * U;
        }
        // this batch's operand buffer is free: park batch bi+2 there
        __syncwarp(gmask);
        if (has_nn) park_ops(slot / LPR, l_s, l_w);
        __syncwarp(gmask);
        slot ^= LPR;
        d_cur = d_nxt;
        d_nxt = l_d;
      }
      cp_async_wait<0>();                         // nothing of this group is in flight: its ring may hold a tail now
      cnt = pos - run_start;
      // remainder of the edge list's last chunk: one edge at a time, uniform operand loads
      for (int k = pos; k < n_edges; ++k) {
        const int64_t e = e_begin + k;
        const uint32_t d = (uint32_t)dst_index[e];
        const int64_t sid = src_index ? src_index[e] : e;
        float we[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          we[j] = 1.f;
          if (HEADW) we[j] = to_acc<T>(weight[e * p.ws_e + hj[j] * p.ws_h]);
          else if (WM == WM_EDGE) we[j] = to_acc<T>(weight[o_edge_perm ? (int64_t)o_edge_perm[e] : e]);
        }
        VecT v[VPL];
#pragma unroll
        for (int j = 0; j < VPL; ++j) v[j] = *reinterpret_cast<const VecT *>(lane_src[j] + sid * row_bytes);
        if (d != ld32) close_run((int64_t)ld32, o_zero_gaps ? (int64_t)d : (int64_t)ld32 + 1);
        add_edge(v, we);
        ++cnt;
        ld32 = d;
      }
      last_dst = (int64_t)ld32;
    } else {
      // batch operands of this lane: edge (b + gl)
      auto load_batch = [&](int64_t b, int nb, int64_t &my_dst, int64_t &my_off, A &my_w) {
        const int64_t my_e = b + gl;
        const bool valid = gl < nb;
        my_dst = valid ? ld_stream(dst_index + my_e, pol) : (int64_t)-2;
        const int64_t s = valid ? (src_index ? ld_stream(src_index + my_e, pol) : my_e) : 0;
        my_off = s * row_bytes;
        my_w = A(1);
        if (WM == WM_EDGE && valid) my_w = to_acc<T>(ld_stream_t<T>(weight + (o_edge_perm ? (int64_t)ld_stream32(o_edge_perm + my_e, pol) : my_e), pol));
      };

      // operands of the current batch (my_*) and of the next one (n_*); the one after that is loaded at the
      // top of every iteration, so index / weight loads are two batches ahead of their first use
      int64_t my_dst, my_off, n_dst = -2, n_off = 0;
      A my_w, n_w = A(1);
      load_batch(e_begin, (int)min((int64_t)LPR, e_end - e_begin), my_dst, my_off, my_w);
      if (e_begin + LPR < e_end) load_batch(e_begin + LPR, (int)min((int64_t)LPR, e_end - e_begin - LPR), n_dst, n_off, n_w);

      // ring: copies sub-batch [kk, kk+U) of a batch (row offsets in `offs`, one per lane) into `stage`
      auto ring_issue = [&](int64_t offs, int kk, int stage) {
  #pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t off = __shfl_sync(gmask, offs, kk + u, LPR);
  #pragma unroll
          for (int j = 0; j < VPL; ++j)
            cp_async_16(ring_s + (uint32_t)(((stage * U + u) * CW + j * LPR * VECW) * sizeof(T)), lane_src[j] + off);
        }
      };
      int stage = 0;              // ring stage of the sub-batch consumed next
      if constexpr (PF > 0) {
        // prologue: the first PF sub-batches of the chunk (only a full first batch uses the ring)
        const bool full0 = (e_end - e_begin) >= LPR;
  #pragma unroll
        for (int q = 0; q < PF; ++q) {
          if (full0) ring_issue(my_off, q * U, q);
          cp_async_commit();
        }
      }

      for (int64_t b = e_begin; b < e_end; b += LPR) {
        const int nb = (int)min((int64_t)LPR, e_end - b);
        // issue the operand loads of the batch after the next before touching this one
        int64_t nn_dst = -2, nn_off = 0;
        A nn_w = A(1);
        if (b + 2 * LPR < e_end) load_batch(b + 2 * LPR, (int)min((int64_t)LPR, e_end - b - 2 * LPR), nn_dst, nn_off, nn_w);
        const bool next_full = (e_end - b) >= 2 * LPR;

        // segment heads of this batch as a bitmask (bit k: edge b+k starts a new dst row)
        int64_t left = __shfl_up_sync(gmask, my_dst, 1, LPR);
        if (gl == 0) left = last_dst;
        const unsigned bmask = (__ballot_sync(gmask, gl < nb && my_dst != left) >> gshift) & low_bits<LPR>();
        const int64_t batch_left = last_dst;
        last_dst = __shfl_sync(gmask, my_dst, nb - 1, LPR);
        const T *wb[VPL];                 // WM_GENERIC: this batch's weights for this lane's heads
        const int ws_e32 = (int)p.ws_e;
  #pragma unroll
        for (int j = 0; j < VPL; ++j) wb[j] = (WM == WM_GENERIC && lane_w[j] != nullptr) ? lane_w[j] + b * p.ws_e : nullptr;

        if (nb == LPR) {
          // ---- full batch: U loads in flight, branch-free when the U edges hold no segment head ----
  #pragma unroll 1
          for (int k0 = 0; k0 < LPR; k0 += U) {
            VecT v[U][VPL];
            A w[U][VPL];
            if constexpr (PF > 0) {
              // keep PF sub-batches in flight: issue the one PF ahead (this batch or the next), then wait for
              // the oldest and read this lane's own pieces back
              const int kk = k0 + PF * U;
              int st = stage + PF;
              if (st >= NS) st -= NS;
              if (kk < LPR) ring_issue(my_off, kk, st);
              else if (next_full) ring_issue(n_off, kk - LPR, st);
              cp_async_commit();
              cp_async_wait<PF>();
  #pragma unroll
              for (int u = 0; u < U; ++u) {
                const A we = (WM == WM_EDGE) ? __shfl_sync(gmask, my_w, k0 + u, LPR) : A(1);
  #pragma unroll
                for (int j = 0; j < VPL; ++j) {
                  v[u][j] = *reinterpret_cast<const VecT *>(ring + ((stage * U + u) * CW + j * LPR * VECW));
                  w[u][j] = we;
                  if (WM == WM_GENERIC && wb[j] != nullptr) w[u][j] = to_acc<T>(__ldg(wb[j] + (k0 + u) * ws_e32));
                }
              }
              stage = (stage + 1 == NS) ? 0 : stage + 1;
            } else {
  #pragma unroll
              for (int u = 0; u < U; ++u) {
                const int64_t off = __shfl_sync(gmask, my_off, k0 + u, LPR);
                const A we = (WM == WM_EDGE) ? __shfl_sync(gmask, my_w, k0 + u, LPR) : A(1);
  #pragma unroll
                for (int j = 0; j < VPL; ++j) {
                  v[u][j] = *reinterpret_cast<const VecT *>(lane_src[j] + off);
                  w[u][j] = we;
                  if (WM == WM_GENERIC && wb[j] != nullptr) w[u][j] = to_acc<T>(__ldg(wb[j] + (k0 + u) * ws_e32));
                }
              }
            }
            const unsigned sub = (bmask >> k0) & low_bits<U>();
            if (sub == 0) {
  #pragma unroll
              for (int u = 0; u < U; ++u) accumulate(v[u], w[u]);
              cnt += U;
            } else {
  #pragma unroll
              for (int u = 0; u < U; ++u) {
                if ((sub >> u) & 1u) {
                  const int k = k0 + u;
                  const int64_t row = (k == 0) ? batch_left : __shfl_sync(gmask, my_dst, (k == 0) ? 0 : k - 1, LPR);
                  close_run(row, o_zero_gaps ? __shfl_sync(gmask, my_dst, k, LPR) : row + 1);
                }
                accumulate(v[u], w[u]);
                ++cnt;
              }
            }
          }
        } else {
          // ---- ragged last batch of the edge list: one edge at a time ---------------------------------
          for (int k = 0; k < nb; ++k) {
            const int64_t off = __shfl_sync(gmask, my_off, k, LPR);
            const A we = (WM == WM_EDGE) ? __shfl_sync(gmask, my_w, k, LPR) : A(1);
            const int64_t row = __shfl_sync(gmask, my_dst, k > 0 ? k - 1 : 0, LPR);
            VecT v[VPL];
            A w[VPL];
  #pragma unroll
            for (int j = 0; j < VPL; ++j) {
              v[j] = *reinterpret_cast<const VecT *>(lane_src[j] + off);
              w[j] = we;
              if (WM == WM_GENERIC && wb[j] != nullptr) w[j] = to_acc<T>(__ldg(wb[j] + k * ws_e32));
            }
            if ((bmask >> k) & 1u) {
              const int64_t prev = (k == 0) ? batch_left : row;
              close_run(prev, o_zero_gaps ? __shfl_sync(gmask, my_dst, k, LPR) : prev + 1);
            }
            accumulate(v, w);
            ++cnt;
          }
        }
        my_dst = n_dst; my_off = n_off; my_w = n_w;
        n_dst = nn_dst; n_off = nn_off; n_w = nn_w;
      }
    }

    // the run still open at the chunk end
    const int64_t cur_row = last_dst;
    if (o_zero_gaps && e_end == E && p.fill_hi > cur_row + 1) fill_gap(cur_row + 1, p.fill_hi);
    const bool continues = (cur_row == next_row);
    if (is_head) {                      // the whole chunk is one run that entered from the left
      park(s_head + g * CW, s_head_cnt, s_head_row, cur_row);
      flags |= FLAG_HEAD | (continues ? FLAG_THROUGH : 0);
    } else if (continues) {
      park(tail_slot(g), s_tail_cnt, s_tail_row, cur_row);
      flags |= FLAG_TAIL;
    } else {
      finalize_store(cur_row, acc, cnt);
    }
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
    if (gl == 0 && first_col_tile) {
      if (from_tile_head) p.head_cnt[tile] = n;
      else { p.tail_cnt[tile] = n; p.tail_row[tile] = row; }
    }
  };

  // a group may own two chains: the one that entered the tile from the left (chunk 0 only; it continues while chunks
  // are THROUGH) and the one its own tail partial starts.  One call site, so that the chain code exists once.
  const bool own_head = (g == 0) && (flags & FLAG_HEAD);
  const bool own_tail = (flags & FLAG_TAIL) != 0;
  for (int which = own_head ? 0 : 1; which < (own_tail ? 2 : 1); ++which) {
    const bool h = (which == 0);
    run_chain(h ? s_head : tail_slot(g), h ? s_head_cnt[0] : s_tail_cnt[g], h ? s_head_row[0] : s_tail_row[g],
              h ? ((flags & FLAG_THROUGH) ? 1 : NG) : g + 1, h);
  }

  // tile-level flags follow from the index alone
  if (tid == 0 && first_col_tile) {
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
// A chain = tail[t] + head[t+1] + ... up to and including the first head that is not THROUGH.
// A CTA looks at 8 consecutive tiles.  Short chains: one warp each, lanes across the columns, the
// chain loop unrolled so that its loads overlap.  Long chains (a hub row cut by hundreds of tiles):
// all 8 warps stride over the chain, then the 8 partials are combined in warp order -- still a fixed
// summation order.
template <typename T, int RED>
__global__ void __launch_bounds__(kThreads)
segment_fixup_kernel(const __grid_constant__ Params p) {
  using A = typename AccOf<T>::type;
  // launched as a programmatic dependent of the main kernel (inst.cu): nothing of its output may be read before this
  asm volatile("griddepcontrol.wait;" ::: "memory");
  constexpr int NW = kThreads / 32;
  constexpr int kLong = 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t W = p.W;
  const A *head = static_cast<const A *>(p.carry_head);
  __shared__ int64_t s_last[NW];
  __shared__ int s_is_long[NW];
  __shared__ A s_part[NW * 32];
  __shared__ __align__(16) float s_part4[sizeof(A) == 4 ? kThreads * 4 : 4];   // long chains, fp32 carries: one float4 per thread

  const int64_t t = (int64_t)blockIdx.x * NW + warp;
  const bool has = (t < p.n_tiles) && (p.flags[t] & FLAG_TAIL);
  int64_t last = t + 1;
  if (has) {
    // chain end: scan the THROUGH flags 32 tiles at a time
    for (;;) {
      const int64_t q = last + lane;
      const bool thr = (q < p.n_tiles) && (p.flags[q] & FLAG_THROUGH);
      const unsigned m = __ballot_sync(0xffffffffu, !thr);
      if (m) { last += __ffs(m) - 1; break; }
      last += 32;
    }
  }
  const bool is_long = has && (last - t > kLong);
  if (lane == 0) { s_last[warp] = last; s_is_long[warp] = is_long ? 1 : 0; }

  if (has && !is_long) {
    long long n = p.tail_cnt[t];
    for (int64_t j = t + 1; j <= last; ++j) n += p.head_cnt[j];
    if (p.mean && p.mean_rowptr != nullptr) n = p.mean_rowptr[p.tail_row[t] + 1] - p.mean_rowptr[p.tail_row[t]];
    const A *tail = static_cast<const A *>(p.carry_tail) + t * W;
    T *dst = static_cast<T *>(p.dst) + p.tail_row[t] * W;
    for (int64_t c = lane; c < W; c += 32) {
      A a = tail[c];
      int64_t j = t + 1;
      for (; j + 3 <= last; j += 4) {
        const A x0 = head[j * W + c], x1 = head[(j + 1) * W + c], x2 = head[(j + 2) * W + c], x3 = head[(j + 3) * W + c];
        a = red_op<RED, A>(red_op<RED, A>(red_op<RED, A>(red_op<RED, A>(a, x0), x1), x2), x3);
      }
      for (; j <= last; ++j) a = red_op<RED, A>(a, head[j * W + c]);
      if (p.mean) a = a / static_cast<A>(n);
      if (p.accumulate) a = to_acc<T>(dst[c]) + a;
      dst[c] = from_acc<T>(a);
    }
  }
  __syncthreads();
  // long chains, one after the other, by the whole CTA
  for (int wv = 0; wv < NW; ++wv) {
    if (!s_is_long[wv]) continue;          // uniform across the CTA
    const int64_t tt = (int64_t)blockIdx.x * NW + wv;
    const int64_t lst = s_last[wv];
    long long n = 0;
    for (int64_t j = tt + 1 + lane; j <= lst; j += 32) n += p.head_cnt[j];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    n += p.tail_cnt[tt];
    if (p.mean && p.mean_rowptr != nullptr) n = p.mean_rowptr[p.tail_row[tt] + 1] - p.mean_rowptr[p.tail_row[tt]];
    if constexpr (sizeof(A) == 4) {
      // fp32 carries in whole 16-byte vectors: VT threads across a carry row, 256 / VT tiles side by side, 8 vector loads
      // in flight per thread (a products-shape shard whose hub row is cut by 741 tiles: 32 -> 9 us; the chain used to be
      // walked one 32-column block at a time with 4 scalar loads in flight).  Fixed order: per-lane partial sums over
      // tiles tl, tl + TL, ..., then the TL partials left to right, then the tail carry.
      if ((W & 3) == 0) {
        const int nvec = (int)(W >> 2);
        int VT = 1;
        while (VT < nvec && VT < kThreads) VT <<= 1;
        const int TL = kThreads / VT;
        const int v = threadIdx.x % VT, tl = threadIdx.x / VT;
        float4 *s_vec = reinterpret_cast<float4 *>(s_part4);
        for (int v0 = 0; v0 < nvec; v0 += VT) {
          const int vv = v0 + v;
          float4 a = make_float4(red_identity<RED, float>(), red_identity<RED, float>(), red_identity<RED, float>(),
                                 red_identity<RED, float>());
          if (vv < nvec) {
            const float4 *hv = reinterpret_cast<const float4 *>(head) + vv;
            const int64_t rowv = W >> 2;
            int64_t j = tt + 1 + tl;
            constexpr int UN = 8;
            for (; j + (int64_t)(UN - 1) * TL <= lst; j += (int64_t)UN * TL) {
              float4 x[UN];
#pragma unroll
              for (int u = 0; u < UN; ++u) x[u] = __ldcg(hv + (j + (int64_t)u * TL) * rowv);
#pragma unroll
              for (int u = 0; u < UN; ++u) {
                a.x = red_op<RED, float>(a.x, x[u].x); a.y = red_op<RED, float>(a.y, x[u].y);
                a.z = red_op<RED, float>(a.z, x[u].z); a.w = red_op<RED, float>(a.w, x[u].w);
              }
            }
            for (; j <= lst; j += TL) {
              const float4 x = __ldcg(hv + j * rowv);
              a.x = red_op<RED, float>(a.x, x.x); a.y = red_op<RED, float>(a.y, x.y);
              a.z = red_op<RED, float>(a.z, x.z); a.w = red_op<RED, float>(a.w, x.w);
            }
          }
          s_vec[tl * VT + v] = a;
          __syncthreads();
          if (tl == 0 && vv < nvec) {
            const float4 t4 = reinterpret_cast<const float4 *>(static_cast<const float *>(p.carry_tail) + tt * W)[vv];
            float r[4] = {t4.x, t4.y, t4.z, t4.w};
            for (int k = 0; k < TL; ++k) {
              const float4 q = s_vec[k * VT + v];
              r[0] = red_op<RED, float>(r[0], q.x); r[1] = red_op<RED, float>(r[1], q.y);
              r[2] = red_op<RED, float>(r[2], q.z); r[3] = red_op<RED, float>(r[3], q.w);
            }
            T *q = static_cast<T *>(p.dst) + p.tail_row[tt] * W + (int64_t)vv * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float o = r[i];
              if (p.mean) o = o / static_cast<float>(n);
              if (p.accumulate) o = to_acc<T>(q[i]) + o;
              q[i] = from_acc<T>(o);
            }
          }
          __syncthreads();
        }
        continue;
      }
    }
    for (int64_t c0 = 0; c0 < W; c0 += 32) {
      const int64_t c = c0 + lane;
      A a = red_identity<RED, A>();
      if (c < W) {
        int64_t j = tt + 1 + warp;
        for (; j + 3 * NW <= lst; j += 4 * NW) {
          const A x0 = head[j * W + c], x1 = head[(j + NW) * W + c], x2 = head[(j + 2 * NW) * W + c], x3 = head[(j + 3 * NW) * W + c];
          a = red_op<RED, A>(red_op<RED, A>(red_op<RED, A>(red_op<RED, A>(a, x0), x1), x2), x3);
        }
        for (; j <= lst; j += NW) a = red_op<RED, A>(a, head[j * W + c]);
      }
      s_part[warp * 32 + lane] = a;
      __syncthreads();
      if (warp == 0 && c < W) {
        A r = static_cast<const A *>(p.carry_tail)[tt * W + c];
#pragma unroll
        for (int k = 0; k < NW; ++k) r = red_op<RED, A>(r, s_part[k * 32 + lane]);
        if (p.mean) r = r / static_cast<A>(n);
        T *q = static_cast<T *>(p.dst) + p.tail_row[tt] * W + c;
        if (p.accumulate) r = to_acc<T>(*q) + r;
        *q = from_acc<T>(r);
      }
      __syncthreads();
    }
  }
}

}  // namespace geot
