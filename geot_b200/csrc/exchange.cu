// exchange.cu -- device helpers of the multi-GPU path (geot_b200/dist.py).  New functionality:
// the reference has no distributed code.
//
// The dst-sharded gather ops overlap the exchange of the src rows with the reduction: a rank's edges are split once
// per graph into a src-local and a src-remote bucket (stable, so both stay dst-sorted); the local bucket is reduced
// while the remote rows travel, the remote bucket afterwards with accumulate (segment_reduce.cuh).  Here:
// push_rows (the fused pack + transfer over peer memory of the "push" transport) and permute_edges (per-head
// weights into bucket order; one weight per edge needs no pass: the kernel reads weight[edge_perm[e]]).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/geot_b200.h"

extern "C" int geot_b200_set_cuda_error(const char *what, int cuda_error);

namespace {

// out[e*words + k] = in[perm[e]*words + k]: edge operands (weights) into bucket order, 4-byte words
__global__ void __launch_bounds__(256)
permute_edges_kernel(const uint32_t *__restrict__ in, const int64_t *__restrict__ perm, uint32_t *__restrict__ out,
                     int64_t E, int words) {
  const int64_t n = E * words;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / words;
    const int k = (int)(i - e * words);
    out[i] = in[perm[e] * words + k];
  }
}
__global__ void __launch_bounds__(256)
permute_edges16_kernel(const uint16_t *__restrict__ in, const int64_t *__restrict__ perm, uint16_t *__restrict__ out,
                       int64_t E, int halves) {
  const int64_t n = E * halves;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / halves;
    const int k = (int)(i - e * halves);
    out[i] = in[perm[e] * halves + k];
  }
}

// 16-byte pieces: packs whole feature rows (the rows a peer asked for) at full vector width
__global__ void __launch_bounds__(256)
permute_rows16_kernel(const uint4 *__restrict__ in, const int64_t *__restrict__ perm, uint4 *__restrict__ out, int64_t E,
                      int vecs) {
  const int64_t n = E * vecs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / vecs;
    const int k = (int)(i - e * vecs);
    out[i] = in[perm[e] * vecs + k];
  }
}

// Fused pack + transfer over peer memory: row e of the send list goes straight into the receive buffer of the
// GPU that asked for it -- bases[dest_peer[e]] is that GPU's buffer mapped into this process (NVLink P2P stores,
// 16-byte vectors), dest_row[e] the row slot the receiver's src ids point at.  No staging buffer, no NCCL call.
template <typename V>
__global__ void __launch_bounds__(256)
push_rows_kernel(const V *__restrict__ x, const int64_t *__restrict__ rows, const int32_t *__restrict__ dest_peer,
                 const int64_t *__restrict__ dest_row, V *const *__restrict__ bases, int64_t n, int vecs) {
  const int64_t total = n * vecs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / vecs;
    const int k = (int)(i - e * vecs);
    V *base = bases[dest_peer[e]];
    base[dest_row[e] * vecs + k] = x[rows[e] * vecs + k];
  }
}

}  // namespace

extern "C" {

int geot_b200_permute_edges(const void *in, const int64_t *perm, void *out, int64_t E, int64_t bytes_per_edge,
                            cudaStream_t stream) {
  if (E < 0 || bytes_per_edge <= 0 || (bytes_per_edge & 1)) return GEOT_ERR_INVALID_ARG;
  if (E == 0) return GEOT_OK;       // nothing to move (an empty tensor's pointer may be null)
  if (!in || !perm || !out) return GEOT_ERR_INVALID_ARG;
  // widest move that both the record size and the base pointers allow (a view may start mid-vector)
  const uintptr_t addr = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out);
  if (addr & 1) return GEOT_ERR_INVALID_ARG;
  const bool vec16 = bytes_per_edge % 16 == 0 && (addr & 15) == 0;
  const bool vec4 = bytes_per_edge % 4 == 0 && (addr & 3) == 0;
  const int64_t n = E * (vec16 ? bytes_per_edge / 16 : (vec4 ? bytes_per_edge / 4 : bytes_per_edge / 2));
  const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (vec16)
    permute_rows16_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint4 *>(in), perm, static_cast<uint4 *>(out), E,
                                                      (int)(bytes_per_edge / 16));
  else if (vec4)
    permute_edges_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint32_t *>(in), perm, static_cast<uint32_t *>(out), E,
                                                     (int)(bytes_per_edge / 4));
  else
    permute_edges16_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint16_t *>(in), perm, static_cast<uint16_t *>(out),
                                                       E, (int)(bytes_per_edge / 2));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return geot_b200_set_cuda_error("permute_edges_kernel", (int)e);
  return GEOT_OK;
}

int geot_b200_push_rows(const void *x, const int64_t *rows, const int32_t *dest_peer, const int64_t *dest_row,
                        void *const *peer_bases, int64_t n, int64_t row_bytes, int peers_aligned16, cudaStream_t stream) {
  if (n < 0 || row_bytes <= 0 || (row_bytes & 1)) return GEOT_ERR_INVALID_ARG;
  if (n == 0) return GEOT_OK;
  if (!x || !rows || !dest_peer || !dest_row || !peer_bases) return GEOT_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(x) & 1) return GEOT_ERR_INVALID_ARG;
  // the peer bases live in device memory: the caller vouches for their alignment (symmetric allocations are)
  const bool vec16 = row_bytes % 16 == 0 && peers_aligned16 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const bool vec4 = row_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0;      // (bases: >= 4-byte aligned rows)
  const int64_t total = n * (vec16 ? row_bytes / 16 : (vec4 ? row_bytes / 4 : row_bytes / 2));
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  if (vec16)
    push_rows_kernel<uint4><<<blocks, 256, 0, stream>>>(static_cast<const uint4 *>(x), rows, dest_peer, dest_row,
                                                        reinterpret_cast<uint4 *const *>(peer_bases), n, (int)(row_bytes / 16));
  else if (vec4)
    push_rows_kernel<uint32_t><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t *>(x), rows, dest_peer, dest_row,
                                                           reinterpret_cast<uint32_t *const *>(peer_bases), n,
                                                           (int)(row_bytes / 4));
  else      // bf16 / fp16 rows with an odd element count
    push_rows_kernel<uint16_t><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t *>(x), rows, dest_peer, dest_row,
                                                           reinterpret_cast<uint16_t *const *>(peer_bases), n,
                                                           (int)(row_bytes / 2));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return geot_b200_set_cuda_error("push_rows_kernel", (int)e);
  return GEOT_OK;
}

}  // extern "C"
