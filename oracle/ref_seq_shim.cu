// ref_seq_shim.cu -- C-callable handles on the REFERENCE's own golden definitions, compiled from the
// reference sources where they lie (nothing is copied): the sequential loops of
// csrc/util/check.cuh:77-111 (segment_coo_sequencial, gather_sequencial, gws_sequencial) and the
// synthetic sorted-index generator csrc/dataloader/dataloader.hpp:21-62 (generateIndex).
// Host code only; built by oracle/Makefile into oracle/_ref/libref_seq.so when /root/reference is
// present.  TEST INFRASTRUCTURE ONLY: used to pin oracle/geot_oracle.c and to generate
// tests/golden/*.npz (tests/golden/make_golden.py).
#include <cstdint>
#include <vector>
#include "util/check.cuh"
#include "dataloader/dataloader.hpp"

extern "C" {

void ref_segment_coo_sequencial_f32(const float *src, const int64_t *index, int nnz, int N,
                                    int dst_len, float *dst) {
  util::segment_coo_sequencial<float, int64_t>(src, index, nnz, N, dst_len, dst);
}

void ref_gather_sequencial_f32(const float *src, const int64_t *index, int nnz, int N, int dst_len,
                               float *dst) {
  util::gather_sequencial<float, int64_t>(src, index, nnz, N, dst_len, dst);
}

void ref_gws_sequencial_f32(float *dst, const float *src, int64_t *row, int64_t *col,
                            const float *weight, int nnz, int N, int dst_len) {
  util::gws_sequencial<float, int64_t>(dst, src, row, col, weight, nnz, N, dst_len);
}

// returns dst_len; out must hold total_count entries
int ref_generate_index(int range, int min_seg, int max_seg, int total_count, double cv,
                       int64_t *out) {
  std::vector<int64_t> v;
  int dst_len = generateIndex<int64_t>(range, min_seg, max_seg, total_count, cv, v);
  for (int i = 0; i < total_count; ++i) out[i] = v[i];
  return dst_len;
}
}
