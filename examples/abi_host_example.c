/*
 * abi_host_example.c -- a plain-C host of libgeot_b200.so (no torch, no C++): the call a cgo / JNI / ctypes stub makes.
 *
 *   gcc -std=c99 -Iinclude examples/abi_host_example.c -Lgeot_b200/lib -lgeot_b200 -Wl,-rpath,$PWD/geot_b200/lib -o abi_host_example
 *
 * gather_weight_scatter on a 6-edge graph with every operand in host memory (geot_b200_segment_reduce_host); needs a
 * B200 to run (there is no CPU path: without a device the call returns GEOT_ERR_CUDA and the text says why).  The
 * pure-host helper geot_b200_host_row_pointers runs anywhere.
 */
#include <stdio.h>
#include <stdlib.h>

#include "geot_b200.h"

int main(void) {
  /* dst-sorted COO edge list: dst 0 <- {1, 2}, dst 1 <- {0}, dst 3 <- {1, 2, 3}; row 2 has no edge */
  const int64_t dst_index[6] = {0, 0, 1, 3, 3, 3};
  const int64_t src_index[6] = {1, 2, 0, 1, 2, 3};
  const float weight[6] = {0.5f, 0.5f, 1.0f, 1.0f, 2.0f, 3.0f};
  const float src[4 * 2] = {1, 10, 2, 20, 3, 30, 4, 40}; /* [4, 2] */
  float dst[4 * 2];
  int64_t rowptr[5];
  int st, i;

  printf("libgeot_b200 version %d, SASS for sm_%d\n", geot_b200_version(), geot_b200_arch());

  st = geot_b200_host_row_pointers(dst_index, 6, 0, 4, rowptr, 1); /* CSR row pointer == geot::coo_to_csr */
  if (st != GEOT_OK) return 1;
  printf("rowptr:");
  for (i = 0; i < 5; ++i) printf(" %lld", (long long)rowptr[i]);
  printf("\n"); /* 0 2 3 3 6 */

  st = geot_b200_segment_reduce_host(src, 4, src_index, dst_index, weight, dst, /*E=*/6, /*S=*/4, /*H=*/1, /*F=*/2,
                                     GEOT_F32, GEOT_SUM, GEOT_W_EDGE);
  if (st != GEOT_OK) {
    printf("segment_reduce_host: %s (%s)\n", geot_b200_status_string(st), geot_b200_last_cuda_error());
    return st == GEOT_ERR_CUDA ? 0 : 1; /* no device here: expected */
  }
  for (i = 0; i < 4; ++i) printf("dst[%d] = %g %g\n", i, dst[2 * i], dst[2 * i + 1]);
  /* dst[0] = 2.5 25, dst[1] = 1 10, dst[2] = 0 0, dst[3] = 20 200 */
  geot_b200_host_arena_release();
  return 0;
}
