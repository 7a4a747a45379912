/* A plain C99 client of include/rubix_b200.h: proves that the header is valid C (no C++, no CUDA, no torch types in
 * any signature) and that a non-Python host can link the library and get ERROR CODES, not crashes, when it has no
 * device or passes bad arguments.  Built and run by tests/test_library_abi.py with gcc -std=c99 -Wall -Wextra -Werror. */
#include <stdio.h>
#include <string.h>

#include "rubix_b200.h"

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAILED: %s (line %d)\n", #cond, __LINE__); ++failures; } \
  } while (0)

int main(void) {
  int w = 0, ws = 0;
  int64_t v = 0;
  CHECK(rbx_version() >= 100);
  CHECK(rbx_last_error() != NULL);
  CHECK(rbx_launch_count() >= 0);
  /* sizes and geometry need no device */
  CHECK(rbx_slab_geometry(3721, 8, 12, &w, &ws) == RBX_OK && w == 466 && ws == 490);
  CHECK(rbx_sort_by_spaxel_workspace_bytes(1000000, 625) >= (size_t)12000000);
  CHECK(rbx_set_option("psub", 128) == RBX_OK && rbx_get_option("psub", &v) == RBX_OK && v == 128);
  CHECK(rbx_set_option("psub", -1) == RBX_OK);
  CHECK(rbx_set_option("no_such_switch", 1) != RBX_OK && strstr(rbx_last_error(), "unknown option") != NULL);
  /* argument validation comes before any CUDA call */
  CHECK(rbx_spaxel_assign(NULL, 5, NULL, 1, NULL, NULL, NULL) == RBX_ERR_INVALID_ARGUMENT);
  CHECK(rbx_ssp_lookup(NULL, NULL, NULL, 0, NULL, NULL) == RBX_ERR_INVALID_ARGUMENT);
  CHECK(strstr(rbx_last_error(), "null plan") != NULL);
  CHECK(rbx_sort_by_spaxel(NULL, 10, 0, NULL, NULL, NULL, NULL, 0, NULL) == RBX_ERR_INVALID_ARGUMENT);
  CHECK(rbx_segment_sum_sorted(NULL, NULL, NULL, 0, 0, NULL, NULL) == RBX_ERR_INVALID_ARGUMENT);
  CHECK(rbx_segment_sum_sorted(NULL, NULL, NULL, 10, 5, (float *)&w, NULL) == RBX_ERR_INVALID_ARGUMENT);
  CHECK(rbx_reduce_cube(NULL, NULL, NULL, 0, 0, NULL) != RBX_OK);
  CHECK(rbx_build_cube_workspace_bytes(NULL, 10, 25) == 0);
  {
    rbx_plan *plan = NULL;
    CHECK(rbx_plan_create(&plan, NULL, 2, NULL, 2, NULL, 2, NULL, NULL, 1, 0.1, RBX_METHOD_LINEAR, 2, NULL) ==
          RBX_ERR_INVALID_ARGUMENT);
    CHECK(plan == NULL);
  }
  printf(failures ? "abi_client: %d check(s) failed\n" : "abi_client: ok\n", failures);
  return failures ? 1 : 0;
}
