/* The C ABI from plain C99 (no C++ anywhere in this translation unit): the header must compile with
 * `gcc -std=c99 -pedantic`, the library must link and its host-only entry points must work without a GPU.
 * Compute entry points need a device: de_context_create reports DE_ERR_CUDA here and tests/ -m gpu cover the rest. */
#include <stdio.h>

#include "dune_eigensolver_b200.h"

int main(void)
{
  double a[4] = {2.0, 1.0, 1.0, 2.0}, w[2], v[4], block[16];
  de_context *ctx = NULL;
  int status;
  if (de_version() <= 0)
    return 1;
  if (de_host_sym_eig(2, a, w, v) != DE_OK) /* eigenvalues 1 and 3 */
    return 2;
  if (de_start_block(2, 8, 123u, block) != DE_OK) /* the reference's start block, eigensolver.hh:50-55 */
    return 3;
  status = de_context_create(0, NULL, &ctx);
  printf("version %d eig %.12g %.12g start %.17g context %d\n", de_version(), w[0], w[1], block[0], status);
  if (status != DE_OK)
    printf("error %s\n", de_last_error_string(NULL));
  else
    de_context_destroy(ctx);
  return 0;
}
