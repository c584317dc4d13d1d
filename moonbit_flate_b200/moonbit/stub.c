/* Native stub linked into the MoonBit package: fb200_create returns its handle
 * through an out-parameter, which MoonBit's FFI cannot express for an opaque
 * type, so this wrapper returns it (and aborts when there is no device: the
 * package has no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>

#include "flate_b200.h"

fb200_ctx *fb200_create_ret(int device)
{
  fb200_ctx *ctx = NULL;
  int rc = fb200_create(&ctx, device);
  if (rc != FB200_OK) {
    fprintf(stderr, "flate_b200: fb200_create failed (%d): no usable sm_100 CUDA device, no CPU fallback\n", rc);
    abort();
  }
  return ctx;
}
