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

/* Writer objects in their pull form (no callback crosses the FFI): the compressed bytes queue up inside the
 * object and the MoonBit side collects them with fb200_writer_take after every write / close. */
fb200_writer *fb200_writer_new_pull(fb200_ctx *ctx) { return fb200_writer_new(ctx, NULL, NULL); }

fb200_writer *fb200_writer_new_dict_pull(fb200_ctx *ctx, const uint8_t *dict, uint64_t n)
{
  return fb200_writer_new_dict(ctx, NULL, NULL, dict, n);
}
