/* Minimal C99 client of include/vkjit_b200.h: builds a trace through the C ABI, checks typing / ref-counts /
 * error reporting and the generated source.  Needs no GPU (device calls must fail with VKJIT_ERR_NO_DEVICE
 * when there is none). */
#include <stdio.h>
#include <string.h>
#include "vkjit_b200.h"

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "FAILED: %s (line %d): %s\n", #x, __LINE__, vkjit_last_error()); return 1; } } while (0)

int main(void) {
  vkjit_ir* ir = NULL;
  vkjit_var a, c, m, z;
  vkjit_type ty;
  uint32_t rc;
  char src[8192];
  size_t len = 0, cubin = 0;
  CHECK(vkjit_abi_version() == VKJIT_B200_ABI_VERSION);
  CHECK(vkjit_ir_create(&ir) == VKJIT_OK);
  CHECK(vkjit_arange(ir, VKJIT_TY_U32, 1000, &a) == VKJIT_OK);
  CHECK(vkjit_const_f32(ir, 0.5f, &c) == VKJIT_OK);
  CHECK(vkjit_bop(ir, VKJIT_BOP_MUL, a, c, &m) == VKJIT_OK);          /* U32 * F32 -> F32 (autocast) */
  CHECK(vkjit_bop(ir, VKJIT_BOP_ADD, m, c, &z) == VKJIT_OK);
  CHECK(vkjit_var_type(ir, z, &ty) == VKJIT_OK && ty == VKJIT_TY_F32);
  CHECK(vkjit_var_ref_count(ir, c, &rc) == VKJIT_OK && rc == 3);      /* handle + two users */
  CHECK(vkjit_debug_codegen(ir, &z, 1, 1, src, sizeof src, &len, &cubin) == VKJIT_OK);
  CHECK(len > 0 && cubin > 0 && strstr(src, "vkjit_trace") != NULL);
  CHECK(vkjit_bop(ir, VKJIT_BOP_ADD, z, 123456, &m) == VKJIT_ERR_INVALID);   /* bad handle -> status, not a crash */
  CHECK(strlen(vkjit_last_error()) > 0);
  if (!vkjit_is_initialized()) {
    vkjit_status st = vkjit_eval(ir, &z, 1);
    CHECK(st == VKJIT_ERR_NO_DEVICE || st == VKJIT_OK);
  }
  CHECK(vkjit_dec_ref(ir, z) == VKJIT_OK);
  CHECK(vkjit_ir_destroy(ir) == VKJIT_OK);
  printf("c client ok (%zu bytes of CUDA C, %zu bytes of cubin)\n", len, cubin);
  return 0;
}
