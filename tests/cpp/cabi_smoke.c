/* Plain C99 consumer of the boundary, as a cgo / JNI / ctypes stub would bind it: dlopen() the library, dlsym() every entry
 * point named on the command line, and drive the calls that need no device ("boundary" tier).
 *   usage: cabi_smoke <path to libglim_b200.so> <symbol> ...
 * Exit code 0 = every symbol resolved and the status protocol behaves; prints one summary line. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "glim_b200.h"

typedef const char* (*status_string_fn)(gb_status);
typedef const char* (*last_error_fn)(void);
typedef int (*device_count_fn)(void);
typedef gb_status (*ctx_create_fn)(int, gb_ctx**);
typedef gb_status (*ctx_destroy_fn)(gb_ctx*);
typedef gb_status (*hessian_blocks_fn)(const gb_linearized6*, double, double*, double*, double*, double*, double*, double*);

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  int missing = 0;
  for (int i = 2; i < argc; i++)
    if (!dlsym(h, argv[i])) { fprintf(stderr, "missing symbol %s\n", argv[i]); missing++; }
  if (missing) return 4;

  status_string_fn status_string;
  *(void**)(&status_string) = dlsym(h, "gb_status_string"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  last_error_fn last_error;
  *(void**)(&last_error) = dlsym(h, "gb_last_error"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  device_count_fn device_count;
  *(void**)(&device_count) = dlsym(h, "gb_device_count"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  ctx_create_fn ctx_create;
  *(void**)(&ctx_create) = dlsym(h, "gb_ctx_create"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  ctx_destroy_fn ctx_destroy;
  *(void**)(&ctx_destroy) = dlsym(h, "gb_ctx_destroy"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  hessian_blocks_fn hessian_blocks;
  *(void**)(&hessian_blocks) = dlsym(h, "gb_hessian_blocks"); /* the POSIX idiom: ISO C has no object -> function pointer cast */
  if (strcmp(status_string(GB_OK), "ok") != 0) return 5;

  /* a host-only entry point: HessianFactor blocks of a record (sign convention of SURVEY A.3) */
  gb_linearized6 lin;
  memset(&lin, 0, sizeof(lin));
  for (int k = 0; k < 6; k++) { lin.b_t[k] = k + 1.0; lin.b_s[k] = -(k + 1.0); lin.H_tt[k * 6 + k] = 2.0; lin.H_ss[k * 6 + k] = 3.0; }
  lin.error = 4.0;
  double G11[36], G12[36], g1[6], G22[36], g2[6], f = 0.0;
  if (hessian_blocks(&lin, 0.5, G11, G12, g1, G22, g2, &f) != GB_OK) return 6;
  if (g1[2] != -3.0 || g2[2] != 3.0 || G11[7] != 2.0 || G22[14] != 3.0 || f != 2.0) return 7;
  if (hessian_blocks(NULL, 1.0, G11, G12, g1, G22, g2, &f) != GB_ERR_INVALID_ARGUMENT) return 8;

  const int ndev = device_count();
  gb_ctx* ctx = NULL;
  const gb_status st = ctx_create(0, &ctx);
  if (ndev <= 0) {
    if (st != GB_ERR_NO_DEVICE || ctx != NULL || strlen(last_error()) == 0) return 9; /* no CPU fallback: loud failure */
  } else {
    if (st != GB_OK || ctx == NULL) { fprintf(stderr, "gb_ctx_create: %s\n", last_error()); return 10; }
    if (ctx_destroy(ctx) != GB_OK) return 11;
  }
  printf("cabi_smoke ok: %d symbols, %d device(s), ctx_create -> %s\n", argc - 2, ndev, status_string(st));
  dlclose(h);
  return 0;
}
