// Thread-local error text + ABI version of libmlg_b200.so.
#include <stdarg.h>

#include "common.cuh"
#include "../../include/mlg_b200.h"

static thread_local char g_err[512] = "";

void mlg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mlg_last_error(void) { return g_err; }
extern "C" int mlg_abi_version(void) { return MLG_ABI_VERSION; }
