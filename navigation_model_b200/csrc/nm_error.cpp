// nm_version / nm_last_error: thread-local error reporting of the C-ABI.
#include <string.h>

#include "nm_common.h"

static thread_local char g_err[512] = "";

char *nm_error_buffer() { return g_err; }

int nm_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" int nm_version(void) { return NM_ABI_VERSION; }
extern "C" const char *nm_last_error(void) { return g_err; }
