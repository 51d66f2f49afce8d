#include "common.cuh"

namespace mog {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace mog

extern "C" int mog_version(void) { return MOG_VERSION; }
extern "C" const char* mog_last_error(void) { return mog::err_buf(); }
