#include "common.cuh"

#include <atomic>

namespace mog {
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }
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

namespace mog { unsigned long long launches(); }
extern "C" unsigned long long mog_launch_count(void) { return mog::launches(); }
extern "C" int mog_version(void) { return MOG_VERSION; }
extern "C" const char* mog_last_error(void) { return mog::err_buf(); }
