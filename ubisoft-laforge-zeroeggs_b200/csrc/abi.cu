// Error plumbing, launch counter and version for the C ABI (include/zeggs_b200.h).
#include <stdarg.h>
#include <atomic>
#include "common.cuh"
#include "../../include/zeggs_b200.h"

namespace zeggs {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace zeggs

extern "C" const char* zeggs_last_error(void) { return zeggs::get_error(); }
extern "C" int zeggs_version(void) { return 100; }
extern "C" long long zeggs_launch_count(void) { return zeggs::g_launches.load(); }
