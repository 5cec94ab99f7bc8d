// api.cu — library-wide C ABI plumbing: thread-local error string, version, launch accounting.
#include <atomic>
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" int b200_version(void) { return 100; }
extern "C" int64_t b200_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }
