#include "common.cuh"
#include <string.h>

namespace mimamo {
static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}
}  // namespace mimamo

extern "C" {
int mimamo_abi_version(void) { return 1; }
const char* mimamo_last_error(void) { return mimamo::t_error; }
uint64_t mimamo_launch_count(void) { return mimamo::g_launches.load(); }
}
