// Error state, launch counter and version of the C-ABI library.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace vilco {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<int> g_act_fmt{VILCO_BF16};
static std::atomic<float> g_grad_scale{1.0f};
int act_fmt() { return g_act_fmt.load(std::memory_order_relaxed); }
float grad_scale() { return g_grad_scale.load(std::memory_order_relaxed); }
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
}  // namespace vilco

extern "C" const char* vilco_last_error(void) { return vilco::g_err; }
extern "C" int vilco_version(void) { return 3; }
extern "C" uint64_t vilco_launch_count(void) { return vilco::g_launches.load(std::memory_order_relaxed); }
extern "C" int vilco_set_plane_format(int fmt) {
  if (fmt != VILCO_BF16 && fmt != VILCO_F16) {
    vilco::set_error("vilco_set_plane_format: fmt must be VILCO_BF16 or VILCO_F16");
    return VILCO_E_ARG;
  }
  vilco::g_act_fmt.store(fmt, std::memory_order_relaxed);
  return VILCO_OK;
}
extern "C" int vilco_get_plane_format(void) { return vilco::act_fmt(); }
extern "C" int vilco_set_grad_scale(float s) {
  if (!(s > 0.f)) {
    vilco::set_error("vilco_set_grad_scale: scale must be positive");
    return VILCO_E_ARG;
  }
  vilco::g_grad_scale.store(s, std::memory_order_relaxed);
  return VILCO_OK;
}
