// Library-wide state: thread-local error message, launch counter, device properties.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace pdf {

static thread_local char g_error[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<bool> g_pdl{true};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed); }

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::atomic<int> g_sm_cap{0};

int num_sms() {
  const int cap = g_sm_cap.load(std::memory_order_relaxed);
  if (cap > 0) return cap;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

}  // namespace pdf

extern "C" int pdf_version(void) { return 100; }
extern "C" int pdf_debug_enable_pdl(int enable) {
  pdf::g_pdl.store(enable != 0, std::memory_order_relaxed);
  return PDF_OK;
}
/* tuning hook: persistent kernels size their grids for `cap` SMs instead of the device's (0 = off): lets two streams share the GPU */
extern "C" int pdf_debug_set_sm_cap(int cap) {
  pdf::g_sm_cap.store(cap < 0 ? 0 : cap, std::memory_order_relaxed);
  return PDF_OK;
}
extern "C" const char* pdf_last_error(void) { return pdf::g_error; }
extern "C" uint64_t pdf_launch_count(void) { return pdf::g_launches.load(std::memory_order_relaxed); }
