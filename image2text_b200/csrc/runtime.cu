// Error plumbing, version, launch counter.
#include "common.cuh"

namespace i2t {

std::atomic<int64_t> g_launches{0};
std::atomic<int> g_pdl{1};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace i2t

extern "C" {
int i2t_version(void) { return 100; }
const char* i2t_last_error(void) { return i2t::err_buf(); }
int64_t i2t_launch_count(void) { return i2t::g_launches.load(); }
void i2t_set_pdl(int enabled) { i2t::g_pdl.store(enabled ? 1 : 0); }
}
