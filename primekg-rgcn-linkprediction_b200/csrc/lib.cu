// Library-level entry points: ABI version, thread-local error text, device check.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace rgcn {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

unsigned long long launches();

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RGCN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace rgcn

extern "C" int64_t rgcn_launch_count(void) { return (int64_t)rgcn::launches(); }

extern "C" int rgcn_abi_version(void) { return RGCN_B200_ABI_VERSION; }

extern "C" int rgcn_last_error(char* buf, size_t buf_len) {
  const char* e = rgcn::err_buf();
  const size_t n = strlen(e);
  if (buf && buf_len) {
    const size_t m = n < buf_len - 1 ? n : buf_len - 1;
    memcpy(buf, e, m);
    buf[m] = 0;
  }
  return (int)n;
}

extern "C" int rgcn_check_device(void) {
  int dev = 0, major = 0;
  RGCN_CUDA(cudaGetDevice(&dev));
  RGCN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    rgcn::set_error("this library is built for sm_100a only; current device has compute capability %d.x", major);
    return RGCN_EUNSUPPORTED;
  }
  return RGCN_OK;
}
