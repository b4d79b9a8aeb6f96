// Library-wide state: error string, launch counter, precision switch, device check, seq lens.
#include <stdarg.h>

#include "common.cuh"

namespace ds2 {
static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
static std::atomic<int> g_prec{DS2_PREC_FP32};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int precision() { return g_prec.load(std::memory_order_relaxed); }
}  // namespace ds2

extern "C" {

const char* ds2_version(void) { return "ds2_b200 0.1.0 (sm_100a)"; }
const char* ds2_last_error(void) { return ds2::g_err; }

int ds2_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  DS2_CHECK_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) {
    ds2::set_error("no CUDA device");
    return DS2_ERR_CUDA;
  }
  int dev = 0;
  DS2_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DS2_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) {
    ds2::set_error("device %s is sm_%d%d; this library is built for sm_100a only", p.name, p.major, p.minor);
    return DS2_ERR_CUDA;
  }
  return DS2_OK;
}

int ds2_set_precision(int prec) {
  DS2_REQUIRE(prec == DS2_PREC_FP32 || prec == DS2_PREC_TF32, "unknown precision %d", prec);
  ds2::g_prec.store(prec);
  return DS2_OK;
}
int ds2_get_precision(void) { return ds2::precision(); }

int64_t ds2_launch_count(int reset) {
  long long v = reset ? ds2::g_launches.exchange(0) : ds2::g_launches.load();
  return (int64_t)v;
}

int ds2_seq_lens_host(const int32_t* in_len, int n, int32_t* out_len) {
  DS2_REQUIRE(in_len && out_len && n >= 0, "ds2_seq_lens_host: bad arguments");
  // model.py:299-310 with the two Conv2d time-axis geometries (k=11,p=5,d=1; strides 2 and 1).
  // Python floor division == C division here because the numerator is >= -1 only for len = 0;
  // use an explicit floor to stay exact for every int32 input.
  const int k = 11, p = 5, dil = 1, strides[2] = {2, 1};
  for (int i = 0; i < n; ++i) {
    long long L = in_len[i];
    for (int c = 0; c < 2; ++c) {
      long long num = L + 2 * p - dil * (k - 1) - 1;
      long long q = num / strides[c];
      if ((num % strides[c] != 0) && ((num < 0) != (strides[c] < 0))) --q;
      L = q + 1;
    }
    out_len[i] = (int32_t)L;
  }
  return DS2_OK;
}

}  // extern "C"
