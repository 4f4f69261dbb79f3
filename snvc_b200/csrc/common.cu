#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace snvc {

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

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace {
const char* const kOptNames[OPT_COUNT] = {"SNVC_CONV_MODE", "SNVC_CONV_STORE", "SNVC_CONV_OCC", "SNVC_CONV_MAXGRID",
                                          "SNVC_CV_SPLIT_OLD", "SNVC_CV_THREADS", "SNVC_ROI_MODE", "SNVC_LIFT_MODE", "SNVC_CV_WALK"};
struct Options {
  char val[OPT_COUNT][32];
  bool set[OPT_COUNT];
  Options() {
    for (int i = 0; i < OPT_COUNT; ++i) assign(i, getenv(kOptNames[i]));
  }
  void assign(int i, const char* v) {
    set[i] = v != nullptr && v[0] != 0;
    snprintf(val[i], sizeof(val[i]), "%s", set[i] ? v : "");
  }
};
Options g_options;   // constructed when the shared library is loaded
}  // namespace

const char* opt(OptId id) { return g_options.set[id] ? g_options.val[id] : nullptr; }

}  // namespace snvc

extern "C" int snvc_set_option(const char* name, const char* value) {
  using namespace snvc;
  if (name == nullptr) {                       // reset: every option unset
    for (int i = 0; i < OPT_COUNT; ++i) g_options.assign(i, nullptr);
    return 0;
  }
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, kOptNames[i]) == 0) {
      g_options.assign(i, value);
      return 0;
    }
  return fail(SNVC_E_BADARG, "unknown option %s", name);
}
extern "C" const char* snvc_get_option(const char* name) {
  using namespace snvc;
  for (int i = 0; name && i < OPT_COUNT; ++i)
    if (strcmp(name, kOptNames[i]) == 0) return opt((OptId)i);
  return nullptr;
}

namespace snvc { int64_t launches(); }
extern "C" int64_t snvc_launch_count(void) { return snvc::launches(); }
extern "C" int snvc_version(void) { return SNVC_ABI_VERSION; }
extern "C" const char* snvc_last_error(void) { return snvc::err_buf(); }
