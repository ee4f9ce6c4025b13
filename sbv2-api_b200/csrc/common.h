// Shared host-side helpers: status/exception plumbing for the C ABI, CUDA error checks.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sbv2_b200.h"

namespace sbv2 {

struct Error : std::runtime_error {
  sbv2_status code;
  Error(sbv2_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(sbv2_status c, const std::string& m) { throw Error(c, m); }

void set_last_error(const std::string& m);

// Runs `f`, converts exceptions into a status + thread-local message. Never lets anything escape.
template <class F>
int guarded(F&& f) noexcept {
  try {
    f();
    return SBV2_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return SBV2_ERR_INTERNAL;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return SBV2_ERR_INTERNAL;
  } catch (...) {
    set_last_error("unknown error");
    return SBV2_ERR_INTERNAL;
  }
}

#define SBV2_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) ::sbv2::fail(SBV2_ERR_INVALID_ARGUMENT, std::string(msg));          \
  } while (0)

}  // namespace sbv2

#ifdef __CUDACC__
#include <cuda_runtime.h>

#include <mutex>
#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      ::sbv2::fail(SBV2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +  \
                                      __FILE__ + ":" + std::to_string(__LINE__));                \
  } while (0)
#endif

#ifdef __CUDACC__
namespace sbv2 {
// Function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize ...) belong to the device/context, not to the
// process: a call site owns one PerDeviceOnce and runs its opt-in once per device the library is used on.  The mutex
// also covers two host threads that drive different models on the same device for the first time.
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;  // bit d: the opt-in ran on device ordinal d (ordinals >= 64 re-run it every time)
  template <class F>
  void run(F&& f) {
    int d = 0;
    CUDA_CHECK(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lock(mu);
    const unsigned long long bit = d < 64 ? (1ull << d) : 0ull;
    if (bit && (done & bit)) return;
    f();
    done |= bit;
  }
};
}  // namespace sbv2
#endif
