// Host-side helpers shared by the .cu translation units of libb200whisper.so
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

namespace b200w {

typedef __nv_bfloat16 bf16;

void set_last_error(const std::string& s);
const char* get_last_error();

struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define B200W_CHECK_CUDA(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _buf[512];                                                                           \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,           \
               cudaGetErrorString(_e));                                                         \
      throw ::b200w::Error(_buf);                                                               \
    }                                                                                           \
  } while (0)

#define B200W_REQUIRE(cond, msg)                                                                \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      char _buf[512];                                                                           \
      snprintf(_buf, sizeof(_buf), "%s:%d: requirement (%s) failed: %s", __FILE__, __LINE__,    \
               #cond, msg);                                                                     \
      throw ::b200w::Error(_buf);                                                               \
    }                                                                                           \
  } while (0)

// Wrap a C-ABI entry point body: 0 on success, -1 on failure with b200w_last_error() set.
#define B200W_API_BEGIN try {
#define B200W_API_END                                \
  }                                                  \
  catch (const std::exception& e) {                  \
    ::b200w::set_last_error(e.what());               \
    return -1;                                       \
  }                                                  \
  catch (...) {                                      \
    ::b200w::set_last_error("unknown C++ exception"); \
    return -1;                                       \
  }                                                  \
  return 0;

int num_sms();
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// launch counter (bench.py reports gpu_launches from this)
extern unsigned long long g_launch_count;
#define B200W_COUNT_LAUNCH() (++::b200w::g_launch_count)

}  // namespace b200w
