// Host-side helpers shared by the .cu translation units of libb200whisper.so
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
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

int num_sms();  // SM count of the CURRENT device (cached per device)
// Device memory for records that hold TMA descriptors (CUtensorMap) read by kernels from GLOBAL memory.  The address
// of such a record is never given out twice in the life of the process: the TMA unit caches descriptors by address,
// and a record freed by one engine and re-allocated (same address, other contents) by the next was fetched stale --
// an intermittent illegal memory access once engines had been destroyed and created (round 2: 2 of ~30 runs of the
// GPU suite).  Bump allocation from chunks that are never freed; ~0.5 KB per distinct decode GEMM call.
void* alloc_descriptor_record(size_t bytes);
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (device, function): remember what each device was given
// (grow-only) instead of a per-process flag, so engines on several GPUs of one process all get configured.
void ensure_dynamic_smem(const void* func, size_t bytes);
template <typename... KArgs>
inline void ensure_dynamic_smem(void (*kernel)(KArgs...), size_t bytes) {
  ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL).  Kernels of the decode step are launched with the
// programmatic-stream-serialization attribute: a kernel's CTAs may become resident (and run their
// prologue: barrier init, TMEM allocation, weight prefetch) while the kernel before it is still
// running.  Contract for every kernel launched through launch_pdl(): pdl_trigger() first, then
// pdl_wait() in every thread BEFORE the first access to anything another kernel wrote or reads --
// before any early return, so that grid completion always implies the predecessor's completion.
// Both are no-ops when the kernel is launched without the attribute (B200W_NO_PDL=1).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  B200W_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
#endif

// launch counter (bench.py reports gpu_launches from this); engines may run on several host threads
extern std::atomic<unsigned long long> g_launch_count;
#define B200W_COUNT_LAUNCH() (::b200w::g_launch_count.fetch_add(1, std::memory_order_relaxed))

}  // namespace b200w
