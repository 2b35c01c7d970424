// Shared helpers for the nefii_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cuda_runtime.h>

#define NEFII_OK 0
#define NEFII_ERR_ARG (-1)
#define NEFII_ERR_CUDA (-2)
#define NEFII_ERR_STATE (-3)

namespace nefii {

// thread-local error text returned by nefii_last_error()
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define NEFII_CHECK_ARG(cond, ...)                                   \
  do {                                                               \
    if (!(cond)) return ::nefii::set_error(NEFII_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define NEFII_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return ::nefii::set_error(NEFII_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                cudaGetErrorString(e__), __FILE__, __LINE__);              \
  } while (0)

#define NEFII_LAUNCH_CHECK()                                                               \
  do {                                                                                     \
    ::nefii::count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return ::nefii::set_error(NEFII_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                                cudaGetErrorString(e__), __FILE__, __LINE__);              \
  } while (0)

// number of kernels this library has launched (nefii_launch_count); bumped by NEFII_LAUNCH_CHECK.  Replays of a captured
// trace graph add the kernels of one trip through each of its loops (tracer_graph.cu).
void count_launch();
long long launches();
void add_launches(long long n);

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200 (grid sizing default; launchers that depend on it read the device attribute)

}  // namespace nefii
