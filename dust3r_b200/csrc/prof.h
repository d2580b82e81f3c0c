// Launch accounting + optional CUDA-event instrumentation of every kernel launch of the library.
//   d3r_launch_count()      : number of kernels launched by this library since the last reset
//   d3r_prof_enable(1)      : bracket every launch with cudaEvents on the launching stream
//   d3r_prof_report(buf,n)  : JSON {tag: {count, ms, flops, bytes}} (synchronises the recorded events)
//   d3r_prof_dump(buf,n)    : JSON list of every recorded launch [{tag, detail, ms, flops, bytes}] in launch order
#pragma once
#include <cuda_runtime.h>

namespace d3r {
namespace prof {
struct Scope {
  Scope(const char* tag, cudaStream_t st, double flops = 0.0, double bytes = 0.0, int launches = 1, const char* detail = nullptr);
  ~Scope();
  int idx;
  cudaStream_t st;
};
}  // namespace prof
}  // namespace d3r
