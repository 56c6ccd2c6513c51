// Internal: launch helper with optional programmatic dependent launch (PDL).
//
// A DGN step at the headline size is ~100 dependent kernels of a few microseconds each, so overlapping the launch
// path of kernel N+1 with the execution of kernel N looks attractive.  Every kernel of this library starts with
// griddepcontrol.wait (returns once every prerequisite grid has completed and its memory is visible; a no-op for a
// plain launch) followed by griddepcontrol.launch_dependents, and nothing is read or written before the wait, so
// the data flow is exactly that of stream order either way.  With DGN_PDL=1 the kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (works under stream capture: the CUDA graph gets programmatic
// edges).  MEASURED on B200 (profiles/README.md): back-to-back launches of one kernel gain ~5 %, but the whole
// captured training step gets SLOWER (0.68 -> 0.80 ms), with the trigger before or after the wait - pre-launched
// grids hold SM resources the running grid and the parallel weight-gradient branch need.  Hence off by default.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>

namespace dgn {

// Wait first, then trigger: the dependent grid is released only once THIS grid really runs, so at most one
// not-yet-runnable grid sits on the SMs.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DGN_PDL"); return e && atoi(e) != 0; }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace dgn
