// Kernel launch helper.  Every kernel of the step is launched with programmatic stream serialisation (PDL, sm_90+):
// each kernel begins with pdl_sync() = griddepcontrol.wait (all memory of the preceding grid is visible after it)
// followed by griddepcontrol.launch_dependents, so the NEXT kernel's CTAs are scheduled and run their prologue while
// this grid drains instead of paying a launch gap per kernel (~410 launches per ResNet-50 step).  Semantics are those
// of plain stream order: nothing before the wait touches global memory.  R3M_PDL=0 launches without the attribute.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace r3m {

bool pdl_enabled();
// > 0: launch without the PDL attribute (Engine::run sets it around the cross-stream join points of a stream capture)
extern thread_local int g_pdl_suppress;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = (pdl_enabled() && g_pdl_suppress == 0) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Cooperative launch (all blocks of the grid are resident together: required by kernels with a software grid barrier,
// so that concurrent work on other streams can never leave part of the grid unscheduled while the rest spins).  No PDL.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeCooperative;
  attr.val.cooperative = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace r3m
