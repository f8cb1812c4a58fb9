// Internals shared by the kernel translation units (sketch.cu, cluster.cu, mlpath.cu, genotype.cu).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "kernels.cuh"

namespace drprg {

std::atomic<uint64_t>& launch_counter();  // kernels launched by this library since load (bench.py's gpu_launches)
#define g_launches (launch_counter())


// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember what every (kernel, device)
// pair was raised to, so a second index on another GPU of the same process (multi-GPU sharding inside the library) gets
// its own call.  Thread-safe; a failure is reported instead of surfacing later as an invalid-value launch error.
template <class F>
static void ensure_dyn_smem(F kernel, size_t bytes) {
    static std::mutex m;
    static std::map<std::pair<const void*, int>, size_t> configured;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    size_t& have = configured[{reinterpret_cast<const void*>(kernel), dev}];
    if (bytes <= have) return;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ") + cudaGetErrorString(e));
    have = bytes;
}

#define FULL 0xffffffffu
__device__ __forceinline__ uint32_t cov_sat(int32_t c) { return c > 65535 ? 65535u : (uint32_t)c; }  // uint16 upstream

}  // namespace drprg
