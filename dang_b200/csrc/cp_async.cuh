// cp_async.cuh -- 16-byte asynchronous global -> shared copies (LDGSTS, L1-bypassing) for the per-thread prefetch
// rings of the streaming kernels (kernels_stream.cuh, kernels_cg_solve.cuh).  A thread only ever reads the slots it
// filled itself, so completion is tracked with cp.async.wait_group alone -- no block-level barrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
