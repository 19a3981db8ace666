// Co-residency experiment (tools/exp_coresident.py): G CTAs that hold `smem_bytes` of shared memory each and spin for
// `ns` nanoseconds -- a stand-in for a communication kernel running next to the persistent kernels.
#include <cstdint>
#include <cuda_runtime.h>
__global__ void spin_kernel(long long ns) {
  extern __shared__ unsigned char s[];
  if (threadIdx.x == 0) s[0] = 1;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while ((long long)(t - t0) < ns);
}
extern "C" int spin_launch(int ctas, int threads, int smem_bytes, long long ns, void* stream) {
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  spin_kernel<<<ctas, threads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(ns);
  return (int)cudaGetLastError();
}
