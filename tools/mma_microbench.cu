// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, SS mode, no-swizzle K-major
// operands resident in smem) as a function of N.  Decides the N-tile of gemm.cuh.
#include <cstdio>
#include <cstdlib>
#include "../oprl_b200/csrc/ptx.cuh"
using namespace oprl;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_loop(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < (128 * 32 + N * 32); i += 128) f[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t sa = ptx::smem_u32(smem);
    const uint32_t sb = sa + 128 * 32 * 4;
    const uint32_t idesc = ptx::idesc_tf32(128, N, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ptx::mma_tf32(slot, ptx::smem_desc(sa + j * 256, 128, 1024),
                      ptx::smem_desc(sb + j * 256, 128, 1024), idesc, 1u);
      }
    }
    long long t1 = clock64();
    ptx::mma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(slot, 256);
  }
}

template <int N>
void run() {
  long long* d;
  cudaMalloc(&d, 16);
  int smem = (128 * 32 + N * 32) * 4;
  cudaFuncSetAttribute(mma_loop<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 256;
  for (int r = 0; r < 2; ++r) mma_loop<N><<<1, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", N, (double)h[0] / (iters * 4),
         (double)h[1] / (iters * 4), cudaGetErrorString(e));
}

int main() {
  run<16>(); run<32>(); run<64>(); run<128>(); run<256>();
  return 0;
}
