// Round-2 probe, part 4 (not part of the product build): what one tcgen05.mma of the chain kernel costs, factor by factor.
// One CTA; warp 0 (and optionally warp 1) issue `iters` rounds of the chain kernel's MMA pattern for one K chunk
// (4 K steps) with everything else stripped; flags switch the suspects on one by one:
//   bit 0  A operand from a rotating ring of TMEM slots instead of the same 64 columns
//   bit 1  B operand buffer with the 144-byte core pitch / 9216-byte row-group stride instead of 128 / 1024
//   bit 2  the 2-MMA pattern (N = 32 against [hi ; lo] + N = 16) instead of three N = 16 MMAs per K step
//   bit 3  eight other warps store to the A ring (tcgen05.st x32, back to back) while the MMAs run
//   bit 4  one tcgen05.commit per chunk (to a dummy mbarrier)
//   bit 5  two issuing warps (M tile 0 / 1: different accumulators)
//   bit 6  B descriptor does not advance (same 8 K columns every step)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/chain_probe4.cu -o build/chain_probe4
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../../oprl_b200/csrc/ptx.cuh"

using namespace oprl;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void __launch_bounds__(320, 1) mma_cost(int flags, int chunks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2], dummy;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = tid; i < 40 * 1024 / 4; i += 320) f[i] = 0.001f * (i & 127);
  if (tid == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::mbar_init(&dummy, (1 << 20) - 1);
    stop = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int pitch = (flags & 2) ? 144 : 128;
  const int sbo = (flags & 2) ? 9216 : 1024;
  const int n_slots = 5;
  const uint32_t a_col0 = 192;
  if (warp >= 2) {
    // fill the A ring once (every lane quarter), then optionally keep storing
    const int q = warp & 3;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.01f * j;
    const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + a_col0;
    for (int s = 0; s < n_slots; ++s) {
      ptx::tmem_st32(ta + 64u * s, v);
      ptx::tmem_st32(ta + 64u * s + 32u, v);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    asm volatile("bar.sync 1, 320;\n" ::: "memory");
    if (flags & 8) {
      int s = warp >= 6 ? 2 : 0;
      while (!stop) {
        ptx::tmem_st32(ta + 64u * s, v);
        ptx::tmem_st32(ta + 64u * s + 32u, v);
        ptx::tmem_st_wait();
        s = (s + 1) % n_slots;
      }
    }
  } else {
    asm volatile("bar.sync 1, 320;\n" ::: "memory");
    ptx::tc_fence_after();
    const int m = warp;
    if (m == 0 || (flags & 32)) {
      const uint32_t idesc16 = ptx::idesc_tf32(128, 16, 0, 0), idesc32 = ptx::idesc_tf32(128, 32, 0, 0);
      const uint32_t kstep = static_cast<uint32_t>(2 * pitch) >> 4;
      const uint32_t dw_hi = ((static_cast<uint32_t>(sbo) >> 4) & 0x3FFFu) | (1u << 14);
      const uint32_t lbo_bits = (static_cast<uint32_t>(pitch) >> 4) << 16;
      const uint32_t dl0 = ((ptx::smem_u32(smem) >> 4) & 0x3FFFu) | lbo_bits;
      const uint32_t d0 = tmem + static_cast<uint32_t>(m * 96);
      uint32_t sl = 0;
      uint32_t dl = dl0;
      long long t0 = clock64();
      for (int c = 0; c < chunks; ++c) {
        if (ptx::elect_one()) {
          const uint32_t ta_hi = tmem + a_col0 + ((flags & 1) ? sl * 64u : 0u);
          const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t d = (flags & 64) ? dl0 : dl + j * kstep;
            if (flags & 4) {
              ptx::mma_tf32_ts2(d0, ta_hi + 8u * j, d, dw_hi, idesc32, (c | j) ? 1u : 0u);
              ptx::mma_tf32_ts2(d0 + 16u, ta_lo + 8u * j, d, dw_hi, idesc16, 1u);
            } else {
              ptx::mma_tf32_ts2(d0, ta_lo + 8u * j, d, dw_hi, idesc16, (c | j) ? 1u : 0u);
              ptx::mma_tf32_ts2(d0, ta_hi + 8u * j, d + 2 * (sbo >> 4), dw_hi, idesc16, 1u);
              ptx::mma_tf32_ts2(d0 + 16u, ta_hi + 8u * j, d, dw_hi, idesc16, (c | j) ? 1u : 0u);
            }
          }
          if (flags & 16) ptx::mma_commit(&dummy);
        }
        __syncwarp();
        dl += 4 * kstep;
        if ((c & 7) == 7) dl = dl0;
        if (++sl == n_slots) sl = 0;
      }
      long long t1 = clock64();
      if (ptx::elect_one()) ptx::mma_commit(&bar[m]);
      __syncwarp();
      ptx::mbar_wait(&bar[m], 0);
      long long t2 = clock64();
      if (lane == 0) {
        out[2 * m] = t1 - t0;
        out[2 * m + 1] = t2 - t0;
      }
    }
    if (warp == 0) {
      if ((flags & 32)) {
        // wait for warp 1 too before stopping the store warps (its barrier is bar[1])
        ptx::mbar_wait(&bar[1], 0);
      }
      stop = 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 64));
  const int smem = 40 * 1024;
  CK(cudaFuncSetAttribute(mma_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int chunks = 64;
  const int cases[] = {0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 23, 31, 32, 39, 47, 63, 64, 64 + 7, 64 + 15};
  for (int flags : cases) {
    long long h[4] = {0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) {
      CK(cudaMemset(d, 0, 64));
      mma_cost<<<1, 320, smem>>>(flags, chunks, d);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
    const int per = (flags & 4) ? 8 : 12;
    printf("J flags %3d [%s%s%s%s%s%s%s]: %d chunks: issue %6lld cyc (%.0f / chunk), retired %6lld cyc (%.0f / chunk, %.1f / mma)", flags,
           flags & 1 ? "ring " : "", flags & 2 ? "pitch144 " : "", flags & 4 ? "2mma " : "3mma ", flags & 8 ? "sttm " : "",
           flags & 16 ? "commit " : "", flags & 32 ? "2warps " : "", flags & 64 ? "fixedB " : "", chunks, h[0], h[0] / double(chunks), h[1],
           h[1] / double(chunks), h[1] / double(chunks * per));
    if (flags & 32) printf("  | warp 1: retired %lld (%.0f / chunk)", h[3], h[3] / double(chunks));
    printf("\n");
  }
  return 0;
}
