// Round-2 probe, part 5 (not part of the product build): the MMA issue loop of the chain kernel with the variants as
// TEMPLATE parameters (no run-time flag branches in the loop, unlike chain_probe4): cycles per K chunk for
//   kTwo     2-MMA pattern (N = 32 + N = 16, alternating instruction descriptors) vs three N = 16 MMAs
//   kRing    A operand from a rotating 5-slot ring
//   kVec     loop state forced into vector registers (values re-read from shared memory each chunk -> R2UR per operand)
//   kWait    an mbarrier.try_wait on an already completed barrier + tcgen05.fence per chunk
//   kCommit  a tcgen05.commit per chunk
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/chain_probe5.cu -o build/chain_probe5
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../oprl_b200/csrc/ptx.cuh"
using namespace oprl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <bool kTwo, bool kRing, bool kVec, int kWait, bool kCommit>
__global__ void __launch_bounds__(128, 1) issue_loop(int chunks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done, dummy;
  __shared__ uint32_t slot_s;
  __shared__ volatile uint32_t state[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = tid; i < 40 * 1024 / 4; i += 128) f[i] = 0.001f * (i & 127);
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_init(&done, 1);
    ptx::mbar_init(&dummy, (1 << 20) - 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive(&bar);  // phase 0 complete: try_wait(parity 0) succeeds at once
  }
  if (warp == 0) ptx::tmem_alloc(&slot_s, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot_s;
  {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.01f * j;
    const uint32_t ta = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 192u;
    for (int s = 0; s < 5; ++s) {
      ptx::tmem_st32(ta + 64u * s, v);
      ptx::tmem_st32(ta + 64u * s + 32u, v);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      constexpr uint32_t pitch = 144, sbo = 9216, kstep = (2 * pitch) >> 4;
      const uint32_t idesc16 = ptx::idesc_tf32(128, 16, 0, 0), idesc32 = ptx::idesc_tf32(128, 32, 0, 0);
      const uint32_t dw_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
      const uint32_t dl0 = ((ptx::smem_u32(smem) >> 4) & 0x3FFFu) | ((pitch >> 4) << 16);
      const uint32_t bar_a = ptx::smem_u32(&bar), dummy_a = ptx::smem_u32(&dummy);
      uint32_t sl = 0, dl = dl0, in_group = 0, dgrp = tmem;
      if (kVec) { state[0] = sl; state[1] = dl; state[2] = dgrp; }
      t0 = clock64();
      for (int c = 0; c < chunks; ++c) {
        if (kVec) { sl = state[0]; dl = state[1]; dgrp = state[2]; }
        if (kWait & 1) {  // 1: try_wait, 2: tcgen05.fence::after_thread_sync, 4: only every second chunk
          if (!(kWait & 4) || !(c & 1)) {
            uint32_t ok;
            do {
              asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                           : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
            } while (!ok);
          }
        }
        if (kWait == 8) {   // wait as ONE asm block with its own loop (CUTLASS style)
          asm volatile(
              "{\n\t.reg .pred P1;\n\t"
              "WAIT_LOOP:\n\t"
              "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
              "@P1 bra WAIT_DONE;\n\t"
              "bra WAIT_LOOP;\n\t"
              "WAIT_DONE:\n\t}\n" ::"r"(bar_a), "r"(0u)
              : "memory");
        }
        if (kWait == 16) {  // poll result made warp-uniform through a vote over the active lanes
          uint32_t ok;
          do {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                         : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
          } while (!__any_sync(__activemask(), ok));
        }
        if (kWait == 32) {  // straight-line: a few predicated retries, no loop
          asm volatile(
              "{\n\t.reg .pred P1;\n\t"
              "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
              "@!P1 mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
              "@!P1 mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
              "@!P1 mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
              "@!P1 trap;\n\t}\n" ::"r"(bar_a), "r"(0u)
              : "memory");
        }
        if (kWait & 2) {
          if (!(kWait & 4) || !(c & 1)) ptx::tc_fence_after();
        }
        const uint32_t ta_hi = tmem + 192u + (kRing ? sl * 64u : 0u);
        const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (kTwo) {
            ptx::mma_tf32_ts2(dgrp, ta_hi + 8u * j, dl + j * kstep, dw_hi, idesc32, (in_group | j) ? 1u : 0u);
            ptx::mma_tf32_ts2(dgrp + 16u, ta_lo + 8u * j, dl + j * kstep, dw_hi, idesc16, 1u);
          } else {
            ptx::mma_tf32_ts2(dgrp, ta_lo + 8u * j, dl + j * kstep, dw_hi, idesc16, (c | j) ? 1u : 0u);
            ptx::mma_tf32_ts2(dgrp, ta_hi + 8u * j, dl + j * kstep + 2 * (sbo >> 4), dw_hi, idesc16, 1u);
            ptx::mma_tf32_ts2(dgrp + 16u, ta_hi + 8u * j, dl + j * kstep, dw_hi, idesc16, (in_group | j) ? 1u : 0u);
          }
        }
        if (kCommit) ptx::mma_commit_addr(dummy_a);
        dl += 4 * kstep;
        if ((c & 7) == 7) dl = dl0;
        if (++sl == 5) sl = 0;
        if (++in_group == 4) { in_group = 0; dgrp = (dgrp == tmem) ? tmem + 32u : tmem; }
        if (kVec) { state[0] = sl; state[1] = dl; state[2] = dgrp; }
      }
      t1 = clock64();
      ptx::mma_commit(&done);
    }
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (t0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <bool kTwo, bool kRing, bool kVec, int kWait, bool kCommit>
static void run(long long* d) {
  const int smem = 40 * 1024, chunks = 64;
  CK(cudaFuncSetAttribute(issue_loop<kTwo, kRing, kVec, kWait, kCommit>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long h[2] = {0, 0};
  for (int r = 0; r < 3; ++r) {
    issue_loop<kTwo, kRing, kVec, kWait, kCommit><<<1, 128, smem>>>(chunks, d);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  const int per = kTwo ? 8 : 12;
  printf("K %s %s %s %s %s: issue %.0f cyc / chunk, retired %.0f / chunk (%.1f / mma)\n", kTwo ? "2mma" : "3mma", kRing ? "ring" : "    ",
         kVec ? "vec" : "   ", kWait == 3 ? "wait+fence" : kWait == 1 ? "wait only " : kWait == 2 ? "fence only" : kWait == 7 ? "w+f / 2chk " : kWait == 8 ? "asm loop  " : kWait == 16 ? "vote loop " : kWait == 32 ? "straight  " : "          ", kCommit ? "commit" : "      ", h[0] / double(chunks), h[1] / double(chunks),
         h[1] / double(chunks * per));
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 64));
  run<false, false, false, 0, false>(d);
  run<false, true, false, 0, false>(d);
  run<true, false, false, 0, false>(d);
  run<true, true, false, 0, false>(d);
  run<true, true, true, 0, false>(d);
  run<true, true, false, 3, false>(d);
  run<true, true, false, 1, false>(d);
  run<true, true, false, 2, false>(d);
  run<true, true, false, 7, false>(d);
  run<true, true, false, 0, true>(d);
  run<true, true, false, 3, true>(d);
  run<true, true, true, 3, true>(d);
  run<true, true, true, 7, true>(d);
  run<false, true, true, 3, true>(d);
  run<false, true, false, 7, true>(d);
  run<true, true, false, 8, true>(d);
  run<true, true, false, 16, true>(d);
  run<true, true, false, 32, true>(d);
  return 0;
}
