// Next-round experiment (not part of the product build): what does it cost to combine the fp32 partial
// tiles (128 x 32 = 16 KB each) of a split-K cluster?
//   mode 0  "gather":     ranks 1..ks-1 push their whole tile into rank 0 (what gemm.cuh does today)
//   mode 1  "all-to-all": every rank pushes, to each peer, the 32/ks columns that peer will finish
// Prints cycles from "all partial tiles ready" to "destination has everything", CTA 0's clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/dsmem_exchange_bench.cu -o build/dsmem_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../../oprl_b200/csrc/ptx.cuh"

using namespace oprl;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int kThreads = 256;
constexpr int kTileFloats = 128 * 32;

// smem: [0] own tile (16 KB) | [1..3] receive slots (16 KB each) | barrier
__global__ void __launch_bounds__(kThreads, 1) exchange_kernel(int mode, int ks, long long* out, float* sink) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* own = reinterpret_cast<float*>(smem_raw);
  float* slots = own + kTileFloats;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 4 * kTileFloats * 4);
  const int tid = threadIdx.x, lane = tid & 31;
  const int kr = static_cast<int>(ptx::cluster_ctarank());
  const int row = tid & 127, half = tid >> 7;
  if (tid == 0) {
    // arrivals expected by this CTA: one per remote warp that sends to it
    const int senders = (mode == 0) ? (kr == 0 ? ks - 1 : 0) : ks - 1;
    ptx::mbar_init(bar, senders > 0 ? senders * (kThreads / 32) : 1);
    ptx::fence_mbar_init();
  }
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = static_cast<float>(kr * 1000 + row + j + half * 16);
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const long long t0 = clock64();
  if (mode == 0) {
    if (kr > 0) {
      const uint32_t slot = ptx::mapa(ptx::smem_u32(slots + (kr - 1) * kTileFloats), 0);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        ptx::st_cluster_v4(slot + static_cast<uint32_t>(((half * 4 + j4) * 128 + row) * 16), v[4 * j4], v[4 * j4 + 1],
                           v[4 * j4 + 2], v[4 * j4 + 3]);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
    } else {
      ptx::mbar_wait_cluster(bar, 0);
      for (int r = 1; r < ks; ++r) {
        const float4* s = reinterpret_cast<const float4*>(slots + (r - 1) * kTileFloats);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 p = s[(half * 4 + j4) * 128 + row];
          v[4 * j4] += p.x; v[4 * j4 + 1] += p.y; v[4 * j4 + 2] += p.z; v[4 * j4 + 3] += p.w;
        }
      }
    }
  } else {
    // columns [c0, c0 + 32/ks) of the tile belong to rank c0 / (32/ks); this thread holds 16 columns
    const int per = 32 / ks;  // 16 (ks = 2) or 8 (ks = 4)
    for (int j0 = 0; j0 < 16; j0 += per) {
      const int owner = (half * 16 + j0) / per;
      if (owner == kr) continue;
      // slot index at the owner: which sender am I, counted without the owner itself
      const int si = kr < owner ? kr : kr - 1;
      const uint32_t slot = ptx::mapa(ptx::smem_u32(slots + si * kTileFloats), owner);
      for (int j4 = 0; j4 < per / 4; ++j4)
        ptx::st_cluster_v4(slot + static_cast<uint32_t>((j4 * 128 + row) * 16), v[j0 + 4 * j4], v[j0 + 4 * j4 + 1],
                           v[j0 + 4 * j4 + 2], v[j0 + 4 * j4 + 3]);
    }
    __syncwarp();
    if (lane == 0)
      for (int r = 0; r < ks; ++r)
        if (r != kr) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), r));
    ptx::mbar_wait_cluster(bar, 0);
    // (every warp arrives on every peer although only half of the warps hold columns of a given owner at
    //  ks = 2: the arrival count above assumes all 8 warps of every sender arrive)
    for (int s = 0; s < ks - 1; ++s) {
      const float4* sp = reinterpret_cast<const float4*>(slots + s * kTileFloats);
      for (int j4 = 0; j4 < per / 4; ++j4) {
        const float4 p = sp[j4 * 128 + row];
        v[0] += p.x + p.y + p.z + p.w;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc += v[j];
  sink[blockIdx.x * kThreads + tid] = acc;
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
  // nobody leaves while a peer may still write into this CTA's slots
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

int main() {
  const int smem = 4 * kTileFloats * 4 + 64;
  CK(cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d_out;
  float* d_sink;
  CK(cudaMalloc(&d_out, 8));
  CK(cudaMalloc(&d_sink, 64 * kThreads * 4));
  for (int mode = 0; mode < 2; ++mode)
    for (int ks = 2; ks <= 4; ks *= 2) {
      long long best = 1 << 30;
      for (int it = 0; it < 20; ++it) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(16 * ks);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ks;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, exchange_kernel, mode, ks, d_out, d_sink));
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
        if (c < best) best = c;
      }
      printf("%s ks=%d: %lld cycles (best of 20, rank 0 of cluster 0)\n", mode ? "all-to-all" : "gather    ", ks, best);
    }
  return 0;
}
