// Round-2 probe, part 2 (not part of the product build):
//   F  what one SM can pull out of L2 with 1-D bulk copies (16 KB chunks, 8-stage ring), for 1 / 32 / 64 / 148
//      CTAs streaming the same 512 KB (L2-resident), unicast and with .multicast::cluster in clusters of 2 / 4
//      (every CTA of the cluster issues 1/cs of the chunks and multicasts them to all)
//   G  one-way DSMEM hop with st.async (data + complete_tx on the remote mbarrier, no release round trip)
//   H  one-way DSMEM hop with one cp.async.bulk shared::cta -> shared::cluster copy
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/chain_probe2.cu -o build/chain_probe2
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

constexpr int kChunk = 16384;
constexpr int kStages = 8;

__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::
          "r"(ptx::smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(ptx::smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(caddr) : "memory");
}

// ------------------------------------------------------------------ F: ingest
__global__ void __launch_bounds__(64, 1) ingest_kernel(const float* src, int nchunks, int cs, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages], empty[kStages], lempty[kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cs > 1 ? ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], cs);
      ptx::mbar_init(&lempty[s], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (cs > 1) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  }
  const long long t0 = clock64();
  if (warp == 0) {
    // every CTA arms its own full barrier for every chunk; chunk c is issued by rank c % cs
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kStages;
      const uint32_t ph = (c / kStages) & 1;
      if (lane == 0) {
        if (static_cast<int>(rank) == c % cs) {
          ptx::mbar_wait(&empty[s], ph ^ 1);  // all CTAs of the cluster have drained this stage
        }
        ptx::mbar_wait(&lempty[s], ph ^ 1);  // this CTA's consumer has drained the stage: safe to re-arm
        ptx::mbar_expect_tx(&full[s], kChunk);
        if (static_cast<int>(rank) == c % cs) {
          const float* g = src + static_cast<size_t>(c % 32) * (kChunk / 4);
          if (cs > 1) bulk_g2s_mc(smem + s * kChunk, g, kChunk, &full[s], static_cast<uint16_t>((1u << cs) - 1));
          else ptx::bulk_g2s(smem + s * kChunk, g, kChunk, &full[s]);
        }
      }
      __syncwarp();
    }
  } else {
    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kStages;
      const uint32_t ph = (c / kStages) & 1;
      ptx::mbar_wait(&full[s], ph);
      acc += reinterpret_cast<const float*>(smem + s * kChunk)[lane];
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&lempty[s]);
        if (cs > 1) {
          // tell the CTA that will issue chunk c + kStages (same stage) that this CTA is done with it
          const int issuer = (c + kStages) % cs;
          mbar_arrive_remote(ptx::mapa(ptx::smem_u32(&empty[s]), issuer));
        } else {
          ptx::mbar_arrive(&empty[s]);
        }
      }
    }
    if (acc == 1.2345f) out[63] = 1;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (cs > 1) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  }
}

// ------------------------------------------------------------------ G/H: hops
__device__ __forceinline__ void st_async_v4(uint32_t caddr, float a, float b, float c, float d, uint32_t bar_caddr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(
                   caddr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar_caddr)
               : "memory");
}
// mode 0: st.async v4 by 128 threads; mode 1: one bulk s2s copy issued by thread 0 after a CTA barrier
__global__ void __launch_bounds__(128, 1) hop2_kernel(int mode, int bytes, int rounds, long long* out, float* sink) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4* rx = reinterpret_cast<float4*>(smem_raw);
  float4* tx = reinterpret_cast<float4*>(smem_raw + 65536);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const uint32_t me = ptx::cluster_ctarank(), peer = me ^ 1u;
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const int n4 = bytes / 16;
  const uint32_t peer_rx = ptx::mapa(ptx::smem_u32(rx), peer);
  const uint32_t peer_bar = ptx::mapa(ptx::smem_u32(&bar), peer);
  float acc = 0.f;
  // the receiver arms its barrier for the first hop before anything is sent
  if (tid == 0) ptx::mbar_expect_tx(&bar, bytes);
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const bool my_turn = ((r & 1) == static_cast<int>(me));
    if (my_turn) {
      const float x = static_cast<float>(r) + acc;
      if (mode == 0) {
        for (int i = tid; i < n4; i += 128) st_async_v4(peer_rx + i * 16, x, x + 1.f, x + 2.f, x + 3.f, peer_bar);
      } else {
        for (int i = tid; i < n4; i += 128) tx[i] = make_float4(x, x + 1.f, x + 2.f, x + 3.f);
        ptx::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) ptx::bulk_s2s_cluster(peer_rx, tx, bytes, peer_bar);
      }
    } else {
      ptx::mbar_wait_cluster(&bar, (r >> 1) & 1);
      for (int i = tid; i < n4; i += 128) acc += rx[i].x;
      __syncthreads();
      if (tid == 0) ptx::mbar_expect_tx(&bar, bytes);  // re-arm for my next receive (two hops later)
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

static void launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int block, int smem, int cs) {
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.numAttrs = 0;
  if (cs > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 8 * 256));
  float* src;
  CK(cudaMalloc(&src, 32 * kChunk));
  CK(cudaMemset(src, 0, 32 * kChunk));
  float* sink;
  CK(cudaMalloc(&sink, 1024));
  CK(cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kChunk));
  const int nchunks = 128;  // 2 MB per CTA
  for (int cs : {1, 2, 4}) {
    for (int grid : {cs, 32, 64, 148 / cs * cs}) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[1];
      launch_cfg(cfg, attr, grid, 64, kStages * kChunk, cs);
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, ingest_kernel, (const float*)src, nchunks, cs, d));
        CK(cudaDeviceSynchronize());
      }
      long long h[256];
      CK(cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("F  ingest: cluster %d grid %3d: %7lld cyc for %d KB per CTA = %.1f B/clk per SM landed (slowest CTA)\n", cs,
             grid, mx, nchunks * kChunk / 1024, static_cast<double>(nchunks) * kChunk / mx);
    }
  }
  CK(cudaFuncSetAttribute(hop2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  for (int mode = 0; mode < 2; ++mode)
    for (int bytes : {2048, 4096, 8192, 16384, 32768}) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[1];
      launch_cfg(cfg, attr, 2, 128, 131072, 2);
      const int rounds = 200;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, hop2_kernel, mode, bytes, rounds, d, sink));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      printf("%s %5d B/hop: %.0f cyc per one-way hop\n", mode == 0 ? "G st.async " : "H bulk s2s ", bytes,
             static_cast<double>(h) / rounds);
    }
  return 0;
}
