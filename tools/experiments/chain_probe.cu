// Round-2 probe (not part of the product build): the constants the batch-slice "chain" kernel is
// designed around.
//   A  cycles per tcgen05.mma.kind::tf32 (M = 128, A operand in TMEM, B in smem) for N = 16..128,
//      issued back to back by one elected lane with pre-built descriptors (96 MMAs = one 256-deep layer, 3xTF32)
//   B  one-way hop between the two CTAs of a cluster through distributed shared memory
//      (st.shared::cluster.v4 by 256 threads + one remote mbarrier arrive per warp), 1..32 KB
//   C  the same payload through global memory / L2 (st.global.v4, remote mbarrier arrive, ld.global.cg)
//   D  barrier.cluster arrive + wait, cluster of 2 / 4 / 8
//   E  launch floor: a chain of 24 near-empty kernels in a CUDA graph with programmatic dependent launch
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/experiments/chain_probe.cu -o build/chain_probe
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

// ------------------------------------------------------------------ A: MMA issue rate
template <int N, bool kTS>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* f = reinterpret_cast<float*>(smem);
  // A (SS mode): 128 x 32 floats; B: 8 chunks of N x 32 floats
  for (int i = threadIdx.x; i < 128 * 32 + 8 * N * 32; i += 128) f[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.01f * j;
    const uint32_t ta = tmem + ((threadIdx.x & ~31u) << 16) + 256u;
    ptx::tmem_st32(ta, v);
    ptx::tmem_st32(ta + 32, v);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t sa = ptx::smem_u32(smem);
    const uint32_t sb = sa + 128 * 32 * 4;
    const uint32_t idesc = ptx::idesc_tf32(128, N, 0, 0);
    const uint64_t da0 = ptx::smem_desc(sa, 128, 1024);
    const uint64_t db0 = ptx::smem_desc(sb, 128, 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const uint64_t db = db0 + static_cast<uint64_t>((c * N * 32 * 4) >> 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dbj = db + static_cast<uint64_t>((j * 256) >> 4);
          if (kTS) {
            ptx::mma_tf32_ts(tmem, tmem + 256u + 32u + 8u * j, dbj, idesc, (c | j) ? 1u : 0u);
            ptx::mma_tf32_ts(tmem, tmem + 256u + 8u * j, dbj, idesc, 1u);
            ptx::mma_tf32_ts(tmem + 128u, tmem + 256u + 8u * j, dbj, idesc, (c | j) ? 1u : 0u);
          } else {
            const uint64_t daj = da0 + static_cast<uint64_t>((j * 256) >> 4);
            ptx::mma_tf32(tmem, daj, dbj, idesc, (c | j) ? 1u : 0u);
            ptx::mma_tf32(tmem, daj, dbj, idesc, 1u);
            ptx::mma_tf32(tmem + 128u, daj, dbj, idesc, (c | j) ? 1u : 0u);
          }
        }
      }
      t1 = clock64();
      ptx::mma_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    t2 = clock64();
    if (t0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

template <int N, bool kTS>
static void run_mma(long long* d) {
  const int smem = (128 * 32 + 8 * N * 32) * 4;
  CK(cudaFuncSetAttribute(mma_rate<N, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long h[2] = {0, 0};
  for (int r = 0; r < 3; ++r) {
    mma_rate<N, kTS><<<1, 128, smem>>>(d);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("A  N=%3d %s: 96 MMAs issue %5lld cyc (%.1f/mma), retired %5lld cyc (%.1f/mma)\n", N, kTS ? "TS" : "SS", h[0],
         h[0] / 96.0, h[1], h[1] / 96.0);
}

// ------------------------------------------------------------------ B/C: cluster hop
// mode 0: DSMEM stores; mode 1: through global memory.  `bytes` per hop, ping-pong `rounds` times.
__global__ void __launch_bounds__(256, 1) hop_kernel(int mode, int bytes, int rounds, float* gbuf, long long* out,
                                                     float* sink) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float4* rx = reinterpret_cast<float4*>(smem_raw);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t me = ptx::cluster_ctarank(), peer = me ^ 1u;
  if (tid == 0) {
    ptx::mbar_init(&bar, 8);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const int n4 = bytes / 16;  // float4 per hop
  const uint32_t peer_rx = ptx::mapa(ptx::smem_u32(rx), peer);
  const uint32_t peer_bar = ptx::mapa(ptx::smem_u32(&bar), peer);
  float4* g_tx = reinterpret_cast<float4*>(gbuf) + static_cast<size_t>(blockIdx.x) * 4096;
  const float4* g_rx = reinterpret_cast<const float4*>(gbuf) + static_cast<size_t>(blockIdx.x ^ 1u) * 4096;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const bool my_turn = ((r & 1) == static_cast<int>(me));
    if (my_turn) {
      const float x = static_cast<float>(r) + acc;
      if (mode == 0) {
        for (int i = tid; i < n4; i += 256) ptx::st_cluster_v4(peer_rx + i * 16, x, x + 1.f, x + 2.f, x + 3.f);
      } else {
        for (int i = tid; i < n4; i += 256) g_tx[i] = make_float4(x, x + 1.f, x + 2.f, x + 3.f);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(peer_bar);
    } else {
      ptx::mbar_wait_cluster(&bar, (r >> 1) & 1);
      if (mode == 0) {
        for (int i = tid; i < n4; i += 256) acc += rx[i].x;
      } else {
        for (int i = tid; i < n4; i += 256) {
          const float4 v = __ldcg(g_rx + i);
          rx[i] = v;
          acc += v.x;
        }
      }
      __syncthreads();
    }
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

// ------------------------------------------------------------------ D: cluster barrier
__global__ void __launch_bounds__(256, 1) cbar_kernel(int rounds, long long* out) {
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// ------------------------------------------------------------------ E: launch floor
__global__ void __launch_bounds__(256, 1) empty_kernel(float* p) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  if (threadIdx.x == 0) p[blockIdx.x] += 1.f;
}

static void launch_cluster(void (*k)(float*), int grid, int cluster, bool pdl, cudaStream_t st, float* p) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CK(cudaLaunchKernelEx(&cfg, k, p));
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 64));
  float* sink;
  CK(cudaMalloc(&sink, 1 << 20));
  CK(cudaMemset(sink, 0, 1 << 20));
  run_mma<16, true>(d);
  run_mma<32, true>(d);
  run_mma<64, true>(d);
  run_mma<128, true>(d);
  run_mma<32, false>(d);
  run_mma<64, false>(d);

  CK(cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int mode = 0; mode < 2; ++mode)
    for (int bytes : {16 * 256, 8192, 16384, 32768, 65536}) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = 65536;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const int rounds = 200;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, hop_kernel, mode, bytes, rounds, sink + 1024, d, sink));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      printf("%s  %5d B/hop: %.0f cyc per one-way hop (%.1f B/clk)\n", mode == 0 ? "B dsmem " : "C global", bytes,
             static_cast<double>(h) / rounds, bytes / (static_cast<double>(h) / rounds));
    }
  for (int cs : {2, 4, 8}) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(256);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaLaunchKernelEx(&cfg, cbar_kernel, 200, d));
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("D  barrier.cluster arrive+wait, cluster of %d: %.0f cyc\n", cs, h / 200.0);
  }
  // E: launch floor in a graph
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (int cluster : {1, 2, 4}) {
    for (int pdl = 0; pdl < 2; ++pdl) {
      for (int grid : {32, 128}) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < 24; ++i) launch_cluster(empty_kernel, grid, cluster, pdl != 0, st, sink);
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e0, st));
        for (int i = 0; i < 200; ++i) CK(cudaGraphLaunch(ge, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("E  empty kernel chain: cluster %d grid %3d pdl %d: %.2f us per launch\n", cluster, grid, pdl,
               ms * 1e3 / (200 * 24));
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
      }
    }
  }
  return 0;
}
