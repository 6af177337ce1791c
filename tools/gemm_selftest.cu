// Standalone GPU self-test for the tcgen05 3xTF32 GEMM (oprl_b200/csrc/gemm.cuh).
// Checks the tensor-core and FFMA variants against an fp64 host reference
// (row-major, tiled, transposed-tiled and column-sum outputs), prints an
// in-kernel phase breakdown, and times stream / CUDA-graph launch chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/gemm_selftest.cu -o build/gemm_selftest
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../oprl_b200/csrc/gemm.cuh"

using namespace oprl;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

struct HostMat {  // logical [rows x cols] row-major + tiled device copies
  int rows, cols;
  std::vector<float> v;
  float* d = nullptr;
  void upload() {
    std::vector<float> tiled((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) tiled[ct_index(rows, r, c)] = v[(size_t)r * cols + c];
    CK(cudaMalloc(&d, tiled.size() * 4));
    CK(cudaMemcpy(d, tiled.data(), tiled.size() * 4, cudaMemcpyHostToDevice));
  }
};

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int g_ksplit = 1;  // CTAs per tile (cluster size) of the launches below; 0 = the engine's automatic choice

// host-side description of one launch (the kernel takes the op table by device pointer)
struct HostLaunch {
  int n_ops;
  long long* prof;
  GemmOp op[kMaxOps];
};

template <bool kSimt>
static void launch(const HostLaunch& Lin, cudaStream_t st = 0) {
  // device copy of the (finalized) op table, re-uploaded only when the host description changed
  struct Cached { HostLaunch h; GemmOp* d; };
  static std::vector<Cached> cache;
  Cached* c = nullptr;
  for (auto& x : cache)
    if (memcmp(&x.h, &Lin, sizeof(HostLaunch)) == 0) c = &x;
  if (!c) {
    Cached n;
    n.h = Lin;
    HostLaunch fin = Lin;
    for (int i = 0; i < fin.n_ops; ++i) gemm_finalize(fin.op[i]);
    CK(cudaMalloc(&n.d, sizeof(GemmOp) * fin.n_ops));
    CK(cudaMemcpy(n.d, fin.op, sizeof(GemmOp) * fin.n_ops, cudaMemcpyHostToDevice));
    cache.push_back(n);
    c = &cache.back();
  }
  GemmLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_ops = Lin.n_ops;
  L.prof = Lin.prof;
  L.ops = c->d;
  int tiles = 0;
  for (int i = 0; i < L.n_ops; ++i) {
    tiles += gemm_tiles(Lin.op[i]);
    L.tile_end[i] = tiles;
  }
  const int ks = g_ksplit > 0 ? g_ksplit : gemm_choose_ksplit(Lin.op, Lin.n_ops, 148);
  L.ksplit = ks;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(tiles * ks);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ks;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ks > 1 ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, gemm_kernel<kSimt>, L));
}

static double run_case(int M, int N, int K, bool simt, int passes) {
  HostMat A, B;
  A.rows = M; A.cols = K; B.rows = N; B.cols = K;
  A.v.resize((size_t)M * K);
  B.v.resize((size_t)N * K);
  for (auto& x : A.v) x = frand();
  for (auto& x : B.v) x = frand();
  A.upload();
  B.upload();
  std::vector<float> bias(N);
  for (auto& x : bias) x = frand();
  float *d_bias, *d_rm, *d_t, *d_tt, *d_cs;
  CK(cudaMalloc(&d_bias, N * 4));
  CK(cudaMemcpy(d_bias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_rm, (size_t)M * N * 4));
  CK(cudaMalloc(&d_t, (size_t)M * N * 4));
  CK(cudaMalloc(&d_tt, (size_t)M * N * 4));
  CK(cudaMalloc(&d_cs, (size_t)(M / 128) * N * 4));
  CK(cudaMemset(d_rm, 0xff, (size_t)M * N * 4));

  HostLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_ops = 1;
  GemmOp& o = L.op[0];
  o.a = A.d; o.a_rows = M;
  o.b = B.d; o.b_rows = N;
  o.M = M; o.N = N; o.K = K;
  o.bias = d_bias; o.bias_n = N; o.act = ACT_RELU;
  o.t = d_t; o.t_rows = M; o.t_c0 = 0; o.t_n = N;
  o.tt = d_tt; o.tt_rows = N;
  o.rm = d_rm; o.rm_ld = N; o.rm_m = M; o.rm_n = N;
  o.colsum = d_cs; o.colsum_ld = N;
  o.passes = passes; o.alpha = 1.f;
  // fused layer-0 weight gradient: dW0[n][k] = sum_m D(m, n) X(m, k), X tiled [M x 64]
  const int KP0 = 64, A_ = 6, A4_ = 8, S_ = 50;  // columns [action 6 | pad 2 | state 50 | pad 6]
  HostMat X;
  X.rows = M; X.cols = KP0;
  X.v.resize((size_t)M * KP0);
  for (auto& x : X.v) x = frand();
  X.upload();
  float *d_part, *d_dw0, *d_csout;
  unsigned int* d_cnt;
  CK(cudaMalloc(&d_part, (size_t)(M / 128) * N * KP0 * 4));
  CK(cudaMalloc(&d_dw0, (size_t)N * (A_ + S_) * 4));
  CK(cudaMalloc(&d_csout, (size_t)N * 4));
  CK(cudaMalloc(&d_cnt, (N / 32) * 4));
  CK(cudaMemset(d_cnt, 0, (N / 32) * 4));
  CK(cudaMemset(d_dw0, 0xff, (size_t)N * (A_ + S_) * 4));
  o.dw0_x = X.d; o.dw0_kp = KP0; o.dw0_part = d_part; o.dw0_out = d_dw0; o.dw0_cnt = d_cnt;
  o.dw0_ld = A_ + S_; o.dw0_n = N - 3; o.dw0_cols = KP0;
  o.dw0_map_a = A_; o.dw0_map_a4 = A4_; o.dw0_map_s = S_;
  o.colsum_out = d_csout; o.colsum_n = N;
  float* d_bout;
  CK(cudaMalloc(&d_bout, (size_t)N * 4));
  o.dw0_ones = A_; o.dw0_bias_out = d_bout;
  if (simt) launch<true>(L); else launch<false>(L);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  size_t mn = (size_t)M * N;
  std::vector<float> out(mn), tv(mn), ttv(mn), cs((size_t)(M / 128) * N);
  CK(cudaMemcpy(out.data(), d_rm, mn * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tv.data(), d_t, mn * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ttv.data(), d_tt, mn * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cs.data(), d_cs, cs.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_terr = 0, max_tterr = 0, max_cerr = 0;
  std::vector<double> colref((size_t)(M / 128) * N, 0.0);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0, mag = 0;
      for (int k = 0; k < K; ++k) {
        double a = A.v[(size_t)m * K + k];
        double b = B.v[(size_t)n * K + k];
        acc += a * b;
        mag += fabs(a * b);
      }
      acc += bias[n];
      if (acc < 0) acc = 0;
      double got = out[(size_t)m * N + n];
      double err = fabs(got - acc) / (mag + 1.0);
      if (!(err <= max_err)) max_err = err;  // NaN-propagating
      double tgot = (double)tv[ct_index(M, m, n)];
      double terr = fabs(tgot - got) / (fabs(got) + 1e-3);
      if (!(terr <= max_terr)) max_terr = terr;
      double ttgot = (double)ttv[ct_index(N, n, m)];
      double tterr = fabs(ttgot - got) / (fabs(got) + 1e-3);
      if (!(tterr <= max_tterr)) max_tterr = tterr;
      colref[(size_t)(m / 128) * N + n] += got;
    }
  for (size_t i = 0; i < cs.size(); ++i) {
    double ce = fabs(cs[i] - colref[i]) / (fabs(colref[i]) + 1.0);
    if (!(ce <= max_cerr)) max_cerr = ce;
  }
  // dW0 and the cross-tile column sums against the device's own D (row-major output)
  double max_dwerr = 0;
  {
    std::vector<float> dw((size_t)N * (A_ + S_)), cso(N);
    CK(cudaMemcpy(dw.data(), d_dw0, dw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cso.data(), d_csout, cso.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> bo(N);
    CK(cudaMemcpy(bo.data(), d_bout, bo.size() * 4, cudaMemcpyDeviceToHost));
    for (int n = 0; n < N - 3; ++n) {  // the ones-column bias gradient against the column sums
      double ce = fabs(bo[n] - cso[n]) / (fabs(cso[n]) + 1.0);
      if (!(ce <= max_cerr)) max_cerr = ce;
    }
    for (int n = 0; n < N; ++n) {
      double ctot = 0;
      for (int m = 0; m < M; ++m) ctot += out[(size_t)m * N + n];
      double ce = fabs(cso[n] - ctot) / (fabs(ctot) + 1.0);
      if (!(ce <= max_cerr)) max_cerr = ce;
      for (int col = 0; col < A_ + S_; ++col) {
        const int kk = col < S_ ? A4_ + col : col - S_;  // reference column -> tiled column
        double acc = 0, mag = 0;
        for (int m = 0; m < M; ++m) {
          double d = out[(size_t)m * N + n], x = X.v[(size_t)m * KP0 + kk];
          acc += d * x;
          mag += fabs(d * x);
        }
        const float got = dw[(size_t)n * (A_ + S_) + col];
        double err;
        if (n >= N - 3) {
          uint32_t bits;
          memcpy(&bits, &got, 4);
          err = bits == 0xffffffffu ? 0 : 1;  // rows >= dw0_n must stay untouched
        } else {
          err = fabs(got - acc) / (mag + 1.0);
        }
        if (!(err <= max_dwerr)) max_dwerr = err;
      }
    }
  }
  printf("  M=%d N=%d K=%d %s passes=%d: rel_err=%.3e tiled=%.3e ttiled=%.3e colsum=%.3e dw0=%.3e\n", M, N, K,
         simt ? "SIMT" : "TC  ", passes, max_err, max_terr, max_tterr, max_cerr, max_dwerr);
  cudaFree(X.d); cudaFree(d_part); cudaFree(d_dw0); cudaFree(d_csout); cudaFree(d_cnt);
  cudaFree(A.d); cudaFree(B.d);
  cudaFree(d_bias); cudaFree(d_rm); cudaFree(d_t); cudaFree(d_tt); cudaFree(d_cs);
  double worst = max_err;
  if (!(max_terr < 1e-6)) worst = 1;
  if (!(max_tterr < 1e-6)) worst = 1;
  if (!(max_cerr < 1e-5)) worst = 1;
  if (!(max_dwerr < 2e-6)) worst = 1;
  return worst;
}

static HostLaunch make_bench(int nops, int M, int N, int K, int passes, bool tt, bool dw0 = false) {
  HostLaunch L;
  memset(&L, 0, sizeof(L));
  L.n_ops = nops;
  for (int i = 0; i < nops; ++i) {
    float *a, *b, *t, *ttp;
    CK(cudaMalloc(&a, (size_t)M * K * 4));
    CK(cudaMalloc(&b, (size_t)N * K * 4));
    CK(cudaMalloc(&t, (size_t)M * N * 4));
    CK(cudaMalloc(&ttp, (size_t)M * N * 4));
    CK(cudaMemset(a, 0, (size_t)M * K * 4));
    CK(cudaMemset(b, 0, (size_t)N * K * 4));
    GemmOp& o = L.op[i];
    o.a = a; o.a_rows = M; o.b = b; o.b_rows = N;
    o.M = M; o.N = N; o.K = K; o.act = ACT_RELU;
    o.t = t; o.t_rows = M; o.t_n = N; o.passes = passes; o.alpha = 1.f;
    if (tt) { o.tt = ttp; o.tt_rows = N; }
    if (dw0 && i == 0) {
      float *x, *part, *out, *cs, *cso;
      unsigned int* cnt;
      CK(cudaMalloc(&x, (size_t)M * 32 * 4));
      CK(cudaMemset(x, 0, (size_t)M * 32 * 4));
      CK(cudaMalloc(&part, (size_t)(M / 128) * N * 32 * 4));
      CK(cudaMalloc(&out, (size_t)N * 32 * 4));
      CK(cudaMalloc(&cs, (size_t)(M / 128) * N * 4));
      CK(cudaMalloc(&cso, (size_t)N * 4));
      CK(cudaMalloc(&cnt, (N / 32) * 4));
      CK(cudaMemset(cnt, 0, (N / 32) * 4));
      o.dw0_x = x; o.dw0_kp = 32; o.dw0_part = part; o.dw0_out = out; o.dw0_cnt = cnt;
      o.dw0_ld = 30; o.dw0_n = N; o.dw0_cols = 32; o.dw0_map_a = 6; o.dw0_map_a4 = 8; o.dw0_map_s = 24;
      o.dw0_ones = 6; o.dw0_bias_out = cso;
    }
  }
  return L;
}

static void spin_warm(const HostLaunch& L, double ms_target) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  float ms = 0;
  while (ms < ms_target) {
    for (int i = 0; i < 200; ++i) launch<false>(L);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
}

static void bench(int nops, int M, int N, int K, int passes, bool simt, bool tt, bool dw0 = false) {
  HostLaunch L = make_bench(nops, M, N, K, passes, tt, dw0);
  if (dw0) printf("(next: op 0 carries the fused dW0 + column-sum epilogue)\n");
  if (!simt) spin_warm(L, 300.0);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 2000;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) { if (simt) launch<true>(L); else launch<false>(L); }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double us = ms * 1e3 / iters;
  double flop = 2.0 * M * N * K * nops;
  printf("bench stream %s ops=%d M=%d N=%d K=%d passes=%d tt=%d: %.2f us/launch  %.2f TFLOP/s (algorithmic fp32)\n",
         simt ? "SIMT" : "TC  ", nops, M, N, K, passes, (int)tt, us, flop / us * 1e-6);
  if (simt) return;
  // the same launch as a 20-node dependent chain inside a CUDA graph
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < 20; ++i) launch<false>(L, st);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 200; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("bench graph  TC   (20-node chain): %.2f us/node\n", ms * 1e3 / (200 * 20));
  // phase profile of CTA 0
  long long* d_prof;
  CK(cudaMalloc(&d_prof, 24 * 8));
  CK(cudaMemset(d_prof, 0, 24 * 8));
  HostLaunch P = L;
  P.prof = d_prof;
  for (int i = 0; i < 3; ++i) launch<false>(P);
  CK(cudaDeviceSynchronize());
  long long h[24];
  CK(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
  double ns = (double)(h[10] - h[9]);
  printf("  prof cycles: setup=%lld issue_all=%lld first_full=%lld last_commit=%lld accum=%lld tmem_ld=%lld bar=%lld math=%lld epi_end=%lld exit=%lld | %.0f ns => %.0f MHz\n",
         h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[11] - h[0], h[12] - h[0], h[7] - h[0],
         h[8] - h[0], ns, (h[8] - h[0]) / ns * 1e3);
  if (g_ksplit > 1)
    printf("  rank 1 (own clock): readout_done=%lld cluster_wait_done=%lld sent=%lld\n", h[17] - h[16], h[18] - h[16], h[19] - h[16]);
  if (dw0)
    printf("  dW0 epilogue: start=%lld partials_stored=%lld ticket_known=%lld\n", h[13] - h[0], h[14] - h[0], h[15] - h[0]);
}

int main(int argc, char** argv) {
  CK(cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  CK(cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);

  int fails = 0;
  printf("== SIMT cross-check path (validates layout / copies / epilogue)\n");
  if (!(run_case(256, 256, 256, true, 3) < 2e-6)) ++fails;
  if (!(run_case(128, 32, 32, true, 3) < 2e-6)) ++fails;
  printf("== tcgen05 path\n");
  if (!(run_case(256, 256, 256, false, 3) < 2e-6)) ++fails;
  if (!(run_case(256, 256, 256, false, 1) < 1e-3)) ++fails;
  if (!(run_case(128, 32, 32, false, 3) < 2e-6)) ++fails;
  if (!(run_case(256, 32, 256, false, 3) < 2e-6)) ++fails;
  if (!(run_case(1024, 512, 96, false, 3) < 2e-6)) ++fails;
  if (!(run_case(256, 512, 512, false, 3) < 2e-6)) ++fails;
  printf("== split-K over clusters (2 / 4 CTAs per tile)\n");
  for (int ks = 2; ks <= 4; ks *= 2) {
    g_ksplit = ks;
    if (!(run_case(256, 256, 256, true, 3) < 2e-6)) ++fails;
    if (!(run_case(256, 256, 256, false, 3) < 2e-6)) ++fails;
    if (!(run_case(128, 32, 32, false, 3) < 2e-6)) ++fails;   // one chunk: ranks > 0 leave at once
    if (!(run_case(256, 64, 96, false, 3) < 2e-6)) ++fails;   // 3 chunks: uneven shares
    if (!(run_case(256, 32, 416, false, 3) < 2e-6)) ++fails;  // 13 chunks
  }
  g_ksplit = 1;
  printf("fails=%d\n", fails);

  if (argc > 1 && atoi(argv[1]) == 0) return fails ? 1 : 0;
  bench(1, 256, 256, 256, 3, false, false);
  for (int ks = 2; ks <= 4; ks *= 2) {
    g_ksplit = ks;
    printf("(next: %d CTAs per tile)\n", ks);
    bench(1, 256, 256, 256, 3, false, false);
    if (ks == 2) bench(3, 256, 256, 256, 3, false, true);
    if (ks == 2) bench(2, 256, 256, 256, 3, false, true, true);
  }
  g_ksplit = 1;
  bench(3, 256, 256, 256, 3, false, false);
  bench(3, 256, 256, 256, 3, false, true);
  bench(2, 256, 256, 256, 3, false, true, true);
  bench(3, 256, 256, 256, 1, false, false);
  bench(3, 256, 256, 32, 3, false, false);
  bench(5, 256, 512, 512, 3, false, false);
  bench(8, 1024, 256, 256, 3, false, false);
  bench(3, 256, 256, 256, 3, true, false);
  printf("SELFTEST %s\n", fails == 0 ? "PASS" : "FAIL");
  return fails ? 1 : 0;
}
