// Grouped small-GEMM kernel for the off-policy update: every MLP layer forward,
// dX and dW product of the DDPG/TD3/SAC/TQC update is one `GemmOp`.
//
//   D[M x N] = sum_k A(m,k) * B(n,k)       (fp32 accumulate in TMEM)
//
// * Operands live in HBM/L2 as plain fp32 in the CT32 "core-tiled" layout.  Both operands
//   are always K-major (tcgen05 kind::tf32 only takes MN-major operands in the
//   32-bit-swizzled SW128_32B layout), so producers that feed a dX (= delta * W) or
//   dW (= delta^T * X) product also write a transposed tiled copy (`tt` output here,
//   W^T from the Adam kernel).
// * fp32-accurate products on tensor cores via 3xTF32:
//       A*B ~= Alo*Bhi + Ahi*Blo + Ahi*Bhi        (tcgen05.mma.kind::tf32)
//   (needed for the reference's 1e-5 parameter-L2 parity bar).
//   The hi/lo split (hi = tf32(x), lo = tf32(x - hi)) is done IN the kernel by the four warps
//   that later run the epilogue: operands cross L2->SMEM once as fp32 instead of twice as
//   pre-split halves (the K loop is bound by the ~36 B/clk a single SM pulls from L2).  The
//   split A operand is written registers -> TMEM (tcgen05.st) and the MMAs read A from tensor
//   memory: with both operands in shared memory the 128 B/clk SMEM port (3 x 4 KB of A per
//   k-step + the splitter's own traffic) was the K-loop limit.
//   The tensor core's accumulator add is not round-to-nearest, so a long accumulation
//   chain loses ~1 ulp per step (measured: one 96-step chain gave 8x the parameter error of
//   the FFMA cross-check).  The chain is therefore cut: the two small cross terms go to
//   their own TMEM accumulator, the Ahi*Bhi terms to one accumulator per group of K
//   chunks, and the epilogue adds the partial sums with ordinary fp32 adds.
// * Tile 128 x 32 per CTA, K streamed in 32-wide chunks through an 8-stage mbarrier ring
//   filled by 1-D bulk async copies (TMA engine, no tensor maps: the producers already
//   wrote UMMA-canonical core matrices); the copy warp starts before TMEM allocation is done;
//   bias and the ReLU-mask tile of the epilogue are prefetched during the K loop; the
//   transposed output is staged through shared memory and stored as coalesced 16-byte vectors.
// * Warp roles: warp0 = copy producer, warp1 = MMA issuer (+TMEM owner), warps2-9 =
//   tf32 splitters during the K loop (two warps per TMEM lane quarter, alternating K chunks so
//   one group's load/convert/store latency hides behind the other's), then epilogue, 16 of the
//   32 tile columns each (TMEM -> regs -> bias/act/mask -> tiled + transposed tiled +
//   row-major + column sums).
// * `kSimt` variant keeps loads/epilogue identical but does the products with plain fp32
//   FFMA from shared memory: the on-device cross-check for the tensor-core path.
#pragma once
#include <cstdint>
#include <cstdio>
#include "ptx.cuh"

namespace oprl {

// ---------------------------------------------------------------- CT32 layout
// Padded [rows x cols] fp32 matrix, rows % 32 == 0, cols % 32 == 0.
// Block (r/8, c/32) is 1 KB contiguous; blocks are ordered column-block major:
//   float offset = ((c/32) * (rows/8) + r/8) * 256 + ((c%32)/4)*32 + (r%8)*4 + c%4
// i.e. inside a block there are 8 UMMA core matrices (8 rows x 16 bytes each).
__host__ __device__ __forceinline__ size_t ct_index(int rows, int r, int c) {
  return (static_cast<size_t>(c >> 5) * (rows >> 3) + (r >> 3)) * 256 + ((c & 31) >> 2) * 32 +
         (r & 7) * 4 + (c & 3);
}
__host__ __device__ __forceinline__ int pad32(int x) { return (x + 31) & ~31; }
__host__ __device__ __forceinline__ int pad128(int x) { return (x + 127) & ~127; }

constexpr int kBM = 128;
constexpr int kBN = 32;
constexpr int kBK = 32;
constexpr int kStages = 8;  // smem ring depth (raw fp32 chunks in flight: L2 latency x bandwidth)
constexpr int kASlots = 4;  // TMEM ring depth for the split A operand
constexpr int kAFloats = kBM * kBK;  // 4096 floats (16 KB)
constexpr int kBFloats = kBN * kBK;  // 1024 floats ( 4 KB)
// smem stage = [A raw][B raw -> B hi][B lo]; the split A operand lives in TMEM
constexpr int kStageFloats = kAFloats + 2 * kBFloats;
constexpr int kStageBytes = kStageFloats * 4;  // 24 KB
constexpr int kATmemCols = 2 * kBK;            // per stage: A hi (32 columns) | A lo (32 columns)
constexpr int kGemmThreads = 320;  // warp 0 copies, warp 1 MMA, warps 2-9 split + epilogue
constexpr int kEN = kBN / 2;        // epilogue columns per thread (two warps per TMEM lane quarter)
constexpr int kMaskBytes = kBM * kBN * 4;  // ReLU-mask tile of the epilogue, prefetched
constexpr int kGemmSmemBytes = kStages * kStageBytes + kMaskBytes + 1280;
constexpr int kTTPitch = kBM + 4;  // smem pitch of the transposed staging tile (conflict-free)
constexpr int kMaxOps = 12;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct GemmOp {
  // operands (CT32 fp32)
  const float* a;
  const float* b;
  // epilogue inputs
  const float* bias;  // v += bias[n] for n < bias_n
  const float* mask;  // v *= (mask[m][n] > 0), CT32 with mask_rows padded rows
  const float* rs;    // row-major [m][n] matrix: v *= (1 - rs^2) for n < rs_n (tanh')
  const float* addm;  // row-major [m][n] matrix added after the activation, n < addm_n
  // outputs
  float* t;       // CT32 output at column offset t_c0, only columns n < t_n
  float* tt;      // transposed CT32 output: element (n, m) of a [tt_rows x M] matrix
  float* rm;      // row-major output, m < rm_m, n < rm_n
  float* colsum;  // per-M-tile partials: colsum[mtile * colsum_ld + n]
  float* colsum_out;         // if set: the last CTA of each N tile writes the total over M tiles here
  unsigned int* colsum_cnt;  // one arrival counter per N tile (self-resetting)
  int a_rows;  // padded row count of the stored A matrix [M.. x K]
  int b_rows;  // padded row count of the stored B matrix [N.. x K]
  int M, N, K;  // padded problem (M % 128, N % 32, K % 32)
  int bias_n, act;
  int mask_rows;
  int rs_ld, rs_n;
  int addm_ld, addm_n;
  int m_valid;  // rows m >= m_valid are forced to 0 before any output (0 = no limit)
  int n_valid;  // columns n >= n_valid are forced to 0 before any output (0 = no limit)
  int t_rows, t_c0, t_n;
  int tt_rows;
  int rm_ld, rm_trans, rm_m, rm_n;
  // optional column map for the row-major store (critic layer-1 keeps its input as
  // [action | pad4 | state]):  n < map_a -> map_s + n ; n >= map_a4 -> n - map_a4.
  int map_a, map_a4, map_s;
  int colsum_ld, colsum_n;  // partial row stride; colsum_out gets columns n < colsum_n
  // Scalar head riding on the last hidden layer (critic.Q1(s, pi(s)) of the DDPG / TD3 actor step,
  // ddpg.py:104, td3.py:135-137): with D = h (post-ReLU) and aux_vec = the head weights w3,
  //   tail_out[(2 * ntile + half) * M + m] = sum_{n in this thread's 16 columns} D(m, n) w3[n]
  // (the consumer adds the 2 * N/32 partials per row: q[m] - b3), and
  //   aux_t(m, n) = D(m, n) > 0 && m < aux_m ? aux_alpha * w3[n] : 0            (tiled [M x N])
  // is dz of this layer for the constant seed dL/dq = aux_alpha -- the backward chain starts from it
  // without a separate head kernel on the critical path.
  const float* aux_vec;
  float* aux_t;
  float* tail_out;
  float aux_alpha;
  int aux_m;
  // Fused layer-0 weight gradient: this op's output D [M x N] is dz_0 (batch x hidden) and
  //   dW_0[n][k] = sum_m D(m, n) * X(m, k)          (X = the tiled layer-0 input, dw0_kp columns)
  // is accumulated in the epilogue with fp32 FMAs (per-CTA partial over its 128 rows, fixed-order
  // sum over the M tiles by the last CTA to arrive) and stored row-major at dw0_out[n * dw0_ld + col],
  // col = the dw0_map_* column map, rows n < dw0_n.  Replaces a 2-CTA GEMM stage.
  const float* dw0_x;
  float* dw0_part;        // [M tiles][N tiles][32][dw0_kp]
  float* dw0_out;
  unsigned int* dw0_cnt;  // one arrival counter per N tile (self-resetting)
  int dw0_kp, dw0_ld, dw0_n, dw0_cols;  // dw0_cols: tiled columns k < dw0_cols are candidates for the store
  int dw0_map_a, dw0_map_a4, dw0_map_s;  // column map of the dW_0 store (same meaning as map_a/map_a4/map_s)
  // bias gradient through the same product: tiled column dw0_ones (a pad column of X, < 0 = none) is
  // read as 1.0, so P[n][dw0_ones] = sum_m D(m, n) = db_0[n]; it is stored to dw0_bias_out[n].
  int dw0_ones;
  float* dw0_bias_out;
  // dw0_defer: stop after the per-M-tile partials are stored -- the Adam kernel adds them (same order) when it
  // fetches the gradient, so the arrival ticket and the last CTA's reduction leave the update's critical path
  // (requires dw0_ones >= 0: the bias gradient rides in the same partials)
  int dw0_defer;
  int passes;   // 3 = 3xTF32 (fp32-accurate), 1 = single tf32 pass
  int group;    // K chunks per hi*hi accumulator (gemm_finalize)
  int n_big;    // number of hi*hi accumulators, <= 7
  float alpha;  // v *= alpha (applied last)
  float clamp;  // if > 0: v = min(max(v, -clamp), clamp) after addm
};

struct GemmLaunch {
  // first cache line of the parameter block: everything the tile -> op lookup needs
  int n_ops;
  // Split-K over a thread-block cluster: ksplit (1, 2 or 4) CTAs share one output tile, the K
  // chunks dealt round-robin.  A single SM pulls only ~31-36 B/clk from L2, so the K loop of one
  // tile (160 KB at K = 256) is ingest-bound; with the chunks spread over ksplit SMs each CTA
  // streams 1/ksplit of them.  Ranks > 0 push their fp32 partial tile into rank 0's shared memory
  // (DSMEM stores + a remote mbarrier arrive); rank 0 adds the partials in rank order and runs the
  // epilogue.  The launch carries cluster dimension (ksplit, 1, 1) and ksplit x the CTAs.
  int ksplit;
  int nsub;               // 1: 128 x 32 tiles, 2: 128 x 64 tiles (gemm_kernel<.., 2>, ksplit == 1 only)
  int tile_end[kMaxOps];  // exclusive prefix sums of gemm_tiles(op[i])
  long long* prof;        // selftest only: per-phase clock64 stamps of CTA 0
  // The op descriptors (~400 B each) live in device memory, written once when the program is built:
  // passing them by value made every GEMM node of the update graph carry 4.4 KB of kernel parameters
  // (54 KB per graph launch to patch on the host).  Each CTA copies its descriptor into shared
  // memory at kernel start -- before the programmatic-dependent-launch wait, the table is static.
  const GemmOp* ops;
};

__host__ __device__ __forceinline__ int gemm_tiles(const GemmOp& o, int nsub = 1) {
  return (o.M / kBM) * ((o.N + kBN * nsub - 1) / (kBN * nsub));
}
// Wide-tile variant (`GemmLaunch::nsub` = 2, tile 128 x 64): the same K loop with N = 64 MMAs -- the split A operand
// is produced once per 64 output columns instead of once per 32, and a launch of T narrow tiles becomes T / 2 CTAs
// (TQC's 160- and 480-tile launches drop from 2 and 4 waves of 148 SMs to 1 and 2).  Shared memory: 5 ring stages of
// 32 KB (A 16 KB, B raw/hi 8 KB, B lo 8 KB) + a 32 KB mask tile; tensor memory: <= 5 accumulators x 64 columns +
// 3 A slots x 64 columns = 512.  The epilogue runs the narrow tile's code twice (columns 0-31, then 32-63).
// An op whose N is an odd multiple of 32 ends in a ragged tile: only 32 rows of B are copied, the MMAs still run 64
// wide over whatever the second half of the B buffer holds, and the second epilogue pass is skipped (those
// accumulator columns are never read).  Not available to ops that use the tanh' factors or the riding scalar head
// (their per-tile state is sized for 32 columns), nor to split-K launches.
constexpr int kWideStages = 5;
constexpr int kWideASlots = 3;
constexpr int kWideMaxBig = 4;
constexpr int kGemmSmemBytesWide = kWideStages * (kAFloats + 4 * kBFloats) * 4 + 2 * kMaskBytes + 1280;
inline bool gemm_wide_ok(const GemmOp& o) {
  return !o.rs && !o.aux_vec;
}
// host: derived fields (TMEM accumulator plan) -- call once per op before launching
inline void gemm_finalize(GemmOp& o, int max_big = 7) {
  const int nchunks = o.K / kBK;
  o.group = (nchunks + max_big - 1) / max_big > 2 ? (nchunks + max_big - 1) / max_big : 2;
  o.n_big = (nchunks + o.group - 1) / o.group;
}
// host: CTAs per tile for one launch -- the largest of {4, 2, 1} that keeps the whole launch in one
// wave of `n_sm` single-CTA SMs and is worth the exchange (an op must have at least 2 chunks per CTA)
// `max4`: the most CTAs a launch of 4-CTA clusters may have.  A cluster needs its 4 SMs free in ONE GPC at the same
// time; at 32 clusters (128 CTAs on 148 SMs) some wait for a second round behind the previous kernel's last CTAs.
// Measured (B200, one box): cap 148 -> 112: DDPG 88.2 -> 87.5, TD3 76.0 -> 72.0 us per update.  Long K loops (more
// than 16 chunks: SAC's batch-1024 weight gradients) keep the 4-way split, there the shorter loop is worth more.
constexpr int kSplit4MaxCtas = 112;
inline int gemm_choose_ksplit(const GemmOp* ops, int n_ops, int n_sm, int max4 = kSplit4MaxCtas) {
  int tiles = 0, max_chunks = 0;
  for (int i = 0; i < n_ops; ++i) {
    tiles += gemm_tiles(ops[i]);
    max_chunks = ops[i].K / kBK > max_chunks ? ops[i].K / kBK : max_chunks;
  }
  for (int ks = 4; ks >= 2; ks >>= 1) {
    if (tiles * ks > n_sm) continue;
    if (ks == 4 && tiles * ks > max4 && max_chunks <= 16) continue;
    if (max_chunks < 2 * ks) continue;
    return ks;
  }
  return 1;
}
// keep a value in a register from here on (stops ptxas from re-reading kernel parameters,
// one dependent constant-bank load per use, inside the latency-critical epilogue)
template <typename T>
__device__ __forceinline__ void pin(T& x) {
  asm volatile("" : "+r"(x));
}
__device__ __forceinline__ void pin(float& x) { asm volatile("" : "+f"(x)); }
template <typename T>
__device__ __forceinline__ void pin_ptr(T*& x) {
  asm volatile("" : "+l"(x));
}

template <bool kSimt, int kNSub = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_kernel(const __grid_constant__ GemmLaunch L) {
  static_assert(kNSub == 1 || (kNSub == 2 && !kSimt), "wide tiles exist for the tensor-core path only");
  // tile geometry of this instantiation (the k* names below shadow the narrow-tile constants of the file)
  constexpr int kBNt = kBN * kNSub;                      // tile columns
  constexpr int kBFl = kBFloats * kNSub;                 // floats of one raw B chunk
  constexpr int kStageFl = kAFloats + 2 * kBFl;          // [A raw][B raw -> hi][B lo]
  constexpr int kStageBy = kStageFl * 4;
  constexpr int kNStages = kNSub == 2 ? kWideStages : kStages;
  constexpr int kNASlots = kNSub == 2 ? kWideASlots : kASlots;
  constexpr int kMaskBy = kMaskBytes * kNSub;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  float* mask_smem = reinterpret_cast<float*>(smem_raw + kNStages * kStageBy);
  uint8_t* ctl = smem_raw + kNStages * kStageBy + kMaskBy;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctl);
  uint64_t* empty = full + kNStages;
  uint64_t* conv = empty + kNStages;
  uint64_t* accum = conv + kNStages;
  uint64_t* mask_bar = accum + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 1);
  float* cs_smem = reinterpret_cast<float*>(ctl + 256);    // [4][32]
  float* bias_smem = reinterpret_cast<float*>(ctl + 768);  // [32]
  float* aux_smem = reinterpret_cast<float*>(ctl + 1024);  // [32]: this tile's slice of GemmOp::aux_vec

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  uint64_t* xbar = reinterpret_cast<uint64_t*>(ctl + 904);  // split-K: partial tiles have landed

  ptx::pdl_trigger();
  // ---- which op / tile is this CTA
  const int ks = (kNSub == 1 && L.ksplit > 1) ? L.ksplit : 1;
  const int kr = ks > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  int t = ks > 1 ? static_cast<int>(blockIdx.x) / ks : static_cast<int>(blockIdx.x);
  int oi = 0;
  while (oi < L.n_ops && t >= L.tile_end[oi]) ++oi;
  if (oi >= L.n_ops) return;
  if (oi > 0) t -= L.tile_end[oi - 1];
  __shared__ __align__(16) GemmOp so;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(L.ops + oi);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&so);
    for (int i = tid; i < static_cast<int>(sizeof(GemmOp) / 4); i += kGemmThreads) dst[i] = __ldg(src + i);
    __syncthreads();
  }
  const GemmOp& o = so;
  const int ntn = (o.N + kBNt - 1) / kBNt;
  const int mt = t / ntn;
  const int m0 = mt * kBM;
  const int n0 = (t % ntn) * kBNt;
  const int nsub = (kNSub == 2 && n0 + kBN >= o.N) ? 1 : kNSub;  // 32-column sub-tiles this CTA owns (ragged last tile: 1)
  const int nchunks = o.K / kBK;
  // this CTA's share of the K chunks: global chunk kr + ks * cl for cl < nloc
  const int nloc = kr < nchunks ? (nchunks - kr + ks - 1) / ks : 0;
  const int active = nchunks < ks ? nchunks : ks;  // CTAs of the cluster that hold a partial tile
  if (nloc == 0) {  // (only with ks > 1) nothing to add: leave after the cluster-wide arrive
    ptx::cluster_arrive();
    return;
  }
  const int passes = o.passes;
  // ring depth: a split-K CTA keeps the last two stages (48 KB) as landing slots for the partial tiles of ranks 1-3
  const int nst = (kNSub == 1 && ks > 1) ? kStages - 2 : kNStages;
  long long* prof = (L.prof && blockIdx.x == 0) ? L.prof : nullptr;
  long long* prof1 = (L.prof && blockIdx.x == 1 && ks > 1) ? L.prof : nullptr;  // rank 1 of tile 0
  if (prof1 && tid == 0) prof1[16] = clock64();
  if (prof && tid == 0) {
    prof[0] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    prof[9] = static_cast<long long>(gt);
  }
  // TMEM map: accumulator column block 0 = cross terms, blocks 1..n_big = hi*hi per group of K
  // chunks (at most 7), then kASlots x (A hi | A lo) operand blocks written by the splitter warps.
  int group = o.group;
  int n_big = o.n_big;
  if (ks > 1) {  // the accumulator plan of gemm_finalize, for this CTA's chunk count
    group = (nloc + 6) / 7 > 2 ? (nloc + 6) / 7 : 2;
    n_big = (nloc + group - 1) / group;
  }
  const uint32_t a_col0 = static_cast<uint32_t>(kBNt * (n_big + 1));
  uint32_t tmem_cols = 32;
  while (tmem_cols < a_col0 + kNASlots * kATmemCols) tmem_cols <<= 1;

  float v[kEN];

  if (warp == 0) {
    // ===================== producer: bulk copies HBM/L2 -> smem ring (fp32, once)
    if (lane < kNStages) {
      ptx::mbar_init(&full[lane], 1);
      ptx::mbar_init(&empty[lane], 1);
      ptx::mbar_init(&conv[lane], 4);  // one arrival per splitter warp
    } else if (lane == kNStages) {
      ptx::mbar_init(accum, 1);
      ptx::mbar_init(mask_bar, 1);
      ptx::mbar_init(xbar, static_cast<uint32_t>(active > 1 ? (active - 1) * 8 : 1));  // one arrival per remote epilogue warp
    }
    ptx::fence_mbar_init();
    __syncwarp();
    if (ks > 1) ptx::cluster_arrive();  // publishes the barrier init to the cluster
    // the other warps wait on this barrier (after TMEM allocation); the copies start now
    asm volatile("bar.arrive 3, %0;\n" ::"n"(kGemmThreads) : "memory");
    if (prof && lane == 0) prof[1] = clock64();
    ptx::pdl_wait();  // operands are the previous kernel's outputs
    const int a_rb = o.a_rows >> 3;
    const int b_rb = o.b_rows >> 3;
    const uint32_t b_bytes = static_cast<uint32_t>(kBFloats * nsub) * 4u;
    const uint32_t tx = kAFloats * 4u + b_bytes;
    int s = 0;
    uint32_t ph = 0;
    for (int cl = 0; cl < nloc; ++cl) {
      const int c = kr + cl * ks;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (ptx::elect_one()) {
        float* st = smem + s * kStageFl;
        ptx::mbar_expect_tx(&full[s], tx);
        ptx::bulk_g2s(st, o.a + (static_cast<size_t>(c) * a_rb + (m0 >> 3)) * 256, kAFloats * 4, &full[s]);
        // (the 32 * kNSub rows of B are kNSub * 4 consecutive 1 KB row-group blocks of this K chunk)
        ptx::bulk_g2s(st + kAFloats, o.b + (static_cast<size_t>(c) * b_rb + (n0 >> 3)) * 256, b_bytes, &full[s]);
        if (c == 0 && o.mask) {  // (chunk 0 belongs to rank 0, the CTA that runs the epilogue)
          // epilogue ReLU mask: each [128 x 32] tile of the saved activation is one contiguous 16 KB
          ptx::mbar_expect_tx(mask_bar, static_cast<uint32_t>(kMaskBytes * nsub));
          for (int h = 0; h < nsub; ++h)
            ptx::bulk_g2s(mask_smem + h * (kMaskBytes / 4),
                          o.mask + (static_cast<size_t>((n0 >> 5) + h) * (o.mask_rows >> 3) + (m0 >> 3)) * 256,
                          kMaskBytes, mask_bar);
        }
      }
      __syncwarp();
      if (++s == nst) {
        s = 0;
        ph ^= 1;
      }
    }
    if (prof && lane == 0) prof[2] = clock64();
  } else {
    if (ks > 1) ptx::cluster_arrive();
    if (!kSimt && warp == 1) ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tc_fence_before();
    asm volatile("bar.sync 3, %0;\n" ::"n"(kGemmThreads) : "memory");
    ptx::tc_fence_after();
  }

  if (!kSimt && warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole K loop (waits included).
    // Its program is serial -- every instruction between two MMAs costs its full latency -- so the loop keeps to
    // the barrier poll, the operand words (descriptor low word advanced by adds, not rebuilt per MMA) and the
    // MMAs; the poll of the NEXT chunk's barrier (~140 cycles round trip) is sampled before this chunk's MMAs
    // are issued.  (Electing per chunk and rebuilding both 64-bit descriptors per MMA cost ~55 cycles per MMA,
    // which -- not the L2 ingest -- was what bound the K loop; tools/experiments/chain_probe5.cu.)
    const uint32_t tmem_d = *tmem_slot;
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_tf32(kBM, kBNt, 0, 0);
      // B smem tile = [row group of 8][8 K-cores][8 rows][16 B]: K cores 128 B apart (LBO), 8-row groups 1 KB
      // apart (SBO); one MMA (K = 8) consumes two K cores = 256 B = 16 descriptor units.
      const uint32_t dw_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14);
      const uint32_t lbo_bits = (128u >> 4) << 16;
      const uint32_t conv0 = ptx::smem_u32(&conv[0]), empty0 = ptx::smem_u32(&empty[0]);
      const uint32_t sb0 = ptx::smem_u32(smem + kAFloats);
      int in_group = 0;
      uint32_t big = tmem_d + static_cast<uint32_t>(kBNt);
      uint32_t ok = 0;
      int s = 0, aslot = 0;
      uint32_t par = 0;
      for (int c = 0; c < nloc; ++c) {  // c: local chunk index
        if (!ok) {
          uint32_t spins = 0;
          while (!ptx::mbar_test_wait_addr(conv0 + s * 8u, par)) {
            if (++spins > (1u << 26)) __trap();
          }
        }
        ptx::tc_fence_after();
        if (prof && c == 0) prof[3] = clock64();
        int s1 = s + 1;
        uint32_t par1 = par;
        if (s1 == nst) {
          s1 = 0;
          par1 ^= 1u;
        }
        ok = (c + 1 < nloc) ? ptx::mbar_test_wait_addr(conv0 + static_cast<uint32_t>(s1) * 8u, par1) : 0u;
        const uint32_t dl_hi = (((sb0 + static_cast<uint32_t>(s) * kStageBy) >> 4) & 0x3FFFu) | lbo_bits;
        const uint32_t dl_lo = dl_hi + ((kBFl * 4u) >> 4);
        const uint32_t ta_hi = tmem_d + a_col0 + static_cast<uint32_t>(aslot * kATmemCols);
        const uint32_t ta_lo = ta_hi + kBK;
#pragma unroll
        for (int j = 0; j < kBK / 8; ++j) {
          const uint32_t big_acc = (in_group | j) ? 1u : 0u;
          if (passes == 3) {
            ptx::mma_tf32_ts2(tmem_d, ta_lo + 8u * j, dl_hi + 16u * j, dw_hi, idesc, (c | j) ? 1u : 0u);
            ptx::mma_tf32_ts2(tmem_d, ta_hi + 8u * j, dl_lo + 16u * j, dw_hi, idesc, 1u);
          }
          ptx::mma_tf32_ts2(big, ta_hi + 8u * j, dl_hi + 16u * j, dw_hi, idesc, big_acc);
        }
        ptx::mma_commit_addr(empty0 + s * 8u);
        if (++in_group == group) {
          in_group = 0;
          big += static_cast<uint32_t>(kBNt);
        }
        s = s1;
        par = par1;
        if (++aslot == kNASlots) aslot = 0;
      }
      ptx::mma_commit(accum);
      if (prof) prof[4] = clock64();
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ===================== splitter + epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // 0: warps 2-5 (even chunks, columns 0-15), 1: warps 6-9
    const int cn0 = half * kEN;        // first tile column this thread finishes
    const int row = q * 32 + lane;
    const int ct = (tid - 64) & 127;   // 0..127 inside the group
    ptx::pdl_wait();  // bias / mask / outputs alias buffers the previous kernel may still use
    if (tid - 64 < kBN) bias_smem[tid - 64] = (o.bias && n0 + tid - 64 < o.bias_n) ? __ldg(o.bias + n0 + tid - 64) : 0.f;
    if (tid - 64 >= kBN && tid - 64 < 2 * kBN) {
      // narrow tile: this tile's slice of the riding head's weights; wide tile (no riding head): the bias of columns 32-63
      if (kNSub == 1) aux_smem[tid - 64 - kBN] = o.aux_vec ? __ldg(o.aux_vec + n0 + tid - 64 - kBN) : 0.f;
      else aux_smem[tid - 64 - kBN] = (o.bias && n0 + tid - 64 < o.bias_n) ? __ldg(o.bias + n0 + tid - 64) : 0.f;
    }
    // epilogue plan, read from the kernel parameters now (while the first copies are in flight)
    enum : uint32_t { F_BIAS = 1, F_RELU = 2, F_TANH = 4, F_MASK = 8, F_RS = 16, F_ADDM = 32, F_CLAMP = 64,
                      F_ALPHA = 128, F_MVALID = 256, F_NVALID = 512, F_T = 1024, F_TT = 2048, F_RM = 4096,
                      F_COLSUM = 8192, F_DW0 = 16384, F_AUX = 32768 };
    uint32_t fl = (o.bias ? F_BIAS : 0u) | (o.act == ACT_RELU ? F_RELU : 0u) | (o.act == ACT_TANH ? F_TANH : 0u) |
                  (o.mask ? F_MASK : 0u) | (o.rs ? F_RS : 0u) | (o.addm ? F_ADDM : 0u) |
                  (o.clamp > 0.f ? F_CLAMP : 0u) | (o.alpha != 1.f ? F_ALPHA : 0u) |
                  (o.m_valid > 0 ? F_MVALID : 0u) | (o.n_valid > 0 ? F_NVALID : 0u) | (o.t ? F_T : 0u) |
                  (o.tt ? F_TT : 0u) | (o.rm ? F_RM : 0u) | (o.colsum ? F_COLSUM : 0u) | (o.dw0_out ? F_DW0 : 0u) |
                  (o.aux_vec ? F_AUX : 0u);
    float* out_t = o.t;
    float* out_tt = o.tt;
    int t_rows = o.t_rows, t_c0 = o.t_c0, t_n = o.t_n, tt_rows = o.tt_rows;
    pin(fl); pin_ptr(out_t); pin_ptr(out_tt); pin(t_rows); pin(t_c0); pin(t_n); pin(tt_rows);
    // tanh' factors of the epilogue (rs), fetched now: their latency hides behind the K loop
    float rsv[kEN];
#pragma unroll
    for (int j = 0; j < kEN; ++j) rsv[j] = 0.f;
    if (fl & F_RS) {
#pragma unroll
      for (int j = 0; j < kEN; ++j)
        if (n0 + cn0 + j < o.rs_n) rsv[j] = __ldg(o.rs + static_cast<size_t>(m0 + row) * o.rs_ld + n0 + cn0 + j);
    }
    // fused dW_0: the [128 x 32] tile of the layer-0 input this CTA's rows multiply (one contiguous
    // 16 KB of the CT32 matrix) is fetched into registers now; its latency hides behind the K loop
    float4 xr[kAFloats / 4 / 256];
    if (fl & F_DW0) {
      const float4* xg = reinterpret_cast<const float4*>(o.dw0_x + static_cast<size_t>(m0 >> 3) * 256);
#pragma unroll
      for (int i = 0; i < kAFloats / 4 / 256; ++i) xr[i] = __ldg(xg + (tid - 64) + i * 256);
    }
    // this thread's 16 columns of the 32-column sub-tile `h`: the hi*hi partial sums of the accumulator groups, then
    // the cross terms, added with ordinary fp32 adds (accumulator blocks are kBNt columns apart)
    auto read_acc = [&](int h) {
      const uint32_t lane_base =
          *tmem_slot + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cn0 + kBN * h);
      float part[kEN];
      ptx::tmem_ld16(lane_base + static_cast<uint32_t>(kBNt), v);
      for (int gi = 1; gi < n_big; ++gi) {
        ptx::tmem_ld16(lane_base + static_cast<uint32_t>(kBNt * (1 + gi)), part);
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] += part[j];
      }
      if (passes == 3) {
        ptx::tmem_ld16(lane_base, part);
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] += part[j];
      }
    };
    if (!kSimt) {
      // ---- K loop: split every landed fp32 chunk into tf32 hi / lo.  A: this thread's row
      // (32 k) goes registers -> TMEM (the MMA reads A from tensor memory, so shared memory only
      // serves the narrow B operand); B: in place (hi) + a second 4 KB buffer (lo).
      const uint32_t ta_lane = *tmem_slot + (static_cast<uint32_t>(q * 32) << 16) + a_col0;
      int s = half;  // ring stage / phase of chunk c, and of chunk c - kNASlots (the A slot's previous user)
      uint32_t ph = 0;
      int sp = (kNASlots & 1) == half ? 0 : 1;  // (first c >= kNASlots of this group's parity) - kNASlots
      uint32_t php = 0;
      for (int c = half; c < nloc; c += 2) {  // c: local chunk index
        ptx::mbar_wait(&full[s], ph);
        const float* st = smem + s * kStageFl;
        float hi[kBK], lo[kBK];
#pragma unroll
        for (int j = 0; j < kBK / 4; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(st + ((row >> 3) * 8 + j) * 32 + (row & 7) * 4);
          ptx::split_tf32(x.x, hi[4 * j + 0], lo[4 * j + 0]);
          ptx::split_tf32(x.y, hi[4 * j + 1], lo[4 * j + 1]);
          ptx::split_tf32(x.z, hi[4 * j + 2], lo[4 * j + 2]);
          ptx::split_tf32(x.w, hi[4 * j + 3], lo[4 * j + 3]);
        }
        if (c >= kNASlots) {
          // the TMEM slot is free once the MMAs of chunk c - kNASlots retired (their commit
          // arrives on that chunk's `empty` barrier)
          ptx::mbar_wait(&empty[sp], php);
          ptx::tc_fence_after();
          sp += 2;
          if (sp >= nst) {
            sp -= nst;
            php ^= 1u;
          }
        }
        const uint32_t ta = ta_lane + static_cast<uint32_t>((c % kNASlots) * kATmemCols);
        ptx::tmem_st32(ta, hi);
        if (passes == 3) ptx::tmem_st32(ta + kBK, lo);
        float4* braw = reinterpret_cast<float4*>(smem + s * kStageFl + kAFloats);
        float4* blo = braw + kBFl / 4;
#pragma unroll
        for (int i = 0; i < kBFl / 4 / 128; ++i) {
          const int idx = ct + i * 128;
          const float4 x = braw[idx];
          float4 h, l;
          ptx::split_tf32(x.x, h.x, l.x);
          ptx::split_tf32(x.y, h.y, l.y);
          ptx::split_tf32(x.z, h.z, l.z);
          ptx::split_tf32(x.w, h.w, l.w);
          braw[idx] = h;
          if (passes == 3) blo[idx] = l;
        }
        ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&conv[s]);
        s += 2;
        if (s >= nst) {
          s -= nst;
          ph ^= 1u;
        }
      }
      ptx::mbar_wait(accum, 0);
      ptx::tc_fence_after();
      if (prof && tid == 64) prof[5] = clock64();
      read_acc(0);
      if (prof && tid == 64) prof[6] = clock64();
    } else {
      // FFMA cross-check path: same smem contents, products on CUDA cores in plain fp32.
#pragma unroll
      for (int j = 0; j < kEN; ++j) v[j] = 0.f;
      int s = 0;
      uint32_t ph = 0;
      for (int c = 0; c < nloc; ++c) {
        ptx::mbar_wait(&full[s], ph);
        const float* sa = smem + s * kStageFloats;
        const float* sb = sa + kAFloats;
        for (int k = 0; k < kBK; ++k) {
          const float a = sa[((row >> 3) * 8 + (k >> 2)) * 32 + (row & 7) * 4 + (k & 3)];
#pragma unroll
          for (int j = 0; j < kEN; ++j) {
            const int n = cn0 + j;
            v[j] = fmaf(a, sb[((n >> 3) * 8 + (k >> 2)) * 32 + (n & 7) * 4 + (k & 3)], v[j]);
          }
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        if (tid == 64) ptx::mbar_arrive(&empty[s]);
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    // every MMA (or FFMA pass) of this tile is done: the smem ring is free for epilogue staging
    asm volatile("bar.sync 1, 256;\n" ::: "memory");  // also publishes bias_smem
    if (kNSub == 1 && active > 1) {
      // ---- split-K: partial tiles travel to rank 0 through distributed shared memory (plain
      // st.shared::cluster stores; a 16 KB cp.async.bulk shared::cta -> shared::cluster copy measured
      // slower, ~2 200 vs ~1 800 cycles: one SM pair moves only ~8 B/clk over DSMEM either way).
      // Slot of rank r = 16 KB number r - 1 behind the (shortened) ring of rank 0; element (row, 4 columns j4)
      // of a thread sits at float4 index (half * 4 + j4) * 128 + row in both CTAs.
      if (kr > 0) {
        if (prof1 && tid == 64) prof1[17] = clock64();
        ptx::cluster_wait();  // rank 0 has initialised xbar (every thread of the cluster arrived at kernel start)
        if (prof1 && tid == 64) prof1[18] = clock64();
        const uint32_t slot = ptx::mapa(ptx::smem_u32(smem + nst * kStageFloats + (kr - 1) * kAFloats), 0);
#pragma unroll
        for (int j4 = 0; j4 < kEN / 4; ++j4)
          ptx::st_cluster_v4(slot + static_cast<uint32_t>(((half * 4 + j4) * 128 + row) * 16), v[4 * j4 + 0],
                             v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        // one remote arrive per warp (remote arrivals on one mbarrier serialise); the warp barrier
        // orders the other lanes' stores before lane 0's release
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(xbar), 0));
        if (prof1 && tid == 64) prof1[19] = clock64();
        fl = 0;  // rank 0 runs the epilogue
      } else {
        ptx::mbar_wait_cluster(xbar, 0);
        for (int r = 1; r < active; ++r) {
          const float4* slot = reinterpret_cast<const float4*>(smem + nst * kStageFloats + (r - 1) * kAFloats);
#pragma unroll
          for (int j4 = 0; j4 < kEN / 4; ++j4) {
            const float4 p4 = slot[(half * 4 + j4) * 128 + row];
            v[4 * j4 + 0] += p4.x;
            v[4 * j4 + 1] += p4.y;
            v[4 * j4 + 2] += p4.z;
            v[4 * j4 + 3] += p4.w;
          }
        }
      }
    }
    if (prof && tid == 64) prof[11] = clock64();

    const int m = m0 + row;
    // The epilogue handles one 32-column sub-tile at a time (a wide tile has two; its second pass re-reads the
    // accumulators and reuses the staging buffers behind a barrier).
#pragma unroll 1
    for (int h = 0; h < nsub; ++h) {
      if (h > 0) {
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        read_acc(h);
      }
      const int n0h = n0 + kBN * h;
      const int nb = n0h + cn0;  // first global column of this thread
      const float* bsm = h ? aux_smem : bias_smem;
      const float* msm = mask_smem + h * (kMaskBytes / 4);
      // ---- epilogue math
      if (fl & F_BIAS) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] += bsm[cn0 + j];
      }
      if (fl & F_RELU) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (fl & F_TANH) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] = tanhf(v[j]);
      }
      if (fl & F_MASK) {
        ptx::mbar_wait(mask_bar, 0);
#pragma unroll
        for (int j4 = 0; j4 < kEN / 4; ++j4) {
          const float4 mk = *reinterpret_cast<const float4*>(msm + ((row >> 3) * 8 + cn0 / 4 + j4) * 32 + (row & 7) * 4);
          v[4 * j4 + 0] = mk.x > 0.f ? v[4 * j4 + 0] : 0.f;
          v[4 * j4 + 1] = mk.y > 0.f ? v[4 * j4 + 1] : 0.f;
          v[4 * j4 + 2] = mk.z > 0.f ? v[4 * j4 + 2] : 0.f;
          v[4 * j4 + 3] = mk.w > 0.f ? v[4 * j4 + 3] : 0.f;
        }
      }
      if (fl & F_RS) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] *= (1.f - rsv[j] * rsv[j]);  // rsv = 0 beyond rs_n
      }
      if (fl & F_ADDM) {
#pragma unroll
        for (int j = 0; j < kEN; ++j)
          if (nb + j < o.addm_n) v[j] += o.addm[static_cast<size_t>(m) * o.addm_ld + nb + j];
      }
      if (fl & F_CLAMP) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] = fminf(fmaxf(v[j], -o.clamp), o.clamp);
      }
      if (fl & F_ALPHA) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] *= o.alpha;
      }
      if ((fl & F_MVALID) && m >= o.m_valid) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) v[j] = 0.f;
      }
      if (fl & F_NVALID) {
#pragma unroll
        for (int j = 0; j < kEN; ++j)
          if (nb + j >= o.n_valid) v[j] = 0.f;
      }
      if (prof && tid == 64) prof[12] = clock64();
      // ---- outputs
      if (fl & F_AUX) {
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < kEN; ++j) dot = fmaf(v[j], aux_smem[cn0 + j], dot);
        o.tail_out[static_cast<size_t>(2 * (n0h >> 5) + half) * o.M + m] = dot;
        const float aa = m < o.aux_m ? o.aux_alpha : 0.f;
#pragma unroll
        for (int j4 = 0; j4 < kEN / 4; ++j4)
          *reinterpret_cast<float4*>(o.aux_t + ct_index(t_rows, m, nb + 4 * j4)) =
              make_float4(v[4 * j4 + 0] > 0.f ? aa * aux_smem[cn0 + 4 * j4 + 0] : 0.f,
                          v[4 * j4 + 1] > 0.f ? aa * aux_smem[cn0 + 4 * j4 + 1] : 0.f,
                          v[4 * j4 + 2] > 0.f ? aa * aux_smem[cn0 + 4 * j4 + 2] : 0.f,
                          v[4 * j4 + 3] > 0.f ? aa * aux_smem[cn0 + 4 * j4 + 3] : 0.f);
      }
      if (fl & F_T) {
#pragma unroll
        for (int j4 = 0; j4 < kEN / 4; ++j4) {
          if (nb + 4 * j4 < t_n) {
            *reinterpret_cast<float4*>(out_t + ct_index(t_rows, m, t_c0 + nb + 4 * j4)) =
                make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          }
        }
      }
      if (fl & F_TT) {
        // transposed output [32 n x 128 m]: stage through smem (lanes = consecutive m, conflict
        // free), then each thread stores 4 coalesced float4 = (n, 4 consecutive m) of the CT32 blocks
        float* sT = smem;  // [32][kTTPitch]
#pragma unroll
        for (int j = 0; j < kEN; ++j) sT[(cn0 + j) * kTTPitch + row] = v[j];
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
#pragma unroll
        for (int i = 0; i < (kBM * kBN / 4) / 256; ++i) {
          const int f = (tid - 64) + i * 256;  // float4 index inside the tile, block-contiguous order
          const int blk = f >> 6;              // 1 KB block: n-group (0..3) x m-chunk (0..3), m-chunk major
          const int mc = blk >> 2, ng = blk & 3;
          const int core = (f >> 3) & 7;       // 4-m core inside the block
          const int nr = f & 7;                // n inside the group
          const int n = ng * 8 + nr, mm = mc * 32 + core * 4;
          const float4 x = *reinterpret_cast<const float4*>(sT + n * kTTPitch + mm);
          *reinterpret_cast<float4*>(out_tt + ct_index(tt_rows, n0h + n, m0 + mm)) = x;
        }
      }
      if ((fl & F_RM) && !o.rm_trans && o.map_a4 <= 0) {
        // plain row-major output (the weight gradients): staged through shared memory so that a warp stores
        // 32 consecutive columns of one row (one 128-byte line) instead of one column of 32 rows (32 sectors)
        float* sR = smem + 32 * kTTPitch + 2 * kBM * 36 + 8 * 32 * 36;  // [128][33], behind the dW_0 buffers
#pragma unroll
        for (int j = 0; j < kEN; ++j) sR[row * 33 + cn0 + j] = v[j];
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        const int we = (tid - 64) >> 5;
        const int col = n0h + lane;
        if (col < o.rm_n) {
          float* dst = o.rm + col;
          const int ld = o.rm_ld;
#pragma unroll 4
          for (int r = 0; r < 16; ++r) {
            const int mr = we * 16 + r;
            if (m0 + mr < o.rm_m) dst[static_cast<size_t>(m0 + mr) * ld] = sR[mr * 33 + lane];
          }
        }
      } else if (fl & F_RM) {
        if (!o.rm_trans) {
          if (m < o.rm_m) {
#pragma unroll
            for (int j = 0; j < kEN; ++j) {
              const int n = nb + j;
              int col = n;
              bool ok = n < o.rm_n;
              if (o.map_a4 > 0) {  // [action | pad4 | state] -> [state | action]
                if (n < o.map_a) col = o.map_s + n;
                else if (n >= o.map_a4) col = n - o.map_a4;
                else ok = false;
                ok = ok && (n < o.map_a4 + o.map_s);
              }
              if (ok) o.rm[static_cast<size_t>(m) * o.rm_ld + col] = v[j];
            }
          }
        } else {
          if (m < o.rm_m) {
#pragma unroll
            for (int j = 0; j < kEN; ++j)
              if (nb + j < o.rm_n) o.rm[static_cast<size_t>(nb + j) * o.rm_ld + m] = v[j];
          }
        }
      }
      if (fl & F_COLSUM) {
#pragma unroll
        for (int j = 0; j < kEN; ++j) {
          float s = v[j];
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          if (lane == j) cs_smem[q * 32 + cn0 + j] = s;
        }
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        if (warp == 2) {
          const float s = cs_smem[lane] + cs_smem[32 + lane] + cs_smem[64 + lane] + cs_smem[96 + lane];
          o.colsum[static_cast<size_t>(mt) * o.colsum_ld + n0h + lane] = s;
          if (o.colsum_out && !(fl & F_DW0)) {  // (with a fused dW_0 its arrival ticket serves both totals)
            // deterministic cross-CTA total: the last M tile to arrive sums all partials in order
            // (release/acquire ticket by lane 0; the warp barrier orders the other lanes' stores before it)
            const int mtiles = o.M / kBM;
            __syncwarp();
            unsigned int ticket = 0;
            if (lane == 0) ticket = ptx::atom_add_acq_rel_gpu(o.colsum_cnt + (n0h / kBN), 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket == static_cast<unsigned int>(mtiles - 1)) {
              float tot = 0.f;
              for (int i0 = 0; i0 < mtiles; i0 += 8) {  // up to 8 tiles in flight
                float ld8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  ld8[i] = (i0 + i < mtiles) ? __ldcg(o.colsum + static_cast<size_t>(i0 + i) * o.colsum_ld + n0h + lane) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) tot += ld8[i];
              }
              if (n0h + lane < o.colsum_n) o.colsum_out[n0h + lane] = tot;
              if (lane == 0) o.colsum_cnt[n0h / kBN] = 0u;
            }
          }
        }
      }
      if (fl & F_DW0) {
        // ---- fused layer-0 weight gradient (see GemmOp::dw0_*): P[n][k] = sum_m D(m, n) X(m, k) over this
        // CTA's 128 rows as a register-tiled fp32 product.  Warp w sums rows 16w..16w+15; inside a warp
        // each lane owns a 4 n x 8 k block (3 LDS.128 per 32 FFMA); the 8 per-warp partials are added
        // through shared memory, the per-CTA result goes to global, the last M tile adds the tiles up.
        if (prof && tid == 64) prof[13] = clock64();
        float* sD = smem + 32 * kTTPitch;     // [128 m][36]: this tile's D, row-major
        float* Xs = sD + kBM * 36;            // [128 m][36]: the X chunk, row-major
        float* Ps = Xs + kBM * 36;            // [8 warps][32 n][36]: per-warp partial sums (k halves swizzled)
        uint32_t* last_flag = reinterpret_cast<uint32_t*>(ctl + 896);
        const int te = tid - 64, we = te >> 5;
        const int nq = lane >> 2, ko = lane & 3;
        const int kp = o.dw0_kp;
        const int ones = o.dw0_ones;
        const int nt = n0h >> 5;
        const int ntn32 = o.N / kBN;  // partial sums are kept per 32-column sub-tile
        float* part = o.dw0_part + (static_cast<size_t>(mt) * ntn32 + nt) * 32 * kp;
#pragma unroll
        for (int j4 = 0; j4 < kEN / 4; ++j4)
          *reinterpret_cast<float4*>(sD + row * 36 + cn0 + 4 * j4) =
              make_float4(v[4 * j4 + 0], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        for (int kc0 = 0; kc0 < (kp >> 5); ++kc0) {
          if (kc0 > 0 || (h > 0 && kp > 32)) {  // (xr still holds chunk 0 unless an earlier pass moved on)
            const float4* xg = reinterpret_cast<const float4*>(
                o.dw0_x + (static_cast<size_t>(kc0) * (o.M >> 3) + (m0 >> 3)) * 256);
#pragma unroll
            for (int i = 0; i < kAFloats / 4 / 256; ++i) xr[i] = __ldg(xg + te + i * 256);
          }
#pragma unroll
          for (int i = 0; i < kAFloats / 4 / 256; ++i) {
            const int f = te + i * 256;  // float4 index inside the CT32 chunk: [row group][k core][row]
            const int m = (f >> 6) * 8 + (f & 7), kc = (f >> 3) & 7;
            if (ones >= 0 && (ones >> 2) == kc0 * 8 + kc) {
              xr[i].x = (ones & 3) == 0 ? 1.f : xr[i].x;
              xr[i].y = (ones & 3) == 1 ? 1.f : xr[i].y;
              xr[i].z = (ones & 3) == 2 ? 1.f : xr[i].z;
              xr[i].w = (ones & 3) == 3 ? 1.f : xr[i].w;
            }
            *reinterpret_cast<float4*>(Xs + m * 36 + 4 * kc) = xr[i];
          }
          asm volatile("bar.sync 2, 256;\n" ::: "memory");
          float acc[4][8];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 4
          for (int mm = 0; mm < 16; ++mm) {
            const int ml = we * 16 + mm;
            const float4 d4 = *reinterpret_cast<const float4*>(sD + ml * 36 + 4 * nq);
            const float4 x0 = *reinterpret_cast<const float4*>(Xs + ml * 36 + 8 * ko);
            const float4 x1 = *reinterpret_cast<const float4*>(Xs + ml * 36 + 8 * ko + 4);
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(dv[a], xv[b], acc[a][b]);
          }
          // float4 stores; the two 4-k halves of odd n-quads are swapped so a quarter warp covers all banks
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              *reinterpret_cast<float4*>(Ps + we * (32 * 36) + (4 * nq + a) * 36 + 8 * ko + 4 * (h ^ (nq & 1))) =
                  make_float4(acc[a][4 * h + 0], acc[a][4 * h + 1], acc[a][4 * h + 2], acc[a][4 * h + 3]);
          asm volatile("bar.sync 2, 256;\n" ::: "memory");
#pragma unroll
          for (int i = 0; i < (kBN * kBK) / 256; ++i) {
            const int oo = te + i * 256;
            const int n = oo >> 5, k = oo & 31;  // lanes = consecutive k
            float tot = 0.f;
            const int ks = k ^ (((n >> 2) & 1) << 2);
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) tot += Ps[w8 * (32 * 36) + n * 36 + ks];
            part[static_cast<size_t>(n) * kp + kc0 * 32 + k] = tot;
          }
          // (the next chunk's Xs stores are ordered behind every warp's reads by the two barriers above)
        }
        // deterministic cross-CTA totals (dW_0 and, if present, the bias column sums): the last M tile
        // to arrive adds the per-tile partials in tile order -- unless the consumer does (dw0_defer)
        const int mtiles = o.M / kBM;
        if (prof && tid == 64) prof[14] = clock64();
        if (o.dw0_defer) continue;
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        if (te == 0)
          *last_flag = (ptx::atom_add_acq_rel_gpu(o.dw0_cnt + nt, 1u) == static_cast<unsigned int>(mtiles - 1)) ? 1u : 0u;
        asm volatile("bar.sync 2, 256;\n" ::: "memory");
        if (prof && tid == 64) prof[15] = clock64();
        if (*last_flag) {
          const float* p0 = o.dw0_part + static_cast<size_t>(nt) * 32 * kp;
          const size_t mt_stride = static_cast<size_t>(ntn32) * 32 * kp;
          const int ma = o.dw0_map_a, ma4 = o.dw0_map_a4, ms = o.dw0_map_s, cols = o.dw0_cols, nv = o.dw0_n - n0h;
          float* outp = o.dw0_out + static_cast<size_t>(n0h) * o.dw0_ld;
          float* bias_out = o.dw0_bias_out;
          const int ld = o.dw0_ld;
          // 32 x kp outputs, 4 per thread per round; every tile's partial of a round is in flight
          // together (up to 8 M tiles; more are added in further passes of the same order)
          for (int base = 0; base < 32 * kp; base += 1024) {
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i0 = 0; i0 < mtiles; i0 += 8) {
              float ld4[8][4];
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  ld4[i][j] = (i0 + i < mtiles) ? __ldcg(p0 + (i0 + i) * mt_stride + base + te + 256 * j) : 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) tot[j] += ld4[i][j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int oo = base + te + 256 * j;
              const int n = oo / kp, kk = oo - n * kp;
              int col = kk;
              bool ok = kk < cols && n < nv;
              if (ma4 > 0) {  // [action | pad4 | state] -> [state | action]
                if (kk < ma) col = ms + kk;
                else if (kk >= ma4) col = kk - ma4;
                else ok = false;
                ok = ok && (kk < ma4 + ms);
              }
              if (ok) outp[static_cast<size_t>(n) * ld + col] = tot[j];
              if (kk == ones && n < nv) bias_out[n0h + n] = tot[j];
            }
          }
          if ((fl & F_COLSUM) && o.colsum_out && we == 0) {
            float tot = 0.f;
            for (int i = 0; i < mtiles; ++i)
              tot += __ldcg(o.colsum + static_cast<size_t>(i) * o.colsum_ld + n0h + lane);
            if (n0h + lane < o.colsum_n) o.colsum_out[n0h + lane] = tot;
          }
          if (te == 0) o.dw0_cnt[nt] = 0u;
        }
      }
    }
  }

  if (prof && tid == 64) prof[7] = clock64();
  // ---- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (!kSimt && warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(*tmem_slot, tmem_cols);
  }
  if (prof && tid == 32) {
    prof[8] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    prof[10] = static_cast<long long>(gt);
  }
}

}  // namespace oprl
