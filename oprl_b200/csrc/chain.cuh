// Batch-slice chain kernel: a whole dependency chain of MLP layers in ONE launch.
//
// The stage path (gemm.cuh) tiles every layer over output features and pays a kernel boundary
// (~3.8 us with programmatic dependent launch) per layer: 15 launches per DDPG update, 57 of 92 us.
// Here the roles of the operands are swapped: the WEIGHTS are the M side of the MMA (128 output
// features per instruction, A operand streamed L2 -> registers -> tf32 hi/lo split -> TMEM by eight
// feeder warps) and a SLICE OF THE BATCH (kNB = 16 rows) is the N side, kept in shared memory as the
// K-major B operand.  One CTA therefore carries its 16 batch rows through every layer of every
// network of the step -- forward, loss, backward dX chain -- without ever talking to another CTA:
// the output of a layer comes back from TMEM with thread = feature, is biased / rectified / split
// and written straight into the shared-memory operand of the next layer.  Cross-CTA traffic is
// left to what genuinely contracts over the batch: the weight gradients (one grouped GEMM launch
// afterwards, fed by the transposed activations / deltas this kernel stores: thread = feature is
// exactly the [feature x batch] layout dW needs) and a handful of bias-gradient / loss partial sums
// (per-CTA partials, fixed-order total by the last CTA to arrive).
//
//   D^T[feat x 16] (+)= W[feat x k] * H^T[k x 16]        tcgen05.mma kind::tf32, M = 128, N = 16 / 32
//
// Arithmetic is the same 3xTF32 scheme with cut accumulation chains as gemm.cuh (cross terms apart from
// the hi*hi terms, one accumulator pair per group of K chunks, fp32 adds in the epilogue) -- but issued as
// TWO MMAs per K step instead of three: with N this small an MMA costs what it takes to read its 4 KB A tile
// out of tensor memory (~40 cycles measured; 8 cycles of math), so A_hi is read once against the
// concatenated rows [H_hi ; H_lo] (N = 32 -> hi*hi and hi*lo side by side) and A_lo once against H_hi.  Narrow layers (<= 8 outputs: tanh policy heads, scalar Q heads, the action columns
// of dX) never touch the tensor core: they are dot products over the feature axis done in the
// epilogue of the layer that produces their input (fp32 FMA, warp butterfly + fixed-order
// cross-warp sum), and K <= 8 layers (the policy head backward) are a few FMAs per thread.
//
// Warp roles (18 warps; a sub-partition of the SM hosts at most five, so 96 registers per thread):
//   0-1   MMA issuers, one per M tile (128 output features) of the layer.  A single warp issuing every MMA is
//         the bottleneck of this design: the ~100 scalar instructions between two chunks (barrier wait, fence,
//         elect, descriptor words into uniform registers, 12 MMAs, commit) are serial latency, ~500 cycles per
//         chunk against ~130 of tensor time.  The two M tiles have separate accumulators, so two warps can
//         issue side by side without changing any accumulation order; the chunk stream interleaves the tiles.
//   2-9   weight feeders, two per TMEM lane quarter alternating 16 KB chunks: L2 -> registers (one chunk
//         prefetched) -> tf32 hi/lo split -> tcgen05.st, sixteen K columns at a time (register budget).
//   10-17 epilogue, one per (lane quarter, M tile): a thread owns one feature row of the layer output.
// The weight stream runs ahead of the MMAs through a ring of TMEM slots, across layer boundaries; the
// accumulators are double-buffered so the read-out of one layer overlaps the MMAs of the next.
//
// Reference semantics: ddpg.py:86-107, td3.py:95-141 (see engine.cu build_chain_ddpg_td3).
#pragma once
#include "kernels.cuh"

namespace oprl {

constexpr int kNB = 16;                  // batch rows (MMA N) per CTA
constexpr int kFeedGroups = 2;            // feeder groups of four warps (one per TMEM lane quarter); chunk g belongs to group g % kFeedGroups
                                         // (3 groups = 22 warps at 80 registers measured no faster: the slot hand-shake, not the feeders' work, paces the ring)
constexpr int kChainWarps = 2 + 4 * kFeedGroups + 8;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kChainThreads = kChainWarps * 32;
constexpr int kCorePitch = 144;          // bytes between K-adjacent 8x16B core matrices of an operand buffer
                                         // (128 + 16: the feature-per-lane scalar stores hit 32 distinct banks)
constexpr int kCSlots = 6;               // most TMEM ring slots of split A chunks (64 columns each; ChainLaunch::n_slots are used)
constexpr int kCAcc = 5;                 // accumulator slots per M tile: cross terms + up to 4 hi*hi groups
constexpr int kCMaxOps = 16;
constexpr int kCMaxChunks = 192;
constexpr int kCMaxBufs = 12;
constexpr int kCMaxVec = 8;
constexpr int kCMaxJ = 6;               // widest narrow head (action dimensions) done in an epilogue
constexpr int kCMaskSlots = 4;
constexpr int kCFeat = 256;              // widest layer (2 M tiles)

enum ChainFlag : int {
  CF_BIAS_RELU = 1,    // v = max(acc + bias[f], 0)
  CF_SAVE_MASK = 2,    // remember v > 0 (bit n of mask[mask_slot][f])
  CF_APPLY_MASK = 4,   // v = mask bit ? acc : 0   (ReLU backward)
  CF_OUT_SMEM = 8,     // write the result as the next layer's operand (tf32 hi / lo)
  CF_OUT_GLOBAL = 16,  // store the result transposed-tiled [feature x batch] for the weight-gradient GEMMs
  CF_MASK_GLOBAL = 32, // also store the mask bits to global memory (consumed by a later chain launch)
  CF_COLSUM = 64,      // per-CTA bias-gradient partial: sum over this CTA's batch rows
};
enum ChainHead : int {
  CH_NONE = 0,
  CH_ACTION = 1,   // a = tanh(W3 h + b3) [+ noise, clamp] -> action rows of an operand buffer / global
  CH_QTARGET = 2,  // target critic head: qn = w3 . h + b3
  CH_QLOSS = 3,    // online critic head + TD target + MSE seed + dz of this layer + head gradients
  CH_QACTOR = 4,   // critic head of the actor step: q (logged) + dz of this layer for the constant seed
  CH_DXA = 5,      // action columns of dX, tanh', then the policy head backward (K = A) -> dz of the actor's last hidden layer
};

struct ChainOp {
  const float* w;       // A operand: CT32 weights [w_rows x 32 * kchunks]
  const float* bias;
  float* gout;          // CF_OUT_GLOBAL: CT32 [gout_rows x Bp]
  unsigned int* gmask;  // CF_MASK_GLOBAL: [n_cta][256]
  const float* hw;      // head weights (CH_ACTION: [J x hw_ld] ; Q heads: [H] ; CH_DXA: critic W0 + S, row stride hw_ld)
  const float* hb;      // head bias
  const float* aux;     // CH_ACTION: noise [B x J] (nullable) ; CH_DXA: tanh(pi) values [Bp x J]
  const float* w2;      // CH_DXA: actor last-layer weights [J x Ha]
  float* hout;          // CH_ACTION: row-major [Bp x J] (nullable) ; CH_DXA: CT32 [128 x Bp] (dz of the policy head, transposed)
  float* hout2;         // CH_ACTION: CT32 [Bp x ..] input matrix whose action columns get the result (nullable)
  int w_rows, mtiles, kchunks;
  int group;            // K chunks per hi*hi accumulator (2: chains of 8 MMAs; 4: chains of 16, two accumulators fewer to read out)
  int in_hi, in_lo, in_sbo, in_bar, in_phase;   // input operand buffer (byte offsets into dynamic smem)
  int out_hi, out_lo, out_sbo, out_bar;         // CF_OUT_SMEM / head output buffer
  int x_hi, x_lo, x_sbo, x_bar;                 // CH_ACTION: operand buffer receiving the action rows (x_bar < 0: none)
  int flags, head, J, hw_ld;
  int mask_slot, mask2_slot;
  int gout_rows;
  int vec_slot, vec2_slot;  // per-CTA partial vectors (CF_COLSUM / head gradients), -1 = none
  int crit;
  int Ha;
  float clamp;
};

struct ChainInput {   // tiled [Bp x 32 * kchunks] matrix whose 16-row slice becomes an operand buffer
  const float* src;
  int rows, kchunks;
  int hi, lo, sbo, bar;
};

struct ChainMmaOp {  // what the MMA-issuing warp needs of one op, in kernel-parameter (constant) space so that its
                     // loop runs on the uniform datapath
  uint32_t in_hi, in_lo;  // byte offsets of the input operand buffer (hi / lo halves) in dynamic shared memory
  uint32_t dw_hi;         // descriptor high word: SBO >> 4 | version 1 << 14
  uint8_t in_bar, in_phase, mtiles, kchunks;
  uint8_t group, pad[3];
};

struct ChainLaunch {
  const ChainOp* ops;
  ChainMmaOp mop[kCMaxOps];
  int n_ops;
  int B, Bp, n_cta;
  int n_in;
  ChainInput in[3];
  // ReLU masks of layers whose forward ran in an EARLIER launch (the actor's pi(s) pass): rebuilt at kernel start
  // from the saved transposed activations (CT32 [gm_rows x Bp], h > 0)
  const float* gm_src[2];
  int gm_rows[2];
  int gm_slot[2];
  int n_gm;
  const float* r;
  const float* d;
  float* part;      // [n_cta][part_stride]: vectors (256 each) then 32 tail scalars
  int part_stride;
  unsigned int* counter;
  int n_vec;
  float* vec_dst[kCMaxVec];
  int vec_n[kCMaxVec];
  int kind;         // 0: critic step (loss scalars, gb3) ; 1: actor step (actor loss, policy-head bias gradient)
  int nq;
  float gamma, inv_count;
  float* gb3[2];
  float* gb_head;   // kind 1: bias gradient of the policy head [J]
  int J;
  DevState* st;
  int bump, bump_actor;
  // TMEM plan: two accumulator regions of d_cols columns (op i uses region i & 1, so the read-out of one
  // op overlaps the MMAs of the next), then n_slots x 64 columns of split A chunks from column a_col0
  int d_cols, a_col0, n_slots;
  int region_mask;  // 1: two accumulator regions (op i uses region i & 1); 0: one region, more ring slots
  long long* prof;  // selftest / profiling: clock64 stamps of CTA 0 (nullable)
  int pitch;        // bytes between K-adjacent core matrices of the operand buffers (kCorePitch)
  int debug;        // timing experiments only: 2 = no weight loads, 4 = no tcgen05.st (results invalid); 8 = waiting warps poll in a tight loop instead of backing off with nanosleep; 16 = no per-chunk commit, 32 = feeders never wait for a free slot, 64 = MMA warps never wait for a filled slot (48 / 112: handshake cost probes)
};

struct ChainCtl {
  uint64_t a_full[kCSlots], a_empty[kCSlots];
  uint64_t d_full[2], d_free[2];
  uint64_t buf_bar[kCMaxBufs];
  uint32_t tmem_slot;
  uint32_t last_flag;
  const float* chunk_src[kCMaxChunks];
  ChainOp ops[kCMaxOps];
  uint32_t mask[kCMaskSlots][kCFeat];
  float red[2][kEpiWarps][kCMaxJ][kNB];
  float qn_s[2][kNB];
  float dq_s[2][kNB];
  float rr[kNB], dd[kNB];
  float dza_s[kCMaxJ][kNB];
  float tail_s[32];
};
constexpr int kChainCtlBytes = ((sizeof(ChainCtl) + 1023) / 1024) * 1024;
constexpr int kChainSmemMax = 232448;

// operand buffer geometry: 16 rows = 2 groups of 8; group stride (SBO) = K/4 cores x kCorePitch
__host__ __device__ __forceinline__ int chain_buf_sbo(int K) { return (K / 4) * kCorePitch; }
__host__ __device__ __forceinline__ int chain_buf_bytes(int K) { return 2 * chain_buf_sbo(K); }  // one of hi / lo

namespace ptx {
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// registers -> TMEM: thread i writes lane (lane_base + i), 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// registers -> TMEM: thread i writes lane (lane_base + i), 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
}  // namespace ptx

// Sum over the 32 lanes of a warp of 16 values per lane: a butterfly that halves the value count at
// every step (8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5).  Lane l ends up with the total of
// column (l >> 1) & 15.  Fixed tree -> bit-reproducible.
__device__ __forceinline__ float warp_colsum16(const float* p, int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  float r8[8], r4[4], r2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = b4 ? p[i + 8] : p[i];
    const float send = b4 ? p[i] : p[i + 8];
    r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = b3 ? r8[i + 4] : r8[i];
    const float send = b3 ? r8[i] : r8[i + 4];
    r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b2 ? r4[i + 2] : r4[i];
    const float send = b2 ? r4[i] : r4[i + 2];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = b1 ? r2[1] : r2[0];
  const float send = b1 ? r2[0] : r2[1];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
// fixed-shape sum over the 16 lanes of a half warp (all 32 lanes call); valid in the lanes with (lane & 15) == 0
__device__ __forceinline__ float half_warp_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 8);
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}

// wait with back-off: a warp that polls an mbarrier in a tight loop takes issue slots and shared-memory
// cycles from the warps it is waiting for
__device__ __forceinline__ void chain_wait(uint64_t* bar, uint32_t parity, bool backoff) {
  if (!backoff) {
    ptx::mbar_wait(bar, parity);
    return;
  }
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (++spins > (1u << 24)) {
      if ((threadIdx.x & 31) == 0) printf("oprl: chain mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// the same wait for ONE thread on a shared::cta address (the elected MMA-issuing thread; no printf in its loop)
__device__ __forceinline__ void chain_wait1(uint32_t bar_saddr, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(bar_saddr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}

#define CHAIN_EPI_BAR() asm volatile("bar.sync 1, 256;\n" ::: "memory")

// one feature row (16 batch columns) -> operand buffer (tf32 hi / lo), K-major core-matrix layout
__device__ __forceinline__ void chain_store_operand(uint8_t* csm, int off_hi, int off_lo, int sbo, int pitch, int f, const float* x) {
  uint8_t* bh = csm + off_hi + (f >> 2) * pitch + (f & 3) * 4;
  uint8_t* bl = csm + off_lo + (f >> 2) * pitch + (f & 3) * 4;
#pragma unroll
  for (int n = 0; n < kNB; ++n) {
    float h, l;
    ptx::split_tf32(x[n], h, l);
    const int off = (n >> 3) * sbo + (n & 7) * 16;
    *reinterpret_cast<float*>(bh + off) = h;
    *reinterpret_cast<float*>(bl + off) = l;
  }
}
// one feature row -> transposed-tiled global matrix [rows x Bp], columns n0 .. n0 + 15
__device__ __forceinline__ void chain_store_gout(float* g, int rows, int f, int n0, const float* x) {
  float* dst = g + ct_index(rows, f, n0);
#pragma unroll
  for (int n4 = 0; n4 < kNB / 4; ++n4)
    *reinterpret_cast<float4*>(dst + n4 * 32) = make_float4(x[4 * n4], x[4 * n4 + 1], x[4 * n4 + 2], x[4 * n4 + 3]);
}
// total of the eight epilogue warps' partials, in warp order
__device__ __forceinline__ float chain_red8(const float (*red)[kCMaxJ][kNB], int j, int n, int nw) {
  float s = red[0][j][n];
  for (int w = 1; w < nw; ++w) s += red[w][j][n];
  return s;
}

__global__ void __launch_bounds__(kChainThreads, 1) chain_kernel(const __grid_constant__ ChainLaunch L) {
  extern __shared__ __align__(1024) uint8_t csm[];
  ChainCtl& C = *reinterpret_cast<ChainCtl*>(csm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int n0 = cta * kNB;
  long long* prof = (L.prof && cta == 0) ? L.prof : nullptr;

  ptx::pdl_trigger();
  if (prof && tid == 0) prof[0] = clock64();
  // ---- static tables (host-written when the program was built): legal before the dependency wait
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(L.ops);
    uint32_t* dst = reinterpret_cast<uint32_t*>(C.ops);
    const int nw = L.n_ops * static_cast<int>(sizeof(ChainOp) / 4);
    for (int i = tid; i < nw; i += kChainThreads) dst[i] = __ldg(src + i);
  }
  if (warp == 0) {
    if (lane < kCSlots) {
      ptx::mbar_init(&C.a_full[lane], 4);   // the four quarter warps of one feeder group
      ptx::mbar_init(&C.a_empty[lane], 1);  // tcgen05.commit of the MMA warp that consumed the chunk
    } else if (lane < kCSlots + 2) {
      ptx::mbar_init(&C.d_full[lane - kCSlots], 2);          // one commit per MMA warp
      ptx::mbar_init(&C.d_free[lane - kCSlots], kEpiWarps);  // one arrival per epilogue warp
    } else if (lane >= 8 && lane < 8 + kCMaxBufs) {
      ptx::mbar_init(&C.buf_bar[lane - 8], kEpiWarps);
    }
    ptx::fence_mbar_init();
    __syncwarp();
    ptx::tmem_alloc(&C.tmem_slot, 512);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // chunk table: flat chunk index -> source of its 16 KB.  Inside an op the M tiles alternate (K chunk 0 of
  // tile 0, K chunk 0 of tile 1, K chunk 1 of tile 0, ...) so that both MMA warps always have work.
  if (tid < L.n_ops) {
    int g0 = 0;
    for (int i = 0; i < tid; ++i) g0 += C.ops[i].mtiles * C.ops[i].kchunks;
    const ChainOp& o = C.ops[tid];
    for (int c = 0; c < o.kchunks; ++c)
      for (int mt = 0; mt < o.mtiles; ++mt)
        C.chunk_src[g0 + c * o.mtiles + mt] = o.w + (static_cast<size_t>(c) * (o.w_rows >> 3) + mt * 16) * 256;
  }
  if (tid < 32) C.tail_s[tid] = 0.f;
  __syncthreads();
  int total_chunks = 0;
  for (int i = 0; i < L.n_ops; ++i) total_chunks += C.ops[i].mtiles * C.ops[i].kchunks;
  const uint32_t tmem = C.tmem_slot;
  const int n_slots = L.n_slots;
  const uint32_t a_col0 = static_cast<uint32_t>(L.a_col0);
  constexpr int pitch = kCorePitch;  // compile-time: the per-K-step descriptor advance becomes an immediate
  const bool backoff = (L.debug & 8) == 0;  // default on; OPRL_B200_CHAIN_DEBUG=8 makes every wait a tight poll

  if (warp < 2) {
    // ================================================================= MMA issuers (warp m: M tile m)
    // ONE elected thread runs the whole loop -- waits included -- inside a single elect.sync region: ptxas then
    // keeps every address, descriptor word and counter in uniform registers and the MMAs issue back to back
    // (~10 cycles each).  Electing per chunk instead left the loop on the vector datapath with an R2UR per MMA
    // operand: 35-45 cycles per MMA, measured (tools/experiments/chain_probe4.cu).
    const int m = warp;
    if (ptx::elect_one()) {
      // This thread's program is serial: every instruction between two MMAs costs its full latency (~5 cycles each,
      // nothing to overlap with), so the per-chunk path is kept to the wait, the operand words and the MMAs --
      // every launch parameter is read into a register once, the ring position advances by adds.
      const uint32_t idesc16 = ptx::idesc_tf32(128, kNB, 0, 0), idesc32 = ptx::idesc_tf32(128, 2 * kNB, 0, 0);
      constexpr uint32_t kstep = static_cast<uint32_t>(2 * pitch) >> 4;  // one K = 8 step, in the descriptor's 16-byte units
      const uint32_t smem_base16 = ptx::smem_u32(csm) >> 4;
      const uint32_t lbo_bits = (static_cast<uint32_t>(pitch) >> 4) << 16;
      const uint32_t full0 = ptx::smem_u32(&C.a_full[0]), empty0 = ptx::smem_u32(&C.a_empty[0]);
      const uint32_t ta0 = tmem + a_col0;
      const uint32_t nsl = static_cast<uint32_t>(n_slots);
      const uint32_t d_cols = static_cast<uint32_t>(L.d_cols);
      const int n_ops = L.n_ops;
      const bool dbg_nowait = (L.debug & 64) != 0, dbg_nocommit = (L.debug & 16) != 0;  // timing probes only
      uint32_t slot0 = 0, par0 = 0;  // ring position / phase parity of the op's first chunk (all chunks, both warps)
      for (int oi = 0; oi < n_ops; ++oi) {
        const ChainMmaOp o = L.mop[oi];
        const uint32_t r = static_cast<uint32_t>(oi & L.region_mask);
        const int k = L.region_mask ? (oi >> 1) : oi;  // use count of the accumulator region
        chain_wait1(ptx::smem_u32(&C.buf_bar[o.in_bar]), static_cast<uint32_t>(o.in_phase & 1));
        if (k > 0) chain_wait1(ptx::smem_u32(&C.d_free[r]), static_cast<uint32_t>((k - 1) & 1));
        ptx::tc_fence_after();
        if (prof && m == 0 && oi < 16) prof[16 + oi] = clock64();
        const uint32_t mtiles = o.mtiles, kchunks = o.kchunks, group = o.group;
        if (static_cast<uint32_t>(m) < mtiles) {
          // descriptor words: low = start >> 4 | LBO >> 4 << 16, high = SBO >> 4 | version 1 << 14
          const uint32_t dw_hi = o.dw_hi;
          uint32_t dl = ((smem_base16 + (o.in_hi >> 4)) & 0x3FFFu) | lbo_bits;
          // accumulators of one M tile: per group of K chunks a pair of column blocks [hi*hi | cross terms]
          const uint32_t n_grp = (kchunks + group - 1) / group;
          uint32_t dgrp = tmem + r * d_cols + static_cast<uint32_t>(m) * n_grp * 2 * kNB;
          uint32_t s = slot0 + static_cast<uint32_t>(m), par = par0;
          if (s >= nsl) {
            s -= nsl;
            par ^= 1u;
          }
          // The barrier poll costs ~140 cycles of latency (mbarrier.try_wait round trip, chain_probe5): the state of
          // the NEXT own chunk's slot is sampled (non-blocking test_wait) before this chunk's MMAs are issued.
          uint32_t ok = 0, left = kchunks;
          for (uint32_t g = 0; g < n_grp; ++g, dgrp += 2 * kNB) {
            const uint32_t in_g = left < group ? left : group;
            left -= in_g;
            for (uint32_t cc = 0; cc < in_g; ++cc) {
              if (!ok && !dbg_nowait) chain_wait1(full0 + s * 8u, par);
              ptx::tc_fence_after();
              if (prof && m == 0 && oi == 4) prof[160 + 2 * (g * group + cc)] = clock64();
              uint32_t ns = s + mtiles, npar = par;
              if (ns >= nsl) {
                ns -= nsl;
                npar ^= 1u;
              }
              ok = ptx::mbar_test_wait_addr(full0 + ns * 8u, npar);  // (past the op's last chunk: a harmless early look)
              const uint32_t ta_hi = ta0 + s * 64u;
              // A_hi . [B_hi ; B_lo] (N = 32: the lo rows follow the hi rows in the operand buffer) -> [hi*hi | hi*lo];
              // A_lo . B_hi -> the cross-term block
              ptx::mma_tf32_ts2(dgrp, ta_hi, dl, dw_hi, idesc32, cc);
              ptx::mma_tf32_ts2(dgrp + kNB, ta_hi + 32u, dl, dw_hi, idesc16, 1u);
#pragma unroll
              for (uint32_t j = 1; j < 4; ++j) {
                ptx::mma_tf32_ts2(dgrp, ta_hi + 8u * j, dl + j * kstep, dw_hi, idesc32, 1u);
                ptx::mma_tf32_ts2(dgrp + kNB, ta_hi + 32u + 8u * j, dl + j * kstep, dw_hi, idesc16, 1u);
              }
              if (!dbg_nocommit) ptx::mma_commit_addr(empty0 + s * 8u);
              if (prof && m == 0 && oi == 4) prof[161 + 2 * (g * group + cc)] = clock64();
              dl += 4 * kstep;
              s = ns;
              par = npar;
            }
            ok = (left > 0) ? ok : 0u;
          }
        }
        ptx::mma_commit_addr(ptx::smem_u32(&C.d_full[r]));  // (a warp that issued nothing for this op arrives at once)
        // ring position of the next op's first chunk
        slot0 += mtiles * kchunks;
        while (slot0 >= nsl) {
          slot0 -= nsl;
          par0 ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + 4 * kFeedGroups) {
    // ================================================================= weight feeders
    const int q = warp & 3, half = (warp - 2) >> 2;  // half: this warp's feeder group
    const int row = q * 32 + lane;
    const uint32_t ta_lane = tmem + (static_cast<uint32_t>(q * 32) << 16) + a_col0;
    const int roff = (row >> 3) * 64 + (row & 7);  // float4 index of this row inside a chunk (+ 8 per k core)
    ptx::pdl_wait();  // the weights are the previous launch's (Adam) output
    if (L.debug & 128) total_chunks = 0;  // (timing probe: no feeders at all)
    long long t_wait = 0, t_work = 0;
    // One chunk of this half is prefetched in registers; the split goes sixteen K columns at a time (register budget).
    // Measured on the way here (tools/chain_prof.py): tcgen05.st x8 instead of x16 costs +250 cycles per chunk (the
    // tensor-memory stores are slow while MMAs run), a second prefetched chunk buys nothing (the ring of TMEM slots,
    // not the L2 latency, paces the feeders).
    float4 nx[8];
    if (half < total_chunks) {
      const float4* sp = reinterpret_cast<const float4*>(C.chunk_src[half]) + roff;
#pragma unroll
      for (int j = 0; j < 8; ++j) nx[j] = __ldg(sp + j * 8);
    }
    int slot = half % n_slots, wraps = half / n_slots;
    const uint32_t empty0 = ptx::smem_u32(&C.a_empty[0]);
    uint32_t free_seen = 0;  // the slot of the upcoming chunk was already seen released (sampled one chunk ahead)
    for (int g = half; g < total_chunks; g += kFeedGroups) {
      const long long t0 = prof ? clock64() : 0;
      if (wraps > 0) {
        // (tight poll: the feeders are the pace setters -- a back-off sleep here was ~300 cycles per chunk)
        if (!free_seen) chain_wait(&C.a_empty[slot], static_cast<uint32_t>((wraps - 1) & 1), (L.debug & 256) != 0);
        ptx::tc_fence_after();
      }
      const long long t1 = prof ? clock64() : 0;
      const uint32_t ta = ta_lane + static_cast<uint32_t>(slot * 64);
      // ring position of this group's next chunk; its release is sampled now (non-blocking), the ~140-cycle
      // round trip of the poll overlaps the split below.  (Safe only because n_slots is a multiple of kFeedGroups:
      // a group then always revisits ITS OWN slots, so the phase it asks about is at most one behind -- with slots
      // shared between groups the parity test can alias a phase two uses back and release a slot still being read.)
      int nslot = slot + kFeedGroups, nwraps = wraps;
      while (nslot >= n_slots) {
        nslot -= n_slots;
        nwraps += 1;
      }
      free_seen = (nwraps > 0 && g + kFeedGroups < total_chunks)
                      ? ptx::mbar_test_wait_addr(empty0 + static_cast<uint32_t>(nslot) * 8u, static_cast<uint32_t>((nwraps - 1) & 1))
                      : 0u;
#pragma unroll
      for (int j4 = 0; j4 < 2; ++j4) {  // sixteen K columns at a time: hi -> columns [16 j4, +16), lo -> 32 + the same
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ptx::split_tf32(nx[4 * j4 + j].x, hi[4 * j + 0], lo[4 * j + 0]);
          ptx::split_tf32(nx[4 * j4 + j].y, hi[4 * j + 1], lo[4 * j + 1]);
          ptx::split_tf32(nx[4 * j4 + j].z, hi[4 * j + 2], lo[4 * j + 2]);
          ptx::split_tf32(nx[4 * j4 + j].w, hi[4 * j + 3], lo[4 * j + 3]);
        }
        if (!(L.debug & 4)) {
          ptx::tmem_st16(ta + 16u * j4, hi);
          ptx::tmem_st16(ta + 32u + 16u * j4, lo);
        } else {
          asm volatile("" ::"f"(hi[0] + lo[7] + hi[3] + lo[2]));
        }
      }
      if (g + kFeedGroups < total_chunks && !(L.debug & 2)) {
        const float4* sp = reinterpret_cast<const float4*>(C.chunk_src[g + kFeedGroups]) + roff;
#pragma unroll
        for (int j = 0; j < 8; ++j) nx[j] = __ldg(sp + j * 8);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&C.a_full[slot]);
      if (prof) {
        const long long t2 = clock64();
        t_wait += t1 - t0;
        t_work += t2 - t1;
        if (warp == 2 && lane == 0 && g >= 22 && g < 38) {
          prof[96 + 3 * ((g - 22) >> 1)] = t0;
          prof[97 + 3 * ((g - 22) >> 1)] = t1;
          prof[98 + 3 * ((g - 22) >> 1)] = t2;
        }
        if (warp < 6 && lane == 0 && g >= 22 && g < 38) prof[200 + 4 * ((g - 22) >> 1) + (warp - 2)] = t2;  // quarter-warp skew
      }
      slot = nslot;
      wraps = nwraps;
    }
    if (prof && warp == 2 && lane == 0) {
      prof[56] = t_wait;   // cycles this feeder waited for a free TMEM slot (MMA / epilogue bound)
      prof[57] = t_work;   // cycles in split + tcgen05.st + wait + arrive
      prof[58] = total_chunks;
    }
  } else {
    // ================================================================= epilogue warps
    const int e = warp - (2 + 4 * kFeedGroups);  // 0..7
    const int q = warp & 3, mt = e >> 2;
    const int f = mt * 128 + q * 32 + lane;  // feature row of the layer output this thread owns
    const int et = e * 32 + lane;            // 0..255
    ptx::pdl_wait();
    if (L.bump && cta == 0 && et == 0) bump_counters(L.st, L.bump_actor);
    // ---- operand buffers that come from global memory (the gathered batch), reward / done, masks
    for (int ii = 0; ii < L.n_in; ++ii) {
      const ChainInput& in = L.in[ii];
      const int nf4 = in.kchunks * 128;  // float4s: 16 rows x 32 k per chunk
      for (int i = et; i < nf4; i += kEpiThreads) {
        const int c = i >> 7, w = i & 127;
        const int grp = w >> 6, j = (w >> 3) & 7, rr8 = w & 7;
        const float4 x = __ldg(reinterpret_cast<const float4*>(
                                   in.src + (static_cast<size_t>(c) * (in.rows >> 3) + (n0 >> 3) + grp) * 256) + j * 8 + rr8);
        float4 h, l;
        ptx::split_tf32(x.x, h.x, l.x);
        ptx::split_tf32(x.y, h.y, l.y);
        ptx::split_tf32(x.z, h.z, l.z);
        ptx::split_tf32(x.w, h.w, l.w);
        const int off = grp * in.sbo + (c * 8 + j) * pitch + rr8 * 16;
        *reinterpret_cast<float4*>(csm + in.hi + off) = h;
        *reinterpret_cast<float4*>(csm + in.lo + off) = l;
      }
    }
    if (et < kNB) {
      const int mrow = min(n0 + et, L.B - 1);
      C.rr[et] = L.r ? __ldg(L.r + mrow) : 0.f;
      C.dd[et] = L.d ? __ldg(L.d + mrow) : 0.f;
    }
    for (int k = 0; k < L.n_gm; ++k) {
      uint32_t bits = 0;
      if (et < L.gm_rows[k]) {
        const float4* src = reinterpret_cast<const float4*>(L.gm_src[k] + ct_index(L.gm_rows[k], et, n0));
#pragma unroll
        for (int n4 = 0; n4 < kNB / 4; ++n4) {
          const float4 h4 = __ldg(src + n4 * 8);
          bits |= (h4.x > 0.f ? 1u : 0u) << (4 * n4) | (h4.y > 0.f ? 2u : 0u) << (4 * n4) | (h4.z > 0.f ? 4u : 0u) << (4 * n4) |
                  (h4.w > 0.f ? 8u : 0u) << (4 * n4);
        }
      }
      C.mask[L.gm_slot[k]][et] = bits;
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (int ii = 0; ii < L.n_in; ++ii) ptx::mbar_arrive(&C.buf_bar[L.in[ii].bar]);
    CHAIN_EPI_BAR();  // rr / dd / masks visible to every epilogue thread
    if (prof && et == 0) prof[1] = clock64();

    float* mypart = L.part + static_cast<size_t>(cta) * L.part_stride;
    int hcount = 0;
    const int hj = et >> 4, hn = et & 15;  // head finisher thread -> (output j, batch column n)
    const bool col_valid = n0 + hn < L.B;
    for (int oi = 0; oi < L.n_ops; ++oi) {
      const ChainOp& o = C.ops[oi];
      const bool act = mt < o.mtiles;  // this thread owns a row of the output
      const int flags = o.flags;
      const int head = o.head;
      const int J = (head == CH_ACTION || head == CH_DXA) ? o.J : 1;
      const int nw = 4 * o.mtiles;  // epilogue warps that hold a partial
      // ---- everything that comes from global memory is requested before the accumulators are waited for
      float bv = 0.f;
      uint32_t mbits = 0u;
      float hwv[kCMaxJ];
#pragma unroll
      for (int j = 0; j < kCMaxJ; ++j) hwv[j] = 0.f;
      if (act) {
        if (flags & CF_BIAS_RELU) bv = __ldg(o.bias + f);
        if (flags & CF_APPLY_MASK) mbits = C.mask[o.mask_slot][f];
        if (head == CH_ACTION) {
#pragma unroll
          for (int j = 0; j < kCMaxJ; ++j)
            if (j < J) hwv[j] = __ldg(o.hw + static_cast<size_t>(j) * o.hw_ld + f);
        } else if (head == CH_DXA) {
#pragma unroll
          for (int j = 0; j < kCMaxJ; ++j)
            if (j < J) hwv[j] = __ldg(o.hw + static_cast<size_t>(f) * o.hw_ld + j);
        } else if (head != CH_NONE) {
          hwv[0] = __ldg(o.hw + f);
        }
      }
      float fin_b = 0.f, fin_aux = 0.f;  // head bias / per-(row, output) extra of the finisher threads
      if (head == CH_ACTION) {
        if (et < J * kNB) {
          fin_b = __ldg(o.hb + hj);
          if (o.aux) fin_aux = __ldg(o.aux + static_cast<size_t>(min(n0 + hn, L.B - 1)) * J + hj);
        }
      } else if (head == CH_DXA) {
        if (et < J * kNB) fin_aux = __ldg(o.aux + static_cast<size_t>(n0 + hn) * J + hj);
      } else if (head != CH_NONE) {
        fin_b = __ldg(o.hb);
      }
      float w2v[kCMaxJ];
      uint32_t m2 = 0u;
      if (head == CH_DXA) {
#pragma unroll
        for (int j = 0; j < kCMaxJ; ++j) w2v[j] = (j < J && f < o.Ha) ? __ldg(o.w2 + static_cast<size_t>(j) * o.Ha + f) : 0.f;
        if (f < o.Ha) m2 = C.mask[o.mask2_slot][f];
      }
      // ---- accumulators -> registers (hi*hi groups in order, then the cross terms), release the region
      const int r = oi & L.region_mask;
      chain_wait(&C.d_full[r], static_cast<uint32_t>((L.region_mask ? (oi >> 1) : oi) & 1), backoff);
      ptx::tc_fence_after();
      if (prof && et == 0 && oi < 16) prof[32 + oi] = clock64();
      float v[kNB];
      if (act) {
        // per group of K chunks a pair [hi*hi | cross]: the hi*hi blocks in order, then the cross blocks
        const int n_grp = (o.kchunks + o.group - 1) / o.group;
        const uint32_t base = tmem + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(r * L.d_cols + mt * n_grp * 2 * kNB);
        float p1[kNB], p2[kNB];
        ptx::tmem_ld16_nowait(base, v);
        ptx::tmem_ld16_nowait(base + (n_grp > 1 ? 2 * kNB : kNB), p1);  // second hi*hi block, or the only cross block
        ptx::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < kNB; ++n) v[n] += p1[n];
        if (n_grp > 1) {
          for (int g = 2; g < n_grp; ++g) {
            ptx::tmem_ld16_nowait(base + g * 2 * kNB, p1);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int n = 0; n < kNB; ++n) v[n] += p1[n];
          }
          for (int g = 0; g < n_grp; g += 2) {
            ptx::tmem_ld16_nowait(base + g * 2 * kNB + kNB, p1);
            if (g + 1 < n_grp) ptx::tmem_ld16_nowait(base + (g + 1) * 2 * kNB + kNB, p2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int n = 0; n < kNB; ++n) v[n] += p1[n];
            if (g + 1 < n_grp) {
#pragma unroll
              for (int n = 0; n < kNB; ++n) v[n] += p2[n];
            }
          }
        }
      } else {
#pragma unroll
        for (int n = 0; n < kNB; ++n) v[n] = 0.f;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&C.d_free[r]);
      const bool stamp = prof && et == 0 && (head == CH_QLOSS || head == CH_DXA);
      if (stamp) {
        prof[62] = oi;
        prof[63] = clock64();  // accumulators in registers
      }

      // ---- layer epilogue
      if (flags & CF_BIAS_RELU) {
#pragma unroll
        for (int n = 0; n < kNB; ++n) v[n] = fmaxf(v[n] + bv, 0.f);
      }
      if (flags & CF_APPLY_MASK) {
#pragma unroll
        for (int n = 0; n < kNB; ++n) v[n] = ((mbits >> n) & 1u) ? v[n] : 0.f;
      }
      if ((flags & (CF_SAVE_MASK | CF_MASK_GLOBAL)) && act) {
        uint32_t bits = 0;
#pragma unroll
        for (int n = 0; n < kNB; ++n) bits |= (v[n] > 0.f ? 1u : 0u) << n;
        if (flags & CF_SAVE_MASK) C.mask[o.mask_slot][f] = bits;
        if (flags & CF_MASK_GLOBAL) o.gmask[static_cast<size_t>(cta) * kCFeat + f] = bits;
      }
      if (head == CH_NONE || head == CH_ACTION || head == CH_QTARGET) {
        // the layer output itself is what the next op / the weight-gradient GEMM consumes: operand first
        if (flags & CF_OUT_SMEM) {
          if (act) chain_store_operand(csm, o.out_hi, o.out_lo, o.out_sbo, pitch, f, v);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        }
      }
      const int rb = hcount & 1;
      if (head != CH_NONE) ++hcount;
      if (head == CH_QACTOR) {
        // actor loss -mean q(s, pi(s)) (ddpg.py:104, td3.py:135-137): the seed dL/dq = -1/count is a constant,
        // so dz of this layer needs no reduction -- it goes out first, q (logged only) afterwards
        if (act) {
          const float seed = -L.inv_count * hwv[0];
          float dz[kNB];
#pragma unroll
          for (int n = 0; n < kNB; ++n) dz[n] = (v[n] > 0.f && n0 + n < L.B) ? seed : 0.f;
          chain_store_operand(csm, o.out_hi, o.out_lo, o.out_sbo, pitch, f, dz);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
      }
      if (stamp) prof[64] = clock64();  // layer epilogue done

      // ---- narrow head: dot products over the feature axis (per-warp butterfly, then the warps in order)
      if (head != CH_NONE && act) {
#pragma unroll
        for (int j = 0; j < kCMaxJ; ++j) {
          if (j < J) {
            float p[kNB];
#pragma unroll
            for (int n = 0; n < kNB; ++n) p[n] = v[n] * hwv[j];
            const float s = warp_colsum16(p, lane);
            if (!(lane & 1)) C.red[rb][e][j][(lane >> 1) & 15] = s;
          }
        }
      }
      if ((head == CH_NONE || head == CH_ACTION || head == CH_QTARGET) && act) {
        // global side outputs of the plain layer, off the critical path
        if (flags & CF_OUT_GLOBAL) chain_store_gout(o.gout, o.gout_rows, f, n0, v);
        if (flags & CF_COLSUM) {
          float s = 0.f;
#pragma unroll
          for (int n = 0; n < kNB; ++n) s += v[n];
          mypart[o.vec_slot * kCFeat + f] = s;
        }
      }
      if (head == CH_NONE) continue;
      if (stamp) prof[65] = clock64();  // head partials written
      CHAIN_EPI_BAR();
      if (stamp) prof[66] = clock64();  // behind barrier A

      if (head == CH_ACTION) {
        if (et < J * kNB) {
          float a = tanhf(chain_red8(C.red[rb], hj, hn, nw) + fin_b);
          if (o.aux) a += fin_aux;
          if (o.clamp > 0.f) a = fminf(fmaxf(a, -o.clamp), o.clamp);
          if (o.x_bar >= 0) {
            float h, l;
            ptx::split_tf32(a, h, l);
            const int off = (hn >> 3) * o.x_sbo + (hj >> 2) * pitch + (hn & 7) * 16 + (hj & 3) * 4;
            *reinterpret_cast<float*>(csm + o.x_hi + off) = h;
            *reinterpret_cast<float*>(csm + o.x_lo + off) = l;
          }
          if (o.hout) o.hout[static_cast<size_t>(n0 + hn) * J + hj] = a;
          if (o.hout2) o.hout2[ct_index(L.Bp, n0 + hn, hj)] = a;
        }
        if (o.x_bar >= 0) {  // every warp arrives behind its own writes (warps without finisher threads: at once)
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.x_bar]);
        }
        continue;
      }
      if (head == CH_QTARGET) {
        if (et < kNB) C.qn_s[o.crit][et] = chain_red8(C.red[rb], 0, et, nw) + fin_b;
        continue;  // read by the same threads in the CH_QLOSS op that follows
      }
      if (head == CH_QLOSS) {
        // TD target and MSE seed per batch row (ddpg.py:94-98, td3.py:105-112): 16 column threads, then everybody
        float c_loss = 0.f, c_q = 0.f, c_y = 0.f, c_e = 0.f, c_dq = 0.f;
        if (et < kNB) {
          const float qv = chain_red8(C.red[rb], 0, et, nw) + fin_b;
          float qn = C.qn_s[0][et];
          if (L.nq == 2) qn = fminf(qn, C.qn_s[1][et]);
          const float y = C.rr[et] + ((1.0f - C.dd[et]) * L.gamma) * qn;
          const float diff = qv - y;
          const float dqv = col_valid ? L.inv_count * (2.0f * diff) : 0.f;
          C.dq_s[rb][et] = dqv;
          if (col_valid) {
            c_loss = diff * diff;
            c_dq = dqv;
            c_q = qv; c_y = y; c_e = diff;
          }
        }
        CHAIN_EPI_BAR();
        float gw = 0.f, gb = 0.f;
        if (act) {
          float dz[kNB];
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            const float dqv = C.dq_s[rb][n];
            dz[n] = v[n] > 0.f ? dqv * hwv[0] : 0.f;
            gw = fmaf(dqv, v[n], gw);
            gb += dz[n];
          }
          chain_store_operand(csm, o.out_hi, o.out_lo, o.out_sbo, pitch, f, dz);
#pragma unroll
          for (int n = 0; n < kNB; ++n) v[n] = dz[n];
        }
        if (stamp) prof[67] = clock64();  // dz in the operand buffer
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        if (stamp) prof[68] = clock64();  // arrived
        // off the critical path: gradients for the GEMM launch / the totals
        if (act) {
          mypart[o.vec_slot * kCFeat + f] = gw;   // head weight gradient
          mypart[o.vec2_slot * kCFeat + f] = gb;  // bias gradient of this layer
          chain_store_gout(o.gout, o.gout_rows, f, n0, v);
        }
        if (e == 0) {  // lanes 0-15 of the first epilogue warp hold the columns
          c_loss = half_warp_sum(c_loss);
          c_q = half_warp_sum(c_q);
          c_y = half_warp_sum(c_y);
          c_e = half_warp_sum(c_e);
          c_dq = half_warp_sum(c_dq);
          if (lane == 0) {
            C.tail_s[0] += c_loss;
            if (o.crit == 0) {
              C.tail_s[1] += c_q;
              C.tail_s[2] += c_y;
              C.tail_s[3] += c_e;
            }
            C.tail_s[4 + o.crit] = c_dq;
          }
        }
        continue;
      }
      if (head == CH_QACTOR) {
        if (e == 0) {
          float c_q = (lane < kNB && n0 + lane < L.B) ? chain_red8(C.red[rb], 0, lane, nw) + fin_b : 0.f;
          c_q = half_warp_sum(c_q);
          if (lane == 0) C.tail_s[0] += c_q;
        }
        continue;
      }
      if (head == CH_DXA) {
        // dL/da = dz0 . W0[:, S:S+A] ; through tanh: dz_head = dL/da * (1 - a^2) ; its bias gradient
        float dza = 0.f;
        if (et < J * kNB) {
          dza = col_valid ? chain_red8(C.red[rb], hj, hn, nw) * (1.f - fin_aux * fin_aux) : 0.f;
          C.dza_s[hj][hn] = dza;
        }
        if (stamp) prof[69] = clock64();  // finisher done
        CHAIN_EPI_BAR();
        if (stamp) prof[70] = clock64();  // behind barrier B
        // policy head backward (K = A): dz1[f][n] = relu'(h1) * sum_j W2[j][f] dza[j][n]
        float gb = 0.f;
        const bool act2 = f < o.Ha;
        if (act2) {
          float dz[kNB];
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < kCMaxJ; ++j)
              if (j < J) sacc = fmaf(w2v[j], C.dza_s[j][n], sacc);
            dz[n] = ((m2 >> n) & 1u) ? sacc : 0.f;
            gb += dz[n];
          }
          chain_store_operand(csm, o.out_hi, o.out_lo, o.out_sbo, pitch, f, dz);
#pragma unroll
          for (int n = 0; n < kNB; ++n) v[n] = dz[n];
        }
        if (stamp) prof[67] = clock64();  // dz in the operand buffer
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        if (stamp) prof[68] = clock64();  // arrived
        // off the critical path
        if (act2) {
          mypart[o.vec_slot * kCFeat + f] = gb;
          chain_store_gout(o.gout, o.gout_rows, f, n0, v);
        }
        if (et < J * kNB) o.hout[ct_index(128, hj, n0 + hn)] = dza;
        {
          const float ssum = half_warp_sum(dza);
          if ((lane & 15) == 0 && et < J * kNB) C.tail_s[8 + hj] = ssum;
        }
        continue;
      }
    }
    if (prof && et == 0) prof[2] = clock64();

    // ---- per-CTA partials -> totals: the last CTA to arrive adds them in CTA order
    CHAIN_EPI_BAR();
    if (et < 32) mypart[L.n_vec * kCFeat + et] = C.tail_s[et];
    CHAIN_EPI_BAR();
    if (et == 0) C.last_flag = (ptx::atom_add_acq_rel_gpu(L.counter, 1u) == static_cast<unsigned int>(L.n_cta - 1)) ? 1u : 0u;
    CHAIN_EPI_BAR();
    if (C.last_flag) {
      for (int vs = 0; vs < L.n_vec; ++vs) {
        if (et >= L.vec_n[vs]) continue;
        float tot = 0.f;
        for (int c0 = 0; c0 < L.n_cta; c0 += 16) {
          float t16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            t16[k] = (c0 + k < L.n_cta) ? __ldcg(L.part + static_cast<size_t>(c0 + k) * L.part_stride + vs * kCFeat + et) : 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) tot += t16[k];
        }
        L.vec_dst[vs][et] = tot;
      }
      if (et < 32) {
        float tot = 0.f;
        for (int c0 = 0; c0 < L.n_cta; c0 += 16) {
          float t16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            t16[k] = (c0 + k < L.n_cta) ? __ldcg(L.part + static_cast<size_t>(c0 + k) * L.part_stride + L.n_vec * kCFeat + et) : 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) tot += t16[k];
        }
        if (L.kind == 0) {
          if (et == 0) L.st->scalars[SC_CRITIC_LOSS] = tot * L.inv_count;
          else if (et == 1) L.st->scalars[SC_Q_MEAN] = tot * L.inv_count;
          else if (et == 2) L.st->scalars[SC_QT_MEAN] = tot * L.inv_count;
          else if (et == 3) L.st->scalars[SC_Q_ERR_MEAN] = tot * L.inv_count;
          else if (et == 4) L.gb3[0][0] = tot;
          else if (et == 5 && L.nq == 2) L.gb3[1][0] = tot;
        } else {
          if (et == 0) L.st->scalars[SC_ACTOR_LOSS] = -L.inv_count * tot;
          else if (et >= 8 && et < 8 + L.J) L.gb_head[et - 8] = tot;
        }
      }
      if (et == 0) *L.counter = 0u;
    }
  }

  // ---- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
  if (prof && tid == 0) prof[3] = clock64();
}

}  // namespace oprl
