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
//   D^T[feat x 16] (+)= W[feat x k] * H^T[k x 16]        tcgen05.mma kind::tf32, M = 128, N = 16
//
// Arithmetic is the same 3xTF32 scheme with cut accumulation chains as gemm.cuh (cross terms in
// their own accumulator, hi*hi terms in one accumulator per two K chunks, fp32 adds in the
// epilogue).  Narrow layers (<= 8 outputs: tanh policy heads, scalar Q heads, the action columns
// of dX) never touch the tensor core: they are dot products over the feature axis done in the
// epilogue of the layer that produces their input (fp32 FMA, warp butterfly + fixed-order
// cross-warp sum), and K <= 8 layers (the policy head backward) are a few FMAs per thread.
//
// Warp roles (17 warps, 96 registers each: a sub-partition of the SM hosts five of them): 0 = MMA
// issuer (+ TMEM owner), 1-8 = weight feeders (two per TMEM lane quarter, alternating 16 KB chunks,
// one chunk prefetched in registers, split and stored eight K columns at a time to stay inside the
// register budget), 9-16 = epilogue (lane quarter x M tile).  The weight stream is the bound (measured 303 cycles per [128 x 32] chunk,
// tools/experiments/chain_probe3.cu) and runs ahead of the MMAs through a 5-slot TMEM ring, across
// layer boundaries, so a layer costs its chunk count x ~0.16 us and nothing else.
//
// Reference semantics: ddpg.py:86-107, td3.py:95-141 (see engine.cu build_chain_ddpg_td3).
#pragma once
#include "kernels.cuh"

namespace oprl {

constexpr int kNB = 16;                  // batch rows (MMA N) per CTA
constexpr int kChainWarps = 17;
constexpr int kChainThreads = kChainWarps * 32;
constexpr int kCorePitch = 144;          // bytes between K-adjacent 8x16B core matrices of an operand buffer
                                         // (128 + 16: the feature-per-lane scalar stores hit 32 distinct banks)
constexpr int kCSlots = 5;               // TMEM ring of split A chunks (64 columns each)
constexpr int kCAcc = 5;                 // accumulators per M tile: cross terms + up to 4 hi*hi groups
constexpr int kCACol0 = 192;             // first TMEM column of the A ring (accumulators: 2 x 5 x 16 = 160)
constexpr int kCMaxOps = 16;
constexpr int kCMaxChunks = 192;
constexpr int kCMaxBufs = 12;
constexpr int kCMaxVec = 8;
constexpr int kCMaxJ = 8;
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

struct ChainLaunch {
  const ChainOp* ops;
  int n_ops;
  int B, Bp, n_cta;
  int n_in;
  ChainInput in[3];
  const unsigned int* gm_src[2];  // mask bits written by an earlier chain launch
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
  long long* prof;  // selftest / profiling: clock64 stamps of CTA 0 (nullable)
};

struct ChainCtl {
  uint64_t a_full[kCSlots], a_empty[kCSlots];
  uint64_t d_full, d_free;
  uint64_t buf_bar[kCMaxBufs];
  uint32_t tmem_slot;
  uint32_t last_flag;
  const float* chunk_src[kCMaxChunks];
  ChainOp ops[kCMaxOps];
  uint32_t mask[kCMaskSlots][kCFeat];
  float red[2][8][kCMaxJ][kNB];
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

#define CHAIN_EPI_BAR() asm volatile("bar.sync 1, 256;\n" ::: "memory")

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
      ptx::mbar_init(&C.a_full[lane], 4);   // the four quarter warps of one feeder half
      ptx::mbar_init(&C.a_empty[lane], 1);  // tcgen05.commit
    } else if (lane == kCSlots) {
      ptx::mbar_init(&C.d_full, 1);
      ptx::mbar_init(&C.d_free, 8);
    } else if (lane >= 8 && lane < 8 + kCMaxBufs) {
      ptx::mbar_init(&C.buf_bar[lane - 8], 8);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
    __syncwarp();
    ptx::tmem_alloc(&C.tmem_slot, 512);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // chunk table: flat chunk index -> source of its 16 KB (op, M tile, K chunk)
  if (tid < L.n_ops) {
    int g0 = 0;
    for (int i = 0; i < tid; ++i) g0 += C.ops[i].mtiles * C.ops[i].kchunks;
    const ChainOp& o = C.ops[tid];
    for (int mt = 0; mt < o.mtiles; ++mt)
      for (int c = 0; c < o.kchunks; ++c)
        C.chunk_src[g0 + mt * o.kchunks + c] = o.w + (static_cast<size_t>(c) * (o.w_rows >> 3) + mt * 16) * 256;
  }
  if (tid < 32) C.tail_s[tid] = 0.f;
  __syncthreads();
  int total_chunks = 0;
  for (int i = 0; i < L.n_ops; ++i) total_chunks += C.ops[i].mtiles * C.ops[i].kchunks;
  const uint32_t tmem = C.tmem_slot;

  if (warp == 0) {
    // ================================================================= MMA issuer
    const uint32_t idesc = ptx::idesc_tf32(128, kNB, 0, 0);
    int g = 0;
    for (int oi = 0; oi < L.n_ops; ++oi) {
      const ChainOp& o = C.ops[oi];
      ptx::mbar_wait(&C.buf_bar[o.in_bar], static_cast<uint32_t>(o.in_phase & 1));
      if (oi > 0) ptx::mbar_wait(&C.d_free, static_cast<uint32_t>((oi - 1) & 1));
      ptx::tc_fence_after();
      if (prof && lane == 0 && oi < 16) prof[16 + oi] = clock64();
      const uint32_t b_hi0 = ptx::smem_u32(csm + o.in_hi);
      const uint32_t b_lo0 = ptx::smem_u32(csm + o.in_lo);
      const uint32_t sbo = static_cast<uint32_t>(o.in_sbo);
      for (int mt = 0; mt < o.mtiles; ++mt) {
        const uint32_t d_cross = tmem + static_cast<uint32_t>(mt * kCAcc * kNB);
        uint32_t big = d_cross + kNB;
        int in_group = 0;
        for (int c = 0; c < o.kchunks; ++c, ++g) {
          const int slot = g % kCSlots;
          ptx::mbar_wait(&C.a_full[slot], static_cast<uint32_t>((g / kCSlots) & 1));
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t ta_hi = tmem + kCACol0 + static_cast<uint32_t>(slot * 64);
            const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t koff = static_cast<uint32_t>((c * 8 + 2 * j) * kCorePitch);
              const uint64_t db_hi = ptx::smem_desc(b_hi0 + koff, kCorePitch, sbo);
              const uint64_t db_lo = ptx::smem_desc(b_lo0 + koff, kCorePitch, sbo);
              ptx::mma_tf32_ts(d_cross, ta_lo + 8u * j, db_hi, idesc, (c | j) ? 1u : 0u);
              ptx::mma_tf32_ts(d_cross, ta_hi + 8u * j, db_lo, idesc, 1u);
              ptx::mma_tf32_ts(big, ta_hi + 8u * j, db_hi, idesc, (in_group | j) ? 1u : 0u);
            }
            ptx::mma_commit(&C.a_empty[slot]);
          }
          __syncwarp();
          if (++in_group == 2) {
            in_group = 0;
            big += kNB;
          }
        }
      }
      if (ptx::elect_one()) ptx::mma_commit(&C.d_full);
      __syncwarp();
    }
  } else if (warp <= 8) {
    // ================================================================= weight feeders
    const int q = warp & 3, half = (warp - 1) >> 2;
    const int row = q * 32 + lane;
    const uint32_t ta_lane = tmem + (static_cast<uint32_t>(q * 32) << 16) + kCACol0;
    const int roff = (row >> 3) * 64 + (row & 7);  // float4 index of this row inside a chunk (+ 8 per k core)
    ptx::pdl_wait();  // the weights are the previous launch's (Adam) output
    float4 nx[8];
    if (half < total_chunks) {
      const float4* s = reinterpret_cast<const float4*>(C.chunk_src[half]) + roff;
#pragma unroll
      for (int j = 0; j < 8; ++j) nx[j] = __ldg(s + j * 8);
    }
    for (int g = half; g < total_chunks; g += 2) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = nx[j];
      if (g + 2 < total_chunks) {
        const float4* s = reinterpret_cast<const float4*>(C.chunk_src[g + 2]) + roff;
#pragma unroll
        for (int j = 0; j < 8; ++j) nx[j] = __ldg(s + j * 8);
      }
      const int slot = g % kCSlots;
      const int use = g / kCSlots;
      if (use > 0) {
        ptx::mbar_wait(&C.a_empty[slot], static_cast<uint32_t>((use - 1) & 1));
        ptx::tc_fence_after();
      }
      const uint32_t ta = ta_lane + static_cast<uint32_t>(slot * 64);
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {  // eight K columns at a time: hi -> columns [8 j2, +8), lo -> 32 + the same
        float hi[8], lo[8];
        ptx::split_tf32(x[2 * j2].x, hi[0], lo[0]);
        ptx::split_tf32(x[2 * j2].y, hi[1], lo[1]);
        ptx::split_tf32(x[2 * j2].z, hi[2], lo[2]);
        ptx::split_tf32(x[2 * j2].w, hi[3], lo[3]);
        ptx::split_tf32(x[2 * j2 + 1].x, hi[4], lo[4]);
        ptx::split_tf32(x[2 * j2 + 1].y, hi[5], lo[5]);
        ptx::split_tf32(x[2 * j2 + 1].z, hi[6], lo[6]);
        ptx::split_tf32(x[2 * j2 + 1].w, hi[7], lo[7]);
        ptx::tmem_st8(ta + 8u * j2, hi);
        ptx::tmem_st8(ta + 32u + 8u * j2, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&C.a_full[slot]);
    }
  } else {
    // ================================================================= epilogue warps
    const int e = warp - 9;           // 0..7
    const int q = warp & 3, mt = e >> 2;
    const int f = mt * 128 + q * 32 + lane;  // feature (row of the layer output) this thread owns
    const int et = e * 32 + lane;            // 0..255
    ptx::pdl_wait();
    if (L.bump && cta == 0 && et == 0) bump_counters(L.st, L.bump_actor);
    // ---- operand buffers that come from global memory (the gathered batch), reward / done, masks
    for (int ii = 0; ii < L.n_in; ++ii) {
      const ChainInput& in = L.in[ii];
      const int nf4 = in.kchunks * 128;  // float4s: 16 rows x 32 k per chunk
      for (int i = et; i < nf4; i += 256) {
        const int c = i >> 7, w = i & 127;
        const int grp = w >> 6, j = (w >> 3) & 7, r = w & 7;
        const float4 x = __ldg(reinterpret_cast<const float4*>(
                                   in.src + (static_cast<size_t>(c) * (in.rows >> 3) + (n0 >> 3) + grp) * 256) + j * 8 + r);
        float4 h, l;
        ptx::split_tf32(x.x, h.x, l.x);
        ptx::split_tf32(x.y, h.y, l.y);
        ptx::split_tf32(x.z, h.z, l.z);
        ptx::split_tf32(x.w, h.w, l.w);
        const int off = grp * in.sbo + (c * 8 + j) * kCorePitch + r * 16;
        *reinterpret_cast<float4*>(csm + in.hi + off) = h;
        *reinterpret_cast<float4*>(csm + in.lo + off) = l;
      }
    }
    if (et < kNB) {
      const int m = min(n0 + et, L.B - 1);
      C.rr[et] = L.r ? __ldg(L.r + m) : 0.f;
      C.dd[et] = L.d ? __ldg(L.d + m) : 0.f;
    }
    for (int k = 0; k < L.n_gm; ++k) C.mask[L.gm_slot[k]][et] = __ldg(L.gm_src[k] + static_cast<size_t>(cta) * kCFeat + et);
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (int ii = 0; ii < L.n_in; ++ii) ptx::mbar_arrive(&C.buf_bar[L.in[ii].bar]);
    CHAIN_EPI_BAR();  // rr / dd / masks visible to every epilogue thread
    if (prof && et == 0) prof[1] = clock64();

    float* mypart = L.part + static_cast<size_t>(cta) * L.part_stride;
    int hcount = 0;
    for (int oi = 0; oi < L.n_ops; ++oi) {
      const ChainOp& o = C.ops[oi];
      const bool act = mt < o.mtiles;  // this thread owns a row of the output
      const int flags = o.flags;
      float bv = 0.f;
      if ((flags & CF_BIAS_RELU) && act) bv = __ldg(o.bias + f);
      uint32_t mbits = 0;
      if ((flags & CF_APPLY_MASK) && act) mbits = C.mask[o.mask_slot][f];
      float hwv[kCMaxJ];
#pragma unroll
      for (int j = 0; j < kCMaxJ; ++j) hwv[j] = 0.f;
      if (act) {
        if (o.head == CH_ACTION) {
#pragma unroll
          for (int j = 0; j < kCMaxJ; ++j)
            if (j < o.J) hwv[j] = __ldg(o.hw + static_cast<size_t>(j) * o.hw_ld + f);
        } else if (o.head == CH_DXA) {
#pragma unroll
          for (int j = 0; j < kCMaxJ; ++j)
            if (j < o.J) hwv[j] = __ldg(o.hw + static_cast<size_t>(f) * o.hw_ld + j);
        } else if (o.head != CH_NONE) {
          hwv[0] = __ldg(o.hw + f);
        }
      }
      // ---- accumulators -> registers (hi*hi groups in order, then the cross terms), release TMEM
      ptx::mbar_wait(&C.d_full, static_cast<uint32_t>(oi & 1));
      ptx::tc_fence_after();
      if (prof && et == 0 && oi < 16) prof[32 + oi] = clock64();
      float v[kNB];
#pragma unroll
      for (int n = 0; n < kNB; ++n) v[n] = 0.f;
      if (act) {
        const uint32_t base = tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * kCAcc * kNB);
        const int n_big = (o.kchunks + 1) >> 1;
        float p1[kNB], p2[kNB];
        ptx::tmem_ld16_nowait(base + kNB, v);
        if (n_big > 1) ptx::tmem_ld16_nowait(base + 2 * kNB, p1);
        ptx::tmem_ld_wait();
        if (n_big > 1) {
#pragma unroll
          for (int n = 0; n < kNB; ++n) v[n] += p1[n];
        }
        if (n_big > 2) {
          ptx::tmem_ld16_nowait(base + 3 * kNB, p1);
          if (n_big > 3) ptx::tmem_ld16_nowait(base + 4 * kNB, p2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < kNB; ++n) v[n] += p1[n];
          if (n_big > 3) {
#pragma unroll
            for (int n = 0; n < kNB; ++n) v[n] += p2[n];
          }
        }
        ptx::tmem_ld16_nowait(base, p1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < kNB; ++n) v[n] += p1[n];
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&C.d_free);

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

      const int head = o.head;
      if (head == CH_NONE || head == CH_ACTION || head == CH_QTARGET) {
        // the layer output itself is what the next op / the weight-gradient GEMM consumes
        if (act) {
          if (flags & CF_OUT_GLOBAL) {
            float* dst = o.gout + ct_index(o.gout_rows, f, n0);
#pragma unroll
            for (int n4 = 0; n4 < kNB / 4; ++n4)
              *reinterpret_cast<float4*>(dst + n4 * 32) = make_float4(v[4 * n4], v[4 * n4 + 1], v[4 * n4 + 2], v[4 * n4 + 3]);
          }
          if (flags & CF_COLSUM) {
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < kNB; ++n) s += v[n];
            mypart[o.vec_slot * kCFeat + f] = s;
          }
          if (flags & CF_OUT_SMEM) {
            uint8_t* bh = csm + o.out_hi + (f >> 2) * kCorePitch + (f & 3) * 4;
            uint8_t* bl = csm + o.out_lo + (f >> 2) * kCorePitch + (f & 3) * 4;
#pragma unroll
            for (int n = 0; n < kNB; ++n) {
              float h, l;
              ptx::split_tf32(v[n], h, l);
              const int off = (n >> 3) * o.out_sbo + (n & 7) * 16;
              *reinterpret_cast<float*>(bh + off) = h;
              *reinterpret_cast<float*>(bl + off) = l;
            }
          }
        }
        if (flags & CF_OUT_SMEM) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        }
      }
      if (head == CH_NONE) continue;

      // ---- narrow head: dot products over the feature axis
      const int rb = hcount & 1;
      ++hcount;
      const int J = (head == CH_ACTION || head == CH_DXA) ? o.J : 1;
      const int nwarps_act = 4 * o.mtiles;
      if (act) {
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
      CHAIN_EPI_BAR();
      float hsum = 0.f;  // finisher thread (j, n) = (et >> 4, et & 15): total over the warps, in warp order
      const int hj = et >> 4, hn = et & 15;
      if (et < J * kNB) {
        for (int w8 = 0; w8 < nwarps_act; ++w8) hsum += C.red[rb][w8][hj][hn];
      }
      const bool col_valid = n0 + hn < L.B;

      if (head == CH_ACTION) {
        if (et < J * kNB) {
          float a = tanhf(hsum + __ldg(o.hb + hj));
          if (o.aux) a += __ldg(o.aux + static_cast<size_t>(min(n0 + hn, L.B - 1)) * J + hj);
          if (o.clamp > 0.f) a = fminf(fmaxf(a, -o.clamp), o.clamp);
          if (o.hout) o.hout[static_cast<size_t>(n0 + hn) * J + hj] = a;
          if (o.hout2) o.hout2[ct_index(L.Bp, n0 + hn, hj)] = a;
          if (o.x_bar >= 0) {
            float h, l;
            ptx::split_tf32(a, h, l);
            const int off = (hn >> 3) * o.x_sbo + (hj >> 2) * kCorePitch + (hn & 7) * 16 + (hj & 3) * 4;
            *reinterpret_cast<float*>(csm + o.x_hi + off) = h;
            *reinterpret_cast<float*>(csm + o.x_lo + off) = l;
          }
        }
        if (o.x_bar >= 0) {
          ptx::fence_proxy_async_smem();
          CHAIN_EPI_BAR();
          if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.x_bar]);
        }
        continue;
      }
      if (head == CH_QTARGET) {
        if (et < kNB) C.qn_s[o.crit][et] = hsum + __ldg(o.hb);
        continue;  // read back by the same threads in CH_QLOSS
      }
      if (head == CH_QLOSS) {
        // TD target and MSE seed per batch row (ddpg.py:94-98, td3.py:105-112)
        float c_loss = 0.f, c_q = 0.f, c_y = 0.f, c_e = 0.f, c_dq = 0.f;
        if (et < kNB) {
          const float qv = hsum + __ldg(o.hb);
          float qn = C.qn_s[0][et];
          if (L.nq == 2) qn = fminf(qn, C.qn_s[1][et]);
          const float y = C.rr[et] + ((1.0f - C.dd[et]) * L.gamma) * qn;
          const float diff = qv - y;
          const float dq = col_valid ? L.inv_count * (2.0f * diff) : 0.f;
          C.dq_s[rb][et] = dq;
          if (col_valid) {
            c_loss = diff * diff;
            c_dq = dq;
            if (o.crit == 0) { c_q = qv; c_y = y; c_e = diff; }
          }
        }
        if (e == 0) {  // lanes 0-15 of the first epilogue warp hold the columns
          c_loss = half_warp_sum(c_loss);
          c_q = half_warp_sum(c_q);
          c_y = half_warp_sum(c_y);
          c_e = half_warp_sum(c_e);
          c_dq = half_warp_sum(c_dq);
          if (lane == 0) {
            C.tail_s[0] += c_loss;
            C.tail_s[1] += c_q;
            C.tail_s[2] += c_y;
            C.tail_s[3] += c_e;
            C.tail_s[4 + o.crit] = c_dq;
          }
        }
        CHAIN_EPI_BAR();
        if (act) {
          float gw = 0.f, gb = 0.f;
          float dz[kNB];
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            const float dq = C.dq_s[rb][n];
            dz[n] = v[n] > 0.f ? dq * hwv[0] : 0.f;
            gw = fmaf(dq, v[n], gw);
            gb += dz[n];
          }
          mypart[o.vec_slot * kCFeat + f] = gw;   // head weight gradient
          mypart[o.vec2_slot * kCFeat + f] = gb;  // bias gradient of this layer
          float* dst = o.gout + ct_index(o.gout_rows, f, n0);
#pragma unroll
          for (int n4 = 0; n4 < kNB / 4; ++n4)
            *reinterpret_cast<float4*>(dst + n4 * 32) = make_float4(dz[4 * n4], dz[4 * n4 + 1], dz[4 * n4 + 2], dz[4 * n4 + 3]);
          uint8_t* bh = csm + o.out_hi + (f >> 2) * kCorePitch + (f & 3) * 4;
          uint8_t* bl = csm + o.out_lo + (f >> 2) * kCorePitch + (f & 3) * 4;
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            float h, l;
            ptx::split_tf32(dz[n], h, l);
            const int off = (n >> 3) * o.out_sbo + (n & 7) * 16;
            *reinterpret_cast<float*>(bh + off) = h;
            *reinterpret_cast<float*>(bl + off) = l;
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        continue;
      }
      if (head == CH_QACTOR) {
        // actor loss -mean q(s, pi(s)) (ddpg.py:104, td3.py:135-137): the seed dL/dq = -1/count is a constant
        if (e == 0) {
          float c_q = (et < kNB && col_valid) ? hsum + __ldg(o.hb) : 0.f;
          c_q = half_warp_sum(c_q);
          if (lane == 0) C.tail_s[0] += c_q;
        }
        if (act) {
          const float seed = -L.inv_count * hwv[0];
          uint8_t* bh = csm + o.out_hi + (f >> 2) * kCorePitch + (f & 3) * 4;
          uint8_t* bl = csm + o.out_lo + (f >> 2) * kCorePitch + (f & 3) * 4;
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            const float dz = (v[n] > 0.f && n0 + n < L.B) ? seed : 0.f;
            float h, l;
            ptx::split_tf32(dz, h, l);
            const int off = (n >> 3) * o.out_sbo + (n & 7) * 16;
            *reinterpret_cast<float*>(bh + off) = h;
            *reinterpret_cast<float*>(bl + off) = l;
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
        continue;
      }
      if (head == CH_DXA) {
        // dL/da = dz0 . W0[:, S:S+A] ; through tanh: dz_head = dL/da * (1 - a^2) ; its bias gradient
        if (et < J * kNB) {
          const float tv = __ldg(o.aux + static_cast<size_t>(n0 + hn) * J + hj);
          const float dza = col_valid ? hsum * (1.f - tv * tv) : 0.f;
          C.dza_s[hj][hn] = dza;
          o.hout[ct_index(128, hj, n0 + hn)] = dza;
          hsum = dza;
        } else {
          hsum = 0.f;
        }
        {
          const float s = half_warp_sum(hsum);
          if ((lane & 15) == 0 && et < J * kNB) C.tail_s[8 + hj] = s;
        }
        CHAIN_EPI_BAR();
        // policy head backward (K = A): dz1[f][n] = relu'(h1) * sum_j W2[j][f] dza[j][n]
        const bool act2 = f < o.Ha;
        if (act2) {
          float w2v[kCMaxJ];
#pragma unroll
          for (int j = 0; j < kCMaxJ; ++j) w2v[j] = j < J ? __ldg(o.w2 + static_cast<size_t>(j) * o.Ha + f) : 0.f;
          const uint32_t m2 = C.mask[o.mask2_slot][f];
          float dz[kNB];
          float gb = 0.f;
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < kCMaxJ; ++j)
              if (j < J) s = fmaf(w2v[j], C.dza_s[j][n], s);
            dz[n] = ((m2 >> n) & 1u) ? s : 0.f;
            gb += dz[n];
          }
          mypart[o.vec_slot * kCFeat + f] = gb;
          float* dst = o.gout + ct_index(o.gout_rows, f, n0);
#pragma unroll
          for (int n4 = 0; n4 < kNB / 4; ++n4)
            *reinterpret_cast<float4*>(dst + n4 * 32) = make_float4(dz[4 * n4], dz[4 * n4 + 1], dz[4 * n4 + 2], dz[4 * n4 + 3]);
          uint8_t* bh = csm + o.out_hi + (f >> 2) * kCorePitch + (f & 3) * 4;
          uint8_t* bl = csm + o.out_lo + (f >> 2) * kCorePitch + (f & 3) * 4;
#pragma unroll
          for (int n = 0; n < kNB; ++n) {
            float h, l;
            ptx::split_tf32(dz[n], h, l);
            const int off = (n >> 3) * o.out_sbo + (n & 7) * 16;
            *reinterpret_cast<float*>(bh + off) = h;
            *reinterpret_cast<float*>(bl + off) = l;
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&C.buf_bar[o.out_bar]);
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
