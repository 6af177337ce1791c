// Grouped small-GEMM kernel for the off-policy update: every MLP layer forward,
// dX and dW product of the DDPG/TD3/SAC/TQC update is one `GemmOp`.
//
//   D[M x N] = sum_k A(m,k) * B(n,k)       (fp32 accumulate in TMEM)
//
// * Operands live in HBM/L2 in the CT32 "core-tiled" layout, pre-split into
//   tf32 hi/lo halves by whoever produced them (previous epilogue / Adam).
//   Both operands are always K-major (tcgen05 kind::tf32 only takes MN-major
//   operands in the 32-bit-swizzled SW128_32B layout), so producers that feed a
//   dX (= delta * W) or dW (= delta^T * X) product also write a transposed tiled
//   copy (`tt_*` outputs here, W^T from the Adam kernel).
// * fp32-accurate products on tensor cores via 3xTF32:
//       A*B ~= Alo*Bhi + Ahi*Blo + Ahi*Bhi        (tcgen05.mma.kind::tf32)
//   (needed for the reference's 1e-5 parameter-L2 parity bar).
//   The tensor core's accumulator add is not round-to-nearest, so a long accumulation
//   chain loses ~1 ulp per step (measured: one 96-step chain gave 8x the parameter error of
//   the FFMA cross-check).  The chain is therefore cut: the two small cross terms go to
//   their own TMEM accumulator, the Ahi*Bhi terms to one accumulator per group of K
//   chunks, and the epilogue adds the partial sums with ordinary fp32 adds.
// * Tile 128 x 32 per CTA, K streamed in 32-wide chunks through a 4-stage
//   mbarrier ring filled by 1-D bulk async copies (TMA engine, no tensor maps:
//   the producers already wrote UMMA-canonical core matrices).
// * Warp roles: warp0 = copy producer, warp1 = MMA issuer (+TMEM owner),
//   warps2-5 = epilogue (TMEM -> regs -> bias/act/mask -> tiled hi/lo + row-major).
// * `kSimt` variant keeps loads/epilogue identical but does the products with
//   FFMA from shared memory: the on-device cross-check for the descriptor path.
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
constexpr int kStages = 4;
constexpr int kAFloats = kBM * kBK;  // 4096 floats (16 KB) per hi / lo
constexpr int kBFloats = kBN * kBK;  // 1024 floats ( 4 KB) per hi / lo
constexpr int kStageFloats = 2 * kAFloats + 2 * kBFloats;
constexpr int kStageBytes = kStageFloats * 4;  // 40 KB
constexpr int kGemmThreads = 192;
constexpr int kGemmSmemBytes = kStages * kStageBytes + 1024;
constexpr int kMaxOps = 12;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct GemmOp {
  // operands (CT32 hi/lo)
  const float* a_hi;
  const float* a_lo;
  const float* b_hi;
  const float* b_lo;
  // epilogue inputs
  const float* bias;     // v += bias[n] for n < bias_n
  const float* mask_hi;  // v *= (mask[m][n] > 0), CT32 with mask_rows padded rows
  const float* rs;       // row-major [m][n] matrix: v *= (1 - rs^2) for n < rs_n (tanh')
  const float* addm;     // row-major [m][n] matrix added after the activation, n < addm_n
  // outputs
  float* t_hi;  // CT32 output (hi/lo) at column offset t_c0, only columns n < t_n
  float* t_lo;
  float* tt_hi;  // transposed CT32 output: element (n, m) of a [tt_rows x M] matrix
  float* tt_lo;
  float* rm;      // row-major output, m < rm_m, n < rm_n
  float* colsum;  // per-M-tile partials: colsum[mtile * colsum_ld + n]
  float* colsum_out;       // if set: the last CTA of each N tile writes the total over M tiles here
  unsigned int* colsum_cnt;  // one arrival counter per N tile (self-resetting)
  int a_rows;  // padded row count of the stored A matrix [M.. x K]
  int b_rows;  // padded row count of the stored B matrix [N.. x K]
  int M, N, K;       // padded problem (M % 128, N % 32, K % 32)
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
  int passes;  // 3 = 3xTF32 (fp32-accurate), 1 = single tf32 pass
  float alpha;  // v *= alpha (applied last)
  float clamp;  // if > 0: v = min(max(v, -clamp), clamp) after addm
};

struct GemmLaunch {
  GemmOp op[kMaxOps];
  int n_ops;
  long long* prof;  // selftest only: per-phase clock64 stamps of CTA 0
};

__host__ __device__ __forceinline__ int gemm_tiles(const GemmOp& o) {
  return (o.M / kBM) * (o.N / kBN);
}

template <bool kSimt>
__global__ void __launch_bounds__(kGemmThreads, 1)
    gemm_kernel(const __grid_constant__ GemmLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* accum = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  float* cs_smem = reinterpret_cast<float*>(smem_raw + kStages * kStageBytes + 256);  // [4][32]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // ---- which op / tile is this CTA
  int t = blockIdx.x;
  int oi = 0;
  for (; oi < L.n_ops; ++oi) {
    int nt = gemm_tiles(L.op[oi]);
    if (t < nt) break;
    t -= nt;
  }
  if (oi >= L.n_ops) return;
  const GemmOp& o = L.op[oi];
  const int ntn = o.N / kBN;
  const int mt = t / ntn;
  const int m0 = mt * kBM;
  const int n0 = (t % ntn) * kBN;
  const int nchunks = o.K / kBK;
  const int passes = o.passes;
  long long* prof = (L.prof && blockIdx.x == 0) ? L.prof : nullptr;
  if (prof && tid == 0) {
    prof[0] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    prof[9] = static_cast<long long>(gt);
  }

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(accum, 1);
    ptx::fence_mbar_init();
  }
  // accumulators: column block 0 = cross terms, blocks 1.. = hi*hi per group of K chunks
  const int group = max(2, (nchunks + 14) / 15);
  const int n_big = (nchunks + group - 1) / group;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(32 * (n_big + 1))) tmem_cols <<= 1;
  if (!kSimt && warp == 1) ptx::tmem_alloc(tmem_slot, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (prof && tid == 0) prof[1] = clock64();

  float v[kBN];

  if (warp == 0) {
    // ===================== producer: bulk copies HBM/L2 -> smem ring
    const int a_rb = o.a_rows >> 3;
    const int b_rb = o.b_rows >> 3;
    const uint32_t tx = (passes == 3 ? 2u : 1u) * (kAFloats + kBFloats) * 4u;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % kStages;
      const uint32_t ph = (c / kStages) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) ptx::mbar_expect_tx(&full[s], tx);
      __syncwarp();
      float* st = smem + s * kStageFloats;
      const int halves = (passes == 3) ? 2 : 1;
      // 4 pieces per stage: A hi, B hi, A lo, B lo -- one lane each
      if (lane < 2 * halves) {
        const int h = lane >> 1;
        if ((lane & 1) == 0) {
          ptx::bulk_g2s(st + h * kAFloats,
                        (h ? o.a_lo : o.a_hi) + (static_cast<size_t>(c) * a_rb + (m0 >> 3)) * 256,
                        kAFloats * 4, &full[s]);
        } else {
          ptx::bulk_g2s(st + 2 * kAFloats + h * kBFloats,
                        (h ? o.b_lo : o.b_hi) + (static_cast<size_t>(c) * b_rb + (n0 >> 3)) * 256,
                        kBFloats * 4, &full[s]);
        }
      }
    }
    if (prof && lane == 0) prof[2] = clock64();
  } else if (!kSimt && warp == 1) {
    // ===================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t tmem_d = *tmem_slot;
      const uint32_t idesc = ptx::idesc_tf32(kBM, kBN, 0, 0);
      // smem tile = [row group of 8][8 K-cores][8 rows][16 B]: K cores 128 B apart (LBO),
      // 8-row groups 1 KB apart (SBO); one MMA (K=8) consumes two K cores = 256 B.
      const uint32_t a_step = 256u, b_step = 256u;
      const uint32_t a_lbo = 128u, a_sbo = 1024u, b_lbo = 128u, b_sbo = 1024u;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kStages;
        const uint32_t ph = (c / kStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (prof && c == 0) prof[3] = clock64();
        const uint32_t sa_hi = ptx::smem_u32(smem + s * kStageFloats);
        const uint32_t sa_lo = sa_hi + kAFloats * 4;
        const uint32_t sb_hi = sa_hi + 2 * kAFloats * 4;
        const uint32_t sb_lo = sb_hi + kBFloats * 4;
#pragma unroll
        for (int j = 0; j < kBK / 8; ++j) {
          const uint64_t da_hi = ptx::smem_desc(sa_hi + j * a_step, a_lbo, a_sbo);
          const uint64_t db_hi = ptx::smem_desc(sb_hi + j * b_step, b_lbo, b_sbo);
          const uint32_t big = tmem_d + 32u * static_cast<uint32_t>(1 + c / group);
          const uint32_t big_acc = ((c % group) | j) ? 1u : 0u;
          if (passes == 3) {
            const uint64_t da_lo = ptx::smem_desc(sa_lo + j * a_step, a_lbo, a_sbo);
            const uint64_t db_lo = ptx::smem_desc(sb_lo + j * b_step, b_lbo, b_sbo);
            ptx::mma_tf32(tmem_d, da_lo, db_hi, idesc, (c | j) ? 1u : 0u);
            ptx::mma_tf32(tmem_d, da_hi, db_lo, idesc, 1u);
          }
          ptx::mma_tf32(big, da_hi, db_hi, idesc, big_acc);
        }
        ptx::mma_commit(&empty[s]);
      }
      ptx::mma_commit(accum);
      if (prof) prof[4] = clock64();
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (!kSimt) {
      ptx::mbar_wait(accum, 0);
      ptx::tc_fence_after();
      if (prof && tid == 64) prof[5] = clock64();
      {
        const uint32_t lane_base = *tmem_slot + (static_cast<uint32_t>(q * 32) << 16);
        float part[kBN];
        ptx::tmem_ld32(lane_base + 32u, v);
        for (int gi = 1; gi < n_big; ++gi) {
          ptx::tmem_ld32(lane_base + 32u * static_cast<uint32_t>(1 + gi), part);
#pragma unroll
          for (int j = 0; j < kBN; ++j) v[j] += part[j];
        }
        if (passes == 3) {
          ptx::tmem_ld32(lane_base, part);
#pragma unroll
          for (int j = 0; j < kBN; ++j) v[j] += part[j];
        }
      }
      if (prof && tid == 64) prof[6] = clock64();
    } else {
      // FFMA cross-check path: same smem contents, products on CUDA cores.
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kStages;
        const uint32_t ph = (c / kStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        const float* sa_hi = smem + s * kStageFloats;
        const float* sa_lo = sa_hi + kAFloats;
        const float* sb_hi = sa_hi + 2 * kAFloats;
        const float* sb_lo = sb_hi + kBFloats;
        for (int k = 0; k < kBK; ++k) {
          const int ia = ((row >> 3) * 8 + (k >> 2)) * 32 + (row & 7) * 4 + (k & 3);
          float a = sa_hi[ia];
          if (passes == 3) a += sa_lo[ia];
#pragma unroll
          for (int j = 0; j < kBN; ++j) {
            const int ib = ((j >> 3) * 8 + (k >> 2)) * 32 + (j & 7) * 4 + (k & 3);
            float b = sb_hi[ib];
            if (passes == 3) b += sb_lo[ib];
            v[j] = fmaf(a, b, v[j]);
          }
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (tid == 64) ptx::mbar_arrive(&empty[s]);
      }
    }

    const int m = m0 + row;
    // ---- epilogue math
    if (o.bias) {
#pragma unroll
      for (int j = 0; j < kBN; ++j)
        if (n0 + j < o.bias_n) v[j] += __ldg(o.bias + n0 + j);
    }
    if (o.act == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (o.act == ACT_TANH) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] = tanhf(v[j]);
    }
    if (o.mask_hi) {
#pragma unroll
      for (int j4 = 0; j4 < kBN / 4; ++j4) {
        const float4 mk =
            *reinterpret_cast<const float4*>(o.mask_hi + ct_index(o.mask_rows, m, n0 + 4 * j4));
        v[4 * j4 + 0] = mk.x > 0.f ? v[4 * j4 + 0] : 0.f;
        v[4 * j4 + 1] = mk.y > 0.f ? v[4 * j4 + 1] : 0.f;
        v[4 * j4 + 2] = mk.z > 0.f ? v[4 * j4 + 2] : 0.f;
        v[4 * j4 + 3] = mk.w > 0.f ? v[4 * j4 + 3] : 0.f;
      }
    }
    if (o.rs) {
#pragma unroll
      for (int j = 0; j < kBN; ++j)
        if (n0 + j < o.rs_n) {
          const float a = o.rs[static_cast<size_t>(m) * o.rs_ld + n0 + j];
          v[j] *= (1.f - a * a);
        }
    }
    if (o.addm) {
#pragma unroll
      for (int j = 0; j < kBN; ++j)
        if (n0 + j < o.addm_n) v[j] += o.addm[static_cast<size_t>(m) * o.addm_ld + n0 + j];
    }
    if (o.clamp > 0.f) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] = fminf(fmaxf(v[j], -o.clamp), o.clamp);
    }
    if (o.alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] *= o.alpha;
    }
    if (o.m_valid > 0 && m >= o.m_valid) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) v[j] = 0.f;
    }
    if (o.n_valid > 0) {
#pragma unroll
      for (int j = 0; j < kBN; ++j)
        if (n0 + j >= o.n_valid) v[j] = 0.f;
    }
    // ---- outputs
    if (o.t_hi) {
#pragma unroll
      for (int j4 = 0; j4 < kBN / 4; ++j4) {
        if (n0 + 4 * j4 < o.t_n) {
          float4 hi, lo;
          ptx::split_tf32(v[4 * j4 + 0], hi.x, lo.x);
          ptx::split_tf32(v[4 * j4 + 1], hi.y, lo.y);
          ptx::split_tf32(v[4 * j4 + 2], hi.z, lo.z);
          ptx::split_tf32(v[4 * j4 + 3], hi.w, lo.w);
          const size_t off = ct_index(o.t_rows, m, o.t_c0 + n0 + 4 * j4);
          *reinterpret_cast<float4*>(o.t_hi + off) = hi;
          *reinterpret_cast<float4*>(o.t_lo + off) = lo;
        }
      }
    }
    if (o.tt_hi) {
      // element (n, m) of the transposed matrix; lanes cover 32 consecutive m.
#pragma unroll
      for (int j = 0; j < kBN; ++j) {
        float hi, lo;
        ptx::split_tf32(v[j], hi, lo);
        const size_t off = ct_index(o.tt_rows, n0 + j, m);
        o.tt_hi[off] = hi;
        o.tt_lo[off] = lo;
      }
    }
    if (o.rm) {
      if (!o.rm_trans) {
        if (m < o.rm_m) {
#pragma unroll
          for (int j = 0; j < kBN; ++j) {
            const int n = n0 + j;
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
          for (int j = 0; j < kBN; ++j)
            if (n0 + j < o.rm_n) o.rm[static_cast<size_t>(n0 + j) * o.rm_ld + m] = v[j];
        }
      }
    }
    if (o.colsum) {
#pragma unroll
      for (int j = 0; j < kBN; ++j) {
        float s = v[j];
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (lane == j) cs_smem[q * 32 + j] = s;
      }
      asm volatile("bar.sync 2, 128;\n" ::: "memory");
      if (warp == 2) {
        const float s = cs_smem[lane] + cs_smem[32 + lane] + cs_smem[64 + lane] + cs_smem[96 + lane];
        o.colsum[static_cast<size_t>(mt) * o.colsum_ld + n0 + lane] = s;
        if (o.colsum_out) {
          // deterministic cross-CTA total: the last M tile to arrive sums all partials in order
          const int mtiles = o.M / kBM;
          __threadfence();
          __syncwarp();
          unsigned int ticket = 0;
          if (lane == 0) ticket = atomicAdd(o.colsum_cnt + (n0 / kBN), 1u);
          ticket = __shfl_sync(0xffffffffu, ticket, 0);
          if (ticket == static_cast<unsigned int>(mtiles - 1)) {
            __threadfence();
            float tot = 0.f;
            for (int i = 0; i < mtiles; ++i)
              tot += __ldcg(o.colsum + static_cast<size_t>(i) * o.colsum_ld + n0 + lane);
            if (n0 + lane < o.colsum_n) o.colsum_out[n0 + lane] = tot;
            if (lane == 0) o.colsum_cnt[n0 / kBN] = 0u;
          }
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
