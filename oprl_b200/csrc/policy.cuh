// SIMT kernels of the stochastic-policy algorithms (SAC / TQC): tanh-Gaussian head
// forward/backward, actor-loss seeds, TQC atom sort + quantile-Huber loss, temperature step.
// All are tiny next to the GEMM chain; they exist to keep the whole update on the device.
#pragma once
#include <cfloat>
#include "kernels.cuh"

namespace oprl {

constexpr float kLogStdMin = -20.f, kLogStdMax = 2.f;  // nn_models.py:11
constexpr int kHeadThreads = 128;
constexpr int kMaxNets = 8;

__device__ __forceinline__ float logsigmoidf(float x) {
  // torch: min(x, 0) - log1p(exp(-|x|))
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

// ------------------------------------------------------------- tanh-Gaussian head, forward
// Reference: GaussianActor.forward (training mode) + TanhNormal (nn_models.py:169-182,197-214).
//   mu, l = split(out); ls = clamp(l, -20, 2); sigma = exp(ls); u = mu + sigma * eps; a = tanh(u)
//   logp = sum_j [ -(u-mu)^2 / (2 sigma^2) - log(sigma) - log(sqrt(2 pi))
//                  - (2 log 2 + logsigmoid(2u) + logsigmoid(-2u)) ]
struct HeadFwdArgs {
  const float* out;  // [Bp x 2A] row-major actor output
  const float* eps;  // [B x A] standard normal draws
  int B, A;
  TM X;          // action columns [0, A) of this tiled matrix receive a
  float* a_rm;   // [Bp x A] (nullable)
  float* logp;   // [Bp]
};
// One thread per (row, action): kHeadThreads / A rows per block (a thread per ROW left a batch of 1024 on 8 SMs, each
// thread walking its A actions one after the other); the row's log-probability is then added up in action order by
// the thread of action 0.
__host__ __device__ __forceinline__ int head_fwd_rows(int A) { return kHeadThreads / A; }
__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(HeadFwdArgs g) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float term[kHeadThreads];
  const int R = head_fwd_rows(g.A);
  const int rl = threadIdx.x / g.A, j = threadIdx.x - rl * g.A;
  const int m = blockIdx.x * R + rl;
  const bool live = rl < R && m < g.B;
  if (live) {
    const float* o = g.out + static_cast<size_t>(m) * 2 * g.A;
    const float mu = o[j];
    const float ls = fminf(fmaxf(o[g.A + j], kLogStdMin), kLogStdMax);
    const float sigma = expf(ls);
    const float u = __fadd_rn(mu, __fmul_rn(sigma, g.eps[static_cast<size_t>(m) * g.A + j]));
    const float a = tanhf(u);
    const float log_det = __fadd_rn(__fadd_rn(1.3862943611198906f, logsigmoidf(2.f * u)), logsigmoidf(-2.f * u));
    const float var = __fmul_rn(sigma, sigma);
    const float diff = __fsub_rn(u, mu);
    const float nlp = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(diff, diff), __fmul_rn(2.f, var)), logf(sigma)),
                                0.9189385332046727f);
    term[threadIdx.x] = __fsub_rn(nlp, log_det);
    store_tiled(g.X, m, j, a);
    if (g.a_rm) g.a_rm[static_cast<size_t>(m) * g.A + j] = a;
  }
  __syncthreads();
  if (live && j == 0) {
    float lp = 0.f;
    for (int k = 0; k < g.A; ++k) lp = __fadd_rn(lp, term[threadIdx.x + k]);
    g.logp[m] = lp;
  }
}

// ------------------------------------------------------------ tanh-Gaussian head, backward
// d out for  L = c * sum_rows logp  +  (critic path through a), with da = sum of the critics'
// input-gradient rows.  Analytic form of what autograd produces:
//   dL/dmu_j = c * 2 a_j + da_j (1 - a_j^2)
//   dL/dsigma_j = c * (-1/sigma_j + 2 a_j eps_j) + da_j (1 - a_j^2) eps_j
//   dL/dl_j = dL/dsigma_j * sigma_j   if -20 <= l_j <= 2 else 0      (clamp backward)
// Writes dz [Bp x pad32(2A)] tiled (+ transpose) and the last-layer bias gradient.
struct HeadBwdArgs {
  const float* out;
  const float* eps;
  const float* a_rm;
  const float* da[kMaxNets];  // [Bp x A] row-major each
  int n_da;
  int B, A;
  float inv_count;  // c = alpha * inv_count
  TM dz, dzT;
  float* db;        // [2A] gradient slot of the last actor bias
  float* partial;   // [gridDim.x x 2A]
  unsigned int* counter;
};
// One thread per (row, action), kHeadBwdThreads / A rows per block; deterministic bias sums: rows of a block in row
// order, then the last block to arrive adds the blocks (G = kHeadBwdThreads / 2A lanes of blocks in parallel, each in
// block order, then the G partial sums in order).
constexpr int kHeadBwdThreads = 256;
__host__ __device__ __forceinline__ int head_bwd_rows(int A) { return kHeadBwdThreads / A; }
__global__ void __launch_bounds__(kHeadBwdThreads) head_bwd_kernel(HeadBwdArgs g, const DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float sh[2 * kHeadBwdThreads];  // [R x 2A] (R * A <= 256), then [G x 2A]
  const int W = 2 * g.A;
  const int R = head_bwd_rows(g.A);
  const int rl = threadIdx.x / g.A, j = threadIdx.x - rl * g.A;
  const int m = blockIdx.x * R + rl;
  const float c = st->alpha * g.inv_count;
  if (rl < R) {
    float dmu = 0.f, dl = 0.f;
    if (m < g.B) {
      const float* o = g.out + static_cast<size_t>(m) * W;
      const float l = o[g.A + j];
      const float ls = fminf(fmaxf(l, kLogStdMin), kLogStdMax);
      const float sigma = expf(ls);
      const float e = g.eps[static_cast<size_t>(m) * g.A + j];
      const float a = g.a_rm[static_cast<size_t>(m) * g.A + j];
      float da = 0.f;
      for (int k = 0; k < g.n_da; ++k) da += g.da[k][static_cast<size_t>(m) * g.A + j];
      const float du = da * (1.f - a * a);  // through tanh
      dmu = c * (2.f * a) + du;
      const float dsigma = c * (2.f * a * e - 1.f / sigma) + du * e;
      dl = (l >= kLogStdMin && l <= kLogStdMax) ? dsigma * sigma : 0.f;
      store_tiled(g.dz, m, j, dmu);
      store_tiled(g.dz, m, g.A + j, dl);
      store_tiled(g.dzT, j, m, dmu);
      store_tiled(g.dzT, g.A + j, m, dl);
    }
    sh[rl * W + j] = dmu;
    sh[rl * W + g.A + j] = dl;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < W) {
    float s = 0.f;
    for (int t = 0; t < R; ++t) s += sh[t * W + threadIdx.x];
    g.partial[blockIdx.x * W + threadIdx.x] = s;
  }
  __syncthreads();
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) ticket = ptx::atom_add_acq_rel_gpu(g.counter, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  const int G = kHeadBwdThreads / W;
  const int grp = threadIdx.x / W, col = threadIdx.x - grp * W;
  if (grp < G) {
    float s = 0.f;
    for (unsigned int b0 = grp; b0 < gridDim.x; b0 += 16u * G) {
      float tv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const unsigned int b = b0 + static_cast<unsigned int>(k * G);
        tv[k] = b < gridDim.x ? __ldcg(g.partial + b * W + col) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) s += tv[k];
    }
    sh[grp * W + col] = s;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < W) {
    float s = 0.f;
    for (int gi = 0; gi < G; ++gi) s += sh[gi * W + threadIdx.x];
    g.db[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *g.counter = 0u;
}

// ---------------------------------------------------------------------- actor-loss seeds
// SAC (sac.py:124-126):  L = alpha * mean(logp) - mean(min(q1, q2)); seeds dL/dq_i per row.
// TQC (tqc.py:163-167):  L = mean(alpha * logp - mean_{i,j} z_ij); seeds are constant
// (pre-filled), this kernel only produces the scalars.
struct ActorSeedArgs {
  const float* q;     // [Bp x nq] row-major critic outputs at (s, pi(s))
  const float* logp;  // [Bp]
  int B, nq;          // nq = 2 (SAC) or n_nets * n_quantiles (TQC)
  int tqc;
  float inv_count;
  float target_entropy;
  float* alpha_x;  // one float behind the actor gradient arena: this rank's share of the mean the
                   // temperature loss multiplies (all-reduced with the arena under data parallelism)
  TM D[2];  // SAC seeds [Bp x 32], column 0
};
__global__ void __launch_bounds__(kTdThreads) actor_seed_kernel(ActorSeedArgs a, DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float sh[kTdThreads];
  float qs = 0.f, ls = 0.f, lts = 0.f;
  for (int m = threadIdx.x; m < a.B; m += kTdThreads) {
    const float* q = a.q + static_cast<size_t>(m) * a.nq;
    if (!a.tqc) {
      const float q1 = q[0], q2 = q[1];
      qs += fminf(q1, q2);
      // torch.min backward: the smaller input gets the gradient, ties split evenly
      const float g1 = (q1 < q2) ? 1.f : (q1 == q2 ? 0.5f : 0.f);
      store_tiled(a.D[0], m, 0, -a.inv_count * g1);
      store_tiled(a.D[1], m, 0, -a.inv_count * (1.f - g1));
    }
    ls += a.logp[m];
    lts += a.logp[m] + a.target_entropy;
  }
  if (a.tqc) {
    // mean over the n_nets * n_quantiles atoms of a row: one warp per row, coalesced, 32 loads per lane in flight (a thread
    // per row walked 125 strided loads one after the other: 13 us for a logging scalar on the critical path)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_nq = 1.f / static_cast<float>(a.nq);
    constexpr int kRows = 8, kPer = 4;  // rows in flight per warp, atoms per lane and row (nq <= 128)
    for (int m0 = warp * kRows; m0 < a.B; m0 += (kTdThreads / 32) * kRows) {
      float x[kRows][kPer];
#pragma unroll
      for (int u = 0; u < kRows; ++u)
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int k = lane + 32 * i;
          x[u][i] = (m0 + u < a.B && k < a.nq) ? a.q[static_cast<size_t>(m0 + u) * a.nq + k] : 0.f;
        }
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        float sum = (x[u][0] + x[u][1]) + (x[u][2] + x[u][3]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0 && m0 + u < a.B) qs += sum * inv_nq;
      }
    }
  }
  const float qsum = block_sum<kTdThreads>(qs, sh);
  const float lsum = block_sum<kTdThreads>(ls, sh);
  const float ltsum = block_sum<kTdThreads>(lts, sh);
  if (threadIdx.x == 0) {
    const float mean_lp = lsum * a.inv_count;
    st->scalars[SC_LOGPI_MEAN] = mean_lp;
    st->scalars[SC_ACTOR_LOSS] = st->alpha * mean_lp - qsum * a.inv_count;
    // what the temperature loss multiplies: SAC target_entropy + mean(logp) (sac.py:133-135),
    // TQC mean(logp + target_entropy) (tqc.py:163); the SAC constant is added in alpha_step
    *a.alpha_x = a.tqc ? ltsum * a.inv_count : mean_lp;
  }
}

// -------------------------------------------------------- temperature (log_alpha) Adam step
// float64 scalar Adam exactly as the reference's optim_alpha (sac.py:66-69,132-141; tqc.py:105-111,
// 175-177):  loss = -log_alpha * x  =>  grad = -x.
struct AlphaStep {
  int enabled;
  double lr;
  const float* alpha_x;
  float add;  // SAC: target_entropy ; TQC: 0
  int world;  // data-parallel learners: sum the per-rank shares (runs after the actor Adam's handshake)
  const float* peer_x[kMaxRanks];
  PubSlot* pub;  // non-null: publish_state() to this host ring (this is the update's last kernel)
};
__device__ __forceinline__ void alpha_step(DevState* st, const AlphaStep& as) {
  float share = 0.f;
  if (as.world > 1) {
    for (int r = 0; r < as.world; ++r) share += *as.peer_x[r];
  } else {
    share = *as.alpha_x;
  }
  const double x = static_cast<double>(as.add + share);
  const double g = -x;
  st->scalars[SC_ALPHA_LOSS] = static_cast<float>(-st->log_alpha * x);
  const int t = st->step[2] + 1;
  st->step[2] = t;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  st->m_alpha = st->m_alpha + (g - st->m_alpha) * (1.0 - b1);
  st->v_alpha = st->v_alpha * b2 + (1.0 - b2) * g * g;
  const double bc1 = 1.0 - pow(b1, static_cast<double>(t));
  const double bc2 = 1.0 - pow(b2, static_cast<double>(t));
  const double denom = sqrt(st->v_alpha) / sqrt(bc2) + eps;
  st->log_alpha += -(as.lr / bc1) * st->m_alpha / denom;
  st->alpha = static_cast<float>(exp(st->log_alpha));
}
__global__ void alpha_step_kernel(DevState* st, AlphaStep as) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0 && as.enabled) alpha_step(st, as);
  if (as.pub && blockIdx.x == 0) {
    __syncthreads();
    publish_state(st, as.pub, threadIdx.x, blockDim.x);
  }
}

// --------------------------------------------------------------- TQC critic loss + seeds
// Reference tqc.py:129-148 + quantile_huber_loss_f (tqc.py:14-36).  One block per batch row:
// sort the n_nets * nq target atoms, drop the top ones, build the targets, accumulate the
// quantile-Huber loss and its gradient w.r.t. every online atom.
struct TqcArgs {
  const float* zn;  // [Bp x NT] target-critic atoms (row-major, NT = n_nets * nq)
  const float* z;   // [Bp x NT] online atoms
  const float* r;
  const float* d;
  const float* logp2;
  float gamma;
  float inv_total;  // 1 / (global rows * n_nets * nq * keep)
  int B, n_nets, nq, keep;
  int bump_actor;
  TM dZ[kMaxNets], dZT[kMaxNets];
  float* db[kMaxNets];  // last-layer bias gradient slots [nq]
  float* dz_rm;         // [Bp x NT] scratch (row-major gradient, for the bias sums)
  float* loss_part;     // [B]
  float* part2;         // [ceil(B / 16)][NT + 1] group partials of the two-level reduction
  unsigned int* counter;  // [1 + ceil(B / 16)]
};
constexpr int kTqcThreads = 128;  // NT <= 128
__global__ void __launch_bounds__(kTqcThreads) tqc_loss_kernel(TqcArgs a, DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float srt[kTqcThreads];
  __shared__ float red[kTqcThreads];
  const int m = blockIdx.x;
  const int NT = a.n_nets * a.nq;
  const int tid = threadIdx.x;
  srt[tid] = tid < NT ? a.zn[static_cast<size_t>(m) * NT + tid] : FLT_MAX;
  __syncthreads();
  // bitonic sort, ascending, 128 keys
  for (int k = 2; k <= kTqcThreads; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = tid ^ j;
      if (ixj > tid) {
        const float x = srt[tid], y = srt[ixj];
        const bool up = (tid & k) == 0;
        if ((x > y) == up) { srt[tid] = y; srt[ixj] = x; }
      }
      __syncthreads();
    }
  }
  // target_k = r + ((1 - d) * gamma) * (z_(k) - alpha * logp')       (tqc.py:143-145)
  const float coef = (1.0f - a.d[m]) * a.gamma;
  const float shift = st->alpha * a.logp2[m];
  const float rew = a.r[m];
  float tk = 0.f;
  if (tid < a.keep) tk = rew + coef * (srt[tid] - shift);
  __syncthreads();
  srt[tid] = tk;
  __syncthreads();
  float loss = 0.f;
  if (tid < NT) {
    const int net = tid / a.nq, j = tid - net * a.nq;
    const float zq = a.z[static_cast<size_t>(m) * NT + tid];
    const float tau = static_cast<float>(j) / static_cast<float>(a.nq) + 0.5f / static_cast<float>(a.nq);
    // four independent accumulator pairs (k mod 4), added in a fixed order: one warp per scheduler walking `keep`
    // dependent adds was the longest phase of this kernel
    float lacc[4] = {0.f, 0.f, 0.f, 0.f}, gacc[4] = {0.f, 0.f, 0.f, 0.f};
    auto term = [&](int k, float& l, float& g) {
      const float delta = srt[k] - zq;
      const float ad = fabsf(delta);
      const float hub = ad > 1.f ? ad - 0.5f : delta * delta * 0.5f;
      const float w = fabsf(tau - (delta < 0.f ? 1.f : 0.f));
      l += w * hub;
      // d/dz: delta = t - z  =>  -(w * huber'(delta))
      g -= w * (ad > 1.f ? (delta > 0.f ? 1.f : -1.f) : delta);
    };
    int k = 0;
    for (; k + 4 <= a.keep; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) term(k + u, lacc[u], gacc[u]);
    }
    for (int u = 0; k < a.keep; ++k, ++u) term(k, lacc[u], gacc[u]);
    loss = (lacc[0] + lacc[1]) + (lacc[2] + lacc[3]);
    float grad = (gacc[0] + gacc[1]) + (gacc[2] + gacc[3]);
    grad *= a.inv_total;
    store_tiled(a.dZ[net], m, j, grad);
    store_tiled(a.dZT[net], j, m, grad);
    a.dz_rm[static_cast<size_t>(m) * NT + tid] = grad;
  }
  const float bl = block_sum<kTqcThreads>(loss, red);
  if (tid == 0) a.loss_part[m] = bl;
  // bias gradients (column sums of dz over the rows) + loss, in a fixed order and two levels: the last block of
  // every 16 consecutive rows adds that group (one batch of 16 loads in flight), the last group to finish adds the
  // groups -- a single last block walking all B rows was 15 us of this kernel.
  __syncthreads();
  __shared__ unsigned int ticket;
  const unsigned int grp = blockIdx.x >> 4, ngrp = (gridDim.x + 15u) >> 4;
  const unsigned int gsize = min(16u, gridDim.x - grp * 16u);
  if (tid == 0) ticket = ptx::atom_add_acq_rel_gpu(a.counter + 1 + grp, 1u);
  __syncthreads();
  if (ticket != gsize - 1) return;
  const int PW = NT + 1;  // group partial: NT column sums + the loss
  {
    float tv[16];
    if (tid < NT) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        tv[k] = static_cast<unsigned int>(k) < gsize ? __ldcg(a.dz_rm + static_cast<size_t>(grp * 16 + k) * NT + tid) : 0.f;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) s += tv[k];
      a.part2[grp * PW + tid] = s;
    }
    if (tid == 0) {
      float lv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) lv[k] = static_cast<unsigned int>(k) < gsize ? __ldcg(a.loss_part + grp * 16 + k) : 0.f;
      float l = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) l += lv[k];
      a.part2[grp * PW + NT] = l;
      a.counter[1 + grp] = 0u;
    }
  }
  __syncthreads();
  if (tid == 0) ticket = ptx::atom_add_acq_rel_gpu(a.counter, 1u);
  __syncthreads();
  if (ticket != ngrp - 1) return;
  if (tid <= NT && tid < kTqcThreads) {
    // thread NT (or thread 0 when NT == 128, below) adds the losses
    float s = 0.f;
    for (unsigned int g0 = 0; g0 < ngrp; g0 += 16) {
      float tv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) tv[k] = g0 + k < ngrp ? __ldcg(a.part2 + (g0 + k) * PW + tid) : 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) s += tv[k];
    }
    if (tid < NT) a.db[tid / a.nq][tid % a.nq] = s;
    else st->scalars[SC_CRITIC_LOSS] = s * a.inv_total;
  }
  if (tid == 0) {
    if (NT == kTqcThreads) {
      float s = 0.f;
      for (unsigned int g0 = 0; g0 < ngrp; ++g0) s += __ldcg(a.part2 + g0 * PW + NT);
      st->scalars[SC_CRITIC_LOSS] = s * a.inv_total;
    }
    bump_counters(st, a.bump_actor);
    *a.counter = 0u;
  }
}

}  // namespace oprl

namespace oprl {

// ------------------------------------------------ scalar critic head, forward + backward (SIMT)
// For critics with ONE output (DDPG / TD3 / SAC, nn_models.py:27-81) the last Linear layer is a
// dot product per row: running it -- and the first backward step it seeds -- as padded tensor-core
// GEMM stages costs three launches of the dependency chain for ~0.1 % of the FLOPs.  This kernel
// does the whole neighbourhood of the loss in one launch, one block per 8 batch rows, one thread
// per (critic, hidden unit):
//   mode 0 (critic step; ddpg.py:94-98, td3.py:95-112, sac.py:96-105)
//     q_i = h2_i . w3_i + b3_i ; qn_i likewise from the target nets ; y = r + (1-d) gamma (min_i qn_i - alpha logp')
//     L = sum_i mean (q_i - y)^2 ; dq_i = (2/count)(q_i - y)
//   mode 1 (actor step; ddpg.py:104, td3.py:135-137, sac.py:124-126)
//     q_i = h2_i . w3_i + b3_i at (s, pi(s)) ; L = alpha mean(logp) - mean(min_i q_i) ; dq_i = -(1/count)[i = argmin]
//   both: dz2_i = dq_i w3_i^T (.) relu'(h2_i)  (tiled + transposed), and in mode 0 the gradients of
//   w3_i, b3_i and of the previous layer's bias (column sums of dz2_i), reduced in a fixed order.
struct CriticHeadArgs {
  int mode, nq, B, H;
  int h_rows;    // padded rows of the tiled h2 / dz2 matrices (Bp)
  int dzT_rows;  // padded rows of the transposed dz2 (pad128(H))
  float gamma, inv_count, target_entropy;
  const float* h2[2];
  const float* h2t[2];
  const float* w3[2];
  const float* b3[2];
  const float* w3t[2];
  const float* b3t[2];
  const float* r;
  const float* d;
  const float* logp2;  // mode 0, SAC (nullable)
  const float* logp;   // mode 1, SAC (nullable)
  float* dz2[2];
  float* dz2T[2];  // nullable (mode 1: no weight gradients needed)
  float* gw3[2];
  float* gb3[2];
  float* gb2[2];
  float* part;   // [gridDim.x][nq * 2 * H + 8]
  float* part2;  // [ceil(gridDim.x / 16)][same]: group partials of the two-level reduction (more than 32 blocks)
  unsigned int* counter;  // [1 + ceil(gridDim.x / 16)]
  float* alpha_x;  // mode 1, SAC: share of mean(logp) behind the actor gradient arena (nullable)
  int bump_actor;
};
constexpr int kHeadRows = 8;
// One thread per (critic, hidden unit): blockDim.x = pad32(H) * nq.  With both critics in one thread (round 1) TD3's
// head stage cost 14.3 us against 7.4 us for DDPG's single critic: twice the loads, shuffles and reduction round
// trips in series.  The critics meet only in the 32-entry result block in shared memory (min over the targets).
__global__ void __launch_bounds__(512) critic_head_kernel(const __grid_constant__ CriticHeadArgs a, DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float sh[];  // [warps][16] dot partials, then [4][8] results
  const int tid = threadIdx.x;
  const int Hp = blockDim.x / a.nq;  // threads per critic (a multiple of 32)
  const int ci = tid / Hp;           // this thread's critic
  const int n = tid - ci * Hp;       // hidden unit
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, wpc = Hp >> 5;
  const int m0 = blockIdx.x * kHeadRows;
  const bool live = n < a.H;
  // Nothing in this kernel reads the step counters / tick, and everything that consumed the old
  // values ran in earlier launches: advance them first, off the critical tail.
  if (a.mode == 0 && blockIdx.x == 0 && tid == 0) bump_counters(st, a.bump_actor);
  // per-row inputs of the loss, fetched now so that their latency overlaps the dot products
  float rr[kHeadRows], dd[kHeadRows], lp[kHeadRows];
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) {
    const int m = min(m0 + r, a.B - 1);
    rr[r] = a.mode == 0 ? a.r[m] : 0.f;
    dd[r] = a.mode == 0 ? a.d[m] : 0.f;
    lp[r] = a.mode == 0 ? (a.logp2 ? a.logp2[m] : 0.f) : (a.logp ? a.logp[m] : 0.f);
  }
  float b3v[2] = {0.f, 0.f}, b3tv[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (i >= a.nq) break;
    b3v[i] = a.b3[i][0];
    if (a.mode == 0) b3tv[i] = a.b3t[i][0];
  }
  // ---- phase 1: the dot products of this block's 8 rows, this thread's critic (online, and target in the critic step)
  float hv[kHeadRows];  // online hidden activations of this thread's unit (kept for phase 2)
  float prod[2][kHeadRows];
  const float w3v = live ? a.w3[ci][n] : 0.f;
  {
    const float w3t = (a.mode == 0 && live) ? a.w3t[ci][n] : 0.f;
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const size_t off = ct_index(a.h_rows, m0 + r, n);
      const float h = live ? a.h2[ci][off] : 0.f;
      hv[r] = h;
      prod[0][r] = h * w3v;
      prod[1][r] = (a.mode == 0 && live) ? a.h2t[ci][off] * w3t : 0.f;
    }
  }
  float* red = sh;  // [nwarps][16]: [online | target][row]
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (s == 1 && a.mode != 0) continue;
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      float x = prod[s][r];
      x += __shfl_xor_sync(0xffffffffu, x, 16);
      x += __shfl_xor_sync(0xffffffffu, x, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 4);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      if (lane == 0) red[warp * 16 + s * kHeadRows + r] = x;
    }
  }
  __syncthreads();
  // res[(2 * target + critic) * 8 + row]: streams 0-1 online critics, 2-3 target critics (critic step only)
  float* res = sh + nwarps * 16;
  if (tid < 32) {
    const int strm = tid >> 3, r = tid & 7;
    const int c = strm & 1, tg = strm >> 1;
    float x = 0.f;
    if (c < a.nq && (tg == 0 || a.mode == 0))
      for (int wv = c * wpc; wv < (c + 1) * wpc; ++wv) x += red[wv * 16 + tg * kHeadRows + r];
    res[tid] = x;
  }
  __syncthreads();
  // ---- per-row loss terms and seeds (every thread computes them redundantly: no extra barrier)
  float dq[2][kHeadRows];
  float loss = 0.f, qsum = 0.f, ysum = 0.f, esum = 0.f, dsum[2] = {0.f, 0.f}, lpsum = 0.f;
  const float alpha = st->alpha;
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) {
    const int m = m0 + r;
    dq[0][r] = dq[1][r] = 0.f;
    if (m >= a.B) continue;
    float q[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) q[i] = i < a.nq ? res[i * kHeadRows + r] + b3v[i] : 0.f;
    if (a.mode == 0) {
      float qn = res[2 * kHeadRows + r] + b3tv[0];
      if (a.nq == 2) qn = fminf(qn, res[3 * kHeadRows + r] + b3tv[1]);
      if (a.logp2) qn = qn - alpha * lp[r];
      const float y = rr[r] + ((1.0f - dd[r]) * a.gamma) * qn;
      ysum += y;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i >= a.nq) break;
        const float diff = q[i] - y;
        loss += diff * diff;
        dq[i][r] = a.inv_count * (2.0f * diff);
        dsum[i] += dq[i][r];
        if (i == 0) { qsum += q[0]; esum += diff; }
      }
    } else {
      if (a.nq == 2) {
        // torch.min backward: the smaller input gets the gradient, ties split evenly
        const float g1 = (q[0] < q[1]) ? 1.f : (q[0] == q[1] ? 0.5f : 0.f);
        dq[0][r] = -a.inv_count * g1;
        dq[1][r] = -a.inv_count * (1.f - g1);
        qsum += fminf(q[0], q[1]);
      } else {
        dq[0][r] = -a.inv_count;
        qsum += q[0];
      }
      if (a.logp) lpsum += lp[r];
    }
  }
  // ---- phase 2: dz2 = dq w3^T (.) relu'(h2), column sums -- this thread's critic
  const int W = a.nq * 2 * a.H + 8;
  float* mypart = a.part + static_cast<size_t>(blockIdx.x) * W;
  {
    float gw = 0.f, gb = 0.f;
    float dzv[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const float dqr = ci == 0 ? dq[0][r] : dq[1][r];
      const float dz = hv[r] > 0.f ? dqr * w3v : 0.f;
      dzv[r] = dz;
      gw = fmaf(dqr, hv[r], gw);
      gb += dz;
      if (live) a.dz2[ci][ct_index(a.h_rows, m0 + r, n)] = dz;
    }
    if (live && a.dz2T[ci]) {
      float* dst = a.dz2T[ci] + ct_index(a.dzT_rows, n, m0);
      *reinterpret_cast<float4*>(dst) = make_float4(dzv[0], dzv[1], dzv[2], dzv[3]);
      *reinterpret_cast<float4*>(dst + 32) = make_float4(dzv[4], dzv[5], dzv[6], dzv[7]);
    }
    if (live && a.mode == 0) {
      mypart[ci * 2 * a.H + n] = gw;
      mypart[ci * 2 * a.H + a.H + n] = gb;
    }
  }
  if (tid == 0) {
    float* tail = mypart + a.nq * 2 * a.H;
    tail[0] = loss; tail[1] = qsum; tail[2] = ysum; tail[3] = esum;
    tail[4] = dsum[0]; tail[5] = dsum[1]; tail[6] = lpsum;
  }
  // ---- last block: fixed-order reduction over blocks.  One release/acquire ticket by one thread
  // (the block barrier orders every thread's partial stores before it) instead of a __threadfence()
  // in every thread on both sides of a relaxed atomic.
  __syncthreads();
  __shared__ unsigned int ticket;
  const float* part = a.part;  // what the final reduction adds up: `nparts` rows of W floats
  unsigned int nparts = gridDim.x;
  if (gridDim.x > 32) {
    // batch > 256 rows: two levels.  The last block of every 16 consecutive blocks adds that group's partials
    // (W columns over the threads, 16 loads in flight each), the last GROUP to finish adds the groups below: one
    // block pulling all of [gridDim.x][W] through its SM (528 KB at batch 1024) was 15 us of this kernel.
    const unsigned int grp = blockIdx.x >> 4, ngrp = (gridDim.x + 15u) >> 4;
    const unsigned int gsize = min(16u, gridDim.x - grp * 16u);
    if (tid == 0) ticket = ptx::atom_add_acq_rel_gpu(a.counter + 1 + grp, 1u);
    __syncthreads();
    if (ticket != gsize - 1) return;
    const float* gp = a.part + static_cast<size_t>(grp) * 16 * W;
    float* dst = a.part2 + static_cast<size_t>(grp) * W;
    for (int c = tid; c < W; c += blockDim.x) {
      float tv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) tv[k] = static_cast<unsigned int>(k) < gsize ? __ldcg(gp + static_cast<size_t>(k) * W + c) : 0.f;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sum += tv[k];
      dst[c] = sum;
    }
    if (tid == 0) a.counter[1 + grp] = 0u;
    __syncthreads();
    if (tid == 0) ticket = ptx::atom_add_acq_rel_gpu(a.counter, 1u);
    __syncthreads();
    if (ticket != ngrp - 1) return;
    part = a.part2;
    nparts = ngrp;
  } else {
    if (tid == 0) ticket = ptx::atom_add_acq_rel_gpu(a.counter, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
  }
  // every load of the reduction is issued before the first add: the per-block scalar tails (warp 0:
  // lane = block, 7 values each) and, in the critic step, the head-weight / bias-gradient partials
  // of this thread's (critic, hidden unit) (batches of sixteen blocks)
  float tl[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const unsigned int nb32 = (nparts + 31) / 32;
  if (warp == 0 && nb32 == 1) {
    if (static_cast<unsigned int>(lane) < nparts) {
      const float* tp = part + static_cast<size_t>(lane) * W + a.nq * 2 * a.H;
#pragma unroll
      for (int k = 0; k < 7; ++k) tl[k] = __ldcg(tp + k);
    }
  }
  if (a.mode == 0 && live) {
    // loads batched sixteen blocks at a time (independent L2 round trips), adds in block order
    float gw = 0.f, gb = 0.f;
    for (unsigned int b0 = 0; b0 < nparts; b0 += 16) {
      float tw[16], tb[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const bool ok = b0 + k < nparts;
        const float* p = part + static_cast<size_t>(ok ? b0 + k : 0) * W + ci * 2 * a.H;
        tw[k] = ok ? __ldcg(p + n) : 0.f;
        tb[k] = ok ? __ldcg(p + a.H + n) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        gw += tw[k];
        gb += tb[k];
      }
    }
    a.gw3[ci][n] = gw;
    a.gb2[ci][n] = gb;
  }
  float t[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (nb32 == 1) {
    // <= 32 blocks: a fixed xor tree over the lanes of warp 0 (deterministic)
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        float x = tl[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        t[k] = x;
      }
    }
  } else {
    // more parts: gathered by (block, k) threads in parallel, summed in block order
    __syncthreads();
    for (unsigned int idx = tid; idx < nparts * 8; idx += blockDim.x)
      sh[idx] = __ldcg(part + static_cast<size_t>(idx >> 3) * W + a.nq * 2 * a.H + (idx & 7));
    __syncthreads();
    if (tid == 0)
      for (unsigned int b = 0; b < nparts; ++b)
        for (int k = 0; k < 7; ++k) t[k] += sh[b * 8 + k];
  }
  if (tid == 0) {
    if (a.mode == 0) {
      st->scalars[SC_CRITIC_LOSS] = t[0] * a.inv_count;
      st->scalars[SC_Q_MEAN] = t[1] * a.inv_count;
      st->scalars[SC_QT_MEAN] = t[2] * a.inv_count;
      st->scalars[SC_Q_ERR_MEAN] = t[3] * a.inv_count;
      a.gb3[0][0] = t[4];
      if (a.nq == 2) a.gb3[1][0] = t[5];
    } else {
      const float mean_lp = t[6] * a.inv_count;
      st->scalars[SC_LOGPI_MEAN] = mean_lp;
      st->scalars[SC_ACTOR_LOSS] = (a.logp ? st->alpha * mean_lp : 0.f) - t[1] * a.inv_count;
      if (a.alpha_x) *a.alpha_x = mean_lp;
    }
    *a.counter = 0u;
  }
}

}  // namespace oprl
