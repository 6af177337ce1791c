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
__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(HeadFwdArgs g) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  const int m = blockIdx.x * kHeadThreads + threadIdx.x;
  if (m >= g.B) return;
  const float* o = g.out + static_cast<size_t>(m) * 2 * g.A;
  float lp = 0.f;
  for (int j = 0; j < g.A; ++j) {
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
    lp += __fsub_rn(nlp, log_det);
    store_tiled(g.X, m, j, a);
    if (g.a_rm) g.a_rm[static_cast<size_t>(m) * g.A + j] = a;
  }
  g.logp[m] = lp;
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
__global__ void __launch_bounds__(kHeadThreads) head_bwd_kernel(HeadBwdArgs g, const DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  extern __shared__ float sh[];  // [kHeadThreads x 2A]
  const int m = blockIdx.x * kHeadThreads + threadIdx.x;
  const int W = 2 * g.A;
  const float c = st->alpha * g.inv_count;
  float* mine = sh + threadIdx.x * W;
  for (int j = 0; j < W; ++j) mine[j] = 0.f;
  if (m < g.B) {
    const float* o = g.out + static_cast<size_t>(m) * W;
    for (int j = 0; j < g.A; ++j) {
      const float l = o[g.A + j];
      const float ls = fminf(fmaxf(l, kLogStdMin), kLogStdMax);
      const float sigma = expf(ls);
      const float e = g.eps[static_cast<size_t>(m) * g.A + j];
      const float a = g.a_rm[static_cast<size_t>(m) * g.A + j];
      float da = 0.f;
      for (int k = 0; k < g.n_da; ++k) da += g.da[k][static_cast<size_t>(m) * g.A + j];
      const float du = da * (1.f - a * a);  // through tanh
      const float dmu = c * (2.f * a) + du;
      const float dsigma = c * (2.f * a * e - 1.f / sigma) + du * e;
      const float dl = (l >= kLogStdMin && l <= kLogStdMax) ? dsigma * sigma : 0.f;
      store_tiled(g.dz, m, j, dmu);
      store_tiled(g.dz, m, g.A + j, dl);
      store_tiled(g.dzT, j, m, dmu);
      store_tiled(g.dzT, g.A + j, m, dl);
      mine[j] = dmu;
      mine[g.A + j] = dl;
    }
  }
  __syncthreads();
  // deterministic column sums: per block, then the last block over blocks
  for (int j = threadIdx.x; j < W; j += kHeadThreads) {
    float s = 0.f;
    for (int t = 0; t < kHeadThreads; ++t) s += sh[t * W + j];
    g.partial[blockIdx.x * W + j] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) ticket = atomicAdd(g.counter, 1u);
  __syncthreads();
  if (ticket == gridDim.x - 1) {
    __threadfence();
    for (int j = threadIdx.x; j < W; j += kHeadThreads) {
      float s = 0.f;
      for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(g.partial + b * W + j);
      g.db[j] = s;
    }
    if (threadIdx.x == 0) *g.counter = 0u;
  }
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
    } else {
      float s = 0.f;
      for (int k = 0; k < a.nq; ++k) s += q[k];
      qs += s / static_cast<float>(a.nq);
    }
    ls += a.logp[m];
    lts += a.logp[m] + a.target_entropy;
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
};
__device__ __forceinline__ void alpha_step(DevState* st, const AlphaStep& as) {
  const double x = static_cast<double>(as.add + *as.alpha_x);
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
  unsigned int* counter;
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
    float grad = 0.f;
    for (int k = 0; k < a.keep; ++k) {
      const float delta = srt[k] - zq;
      const float ad = fabsf(delta);
      const float hub = ad > 1.f ? ad - 0.5f : delta * delta * 0.5f;
      const float w = fabsf(tau - (delta < 0.f ? 1.f : 0.f));
      loss += w * hub;
      // d/dz: delta = t - z  =>  -(w * huber'(delta))
      grad -= w * (ad > 1.f ? (delta > 0.f ? 1.f : -1.f) : delta);
    }
    grad *= a.inv_total;
    store_tiled(a.dZ[net], m, j, grad);
    store_tiled(a.dZT[net], j, m, grad);
    a.dz_rm[static_cast<size_t>(m) * NT + tid] = grad;
  }
  const float bl = block_sum<kTqcThreads>(loss, red);
  if (tid == 0) a.loss_part[m] = bl;
  // last block: bias gradients (column sums over rows) + loss, in a fixed order
  __threadfence();
  __syncthreads();
  __shared__ unsigned int ticket;
  if (tid == 0) ticket = atomicAdd(a.counter, 1u);
  __syncthreads();
  if (ticket == gridDim.x - 1) {
    __threadfence();
    if (tid < NT) {
      float s = 0.f;
      for (int r = 0; r < a.B; ++r) s += __ldcg(a.dz_rm + static_cast<size_t>(r) * NT + tid);
      a.db[tid / a.nq][tid % a.nq] = s;
    }
    float l = 0.f;
    for (int r = tid; r < a.B; r += kTqcThreads) l += __ldcg(a.loss_part + r);
    const float tot = block_sum<kTqcThreads>(l, red);
    if (tid == 0) {
      st->scalars[SC_CRITIC_LOSS] = tot * a.inv_total;
      st->tick += 1;
      st->step[1] += 1;
      st->step[0] += a.bump_actor;
      st->ext_noise = 0;
      *a.counter = 0u;
    }
  }
}

}  // namespace oprl
