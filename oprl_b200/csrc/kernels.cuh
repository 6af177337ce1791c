// SIMT kernels around the GEMM chain: replay-buffer gather (K0 in SURVEY.md §2b),
// TD target + MSE seed gradient (K3/K4), fused Adam + Polyak + tf32 re-tiling
// (K6/K9/K10), counters + noise draws.  All HBM/latency bound.
#pragma once
#include <curand_kernel.h>
#include "gemm.cuh"

namespace oprl {

struct TM {  // CT32 tiled fp32 matrix (gemm.cuh)
  float* p;
  int rows;  // padded
  int cols;  // padded
};

// Engine scalars that live on the device so a captured CUDA graph can replay.
struct DevState {
  unsigned long long tick;  // number of updates completed (noise / sampling stream offset)
  int step[4];              // Adam step counts: 0 actor, 1 critic, 2 alpha
  int ext_noise;            // bit w: the caller injected the w-th normal draw of the next update
  int pad0;
  double log_alpha, m_alpha, v_alpha;  // SAC/TQC temperature, float64 like the reference
  float alpha;                         // exp(log_alpha) rounded to fp32
  float pad1;
  float scalars[32];  // see enum Scalar
  // torch.optim.Adam bias corrections, kept current by the kernel that advances the step
  // counters so that the Adam launches do no double-precision pow() on their critical path:
  //   step_size = lr / (1 - beta1^t),  bc2_sqrt = sqrt(1 - beta2^t)      (0 actor, 1 critic)
  double lr[2];
  double b1pow[2], b2pow[2];  // beta1^t, beta2^t as running products
  float step_size[2], bc2_sqrt[2];
};

constexpr double kBeta1 = 0.9, kBeta2 = 0.999;

// Host-visible copy of the update's scalars: the last kernel of an update writes them into slot (tick - 1) % kHostRing
// of a pinned, device-mapped ring which the host polls -- no D2H copy, no event, in loops that read the losses of
// every update (OPRL_B200_HOST_SCALARS=1).  The slot is a sequence-locked record: twelve 16-byte units, each three
// payload words + the low word of the tick, each written by ONE aligned 16-byte store, so no system-scope fence has
// to order "payload, then sequence number" (that fence held the update's last kernel ~1.5 us: PCIe round trip); the
// host accepts the record when every unit carries the tick it waits for and still does after the payload was read.
// Payload word p: scalars[p] for p < 32, alpha for p = 32, zero above.
// Call with all threads of one block, after a barrier that orders the block's own writes to *st.
constexpr int kHostRing = 8;
constexpr int kPubUnits = 12;
struct alignas(256) PubSlot {
  unsigned int w[64];  // kPubUnits x {payload, payload, payload, tick}
};
__device__ __forceinline__ void publish_state(const DevState* st, PubSlot* ring, int tid, int nthreads) {
  (void)nthreads;
  if (!ring || tid >= kPubUnits) return;
  const unsigned long long tick = *reinterpret_cast<const volatile unsigned long long*>(&st->tick);
  PubSlot* slot = ring + ((tick - 1ull) % kHostRing);
  unsigned int v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int p = 3 * tid + k;
    float f = 0.f;
    if (p < 32) f = *reinterpret_cast<const volatile float*>(&st->scalars[p]);
    else if (p == 32) f = *reinterpret_cast<const volatile float*>(&st->alpha);
    v[k] = __float_as_uint(f);
  }
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(slot->w + 4 * tid), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(static_cast<unsigned int>(tick))
               : "memory");
}

// Advance the per-update counters (one thread, at the end of the loss kernel): everything that
// consumed the old tick (sampling, noise) ran in earlier launches; the Adam launches that need
// the new step counts run later.
__device__ __forceinline__ void bump_counters(DevState* st, int bump_actor) {
  st->tick += 1;
  st->ext_noise = 0;
#pragma unroll
  for (int opt = 0; opt < 2; ++opt) {
    if (opt == 0 && !bump_actor) continue;
    st->step[opt] += 1;
    st->b1pow[opt] *= kBeta1;
    st->b2pow[opt] *= kBeta2;
    st->step_size[opt] = static_cast<float>(st->lr[opt] / (1.0 - st->b1pow[opt]));
    st->bc2_sqrt[opt] = static_cast<float>(sqrt(1.0 - st->b2pow[opt]));
  }
}

enum Scalar : int {
  SC_CRITIC_LOSS = 0,
  SC_ACTOR_LOSS = 1,
  SC_ALPHA_LOSS = 2,
  SC_Q_MEAN = 3,
  SC_QT_MEAN = 4,
  SC_LOGPI_MEAN = 5,
  SC_Q_ERR_MEAN = 6,
  SC_ALPHA = 7,
};

__device__ __forceinline__ void store_tiled(const TM& t, int r, int c, float x) {
  t.p[ct_index(t.rows, r, c)] = x;
}

// Standard-normal draws of one update.  raw[i] is either injected by the caller
// (parity tests: the reference's torch.randn stream) or drawn here with Philox;
// out[i] = clip(raw[i] * scale, +-clip)  (TD3 target-policy smoothing, td3.py:98-100).
struct NoiseSpec {
  float* raw;   // [n]
  float* out;   // [n] (may alias raw when scale == 1 and clip <= 0), nullable
  int n;
  float scale;
  float clip;  // <= 0: no clip
};

// -------------------------------------------------------------------- gather
// Reference: EpisodicReplayBuffer.sample, episodic_buffer.py:123-133 -- five
// advanced-index gathers; next_state = states[ep, step + 1] is the adjacent row.
struct GatherArgs {
  // sources: replay storage (gathered) or dense user batch (dense = 1)
  const float* states;   // [E, L+1, S]   | dense: [B, S]
  const float* actions;  // [E, L, A]     | dense: [B, A]
  const float* rewards;  // [E, L, 1]     | dense: [B]
  const float* dones;    // [E, L, 1]     | dense: [B]
  const float* next_states;  // dense only: [B, S]
  const int* ep_step;        // [B][2] (episode, step) pairs, or nullptr -> device sampling
  const int* prefix;         // [n_eps + 1] transition prefix sums (device sampling)
  int n_eps, n_trans;
  int L, S, A, A4, B;
  int dense;
  unsigned long long seed;
  // outputs
  float *bs, *ba, *br, *bd, *bs2;  // row-major batch arena the caller sees (every pointer nullable)
  float *wr, *wd;                   // rewards / dones of this batch as the update's kernels read them
  // RNG stream offset = number of updates completed, and the injected-noise mask: passed by the
  // host (which counts the updates it has launched) so that a gather running ahead of the previous
  // update -- on another stream, into the other working set -- does not race with the device counters
  unsigned long long tick;
  int ext;
  int dev_tick;  // 1: read DevState::tick instead (prefetching gather inside the update graph, placed
                 // behind the loss kernel that advances the counter -- nobody writes it after that)
  // n-step returns (north_star; the reference stores `gamma` in the buffer and never uses it, episodic_buffer.py:18):
  // with n_step > 1 a sampled (episode, step) yields  R = sum_{k<m} gamma^k r_{t+k},  s' = s_{t+m},
  // m = min(n_step, steps up to and including the first done, steps left in the episode), and the EFFECTIVE done
  // d' = 1 - (1 - d_{t+m-1}) gamma^{m-1}, so that the unchanged 1-step target r + (1 - d') gamma Q'(s') is the n-step
  // target R + (1 - d) gamma^m Q'(s_{t+m}).  n_step <= 1: the reference's 1-step transition, bit for bit.
  int n_step;
  float nstep_gamma;
  int* out_ep_step;                 // [B][2] what was sampled (device sampling), nullable
  TM X, XT, Xn, Xp;
  NoiseSpec noise[2];  // blocks >= B draw the update's normals (n == 0: none)
};

constexpr int kGatherThreads = 64;   // noise blocks / row-major gather
constexpr int kGatherRows = 8;       // batch rows per block, one warp each
constexpr int kGatherBlock = 32 * kGatherRows;
constexpr int kPrefixSmem = 4096;    // episode prefix sums cached in shared memory up to this many
__global__ void __launch_bounds__(kGatherBlock)
    gather_kernel(const __grid_constant__ GatherArgs g, const DevState* st) {
  const int row_blocks = (g.B + kGatherRows - 1) / kGatherRows;
  if (static_cast<int>(blockIdx.x) >= row_blocks) {
    // ---- noise blocks: 4 normals per thread, one Philox subsequence per quad
    const int total = g.noise[0].n + g.noise[1].n;
    const unsigned long long tick = g.dev_tick ? st->tick : g.tick;
    const int ext = g.ext;
    const int nb = gridDim.x - row_blocks;
    for (int base = ((blockIdx.x - row_blocks) * kGatherBlock + threadIdx.x) * 4; base < total;
         base += nb * kGatherBlock * 4) {
      curandStatePhilox4_32_10_t rng;
      curand_init(g.seed, static_cast<unsigned long long>(base >> 2), tick * 8ull, &rng);
      const float4 z = curand_normal4(&rng);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = base + k;
        if (i >= total) break;
        const int w = (i < g.noise[0].n) ? 0 : 1;
        const NoiseSpec& sp = g.noise[w];
        const int j = w ? i - g.noise[0].n : i;
        float v;
        if (ext & (1 << w)) v = sp.raw[j];
        else { v = zz[k]; sp.raw[j] = v; }
        if (sp.out) {
          v *= sp.scale;
          if (sp.clip > 0.f) v = fminf(fmaxf(v, -sp.clip), sp.clip);
          sp.out[j] = v;
        }
      }
    }
    return;
  }
  // ---- gather blocks: one warp per sampled transition
  __shared__ int s_prefix[kPrefixSmem];
  const bool device_draw = !g.dense && !g.ep_step;
  const bool cached = device_draw && g.n_eps + 1 <= kPrefixSmem;
  if (cached) {
    // one coalesced pass instead of a dependent global-memory binary search per row
    for (int i = threadIdx.x; i <= g.n_eps; i += kGatherBlock) s_prefix[i] = g.prefix[i];
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kGatherRows + warp;
  if (b >= g.B) return;
  int ep = 0, step = 0;
  if (!g.dense) {
    if (g.ep_step) {
      ep = g.ep_step[2 * b];
      step = g.ep_step[2 * b + 1];
    } else {
      if (lane == 0) {
        curandStatePhilox4_32_10_t rng;
        curand_init(g.seed ^ 0x9E3779B97F4A7C15ull, static_cast<unsigned long long>(b), (g.dev_tick ? st->tick : g.tick) * 4ull, &rng);
        const unsigned int u = curand(&rng);
        const int t = static_cast<int>((static_cast<unsigned long long>(u) * g.n_trans) >> 32);
        const int* pre = cached ? s_prefix : g.prefix;
        int lo = 0, hi = g.n_eps;  // find ep with prefix[ep] <= t < prefix[ep+1]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (pre[mid] <= t) lo = mid; else hi = mid;
        }
        ep = lo;
        step = t - pre[lo];
        if (g.out_ep_step) {
          g.out_ep_step[2 * b] = ep;
          g.out_ep_step[2 * b + 1] = step;
        }
      }
      ep = __shfl_sync(0xffffffffu, ep, 0);
      step = __shfl_sync(0xffffffffu, step, 0);
    }
  }
  const float* src_a;
  const float* src_s;
  const float* src_s2;
  float r, d;
  if (g.dense) {
    src_a = g.actions + static_cast<size_t>(b) * g.A;
    src_s = g.states + static_cast<size_t>(b) * g.S;
    src_s2 = g.next_states + static_cast<size_t>(b) * g.S;
    r = g.rewards[b];
    d = g.dones[b];
  } else {
    const size_t row = static_cast<size_t>(ep) * g.L + step;
    const size_t srow = static_cast<size_t>(ep) * (g.L + 1) + step;
    src_a = g.actions + row * g.A;
    src_s = g.states + srow * g.S;
    src_s2 = src_s + g.S;
    r = g.rewards[row];
    d = g.dones[row];
    if (g.n_step > 1) {
      // (every lane computes the same few terms: n is small and the loads are broadcast)
      const int ep_len = g.prefix[ep + 1] - g.prefix[ep];
      float gpow = 1.f;  // gamma^(m-1)
      int m = 1;
      while (m < g.n_step && d == 0.f && step + m < ep_len) {
        gpow = __fmul_rn(gpow, g.nstep_gamma);
        r = __fadd_rn(r, __fmul_rn(gpow, g.rewards[row + m]));
        d = g.dones[row + m];
        ++m;
      }
      src_s2 = src_s + static_cast<size_t>(m) * g.S;
      d = __fsub_rn(1.f, __fmul_rn(__fsub_rn(1.f, d), gpow));
    }
  }
  // s and s' are ADJACENT rows of the replay storage: one contiguous 2S-float read.  When S is a multiple of 4 (and
  // so is the padded action width, which makes every state quad land inside one 4-column core of the tiled
  // matrices) the row goes as 16-byte vectors: a lane = four state columns -> one LDG.128, one 16-byte store per
  // row-major / tiled destination (walker: 12 lanes cover both states instead of 48 scalar round trips).
  if ((g.S & 3) == 0 && (g.A4 & 3) == 0 && (reinterpret_cast<uintptr_t>(src_s) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(src_s2) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.bs) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(g.bs2) & 15) == 0) {
    const int q4 = g.S >> 2;  // quads per state row
    for (int c = lane; c < 2 * q4; c += 32) {
      const bool nxt = c >= q4;
      const int j = (nxt ? c - q4 : c) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>((nxt ? src_s2 : src_s) + j));
      if (!nxt) {
        if (g.bs) *reinterpret_cast<float4*>(g.bs + static_cast<size_t>(b) * g.S + j) = v;
        *reinterpret_cast<float4*>(g.X.p + ct_index(g.X.rows, b, g.A4 + j)) = v;
        *reinterpret_cast<float4*>(g.Xp.p + ct_index(g.Xp.rows, b, g.A4 + j)) = v;
        store_tiled(g.XT, g.A4 + j + 0, b, v.x);
        store_tiled(g.XT, g.A4 + j + 1, b, v.y);
        store_tiled(g.XT, g.A4 + j + 2, b, v.z);
        store_tiled(g.XT, g.A4 + j + 3, b, v.w);
      } else {
        if (g.bs2) *reinterpret_cast<float4*>(g.bs2 + static_cast<size_t>(b) * g.S + j) = v;
        *reinterpret_cast<float4*>(g.Xn.p + ct_index(g.Xn.rows, b, g.A4 + j)) = v;
      }
    }
    for (int c = lane; c < g.A; c += 32) {
      const float v = src_a[c];
      if (g.ba) g.ba[static_cast<size_t>(b) * g.A + c] = v;
      store_tiled(g.X, b, c, v);
      store_tiled(g.XT, c, b, v);
    }
  } else {
  const int W = g.A + 2 * g.S;
  for (int c = lane; c < W; c += 32) {
    if (c < g.A) {
      const float v = src_a[c];
      if (g.ba) g.ba[static_cast<size_t>(b) * g.A + c] = v;
      store_tiled(g.X, b, c, v);
      store_tiled(g.XT, c, b, v);
    } else if (c < g.A + g.S) {
      const int j = c - g.A;
      const float v = src_s[j];
      if (g.bs) g.bs[static_cast<size_t>(b) * g.S + j] = v;
      store_tiled(g.X, b, g.A4 + j, v);
      store_tiled(g.XT, g.A4 + j, b, v);
      store_tiled(g.Xp, b, g.A4 + j, v);
    } else {
      const int j = c - g.A - g.S;
      const float v = src_s2[j];
      if (g.bs2) g.bs2[static_cast<size_t>(b) * g.S + j] = v;
      store_tiled(g.Xn, b, g.A4 + j, v);
    }
  }
  }
  if (lane == 0) {
    if (g.br) g.br[b] = r;
    if (g.bd) g.bd[b] = d;
    g.wr[b] = r;
    g.wd[b] = d;
  }
}

// Plain row-major gather (EpisodicReplayBuffer.sample without an attached engine,
// episodic_buffer.py:127-133): one block per sampled transition.
struct RowGatherArgs {
  const float *states, *actions, *rewards, *dones;
  const int* ep_step;
  int L, S, A, B;
  float *s, *a, *r, *d, *s2;
};
__global__ void __launch_bounds__(kGatherThreads) gather_rows_kernel(RowGatherArgs g) {
  const int b = blockIdx.x;
  const int ep = g.ep_step[2 * b], step = g.ep_step[2 * b + 1];
  const size_t row = static_cast<size_t>(ep) * g.L + step;
  const float* src_s = g.states + (static_cast<size_t>(ep) * (g.L + 1) + step) * g.S;
  const float* src_a = g.actions + row * g.A;
  for (int c = threadIdx.x; c < 2 * g.S + g.A; c += blockDim.x) {
    if (c < g.S) g.s[static_cast<size_t>(b) * g.S + c] = src_s[c];
    else if (c < 2 * g.S) g.s2[static_cast<size_t>(b) * g.S + c - g.S] = src_s[c];  // adjacent row
    else g.a[static_cast<size_t>(b) * g.A + c - 2 * g.S] = src_a[c - 2 * g.S];
  }
  if (threadIdx.x == 0) {
    g.r[b] = g.rewards[row];
    g.d[b] = g.dones[row];
  }
}

// ---------------------------------------------------------------- ingest scatter
// EpisodicReplayBuffer.add_transition / add_episode (episodic_buffer.py:81-112): the host stages whole
// transitions in a pinned ring, one row = [state S | action A | reward | done | episode (int bits) | step (int bits)];
// one H2D copy of the staged rows and ONE launch of this kernel place them in the replay storage (one warp per
// transition), instead of two small copies and two fill kernels per transition.
struct ScatterArgs {
  float *states, *actions, *rewards, *dones;
  const float* staged;  // [n][S + A + 4]
  int n, L, S, A;
};
constexpr int kScatterRows = 8;
__global__ void __launch_bounds__(32 * kScatterRows) scatter_transitions_kernel(ScatterArgs g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kScatterRows + warp;
  if (row >= g.n) return;
  const int W = g.S + g.A + 4;
  const float* src = g.staged + static_cast<size_t>(row) * W;
  const int ep = __float_as_int(src[g.S + g.A + 2]);
  const int step = __float_as_int(src[g.S + g.A + 3]);
  float* ds = g.states + (static_cast<size_t>(ep) * (g.L + 1) + step) * g.S;
  float* da = g.actions + (static_cast<size_t>(ep) * g.L + step) * g.A;
  for (int c = lane; c < g.S; c += 32) ds[c] = src[c];
  for (int c = lane; c < g.A; c += 32) da[c] = src[g.S + c];
  if (lane == 0) {
    g.rewards[static_cast<size_t>(ep) * g.L + step] = src[g.S + g.A];
    g.dones[static_cast<size_t>(ep) * g.L + step] = src[g.S + g.A + 1];
  }
}

// ---------------------------------------------------------------- block reduce
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* sh) {
  // fixed-shape tree -> bit-reproducible
  sh[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

// ------------------------------------------------------- TD target (DDPG / TD3)
// Reference: ddpg.py:94-98, td3.py:95-112.
//   y = r + ((1 - d) * gamma) * min_i q'_i ;  L = sum_i mean((q_i - y)^2)
//   dL/dq_i = (1/B) * (2 * (q_i - y))
// Writes the seed gradient as tiled [Bp x 32] matrices (column 0) + transposes,
// the last-layer bias gradients, and the logging scalars.  One block.
struct TdArgs {
  const float* qn;  // [Bp x nq] target-critic outputs
  const float* q;   // [Bp x nq] online critic outputs
  const float* r;
  const float* d;
  const float* logpi_next;  // SAC: [Bp] log pi(a'|s') (nullable)
  float gamma;
  float inv_count;  // 1 / (global batch rows): the losses are means over all learners' rows
  int B, nq;
  int bump_actor;   // this update also steps the actor (Adam step counter)
  TM D3[2];
  TM D3T[2];
  float* db3[2];  // gradient slot of each critic's output bias
  float* y_out;   // [Bp] (nullable)
};
constexpr int kTdThreads = 256;
__global__ void __launch_bounds__(kTdThreads) td_kernel(TdArgs a, DevState* st) {
  ptx::pdl_trigger();
  ptx::pdl_wait();
  __shared__ float sh[kTdThreads];
  const float invB = a.inv_count;
  float loss[2] = {0.f, 0.f}, dsum[2] = {0.f, 0.f}, qsum = 0.f, ysum = 0.f, esum = 0.f;
  const float alpha = st->alpha;
  for (int m = threadIdx.x; m < a.B; m += kTdThreads) {
    float qn = a.qn[static_cast<size_t>(m) * a.nq];
    if (a.nq == 2) qn = fminf(qn, a.qn[static_cast<size_t>(m) * 2 + 1]);
    if (a.logpi_next) qn = qn - alpha * a.logpi_next[m];
    const float y = a.r[m] + ((1.0f - a.d[m]) * a.gamma) * qn;
    if (a.y_out) a.y_out[m] = y;
    ysum += y;
    for (int i = 0; i < a.nq; ++i) {
      const float q = a.q[static_cast<size_t>(m) * a.nq + i];
      const float diff = q - y;
      loss[i] += diff * diff;
      const float dq = invB * (2.0f * diff);
      dsum[i] += dq;
      store_tiled(a.D3[i], m, 0, dq);
      store_tiled(a.D3T[i], 0, m, dq);
      if (i == 0) { qsum += q; esum += diff; }
    }
  }
  float total = 0.f;
  for (int i = 0; i < a.nq; ++i) {
    const float l = block_sum<kTdThreads>(loss[i], sh);
    const float ds = block_sum<kTdThreads>(dsum[i], sh);
    total += l * invB;
    if (threadIdx.x == 0) a.db3[i][0] = ds;
  }
  const float qs = block_sum<kTdThreads>(qsum, sh);
  const float ys = block_sum<kTdThreads>(ysum, sh);
  const float es = block_sum<kTdThreads>(esum, sh);
  if (threadIdx.x == 0) {
    st->scalars[SC_CRITIC_LOSS] = total;
    st->scalars[SC_Q_MEAN] = qs * invB;
    st->scalars[SC_QT_MEAN] = ys * invB;
    st->scalars[SC_Q_ERR_MEAN] = es * invB;
    bump_counters(st, a.bump_actor);
  }
}

// ------------------------------------------------ Adam + Polyak + tf32 re-tiling
// Reference: torch.optim.Adam single-tensor path (ddpg.py:51,56,101,107) and the
// Polyak loops (ddpg.py:72-84, nn_functions.py:5-10).  One pass over
// [theta, g, m, v, theta_target]; also refreshes the CT32 hi/lo operand copies
// (W and W^T, target W) that the GEMM kernel consumes (plain fp32; the GEMM splits to tf32).
// One tensor of a parameter group.  The five arenas of a group (theta, grad, m, v, target) share one layout, so a
// segment carries its element offset (`goff`) and the kernel gets the arena bases as a parameter (AdamArenas, constant
// bank): five pointers fewer in the registers of every thread than one pointer per array.
struct AdamSeg {
  float* w;   // tiled [rows_pad x cols_pad] (nullable for biases)
  float* wt;  // tiled transposed (nullable)
  float* tw;  // tiled target weights (nullable)
  // Deferred layer-0 gradient (GemmOp::dw0_defer): when the launch says so (gp_mt > 0) the gradient of element
  // (r, c) of this tensor is the sum over M tiles mt < gp_mt of gpart[mt * w_rows * wt_rows + r * wt_rows + c], c = the
  // tiled column of the element (weights) or gp_ones (the bias, whose gradient rides in a pad column of the same
  // product).  nullptr: always read the gradient arena.
  const float* gpart;
  int gp_off;  // the same partials as an offset into the (exported) gradient arena: where a data-parallel peer's copy is
  int gp_ones;
  int w_rows, wt_rows;   // padded row counts of the tiled copies W [w_rows x wt_rows], W^T [wt_rows x w_rows]
  int n, rows, cols;     // row-major [rows x cols]
  int split, off_lo, off_hi;  // tiled col = j < split ? j + off_lo : j - split + off_hi
  int opt;                    // 0 actor, 1 critic
  int goff;                   // element offset of this tensor inside the group's arenas
};
struct AdamArenas {
  float* theta;
  float* grad;
  float* m;
  float* v;
  float* target;  // nullable
};
constexpr int kDeferMaxMt = 8;  // M tiles (of 128 batch rows) a deferred layer-0 gradient may have

// ---- fused gradient all-reduce (data-parallel learners on one NVLink domain) ----------------
// Every rank maps every other rank's gradient arena and flag block (CUDA IPC).  The Adam kernel
// itself performs the reduction: after a flag handshake ("my gradients of step t are complete")
// each thread sums the same element from all ranks IN RANK ORDER -- so all replicas compute
// bit-identical sums -- and goes straight on to the Adam update.  No NCCL launch, no extra pass
// over the gradients, no parameter broadcast.
constexpr int kMaxRanks = 8;
struct CommArgs {
  int world, rank;  // world <= 1: disabled
  int group;        // 0 actor, 1 critic: which flag rows to use
  int exit_barrier; // also wait until every peer has finished reading this rank's gradients
  const float* peer_grad[kMaxRanks];    // this group's gradient arena on every rank (own included)
  unsigned int* peer_flags[kMaxRanks];  // every rank's flag block: [2 groups][2 phases][kMaxRanks]
  unsigned int* done_counter;           // local: blocks of this launch that finished reading
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// phase 0: "gradients of step `epoch` are complete here"; phase 1: "I have read everybody's".
// Both are called by the first `world` threads of a block, thread r handling rank r, so the remote
// flag stores and the polls of the local flag block run side by side instead of one after another.
__device__ __forceinline__ void comm_signal(const CommArgs& cm, int phase, unsigned int epoch, int r) {
  __threadfence_system();
  st_release_sys(cm.peer_flags[r] + (cm.group * 2 + phase) * kMaxRanks + cm.rank, epoch);
}
__device__ __forceinline__ void comm_wait(const CommArgs& cm, int phase, unsigned int epoch, int r) {
  const unsigned int* mine = cm.peer_flags[cm.rank] + (cm.group * 2 + phase) * kMaxRanks;
  unsigned int spins = 0;
  // steps only grow: ">=" tolerates a peer that is already one handshake ahead
  while (static_cast<int>(ld_acquire_sys(mine + r) - epoch) < 0) {
    if (++spins > (1u << 28)) {
      printf("oprl: rank %d timed out waiting for rank %d (group %d phase %d epoch %u)\n", cm.rank, r, cm.group,
             phase, epoch);
      __trap();
    }
  }
}
struct AdamHyper {  // python doubles of torch.optim.Adam, rounded to fp32 where torch does
  double lr[2];
  double beta1, beta2;
  float w1, w2;  // (float)(1 - beta1), (float)(1 - beta2)
  float beta2f, eps;
  float tau, one_minus_tau;  // (float)tau, (float)(1 - tau)
};
constexpr int kAdamThreads = 256;
constexpr int kAdamPatchSmem = 2 * 32 * 33 * 4;  // dynamic shared memory of a launch with 32 x 32 patches
// Optional duty of block 0 of an Adam launch: the actor loss of the DDPG / TD3 actor step,
//   out = scale * sum_{m < B} (b3 + sum_{p < nparts} part[p * ld + m])        (= -mean q(s, pi(s)))
// from the per-tile partial head dot products the critic forward GEMM left behind (GemmOp::tail_out).
// A logging scalar only -- nothing on the gradient path waits for it.
struct LossTail {
  PubSlot* pub;       // non-null: this launch is the update's last kernel -- publish_state() to this ring
  const float* part;  // nullptr = no duty
  const float* b3;
  float* out;
  int nparts, ld, B;
  float scale;
};
// mode: bit0 = Adam step, bit1 = Polyak, bit2 = (re)tile online weights, bit3 = tile targets
// kPatch: compiled with the 32 x 32 patch path (groups of large matrices, TQC).  The plain instantiation is what the
// latency-bound DDPG / TD3 / SAC launches run: the patch path's extra registers (48 -> 64) alone cost 1-2 us per
// update there (measured A/B on one box), so it is kept out of their kernel.
// kDP: compiled with the in-kernel gradient all-reduce (data-parallel learners); the single-learner instantiation
// carries none of its registers.
template <bool kPatch, bool kDP>
__global__ void __launch_bounds__(kAdamThreads)
    adam_kernel(const AdamSeg* segs, const int2* blocks, const __grid_constant__ AdamArenas ar, AdamHyper hp,
                const DevState* st, int mode, const __grid_constant__ CommArgs cm, const LossTail lt, int gp_mt) {
  ptx::pdl_trigger();
  // block -> (tensor, first element): one block per 256 consecutive elements of one tensor, so the
  // grid holds no idle blocks (a [6] bias does not get the grid width of a [256 x 256] weight)
  // both tables are written by the host when the program is built, never by a kernel: read them
  // (two dependent L2 round trips) before waiting for the previous launch, not after
  const int2 bt = blocks[blockIdx.x];
  const AdamSeg sg = segs[bt.x];
  ptx::pdl_wait();
  const bool reduce = kDP && cm.world > 1 && (mode & 1);
  const unsigned int epoch = static_cast<unsigned int>(st->step[sg.opt]);
  if (reduce) {
    if (static_cast<int>(threadIdx.x) < cm.world) {
      if (blockIdx.x == 0) comm_signal(cm, 0, epoch, threadIdx.x);
      comm_wait(cm, 0, epoch, threadIdx.x);
    }
    __syncthreads();
  }
  const float s_step_size = st->step_size[sg.opt];
  const float s_bc2_sqrt = st->bc2_sqrt[sg.opt];
  // one element: all-reduce (rank order) -> Adam -> Polyak; returns the new online / target values
  auto element = [&](int i, float& p, float& tp) {
    p = ar.theta[sg.goff + i];
    if (mode & 1) {
      float g;
      if (reduce && gp_mt > 0 && sg.gpart) {
        // deferred layer-0 gradient under data parallelism: every rank's per-M-tile partials (behind its exported
        // arena), all requested before the first add; per rank in tile order, then the ranks in rank order --
        // the association the non-deferred path has (epilogue: tiles, Adam: ranks)
        int r = i, c = sg.gp_ones;
        if (c < 0) {
          r = i / sg.cols;
          const int j = i - r * sg.cols;
          c = j < sg.split ? j + sg.off_lo : j - sg.split + sg.off_hi;
        }
        const size_t off = static_cast<size_t>(sg.gp_off) + static_cast<size_t>(r) * sg.wt_rows + c;
        const size_t stride = static_cast<size_t>(sg.w_rows) * sg.wt_rows;
        g = 0.f;
        for (int t0 = 0; t0 < gp_mt; t0 += 2) {
          float pv[kMaxRanks][2];
#pragma unroll
          for (int q = 0; q < kMaxRanks; ++q)
#pragma unroll
            for (int u = 0; u < 2; ++u)
              pv[q][u] = (q < cm.world && t0 + u < gp_mt) ? cm.peer_grad[q][off + (t0 + u) * stride] : 0.f;
          // (more than two tiles per rank, batch > 256: the rank sums are continued tile pair by tile pair, which
          // changes the association against the single-learner order by rounding only)
#pragma unroll
          for (int q = 0; q < kMaxRanks; ++q)
            if (q < cm.world) g += pv[q][0] + pv[q][1];
        }
        ar.grad[sg.goff + i] = g;
      } else if (reduce) {
        // every rank's copy of this element is requested before the first add: one NVLink round
        // trip instead of world - 1 dependent ones; the sum itself stays in rank order
        float gr[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) gr[r] = r < cm.world ? cm.peer_grad[r][sg.goff + i] : 0.f;
        g = 0.f;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < cm.world) g += gr[r];
      } else if (gp_mt > 0 && sg.gpart) {
        // layer-0 gradient left as per-M-tile partials by the fused GEMM epilogue: add them here, in tile order
        // (what the epilogue's last CTA did behind an arrival ticket), and keep the arena complete
        int r = i, c = sg.gp_ones;
        if (c < 0) {
          r = i / sg.cols;
          const int j = i - r * sg.cols;
          c = j < sg.split ? j + sg.off_lo : j - sg.split + sg.off_hi;
        }
        const float* pp = sg.gpart + static_cast<size_t>(r) * sg.wt_rows + c;
        const size_t stride = static_cast<size_t>(sg.w_rows) * sg.wt_rows;
        g = 0.f;
        for (int t = 0; t < gp_mt; t += 4) {  // four tiles in flight (a 256-row batch has two, 1024 rows eight)
          const float p0 = __ldcg(pp + t * stride);
          const float p1 = t + 1 < gp_mt ? __ldcg(pp + (t + 1) * stride) : 0.f;
          const float p2 = t + 2 < gp_mt ? __ldcg(pp + (t + 2) * stride) : 0.f;
          const float p3 = t + 3 < gp_mt ? __ldcg(pp + (t + 3) * stride) : 0.f;
          g += p0;
          g += p1;
          g += p2;
          g += p3;
        }
        ar.grad[sg.goff + i] = g;
      } else {
        g = ar.grad[sg.goff + i];
      }
      float m = ar.m[sg.goff + i];
      float v = ar.v[sg.goff + i];
      m = fmaf(hp.w1, g - m, m);                              // exp_avg.lerp_(grad, 1 - beta1)
      v = __fadd_rn(__fmul_rn(v, hp.beta2f), __fmul_rn(__fmul_rn(hp.w2, g), g));  // mul_().addcmul_()
      const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s_bc2_sqrt), hp.eps);
      p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-s_step_size, m), denom));  // addcdiv_: p + (value*m)/denom
      ar.m[sg.goff + i] = m;
      ar.v[sg.goff + i] = v;
      ar.theta[sg.goff + i] = p;
    }
    tp = 0.f;
    if (ar.target) {
      tp = ar.target[sg.goff + i];
      if (mode & 2) {
        tp = __fadd_rn(__fmul_rn(hp.tau, p), __fmul_rn(hp.one_minus_tau, tp));
        ar.target[sg.goff + i] = tp;
      }
    }
  };
  if (!kPatch || bt.y >= 0) {
    const int i = bt.y + threadIdx.x;
    if (i < sg.n) {
      float p, tp;
      element(i, p, tp);
      if (sg.w) {
        const int r = i / sg.cols;
        const int j = i - r * sg.cols;
        const int c = j < sg.split ? j + sg.off_lo : j - sg.split + sg.off_hi;
        if (mode & 4) {
          sg.w[ct_index(sg.w_rows, r, c)] = p;
          if (sg.wt) sg.wt[ct_index(sg.wt_rows, c, r)] = p;
        }
        if (sg.tw && (mode & 8)) sg.tw[ct_index(sg.w_rows, r, c)] = tp;
      }
    }
  } else {
    // Patch mode (large plain matrices, rows % 32 == cols % 32 == 0, identity column map): this block owns the
    // 32 x 32 patch -1 - bt.y of the row-major tensor -- 128-byte row segments on the read side, and on the write
    // side exactly one contiguous 4 KB run of each tiled copy (W: 4 row-blocks of one column block; W^T: 4 row-blocks
    // of the transposed one), stored as float4 through a shared-memory transpose.  The element-wise path above
    // scatters 4-byte stores over 16-byte core rows: 8x the L2 write sectors for W^T (TQC: 46 us of Adam).
    // (dynamic shared memory, requested only by launches that have patches: a static 8 KB in every Adam block kept the
    // next GEMM's 214 KB CTAs from becoming resident under programmatic dependent launch -- +1..2 us per update)
    extern __shared__ float adam_patch_smem[];
    float (*pp)[33] = reinterpret_cast<float (*)[33]>(adam_patch_smem);
    float (*pt)[33] = reinterpret_cast<float (*)[33]>(adam_patch_smem + 32 * 33);
    const int pc_n = sg.cols >> 5;
    const int patch = -1 - bt.y;
    const int r0 = (patch / pc_n) << 5, c0 = (patch % pc_n) << 5;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // all 4 x 5 loads of this thread are issued before the first store (the element-wise lambda interleaves loads
    // and stores through pointers the compiler must assume to alias: four serial DRAM round trips per block)
    float ep[4], eg[4], em[4], ev[4], et[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = (r0 + wrp + 8 * q) * sg.cols + c0 + lane;
      ep[q] = ar.theta[sg.goff + i];
      et[q] = ar.target ? ar.target[sg.goff + i] : 0.f;
      if (mode & 1) {
        em[q] = ar.m[sg.goff + i];
        ev[q] = ar.v[sg.goff + i];
        if (reduce) {
          float gr[kMaxRanks];
#pragma unroll
          for (int r = 0; r < kMaxRanks; ++r) gr[r] = r < cm.world ? cm.peer_grad[r][sg.goff + i] : 0.f;
          float g = 0.f;
#pragma unroll
          for (int r = 0; r < kMaxRanks; ++r)
            if (r < cm.world) g += gr[r];
          eg[q] = g;
        } else {
          eg[q] = ar.grad[sg.goff + i];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rr = wrp + 8 * q;
      const int i = (r0 + rr) * sg.cols + c0 + lane;
      float p = ep[q], tp = et[q];
      if (mode & 1) {
        const float g = eg[q];
        const float m = fmaf(hp.w1, g - em[q], em[q]);
        const float v = __fadd_rn(__fmul_rn(ev[q], hp.beta2f), __fmul_rn(__fmul_rn(hp.w2, g), g));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s_bc2_sqrt), hp.eps);
        p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-s_step_size, m), denom));
        ar.m[sg.goff + i] = m;
        ar.v[sg.goff + i] = v;
        ar.theta[sg.goff + i] = p;
      }
      if (ar.target && (mode & 2)) {
        tp = __fadd_rn(__fmul_rn(hp.tau, p), __fmul_rn(hp.one_minus_tau, tp));
        ar.target[sg.goff + i] = tp;
      }
      pp[rr][lane] = p;
      pt[rr][lane] = tp;
    }
    __syncthreads();
    const int t = threadIdx.x;
    const int hi = t >> 6, mid = (t >> 3) & 7, lo = t & 7;
    if (mode & 4) {
      float4* wdst = reinterpret_cast<float4*>(sg.w + ct_index(sg.w_rows, r0, c0));
      wdst[t] = make_float4(pp[hi * 8 + lo][mid * 4], pp[hi * 8 + lo][mid * 4 + 1], pp[hi * 8 + lo][mid * 4 + 2],
                            pp[hi * 8 + lo][mid * 4 + 3]);
      if (sg.wt) {
        float4* tdst = reinterpret_cast<float4*>(sg.wt + ct_index(sg.wt_rows, c0, r0));
        tdst[t] = make_float4(pp[mid * 4][hi * 8 + lo], pp[mid * 4 + 1][hi * 8 + lo], pp[mid * 4 + 2][hi * 8 + lo],
                              pp[mid * 4 + 3][hi * 8 + lo]);
      }
    }
    if (sg.tw && (mode & 8)) {
      float4* wdst = reinterpret_cast<float4*>(sg.tw + ct_index(sg.w_rows, r0, c0));
      wdst[t] = make_float4(pt[hi * 8 + lo][mid * 4], pt[hi * 8 + lo][mid * 4 + 1], pt[hi * 8 + lo][mid * 4 + 2],
                            pt[hi * 8 + lo][mid * 4 + 3]);
    }
  }
  if (lt.part && blockIdx.x == 0) {
    __shared__ float lt_sh[kAdamThreads / 32];
    float acc = 0.f;
    const float b3 = lt.b3[0];
    for (int m = threadIdx.x; m < lt.B; m += kAdamThreads) {
      float q = b3;
      for (int p = 0; p < lt.nparts; ++p) q += __ldcg(lt.part + static_cast<size_t>(p) * lt.ld + m);
      acc += q;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) lt_sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < kAdamThreads / 32; ++w) tot += lt_sh[w];
      *lt.out = lt.scale * tot;
    }
  }
  if (lt.pub && blockIdx.x == 0) {
    __syncthreads();  // actor_loss above is part of what gets published
    publish_state(st, lt.pub, threadIdx.x, kAdamThreads);
  }
  if (reduce && cm.exit_barrier) {
    // nobody overwrites gradients another rank may still be reading: the last block of this launch
    // announces "done reading" and waits for the same from every peer
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(cm.done_counter, 1u) == gridDim.x - 1) {
        *cm.done_counter = 0u;
        for (int r = 0; r < cm.world; ++r) comm_signal(cm, 1, epoch, r);
        for (int r = 0; r < cm.world; ++r) comm_wait(cm, 1, epoch, r);
      }
    }
  }
}

}  // namespace oprl
