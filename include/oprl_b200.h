/* oprl_b200 -- C ABI of the B200-native off-policy update engine.
 *
 * The reference (schatty/oprl) has no FFI: its hot path is three duck-typed Python
 * protocols.  This header is the boundary a maintainer would bind (ctypes / cffi)
 * to replace that path; each entry point cites the reference interface it stands
 * in for (paths relative to the reference root).
 *
 *   ReplayBufferProtocol.sample      src/oprl/buffers/protocols.py:19-21,
 *                                    src/oprl/buffers/episodic_buffer.py:114-133
 *   AlgorithmProtocol.update         src/oprl/algos/protocols.py:31-38,
 *                                    src/oprl/algos/{ddpg.py:61-107,td3.py:71-146,
 *                                    sac.py:75-155,tqc.py:116-189}
 *   AlgorithmProtocol.create         src/oprl/algos/protocols.py:27 (parameter/optimizer state)
 *   soft_update                      src/oprl/algos/nn_functions.py:5-10
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success
 * and a negative code on failure with a message in oprl_last_error() (thread
 * local).  Device memory for parameters, optimizer state, replay storage and the
 * sampled batch is owned by the caller (e.g. torch tensors) and only borrowed;
 * the engine owns its activation workspace.  All work is enqueued on the
 * engine's stream; oprl_sync() or a scalar read-back waits for it.
 */
#ifndef OPRL_B200_H_
#define OPRL_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oprl_engine oprl_engine;

enum { OPRL_ALGO_DDPG = 0, OPRL_ALGO_TD3 = 1, OPRL_ALGO_SAC = 2, OPRL_ALGO_TQC = 3 };
/* GEMM arithmetic: 3xTF32 on tcgen05 is fp32-accurate (parity mode); single-pass TF32
 * is the labelled fast mode; SIMT is the FFMA cross-check of the same data path. */
enum { OPRL_GEMM_TC_3XTF32 = 0, OPRL_GEMM_TC_TF32 = 1, OPRL_GEMM_SIMT = 2 };
enum { OPRL_NET_ACTOR = 0, OPRL_NET_CRITIC = 1 };
/* oprl_update flags */
enum { OPRL_UPDATE_ACTOR = 1 /* also run the actor step (+ Polyak updates tied to it) */ };
/* oprl_update segments (multi-GPU learners all-reduce the gradient arena between them) */
enum {
  OPRL_SEG_ALL = -1,
  OPRL_SEG_CRITIC_GRAD = 0,           /* target-Q, critic forward/backward -> critic gradient arena */
  OPRL_SEG_CRITIC_STEP_ACTOR_GRAD = 1, /* critic Adam(+Polyak), actor forward/backward -> actor gradient arena */
  OPRL_SEG_ACTOR_STEP = 2              /* actor Adam (+Polyak), temperature step */
};

typedef struct oprl_cfg {
  int algo;          /* OPRL_ALGO_* */
  int state_dim;     /* ddpg.py:19 */
  int action_dim;    /* ddpg.py:20 */
  int actor_hidden;  /* 256: ddpg.py:43 */
  int actor_layers;  /* hidden layers: 2 */
  int critic_hidden; /* 256: nn_models.py:31 ; TQC 512: tqc.py:49 */
  int critic_layers; /* hidden layers: 2 ; TQC 3 */
  int n_critics;     /* DDPG 1, TD3/SAC 2, TQC n_nets (tqc.py:73) */
  int n_quantiles;   /* TQC 25 (tqc.py:72), else 1 */
  int top_quantiles_to_drop; /* tqc.py:71 */
  int tune_alpha;    /* sac.py:22 ; TQC always 1 */
  int gemm_mode;     /* OPRL_GEMM_* */
  int device;        /* CUDA ordinal */
  int world_size;    /* data-parallel learners sharing the minibatch (losses are means over world_size * B rows) */
  /* python floats of the reference dataclasses, as doubles (rounded to fp32 where torch would) */
  double gamma, tau, lr_actor, lr_critic, lr_alpha;
  double policy_noise, noise_clip, max_action; /* td3.py:21-28 */
  double alpha_init, target_entropy;           /* sac.py:27,70 */
  unsigned long long seed;                     /* device-side sampling / noise streams */
} oprl_cfg;

typedef struct oprl_state { /* optimizer / RNG counters for checkpoint + tests */
  unsigned long long tick;
  int step_actor, step_critic, step_alpha, pad;
  double log_alpha, m_alpha, v_alpha;
} oprl_state;

const char* oprl_last_error(void);
int oprl_abi_version(void);

int oprl_engine_create(const oprl_cfg* cfg, oprl_engine** out);
void oprl_engine_destroy(oprl_engine* e);

/* Number of fp32 elements of a network group's flat parameter arena, laid out
 * exactly as torch's parameters() of the reference module: per net, per layer,
 * weight [out,in] then bias [out] (nn_models.py:98-104).  The GRADIENT array of each group
 * must be OPRL_GRAD_TAIL floats longer: the tail carries per-rank partial sums (temperature
 * loss) that ride along with the data-parallel gradient all-reduce. */
#define OPRL_GRAD_TAIL 8
long long oprl_engine_arena_floats(const oprl_engine* e, int net);
/* Borrow the caller's flat device arrays (each `arena_floats` long).  theta_target
 * may be NULL for groups without a target network (SAC/TQC actor). */
int oprl_engine_bind_arena(oprl_engine* e, int net, float* theta, float* grad, float* m,
                           float* v, float* theta_target);
/* Re-derive the engine's tiled tf32 operand copies after the caller edited theta /
 * theta_target (load_state_dict, initial weights). */
int oprl_engine_sync_params(oprl_engine* e);

/* Replay storage (device pointers), shapes as episodic_buffer.py:29-55:
 * states [E, L+1, S], actions [E, L, A], rewards [E, L, 1], dones [E, L, 1]. */
int oprl_buffer_bind(oprl_engine* e, const float* states, const float* actions,
                     const float* rewards, const float* dones, int E, int L);
/* Host prefix sums of episode lengths (n_eps + 1 ints) for on-device index sampling.
 * Also the ordering point for replay writes: transitions written to the bound storage become
 * visible to oprl_step's device-side sampling through this call (or oprl_buffer_bind) -- the first
 * gather after it is ordered behind everything enqueued on the launch stream so far, later gathers
 * may run ahead of the update in flight (see oprl_step). */
int oprl_buffer_set_prefix(oprl_engine* e, const int* prefix_host, int n_eps);
/* n-step return assembly in the gather (BASELINE north_star; the reference keeps `gamma` in the buffer,
 * src/oprl/buffers/episodic_buffer.py:18, but assembles 1-step transitions only -- this is an extension, off by
 * default).  With n_step > 1, oprl_sample / oprl_step return for a drawn (episode, t):
 *   reward R = sum_{k<m} gamma^k r_{t+k}, next_state = s_{t+m}, done d' = 1 - (1 - d_{t+m-1}) gamma^(m-1),
 *   m = min(n_step, steps up to and including the first done, steps left in the episode),
 * so the unchanged 1-step TD target r + (1 - d') gamma Q'(s') equals the n-step target.  Needs
 * oprl_buffer_set_prefix (episode lengths).  n_step = 1 restores the reference's transitions bit for bit. */
int oprl_buffer_set_nstep(oprl_engine* e, int n_step, double gamma);

/* Row-major batch arrays the gather writes (what sample() returns): device pointers
 * s [cap,S], a [cap,A], r [cap], d [cap], s2 [cap,S]. */
int oprl_batch_bind(oprl_engine* e, float* s, float* a, float* r, float* d, float* s2, int cap);

/* sample(): gather B transitions.  ep_step_host = B (episode, step) int pairs chosen
 * by the host (reference RNG parity), or NULL to draw uniformly on the device. */
int oprl_sample(oprl_engine* e, const int* ep_step_host, int B);
/* Generic path: load a caller-provided dense device batch instead of gathering. */
int oprl_load_batch(oprl_engine* e, const float* s, const float* a, const float* r,
                    const float* d, const float* s2, int B);
/* Same with HOST arrays (fp32, contiguous): packed into a pinned staging ring (the arrays may be
 * reused as soon as the call returns), copied H2D and loaded into the engine's other working set on
 * a copy stream, beside the update still running on the launch stream -- the path a CPU-resident
 * replay buffer / trainer would use. */
int oprl_load_batch_host(oprl_engine* e, const float* s, const float* a, const float* r,
                         const float* d, const float* s2, int B);
/* Inject the standard-normal draws of the next update (parity tests).  which = 0:
 * first draw of the update (TD3 smoothing noise / SAC-TQC next-action noise),
 * 1: second draw (SAC-TQC actor-step noise).  noise: device pointer [B, A]. */
int oprl_set_noise(oprl_engine* e, int which, const float* noise_dev, int n);

/* update(): one gradient update on the batch last sampled / loaded (calling it again repeats the
 * update on the same batch). */
int oprl_update(oprl_engine* e, int flags, int segment);
/* Device-resident learner step = oprl_sample(e, NULL, B) + oprl_update(e, flags, OPRL_SEG_ALL), the
 * loop body of distrib/policy_update_worker.py:66-68 without the host round trip.  In a steady loop
 * (same B, no oprl_buffer_set_prefix / _bind / oprl_set_noise in between) the batch of step t+1 is
 * gathered as a parallel branch of step t's update graph; the uniform draws are the ones the
 * in-order sequence would have made (the RNG offset is the update count, not the launch order).
 * The row-major batch arrays of oprl_batch_bind are NOT written on this path. */
int oprl_step(oprl_engine* e, int B, int flags);

/* Logging scalars (synchronises the stream):
 * 0 critic_loss 1 actor_loss 2 alpha_loss 3 mean q 4 mean q_target 5 mean logpi
 * 6 mean (q - q_target) 7 alpha */
int oprl_get_scalars(oprl_engine* e, float* out_host, int n);
/* Pipelined form of the same read-back (the reference logs its losses with a blocking `.item()`
 * per scalar, algos/td3.py:118-131, sac.py:112-150; a learner loop can instead consume the scalars
 * of update t while update t+1 is already running): _enqueue appends a D2H copy of the scalars of
 * everything launched so far to the stream and returns a ticket >= 0; _wait blocks on that copy
 * only.  A ticket stays valid for 8 further enqueues / updates.  In a loop that reads every update the
 * engine switches to a program variant whose last kernel publishes the scalars into a pinned,
 * device-mapped ring itself: _enqueue then costs no GPU work and _wait polls host memory
 * (single-learner engines; OPRL_B200_HOST_SCALARS=0 keeps the D2H copy). */
int oprl_scalars_enqueue(oprl_engine* e);
int oprl_scalars_wait(oprl_engine* e, int ticket, float* out_host, int n);
int oprl_get_state(oprl_engine* e, oprl_state* out);
int oprl_set_state(oprl_engine* e, const oprl_state* in);
int oprl_sync(oprl_engine* e);
void* oprl_stream(oprl_engine* e); /* cudaStream_t the engine currently launches into */
/* Redirect the engine's launches to the caller's stream (e.g. torch's current stream) so that
 * ordinary stream ordering holds between the caller's tensors and the engine; NULL = engine-owned. */
int oprl_engine_set_stream(oprl_engine* e, void* stream);

/* Data-parallel learners: every rank runs the same program on its B rows of a world_size * B
 * minibatch; all losses / gradient seeds are scaled by 1 / (world_size * B) so that an
 * all-reduce(SUM) of the gradient arenas between the OPRL_SEG_* segments yields the gradients
 * of the global-mean losses. */
int oprl_engine_set_world_size(oprl_engine* e, int world_size);

/* Fused NVLink all-reduce (one process per GPU, one node): instead of an NCCL call between the
 * segments, every rank maps its peers' gradient arenas through CUDA IPC and the Adam kernel sums
 * them in rank order after a flag handshake.  oprl_comm_init allocates this rank's exportable
 * gradient arenas + flag block and writes three 64-byte IPC handles to handles_out;
 * exchange them (e.g. torch.distributed.all_gather_object) and pass all world x 3 handles, rank
 * major, plus each rank's CUDA device ordinal to oprl_comm_connect.  Afterwards oprl_update(...,
 * OPRL_SEG_ALL) performs the whole data-parallel update. */
int oprl_comm_init(oprl_engine* e, int rank, int world, void* handles_out);
int oprl_comm_connect(oprl_engine* e, const void* all_handles, const int* device_of_rank);

/* Engine-less sample(): plain row-major gather of B host-chosen (episode, step) pairs out of
 * replay storage shaped as in oprl_buffer_bind (episodic_buffer.py:127-133); next_state is the
 * adjacent row states[ep, step + 1].  ep_step_dev: device scratch of 2*B ints. */
int oprl_gather_rows(const float* states, const float* actions, const float* rewards,
                     const float* dones, int E, int L, int S, int A, const int* ep_step_host,
                     int* ep_step_dev, int B, float* s, float* a, float* r, float* d, float* s2,
                     void* stream);
/* Ingest edge -- EpisodicReplayBuffer.add_transition / add_episode (src/oprl/buffers/episodic_buffer.py:81-112).
 * `staged_dev` holds n transitions already copied to the device, one row of S + A + 4 floats each:
 * [state | action | reward | done | episode index (int32 bits) | step index (int32 bits)]; one launch writes them
 * into the replay storage (states [E, L+1, S], actions [E, L, A], rewards / dones [E, L, 1]).  The caller
 * guarantees 0 <= episode < E and 0 <= step < L (the Python buffer stages them from its own ring bookkeeping). */
int oprl_scatter_transitions(float* states, float* actions, float* rewards, float* dones, int E, int L, int S, int A,
                             const float* staged_dev, int n, void* stream);

/* number of kernel launches one oprl_update(flags, OPRL_SEG_ALL) enqueues at batch B */
int oprl_update_launches(oprl_engine* e, int B, int flags);

/* Measurement hook for bench.py's roofline: what = 0 replays only the tcgen05 GEMM launches of
 * one update `iters` times, what = 2 only its SIMT launches, what = 3 only its batch-slice chain launches
 * (csrc/chain.cuh), what = 1 the gather (device index draw); total milliseconds by CUDA
 * events on the launch stream.  Leaves activations / batch in an unspecified state. */
int oprl_profile(oprl_engine* e, int B, int flags, int what, int iters, float* ms_total,
                 int* launches_per_iter);
/* Debug hook (OPRL_B200_CHAIN_PROF=1): 256 clock64 stamps of CTA 0 of the critic-step (which = 0) or
 * actor-step (1) chain launch of the last update; no reference counterpart. */
int oprl_chain_prof(oprl_engine* e, int B, int flags, int which, long long* out64);

#ifdef __cplusplus
}
#endif
#endif /* OPRL_B200_H_ */
