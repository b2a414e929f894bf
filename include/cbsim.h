/*
 * cbsim.h — C ABI of libcbsim.so: the B200 batched continuous-env step path for C-CyberBattleSim.
 *
 * Plain C, plain pointers and sizes, no torch / C++ types.  One handle per GPU, one host thread per
 * handle.  Every call returns 0 on success or a negative cbs_status; the message is available through
 * cbs_last_error().  No C++ exception crosses this boundary.  Functions that take a `stream` enqueue
 * work on that CUDA stream (pass the integer value of a cudaStream_t; 0 = legacy default stream) and
 * return without synchronising, unless stated otherwise.
 *
 * Each entry point names the reference interface (zsh239040/C-CyberBattleSim, path relative to
 * cyberbattle/) it replaces.  There is no CPU fallback: if no CUDA device is usable cbs_create fails.
 *
 * Pointer suffixes:  _host = host memory, _dev = device memory on the handle's GPU.
 */
#ifndef CBSIM_H
#define CBSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBS_ABI_VERSION 6

/* dimensions fixed by the reference's defaults (agents/config/train_config.yaml:19,32;
 * gae/config/train_config.yaml:8,11; _env/cyberbattle_env_compressed.py:112-142) */
#define CBS_NODE_EMB_DIM 64
#define CBS_VULN_EMB_DIM 768
#define CBS_OUTCOME_DIM 9
#define CBS_ACTION_DIM 905 /* 64 + 64 + 768 + 9 */
#define CBS_OBS_DIM 194    /* graph_embeddings[192] (mean|max|min) + discrete_features[2] (discovered, owned) */
#define CBS_OBS_DIM_NODE_GOAL 258 /* *_node goals: + the interest node's embedding[64] before the discrete features (compressed:119-125) */
#define CBS_NUM_STATS 14   /* cyberbattle_env.py:517-524 get_statistics() tuple */
#define CBS_NUM_REWARDS 10
#define CBS_NUM_PENALTIES 18
#define CBS_MAX_NODES 128
#define CBS_INFO_INTS 8

typedef enum {
  CBS_OK = 0,
  CBS_ERR_INVALID_ARG = -1,
  CBS_ERR_CUDA = -2,
  CBS_ERR_NOT_READY = -3, /* scenarios not loaded / env not reset */
  CBS_ERR_CAPACITY = -4,  /* a per-env capacity (snapshot slots, edges) overflowed on the device */
  CBS_ERR_NO_DEVICE = -5
} cbs_status;

typedef struct cbs_handle cbs_handle;

/* Environment configuration == kwargs of CyberBattleEnv.__init__ (_env/cyberbattle_env.py:38-60) and
 * CyberBattleCompressedEnv.__init__ (_env/cyberbattle_env_compressed.py:74-89) that the step path reads,
 * plus the reward tables of agents/config/rewards_config.yaml in the order of constants.REWARD_KEYS /
 * PENALTY_KEYS. */
typedef struct {
  int32_t abi_version;          /* must be CBS_ABI_VERSION */
  int32_t device;               /* CUDA device ordinal */
  int32_t num_envs;             /* envs held by this handle (this GPU's shard) */
  int64_t global_env_offset;    /* index of env 0 in the whole batch: keys the Philox streams so results do not depend on the GPU count */
  uint64_t seed;                /* Philox key */
  int32_t goal;                 /* 0 control, 1 discovery, 2 disruption, 3 control_node, 4 discovery_node, 5 disruption_node
                                   (cyberbattle_env.py:467-514); 3..5 need cbs_scenario_tables.sc_interest */
  int32_t episode_iterations;   /* constant cut-off (cyberbattle_env.py:366) */
  double proportional_cutoff_coefficient; /* 0 = off (cyberbattle_env.py:361,457-460) */
  double winning_reward, losing_reward;
  int32_t absolute_reward;      /* cyberbattle_env.py:379 */
  int32_t stop_at_goal_reached; /* cyberbattle_env.py:348 */
  int32_t remove_main_obstacles, remove_all_obstacles; /* compressed:536-543 */
  int32_t switch_interval;      /* RandomSwitchEnv.switch_interval (cyberbattle_env_switch.py:218-220): a reset draws a new scenario when
                                   (episodes finished + 1) % (switch_interval + 1) == 0, so 0 = at every reset; < 0 = never */
  int32_t auto_reset;           /* 1: cbs_observe resets finished envs in place (DummyVecEnv semantics) */
  double rewards[CBS_NUM_REWARDS];
  double penalties[CBS_NUM_PENALTIES];
  int32_t max_slots;            /* embedding-snapshot slots per env; 0 = derive from the cut-offs */
  int32_t max_edges;            /* visible-graph edges per env; 0 = derive from the cut-offs */
  float decode_margin;          /* cosine-score margin inside which candidates are re-scored in float64; 0 = default */
  int32_t decode_gemm;          /* 0 = default (tcgen05 TF32 when built in), 1 = force the SIMT fp32 path */
  /* static defender (cyberbattle_env.py:52,331-332,416-430): 0 none, 1 ScanAndReimageCompromisedMachines
   * (_env/static_defender.py:27-60: every scan_frequency steps scan_capacity nodes are drawn with replacement; an owned,
   * Running node without defense evasion is detected with detect_probability and, if re-imageable, goes to Imaging for
   * 15 steps, static_defender_actions.py:19,37-68).  Needs cbs_scenario_tables.nd_reimageable. */
  int32_t static_defender;      /* 2 = ExternalRandomEvents (_env/static_defender.py:63-161): every step every node without defense
                                   evasion suffers, with random_event_probability, one of start / stop a service, remove / add a
                                   firewall rule (simulation/static_defender_actions.py:96-168).  Needs nd_ev_init, vi_svc_slot,
                                   out_slot of cbs_scenario_tables and ev_proj of cbs_gae_tables. */
  int32_t scan_capacity;        /* 1..8 */
  int32_t scan_frequency;       /* >= 1 */
  int32_t precise_graph_encoding; /* compressed:455-462: re-encode the visible graph on every step */
  double detect_probability;
  /* compressed:86,419-427,498-506: at every table-maintaining encode of a step, the rows of the (source, target) pairs
   * from which the action's source or target node can be reached in the visible graph are overwritten with the current
   * node embeddings (their place in the table's insertion order is kept).  Under the re-imaging defender a step whose desired
   * outcome does not re-encode by itself refreshes around the nodes the defender re-imaged in that step (:423-427); with the
   * events defender the reference itself raises (nx.has_path on nodes outside the visible graph), so that pair is rejected. */
  int32_t precise_action_space_positions;
  /* compressed:82,570-590 `distance_metric` of find_closest_action_embedding: 0 'cosine' (scipy cdist, the reference's default),
   * 1 'l1', 2 'l2', 3 'inf' (np.linalg.norm(action - rows, ord, axis=1), :571-576).  Anything else is rejected like the
   * reference's ValueError (:578-579).  1..3 decode in float64 (k_decode_metric.cu) and run the transition as its own launch. */
  int32_t distance_metric;
  /* compressed:83,521-522,553-567 `sample_subset_samples` (100 in agents/config/train_config.yaml:29): after every table-maintaining
   * encode at most this many action-table rows per outcome class stay in the table; 0 = the whole table.  The reference draws
   * the subset with np.random.choice; here row i of an over-full class gets the key philox(seed, global env index, c, identity_i).x
   * with c = the env's lifetime count of balance calls and identity = 0x80000000 | source << 23 | target << 16 | kind << 12 |
   * scenario-local vulnerability index, and the k smallest (key, table position) stay (ccbs_b200.philox.subset_keep). */
  int32_t sample_subset_samples;
  double random_event_probability; /* static_defender == 2 (train_config.yaml random_event_probability_min / _max) */
} cbs_config;

/* Immutable scenario tables, produced by ccbs_b200.scenario.compile_scenarios (host arrays; copied to the
 * device by cbs_load_scenarios).  They flatten simulation/model.py NodeInfo / VulnerabilityInfo /
 * PredictedResult / FirewallConfiguration and the reach counts of cyberbattle_env.py:205-217. */
typedef struct {
  int32_t num_scenarios, max_nodes, words;
  int32_t num_nodes_total, num_inst, num_rows, num_recon, num_ports_total, num_uvuln_total, num_global_vulns;
  int64_t num_instof;
  const int32_t* sc_num_nodes;           /* [S] */
  const int32_t* sc_node_off;            /* [S+1] */
  const int32_t* sc_port_off;            /* [S+1] */
  const int32_t* sc_uvuln_off;           /* [S+1] */
  const int32_t* sc_num_uvuln;           /* [S] */
  const int64_t* sc_instof_off;          /* [S+1] */
  const int32_t* sc_discoverable_amount; /* [S] */
  const uint32_t* sc_init_has_data;      /* [S][words] */
  const uint32_t* sc_init_visible;       /* [S][words] */
  const int32_t* sc_feasible_off;        /* [S+1] for the configured goal */
  const int32_t* feasible_starters;      /* [sc_feasible_off[S]] */
  int32_t num_feasible;
  const int32_t* sc_interest;            /* [S] interest node of each scenario (cyberbattle_env.py:127-131), NULL for the network-wide goals */
  const int32_t* nd_value;               /* [Nn] */
  const uint8_t* nd_level_at_access;     /* [Nn] */
  const uint8_t* nd_reimageable;         /* [Nn] NodeInfo.reimageable (model.py:312); only read with a static defender */
  const int32_t* nd_ownable;             /* [Nn] */
  const int32_t* nd_discoverable;        /* [Nn] */
  const int32_t* nd_disruptable;         /* [Nn] */
  const int32_t* nd_row_off;             /* [2*Nn+1] local|remote candidate lists */
  const uint32_t* outblock;              /* [ports][words] */
  const int32_t* uvuln_global;           /* [num_uvuln_total] scenario-local vuln index -> row of vemb */
  const int32_t* inst_of;                /* [num_instof] */
  const int32_t* vi_port;                /* [I] */
  const uint32_t* vi_flags;              /* [I] */
  const uint16_t* vi_kinds_any;          /* [I] */
  const uint16_t* vi_kinds_remote;       /* [I] */
  const double* vi_success;              /* [I] */
  const double* vi_cost;                 /* [I] */
  const int32_t* vi_recon_any;           /* [I][2] */
  const int32_t* vi_recon_remote;        /* [I][2] */
  const int32_t* vi_ulocal;              /* [I] */
  const uint8_t* recon_nodes;            /* [num_recon] */
  const uint32_t* row_packed;            /* [R] */
  const int32_t* row_inst;               /* [R] */
  const float* vemb32;                   /* [Ug][768] */
  const double* vemb64;                  /* [Ug][768] */
  const double* vnorm2;                  /* [Ug] */
  /* events defender only (may be NULL otherwise) */
  const uint16_t* nd_ev_init;            /* [Nn][4] running services, incoming BLOCK, outgoing BLOCK (bit i = the node's i-th service / its port), service count */
  const uint8_t* vi_svc_slot;            /* [I] the target node's service slot of the vulnerability's port, 0xFF = none */
  const uint8_t* out_slot;               /* [ports][max_nodes] node's service slot of a scenario port, 0xFF = none */
} cbs_scenario_tables;

/* Folded graph-encoder tables (ccbs_b200.gae.fold_gae) for gae/model.py:25-82 GAEEncoder.forward with the
 * default layer config, eval mode. */
typedef struct {
  const float* node_static; /* [Nn][2][18][64]: the node's folded rows while not visible / while visible */
  const float* dyn_proj;    /* [6][18][64] */
  const float* vuln_h;      /* [Ug][16] */
  const float* nn0_b;       /* [16] */
  const float* bn1_scale;   /* [64] */
  const float* bn1_shift;   /* [64] */
  const float* gcn_wt;      /* [64][64] (in, out) */
  const float* bn2_scale;   /* [64] */
  const float* bn2_shift;   /* [64] */
  const float* ev_proj;     /* [30][18][64] firewall-in[10] | firewall-out[10] | service-running[10] feature columns (events defender; may be NULL) */
} cbs_gae_tables;

/* ---- lifetime ---------------------------------------------------------------------------------- */
int cbs_abi_version(void);
/* replaces: construction of CyberBattleCompressedEnv objects (utils/envs_utils.py:23-32) for num_envs envs */
int cbs_create(const cbs_config* cfg, cbs_handle** out);
void cbs_destroy(cbs_handle* h);
const char* cbs_last_error(const cbs_handle* h); /* h may be NULL: last error of a failed cbs_create */

/* replaces: pickle.load of scenario envs (cyberbattle_env_switch.py:72-89) + env.set_graph_encoder (compressed:643) */
int cbs_load_scenarios(cbs_handle* h, const cbs_scenario_tables* t, const cbs_gae_tables* g);
/* scenario of every env ([num_envs], host).  replaces RandomSwitchEnv._switch_environment (switch.py:102-106) */
int cbs_set_scenarios(cbs_handle* h, const int32_t* scenario_of_env_host);
/* optional deterministic starter nodes: queue_host[num_envs][qlen]; episode e of env b starts at
 * queue[b][e % qlen].  NULL restores random starters (Philox over the feasible set == the rejection loop of
 * cyberbattle_env.py:195-248). */
int cbs_set_starter_queue(cbs_handle* h, const int32_t* queue_host, int32_t qlen);
/* Row pitch, in floats, of the action tensors handed to cbs_decode / cbs_step (default 905 = dense).  With a pitch
 * that is a multiple of 4 floats the tensor-core contraction reads the tensor in place through TMA; a dense tensor
 * is repacked first (TMA cannot address 3620-byte rows). */
int cbs_set_action_stride(cbs_handle* h, int32_t stride_floats);
/* Declares that the action tensors handed to cbs_decode / cbs_step are COMPLETE well before the call is made - written by work
 * that finished before the previous library call on the stream did (a pre-recorded trace, a ring of pre-generated batches; NOT
 * the output of a policy kernel enqueued just before the call).  The step's kernels are launched with programmatic dependencies;
 * with this declaration the tensor-core contraction of step t+1 reads its actions while step t's observation kernel is still
 * draining (0.101 -> 0.093 ms per step at 8192 envs) instead of waiting for it first.  Off by default (always correct; costs
 * nothing when a policy runs between the steps, since nothing could overlap then); the environment variable
 * CBS_ACTIONS_PRESTAGED=1 turns it on for every handle of the process.  cbs_replay applies it to its steps 1..T-1 by itself
 * (its whole action slab is staged before the call).  No reference counterpart (the reference steps one env on the host). */
int cbs_set_actions_prestaged(cbs_handle* h, int32_t on);
/* (events defender: detect_uniforms_dev is float32 [num_envs][max_nodes][4] = per node { function index 0 start service / 1 firewall
 * remove / 2 stop service / 3 firewall add, event uniform, pick uniform, side uniform } replacing random.choice /
 * numpy.random.random of _env/static_defender.py:80-161; scan_nodes_dev is ignored.) */
/* Test hook for the static defender's randomness: scan_nodes_dev [num_envs][scan_capacity] int32 replaces the
 * random.choices draw of _env/static_defender.py:48, detect_uniforms_dev [num_envs][scan_capacity] float32 replaces the
 * numpy.random.random() calls of :53 (consumed in call order, as the reference consumes its stream).  The pointers are
 * read by every following step until cleared with NULLs (then Philox streams 3.. and 5.. are used). */
int cbs_set_defender_draws(cbs_handle* h, const int32_t* scan_nodes_dev, const float* detect_uniforms_dev);
/* replaces set_cut_off / set_proportional_cutoff_coefficient (cyberbattle_env_switch.py:198-203) */
int cbs_set_cutoffs(cbs_handle* h, int32_t episode_iterations, double proportional_cutoff_coefficient);

/* ---- the step path ----------------------------------------------------------------------------- */
/* replaces RandomSwitchEnv.reset -> CyberBattleCompressedEnv.reset (switch.py:151-167, compressed:158-189,
 * cyberbattle_env.py:134-186).  env_mask_dev: [num_envs] bytes, non-zero = reset; NULL = all.
 * obs_dev (optional): [num_envs][obs_len] float32 (obs_len = 194, or 258 for *_node goals), written for every env. */
int cbs_reset(cbs_handle* h, const uint8_t* env_mask_dev, float* obs_dev, uintptr_t stream);

/* replaces find_closest_action_embedding (compressed:570-590: scipy cdist 'cosine', or np.linalg.norm with ord 1 / 2 / inf
 * for cbs_config.distance_metric 1..3, + argmin over the action table).  actions_dev: [num_envs][905] float32.
 * sel_dev: [num_envs][4] = source node, target node, scenario-local vulnerability index, outcome kind.
 * dist_dev: [num_envs] float64 distance in the configured metric. */
int cbs_decode(cbs_handle* h, const float* actions_dev, int32_t* sel_dev, double* dist_dev, uintptr_t stream);

/* replaces CyberBattleEnv.step_attacker_env (cyberbattle_env.py:299-394) incl.
 * AttackerAgentActions.exploit_{local,remote}_vulnerability (simulation/attacker_actions.py:92-547), goal /
 * termination checks and the distance penalty of compressed:430.
 * sel_dev as produced by cbs_decode (any indices are accepted: invalid ones take the reference's penalty
 * branches).  dist_dev may be NULL (distance 0).  uniforms_dev: [num_envs] float32 success-rate draws, or NULL
 * to draw from Philox(seed, global env index, env step counter).
 * Outputs (each may be NULL): reward_dev float32, done_dev = done|truncated (compressed:451), truncated_dev,
 * outcome_dev = obtained outcome code (ccbs_b200.constants K_* / OC_*). */
int cbs_transition(cbs_handle* h, const int32_t* sel_dev, const double* dist_dev, const float* uniforms_dev,
                   float* reward_dev, uint8_t* done_dev, uint8_t* truncated_dev, uint8_t* outcome_dev,
                   uintptr_t stream);

/* K transitions per env in one launch ("K-step persistent kernel over fixed action traces"): sel_dev [k_steps][num_envs][4],
 * dist_dev [k_steps][num_envs] or NULL, uniforms_dev [k_steps][num_envs] or NULL (Philox), reward_dev / done_dev [k_steps][num_envs]
 * or NULL.  Step k of env b applies sel_dev[k][b] exactly like cbs_transition (same code), but the env's record stays in
 * registers between the steps and is written back once.  No cbs_observe runs between those steps: the visible graph, the action
 * table and the observation are NOT advanced and an env that finishes stays finished (reward 0, done 1) — this is
 * CyberBattleEnv.step_attacker_env (cyberbattle_env.py:299-394) alone, for pre-decoded traces and for the transition roofline.
 * Scenarios of <= 32 nodes, no static defender.  Follow it with cbs_reset before stepping the envs normally again. */
int cbs_transition_ksteps(cbs_handle* h, const int32_t* sel_dev, const double* dist_dev, const float* uniforms_dev, int32_t k_steps,
                          float* reward_dev, uint8_t* done_dev, uintptr_t stream);

/* replaces update_evolving_visible_graph_after_step + encode + create_continuous_action_space
 * (compressed:399-428, 465-550) and, with auto_reset, the VecEnv reset of finished envs.
 * obs_dev: [num_envs][CBS_OBS_DIM] float32, or NULL: the observation cache itself is addressable through
 * cbs_state_ptr(CBS_F_OBS) (zero copy).  The terminal observation of envs that finished in this step stays
 * readable through cbs_read_state(CBS_F_TERMINAL_OBS).  Must follow a cbs_transition (it consumes the worklist
 * that call produced). */
int cbs_observe(cbs_handle* h, float* obs_dev, uintptr_t stream);

/* decode + transition + observe, device buffers.  info_dev (optional): [num_envs][CBS_INFO_INTS] int32 =
 * source, target, vulnerability, desired outcome kind, obtained outcome code, end_episode_reason, step_count,
 * truncated. */
int cbs_step(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, float* obs_dev, float* reward_dev,
             uint8_t* done_dev, int32_t* info_dev, uintptr_t stream);

/* One step like cbs_step (same launches), with a CUDA event recorded on `stream` between the kernels.  out_ms[5] =
 * { decode_gemm, decode_select + transition (one fused kernel), observe, 0, 0 } in milliseconds.  Synchronises the
 * stream.  Measurement aid for bench.py. */
int cbs_profile_step(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, float* out_ms, uintptr_t stream);

/* ---- K-step replay (BASELINE configs[0]: one env, 10 000 steps, parity trace) -------------------------------------------------
 * T whole steps (decode -> transition -> observe with in-place resets, exactly the launches of cbs_step) over pre-staged inputs,
 * with every step's decoded action, outcome, reward, distance, state records and observation written to a device log — no host
 * round trip between the steps, one synchronisation (the caller's) at the end.  actions_dev [T][num_envs][905] float32,
 * uniforms_dev [T][num_envs] float32 or NULL (Philox).  Starters come from cbs_set_starter_queue (or Philox).  The log covers envs
 * [first_env, first_env + num_logged); every pointer is device memory, [T][num_logged][...], and may be NULL to skip the field:
 *   after the transition, before the observe (i.e. before an in-place reset):
 *     sel int32[4] · meta int32[4] = { obtained outcome code, flags (bit0 done, bit1 truncated, bits2-3 end reason, bits 8.. scenario in force), step count,
 *     episodes finished before this step } · reward, dist float64 · masks uint32[mask_pitch] · disc_order uint8[max_nodes] ·
 *     owned_order uint8[owned_len] (owned_len = 2 * max_nodes under a static defender: env.owned_nodes itself, else max_nodes) ·
 *     counters int32[8] = { stepcount, num_iterations, discovered_amount, ownable, discoverable, disruptable, discoverable_amount,
 *     n_discovered | n_owned << 16 }
 *   after the observe:
 *     obs float32[obs_len] (the terminal observation when the episode ended in this step) · and, only meaningful when it did:
 *     reset_obs float32[obs_len], reset_masks uint32[mask_pitch] (first observation / masks of the next episode), stats float64[14]
 * Forced decodes (parity tests: follow the recorded pick on a verified near-tie of two table rows): at every step t with
 * force_steps_host[t] != 0 (host array [num_steps]) the decode and the transition run as separate launches and, in between, every
 * env b with force_sel[t][b][0] >= 0 (device, int32 [num_steps][num_envs][4]) takes that action and force_dist[t][b] (device,
 * float64) instead of its own decode.  All three NULL = no forcing.
 * replaces: the reference's own per-step loop `for t: env.step(action[t])` (agents/test_agent.py, utils/test_utils.py:78-140). */
typedef struct {
  int32_t first_env, num_logged;
  int32_t* sel; int32_t* meta; double* reward; double* dist; uint32_t* masks; uint8_t* disc_order; uint8_t* owned_order;
  int32_t* counters; float* obs; float* reset_obs; uint32_t* reset_masks; double* stats;
  const int32_t* force_sel; const double* force_dist; const uint8_t* force_steps_host;
} cbs_replay_log;
int cbs_replay(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, int32_t num_steps, const cbs_replay_log* log,
               uintptr_t stream);

/* Same through HOST buffers (pinned memory recommended): copies actions in, runs the step, copies
 * obs / reward / done (/ info) out, and synchronises.  This is the call the VecEnv adapter makes. */
int cbs_step_host(cbs_handle* h, const float* actions_host, const float* uniforms_host, float* obs_host,
                  float* reward_host, uint8_t* done_host, int32_t* info_host);

/* The same step enqueued on the handle's own stream WITHOUT the final synchronisation, and the wait for it.  Several
 * handles that each hold a slice of one env batch on the same GPU can be stepped as a pipeline: handle k+1's
 * host-to-device copy runs under handle k's kernels and device-to-host copies (ccbs_b200.host_pipeline.ShardedHostEnv).
 * The host buffers must stay valid (and should be pinned) until cbs_host_sync returns. */
int cbs_step_host_async(cbs_handle* h, const float* actions_host, const float* uniforms_host, float* obs_host,
                        float* reward_host, uint8_t* done_host, int32_t* info_host);
int cbs_host_sync(cbs_handle* h);

/* ---- introspection (parity tests, statistics) ----------------------------------------------------- */
typedef enum {
  CBS_F_MASKS = 0,        /* uint32 [B][mask_pitch]: plane p, word w of env b at b * mask_pitch + p * words + w */
  CBS_F_DISC_ORDER = 1,   /* uint8 [B][max_nodes] */
  CBS_F_OWNED_ORDER = 2,  /* uint8 [B][max_nodes] every node that entered env.owned_nodes, in first-entry order */
  CBS_F_SCALARS = 3,      /* int32 [scalar_pitch / 8][B][8]: sector-major; scalar k of env b at ((k >> 3) * B + b) * 8 + (k & 7) */
  CBS_F_TERMINAL_OBS = 4, /* float32 [B][194] */
  CBS_F_OBS = 5,          /* float32 [B][194] cached observation */
  CBS_F_LAST_STATS = 6,   /* float64 [B][14] get_statistics() of the last finished episode */
  CBS_F_STAT_ACCUM = 7,   /* float64 [CBS_NUM_ACCUM] sums over finished episodes (NCCL all-reduce input) */
  CBS_F_PAIR_SLOT = 8,    /* uint8 [B][max_nodes*max_nodes] */
  CBS_F_DIST = 9,         /* float64 [B] last decode distance */
  CBS_F_REWARD64 = 10,    /* float64 [B] last step reward */
  CBS_F_ERRFLAG = 11,     /* int32 [1] device-side capacity error flag */
  CBS_F_VT = 12,          /* float32 [B][vt_stride] action x vulnerability-embedding products of the last decode */
  CBS_F_OWNED_RAW = 13,   /* uint8 [B][2*max_nodes] env.owned_nodes as the reference holds it under a defender (removals, duplicates) */
  CBS_F_REIMAGE_LEFT = 14,/* uint8 [B][max_nodes] node_reimaging_progress of nodes whose Imaging bit is set */
  CBS_F_Z_HIST = 15,      /* float32 [B][slots][max_nodes][64] node-embedding snapshots the action-table rows refer to (cbs_capacities: slots) */
  CBS_F_SEL = 16,         /* int32 [B][4] last decoded / applied action (source, target, vulnerability, outcome kind); a caller that
                             hands this very buffer to cbs_transition saves the copy */
  CBS_F_EV_CUR = 18,      /* uint16 [B][max_nodes][4] events defender: per node { running services, incoming BLOCK, outgoing BLOCK, - } bit
                             sets over the node's service slots */
  CBS_F_EV_X = 19,        /* uint16 [B][max_nodes][4] the same as cached in the node's feature vector of the visible graph */
  CBS_F_MARGIN_EDGE = 20, /* int32 [1] decodes whose float64 winner had a float32 scan score in the outer half of the re-score margin
                             (cbs_config.decode_margin): 0 in every test and bench run; non-zero = widen the margin */
  CBS_F_DIVERGENCE = 17   /* int32 [1] number of env-steps at which the reference itself raises and this library goes on: the re-imaging
                             defender detecting a persistent node in the very step it comes back (owned_nodes.remove of an absent
                             node, cyberbattle_env.py:425 -> ValueError); the removal is a no-op here */
} cbs_field;
#define CBS_NUM_MASKS 15
#define CBS_NUM_SCALARS 26
#define CBS_NUM_ACCUM 20
/* synchronous device->host copy of one state field; bytes must equal the field size (query with dst NULL). */
int64_t cbs_read_state(cbs_handle* h, int32_t field, void* dst_host, int64_t bytes);
/* All state arrays a parity test or a zero-copy consumer needs, in one call (device pointers, valid for the handle's
 * lifetime; layouts as documented at cbs_field).  SURVEY 8(b): `cbs_get_state(h, cbs_state_view*)`. */
typedef struct {
  int32_t num_envs, max_nodes, words, mask_pitch, scalar_pitch, obs_dim, slots, num_masks;
  uint32_t* masks;         /* CBS_F_MASKS */
  int32_t* scalars;        /* CBS_F_SCALARS (sector-major) */
  uint8_t* disc_order;     /* CBS_F_DISC_ORDER */
  uint8_t* owned_order;    /* CBS_F_OWNED_ORDER */
  uint8_t* pair_slot;      /* CBS_F_PAIR_SLOT */
  float* obs;              /* CBS_F_OBS */
  float* terminal_obs;     /* CBS_F_TERMINAL_OBS */
  int32_t* sel;            /* CBS_F_SEL */
  double* dist;            /* CBS_F_DIST */
  double* reward64;        /* CBS_F_REWARD64 */
  double* last_stats;      /* CBS_F_LAST_STATS */
  double* stat_accum;      /* CBS_F_STAT_ACCUM */
} cbs_state_view;
int cbs_get_state(cbs_handle* h, cbs_state_view* out);

/* get_statistics() (cyberbattle_env.py:517-524) of the last finished episode of every env, copied to the host:
 * out_host[B][14] = owned, discovered, not_discovered, disrupted, num_nodes, ownable, discoverable, disruptable,
 * network_availability, reimaged, num_events, discovered_amount, discoverable_amount, goal_reached.  float64 because
 * network_availability is a ratio (SURVEY 8(b) sketched int64).  Synchronises the device. */
int cbs_episode_stats(cbs_handle* h, double* out_host);

/* device pointer of a state field (for zero-copy consumers such as an NCCL all-reduce of CBS_F_STAT_ACCUM) */
void* cbs_state_ptr(cbs_handle* h, int32_t field);
int cbs_reset_stat_accum(cbs_handle* h, uintptr_t stream);
/* debugging aid: per-env {cycles, rows scanned, live pairs, float64 re-scores, pair combinations, start clock} of the
 * following cbs_decode calls are written to trace_dev ([num_envs][6] int64, device); NULL switches it off */
int cbs_debug_select_trace(cbs_handle* h, long long* trace_dev);
/* debug aid: per-env {start ns (globaltimer), duration ns, flags at entry, nodes << 16 | edges} of the env's last item in
 * the observe kernel, followed by five phase end times; trace_dev = int64[num_envs][9] device buffer, or NULL to switch tracing off */
int cbs_debug_observe_trace(cbs_handle* h, long long* trace_dev);
/* number of kernels this library launched since creation (bench.py reports it) */
int64_t cbs_launch_count(const cbs_handle* h);
/* cudaDeviceSynchronize + device error flag (capacity overflow, empty action table) -> CBS_ERR_CAPACITY */
int cbs_sync(cbs_handle* h);
/* out3 = { sizeof(cbs_config), sizeof(cbs_scenario_tables), sizeof(cbs_gae_tables) } — lets a foreign-function
 * binding verify its struct layout without a GPU */
int cbs_struct_sizes(int32_t* out3);
/* bytes of device memory held by the handle (tables + env state) */
int64_t cbs_state_bytes(const cbs_handle* h);
/* out8 = { node capacity, snapshot slots, edge capacity, 1 if the tcgen05 decode GEMM is active, vt_stride, observation length,
 *          mask_pitch (uint32 words per env in CBS_F_MASKS), scalar_pitch (int32 words per env in CBS_F_SCALARS) } */
int cbs_capacities(const cbs_handle* h, int32_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* CBSIM_H */
