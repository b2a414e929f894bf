// cbs_types.h — device-side layout shared by all kernels of libcbsim.
// Numbers mirror c-cyberbattlesim_b200/constants.py (tests/test_constants.py keeps them in sync).
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cbs {

constexpr int NODE_EMB = 64;
constexpr int VULN_EMB = 768;
constexpr int OUTCOME_DIM = 9;
constexpr int ACTION_DIM = 905;
constexpr int OBS_GRAPH = 192;
constexpr int OBS_DIM = 194;
constexpr int NN_CH = 16;
constexpr int PROJ_ROWS = 18;  // 16 hidden + bias slot + root term
constexpr int NUM_DYN = 6;     // persistence, collected, exfiltrated, evasion, privilege, status
constexpr int MAX_NODES = 128;
#ifndef CBS_OBS_SMEM_NODES
#define CBS_OBS_SMEM_NODES 32   // visible graphs up to this many nodes keep their embeddings in shared memory (k_observe.cu)
#endif
constexpr int SCAL_PITCH = 32;  // int32 words per env in State::scal: four sectors of 8 words
// reset cache: what the first encode of an episode produces depends on (scenario, starter) only — one entry per scenario node
constexpr int RC_Z = 264;       // floats [0, obs_dim) observation, [RC_Z, RC_Z + 64) the starter's embedding, [RC_N2] its squared norm
constexpr int RC_N2 = RC_Z + NODE_EMB;
constexpr int RC_PITCH = 336;
constexpr int OBS_CLASSES = 16; // observe work classes, heaviest first (transition.cuh): cost buckets of 2.5 us.  The observe kernel's duration
                                // is set by how well its ~2 heavy items per warp pack, so they are claimed longest first
constexpr int SUB_CLASSES = 10;   // outcome kinds that can have action-table rows (K_DOS .. K_LATERAL; Execution never does)
constexpr int SUB_META = 16;      // int32 per env in State::sub_meta: [0, 10) rows alive per class, [10] classes ranked so far, [11], [12] the
                                  // classes' first-appearance rank (4 bits per kind, 15 = not in the table yet), [13] lifetime count of balance calls
constexpr int SUB_MAX_ROWS_PER_PAIR = 256;   // candidate rows of one (source, target) pair addressable by a list entry
constexpr int SCHED_BINS = 8;   // decode cost bins: rows < 64, < 128, ..., >= 4096 (longest-first scheduling)

// outcome kinds (simulation/model.py:66-193)
enum Kind : int { K_DOS = 0, K_DISCOVERY, K_COLLECTION, K_EXFILTRATION, K_RECON, K_EVASION, K_PERSISTENCE,
                  K_PRIVESC, K_CREDACCESS, K_LATERAL, K_EXECUTION, N_KINDS };
// obtained-outcome codes for failures (simulation/attacker_actions.py)
enum Code : int { OC_INVALID_SRC_NOT_OWNED = 16, OC_INVALID_TGT_NOT_DISCOVERED, OC_SRC_NOT_RUNNING, OC_TGT_NOT_RUNNING,
                  OC_NO_VULNERABILITY, OC_NO_PRIVILEGE, OC_OUTCOME_NOT_PRESENT, OC_PORT_NOT_LISTENING, OC_FW_OUTGOING,
                  OC_FW_INCOMING, OC_UNSUCCESSFUL, OC_NO_NEEDED, OC_REPEATED, OC_REMOTE_OUTCOME_LOCAL };
enum Metric : int { METRIC_COSINE = 0, METRIC_L1 = 1, METRIC_L2 = 2, METRIC_INF = 3 };   // cbs_config.distance_metric (compressed:571-576)
enum Goal : int { GOAL_CONTROL = 0, GOAL_DISCOVERY = 1, GOAL_DISRUPTION = 2, GOAL_CONTROL_NODE = 3, GOAL_DISCOVERY_NODE = 4,
                  GOAL_DISRUPTION_NODE = 5 };
enum Mask : int { M_OWNED = 0, M_DISCOVERED, M_VISIBLE, M_HAS_DATA, M_COLLECTED, M_EXFILTRATED, M_PERSISTENCE,
                  M_EVASION, M_STOPPED, M_PRIV_USER, M_PRIV_ROOT,
                  // planes that only change under the re-imaging static defender (_env/static_defender.py:27-60)
                  M_IMAGING,      // status == Imaging (simulation/static_defender_actions.py:37-52)
                  M_X_IMAGING,    // the node's cached feature vector in the visible graph says Imaging (compressed:472-479)
                  M_EVER_OWNED,   // last_owned_at is not None (attacker_actions.py:87)
                  M_OWN_STALE,    // re-imaged after it was last owned (attacker_actions.py:561-573)
                  N_MASKS };
constexpr int REIMAGING_DURATION = 15;   // static_defender_actions.py:19
constexpr int MAX_SCAN_CAPACITY = 8;
enum Reward : int { R_VALUE = 0, R_COST, R_NODE_DISCOVERED, R_COLLECTED, R_EXFILTRATED, R_PERSISTENCE, R_PRIVESC,
                    R_VISIBILITY, R_DOS, R_EVASION, N_REWARDS };
enum Penalty : int { P_NO_VULN = 0, P_NO_PRIV, P_SUCCESS_FAILED, P_NO_DATA_COLLECT, P_NO_DATA_EXFIL, P_ALREADY_PERSISTENT,
                     P_ALREADY_STOPPED, P_ALREADY_OWNED, P_ALREADY_VISIBLE, P_ALREADY_EVASION, P_UNOPEN_PORT,
                     P_PRIVESC_NOT_OWNED, P_PRIVESC_ALREADY, P_OUTCOME_NOT_VALID, P_FW_LOCAL, P_FW_REMOTE,
                     P_INVALID_ACTION, P_DISTANCE, N_PENALTIES };
// per-env int32 scalars, grouped into four 32-byte sectors per env; each sector is its own dense array
// (State::scal[sector][B][8]) so that a kernel streams only the sectors it needs.  The transition reads and rewrites
// sector 0 on every step, touches sector 1 only for outcomes that change a list, and never reads sectors 2-3 (episode
// constants come from the tables via S_SCST).
enum Scalar : int {
  // sector 0 — rewritten by every transition
  S_FLAGS = 0, S_STEPCOUNT, S_NUM_ITER, S_TOTAL_STEPS, S_OUTCOME,
  S_SCST,           // scenario << 8 | starter node (copy of S_SCENARIO / S_STARTER, written by the reset)
  S_EP_RETURN,      // float64 episode return, two words (8-byte aligned)
  S_EP_RETURN_HI,
  // sector 1 — list lengths and counters
  S_N_DISC = 8, S_N_OWNED, S_DISC_AMOUNT,
  S_N_OWNED_RAW,    // len(env.owned_nodes) under a defender (the list can hold duplicates, cyberbattle_env.py:425-430)
  S_N_REIMAGED,     // len(overall_reimaged) == num_events of the episode (cyberbattle_env.py:419-422)
  S_N_SLOTS, S_N_EDGES, S_N_ENCODES,
  // sector 2 — constants of the episode (functions of scenario and starter)
  S_SCENARIO = 16, S_STARTER, S_NODE_OFF, S_OWNABLE, S_DISCOVERABLE, S_DISRUPTABLE, S_PROP_NODES, S_DISCOVERABLE_AMOUNT,
  // sector 3
  S_EPISODES = 24,
  S_UVULN_OFF,      // T.sc_uvuln_off[scenario], kept per env by the reset (one dependent table read less per edge update)
  N_SCALARS };
// S_FLAGS bits
constexpr int FL_DONE = 1, FL_TRUNC = 2, FL_REASON_SHIFT = 2 /*2 bits*/, FL_ADD_EDGE = 16, FL_REENCODE = 32,
              FL_NEEDS_RESET = 64, FL_FINISHED_THIS_STEP = 128,
              FL_DIRTY = 256,  // node features / edges / node sets changed since the last encode
              FL_INTEREST_IN_GRAPH = 512,  // *_node goals: the interest node was added to the visible graph (compressed:254-256)
              FL_PENDING_SHIFT = 16;       // bits 16..31, sample_subset_samples: table builds skipped since the last one that ran (the
                                           // encode of an unchanged graph is skipped, but the reference's balance step still draws)
// vi_flags bits (scenario.py VI_*)
constexpr uint32_t VI_LISTENING = 1u, VI_IN_ALLOWED = 2u;
constexpr int VI_PRIVREQ_SHIFT = 2, VI_LEVEL_ANY_SHIFT = 4, VI_LEVEL_REMOTE_SHIFT = 6;
// stat accumulators (sums over finished episodes on this GPU)
enum Accum : int { A_EPISODES = 0, A_RETURN, A_LENGTH, A_WINS, A_LOST, A_CUTOFF, A_STAT0 /* .. A_STAT0+13 */, N_ACCUM = 20 };

struct Tables {  // immutable, device pointers
  int num_scenarios, max_nodes, words, num_global_vulns;
  const int32_t *sc_num_nodes, *sc_node_off, *sc_port_off, *sc_uvuln_off, *sc_num_uvuln, *sc_discoverable_amount,
      *sc_feasible_off, *feasible_starters, *sc_interest;
  const int64_t* sc_instof_off;
  const uint32_t *sc_init_has_data, *sc_init_visible;
  const int32_t *nd_value, *nd_ownable, *nd_discoverable, *nd_disruptable, *nd_row_off;
  const uint8_t* nd_level_at_access;
  const uint8_t* nd_reimageable;
  const uint32_t* outblock;
  const int32_t *inst_of, *vi_port, *vi_recon_any, *vi_recon_remote, *vi_ulocal, *row_inst, *uvuln_global, *row_ulocal;
  const uint32_t *vi_flags, *row_packed;
  const uint16_t *vi_kinds_any, *vi_kinds_remote;
  const double *vi_success, *vi_cost, *vemb64, *vnorm2;
  const uint8_t* recon_nodes;
  const float* vemb32;
  // derived at load time for the transition (cbs_load_scenarios): one 32-byte record per scenario and per vulnerability
  // instance, so that each look-up level of the transition is two 128-bit loads from one sector instead of 4-7 gathers
  const int4* sc_pack;         // [S][2]  { num_nodes, node_off, num_uvuln, port_off } { instof_off lo, hi, interest node or -1, 0 }
  const uint4* vi_pack;        // [I][2]  { vi_flags | len_any << 8 | len_remote << 16, kinds_any | kinds_remote << 16,
                               //           port (words > 1) or the port's outgoing-firewall node mask itself (words == 1), recon offset }
                               //         { success rate (float64), cost (float64) }
  const uint8_t* recon_pack;   // per instance: the "any type" Reconnaissance node list, then the "REMOTE only" one, each padded to 8 bytes
  const uint2* recon_mask;     // [I] node sets of those two lists as bitmasks (scenarios of <= 32 nodes; zeros otherwise)
  // ExternalRandomEvents defender
  const uint16_t* nd_ev_init;  // [Nn][4] running | incoming BLOCK | outgoing BLOCK bits per service slot, service count
  const uint8_t* out_slot;     // [ports][max_nodes] node's service slot of a scenario port (0xFF none)
  const float* ev_proj;        // [30][18][64] folded firewall / running feature columns
  // GAE
  const float *node_static, *dyn_proj, *vuln_h, *nn0_b, *bn1_scale, *bn1_shift, *gcn_wt, *bn2_scale, *bn2_shift;
};

struct Params {  // configuration, by value
  int B, ncap, words, slots, ecap;
  int obs_dim;      // 194, or 258 for the *_node goals (64 extra floats: the interest node's embedding)
  long long global_env_offset;
  unsigned long long seed;
  int goal, episode_iterations, absolute_reward, stop_at_goal, remove_main, remove_all, switch_interval, auto_reset;
  double prop_coeff, winning_reward, losing_reward;
  double rew[N_REWARDS], pen[N_PENALTIES];
  float margin;
  int qlen;
  int act_stride;   // row pitch (floats) of the action tensors handed to decode; 905 = dense
  // static defender ScanAndReimageCompromisedMachines (_env/static_defender.py:27-60); defender == 0: none
  int defender, scan_capacity, scan_frequency;
  double detect_prob;
  int always_encode;   // defender or precise_graph_encoding: every step re-encodes (compressed:401,455-462)
  int ocap;            // capacity of owned_raw
  int mpitch;          // uint32 words per env in State::masks (N_MASKS * words rounded up to 16)
  int precise_graph;       // precise_graph_encoding (compressed:455-462)
  int precise_positions;   // precise_action_space_positions (compressed:419-427,498-506): table rows are refreshed, see build_table
  int metric;              // enum Metric; != METRIC_COSINE: k_decode_metric.cu decodes, the transition runs as its own launch
  double event_prob;       // defender == 2: random_event_probability
  int subset_k;            // sample_subset_samples (compressed:521-522,553-567): rows kept per outcome class, 0 = the whole table
};

struct State {  // mutable, device pointers
  // Per-env records: every kernel on the step path handles an env with one warp (or one lane / thread), so the env's few
  // dozen words must share 32-byte sectors; plane-major SoA put each of them in a different sector.
  uint32_t* masks;       // [B][mpitch]  plane p, word w at p * words + w
  int32_t* scal;         // [4][B][8]  sector-major (enum Scalar: sector = plane >> 3)
  uint8_t* disc_order;   // [B][ncap]
  uint8_t* owned_order;  // [B][ncap]
  uint8_t* pair_slot;    // [B][ncap*ncap]   0xFF = pair not in the action table
  // defender only (1-byte dummies otherwise):
  uint8_t* owned_raw;    // [B][ocap]  env.owned_nodes exactly as the reference keeps it (removals, duplicates); owned_order
                         //            stays the append-only list of every node that ever was a source of table rows
  uint8_t* reimage_left; // [B][ncap]  node_reimaging_progress (static_defender_actions.py:23)
  uint8_t* pair_opos;    // [B][ncap*ncap]  position of the source in owned_raw when the pair entered the table (tie order)
  // precise_action_space_positions only (1-byte dummy otherwise):
  uint32_t* changed;     // [B][words]  defender + precise_action_space_positions: nodes the defender re-imaged in the last step
                         //             (env.changed_nodes, cyberbattle_env.py:418: the refresh seeds of compressed:423-427)
  uint8_t* pair_epoch;   // [B][ncap*ncap]  slot at which the pair FIRST entered the table: its place in the insertion order;
                         //                 pair_slot then names the snapshot its rows currently carry (refreshed over time)
  // events defender only: per node { running services, incoming BLOCK, outgoing BLOCK, unused } bit sets over the node's
  // service slots — the live ones and the ones the node's CACHED feature vector in the visible graph was built from
  uint16_t* ev_cur;      // [B][ncap][4]
  uint16_t* ev_x;        // [B][ncap][4]
  const int32_t* def_nodes;     // [B][scan_capacity] test override of the scan draws (random.choices), or nullptr -> Philox
  const float* def_uniforms;    // [B][scan_capacity] test override of the detection uniforms, consumed in call order
                                //   (events defender: [B][ncap][4] = function index, event / pick / side uniforms per node)
  float* z_hist;         // [B][slots][ncap][64]  node embeddings of the encode that created the slot
  float* zn2_hist;       // [B][slots][ncap]      their squared norms
  __half* z16_hist;      // [B][slots][ncap][64]  half-precision copy read by the approximate decode scan
  uint8_t* edge_src;     // [B][ecap]
  uint8_t* edge_dst;     // [B][ecap]
  int32_t* edge_cnt;     // [B][ecap]   live accumulator length (0 after the compressed:237 reset)
  float* edge_sum;       // [B][ecap][16]  running sum of W1*emb over the accumulator
  float* edge_m;         // [B][ecap][16]  W1 * (stored edge attribute)
  float* obs;            // [B][OBS_DIM] cached observation
  float* term_obs;       // [B][OBS_DIM]
  int32_t* sel;          // [B][4]
  double* dist;          // [B]
  double* reward64;      // [B]
  double* last_stats;    // [B][14]
  double* accum;         // [N_ACCUM]
  int32_t* starter_queue;  // [B][qlen] or nullptr
  float* vt;             // [B][Ug]  action x vulnerability-embedding products (decode GEMM output)
  float* scratch;        // [B][2][ncap][64] encode scratch when ncap > 32
  float* reset_cache;          // [total scenario nodes][RC_PITCH], filled by the first reset from each (scenario, starter)
  int32_t* reset_cache_flag;   // [total scenario nodes] 1 = entry valid
  int32_t* errflag;      // [4]  [0] capacity / domain error code (cbs_sync), [1] steps at which the reference itself would have raised
                         //      (CBS_F_DIVERGENCE), [2] decodes whose winner sat in the outer half of the re-score margin (CBS_F_MARGIN_EDGE)
  int32_t* worklist;     // [OBS_CLASSES][B] envs whose step needs graph work, by cost class
  int32_t* work_ctr;     // [1] finished-warp counter, [2] next item (dynamic scheduling), [4 .. 4 + OBS_CLASSES) class list lengths
  int32_t* work_est;     // [B] candidate rows in the env's action table (decode cost estimate)
  int32_t* bin_cnt;      // [2][SCHED_BINS + 1] envs per cost bin (double buffered; filled by the transition for the next decode), [SCHED_BINS] = finished-warp counter
  int32_t* bin_list;     // [2][SCHED_BINS][B] env ids per bin
  // sample_subset_samples only (1-element dummies otherwise): the action table is then EXPLICIT, at most subset_k rows per outcome class
  uint32_t* sub_rows;    // [B][SUB_CLASSES][subset_k]  source | target << 7 | row within the pair << 14 | insertion epoch << 22
  int32_t* sub_meta;     // [B][SUB_META]
  uint32_t* sub_alive;   // [B][ncap*ncap][8]  with precise_action_space_positions: which rows of a pair are in the lists
  uint16_t* sub_newp;    // [B][ncap*ncap]     scratch (graphs beyond the shared-memory buffers): the pairs one table build touches
};


// ---- programmatic dependent launch (the step's three kernels overlap their heads with the previous kernel's tail) ----
// A kernel launched through launch_pdl may START before the previous kernel on the stream has finished (once every CTA of that
// kernel has called pdl_trigger or exited); it must not read or write anything an earlier kernel touches before pdl_wait(),
// which returns when the previous kernel has completed and its writes are visible.  Launched with a plain <<<>>> (or with
// CBS_NO_PDL=1) both calls are no-ops and the stream order is the usual one.  Works under stream capture (programmatic edges).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per DEVICE: remember what was granted per (device, kernel), not per process
inline cudaError_t ensure_dyn_smem(const void* func, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = granted[std::make_pair(dev, func)];
  if (smem <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) cur = smem;
  return e;
}

inline bool pdl_enabled() {
  static const bool on = getenv("CBS_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool allow, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (allow && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

}  // namespace cbs
