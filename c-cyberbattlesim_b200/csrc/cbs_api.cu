// cbs_api.cu — C ABI (include/cbsim.h): handle lifetime, table upload, kernel launches.
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cbsim.h"
#include <cuda_fp16.h>

#include "cbs_types.h"

namespace cbs {
cudaError_t launch_decode_gemm_simt(const float*, int, const float*, float*, int, int, int, cudaStream_t);
cudaError_t launch_decode_gemm_tc(const float*, int, const float*, float*, float*, int, int, int, int32_t*, int, cudaStream_t);
cudaError_t launch_decode_gemm_f16(const float*, int, const __half*, float*, int, int, int, int32_t*, int, cudaStream_t);
cudaError_t convert_vemb_f16(const float*, __half*, size_t, cudaStream_t);
bool decode_gemm_f16_applies(int, int);
bool decode_gemm_tc_available();
cudaError_t launch_decode_select(const Tables&, const Params&, const State&, const float*, int, int, int, const float*, float*, uint8_t*,
                                 int32_t*, double*, cudaStream_t);
cudaError_t launch_decode_metric(const Tables&, const Params&, const State&, const float*, double*, int, int, int32_t*, double*, cudaStream_t);
cudaError_t launch_observe(const Tables&, const Params&, const State&, const uint8_t*, int, int, cudaStream_t);
cudaError_t launch_transition(const Tables&, const Params&, const State&, const int32_t*, const double*, const float*, int, float*,
                              uint8_t*, uint8_t*, uint8_t*, int, cudaStream_t);
cudaError_t launch_transition_ksteps(const Tables&, const Params&, const State&, const int32_t*, const double*, const float*, int, float*,
                                     uint8_t*, cudaStream_t);

__global__ void info_kernel(Params P, State S, int32_t* __restrict__ info) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int4 sl = reinterpret_cast<const int4*>(S.sel)[b];
  const int32_t* sc = S.scal + (size_t)b * 8;   // sector 0
  const int flags = sc[S_FLAGS];
  int32_t* o = info + (size_t)b * CBS_INFO_INTS;
  o[0] = sl.x; o[1] = sl.y; o[2] = sl.z; o[3] = sl.w;
  o[4] = sc[S_OUTCOME];
  o[5] = (flags >> FL_REASON_SHIFT) & 3;
  o[6] = sc[S_STEPCOUNT];
  o[7] = ((flags & FL_TRUNC) ? 1 : 0) | ((sc[S_SCST] >> 8) << 8);   // bit 0 truncated, bits 8.. the scenario in force during the step
}

// cbs_replay: one warp per logged env copies the env's records into row (t, e) of the log.  phase 0 = after the transition,
// phase 1 = after the observe.
__global__ void replay_log_kernel(Params P, State S, cbs_replay_log L, int t, int phase, int owned_len) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= L.num_logged) return;
  const int b = L.first_env + e;
  const size_t row = (size_t)t * L.num_logged + e;
  auto sc = [&](int plane) { return S.scal[((size_t)(plane >> 3) * P.B + b) * 8 + (plane & 7)]; };
  if (phase == 0) {
    const int flags = sc(S_FLAGS);
    if (lane < 4) {
      if (L.sel) L.sel[row * 4 + lane] = S.sel[(size_t)b * 4 + lane];
      if (L.meta) L.meta[row * 4 + lane] = lane == 0 ? sc(S_OUTCOME) : (lane == 1 ? ((flags & 15) | ((sc(S_SCST) >> 8) << 8)) : (lane == 2 ? sc(S_STEPCOUNT) : sc(S_EPISODES)));
    }
    if (lane == 0) {
      if (L.reward) L.reward[row] = S.reward64[b];
      if (L.dist) L.dist[row] = S.dist[b];
    }
    if (L.masks) for (int i = lane; i < P.mpitch; i += 32) L.masks[row * P.mpitch + i] = S.masks[(size_t)b * P.mpitch + i];
    if (L.disc_order) for (int i = lane; i < P.ncap; i += 32) L.disc_order[row * P.ncap + i] = S.disc_order[(size_t)b * P.ncap + i];
    if (L.owned_order)
      for (int i = lane; i < owned_len; i += 32)
        L.owned_order[row * owned_len + i] = P.defender ? S.owned_raw[(size_t)b * P.ocap + i] : S.owned_order[(size_t)b * P.ncap + i];
    if (L.counters && lane < 8) {
      const int planes[8] = {S_STEPCOUNT, S_NUM_ITER, S_DISC_AMOUNT, S_OWNABLE, S_DISCOVERABLE, S_DISRUPTABLE, S_DISCOVERABLE_AMOUNT, S_N_DISC};
      int v = sc(planes[lane]);
      if (lane == 7) v |= sc(P.defender ? S_N_OWNED_RAW : S_N_OWNED) << 16;
      L.counters[row * 8 + lane] = v;
    }
  } else {
    // the step's own done flag was logged in phase 0 (an in-place reset has cleared it by now)
    const bool finished = L.meta ? (L.meta[row * 4 + 1] & 3) != 0 : false;
    for (int i = lane; i < P.obs_dim; i += 32) {
      const float cur = S.obs[(size_t)b * P.obs_dim + i];
      if (L.obs) L.obs[row * P.obs_dim + i] = finished ? S.term_obs[(size_t)b * P.obs_dim + i] : cur;
      if (L.reset_obs) L.reset_obs[row * P.obs_dim + i] = finished ? cur : 0.f;
    }
    if (finished) {
      if (L.reset_masks) for (int i = lane; i < P.mpitch; i += 32) L.reset_masks[row * P.mpitch + i] = S.masks[(size_t)b * P.mpitch + i];
      if (L.stats && lane < 14) L.stats[row * 14 + lane] = S.last_stats[(size_t)b * 14 + lane];
    }
  }
}

__global__ void replay_force_kernel(Params P, State S, const int32_t* __restrict__ force_sel, const double* __restrict__ force_dist) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int4 f = reinterpret_cast<const int4*>(force_sel)[b];
  if (f.x < 0) return;
  reinterpret_cast<int4*>(S.sel)[b] = f;
  S.dist[b] = force_dist[b];
}

__global__ void init_flags_kernel(int32_t* scal, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) scal[(size_t)b * 8 + S_FLAGS] = FL_NEEDS_RESET;   // sector 0
}
}  // namespace cbs

using namespace cbs;
namespace cbs { extern long long* g_sel_trace; extern long long* g_obs_trace; }

static thread_local std::string g_create_error;
static const char* const kErrflagMessage =
    "device reported capacity/domain error %d (1 snapshot slots, 2 edges, 3 empty action table, 4 worklist, 5 owned-node list, "
    "6 sub-sampled action table, 8 node embedding beyond the half-precision range, 7 removal of an absent owned node (the reference raises ValueError there), 9 tensor-core pipeline timeout)";

struct cbs_handle {
  cbs_config cfg{};
  Params P{};
  Tables T{};
  State S{};
  std::vector<void*> table_allocs, state_allocs;
  std::string err;
  bool loaded = false;
  int Ug = 0, vt_stride = 0;
  int64_t launches = 0;
  bool use_tc = false;
  bool actions_prestaged = getenv("CBS_ACTIONS_PRESTAGED") != nullptr;   // cbs_set_actions_prestaged
  int sched_buf = 0;     // cost-bin buffer the next decode reads (the transitions of that step fill the other one)
  int num_sms = 148;
  float* a_packed = nullptr;   // [B][768] 16-byte aligned copy of the vulnerability part of the action (TMA source)
  __half* vemb16 = nullptr;    // [Ug][768] the vulnerability embeddings in half precision (B operand of the FP16 contraction)
  double* vt64 = nullptr;      // [B][Ug] vulnerability part of the l1 / l2 / inf distances (distance_metric != cosine only)
  // host-step staging
  cudaStream_t hstream = nullptr;
  float *h_actions = nullptr, *h_uniforms = nullptr, *h_reward = nullptr;
  uint8_t* h_done = nullptr;
  int32_t* h_info = nullptr;
  int32_t* h_errflag = nullptr;   // pinned host copy of State::errflag, refreshed by every cbs_step_host_async
  // io scratch for cbs_step
  int32_t* d_sel = nullptr;
  double* d_dist = nullptr;
  size_t state_bytes = 0;
};

static int fail(cbs_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CK(h, expr)                                                                                     \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return fail(h, CBS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

template <typename Tp>
static int dalloc(cbs_handle* h, std::vector<void*>& pool, Tp** out, size_t count, bool zero = true) {
  void* p = nullptr;
  const size_t bytes = (count ? count : 1) * sizeof(Tp);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(h, CBS_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  if (zero) cudaMemset(p, 0, bytes);
  pool.push_back(p);
  h->state_bytes += bytes;
  *out = static_cast<Tp*>(p);
  return 0;
}

template <typename Tp>
static int upload(cbs_handle* h, const Tp** out, const Tp* src, size_t count) {
  Tp* p = nullptr;
  int rc = dalloc(h, h->table_allocs, &p, count, false);
  if (rc) return rc;
  if (count) {
    if (!src) return fail(h, CBS_ERR_INVALID_ARG, "null table pointer");
    cudaError_t e = cudaMemcpy(p, src, count * sizeof(Tp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(h, CBS_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return 0;
}

// vi_port is scenario-local: the first port of the scenario each vulnerability instance belongs to, found through
// inst_of (inst_of[sc_instof_off[s] .. sc_instof_off[s+1]) lists the instances of scenario s, or -1)
static std::vector<int32_t> instance_port_offsets(const cbs_scenario_tables* t) {
  // inst_of[sc_instof_off[s] .. sc_instof_off[s+1]) holds the instances of scenario s (or -1)
  std::vector<int32_t> off((size_t)(t->num_inst > 0 ? t->num_inst : 1), 0);
  for (int s = 0; s < t->num_scenarios; ++s)
    for (int64_t k = t->sc_instof_off[s]; k < t->sc_instof_off[s + 1]; ++k)
      if (t->inst_of[k] >= 0) off[t->inst_of[k]] = t->sc_port_off[s];
  return off;
}


// Per-env capacities an episode of the given cut-offs can need (snapshot slots = table-growing encodes, visible-graph edges)
static void derive_capacities(int max_nodes, int episode_iterations, double prop_coeff, bool defender, bool precise, int* slots_out,
                              int* ecap_out, int* max_steps_out) {
  int max_steps = episode_iterations;
  if (prop_coeff > 0) {
    const double lim = (double)(max_nodes - 1) * prop_coeff;
    const int l = (int)lim + ((double)(int)lim < lim ? 1 : 0);
    if (l < max_steps) max_steps = l;
  }
  max_steps += 1;   // the cut-offs test the pre-increment counter (cyberbattle_env.py:361-366, :394)
  // table-growing encodes: each adds an owned or a discovered node ... or, under a defender, sees a node come back from
  // re-imaging with pairs still missing
  int slots = (defender ? 4 : 2) * max_nodes - 1;
  if (max_steps + 1 < slots) slots = max_steps + 1;
  if (precise) slots = max_steps + 1;   // every table-maintaining encode may refresh rows: one snapshot per step
  int ecap = max_nodes * max_nodes;
  if (max_steps < ecap) ecap = max_steps;
  *slots_out = slots; *ecap_out = ecap; *max_steps_out = max_steps;
}

extern "C" {

int cbs_abi_version(void) { return CBS_ABI_VERSION; }

const char* cbs_last_error(const cbs_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cbs_create(const cbs_config* cfg, cbs_handle** out) {
  if (!cfg || !out) return fail(nullptr, CBS_ERR_INVALID_ARG, "cbs_create: null argument");
  if (cfg->abi_version != CBS_ABI_VERSION) return fail(nullptr, CBS_ERR_INVALID_ARG, "ABI version mismatch (%d vs %d)", cfg->abi_version, CBS_ABI_VERSION);
  if (cfg->num_envs <= 0) return fail(nullptr, CBS_ERR_INVALID_ARG, "num_envs must be positive");
  if (cfg->goal < 0 || cfg->goal > 5) return fail(nullptr, CBS_ERR_INVALID_ARG, "unsupported goal %d", cfg->goal);
  if (cfg->static_defender < 0 || cfg->static_defender > 2)
    return fail(nullptr, CBS_ERR_INVALID_ARG, "static_defender must be 0 (none), 1 (scan and re-image) or 2 (external random events)");
  if (cfg->static_defender == 2 && !(cfg->random_event_probability >= 0.0 && cfg->random_event_probability <= 1.0))
    return fail(nullptr, CBS_ERR_INVALID_ARG, "random_event_probability must be in [0, 1]");
  if (cfg->static_defender == 1 && (cfg->scan_capacity < 1 || cfg->scan_capacity > MAX_SCAN_CAPACITY || cfg->scan_frequency < 1))
    return fail(nullptr, CBS_ERR_INVALID_ARG, "scan_capacity must be in 1..%d and scan_frequency >= 1", MAX_SCAN_CAPACITY);
  if (cfg->precise_action_space_positions && cfg->static_defender == 2)   // the reference raises networkx.NodeNotFound there (compressed:423-427,498-500)
    return fail(nullptr, CBS_ERR_INVALID_ARG, "precise_action_space_positions cannot be used with the events defender");
  if (cfg->distance_metric < METRIC_COSINE || cfg->distance_metric > METRIC_INF)
    return fail(nullptr, CBS_ERR_INVALID_ARG, "Unsupported metric %d. Use 0 cosine, 1 l1, 2 l2 or 3 inf", cfg->distance_metric);   // compressed:578-579
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, CBS_ERR_NO_DEVICE, "no CUDA device available (libcbsim has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CBS_ERR_INVALID_ARG, "device %d out of range", cfg->device);
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return fail(nullptr, CBS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cbs_handle* h = new cbs_handle();
  h->cfg = *cfg;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  Params& P = h->P;
  P.B = cfg->num_envs;
  P.global_env_offset = cfg->global_env_offset;
  P.seed = cfg->seed;
  P.goal = cfg->goal;
  P.obs_dim = cfg->goal >= GOAL_CONTROL_NODE ? OBS_DIM + NODE_EMB : OBS_DIM;
  P.episode_iterations = cfg->episode_iterations;
  P.prop_coeff = cfg->proportional_cutoff_coefficient;
  P.winning_reward = cfg->winning_reward;
  P.losing_reward = cfg->losing_reward;
  P.absolute_reward = cfg->absolute_reward;
  P.stop_at_goal = cfg->stop_at_goal_reached;
  P.remove_main = cfg->remove_main_obstacles;
  P.remove_all = cfg->remove_all_obstacles;
  P.switch_interval = cfg->switch_interval;
  P.auto_reset = cfg->auto_reset;
  for (int i = 0; i < N_REWARDS; ++i) P.rew[i] = cfg->rewards[i];
  for (int i = 0; i < N_PENALTIES; ++i) P.pen[i] = cfg->penalties[i];
  P.qlen = 0;
  P.act_stride = ACTION_DIM;
  P.defender = cfg->static_defender;
  P.scan_capacity = cfg->scan_capacity;
  P.scan_frequency = cfg->scan_frequency;
  P.detect_prob = cfg->detect_probability;
  P.event_prob = cfg->random_event_probability;
  P.always_encode = (cfg->static_defender || cfg->precise_graph_encoding) ? 1 : 0;
  P.precise_positions = cfg->precise_action_space_positions ? 1 : 0;
  P.precise_graph = cfg->precise_graph_encoding ? 1 : 0;
  P.metric = cfg->distance_metric;
  P.subset_k = cfg->sample_subset_samples > 0 ? cfg->sample_subset_samples : 0;
  *out = h;
  return CBS_OK;
}

void cbs_destroy(cbs_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : h->table_allocs) cudaFree(p);
  for (void* p : h->state_allocs) cudaFree(p);
  if (h->hstream) cudaStreamDestroy(h->hstream);
  if (h->h_errflag) cudaFreeHost(h->h_errflag);
  delete h;
}

int cbs_load_scenarios(cbs_handle* h, const cbs_scenario_tables* t, const cbs_gae_tables* g) {
  if (!h || !t || !g) return fail(h, CBS_ERR_INVALID_ARG, "cbs_load_scenarios: null argument");
  if (h->loaded) return fail(h, CBS_ERR_INVALID_ARG, "scenarios already loaded (create a new handle)");
  if (t->max_nodes < 1 || t->max_nodes > CBS_MAX_NODES) return fail(h, CBS_ERR_INVALID_ARG, "max_nodes %d out of range", t->max_nodes);
  if (t->words != (t->max_nodes + 31) / 32) return fail(h, CBS_ERR_INVALID_ARG, "words does not match max_nodes");
  if (h->P.subset_k) {   // limits of the explicit table's 32-bit row entries and of the Philox row identity (subset.cuh)
    for (int n = 0; n < t->num_nodes_total; ++n)
      if (t->nd_row_off[2 * n + 2] - t->nd_row_off[2 * n] > SUB_MAX_ROWS_PER_PAIR)
        return fail(h, CBS_ERR_INVALID_ARG, "sample_subset_samples: node %d has more than %d candidate rows", n, SUB_MAX_ROWS_PER_PAIR);
    for (int s = 0; s < t->num_scenarios; ++s)
      if (t->sc_num_uvuln[s] > 4096) return fail(h, CBS_ERR_INVALID_ARG, "sample_subset_samples: scenario %d has more than 4096 vulnerabilities", s);
  }
  CK(h, cudaSetDevice(h->cfg.device));
  Tables& T = h->T;
  Params& P = h->P;
  const int S_ = t->num_scenarios, Nn = t->num_nodes_total, I = t->num_inst;
  T.num_scenarios = S_; T.max_nodes = t->max_nodes; T.words = t->words; T.num_global_vulns = t->num_global_vulns;
  int rc = 0;
#define UP(field, count) if ((rc = upload(h, &T.field, t->field, (size_t)(count)))) return rc
  UP(sc_num_nodes, S_); UP(sc_node_off, S_ + 1); UP(sc_port_off, S_ + 1); UP(sc_uvuln_off, S_ + 1); UP(sc_num_uvuln, S_);
  UP(sc_instof_off, S_ + 1); UP(sc_discoverable_amount, S_);
  UP(sc_init_has_data, (size_t)S_ * t->words); UP(sc_init_visible, (size_t)S_ * t->words);
  UP(sc_feasible_off, S_ + 1); UP(feasible_starters, t->num_feasible);
  if (h->cfg.goal >= GOAL_CONTROL_NODE) {
    if (!t->sc_interest) return fail(h, CBS_ERR_INVALID_ARG, "*_node goals need sc_interest (one interest node per scenario)");
    for (int s = 0; s < S_; ++s)
      if (t->sc_interest[s] < 0 || t->sc_interest[s] >= t->sc_num_nodes[s]) return fail(h, CBS_ERR_INVALID_ARG, "interest node of scenario %d out of range", s);
    UP(sc_interest, S_);
  }
  UP(nd_value, Nn); UP(nd_level_at_access, Nn);
  if (h->cfg.static_defender) {
    if (!t->nd_reimageable) return fail(h, CBS_ERR_INVALID_ARG, "the static defender needs nd_reimageable");
    UP(nd_reimageable, Nn);
  }
  if (h->cfg.static_defender == 2) {
    if (!t->nd_ev_init || !t->vi_svc_slot || !t->out_slot || !g->ev_proj)
      return fail(h, CBS_ERR_INVALID_ARG, "the events defender needs nd_ev_init, vi_svc_slot, out_slot and the folded ev_proj");
    UP(nd_ev_init, (size_t)Nn * 4); UP(out_slot, (size_t)t->num_ports_total * t->max_nodes);
  }
  UP(nd_ownable, Nn); UP(nd_discoverable, Nn); UP(nd_disruptable, Nn);
  UP(nd_row_off, 2 * (size_t)Nn + 1); UP(outblock, (size_t)t->num_ports_total * t->words);
  UP(uvuln_global, t->num_uvuln_total); UP(inst_of, t->num_instof);
  UP(vi_port, I); UP(vi_flags, I); UP(vi_kinds_any, I); UP(vi_kinds_remote, I); UP(vi_success, I); UP(vi_cost, I);
  UP(vi_recon_any, 2 * (size_t)I); UP(vi_recon_remote, 2 * (size_t)I); UP(vi_ulocal, I); UP(recon_nodes, t->num_recon);
  UP(row_packed, t->num_rows); UP(row_inst, t->num_rows);
  {   // derived: scenario-local vulnerability index of every candidate row (saves a dependent lookup in decode)
    std::vector<int32_t> ru((size_t)t->num_rows);
    for (int r = 0; r < t->num_rows; ++r) ru[r] = t->vi_ulocal[t->row_inst[r]];
    if ((rc = upload(h, &T.row_ulocal, ru.data(), ru.size()))) return rc;
  }
  {   // derived: packed scenario / instance records and back-to-back Reconnaissance lists (cbs_types.h, Tables::sc_pack ..)
    std::vector<int32_t> sp((size_t)S_ * 8, 0);
    for (int s = 0; s < S_; ++s) {
      int32_t* r = &sp[(size_t)s * 8];
      r[0] = t->sc_num_nodes[s]; r[1] = t->sc_node_off[s]; r[2] = t->sc_num_uvuln[s]; r[3] = t->sc_port_off[s];
      const int64_t io = t->sc_instof_off[s];
      r[4] = (int32_t)(uint32_t)(io & 0xFFFFFFFFll); r[5] = (int32_t)(io >> 32);
      r[6] = (h->cfg.goal >= GOAL_CONTROL_NODE) ? t->sc_interest[s] : -1;
    }
    if ((rc = upload(h, reinterpret_cast<const int32_t**>(&T.sc_pack), sp.data(), sp.size()))) return rc;
    std::vector<uint32_t> vp((size_t)(I > 0 ? I : 1) * 8, 0u);
    const std::vector<int32_t> inst_port_off = instance_port_offsets(t);
    std::vector<uint8_t> rp;
    std::vector<uint32_t> rmask((size_t)(I > 0 ? I : 1) * 2, 0u);
    for (int i = 0; i < I; ++i) {
      uint32_t* r = &vp[(size_t)i * 8];
      const int oa = t->vi_recon_any[2 * i], la = t->vi_recon_any[2 * i + 1];
      const int orr = t->vi_recon_remote[2 * i], lr = t->vi_recon_remote[2 * i + 1];
      if (la < 0 || la > CBS_MAX_NODES || lr < 0 || lr > CBS_MAX_NODES || (t->vi_flags[i] >> 8))   // (list lengths <= 128 < 256: bits 8..23)
        return fail(h, CBS_ERR_INVALID_ARG, "vulnerability instance %d: malformed flags / reconnaissance list", i);
      r[0] = t->vi_flags[i] | ((uint32_t)la << 8) | ((uint32_t)lr << 16) |
             ((uint32_t)(h->cfg.static_defender == 2 ? t->vi_svc_slot[i] : 0xFF) << 24);   // events defender: the target's service slot of the port
      r[1] = (uint32_t)t->vi_kinds_any[i] | ((uint32_t)t->vi_kinds_remote[i] << 16);
      // the port's outgoing-firewall node mask itself when a plane is one word, else the port index into `outblock`
      r[2] = (t->words == 1 && h->cfg.static_defender != 2) ? t->outblock[(size_t)(inst_port_off[i] + t->vi_port[i])] : (uint32_t)t->vi_port[i];
      r[3] = (uint32_t)rp.size();           // both lists start on an 8-byte boundary (the transition reads 8 ids per load)
      if (t->words == 1) {
        for (int k = 0; k < la; ++k) rmask[2 * (size_t)i] |= 1u << (t->recon_nodes[oa + k] & 31);
        for (int k = 0; k < lr; ++k) rmask[2 * (size_t)i + 1] |= 1u << (t->recon_nodes[orr + k] & 31);
      }
      rp.insert(rp.end(), t->recon_nodes + oa, t->recon_nodes + oa + la);
      rp.resize((rp.size() + 7) & ~size_t(7), 0);
      rp.insert(rp.end(), t->recon_nodes + orr, t->recon_nodes + orr + lr);
      rp.resize((rp.size() + 7) & ~size_t(7), 0);
      memcpy(&r[4], &t->vi_success[i], 8);
      memcpy(&r[6], &t->vi_cost[i], 8);
    }
    if ((rc = upload(h, reinterpret_cast<const uint32_t**>(&T.vi_pack), vp.data(), vp.size()))) return rc;
    if ((rc = upload(h, &T.recon_pack, rp.data(), rp.size()))) return rc;
    if ((rc = upload(h, reinterpret_cast<const uint32_t**>(&T.recon_mask), rmask.data(), rmask.size()))) return rc;
  }
  UP(vemb32, (size_t)t->num_global_vulns * VULN_EMB); UP(vemb64, (size_t)t->num_global_vulns * VULN_EMB);
  UP(vnorm2, t->num_global_vulns);
#undef UP
#define UPG(field, count) if ((rc = upload(h, &T.field, g->field, (size_t)(count)))) return rc
  UPG(node_static, (size_t)Nn * 2 * PROJ_ROWS * NODE_EMB); UPG(dyn_proj, NUM_DYN * PROJ_ROWS * NODE_EMB);
  UPG(vuln_h, (size_t)t->num_global_vulns * NN_CH); UPG(nn0_b, NN_CH); UPG(bn1_scale, NODE_EMB); UPG(bn1_shift, NODE_EMB);
  UPG(gcn_wt, NODE_EMB * NODE_EMB); UPG(bn2_scale, NODE_EMB); UPG(bn2_shift, NODE_EMB);
  if (h->cfg.static_defender == 2) UPG(ev_proj, 30 * PROJ_ROWS * NODE_EMB);
#undef UPG

  // capacities
  P.ncap = ((t->max_nodes + 3) / 4) * 4;
  P.words = t->words;
  int slots = 0, ecap = 0, max_steps = 0;
  derive_capacities(t->max_nodes, h->cfg.episode_iterations, P.prop_coeff, h->cfg.static_defender != 0, P.precise_positions != 0, &slots, &ecap,
                    &max_steps);
  if (slots > 255) {
    if (P.precise_positions && h->cfg.max_slots <= 0)
      return fail(h, CBS_ERR_INVALID_ARG, "precise_action_space_positions needs one snapshot slot per episode step: episodes of up to %d "
                                          "steps exceed the 255-slot limit", max_steps);
    slots = 255;
  }
  P.slots = h->cfg.max_slots > 0 ? h->cfg.max_slots : slots;
  if (P.slots > 255) return fail(h, CBS_ERR_INVALID_ARG, "max_slots must be <= 255");
  P.ecap = h->cfg.max_edges > 0 ? h->cfg.max_edges : ecap;
  h->Ug = t->num_global_vulns;
  h->vt_stride = ((h->Ug + 63) / 64) * 64;
  h->use_tc = (h->cfg.decode_gemm == 0) && decode_gemm_tc_available();
  // float32-scan error budget: half-precision snapshot copies (~6e-5 worst case) + TF32 products (~2e-5)
  P.margin = h->cfg.decode_margin > 0 ? h->cfg.decode_margin : (h->use_tc ? 1e-3f : 5e-4f);

  State& S = h->S;
  const size_t B = P.B;
  h->state_bytes = 0;
#define AL(field, count) if ((rc = dalloc(h, h->state_allocs, &S.field, (size_t)(count)))) return rc
  static_assert(N_SCALARS <= SCAL_PITCH, "scalar record does not fit its line");
  P.mpitch = ((N_MASKS * P.words + 15) / 16) * 16;
  AL(masks, (size_t)P.mpitch * B); AL(scal, (size_t)SCAL_PITCH * B);
  AL(disc_order, B * P.ncap); AL(owned_order, B * P.ncap); AL(pair_slot, B * P.ncap * P.ncap);
  P.ocap = 2 * P.ncap;
  if (P.defender) { AL(owned_raw, B * P.ocap); AL(reimage_left, B * P.ncap); AL(pair_opos, B * P.ncap * P.ncap); }
  else { AL(owned_raw, 1); AL(reimage_left, 1); AL(pair_opos, 1); }
  AL(pair_epoch, P.precise_positions ? B * P.ncap * P.ncap : 1);
  AL(changed, (P.precise_positions && P.defender) ? B * P.words : 1);
  AL(ev_cur, P.defender == 2 ? B * P.ncap * 4 : 1); AL(ev_x, P.defender == 2 ? B * P.ncap * 4 : 1);
  static_assert(OBS_DIM + NODE_EMB <= RC_Z, "reset-cache entry too small for the observation");
  AL(reset_cache, (size_t)Nn * RC_PITCH); AL(reset_cache_flag, Nn);
  AL(z_hist, B * P.slots * P.ncap * NODE_EMB); AL(zn2_hist, B * P.slots * P.ncap);
  AL(z16_hist, B * P.slots * P.ncap * NODE_EMB); AL(worklist, (size_t)OBS_CLASSES * B); AL(work_ctr, 4 + OBS_CLASSES + 2); AL(work_est, B); AL(bin_cnt, 2 * (SCHED_BINS + 1)); AL(bin_list, (size_t)2 * SCHED_BINS * B);
  AL(edge_src, B * P.ecap); AL(edge_dst, B * P.ecap); AL(edge_cnt, B * P.ecap);
  AL(edge_sum, B * P.ecap * NN_CH); AL(edge_m, B * P.ecap * NN_CH);
  AL(obs, B * P.obs_dim); AL(term_obs, B * P.obs_dim); AL(sel, B * 4); AL(dist, B); AL(reward64, B);
  AL(last_stats, B * 14); AL(accum, N_ACCUM); AL(vt, B * h->vt_stride); AL(errflag, 4);
  AL(scratch, P.ncap > CBS_OBS_SMEM_NODES ? B * 2 * P.ncap * NODE_EMB : 1);   // only graphs beyond the shared-memory buffers use it
  AL(sub_rows, P.subset_k ? B * SUB_CLASSES * P.subset_k : 1); AL(sub_meta, P.subset_k ? B * SUB_META : 1);
  AL(sub_alive, (P.subset_k && P.precise_positions) ? B * P.ncap * P.ncap * (SUB_MAX_ROWS_PER_PAIR / 32) : 1);
  AL(sub_newp, (P.subset_k && P.ncap > CBS_OBS_SMEM_NODES) ? B * P.ncap * P.ncap : 1);
#undef AL
  if ((rc = dalloc(h, h->state_allocs, &h->d_sel, B * 4))) return rc;
  if ((rc = dalloc(h, h->state_allocs, &h->d_dist, B))) return rc;
  if (h->use_tc && (rc = dalloc(h, h->state_allocs, &h->a_packed, B * VULN_EMB))) return rc;
  if (h->use_tc) {
    const size_t n = (size_t)h->Ug * VULN_EMB;
    if ((rc = dalloc(h, h->state_allocs, &h->vemb16, n, false))) return rc;
    CK(h, convert_vemb_f16(h->T.vemb32, h->vemb16, n, 0));
    CK(h, cudaDeviceSynchronize());
  }
  if (P.metric != METRIC_COSINE && (rc = dalloc(h, h->state_allocs, &h->vt64, B * (size_t)h->Ug, false))) return rc;
  init_flags_kernel<<<(P.B + 255) / 256, 256>>>(S.scal, P.B);
  CK(h, cudaGetLastError());
  CK(h, cudaDeviceSynchronize());
  h->loaded = true;
  return CBS_OK;
}

int cbs_set_scenarios(cbs_handle* h, const int32_t* sc_host) {
  if (!h || !sc_host) return fail(h, CBS_ERR_INVALID_ARG, "cbs_set_scenarios: null argument");
  if (!h->loaded) return fail(h, CBS_ERR_NOT_READY, "load scenarios first");
  for (int b = 0; b < h->P.B; ++b)
    if (sc_host[b] < 0 || sc_host[b] >= h->T.num_scenarios) return fail(h, CBS_ERR_INVALID_ARG, "scenario id %d out of range (env %d)", sc_host[b], b);
  CK(h, cudaSetDevice(h->cfg.device));
  // one int32 per env into the sector that holds S_SCENARIO (State::scal is [sector][B][8])
  CK(h, cudaMemcpy2D(h->S.scal + ((size_t)(S_SCENARIO >> 3) * h->P.B) * 8 + (S_SCENARIO & 7), 8 * sizeof(int32_t), sc_host,
                     sizeof(int32_t), sizeof(int32_t), h->P.B, cudaMemcpyHostToDevice));
  return CBS_OK;
}

int cbs_set_starter_queue(cbs_handle* h, const int32_t* q, int32_t qlen) {
  if (!h) return CBS_ERR_INVALID_ARG;
  if (!h->loaded) return fail(h, CBS_ERR_NOT_READY, "load scenarios first");
  CK(h, cudaSetDevice(h->cfg.device));
  CK(h, cudaDeviceSynchronize());
  if (!q || qlen <= 0) { h->S.starter_queue = nullptr; h->P.qlen = 0; return CBS_OK; }
  int32_t* d = nullptr;
  int rc = dalloc(h, h->state_allocs, &d, (size_t)h->P.B * qlen, false);
  if (rc) return rc;
  CK(h, cudaMemcpy(d, q, sizeof(int32_t) * (size_t)h->P.B * qlen, cudaMemcpyHostToDevice));
  h->S.starter_queue = d;
  h->P.qlen = qlen;
  return CBS_OK;
}

int cbs_set_action_stride(cbs_handle* h, int32_t stride_floats) {
  if (!h) return CBS_ERR_INVALID_ARG;
  if (stride_floats < ACTION_DIM) return fail(h, CBS_ERR_INVALID_ARG, "action stride must be >= %d floats", ACTION_DIM);
  h->P.act_stride = stride_floats;
  return CBS_OK;
}

int cbs_set_actions_prestaged(cbs_handle* h, int32_t on) {
  if (!h) return CBS_ERR_INVALID_ARG;
  h->actions_prestaged = on != 0;
  return CBS_OK;
}

int cbs_set_defender_draws(cbs_handle* h, const int32_t* scan_nodes_dev, const float* detect_uniforms_dev) {
  if (!h) return CBS_ERR_INVALID_ARG;
  if (!h->P.defender) return fail(h, CBS_ERR_INVALID_ARG, "no static defender configured");
  h->S.def_nodes = scan_nodes_dev;
  h->S.def_uniforms = detect_uniforms_dev;
  return CBS_OK;
}

int cbs_set_cutoffs(cbs_handle* h, int32_t episode_iterations, double prop) {
  if (!h) return CBS_ERR_INVALID_ARG;
  if (h->loaded) {   // the per-env buffers were sized from the cut-offs in force at load time: longer episodes must still fit
    int slots = 0, ecap = 0, max_steps = 0;
    derive_capacities(h->T.max_nodes, episode_iterations, prop, h->P.defender != 0, h->P.precise_positions != 0, &slots, &ecap, &max_steps);
    if (slots > 255) slots = 255;
    if (slots > h->P.slots || ecap > h->P.ecap)
      return fail(h, CBS_ERR_CAPACITY, "cut-offs (%d iterations, coefficient %g) allow episodes of %d steps, which need %d snapshot slots and "
                  "%d edges per env; this handle was created with %d / %d.  Create it with cbs_config.max_slots / max_edges (or the "
                  "larger cut-offs) instead", episode_iterations, prop, max_steps, slots, ecap, h->P.slots, h->P.ecap);
  }
  h->P.episode_iterations = episode_iterations;
  h->P.prop_coeff = prop;
  return CBS_OK;
}

static int launch_gemm(cbs_handle* h, const float* actions_dev, cudaStream_t st, bool prestaged = false);

static int check_ready(cbs_handle* h) {
  if (!h) return CBS_ERR_INVALID_ARG;
  if (!h->loaded) return fail(h, CBS_ERR_NOT_READY, "scenarios not loaded");
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return fail(h, CBS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return 0;
}

int cbs_reset(cbs_handle* h, const uint8_t* env_mask_dev, float* obs_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  CK(h, launch_observe(h->T, h->P, h->S, env_mask_dev, 1, h->num_sms, (cudaStream_t)stream));
  h->launches += 1;
  if (obs_dev) CK(h, cudaMemcpyAsync(obs_dev, h->S.obs, (size_t)h->P.B * h->P.obs_dim * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CBS_OK;
}

int cbs_decode(cbs_handle* h, const float* actions_dev, int32_t* sel_dev, double* dist_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions_dev) return fail(h, CBS_ERR_INVALID_ARG, "cbs_decode: actions is null");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->P.metric != METRIC_COSINE) {
    CK(h, launch_decode_metric(h->T, h->P, h->S, actions_dev, h->vt64, h->Ug, h->sched_buf, sel_dev, dist_dev, st));
    h->launches += 2;
    return CBS_OK;
  }
  if ((rc = launch_gemm(h, actions_dev, st))) return rc;
  CK(h, launch_decode_select(h->T, h->P, h->S, actions_dev, h->vt_stride, h->sched_buf, 0, nullptr, nullptr, nullptr, sel_dev, dist_dev, st));
  h->launches += 1;
  return CBS_OK;
}

int cbs_transition(cbs_handle* h, const int32_t* sel_dev, const double* dist_dev, const float* uniforms_dev, float* reward_dev,
                   uint8_t* done_dev, uint8_t* truncated_dev, uint8_t* outcome_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!sel_dev) return fail(h, CBS_ERR_INVALID_ARG, "cbs_transition: sel is null");
  CK(h, launch_transition(h->T, h->P, h->S, sel_dev, dist_dev, uniforms_dev, h->sched_buf ^ 1, reward_dev, done_dev, truncated_dev,
                          outcome_dev, h->num_sms, (cudaStream_t)stream));
  h->sched_buf ^= 1;
  h->launches += 1;
  return CBS_OK;
}

int cbs_transition_ksteps(cbs_handle* h, const int32_t* sel_dev, const double* dist_dev, const float* uniforms_dev, int32_t k_steps,
                          float* reward_dev, uint8_t* done_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!sel_dev || k_steps < 1) return fail(h, CBS_ERR_INVALID_ARG, "cbs_transition_ksteps: sel is null or k_steps < 1");
  if (h->P.defender || h->P.words != 1)
    return fail(h, CBS_ERR_INVALID_ARG, "cbs_transition_ksteps keeps an env's record in registers: scenarios of <= 32 nodes, no static defender");
  CK(h, launch_transition_ksteps(h->T, h->P, h->S, sel_dev, dist_dev, uniforms_dev, k_steps, reward_dev, done_dev, (cudaStream_t)stream));
  h->launches += 1;
  return CBS_OK;
}

int cbs_observe(cbs_handle* h, float* obs_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  CK(h, launch_observe(h->T, h->P, h->S, nullptr, 0, h->num_sms, (cudaStream_t)stream));
  h->launches += 1;
  if (obs_dev) CK(h, cudaMemcpyAsync(obs_dev, h->S.obs, (size_t)h->P.B * h->P.obs_dim * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CBS_OK;
}

// prestaged: this call's actions are known to be complete (cbs_replay's later steps); else the handle's declaration decides
static int launch_gemm(cbs_handle* h, const float* actions_dev, cudaStream_t st, bool prestaged) {
  const int wait_early = (prestaged || h->actions_prestaged) ? 0 : 1;
  if (h->use_tc) {
    if (h->vemb16 && decode_gemm_f16_applies(h->P.B, h->Ug))
      CK(h, launch_decode_gemm_f16(actions_dev, h->P.act_stride, h->vemb16, h->S.vt, h->P.B, h->Ug, h->vt_stride, h->S.errflag, wait_early, st));
    else
      CK(h, launch_decode_gemm_tc(actions_dev, h->P.act_stride, h->T.vemb32, h->a_packed, h->S.vt, h->P.B, h->Ug, h->vt_stride, h->S.errflag, wait_early, st));
    h->launches += 1;
  } else {
    CK(h, launch_decode_gemm_simt(actions_dev, h->P.act_stride, h->T.vemb32, h->S.vt, h->P.B, h->Ug, h->vt_stride, st));
    h->launches += 1;
  }
  return CBS_OK;
}

int cbs_step(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, float* obs_dev, float* reward_dev,
             uint8_t* done_dev, int32_t* info_dev, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions_dev) return fail(h, CBS_ERR_INVALID_ARG, "cbs_step: actions is null");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->P.metric != METRIC_COSINE) {
    // l1 / l2 / inf decode (k_decode_metric.cu), then the transition as its own launch on the handle's own sel / dist
    CK(h, launch_decode_metric(h->T, h->P, h->S, actions_dev, h->vt64, h->Ug, h->sched_buf, nullptr, nullptr, st));
    CK(h, launch_transition(h->T, h->P, h->S, h->S.sel, h->S.dist, uniforms_dev, h->sched_buf ^ 1, reward_dev, done_dev, nullptr, nullptr, h->num_sms, st));
    h->sched_buf ^= 1;
    h->launches += 3;
  } else {
    if ((rc = launch_gemm(h, actions_dev, st))) return rc;
    // fused: every warp runs the transition of its env right after decoding it (no separate transition launch)
    CK(h, launch_decode_select(h->T, h->P, h->S, actions_dev, h->vt_stride, h->sched_buf, 1, uniforms_dev, reward_dev, done_dev,
                               nullptr, nullptr, st));
    h->sched_buf ^= 1;
    h->launches += 1;
  }
  if (info_dev) {   // before observe: an auto-reset clears the per-step flags
    info_kernel<<<(h->P.B + 255) / 256, 256, 0, st>>>(h->P, h->S, info_dev);
    CK(h, cudaGetLastError());
    h->launches += 1;
  }
  return cbs_observe(h, obs_dev, stream);
}

int cbs_replay(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, int32_t num_steps, const cbs_replay_log* log,
               uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions_dev || num_steps < 1) return fail(h, CBS_ERR_INVALID_ARG, "cbs_replay: actions is null or num_steps < 1");
  if (log && (log->first_env < 0 || log->num_logged < 0 || log->first_env + log->num_logged > h->P.B))
    return fail(h, CBS_ERR_INVALID_ARG, "cbs_replay: logged env range [%d, %d) outside the handle's %d envs", log->first_env,
                log->first_env + log->num_logged, h->P.B);
  if (log && log->num_logged > 0 && !log->meta && (log->obs || log->reset_obs || log->reset_masks || log->stats))
    return fail(h, CBS_ERR_INVALID_ARG, "cbs_replay: the post-observe log fields need `meta` (it carries the step's done flag)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = h->P.B;
  const int owned_len = h->P.defender ? h->P.ocap : h->P.ncap;
  const bool logging = log && log->num_logged > 0;
  const int lgrid = logging ? (log->num_logged + 3) / 4 : 0;
  for (int t = 0; t < num_steps; ++t) {
    const float* a = actions_dev + (size_t)t * B * h->P.act_stride;
    const float* u = uniforms_dev ? uniforms_dev + (size_t)t * B : nullptr;
    const bool forced = log && log->force_steps_host && log->force_sel && log->force_dist && log->force_steps_host[t];
    if (h->P.metric != METRIC_COSINE || forced) {
      if (h->P.metric != METRIC_COSINE) {
        CK(h, launch_decode_metric(h->T, h->P, h->S, a, h->vt64, h->Ug, h->sched_buf, nullptr, nullptr, st));
        h->launches += 2;
      } else {
        if ((rc = launch_gemm(h, a, st, t > 0))) return rc;
        CK(h, launch_decode_select(h->T, h->P, h->S, a, h->vt_stride, h->sched_buf, 0, nullptr, nullptr, nullptr, nullptr, nullptr, st));
        h->launches += 1;
      }
      if (forced) {
        replay_force_kernel<<<(h->P.B + 255) / 256, 256, 0, st>>>(h->P, h->S, log->force_sel + (size_t)t * B * 4, log->force_dist + (size_t)t * B);
        h->launches += 1;
      }
      CK(h, launch_transition(h->T, h->P, h->S, h->S.sel, h->S.dist, u, h->sched_buf ^ 1, nullptr, nullptr, nullptr, nullptr, h->num_sms, st));
      h->launches += 1;
    } else {
      if ((rc = launch_gemm(h, a, st, t > 0))) return rc;   // (the whole [T, B, pitch] slab was staged before the call)
      CK(h, launch_decode_select(h->T, h->P, h->S, a, h->vt_stride, h->sched_buf, 1, u, nullptr, nullptr, nullptr, nullptr, st));
      h->launches += 1;
    }
    h->sched_buf ^= 1;
    if (logging) {
      replay_log_kernel<<<lgrid, 128, 0, st>>>(h->P, h->S, *log, t, 0, owned_len);
      h->launches += 1;
    }
    CK(h, launch_observe(h->T, h->P, h->S, nullptr, 0, h->num_sms, st));
    h->launches += 1;
    if (logging) {
      replay_log_kernel<<<lgrid, 128, 0, st>>>(h->P, h->S, *log, t, 1, owned_len);
      h->launches += 1;
    }
    CK(h, cudaGetLastError());
  }
  return CBS_OK;
}

int cbs_profile_step(cbs_handle* h, const float* actions_dev, const float* uniforms_dev, float* out_ms, uintptr_t stream) {
  // the same three launches as cbs_step with a CUDA event between them, on the launching stream; out_ms[5] =
  // { decode_gemm, decode_select + transition (fused kernel), observe, 0, 0 } in milliseconds.  Synchronises.
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions_dev || !out_ms) return fail(h, CBS_ERR_INVALID_ARG, "cbs_profile_step: null argument");
  if (h->P.metric != METRIC_COSINE) return fail(h, CBS_ERR_INVALID_ARG, "cbs_profile_step times the cosine decode's kernels only");
  cudaStream_t st = (cudaStream_t)stream;
  struct Events {   // destroyed on every return path (CK returns early on a CUDA error)
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } evs;
  cudaEvent_t* ev = evs.e;
  for (int i = 0; i < 4; ++i) CK(h, cudaEventCreate(&ev[i]));
  CK(h, cudaEventRecord(ev[0], st));
  if ((rc = launch_gemm(h, actions_dev, st))) return rc;
  CK(h, cudaEventRecord(ev[1], st));
  CK(h, launch_decode_select(h->T, h->P, h->S, actions_dev, h->vt_stride, h->sched_buf, 1, uniforms_dev, nullptr, nullptr, nullptr,
                             nullptr, st));
  h->sched_buf ^= 1;
  CK(h, cudaEventRecord(ev[2], st));
  CK(h, launch_observe(h->T, h->P, h->S, nullptr, 0, h->num_sms, st));
  CK(h, cudaEventRecord(ev[3], st));
  CK(h, cudaStreamSynchronize(st));
  for (int i = 0; i < 3; ++i) CK(h, cudaEventElapsedTime(&out_ms[i], ev[i], ev[i + 1]));
  out_ms[3] = out_ms[4] = 0.f;
  h->launches += 2;
  return CBS_OK;
}

static int ensure_host_staging(cbs_handle* h) {
  if (h->hstream) return 0;
  const size_t B = h->P.B;
  int rc;
  CK(h, cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
  if ((rc = dalloc(h, h->state_allocs, &h->h_actions, B * ACTION_DIM, false))) return rc;
  if ((rc = dalloc(h, h->state_allocs, &h->h_uniforms, B, false))) return rc;
  if ((rc = dalloc(h, h->state_allocs, &h->h_reward, B, false))) return rc;
  if ((rc = dalloc(h, h->state_allocs, &h->h_done, B, false))) return rc;
  if ((rc = dalloc(h, h->state_allocs, &h->h_info, B * CBS_INFO_INTS, false))) return rc;
  CK(h, cudaMallocHost(&h->h_errflag, sizeof(int32_t)));
  *h->h_errflag = 0;
  return 0;
}

int cbs_step_host_async(cbs_handle* h, const float* actions_host, const float* uniforms_host, float* obs_host, float* reward_host,
                        uint8_t* done_host, int32_t* info_host) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!actions_host) return fail(h, CBS_ERR_INVALID_ARG, "cbs_step_host: actions is null");
  if ((rc = ensure_host_staging(h))) return rc;
  const size_t B = h->P.B;
  cudaStream_t st = h->hstream;
  // one dense copy: a pitched (2-D) host-to-device copy of 3620-byte rows runs at less than half the PCIe rate
  // (measured 4.4 M vs 9.5 M env-steps/s end to end), which costs more than the repack kernel it would save
  CK(h, cudaMemcpyAsync(h->h_actions, actions_host, B * ACTION_DIM * sizeof(float), cudaMemcpyHostToDevice, st));
  const int saved_stride = h->P.act_stride;
  h->P.act_stride = ACTION_DIM;
  if (uniforms_host) CK(h, cudaMemcpyAsync(h->h_uniforms, uniforms_host, B * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = cbs_step(h, h->h_actions, uniforms_host ? h->h_uniforms : nullptr, nullptr, h->h_reward, h->h_done,
                info_host ? h->h_info : nullptr, (uintptr_t)st);
  h->P.act_stride = saved_stride;
  if (rc) return rc;
  if (obs_host) CK(h, cudaMemcpyAsync(obs_host, h->S.obs, B * h->P.obs_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (reward_host) CK(h, cudaMemcpyAsync(reward_host, h->h_reward, B * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (done_host) CK(h, cudaMemcpyAsync(done_host, h->h_done, B, cudaMemcpyDeviceToHost, st));
  if (info_host) CK(h, cudaMemcpyAsync(info_host, h->h_info, B * CBS_INFO_INTS * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(h, cudaMemcpyAsync(h->h_errflag, h->S.errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));   // checked by cbs_host_sync
  return CBS_OK;
}

int cbs_host_sync(cbs_handle* h) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (h->hstream) CK(h, cudaStreamSynchronize(h->hstream));
  if (h->h_errflag && *h->h_errflag) return fail(h, CBS_ERR_CAPACITY, kErrflagMessage, *h->h_errflag);
  return CBS_OK;
}

int cbs_step_host(cbs_handle* h, const float* actions_host, const float* uniforms_host, float* obs_host, float* reward_host,
                  uint8_t* done_host, int32_t* info_host) {
  int rc = cbs_step_host_async(h, actions_host, uniforms_host, obs_host, reward_host, done_host, info_host);
  if (rc) return rc;
  return cbs_host_sync(h);
}

static int field_ptr(cbs_handle* h, int32_t field, void** p, int64_t* bytes) {
  const Params& P = h->P;
  const State& S = h->S;
  const int64_t B = P.B;
  switch (field) {
    case CBS_F_MASKS: *p = S.masks; *bytes = (int64_t)P.mpitch * B * 4; break;
    case CBS_F_DISC_ORDER: *p = S.disc_order; *bytes = B * P.ncap; break;
    case CBS_F_OWNED_ORDER: *p = S.owned_order; *bytes = B * P.ncap; break;
    case CBS_F_OWNED_RAW: *p = S.owned_raw; *bytes = P.defender ? B * P.ocap : 1; break;
    case CBS_F_Z_HIST: *p = S.z_hist; *bytes = B * P.slots * P.ncap * NODE_EMB * 4; break;
    case CBS_F_REIMAGE_LEFT: *p = S.reimage_left; *bytes = P.defender ? B * P.ncap : 1; break;
    case CBS_F_SCALARS: *p = S.scal; *bytes = (int64_t)SCAL_PITCH * B * 4; break;
    case CBS_F_TERMINAL_OBS: *p = S.term_obs; *bytes = B * P.obs_dim * 4; break;
    case CBS_F_OBS: *p = S.obs; *bytes = B * P.obs_dim * 4; break;
    case CBS_F_LAST_STATS: *p = S.last_stats; *bytes = B * 14 * 8; break;
    case CBS_F_STAT_ACCUM: *p = S.accum; *bytes = N_ACCUM * 8; break;
    case CBS_F_PAIR_SLOT: *p = S.pair_slot; *bytes = B * P.ncap * P.ncap; break;
    case CBS_F_DIST: *p = S.dist; *bytes = B * 8; break;
    case CBS_F_REWARD64: *p = S.reward64; *bytes = B * 8; break;
    case CBS_F_ERRFLAG: *p = S.errflag; *bytes = 4; break;
    case CBS_F_SEL: *p = S.sel; *bytes = B * 16; break;
    case CBS_F_DIVERGENCE: *p = S.errflag + 1; *bytes = 4; break;
    case CBS_F_MARGIN_EDGE: *p = S.errflag + 2; *bytes = 4; break;
    case CBS_F_EV_CUR: *p = S.ev_cur; *bytes = P.defender == 2 ? B * P.ncap * 8 : 2; break;
    case CBS_F_EV_X: *p = S.ev_x; *bytes = P.defender == 2 ? B * P.ncap * 8 : 2; break;
    case CBS_F_VT: *p = S.vt; *bytes = B * h->vt_stride * 4; break;
    default: return fail(h, CBS_ERR_INVALID_ARG, "unknown state field %d", field);
  }
  return 0;
}

int64_t cbs_read_state(cbs_handle* h, int32_t field, void* dst_host, int64_t bytes) {
  int rc = check_ready(h);
  if (rc) return rc;
  void* p = nullptr;
  int64_t n = 0;
  if ((rc = field_ptr(h, field, &p, &n))) return rc;
  if (!dst_host) return n;
  if (bytes != n) return fail(h, CBS_ERR_INVALID_ARG, "field %d holds %lld bytes, caller passed %lld", field, (long long)n, (long long)bytes);
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(dst_host, p, (size_t)n, cudaMemcpyDeviceToHost));
  return n;
}

void* cbs_state_ptr(cbs_handle* h, int32_t field) {
  if (!h || !h->loaded) return nullptr;
  void* p = nullptr;
  int64_t n = 0;
  if (field_ptr(h, field, &p, &n)) return nullptr;
  return p;
}

int cbs_get_state(cbs_handle* h, cbs_state_view* out) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!out) return fail(h, CBS_ERR_INVALID_ARG, "cbs_get_state: null argument");
  const Params& P = h->P;
  const State& S = h->S;
  out->num_envs = P.B; out->max_nodes = P.ncap; out->words = P.words; out->mask_pitch = P.mpitch; out->scalar_pitch = SCAL_PITCH;
  out->obs_dim = P.obs_dim; out->slots = P.slots; out->num_masks = N_MASKS;
  out->masks = S.masks; out->scalars = S.scal; out->disc_order = S.disc_order; out->owned_order = S.owned_order;
  out->pair_slot = S.pair_slot; out->obs = S.obs; out->terminal_obs = S.term_obs; out->sel = S.sel; out->dist = S.dist;
  out->reward64 = S.reward64; out->last_stats = S.last_stats; out->stat_accum = S.accum;
  return CBS_OK;
}

int cbs_episode_stats(cbs_handle* h, double* out_host) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!out_host) return fail(h, CBS_ERR_INVALID_ARG, "cbs_episode_stats: null argument");
  CK(h, cudaDeviceSynchronize());
  CK(h, cudaMemcpy(out_host, h->S.last_stats, (size_t)h->P.B * 14 * sizeof(double), cudaMemcpyDeviceToHost));
  return CBS_OK;
}

int cbs_reset_stat_accum(cbs_handle* h, uintptr_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  CK(h, cudaMemsetAsync(h->S.accum, 0, N_ACCUM * sizeof(double), (cudaStream_t)stream));
  return CBS_OK;
}

int cbs_debug_select_trace(cbs_handle* h, long long* trace_dev) {
  (void)h;
  cbs::g_sel_trace = trace_dev;   // [num_envs][6] int64 device buffer, or NULL to switch tracing off
  return CBS_OK;
}

int cbs_debug_observe_trace(cbs_handle* h, long long* trace_dev) {
  (void)h;
  cbs::g_obs_trace = trace_dev;   // [num_envs][4] int64 device buffer, or NULL to switch tracing off
  return CBS_OK;
}

int64_t cbs_launch_count(const cbs_handle* h) { return h ? h->launches : 0; }

int cbs_sync(cbs_handle* h) {
  int rc = check_ready(h);
  if (rc) return rc;
  CK(h, cudaDeviceSynchronize());
  int flag = 0;
  CK(h, cudaMemcpy(&flag, h->S.errflag, 4, cudaMemcpyDeviceToHost));
  if (flag) return fail(h, CBS_ERR_CAPACITY, kErrflagMessage, flag);
  return CBS_OK;
}

int cbs_struct_sizes(int32_t* out3) {
  if (!out3) return CBS_ERR_INVALID_ARG;
  out3[0] = (int32_t)sizeof(cbs_config); out3[1] = (int32_t)sizeof(cbs_scenario_tables); out3[2] = (int32_t)sizeof(cbs_gae_tables);
  return CBS_OK;
}

int64_t cbs_state_bytes(const cbs_handle* h) { return h ? (int64_t)h->state_bytes : 0; }
int cbs_capacities(const cbs_handle* h, int32_t* out4 /* 8 ints */) {
  if (!h || !out4) return CBS_ERR_INVALID_ARG;
  out4[0] = h->P.ncap; out4[1] = h->P.slots; out4[2] = h->P.ecap; out4[3] = h->use_tc ? 1 : 0;
  out4[4] = h->vt_stride;
  out4[5] = h->P.obs_dim;
  out4[6] = h->P.mpitch;
  out4[7] = SCAL_PITCH;
  return CBS_OK;
}

}  // extern "C"
