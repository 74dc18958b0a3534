// engine.cu -- host side of libjxb.so: the C ABI of include/jxb.h.
//
// One engine per process, bound to one B200.  A model owns its struct-of-arrays agent
// state, env scalars, key table and history ring in HBM; jxb_model_run() enqueues the
// whole time loop (CUDA graph of fused step kernels) on the engine's stream and reads
// the history back once at the end (jaxabm/model.py:218-262 without the per-step host
// round trip).  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jxb.h"
#include "common.cuh"
#include "rules.cuh"
#include "schelling.cuh"
#include "schelling_bits.cuh"
#include "grid_shard.cuh"
#include "sir.cuh"
#include "ensemble.cuh"
#include "record.cuh"

using namespace jxb;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(JXB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                  cudaGetErrorString(_e));                                                \
  } while (0)

extern "C" const char* jxb_last_error(void) { return g_err; }
extern "C" int jxb_version(void) { return JXB_VERSION; }

// ---------------------------------------------------------------------------------------
// registries: state layout of every rule, env/metric layout of every program
// dtype: 0 f32, 1 i32, 2 bool(u8), 3 f64 (a Python float in the reference)
// ---------------------------------------------------------------------------------------
struct FieldSpec { const char* name; int dtype; int width; };
struct RuleSpec { int rule; int nf; FieldSpec f[kMaxFields]; };
struct SlotSpec { const char* name; int dtype; double dflt; };
struct ProgramSpec {
  int program; int has_env_fn;
  int n_env; SlotSpec env[kMaxEnv];
  int n_metrics; SlotSpec metrics[kMaxMetrics];
};

static const RuleSpec kRules[] = {
    {JXB_RULE_RANDOM_WALKER, 4, {{"position", 0, 2}, {"velocity", 0, 2}, {"color", 1, 1}, {"steps_taken", 1, 1}}},
    {JXB_RULE_SCALED_WALKER, 4, {{"position", 0, 2}, {"velocity", 0, 2}, {"color", 1, 1}, {"steps_taken", 1, 1}}},
    {JXB_RULE_CONSUMER, 4, {{"savings", 0, 1}, {"consumption", 0, 1}, {"utility", 0, 1}, {"income", 0, 1}}},
    {JXB_RULE_PRODUCER, 3, {{"capital", 0, 1}, {"production", 0, 1}, {"profit", 0, 1}}},
    {JXB_RULE_GROWTH, 1, {{"value", 0, 1}}},
    {JXB_RULE_INCREMENT, 1, {{"value", 0, 1}}},
    {JXB_RULE_WEALTH, 2, {{"wealth", 0, 1}, {"productivity", 0, 1}}},
    {JXB_RULE_SCHELLING, 4, {{"type", 1, 1}, {"position", 1, 2}, {"satisfied", 2, 1}, {"moves", 1, 1}}},
    {JXB_RULE_SIR, 1, {{"state", 1, 1}}},
    {JXB_RULE_HOUSEHOLD, 15,
     {{"savings", 0, 1}, {"income", 0, 1}, {"bank_deposits", 0, 1}, {"cash", 0, 1}, {"debt", 0, 1},
      {"propensity_to_consume", 0, 1}, {"propensity_to_save", 0, 1}, {"risk_aversion", 0, 1}, {"employed", 2, 1},
      {"productivity", 0, 1}, {"labor_supply", 0, 1}, {"consumption", 0, 1}, {"utility", 0, 1},
      {"taxes_paid", 0, 1}, {"transfers_received", 0, 1}}},
    {JXB_RULE_CONSUMER_FIRM, 19,
     {{"capital_stock", 0, 1}, {"production_capacity", 0, 1}, {"inventory", 0, 1}, {"cash", 0, 1}, {"revenue", 0, 1},
      {"profit", 0, 1}, {"debt", 0, 1}, {"production_efficiency", 0, 1}, {"labor_demand", 0, 1},
      {"energy_usage", 0, 1}, {"goods_produced", 0, 1}, {"goods_sold", 0, 1}, {"price", 0, 1}, {"markup_rate", 0, 1},
      {"labor_elasticity", 0, 1}, {"capital_elasticity", 0, 1}, {"energy_elasticity", 0, 1}, {"age", 1, 1},
      {"is_active", 2, 1}}},
};

static const ProgramSpec kPrograms[] = {
    {JXB_PROGRAM_NONE, 0, 0, {}, 0, {}},
    {JXB_PROGRAM_RANDOM_WALK, 1, 7,
     {{"bounds_lo", 0, 0.0}, {"bounds_hi", 0, 1.0}, {"time", 1, 0}, {"mean_x", 3, 0.5}, {"mean_y", 3, 0.5},
      {"num_red", 1, 0}, {"num_blue", 1, 0}},
     7,
     {{"mean_x", 3, 0}, {"mean_y", 3, 0}, {"mean_distance", 0, 0}, {"max_distance", 0, 0}, {"num_red", 1, 0},
      {"num_blue", 1, 0}, {"time", 1, 0}}},
    {JXB_PROGRAM_MARKET, 1, 5,
     {{"price_level", 0, 1.0}, {"gdp", 0, 0}, {"unemployment", 0, 0}, {"total_consumption", 0, 0},
      {"total_production", 0, 0}},
     5,
     {{"gdp", 0, 0}, {"price_level", 0, 0}, {"unemployment", 0, 0}, {"avg_utility", 0, 0}, {"avg_profit", 0, 0}}},
    {JXB_PROGRAM_GROWTH, 1, 2, {{"price_level", 3, 1.0}, {"interest_rate", 3, 0.05}}, 3,
     {{"avg_value", 0, 0}, {"price_level", 3, 0}, {"price_gap", 0, 0}}},
    {JXB_PROGRAM_COUNTER, 1, 2, {{"counter", 1, 0}, {"increment", 0, 1.0}}, 2,
     {{"total_value", 0, 0}, {"step_counter", 1, 0}}},
    {JXB_PROGRAM_SCHELLING, 1, 3,
     {{"segregation_index", 0, 0}, {"percent_satisfied", 0, 0}, {"total_moves", 1, 0}}, 3,
     {{"percent_satisfied", 0, 0}, {"segregation_index", 0, 0}, {"total_moves", 1, 0}}},
    {JXB_PROGRAM_SIR, 0, 0, {}, 3, {{"count_S", 1, 0}, {"count_I", 1, 0}, {"count_R", 1, 0}}},
    // env slot order == the EE_* enum of csrc/economy.cuh; defaults are the .get() defaults of the reference
    {JXB_PROGRAM_ECONOMY, 1, EE_COUNT,
     {{"wage_rate", 0, 1.0}, {"price_level", 0, 1.0}, {"interest_rate", 0, 0.05}, {"gdp", 0, 0.1}, {"gdp_growth", 0, 0.0},
      {"total_labor_supply", 0, 0.0}, {"total_labor_demand", 0, 0.0}, {"employment_rate", 0, 1.0},
      {"unemployment_rate", 0, 0.0}, {"time_step", 1, 0}, {"climate_impact", 3, 1.0}, {"pandemic_impact", 3, 1.0},
      {"tax_rate", 3, 0.2}, {"energy_price", 3, 1.0}, {"job_market_condition", 3, 1.0}, {"goods_availability", 3, 1.0},
      {"consumer_goods_demand", 3, 100.0}, {"consumer_goods_price", 3, 1.0}, {"capital_goods_price", 3, 2.0},
      {"inflation_rate", 3, 0.0}, {"debt_to_gdp", 3, 0.0}, {"avg_utility", 3, 0.0}, {"income_per_capita", 3, 0.0},
      {"govt_spending", 3, 0.0}, {"energy_supply", 3, 0.0}, {"capital_goods_demand", 3, 0.0}, {"total_income", 0, 0.0},
      {"_total_income_set", 1, 0}, {"_gini", 0, 0.0}},
     29,
     {{"gdp", 0, 0}, {"gdp_growth", 0, 0}, {"inflation", 0, 0}, {"unemployment", 0, 0}, {"wage_rate", 0, 0},
      {"interest_rate", 0, 0}, {"goods_availability", 0, 0}, {"labor_market_tightness", 0, 0}, {"consumer_price", 0, 0},
      {"capital_price", 0, 0}, {"energy_price", 0, 0}, {"utility", 0, 0}, {"income_per_capita", 0, 0},
      {"inequality", 0, 0}, {"consumer_sector_share", 0, 0}, {"capital_sector_share", 0, 0},
      {"energy_sector_share", 0, 0}, {"govt_sector_share", 0, 0}, {"technology_level", 0, 0},
      {"capital_investment", 0, 0}, {"energy_demand", 0, 0}, {"energy_supply", 0, 0}, {"renewable_share", 0, 0},
      {"carbon_emissions", 0, 0}, {"sustainability_index", 0, 0}, {"debt_to_gdp", 0, 0}, {"climate_impact", 0, 0},
      {"pandemic_impact", 0, 0}, {"economic_health", 0, 0}}},
};

static const RuleSpec* find_rule(int rule) {
  for (const auto& r : kRules)
    if (r.rule == rule) return &r;
  return nullptr;
}
static const ProgramSpec* find_program(int p) {
  for (const auto& r : kPrograms)
    if (r.program == p) return &r;
  return nullptr;
}
static size_t dtype_size(int dt) { return dt == 2 ? 1 : (dt == 3 ? 8 : 4); }

// ---------------------------------------------------------------------------------------
// NCCL through dlopen (only touched when a population is sharded across processes)
// ---------------------------------------------------------------------------------------
struct NcclId { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return JXB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail(JXB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  *(void**)&g_nccl.GetUniqueId = dlsym(g_nccl.lib, "ncclGetUniqueId");
  *(void**)&g_nccl.CommInitRank = dlsym(g_nccl.lib, "ncclCommInitRank");
  *(void**)&g_nccl.AllReduce = dlsym(g_nccl.lib, "ncclAllReduce");
  *(void**)&g_nccl.CommDestroy = dlsym(g_nccl.lib, "ncclCommDestroy");
  *(void**)&g_nccl.GetErrorString = dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return fail(JXB_ERR_NCCL, "libnccl is missing required symbols");
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// engine / model objects
// ---------------------------------------------------------------------------------------
struct jxb_engine {
  int device = 0;
  int sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  // caching allocator: device blocks of destroyed models are kept and handed to the next model of
  // a similar shape (cudaMalloc / cudaFree of ~0.5 GB of columns costs milliseconds per model)
  std::vector<std::pair<void*, size_t>> pool;
  size_t pool_bytes = 0;
  // peer-memory exchange (CUDA IPC over NVLink): own buffer + every rank's buffer as mapped here
  XchgBuf* xlocal = nullptr;
  XchgBuf* xpeer[kMaxPeers] = {};
  bool p2p = false;
  const void* hist_owner = nullptr;     // the sharded economy model currently counting into xlocal's histogram region
  // receive areas of destroyed sharded models: an allocation whose IPC handle was exported must not be freed
  // while a peer process may still have it mapped, and the peers close their mappings whenever THEY destroy
  // their model -- so the areas are only released with the engine; until then they are handed to the next
  // sharded model that needs an area of the same size (take_retired_area)
  std::vector<std::pair<void*, size_t>> retired_areas;
};

struct jxb_model {
  jxb_engine* eng = nullptr;
  jxb_model_desc desc{};
  const ProgramSpec* prog = nullptr;
  const RuleSpec* rules[JXB_MAX_TYPES] = {};
  ModelDev dev{};
  std::vector<std::pair<void*, size_t>> allocs;
  std::vector<std::pair<void*, size_t>> net_allocs;     // CSR + SIR buffers of the current network (released on rebuild)
  Key rng{0, 0};
  bool initialized = false;
  bool collections_ready[JXB_MAX_TYPES] = {};
  long long time_step = 0;
  // per-run buffers
  uint32_t* d_keys = nullptr; std::vector<uint32_t> h_keys; size_t keys_cap = 0;
  double* d_metrics = nullptr; int* d_rec = nullptr; size_t rec_cap = 0;
  // Schelling
  bool has_grid = false; SchellingDev sd{}; long long pad = 0; float* d_ratio = nullptr;
  bool grid_built = false; bool sat_dirty = false; long long n_empty_cells = 0; int sch_blocks = 0;
  // bit-sliced variant (row length a multiple of 1024): the planes are the live grid during a run
  bool sch_bits = false; SchellingBitsDev sb{}; bool ct_stale = false;
  // persistent bit-sliced kernel: the agent id and its move count travel with the cell (sb.cell_am); the
  // 'position' / 'moves' columns are derived from it when they are read (pos_stale = they are behind)
  bool sch_packed = false; bool pos_stale = false;
  // row-band decomposition over the GPUs of a box (csrc/grid_shard.cuh): receive area + peers' areas
  bool grid_sharded = false; bool gs_attached = false; GridShardDev gs{}; unsigned char* gs_area = nullptr;
  size_t gs_area_bytes = 0; void* gs_opened[kMaxPeers] = {};
  // SIR
  bool has_net = false; SirDev sv{}; bool net_built = false; long long nnz = 0;
  // node-range sharding of the network over ranks (csrc/sir.cuh): IPC-shared area [hdr | bitmap 0 | bitmap 1]
  bool net_sharded = false; bool ns_attached = false; unsigned char* ns_area = nullptr; size_t ns_area_bytes = 0;
  void* ns_opened[kMaxPeers] = {};
  // SIR formulation (JXB_SIR_MODE): 0 "pull" = CSR ballot-segmented sweep of ALL edges, fused
  // transitions; 1 "push" = infected rows scatter-add into k32 + transition kernel; 2 "pull_s" = pull
  // over the susceptible rows only; 3 "auto" (default) = direction-optimising, the step's tail picks
  // push or pull_s for the next step on the device
  int sir_mode = 3; int sir_tblocks = 0; int sir_pull_cps = 8;
  // traced model (jxb_model_create_traced): layout owned by the model, kernels in a generated library
  bool traced = false; RuleSpec traced_rules[JXB_MAX_TYPES]; ProgramSpec traced_prog;
  std::vector<std::string> traced_names; int traced_acc = 0, traced_variants = 1; bool traced_started = false;
  int (*traced_init)(const void*, int, unsigned int, unsigned int, int, int, void*) = nullptr;
  int (*traced_step)(const void*, int, int, void*) = nullptr;
  // economy (C4-B)
  bool has_eco = false; EcoDev eco{}; int eco_hh = -1;
  // graphs: cached executable graphs of `chunk` consecutive steps
  cudaGraphExec_t graph1 = nullptr, graphK = nullptr; int chunkK = 0;
  const void* sig_keys = nullptr; const void* sig_metrics = nullptr; const void* sig_rec = nullptr; int sig_ci = 0;
  // per-agent series (csrc/record.cuh): recorded columns, their device ring and its layout
  struct RecField { int type, field; size_t bytes, stride; size_t off; };
  std::vector<RecField> rec_fields; unsigned char* d_series = nullptr; size_t series_bytes = 0, series_slots = 0;
  int series_nrec = 0; const void* sig_series = nullptr; size_t sig_series_slots = 0;
  // filter scratch (jxb_collection_filter_select -> _gather): exclusive scan of the selection flags
  unsigned int* filter_pos = nullptr; size_t filter_pos_bytes = 0; int filter_type = -1; long long filter_count = -1;
  // profiling of the dominant kernel
  bool profile = false; double prof_seconds = 0; int64_t prof_launches = 0;
  std::vector<cudaEvent_t> prof_events;
  std::vector<cudaEvent_t> gs_trace;      // JXB_GS_TRACE=1 (with JXB_NO_GRAPH=1): events around each band kernel of each step
  int step_blocks = 0;
};

static size_t pool_cap_bytes() {
  static const size_t cap = [] {
    const char* ev = getenv("JXB_POOL_MB");
    return (size_t)(ev ? atoll(ev) : 16384) << 20;
  }();
  return cap;
}

static cudaError_t pool_alloc(jxb_engine* eng, void** out, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
  int best = -1;
  for (int i = 0; i < (int)eng->pool.size(); ++i) {
    const size_t b = eng->pool[i].second;
    if (b >= bytes && b <= bytes + bytes / 8 + (1u << 16) && (best < 0 || b < eng->pool[best].second)) best = i;
  }
  if (best >= 0) {
    *out = eng->pool[best].first;
    eng->pool_bytes -= eng->pool[best].second;
    eng->pool.erase(eng->pool.begin() + best);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess && !eng->pool.empty()) {          // give the cached blocks back and retry
    cudaGetLastError();
    for (auto& b : eng->pool) cudaFree(b.first);
    eng->pool.clear();
    eng->pool_bytes = 0;
    e = cudaMalloc(out, bytes);
  }
  return e;
}

static void pool_free(jxb_engine* eng, void* p, size_t bytes) {
  if (!p) return;
  bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
  eng->pool.emplace_back(p, bytes);
  eng->pool_bytes += bytes;
  while (eng->pool_bytes > pool_cap_bytes() && !eng->pool.empty()) {
    cudaFree(eng->pool.front().first);
    eng->pool_bytes -= eng->pool.front().second;
    eng->pool.erase(eng->pool.begin());
  }
}

template <class T>
static int dev_alloc(jxb_model* m, T** p, size_t count, std::vector<std::pair<void*, size_t>>* list = nullptr) {
  void* q = nullptr;
  const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
  CK(pool_alloc(m->eng, &q, bytes));
  (list ? *list : m->allocs).emplace_back(q, bytes);
  *p = (T*)q;
  return JXB_OK;
}

// receive areas whose IPC handle was exported: recycled by exact size for the next sharded model of the same
// shape (bench loops, sweeps and test suites create and destroy many), never freed before the engine goes
static void* take_retired_area(jxb_engine* eng, size_t bytes) {
  for (size_t i = 0; i < eng->retired_areas.size(); ++i)
    if (eng->retired_areas[i].second == bytes) {
      void* p = eng->retired_areas[i].first;
      eng->retired_areas.erase(eng->retired_areas.begin() + i);
      return p;
    }
  return nullptr;
}

extern "C" int jxb_engine_create(int device, jxb_engine** out) {
  if (!out) return fail(JXB_ERR_INVALID, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(JXB_ERR_NO_DEVICE, "no CUDA device visible (%s); libjxb has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(JXB_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  CK(cudaSetDevice(device));
  jxb_engine* eng = new jxb_engine();
  eng->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  eng->sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&eng->ev0));
  CK(cudaEventCreate(&eng->ev1));
  *out = eng;
  return JXB_OK;
}

extern "C" int jxb_engine_destroy(jxb_engine* eng) {
  if (!eng) return JXB_OK;
  cudaSetDevice(eng->device);
  if (eng->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(eng->nccl_comm);
  for (int p = 0; p < kMaxPeers; ++p)
    if (eng->xpeer[p] && eng->xpeer[p] != eng->xlocal) cudaIpcCloseMemHandle(eng->xpeer[p]);
  if (eng->xlocal) cudaFree(eng->xlocal);
  for (auto& a : eng->retired_areas) cudaFree(a.first);
  for (auto& b : eng->pool) cudaFree(b.first);
  cudaEventDestroy(eng->ev0);
  cudaEventDestroy(eng->ev1);
  cudaStreamDestroy(eng->stream);
  delete eng;
  return JXB_OK;
}

extern "C" int jxb_engine_sm_count(jxb_engine* eng, int* out) {
  if (!eng || !out) return fail(JXB_ERR_INVALID, "null argument");
  *out = eng->sms;
  return JXB_OK;
}

extern "C" int jxb_engine_launch_count(jxb_engine* eng, int64_t* out) {
  if (!eng || !out) return fail(JXB_ERR_INVALID, "null argument");
  *out = eng->launches;
  return JXB_OK;
}

extern "C" int jxb_nccl_unique_id(void* id, size_t bytes) {
  if (bytes < sizeof(NcclId)) return fail(JXB_ERR_INVALID, "need %zu bytes", sizeof(NcclId));
  int rc = load_nccl();
  if (rc) return rc;
  int r = g_nccl.GetUniqueId(id);
  if (r) return fail(JXB_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return JXB_OK;
}

extern "C" int jxb_engine_attach_nccl(jxb_engine* eng, const void* id, size_t bytes, int rank, int world) {
  if (!eng || !id || bytes < sizeof(NcclId)) return fail(JXB_ERR_INVALID, "bad nccl id");
  int rc = load_nccl();
  if (rc) return rc;
  CK(cudaSetDevice(eng->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  int r = g_nccl.CommInitRank(&eng->nccl_comm, world, nid, rank);
  if (r) return fail(JXB_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  eng->rank = rank;
  eng->world = world;
  return JXB_OK;
}

// peer-memory exchange: every rank exports the IPC handle of its XchgBuf, the host shim
// all-gathers the handles (torch.distributed) and every rank maps its peers' buffers.
extern "C" int jxb_engine_p2p_export(jxb_engine* eng, void* handle_out, size_t bytes) {
  if (!eng || !handle_out || bytes < sizeof(cudaIpcMemHandle_t))
    return fail(JXB_ERR_INVALID, "need %zu bytes for the IPC handle", sizeof(cudaIpcMemHandle_t));
  CK(cudaSetDevice(eng->device));
  if (!eng->xlocal) {
    // [XchgBuf | pad | income histogram of the sharded economy's Gini] -- one allocation, one IPC handle
    static_assert(sizeof(XchgBuf) <= kXchgHistOffset, "XchgBuf must fit in front of the histogram region");
    const size_t bytes = kXchgHistOffset + (size_t)kGiniBins * sizeof(unsigned int);
    CK(cudaMalloc((void**)&eng->xlocal, bytes));
    CK(cudaMemset(eng->xlocal, 0, bytes));
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, eng->xlocal));
  memset(handle_out, 0, bytes);
  memcpy(handle_out, &h, sizeof(h));
  return JXB_OK;
}

extern "C" int jxb_engine_p2p_attach(jxb_engine* eng, const void* handles, size_t bytes_each, int rank, int world) {
  if (!eng || !handles || bytes_each < sizeof(cudaIpcMemHandle_t)) return fail(JXB_ERR_INVALID, "bad IPC handle table");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(JXB_ERR_INVALID, "rank %d / world %d out of range (max %d peers)", rank, world, kMaxPeers);
  if (!eng->xlocal) return fail(JXB_ERR_STATE, "call jxb_engine_p2p_export first");
  CK(cudaSetDevice(eng->device));
  for (int p = 0; p < world; ++p) {
    if (p == rank) { eng->xpeer[p] = eng->xlocal; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * bytes_each, sizeof(h));
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    eng->xpeer[p] = (XchgBuf*)q;
  }
  eng->rank = rank;
  eng->world = world;
  eng->p2p = true;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// model construction
// ---------------------------------------------------------------------------------------
static bool program_accepts(int program, int rule) {
  switch (program) {
    case JXB_PROGRAM_NONE: return true;    // any registered layout as a passive container (AgentCollection.filter results);
                                           // rules that need a grid / network / economy program do nothing under it
    case JXB_PROGRAM_RANDOM_WALK: return rule == JXB_RULE_RANDOM_WALKER || rule == JXB_RULE_SCALED_WALKER;
    case JXB_PROGRAM_MARKET: return rule == JXB_RULE_CONSUMER || rule == JXB_RULE_PRODUCER;
    case JXB_PROGRAM_GROWTH: return rule == JXB_RULE_GROWTH;
    case JXB_PROGRAM_COUNTER: return rule == JXB_RULE_INCREMENT || rule == JXB_RULE_GROWTH;
    case JXB_PROGRAM_SCHELLING: return rule == JXB_RULE_SCHELLING;
    case JXB_PROGRAM_SIR: return rule == JXB_RULE_SIR;
    case JXB_PROGRAM_ECONOMY: return rule == JXB_RULE_HOUSEHOLD || rule == JXB_RULE_CONSUMER_FIRM;
  }
  return false;
}

static int validate_desc(const jxb_model_desc* d, bool traced = false) {
  if (!d) return fail(JXB_ERR_INVALID, "desc is NULL");
  if (traced != (d->program == JXB_PROGRAM_TRACED))
    return fail(JXB_ERR_INVALID, "JXB_PROGRAM_TRACED models are created with jxb_model_create_traced (and only those)");
  if (!traced && !find_program(d->program)) return fail(JXB_ERR_INVALID, "unknown program %d", d->program);
  if (d->rng_mode != JXB_RNG_LEGACY && d->rng_mode != JXB_RNG_PARTITIONABLE)
    return fail(JXB_ERR_INVALID, "unknown rng_mode %d", d->rng_mode);
  if (d->n_types < 1 || d->n_types > JXB_MAX_TYPES)
    return fail(JXB_ERR_INVALID, "No agent collections added to model");   // model.py:125-126
  for (int i = 0; i < d->n_types; ++i) {
    const jxb_type_desc& t = d->types[i];
    if (traced ? t.rule != JXB_RULE_TRACED : !find_rule(t.rule))
      return fail(JXB_ERR_INVALID, "collection %d: rule %d is not a registered agent rule", i, t.rule);
    if (t.n_agents <= 0) return fail(JXB_ERR_INVALID, "num_agents must be a positive integer");
    if (!traced && !program_accepts(d->program, t.rule))
      return fail(JXB_ERR_INVALID, "program %d does not accept rule %d", d->program, t.rule);
    if (t.global_n < t.n_agents || t.global_offset < 0 || t.global_offset + t.n_agents > t.global_n)
      return fail(JXB_ERR_INVALID, "collection %d: bad shard [%lld,+%lld) of %lld", i,
                  (long long)t.global_offset, (long long)t.n_agents, (long long)t.global_n);
  }
  if ((d->program == JXB_PROGRAM_SCHELLING || d->program == JXB_PROGRAM_SIR) && d->n_types != 1)
    return fail(JXB_ERR_INVALID, "grid / network programs take exactly one collection");
  if (d->program == JXB_PROGRAM_SCHELLING) {
    if (d->grid_w <= 0 || d->grid_h <= 0) return fail(JXB_ERR_INVALID, "Schelling needs a Grid shape");
    if ((long long)d->grid_w * d->grid_h >= (1ll << 31)) return fail(JXB_ERR_UNSUPPORTED, "grid too large");
    if (d->types[0].n_agents > (long long)d->grid_w * d->grid_h)
      return fail(JXB_ERR_INVALID, "more agents than cells");
  }
  return JXB_OK;
}

static void fill_type_dev(const jxb_type_desc& t, TypeDev& td) {
  td.n = t.n_agents;
  td.goff = t.global_offset;
  td.gn = t.global_n;
  td.rule = t.rule;
  for (int k = 0; k < JXB_MAX_PARAMS; ++k) td.p[k] = k < t.n_params ? t.params[k] : 0.f;
}

static int plan_step_blocks(jxb_model* m);
static int grid_shard_prepare(jxb_model* m, int row_begin, int row_end);

static int model_create(jxb_engine* eng, const jxb_model_desc* d, const jxb_traced_spec* ts, jxb_model** out);

extern "C" int jxb_model_create(jxb_engine* eng, const jxb_model_desc* d, jxb_model** out) {
  return model_create(eng, d, nullptr, out);
}

extern "C" int jxb_model_create_traced(jxb_engine* eng, const jxb_model_desc* d, const jxb_traced_spec* ts, jxb_model** out) {
  if (!ts || !ts->launch_init || !ts->launch_step) return fail(JXB_ERR_INVALID, "traced spec / launchers missing");
  if (ts->n_env < 0 || ts->n_env > kMaxEnv || ts->n_metrics < 0 || ts->n_metrics > kMaxMetrics || ts->n_acc < 1 ||
      ts->n_acc > 64 || ts->n_variants < 1)
    return fail(JXB_ERR_INVALID, "traced spec out of range (<= %d env entries, <= %d metrics, <= 64 reductions)", kMaxEnv, kMaxMetrics);
  return model_create(eng, d, ts, out);
}

static int model_create(jxb_engine* eng, const jxb_model_desc* d, const jxb_traced_spec* ts, jxb_model** out) {
  if (!eng || !out) return fail(JXB_ERR_INVALID, "null argument");
  int rc = validate_desc(d, ts != nullptr);
  if (rc) return rc;
  CK(cudaSetDevice(eng->device));
  jxb_model* m = new jxb_model();
  m->eng = eng;
  m->desc = *d;
  m->prog = ts ? nullptr : find_program(d->program);
  if (ts) {
    // the model owns its layout tables (names are copied: the caller's strings need not outlive the call)
    m->traced = true;
    m->traced_names.reserve(JXB_MAX_TYPES * JXB_MAX_FIELDS + kMaxEnv + kMaxMetrics);
    auto keep = [&](const char* sname) { m->traced_names.emplace_back(sname ? sname : ""); return m->traced_names.back().c_str(); };
    for (int i = 0; i < d->n_types; ++i) {
      RuleSpec& r = m->traced_rules[i];
      r.rule = JXB_RULE_TRACED;
      r.nf = ts->n_fields[i];
      if (r.nf < 1 || r.nf > kMaxFields) { delete m; return fail(JXB_ERR_INVALID, "collection %d: 1..%d fields", i, kMaxFields); }
      for (int f = 0; f < r.nf; ++f) {
        const int width = ts->field_widths[i][f];
        if (width < 1 || width > 4) { delete m; return fail(JXB_ERR_INVALID, "collection %d field %d: width 1..4", i, f); }
        r.f[f] = FieldSpec{keep(ts->field_names[i][f]), ts->field_dtypes[i][f], width};
      }
    }
    ProgramSpec& p = m->traced_prog;
    p.program = JXB_PROGRAM_TRACED; p.has_env_fn = ts->has_env_fn; p.n_env = ts->n_env; p.n_metrics = ts->n_metrics;
    for (int k = 0; k < ts->n_env; ++k) p.env[k] = SlotSpec{keep(ts->env_names[k]), ts->env_dtypes[k], ts->env_init[k]};
    for (int k = 0; k < ts->n_metrics; ++k) p.metrics[k] = SlotSpec{keep(ts->metric_names[k]), ts->metric_dtypes[k], 0.0};
    m->prog = &p;
    m->traced_acc = ts->n_acc;
    m->traced_variants = ts->n_variants;
    *(void**)&m->traced_init = ts->launch_init;
    *(void**)&m->traced_step = ts->launch_step;
    if (ts->n_consts < 0 || ts->n_consts > 4096 || (ts->n_consts && !ts->consts)) { delete m; return fail(JXB_ERR_INVALID, "bad constant table"); }
  }
  ModelDev& md = m->dev;
  memset(&md, 0, sizeof(md));
  md.n_types = d->n_types;
  md.program = d->program;
  md.collect_interval = 1;
  md.has_env_fn = m->prog->has_env_fn;
  md.world_size = d->world_size > 1 ? d->world_size : 1;
  md.rank = d->rank;
  md.exchange = 0;
  if (md.world_size > 1 && d->program == JXB_PROGRAM_SIR) {
    const jxb_type_desc& t0 = d->types[0];
    if (md.world_size > kMaxPeers || md.rank < 0 || md.rank >= md.world_size || (t0.global_offset % 32) != 0 ||
        (t0.global_offset + t0.n_agents != t0.global_n && (t0.n_agents % 32) != 0)) {
      delete m;
      return fail(JXB_ERR_INVALID, "a sharded Network splits the agents at multiples of 32 over at most %d ranks", kMaxPeers);
    }
    m->net_sharded = true;
  }
  // a sharded grid / network has its own receive areas
  if (md.world_size > 1 && d->program != JXB_PROGRAM_SCHELLING && d->program != JXB_PROGRAM_SIR) {
    static const bool force_nccl = getenv("JXB_EXCHANGE") && !strcmp(getenv("JXB_EXCHANGE"), "nccl");
    if (eng->p2p && !force_nccl) {
      if (eng->world != md.world_size || eng->rank != md.rank) {
        delete m;
        return fail(JXB_ERR_INVALID, "model shard (rank %d of %d) does not match the attached peers (rank %d of %d)",
                    md.rank, md.world_size, eng->rank, eng->world);
      }
      md.exchange = 1;
      for (int p = 0; p < md.world_size; ++p) md.xpeer[p] = eng->xpeer[p];
    } else if (eng->nccl_comm) {
      md.exchange = 2;
    } else {
      delete m;
      return fail(JXB_ERR_STATE, "sharded model but neither peer memory (jxb_engine_p2p_attach) nor an NCCL "
                                 "communicator (jxb_engine_attach_nccl) is attached");
    }
  }
  for (int k = 0; k < JXB_MAX_PARAMS; ++k) md.mp[k] = k < d->n_params ? d->params[k] : 0.0;
#define TRY(x) do { rc = (x); if (rc) { jxb_model_destroy(m); return rc; } } while (0)
  for (int i = 0; i < d->n_types; ++i) {
    const jxb_type_desc& t = d->types[i];
    m->rules[i] = ts ? &m->traced_rules[i] : find_rule(t.rule);
    fill_type_dev(t, md.t[i]);
    for (int f = 0; f < m->rules[i]->nf; ++f) {
      const FieldSpec& fs = m->rules[i]->f[f];
      unsigned char* p = nullptr;
      // round up so vector tails of the 4-wide loops never fault
      size_t bytes = (size_t)(t.n_agents + 8) * fs.width * dtype_size(fs.dtype);
      TRY(dev_alloc(m, &p, bytes));
      cudaMemsetAsync(p, 0, bytes, eng->stream);
      md.t[i].f[f] = p;
    }
  }
  TRY(dev_alloc(m, &md.env, kMaxEnv));
  {
    double h[kMaxEnv] = {0};
    for (int i = 0; i < m->prog->n_env; ++i) h[i] = m->prog->env[i].dflt;
    if (cudaMemcpy(md.env, h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) {
      jxb_model_destroy(m);
      return fail(JXB_ERR_CUDA, "env upload failed");
    }
  }
  TRY(dev_alloc(m, &md.ctrl, 1));
  cudaMemset(md.ctrl, 0, sizeof(Ctrl));
  if (ts) {
    double* d_c = nullptr;
    TRY(dev_alloc(m, &d_c, (size_t)std::max(ts->n_consts, 1)));
    if (ts->n_consts && cudaMemcpy(d_c, ts->consts, (size_t)ts->n_consts * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      jxb_model_destroy(m);
      return fail(JXB_ERR_CUDA, "constant table upload failed");
    }
    md.consts = d_c;
  }
  TRY(dev_alloc(m, &md.allreduce_buf, kAcc));
  TRY(plan_step_blocks(m));
  TRY(dev_alloc(m, &md.partials, (size_t)std::max(m->step_blocks, 1) * std::max(kAcc, m->traced_acc)));

  if (d->program == JXB_PROGRAM_SCHELLING) {
    SchellingDev& sd = m->sd;
    sd.W = d->grid_w; sd.H = d->grid_h; sd.periodic = d->grid_periodic ? 1 : 0;
    sd.cells = (long long)sd.W * sd.H;
    sd.ntiles = (int)((sd.cells + kTileCells - 1) / kTileCells);
    m->pad = ((long long)sd.H + 16 + 15) / 16 * 16;
    signed char* base = nullptr;
    TRY(dev_alloc(m, &base, (size_t)(sd.cells + 2 * m->pad + 16)));
    sd.ct = base + m->pad;
    TRY(dev_alloc(m, &sd.cell_agent, (size_t)sd.cells));
    const long long n = d->types[0].n_agents;
    m->n_empty_cells = sd.cells - n;
    sd.n_empty = (unsigned int)(sd.cells - n);
    TRY(dev_alloc(m, &sd.U, (size_t)n + 1));
    TRY(dev_alloc(m, &sd.MA, (size_t)n + 1));
    TRY(dev_alloc(m, &sd.E, (size_t)(sd.cells - n) + 1));
    TRY(dev_alloc(m, &sd.mask16, (size_t)(sd.cells + 15) / 16 + 16));
    {
      // persistent cooperative grid: every CTA must be co-resident
      int occ = 0;
      cudaError_t oe = (sd.H % 16 == 0)
          ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, schelling_run_kernel<true, 1>, kThreads, 0)
          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, schelling_run_kernel<false, 1>, kThreads, 0);
      if (oe != cudaSuccess || occ < 1) { jxb_model_destroy(m); return fail(JXB_ERR_CUDA, "occupancy query failed"); }
      int bps = occ;
      if (const char* ev = getenv("JXB_SCH_BPS")) bps = std::max(1, std::min(occ, atoi(ev)));
      else bps = std::min(occ, 4);
      m->sch_blocks = (int)std::max<long long>(1, std::min<long long>((long long)eng->sms * bps, (sd.cells + 511) / 512));
      const int strips = sd.H / 1024;
      static const bool no_bits = getenv("JXB_SCH_NO_BITS") != nullptr;
      if (!no_bits && sd.H % 1024 == 0 && (strips == 1 || strips == 2 || strips == 4 || strips == 8)) {
        int occb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, schelling_bits_kernel<1>, kThreads, 0) == cudaSuccess && occb >= 1) {
          m->sch_bits = true;
          int bpsb = std::min(occb, 4);
          if (const char* ev = getenv("JXB_SCH_BPS")) bpsb = std::max(1, std::min(occb, atoi(ev)));
          // at least ~2 strip-rows per warp, all CTAs co-resident
          m->sch_blocks = (int)std::max<long long>(1, std::min<long long>((long long)eng->sms * bpsb,
                                                                          ((long long)sd.W * strips + 15) / 16));
        } else {
          cudaGetLastError();
        }
      }
    }
    if (md.world_size > 1) {
      // row bands over the ranks: always the bit-plane layout (any row length that fills whole words)
      if (sd.H % 32 != 0) { jxb_model_destroy(m); return fail(JXB_ERR_UNSUPPORTED, "a sharded Grid needs a row length that is a multiple of 32 cells"); }
      if (md.world_size > kMaxPeers || md.rank < 0 || md.rank >= md.world_size || sd.W < md.world_size) {
        jxb_model_destroy(m);
        return fail(JXB_ERR_INVALID, "cannot split %d grid rows over rank %d of %d (at most %d ranks)", sd.W, md.rank, md.world_size, kMaxPeers);
      }
      m->sch_bits = true;
      m->grid_sharded = true;
    } else if (sd.H % 32 == 0) {
      // ONE GPU: the band kernels (4 launches per step, whole grid = one band) also serve the large
      // grids whose rows are outside the persistent bit-sliced kernel's shapes (not a multiple of 1024
      // cells, or more than 8192) -- ~2x the byte/LUT kernel there; small grids stay on the persistent
      // kernels (one launch for the whole run).  JXB_GRID_BANDS=1 / 0 forces / forbids it.
      const char* ev = getenv("JXB_GRID_BANDS");
      const bool want = ev ? atoi(ev) != 0 : (!m->sch_bits && sd.cells >= (1ll << 24));
      if (want) { m->sch_bits = true; m->grid_sharded = true; }
    }
    if (m->sch_bits) {
      SchellingBitsDev& sb = m->sb;
      sb.wpr = sd.H / 32;
      sb.spr = sd.H / 1024;
      const size_t plane_words = (size_t)(sd.W + 2) * sb.wpr + 32;
      unsigned int* p0 = nullptr; unsigned int* p1 = nullptr;
      TRY(dev_alloc(m, &p0, plane_words));
      TRY(dev_alloc(m, &p1, plane_words));
      cudaMemsetAsync(p0, 0, plane_words * 4, eng->stream);
      cudaMemsetAsync(p1, 0, plane_words * 4, eng->stream);
      sb.occ = p0 + sb.wpr;
      sb.t1 = p1 + sb.wpr;
      TRY(dev_alloc(m, &sb.umask, (size_t)(sd.cells >> 5) + 32));
      TRY(dev_alloc(m, &sb.cell_am, (size_t)sd.cells));
      m->sch_packed = true;
    }
    {
      BlkPart* bp = nullptr;
      TRY(dev_alloc(m, &bp, (size_t)2 * m->sch_blocks));
      sd.blk_part = bp;
    }
    // least number of same-type neighbours that satisfies an agent with o occupied neighbours,
    // evaluated in the rule's float32 arithmetic: same/o >= threshold
    const float thr = d->n_params > 0 ? (float)d->params[0] : 0.5f;
    sd.need_lut = 0;
    for (int o = 1; o <= 8; ++o) {
      int need = o + 1;
      for (int sm = 0; sm <= o; ++sm)
        if (((float)sm / (float)o) >= thr) { need = sm; break; }
      sd.need_lut |= (unsigned long long)need << (4 * o);
      for (int q = 0; q < 4; ++q) m->sb.need_sel[o][q] = ((need >> q) & 1) ? 0xFFFFFFFFu : 0u;
    }
    m->has_grid = true;
    if (m->grid_sharded && md.world_size == 1) {     // single band: no peers to map
      TRY(grid_shard_prepare(m, 0, sd.W));
      m->gs_attached = true;
    }
  }
  if (d->program == JXB_PROGRAM_SIR) m->has_net = true;
  if (d->program == JXB_PROGRAM_ECONOMY) {
    if (md.world_size > 1 && md.exchange != 1) {
      jxb_model_destroy(m);
      return fail(JXB_ERR_STATE, "a sharded economy runs on the peer-memory exchange (env partial sums and the ranks' income "
                                 "histograms); attach it with jxb_engine_p2p_export / jxb_engine_p2p_attach");
    }
    if (md.world_size > 1 && eng->hist_owner) {
      jxb_model_destroy(m);
      return fail(JXB_ERR_STATE, "one sharded economy model at a time per engine (its income histogram lives in the engine's "
                                 "exchange allocation); destroy the previous one first");
    }
    EcoDev& ed = m->eco;
    TRY(dev_alloc(m, &ed.partials, (size_t)std::max(m->step_blocks, 1) * kEcoRow));
    TRY(dev_alloc(m, &ed.bin_count, (size_t)kGiniBins));
    TRY(dev_alloc(m, &ed.tile_range, 4));
    {
      const int empty[4] = {1, 0, -1, 0};
      cudaMemcpyAsync(ed.tile_range, empty, sizeof(empty), cudaMemcpyHostToDevice, eng->stream);
    }
    if (md.world_size > 1) {
      ed.hist = (unsigned int*)((unsigned char*)eng->xlocal + kXchgHistOffset);
      cudaMemsetAsync(ed.hist, 0, (size_t)kGiniBins * 4, eng->stream);
      eng->hist_owner = m;
    } else {
      ed.hist = ed.bin_count;
    }
    TRY(dev_alloc(m, &ed.bin_base, (size_t)kGiniBins));
    TRY(dev_alloc(m, &ed.scan_sums, (size_t)kGiniBins / kGiniScanTile));
    TRY(dev_alloc(m, &ed.ticket2, 4));
    cudaMemsetAsync(ed.bin_count, 0, (size_t)kGiniBins * 4, eng->stream);
    cudaMemsetAsync(ed.ticket2, 0, 16, eng->stream);
    for (int i = 0; i < d->n_types; ++i)
      if (d->types[i].rule == JXB_RULE_HOUSEHOLD) m->eco_hh = i;
    const long long nh = m->eco_hh >= 0 ? d->types[m->eco_hh].n_agents : 0;
    ed.gini_blocks = (int)std::max<long long>(1, std::min<long long>((nh + kThreads - 1) / kThreads, (long long)eng->sms * 8));
    TRY(dev_alloc(m, &ed.gini_partials, (size_t)ed.gini_blocks * 2));
    m->has_eco = true;
  }
#undef TRY
  cudaStreamSynchronize(eng->stream);
  *out = m;
  return JXB_OK;
}

extern "C" int jxb_model_destroy(jxb_model* m) {
  if (!m) return JXB_OK;
  cudaSetDevice(m->eng->device);
  cudaStreamSynchronize(m->eng->stream);
  if (m->graph1) cudaGraphExecDestroy(m->graph1);
  if (m->graphK) cudaGraphExecDestroy(m->graphK);
  for (auto e : m->prof_events) cudaEventDestroy(e);
  for (auto& a : m->allocs) pool_free(m->eng, a.first, a.second);
  for (int p = 0; p < kMaxPeers; ++p)
    if (m->gs_opened[p]) cudaIpcCloseMemHandle(m->gs_opened[p]);
  for (int p = 0; p < kMaxPeers; ++p)
    if (m->ns_opened[p]) cudaIpcCloseMemHandle(m->ns_opened[p]);
  // exported areas outlive the model (see jxb_engine::retired_areas); a single-band grid never exported its area
  if (m->gs_area) { if (m->dev.world_size > 1) m->eng->retired_areas.emplace_back(m->gs_area, m->gs_area_bytes); else cudaFree(m->gs_area); }
  if (m->ns_area) m->eng->retired_areas.emplace_back(m->ns_area, m->ns_area_bytes);
  for (auto& a : m->net_allocs) pool_free(m->eng, a.first, a.second);
  if (m->eng->hist_owner == m) m->eng->hist_owner = nullptr;
  pool_free(m->eng, m->d_series, m->series_bytes);
  pool_free(m->eng, m->filter_pos, m->filter_pos_bytes);
  pool_free(m->eng, m->d_keys, m->keys_cap * 4);
  pool_free(m->eng, m->d_metrics, m->rec_cap * kMaxMetrics * sizeof(double));
  pool_free(m->eng, m->d_rec, m->rec_cap * sizeof(int));
  delete m;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------
#define NEED(m) if (!(m)) return fail(JXB_ERR_INVALID, "model is NULL")
#define NEED_TYPE(m, type) \
  if ((type) < 0 || (type) >= (m)->desc.n_types) return fail(JXB_ERR_INVALID, "type index %d out of range", (type))

extern "C" int jxb_model_n_fields(jxb_model* m, int type, int* out) {
  NEED(m); NEED_TYPE(m, type);
  *out = m->rules[type]->nf;
  return JXB_OK;
}
extern "C" int jxb_model_field_info(jxb_model* m, int type, int field, const char** name, int* dtype, int* width) {
  NEED(m); NEED_TYPE(m, type);
  if (field < 0 || field >= m->rules[type]->nf) return fail(JXB_ERR_INVALID, "field %d out of range", field);
  const FieldSpec& f = m->rules[type]->f[field];
  if (name) *name = f.name;
  if (dtype) *dtype = f.dtype;
  if (width) *width = f.width;
  return JXB_OK;
}
extern "C" int jxb_model_n_env(jxb_model* m, int* out) { NEED(m); *out = m->prog->n_env; return JXB_OK; }
extern "C" int jxb_model_env_info(jxb_model* m, int slot, const char** name, int* dtype) {
  NEED(m);
  if (slot < 0 || slot >= m->prog->n_env) return fail(JXB_ERR_INVALID, "env slot %d out of range", slot);
  if (name) *name = m->prog->env[slot].name;
  if (dtype) *dtype = m->prog->env[slot].dtype;
  return JXB_OK;
}
extern "C" int jxb_model_n_metrics(jxb_model* m, int* out) { NEED(m); *out = m->prog->n_metrics; return JXB_OK; }
extern "C" int jxb_model_metric_info(jxb_model* m, int k, const char** name, int* dtype) {
  NEED(m);
  if (k < 0 || k >= m->prog->n_metrics) return fail(JXB_ERR_INVALID, "metric %d out of range", k);
  if (name) *name = m->prog->metrics[k].name;
  if (dtype) *dtype = m->prog->metrics[k].dtype;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// state access
// ---------------------------------------------------------------------------------------
static size_t field_bytes(jxb_model* m, int type, int field) {
  const FieldSpec& f = m->rules[type]->f[field];
  return (size_t)m->desc.types[type].n_agents * f.width * dtype_size(f.dtype);
}

static int sir_sync_from_api(jxb_model* m);
static int sir_sync_to_api(jxb_model* m);
static int materialize_field(jxb_model* m, int type, int field, cudaStream_t s, bool in_step);
static int schelling_export_satisfied(jxb_model* m, cudaStream_t s, bool clear_dirty);

static int check_field(jxb_model* m, int type, int field, size_t bytes) {
  NEED(m); NEED_TYPE(m, type);
  if (field < 0 || field >= m->rules[type]->nf) return fail(JXB_ERR_INVALID, "field %d out of range", field);
  if (bytes != field_bytes(m, type, field))
    return fail(JXB_ERR_INVALID, "field '%s': expected %zu bytes, got %zu", m->rules[type]->f[field].name,
                field_bytes(m, type, field), bytes);
  return JXB_OK;
}

// ONE grid over several ranks: the API columns of a rank only hold its band's view after a read
static inline bool grid_multi_rank(const jxb_model* m) { return m->grid_sharded && m->dev.world_size > 1; }

// a Schelling state column was written by the caller (upload / fill)
static int grid_after_write(jxb_model* m, int field) {
  if (!m->has_grid) return JXB_OK;
  if (field == 0 || field == 1) m->grid_built = false;     // cell binning is rebuilt (slot order restarts)
  if (field == 2) m->sat_dirty = false;
  if (field == 3 && m->sch_packed && m->grid_built) {
    // the move counts travel with the cells: refresh them from the column, keep the grid and the slot order
    cell_am_set_moves_kernel<<<m->eng->sms * 8, 256, 0, m->eng->stream>>>(m->sd, m->sb, (const int*)m->dev.t[0].f[3]);
    m->eng->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->eng->stream));
    if (grid_multi_rank(m)) m->pos_stale = true;      // the column is whole again on every rank: reads re-derive the band view
  }
  return JXB_OK;
}

extern "C" int jxb_model_upload(jxb_model* m, int type, int field, const void* host, size_t bytes) {
  int rc = check_field(m, type, field, bytes);
  if (rc) return rc;
  CK(cudaSetDevice(m->eng->device));
  if (m->sch_packed && !grid_multi_rank(m) && m->pos_stale && m->grid_built) {
    // the next rebuild packs 'position' AND 'moves' from the API columns: both must be current before one of them
    // (or 'type') is overwritten
    rc = materialize_field(m, 0, 1, m->eng->stream, false);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(m->dev.t[type].f[field], host, bytes, cudaMemcpyHostToDevice, m->eng->stream));
  CK(cudaStreamSynchronize(m->eng->stream));
  if (m->has_net) return sir_sync_from_api(m);
  return grid_after_write(m, field);
}

// Bring the API-visible column (type, field) up to date from the engine's packed layout, stream-ordered and
// without a host sync.  in_step: called between the launches of a step (possibly under graph capture), where the
// host-side copies of time_step / dirty flags are stale -- everything is derived on the device.
static int materialize_field(jxb_model* m, int type, int field, cudaStream_t s, bool in_step) {
  (void)type;
  if (m->has_net && m->net_built) {
    if (in_step) {
      sir_unpack_cur_kernel<<<m->eng->sms * 8, 256, 0, s>>>(m->sv, m->dev.ctrl, (int*)m->dev.t[0].f[0], m->desc.types[0].n_agents);
      m->eng->launches++;
    } else {
      int rc = sir_sync_to_api(m);
      if (rc) return rc;
    }
  }
  if (m->sch_packed && !m->grid_sharded && (field == 1 || field == 3) && m->grid_built && (in_step || m->pos_stale)) {
    cell_am_unpack_kernel<<<m->eng->sms * 8, 256, 0, s>>>(m->sd, m->sb, (int2*)m->dev.t[0].f[1], (int*)m->dev.t[0].f[3]);
    m->eng->launches++;
    CK(cudaGetLastError());
    if (!in_step) m->pos_stale = false;
  }
  if (m->has_grid && field == 2 && (in_step || m->sat_dirty)) { int rc = schelling_export_satisfied(m, s, !in_step); if (rc) return rc; }
  if (m->grid_sharded && (field == 1 || field == 3) && m->grid_built && (in_step || m->pos_stale)) {
    // this rank's view: the agents sitting in its band; -1 / 0 for everybody else (host: max / sum over ranks)
    CK(cudaMemsetAsync(m->dev.t[0].f[1], 0xFF, field_bytes(m, 0, 1), s));
    CK(cudaMemsetAsync(m->dev.t[0].f[3], 0, field_bytes(m, 0, 3), s));
    grid_shard_unpack_kernel<<<m->eng->sms * 8, 256, 0, s>>>(m->sd, m->sb, m->gs, (int2*)m->dev.t[0].f[1], (int*)m->dev.t[0].f[3]);
    m->eng->launches++;
    CK(cudaGetLastError());
    if (!in_step) m->pos_stale = false;
  }
  return JXB_OK;
}

extern "C" int jxb_model_download(jxb_model* m, int type, int field, void* host, size_t bytes) {
  int rc = check_field(m, type, field, bytes);
  if (rc) return rc;
  CK(cudaSetDevice(m->eng->device));
  rc = materialize_field(m, type, field, m->eng->stream, false);
  if (rc) return rc;
  CK(cudaMemcpyAsync(host, m->dev.t[type].f[field], bytes, cudaMemcpyDeviceToHost, m->eng->stream));
  CK(cudaStreamSynchronize(m->eng->stream));
  // 'moves' of a sharded Grid is a sum over ranks; until the first rebuild splits the uploaded values by
  // band, every rank still holds the caller's whole column: only rank 0 reports it
  if (m->grid_sharded && field == 3 && !m->grid_built && m->dev.rank != 0) memset(host, 0, bytes);
  return JXB_OK;
}

extern "C" int jxb_model_fill(jxb_model* m, int type, int field, const void* value, size_t bytes) {
  NEED(m); NEED_TYPE(m, type);
  if (field < 0 || field >= m->rules[type]->nf) return fail(JXB_ERR_INVALID, "field %d out of range", field);
  const FieldSpec& f = m->rules[type]->f[field];
  if (bytes != f.width * dtype_size(f.dtype) || bytes > 16)
    return fail(JXB_ERR_INVALID, "fill '%s': expected %zu bytes", f.name, f.width * dtype_size(f.dtype));
  CK(cudaSetDevice(m->eng->device));
  uint4 v = {0, 0, 0, 0};
  memcpy(&v, value, bytes);
  if (m->sch_packed && !grid_multi_rank(m) && m->pos_stale && m->grid_built) {
    int rc = materialize_field(m, 0, 1, m->eng->stream, false);
    if (rc) return rc;
  }
  const long long n = m->desc.types[type].n_agents;
  int blocks = (int)std::min<long long>((n * (long long)bytes + 255) / 256, m->eng->sms * 8);
  fill_kernel<<<blocks, 256, 0, m->eng->stream>>>((unsigned char*)m->dev.t[type].f[field], n, (int)bytes, v);
  m->eng->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->eng->stream));
  if (m->has_net) return sir_sync_from_api(m);
  return grid_after_write(m, field);
}

extern "C" int jxb_model_set_env(jxb_model* m, int slot, double value) {
  NEED(m);
  if (slot < 0 || slot >= m->prog->n_env) return fail(JXB_ERR_INVALID, "env slot %d out of range", slot);
  CK(cudaSetDevice(m->eng->device));
  CK(cudaMemcpy(m->dev.env + slot, &value, sizeof(double), cudaMemcpyHostToDevice));
  return JXB_OK;
}

extern "C" int jxb_model_get_env(jxb_model* m, int slot, double* value) {
  NEED(m);
  if (slot < 0 || slot >= m->prog->n_env) return fail(JXB_ERR_INVALID, "env slot %d out of range", slot);
  CK(cudaSetDevice(m->eng->device));
  CK(cudaStreamSynchronize(m->eng->stream));
  CK(cudaMemcpy(value, m->dev.env + slot, sizeof(double), cudaMemcpyDeviceToHost));
  return JXB_OK;
}

extern "C" int jxb_model_set_type_param(jxb_model* m, int type, int index, float value) {
  NEED(m); NEED_TYPE(m, type);
  if (index < 0 || index >= JXB_MAX_PARAMS) return fail(JXB_ERR_INVALID, "param index %d out of range", index);
  m->dev.t[type].p[index] = value;
  m->desc.types[type].params[index] = value;
  // kernel arguments are baked into the cached step graphs
  if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }
  if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
  return JXB_OK;
}

extern "C" int jxb_model_time_step(jxb_model* m, int64_t* out) {
  NEED(m);
  *out = m->time_step;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// Schelling grid helpers
// ---------------------------------------------------------------------------------------
extern "C" int jxb_model_grid_rebuild(jxb_model* m) {
  NEED(m);
  if (!m->has_grid) return fail(JXB_ERR_STATE, "model has no Grid");
  if (m->grid_sharded && !m->gs_attached)
    return fail(JXB_ERR_STATE, "sharded Grid: call jxb_model_grid_shard_export / jxb_model_grid_shard_attach first");
  CK(cudaSetDevice(m->eng->device));
  cudaStream_t s = m->eng->stream;
  int* d_err = nullptr;
  const size_t err_bytes = (2 + (size_t)m->sd.ntiles) * sizeof(int);
  CK(pool_alloc(m->eng, (void**)&d_err, err_bytes));     // pooled: cudaFree would synchronise the whole device
  CK(cudaMemsetAsync(d_err, 0, 2 * sizeof(int), s));
  const int blocks = m->eng->sms * 8;
  grid_clear_kernel<<<blocks, 256, 0, s>>>(m->sd, m->pad);
  grid_scatter_kernel<<<blocks, 256, 0, s>>>(m->sd, (const int*)m->dev.t[0].f[0], (const int2*)m->dev.t[0].f[1],
                                            m->desc.types[0].n_agents, d_err);
  // env['empty_cells']: ascending empty cells of the freshly built grid
  empty_count_kernel<<<m->sd.ntiles, kThreads, 0, s>>>(m->sd, (unsigned int*)(d_err + 2));
  empty_write_kernel<<<m->sd.ntiles, kThreads, 0, s>>>(m->sd, (const unsigned int*)(d_err + 2), (unsigned int*)(d_err + 1));
  m->eng->launches += 4;
  int err = 0;
  unsigned int n_empty = 0;
  CK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&n_empty, d_err + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  pool_free(m->eng, d_err, err_bytes);
  if (!err && n_empty != m->sd.n_empty) err = 2;
  CK(cudaGetLastError());
  if (err == 1) return fail(JXB_ERR_INVALID, "an agent position lies outside the grid");
  if (err == 2) return fail(JXB_ERR_INVALID, "two agents share a grid cell");
  if (m->sch_bits) {
    planes_from_ct_kernel<<<m->eng->sms * 8, 256, 0, s>>>(m->sd, m->sb);
    m->eng->launches++;
    CK(cudaGetLastError());
    m->ct_stale = false;
  }
  if (m->sch_packed) {
    cell_am_pack_kernel<<<blocks, 256, 0, s>>>(m->sd, m->sb, (const int*)m->dev.t[0].f[3]);
    m->eng->launches++;
    CK(cudaGetLastError());
    m->pos_stale = grid_multi_rank(m);      // every rank holds the whole columns now: reads derive the band view
  }
  if (m->grid_sharded) {
    // my range of the empty-cell slots moves into the receive area, where the peers reach it
    const GridShardDev& gs = m->gs;
    const unsigned int e = m->sd.n_empty, lo = std::min(e, (unsigned int)gs.rank * gs.eper), hi = std::min(e, lo + gs.eper);
    if (hi > lo) CK(cudaMemcpyAsync(m->gs_area + gs.slot_off, m->sd.E + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  m->grid_built = true;
  return JXB_OK;
}

extern "C" int jxb_model_download_grid(jxb_model* m, int32_t* host, size_t bytes) {
  NEED(m);
  if (!m->has_grid) return fail(JXB_ERR_STATE, "model has no Grid");
  if (bytes != (size_t)m->sd.cells * 4) return fail(JXB_ERR_INVALID, "grid is %lld int32", m->sd.cells);
  CK(cudaSetDevice(m->eng->device));
  if (!m->grid_built) { int rc = jxb_model_grid_rebuild(m); if (rc) return rc; }
  if (m->sch_bits && m->ct_stale) {
    ct_from_planes_kernel<<<m->eng->sms * 8, 256, 0, m->eng->stream>>>(m->sd, m->sb);
    m->eng->launches++;
    m->ct_stale = false;
  }
  int* tmp = nullptr;
  CK(cudaMalloc(&tmp, bytes));
  grid_export_kernel<<<m->eng->sms * 8, 256, 0, m->eng->stream>>>(m->sd, tmp);
  m->eng->launches++;
  CK(cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, m->eng->stream));
  CK(cudaStreamSynchronize(m->eng->stream));
  cudaFree(tmp);
  return JXB_OK;
}

extern "C" int jxb_model_download_empty_cells(jxb_model* m, int32_t* host, size_t bytes) {
  NEED(m);
  if (!m->has_grid) return fail(JXB_ERR_STATE, "model has no Grid");
  if (bytes != (size_t)m->sd.n_empty * 4) return fail(JXB_ERR_INVALID, "empty_cells holds %u int32", m->sd.n_empty);
  CK(cudaSetDevice(m->eng->device));
  if (!m->grid_built) { int rc = jxb_model_grid_rebuild(m); if (rc) return rc; }
  if (m->grid_sharded) {
    // the slots live in the ranks' receive areas (call after a barrier: the peers must have finished their run)
    const GridShardDev& gs = m->gs;
    const unsigned int e = m->sd.n_empty;
    for (int q = 0; q < gs.world; ++q) {
      const unsigned int lo = std::min(e, (unsigned int)q * gs.eper), hi = std::min(e, lo + gs.eper);
      if (hi > lo) CK(cudaMemcpyAsync(host + lo, gs.peer[q] + gs.slot_off, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, m->eng->stream));
    }
  } else if (bytes) {
    CK(cudaMemcpyAsync(host, m->sd.E, bytes, cudaMemcpyDeviceToHost, m->eng->stream));
  }
  CK(cudaStreamSynchronize(m->eng->stream));
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// Grid row-band sharding (csrc/grid_shard.cuh)
// ---------------------------------------------------------------------------------------
// band of this rank + its receive area, step-info block and per-CTA partials (idempotent)
static int grid_shard_prepare(jxb_model* m, int row_begin, int row_end) {
  const int G = m->dev.world_size, W = m->sd.W, H = m->sd.H;
  const int max_rows = (W + G - 1) / G;
  if (row_begin < 0 || row_end > W || row_end <= row_begin || row_end - row_begin > max_rows)
    return fail(JXB_ERR_INVALID, "band [%d,%d) of %d rows: every rank owns 1..%d consecutive rows", row_begin, row_end, W, max_rows);
  CK(cudaSetDevice(m->eng->device));
  GridShardDev& gs = m->gs;
  gs.rank = m->dev.rank; gs.world = G; gs.X0 = row_begin; gs.X1 = row_end;
  if (G == 1) { gs.xb[0] = 0; gs.xb[1] = W; }
  // a segment takes the cell records of one sender (targets that were empty cells of my band) from the front and its
  // halo records (sets in my two halo rows, clears from the sender's boundary rows) from the back
  // (a slot owner forwards at most one record per slot it holds)
  gs.halo_cap = (unsigned int)(4 * H);
  gs.eper = std::max(1u, (m->sd.n_empty + (unsigned int)G - 1u) / (unsigned int)G);
  gs.nfwd = (unsigned int)m->eng->sms * 2u;                       // checked to be the same on every rank at attach
  gs.fchunk = (gs.eper + gs.nfwd - 1u) / gs.nfwd;                  // a forward CTA serves at most this many requests
  gs.cap = gs.nfwd * gs.fchunk + gs.halo_cap;
  gs.slot_off = (sizeof(GridXchgHdr) + (size_t)2 * G * gs.nfwd * 4 + 15) / 16 * 16;
  gs.req_off = (gs.slot_off + (size_t)gs.eper * 4 + 15) / 16 * 16;
  gs.rec_off = gs.req_off + (size_t)2 * G * gs.eper * sizeof(uint4);
  if (!m->gs_area) {
    m->gs_area_bytes = gs.rec_off + (size_t)2 * G * gs.cap * sizeof(uint4);
    m->gs_area = G > 1 ? (unsigned char*)take_retired_area(m->eng, m->gs_area_bytes) : nullptr;
    if (!m->gs_area) CK(cudaMalloc((void**)&m->gs_area, m->gs_area_bytes));      // its own allocation: IPC handles name whole allocations
    // flags / counts of a recycled area hold the step tags of its previous model: clear them before the handle
    // leaves this call (the peers only store into the area after the attach barrier that follows)
    CK(cudaMemset(m->gs_area, 0, gs.slot_off));
    CK(cudaDeviceSynchronize());
    int rc;
    if ((rc = dev_alloc(m, &gs.info, 1))) return rc;
    CK(cudaMemset(gs.info, 0, sizeof(GridStepInfo)));
    const int nrows = row_end - row_begin, strips = (m->sb.wpr + 31) / 32;
    // ~2 strip-rows per warp, at most 2 CTAs per SM, at least one row per CTA
    gs.blocks = (int)std::max<long long>(1, std::min<long long>(std::min<long long>((long long)m->eng->sms * 2, nrows),
                                                                ((long long)nrows * strips + 15) / 16));
    if ((rc = dev_alloc(m, &gs.part, (size_t)gs.blocks))) return rc;
    if ((rc = dev_alloc(m, &gs.sendcnt, (size_t)3 * kMaxPeers))) return rc;
    CK(cudaMemset(gs.sendcnt, 0, 3 * kMaxPeers * sizeof(unsigned int)));
  }
  gs.peer[gs.rank] = m->gs_area;
  gs.self = m->gs_area;
  return JXB_OK;
}

extern "C" int jxb_model_grid_shard_export(jxb_model* m, int row_begin, int row_end, void* handle_out, size_t bytes) {
  NEED(m);
  if (!m->grid_sharded || m->dev.world_size < 2)
    return fail(JXB_ERR_STATE, "model is not a sharded Grid (desc.world_size > 1 with a Schelling program)");
  if (!handle_out || bytes < JXB_GRID_HANDLE_BYTES)
    return fail(JXB_ERR_INVALID, "need %d bytes for the IPC handle + the band", JXB_GRID_HANDLE_BYTES);
  int rc = grid_shard_prepare(m, row_begin, row_end);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, m->gs_area));
  memset(handle_out, 0, bytes);
  memcpy(handle_out, &h, sizeof(h));
  const int32_t band[3] = {row_begin, row_end, (int32_t)m->gs.nfwd};
  memcpy((char*)handle_out + sizeof(h), band, sizeof(band));
  return JXB_OK;
}

extern "C" int jxb_model_grid_shard_attach(jxb_model* m, const void* handles, size_t bytes_each, int n_ranks) {
  NEED(m);
  if (!m->grid_sharded || !m->gs_area) return fail(JXB_ERR_STATE, "call jxb_model_grid_shard_export first");
  if (!handles || n_ranks != m->dev.world_size) return fail(JXB_ERR_INVALID, "need the handles of all %d ranks", m->dev.world_size);
  if (bytes_each < JXB_GRID_HANDLE_BYTES) return fail(JXB_ERR_INVALID, "handle entries too small");
  CK(cudaSetDevice(m->eng->device));
  // the bands must tile the rows in rank order (a mover's record goes to the owner of its target row)
  int next = 0;
  for (int p = 0; p < n_ranks; ++p) {
    int32_t band[3];
    memcpy(band, (const char*)handles + (size_t)p * bytes_each + sizeof(cudaIpcMemHandle_t), sizeof(band));
    if (band[0] != next || band[1] <= band[0]) return fail(JXB_ERR_INVALID, "rank %d's band [%d,%d) does not continue at row %d", p, band[0], band[1], next);
    if ((unsigned int)band[2] != m->gs.nfwd) return fail(JXB_ERR_INVALID, "rank %d runs on a different GPU model (%d vs %u forward CTAs)", p, band[2], m->gs.nfwd);
    m->gs.xb[p] = band[0];
    next = band[1];
  }
  if (next != m->sd.W) return fail(JXB_ERR_INVALID, "the bands cover %d of %d rows", next, m->sd.W);
  m->gs.xb[n_ranks] = next;
  for (int p = 0; p < n_ranks; ++p) {
    if (p == m->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * bytes_each, sizeof(h));
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    m->gs_opened[p] = q;
    m->gs.peer[p] = (unsigned char*)q;
  }
  m->gs_attached = true;
  // kernel arguments are baked into cached graphs
  if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }
  if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
  return JXB_OK;
}

static int schelling_export_satisfied(jxb_model* m, cudaStream_t s, bool clear_dirty) {
  unsigned char* sat = (unsigned char*)m->dev.t[0].f[2];
  CK(cudaMemsetAsync(sat, 1, (size_t)m->desc.types[0].n_agents, s));
  if (m->grid_sharded) grid_shard_export_satisfied_kernel<<<m->eng->sms * 4, 256, 0, s>>>(m->sb, m->sd, m->gs, sat);   // host: min over ranks
  else if (m->sch_packed) satisfied_export_packed_kernel<<<m->eng->sms * 4, 256, 0, s>>>(m->sd, m->sb, m->dev.ctrl, sat);
  else satisfied_export_kernel<<<m->eng->sms * 4, 256, 0, s>>>(m->sd, m->dev.ctrl, sat);
  m->eng->launches++;
  CK(cudaGetLastError());
  if (clear_dirty) m->sat_dirty = false;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// SIR network helpers
// ---------------------------------------------------------------------------------------
extern "C" int jxb_model_set_network(jxb_model* m, const int32_t* edges, int64_t n_edges) {
  NEED(m);
  if (!m->has_net) return fail(JXB_ERR_STATE, "model has no Network");
  if (n_edges < 0 || n_edges >= (1ll << 31)) return fail(JXB_ERR_UNSUPPORTED, "edge count out of range");
  CK(cudaSetDevice(m->eng->device));
  const long long n = m->desc.types[0].n_agents;
  // bin by source on the device (counting sort: degree histogram, exclusive scan, cursor scatter);
  // neighbour order inside a row is irrelevant to the rule
  cudaStream_t st = m->eng->stream;
  int rc;
  if (!m->net_allocs.empty()) {
    // rebuild (Model.add_env_state('network_edges') after initialize(), Network.add_edge): the previous CSR,
    // state and counter buffers go back to the pool; their pointers are baked into the cached step graphs
    if (m->net_built) { rc = sir_sync_to_api(m); if (rc) return rc; }     // the live state lives in the packed buffers
    CK(cudaStreamSynchronize(st));
    if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }
    if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
    for (auto& a : m->net_allocs) pool_free(m->eng, a.first, a.second);
    m->net_allocs.clear();
    m->net_built = false;
  }
  unsigned int* d_rp = nullptr; int* d_col = nullptr;
  if ((rc = dev_alloc(m, &d_rp, (size_t)n + 2, &m->net_allocs))) return rc;
  if ((rc = dev_alloc(m, &d_col, (size_t)std::max<int64_t>(n_edges, 1), &m->net_allocs))) return rc;
  // the row extents come back into a page-locked block (cached by size): 40 MB at C3, ~1 ms instead of ~4
  void* h_rp = nullptr;
  if ((rc = jxb_host_alloc(((size_t)n + 2) * 4, &h_rp))) return rc;
  struct HostBlock { void* p; ~HostBlock() { jxb_host_free(p); } } h_rp_guard{h_rp};
  unsigned int* row_ptr = (unsigned int*)h_rp;
  {
    void* d_edges = nullptr; void* d_cursor = nullptr; void* d_sums = nullptr; void* d_flag = nullptr;
    const int ntiles = (int)((n + kScanTile - 1) / kScanTile);
    const size_t eb = (size_t)std::max<int64_t>(n_edges, 1) * 8, cb = ((size_t)n + 2) * 4, sb = ((size_t)ntiles + 2) * 4;
    auto release = [&]() {
      cudaStreamSynchronize(st);
      pool_free(m->eng, d_edges, eb); pool_free(m->eng, d_cursor, cb); pool_free(m->eng, d_sums, sb); pool_free(m->eng, d_flag, 16);
    };
#define NCK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { release(); return fail(JXB_ERR_CUDA, \
    "%s failed: %s", #call, cudaGetErrorString(_e)); } } while (0)
    NCK(pool_alloc(m->eng, &d_edges, eb));
    NCK(pool_alloc(m->eng, &d_cursor, cb));
    NCK(pool_alloc(m->eng, &d_sums, sb));
    NCK(pool_alloc(m->eng, &d_flag, 16));
    NCK(cudaMemsetAsync(d_rp, 0, cb, st));
    NCK(cudaMemsetAsync(d_flag, 0, 16, st));
    if (n_edges) NCK(cudaMemcpyAsync(d_edges, edges, (size_t)n_edges * 8, cudaMemcpyHostToDevice, st));
    const int g = m->eng->sms * 8;
    csr_count_kernel<<<g, kThreads, 0, st>>>((const int2*)d_edges, n_edges, n, (long long)m->desc.types[0].global_n, d_rp, (int*)d_flag);
    scan_tile_sums_kernel<<<ntiles, kThreads, 0, st>>>(d_rp, n, (unsigned int*)d_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>((unsigned int*)d_sums, ntiles);
    scan_apply_kernel<<<ntiles, kThreads, 0, st>>>(d_rp, n, (const unsigned int*)d_sums, d_rp, (unsigned int*)d_cursor);
    csr_fill_kernel<<<g, kThreads, 0, st>>>((const int2*)d_edges, n_edges, n, (long long)m->desc.types[0].global_n,
                                            (unsigned int*)d_cursor, d_col);
    m->eng->launches += 5;
    int bad = 0;
    NCK(cudaMemcpyAsync(&bad, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    NCK(cudaMemcpyAsync(row_ptr, d_rp, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, st));
    NCK(cudaStreamSynchronize(st));
    NCK(cudaGetLastError());
#undef NCK
    release();
    if (bad) return fail(JXB_ERR_INVALID, "an edge references an agent outside [0,%lld) (sources are local rows, targets global ids)",
                         (long long)m->desc.types[0].global_n);
    if (row_ptr[n] != (unsigned int)n_edges) return fail(JXB_ERR_CUDA, "CSR build lost edges (%u of %lld)", row_ptr[n], (long long)n_edges);
  }
  // row blocks: whole 32-row groups, greedy up to kSirTile entries / 1024 rows
  std::vector<int> rb;
  rb.push_back(0);
  long long r = 0;
  while (r < n) {
    long long end = r;
    const unsigned int e0 = row_ptr[r];
    while (end < n) {
      const long long g_end = std::min<long long>(end + 32, n);
      const bool first = end == r;
      if (!first && (row_ptr[g_end] - e0 > (unsigned)kSirTile || g_end - r > kThreads * kSirRowsPerThread)) break;
      end = g_end;
    }
    rb.push_back((int)end);
    r = end;
  }
  SirDev& sv = m->sv;
  int* d_rb; float* d_esc;
  if ((rc = dev_alloc(m, &d_rb, rb.size(), &m->net_allocs))) return rc;
  if ((rc = dev_alloc(m, &d_esc, kSirKCap + 1, &m->net_allocs))) return rc;
  if (m->net_sharded) {
    // the two GLOBAL bitmaps live in one IPC-shareable allocation behind the exchange header
    const long long gn = m->desc.types[0].global_n;
    const size_t words = (size_t)(gn + 31) / 32 + 1;
    sv.bits_stride = (words * 4 + 255) & ~(size_t)255;
    if (!m->ns_area) {
      m->ns_area_bytes = sizeof(SirXchgHdr) + 2 * sv.bits_stride;
      m->ns_area = (unsigned char*)take_retired_area(m->eng, m->ns_area_bytes);
      if (!m->ns_area) CK(cudaMalloc((void**)&m->ns_area, m->ns_area_bytes));
    }
    CK(cudaMemset(m->ns_area, 0, m->ns_area_bytes));
    CK(cudaDeviceSynchronize());
    sv.world = m->dev.world_size; sv.rank = m->dev.rank;
    sv.gw0 = (unsigned int)(m->desc.types[0].global_offset / 32);
    sv.gwords = (unsigned int)((n + 31) / 32);
    sv.self = m->ns_area;
    sv.peer[sv.rank] = m->ns_area;
  }
  for (int b = 0; b < 2; ++b) {
    if ((rc = dev_alloc(m, &sv.state8[b], (size_t)n + 32, &m->net_allocs))) return rc;
    cudaMemset(sv.state8[b], 0, (size_t)n + 32);
    if (m->net_sharded) {
      sv.infbits[b] = (unsigned int*)(m->ns_area + sizeof(SirXchgHdr) + (size_t)b * sv.bits_stride);
    } else {
      if ((rc = dev_alloc(m, &sv.infbits[b], (size_t)(n + 31) / 32 + 1, &m->net_allocs))) return rc;
      cudaMemset(sv.infbits[b], 0, ((size_t)(n + 31) / 32 + 1) * 4);
    }
  }
  CK(cudaMemcpy(d_rb, rb.data(), rb.size() * 4, cudaMemcpyHostToDevice));
  {
    // escape[k] = (1-beta)^k as a float32 product chain (DESIGN.md "SIR rule")
    std::vector<float> q(kSirKCap + 1);
    const float b = 1.0f - m->desc.types[0].params[0];
    q[0] = 1.0f;
    for (int k = 1; k <= kSirKCap; ++k) { volatile float v = q[k - 1] * b; q[k] = v; }
    CK(cudaMemcpy(d_esc, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
  }
  sv.row_ptr = d_rp; sv.col = d_col; sv.rb = d_rb; sv.nrb = (int)rb.size() - 1; sv.escape = d_esc;
  if ((rc = dev_alloc(m, &sv.partials, (size_t)std::max<long long>(std::max<long long>(sv.nrb, (long long)m->eng->sms * 8), (n + kThreads * kSirRowsPerThread - 1) / (kThreads * kSirRowsPerThread)) * 3 + 3, &m->net_allocs))) return rc;
  m->nnz = n_edges;
  m->net_built = true;
  {
    // push formulation: counters + the static list of heavy rows
    sv.big_len = 256u;
    if (const char* ev = getenv("JXB_SIR_BIG")) sv.big_len = (unsigned int)std::max(1, atoi(ev));
    std::vector<int> heavy;
    for (long long i = 0; i < n; ++i)
      if (row_ptr[i + 1] - row_ptr[i] > (unsigned)kSirHeavy) heavy.push_back((int)i);
    int* d_heavy = nullptr;
    if ((rc = dev_alloc(m, &d_heavy, heavy.size() + 1, &m->net_allocs))) return rc;
    if (!heavy.empty()) CK(cudaMemcpy(d_heavy, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice));
    if ((rc = dev_alloc(m, &sv.k32, (size_t)n + 32, &m->net_allocs))) return rc;
    CK(cudaMemset(sv.k32, 0, ((size_t)n + 32) * 4));
    sv.heavy = d_heavy;
    sv.n_heavy = (int)heavy.size();
    m->sir_tblocks = (int)std::max<long long>(1, std::min<long long>((n + kThreads * kSirRowsPerThread - 1) / (kThreads * kSirRowsPerThread), (long long)m->eng->sms * 8));
    if (const char* cps = getenv("JXB_SIR_PULL_CPS")) m->sir_pull_cps = std::max(1, std::min(8, atoi(cps)));
    const char* mode = getenv("JXB_SIR_MODE");
    m->sir_mode = 3;
    if (mode && !strcmp(mode, "pull")) m->sir_mode = 0;
    else if (mode && !strcmp(mode, "push")) m->sir_mode = 1;
    else if (mode && !strcmp(mode, "pull_s")) m->sir_mode = 2;
    if (m->net_sharded) m->sir_mode = 2;     // a rank only holds its own rows: pull over them (a push would scatter remotely)
    sv.auto_mode = m->sir_mode == 3;
    const int dev_mode = m->sir_mode == 2 ? 0 : 1;       // auto starts in the push direction
    CK(cudaMemcpy(&m->dev.ctrl->sir_mode, &dev_mode, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&m->dev.ctrl->sir_mode_next, &dev_mode, sizeof(int), cudaMemcpyHostToDevice));
    const size_t max_ctas = (size_t)std::max<long long>(m->sir_tblocks, (long long)m->eng->sms * 8);
    if ((rc = dev_alloc(m, &sv.degsum, max_ctas * 2 + 2, &m->net_allocs))) return rc;
  }
  return sir_sync_from_api(m);
}

// ---------------------------------------------------------------------------------------
// Network node-range sharding (csrc/sir.cuh)
// ---------------------------------------------------------------------------------------
extern "C" int jxb_model_net_shard_export(jxb_model* m, void* handle_out, size_t bytes) {
  NEED(m);
  if (!m->net_sharded || !m->net_built || !m->ns_area)
    return fail(JXB_ERR_STATE, "not a sharded Network model with its network set (desc.world_size > 1, jxb_model_set_network)");
  if (!handle_out || bytes < sizeof(cudaIpcMemHandle_t))
    return fail(JXB_ERR_INVALID, "need %zu bytes for the IPC handle", sizeof(cudaIpcMemHandle_t));
  CK(cudaSetDevice(m->eng->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, m->ns_area));
  memset(handle_out, 0, bytes);
  memcpy(handle_out, &h, sizeof(h));
  return JXB_OK;
}

extern "C" int jxb_model_net_shard_attach(jxb_model* m, const void* handles, size_t bytes_each, int n_ranks) {
  NEED(m);
  if (!m->net_sharded || !m->ns_area) return fail(JXB_ERR_STATE, "call jxb_model_net_shard_export first");
  if (!handles || n_ranks != m->dev.world_size || bytes_each < sizeof(cudaIpcMemHandle_t))
    return fail(JXB_ERR_INVALID, "need the IPC handles of all %d ranks", m->dev.world_size);
  CK(cudaSetDevice(m->eng->device));
  for (int p = 0; p < n_ranks; ++p) {
    if (p == m->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)p * bytes_each, sizeof(h));
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    m->ns_opened[p] = q;
    m->sv.peer[p] = (unsigned char*)q;
  }
  m->ns_attached = true;
  if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }
  if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
  return JXB_OK;
}

// hand my slice of the current infected bitmap to every peer (after init / an upload of 'state'); the host
// shim barriers afterwards, so that every rank's copy is whole before anybody steps
extern "C" int jxb_model_net_shard_sync(jxb_model* m) {
  NEED(m);
  if (!m->net_sharded || !m->ns_attached) return fail(JXB_ERR_STATE, "sharded Network: export / attach the peers first");
  CK(cudaSetDevice(m->eng->device));
  sir_shard_sync_kernel<<<m->eng->sms * 2, 256, 0, m->eng->stream>>>(m->sv, (int)(m->time_step & 1));
  m->eng->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->eng->stream));
  return JXB_OK;
}

static int sir_sync_from_api(jxb_model* m) {
  if (!m->net_built) return JXB_OK;
  const long long n = m->desc.types[0].n_agents;
  const int cur = (int)(m->time_step & 1);
  const int blocks = (int)((n + 255) / 256);
  sir_pack_kernel<<<blocks, 256, 0, m->eng->stream>>>((const int*)m->dev.t[0].f[0], m->sv.state8[cur],
                                                      m->sv.infbits[cur] + (m->net_sharded ? m->sv.gw0 : 0u), n);
  m->eng->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->eng->stream));
  return JXB_OK;
}

static int sir_sync_to_api(jxb_model* m) {
  if (!m->net_built) return JXB_OK;
  const long long n = m->desc.types[0].n_agents;
  const int cur = (int)(m->time_step & 1);
  sir_unpack_kernel<<<m->eng->sms * 8, 256, 0, m->eng->stream>>>(m->sv.state8[cur], (int*)m->dev.t[0].f[0], n);
  m->eng->launches++;
  CK(cudaGetLastError());
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// initialisation (model.py:118-144, agent.py:92-130)
// ---------------------------------------------------------------------------------------
static int launch_init(jxb_model* m, int type, Key key) {
  const TypeDev& t = m->dev.t[type];
  int blocks = (int)std::min<long long>((t.n + 255) / 256, (long long)m->eng->sms * 16);
  if (m->traced) {
    const int r = m->traced_init(&m->dev, type, key.a, key.b, m->desc.rng_mode, blocks, (void*)m->eng->stream);
    m->eng->launches++;
    if (r) return fail(JXB_ERR_CUDA, "traced init kernel launch failed (%d)", r);
    m->collections_ready[type] = true;
    return JXB_OK;
  }
  if (m->desc.rng_mode == JXB_RNG_PARTITIONABLE)
    init_kernel<1><<<blocks, 256, 0, m->eng->stream>>>(t, key);
  else
    init_kernel<0><<<blocks, 256, 0, m->eng->stream>>>(t, key);
  m->eng->launches++;
  CK(cudaGetLastError());
  m->collections_ready[type] = true;
  return JXB_OK;
}

extern "C" int jxb_collection_init(jxb_model* m, int type, uint32_t k0, uint32_t k1) {
  NEED(m); NEED_TYPE(m, type);
  CK(cudaSetDevice(m->eng->device));
  int rc = launch_init(m, type, Key{k0, k1});
  if (rc) return rc;
  CK(cudaStreamSynchronize(m->eng->stream));
  if (m->has_net) return sir_sync_from_api(m);
  if (m->has_grid) m->grid_built = false;
  return JXB_OK;
}

// Economy: where the first step's kernels put their shared-memory histogram window (later steps take it from the
// previous step's scan).  A strided sample of the freshly initialised incomes, binned on the host.
static int eco_seed_window(jxb_model* m) {
  const int hh = m->eco_hh;
  if (hh < 0) return JXB_OK;
  const long long n = m->desc.types[hh].n_agents;
  const long long stride = std::max<long long>(1, n / 8192);
  const size_t rows = (size_t)((n + stride - 1) / stride);
  std::vector<float> sample(rows);
  CK(cudaMemcpy2D(sample.data(), sizeof(float), m->dev.t[hh].f[1], (size_t)stride * sizeof(float), sizeof(float), rows,
                  cudaMemcpyDeviceToHost));
  const int tiles = kGiniBins / kGiniScanTile;
  std::vector<unsigned int> cnt((size_t)tiles + 1, 0u);
  for (float x : sample) {
    unsigned int b;
    memcpy(&b, &x, 4);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);            // gini_bin (csrc/economy.cuh)
    cnt[(b >> (32 - kGiniBits)) / kGiniScanTile] += 1u;
  }
  int best = 0;
  for (int t = 0; t + 1 < tiles; ++t)
    if (cnt[t] + cnt[t + 1] > cnt[best] + cnt[best + 1]) best = t;
  CK(cudaMemcpy(m->eco.tile_range + 2, &best, sizeof(int), cudaMemcpyHostToDevice));
  return JXB_OK;
}

extern "C" int jxb_model_init(jxb_model* m, uint32_t k0, uint32_t k1) {
  NEED(m);
  CK(cudaSetDevice(m->eng->device));
  const int C = m->desc.n_types, mode = m->desc.rng_mode;
  const Key root{k0, k1};
  m->rng = split_child(mode, root, 0, C + 1);                       // model.py:129-130
  for (int i = 0; i < C; ++i) {
    int rc = launch_init(m, i, split_child(mode, root, i + 1, C + 1));   // model.py:133-137
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(m->eng->stream));
  m->initialized = true;
  if (m->has_eco) { int rc = eco_seed_window(m); if (rc) return rc; }
  if (m->has_net) return sir_sync_from_api(m);
  if (m->has_grid) m->grid_built = false;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// step launch plan
// ---------------------------------------------------------------------------------------
static int plan_step_blocks(jxb_model* m) {
  ModelDev& md = m->dev;
  if (md.program == JXB_PROGRAM_ECONOMY) {
    // one launch per collection, each a single resident wave of its own kernel
    int begin = 0;
    for (int i = 0; i < md.n_types; ++i) {
      int occ = 0;
      if (md.t[i].rule == JXB_RULE_HOUSEHOLD)
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, economy_step_kernel<1, JXB_RULE_HOUSEHOLD>, kThreads, 0));
      else
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, economy_step_kernel<1, JXB_RULE_CONSUMER_FIRM>, kThreads, 0));
      if (occ < 1) occ = 1;
      // households: four agents per thread per iteration
      const long long per_thread = md.t[i].rule == JXB_RULE_HOUSEHOLD ? 4 : 1;
      const long long need = std::max<long long>(1, (md.t[i].n / per_thread + kThreads - 1) / kThreads);
      const int nb = (int)std::min<long long>(need, (long long)m->eng->sms * occ);
      md.t[i].block_begin = begin;
      md.t[i].block_count = nb;
      begin += nb;
    }
    md.grid_blocks = begin;
    m->step_blocks = begin;
    return JXB_OK;
  }
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel<1>, kThreads, 0));
  if (occ < 1) occ = 1;
  const int budget = m->eng->sms * occ;
  long long total = 0;
  for (int i = 0; i < md.n_types; ++i) total += md.t[i].n;
  int begin = 0;
  for (int i = 0; i < md.n_types; ++i) {
    const long long need = std::max<long long>(1, (md.t[i].n / kVec + kThreads - 1) / kThreads);
    long long share = std::max<long long>(1, (long long)((double)budget * (double)md.t[i].n / (double)total));
    int nb = (int)std::min(need, share);
    md.t[i].block_begin = begin;
    md.t[i].block_count = nb;
    begin += nb;
  }
  md.grid_blocks = begin;
  m->step_blocks = begin;
  return JXB_OK;
}

// per-agent series (csrc/record.cuh): after the step's tail advanced the device-side counters, copy every
// recorded column into its ring slot if this step recorded a history row.  Capture-safe.
static int enqueue_snapshots(jxb_model* m, cudaStream_t s) {
  for (const auto& rf : m->rec_fields) {
    int rc = materialize_field(m, rf.type, rf.field, s, true);
    if (rc) return rc;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((rf.bytes / 16 + kThreads - 1) / kThreads, (size_t)m->eng->sms * 8));
    series_snapshot_kernel<<<blocks, kThreads, 0, s>>>(m->dev.ctrl, m->dev.collect_interval, (const uint4*)m->dev.t[rf.type].f[rf.field],
                                                       m->d_series + rf.off, (unsigned long long)rf.bytes, (unsigned long long)rf.stride);
    m->eng->launches++;
  }
  CK(cudaGetLastError());
  return JXB_OK;
}

// enqueue one Model.step (model.py:146-216) on the stream; capture-safe (no host-varying args)
static int enqueue_step(jxb_model* m, cudaStream_t s, bool timed) {
  jxb_engine* eng = m->eng;
  const bool part = m->desc.rng_mode == JXB_RNG_PARTITIONABLE;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    m->prof_events.push_back(e0); m->prof_events.push_back(e1);
  }
  switch (m->desc.program) {
    case JXB_PROGRAM_SIR: {
      if (m->net_sharded) {
        // node-range shard: pull over my rows (new bitmap words stored into every rank's copy) + flag wait
        if (!m->ns_attached) return fail(JXB_ERR_STATE, "sharded Network step without attached peers");
        const int pgrid = m->eng->sms * m->sir_pull_cps;
        if (timed) cudaEventRecord(e0, s);
        if (part) sir_pull_s_kernel<1, true><<<pgrid, kThreads, 0, s>>>(m->sv, m->dev);
        else sir_pull_s_kernel<0, true><<<pgrid, kThreads, 0, s>>>(m->sv, m->dev);
        if (timed) cudaEventRecord(e1, s);
        sir_shard_wait_kernel<<<1, 32, 0, s>>>(m->sv, m->dev);
        eng->launches += 2;
        break;
      }
      if (timed) cudaEventRecord(e0, s);
      if (m->sir_mode == 0) {
        if (part) sir_step_kernel<1><<<m->sv.nrb, kThreads, 0, s>>>(m->sv, m->dev);
        else sir_step_kernel<0><<<m->sv.nrb, kThreads, 0, s>>>(m->sv, m->dev);
      } else {
        const int pgrid = m->eng->sms * 8;
        if (m->sir_mode == 3) {
          sir_begin_step_kernel<<<1, 32, 0, s>>>(m->dev.ctrl);
          eng->launches += 1;
        }
        if (m->sir_mode != 2) {          // push direction (self-gated on ctrl->sir_mode in auto mode)
          sir_push_kernel<<<pgrid, kThreads, 0, s>>>(m->sv, m->dev);
          if (part) sir_transition_kernel<1><<<m->sir_tblocks, kThreads, 0, s>>>(m->sv, m->dev);
          else sir_transition_kernel<0><<<m->sir_tblocks, kThreads, 0, s>>>(m->sv, m->dev);
          eng->launches += 1;
        }
        if (m->sir_mode != 1) {          // pull direction
          const int qgrid = m->eng->sms * m->sir_pull_cps;
          if (part) sir_pull_s_kernel<1><<<qgrid, kThreads, 0, s>>>(m->sv, m->dev);
          else sir_pull_s_kernel<0><<<qgrid, kThreads, 0, s>>>(m->sv, m->dev);
          if (m->sir_mode == 3) eng->launches += 1;
        }
      }
      if (timed) cudaEventRecord(e1, s);
      eng->launches += 1;
      break;
    }
    case JXB_PROGRAM_SCHELLING: {     // row-band shard of a grid (the single-GPU run is one persistent launch)
      if (!m->grid_sharded || !m->gs_attached) return fail(JXB_ERR_STATE, "sharded Grid step without attached peers");
      static const bool trace = getenv("JXB_GS_TRACE") && atoi(getenv("JXB_GS_TRACE")) != 0;
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &cap);
      const bool tr = trace && cap == cudaStreamCaptureStatusNone;
      auto mark = [&]() { if (tr) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); m->gs_trace.push_back(e); } };
      mark();
      if (timed) cudaEventRecord(e0, s);
      grid_shard_sweep_kernel<<<m->gs.blocks, kThreads, 0, s>>>(m->sd, m->sb, m->gs);
      mark();
      grid_shard_counts_kernel<<<1, kThreads, 0, s>>>(m->sd, m->gs, m->dev);
      mark();
      if (part) grid_shard_moveout_kernel<1><<<m->gs.blocks, kThreads, 0, s>>>(m->sd, m->sb, m->gs, m->dev);
      else grid_shard_moveout_kernel<0><<<m->gs.blocks, kThreads, 0, s>>>(m->sd, m->sb, m->gs, m->dev);
      mark();
      grid_shard_forward_kernel<<<m->gs.nfwd, kThreads, 0, s>>>(m->sd, m->gs);
      mark();
      grid_shard_apply_kernel<<<eng->sms * 8, 256, 0, s>>>(m->sd, m->sb, m->gs);
      mark();
      if (timed) cudaEventRecord(e1, s);        // profile: the five band kernels of a step are timed as one unit
      eng->launches += 5;
      break;
    }
    case JXB_PROGRAM_TRACED: {
      if (timed) cudaEventRecord(e0, s);
      const int variant = m->traced_started ? m->traced_variants - 1 : 0;
      const int r = m->traced_step(&m->dev, m->desc.rng_mode, variant, (void*)s);
      if (timed) cudaEventRecord(e1, s);
      eng->launches += 1;
      if (r) return fail(JXB_ERR_CUDA, "traced step kernel launch failed (%d)", r);
      break;
    }
    case JXB_PROGRAM_ECONOMY: {
      if (timed) cudaEventRecord(e0, s);
      for (int ti = 0; ti < m->desc.n_types; ++ti) {
        const TypeDev& t = m->dev.t[ti];
        const bool hhk = t.rule == JXB_RULE_HOUSEHOLD;
        if (part) {
          if (hhk) economy_step_kernel<1, JXB_RULE_HOUSEHOLD><<<t.block_count, kThreads, 0, s>>>(m->dev, m->eco, ti, t.block_begin, m->step_blocks);
          else economy_step_kernel<1, JXB_RULE_CONSUMER_FIRM><<<t.block_count, kThreads, 0, s>>>(m->dev, m->eco, ti, t.block_begin, m->step_blocks);
        } else {
          if (hhk) economy_step_kernel<0, JXB_RULE_HOUSEHOLD><<<t.block_count, kThreads, 0, s>>>(m->dev, m->eco, ti, t.block_begin, m->step_blocks);
          else economy_step_kernel<0, JXB_RULE_CONSUMER_FIRM><<<t.block_count, kThreads, 0, s>>>(m->dev, m->eco, ti, t.block_begin, m->step_blocks);
        }
        eng->launches += 1;
      }
      if (timed) cudaEventRecord(e1, s);
      const int hh = m->eco_hh;
      const int tiles = kGiniBins / kGiniScanTile;
      if (hh >= 0 && m->dev.world_size > 1) {
        // sharded population: the whole population's income counts = the sum of the ranks' histograms, read out of
        // the peers' exchange allocations (the step kernels' exchange ordered them); no NCCL call on the step path
        gini_gather_kernel<<<tiles, kThreads, 0, s>>>(m->dev, m->eco);
        eng->launches += 1;
      }
      if (hh >= 0) {
        gini_scan_sums_kernel<<<tiles, kThreads, 0, s>>>(m->eco, m->eco.bin_count, m->eco.scan_sums);
        gini_scan_top_kernel<<<1, 1024, 0, s>>>(m->eco.scan_sums, tiles, m->eco.tile_range + 2);
        gini_scan_apply_kernel<<<tiles, kThreads, 0, s>>>(m->eco, m->eco.bin_count, m->eco.scan_sums, m->eco.bin_base);
        eng->launches += 3;
      }
      gini_accumulate_kernel<<<hh >= 0 ? m->eco.gini_blocks : 1, kThreads, 0, s>>>(m->dev, m->eco, hh);
      eng->launches += 1;
      // this rank's bins back to zero for the next step's histogram (populated tiles only)
      if (hh >= 0) {
        gini_clear_kernel<<<tiles, kThreads, 0, s>>>(m->eco);
        eng->launches += 1;
      }
      break;
    }
    default: {
      if (timed) cudaEventRecord(e0, s);
      if (part) step_kernel<1><<<m->step_blocks, kThreads, 0, s>>>(m->dev);
      else step_kernel<0><<<m->step_blocks, kThreads, 0, s>>>(m->dev);
      if (timed) cudaEventRecord(e1, s);
      eng->launches += 1;
      if (m->dev.exchange == 2) {
        // NCCL path (JXB_EXCHANGE=nccl or no peer memory): sum/max/int slots as three small all-reduces
        if (!eng->nccl_comm) return fail(JXB_ERR_STATE, "sharded model but no NCCL communicator attached");
        double* buf = m->dev.allreduce_buf;
        int r = g_nccl.AllReduce(buf, buf, kFSum, /*ncclFloat64*/ 8, /*ncclSum*/ 0, eng->nccl_comm, s);
        if (!r) r = g_nccl.AllReduce(buf + kFSum, buf + kFSum, kFMax, 8, /*ncclMax*/ 2, eng->nccl_comm, s);
        if (!r) r = g_nccl.AllReduce(buf + kFSum + kFMax, buf + kFSum + kFMax, kISum, 8, 0, eng->nccl_comm, s);
        if (r) return fail(JXB_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
        tail_kernel<<<1, 32, 0, s>>>(m->dev);
        eng->launches += 1;
      }
    }
  }
  CK(cudaGetLastError());
  return enqueue_snapshots(m, s);
}

static int build_graph(jxb_model* m, int chunk, cudaGraphExec_t* out) {
  cudaStream_t s = m->eng->stream;
  cudaGraph_t g = nullptr;
  const int64_t before = m->eng->launches;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < chunk; ++i) {
    int rc = enqueue_step(m, s, false);
    if (rc) { cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); return rc; }
  }
  CK(cudaStreamEndCapture(s, &g));
  m->eng->launches = before;   // capture enqueued nothing
  cudaError_t e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(JXB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  return JXB_OK;
}

// Schelling: the whole run is ONE cooperative launch (csrc/schelling.cuh)
static int launch_schelling(jxb_model* m, int steps, cudaStream_t s) {
  const bool fast = (m->sd.H % 16) == 0, part = m->desc.rng_mode == JXB_RNG_PARTITIONABLE;
  if (m->sch_bits) {
    void* args[] = {(void*)&m->sd, (void*)&m->sb, (void*)&m->dev, (void*)&steps};
    const void* fn = part ? (const void*)schelling_bits_kernel<1> : (const void*)schelling_bits_kernel<0>;
    CK(cudaLaunchCooperativeKernel(fn, dim3(m->sch_blocks), dim3(kThreads), args, 0, s));
    m->ct_stale = true;
    m->pos_stale = true;
    m->eng->launches += 1;
    return JXB_OK;
  }
  void* args[] = {(void*)&m->sd, (void*)&m->dev, (void*)&steps};
  const void* fn = fast ? (part ? (const void*)schelling_run_kernel<true, 1> : (const void*)schelling_run_kernel<true, 0>)
                        : (part ? (const void*)schelling_run_kernel<false, 1> : (const void*)schelling_run_kernel<false, 0>);
  CK(cudaLaunchCooperativeKernel(fn, dim3(m->sch_blocks), dim3(kThreads), args, 0, s));
  m->eng->launches += 1;
  return JXB_OK;
}

static int launches_per_step(jxb_model* m);
static int launches_per_step_all(jxb_model* m) {
  int extra = 0;
  for (const auto& rf : m->rec_fields) {
    extra += 1;
    if (m->has_net) extra += 1;
    if (m->has_grid && rf.field == 2) extra += 1;
    if (m->grid_sharded && (rf.field == 1 || rf.field == 3)) extra += 1;
  }
  return launches_per_step(m) + extra;
}
static int launches_per_step(jxb_model* m) {
  if (m->grid_sharded) return 5;
  if (m->net_sharded) return 2;
  if (m->desc.program == JXB_PROGRAM_SIR) return m->sir_mode == 3 ? 4 : (m->sir_mode == 1 ? 2 : 1);
  if (m->desc.program == JXB_PROGRAM_ECONOMY) return m->desc.n_types + (m->eco_hh >= 0 ? (m->dev.world_size > 1 ? 6 : 5) : 1);
  return (m->dev.exchange == 2) ? 2 : 1;     // + the NCCL kernels, which are not ours
}

extern "C" int jxb_model_set_profile(jxb_model* m, int enable) {
  NEED(m);
  m->profile = enable != 0;
  return JXB_OK;
}

extern "C" int jxb_model_profile(jxb_model* m, double* seconds, int64_t* launches, const char** name) {
  NEED(m);
  if (seconds) *seconds = m->prof_seconds;
  if (launches) *launches = m->prof_launches;
  if (name) {
    switch (m->desc.program) {
      case JXB_PROGRAM_SCHELLING: *name = m->grid_sharded ? "grid_shard_sweep_kernel" : (m->sch_bits ? "schelling_bits_kernel" : "schelling_run_kernel"); break;
      case JXB_PROGRAM_ECONOMY: *name = "economy_step_kernel"; break;
      case JXB_PROGRAM_TRACED: *name = "jxc_step_kernel (traced)"; break;
      case JXB_PROGRAM_SIR:
        *name = m->sir_mode == 0 ? "sir_step_kernel" : (m->sir_mode == 1 ? "sir_push_kernel+sir_transition_kernel"
                : (m->sir_mode == 2 ? "sir_pull_s_kernel" : "sir_push_kernel+sir_transition_kernel|sir_pull_s_kernel"));
        break;
      default: *name = "step_kernel";
    }
  }
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// Model.run (model.py:218-262)
// ---------------------------------------------------------------------------------------
extern "C" int jxb_model_run(jxb_model* m, int steps, int collect_interval, double* metrics_out,
                             int32_t* steps_out, int* n_records_out, double* device_seconds_out) {
  NEED(m);
  if (!m->initialized)
    return fail(JXB_ERR_STATE, "Model must be initialized before stepping. Call initialize() first.");
  if (steps < 0) return fail(JXB_ERR_INVALID, "steps must be >= 0");
  if (collect_interval < 1) return fail(JXB_ERR_INVALID, "collect_interval must be >= 1");
  jxb_engine* eng = m->eng;
  CK(cudaSetDevice(eng->device));
  cudaStream_t s = eng->stream;
  if (m->has_grid && !m->grid_built) { int rc = jxb_model_grid_rebuild(m); if (rc) return rc; }
  if (m->has_net && !m->net_built) return fail(JXB_ERR_STATE, "SIR model has no network; call jxb_model_set_network");

  const int C = m->desc.n_types, mode = m->desc.rng_mode, stride = (C + 1) * 2;
  // ---- key schedule of the whole run (model.py:156,164,183), host scalar work ----------
  if ((size_t)steps * stride > m->keys_cap) {
    CK(cudaStreamSynchronize(s));
    pool_free(eng, m->d_keys, m->keys_cap * 4);
    m->keys_cap = (size_t)std::max(steps, 16) * stride;
    m->d_keys = nullptr;
    CK(pool_alloc(eng, (void**)&m->d_keys, m->keys_cap * 4));
    m->h_keys.resize(m->keys_cap);
  }
  for (int t = 0; t < steps; ++t) {
    Key step_key = split_child(mode, m->rng, 1, 2);
    m->rng = split_child(mode, m->rng, 0, 2);
    uint32_t* row = m->h_keys.data() + (size_t)t * stride;
    for (int c = 0; c < C; ++c) {
      Key ck = split_child(mode, step_key, 1, 2);
      step_key = split_child(mode, step_key, 0, 2);
      row[2 * c] = ck.a; row[2 * c + 1] = ck.b;
    }
    Key uk{0, 0};
    if (m->prog->has_env_fn) uk = split_child(mode, step_key, 1, 2);
    row[2 * C] = uk.a; row[2 * C + 1] = uk.b;
  }
  if (steps) CK(cudaMemcpyAsync(m->d_keys, m->h_keys.data(), (size_t)steps * stride * 4, cudaMemcpyHostToDevice, s));
  m->dev.keys = m->d_keys;

  // ---- history ring ---------------------------------------------------------------------
  const long long t0 = m->time_step;
  const int n_rec = (int)((t0 + steps) / collect_interval - t0 / collect_interval);
  if ((size_t)n_rec + 1 > m->rec_cap) {
    CK(cudaStreamSynchronize(s));
    pool_free(eng, m->d_metrics, m->rec_cap * kMaxMetrics * sizeof(double));
    pool_free(eng, m->d_rec, m->rec_cap * sizeof(int));
    m->rec_cap = (size_t)n_rec + 64;
    m->d_metrics = nullptr; m->d_rec = nullptr;
    CK(pool_alloc(eng, (void**)&m->d_metrics, m->rec_cap * kMaxMetrics * sizeof(double)));
    CK(pool_alloc(eng, (void**)&m->d_rec, m->rec_cap * sizeof(int)));
  }
  m->dev.metrics = m->d_metrics;
  m->dev.record_steps = m->d_rec;
  m->dev.collect_interval = collect_interval;
  m->series_nrec = 0;
  if (!m->rec_fields.empty()) {
    // ring of the recorded columns: field k owns rec_cap slots of stride_k bytes behind off_k
    const size_t slots = std::max<size_t>(std::max<size_t>((size_t)n_rec, 1), m->series_slots);
    size_t total = 0;
    for (auto& rf : m->rec_fields) { rf.off = total; total += rf.stride * slots; }
    static const size_t cap_bytes = (size_t)(getenv("JXB_SERIES_MAX_MB") ? atoll(getenv("JXB_SERIES_MAX_MB")) : 32768) << 20;
    if (total > cap_bytes)
      return fail(JXB_ERR_INVALID, "recording %d snapshots of the selected agent columns needs %.1f GB of HBM (limit JXB_SERIES_MAX_MB = %zu MB): "
                  "raise collect_interval or record fewer columns", n_rec, (double)total / 1e9, cap_bytes >> 20);
    if (total > m->series_bytes || m->series_slots != slots) {
      CK(cudaStreamSynchronize(s));
      pool_free(eng, m->d_series, m->series_bytes);
      m->d_series = nullptr; m->series_bytes = 0;
      CK(pool_alloc(eng, (void**)&m->d_series, total));
      m->series_bytes = total;
      m->series_slots = slots;
    }
    m->series_nrec = n_rec;
  }
  {
    // reset the per-run counters, keep the persistent ones
    int zeros[2] = {0, 0};
    CK(cudaMemcpyAsync(&m->dev.ctrl->step_in_run, zeros, sizeof(zeros), cudaMemcpyHostToDevice, s));
  }
  static const bool use_graph = getenv("JXB_NO_GRAPH") == nullptr;
  const bool persistent = m->desc.program == JXB_PROGRAM_SCHELLING && !m->grid_sharded;
  const bool graphs = use_graph && !persistent && !m->profile && m->dev.exchange != 2 && steps > 0 &&
                      !(m->traced && !m->traced_started);        // first step of a traced model uses variant 0
  if (graphs) {
    // kernel arguments (the ModelDev snapshot) are baked into a captured graph: rebuild the
    // two cached graphs (1 step, 32 steps) whenever a pointer or the interval changed
    const bool changed = !m->graph1 || m->sig_keys != m->d_keys || m->sig_metrics != m->d_metrics ||
                         m->sig_rec != m->d_rec || m->sig_ci != collect_interval || m->sig_series != m->d_series ||
                         m->sig_series_slots != m->series_slots;
    if (changed) {
      if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }
      if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
      int rc = build_graph(m, 1, &m->graph1);
      if (rc) return rc;
      m->chunkK = 32;
      rc = build_graph(m, m->chunkK, &m->graphK);
      if (rc) return rc;
      m->sig_keys = m->d_keys; m->sig_metrics = m->d_metrics; m->sig_rec = m->d_rec; m->sig_ci = collect_interval;
      m->sig_series = m->d_series; m->sig_series_slots = m->series_slots;
    }
  }
  for (auto e : m->prof_events) cudaEventDestroy(e);
  m->prof_events.clear();

  CK(cudaEventRecord(eng->ev0, s));
  if (persistent && m->rec_fields.empty()) {
    if (steps > 0) { int rc = launch_schelling(m, steps, s); if (rc) return rc; }
  } else if (persistent) {
    // per-agent series: the persistent kernel runs up to the next recording step, then the snapshots are taken
    long long t = t0;
    int left = steps;
    while (left > 0) {
      const int c = (int)std::min<long long>(left, collect_interval - (t % collect_interval));
      int rc = launch_schelling(m, c, s);
      if (!rc) rc = enqueue_snapshots(m, s);
      if (rc) return rc;
      t += c; left -= c;
    }
  } else if (graphs) {
    int left = steps;
    while (left >= m->chunkK) { CK(cudaGraphLaunch(m->graphK, s)); left -= m->chunkK; }
    while (left > 0) { CK(cudaGraphLaunch(m->graph1, s)); --left; }
    eng->launches += (int64_t)steps * launches_per_step_all(m);
  } else {
    for (int t = 0; t < steps; ++t) {
      int rc = enqueue_step(m, s, m->profile);
      if (rc) return rc;
      m->traced_started = true;
    }
  }
  CK(cudaEventRecord(eng->ev1, s));
  if (metrics_out && n_rec)
    CK(cudaMemcpyAsync(metrics_out, m->d_metrics, (size_t)n_rec * kMaxMetrics * sizeof(double),
                       cudaMemcpyDeviceToHost, s));
  if (steps_out && n_rec)
    CK(cudaMemcpyAsync(steps_out, m->d_rec, (size_t)n_rec * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
  if (device_seconds_out) *device_seconds_out = (double)ms * 1e-3;
  if (n_records_out) *n_records_out = n_rec;
  if (m->profile && persistent) {
    m->prof_seconds = (double)ms * 1e-3;      // the persistent kernel IS the run: time per step = total / steps
    m->prof_launches = steps;
  } else if (m->profile) {
    double tot = 0;
    for (size_t i = 0; i + 1 < m->prof_events.size(); i += 2) {
      float k = 0;
      cudaEventElapsedTime(&k, m->prof_events[i], m->prof_events[i + 1]);
      tot += k;
    }
    m->prof_seconds = tot * 1e-3;
    m->prof_launches = (int64_t)m->prof_events.size() / 2;
  }
  if (!m->gs_trace.empty()) {
    // per-kernel device time of the band steps, summed over the run and for the first steps
    static const char* names[5] = {"sweep", "counts", "moveout", "forward", "apply"};
    const size_t nst = m->gs_trace.size() / 6;
    double sum[5] = {0, 0, 0, 0, 0};
    std::string first;
    for (size_t st = 0; st < nst; ++st)
      for (int k = 0; k < 5; ++k) {
        float us = 0;
        cudaEventElapsedTime(&us, m->gs_trace[st * 6 + k], m->gs_trace[st * 6 + k + 1]);
        sum[k] += us * 1e3;
        if (st < 2) { char b[64]; snprintf(b, sizeof b, " %s[%zu]=%.1f", names[k], st, us * 1e3); first += b; }
      }
    fprintf(stderr, "[jxb gs_trace] rank %d, %zu steps, us: sweep %.1f counts %.1f moveout %.1f forward %.1f apply %.1f |%s\n",
            m->dev.rank, nst, sum[0], sum[1], sum[2], sum[3], sum[4], first.c_str());
    for (auto e : m->gs_trace) cudaEventDestroy(e);
    m->gs_trace.clear();
  }
  m->time_step = t0 + steps;
  if (m->has_grid && steps > 0) m->sat_dirty = true;
  if (m->net_sharded) {
    unsigned int nerr = 0;
    CK(cudaMemcpy(&nerr, &((SirXchgHdr*)m->ns_area)->err, sizeof(nerr), cudaMemcpyDeviceToHost));
    if (nerr) return fail(JXB_ERR_NCCL, "sharded Network: a rank did not publish its step within the spin budget");
  }
  if (m->grid_sharded) {
    if (steps > 0) { m->ct_stale = true; m->pos_stale = true; }
    unsigned int gerr = 0;
    CK(cudaMemcpy(&gerr, &((GridXchgHdr*)m->gs_area)->err, sizeof(gerr), cudaMemcpyDeviceToHost));
    if (gerr) return fail(JXB_ERR_NCCL, "sharded Grid: a rank did not publish its band within the spin budget");
  }
  if (m->dev.exchange == 1) {
    unsigned int xerr = 0;
    CK(cudaMemcpy(&xerr, &eng->xlocal->err, sizeof(xerr), cudaMemcpyDeviceToHost));
    if (xerr) return fail(JXB_ERR_NCCL, "peer exchange timed out: a rank did not publish its partial sums");
  }
  return JXB_OK;
}

// AgentCollection.update with the caller's key (agent.py:132-177); env/metrics untouched
extern "C" int jxb_collection_update(jxb_model* m, int type, uint32_t k0, uint32_t k1) {
  NEED(m); NEED_TYPE(m, type);
  if (!m->collections_ready[type])
    return fail(JXB_ERR_STATE, "Agent collection not initialized. Call init() first.");   // agent.py:150-151
  if (m->has_grid || m->has_net || m->has_eco || m->traced)
    return fail(JXB_ERR_UNSUPPORTED, "grid / network / economy / traced collections step through Model.step only");
  CK(cudaSetDevice(m->eng->device));
  const TypeDev& t = m->dev.t[type];
  const int blocks = (int)std::max<long long>(1, std::min<long long>((t.n / kVec + kThreads - 1) / kThreads,
                                                                     (long long)m->eng->sms * 8));
  (void)k0; (void)k1;   // registered well-mixed rules draw nothing in update(); the key is unobservable
  if (m->desc.rng_mode == JXB_RNG_PARTITIONABLE)
    collection_update_kernel<1><<<blocks, kThreads, 0, m->eng->stream>>>(m->dev, type);
  else
    collection_update_kernel<0><<<blocks, kThreads, 0, m->eng->stream>>>(m->dev, type);
  m->eng->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->eng->stream));
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// record / select (csrc/record.cuh)
// ---------------------------------------------------------------------------------------
extern "C" int jxb_model_record_fields(jxb_model* m, int n, const int32_t* types, const int32_t* fields) {
  NEED(m);
  if (n < 0 || n > 8 || (n && (!types || !fields))) return fail(JXB_ERR_INVALID, "record up to 8 (collection, field) columns");
  std::vector<jxb_model::RecField> rf;
  for (int k = 0; k < n; ++k) {
    NEED_TYPE(m, types[k]);
    if (fields[k] < 0 || fields[k] >= m->rules[types[k]]->nf) return fail(JXB_ERR_INVALID, "field %d out of range", fields[k]);
    if (m->grid_sharded && m->dev.world_size > 1)
      return fail(JXB_ERR_UNSUPPORTED, "per-agent series of a Grid sharded over ranks: record on the single-GPU model");
    const size_t bytes = field_bytes(m, types[k], fields[k]);
    rf.push_back({types[k], fields[k], bytes, (bytes + 15) / 16 * 16, 0});
  }
  CK(cudaSetDevice(m->eng->device));
  CK(cudaStreamSynchronize(m->eng->stream));
  m->rec_fields = rf;
  m->series_nrec = 0;
  m->series_slots = 0;
  if (m->graph1) { cudaGraphExecDestroy(m->graph1); m->graph1 = nullptr; }     // the snapshot launches are part of the step graph
  if (m->graphK) { cudaGraphExecDestroy(m->graphK); m->graphK = nullptr; }
  return JXB_OK;
}

extern "C" int jxb_model_series_info(jxb_model* m, int k, int* n_records, size_t* bytes_per_record) {
  NEED(m);
  if (k < 0 || k >= (int)m->rec_fields.size()) return fail(JXB_ERR_INVALID, "series %d out of range", k);
  if (n_records) *n_records = m->series_nrec;
  if (bytes_per_record) *bytes_per_record = m->rec_fields[k].bytes;
  return JXB_OK;
}

extern "C" int jxb_model_series_download(jxb_model* m, int k, void* host, size_t bytes) {
  NEED(m);
  if (k < 0 || k >= (int)m->rec_fields.size()) return fail(JXB_ERR_INVALID, "series %d out of range", k);
  const auto& rf = m->rec_fields[k];
  if (bytes != rf.bytes * (size_t)m->series_nrec)
    return fail(JXB_ERR_INVALID, "series %d holds %d records of %zu bytes", k, m->series_nrec, rf.bytes);
  if (!bytes) return JXB_OK;
  CK(cudaSetDevice(m->eng->device));
  CK(cudaMemcpy2DAsync(host, rf.bytes, m->d_series + rf.off, rf.stride, rf.bytes, (size_t)m->series_nrec,
                       cudaMemcpyDeviceToHost, m->eng->stream));
  CK(cudaStreamSynchronize(m->eng->stream));
  return JXB_OK;
}

static int pred_check(jxb_model* m, int type, const jxb_pred_ins* prog, int n_ins) {
  if (n_ins < 1 || n_ins > kPredMaxIns) return fail(JXB_ERR_INVALID, "predicate programs hold 1..%d instructions", kPredMaxIns);
  int sp = 0;
  for (int k = 0; k < n_ins; ++k) {
    const jxb_pred_ins& in = prog[k];
    int pop = 0, push = 1;
    if (in.op == JP_LOAD_F32 || in.op == JP_LOAD_I32 || in.op == JP_LOAD_U8) {
      if (in.a < 0 || in.a >= m->rules[type]->nf) return fail(JXB_ERR_INVALID, "predicate: field %d out of range", in.a);
      const FieldSpec& f = m->rules[type]->f[in.a];
      const int want = in.op == JP_LOAD_F32 ? 0 : (in.op == JP_LOAD_I32 ? 1 : 2);
      if (f.dtype != want || in.b < 0 || in.b >= f.width)
        return fail(JXB_ERR_INVALID, "predicate: load of '%s' does not match its dtype / width", f.name);
    } else if (in.op == JP_CONST_F32 || in.op == JP_CONST_I32) {
    } else if ((in.op >= JP_ADD_F && in.op <= JP_MAX_I) || (in.op >= JP_LT_F && in.op <= JP_XOR)) {
      pop = 2;
    } else if ((in.op >= JP_NEG_F && in.op <= JP_F2B) || in.op == JP_NOT || (in.op >= JP_SQRT_F && in.op <= JP_LOG_F)) {
      pop = 1;
    } else if (in.op == JP_SELECT) {
      pop = 3;
    } else {
      return fail(JXB_ERR_INVALID, "predicate: unknown opcode %d", in.op);
    }
    if (sp < pop) return fail(JXB_ERR_INVALID, "predicate: stack underflow at instruction %d", k);
    sp += push - pop;
    if (sp > kPredStack) return fail(JXB_ERR_INVALID, "predicate: expression deeper than %d", kPredStack);
  }
  if (sp != 1) return fail(JXB_ERR_INVALID, "predicate: program leaves %d values", sp);
  return JXB_OK;
}

extern "C" int jxb_collection_filter_select(jxb_model* m, int type, const jxb_pred_ins* prog, int n_ins,
                                            const uint8_t* host_mask, size_t mask_bytes, int64_t* count_out) {
  NEED(m); NEED_TYPE(m, type);
  if (!count_out) return fail(JXB_ERR_INVALID, "count_out is NULL");
  if (!m->collections_ready[type] && !m->initialized) return fail(JXB_ERR_STATE, "Agent collection not initialized. Call init() first.");
  const long long n = m->desc.types[type].n_agents;
  if ((prog != nullptr) == (host_mask != nullptr)) return fail(JXB_ERR_INVALID, "give either a predicate program or a host mask");
  if (host_mask && mask_bytes != (size_t)n) return fail(JXB_ERR_INVALID, "the mask holds one byte per agent (%lld)", n);
  if (prog) { int rc = pred_check(m, type, prog, n_ins); if (rc) return rc; }
  if (m->grid_sharded && m->dev.world_size > 1) return fail(JXB_ERR_UNSUPPORTED, "filter on a Grid sharded over ranks");
  jxb_engine* eng = m->eng;
  CK(cudaSetDevice(eng->device));
  cudaStream_t s = eng->stream;
  PredProgram pg;
  memset(&pg, 0, sizeof(pg));
  FilterCols fc;
  memset(&fc, 0, sizeof(fc));
  const RuleSpec* rs = m->rules[type];
  for (int f = 0; f < rs->nf; ++f) { fc.f[f] = m->dev.t[type].f[f]; fc.width[f] = rs->f[f].width; }
  if (prog) {
    pg.n = n_ins;
    memcpy(pg.ins, prog, (size_t)n_ins * sizeof(jxb_pred_ins));
    for (int k = 0; k < n_ins; ++k)
      if (prog[k].op == JP_LOAD_F32 || prog[k].op == JP_LOAD_I32 || prog[k].op == JP_LOAD_U8) {
        int rc = materialize_field(m, type, prog[k].a, s, false);
        if (rc) return rc;
      }
  }
  pool_free(eng, m->filter_pos, m->filter_pos_bytes);
  m->filter_pos = nullptr; m->filter_pos_bytes = 0; m->filter_type = -1; m->filter_count = -1;
  const int ntiles = (int)((n + kScanTile - 1) / kScanTile);
  const size_t pos_bytes = ((size_t)n + 2) * 4, sums_bytes = ((size_t)ntiles + 2) * 4;
  unsigned int* d_flags = nullptr; unsigned int* d_sums = nullptr; unsigned char* d_mask = nullptr;
  CK(pool_alloc(eng, (void**)&m->filter_pos, pos_bytes));
  m->filter_pos_bytes = pos_bytes;
  CK(pool_alloc(eng, (void**)&d_flags, pos_bytes));
  CK(pool_alloc(eng, (void**)&d_sums, sums_bytes));
  if (host_mask) {
    CK(pool_alloc(eng, (void**)&d_mask, (size_t)n));
    CK(cudaMemcpyAsync(d_mask, host_mask, (size_t)n, cudaMemcpyHostToDevice, s));
  }
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)eng->sms * 8));
  filter_flags_kernel<<<blocks, kThreads, 0, s>>>(pg, fc, d_mask, n, d_flags);
  scan_tile_sums_kernel<<<ntiles, kThreads, 0, s>>>(d_flags, n, d_sums);
  scan_sums_kernel<<<1, 1024, 0, s>>>(d_sums, ntiles);
  scan_apply_kernel<<<ntiles, kThreads, 0, s>>>(d_flags, n, d_sums, m->filter_pos, nullptr);
  eng->launches += 4;
  unsigned int count = 0;
  cudaError_t e = cudaMemcpyAsync(&count, m->filter_pos + n, sizeof(count), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaGetLastError();
  pool_free(eng, d_flags, pos_bytes); pool_free(eng, d_sums, sums_bytes);
  if (d_mask) pool_free(eng, d_mask, (size_t)n);
  if (e != cudaSuccess) return fail(JXB_ERR_CUDA, "filter select failed: %s", cudaGetErrorString(e));
  m->filter_type = type;
  m->filter_count = count;
  *count_out = count;
  return JXB_OK;
}

extern "C" int jxb_collection_filter_gather(jxb_model* src, int type, jxb_model* dst, int dst_type) {
  NEED(src); NEED(dst); NEED_TYPE(src, type); NEED_TYPE(dst, dst_type);
  if (src->filter_type != type || src->filter_count < 0 || !src->filter_pos)
    return fail(JXB_ERR_STATE, "call jxb_collection_filter_select on this collection first");
  if (src->eng != dst->eng) return fail(JXB_ERR_INVALID, "source and destination live on different engines");
  const RuleSpec* a = src->rules[type]; const RuleSpec* b = dst->rules[dst_type];
  bool same = a->nf == b->nf;
  for (int f = 0; same && f < a->nf; ++f) same = a->f[f].dtype == b->f[f].dtype && a->f[f].width == b->f[f].width;
  if (!same) return fail(JXB_ERR_INVALID, "the destination collection has a different state layout");
  if (dst->desc.types[dst_type].n_agents != src->filter_count)
    return fail(JXB_ERR_INVALID, "the destination collection must hold exactly the %lld selected agents", src->filter_count);
  jxb_engine* eng = src->eng;
  CK(cudaSetDevice(eng->device));
  cudaStream_t s = eng->stream;
  const long long n = src->desc.types[type].n_agents;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)eng->sms * 8));
  for (int f = 0; f < a->nf; ++f) {
    int rc = materialize_field(src, type, f, s, false);
    if (rc) return rc;
    const int elem = a->f[f].width * (int)dtype_size(a->f[f].dtype);
    filter_scatter_kernel<<<blocks, kThreads, 0, s>>>(src->filter_pos, n, (const unsigned char*)src->dev.t[type].f[f],
                                                      (unsigned char*)dst->dev.t[dst_type].f[f], elem);
    eng->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  pool_free(eng, src->filter_pos, src->filter_pos_bytes);
  src->filter_pos = nullptr; src->filter_pos_bytes = 0; src->filter_type = -1; src->filter_count = -1;
  dst->collections_ready[dst_type] = true;
  if (dst->has_net) return sir_sync_from_api(dst);
  if (dst->has_grid) dst->grid_built = false;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// page-locked host blocks (cached by size)
// ---------------------------------------------------------------------------------------
static std::mutex g_host_mu;
static std::multimap<size_t, void*> g_host_free;        // size -> cached block
static std::map<void*, size_t> g_host_live;             // block -> size
static size_t g_host_cached = 0;

extern "C" int jxb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(JXB_ERR_INVALID, "out is NULL");
  bytes = (std::max<size_t>(bytes, 64) + 4095) & ~(size_t)4095;
  std::lock_guard<std::mutex> lk(g_host_mu);
  auto it = g_host_free.lower_bound(bytes);
  if (it != g_host_free.end() && it->first <= bytes + bytes / 8 + (1u << 16)) {
    *out = it->second;
    g_host_live[it->second] = it->first;
    g_host_cached -= it->first;
    g_host_free.erase(it);
    return JXB_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    for (auto& kv : g_host_free) cudaFreeHost(kv.second);
    g_host_free.clear();
    g_host_cached = 0;
    e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) return fail(JXB_ERR_CUDA, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
  }
  g_host_live[p] = bytes;
  *out = p;
  return JXB_OK;
}

extern "C" int jxb_host_free(void* p) {
  if (!p) return JXB_OK;
  std::lock_guard<std::mutex> lk(g_host_mu);
  auto it = g_host_live.find(p);
  if (it == g_host_live.end()) return fail(JXB_ERR_INVALID, "not a block of jxb_host_alloc");
  const size_t bytes = it->second;
  g_host_live.erase(it);
  if (g_host_cached + bytes > ((size_t)4 << 30)) { cudaFreeHost(p); return JXB_OK; }
  g_host_free.emplace(bytes, p);
  g_host_cached += bytes;
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// ensembles
// ---------------------------------------------------------------------------------------
extern "C" int jxb_ensemble_run(jxb_engine* eng, const jxb_model_desc* d, int R, int n_swept,
                                const int32_t* slots, const double* params, const uint32_t* seeds,
                                const double* env_init, int steps, double* last_metrics_out,
                                double* device_seconds_out) {
  if (!eng) return fail(JXB_ERR_INVALID, "engine is NULL");
  int rc = validate_desc(d);
  if (rc) return rc;
  if (R <= 0 || steps < 0 || !seeds || !last_metrics_out) return fail(JXB_ERR_INVALID, "bad ensemble arguments");
  if (n_swept < 0 || n_swept > kEnsMaxSwept || (n_swept && (!slots || !params)))
    return fail(JXB_ERR_INVALID, "bad swept-parameter table");
  if (d->program == JXB_PROGRAM_SCHELLING || d->program == JXB_PROGRAM_SIR)
    return fail(JXB_ERR_UNSUPPORTED, "grid / network programs run replicas through jxb_model_run");
  CK(cudaSetDevice(eng->device));
  cudaStream_t s = eng->stream;
  const ProgramSpec* prog = find_program(d->program);
  EnsDev ed;
  memset(&ed, 0, sizeof(ed));
  ed.program = d->program; ed.n_types = d->n_types; ed.has_env_fn = prog->has_env_fn;
  ed.n_env = prog->n_env; ed.n_metrics = prog->n_metrics;
  ed.R = R; ed.steps = steps; ed.n_swept = n_swept;
  for (int i = 0; i < n_swept; ++i) {
    ed.slots[i] = slots[i];
    const int sl = slots[i];
    const bool ok = (sl >= 0 && sl < JXB_MAX_PARAMS) ||
                    (sl >= 100 && (sl - 100) / 16 < d->n_types);
    if (!ok) return fail(JXB_ERR_INVALID, "swept slot %d is not a parameter of this model", sl);
  }
  for (int k = 0; k < JXB_MAX_PARAMS; ++k) ed.mp[k] = k < d->n_params ? d->params[k] : 0.0;
  for (int k = 0; k < prog->n_env; ++k) ed.env0[k] = env_init ? env_init[k] : prog->env[k].dflt;
  // state layout of ONE CTA for a cluster of `cs` CTAs per replica: every collection is sliced by
  // agent index (slice = multiple of 4 agents); pick the smallest cluster whose slice fits shared memory
  auto layout = [&](int cs, long long* slice) -> size_t {
    size_t off = 0;
    for (int i = 0; i < d->n_types; ++i) {
      const long long n = d->types[i].n_agents;
      slice[i] = cs == 1 ? n : ((n + cs - 1) / cs + 3) / 4 * 4;
      const RuleSpec* rs = find_rule(d->types[i].rule);
      for (int f = 0; f < rs->nf; ++f) {
        ed.t[i].f[f] = (void*)off;
        size_t bytes = (size_t)(slice[i] + 8) * rs->f[f].width * dtype_size(rs->f[f].dtype);
        off += (bytes + 15) / 16 * 16;
      }
    }
    return off;
  };
  for (int i = 0; i < d->n_types; ++i) fill_type_dev(d->types[i], ed.t[i]);
  // JXB_ENS_MAX_CLUSTER / JXB_ENS_MIN_CLUSTER bound the cluster size (read per call: the parity tests force
  // every shape -- 1, 2, 4, 8 CTAs per replica and the L2-scratch path -- on the same replicas);
  // JXB_ENS_FORCE_SCRATCH=1 puts the state into the per-CTA L2 slot even when it would fit shared memory
  const int max_cluster = getenv("JXB_ENS_MAX_CLUSTER") ? atoi(getenv("JXB_ENS_MAX_CLUSTER")) : 8;
  const int min_cluster = getenv("JXB_ENS_MIN_CLUSTER") ? std::max(1, atoi(getenv("JXB_ENS_MIN_CLUSTER"))) : 1;
  const bool force_scratch = getenv("JXB_ENS_FORCE_SCRATCH") && atoi(getenv("JXB_ENS_FORCE_SCRATCH")) != 0;
  int cs = 1;
  size_t off = layout(1, ed.slice);
  if ((off > kEnsSmemBudget || min_cluster > 1) && !force_scratch) {
    bool found = false;
    for (int c = 2; c <= std::min(max_cluster, (int)kMaxPeers); c *= 2) {
      if (c < min_cluster) continue;
      long long sl[JXB_MAX_TYPES];
      const size_t o = layout(c, sl);
      if (o <= kEnsSmemBudget) { cs = c; off = o; found = true; for (int i = 0; i < d->n_types; ++i) ed.slice[i] = sl[i]; break; }
    }
    if (!found) off = layout(1, ed.slice);     // nothing fits: per-CTA L2 scratch slot
  }
  ed.cluster = cs;
  ed.state_bytes = off;
  ed.use_smem = off <= kEnsSmemBudget && !force_scratch;
  int grid;
  size_t dyn = 0;
  if (ed.use_smem) {
    dyn = off;
    const int per_sm = (cs == 1 && off <= kEnsSmemBudget / 2) ? 2 : 1;
    const int clusters = std::max(1, std::min(R, eng->sms * per_sm / cs));
    grid = clusters * cs;
  } else {
    // keep all live replica slots inside L2 (~64 MB) when possible, at least one CTA per SM
    long long fit = (long long)((64ull << 20) / off);
    grid = (int)std::min<long long>(R, std::max<long long>(eng->sms, std::min<long long>(fit, 2ll * eng->sms)));
  }
  double* d_params = nullptr; uint32_t* d_seeds = nullptr; double* d_out = nullptr; unsigned char* d_scratch = nullptr;
  auto cleanup = [&]() { cudaFree(d_params); cudaFree(d_seeds); cudaFree(d_out); cudaFree(d_scratch); };
#define ECK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return fail(JXB_ERR_CUDA, \
    "%s failed: %s", #call, cudaGetErrorString(_e)); } } while (0)
  ECK(cudaMalloc(&d_params, std::max<size_t>((size_t)R * std::max(n_swept, 1) * sizeof(double), 16)));
  ECK(cudaMalloc(&d_seeds, (size_t)R * 4));
  ECK(cudaMalloc(&d_out, (size_t)R * kMaxMetrics * sizeof(double)));
  if (!ed.use_smem) ECK(cudaMalloc(&d_scratch, (size_t)grid * off));
  if (n_swept) ECK(cudaMemcpyAsync(d_params, params, (size_t)R * n_swept * sizeof(double), cudaMemcpyHostToDevice, s));
  ECK(cudaMemcpyAsync(d_seeds, seeds, (size_t)R * 4, cudaMemcpyHostToDevice, s));
  ed.params = d_params; ed.seeds = d_seeds; ed.out = d_out; ed.scratch = d_scratch;
  const bool part = d->rng_mode == JXB_RNG_PARTITIONABLE;
  if (dyn > 48 * 1024) {
    if (part) ECK(cudaFuncSetAttribute(ensemble_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    else ECK(cudaFuncSetAttribute(ensemble_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  }
  ECK(cudaEventRecord(eng->ev0, s));
  if (cs > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kEnsThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (part) ECK(cudaLaunchKernelEx(&cfg, ensemble_kernel<1>, ed));
    else ECK(cudaLaunchKernelEx(&cfg, ensemble_kernel<0>, ed));
  } else {
    if (part) ensemble_kernel<1><<<grid, kEnsThreads, dyn, s>>>(ed);
    else ensemble_kernel<0><<<grid, kEnsThreads, dyn, s>>>(ed);
  }
  eng->launches++;
  ECK(cudaGetLastError());
  ECK(cudaEventRecord(eng->ev1, s));
  std::vector<double> tmp((size_t)R * kMaxMetrics);
  ECK(cudaMemcpyAsync(tmp.data(), d_out, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  ECK(cudaStreamSynchronize(s));
  float ms = 0;
  ECK(cudaEventElapsedTime(&ms, eng->ev0, eng->ev1));
#undef ECK
  if (device_seconds_out) *device_seconds_out = ms * 1e-3;
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < prog->n_metrics; ++k)
      last_metrics_out[(size_t)r * prog->n_metrics + k] = tmp[(size_t)r * kMaxMetrics + k];
  cleanup();
  return JXB_OK;
}

// ---------------------------------------------------------------------------------------
// host-only key algebra
// ---------------------------------------------------------------------------------------
static bool mode_ok(int mode) { return mode == JXB_RNG_LEGACY || mode == JXB_RNG_PARTITIONABLE; }

extern "C" int jxb_prng_threefry2x32(const uint32_t key[2], const uint32_t ctr[2], uint32_t out[2]) {
  threefry2x32(key[0], key[1], ctr[0], ctr[1], out[0], out[1]);
  return JXB_OK;
}

extern "C" int jxb_prng_split(int mode, const uint32_t key[2], int n, uint32_t* out) {
  if (!mode_ok(mode) || n < 0 || !out) return fail(JXB_ERR_INVALID, "bad split arguments");
  for (int j = 0; j < n; ++j) {
    Key c = split_child(mode, Key{key[0], key[1]}, (uint64_t)j, (uint64_t)n);
    out[2 * j] = c.a; out[2 * j + 1] = c.b;
  }
  return JXB_OK;
}

static uint32_t host_bits(int mode, Key k, uint64_t j, uint64_t n) {
  return mode == 1 ? bits_elem<1>(k, j, n) : bits_elem<0>(k, j, n);
}

extern "C" int jxb_prng_bits(int mode, const uint32_t key[2], int64_t n, uint32_t* out) {
  if (!mode_ok(mode) || n < 0 || !out) return fail(JXB_ERR_INVALID, "bad bits arguments");
  for (int64_t j = 0; j < n; ++j) out[j] = host_bits(mode, Key{key[0], key[1]}, (uint64_t)j, (uint64_t)n);
  return JXB_OK;
}

extern "C" int jxb_prng_uniform(int mode, const uint32_t key[2], int64_t n, float lo, float hi, float* out) {
  if (!mode_ok(mode) || n < 0 || !out) return fail(JXB_ERR_INVALID, "bad uniform arguments");
  for (int64_t j = 0; j < n; ++j)
    out[j] = bits_to_uniform(host_bits(mode, Key{key[0], key[1]}, (uint64_t)j, (uint64_t)n), lo, hi);
  return JXB_OK;
}

extern "C" int jxb_prng_feistel(uint32_t n, const uint32_t rk[4], uint32_t idx, int inverse, uint32_t* out) {
  if (!rk || !out || (n > 1 && idx >= n)) return fail(JXB_ERR_INVALID, "bad feistel arguments");
  const Feistel f = make_feistel(n, rk);
  *out = inverse ? feistel_inverse(f, idx) : feistel_permute(f, idx);
  return JXB_OK;
}

extern "C" int jxb_prng_randint(int mode, const uint32_t key[2], int64_t n, int32_t lo, int32_t hi, int32_t* out) {
  if (!mode_ok(mode) || n < 0 || !out) return fail(JXB_ERR_INVALID, "bad randint arguments");
  const Key k{key[0], key[1]};
  const Key k1 = split_child(mode, k, 0, 2), k2 = split_child(mode, k, 1, 2);
  uint32_t span = hi <= lo ? 1u : (uint32_t)((int64_t)hi - (int64_t)lo);
  uint32_t mult = (1u << 16) % span;
  mult = (uint32_t)(mult * mult) % span;
  for (int64_t j = 0; j < n; ++j) {
    const uint32_t hb = host_bits(mode, k1, (uint64_t)j, (uint64_t)n);
    const uint32_t lb = host_bits(mode, k2, (uint64_t)j, (uint64_t)n);
    const uint32_t off = (uint32_t)((hb % span) * mult + (lb % span)) % span;
    out[j] = (int32_t)((int64_t)lo + (int64_t)off);
  }
  return JXB_OK;
}
