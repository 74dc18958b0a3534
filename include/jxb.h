/*
 * jxb.h -- C ABI of libjxb.so, the B200-native execution engine for the JaxABM hot path.
 *
 * The reference (a11to1n3/JaxABM) is pure Python on JAX and has no FFI of its own
 * (SURVEY.md F1); this header therefore *defines* the boundary.  Every entry point
 * names the reference interface it stands in for (path:line under /root/reference).
 * The host side (the jaxabm_b200 Python package) keeps the reference's Python API verbatim and
 * binds these symbols with ctypes (see INTEGRATION.md for the stub a maintainer of
 * the reference would add).
 *
 * Conventions
 *   - plain C types only; all pointers are HOST pointers owned by the caller and are
 *     only touched for the duration of the call;
 *   - every function returns JXB_OK (0) or a negative jxb_status; jxb_last_error()
 *     returns a thread-local NUL-terminated message owned by the library;
 *   - the engine owns all device memory, its CUDA stream, events and graphs;
 *   - calls on one jxb_model must be serialised by the caller;
 *   - there is no CPU fallback: any call that needs a device fails with
 *     JXB_ERR_NO_DEVICE when none is present.  The jxb_prng_* helpers are host-only
 *     scalar key algebra (the reference does the same work on the host through
 *     jax.random) and work without a GPU.
 */
#ifndef JXB_H_
#define JXB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JXB_VERSION 100 /* 0.1.0 */

typedef enum {
  JXB_OK = 0,
  JXB_ERR_INVALID = -1,    /* bad argument / unknown rule or program          */
  JXB_ERR_NO_DEVICE = -2,  /* no CUDA device (there is no CPU fallback)        */
  JXB_ERR_CUDA = -3,       /* a CUDA runtime call failed; see jxb_last_error() */
  JXB_ERR_STATE = -4,      /* call order violated (e.g. run before init)       */
  JXB_ERR_NCCL = -5,
  JXB_ERR_UNSUPPORTED = -6
} jxb_status;

/* JAX threefry stream layouts (flag jax_threefry_partitionable; default flipped in
 * JAX 0.5.0).  The reference pins only jax>=0.4.1 (requirements.txt:8-9).          */
enum { JXB_RNG_LEGACY = 0, JXB_RNG_PARTITIONABLE = 1 };

/* Registered agent rules: one hand-written fused update kernel each.               */
enum {
  JXB_RULE_RANDOM_WALKER = 1,   /* examples/basic_example.py:20-69                   */
  JXB_RULE_SCALED_WALKER = 2,   /* same update; keyed init (roofline variant, N=2^26)*/
  JXB_RULE_CONSUMER = 3,        /* tests/integration/test_integration.py:20-67       */
  JXB_RULE_PRODUCER = 4,        /* tests/integration/test_integration.py:70-121      */
  JXB_RULE_GROWTH = 5,          /* tests/unit/test_analysis.py:22-39 (DummyAgent)    */
  JXB_RULE_INCREMENT = 6,       /* tests/unit/test_model.py:43-48   (DummyAgent)     */
  JXB_RULE_WEALTH = 7,          /* tests/unit/test_agent.py:44-100  (TestAgent)      */
  JXB_RULE_SCHELLING = 8,       /* layout examples/models/schelling_model.py:26-31   */
  JXB_RULE_SIR = 9,             /* layout jaxabm/agentpy.py:557,574-582              */
  JXB_RULE_HOUSEHOLD = 10,      /* examples/models/advanced_economic_model.py:57-296 */
  JXB_RULE_CONSUMER_FIRM = 11,  /* examples/models/advanced_economic_model.py:299-581*/
  JXB_RULE_TRACED = 12          /* user AgentType traced into a generated kernel (jxb_model_create_traced) */
};

/* Registered model programs: the update_state_fn + metrics_fn pair that runs inside
 * the step (jaxabm/model.py:182-200), fused into the tail of the step's kernels.    */
enum {
  JXB_PROGRAM_NONE = 0,         /* no env fn, no metrics (run() returns {})          */
  JXB_PROGRAM_RANDOM_WALK = 1,  /* examples/basic_example.py:72-182                  */
  JXB_PROGRAM_MARKET = 2,       /* tests/integration/test_integration.py:125-183     */
  JXB_PROGRAM_GROWTH = 3,       /* tests/unit/test_analysis.py:105-128               */
  JXB_PROGRAM_COUNTER = 4,      /* tests/unit/test_model.py:20-40                    */
  JXB_PROGRAM_SCHELLING = 5,    /* builder-authored rule, DESIGN.md                  */
  JXB_PROGRAM_SIR = 6,          /* builder-authored rule, DESIGN.md                  */
  JXB_PROGRAM_ECONOMY = 7,      /* examples/models/advanced_economic_model.py:1461-1905 */
  JXB_PROGRAM_TRACED = 8        /* user update_state_fn / metrics_fn traced into the generated kernel's tail */
};

#define JXB_MAX_TYPES 4
#define JXB_MAX_PARAMS 16
#define JXB_MAX_METRICS 32      /* row stride (doubles) of metrics_out / last_metrics_out scratch */

/* One agent collection (jaxabm/agent.py:69-90: AgentCollection(agent_type, num_agents)). */
typedef struct {
  int32_t rule;                   /* JXB_RULE_*                                      */
  int64_t n_agents;               /* agents held by THIS engine (local shard)        */
  int64_t global_offset;          /* index of local agent 0 in the whole population  */
  int64_t global_n;               /* whole-population size (== n_agents on one GPU)  */
  int32_t n_params;
  float params[JXB_MAX_PARAMS];   /* rule constants (the agent_type's attributes)    */
} jxb_type_desc;

/* jaxabm/model.py:25-58 (Model.__init__) + add_agent_collection/add_env_state.      */
typedef struct {
  int32_t program;                /* JXB_PROGRAM_*                                   */
  int32_t rng_mode;               /* JXB_RNG_*                                       */
  int32_t n_types;
  jxb_type_desc types[JXB_MAX_TYPES];   /* insertion order == key order (model.py:163) */
  int32_t n_params;
  double params[JXB_MAX_PARAMS];  /* Model(params=...) entries the program reads     */
  int32_t grid_w, grid_h;         /* jaxabm/agentpy.py:480-493 Grid(shape, periodic) */
  int32_t grid_periodic;
  int32_t world_size, rank;       /* population sharding (1, 0 on a single GPU)      */
} jxb_model_desc;

typedef struct jxb_engine jxb_engine;
typedef struct jxb_model jxb_model;

int jxb_version(void);
const char* jxb_last_error(void);

/* ---- engine: one per process, bound to one device ------------------------------ */
int jxb_engine_create(int device, jxb_engine** out);
int jxb_engine_destroy(jxb_engine*);
int jxb_engine_sm_count(jxb_engine*, int* out);
/* kernels launched by this engine since creation (bench.py's "gpu_launches").       */
int jxb_engine_launch_count(jxb_engine*, int64_t* out);

/* ---- model ------------------------------------------------------------------------ */
/* jaxabm/model.py:25-58,60-99: construct with collections, env layout and params.   */
int jxb_model_create(jxb_engine*, const jxb_model_desc*, jxb_model** out);
int jxb_model_destroy(jxb_model*);

/* Field/env introspection so the host shim can size buffers (names mirror the
 * reference's state-dict keys).  dtype: 0=f32, 1=i32, 2=bool(u8), 3=f64 (Python float). */
int jxb_model_n_fields(jxb_model*, int type, int* out);
int jxb_model_field_info(jxb_model*, int type, int field, const char** name,
                         int* dtype, int* width);
int jxb_model_n_env(jxb_model*, int* out);
int jxb_model_env_info(jxb_model*, int slot, const char** name, int* dtype);
int jxb_model_n_metrics(jxb_model*, int* out);
int jxb_model_metric_info(jxb_model*, int metric, const char** name, int* dtype);

/* Traced models (jaxabm/agent.py:168-177 evaluates arbitrary user Python under jax.vmap; here the
 * host shim traces it -- jaxabm_b200/trace.py -- into a CUDA source compiled for sm_100a).  The
 * spec gives the state / env / metric layout the trace found and the two launcher entry points of
 * the generated library:
 *   int launch_init(const void* model_dev, int type, unsigned k0, unsigned k1, int rng_mode, int blocks, void* stream)
 *   int launch_step(const void* model_dev, int rng_mode, int variant, void* stream)
 * desc->program must be JXB_PROGRAM_TRACED and every collection's rule JXB_RULE_TRACED.  variant 0 is
 * used for the first step after construction (env entries still have their Python types), the
 * last variant for every later step.                                                        */
#define JXB_MAX_FIELDS 20
#define JXB_MAX_ENV 32
typedef struct {
  int32_t n_fields[JXB_MAX_TYPES];
  const char* field_names[JXB_MAX_TYPES][JXB_MAX_FIELDS];
  int32_t field_dtypes[JXB_MAX_TYPES][JXB_MAX_FIELDS];     /* 0 f32, 1 i32, 2 bool */
  int32_t n_env;
  const char* env_names[JXB_MAX_ENV];
  int32_t env_dtypes[JXB_MAX_ENV];                         /* 0 f32, 1 i32, 2 bool, 3 f64 (Python float) */
  double env_init[JXB_MAX_ENV];
  int32_t n_metrics;
  const char* metric_names[JXB_MAX_METRICS];
  int32_t metric_dtypes[JXB_MAX_METRICS];
  int32_t has_env_fn;
  int32_t n_acc;                                           /* reduction slots per CTA partial row */
  int32_t n_variants;
  void* launch_init;
  void* launch_step;
  int32_t n_consts;                                        /* float constants of the traced code ...      */
  const double* consts;                                    /* ... uploaded once; the kernels read them     */
  int32_t field_widths[JXB_MAX_TYPES][JXB_MAX_FIELDS];     /* components per agent (1; 2 for an f32[N,2] position ...) */
} jxb_traced_spec;
int jxb_model_create_traced(jxb_engine*, const jxb_model_desc*, const jxb_traced_spec*, jxb_model** out);

/* AgentCollection._states access (jaxabm/agent.py:179-196), host <-> HBM.           */
int jxb_model_upload(jxb_model*, int type, int field, const void* host, size_t bytes);
int jxb_model_download(jxb_model*, int type, int field, void* host, size_t bytes);
/* vmap broadcast of an unbatched init value (jaxabm/agent.py:125-130).              */
int jxb_model_fill(jxb_model*, int type, int field, const void* value, size_t bytes);
/* Model.add_env_state (jaxabm/model.py:76-99); scalars travel as double.            */
int jxb_model_set_env(jxb_model*, int slot, double value);
int jxb_model_get_env(jxb_model*, int slot, double* value);
/* change one constant of a collection's agent_type after construction (e.g. the
 * 'wage_rate' a caller passes in model_state, tests/unit/test_agent.py:75-78).     */
int jxb_model_set_type_param(jxb_model*, int type, int index, float value);

/* Network env (jaxabm/agentpy.py:545-582): directed adjacency list int32[E,2]; the
 * engine bins it by source into CSR in HBM.  Grid env for Schelling is derived from
 * the uploaded 'position'/'type' fields by jxb_model_grid_rebuild.                  */
int jxb_model_set_network(jxb_model*, const int32_t* edges, int64_t n_edges);
int jxb_model_grid_rebuild(jxb_model*);
/* env['grid'] (int32[W,H], -1 empty) as the reference lays it out
 * (examples/models/schelling_model.py:119-131).                                     */
int jxb_model_download_grid(jxb_model*, int32_t* host, size_t bytes);
/* env['empty_cells'] (examples/models/schelling_model.py:133-139) as cell ids x*H+y in slot
 * order, int32[W*H - n_agents].                                                          */
int jxb_model_download_empty_cells(jxb_model*, int32_t* host, size_t bytes);

/* Model.initialize (jaxabm/model.py:118-144): keys = split(PRNGKey(seed), C+1);
 * collection i is initialised on the device from keys[i+1] (agent.py:92-130).       */
int jxb_model_init(jxb_model*, uint32_t seed_key0, uint32_t seed_key1);
/* AgentCollection.init(key, cfg) for one collection (jaxabm/agent.py:92-130).       */
int jxb_collection_init(jxb_model*, int type, uint32_t key0, uint32_t key1);
/* AgentCollection.update(model_state, key, cfg) (jaxabm/agent.py:132-177): one fused
 * update of one collection with the caller's key; env/metrics untouched.            */
int jxb_collection_update(jxb_model*, int type, uint32_t key0, uint32_t key1);

/* Model.run(steps) (jaxabm/model.py:218-262) = `steps` x Model.step (model.py:146-216)
 * with no host round-trip inside.  metrics_out: [n_records][n_metrics] doubles where
 * n_records = number of t in (t0, t0+steps] with t % collect_interval == 0; steps_out
 * receives those t.  A metrics row is JXB_MAX_METRICS doubles wide.  Either may be NULL.  device_seconds_out: CUDA-event time of the
 * step loop on the engine's stream.                                                 */
int jxb_model_run(jxb_model*, int steps, int collect_interval, double* metrics_out,
                  int32_t* steps_out, int* n_records_out, double* device_seconds_out);
int jxb_model_time_step(jxb_model*, int64_t* out);
/* Seconds of the dominant kernel of the last run (CUDA events around every launch of
 * it when profiling is enabled with jxb_model_set_profile(m, 1)).                   */
int jxb_model_set_profile(jxb_model*, int enable);
int jxb_model_profile(jxb_model*, double* dominant_kernel_seconds, int64_t* launches,
                      const char** kernel_name);

/* Page-locked host blocks for results (AgentCollection.states reads): a download into such a block
 * runs at PCIe speed without a staging copy.  Freed blocks are cached by size inside the library
 * (cudaMallocHost of a 100 MB block costs tens of milliseconds).                                 */
int jxb_host_alloc(size_t bytes, void** out);
int jxb_host_free(void* p);

/* ---- record / select ------------------------------------------------ */
/* Per-agent time series: the facade's Results is meant to carry 'agents.<name>.<var>' entries
 * (jaxabm/agentpy.py:1103-1106; the reference never fills them).  Opt-in: name up to 8 (collection, field)
 * columns; every later jxb_model_run snapshots them on the device whenever it records a history row
 * (t % collect_interval == 0) and keeps the snapshots in HBM until the next run.  n = 0 switches it off.   */
int jxb_model_record_fields(jxb_model*, int n, const int32_t* types, const int32_t* fields);
/* series k of the last run: n_records snapshots of bytes_per_record bytes each (the column, in agent order) */
int jxb_model_series_info(jxb_model*, int k, int* n_records, size_t* bytes_per_record);
int jxb_model_series_download(jxb_model*, int k, void* host, size_t bytes);

/* AgentCollection.filter(condition) (jaxabm/agent.py:213-243) as a device stream compaction.
 * select: flag every agent of collection `type` either by a postfix predicate program over its own state
 * columns (the host shim traces `condition` into it; opcodes JP_* in csrc/record.cuh) or, when `prog` is NULL,
 * by a host-evaluated mask of n_agents bytes; returns the number of selected agents.
 * gather: scatter the selected agents' rows of every state column, in agent order, into collection
 * `dst_type` of `dst` -- a model whose collection has the same rule and exactly `count` agents.          */
typedef struct { int32_t op; int32_t a; int32_t b; float f; } jxb_pred_ins;
int jxb_collection_filter_select(jxb_model*, int type, const jxb_pred_ins* prog, int n_ins,
                                 const uint8_t* host_mask, size_t mask_bytes, int64_t* count_out);
int jxb_collection_filter_gather(jxb_model* src, int type, jxb_model* dst, int dst_type);

/* ---- ensembles (jaxabm/analysis.py:113-157 and :434-476) --------------------------- */
/* R independent replicas of `desc`; replica r overrides params by
 * param_slots/params[r][n_swept] (slot < 100: model param index; slot >= 100:
 * 100 + 16*collection + rule-param index) and is seeded PRNGKey(seeds[r]).  env_init
 * ([n_env] doubles, or NULL for the program defaults) is the add_env_state() layout every
 * replica starts from.  The whole time loop of a replica runs on-chip.
 * last_metrics_out: [R][n_metrics] doubles = results[m][-1] of every run.               */
int jxb_ensemble_run(jxb_engine*, const jxb_model_desc* desc, int n_replicas,
                     int n_swept, const int32_t* param_slots, const double* params,
                     const uint32_t* seeds, const double* env_init, int steps,
                     double* last_metrics_out, double* device_seconds_out);

/* ---- population sharding across processes (one process per GPU) -------------------- */
/* Attach an NCCL communicator: id_bytes = ncclUniqueId broadcast by the host shim
 * over torch.distributed.  Used for the per-step env partial-sum all-reduce.         */
int jxb_nccl_unique_id(void* id_bytes_out, size_t bytes);
int jxb_engine_attach_nccl(jxb_engine*, const void* id_bytes, size_t bytes, int rank,
                           int world_size);

/* Peer-memory exchange (the product path): every rank exports the CUDA IPC handle of its
 * exchange buffer, the host shim all-gathers the JXB_IPC_HANDLE_BYTES-byte handles and every
 * rank maps its peers' buffers over NVLink.  The step kernel's last CTA then stores its
 * partial-sum row into every peer, waits for the peers' rows and folds them in rank order:
 * update + reduction + exchange in ONE kernel, no NCCL launch on the step path.
 * JXB_EXCHANGE=nccl selects the NCCL all-reduce path instead (kept as the comparison arm). */
#define JXB_IPC_HANDLE_BYTES 64
int jxb_engine_p2p_export(jxb_engine*, void* handle_out, size_t bytes);
int jxb_engine_p2p_attach(jxb_engine*, const void* handles, size_t bytes_each, int rank,
                          int world_size);

/* ---- ONE grid split into row bands over the GPUs of a box (Grid) ------ */
/* jaxabm/agentpy.py:480-527 keeps one env['grid'] for all agents; here rank r of desc.world_size owns
 * the rows [row_begin, row_end) (consecutive bands in rank order that tile the grid, at most
 * ceil(W / world) rows each) and every rank is handed the whole per-agent columns.  export allocates
 * this rank's receive area (its range of the empty-cell slots + the record segments) and returns
 * JXB_GRID_HANDLE_BYTES: the area's CUDA IPC handle followed by the band and a launch shape as int32[4]; the host shim
 * all-gathers the entries and attach maps the peers' areas (one process per rank; the ranks may share
 * a device).  After that jxb_model_run steps the band with no host round trip: a rank walks only the
 * movers of its own rows and stores each as a request into the area of the rank that holds the mover's
 * slot; that rank rewrites the slot and forwards the mover to the owner of the target row -- posted
 * stores over NVLink and step flags only (no NCCL call, no remote load on the step path).
 * Downloads of a shard return ITS view; the host combines the ranks: 'position' max (-1 = agent is
 * not in my band), 'satisfied' min, 'moves' sum (0 = not in my band), 'type' any; env grid rows
 * [row_begin, row_end); empty_cells is gathered from the ranks' areas (barrier first).              */
#define JXB_GRID_HANDLE_BYTES 80
int jxb_model_grid_shard_export(jxb_model*, int row_begin, int row_end, void* handle_out, size_t bytes);
int jxb_model_grid_shard_attach(jxb_model*, const void* handles, size_t bytes_each, int n_ranks);

/* ---- ONE network split by node ranges over the GPUs of a box (Network) - */
/* jaxabm/agentpy.py:545-582 keeps one env['network_edges'] for all agents; here rank r owns the agents
 * [global_offset, global_offset + n_agents) of types[0] (boundaries at multiples of 32) and passes
 * jxb_model_set_network only the edges whose SOURCE it owns, sources re-based to local rows, targets
 * left as global ids.  Every rank holds the two global "is infected" bitmaps inside an IPC-shared
 * receive area; export / attach map the peers' areas (as for the grid), sync hands this rank's slice
 * of the freshly packed bitmap to every peer (call it after init / an upload of 'state', then barrier).
 * jxb_model_run then steps the rank's rows in the pull direction: the aggregation kernel stores the
 * new bitmap word of each 32-row group straight into every rank's copy over NVLink and releases a
 * step flag; a one-warp wait kernel folds the ranks' exact S/I/R counts.  'state' reads return this
 * rank's slice.                                                                                          */
int jxb_model_net_shard_export(jxb_model*, void* handle_out, size_t bytes);
int jxb_model_net_shard_attach(jxb_model*, const void* handles, size_t bytes_each, int n_ranks);
int jxb_model_net_shard_sync(jxb_model*);

/* ---- host-only scalar key algebra (jax.random on the reference's host path) -------- */
/* jax.random.split(key, n) -> out[n][2] (jaxabm/model.py:129,156; analysis.py:438).  */
int jxb_prng_split(int rng_mode, const uint32_t key[2], int n, uint32_t* out);
/* jax.random.bits(key, (n,)) 32-bit.                                                  */
int jxb_prng_bits(int rng_mode, const uint32_t key[2], int64_t n, uint32_t* out);
/* jax.random.uniform(key, (n,), minval, maxval) float32 (analysis.py:81).            */
int jxb_prng_uniform(int rng_mode, const uint32_t key[2], int64_t n, float lo, float hi,
                     float* out);
/* jax.random.randint(key, (n,), lo, hi) int32 (agentpy.py:510-512; analysis.py:441). */
int jxb_prng_randint(int rng_mode, const uint32_t key[2], int64_t n, int32_t lo,
                     int32_t hi, int32_t* out);
/* the keyed bijection of [0, n) that matches Schelling movers to empty-cell slots (DESIGN.md "Schelling
 * rule": 4-round balanced Feistel with cycle walking, round keys = random_bits(coll_key, (8,))[0:4] / [4:8]);
 * inverse != 0 evaluates the inverse permutation (the kernels walk the unsatisfied agents in cell order and
 * ask which mover index each one is).                                                  */
int jxb_prng_feistel(uint32_t n, const uint32_t round_keys[4], uint32_t idx, int inverse, uint32_t* out);
/* raw block function, for known-answer tests.                                         */
int jxb_prng_threefry2x32(const uint32_t key[2], const uint32_t ctr[2], uint32_t out[2]);

#ifdef __cplusplus
}
#endif
#endif /* JXB_H_ */
