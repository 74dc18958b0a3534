// record.cuh -- the record / select side of the path:
//   * per-agent time series: the reference's facade intends Results keys 'agents.<name>.<var>'
//     (jaxabm/agentpy.py:1103-1106) but never fills them; here selected state columns are snapshotted into a
//     device ring every collect_interval steps (a copy kernel at the end of the step's launches, gated on the
//     device-resident step counter, so it replays inside the captured step graph) and read back once per run();
//   * AgentCollection.filter (jaxabm/agent.py:213-243) as a stream compaction: a flag per agent (from a small
//     postfix predicate program evaluated per agent, or from a mask the host evaluated), an exclusive scan, and
//     one ordered scatter per state column straight into the new collection's columns -- the selected agents keep
//     their order (v[mask] semantics) and no column travels through the host.
#pragma once
#include "common.cuh"

namespace jxb {

// ---------------------------------------------------------------------------------------------------------
// series snapshots
// ---------------------------------------------------------------------------------------------------------
// Runs after the step's tail advanced ctrl->time_step / n_recorded: when this step recorded a history row, copy
// the column into ring slot n_recorded - 1.  16-byte chunks; columns are (N + 8)-padded so the tail chunk is safe.
__global__ void __launch_bounds__(kThreads) series_snapshot_kernel(const Ctrl* ctrl, int collect_interval, const uint4* src,
                                                                   unsigned char* ring, unsigned long long bytes_per_snap,
                                                                   unsigned long long slot_stride) {
  const long long t = ctrl->time_step;
  if ((t % collect_interval) != 0 || ctrl->n_recorded <= 0) return;
  uint4* dst = (uint4*)(ring + (unsigned long long)(ctrl->n_recorded - 1) * slot_stride);
  const unsigned long long chunks = (bytes_per_snap + 15) / 16;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks;
       i += (unsigned long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------------
// filter: predicate program
// ---------------------------------------------------------------------------------------------------------
// Postfix program over the agent's own state (jxb_pred_ins of include/jxb.h).  Values on the stack are 32-bit
// patterns whose type (float32 / int32 / bool) the host-side tracer tracked with JAX's promotion rules and baked
// into the opcodes.
enum {
  JP_LOAD_F32 = 1, JP_LOAD_I32, JP_LOAD_U8,          // a = field, b = component            -> push
  JP_CONST_F32, JP_CONST_I32,                        // f / a                                -> push
  JP_ADD_F, JP_SUB_F, JP_MUL_F, JP_DIV_F, JP_MIN_F, JP_MAX_F,
  JP_ADD_I, JP_SUB_I, JP_MUL_I, JP_MIN_I, JP_MAX_I,
  JP_NEG_F, JP_NEG_I, JP_ABS_F, JP_ABS_I,
  JP_I2F, JP_F2I, JP_B2I, JP_B2F, JP_I2B, JP_F2B,
  JP_LT_F, JP_LE_F, JP_GT_F, JP_GE_F, JP_EQ_F, JP_NE_F,
  JP_LT_I, JP_LE_I, JP_GT_I, JP_GE_I, JP_EQ_I, JP_NE_I,
  JP_AND, JP_OR, JP_XOR, JP_NOT,
  JP_SELECT,                                         // cond, a, b -> cond ? a : b
  JP_SQRT_F, JP_EXP_F, JP_LOG_F,
  JP_OP_COUNT
};

constexpr int kPredMaxIns = 96;
constexpr int kPredStack = 16;

struct PredProgram {
  int n;
  jxb_pred_ins ins[kPredMaxIns];
};

struct FilterCols {
  const void* f[kMaxFields];
  int width[kMaxFields];
};

__device__ __forceinline__ float jp_nanmin(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float jp_nanmax(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__device__ inline bool pred_eval(const PredProgram& pg, const FilterCols& fc, long long i) {
  unsigned int st[kPredStack];
  int sp = 0;
#define F(x) __uint_as_float(x)
#define U(x) __float_as_uint(x)
  for (int k = 0; k < pg.n; ++k) {
    const jxb_pred_ins in = pg.ins[k];
    switch (in.op) {
      case JP_LOAD_F32: st[sp++] = U(((const float*)fc.f[in.a])[i * fc.width[in.a] + in.b]); break;
      case JP_LOAD_I32: st[sp++] = (unsigned int)((const int*)fc.f[in.a])[i * fc.width[in.a] + in.b]; break;
      case JP_LOAD_U8: st[sp++] = ((const unsigned char*)fc.f[in.a])[i * fc.width[in.a] + in.b] != 0; break;
      case JP_CONST_F32: st[sp++] = U(in.f); break;
      case JP_CONST_I32: st[sp++] = (unsigned int)in.a; break;
#define BIN_F(OP, EXPR) case OP: { const float b = F(st[--sp]), a = F(st[sp - 1]); st[sp - 1] = EXPR; break; }
#define BIN_I(OP, EXPR) case OP: { const int b = (int)st[--sp], a = (int)st[sp - 1]; st[sp - 1] = (unsigned int)(EXPR); break; }
      BIN_F(JP_ADD_F, U(a + b)) BIN_F(JP_SUB_F, U(a - b)) BIN_F(JP_MUL_F, U(a * b)) BIN_F(JP_DIV_F, U(a / b))
      BIN_F(JP_MIN_F, U(jp_nanmin(a, b))) BIN_F(JP_MAX_F, U(jp_nanmax(a, b)))
      BIN_I(JP_ADD_I, a + b) BIN_I(JP_SUB_I, a - b) BIN_I(JP_MUL_I, a * b) BIN_I(JP_MIN_I, min(a, b)) BIN_I(JP_MAX_I, max(a, b))
      BIN_F(JP_LT_F, a < b) BIN_F(JP_LE_F, a <= b) BIN_F(JP_GT_F, a > b) BIN_F(JP_GE_F, a >= b) BIN_F(JP_EQ_F, a == b) BIN_F(JP_NE_F, a != b)
      BIN_I(JP_LT_I, a < b) BIN_I(JP_LE_I, a <= b) BIN_I(JP_GT_I, a > b) BIN_I(JP_GE_I, a >= b) BIN_I(JP_EQ_I, a == b) BIN_I(JP_NE_I, a != b)
      BIN_I(JP_AND, (a != 0) && (b != 0)) BIN_I(JP_OR, (a != 0) || (b != 0)) BIN_I(JP_XOR, (a != 0) != (b != 0))
#undef BIN_F
#undef BIN_I
      case JP_NOT: st[sp - 1] = st[sp - 1] == 0; break;
      case JP_NEG_F: st[sp - 1] = U(-F(st[sp - 1])); break;
      case JP_NEG_I: st[sp - 1] = (unsigned int)(-(int)st[sp - 1]); break;
      case JP_ABS_F: st[sp - 1] = U(fabsf(F(st[sp - 1]))); break;
      case JP_ABS_I: st[sp - 1] = (unsigned int)abs((int)st[sp - 1]); break;
      case JP_I2F: st[sp - 1] = U((float)(int)st[sp - 1]); break;
      case JP_F2I: st[sp - 1] = (unsigned int)(int)F(st[sp - 1]); break;
      case JP_B2I: break;
      case JP_B2F: st[sp - 1] = U(st[sp - 1] ? 1.0f : 0.0f); break;
      case JP_I2B: st[sp - 1] = st[sp - 1] != 0; break;
      case JP_F2B: st[sp - 1] = F(st[sp - 1]) != 0.0f; break;
      case JP_SELECT: { const unsigned int b = st[--sp], a = st[--sp]; st[sp - 1] = st[sp - 1] ? a : b; break; }
      case JP_SQRT_F: st[sp - 1] = U(sqrtf(F(st[sp - 1]))); break;
      case JP_EXP_F: st[sp - 1] = U(expf(F(st[sp - 1]))); break;
      case JP_LOG_F: st[sp - 1] = U(logf(F(st[sp - 1]))); break;
      default: break;
    }
  }
#undef F
#undef U
  return sp > 0 && st[sp - 1] != 0;
}

// flag[i] = predicate(agent i) (program) or mask[i] != 0 (host-evaluated mask)
__global__ void __launch_bounds__(kThreads) filter_flags_kernel(const PredProgram pg, const FilterCols fc, const unsigned char* mask,
                                                                long long n, unsigned int* flags) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    flags[i] = mask ? (mask[i] != 0) : (pred_eval(pg, fc, i) ? 1u : 0u);
}

// pos = exclusive scan of the flags (pos[n] = count): agent i is selected iff pos[i + 1] != pos[i] and lands at
// pos[i] -- ascending agent order is preserved.  elem = bytes of one agent's entry of this column.
__global__ void __launch_bounds__(kThreads) filter_scatter_kernel(const unsigned int* pos, long long n, const unsigned char* src,
                                                                  unsigned char* dst, int elem) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int p = pos[i];
    if (pos[i + 1] == p) continue;
    if (elem == 4) ((unsigned int*)dst)[p] = ((const unsigned int*)src)[i];
    else if (elem == 8) ((uint2*)dst)[p] = ((const uint2*)src)[i];
    else if (elem == 1) dst[p] = src[i];
    else for (int b = 0; b < elem; ++b) dst[(size_t)p * elem + b] = src[(size_t)i * elem + b];
  }
}

}  // namespace jxb
