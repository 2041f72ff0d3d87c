// C-ABI implementation (include/gww.h): host orchestration of the sm_100a kernels.
// One process per GPU; all work is enqueued on the caller's stream; no allocation on the hot path.
#include "../../include/gww.h"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "attention_tc.cuh"
#include "attention_persist.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"
#include "logmel.cuh"
#include "qfront.cuh"
#include "qadapter_tc.cuh"
#include "whiten.cuh"

using namespace gww;

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(GWW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                  __FILE__, __LINE__);                                                       \
  } while (0)
#define GWW_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != GWW_OK) return _r; \
  } while (0)
#define LAUNCH_CHECK()  \
  do {                  \
    ++g_launches;       \
    CU_TRY(cudaGetLastError()); \
  } while (0)

extern "C" const char* gww_last_error(void) { return g_err.c_str(); }
extern "C" const char* gww_version(void) { return "gw-whisper-b200 0.2 (sm_100a, " GWW_OPERAND_NAME " operands)"; }
extern "C" const char* gww_operand_dtype(void) { return GWW_OPERAND_NAME; }
extern "C" long gww_launch_count(void) { return g_launches.load(); }

// Process-global state is keyed by the device ordinal (a process may drive more than one GPU) and guarded
// by g_state_mu where it is mutated after start-up.
constexpr int kMaxDevices = 64;
static int g_sms_by_dev[kMaxDevices] = {0};
static std::mutex g_state_mu;
static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
#define g_num_sms (g_sms_by_dev[current_device()])
static std::atomic<int> g_prune_last{1};
extern "C" int gww_set_last_layer_pruning(int enable) { return g_prune_last.exchange(enable ? 1 : 0); }
extern "C" int gww_device_ok(void) {
  if (g_num_sms > 0) return GWW_OK;   // checked once per device
  int dev = 0, count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(GWW_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
  }
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(GWW_ERR_NO_DEVICE, "device %s is sm_%d%d; kernels are built for sm_100a only",
                prop.name, prop.major, prop.minor);
  if (dev < 0 || dev >= kMaxDevices) return fail(GWW_ERR_NO_DEVICE, "device ordinal %d out of range", dev);
  g_sms_by_dev[dev] = prop.multiProcessorCount;
  return GWW_OK;
}

// cudaFuncSetAttribute is per (function, device): remember which pairs were already opted in.
static std::unordered_map<uint64_t, int> g_attr_done;
template <typename K>
static int ensure_smem_attr(K kern, int bytes) {
  const uint64_t key = (uint64_t)(uintptr_t)reinterpret_cast<const void*>(kern) * 64ull + (uint64_t)current_device();
  std::lock_guard<std::mutex> lk(g_state_mu);
  auto it = g_attr_done.find(key);
  if (it != g_attr_done.end() && it->second >= bytes) return GWW_OK;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  g_attr_done[key] = bytes;
  return GWW_OK;
}


// ------------------------------------------------------------------------------------------------
// optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline)
// ------------------------------------------------------------------------------------------------
enum ProfKind : int {
  PK_LOGMEL = 0, PK_FEATS_TM, PK_GEMM_CONV1, PK_GEMM_CONV2, PK_LN, PK_GEMM_QKV, PK_ATTN, PK_GEMM_O,
  PK_GEMM_FC1, PK_GEMM_FC2, PK_HEAD, PK_OTHER, PK_QSCAN, PK_QADAPTER, PK_ATTN_LAST, PK_WHITEN, PK_QA_CONV1, PK_QA_CONV2,
  PK_QA_CONV3, PK_QA_POOL, PK_COUNT
};
static const char* kProfNames[PK_COUNT] = {"logmel", "feats_to_timemajor", "gemm_conv1", "gemm_conv2",
                                           "layernorm", "gemm_qkv", "attention", "gemm_out_proj",
                                           "gemm_fc1", "gemm_fc2", "head", "other", "qscan", "qadapter",
                                           "attention_last_row", "whiten", "qadapter_conv1", "qadapter_conv2",
                                           "qadapter_conv3", "qadapter_pool"};
struct ProfRec { cudaEvent_t a, b; int kind; };
// The profiler is a single-device, single-thread diagnostic (bench.py): records are kept in a deque so that
// pointers stay valid while it grows, the bookkeeping is under g_prof_mu, and records made on another device
// than the one gww_profile_begin() ran on are skipped.
static std::atomic<bool> g_prof_on{false};
static std::deque<ProfRec> g_prof_recs;
static size_t g_prof_used = 0;
static int g_prof_dev = 0;
static std::mutex g_prof_mu;
struct ProfScope {
  ProfRec* r = nullptr;
  cudaStream_t s;
  ProfScope(int kind, cudaStream_t stream) : s(stream) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    {
      std::lock_guard<std::mutex> lk(g_prof_mu);
      if (current_device() != g_prof_dev) return;
      if (g_prof_used == g_prof_recs.size()) {
        ProfRec n{};
        if (cudaEventCreate(&n.a) != cudaSuccess || cudaEventCreate(&n.b) != cudaSuccess) return;
        g_prof_recs.push_back(n);
      }
      r = &g_prof_recs[g_prof_used++];
    }
    r->kind = kind;
    cudaEventRecord(r->a, s);
  }
  ~ProfScope() { if (r) cudaEventRecord(r->b, s); }
};
extern "C" int gww_profile_num_kinds(void) { return PK_COUNT; }
extern "C" const char* gww_profile_kind_name(int k) { return (k >= 0 && k < PK_COUNT) ? kProfNames[k] : ""; }
extern "C" int gww_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_used = 0;
  g_prof_dev = current_device();
  g_prof_on = true;
  return GWW_OK;
}
// Stops profiling, waits for the recorded events and accumulates milliseconds / launch counts.
extern "C" int gww_profile_end(double* ms_by_kind, long* count_by_kind) {
  g_prof_on = false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < PK_COUNT; ++k) { ms_by_kind[k] = 0.0; count_by_kind[k] = 0; }
  for (size_t i = 0; i < g_prof_used; ++i) {
    ProfRec& r = g_prof_recs[i];
    CU_TRY(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_kind[r.kind] += ms;
    count_by_kind[r.kind] += 1;
  }
  g_prof_used = 0;
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int get_encode() {
  if (g_encode) return GWW_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(GWW_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return GWW_OK;
}

// Encoded tensor maps are cached: the same (buffer, shape) pairs recur every chunk and
// cuTensorMapEncodeTiled costs tens of microseconds of host time per call.
struct MapKey {
  uint64_t v[16];
  bool operator==(const MapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static std::mutex g_map_mu;

// dims/box are innermost-first; strides_bytes has rank-1 entries (dim 1..rank-1).
static int make_map_uncached(CUtensorMap* m, bool f32, int rank, const void* base, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, bool swizzle);
static int make_map(CUtensorMap* m, bool f32, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle = true) {
  MapKey k{};
  k.v[0] = (uint64_t)f32 | ((uint64_t)rank << 8) | ((uint64_t)swizzle << 16);
  k.v[1] = (uint64_t)(uintptr_t)base;
  for (int i = 0; i < rank; ++i) { k.v[2 + i] = dims[i]; k.v[11 + i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) k.v[7 + i] = strides_bytes[i];
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_map_cache.find(k);
  if (it != g_map_cache.end()) { *m = it->second; return GWW_OK; }
  GWW_TRY(make_map_uncached(m, f32, rank, base, dims, strides_bytes, box, swizzle));
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache.emplace(k, *m);
  return GWW_OK;
}
static int make_map_uncached(CUtensorMap* m, bool f32, int rank, const void* base, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, bool swizzle) {
  GWW_TRY(get_encode());
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                            : (GWW_OPERAND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16),
                        rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(GWW_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank=%d base=%p dims=[%llu,%llu,%llu,%llu] "
                "strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                (int)r, rank, base, (unsigned long long)dims[0],
                (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0),
                (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0],
                rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// GEMM dispatch
// ------------------------------------------------------------------------------------------------
constexpr int kStatSlotsMax = 16;
struct GemmCall {
  // A operand: 4-D view {C, P, R, Bt} of a bf16 activation
  const void* a_base;
  uint64_t a_dims[4];
  uint64_t a_strides[3];
  // W: [N, Ktot] bf16
  const void* w_base;
  int ktot;
  // C: 3-D view {N, rows, batch}
  void* c_base;
  uint64_t c_strides[2];
  GemmParams p;
  int epi;
  int block_n;
  int kind = PK_OTHER;
};

template <int BN, int EPI, int MC>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR,
                         const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, EPI, MC>;
  GWW_TRY(ensure_smem_attr(kern, GemmSmem<BN, MC>::kTotal));
  const int tiles_m = ((p.rows + 127) / 128) * p.batch;
  const int items = ((tiles_m + MC - 1) / MC) * ((p.n + BN - 1) / BN);
  const int max_groups = g_num_sms / MC;
  const int grid = (items < max_groups ? items : max_groups) * MC;
  if constexpr (MC == 1) {
    kern<<<grid, gemm_threads<EPI, BN>(), GemmSmem<BN, MC>::kTotal, stream>>>(tmA, tmB, tmR, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(gemm_threads<EPI, BN>());
    cfg.dynamicSmemBytes = GemmSmem<BN, MC>::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MC;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CU_TRY(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmR, p));
  }
  LAUNCH_CHECK();
  return GWW_OK;
}

template <int BN, int MC>
static int launch_gemm_bn(int epi, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& r,
                          const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_BIAS_BF16: return launch_gemm_t<BN, EPI_BIAS_BF16, MC>(a, b, r, p, s);
    case EPI_BIAS_GELU_BF16: return launch_gemm_t<BN, EPI_BIAS_GELU_BF16, MC>(a, b, r, p, s);
    case EPI_BIAS_RESID_F32: return launch_gemm_t<BN, EPI_BIAS_RESID_F32, MC>(a, b, r, p, s);
    case EPI_BIAS_GELU_POS_F32: return launch_gemm_t<BN, EPI_BIAS_GELU_POS_F32, MC>(a, b, r, p, s);
  }
  return fail(GWW_ERR_INVALID, "unknown epilogue %d", epi);
}

// CTA-pair mode (tcgen05.mma.cta_group::2) needs enough tile pairs to keep every SM pair busy.  Measured
// inside the whisper-base step (r1, ms per step, pair vs single): qkv 24.3 / 28.3, fc2 29.0 / 33.0, but
// out_proj 17.7 / 16.3 and fc1 (GELU epilogue) 32.4 / 30.9 -- the small-N K=512 residual GEMM is HBM-bound
// and the GELU GEMM is bound by its epilogue's issue slots, where the pair's lock-step costs more than
// its halved operand traffic saves.
static int gemm_pair_mode(const GemmParams& p, int epi, int ktot, int kind) {
  // tuning override per GEMM kind: GWW_GEMM_MC_QKV / _O / _FC1 / _FC2 / _CONV1 / _CONV2 = 1 | 2 (read once)
  static int per_kind[PK_COUNT];
  static bool per_kind_init = false;
  if (!per_kind_init) {
    const struct { int k; const char* name; } tbl[] = {{PK_GEMM_QKV, "GWW_GEMM_MC_QKV"}, {PK_GEMM_O, "GWW_GEMM_MC_O"},
                                                       {PK_GEMM_FC1, "GWW_GEMM_MC_FC1"}, {PK_GEMM_FC2, "GWW_GEMM_MC_FC2"},
                                                       {PK_GEMM_CONV1, "GWW_GEMM_MC_CONV1"}, {PK_GEMM_CONV2, "GWW_GEMM_MC_CONV2"}};
    for (int i = 0; i < PK_COUNT; ++i) per_kind[i] = 0;
    for (const auto& e : tbl) { const char* v = getenv(e.name); per_kind[e.k] = v ? atoi(v) : 0; }
    per_kind_init = true;
  }
  if (kind >= 0 && kind < PK_COUNT && (per_kind[kind] == 1 || per_kind[kind] == 2)) {
    const int tiles_m_k = ((p.rows + 127) / 128) * p.batch;
    return (per_kind[kind] == 2 && tiles_m_k < 2) ? 1 : per_kind[kind];
  }
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GWW_GEMM_MC");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1 || forced == 2) return forced;
  const int tiles_m = ((p.rows + 127) / 128) * p.batch;
  if (tiles_m < g_num_sms) return 1;
  if (epi == EPI_BIAS_GELU_BF16) return 1;
  // whisper-tiny's qkv projection (K = 384: six k-blocks per tile): measured inside the MLGWSC-1 search, single CTA
  // 400 ms per hour of strain against 492 for the pair (r2); fc2 / conv2 / fc1 of the same model are best as chosen below
  if (epi == EPI_BIAS_BF16 && ktot <= 384) return 1;
  if (epi == EPI_BIAS_RESID_F32 && ktot <= 768 && p.n <= 768) return 1;
  return 2;
}

static int run_gemm(const GemmCall& g, cudaStream_t stream) {
  if (g.p.n % 64 != 0) return fail(GWW_ERR_INVALID, "gemm: N=%d must be a multiple of 64", g.p.n);
  if (g.ktot % 64 != 0) return fail(GWW_ERR_INVALID, "gemm: K=%d must be a multiple of 64", g.ktot);
  const bool out_f32 = (g.epi == EPI_BIAS_RESID_F32 || g.epi == EPI_BIAS_GELU_POS_F32);
  if (out_f32 && g.p.n % g.block_n != 0)
    return fail(GWW_ERR_INVALID, "gemm: f32 epilogues need N %% block_n == 0 (N=%d)", g.p.n);
  CUtensorMap tmA, tmB;
  const uint32_t abox[4] = {64, 1, 128, 1};
  GWW_TRY(make_map(&tmA, false, 4, g.a_base, g.a_dims, g.a_strides, abox));
  const uint64_t wdims[2] = {(uint64_t)g.ktot, (uint64_t)g.p.n};
  const uint64_t wstr[1] = {(uint64_t)g.ktot * 2};
  const int mc = gemm_pair_mode(g.p, g.epi, g.ktot, g.kind);
  const uint32_t wbox[2] = {64, (uint32_t)(g.block_n / mc)};
  GWW_TRY(make_map(&tmB, false, 2, g.w_base, wdims, wstr, wbox));
  GemmParams p = g.p;
  if (p.n > GemmSmem<256, 1>::kBiasFloats)
    return fail(GWW_ERR_INVALID, "gemm: N=%d exceeds the %d-column bias buffer", p.n, GemmSmem<256, 1>::kBiasFloats);
  {
    const int tiles_per_batch = (p.rows + 127) / 128;
    const int tiles_n = (p.n + g.block_n - 1) / g.block_n;
    if ((long)tiles_per_batch * p.batch * tiles_n >= (1l << 24))
      return fail(GWW_ERR_INVALID, "gemm: too many tiles for the fast tile-index division");
    p.magic_tiles_n = gemm_div_magic(tiles_n);
    p.magic_tiles_per_batch = gemm_div_magic(tiles_per_batch);
  }
  const uint64_t esz = out_f32 ? 4 : 2;
  p.c = g.c_base;
  p.c_row_stride = (long)(g.c_strides[0] / esz);
  p.c_batch_stride = (long)(g.c_strides[1] / esz);
  if ((reinterpret_cast<uintptr_t>(g.c_base) & 31) != 0 || (p.c_row_stride * esz) % 32 != 0 ||
      (p.c_batch_stride * esz) % 32 != 0)
    return fail(GWW_ERR_INVALID, "gemm: output must be 32-byte aligned (base and strides)");
  CUtensorMap tmR = tmB;   // placeholder unless the residual is prefetched
  p.prefetch_resid = 0;
  if (g.epi == EPI_BIAS_RESID_F32 && p.resid != nullptr) {
    const uint64_t rdims[2] = {(uint64_t)p.n, (uint64_t)p.batch * (uint64_t)p.rows};
    const uint64_t rstr[1] = {(uint64_t)p.n * 4};
    const uint32_t rbox[2] = {(uint32_t)g.block_n, 128};
    GWW_TRY(make_map(&tmR, true, 2, p.resid, rdims, rstr, rbox, false));
    p.prefetch_resid = 1;
  }
  ProfScope ps(g.kind, stream);
  switch (g.block_n * 10 + mc) {
    case 1281: return launch_gemm_bn<128, 1>(g.epi, tmA, tmB, tmR, p, stream);
    case 1921: return launch_gemm_bn<192, 1>(g.epi, tmA, tmB, tmR, p, stream);
    case 2561: return launch_gemm_bn<256, 1>(g.epi, tmA, tmB, tmR, p, stream);
    case 1282: return launch_gemm_bn<128, 2>(g.epi, tmA, tmB, tmR, p, stream);
    case 1922: return launch_gemm_bn<192, 2>(g.epi, tmA, tmB, tmR, p, stream);
    case 2562: return launch_gemm_bn<256, 2>(g.epi, tmA, tmB, tmR, p, stream);
  }
  return fail(GWW_ERR_INVALID, "gemm: block_n must be 128, 192 or 256 (got %d)", g.block_n);
}

// plain Linear: C[M,N] = epi(A[M,K] W[N,K]^T)
// LayerNorm-fold plumbing of one Linear: producer side (writes xb + stats) or consumer side (reads them)
struct LnFold {
  op16_t* xb = nullptr;
  float* stats_out = nullptr;
  const float* stats_in = nullptr;
  const float* c1 = nullptr;
  int in_slots = 0;     // slots the producer of stats_in wrote
  int d_model = 0;
  float* guard = nullptr;
};
// |mean| / std of a residual row above which the folded LayerNorm is considered unsafe (validated up to 1.2 on
// random-init weights in round 1; the operand rounding of x, 2^-12 for fp16 / 2^-9 for bf16 relative to |mean|, is
// amplified by this ratio in the normalised value)
constexpr float kLnGuardLimit = GWW_OPERAND_BF16 ? 1.5f : 4.0f;
static int stat_slots_of(int n, int block_n) { return 2 * ((n + block_n - 1) / block_n); }
static void apply_fold(GemmParams& p, const LnFold& lf, int block_n) {
  p.xb = lf.xb;
  p.stats_out = lf.stats_out;
  p.stats_in = lf.stats_in;
  p.ln_c1 = lf.c1;
  p.stat_slots = lf.stats_in ? lf.in_slots : stat_slots_of(p.n, block_n);
  p.ln_inv_k = lf.d_model > 0 ? 1.0f / (float)lf.d_model : 0.f;
  p.ln_eps = 1e-5f;
  p.ln_guard = lf.stats_in ? lf.guard : nullptr;
  p.ln_guard_min = 0.5f * kLnGuardLimit;
}

static int run_linear(const void* A, const void* W, void* C, const float* bias, const float* resid,
                      long M, int N, int K, int epi, int block_n, cudaStream_t stream,
                      int kind = PK_OTHER, const LnFold& lf = LnFold(), long ldc = 0) {
  // ldc: elements between consecutive output rows (0 = N: a dense [M, N] output)
  GemmCall g{};
  g.kind = kind;
  g.a_base = A;
  g.a_dims[0] = K; g.a_dims[1] = 1; g.a_dims[2] = M; g.a_dims[3] = 1;
  g.a_strides[0] = (uint64_t)K * 2; g.a_strides[1] = (uint64_t)K * 2; g.a_strides[2] = (uint64_t)M * K * 2;
  g.w_base = W;
  g.ktot = K;
  g.c_base = C;
  const bool out_f32 = (epi == EPI_BIAS_RESID_F32 || epi == EPI_BIAS_GELU_POS_F32);
  const uint64_t esz = out_f32 ? 4 : 2;
  g.c_strides[0] = (uint64_t)(ldc > 0 ? ldc : N) * esz;
  g.c_strides[1] = (uint64_t)M * g.c_strides[0];
  g.p.rows = (int)M; g.p.batch = 1; g.p.n = N; g.p.kb_per_tap = K / 64; g.p.taps = 1; g.p.p_mod = 1;
  g.p.bias = bias; g.p.resid = resid; g.p.pos = nullptr;
  g.epi = epi;
  g.block_n = block_n;
  apply_fold(g.p, lf, block_n);
  if (g.p.stats_out && g.p.stat_slots > kStatSlotsMax)
    return fail(GWW_ERR_INVALID, "gemm: %d statistic slots exceed the workspace layout", g.p.stat_slots);
  return run_gemm(g, stream);
}

static int pick_block_n(int n) {
  if (n % 256 == 0) return 256;
  if (n % 192 == 0) return 192;
  return 128;
}

// ------------------------------------------------------------------------------------------------
// attention / layernorm launchers
// ------------------------------------------------------------------------------------------------
template <int NT>
static int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmO, const AttnParams& ap, long n, int T, int d,
                            cudaStream_t stream) {
  auto kern = attention_tc_kernel<NT>;
  GWW_TRY(ensure_smem_attr(kern, AttnCfg<NT>::kSmemBytes));
  if (n > 32768) return fail(GWW_ERR_INVALID, "attention: more than 32768 det-windows per call");   // gridDim.z limit
  dim3 grid((T + 128 * NT - 1) / (128 * NT), d / 64, (unsigned)n);
  ProfScope ps(PK_ATTN, stream);
  kern<<<grid, AttnCfg<NT>::kThreads, AttnCfg<NT>::kSmemBytes, stream>>>(tmQ, tmO, ap);
  LAUNCH_CHECK();
  return GWW_OK;
}

// query tiles per CTA: 1 = two independent CTAs per SM (default), 2 = one CTA with both tiles
static int attention_tiles_per_cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GWW_ATTN_NT");
    v = (e && atoi(e) == 2) ? 2 : 1;
  }
  return v;
}

// persistent one-CTA-per-SM kernel (default) vs one CTA per work item (GWW_ATTN_PERSIST=0)
static bool attention_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GWW_ATTN_PERSIST");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// persistent TMA-fed conv1 of the Q-Adapter (default) vs one CTA per tile with plain loads (GWW_QA_CONV1_TMA=0)
static bool qadapter_conv1_tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GWW_QA_CONV1_TMA");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

static int run_attention(const void* qkv, void* out, long n, int T, int d, cudaStream_t stream) {
  if (d % 64 != 0) return fail(GWW_ERR_INVALID, "attention: d_model %% 64 != 0");
  CUtensorMap tmQ, tmO;
  const uint64_t qd[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)n};
  const uint64_t qs[2] = {(uint64_t)6 * d, (uint64_t)T * 6 * d};
  const uint32_t qb[3] = {64, 128, 1};
  GWW_TRY(make_map(&tmQ, false, 3, qkv, qd, qs, qb));
  const uint64_t od[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)n};
  const uint64_t os[2] = {(uint64_t)2 * d, (uint64_t)T * 2 * d};
  const uint32_t ob[3] = {64, 32, 1};
  GWW_TRY(make_map(&tmO, false, 3, out, od, os, ob));
  if (attention_persistent()) {
    GWW_TRY(ensure_smem_attr(attention_persist_kernel, kApSmemBytes));
    AttnPersistParams pp;
    pp.T = T; pp.d_model = d; pp.nkv = (T + 127) / 128; pp.n_heads = d / 64; pp.n_qpairs = (T + 255) / 256;
    const long items = (long)pp.n_qpairs * pp.n_heads * n;
    if (items > 0x7fffffffL) return fail(GWW_ERR_INVALID, "attention: too many work items");
    pp.n_items = (int)items;
    const int grid = items < g_num_sms ? (int)items : g_num_sms;
    ProfScope ps(PK_ATTN, stream);
    attention_persist_kernel<<<grid, 384, kApSmemBytes, stream>>>(tmQ, tmO, pp);
    LAUNCH_CHECK();
    return GWW_OK;
  }
  AttnParams ap;
  ap.T = T; ap.d_model = d; ap.nkv = (T + 127) / 128;
  if (attention_tiles_per_cta() == 2) return launch_attention<2>(tmQ, tmO, ap, n, T, d, stream);
  return launch_attention<1>(tmQ, tmO, ap, n, T, d, stream);
}

template <typename OutT>
static int run_ln_t(const float* x, OutT* out, const float* g, const float* b, long rows, int d,
                    long in_off, long in_stride, cudaStream_t stream) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const float eps = 1e-5f;
  ProfScope ps(PK_LN, stream);
  switch (d) {
    case 384: layernorm_kernel<3, OutT><<<grid, 256, 0, stream>>>(x, out, g, b, rows, in_off, in_stride, eps); break;
    case 512: layernorm_kernel<4, OutT><<<grid, 256, 0, stream>>>(x, out, g, b, rows, in_off, in_stride, eps); break;
    case 768: layernorm_kernel<6, OutT><<<grid, 256, 0, stream>>>(x, out, g, b, rows, in_off, in_stride, eps); break;
    case 1024: layernorm_kernel<8, OutT><<<grid, 256, 0, stream>>>(x, out, g, b, rows, in_off, in_stride, eps); break;
    case 1280: layernorm_kernel<10, OutT><<<grid, 256, 0, stream>>>(x, out, g, b, rows, in_off, in_stride, eps); break;
    default: return fail(GWW_ERR_INVALID, "layernorm: unsupported d=%d", d);
  }
  LAUNCH_CHECK();
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// log-mel tables
// ------------------------------------------------------------------------------------------------
static LogmelTables g_lm_by_dev[kMaxDevices]{};
static bool g_lm_ready_by_dev[kMaxDevices] = {false};
static std::mutex g_lm_mu;
#define g_lm (g_lm_by_dev[current_device()])
#define g_lm_ready (g_lm_ready_by_dev[current_device()])

static double hz_to_mel(double f) {
  return (f >= 1000.0) ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) {
  return (m >= 15.0) ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0;
}

template <typename T>
static int upload(const std::vector<T>& h, const T** dptr) {
  T* d = nullptr;
  CU_TRY(cudaMalloc(&d, h.size() * sizeof(T)));
  CU_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = d;
  return GWW_OK;
}

static int logmel_tables_init() {
  std::lock_guard<std::mutex> lk(g_lm_mu);
  if (g_lm_ready) return GWW_OK;
  const double PI = 3.14159265358979323846;
  std::vector<double2> t2048(1024), t16000(16000), t125(125), t400(400);
  for (int j = 0; j < 1024; ++j) t2048[j] = make_double2(std::cos(2 * PI * j / 2048), -std::sin(2 * PI * j / 2048));
  for (int j = 0; j < 16000; ++j) t16000[j] = make_double2(std::cos(2 * PI * j / 16000), std::sin(2 * PI * j / 16000));
  for (int j = 0; j < 125; ++j) t125[j] = make_double2(std::cos(2 * PI * j / 125), std::sin(2 * PI * j / 125));
  for (int j = 0; j < 400; ++j) t400[j] = make_double2(std::cos(2 * PI * j / 400), std::sin(2 * PI * j / 400));
  // slaney mel filter bank, 201 bins x 80 filters (HF audio_utils.mel_filter_bank,
  // feature_extraction_whisper.py:94-102): triangular, slaney area normalisation
  std::vector<double> hz(82);
  const double mlo = hz_to_mel(0.0), mhi = hz_to_mel(8000.0);
  for (int i = 0; i < 82; ++i) hz[i] = mel_to_hz(mlo + (mhi - mlo) * i / 81.0);
  std::vector<int> lo(80), cnt(80), off(80);
  std::vector<double> wts;
  for (int m = 0; m < 80; ++m) {
    const double enorm = 2.0 / (hz[m + 2] - hz[m]);
    int first = -1, last = -1;
    std::vector<double> row(201);
    for (int k = 0; k < 201; ++k) {
      const double f = 8000.0 * k / 200.0;
      const double down = (f - hz[m]) / (hz[m + 1] - hz[m]);
      const double up = (hz[m + 2] - f) / (hz[m + 2] - hz[m + 1]);
      const double v = std::fmax(0.0, std::fmin(down, up)) * enorm;
      row[k] = v;
      if (v > 0.0) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    lo[m] = first; cnt[m] = last - first + 1; off[m] = (int)wts.size();
    for (int k = first; k <= last; ++k) wts.push_back(row[k]);
  }
  if (wts.empty()) wts.push_back(0.0);
  GWW_TRY(upload(t2048, &g_lm.tw2048));
  GWW_TRY(upload(t16000, &g_lm.tw16000));
  GWW_TRY(upload(t125, &g_lm.tw125));
  GWW_TRY(upload(t400, &g_lm.tw400));
  GWW_TRY(upload(lo, &g_lm.mel_lo));
  GWW_TRY(upload(cnt, &g_lm.mel_cnt));
  GWW_TRY(upload(off, &g_lm.mel_off));
  GWW_TRY(upload(wts, &g_lm.mel_w));
  GWW_TRY(ensure_smem_attr(logmel_kernel, kLmSmemBytes));
  g_lm_ready = true;
  return GWW_OK;
}

// zero the first `n16` 16-byte words of every row (row pitch `pitch16` words)
__global__ void zero_rows_kernel(uint4* base, size_t pitch16, int n16) {
  uint4* row = base + blockIdx.x * pitch16;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) row[i] = make_uint4(0u, 0u, 0u, 0u);
}

// strain addressing: det-window w = b*D + i reads strain + b*win_stride + i*det_stride
__global__ void __launch_bounds__(256)
gather_windows_kernel(const float* __restrict__ strain, float* __restrict__ out, long n_dw, int D,
                      long win_stride, long det_stride) {
  const long w = blockIdx.x;
  if (w >= n_dw) return;
  const float* src = strain + (w / D) * win_stride + (w % D) * det_stride;
  for (int i = threadIdx.x; i < 2048; i += 256) out[w * 2048 + i] = src[i];
}

static int run_logmel(const float* strain_contig, long n, float* out_f32, op16_t* out_tm,
                      cudaStream_t stream, const float* audio_in = nullptr, float* audio_out = nullptr) {
  GWW_TRY(logmel_tables_init());
  const int grid = (int)(n < (long)g_num_sms ? n : (long)g_num_sms);
  ProfScope ps(PK_LOGMEL, stream);
  logmel_kernel<<<grid, kLmThreads, kLmSmemBytes, stream>>>(strain_contig, n, out_f32, out_tm, g_lm, audio_in, audio_out);
  LAUNCH_CHECK();
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct LayerDev {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  op16_t *qkv_w, *o_w, *fc1_w, *fc2_w;
  float *qkv_b, *o_b, *fc1_b, *fc2_b;
  // LayerNorm folded into the consuming Linear (gemm_tc.cuh "LayerNorm folding"): W (.) gamma in bf16,
  // c1[n] = sum_k of those bf16 weights, c2[n] = sum_k beta_k W_nk + b_n
  op16_t *qkv_wf, *fc1_wf;
  float *qkv_c1, *qkv_c2, *fc1_c1, *fc1_c2;
};
struct gww_model {
  gww_encoder_config_t cfg;
  op16_t *conv1_w, *conv2_w;   // [d, 384], [d, 3d]
  op16_t* conv1_wp = nullptr;  // [d, 256]: the three taps' 80 channels contiguous (k = 80 tap + c), zero-padded
  float *conv1_b, *conv2_b, *pos_emb, *lnp_g, *lnp_b;
  std::vector<LayerDev> layers;
  std::vector<void*> owned;
  std::vector<void*> head_owned;   // device buffers of the current head (freed when the head is replaced)
  // LayerNorm-fold guard (see kLnGuardLimit): device max of |mean|/std, its pinned host mirror, and the state machine
  float* ln_guard_dev = nullptr;
  float* ln_guard_host = nullptr;
  cudaEvent_t ln_guard_ev = nullptr;
  bool ln_guard_pending = false;
  bool ln_fold_checked = false;    // the first folded chunk of this model has been verified synchronously
  bool ln_fold_off = false;        // the guard tripped: this model runs the stand-alone LayerNorm kernel
  float ln_ratio_seen = 0.f;
  bool has_head = false;
  HeadParams head{};
  float* head_scratch = nullptr;   // 2 x rows x kHeadMaxWidth ping-pong activations
  long head_scratch_rows = 0;
};

static int dev_f32(gww_model* m, const float* h, size_t n, float** out) {
  float* d = nullptr;
  CU_TRY(cudaMalloc(&d, n * sizeof(float)));
  m->owned.push_back(d);
  if (h) CU_TRY(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
  else CU_TRY(cudaMemset(d, 0, n * sizeof(float)));
  *out = d;
  return GWW_OK;
}
static int dev_bf16(gww_model* m, const std::vector<float>& h, op16_t** out) {
  std::vector<op16_t> hb(h.size());
  for (size_t i = 0; i < h.size(); ++i) hb[i] = float_to_op16(h[i]);
  op16_t* d = nullptr;
  CU_TRY(cudaMalloc(&d, hb.size() * sizeof(op16_t)));
  m->owned.push_back(d);
  CU_TRY(cudaMemcpy(d, hb.data(), hb.size() * sizeof(op16_t), cudaMemcpyHostToDevice));
  *out = d;
  return GWW_OK;
}

// W' = diag(m / ||W0 + s B A||_row) (W0 + s B A)   (PEFT 0.12 DoRA merge; fp32 with f64 norms)
static void dora_merge(const float* W0, int dout, int din, const gww_dora_t& a, std::vector<float>& out) {
  out.assign(W0, W0 + (size_t)dout * din);
  if (a.lora_A == nullptr || a.lora_B == nullptr || a.magnitude == nullptr || a.r <= 0) return;
  for (int o = 0; o < dout; ++o) {
    float* row = out.data() + (size_t)o * din;
    for (int r = 0; r < a.r; ++r) {
      const float br = a.scale * a.lora_B[(size_t)o * a.r + r];
      const float* ar = a.lora_A + (size_t)r * din;
      for (int i = 0; i < din; ++i) row[i] += br * ar[i];
    }
    double nrm = 0.0;
    for (int i = 0; i < din; ++i) nrm += (double)row[i] * row[i];
    const float sc = (float)((double)a.magnitude[o] / std::sqrt(nrm));
    for (int i = 0; i < din; ++i) row[i] *= sc;
  }
}

// folds LayerNorm(gamma, beta) into the Linear (W [N,K], b [N]) that consumes it
static void ln_fold(const std::vector<float>& W, const float* b, const float* gamma, const float* beta, int N, int K,
                    std::vector<float>& Wf, std::vector<float>& c1, std::vector<float>& c2) {
  Wf.resize((size_t)N * K);
  c1.assign(N, 0.f);
  c2.assign(N, 0.f);
  for (int n = 0; n < N; ++n) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < K; ++k) {
      const float wg = W[(size_t)n * K + k] * gamma[k];
      Wf[(size_t)n * K + k] = wg;
      s1 += (double)op16_to_float(float_to_op16(wg));   // the sum of what the tensor core will multiply by
      s2 += (double)beta[k] * (double)W[(size_t)n * K + k];
    }
    c1[n] = (float)s1;
    c2[n] = (float)(s2 + (b ? (double)b[n] : 0.0));
  }
}

extern "C" int gww_model_create(const gww_encoder_config_t* cfg, const gww_encoder_weights_t* w,
                                gww_model_t** out) {
  if (!cfg || !w || !out) return fail(GWW_ERR_INVALID, "model_create: null argument");
  GWW_TRY(gww_device_ok());
  const int d = cfg->d_model, f = cfg->ffn_dim, L = cfg->n_layers;
  if (d % 128 != 0 || cfg->n_heads * 64 != d || f % 64 != 0 || L <= 0)
    return fail(GWW_ERR_INVALID, "model_create: unsupported geometry d=%d heads=%d ffn=%d layers=%d",
                d, cfg->n_heads, f, L);
  gww_model* m = new gww_model();
  m->cfg = *cfg;
  int rc = GWW_OK;
  auto guard = [&](int r) { if (r != GWW_OK && rc == GWW_OK) rc = r; return r; };
  {
    std::vector<float> p((size_t)d * 384, 0.f);
    for (int co = 0; co < d; ++co)
      for (int ci = 0; ci < 80; ++ci)
        for (int t = 0; t < 3; ++t) p[(size_t)co * 384 + t * 128 + ci] = w->conv1_w[((size_t)co * 80 + ci) * 3 + t];
    guard(dev_bf16(m, p, &m->conv1_w));
    std::vector<float> pp((size_t)d * 256, 0.f);
    for (int co = 0; co < d; ++co)
      for (int ci = 0; ci < 80; ++ci)
        for (int t = 0; t < 3; ++t) pp[(size_t)co * 256 + t * 80 + ci] = w->conv1_w[((size_t)co * 80 + ci) * 3 + t];
    guard(dev_bf16(m, pp, &m->conv1_wp));
    std::vector<float> p2((size_t)d * 3 * d);
    for (int co = 0; co < d; ++co)
      for (int ci = 0; ci < d; ++ci)
        for (int t = 0; t < 3; ++t) p2[(size_t)co * 3 * d + (size_t)t * d + ci] = w->conv2_w[((size_t)co * d + ci) * 3 + t];
    guard(dev_bf16(m, p2, &m->conv2_w));
  }
  guard(dev_f32(m, w->conv1_b, d, &m->conv1_b));
  guard(dev_f32(m, w->conv2_b, d, &m->conv2_b));
  guard(dev_f32(m, w->pos_emb, (size_t)GWW_N_CTX * d, &m->pos_emb));
  guard(dev_f32(m, w->ln_post_g, d, &m->lnp_g));
  guard(dev_f32(m, w->ln_post_b, d, &m->lnp_b));
  m->layers.resize(L);
  for (int l = 0; l < L && rc == GWW_OK; ++l) {
    const gww_layer_weights_t& lw = w->layers[l];
    LayerDev& ld = m->layers[l];
    guard(dev_f32(m, lw.ln1_g, d, &ld.ln1_g));
    guard(dev_f32(m, lw.ln1_b, d, &ld.ln1_b));
    guard(dev_f32(m, lw.ln2_g, d, &ld.ln2_g));
    guard(dev_f32(m, lw.ln2_b, d, &ld.ln2_b));
    std::vector<float> q, k, v, o;
    dora_merge(lw.q_w, d, d, lw.dora_q, q);
    dora_merge(lw.k_w, d, d, lw.dora_k, k);
    dora_merge(lw.v_w, d, d, lw.dora_v, v);
    dora_merge(lw.o_w, d, d, lw.dora_o, o);
    std::vector<float> qkv((size_t)3 * d * d), qkvb((size_t)3 * d, 0.f);
    const float qs = 0.125f;  // head_dim^-0.5, exact in bf16; applied after bias in the reference (:310)
    for (size_t i = 0; i < (size_t)d * d; ++i) {
      qkv[i] = q[i] * qs;
      qkv[(size_t)d * d + i] = k[i];
      qkv[(size_t)2 * d * d + i] = v[i];
    }
    for (int i = 0; i < d; ++i) {
      qkvb[i] = (lw.q_b ? lw.q_b[i] : 0.f) * qs;
      qkvb[2 * d + i] = lw.v_b ? lw.v_b[i] : 0.f;
    }
    guard(dev_bf16(m, qkv, &ld.qkv_w));
    guard(dev_f32(m, qkvb.data(), 3 * d, &ld.qkv_b));
    {
      std::vector<float> wf, c1, c2;
      ln_fold(qkv, qkvb.data(), lw.ln1_g, lw.ln1_b, 3 * d, d, wf, c1, c2);
      guard(dev_bf16(m, wf, &ld.qkv_wf));
      guard(dev_f32(m, c1.data(), 3 * d, &ld.qkv_c1));
      guard(dev_f32(m, c2.data(), 3 * d, &ld.qkv_c2));
      const std::vector<float> w1(lw.fc1_w, lw.fc1_w + (size_t)f * d);
      ln_fold(w1, lw.fc1_b, lw.ln2_g, lw.ln2_b, f, d, wf, c1, c2);
      guard(dev_bf16(m, wf, &ld.fc1_wf));
      guard(dev_f32(m, c1.data(), f, &ld.fc1_c1));
      guard(dev_f32(m, c2.data(), f, &ld.fc1_c2));
    }
    guard(dev_bf16(m, o, &ld.o_w));
    guard(dev_f32(m, lw.o_b, d, &ld.o_b));
    guard(dev_bf16(m, std::vector<float>(lw.fc1_w, lw.fc1_w + (size_t)f * d), &ld.fc1_w));
    guard(dev_f32(m, lw.fc1_b, f, &ld.fc1_b));
    guard(dev_bf16(m, std::vector<float>(lw.fc2_w, lw.fc2_w + (size_t)d * f), &ld.fc2_w));
    guard(dev_f32(m, lw.fc2_b, d, &ld.fc2_b));
  }
  if (rc == GWW_OK) {
    if (cudaMalloc(&m->ln_guard_dev, sizeof(float)) != cudaSuccess || cudaMemset(m->ln_guard_dev, 0, sizeof(float)) != cudaSuccess ||
        cudaMallocHost(&m->ln_guard_host, sizeof(float)) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ln_guard_ev, cudaEventDisableTiming) != cudaSuccess)
      rc = fail(GWW_ERR_CUDA, "model_create: cannot allocate the LayerNorm-fold guard");
    else
      *m->ln_guard_host = 0.f;
  }
  if (rc != GWW_OK) {
    gww_model_destroy(m);
    return rc;
  }
  *out = m;
  return GWW_OK;
}

extern "C" int gww_model_ln_fold_state(const gww_model_t* m, float* max_ratio, int* fold_active) {
  if (!m) return fail(GWW_ERR_INVALID, "ln_fold_state: null model");
  float v = 0.f;
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(&v, m->ln_guard_dev, sizeof(float), cudaMemcpyDeviceToHost));
  if (v < m->ln_ratio_seen) v = m->ln_ratio_seen;
  if (max_ratio) *max_ratio = v;
  if (fold_active) *fold_active = (m->ln_fold_off ? 0 : 1);
  return GWW_OK;
}

extern "C" int gww_model_set_head(gww_model_t* m, const gww_head_weights_t* h) {
  if (!m || !h) return fail(GWW_ERR_INVALID, "set_head: null argument");
  if (h->n_layers < 1 || h->n_layers > kHeadMaxLayers)
    return fail(GWW_ERR_INVALID, "set_head: n_layers=%d out of range", h->n_layers);
  HeadParams hp{};
  hp.n_layers = h->n_layers;
  hp.softmax = h->softmax;
  for (int i = 0; i <= h->n_layers; ++i) {
    if (h->dims[i] < 1 || h->dims[i] > kHeadMaxWidth)
      return fail(GWW_ERR_INVALID, "set_head: layer width %d out of range", h->dims[i]);
    hp.dims[i] = h->dims[i];
  }
  // replace, not accumulate: two classifiers alternating on one encoder must not leak a head per switch.
  // cudaFree waits for in-flight work that still reads the old buffers.
  for (void* p : m->head_owned) cudaFree(p);
  m->head_owned.clear();
  m->has_head = false;
  for (int i = 0; i < h->n_layers; ++i) {
    if (!h->w[i] || !h->b[i]) return fail(GWW_ERR_INVALID, "set_head: null weight pointer in layer %d", i);
    float *dw = nullptr, *db = nullptr;
    const size_t nw = (size_t)h->dims[i] * h->dims[i + 1], nb = (size_t)h->dims[i + 1];
    CU_TRY(cudaMalloc(&dw, nw * sizeof(float)));
    m->head_owned.push_back(dw);
    CU_TRY(cudaMalloc(&db, nb * sizeof(float)));
    m->head_owned.push_back(db);
    CU_TRY(cudaMemcpy(dw, h->w[i], nw * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(db, h->b[i], nb * sizeof(float), cudaMemcpyHostToDevice));
    hp.w[i] = dw;
    hp.b[i] = db;
  }
  GWW_TRY(ensure_smem_attr(head_linear_kernel, kHeadWin * kHeadMaxWidth * 4));
  m->head = hp;
  m->has_head = true;
  return GWW_OK;
}

extern "C" void gww_model_destroy(gww_model_t* m) {
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  for (void* p : m->head_owned) cudaFree(p);
  if (m->ln_guard_dev) cudaFree(m->ln_guard_dev);
  if (m->ln_guard_host) cudaFreeHost(m->ln_guard_host);
  if (m->ln_guard_ev) cudaEventDestroy(m->ln_guard_ev);
  if (m->head_scratch) cudaFree(m->head_scratch);
  delete m;
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct Workspace {
  op16_t* feats_tm;  // [chunk, 3002, 80]
  float* x;                 // [chunk*1500, d]  residual stream
  op16_t* h;         // [chunk*1500, d]  LN output / attention output
  op16_t* g;         // [chunk*1500, ffn] fc1 output; aliases qkv [.,3d] and conv1 out [chunk,3001,d]
  float* pooled;            // [chunk, d]
  float* head_out;          // [chunk, 64]
  float* head_scratch;      // [2, chunk, kHeadMaxWidth]
  float* gather;            // [chunk, 2048] contiguous strain windows
  float* x_last;            // [chunk, d] residual rows of the last token (pruned final layer)
  op16_t* xb;        // [chunk*1500, d]  bf16 copy of the residual stream (LayerNorm fold: A operand of qkv / fc1)
  float* stats;             // [chunk*1500, kStatSlots, 2] per-row partial (sum, sum of squares) of x
  size_t total;
};
constexpr int kStatSlots = 16;   // 2 * ceil(d / block_n) <= 16 for d <= 1024 with block_n >= 128
static size_t align_up(size_t v) { return (v + 1023) & ~(size_t)1023; }
static Workspace carve(const gww_model* m, int chunk, uint8_t* base) {
  const size_t d = m->cfg.d_model, f = m->cfg.ffn_dim, M = (size_t)chunk * GWW_N_CTX;
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return base + o; };
  w.feats_tm = (op16_t*)take((size_t)chunk * 3002 * 80 * 2);
  w.x = (float*)take(M * d * 4);
  w.h = (op16_t*)take(M * d * 2);
  size_t gbytes = M * f * 2;
  const size_t h1bytes = ((size_t)chunk * 3001 + 2) * d * 2;
  if (h1bytes > gbytes) gbytes = h1bytes;
  if (M * 3 * d * 2 > gbytes) gbytes = M * 3 * d * 2;
  if (M * d * 4 > gbytes) gbytes = M * d * 4;
  w.g = (op16_t*)take(gbytes);
  w.pooled = (float*)take((size_t)chunk * d * 4);
  w.head_out = (float*)take((size_t)chunk * 64 * 4);
  w.head_scratch = (float*)take((size_t)2 * chunk * kHeadMaxWidth * 4);
  w.gather = (float*)take((size_t)chunk * 2048 * 4);
  w.x_last = (float*)take((size_t)chunk * d * 4);
  w.xb = (op16_t*)take(M * d * 2);
  w.stats = (float*)take(M * kStatSlots * 2 * 4);
  w.total = off;
  return w;
}
extern "C" size_t gww_workspace_bytes(const gww_model_t* m, int chunk) {
  if (!m || chunk <= 0) return 0;
  return carve(m, chunk, nullptr).total + 1024;
}

// ------------------------------------------------------------------------------------------------
// encoder on one chunk whose time-major bf16 features are already in ws.feats_tm
// ------------------------------------------------------------------------------------------------
// LayerNorm folded into the neighbouring GEMMs (default) or the stand-alone LayerNorm kernel (GWW_LN_FOLD=0)
static bool ln_fold_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GWW_LN_FOLD");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

static int encoder_chunk_impl(const gww_model* m, const Workspace& ws, int nc, float* last_hidden,
                              float* pooled, int use_last_token, cudaStream_t stream, bool fold);

// Runs one chunk; owns the LayerNorm-fold guard: the first folded chunk of a model is checked synchronously (and
// redone with the stand-alone LayerNorm if the residual stream's |mean|/std exceeds kLnGuardLimit), later chunks are
// checked without blocking (the running maximum comes back through pinned memory behind an event).
static int encoder_chunk(const gww_model* cm, const Workspace& ws, int nc, float* last_hidden,
                         float* pooled, int use_last_token, cudaStream_t stream) {
  gww_model* m = const_cast<gww_model*>(cm);
  bool fold = ln_fold_enabled() && !m->ln_fold_off;
  auto trip = [&](float ratio) {
    m->ln_fold_off = true;
    m->ln_ratio_seen = ratio;
    fprintf(stderr, "gww: residual stream has |mean|/std = %.2f (> %.1f): LayerNorm folding disabled for this model, "
                    "using the stand-alone LayerNorm kernel\n", ratio, kLnGuardLimit);
  };
  if (fold && m->ln_guard_pending && cudaEventQuery(m->ln_guard_ev) == cudaSuccess) {
    m->ln_guard_pending = false;
    if (*m->ln_guard_host > kLnGuardLimit) { trip(*m->ln_guard_host); fold = false; }
  }
  GWW_TRY(encoder_chunk_impl(m, ws, nc, last_hidden, pooled, use_last_token, stream, fold));
  if (!fold) return GWW_OK;
  if (!m->ln_fold_checked) {
    float v = 0.f;
    CU_TRY(cudaMemcpyAsync(&v, m->ln_guard_dev, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    m->ln_fold_checked = true;
    if (v > kLnGuardLimit) {
      trip(v);
      return encoder_chunk_impl(m, ws, nc, last_hidden, pooled, use_last_token, stream, false);
    }
  } else if (!m->ln_guard_pending) {
    CU_TRY(cudaMemcpyAsync(m->ln_guard_host, m->ln_guard_dev, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaEventRecord(m->ln_guard_ev, stream));
    m->ln_guard_pending = true;
  }
  return GWW_OK;
}

static int encoder_chunk_impl(const gww_model* m, const Workspace& ws, int nc, float* last_hidden,
                              float* pooled, int use_last_token, cudaStream_t stream, const bool fold) {
  const int d = m->cfg.d_model, f = m->cfg.ffn_dim;
  const long M = (long)nc * GWW_N_CTX;
  const int bn_d = pick_block_n(d), bn_3d = pick_block_n(3 * d), bn_f = pick_block_n(f);
  // out_proj (K = N = d) is HBM-bound (10 d bytes per row against 2 d^2 FLOP): a 128-wide N tile leaves room
  // for more operand stages in flight and measured 0.355 vs 0.449 ms (whisper-base, 384k rows; 5.5 TB/s)
  static const int bn_o_env = getenv("GWW_GEMM_BN_O") ? atoi(getenv("GWW_GEMM_BN_O")) : 0;   // tuning override
  const int bn_o = (bn_o_env == 128 || bn_o_env == 192 || bn_o_env == 256) ? bn_o_env : ((d % 128 == 0) ? 128 : bn_d);
  // producer side of a residual GEMM / consumer side of the Linear after a LayerNorm
  auto produce = [&]() { LnFold lf; if (fold) { lf.xb = ws.xb; lf.stats_out = ws.stats; lf.d_model = d; } return lf; };
  auto consume = [&](const float* c1, int slots) {
    LnFold lf; lf.stats_in = ws.stats; lf.c1 = c1; lf.in_slots = slots; lf.d_model = d; lf.guard = m->ln_guard_dev; return lf;
  };
  op16_t* h1 = ws.g;   // conv1 output [nc, 3001, d], row 0 of each sample = zero pad
  {  // zero pad row (t = -1) of every sample
    zero_rows_kernel<<<nc, 128, 0, stream>>>(reinterpret_cast<uint4*>(h1), (size_t)3001 * d * 2 / 16, d * 2 / 16);
    LAUNCH_CHECK();
  }
  {  // conv1 (k=3, pad=1) + GELU : feats_tm [nc,3002,80] -> h1[:,1:,:]
    GemmCall g{};
    g.a_base = ws.feats_tm;
    // time-major features: the three 80-channel rows t-1, t, t+1 of an output row are CONTIGUOUS (240 elements from
    // padded row t), so the im2col operand is a tensor map whose row stride (80 elements) is smaller than its row
    // length (240): K = 240 -> 4 k-blocks (columns 240..255 out of range = zero) instead of three taps x 128
    // (VERDICT r1 5c without replicating the features).  GWW_CONV1_PACKED=0: the three row-shifted boxes of round 1.
    static const bool packed = !(getenv("GWW_CONV1_PACKED") && atoi(getenv("GWW_CONV1_PACKED")) == 0);
    if (packed) {
      g.a_dims[0] = 240; g.a_dims[1] = 1; g.a_dims[2] = 3000; g.a_dims[3] = nc;
      g.a_strides[0] = 160; g.a_strides[1] = 160; g.a_strides[2] = 3002ull * 160;
      g.w_base = m->conv1_wp; g.ktot = 256;
      g.p.kb_per_tap = 4; g.p.taps = 1;
    } else {
      g.a_dims[0] = 80; g.a_dims[1] = 1; g.a_dims[2] = 3002; g.a_dims[3] = nc;
      g.a_strides[0] = 160; g.a_strides[1] = 160; g.a_strides[2] = 3002ull * 160;
      g.w_base = m->conv1_w; g.ktot = 384;
      g.p.kb_per_tap = 2; g.p.taps = 3;
    }
    g.c_base = h1 + d;
    g.c_strides[0] = (uint64_t)d * 2; g.c_strides[1] = 3001ull * d * 2;
    g.p.rows = 3000; g.p.batch = nc; g.p.n = d; g.p.p_mod = 1;
    g.p.bias = m->conv1_b; g.p.resid = nullptr; g.p.pos = nullptr;
    g.epi = EPI_BIAS_GELU_BF16; g.block_n = bn_d; g.kind = PK_GEMM_CONV1;
    GWW_TRY(run_gemm(g, stream));
  }
  {  // conv2 (k=3, stride 2, pad=1) + GELU + embed_positions : h1 -> x [nc,1500,d] f32
    GemmCall g{};
    g.a_base = h1;
    g.a_dims[0] = d; g.a_dims[1] = 2; g.a_dims[2] = 1501; g.a_dims[3] = nc;
    g.a_strides[0] = (uint64_t)d * 2; g.a_strides[1] = (uint64_t)d * 4; g.a_strides[2] = 3001ull * d * 2;
    g.w_base = m->conv2_w; g.ktot = 3 * d;
    g.c_base = ws.x;
    g.c_strides[0] = (uint64_t)d * 4; g.c_strides[1] = (uint64_t)GWW_N_CTX * d * 4;
    g.p.rows = GWW_N_CTX; g.p.batch = nc; g.p.n = d; g.p.kb_per_tap = d / 64; g.p.taps = 3; g.p.p_mod = 2;
    g.p.bias = m->conv2_b; g.p.resid = nullptr; g.p.pos = m->pos_emb;
    g.epi = EPI_BIAS_GELU_POS_F32; g.block_n = bn_d; g.kind = PK_GEMM_CONV2;
    if (fold) {   // the first layer's LN1 statistics and bf16 copy come out of the conv stem
      LnFold lf; lf.xb = ws.xb; lf.stats_out = ws.stats; lf.d_model = d;
      apply_fold(g.p, lf, bn_d);
    }
    GWW_TRY(run_gemm(g, stream));
  }
  int slots_in = stat_slots_of(d, bn_d);   // slots written by the GEMM that last produced x
  op16_t* qkv = ws.g;
  // when only last_hidden_state[:, -1, :] is consumed, the final layer needs all tokens'
  // K and V but only the last token's query row, out-projection, MLP and final LayerNorm.
  const bool prune = (last_hidden == nullptr && pooled != nullptr && use_last_token && g_prune_last.load() != 0);
  for (size_t li = 0; li < m->layers.size(); ++li) {
    const LayerDev& ld = m->layers[li];
    if (prune && li + 1 == m->layers.size()) {
      const int T = GWW_N_CTX;
      op16_t* hl = ws.h;                          // [nc, d] attention output of the last token
      op16_t* hl2 = ws.h + (size_t)nc * d;        // [nc, d] LN2 output
      // K and V projections of all tokens into columns [d, 3d) of the qkv rows (rows d.. of the fused weight); the
      // query projection only for the last token's row (stand-alone LayerNorm of that row, written in place)
      const int bn_2d = pick_block_n(2 * d);
      if (fold) {
        GWW_TRY(run_linear(ws.xb, ld.qkv_wf + (size_t)d * d, qkv + d, ld.qkv_c2 + d, nullptr, M, 2 * d, d, EPI_BIAS_BF16, bn_2d,
                           stream, PK_GEMM_QKV, consume(ld.qkv_c1 + d, slots_in), 3L * d));
      } else {
        GWW_TRY(run_ln_t<op16_t>(ws.x, ws.h, ld.ln1_g, ld.ln1_b, M, d, 0, 1, stream));
        GWW_TRY(run_linear(ws.h, ld.qkv_w + (size_t)d * d, qkv + d, ld.qkv_b + d, nullptr, M, 2 * d, d, EPI_BIAS_BF16, bn_2d, stream,
                           PK_GEMM_QKV, LnFold(), 3L * d));
      }
      GWW_TRY(run_ln_t<op16_t>(ws.x, hl2, ld.ln1_g, ld.ln1_b, nc, d, T - 1, T, stream));
      GWW_TRY(run_linear(hl2, ld.qkv_w, qkv + (size_t)(T - 1) * 3 * d, ld.qkv_b, nullptr, nc, d, d, EPI_BIAS_BF16, bn_d, stream,
                         PK_GEMM_QKV, LnFold(), (long)T * 3 * d));
      {
        ProfScope ps(PK_ATTN_LAST, stream);
        const size_t smem = (size_t)(((T + 3) & ~3) + 8 * 64 + 16) * sizeof(float);
        last_row_attention_kernel<<<dim3(d / 64, nc), 256, smem, stream>>>(qkv, hl, T, d);
        LAUNCH_CHECK();
        gather_last_rows_kernel<<<nc, 128, 0, stream>>>(ws.x, ws.x_last, T, d);
        LAUNCH_CHECK();
      }
      GWW_TRY(run_linear(hl, ld.o_w, ws.x_last, ld.o_b, ws.x_last, nc, d, d, EPI_BIAS_RESID_F32, bn_d, stream, PK_GEMM_O));
      GWW_TRY(run_ln_t<op16_t>(ws.x_last, hl2, ld.ln2_g, ld.ln2_b, nc, d, 0, 1, stream));
      GWW_TRY(run_linear(hl2, ld.fc1_w, ws.g, ld.fc1_b, nullptr, nc, f, d, EPI_BIAS_GELU_BF16, bn_f, stream, PK_GEMM_FC1));
      GWW_TRY(run_linear(ws.g, ld.fc2_w, ws.x_last, ld.fc2_b, ws.x_last, nc, d, f, EPI_BIAS_RESID_F32, bn_d, stream, PK_GEMM_FC2));
      GWW_TRY(run_ln_t<float>(ws.x_last, pooled, m->lnp_g, m->lnp_b, nc, d, 0, 1, stream));
      return GWW_OK;
    }
    if (fold) {
      // no LayerNorm kernels: x's bf16 copy and row statistics come out of the GEMM that produced x
      GWW_TRY(run_linear(ws.xb, ld.qkv_wf, qkv, ld.qkv_c2, nullptr, M, 3 * d, d, EPI_BIAS_BF16, bn_3d, stream, PK_GEMM_QKV,
                         consume(ld.qkv_c1, slots_in)));
      GWW_TRY(run_attention(qkv, ws.h, nc, GWW_N_CTX, d, stream));
      GWW_TRY(run_linear(ws.h, ld.o_w, ws.x, ld.o_b, ws.x, M, d, d, EPI_BIAS_RESID_F32, bn_o, stream, PK_GEMM_O, produce()));
      GWW_TRY(run_linear(ws.xb, ld.fc1_wf, ws.g, ld.fc1_c2, nullptr, M, f, d, EPI_BIAS_GELU_BF16, bn_f, stream, PK_GEMM_FC1,
                         consume(ld.fc1_c1, stat_slots_of(d, bn_o))));
      GWW_TRY(run_linear(ws.g, ld.fc2_w, ws.x, ld.fc2_b, ws.x, M, d, f, EPI_BIAS_RESID_F32, bn_d, stream, PK_GEMM_FC2, produce()));
      slots_in = stat_slots_of(d, bn_d);
      continue;
    }
    GWW_TRY(run_ln_t<op16_t>(ws.x, ws.h, ld.ln1_g, ld.ln1_b, M, d, 0, 1, stream));
    GWW_TRY(run_linear(ws.h, ld.qkv_w, qkv, ld.qkv_b, nullptr, M, 3 * d, d, EPI_BIAS_BF16, bn_3d, stream, PK_GEMM_QKV));
    GWW_TRY(run_attention(qkv, ws.h, nc, GWW_N_CTX, d, stream));
    GWW_TRY(run_linear(ws.h, ld.o_w, ws.x, ld.o_b, ws.x, M, d, d, EPI_BIAS_RESID_F32, bn_o, stream, PK_GEMM_O));
    GWW_TRY(run_ln_t<op16_t>(ws.x, ws.h, ld.ln2_g, ld.ln2_b, M, d, 0, 1, stream));
    GWW_TRY(run_linear(ws.h, ld.fc1_w, ws.g, ld.fc1_b, nullptr, M, f, d, EPI_BIAS_GELU_BF16, bn_f, stream, PK_GEMM_FC1));
    GWW_TRY(run_linear(ws.g, ld.fc2_w, ws.x, ld.fc2_b, ws.x, M, d, f, EPI_BIAS_RESID_F32, bn_d, stream, PK_GEMM_FC2));
  }
  if (last_hidden != nullptr)
    GWW_TRY(run_ln_t<float>(ws.x, last_hidden, m->lnp_g, m->lnp_b, M, d, 0, 1, stream));
  if (pooled != nullptr) {
    if (use_last_token) {
      GWW_TRY(run_ln_t<float>(ws.x, pooled, m->lnp_g, m->lnp_b, nc, d, GWW_N_CTX - 1, GWW_N_CTX, stream));
    } else {
      float* full = last_hidden;
      if (full == nullptr) {
        full = reinterpret_cast<float*>(ws.g);
        GWW_TRY(run_ln_t<float>(ws.x, full, m->lnp_g, m->lnp_b, M, d, 0, 1, stream));
      }
      mean_pool_kernel<0><<<nc, 256, 0, stream>>>(full, pooled, GWW_N_CTX, d);
      LAUNCH_CHECK();
    }
  }
  return GWW_OK;
}

static int check_ws(const gww_model* m, int chunk, void* workspace, size_t bytes, Workspace* ws) {
  if (!m) return fail(GWW_ERR_INVALID, "null model");
  if (chunk <= 0 || chunk > 32768) return fail(GWW_ERR_INVALID, "chunk=%d out of range", chunk);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  *ws = carve(m, chunk, base);
  const size_t need = ws->total + (size_t)(base - reinterpret_cast<uint8_t*>(workspace));
  if (workspace == nullptr || bytes < need)
    return fail(GWW_ERR_WORKSPACE, "workspace too small: have %zu need %zu bytes", bytes, need);
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// entry points
// ------------------------------------------------------------------------------------------------
extern "C" int gww_logmel_frontend(const float* strain, long n, float* feats, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !feats || n < 0) return fail(GWW_ERR_INVALID, "logmel_frontend: bad argument");
  if (n == 0) return GWW_OK;
  return run_logmel(strain, n, feats, nullptr, (cudaStream_t)stream);
}

extern "C" int gww_resample_16k(const float* strain, long n, float* audio, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !audio || n < 0) return fail(GWW_ERR_INVALID, "resample_16k: bad argument");
  if ((reinterpret_cast<uintptr_t>(audio) & 15) != 0) return fail(GWW_ERR_INVALID, "resample_16k: audio must be 16-byte aligned");
  if (n == 0) return GWW_OK;
  return run_logmel(strain, n, nullptr, nullptr, (cudaStream_t)stream, nullptr, audio);
}

extern "C" int gww_logmel_from_16k(const float* audio, long n, float* feats, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!audio || !feats || n < 0) return fail(GWW_ERR_INVALID, "logmel_from_16k: bad argument");
  if ((reinterpret_cast<uintptr_t>(audio) & 15) != 0) return fail(GWW_ERR_INVALID, "logmel_from_16k: audio must be 16-byte aligned");
  if (n == 0) return GWW_OK;
  return run_logmel(nullptr, n, feats, nullptr, (cudaStream_t)stream, audio, nullptr);
}

extern "C" int gww_encoder_forward(const gww_model_t* m, const float* feats, long n, float* last_hidden,
                                   float* pooled, int use_last_token, void* workspace,
                                   size_t workspace_bytes, int chunk, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!feats || n < 0 || (!last_hidden && !pooled)) return fail(GWW_ERR_INVALID, "encoder_forward: bad argument");
  Workspace ws;
  GWW_TRY(check_ws(m, chunk, workspace, workspace_bytes, &ws));
  cudaStream_t s = (cudaStream_t)stream;
  const int d = m->cfg.d_model;
  for (long c0 = 0; c0 < n; c0 += chunk) {
    const int nc = (int)((n - c0 < chunk) ? n - c0 : chunk);
    dim3 grid(47, nc);
    {
      ProfScope ps(PK_FEATS_TM, s);
      feats_to_timemajor_kernel<<<grid, 256, 0, s>>>(feats + c0 * 80L * 3000L, ws.feats_tm);
    }
    LAUNCH_CHECK();
    GWW_TRY(encoder_chunk(m, ws, nc, last_hidden ? last_hidden + c0 * (long)GWW_N_CTX * d : nullptr,
                          pooled ? pooled + c0 * d : nullptr, use_last_token, s));
  }
  return GWW_OK;
}

static int run_head(const gww_model* m, const float* reps, long B, float* out, cudaStream_t s,
                    float* scratch = nullptr) {
  if (!m->has_head) return fail(GWW_ERR_INVALID, "head_forward: no head set on this model");
  if (B == 0) return GWW_OK;
  const HeadParams& hp = m->head;
  if (scratch == nullptr) {   // standalone call: model-owned scratch, grown on demand (not the hot path)
    gww_model* mm = const_cast<gww_model*>(m);
    if (mm->head_scratch_rows < B) {
      CU_TRY(cudaStreamSynchronize(s));
      if (mm->head_scratch) CU_TRY(cudaFree(mm->head_scratch));
      CU_TRY(cudaMalloc(&mm->head_scratch, (size_t)2 * B * kHeadMaxWidth * sizeof(float)));
      mm->head_scratch_rows = B;
    }
    scratch = mm->head_scratch;
  }
  float* bufs[2] = {scratch, scratch + (size_t)B * kHeadMaxWidth};
  const float* cur = reps;
  const int C = hp.dims[hp.n_layers];
  ProfScope ps(PK_HEAD, s);
  for (int L = 0; L < hp.n_layers; ++L) {
    const int din = hp.dims[L], dout = hp.dims[L + 1];
    const bool last = (L == hp.n_layers - 1);
    float* dst = (last && !hp.softmax) ? out : bufs[L & 1];
    dim3 grid((dout + 7) / 8, (unsigned)((B + kHeadWin - 1) / kHeadWin));
    head_linear_kernel<<<grid, 256, (size_t)kHeadWin * din * sizeof(float), s>>>(
        cur, hp.w[L], hp.b[L], dst, (int)B, din, dout, last ? 0 : 1);
    LAUNCH_CHECK();
    cur = dst;
  }
  if (hp.softmax) {
    row_softmax_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(cur, out, (int)B, C);
    LAUNCH_CHECK();
  }
  return GWW_OK;
}

extern "C" int gww_head_forward(const gww_model_t* m, const float* reps, long B, float* out, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!m || !reps || !out || B < 0) return fail(GWW_ERR_INVALID, "head_forward: bad argument");
  return run_head(m, reps, B, out, (cudaStream_t)stream);
}

// strain windows addressed as strain + b*win_stride + i*det_stride -> out [B, C]
static int forward_logmel_strided(const gww_model* m, const float* strain, long B, int D,
                                  long win_stride, long det_stride, bool contiguous, float* out,
                                  float* pooled_out, const Workspace& ws, int chunk, cudaStream_t s) {
  if (!m->has_head) return fail(GWW_ERR_INVALID, "forward: no head set on this model");
  if (m->head.dims[0] != m->cfg.d_model * D)
    return fail(GWW_ERR_INVALID, "forward: head input width %d != d_model*D = %d", m->head.dims[0],
                m->cfg.d_model * D);
  const int wchunk = chunk / D;   // whole windows per chunk
  if (wchunk < 1) return fail(GWW_ERR_INVALID, "forward: chunk=%d smaller than D=%d", chunk, D);
  const int C = m->head.dims[m->head.n_layers];
  const int d = m->cfg.d_model;
  for (long b0 = 0; b0 < B; b0 += wchunk) {
    const int nb = (int)((B - b0 < wchunk) ? B - b0 : wchunk);
    const int nc = nb * D;
    const float* src;
    if (contiguous) {
      src = strain + b0 * win_stride;
    } else {
      gather_windows_kernel<<<nc, 256, 0, s>>>(strain + b0 * win_stride, ws.gather, nc, D, win_stride, det_stride);
      LAUNCH_CHECK();
      src = ws.gather;
    }
    GWW_TRY(run_logmel(src, nc, nullptr, ws.feats_tm, s));
    GWW_TRY(encoder_chunk(m, ws, nc, nullptr, ws.pooled, 1, s));
    if (pooled_out)
      CU_TRY(cudaMemcpyAsync(pooled_out + b0 * D * d, ws.pooled, (size_t)nc * d * 4, cudaMemcpyDeviceToDevice, s));
    GWW_TRY(run_head(m, ws.pooled, nb, out + b0 * C, s, ws.head_scratch));
  }
  return GWW_OK;
}

extern "C" int gww_forward_windows_logmel(const gww_model_t* m, const float* strain, long B, int D,
                                          float* out, float* pooled_out, void* workspace,
                                          size_t workspace_bytes, int chunk, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !out || B < 0 || D < 1) return fail(GWW_ERR_INVALID, "forward_windows: bad argument");
  Workspace ws;
  GWW_TRY(check_ws(m, chunk, workspace, workspace_bytes, &ws));
  return forward_logmel_strided(m, strain, B, D, (long)D * 2048, 2048, true, out, pooled_out, ws, chunk,
                                (cudaStream_t)stream);
}

__global__ void take_col0_kernel(const float* __restrict__ out, int C, long n, float* __restrict__ scores) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scores[i] = out[i * C];
}

extern "C" int gww_threshold_compact(const float* out, int C, long n, float thr, long idx_base,
                                     long* trig_idx, float* trig_score, int* trig_count, int capacity,
                                     void* stream) {
  GWW_TRY(gww_device_ok());
  if (!out || !trig_idx || !trig_score || !trig_count || n < 0 || n > 0x7fffffffL)
    return fail(GWW_ERR_INVALID, "threshold_compact: bad argument");
  if (n == 0) return GWW_OK;
  threshold_compact_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(out, C, (int)n, thr, idx_base, trig_idx,
                                                                trig_score, trig_count, capacity);
  LAUNCH_CHECK();
  return GWW_OK;
}

extern "C" int gww_stream_search_logmel(const gww_model_t* m, const float* strain, int D, long n_samples,
                                        int hop, long first_window, long n_windows, float thr,
                                        float* scores, long* trig_idx, float* trig_score,
                                        int* trig_count, int capacity, void* workspace,
                                        size_t workspace_bytes, int chunk, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !scores || D < 1 || hop < 1 || n_windows < 0 || first_window < 0)
    return fail(GWW_ERR_INVALID, "stream_search: bad argument");
  if (n_samples < 2048 || (first_window + n_windows - 1) * hop + 2048 > n_samples)
    if (n_windows > 0) return fail(GWW_ERR_INVALID, "stream_search: windows exceed the segment");
  Workspace ws;
  GWW_TRY(check_ws(m, chunk, workspace, workspace_bytes, &ws));
  if (!m->has_head) return fail(GWW_ERR_INVALID, "stream_search: no head set");
  cudaStream_t s = (cudaStream_t)stream;
  const int C = m->head.dims[m->head.n_layers];
  if (C > 64) return fail(GWW_ERR_INVALID, "stream_search: head has %d outputs (> 64)", C);
  const int wchunk = chunk / D;
  if (wchunk < 1) return fail(GWW_ERR_INVALID, "stream_search: chunk smaller than D");
  for (long k0 = 0; k0 < n_windows; k0 += wchunk) {
    const int nb = (int)((n_windows - k0 < wchunk) ? n_windows - k0 : wchunk);
    const float* base = strain + (first_window + k0) * hop;
    GWW_TRY(forward_logmel_strided(m, base, nb, D, hop, n_samples, false, ws.head_out, nullptr, ws, chunk, s));
    take_col0_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ws.head_out, C, nb, scores + k0);
    LAUNCH_CHECK();
    if (trig_idx && trig_score && trig_count)
      GWW_TRY(gww_threshold_compact(ws.head_out, C, nb, thr, first_window + k0, trig_idx, trig_score,
                                    trig_count, capacity, stream));
  }
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// front end B: Q-transform + Q-Adapter
// ------------------------------------------------------------------------------------------------
struct QRowHost {
  QRow r;
  double q;
  float freq;
};
struct gww_qfront {
  int spec_f = 512, spec_t = 512;
  int n_planes = 0, n_rows = 0, n_tiles = 0;
  std::vector<double> qs;
  std::vector<QRowHost> rows;            // plane-major, ascending frequency
  std::vector<QRow> h_sorted, h_orig;     // host copies of the device plan tables
  std::vector<float> h_window;
  QPlan plan{};
  bool on_device = false;                // tables uploaded (done lazily: the plan itself needs no GPU)
  bool has_adapter = false;
  QAdapterDev ad{};
  int n_detectors = 0;
  int c1 = 16, c2 = 32, c3 = 64;         // adapter CNN widths (inference.py:322-332; train.py:118-123 uses 32/64/128)
  std::vector<void*> owned;
};

template <typename T>
static int qf_upload(gww_qfront* qf, const std::vector<T>& h, const T** dptr) {
  T* d = nullptr;
  CU_TRY(cudaMalloc(&d, (h.empty() ? 1 : h.size()) * sizeof(T)));
  qf->owned.push_back(d);
  if (!h.empty()) CU_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = d;
  return GWW_OK;
}

extern "C" void gww_qfront_destroy(gww_qfront_t* qf) {
  if (!qf) return;
  for (void* p : qf->owned) cudaFree(p);
  delete qf;
}

// Tiling plan of ml4gw QScan / GWpy QTiling (oracle/qscan.py; SURVEY.md section 8a "Q1 tiling plan").
extern "C" int gww_qfront_create(double duration, double sample_rate, double qmin, double qmax,
                                 double mismatch, int spec_f, int spec_t, gww_qfront_t** out) {
  if (!out) return fail(GWW_ERR_INVALID, "qfront_create: null argument");
  if (duration * sample_rate != 2048.0)
    return fail(GWW_ERR_INVALID, "qfront_create: only duration*sample_rate == 2048 samples is supported");
  if (!(qmin > 0 && qmax > qmin) || mismatch <= 0 || spec_t < 64 || spec_t > 512 || (spec_t % 64) != 0 ||
      spec_f < 64 || spec_f > 512 || (spec_f % 64) != 0)
    return fail(GWW_ERR_INVALID, "qfront_create: unsupported parameters (spectrogram sides must be multiples of 64 in [64, 512])");
  const double PI = 3.14159265358979323846;
  gww_qfront* qf = new gww_qfront();
  qf->spec_f = spec_f; qf->spec_t = spec_t;
  const double deltam = 2.0 * std::sqrt(mismatch / 3.0);
  const double cumum = std::log(qmax / qmin) / std::sqrt(2.0);
  const int nplanes = (int)std::fmax(std::ceil(cumum / deltam), 1.0);
  const double dq = cumum / nplanes;
  if (nplanes > 8) { delete qf; return fail(GWW_ERR_INVALID, "qfront_create: more than 8 Q planes"); }
  std::vector<float> window;
  int eoff = 0;
  for (int ip = 0; ip < nplanes; ++ip) {
    const double q = qmin * std::exp(std::sqrt(2.0) * dq * (ip + 0.5));
    qf->qs.push_back(q);
    const double qprime = q / std::sqrt(11.0);
    const double minf = 50.0 * q / (2.0 * PI * duration);
    const double maxf = sample_rate / 2.0 / (1.0 + 1.0 / qprime);
    const double fcum = std::log(maxf / minf) * std::sqrt(2.0 + q * q) / 2.0;
    const int nfreq = (int)std::fmax(1.0, std::ceil(fcum / deltam));
    const double fstep = fcum / nfreq;
    const float fstepmin = (float)(1.0 / duration);
    const double base = std::exp(2.0 / std::sqrt(2.0 + q * q) * fstep);
    std::vector<float> freqs;
    for (int i = 0; i < nfreq; ++i) {
      const float f32 = (float)(minf * std::pow(base, i + 0.5));     // torch.Tensor([...]) is float32
      const float fl = std::floor(f32 / fstepmin) * fstepmin;
      if (freqs.empty() || fl != freqs.back()) freqs.push_back(fl);   // ascending input: unique == dedup
    }
    for (float f32 : freqs) {
      const double f = (double)f32;
      QRowHost rh{};
      rh.q = q; rh.freq = f32;
      const int ws = 2 * (int)(f / qprime * duration) + 1;
      const double tcum = duration * 2.0 * PI * f / q;
      const int log2n = (int)std::ceil(std::log2(tcum / deltam));
      const int n = 1 << log2n;
      if (n < 32 || n > 2048 || ws > n) { gww_qfront_destroy(qf); return fail(GWW_ERR_INVALID, "qfront_create: row with %d tiles / window %d unsupported", n, ws); }
      const int pad = n - ws;
      const int half = (ws - 1) / 2;
      rh.r.n = n; rh.r.log2n = log2n; rh.r.ws = ws;
      rh.r.left = (int)((pad - 1) / 2.0);
      rh.r.idx0 = (int)std::nearbyint((double)(-half) + 1.0 + f * duration);
      rh.r.woff = (int)window.size();
      rh.r.eoff = eoff;
      rh.r.plane = ip;
      if (rh.r.idx0 < 0 || rh.r.idx0 + ws - 1 > 1024) { gww_qfront_destroy(qf); return fail(GWW_ERR_INVALID, "qfront_create: window of f=%g exceeds the spectrum", f); }
      // window in float32 with torch's operation order (QTile.get_window)
      const double norm = (double)n / (duration * sample_rate) * std::sqrt(315.0 * qprime / (128.0 * f));
      for (int i = 0; i < ws; ++i) {
        const float wf = (float)(i - half) / (float)duration;
        const float xf = (wf * (float)qprime) / (float)f;
        const float a = 1.0f - xf * xf;
        window.push_back((a * a) * (float)norm);
      }
      eoff += n;
      qf->rows.push_back(rh);
    }
  }
  qf->n_planes = nplanes;
  qf->n_rows = (int)qf->rows.size();
  qf->n_tiles = eoff;
  // device plan
  std::vector<QRow> orig, sorted_rows;
  for (const QRowHost& r : qf->rows) orig.push_back(r.r);
  for (int pass = 0; pass < 2; ++pass)
    for (int want = (pass == 0 ? 512 : 2048); want >= (pass == 0 ? 32 : 1024); want >>= 1)
      for (const QRow& r : orig)
        if (r.n == want) sorted_rows.push_back(r);
  int n_warp = 0;
  for (const QRow& r : sorted_rows) n_warp += (r.n <= 512) ? 1 : 0;
  QPlan& pl = qf->plan;
  qf->h_sorted = sorted_rows; qf->h_orig = orig; qf->h_window = window;
  pl.n_rows = qf->n_rows; pl.n_rows_warp = n_warp; pl.n_tiles = qf->n_tiles; pl.n_planes = nplanes;
  int maxrows = 0;
  for (int ip = 0, r0 = 0; ip < nplanes; ++ip) {
    int cnt = 0;
    for (const QRow& r : orig) cnt += (r.plane == ip) ? 1 : 0;
    pl.plane_row0[ip] = r0; pl.plane_nrows[ip] = cnt;
    r0 += cnt;
    if (cnt > maxrows) maxrows = cnt;
  }
  if (maxrows > kQiMaxRows) { gww_qfront_destroy(qf); return fail(GWW_ERR_INVALID, "qfront_create: plane with %d rows (> %d)", maxrows, kQiMaxRows); }
  *out = qf;
  return GWW_OK;
}

// uploads the plan tables and opts the kernels into their shared-memory sizes (first compute call)
static int qf_ensure_device(gww_qfront* qf) {
  if (qf->on_device) return GWW_OK;
  GWW_TRY(gww_device_ok());
  const double PI = 3.14159265358979323846;
  std::vector<float2> tw(2048);
  for (int k = 0; k < 2048; ++k)
    tw[k] = make_float2((float)std::cos(2.0 * PI * k / 2048.0), (float)std::sin(2.0 * PI * k / 2048.0));
  QPlan& pl = qf->plan;
  GWW_TRY(qf_upload(qf, qf->h_sorted, &pl.rows));
  GWW_TRY(qf_upload(qf, qf->h_orig, &pl.orig));
  GWW_TRY(qf_upload(qf, qf->h_window, &pl.window));
  GWW_TRY(qf_upload(qf, tw, &pl.tw2048));
  GWW_TRY(ensure_smem_attr(qscan_tiles_kernel, kQsSmemBytes));
  GWW_TRY(ensure_smem_attr(qscan_interp_kernel, kQiMaxRows * 512 * 4));
  GWW_TRY(ensure_smem_attr(qadapter_conv2_kernel, kC2SmemBytes));
  GWW_TRY(ensure_smem_attr(qadapter_conv3_kernel, kC3SmemBytes));
  qf->on_device = true;
  return GWW_OK;
}

extern "C" int gww_qfront_info(const gww_qfront_t* qf, int* n_planes, int* n_rows, int* n_tiles) {
  if (!qf) return fail(GWW_ERR_INVALID, "qfront_info: null handle");
  if (n_planes) *n_planes = qf->n_planes;
  if (n_rows) *n_rows = qf->n_rows;
  if (n_tiles) *n_tiles = qf->n_tiles;
  return GWW_OK;
}

extern "C" int gww_qfront_plan(const gww_qfront_t* qf, double* q_of_plane, int* plane_of_row, float* freq,
                               int* ntiles, int* windowsize, int* tile_offset) {
  if (!qf) return fail(GWW_ERR_INVALID, "qfront_plan: null handle");
  if (q_of_plane) for (int i = 0; i < qf->n_planes; ++i) q_of_plane[i] = qf->qs[i];
  for (int i = 0; i < qf->n_rows; ++i) {
    const QRowHost& r = qf->rows[i];
    if (plane_of_row) plane_of_row[i] = r.r.plane;
    if (freq) freq[i] = r.freq;
    if (ntiles) ntiles[i] = r.r.n;
    if (windowsize) windowsize[i] = r.r.ws;
    if (tile_offset) tile_offset[i] = r.r.eoff;
  }
  return GWW_OK;
}

extern "C" int gww_qfront_set_adapter(gww_qfront_t* qf, const gww_qadapter_weights_t* w) {
  if (!qf || !w) return fail(GWW_ERR_INVALID, "qfront_set_adapter: null argument");
  if (!w->conv1_w || !w->conv1_b || !w->conv2_w || !w->conv2_b || !w->conv3_w || !w->conv3_b ||
      !w->conv4_w || !w->conv4_b || !w->film_gamma || !w->film_beta)
    return fail(GWW_ERR_INVALID, "qfront_set_adapter: null weight pointer");
  if (w->n_detectors < 1 || w->n_detectors > 8)
    return fail(GWW_ERR_INVALID, "qfront_set_adapter: n_detectors=%d out of range", w->n_detectors);
  GWW_TRY(qf_ensure_device(qf));
  const int C1 = w->c1 > 0 ? w->c1 : 16, C2 = w->c2 > 0 ? w->c2 : 32, C3 = w->c3 > 0 ? w->c3 : 64;
  if (C1 % 16 || C2 % 16 || C3 % 16 || C1 > 64 || C2 > 128 || C3 > 256)
    return fail(GWW_ERR_INVALID, "qfront_set_adapter: channel widths (%d, %d, %d) must be multiples of 16, <= (64, 128, 256)", C1, C2, C3);
  qf->c1 = C1; qf->c2 = C2; qf->c3 = C3;
  const bool std_geom = (C1 == 16 && C2 == 32 && C3 == 64);
  std::vector<float> w1(9 * C1), b1(w->conv1_b, w->conv1_b + C1), w2((size_t)9 * C1 * C2), b2(w->conv2_b, w->conv2_b + C2),
      w3((size_t)9 * C2 * C3), b3(w->conv3_b, w->conv3_b + C3), w4(w->conv4_w, w->conv4_w + C3);
  for (int c = 0; c < C1; ++c)
    for (int t = 0; t < 9; ++t) w1[t * C1 + c] = w->conv1_w[c * 9 + t];
  for (int co = 0; co < C2; ++co)
    for (int ci = 0; ci < C1; ++ci)
      for (int t = 0; t < 9; ++t) w2[((size_t)t * C1 + ci) * C2 + co] = w->conv2_w[((size_t)co * C1 + ci) * 9 + t];
  for (int co = 0; co < C3; ++co)
    for (int ci = 0; ci < C2; ++ci)
      for (int t = 0; t < 9; ++t) w3[((size_t)t * C2 + ci) * C3 + co] = w->conv3_w[((size_t)co * C2 + ci) * 9 + t];
  QAdapterDev ad{};
  GWW_TRY(qf_upload(qf, w1, &ad.w1)); GWW_TRY(qf_upload(qf, b1, &ad.b1));
  GWW_TRY(qf_upload(qf, w2, &ad.w2)); GWW_TRY(qf_upload(qf, b2, &ad.b2));
  GWW_TRY(qf_upload(qf, w3, &ad.w3)); GWW_TRY(qf_upload(qf, b3, &ad.b3));
  GWW_TRY(qf_upload(qf, w4, &ad.w4));
  ad.w2p = nullptr; ad.w3p = nullptr;
  if (std_geom) {  // bf16 hi/lo images of the conv2 / conv3 weights in the UMMA canonical layout (tensor-core path)
    uint4 *p2 = nullptr, *p3 = nullptr;
    CU_TRY(cudaMalloc(&p2, (size_t)QtCfg<16, 32>::kWBytes));
    qf->owned.push_back(p2);
    CU_TRY(cudaMalloc(&p3, (size_t)QtCfg<32, 64>::kWBytes));
    qf->owned.push_back(p3);
    qt_pack_weights_kernel<16, 32><<<(QtCfg<16, 32>::kWElems + 255) / 256, 256>>>(ad.w2, p2);
    LAUNCH_CHECK();
    qt_pack_weights_kernel<32, 64><<<(QtCfg<32, 64>::kWElems + 255) / 256, 256>>>(ad.w3, p3);
    LAUNCH_CHECK();
    CU_TRY(cudaDeviceSynchronize());
    ad.w2p = p2; ad.w3p = p3;
  }
  ad.b4 = w->conv4_b[0];
  ad.scale = w->scale; ad.bias = w->bias;
  for (int i = 0; i < 8; ++i) { ad.gamma[i] = 1.f; ad.beta[i] = 0.f; }
  for (int i = 0; i < w->n_detectors; ++i) { ad.gamma[i] = w->film_gamma[i]; ad.beta[i] = w->film_beta[i]; }
  qf->ad = ad;
  qf->n_detectors = w->n_detectors;
  qf->has_adapter = true;
  return GWW_OK;
}

struct QWorkspace {
  float* tiles;            // [n, n_tiles]
  unsigned int* plane_max; // [8]
  int* plane_idx;          // [1]
  float* spec;             // [n, F, T]
  float* act1;             // [n, F/2, T/2, 16]
  float* act2;             // [n, F/4, T/4, 32]
  float* map;              // [n, F/4, T/4]
  size_t total;
};
static QWorkspace qcarve(const gww_qfront* qf, long n, uint8_t* base) {
  QWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return base + o; };
  const size_t F = qf->spec_f, T = qf->spec_t, N = (size_t)n;
  w.tiles = (float*)take(N * qf->n_tiles * 4);
  w.plane_max = (unsigned int*)take(64);
  w.plane_idx = (int*)take(64);
  w.spec = (float*)take(N * F * T * 4);
  w.act1 = (float*)take(N * (F / 2) * (T / 2) * (size_t)qf->c1 * 4);
  w.act2 = (float*)take(N * (F / 4) * (T / 4) * (size_t)qf->c2 * 4);
  w.map = (float*)take(N * (F / 4) * (T / 4) * 4);
  w.total = off;
  return w;
}
extern "C" size_t gww_qfront_workspace_bytes(const gww_qfront_t* qf, long n) {
  if (!qf || n <= 0) return 0;
  return qcarve(qf, n, nullptr).total + 1024;
}
static int qcheck_ws(const gww_qfront* qf, long n, void* workspace, size_t bytes, QWorkspace* ws) {
  if (!qf) return fail(GWW_ERR_INVALID, "null qfront handle");
  GWW_TRY(qf_ensure_device(const_cast<gww_qfront*>(qf)));
  if (n <= 0 || n > 65535) return fail(GWW_ERR_INVALID, "qfront: n=%ld out of range (1..65535 windows per call)", n);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  *ws = qcarve(qf, n, base);
  const size_t need = ws->total + (size_t)(base - reinterpret_cast<uint8_t*>(workspace));
  if (workspace == nullptr || bytes < need)
    return fail(GWW_ERR_WORKSPACE, "qfront workspace too small: have %zu need %zu bytes", bytes, need);
  return GWW_OK;
}

// one QScan call over n windows: tiles (all planes) -> plane choice -> spec
static int run_qscan(const gww_qfront* qf, const float* strain, long n, long win_stride, float* spec,
                     float* tiles, int* plane_idx, const QWorkspace& ws, cudaStream_t s) {
  ProfScope ps(PK_QSCAN, s);
  CU_TRY(cudaMemsetAsync(ws.plane_max, 0, 64, s));
  qscan_tiles_kernel<<<(unsigned)n, kQsThreads, kQsSmemBytes, s>>>(strain, n, win_stride, tiles, ws.plane_max, qf->plan);
  LAUNCH_CHECK();
  const int R = kQiMaxRows;
  qscan_interp_kernel<<<dim3((unsigned)n, kQiSplit), kQiThreads, (size_t)R * 512 * 4, s>>>(
      tiles, ws.plane_max, spec, plane_idx, qf->spec_f, qf->spec_t, qf->plan);
  LAUNCH_CHECK();
  return GWW_OK;
}

// tensor-core convolutions (default) vs the fp32 CUDA-core kernels of round 1 (GWW_QADAPTER_TC=0)
static bool qadapter_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GWW_QADAPTER_TC");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}

// adapter CNN on spec [n,F,T] -> f32 [n,80,3000] and/or bf16 time-major rows of feats_tm
static int run_qadapter(const gww_qfront* qf, const float* spec, long n, int det, float* feats_f32,
                        op16_t* feats_tm, long tm_stride_w, long tm_off, const QWorkspace& ws,
                        cudaStream_t s) {
  if (!qf->has_adapter) return fail(GWW_ERR_INVALID, "qadapter: no adapter weights set");
  if (det < 0 || det >= qf->n_detectors) return fail(GWW_ERR_INVALID, "qadapter: det_idx=%d out of range", det);
  const int F = qf->spec_f, T = qf->spec_t;
  const bool std_geom = (qf->c1 == 16 && qf->c2 == 32 && qf->c3 == 64);
  if (!std_geom) {
    // other adapter widths (MLGWSC-1/train.py geometry): generic fp32 kernels
    ProfScope ps(PK_QADAPTER, s);
    const int C1 = qf->c1, C2 = qf->c2, C3 = qf->c3;
    auto k1 = qadapter_conv_generic_kernel<true, false>;
    auto k3 = qadapter_conv_generic_kernel<false, true>;
    GWW_TRY(ensure_smem_attr(k1, 18 * 18 * 128 * 4));
    GWW_TRY(ensure_smem_attr(k3, 18 * 18 * 128 * 4));
    k1<<<dim3((T + 15) / 16, (F + 15) / 16, (unsigned)n), 256, (size_t)18 * 18 * 1 * 4, s>>>(spec, qf->ad.w1, qf->ad.b1, nullptr, 0.f,
                                                                                         ws.act1, F, T, 1, C1);
    LAUNCH_CHECK();
    k1<<<dim3((T / 2 + 15) / 16, (F / 2 + 15) / 16, (unsigned)n), 256, (size_t)18 * 18 * C1 * 4, s>>>(ws.act1, qf->ad.w2, qf->ad.b2, nullptr,
                                                                                                   0.f, ws.act2, F / 2, T / 2, C1, C2);
    LAUNCH_CHECK();
    k3<<<dim3((T / 4 + 15) / 16, (F / 4 + 15) / 16, (unsigned)n), 256, (size_t)18 * 18 * C2 * 4, s>>>(ws.act2, qf->ad.w3, qf->ad.b3, qf->ad.w4,
                                                                                                   qf->ad.b4, ws.map, F / 4, T / 4, C2, C3);
    LAUNCH_CHECK();
  } else if (qadapter_tc_enabled() && F % 64 == 0 && T % 32 == 0) {
    // tensor-core path: conv1 (CUDA cores, fp32) writes bf16 hi/lo planes; conv2 / conv3 are tcgen05 implicit GEMMs
    if (qadapter_conv1_tma_enabled()) {
      CUtensorMap tm1;
      const uint64_t d1[3] = {(uint64_t)T, (uint64_t)F, (uint64_t)n};
      const uint64_t s1[2] = {(uint64_t)T * 4, (uint64_t)T * F * 4};
      const uint32_t b1[3] = {kC1Pitch, 34, 1};
      GWW_TRY(make_map(&tm1, true, 3, spec, d1, s1, b1, false));
      const long tiles1 = (long)(F / 32) * (T / 32) * n;
      const int grid1 = (int)std::min<long>(4L * g_num_sms, tiles1);
      ProfScope ps(PK_QA_CONV1, s);
      qadapter_conv1_tma_kernel<<<grid1, 256, 0, s>>>(tm1, ws.act1, F, T, n, qf->ad);
    } else {
      ProfScope ps(PK_QA_CONV1, s);
      qadapter_conv1_kernel<true><<<dim3(T / 32, F / 32, (unsigned)n), 256, 0, s>>>(spec, ws.act1, F, T, qf->ad);
    }
    LAUNCH_CHECK();
    using C2 = QtCfg<16, 32>;
    using C3 = QtCfg<32, 64>;
    auto k2 = qadapter_conv_tc_kernel<16, 32, 0>;
    auto k3 = qadapter_conv_tc_kernel<32, 64, 1>;
    GWW_TRY(ensure_smem_attr(k2, C2::kSmemBytes));
    GWW_TRY(ensure_smem_attr(k3, C3::kSmemBytes));
    const long tiles2 = (long)(F / 2 / kQtTileH) * (T / 2 / kQtTileW) * n, tiles3 = (long)(F / 4 / kQtTileH) * (T / 4 / kQtTileW) * n;
    const int sms = g_num_sms;
    const int grid2 = (int)std::min<long>(sms, (tiles2 + C2::kGroups - 1) / C2::kGroups);
    const int grid3 = (int)std::min<long>(sms, (tiles3 + C3::kGroups - 1) / C3::kGroups);
    // the plane-format activations as 4-D maps {4 W words, H, planes, n}: one haloed tile = one TMA box
    CUtensorMap tm2, tm3;
    {
      const int H2 = F / 2, W2 = T / 2, H3 = F / 4, W3 = T / 4;
      const uint64_t d2[4] = {(uint64_t)4 * W2, (uint64_t)H2, (uint64_t)C2::kPlanes, (uint64_t)n};
      const uint64_t s2[3] = {(uint64_t)16 * W2, (uint64_t)16 * W2 * H2, (uint64_t)16 * W2 * H2 * C2::kPlanes};
      const uint32_t b2[4] = {4 * kQtHaloW, kQtHaloH, (uint32_t)C2::kPlanes, 1};
      GWW_TRY(make_map(&tm2, true, 4, ws.act1, d2, s2, b2, false));
      const uint64_t d3[4] = {(uint64_t)4 * W3, (uint64_t)H3, (uint64_t)C3::kPlanes, (uint64_t)n};
      const uint64_t s3[3] = {(uint64_t)16 * W3, (uint64_t)16 * W3 * H3, (uint64_t)16 * W3 * H3 * C3::kPlanes};
      const uint32_t b3[4] = {4 * kQtHaloW, kQtHaloH, (uint32_t)C3::kPlanes, 1};
      GWW_TRY(make_map(&tm3, true, 4, ws.act2, d3, s3, b3, false));
    }
    {
      ProfScope ps(PK_QA_CONV2, s);
      k2<<<grid2, C2::kThreads, C2::kSmemBytes, s>>>(tm2, qf->ad.w2p, qf->ad.b2, nullptr, 0.f, ws.act2, F / 2, T / 2, n);
    }
    LAUNCH_CHECK();
    {
      ProfScope ps(PK_QA_CONV3, s);
      k3<<<grid3, C3::kThreads, C3::kSmemBytes, s>>>(tm3, qf->ad.w3p, qf->ad.b3, qf->ad.w4, qf->ad.b4, ws.map, F / 4, T / 4, n);
    }
    LAUNCH_CHECK();
  } else {
    ProfScope ps(PK_QADAPTER, s);
    qadapter_conv1_kernel<false><<<dim3(T / 32, F / 32, (unsigned)n), 256, 0, s>>>(spec, ws.act1, F, T, qf->ad);
    LAUNCH_CHECK();
    qadapter_conv2_kernel<<<dim3(T / 32, F / 32, (unsigned)n), 256, kC2SmemBytes, s>>>(ws.act1, ws.act2, F / 2, T / 2, qf->ad);
    LAUNCH_CHECK();
    qadapter_conv3_kernel<<<dim3(T / 64, F / 64, (unsigned)n), 256, kC3SmemBytes, s>>>(ws.act2, ws.map, F / 4, T / 4, qf->ad);
    LAUNCH_CHECK();
  }
  {
    ProfScope ps(PK_QA_POOL, s);
    qadapter_pool_kernel<<<dim3((GWW_N_FRAMES + 127) / 128, (unsigned)n), 256, 0, s>>>(
        ws.map, feats_f32, feats_tm, tm_stride_w, tm_off, F / 4, T / 4, GWW_N_MELS, GWW_N_FRAMES, det, qf->ad);
  }
  LAUNCH_CHECK();
  return GWW_OK;
}

extern "C" int gww_qscan(const gww_qfront_t* qf, const float* strain, long n, long win_stride, float* spec,
                         float* tiles, int* plane_idx, void* workspace, size_t workspace_bytes,
                         void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !spec || win_stride < 1) return fail(GWW_ERR_INVALID, "qscan: bad argument");
  if (n == 0) return GWW_OK;
  QWorkspace ws;
  GWW_TRY(qcheck_ws(qf, n, workspace, workspace_bytes, &ws));
  return run_qscan(qf, strain, n, win_stride, spec, tiles ? tiles : ws.tiles, plane_idx ? plane_idx : ws.plane_idx,
                   ws, (cudaStream_t)stream);
}

extern "C" int gww_qadapter(const gww_qfront_t* qf, const float* spec, long n, int det_idx, float* feats_f32,
                            void* workspace, size_t workspace_bytes, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!spec || !feats_f32) return fail(GWW_ERR_INVALID, "qadapter: bad argument");
  if (n == 0) return GWW_OK;
  QWorkspace ws;
  GWW_TRY(qcheck_ws(qf, n, workspace, workspace_bytes, &ws));
  return run_qadapter(qf, spec, n, det_idx, feats_f32, nullptr, 0, 0, ws, (cudaStream_t)stream);
}

// strain windows: window b of detector i at strain + b*win_stride + i*det_stride
static int forward_qscan_strided(const gww_model* m, const gww_qfront* qf, const float* strain, long B, int D,
                                 long win_stride, long det_stride, int use_last_token, float* out,
                                 const Workspace& ws, const QWorkspace& qws, cudaStream_t s) {
  if (!m->has_head) return fail(GWW_ERR_INVALID, "forward_qscan: no head set on this model");
  if (m->head.dims[0] != m->cfg.d_model * D)
    return fail(GWW_ERR_INVALID, "forward_qscan: head input width %d != d_model*D = %d", m->head.dims[0],
                m->cfg.d_model * D);
  if (D > qf->n_detectors) return fail(GWW_ERR_INVALID, "forward_qscan: D=%d > adapter detectors %d", D, qf->n_detectors);
  for (int i = 0; i < D; ++i) {
    GWW_TRY(run_qscan(qf, strain + i * det_stride, B, win_stride, qws.spec, qws.tiles, qws.plane_idx, qws, s));
    GWW_TRY(run_qadapter(qf, qws.spec, B, i, nullptr, ws.feats_tm, D, i, qws, s));
  }
  GWW_TRY(encoder_chunk(m, ws, (int)(B * D), nullptr, ws.pooled, use_last_token, s));
  return run_head(m, ws.pooled, B, out, s, ws.head_scratch);
}

extern "C" int gww_forward_windows_qscan(const gww_model_t* m, const gww_qfront_t* qf, const float* strain,
                                         long B, int D, int use_last_token, float* out, void* workspace,
                                         size_t workspace_bytes, void* q_workspace, size_t q_workspace_bytes,
                                         void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !out || B < 0 || D < 1) return fail(GWW_ERR_INVALID, "forward_windows_qscan: bad argument");
  if (B == 0) return GWW_OK;
  Workspace ws;
  GWW_TRY(check_ws(m, (int)(B * D), workspace, workspace_bytes, &ws));
  QWorkspace qws;
  GWW_TRY(qcheck_ws(qf, B, q_workspace, q_workspace_bytes, &qws));
  return forward_qscan_strided(m, qf, strain, B, D, (long)D * 2048, 2048, use_last_token, out, ws, qws,
                               (cudaStream_t)stream);
}

extern "C" int gww_stream_search_qscan(const gww_model_t* m, const gww_qfront_t* qf, const float* strain, int D,
                                       long n_samples, int hop, long first_window, long n_windows, int batch,
                                       float thr, float* scores, long* trig_idx, float* trig_score,
                                       int* trig_count, int capacity, void* workspace, size_t workspace_bytes,
                                       void* q_workspace, size_t q_workspace_bytes, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || !scores || D < 1 || hop < 1 || n_windows < 0 || first_window < 0 || batch < 1)
    return fail(GWW_ERR_INVALID, "stream_search_qscan: bad argument");
  if (n_windows == 0) return GWW_OK;
  if (n_samples < 2048 || (first_window + n_windows - 1) * hop + 2048 > n_samples)
    return fail(GWW_ERR_INVALID, "stream_search_qscan: windows exceed the segment");
  Workspace ws;
  GWW_TRY(check_ws(m, batch * D, workspace, workspace_bytes, &ws));
  QWorkspace qws;
  GWW_TRY(qcheck_ws(qf, batch, q_workspace, q_workspace_bytes, &qws));
  if (!m->has_head) return fail(GWW_ERR_INVALID, "stream_search_qscan: no head set");
  cudaStream_t s = (cudaStream_t)stream;
  const int C = m->head.dims[m->head.n_layers];
  if (C > 64) return fail(GWW_ERR_INVALID, "stream_search_qscan: head has %d outputs (> 64)", C);
  for (long k0 = 0; k0 < n_windows; k0 += batch) {
    const int nb = (int)((n_windows - k0 < batch) ? n_windows - k0 : batch);
    const float* base = strain + (first_window + k0) * hop;
    GWW_TRY(forward_qscan_strided(m, qf, base, nb, D, hop, n_samples, 1, ws.head_out, ws, qws, s));
    take_col0_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ws.head_out, C, nb, scores + k0);
    LAUNCH_CHECK();
    if (trig_idx && trig_score && trig_count)
      GWW_TRY(gww_threshold_compact(ws.head_out, C, nb, thr, first_window + k0, trig_idx, trig_score,
                                    trig_count, capacity, stream));
  }
  return GWW_OK;
}

// ------------------------------------------------------------------------------------------------
// whitening (MLGWSC-1/inference.py:56-137; kernels in whiten.cuh)
// ------------------------------------------------------------------------------------------------
struct WhitenPlan {
  long n, n_seg, first, nk;
  int seg_len, log2n, seg_stride, nb, L, H, cs_chunks, wrap;
  // workspace
  double2* tw; double *seg_psd, *psd0, *inv_asd, *mag, *partial, *q, *qt, *w;
  size_t total;
};
static int whiten_plan(long n, int seg_len, int seg_stride, int max_filter_len, int fir_half, uint8_t* base, WhitenPlan* pl) {
  if (n < 2 || (n & 1)) return fail(GWW_ERR_INVALID, "whiten: the number of samples must be even (got %ld)", n);
  int log2n = 0;
  while ((1 << log2n) < seg_len) ++log2n;
  if ((1 << log2n) != seg_len || seg_len < 16 || seg_len > kWhMaxSegLen)
    return fail(GWW_ERR_INVALID, "whiten: Welch segment length %d must be a power of two in [16, %d]", seg_len, kWhMaxSegLen);
  if (seg_stride < 1 || seg_stride > seg_len) return fail(GWW_ERR_INVALID, "whiten: bad segment stride %d", seg_stride);
  if (max_filter_len < 2 || (max_filter_len & 1) || max_filter_len > 4096 || max_filter_len > n)
    return fail(GWW_ERR_INVALID, "whiten: max_filter_len=%d must be even, <= 4096 and <= n", max_filter_len);
  // pycbc.psd.welch segmentation
  long n_seg = n / seg_stride;
  if ((n_seg - 1) * seg_stride + seg_len > n) n_seg -= 1;
  while (n_seg >= 1 && (n_seg - 1) * seg_stride + seg_len > n) n_seg -= 1;
  if (n_seg < 1) return fail(GWW_ERR_INVALID, "whiten: %ld samples are too few for one Welch segment of %d", n, seg_len);
  const long data_len = (n_seg - 1) * seg_stride + seg_len;
  const long diff = n - data_len;
  long first = diff / 2;
  if (diff % 2) first += 1;
  pl->n = n; pl->n_seg = n_seg; pl->first = first; pl->nk = n / 2 + 1;
  pl->seg_len = seg_len; pl->log2n = log2n; pl->seg_stride = seg_stride; pl->nb = seg_len / 2 + 1;
  pl->L = max_filter_len;
  long H = fir_half > 0 ? fir_half : 8192;
  if (H < max_filter_len / 2) H = max_filter_len / 2;
  if (H > 11264) H = 11264;                  // shared-memory tile of the FIR: (2048 + 2H) * 9/8 doubles <= 227 KB
  // short segment: the filter covers the whole circle and the result is exact; the taps -n/2 and +n/2 are the
  // same sample, so w[n/2] is halved (pl->wrap)
  pl->wrap = 0;
  if (H >= n / 2) { H = n / 2; pl->wrap = 1; }
  pl->H = (int)H;
  pl->cs_chunks = (int)((pl->nk + kCsChunk - 1) / kCsChunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return base + o; };
  pl->tw = (double2*)take((size_t)seg_len / 2 * sizeof(double2));
  pl->seg_psd = (double*)take((size_t)n_seg * pl->nb * 8);
  pl->psd0 = (double*)take((size_t)pl->nb * 8);
  pl->inv_asd = (double*)take((size_t)pl->nk * 8);
  pl->mag = (double*)take((size_t)pl->nk * 8);
  const int max_out = (pl->H + 1 > pl->L / 2 + 1) ? pl->H + 1 : pl->L / 2 + 1;
  pl->partial = (double*)take((size_t)pl->cs_chunks * max_out * 8);
  pl->q = (double*)take((size_t)(pl->L / 2 + 1) * 8);
  pl->qt = (double*)take((size_t)pl->L * 8);
  pl->w = (double*)take((size_t)(pl->H + 1) * 8);
  pl->total = off;
  return GWW_OK;
}

extern "C" size_t gww_whiten_workspace_bytes(long n, int seg_len, int seg_stride, int max_filter_len, int fir_half) {
  WhitenPlan pl;
  if (whiten_plan(n, seg_len, seg_stride, max_filter_len, fir_half, nullptr, &pl) != GWW_OK) return 0;
  return pl.total + 1024;
}

static double whiten_median_bias(long n) {   // pycbc.psd.estimate.median_bias
  if (n >= 1000) return std::log(2.0);
  double ans = 1.0;
  for (long i = 1; i < (long)((n - 1) / 2 + 1); ++i) ans += 1.0 / (2 * i + 1) - 1.0 / (2 * i);
  return ans;
}

extern "C" int gww_whiten(const double* strain, long n, double delta_t, int seg_len, int seg_stride,
                          int max_filter_len, double low_frequency_cutoff, int trunc_hann, int remove_corrupted,
                          int fir_half, double* white, float* white_f32, double* psd_out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!strain || (!white && !white_f32) || !(delta_t > 0)) return fail(GWW_ERR_INVALID, "whiten: bad argument");
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  WhitenPlan pl;
  GWW_TRY(whiten_plan(n, seg_len, seg_stride, max_filter_len, fir_half, base, &pl));
  if (workspace == nullptr || workspace_bytes < pl.total + (size_t)(base - reinterpret_cast<uint8_t*>(workspace)))
    return fail(GWW_ERR_WORKSPACE, "whiten: workspace too small (need %zu bytes)", pl.total + 1024);
  cudaStream_t s = (cudaStream_t)stream;
  ProfScope ps(PK_WHITEN, s);
  // 1. Welch PSD
  whiten_twiddle_kernel<<<(seg_len / 2 + 255) / 256, 256, 0, s>>>(pl.tw, seg_len);
  LAUNCH_CHECK();
  GWW_TRY(ensure_smem_attr(welch_segments_kernel, 2 * kWhMaxSegLen * 16));
  {
    const long grid = pl.n_seg < 8L * g_num_sms ? pl.n_seg : 8L * g_num_sms;
    welch_segments_kernel<<<(unsigned)grid, 256, (size_t)2 * seg_len * 16, s>>>(strain, pl.first, seg_len, pl.log2n, seg_stride,
                                                                             pl.n_seg, delta_t, pl.tw, pl.seg_psd);
    LAUNCH_CHECK();
  }
  {
    double sumw2 = 0.0;
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < seg_len; ++i) {
      const double w = 0.5 - 0.5 * std::cos(2.0 * PI * i / (seg_len - 1));
      sumw2 += w * w;
    }
    const double delta_f = 1.0 / delta_t / seg_len;
    const double scale = 2.0 * delta_f * seg_len / sumw2;
    welch_median_kernel<<<pl.nb, 256, 0, s>>>(pl.seg_psd, pl.n_seg, pl.nb, 1.0 / whiten_median_bias(pl.n_seg), scale, pl.psd0);
    LAUNCH_CHECK();
    if (psd_out) CU_TRY(cudaMemcpyAsync(psd_out, pl.psd0, (size_t)pl.nb * 8, cudaMemcpyDeviceToDevice, s));
  }
  // 2. interpolate + inverse ASD
  const double psd_df = 1.0 / delta_t / seg_len;
  const double df = 1.0 / ((double)n * delta_t);
  long kmin = 1;
  if (low_frequency_cutoff > 0.0) kmin = (long)(low_frequency_cutoff / df);
  inv_asd_kernel<<<(unsigned)((pl.nk + 255) / 256), 256, 0, s>>>(pl.psd0, pl.nb, psd_df, df, pl.nk, kmin, pl.inv_asd);
  LAUNCH_CHECK();
  // 3. q = irfft(inv_asd) at taps 0..L/2, truncation window
  const int nq = pl.L / 2 + 1;
  cosine_series_kernel<<<dim3(pl.cs_chunks, (nq + kCsThreads - 1) / kCsThreads), kCsThreads, 0, s>>>(pl.inv_asd, pl.nk, n, nq, pl.partial);
  LAUNCH_CHECK();
  cosine_series_reduce_kernel<<<(nq + 255) / 256, 256, 0, s>>>(pl.partial, pl.cs_chunks, nq, 1.0 / (double)n, pl.q, -1);
  LAUNCH_CHECK();
  trunc_window_kernel<<<(pl.L + 255) / 256, 256, 0, s>>>(pl.q, pl.L, trunc_hann, pl.qt);
  LAUNCH_CHECK();
  // 4. |Q|
  GWW_TRY(ensure_smem_attr(filter_mag_kernel, 4096 * 8));
  filter_mag_kernel<<<(unsigned)((pl.nk + 255) / 256), 256, (size_t)pl.L * 8, s>>>(pl.qt, pl.L, pl.nk, n, pl.mag);
  LAUNCH_CHECK();
  // 5. w = irfft(|Q|) at taps 0..H, circular FIR, crop
  const int nw = pl.H + 1;
  cosine_series_kernel<<<dim3(pl.cs_chunks, (nw + kCsThreads - 1) / kCsThreads), kCsThreads, 0, s>>>(pl.mag, pl.nk, n, nw, pl.partial);
  LAUNCH_CHECK();
  cosine_series_reduce_kernel<<<(nw + 255) / 256, 256, 0, s>>>(pl.partial, pl.cs_chunks, nw, 1.0 / (double)n, pl.w,
                                                                pl.wrap ? nw - 1 : -1);
  LAUNCH_CHECK();
  const long n0 = remove_corrupted ? pl.L / 2 : 0;
  const long n_out = remove_corrupted ? n - pl.L : n;
  if (n_out <= 0) return fail(GWW_ERR_INVALID, "whiten: nothing left after removing the corrupted edges");
  const size_t fir_smem = (size_t)(kFirTile + 2 * pl.H + (kFirTile + 2 * pl.H) / 8 + 8) * 8;
  GWW_TRY(ensure_smem_attr(fir_apply_kernel, (int)fir_smem));
  fir_apply_kernel<<<(unsigned)((n_out + kFirTile - 1) / kFirTile), kFirThreads, fir_smem, s>>>(strain, n, pl.w, pl.H, n0, n_out,
                                                                                             white, white_f32);
  LAUNCH_CHECK();
  return GWW_OK;
}

// ---- building blocks -------------------------------------------------------------------------------
extern "C" int gww_gemm_bf16(const void* A, const void* W, void* C, const float* bias, const float* resid,
                             const float* pos, long M, int N, int K, int epilogue, int block_n,
                             void* stream) {
  GWW_TRY(gww_device_ok());
  if (!A || !W || !C || M <= 0) return fail(GWW_ERR_INVALID, "gemm: bad argument");
  if (epilogue == EPI_BIAS_GELU_POS_F32) {
    // positional table indexed by the row inside a batch entry: expose as batch=1, rows=M
    GemmCall g{};
    g.a_base = A;
    g.a_dims[0] = K; g.a_dims[1] = 1; g.a_dims[2] = M; g.a_dims[3] = 1;
    g.a_strides[0] = (uint64_t)K * 2; g.a_strides[1] = (uint64_t)K * 2; g.a_strides[2] = (uint64_t)M * K * 2;
    g.w_base = W; g.ktot = K; g.c_base = C;
    g.c_strides[0] = (uint64_t)N * 4; g.c_strides[1] = (uint64_t)M * N * 4;
    g.p.rows = (int)M; g.p.batch = 1; g.p.n = N; g.p.kb_per_tap = K / 64; g.p.taps = 1; g.p.p_mod = 1;
    g.p.bias = bias; g.p.resid = nullptr; g.p.pos = pos;
    g.epi = epilogue; g.block_n = block_n ? block_n : pick_block_n(N);
    return run_gemm(g, (cudaStream_t)stream);
  }
  return run_linear(A, W, C, bias, resid, M, N, K, epilogue, block_n ? block_n : pick_block_n(N),
                    (cudaStream_t)stream);
}

extern "C" int gww_attention(const void* qkv, void* out, long n, int T, int d_model, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!qkv || !out || n <= 0) return fail(GWW_ERR_INVALID, "attention: bad argument");
  return run_attention(qkv, out, n, T, d_model, (cudaStream_t)stream);
}

extern "C" int gww_layernorm(const float* x, void* out, const float* gamma, const float* beta, long rows,
                             int d, int out_bf16, void* stream) {
  GWW_TRY(gww_device_ok());
  if (!x || !out || rows <= 0) return fail(GWW_ERR_INVALID, "layernorm: bad argument");
  if (out_bf16) return run_ln_t<op16_t>(x, (op16_t*)out, gamma, beta, rows, d, 0, 1, (cudaStream_t)stream);
  return run_ln_t<float>(x, (float*)out, gamma, beta, rows, d, 0, 1, (cudaStream_t)stream);
}
