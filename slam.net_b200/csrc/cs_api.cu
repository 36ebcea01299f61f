// cs_api.cu — the C ABI of include/coreslam_b200.h over the kernels in cs_kernels.cuh.
//
// One cs_processor replaces one CoreSLAMProcessor (CoreSLAM/CoreSLAMProcessor.cs): it owns the
// device-resident HoleMap, the pinned staging block the scan is written into, one CUDA stream and a
// mapped result slot the finalize kernel stores the pose into.  No CPU fallback exists: every compute
// entry point launches kernels or fails.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost a pointer test unless a profiler is attached

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <limits>
#include <cstring>
#include <atomic>
#include <exception>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/coreslam_b200.h"
#include "cs_kernels.cuh"
#include "cs_obstacle.cuh"
#include "cs_wedge.cuh"

static_assert(sizeof(CsDevResult) == sizeof(cs_result), "cs_result layout");
static_assert(sizeof(cs_config) == 72, "cs_config layout (ctypes / P/Invoke mirror it)");
static_assert(sizeof(CsStepHeader) == 48, "CsStepHeader");

namespace {

thread_local std::string g_create_error;

constexpr size_t kHdrBytes = 64;
// NVTX range over one C-ABI call: shows up as a named span in Nsight Systems / ncu --nvtx
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

constexpr int kRayCopies = 8;  // copies of the per-ray draw parameters of a stand-alone session (see CsSession::rays)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Timing {
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool valid = false;
};

}  // namespace

// Production mode (on-device Philox candidates): the candidates of scan k+1 are a function of (seed, k+1) alone, so their sort
// is queued right behind the draw kernel of scan k and runs in its shadow; the search of scan k+1 then starts without a sort
// in front.  The same holds for the replay of a device-resident scan log with candidate tables (cs_replay): the next scan's
// table is already in HBM.  (Not for cs_update with a host table: it arrives with the scan.)  The descriptor says what was
// sorted into which half of CsSession::s2_sorted; a step that asks for anything else (another candidate mode, table, slice or
// scan number, another upload of the log) sorts as usual and overwrites it.
struct Presort {
  bool valid = false;
  unsigned scan_index = 0;
  int cand_first = 0, cand_count = 0, slot = 0;
  int cand_mode = 0;
  const float* cand = nullptr;  // replay of a device-resident log with candidate tables: the table that was sorted ...
  uint64_t tag = 0;             // ... and the upload of the log it belongs to
};

struct cs_processor {
  cs_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool tiled = true;
  int size = 0, pitch_tiles = 0, max_points = 0, n_cand = 0;
  float scale = 0.f;
  size_t map_cells = 0;  // allocated cells (tiled: pitch_tiles^2 * 64)

  CsSession hs{};            // host mirror of the session descriptor
  CsSession* d_sess = nullptr;
  uint16_t* d_map = nullptr;
  uint16_t* d_linear = nullptr;  // lazily allocated row-major scratch for upload/download
  uint8_t* d_packed = nullptr;
  int4* d_rays = nullptr;       // kRayCopies copies of ray_stride entries
  int* d_batch_max = nullptr;   // kRayCopies copies of batch_stride entries
  unsigned long long* d_prep_words = nullptr;
  int ray_stride = 0, batch_stride = 0;
  int* d_ray_dbg = nullptr;
  int* d_distances = nullptr;
  // wedge integration scratch (CsSession::w_*)
  int2* d_w_rk = nullptr;
  float2* d_w_bkey = nullptr;
  int* d_w_top = nullptr;
  unsigned long long* d_spec = nullptr;  // glue table (CsStepArgs::spec): 8 words per flat candidate
  int w_slot = 0;
  long long* d_ring_cycles = nullptr;
  unsigned long long* d_checksum = nullptr;

  // slab search scratch (CsSession::s2_*), allocated when the handle has enough candidates for it to pay
  float4* d_s2_sorted = nullptr;
  float4* d_s2_tmp = nullptr;
  unsigned long long* d_s2_meta = nullptr;
  unsigned long long* d_s2_acc = nullptr;
  unsigned* d_s2_ghist = nullptr;
  int s2_cap = 0, s2_toggle = 0;
  Presort presort;

  // ObstacleMap (cfg.obstacle_map_size > 0): CoreSLAM/ObstacleMap.cs, CoreSLAMProcessor.cs:53, :132-133
  CsObstacle ho{};               // host mirror of the descriptor
  CsObstacle* d_obst = nullptr;
  int unmapped_obstacle_hits = -5;  // :98

  // asynchronous map export (cs_map_export_begin / _wait): snapshot on the main stream, D2H on a side stream
  cudaStream_t export_stream = nullptr;
  cudaEvent_t ev_export_snap = nullptr, ev_export_done = nullptr;
  bool export_busy = false;
  int8_t* d_obst_snap = nullptr;

  float2* d_cloud = nullptr;  // scan points computed on the device from raw segments (cs_update_segments)

  // staging: [hdr 64][points][cand][cand_cs]; segments path: [hdr 64][rays][seg_first][seg_poses][cand]
  size_t stage_bytes = 0;
  uint8_t* h_stage = nullptr;  // pinned
  uint8_t* d_stage = nullptr;
  // cs_update's staging runs ahead of the stream: the H2D copy of scan k goes through a copy stream into the buffer scan
  // k-1 is NOT using (d_stage / d_stage_b alternate), so it overlaps the integration of scan k-1 instead of queueing
  // behind it; the main stream only waits for the copy's event.  ev_stage_free[i]: last step that reads buffer i is enqueued.
  uint8_t* d_stage_b = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_stage_free[2] = {nullptr, nullptr};
  int stage_flip = 0, stage_cur = 0;

  // mapped result slot
  uint8_t* h_slot = nullptr;  // pinned+mapped: CsDevResult at 0, seq flag at 64, stuck-poll word (CsSpin) at 96
  uint8_t* d_slot = nullptr;
  unsigned seq = 0;

  // host mirrors of the deterministic parts of the state machine
  int parity = 0;      // which CsSession::state / key slot is current
  int scan_count = 0;  // CoreSLAMProcessor.cs:34 (saturates at PositionSearchBeginning)
  int search_begin = 5;
  unsigned update_count = 0;  // Philox scan index

  // split-phase update in flight (cs_update_begin ... cs_update_finish)
  bool pending = false;
  CsStepArgs pending_args{};
  int pending_points = 0, pending_rings = 0;

  // candidate-split group (cs_group_*): this rank's exchange table and the (peer-mapped) tables of all ranks
  int group_rank = 0, group_world = 0;
  bool group_shares_device = false;  // another rank of the group runs on this handle's device (same process)
  unsigned xchg_seq = 0;
  unsigned long long* d_xchg = nullptr;
  unsigned long long* xchg_peer[CS_GROUP_MAX] = {};
  bool xchg_ipc[CS_GROUP_MAX] = {};  // opened with cudaIpcOpenMemHandle (to be closed on detach)

  uint64_t launches = 0;
  unsigned step_counter = 0;
  Timing tm;
  cs_timing last_timing{};
  std::string error;
  bool poisoned = false;
};

struct cs_scanlog {
  int device = 0, n_scans = 0, max_points = 0, n_offsets = 0;
  std::vector<CsStepHeader> h_hdr;
  std::vector<float> h_points;   // n_scans * max_points * 2
  std::vector<float> h_offsets;  // n_scans * n_offsets * 3
  std::vector<double> h_max_range;  // per scan, sizes the update kernel's grid
  CsStepHeader* d_hdr = nullptr;
  float2* d_points = nullptr;
  float* d_offsets = nullptr;
  CsDevResult* d_results = nullptr;
  bool uploaded = false;
  uint64_t generation = 0;  // counts the uploads: what a handle sorted ahead from this log is only good for the same upload
};

namespace {

cs_status fail(cs_processor* h, cs_status code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) {
    h->error = buf;
    if (code == CS_ERR_CUDA) h->poisoned = true;
  } else {
    g_create_error = buf;
  }
  return code;
}

#define CS_CUDA(h, expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return fail((h), CS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// CS_FLAG_DEBUG_BOUNDED_SPIN: a device-side poll of an earlier step gave up (CsSpin, cs_kernels.cuh): the handle fails
const char* stuck_site_name(unsigned w) {
  switch (w & 0xffu) {
    case CS_STUCK_POSE: return "the pose of the step (search kernel / set-up kernel never published it)";
    case CS_STUCK_RAYS: return "the rays of the step (a preparing block never arrived)";
    case CS_STUCK_HANDOFF: return "the hand-off word of a contested cell";
    default: return "an unknown word";
  }
}
#define CS_CHECK_STUCK(h, slot, failfn)                                                                         \
  do {                                                                                                          \
    const unsigned _w = (slot) ? *reinterpret_cast<volatile unsigned*>(slot) : 0u;                              \
    if (_w)                                                                                                     \
      return failfn((h), CS_ERR_CUDA, "bounded polls: block %u of the draw kernel gave up waiting for %s", _w >> 8, \
                    stuck_site_name(_w));                                                                       \
  } while (0)

#define CS_CHECK_HANDLE(h)                                                          \
  do {                                                                              \
    if (!(h)) return CS_ERR_INVALID_ARGUMENT;                                       \
    if ((h)->poisoned) return CS_ERR_CUDA;                                          \
    CS_CHECK_STUCK(h, (h)->h_slot ? (h)->h_slot + 96 : nullptr, fail);              \
    cudaError_t _e = cudaSetDevice((h)->device);                                    \
    if (_e != cudaSuccess) return fail((h), CS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); \
  } while (0)

cs_status push_session(cs_processor* h) {
  CS_CUDA(h, cudaMemcpyAsync(h->d_sess, &h->hs, sizeof(CsSession), cudaMemcpyHostToDevice, h->stream));
  return CS_OK;
}

// session fields the host owns; state/key/x1.. are device-owned, so properties are patched individually
cs_status patch_session(cs_processor* h, size_t offset, const void* src, size_t bytes) {
  CS_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->d_sess) + offset, src, bytes, cudaMemcpyHostToDevice,
                             h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));  // src may be a stack temporary
  return CS_OK;
}

template <typename F>
void dispatch_layout(bool tiled, F&& f) {
  if (tiled) f(std::true_type{});
  else f(std::false_type{});
}

cs_status launch_fill(cs_processor* h, uint16_t value) {
  cs_fill_kernel<<<148 * 4, 256, 0, h->stream>>>(h->d_map, h->map_cells, value);
  h->launches++;
  CS_CUDA(h, cudaGetLastError());
  return CS_OK;
}

struct StagePlan {
  size_t off_points, off_cand, off_cs, total;
};

StagePlan plan_stage(int n_points, int n_cand_floats3, bool with_cs) {
  StagePlan p;
  p.off_points = kHdrBytes;
  p.off_cand = align_up(p.off_points + (size_t)n_points * 8, 16);
  p.off_cs = align_up(p.off_cand + (size_t)n_cand_floats3 * 12, 16);
  p.total = align_up(p.off_cs + (with_cs ? (size_t)(n_cand_floats3 + 1) * 8 : 0), 16);
  return p;
}

// Upper bound on the ring count of a scan: |rotated, scaled point| + half the hole width, plus rounding slack.
// The 512-thread instances of the rings kernel use more dynamic shared memory than the default limit.
cudaError_t rings_allow_shared_memory() {
  const int one = (int)CS_RING_SMEM(CS_RING_MAX_THREADS, CS_RING_MAX_SLOT_BITS, 0);
  const int two = (int)CS_RING_SMEM(CS_RING_MAX_THREADS, CS_RING_MAX_SLOT_BITS, 1);
  cudaError_t e = cudaFuncSetAttribute(cs_rings_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, one);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_rings_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, one);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_rings_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_rings_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two);
  return e;
}

// Wedge integration (cs_wedge.cuh): levels a ray of a size x size map can reach, the words of one half of CsSession::w_top,
// and the dynamic shared memory of the kernel (one 16-bit count per level and key sector).
int wedge_levels(int size) {
  const int k = size > 1 ? size - 1 : 1;
  int L = 0;
  if (k < 64) { while ((2 << L) <= k) L++; } else L = 5 + (k >> 6);
  return L + 1;
}
size_t wedge_top_words(int size) { return (size_t)wedge_levels(size) * (CS_W_SECTORS + 1); }
size_t wedge_smem(int) { return 0; }
cudaError_t wedge_allow_shared_memory();

int rings_hint_of(int size, float scale, float hole_width, double max_range) {
  if (!(max_range == max_range) || max_range > 1e30) return size;
  double cells = max_range * (double)scale * 1.00001 + 0.5 * (double)hole_width * (double)scale + 4.0;
  if (cells >= (double)size) return size;
  return (int)cells + 1;
}
int rings_hint(const cs_processor* h, double max_range) { return rings_hint_of(h->size, h->scale, h->hs.hole_width, max_range); }

// Largest range of a scan (for the rings hint: an upper bound on the rings the scan can reach; NaN sticks, overflow gives
// +inf, both mean "all rings").  Four independent running maxima: a batch update calls this for every session on one host
// thread (1024 x 360 points per cfg5 step), where a single dependent chain cost as much as the staging copy itself.
double max_range_of(const float* points, int n) {
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  bool nan_seen = false;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    const float* p = points + 2 * (size_t)i;
    const float r0 = p[0] * p[0] + p[1] * p[1], r1 = p[2] * p[2] + p[3] * p[3];
    const float r2 = p[4] * p[4] + p[5] * p[5], r3 = p[6] * p[6] + p[7] * p[7];
    nan_seen |= (r0 != r0) | (r1 != r1) | (r2 != r2) | (r3 != r3);
    m0 = r0 > m0 ? r0 : m0;
    m1 = r1 > m1 ? r1 : m1;
    m2 = r2 > m2 ? r2 : m2;
    m3 = r3 > m3 ? r3 : m3;
  }
  for (; i < n; i++) {
    const float r = points[2 * i] * points[2 * i] + points[2 * i + 1] * points[2 * i + 1];
    nan_seen |= (r != r);
    m0 = r > m0 ? r : m0;
  }
  if (nan_seen) return std::numeric_limits<double>::quiet_NaN();
  const float m = std::fmax(std::fmax(m0, m1), std::fmax(m2, m3));
  return std::sqrt((double)m) * 1.000001;  // float rounding of r^2 (2^-24 relative) stays inside the bound
}

// What a step is launched on: one processor (n_sessions = 1) or a batch of independent sessions
// (grid.y = n_sessions; every kernel indexes its session by blockIdx.y).
// Experiment knobs (environment, read once): CS_TUNE_SEARCH_WARPS (2/4/8), CS_TUNE_RING_SPAN, CS_TUNE_RING_THREADS.
// Unset = the built-in choice.  They change launch shapes only, never results.
struct Tune {
  int search_warps = 0, ring_span = 0, ring_threads = 0, ring_slot_bits = 0, ring_blocks_per_sm = 0, ring_small = 0;
  int search2 = 0, s2_points = 0, s2_threads = 0, s2_min_cand = 0, s2_sort_one_block = 0, copy_stream = 0;
  int integrate = 0, w_general = 0, w_blocks = 0, w_prefetch = 0, w_sub = 0, w_prev = 0, w_carveout = 0, s2_carveout = 0;
  int spin_ms = 0, fault = 0, w_resident = 0, spec = 0, presort = 0;
  Tune() {
    auto geti = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; };
    search_warps = geti("CS_TUNE_SEARCH_WARPS");
    search2 = geti("CS_TUNE_SEARCH2");          // -1: never use the slab search, 0: built-in rule
    s2_points = geti("CS_TUNE_S2_POINTS");      // points per cluster
    s2_threads = geti("CS_TUNE_S2_THREADS");    // candidates per slab
    s2_sort_one_block = geti("CS_TUNE_S2_SORT_ONE_BLOCK");  // 1: sort generated candidates with one block whenever they fit
    copy_stream = geti("CS_TUNE_COPY_STREAM");  // -1: cs_update stages its inputs on the main stream
    s2_min_cand = geti("CS_TUNE_S2_MIN_CAND");  // fewest candidates the slab search is used for
    integrate = geti("CS_TUNE_INTEGRATE");      // 1: the rings kernel draws the scan instead of the wedge kernel (A/B runs)
    w_general = geti("CS_TUNE_W_GENERAL");      // 1: every task of the wedge kernel takes its general path (tests)
    w_blocks = geti("CS_TUNE_W_BLOCKS");        // blocks of the wedge kernel per session
    w_prefetch = geti("CS_TUNE_W_PREFETCH");    // -1: no L2 prefetch of the map around the pose
    w_carveout = geti("CS_TUNE_W_CARVEOUT");    // shared-memory carve-out (percent) of the wedge kernel
    s2_carveout = geti("CS_TUNE_S2_CARVEOUT");  // ... of the slab-search and sort kernels
    w_prev = geti("CS_TUNE_W_PREV");            // -1: every scan builds its task table from its own counts
    w_resident = geti("CS_TUNE_W_RESIDENT");    // 4: one small scan alone also runs the four-blocks-per-SM instance of the wedge kernel
    presort = geti("CS_TUNE_PRESORT");          // -1: production mode sorts a scan's candidates in front of its search, never ahead
    spec = geti("CS_TUNE_SPEC");                // -1: no glue table (the publishing thread always computes the glue itself)
    w_sub = geti("CS_TUNE_W_SUB");              // most warps a task's rings are split over (1, 2, 4, 8)
    spin_ms = geti("CS_TUNE_SPIN_MS");          // CS_FLAG_DEBUG_BOUNDED_SPIN: milliseconds a device-side poll lasts (default 2000)
    fault = geti("CS_TUNE_FAULT");              // tests of the bounded polls: 1 = the draw kernel waits for a pose tag nobody publishes
    ring_span = geti("CS_TUNE_RING_SPAN");
    ring_threads = geti("CS_TUNE_RING_THREADS");
    ring_slot_bits = geti("CS_TUNE_RING_SLOT_BITS");
    ring_small = geti("CS_TUNE_RING_SMALL");  // -1: batches use the 512-thread instance of the rings kernel too
    ring_blocks_per_sm = geti("CS_TUNE_RING_BLOCKS_PER_SM");  // resident blocks per SM the one-session rings grid is capped at
  }
};
const Tune& tune() { static Tune t; return t; }

cudaError_t wedge_allow_shared_memory() {
  // Experiment knobs: the shared-memory carve-out the draw and search kernels ask for (percent of the SM's 228 KB).  A
  // kernel whose carve-out differs from its predecessor's cannot share an SM with it.
  cudaError_t e = cudaSuccess;
  if (tune().w_carveout > 0) {
    e = cudaFuncSetAttribute(cs_wedge_kernel<true, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().w_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_wedge_kernel<false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().w_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_wedge_kernel<true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().w_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_wedge_kernel<false, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().w_carveout);
  }
  if (e == cudaSuccess && tune().s2_carveout > 0) {
    e = cudaFuncSetAttribute(cs_search2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().s2_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_search2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().s2_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_sort_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().s2_carveout);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cs_sort_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, tune().s2_carveout);
  }
  return e;
}


int device_sm_count(int device) {
  static int cache[64] = {0};
  if (device < 0 || device >= 64) return 148;
  if (!cache[device]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    cache[device] = n;
  }
  return cache[device];
}

// Warps (= candidates) per block of the search kernel.  All blocks of one scan are resident at once, and a
// warp's run time is fixed by the scan, so the kernel ends when the SM with the most warps ends: pick the
// block size whose busiest SM carries the fewest warps (4097 candidates on 148 SMs: 8 warps -> 4 blocks = 32
// warps on some SMs against 27.7 on average; 4 warps -> 7 blocks = 28).  Bigger blocks win ties (the scan is
// staged once per block).
int cs_search_warps(long long cand_count, int n_sessions, int num_sms) {
  if (tune().search_warps == 2 || tune().search_warps == 4 || tune().search_warps == 8) return tune().search_warps;
  const long long total = cand_count * (long long)n_sessions;
  if (total > (long long)num_sms * 40) return CS_SEARCH_WARPS;  // more than one wave anyway
  int best = CS_SEARCH_WARPS;
  long long best_cost = -1;
  for (int w = CS_SEARCH_WARPS; w >= 2; w >>= 1) {
    const long long blocks = ((cand_count + w - 1) / w) * n_sessions;
    const long long per_sm = (blocks + num_sms - 1) / num_sms;
    if (per_sm > 16) continue;  // keep the whole grid resident (shared memory: 16 x 16 KB)
    const long long cost = per_sm * w;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = w; }
  }
  return best;
}

// Slab search (cs_sort_kernel + cs_search2_kernel) instead of the warp-per-candidate kernel: one session alone whose
// candidate count fills the machine with slabs.  Block (c, s) = points [c*points, ...) x sorted candidates [s*threads, ...).
constexpr int kS2MinCand = 1024;
constexpr int kS2MinCandBatch = 256;
struct S2Plan {
  int points = 0, threads = 0, clusters = 0, slabs = 0;
};
int cs_s2_min_cand(uint32_t flags, int n_sessions = 1) {  // fewest candidates (searchPose included) the slab search is used for; 0 = never
  if ((flags & CS_FLAG_SEARCH_WARP) || tune().search2 < 0) return 0;
  if (flags & CS_FLAG_SEARCH_SLAB) return 1;
  if (tune().s2_min_cand > 0) return tune().s2_min_cand;
  // a batch fills the machine with its sessions; what the slabs need is enough candidates to amortise a block's set-up
  return n_sessions > 1 ? kS2MinCandBatch : kS2MinCand;
}
bool cs_plan_search2(int n_sessions, int s2_cap, int min_cand, long long cand_count, int n_points, int num_sms, S2Plan* p,
                     int max_threads = CS_S2_MAX_THREADS) {
  if (min_cand <= 0 || n_sessions < 1 || s2_cap <= 0 || cand_count > s2_cap || n_points < 1) return false;
  if (cand_count < min_cand) return false;
  if (n_sessions > 1 && (cand_count > CS_SORT_THREADS * CS_SORT_REG || n_sessions > 65535)) return false;  // batches: one sort block per session
  // the plan depends on (candidates, points, SMs) only: remember the last one (a replay asks for the same every scan)
  struct Memo { long long cand = -1; int points = -1, sms = -1, sessions = -1, max_threads = -1; bool ok = false; S2Plan plan; };
  static thread_local Memo memo;
  if (memo.cand == cand_count && memo.points == n_points && memo.sms == num_sms && memo.sessions == n_sessions && memo.max_threads == max_threads) {
    *p = memo.plan;
    return memo.ok;
  }
  memo.cand = cand_count; memo.points = n_points; memo.sms = num_sms; memo.sessions = n_sessions; memo.max_threads = max_threads; memo.ok = false;
  // Launch shape: the kernel is one resident wave of blocks (clusters x slabs) and ends when the busiest SM ends.  A lane
  // pays a fixed set-up (candidate pose, cos/sin: ~kSetup instruction slots) plus ~kLookup per point of its cluster, in
  // whole batches of CS_S2_BATCH; a block's cost is that times its warps, and an SM issues about kIpcPerWarp
  // instructions per clock and resident warp, up to kIpcMax.  Pick the (points, threads) pair with the shortest busiest SM.
  constexpr double kSetup = 450.0, kLookup = 26.0, kIpcPerWarp = 0.12, kIpcMax = 3.0;
  int best_points = 0, best_threads = 0;
  double best_cost = 0.0;
  const int p_lo = tune().s2_points > 0 ? tune().s2_points : CS_S2_BATCH;
  const int p_hi = tune().s2_points > 0 ? tune().s2_points : CS_S2_MAX_POINTS;
  const int t_lo = tune().s2_threads > 0 ? (tune().s2_threads + 31) / 32 * 32 : 64;
  const int t_hi = tune().s2_threads > 0 ? (t_lo < max_threads ? t_lo : max_threads) : max_threads;
  for (int points = p_lo; points <= p_hi && points <= CS_S2_MAX_POINTS; points += 4) {
    const long long clusters = (n_points + points - 1) / points;
    if (clusters > CS_S2_MAX_CLUSTERS) continue;
    const int batches = ((int)((n_points + clusters - 1) / clusters) + CS_S2_BATCH - 1) / CS_S2_BATCH;  // of the fullest cluster
    for (int threads = t_hi; threads >= t_lo; threads -= 32) {  // bigger blocks win ties: one cluster's lines shared by more warps
      const long long slabs = (cand_count + threads - 1) / threads;
      if (slabs > 65535) continue;
      const long long blocks = clusters * slabs * n_sessions;
      const long long per_sm = (blocks + num_sms - 1) / num_sms;
      const double warps = (double)per_sm * (threads / 32);
      const double resident = warps < 48.0 ? warps : 48.0;
      const double ipc = resident * kIpcPerWarp < kIpcMax ? resident * kIpcPerWarp : kIpcMax;
      const double cost = warps * (kSetup + kLookup * CS_S2_BATCH * batches) / ipc;
      if (best_points == 0 || cost < best_cost * 0.999) { best_cost = cost; best_points = points; best_threads = threads; }
    }
  }
  if (best_points == 0) return false;
  const long long clusters = (n_points + best_points - 1) / best_points;
  const long long slabs = (cand_count + best_threads - 1) / best_threads;
  if (slabs > 65535 || clusters > CS_S2_MAX_CLUSTERS) return false;
  // even out: the same number of blocks with the smallest equal shares
  p->clusters = (int)clusters;
  p->points = (int)((n_points + clusters - 1) / clusters);
  p->slabs = (int)slabs;
  p->threads = (int)(((cand_count + slabs - 1) / slabs + 31) / 32 * 32);
  memo.plan = *p;
  memo.ok = true;
  return true;
}

struct LaunchCtx {
  int s2_cap = 0;            // capacity of the slab-search scratch (0: none)
  int s2_min_cand = 0;       // cs_s2_min_cand of the handle
  int* s2_toggle = nullptr;  // which half of CsSession::s2_sorted the next sort writes
  Presort* presort = nullptr;  // one session alone: the sort queued ahead for the next scan, if any
  const float* next_cand = nullptr;  // cs_replay with tables: the next scan's table (nullptr: none) and the log's upload number
  uint64_t log_tag = 0;
  const CsSession* hs = nullptr;  // host mirror of the session (host-owned constants for the slab kernels)
  int num_sms;
  cudaStream_t stream;
  CsSession* d_sess;
  bool tiled;
  int n_sessions;
  uint64_t* launches;
  unsigned* step_counter;  // source of CsStepArgs::step_id
  int* w_slot = nullptr;   // which half of CsSession::w_top the next drawn step counts into
  long long* diag;
  int diag_rings;
  volatile unsigned* stuck_dev = nullptr;  // CS_FLAG_DEBUG_BOUNDED_SPIN: device address of the mapped-host word (CsSpin)
  unsigned long long* spec = nullptr;      // glue table of the handle (one session alone)
  cudaEvent_t ev_pose;   // optional: recorded once the pose is out
  cudaEvent_t ev_done;   // optional: recorded after the rings kernel
};

enum { CS_PHASE_SEARCH = 1, CS_PHASE_FINISH = 2, CS_PHASE_ALL = 3 };

// One step on the stream: [search kernel, whose last block publishes the pose and prepares the rays] or
// [set-up kernel] -> (big scans: multi-block ray preparation) -> rings kernel.  Split phases (multi-GPU
// candidate split): SEARCH launches only the search over this GPU's candidate slice and leaves the packed
// arg-min in CsSession::key[parity]; FINISH — after the caller's 8-byte min exchange — launches the set-up
// kernel (which decodes the reduced key) and the rings kernel.
// Launch with the programmatic-stream-serialization attribute: the kernel may become resident while its
// predecessor in the stream drains; every kernel here calls cs_pdl_wait() before it reads anything the
// predecessor wrote.
// One-shot: the next launch_pdl on this thread is a plain (fully serialised) launch.  Set after a kernel that PRODUCES
// inputs of the step (cs_cloud_kernel writes the scan points): the step's first kernel reads its inputs before its
// dependency wait, which is only sound when they were complete before it could start.
thread_local bool g_next_launch_plain = false;

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  if (g_next_launch_plain) {
    g_next_launch_plain = false;
    kernel<<<grid, block, smem, stream>>>(args...);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// The candidate sort of a step (a.s2_* set): one block sorts a table of up to 8192 candidates out of its registers; generated
// (Philox) candidates are spread over the SMs, one per thread, as soon as there are more than one block's threads of them
// (batches: one block per session).  warm: one session alone with a tiled map — the sort grid's other blocks (or the
// histogram kernel's threads) pull the map within reach of the step into L2 beside the sort.
cudaError_t launch_sort(const LaunchCtx& c, CsStepArgs& a, bool warm) {
  cudaError_t e = cudaSuccess;
  warm = warm && c.n_sessions == 1 && c.tiled && tune().w_prefetch >= 0 && tune().w_prefetch != 2;
  const bool one_block = a.cand_count <= CS_SORT_THREADS * CS_SORT_REG &&
                         (a.cand_mode != CS_CAND_PHILOX || a.cand_count <= CS_SORT_THREADS || c.n_sessions > 1 || tune().s2_sort_one_block > 0);
  if (warm) a.w_prefetch = 1;
  if (one_block) {
    e = launch_pdl(a.cand_mode == CS_CAND_PHILOX ? cs_sort_kernel<true> : cs_sort_kernel<false>,
                   dim3(warm ? 1u + (unsigned)c.num_sms : 1u, (unsigned)c.n_sessions), dim3(CS_SORT_THREADS), 0, c.stream, c.d_sess, a);
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  } else {
    const unsigned sort_blocks = (unsigned)((a.cand_count + CS_SORT_MB_CHUNK - 1) / CS_SORT_MB_CHUNK);
    e = launch_pdl(a.cand_mode == CS_CAND_PHILOX ? cs_sort_hist_kernel<true> : cs_sort_hist_kernel<false>, dim3(sort_blocks),
                   dim3(CS_SORT_THREADS), 0, c.stream, c.d_sess, a);
    if (e != cudaSuccess) return e;
    e = launch_pdl(cs_sort_scatter_kernel, dim3(sort_blocks), dim3(CS_SORT_THREADS), 0, c.stream, c.d_sess, a);
    if (e != cudaSuccess) return e;
    (*c.launches) += 2;
  }
  return e;
}

cudaError_t launch_step_ctx(const LaunchCtx& c, CsStepArgs a, int n_points, int rings, int phases) {
  // An empty cloud (Update with segments that carry no rays): CalculateDistance returns int.MaxValue for every pose (:251-258),
  // so searchPose wins (:630-648, strict <) without a single lookup: no search kernel, the glue decodes (MaxValue, index 0).
  // (a.empty_cloud is set by the caller that knows the scan: stage_update.)
  const bool searching = (a.step_mode != CS_STEP_INTEGRATE_ONLY) && a.do_search && !a.empty_cloud;
  const bool draws = a.step_mode != CS_STEP_SEARCH_ONLY && n_points > 0;
  const bool fused = searching && phases == CS_PHASE_ALL;
  a.fuse_publish = fused ? 1 : 0;
  a.max_ring_hint = rings - 1;
  a.w_prefetch = 0;  // the map around the pose goes to L2 ahead of its use (one session alone; batches hide the latency)
  a.diag = c.diag;
  a.diag_rings = c.diag_rings;
  a.spec = (tune().spec >= 0 && fused && draws && a.step_mode == CS_STEP_UPDATE) ? c.spec : nullptr;
  a.stuck_flag = c.stuck_dev;
  a.spin_ns = c.stuck_dev ? (long long)(tune().spin_ms > 0 ? tune().spin_ms : 2000) * 1000000ll : 0;
  if (phases & CS_PHASE_FINISH) {  // a call that publishes a pose takes the next step id: never 0, alternating parity
    if (++(*c.step_counter) == 0) *c.step_counter = 2;
    a.step_id = *c.step_counter;
  }
  cudaError_t e = cudaSuccess;
  S2Plan s2;
  bool used_s2 = false;
  if ((phases & CS_PHASE_SEARCH) && searching &&
      cs_plan_search2(c.n_sessions, c.s2_cap, c.s2_min_cand, a.cand_count, a.s2_host_points, c.num_sms, &s2,
                      a.spec ? CS_S2_MAX_THREADS - 32 : CS_S2_MAX_THREADS)) {  // (the glue table's service warp rides along)
    used_s2 = true;
    a.s2_points = s2.points;
    a.s2_slab = s2.threads;
    a.s2_slot = (*c.s2_toggle ^= 1);
    a.s2_batch = c.n_sessions > 1 ? 1 : 0;
    // (Measured and dropped: letting the candidate sort — the first kernel of the step, one block — prefetch the map around
    // the search pose: 65 k prefetches from one SM take 40 us.  CS_TUNE_W_PREFETCH=1 still selects it for experiments.)
    a.s2_map = c.hs->map;
    a.s2_sorted = c.hs->s2_sorted + (size_t)a.s2_slot * c.hs->s2_cap;
    a.s2_tmp = c.hs->s2_tmp;
    a.s2_meta = c.hs->s2_meta;
    a.s2_acc = c.hs->s2_acc;
    a.s2_ghist = c.hs->s2_ghist;
    a.s2_seed = c.hs->seed;
    a.s2_size = c.hs->size;
    a.s2_pitch_tiles = c.hs->pitch_tiles;
    a.s2_scale = c.hs->scale;
    a.s2_sigma_xy = c.hs->sigma_xy;
    a.s2_sigma_theta = c.hs->sigma_theta;
    // (Measured and dropped: a warm-up kernel in front of the sort, one block per SM loading the map within reach of the step while
    // the one-block sort runs: one more link in the dependency chain costs more than the cold lookups — cfg2 41.8 -> 42.6 us.)
    const bool presorted = c.presort && c.presort->valid && !c.ev_done /* (CS_FLAG_TIMING measures the whole stage) */ && c.presort->cand_mode == a.cand_mode && c.presort->scan_index == a.scan_index &&
                           c.presort->cand_first == a.cand_first && c.presort->cand_count == a.cand_count && c.presort->slot == a.s2_slot &&
                           (a.cand_mode == CS_CAND_PHILOX || (a.cand_mode == CS_CAND_OFFSETS && c.presort->cand == a.cand && c.presort->tag == c.log_tag && c.log_tag != 0));
    if (c.presort) c.presort->valid = false;
    if (!presorted) {
      e = launch_sort(c, a, /*warm=*/true);
      if (e != cudaSuccess) return e;
    }
    dispatch_layout(c.tiled, [&](auto T) {
      e = launch_pdl(cs_search2_kernel<decltype(T)::value>, dim3((unsigned)s2.clusters, (unsigned)s2.slabs, (unsigned)c.n_sessions),
                     dim3(s2.threads + (a.spec ? 32 : 0)), 0, c.stream, c.d_sess, a);
    });
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  } else if ((phases & CS_PHASE_SEARCH) && searching) {
    a.spec = nullptr;  // (the glue table is filled by the slab search's service warps)
    const int warps = cs_search_warps(a.cand_count, c.n_sessions, c.num_sms);
    dim3 grid((unsigned)((a.cand_count + warps - 1) / warps), (unsigned)c.n_sessions);
    int chunk = (n_points + 1) & ~1;  // even: the staging loop moves two points per 16-byte load
    if (chunk > CS_SEARCH_CHUNK) chunk = CS_SEARCH_CHUNK;
    if (chunk < 2) chunk = 2;
    a.search_chunk = chunk;
    dispatch_layout(c.tiled, [&](auto T) {
      e = launch_pdl(cs_search_kernel<decltype(T)::value>, grid, dim3(warps * 32), (size_t)chunk * sizeof(float2), c.stream,
                     c.d_sess, a);
    });
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  }
  if (!(phases & CS_PHASE_FINISH)) return cudaGetLastError();
  if (!fused) {
    e = launch_pdl(cs_setup_kernel, dim3(1, (unsigned)c.n_sessions), dim3(CS_SETUP_THREADS), 0, c.stream, c.d_sess, a);
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  }
  // the pose is out once this kernel has finished
  if (c.ev_pose) cudaEventRecord(c.ev_pose, c.stream);
  if (draws && tune().integrate != 1 && c.w_slot) {
    // The wedge kernel: (ring range x angular wedge) tasks, one warp each (cs_wedge.cuh).  The first blocks prepare the rays,
    // 128 each (big scans: bigger groups, so that at most 64 blocks prepare); every warp of the grid then takes tasks (small
    // scans: by block tickets; otherwise round-robin).  One session alone gets a resident wave — three blocks per SM of the
    // 80-register instance for a small scan, four of the 64-register one for a big scan — a session of a batch four blocks.
    int group = (((n_points + 63) / 64) + 31) / 32 * 32;
    if (group < 128) group = 128;
    if (group > CS_W_THREADS) group = CS_W_THREADS;  // one ray per thread of a preparing block
    a.prep_group = group;
    const int nprep = (n_points + group - 1) / group;
    // (one small scan alone: the 80-register instance, three blocks per SM; see cs_wedge_kernel)
    const bool latency = c.n_sessions == 1 && n_points <= 2048 && tune().w_resident != 4;
    int blocks = c.n_sessions == 1 ? (latency ? 3 : 4) * c.num_sms : 4;
    if (tune().w_blocks > 0) blocks = tune().w_blocks;
    if (blocks < nprep) blocks = nprep;
    // A rank of a candidate-split group that shares its device with another rank (several handles of one process on one GPU)
    // must not keep a grid of polling blocks resident while its search kernel waits for the other ranks' keys — the ranks
    // would starve each other: its draw kernel starts after the search.  Ranks on devices of their own overlap as usual.
    if (a.xchg_world > 1 && a.xchg_serial_draw) g_next_launch_plain = true;
    {  // the counters rotate through three thirds: counted into now / counted into by the previous drawn step / zeroed now
      const int d = (*c.w_slot)++;
      a.w_slot = d % 3;
      // The previous scan's counts size this scan's wedges (the table is then ready before the pose) where a scan is a latency
      // chain; a big scan is throughput, its wedges are narrow (a few degrees of heading change between scans would
      // overfill some) and it builds the table from its own counts.
      a.w_prev = (d > 0 && tune().w_prev >= 0 && n_points <= 2048) ? (d + 2) % 3 : -1;
      a.w_zero = (d + 1) % 3;
      if (*c.w_slot >= 3000) *c.w_slot -= 2997;  // (keeps d > 0 and d mod 3)
    }
    a.w_general = tune().w_general > 0 ? 1 : 0;
    if (a.w_prefetch == 0 && c.n_sessions == 1 && c.tiled && tune().w_prefetch >= 0) a.w_prefetch = 2;
    // (Measured and left as an experiment knob, CS_TUNE_W_PREFETCH=3: in batches, the preparing thread of a ray prefetching the
    // tiles along it — cfg5 2.01 -> 2.08 ms per step: the draw tasks of a batch already keep the memory system busy.)
    if (c.n_sessions > 1 && c.tiled && tune().w_prefetch == 3) a.w_prefetch = 3;
    a.w_sub_max = tune().w_sub;
    CsStepArgs draw_args = a;
    if (tune().fault == 1 && c.stuck_dev) draw_args.step_id ^= 0x40000000u;  // (fault injection: a pose tag nobody publishes)
    dispatch_layout(c.tiled, [&](auto T) {
      e = launch_pdl(latency ? cs_wedge_kernel<decltype(T)::value, 3> : cs_wedge_kernel<decltype(T)::value, 4>,
                     dim3((unsigned)blocks, (unsigned)c.n_sessions), dim3(CS_W_THREADS), wedge_smem(c.hs->size), c.stream, c.d_sess, draw_args);
    });
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  } else if (draws) {
    int threads = ((n_points + CS_RING_RPT - 1) / CS_RING_RPT + 31) / 32 * 32;
    if (tune().ring_threads >= 32) threads = tune().ring_threads / 32 * 32;
    if (threads < 32) threads = 32;
    if (threads > CS_RING_MAX_THREADS) threads = CS_RING_MAX_THREADS;
    if (rings < 1) rings = 1;
    // One session alone: a grid of at most one resident wave (2 blocks per SM) whose blocks draw work units of
    // `span` rings from a ticket counter.  A batch of sessions: one unit of 64 rings per block, grid.y = sessions (measured on cfg5: 8 -> 64 rings per unit, a 2048-entry slot table and the small-block instance of the kernel take the rings from 4.8 to 2.8 ms per step).
    const bool dynamic = c.n_sessions == 1;
    // scans of several rounds (more rays than 2 x threads): two rings per unit, so that a round's rays serve two rings and
    // the second ring's evaluation hides the first one's map loads (cfg3: 208 -> 183 us; 4 and more rings per unit lose to imbalance)
    const bool multi_round = n_points > threads * CS_RING_RPT;
    // Batches: 64 rings per block; 32 for batches of at most 256 sessions, where the grid is only a wave or two of blocks and
    // shorter blocks leave a shorter tail (128 sessions, the per-GPU share of cfg5 on eight GPUs: 0.49 -> 0.455 ms per step).
    int span = dynamic ? (multi_round ? 2 : 1) : (c.n_sessions <= 256 ? 32 : 64);
    if (tune().ring_span > 0) span = tune().ring_span;
    if (span < 1) span = 1;
    if (span > CS_RING_MAX_SPAN) span = CS_RING_MAX_SPAN;
    a.ring_span = span;
    a.ring_dynamic = dynamic ? 1 : 0;
    // slot table: a session alone keeps every ring up to k = 1024 in one window; a batch trades windows on its outer
    // rings for resident blocks (shared memory per block is what limits them)
    a.ring_slot_bits = dynamic ? CS_RING_MAX_SLOT_BITS : 11;
    if (tune().ring_slot_bits >= 8 && tune().ring_slot_bits <= CS_RING_MAX_SLOT_BITS) a.ring_slot_bits = tune().ring_slot_bits;
    int blocks = (rings + span - 1) / span;
    const int per_sm = tune().ring_blocks_per_sm > 0 ? tune().ring_blocks_per_sm : 2;
    if (dynamic && blocks > per_sm * c.num_sms) blocks = per_sm * c.num_sms;
    // the first blocks also prepare the rays, 128 each (one warp per SM sub-partition: the preparation is a
    // latency chain, not throughput); big scans use bigger groups so that at most 64 blocks prepare
    int group = (((n_points + 63) / 64) + 31) / 32 * 32;
    if (group < 128) group = 128;
    a.prep_group = group;
    const int nprep = (n_points + group - 1) / group;
    if (blocks < nprep) blocks = nprep;
    const bool small_rings = !dynamic && !multi_round && threads <= CS_RING_SMALL_THREADS && tune().ring_small >= 0;
    dispatch_layout(c.tiled, [&](auto T) {
      constexpr bool tiled = decltype(T)::value;
      e = launch_pdl(small_rings ? cs_rings_kernel<tiled, true, false>
                                 : (multi_round ? cs_rings_kernel<tiled, false, true> : cs_rings_kernel<tiled, false, false>),
                     dim3((unsigned)blocks, (unsigned)c.n_sessions), dim3(threads),
                     CS_RING_SMEM(threads, a.ring_slot_bits, multi_round), c.stream, c.d_sess, a);
    });
    if (e != cudaSuccess) return e;
    (*c.launches)++;
  }
  // ---- production mode: the next scan's candidate sort, queued in the shadow of this scan's draw kernel (see Presort).  Not
  // under CS_FLAG_TIMING (the events around the stages would count it to the wrong one).
  const bool ahead_table = a.cand_mode == CS_CAND_OFFSETS && c.next_cand != nullptr && c.log_tag != 0 && !a.cand_cs;
  if (c.presort && tune().presort >= 0 && used_s2 && fused && draws && (a.cand_mode == CS_CAND_PHILOX || ahead_table) && c.n_sessions == 1 &&
      !c.ev_done) {
    CsStepArgs b = a;
    if (ahead_table) b.cand = c.next_cand;
    b.scan_index = a.scan_index + 1;
    b.s2_slot = a.s2_slot ^ 1;
    b.s2_sorted = c.hs->s2_sorted + (size_t)b.s2_slot * c.hs->s2_cap;
    b.w_prefetch = 0;  // (the next scan's header is not known yet: its draw kernel's blocks warm the map themselves)
    e = launch_sort(c, b, /*warm=*/false);
    if (e != cudaSuccess) return e;
    c.presort->valid = true;
    c.presort->scan_index = b.scan_index;
    c.presort->cand_first = a.cand_first;
    c.presort->cand_count = a.cand_count;
    c.presort->slot = b.s2_slot;
    c.presort->cand_mode = a.cand_mode;
    c.presort->cand = ahead_table ? c.next_cand : nullptr;
    c.presort->tag = ahead_table ? c.log_tag : 0;
  }
  if (c.ev_done) cudaEventRecord(c.ev_done, c.stream);
  return cudaGetLastError();
}

// UpdateObstacleMap (CoreSLAMProcessor.cs:540-593) for n_sessions sessions: the ray kernel (hits + no-hit marks), then the
// sweep over the marked tiles.  Plain stream order: both follow the rings kernel of the same step.
cudaError_t launch_obstacle_update(cudaStream_t stream, const CsSession* d_sess, const CsObstacle* d_obst, const CsObstacle& ho,
                                   const CsStepArgs& a, int n_points, int n_sessions, int num_sms, uint64_t* launches) {
  const int warps = CS_OBST_RAY_THREADS / 32;
  int blocks = (n_points + warps - 1) / warps;
  if (blocks > num_sms * 8) blocks = num_sms * 8;
  if (blocks < 1) blocks = 1;
  cs_obstacle_rays_kernel<<<dim3((unsigned)blocks, (unsigned)n_sessions), CS_OBST_RAY_THREADS, 0, stream>>>(d_sess, d_obst, a);
  (*launches)++;
  const int n_words = ho.words_per_row * (ho.rows / 4);
  int sblocks = (n_words + 255) / 256;
  if (sblocks > num_sms * 8) sblocks = num_sms * 8;
  cs_obstacle_sweep_kernel<<<dim3((unsigned)sblocks, (unsigned)n_sessions), 256, 0, stream>>>(d_obst);
  (*launches)++;
  return cudaGetLastError();
}

cs_status launch_step(cs_processor* h, CsStepArgs a, int n_points, int rings, bool timing, int ev_base, int phases = CS_PHASE_ALL,
                      const float* next_cand = nullptr, uint64_t log_tag = 0) {
  LaunchCtx c{};
  c.num_sms = device_sm_count(h->device);
  c.stream = h->stream;
  c.d_sess = h->d_sess;
  c.tiled = h->tiled;
  c.n_sessions = 1;
  c.launches = &h->launches;
  c.step_counter = &h->step_counter;
  c.w_slot = &h->w_slot;
  c.diag = h->d_ring_cycles;
  c.diag_rings = h->size;
  c.s2_cap = h->s2_cap;
  c.s2_min_cand = cs_s2_min_cand(h->cfg.flags);
  c.s2_toggle = &h->s2_toggle;
  c.presort = &h->presort;
  c.next_cand = next_cand;
  c.log_tag = log_tag;
  c.hs = &h->hs;
  c.spec = h->d_spec;
  if (h->cfg.flags & CS_FLAG_DEBUG_BOUNDED_SPIN) c.stuck_dev = reinterpret_cast<volatile unsigned*>(h->d_slot + 96);
  a.hdr_stride = 1;
  const bool want_pose_event = timing || (a.seq_flag && (h->cfg.flags & CS_FLAG_NO_HOST_SPIN));
  c.ev_pose = (want_pose_event && (phases & CS_PHASE_FINISH)) ? h->tm.ev[ev_base + 0] : nullptr;
  c.ev_done = (timing && (phases & CS_PHASE_FINISH)) ? h->tm.ev[ev_base + 1] : nullptr;
  CS_CUDA(h, launch_step_ctx(c, a, n_points, rings, phases));
  if (h->d_obst && (phases & CS_PHASE_FINISH) && a.step_mode != CS_STEP_SEARCH_ONLY && n_points > 0) {
    // UpdateObstacleMap (:751) follows UpdateHoleMap (:750): same pose (CsSession::cur_pose), same cloud
    CS_CUDA(h, launch_obstacle_update(c.stream, h->d_sess, h->d_obst, h->ho, a, n_points, 1, c.num_sms, &h->launches));
  }
  if (c.ev_done) cudaEventRecord(h->tm.ev[ev_base + 2], h->stream);
  if (h->copy_stream) CS_CUDA(h, cudaEventRecord(h->ev_stage_free[h->stage_cur], h->stream));  // this step's readers of its staging buffer are enqueued
  return CS_OK;
}

// A handle that belongs to a candidate-split group evaluates its slice of the flat candidate indices and exchanges the
// arg-min inside the search kernel (cs_exchange_min); the exchange number advances with every searched scan, identically on
// every rank.
void apply_group(cs_processor* h, CsStepArgs& a) {
  if (h->group_world <= 1 || !a.do_search || a.empty_cloud) return;
  const long long n_flat = (long long)h->n_cand + 1;
  const long long lo = (long long)h->group_rank * n_flat / h->group_world, hi = (long long)(h->group_rank + 1) * n_flat / h->group_world;
  a.cand_first = (int)lo;
  a.cand_count = (int)(hi - lo);
  a.xchg_world = h->group_world;
  a.xchg_serial_draw = h->group_shares_device ? 1 : 0;
  a.xchg_rank = h->group_rank;
  a.xchg_seq = ++h->xchg_seq;
  for (int p = 0; p < h->group_world; p++) a.xchg_peer[p] = h->xchg_peer[p];
}

cs_status wait_for_pose(cs_processor* h, unsigned seq, cudaEvent_t fallback_event) {
  auto t0 = std::chrono::steady_clock::now();
  if (h->cfg.flags & CS_FLAG_NO_HOST_SPIN) {
    CS_CUDA(h, cudaEventSynchronize(fallback_event));
  } else {
    volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(h->h_slot + 64);
    unsigned spins = 0;
    while (*flag != seq) {
      if ((++spins & 0x3fff) == 0) {
        cudaError_t e = cudaStreamQuery(h->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady)
          return fail(h, CS_ERR_CUDA, "stream error while waiting for the pose: %s", cudaGetErrorString(e));
        if (e == cudaSuccess && *flag != seq)
          return fail(h, CS_ERR_CUDA, "stream drained but the pose flag was never written");
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  auto t1 = std::chrono::steady_clock::now();
  h->last_timing.host_wait_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  return CS_OK;
}

void copy_result(cs_result* out, const CsDevResult* r) {
  out->pose[0] = r->pose[0]; out->pose[1] = r->pose[1]; out->pose[2] = r->pose[2];
  out->distance = r->distance;
  out->index = r->index;
  out->searched = r->searched;
  out->visits = r->visits;
}

cs_status collect_timing(cs_processor* h, bool had_h2d) {
  CS_CUDA(h, cudaEventSynchronize(h->tm.ev[4]));
  float ms = 0;
  cs_timing& t = h->last_timing;
  if (had_h2d) { cudaEventElapsedTime(&ms, h->tm.ev[0], h->tm.ev[1]); t.h2d_ms = ms; } else t.h2d_ms = 0;
  cudaEventElapsedTime(&ms, h->tm.ev[1], h->tm.ev[2]); t.search_ms = ms;
  cudaEventElapsedTime(&ms, h->tm.ev[2], h->tm.ev[3]); t.integrate_ms = ms;  // rings kernel (+ set-up kernels of big scans)
  t.finalize_ms = 0.f;
  cudaEventElapsedTime(&ms, h->tm.ev[0], h->tm.ev[4]); t.total_device_ms = ms;
  t.obstacle_ms = 0.f;
  if (h->d_obst) { cudaEventElapsedTime(&ms, h->tm.ev[3], h->tm.ev[4]); t.obstacle_ms = ms; }
  return CS_OK;
}

bool finite3(const float* p) { return p[0] == p[0] && p[1] == p[1] && p[2] == p[2]; }

}  // namespace

// =====================================================================================================
extern "C" {

int32_t cs_abi_version(void) { return CS_ABI_VERSION; }

const char* cs_last_error(const cs_processor* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int32_t cs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

cs_status cs_create(const cs_config* cfg, cs_processor** out) {
  if (!cfg || !out) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_create: null argument");
  *out = nullptr;
  if (cfg->hole_map_size < 8 || cfg->hole_map_size > 16384)
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "hole_map_size must be in [8, 16384], got %d", cfg->hole_map_size);
  if (!(cfg->physical_map_size > 0.f))
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "physical_map_size must be positive");
  if (cfg->iterations_per_thread < 0 || cfg->iterations_per_thread > (1 << 24))
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "iterations_per_thread out of range");
  const int threads = cfg->num_search_threads > 0 ? cfg->num_search_threads : 1;
  const long long n_cand = (long long)threads * cfg->iterations_per_thread;
  if (n_cand > (1ll << 26)) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "too many candidates per scan");
  const int max_points = cfg->max_points > 0 ? cfg->max_points : 16384;
  if (max_points > 65536) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "max_points must be <= 65536");
  if (cfg->obstacle_map_size < 0 || cfg->obstacle_map_size > 16384)
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "obstacle_map_size must be 0 (no ObstacleMap) or in [1, 16384], got %d",
                cfg->obstacle_map_size);

  int ndev = cs_device_count();
  if (ndev <= 0)
    return fail(nullptr, CS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path (cudaGetDeviceCount = 0)");
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", cfg->device, ndev);

  cs_processor* h = new cs_processor();
  h->cfg = *cfg;
  h->device = cfg->device;
  h->tiled = !(cfg->flags & CS_FLAG_ROW_MAJOR_MAP);
  h->size = cfg->hole_map_size;
  h->pitch_tiles = (h->size + 7) / 8;
  h->max_points = max_points;
  h->n_cand = (int)n_cand;
  h->scale = (float)cfg->hole_map_size / cfg->physical_map_size;  // HoleMap.cs:20
  h->map_cells = h->tiled ? (size_t)h->pitch_tiles * h->pitch_tiles * 64 : (size_t)h->size * h->size;

  auto bail = [&](cs_status st) {
    g_create_error = h->error;
    cs_destroy(h);
    return st;
  };
#define CS_CREATE_CUDA(expr)                                                                           \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      fail(h, _e == cudaErrorMemoryAllocation ? CS_ERR_OUT_OF_MEMORY : CS_ERR_CUDA, "%s failed: %s", #expr, \
           cudaGetErrorString(_e));                                                                    \
      return bail(_e == cudaErrorMemoryAllocation ? CS_ERR_OUT_OF_MEMORY : CS_ERR_CUDA);               \
    }                                                                                                  \
  } while (0)

  CS_CREATE_CUDA(cudaSetDevice(h->device));
  if (cfg->stream) {
    h->stream = (cudaStream_t)cfg->stream;
  } else {
    CS_CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  CS_CREATE_CUDA(cudaMalloc(&h->d_map, h->map_cells * sizeof(uint16_t)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_sess, sizeof(CsSession)));
  if (cfg->flags & CS_FLAG_DEBUG_RAYS) CS_CREATE_CUDA(cudaMalloc(&h->d_ray_dbg, (size_t)max_points * 6 * sizeof(int)));
  h->ray_stride = (max_points + 31) / 32 * 32;
  h->batch_stride = (h->ray_stride / 32 + 31) / 32 * 32;  // whole 128-byte lines per copy
  CS_CREATE_CUDA(cudaMalloc(&h->d_rays, (size_t)kRayCopies * h->ray_stride * sizeof(int4)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_batch_max, (size_t)kRayCopies * h->batch_stride * sizeof(int)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_prep_words, (size_t)2 * kRayCopies * 16 * sizeof(unsigned long long)));
  CS_CREATE_CUDA(cudaMemset(h->d_prep_words, 0, (size_t)2 * kRayCopies * 16 * sizeof(unsigned long long)));
  CS_CREATE_CUDA(rings_allow_shared_memory());
  CS_CREATE_CUDA(cudaMalloc(&h->d_w_rk, (size_t)h->ray_stride * sizeof(int2)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_w_bkey, (size_t)(h->ray_stride / 32 + 1) * sizeof(float2)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_w_top, (size_t)3 * wedge_top_words(h->size) * sizeof(int)));
  CS_CREATE_CUDA(cudaMemset(h->d_w_top, 0, (size_t)3 * wedge_top_words(h->size) * sizeof(int)));
  CS_CREATE_CUDA(wedge_allow_shared_memory());
  CS_CREATE_CUDA(cudaMalloc(&h->d_spec, ((size_t)n_cand + 1) * 8 * sizeof(unsigned long long)));
  CS_CREATE_CUDA(cudaMemset(h->d_spec, 0, ((size_t)n_cand + 1) * 8 * sizeof(unsigned long long)));  // (tag 0 is never a step's)
  CS_CREATE_CUDA(cudaMalloc(&h->d_distances, ((size_t)n_cand + 1) * sizeof(int)));
  CS_CREATE_CUDA(cudaMalloc(&h->d_checksum, sizeof(unsigned long long)));
  if (cs_s2_min_cand(cfg->flags) > 0 && n_cand + 1 >= cs_s2_min_cand(cfg->flags)) {
    h->s2_cap = (int)n_cand + 1;
    CS_CREATE_CUDA(cudaMalloc(&h->d_s2_sorted, (size_t)2 * h->s2_cap * sizeof(float4)));
    CS_CREATE_CUDA(cudaMalloc(&h->d_s2_tmp, (size_t)h->s2_cap * sizeof(float4)));
    CS_CREATE_CUDA(cudaMalloc(&h->d_s2_meta, (size_t)h->s2_cap * sizeof(unsigned long long)));
    CS_CREATE_CUDA(cudaMalloc(&h->d_s2_acc, (size_t)h->s2_cap * sizeof(unsigned long long)));
    CS_CREATE_CUDA(cudaMemsetAsync(h->d_s2_acc, 0, (size_t)h->s2_cap * sizeof(unsigned long long), h->stream));
    CS_CREATE_CUDA(cudaMalloc(&h->d_s2_ghist, (size_t)(CS_SORT_BINS + 1) * sizeof(unsigned)));
    CS_CREATE_CUDA(cudaMemsetAsync(h->d_s2_ghist, 0, (size_t)(CS_SORT_BINS + 1) * sizeof(unsigned), h->stream));
  }
  h->stage_bytes = plan_stage(max_points, (int)n_cand, true).total + 16 * ((size_t)max_points + 1) + 64;  // + segment tables
  CS_CREATE_CUDA(cudaMalloc(&h->d_cloud, (size_t)max_points * sizeof(float2)));
  CS_CREATE_CUDA(cudaHostAlloc(&h->h_stage, h->stage_bytes, cudaHostAllocDefault));
  CS_CREATE_CUDA(cudaMalloc(&h->d_stage, h->stage_bytes));
  if (tune().copy_stream >= 0) {
    CS_CREATE_CUDA(cudaMalloc(&h->d_stage_b, h->stage_bytes));
    CS_CREATE_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CS_CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    for (auto& e : h->ev_stage_free) CS_CREATE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  CS_CREATE_CUDA(cudaHostAlloc(&h->h_slot, 128, cudaHostAllocMapped));
  CS_CREATE_CUDA(cudaHostGetDevicePointer((void**)&h->d_slot, h->h_slot, 0));
  memset(h->h_slot, 0, 128);
  for (auto& e : h->tm.ev) CS_CREATE_CUDA(cudaEventCreate(&e));
  if (cfg->obstacle_map_size > 0) {  // :132-133
    CsObstacle& o = h->ho;
    o.size = cfg->obstacle_map_size;
    o.pitch = (o.size + 7) / 8 * 8;
    o.rows = (o.size + 3) / 4 * 4;
    o.words_per_row = o.pitch / 8;
    o.scale = (float)cfg->obstacle_map_size / cfg->physical_map_size;  // ObstacleMap.cs:20
    o.max_hits = 10;                                                   // :103
    CS_CREATE_CUDA(cudaMalloc(&o.pixels, (size_t)o.pitch * o.rows));
    CS_CREATE_CUDA(cudaMalloc(&o.no_hit, (size_t)o.words_per_row * (o.rows / 4) * sizeof(uint32_t)));
    CS_CREATE_CUDA(cudaMalloc(&o.touched, sizeof(long long)));
    CS_CREATE_CUDA(cudaMalloc(&h->d_obst, sizeof(CsObstacle)));
    CS_CREATE_CUDA(cudaMemcpy(h->d_obst, &o, sizeof(CsObstacle), cudaMemcpyHostToDevice));
  }

  // Optional L2 persistence window over the map (the gathers are served from L2 either way when the
  // map fits the 126 MB L2; the window only matters next to other L2-hungry work).
  {
    cudaDeviceProp prop;
    if ((cfg->flags & CS_FLAG_L2_PERSIST) && cudaGetDeviceProperties(&prop, h->device) == cudaSuccess &&
        prop.persistingL2CacheMaxSize > 0) {
      size_t bytes = h->map_cells * sizeof(uint16_t);
      size_t win = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
      size_t carve = win < (size_t)prop.persistingL2CacheMaxSize ? win : (size_t)prop.persistingL2CacheMaxSize;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = h->d_map;
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = win <= carve ? 1.0f : (float)carve / (float)win;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      }
      cudaGetLastError();
    }
  }

  CsSession& s = h->hs;
  memset(&s, 0, sizeof(s));
  s.map = h->d_map;
  s.size = h->size;
  s.pitch_tiles = h->pitch_tiles;
  s.scale = h->scale;
  s.sigma_xy = cfg->sigma_xy;
  s.sigma_theta = cfg->sigma_theta;
  s.iters = cfg->iterations_per_thread;
  s.threads = cfg->num_search_threads;
  s.n_cand = h->n_cand;
  s.quality = 50;        // :82
  s.hole_width = 0.6f;   // :87
  s.search_begin = 5;    // :92
  s.seed = cfg->seed;
  s.rays = h->d_rays;
  s.batch_max = h->d_batch_max;
  s.ray_copies = kRayCopies;
  s.ray_stride = h->ray_stride;
  s.batch_stride = h->batch_stride;
  s.prep_words = h->d_prep_words;
  s.w_rk = h->d_w_rk;
  s.w_bkey = h->d_w_bkey;
  s.w_top = h->d_w_top;
  s.w_levels = wedge_levels(h->size);
  s.s2_sorted = h->d_s2_sorted;
  s.s2_tmp = h->d_s2_tmp;
  s.s2_meta = h->d_s2_meta;
  s.s2_acc = h->d_s2_acc;
  s.s2_ghist = h->d_s2_ghist;
  s.s2_cap = h->s2_cap;
  s.ray_dbg = h->d_ray_dbg;
  s.distances = (cfg->flags & CS_FLAG_KEEP_DISTANCES) ? h->d_distances : nullptr;
  cs_status st = cs_reset(h);
  if (st != CS_OK) return bail(st);
  *out = h;
  return CS_OK;
#undef CS_CREATE_CUDA
}

cs_status cs_destroy(cs_processor* h) {
  if (!h) return CS_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& e : h->tm.ev)
    if (e) cudaEventDestroy(e);
  cudaFree(h->d_map);
  cudaFree(h->d_linear);
  cudaFree(h->d_packed);
  cudaFree(h->d_sess);
  cudaFree(h->d_rays);
  cudaFree(h->d_batch_max);
  cudaFree(h->d_prep_words);
  for (int p = 0; p < CS_GROUP_MAX; p++)
    if (h->xchg_ipc[p] && h->xchg_peer[p]) cudaIpcCloseMemHandle(h->xchg_peer[p]);
  cudaFree(h->d_xchg);
  cudaFree(h->d_ray_dbg);
  cudaFree(h->d_w_rk);
  cudaFree(h->d_w_bkey);
  cudaFree(h->d_w_top);
  cudaFree(h->d_spec);
  cudaFree(h->d_distances);
  cudaFree(h->d_ring_cycles);
  cudaFree(h->d_checksum);
  cudaFree(h->d_cloud);
  if (h->export_stream) cudaStreamSynchronize(h->export_stream);
  if (h->ev_export_snap) cudaEventDestroy(h->ev_export_snap);
  if (h->ev_export_done) cudaEventDestroy(h->ev_export_done);
  if (h->export_stream) cudaStreamDestroy(h->export_stream);
  cudaFree(h->d_obst_snap);
  cudaFree(h->d_s2_sorted);
  cudaFree(h->d_s2_tmp);
  cudaFree(h->d_s2_meta);
  cudaFree(h->d_s2_acc);
  cudaFree(h->d_s2_ghist);
  cudaFree(h->d_stage);
  cudaFree(h->d_stage_b);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  for (auto& e : h->ev_stage_free)
    if (e) cudaEventDestroy(e);
  cudaFree(h->ho.pixels);
  cudaFree(h->ho.no_hit);
  cudaFree(h->ho.touched);
  cudaFree(h->d_obst);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->h_slot) cudaFreeHost(h->h_slot);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return CS_OK;
}

cs_status cs_reset(cs_processor* h) {  // CoreSLAMProcessor.cs:167-175
  CS_CHECK_HANDLE(h);
  CsSession& s = h->hs;
  memset(s.state, 0, sizeof(s.state));
  for (int k = 0; k < 3; k++) s.state[0].pose[k] = h->cfg.start_pose[k];  // :172
  s.state[0].scan_count = 0;                                               // :174 (lastOdometryPose = 0, :173)
  s.key[0] = s.key[1] = ~0ull;
  s.search_done = 0;
  s.visits_slot[0] = s.visits_slot[1] = 0;
  s.ring_ticket[0] = s.ring_ticket[1] = 0;
  memset(s.ll_pose, 0, sizeof(s.ll_pose));
  h->parity = 0;
  h->scan_count = 0;
  h->update_count = 0;
  h->search_begin = s.search_begin;
  cs_status st = push_session(h);
  if (st != CS_OK) return st;
  st = launch_fill(h, (uint16_t)((CS_TS_OBSTACLE + CS_TS_NO_OBSTACLE) / 2));  // :169
  if (st != CS_OK) return st;
  if (h->d_obst) {  // :170 ArrayEx.Fill(ObstacleMap.Pixels, UnmappedObstacleHits)
    cs_obstacle_fill_kernel<<<148 * 2, 256, 0, h->stream>>>(h->ho, h->unmapped_obstacle_hits);
    h->launches++;
    CS_CUDA(h, cudaGetLastError());
  }
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

// ---- ObstacleMap (CoreSLAM/ObstacleMap.cs; CoreSLAMProcessor.cs:53, :98, :103) --------------------------
#define CS_NEED_OBSTACLE(h)                                                                                  \
  do {                                                                                                       \
    if (!(h)->d_obst) return fail((h), CS_ERR_STATE, "the handle was created without an ObstacleMap (obstacle_map_size = 0)"); \
  } while (0)

cs_status cs_set_unmapped_obstacle_hits(cs_processor* h, int32_t hits) {
  CS_CHECK_HANDLE(h);
  if (hits < -128 || hits > 127) return fail(h, CS_ERR_INVALID_ARGUMENT, "UnmappedObstacleHits is an sbyte");
  h->unmapped_obstacle_hits = hits;  // takes effect at the next Reset, as in the reference (:96)
  return CS_OK;
}

cs_status cs_set_max_obstacle_hits(cs_processor* h, int32_t hits) {
  CS_CHECK_HANDLE(h);
  if (hits < -128 || hits > 127) return fail(h, CS_ERR_INVALID_ARGUMENT, "MaxObstacleHits is an sbyte");
  h->ho.max_hits = hits;
  if (!h->d_obst) return CS_OK;
  CS_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->d_obst) + offsetof(CsObstacle, max_hits), &h->ho.max_hits, sizeof(int),
                             cudaMemcpyHostToDevice, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_get_obstacle_map_info(const cs_processor* h, int32_t* size, float* scale) {
  if (!h) return CS_ERR_INVALID_ARGUMENT;
  if (size) *size = h->ho.size;
  if (scale) *scale = h->ho.scale;
  return CS_OK;
}

cs_status cs_obstacle_map_download(cs_processor* h, int8_t* pixels) {
  CS_CHECK_HANDLE(h);
  CS_NEED_OBSTACLE(h);
  if (!pixels) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pixels");
  const CsObstacle& o = h->ho;
  CS_CUDA(h, cudaMemcpy2DAsync(pixels, (size_t)o.size, o.pixels, (size_t)o.pitch, (size_t)o.size, (size_t)o.size,
                               cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_obstacle_map_upload(cs_processor* h, const int8_t* pixels) {
  CS_CHECK_HANDLE(h);
  CS_NEED_OBSTACLE(h);
  if (!pixels) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pixels");
  const CsObstacle& o = h->ho;
  CS_CUDA(h, cudaMemcpy2DAsync(o.pixels, (size_t)o.pitch, pixels, (size_t)o.size, (size_t)o.size, (size_t)o.size,
                               cudaMemcpyHostToDevice, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_obstacle_map_fill(cs_processor* h, int32_t value) {
  CS_CHECK_HANDLE(h);
  CS_NEED_OBSTACLE(h);
  if (value < -128 || value > 127) return fail(h, CS_ERR_INVALID_ARGUMENT, "ObstacleMap pixels are sbyte");
  cs_obstacle_fill_kernel<<<148 * 2, 256, 0, h->stream>>>(h->ho, value);
  h->launches++;
  CS_CUDA(h, cudaGetLastError());
  return CS_OK;
}

cs_status cs_get_obstacle_visits(cs_processor* h, int64_t* touched) {
  CS_CHECK_HANDLE(h);
  CS_NEED_OBSTACLE(h);
  if (!touched) return fail(h, CS_ERR_INVALID_ARGUMENT, "null out");
  long long v = 0;
  CS_CUDA(h, cudaMemcpyAsync(&v, h->ho.touched, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  *touched = v;
  return CS_OK;
}

cs_status cs_set_quality(cs_processor* h, int32_t quality) {
  CS_CHECK_HANDLE(h);
  if (quality < 1 || quality > 255) return fail(h, CS_ERR_INVALID_ARGUMENT, "Quality must be 1..255 (CoreSLAMProcessor.cs:76-82)");
  h->hs.quality = quality;
  return patch_session(h, offsetof(CsSession, quality), &h->hs.quality, sizeof(int));
}

cs_status cs_set_hole_width(cs_processor* h, float metres) {
  CS_CHECK_HANDLE(h);
  if (!(metres >= 0.f) || !(metres * h->scale < 60000.f))
    return fail(h, CS_ERR_INVALID_ARGUMENT, "HoleWidth*Scale must be in [0, 60000) cells");
  h->hs.hole_width = metres;
  return patch_session(h, offsetof(CsSession, hole_width), &h->hs.hole_width, sizeof(float));
}

cs_status cs_set_position_search_beginning(cs_processor* h, int32_t scans) {
  CS_CHECK_HANDLE(h);
  h->hs.search_begin = scans;
  h->search_begin = scans;
  return patch_session(h, offsetof(CsSession, search_begin), &h->hs.search_begin, sizeof(int));
}

cs_status cs_get_pose(cs_processor* h, float pose[3]) {
  CS_CHECK_HANDLE(h);
  if (!pose) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pose");
  CsState st;
  CS_CUDA(h, cudaMemcpyAsync(&st, reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, state) + h->parity * sizeof(CsState),
                             sizeof(CsState), cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  pose[0] = st.pose[0]; pose[1] = st.pose[1]; pose[2] = st.pose[2];
  return CS_OK;
}

cs_status cs_set_pose(cs_processor* h, const float pose[3], const float last_odometry[3], int32_t scan_count) {
  CS_CHECK_HANDLE(h);
  if (!pose || !last_odometry) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pose");
  CsState st{};
  for (int k = 0; k < 3; k++) { st.pose[k] = pose[k]; st.last_odo[k] = last_odometry[k]; }
  st.scan_count = scan_count;
  h->scan_count = scan_count;
  return patch_session(h, offsetof(CsSession, state) + h->parity * sizeof(CsState), &st, sizeof(CsState));
}

cs_status cs_get_map_info(const cs_processor* h, int32_t* size, float* scale) {
  if (!h) return CS_ERR_INVALID_ARGUMENT;
  if (size) *size = h->size;
  if (scale) *scale = h->scale;
  return CS_OK;
}

// -----------------------------------------------------------------------------------------------------
cs_status cs_search(cs_processor* h, const float* points, int32_t n_points, const float search_pose[3],
                    const float* cand_poses, const float* cand_cs, int32_t n_cand, uint32_t scan_index,
                    cs_result* best, int32_t* distances) {
  NvtxRange nvtx_range("cs_search");
  CS_CHECK_HANDLE(h);
  if (!points || !search_pose || !best || n_points <= 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_search: bad argument");
  if (n_points > h->max_points) return fail(h, CS_ERR_CAPACITY, "n_points %d > max_points %d", n_points, h->max_points);
  if (n_cand < 0 || n_cand > h->n_cand) return fail(h, CS_ERR_CAPACITY, "n_cand %d > T*I = %d", n_cand, h->n_cand);
  if (!cand_poses && n_cand != h->n_cand) return fail(h, CS_ERR_INVALID_ARGUMENT, "Philox mode evaluates exactly T*I candidates");

  StagePlan sp = plan_stage(n_points, n_cand, cand_cs != nullptr);
  h->stage_cur = 0;
  CsStepHeader* hdr = reinterpret_cast<CsStepHeader*>(h->h_stage);
  memset(hdr, 0, kHdrBytes);
  hdr->odo[0] = search_pose[0]; hdr->odo[1] = search_pose[1]; hdr->odo[2] = search_pose[2];
  hdr->n_points = n_points;
  memcpy(h->h_stage + sp.off_points, points, (size_t)n_points * 8);
  if (cand_poses) memcpy(h->h_stage + sp.off_cand, cand_poses, (size_t)n_cand * 12);
  if (cand_cs) memcpy(h->h_stage + sp.off_cs, cand_cs, (size_t)(n_cand + 1) * 8);
  CS_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, sp.total, cudaMemcpyHostToDevice, h->stream));

  // distances requested: point the session at the buffer for this call
  int* dist_ptr = distances ? h->d_distances : h->hs.distances;
  if (dist_ptr != h->hs.distances) {
    CS_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, distances), &dist_ptr,
                               sizeof(int*), cudaMemcpyHostToDevice, h->stream));
  }

  CsStepArgs a{};
  a.hdr = reinterpret_cast<const CsStepHeader*>(h->d_stage);
  a.points = reinterpret_cast<const float2*>(h->d_stage + sp.off_points);
  a.cand = cand_poses ? reinterpret_cast<const float*>(h->d_stage + sp.off_cand) : nullptr;
  a.cand_cs = cand_cs ? reinterpret_cast<const float*>(h->d_stage + sp.off_cs) : nullptr;
  a.result = reinterpret_cast<CsDevResult*>(h->d_slot);
  a.seq_flag = reinterpret_cast<volatile unsigned*>(h->d_slot + 64);
  a.seq_value = ++h->seq;
  a.scan_index = scan_index;
  a.cand_mode = cand_poses ? CS_CAND_ABSOLUTE : CS_CAND_PHILOX;
  a.step_mode = CS_STEP_SEARCH_ONLY;
  a.parity = h->parity;
  a.do_search = 1;
  a.n_cand = n_cand;
  a.cand_first = 0;
  a.cand_count = n_cand + 1;
  a.s2_host_points = n_points;
  cs_status st = launch_step(h, a, 0, 0, false, 0);
  if (st != CS_OK) return st;
  if (distances) {
    CS_CUDA(h, cudaMemcpyAsync(distances, h->d_distances, (size_t)(n_cand + 1) * sizeof(int), cudaMemcpyDeviceToHost,
                               h->stream));
    int* restore = h->hs.distances;
    CS_CUDA(h, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, distances), &restore,
                               sizeof(int*), cudaMemcpyHostToDevice, h->stream));
  }
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  copy_result(best, reinterpret_cast<const CsDevResult*>(h->h_slot));
  best->visits = 0;
  return CS_OK;
}

cs_status cs_integrate(cs_processor* h, const float* points, int32_t n_points, const float pose[3],
                       const float* pose_cs, int64_t* visits) {
  NvtxRange nvtx_range("cs_integrate");
  CS_CHECK_HANDLE(h);
  if (!points || !pose || n_points < 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_integrate: bad argument");
  if (n_points > h->max_points) return fail(h, CS_ERR_CAPACITY, "n_points %d > max_points %d", n_points, h->max_points);
  // No sync needed before refilling the pinned block: the previous call returned only after its pose
  // flag, which the device writes after that call's H2D copy completed; the device-side block is
  // protected by stream order.
  StagePlan sp = plan_stage(n_points, 0, false);
  h->stage_cur = 0;
  CsStepHeader* hdr = reinterpret_cast<CsStepHeader*>(h->h_stage);
  memset(hdr, 0, kHdrBytes);
  hdr->odo[0] = pose[0]; hdr->odo[1] = pose[1]; hdr->odo[2] = pose[2];
  hdr->n_points = n_points;
  if (pose_cs) { hdr->cs[0] = pose_cs[0]; hdr->cs[1] = pose_cs[1]; hdr->has_cs = 1; }
  memcpy(h->h_stage + sp.off_points, points, (size_t)n_points * 8);
  CS_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, sp.total, cudaMemcpyHostToDevice, h->stream));
  CsStepArgs a{};
  a.hdr = reinterpret_cast<const CsStepHeader*>(h->d_stage);
  a.points = reinterpret_cast<const float2*>(h->d_stage + sp.off_points);
  a.result = reinterpret_cast<CsDevResult*>(h->d_slot);
  a.step_mode = CS_STEP_INTEGRATE_ONLY;
  a.parity = h->parity;
  cs_status st = launch_step(h, a, n_points, rings_hint(h, max_range_of(points, n_points)), false, 0);
  if (st != CS_OK) return st;
  if (visits) {
    long long v = 0;
    CS_CUDA(h, cudaMemcpyAsync(&v, reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, visits_slot) + (h->step_counter & 1u) * sizeof(long long), sizeof(v),
                               cudaMemcpyDeviceToHost, h->stream));
    CS_CUDA(h, cudaStreamSynchronize(h->stream));
    *visits = v;
  }
  return CS_OK;
}

// Stages one scan and fills the step arguments of an Update; shared by cs_update and cs_update_begin.
// Raw scan segments of one Update (cs_update_segments): the cloud is computed on the device.
struct SegInput {
  const float* rays;        // n_points * (angle, radius)
  const int32_t* seg_first; // n_segments + 1
  const float* seg_poses;   // n_segments * 3
  int n_segments;
};

static cs_status launch_cloud(cs_processor* h, const uint8_t* d_rays, const uint8_t* d_first, const uint8_t* d_poses, int n_rays,
                              int n_segments, const float odo[3]) {
  cs_cloud_kernel<<<(n_rays + CS_CLOUD_THREADS - 1) / CS_CLOUD_THREADS, CS_CLOUD_THREADS, 0, h->stream>>>(
      reinterpret_cast<const float2*>(d_rays), reinterpret_cast<const int*>(d_first), reinterpret_cast<const float*>(d_poses),
      n_segments, n_rays, odo[0], odo[1], odo[2], h->d_cloud);
  h->launches++;
  CS_CUDA(h, cudaGetLastError());
  return CS_OK;
}

static cs_status check_segments(cs_processor* h, const float* rays, const int32_t* seg_first, const float* seg_poses, int32_t n_rays,
                                int32_t n_segments) {
  // n_rays == 0 is legal: segments whose Rays lists are empty still advance the processor state (:719-747, every distance
  // is int.MaxValue and searchPose wins); only a negative count or missing tables are argument errors
  if ((!rays && n_rays > 0) || !seg_first || !seg_poses || n_rays < 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "segments: bad argument");
  if (n_segments <= 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "Sequence contains no elements (segments.Last(), CoreSLAMProcessor.cs:719)");
  if (n_rays > h->max_points) return fail(h, CS_ERR_CAPACITY, "n_rays %d > max_points %d", n_rays, h->max_points);
  if (n_segments > h->max_points) return fail(h, CS_ERR_CAPACITY, "n_segments %d > max_points %d", n_segments, h->max_points);
  if (seg_first[0] != 0 || seg_first[n_segments] != n_rays) return fail(h, CS_ERR_INVALID_ARGUMENT, "seg_first must run from 0 to n_rays");
  for (int s = 0; s < n_segments; s++)
    if (seg_first[s + 1] < seg_first[s]) return fail(h, CS_ERR_INVALID_ARGUMENT, "seg_first must be non-decreasing");
  return CS_OK;
}

// upper bound of |point| over the cloud of the segments: |segment.Pose - odometry| + |radius|
static double max_range_of_segments(const float* rays, const int32_t* seg_first, const float* seg_poses, int n_segments, const float odo[3]) {
  double m = 0.0;
  bool nan_seen = false;  // a NaN radius or segment pose: "all rings" (rings_hint_of), whatever the other rays reach
  for (int s = 0; s < n_segments; s++) {
    const double dx = (double)seg_poses[3 * s] - odo[0], dy = (double)seg_poses[3 * s + 1] - odo[1];
    const double d0 = std::sqrt(dx * dx + dy * dy);
    for (int i = seg_first[s]; i < seg_first[s + 1]; i++) {
      const double r = d0 + std::fabs((double)rays[2 * i + 1]);
      if (r != r) nan_seen = true;
      else if (r > m) m = r;
    }
  }
  return nan_seen ? std::numeric_limits<double>::quiet_NaN() : m;
}

static cs_status stage_update(cs_processor* h, const float* points, int32_t n_points, const float odometry_pose[3],
                              const float* cand_offsets, bool timing, CsStepArgs* out_args, const SegInput* seg = nullptr,
                              bool split_phase = false) {
  if ((!points && !seg && n_points > 0) || !odometry_pose || n_points < 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_update: bad argument");
  if (n_points > h->max_points) return fail(h, CS_ERR_CAPACITY, "n_points %d > max_points %d", n_points, h->max_points);
  if (!finite3(odometry_pose)) return fail(h, CS_ERR_INVALID_ARGUMENT, "odometry pose is NaN");
  if (h->pending) return fail(h, CS_ERR_STATE, "a split-phase update is in flight: call cs_update_finish first");

  // The previous call's integration may still be running: the pinned block is free again (its H2D copy
  // finished before that call's pose flag), the device block is protected by stream order.
  const bool do_search = h->scan_count >= h->search_begin;  // :726
  const bool with_offsets = cand_offsets != nullptr && do_search;
  StagePlan sp = plan_stage(n_points, with_offsets ? h->n_cand : 0, false);
  CsStepHeader* hdr = reinterpret_cast<CsStepHeader*>(h->h_stage);
  memset(hdr, 0, kHdrBytes);
  hdr->odo[0] = odometry_pose[0]; hdr->odo[1] = odometry_pose[1]; hdr->odo[2] = odometry_pose[2];
  hdr->n_points = n_points;
  size_t off_first = 0, off_poses = 0;
  if (seg) {  // the points slot carries the raw rays; the segment tables go behind the candidate table
    if (n_points > 0) memcpy(h->h_stage + sp.off_points, seg->rays, (size_t)n_points * 8);
    off_first = sp.total;
    off_poses = align_up(off_first + (size_t)(seg->n_segments + 1) * 4, 16);
    sp.total = align_up(off_poses + (size_t)seg->n_segments * 12, 16);
    memcpy(h->h_stage + off_first, seg->seg_first, (size_t)(seg->n_segments + 1) * 4);
    memcpy(h->h_stage + off_poses, seg->seg_poses, (size_t)seg->n_segments * 12);
  } else if (n_points > 0) {
    memcpy(h->h_stage + sp.off_points, points, (size_t)n_points * 8);
  }
  if (with_offsets) memcpy(h->h_stage + sp.off_cand, cand_offsets, (size_t)h->n_cand * 12);

  if (timing) cudaEventRecord(h->tm.ev[0], h->stream);
  uint8_t* d_base = h->d_stage;
  if (h->copy_stream) {
    // run ahead of the stream: copy into the buffer the scan in flight is not reading, behind nothing but that buffer's
    // last readers (two scans ago); the main stream picks the copy up through its event
    h->stage_cur = (h->stage_flip ^= 1);
    d_base = h->stage_cur ? h->d_stage_b : h->d_stage;
    CS_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_stage_free[h->stage_cur], 0));
    CS_CUDA(h, cudaMemcpyAsync(d_base, h->h_stage, sp.total, cudaMemcpyHostToDevice, h->copy_stream));
    CS_CUDA(h, cudaEventRecord(h->ev_copy, h->copy_stream));
    CS_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
  } else {
    h->stage_cur = 0;
    CS_CUDA(h, cudaMemcpyAsync(d_base, h->h_stage, sp.total, cudaMemcpyHostToDevice, h->stream));
  }
  if (seg && n_points > 0) {  // ScanSegmentsToCloud (:723) on the device; the step's first kernel must not start before it is complete
    cs_status cst = launch_cloud(h, d_base + sp.off_points, d_base + off_first, d_base + off_poses, n_points,
                                 seg->n_segments, odometry_pose);
    if (cst != CS_OK) return cst;
    g_next_launch_plain = true;
  }
  if (timing) cudaEventRecord(h->tm.ev[1], h->stream);

  CsStepArgs a{};
  a.hdr = reinterpret_cast<const CsStepHeader*>(d_base);
  a.points = seg ? h->d_cloud : reinterpret_cast<const float2*>(d_base + sp.off_points);
  a.cand = with_offsets ? reinterpret_cast<const float*>(d_base + sp.off_cand) : nullptr;
  a.result = reinterpret_cast<CsDevResult*>(h->d_slot);
  a.seq_flag = reinterpret_cast<volatile unsigned*>(h->d_slot + 64);
  a.seq_value = ++h->seq;
  a.scan_index = h->update_count;
  a.cand_mode = with_offsets ? CS_CAND_OFFSETS : CS_CAND_PHILOX;
  a.step_mode = CS_STEP_UPDATE;
  a.parity = h->parity;
  a.do_search = do_search ? 1 : 0;
  a.n_cand = h->n_cand;
  a.cand_first = 0;
  a.cand_count = h->n_cand + 1;
  a.s2_host_points = n_points;
  a.empty_cloud = (do_search && n_points == 0) ? 1 : 0;
  if (!split_phase) apply_group(h, a);  // (cs_update_begin / _finish: the caller exchanges the key between the phases)
  *out_args = a;
  return CS_OK;
}

// Host-side bookkeeping after an Update has been enqueued, the wait for the pose, and the result copy.
static cs_status complete_update(cs_processor* h, const CsStepArgs& a, bool timing, cs_result* out) {
  h->parity ^= 1;
  h->update_count++;
  if (!a.do_search) h->scan_count++;  // :741
  cs_status st = wait_for_pose(h, a.seq_value, h->tm.ev[2]);
  if (st != CS_OK) return st;
  if (timing) {
    st = collect_timing(h, true);
    if (st != CS_OK) return st;
  }
  CS_CHECK_STUCK(h, h->h_slot + 96, fail);
  if (reinterpret_cast<const CsDevResult*>(h->h_slot)->searched < 0)
    return fail(h, CS_ERR_NCCL, "candidate-split group: a rank never delivered its arg-min (exchange timed out after %.1f s)",
                (double)CS_XCHG_TIMEOUT_NS * 1e-9);
  if (out) {
    copy_result(out, reinterpret_cast<const CsDevResult*>(h->h_slot));
    out->visits = -1;  // counted on the device while the integration runs; cs_get_visits() after cs_sync()
    if (timing) {
      long long v = 0;
      CS_CUDA(h, cudaMemcpyAsync(&v, reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, visits_slot) + (h->step_counter & 1u) * sizeof(long long), sizeof(v),
                                 cudaMemcpyDeviceToHost, h->stream));
      CS_CUDA(h, cudaStreamSynchronize(h->stream));
      out->visits = v;
    }
  }
  return CS_OK;
}

cs_status cs_update(cs_processor* h, const float* points, int32_t n_points, const float odometry_pose[3],
                    const float* cand_offsets, cs_result* out) {
  NvtxRange nvtx_range("cs_update");
  CS_CHECK_HANDLE(h);
  const bool timing = (h->cfg.flags & CS_FLAG_TIMING) != 0;
  CsStepArgs a{};
  cs_status st = stage_update(h, points, n_points, odometry_pose, cand_offsets, timing, &a);
  if (st != CS_OK) return st;
  st = launch_step(h, a, n_points, rings_hint(h, max_range_of(points, n_points)), timing, 2);
  if (st != CS_OK) return st;
  return complete_update(h, a, timing, out);
}

// CoreSLAMProcessor.Update(List<ScanSegment>) whole (:717-752): ScanSegmentsToCloud included.
cs_status cs_update_segments(cs_processor* h, const float* rays, const int32_t* seg_first, const float* seg_poses, int32_t n_rays,
                             int32_t n_segments, const float* cand_offsets, cs_result* out) {
  NvtxRange nvtx_range("cs_update_segments");
  CS_CHECK_HANDLE(h);
  cs_status st = check_segments(h, rays, seg_first, seg_poses, n_rays, n_segments);
  if (st != CS_OK) return st;
  const float* odo = seg_poses + 3 * (size_t)(n_segments - 1);  // :719 odoPose = segments.Last().Pose
  const bool timing = (h->cfg.flags & CS_FLAG_TIMING) != 0;
  SegInput seg{rays, seg_first, seg_poses, n_segments};
  CsStepArgs a{};
  st = stage_update(h, nullptr, n_rays, odo, cand_offsets, timing, &a, &seg);
  if (st != CS_OK) { g_next_launch_plain = false; return st; }
  st = launch_step(h, a, n_rays, rings_hint(h, max_range_of_segments(rays, seg_first, seg_poses, n_segments, odo)), timing, 2);
  g_next_launch_plain = false;
  if (st != CS_OK) return st;
  return complete_update(h, a, timing, out);
}

cs_status cs_segments_to_cloud(cs_processor* h, const float* rays, const int32_t* seg_first, const float* seg_poses, int32_t n_rays,
                               int32_t n_segments, const float odometry_pose[3], float* points_out) {
  CS_CHECK_HANDLE(h);
  cs_status st = check_segments(h, rays, seg_first, seg_poses, n_rays, n_segments);
  if (st != CS_OK) return st;
  if (!odometry_pose || !points_out) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_segments_to_cloud: null argument");
  if (h->pending) return fail(h, CS_ERR_STATE, "a split-phase update is in flight: call cs_update_finish first");
  const size_t off_rays = kHdrBytes;
  const size_t off_first = align_up(off_rays + (size_t)n_rays * 8, 16);
  const size_t off_poses = align_up(off_first + (size_t)(n_segments + 1) * 4, 16);
  const size_t total = align_up(off_poses + (size_t)n_segments * 12, 16);
  memcpy(h->h_stage + off_rays, rays, (size_t)n_rays * 8);
  memcpy(h->h_stage + off_first, seg_first, (size_t)(n_segments + 1) * 4);
  memcpy(h->h_stage + off_poses, seg_poses, (size_t)n_segments * 12);
  CS_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, total, cudaMemcpyHostToDevice, h->stream));
  st = launch_cloud(h, h->d_stage + off_rays, h->d_stage + off_first, h->d_stage + off_poses, n_rays, n_segments, odometry_pose);
  if (st != CS_OK) return st;
  CS_CUDA(h, cudaMemcpyAsync(points_out, h->d_cloud, (size_t)n_rays * 8, cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

// ---- multi-GPU candidate split (BASELINE cfg4) ---------------------------------------------
cs_status cs_update_begin(cs_processor* h, const float* points, int32_t n_points, const float odometry_pose[3],
                          const float* cand_offsets, int32_t cand_first, int32_t cand_count, uint64_t** key_device) {
  NvtxRange nvtx_range("cs_update_begin");
  CS_CHECK_HANDLE(h);
  if (cand_first < 0 || cand_count < 0 || (long long)cand_first + cand_count > (long long)h->n_cand + 1)
    return fail(h, CS_ERR_INVALID_ARGUMENT, "candidate slice [%d, %d) outside [0, T*I+1 = %d)", cand_first,
                cand_first + cand_count, h->n_cand + 1);
  CsStepArgs a{};
  cs_status st = stage_update(h, points, n_points, odometry_pose, cand_offsets, false, &a, nullptr, true);
  if (st != CS_OK) return st;
  a.cand_first = cand_first;
  a.cand_count = cand_count;
  const int rings = rings_hint(h, max_range_of(points, n_points));
  if (cand_count > 0) {
    st = launch_step(h, a, n_points, rings, false, 2, CS_PHASE_SEARCH);
    if (st != CS_OK) return st;
  }
  h->pending = true;
  h->pending_args = a;
  h->pending_points = n_points;
  h->pending_rings = rings;
  if (key_device)  // NULL when this scan runs no search (:726): nothing to exchange
    *key_device = (a.do_search && cand_count > 0)
                      ? reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, key) +
                                                    sizeof(unsigned long long) * (size_t)a.parity)
                      : nullptr;
  return CS_OK;
}

cs_status cs_update_finish(cs_processor* h, cs_result* out) {
  NvtxRange nvtx_range("cs_update_finish");
  CS_CHECK_HANDLE(h);
  if (!h->pending) return fail(h, CS_ERR_STATE, "cs_update_finish without cs_update_begin");
  h->pending = false;
  const CsStepArgs a = h->pending_args;
  cs_status st = launch_step(h, a, h->pending_points, h->pending_rings, false, 2, CS_PHASE_FINISH);
  if (st != CS_OK) return st;
  return complete_update(h, a, false, out);
}

// ---- candidate-split group with the exchange inside the search kernel (SURVEY 8b "cs_group_create", 8e) ------------------
static cs_status group_table(cs_processor* h) {
  if (!h->d_xchg) {
    const size_t bytes = (size_t)2 * CS_GROUP_MAX * 2 * sizeof(unsigned long long);
    CS_CUDA(h, cudaMalloc(&h->d_xchg, bytes));
    CS_CUDA(h, cudaMemset(h->d_xchg, 0, bytes));
  }
  return CS_OK;
}

cs_status cs_group_export(cs_processor* h, cs_ipc_handle* out) {
  CS_CHECK_HANDLE(h);
  if (!out) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_group_export: null out");
  cs_status st = group_table(h);
  if (st != CS_OK) return st;
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(cs_ipc_handle), "cs_ipc_handle too small");
  cudaIpcMemHandle_t mh;
  CS_CUDA(h, cudaIpcGetMemHandle(&mh, h->d_xchg));
  memset(out, 0, sizeof(*out));
  memcpy(out, &mh, sizeof(mh));
  return CS_OK;
}

cs_status cs_group_detach(cs_processor* h) {
  CS_CHECK_HANDLE(h);
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int p = 0; p < CS_GROUP_MAX; p++) {
    if (h->xchg_ipc[p] && h->xchg_peer[p]) cudaIpcCloseMemHandle(h->xchg_peer[p]);
    h->xchg_peer[p] = nullptr;
    h->xchg_ipc[p] = false;
  }
  h->group_world = 0;
  h->group_rank = 0;
  h->group_shares_device = false;
  h->xchg_seq = 0;
  if (h->d_xchg) CS_CUDA(h, cudaMemset(h->d_xchg, 0, (size_t)2 * CS_GROUP_MAX * 2 * sizeof(unsigned long long)));
  cudaGetLastError();
  return CS_OK;
}

static cs_status group_check(cs_processor* h, int32_t rank, int32_t world) {
  if (world < 1 || world > CS_GROUP_MAX || rank < 0 || rank >= world)
    return fail(h, CS_ERR_INVALID_ARGUMENT, "group: rank %d of %d (at most %d ranks)", rank, world, CS_GROUP_MAX);
  if ((long long)h->n_cand + 1 < world) return fail(h, CS_ERR_INVALID_ARGUMENT, "group: fewer candidates than ranks");
  if (h->pending) return fail(h, CS_ERR_STATE, "a split-phase update is in flight");
  return CS_OK;
}

cs_status cs_group_attach(cs_processor* h, int32_t rank, int32_t world, const cs_ipc_handle* handles) {
  CS_CHECK_HANDLE(h);
  cs_status st = group_check(h, rank, world);
  if (st != CS_OK) return st;
  if (!handles && world > 1) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_group_attach: null handles");
  st = cs_group_detach(h);
  if (st == CS_OK) st = group_table(h);
  if (st != CS_OK) return st;
  for (int p = 0; p < world; p++) {
    if (p == rank) { h->xchg_peer[p] = h->d_xchg; continue; }
    cudaIpcMemHandle_t mh;
    memcpy(&mh, &handles[p], sizeof(mh));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cs_group_detach(h);
      return fail(h, CS_ERR_NCCL, "cs_group_attach: cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
    }
    h->xchg_peer[p] = static_cast<unsigned long long*>(ptr);
    h->xchg_ipc[p] = true;
  }
  h->group_rank = rank;
  h->group_world = world;
  return CS_OK;
}

cs_status cs_group_attach_local(cs_processor* h, int32_t rank, int32_t world, cs_processor* const* peers) {
  CS_CHECK_HANDLE(h);
  cs_status st = group_check(h, rank, world);
  if (st != CS_OK) return st;
  if (!peers && world > 1) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_group_attach_local: null peers");
  st = cs_group_detach(h);
  if (st == CS_OK) st = group_table(h);
  if (st != CS_OK) return st;
  for (int p = 0; p < world; p++) {
    cs_processor* q = (p == rank) ? h : peers[p];
    if (!q) { cs_group_detach(h); return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_group_attach_local: peer %d is null", p); }
    if (q != h) {
      if (cudaSetDevice(q->device) != cudaSuccess || group_table(q) != CS_OK) {
        cudaSetDevice(h->device);
        cs_group_detach(h);
        return fail(h, CS_ERR_CUDA, "cs_group_attach_local: peer %d has no exchange table", p);
      }
      cudaSetDevice(h->device);
      if (q->device == h->device) h->group_shares_device = true;
      if (q->device != h->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          cs_group_detach(h);
          return fail(h, CS_ERR_NCCL, "cs_group_attach_local: no peer access to device %d: %s", q->device, cudaGetErrorString(e));
        }
        cudaGetLastError();
      }
    }
    h->xchg_peer[p] = q->d_xchg;
  }
  h->group_rank = rank;
  h->group_world = world;
  return CS_OK;
}

cs_status cs_sync(cs_processor* h) {
  CS_CHECK_HANDLE(h);
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

// -----------------------------------------------------------------------------------------------------
static cs_status export_wait(cs_processor* h) {
  if (h->export_busy) {
    CS_CUDA(h, cudaEventSynchronize(h->ev_export_done));
    h->export_busy = false;
  }
  return CS_OK;
}

static cs_status ensure_linear(cs_processor* h) {
  cs_status st = export_wait(h);  // the staging buffers below may still be read by an export's D2H copy
  if (st != CS_OK) return st;
  if (!h->d_linear) CS_CUDA(h, cudaMalloc(&h->d_linear, (size_t)h->size * h->size * sizeof(uint16_t)));
  return CS_OK;
}

cs_status cs_map_download(cs_processor* h, uint16_t* pixels) {
  NvtxRange nvtx_range("cs_map_download");
  CS_CHECK_HANDLE(h);
  if (!pixels) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pixels");
  cs_status st = ensure_linear(h);
  if (st != CS_OK) return st;
  dispatch_layout(h->tiled, [&](auto T) {
    cs_relayout_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->d_linear, h->size, h->pitch_tiles, 0);
  });
  h->launches++;
  CS_CUDA(h, cudaMemcpyAsync(pixels, h->d_linear, (size_t)h->size * h->size * 2, cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_map_upload(cs_processor* h, const uint16_t* pixels) {
  NvtxRange nvtx_range("cs_map_upload");
  CS_CHECK_HANDLE(h);
  if (!pixels) return fail(h, CS_ERR_INVALID_ARGUMENT, "null pixels");
  cs_status st = ensure_linear(h);
  if (st != CS_OK) return st;
  CS_CUDA(h, cudaMemcpyAsync(h->d_linear, pixels, (size_t)h->size * h->size * 2, cudaMemcpyHostToDevice, h->stream));
  dispatch_layout(h->tiled, [&](auto T) {
    cs_relayout_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->d_linear, h->size, h->pitch_tiles, 1);
  });
  h->launches++;
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_map_fill(cs_processor* h, uint16_t value) {
  CS_CHECK_HANDLE(h);
  cs_status st = launch_fill(h, value);
  if (st != CS_OK) return st;
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_map_packed(cs_processor* h, uint8_t* packed) {
  CS_CHECK_HANDLE(h);
  if (!packed) return fail(h, CS_ERR_INVALID_ARGUMENT, "null packed");
  {
    cs_status wst = export_wait(h);
    if (wst != CS_OK) return wst;
  }
  size_t n = ((size_t)h->size * h->size) / 2;
  if (!h->d_packed) CS_CUDA(h, cudaMalloc(&h->d_packed, n));
  dispatch_layout(h->tiled, [&](auto T) {
    cs_pack_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->size, h->pitch_tiles, h->d_packed);
  });
  h->launches++;
  CS_CUDA(h, cudaMemcpyAsync(packed, h->d_packed, n, cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

// Map export without stalling the update stream: the viewer formats of the reference — Gray16
// (MainWindow.xaml.cs:227-229 writes HoleMap.Pixels into a Gray16 bitmap), 4-bpp packed (HoleMap.GetPackedPixels,
// HoleMap.cs:44-55) and the ObstacleMap's sbyte grid — are snapshotted by a short kernel in stream order (after the
// last integration, before the next), and the device-to-host copy of the snapshot runs on a side stream while the next
// Updates proceed.
cs_status cs_map_export_begin(cs_processor* h, int32_t format, void* dst) {
  CS_CHECK_HANDLE(h);
  if (!dst) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_map_export_begin: null destination");
  if (format < CS_EXPORT_GRAY16 || format > CS_EXPORT_OBSTACLE_I8) return fail(h, CS_ERR_INVALID_ARGUMENT, "unknown export format %d", format);
  if (format == CS_EXPORT_OBSTACLE_I8 && !h->d_obst) return fail(h, CS_ERR_STATE, "this processor has no ObstacleMap");
  if (!h->export_stream) {
    CS_CUDA(h, cudaStreamCreateWithFlags(&h->export_stream, cudaStreamNonBlocking));
    CS_CUDA(h, cudaEventCreateWithFlags(&h->ev_export_snap, cudaEventDisableTiming));
    CS_CUDA(h, cudaEventCreateWithFlags(&h->ev_export_done, cudaEventDisableTiming));
  }
  // the previous export's copy must be through with the snapshot buffers before they are overwritten (device-side wait)
  if (h->export_busy) CS_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_export_done, 0));
  const size_t cells = (size_t)h->size * h->size;
  const void* src = nullptr;
  size_t bytes = 0;
  if (format == CS_EXPORT_GRAY16) {
    if (!h->d_linear) CS_CUDA(h, cudaMalloc(&h->d_linear, cells * sizeof(uint16_t)));
    dispatch_layout(h->tiled, [&](auto T) {
      cs_relayout_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->d_linear, h->size, h->pitch_tiles, 0);
    });
    h->launches++;
    src = h->d_linear; bytes = cells * 2;
  } else if (format == CS_EXPORT_PACKED4) {
    if (!h->d_packed) CS_CUDA(h, cudaMalloc(&h->d_packed, cells / 2));
    dispatch_layout(h->tiled, [&](auto T) {
      cs_pack_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->size, h->pitch_tiles, h->d_packed);
    });
    h->launches++;
    src = h->d_packed; bytes = cells / 2;
  } else {
    const CsObstacle& o = h->ho;
    bytes = (size_t)o.size * o.size;
    if (!h->d_obst_snap) CS_CUDA(h, cudaMalloc(&h->d_obst_snap, bytes));
    CS_CUDA(h, cudaMemcpy2DAsync(h->d_obst_snap, (size_t)o.size, o.pixels, (size_t)o.pitch, (size_t)o.size, (size_t)o.size,
                                 cudaMemcpyDeviceToDevice, h->stream));
    src = h->d_obst_snap;
  }
  CS_CUDA(h, cudaGetLastError());
  CS_CUDA(h, cudaEventRecord(h->ev_export_snap, h->stream));
  CS_CUDA(h, cudaStreamWaitEvent(h->export_stream, h->ev_export_snap, 0));
  CS_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->export_stream));
  CS_CUDA(h, cudaEventRecord(h->ev_export_done, h->export_stream));
  h->export_busy = true;
  return CS_OK;
}

cs_status cs_map_export_wait(cs_processor* h) {
  CS_CHECK_HANDLE(h);
  return export_wait(h);
}

cs_status cs_map_checksum(cs_processor* h, uint64_t* checksum) {
  CS_CHECK_HANDLE(h);
  if (!checksum) return fail(h, CS_ERR_INVALID_ARGUMENT, "null checksum");
  CS_CUDA(h, cudaMemsetAsync(h->d_checksum, 0, sizeof(unsigned long long), h->stream));
  dispatch_layout(h->tiled, [&](auto T) {
    cs_checksum_kernel<decltype(T)::value><<<148 * 8, 256, 0, h->stream>>>(h->d_map, h->size, h->pitch_tiles, h->d_checksum);
  });
  h->launches++;
  unsigned long long v = 0;
  CS_CUDA(h, cudaMemcpyAsync(&v, h->d_checksum, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  *checksum = v;
  return CS_OK;
}

uint64_t cs_host_map_checksum(const uint16_t* pixels, int32_t size) {
  unsigned long long acc = 0;
  const size_t n = (size_t)size * (size_t)size;
  for (size_t i = 0; i < n; i++) acc += ((unsigned long long)pixels[i] + 1ull) * cs_mix64((unsigned long long)i);
  return acc;
}

// -----------------------------------------------------------------------------------------------------
cs_status cs_get_timing(cs_processor* h, cs_timing* t) {
  if (!h || !t) return CS_ERR_INVALID_ARGUMENT;
  *t = h->last_timing;
  return CS_OK;
}

cs_status cs_get_distances(cs_processor* h, int32_t* distances, int32_t count) {
  CS_CHECK_HANDLE(h);
  if (!h->hs.distances) return fail(h, CS_ERR_STATE, "handle was created without CS_FLAG_KEEP_DISTANCES");
  if (!distances || count < 0 || count > h->n_cand + 1) return fail(h, CS_ERR_INVALID_ARGUMENT, "bad count");
  CS_CUDA(h, cudaMemcpyAsync(distances, h->d_distances, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_get_rays(cs_processor* h, int32_t* rays, int32_t n_points) {
  CS_CHECK_HANDLE(h);
  if (!h->d_ray_dbg) return fail(h, CS_ERR_STATE, "handle was created without CS_FLAG_DEBUG_RAYS");
  if (!rays || n_points < 0 || n_points > h->max_points) return fail(h, CS_ERR_INVALID_ARGUMENT, "bad n_points");
  CS_CUDA(h, cudaMemcpyAsync(rays, h->d_ray_dbg, (size_t)n_points * 6 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  return CS_OK;
}

cs_status cs_get_visits(cs_processor* h, int64_t* visits) {
  CS_CHECK_HANDLE(h);
  if (!visits) return fail(h, CS_ERR_INVALID_ARGUMENT, "null visits");
  long long v = 0;
  CS_CUDA(h, cudaMemcpyAsync(&v, reinterpret_cast<uint8_t*>(h->d_sess) + offsetof(CsSession, visits_slot) + (h->step_counter & 1u) * sizeof(long long), sizeof(v),
                             cudaMemcpyDeviceToHost, h->stream));
  CS_CUDA(h, cudaStreamSynchronize(h->stream));
  *visits = v;
  return CS_OK;
}

cs_status cs_get_ring_cycles(cs_processor* h, int64_t* cycles, int32_t count) {
  CS_CHECK_HANDLE(h);
  const size_t slots = ((size_t)h->size + CS_DIAG_SEARCH_BLOCKS) * 8;
  if (count < 0 || (size_t)count > slots) return fail(h, CS_ERR_INVALID_ARGUMENT, "bad count");
  if (!h->d_ring_cycles) {  // first call switches the diagnostics on
    CS_CUDA(h, cudaMalloc(&h->d_ring_cycles, slots * sizeof(long long)));
    CS_CUDA(h, cudaMemsetAsync(h->d_ring_cycles, 0, slots * sizeof(long long), h->stream));
    h->hs.ring_cycles = h->d_ring_cycles;
    return patch_session(h, offsetof(CsSession, ring_cycles), &h->hs.ring_cycles, sizeof(long long*));
  }
  if (cycles && count > 0) {
    CS_CUDA(h, cudaMemcpyAsync(cycles, h->d_ring_cycles, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CS_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return CS_OK;
}

cs_status cs_get_search_plan(cs_processor* h, int32_t n_points, int32_t n_cand, int32_t plan[5]) {
  CS_CHECK_HANDLE(h);
  if (!plan || n_points < 1 || n_cand < 0) return fail(h, CS_ERR_INVALID_ARGUMENT, "cs_get_search_plan: bad argument");
  const int sms = device_sm_count(h->device);
  S2Plan s2;
  // (as cs_update launches it: a block carries the glue table's service warp beside the slab's threads)
  const int cap = (tune().spec >= 0 && h->d_spec) ? CS_S2_MAX_THREADS - 32 : CS_S2_MAX_THREADS;
  if (cs_plan_search2(1, h->s2_cap, cs_s2_min_cand(h->cfg.flags), (long long)n_cand + 1, n_points, sms, &s2, cap)) {
    plan[0] = 1; plan[1] = s2.clusters; plan[2] = s2.slabs; plan[3] = s2.threads; plan[4] = s2.points;
  } else {
    const int warps = cs_search_warps((long long)n_cand + 1, 1, sms);
    plan[0] = 0; plan[1] = (int)(((long long)n_cand + 1 + warps - 1) / warps); plan[2] = 1; plan[3] = warps * 32; plan[4] = n_points;
  }
  return CS_OK;
}

cs_status cs_get_launch_count(cs_processor* h, uint64_t* launches) {
  if (!h || !launches) return CS_ERR_INVALID_ARGUMENT;
  *launches = h->launches;
  return CS_OK;
}

cs_status cs_pinned_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return CS_ERR_INVALID_ARGUMENT;
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    fail(nullptr, CS_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return CS_ERR_OUT_OF_MEMORY;
  }
  return CS_OK;
}

cs_status cs_pinned_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
  return CS_OK;
}

// -----------------------------------------------------------------------------------------------------
cs_status cs_scanlog_create(int32_t device, int32_t n_scans, int32_t max_points, int32_t n_offsets, cs_scanlog** out) {
  if (!out || n_scans <= 0 || max_points <= 0 || n_offsets < 0 || max_points > 65536)
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_create: bad argument");
  *out = nullptr;
  if (cs_device_count() <= device) return fail(nullptr, CS_ERR_NO_DEVICE, "no such CUDA device");
  cs_scanlog* log = nullptr;
  try {  // a file header with a huge n_scans must come back as a status, not as std::bad_alloc through the C boundary
    log = new cs_scanlog();
    log->device = device;
    log->n_scans = n_scans;
    log->max_points = (max_points + 1) & ~1;  // keep every scan's points 16-byte aligned
    log->n_offsets = n_offsets;
    log->h_hdr.assign((size_t)n_scans, CsStepHeader{});
    log->h_points.assign((size_t)n_scans * log->max_points * 2, 0.f);
    log->h_offsets.assign((size_t)n_scans * n_offsets * 3, 0.f);
    log->h_max_range.assign((size_t)n_scans, 0.0);
  } catch (const std::exception&) {  // bad_alloc, length_error
    delete log;
    return fail(nullptr, CS_ERR_OUT_OF_MEMORY, "cs_scanlog_create: host allocation failed (%d scans x %d points, %d offsets)", n_scans,
                max_points, n_offsets);
  }
  cudaSetDevice(device);
  bool ok = cudaMalloc(&log->d_hdr, sizeof(CsStepHeader) * n_scans) == cudaSuccess &&
            cudaMalloc(&log->d_points, sizeof(float2) * (size_t)n_scans * log->max_points) == cudaSuccess &&
            cudaMalloc(&log->d_offsets, sizeof(float) * 3 * (size_t)n_scans * (n_offsets > 0 ? n_offsets : 1)) == cudaSuccess &&
            cudaMalloc(&log->d_results, sizeof(CsDevResult) * n_scans) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    cs_scanlog_destroy(log);
    return fail(nullptr, CS_ERR_OUT_OF_MEMORY, "cs_scanlog_create: device allocation failed");
  }
  *out = log;
  return CS_OK;
}

cs_status cs_scanlog_set(cs_scanlog* log, int32_t scan, const float* points, int32_t n_points,
                         const float odometry_pose[3], const float* cand_offsets) {
  if (!log || !points || !odometry_pose || scan < 0 || scan >= log->n_scans || n_points <= 0 || n_points > log->max_points)
    return CS_ERR_INVALID_ARGUMENT;
  CsStepHeader& hd = log->h_hdr[scan];
  memset(&hd, 0, sizeof(hd));
  hd.odo[0] = odometry_pose[0]; hd.odo[1] = odometry_pose[1]; hd.odo[2] = odometry_pose[2];
  hd.n_points = n_points;
  memcpy(&log->h_points[(size_t)scan * log->max_points * 2], points, (size_t)n_points * 8);
  log->h_max_range[scan] = max_range_of(points, n_points);
  if (cand_offsets && log->n_offsets > 0)
    memcpy(&log->h_offsets[(size_t)scan * log->n_offsets * 3], cand_offsets, (size_t)log->n_offsets * 12);
  log->uploaded = false;
  return CS_OK;
}

cs_status cs_scanlog_upload(cs_scanlog* log) {
  if (!log) return CS_ERR_INVALID_ARGUMENT;
  cudaSetDevice(log->device);
  bool ok = cudaMemcpy(log->d_hdr, log->h_hdr.data(), sizeof(CsStepHeader) * log->n_scans, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(log->d_points, log->h_points.data(), log->h_points.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok && log->n_offsets > 0)
    ok = cudaMemcpy(log->d_offsets, log->h_offsets.data(), log->h_offsets.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) return fail(nullptr, CS_ERR_CUDA, "cs_scanlog_upload: %s", cudaGetErrorString(cudaGetLastError()));
  log->uploaded = true;
  static std::atomic<uint64_t> g_log_generation{1};
  log->generation = g_log_generation.fetch_add(1);
  return CS_OK;
}

// ---- scan-log files ---------------------------------------------------------------
// The reference has no log format (the simulator generates scans live, MainWindow.xaml.cs:380-407); this one lets a
// recorded drive be replayed through both implementations.  Little-endian, see include/coreslam_b200.h.
namespace {
struct CslgHeader {
  char magic[4];
  uint32_t version, n_scans, max_points, n_offsets, reserved[3];
};
static_assert(sizeof(CslgHeader) == 32, "CSLG header");

struct FileCloser {
  FILE* f;
  ~FileCloser() { if (f) fclose(f); }
};

cs_status cslg_open(const char* path, FILE** f, CslgHeader* hd) {
  if (!path) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "scan log: null path");
  *f = fopen(path, "rb");
  if (!*f) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "scan log: cannot open %s", path);
  if (fread(hd, sizeof(*hd), 1, *f) != 1 || memcmp(hd->magic, "CSLG", 4) != 0 || hd->version != 1 || hd->n_scans == 0 ||
      hd->max_points == 0 || hd->max_points > 65536 || hd->n_offsets > (1u << 26)) {
    fclose(*f);
    *f = nullptr;
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "scan log: %s is not a CSLG version 1 file", path);
  }
  return CS_OK;
}

// reads the next record; points/offsets may be NULL to skip their payload
cs_status cslg_read_record(FILE* f, const CslgHeader& hd, int32_t* n_points, float odo[3], float* points, float* offsets) {
  uint32_t n = 0;
  if (fread(&n, 4, 1, f) != 1 || n == 0 || n > hd.max_points || fread(odo, 4, 3, f) != 3)
    return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "scan log: truncated or corrupt record");
  *n_points = (int32_t)n;
  const long pts_bytes = (long)n * 8, off_bytes = (long)hd.n_offsets * 12;
  bool ok = points ? fread(points, 8, n, f) == n : fseek(f, pts_bytes, SEEK_CUR) == 0;
  if (ok && hd.n_offsets) ok = offsets ? fread(offsets, 12, hd.n_offsets, f) == hd.n_offsets : fseek(f, off_bytes, SEEK_CUR) == 0;
  return ok ? CS_OK : fail(nullptr, CS_ERR_INVALID_ARGUMENT, "scan log: truncated record");
}
}  // namespace

cs_status cs_scanlog_save(const cs_scanlog* log, const char* path) {
  if (!log || !path) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_save: null argument");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_save: cannot create %s", path);
  FileCloser closer{f};
  CslgHeader hd{};
  memcpy(hd.magic, "CSLG", 4);
  hd.version = 1; hd.n_scans = (uint32_t)log->n_scans; hd.max_points = (uint32_t)log->max_points; hd.n_offsets = (uint32_t)log->n_offsets;
  bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
  for (int k = 0; ok && k < log->n_scans; k++) {
    const CsStepHeader& sh = log->h_hdr[k];
    if (sh.n_points <= 0) return fail(nullptr, CS_ERR_STATE, "cs_scanlog_save: scan %d was never set", k);
    const uint32_t n = (uint32_t)sh.n_points;
    ok = fwrite(&n, 4, 1, f) == 1 && fwrite(sh.odo, 4, 3, f) == 3 &&
         fwrite(&log->h_points[(size_t)k * log->max_points * 2], 8, n, f) == n;
    if (ok && log->n_offsets > 0)
      ok = fwrite(&log->h_offsets[(size_t)k * log->n_offsets * 3], 12, (size_t)log->n_offsets, f) == (size_t)log->n_offsets;
  }
  return ok ? CS_OK : fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_save: write to %s failed", path);
}

cs_status cs_scanlog_file_info(const char* path, int32_t* n_scans, int32_t* max_points, int32_t* n_offsets) {
  FILE* f = nullptr;
  CslgHeader hd;
  cs_status st = cslg_open(path, &f, &hd);
  if (st != CS_OK) return st;
  fclose(f);
  if (n_scans) *n_scans = (int32_t)hd.n_scans;
  if (max_points) *max_points = (int32_t)hd.max_points;
  if (n_offsets) *n_offsets = (int32_t)hd.n_offsets;
  return CS_OK;
}

cs_status cs_scanlog_file_read(const char* path, int32_t scan, float* points, int32_t* n_points, float odometry_pose[3],
                               float* cand_offsets) {
  if (!n_points || !odometry_pose || scan < 0) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_file_read: bad argument");
  FILE* f = nullptr;
  CslgHeader hd;
  cs_status st = cslg_open(path, &f, &hd);
  if (st != CS_OK) return st;
  FileCloser closer{f};
  if ((uint32_t)scan >= hd.n_scans) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_file_read: scan %d of %u", scan, hd.n_scans);
  for (int k = 0; k <= scan; k++) {
    const bool want = k == scan;
    st = cslg_read_record(f, hd, n_points, odometry_pose, want ? points : nullptr, want ? cand_offsets : nullptr);
    if (st != CS_OK) return st;
  }
  return CS_OK;
}

cs_status cs_scanlog_load(int32_t device, const char* path, cs_scanlog** out) {
  if (!out) return fail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_scanlog_load: null out");
  *out = nullptr;
  FILE* f = nullptr;
  CslgHeader hd;
  cs_status st = cslg_open(path, &f, &hd);
  if (st != CS_OK) return st;
  FileCloser closer{f};
  cs_scanlog* log = nullptr;
  st = cs_scanlog_create(device, (int32_t)hd.n_scans, (int32_t)hd.max_points, (int32_t)hd.n_offsets, &log);
  if (st != CS_OK) return st;
  std::vector<float> pts, off;
  try {
    pts.resize((size_t)hd.max_points * 2);
    off.resize((size_t)hd.n_offsets * 3);
  } catch (const std::exception&) {
    cs_scanlog_destroy(log);
    return fail(nullptr, CS_ERR_OUT_OF_MEMORY, "cs_scanlog_load: host allocation failed");
  }
  for (uint32_t k = 0; k < hd.n_scans; k++) {
    int32_t n = 0;
    float odo[3];
    st = cslg_read_record(f, hd, &n, odo, pts.data(), hd.n_offsets ? off.data() : nullptr);
    if (st == CS_OK) st = cs_scanlog_set(log, (int32_t)k, pts.data(), n, odo, hd.n_offsets ? off.data() : nullptr);
    if (st != CS_OK) {
      cs_scanlog_destroy(log);
      return st;
    }
  }
  st = cs_scanlog_upload(log);
  if (st != CS_OK) {
    cs_scanlog_destroy(log);
    return st;
  }
  *out = log;
  return CS_OK;
}

cs_status cs_scanlog_destroy(cs_scanlog* log) {
  if (!log) return CS_OK;
  cudaSetDevice(log->device);
  cudaFree(log->d_hdr);
  cudaFree(log->d_points);
  cudaFree(log->d_offsets);
  cudaFree(log->d_results);
  cudaGetLastError();
  delete log;
  return CS_OK;
}

cs_status cs_replay(cs_processor* h, const cs_scanlog* log, int32_t first, int32_t count, cs_result* results) {
  NvtxRange nvtx_range("cs_replay");
  CS_CHECK_HANDLE(h);
  if (!log || !log->uploaded) return fail(h, CS_ERR_STATE, "scan log not uploaded");
  if (log->device != h->device) return fail(h, CS_ERR_INVALID_ARGUMENT, "scan log lives on another device");
  if (first < 0 || count < 0 || first + count > log->n_scans) return fail(h, CS_ERR_INVALID_ARGUMENT, "scan range");
  if (log->n_offsets > 0 && log->n_offsets != h->n_cand)
    return fail(h, CS_ERR_INVALID_ARGUMENT, "scan log carries %d offsets per scan, handle needs T*I = %d", log->n_offsets, h->n_cand);
  if (log->max_points > h->max_points) return fail(h, CS_ERR_CAPACITY, "scan log max_points exceeds the handle's");
  const bool timing = (h->cfg.flags & CS_FLAG_TIMING) != 0;
  const bool per_kernel = timing && count == 1;
  if (timing) {
    cudaEventRecord(h->tm.ev[0], h->stream);
    if (per_kernel) cudaEventRecord(h->tm.ev[1], h->stream);
  }
  for (int i = 0; i < count; i++) {
    const int sidx = first + i;
    const bool do_search = h->scan_count >= h->search_begin;
    CsStepArgs a{};
    a.hdr = log->d_hdr + sidx;
    a.points = log->d_points + (size_t)sidx * log->max_points;
    a.cand = log->n_offsets > 0 ? log->d_offsets + (size_t)sidx * log->n_offsets * 3 : nullptr;
    a.result = log->d_results + sidx;
    a.visits_out = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(log->d_results + sidx) + offsetof(CsDevResult, visits));
    a.scan_index = h->update_count;
    a.cand_mode = log->n_offsets > 0 ? CS_CAND_OFFSETS : CS_CAND_PHILOX;
    a.step_mode = CS_STEP_UPDATE;
    a.parity = h->parity;
    a.do_search = do_search ? 1 : 0;
    a.n_cand = h->n_cand;
    a.cand_first = 0;
    a.cand_count = h->n_cand + 1;
    a.s2_host_points = log->h_hdr[sidx].n_points;
    apply_group(h, a);
    // (the next scan's table, if the log has one: its sort is queued behind this scan's draw kernel, see Presort)
    const float* next_cand = (log->n_offsets > 0 && sidx + 1 < log->n_scans) ? log->d_offsets + (size_t)(sidx + 1) * log->n_offsets * 3 : nullptr;
    cs_status st = launch_step(h, a, log->h_hdr[sidx].n_points, rings_hint(h, log->h_max_range[sidx]), per_kernel, 2, CS_PHASE_ALL,
                               next_cand, log->generation);
    if (st != CS_OK) return st;
    h->parity ^= 1;
    h->update_count++;
    if (!do_search) h->scan_count++;
  }
  if (timing && !per_kernel) cudaEventRecord(h->tm.ev[4], h->stream);
  if (results && count > 0) {
    static_assert(sizeof(cs_result) == sizeof(CsDevResult), "");
    CS_CUDA(h, cudaMemcpyAsync(results, log->d_results + first, sizeof(CsDevResult) * count, cudaMemcpyDeviceToHost, h->stream));
  }
  // results == NULL: fire and forget — the scans stay queued on the stream (cs_sync / cs_get_pose wait for them)
  if (results || timing) CS_CUDA(h, cudaStreamSynchronize(h->stream));
  if (per_kernel) {
    cs_status st = collect_timing(h, false);
    if (st != CS_OK) return st;
  } else if (timing) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->tm.ev[0], h->tm.ev[4]);
    h->last_timing = cs_timing{};
    h->last_timing.total_device_ms = ms;
  }
  return CS_OK;
}

// =====================================================================================================
// Batches of independent sessions on one GPU (SURVEY 8e "sessions", BASELINE cfg5): one launch per
// kernel for all sessions (grid.y = session), every session with its own map, pose, parameters and
// Philox stream.  No data-path communication between sessions or GPUs.
// =====================================================================================================
}  // extern "C"

struct cs_batch {
  int device = 0, n = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false, tiled = true;
  int size = 0, pitch_tiles = 0, max_points = 0, n_cand = 0;
  int user_max_points = 0;  // the caller's cfg.max_points: stride of the caller's `points` array (max_points above is the
                            // internal stride of the staging / device blocks, rounded up to even for 16-byte alignment)
  float scale = 0.f;
  size_t map_cells = 0;
  std::vector<CsSession> hs;
  CsSession* d_sess = nullptr;
  uint16_t* d_maps = nullptr;
  uint16_t* d_linear = nullptr;
  int4* d_rays = nullptr;
  int* d_batch_max = nullptr;
  unsigned long long* d_prep_words = nullptr;
  unsigned long long* d_checksum = nullptr;
  // wedge integration scratch of all sessions (CsSession::w_*)
  int2* d_w_rk = nullptr;
  float2* d_w_bkey = nullptr;
  int* d_w_top = nullptr;
  int w_slot = 0;
  // slab-search scratch of all sessions (CsSession::s2_*): per session 2*cap sorted + cap tmp entries, cap meta + cap acc words
  float4* d_s2_entries = nullptr;
  unsigned long long* d_s2_words = nullptr;
  int s2_cap = 0, s2_toggle = 0;
  uint32_t flags = 0;
  // staging: [n headers][n * max_points points][n * n_cand offsets]
  size_t off_points = 0, off_cand = 0, stage_bytes = 0;
  uint8_t* h_stage = nullptr;
  uint8_t* d_stage = nullptr;
  CsDevResult* d_results = nullptr;
  CsDevResult* h_results = nullptr;  // pinned
  uint8_t* h_stuck = nullptr;        // CS_FLAG_DEBUG_BOUNDED_SPIN: pinned+mapped stuck-poll word (CsSpin), and its device address
  uint8_t* d_stuck = nullptr;
  int parity = 0, scan_count = 0, search_begin = 5;
  unsigned update_count = 0;
  uint64_t launches = 0;
  unsigned step_counter = 0;
  // cs_batch_submit / cs_batch_collect: two staging + result slots, so that the host stages step k+1 while step k runs
  uint8_t* pipe_h_stage[2] = {nullptr, nullptr};
  uint8_t* pipe_d_stage[2] = {nullptr, nullptr};
  CsDevResult* pipe_d_results[2] = {nullptr, nullptr};
  CsDevResult* pipe_h_results[2] = {nullptr, nullptr};
  cudaEvent_t pipe_done[2] = {nullptr, nullptr};
  unsigned pipe_submitted = 0, pipe_collected = 0;
  std::string error;
  bool poisoned = false;
};

namespace {
cs_status bfail(cs_batch* b, cs_status code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (b) {
    b->error = buf;
    if (code == CS_ERR_CUDA) b->poisoned = true;
  } else {
    g_create_error = buf;
  }
  return code;
}
#define CS_BCUDA(b, expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return bfail((b), CS_ERR_CUDA, "%s failed: %s (line %d)", #expr, cudaGetErrorString(_e), __LINE__); \
  } while (0)
#define CS_CHECK_BATCH(b)                                                           \
  do {                                                                              \
    if (!(b)) return CS_ERR_INVALID_ARGUMENT;                                       \
    if ((b)->poisoned) return CS_ERR_CUDA;                                          \
    CS_CHECK_STUCK(b, (b)->h_stuck, bfail);                                         \
    if (cudaSetDevice((b)->device) != cudaSuccess) return bfail((b), CS_ERR_CUDA, "cudaSetDevice failed"); \
  } while (0)

LaunchCtx batch_ctx(cs_batch* b) {
  LaunchCtx c{};
  c.num_sms = device_sm_count(b->device);
  c.stream = b->stream;
  c.d_sess = b->d_sess;
  c.tiled = b->tiled;
  c.n_sessions = b->n;
  c.launches = &b->launches;
  c.step_counter = &b->step_counter;
  c.w_slot = &b->w_slot;
  c.s2_cap = b->s2_cap;
  c.s2_min_cand = cs_s2_min_cand(b->flags, b->n);
  c.s2_toggle = &b->s2_toggle;
  c.hs = &b->hs[0];
  c.stuck_dev = reinterpret_cast<volatile unsigned*>(b->d_stuck);
  return c;
}

void batch_reset_host(cs_batch* b, const cs_config* cfgs) {
  for (int j = 0; j < b->n; j++) {
    CsSession& s = b->hs[j];
    memset(s.state, 0, sizeof(s.state));
    if (cfgs)
      for (int k = 0; k < 3; k++) s.state[0].pose[k] = cfgs[j].start_pose[k];
    s.key[0] = s.key[1] = ~0ull;
    s.search_done = 0;
    s.visits_slot[0] = s.visits_slot[1] = 0;
    s.ring_ticket[0] = s.ring_ticket[1] = 0;
    memset(s.ll_pose, 0, sizeof(s.ll_pose));
  }
  b->parity = 0;
  b->scan_count = 0;
  b->update_count = 0;
}
}  // namespace

extern "C" {

const char* cs_batch_last_error(const cs_batch* b) { return b ? b->error.c_str() : g_create_error.c_str(); }

cs_status cs_batch_create(const cs_config* cfgs, int32_t n_sessions, cs_batch** out) {
  if (!cfgs || !out || n_sessions <= 0 || n_sessions > 65535)
    return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_batch_create: bad argument (1..65535 sessions)");
  *out = nullptr;
  const cs_config& c0 = cfgs[0];
  if (c0.hole_map_size < 8 || c0.hole_map_size > 16384 || !(c0.physical_map_size > 0.f) || c0.iterations_per_thread < 0)
    return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_batch_create: bad map size / iterations");
  for (int j = 0; j < n_sessions; j++)
    if (cfgs[j].obstacle_map_size != 0)
      return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "cs_batch_create: session batches carry no ObstacleMap (obstacle_map_size must be 0)");
  for (int j = 1; j < n_sessions; j++) {
    const cs_config& c = cfgs[j];
    if (c.hole_map_size != c0.hole_map_size || c.physical_map_size != c0.physical_map_size ||
        c.iterations_per_thread != c0.iterations_per_thread || c.num_search_threads != c0.num_search_threads ||
        c.device != c0.device || c.max_points != c0.max_points || c.flags != c0.flags)
      return bfail(nullptr, CS_ERR_INVALID_ARGUMENT,
                   "cs_batch_create: session %d differs in map size / iterations / threads / device / max_points / flags "
                   "(only start_pose, sigma_xy, sigma_theta and seed may vary)", j);
  }
  const int threads = c0.num_search_threads > 0 ? c0.num_search_threads : 1;
  const long long n_cand = (long long)threads * c0.iterations_per_thread;
  if (n_cand > (1ll << 26)) return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "too many candidates per scan");
  const int max_points = c0.max_points > 0 ? c0.max_points : 16384;
  if (max_points > 65536) return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "max_points must be <= 65536");
  int ndev = cs_device_count();
  if (ndev <= 0) return bfail(nullptr, CS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  if (c0.device < 0 || c0.device >= ndev) return bfail(nullptr, CS_ERR_INVALID_ARGUMENT, "device %d out of range", c0.device);

  cs_batch* b = new cs_batch();
  b->device = c0.device;
  b->n = n_sessions;
  b->tiled = !(c0.flags & CS_FLAG_ROW_MAJOR_MAP);
  b->size = c0.hole_map_size;
  b->pitch_tiles = (b->size + 7) / 8;
  b->user_max_points = max_points;
  b->max_points = (max_points + 1) & ~1;
  b->n_cand = (int)n_cand;
  b->scale = (float)c0.hole_map_size / c0.physical_map_size;
  b->map_cells = b->tiled ? (size_t)b->pitch_tiles * b->pitch_tiles * 64 : (size_t)b->size * b->size;
  b->hs.assign((size_t)n_sessions, CsSession{});
  b->off_points = align_up(sizeof(CsStepHeader) * (size_t)n_sessions, 16);
  b->off_cand = align_up(b->off_points + (size_t)n_sessions * b->max_points * 8, 16);
  b->stage_bytes = align_up(b->off_cand + (size_t)n_sessions * (size_t)n_cand * 12, 16);

  bool ok = cudaSetDevice(b->device) == cudaSuccess;
  if (ok && c0.stream) b->stream = (cudaStream_t)c0.stream;
  else if (ok) { ok = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess; b->own_stream = ok; }
  ok = ok && cudaMalloc(&b->d_maps, b->map_cells * sizeof(uint16_t) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_sess, sizeof(CsSession) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_rays, (size_t)b->max_points * sizeof(int4) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_batch_max, ((size_t)b->max_points / 32 + 1) * sizeof(int) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_prep_words, (size_t)n_sessions * 2 * 16 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMemset(b->d_prep_words, 0, (size_t)n_sessions * 2 * 16 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_checksum, sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_w_rk, (size_t)b->max_points * sizeof(int2) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_w_bkey, ((size_t)b->max_points / 32 + 1) * sizeof(float2) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_w_top, (size_t)n_sessions * 3 * wedge_top_words(b->size) * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMemset(b->d_w_top, 0, (size_t)n_sessions * 3 * wedge_top_words(b->size) * sizeof(int)) == cudaSuccess;
  ok = ok && wedge_allow_shared_memory() == cudaSuccess;
  b->flags = c0.flags;
  {
    const int min_cand = cs_s2_min_cand(c0.flags, n_sessions);
    if (n_sessions > 1 && min_cand > 0 && n_cand + 1 >= min_cand && n_cand + 1 <= CS_SORT_THREADS * CS_SORT_REG) {
      b->s2_cap = n_cand + 1;
      const size_t cap = (size_t)b->s2_cap;
      ok = ok && cudaMalloc(&b->d_s2_entries, (size_t)n_sessions * 3 * cap * sizeof(float4)) == cudaSuccess;
      ok = ok && cudaMalloc(&b->d_s2_words, (size_t)n_sessions * 2 * cap * sizeof(unsigned long long)) == cudaSuccess;
      ok = ok && cudaMemsetAsync(b->d_s2_words, 0, (size_t)n_sessions * 2 * cap * sizeof(unsigned long long), b->stream) == cudaSuccess;
    }
  }
  ok = ok && cudaHostAlloc(&b->h_stage, b->stage_bytes, cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_stage, b->stage_bytes) == cudaSuccess;
  ok = ok && cudaMalloc(&b->d_results, sizeof(CsDevResult) * (size_t)n_sessions) == cudaSuccess;
  ok = ok && cudaHostAlloc(&b->h_results, sizeof(CsDevResult) * (size_t)n_sessions, cudaHostAllocDefault) == cudaSuccess;
  if (b->flags & CS_FLAG_DEBUG_BOUNDED_SPIN) {
    ok = ok && cudaHostAlloc(&b->h_stuck, 64, cudaHostAllocMapped) == cudaSuccess;
    ok = ok && cudaHostGetDevicePointer((void**)&b->d_stuck, b->h_stuck, 0) == cudaSuccess;
    if (ok) memset(b->h_stuck, 0, 64);
  }
  ok = ok && rings_allow_shared_memory() == cudaSuccess;
  if (!ok) {
    cudaError_t e = cudaGetLastError();
    bfail(nullptr, e == cudaErrorMemoryAllocation ? CS_ERR_OUT_OF_MEMORY : CS_ERR_CUDA, "cs_batch_create: %s", cudaGetErrorString(e));
    cs_batch_destroy(b);
    return e == cudaErrorMemoryAllocation ? CS_ERR_OUT_OF_MEMORY : CS_ERR_CUDA;
  }
  for (int j = 0; j < n_sessions; j++) {
    CsSession& s = b->hs[j];
    memset(&s, 0, sizeof(s));
    s.map = b->d_maps + (size_t)j * b->map_cells;
    s.size = b->size;
    s.pitch_tiles = b->pitch_tiles;
    s.scale = b->scale;
    s.sigma_xy = cfgs[j].sigma_xy;
    s.sigma_theta = cfgs[j].sigma_theta;
    s.iters = c0.iterations_per_thread;
    s.threads = c0.num_search_threads;
    s.n_cand = b->n_cand;
    s.quality = 50;
    s.hole_width = 0.6f;
    s.search_begin = 5;
    s.seed = cfgs[j].seed;
    s.rays = b->d_rays + (size_t)j * b->max_points;
    s.batch_max = b->d_batch_max + (size_t)j * ((size_t)b->max_points / 32 + 1);
    s.ray_copies = 1;  // a session of a batch has few blocks: nothing to spread
    s.ray_stride = b->max_points;
    s.batch_stride = b->max_points / 32 + 1;
    s.prep_words = b->d_prep_words + (size_t)j * 2 * 16;
    s.w_rk = b->d_w_rk + (size_t)j * b->max_points;
    s.w_bkey = b->d_w_bkey + (size_t)j * ((size_t)b->max_points / 32 + 1);
    s.w_top = b->d_w_top + (size_t)j * 3 * wedge_top_words(b->size);
    s.w_levels = wedge_levels(b->size);
    if (b->s2_cap > 0) {
      const size_t cap = (size_t)b->s2_cap;
      s.s2_sorted = b->d_s2_entries + (size_t)j * 3 * cap;
      s.s2_tmp = s.s2_sorted + 2 * cap;
      s.s2_meta = b->d_s2_words + (size_t)j * 2 * cap;
      s.s2_acc = s.s2_meta + cap;
      s.s2_cap = b->s2_cap;
    }
  }
  batch_reset_host(b, cfgs);
  cs_fill_kernel<<<148 * 8, 256, 0, b->stream>>>(b->d_maps, b->map_cells * (size_t)n_sessions,
                                                 (uint16_t)((CS_TS_OBSTACLE + CS_TS_NO_OBSTACLE) / 2));
  b->launches++;
  cudaMemcpyAsync(b->d_sess, b->hs.data(), sizeof(CsSession) * (size_t)n_sessions, cudaMemcpyHostToDevice, b->stream);
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
    bfail(nullptr, CS_ERR_CUDA, "cs_batch_create: %s", cudaGetErrorString(cudaGetLastError()));
    cs_batch_destroy(b);
    return CS_ERR_CUDA;
  }
  *out = b;
  return CS_OK;
}

cs_status cs_batch_destroy(cs_batch* b) {
  if (!b) return CS_OK;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->d_maps);
  cudaFree(b->d_linear);
  cudaFree(b->d_sess);
  cudaFree(b->d_rays);
  cudaFree(b->d_batch_max);
  cudaFree(b->d_prep_words);
  cudaFree(b->d_checksum);
  cudaFree(b->d_w_rk);
  cudaFree(b->d_w_bkey);
  cudaFree(b->d_w_top);
  cudaFree(b->d_s2_entries);
  cudaFree(b->d_s2_words);
  cudaFree(b->d_stage);
  cudaFree(b->d_results);
  if (b->h_stage) cudaFreeHost(b->h_stage);
  if (b->h_results) cudaFreeHost(b->h_results);
  if (b->h_stuck) cudaFreeHost(b->h_stuck);
  for (int i = 0; i < 2; i++) {
    cudaFree(b->pipe_d_stage[i]);
    cudaFree(b->pipe_d_results[i]);
    if (b->pipe_h_stage[i]) cudaFreeHost(b->pipe_h_stage[i]);
    if (b->pipe_h_results[i]) cudaFreeHost(b->pipe_h_results[i]);
    if (b->pipe_done[i]) cudaEventDestroy(b->pipe_done[i]);
  }
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  cudaGetLastError();
  delete b;
  return CS_OK;
}

int32_t cs_batch_size(const cs_batch* b) { return b ? b->n : 0; }

// session < 0: all sessions.  Quality 1..255 (:76-82), HoleWidth metres (:87).
cs_status cs_batch_set_params(cs_batch* b, int32_t session, int32_t quality, float hole_width) {
  CS_CHECK_BATCH(b);
  if (session >= b->n) return bfail(b, CS_ERR_INVALID_ARGUMENT, "session %d out of range", session);
  if (quality < 1 || quality > 255) return bfail(b, CS_ERR_INVALID_ARGUMENT, "Quality must be 1..255");
  if (!(hole_width >= 0.f) || !(hole_width * b->scale < 60000.f)) return bfail(b, CS_ERR_INVALID_ARGUMENT, "HoleWidth out of range");
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  for (int j = (session < 0 ? 0 : session); j < (session < 0 ? b->n : session + 1); j++) {
    b->hs[j].quality = quality;
    b->hs[j].hole_width = hole_width;
    CS_BCUDA(b, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(b->d_sess + j) + offsetof(CsSession, quality), &b->hs[j].quality,
                                sizeof(int), cudaMemcpyHostToDevice, b->stream));
    CS_BCUDA(b, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(b->d_sess + j) + offsetof(CsSession, hole_width), &b->hs[j].hole_width,
                                sizeof(float), cudaMemcpyHostToDevice, b->stream));
  }
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  return CS_OK;
}

static float batch_max_hole_width(const cs_batch* b) {
  float w = 0.f;
  for (const CsSession& s : b->hs) w = s.hole_width > w ? s.hole_width : w;
  return w;
}

// One CoreSLAMProcessor.Update (:717-752) for every session of the batch.
//   points        n_sessions * max_points * (x, y); session j uses the first n_points[j]
//   odometry      n_sessions * (x, y, theta)
//   cand_offsets  n_sessions * T*I * (dx, dy, dtheta) verification tables, or NULL: per-session Philox streams
//   results       optional, n_sessions records (blocks until the poses are back)
namespace {
// Stages one Update of every session into (h_stage -> d_stage) and launches its kernels; the result records go to d_results.
cs_status batch_stage_and_launch(cs_batch* b, uint8_t* h_stage, uint8_t* d_stage, CsDevResult* d_results, const float* points,
                                 const int32_t* n_points, const float* odometry, const float* cand_offsets) {
  const bool do_search = b->scan_count >= b->search_begin;
  const bool with_offsets = cand_offsets != nullptr && do_search;
  int max_n = 0;
  double max_range = 0.0;
  bool range_nan = false;  // a NaN point anywhere in the batch: every ring is launched (NaN does not order, so it is tracked apart)
  for (int j = 0; j < b->n; j++) {
    const int np = n_points[j];
    if (np <= 0 || np > b->user_max_points)
      return bfail(b, CS_ERR_CAPACITY, "session %d: n_points %d outside 1..%d", j, np, b->user_max_points);
    const float* odo = odometry + 3 * (size_t)j;
    if (!finite3(odo)) return bfail(b, CS_ERR_INVALID_ARGUMENT, "session %d: odometry pose is NaN", j);
    CsStepHeader* hdr = reinterpret_cast<CsStepHeader*>(h_stage) + j;
    memset(hdr, 0, sizeof(CsStepHeader));
    hdr->odo[0] = odo[0]; hdr->odo[1] = odo[1]; hdr->odo[2] = odo[2];
    hdr->n_points = np;
    const float* src = points + (size_t)j * b->user_max_points * 2;  // the caller's stride is the caller's cfg.max_points
    memcpy(h_stage + b->off_points + (size_t)j * b->max_points * 8, src, (size_t)np * 8);
    const double r = max_range_of(src, np);
    if (r != r) range_nan = true;
    else if (r > max_range) max_range = r;
    if (np > max_n) max_n = np;
  }
  if (range_nan) max_range = std::numeric_limits<double>::quiet_NaN();
  if (with_offsets) memcpy(h_stage + b->off_cand, cand_offsets, (size_t)b->n * b->n_cand * 12);
  const size_t bytes = with_offsets ? b->stage_bytes : b->off_cand;
  CS_BCUDA(b, cudaMemcpyAsync(d_stage, h_stage, bytes, cudaMemcpyHostToDevice, b->stream));

  CsStepArgs a{};
  a.hdr = reinterpret_cast<const CsStepHeader*>(d_stage);
  a.hdr_stride = 1;
  a.points = reinterpret_cast<const float2*>(d_stage + b->off_points);
  a.points_stride = (size_t)b->max_points;
  a.cand = with_offsets ? reinterpret_cast<const float*>(d_stage + b->off_cand) : nullptr;
  a.cand_stride = (size_t)b->n_cand * 3;
  a.result = d_results;
  a.result_stride = 1;
  a.scan_index = b->update_count;
  a.cand_mode = with_offsets ? CS_CAND_OFFSETS : CS_CAND_PHILOX;
  a.step_mode = CS_STEP_UPDATE;
  a.parity = b->parity;
  a.do_search = do_search ? 1 : 0;
  a.n_cand = b->n_cand;
  a.cand_first = 0;
  a.cand_count = b->n_cand + 1;
  a.s2_host_points = max_n;
  CS_BCUDA(b, launch_step_ctx(batch_ctx(b), a, max_n, rings_hint_of(b->size, b->scale, batch_max_hole_width(b), max_range), CS_PHASE_ALL));
  b->parity ^= 1;
  b->update_count++;
  if (!do_search) b->scan_count++;
  return CS_OK;
}
}  // namespace

cs_status cs_batch_update(cs_batch* b, const float* points, const int32_t* n_points, const float* odometry,
                          const float* cand_offsets, cs_result* results) {
  NvtxRange nvtx_range("cs_batch_update");
  CS_CHECK_BATCH(b);
  if (!points || !n_points || !odometry) return bfail(b, CS_ERR_INVALID_ARGUMENT, "cs_batch_update: null argument");
  // the pinned block may still be the source of the previous call's copy
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  cs_status st = batch_stage_and_launch(b, b->h_stage, b->d_stage, b->d_results, points, n_points, odometry, cand_offsets);
  if (st != CS_OK) return st;
  if (results) {
    CS_BCUDA(b, cudaMemcpyAsync(b->h_results, b->d_results, sizeof(CsDevResult) * (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
    CS_BCUDA(b, cudaStreamSynchronize(b->stream));
    for (int j = 0; j < b->n; j++) {
      copy_result(results + j, b->h_results + j);
      results[j].visits = -1;
    }
  }
  return CS_OK;
}

// The same Update, pipelined: cs_batch_submit stages and queues a step and returns; cs_batch_collect waits for the oldest
// submitted step and hands out its result records.  At most two steps may be waiting for their collect, so a caller that
// alternates submit(k+1), collect(k) has the host staging of step k+1 (pinned copy, range scan, H2D) running while the device
// works on step k, and consecutive steps stay chained on the stream (no host round trip between them).
cs_status cs_batch_submit(cs_batch* b, const float* points, const int32_t* n_points, const float* odometry,
                          const float* cand_offsets) {
  NvtxRange nvtx_range("cs_batch_submit");
  CS_CHECK_BATCH(b);
  if (!points || !n_points || !odometry) return bfail(b, CS_ERR_INVALID_ARGUMENT, "cs_batch_submit: null argument");
  if (b->pipe_submitted - b->pipe_collected >= 2u)
    return bfail(b, CS_ERR_STATE, "cs_batch_submit: two submitted steps are waiting for cs_batch_collect");
  const int slot = (int)(b->pipe_submitted & 1u);
  if (!b->pipe_done[slot]) {  // first use of the slot: all five resources or none (the event, created last, is the guard)
    uint8_t *hs = nullptr, *ds = nullptr;
    CsDevResult *dr = nullptr, *hr = nullptr;
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaHostAlloc(&hs, b->stage_bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&ds, b->stage_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dr, sizeof(CsDevResult) * (size_t)b->n);
    if (e == cudaSuccess) e = cudaHostAlloc(&hr, sizeof(CsDevResult) * (size_t)b->n, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      if (hs) cudaFreeHost(hs);
      if (hr) cudaFreeHost(hr);
      cudaFree(ds);
      cudaFree(dr);
      cudaGetLastError();
      return bfail(b, e == cudaErrorMemoryAllocation ? CS_ERR_OUT_OF_MEMORY : CS_ERR_CUDA, "cs_batch_submit: staging slot: %s",
                   cudaGetErrorString(e));
    }
    b->pipe_h_stage[slot] = hs; b->pipe_d_stage[slot] = ds; b->pipe_d_results[slot] = dr; b->pipe_h_results[slot] = hr;
    b->pipe_done[slot] = ev;
  }
  // The slot's pinned block was the source of the copy two submits ago; that step has been collected (its pipe_done event,
  // recorded after the copy, was waited for), so the block is free.  The device block is protected by stream order.
  cs_status st = batch_stage_and_launch(b, b->pipe_h_stage[slot], b->pipe_d_stage[slot], b->pipe_d_results[slot], points, n_points,
                                        odometry, cand_offsets);
  if (st != CS_OK) return st;
  CS_BCUDA(b, cudaMemcpyAsync(b->pipe_h_results[slot], b->pipe_d_results[slot], sizeof(CsDevResult) * (size_t)b->n,
                              cudaMemcpyDeviceToHost, b->stream));
  CS_BCUDA(b, cudaEventRecord(b->pipe_done[slot], b->stream));
  b->pipe_submitted++;
  return CS_OK;
}

cs_status cs_batch_collect(cs_batch* b, cs_result* results /* n_sessions records, or NULL to only wait */) {
  NvtxRange nvtx_range("cs_batch_collect");
  CS_CHECK_BATCH(b);
  if (b->pipe_collected == b->pipe_submitted) return bfail(b, CS_ERR_STATE, "cs_batch_collect: nothing was submitted");
  const int slot = (int)(b->pipe_collected & 1u);
  CS_BCUDA(b, cudaEventSynchronize(b->pipe_done[slot]));
  if (results) {
    for (int j = 0; j < b->n; j++) {
      copy_result(results + j, b->pipe_h_results[slot] + j);
      results[j].visits = -1;
    }
  }
  b->pipe_collected++;
  return CS_OK;
}

// Parameter sweep: every session consumes the same device-resident scan log (its own map, pose, parameters
// and — when the log carries no candidate tables — its own Philox stream).  results: optional, the
// n_sessions records of the LAST scan replayed.
cs_status cs_batch_replay(cs_batch* b, const cs_scanlog* log, int32_t first, int32_t count, cs_result* results) {
  NvtxRange nvtx_range("cs_batch_replay");
  CS_CHECK_BATCH(b);
  if (!log || !log->uploaded) return bfail(b, CS_ERR_STATE, "scan log not uploaded");
  if (log->device != b->device) return bfail(b, CS_ERR_INVALID_ARGUMENT, "scan log lives on another device");
  if (first < 0 || count < 0 || first + count > log->n_scans) return bfail(b, CS_ERR_INVALID_ARGUMENT, "scan range");
  if (log->n_offsets > 0 && log->n_offsets != b->n_cand) return bfail(b, CS_ERR_INVALID_ARGUMENT, "scan log offsets != T*I");
  if (log->max_points > b->max_points) return bfail(b, CS_ERR_CAPACITY, "scan log max_points exceeds the batch's");
  const float hw = batch_max_hole_width(b);
  for (int i = 0; i < count; i++) {
    const int sidx = first + i;
    const bool do_search = b->scan_count >= b->search_begin;
    CsStepArgs a{};
    a.hdr = log->d_hdr + sidx;
    a.hdr_stride = 0;
    a.points = log->d_points + (size_t)sidx * log->max_points;
    a.points_stride = 0;
    a.cand = log->n_offsets > 0 ? log->d_offsets + (size_t)sidx * log->n_offsets * 3 : nullptr;
    a.cand_stride = 0;
    a.result = b->d_results;
    a.result_stride = 1;
    a.scan_index = b->update_count;
    a.cand_mode = log->n_offsets > 0 ? CS_CAND_OFFSETS : CS_CAND_PHILOX;
    a.step_mode = CS_STEP_UPDATE;
    a.parity = b->parity;
    a.do_search = do_search ? 1 : 0;
    a.n_cand = b->n_cand;
    a.cand_first = 0;
    a.cand_count = b->n_cand + 1;
    a.s2_host_points = log->h_hdr[sidx].n_points;
    CS_BCUDA(b, launch_step_ctx(batch_ctx(b), a, log->h_hdr[sidx].n_points,
                                rings_hint_of(b->size, b->scale, hw, log->h_max_range[sidx]), CS_PHASE_ALL));
    b->parity ^= 1;
    b->update_count++;
    if (!do_search) b->scan_count++;
  }
  if (results && count > 0) {
    CS_BCUDA(b, cudaMemcpyAsync(b->h_results, b->d_results, sizeof(CsDevResult) * (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
    CS_BCUDA(b, cudaStreamSynchronize(b->stream));
    for (int j = 0; j < b->n; j++) {
      copy_result(results + j, b->h_results + j);
      results[j].visits = -1;
    }
  }  // results == NULL: fire and forget (cs_batch_sync waits)
  return CS_OK;
}

cs_status cs_batch_sync(cs_batch* b) {
  CS_CHECK_BATCH(b);
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  return CS_OK;
}

cs_status cs_batch_get_poses(cs_batch* b, float* poses /* n_sessions * 3 */) {
  CS_CHECK_BATCH(b);
  if (!poses) return bfail(b, CS_ERR_INVALID_ARGUMENT, "null poses");
  std::vector<CsSession> tmp((size_t)b->n);
  CS_BCUDA(b, cudaMemcpyAsync(tmp.data(), b->d_sess, sizeof(CsSession) * (size_t)b->n, cudaMemcpyDeviceToHost, b->stream));
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  for (int j = 0; j < b->n; j++)
    for (int k = 0; k < 3; k++) poses[3 * j + k] = tmp[j].state[b->parity].pose[k];
  return CS_OK;
}

cs_status cs_batch_map_download(cs_batch* b, int32_t session, uint16_t* pixels) {
  CS_CHECK_BATCH(b);
  if (!pixels || session < 0 || session >= b->n) return bfail(b, CS_ERR_INVALID_ARGUMENT, "cs_batch_map_download: bad argument");
  if (!b->d_linear) CS_BCUDA(b, cudaMalloc(&b->d_linear, (size_t)b->size * b->size * sizeof(uint16_t)));
  dispatch_layout(b->tiled, [&](auto T) {
    cs_relayout_kernel<decltype(T)::value><<<148 * 8, 256, 0, b->stream>>>(b->hs[session].map, b->d_linear, b->size, b->pitch_tiles, 0);
  });
  b->launches++;
  CS_BCUDA(b, cudaMemcpyAsync(pixels, b->d_linear, (size_t)b->size * b->size * 2, cudaMemcpyDeviceToHost, b->stream));
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  return CS_OK;
}

cs_status cs_batch_map_checksums(cs_batch* b, uint64_t* checksums /* n_sessions */) {
  CS_CHECK_BATCH(b);
  if (!checksums) return bfail(b, CS_ERR_INVALID_ARGUMENT, "null checksums");
  for (int j = 0; j < b->n; j++) {
    CS_BCUDA(b, cudaMemsetAsync(b->d_checksum, 0, sizeof(unsigned long long), b->stream));
    dispatch_layout(b->tiled, [&](auto T) {
      cs_checksum_kernel<decltype(T)::value><<<148 * 4, 256, 0, b->stream>>>(b->hs[j].map, b->size, b->pitch_tiles, b->d_checksum);
    });
    b->launches++;
    CS_BCUDA(b, cudaMemcpyAsync(checksums + j, b->d_checksum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
  }
  CS_BCUDA(b, cudaStreamSynchronize(b->stream));
  return CS_OK;
}

cs_status cs_batch_get_launch_count(cs_batch* b, uint64_t* launches) {
  if (!b || !launches) return CS_ERR_INVALID_ARGUMENT;
  *launches = b->launches;
  return CS_OK;
}

cs_status cs_set_flags(cs_processor* h, uint32_t flags) {
  CS_CHECK_HANDLE(h);
  const uint32_t fixed = CS_FLAG_ROW_MAJOR_MAP | CS_FLAG_L2_PERSIST | CS_FLAG_DEBUG_RAYS | CS_FLAG_SEARCH_WARP | CS_FLAG_SEARCH_SLAB;
  if ((flags & fixed) != (h->cfg.flags & fixed))
    return fail(h, CS_ERR_INVALID_ARGUMENT, "map layout / L2 window / debug rays / search kernel choice are fixed at creation");
  h->cfg.flags = flags;
  int* dist = (flags & CS_FLAG_KEEP_DISTANCES) ? h->d_distances : nullptr;
  h->hs.distances = dist;
  return patch_session(h, offsetof(CsSession, distances), &dist, sizeof(int*));
}

// Random 2-byte gather micro-benchmark: the measured ceiling the search kernel's lookup rate is
// compared with.  Every thread issues `per_thread` independent loads at hashed cell
// indices of a table of `cells` uint16 (L2-resident when it fits), eight in flight at a time.
__global__ void cs_gather_peak_kernel(const uint16_t* __restrict__ table, unsigned mask, int per_thread,
                                      unsigned long long* __restrict__ sink) {
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long z = cs_mix64(tid);
  unsigned acc = 0;
  for (int i = 0; i < per_thread; i += 8) {
    unsigned v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      z = z * 6364136223846793005ull + 1442695040888963407ull;
      v[k] = __ldg(table + ((unsigned)(z >> 33) & mask));
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc += v[k];
  }
  if (acc == 0xffffffffu) atomicAdd(sink, 1ull);
}

int32_t cs_rings_hint(int32_t size_pixels, float size_meters, float hole_width, const float* points, int32_t n_points) {
  if (size_pixels <= 0 || !points || n_points < 0) return 0;
  const float scale = (float)size_pixels / size_meters;  // HoleMap.cs:20
  return rings_hint_of(size_pixels, scale, hole_width, max_range_of(points, n_points));
}

cs_status cs_gather_peak(int32_t device, int64_t cells, int32_t per_thread, int32_t repeats, double* lookups_per_s) {
  if (!lookups_per_s || cells < 1024 || per_thread < 8 || repeats < 1) return CS_ERR_INVALID_ARGUMENT;
  if (cs_device_count() <= device) return fail(nullptr, CS_ERR_NO_DEVICE, "no such CUDA device");
  cudaSetDevice(device);
  unsigned pow2 = 1024;
  while ((int64_t)pow2 * 2 <= cells && pow2 < (1u << 30)) pow2 *= 2;
  uint16_t* table = nullptr;
  unsigned long long* sink = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(&table, (size_t)pow2 * 2) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(table);
    return fail(nullptr, CS_ERR_OUT_OF_MEMORY, "cs_gather_peak: allocation failed");
  }
  cudaMemset(table, 1, (size_t)pow2 * 2);
  cudaMemset(sink, 0, 8);
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256;
  double best = 0;
  for (int r = 0; r < repeats + 2; r++) {
    cudaEventRecord(e0);
    cs_gather_peak_kernel<<<blocks, threads>>>(table, pow2 - 1, per_thread, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * threads * (double)(per_thread / 8 * 8) / (ms * 1e-3);
    if (r >= 2 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(table);
  cudaFree(sink);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, CS_ERR_CUDA, "cs_gather_peak: %s", cudaGetErrorString(e));
  *lookups_per_s = best;
  return CS_OK;
}

// -----------------------------------------------------------------------------------------------------
void cs_philox_offsets(uint64_t seed, uint32_t scan_index, int32_t n, float sigma_xy, float sigma_theta, float* offsets) {
  for (int32_t i = 0; i < n; i++) cs_gauss3(seed, scan_index, (uint32_t)i, sigma_xy, sigma_theta, offsets + 3 * (size_t)i);
}

void cs_host_sincos(const float* angles, int32_t n, float* cos_out, float* sin_out) {
  for (int32_t i = 0; i < n; i++) {
    if (cos_out) cos_out[i] = cs_cosf(angles[i]);
    if (sin_out) sin_out[i] = cs_sinf(angles[i]);
  }
}

float cs_host_normalize_angle(float a) { return cs_normalize_angle(a); }

cs_status cs_device_sincos(int32_t device, const float* angles, int32_t n, float* cos_out, float* sin_out) {
  if (!angles || !cos_out || !sin_out || n <= 0) return CS_ERR_INVALID_ARGUMENT;
  if (cs_device_count() <= device) return fail(nullptr, CS_ERR_NO_DEVICE, "no such CUDA device");
  cudaSetDevice(device);
  float *d_in = nullptr, *d_c = nullptr, *d_s = nullptr;
  size_t bytes = (size_t)n * sizeof(float);
  bool ok = cudaMalloc(&d_in, bytes) == cudaSuccess && cudaMalloc(&d_c, bytes) == cudaSuccess &&
            cudaMalloc(&d_s, bytes) == cudaSuccess && cudaMemcpy(d_in, angles, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    cs_sincos_kernel<<<(n + 255) / 256, 256>>>(d_in, n, d_c, d_s);
    ok = cudaMemcpy(cos_out, d_c, bytes, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(sin_out, d_s, bytes, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(d_in); cudaFree(d_c); cudaFree(d_s);
  if (!ok) return fail(nullptr, CS_ERR_CUDA, "cs_device_sincos: %s", cudaGetErrorString(cudaGetLastError()));
  return CS_OK;
}

}  // extern "C"
