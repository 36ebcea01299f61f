// cs_kernels.cuh — sm_100a kernels of the CoreSLAM scan-to-map hot path.
//
//   cs_search_kernel     CalculateDistanceSISD + MonteCarloSearch + ParallelMonteCarloSearch arg-min
//                        (CoreSLAM/CoreSLAMProcessor.cs:226-259, 624-653, 674-710)
//   cs_finalize_kernel   Update glue (:717-752: gate, searchPose, NormalizeAngle, state) and the per-ray
//                        part of UpdateHoleMap / DrawLaserRayOnHoleMap / ClipRay (:496-534, 359-402, 320-345)
//   cs_integrate_kernel  the draw loop (:404-442) re-organised by rings (see below) so the ordered
//                        read-modify-write is exact without atomics or sorting
//
// No tensor cores: nothing here is a contraction.  The search is a 2-byte gather per (candidate, point)
// out of an L2-resident map; the integration is an ordered 2-byte read-modify-write per visited cell.
// Compile with -fmad=false: every float operation below must round once, like RyuJIT's scalar SSE code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_math.h"

#define CS_TS_NO_OBSTACLE 65500  // CoreSLAMProcessor.cs:21
#define CS_TS_OBSTACLE 0         // CoreSLAMProcessor.cs:22

enum CsCandMode { CS_CAND_PHILOX = 0, CS_CAND_OFFSETS = 1, CS_CAND_ABSOLUTE = 2 };
enum CsStepMode { CS_STEP_UPDATE = 0, CS_STEP_SEARCH_ONLY = 1, CS_STEP_INTEGRATE_ONLY = 2 };

struct CsState {  // CoreSLAMProcessor.cs:34-35, 106
  float pose[3];
  float last_odo[3];
  int scan_count;
  int pad;
};

struct CsStepHeader {  // per-scan input, 48 B
  float odo[3];        // odometry pose (UPDATE) or explicit pose (SEARCH_ONLY / INTEGRATE_ONLY)
  int n_points;
  float cs[2];         // optional host (cos, sin) of the explicit pose
  int has_cs;
  int pad[5];
};

struct CsRay {  // 32 B: everything the draw loop needs to know about one ray, in closed form
  int dxc, dyc;   // clipped extents along the major / minor axis (:370-371 after the :383-385 swap)
  int a0, b0;     // pixval descends for major steps a0..b0 (:406-416); ascends from max(a0, b0+1) (:418-427)
  int incv;       // :398
  int kc;         // how many of the ascending steps take the +1 carry (:422-426)
  int nd_total;   // number of descending steps once past b0
  int flags;      // bit0 valid, bit1 steep (major axis = y), bit2 major step negative, bit3 minor step negative
};

struct CsDevResult {  // device twin of cs_result (include/coreslam_b200.h)
  float pose[3];
  int distance;
  int index;
  int searched;
  long long visits;
};

struct CsSession {  // one CoreSLAMProcessor, device resident
  uint16_t* map;    // HoleMap.Pixels; tiled: 8x8-cell tiles of 128 B, 4x4-cell quadrants of 32 B
  int size;         // HoleMap.Size
  int pitch_tiles;  // tiles per tile-row (tiled layout)
  float scale;      // HoleMap.Scale
  float sigma_xy, sigma_theta;
  int iters, threads, n_cand;  // n_cand = max(threads,1) * iters random candidates (+1 for searchPose)
  int quality;       // :82
  float hole_width;  // :87
  int search_begin;  // :92
  unsigned long long seed;
  CsState state[2];            // slot 0 is live (slot 1 spare)
  unsigned long long key[2];   // packed (distance << 32 | flat index) arg-min of the running search
  CsRay* rays;                 // capacity max_points
  int* ray_dbg;                // optional 6 ints per ray (x1,y1,x2,y2,xp,yp)
  int* distances;              // optional n_cand+1
  long long* ring_cycles;      // optional diagnostics: cycles each ring's warp spent in the integrate kernel
  int x1, y1;                  // ray origin cell of the current integration (:505-506)
  int max_ring;                // max dxc over valid rays, -1 if nothing to draw
  int n_rays;
  long long visits;
};

struct CsStepArgs {  // by-value kernel argument; session j uses element j of every array
  const CsStepHeader* hdr;
  const float2* points;  size_t points_stride;   // stride in float2 between sessions
  const float* cand;     size_t cand_stride;     // offsets or absolute poses, 3 floats per candidate
  const float* cand_cs;                          // optional (n_cand+1)*2 host cos/sin
  CsDevResult* result;   size_t result_stride;   // where finalize writes (device or mapped host memory)
  volatile unsigned* seq_flag;                   // optional mapped-host flag, set to seq_value when the pose is out
  unsigned seq_value;
  unsigned scan_index;
  int cand_mode;   // CsCandMode
  int step_mode;   // CsStepMode
  int parity;      // which state/key slot this handle uses (0)
  int do_search;   // host mirror of scanCount >= PositionSearchBeginning (:726)
  int n_cand;      // random candidates evaluated this step (excludes searchPose)
  int cand_first;  // first flat index evaluated by this GPU (multi-GPU candidate split), normally 0
  int cand_count;  // number of flat indices evaluated by this GPU, normally n_cand+1
};

// ---------------------------------------------------------------------------------------------------
// map addressing
// ---------------------------------------------------------------------------------------------------
template <bool TILED>
__device__ __forceinline__ uint32_t cs_cell_offset(int x, int y, int size, int pitch_tiles) {
  if (TILED) {
    uint32_t tile = (uint32_t)(y >> 3) * (uint32_t)pitch_tiles + (uint32_t)(x >> 3);
    uint32_t in = (uint32_t)(x & 3) | ((uint32_t)(y & 3) << 2) | ((uint32_t)(x & 4) << 2) | ((uint32_t)(y & 4) << 3);
    return tile * 64u + in;
  } else {
    return (uint32_t)y * (uint32_t)size + (uint32_t)x;
  }
}

// wrapping int32 arithmetic (C# unchecked)
__device__ __forceinline__ int cs_wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int cs_wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int cs_wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int cs_wabs(int a) { return a < 0 ? cs_wsub(0, a) : a; }
__device__ __forceinline__ int cs_sign(int a) { return (a > 0) - (a < 0); }
__device__ __forceinline__ int cs_wdiv(int a, int b) { return (b == 0 || (a == (int)0x80000000 && b == -1)) ? 0 : a / b; }

// ---------------------------------------------------------------------------------------------------
// candidate enumeration: flat index 0 = searchPose, 1 + t*I + i = thread t's i-th candidate
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cs_search_pose(const CsSession& S, const CsStepHeader& h, const CsStepArgs& a,
                                               float sp[3]) {
  if (a.step_mode == CS_STEP_UPDATE) {
    const CsState& st = S.state[a.parity];
#pragma unroll
    for (int k = 0; k < 3; k++) sp[k] = __fadd_rn(st.pose[k], __fsub_rn(h.odo[k], st.last_odo[k]));  // :728
  } else {
    sp[0] = h.odo[0]; sp[1] = h.odo[1]; sp[2] = h.odo[2];
  }
}

__device__ __forceinline__ void cs_candidate_pose(const CsSession& S, const CsStepArgs& a, const float* cand,
                                                  const float sp[3], int idx, float pose[3]) {
  if (idx == 0) {
    pose[0] = sp[0]; pose[1] = sp[1]; pose[2] = sp[2];
    return;
  }
  if (a.cand_mode == CS_CAND_ABSOLUTE) {
    const float* p = cand + 3 * (size_t)(idx - 1);
    pose[0] = p[0]; pose[1] = p[1]; pose[2] = p[2];
    return;
  }
  float off[3];
  if (a.cand_mode == CS_CAND_OFFSETS) {
    const float* p = cand + 3 * (size_t)(idx - 1);
    off[0] = p[0]; off[1] = p[1]; off[2] = p[2];
  } else {
    cs_gauss3(S.seed, a.scan_index, (uint32_t)(idx - 1), S.sigma_xy, S.sigma_theta, off);
  }
  pose[0] = __fadd_rn(sp[0], off[0]);  // :635-637
  pose[1] = __fadd_rn(sp[1], off[1]);
  pose[2] = __fadd_rn(sp[2], off[2]);
}

// ---------------------------------------------------------------------------------------------------
// search: one warp per candidate pose, lanes stride over the scan in 32-point strips
// ---------------------------------------------------------------------------------------------------
#define CS_SEARCH_WARPS 8
#define CS_SEARCH_CHUNK 2048  // points staged in shared memory per pass (16 KB)

template <bool TILED>
__global__ void __launch_bounds__(CS_SEARCH_WARPS * 32)
cs_search_kernel(CsSession* __restrict__ sessions, CsStepArgs a) {
  __shared__ float2 s_pts[CS_SEARCH_CHUNK];
  __shared__ unsigned long long s_key[CS_SEARCH_WARPS];

  const int sj = blockIdx.y;
  CsSession& S = sessions[sj];
  const CsStepHeader& hdr = a.hdr[sj];
  const float2* __restrict__ points = a.points + (size_t)sj * a.points_stride;
  const float* cand = a.cand ? a.cand + (size_t)sj * a.cand_stride : nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int local = blockIdx.x * CS_SEARCH_WARPS + warp;
  const int idx = a.cand_first + local;  // flat candidate index
  const bool valid = local < a.cand_count;
  const int P = hdr.n_points;
  const int size = S.size, pitch_tiles = S.pitch_tiles;
  const uint16_t* __restrict__ map = S.map;
  const float scale = S.scale;

  float px = 0.f, py = 0.f, c = 0.f, s = 0.f;
  if (valid) {
    float sp[3], pose[3];
    cs_search_pose(S, hdr, a, sp);
    cs_candidate_pose(S, a, cand, sp, idx, pose);
    float ct, st;
    if (a.cand_cs) {
      ct = a.cand_cs[2 * (size_t)idx];
      st = a.cand_cs[2 * (size_t)idx + 1];
    } else {
      ct = cs_cosf(pose[2]);
      st = cs_sinf(pose[2]);
    }
    px = __fadd_rn(__fmul_rn(pose[0], scale), 0.5f);  // :232
    py = __fadd_rn(__fmul_rn(pose[1], scale), 0.5f);  // :233
    c = __fmul_rn(ct, scale);                         // :234
    s = __fmul_rn(st, scale);                         // :235
  }

  unsigned sum = 0;  // <= 65535 * 65536 < 2^32 (max_points <= 65536)
  unsigned nb = 0;
  for (int base = 0; base < P; base += CS_SEARCH_CHUNK) {
    const int n = min(CS_SEARCH_CHUNK, P - base);
    __syncthreads();
    // vectorised staging: two points per 16-byte load
    const float4* src4 = reinterpret_cast<const float4*>(points + base);
    float4* dst4 = reinterpret_cast<float4*>(s_pts);
    for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) dst4[i] = __ldg(src4 + i);
    if ((n & 1) && threadIdx.x == 0) s_pts[n - 1] = points[base + n - 1];
    __syncthreads();
    if (valid) {
#pragma unroll 8
      for (int i = lane; i < n; i += 32) {
        const float2 p = s_pts[i];
        // :240-241, left-associative, one rounding per operation
        float fx = __fsub_rn(__fadd_rn(px, __fmul_rn(c, p.x)), __fmul_rn(s, p.y));
        float fy = __fadd_rn(__fadd_rn(py, __fmul_rn(s, p.x)), __fmul_rn(c, p.y));
        // (int) cast + bounds test (:244).  fmaxf drops NaN to -2 (out of bounds, like cvttss2si's
        // 0x80000000); +overflow saturates to INT_MAX, also out of bounds.
        int x = __float2int_rz(fmaxf(fx, -2.0f));
        int y = __float2int_rz(fmaxf(fy, -2.0f));
        bool in = ((unsigned)x < (unsigned)size) && ((unsigned)y < (unsigned)size);
        if (in) {
          sum += (unsigned)__ldg(map + cs_cell_offset<TILED>(x, y, size, pitch_tiles));  // :246
          nb++;
        }
      }
    }
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }

  unsigned long long key = ~0ull;
  if (valid) {
    int d = nb > 0 ? (int)(((unsigned long long)sum * 1024ull) / (unsigned long long)P) : 2147483647;  // :251-258
    key = ((unsigned long long)(unsigned)d << 32) | (unsigned)idx;
    if (lane == 0 && S.distances) S.distances[idx] = d;
  }
  if (lane == 0) s_key[warp] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long k = s_key[0];
#pragma unroll
    for (int w = 1; w < CS_SEARCH_WARPS; w++) k = min(k, s_key[w]);
    if (k != ~0ull) atomicMin(&S.key[a.parity], k);
  }
}

// ---------------------------------------------------------------------------------------------------
// ray set-up (ClipRay :320-345 and the prologue of DrawLaserRayOnHoleMap :361-402)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cs_clip_ray(int size, int& xyc, int& yxc, int xy, int yx) {
  if (xyc < 0) {
    if (xyc == xy) return false;
    yxc = cs_wadd(yxc, cs_wdiv(cs_wmul(cs_wsub(yxc, yx), cs_wsub(0, xyc)), cs_wsub(xyc, xy)));
    xyc = 0;
  }
  if (xyc >= size) {
    if (xyc == xy) return false;
    yxc = cs_wadd(yxc, cs_wdiv(cs_wmul(cs_wsub(yxc, yx), cs_wsub(size - 1, xyc)), cs_wsub(xyc, xy)));
    xyc = size - 1;
  }
  return true;
}

__device__ __forceinline__ CsRay cs_make_ray(int size, int x1, int y1, int x2, int y2, int xp, int yp) {
  CsRay r;
  r.dxc = 0; r.dyc = 0; r.a0 = 0; r.b0 = -1; r.incv = 0; r.kc = 0; r.nd_total = 0; r.flags = 0;
  int x2c = x2, y2c = y2;
  if (!cs_clip_ray(size, x2c, y2c, x1, y1)) return r;  // :365
  if (!cs_clip_ray(size, y2c, x2c, y1, x1)) return r;  // :366
  int dx = cs_wabs(cs_wsub(x2, x1)), dy = cs_wabs(cs_wsub(y2, y1));        // :368-369
  int dxc = cs_wabs(cs_wsub(x2c, x1)), dyc = cs_wabs(cs_wsub(y2c, y1));    // :370-371
  int sx = cs_sign(cs_wsub(x2, x1)), sy = cs_sign(cs_wsub(y2, y1));        // :372-373
  int D, smaj, smin;
  bool steep;
  if (dx > dy) {  // :377
    steep = false;
    D = cs_wabs(cs_wsub(xp, x2));
    smaj = sx; smin = sy;
  } else {
    steep = true;
    dx = dy;
    int t = dxc; dxc = dyc; dyc = t;
    D = cs_wabs(cs_wsub(yp, y2));
    smaj = sy; smin = sx;
  }
  if (D == 0) return r;  // :389-392
  // the walk below never leaves the box [start, clipped end]; reject anything a wrapped clip produced
  if (dxc >= size || dyc >= size || D < 0) return r;
  const int value = CS_TS_OBSTACLE;
  int incv = (value - CS_TS_NO_OBSTACLE) / D;                     // :398
  int rem = -(value - CS_TS_NO_OBSTACLE - cs_wmul(D, incv));      // -incerrorv >= 0, :399
  // zone thresholds in 64 bit: descending while t2 < x <= t1, ascending while x > t2 and x > t1 (:406-408)
  long long t2 = (long long)cs_wsub(dx, cs_wmul(2, D)) + 1;       // first x with x > dx - 2*derrorv
  long long t1 = (long long)cs_wsub(dx, D);
  long long a0 = t2 > 0 ? t2 : 0;
  long long lim = (long long)size + 1;                            // beyond any x <= dxc
  if (a0 > lim) a0 = lim;
  if (t1 > lim) t1 = lim;
  if (t1 < -1) t1 = -1;
  long long nd_total = t1 - a0 + 1;
  if (nd_total < 0) nd_total = 0;
  long long e0 = (long long)(D / 2) - nd_total * (long long)rem;  // errorv entering the ascending zone (:397, :411)
  long long need = -e0 - (long long)rem;
  long long kc = need > 0 ? (need + (long long)rem + (long long)D - 1) / ((long long)rem + (long long)D) : 0;
  if (kc > lim) kc = lim;
  r.dxc = dxc; r.dyc = dyc;
  r.a0 = (int)a0; r.b0 = (int)t1;
  r.incv = incv; r.kc = (int)kc; r.nd_total = (int)nd_total;
  r.flags = 1 | (steep ? 2 : 0) | (smaj < 0 ? 4 : 0) | (smin < 0 ? 8 : 0);
  return r;
}

// pixval written at major step x of ray r (closed form of :402-428)
__device__ __forceinline__ int cs_ray_pixval(const CsRay& r, int x) {
  if (x <= r.b0) {
    int nd = x - r.a0 + 1;
    nd = nd > 0 ? nd : 0;
    return CS_TS_NO_OBSTACLE + nd * r.incv;
  }
  int c0 = max(r.a0, r.b0 + 1);
  if (x < c0) return CS_TS_NO_OBSTACLE;
  int j = x - c0 + 1;
  return CS_TS_NO_OBSTACLE + (r.nd_total - j) * r.incv + min(j, r.kc);
}

// minor-axis offset after x major steps (closed form of the Bresenham error walk :394-396, 433-441)
__device__ __forceinline__ int cs_ray_minor(const CsRay& r, int x) {
  if (x == 0) return 0;
  unsigned num = 2u * (unsigned)r.dyc * (unsigned)x + (unsigned)r.dxc - 1u;
  unsigned m = num / (2u * (unsigned)r.dxc);
  return min((int)m, x);
}

// ---------------------------------------------------------------------------------------------------
// finalize: one block per session
// ---------------------------------------------------------------------------------------------------
#define CS_FINALIZE_THREADS 512

__global__ void __launch_bounds__(CS_FINALIZE_THREADS)
cs_finalize_kernel(CsSession* __restrict__ sessions, CsStepArgs a) {
  __shared__ float s_pose[3];
  __shared__ float s_cs[2];
  __shared__ int s_red_ring[CS_FINALIZE_THREADS / 32];
  __shared__ long long s_red_vis[CS_FINALIZE_THREADS / 32];

  const int sj = blockIdx.x;
  CsSession& S = sessions[sj];
  const CsStepHeader& hdr = a.hdr[sj];
  const float2* __restrict__ points = a.points + (size_t)sj * a.points_stride;
  const float* cand = a.cand ? a.cand + (size_t)sj * a.cand_stride : nullptr;
  CsDevResult* result = a.result ? a.result + (size_t)sj * a.result_stride : nullptr;

  if (threadIdx.x == 0) {
    float pose[3];
    int dist = 2147483647, index = 0, searched = 0;
    bool have_cs = false;
    float ct = 0.f, st = 0.f;
    if (a.step_mode == CS_STEP_INTEGRATE_ONLY) {
      pose[0] = hdr.odo[0]; pose[1] = hdr.odo[1]; pose[2] = hdr.odo[2];
      if (hdr.has_cs) { have_cs = true; ct = hdr.cs[0]; st = hdr.cs[1]; }
    } else {
      float sp[3];
      cs_search_pose(S, hdr, a, sp);
      if (a.do_search) {
        unsigned long long key = S.key[a.parity];
        dist = (int)(unsigned)(key >> 32);
        index = (int)(unsigned)(key & 0xffffffffu);
        searched = 1;
        cs_candidate_pose(S, a, cand, sp, index, pose);
      } else {
        pose[0] = hdr.odo[0]; pose[1] = hdr.odo[1]; pose[2] = hdr.odo[2];  // :742
      }
      if (a.step_mode == CS_STEP_UPDATE) {
        pose[2] = cs_normalize_angle(pose[2]);  // :746
        // in place: kernels of one handle are stream ordered, and only this thread touches the state here
        CsState& st1 = S.state[a.parity];
        st1.pose[0] = pose[0]; st1.pose[1] = pose[1]; st1.pose[2] = pose[2];  // :747
        st1.last_odo[0] = hdr.odo[0]; st1.last_odo[1] = hdr.odo[1]; st1.last_odo[2] = hdr.odo[2];  // :745
        st1.scan_count += (a.do_search ? 0 : 1);  // :741
      }
    }
    S.key[a.parity] = ~0ull;  // re-arm the arg-min for the next search
    if (result) {
      result->pose[0] = pose[0]; result->pose[1] = pose[1]; result->pose[2] = pose[2];
      result->distance = dist;
      result->index = index;
      result->searched = searched;
      if (a.seq_flag) {
        __threadfence_system();
        *a.seq_flag = a.seq_value;
      }
    }
    if (!have_cs) { ct = cs_cosf(pose[2]); st = cs_sinf(pose[2]); }
    s_pose[0] = pose[0]; s_pose[1] = pose[1]; s_pose[2] = pose[2];
    s_cs[0] = ct; s_cs[1] = st;
  }
  __syncthreads();

  int max_ring = -1;
  long long visits = 0;
  const int n = (a.step_mode == CS_STEP_SEARCH_ONLY) ? 0 : hdr.n_points;
  const float scale = S.scale;
  const int size = S.size;
  const float px = __fadd_rn(__fmul_rn(s_pose[0], scale), 0.5f);  // :499
  const float py = __fadd_rn(__fmul_rn(s_pose[1], scale), 0.5f);  // :500
  const float c = __fmul_rn(s_cs[0], scale);                       // :501
  const float s = __fmul_rn(s_cs[1], scale);                       // :502
  const int x1 = cs_cvt_i32(px), y1 = cs_cvt_i32(py);              // :505-506
  const bool on_map = !(x1 < 0 || x1 >= size || y1 < 0 || y1 >= size);  // :509-512
  const float hw = S.hole_width;

  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    CsRay r;
    r.dxc = 0; r.dyc = 0; r.a0 = 0; r.b0 = -1; r.incv = 0; r.kc = 0; r.nd_total = 0; r.flags = 0;
    int x2 = 0, y2 = 0, xp = 0, yp = 0;
    if (on_map) {
      const float2 p = points[i];
      float x2p = __fsub_rn(__fmul_rn(c, p.x), __fmul_rn(s, p.y));  // :519
      float y2p = __fadd_rn(__fmul_rn(s, p.x), __fmul_rn(c, p.y));  // :520
      xp = cs_cvt_i32(__fadd_rn(px, x2p));                          // :521
      yp = cs_cvt_i32(__fadd_rn(py, y2p));                          // :522
      float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(x2p, x2p), __fmul_rn(y2p, y2p)));  // :524
      float add = __fdiv_rn(__fdiv_rn(__fmul_rn(hw, scale), 2.0f), dist);            // :525
      float k1 = __fadd_rn(1.0f, add);
      x2p = __fmul_rn(x2p, k1);                                     // :527
      y2p = __fmul_rn(y2p, k1);                                     // :528
      x2 = cs_cvt_i32(__fadd_rn(px, x2p));                          // :529
      y2 = cs_cvt_i32(__fadd_rn(py, y2p));                          // :530
      r = cs_make_ray(size, x1, y1, x2, y2, xp, yp);
      if (r.flags & 1) {
        max_ring = max(max_ring, r.dxc);
        visits += (long long)r.dxc + 1;
      }
    }
    S.rays[i] = r;
    if (S.ray_dbg) {
      int* d = S.ray_dbg + 6 * (size_t)i;
      d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2; d[4] = xp; d[5] = yp;
    }
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    max_ring = max(max_ring, __shfl_xor_sync(0xffffffffu, max_ring, o));
    visits += __shfl_xor_sync(0xffffffffu, visits, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red_ring[threadIdx.x >> 5] = max_ring;
    s_red_vis[threadIdx.x >> 5] = visits;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < CS_FINALIZE_THREADS / 32; w++) {
      max_ring = max(max_ring, s_red_ring[w]);
      visits += s_red_vis[w];
    }
    S.x1 = x1; S.y1 = y1;
    S.max_ring = max_ring;
    S.n_rays = n;
    S.visits = visits;
    if (result) result->visits = visits;
  }
}

// ---------------------------------------------------------------------------------------------------
// integration by rings.
//
// Every ray starts in the same cell (x1,y1) and advances exactly one cell along its major axis per step,
// so the cell written at step k lies on the square ring of Chebyshev radius k around the start.  Rings
// are therefore independent of each other, and inside a ring the only ordering that matters is the
// reference's ray order (foreach over cloud.Points, :517).  One warp owns one ring: it walks the rays
// in index order 32 at a time, each lane evaluates its ray's cell and pixval at step k in closed form,
// lanes that hit the same cell are applied in lane (= ray) order by the group's first lane, and
// consecutive 32-ray batches are ordered by program order.  No atomics, no sort, bit-exact.
// ---------------------------------------------------------------------------------------------------
#define CS_INT_WARPS 8
#define CS_INT_CHUNK 256  // rays staged per pass (8 KB)

__device__ __forceinline__ int cs_blend(int old, int pixval, int alpha) {
  return (int)(uint16_t)(((256 - alpha) * old + alpha * pixval) >> 8);  // :431
}

template <bool TILED>
__global__ void __launch_bounds__(CS_INT_WARPS * 32)
cs_integrate_kernel(CsSession* __restrict__ sessions) {
  __shared__ CsRay s_rays[CS_INT_CHUNK];
  const int sj = blockIdx.y;
  CsSession& S = sessions[sj];
  const int max_ring = S.max_ring;
  if ((int)(blockIdx.x * CS_INT_WARPS) > max_ring) return;  // whole block beyond the longest ray

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * CS_INT_WARPS + warp;  // this warp's ring
  const int n = S.n_rays;
  const int size = S.size, pitch_tiles = S.pitch_tiles;
  const int x1 = S.x1, y1 = S.y1;
  const int alpha = S.quality;
  uint16_t* __restrict__ map = S.map;
  const CsRay* __restrict__ rays = S.rays;
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long t_begin = S.ring_cycles ? clock64() : 0;

  for (int base = 0; base < n; base += CS_INT_CHUNK) {
    const int cn = min(CS_INT_CHUNK, n - base);
    __syncthreads();
    {
      // 256 rays * 32 B, two 16-byte loads per thread
      const int4* src = reinterpret_cast<const int4*>(rays + base);
      int4* dst = reinterpret_cast<int4*>(s_rays);
      for (int i = threadIdx.x; i < cn * 2; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (k > max_ring) continue;

    for (int b = 0; b < cn; b += 32) {
      const int ri = b + lane;
      bool active = false;
      uint32_t cell = 0;
      int pixval = 0;
      if (ri < cn) {
        const CsRay r = s_rays[ri];
        if ((r.flags & 1) && k <= r.dxc) {
          int m = cs_ray_minor(r, k);
          int dmaj = (r.flags & 4) ? -k : k;
          int dmin = (r.flags & 8) ? -m : m;
          int x = (r.flags & 2) ? x1 + dmin : x1 + dmaj;
          int y = (r.flags & 2) ? y1 + dmaj : y1 + dmin;
          if ((unsigned)x < (unsigned)size && (unsigned)y < (unsigned)size) {
            active = true;
            cell = cs_cell_offset<TILED>(x, y, size, pitch_tiles);
            pixval = cs_ray_pixval(r, k);
          }
        }
      }
      const unsigned am = __ballot_sync(0xffffffffu, active);
      if (am == 0) continue;
      unsigned peers = 0;
      if (active) peers = __match_any_sync(am, cell);
      const bool leader = active && ((peers & lt_mask) == 0);
      const int cnt = __popc(peers);
      const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
      int v = 0;
      if (leader) v = (int)__ldcg(map + cell);
      if (maxcnt == 1) {
        if (leader) v = cs_blend(v, pixval, alpha);
      } else {
        // ordered application inside each same-cell group: the leader pulls its peers' pixvals in
        // lane order (= ray order)
        unsigned rest = peers;
        for (int t = 0; t < maxcnt; t++) {
          int src = lane;
          if (leader && rest) {
            src = __ffs(rest) - 1;
            rest &= rest - 1;
          } else if (leader) {
            src = -1;
          }
          int pv = __shfl_sync(0xffffffffu, pixval, src < 0 ? lane : src);
          if (leader && src >= 0) v = cs_blend(v, pv, alpha);
        }
      }
      if (leader) __stcg(map + cell, (uint16_t)v);
      __syncwarp();
    }
  }
  if (S.ring_cycles && lane == 0 && k <= max_ring) S.ring_cycles[k] = clock64() - t_begin;
}

// ---------------------------------------------------------------------------------------------------
// map helpers
// ---------------------------------------------------------------------------------------------------
__global__ void cs_fill_kernel(uint16_t* __restrict__ map, size_t n_cells, uint16_t value) {
  // 8 cells per 16-byte store
  const uint32_t v2 = (uint32_t)value | ((uint32_t)value << 16);
  const uint4 v = make_uint4(v2, v2, v2, v2);
  size_t n16 = n_cells / 8;
  uint4* p = reinterpret_cast<uint4*>(map);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n16 * 8; i < n_cells; i++) map[i] = value;
}

// row-major <-> device layout; to_device != 0: linear -> map
template <bool TILED>
__global__ void cs_relayout_kernel(uint16_t* __restrict__ map, uint16_t* __restrict__ linear, int size, int pitch_tiles,
                                   int to_device) {
  const size_t n = (size_t)size * (size_t)size;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int y = (int)(i / (size_t)size), x = (int)(i % (size_t)size);
    uint32_t o = cs_cell_offset<TILED>(x, y, size, pitch_tiles);
    if (to_device) map[o] = linear[i];
    else linear[i] = map[o];
  }
}

__device__ __host__ __forceinline__ unsigned long long cs_mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// checksum = sum over cells of (value+1) * mix64(row-major index), mod 2^64 — order independent
template <bool TILED>
__global__ void cs_checksum_kernel(const uint16_t* __restrict__ map, int size, int pitch_tiles,
                                   unsigned long long* __restrict__ out) {
  const size_t n = (size_t)size * (size_t)size;
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int y = (int)(i / (size_t)size), x = (int)(i % (size_t)size);
    unsigned long long v = map[cs_cell_offset<TILED>(x, y, size, pitch_tiles)];
    acc += (v + 1ull) * cs_mix64((unsigned long long)i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// HoleMap.GetPackedPixels (HoleMap.cs:44-55): two cells per byte, top 4 bits each
template <bool TILED>
__global__ void cs_pack_kernel(const uint16_t* __restrict__ map, int size, int pitch_tiles, uint8_t* __restrict__ out) {
  const size_t n = ((size_t)size * (size_t)size) / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    size_t i0 = 2 * i, i1 = 2 * i + 1;
    unsigned a = map[cs_cell_offset<TILED>((int)(i0 % size), (int)(i0 / size), size, pitch_tiles)];
    unsigned b = map[cs_cell_offset<TILED>((int)(i1 % size), (int)(i1 / size), size, pitch_tiles)];
    out[i] = (uint8_t)(((a >> 12) << 4) | (b >> 12));
  }
}

__global__ void cs_sincos_kernel(const float* __restrict__ in, int n, float* __restrict__ c, float* __restrict__ s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    c[i] = cs_cosf(in[i]);
    s[i] = cs_sinf(in[i]);
  }
}
