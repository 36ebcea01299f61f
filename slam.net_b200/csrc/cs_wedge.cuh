// cs_wedge.cuh — the draw loop of DrawLaserRayOnHoleMap (CoreSLAM/CoreSLAMProcessor.cs:404-442) as independent
// (ring range x angular wedge) tasks, one warp per task, no block-wide synchronisation in the draw loop.
//
// Why it is exact.  Every ray of a scan starts in the same cell (x1,y1) and advances one cell along its major axis per
// step, so its step k lies on the square ring of Chebyshev radius k; rings never share a cell.  On ring k ray r sits at
// position  p = c*k + g*m(k)  along the ring (c in {1,3,5,7} the side, g = +-1), and the Bresenham walk :394-396, 433-441
// has the closed form  m(k) = min(k, ceil(k*s - 1/2)),  s = dyc/dxc  — so  |p - k*kappa| <= 1/2  with the ray's angular
// key  kappa = c + g*s  in [0, 8].  A wedge is a key interval [beta_w, beta_w+1); on ring k it OWNS the positions
// ceil(k*beta_w) <= p < ceil(k*beta_w+1)  (every position of every ring belongs to exactly one wedge, whatever the
// boundaries are), and only rays with  beta_w - 1/(2k) <= kappa < beta_w+1 + 1/(2k)  can land there.  A task therefore
// filters its candidate rays by key (conservatively, in float), keeps them in RAY ORDER, decides ownership of every visit
// exactly (integers), and applies the visits of one cell in ray order: all visitors of a cell are candidates of the one
// task that owns the cell, so the reference's order (foreach over cloud.Points, :517) is kept without atomics, sorting or
// a slot table, for any ray order.  tools/wedge_model.py is the host model of this decomposition.
//
// Two paths per task, chosen by the candidate count:
//   fast     <= 32 candidates: one ray per lane, Bresenham state carried in registers from ring to ring (no per-ring
//            division), visits of one cell inside the warp found with match.any and applied by the lowest lane in lane
//            (= ray) order; rings are processed four at a time so four map loads per lane are in flight.
//   general  any number of candidates (the dense centre, where every ray crosses every wedge; clustered or multi-turn
//            scans): per ring the wedge's cells are staged in shared memory, the candidate rays are evaluated in closed
//            form 32 at a time in ray order, same-cell groups fold into the staged value (uniform groups as a count
//            with the exact fixed-point early-out), then the touched cells go back.
// The first blocks prepare the rays exactly as the rings kernel's do (cs_ray_from_point), plus each ray's key, the key
// range of every 32 rays, and per level the number of rays that reach it (which sizes the wedges: ~26 candidates each).
#pragma once
#include "cs_kernels.cuh"

#define CS_W_THREADS 256
#define CS_W_WARPS (CS_W_THREADS / 32)
#define CS_W_LEVELS 264   // rings [1,1] [2,3] [4,7] [8,15] [16,31] [32,63], then 64 rings each: ring 16383 is level 260
#define CS_W_FIX 14       // wedge boundaries are multiples of 2^-14 (keys span [0, 8])
#define CS_W_CAP 256      // cells of a wedge staged per pass of the general path
#define CS_W_G 4          // rings in flight per lane on the fast path
#define CS_W_MAX_WEDGES 8192

__device__ __forceinline__ int cs_w_level_first(int L) { return L < 6 ? (1 << L) : 64 * (L - 5); }
__device__ __forceinline__ int cs_w_level_last(int L) { return L < 6 ? (2 << L) - 1 : 64 * (L - 5) + 63; }
__device__ __forceinline__ int cs_w_level_of(int k) { return k < 64 ? 31 - __clz(k) : 5 + (k >> 6); }  // k >= 1

// wedges of a level reached by `alive` rays whose first ring is k0: about 26 candidates per wedge, of which
// alive / (8 k0) come from the +-1/(2 k0) margins; where the margins alone exceed that (the centre) a wedge is as wide
// as its margins.  Never more wedges than ring k0 has cells.
__device__ __forceinline__ int cs_w_wedges(int alive, int k0) {
  if (alive <= 0) return 0;
  const int halo2 = alive / (8 * k0);
  const int own = halo2 > 13 ? halo2 : 26 - halo2;
  int W = (alive + own - 1) / own;
  W = min(W, min(8 * k0, CS_W_MAX_WEDGES));
  return max(W, 1);
}
__device__ __forceinline__ unsigned cs_w_beta(int w, int W) { return (unsigned)(((unsigned)w * (8u << CS_W_FIX)) / (unsigned)W); }
__device__ __forceinline__ int cs_w_bound(int k, unsigned beta) { return (int)(((unsigned)k * beta + ((1u << CS_W_FIX) - 1u)) >> CS_W_FIX); }

// side constant c and minor sign g of  p = c*k + g*m  (cs_ring_visit's position, before the 8k -> 0 wrap)
__device__ __forceinline__ void cs_w_side(int flags, int& c, int& g) {
  const bool steep = (flags & 2) != 0, majneg = (flags & 4) != 0, minneg = (flags & 8) != 0;
  c = steep ? (majneg ? 7 : 3) : (majneg ? 5 : 1);
  g = ((steep != majneg) ? -1 : 1) * (minneg ? -1 : 1);
}

struct CsWTask {
  int k0, k1;          // rings, inclusive
  unsigned blo, bhi;   // wedge boundaries (fixed point)
  float flo, fhi;      // key range a candidate must fall in (margins included)
  float wrap_lo;       // wedge 0 also takes keys >= wrap_lo (positions 8k wrap to 0); 9 = no
};

// ---- ray preparation by the first blocks (UpdateHoleMap :517-530, ClipRay, the prologue of DrawLaserRayOnHoleMap), one ray
// per lane: packed ray, (key, dxc), the warp's key range and largest dxc, and the per-level counts.
__device__ __forceinline__ void cs_w_prepare(CsSession& S, const CsRayFrame& f, const float2 p, int i, bool in_range, int* alive,
                                             long long& visits) {
  const unsigned full = 0xffffffffu;
  CsRay r;
  r.dxc = 0; r.dyc = 0; r.a0 = 0; r.b0 = -1; r.incv = 0; r.kc = 0; r.nd_total = 0; r.flags = 0;
  if (in_range) r = cs_ray_from_point(f, p, S.ray_dbg ? S.ray_dbg + 6 * (size_t)i : nullptr);
  const bool valid = (r.flags & 1) != 0;
  float key = 0.f;
  if (valid) {
    visits += (long long)r.dxc + 1;
    int c, g;
    cs_w_side(r.flags, c, g);
    const float s = r.dxc > 0 ? __fdiv_rn((float)min(r.dyc, r.dxc), (float)r.dxc) : 0.f;
    key = (float)c + (float)g * s;
  }
  if (in_range) {
    S.rays[i] = cs_pack_ray(r);
    S.w_rk[i] = make_int2(__float_as_int(key), valid ? r.dxc : -1);
  }
  const int lane = threadIdx.x & 31;
  const int bm = __reduce_max_sync(full, valid ? r.dxc : -1);
  // keys are in [0, 8]: order-preserving as unsigned bit patterns
  const unsigned kmin = __reduce_min_sync(full, valid ? __float_as_uint(key) : 0x7f800000u);
  const unsigned kmax = __reduce_max_sync(full, valid ? __float_as_uint(key) : 0u);
  if (lane == 0 && in_range) {
    S.batch_max[i >> 5] = bm;
    S.w_bkey[i >> 5] = make_float2(__uint_as_float(kmin), __uint_as_float(kmax));
  }
  // per level: rays of this warp that reach its first ring
  const unsigned nvalid = __popc(__ballot_sync(full, valid));
  if (lane == 0 && nvalid) atomicAdd(&alive[CS_W_LEVELS], (int)nvalid);
  if (bm >= 1) {
    const int top = cs_w_level_of(bm);
    for (int L = 0; L <= top; L++) {
      const unsigned m = __ballot_sync(full, valid && r.dxc >= cs_w_level_first(L));
      if (lane == 0 && m) atomicAdd(&alive[L], (int)__popc(m));
    }
  }
}

// for every batch of 32 consecutive rays that can hold a candidate of task t (by its key range and largest dxc): f(batch)
template <typename F>
__device__ __forceinline__ void cs_w_for_batches(const CsSession& S, const CsWTask& t, int n, F&& f) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int nb = (n + 31) >> 5;
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int bi = b0 + lane;
    bool ov = false;
    if (bi < nb && __ldcg(S.batch_max + bi) >= t.k0) {
      const float2 kk = __ldcg(S.w_bkey + bi);
      ov = (kk.y >= t.flo && kk.x < t.fhi) || kk.y >= t.wrap_lo;
    }
    unsigned om = __ballot_sync(full, ov);
    while (om) {
      const int b = b0 + __ffs(om) - 1;
      om &= om - 1;
      f(b);
    }
  }
}
__device__ __forceinline__ bool cs_w_is_candidate(const CsWTask& t, const int2 rk) {
  const float key = __int_as_float(rk.x);
  return rk.y >= t.k0 && ((key >= t.flo && key < t.fhi) || key >= t.wrap_lo);
}

// cell of position p on ring k around (x1, y1) (inverse of cs_ring_visit's position)
__device__ __forceinline__ void cs_w_pos_to_xy(int p, int k, int x1, int y1, int& x, int& y) {
  if (p <= 2 * k) { x = x1 + k; y = y1 + (p - k); }
  else if (p <= 4 * k) { y = y1 + k; x = x1 + (3 * k - p); }
  else if (p <= 6 * k) { x = x1 - k; y = y1 + (5 * k - p); }
  else { y = y1 - k; x = x1 + (p - 7 * k); }
}

// ---- ring 0: the start cell, visited by every valid ray, in ray order (one warp)
template <bool TILED>
__device__ void cs_w_ring0(const CsSession& S, uint16_t* __restrict__ map, int n, int x1, int y1, int size, int pitch_tiles, int alpha) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const uint32_t cell = cs_cell_offset<TILED>(x1, y1, size, pitch_tiles);
  CsBlendState st;
  st.val = 0; st.last_pv = -1; st.fixed = false;
  bool loaded = false;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int pv = 0;
    bool valid = false;
    if (i < n) {
      const CsRay r = cs_unpack_ray(__ldcg(S.rays + i));
      valid = (r.flags & 1) != 0;
      if (valid) pv = cs_ray_pixval(r, 0);
    }
    unsigned act = __ballot_sync(full, valid);
    if (!act) continue;
    if (!loaded) { st.val = (int)__ldcg(map + cell); loaded = true; }  // every lane keeps the same running value
    const int first = __ffs(act) - 1;
    const int pv0 = __shfl_sync(full, pv, first);
    if (__ballot_sync(full, valid && pv != pv0) == 0u) {
      st.apply_n(pv0, __popc(act), alpha);  // one pixval: a count, with the exact fixed-point early-out
      st.fixed = false; st.last_pv = -1;
    } else {
      while (act) {
        const int l = __ffs(act) - 1;
        act &= act - 1;
        st.apply(__shfl_sync(full, pv, l), alpha);
      }
    }
  }
  if (loaded && lane == 0) __stcg(map + cell, (uint16_t)st.val);
}

// ---- general path: any number of candidates
template <bool TILED>
__device__ void cs_w_general(const CsSession& S, uint16_t* __restrict__ map, const CsWTask& t, int n, int x1, int y1, int size,
                             int pitch_tiles, int alpha, unsigned* s_val, uint32_t* s_cell) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (int k = t.k0; k <= t.k1; k++) {
    const int lo = cs_w_bound(k, t.blo), hi = cs_w_bound(k, t.bhi);
    for (int c0 = lo; c0 < hi; c0 += CS_W_CAP) {
      const int cn = min(CS_W_CAP, hi - c0);
      for (int i = lane; i < cn; i += 32) {  // stage the wedge's cells of this ring
        int x, y;
        cs_w_pos_to_xy(c0 + i, k, x1, y1, x, y);
        const bool on = (unsigned)x < (unsigned)size && (unsigned)y < (unsigned)size;
        const uint32_t cell = on ? cs_cell_offset<TILED>(x, y, size, pitch_tiles) : 0xffffffffu;
        s_cell[i] = cell;
        s_val[i] = on ? (unsigned)__ldcg(map + cell) : 0u;
      }
      __syncwarp();
      cs_w_for_batches(S, t, n, [&](int b) {
        const int i = b * 32 + lane;
        bool inw = false;
        int pos = 0, pv = 0;
        uint32_t cell = 0;
        if (i < n && cs_w_is_candidate(t, __ldcg(S.w_rk + i))) {
          const CsRay r = cs_unpack_ray(__ldcg(S.rays + i));
          inw = cs_ring_visit<TILED>(r, k, x1, y1, size, pitch_tiles, pos, cell, pv) && pos >= c0 && pos < c0 + cn;
        }
        const unsigned act = __ballot_sync(full, inw);
        if (!act) return;
        unsigned rest = 0u;   // leaders of groups with several pixvals: the other members, applied in lane order below
        bool lead = false;
        int v = 0;
        if (inw) {
          const unsigned grp = __match_any_sync(act, pos);
          const unsigned same = __match_any_sync(act, ((unsigned long long)(unsigned)pos << 32) | (unsigned)pv);
          lead = lane == __ffs(grp) - 1;
          if (lead) {
            v = (int)(s_val[pos - c0] & 0xffffu);
            if (same == grp) {  // one pixval: blends of one value commute with themselves — apply it as a count
              CsBlendState st;
              st.val = v; st.last_pv = -1; st.fixed = false;
              st.apply_n(pv, __popc(grp), alpha);
              v = st.val;
            } else {
              v = cs_blend(v, pv, alpha);
              rest = grp & ~(1u << lane);
            }
          }
        }
        unsigned any_rest = __ballot_sync(full, rest != 0u);
        while (any_rest) {  // warp-uniform: every leader with members left fetches its next one
          const int src = rest ? __ffs(rest) - 1 : lane;
          const int pvj = __shfl_sync(full, pv, src);
          if (rest) { v = cs_blend(v, pvj, alpha); rest &= rest - 1; }
          any_rest = __ballot_sync(full, rest != 0u);
        }
        if (lead) s_val[pos - c0] = (unsigned)v | 0x10000u;  // touched
        __syncwarp();
      });
      for (int i = lane; i < cn; i += 32) {
        const unsigned w = s_val[i];
        if (w & 0x10000u) __stcg(map + s_cell[i], (uint16_t)(w & 0xffffu));
      }
      __syncwarp();
    }
  }
}

// ---- fast path: at most 32 candidates (s_list, in ray order), one per lane
template <bool TILED>
__device__ void cs_w_fast(const CsSession& S, uint16_t* __restrict__ map, const CsWTask& t, int ncand, const int* s_list, int x1, int y1,
                          int size, int pitch_tiles, int alpha) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // this lane's ray, and its walk state at ring k0 - 1
  CsRay r;
  r.dxc = -1; r.dyc = 0; r.a0 = 0; r.b0 = -1; r.incv = 0; r.kc = 0; r.nd_total = 0; r.flags = 0;
  if (lane < ncand) r = cs_unpack_ray(__ldcg(S.rays + s_list[lane]));
  const int dxc = (lane < ncand) ? r.dxc : -1;
  const int dyc = min(r.dyc, r.dxc);  // a clipped dyc > dxc walks the diagonal (m = k), like dyc = dxc
  const int den2 = 2 * max(dxc, 1), dy2 = 2 * dyc;
  int cs_, g;
  cs_w_side(r.flags, cs_, g);
  const bool steep = (r.flags & 2) != 0;
  const int smaj = (r.flags & 4) ? -1 : 1, smin = (r.flags & 8) ? -1 : 1;
  const int dxM = steep ? 0 : smaj, dyM = steep ? smaj : 0;  // per ring
  const int dxN = steep ? smin : 0, dyN = steep ? 0 : smin;  // per minor step
  int k = t.k0 - 1;
  int q, rem;
  {
    const int num = dy2 * k + max(dxc, 1) - 1;  // (2 dyc k + dxc - 1) = q * 2 dxc + rem
    q = (k == 0) ? 0 : cs_ray_minor(r, k);
    if (dxc < 1) q = 0;
    rem = num - q * den2;
  }
  int x = x1 + dxM * k + dxN * q, y = y1 + dyM * k + dyN * q;
  int pos = cs_ * k + g * q;
  unsigned accl = (unsigned)k * t.blo, acch = (unsigned)k * t.bhi;
  const int c0 = max(r.a0, r.b0 + 1);

  for (int kb = t.k0; kb <= t.k1; kb += CS_W_G) {
    uint32_t cell[CS_W_G];
    int pv[CS_W_G], val[CS_W_G];
    unsigned grp[CS_W_G];
    unsigned leadbits = 0u, cfbits = 0u;
#pragma unroll
    for (int j = 0; j < CS_W_G; j++) {
      k = kb + j;
      rem += dy2;
      const bool wrap = rem >= den2;
      if (wrap) { rem -= den2; x += dxN; y += dyN; pos += g; }
      x += dxM; y += dyM; pos += cs_;
      accl += t.blo; acch += t.bhi;
      const int lo = (int)((accl + ((1u << CS_W_FIX) - 1u)) >> CS_W_FIX), hi = (int)((acch + ((1u << CS_W_FIX) - 1u)) >> CS_W_FIX);
      const int posn = (pos == 8 * k) ? 0 : pos;
      const bool inw = k <= t.k1 && k <= dxc && posn >= lo && posn < hi && (unsigned)x < (unsigned)size && (unsigned)y < (unsigned)size;
      // pixval (closed form of :402-428, cs_ray_pixval with c0 hoisted)
      int p;
      if (k <= r.b0) p = CS_TS_NO_OBSTACLE + max(k - r.a0 + 1, 0) * r.incv;
      else if (k < c0) p = CS_TS_NO_OBSTACLE;
      else { const int jj = k - c0 + 1; p = CS_TS_NO_OBSTACLE + (r.nd_total - jj) * r.incv + min(jj, r.kc); }
      pv[j] = p;
      cell[j] = cs_cell_offset<TILED>(x, y, size, pitch_tiles);
      grp[j] = 0u;
      val[j] = 0;
      const unsigned act = __ballot_sync(full, inw);
      if (act) {
        bool lead = false;
        if (inw) {
          grp[j] = __match_any_sync(act, posn);
          lead = lane == __ffs(grp[j]) - 1;
          if (lead) val[j] = (int)__ldcg(map + cell[j]);
        }
        if (lead) leadbits |= 1u << j;
        if (__ballot_sync(full, inw && !lead)) cfbits |= 1u << j;  // some cell of this ring has several visitors in this warp
      }
    }
#pragma unroll
    for (int j = 0; j < CS_W_G; j++) {
      const bool lead = (leadbits >> j) & 1u;
      int v = cs_blend(val[j], pv[j], alpha);
      if ((cfbits >> j) & 1u) {  // warp-uniform
        unsigned rest = lead ? (grp[j] & ~(1u << lane)) : 0u;
        unsigned any_rest = __ballot_sync(full, rest != 0u);
        while (any_rest) {
          const int src = rest ? __ffs(rest) - 1 : lane;
          const int pvj = __shfl_sync(full, pv[j], src);
          if (rest) { v = cs_blend(v, pvj, alpha); rest &= rest - 1; }
          any_rest = __ballot_sync(full, rest != 0u);
        }
      }
      if (lead) __stcg(map + cell[j], (uint16_t)v);
    }
  }
}

template <bool TILED>
__global__ void __launch_bounds__(CS_W_THREADS, 4)
cs_wedge_kernel(CsSession* __restrict__ sessions, CsStepArgs a) {
  __shared__ int s_first[CS_W_LEVELS + 2];   // first task of each level (task 0 is ring 0); s_first[nlev] = number of tasks
  __shared__ short s_wedges[CS_W_LEVELS];    // wedges per level
  __shared__ int s_nlev;
  __shared__ float sh_pose[5];
  __shared__ long long sh_vis[CS_W_WARPS];
  __shared__ unsigned s_val[CS_W_WARPS][CS_W_CAP];
  __shared__ uint32_t s_cell[CS_W_WARPS][CS_W_CAP];
  __shared__ int s_list[CS_W_WARPS][32];

  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  cs_pdl_launch_dependents();  // the next step's search may become resident once every block here has started
  // No griddepcontrol.wait in front: like the rings kernel, this one is resident before the search in front has ended; its
  // preparing blocks poll the pose words, everybody else the preparing blocks' count (see cs_rings_kernel).

  const int sj = blockIdx.y;
  CsSession& S = sessions[sj];
  const CsStepHeader& hdr = a.hdr[(size_t)sj * a.hdr_stride];
  const int n = hdr.n_points;
  const int size = S.size, pitch_tiles = S.pitch_tiles;
  const float scale = S.scale;
  const int alpha = S.quality;
  uint16_t* __restrict__ map = S.map;
  const int copies = S.ray_copies;
  const unsigned slot = a.step_id & 1u;
  const unsigned long long ll_tag = (unsigned long long)a.step_id << 32;
  int* alive = S.w_alive + (size_t)a.w_slot * (CS_W_LEVELS + 1);

  // ---- ray preparation: the first nprep blocks take a.prep_group rays each, one ray per thread
  const int group = a.prep_group;
  const int nprep = (n + group - 1) / group;
  if ((int)blockIdx.x < nprep) {
    const float2* __restrict__ points = a.points + (size_t)sj * a.points_stride;
    const int g_begin = blockIdx.x * group, g_end = min(n, g_begin + group);
    if (blockIdx.x == 0) {  // the other half of the counters is this step's to re-arm (nobody reads or counts into it now)
      int* other = S.w_alive + (size_t)(a.w_slot ^ 1) * (CS_W_LEVELS + 1);
      for (int i = tid; i <= CS_W_LEVELS; i += CS_W_THREADS) other[i] = 0;
    }
    if (tid < 5) {
      volatile unsigned long long* ll = S.ll_pose + tid;
      unsigned long long w;
      while (((w = *ll) & 0xffffffff00000000ull) != ll_tag) {}
      sh_pose[tid] = __uint_as_float((unsigned)w);
    }
    __syncthreads();
    const float pose[3] = {sh_pose[0], sh_pose[1], sh_pose[2]};
    const float cs[2] = {sh_pose[3], sh_pose[4]};
    const CsRayFrame f = cs_ray_frame(S, pose, cs);
    long long vis = 0;
    for (int base = g_begin; base < g_end; base += CS_W_THREADS) {  // whole warps: the warp reductions need every lane
      const int i = base + tid;
      if (base + warp * 32 < g_end) {
        const bool in_range = i < g_end;
        const float2 p = in_range ? __ldg(points + i) : make_float2(1.f, 0.f);
        cs_w_prepare(S, f, p, i, in_range, alive, vis);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vis += __shfl_xor_sync(full, vis, o);
    if (lane == 0) sh_vis[warp] = vis;
    __threadfence();  // this thread's stores and counts are visible device-wide before the block's arrival is counted
    __syncthreads();
    if (tid < copies) atomicAdd(S.prep_words + ((size_t)slot * copies + tid) * 16, 1ull);
    if (tid == 0) {
      for (int w = 1; w < CS_W_WARPS; w++) vis += sh_vis[w];
      if (vis) {
        atomicAdd((unsigned long long*)&S.visits_slot[slot], (unsigned long long)vis);
        if (a.visits_out) atomicAdd((unsigned long long*)a.visits_out, (unsigned long long)vis);
      }
    }
  }
  // ---- everybody: wait for the preparing blocks, then build the task table of this scan
  if (tid == 0) {
    volatile unsigned long long* pw = S.prep_words + ((size_t)slot * copies + (cs_smid() % copies)) * 16;
    while (*pw != (unsigned long long)nprep) {}
    __threadfence();
  }
  __syncthreads();
  if (warp == 0) {
    int total = 1, nlev = 0;  // task 0: ring 0
    for (int L0 = 0; L0 < CS_W_LEVELS; L0 += 32) {
      const int L = L0 + lane;
      const int al = L < CS_W_LEVELS ? __ldcg(alive + L) : 0;
      const int W = al > 0 ? cs_w_wedges(al, cs_w_level_first(L)) : 0;
      int incl = W;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += u;
      }
      if (L < CS_W_LEVELS) { s_first[L] = total + incl - W; s_wedges[L] = (short)W; }
      const unsigned has = __ballot_sync(full, W > 0);
      if (has) nlev = L0 + 32 - __clz(has);
      total += __shfl_sync(full, incl, 31);
      if (!has) break;  // levels are reached by fewer and fewer rays: an empty chunk ends the scan
    }
    if (lane == 0) { s_first[nlev] = total; s_nlev = nlev; }
  }
  __syncthreads();
  const int nlev = s_nlev;
  const int n_tasks = s_first[nlev];
  const int n_valid = __ldcg(alive + CS_W_LEVELS);
  const float pose_x = __uint_as_float((unsigned)__ldcg(&S.ll_pose[0]));
  const float pose_y = __uint_as_float((unsigned)__ldcg(&S.ll_pose[1]));
  const int x1 = cs_cvt_i32(__fadd_rn(__fmul_rn(pose_x, scale), 0.5f));  // :499, :505
  const int y1 = cs_cvt_i32(__fadd_rn(__fmul_rn(pose_y, scale), 0.5f));  // :500, :506

  // ---- tasks: static round-robin over the warps of the session's blocks, the (long) centre tasks first, starting with
  // the warps of the blocks that did not prepare rays
  if (n_valid > 0) {
    const int total_warps = gridDim.x * CS_W_WARPS;
    int gw = (int)blockIdx.x * CS_W_WARPS + warp - min(nprep, (int)gridDim.x - 1) * CS_W_WARPS;
    if (gw < 0) gw += total_warps;
    for (int task = gw; task < n_tasks; task += total_warps) {
      if (task == 0) {
        cs_w_ring0<TILED>(S, map, n, x1, y1, size, pitch_tiles, alpha);
        continue;
      }
      int L = 0;  // level of the task: last L with s_first[L] <= task
      {
        int lo = 0, hi = nlev - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_first[mid] <= task) lo = mid; else hi = mid - 1;
        }
        L = lo;
      }
      const int W = s_wedges[L], w = task - s_first[L];
      if (W <= 0 || w >= W) continue;  // (empty level inside the table: its range of tasks is empty)
      CsWTask t;
      t.k0 = cs_w_level_first(L);
      t.k1 = cs_w_level_last(L);
      t.blo = cs_w_beta(w, W);
      t.bhi = cs_w_beta(w + 1, W);
      const float eps = 0.5f / (float)t.k0 + 1e-4f;
      t.flo = (float)t.blo * (1.0f / (float)(1 << CS_W_FIX)) - eps;
      t.fhi = (float)t.bhi * (1.0f / (float)(1 << CS_W_FIX)) + eps;
      t.wrap_lo = (w == 0) ? 8.0f - eps : 9.0f;
      // candidates, in ray order; the largest dxc among them bounds the rings of the task
      int ncand = 0, kmax = 0;
      cs_w_for_batches(S, t, n, [&](int b) {
        const int i = b * 32 + lane;
        int2 rk = make_int2(0, -1);
        if (i < n) rk = __ldcg(S.w_rk + i);
        const bool cand = i < n && cs_w_is_candidate(t, rk);
        const unsigned m = __ballot_sync(full, cand);
        if (cand) {
          const int at = ncand + __popc(m & ((1u << lane) - 1u));
          if (at < 32) s_list[warp][at] = i;
        }
        ncand += __popc(m);
        kmax = max(kmax, __reduce_max_sync(full, cand ? rk.y : 0));
      });
      if (ncand == 0) continue;
      t.k1 = min(t.k1, kmax);
      __syncwarp();
      if (ncand <= 32 && !a.w_general)
        cs_w_fast<TILED>(S, map, t, ncand, s_list[warp], x1, y1, size, pitch_tiles, alpha);
      else
        cs_w_general<TILED>(S, map, t, n, x1, y1, size, pitch_tiles, alpha, s_val[warp], s_cell[warp]);
      __syncwarp();
    }
  }
  cs_pdl_wait();  // the kernel in front has long finished; this only makes "this grid done" imply "that grid done"
}
