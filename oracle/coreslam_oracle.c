/* coreslam_oracle.c — see coreslam_oracle.h.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED against an
 * executed reference (no .NET here), pinned by KAT-A..E + an independent transliteration.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fwrapv  (float32, one rounding per operation and
 * wrapping int32, which is what RyuJIT emits for the cited C#).  Citations: /root/reference.
 */
#include "coreslam_oracle.h"

#include <emmintrin.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* HoleMap.cs:17-22 */
or_holemap* or_holemap_create(int size_pixels, float size_meters) {
  or_holemap* m = (or_holemap*)calloc(1, sizeof(*m));
  m->size = size_pixels;
  m->scale = (float)size_pixels / size_meters; /* int -> float, then float divide */
  m->pixels = (uint16_t*)calloc((size_t)size_pixels * (size_t)size_pixels, sizeof(uint16_t));
  return m;
}

void or_holemap_destroy(or_holemap* m) {
  if (!m) return;
  free(m->pixels);
  free(m);
}

/* HoleMap.cs:44-55 */
void or_holemap_packed(const or_holemap* m, uint8_t* out) {
  size_t n = ((size_t)m->size * (size_t)m->size) / 2;
  for (size_t i = 0; i < n; i++)
    out[i] = (uint8_t)(((m->pixels[i * 2] >> 12) << 4) | (m->pixels[i * 2 + 1] >> 12));
}

/* RyuJIT x64 lowers (int)float to cvttss2si: truncation, 0x80000000 for NaN / out of range. */
int32_t or_cvt(float f) { return _mm_cvttss_si32(_mm_set_ss(f)); }

/* MathEx.cs:116-138 */
static float normalize_angle_pos(float angle) {
  float pi2 = (float)M_PI * 2.0f; /* MathF.PI * 2.0f */
  return fmodf(fmodf(angle, pi2) + pi2, pi2);
}

float or_normalize_angle(float angle) {
  float a = normalize_angle_pos(angle);
  if (a > (float)M_PI) a -= 2.0f * (float)M_PI;
  return a;
}

/* CoreSLAMProcessor.cs:226-259 */
int32_t or_distance(const or_holemap* m, const float* points, int n, const float pose[3]) {
  int nb_points = 0;
  int64_t sum = 0;
  const float scale = m->scale;
  const int size = m->size;

  float px = pose[0] * scale + 0.5f;  /* :232 */
  float py = pose[1] * scale + 0.5f;  /* :233 */
  float c = cosf(pose[2]) * scale;    /* :234, MathF.Cos -> libm cosf */
  float s = sinf(pose[2]) * scale;    /* :235 */

  for (int i = 0; i < n; i++) {
    float X = points[2 * i], Y = points[2 * i + 1];
    float fx = px + c * X; /* :240, left-associative, every op rounded to float */
    fx = fx - s * Y;
    float fy = py + s * X; /* :241 */
    fy = fy + c * Y;
    int32_t x = or_cvt(fx);
    int32_t y = or_cvt(fy);
    if (x >= 0 && x < size && y >= 0 && y < size) { /* :244 */
      sum += m->pixels[y * size + x];
      nb_points++;
    }
  }
  if (nb_points > 0) return (int32_t)((sum * 1024) / n); /* :253 — divides by ALL points */
  return INT32_MAX;                                      /* :257 */
}

/* CoreSLAMProcessor.cs:320-345 */
int or_clip_ray(int size, int32_t* xyc, int32_t* yxc, int32_t xy, int32_t yx) {
  if (*xyc < 0) {
    if (*xyc == xy) return 0;
    *yxc += (*yxc - yx) * (-*xyc) / (*xyc - xy);
    *xyc = 0;
  }
  if (*xyc >= size) {
    if (*xyc == xy) return 0;
    *yxc += (*yxc - yx) * (size - 1 - *xyc) / (*xyc - xy);
    *xyc = size - 1;
  }
  return 1;
}

static int32_t iabs(int32_t v) { return v < 0 ? -v : v; }
static int32_t isign(int32_t v) { return (v > 0) - (v < 0); }

/* CoreSLAMProcessor.cs:359-443 */
int64_t or_draw_ray(or_holemap* m, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t xp, int32_t yp,
                    int32_t value, int32_t alpha, int32_t* trace, int trace_cap) {
  const int size = m->size;
  int32_t x2c = x2, y2c = y2;
  if (!or_clip_ray(size, &x2c, &y2c, x1, y1)) return 0; /* :365 */
  if (!or_clip_ray(size, &y2c, &x2c, y1, x1)) return 0; /* :366 */

  int32_t dx = iabs(x2 - x1), dy = iabs(y2 - y1);       /* :368-369 (unclipped) */
  int32_t dxc = iabs(x2c - x1), dyc = iabs(y2c - y1);   /* :370-371 (clipped) */
  int32_t incptrx = isign(x2 - x1);                     /* :372 */
  int32_t incptry = isign(y2 - y1) * size;              /* :373 */
  int32_t sincv = isign(value - OR_TS_NO_OBSTACLE);     /* :374 */
  int32_t derrorv;

  if (dx > dy) {
    derrorv = iabs(xp - x2); /* :379 */
  } else {
    int32_t t;
    dx = dy;                                   /* :383 */
    t = dxc; dxc = dyc; dyc = t;               /* :384 */
    t = incptrx; incptrx = incptry; incptry = t; /* :385 */
    derrorv = iabs(yp - y2);                   /* :386 */
  }
  if (derrorv == 0) return 0; /* :389-392 */

  int32_t error = 2 * dyc - dxc;                            /* :394 */
  int32_t horiz = 2 * dyc;                                  /* :395 */
  int32_t diago = 2 * (dyc - dxc);                          /* :396 */
  int32_t errorv = derrorv / 2;                             /* :397 */
  int32_t incv = (value - OR_TS_NO_OBSTACLE) / derrorv;     /* :398 */
  int32_t incerrorv = value - OR_TS_NO_OBSTACLE - derrorv * incv; /* :399 */
  int32_t ptr = y1 * size + x1;                             /* :401 */
  int32_t pixval = OR_TS_NO_OBSTACLE;                       /* :402 */
  int64_t visits = 0;

  for (int32_t x = 0; x <= dxc; x++, ptr += incptrx) { /* :404 */
    if (x > dx - 2 * derrorv) {                         /* :406 */
      if (x <= dx - derrorv) {                          /* :408 */
        pixval += incv;
        errorv += incerrorv;
        if (errorv > derrorv) {
          pixval += sincv;
          errorv -= derrorv;
        }
      } else {
        pixval -= incv;
        errorv -= incerrorv;
        if (errorv < 0) {
          pixval -= sincv;
          errorv += derrorv;
        }
      }
    }
    /* :431 */
    uint16_t nv = (uint16_t)(((256 - alpha) * (int32_t)m->pixels[ptr] + alpha * pixval) >> 8);
    m->pixels[ptr] = nv;
    if (trace && visits < trace_cap) {
      trace[3 * visits + 0] = ptr;
      trace[3 * visits + 1] = pixval;
      trace[3 * visits + 2] = nv;
    }
    visits++;
    if (error > 0) { /* :433-441 */
      ptr += incptry;
      error += diago;
    } else {
      error += horiz;
    }
  }
  return visits;
}

/* CoreSLAMProcessor.cs:496-534 */
int64_t or_update_hole_map(or_holemap* m, const float* points, int n, const float pose[3], float hole_width,
                           int quality, int32_t* rays_out) {
  const float scale = m->scale;
  float px = pose[0] * scale + 0.5f; /* :499 */
  float py = pose[1] * scale + 0.5f;
  float c = cosf(pose[2]) * scale;   /* :501 */
  float s = sinf(pose[2]) * scale;
  int32_t x1 = or_cvt(px), y1 = or_cvt(py); /* :505-506 */
  int64_t visits = 0;

  if (x1 < 0 || x1 >= m->size || y1 < 0 || y1 >= m->size) return 0; /* :509-512 */

  for (int i = 0; i < n; i++) { /* :517 */
    float X = points[2 * i], Y = points[2 * i + 1];
    float x2p = c * X - s * Y; /* :519 */
    float y2p = s * X + c * Y; /* :520 */
    int32_t xp = or_cvt(px + x2p);
    int32_t yp = or_cvt(py + y2p);
    float dist = sqrtf(x2p * x2p + y2p * y2p);    /* :524 */
    float add = hole_width * scale / 2.0f / dist; /* :525 */
    x2p *= (1.0f + add);
    y2p *= (1.0f + add);
    int32_t x2 = or_cvt(px + x2p);
    int32_t y2 = or_cvt(py + y2p);
    if (rays_out) {
      int32_t* r = rays_out + 6 * i;
      r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = xp; r[5] = yp;
    }
    visits += or_draw_ray(m, x1, y1, x2, y2, xp, yp, OR_TS_OBSTACLE, quality, NULL, 0); /* :532 */
  }
  return visits;
}

/* CoreSLAMProcessor.cs:624-653 */
void or_monte_carlo_search(const or_holemap* m, const float* points, int n, const float search_pose[3],
                           const float* offsets, int iterations, float best_pose[3], int32_t* best_distance,
                           int32_t* distances_out) {
  float best[3] = {search_pose[0], search_pose[1], search_pose[2]};
  int32_t current = or_distance(m, points, n, search_pose); /* :627 */
  int32_t bestd = current;
  if (distances_out) distances_out[0] = current;

  for (int k = 0; k < iterations; k++) { /* :630 */
    float cur[3];
    cur[0] = search_pose[0] + offsets[3 * k + 0]; /* :635 */
    cur[1] = search_pose[1] + offsets[3 * k + 1];
    cur[2] = search_pose[2] + offsets[3 * k + 2];
    current = or_distance(m, points, n, cur); /* :641 */
    if (distances_out) distances_out[1 + k] = current;
    if (current < bestd) { /* :644, strict */
      bestd = current;
      best[0] = cur[0]; best[1] = cur[1]; best[2] = cur[2];
    }
  }
  *best_distance = bestd;
  best_pose[0] = best[0]; best_pose[1] = best[1]; best_pose[2] = best[2];
}

/* CoreSLAMProcessor.cs:674-710 (serial execution of the per-thread bodies) */
void or_parallel_search(const or_holemap* m, const float* points, int n, const float search_pose[3],
                        const float* offsets, int iterations, int threads, float best_pose[3],
                        int32_t* best_distance, int32_t* distances_out, int32_t* best_index_out) {
  int32_t bestd = INT32_MAX; /* :695 */
  float best[3] = {search_pose[0], search_pose[1], search_pose[2]}; /* :696 */
  int32_t best_index = 0;
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(iterations + 1));

  for (int t = 0; t < threads; t++) {
    float pose_t[3];
    int32_t dist_t;
    or_monte_carlo_search(m, points, n, search_pose, offsets + (size_t)3 * t * iterations, iterations, pose_t,
                          &dist_t, tmp);
    if (distances_out) {
      distances_out[0] = tmp[0];
      memcpy(distances_out + 1 + (size_t)t * iterations, tmp + 1, sizeof(int32_t) * (size_t)iterations);
    }
    if (dist_t < bestd) { /* :700, strict: lowest thread wins ties */
      bestd = dist_t;
      best[0] = pose_t[0]; best[1] = pose_t[1]; best[2] = pose_t[2];
      /* flat index of thread t's winner: first minimum in (searchPose, its candidates) */
      int32_t idx = 0, dmin = tmp[0];
      for (int k = 0; k < iterations; k++)
        if (tmp[1 + k] < dmin) { dmin = tmp[1 + k]; idx = 1 + t * iterations + k; }
      best_index = idx;
    }
  }
  free(tmp);
  *best_distance = bestd;
  best_pose[0] = best[0]; best_pose[1] = best[1]; best_pose[2] = best[2];
  if (best_index_out) *best_index_out = best_index;
}

/* CoreSLAMProcessor.cs:187-207 */
void or_segment_to_cloud(const float* rays, int n, const float segment_pose[3], const float odometry_pose[3],
                         float* points_out) {
  float px = segment_pose[0] - odometry_pose[0]; /* :194 */
  float py = segment_pose[1] - odometry_pose[1];
  float pz = segment_pose[2] - odometry_pose[2];
  for (int i = 0; i < n; i++) {
    float angle = rays[2 * i], radius = rays[2 * i + 1];
    points_out[2 * i + 0] = px + radius * cosf(angle + pz); /* :200 */
    points_out[2 * i + 1] = py + radius * sinf(angle + pz); /* :201 */
  }
}

/* ------------------------------------------------------------------------------------------- */
/* CoreSLAM/ObstacleMap.cs:17-22 */
or_obstaclemap* or_obstaclemap_create(int size_pixels, float size_meters) {
  or_obstaclemap* m = (or_obstaclemap*)calloc(1, sizeof(*m));
  m->size = size_pixels;
  m->scale = (float)size_pixels / size_meters; /* ObstacleMap.cs:20, int -> float, then divide */
  size_t n = (size_t)size_pixels * (size_t)size_pixels;
  m->pixels = (int8_t*)calloc(n ? n : 1, 1);
  m->no_hit = (uint8_t*)calloc(n ? n : 1, 1);
  return m;
}

void or_obstaclemap_destroy(or_obstaclemap* m) {
  if (!m) return;
  free(m->pixels);
  free(m->no_hit);
  free(m);
}

/* CoreSLAMProcessor.cs:456-490 */
int64_t or_draw_ray_obstacle(or_obstaclemap* m, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int max_obstacle_hits) {
  const int size = m->size;
  int32_t ddx = (int32_t)((uint32_t)x2 - (uint32_t)x1), ddy = (int32_t)((uint32_t)y2 - (uint32_t)y1);
  if (ddx == INT32_MIN || ddy == INT32_MIN) return -1; /* Math.Abs(int.MinValue) throws */
  int32_t dx = ddx < 0 ? -ddx : ddx, sx = (ddx > 0) - (ddx < 0); /* :458 */
  int32_t dy = ddy < 0 ? -ddy : ddy, sy = (ddy > 0) - (ddy < 0); /* :459 */
  int32_t err = (dx > dy ? dx : -dy) / 2, e2;                    /* :460 */
  int64_t touched = 0;
  for (;;) { /* :462 */
    if (x1 < 0 || x1 >= size || y1 < 0 || y1 >= size) { /* :465-466 */
      break;
    } else if (x1 == x2 && y1 == y2) { /* :471 */
      int8_t* px = &m->pixels[(size_t)y1 * size + x1];
      if (*px < max_obstacle_hits) (*px)++; /* :474-477 */
      touched++;
      break;
    } else {
      m->no_hit[(size_t)y1 * size + x1] = 1; /* :483 */
      touched++;
    }
    e2 = err;                                                                   /* :486 */
    if (e2 > -dx) { err = (int32_t)((uint32_t)err - (uint32_t)dy); x1 += sx; } /* :487 */
    if (e2 < dy) { err = (int32_t)((uint32_t)err + (uint32_t)dx); y1 += sy; }  /* :488 */
  }
  return touched;
}

/* CoreSLAMProcessor.cs:540-593 */
int64_t or_update_obstacle_map(or_obstaclemap* m, const float* points, int n, const float pose[3],
                               int max_obstacle_hits) {
  const int size = m->size;
  memset(m->no_hit, 0, (size_t)size * (size_t)size); /* :542 */
  float px = pose[0] * m->scale + 0.5f; /* :545 */
  float py = pose[1] * m->scale + 0.5f; /* :546 */
  float c = cosf(pose[2]) * m->scale;   /* :547 */
  float s = sinf(pose[2]) * m->scale;   /* :548 */
  int32_t x1 = or_cvt(px), y1 = or_cvt(py); /* :553-554 */
  if (x1 < 0 || x1 >= size || y1 < 0 || y1 >= size) return 0; /* :557-560 */
  int64_t touched = 0;
  for (int i = 0; i < n; i++) { /* :563 */
    float X = points[2 * i], Y = points[2 * i + 1];
    float fx = px + c * X; /* :566, left-associative, one rounding per operation */
    fx = fx - s * Y;
    float fy = py + s * X; /* :567 */
    fy = fy + c * Y;
    int32_t x2 = or_cvt(fx), y2 = or_cvt(fy);
    int64_t t = or_draw_ray_obstacle(m, x1, y1, x2, y2, max_obstacle_hits); /* :570 */
    if (t > 0) touched += t;
  }
  for (int y = 0; y < size; y++) { /* :576-592 */
    for (int x = 0; x < size; x++) {
      size_t k = (size_t)y * size + x;
      if (m->no_hit[k]) {
        if (m->pixels[k] < 0) m->pixels[k]++;
        else if (m->pixels[k] > 0) m->pixels[k]--;
      }
    }
  }
  return touched;
}

/* ------------------------------------------------------------------------------------------- */
/* CoreSLAMProcessor.cs:119-175 */
or_processor* or_processor_create_full(float physical_map_size, int hole_map_size, int obstacle_map_size,
                                       const float start_pose[3], float sigma_xy, float sigma_theta,
                                       int iterations_per_thread, int num_search_threads) {
  or_processor* p = or_processor_create(physical_map_size, hole_map_size, start_pose, sigma_xy, sigma_theta,
                                        iterations_per_thread, num_search_threads);
  if (obstacle_map_size > 0) {
    p->omap = or_obstaclemap_create(obstacle_map_size, physical_map_size); /* :132 */
    or_processor_reset(p);
  }
  return p;
}

or_processor* or_processor_create(float physical_map_size, int hole_map_size, const float start_pose[3],
                                  float sigma_xy, float sigma_theta, int iterations_per_thread,
                                  int num_search_threads) {
  or_processor* p = (or_processor*)calloc(1, sizeof(*p));
  p->physical_map_size = physical_map_size;
  memcpy(p->start_pose, start_pose, sizeof(float) * 3);
  p->sigma_xy = sigma_xy;
  p->sigma_theta = sigma_theta;
  p->iterations_per_thread = iterations_per_thread;
  p->num_search_threads = num_search_threads;
  p->quality = 50;
  p->hole_width = 0.6f;
  p->position_search_beginning = 5;
  p->unmapped_obstacle_hits = -5; /* :98 */
  p->max_obstacle_hits = 10;      /* :103 */
  p->map = or_holemap_create(hole_map_size, physical_map_size); /* :131 */
  or_processor_reset(p);                                        /* :140 */
  return p;
}

void or_processor_destroy(or_processor* p) {
  if (!p) return;
  or_holemap_destroy(p->map);
  or_obstaclemap_destroy(p->omap);
  free(p);
}

/* CoreSLAMProcessor.cs:167-175 */
void or_processor_reset(or_processor* p) {
  size_t n = (size_t)p->map->size * (size_t)p->map->size;
  uint16_t v = (uint16_t)((OR_TS_OBSTACLE + OR_TS_NO_OBSTACLE) / 2); /* :169 -> 32750 */
  for (size_t i = 0; i < n; i++) p->map->pixels[i] = v;
  if (p->omap) /* :170 ArrayEx.Fill(ObstacleMap.Pixels, UnmappedObstacleHits) */
    memset(p->omap->pixels, (int8_t)p->unmapped_obstacle_hits, (size_t)p->omap->size * (size_t)p->omap->size);
  p->obstacle_visits = 0;
  memcpy(p->pose, p->start_pose, sizeof(float) * 3); /* :172 */
  memset(p->last_odometry_pose, 0, sizeof(float) * 3); /* :173 */
  p->scan_count = 0; /* :174 */
  p->visits = 0;
  p->last_distance = INT32_MAX;
  p->last_index = 0;
}

static void update_common(or_processor* p, or_worker* w, const float* points, int n, const float odo[3],
                          const float* offsets) {
  float new_pose[3];
  if (p->scan_count >= p->position_search_beginning) { /* :726 */
    float search_pose[3];
    for (int k = 0; k < 3; k++) search_pose[k] = p->pose[k] + (odo[k] - p->last_odometry_pose[k]); /* :728 */
    int threads = p->num_search_threads;
    if (threads <= 0) { /* :736 SingleMonteCarloSearch */
      or_monte_carlo_search(p->map, points, n, search_pose, offsets, p->iterations_per_thread, new_pose,
                            &p->last_distance, NULL);
      p->last_index = -1;
    } else if (w) {
      or_parallel_search_mt(w, p->map, points, n, search_pose, offsets, p->iterations_per_thread, new_pose,
                            &p->last_distance);
      p->last_index = -1;
    } else {
      or_parallel_search(p->map, points, n, search_pose, offsets, p->iterations_per_thread, threads, new_pose,
                         &p->last_distance, NULL, &p->last_index);
    }
  } else {
    p->scan_count++; /* :741 */
    memcpy(new_pose, odo, sizeof(float) * 3);
    p->last_distance = INT32_MAX;
    p->last_index = 0;
  }
  memcpy(p->last_odometry_pose, odo, sizeof(float) * 3); /* :745 */
  new_pose[2] = or_normalize_angle(new_pose[2]);         /* :746 */
  memcpy(p->pose, new_pose, sizeof(float) * 3);          /* :747 */
  p->visits = or_update_hole_map(p->map, points, n, p->pose, p->hole_width, p->quality, NULL); /* :750 */
  if (p->omap) p->obstacle_visits = or_update_obstacle_map(p->omap, points, n, p->pose, p->max_obstacle_hits); /* :751 */
}

/* CoreSLAMProcessor.cs:717-752 */
void or_processor_update(or_processor* p, const float* points, int n, const float odometry_pose[3],
                         const float* offsets) {
  update_common(p, NULL, points, n, odometry_pose, offsets);
}

void or_processor_update_mt(or_processor* p, or_worker* w, const float* points, int n,
                            const float odometry_pose[3], const float* offsets) {
  update_common(p, w, points, n, odometry_pose, offsets);
}

/* ------------------------------------------------------------------------------------------- */
/* Persistent worker pool — BaseSLAM/ParallelWorker.cs:34-117.  One queue slot + signal per thread
 * (SignalConcurrentQueue.cs:14-50), Work() enqueues one item per thread and waits for all. */
typedef void (*or_action)(int index, void* ctx);

typedef struct {
  pthread_t thread;
  pthread_mutex_t mu;
  pthread_cond_t enqueued; /* EnqueuedItemSignal */
  pthread_cond_t done;     /* per-item WaitHandle */
  or_action action;
  void* ctx;
  int pending, finished, cancel, index;
} or_slot;

struct or_worker {
  int num_threads;
  or_slot* slots;
};

static void* work_loop(void* arg) { /* ParallelWorker.cs:67-91 */
  or_slot* s = (or_slot*)arg;
  pthread_mutex_lock(&s->mu);
  for (;;) {
    while (!s->pending && !s->cancel) pthread_cond_wait(&s->enqueued, &s->mu);
    if (s->cancel) break;
    or_action a = s->action;
    void* ctx = s->ctx;
    s->pending = 0;
    pthread_mutex_unlock(&s->mu);
    a(s->index, ctx);
    pthread_mutex_lock(&s->mu);
    s->finished = 1;
    pthread_cond_signal(&s->done);
  }
  pthread_mutex_unlock(&s->mu);
  return NULL;
}

or_worker* or_worker_create(int num_threads) { /* ParallelWorker.cs:34-56 */
  or_worker* w = (or_worker*)calloc(1, sizeof(*w));
  w->num_threads = num_threads;
  w->slots = (or_slot*)calloc((size_t)num_threads, sizeof(or_slot));
  for (int i = 0; i < num_threads; i++) {
    or_slot* s = &w->slots[i];
    s->index = i;
    pthread_mutex_init(&s->mu, NULL);
    pthread_cond_init(&s->enqueued, NULL);
    pthread_cond_init(&s->done, NULL);
    pthread_create(&s->thread, NULL, work_loop, s);
  }
  return w;
}

void or_worker_destroy(or_worker* w) { /* ParallelWorker.cs:132-139 */
  if (!w) return;
  for (int i = 0; i < w->num_threads; i++) {
    or_slot* s = &w->slots[i];
    pthread_mutex_lock(&s->mu);
    s->cancel = 1;
    pthread_cond_signal(&s->enqueued);
    pthread_mutex_unlock(&s->mu);
    pthread_join(s->thread, NULL);
    pthread_mutex_destroy(&s->mu);
    pthread_cond_destroy(&s->enqueued);
    pthread_cond_destroy(&s->done);
  }
  free(w->slots);
  free(w);
}

static void worker_work(or_worker* w, or_action a, void* ctx) { /* ParallelWorker.cs:98-117, waitResult = true */
  for (int i = 0; i < w->num_threads; i++) {
    or_slot* s = &w->slots[i];
    pthread_mutex_lock(&s->mu);
    s->action = a;
    s->ctx = ctx;
    s->finished = 0;
    s->pending = 1;
    pthread_cond_signal(&s->enqueued);
    pthread_mutex_unlock(&s->mu);
  }
  for (int i = 0; i < w->num_threads; i++) { /* WaitHandle.WaitAll */
    or_slot* s = &w->slots[i];
    pthread_mutex_lock(&s->mu);
    while (!s->finished) pthread_cond_wait(&s->done, &s->mu);
    pthread_mutex_unlock(&s->mu);
  }
}

typedef struct {
  const or_holemap* m;
  const float* points;
  int n;
  const float* search_pose;
  const float* offsets;
  int iterations;
  float (*poses)[3];
  int32_t* distances;
} search_ctx;

static void search_action(int index, void* vctx) { /* CoreSLAMProcessor.cs:680-689 */
  search_ctx* c = (search_ctx*)vctx;
  or_monte_carlo_search(c->m, c->points, c->n, c->search_pose, c->offsets + (size_t)3 * index * c->iterations,
                        c->iterations, c->poses[index], &c->distances[index], NULL);
}

void or_parallel_search_mt(or_worker* w, const or_holemap* m, const float* points, int n,
                           const float search_pose[3], const float* offsets, int iterations,
                           float best_pose[3], int32_t* best_distance) {
  int T = w->num_threads;
  float(*poses)[3] = (float(*)[3])malloc(sizeof(float[3]) * (size_t)T);
  int32_t* distances = (int32_t*)malloc(sizeof(int32_t) * (size_t)T);
  search_ctx ctx = {m, points, n, search_pose, offsets, iterations, poses, distances};
  worker_work(w, search_action, &ctx);
  int32_t bestd = INT32_MAX; /* :695-705 */
  float best[3] = {search_pose[0], search_pose[1], search_pose[2]};
  for (int i = 0; i < T; i++) {
    if (distances[i] < bestd) {
      bestd = distances[i];
      best[0] = poses[i][0]; best[1] = poses[i][1]; best[2] = poses[i][2];
    }
  }
  free(poses);
  free(distances);
  *best_distance = bestd;
  best_pose[0] = best[0]; best_pose[1] = best[1]; best_pose[2] = best[2];
}

/* host libm over arrays: what MathF.Cos / MathF.Sin / NormalizeAngle give on this machine */
void or_libm_sincos(const float* in, int64_t n, float* cos_out, float* sin_out) {
  for (int64_t i = 0; i < n; i++) {
    cos_out[i] = cosf(in[i]);
    sin_out[i] = sinf(in[i]);
  }
}

void or_normalize_angle_array(const float* in, int64_t n, float* out) {
  for (int64_t i = 0; i < n; i++) out[i] = or_normalize_angle(in[i]);
}

/* zlib CRC-32 (poly 0xEDB88320), for map checksums in KATs */
uint32_t or_crc32(const void* data, uint64_t nbytes) {
  static uint32_t table[256];
  static int init = 0;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
      table[i] = c;
    }
    init = 1;
  }
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = 0xFFFFFFFFu;
  for (uint64_t i = 0; i < nbytes; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
