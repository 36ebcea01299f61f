/* coreslam_oracle.h — CPU restatement of SLAM.NET's CoreSLAM scan-to-map hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under slam.net_b200/ links, imports or calls this; it is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED against an executed reference: the reference is C#/.NET 6 and no .NET runtime
 * exists in this environment, and the reference ships no tests, golden vectors or fixtures for
 * this path.  The restatement is instead pinned by (a) the hand-derived
 * known-answer vectors KAT-A..E of SURVEY.md section 8c (tests/test_oracle_kat.py) and (b) an
 * independent statement-by-statement Python transliteration (oracle/transliteration.py) that
 * must agree with it bit-for-bit on randomized inputs (tests/test_oracle_cross.py).
 *
 * All file:line citations are relative to /root/reference (mikkleini/slam.net @ 7abc587).
 */
#ifndef CORESLAM_ORACLE_H
#define CORESLAM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_TS_NO_OBSTACLE 65500 /* CoreSLAM/CoreSLAMProcessor.cs:21 */
#define OR_TS_OBSTACLE 0        /* CoreSLAM/CoreSLAMProcessor.cs:22 */

/* CoreSLAM/HoleMap.cs:17-37 */
typedef struct {
  int size;         /* Size  (pixels per edge) */
  float scale;      /* Scale = sizePixels / sizeMeters (HoleMap.cs:20) */
  uint16_t* pixels; /* Pixels[size*size], row-major y*size+x */
} or_holemap;

or_holemap* or_holemap_create(int size_pixels, float size_meters);
void or_holemap_destroy(or_holemap* m);
/* HoleMap.GetPackedPixels, HoleMap.cs:44-55; out has size*size/2 bytes */
void or_holemap_packed(const or_holemap* m, uint8_t* out);

/* (int)float as .NET 6 x64 compiles it (cvttss2si) */
int32_t or_cvt(float f);

/* MathEx.NormalizeAngle, BaseSLAM/MathEx.cs:116-138 */
float or_normalize_angle(float angle);

/* CalculateDistanceSISD, CoreSLAMProcessor.cs:226-259.  points = n * (x,y) */
int32_t or_distance(const or_holemap* m, const float* points, int n, const float pose[3]);

/* ClipRay, CoreSLAMProcessor.cs:320-345 */
int or_clip_ray(int size, int32_t* xyc, int32_t* yxc, int32_t xy, int32_t yx);

/* DrawLaserRayOnHoleMap, CoreSLAMProcessor.cs:359-443.  Returns the number of cells written.
 * If trace != NULL it receives up to trace_cap triples (cell index, pixval, new value). */
int64_t or_draw_ray(or_holemap* m, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int32_t xp, int32_t yp,
                    int32_t value, int32_t alpha, int32_t* trace, int trace_cap);

/* UpdateHoleMap, CoreSLAMProcessor.cs:496-534.  Returns total cells written (visits).
 * rays_out (optional, n*6 ints) receives x1,y1,x2,y2,xp,yp per point. */
int64_t or_update_hole_map(or_holemap* m, const float* points, int n, const float pose[3], float hole_width,
                           int quality, int32_t* rays_out);

/* MonteCarloSearch, CoreSLAMProcessor.cs:624-653.  offsets = iterations * (dx,dy,dtheta), i.e. the
 * values the reference dequeues in the order X, Y, Theta (:633-638).  distances_out (optional)
 * receives iterations+1 values: [0] = searchPose, [1+i] = candidate i. */
void or_monte_carlo_search(const or_holemap* m, const float* points, int n, const float search_pose[3],
                           const float* offsets, int iterations, float best_pose[3], int32_t* best_distance,
                           int32_t* distances_out);

/* ParallelMonteCarloSearch, CoreSLAMProcessor.cs:674-710, run serially: thread t gets
 * offsets[t*iterations .. (t+1)*iterations).  distances_out (optional) receives
 * 1 + threads*iterations values in the flat order of SURVEY.md section 8a6.
 * best_index_out (optional): flat index of the winner (0 = searchPose). */
void or_parallel_search(const or_holemap* m, const float* points, int n, const float search_pose[3],
                        const float* offsets, int iterations, int threads, float best_pose[3],
                        int32_t* best_distance, int32_t* distances_out, int32_t* best_index_out);

/* ScanSegmentsToCloud, CoreSLAMProcessor.cs:187-207.  One segment: rays = n * (angle, radius). */
void or_segment_to_cloud(const float* rays, int n, const float segment_pose[3], const float odometry_pose[3],
                         float* points_out);

/* ---- ObstacleMap half of Update ------------------------------------------------- */
/* CoreSLAM/ObstacleMap.cs:11-44: sbyte[size,size] (first index Y), Scale = sizePixels / sizeMeters.
 * no_hit is CoreSLAMProcessor's private bool[,] noHitMap (:30, :133). */
typedef struct {
  int size;
  float scale;
  int8_t* pixels;  /* row-major y*size+x */
  uint8_t* no_hit; /* row-major y*size+x, 0/1 */
} or_obstaclemap;

or_obstaclemap* or_obstaclemap_create(int size_pixels, float size_meters);
void or_obstaclemap_destroy(or_obstaclemap* m);
/* DrawLaserRayOnObstacleMap, CoreSLAMProcessor.cs:456-490.  Returns the number of loop iterations that
 * touched the map (no-hit marks + the hit).  A ray whose Math.Abs argument is int.MinValue (the reference
 * throws OverflowException there) is skipped and returns -1. */
int64_t or_draw_ray_obstacle(or_obstaclemap* m, int32_t x1, int32_t y1, int32_t x2, int32_t y2, int max_obstacle_hits);
/* UpdateObstacleMap, CoreSLAMProcessor.cs:540-593.  Returns the cells touched by the rays (sum of the above). */
int64_t or_update_obstacle_map(or_obstaclemap* m, const float* points, int n, const float pose[3],
                               int max_obstacle_hits);

/* ---- CoreSLAMProcessor state machine: ctor :119-162, Reset :167-175, Update :717-752 ---- */
typedef struct {
  or_holemap* map;
  float physical_map_size;
  float start_pose[3];
  float sigma_xy, sigma_theta;
  int iterations_per_thread, num_search_threads;
  int quality;                   /* :82, default 50 */
  float hole_width;              /* :87, default 0.6f */
  int position_search_beginning; /* :92, default 5 */
  float pose[3];                 /* :106 */
  float last_odometry_pose[3];   /* :35 */
  int scan_count;                /* :34 */
  int64_t visits;                /* cells written by the last Update (measurement only) */
  int32_t last_distance;         /* winner's distance of the last search, INT32_MAX if none */
  int32_t last_index;
  or_obstaclemap* omap;          /* :53, NULL when created without an obstacle map */
  int unmapped_obstacle_hits;    /* :98, default -5 (sbyte) */
  int max_obstacle_hits;         /* :103, default 10 (sbyte) */
  int64_t obstacle_visits;       /* map cells touched by the rays of the last UpdateObstacleMap */
} or_processor;

or_processor* or_processor_create(float physical_map_size, int hole_map_size, const float start_pose[3],
                                  float sigma_xy, float sigma_theta, int iterations_per_thread,
                                  int num_search_threads);
/* the full constructor (:119-120): obstacle_map_size > 0 also creates the ObstacleMap, and Update then ends
 * with UpdateObstacleMap (:751) */
or_processor* or_processor_create_full(float physical_map_size, int hole_map_size, int obstacle_map_size,
                                       const float start_pose[3], float sigma_xy, float sigma_theta,
                                       int iterations_per_thread, int num_search_threads);
void or_processor_destroy(or_processor* p);
void or_processor_reset(or_processor* p);
/* Update with an already-built cloud (points in the lidar frame relative to odometry_pose) and the
 * T*I candidate offsets that the reference would have dequeued for this scan (ignored while
 * scan_count < position_search_beginning).  UpdateObstacleMap (:751) runs when the processor has an
 * obstacle map (or_processor_create_full). */
void or_processor_update(or_processor* p, const float* points, int n, const float odometry_pose[3],
                         const float* offsets);

/* ---- threaded CPU baseline mirroring BaseSLAM/ParallelWorker.cs:34-117 ---- */
typedef struct or_worker or_worker;
or_worker* or_worker_create(int num_threads);
void or_worker_destroy(or_worker* w);
/* ParallelMonteCarloSearch on `threads` persistent workers (fan-out, WaitAll, serial arg-min). */
void or_parallel_search_mt(or_worker* w, const or_holemap* m, const float* points, int n,
                           const float search_pose[3], const float* offsets, int iterations,
                           float best_pose[3], int32_t* best_distance);
/* or_processor_update using the worker pool for the search */
void or_processor_update_mt(or_processor* p, or_worker* w, const float* points, int n,
                            const float odometry_pose[3], const float* offsets);

void or_libm_sincos(const float* in, int64_t n, float* cos_out, float* sin_out);
void or_normalize_angle_array(const float* in, int64_t n, float* out);
uint32_t or_crc32(const void* data, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
