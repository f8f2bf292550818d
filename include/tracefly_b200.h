/* tracefly_b200.h — C-ABI of the B200-native TraceFly GA hot path.
 *
 * One shared library (libtracefly_b200.so, sm_100a only, no CPU fallback) whose
 * entry points are what a binding for the reference's GA path would call.  All
 * signatures are plain C: pointers, sizes, PODs.  No torch / C++ types.
 *
 * Reference files (under the upstream repo root):
 *   HR  = UAV_Navigation/path_planning_GA/path_planner_hrmaga.py
 *   PF  = UAV_Navigation/path_planning_GA/GA_complex _PF.py
 *   GAE = uav-backend/app/routers/GA_enhanced.py
 *
 * Conventions
 *   - every `*_dev` pointer is CALLER-OWNED device memory (e.g. a torch tensor's
 *     data_ptr()), valid until `stream` reaches the operation; the library never
 *     allocates per call (scratch is sized by tf_*_workspace_bytes and passed in).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     every call is asynchronous on that stream.
 *   - return value: TF_OK (0) or a negative tf_status; never throws, never exits.
 *   - population layout: f32[P][W][4] (x, y, z, pad) — one waypoint = one float4 —
 *     or, for the reference-order f64 entry points, f64[P][W][4].
 *   - lens_dev: i32[P] valid waypoint count per chromosome (<= W), or NULL = all W.
 *   - collision flags: u32[P][ceil(W/32)], bit i of word k = waypoint 32k+i collides.
 */
#ifndef TRACEFLY_B200_H
#define TRACEFLY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_VERSION 100 /* 0.1.0 */

typedef enum tf_status {
    TF_OK = 0,
    TF_ERR_INVALID = -1,     /* NULL pointer / bad enum / bad flag            */
    TF_ERR_SHAPE = -2,       /* P, W, N, k out of the supported range         */
    TF_ERR_UNSUPPORTED = -3, /* combination not built                          */
    TF_ERR_CUDA = -4,        /* a CUDA runtime call failed (see tf_last_cuda_error) */
    TF_ERR_NOMEM = -5,
    TF_ERR_NODEVICE = -6,    /* no sm_100 device: there is no fallback path    */
    TF_ERR_WORKSPACE = -7    /* workspace too small                            */
} tf_status;

#define TF_W_MAX 256   /* waypoints per chromosome supported by the warp-per-chromosome kernels */
#define TF_TOPK_MAX 8192

/* ---- environment ---------------------------------------------------------------- */

/* Obstacle semantics: which reference collision / distance rules apply. */
typedef enum tf_semantics {
    TF_SEM_NONE = 0,
    TF_SEM_HR = 1, /* HR:17-28  sphere test ||p-c|| <= size + safety_radius; nearest distance to CENTRES */
    TF_SEM_PF = 2  /* PF:28-59  cube / sphere / cylinder(=AABB s,s,2s) collision; distances per PF:46-59 */
} tf_semantics;

typedef enum tf_kind {
    TF_KIND_SPHERE = 0,   /* PF:37-39, PF:54-55 */
    TF_KIND_CUBE = 1,     /* PF:34-36 collision = Chebyshev <= size; PF:52-53 distance is spherical */
    TF_KIND_CYLINDER = 2, /* PF:40-42 collision = AABB (s,s,2s); PF:56-57 distance = max(0,|dz|-2s) */
    TF_KIND_AABB = 3      /* extension: half extents (hx,hy,hz), true box distance */
} tf_kind;

/* 32-byte host record; `size` < 0 means "absent" -> HR default 1.0 (HR:20). */
typedef struct tf_obstacle {
    float cx, cy, cz, size;
    float hx, hy, hz; /* TF_KIND_AABB only */
    int32_t kind;     /* tf_kind; ignored under TF_SEM_HR */
} tf_obstacle;

/* Moving sphere (extension, BASELINE configs[3]): c(t) = c0 + v * t, t = waypoint index;
 * HR rules (radius = size + safety_radius, distance to centre). */
typedef struct tf_dyn_obstacle {
    float cx, cy, cz, size;
    float vx, vy, vz, _pad;
} tf_dyn_obstacle;

typedef struct tf_env tf_env; /* opaque; owns its device copy of the obstacle tables */

/* obs_host: n records on the HOST (copied; kind-sorted and pre-squared on the way in). */
int tf_env_create(int semantics, const tf_obstacle* obs_host, int n, float safety_radius, tf_env** out);
/* Same, from float64 records (n x 8 doubles: cx cy cz size hx hy hz kind).  The FP64
 * reference-order kernels then see positions / sizes / safety radius unrounded. */
int tf_env_create_f64(int semantics, const double* obs64_host, int n, double safety_radius, tf_env** out);
int tf_env_set_dynamic(tf_env* env, const tf_dyn_obstacle* dyn_host, int n, float safety_radius);
/* Optional acceleration of the static analytic obstacles (HR and PF rules): a uniform grid over
 * [box_lo, box_hi) whose `cell`-sized cells list (up to `cap`) the spheres / cubes / aabbs that can be
 * nearest to a point of the cell plus every obstacle that can collide with one, and — PF cylinders
 * measure distance along z only (PF:56-57) — a 1-D index of `z_cell`-high slabs listing the cylinders
 * that can be nearest.  A waypoint then visits tens of obstacles instead of all N, and the obstacle
 * tables are no longer staged in shared memory.  Results are bit-identical to the full loop; points
 * outside the box and overflowed lists fall back to it. */
int tf_env_build_static_index(tf_env* env, const float box_lo[3], const float box_hi[3], float cell, int cap,
                              float z_cell, int z_cap, void* stream);
int tf_env_static_index_stats(const tf_env* env, int64_t* n_lists, int64_t* n_overflow, double* mean_len,
                              int* max_len, double* z_mean_len);
/* Optional acceleration of the moving spheres: a uniform grid over [box_lo, box_hi) with `cell`-sized
 * cells, one copy per band of `band` consecutive integer times, whose entries list (up to `cap`) the
 * spheres that can be nearest to, or collide with, a point of the cell at a time of the band
 * (t < time_slices).  Results are bit-identical to the brute-force loop; points outside the box,
 * fractional times and overflowed cells fall back to it — so make the box generous: a warp waits for
 * its slowest lane, and one waypoint a few cells past the wall costs the whole warp the full loop. */
int tf_env_build_dynamic_index(tf_env* env, int time_slices, int band, const float box_lo[3],
                               const float box_hi[3], float cell, int cap, void* stream);
/* Diagnostics of the index: number of lists, overflowed lists, mean / max list length. */
int tf_env_dynamic_index_stats(const tf_env* env, int64_t* n_lists, int64_t* n_overflow, double* mean_len, int* max_len);
/* Bit-packed occupancy grid in BRICK layout (see tf_voxel_*): caller-owned device memory,
 * must stay alive while the env uses it.  nx, ny, nz multiples of 8.  NULL detaches. */
int tf_env_set_voxels(tf_env* env, const uint32_t* bricks_dev, int nx, int ny, int nz);
/* Hazard heat-map f32[nt][ny][nx] on the device; the env keeps its own (corner-packed, transposed) copy, so the
 * caller's buffer may be released after the call.  A waypoint i at (x, y)
 * samples tri-linearly at gx = x*sx + ox, gy = y*sy + oy, gt = i*st (clamped).  NULL detaches. */
int tf_env_set_hazard(tf_env* env, const float* hazard_dev, int nt, int ny, int nx,
                      float sx, float sy, float ox, float oy, float st);
int tf_env_destroy(tf_env* env);

/* Brick layout: the grid is tiled in 8x8x8-voxel bricks of 16 u32 (64 B); brick index =
 * bx + (nx/8) * (by + (ny/8) * bz); inside a brick word = (z&7)*2 + ((y&7)>>2),
 * bit = ((y&3)<<3) | (x&7).  One 32-B sector therefore covers 8x8x4 voxels. */
int64_t tf_voxel_grid_bytes(int nx, int ny, int nz);
/* Rasterise analytic obstacles (PF collision rules; HR if semantics == TF_SEM_HR with
 * radius size + safety_radius) into a zeroed brick grid: voxel (x,y,z) is set iff the
 * lattice point (x,y,z) collides.  obs_dev: tf_obstacle[n] on the DEVICE. */
int tf_voxelize(int semantics, const tf_obstacle* obs_dev, int n, float safety_radius,
                uint32_t* bricks_dev, int nx, int ny, int nz, void* stream);
/* LiDAR point cloud f32[n][4] (x, y, z, pad) on the device -> occupancy bits: each point marks voxel
 * int((p - origin) / voxel_size) — downsample_lidar's own arithmetic (lidar_compression.py:242-246: float64
 * division, truncation toward zero; origin = 0 there); the set bits are its unique voxels.  Voxels outside the
 * grid are dropped. */
int tf_voxelize_points(const float* points_dev, int64_t n, const double origin[3], double voxel_size,
                       uint32_t* bricks_dev, int nx, int ny, int nz, void* stream);
/* Linear bit grid (bit index x + nx*(y + ny*z), LSB first in u32 words) -> brick layout. */
int tf_voxel_pack_linear(const uint32_t* linear_dev, uint32_t* bricks_dev, int nx, int ny, int nz, void* stream);

/* ---- fitness (replaces Path.calculate_fitness HR:42-59 / PF:144-170) -------------- */

typedef struct tf_params {
    float w_safety, w_energy, w_wind, w_time, w_turn; /* HR:55-59: 0.4 0.2 0.2 0.1 0.1 */
    int32_t dist_zero_guard;    /* 1: time_score = 0 when distance == 0 (HR:53); 0: 1/0 = inf (PF:163) */
    float hazard_weight;        /* extension: wind = 1/(1 + alt + hazard_weight * sum hazard) ; 0 = off */
    int32_t samples_per_segment; /* extension: S >= 1 collision/proximity samples per segment; 1 = reference */
} tf_params;

void tf_params_default(tf_params* p);

/* FP32 throughput kernel, one warp per chromosome, fixed reduction order.
 * fit_dev f32[P]; coll_dev u32[P][ceil(W/32)] or NULL. */
int tf_fitness(const tf_env* env, const float* pop_dev, const int32_t* lens_dev, int P, int W,
               const tf_params* params, float* fit_dev, uint32_t* coll_dev, void* stream);

/* FP64 reference-order kernel: the reference's exact operation order (sequential sums,
 * sqrt(dot) norms), so that on identical inputs fitness is BIT-IDENTICAL to HR/PF in
 * float64 and the stable ranking cannot diverge.  pop64_dev f64[P][W][4]; fit64_dev f64[P]. */
int tf_fitness_f64(const tf_env* env, const double* pop64_dev, const int32_t* lens_dev, int P, int W,
                   const tf_params* params, double* fit64_dev, uint32_t* coll_dev, void* stream);

/* Point queries in the reference's float64 arithmetic: is_collision (HR:17-23) /
 * distance_to_nearest_obstacle (HR:25-28) / ComplexEnvironment.is_obstacle (PF:28-44) /
 * distance_to_obstacles (PF:46-59) for n points.  pts64_dev f64[n][4]; collide_dev i32[n] or
 * NULL; dnear_dev f64[n] or NULL (inf when the environment has no obstacles). */
int tf_point_query_f64(const tf_env* env, const double* pts64_dev, int n, int32_t* collide_dev,
                       double* dnear_dev, void* stream);

/* ---- ranking (replaces population.sort(key=fitness, reverse=True) HR:114 / PF:193) -- */

/* Stable descending top-k: order = (fitness desc, index asc); `worst` != 0 returns instead
 * the k LAST entries of that order, worst first.  k <= TF_TOPK_MAX.  idx_dev i32[k]. */
int64_t tf_topk_workspace_bytes(int P, int k, int f64);
int tf_rank_topk(const float* fit_dev, int P, int k, int worst, int32_t* idx_dev,
                 void* workspace_dev, int64_t workspace_bytes, void* stream);
int tf_rank_topk_f64(const double* fit64_dev, int P, int k, int worst, int32_t* idx_dev,
                     void* workspace_dev, int64_t workspace_bytes, void* stream);

/* The k best (worst != 0: k worst) chromosomes under the same stable order, as a SET in ASCENDING
 * INDEX order — no rank order, no sort: the victims of island migration.  Any k <= P. */
int64_t tf_select_set_workspace_bytes(int P);
int tf_select_set(const float* fit_dev, int P, int k, int worst, int32_t* idx_dev,
                  void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ---- selection + crossover + mutation (HR:85-99,114-122; PF:192-208) --------------- */

typedef enum tf_selection {
    TF_SEL_TRUNCATION = 0, /* HR:118: 2 distinct uniform picks from the sorted top n_top */
    TF_SEL_TOURNAMENT = 1, /* extension: best of tournament_k uniform entrants, per parent */
    TF_SEL_TOP2 = 2        /* PF:192-194 select_parents: every pair's parents are the two best, in rank order */
} tf_selection;
typedef enum tf_crossover {
    TF_CX_SINGLE_POINT = 0, /* HR:85-89: cp in [1, min(L1,L2)-2]; c1 = p1[:cp]+p2[cp:], c2 = p2[:cp]+p1[cp:] */
    TF_CX_UNIFORM = 1       /* extension: interior waypoints swapped by an iid bit mask */
} tf_crossover;
typedef enum tf_mutation {
    TF_MUT_LATTICE = 0,     /* HR:91-99: with prob rate, one interior waypoint += d, d in {-1,0,1}^3 */
    TF_MUT_LATTICE_PF = 1,  /* PF:202-208: any waypoint (endpoints included) += d, d in {-1,+1}^3 */
    TF_MUT_GAUSSIAN = 2     /* extension: one interior waypoint += sigma * z, z ~ N(0,1)^3 (Irwin-Hall 8) */
} tf_mutation;

typedef struct tf_ga_params {
    int32_t n_elite;       /* HR:115: 2 — first n_elite of the ranking are copied to slots 0.. */
    int32_t selection;     /* tf_selection */
    int32_t n_top;         /* truncation pool size (HR:118: 10) */
    int32_t tournament_k;  /* 1..32 */
    int32_t crossover;     /* tf_crossover */
    int32_t mutation;      /* tf_mutation */
    float mutation_rate;   /* HR:62: 0.1 */
    float sigma;           /* TF_MUT_GAUSSIAN */
    uint32_t stream_id;    /* Philox counter word 2: island / rank id */
    int32_t reserve_sms;   /* tf_generation only: SMs (0 = none, at most half) this launch of the persistent kernel
                              leaves unoccupied, so that kernels on other streams (elite selection, NCCL) run beside
                              it.  A launch argument, not library state: safe with several host threads / streams. */
} tf_ga_params;

void tf_ga_params_default(tf_ga_params* p);

/* Draw tape: host-supplied random draws, one 32-byte record per CHILD PAIR j (children
 * land in slots n_elite+2j, n_elite+2j+1; a slot >= P is dropped, HR:122).
 *   pa, pb : positions in the ranking (truncation) or population indices (tournament)
 *   cp     : crossover point
 *   mp1/2  : mutated waypoint of child 1/2, or -1 = gate failed (rand() >= rate)
 *   d1/d2  : packed delta, 2 bits per axis: (dx+1) | (dy+1)<<2 | (dz+1)<<4
 * This is exactly the information HR:118-120 draws from np.random, in that order. */
typedef struct tf_draw {
    int32_t pa, pb, cp, mp1, d1, mp2, d2, _pad;
} tf_draw;

/* The draws of HRMAGAGenerator.evolve (HR:117-120) for `generations` x `npairs` child pairs, taken from a COPY of
 * NumPy's legacy global generator exactly as the reference takes them (host function, no GPU work):
 * mt_key u32[624] / *mt_pos are np.random.get_state()[1:3] on entry and the state after the draws on return, so the
 * caller can np.random.set_state() what the reference would have consumed.  n_top = min(10, P) (HR:118), W >= 3. */
int tf_np_hr_tape(uint32_t* mt_key, int32_t* mt_pos, int generations, int npairs, int n_top, int W,
                  double mutation_rate, tf_draw* tape_out);

/* One GA step: next_pop <- elites + children.  tape_dev == NULL -> counter-based
 * Philox4x32-10 keyed by `seed`, counter (pair, generation, stream_id, call).
 * top_idx_dev: i32[max(n_elite, n_top)] from tf_rank_topk.  Optional parity outputs:
 * sel_dev i32[npairs][2] (population indices of both parents), cp_dev i32[npairs]. */
int tf_ga_step(const float* pop_dev, const int32_t* lens_dev, const float* fit_dev,
               const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
               const tf_ga_params* ga, int P, int W, float* next_pop_dev, int32_t* next_lens_dev,
               int32_t* sel_dev, int32_t* cp_dev, void* stream);
int tf_ga_step_f64(const double* pop64_dev, const int32_t* lens_dev, const double* fit64_dev,
                   const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
                   const tf_ga_params* ga, int P, int W, double* next_pop64_dev, int32_t* next_lens_dev,
                   int32_t* sel_dev, int32_t* cp_dev, void* stream);

/* Fused generation (FP32, fixed W): builds each child in registers, writes it to
 * next_pop and evaluates its fitness in the same warp — the child is never re-read.
 * Results are identical to tf_ga_step followed by tf_fitness. */
int tf_generation(const tf_env* env, const float* pop_dev, const float* fit_dev,
                  const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
                  const tf_ga_params* ga, const tf_params* params, int P, int W,
                  float* next_pop_dev, float* next_fit_dev, uint32_t* next_coll_dev,
                  int32_t* sel_dev, int32_t* cp_dev, void* stream);
/* tf_generation followed by the ranking of its children: next_top_dev i32[next_k] = what
 * tf_rank_topk(next_fit_dev, P, next_k, worst = 0) returns, i.e. the top_idx of the NEXT generation (HR:114 runs once
 * per generation; this removes its launch from the generation loop).  Fused into the persistent generation kernel for
 * next_k <= 2 on large populations (every warp keeps its two best children, the last warp of the grid merges them); any
 * other case runs tf_rank_topk behind the generation — same result.  workspace: tf_topk_workspace_bytes(P, next_k, 0). */
int tf_generation_ranked(const tf_env* env, const float* pop_dev, const float* fit_dev,
                         const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
                         const tf_ga_params* ga, const tf_params* params, int P, int W,
                         float* next_pop_dev, float* next_fit_dev, uint32_t* next_coll_dev,
                         int32_t* sel_dev, int32_t* cp_dev, int32_t* next_top_dev, int next_k,
                         void* workspace_dev, int64_t workspace_bytes, void* stream);

/* HR:126-128 best tracking: if fit[slot] > *best_fit then { *best_fit = fit[slot];
 * best_path <- pop[slot]; *best_gen = generation }. */
int tf_track_best(const float* pop_dev, const float* fit_dev, int slot, int W, uint32_t generation,
                  float* best_fit_dev, float* best_path_dev, int32_t* best_gen_dev, void* stream);
int tf_track_best_f64(const double* pop64_dev, const double* fit64_dev, int slot, int W, uint32_t generation,
                      double* best_fit_dev, double* best_path_dev, int32_t* best_gen_dev, void* stream);

/* ---- initial populations --------------------------------------------- */

/* HR:68-80 diagonal lattice walk start -> end, replicated P times (HR:82-83).  Returns the
 * walk length through *w_out when pop_dev == NULL (sizing call, host only). */
int tf_init_diagonal(const double start[3], const double end[3], int P, int W, float* pop_dev,
                     int* w_out, void* stream);
int tf_init_diagonal_f64(const double start[3], const double end[3], int P, int W, double* pop64_dev,
                         int* w_out, void* stream);
/* Seeded +-1 lattice random walks (Philox) from per-chromosome or common start, clipped to
 * [lo, hi] per axis: the synthetic populations of BASELINE configs[1..4]. */
int tf_init_walks(uint64_t seed, const float* start_dev /* f32[P][4] or NULL */, const float start[3],
                  const float lo[3], const float hi[3], int P, int W, float* pop_dev, void* stream);

/* ---- island migration (either side of the NCCL all-gather) --------------- */

/* rows_dev f32[E][W][4], rfit_dev f32[E] <- pop[elite_idx[e]], fit[elite_idx[e]]. */
int tf_pack_elites(const float* pop_dev, const float* fit_dev, const int32_t* elite_idx_dev, int E, int W,
                   float* rows_dev, float* rfit_dev, void* stream);
/* gathered rows f32[world][E][W][4] / fit f32[world][E]; the j-th FOREIGN elite (source rank
 * ascending, skipping self_rank, then position in its rows) overwrites pop[worst_idx[j]], fit[...].
 * worst_idx: the (world-1)*E victims, e.g. from tf_select_set(..., worst = 1). */
int tf_replace_worst(float* pop_dev, float* fit_dev, const int32_t* worst_idx_dev, int W,
                     const float* gathered_rows_dev, const float* gathered_fit_dev,
                     int world, int self_rank, int E, void* stream);

/* ---- backend GA fitness --------------------------------------------------- */
/* GeneticAlgorithm.fitness_function of uav-backend/app/routers/GA_enhanced.py (GAE:20-26) for a ragged
 * population: fit_dev i32[P] = -len - #{i in [1, len-2] : path[i-1] != path[i+1]}. */
int tf_fitness_gae(const float* pop_dev, const int32_t* lens_dev, int P, int W, int32_t* fit_dev, void* stream);

/* Roulette selection of the backend GA (GAE:28-34: random.choices(population, [s / total for s in scores], k=2)),
 * in two steps so that one O(P) pass serves every draw of a generation:
 *   tf_roulette_cumweights  cum[i] = w_0 + ... + w_i with w_i = score_i / sum(score), accumulated LEFT TO RIGHT in
 *                           float64 exactly like itertools.accumulate (a sequential pass by design: a parallel scan
 *                           would round differently and could flip a bisect).  Scores and their sum are negative
 *                           integers, so worse paths get larger weights — the reference's quirk, kept.
 *   tf_select_roulette      idx[i] = bisect_right(cum, u[i] * (cum[P-1] + 0.0), 0, P - 1) for n uniform draws u
 *                           (what random.choices does with each random()). */
int tf_roulette_cumweights(const int32_t* score_dev, int P, double* cum_dev, void* stream);
int tf_select_roulette(const double* cum_dev, int P, const double* u_dev, int n, int32_t* idx_dev, void* stream);

/* ---- initial population of the shaped-obstacle GA ------------------------------ */
/* `PathPlanner.find_path(dodge_strategy, height_variation)` ("GA_complex _PF.py":93-132) for n paths at once,
 * one thread per path: from `source`, step by sign(destination - current) per axis (z biased by the path's
 * height variation, PF:111-116); if the next point is an obstacle (PF:28-44 rules of `env`, evaluated in
 * float64 like the reference) try the single dodge move instead, or stay (PF:118-129); stop at the
 * destination or when a point repeats (PF:105-106).  `GeneticAlgorithm.initialize_population` (PF:181-188)
 * is one call with n = population_size and the height variations it draws.
 * paths64_dev: f64[n][cap][4] (x, y, z, 0), the layout tf_fitness_f64 reads; lens_dev: i32[n].
 * A path that needs more than `cap` waypoints is cut at cap and counted in *overflow_dev (i32, device,
 * caller-zeroed, may be NULL): the caller retries with a larger cap. */
enum { TF_DODGE_UP = 0, TF_DODGE_DOWN = 1, TF_DODGE_LEFT = 2, TF_DODGE_RIGHT = 3, TF_DODGE_FORWARD = 4, TF_DODGE_BACKWARD = 5 };
int tf_find_path_pf(const tf_env* env, const double source[3], const double destination[3],
                    const int32_t* height_variation_dev, int n, int dodge, int cap, double* paths64_dev,
                    int32_t* lens_dev, int32_t* overflow_dev, void* stream);

/* ---- initial population of the backend GA --------------------------------------- */
/* `UAVPathPlanner.find_path()` (GA_enhanced.py GAE:166-179: straight lines GAE:143-151, random one-step dodge
 * GAE:153-164, voxel-set membership GAE:140-141) for n_paths paths, which is `GeneticAlgorithm.initialize_population`
 * (GAE:13-18).  The reference draws one random.shuffle per dodge from Python's global generator, path after path, so
 * the paths are walked IN SEQUENCE by one warp (lanes share the membership scan) and consume shuffles_dev — packed
 * orders from tf_py_shuffle6_tape — in that order.  obstacles_dev i32[n_obs][4] (x, y, z, 0); bounds = xlo xhi ylo yhi
 * zlo zhi (inclusive, GAE:132-138).  paths_dev i32[n_paths][cap][4]; lens_dev i32[n_paths];
 * result_dev i32[2] = {shuffles consumed, status: 0 ok / 1 cap too small / 2 tape exhausted}. */
int tf_find_path_gae(const int32_t* obstacles_dev, int n_obs, const int32_t bounds[6], const int32_t source[3],
                     const int32_t destination[3], const int32_t* shuffles_dev, int n_shuffles, int n_paths, int cap,
                     int32_t* paths_dev, int32_t* lens_dev, int32_t* result_dev, void* stream);
/* n shuffles of the six dodge moves, taken from a COPY of CPython's global generator (random.getstate()[1]: 624 key
 * words, then the position) exactly as random.shuffle takes them (host function).  out[i]: the order, 3 bits per
 * position.  mt_key / *mt_pos are updated to the state after the draws. */
int tf_py_shuffle6_tape(uint32_t* mt_key, int32_t* mt_pos, int n, int32_t* out);

/* ---- compact host format -------------------------------------------------- */
/* int16[n][4] (x, y, z, pad) <-> f32[n][4]: lattice chromosomes at 8 B per waypoint for the PCIe leg.
 * Packing rounds to nearest and saturates; it is exact for integer coordinates in [-32768, 32767]. */
int tf_unpack_i16(const int16_t* in_dev, float* pop_dev, int64_t n_waypoints, void* stream);
/* f32[n][3] (x, y, z) <-> f32[n][4]: the reference's waypoints are 3-tuples (HR:33-36), so a caller's array
 * form is [P][W][3]; 12 B per waypoint on the PCIe leg, the pad lane is written as 0. */
int tf_unpack_xyz(const float* in_dev, float* pop_dev, int64_t n_waypoints, void* stream);
int tf_pack_xyz(const float* pop_dev, float* out_dev, int64_t n_waypoints, void* stream);
int tf_pack_i16(const float* pop_dev, int16_t* out_dev, int64_t n_waypoints, void* stream);

/* Lossy front ends of the reference's waypoint compression (UAV_Navigation/data_compression/
 * ga_waypoints_compression.py), applied per path to paths64_dev f64[n_paths][W][4] (what tf_track_best_f64 keeps; use
 * f32 = 1 for f32[n_paths][W][4] populations, widened to float64 first) — entropy coding (zlib / LZMA) stays on the host:
 *   TF_CODEC_QUANT_F32   compress_waypoints (:2090-2093): np.round(w / q) * q as float32; param = q (0.05)
 *                        out_dev f32[n_paths][W][3]   (host: zlib.compress(out.tobytes()) == the reference's bytes)
 *   TF_CODEC_DELTA_ROUND delta_encode + quantize (:875-884): np.round(np.diff(w, axis=0, prepend=w[:1]), d);
 *                        param = d decimals (3)       out_dev f64[n_paths][W][3]   (host: struct.pack + lzma) */
enum { TF_CODEC_QUANT_F32 = 0, TF_CODEC_DELTA_ROUND = 1 };
int tf_encode_waypoints(const void* paths_dev, int f32, int64_t n_paths, int W, int mode, double param, void* out_dev,
                        void* stream);

/* ---- misc ------------------------------------------------------------------------------ */
int tf_version(void);
const char* tf_strerror(int status);
int tf_last_cuda_error(void);               /* cudaError_t of the last failing runtime call on this thread */
int tf_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes);
/* Philox4x32-10 block, exposed so host code and tests can reproduce device draws. */
void tf_philox4x32_10(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                      uint32_t out[4]);
/* Self-test of the library's FP32 square root (rsqrt seed + fma Newton step, used for segment lengths and
 * obstacle distances; the FP32 model of HR:51-52 / PF:162 needs it correctly rounded): adds to
 * *mismatches_dev (u64, device, caller-zeroed) the number of bit patterns in [first_bits, first_bits + count)
 * whose result differs from sqrt.rn.f32 (scalar and batched form).  count <= 2^32. */
int tf_selftest_sqrt(uint32_t first_bits, uint64_t count, uint64_t* mismatches_dev, void* stream);
/* Self-test of the library's FP32 reciprocal (MUFU.RCP seed + fma Newton step for 1 <= y < 2^126, div.rn elsewhere;
 * the proximity term 1 / (1 + d) of HR:46 in the FP32 model): adds to *mismatches_dev the number of bit patterns y in
 * [first_bits, first_bits + count) for which it differs from div.rn.f32(1, y) (each pattern is checked through the scalar
 * and the batched form).  count <= 2^32. */
int tf_selftest_rcp(uint32_t first_bits, uint64_t count, uint64_t* mismatches_dev, void* stream);
/* Options.  TF_OPT_DISABLE_STREAM = 1 forces the LSU-path kernels even where the TMA-staged
 * kernels apply (fixed length W <= 256; with an obstacle table W = 32 / 64); both give bit-identical results. */
#define TF_OPT_DISABLE_STREAM 1
/* TF_OPT_STREAM_SCHEDULE: how the TMA-staged kernels walk a tile.  1 = one chromosome per warp step,
 * 2 = two (one per half-warp, dynamic tile tickets), 0 (default) = the library picks per kernel variant.
 * Bit-identical results either way; exposed for A/B measurements and the equivalence tests. */
#define TF_OPT_STREAM_SCHEDULE 2
/* TF_OPT_RESERVE_SMS: process-wide DEFAULT of tf_ga_params.reserve_sms (the larger of the two applies).  Prefer
 * the per-launch field: an option toggled around a launch races with other host threads. */
#define TF_OPT_RESERVE_SMS 3
int tf_set_option(int option, int value);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
int64_t tf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TRACEFLY_B200_H */
