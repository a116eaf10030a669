/*
 * uam_b200.h -- C-ABI of the B200-native path scorer for nomaporon/uam_path_planning.
 *
 * The reference has no FFI of its own (it is pure Python); this header is the boundary a
 * maintainer binds with ctypes from the reference's Python classes.  Each
 * entry point names the reference interface it replaces; file:line are relative to
 * geo_simulation_project/ in the reference tree.
 *
 * Conventions
 *   - plain C: pointers, sizes, doubles.  No C++ / torch types.
 *   - every call returns int: 0 = UAM_OK, negative = error; never throws, never aborts.
 *     uam_last_error(ctx) gives the message of the last failing call on that ctx.
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensor data_ptr());
 *     pointers named h_* (and all small parameter tables) are HOST pointers.
 *   - `stream` is the caller's cudaStream_t passed as void* (NULL = CUDA's default stream, as in
 *     every CUDA API).  Device-pointer calls are asynchronous and stream-ordered; host-buffer
 *     calls (*_host) run on the ctx's own streams and return when the results are in the
 *     caller's host buffers.
 *   - a ctx is bound to one GPU and is not thread-safe.  No global state.  There is no CPU
 *     fallback: without a CUDA device uam_ctx_create fails with UAM_ERR_CUDA.
 *
 * Layouts (kept from the reference)
 *   paths   z_  : (B, 2*(N+2)) float64, C-contiguous, interleaved [xs,ys,x1,y1,...,xN,yN,xg,yg]
 *                 = start + N free waypoints + goal      (path_generation/solver.py:59-66)
 *   params  p   : [ms_x, ms_y, mg_x, mg_y, maxratio, maxalpha, enlargement, w_0 .. w_{R-1}]
 *                 (solver.py:60-68; p[0:4] = map.x_start / map.x_goal as used by
 *                 Problem.length_of, problem.py:137-140; weights in region insertion order)
 *   rasters     : (L, H, W) float32 C-order, row <-> y, col <-> x, affine (x0, dx, y0, dy),
 *                 cell centre (x0 + (j+1/2) dx, y0 + (i+1/2) dy); occupancy (H, W) uint8
 *                 (rasterio band layout as read at map_generation/data_manager.py:13)
 */
#ifndef UAM_B200_H
#define UAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uam_ctx uam_ctx;

enum {
    UAM_OK = 0,
    UAM_ERR_INVALID = -1,      /* bad argument */
    UAM_ERR_CUDA = -2,         /* CUDA runtime error (message in uam_last_error) */
    UAM_ERR_NOMEM = -3,
    UAM_ERR_STATE = -4,        /* required map state missing (no shapes / no raster uploaded) */
    UAM_ERR_UNSUPPORTED = -5
};

/* option flags = Problem.options (path_generation/problem.py:12-17) */
enum {
    UAM_LENGTH_SMOOTH   = 1 << 0,
    UAM_PENALTY_SMOOTH  = 1 << 1,
    UAM_OBSTACLE_SMOOTH = 1 << 2,
    UAM_MAXRATIO_SMOOTH = 1 << 3,
    /* the length term's first pair is (map.x_start, z_0) (problem.py:137-145).  With this flag the
     * scorer uses each path's own z_0 instead of p[0:2], i.e. the term is 0 -- for batches of
     * independent start/goal queries. */
    UAM_OWN_START       = 1 << 4
};

/* inequality record kinds of uam_map_set_shapes: 8 doubles per record, rec[0] = kind */
enum {
    UAM_EDGE_LINE = 0,     /* rec = {0, Ax, Ay, Bx-Ax, By-Ay, sgn, 0, 0}   h = -sgn*((By-Ay)(x-Ax) - (Bx-Ax)(y-Ay))
                              path_generation/polygon.py:69-71,98 */
    UAM_EDGE_ELLIPSE = 1,  /* rec = {1, cx, cy, r1, r2, 0, 0, 0}           h = ((x-cx)/r1)^2 + ((y-cy)/r2)^2 - 1
                              path_generation/ball.py:33-37 */
    UAM_EDGE_BOX = 2       /* rec = {2, axis, sign, c, r, 0, 0, 0}         h = sign>0 ? x_d - c - r : -x_d + c - r
                              path_generation/square.py:29-51 */
};

/* ---- context ------------------------------------------------------------------------------- */
int uam_ctx_create(int device, uam_ctx** ctx);
int uam_ctx_destroy(uam_ctx* ctx);
const char* uam_last_error(const uam_ctx* ctx);
const char* uam_version(void);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int uam_launch_count(const uam_ctx* ctx, uint64_t* n);
/* wait for the ctx's own streams */
int uam_sync(uam_ctx* ctx);
/* tuning knobs (defaults are the measured-best settings) */
enum {
    UAM_OPT_RASTER_LAYOUT = 1,        /* texel layout used by the next uam_map_set_raster*: 0 row-major, 1 tiled */
    UAM_OPT_INTEGRAL_VARIANT = 2,     /* integral mode: -1 auto (default: 2 for batches of >= 2^18 segments, else 0),
                                         0 warp per path / lane per sample, 1 lane pair per sample,
                                         2 segments binned by raster tile, warp per 32 sorted segments,
                                         3 segments cut at 64 x 64-cell tile boundaries, pieces sorted by tile, the
                                           tile staged in shared memory by one bulk async copy (cp.async.bulk) */
    UAM_OPT_L2_FETCH_GRANULARITY = 3, /* cudaLimitMaxL2FetchGranularity for this device: 32, 64 or 128 bytes */
    UAM_OPT_TIME_KERNELS = 4,         /* 1: bracket the dominant raster-scoring kernel of every device-pointer call with
                                         CUDA events on the caller's stream (resets the statistics) */
    UAM_OPT_GRID_DELTA = 6,           /* grid search: width of the distance window relaxed per round (0 = automatic: the cost
                                         of crossing one 32-cell tile at the mean cell cost); ordering only, same results */
    UAM_OPT_COMBINE_LAYERS = 5,       /* 1 (default): large-batch integral mode samples "quad texels": the layers are folded
                                         into ONE layer sum_l w_l * layer_l (the penalty is linear in the layers) and every
                                         cell stores its 2 x 2 bilinear footprint as one float4, so a tap is one 16-byte
                                         load; occupancy flags ride in the sign bits when all values are >= 0, else in a
                                         bit-plane.  Rebuilt when the weights or the raster change.  2: as 1 but always the
                                         bit-plane form.  0: always sample every layer texel by texel */
    UAM_OPT_HOST_CHUNKS = 7,          /* uam_score_paths_raster_host: chunks per call flowing through the copy / score
                                         pipeline (0 = default 4) */
    UAM_OPT_HOST_TAPER = 8,           /* ... chunk sizes change linearly from the first to the last chunk: t > 0 the last chunk
                                         is t percent smaller than the first, t < 0 the first is |t| percent smaller than the
                                         last, 0 (default) equal chunks; -95 .. 95 */
    UAM_OPT_RASTERIZER = 10,          /* 1 (default): uam_rasterize_occupancy / uam_rasterize_layers find, per raster row and shape, the
                                         interval of cells inside the shape by bisection with the exact fp64 predicate (the value of
                                         an inequality along a row is monotone in the column) and only touch those cells; 0: every
                                         cell of every surviving tile is evaluated (round-1 kernels); 2: as 1 with the tile form of the
                                         layer kernel (row intervals by bisection inside 16 x 64 tiles); 3: as 1 with the sampled row
                                         form of the layer kernel (coarse spans from 32 sample columns per row).  Same bits in every mode */
    UAM_OPT_CCL_TILES = 11,           /* 1 (default): uam_label_components labels 32 x 32-cell tiles in shared memory first and unites only
                                         the pairs across tile borders in HBM; 0: global union-find over all cells (round 1).  Same labels */
    UAM_OPT_GRID_GRAPH = 12,          /* 1 (default): the relaxation rounds of uam_grid_search* are looped on the device -- one CUDA graph
                                         with a WHILE conditional node whose body is a round (min key, selection, tile relaxation) and a
                                         one-thread kernel that re-arms the loop while tiles are pending; 0: the host enqueues the rounds
                                         and reads the active count back every 8 rounds.  Same results */
    UAM_OPT_GRID_HALF_CAP = 13,       /* grid search: half sweeps (32 row steps) a tile activation may run before the tile is handed to the
                                         next round with what it has (default 4 = two double sweeps; 0: to the fixed point).  Scheduling
                                         only, same results */
    UAM_OPT_SHAPE_GRID = 9            /* 1 (default): the analytic scorer / point queries look up per-cell candidate lists over
                                         the shapes (a shape with one inequality > max(e, 1e-14) on a whole cell contributes
                                         exact zeros there and is left out; same bits).  0: every shape at every point */
};
/* statistics of UAM_OPT_TIME_KERNELS: mean device time (ms) of the dominant scoring kernel (uam_k_score_tiles / uam_k_score_groups /
 * uam_k_score_raster_int / uam_k_score_raster_wp) over the timed calls, and their number */
enum { UAM_STAT_SCORE_KERNEL_MS_MEAN = 1, UAM_STAT_SCORE_KERNEL_COUNT = 2,
       /* counted work of the last uam_grid_search*: tile activations, double sweeps (one = 64 row steps of 32 cells, each
          relaxing the 8 in-plane edges of its cells), relaxation rounds */
       UAM_STAT_GRID_ACTIVATIONS = 3, UAM_STAT_GRID_SWEEPS = 4, UAM_STAT_GRID_ROUNDS = 5,
       /* shape grid of the analytic scorer (UAM_OPT_SHAPE_GRID) as last built: cells (0 = none) and list entries in all */
       UAM_STAT_SHAPE_GRID_CELLS = 6, UAM_STAT_SHAPE_GRID_ITEMS = 7,
       /* what the host enqueued for the last uam_grid_search*: kernel launches + memsets + graph launches (UAM_OPT_GRID_GRAPH) */
       UAM_STAT_GRID_HOST_SUBMISSIONS = 8 };
int uam_ctx_get_stat(uam_ctx* ctx, int stat, double* value);
int uam_ctx_set_option(uam_ctx* ctx, int option, int64_t value);

/* ---- map: shape tables ------------------------------------------------------------------------
 * Replaces the object graph RegionMap / Map / QuadraticObstacle / Function
 * (path_generation/region_map.py:8-61, map.py:7-43, quadratic_obstacle.py:8-39, function.py:119).
 *   h_edges        n_edges x 8 doubles, inequality records in reference order
 *   h_shape_off    n_shapes+1 prefix into h_edges
 *   h_shape_region n_shapes: -1 = hard obstacle (map.obstacles), else region index (insertion order)
 *   h_shape_center n_shapes x 2: QuadraticObstacle.center; NaN = none (penalty not normalised,
 *                  problem.py:76-77)
 * Obstacles keep insertion order among themselves (it is the order of get_nonlincon's blocks). */
int uam_map_set_shapes(uam_ctx* ctx, const double* h_edges, int n_edges, const int32_t* h_shape_off,
                       const int32_t* h_shape_region, const double* h_shape_center, int n_shapes,
                       int n_regions);

/* ---- map: rasters -------------------------------------------------------------------------------
 * Upload L float32 cost layers + uint8 occupancy; the library re-lays them out on the device
 * (texel-interleaved).  L in [1,3].  Host and device-pointer forms. */
int uam_map_set_raster(uam_ctx* ctx, const float* h_layers, int L, int H, int W, double x0, double dx,
                       double y0, double dy, const uint8_t* h_occupancy);
int uam_map_set_raster_device(uam_ctx* ctx, const float* d_layers, int L, int H, int W, double x0,
                              double dx, double y0, double dy, const uint8_t* d_occupancy, void* stream);

/* ---- candidates ---------------------------------------------------------------------------------------
 * Solver.create_x_init (path_generation/solver.py:103-136) for B displacements at once: d_z (B, 2(N+2)) float64
 * receives [start, arc through start and goal with sagitta d_b |goal-start|/2 (straight line for d_b = 0), goal].
 * h_ends = [x_start(2), x_goal(2)]; |d_b| <= 1 is the caller's check (the reference raises ValueError). */
int uam_make_arc_paths(uam_ctx* ctx, const double* h_ends, int N, const double* d_displacement, int64_t B,
                       double* d_z, void* stream);

/* The general form: row b of d_cand = {xs, ys, xg, yg, displacement} -- every candidate its own start / goal (a batch of
 * independent start/goal queries, 40 bytes each) -- and, for jitter_sigma > 0, N(0, sigma^2) added to both coordinates of
 * the N interior waypoints (start and goal stay fixed, as in the reference's flow create_x_init -> solve, main.py:160-171).
 * The normals come from a counter-based generator keyed by (seed, index0 + b, waypoint): a candidate depends on its global
 * index only, not on the batch split or the number of GPUs. */
int uam_make_candidates(uam_ctx* ctx, const double* d_cand, int64_t B, int N, double jitter_sigma, uint64_t seed,
                        uint64_t index0, double* d_z, void* stream);

/* ---- path scoring: analytic shapes -------------------------------------------------------------
 * Replaces, batched over B paths:
 *   d_cost[b]    = Problem.get_cost(z_)                      path_generation/problem.py:38-44
 *                  (length term incl. the dropped-last-segment behaviour of problem.py:39,140-145)
 *   d_collide[b] = any_j Map.collides(z_j)                   map.py:41-43, quadratic_obstacle.py:89-94
 *   d_g[b,:]     = Problem.get_nonlincon(z_)  (nullable)     problem.py:84-114, length 3N + n_obs(N+2)
 * All fp64 in the reference's operation order (no FMA contraction): inequalities, collision flags
 * and the zero pattern of g are bit-exact; costs differ from the reference only through the order
 * in which the per-waypoint terms are summed (warp-shuffle tree), ~1e-15 relative.
 * uam_analytic_g_len gives the row length of d_g for this map and N. */
int uam_score_paths_analytic(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p,
                             int n_p, int flags, double* d_cost, uint8_t* d_collide, double* d_g,
                             void* stream);
int uam_score_paths_analytic_host(uam_ctx* ctx, const double* h_z, int64_t B, int N, const double* h_p,
                                  int n_p, int flags, double* h_cost, uint8_t* h_collide, double* h_g);
int uam_analytic_g_len(const uam_ctx* ctx, int N, int64_t* len);

/* Gradient of Problem.get_cost with respect to every waypoint: d_grad (B, 2(N+2)) float64, same layout as d_z
 * (columns 2 .. 2N+1 are the solver's decision variables z_1..z_N, solver.py:59); d_cost nullable.  The reference gets
 * this derivative from CasADi's algorithmic differentiation inside OpEn (solver.py:82-101); here it is analytic, fp64.
 * Needs UAM_PENALTY_SMOOTH.  The length term follows get_cost (last segment absent, problem.py:39,140-145). */
int uam_grad_paths_analytic(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p, int flags,
                            double* d_cost, double* d_grad, void* stream);

/* Problem.length_of(x, smooth) (problem.py:130-146) for B rows of M points each (x is (B, 2M) float64):
 * y = [map.x_start; x; map.x_goal], out = sum of nrm(y_{k+1} - y_k) over the FIRST N+1 pairs (N <= M).
 * M = N reproduces Solver.solve's call (solver.py:49), M = N+2 the call inside get_cost (problem.py:39).
 * h_ends = [x_start(2), x_goal(2)]. */
int uam_length_of(uam_ctx* ctx, const double* d_x, int64_t B, int M, int N, const double* h_ends, int smooth,
                  double* d_out, void* stream);
int uam_length_of_host(uam_ctx* ctx, const double* h_x, int64_t B, int M, int N, const double* h_ends,
                       int smooth, double* h_out);

/* ---- path scoring: rasters ----------------------------------------------------------------------
 * Same cost functional with P(x) = sum_l w_l * bilinear(layer_l, x) and collision = occupancy of
 * the nearest cell.  samples_per_cell == 0: one sample per waypoint (the reference's sampling,
 * problem.py:42-43).  > 0: every segment is sampled S_k = max(1, ceil(|dz_k|_cells * samples_per_cell))
 * times (left-endpoint rule, mean per segment; S_k capped at 2^20) -- build-defined line-integral
 * extension.  p[7:] are the L layer weights.  World->pixel coordinates, S_k and sample positions are
 * fp64 (same cells and sample counts as the oracle, bit for bit); texel arithmetic and the penalty
 * sum are fp32; the length term is fp64; d_cost is float32.  d_nsamples (nullable, int64 per path)
 * receives the number of raster samples taken for the path. */
int uam_score_paths_raster(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p,
                           int flags, double samples_per_cell, float* d_cost, uint8_t* d_collide,
                           int64_t* d_nsamples, void* stream);
int uam_score_paths_raster_host(uam_ctx* ctx, const double* h_z, int64_t B, int N, const double* h_p,
                                int n_p, int flags, double samples_per_cell, float* h_cost,
                                uint8_t* h_collide);

/* Scoring + best candidate in one call: as uam_score_paths_raster, and *d_key receives the best key (see uam_best) of the
 * batch -- of ALL ranks' batches when a peer group is attached (uam_peer_attach): the last kernel of the step finds the
 * local min, pushes it into every peer's symmetric block over NVLink and waits for theirs (no separate collective).
 * Every rank of the group must make the same sequence of *_best / uam_best_allreduce calls (B == 0 is allowed). */
int uam_score_paths_raster_best(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p,
                                int flags, double samples_per_cell, float* d_cost, uint8_t* d_collide,
                                int64_t global_offset, uint64_t* d_key, void* stream);

/* Asynchronous host-buffer scoring: a ring of 3 slots (stream + staging each).  submit queues upload -> (candidate
 * generation ->) whole-batch scoring -> download on the slot's stream and returns a ticket at once; uam_raster_wait(ticket)
 * returns when the results are in h_cost / h_collide / *h_key (nullable; *h_key = this rank's best key with
 * global_offset + b as index).  With two tickets in flight the upload of batch s+1 overlaps the kernels of batch s.  The
 * host buffers must stay valid (and should be pinned) until the wait; a 4th submission first waits for the oldest ticket.
 *   uam_raster_submit_paths_host       h_z (B, 2(N+2)) float64: the caller's waypoints (1 KiB per path at N = 62)
 *   uam_raster_submit_candidates_host  h_cand (B, 5) float64 {xs, ys, xg, yg, displacement}: the candidates are generated on
 *                                      the device (uam_make_candidates with index0 = global_offset), 40 bytes per path --
 *                                      the reference's own flow (displacement -> create_x_init -> score, main.py:160-171) */
int uam_raster_submit_paths_host(uam_ctx* ctx, const double* h_z, int64_t B, int N, const double* h_p, int n_p,
                                 int flags, double samples_per_cell, float* h_cost, uint8_t* h_collide,
                                 uint64_t* h_key, int64_t global_offset, int* ticket);
int uam_raster_submit_candidates_host(uam_ctx* ctx, const double* h_cand, int64_t B, int N, double jitter_sigma,
                                      uint64_t seed, const double* h_p, int n_p, int flags, double samples_per_cell,
                                      float* h_cost, uint8_t* h_collide, uint64_t* h_key, int64_t global_offset,
                                      int* ticket);
int uam_raster_wait(uam_ctx* ctx, int ticket);

/* ---- point queries --------------------------------------------------------------------------------
 * Replaces Problem.get_penalty_function(region)(x) (problem.py:59-82), get_total_penalty_function
 * (:49-56) and Map.collides(x) (map.py:41-43) for M points x (M,2) float64.  Outputs nullable:
 *   d_region_pen (M, R) float64 weighted per-region penalties; d_obst_pen (M) float64 =
 *   get_penalty_function(None)(x); d_collide (M) uint8.  All fp64 (this is the exact path). */
int uam_eval_points(uam_ctx* ctx, const double* d_x, int64_t M, const double* h_p, int n_p, int flags,
                    double* d_region_pen, double* d_obst_pen, uint8_t* d_collide, void* stream);
int uam_eval_points_host(uam_ctx* ctx, const double* h_x, int64_t M, const double* h_p, int n_p, int flags,
                         double* h_region_pen, double* h_obst_pen, uint8_t* h_collide);

/* Function.__call__ (function.py:119-120): h_i(x_m) of n_rec raw inequality records (8 doubles each, the
 * record kinds above) at M points, h_out[i*M + m]; fp64, bit-exact with the reference's closures
 * (polygon.py:69-71,98; ball.py:33-37; square.py:29-51). */
int uam_eval_inequalities_host(uam_ctx* ctx, const double* h_records, int n_rec, const double* h_x, int64_t M,
                               double* h_out);

/* ---- best candidate --------------------------------------------------------------------------------
 * Replaces the running min of path_generation/main.py:162-180.  *d_key = min over b of
 *     key(cost[b], global_offset + b) = (img(float32 cost) << 31) | index        (63 bits, index < 2^31)
 * img = order-preserving image of the float32 bit pattern (all bits flipped for a negative value, top bit set otherwise;
 * NaN -> 0xffffffff), so the unsigned order of the keys is the float order of the costs for EVERY value (negative costs
 * are reachable through negative layer weights), a NaN never wins, ties resolve to the smaller index (the strict `<` of
 * main.py:175) and the key is a non-negative int64: the caller min-reduces it across ranks with a signed or unsigned min
 * alike (NCCL).  An empty batch leaves 2^63 - 1.  d_key must be pre-set (to 2^63 - 1) by the caller or by passing
 * reset != 0.  d_cost is float32 (raster scorer) or float64 (analytic scorer, rounded to float32 for the key) according
 * to cost_is_f64. */
int uam_best(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
             uint64_t* d_key, int reset, void* stream);

/* The same with the cross-rank exchange built in: *d_key = min over the attached peer group (or this rank alone when no
 * group is attached) of the batch keys; B == 0 takes part with "no candidate".  One kernel: local min + push / wait over
 * NVLink peer memory in its last CTA. */
int uam_best_allreduce(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
                       uint64_t* d_key, void* stream);
/* Peer group of the ranks of one box (one process per GPU).  uam_peer_export writes the 64-byte CUDA IPC handle of this
 * ctx's symmetric block; the caller all-gathers the handles (torch.distributed, any backend) and passes all `world` of them
 * (rank order, 64 bytes each) to uam_peer_attach, which maps the peers' blocks.  uam_peer_status: *timed_out != 0 if a
 * wait for the peers ever gave up (2 s; the key of that step is then a partial min). */
int uam_peer_export(uam_ctx* ctx, void* h_handle64);
int uam_peer_attach(uam_ctx* ctx, int rank, int world, const void* h_handles);
int uam_peer_status(uam_ctx* ctx, int* timed_out);

/* ---- map rebuild (map_generation) --------------------------------------------------------------------
 * uam_dem_mask: mask = image > threshold, or image == -9999 when threshold == -9999
 *               (map_generation/data_manager.py:14-17).
 * uam_rasterize_occupancy: occ[i,j] = Map.collides(cell centre)  (map.py:41-43), fp64, bit-exact.
 * uam_rasterize_layers: layer[l,i,j] = float32(sum_{s in region l} psi_s(xc;e)/psi_s(c_s;e)), unweighted
 *               (problem.py:72-80 without w; quadratic_obstacle.py:33-35).
 * uam_edt: exact squared Euclidean distance (cells) to the nearest occupied cell and clearance =
 *               sqrt(d2)*cell (build-defined extension; no reference counterpart). */
int uam_dem_mask(uam_ctx* ctx, const float* d_image, int64_t n, float threshold, uint8_t* d_mask,
                 void* stream);
int uam_rasterize_occupancy(uam_ctx* ctx, int H, int W, double x0, double dx, double y0, double dy,
                            uint8_t* d_occ, void* stream);
int uam_rasterize_layers(uam_ctx* ctx, int H, int W, double x0, double dx, double y0, double dy,
                         double enlargement, float* d_layers, void* stream);
int uam_edt(uam_ctx* ctx, const uint8_t* d_occ, int H, int W, double cell, int32_t* d_dist2,
            float* d_clearance, void* stream);

/* ---- polygon front-end of map_generation ------------------------------------------------
 * The data-parallel core of DataManager.load_dem_polygons_from_geotiff (map_generation/data_manager.py:11-19: one polygon
 * per connected region of the mask, as rasterio.features.shapes yields them with its default 4-connectivity) and of
 * DataProcessor.process_polygons (map_generation/data_processor.py:16-34,67-71: area filter + cv2.minAreaRect of each
 * polygon's exterior ring).
 * uam_label_components: d_mask (H,W) uint8 -> d_labels (H,W) int32: 0 = background, components numbered 1..n in raster-scan
 *                       order of their first cell (scipy.ndimage.label's numbering); connectivity 4 or 8; *h_n_components
 *                       (host) receives n.  Synchronises `stream`.
 * uam_component_stats:  d_area (n) int64 = cells per component (polygon.area / cell area); d_bbox (n,4) int32 =
 *                       {row min, row max, col min, col max}.  Synchronises `stream`.
 * uam_component_rects:  minimum-area enclosing rectangle of the cell corners of each listed component (d_ids: K distinct
 *                       labels) = of the polygon's exterior ring; d_rect (K,4,2) float64 world coordinates of the corners
 *                       (x0 + col*dx, y0 + row*dy at cell corners), consecutive around the rectangle, the first two on the
 *                       line of the chosen hull edge; d_info (K,2) int32 (nullable) = {hull vertices, chosen edge}.  Exact:
 *                       every hull edge is tried with integer extents and a 128-bit area comparison, ties to the first
 *                       edge.  Needs square cells and H, W <= 32766.  Synchronises `stream`. */
int uam_label_components(uam_ctx* ctx, const uint8_t* d_mask, int H, int W, int connectivity, int32_t* d_labels,
                         int32_t* h_n_components, void* stream);
int uam_component_stats(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int n_components, int64_t* d_area,
                        int32_t* d_bbox, void* stream);
int uam_component_rects(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int n_components, const int32_t* d_bbox,
                        const int32_t* d_ids, int K, double x0, double dx, double y0, double dy, double* d_rect,
                        int32_t* d_info, void* stream);

/* Large-polygon split of DataProcessor._divide_and_approximate_polygon (map_generation/data_processor.py:34-53): mask of
 * box (box_row, box_col) of the divisions x divisions box grid over component `label`'s bounding box h_bbox = {row min, row
 * max, col min, col max} (host), on the grid refined `divisions` times: d_mask (nr, nc) uint8 with nr, nc = the bounding
 * box's extents in cells.  Sub-cell (r, c) has the corners (col min + (box_col nc + c) / divisions, row min + (box_row nr +
 * r) / divisions) .. + 1 / divisions in cell units.  The box edges are sub-cell boundaries, so the 4-connected regions of the
 * mask (uam_label_components) are the pieces of polygon.intersection(box) and uam_component_rects on them, with the affine of
 * the refined grid, gives each piece's minimum-area rectangle. */
int uam_component_submask(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int label, const int32_t* h_bbox,
                          int divisions, int box_row, int box_col, uint8_t* d_mask, void* stream);

/* ---- grid search / cost-to-go (build-defined extension; the reference has none: SURVEY.md section 0) ----------
 * Q independent single-source cost-to-go sweeps on an 8-connected H x W grid, optionally stacked in `bands` altitude
 * bands, with integer edge costs: in-plane step(u,v) * (cost[b,u] + cost[b,v]), step = 2 (axis) / 3 (diagonal); band
 * change at a fixed cell 2 * (cost[b,v] + cost[b+-1,v]); blocked cells (nullable) are impassable.
 * uam_grid_search:       d_cost (H,W) uint16, d_sources (Q,2) int32 (row, col); d_dist (Q,H,W) int64 (2^62 =
 *                        unreachable); d_parent (Q,H,W) int32 (nullable): flat index of the best predecessor, ties to
 *                        the lowest neighbour slot in the order (-1,-1),(-1,0),(-1,1),(0,-1),(0,1),(1,-1),(1,0),(1,1);
 *                        parent[source] = source, unreachable = -1.
 * uam_grid_search_bands: d_cost / d_blocked (bands,H,W), d_sources (Q,3) int32 (band, row, col); d_dist / d_parent
 *                        (Q,bands,H,W); parent = flat index into (bands,H,W); two more predecessor slots after the eight
 *                        in-plane ones: band below (8), band above (9).
 * Frontier-parallel tile relaxation (warp per 32 x 32 tile, Gauss-Seidel row sweeps with (min,+) warp scans); results
 * are bit-identical to Dijkstra.  Synchronises `stream` internally (the number of rounds is data dependent).
 * Precondition for predecessors: every passable cell has cost >= 1 (a weight-0 edge between two cost-0 cells would let
 * them choose each other); with d_parent != NULL a grid with passable cost-0 cells is refused (UAM_ERR_UNSUPPORTED),
 * with d_parent == NULL distances are computed as usual (they stay exact for weights >= 0). */
/* uam_grid_search_goals: start/goal queries.  As uam_grid_search_bands (bands >= 1; d_sources and d_goals are (Q,3) int32
 *                        {band, row, col}), but a query stops as soon as nothing pending can lower its goal's distance:
 *                        dist is exact for the goal and for every node closer to the source than the goal (nodes farther
 *                        away hold upper bounds or 2^62), and the predecessors of those nodes are the ones of the full
 *                        search (edge weights are positive for cost >= 1), so the extracted path is the same.  A goal
 *                        outside the grid leaves the query unbounded.
 * uam_grid_extract_paths: one path per query from d_parent (Q,bands,H,W): d_path (Q,max_len) int32 flat node ids from the
 *                        source to the goal, d_len (Q) = nodes on the path, 0 = goal not reached, -k = the path has
 *                        k > max_len nodes (nothing written). */
int uam_grid_search_goals(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int bands, int H, int W,
                          const int32_t* d_sources, const int32_t* d_goals, int Q, int64_t* d_dist, int32_t* d_parent,
                          void* stream);
int uam_grid_extract_paths(uam_ctx* ctx, const int32_t* d_parent, int bands, int H, int W, const int32_t* d_sources,
                           const int32_t* d_goals, int Q, int max_len, int32_t* d_path, int32_t* d_len, void* stream);
int uam_grid_search(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int H, int W,
                    const int32_t* d_sources, int Q, int64_t* d_dist, int32_t* d_parent, void* stream);
int uam_grid_search_bands(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int bands, int H, int W,
                          const int32_t* d_sources, int Q, int64_t* d_dist, int32_t* d_parent, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UAM_B200_H */
