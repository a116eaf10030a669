// Batched cost-to-go sweeps on an 8-connected grid with optional altitude bands (build-defined extension: the
// reference has no grid search; BASELINE.json config 5).  Q independent single-source problems share every launch.
//
// Graph: node = (band, row, col).  In-plane edge u -> v costs step(u,v) * (cost[b,u] + cost[b,v]) with step = 2
// (axis) / 3 (diagonal); a band change at a fixed cell costs 2 * (cost[b,v] + cost[b',v]), b' = b +- 1.  All integer,
// so shortest distances are unique numbers and a label-correcting relaxation in ANY order converges to the same
// bits as Dijkstra.  Frontier-parallel relaxation, organised by 32 x 32 tiles of one band:
// Every (query, band, tile) triple has a KEY: the smallest distance that arrived at its border since it was last
// relaxed (2^62 = nothing pending).  Tiles are relaxed in rough distance order (delta-stepping at tile granularity):
//   round:  uam_k_grid_minkey   per query, the smallest pending key;
//           uam_k_grid_select   collects the triples whose key is within `delta` of their query's minimum and
//                               clears their keys (the others wait: relaxing them now would be redone when the
//                               shorter fronts arrive -- measured 32 activations per tile without the ordering);
//           uam_k_grid_relax    one WARP per active triple: loads the tile + 1-cell halo of dist (as 32-bit offsets
//                               from the tile's key; cells below the key are frozen) / cost into its 13.6 KB slice
//                               of shared memory, folds in the candidates from the bands above / below (those do
//                               not change during the activation), then alternates a top-down and a bottom-up
//                               Gauss-Seidel sweep until a sweep in each direction moves nothing.  A sweep step handles one row: the three
//                               neighbours in the previous row, then the exact closure along the row in both
//                               directions as two (min,+) warp scans -- with S the prefix sum of the horizontal
//                               edge weights, min_k<=j (d_k + S_j - S_k) = S_j + prefixmin(d - S) and
//                               min_k>=j (d_k + S_k - S_j) = suffixmin(d + S) - S_j.  Information crosses the
//                               whole tile in one sweep; no block barriers.  Cells that dropped are written back
//                               and the smallest dropped value on each side goes into the key of the neighbour
//                               tile behind that side (atomicMin), the smallest overall into the keys of the same
//                               tile in the bands above / below.
//   rounds repeat until the selection finds no tile.  The loop runs on the device: one CUDA graph whose WHILE conditional node has
//   a round as its body (uam_k_grid_loop_cond re-arms it); UAM_OPT_GRID_GRAPH = 0 is the host-driven loop (active count read
//   back every 8 rounds).  Inside a round the warps take list entries through a work counter; an activation builds its per-row
//   tables once (prefix sums of the horizontal edge weights, dead-edge ballots), runs at most UAM_OPT_GRID_HALF_CAP half sweeps
//   (default: two double sweeps) and, if the tile is not at its fixed point by then, leaves it pending for the next round --
//   a round is bulk-synchronous, so it must not wait for its longest activation.
//   uam_k_grid_parent then picks each cell's predecessor: argmin over the neighbours in a fixed slot order with a
//   strict '<' -- deterministic, identical to the oracle's post-pass.
#include <algorithm>

#include "uam_internal.cuh"

namespace {

#define GT 32                          // tile side
#define GH (GT + 2)                    // with halo
#define UAM_GRID_INF (1ll << 62)
#define UAM_GRID_WARPS 8

__device__ __constant__ int c_di[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
__device__ __constant__ int c_dj[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
__device__ __constant__ int c_st[8] = {3, 2, 3, 2, 2, 3, 2, 3};

struct UamGridGeo {
    int H, W, bands, Q;
    int tiles_x, tiles_y;
    int src_stride;          // 2: sources are (row, col) in band 0; 3: (band, row, col)
    int half_cap;            // half sweeps per activation before the tile is handed to the next round (0: to the fixed point)
};

__global__ void __launch_bounds__(256)
uam_k_grid_init(long long* __restrict__ dist, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) dist[t] = UAM_GRID_INF;
}

__device__ __forceinline__ bool uam_grid_source(const int* __restrict__ sources, const UamGridGeo g, int q, int& b, int& i, int& j) {
    if (g.src_stride == 3) { b = sources[3 * q]; i = sources[3 * q + 1]; j = sources[3 * q + 2]; }
    else { b = 0; i = sources[2 * q]; j = sources[2 * q + 1]; }
    return b >= 0 && b < g.bands && i >= 0 && i < g.H && j >= 0 && j < g.W;
}

__global__ void uam_k_grid_seed(long long* __restrict__ dist, const uint8_t* __restrict__ blocked,
                                const int* __restrict__ sources, UamGridGeo g, unsigned long long* __restrict__ keys) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.Q) return;
    int sb, si, sj;
    if (!uam_grid_source(sources, g, q, sb, si, sj)) return;
    const size_t cells = (size_t)g.H * g.W;
    const size_t c = (size_t)sb * cells + (size_t)si * g.W + sj;
    if (blocked && blocked[c]) return;
    dist[(size_t)q * g.bands * cells + c] = 0;
    // the source is an arrival of value 0 for its own tile AND for every tile whose halo (or band stack) holds it: a tile
    // relaxed with a larger key would treat the source cell as frozen
    const size_t tiles = (size_t)g.tiles_x * g.tiles_y;
    const int ti = si / GT, tj = sj / GT;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            const int ni = ti + di, nj = tj + dj;
            if (ni >= 0 && ni < g.tiles_y && nj >= 0 && nj < g.tiles_x)
                keys[((size_t)q * g.bands + sb) * tiles + (size_t)ni * g.tiles_x + nj] = 0ull;
        }
    if (sb > 0) keys[((size_t)q * g.bands + sb - 1) * tiles + (size_t)ti * g.tiles_x + tj] = 0ull;
    if (sb + 1 < g.bands) keys[((size_t)q * g.bands + sb + 1) * tiles + (size_t)ti * g.tiles_x + tj] = 0ull;
}

// per query: smallest pending key (grid = Q x parts CTAs)
__global__ void __launch_bounds__(256)
uam_k_grid_minkey(const unsigned long long* __restrict__ keys, size_t per_q, int parts, unsigned long long* __restrict__ minkey) {
    const int q = blockIdx.x / parts, part = blockIdx.x - q * parts;
    const unsigned long long* kq = keys + (size_t)q * per_q;
    unsigned long long m = (unsigned long long)UAM_GRID_INF;
    for (size_t t = (size_t)part * blockDim.x + threadIdx.x; t < per_q; t += (size_t)parts * blockDim.x) {
        const unsigned long long k = kq[t];
        m = k < m ? k : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t < m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m < (unsigned long long)UAM_GRID_INF) atomicMin(&minkey[q], m);
}

__global__ void __launch_bounds__(256)
uam_k_grid_select(unsigned long long* __restrict__ keys, size_t n, size_t per_q, const unsigned long long* __restrict__ minkey,
                  unsigned long long delta, unsigned* __restrict__ list, unsigned long long* __restrict__ list_key,
                  unsigned* __restrict__ count, const long long* __restrict__ dist, const long long* __restrict__ goal_at) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const unsigned long long k = keys[t];
        // goal-bounded query: an arrival that is not below the goal's current distance cannot lower it (edge weights are
        // >= 0) -- the tile stays pending but is never relaxed; the query is finished when nothing below the bound is left
        if (goal_at) {
            const long long ga = goal_at[t / per_q];
            if (ga >= 0 && k >= (unsigned long long)dist[ga]) continue;
        }
        if (k < (unsigned long long)UAM_GRID_INF && k <= minkey[t / per_q] + delta) {
            keys[t] = (unsigned long long)UAM_GRID_INF;
            const unsigned pos = atomicAdd(count, 1u);
            list[pos] = (unsigned)t;
            list_key[pos] = k;
        }
    }
}

// Distances inside an activation are 32-bit and RELATIVE TO THE TILE'S KEY.  Everything this activation can produce is
// >= key (a relaxation chain starts at a border value that arrived since the last relaxation, and the key is their
// minimum), so a cell whose distance is below the key can neither drop nor pass on anything the tile has not already
// seen: it is frozen and treated like a blocked cell.  Cells at key + UAM_GRID_LIM or more do not fit: they enter as
// "unreached"; if one of them is still that far after the sweeps the tile is relaxed again with the smallest such
// value as its key (a wall inside the tile separating fronts 10^9 apart -- rare).
#define UAM_GRID_INF32 0x3fffffff
#define UAM_GRID_LIM 0x38000000          // INF32 - 2^27: room for any in-tile growth (34 cells x 3 x 131070 < 2^24)

// One sweep step: row li of the tile takes candidates from row ln (li - 1 or li + 1) and closes along the row.
// Lane l owns column l + 1.  Edge e_l joins columns l and l + 1 (l = 0..32); it is dead when either end is blocked.
// S_l = weight of e_0..e_l (prefix sum, dead edges count 0); a candidate from column k to column j is
// d_k + |S_j - S_k| and is valid iff no dead edge lies between them -- checked on the ballot of dead edges, so the two
// min-scans (from the left, from the right) need no sentinel weights.  span_l[s] / span_r[s]: the edges between this
// lane and the lane 2^s to its left / right (all ones when there is no such lane).
// Per activation and tile row (the costs do not change while a tile is relaxed): prefix sums S of the horizontal edge
// weights, the ballot of dead edges and S_33 (-1: the edge to the right halo column is dead).  uam_grid_row_step used to
// rebuild them in every one of the ~160 row steps of an activation -- a five-step shuffle scan on the critical path in front of
// the two closure scans.
__device__ __forceinline__ void uam_grid_row_tables(const int* __restrict__ C, int* __restrict__ Srow, unsigned* __restrict__ deadm,
                                                    int* __restrict__ S33v, int lane) {
    for (int li = 1; li <= GT; ++li) {
        const int cv = C[li * GH + lane + 1];
        const int cl = C[li * GH + lane];                 // column to the left (lane 0: halo column 0)
        const bool dead_l = cv < 0 || cl < 0;             // e_lane
        const unsigned dead = __ballot_sync(0xffffffffu, dead_l);
        int S = dead_l ? 0 : 2 * (cv + cl);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, S, o);
            if (lane >= o) S += t;
        }
        const int c32 = C[li * GH + GT], c33 = C[li * GH + GT + 1];
        const bool dead32 = c32 < 0 || c33 < 0;
        const int S33 = __shfl_sync(0xffffffffu, S, 31) + (dead32 ? 0 : 2 * (c32 + c33));
        Srow[(li - 1) * 32 + lane] = S;
        if (lane == 0) {
            deadm[li - 1] = dead;
            S33v[li - 1] = dead32 ? -1 : S33;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ bool uam_grid_row_step(int* __restrict__ D, const int* __restrict__ C, const int* __restrict__ Srow,
                                                  const unsigned* __restrict__ deadm, const int* __restrict__ S33v, int li, int ln,
                                                  int lane, const unsigned below, const unsigned (&span_l)[5],
                                                  const unsigned (&span_r)[5]) {
    const int lj = lane + 1;
    const int cv = C[li * GH + lj];
    const int d0 = D[li * GH + lj];
    const unsigned dead = deadm[li - 1];
    const int S = Srow[(li - 1) * 32 + lane];
    const int S33 = S33v[li - 1];
    const bool dead32 = S33 < 0;
    // candidates from the previous row
    int d = d0;
    if (cv >= 0) {
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            const int cn = C[ln * GH + lj + dj];
            if (cn >= 0) d = min(d, D[ln * GH + lj + dj] + (dj ? 3 : 2) * (cn + cv));
        }
    }
    // closure along the row: m = min over valid k <= j of (d_k - S_k), gm = min over valid k >= j of (d_k + S_k)
    int m = d - S, gm = d + S;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int tm = __shfl_up_sync(0xffffffffu, m, 1 << s);
        const int tg = __shfl_down_sync(0xffffffffu, gm, 1 << s);
        if (!(dead & span_l[s])) m = min(m, tm);
        if (!(dead & span_r[s])) gm = min(gm, tg);
    }
    // halo columns: column 0 (S = 0) reaches lane j iff e_0..e_j alive; column 33 iff e_{j+1}..e_32 alive
    if (!(dead & below)) m = min(m, D[li * GH]);
    if (!(dead & ~below) && !dead32) gm = min(gm, D[li * GH + GT + 1] + S33);
    const int nd = min(d, min(m + S, gm - S));
    const bool drop = cv >= 0 && nd < d0 && nd < UAM_GRID_INF32;
    if (drop) D[li * GH + lj] = nd;
    return drop;
}

#define UAM_GRID_WARP_SMEM (GH * GH * 4 + GH * GH * 4 + GT * 32 * 4 + GT * 4 + GT * 4)      // D | C | S | dead | S33: 13 600 B

__global__ void __launch_bounds__(UAM_GRID_WARPS * 32, 2)
uam_k_grid_relax(const uint16_t* __restrict__ cost, const uint8_t* __restrict__ blocked, UamGridGeo g,
                 const unsigned* __restrict__ list, const unsigned long long* __restrict__ list_key,
                 unsigned* __restrict__ count, long long* __restrict__ dist, unsigned long long* __restrict__ keys,
                 unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char uam_grid_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* D = reinterpret_cast<int*>(uam_grid_smem + (size_t)warp * UAM_GRID_WARP_SMEM);      // distance - key
    int* C = D + GH * GH;                                                                      // cell cost, -1 = blocked / outside / frozen
    int* Srow = C + GH * GH;                                                                   // per-row tables of the row step
    unsigned* deadm = reinterpret_cast<unsigned*>(Srow + GT * 32);
    int* S33v = reinterpret_cast<int*>(deadm + GT);
    const int H = g.H, W = g.W;
    const size_t cells = (size_t)H * W;
    const unsigned tiles = (unsigned)(g.tiles_x * g.tiles_y);
    const unsigned n = *count;
    // lane-constant masks of the scan steps
    const unsigned below = (2u << lane) - 1u;         // bits 0..lane
    unsigned span_l[5], span_r[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int o = 1 << s;
        // edges between lane - o and lane: e_{lane-o+1}..e_lane; between lane and lane + o: e_{lane+1}..e_{lane+o}
        span_l[s] = lane >= o ? (below & ~((2u << (lane - o)) - 1u)) : 0xffffffffu;
        span_r[s] = lane + o < 32 ? (((2u << (lane + o)) - 1u) & ~below) : 0xffffffffu;
    }
    // the entries are handed out through a counter (count[1], zeroed with the list length): activations differ widely in
    // length (1 .. a dozen double sweeps), a fixed share per warp would leave the round waiting for the unluckiest warp
    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = atomicAdd(count + 1, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n) break;
        const unsigned ent = list[w];
        long long key = (long long)list_key[w];
        const unsigned qb = ent / tiles, tile = ent - qb * tiles;
        const int q = (int)(qb / (unsigned)g.bands), b = (int)(qb - (unsigned)q * g.bands);
        const int ti = (int)(tile / (unsigned)g.tiles_x), tj = (int)(tile - (unsigned)ti * g.tiles_x);
        const int i0 = ti * GT - 1, j0 = tj * GT - 1;
        long long* dq = dist + ((size_t)q * g.bands + b) * cells;
        const uint16_t* cb = cost + (size_t)b * cells;
        const uint8_t* bb = blocked ? blocked + (size_t)b * cells : nullptr;
        long long m_all = UAM_GRID_INF, m_top = UAM_GRID_INF, m_bot = UAM_GRID_INF, m_side = UAM_GRID_INF;   // m_side: this lane's column
        unsigned n_half = 0;                      // half sweeps (32 row steps each)
        long long repost = UAM_GRID_INF;          // the sweeps were cut off at g.half_cap under this key: the tile is not at its fixed point yet
        for (;;) {                               // one trip unless some cell sits 10^9 above the key (see above)
            long long far_min = UAM_GRID_INF;
            __syncwarp();
            // ---- load tile + halo (34 rows x 34 columns: lanes cover columns 0..31, lanes 0..1 also 32..33); the loads of
            //      several rows are issued together (the activation's latency is what bounds a round) ------------------
            // a tile whose halo lies inside the raster (all but the border tiles) needs no bounds tests and 32-bit offsets from
            // the tile's corner; its two right halo columns are read with the lanes down the rows (4 strided loads per lane
            // instead of 34 trips with two active lanes) -- the load was 5.7 k of the ~30 k instructions of an activation
            const bool interior = i0 >= 0 && j0 >= 0 && i0 + GH <= H && j0 + GH <= W;          // (warp-uniform)
            if (interior) {
                const long long* dp = dq + (size_t)i0 * W + j0;
                const uint16_t* cp = cb + (size_t)i0 * W + j0;
                const uint8_t* bp = bb ? bb + (size_t)i0 * W + j0 : nullptr;
                auto put = [&](int r, int c, long long dv, int cv, uint8_t bl) {
                    const long long rel = dv - key;
                    const bool frozen = bl || rel < 0;
                    const bool far = !frozen && rel >= UAM_GRID_LIM;
                    if (far && dv < UAM_GRID_INF) far_min = dv < far_min ? dv : far_min;
                    D[r * GH + c] = (frozen || far) ? UAM_GRID_INF32 : (int)rel;
                    C[r * GH + c] = frozen ? -1 : cv;
                };
#pragma unroll 1
                for (int r0 = 0; r0 < 30; r0 += 6) {          // rows 0 .. 29, six at a time
                    long long dv[6];
                    int cvv[6];
                    uint8_t bl[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const unsigned o = (unsigned)(r0 + k) * (unsigned)W + (unsigned)lane;
                        dv[k] = dp[o];
                        cvv[k] = (int)cp[o];
                        bl[k] = bp ? bp[o] : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) put(r0 + k, lane, dv[k], cvv[k], bl[k]);
                }
                {                                             // rows 30 .. 33
                    long long dv[4];
                    int cvv[4];
                    uint8_t bl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned o = (unsigned)(30 + k) * (unsigned)W + (unsigned)lane;
                        dv[k] = dp[o];
                        cvv[k] = (int)cp[o];
                        bl[k] = bp ? bp[o] : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) put(30 + k, lane, dv[k], cvv[k], bl[k]);
                }
                {   // columns 32 and 33: element (row, column) = (lane + 32 a, 32 + c)
                    long long dv[4];
                    int cvv[4];
                    uint8_t bl[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int r = lane + 32 * (t >> 1), c = 32 + (t & 1);
                        const bool ok = r < GH;
                        const unsigned o = ok ? (unsigned)r * (unsigned)W + (unsigned)c : 0u;
                        dv[t] = ok ? dp[o] : UAM_GRID_INF;
                        cvv[t] = ok ? (int)cp[o] : -1;
                        bl[t] = (ok && bp) ? bp[o] : 0;
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int r = lane + 32 * (t >> 1), c = 32 + (t & 1);
                        if (r < GH) put(r, c, dv[t], cvv[t], bl[t]);
                    }
                }
            } else {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int lj = lane + 32 * pass;
                const int j = j0 + lj;
                const bool col_ok = lj < GH && j >= 0 && j < W;
                for (int r0 = 0; r0 < GH; r0 += 6) {
                    long long dv[6];
                    int cvv[6];
                    uint8_t bl[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int i = i0 + r0 + k;
                        const bool ok = col_ok && r0 + k < GH && i >= 0 && i < H;
                        const size_t gi = ok ? (size_t)i * W + j : 0;
                        dv[k] = ok ? dq[gi] : UAM_GRID_INF;
                        cvv[k] = ok ? (int)cb[gi] : -1;
                        bl[k] = (ok && bb) ? bb[gi] : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (lj < GH && r0 + k < GH) {
                            const long long rel = dv[k] - key;
                            const bool frozen = bl[k] || rel < 0;
                            const bool far = !frozen && rel >= UAM_GRID_LIM;
                            if (far && dv[k] < UAM_GRID_INF) far_min = dv[k] < far_min ? dv[k] : far_min;
                            D[(r0 + k) * GH + lj] = (frozen || far) ? UAM_GRID_INF32 : (int)rel;
                            C[(r0 + k) * GH + lj] = frozen ? -1 : cvv[k];
                        }
                    }
                }
            }
            }
            __syncwarp();
            // ---- candidates from the bands below / above (fixed during this activation) ------------------------------
            for (int db = -1; db <= 1; db += 2) {
                const int b2 = b + db;
                if (b2 < 0 || b2 >= g.bands) continue;
                const long long* d2 = dist + ((size_t)q * g.bands + b2) * cells;
                const uint16_t* c2 = cost + (size_t)b2 * cells;
                const uint8_t* bl2 = blocked ? blocked + (size_t)b2 * cells : nullptr;
                const int j = j0 + lane + 1;
                const size_t corner = interior ? (size_t)i0 * W + j : 0;
                for (int r0 = 1; r0 <= GT; r0 += 8) {
                    long long dv[8];
                    int cvv[8];
                    uint8_t bl[8];
                    if (interior) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const unsigned o = (unsigned)(r0 + k) * (unsigned)W;
                            dv[k] = d2[corner + o];
                            cvv[k] = (int)c2[corner + o];
                            bl[k] = bl2 ? bl2[corner + o] : 0;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int i = i0 + r0 + k;
                            const bool ok = i < H && j < W;
                            const size_t gi = ok ? (size_t)i * W + j : 0;
                            dv[k] = ok ? d2[gi] : UAM_GRID_INF;
                            cvv[k] = ok ? (int)c2[gi] : 0;
                            bl[k] = (ok && bl2) ? bl2[gi] : 0;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int li = r0 + k;
                        const int cv = C[li * GH + lane + 1];
                        if (cv >= 0 && !bl[k] && dv[k] < UAM_GRID_INF) {
                            const long long rel = dv[k] + (long long)(2 * (cvv[k] + cv)) - key;
                            if (rel >= UAM_GRID_LIM) far_min = dv[k] < far_min ? dv[k] : far_min;       // source too far for this key
                            else if (rel >= 0 && (int)rel < D[li * GH + lane + 1]) D[li * GH + lane + 1] = (int)rel;
                        }
                    }
                }
            }
            __syncwarp();
            // ---- Gauss-Seidel sweeps to the local fixed point -----------------------------------------------------------
            // The tile is at its fixed point when a top-down and a bottom-up sweep IN A ROW move nothing (the first leaves the
            // state as it was, so the second has checked the same state): the loop ends after two quiet half sweeps, whichever
            // direction comes last -- not only after a quiet top-down + bottom-up pair, which cost half a double sweep more per
            // activation on average.
            {
                uam_grid_row_tables(C, Srow, deadm, S33v, lane);
                int quiet = 0;
                const unsigned cap = g.half_cap > 0 ? n_half + (unsigned)g.half_cap : 0xffffffffu;
                for (;;) {
                    bool ch = false;
                    ++n_half;
                    for (int li = 1; li <= GT; ++li) {
                        ch |= uam_grid_row_step(D, C, Srow, deadm, S33v, li, li - 1, lane, below, span_l, span_r);
                        __syncwarp();
                    }
                    quiet = __any_sync(0xffffffffu, ch) ? 0 : quiet + 1;
                    if (quiet >= 2) break;
                    ch = false;
                    ++n_half;
                    for (int li = GT; li >= 1; --li) {
                        ch |= uam_grid_row_step(D, C, Srow, deadm, S33v, li, li + 1, lane, below, span_l, span_r);
                        __syncwarp();
                    }
                    quiet = __any_sync(0xffffffffu, ch) ? 0 : quiet + 1;
                    if (quiet >= 2) break;
                    if (n_half >= cap) { repost = key < repost ? key : repost; break; }      // (warp-uniform) continue in the next round
                }
            }
            // ---- write back the cells that dropped; smallest dropped value per side ------------------------------------
            if (interior) {
                long long* dp = dq + (size_t)i0 * W + j0 + lane + 1;
                for (int r0 = 1; r0 <= GT; r0 += 8) {
                    long long old[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) old[k] = dp[(unsigned)(r0 + k) * (unsigned)W];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int li = r0 + k;
                        const int rel = D[li * GH + lane + 1];
                        const long long now = key + (long long)rel;
                        if (rel < UAM_GRID_INF32 && now < old[k]) {
                            dp[(unsigned)li * (unsigned)W] = now;
                            m_all = now < m_all ? now : m_all;
                            m_side = now < m_side ? now : m_side;
                            if (li == 1) m_top = now < m_top ? now : m_top;
                            if (li == GT) m_bot = now < m_bot ? now : m_bot;
                        }
                    }
                }
            } else {
                const int lj = lane + 1;
                const int j = j0 + lj;
                for (int r0 = 1; r0 <= GT; r0 += 8) {
                    long long old[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int i = i0 + r0 + k;
                        old[k] = (i < H && j < W) ? dq[(size_t)i * W + j] : 0;       // 0: nothing is below it
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int li = r0 + k;
                        const int rel = D[li * GH + lj];
                        const long long now = key + (long long)rel;
                        if (rel < UAM_GRID_INF32 && now < old[k]) {
                            dq[(size_t)(i0 + li) * W + j] = now;
                            m_all = now < m_all ? now : m_all;
                            m_side = now < m_side ? now : m_side;
                            if (li == 1) m_top = now < m_top ? now : m_top;
                            if (li == GT) m_bot = now < m_bot ? now : m_bot;
                        }
                    }
                }
            }
            // another trip?  only when something finite did not fit under this key
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const long long t = __shfl_xor_sync(0xffffffffu, far_min, o);
                far_min = t < far_min ? t : far_min;
            }
            if (far_min >= UAM_GRID_INF) break;
            key = far_min;
        }
        // corners (single cells), left / right columns (lanes 0 / 31), top / bottom rows and overall (warp minima)
        const long long c_tl = __shfl_sync(0xffffffffu, m_top, 0), c_tr = __shfl_sync(0xffffffffu, m_top, 31);
        const long long c_bl = __shfl_sync(0xffffffffu, m_bot, 0), c_br = __shfl_sync(0xffffffffu, m_bot, 31);
        const long long m_left = __shfl_sync(0xffffffffu, m_side, 0), m_right = __shfl_sync(0xffffffffu, m_side, 31);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long t = __shfl_xor_sync(0xffffffffu, m_all, o);
            m_all = t < m_all ? t : m_all;
            t = __shfl_xor_sync(0xffffffffu, m_top, o);
            m_top = t < m_top ? t : m_top;
            t = __shfl_xor_sync(0xffffffffu, m_bot, o);
            m_bot = t < m_bot ? t : m_bot;
        }
        if (lane == 0) {                      // counted work: tile activations, double sweeps
            atomicAdd(&stats[0], 1ull);
            atomicAdd(&stats[1], (unsigned long long)n_half);
        }
        if (m_all < UAM_GRID_INF) {
            unsigned long long* kq = keys + (size_t)q * g.bands * tiles;
            if (lane < 8) {
                // slot order of c_di / c_dj: TL, T, TR, L, R, BL, B, BR
                const long long v = lane == 0 ? c_tl : lane == 1 ? m_top : lane == 2 ? c_tr : lane == 3 ? m_left : lane == 4 ? m_right
                                  : lane == 5 ? c_bl : lane == 6 ? m_bot : c_br;
                const int ni = ti + c_di[lane], nj = tj + c_dj[lane];
                if (v < UAM_GRID_INF && ni >= 0 && ni < g.tiles_y && nj >= 0 && nj < g.tiles_x)
                    atomicMin(&kq[(size_t)b * tiles + (size_t)ni * g.tiles_x + nj], (unsigned long long)v);
            } else if (lane == 8 && b > 0) {
                atomicMin(&kq[(size_t)(b - 1) * tiles + tile], (unsigned long long)m_all);
            } else if (lane == 9 && b + 1 < g.bands) {
                atomicMin(&kq[(size_t)(b + 1) * tiles + tile], (unsigned long long)m_all);
            }
        }
        // cut off before the fixed point: the tile stays pending with the key it was relaxed with (what it wrote back are valid
        // upper bounds; the next round goes on from them)
        if (repost < UAM_GRID_INF && lane == 10)
            atomicMin(&keys[((size_t)q * g.bands + b) * tiles + tile], (unsigned long long)repost);
    }
}

// predecessor slots: 0..7 in-plane (c_di / c_dj), 8 = band below, 9 = band above
__global__ void __launch_bounds__(256)
uam_k_grid_parent(const uint16_t* __restrict__ cost, UamGridGeo g, const long long* __restrict__ dist,
                  const int* __restrict__ sources, int* __restrict__ parent) {
    const int H = g.H, W = g.W;
    const size_t cells = (size_t)H * W;
    const size_t nodes = cells * g.bands;
    const size_t total = nodes * g.Q;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int q = (int)(t / nodes);
        const size_t v = t - (size_t)q * nodes;
        const long long* dq = dist + (size_t)q * nodes;
        int p = -1;
        if (dq[v] < UAM_GRID_INF) {
            const int vb = (int)(v / cells);
            const size_t vc = v - (size_t)vb * cells;
            const int vi = (int)(vc / W), vj = (int)(vc - (size_t)vi * W);
            int sb, si, sj;
            uam_grid_source(sources, g, q, sb, si, sj);
            if (vb == sb && vi == si && vj == sj) {
                p = (int)v;
            } else {
                const int cv = cost[v];
                long long best = UAM_GRID_INF;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int ui = vi + c_di[s], uj = vj + c_dj[s];
                    if (ui < 0 || ui >= H || uj < 0 || uj >= W) continue;
                    const size_t u = (size_t)vb * cells + (size_t)ui * W + uj;
                    const long long du = dq[u];
                    if (du >= UAM_GRID_INF) continue;
                    const long long nd = du + (long long)(c_st[s] * ((int)cost[u] + cv));
                    if (nd < best) { best = nd; p = (int)u; }
                }
#pragma unroll
                for (int db = -1; db <= 1; db += 2) {
                    const int ub = vb + db;
                    if (ub < 0 || ub >= g.bands) continue;
                    const size_t u = (size_t)ub * cells + vc;
                    const long long du = dq[u];
                    if (du >= UAM_GRID_INF) continue;
                    const long long nd = du + (long long)(2 * ((int)cost[u] + cv));
                    if (nd < best) { best = nd; p = (int)u; }
                }
            }
        }
        parent[t] = p;
    }
}

// flat index of each query's goal inside dist (q * nodes + node), -1 = no bound for this query (goal outside the grid)
__global__ void uam_k_grid_goal_index(const int* __restrict__ goals, UamGridGeo g, long long* __restrict__ goal_at) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.Q) return;
    int b, i, j;
    const size_t cells = (size_t)g.H * g.W;
    goal_at[q] = uam_grid_source(goals, g, q, b, i, j) ? (long long)((size_t)q * g.bands * cells + (size_t)b * cells + (size_t)i * g.W + j) : -1ll;
}

// One thread per query: follow the predecessors from the goal back to the source, then write the nodes source -> goal.
// len[q] = nodes on the path; 0 = goal not reached (or source / goal outside the grid); -k = the path has k > max_len nodes.
__global__ void uam_k_grid_extract(const int* __restrict__ parent, UamGridGeo g, const int* __restrict__ sources,
                                   const int* __restrict__ goals, int max_len, int* __restrict__ path, int* __restrict__ len) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.Q) return;
    const size_t cells = (size_t)g.H * g.W, nodes = cells * g.bands;
    int sb, si, sj, gb, gi, gj;
    if (!uam_grid_source(sources, g, q, sb, si, sj) || !uam_grid_source(goals, g, q, gb, gi, gj)) { len[q] = 0; return; }
    const int s = (int)((size_t)sb * cells + (size_t)si * g.W + sj), t = (int)((size_t)gb * cells + (size_t)gi * g.W + gj);
    const int* pq = parent + (size_t)q * nodes;
    long long n = 1;
    int v = t;
    while (v != s) {
        v = pq[v];
        if (v < 0 || n > (long long)nodes) { len[q] = 0; return; }        // unreachable (or not a tree: cannot happen)
        ++n;
    }
    if (n > max_len) { len[q] = (int)-n; return; }
    len[q] = (int)n;
    int* out = path + (size_t)q * max_len;
    v = t;
    for (long long k = n - 1; k >= 0; --k) {
        out[k] = v;
        if (k) v = pq[v];
    }
}

// sum of the cell costs (-> mean, for the automatic delta) and number of passable cells of cost 0
__global__ void __launch_bounds__(256)
uam_k_grid_cost_sum(const uint16_t* __restrict__ cost, const uint8_t* __restrict__ blocked, size_t n,
                    unsigned long long* __restrict__ sum) {
    unsigned long long acc = 0, zeros = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const unsigned c = cost[t];
        acc += c;
        zeros += (c == 0 && !(blocked && blocked[t])) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum, acc);
        if (zeros) atomicAdd(sum + 1, zeros);
    }
}

int uam_grid_mean_cost(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, size_t n, cudaStream_t st,
                       unsigned long long* mean, unsigned long long* zero_cells) {
    unsigned long long* d_sum = (unsigned long long*)ctx->d_scratch;        // scratch: the keys are initialised after this
    UAM_CUDA(ctx, cudaMemsetAsync(d_sum, 0, 16, st));
    uam_k_grid_cost_sum<<<ctx->sm_count * 8, 256, 0, st>>>(d_cost, d_blocked, n, d_sum);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_cost_sum");
    unsigned long long h[2] = {0, 0};
    UAM_CUDA(ctx, cudaMemcpyAsync(h, d_sum, 16, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    *mean = std::max<unsigned long long>(1ull, h[0] / (unsigned long long)n);
    *zero_cells = h[1];
    return UAM_OK;
}

// Tail of one relaxation round inside the CUDA-graph WHILE loop: another round unless the selection found no tile (or the
// safety cap is hit).  stats[2] counts the rounds, stats[3] keeps the last round's active count.  The kernel also resets what the
// next round starts from (the list length / work counter and the per-query minima), so a round is three kernels and this one --
// no memset nodes.
__global__ void uam_k_grid_loop_cond(cudaGraphConditionalHandle handle, unsigned* __restrict__ count,
                                     unsigned long long* __restrict__ minkey, int Q, unsigned long long* __restrict__ stats,
                                     unsigned long long max_rounds) {
    for (int q = threadIdx.x; q < Q; q += blockDim.x) minkey[q] = ~0ull;
    if (threadIdx.x == 0) {
        const unsigned long long r = stats[2] + 1ull;
        const unsigned n = count[0];
        stats[2] = r;
        stats[3] = n;
        count[0] = 0u;
        count[1] = 0u;
        cudaGraphSetConditional(handle, (n != 0u && r < max_rounds) ? 1u : 0u);
    }
}

int uam_grid_search_impl(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int bands, int H, int W,
                         const int32_t* d_sources, int src_stride, int Q, int64_t* d_dist, int32_t* d_parent, void* stream,
                         const int32_t* d_goals = nullptr) {
    if (!ctx) return UAM_ERR_INVALID;
    if (H < 1 || W < 1 || bands < 1 || Q < 0 || !d_cost || !d_dist || (Q > 0 && !d_sources))
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_grid_search");
    if ((size_t)H * W * bands >= 0x7fffffffull) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "grid has 2^31 or more nodes");
    if (Q == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_NVTX("uam.grid_search");
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UamGridGeo g;
    g.H = H; g.W = W; g.bands = bands; g.Q = Q; g.src_stride = src_stride;
    g.half_cap = ctx->grid_half_cap;
    g.tiles_x = (W + GT - 1) / GT;
    g.tiles_y = (H + GT - 1) / GT;
    const size_t tiles = (size_t)g.tiles_x * g.tiles_y;
    const size_t n_flags = tiles * bands * Q;
    if (n_flags >= 0xffffffffull) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "too many (query, band, tile) triples");
    // scratch: keys (u64) | list_key (u64) | minkey (u64 x Q) | stats (u64 x 4) | goal_at (i64 x Q) | list (u32) | count (u32 x 2)
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, n_flags * 20 + (size_t)Q * 16 + 256));
    unsigned long long* keys = (unsigned long long*)ctx->d_scratch;
    // delta = cost of crossing about two tiles at the grid's mean cell cost (ordering only: any value gives the same result)
    unsigned long long delta = ctx->grid_delta > 0 ? (unsigned long long)ctx->grid_delta : 0ull;
    {
        unsigned long long mean = 1, zero_cells = 0;
        UAM_TRY(uam_grid_mean_cost(ctx, d_cost, d_blocked, (size_t)H * W * bands, st, &mean, &zero_cells));
        // Edge weights are positive only for cost >= 1.  Two adjacent passable cells of cost 0 are joined by a weight-0 edge:
        // the distances stay exact, but the predecessor rule (argmin of du + w, strict '<' in slot order) could make the
        // two cells each other's parent.  Predecessors are therefore refused on such grids (block the cells or use cost 1).
        if (zero_cells && d_parent)
            return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "%llu passable cells have cost 0: predecessors need cost >= 1 on every passable cell "
                                                      "(pass d_parent = NULL for distances only)", zero_cells);
        // measured on C5 (profiles/r02_sweep_grid_delta.json): one band, 64 queries: 4x / 8x / 16x mean x GT -> 183 / 193 / 180
        // queries/s; eight bands (the fronts also climb between bands: a narrower window re-relaxes fewer tiles), 16 queries:
        // 1x / 2x / 4x / 8x -> 13.8 / 14.3 / 13.6 / 11.3 queries/s
        if (!delta) delta = mean * (bands > 1 ? 2ull : 8ull) * GT;
    }
    unsigned long long* list_key = keys + n_flags;
    unsigned long long* minkey = list_key + n_flags;
    unsigned long long* stats = minkey + Q;
    long long* goal_at = (long long*)(stats + 4);
    unsigned* list = (unsigned*)(goal_at + Q);
    unsigned* count = list + n_flags;
    UAM_CUDA(ctx, cudaMemsetAsync(stats, 0, 32, st));
    const int grid_fill = ctx->sm_count * 16;
    uam_k_grid_init<<<grid_fill, 256, 0, st>>>((long long*)d_dist, (size_t)H * W * bands * Q);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_init");
    uam_k_grid_init<<<grid_fill, 256, 0, st>>>((long long*)keys, n_flags);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_init");
    uam_k_grid_seed<<<(Q + 127) / 128, 128, 0, st>>>((long long*)d_dist, d_blocked, d_sources, g, keys);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_seed");
    if (d_goals) {
        uam_k_grid_goal_index<<<(Q + 127) / 128, 128, 0, st>>>(d_goals, g, goal_at);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_goal_index");
    }
    const size_t per_q = tiles * bands;
    const int parts = (int)std::max<size_t>(1, std::min<size_t>(64, per_q / 2048));
    const size_t smem = (size_t)UAM_GRID_WARP_SMEM * UAM_GRID_WARPS;
    UAM_CUDA(ctx, cudaFuncSetAttribute(uam_k_grid_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid_relax = ctx->sm_count * 2;      // 2 CTAs of 8 warps per SM (13.6 KB of shared memory per warp)
    const long long max_rounds = 1ll << 40;      // the loop ends when no key is pending
    unsigned h_count = 1;
    long long rounds_done = 0;
    // one relaxation round: every argument lives in device memory and none changes from round to round
    auto enqueue_round = [&](cudaStream_t s, bool resets) -> int {
        if (resets) {                    // (the graph loop's tail kernel does them for the next round)
            UAM_CUDA(ctx, cudaMemsetAsync(count, 0, 8, s));
            UAM_CUDA(ctx, cudaMemsetAsync(minkey, 0xff, (size_t)Q * 8, s));
        }
        uam_k_grid_minkey<<<Q * parts, 256, 0, s>>>(keys, per_q, parts, minkey);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_minkey");
        uam_k_grid_select<<<ctx->sm_count * 4, 256, 0, s>>>(keys, n_flags, per_q, minkey, delta, list, list_key, count,
                                                            (const long long*)d_dist, d_goals ? goal_at : nullptr);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_select");
        uam_k_grid_relax<<<grid_relax, UAM_GRID_WARPS * 32, smem, s>>>(d_cost, d_blocked, g, list, list_key, count, (long long*)d_dist, keys, stats);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_relax");
        return UAM_OK;
    };
    // The round loop runs ON THE DEVICE: a CUDA graph whose only top-level node is a WHILE conditional node; its body is one
    // round (captured from the launches above) + a one-thread kernel that keeps the loop going while the selection still finds
    // tiles.  One graph launch per call instead of ~5 launches per round and a host read-back of the active count every 8 rounds
    // (round 1: 5 496 launches and 472 host-synchronised rounds for one 64-query call).  Falls back to the host loop if the
    // graph cannot be built (UAM_OPT_GRID_GRAPH = 0 forces that).
    bool looped = false;
    if (ctx->grid_graph) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaStream_t cap = nullptr;
        bool ok = cudaGraphCreate(&graph, 0) == cudaSuccess;
        cudaGraphConditionalHandle handle = 0;
        ok = ok && cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
        cudaGraph_t body = nullptr;
        if (ok) {
            cudaGraphNodeParams np = {};
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = handle;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            ok = cudaGraphAddNode(&node, graph, nullptr, 0, &np) == cudaSuccess;
            if (ok) body = np.conditional.phGraph_out[0];
        }
        ok = ok && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) == cudaSuccess;
        if (ok) {
            ok = cudaStreamBeginCaptureToGraph(cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int rc = enqueue_round(cap, false);
                uam_k_grid_loop_cond<<<1, 256, 0, cap>>>(handle, count, minkey, Q, stats, 1ull << 24);
                cudaGraph_t captured = nullptr;
                ok = cudaStreamEndCapture(cap, &captured) == cudaSuccess && rc == UAM_OK;
            }
        }
        ok = ok && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (ok) {
            ok = cudaMemsetAsync(count, 0, 8, st) == cudaSuccess && cudaMemsetAsync(minkey, 0xff, (size_t)Q * 8, st) == cudaSuccess &&
                 cudaGraphLaunch(exec, st) == cudaSuccess;
            if (ok) {
                unsigned long long h_tail[2] = {0, 0};      // rounds, active count of the last round
                ok = cudaMemcpyAsync(h_tail, stats + 2, 16, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                     cudaStreamSynchronize(st) == cudaSuccess;
                const unsigned long long h_rounds = h_tail[0];
                h_count = (unsigned)h_tail[1];
                rounds_done = (long long)h_rounds;
                if (h_rounds) ctx->launches += 4ull * h_rounds - 3ull;      // kernels the loop ran (3 were counted during the capture)
                if (!ok) {                      // the loop itself failed: nothing to fall back to
                    if (exec) cudaGraphExecDestroy(exec);
                    if (graph) cudaGraphDestroy(graph);
                    if (cap) cudaStreamDestroy(cap);
                    return uam_fail(ctx, UAM_ERR_CUDA, "grid search graph loop failed: %s", cudaGetErrorString(cudaGetLastError()));
                }
                looped = true;
            }
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (cap) cudaStreamDestroy(cap);
        if (!looped) (void)cudaGetLastError();      // graph not available: clear the error, run the host loop
    }
    for (long long round = 0; !looped && round < max_rounds && h_count; ++round) {
        rounds_done = round + 1;
        UAM_TRY(enqueue_round(st, true));
        if ((round & 7) == 7) {     // termination check every 8 rounds: rounds with an empty list are no-ops
            UAM_CUDA(ctx, cudaMemcpyAsync(&h_count, count, 4, cudaMemcpyDeviceToHost, st));
            UAM_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    if (h_count) return uam_fail(ctx, UAM_ERR_STATE, "grid search did not converge");
    {
        unsigned long long h_stats[2] = {0, 0};
        UAM_CUDA(ctx, cudaMemcpyAsync(h_stats, stats, 16, cudaMemcpyDeviceToHost, st));
        UAM_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->grid_activations = (double)h_stats[0];
        ctx->grid_sweeps = 0.5 * (double)h_stats[1];          // counted in half sweeps, reported in double sweeps
        ctx->grid_rounds = (double)rounds_done;
        // set-up (memset, 2 x init, seed [, goal index]) + the rounds (2 memsets + one graph launch, or 2 memsets + 3 kernels per round) [+ parent]
        ctx->grid_host_submissions = 4.0 + (d_goals ? 1.0 : 0.0) + (looped ? 3.0 : 5.0 * (double)rounds_done) + (d_parent ? 1.0 : 0.0);
    }
    if (d_parent) {
        uam_k_grid_parent<<<grid_fill, 256, 0, st>>>(d_cost, g, (const long long*)d_dist, d_sources, d_parent);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_parent");
    }
    return UAM_OK;
}

}  // namespace

extern "C" int uam_grid_search(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int H, int W,
                               const int32_t* d_sources, int Q, int64_t* d_dist, int32_t* d_parent, void* stream) {
    return uam_grid_search_impl(ctx, d_cost, d_blocked, 1, H, W, d_sources, 2, Q, d_dist, d_parent, stream);
}

extern "C" int uam_grid_search_goals(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int bands, int H, int W,
                                     const int32_t* d_sources, const int32_t* d_goals, int Q, int64_t* d_dist, int32_t* d_parent,
                                     void* stream) {
    if (ctx && !d_goals && Q > 0) return uam_fail(ctx, UAM_ERR_INVALID, "goals pointer is NULL");
    return uam_grid_search_impl(ctx, d_cost, d_blocked, bands, H, W, d_sources, 3, Q, d_dist, d_parent, stream, d_goals);
}

extern "C" int uam_grid_extract_paths(uam_ctx* ctx, const int32_t* d_parent, int bands, int H, int W, const int32_t* d_sources,
                                      const int32_t* d_goals, int Q, int max_len, int32_t* d_path, int32_t* d_len, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (H < 1 || W < 1 || bands < 1 || Q < 0 || max_len < 0) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_grid_extract_paths");
    if (Q == 0) return UAM_OK;
    if (!d_parent || !d_sources || !d_goals || !d_len || (max_len > 0 && !d_path)) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UamGridGeo g;
    g.H = H; g.W = W; g.bands = bands; g.Q = Q; g.src_stride = 3;
    g.tiles_x = (W + GT - 1) / GT;
    g.tiles_y = (H + GT - 1) / GT;
    uam_k_grid_extract<<<(Q + 63) / 64, 64, 0, uam_pick_stream(ctx, stream)>>>(d_parent, g, d_sources, d_goals, max_len, d_path, d_len);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_extract");
    return UAM_OK;
}

extern "C" int uam_grid_search_bands(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int bands, int H, int W,
                                     const int32_t* d_sources, int Q, int64_t* d_dist, int32_t* d_parent, void* stream) {
    return uam_grid_search_impl(ctx, d_cost, d_blocked, bands, H, W, d_sources, 3, Q, d_dist, d_parent, stream);
}
