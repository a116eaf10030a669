// Polygon front-end of map_generation on the GPU: the data-parallel core of
//   DataManager.load_dem_polygons_from_geotiff  (data_manager.py:11-19: mask -> rasterio.features.shapes -> one polygon
//                                                per connected region of the mask, 4-connectivity)
//   DataProcessor.process_polygons              (data_processor.py:16-34,67-71: area filter, cv2.minAreaRect of each
//                                                polygon's exterior ring)
// as three steps on device arrays:
//   uam_label_components   connected-component labelling, labels 1..n in raster-scan order of each component's first cell
//                          (= scipy.ndimage.label's numbering); union-find with atomicMin on flat indices, so a component's
//                          root is its smallest flat index and the result does not depend on the schedule
//   uam_component_stats    cells per component (polygon.area / cell area: holes are not counted, like the polygon's
//                          interior rings) and bounding box, warp-aggregated atomics
//   uam_component_rects    minimum-area enclosing rectangle of each chosen component's cell corners = of the exterior ring
//                          cv2.minAreaRect gets.  Per component: column extremes per grid line (atomics from run ends
//                          only), convex hull by two monotone stacks, then every hull edge is tried as the rectangle's
//                          direction (a minimum-area rectangle has a side on a hull edge) with exact integer extents and a
//                          128-bit cross-multiplied area comparison, ties to the first edge: deterministic, no float
//                          comparisons decide which rectangle wins.
#include <algorithm>
#include <climits>

#include "uam_internal.cuh"

namespace {

// ---- union-find on flat indices -------------------------------------------------------------------------------------
__device__ __forceinline__ int uam_uf_find(const int* __restrict__ L, int i) {
    int p = L[i];
    while (p != i) {
        i = p;
        p = L[i];
    }
    return i;
}

__device__ __forceinline__ void uam_uf_union(int* L, int a, int b) {
    bool done = false;
    while (!done) {
        a = uam_uf_find(L, a);
        b = uam_uf_find(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = old == b;
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = old == a;
            a = old;
        } else {
            done = true;
        }
    }
}

// Row pass: every foreground cell starts as the first cell of its horizontal run (a run = consecutive foreground cells of
// one row), found with a ballot per 32 cells + the carry across 32-cell words done by a union in the merge pass.
__global__ void __launch_bounds__(256)
uam_k_ccl_init(const uint8_t* __restrict__ mask, int H, int W, int* __restrict__ L) {
    const long long n = (long long)H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = i < n;
    const bool fg = in && mask[i] != 0;
    const int col = in ? (int)(i % W) : 0;
    // cells of this warp that continue a run inside the warp: foreground, not at column 0, left neighbour (lane - 1) foreground
    const unsigned fgm = __ballot_sync(0xffffffffu, fg);
    if (!in) return;
    if (!fg) { L[i] = -1; return; }
    // start of the run inside this warp's 32 cells: walk left over set bits while the row does not wrap
    int back = 0;
    if (lane > 0) {
        const unsigned below = fgm << (32 - lane);        // bit 31 = lane - 1, bit 30 = lane - 2, ...
        back = min(__clz(~below), min(lane, col));        // consecutive foreground lanes to the left, inside the row
    }
    L[i] = (int)(i - back);
}

// Merge pass: unite a cell with its upper neighbour(s), and a run that starts at a warp boundary with the cell to its left.
__global__ void __launch_bounds__(256)
uam_k_ccl_merge(const uint8_t* __restrict__ mask, int H, int W, int conn8, int* __restrict__ L) {
    const long long n = (long long)H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || mask[i] == 0) return;
    const int row = (int)(i / W), col = (int)(i - (long long)row * W);
    // a run that continues across a 32-cell word boundary: the init pass linked cells inside a word only
    if (col > 0 && (i & 31) == 0 && mask[i - 1] != 0) uam_uf_union(L, (int)i, (int)(i - 1));
    if (row > 0) {
        const uint8_t up = mask[i - W];
        // one union per run of the upper row that touches this cell: skip when the left cell is foreground and shares `up`
        if (up != 0 && !(col > 0 && mask[i - 1] != 0 && mask[i - W - 1] != 0)) uam_uf_union(L, (int)i, (int)(i - W));
        if (conn8) {
            if (col > 0 && mask[i - W - 1] != 0 && up == 0 && !(mask[i - 1] != 0)) uam_uf_union(L, (int)i, (int)(i - W - 1));
            if (col + 1 < W && mask[i - W + 1] != 0 && up == 0) uam_uf_union(L, (int)i, (int)(i - W + 1));
        }
    }
}

// ---- block-local labelling (round 2) --------------------------------------------------------------------------------
// The two kernels above chase pointers through HBM for every cell (ncu r01: long-scoreboard stalls of 28-31 per issue).
// Here a CTA first labels a 32 x 32-cell tile entirely in shared memory -- runs per tile row by ballot, unions with the
// row above by shared-memory atomicMin, local flatten -- and writes for every cell the GLOBAL index of its tile-local
// root (the smallest index of its local component, hence <= the cell's own index: the union-find invariant holds).
// Only the pairs that cross a tile border are then united in HBM (1/16 of the cells), and the global flatten finds a
// root in one or two hops.  The pairing rules are the ones of uam_k_ccl_merge (a union is idempotent, so a rule applied by
// both kernels costs nothing but time); the result is the same partition, hence the same roots and the same labels.
#define UAM_CCL_T 32
__device__ __forceinline__ int uam_sm_find(const int* lab, int i) {
    int p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}
__device__ __forceinline__ void uam_sm_union(int* lab, int a, int b) {
    bool done = false;
    while (!done) {
        a = uam_sm_find(lab, a);
        b = uam_sm_find(lab, b);
        if (a < b) { const int old = atomicMin(&lab[b], a); done = old == b; b = old; }
        else if (b < a) { const int old = atomicMin(&lab[a], b); done = old == a; a = old; }
        else done = true;
    }
}

__global__ void __launch_bounds__(256)
uam_k_ccl_tile(const uint8_t* __restrict__ mask, int H, int W, int conn8, int* __restrict__ L) {
    __shared__ int lab[UAM_CCL_T * UAM_CCL_T];
    __shared__ unsigned rowbits[UAM_CCL_T];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = blockIdx.y * UAM_CCL_T, j0 = blockIdx.x * UAM_CCL_T;
    const int j = j0 + lane;
    // rows of the tile: warp w owns rows w, w + 8, w + 16, w + 24
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp + 8 * k, i = i0 + r;
        const bool fg = i < H && j < W && mask[(size_t)i * W + j] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, fg);
        if (lane == 0) rowbits[r] = m;
        int v = -1;
        if (fg) {
            const unsigned below = lane ? (m << (32 - lane)) : 0u;      // bit 31 = lane - 1, ...
            v = r * UAM_CCL_T + lane - (lane ? min(__clz(~below), lane) : 0);
        }
        lab[r * UAM_CCL_T + lane] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp + 8 * k;
        if (r == 0) continue;
        const unsigned cur = rowbits[r], up = rowbits[r - 1];
        const unsigned me = 1u << lane;
        if (!(cur & me)) continue;
        const bool left = lane > 0 && (cur & (me >> 1)), upleft = lane > 0 && (up & (me >> 1));
        const int idx = r * UAM_CCL_T + lane;
        if ((up & me) && !(left && upleft)) uam_sm_union(lab, idx, idx - UAM_CCL_T);
        if (conn8 && !(up & me)) {
            if (upleft && !left) uam_sm_union(lab, idx, idx - UAM_CCL_T - 1);
            if (lane < 31 && (up & (me << 1))) uam_sm_union(lab, idx, idx - UAM_CCL_T + 1);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp + 8 * k, i = i0 + r;
        if (i >= H || j >= W) continue;
        const int v = lab[r * UAM_CCL_T + lane];
        int out = -1;
        if (v >= 0) {
            const int root = uam_sm_find(lab, r * UAM_CCL_T + lane);
            out = (i0 + (root >> 5)) * W + j0 + (root & 31);
        }
        L[(size_t)i * W + j] = out;
    }
}

// the pairs of uam_k_ccl_merge's rules that cross a tile border: border rows (row % 32 == 0) and border columns
__global__ void __launch_bounds__(256)
uam_k_ccl_borders(const uint8_t* __restrict__ mask, int H, int W, int conn8, int* __restrict__ L) {
    const long long n_rows = (long long)((H - 1) / UAM_CCL_T) * W;             // cells of the rows 32, 64, ...
    const long long n_cols = (long long)((W - 1) / UAM_CCL_T) * H;             // cells of the columns 32, 64, ...
    const long long total = n_rows + n_cols * (conn8 ? 2 : 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int row, col, part;
        if (t < n_rows) { part = 0; row = (int)(t / W + 1) * UAM_CCL_T; col = (int)(t % W); }
        else if (t < n_rows + n_cols) { part = 1; const long long u = t - n_rows; col = (int)(u / H + 1) * UAM_CCL_T; row = (int)(u % H); }
        else { part = 2; const long long u = t - n_rows - n_cols; col = (int)(u / H + 1) * UAM_CCL_T - 1; row = (int)(u % H); }
        const long long i = (long long)row * W + col;
        if (mask[i] == 0) continue;
        const bool left = col > 0 && mask[i - 1] != 0;
        if (part == 0) {
            const bool up = mask[i - W] != 0;
            const bool upleft = col > 0 && mask[i - W - 1] != 0;
            if (up && !(left && upleft)) uam_uf_union(L, (int)i, (int)(i - W));
            if (conn8 && !up) {
                if (upleft && !left) uam_uf_union(L, (int)i, (int)(i - W - 1));
                if (col + 1 < W && mask[i - W + 1] != 0) uam_uf_union(L, (int)i, (int)(i - W + 1));
            }
        } else if (part == 1) {
            if (left) uam_uf_union(L, (int)i, (int)(i - 1));
            // the up-left diagonal across the vertical border (rows on a horizontal border were handled by part 0)
            if (conn8 && row % UAM_CCL_T != 0 && !left && mask[i - W] == 0 && mask[i - W - 1] != 0) uam_uf_union(L, (int)i, (int)(i - W - 1));
        } else {
            // column 32 k - 1: the up-right diagonal across the vertical border
            if (row % UAM_CCL_T != 0 && col + 1 < W && mask[i - W] == 0 && mask[i - W + 1] != 0) uam_uf_union(L, (int)i, (int)(i - W + 1));
        }
    }
}

__global__ void __launch_bounds__(256)
uam_k_ccl_flatten(long long n, int* __restrict__ L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = L[i];
    if (v >= 0) L[i] = uam_uf_find(L, (int)i);
}

// ---- exclusive scan of a per-element count (three passes; element order = index order) ------------------------------
#define UAM_SCAN_PER_THREAD 16
#define UAM_SCAN_CHUNK (256 * UAM_SCAN_PER_THREAD)

struct UamIsRoot {
    const int* L;
    __device__ int operator()(long long i) const { return L[i] == (int)i ? 1 : 0; }
};
struct UamIntArray {
    const int* v;
    __device__ int operator()(long long i) const { return v[i]; }
};

template <class F>
__global__ void __launch_bounds__(256)
uam_k_scan_counts(F f, long long n, unsigned long long* __restrict__ block_sum) {
    __shared__ unsigned long long warp_sum[8];
    const long long base = (long long)blockIdx.x * UAM_SCAN_CHUNK + (long long)threadIdx.x * UAM_SCAN_PER_THREAD;
    unsigned long long c = 0;
    for (int k = 0; k < UAM_SCAN_PER_THREAD; ++k)
        if (base + k < n) c += (unsigned long long)f(base + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sum[w];
        block_sum[blockIdx.x] = t;
    }
}

// exclusive scan of block_sum[0..nb) in place, total to block_sum[nb] (single CTA)
__global__ void __launch_bounds__(1024)
uam_k_scan_blocks(unsigned long long* __restrict__ block_sum, int nb) {
    __shared__ unsigned long long warp_tot[32];
    const int per = (nb + 1023) / 1024;
    const int lo = min((int)threadIdx.x * per, nb), hi = min(lo + per, nb);
    unsigned long long local = 0;
    for (int i = lo; i < hi; ++i) local += block_sum[i];
    unsigned long long incl = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = warp_tot[lane];
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;
    }
    __syncthreads();
    unsigned long long run = warp_tot[warp] + incl - local;
    for (int i = lo; i < hi; ++i) {
        const unsigned long long c = block_sum[i];
        block_sum[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) block_sum[nb] = run;
}

// third pass: out(i, exclusive prefix) for every element with a non-zero count
template <class F, class OUT>
__global__ void __launch_bounds__(256)
uam_k_scan_apply(F f, long long n, const unsigned long long* __restrict__ block_sum, OUT out) {
    __shared__ unsigned long long warp_sum[8];
    const long long base = (long long)blockIdx.x * UAM_SCAN_CHUNK + (long long)threadIdx.x * UAM_SCAN_PER_THREAD;
    int cnt[UAM_SCAN_PER_THREAD];
    unsigned long long c = 0;
    for (int k = 0; k < UAM_SCAN_PER_THREAD; ++k) {
        cnt[k] = base + k < n ? f(base + k) : 0;
        c += (unsigned long long)cnt[k];
    }
    unsigned long long incl = c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long run = block_sum[blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) run += warp_sum[w];
    for (int k = 0; k < UAM_SCAN_PER_THREAD; ++k) {
        if (base + k < n) out(base + k, run, cnt[k]);
        run += (unsigned long long)cnt[k];
    }
}

struct UamRootCode {       // a root gets the code -2 - rank (rank = number of roots before it in raster-scan order)
    int* L;
    __device__ void operator()(long long i, unsigned long long prefix, int cnt) const {
        if (cnt) L[i] = -2 - (int)prefix;
    }
};
struct UamOffsets {
    long long* off;
    __device__ void operator()(long long i, unsigned long long prefix, int) const { off[i] = (long long)prefix; }
};

// labels: 0 = background, 1 + rank of the component's root
__global__ void __launch_bounds__(256)
uam_k_ccl_relabel(long long n, const int* __restrict__ L, int32_t* __restrict__ labels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = L[i];
    int lab = 0;
    if (v < -1) lab = -1 - v;                 // a root: code -2 - rank -> 1 + rank
    else if (v >= 0) lab = -1 - L[v];         // v is a root index (flattened); its entry holds the code
    labels[i] = lab;
}

// ---- per-component statistics ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
uam_k_comp_init(int n_comp, unsigned long long* __restrict__ area, int* __restrict__ bbox) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_comp) return;
    area[c] = 0ull;
    bbox[4 * c + 0] = INT_MAX; bbox[4 * c + 1] = -1; bbox[4 * c + 2] = INT_MAX; bbox[4 * c + 3] = -1;   // rmin rmax cmin cmax
}

// A CTA owns 4096 consecutive cells, a thread 16 consecutive ones.  A thread keeps one open entry {label, cells, bbox}
// while it walks its cells (a row run keeps its label), entries closed early go straight to the global atomics (region
// borders only); the entries still open at the end are combined per warp (match_any) and then per CTA (a short list in
// shared memory), so a region that covers whole CTAs costs 5 atomics per 4096 cells instead of 5 per warp.
#define UAM_STATS_PER_THREAD 16
struct UamCompEntry {
    int lab, rmin, rmax, cmin, cmax;
    unsigned cnt;
};

__device__ __forceinline__ void uam_comp_flush(const UamCompEntry& e, unsigned long long* __restrict__ area, int* __restrict__ bbox) {
    const int c = e.lab - 1;
    atomicAdd(&area[c], (unsigned long long)e.cnt);
    atomicMin(&bbox[4 * c + 0], e.rmin); atomicMax(&bbox[4 * c + 1], e.rmax);
    atomicMin(&bbox[4 * c + 2], e.cmin); atomicMax(&bbox[4 * c + 3], e.cmax);
}

__global__ void __launch_bounds__(256)
uam_k_comp_stats(const int32_t* __restrict__ labels, int H, int W, int n_comp, unsigned long long* __restrict__ area,
                 int* __restrict__ bbox, unsigned* __restrict__ bad) {
    __shared__ UamCompEntry s_ent[8];
    const long long n = (long long)H * W;
    const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * UAM_STATS_PER_THREAD;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    UamCompEntry cur;
    cur.lab = 0; cur.cnt = 0; cur.rmin = cur.rmax = cur.cmin = cur.cmax = 0;
    if (base < n) {
        int row = (int)(base / W), col = (int)(base - (long long)row * W);
        for (int k = 0; k < UAM_STATS_PER_THREAD && base + k < n; ++k) {
            int lab = labels[base + k];
            if (lab < 0 || lab > n_comp) { atomicOr(bad, 1u); lab = 0; }
            if (lab != cur.lab) {
                if (cur.lab) uam_comp_flush(cur, area, bbox);
                cur.lab = lab; cur.cnt = 0; cur.rmin = cur.rmax = row; cur.cmin = cur.cmax = col;
            }
            if (lab) {
                cur.cnt += 1;
                cur.rmax = row;                       // rows only grow along the walk
                cur.cmin = min(cur.cmin, col); cur.cmax = max(cur.cmax, col);
            }
            if (++col == W) { col = 0; ++row; }
        }
    }
    // per warp: lanes whose open entries carry the same label combine
    const unsigned peers = __match_any_sync(0xffffffffu, cur.lab);
    const int leader = __ffs(peers) - 1;
    // every lane publishes its entry in turn, the leader of a group folds its peers
    UamCompEntry tot = cur;
    for (int src = 0; src < 32; ++src) {
        const int l = __shfl_sync(0xffffffffu, cur.lab, src);
        const unsigned c = __shfl_sync(0xffffffffu, cur.cnt, src);
        const int r0 = __shfl_sync(0xffffffffu, cur.rmin, src), r1 = __shfl_sync(0xffffffffu, cur.rmax, src);
        const int c0 = __shfl_sync(0xffffffffu, cur.cmin, src), c1 = __shfl_sync(0xffffffffu, cur.cmax, src);
        if (lane == leader && src != lane && l == cur.lab && cur.lab) {
            tot.cnt += c;
            tot.rmin = min(tot.rmin, r0); tot.rmax = max(tot.rmax, r1);
            tot.cmin = min(tot.cmin, c0); tot.cmax = max(tot.cmax, c1);
        }
    }
    // per CTA: the first leader of each warp whose label equals the warp-0 entry's label joins it, the others go to global
    const bool is_leader = cur.lab != 0 && lane == leader;
    // the lowest-lane leader of the warp represents it in shared memory
    const unsigned leaders = __ballot_sync(0xffffffffu, is_leader);
    const bool rep = is_leader && lane == __ffs(leaders) - 1;
    if (lane == 0) s_ent[warp].lab = 0;
    __syncwarp();
    if (rep) s_ent[warp] = tot;
    else if (is_leader) uam_comp_flush(tot, area, bbox);
    __syncthreads();
    if (threadIdx.x < 8) {
        // entry w is folded into the first earlier entry with the same label; the first of each label flushes the sum
        const UamCompEntry mine = s_ent[threadIdx.x];
        if (mine.lab) {
            bool first = true;
            for (int w = 0; w < (int)threadIdx.x; ++w) first = first && s_ent[w].lab != mine.lab;
            if (first) {
                UamCompEntry sum = mine;
                for (int w = threadIdx.x + 1; w < 8; ++w) {
                    const UamCompEntry o = s_ent[w];
                    if (o.lab == mine.lab) {
                        sum.cnt += o.cnt;
                        sum.rmin = min(sum.rmin, o.rmin); sum.rmax = max(sum.rmax, o.rmax);
                        sum.cmin = min(sum.cmin, o.cmin); sum.cmax = max(sum.cmax, o.cmax);
                    }
                }
                uam_comp_flush(sum, area, bbox);
            }
        }
    }
}

// ---- minimum-area rectangles -------------------------------------------------------------------------------------
// slot[label] = position of the component in the caller's list (-1 = not asked for); heights = grid lines it spans
__global__ void uam_k_rect_slots(const int32_t* __restrict__ ids, int K, int n_comp, const int* __restrict__ bbox,
                                 int* __restrict__ slot, int* __restrict__ lines, unsigned* __restrict__ bad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int id = ids[k];
    if (id < 1 || id > n_comp || bbox[4 * (id - 1) + 1] < 0) { atomicOr(bad, 2u); lines[k] = 0; return; }
    if (atomicExch(&slot[id], k) != -1) atomicOr(bad, 4u);       // listed twice
    lines[k] = bbox[4 * (id - 1) + 1] - bbox[4 * (id - 1) + 0] + 2;      // grid lines rmin .. rmax + 1
}

__global__ void __launch_bounds__(256)
uam_k_fill_i32(int* __restrict__ p, long long n, int v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Column extremes per grid line: a cell (row, col) of component c touches the grid lines y = row and y = row + 1 with corner
// x from col to col + 1.  Only run ends write (the left end gives the minimum, the right end the maximum).
__global__ void __launch_bounds__(256)
uam_k_rect_extremes(const int32_t* __restrict__ labels, int H, int W, const int* __restrict__ slot,
                    const int32_t* __restrict__ ids, const int* __restrict__ bbox, const long long* __restrict__ off,
                    int* __restrict__ xl, int* __restrict__ xr) {
    const long long n = (long long)H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lab = labels[i];
    if (lab == 0) return;
    const int k = slot[lab];
    if (k < 0) return;
    const int row = (int)(i / W), col = (int)(i - (long long)row * W);
    const bool left_end = col == 0 || labels[i - 1] != lab;
    const bool right_end = col + 1 == W || labels[i + 1] != lab;
    if (!left_end && !right_end) return;
    const long long o = off[k] + (row - bbox[4 * (lab - 1)]);
    if (left_end) { atomicMin(&xl[o], col); atomicMin(&xl[o + 1], col); }
    if (right_end) { atomicMax(&xr[o], col + 1); atomicMax(&xr[o + 1], col + 1); }
}

__device__ __forceinline__ long long uam_cross(long long ax, long long ay, long long bx, long long by) { return ax * by - ay * bx; }

// One CTA per listed component.  hull_all: scratch for 4 * lines + 2 vertices (int2) per component at 4 * off[k] + 2 * k.
// rect out: 4 corners (x, y) in world coordinates, counter-clockwise in pixel space starting at the corner that lies on the
// chosen hull edge's line at the smallest projection; info out: hull vertices, chosen edge.
__global__ void __launch_bounds__(128)
uam_k_rect_hull(const int32_t* __restrict__ ids, int K, const int* __restrict__ bbox, const long long* __restrict__ off,
                const int* __restrict__ xl, const int* __restrict__ xr, int2* __restrict__ hull_all, double x0, double dx,
                double y0, double dy, double* __restrict__ rect, int* __restrict__ info) {
    const int k = blockIdx.x;
    if (k >= K) return;
    const int id = ids[k];
    const int rmin = bbox[4 * (id - 1)], lines = bbox[4 * (id - 1) + 1] - rmin + 2;
    const long long o = off[k];
    // scratch of this component: [0, 2 lines) the threads' local chains (left chains in [0, lines), right chains in
    // [lines, 2 lines), each thread inside the slice of its own grid lines), then the final hull (at most 2 lines vertices)
    int2* loc = hull_all + 4 * o + 2 * k;
    int2* hull = loc + 2 * lines;
    // Phase A, parallel: thread t reduces the grid lines [t c, (t + 1) c) to their local chains.  The hull of the union is
    // the hull of the local hulls, so phase B only sees the few points that survive locally.
    // Left chain, top to bottom: x as a function of y must be convex (the hull's left side) -- pop while the turn
    // a -> b -> p is not strictly convex, cross(b - a, p - a) >= 0; the right chain is the mirror image (<= 0).
    const int c = (lines + (int)blockDim.x - 1) / (int)blockDim.x;
    __shared__ int s_nl[128], s_nr[128];
    {
        const int t0 = min((int)threadIdx.x * c, lines), t1 = min(t0 + c, lines);
        int2* L = loc + t0;
        int2* R = loc + lines + t0;
        int nl = 0, nr = 0;
        for (int t = t0; t < t1; ++t) {
            const int2 p = make_int2(xl[o + t], rmin + t);
            while (nl >= 2 && uam_cross(L[nl - 1].x - L[nl - 2].x, L[nl - 1].y - L[nl - 2].y, p.x - L[nl - 2].x, p.y - L[nl - 2].y) >= 0) --nl;
            L[nl++] = p;
            const int2 q = make_int2(xr[o + t], rmin + t);
            while (nr >= 2 && uam_cross(R[nr - 1].x - R[nr - 2].x, R[nr - 1].y - R[nr - 2].y, q.x - R[nr - 2].x, q.y - R[nr - 2].y) <= 0) --nr;
            R[nr++] = q;
        }
        s_nl[threadIdx.x] = nl;
        s_nr[threadIdx.x] = nr;
    }
    __syncthreads();
    // Phase B, thread 0: the local chains in order through the same monotone stacks; the polygon is the left chain
    // downwards followed by the right chain upwards (counter-clockwise on the screen: x right, y down).
    __shared__ int s_nh2;
    if (threadIdx.x == 0) {
        int nl = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const int2* L = loc + min(t * c, lines);
            for (int i = 0; i < s_nl[t]; ++i) {
                const int2 p = L[i];
                while (nl >= 2 && uam_cross(hull[nl - 1].x - hull[nl - 2].x, hull[nl - 1].y - hull[nl - 2].y, p.x - hull[nl - 2].x,
                                            p.y - hull[nl - 2].y) >= 0) --nl;
                hull[nl++] = p;
            }
        }
        // right chain: built top to bottom behind the left chain, then reversed in place
        int2* rc = hull + nl;
        int nr = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const int2* R = loc + lines + min(t * c, lines);
            for (int i = 0; i < s_nr[t]; ++i) {
                const int2 q = R[i];
                while (nr >= 2 && uam_cross(rc[nr - 1].x - rc[nr - 2].x, rc[nr - 1].y - rc[nr - 2].y, q.x - rc[nr - 2].x, q.y - rc[nr - 2].y) <= 0) --nr;
                rc[nr++] = q;
            }
        }
        for (int i = 0, j = nr - 1; i < j; ++i, --j) { const int2 tmp = rc[i]; rc[i] = rc[j]; rc[j] = tmp; }
        // The chains never share a vertex (xr > xl on every grid line) and the turns at the four chain ends are strict
        // (the end points are the extreme points of the top / bottom grid line), so hull[0 .. nl + nr) is the hull.
        s_nh2 = nl + nr;
    }
    __syncthreads();
    const int n = s_nh2;
    // every hull edge as the rectangle's direction: exact integer extents, area = (du * dv) / |e|^2
    unsigned long long best_num = 0, best_den = 0;
    int best_edge = -1;
    long long b_umin = 0, b_umax = 0, b_v = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int2 a = hull[e], b = hull[(e + 1) % n];
        const long long ex = b.x - a.x, ey = b.y - a.y;
        long long umin = 0, umax = 0, vext = 0;          // projections of (p - a) on e and on the inward normal
        for (int q = 0; q < n; ++q) {
            const long long px = hull[q].x - a.x, py = hull[q].y - a.y;
            const long long u = px * ex + py * ey, v = uam_cross(px, py, ex, ey);      // v >= 0 for this orientation
            umin = min(umin, u); umax = max(umax, u);
            vext = max(vext, v < 0 ? -v : v);
        }
        const unsigned long long num = (unsigned long long)(umax - umin) * (unsigned long long)vext;
        const unsigned long long den = (unsigned long long)(ex * ex + ey * ey);
        // num / den < best_num / best_den  <=>  num * best_den < best_num * den   (128-bit products)
        bool better = best_edge < 0;
        if (!better) {
            const unsigned __int128 l = (unsigned __int128)num * best_den, r = (unsigned __int128)best_num * den;
            better = l < r;
        }
        if (better) { best_num = num; best_den = den; best_edge = e; b_umin = umin; b_umax = umax; b_v = vext; }
    }
    // block argmin, ties to the lowest edge index
    __shared__ unsigned long long s_num[128], s_den[128];
    __shared__ int s_edge[128];
    __shared__ long long s_u0[128], s_u1[128], s_v[128];
    s_num[threadIdx.x] = best_num; s_den[threadIdx.x] = best_den; s_edge[threadIdx.x] = best_edge;
    s_u0[threadIdx.x] = b_umin; s_u1[threadIdx.x] = b_umax; s_v[threadIdx.x] = b_v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int w = -1;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            if (s_edge[t] < 0) continue;
            bool better = w < 0;
            if (!better) {
                const unsigned __int128 l = (unsigned __int128)s_num[t] * s_den[w], r = (unsigned __int128)s_num[w] * s_den[t];
                better = l < r || (l == r && s_edge[t] < s_edge[w]);
            }
            if (better) w = t;
        }
        const int e = s_edge[w];
        const int2 a = hull[e], b = hull[(e + 1) % n];
        const double ex = (double)(b.x - a.x), ey = (double)(b.y - a.y), den = (double)s_den[w];
        // inward normal of a hull edge in this orientation: v = cross(p - a, e) >= 0  ->  n = (ey, -ex)
        const double nx = ey, ny = -ex;
        const double u0 = (double)s_u0[w] / den, u1 = (double)s_u1[w] / den, vv = (double)s_v[w] / den;
        const double cx[4] = {a.x + u0 * ex, a.x + u1 * ex, a.x + u1 * ex + vv * nx, a.x + u0 * ex + vv * nx};
        const double cy[4] = {a.y + u0 * ey, a.y + u1 * ey, a.y + u1 * ey + vv * ny, a.y + u0 * ey + vv * ny};
        for (int c = 0; c < 4; ++c) {
            rect[8 * (size_t)k + 2 * c + 0] = x0 + cx[c] * dx;
            rect[8 * (size_t)k + 2 * c + 1] = y0 + cy[c] * dy;
        }
        if (info) { info[2 * k] = n; info[2 * k + 1] = e; }
    }
}

template <class F, class OUT>
int uam_scan(uam_ctx* ctx, F f, long long n, OUT out, unsigned long long* block_sum, unsigned long long* h_total, cudaStream_t st) {
    const int nb = (int)((n + UAM_SCAN_CHUNK - 1) / UAM_SCAN_CHUNK);
    uam_k_scan_counts<<<nb, 256, 0, st>>>(f, n, block_sum);
    UAM_CHECK_LAUNCH(ctx, "uam_k_scan_counts");
    uam_k_scan_blocks<<<1, 1024, 0, st>>>(block_sum, nb);
    UAM_CHECK_LAUNCH(ctx, "uam_k_scan_blocks");
    uam_k_scan_apply<<<nb, 256, 0, st>>>(f, n, block_sum, out);
    UAM_CHECK_LAUNCH(ctx, "uam_k_scan_apply");
    if (h_total) {
        UAM_CUDA(ctx, cudaMemcpyAsync(h_total, block_sum + nb, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        UAM_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return UAM_OK;
}

}  // namespace

extern "C" int uam_label_components(uam_ctx* ctx, const uint8_t* d_mask, int H, int W, int connectivity, int32_t* d_labels,
                                    int32_t* h_n_components, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.map.label_components");
    if (H < 0 || W < 0) return uam_fail(ctx, UAM_ERR_INVALID, "negative raster size");
    if (connectivity != 4 && connectivity != 8) return uam_fail(ctx, UAM_ERR_INVALID, "connectivity must be 4 or 8");
    const long long n = (long long)H * W;
    if (n >= (1ll << 31) - 2) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "rasters of 2^31 cells or more are not supported");
    if (h_n_components) *h_n_components = 0;
    if (n == 0) return UAM_OK;
    if (!d_mask || !d_labels) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const int nb = (int)((n + UAM_SCAN_CHUNK - 1) / UAM_SCAN_CHUNK);
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, (size_t)n * 4 + (size_t)(nb + 2) * 8 + 256));
    unsigned long long* block_sum = (unsigned long long*)ctx->d_scratch;
    int* L = (int*)(block_sum + nb + 2);
    const unsigned ctas = (unsigned)((n + 255) / 256);
    if (ctx->ccl_tiles && (H + UAM_CCL_T - 1) / UAM_CCL_T <= 65535) {
        // tile-local labelling in shared memory, then only the pairs across tile borders are united in HBM
        dim3 tgrid((W + UAM_CCL_T - 1) / UAM_CCL_T, (H + UAM_CCL_T - 1) / UAM_CCL_T);
        uam_k_ccl_tile<<<tgrid, 256, 0, st>>>(d_mask, H, W, connectivity == 8 ? 1 : 0, L);
        UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_tile");
        uam_k_ccl_borders<<<ctx->sm_count * 16, 256, 0, st>>>(d_mask, H, W, connectivity == 8 ? 1 : 0, L);
        UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_borders");
    } else {
        uam_k_ccl_init<<<ctas, 256, 0, st>>>(d_mask, H, W, L);
        UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_init");
        uam_k_ccl_merge<<<ctas, 256, 0, st>>>(d_mask, H, W, connectivity == 8 ? 1 : 0, L);
        UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_merge");
    }
    uam_k_ccl_flatten<<<ctas, 256, 0, st>>>(n, L);
    UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_flatten");
    unsigned long long total = 0;
    UAM_TRY(uam_scan(ctx, UamIsRoot{L}, n, UamRootCode{L}, block_sum, &total, st));
    uam_k_ccl_relabel<<<ctas, 256, 0, st>>>(n, L, d_labels);
    UAM_CHECK_LAUNCH(ctx, "uam_k_ccl_relabel");
    if (h_n_components) *h_n_components = (int32_t)total;
    return UAM_OK;
}

extern "C" int uam_component_stats(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int n_components, int64_t* d_area,
                                   int32_t* d_bbox, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.map.component_stats");
    if (H < 0 || W < 0 || n_components < 0) return uam_fail(ctx, UAM_ERR_INVALID, "negative size");
    if (n_components == 0) return UAM_OK;
    if (!d_labels || !d_area || !d_bbox) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const long long n = (long long)H * W;
    UAM_TRY(uam_reserve(ctx, &ctx->d_cull_scratch, &ctx->cull_scratch_bytes, 256));
    unsigned* bad = (unsigned*)ctx->d_cull_scratch;
    UAM_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, st));
    uam_k_comp_init<<<(n_components + 255) / 256, 256, 0, st>>>(n_components, (unsigned long long*)d_area, d_bbox);
    UAM_CHECK_LAUNCH(ctx, "uam_k_comp_init");
    if (n > 0) {
        const long long per_cta = 256ll * UAM_STATS_PER_THREAD;
        uam_k_comp_stats<<<(unsigned)((n + per_cta - 1) / per_cta), 256, 0, st>>>(d_labels, H, W, n_components, (unsigned long long*)d_area, d_bbox, bad);
        UAM_CHECK_LAUNCH(ctx, "uam_k_comp_stats");
    }
    unsigned h_bad = 0;
    UAM_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad) return uam_fail(ctx, UAM_ERR_INVALID, "labels outside 0 .. n_components");
    return UAM_OK;
}

extern "C" int uam_component_rects(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int n_components, const int32_t* d_bbox,
                                   const int32_t* d_ids, int K, double x0, double dx, double y0, double dy, double* d_rect,
                                   int32_t* d_info, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.map.component_rects");
    if (H < 0 || W < 0 || n_components < 0 || K < 0) return uam_fail(ctx, UAM_ERR_INVALID, "negative size");
    if (K == 0) return UAM_OK;
    if (!d_labels || !d_bbox || !d_ids || !d_rect) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    if (!(fabs(fabs(dx) - fabs(dy)) <= 1e-12 * fabs(dx)) || dx == 0.0)
        return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "minimum-area rectangles need square cells (|dx| == |dy|)");
    if (H > 32766 || W > 32766) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "rasters beyond 32766 cells per side: the exact area comparison is sized for 2^15");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const long long n = (long long)H * W;
    const int nb = (K + UAM_SCAN_CHUNK - 1) / UAM_SCAN_CHUNK;
    // fixed part of the scratch: flags | block sums | slot[n_components + 1] | lines[K] | off[K + 1]
    const size_t fixed = 256 + (size_t)(nb + 2) * 8 + (size_t)(n_components + 1) * 4 + (size_t)K * 4 + 16 + (size_t)(K + 1) * 8;
    UAM_TRY(uam_reserve(ctx, &ctx->d_cull_scratch, &ctx->cull_scratch_bytes, fixed));
    unsigned* bad = (unsigned*)ctx->d_cull_scratch;
    unsigned long long* block_sum = (unsigned long long*)((char*)ctx->d_cull_scratch + 256);
    long long* off = (long long*)(block_sum + nb + 2);
    int* slot = (int*)(off + K + 1);
    int* lines = slot + n_components + 1;
    UAM_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, st));
    uam_k_fill_i32<<<(n_components + 1 + 255) / 256, 256, 0, st>>>(slot, n_components + 1, -1);
    UAM_CHECK_LAUNCH(ctx, "uam_k_fill_i32");
    uam_k_rect_slots<<<(K + 255) / 256, 256, 0, st>>>(d_ids, K, n_components, d_bbox, slot, lines, bad);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rect_slots");
    unsigned long long total = 0;
    UAM_TRY(uam_scan(ctx, UamIntArray{lines}, (long long)K, UamOffsets{off}, block_sum, &total, st));
    unsigned h_bad = 0;
    UAM_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad) return uam_fail(ctx, UAM_ERR_INVALID, "component ids must be distinct labels in 1 .. n_components of non-empty components");
    // per-line extremes (2 x total ints) + chain / hull scratch (4 * total + 2 * K int2)
    const size_t need = (size_t)total * 8 + ((size_t)total * 4 + (size_t)K * 2) * 8 + 256;
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, need));
    int* xl = (int*)ctx->d_scratch;
    int* xr = xl + total;
    int2* hull = (int2*)(xr + total);           // 8 * total bytes in: still 8-byte aligned
    uam_k_fill_i32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xl, (long long)total, INT_MAX);
    UAM_CHECK_LAUNCH(ctx, "uam_k_fill_i32");
    uam_k_fill_i32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xr, (long long)total, -1);
    UAM_CHECK_LAUNCH(ctx, "uam_k_fill_i32");
    uam_k_rect_extremes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_labels, H, W, slot, d_ids, d_bbox, off, xl, xr);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rect_extremes");
    uam_k_rect_hull<<<K, 128, 0, st>>>(d_ids, K, d_bbox, off, xl, xr, hull, x0, dx, y0, dy, d_rect, d_info);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rect_hull");
    return UAM_OK;
}

// ---- large-polygon split (map_generation/data_processor.py:34-53) -----------------------------------------------------
// The reference cuts a polygon larger than `large_area` with a divisions x divisions grid of boxes over its bounding box
// and approximates every piece of polygon.intersection(box) by its own rectangle.  The box edges fall inside cells; on a
// grid refined `div` times they fall on (sub-)cell boundaries, so box (bj, bk) of a component whose bounding box is nr x nc
// cells is exactly the nr x nc block of sub-cells starting at sub-row bk * nr, sub-column bj * nc, and sub-cell (r, c) of
// the block lies in cell (row0 + (bk * nr + r) / div, col0 + (bj * nc + c) / div).  This kernel writes that block's mask;
// its 4-connected regions are the pieces (two clipped cells are connected iff they share an edge of positive length inside
// the box), and the hull of a piece's sub-cell corners is the hull of the clipped polygon's exterior ring.
__global__ void __launch_bounds__(256)
uam_k_component_submask(const int32_t* __restrict__ labels, int W, int label, int row0, int col0, int sub_row0, int sub_col0,
                        int nr, int nc, int div, uint8_t* __restrict__ mask) {
    const long long n = (long long)nr * nc;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const int r = (int)(t / nc), c = (int)(t - (long long)r * nc);
        const int i = row0 + (sub_row0 + r) / div, j = col0 + (sub_col0 + c) / div;
        mask[t] = labels[(size_t)i * W + j] == label ? 1 : 0;
    }
}

extern "C" int uam_component_submask(uam_ctx* ctx, const int32_t* d_labels, int H, int W, int label, const int32_t* h_bbox,
                                     int divisions, int box_row, int box_col, uint8_t* d_mask, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (!d_labels || !h_bbox || !d_mask || divisions < 1 || box_row < 0 || box_row >= divisions || box_col < 0 || box_col >= divisions)
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_component_submask");
    const int r0 = h_bbox[0], r1 = h_bbox[1], c0 = h_bbox[2], c1 = h_bbox[3];
    if (r0 < 0 || r1 >= H || c0 < 0 || c1 >= W || r1 < r0 || c1 < c0) return uam_fail(ctx, UAM_ERR_INVALID, "bounding box outside the raster");
    const int nr = r1 - r0 + 1, nc = c1 - c0 + 1;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long n = (long long)nr * nc;
    const long long ctas = std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    uam_k_component_submask<<<(unsigned)ctas, 256, 0, uam_pick_stream(ctx, stream)>>>(d_labels, W, label, r0, c0, box_row * nr, box_col * nc,
                                                                                      nr, nc, divisions, d_mask);
    UAM_CHECK_LAUNCH(ctx, "uam_k_component_submask");
    return UAM_OK;
}
