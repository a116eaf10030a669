/* C restatement of the raster path scorer and the grid search of oracle/uam_oracle.py -- TEST INFRASTRUCTURE ONLY.
 *
 * Same operation order as the numpy oracle (which is pinned against the reference's own Python, see
 * tests/test_oracle_golden.py), one path / one query per OpenMP thread, so that
 *   - parity can be checked against the CUDA path at BASELINE.json's FULL sizes (the numpy oracle needs minutes
 *     for a few thousand paths), and
 *   - the CPU baseline of bench.py can use every host core.
 * Compiled by oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off: no FMA contraction, the numpy oracle has none).
 * Follows: Problem.get_cost / length_of (reference path_generation/problem.py:38-44,130-146) for the cost functional
 * and its length term; the raster sampling and the grid search are build-defined extensions (no reference
 * counterpart) defined by oracle/uam_oracle.py::score_paths_raster / grid_search.
 * Nothing in the product (uam_path_planning_b200/) links or calls this file. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

/* oracle/uam_oracle.py::sample_uv: bilinear value per layer (float64 maths on float32 texels, fraction rounded to
 * float32) weighted by w, + nearest-cell occupancy; (u, v) already clamped to [0, W-1] x [0, H-1] */
static inline double sample_uv(const float* layers, const uint8_t* occ, int L, int H, int W, const double* w, double u,
                               double v, int* occupied) {
    long j0 = (long)fmin(floor(u), (double)(W - 2));
    long i0 = (long)fmin(floor(v), (double)(H - 2));
    const double fx = (double)(float)(u - (double)j0);
    const double fy = (double)(float)(v - (double)i0);
    double acc = 0.0;
    for (int l = 0; l < L; ++l) {
        const float* p = layers + (size_t)l * H * W;
        const double t00 = p[(size_t)i0 * W + j0], t01 = p[(size_t)i0 * W + j0 + 1];
        const double t10 = p[(size_t)(i0 + 1) * W + j0], t11 = p[(size_t)(i0 + 1) * W + j0 + 1];
        const double top = t00 + fx * (t01 - t00);
        const double bot = t10 + fx * (t11 - t10);
        acc += w[l] * (top + fy * (bot - top));
    }
    const long jn = j0 + (fx >= 0.5), im = i0 + (fy >= 0.5);
    *occupied = occ ? occ[(size_t)im * W + jn] != 0 : 0;
    return acc;
}

/* oracle/uam_oracle.py::score_paths_raster.  Z (B, 2*Wp) interleaved xy incl. start and goal; x_start nullable.
 * Returns 0; cost (B) float64, collide (B) uint8, nsamples (B) int64 (each nullable). */
int uam_oc_score_paths_raster(const float* layers, const uint8_t* occ, int L, int H, int W, double x0, double dx, double y0,
                              double dy, const double* Z, int64_t B, int Wp, const double* w, double spc, int length_smooth,
                              const double* x_start, double* cost, uint8_t* collide, int64_t* nsamples, int nthreads) {
    const int N = Wp - 2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t b = 0; b < B; ++b) {
        const double* P = Z + (size_t)b * 2 * Wp;
        /* length term (quirk Q1): |z_0 - m_s| + sum_{k=0}^{N-1} |dz_k| */
        double Lsum = 0.0;
        for (int k = 0; k < N; ++k) {
            const double ex = P[2 * (k + 1)] - P[2 * k], ey = P[2 * (k + 1) + 1] - P[2 * k + 1];
            const double d = sqrt(ex * ex + ey * ey);
            Lsum += length_smooth ? d * d : d;
        }
        if (x_start) {
            const double ex = P[0] - x_start[0], ey = P[1] - x_start[1];
            const double d0 = sqrt(ex * ex + ey * ey);
            Lsum = Lsum + (length_smooth ? d0 * d0 : d0);
        }
        double pen = 0.0;
        int col = 0;
        int64_t ns = 1;
        double Uk = (P[0] - x0) / dx - 0.5, Vk = (P[1] - y0) / dy - 0.5;
        for (int k = 0; k < Wp - 1; ++k) {
            const double Un = (P[2 * (k + 1)] - x0) / dx - 0.5, Vn = (P[2 * (k + 1) + 1] - y0) / dy - 0.5;
            const double dU = Un - Uk, dV = Vn - Vk;
            int64_t S = 1;
            if (spc > 0.0) S = (int64_t)fmax(1.0, ceil(sqrt(dU * dU + dV * dV) * spc));
            const double su = dU / (double)S, sv = dV / (double)S;
            double acc = 0.0;
            for (int64_t s = 0; s < S; ++s) {
                const double u = clampd(Uk + (double)s * su, 0.0, (double)(W - 1));
                const double v = clampd(Vk + (double)s * sv, 0.0, (double)(H - 1));
                int o;
                acc += sample_uv(layers, occ, L, H, W, w, u, v, &o);
                col |= o;
            }
            pen += acc / (double)S;
            ns += S;
            Uk = Un;
            Vk = Vn;
        }
        {
            int o;
            pen += sample_uv(layers, occ, L, H, W, w, clampd(Uk, 0.0, (double)(W - 1)), clampd(Vk, 0.0, (double)(H - 1)), &o);
            col |= o;
        }
        if (cost) cost[b] = (double)(N + 1) * Lsum + pen / (double)N;
        if (collide) collide[b] = (uint8_t)col;
        if (nsamples) nsamples[b] = ns;
    }
    return 0;
}

/* ---- grid search: oracle/uam_oracle.py::grid_search (binary-heap Dijkstra + the fixed-order parent post-pass) -------- */
typedef struct { int64_t d; int32_t v; } HeapEnt;

static void heap_push(HeapEnt** h, size_t* n, size_t* cap, int64_t d, int32_t v) {
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 1024;
        *h = (HeapEnt*)realloc(*h, *cap * sizeof(HeapEnt));
    }
    size_t i = (*n)++;
    while (i > 0) {
        const size_t p = (i - 1) / 2;
        if ((*h)[p].d <= d) break;
        (*h)[i] = (*h)[p];
        i = p;
    }
    (*h)[i].d = d;
    (*h)[i].v = v;
}

static HeapEnt heap_pop(HeapEnt* h, size_t* n) {
    const HeapEnt top = h[0];
    const HeapEnt last = h[--(*n)];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && h[c + 1].d < h[c].d) ++c;
        if (h[c].d >= last.d) break;
        h[i] = h[c];
        i = c;
    }
    if (*n) h[i] = last;
    return top;
}

static const int NB_DI[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static const int NB_DJ[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
static const int NB_ST[8] = {3, 2, 3, 2, 2, 3, 2, 3};

/* neighbours of v in slot order: 8 in-plane, band below, band above; returns how many */
static inline int neighbours(int32_t v, int bands, int H, int W, int32_t* out, int* step) {
    const int32_t cells = H * W;
    const int vb = v / cells, vc = v - vb * cells, vi = vc / W, vj = vc - vi * W;
    int n = 0;
    for (int s = 0; s < 8; ++s) {
        const int ui = vi + NB_DI[s], uj = vj + NB_DJ[s];
        if (ui >= 0 && ui < H && uj >= 0 && uj < W) { out[n] = vb * cells + ui * W + uj; step[n++] = NB_ST[s]; }
    }
    if (vb > 0) { out[n] = v - cells; step[n++] = 2; }
    if (vb + 1 < bands) { out[n] = v + cells; step[n++] = 2; }
    return n;
}

/* one query: source = flat node index (band, row, col); dist (nodes) int64, parent (nodes) int32 nullable */
static void grid_search_one(const uint16_t* cost, const uint8_t* blocked, int bands, int H, int W, int32_t src, int64_t* dist,
                            int32_t* parent) {
    const int64_t INF = (int64_t)1 << 62;
    const int32_t nodes = bands * H * W;
    for (int32_t v = 0; v < nodes; ++v) dist[v] = INF;
    if (src >= 0 && src < nodes && !(blocked && blocked[src])) {
        HeapEnt* heap = NULL;
        size_t n = 0, cap = 0;
        dist[src] = 0;
        heap_push(&heap, &n, &cap, 0, src);
        int32_t nb[10];
        int st[10];
        while (n) {
            const HeapEnt e = heap_pop(heap, &n);
            if (e.d != dist[e.v]) continue;
            const int m = neighbours(e.v, bands, H, W, nb, st);
            for (int k = 0; k < m; ++k) {
                const int32_t u = nb[k];
                if (blocked && blocked[u]) continue;
                const int64_t nd = e.d + (int64_t)st[k] * ((int64_t)cost[e.v] + (int64_t)cost[u]);
                if (nd < dist[u]) {
                    dist[u] = nd;
                    heap_push(&heap, &n, &cap, nd, u);
                }
            }
        }
        free(heap);
    }
    if (!parent) return;
    for (int32_t v = 0; v < nodes; ++v) {
        int32_t p = -1;
        if (dist[v] < INF) {
            if (v == src) p = v;
            else {
                int32_t nb[10];
                int st[10];
                const int m = neighbours(v, bands, H, W, nb, st);
                int64_t best = INF;
                for (int k = 0; k < m; ++k) {
                    const int32_t u = nb[k];
                    if (dist[u] >= INF) continue;
                    const int64_t nd = dist[u] + (int64_t)st[k] * ((int64_t)cost[u] + (int64_t)cost[v]);
                    if (nd < best) { best = nd; p = u; }
                }
            }
        }
        parent[v] = p;
    }
}

/* Q queries, one per OpenMP thread.  cost / blocked (bands,H,W); sources (Q,3) int32 (band,row,col);
 * dist (Q,bands,H,W) int64; parent (Q,bands,H,W) int32 nullable. */
int uam_oc_grid_search(const uint16_t* cost, const uint8_t* blocked, int bands, int H, int W, const int32_t* sources, int Q,
                       int64_t* dist, int32_t* parent, int nthreads) {
    const size_t nodes = (size_t)bands * H * W;
    if (nodes >= 0x7fffffffull) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < Q; ++q) {
        const int sb = sources[3 * q], si = sources[3 * q + 1], sj = sources[3 * q + 2];
        int32_t src = -1;
        if (sb >= 0 && sb < bands && si >= 0 && si < H && sj >= 0 && sj < W) src = (int32_t)((size_t)sb * H * W + (size_t)si * W + sj);
        grid_search_one(cost, blocked, bands, H, W, src, dist + (size_t)q * nodes, parent ? parent + (size_t)q * nodes : NULL);
    }
    return 0;
}

int uam_oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
