// mcts.cu — C-ABI implementation of include/ipp_mcts.h for sm_100a: the batched MCTS-zero rollout
// loop.  One warp per tree; every tree of the batch advances by one simulation per
// (simulate_begin, simulate_end) pair:
//
//   mcts_select_kernel   PUCT descent root -> leaf (compute_uct incl. forced playouts at the root,
//                        planning/mcts_zero/mcts.py:280-296), creation of the child of a new edge, and — same launch,
//                        mcts_rollout_body — the reward of the path's NEW prediction step; the steps above it were rolled
//                        out when their edges were created: an edge caches its reward, a node the variances its step left
//                        behind (overlay)
//   -- evaluator call-out (policy/value network; not part of this library) --
//   mcts_expand_kernel   mask + normalise the leaf's priors (mcts.py:196-237), back the value up
//                        (mcts.py:248-265)
//
// Tree storage per tree: `max_nodes` nodes, each with a 16 B header, the dense prior over its W window slots
// (candidate actions = lattice cells within max_valid_action_distance of the node, all altitude levels;
// -1 = invalid action, -2 = already an edge) and its best not-yet-visited slot; plus a pool of at most
// num_simulations EDGES {parent node, slot, prior, Q, N, child} (a simulation adds at most one).  An unvisited
// action has Q = N = 0, so among them the arg max of the PUCT score is the arg max of the prior: a descent
// step compares the node's visited children (a scan of the small edge pool) with that one cached candidate
// instead of scanning W (~1 900) slots; the prior array is rescanned only at the node that grows a new edge.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/ipp_mcts.h"
#include "engine_internal.h"
#include "step_kernel.cuh"

using namespace ipp;

namespace {

constexpr int kTreeWarps = 4;  // trees per CTA
// experiment switches (measured in DESIGN.md section 9)
#ifndef IPP_MCTS_PUCT_TABLE
#define IPP_MCTS_PUCT_TABLE 1  // exploration constant and sqrt(Ns + 1) of compute_uct from a table by visit count: 0.0710 -> 0.0673 ms
#endif
#ifndef IPP_MCTS_EDGE_CACHE
#define IPP_MCTS_EDGE_CACHE 0  // a tree's edge pool read once per simulation into the lanes' registers (<= 128 edges): 10 % SLOWER
#endif                         // (0.0780 vs 0.0710 ms per simulation: 20 more live registers per lane), kept as a switch

struct TreeDims {
    int T, M, W, D, r, L, H;  // trees, nodes per tree, window slots, window width, radius, levels, episode horizon
    int Wp;                   // W rounded up to whole 16-byte groups: stride of a node's prior row and of the slot tables
    int E;                    // edge-pool capacity per tree
    int max_path;             // H + 1
    int first_env;
    float c_init, c_base, gamma, forced_k, max_dist, dir_eps;
};

struct __align__(16) Edge {
    int parent, slot;
    float prior, q;
    int n, child;
    int pad[2];  // [0] action id (cached when the child node is created), [1] reward bits (cached by the edge's first rollout)
};
static_assert(sizeof(Edge) == 32, "Edge layout");
constexpr int kNoNode = -1;

struct TreeArrays {
    int4 *hdr;       // [T][M]  {packed position, budget bits, Ns, depth | expanded << 8}
    float *P;        // [T][M][W]  masked priors BEFORE normalisation (policy * mask, root: mixed with the noise); -1 invalid, -2 visited
                     //            (prior moved into the edge).  Normalised prior = P * pscale (mcts.py:228-234).
    float *pscale;   // [T][M]     1 / sum of the node's row; negative: the row sums to 0 and every valid action has prior -pscale
    int2 *bu;        // [T][M]  best unvisited slot of the node {slot (-1: none), prior bits}
    Edge *edges;     // [T][E]
    int *n_edges;    // [T]
    int *n_nodes;    // [T]
    double *root_pose;  // [T][3]
    // per-simulation scratch
    int *path_edge;    // [T][max_path]
    int *path_action;  // [T][max_path]  (-1 padded)
    float *path_reward;  // [T][max_path]
    int *leaf;           // [T][IPP_MCTS_LEAF_WORDS]
    const int *slot_info;   // [W]     slot -> level | a << 8 | b << 16 (window column / row index)
    const float *slot_dist; // [L][W]  distance from a LATTICE node at level ln to slot s (valid when the lattice arithmetic is exact)
    int tables_ok;
    float2 *puct;        // [M + 1]  {exploration constant, sqrt(Ns + 1)} of compute_uct by node visit count (mcts.py:282-284)
    float *overlay;      // [T][M][tile]  variances of a node's footprint after its prediction step (row-major, pitch = its nx)
    int tile;            // floats per overlay (largest footprint)
};

__device__ __forceinline__ int pack_pos(int col, int row, int lvl) { return col | (row << 12) | ((lvl + 1) << 24); }
__device__ __forceinline__ void unpack_pos(int w, int &col, int &row, int &lvl) {
    col = w & 0xFFF;
    row = (w >> 12) & 0xFFF;
    lvl = (w >> 24) - 1;
}

// pose of a node: the root keeps the caller's previous action (any pose), every other node sits on the action lattice
__device__ __forceinline__ void node_pose(const StepParams &p, const TreeArrays &a, int t, int node, int col, int row, int lvl, double &x, double &y,
                                          double &h) {
    if (node == 0) {
        x = a.root_pose[3 * t];
        y = a.root_pose[3 * t + 1];
        h = a.root_pose[3 * t + 2];
    } else {
        x = __dadd_rn(__dmul_rn(p.res, (double)col), __dmul_rn(0.5, p.res));
        y = __dadd_rn(__dmul_rn(p.res, (double)row), __dmul_rn(0.5, p.res));
        h = p.lut[lvl].alt;
    }
}

// order-preserving map float -> unsigned (for the hardware warp reductions, REDUX): a < b  <=>  fkey(a) < fkey(b)
__device__ __forceinline__ unsigned fkey(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
// arg max over the warp of (value, lowest slot among equal values): returns the winning lane (every lane takes part)
__device__ __forceinline__ int warp_argmax_lowest_slot(float value, int slot, float &best, int &best_slot) {
    const unsigned k = fkey(value);
    const unsigned kmax = __reduce_max_sync(0xffffffffu, k);
    const int cand = k == kmax ? slot : 0x7fffffff;
    best_slot = __reduce_min_sync(0xffffffffu, cand);
    best = fkey_inv(kmax);
    return __ffs(__ballot_sync(0xffffffffu, k == kmax && slot == best_slot)) - 1;
}

// Per-lane best (largest prior >= 0, lowest slot among equals) of a node's prior row, 16-byte loads over the aligned middle of the
// row (rows start at any 4-byte offset: W is not a multiple of 4).
__device__ __forceinline__ void best_prior_scan(const float *P, int W, int lane, float &bp, int &bs) {
    auto take = [&](float pr, int s) {
        if (pr >= 0.0f && (pr > bp || (pr == bp && s < bs))) {
            bp = pr;
            bs = s;
        }
    };
    const int head = min(W, (int)((4u - (unsigned)(((uintptr_t)P >> 2) & 3u)) & 3u));
    const int nvec = (W - head) >> 2;
    if (lane < head) take(P[lane], lane);
    const float4 *P4 = reinterpret_cast<const float4 *>(P + head);
    for (int v0 = 0; v0 < nvec; v0 += 128) {  // four 16-byte groups per lane per pass, every load issued before the first compare
        float4 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = P4[min(v0 + lane + 32 * j, nvec - 1)];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = v0 + lane + 32 * j;
            if (v >= nvec) continue;
            const int s0 = head + 4 * v;
            take(q[j].x, s0);
            take(q[j].y, s0 + 1);
            take(q[j].z, s0 + 2);
            take(q[j].w, s0 + 3);
        }
    }
    const int tail0 = head + 4 * nvec;
    if (tail0 + lane < W) take(P[tail0 + lane], tail0 + lane);
}

struct Slot {
    int lvl, col, row;
    bool in_grid;
};
__device__ __forceinline__ Slot slot_cell(const TreeDims &d, const StepParams &p, int s, int ccol, int crow) {
    const int DD = d.D * d.D;
    Slot o;
    o.lvl = s / DD;
    const int rem = s - o.lvl * DD;
    const int a = rem / d.D;
    o.col = ccol + a - d.r;
    o.row = crow + (rem - a * d.D) - d.r;
    o.in_grid = o.col >= 0 && o.col < p.X && o.row >= 0 && o.row < p.Y;
    return o;
}
__device__ __forceinline__ void slot_pose(const StepParams &p, const Slot &c, double &x, double &y, double &h) {
    x = __dadd_rn(__dmul_rn(p.res, (double)c.col), __dmul_rn(0.5, p.res));
    y = __dadd_rn(__dmul_rn(p.res, (double)c.row), __dmul_rn(0.5, p.res));
    h = p.lut[c.lvl].alt;
}

// ------------------------------------------------------------------------------------------------
// selection: one warp walks one tree from the root to a leaf
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__device__ __forceinline__ void mcts_rollout_body(const StepParams &p, const TreeDims &d, const TreeArrays &a, uint32_t flags, int t, int lane,
                                                  int4 *s_rect_w, int *s_node_w);

// The rollout of the path's new prediction step (mcts_rollout_body, below) runs at the end of the same launch: it needs nothing
// but the path this warp has just walked, and its memory latency hides under the other warps' descents.
#ifndef IPP_MCTS_SELECT_MINBLOCKS
#define IPP_MCTS_SELECT_MINBLOCKS 8  // resident CTAs per SM the kernel is compiled for (register cap 65536 / (128 * this) = 64): 10 / 12 / 16 spill and are slower
#endif
template <int LAYOUT>
__global__ void __launch_bounds__(kTreeWarps * 32, IPP_MCTS_SELECT_MINBLOCKS) mcts_select_kernel(const __grid_constant__ StepParams p, TreeDims d, TreeArrays a, uint32_t flags,
                                                                      int stage_edges) {
    extern __shared__ int4 s_pool[];                        // [kTreeWarps][E][2]: the trees' edge pools (stage_edges != 0)
    __shared__ int4 s_rect[kTreeWarps][IPP_MCTS_MAX_PATH];  // {xl, yu, nx, ny} of the nodes on the path (rollout)
    __shared__ int s_node[kTreeWarps][IPP_MCTS_MAX_PATH];
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * kTreeWarps + (threadIdx.x >> 5);
    if (t >= d.T) return;
    int4 *hdr = a.hdr + (size_t)t * d.M;
    int node = 0, depth = 0, len = 0, kind = IPP_MCTS_LEAF_TERMINAL;
    int ccol = 0, crow = 0, lvl = -1;
    float budget = 0.0f;
    // The tree's edges do not change while the warp walks down (a new edge ends the descent): when they fit the lanes' registers
    // (<= 32 * kEdgeCache edges: one 16-byte load {parent, slot, prior, q} + the visit count per edge) they are read ONCE per
    // simulation and every level filters them by parent; larger pools are scanned in two passes per level.
    constexpr int kEdgeCache = 4;
    Edge *edges = a.edges + (size_t)t * d.E;
    const int ne = a.n_edges[t];
    // Experiment switch (stage_edges, off by default — see the launch site): the pool does not change while the warp walks down
    // and every level scans all of it, so it can be copied into shared memory once per simulation.
    const Edge *pool = edges;
    if (stage_edges) {
        int4 *mine = s_pool + (size_t)(threadIdx.x >> 5) * d.E * 2;
        const int4 *src = reinterpret_cast<const int4 *>(edges);
        for (int i0 = 0; i0 < 2 * ne; i0 += 128) {
            int4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = src[min(i0 + lane + 32 * j, 2 * ne - 1)];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + lane + 32 * j < 2 * ne) mine[i0 + lane + 32 * j] = v[j];
        }
        __syncwarp();
        pool = reinterpret_cast<const Edge *>(mine);
    }
    const bool cached = (IPP_MCTS_EDGE_CACHE != 0) && ne <= 32 * kEdgeCache;
    int4 ea[kEdgeCache];
    int en[kEdgeCache];
#pragma unroll
    for (int i = 0; i < kEdgeCache; ++i) {
        const int e = lane + 32 * i;
        ea[i] = make_int4(-2, 0, 0, 0);  // parent -2: no node
        en[i] = 0;
        if (cached && e < ne) {
            ea[i] = *reinterpret_cast<const int4 *>(edges + e);
            en[i] = edges[e].n;
        }
    }
    while (true) {
        const int4 h = hdr[node];
        unpack_pos(h.x, ccol, crow, lvl);
        budget = __int_as_float(h.y);
        const int Ns = h.z;
        const bool expanded = ((h.w >> 8) & 1) != 0;
        depth = h.w & 0xFF;
        if (depth > d.H || !(budget > 0.0f)) {  // mcts.py:175-176
            kind = IPP_MCTS_LEAF_TERMINAL;
            break;
        }
        if (!expanded) {  // mcts.py:185: leaf -> evaluator
            kind = IPP_MCTS_LEAF_EVAL;
            break;
        }
        float *P = a.P + ((size_t)t * d.M + node) * d.Wp;
        // normalize_q_values (mcts.py:267-278) over the dense action vector: every action that is not an edge has Q = 0
        float qmin = 0.0f, qmax = 0.0f;
        if (cached) {
#pragma unroll
            for (int i = 0; i < kEdgeCache; ++i) {
                if (ea[i].x == node) {
                    const float q = __int_as_float(ea[i].w);
                    qmin = fminf(qmin, q);
                    qmax = fmaxf(qmax, q);
                }
            }
        } else {
            for (int e = lane; e < ne; e += 32) {
                if (pool[e].parent != node) continue;
                const float q = pool[e].q;
                qmin = fminf(qmin, q);
                qmax = fmaxf(qmax, q);
            }
        }
        qmin = fkey_inv(__reduce_min_sync(0xffffffffu, fkey(qmin)));
        qmax = fkey_inv(__reduce_max_sync(0xffffffffu, fkey(qmax)));
        const float qscale = qmax > qmin ? 1.0f / (qmax - qmin) : 0.0f;  // all zero -> values unchanged (= 0)
        // compute_uct (mcts.py:280-296): the node's edges, then its best unvisited action (Q = N = 0)
#if IPP_MCTS_PUCT_TABLE
        const float2 pt = a.puct[Ns];  // {c_init + log((Ns + c_base + 1) / c_base), sqrt(Ns + 1)}: a visit count is at most M
        const float prior_c = pt.x, sq = pt.y;
#else
        const float prior_c = d.c_init + logf(((float)Ns + d.c_base + 1.0f) / d.c_base);
        const float sq = sqrtf((float)Ns + 1.0f);
#endif
        const bool force = depth == 0;
        float best = -INFINITY;
        int best_s = 0x7fffffff, best_e = -1;
        int best_child = kNoNode, best_action = 0;  // of the lane's best edge: the winner's come along through shuffles, no second load
        auto consider = [&](int e, int slot, float prior, float q, int visits, int child_, int action_) {
            const float n = (float)visits;
            float u = (q - qmin) * qscale + prior_c * prior * (sq / (1.0f + n));
            if (force && n > 0.0f && n < ceilf(sqrtf(d.forced_k * prior * (float)Ns))) u = INFINITY;
            if (u > best || (u == best && slot < best_s)) {
                best = u;
                best_s = slot;
                best_e = e;
                best_child = child_;
                best_action = action_;
            }
        };
        if (cached) {
#pragma unroll
            for (int i = 0; i < kEdgeCache; ++i)
                if (ea[i].x == node)
                    consider(lane + 32 * i, ea[i].y, __int_as_float(ea[i].z), __int_as_float(ea[i].w), en[i], edges[lane + 32 * i].child,
                             edges[lane + 32 * i].pad[0]);
        } else {
            for (int e = lane; e < ne; e += 32) {
                const Edge ed = pool[e];
                if (ed.parent != node) continue;
                consider(e, ed.slot, ed.prior, ed.q, ed.n, ed.child, ed.pad[0]);
            }
        }
        const int2 bu = a.bu[(size_t)t * d.M + node];
        if (lane == 0 && bu.x >= 0) {
            const float u = (0.0f - qmin) * qscale + prior_c * __int_as_float(bu.y) * sq;
            if (u > best || (u == best && bu.x < best_s)) {
                best = u;
                best_s = bu.x;
                best_e = -1;
            }
        }
        {
            const int win = warp_argmax_lowest_slot(best, best_s, best, best_s);
            best_e = __shfl_sync(0xffffffffu, best_e, win);
            best_child = __shfl_sync(0xffffffffu, best_child, win);
            best_action = __shfl_sync(0xffffffffu, best_action, win);
        }
        int child = kNoNode;
        if (best_e < 0) {
            // a new edge: move the slot's prior into the pool and find the node's next best unvisited action
            best_e = ne;
            if (lane == 0) {
                Edge ed;
                ed.parent = node, ed.slot = best_s, ed.prior = __int_as_float(bu.y), ed.q = 0.0f, ed.n = 0, ed.child = kNoNode;
                ed.pad[0] = ed.pad[1] = 0;
                edges[ne] = ed;
                a.n_edges[t] = ne + 1;
                P[best_s] = -2.0f;
            }
            __syncwarp();
            float bp = -1.0f;
            int bs = 0x7fffffff;
            best_prior_scan(P, d.W, lane, bp, bs);
            warp_argmax_lowest_slot(bp, bs, bp, bs);
            if (lane == 0) {
                const float sc = a.pscale[(size_t)t * d.M + node];  // rows hold unnormalised priors
                a.bu[(size_t)t * d.M + node] = make_int2(bp >= 0.0f ? bs : -1, __float_as_int(bp >= 0.0f ? (sc > 0.0f ? bp * sc : -sc) : bp));
            }
        } else {
            child = best_child;
        }
        if (child != kNoNode) {
            // a visited edge whose child exists: its action id was cached when the child was created, the child's header has the
            // rest — no slot decode, no poses, no cost on the way down
            if (lane == 0) {
                a.path_edge[(size_t)t * d.max_path + len] = best_e;
                a.path_action[(size_t)t * d.max_path + len] = best_action;
            }
            ++len;
            node = child;
            continue;
        }
        // the chosen edge
        const Slot c = slot_cell(d, p, best_s, ccol, crow);
        double nx, ny, nh, ax, ay, ah;
        node_pose(p, a, t, node, ccol, crow, lvl, nx, ny, nh);
        slot_pose(p, c, ax, ay, ah);
        const float cost = job_cost(p, ax, ay, ah, nx, ny, nh);
        const float child_budget = budget - cost;  // mcts.py:247
        if (lane == 0) {
            a.path_edge[(size_t)t * d.max_path + len] = best_e;
            a.path_action[(size_t)t * d.max_path + len] = c.lvl * (p.X * p.Y) + p.X * c.col + c.row;
        }
        ++len;
        {
            const int n_nodes = a.n_nodes[t];
            if (depth + 1 > d.H || !(child_budget > 0.0f) || n_nodes >= d.M) {
                kind = IPP_MCTS_LEAF_TERMINAL;  // simulate() returns 0 before any node exists (mcts.py:175-176)
                depth += 1;
                budget = child_budget;
                ccol = c.col, crow = c.row, lvl = c.lvl;
                node = -1;
                break;
            }
            if (lane == 0) {
                hdr[n_nodes] = make_int4(pack_pos(c.col, c.row, c.lvl), __float_as_int(child_budget), 0, depth + 1);
                a.bu[(size_t)t * d.M + n_nodes] = make_int2(-1, 0);
                edges[best_e].child = n_nodes;
                edges[best_e].pad[0] = c.lvl * (p.X * p.Y) + p.X * c.col + c.row;
                a.n_nodes[t] = n_nodes + 1;
            }
            __syncwarp();
            node = n_nodes;
            depth += 1;
            budget = child_budget;
            ccol = c.col, crow = c.row, lvl = c.lvl;
            kind = IPP_MCTS_LEAF_EVAL;
            break;
        }
    }
    if (lane == 0) {
        for (int k = len; k < d.max_path; ++k) a.path_action[(size_t)t * d.max_path + k] = -1;
        int *lf = a.leaf + (size_t)t * IPP_MCTS_LEAF_WORDS;
        lf[0] = kind;
        lf[1] = node;
        lf[2] = ccol;
        lf[3] = crow;
        lf[4] = lvl;
        lf[5] = depth;
        lf[6] = __float_as_int(budget);
        lf[7] = len;
    }
    __syncwarp();  // the path and the leaf record written by lane 0 are read by every lane below
    if (flags & 0x80000000u) return;  // timing probe (IPP_MCTS_SKIP_ROLLOUT=1): the descent alone, rewards left stale
    mcts_rollout_body<LAYOUT>(p, d, a, flags, t, lane, s_rect[threadIdx.x >> 5], s_node[threadIdx.x >> 5]);
}

// ------------------------------------------------------------------------------------------------
// rollout: the reward of the path's newest prediction step
// ------------------------------------------------------------------------------------------------
// The reference replays the whole path every simulation: one simulate_prediction_step per level on a copy of the covariance
// (planning/mcts_zero/mcts.py:239-246 -> planning/common/optimization.py:14-30).  A prediction step is deterministic given
// the path above it, so the search memoises it: an edge keeps the reward of its step (first rollout), a node the variances
// its step left on its footprint (<= `tile` floats in HBM).  A simulation then costs ONE footprint — the new edge's: its
// variances come from the latest ancestor overlay that covers a cell, else from the env's belief (which must not change
// between ipp_mcts_begin and the search's last simulation) — instead of one footprint per level; paths that end on an edge
// visited before (terminal edges) cost nothing.  Same arithmetic as ipp_rollout_kernel (rollout_kernel.cuh): bit-identical
// rewards, hence identical trees.
template <int LAYOUT>
__device__ __forceinline__ void mcts_rollout_body(const StepParams &p, const TreeDims &d, const TreeArrays &a, uint32_t flags, int t, int lane,
                                                  int4 *s_rect_w /* [IPP_MCTS_MAX_PATH] */, int *s_node_w /* [IPP_MCTS_MAX_PATH] */) {
    const int *lf = a.leaf + (size_t)t * IPP_MCTS_LEAF_WORDS;
    const int len = lf[7];
    if (len == 0) return;
    const int *pe = a.path_edge + (size_t)t * d.max_path;
    const int *pa = a.path_action + (size_t)t * d.max_path;
    float *pr = a.path_reward + (size_t)t * d.max_path;
    Edge *edges = a.edges + (size_t)t * d.E;

    // steps above the new one: cached rewards; rectangles and nodes of the path for the overlay look-up
    bool last_cached = false;
    if (lane < len) {
        const Edge ed = edges[pe[lane]];
        // the last step is new when its edge has never been backed up; a path that ends on an unexpanded node is rolled out
        // again (same values): the node's overlay must exist whatever happened before
        bool cached = lane < len - 1 || ed.n > 0;
        if (lane == len - 1 && lf[0] == IPP_MCTS_LEAF_EVAL && lf[1] > 0 && lf[1] == ed.child) cached = false;
        if (cached) pr[lane] = __int_as_float(ed.pad[1]);
        if (lane == len - 1) last_cached = cached;
        if (lane < len - 1) {
            int lvl, col, row;
            decode_id(p, pa[lane], lvl, col, row);
            Geom g;
            clip_footprint(p, col, row, p.lut[lvl].rx, p.lut[lvl].ry, g);
            s_rect_w[lane] = make_int4(g.xl, g.yu, g.nx, g.ny);
            s_node_w[lane] = ed.child;
        }
    }
    last_cached = __shfl_sync(0xffffffffu, last_cached, len - 1);
    if (last_cached) return;
    __syncwarp();

    const int k = len - 1;  // the new step
    const int env = d.first_env + t;
    const Belief<LAYOUT> bel(p, (size_t)env);
    const bool adaptive = (flags & IPP_FLAG_ADAPTIVE) != 0;
    const bool entropy = (flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;
    int lvl, col, row;
    decode_id(p, pa[k], lvl, col, row);
    const Geom g = geom_from_cell(p, lvl, col, row);
    FuseCtx fc;
    fc.rf = g.rf;
    fc.R = g.R;
    fc.invR = fast_rcp(g.R);
    const int nqx = (g.nx + 1) >> 1, nqy = (g.ny + 1) >> 1, nq = nqx * nqy;
    const float inv_nqx = __frcp_rn((float)nqx);
    // the node this step creates keeps its variances for its descendants (none when the search stopped at a terminal edge)
    const int new_node = lf[0] == IPP_MCTS_LEAF_EVAL ? lf[1] : -1;
    float *mine = new_node > 0 ? a.overlay + ((size_t)t * d.M + new_node) * a.tile : nullptr;
    const float *ov = a.overlay + (size_t)t * d.M * a.tile;

    float acc = 0.0f;
    for (int q = lane; q < nq; q += 32) {
        const int qyy = fdiv(q, nqx, inv_nqx), qxx = q - qyy * nqx;
        const int r0 = 2 * qyy, c0 = 2 * qxx;
        const bool cok = c0 + 1 < g.nx, rok = r0 + 1 < g.ny;
        const bool ok[4] = {true, cok, rok, cok && rok};
        float m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
        const int R0 = g.yu + r0, C0 = g.xl + c0;
        // the quad's variance: the latest path node whose footprint covers it, else the env's belief; a quad that straddles a
        // rectangle border goes cell by cell
        int src = -1;  // >= 0: overlay of that path node; -1: belief; -2: mixed
        int4 rc = make_int4(0, 0, 0, 0);
        for (int s = k - 1; s >= 0; --s) {
            rc = s_rect_w[s];
            const int dx0 = C0 - rc.x, dy0 = R0 - rc.y, dx1 = dx0 + (cok ? 1 : 0), dy1 = dy0 + (rok ? 1 : 0);
            const bool in_x0 = (unsigned)dx0 < (unsigned)rc.z, in_x1 = (unsigned)dx1 < (unsigned)rc.z;
            const bool in_y0 = (unsigned)dy0 < (unsigned)rc.w, in_y1 = (unsigned)dy1 < (unsigned)rc.w;
            if (in_x0 && in_x1 && in_y0 && in_y1) {
                src = s;
                break;
            }
            if ((in_x0 || in_x1) && (in_y0 || in_y1)) {
                src = -2;
                break;
            }
        }
        if (src >= 0) {
            const float *tl = ov + (size_t)s_node_w[src] * a.tile + (R0 - rc.y) * rc.z + (C0 - rc.x);
            v[0] = __ldcg(tl);
            if (cok) v[1] = __ldcg(tl + 1);
            if (rok) v[2] = __ldcg(tl + rc.z);
            if (cok && rok) v[3] = __ldcg(tl + rc.z + 1);
            if (adaptive) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (ok[c]) m[c] = bel.load_mean(Belief<LAYOUT>::idx(p, R0 + (c >> 1), C0 + (c & 1)));
            }
        } else if (src == -1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (!ok[c]) continue;
                const int off = Belief<LAYOUT>::idx(p, R0 + (c >> 1), C0 + (c & 1));
                if (LAYOUT == IPP_LAYOUT_PLANES || LAYOUT == IPP_LAYOUT_SPLIT || !adaptive) {
                    v[c] = bel.load_var(off);
                    if (adaptive) m[c] = bel.load_mean(off);
                } else {
                    bel.load(off, m[c], v[c]);  // one 8-byte load for {mean, var}
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (!ok[c]) continue;
                const int R = R0 + (c >> 1), C = C0 + (c & 1);
                bool found = false;
                for (int s = k - 1; s >= 0 && !found; --s) {
                    const int4 r4 = s_rect_w[s];
                    const int dx = C - r4.x, dy = R - r4.y;
                    if ((unsigned)dx < (unsigned)r4.z && (unsigned)dy < (unsigned)r4.w) {
                        v[c] = __ldcg(ov + (size_t)s_node_w[s] * a.tile + dy * r4.z + dx);
                        found = true;
                    }
                }
                const int off = Belief<LAYOUT>::idx(p, R, C);
                if (!found) v[c] = bel.load_var(off);
                if (adaptive) m[c] = bel.load_mean(off);  // the mean never changes in a prediction step
            }
        }
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        float mn[4], vn[4];
        bool msk[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) msk[c] = ok[c] && (!adaptive || (fmaf(p.kappa, v[c], m[c]) >= p.thr));
        acc += kalman_quad_rt(entropy, adaptive, fc, cok, rok, m, v, z, msk, mn, vn);
        if (mine) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (ok[c]) mine[(r0 + (c >> 1)) * g.nx + c0 + (c & 1)] = vn[c];
        }
    }
    float accd = acc;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, s);
    if (lane == 0) {
        // cost from the pose the step starts at: the root's (any pose) or the lattice cell of the step above
        double qx, qy, qh;
        if (k == 0) {
            qx = a.root_pose[3 * t], qy = a.root_pose[3 * t + 1], qh = a.root_pose[3 * t + 2];
        } else {
            int l2, c2, r2;
            decode_id(p, pa[k - 1], l2, c2, r2);
            qx = __dadd_rn(__dmul_rn(p.res, (double)c2), __dmul_rn(0.5, p.res));
            qy = __dadd_rn(__dmul_rn(p.res, (double)r2), __dmul_rn(0.5, p.res));
            qh = p.lut[l2].alt;
        }
        const float cost = job_cost(p, g.px, g.py, g.ph, qx, qy, qh);
        const float reward = accd * fast_rcp(cost + 1.0f);
        pr[k] = reward;
        edges[pe[k]].pad[1] = __float_as_int(reward);
    }
}

// ------------------------------------------------------------------------------------------------
// expansion + backup
// ------------------------------------------------------------------------------------------------
#ifndef IPP_MCTS_EXPAND_MINBLOCKS
#define IPP_MCTS_EXPAND_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(kTreeWarps * 32, IPP_MCTS_EXPAND_MINBLOCKS) mcts_expand_kernel(const __grid_constant__ StepParams p, TreeDims d, TreeArrays a,
                                                                      const float *priors_window, const float *priors_dense,
                                                                      const float *values, const float *root_noise, int num_actions) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * kTreeWarps + (threadIdx.x >> 5);
    if (t >= d.T) return;
    const int *lf = a.leaf + (size_t)t * IPP_MCTS_LEAF_WORDS;
    const int kind = lf[0], node = lf[1], len = lf[7];
    int4 *hdr = a.hdr + (size_t)t * d.M;
    float value = 0.0f;
    if (kind == IPP_MCTS_LEAF_EVAL) {
        int ccol, crow, lvl;
        const int4 h = hdr[node];
        unpack_pos(h.x, ccol, crow, lvl);
        const float budget = __int_as_float(h.y);
        double nx, ny, nh;
        node_pose(p, a, t, node, ccol, crow, lvl, nx, ny, nh);
        float *P = a.P + ((size_t)t * d.M + node) * d.Wp;
        const bool noisy = node == 0 && root_noise != nullptr && d.dir_eps > 0.0f;
        // get_next_actions_mask (mcts.py:148-158) and policy * mask (mcts.py:220).  Slot s = (lvl * D + a) * D + b is the cell
        // (ccol + a - r, crow + b - r) at altitude level lvl: lanes run over b, the warp over (lvl, a), so the column part of
        // the pose / distance / in-grid test is warp-uniform, the row part a per-lane constant, and no slot index is ever
        // divided (the s = lane, lane + 32, ... sweep spent ~115 instructions per slot on two integer divisions and three
        // fp64 poses; this one ~15).  Same arithmetic as slot_pose() / job_dist(): bit-identical priors.
        // Every reference call site leaves uav_specificaion = None: the mask compares the Euclidean distance with the budget
        // even when costs are flight times (quirk, reproduced; the budget is decremented by the cost).
        float sum = 0.0f;
        int n_valid = 0;
        float bp = -1.0f;  // best (largest prior, lowest slot) action of this lane's slots: the node's first unvisited candidate
        int bs = 0x7fffffff;
        const double half_res = __dmul_rn(0.5, p.res);
        const int N = p.X * p.Y;
        if (node != 0 && a.tables_ok) {
            // a lattice node: geometry from the per-level slot tables (22 KB, cache resident), lanes over consecutive slots
            const float4 *dist_row = reinterpret_cast<const float4 *>(a.slot_dist + (size_t)lvl * d.Wp);
            const int4 *info_row = reinterpret_cast<const int4 *>(a.slot_info);
            // a lane takes four consecutive slots (one 16-byte group of the tables and of the node's row), two groups per pass with
            // every load issued before the first use (the evaluator's prior row streams from HBM at any 4-byte alignment: scalar
            // loads; invalid slots are read too, then masked)
            const float *prow = priors_window ? priors_window + (size_t)t * d.W : nullptr;
            const int nv = d.Wp >> 2;
            float4 *P4 = reinterpret_cast<float4 *>(P);
            for (int v0 = 0; v0 < nv; v0 += 64) {
                int4 inf[2];
                float4 dst[2];
                float prv[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int v = min(v0 + lane + 32 * j, nv - 1);
                    inf[j] = __ldg(info_row + v);
                    dst[j] = __ldg(dist_row + v);
#pragma unroll
                    for (int c = 0; c < 4; ++c) prv[j][c] = prow ? __ldg(prow + min(4 * v + c, d.W - 1)) : 1.0f;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int v = v0 + lane + 32 * j;
                    if (v >= nv) continue;
                    const int infs[4] = {inf[j].x, inf[j].y, inf[j].z, inf[j].w};
                    const float dsts[4] = {dst[j].x, dst[j].y, dst[j].z, dst[j].w};
                    float out[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int sl = 4 * v + c;
                        const int l = infs[c] & 255, col = ccol + ((infs[c] >> 8) & 255) - d.r, row = crow + (infs[c] >> 16) - d.r;
                        const float dist = dsts[c];  // 0 in the padding of the last group
                        float pr = -1.0f;
                        if (col >= 0 && col < p.X && row >= 0 && row < p.Y && dist > 0.0f && dist <= budget && dist < d.max_dist) {
                            pr = (!prow && priors_dense) ? priors_dense[(size_t)t * num_actions + l * N + p.X * col + row] : prv[j][c];
                            pr = fmaxf(pr, 0.0f);
                            sum += pr;
                            ++n_valid;
                            if (pr > bp || (pr == bp && sl < bs)) {
                                bp = pr;
                                bs = sl;
                            }
                        }
                        out[c] = pr;
                    }
                    P4[v] = make_float4(out[0], out[1], out[2], out[3]);
                }
            }
        } else
        for (int b0 = 0; b0 < d.D; b0 += 32) {
            const int b = b0 + lane;
            const int row = crow + b - d.r;
            const bool row_ok = b < d.D && row >= 0 && row < p.Y;
            const float dy = (float)(__dadd_rn(__dmul_rn(p.res, (double)row), half_res) - ny);
            for (int l = 0; l < d.L; ++l) {
                const float dz = (float)(p.lut[l].alt - nh);
                const float dyz = fmaf(dy, dy, dz * dz);
                for (int a_ = 0; a_ < d.D; ++a_) {
                    const int col = ccol + a_ - d.r;
                    const bool col_ok = col >= 0 && col < p.X;
                    const float dx = (float)(__dadd_rn(__dmul_rn(p.res, (double)col), half_res) - nx);
                    const float dist = fast_sqrt(fmaf(dx, dx, dyz));
                    const int sl = (l * d.D + a_) * d.D + b;
                    float pr = -1.0f;
                    if (row_ok && col_ok && dist > 0.0f && dist <= budget && dist < d.max_dist) {
                        pr = priors_window ? priors_window[(size_t)t * d.W + sl]
                                           : (priors_dense ? priors_dense[(size_t)t * num_actions + l * N + p.X * col + row] : 1.0f);
                        pr = fmaxf(pr, 0.0f);
                        if (noisy) pr = (1.0f - d.dir_eps) * pr;
                        sum += pr;
                        if (noisy) pr += d.dir_eps * root_noise[(size_t)t * d.W + sl];
                        ++n_valid;
                        if (pr > bp || (pr == bp && sl < bs)) {
                            bp = pr;
                            bs = sl;
                        }
                    }
                    if (b < d.D) P[sl] = pr;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
        }
        __syncwarp();
        if (n_valid > 0) {  // else: simulate() returns 0 and the node stays a leaf (mcts.py:197-199)
            // Ps /= sum(Ps) (mcts.py:228-234).  With exploration noise the sum runs over ALL actions, invalid ones
            // included (noise is added after masking, mcts.py:160-164,225-226): (1-eps) * sum_valid(p) + eps * 1.
            // The row keeps the masked priors as they are; the node keeps 1 / sum (or, for a row that sums to 0, the uniform
            // prior of its valid actions), and every reader scales: no second pass over the row.  The arg max is taken on the
            // unscaled values (scaling is monotone; the reference compares in fp64, where distinct priors stay distinct).
            const float total = noisy ? sum + d.dir_eps : sum;
            const float scale = total > 0.0f ? 1.0f / total : 0.0f;
            const float uniform = 1.0f / (float)n_valid;
            warp_argmax_lowest_slot(bp, bs, bp, bs);
            if (lane == 0) {
                hdr[node] = make_int4(h.x, h.y, 0, (h.w & 0xFF) | (1 << 8));
                a.pscale[(size_t)t * d.M + node] = total > 0.0f ? scale : -uniform;
                a.bu[(size_t)t * d.M + node] = make_int2(bs, __float_as_int(total > 0.0f ? bp * scale : uniform));
            }
            value = values ? values[t] : 0.0f;
        }
    }
    // backup (mcts.py:248-265), deepest edge first.  Lane k owns edge k of the path: the loads (reward, edge, parent header) of all
    // levels are in flight at once, the discounted sum runs down the path through shuffles, every lane writes its own edge back
    // (the edges of a path have distinct parents).
    float r_k = 0.0f, q_k = 0.0f;
    int n_k = 0, z_k = 0, parent_k = 0;
    Edge *ed = nullptr;
    if (lane < len) {
        r_k = a.path_reward[(size_t)t * d.max_path + lane];
        ed = a.edges + (size_t)t * d.E + a.path_edge[(size_t)t * d.max_path + lane];
        n_k = ed->n;
        q_k = ed->q;
        parent_k = ed->parent;
        z_k = hdr[parent_k].z;
    }
    float mine = 0.0f;
    for (int k = len - 1; k >= 0; --k) {  // warp-uniform
        value = __shfl_sync(0xffffffffu, r_k, k) + d.gamma * value;
        if (lane == k) mine = value;
    }
    if (lane < len) {
        ed->q = n_k > 0 ? ((float)n_k * q_k + mine) / (float)(n_k + 1) : mine;
        ed->n = n_k + 1;
        hdr[parent_k].z = z_k + 1;
    }
}

__global__ void mcts_begin_kernel(TreeDims d, TreeArrays a, const StepParams p, const float *budgets) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    const double x = a.root_pose[3 * t], y = a.root_pose[3 * t + 1];
    const int col = clampi((int)floor(__ddiv_rn(x, p.res)), 0, p.X - 1), row = clampi((int)floor(__ddiv_rn(y, p.res)), 0, p.Y - 1);
    a.hdr[(size_t)t * d.M] = make_int4(pack_pos(col, row, -1), __float_as_int(budgets[t]), 0, 0);
    a.bu[(size_t)t * d.M] = make_int2(-1, 0);
    a.n_nodes[t] = 1;
    a.n_edges[t] = 0;
}

// dense root statistics from the prior array and the root's edges: one CTA per tree
__global__ void mcts_root_export_kernel(const __grid_constant__ StepParams p, TreeDims d, TreeArrays a, float *ps, float *qsa, int *nsa, int *ids,
                                        int *ns) {
    const int t = blockIdx.x;
    const int4 h = a.hdr[(size_t)t * d.M];
    int ccol, crow, lvl;
    unpack_pos(h.x, ccol, crow, lvl);
    const bool expanded = ((h.w >> 8) & 1) != 0;
    const float sc = a.pscale[(size_t)t * d.M];
    const size_t base = (size_t)t * d.M * d.Wp;
    for (int s = threadIdx.x; s < d.W; s += blockDim.x) {
        const Slot c = slot_cell(d, p, s, ccol, crow);
        const size_t o = (size_t)t * d.W + s;
        if (ps) {
            float pv = expanded ? a.P[base + s] : -1.0f;
            if (pv >= 0.0f) pv = sc > 0.0f ? pv * sc : -sc;  // rows hold unnormalised priors
            ps[o] = pv;
        }
        if (qsa) qsa[o] = 0.0f;
        if (nsa) nsa[o] = 0;
        if (ids) ids[o] = c.in_grid ? c.lvl * (p.X * p.Y) + p.X * c.col + c.row : -1;
    }
    __syncthreads();
    const Edge *edges = a.edges + (size_t)t * d.E;
    const int ne = a.n_edges[t];
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const Edge ed = edges[e];
        if (ed.parent != 0) continue;
        const size_t o = (size_t)t * d.W + ed.slot;
        if (ps) ps[o] = ed.prior;
        if (qsa) qsa[o] = ed.q;
        if (nsa) nsa[o] = ed.n;
    }
    if (ns && threadIdx.x == 0) ns[t] = h.z;
}

// Window geometry of a lattice node: slot -> (level, column, row) indices and its distance to the node, per node level.  Same
// expressions as the expansion's general path (slot_pose / job_dist), evaluated at offsets: identical bits whenever res * cell
// + res / 2 is exact in fp64 (resolutions with a float32 mantissa; checked by the host), so that pose differences depend on
// the cell offsets only.
__global__ void slot_table_kernel(TreeDims d, const StepParams p, int *info, float *dist) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.Wp) return;
    if (s >= d.W) {  // padding of the last 16-byte group: distance 0 = never a valid action
        info[s] = 0;
        for (int ln = 0; ln < d.L; ++ln) dist[(size_t)ln * d.Wp + s] = 0.0f;
        return;
    }
    const int DD = d.D * d.D;
    const int l = s / DD, rem = s - l * DD, a_ = rem / d.D, b = rem - a_ * d.D;
    info[s] = l | (a_ << 8) | (b << 16);
    const float dx = (float)__dmul_rn(p.res, (double)(a_ - d.r));
    const float dy = (float)__dmul_rn(p.res, (double)(b - d.r));
    for (int ln = 0; ln < d.L; ++ln) {
        const float dz = (float)(p.lut[l].alt - p.lut[ln].alt);
        const float dyz = fmaf(dy, dy, dz * dz);
        dist[(size_t)ln * d.Wp + s] = fast_sqrt(fmaf(dx, dx, dyz));
    }
}

// the visit-count dependent factors of compute_uct (mcts.py:280-296), tabulated once per search object
__global__ void puct_table_kernel(float2 *tab, int n, float c_init, float c_base) {
    const int Ns = blockIdx.x * blockDim.x + threadIdx.x;
    if (Ns < n) tab[Ns] = make_float2(c_init + logf(((float)Ns + c_base + 1.0f) / c_base), sqrtf((float)Ns + 1.0f));
}

__global__ void iota_kernel(int *out, int n, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
struct ipp_mcts {
    ipp_engine *env = nullptr;
    ipp_mcts_config cfg{};
    TreeDims d{};
    TreeArrays a{};
    StepParams sp{};
    int num_actions = 0;
    int *d_env_index = nullptr;
    float *d_budgets = nullptr;
    // staging for host-side evaluator outputs / exports
    float *d_in_priors = nullptr, *d_in_values = nullptr, *d_in_noise = nullptr;
    size_t cap_in_priors = 0;
    float *d_out_f[2] = {nullptr, nullptr};
    int *d_out_i[3] = {nullptr, nullptr, nullptr};
    std::vector<void *> owned;
    cudaStream_t stream = nullptr;
    uint64_t device_bytes = 0, launches = 0;
    int simulations = 0;
    bool begun = false, pending = false;
    std::string err;
};

static thread_local std::string g_mcts_create_err;

static int mfail(ipp_mcts *m, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m)
        m->err = buf;
    else
        g_mcts_create_err = buf;
    return code;
}

#define MCU(m, call)                                                                                              \
    do {                                                                                                          \
        cudaError_t _s = (call);                                                                                  \
        if (_s != cudaSuccess)                                                                                    \
            return mfail((m), _s == cudaErrorMemoryAllocation ? IPP_ERR_NOMEM : IPP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                         cudaGetErrorString(_s), __FILE__, __LINE__);                                             \
    } while (0)

template <typename T>
static int malloc_dev(ipp_mcts *m, T **p, size_t n) {
    MCU(m, cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    m->owned.push_back((void *)*p);
    m->device_bytes += n * sizeof(T);
    return IPP_OK;
}

extern "C" int ipp_mcts_create(ipp_engine *env, const ipp_mcts_config *cfg, ipp_mcts **out) {
    if (!env || !cfg || !out) return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: NULL argument");
    *out = nullptr;
    if (cfg->struct_bytes != sizeof(ipp_mcts_config)) return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: ABI mismatch");
    ipp_info info;
    if (ipp_get_info(env, &info) != IPP_OK) return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: bad engine");
    if (cfg->n_trees < 1 || cfg->first_env < 0 || cfg->first_env + cfg->n_trees > info.batch)
        return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: trees [%d, %d) outside the engine batch %d", cfg->first_env,
                     cfg->first_env + cfg->n_trees, info.batch);
    if (cfg->num_simulations < 1 || cfg->num_simulations > 65000) return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: num_simulations outside [1, 65000]");
    if (cfg->episode_horizon < 0 || cfg->episode_horizon + 1 > IPP_MCTS_MAX_PATH)
        return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: episode_horizon outside [0, %d]", IPP_MCTS_MAX_PATH - 1);
    if (!(cfg->max_valid_action_distance > 0) || !(cfg->puct_base > 0))
        return mfail(nullptr, IPP_ERR_INVALID, "ipp_mcts_create: max_valid_action_distance and puct_base must be > 0");
    if (info.x_dim > 4095 || info.y_dim > 4095) return mfail(nullptr, IPP_ERR_UNSUPPORTED, "ipp_mcts_create: grids beyond 4095 cells per side");
    if (info.x_dim != info.y_dim)  // level * N + x_dim * col + row is a bijection on square grids only (the dense prior is indexed by it)
        return mfail(nullptr, IPP_ERR_UNSUPPORTED, "ipp_mcts_create: the reference's action ids are ambiguous on non-square grids (%d x %d)", info.x_dim, info.y_dim);

    ipp_mcts *m = new (std::nothrow) ipp_mcts();
    if (!m) return mfail(nullptr, IPP_ERR_NOMEM, "ipp_mcts_create: out of host memory");
    m->env = env;
    m->cfg = *cfg;
    ipp_internal_step_params(env, &m->sp);
    m->stream = ipp_internal_stream(env);
    m->num_actions = info.num_actions;
    TreeDims &d = m->d;
    d.T = cfg->n_trees;
    d.M = cfg->num_simulations + 1;
    d.E = cfg->num_simulations;
    d.L = info.num_altitude_levels;
    d.r = (int)std::floor(cfg->max_valid_action_distance / m->sp.res) + 1;
    d.D = 2 * d.r + 1;
    d.W = d.L * d.D * d.D;
    d.Wp = (d.W + 3) & ~3;
    d.H = cfg->episode_horizon;
    d.max_path = d.H + 1;
    d.first_env = cfg->first_env;
    d.c_init = (float)cfg->puct_init;
    d.c_base = (float)cfg->puct_base;
    d.gamma = (float)cfg->gamma;
    d.forced_k = (float)cfg->forced_playout_factor;
    d.max_dist = (float)cfg->max_valid_action_distance;
    d.dir_eps = (float)cfg->dirichlet_eps;

    auto bail = [&](int rc) {
        g_mcts_create_err = m->err;
        ipp_mcts_destroy(m);
        return rc;
    };
    const size_t TM = (size_t)d.T * d.M, TMW = TM * d.Wp, TP = (size_t)d.T * d.max_path;
    int rc;
    TreeArrays &a = m->a;
    if ((rc = malloc_dev(m, &a.hdr, TM)) || (rc = malloc_dev(m, &a.P, TMW)) || (rc = malloc_dev(m, &a.pscale, TM)) || (rc = malloc_dev(m, &a.bu, TM)) ||
        (rc = malloc_dev(m, &a.edges, (size_t)d.T * d.E)) || (rc = malloc_dev(m, &a.n_edges, (size_t)d.T)) ||
        (rc = malloc_dev(m, &a.n_nodes, (size_t)d.T)) || (rc = malloc_dev(m, &a.root_pose, 3 * (size_t)d.T)) ||
        (rc = malloc_dev(m, &a.path_edge, TP)) || (rc = malloc_dev(m, &a.path_action, TP)) ||
        (rc = malloc_dev(m, &a.path_reward, TP)) || (rc = malloc_dev(m, &a.leaf, (size_t)d.T * IPP_MCTS_LEAF_WORDS)) ||
        (rc = malloc_dev(m, &m->d_env_index, (size_t)d.T)) || (rc = malloc_dev(m, &m->d_budgets, (size_t)d.T)))
        return bail(rc);
    a.tile = 1;
    for (int k = 0; k < info.num_altitude_levels; ++k)
        a.tile = std::max(a.tile, std::min(2 * info.radius_x[k] + 1, info.x_dim) * std::min(2 * info.radius_y[k] + 1, info.y_dim));
    if ((rc = malloc_dev(m, &a.overlay, TM * (size_t)a.tile)) || (rc = malloc_dev(m, &a.puct, (size_t)d.M + 1))) return bail(rc);
    puct_table_kernel<<<(d.M + 256) / 256, 256, 0, m->stream>>>(a.puct, d.M + 1, d.c_init, d.c_base);
    m->launches++;
    {
        int *info_tab = nullptr;
        float *dist_tab = nullptr;
        if ((rc = malloc_dev(m, &info_tab, (size_t)d.Wp)) || (rc = malloc_dev(m, &dist_tab, (size_t)d.L * d.Wp))) return bail(rc);
        slot_table_kernel<<<(d.Wp + 255) / 256, 256, 0, m->stream>>>(d, m->sp, info_tab, dist_tab);
        m->launches++;
        a.slot_info = info_tab;
        a.slot_dist = dist_tab;
        // pose differences depend on cell offsets only when res * cell + res / 2 is exact in fp64: resolutions with a float32
        // mantissa (24 bits + 12 bits of cell index); window indices must fit the 8-bit fields
        a.tables_ok = (m->sp.res == (double)(float)m->sp.res) && d.D <= 255 && d.L <= 255 && getenv("IPP_MCTS_NO_TABLES") == nullptr;
    }
    iota_kernel<<<(d.T + 255) / 256, 256, 0, m->stream>>>(m->d_env_index, d.T, d.first_env);
    m->launches++;
    if (cudaStreamSynchronize(m->stream) != cudaSuccess) return bail(mfail(m, IPP_ERR_CUDA, "ipp_mcts_create: device initialisation failed"));
    *out = m;
    return IPP_OK;
}

extern "C" void ipp_mcts_destroy(ipp_mcts *m) {
    if (!m) return;
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (void *p : m->owned) cudaFree(p);
    delete m;
}

extern "C" const char *ipp_mcts_last_error(const ipp_mcts *m) { return m ? m->err.c_str() : g_mcts_create_err.c_str(); }

extern "C" int ipp_mcts_get_info(const ipp_mcts *m, ipp_mcts_info *out) {
    if (!m || !out) return IPP_ERR_INVALID;
    memset(out, 0, sizeof *out);
    out->n_trees = m->d.T;
    out->max_nodes = m->d.M;
    out->levels = m->d.L;
    out->window_dim = m->d.D;
    out->window_radius = m->d.r;
    out->window_slots = m->d.W;
    out->max_path = m->d.max_path;
    out->simulations = m->simulations;
    out->device_bytes = m->device_bytes;
    out->launches = m->launches;
    if (m->begun) {  // edges of all pools (a 4-byte counter per tree, read back on demand)
        std::vector<int> ne((size_t)m->d.T);
        if (cudaMemcpyAsync(ne.data(), m->a.n_edges, ne.size() * sizeof(int), cudaMemcpyDeviceToHost, m->stream) == cudaSuccess &&
            cudaStreamSynchronize(m->stream) == cudaSuccess)
            for (int v : ne) out->edges += (uint64_t)v;
        else
            cudaGetLastError();
    }
    return IPP_OK;
}

extern "C" int ipp_mcts_begin(ipp_mcts *m, const double *root_poses, const float *budgets) {
    if (!m || !budgets) return IPP_ERR_INVALID;
    const TreeDims &d = m->d;
    ipp_internal_step_params(m->env, &m->sp);  // belief pointers / flags of the engine as of now
    std::vector<double> all;
    if (!root_poses) {
        ipp_info info;
        ipp_get_info(m->env, &info);
        all.resize(3 * (size_t)info.batch);
        if (ipp_get_prev_pose(m->env, all.data()) != IPP_OK) return mfail(m, IPP_ERR_CUDA, "ipp_mcts_begin: %s", ipp_last_error(m->env));
        root_poses = all.data() + 3 * (size_t)d.first_env;
    }
    MCU(m, cudaMemcpyAsync(m->a.root_pose, root_poses, 3 * (size_t)d.T * sizeof(double), cudaMemcpyHostToDevice, m->stream));
    MCU(m, cudaMemcpyAsync(m->d_budgets, budgets, (size_t)d.T * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    mcts_begin_kernel<<<(d.T + 255) / 256, 256, 0, m->stream>>>(d, m->a, m->sp, m->d_budgets);
    m->launches++;
    MCU(m, cudaGetLastError());
    MCU(m, cudaStreamSynchronize(m->stream));  // the host buffers may go away
    m->simulations = 0;
    m->begun = true;
    m->pending = false;
    return IPP_OK;
}

extern "C" int ipp_mcts_simulate_begin(ipp_mcts *m, int32_t *leaf_info) {
    if (!m) return IPP_ERR_INVALID;
    if (!m->begun) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_simulate_begin: call ipp_mcts_begin first");
    if (m->pending) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_simulate_begin: the previous simulation has not been ended");
    if (m->simulations >= m->cfg.num_simulations) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_simulate_begin: node capacity (num_simulations) used up");
    const TreeDims &d = m->d;
    const int blocks = (d.T + kTreeWarps - 1) / kTreeWarps;
    // PUCT descent + the reward of the path's new prediction step (the steps above it are cached in their edges); the belief
    // is only read
    {
        uint32_t fl = m->cfg.step_flags & (IPP_REWARD_MASK | IPP_FLAG_ADAPTIVE);
        if (getenv("IPP_MCTS_SKIP_ROLLOUT")) fl |= 0x80000000u;  // timing probe only: 0.141 vs 0.193 ms per simulation on growing trees
        const int layout = ipp_internal_layout(m->env);
        void (*kern)(const StepParams, TreeDims, TreeArrays, uint32_t, int) =
            layout == IPP_LAYOUT_TILED   ? mcts_select_kernel<IPP_LAYOUT_TILED>
            : layout == IPP_LAYOUT_SUPER ? mcts_select_kernel<IPP_LAYOUT_SUPER>
            : layout == IPP_LAYOUT_SPLIT ? mcts_select_kernel<IPP_LAYOUT_SPLIT>
            : layout == IPP_LAYOUT_MV    ? mcts_select_kernel<IPP_LAYOUT_MV>
                                         : mcts_select_kernel<IPP_LAYOUT_PLANES>;
        // experiment switch (IPP_MCTS_STAGE_EDGES=1): the trees' edge pools staged in shared memory once per simulation (they fit
        // beside 8 resident CTAs per SM up to 192 edges per tree) — measured 1.5 % slower on growing trees (0.1940 vs 0.1907 ms per
        // simulation): the scans' loads were not what the descent waits for
        const size_t pool_bytes = (size_t)kTreeWarps * d.E * sizeof(Edge);
        const char *st_env = getenv("IPP_MCTS_STAGE_EDGES");
        const int stage = pool_bytes <= 24 * 1024 && st_env != nullptr && st_env[0] == '1';
        kern<<<blocks, kTreeWarps * 32, stage ? pool_bytes : 0, m->stream>>>(m->sp, d, m->a, fl, stage);
        m->launches++;
        MCU(m, cudaGetLastError());
    }
    if (leaf_info) {
        MCU(m, cudaMemcpyAsync(leaf_info, m->a.leaf, (size_t)d.T * IPP_MCTS_LEAF_WORDS * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        MCU(m, cudaStreamSynchronize(m->stream));
    }
    m->pending = true;
    return IPP_OK;
}

static int stage_in(ipp_mcts *m, float **slot, const float *src, size_t n, const float **dev) {
    *dev = nullptr;
    if (!src) return IPP_OK;
    if (!*slot) {
        int rc = malloc_dev(m, slot, n);
        if (rc != IPP_OK) return rc;
    }
    MCU(m, cudaMemcpyAsync(*slot, src, n * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    *dev = *slot;
    return IPP_OK;
}

extern "C" int ipp_mcts_simulate_end(ipp_mcts *m, const float *priors_window, const float *priors_dense, const float *values,
                                     const float *root_noise, int32_t inputs_are_device) {
    if (!m) return IPP_ERR_INVALID;
    if (!m->pending) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_simulate_end: no simulation in flight");
    if (priors_window && priors_dense) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_simulate_end: give window OR dense priors");
    const TreeDims &d = m->d;
    const float *pw = priors_window, *pd = priors_dense, *pv = values, *pn = root_noise;
    if (!inputs_are_device) {
        int rc;
        const size_t n_pri = priors_dense ? (size_t)d.T * m->num_actions : (size_t)d.T * d.W;
        if (priors_window || priors_dense) {
            if (m->cap_in_priors < n_pri) {  // (re)allocate the priors staging for the larger of the two forms
                float *p = nullptr;
                if ((rc = malloc_dev(m, &p, n_pri)) != IPP_OK) return rc;
                m->d_in_priors = p;
                m->cap_in_priors = n_pri;
            }
            MCU(m, cudaMemcpyAsync(m->d_in_priors, priors_window ? priors_window : priors_dense, n_pri * sizeof(float), cudaMemcpyHostToDevice, m->stream));
            pw = priors_window ? m->d_in_priors : nullptr;
            pd = priors_dense ? m->d_in_priors : nullptr;
        }
        if ((rc = stage_in(m, &m->d_in_values, values, (size_t)d.T, &pv)) != IPP_OK) return rc;
        if ((rc = stage_in(m, &m->d_in_noise, root_noise, (size_t)d.T * d.W, &pn)) != IPP_OK) return rc;
    }
    const int blocks = (d.T + kTreeWarps - 1) / kTreeWarps;
    mcts_expand_kernel<<<blocks, kTreeWarps * 32, 0, m->stream>>>(m->sp, d, m->a, pw, pd, pv, pn, m->num_actions);
    m->launches++;
    MCU(m, cudaGetLastError());
    if (!inputs_are_device) MCU(m, cudaStreamSynchronize(m->stream));  // host buffers may be reused by the caller
    m->pending = false;
    m->simulations++;
    return IPP_OK;
}

extern "C" int ipp_mcts_root_stats(ipp_mcts *m, float *ps, float *qsa, int32_t *nsa, int32_t *action_ids, int32_t *ns) {
    if (!m) return IPP_ERR_INVALID;
    if (!m->begun) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_root_stats: call ipp_mcts_begin first");
    const TreeDims &d = m->d;
    const size_t TW = (size_t)d.T * d.W;
    int rc;
    for (int k = 0; k < 2; ++k)
        if (!m->d_out_f[k] && (rc = malloc_dev(m, &m->d_out_f[k], TW)) != IPP_OK) return rc;
    for (int k = 0; k < 3; ++k)
        if (!m->d_out_i[k] && (rc = malloc_dev(m, &m->d_out_i[k], k == 2 ? (size_t)d.T : TW)) != IPP_OK) return rc;
    mcts_root_export_kernel<<<d.T, 256, 0, m->stream>>>(m->sp, d, m->a, m->d_out_f[0], m->d_out_f[1], m->d_out_i[0], m->d_out_i[1], m->d_out_i[2]);
    m->launches++;
    MCU(m, cudaGetLastError());
    if (ps) MCU(m, cudaMemcpyAsync(ps, m->d_out_f[0], TW * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    if (qsa) MCU(m, cudaMemcpyAsync(qsa, m->d_out_f[1], TW * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    if (nsa) MCU(m, cudaMemcpyAsync(nsa, m->d_out_i[0], TW * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    if (action_ids) MCU(m, cudaMemcpyAsync(action_ids, m->d_out_i[1], TW * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    if (ns) MCU(m, cudaMemcpyAsync(ns, m->d_out_i[2], (size_t)d.T * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    MCU(m, cudaStreamSynchronize(m->stream));
    return IPP_OK;
}

extern "C" int ipp_mcts_get_paths(ipp_mcts *m, int32_t *actions, float *rewards) {
    if (!m) return IPP_ERR_INVALID;
    if (!m->pending) return mfail(m, IPP_ERR_INVALID, "ipp_mcts_get_paths: no simulation in flight");
    const size_t TP = (size_t)m->d.T * m->d.max_path;
    if (actions) MCU(m, cudaMemcpyAsync(actions, m->a.path_action, TP * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    if (rewards) MCU(m, cudaMemcpyAsync(rewards, m->a.path_reward, TP * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    MCU(m, cudaStreamSynchronize(m->stream));
    return IPP_OK;
}

extern "C" void *ipp_mcts_device_ptr(ipp_mcts *m, int32_t which) {
    if (!m) return nullptr;
    switch (which) {
        case IPP_MCTS_PTR_LEAF_INFO: return m->a.leaf;
        case IPP_MCTS_PTR_PATH_ACTIONS: return m->a.path_action;
        case IPP_MCTS_PTR_PATH_REWARDS: return m->a.path_reward;
        default: return nullptr;
    }
}
