// rollout_kernel.cuh — path rollouts for tree search (sm_100a): one warp replays a whole action path
// of <= kMaxHorizon prediction steps from an env's CURRENT belief without writing a byte of it.
//
// Reference: MCTS.simulate descends the tree with one simulate_prediction_step per level
// (planning/mcts_zero/mcts.py:239-246 -> planning/common/optimization.py:14-30), each of which copies and
// updates the full covariance.  Here the variance a later step sees is the env's variance overlaid with
// the footprints the earlier steps of the same path produced; those (<= 529 cells each) live in shared
// memory, so a path costs 4 B per footprint cell of HBM reads (8 with the adaptive mask) and no writes,
// no scratch env, no undo.  Rewards are the reference's: sum_mask(v - v') / (cost + 1) per step
// (planning/common/rewards.py:15-31), cost between consecutive path poses (planning/common/actions.py:8-41).
//
//   grid = ceil(n_jobs / warps), block = warps * 32, dynamic smem = warps * (horizon - 1) * tile floats
#pragma once
#include "step_kernel.cuh"

namespace ipp {

constexpr int kMaxHorizon = 8;
constexpr int kRolloutWarps = 4;

struct RolloutParams {
    StepParams base;              // env_index, prev_in, prev_state, flags, belief pointers, LUT
    const int32_t *path_actions;  // [n_jobs][horizon] action ids, < 0 terminates the path
    float *rewards;               // [n_jobs][horizon]; entries past the end of a path are 0
    int horizon;
    int tile_floats;              // capacity of one overlay tile (largest footprint, cells)
};

template <int LAYOUT>
__global__ void __launch_bounds__(kRolloutWarps * 32) ipp_rollout_kernel(const __grid_constant__ RolloutParams rp) {
    extern __shared__ float s_tiles[];  // [warp][horizon - 1][tile_floats]
    __shared__ int4 s_rect[kRolloutWarps][kMaxHorizon];  // {xl, yu, nx, ny} of the earlier steps

    const StepParams &p = rp.base;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int job = blockIdx.x * kRolloutWarps + wib;
    if (job >= p.n_jobs) return;
    const int H = rp.horizon;
    float *tiles = s_tiles + (size_t)wib * (H - 1) * rp.tile_floats;
    const int env = p.env_index ? __ldg(p.env_index + job) : job;
    const Belief<LAYOUT> bel(p, (size_t)env);
    const bool adaptive = (p.flags & IPP_FLAG_ADAPTIVE) != 0;
    const bool entropy = (p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;

    const double *pv = p.prev_in != nullptr ? p.prev_in + 3 * (size_t)job : p.prev_state + 3 * (size_t)env;
    double qx = pv[0], qy = pv[1], qh = pv[2];  // pose the next step starts from

    int k = 0;
    for (; k < H; ++k) {
        const int id = __ldg(rp.path_actions + (size_t)job * H + k);
        if (id < 0) break;
        int lvl, col, row;
        decode_id(p, id, lvl, col, row);
        const Geom g = geom_from_cell(p, lvl, col, row);
        FuseCtx fc;
        fc.rf = g.rf;
        fc.R = g.R;
        fc.invR = fast_rcp(g.R);
        const int nqx = (g.nx + 1) >> 1, nqy = (g.ny + 1) >> 1, nq = nqx * nqy;
        const float inv_nqx = __frcp_rn((float)nqx);
        float *mine = tiles + (size_t)k * rp.tile_floats;  // written only when a later step may read it (k < H - 1)
        float acc = 0.0f;
        for (int q = lane; q < nq; q += 32) {
            const int qyy = fdiv(q, nqx, inv_nqx), qxx = q - qyy * nqx;
            const int r0 = 2 * qyy, c0 = 2 * qxx;
            const bool cok = c0 + 1 < g.nx, rok = r0 + 1 < g.ny;
            const bool ok[4] = {true, cok, rok, cok && rok};
            float m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
            const int R0 = g.yu + r0, C0 = g.xl + c0;
            // Where does the quad's variance come from?  The latest earlier step of this path whose footprint covers it, else the
            // env's belief in HBM.  Classify the whole quad against each earlier rectangle (one 16-byte shared load per step):
            // fully inside -> its overlay tile, disjoint from all -> HBM, straddling a border -> cell by cell.
            int src = -1;        // >= 0: overlay tile of that step; -1: HBM; -2: mixed
            int4 rc = make_int4(0, 0, 0, 0);
            for (int s = k - 1; s >= 0; --s) {
                rc = s_rect[wib][s];
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
                const float *t = tiles + (size_t)src * rp.tile_floats + (R0 - rc.y) * rc.z + (C0 - rc.x);
                v[0] = t[0];
                if (cok) v[1] = t[1];
                if (rok) v[2] = t[rc.z];
                if (cok && rok) v[3] = t[rc.z + 1];
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
                        const int4 r4 = s_rect[wib][s];
                        const int dx = C - r4.x, dy = R - r4.y;
                        if ((unsigned)dx < (unsigned)r4.z && (unsigned)dy < (unsigned)r4.w) {
                            v[c] = tiles[(size_t)s * rp.tile_floats + dy * r4.z + dx];
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
            if (k < H - 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (ok[c]) mine[(r0 + (c >> 1)) * g.nx + c0 + (c & 1)] = vn[c];
            }
        }
        float accd = acc;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, s);
        if (lane == 0) {
            const float cost = job_cost(p, g.px, g.py, g.ph, qx, qy, qh);
            rp.rewards[(size_t)job * H + k] = accd * fast_rcp(cost + 1.0f);
            s_rect[wib][k] = make_int4(g.xl, g.yu, g.nx, g.ny);
        }
        qx = g.px;
        qy = g.py;
        qh = g.ph;
        __syncwarp();  // tile k and its rectangle are visible to the next step
    }
    if (lane == 0)
        for (; k < H; ++k) rp.rewards[(size_t)job * H + k] = 0.0f;
}

}  // namespace ipp
