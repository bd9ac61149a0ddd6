// step_kernel.cuh — the general fused step kernel (sm_100a): one warp per job, footprint gathered
// straight from HBM with LSU loads.  Handles every mode and input form of the C ABI (action ids or
// fp64 poses, env_index jobs, predict-only, measure-only, caller-supplied measurements, log-odds).
// The throughput path for the headline configuration is the cp.async-staged persistent kernel in
// step_async.cuh; both share quad_math.cuh and produce bit-identical results.
//
// The footprint is tiled by 2x2-cell "quads" anchored at its top-left cell; lane l handles quads
// l, l+32, ...  A quad is exactly one measurement block at resolution factor 2 and four independent
// measurements at resolution factor 1, so the block sums of the Kalman update never leave a thread;
// the only cross-lane step is the final reward reduction (warp shuffle tree, fp64).
//
// HBM-bound gather-update-reduce: per covered cell the kernel reads gt, mean, var and writes
// mean, var (20 B), nothing else touches DRAM.  No tensor cores on purpose.
#pragma once
#include "quad_math.cuh"

namespace ipp {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kTapCap = 16;  // tap-table entries per axis staged in shared memory (same in both kernels)

// ---------------------------------------------------------------------------------------------
// belief accessors for the two HBM layouts
// ---------------------------------------------------------------------------------------------
// interleaved {mean,var}; IPP_LAYOUT_MV: row-major map, IPP_LAYOUT_TILED: 4x4-cell tiles, IPP_LAYOUT_SUPER: 192-byte
// super-tiles (quad_math.cuh)
template <int LAYOUT>
struct Belief {
    static_assert(LAYOUT == IPP_LAYOUT_MV || LAYOUT == IPP_LAYOUT_TILED || LAYOUT == IPP_LAYOUT_SUPER, "unknown layout");
    float2 *mv;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : mv(reinterpret_cast<float2 *>(p.mean) + env * p.plane) {}
    static __device__ __forceinline__ int idx(const StepParams &p, int R, int C) {
        return LAYOUT == IPP_LAYOUT_TILED ? tiled_mv_index(p.txm, R, C) : (LAYOUT == IPP_LAYOUT_SUPER ? super_mv_index(p.txm, R, C) : R * p.X + C);
    }
    static __device__ __forceinline__ int gidx(const StepParams &p, int R, int C) {
        return LAYOUT == IPP_LAYOUT_TILED ? tiled_gt_index(p.txg, R, C) : (LAYOUT == IPP_LAYOUT_SUPER ? super_gt_index(p.txm, R, C) : R * p.X + C);
    }
    __device__ __forceinline__ void load(int i, float &mean, float &var) const {
        const float2 t = __ldcg(mv + i);
        mean = t.x;
        var = t.y;
    }
    __device__ __forceinline__ float load_mean(int i) const { return __ldcg(&mv[i].x); }
    __device__ __forceinline__ float load_var(int i) const { return __ldcg(&mv[i].y); }
    __device__ __forceinline__ void store(int i, float mean, float var) const { mv[i] = make_float2(mean, var); }
    __device__ __forceinline__ void store_mean(int i, float mean) const { mv[i].x = mean; }
    __device__ __forceinline__ void store_var(int i, float var) const { mv[i].y = var; }
};

template <>
struct Belief<IPP_LAYOUT_PLANES> {
    float *m, *v;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : m(p.mean + env * p.plane), v(p.var + env * p.plane) {}
    static __device__ __forceinline__ int idx(const StepParams &p, int R, int C) { return R * p.X + C; }
    static __device__ __forceinline__ int gidx(const StepParams &p, int R, int C) { return R * p.X + C; }
    __device__ __forceinline__ void load(int i, float &mean, float &var) const {
        mean = __ldcg(m + i);
        var = __ldcg(v + i);
    }
    __device__ __forceinline__ float load_mean(int i) const { return __ldcg(m + i); }
    __device__ __forceinline__ float load_var(int i) const { return __ldcg(v + i); }
    __device__ __forceinline__ void store(int i, float mean, float var) const {
        m[i] = mean;
        v[i] = var;
    }
    __device__ __forceinline__ void store_mean(int i, float mean) const { m[i] = mean; }
    __device__ __forceinline__ void store_var(int i, float var) const { v[i] = var; }
};

// IPP_LAYOUT_SPLIT: var[tile][16] and {mean[16] | gt[16]}[tile] (quad_math.cuh); i = the compact tile index
template <>
struct Belief<IPP_LAYOUT_SPLIT> {
    float *m, *v;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : m(p.mean + env * p.plane_gt), v(p.var + env * p.plane) {}
    static __device__ __forceinline__ int idx(const StepParams &p, int R, int C) { return split_index(p.txm, R, C); }
    static __device__ __forceinline__ int gidx(const StepParams &p, int R, int C) { return split_mean_of(split_index(p.txm, R, C)); }
    __device__ __forceinline__ void load(int i, float &mean, float &var) const {
        mean = __ldcg(m + split_mean_of(i));
        var = __ldcg(v + i);
    }
    __device__ __forceinline__ float load_mean(int i) const { return __ldcg(m + split_mean_of(i)); }
    __device__ __forceinline__ float load_var(int i) const { return __ldcg(v + i); }
    __device__ __forceinline__ void store(int i, float mean, float var) const {
        m[split_mean_of(i)] = mean;
        v[i] = var;
    }
    __device__ __forceinline__ void store_mean(int i, float mean) const { m[split_mean_of(i)] = mean; }
    __device__ __forceinline__ void store_var(int i, float var) const { v[i] = var; }
};

template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(kThreads) ipp_step_kernel(const __grid_constant__ StepParams p) {
    __shared__ float2 s_taps[kWarpsPerBlock][2 * kTapCap * 3];

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int job = blockIdx.x * kWarpsPerBlock + wib;
    if (job >= p.n_jobs) return;

    const int env = p.env_index ? __ldg(p.env_index + job) : job;
    const Geom g = decode(p, job);

    const int nqx = (g.nx + 1) >> 1, nqy = (g.ny + 1) >> 1;
    const int nq = nqx * nqy;
    const bool quirk = (p.flags & IPP_FLAG_NO_DSIZE_QUIRK) == 0;
    const bool adaptive = (p.flags & IPP_FLAG_ADAPTIVE) != 0;
    const bool commit = (p.flags & IPP_FLAG_NO_COMMIT) == 0 && !p.measure_only;
    const bool simulate = (MODE != MODE_PREDICT) && (p.z_in == nullptr);
    const bool need_taps = simulate && g.rf == 2;

    // INTER_AREA geometry at rf = 2 (rows / cols of the down-sampled measurement D)
    const int out_r = quirk ? nqx : nqy;
    const int out_c = quirk ? nqy : nqx;
    int tap_mode = TAPS_FAST;
    TapView tapv;
    tapv.rows = s_taps[wib];
    tapv.cols = s_taps[wib] + 3 * kTapCap;
    if (MODE != MODE_PREDICT && need_taps) {
        if (out_r > g.ny || out_c > g.nx) {
            if (lane == 0) *(volatile int *)p.status = 1;  // mapped host word, bit 0 is the only bit
            return;
        }
        tap_mode = build_tap_tables<kTapCap>(s_taps[wib], lane, g.ny, g.nx, out_r, out_c);
    }

    FuseCtx fc;
    fc.rf = g.rf;
    fc.R = g.R;
    fc.invR = fast_rcp(g.R);
    const bool entropy = (p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;

    const float inv_nqx = __frcp_rn((float)nqx);
    const float inv_outc = __frcp_rn((float)out_c);
    const Belief<LAYOUT> bel(p, (size_t)env);
    const float *gt = p.gt + (size_t)env * p.plane_gt;
    const size_t nrow = (size_t)job * (size_t)p.noise_stride;

    float acc = 0.0f;  // per-lane partial (<= a few dozen quads); fp64 tree across the warp
    float nrm_cache[4] = {0.f, 0.f, 0.f, 0.f};  // rf = 2: one Philox call serves four passes
    int it = 0;  // pass: lane holds quad q = lane + 32 * it (the layout the noise stream is defined on, quad_math.cuh)
    for (int q = lane; q < nq; q += 32, ++it) {
        const int qy = fdiv(q, nqx, inv_nqx), qx = q - qy * nqx;
        const int r0 = 2 * qy, c0 = 2 * qx;
        const bool cok = c0 + 1 < g.nx, rok = r0 + 1 < g.ny;
        const bool ok[4] = {true, cok, rok, cok && rok};
        // cell offsets inside the env's belief / ground-truth arrays (clamped: cells past the footprint are never touched)
        const int R0 = g.yu + r0, C0 = g.xl + c0, R1 = R0 + (rok ? 1 : 0), C1 = C0 + (cok ? 1 : 0);
        const int off[4] = {Belief<LAYOUT>::idx(p, R0, C0), Belief<LAYOUT>::idx(p, R0, C1), Belief<LAYOUT>::idx(p, R1, C0),
                            Belief<LAYOUT>::idx(p, R1, C1)};

        // ---- gather belief ------------------------------------------------------------------
        float m[4], v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m[k] = 0.0f;
            v[k] = 0.0f;
            if (ok[k]) {
                if (MODE == MODE_KALMAN) {
                    bel.load(off[k], m[k], v[k]);
                } else if (MODE == MODE_PREDICT) {
                    v[k] = bel.load_var(off[k]);
                    if (adaptive) m[k] = bel.load_mean(off[k]);
                } else {
                    m[k] = bel.load_mean(off[k]);
                }
            }
        }

        // ---- measurement ----------------------------------------------------------------------
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE != MODE_PREDICT) {
            if (!simulate) {
                if (g.rf == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (ok[k]) z[k] = p.z_in[nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)];
                } else {
                    z[0] = p.z_in[nrow + q];
                }
            } else {
                float eps[4];
                if (p.noise != nullptr) {
                    if (g.rf == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? __ldg(p.noise + nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)) : 0.0f;
                    } else {
                        eps[0] = __ldg(p.noise + nrow + q);
                    }
                } else {
                    draw_normals(p, g.rf, q, lane, it, (uint32_t)env + p.env_id_offset, nrm_cache, eps);
                }
                if (g.rf == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (ok[k]) z[k] = __saturatef(fmaf(g.s2, eps[k], __ldg(gt + Belief<LAYOUT>::gidx(p, R0 + (k >> 1), C0 + (k & 1)))));
                } else {
                    // D[pr, pc] with the measurement's flat index q: (pr, pc) = (q / out_c, q % out_c)
                    const int pr = fdiv(q, out_c, inv_outc), pc = q - pr * out_c;
                    float d;
                    if (LAYOUT == IPP_LAYOUT_TILED)
                        d = downsample(tap_mode, GtTiled{gt, p.txg, g.yu, g.xl}, tapv, pr, pc, g.ny, g.nx, out_r, out_c);
                    else if (LAYOUT == IPP_LAYOUT_SUPER)
                        d = downsample(tap_mode, GtSuper{gt, p.txm, g.yu, g.xl}, tapv, pr, pc, g.ny, g.nx, out_r, out_c);
                    else if (LAYOUT == IPP_LAYOUT_SPLIT)
                        d = downsample(tap_mode, GtSplit{gt, p.txm, g.yu, g.xl}, tapv, pr, pc, g.ny, g.nx, out_r, out_c);
                    else
                        d = downsample(tap_mode, GtRowMajor{gt + g.yu * p.X + g.xl, p.X}, tapv, pr, pc, g.ny, g.nx, out_r, out_c);
                    z[0] = __saturatef(fmaf(g.s2, eps[0], d));
                }
            }
            if (p.z_out != nullptr) {
                if (g.rf == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (ok[k]) p.z_out[nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)] = z[k];
                } else {
                    p.z_out[nrow + q] = z[0];
                }
            }
        }
        if (p.measure_only) continue;

        // ---- fusion + reward --------------------------------------------------------------------
        if (MODE == MODE_LOGODDS) {
            const float gain = 0.5f * fc.invR;
            float dh = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!ok[k]) continue;
                const float zz = g.rf == 1 ? z[k] : z[0];
                const float l1 = fminf(fmaxf(fmaf(2.0f * zz - 1.0f, gain, m[k]), -30.0f), 30.0f);
                dh += bernoulli_entropy(m[k]) - bernoulli_entropy(l1);
                if (commit) bel.store_mean(off[k], l1);
            }
            acc += dh;
            continue;
        }

        float mn[4], vn[4];
        bool msk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!adaptive || (fmaf(p.kappa, v[k], m[k]) >= p.thr));
        acc += kalman_quad_rt(entropy, adaptive, fc, cok, rok, m, v, z, msk, mn, vn);
        if (commit) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!ok[k]) continue;
                if (MODE == MODE_KALMAN)
                    bel.store(off[k], mn[k], vn[k]);
                else
                    bel.store_var(off[k], vn[k]);
            }
        }
    }

    if (p.reward == nullptr && p.measure_only) return;

    // ---- warp-shuffle reduction of the per-env information gain ---------------------------------
    float accd = acc;  // fp32 tree: <= 32 partials of similar size, relative error ~3e-7
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, s);

    if (lane == 0) {
        const double *pv = p.prev_in != nullptr ? p.prev_in + 3 * (size_t)job : p.prev_state + 3 * (size_t)env;
        const float cost = job_cost(p, g.px, g.py, g.ph, pv[0], pv[1], pv[2]);
        if (p.reward != nullptr) p.reward[job] = accd * fast_rcp(cost + 1.0f);
        if (commit && (p.flags & IPP_FLAG_KEEP_PREV) == 0) {
            double *ps = p.prev_state + 3 * (size_t)env;
            ps[0] = g.px;
            ps[1] = g.py;
            ps[2] = g.ph;
        }
    }
}

}  // namespace ipp
