// step_kernel.cuh — the fused per-step kernel of the batched IPP environment engine (sm_100a).
//
// One warp owns one job (= one env step).  The footprint is tiled by 2x2-cell "quads" anchored at
// its top-left cell; lane l handles quads l, l+32, ...  A quad is exactly one measurement block at
// resolution factor 2 and four independent measurements at resolution factor 1, so the block sums
// of the Kalman update never leave a thread (no shuffles, no shared memory inside the loop); the
// only cross-lane step is the final reward reduction (warp shuffle tree, fp64).
//
// HBM-bound gather-update-reduce: per covered cell the kernel reads gt, mean, var and writes
// mean, var (20 B), nothing else touches DRAM.  No tensor cores on purpose.
//
// Reference semantics reproduced here (paths under the reference tree):
//   footprint            sensors/cameras.py:34-75
//   resolution factor    sensors/cameras.py:122-125
//   sigma2(h), R         sensors/models/sensor_models.py:27-36
//   measurement blocks   sensors/models/sensor_models.py:54-81  (partial block weight 1/rf)
//   measurement          simulations/simulations.py:26-34, simulations/sensor_manipulations.py:7-57
//                        (cv2 INTER_AREA incl. the dsize swap; noise variance used as std; clip)
//   Kalman update        mapping/mappings.py:155-197 restricted to a diagonal covariance
//   adaptive mask/reward planning/common/rewards.py:8-31
//   cost                 planning/common/actions.py:8-41
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ipp_b200.h"

namespace ipp {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

// kernel modes (template parameter)
constexpr int MODE_KALMAN = 0;   // full step: measure + mean/var update + reward
constexpr int MODE_PREDICT = 1;  // covariance-only (simulate_prediction_step)
constexpr int MODE_LOGODDS = 2;  // extension: log-odds fusion + Shannon entropy

struct AltLevel {
    double alt;   // altitude [m]
    int rx, ry;   // footprint radius in cells
    int rf;       // resolution factor
    float s2;     // sigma2(h)
    float R;      // rf^3 * sigma2(h)
    int pad;
};

struct StepParams {
    // belief / world (layout PLANES: mean, var separate; layout MV: mean points at float2 base)
    float *mean;
    float *var;
    const float *gt;
    size_t plane;  // y_dim * x_dim
    int X, Y;
    int n_jobs;
    int batch;
    // per-job inputs
    const int32_t *env_index;   // nullable
    const int32_t *action_ids;  // one of action_ids / poses
    const double *poses;
    const double *prev_in;      // nullable: explicit previous actions [n_jobs][3]
    double *prev_state;         // engine previous actions [batch][3]
    const float *noise;         // nullable -> Philox
    const float *z_in;          // nullable: measurements supplied by the caller
    float *z_out;               // nullable
    int noise_stride;
    float *reward;              // nullable (measure-only)
    int *status;                // device status word (bit 0: unsupported up-sampling footprint)
    // configuration
    double res, tan_x, tan_y, coeff_a, coeff_b, rf_alt, max_v, max_a;
    float thr, kappa;
    int cost_mode;
    int n_levels;
    uint32_t flags;
    uint32_t measure_only;
    uint32_t seed_lo, seed_hi, step_lo, step_hi;
    uint32_t env_id_offset;
    AltLevel lut[IPP_MAX_ALTITUDE_LEVELS];
};

// ---------------------------------------------------------------------------------------------
// belief accessors for the two HBM layouts
// ---------------------------------------------------------------------------------------------
template <int LAYOUT>
struct Belief;

template <>
struct Belief<IPP_LAYOUT_PLANES> {
    float *m, *v;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : m(p.mean + env * p.plane), v(p.var + env * p.plane) {}
    __device__ __forceinline__ void load(int i, float &mean, float &var) const {
        mean = m[i];
        var = v[i];
    }
    __device__ __forceinline__ float load_mean(int i) const { return m[i]; }
    __device__ __forceinline__ float load_var(int i) const { return v[i]; }
    __device__ __forceinline__ void store(int i, float mean, float var) const {
        m[i] = mean;
        v[i] = var;
    }
    __device__ __forceinline__ void store_mean(int i, float mean) const { m[i] = mean; }
    __device__ __forceinline__ void store_var(int i, float var) const { v[i] = var; }
};

template <>
struct Belief<IPP_LAYOUT_MV> {
    float2 *mv;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : mv(reinterpret_cast<float2 *>(p.mean) + env * p.plane) {}
    __device__ __forceinline__ void load(int i, float &mean, float &var) const {
        float2 t = mv[i];
        mean = t.x;
        var = t.y;
    }
    __device__ __forceinline__ float load_mean(int i) const { return mv[i].x; }
    __device__ __forceinline__ float load_var(int i) const { return mv[i].y; }
    __device__ __forceinline__ void store(int i, float mean, float var) const { mv[i] = make_float2(mean, var); }
    __device__ __forceinline__ void store_mean(int i, float mean) const { mv[i].x = mean; }
    __device__ __forceinline__ void store_var(int i, float var) const { mv[i].y = var; }
};

// ---------------------------------------------------------------------------------------------
// counter-based RNG: Philox4x32-10 (Random123) + Box-Muller.  Mirrored in oracle/ipp_oracle.py
// (device_normals / device_noise_field).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float r = sqrtf(-2.0f * logf(u01(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// ---------------------------------------------------------------------------------------------
// per-job geometry (footprint, sensor model, cost) — computed redundantly by every lane (SIMT: one
// issue slot either way); integer / fp64 so that floor() and the clip agree with NumPy bit for bit.
// ---------------------------------------------------------------------------------------------
struct Geom {
    int xl, yu, nx, ny, rf;
    float s2, R;
    double px, py, ph;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ Geom decode(const StepParams &p, int job) {
    Geom g;
    int cx, cy, rx, ry;
    if (p.action_ids != nullptr) {
        // planning/common/actions.py:73-91: id = level*N + x_dim*col + row; pose = res*idx + res/2.
        const int id = p.action_ids[job];
        const int N = p.X * p.Y;
        int lvl = id / N;
        lvl = clampi(lvl, 0, p.n_levels - 1);
        const int i = id - lvl * N;
        int col = i / p.X;
        int row = i - col * p.X;
        col = clampi(col, 0, p.X - 1);
        row = clampi(row, 0, p.Y - 1);
        const AltLevel &L = p.lut[lvl];
        cx = col;
        cy = row;
        rx = L.rx;
        ry = L.ry;
        g.rf = L.rf;
        g.s2 = L.s2;
        g.R = L.R;
        g.px = __dadd_rn(__dmul_rn(p.res, (double)col), __dmul_rn(0.5, p.res));
        g.py = __dadd_rn(__dmul_rn(p.res, (double)row), __dmul_rn(0.5, p.res));
        g.ph = L.alt;
    } else {
        g.px = p.poses[3 * (size_t)job + 0];
        g.py = p.poses[3 * (size_t)job + 1];
        g.ph = p.poses[3 * (size_t)job + 2];
        // sensors/cameras.py:44-45,62-66 — same operation order, no fma contraction.
        const double xm = __dmul_rn(__dmul_rn(2.0, g.ph), p.tan_x);
        const double ym = __dmul_rn(__dmul_rn(2.0, g.ph), p.tan_y);
        const double wx = floor(__ddiv_rn(xm, p.res));
        const double wy = floor(__ddiv_rn(ym, p.res));
        const double fcx = floor(__ddiv_rn(g.px, p.res));
        const double fcy = floor(__ddiv_rn(g.py, p.res));
        const double frx = floor(__dmul_rn(0.5, wx));
        const double fry = floor(__dmul_rn(0.5, wy));
        const double lim = 1.0e9;
        cx = (int)fmin(fmax(fcx, -lim), lim);
        cy = (int)fmin(fmax(fcy, -lim), lim);
        rx = (int)fmin(fmax(frx, 0.0), lim);
        ry = (int)fmin(fmax(fry, 0.0), lim);
        g.rf = g.ph > p.rf_alt ? 2 : 1;
        const double s2 = p.coeff_a * (1.0 - exp(-p.coeff_b * g.ph));
        g.s2 = (float)s2;
        g.R = (float)((double)(g.rf * g.rf * g.rf) * s2);
    }
    const long long xl = (long long)cx - rx, xr = (long long)cx + rx;
    const long long yu = (long long)cy - ry, yd = (long long)cy + ry;
    const int xli = (int)(xl < 0 ? 0 : (xl > p.X - 1 ? p.X - 1 : xl));
    const int xri = (int)(xr < 0 ? 0 : (xr > p.X - 1 ? p.X - 1 : xr));
    const int yui = (int)(yu < 0 ? 0 : (yu > p.Y - 1 ? p.Y - 1 : yu));
    const int ydi = (int)(yd < 0 ? 0 : (yd > p.Y - 1 ? p.Y - 1 : yd));
    g.xl = xli;
    g.yu = yui;
    g.nx = xri - xli + 1;
    g.ny = ydi - yui + 1;
    return g;
}

__device__ __forceinline__ double job_cost(const StepParams &p, const Geom &g, double qx, double qy, double qh) {
    // planning/common/actions.py:15-16 / 32-41
    const double dx = g.px - qx, dy = g.py - qy, dz = g.ph - qh;
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    if (p.cost_mode == IPP_COST_DISTANCE) return d;
    const double d_acc = fmin(d * 0.5, (p.max_v * p.max_v) / (2.0 * p.max_a));
    const double d_const = d - 2.0 * d_acc;
    return d_const / p.max_v + 2.0 * sqrt(2.0 * d_acc / p.max_a);
}

// One axis of cv2 INTER_AREA decimation: output sample o of n_out integrates the input over
// [o*s, (o+1)*s), s = n_in/n_out.  Exact integer overlaps in units of 1/n_out; weight =
// overlap / n_in.  (opencv resize.cpp computeResizeAreaTab; reference call site
// simulations/sensor_manipulations.py:20-22.)
struct Taps {
    int start;   // first input index
    int count;   // number of taps
    int a1, a2;  // o*n_in, (o+1)*n_in
};
__device__ __forceinline__ Taps make_taps(int o, int n_in, int n_out) {
    Taps t;
    t.a1 = o * n_in;
    t.a2 = t.a1 + n_in;
    t.start = t.a1 / n_out;
    const int end = (t.a2 + n_out - 1) / n_out;  // exclusive
    t.count = end - t.start;
    return t;
}
__device__ __forceinline__ float tap_weight(const Taps &t, int k, int n_in, int n_out, float inv_n_in) {
    const int i = t.start + k;
    const int lo = max(t.a1, i * n_out);
    const int hi = min(t.a2, (i + 1) * n_out);
    return hi > lo ? (float)(hi - lo) * inv_n_in : 0.0f;
}

// Shannon entropy [nats] of Bernoulli(sigmoid(l)):  log1p(e^-|l|) + |l| e^-|l| / (1 + e^-|l|)
__device__ __forceinline__ float bernoulli_entropy(float l) {
    const float a = fabsf(l);
    const float e = expf(-a);
    return log1pf(e) + a * e / (1.0f + e);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(kThreads) ipp_step_kernel(const __grid_constant__ StepParams p) {
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (job >= p.n_jobs) return;

    const int env = p.env_index ? p.env_index[job] : job;
    const Geom g = decode(p, job);

    const int nqx = (g.nx + 1) >> 1, nqy = (g.ny + 1) >> 1;
    const int nq = nqx * nqy;
    const bool quirk = (p.flags & IPP_FLAG_NO_DSIZE_QUIRK) == 0;
    const bool adaptive = (p.flags & IPP_FLAG_ADAPTIVE) != 0;
    const bool entropy = (p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;
    const bool commit = (p.flags & IPP_FLAG_NO_COMMIT) == 0 && !p.measure_only;
    const bool simulate = (MODE != MODE_PREDICT) && (p.z_in == nullptr);

    // INTER_AREA geometry at rf = 2 (output rows/cols of the down-sampled measurement)
    const int out_r = quirk ? nqx : nqy;  // rows of D
    const int out_c = quirk ? nqy : nqx;  // cols of D
    if (MODE != MODE_PREDICT && g.rf == 2 && simulate && (out_r > g.ny || out_c > g.nx)) {
        if (lane == 0) atomicOr(p.status, 1);
        return;
    }
    const float inv_ny = 1.0f / (float)g.ny, inv_nx = 1.0f / (float)g.nx;

    const Belief<LAYOUT> bel(p, (size_t)env);
    const float *gt = p.gt + (size_t)env * p.plane;
    const int X = p.X;
    const int origin = g.yu * X + g.xl;
    const size_t nrow = (size_t)job * (size_t)p.noise_stride;

    double acc = 0.0;

    for (int q = lane; q < nq; q += 32) {
        const int qy = q / nqx, qx = q - qy * nqx;
        const int r0 = 2 * qy, c0 = 2 * qx;
        const bool cok = c0 + 1 < g.nx, rok = r0 + 1 < g.ny;
        const bool ok[4] = {true, cok, rok, cok && rok};
        const int i00 = origin + r0 * X + c0;
        const int off[4] = {i00, i00 + 1, i00 + X, i00 + X + 1};

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
                        for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? p.noise[nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)] : 0.0f;
                    } else {
                        eps[0] = p.noise[nrow + q];
                    }
                } else {
                    uint32_t rnd[4];
                    philox4x32_10((uint32_t)q, (uint32_t)env + p.env_id_offset, p.step_lo, p.step_hi, p.seed_lo, p.seed_hi, rnd);
                    box_muller(rnd[0], rnd[1], eps[0], eps[1]);
                    if (g.rf == 1) box_muller(rnd[2], rnd[3], eps[2], eps[3]);
                }
                if (g.rf == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (ok[k]) z[k] = fminf(fmaxf(fmaf(g.s2, eps[k], __ldg(gt + off[k])), 0.0f), 1.0f);
                } else {
                    // D[pr, pc] with the measurement's flat index q: (pr, pc) = (q / out_c, q % out_c)
                    const int pr = q / out_c, pc = q - pr * out_c;
                    const Taps tr = make_taps(pr, g.ny, out_r);
                    const Taps tc = make_taps(pc, g.nx, out_c);
                    float d = 0.0f;
                    if (tr.count <= 3 && tc.count <= 3) {
                        float wr[3], wc[3];
                        int ro[3], co[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            wr[k] = tap_weight(tr, k, g.ny, out_r, inv_ny);
                            wc[k] = tap_weight(tc, k, g.nx, out_c, inv_nx);
                            ro[k] = origin + min(tr.start + k, g.ny - 1) * X;
                            co[k] = min(tc.start + k, g.nx - 1);
                        }
                        float gv[9];
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int b = 0; b < 3; ++b) gv[3 * a + b] = __ldg(gt + ro[a] + co[b]);
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float rowsum = fmaf(wc[2], gv[3 * a + 2], fmaf(wc[1], gv[3 * a + 1], wc[0] * gv[3 * a]));
                            d = fmaf(wr[a], rowsum, d);
                        }
                    } else {
                        for (int a = 0; a < tr.count; ++a) {
                            const float wra = tap_weight(tr, a, g.ny, out_r, inv_ny);
                            const int rbase = origin + min(tr.start + a, g.ny - 1) * X;
                            float rowsum = 0.0f;
                            for (int b = 0; b < tc.count; ++b)
                                rowsum = fmaf(tap_weight(tc, b, g.nx, out_c, inv_nx), __ldg(gt + rbase + min(tc.start + b, g.nx - 1)), rowsum);
                            d = fmaf(wra, rowsum, d);
                        }
                    }
                    z[0] = fminf(fmaxf(fmaf(g.s2, eps[0], d), 0.0f), 1.0f);
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
            const float gain = 1.0f / (2.0f * g.R);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!ok[k]) continue;
                const float zz = g.rf == 1 ? z[k] : z[0];
                const float l1 = fminf(fmaxf(fmaf(2.0f * zz - 1.0f, gain, m[k]), -30.0f), 30.0f);
                acc += (double)(bernoulli_entropy(m[k]) - bernoulli_entropy(l1));
                if (commit) bel.store_mean(off[k], l1);
            }
            continue;
        }

        float mn[4], vn[4], dl[4];
        if (g.rf == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float S = v[k] + g.R;
                const float invS = 1.0f / S;
                const float gain = v[k] * invS;
                vn[k] = gain * g.R;                      // v R / (v + R)  ==  v - v^2/S, cancellation-free
                mn[k] = fmaf(gain, z[k] - m[k], m[k]);
                dl[k] = entropy ? 0.5f * logf(S / g.R) : v[k] * gain;
            }
        } else {
            const int cnt = 1 + (int)cok + (int)rok + (int)(cok && rok);
            const float w = cnt == 4 ? 0.25f : 0.5f;  // sensor_models.py:76-79
            const float w2 = w * w;
            const float sv = (v[0] + v[1]) + (v[2] + v[3]);
            const float sm = (m[0] + m[1]) + (m[2] + m[3]);
            const float S = fmaf(w2, sv, g.R);
            const float invS = 1.0f / S;
            const float innov = z[0] - w * sm;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float rest = S - w2 * v[k];         // w^2 * sum_{j != k} v_j + R  > 0
                vn[k] = v[k] * rest * invS;
                mn[k] = fmaf(w * v[k] * invS, innov, m[k]);
                dl[k] = entropy ? 0.5f * logf(S / rest) : w2 * v[k] * v[k] * invS;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
            const bool in_mask = !adaptive || (fmaf(p.kappa, v[k], m[k]) >= p.thr);
            if (in_mask) acc += (double)dl[k];
            if (commit) {
                if (MODE == MODE_KALMAN)
                    bel.store(off[k], mn[k], vn[k]);
                else
                    bel.store_var(off[k], vn[k]);
            }
        }
    }

    if (p.reward == nullptr && p.measure_only) return;

    // ---- warp-shuffle reduction of the per-env information gain ---------------------------------
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);

    if (lane == 0) {
        double qx, qy, qh;
        if (p.prev_in != nullptr) {
            qx = p.prev_in[3 * (size_t)job + 0];
            qy = p.prev_in[3 * (size_t)job + 1];
            qh = p.prev_in[3 * (size_t)job + 2];
        } else {
            qx = p.prev_state[3 * (size_t)env + 0];
            qy = p.prev_state[3 * (size_t)env + 1];
            qh = p.prev_state[3 * (size_t)env + 2];
        }
        const double cost = job_cost(p, g, qx, qy, qh);
        if (p.reward != nullptr) p.reward[job] = (float)(acc / (cost + 1.0));
        if (commit && (p.flags & IPP_FLAG_KEEP_PREV) == 0) {
            p.prev_state[3 * (size_t)env + 0] = g.px;
            p.prev_state[3 * (size_t)env + 1] = g.py;
            p.prev_state[3 * (size_t)env + 2] = g.ph;
        }
    }
}

}  // namespace ipp
