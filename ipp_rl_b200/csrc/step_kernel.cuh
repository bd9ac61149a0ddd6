// step_kernel.cuh — the fused per-step kernel of the batched IPP environment engine (sm_100a).
//
// One warp owns one job (= one env step).  The footprint is tiled by 2x2-cell "quads" anchored at
// its top-left cell; lane l handles quads l, l+32, ...  A quad is exactly one measurement block at
// resolution factor 2 and four independent measurements at resolution factor 1, so the block sums
// of the Kalman update never leave a thread (no shuffles inside the loop); the only cross-lane step
// is the final reward reduction (warp shuffle tree, fp64).  The cv2 INTER_AREA tap tables of a
// footprint are built once per env by the warp and staged in shared memory.
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
constexpr int kTapCap = 64;  // tap-table entries per axis staged in shared memory

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
    double res, tan_x, tan_y, coeff_a, coeff_b, rf_alt;
    float max_v, max_a;
    float thr, kappa;
    int cost_mode;
    int n_levels;
    float inv_N, inv_X;  // 1/(X*Y), 1/X for the action-id decode
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

template <>
struct Belief<IPP_LAYOUT_MV> {
    float2 *mv;
    __device__ __forceinline__ Belief(const StepParams &p, size_t env) : mv(reinterpret_cast<float2 *>(p.mean) + env * p.plane) {}
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

// ---------------------------------------------------------------------------------------------
// counter-based RNG: Philox4x32-10 (Random123) + Box-Muller.  Mirrored in oracle/ipp_oracle.py
// (device_normals / device_noise_field) and oracle/ipp_oracle.c.
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

// u = (x + 0.5) * 2^-32 in (0, 1)
__device__ __forceinline__ float u01(uint32_t x) { return fmaf(__uint2float_rn(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// n0, n1 = sqrt(-2 ln u(a)) * (cos, sin)(pi * (2 u(b) - 1)); the angle lies in [-pi, pi) where the
// SFU sin/cos have their best absolute accuracy (2^-21.4).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float r = sqrtf(-2.0f * logf(u01(a)));
    const float th = 3.14159265358979f * fmaf(2.0f, u01(b), -1.0f);
    n0 = r * __cosf(th);
    n1 = r * __sinf(th);
}

// floor(n / d) for 0 <= n < 2^31, d >= 1 and n/d < 2^20: float estimate (error < 1) + one correction.
__device__ __forceinline__ int fdiv(int n, int d, float inv_d) {
    int q = (int)(__int2float_rz(n) * inv_d);
    const int r = n - q * d;
    q += (r >= d) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}

// ---------------------------------------------------------------------------------------------
// per-job geometry (footprint, sensor model) — computed redundantly by every lane (SIMT: one
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
        const int N = p.X * p.Y;
        const int id = clampi(__ldg(p.action_ids + job), 0, p.n_levels * N - 1);
        const int lvl = fdiv(id, N, p.inv_N);
        const int i = id - lvl * N;
        int col = fdiv(i, p.X, p.inv_X);
        int row = i - col * p.X;
        col = min(col, p.X - 1);
        row = min(row, p.Y - 1);
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

// planning/common/actions.py:15-16 / 32-41.  The pose difference is formed in fp64; the norm and the
// trapezoidal-profile time are fp32 (relative error ~1e-7, two orders below the parity tolerance).
__device__ __forceinline__ float job_cost(const StepParams &p, const Geom &g, double qx, double qy, double qh) {
    const float dx = (float)(g.px - qx), dy = (float)(g.py - qy), dz = (float)(g.ph - qh);
    const float d = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    if (p.cost_mode == IPP_COST_DISTANCE) return d;
    const float d_acc = fminf(d * 0.5f, (p.max_v * p.max_v) / (2.0f * p.max_a));
    const float d_const = d - 2.0f * d_acc;
    return d_const / p.max_v + 2.0f * sqrtf(2.0f * d_acc / p.max_a);
}

// One axis of cv2 INTER_AREA decimation: output sample o of n_out integrates the input over
// [o*s, (o+1)*s), s = n_in/n_out.  Exact integer overlaps in units of 1/n_out; weight =
// overlap / n_in.  (opencv resize.cpp computeResizeAreaTab; reference call site
// simulations/sensor_manipulations.py:20-22.)  Entry = {first input index, w0, w1, w2}; count > 3
// (scale > 2, only for clipped non-square footprints) is flagged with a negative start.
__device__ __forceinline__ float4 make_tap_entry(int o, int n_in, int n_out) {
    const int a1 = o * n_in, a2 = a1 + n_in;
    const int start = a1 / n_out;
    const int end = (a2 + n_out - 1) / n_out;  // exclusive
    const float inv = 1.0f / (float)n_in;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = start + k;
        const int lo = max(a1, i * n_out), hi = min(a2, (i + 1) * n_out);
        w[k] = hi > lo ? (float)(hi - lo) * inv : 0.0f;
    }
    return make_float4(__int_as_float(end - start > 3 ? -1 - start : start), w[0], w[1], w[2]);
}

// generic weight of input i for output o (slow path)
__device__ __forceinline__ float tap_weight_generic(int o, int i, int n_in, int n_out) {
    const int a1 = o * n_in, a2 = a1 + n_in;
    const int lo = max(a1, i * n_out), hi = min(a2, (i + 1) * n_out);
    return hi > lo ? (float)(hi - lo) / (float)n_in : 0.0f;
}

// Shannon entropy [nats] of Bernoulli(sigmoid(l)):  log1p(e^-|l|) + |l| e^-|l| / (1 + e^-|l|)
__device__ __forceinline__ float bernoulli_entropy(float l) {
    const float a = fabsf(l);
    const float e = __expf(-a);
    return log1pf(e) + a * e * __frcp_rn(1.0f + e);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int LAYOUT, int MODE>
__global__ void __launch_bounds__(kThreads) ipp_step_kernel(const __grid_constant__ StepParams p) {
    __shared__ float4 s_taps[kWarpsPerBlock][2 * kTapCap];

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
    const bool entropy = (p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;
    const bool commit = (p.flags & IPP_FLAG_NO_COMMIT) == 0 && !p.measure_only;
    const bool simulate = (MODE != MODE_PREDICT) && (p.z_in == nullptr);
    const bool downsample = simulate && g.rf == 2;

    // INTER_AREA geometry at rf = 2 (rows / cols of the down-sampled measurement D)
    const int out_r = quirk ? nqx : nqy;
    const int out_c = quirk ? nqy : nqx;
    bool generic_taps = false;
    if (MODE != MODE_PREDICT && downsample) {
        if (out_r > g.ny || out_c > g.nx) {
            if (lane == 0) atomicOr(p.status, 1);
            return;
        }
        bool bad = out_r > kTapCap || out_c > kTapCap;
        if (!bad) {
            for (int idx = lane; idx < out_r + out_c; idx += 32) {
                const bool is_row = idx < out_r;
                const float4 e = is_row ? make_tap_entry(idx, g.ny, out_r) : make_tap_entry(idx - out_r, g.nx, out_c);
                s_taps[wib][is_row ? idx : kTapCap + idx - out_r] = e;
                bad |= __float_as_int(e.x) < 0;
            }
        }
        generic_taps = __any_sync(0xffffffffu, bad);
        __syncwarp();
    }

    const float inv_nqx = __frcp_rn((float)nqx);
    const float inv_outc = __frcp_rn((float)out_c);
    const Belief<LAYOUT> bel(p, (size_t)env);
    const float *gt = p.gt + (size_t)env * p.plane;
    const int X = p.X;
    const int origin = g.yu * X + g.xl;
    const size_t nrow = (size_t)job * (size_t)p.noise_stride;
    const float invR = __frcp_rn(g.R);

    double acc = 0.0;

    for (int q = lane; q < nq; q += 32) {
        const int qy = fdiv(q, nqx, inv_nqx), qx = q - qy * nqx;
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
                        for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? __ldg(p.noise + nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)) : 0.0f;
                    } else {
                        eps[0] = __ldg(p.noise + nrow + q);
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
                        if (ok[k]) z[k] = __saturatef(fmaf(g.s2, eps[k], __ldg(gt + off[k])));
                } else {
                    // D[pr, pc] with the measurement's flat index q: (pr, pc) = (q / out_c, q % out_c)
                    const int pr = fdiv(q, out_c, inv_outc), pc = q - pr * out_c;
                    float d = 0.0f;
                    if (!generic_taps) {
                        const float4 tr = s_taps[wib][pr], tc = s_taps[wib][kTapCap + pc];
                        const int rs = __float_as_int(tr.x), cs = __float_as_int(tc.x);
                        const float wr[3] = {tr.y, tr.z, tr.w};
                        const int cb[3] = {cs, min(cs + 1, g.nx - 1), min(cs + 2, g.nx - 1)};
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float *row = gt + origin + min(rs + a, g.ny - 1) * X;
                            const float rowsum = fmaf(tc.w, __ldg(row + cb[2]), fmaf(tc.z, __ldg(row + cb[1]), tc.y * __ldg(row + cb[0])));
                            d = fmaf(wr[a], rowsum, d);
                        }
                    } else {
                        const int rs = (pr * g.ny) / out_r, re = ((pr + 1) * g.ny + out_r - 1) / out_r;
                        const int cs = (pc * g.nx) / out_c, ce = ((pc + 1) * g.nx + out_c - 1) / out_c;
                        for (int a = rs; a < re; ++a) {
                            const float *row = gt + origin + min(a, g.ny - 1) * X;
                            float rowsum = 0.0f;
                            for (int b = cs; b < ce; ++b) rowsum = fmaf(tap_weight_generic(pc, b, g.nx, out_c), __ldg(row + min(b, g.nx - 1)), rowsum);
                            d = fmaf(tap_weight_generic(pr, a, g.ny, out_r), rowsum, d);
                        }
                    }
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
            const float gain = 0.5f * invR;
            float dh = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!ok[k]) continue;
                const float zz = g.rf == 1 ? z[k] : z[0];
                const float l1 = fminf(fmaxf(fmaf(2.0f * zz - 1.0f, gain, m[k]), -30.0f), 30.0f);
                dh += bernoulli_entropy(m[k]) - bernoulli_entropy(l1);
                if (commit) bel.store_mean(off[k], l1);
            }
            acc += (double)dh;
            continue;
        }

        float mn[4], vn[4];
        float gain_q = 0.0f;  // this quad's contribution to the information gain
        bool msk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!adaptive || (fmaf(p.kappa, v[k], m[k]) >= p.thr));

        if (g.rf == 1) {
            float prod = 1.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float S = v[k] + g.R;
                const float gain = v[k] * __frcp_rn(S);
                vn[k] = gain * g.R;  // v R / (v + R)  ==  v - v^2/S, cancellation-free
                mn[k] = fmaf(gain, z[k] - m[k], m[k]);
                if (entropy)
                    prod *= msk[k] ? S * invR : 1.0f;  // v/v' = S/R
                else
                    gain_q += msk[k] ? v[k] * gain : 0.0f;
            }
            if (entropy) gain_q = 0.5f * __logf(prod);
        } else {
            const int cnt = 1 + (int)cok + (int)rok + (int)(cok && rok);
            const float w = cnt == 4 ? 0.25f : 0.5f;  // sensor_models.py:76-79
            const float w2 = w * w;
            const float sv = (v[0] + v[1]) + (v[2] + v[3]);
            const float sm = (m[0] + m[1]) + (m[2] + m[3]);
            const float S = fmaf(w2, sv, g.R);
            const float invS = __frcp_rn(S);
            const float innov = z[0] - w * sm;
            float rest[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rest[k] = fmaf(-w2, v[k], S);  // w^2 * sum_{j != k} v_j + R  > 0
                const float vk_invS = v[k] * invS;
                vn[k] = vk_invS * rest[k];
                mn[k] = fmaf(w * vk_invS, innov, m[k]);
                if (!entropy) gain_q += msk[k] ? w2 * v[k] * vk_invS : 0.0f;
            }
            if (entropy) {
                // prod_k v_k / v'_k = prod_k S / rest_k over the masked cells
                const float a = (msk[0] ? rest[0] : S) * (msk[1] ? rest[1] : S);
                const float b = (msk[2] ? rest[2] : S) * (msk[3] ? rest[3] : S);
                const float S2 = S * S;
                gain_q = 0.5f * __logf((S2 * __frcp_rn(a)) * (S2 * __frcp_rn(b)));
            }
        }
        acc += (double)gain_q;
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
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);

    if (lane == 0) {
        const double *pv = p.prev_in != nullptr ? p.prev_in + 3 * (size_t)job : p.prev_state + 3 * (size_t)env;
        const float cost = job_cost(p, g, pv[0], pv[1], pv[2]);
        if (p.reward != nullptr) p.reward[job] = (float)acc * __frcp_rn(cost + 1.0f);
        if (commit && (p.flags & IPP_FLAG_KEEP_PREV) == 0) {
            double *ps = p.prev_state + 3 * (size_t)env;
            ps[0] = g.px;
            ps[1] = g.py;
            ps[2] = g.ph;
        }
    }
}

}  // namespace ipp
