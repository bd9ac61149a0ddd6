// quad_math.cuh — device helpers shared by the two step kernels (step_kernel.cuh: LSU gather,
// step_tma.cuh: TMA-staged persistent pipeline): parameters, action decode / footprint, sensor
// model, counter-based RNG, cv2 INTER_AREA taps, the per-quad Kalman fusion and the cost model.
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
    size_t plane;     // cells per env in the belief arrays (y_dim * x_dim; TILED: padded to whole 4x4 tiles; SPLIT: floats of the var array)
    size_t plane_gt;  // cells per env in the ground-truth array (TILED: padded to whole 8x4 tiles; SPLIT: floats of the {mean | gt} array)
    int txm, txg;     // TILED / SUPER: tiles per tile-row of the belief (ceil(X/4)) and of the ground truth (ceil(X/8); SUPER: = txm)
    int ts_mv, ts_gt; // TILED / SUPER: tile stride of the belief [float2] (16 / 24) and of the ground truth [float] (32 / 48)
    int gw_shift;     // TILED / SUPER: log2 of the ground-truth tile width (3 / 2)
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
    float inv_v, inv_a, d_acc_max;  // 1/max_v, 1/max_a, max_v^2 / (2 max_a)
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
// ln(u) on the SFU: lg2 is accurate to ~2^-22 ABSOLUTE, which is not enough next to u = 1 where ln u
// itself is tiny, so that corner uses the series of ln(1 + t), t = u - 1 (exact in fp32 for u >= 0.5).
__device__ __forceinline__ float fast_ln01(float u) {
    const float t = u - 1.0f;
    const float series = t * fmaf(t, fmaf(t, fmaf(t, -0.25f, 0.33333334f), -0.5f), 1.0f);
    return t > -0.03125f ? series : 0.69314718f * __log2f(u);
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
    const float r2 = fmaxf(-2.0f * fast_ln01(u01(a)), 1.0e-30f);
    const float r = r2 * rsqrtf(r2);
    const float th = 3.14159265358979f * fmaf(2.0f, u01(b), -1.0f);
    n0 = r * __cosf(th);
    n1 = r * __sinf(th);
}

// The engine's noise stream (mirrored by oracle device_noise_field).  A footprint is tiled by 2x2-cell quads anchored at
// its top-left cell, quad q = qy * nqx + qx (nqx = ceil(nx / 2) quads per row); every step kernel sweeps the quads as
// q = lane + 32 * pass.  Philox group g draws philox4x32_10(counter = (g, env, step), key = seed) -> four normals.
//   rf = 1: quad q draws group q; its four cells use the group's four normals.
//   rf = 2: measurement i = block q uses normal (pass & 3) of group lane + 32 * (pass >> 2), i.e. normal (i >> 5) & 3 of group
//           (i & 31) + 32 * (i >> 7): a lane keeps the four normals of ONE Philox call for four consecutive passes — one
//           call per four measurement blocks instead of one each, no cross-lane traffic.
__device__ __forceinline__ void philox_normals(const StepParams &p, uint32_t group, uint32_t env_id, float (&n)[4]) {
    uint32_t rnd[4];
    philox4x32_10(group, env_id, p.step_lo, p.step_hi, p.seed_lo, p.seed_hi, rnd);
    box_muller(rnd[0], rnd[1], n[0], n[1]);
    box_muller(rnd[2], rnd[3], n[2], n[3]);
}
// the caller holds quad q = lane + 32 * pass; `cache` must persist across the passes of one env
__device__ __forceinline__ void draw_normals(const StepParams &p, int rf, int q, int lane, int pass, uint32_t env_id, float (&cache)[4],
                                             float (&eps)[4]) {
    if (rf == 1) {
        philox_normals(p, (uint32_t)q, env_id, eps);
    } else {
        const int t = pass & 3;
        if (t == 0) philox_normals(p, (uint32_t)(lane + 32 * (pass >> 2)), env_id, cache);  // warp-uniform: every lane is in the same pass
        eps[0] = t == 0 ? cache[0] : (t == 1 ? cache[1] : (t == 2 ? cache[2] : cache[3]));
    }
}

// floor(n / d) for 0 <= n < 2^31, d >= 1 and n/d < 2^20: float estimate (error < 1) + one correction.
__device__ __forceinline__ int fdiv(int n, int d, float inv_d) {
    int q = (int)(__int2float_rz(n) * inv_d);
    const int r = n - q * d;
    q += (r >= d) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// IPP_LAYOUT_TILED: every 128-byte line of HBM holds a compact 2-D tile, so that the lines a footprint pulls in
// (L2 fetches whole 128 B lines from DRAM on this part) are mostly cells the step needs:
//   belief  float2 {mean,var}: 4 x 4 cells per line;   ground truth float: 8 (x) x 4 (y) cells per line;
//   tiles row-major over the map, cells row-major inside a tile.
__device__ __forceinline__ int tiled_mv_index(int txm, int R, int C) { return (((R >> 2) * txm + (C >> 2)) << 4) + ((R & 3) << 2) + (C & 3); }
__device__ __forceinline__ int tiled_gt_index(int txg, int R, int C) { return (((R >> 2) * txg + (C >> 3)) << 5) + ((R & 3) << 3) + (C & 7); }
// IPP_LAYOUT_SUPER: one 192-byte super-tile per 4 x 4 cells = [16 x float2 {mean,var} | 16 x float gt], super-tiles row-major
// over the map: the tiles a footprint touches in one tile row are ONE contiguous run of ntx * 192 bytes holding everything the
// fused step needs (belief and ground truth), i.e. one bulk copy (cp.async.bulk) per tile row and ~1 KB DRAM bursts.
// The belief index is in float2 units (24 per super-tile), the ground-truth index in floats relative to base + 32 floats
// (48 per super-tile).
constexpr int kSuperTileBytes = 192;
__device__ __forceinline__ int super_mv_index(int tx, int R, int C) { return ((R >> 2) * tx + (C >> 2)) * 24 + ((R & 3) << 2) + (C & 3); }
__device__ __forceinline__ int super_gt_index(int tx, int R, int C) { return ((R >> 2) * tx + (C >> 2)) * 48 + ((R & 3) << 2) + (C & 3); }
// IPP_LAYOUT_SPLIT: the super-tile cut in two arrays over the same 4 x 4-cell tiling: var[tile][16] (64 B) and
// {mean[16] | gt[16]}[tile] (128 B, line aligned).  One compact index i = tile * 16 + cell serves both: the variance sits at
// var[i], the mean at meangt[i + (i & ~15)] (= tile * 32 + cell), the ground truth 16 floats behind the mean.
constexpr int kSplitVarTileBytes = 64, kSplitMgTileBytes = 128;
__device__ __forceinline__ int split_index(int tx, int R, int C) { return (((R >> 2) * tx + (C >> 2)) << 4) + ((R & 3) << 2) + (C & 3); }
__device__ __forceinline__ int split_mean_of(int i) { return i + (i & ~15); }
// run-time forms for the streaming kernels (layout TILED, SUPER or SPLIT; strides from StepParams / TiledDims)
__device__ __forceinline__ size_t tiled_mv_index_rt(int txm, int ts_mv, int R, int C) {
    return (size_t)((R >> 2) * txm + (C >> 2)) * ts_mv + ((R & 3) << 2) + (C & 3);
}
__device__ __forceinline__ size_t tiled_gt_index_rt(int txg, int ts_gt, int gw_shift, int R, int C) {
    return (size_t)((R >> 2) * txg + (C >> gw_shift)) * ts_gt + ((R & 3) << gw_shift) + (C & ((1 << gw_shift) - 1));
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

// planning/common/actions.py:73-91: id = level*N + x_dim*col + row
// Ids outside [0, levels * N) are clamped (memory safety) and reported through the status word (bit 1 -> IPP_ERR_INVALID).
// On non-square grids the reference's formula is not a bijection (col or row can leave the map for ids < levels * N): those
// are clamped onto the border cell, as the host-side action table drops them.
__device__ __forceinline__ void decode_id(const StepParams &p, int id_raw, int &lvl, int &col, int &row) {
    const int N = p.X * p.Y;
    const int id = clampi(id_raw, 0, p.n_levels * N - 1);
    lvl = fdiv(id, N, p.inv_N);
    const int i = id - lvl * N;
    col = fdiv(i, p.X, p.inv_X);
    row = i - col * p.X;
    if (id != id_raw) *(volatile int *)p.status = 2;
    col = min(col, p.X - 1);
    row = min(row, p.Y - 1);
}

__device__ __forceinline__ void clip_footprint(const StepParams &p, int cx, int cy, int rx, int ry, Geom &g) {
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
}

__device__ __forceinline__ Geom geom_from_cell(const StepParams &p, int lvl, int col, int row) {
    Geom g;
    const AltLevel &L = p.lut[lvl];
    g.rf = L.rf;
    g.s2 = L.s2;
    g.R = L.R;
    // pose = res*idx + res/2 (actions.py:80-83), same operation order as NumPy
    g.px = __dadd_rn(__dmul_rn(p.res, (double)col), __dmul_rn(0.5, p.res));
    g.py = __dadd_rn(__dmul_rn(p.res, (double)row), __dmul_rn(0.5, p.res));
    g.ph = L.alt;
    clip_footprint(p, col, row, L.rx, L.ry, g);
    return g;
}

__device__ __forceinline__ Geom decode(const StepParams &p, int job) {
    if (p.action_ids != nullptr) {
        int lvl, col, row;
        decode_id(p, __ldg(p.action_ids + job), lvl, col, row);
        return geom_from_cell(p, lvl, col, row);
    }
    Geom g;
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
    const int cx = (int)fmin(fmax(fcx, -lim), lim);
    const int cy = (int)fmin(fmax(fcy, -lim), lim);
    const int rx = (int)fmin(fmax(frx, 0.0), lim);
    const int ry = (int)fmin(fmax(fry, 0.0), lim);
    g.rf = g.ph > p.rf_alt ? 2 : 1;
    const double s2 = p.coeff_a * (1.0 - exp(-p.coeff_b * g.ph));
    g.s2 = (float)s2;
    g.R = (float)((double)(g.rf * g.rf * g.rf) * s2);
    clip_footprint(p, cx, cy, rx, ry, g);
    return g;
}

// planning/common/actions.py:15-16 / 32-41.  The pose difference is formed in fp64; the norm and the
// trapezoidal-profile time are fp32 (relative error ~1e-7, two orders below the parity tolerance).
__device__ __forceinline__ float fast_sqrt(float x) { return x > 0.0f ? x * rsqrtf(x) : 0.0f; }

__device__ __forceinline__ float job_cost_from_dist(const StepParams &p, float d) {
    if (p.cost_mode == IPP_COST_DISTANCE) return d;
    const float d_acc = fminf(d * 0.5f, p.d_acc_max);  // min(d/2, v^2 / (2a))
    const float d_const = d - 2.0f * d_acc;
    return fmaf(d_const, p.inv_v, 2.0f * fast_sqrt(2.0f * d_acc * p.inv_a));
}
__device__ __forceinline__ float job_dist(double px, double py, double ph, double qx, double qy, double qh) {
    const float dx = (float)(px - qx), dy = (float)(py - qy), dz = (float)(ph - qh);
    return fast_sqrt(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
}
__device__ __forceinline__ float job_cost(const StepParams &p, double px, double py, double ph, double qx, double qy, double qh) {
    return job_cost_from_dist(p, job_dist(px, py, ph, qx, qy, qh));
}

// One axis of cv2 INTER_AREA decimation: output sample o of n_out integrates the input over
// [o*s, (o+1)*s), s = n_in/n_out.  Exact integer overlaps in units of 1/n_out; weight =
// overlap / n_in.  (opencv resize.cpp computeResizeAreaTab; reference call site
// simulations/sensor_manipulations.py:20-22.)
// A table entry is 6 floats = 3 x float2: {first input index, w0}, {w1, w2}, {w3, w4}.  Square
// footprints need 3 taps (scale < 2); clipped, non-square ones up to 5 (scale < 4); anything beyond
// (exotic FoV / grid shapes) takes the generic loop.
constexpr int TAPS_FAST = 0, TAPS_WIDE = 1, TAPS_GENERIC = 2;

__device__ __forceinline__ int make_tap_entry(float2 *e /* [3] */, int o, int n_in, int n_out) {
    const int a1 = o * n_in, a2 = a1 + n_in;
    const int start = a1 / n_out;
    const int end = (a2 + n_out - 1) / n_out;  // exclusive
    const float inv = 1.0f / (float)n_in;
    float w[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int i = start + k;
        const int lo = max(a1, i * n_out), hi = min(a2, (i + 1) * n_out);
        w[k] = hi > lo ? (float)(hi - lo) * inv : 0.0f;
    }
    e[0] = make_float2(__int_as_float(start), w[0]);
    e[1] = make_float2(w[1], w[2]);
    e[2] = make_float2(w[3], w[4]);
    const int cnt = end - start;
    return cnt <= 3 ? TAPS_FAST : (cnt <= 5 ? TAPS_WIDE : TAPS_GENERIC);
}

// generic weight of input i for output o (slow path)
__device__ __forceinline__ float tap_weight_generic(int o, int i, int n_in, int n_out, float inv_n_in) {
    const int a1 = o * n_in, a2 = a1 + n_in;
    const int lo = max(a1, i * n_out), hi = min(a2, (i + 1) * n_out);
    return hi > lo ? (float)(hi - lo) * inv_n_in : 0.0f;
}

// Build the row / column tap tables of a rf=2 footprint (one warp): tab[3*idx .. 3*idx+2] for rows,
// tab[3*(CAP+idx) ..] for columns.  Returns TAPS_FAST / TAPS_WIDE / TAPS_GENERIC for the footprint.
template <int CAP>
__device__ __forceinline__ int build_tap_tables(float2 *tab /* [2*CAP*3] */, int lane, int ny, int nx, int out_r, int out_c) {
    int mode = (out_r > CAP || out_c > CAP) ? TAPS_GENERIC : TAPS_FAST;
    if (mode != TAPS_GENERIC) {
        for (int idx = lane; idx < out_r + out_c; idx += 32) {
            const bool is_row = idx < out_r;
            float2 *e = tab + 3 * (is_row ? idx : CAP + idx - out_r);
            mode = max(mode, is_row ? make_tap_entry(e, idx, ny, out_r) : make_tap_entry(e, idx - out_r, nx, out_c));
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mode = max(mode, __shfl_xor_sync(0xffffffffu, mode, s));
    __syncwarp();
    return mode;
}

// D[pr, pc] of the down-sampled measurement from a ground-truth tile `g` (row pitch `pitch`, origin =
// footprint corner; global memory or shared memory).  Rows / columns past the footprint carry weight 0
// and are clamped.
struct TapView {
    const float2 *rows, *cols;  // entry k at [3k .. 3k+2]
};

// Ground-truth views: at(r, c) = value at row r, column c of the footprint (origin = its top-left cell).
struct GtShared {  // staged tile in shared memory, row pitch `pitch` floats
    const float *g;
    int pitch;
    __device__ __forceinline__ float at(int r, int c) const { return g[r * pitch + c]; }
};
struct GtRowMajor {  // global memory, row-major map; g points at the footprint origin
    const float *g;
    int pitch;
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(g + r * pitch + c); }
};
struct GtTiled {  // global memory, IPP_LAYOUT_TILED; g points at the env's plane
    const float *g;
    int txg, yu, xl;
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(g + tiled_gt_index(txg, yu + r, xl + c)); }
};

struct GtSuper {  // global memory, IPP_LAYOUT_SUPER; g points at the env's plane (+ 32 floats)
    const float *g;
    int tx, yu, xl;
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(g + super_gt_index(tx, yu + r, xl + c)); }
};

struct GtSplit {  // global memory, IPP_LAYOUT_SPLIT; g points at the env's {mean | gt} array + 16 floats
    const float *g;
    int tx, yu, xl;
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(g + split_mean_of(split_index(tx, yu + r, xl + c))); }
};

template <class G>
__device__ __forceinline__ float downsample_fast(const G &g, const TapView &t, int pr, int pc, int ny, int nx) {
    const float2 r0 = t.rows[3 * pr], r1 = t.rows[3 * pr + 1];
    const float2 c0 = t.cols[3 * pc], c1 = t.cols[3 * pc + 1];
    const int rs = __float_as_int(r0.x), cs = __float_as_int(c0.x);
    const float wr[3] = {r0.y, r1.x, r1.y};
    const int cb[3] = {cs, min(cs + 1, nx - 1), min(cs + 2, nx - 1)};
    float d = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int rr = min(rs + a, ny - 1);
        const float rowsum = fmaf(c1.y, g.at(rr, cb[2]), fmaf(c1.x, g.at(rr, cb[1]), c0.y * g.at(rr, cb[0])));
        d = fmaf(wr[a], rowsum, d);
    }
    return d;
}

// clipped non-square footprints: up to 5 taps per axis (kept out of line: ~15 % of the rf=2 envs)
template <class G>
__device__ __noinline__ float downsample_wide(G g, TapView t, int pr, int pc, int ny, int nx) {
    const float2 r0 = t.rows[3 * pr], r1 = t.rows[3 * pr + 1], r2 = t.rows[3 * pr + 2];
    const float2 c0 = t.cols[3 * pc], c1 = t.cols[3 * pc + 1], c2 = t.cols[3 * pc + 2];
    const int rs = __float_as_int(r0.x), cs = __float_as_int(c0.x);
    const float wr[5] = {r0.y, r1.x, r1.y, r2.x, r2.y};
    const float wc[5] = {c0.y, c1.x, c1.y, c2.x, c2.y};
    float d = 0.0f;
#pragma unroll 1
    for (int a = 0; a < 5; ++a) {
        const int rr = min(rs + a, ny - 1);
        float rowsum = 0.0f;
#pragma unroll
        for (int b = 0; b < 5; ++b) rowsum = fmaf(wc[b], g.at(rr, min(cs + b, nx - 1)), rowsum);
        d = fmaf(wr[a], rowsum, d);
    }
    return d;
}

// anything else (decimation scale >= 4 or footprints wider than the tap tables)
template <class G>
__device__ __noinline__ float downsample_generic(G g, int pr, int pc, int ny, int nx, int out_r, int out_c) {
    const int rs = (pr * ny) / out_r, re = ((pr + 1) * ny + out_r - 1) / out_r;
    const int cs = (pc * nx) / out_c, ce = ((pc + 1) * nx + out_c - 1) / out_c;
    const float inv_ny = 1.0f / (float)ny, inv_nx = 1.0f / (float)nx;
    float d = 0.0f;
#pragma unroll 1
    for (int a = rs; a < re; ++a) {
        const int rr = min(a, ny - 1);
        float rowsum = 0.0f;
#pragma unroll 1
        for (int b = cs; b < ce; ++b) rowsum = fmaf(tap_weight_generic(pc, b, nx, out_c, inv_nx), g.at(rr, min(b, nx - 1)), rowsum);
        d = fmaf(tap_weight_generic(pr, a, ny, out_r, inv_ny), rowsum, d);
    }
    return d;
}

template <class G>
__device__ __forceinline__ float downsample(int mode, const G &g, const TapView &t, int pr, int pc, int ny, int nx, int out_r, int out_c) {
    if (mode == TAPS_FAST) return downsample_fast(g, t, pr, pc, ny, nx);
    if (mode == TAPS_WIDE) return downsample_wide(g, t, pr, pc, ny, nx);
    return downsample_generic(g, pr, pc, ny, nx, out_r, out_c);
}

// Shannon entropy [nats] of Bernoulli(sigmoid(l)):  log1p(e^-|l|) + |l| e^-|l| / (1 + e^-|l|)
__device__ __forceinline__ float bernoulli_entropy(float l) {
    const float a = fabsf(l);
    const float e = __expf(-a);
    return log1pf(e) + a * e * __frcp_rn(1.0f + e);
}

// ---------------------------------------------------------------------------------------------
// Per-quad fusion.  A quad = 2x2 cells {(r0,c0), (r0,c0+1), (r0+1,c0), (r0+1,c0+1)}; cells outside
// the footprint carry m = v = 0 and ok = false.  rf = 1: four scalar Kalman updates; rf = 2: one
// block update with the reference's weight 1/4 (full block) or 1/2 (partial block).
// Returns the quad's information gain (trace reduction, or 0.5*ln prod v/v' for the entropy mode)
// over the cells selected by msk[].
// ---------------------------------------------------------------------------------------------
struct FuseCtx {
    int rf;
    float R, invR;
};

// SFU reciprocal / log2 (1 instruction each, ~1 ulp): two orders below the 1e-5 parity tolerance
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ENTROPY and ADAPTIVE are compile-time: with run-time flags both reward variants are issued predicated-off
// (ncu: ~50 wasted warp-instructions per env).
template <bool ENTROPY, bool ADAPTIVE>
__device__ __forceinline__ float kalman_quad(const FuseCtx &c, bool cok, bool rok, const float (&m)[4], const float (&v)[4],
                                             const float (&z)[4], const bool (&msk)[4], float (&mn)[4], float (&vn)[4]) {
    constexpr float kHalfLn2 = 0.34657359f;  // 0.5 * ln 2
    float gain_q = 0.0f;
    if (c.rf == 1) {
        float prod = 1.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float S = v[k] + c.R;
            const float gain = v[k] * fast_rcp(S);
            vn[k] = gain * c.R;  // v R / (v + R)  ==  v - v^2/S, cancellation-free
            mn[k] = fmaf(gain, z[k] - m[k], m[k]);
            const bool in = ADAPTIVE ? msk[k] : true;  // cells outside the footprint have v = 0: S/R = 1, v*gain = 0
            if (ENTROPY)
                prod *= in ? S * c.invR : 1.0f;  // v/v' = S/R
            else
                gain_q += in ? __fmul_rn(v[k], gain) : 0.0f;  // __fmul_rn: no context-dependent fma contraction
        }
        if (ENTROPY) gain_q = __fmul_rn(kHalfLn2, fast_lg2(prod));  // __fmul_rn: the caller's "acc +=" must not fuse
    } else {
        const float w = (cok && rok) ? 0.25f : 0.5f;  // sensor_models.py:76-79: partial blocks weigh 1/rf
        const float w2 = w * w;
        const float sv = (v[0] + v[1]) + (v[2] + v[3]);
        const float sm = (m[0] + m[1]) + (m[2] + m[3]);
        const float S = fmaf(w2, sv, c.R);
        const float invS = fast_rcp(S);
        const float innov = fmaf(-w, sm, z[0]);
        float rest[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            rest[k] = fmaf(-w2, v[k], S);  // w^2 * sum_{j != k} v_j + R  > 0
            const float vk_invS = v[k] * invS;
            vn[k] = vk_invS * rest[k];
            mn[k] = fmaf(w * vk_invS, innov, m[k]);
            if (!ENTROPY) gain_q += (ADAPTIVE ? msk[k] : true) ? __fmul_rn(w2 * v[k], vk_invS) : 0.0f;
        }
        if (ENTROPY) {
            // 0.5 ln prod_k v_k / v'_k = 0.5 ln( S^4 / prod_k f_k ),  f_k = rest_k for the masked cells, else S
            // (cells outside the footprint have v = 0, i.e. rest = S): two independent SFU logs, no division
            const float a = ((ADAPTIVE && !msk[0]) ? S : rest[0]) * ((ADAPTIVE && !msk[1]) ? S : rest[1]);
            const float b = ((ADAPTIVE && !msk[2]) ? S : rest[2]) * ((ADAPTIVE && !msk[3]) ? S : rest[3]);
            gain_q = __fmul_rn(kHalfLn2, fmaf(4.0f, fast_lg2(S), -fast_lg2(a * b)));
        }
    }
    return gain_q;
}

// run-time dispatch for callers that are not specialised themselves (the general kernel)
__device__ __forceinline__ float kalman_quad_rt(bool entropy, bool adaptive, const FuseCtx &c, bool cok, bool rok, const float (&m)[4],
                                                const float (&v)[4], const float (&z)[4], const bool (&msk)[4], float (&mn)[4],
                                                float (&vn)[4]) {
    if (entropy) {
        return adaptive ? kalman_quad<true, true>(c, cok, rok, m, v, z, msk, mn, vn) : kalman_quad<true, false>(c, cok, rok, m, v, z, msk, mn, vn);
    }
    return adaptive ? kalman_quad<false, true>(c, cok, rok, m, v, z, msk, mn, vn) : kalman_quad<false, false>(c, cok, rok, m, v, z, msk, mn, vn);
}

}  // namespace ipp
