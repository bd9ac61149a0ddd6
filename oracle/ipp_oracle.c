/*
 * ipp_oracle.c — plain-C, float64 restatement of the reference's per-step hot path
 * (TEST INFRASTRUCTURE / CPU BASELINE — NOT PRODUCT CODE; never linked into libipp_b200.so).
 *
 * Same algorithm as oracle/ipp_oracle.py (which is pinned against the real reference by
 * tests/golden/make_golden.py), written as scalar loops with an OpenMP loop over envs so that it can
 * serve as the all-host-cores CPU baseline in bench.py.  tests/test_oracle_c.py checks it against the
 * NumPy oracle and the golden vectors.
 *
 * Reference lines followed (paths under the reference tree):
 *   footprint                sensors/cameras.py:34-75
 *   resolution factor        sensors/cameras.py:122-125
 *   sigma2(h), R             sensors/models/sensor_models.py:27-36
 *   measurement blocks       sensors/models/sensor_models.py:54-81
 *   take_measurement         simulations/simulations.py:26-34, simulations/sensor_manipulations.py:7-57
 *   INTER_AREA               opencv resize.cpp computeResizeAreaTab (third-party; call site
 *                            sensor_manipulations.py:20-22, dsize swap included)
 *   Kalman update (diag)     mapping/mappings.py:155-197
 *   adaptive mask, reward    planning/common/rewards.py:8-31
 *   cost                     planning/common/actions.py:8-41
 *   Philox4x32-10/Box-Muller the engine's throughput-mode noise definition (oracle/ipp_oracle.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct orc_cfg {
    int32_t x_dim, y_dim;
    int32_t cost_mode; /* 0 distance, 1 flight time */
    int32_t pad;
    double res, tan_x, tan_y, coeff_a, coeff_b, rf_alt, max_v, max_a, thr, kappa;
} orc_cfg;

#define ORC_REWARD_ENTROPY 1
#define ORC_FLAG_ADAPTIVE 4
#define ORC_FLAG_NO_DSIZE_QUIRK 8
#define ORC_FLAG_PREDICT_ONLY 256

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int clampi(long long v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }

static void footprint(const orc_cfg *c, const double *pose, int *xl, int *xr, int *yu, int *yd) {
    const double xm = 2 * pose[2] * c->tan_x, ym = 2 * pose[2] * c->tan_y;
    const double wx = floor(xm / c->res), wy = floor(ym / c->res);
    const double cx = floor(pose[0] / c->res), cy = floor(pose[1] / c->res);
    const double rx = floor(0.5 * wx), ry = floor(0.5 * wy);
    *xl = clampi((long long)(cx - rx), 0, c->x_dim - 1);
    *xr = clampi((long long)(cx + rx), 0, c->x_dim - 1);
    *yu = clampi((long long)(cy - ry), 0, c->y_dim - 1);
    *yd = clampi((long long)(cy + ry), 0, c->y_dim - 1);
}

static double job_cost(const orc_cfg *c, const double *a, const double *p) {
    const double dx = a[0] - p[0], dy = a[1] - p[1], dz = a[2] - p[2];
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    if (c->cost_mode == 0) return d;
    const double d_acc = fmin(d * 0.5, (c->max_v * c->max_v) / (2 * c->max_a));
    const double d_const = d - 2 * d_acc;
    return d_const / c->max_v + 2 * sqrt(2 * d_acc / c->max_a);
}

/* one axis of cv2 INTER_AREA: weights of output sample d over inputs [*first, *first + n) */
static int area_taps(int n_in, int n_out, int d, int *first, double *w /* >= n_in */) {
    if (n_in % n_out == 0) {
        const int k = n_in / n_out;
        *first = d * k;
        for (int i = 0; i < k; ++i) w[i] = 1.0 / k;
        return k;
    }
    const double scale = (double)n_in / n_out;
    const double f1 = d * scale, f2 = f1 + scale;
    const double cell = fmin(scale, n_in - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    if (s2 > n_in - 1) s2 = n_in - 1;
    if (s1 > s2) s1 = s2;
    int n = 0, start = s1;
    if (s1 - f1 > 1e-3) {
        start = s1 - 1;
        w[n++] = (double)(float)((s1 - f1) / cell);
    }
    for (int s = s1; s < s2; ++s) w[n++] = (double)(float)(1.0 / cell);
    if (f2 - s2 > 1e-3) w[n++] = (double)(float)(fmin(fmin(f2 - s2, 1.0), cell) / cell);
    *first = start;
    return n;
}

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = (uint32_t)p1;
        c[2] = n2;
        c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

static void device_normals(uint64_t seed, uint32_t env, uint64_t step, uint32_t group, double out[4]) {
    uint32_t c[4] = {group, env, (uint32_t)(step & 0xffffffffu), (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
    const double k = 1.0 / 4294967296.0, pi = 3.14159265358979323846;
    const double u0 = (c[0] + 0.5) * k, u1 = (c[1] + 0.5) * k, u2 = (c[2] + 0.5) * k, u3 = (c[3] + 0.5) * k;
    const double r0 = sqrt(-2.0 * log(u0)), r1 = sqrt(-2.0 * log(u2));
    out[0] = r0 * cos(pi * (2.0 * u1 - 1.0));
    out[1] = r0 * sin(pi * (2.0 * u1 - 1.0));
    out[2] = r1 * cos(pi * (2.0 * u3 - 1.0));
    out[3] = r1 * sin(pi * (2.0 * u3 - 1.0));
}

/*
 * One step for B envs (fp64 state, in place).  gt/mean/var: [B][Y][X]; prev, actions: [B][3];
 * eps: NULL -> Philox (seed, env_offset + b, step) else [B][stride] standard normals in the C order of
 * the measurement array; z_out (nullable): [B][stride].  reward: [B].
 * flags: bit0 entropy reward, ORC_FLAG_ADAPTIVE, ORC_FLAG_NO_DSIZE_QUIRK, ORC_FLAG_PREDICT_ONLY
 * (covariance only: gt/eps unused, mean untouched).  Returns 0, or -4 on an up-sampling footprint.
 */
int orc_step(const orc_cfg *c, int B, const double *gt, double *mean, double *var, double *prev, const double *actions,
             const double *eps, int stride, uint64_t seed, int64_t env_offset, uint64_t step, int flags, double *reward,
             double *z_out) {
    const int X = c->x_dim, Y = c->y_dim;
    const size_t plane = (size_t)X * Y;
    const int entropy = (flags & 3) == ORC_REWARD_ENTROPY;
    const int adaptive = (flags & ORC_FLAG_ADAPTIVE) != 0;
    const int quirk = (flags & ORC_FLAG_NO_DSIZE_QUIRK) == 0;
    const int predict = (flags & ORC_FLAG_PREDICT_ONLY) != 0;
    int status = 0;

#pragma omp parallel for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
        const double *a = actions + 3 * (size_t)b;
        const double *g = gt ? gt + (size_t)b * plane : NULL;
        double *m = mean + (size_t)b * plane, *v = var + (size_t)b * plane;
        int xl, xr, yu, yd;
        footprint(c, a, &xl, &xr, &yu, &yd);
        const int nx = xr - xl + 1, ny = yd - yu + 1;
        const int rf = a[2] > c->rf_alt ? 2 : 1;
        const double s2 = c->coeff_a * (1 - exp(-c->coeff_b * a[2]));
        const double R = (double)(rf * rf * rf) * s2;
        const int nbx = (nx + rf - 1) / rf, nby = (ny + rf - 1) / rf;
        const int M = nbx * nby;
        double *z = NULL;

        if (!predict) {
            /* ---- measurement: crop -> INTER_AREA (dsize swap) -> + sigma2*eps -> clip ------------ */
            z = (double *)malloc(sizeof(double) * (size_t)(rf == 1 ? nx * ny : M));
            int out_r = ny, out_c = nx;
            if (rf == 2) {
                out_r = quirk ? nbx : nby;
                out_c = quirk ? nby : nbx;
                if (out_r > ny || out_c > nx) {
#pragma omp atomic write
                    status = -4;
                    free(z);
                    continue;
                }
                double wr[64], wc[64];
                double *wrow = ny > 64 ? (double *)malloc(sizeof(double) * ny) : wr;
                double *wcol = nx > 64 ? (double *)malloc(sizeof(double) * nx) : wc;
                for (int pr = 0; pr < out_r; ++pr) {
                    int r0;
                    const int nr = area_taps(ny, out_r, pr, &r0, wrow);
                    for (int pc = 0; pc < out_c; ++pc) {
                        int c0;
                        const int nc = area_taps(nx, out_c, pc, &c0, wcol);
                        double d = 0.0;
                        for (int i = 0; i < nr; ++i) {
                            double rs = 0.0;
                            for (int j = 0; j < nc; ++j) rs += wcol[j] * g[(size_t)(yu + r0 + i) * X + xl + c0 + j];
                            d += wrow[i] * rs;
                        }
                        z[pr * out_c + pc] = d;
                    }
                }
                if (wrow != wr) free(wrow);
                if (wcol != wc) free(wcol);
            } else {
                for (int r = 0; r < ny; ++r)
                    for (int cc = 0; cc < nx; ++cc) z[r * nx + cc] = g[(size_t)(yu + r) * X + xl + cc];
            }
            const int nz = out_r * out_c;
            if (eps) {
                for (int i = 0; i < nz; ++i) z[i] += s2 * eps[(size_t)b * stride + i];
            } else {
                const int nqx = (nx + 1) / 2;
                if (rf == 1) {
                    for (int r = 0; r < ny; ++r)
                        for (int cc = 0; cc < nx; ++cc) {
                            double n4[4];
                            device_normals(seed, (uint32_t)(env_offset + b), step, (uint32_t)((r / 2) * nqx + cc / 2), n4);
                            z[r * nx + cc] += s2 * n4[2 * (r % 2) + (cc % 2)];
                        }
                } else {
                    for (int i = 0; i < nz; ++i) {
                        double n4[4];
                        device_normals(seed, (uint32_t)(env_offset + b), step, (uint32_t)((i & 31) + 32 * (i >> 7)), n4);
                        z[i] += s2 * n4[(i >> 5) & 3];
                    }
                }
            }
            for (int i = 0; i < nz; ++i) z[i] = fmin(fmax(z[i], 0.0), 1.0);
            if (z_out) memcpy(z_out + (size_t)b * stride, z, sizeof(double) * nz);
        }

        /* ---- block Kalman update + reward ----------------------------------------------------------- */
        double acc = 0.0;
        for (int i = 0; i < M; ++i) {
            const int by = i / nbx, bx = i - nbx * by;
            const int x0 = bx * rf, y0 = by * rf;
            const int x1 = x0 + rf < nx ? x0 + rf : nx, y1 = y0 + rf < ny ? y0 + rf : ny;
            const int cnt = (x1 - x0) * (y1 - y0);
            const double w = cnt >= rf * rf ? 1.0 / (rf * rf) : 1.0 / rf;
            double sv = 0.0, sm = 0.0;
            for (int r = y0; r < y1; ++r)
                for (int cc = x0; cc < x1; ++cc) {
                    const size_t k = (size_t)(yu + r) * X + xl + cc;
                    sv += v[k];
                    sm += m[k];
                }
            const double S = w * w * sv + R;
            const double innov = predict ? 0.0 : z[i] - w * sm;
            for (int r = y0; r < y1; ++r)
                for (int cc = x0; cc < x1; ++cc) {
                    const size_t k = (size_t)(yu + r) * X + xl + cc;
                    const double vk = v[k], mk = m[k];
                    const double vn = vk - (w * vk) * (w * vk) / S;
                    const int in_mask = !adaptive || (mk + c->kappa * vk >= c->thr);
                    if (in_mask) acc += entropy ? 0.5 * log(vk / vn) : (vk - vn);
                    v[k] = vn;
                    if (!predict) m[k] = mk + (w * vk / S) * innov;
                }
        }
        const double cost = job_cost(c, a, prev + 3 * (size_t)b);
        reward[b] = acc / (cost + 1.0);
        prev[3 * (size_t)b + 0] = a[0];
        prev[3 * (size_t)b + 1] = a[1];
        prev[3 * (size_t)b + 2] = a[2];
        free(z);
    }
    return status;
}
