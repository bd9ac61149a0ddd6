// ipp_engine.cu — C-ABI implementation of include/ipp_b200.h for sm_100a.
//
// Host-side runtime of the batched IPP environment engine: HBM layout, reset / ground-truth /
// state transfer, the fused step launch (step_kernel.cuh), the evaluation-metric reduction and the
// error / stream plumbing.  No PyTorch, no Python: plain CUDA runtime behind extern "C".
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include <cstdlib>

#include "rollout_kernel.cuh"
#include "step_async.cuh"
#include "step_bulk.cuh"
#include "engine_internal.h"

using namespace ipp;

// ------------------------------------------------------------------------------------------------
// engine object
// ------------------------------------------------------------------------------------------------
struct ipp_engine {
    ipp_config cfg{};
    int n_levels = 0;
    AltLevel lut[IPP_MAX_ALTITUDE_LEVELS]{};
    int max_meas = 0;
    int sm_count = 0;
    size_t plane = 0;     // y_dim * x_dim (dense maps of the ABI)
    size_t plane_mv = 0;  // cells per env in the belief arrays (TILED: whole 4x4 tiles)
    size_t plane_gt = 0;  // cells per env in the ground-truth array (TILED: whole 8x4 tiles)
    int txm = 0, txg = 0, tiles_y = 0;
    int ts_mv = 0, ts_gt = 0, gw_shift = 0;  // tile strides of the belief [float2] / ground truth [float], log2 gt tile width
    bool tiled() const { return cfg.layout == IPP_LAYOUT_TILED || cfg.layout == IPP_LAYOUT_SUPER || cfg.layout == IPP_LAYOUT_SPLIT; }
    bool split() const { return cfg.layout == IPP_LAYOUT_SPLIT; }
    // HBM
    float *d_mean = nullptr;  // PLANES: float[B*plane]; MV / TILED / SUPER: float2[B*plane_mv]; SPLIT: {mean x 16 | gt x 16}[B*tiles]
    float *d_var = nullptr;   // PLANES: float[B*plane]; SPLIT: float[B*tiles*16]
    float *d_gt = nullptr;
    bool gt_aliases_mean = false;  // SUPER / SPLIT: d_gt points into d_mean's allocation
    double *d_prev = nullptr;  // [B][3]
    int *d_status = nullptr;  // device alias of h_status (mapped pinned host word: no copy needed to read it back)
    // staging for the host entry points
    int32_t *d_actions = nullptr;  // [cap_jobs]
    double *d_poses = nullptr;     // [cap_jobs][3]
    double *d_prev_in = nullptr;   // [cap_jobs][3]
    int32_t *d_env_index = nullptr;
    float *d_reward = nullptr;  // [cap_jobs]
    size_t cap_jobs = 0;
    float *d_noise = nullptr;
    float *d_z = nullptr;
    size_t cap_noise = 0, cap_z = 0;
    float *d_metrics = nullptr;
    float *d_scratch = nullptr;  // dense [n][plane] staging for MV get/set
    size_t cap_scratch = 0;
    int *h_status = nullptr;  // pinned + mapped
    // pipelined host steps (ipp_step_submit / ipp_step_wait): H2D of the next step's ids on a copy stream under the running kernel
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[IPP_STEP_SLOTS] = {nullptr, nullptr}, ev_done[IPP_STEP_SLOTS] = {nullptr, nullptr};
    int32_t *d_actions_slot[IPP_STEP_SLOTS] = {nullptr, nullptr};
    float *d_reward_slot[IPP_STEP_SLOTS] = {nullptr, nullptr};
    bool slot_busy[IPP_STEP_SLOTS] = {false, false};
    int zero_copy = IPP_ZERO_COPY_REWARDS | IPP_ZERO_COPY_IDS_FETCH;  // IPP_OPT_ZERO_COPY / env IPP_ZERO_COPY
    uint64_t zero_copy_steps = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // cp.async-staged persistent path (step_async.cuh) and the path switch
    bool async_ok = false;
    int step_path = IPP_PATH_ASYNC;  // requested path (ipp_set_option / IPP_STEP_PATH)
    int async_warps = 0, async_double_warps = 0, async_mv_tile = 0, async_gt_tile = 0;
    bool async_vec16 = false;
    float2 *d_level_taps = nullptr;
    int level_tap_mode[kLevelTabs] = {-1, -1, -1, -1};
    size_t async_smem = 0;
    uint64_t path_launches[2] = {0, 0};
    // bulk-copy persistent path (step_bulk.cuh, IPP_LAYOUT_SUPER)
    bool bulk_ok = false;
    int bulk_warps = 0, bulk_ring = 0;
    size_t bulk_smem = 0;
    int bulk_pwarps = 0, bulk_pring = 0;  // SPLIT: the covariance-only step (variance runs alone, more warps per CTA)
    size_t bulk_psmem = 0;
    uint64_t bulk_predict_launches = 0;
    unsigned int *d_tickets = nullptr;
    int ticket_parity = 0;
    unsigned int *d_slice_state = nullptr;  // in-kernel fetch of host action ids (BulkParams::host_ids)
    unsigned int slice_epoch = 0;
    uint64_t ids_fetched_steps = 0;
    uint64_t launches = 0;
    uint64_t steps = 0;
    uint64_t device_bytes = 0;
    std::string err;
};

static thread_local std::string g_create_err;

static int fail(ipp_engine *e, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e)
        e->err = buf;
    else
        g_create_err = buf;
    return code;
}

#define CU(e, call)                                                                                  \
    do {                                                                                             \
        cudaError_t _s = (call);                                                                     \
        if (_s != cudaSuccess) {                                                                     \
            const int _code = (_s == cudaErrorMemoryAllocation) ? IPP_ERR_NOMEM : IPP_ERR_CUDA;      \
            return fail((e), _code, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_s), __FILE__, __LINE__); \
        }                                                                                            \
    } while (0)

template <typename T>
static int dev_alloc(ipp_engine *e, T **p, size_t n) {
    CU(e, cudaMalloc((void **)p, n * sizeof(T)));
    e->device_bytes += n * sizeof(T);
    return IPP_OK;
}

template <typename T>
static int ensure(ipp_engine *e, T **p, size_t *cap, size_t n) {
    if (*cap >= n && *p) return IPP_OK;
    if (*p) {
        CU(e, cudaStreamSynchronize(e->stream));
        CU(e, cudaFree(*p));
        e->device_bytes -= *cap * sizeof(T);
        *p = nullptr;
        *cap = 0;
    }
    int rc = dev_alloc(e, p, n);
    if (rc == IPP_OK) *cap = n;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// auxiliary kernels (full-map streaming passes; trivially HBM-bound, coalesced)
// ------------------------------------------------------------------------------------------------
__global__ void reset_kernel(float *mean, float *var, int layout, size_t plane, int batch, float prior_mean, float prior_var,
                             const float *prior_var_env) {
    const size_t total = plane * (size_t)batch;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float pv = prior_var_env ? prior_var_env[i / plane] : prior_var;
        if (layout == IPP_LAYOUT_SUPER && (i % 24) >= 16) continue;  // the ground-truth third of a super-tile
        if (layout == IPP_LAYOUT_SPLIT) {  // i runs over the var array; the mean sits at tile * 32 + cell of the {mean | gt} array
            mean[i + (i & ~(size_t)15)] = prior_mean;
            var[i] = pv;
        } else if (layout != IPP_LAYOUT_PLANES) {
            reinterpret_cast<float2 *>(mean)[i] = make_float2(prior_mean, pv);
        } else {
            mean[i] = prior_mean;
            var[i] = pv;
        }
    }
}

__global__ void fill_prev_kernel(double *prev, int batch, double x, double y, double h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) {
        prev[3 * (size_t)i + 0] = x;
        prev[3 * (size_t)i + 1] = y;
        prev[3 * (size_t)i + 2] = h;
    }
}

// dense [n][plane] <-> interleaved float2 (MV layout)
__global__ void mv_unpack_kernel(const float2 *mv, float *mean, float *var, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 t = mv[i];
        if (mean) mean[i] = t.x;
        if (var) var[i] = t.y;
    }
}
__global__ void mv_pack_kernel(float2 *mv, const float *mean, const float *var, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 t = mv[i];
        if (mean) t.x = mean[i];
        if (var) t.y = var[i];
        mv[i] = t;
    }
}

// dense [n][Y][X] <-> IPP_LAYOUT_TILED (quad_math.cuh): one thread per dense cell
struct TiledDims {
    int X, Y, txm, txg, ts_mv, ts_gt, gw_shift;
    size_t plane, plane_mv, plane_gt;
};
__global__ void tiled_unpack_kernel(const float2 *mv, float *mean, float *var, TiledDims d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t env = i / d.plane;
        const int c = (int)(i - env * d.plane), R = c / d.X, C = c - R * d.X;
        const float2 t = mv[env * d.plane_mv + tiled_mv_index_rt(d.txm, d.ts_mv, R, C)];
        if (mean) mean[i] = t.x;
        if (var) var[i] = t.y;
    }
}
__global__ void tiled_pack_kernel(float2 *mv, const float *mean, const float *var, TiledDims d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t env = i / d.plane;
        const int c = (int)(i - env * d.plane), R = c / d.X, C = c - R * d.X;
        float2 *o = mv + env * d.plane_mv + tiled_mv_index_rt(d.txm, d.ts_mv, R, C);
        float2 t = *o;
        if (mean) t.x = mean[i];
        if (var) t.y = var[i];
        *o = t;
    }
}
// dense [n][Y][X] <-> IPP_LAYOUT_SPLIT: var[tile][16], {mean[16] | gt[16]}[tile]
__global__ void split_unpack_kernel(const float *mg, const float *vr, float *mean, float *var, TiledDims d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t env = i / d.plane;
        const int c = (int)(i - env * d.plane), R = c / d.X, C = c - R * d.X;
        const int k = split_index(d.txm, R, C);
        if (mean) mean[i] = mg[env * d.plane_gt + split_mean_of(k)];
        if (var) var[i] = vr[env * d.plane_mv + k];
    }
}
__global__ void split_pack_kernel(float *mg, float *vr, const float *mean, const float *var, TiledDims d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t env = i / d.plane;
        const int c = (int)(i - env * d.plane), R = c / d.X, C = c - R * d.X;
        const int k = split_index(d.txm, R, C);
        if (mean) mg[env * d.plane_gt + split_mean_of(k)] = mean[i];
        if (var) vr[env * d.plane_mv + k] = var[i];
    }
}
// to_tiled != 0: dense -> tiled, else tiled -> dense
__global__ void tiled_gt_kernel(float *tiled, float *dense, TiledDims d, size_t n, int to_tiled) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t env = i / d.plane;
        const int c = (int)(i - env * d.plane), R = c / d.X, C = c - R * d.X;
        float *t = tiled + env * d.plane_gt + tiled_gt_index_rt(d.txg, d.ts_gt, d.gw_shift, R, C);
        if (to_tiled)
            *t = dense[i];
        else
            dense[i] = *t;
    }
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

// Smooth synthetic field in [0,1]: normalised sum of 6 random plane waves per env.
__global__ void synth_gt_kernel(float *gt, size_t plane, size_t plane_gt, int X, int txg /* > 0: TILED / SUPER */, int ts_gt, int gw_shift,
                                int batch, uint32_t seed, uint32_t env_off) {
    const int env = blockIdx.y;
    __shared__ float fx[6], fy[6], ph[6], am[6];
    if (threadIdx.x < 6) {
        const uint32_t h0 = mix32(seed ^ mix32((uint32_t)env + env_off) ^ (0x9E3779B9u * (threadIdx.x + 1)));
        const uint32_t h1 = mix32(h0 + 0x85ebca6bu), h2 = mix32(h1 + 0xc2b2ae35u);
        const float k = 0.02f + 0.10f * (float)threadIdx.x / 6.0f;
        const float ang = 6.2831853f * (h0 * 2.3283064e-10f);
        fx[threadIdx.x] = k * cosf(ang) * 6.2831853f;
        fy[threadIdx.x] = k * sinf(ang) * 6.2831853f;
        ph[threadIdx.x] = 6.2831853f * (h1 * 2.3283064e-10f);
        am[threadIdx.x] = (0.5f + 0.5f * (h2 * 2.3283064e-10f)) / (1.0f + threadIdx.x);
    }
    __syncthreads();
    float norm = 0.f;
    for (int k = 0; k < 6; ++k) norm += am[k];
    float *g = gt + (size_t)env * plane_gt;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
        const int C = (int)(i % X), R = (int)(i / X);
        const float x = (float)C, y = (float)R;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += am[k] * __sinf(fx[k] * x + fy[k] * y + ph[k]);
        g[txg > 0 ? tiled_gt_index_rt(txg, ts_gt, gw_shift, R, C) : i] = fminf(fmaxf(0.5f + 0.5f * s / norm, 0.0f), 1.0f);
    }
}

// Evaluation metrics: one CTA per env, two streaming passes, fp64 block reduction.
// planning/evaluation_metrics.py:4-58 as called by planning/missions.py:176-203.
constexpr int kEvalThreads = 256;

template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], double *smem, bool is_min_first2, bool is_max_third) {
    // v[0], v[1] reduced with min, v[2] with max when flagged; everything else summed
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double x = v[k];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, s);
            if (is_min_first2 && k < 2)
                x = fmin(x, y);
            else if (is_max_third && k == 2)
                x = fmax(x, y);
            else
                x += y;
        }
        if (lane == 0) smem[warp * N + k] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int nw = kEvalThreads / 32;
            double x = lane < nw ? smem[lane * N + k] : ((is_min_first2 && k < 2) ? 1e300 : ((is_max_third && k == 2) ? -1e300 : 0.0));
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const double y = __shfl_xor_sync(0xffffffffu, x, s);
                if (is_min_first2 && k < 2)
                    x = fmin(x, y);
                else if (is_max_third && k == 2)
                    x = fmax(x, y);
                else
                    x += y;
            }
            if (lane == 0) smem[k] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = smem[k];
    __syncthreads();
}

__global__ void __launch_bounds__(kEvalThreads) eval_kernel(const float *mean, const float *var, const float *gt, int layout, TiledDims d,
                                                            float thr, float *metrics) {
    __shared__ double smem[(kEvalThreads / 32) * 10];
    const int env = blockIdx.x;
    const size_t plane = d.plane;
    const bool split = layout == IPP_LAYOUT_SPLIT;
    const bool tiled = layout == IPP_LAYOUT_TILED || layout == IPP_LAYOUT_SUPER || split;
    const bool planar = layout == IPP_LAYOUT_PLANES || split;  // mean and var in separate arrays
    const float *g = gt + (size_t)env * d.plane_gt;
    const float *m = split ? mean + (size_t)env * d.plane_gt : mean + (size_t)env * d.plane_mv * (planar ? 1 : 2);
    const float *v = planar ? var + (size_t)env * d.plane_mv : m + 1;
    const int es = planar ? 1 : 2;
    // cell i of the dense map -> offsets inside the env's belief / ground-truth arrays (SPLIT: the compact tile index; the mean
    // array is addressed through mo())
    auto bi = [&](size_t i) -> size_t { return tiled ? tiled_mv_index_rt(d.txm, d.ts_mv, (int)(i / d.X), (int)(i % d.X)) : i; };
    auto mo = [&](size_t b) -> size_t { return split ? b + (b & ~(size_t)15) : b * es; };
    auto gi_of = [&](size_t i) -> size_t { return tiled ? tiled_gt_index_rt(d.txg, d.ts_gt, d.gw_shift, (int)(i / d.X), (int)(i % d.X)) : i; };

    // Rows of whole 4-cell groups (x_dim % 4 == 0): a thread takes four consecutive cells of a row — contiguous in every layout
    // (a tile row is four cells) — with 16-byte loads and no per-cell index arithmetic; any other width goes cell by cell.
    const bool vec = (d.X & 3) == 0;
    const int X4 = d.X >> 2, nvec = d.Y * X4;
    auto load4 = [&](int R, int C0, float (&gv)[4], float (&mvv)[4], float (&vv)[4]) {
        if (layout == IPP_LAYOUT_PLANES) {
            const size_t k = (size_t)R * d.X + C0;
            const float4 g4 = *reinterpret_cast<const float4 *>(g + k), m4 = *reinterpret_cast<const float4 *>(m + k),
                         v4 = *reinterpret_cast<const float4 *>(v + k);
            gv[0] = g4.x, gv[1] = g4.y, gv[2] = g4.z, gv[3] = g4.w;
            mvv[0] = m4.x, mvv[1] = m4.y, mvv[2] = m4.z, mvv[3] = m4.w;
            vv[0] = v4.x, vv[1] = v4.y, vv[2] = v4.z, vv[3] = v4.w;
            return;
        }
        const size_t kb = tiled ? tiled_mv_index_rt(d.txm, d.ts_mv, R, C0) : (size_t)R * d.X + C0;
        const size_t kg = tiled ? tiled_gt_index_rt(d.txg, d.ts_gt, d.gw_shift, R, C0) : (size_t)R * d.X + C0;
        const float4 g4 = *reinterpret_cast<const float4 *>(g + kg);
        gv[0] = g4.x, gv[1] = g4.y, gv[2] = g4.z, gv[3] = g4.w;
        if (split) {
            const float4 m4 = *reinterpret_cast<const float4 *>(m + mo(kb)), v4 = *reinterpret_cast<const float4 *>(v + kb);
            mvv[0] = m4.x, mvv[1] = m4.y, mvv[2] = m4.z, mvv[3] = m4.w;
            vv[0] = v4.x, vv[1] = v4.y, vv[2] = v4.z, vv[3] = v4.w;
        } else {  // interleaved {mean, var}
            const float4 t0 = *reinterpret_cast<const float4 *>(m + 2 * kb), t1 = *reinterpret_cast<const float4 *>(m + 2 * kb + 4);
            mvv[0] = t0.x, vv[0] = t0.y, mvv[1] = t0.z, vv[1] = t0.w, mvv[2] = t1.x, vv[2] = t1.y, mvv[3] = t1.z, vv[3] = t1.w;
        }
    };

    // pass 1: min(gt), min(mean), max(gt), sum(gt)
    double a[4] = {1e300, 1e300, -1e300, 0.0};
    if (vec) {
        for (int i = threadIdx.x; i < nvec; i += kEvalThreads) {
            const int R = i / X4, C0 = (i - R * X4) << 2;
            float gv[4], mvv[4], vv[4];
            load4(R, C0, gv, mvv, vv);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double gi = gv[c], mi = mvv[c];
                a[0] = fmin(a[0], gi);
                a[1] = fmin(a[1], mi);
                a[2] = fmax(a[2], gi);
                a[3] += gi;
            }
        }
    } else {
        for (size_t i = threadIdx.x; i < plane; i += kEvalThreads) {
            const double gi = g[gi_of(i)], mi = m[mo(bi(i))];
            a[0] = fmin(a[0], gi);
            a[1] = fmin(a[1], mi);
            a[2] = fmax(a[2], gi);
            a[3] += gi;
        }
    }
    block_reduce<4>(a, smem, true, true);
    const double gmin = a[0], mmin = a[1], gmax = a[2], gsum = a[3];
    const double range = gmax - gmin;
    const double n = (double)plane;
    const double wsum = (gsum - n * mmin) / range;  // sum of un-normalised weights (evaluation_metrics.py:34-35)

    // pass 2
    double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto accumulate = [&](double gi, double mi, double vi) {
        const double sq = (gi - mi) * (gi - mi);
        const double w = ((gi - mmin) / range) / wsum;
        const double ll = 0.5 * log(2.0 * 3.141592653589793 * vi) + sq / 2.0 * vi;  // (:44) multiplies by P_ii
        const bool in = (float)gi >= thr;
        s[0] += sq;
        s[1] += w * sq;
        s[2] += ll;
        s[3] += w * ll;
        s[4] += vi;
        s[5] += in ? vi : 0.0;
        s[6] += in ? 0.0 : vi;
        s[7] += in ? 1.0 : 0.0;
        s[8] += in ? sq : 0.0;
    };
    if (vec) {
        for (int i = threadIdx.x; i < nvec; i += kEvalThreads) {
            const int R = i / X4, C0 = (i - R * X4) << 2;
            float gv[4], mvv[4], vv[4];
            load4(R, C0, gv, mvv, vv);
#pragma unroll
            for (int c = 0; c < 4; ++c) accumulate((double)gv[c], (double)mvv[c], (double)vv[c]);
        }
    } else {
        for (size_t i = threadIdx.x; i < plane; i += kEvalThreads) {
            const size_t b = bi(i);
            accumulate((double)g[gi_of(i)], (double)m[mo(b)], (double)v[b * es]);
        }
    }
    block_reduce<10>(s, smem, false, false);
    if (threadIdx.x == 0) {
        float *o = metrics + (size_t)env * IPP_NUM_METRICS;
        const double n_in = s[7], n_out = n - s[7];
        o[0] = (float)sqrt(s[0] / n);
        o[1] = (float)sqrt(s[1] / n);
        o[2] = (float)(s[2] / n);
        o[3] = (float)(s[3] / n);
        o[4] = (float)s[4];
        const double mu_in = s[5] / n_in, mu_out = s[6] / n_out;
        o[5] = (float)((mu_out - mu_in) / mu_out);
        o[6] = (float)sqrt(s[8] / n_in);
        o[7] = (float)s[5];
    }
}

// ------------------------------------------------------------------------------------------------
// configuration -> altitude LUT (fp64, same operation order as the reference)
// ------------------------------------------------------------------------------------------------
static void footprint_radius(const ipp_config &c, double h, int *rx, int *ry) {
    // sensors/cameras.py:44-45, 62-66
    const double xm = 2 * h * c.tan_half_x, ym = 2 * h * c.tan_half_y;
    const double wx = std::floor(xm / c.resolution), wy = std::floor(ym / c.resolution);
    *rx = (int)std::floor(0.5 * wx);
    *ry = (int)std::floor(0.5 * wy);
}

static int build_lut(ipp_engine *e) {
    ipp_config &c = e->cfg;
    // planning/common/actions.py:74: linspace(min, max, int((max-min)/spacing)+1)
    const int n = (int)((c.max_altitude - c.min_altitude) / c.altitude_spacing) + 1;
    if (n < 1 || n > IPP_MAX_ALTITUDE_LEVELS) return fail(e, IPP_ERR_INVALID, "altitude levels %d outside [1,%d]", n, IPP_MAX_ALTITUDE_LEVELS);
    e->n_levels = n;
    int max_cells_rf1 = 1, max_blocks_rf2 = 1;
    for (int k = 0; k < n; ++k) {
        // numpy.linspace: start + k*step with step = (stop-start)/(n-1); last point forced to stop
        const double step = n > 1 ? (c.max_altitude - c.min_altitude) / (double)(n - 1) : 0.0;
        double h = c.min_altitude + (double)k * step;
        if (k == n - 1 && n > 1) h = c.max_altitude;
        AltLevel &L = e->lut[k];
        L.alt = h;
        footprint_radius(c, h, &L.rx, &L.ry);
        L.rf = h > c.rf_altitude ? 2 : 1;
        const double s2 = c.coeff_a * (1 - std::exp(-c.coeff_b * h));
        L.s2 = (float)s2;
        L.R = (float)((double)(L.rf * L.rf * L.rf) * s2);
        L.pad = 0;
    }
    // bound of measurements per step for any altitude in (0, max_altitude] (pose mode included)
    {
        int rx, ry;
        const double h1 = std::fmin(c.max_altitude, c.rf_altitude);
        footprint_radius(c, h1, &rx, &ry);
        const int nx = std::min(2 * rx + 1, c.x_dim), ny = std::min(2 * ry + 1, c.y_dim);
        max_cells_rf1 = nx * ny;
        footprint_radius(c, c.max_altitude, &rx, &ry);
        const int nx2 = std::min(2 * rx + 1, c.x_dim), ny2 = std::min(2 * ry + 1, c.y_dim);
        max_blocks_rf2 = ((nx2 + 1) / 2) * ((ny2 + 1) / 2);
        if (c.max_altitude <= c.rf_altitude) max_blocks_rf2 = 0;
    }
    e->max_meas = std::max(max_cells_rf1, max_blocks_rf2);
    return IPP_OK;
}

// ------------------------------------------------------------------------------------------------
// persistent-path set-up
// ------------------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

typedef void (*async_kernel_t)(const AsyncParams);
// bit 0: entropy reward, bit 1: adaptive mask, bit 2: extras (host noise / measurement read-back), bit 3: TILED layout
template <int V>
static async_kernel_t async_variant_t() {
    return ipp_step_async_kernel<(V & 1) != 0, (V & 2) != 0, (V & 4) != 0, (V & 8) != 0>;
}
static async_kernel_t async_variant(int v) {
    static const async_kernel_t table[16] = {async_variant_t<0>(),  async_variant_t<1>(),  async_variant_t<2>(),  async_variant_t<3>(),
                                             async_variant_t<4>(),  async_variant_t<5>(),  async_variant_t<6>(),  async_variant_t<7>(),
                                             async_variant_t<8>(),  async_variant_t<9>(),  async_variant_t<10>(), async_variant_t<11>(),
                                             async_variant_t<12>(), async_variant_t<13>(), async_variant_t<14>(), async_variant_t<15>()};
    return table[v & 15];
}

// cp.async-staged persistent path: needs the MV layout and footprints that fit two slots per warp.
static int setup_async(ipp_engine *e) {
    const ipp_config &c = e->cfg;
    e->async_ok = false;
    int rc;
    if ((rc = dev_alloc(e, &e->d_tickets, 2)) != IPP_OK) return rc;
    CU(e, cudaMemsetAsync(e->d_tickets, 0, 2 * sizeof(unsigned int), e->stream));
    if (c.layout != IPP_LAYOUT_MV && c.layout != IPP_LAYOUT_TILED) return IPP_OK;
    if (c.x_dim > 32767 || c.y_dim > 32767 || e->n_levels > 255) return IPP_OK;  // EnvPlan packs cell coordinates into 16 bits
    // tile capacities for the largest footprint; with 16-byte staging (x_dim % 4 == 0) the tiles hold the
    // aligned superset of each row: {mean,var} pitch = roundup(1 + fw, 2) cells, gt pitch = roundup(3 + fw, 4) floats
    e->async_vec16 = (c.x_dim % 4 == 0);
    const char *v16 = getenv("IPP_ASYNC_VEC16");
    if (v16 && v16[0] == '0') e->async_vec16 = false;
    if (c.layout == IPP_LAYOUT_TILED) e->async_vec16 = true;  // tiles are always staged as 16-byte aligned supersets
    int mv_cells = 0, gt_cells = 0;
    for (int k = 0; k < e->n_levels; ++k) {
        const int fw = std::min(2 * e->lut[k].rx + 1, c.x_dim), fh = std::min(2 * e->lut[k].ry + 1, c.y_dim);
        if (fw > 255 || fh > 255) return IPP_OK;  // EnvPlan packs footprint sizes into 8 bits
        const int pm = e->async_vec16 ? round_up(fw + 1, 2) : round_up(fw, 2);
        const int pg = e->async_vec16 ? round_up(fw + 3, 4) : round_up(fw, 2);
        mv_cells = std::max(mv_cells, pm * fh);
        gt_cells = std::max(gt_cells, pg * fh);
        // the kernel divides quad indices with a 16-bit magic multiplier: exact while quads * quads-per-row < 2^15
        const int nqx = (fw + 1) / 2, nqy = (fh + 1) / 2;
        if ((long long)nqx * nqy * std::max(nqx, nqy) >= 32768) return IPP_OK;
    }
    // TILED + IPP_DIRECT_MV: the belief is not staged at all (bulk L2 prefetch + direct loads), a slot holds the ground truth only
    const bool direct_mv = c.layout == IPP_LAYOUT_TILED && (IPP_DIRECT_MV != 0);
    const int mv_tile = direct_mv ? 0 : round_up(mv_cells * 8, 16), gt_tile = round_up(gt_cells * 4, 16);
    // shared memory = [slots] stage tiles | [warps] plan ring | [warps] per-env tap tables | level tap tables.  Every warp
    // owns one slot; what is left becomes second (prefetch) slots of the first `double_warps` warps.
    const size_t per_slot = (size_t)(mv_tile + gt_tile);
    const size_t per_warp_fixed = kTapFloats2 * sizeof(float2) + (kPlanRing + 1) * sizeof(EnvPlan);
    const size_t per_cta = (size_t)kLevelTabs * kTapFloats2 * sizeof(float2);
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device) != cudaSuccess) return IPP_OK;
    if ((size_t)dev_smem < per_cta + 4 * (2 * per_slot + per_warp_fixed)) return IPP_OK;  // footprints too large to pipeline in shared memory
    int warps = kAsyncMaxWarps;
    if (const char *wenv = getenv("IPP_ASYNC_WARPS")) warps = std::max(4, std::min(kAsyncMaxWarps, atoi(wenv)));
    warps = (int)std::min<size_t>(warps, ((size_t)dev_smem - per_cta) / (per_slot + per_warp_fixed));
    const size_t spare = (size_t)dev_smem - per_cta - (size_t)warps * (per_slot + per_warp_fixed);
    const int double_warps = (int)std::min<size_t>(warps, spare / per_slot);
    e->async_warps = warps;
    e->async_double_warps = double_warps;
    e->async_mv_tile = mv_tile;
    e->async_gt_tile = gt_tile;
    e->async_smem = per_slot * (warps + double_warps) + per_warp_fixed * warps + per_cta;
    for (int v = 0; v < 16; ++v)
        // the attribute is per function, not per engine: always ask for the device maximum, or a second engine with smaller
        // footprints would lower the limit under the first one
        if (cudaFuncSetAttribute(async_variant(v), cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem) != cudaSuccess) {
            cudaGetLastError();
            return IPP_OK;
        }
    // tap tables of the unclipped footprint of the first kLevelTabs levels (dsize-quirk orientation)
    int h_dims[4 * kLevelTabs] = {0};
    for (int k = 0; k < std::min(e->n_levels, kLevelTabs); ++k) {
        const int fw = 2 * e->lut[k].rx + 1, fh = 2 * e->lut[k].ry + 1;
        if (e->lut[k].rf != 2 || fw > c.x_dim || fh > c.y_dim) continue;  // dims stay 0 -> no table for this level
        h_dims[4 * k + 0] = fh;
        h_dims[4 * k + 1] = fw;
        h_dims[4 * k + 2] = (fw + 1) / 2;  // out_r = nqx (quirk)
        h_dims[4 * k + 3] = (fh + 1) / 2;  // out_c = nqy
    }
    int *d_dims = nullptr, *d_modes = nullptr;
    if ((rc = dev_alloc(e, &e->d_level_taps, (size_t)kLevelTabs * kTapFloats2)) != IPP_OK) return rc;
    CU(e, cudaMalloc((void **)&d_dims, sizeof h_dims));
    CU(e, cudaMalloc((void **)&d_modes, kLevelTabs * sizeof(int)));
    CU(e, cudaMemcpyAsync(d_dims, h_dims, sizeof h_dims, cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaMemsetAsync(e->d_level_taps, 0, (size_t)kLevelTabs * kTapFloats2 * sizeof(float2), e->stream));
    build_level_taps_kernel<<<kLevelTabs, 32, 0, e->stream>>>(e->d_level_taps, d_dims, d_modes);
    e->launches++;
    CU(e, cudaMemcpyAsync(e->level_tap_mode, d_modes, kLevelTabs * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    cudaFree(d_dims);
    cudaFree(d_modes);
    e->async_ok = true;
    return IPP_OK;
}

// ------------------------------------------------------------------------------------------------
// bulk-copy persistent path (IPP_LAYOUT_SUPER)
// ------------------------------------------------------------------------------------------------
typedef void (*bulk_kernel_t)(const BulkParams);
// bit 0: entropy reward, bit 1: adaptive mask, bit 2: extras (host noise / measurement read-back), bit 3: predict-only,
// bit 4: IPP_LAYOUT_SPLIT (else IPP_LAYOUT_SUPER)
template <int V>
static bulk_kernel_t bulk_variant_t() {
    return ipp_step_bulk_kernel<(V & 8) ? MODE_PREDICT : MODE_KALMAN, (V & 1) != 0, (V & 2) != 0, (V & 4) != 0 && (V & 8) == 0, (V & 16) != 0>;
}
static bulk_kernel_t bulk_variant(int v) {
    static const bulk_kernel_t table[32] = {
        bulk_variant_t<0>(),  bulk_variant_t<1>(),  bulk_variant_t<2>(),  bulk_variant_t<3>(),  bulk_variant_t<4>(),  bulk_variant_t<5>(),
        bulk_variant_t<6>(),  bulk_variant_t<7>(),  bulk_variant_t<8>(),  bulk_variant_t<9>(),  bulk_variant_t<10>(), bulk_variant_t<11>(),
        bulk_variant_t<8>(),  bulk_variant_t<9>(),  bulk_variant_t<10>(), bulk_variant_t<11>(), bulk_variant_t<16>(), bulk_variant_t<17>(),
        bulk_variant_t<18>(), bulk_variant_t<19>(), bulk_variant_t<20>(), bulk_variant_t<21>(), bulk_variant_t<22>(), bulk_variant_t<23>(),
        bulk_variant_t<24>(), bulk_variant_t<25>(), bulk_variant_t<26>(), bulk_variant_t<27>(), bulk_variant_t<24>(), bulk_variant_t<25>(),
        bulk_variant_t<26>(), bulk_variant_t<27>()};
    return table[v & 31];
}
// does this variant run the covariance-only configuration (variance runs alone)?
static bool bulk_variance_only(int v) { return (v & 16) && (v & 8) && !(v & 2); }

static int setup_bulk(ipp_engine *e) {
    const ipp_config &c = e->cfg;
    e->bulk_ok = false;
    const bool split = c.layout == IPP_LAYOUT_SPLIT;
    if (c.layout != IPP_LAYOUT_SUPER && !split) return IPP_OK;
    if (c.x_dim > 32767 || c.y_dim > 32767 || e->n_levels > 255) return IPP_OK;  // BulkPlan packs cell coordinates into 16 bits
    int max_fp = 0;  // bytes of the largest staged footprint at its worst tile alignment (192 bytes per tile in either layout)
    int max_tiles = 0;
    for (int k = 0; k < e->n_levels; ++k) {
        const int fw = std::min(2 * e->lut[k].rx + 1, c.x_dim), fh = std::min(2 * e->lut[k].ry + 1, c.y_dim);
        if (fw > 255 || fh > 255) return IPP_OK;  // BulkPlan packs footprint sizes into 8 bits
        const int ntx = std::min(e->txm, (fw + 2) / 4 + 1), ntr = std::min(e->tiles_y, (fh + 2) / 4 + 1);
        if (ntr > (split ? 16 : 32)) return IPP_OK;  // one lane per staged run
        max_tiles = std::max(max_tiles, ntx * ntr);
        max_fp = std::max(max_fp, ntx * ntr * kSuperTileBytes);
        const int nqx = (fw + 1) / 2, nqy = (fh + 1) / 2;
        if ((long long)nqx * nqy * std::max(nqx, nqy) >= 32768) return IPP_OK;  // 16-bit magic division of quad indices
    }
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device) != cudaSuccess) return IPP_OK;
    const size_t per_warp_fixed = kBulkPlanRing * sizeof(BulkPlan) + kBulkTapFloats2 * sizeof(float2) + kBulkDepth * 8;
    const size_t per_cta = (size_t)kBulkMaxLevels * 2 * sizeof(float4);
    const size_t avail = (size_t)dev_smem - per_cta;
    int warps = kBulkMaxWarps;
    if (const char *wenv = getenv("IPP_BULK_WARPS")) warps = std::max(1, std::min(kBulkMaxWarps, atoi(wenv)));
    while (warps > 1 && avail / warps < per_warp_fixed + (size_t)max_fp) --warps;
    if (avail / warps < per_warp_fixed + (size_t)max_fp) return IPP_OK;  // footprints too large to stage in shared memory
    int ring = (int)((avail / warps - per_warp_fixed) / 16 * 16);
    if (const char *renv = getenv("IPP_BULK_RING")) ring = std::max(max_fp, std::min(ring, atoi(renv) / 16 * 16));
    e->bulk_warps = warps;
    e->bulk_ring = ring;
    e->bulk_smem = (size_t)warps * ((size_t)ring + per_warp_fixed) + per_cta;
    if (split) {  // covariance-only step: 64 bytes per staged tile, up to kBulkPredictWarps warps
        const int max_fpv = max_tiles * kSplitVarTileBytes;
        int pw = kBulkPredictWarps;
        if (const char *wenv = getenv("IPP_BULK_PREDICT_WARPS")) pw = std::max(1, std::min(kBulkPredictWarps, atoi(wenv)));
        while (pw > 1 && avail / pw < per_warp_fixed + (size_t)max_fpv) --pw;
        int pring = (int)((avail / pw - per_warp_fixed) / 16 * 16);
        if (const char *renv = getenv("IPP_BULK_PREDICT_RING")) pring = std::max(max_fpv, std::min(pring, atoi(renv) / 16 * 16));
        e->bulk_pwarps = pw;
        e->bulk_pring = pring;
        e->bulk_psmem = (size_t)pw * ((size_t)pring + per_warp_fixed) + per_cta;
    }
    for (int v = 0; v < 32; ++v)
        if (cudaFuncSetAttribute(bulk_variant(v), cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem) != cudaSuccess) {  // per function, not per engine
            cudaGetLastError();
            return IPP_OK;
        }
    e->bulk_ok = true;
    return IPP_OK;
}

static int launch_bulk(ipp_engine *e, const StepParams &p, bool predict, const int32_t *host_ids = nullptr) {
    const int variant = (((p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY) ? 1 : 0) | ((p.flags & IPP_FLAG_ADAPTIVE) ? 2 : 0) |
                        ((!predict && (p.noise != nullptr || p.z_out != nullptr)) ? 4 : 0) | (predict ? 8 : 0) |
                        (e->cfg.layout == IPP_LAYOUT_SPLIT ? 16 : 0);
    const bool vonly = bulk_variance_only(variant);
    const int warps = vonly ? e->bulk_pwarps : e->bulk_warps;
    BulkParams bp;
    bp.base = p;
    bp.tickets = e->d_tickets;
    bp.parity = e->ticket_parity;
    bp.warps = warps;
    bp.ring_bytes = vonly ? e->bulk_pring : e->bulk_ring;
    bp.host_ids = nullptr;
    bp.slice_state = nullptr;
    bp.epoch = 0;
    if (host_ids) {
        const size_t n_slices = ((size_t)e->cfg.batch + kIdSlice - 1) / kIdSlice;
        if (!e->d_slice_state || e->slice_epoch > 0xfffffff0u) {
            int rc;
            if (!e->d_slice_state && (rc = dev_alloc(e, &e->d_slice_state, n_slices)) != IPP_OK) return rc;
            CU(e, cudaMemsetAsync(e->d_slice_state, 0, n_slices * sizeof(unsigned int), e->stream));
            e->slice_epoch = 0;
        }
        e->slice_epoch += 2;
        bp.host_ids = host_ids;
        bp.slice_state = e->d_slice_state;
        bp.epoch = e->slice_epoch;
        e->ids_fetched_steps++;
    }
    const int needed = (p.n_jobs + warps - 1) / warps;
    const int grid = std::max(1, std::min(e->sm_count, needed));
    bulk_variant(variant)<<<grid, warps * 32, vonly ? e->bulk_psmem : e->bulk_smem, e->stream>>>(bp);
    e->ticket_parity ^= 1;
    e->launches++;
    e->path_launches[IPP_PATH_ASYNC]++;
    if (predict) e->bulk_predict_launches++;
    CU(e, cudaGetLastError());
    return IPP_OK;
}

static int launch_async(ipp_engine *e, const StepParams &p) {
    AsyncParams ap;
    ap.base = p;
    ap.tickets = e->d_tickets;
    ap.level_taps = e->d_level_taps;
    memcpy(ap.level_tap_mode, e->level_tap_mode, sizeof ap.level_tap_mode);
    ap.parity = e->ticket_parity;
    ap.warps = e->async_warps;
    ap.double_warps = e->async_double_warps;
    ap.mv_tile_bytes = e->async_mv_tile;
    ap.gt_tile_bytes = e->async_gt_tile;
    ap.vec16 = e->async_vec16 ? 1 : 0;
    const int needed = (p.n_jobs + e->async_warps - 1) / e->async_warps;
    const int grid = std::max(1, std::min(e->sm_count, needed));
    const int variant = (((p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY) ? 1 : 0) | ((p.flags & IPP_FLAG_ADAPTIVE) ? 2 : 0) |
                        ((p.noise != nullptr || p.z_out != nullptr) ? 4 : 0) | (e->cfg.layout == IPP_LAYOUT_TILED ? 8 : 0);
    async_variant(variant)<<<grid, e->async_warps * 32, e->async_smem, e->stream>>>(ap);
    e->ticket_parity ^= 1;
    e->launches++;
    e->path_launches[IPP_PATH_ASYNC]++;
    CU(e, cudaGetLastError());
    return IPP_OK;
}

// the path a Kalman step on action ids takes right now
static int effective_path(const ipp_engine *e) {
    if (e->step_path != IPP_PATH_LSU && (e->async_ok || e->bulk_ok)) return IPP_PATH_ASYNC;
    return IPP_PATH_LSU;
}

// ------------------------------------------------------------------------------------------------
// life cycle
// ------------------------------------------------------------------------------------------------
extern "C" int ipp_create(const ipp_config *cfg, ipp_engine **out) {
    if (!cfg || !out) return fail(nullptr, IPP_ERR_INVALID, "ipp_create: NULL argument");
    *out = nullptr;
    if (cfg->struct_bytes != sizeof(ipp_config) || cfg->abi_version != IPP_ABI_VERSION)
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: ABI mismatch (struct_bytes %u vs %zu, version %u vs %d)", cfg->struct_bytes,
                    sizeof(ipp_config), cfg->abi_version, IPP_ABI_VERSION);
    if (cfg->batch < 1 || cfg->x_dim < 1 || cfg->y_dim < 1) return fail(nullptr, IPP_ERR_INVALID, "ipp_create: batch/x_dim/y_dim must be >= 1");
    if ((double)cfg->x_dim * cfg->y_dim > 1.0e9) return fail(nullptr, IPP_ERR_INVALID, "ipp_create: grid too large");
    if (!(cfg->resolution > 0)) return fail(nullptr, IPP_ERR_INVALID, "ipp_create: environment.resolution must be > 0");
    if (!(cfg->angle_x_deg > 0 && cfg->angle_x_deg < 180 && cfg->angle_y_deg > 0 && cfg->angle_y_deg < 180))
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: field_of_view angles must be in (0, 180)");
    if (cfg->layout != IPP_LAYOUT_PLANES && cfg->layout != IPP_LAYOUT_MV && cfg->layout != IPP_LAYOUT_TILED && cfg->layout != IPP_LAYOUT_SUPER &&
        cfg->layout != IPP_LAYOUT_SPLIT)
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: unknown layout %d", cfg->layout);
    if (cfg->cost_mode != IPP_COST_DISTANCE && cfg->cost_mode != IPP_COST_FLIGHT_TIME)
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: unknown cost_mode %d", cfg->cost_mode);
    if (cfg->cost_mode == IPP_COST_FLIGHT_TIME && !(cfg->max_v > 0 && cfg->max_a > 0))
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: uav max_v / max_a must be > 0");
    if (!(cfg->altitude_spacing > 0) || !(cfg->max_altitude >= cfg->min_altitude) || !(cfg->min_altitude > 0))
        return fail(nullptr, IPP_ERR_INVALID, "ipp_create: need 0 < min_altitude <= max_altitude and altitude_spacing > 0");

    ipp_engine *e = new (std::nothrow) ipp_engine();
    if (!e) return fail(nullptr, IPP_ERR_NOMEM, "ipp_create: out of host memory");
    e->cfg = *cfg;
    const double kPi = 3.14159265358979323846;
    if (e->cfg.tan_half_x == 0) e->cfg.tan_half_x = std::tan(0.5 * (e->cfg.angle_x_deg * (kPi / 180.0)));
    if (e->cfg.tan_half_y == 0) e->cfg.tan_half_y = std::tan(0.5 * (e->cfg.angle_y_deg * (kPi / 180.0)));
    e->plane = (size_t)cfg->x_dim * cfg->y_dim;
    e->plane_mv = e->plane_gt = e->plane;
    if (cfg->layout == IPP_LAYOUT_TILED) {
        e->txm = (cfg->x_dim + 3) / 4;
        e->txg = (cfg->x_dim + 7) / 8;
        e->tiles_y = (cfg->y_dim + 3) / 4;
        e->plane_mv = (size_t)e->tiles_y * e->txm * 16;
        e->plane_gt = (size_t)e->tiles_y * e->txg * 32;
        e->ts_mv = 16;
        e->ts_gt = 32;
        e->gw_shift = 3;
    } else if (cfg->layout == IPP_LAYOUT_SUPER) {  // one array: 192-byte super-tiles {mean,var} x 16 | gt x 16
        e->txm = e->txg = (cfg->x_dim + 3) / 4;
        e->tiles_y = (cfg->y_dim + 3) / 4;
        e->plane_mv = (size_t)e->tiles_y * e->txm * 24;  // float2 units
        e->plane_gt = (size_t)e->tiles_y * e->txm * 48;  // float units
        e->ts_mv = 24;
        e->ts_gt = 48;
        e->gw_shift = 2;
    } else if (cfg->layout == IPP_LAYOUT_SPLIT) {  // var[tiles][16] and {mean x 16 | gt x 16}[tiles]
        e->txm = e->txg = (cfg->x_dim + 3) / 4;
        e->tiles_y = (cfg->y_dim + 3) / 4;
        e->plane_mv = (size_t)e->tiles_y * e->txm * 16;  // floats of the var array
        e->plane_gt = (size_t)e->tiles_y * e->txm * 32;  // floats of the {mean | gt} array
        e->ts_mv = 16;
        e->ts_gt = 32;
        e->gw_shift = 2;
    }

    auto bail = [&](int rc) {
        g_create_err = e->err;
        ipp_destroy(e);
        return rc;
    };
    int rc = build_lut(e);
    if (rc != IPP_OK) return bail(rc);

    cudaError_t s = cudaSetDevice(cfg->device);
    if (s != cudaSuccess) return bail(fail(e, IPP_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(s)));
    cudaDeviceProp prop;
    s = cudaGetDeviceProperties(&prop, cfg->device);
    if (s != cudaSuccess) return bail(fail(e, IPP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(s)));
    e->sm_count = prop.multiProcessorCount;
    // Footprint rows are short unaligned segments (36-184 B): ask L2 to fetch 32 B sectors from HBM
    // instead of the default 64 B pairs (ncu: DRAM read bytes 1.7x the L2 miss bytes otherwise).
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    if (prop.major < 10) return bail(fail(e, IPP_ERR_UNSUPPORTED, "device %d is sm_%d%d; this engine is built for sm_100a only", cfg->device, prop.major, prop.minor));

    if (cfg->stream) {
        e->stream = (cudaStream_t)cfg->stream;
    } else {
        s = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
        if (s != cudaSuccess) return bail(fail(e, IPP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(s)));
        e->own_stream = true;
    }
    const size_t cells = e->plane_mv * (size_t)cfg->batch, cells_gt = e->plane_gt * (size_t)cfg->batch;
    if (cfg->layout == IPP_LAYOUT_SPLIT) {
        if ((rc = dev_alloc(e, &e->d_mean, cells_gt)) != IPP_OK) return bail(rc);
        if ((rc = dev_alloc(e, &e->d_var, cells)) != IPP_OK) return bail(rc);
        e->d_gt = e->d_mean + 16;  // the ground-truth half of tile 0 (same allocation)
        e->gt_aliases_mean = true;
        cudaMemsetAsync(e->d_mean, 0, cells_gt * sizeof(float), e->stream);  // padding cells of partial tiles are staged: keep them finite
        cudaMemsetAsync(e->d_var, 0, cells * sizeof(float), e->stream);
    } else if (cfg->layout != IPP_LAYOUT_PLANES) {
        if ((rc = dev_alloc(e, &e->d_mean, 2 * cells)) != IPP_OK) return bail(rc);
    } else {
        if ((rc = dev_alloc(e, &e->d_mean, cells)) != IPP_OK) return bail(rc);
        if ((rc = dev_alloc(e, &e->d_var, cells)) != IPP_OK) return bail(rc);
    }
    if (cfg->layout == IPP_LAYOUT_SUPER) {
        e->d_gt = e->d_mean + 32;  // the ground-truth third of super-tile 0 (same allocation)
        e->gt_aliases_mean = true;
        cudaMemsetAsync(e->d_mean, 0, 2 * cells * sizeof(float), e->stream);
    } else if (cfg->layout != IPP_LAYOUT_SPLIT && (rc = dev_alloc(e, &e->d_gt, cells_gt)) != IPP_OK)
        return bail(rc);
    if (cfg->layout == IPP_LAYOUT_TILED) {  // padding cells of partial tiles are staged (never used): keep them finite
        cudaMemsetAsync(e->d_gt, 0, cells_gt * sizeof(float), e->stream);
        cudaMemsetAsync(e->d_mean, 0, 2 * cells * sizeof(float), e->stream);
    }
    if ((rc = dev_alloc(e, &e->d_prev, 3 * (size_t)cfg->batch)) != IPP_OK) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_metrics, (size_t)cfg->batch * IPP_NUM_METRICS)) != IPP_OK) return bail(rc);
    // status word: mapped pinned host memory, written by the kernels only on the (rare) error path, read by the
    // host after a stream synchronisation without a device->host copy
    s = cudaHostAlloc((void **)&e->h_status, sizeof(int), cudaHostAllocMapped);
    if (s == cudaSuccess) s = cudaHostGetDevicePointer((void **)&e->d_status, e->h_status, 0);
    if (s != cudaSuccess) return bail(fail(e, IPP_ERR_CUDA, "status init: %s", cudaGetErrorString(s)));
    *e->h_status = 0;
    if (const char *zc = getenv("IPP_ZERO_COPY")) {
        e->zero_copy = 0;
        if (strchr(zc, 'r')) e->zero_copy |= IPP_ZERO_COPY_REWARDS;
        if (strchr(zc, 'i')) e->zero_copy |= IPP_ZERO_COPY_IDS;
        if (strchr(zc, 'f')) e->zero_copy |= IPP_ZERO_COPY_IDS_FETCH;
    }
    if ((rc = setup_async(e)) != IPP_OK) return bail(rc);
    if ((rc = setup_bulk(e)) != IPP_OK) return bail(rc);
    if (const char *sp = getenv("IPP_STEP_PATH")) {
        if (!strcmp(sp, "lsu")) e->step_path = IPP_PATH_LSU;
        if (!strcmp(sp, "async")) e->step_path = IPP_PATH_ASYNC;
    }
    *out = e;
    return IPP_OK;
}

extern "C" void ipp_destroy(ipp_engine *e) {
    if (!e) return;
    if (e->stream) cudaStreamSynchronize(e->stream);
    void *ptrs[] = {e->d_mean, e->d_var, e->gt_aliases_mean ? nullptr : e->d_gt, e->d_prev, e->d_actions, e->d_poses, e->d_prev_in, e->d_env_index,
                    e->d_reward, e->d_noise, e->d_z, e->d_metrics, e->d_scratch, e->d_tickets, e->d_level_taps, e->d_slice_state};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (e->h_status) cudaFreeHost(e->h_status);
    for (int k = 0; k < IPP_STEP_SLOTS; ++k) {
        if (e->ev_h2d[k]) cudaEventDestroy(e->ev_h2d[k]);
        if (e->ev_done[k]) cudaEventDestroy(e->ev_done[k]);
        if (e->d_actions_slot[k]) cudaFree(e->d_actions_slot[k]);
        if (e->d_reward_slot[k]) cudaFree(e->d_reward_slot[k]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" const char *ipp_last_error(const ipp_engine *e) { return e ? e->err.c_str() : g_create_err.c_str(); }

extern "C" int ipp_get_info(const ipp_engine *e, ipp_info *out) {
    if (!e || !out) return IPP_ERR_INVALID;
    memset(out, 0, sizeof *out);
    out->batch = e->cfg.batch;
    out->x_dim = e->cfg.x_dim;
    out->y_dim = e->cfg.y_dim;
    out->layout = e->cfg.layout;
    out->num_altitude_levels = e->n_levels;
    out->num_actions = (int32_t)((size_t)e->n_levels * e->plane);
    out->max_measurements = e->max_meas;
    out->sm_count = e->sm_count;
    out->launches = e->launches;
    out->steps = e->steps;
    out->device_bytes = e->device_bytes;
    for (int k = 0; k < e->n_levels; ++k) {
        out->altitude[k] = e->lut[k].alt;
        out->radius_x[k] = e->lut[k].rx;
        out->radius_y[k] = e->lut[k].ry;
    }
    return IPP_OK;
}

static TiledDims tiled_dims(const ipp_engine *e) {
    TiledDims d;
    d.X = e->cfg.x_dim;
    d.Y = e->cfg.y_dim;
    d.txm = e->txm;
    d.txg = e->txg;
    d.ts_mv = e->ts_mv;
    d.ts_gt = e->ts_gt;
    d.gw_shift = e->gw_shift;
    d.plane = e->plane;
    d.plane_mv = e->plane_mv;
    d.plane_gt = e->plane_gt;
    return d;
}

// the kernels' status word (mapped host memory): bit 0 unsupported footprint, bit 1 invalid action id
static int status_error(ipp_engine *e) {
    const int st = *(volatile int *)e->h_status;
    if (st == 0) return IPP_OK;
    *(volatile int *)e->h_status = 0;
    if (st & 4) return fail(e, IPP_ERR_CUDA, "the step kernel timed out waiting for a slice of host action ids (IPP_ZERO_COPY_IDS_FETCH)");
    if (st & 2)
        return fail(e, IPP_ERR_INVALID, "action id outside the action table (planning/common/actions.py:73-91: level * N + x_dim * col + row); the step ran on the clamped id");
    return fail(e, IPP_ERR_UNSUPPORTED, "footprint needs cv2 INTER_AREA with an up-sampling axis (non-square FoV/grid corner case); not supported");
}

static int check_status(ipp_engine *e) {
    // the status word lives in mapped host memory: visible here once the stream has drained
    CU(e, cudaStreamSynchronize(e->stream));
    for (int k = 0; k < IPP_STEP_SLOTS; ++k) e->slot_busy[k] = false;  // submitted steps run on this stream: drained too
    return status_error(e);
}

extern "C" int ipp_sync(ipp_engine *e) {
    if (!e) return IPP_ERR_INVALID;
    return check_status(e);
}

static int ensure_job_buffers(ipp_engine *e, size_t n);

static int grid_for(size_t n, int threads, int sm_count) {
    const size_t want = (n + threads - 1) / threads;
    const size_t cap = (size_t)sm_count * 16;
    return (int)std::max<size_t>(1, std::min(want, cap));
}

extern "C" int ipp_reset(ipp_engine *e, float prior_mean, float prior_var, const float *prior_var_per_env, const double *init_pose) {
    if (!e) return IPP_ERR_INVALID;
    if (!(prior_var > 0) && !prior_var_per_env) return fail(e, IPP_ERR_INVALID, "ipp_reset: prior variance must be > 0");
    const int B = e->cfg.batch;
    float *d_pv = nullptr;
    if (prior_var_per_env) {
        int rc = ensure_job_buffers(e, (size_t)B);  // borrow the reward staging as a [B] float buffer
        if (rc != IPP_OK) return rc;
        CU(e, cudaMemcpyAsync(e->d_reward, prior_var_per_env, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        d_pv = e->d_reward;
    }
    reset_kernel<<<grid_for(e->plane_mv * (size_t)B, 256, e->sm_count), 256, 0, e->stream>>>(e->d_mean, e->d_var, e->cfg.layout, e->plane_mv, B,
                                                                                             prior_mean, prior_var, d_pv);
    const double dflt[3] = {2.0, 2.0, 14.0};  // planning/missions.py:69
    const double *ip = init_pose ? init_pose : dflt;
    fill_prev_kernel<<<(B + 255) / 256, 256, 0, e->stream>>>(e->d_prev, B, ip[0], ip[1], ip[2]);
    e->launches += 2;
    e->steps = 0;
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

// prior fill with a per-env {level, relative spread}: cell variance = level * (1 + spread * N(0, 1)) (Philox per cell)
__global__ void reset_shuffled_kernel(float *mean, float *var, int layout, size_t plane, int batch, float prior_mean, const float *scale,
                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t env0) {
    const size_t total = plane * (size_t)batch;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (layout == IPP_LAYOUT_SUPER && (i % 24) >= 16) continue;
        const size_t env = i / plane;
        const float level = scale[2 * env], spread = scale[2 * env + 1];
        float pv = level;
        if (spread > 0.0f) {
            const uint32_t cell = (uint32_t)(i - env * plane);
            uint32_t r[4];
            philox4x32_10(cell >> 1, env0 + (uint32_t)env, 11u, 0x5eedu, seed_lo, seed_hi, r);
            float n0, n1;
            box_muller(r[0], r[1], n0, n1);
            pv = fmaxf(level * (1.0f + spread * ((cell & 1) ? n1 : n0)), 1.0e-6f);
        }
        if (layout == IPP_LAYOUT_SPLIT) {
            mean[i + (i & ~(size_t)15)] = prior_mean;
            var[i] = pv;
        } else if (layout != IPP_LAYOUT_PLANES) {
            reinterpret_cast<float2 *>(mean)[i] = make_float2(prior_mean, pv);
        } else {
            mean[i] = prior_mean;
            var[i] = pv;
        }
    }
}

extern "C" int ipp_reset_shuffled(ipp_engine *e, float prior_mean, int32_t gp_mode, float p0, uint64_t seed, const double *init_pose) {
    if (!e) return IPP_ERR_INVALID;
    if (!(p0 > 0) || (!gp_mode && !(p0 > 0.1f))) return fail(e, IPP_ERR_INVALID, "ipp_reset_shuffled: signal_variance must be > 0 / prior_cov_mean > 0.1");
    const int B = e->cfg.batch;
    int rc = ensure_job_buffers(e, 2 * (size_t)B);  // borrow the reward staging as a [B][2] float buffer
    if (rc != IPP_OK) return rc;
    if ((rc = ipp_internal_shuffled_prior(e, e->d_reward, B, gp_mode, p0, p0, seed)) != IPP_OK) return fail(e, rc, "ipp_reset_shuffled: launch failed");
    reset_shuffled_kernel<<<grid_for(e->plane_mv * (size_t)B, 256, e->sm_count), 256, 0, e->stream>>>(
        e->d_mean, e->d_var, e->cfg.layout, e->plane_mv, B, prior_mean, e->d_reward, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32),
        (uint32_t)e->cfg.env_id_offset);
    const double dflt[3] = {2.0, 2.0, 14.0};  // planning/missions.py:69
    const double *ip = init_pose ? init_pose : dflt;
    fill_prev_kernel<<<(B + 255) / 256, 256, 0, e->stream>>>(e->d_prev, B, ip[0], ip[1], ip[2]);
    e->launches += 2;
    e->steps = 0;
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

static int check_range(ipp_engine *e, int32_t first, int32_t n, const char *who) {
    if (first < 0 || n < 0 || (int64_t)first + n > e->cfg.batch) return fail(e, IPP_ERR_INVALID, "%s: env range [%d, %d) outside batch %d", who, first, first + n, e->cfg.batch);
    return IPP_OK;
}

// dense <-> tiled ground truth through the dense scratch buffer (TILED layout)
static int gt_transfer_tiled(ipp_engine *e, float *user, int32_t first_env, int32_t n_env, int32_t user_is_device, bool upload) {
    const size_t n = (size_t)n_env * e->plane;
    float *dense = user;
    int rc;
    if (!user_is_device) {
        if ((rc = ensure(e, &e->d_scratch, &e->cap_scratch, n)) != IPP_OK) return rc;
        dense = e->d_scratch;
        if (upload) CU(e, cudaMemcpyAsync(dense, user, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    }
    tiled_gt_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(e->d_gt + (size_t)first_env * e->plane_gt, dense, tiled_dims(e), n,
                                                                         upload ? 1 : 0);
    e->launches++;
    CU(e, cudaGetLastError());
    if (!user_is_device && !upload) CU(e, cudaMemcpyAsync(user, dense, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_set_ground_truth(ipp_engine *e, const float *gt, int32_t first_env, int32_t n_env, int32_t src_is_device) {
    if (!e || !gt) return IPP_ERR_INVALID;
    int rc = check_range(e, first_env, n_env, "ipp_set_ground_truth");
    if (rc != IPP_OK) return rc;
    if (n_env == 0) return IPP_OK;
    if (e->tiled()) return gt_transfer_tiled(e, const_cast<float *>(gt), first_env, n_env, src_is_device, true);
    CU(e, cudaMemcpyAsync(e->d_gt + (size_t)first_env * e->plane, gt, (size_t)n_env * e->plane * sizeof(float),
                          src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_get_ground_truth(ipp_engine *e, float *gt, int32_t first_env, int32_t n_env, int32_t dst_is_device) {
    if (!e || !gt) return IPP_ERR_INVALID;
    int rc = check_range(e, first_env, n_env, "ipp_get_ground_truth");
    if (rc != IPP_OK) return rc;
    if (n_env == 0) return IPP_OK;
    if (e->tiled()) return gt_transfer_tiled(e, gt, first_env, n_env, dst_is_device, false);
    CU(e, cudaMemcpyAsync(gt, e->d_gt + (size_t)first_env * e->plane, (size_t)n_env * e->plane * sizeof(float),
                          dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_synth_ground_truth(ipp_engine *e, uint64_t seed) {
    if (!e) return IPP_ERR_INVALID;
    dim3 grid((unsigned)std::min<size_t>((e->plane + 255) / 256, 64), (unsigned)e->cfg.batch);
    if (e->cfg.batch > 65535) {
        // grid.y limit: launch in slabs
        for (int first = 0; first < e->cfg.batch; first += 65535) {
            const int n = std::min(65535, e->cfg.batch - first);
            dim3 g2(grid.x, (unsigned)n);
            synth_gt_kernel<<<g2, 256, 0, e->stream>>>(e->d_gt + (size_t)first * e->plane_gt, e->plane, e->plane_gt, e->cfg.x_dim, e->txg, e->ts_gt, e->gw_shift, n,
                                                       (uint32_t)seed, (uint32_t)(e->cfg.env_id_offset + first));
            e->launches++;
        }
    } else {
        synth_gt_kernel<<<grid, 256, 0, e->stream>>>(e->d_gt, e->plane, e->plane_gt, e->cfg.x_dim, e->txg, e->ts_gt, e->gw_shift, e->cfg.batch, (uint32_t)seed,
                                                     (uint32_t)e->cfg.env_id_offset);
        e->launches++;
    }
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_get_state(ipp_engine *e, float *mean, float *var, int32_t first_env, int32_t n_env, int32_t dst_is_device) {
    if (!e) return IPP_ERR_INVALID;
    int rc = check_range(e, first_env, n_env, "ipp_get_state");
    if (rc != IPP_OK) return rc;
    const size_t n = (size_t)n_env * e->plane, off = (size_t)first_env * e->plane;
    const cudaMemcpyKind kind = dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (e->cfg.layout == IPP_LAYOUT_PLANES) {
        if (mean) CU(e, cudaMemcpyAsync(mean, e->d_mean + off, n * sizeof(float), kind, e->stream));
        if (var) CU(e, cudaMemcpyAsync(var, e->d_var + off, n * sizeof(float), kind, e->stream));
    } else {
        float *dm = mean, *dv = var;
        if (!dst_is_device) {
            if ((rc = ensure(e, &e->d_scratch, &e->cap_scratch, 2 * n)) != IPP_OK) return rc;
            dm = mean ? e->d_scratch : nullptr;
            dv = var ? e->d_scratch + n : nullptr;
        }
        if (e->split())
            split_unpack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(e->d_mean + (size_t)first_env * e->plane_gt,
                                                                                     e->d_var + (size_t)first_env * e->plane_mv, dm, dv, tiled_dims(e), n);
        else if (e->tiled())
            tiled_unpack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(
                reinterpret_cast<const float2 *>(e->d_mean) + (size_t)first_env * e->plane_mv, dm, dv, tiled_dims(e), n);
        else
            mv_unpack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(reinterpret_cast<const float2 *>(e->d_mean) + off, dm, dv, n);
        e->launches++;
        if (!dst_is_device) {
            if (mean) CU(e, cudaMemcpyAsync(mean, dm, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
            if (var) CU(e, cudaMemcpyAsync(var, dv, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        }
    }
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_set_state(ipp_engine *e, const float *mean, const float *var, int32_t first_env, int32_t n_env, int32_t src_is_device) {
    if (!e) return IPP_ERR_INVALID;
    int rc = check_range(e, first_env, n_env, "ipp_set_state");
    if (rc != IPP_OK) return rc;
    const size_t n = (size_t)n_env * e->plane, off = (size_t)first_env * e->plane;
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (e->cfg.layout == IPP_LAYOUT_PLANES) {
        if (mean) CU(e, cudaMemcpyAsync(e->d_mean + off, mean, n * sizeof(float), kind, e->stream));
        if (var) CU(e, cudaMemcpyAsync(e->d_var + off, var, n * sizeof(float), kind, e->stream));
    } else {
        const float *dm = mean, *dv = var;
        if (!src_is_device) {
            if ((rc = ensure(e, &e->d_scratch, &e->cap_scratch, 2 * n)) != IPP_OK) return rc;
            if (mean) {
                CU(e, cudaMemcpyAsync(e->d_scratch, mean, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
                dm = e->d_scratch;
            }
            if (var) {
                CU(e, cudaMemcpyAsync(e->d_scratch + n, var, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
                dv = e->d_scratch + n;
            }
        }
        if (e->split())
            split_pack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(e->d_mean + (size_t)first_env * e->plane_gt,
                                                                                   e->d_var + (size_t)first_env * e->plane_mv, dm, dv, tiled_dims(e), n);
        else if (e->tiled())
            tiled_pack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(
                reinterpret_cast<float2 *>(e->d_mean) + (size_t)first_env * e->plane_mv, dm, dv, tiled_dims(e), n);
        else
            mv_pack_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, e->stream>>>(reinterpret_cast<float2 *>(e->d_mean) + off, dm, dv, n);
        e->launches++;
    }
    CU(e, cudaGetLastError());
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" int ipp_set_prev_pose(ipp_engine *e, const double *poses) {
    if (!e || !poses) return IPP_ERR_INVALID;
    CU(e, cudaMemcpyAsync(e->d_prev, poses, 3 * (size_t)e->cfg.batch * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}
extern "C" int ipp_get_prev_pose(ipp_engine *e, double *poses) {
    if (!e || !poses) return IPP_ERR_INVALID;
    CU(e, cudaMemcpyAsync(poses, e->d_prev, 3 * (size_t)e->cfg.batch * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

// ------------------------------------------------------------------------------------------------
// the hot path
// ------------------------------------------------------------------------------------------------
static void fill_params(const ipp_engine *e, StepParams &p) {
    const ipp_config &c = e->cfg;
    memset(&p, 0, sizeof p);
    p.mean = e->d_mean;
    p.var = e->d_var;
    p.gt = e->d_gt;
    p.plane = e->plane_mv;
    p.plane_gt = e->plane_gt;
    p.txm = e->txm;
    p.txg = e->txg;
    p.ts_mv = e->ts_mv;
    p.ts_gt = e->ts_gt;
    p.gw_shift = e->gw_shift;
    p.X = c.x_dim;
    p.Y = c.y_dim;
    p.batch = c.batch;
    p.prev_state = e->d_prev;
    p.status = e->d_status;
    p.res = c.resolution;
    p.tan_x = c.tan_half_x;
    p.tan_y = c.tan_half_y;
    p.coeff_a = c.coeff_a;
    p.coeff_b = c.coeff_b;
    p.rf_alt = c.rf_altitude;
    if (c.cost_mode == IPP_COST_FLIGHT_TIME) {
        p.inv_v = (float)(1.0 / c.max_v);
        p.inv_a = (float)(1.0 / c.max_a);
        p.d_acc_max = (float)((c.max_v * c.max_v) / (2.0 * c.max_a));
    }
    p.inv_N = 1.0f / (float)((double)c.x_dim * c.y_dim);
    p.inv_X = 1.0f / (float)c.x_dim;
    p.thr = (float)c.value_threshold;
    p.kappa = (float)c.interval_factor;
    p.cost_mode = c.cost_mode;
    p.n_levels = e->n_levels;
    p.seed_lo = (uint32_t)(c.seed & 0xffffffffu);
    p.seed_hi = (uint32_t)(c.seed >> 32);
    p.step_lo = (uint32_t)(e->steps & 0xffffffffu);
    p.step_hi = (uint32_t)(e->steps >> 32);
    p.env_id_offset = (uint32_t)c.env_id_offset;
    memcpy(p.lut, e->lut, sizeof p.lut);
}

template <int MODE>
static void launch_mode(ipp_engine *e, const StepParams &p) {
    const int blocks = (p.n_jobs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (e->cfg.layout == IPP_LAYOUT_TILED)
        ipp_step_kernel<IPP_LAYOUT_TILED, MODE><<<blocks, kThreads, 0, e->stream>>>(p);
    else if (e->cfg.layout == IPP_LAYOUT_SUPER)
        ipp_step_kernel<IPP_LAYOUT_SUPER, MODE><<<blocks, kThreads, 0, e->stream>>>(p);
    else if (e->cfg.layout == IPP_LAYOUT_SPLIT)
        ipp_step_kernel<IPP_LAYOUT_SPLIT, MODE><<<blocks, kThreads, 0, e->stream>>>(p);
    else if (e->cfg.layout == IPP_LAYOUT_MV)
        ipp_step_kernel<IPP_LAYOUT_MV, MODE><<<blocks, kThreads, 0, e->stream>>>(p);
    else
        ipp_step_kernel<IPP_LAYOUT_PLANES, MODE><<<blocks, kThreads, 0, e->stream>>>(p);
    e->launches++;
}

static int launch_step(ipp_engine *e, StepParams &p, int mode) {
    if (p.n_jobs <= 0) return IPP_OK;
    if (mode == MODE_KALMAN)
        launch_mode<MODE_KALMAN>(e, p);
    else if (mode == MODE_PREDICT)
        launch_mode<MODE_PREDICT>(e, p);
    else
        launch_mode<MODE_LOGODDS>(e, p);
    CU(e, cudaGetLastError());
    return IPP_OK;
}

static int validate_step_args(ipp_engine *e, const void *ids, const void *poses, const char *who) {
    if (!e) return IPP_ERR_INVALID;
    if ((ids == nullptr) == (poses == nullptr)) return fail(e, IPP_ERR_INVALID, "%s: exactly one of action_ids / poses must be given", who);
    return IPP_OK;
}

// does a Kalman step on action ids with these flags take the bulk-copy persistent kernel?
static bool takes_bulk(const ipp_engine *e, const void *action_ids, uint32_t flags) {
    return action_ids != nullptr && (flags & (IPP_FLAG_LOGODDS | IPP_FLAG_NO_COMMIT)) == 0 && effective_path(e) == IPP_PATH_ASYNC && e->bulk_ok;
}

// host_ids: device alias of the caller's mapped host id buffer the bulk kernel fetches into action_ids itself, or nullptr
static int step_device_impl(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise, int32_t noise_stride,
                            float *reward, float *measurements, uint32_t flags, const int32_t *host_ids) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_step_device");
    if (rc != IPP_OK) return rc;
    if ((noise || measurements) && noise_stride < e->max_meas)
        return fail(e, IPP_ERR_INVALID, "ipp_step: noise_stride %d < max_measurements %d", noise_stride, e->max_meas);
    StepParams p;
    fill_params(e, p);
    p.n_jobs = e->cfg.batch;
    p.action_ids = action_ids;
    p.poses = poses;
    p.noise = noise;
    p.noise_stride = noise_stride;
    p.reward = reward;
    p.z_out = measurements;
    p.flags = flags;
    const int path = (action_ids != nullptr && (flags & (IPP_FLAG_LOGODDS | IPP_FLAG_NO_COMMIT)) == 0) ? effective_path(e) : IPP_PATH_LSU;
    if (path == IPP_PATH_ASYNC)
        rc = e->bulk_ok ? launch_bulk(e, p, false, host_ids) : launch_async(e, p);
    else {
        rc = launch_step(e, p, (flags & IPP_FLAG_LOGODDS) ? MODE_LOGODDS : MODE_KALMAN);
        e->path_launches[IPP_PATH_LSU]++;
    }
    if (rc == IPP_OK) e->steps++;
    return rc;
}

extern "C" int ipp_step_device(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise, int32_t noise_stride,
                               float *reward, float *measurements, uint32_t flags) {
    return step_device_impl(e, action_ids, poses, noise, noise_stride, reward, measurements, flags, nullptr);
}

static int ensure_job_buffers(ipp_engine *e, size_t n) {
    if (e->cap_jobs >= n && e->d_actions && e->d_poses && e->d_prev_in && e->d_env_index && e->d_reward) return IPP_OK;
    CU(e, cudaStreamSynchronize(e->stream));
    void **ps[] = {(void **)&e->d_actions, (void **)&e->d_poses, (void **)&e->d_prev_in, (void **)&e->d_env_index, (void **)&e->d_reward};
    for (void **pp : ps)
        if (*pp) {
            cudaFree(*pp);
            *pp = nullptr;
        }
    e->cap_jobs = 0;
    int rc;
    if ((rc = dev_alloc(e, &e->d_actions, n)) != IPP_OK) return rc;
    if ((rc = dev_alloc(e, &e->d_poses, 3 * n)) != IPP_OK) return rc;
    if ((rc = dev_alloc(e, &e->d_prev_in, 3 * n)) != IPP_OK) return rc;
    if ((rc = dev_alloc(e, &e->d_env_index, n)) != IPP_OK) return rc;
    if ((rc = dev_alloc(e, &e->d_reward, n)) != IPP_OK) return rc;
    e->cap_jobs = n;
    return IPP_OK;
}

// Device alias of a caller's HOST buffer when it is page-locked and mapped (cudaHostAlloc / cudaHostRegister /
// torch pin_memory under unified addressing), else NULL.  Lets the fused kernel read its action ids from and write
// its rewards to the caller's buffer directly: the transfer rides inside the kernel (4 B per env each way) instead
// of being two more stream operations around it.
static void *mapped_alias(const void *host) {
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

extern "C" int ipp_step(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise, int32_t noise_stride, float *reward,
                        float *measurements, uint32_t flags) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_step");
    if (rc != IPP_OK) return rc;
    const size_t B = (size_t)e->cfg.batch;
    if ((rc = ensure_job_buffers(e, B)) != IPP_OK) return rc;
    if ((noise || measurements) && noise_stride < e->max_meas)
        return fail(e, IPP_ERR_INVALID, "ipp_step: noise_stride %d < max_measurements %d", noise_stride, e->max_meas);
    // zero-copy: mapped pinned caller buffers go to the kernel as they are (IPP_OPT_ZERO_COPY)
    const int32_t *ids_dev = nullptr, *ids_fetch = nullptr;
    float *reward_dev = nullptr;
    if (action_ids && (e->zero_copy & IPP_ZERO_COPY_IDS_FETCH) && takes_bulk(e, action_ids, flags)) {
        ids_fetch = (const int32_t *)mapped_alias(action_ids);  // the kernel pulls the ids into d_actions itself
        if (((uintptr_t)ids_fetch & 15) != 0) ids_fetch = nullptr;
        if (ids_fetch) ids_dev = e->d_actions;
    }
    if (action_ids && !ids_dev && (e->zero_copy & IPP_ZERO_COPY_IDS)) ids_dev = (const int32_t *)mapped_alias(action_ids);
    if (reward && (e->zero_copy & IPP_ZERO_COPY_REWARDS)) reward_dev = (float *)mapped_alias(reward);
    if (action_ids && !ids_dev) {
        CU(e, cudaMemcpyAsync(e->d_actions, action_ids, B * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
        ids_dev = e->d_actions;
    }
    if (poses) CU(e, cudaMemcpyAsync(e->d_poses, poses, 3 * B * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (noise) {
        if ((rc = ensure(e, &e->d_noise, &e->cap_noise, B * (size_t)noise_stride)) != IPP_OK) return rc;
        CU(e, cudaMemcpyAsync(e->d_noise, noise, B * (size_t)noise_stride * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    }
    if (measurements) {
        if ((rc = ensure(e, &e->d_z, &e->cap_z, B * (size_t)noise_stride)) != IPP_OK) return rc;
        // entries past an env's measurement count are unspecified by the kernels: hand back zeros, not stale device memory
        CU(e, cudaMemsetAsync(e->d_z, 0, B * (size_t)noise_stride * sizeof(float), e->stream));
    }
    rc = step_device_impl(e, ids_dev, poses ? e->d_poses : nullptr, noise ? e->d_noise : nullptr, noise_stride,
                          reward_dev ? reward_dev : e->d_reward, measurements ? e->d_z : nullptr, flags, ids_fetch);
    if (rc != IPP_OK) return rc;
    if (reward_dev)
        e->zero_copy_steps++;
    else if (reward)
        CU(e, cudaMemcpyAsync(reward, e->d_reward, B * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    if (measurements) CU(e, cudaMemcpyAsync(measurements, e->d_z, B * (size_t)noise_stride * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    return check_status(e);
}

// Pipelined host steps.  ipp_step_submit(slot) queues {H2D of the ids on the copy stream, the fused kernel on the engine's
// stream after it, rewards straight into the caller's pinned buffer or a D2H copy} and returns; ipp_step_wait(slot) blocks
// until that step is complete.  With two slots the host prepares and uploads step t+1 while step t computes.
extern "C" int ipp_step_submit(ipp_engine *e, int32_t slot, const int32_t *action_ids, float *reward, uint32_t flags) {
    if (!e) return IPP_ERR_INVALID;
    if (slot < 0 || slot >= IPP_STEP_SLOTS) return fail(e, IPP_ERR_INVALID, "ipp_step_submit: slot %d outside [0, %d)", slot, IPP_STEP_SLOTS);
    if (!action_ids) return fail(e, IPP_ERR_INVALID, "ipp_step_submit: action_ids == NULL");
    if (e->slot_busy[slot]) return fail(e, IPP_ERR_INVALID, "ipp_step_submit: slot %d is still in flight (ipp_step_wait first)", slot);
    const size_t B = (size_t)e->cfg.batch;
    int rc;
    if (!e->copy_stream) {
        CU(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < IPP_STEP_SLOTS; ++k) {
            CU(e, cudaEventCreateWithFlags(&e->ev_h2d[k], cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming));
            if ((rc = dev_alloc(e, &e->d_actions_slot[k], B)) != IPP_OK) return rc;
            if ((rc = dev_alloc(e, &e->d_reward_slot[k], B)) != IPP_OK) return rc;
        }
    }
    // the slot's id buffer was last read by the kernel of the step waited for before this call: free to overwrite
    CU(e, cudaMemcpyAsync(e->d_actions_slot[slot], action_ids, B * sizeof(int32_t), cudaMemcpyHostToDevice, e->copy_stream));
    CU(e, cudaEventRecord(e->ev_h2d[slot], e->copy_stream));
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_h2d[slot], 0));
    float *reward_dev = (reward && (e->zero_copy & IPP_ZERO_COPY_REWARDS)) ? (float *)mapped_alias(reward) : nullptr;
    rc = ipp_step_device(e, e->d_actions_slot[slot], nullptr, nullptr, 0, reward_dev ? reward_dev : e->d_reward_slot[slot], nullptr, flags);
    if (rc != IPP_OK) return rc;
    if (reward_dev)
        e->zero_copy_steps++;
    else if (reward)
        CU(e, cudaMemcpyAsync(reward, e->d_reward_slot[slot], B * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaEventRecord(e->ev_done[slot], e->stream));
    e->slot_busy[slot] = true;
    return IPP_OK;
}

extern "C" int ipp_step_wait(ipp_engine *e, int32_t slot) {
    if (!e) return IPP_ERR_INVALID;
    if (slot < 0 || slot >= IPP_STEP_SLOTS) return fail(e, IPP_ERR_INVALID, "ipp_step_wait: slot %d outside [0, %d)", slot, IPP_STEP_SLOTS);
    if (!e->slot_busy[slot]) return IPP_OK;
    CU(e, cudaEventSynchronize(e->ev_done[slot]));
    e->slot_busy[slot] = false;
    return status_error(e);
}

// Measurement only (Sensor.take_measurement, sensors/cameras.py:108-116) and update with a
// caller-supplied measurement (Mapping.update_grid_map(pos, data), mapping/mappings.py:114-153):
// the two halves of ipp_step for the B=1 facade, same kernel.
extern "C" int ipp_measure(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *noise, int32_t stride,
                           float *measurements, uint32_t flags) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_measure");
    if (rc != IPP_OK) return rc;
    if (!measurements) return fail(e, IPP_ERR_INVALID, "ipp_measure: measurements == NULL");
    if (stride < e->max_meas) return fail(e, IPP_ERR_INVALID, "ipp_measure: stride %d < max_measurements %d", stride, e->max_meas);
    const size_t B = (size_t)e->cfg.batch;
    if ((rc = ensure_job_buffers(e, B)) != IPP_OK) return rc;
    if ((rc = ensure(e, &e->d_z, &e->cap_z, B * (size_t)stride)) != IPP_OK) return rc;
    CU(e, cudaMemsetAsync(e->d_z, 0, B * (size_t)stride * sizeof(float), e->stream));
    if (action_ids) CU(e, cudaMemcpyAsync(e->d_actions, action_ids, B * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (poses) CU(e, cudaMemcpyAsync(e->d_poses, poses, 3 * B * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (noise) {
        if ((rc = ensure(e, &e->d_noise, &e->cap_noise, B * (size_t)stride)) != IPP_OK) return rc;
        CU(e, cudaMemcpyAsync(e->d_noise, noise, B * (size_t)stride * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    }
    StepParams p;
    fill_params(e, p);
    p.n_jobs = e->cfg.batch;
    p.action_ids = action_ids ? e->d_actions : nullptr;
    p.poses = poses ? e->d_poses : nullptr;
    p.noise = noise ? e->d_noise : nullptr;
    p.noise_stride = stride;
    p.z_out = e->d_z;
    p.flags = flags;
    p.measure_only = 1;
    if ((rc = launch_step(e, p, MODE_KALMAN)) != IPP_OK) return rc;
    if (!noise) e->steps++;  // a Philox draw was consumed
    CU(e, cudaMemcpyAsync(measurements, e->d_z, B * (size_t)stride * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    return check_status(e);
}

extern "C" int ipp_update(ipp_engine *e, const int32_t *action_ids, const double *poses, const float *measurements, int32_t stride,
                          float *reward, uint32_t flags) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_update");
    if (rc != IPP_OK) return rc;
    if (!measurements) return fail(e, IPP_ERR_INVALID, "ipp_update: measurements == NULL");
    if (stride < e->max_meas) return fail(e, IPP_ERR_INVALID, "ipp_update: stride %d < max_measurements %d", stride, e->max_meas);
    const size_t B = (size_t)e->cfg.batch;
    if ((rc = ensure_job_buffers(e, B)) != IPP_OK) return rc;
    if ((rc = ensure(e, &e->d_z, &e->cap_z, B * (size_t)stride)) != IPP_OK) return rc;
    if (action_ids) CU(e, cudaMemcpyAsync(e->d_actions, action_ids, B * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (poses) CU(e, cudaMemcpyAsync(e->d_poses, poses, 3 * B * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaMemcpyAsync(e->d_z, measurements, B * (size_t)stride * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    StepParams p;
    fill_params(e, p);
    p.n_jobs = e->cfg.batch;
    p.action_ids = action_ids ? e->d_actions : nullptr;
    p.poses = poses ? e->d_poses : nullptr;
    p.z_in = e->d_z;
    p.noise_stride = stride;
    p.reward = e->d_reward;
    p.flags = flags;
    if ((rc = launch_step(e, p, (flags & IPP_FLAG_LOGODDS) ? MODE_LOGODDS : MODE_KALMAN)) != IPP_OK) return rc;
    if (reward) CU(e, cudaMemcpyAsync(reward, e->d_reward, B * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    return check_status(e);
}

extern "C" int ipp_predict_device(ipp_engine *e, int32_t n_jobs, const int32_t *env_index, const int32_t *action_ids, const double *poses,
                                  const double *prev_poses, float *reward, uint32_t flags) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_predict_device");
    if (rc != IPP_OK) return rc;
    if (n_jobs < 0) return fail(e, IPP_ERR_INVALID, "ipp_predict: n_jobs < 0");
    if (!env_index && n_jobs != e->cfg.batch) return fail(e, IPP_ERR_INVALID, "ipp_predict: env_index == NULL requires n_jobs == batch");
    StepParams p;
    fill_params(e, p);
    p.n_jobs = n_jobs;
    p.env_index = env_index;
    p.action_ids = action_ids;
    p.poses = poses;
    p.prev_in = prev_poses;
    p.reward = reward;
    p.flags = flags;
    // whole-batch prediction steps on action ids (the rollout loop of the planners) take the persistent path
    if (!env_index && action_ids && e->bulk_ok && effective_path(e) == IPP_PATH_ASYNC && (flags & IPP_FLAG_LOGODDS) == 0)
        return launch_bulk(e, p, true);
    rc = launch_step(e, p, MODE_PREDICT);
    if (rc == IPP_OK) e->path_launches[IPP_PATH_LSU]++;
    return rc;
}

extern "C" int ipp_predict(ipp_engine *e, int32_t n_jobs, const int32_t *env_index, const int32_t *action_ids, const double *poses,
                           const double *prev_poses, float *reward, uint32_t flags) {
    int rc = validate_step_args(e, action_ids, poses, "ipp_predict");
    if (rc != IPP_OK) return rc;
    if (n_jobs < 0) return fail(e, IPP_ERR_INVALID, "ipp_predict: n_jobs < 0");
    if (n_jobs == 0) return IPP_OK;
    const size_t J = (size_t)n_jobs;
    if (env_index) {
        for (size_t j = 0; j < J; ++j)
            if (env_index[j] < 0 || env_index[j] >= e->cfg.batch) return fail(e, IPP_ERR_INVALID, "ipp_predict: env_index[%zu] = %d outside batch", j, env_index[j]);
    }
    if ((rc = ensure_job_buffers(e, std::max(J, (size_t)e->cfg.batch))) != IPP_OK) return rc;
    if (env_index) CU(e, cudaMemcpyAsync(e->d_env_index, env_index, J * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (action_ids) CU(e, cudaMemcpyAsync(e->d_actions, action_ids, J * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (poses) CU(e, cudaMemcpyAsync(e->d_poses, poses, 3 * J * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    if (prev_poses) CU(e, cudaMemcpyAsync(e->d_prev_in, prev_poses, 3 * J * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    rc = ipp_predict_device(e, n_jobs, env_index ? e->d_env_index : nullptr, action_ids ? e->d_actions : nullptr, poses ? e->d_poses : nullptr,
                            prev_poses ? e->d_prev_in : nullptr, e->d_reward, flags);
    if (rc != IPP_OK) return rc;
    if (reward) CU(e, cudaMemcpyAsync(reward, e->d_reward, J * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    return check_status(e);
}

// Path rollouts (rollout_kernel.cuh): `horizon` chained prediction steps per job from the env's current belief,
// nothing written back.
extern "C" int ipp_rollout_device(ipp_engine *e, int32_t n_jobs, int32_t horizon, const int32_t *env_index, const int32_t *path_action_ids,
                                  const double *prev_poses, float *rewards, uint32_t flags) {
    if (!e || !path_action_ids || !rewards) return IPP_ERR_INVALID;
    if (n_jobs < 0) return fail(e, IPP_ERR_INVALID, "ipp_rollout: n_jobs < 0");
    if (horizon < 1 || horizon > kMaxHorizon) return fail(e, IPP_ERR_INVALID, "ipp_rollout: horizon %d outside [1, %d]", horizon, kMaxHorizon);
    if (!env_index && n_jobs != e->cfg.batch) return fail(e, IPP_ERR_INVALID, "ipp_rollout: env_index == NULL requires n_jobs == batch");
    if (flags & IPP_FLAG_LOGODDS) return fail(e, IPP_ERR_UNSUPPORTED, "ipp_rollout: log-odds rollouts are not implemented");
    if (n_jobs == 0) return IPP_OK;
    RolloutParams rp;
    fill_params(e, rp.base);
    rp.base.n_jobs = n_jobs;
    rp.base.env_index = env_index;
    rp.base.prev_in = prev_poses;
    rp.base.flags = flags;
    rp.path_actions = path_action_ids;
    rp.rewards = rewards;
    rp.horizon = horizon;
    int cells = 1;
    for (int k = 0; k < e->n_levels; ++k)
        cells = std::max(cells, std::min(2 * e->lut[k].rx + 1, e->cfg.x_dim) * std::min(2 * e->lut[k].ry + 1, e->cfg.y_dim));
    rp.tile_floats = cells;
    const size_t smem = (size_t)kRolloutWarps * (horizon - 1) * cells * sizeof(float);
    void (*kern)(const RolloutParams) = e->cfg.layout == IPP_LAYOUT_TILED ? ipp_rollout_kernel<IPP_LAYOUT_TILED>
                                        : e->cfg.layout == IPP_LAYOUT_SUPER ? ipp_rollout_kernel<IPP_LAYOUT_SUPER>
                                        : e->cfg.layout == IPP_LAYOUT_SPLIT ? ipp_rollout_kernel<IPP_LAYOUT_SPLIT>
                                        : e->cfg.layout == IPP_LAYOUT_MV  ? ipp_rollout_kernel<IPP_LAYOUT_MV>
                                                                          : ipp_rollout_kernel<IPP_LAYOUT_PLANES>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return fail(e, IPP_ERR_UNSUPPORTED, "ipp_rollout: %zu B of shared memory per CTA needed (horizon %d x %d-cell footprints)", smem, horizon, cells);
        }
    }
    kern<<<(n_jobs + kRolloutWarps - 1) / kRolloutWarps, kRolloutWarps * 32, smem, e->stream>>>(rp);
    e->launches++;
    CU(e, cudaGetLastError());
    return IPP_OK;
}

extern "C" int ipp_rollout(ipp_engine *e, int32_t n_jobs, int32_t horizon, const int32_t *env_index, const int32_t *path_action_ids,
                           const double *prev_poses, float *rewards, uint32_t flags) {
    if (!e || !path_action_ids || !rewards) return IPP_ERR_INVALID;
    if (n_jobs <= 0 || horizon < 1) return n_jobs == 0 ? IPP_OK : fail(e, IPP_ERR_INVALID, "ipp_rollout: bad n_jobs / horizon");
    const size_t J = (size_t)n_jobs, JH = J * (size_t)horizon;
    if (env_index)
        for (size_t j = 0; j < J; ++j)
            if (env_index[j] < 0 || env_index[j] >= e->cfg.batch) return fail(e, IPP_ERR_INVALID, "ipp_rollout: env_index[%zu] = %d outside batch", j, env_index[j]);
    int rc;
    if ((rc = ensure_job_buffers(e, std::max(JH, (size_t)e->cfg.batch))) != IPP_OK) return rc;
    if (env_index) CU(e, cudaMemcpyAsync(e->d_env_index, env_index, J * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaMemcpyAsync(e->d_actions, path_action_ids, JH * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    if (prev_poses) CU(e, cudaMemcpyAsync(e->d_prev_in, prev_poses, 3 * J * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    rc = ipp_rollout_device(e, n_jobs, horizon, env_index ? e->d_env_index : nullptr, e->d_actions, prev_poses ? e->d_prev_in : nullptr,
                            e->d_reward, flags);
    if (rc != IPP_OK) return rc;
    CU(e, cudaMemcpyAsync(rewards, e->d_reward, JH * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    return check_status(e);
}

// engine_internal.h: what the tree-search translation unit needs from the engine
int ipp_internal_step_params(const ipp_engine *e, ipp::StepParams *out) {
    if (!e || !out) return IPP_ERR_INVALID;
    fill_params(e, *out);
    return IPP_OK;
}
cudaStream_t ipp_internal_stream(const ipp_engine *e) { return e->stream; }
void ipp_internal_count_launches(ipp_engine *e, int n) { e->launches += (uint64_t)n; }
int ipp_internal_fail(ipp_engine *e, int code, const char *msg) { return fail(e, code, "%s", msg); }
int ipp_internal_layout(const ipp_engine *e) { return e->cfg.layout; }

extern "C" int ipp_eval_device(ipp_engine *e, float *metrics) {
    if (!e || !metrics) return IPP_ERR_INVALID;
    eval_kernel<<<e->cfg.batch, kEvalThreads, 0, e->stream>>>(e->d_mean, e->d_var, e->d_gt, e->cfg.layout, tiled_dims(e),
                                                              (float)e->cfg.value_threshold, metrics);
    e->launches++;
    CU(e, cudaGetLastError());
    return IPP_OK;
}

extern "C" int ipp_eval(ipp_engine *e, float *metrics) {
    if (!e || !metrics) return IPP_ERR_INVALID;
    int rc = ipp_eval_device(e, e->d_metrics);
    if (rc != IPP_OK) return rc;
    CU(e, cudaMemcpyAsync(metrics, e->d_metrics, (size_t)e->cfg.batch * IPP_NUM_METRICS * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return IPP_OK;
}

extern "C" void *ipp_device_ptr(ipp_engine *e, int32_t which) {
    if (!e) return nullptr;
    switch (which) {
        case IPP_PTR_MEAN: return e->d_mean;
        case IPP_PTR_VAR: return (e->cfg.layout != IPP_LAYOUT_PLANES && e->cfg.layout != IPP_LAYOUT_SPLIT) ? (void *)(e->d_mean + 1) : (void *)e->d_var;
        case IPP_PTR_GT: return e->d_gt;
        case IPP_PTR_REWARD:
            if (ensure_job_buffers(e, (size_t)e->cfg.batch) != IPP_OK) return nullptr;
            return e->d_reward;
        case IPP_PTR_STREAM: return (void *)e->stream;
        default: return nullptr;
    }
}

extern "C" int ipp_set_option(ipp_engine *e, int32_t option, int64_t value) {
    if (!e) return IPP_ERR_INVALID;
    switch (option) {
        case IPP_OPT_STEP_PATH:
            if (value < IPP_PATH_LSU || value > IPP_PATH_ASYNC) return fail(e, IPP_ERR_INVALID, "ipp_set_option: step path %lld unknown", (long long)value);
            e->step_path = (int)value;
            return IPP_OK;
        case IPP_OPT_ZERO_COPY:
            if (value < 0 || value > (IPP_ZERO_COPY_REWARDS | IPP_ZERO_COPY_IDS | IPP_ZERO_COPY_IDS_FETCH))
                return fail(e, IPP_ERR_INVALID, "ipp_set_option: zero-copy mask %lld unknown", (long long)value);
            e->zero_copy = (int)value;
            return IPP_OK;
        default:
            return fail(e, IPP_ERR_INVALID, "ipp_set_option: unknown option %d", option);
    }
}

extern "C" int64_t ipp_get_option(const ipp_engine *e, int32_t option) {
    if (!e) return -1;
    switch (option) {
        case IPP_OPT_STEP_PATH: return effective_path(e);
        case IPP_OPT_LAUNCHES_LSU: return (int64_t)e->path_launches[IPP_PATH_LSU];
        case IPP_OPT_LAUNCHES_ASYNC: return (int64_t)e->path_launches[IPP_PATH_ASYNC];
        case IPP_OPT_ZERO_COPY: return e->zero_copy;
        case IPP_OPT_ZERO_COPY_STEPS: return (int64_t)e->zero_copy_steps;
        case IPP_OPT_IDS_FETCH_STEPS: return (int64_t)e->ids_fetched_steps;
        default: return -1;
    }
}

extern "C" int ipp_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return IPP_ERR_INVALID;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? IPP_OK : IPP_ERR_NOMEM;
}
extern "C" int ipp_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? IPP_OK : IPP_ERR_CUDA; }
