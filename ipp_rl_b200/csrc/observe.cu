// observe.cu — observation (network-input) planes straight from the belief in HBM (sm_100a).
//
// Reference: generate_input_feature_planes (planning/common/features.py:83-151) builds, per history entry, N x N
// planes from the dense covariance: [min-max normalised state, x, y, z position planes, budget plane] and one
// N x N action-cost plane (:61-71).  A per-cell engine has no N x N state; this kernel emits the (y_dim, x_dim)
// restriction of each of those planes for the CURRENT belief, in NCHW order, one CTA per env:
//
//   0  variance / max(variance)      = the diagonal of min_max_normalize(state) for a diagonal state (whose N x N
//                                       minimum is the off-diagonal 0); cells failing the adaptive mask
//                                       (mean + kappa * var >= threshold, rewards.py:8-12) are zeroed first (:94-99)
//   1  x / (x_dim * res)    2  y / (x_dim * res)  [sic, features.py:51]    3  (h - min_alt) / (max_alt - min_alt)
//   4  remaining budget / initial budget
//   5  (optional) cost from the current pose, dropped to min_altitude, to every cell centre at min_altitude,
//      min-max normalised = any column of the reference's cost plane, re-indexed from action id to (row, col)
//
// History (input_history_length > 1) is the caller's ring buffer of these outputs.  HBM traffic: 8 B/cell read
// (twice: max, then write pass) + 4 B/cell/plane written.
#include <cmath>

#include <cstdlib>

#include "engine_internal.h"

using namespace ipp;

namespace {

constexpr int kObsThreads = 256;

struct ObsParams {
    StepParams sp;
    int layout;
    int first_env;
    int planes;  // 5 or 6
    int adaptive;
    double min_alt, max_alt;
    const double *poses;   // [n][3] or nullptr -> sp.prev_state of the env
    const float *budgets;  // [n] remaining / initial
    float *out;            // [n][planes][Y][X]
};

__device__ __forceinline__ float block_max(float v, float *smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = smem[0];
    for (int w = 1; w < kObsThreads / 32; ++w) r = fmaxf(r, smem[w]);
    __syncthreads();
    return r;
}

// Vector path (x_dim % 4 == 0, maps up to 1024 cells per side): a thread handles four consecutive cells of a row — in every
// layout those are contiguous in the belief arrays (a tile row is four cells) — with 16-byte loads and 16-byte stores into the
// dense output planes; the fp64 part of the cost plane (cell centre - pose, per column and per row) is tabulated once per env in
// shared memory, so that a cell costs a handful of fp32 instructions.  Same expressions as the generic kernel below: same bits.
constexpr int kObsMaxDim = 1024;
__global__ void __launch_bounds__(kObsThreads) observe_vec_kernel(const __grid_constant__ ObsParams op) {
    __shared__ float smem[kObsThreads / 32];
    __shared__ float s_dx[kObsMaxDim], s_dy[kObsMaxDim];
    const StepParams &p = op.sp;
    const int j = blockIdx.x, env = op.first_env + j;
    const int X = p.X, Y = p.Y, X4 = X >> 2;
    const size_t N = (size_t)X * Y;
    const int layout = op.layout;
    const bool costs = op.planes > 5;
    const float *mean_pl = p.mean + (size_t)env * (layout == IPP_LAYOUT_SPLIT ? p.plane_gt : p.plane), *var_pl = p.var + (size_t)env * p.plane;
    const float2 *mv = reinterpret_cast<const float2 *>(p.mean) + (size_t)env * p.plane;
    const double *pose = op.poses ? op.poses + 3 * (size_t)j : p.prev_state + 3 * (size_t)env;
    const double px = pose[0], py = pose[1], ph = pose[2];
    if (costs) {  // job_dist()'s fp64 differences, per column / per row
        for (int c = threadIdx.x; c < X; c += kObsThreads) s_dx[c] = (float)(__dadd_rn(__dmul_rn(p.res, (double)c), __dmul_rn(0.5, p.res)) - px);
        for (int r = threadIdx.x; r < Y; r += kObsThreads) s_dy[r] = (float)(__dadd_rn(__dmul_rn(p.res, (double)r), __dmul_rn(0.5, p.res)) - py);
    }
    __syncthreads();
    const float dz = (float)(op.min_alt - op.min_alt);
    // four cells (R, C0 .. C0 + 3): variances (masked in adaptive mode)
    auto load4 = [&](int R, int C0, float (&v)[4]) {
        float m[4] = {0.f, 0.f, 0.f, 0.f};
        if (layout == IPP_LAYOUT_PLANES) {
            const float4 t = *reinterpret_cast<const float4 *>(var_pl + (size_t)R * X + C0);
            v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
            if (op.adaptive) {
                const float4 u = *reinterpret_cast<const float4 *>(mean_pl + (size_t)R * X + C0);
                m[0] = u.x, m[1] = u.y, m[2] = u.z, m[3] = u.w;
            }
        } else if (layout == IPP_LAYOUT_SPLIT) {
            const int k = split_index(p.txm, R, C0);
            const float4 t = *reinterpret_cast<const float4 *>(var_pl + k);
            v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
            if (op.adaptive) {
                const float4 u = *reinterpret_cast<const float4 *>(mean_pl + split_mean_of(k));
                m[0] = u.x, m[1] = u.y, m[2] = u.z, m[3] = u.w;
            }
        } else {
            const size_t k = layout == IPP_LAYOUT_MV ? (size_t)R * X + C0 : tiled_mv_index_rt(p.txm, p.ts_mv, R, C0);
            const float4 t0 = *reinterpret_cast<const float4 *>(mv + k), t1 = *reinterpret_cast<const float4 *>(mv + k + 2);
            m[0] = t0.x, v[0] = t0.y, m[1] = t0.z, v[1] = t0.w, m[2] = t1.x, v[2] = t1.y, m[3] = t1.z, v[3] = t1.w;
        }
        if (op.adaptive) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (!(fmaf(p.kappa, v[c], m[c]) >= p.thr)) v[c] = 0.0f;
        }
    };
    auto cost_at = [&](int R, int C) {
        const float dx = s_dx[C], dy = s_dy[R];
        return job_cost_from_dist(p, fast_sqrt(fmaf(dx, dx, fmaf(dy, dy, dz * dz))));
    };
    const int nvec = Y * X4;
    // pass 1: max of the (masked) variance; extremes of the cost plane
    float vmax = 0.0f, cmax = 0.0f, cmin_neg = -INFINITY;
    for (int i = threadIdx.x; i < nvec; i += kObsThreads) {
        const int R = i / X4, C0 = (i - R * X4) << 2;
        float v[4];
        load4(R, C0, v);
        vmax = fmaxf(fmaxf(vmax, fmaxf(v[0], v[1])), fmaxf(v[2], v[3]));
        if (costs) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float cst = cost_at(R, C0 + c);
                cmax = fmaxf(cmax, cst);
                cmin_neg = fmaxf(cmin_neg, -cst);
            }
        }
    }
    vmax = block_max(vmax, smem);
    if (costs) {
        cmax = block_max(cmax, smem);
        cmin_neg = block_max(cmin_neg, smem);
    }
    const float cmin = -cmin_neg;
    const float xs = (float)(px / ((double)X * p.res)), ys = (float)(py / ((double)X * p.res));
    const float zs = (float)((ph - op.min_alt) / (op.max_alt - op.min_alt));
    const float bs = op.budgets ? op.budgets[j] : 1.0f;
    float *o = op.out + (size_t)j * op.planes * N;
    const float4 x4 = make_float4(xs, xs, xs, xs), y4 = make_float4(ys, ys, ys, ys), z4 = make_float4(zs, zs, zs, zs), b4 = make_float4(bs, bs, bs, bs);
    for (int i = threadIdx.x; i < nvec; i += kObsThreads) {
        const int R = i / X4, C0 = (i - R * X4) << 2;
        float v[4];
        load4(R, C0, v);  // the env's map is still in L2 from pass 1
        const size_t at = (size_t)R * X + C0;
        *reinterpret_cast<float4 *>(o + at) = make_float4(v[0] / vmax, v[1] / vmax, v[2] / vmax, v[3] / vmax);
        *reinterpret_cast<float4 *>(o + N + at) = x4;
        *reinterpret_cast<float4 *>(o + 2 * N + at) = y4;
        *reinterpret_cast<float4 *>(o + 3 * N + at) = z4;
        *reinterpret_cast<float4 *>(o + 4 * N + at) = b4;
        if (costs) {
            float c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float cst = cost_at(R, C0 + k);
                c[k] = cmax > cmin ? (cst - cmin) / (cmax - cmin) : cst / cmax;
            }
            *reinterpret_cast<float4 *>(o + 5 * N + at) = make_float4(c[0], c[1], c[2], c[3]);
        }
    }
}

// Generic path (any x_dim): one cell per thread and pass.
__global__ void __launch_bounds__(kObsThreads) observe_kernel(const __grid_constant__ ObsParams op) {
    __shared__ float smem[kObsThreads / 32];
    const StepParams &p = op.sp;
    const int j = blockIdx.x, env = op.first_env + j;
    const int X = p.X, Y = p.Y;
    const size_t N = (size_t)X * Y;
    const bool planes_layout = op.layout == IPP_LAYOUT_PLANES, tiled = op.layout == IPP_LAYOUT_TILED || op.layout == IPP_LAYOUT_SUPER;
    const bool split = op.layout == IPP_LAYOUT_SPLIT;
    const float *mean_pl = p.mean + (size_t)env * (split ? p.plane_gt : p.plane), *var_pl = p.var + (size_t)env * p.plane;
    const float2 *mv = reinterpret_cast<const float2 *>(p.mean) + (size_t)env * p.plane;
    auto load = [&](size_t i, float &m, float &v) {
        if (planes_layout) {
            m = mean_pl[i];
            v = var_pl[i];
        } else if (split) {
            const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
            const int k = split_index(p.txm, R, C);
            m = mean_pl[split_mean_of(k)];
            v = var_pl[k];
        } else {
            const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
            const float2 t = mv[tiled ? tiled_mv_index_rt(p.txm, p.ts_mv, R, C) : i];
            m = t.x;
            v = t.y;
        }
    };
    const double *pose = op.poses ? op.poses + 3 * (size_t)j : p.prev_state + 3 * (size_t)env;
    const double px = pose[0], py = pose[1], ph = pose[2];

    // pass 1: max of the (masked) variance; extremes of the cost plane
    float vmax = 0.0f, cmax = 0.0f, cmin_neg = -INFINITY;
    for (size_t i = threadIdx.x; i < N; i += kObsThreads) {
        float m, v;
        load(i, m, v);
        if (op.adaptive && !(fmaf(p.kappa, v, m) >= p.thr)) v = 0.0f;
        vmax = fmaxf(vmax, v);
        if (op.planes > 5) {
            const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
            const double cx = __dadd_rn(__dmul_rn(p.res, (double)C), __dmul_rn(0.5, p.res));
            const double cy = __dadd_rn(__dmul_rn(p.res, (double)R), __dmul_rn(0.5, p.res));
            const float c = job_cost(p, cx, cy, op.min_alt, px, py, op.min_alt);
            cmax = fmaxf(cmax, c);
            cmin_neg = fmaxf(cmin_neg, -c);
        }
    }
    vmax = block_max(vmax, smem);
    if (op.planes > 5) {
        cmax = block_max(cmax, smem);
        cmin_neg = block_max(cmin_neg, smem);
    }
    const float cmin = -cmin_neg;
    // min_max_normalize (features.py:74-81): min == max -> x / max
    const float xs = (float)(px / ((double)X * p.res)), ys = (float)(py / ((double)X * p.res));
    const float zs = (float)((ph - op.min_alt) / (op.max_alt - op.min_alt));
    const float bs = op.budgets ? op.budgets[j] : 1.0f;
    float *o = op.out + (size_t)j * op.planes * N;
    for (size_t i = threadIdx.x; i < N; i += kObsThreads) {
        float m, v;
        load(i, m, v);
        if (op.adaptive && !(fmaf(p.kappa, v, m) >= p.thr)) v = 0.0f;
        o[i] = v / vmax;  // the N x N minimum of a diagonal state is the off-diagonal 0
        o[N + i] = xs;
        o[2 * N + i] = ys;
        o[3 * N + i] = zs;
        o[4 * N + i] = bs;
        if (op.planes > 5) {
            const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
            const double cx = __dadd_rn(__dmul_rn(p.res, (double)C), __dmul_rn(0.5, p.res));
            const double cy = __dadd_rn(__dmul_rn(p.res, (double)R), __dmul_rn(0.5, p.res));
            const float c = job_cost(p, cx, cy, op.min_alt, px, py, op.min_alt);
            o[5 * N + i] = cmax > cmin ? (c - cmin) / (cmax - cmin) : c / cmax;
        }
    }
}

}  // namespace

extern "C" int ipp_observe(ipp_engine *e, int32_t first_env, int32_t n_env, const double *poses, const float *budget_ratio, uint32_t flags,
                           float *out, int32_t out_is_device) {
    if (!e || !out) return IPP_ERR_INVALID;
    ObsParams op;
    ipp_internal_step_params(e, &op.sp);
    if (first_env < 0 || n_env < 0 || first_env + n_env > op.sp.batch) return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_observe: env range outside the batch");
    if (n_env == 0) return IPP_OK;
    cudaStream_t stream = ipp_internal_stream(e);
    op.layout = ipp_internal_layout(e);
    op.first_env = first_env;
    op.planes = (flags & IPP_OBS_COSTS) ? 6 : 5;
    op.adaptive = (flags & IPP_FLAG_ADAPTIVE) ? 1 : 0;
    op.min_alt = op.sp.lut[0].alt;
    op.max_alt = op.sp.lut[op.sp.n_levels - 1].alt;
    const size_t N = (size_t)op.sp.X * op.sp.Y, total = (size_t)n_env * op.planes * N;
    double *d_poses = nullptr;
    float *d_budget = nullptr, *d_out = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_poses);
        cudaFree(d_budget);
        if (!out_is_device) cudaFree(d_out);
    };
    bool ok = true;
    if (poses) {
        ok = ok && cudaMalloc((void **)&d_poses, 3 * (size_t)n_env * sizeof(double)) == cudaSuccess;
        if (ok) cudaMemcpyAsync(d_poses, poses, 3 * (size_t)n_env * sizeof(double), cudaMemcpyHostToDevice, stream);
    }
    if (budget_ratio) {
        ok = ok && cudaMalloc((void **)&d_budget, (size_t)n_env * sizeof(float)) == cudaSuccess;
        if (ok) cudaMemcpyAsync(d_budget, budget_ratio, (size_t)n_env * sizeof(float), cudaMemcpyHostToDevice, stream);
    }
    if (out_is_device)
        d_out = out;
    else
        ok = ok && cudaMalloc((void **)&d_out, total * sizeof(float)) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        cleanup();
        return ipp_internal_fail(e, IPP_ERR_NOMEM, "ipp_observe: staging allocation failed");
    }
    op.poses = d_poses;
    op.budgets = d_budget;
    op.out = d_out;
    // 16-byte path: rows of whole 4-cell groups, output planes 16-byte aligned (cudaMalloc'ed or caller-aligned)
    if (op.sp.X % 4 == 0 && op.sp.X <= kObsMaxDim && op.sp.Y <= kObsMaxDim && ((uintptr_t)d_out & 15) == 0 && getenv("IPP_OBS_GENERIC") == nullptr)
        observe_vec_kernel<<<n_env, kObsThreads, 0, stream>>>(op);
    else
        observe_kernel<<<n_env, kObsThreads, 0, stream>>>(op);
    ipp_internal_count_launches(e, 1);
    if (!out_is_device) cudaMemcpyAsync(out, d_out, total * sizeof(float), cudaMemcpyDeviceToHost, stream);
    const cudaError_t s = cudaStreamSynchronize(stream);
    cleanup();
    if (s != cudaSuccess || cudaGetLastError() != cudaSuccess) return ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_observe: CUDA failure");
    return IPP_OK;
}
