// fields.cu — on-device reset for a whole env batch (sm_100a): the piecewise-constant ground truths and the shuffled priors.
//
// Reference:
//   HotspotRandomField.create_ground_truth_map   simulations/simulations.py:57-92   two square hot spots on a low background
//   SplitRandomField.create_ground_truth_map     simulations/simulations.py:102-125 a high and a low half along a row / column
//   Mapping.init_priors with shuffle_prior_cov   mapping/mappings.py:219-240        per-env prior (co)variance scale
// The reference draws from NumPy's global MT19937 (uniform / randint / rand / normal) once per env on the host; here every
// env draws the same SEQUENCE OF DECISIONS from Philox4x32-10 keyed by (seed; global env id, draw index), so a batch is
// generated in one launch and an env's world does not depend on how the batch is sharded.  Parity is statistical (same
// distributions, same construction — the host twins in ipp_rl_b200/simulations are pinned bit for bit against the reference).
#include <cmath>

#include "engine_internal.h"

using namespace ipp;

namespace {

// u-th uniform in [0, 1) of env `env` (draw indices are consumed in the reference's call order)
__device__ __forceinline__ float draw_u(uint32_t seed_lo, uint32_t seed_hi, uint32_t env, uint32_t stream, uint32_t k) {
    uint32_t r[4];
    philox4x32_10(k >> 2, env, stream, 0x1f1e1d5u, seed_lo, seed_hi, r);
    return (float)(r[k & 3] >> 8) * (1.0f / 16777216.0f);
}
// np.random.randint(low, high): low + floor(u * (high - low)), clamped
__device__ __forceinline__ int draw_int(float u, int low, int high) { return high > low ? min(high - 1, low + (int)(u * (float)(high - low))) : low; }

constexpr int kFieldThreads = 256;
constexpr int kMaxTries = 64;  // second hot-spot centre: rejection loop of the reference, bounded here

// one CTA per env
__global__ void __launch_bounds__(kFieldThreads) field_kernel(float *gt, int kind, int radius, int X, int Y, size_t plane_gt, int txg, int ts_gt,
                                                              int gw_shift, uint32_t seed_lo, uint32_t seed_hi, uint32_t env0) {
    __shared__ float s_val[2];
    __shared__ int s_rect[8];
    const int env = blockIdx.x;
    if (threadIdx.x == 0) {
        const uint32_t id = env0 + (uint32_t)env;
        uint32_t k = 0;
        auto u = [&]() { return draw_u(seed_lo, seed_hi, id, (uint32_t)kind, k++); };
        if (kind == IPP_FIELD_HOTSPOT) {
            s_val[0] = 0.7f + 0.3f * u();  // high_interest_value ~ U(0.7, 1)
            s_val[1] = 0.3f * u();         // low_interest_value  ~ U(0, 0.3)
            const int cy = draw_int(u(), radius, Y), cx = draw_int(u(), radius, X);
            s_rect[0] = max(cy - radius, 0), s_rect[1] = min(cy + radius, Y), s_rect[2] = max(cx - radius, 0), s_rect[3] = min(cx + radius, X);
            int ty = cy, tx = cx;
            bool found = false;
            for (int tries = 0; tries < kMaxTries && !found; ++tries) {
                ty = draw_int(u(), radius, Y);
                tx = draw_int(u(), radius, X);
                found = !(abs(ty - cy) <= radius || abs(tx - cx) <= radius);
            }
            if (found) {
                s_rect[4] = max(ty - radius, 0), s_rect[5] = min(ty + radius, Y), s_rect[6] = max(tx - radius, 0), s_rect[7] = min(tx + radius, X);
            } else {  // grids too small for a second, clear hot spot (the reference would loop forever): keep one
                s_rect[4] = s_rect[5] = s_rect[6] = s_rect[7] = 0;
            }
        } else {  // IPP_FIELD_SPLIT
            const float high = 0.65f + 0.35f * u(), low = 0.35f * u();
            const bool swap = u() > 0.5f;
            s_val[0] = swap ? low : high;  // first_value
            s_val[1] = swap ? high : low;  // second_value
            if (u() > 0.5f) {              // split along y: rows [0, cut) first, the rest second
                const int lo = (int)ceilf((float)Y * 0.33f), hi = (int)ceilf((float)Y * 0.66f);
                s_rect[0] = 0, s_rect[1] = draw_int(u(), lo, hi + 1), s_rect[2] = 0, s_rect[3] = X;
            } else {
                const int lo = (int)floorf((float)X * 0.33f), hi = (int)ceilf((float)X * 0.66f);
                s_rect[0] = 0, s_rect[1] = Y, s_rect[2] = 0, s_rect[3] = draw_int(u(), lo, hi + 1);
            }
            s_rect[4] = s_rect[5] = s_rect[6] = s_rect[7] = 0;
        }
    }
    __syncthreads();
    const float inside = s_val[0], outside = s_val[1];
    float *g = gt + (size_t)env * plane_gt;
    const size_t plane = (size_t)X * Y;
    for (size_t i = threadIdx.x; i < plane; i += kFieldThreads) {
        const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
        const bool in = (R >= s_rect[0] && R < s_rect[1] && C >= s_rect[2] && C < s_rect[3]) ||
                        (R >= s_rect[4] && R < s_rect[5] && C >= s_rect[6] && C < s_rect[7]);
        g[txg > 0 ? tiled_gt_index_rt(txg, ts_gt, gw_shift, R, C) : i] = in ? inside : outside;
    }
}

// per-env prior scale of Mapping.init_priors(shuffle_prior_cov=True)
__global__ void shuffled_prior_kernel(float *out /* [n][2]: variance level, relative spread */, int n, int gp_mode, float p0, float p1, int n_cells,
                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t env0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = env0 + (uint32_t)i;
    if (gp_mode) {  // signal_variance ~ U(0.8 sv, 1.2 sv): the Matern prior's diagonal (mappings.py:238-239)
        out[2 * i] = p0 * (0.8f + 0.4f * draw_u(seed_lo, seed_hi, id, 7u, 0));
        out[2 * i + 1] = 0.0f;
    } else {  // prior_cov_mean ~ U(0.1, prior_cov_mean), prior_cov_std = prior_cov_mean (mappings.py:221-223)
        const float mu = 0.1f + (p0 - 0.1f) * draw_u(seed_lo, seed_hi, id, 7u, 0);
        (void)p1;
        const float m2 = 2.0f * mu * mu;  // E[a^2] with sd = mu
        // diag(A A^T) / ||A||_F, A ~ N(mu, mu) of size N x N: row sums of squares are ~ N m2 +- sqrt(N (2 s^4 + 4 mu^2 s^2)),
        // the Frobenius norm ~ N sqrt(m2) (its own fluctuation is O(1/N)): level sqrt(m2), relative spread sqrt(6 mu^4 / N) / m2
        out[2 * i] = sqrtf(m2);
        out[2 * i + 1] = sqrtf(6.0f * mu * mu * mu * mu / (float)n_cells) / m2;
    }
}

}  // namespace

extern "C" int ipp_generate_field(ipp_engine *e, int32_t kind, int32_t cluster_radius, uint64_t seed, int32_t first_env, int32_t n_env) {
    if (!e) return IPP_ERR_INVALID;
    StepParams p;
    ipp_internal_step_params(e, &p);
    if (kind != IPP_FIELD_HOTSPOT && kind != IPP_FIELD_SPLIT) return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_generate_field: unknown field kind");
    if (first_env < 0 || n_env < 0 || first_env + n_env > p.batch) return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_generate_field: env range outside the batch");
    if (cluster_radius < 0) return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_generate_field: cluster_radius < 0");
    if (kind == IPP_FIELD_HOTSPOT && (cluster_radius >= p.X || cluster_radius >= p.Y))
        return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_generate_field: cluster_radius must be smaller than the grid");
    if (n_env == 0) return IPP_OK;
    cudaStream_t stream = ipp_internal_stream(e);
    const int tiled = ipp_internal_layout(e) == IPP_LAYOUT_TILED || ipp_internal_layout(e) == IPP_LAYOUT_SUPER || ipp_internal_layout(e) == IPP_LAYOUT_SPLIT;
    field_kernel<<<n_env, kFieldThreads, 0, stream>>>(const_cast<float *>(p.gt) + (size_t)first_env * p.plane_gt, kind, cluster_radius, p.X, p.Y, p.plane_gt,
                                                      tiled ? p.txg : 0, p.ts_gt, p.gw_shift, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32),
                                                      p.env_id_offset + (uint32_t)first_env);
    ipp_internal_count_launches(e, 1);
    if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_generate_field: CUDA failure");
    return IPP_OK;
}

// shared with ipp_engine.cu (ipp_reset_shuffled): fills scale[n][2] on the engine's stream
int ipp_internal_shuffled_prior(ipp_engine *e, float *scale, int n, int gp_mode, float p0, float p1, uint64_t seed) {
    StepParams p;
    ipp_internal_step_params(e, &p);
    shuffled_prior_kernel<<<(n + 127) / 128, 128, 0, ipp_internal_stream(e)>>>(scale, n, gp_mode, p0, p1, p.X * p.Y, (uint32_t)(seed & 0xffffffffu),
                                                                                (uint32_t)(seed >> 32), p.env_id_offset);
    ipp_internal_count_launches(e, 1);
    return cudaGetLastError() == cudaSuccess ? IPP_OK : IPP_ERR_CUDA;
}
