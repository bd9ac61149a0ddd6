// experience.cu — device-resident experience ring (include/ipp_experience.h), sm_100a.
//
// Reference data path replaced (all file:line into the reference tree):
//   value targets            planning/mcts_zero/episode_generators.py:158-164, planning/common/rewards.py:34-35
//   one bz2 pickle / sample  episode_generators.py:186-192  ->  rows of a ring in HBM (plain stream copies)
//   uniform / prioritised sampling, importance weights, priority update   planning/mcts_zero/replay_buffers.py:83-141
//   random-shift augmentation (ReplicationPad2d(4) + RandomCrop)          replay_buffers.py:58-77
//
// Kernels: value_targets_kernel (one thread per (episode, step)), fill / scatter of priorities, the sampling chain
// pow -> inclusive scan (cub) -> search_kernel (inverse CDF + importance weights, one CTA: n is a training batch),
// gather_obs_kernel (row gather with clamped shifts; HBM-bound: 4 B read + 4 B written per element) and
// gather_rows_kernel (policy / mask / scalars).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <string>

#include "../../include/ipp_experience.h"

namespace {

constexpr int kThreads = 256;

struct GatherParams {
    const float *obs;      // ring [capacity][C][Y][X]
    const int64_t *idx;    // [n]
    const int8_t *shifts;  // [n][2] or nullptr
    float *out;            // [n][C][Y][X]
    int n, C, Y, X;
};

}  // namespace

struct ipp_ring {
    ipp_ring_config cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t obs_floats = 0;  // C*Y*X
    float *d_obs = nullptr, *d_policy = nullptr, *d_value = nullptr, *d_reward = nullptr, *d_priority = nullptr;
    uint8_t *d_mask = nullptr;
    double *d_w = nullptr, *d_cdf = nullptr;  // [capacity] priorities^alpha and their inclusive scan
    void *d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    float *d_max = nullptr;  // [1] running maximum priority
    // per-call staging (grown on demand)
    int64_t *d_idx = nullptr;
    float *d_weights = nullptr;
    double *d_uniform = nullptr;
    int8_t *d_shifts = nullptr;
    size_t cap_n = 0;
    unsigned char *d_stage = nullptr;  // outputs / inputs of host callers
    size_t cap_stage = 0;
    int64_t size = 0, head = 0;
    uint64_t pushed = 0, draws = 0, launches = 0, device_bytes = 0;
    int last_n = 0;
    std::string err;
};

static thread_local std::string g_ring_create_err;

static int rfail(ipp_ring *r, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    (r ? r->err : g_ring_create_err) = buf;
    return code;
}

#define RCU(r, call)                                                                                                  \
    do {                                                                                                              \
        cudaError_t _s = (call);                                                                                      \
        if (_s != cudaSuccess)                                                                                        \
            return rfail((r), _s == cudaErrorMemoryAllocation ? IPP_ERR_NOMEM : IPP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                         cudaGetErrorString(_s), __FILE__, __LINE__);                                                 \
    } while (0)

template <typename T>
static int ralloc(ipp_ring *r, T **p, size_t n) {
    RCU(r, cudaMalloc((void **)p, n * sizeof(T)));
    r->device_bytes += n * sizeof(T);
    return IPP_OK;
}

template <typename T>
static int rensure(ipp_ring *r, T **p, size_t *cap, size_t n) {
    if (*cap >= n && *p) return IPP_OK;
    if (*p) {
        RCU(r, cudaStreamSynchronize(r->stream));
        RCU(r, cudaFree(*p));
        r->device_bytes -= *cap * sizeof(T);
        *p = nullptr;
        *cap = 0;
    }
    int rc = ralloc(r, p, n);
    if (rc == IPP_OK) *cap = n;
    return rc;
}

static int ensure_n(ipp_ring *r, size_t n) {
    if (r->cap_n >= n) return IPP_OK;
    RCU(r, cudaStreamSynchronize(r->stream));
    void *old[] = {r->d_idx, r->d_weights, r->d_uniform, r->d_shifts};
    for (void *p : old)
        if (p) cudaFree(p);
    r->d_idx = nullptr, r->d_weights = nullptr, r->d_uniform = nullptr, r->d_shifts = nullptr;
    r->device_bytes -= r->cap_n * (sizeof(int64_t) + sizeof(float) + sizeof(double) + 2);
    r->cap_n = 0;
    int rc;
    if ((rc = ralloc(r, &r->d_idx, n)) != IPP_OK) return rc;
    if ((rc = ralloc(r, &r->d_weights, n)) != IPP_OK) return rc;
    if ((rc = ralloc(r, &r->d_uniform, n)) != IPP_OK) return rc;
    if ((rc = ralloc(r, &r->d_shifts, 2 * n)) != IPP_OK) return rc;
    r->cap_n = n;
    return IPP_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void value_targets_kernel(const float *__restrict__ rewards, const int32_t *__restrict__ lengths, int n_ep, int T, double gamma,
                                     int horizon, float *__restrict__ values, float *__restrict__ totals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ep * T) return;
    const int e = t / T, i = t - e * T;
    const int len = min(max(lengths[e], 0), T);
    const float *rw = rewards + (size_t)e * T;
    float v = 0.0f;
    if (i < len) {
        const int hi = min(i + horizon, len);
        double s = 0.0;  // Python's sum(): left to right in fp64, gamma ** j with the absolute step j
        for (int j = i; j < hi; ++j) s += pow(gamma, (double)j) * (double)rw[j];
        v = (float)(sqrt(s + 1.0) - 1.0);
    }
    values[t] = v;
    if (totals != nullptr && i == 0) {
        double s = 0.0;
        for (int j = 0; j < len; ++j) s += pow(gamma, (double)j) * (double)rw[j];
        totals[e] = (float)s;
    }
}

__global__ void fill_priority_kernel(float *prio, int64_t capacity, int64_t head, int n, float value, float *running_max, int size_before) {
    // value <= 0: the running maximum (1 when the ring was empty)
    const float v = value > 0.0f ? value : (size_before > 0 ? *running_max : 1.0f);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) prio[(head + k) % capacity] = v;
    if (blockIdx.x == 0 && threadIdx.x == 0) *running_max = size_before > 0 ? fmaxf(*running_max, v) : v;
}

__global__ void set_all_priorities_kernel(float *prio, int64_t size, float value, float *running_max) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < size; k += (int64_t)gridDim.x * blockDim.x) prio[k] = value;
    if (blockIdx.x == 0 && threadIdx.x == 0) *running_max = value;
}

__global__ void pow_kernel(const float *__restrict__ prio, int64_t size, double alpha, double *__restrict__ w) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < size; k += (int64_t)gridDim.x * blockDim.x)
        w[k] = pow((double)prio[k], alpha);
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

// One CTA: inverse-CDF draw + importance weights (n = a training batch).  cdf == nullptr: uniform over [0, size).
__global__ void __launch_bounds__(1024) search_kernel(const double *__restrict__ cdf, const double *__restrict__ w, int64_t size, int n,
                                                      double beta, const double *__restrict__ uniforms, uint64_t seed, uint64_t draw,
                                                      int64_t *__restrict__ idx_out, float *__restrict__ weights) {
    __shared__ double s_max[32];
    const double total = cdf ? cdf[size - 1] : 1.0;
    double my_max = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double u;
        if (uniforms) {
            u = uniforms[k];
        } else {  // 53 random bits -> [0, 1), as np.random.random_sample builds its doubles
            uint32_t rnd[4];
            philox4x32_10((uint32_t)k, (uint32_t)(draw & 0xffffffffu), (uint32_t)(draw >> 32), 0x45585052u, (uint32_t)(seed & 0xffffffffu),
                          (uint32_t)(seed >> 32), rnd);
            u = ((double)(rnd[0] >> 5) * 67108864.0 + (double)(rnd[1] >> 6)) * (1.0 / 9007199254740992.0);
        }
        int64_t lo = 0;
        double wt = 1.0;
        if (cdf) {
            int64_t hi = size;  // first i with cdf[i] / total > u  (searchsorted side='right' on the normalised cdf)
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (cdf[mid] / total <= u)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo > size - 1) lo = size - 1;
            wt = pow((w[lo] / total) * (double)size, -beta);
        } else {
            lo = (int64_t)(u * (double)size);
            if (lo > size - 1) lo = size - 1;
        }
        idx_out[k] = lo;
        if (weights) weights[k] = 1.0f;  // uniform draw; the prioritised weights are normalised below
        my_max = fmax(my_max, wt);
    }
    for (int o = 16; o > 0; o >>= 1) my_max = fmax(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = my_max;
    __syncthreads();
    double mx = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) mx = fmax(mx, s_max[k]);
    if (weights && cdf)
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            // idx_out[k] was written by this very thread; fp64 ratio rounded once to float32 as in replay_buffers.py:132
            const int64_t i = idx_out[k];
            const double wt = pow((w[i] / total) * (double)size, -beta);
            weights[k] = (float)(wt / mx);
        }
}

// out[s][c][y][x] = obs[idx[s]][c][clamp(y + dy)][clamp(x + dx)]; one thread per 4 consecutive x (X % 4 == 0) or per x.
template <bool VEC4>
__global__ void __launch_bounds__(kThreads) gather_obs_kernel(const GatherParams g) {
    const int XV = VEC4 ? g.X >> 2 : g.X;
    const size_t per_sample = (size_t)g.C * g.Y * XV;
    const size_t total = per_sample * g.n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(t / per_sample);
        size_t rem = t - (size_t)s * per_sample;
        const int xv = (int)(rem % XV);
        rem /= XV;
        const int y = (int)(rem % g.Y), c = (int)(rem / g.Y);
        int dy = 0, dx = 0;
        if (g.shifts) {
            dy = g.shifts[2 * s];
            dx = g.shifts[2 * s + 1];
        }
        const int ys = min(max(y + dy, 0), g.Y - 1);
        const float *src = g.obs + ((size_t)g.idx[s] * g.C + c) * ((size_t)g.Y * g.X) + (size_t)ys * g.X;
        float *dst = g.out + (((size_t)s * g.C + c) * g.Y + y) * (size_t)g.X;
        if (VEC4) {
            float4 v;
            if (dx == 0) {
                v = __ldcs(reinterpret_cast<const float4 *>(src) + xv);
            } else {
                const int x0 = 4 * xv + dx;
                v.x = src[min(max(x0, 0), g.X - 1)];
                v.y = src[min(max(x0 + 1, 0), g.X - 1)];
                v.z = src[min(max(x0 + 2, 0), g.X - 1)];
                v.w = src[min(max(x0 + 3, 0), g.X - 1)];
            }
            __stcs(reinterpret_cast<float4 *>(dst) + xv, v);
        } else {
            dst[xv] = src[min(max(xv + dx, 0), g.X - 1)];
        }
    }
}

template <typename T>
__global__ void gather_rows_kernel(const T *__restrict__ src, const int64_t *__restrict__ idx, int n, int width, T *__restrict__ out) {
    const size_t total = (size_t)n * width;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(t / width), k = (int)(t - (size_t)s * width);
        out[t] = src[(size_t)idx[s] * width + k];
    }
}

// priorities[idx[k]] = values[k]; of duplicate indices the LAST occurrence wins, as in NumPy's fancy assignment
// (replay_buffers.py:141) — decided by a forward scan (n is a training batch).
__global__ void scatter_priorities_kernel(float *prio, const int64_t *__restrict__ idx, const float *__restrict__ values, int n, int64_t size,
                                          float *running_max) {
    float m = 0.0f;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int64_t i = idx[k];
        if (i < 0 || i >= size) continue;
        bool last = true;
        for (int j = k + 1; j < n && last; ++j) last = idx[j] != i;
        if (last) {
            prio[i] = values[k];
            m = fmaxf(m, values[k]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<int *>(running_max), __float_as_int(m));  // positive floats order as ints
}

static inline int blocks_for(size_t work, int cap = 148 * 8) {
    size_t b = (work + kThreads - 1) / kThreads;
    return (int)(b < 1 ? 1 : (b > (size_t)cap ? (size_t)cap : b));
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" int ipp_ring_create(const ipp_ring_config *cfg, ipp_ring **out) {
    if (!cfg || !out) return rfail(nullptr, IPP_ERR_INVALID, "ipp_ring_create: NULL argument");
    *out = nullptr;
    if (cfg->struct_bytes != sizeof(ipp_ring_config)) return rfail(nullptr, IPP_ERR_INVALID, "ipp_ring_create: struct_bytes mismatch (header / library version skew)");
    if (cfg->capacity < 1 || cfg->channels < 1 || cfg->y_dim < 1 || cfg->x_dim < 1 || cfg->policy_slots < 1)
        return rfail(nullptr, IPP_ERR_INVALID, "ipp_ring_create: capacity, channels, dims and policy_slots must be >= 1");
    cudaError_t s = cudaSetDevice(cfg->device);
    if (s != cudaSuccess) return rfail(nullptr, IPP_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(s));
    ipp_ring *r = new ipp_ring();
    r->cfg = *cfg;
    r->obs_floats = (size_t)cfg->channels * cfg->y_dim * cfg->x_dim;
    auto bail = [&](int rc) {
        g_ring_create_err = r->err;
        ipp_ring_destroy(r);
        return rc;
    };
    if (cfg->stream) {
        r->stream = (cudaStream_t)cfg->stream;
    } else {
        s = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
        if (s != cudaSuccess) return bail(rfail(r, IPP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(s)));
        r->own_stream = true;
    }
    const size_t cap = (size_t)cfg->capacity, P = (size_t)cfg->policy_slots;
    int rc;
    if ((rc = ralloc(r, &r->d_obs, cap * r->obs_floats)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_policy, cap * P)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_mask, cap * P)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_value, cap)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_reward, cap)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_priority, cap)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_w, cap)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_cdf, cap)) != IPP_OK) return bail(rc);
    if ((rc = ralloc(r, &r->d_max, 1)) != IPP_OK) return bail(rc);
    cudaMemsetAsync(r->d_max, 0, sizeof(float), r->stream);
    cub::DeviceScan::InclusiveSum(nullptr, r->scan_tmp_bytes, r->d_w, r->d_cdf, (int)std::min<size_t>(cap, 0x7fffffff), r->stream);
    if ((rc = ralloc(r, (unsigned char **)&r->d_scan_tmp, r->scan_tmp_bytes + 16)) != IPP_OK) return bail(rc);
    *out = r;
    return IPP_OK;
}

extern "C" void ipp_ring_destroy(ipp_ring *r) {
    if (!r) return;
    if (r->stream) cudaStreamSynchronize(r->stream);
    void *ptrs[] = {r->d_obs, r->d_policy, r->d_mask, r->d_value, r->d_reward, r->d_priority, r->d_w, r->d_cdf, r->d_scan_tmp, r->d_max,
                    r->d_idx, r->d_weights, r->d_uniform, r->d_shifts, r->d_stage};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (r->own_stream && r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

extern "C" const char *ipp_ring_last_error(const ipp_ring *r) { return r ? r->err.c_str() : g_ring_create_err.c_str(); }

extern "C" int ipp_ring_get_info(const ipp_ring *r, ipp_ring_info *out) {
    if (!r || !out) return IPP_ERR_INVALID;
    out->capacity = r->cfg.capacity;
    out->size = r->size;
    out->head = r->head;
    out->pushed = r->pushed;
    out->device_bytes = r->device_bytes;
    out->launches = r->launches;
    return IPP_OK;
}

extern "C" int ipp_ring_value_targets(ipp_ring *r, const float *rewards, const int32_t *lengths, int32_t n_ep, int32_t T, double gamma,
                                      int32_t horizon, float *values, float *totals, int32_t is_device) {
    if (!r) return IPP_ERR_INVALID;
    if (!rewards || !lengths || !values) return rfail(r, IPP_ERR_INVALID, "ipp_ring_value_targets: NULL argument");
    if (n_ep < 0 || T < 1 || horizon < 1) return rfail(r, IPP_ERR_INVALID, "ipp_ring_value_targets: n_episodes >= 0, max_steps >= 1, horizon >= 1");
    if (n_ep == 0) return IPP_OK;
    const size_t nt = (size_t)n_ep * T;
    const float *d_rw = rewards;
    const int32_t *d_len = lengths;
    float *d_val = values, *d_tot = totals;
    if (!is_device) {
        // staging layout: rewards | values | totals | lengths
        const size_t bytes = nt * 4 * 2 + (size_t)n_ep * 8;
        int rc = rensure(r, &r->d_stage, &r->cap_stage, bytes);
        if (rc != IPP_OK) return rc;
        float *base = reinterpret_cast<float *>(r->d_stage);
        RCU(r, cudaMemcpyAsync(base, rewards, nt * 4, cudaMemcpyHostToDevice, r->stream));
        RCU(r, cudaMemcpyAsync(base + 2 * nt + n_ep, lengths, (size_t)n_ep * 4, cudaMemcpyHostToDevice, r->stream));
        d_rw = base;
        d_val = base + nt;
        d_tot = totals ? base + 2 * nt : nullptr;
        d_len = reinterpret_cast<const int32_t *>(base + 2 * nt + n_ep);
    }
    value_targets_kernel<<<(int)((nt + kThreads - 1) / kThreads), kThreads, 0, r->stream>>>(d_rw, d_len, n_ep, T, gamma, horizon, d_val, d_tot);
    r->launches++;
    RCU(r, cudaGetLastError());
    if (!is_device) {
        RCU(r, cudaMemcpyAsync(values, d_val, nt * 4, cudaMemcpyDeviceToHost, r->stream));
        if (totals) RCU(r, cudaMemcpyAsync(totals, d_tot, (size_t)n_ep * 4, cudaMemcpyDeviceToHost, r->stream));
        RCU(r, cudaStreamSynchronize(r->stream));
    }
    return IPP_OK;
}

// copy n rows of `width` elements into ring slots [head, head + n) (two segments when the range wraps)
template <typename T>
static int push_rows(ipp_ring *r, T *ring, const T *src, size_t width, int64_t head, int n, int fill_byte) {
    const int64_t cap = r->cfg.capacity;
    const int64_t first = std::min<int64_t>(n, cap - head);
    if (src) {
        RCU(r, cudaMemcpyAsync(ring + (size_t)head * width, src, (size_t)first * width * sizeof(T), cudaMemcpyDefault, r->stream));
        if (first < n)
            RCU(r, cudaMemcpyAsync(ring, src + (size_t)first * width, (size_t)(n - first) * width * sizeof(T), cudaMemcpyDefault, r->stream));
    } else {
        RCU(r, cudaMemsetAsync(ring + (size_t)head * width, fill_byte, (size_t)first * width * sizeof(T), r->stream));
        if (first < n) RCU(r, cudaMemsetAsync(ring, fill_byte, (size_t)(n - first) * width * sizeof(T), r->stream));
    }
    return IPP_OK;
}

extern "C" int ipp_ring_push(ipp_ring *r, int32_t n, const float *obs, const float *policy, const uint8_t *valid_mask, const float *values,
                             const float *rewards, float priority, int32_t is_device) {
    if (!r) return IPP_ERR_INVALID;
    if (n < 0 || n > r->cfg.capacity) return rfail(r, IPP_ERR_INVALID, "ipp_ring_push: n = %d outside [0, capacity]", n);
    if (n == 0) return IPP_OK;
    if (!obs || !values || !rewards) return rfail(r, IPP_ERR_INVALID, "ipp_ring_push: obs, values and rewards are required");
    (void)is_device;  // cudaMemcpyDefault resolves host / device sources under unified addressing
    const size_t P = (size_t)r->cfg.policy_slots;
    int rc;
    if ((rc = push_rows(r, r->d_obs, obs, r->obs_floats, r->head, n, 0)) != IPP_OK) return rc;
    if ((rc = push_rows(r, r->d_policy, policy, P, r->head, n, 0)) != IPP_OK) return rc;
    if ((rc = push_rows(r, r->d_mask, valid_mask, P, r->head, n, 1)) != IPP_OK) return rc;
    if ((rc = push_rows(r, r->d_value, values, 1, r->head, n, 0)) != IPP_OK) return rc;
    if ((rc = push_rows(r, r->d_reward, rewards, 1, r->head, n, 0)) != IPP_OK) return rc;
    fill_priority_kernel<<<blocks_for((size_t)n, 64), kThreads, 0, r->stream>>>(r->d_priority, r->cfg.capacity, r->head, n, priority, r->d_max,
                                                                                 (int)std::min<int64_t>(r->size, 0x7fffffff));
    r->launches++;
    RCU(r, cudaGetLastError());
    if (!is_device) RCU(r, cudaStreamSynchronize(r->stream));  // the caller may reuse its host buffers
    r->head = (r->head + n) % r->cfg.capacity;
    r->size = std::min<int64_t>(r->size + n, r->cfg.capacity);
    r->pushed += (uint64_t)n;
    return IPP_OK;
}

extern "C" int ipp_ring_reset_priorities(ipp_ring *r) {
    if (!r) return IPP_ERR_INVALID;
    if (r->size == 0) return IPP_OK;
    set_all_priorities_kernel<<<blocks_for((size_t)r->size), kThreads, 0, r->stream>>>(r->d_priority, r->size, (float)(1.0 / (double)r->size),
                                                                                        r->d_max);
    r->launches++;
    RCU(r, cudaGetLastError());
    return IPP_OK;
}

extern "C" int ipp_ring_sample(ipp_ring *r, int32_t n, double alpha, double beta, const double *uniforms, uint64_t seed, int64_t *indices,
                               float *weights, int32_t is_device) {
    if (!r) return IPP_ERR_INVALID;
    if (n < 1) return rfail(r, IPP_ERR_INVALID, "ipp_ring_sample: n must be >= 1");
    if (r->size == 0) return rfail(r, IPP_ERR_INVALID, "ipp_ring_sample: the ring is empty");
    int rc = ensure_n(r, (size_t)n);
    if (rc != IPP_OK) return rc;
    const double *d_u = nullptr;
    if (uniforms) {
        if (is_device) {
            d_u = uniforms;
        } else {
            RCU(r, cudaMemcpyAsync(r->d_uniform, uniforms, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, r->stream));
            d_u = r->d_uniform;
        }
    }
    const bool prioritised = alpha >= 0.0;
    if (prioritised) {
        pow_kernel<<<blocks_for((size_t)r->size), kThreads, 0, r->stream>>>(r->d_priority, r->size, alpha, r->d_w);
        size_t tmp = r->scan_tmp_bytes;
        RCU(r, cub::DeviceScan::InclusiveSum(r->d_scan_tmp, tmp, r->d_w, r->d_cdf, (int)r->size, r->stream));
        r->launches += 2;
    }
    search_kernel<<<1, 1024, 0, r->stream>>>(prioritised ? r->d_cdf : nullptr, r->d_w, r->size, n, beta, d_u, seed, r->draws, r->d_idx,
                                             r->d_weights);
    r->launches++;
    r->draws++;
    r->last_n = n;
    RCU(r, cudaGetLastError());
    const cudaMemcpyKind kind = is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (indices) RCU(r, cudaMemcpyAsync(indices, r->d_idx, (size_t)n * sizeof(int64_t), kind, r->stream));
    if (weights) RCU(r, cudaMemcpyAsync(weights, r->d_weights, (size_t)n * sizeof(float), kind, r->stream));
    if (!is_device) RCU(r, cudaStreamSynchronize(r->stream));
    return IPP_OK;
}

extern "C" int ipp_ring_gather(ipp_ring *r, int32_t n, const int64_t *indices, const int8_t *shifts, float *obs, float *policy,
                               uint8_t *valid_mask, float *values, float *rewards, int32_t is_device) {
    if (!r) return IPP_ERR_INVALID;
    if (n < 1) return rfail(r, IPP_ERR_INVALID, "ipp_ring_gather: n must be >= 1");
    if (!indices && r->last_n != n) return rfail(r, IPP_ERR_INVALID, "ipp_ring_gather: indices == NULL needs a preceding ipp_ring_sample of the same n");
    int rc = ensure_n(r, (size_t)n);
    if (rc != IPP_OK) return rc;
    const int64_t *d_idx = r->d_idx;
    if (indices) {
        if (is_device) {
            d_idx = indices;
        } else {
            for (int k = 0; k < n; ++k)
                if (indices[k] < 0 || indices[k] >= r->size) return rfail(r, IPP_ERR_INVALID, "ipp_ring_gather: index %lld outside [0, %lld)", (long long)indices[k], (long long)r->size);
            RCU(r, cudaMemcpyAsync(r->d_idx, indices, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, r->stream));
            r->last_n = n;
        }
    }
    const int8_t *d_sh = nullptr;
    if (shifts) {
        if (is_device) {
            d_sh = shifts;
        } else {
            RCU(r, cudaMemcpyAsync(r->d_shifts, shifts, 2 * (size_t)n, cudaMemcpyHostToDevice, r->stream));
            d_sh = r->d_shifts;
        }
    }
    const size_t P = (size_t)r->cfg.policy_slots;
    // host callers: outputs are produced in the staging buffer and copied out
    size_t off_obs = 0, off_pol = 0, off_mask = 0, off_val = 0, off_rw = 0, bytes = 0;
    if (!is_device) {
        auto take = [&](size_t b) {
            const size_t o = bytes;
            bytes += (b + 255) & ~(size_t)255;
            return o;
        };
        if (obs) off_obs = take((size_t)n * r->obs_floats * 4);
        if (policy) off_pol = take((size_t)n * P * 4);
        if (valid_mask) off_mask = take((size_t)n * P);
        if (values) off_val = take((size_t)n * 4);
        if (rewards) off_rw = take((size_t)n * 4);
        if ((rc = rensure(r, &r->d_stage, &r->cap_stage, bytes)) != IPP_OK) return rc;
    }
    float *o_obs = is_device ? obs : reinterpret_cast<float *>(r->d_stage + off_obs);
    float *o_pol = is_device ? policy : reinterpret_cast<float *>(r->d_stage + off_pol);
    uint8_t *o_mask = is_device ? valid_mask : r->d_stage + off_mask;
    float *o_val = is_device ? values : reinterpret_cast<float *>(r->d_stage + off_val);
    float *o_rw = is_device ? rewards : reinterpret_cast<float *>(r->d_stage + off_rw);
    if (obs) {
        GatherParams g{r->d_obs, d_idx, d_sh, o_obs, n, r->cfg.channels, r->cfg.y_dim, r->cfg.x_dim};
        const bool vec4 = (r->cfg.x_dim & 3) == 0;
        const size_t work = (size_t)n * r->obs_floats / (vec4 ? 4 : 1);
        if (vec4)
            gather_obs_kernel<true><<<blocks_for(work, 148 * 16), kThreads, 0, r->stream>>>(g);
        else
            gather_obs_kernel<false><<<blocks_for(work, 148 * 16), kThreads, 0, r->stream>>>(g);
        r->launches++;
    }
    if (policy) gather_rows_kernel<float><<<blocks_for((size_t)n * P), kThreads, 0, r->stream>>>(r->d_policy, d_idx, n, (int)P, o_pol), r->launches++;
    if (valid_mask) gather_rows_kernel<uint8_t><<<blocks_for((size_t)n * P), kThreads, 0, r->stream>>>(r->d_mask, d_idx, n, (int)P, o_mask), r->launches++;
    if (values) gather_rows_kernel<float><<<blocks_for((size_t)n), kThreads, 0, r->stream>>>(r->d_value, d_idx, n, 1, o_val), r->launches++;
    if (rewards) gather_rows_kernel<float><<<blocks_for((size_t)n), kThreads, 0, r->stream>>>(r->d_reward, d_idx, n, 1, o_rw), r->launches++;
    RCU(r, cudaGetLastError());
    if (!is_device) {
        if (obs) RCU(r, cudaMemcpyAsync(obs, o_obs, (size_t)n * r->obs_floats * 4, cudaMemcpyDeviceToHost, r->stream));
        if (policy) RCU(r, cudaMemcpyAsync(policy, o_pol, (size_t)n * P * 4, cudaMemcpyDeviceToHost, r->stream));
        if (valid_mask) RCU(r, cudaMemcpyAsync(valid_mask, o_mask, (size_t)n * P, cudaMemcpyDeviceToHost, r->stream));
        if (values) RCU(r, cudaMemcpyAsync(values, o_val, (size_t)n * 4, cudaMemcpyDeviceToHost, r->stream));
        if (rewards) RCU(r, cudaMemcpyAsync(rewards, o_rw, (size_t)n * 4, cudaMemcpyDeviceToHost, r->stream));
        RCU(r, cudaStreamSynchronize(r->stream));
    }
    return IPP_OK;
}

extern "C" int ipp_ring_update_priorities(ipp_ring *r, int32_t n, const int64_t *indices, const float *priorities, int32_t is_device) {
    if (!r) return IPP_ERR_INVALID;
    if (n < 1 || !indices || !priorities) return rfail(r, IPP_ERR_INVALID, "ipp_ring_update_priorities: n >= 1, indices and priorities required");
    const int64_t *d_idx = indices;
    const float *d_p = priorities;
    if (!is_device) {
        int rc = rensure(r, &r->d_stage, &r->cap_stage, (size_t)n * 12);
        if (rc != IPP_OK) return rc;
        RCU(r, cudaMemcpyAsync(r->d_stage, indices, (size_t)n * 8, cudaMemcpyHostToDevice, r->stream));
        RCU(r, cudaMemcpyAsync(r->d_stage + (size_t)n * 8, priorities, (size_t)n * 4, cudaMemcpyHostToDevice, r->stream));
        d_idx = reinterpret_cast<const int64_t *>(r->d_stage);
        d_p = reinterpret_cast<const float *>(r->d_stage + (size_t)n * 8);
    }
    scatter_priorities_kernel<<<blocks_for((size_t)n, 64), kThreads, 0, r->stream>>>(r->d_priority, d_idx, d_p, n, r->size, r->d_max);
    r->launches++;
    RCU(r, cudaGetLastError());
    if (!is_device) RCU(r, cudaStreamSynchronize(r->stream));
    return IPP_OK;
}

extern "C" int ipp_ring_get_priorities(ipp_ring *r, float *priorities) {
    if (!r || !priorities) return IPP_ERR_INVALID;
    if (r->size == 0) return IPP_OK;
    RCU(r, cudaMemcpyAsync(priorities, r->d_priority, (size_t)r->size * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    RCU(r, cudaStreamSynchronize(r->stream));
    return IPP_OK;
}

extern "C" void *ipp_ring_device_ptr(ipp_ring *r, int32_t which) {
    if (!r) return nullptr;
    switch (which) {
        case IPP_RING_PTR_OBS: return r->d_obs;
        case IPP_RING_PTR_POLICY: return r->d_policy;
        case IPP_RING_PTR_MASK: return r->d_mask;
        case IPP_RING_PTR_VALUE: return r->d_value;
        case IPP_RING_PTR_REWARD: return r->d_reward;
        case IPP_RING_PTR_PRIORITY: return r->d_priority;
        case IPP_RING_PTR_LAST_INDICES: return r->d_idx;
        case IPP_RING_PTR_LAST_WEIGHTS: return r->d_weights;
        case IPP_RING_PTR_STREAM: return (void *)r->stream;
        default: return nullptr;
    }
}
