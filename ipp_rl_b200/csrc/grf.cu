// grf.cu — on-device ground-truth generation for the reset path (sm_100a): Gaussian random fields for a whole
// env batch, resident in HBM from the first byte.
//
// Reference: simulations/ground_truths.py:14-33 (gaussian_random_field) as used by GaussianRandomField
// (simulations/simulations.py:37-48): white noise -> fft2 -> times sqrt(pk(|k|)), pk(k) = k^-cluster_radius, 0 at k = 0
// -> ifft2 -> real part -> min-max normalisation to [0, 1].  The reference fills the amplitude with a Python double
// loop (0.06 s per 200x200 map: about an hour for 65 536 envs); here per chunk of envs:
//
//   white_noise_kernel      Philox4x32-10 + Box-Muller, keyed by (seed, global env id, cell)   [or caller-supplied noise]
//   cufftExecR2C            batched 2-D real-to-complex transform (library FFT: cuFFT, loaded with dlopen at first use)
//   spectrum_scale_kernel   times the amplitude table (built on the host in fp64 with the reference's index quirks)
//   cufftExecC2R            batched inverse
//   normalise_kernel        per-env min / max (block reduction) and the write into the engine's ground-truth layout
//
// The amplitude table is Hermitian-symmetrised, (A(k) + A(-k)) / 2: for a real input that is exactly "real part of
// the inverse transform", also for odd dimensions where the reference's wave-number list leaves one spectrum
// row / column at zero (fft_indices, ground_truths.py:7-11).
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "engine_internal.h"

using namespace ipp;

namespace {

// ---- the five cuFFT entry points, resolved at run time ---------------------------------------------------------
typedef int cufftHandle_t;
typedef int (*fn_plan_many)(cufftHandle_t *, int, int *, int *, int, int, int *, int, int, int, int);
typedef int (*fn_set_stream)(cufftHandle_t, cudaStream_t);
typedef int (*fn_exec_r2c)(cufftHandle_t, float *, float2 *);
typedef int (*fn_exec_c2r)(cufftHandle_t, float2 *, float *);
typedef int (*fn_destroy)(cufftHandle_t);
constexpr int kCufftR2C = 0x2a, kCufftC2R = 0x2c;

struct CufftApi {
    void *lib = nullptr;
    fn_plan_many plan_many = nullptr;
    fn_set_stream set_stream = nullptr;
    fn_exec_r2c exec_r2c = nullptr;
    fn_exec_c2r exec_c2r = nullptr;
    fn_destroy destroy = nullptr;
    std::string err;
};

CufftApi &cufft() {
    static CufftApi api;
    if (api.lib || !api.err.empty()) return api;
    const char *names[] = {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so.12", "libcufft.so.10"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.err = "cuFFT not found (dlopen libcufft.so.11)";
        return api;
    }
    api.plan_many = (fn_plan_many)dlsym(api.lib, "cufftPlanMany");
    api.set_stream = (fn_set_stream)dlsym(api.lib, "cufftSetStream");
    api.exec_r2c = (fn_exec_r2c)dlsym(api.lib, "cufftExecR2C");
    api.exec_c2r = (fn_exec_c2r)dlsym(api.lib, "cufftExecC2R");
    api.destroy = (fn_destroy)dlsym(api.lib, "cufftDestroy");
    if (!api.plan_many || !api.set_stream || !api.exec_r2c || !api.exec_c2r || !api.destroy) {
        api.err = "cuFFT symbols missing";
        api.lib = nullptr;
    }
    return api;
}

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void white_noise_kernel(float *out, size_t plane, int n_env, uint32_t seed_lo, uint32_t seed_hi, uint32_t env0) {
    // four normals per Philox call: thread i fills cells 4i .. 4i+3 of env blockIdx.y
    const int env = blockIdx.y;
    float *o = out + (size_t)env * plane;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; 4 * g < plane; g += (size_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10((uint32_t)g, env0 + (uint32_t)env, 0x47524631u /* "GRF1" */, (uint32_t)(g >> 32), seed_lo, seed_hi, r);
        float n[4];
        box_muller(r[0], r[1], n[0], n[1]);
        box_muller(r[2], r[3], n[2], n[3]);
        for (int k = 0; k < 4; ++k)
            if (4 * g + k < plane) o[4 * g + k] = n[k];
    }
}

__global__ void spectrum_scale_kernel(float2 *spec, const float *amp, size_t half_plane, int n_env) {
    const size_t total = half_plane * (size_t)n_env;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float a = amp[i % half_plane];
        float2 s = spec[i];
        s.x *= a;
        s.y *= a;
        spec[i] = s;
    }
}

// one CTA per env: min / max of the field, then (f - min) / (max - min) into the engine's ground-truth layout
__global__ void __launch_bounds__(256) normalise_kernel(const float *field, float *gt, size_t plane, size_t plane_gt, int X, int txg, int ts_gt,
                                                        int gw_shift) {
    __shared__ float s_min[8], s_max[8];
    const int env = blockIdx.x;
    const float *f = field + (size_t)env * plane;
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = threadIdx.x; i < plane; i += blockDim.x) {
        const float v = f[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = lo;
        s_max[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    lo = s_min[0];
    hi = s_max[0];
    for (int w = 1; w < 8; ++w) {
        lo = fminf(lo, s_min[w]);
        hi = fmaxf(hi, s_max[w]);
    }
    const float span = hi - lo;  // (f - lo) / span: exactly 0 and 1 at the extremes, as the reference's division
    float *g = gt + (size_t)env * plane_gt;
    for (size_t i = threadIdx.x; i < plane; i += blockDim.x) {
        const int R = (int)(i / X), C = (int)(i - (size_t)R * X);
        g[txg > 0 ? tiled_gt_index_rt(txg, ts_gt, gw_shift, R, C) : i] = (f[i] - lo) / span;
    }
}

// reference wave numbers (ground_truths.py:7-11); entries past the list's length are "no amplitude"
std::vector<int> fft_indices(int n) {
    std::vector<int> v;
    for (int i = 0; i <= n / 2; ++i) v.push_back(i);
    for (int i = n / 2 - 1; i >= 1; --i) v.push_back(-i);
    return v;
}

}  // namespace

// Replaces GaussianRandomField.create_ground_truth_map (simulations/simulations.py:43-48) for envs
// [first_env, first_env + n_env).  Declared in include/ipp_b200.h.
extern "C" int ipp_generate_ground_truth(ipp_engine *e, double cluster_radius, uint64_t seed, const float *white_noise, int32_t first_env,
                                         int32_t n_env) {
    if (!e) return IPP_ERR_INVALID;
    StepParams p;
    ipp_internal_step_params(e, &p);
    if (first_env < 0 || n_env < 0 || first_env + n_env > p.batch) return ipp_internal_fail(e, IPP_ERR_INVALID, "ipp_generate_ground_truth: env range outside the batch");
    if (n_env == 0) return IPP_OK;
    CufftApi &fft = cufft();
    if (!fft.lib) return ipp_internal_fail(e, IPP_ERR_UNSUPPORTED, ("ipp_generate_ground_truth: " + fft.err).c_str());
    cudaStream_t stream = ipp_internal_stream(e);
    const int X = p.X, Y = p.Y, XH = X / 2 + 1;
    const size_t plane = (size_t)X * Y, half_plane = (size_t)Y * XH;

    // amplitude table A[i][j] = sqrt(pk(|k|)), pk(k) = k^-r, on the half spectrum, symmetrised (see file header)
    std::vector<double> full((size_t)Y * X, 0.0);
    const std::vector<int> ky = fft_indices(Y), kx = fft_indices(X);
    for (size_t i = 0; i < ky.size() && i < (size_t)Y; ++i)
        for (size_t j = 0; j < kx.size() && j < (size_t)X; ++j) {
            if (ky[i] == 0 && kx[j] == 0) continue;
            const double k = std::sqrt((double)ky[i] * ky[i] + (double)kx[j] * kx[j]);
            full[i * X + j] = std::sqrt(std::pow(k, -cluster_radius));
        }
    std::vector<float> amp(half_plane);
    for (int i = 0; i < Y; ++i)
        for (int j = 0; j < XH; ++j) {
            const int im = (Y - i) % Y, jm = (X - j) % X;
            amp[(size_t)i * XH + j] = (float)(0.5 * (full[(size_t)i * X + j] + full[(size_t)im * X + jm]));
        }

    // chunking bounds the scratch: real field + half spectrum per env in flight
    const size_t per_env = plane * sizeof(float) + half_plane * sizeof(float2);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_env, ((size_t)1 << 30) / per_env));
    float *d_field = nullptr, *d_amp = nullptr;
    float2 *d_spec = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_field);
        cudaFree(d_spec);
        cudaFree(d_amp);
    };
    if (cudaMalloc((void **)&d_field, (size_t)chunk * plane * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&d_spec, (size_t)chunk * half_plane * sizeof(float2)) != cudaSuccess ||
        cudaMalloc((void **)&d_amp, half_plane * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        cleanup();
        return ipp_internal_fail(e, IPP_ERR_NOMEM, "ipp_generate_ground_truth: scratch allocation failed");
    }
    cudaMemcpyAsync(d_amp, amp.data(), half_plane * sizeof(float), cudaMemcpyHostToDevice, stream);

    int rc = IPP_OK, launches = 0;
    cufftHandle_t fwd = 0, inv = 0;
    int planned_for = 0;
    for (int done = 0; done < n_env && rc == IPP_OK; done += chunk) {
        const int n = std::min(chunk, n_env - done);
        if (n != planned_for) {
            if (planned_for) {
                fft.destroy(fwd);
                fft.destroy(inv);
            }
            int dims[2] = {Y, X};
            if (fft.plan_many(&fwd, 2, dims, nullptr, 1, 0, nullptr, 1, 0, kCufftR2C, n) != 0 ||
                fft.plan_many(&inv, 2, dims, nullptr, 1, 0, nullptr, 1, 0, kCufftC2R, n) != 0) {
                rc = ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_generate_ground_truth: cufftPlanMany failed");
                break;
            }
            fft.set_stream(fwd, stream);
            fft.set_stream(inv, stream);
            planned_for = n;
        }
        if (white_noise) {
            cudaMemcpyAsync(d_field, white_noise + (size_t)done * plane, (size_t)n * plane * sizeof(float), cudaMemcpyHostToDevice, stream);
        } else {
            for (int y0 = 0; y0 < n; y0 += 65535) {  // grid.y limit
                const int ny = std::min(65535, n - y0);
                dim3 grid((unsigned)std::min<size_t>((plane / 4 + 255) / 256, 64), (unsigned)ny);
                white_noise_kernel<<<grid, 256, 0, stream>>>(d_field + (size_t)y0 * plane, plane, ny, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32),
                                                             p.env_id_offset + (uint32_t)(first_env + done + y0));
                ++launches;
            }
        }
        if (fft.exec_r2c(fwd, d_field, d_spec) != 0) {
            rc = ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_generate_ground_truth: cufftExecR2C failed");
            break;
        }
        spectrum_scale_kernel<<<(unsigned)std::min<size_t>(((size_t)n * half_plane + 255) / 256, 148 * 16), 256, 0, stream>>>(d_spec, d_amp, half_plane, n);
        if (fft.exec_c2r(inv, d_spec, d_field) != 0) {
            rc = ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_generate_ground_truth: cufftExecC2R failed");
            break;
        }
        normalise_kernel<<<n, 256, 0, stream>>>(d_field, const_cast<float *>(p.gt) + (size_t)(first_env + done) * p.plane_gt, plane, p.plane_gt, X, p.txg, p.ts_gt, p.gw_shift);
        launches += 4;
    }
    cudaError_t s = cudaStreamSynchronize(stream);
    if (planned_for) {
        fft.destroy(fwd);
        fft.destroy(inv);
    }
    cleanup();
    ipp_internal_count_launches(e, launches);
    if (rc == IPP_OK && (s != cudaSuccess || cudaGetLastError() != cudaSuccess)) rc = ipp_internal_fail(e, IPP_ERR_CUDA, "ipp_generate_ground_truth: CUDA failure");
    return rc;
}
