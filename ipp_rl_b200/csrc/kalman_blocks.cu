// kalman_blocks.cu — static Kalman update on a diagonal covariance (sm_100a), stateless C-ABI entry point.
//
// Reference: Mapping.kalman_filter_update(P, H, R, grid_mean, observation, cov_only), mapping/mappings.py:155-215.
// For a diagonal P and a measurement model whose rows have disjoint supports and one weight per row (what
// AltitudeSensorModel.measurement_model_matrix builds, sensors/models/sensor_models.py:38-81) the innovation covariance
// S = H P H^T + R is diagonal and the dense update collapses to a closed form per measurement block i with cells C_i:
//     S_i = w_i^2 * sum_{k in C_i} v_k + R_i,   v'_j = v_j - (w_i v_j)^2 / S_i,   x'_j = x_j + (w_i v_j / S_i) (z_i - w_i sum_k x_k)
// The off-diagonals the dense update creates inside a block are dropped, as everywhere in this engine.
// fp64 like the reference (this is the B = 1 API surface, not the fp32 throughput path); one thread per measurement.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ipp_b200.h"

__global__ void kalman_blocks_kernel(int n_meas, const int32_t *row_ptr, const int32_t *cols, const double *weight, const double *noise_var,
                                     const double *obs, double *var, double *mean) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_meas) return;
    const int a = row_ptr[i], b = row_ptr[i + 1];
    const double w = weight[i];
    double sv = 0.0, sm = 0.0;
    for (int k = a; k < b; ++k) {
        sv += var[cols[k]];
        if (mean) sm += mean[cols[k]];
    }
    const double S = w * w * sv + noise_var[i];
    const double innov = (mean && obs) ? obs[i] - w * sm : 0.0;
    for (int k = a; k < b; ++k) {
        const int c = cols[k];
        const double v = var[c];
        const double g = w * v / S;
        var[c] = v - g * (w * v);
        if (mean && obs) mean[c] += g * innov;
    }
}

extern "C" int ipp_kalman_blocks(int32_t device, int32_t n_cells, int32_t n_meas, const int32_t *row_ptr, const int32_t *cols,
                                 const double *weight, const double *noise_var, const double *obs, double *var, double *mean) {
    if (n_cells < 1 || n_meas < 0 || !row_ptr || !weight || !noise_var || !var) return IPP_ERR_INVALID;
    if (n_meas == 0) return IPP_OK;
    if (!cols) return IPP_ERR_INVALID;
    const int nnz = row_ptr[n_meas];
    if (nnz < 0 || row_ptr[0] != 0) return IPP_ERR_INVALID;
    for (int i = 0; i < n_meas; ++i)
        if (row_ptr[i + 1] < row_ptr[i]) return IPP_ERR_INVALID;
    for (int k = 0; k < nnz; ++k)
        if (cols[k] < 0 || cols[k] >= n_cells) return IPP_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return IPP_ERR_CUDA;
    const size_t by_i = (size_t)(n_meas + 1) * sizeof(int32_t) + (size_t)nnz * sizeof(int32_t);
    const size_t by_d = (size_t)n_meas * 3 * sizeof(double) + (size_t)n_cells * 2 * sizeof(double);
    unsigned char *buf = nullptr;
    if (cudaMalloc((void **)&buf, by_d + by_i) != cudaSuccess) return IPP_ERR_NOMEM;
    double *d_w = reinterpret_cast<double *>(buf), *d_r = d_w + n_meas, *d_z = d_r + n_meas, *d_var = d_z + n_meas, *d_mean = d_var + n_cells;
    int32_t *d_ptr = reinterpret_cast<int32_t *>(d_mean + n_cells), *d_cols = d_ptr + n_meas + 1;
    cudaError_t s = cudaSuccess;
    auto up = [&](void *dst, const void *src, size_t n) {
        if (s == cudaSuccess && src) s = cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
    };
    up(d_w, weight, (size_t)n_meas * sizeof(double));
    up(d_r, noise_var, (size_t)n_meas * sizeof(double));
    up(d_z, obs, (size_t)n_meas * sizeof(double));
    up(d_var, var, (size_t)n_cells * sizeof(double));
    up(d_mean, mean, (size_t)n_cells * sizeof(double));
    up(d_ptr, row_ptr, (size_t)(n_meas + 1) * sizeof(int32_t));
    up(d_cols, cols, (size_t)nnz * sizeof(int32_t));
    if (s == cudaSuccess) {
        kalman_blocks_kernel<<<(n_meas + 127) / 128, 128>>>(n_meas, d_ptr, d_cols, d_w, d_r, obs ? d_z : nullptr, d_var, mean ? d_mean : nullptr);
        s = cudaGetLastError();
    }
    if (s == cudaSuccess) s = cudaMemcpy(var, d_var, (size_t)n_cells * sizeof(double), cudaMemcpyDeviceToHost);
    if (s == cudaSuccess && mean && obs) s = cudaMemcpy(mean, d_mean, (size_t)n_cells * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return s == cudaSuccess ? IPP_OK : IPP_ERR_CUDA;
}
