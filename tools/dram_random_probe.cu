// dram_random_probe.cu — how fast can a B200 read (and write back) randomly placed contiguous segments?
// The IPP step touches, per env, 2 x (9..23) row segments of 36-184 B scattered over a 31 GB working set;
// this probe measures the achievable DRAM throughput of exactly that kind of traffic as a function of the
// segment length, with plenty of loads in flight (so that it is the memory system, not the issue rate,
// that limits).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dram_random_probe.cu -o build/dram_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Each warp processes `groups` groups; a group = ROWS segments of SEG_BYTES bytes, `pitch` bytes apart, starting
// at a random (4-byte aligned) offset: the shape of one footprint plane.  mode 0: read; mode 1: read+write back.
template <int MODE>
__global__ void probe(float *buf, size_t n_floats, int seg_floats, int rows, size_t pitch_floats, int groups, uint32_t seed, float *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f;
    const size_t span = (size_t)rows * pitch_floats + seg_floats;
    for (int g = 0; g < groups; ++g) {
        const uint32_t h = mix(seed ^ mix((uint32_t)(warp * 7919u + g)));
        const uint32_t h2 = mix(h + 0x9E3779B9u);
        size_t base = ((((size_t)h << 32) | h2) % (n_floats - span));
        // lanes cover the segment(s): element e of the group = row e / seg, col e % seg
        const int total = rows * seg_floats;
        for (int e = lane; e < total; e += 32) {
            const int r = e / seg_floats, c = e - r * seg_floats;
            float *p = buf + base + (size_t)r * pitch_floats + c;
            const float v = __ldcg(p);
            if (MODE == 1) *p = v * 1.0000001f; else acc += v;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
    const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 24.0) * (1ull << 30);
    float *buf, *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, bytes);
    const size_t n = bytes / 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, argc > 2 ? atoi(argv[2]) : 32);
    printf("working set %.1f GB, L2 fetch granularity %s B\n", bytes / 1e9, argc > 2 ? argv[2] : "32");
    printf("%10s %6s %10s | %12s %12s\n", "seg_bytes", "rows", "pitch_B", "read GB/s", "r+w GB/s(x2)");
    const int segs[] = {32, 64, 96, 128, 184, 256, 368, 512, 1024, 4096};
    for (int si = 0; si < 10; ++si) {
        for (int rows_i = 0; rows_i < 2; ++rows_i) {
            const int seg = segs[si] / 4;
            const int rows = rows_i == 0 ? 1 : 23;
            const size_t pitch = 1600 / 4;  // {mean,var} row pitch of a 200-wide map
            if (rows > 1 && seg > (int)pitch) continue;
            const int warps_total = 148 * 64;  // full occupancy
            const int groups = (int)((size_t)(64u << 20) / ((size_t)rows * seg * 4) / warps_total) + 1;  // ~64 MB... per launch x
            double gbs[2];
            for (int mode = 0; mode < 2; ++mode) {
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) probe<0><<<148 * 8, 256>>>(buf, n, seg, rows, pitch, groups * 8, 1234u + rep, sink);
                    else probe<1><<<148 * 8, 256>>>(buf, n, seg, rows, pitch, groups * 8, 99u + rep, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double moved = (double)warps_total * groups * 8 * rows * seg * 4 * (mode == 0 ? 1 : 2);
                gbs[mode] = moved / (ms * 1e-3) / 1e9;
            }
            printf("%10d %6d %10zu | %12.1f %12.1f\n", seg * 4, rows, pitch * 4, gbs[0], gbs[1]);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
