// tma_probe.cu — stand-alone probe: 3-D tiled TMA box loads at arbitrary (unaligned) element
// coordinates from a [B][Y][X] float tensor, descriptor passed (a) as __grid_constant__ kernel
// parameter and (b) from global memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a tools/tma_probe.cu -o /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int use_global, int c0, int c1, int c2, int box_elems,
                      float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *m = use_global ? gmap : &pmap;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(box_elems * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(saddr(tile)),
            "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(saddr(bar))
            : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(saddr(bar))
                     : "memory");
    }
    for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int a0 = argc > 1 ? atoi(argv[1]) : 6, a1 = argc > 2 ? atoi(argv[2]) : 3, abw = argc > 3 ? atoi(argv[3]) : 20;
    const int X = 48, Y = 24, B = 4;  // inner dim 48 floats (e.g. 24 {mean,var} cells)
    const int bw = abw, bh = 9;
    std::vector<float> h((size_t)X * Y * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 8192);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t gdim[3] = {X, Y, B};
    const cuuint64_t gstr[2] = {(cuuint64_t)X * 4, (cuuint64_t)X * Y * 4};
    const cuuint32_t box[3] = {bw, bh, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap *gmap;
    cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int use_global = 0; use_global < 2; ++use_global) {
        for (int trial = 0; trial < 3; ++trial) {
            const int c0 = trial == 0 ? a0 : (trial == 1 ? 38 : 2), c1 = trial == 0 ? a1 : (trial == 1 ? 20 : 0), c2 = 2;
            cudaMemset(out, 0, 8192);
            probe<<<1, 128, 16384>>>(map, gmap, use_global, c0, c1, c2, bw * bh, out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> o(bw * bh);
            cudaMemcpy(o.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int y = 0; y < bh; ++y)
                for (int x = 0; x < bw; ++x) {
                    const int gx = c0 + x, gy = c1 + y;
                    const float want = (gx < X && gy < Y) ? h[((size_t)c2 * Y + gy) * X + gx] : 0.0f;
                    if (o[y * bw + x] != want) ++bad;
                }
            printf("desc=%s coord=(%d,%d,%d): %s, mismatches=%d\n", use_global ? "global" : "param", c0, c1, c2, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
