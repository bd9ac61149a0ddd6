// dram_probe2.cu — memory-system ceiling for footprint-shaped traffic on B200, by access instruction.
//
// Every warp fetches randomly placed, 32-byte aligned segments of SEG bytes (16-byte chunks, lanes
// side by side, several segments per warp instruction when SEG < 512) from a working set far larger
// than L2, with 4 independent requests in flight per lane and no arithmetic besides a hash per segment.
// Reported: useful GB/s per (instruction, segment length, rows per group).  Run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum
// to see the DRAM bytes each instruction costs per useful byte (L2 fetch granularity).
//
//   mode 0  ld.global.nc.L1::no_allocate.v4   (LDG.128, registers)
//   mode 1  cp.async.cg 16 B                  (LDGSTS.BYPASS, L2 -> smem)
//   mode 2  cp.async.ca 16 B                  (LDGSTS through L1)
//   mode 3  cp.async.bulk (1-D TMA copy, one per segment, mbarrier completion)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dram_probe2.cu -o build/dram_probe2
// Usage: dram_probe2 [GB=24] [granularity=32] [mode=-1 (all)] [seg=-1 (all)] [rows=-1]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kWarps = 16;
constexpr int kSlotBytes = 12288;  // per-warp landing zone for the smem modes

// group g of a warp = `rows` segments of seg bytes, `pitch` bytes apart (a footprint plane), random 32 B aligned base
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) probe(const uint4 *buf, uint32_t n_blocks /* 64 KB blocks */, int seg, int rows, int pitch,
                                                     int groups, uint32_t seed, uint32_t *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * kWarps + w;
    unsigned char *slot = smem + w * kSlotBytes;
    __shared__ __align__(8) unsigned long long bars[kWarps];
    const uint32_t bar = smem_u32(&bars[w]);
    if (MODE == 3) {
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        __syncwarp();
    }
    uint32_t phase = 0;
    const int chunks = seg >> 4;                       // 16 B chunks per segment
    const int spi = chunks >= 32 ? 1 : 32 / chunks;    // segments per warp instruction
    const int my_seg = chunks >= 32 ? 0 : lane / chunks, my_chunk = chunks >= 32 ? lane : lane % chunks;
    const uint32_t span = (uint32_t)(rows - 1) * pitch + seg;
    const uint32_t slack_sectors = (65536u - span) / 32u;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int g = 0; g < groups; ++g) {
        const uint32_t h = mix(seed ^ mix(gw * 7919u + g));
        const size_t base = (size_t)__umulhi(h, n_blocks) * 65536u + (size_t)__umulhi(mix(h + 0x9E3779B9u), slack_sectors) * 32u;
        const unsigned char *gp = reinterpret_cast<const unsigned char *>(buf) + base;
        if (MODE == 3) {
            const uint32_t bytes = (uint32_t)rows * seg;
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            __syncwarp();
            for (int r = lane; r < rows; r += 32) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(slot + (r * seg) % (kSlotBytes - 4096))),
                             "l"(gp + (size_t)r * pitch), "r"((uint32_t)seg), "r"(bar)
                             : "memory");
            }
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done)
                             : "r"(bar), "r"(phase)
                             : "memory");
            }
            phase ^= 1;
            continue;
        }
        // rows are handled spi at a time (short segments) or one at a time in 512 B pieces (long ones)
        const int pieces = chunks >= 32 ? chunks / 32 : 1;
        for (int r0 = 0; r0 < rows; r0 += spi) {
            const int r = r0 + my_seg;
            const bool on = r < rows && (chunks >= 32 || lane < spi * chunks);
            for (int pc = 0; pc < pieces; ++pc) {
                const unsigned char *src = gp + (size_t)r * pitch + (size_t)(pc * 32 + my_chunk) * 16;
                if (!on) continue;
                if (MODE == 0) {
                    uint4 v;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                } else if (MODE == 1) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + ((r * seg + pc * 512) % (kSlotBytes - 512)) + my_chunk * 16)), "l"(src) : "memory");
                } else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(slot + ((r * seg + pc * 512) % (kSlotBytes - 512)) + my_chunk * 16)), "l"(src) : "memory");
                }
            }
        }
        if (MODE == 1 || MODE == 2) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 3;" ::: "memory");  // four groups in flight per warp
        }
    }
    if (MODE == 1 || MODE == 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) sink[0] = acc.x;
}

typedef void (*kern_t)(const uint4 *, uint32_t, int, int, int, int, uint32_t, uint32_t *);

int main(int argc, char **argv) {
    const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 24.0) * (1ull << 30);
    const int gran = argc > 2 ? atoi(argv[2]) : 32;
    const int only_mode = argc > 3 ? atoi(argv[3]) : -1, only_seg = argc > 4 ? atoi(argv[4]) : -1, only_rows = argc > 5 ? atoi(argv[5]) : -1;
    uint4 *buf;
    uint32_t *sink;
    const bool early = getenv("PROBE_LIMIT_EARLY") != nullptr;
    if (early) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    if (!early) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t lim = 0;
    cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity now %zu (set %s allocation)\n", lim, early ? "before" : "after");
    kern_t ks[4] = {probe<0>, probe<1>, probe<2>, probe<3>};
    const char *names[4] = {"ldg.nc.128", "cp.async.cg16", "cp.async.ca16", "cp.async.bulk"};
    const size_t smem = (size_t)kWarps * kSlotBytes;
    for (auto k : ks) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("working set %.1f GB, L2 fetch granularity limit %d B, %d warps/SM\n", bytes / 1e9, gran, kWarps);
    printf("%-14s %6s %5s %7s | %10s\n", "instr", "seg_B", "rows", "pitch_B", "useful GB/s");
    const int segs[] = {32, 64, 128, 192, 256, 512, 1024, 2048};
    const int rowss[] = {1, 23};
    for (int mode = 0; mode < 4; ++mode) {
        if (only_mode >= 0 && mode != only_mode) continue;
        for (int seg : segs) {
            if (only_seg >= 0 && seg != only_seg) continue;
            for (int rows : rowss) {
                if (only_rows >= 0 && rows != only_rows) continue;
                const int pitch = seg <= 1600 ? 1600 : 4096;
                if ((size_t)(rows - 1) * pitch + seg > 60000) continue;
                const size_t per_group = (size_t)rows * seg;
                const int groups = (int)((size_t)(512u << 20) / per_group / (148 * kWarps)) + 1;  // ~512 MB per launch
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    ks[mode]<<<148, kWarps * 32, smem>>>(buf, (uint32_t)(bytes >> 16), seg, rows, pitch, groups, 1234u + rep, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                const double moved = (double)148 * kWarps * groups * per_group;
                printf("%-14s %6d %5d %7d | %10.1f\n", names[mode], seg, rows, pitch, moved / (ms * 1e-3) / 1e9);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
