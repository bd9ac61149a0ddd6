// step_bulk.cuh — the throughput path of the fused step on IPP_LAYOUT_SUPER (sm_100a): a persistent kernel whose warps
// stage whole footprints in shared memory with bulk asynchronous copies (cp.async.bulk / UBLKCP, the TMA unit's 1-D form)
// completed through mbarriers.
//
// Layout (quad_math.cuh): one 192-byte super-tile per 4 x 4 cells = [16 x {mean,var} | 16 x gt].  The tiles a footprint
// touches in one tile row are one contiguous run of ntx * 192 bytes, so a 23 x 23 footprint is 6-7 bulk copies of ~1.2 KB
// (issued by ONE lane, no per-lane address arithmetic, no registers held) where the cp.async kernel (step_async.cuh) issues
// ~460 16-byte copies from all lanes, and DRAM serves ~1 KB bursts instead of isolated 128-byte lines.  Shared memory keeps
// the layout of the run, so one tile-local index serves the staged read and the HBM write-back of a quad.
//
// Per warp: a byte ring in shared memory holding up to kBulkDepth footprints in flight (small footprints -> deeper
// prefetch), one mbarrier per in-flight footprint, a ring of 48-byte plans (everything warp-uniform about an env-step is
// computed once, by one lane, when the warp takes a chunk of tickets: lane i plans ticket base + i).  No block-level
// synchronisation in the loop.  Work distribution: global ticket counter, guided self-scheduling (as step_async.cuh).
//
// INTER_AREA (rf = 2) on unscrambled odd footprints needs no tap tables: output o of n integrates inputs {2o-1, 2o, 2o+1}
// with weights {o, n, n-1-o} / (2n-1) — the quad's own 2 x 2 cells plus the row above and the column to the left; the
// weights are formed exactly as make_tap_entry() forms them, so the result is bit-identical to the table path of the
// general kernel.  Clipped non-square footprints (dsize quirk) take the table path.
//
// IPP_LAYOUT_SPLIT (template parameter SPLIT): the super-tile cut in two arrays, var[tile][16] and {mean[16] | gt[16]}[tile].
// The full step stages two runs per tile row (lanes 0.. issue the variance runs, lanes 16.. the {mean | gt} runs: the same
// bytes as one super-tile run); the covariance-only step (MODE_PREDICT without the adaptive mask) stages, reads and writes
// the variance runs ALONE — a third of the bytes — and runs with more warps per CTA (no ground truth, few registers).
//
//   grid = #SMs (persistent, 1 CTA / SM), block = up to 16 warps (24 for the covariance-only step), dynamic smem ~ 227 KB
//   smem = [warp] byte ring | [warp] plan ring | [warp] tap tables | [warp] mbarriers
#pragma once
#include "step_kernel.cuh"

namespace ipp {

#ifndef IPP_BULK_MAX_WARPS
#define IPP_BULK_MAX_WARPS 16
#endif
#ifndef IPP_BULK_DEPTH
#define IPP_BULK_DEPTH 4  // footprints in flight per warp (power of two)
#endif
#ifndef IPP_BULK_CHUNK
#define IPP_BULK_CHUNK 8
#endif
#ifndef IPP_BULK_GUIDE
#define IPP_BULK_GUIDE 3
#endif
#ifndef IPP_BULK_GUIDE_PREDICT
#define IPP_BULK_GUIDE_PREDICT 2  // covariance-only step (variance runs alone): little work per env, larger chunks pay
#endif
#ifndef IPP_BULK_TARGET_DEPTH
#define IPP_BULK_TARGET_DEPTH 3  // after a consumed footprint: stage one, and more while fewer than this many are in flight
#endif
#ifndef IPP_BULK_ENDGAME_DEPTH
#define IPP_BULK_ENDGAME_DEPTH 2  // in-flight footprints per warp once the chunks have shrunk to one ticket
#endif
#ifndef IPP_BULK_PREDICT_WARPS
#define IPP_BULK_PREDICT_WARPS 24  // covariance-only step: <= 85 registers per thread
#endif
constexpr int kBulkMaxWarps = IPP_BULK_MAX_WARPS;
constexpr int kBulkPredictWarps = IPP_BULK_PREDICT_WARPS;
constexpr int kBulkDepth = IPP_BULK_DEPTH;
constexpr int kBulkChunk = IPP_BULK_CHUNK;
constexpr int kBulkPlanRing = 16;  // live plans per warp: <= kBulkDepth in flight + <= 3 queued + a fresh chunk of <= 8
constexpr int kBulkTapFloats2 = 2 * kTapCap * 3;
static_assert((kBulkDepth & (kBulkDepth - 1)) == 0 && kBulkDepth >= 2 && kBulkDepth <= 8, "depth");
static_assert(kBulkDepth + 3 + kBulkChunk <= kBulkPlanRing, "plan ring too small");

struct BulkParams {
    StepParams base;
    unsigned int *tickets;  // [2] ping-pong work counters
    int parity;             // counter consumed by this launch; the other one is zeroed for the next
    int warps;              // warps per CTA
    int ring_bytes;         // per-warp staging ring (multiple of 16, >= the largest footprint)
    // action ids fetched by the kernel itself from the caller's mapped pinned host buffer (ipp_step, IPP_ZERO_COPY_IDS): the
    // ids are pulled over PCIe in 512-byte slices into base.action_ids (device memory) while the first footprints are
    // already being planned, instead of by a separate H2D copy the kernel has to wait for (16 us at 65 536 envs).
    const int32_t *host_ids;    // device alias of the host buffer (16-byte aligned), or nullptr: base.action_ids is complete
    unsigned int *slice_state;  // [ceil(n_jobs / 128)] epoch + 1: being copied, epoch + 2: in device memory
    unsigned int epoch;         // even, grows by 2 per launch of this kind
};
constexpr int kIdSlice = 128;  // ids per slice: one 16-byte load per lane

// Make slice s of the host ids available in device memory (warp-collective).  Whoever gets there first copies it, everyone else
// waits for that warp — which is running, since it set the state — so no CTA ever waits for one that has not started.
__device__ __forceinline__ void bulk_fetch_ids(const BulkParams &bp, int s, int lane) {
    const unsigned ready = bp.epoch + 2u, busy = bp.epoch + 1u;
    volatile unsigned *w = bp.slice_state + s;
    unsigned st = 0;
    if (lane == 0) {
        st = *w;
        if (st != ready && st != busy) st = atomicCAS(bp.slice_state + s, st, busy) == st ? 0xffffffffu : *w;
    }
    st = __shfl_sync(0xffffffffu, st, 0);
    if (st == 0xffffffffu) {  // this warp copies
        const int i0 = s * kIdSlice + lane * 4;
        int32_t *dst = const_cast<int32_t *>(bp.base.action_ids);
        if (i0 + 3 < bp.base.n_jobs) {
            int4 v;
            asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(bp.host_ids + i0));
            __stcg(reinterpret_cast<int4 *>(dst + i0), v);
        } else {
            for (int i = i0; i < bp.base.n_jobs && i < i0 + 4; ++i) {
                int v;
                asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(bp.host_ids + i));
                __stcg(dst + i, v);
            }
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) *w = ready;
    } else {
        unsigned spins = 0;
        while (st != ready) {
            if (lane == 0) {
                __nanosleep(100);
                st = *w;
                if (++spins > (1u << 24)) {  // ~2 s: the copying warp died (never in a healthy launch) — report instead of hanging
                    *(volatile int *)bp.base.status = 4;
                    st = ready;
                }
            }
            st = __shfl_sync(0xffffffffu, st, 0);
        }
    }
}

// Per-env plan (48 B in shared memory), written by the planning lane; ring_off by the lane that starts the copies.
struct __align__(16) BulkPlan {
    int job;           // < 0: out of work
    int geo;           // xl | yu << 16
    int dims;          // nx | ny << 8 | nqx << 16 | nqy << 24
    int tiles;         // ntx | ntr << 8 | lvl << 16 | flags << 24   (flags: 1 rf == 2, 2 analytic INTER_AREA taps, 4 unsupported,
                       //                                              8 analytic weights = the level's table entry)
    unsigned src_lo;   // byte offset of the footprint's first super-tile from the start of the belief array (64 bit); SPLIT: of its
                       // first 64-byte tile from the start of the var array (the {mean | gt} array: twice that)
    unsigned src_hi;
    int sizes;         // bytes per staged tile row (SPLIT: of the variance run) | ntr << 16
    int ring_off;      // byte offset of the staged footprint in the warp's ring
    int magic_x;       // floor(65536 / nqx) + 1
    int spare;
    float inv_cost1;   // 1 / (cost + 1)
    int outs;          // out_r | out_c << 8
};
static_assert(sizeof(BulkPlan) == 48, "BulkPlan layout");
constexpr int kBulkMaxLevels = IPP_MAX_ALTITUDE_LEVELS;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared (src / dst 16-byte aligned, bytes % 16 == 0), completion on an mbarrier
// IPP_BULK_LDHINT (experiment switch): L2 eviction priority of the staged lines: 0 none, 1 evict_first, 2 evict_last.
#ifndef IPP_BULK_LDHINT
#define IPP_BULK_LDHINT 0
#endif
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
#if IPP_BULK_LDHINT == 0
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
#else
    uint64_t pol;
#if IPP_BULK_LDHINT == 1
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
#endif
}
// 1-D bulk copy shared -> global (the TMA unit's store), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
// IPP_BULK_PREDICT_TMASTORE (experiment switch, default off): the covariance-only step (SPLIT, variance runs alone) updates the
// staged runs in shared memory and writes them back with one bulk store per tile row (whole 64-byte tiles: full-line writes from
// the TMA unit instead of 8-byte scattered stores from every lane).  Measured at C3: 57.5 us per launch against 56.6 us with
// the per-lane stores (the write-back is not bound by store transactions; +33 % bytes for the cells around the footprint).
// IPP_BULK_NOSTORE (experiment switch): the full step on super-tiles computes everything and skips the write-back (wrong results:
// timing probe only) — what do the stores cost?
#ifndef IPP_BULK_NOSTORE
#define IPP_BULK_NOSTORE 0
#endif
#ifndef IPP_BULK_PREDICT_TMASTORE
#define IPP_BULK_PREDICT_TMASTORE 0
#endif
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// Write-back of a quad row.  IPP_BULK_ST selects the cache operator (experiment switch): 0 default (write-back), 1 .cs (streaming /
// evict-first), 2 .wt (write-through), 3 .cg.  Measured at C3: see DESIGN.md 3.5.
#ifndef IPP_BULK_ST
#define IPP_BULK_ST 0
#endif
__device__ __forceinline__ void st_row4(unsigned char *p, float a, float b, float c, float d) {
#if IPP_BULK_ST == 1
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
#elif IPP_BULK_ST == 2
    __stwt(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
#elif IPP_BULK_ST == 3
    __stcg(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
#else
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
#endif
}
__device__ __forceinline__ void st_row2(unsigned char *p, float a, float b) {
#if IPP_BULK_ST == 1
    __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b));
#elif IPP_BULK_ST == 2
    __stwt(reinterpret_cast<float2 *>(p), make_float2(a, b));
#elif IPP_BULK_ST == 3
    __stcg(reinterpret_cast<float2 *>(p), make_float2(a, b));
#else
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
#endif
}

__device__ __forceinline__ void st_row1(unsigned char *p, float a) { *reinterpret_cast<float *>(p) = a; }

// staged ground truth of a footprint, IPP_LAYOUT_SPLIT ({mean | gt} runs in shared memory)
struct GtSplitShared {
    uint32_t base;  // shared address of the staged {mean | gt} runs
    int ntx, a, b;
    __device__ __forceinline__ float at(int r, int c) const {
        const int rr = b + r, cc = a + c;
        return lds32(base + (uint32_t)(((rr >> 2) * ntx + (cc >> 2)) * kSplitMgTileBytes + 64 + ((rr & 3) << 4) + ((cc & 3) << 2)));
    }
};

// staged ground truth of a footprint (super-tile runs in shared memory); (r, c) relative to the footprint's top-left cell
struct GtSuperShared {
    uint32_t base;  // shared address of the slot
    int ntx, a, b;  // tiles per staged tile row; xl & 3; yu & 3
    __device__ __forceinline__ float at(int r, int c) const {
        const int rr = b + r, cc = a + c;
        return lds32(base + (uint32_t)(((rr >> 2) * ntx + (cc >> 2)) * kSuperTileBytes + 128 + ((rr & 3) << 4) + ((cc & 3) << 2)));
    }
};

// Plan one env-step (one lane).  Also advances the env's stored previous action (when the step commits and KEEP_PREV is
// off) and raises the status word for footprints the INTER_AREA path cannot serve.
template <int MODE, bool SPLIT>
__device__ __forceinline__ void bulk_plan_env(const StepParams &p, bool quirk, bool write_prev, int job, BulkPlan *out) {
    if (job < 0) {
        out->job = -1;
        return;
    }
    const int id = __ldcg(p.action_ids + job);  // L2: with BulkParams::host_ids the ids are written by this very launch
    double *ps = p.prev_state + 3 * (size_t)job;
    const double *pv = (MODE == MODE_PREDICT && p.prev_in != nullptr) ? p.prev_in + 3 * (size_t)job : ps;
    const double q0 = pv[0], q1 = pv[1], q2 = pv[2];
    int lvl, col, row;
    decode_id(p, id, lvl, col, row);
    const AltLevel &L = p.lut[lvl];
    const int xl = max(col - L.rx, 0), xr = min(col + L.rx, p.X - 1);
    const int yu = max(row - L.ry, 0), yd = min(row + L.ry, p.Y - 1);
    const int nx = xr - xl + 1, ny = yd - yu + 1;
    const int nqx = (nx + 1) >> 1, nqy = (ny + 1) >> 1;
    const int out_r = quirk ? nqx : nqy, out_c = quirk ? nqy : nqx;
    const int ntx = (xr >> 2) - (xl >> 2) + 1, ntr = (yd >> 2) - (yu >> 2) + 1;
    int flags = 0;
    if (L.rf == 2) {
        flags |= 1;
        if (MODE != MODE_PREDICT) {
            if (out_r > ny || out_c > nx) flags |= 4;
            // D[pr, pc] = D[qy, qx] and both axes decimate 2n-1 -> n: the quad's own cells + the row above / column to the left
            if ((nx & 1) && (ny & 1) && out_r == nqy && out_c == nqx) flags |= 2;
            if (nx == min(2 * L.rx + 1, p.X) && ny == min(2 * L.ry + 1, p.Y)) flags |= 8;
        }
    }
    const int magic_x = (int)((uint32_t)(65536.0f * fast_rcp((float)nqx) * 1.00000012f) + 1u);
    const double px = __dadd_rn(__dmul_rn(p.res, (double)col), __dmul_rn(0.5, p.res));
    const double py = __dadd_rn(__dmul_rn(p.res, (double)row), __dmul_rn(0.5, p.res));
    const float cost = job_cost(p, px, py, L.alt, q0, q1, q2);
    if (write_prev) {
        ps[0] = px;
        ps[1] = py;
        ps[2] = L.alt;
    }
    if (flags & 4) *(volatile int *)p.status = 1;  // mapped host word, bit 0 is the only bit
    const int tile_bytes = SPLIT ? kSplitVarTileBytes : kSuperTileBytes;
    const unsigned long long src = (unsigned long long)job * (p.plane * (SPLIT ? sizeof(float) : sizeof(float2))) +
                                   (unsigned long long)((yu >> 2) * p.txm + (xl >> 2)) * tile_bytes;
    int4 *o = reinterpret_cast<int4 *>(out);
    o[0] = make_int4(job, xl | (yu << 16), nx | (ny << 8) | (nqx << 16) | (nqy << 24), ntx | (ntr << 8) | (lvl << 16) | (flags << 24));
    o[1] = make_int4((int)(unsigned)(src & 0xffffffffull), (int)(unsigned)(src >> 32), (ntx * tile_bytes) | (ntr << 16), 0);
    o[2] = make_int4(magic_x, 0, __float_as_int(fast_rcp(cost + 1.0f)), out_r | (out_c << 8));
}

// MODE: MODE_KALMAN (full step) or MODE_PREDICT (covariance only: no ground truth, no noise; the staged run still carries
// both).  ENTROPY / ADAPTIVE: reward variant and adaptive mask; EXTRAS: host-supplied noise and measurement read-back.
template <int MODE, bool ENTROPY, bool ADAPTIVE, bool EXTRAS, bool SPLIT>
__global__ void __launch_bounds__((MODE == MODE_PREDICT && SPLIT && !ADAPTIVE ? kBulkPredictWarps : kBulkMaxWarps) * 32, 1)
    ipp_step_bulk_kernel(const __grid_constant__ BulkParams bp) {
    // SPLIT: does this variant stage the {mean | gt} runs?  (the mask needs the mean, the measurement the ground truth)
    constexpr bool kNeedMg = !SPLIT || MODE != MODE_PREDICT || ADAPTIVE;
    constexpr bool kTmaStore = SPLIT && !kNeedMg && (IPP_BULK_PREDICT_TMASTORE != 0);  // write-back through shared memory + bulk stores
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const StepParams &p = bp.base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned char *after = smem_raw + (size_t)bp.warps * bp.ring_bytes;
    BulkPlan *plans = reinterpret_cast<BulkPlan *>(after) + w * kBulkPlanRing;
    after += (size_t)bp.warps * kBulkPlanRing * sizeof(BulkPlan);
    float2 *taps = reinterpret_cast<float2 *>(after) + (size_t)w * kBulkTapFloats2;
    after += (size_t)bp.warps * kBulkTapFloats2 * sizeof(float2);
    float4 *lvl_tab = reinterpret_cast<float4 *>(after);  // [level]{R, 1/R, s2, 0}, {1/nx, nqx/nx, 1/ny, nqy/ny} of the unclipped footprint
    after += (size_t)kBulkMaxLevels * 2 * sizeof(float4);
    const uint32_t bars = smem_addr(after) + (uint32_t)(w * kBulkDepth * 8);
    const uint32_t ring = smem_addr(smem_raw) + (uint32_t)(w * bp.ring_bytes);
    const int cap = bp.ring_bytes;

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kBulkDepth; ++k) mbar_init(bars + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((int)threadIdx.x < p.n_levels) {
        const AltLevel &L = p.lut[threadIdx.x];
        const int fw = min(2 * L.rx + 1, p.X), fh = min(2 * L.ry + 1, p.Y);
        const float inx = __frcp_rn((float)fw), iny = __frcp_rn((float)fh);  // == 1.0f / (float)n_in of make_tap_entry
        lvl_tab[2 * threadIdx.x] = make_float4(L.R, fast_rcp(L.R), L.s2, 0.0f);
        lvl_tab[2 * threadIdx.x + 1] = make_float4(inx, (float)((fw + 1) >> 1) * inx, iny, (float)((fh + 1) >> 1) * iny);
    }
    __syncthreads();

    unsigned int *ticket = bp.tickets + bp.parity;
    if (blockIdx.x == 0 && threadIdx.x == 0) bp.tickets[bp.parity ^ 1] = 0u;  // for the next launch

    // ids still on the host: the warps of the grid pull one slice each (one PCIe round trip for the whole batch)
    const bool fetch_ids = bp.host_ids != nullptr;
    if (fetch_ids) {
        const int n_slices = (p.n_jobs + kIdSlice - 1) / kIdSlice;
        for (int sl = blockIdx.x + gridDim.x * w; sl < n_slices; sl += gridDim.x * bp.warps) bulk_fetch_ids(bp, sl, lane);
    }

    const int n_jobs = p.n_jobs;
    const bool quirk = (p.flags & IPP_FLAG_NO_DSIZE_QUIRK) == 0;
    const bool commit = (p.flags & IPP_FLAG_NO_COMMIT) == 0;
    const bool write_prev = commit && (p.flags & IPP_FLAG_KEEP_PREV) == 0;
    const int grow = p.txm * (SPLIT ? kSplitVarTileBytes : kSuperTileBytes);  // HBM tile-row stride [bytes] (SPLIT: of the var array)
    unsigned char *plane0 = reinterpret_cast<unsigned char *>(SPLIT ? p.var : p.mean);
    unsigned char *plane_mg = reinterpret_cast<unsigned char *>(p.mean);  // SPLIT: the {mean | gt} array

    // ---- plan ring: [c_pos, f_pos) staged (in flight), [f_pos, q_tail) planned, not yet staged ------------------------
    int c_pos = 0, f_pos = 0, q_tail = 0;
    int n_if = 0, n_wait = 0;     // staged / planned-not-staged
    unsigned int fi = 0, ci = 0;  // staged / consumed so far (mbarrier slot and phase)
    int r_head = 0, r_tail = 0;   // byte ring: oldest staged footprint / next free byte
    int max_if = kBulkDepth;

    // Stage the next planned footprint if the ring has room for it (warp-uniform; one lane per tile row issues its copy).
    auto try_fill = [&]() -> bool {
        if (n_wait == 0 || n_if >= max_if) return false;
        BulkPlan *pl = plans + f_pos;
        if (pl->job >= 0) {
            const int4 b = *reinterpret_cast<const int4 *>(&pl->src_lo);
            const int row_bytes = b.z & 0xffff, ntr = b.z >> 16;
            const int bytes = ntr * row_bytes * ((SPLIT && kNeedMg) ? 3 : 1);
            int off = -1;
            if (n_if == 0)
                off = 0;
            else if (r_tail > r_head) {
                if (r_tail + bytes <= cap)
                    off = r_tail;
                else if (bytes <= r_head)
                    off = 0;
            } else if (r_tail < r_head) {
                if (r_tail + bytes <= r_head) off = r_tail;
            }
            if (off < 0) return false;
            if (kTmaStore) {  // bulk stores of consumed footprints may still be reading the ring space this one reuses
                bulk_wait_read0();
                __syncwarp();
            }
            const uint32_t bar = bars + 8u * (fi & (kBulkDepth - 1));
            if (lane == 0) {
                mbar_expect_tx(bar, (uint32_t)bytes);
                pl->ring_off = off;
            }
            const unsigned long long src0 = ((unsigned long long)(unsigned)b.y << 32) | (unsigned)b.x;
            if (lane < ntr) {
                const unsigned char *src = plane0 + src0 + (size_t)lane * grow;
                bulk_g2s(ring + (uint32_t)(off + lane * row_bytes), src, (uint32_t)row_bytes, bar);
            } else if (SPLIT && kNeedMg && lane >= 16 && lane - 16 < ntr) {  // the {mean | gt} runs: twice the offsets and sizes
                const unsigned char *src = plane_mg + 2 * src0 + (size_t)(lane - 16) * (2 * grow);
                bulk_g2s(ring + (uint32_t)(off + ntr * row_bytes + (lane - 16) * 2 * row_bytes), src, (uint32_t)(2 * row_bytes), bar);
            }
            if (n_if == 0) r_head = off;
            r_tail = off + bytes;
        }
        ++fi;
        f_pos = (f_pos + 1) & (kBulkPlanRing - 1);
        ++n_if;
        --n_wait;
        return true;
    };

    // The first pass of the loop finds nothing staged: it only takes the warp's first chunk of tickets, plans and stages it.
    unsigned int chunk_base = 0;
    bool exhausted = false;  // no ticket left behind this warp's last chunk
    constexpr int kGuide = (SPLIT && !kNeedMg) ? IPP_BULK_GUIDE_PREDICT : IPP_BULK_GUIDE;
    const float inv_guide = __frcp_rn((float)(max(kGuide, 1) * (int)gridDim.x * bp.warps));

#pragma unroll 1
    while (true) {
        // (A) ask for the next chunk of tickets early — the atomic's result is consumed only after this env has been fused
        const bool request = n_wait <= 2 && !exhausted;
        unsigned int fresh = 0;
        int req = kBulkChunk;
        if (request) {
            const int left = n_jobs - (int)min(chunk_base, (unsigned)n_jobs);  // as of this warp's previous chunk
            req = min(kBulkChunk, max(1, (int)((float)left * inv_guide)));
            if (req == 1) max_if = IPP_BULK_ENDGAME_DEPTH;  // end game: do not hoard work
            if (lane == 0) fresh = atomicAdd(ticket, (unsigned)req);
        }

        // (B) fuse the oldest staged env
        if (n_if == 0) {  // nothing staged, hence nothing planned either: done unless a chunk has just been asked for
            if (!request) break;
        } else {
            const BulkPlan *pl = plans + c_pos;
            const int4 pa = *reinterpret_cast<const int4 *>(pl);
            const int job = pa.x;
            if (job < 0) break;  // warp-uniform: tickets are monotonic, every later plan is empty too
            const int4 pb = *reinterpret_cast<const int4 *>(&pl->src_lo);
            const int4 pc = *reinterpret_cast<const int4 *>(&pl->magic_x);
            mbar_wait(bars + 8u * (ci & (kBulkDepth - 1)), (ci / kBulkDepth) & 1u);

            const int xl = pa.y & 0xffff, yu = pa.y >> 16;
            const int nx = pa.z & 255, ny = (pa.z >> 8) & 255, nqx = (pa.z >> 16) & 255, nqy = (pa.z >> 24) & 255;
            const int ntx = pa.w & 255, lvl = (pa.w >> 16) & 255, pflags = pa.w >> 24;
            const int rf = (pflags & 1) ? 2 : 1;
            const bool analytic = (pflags & 2) != 0, unsupported = (pflags & 4) != 0;
            const uint32_t slot = ring + (uint32_t)pb.w;
            const int srow = pb.z & 0xffff;  // staged tile-row stride [bytes]
            const int a4 = xl & 3, b4 = yu & 3;
            unsigned char *gbase = plane0 + (((unsigned long long)(unsigned)pb.y << 32) | (unsigned)pb.x);

            const float4 la = lvl_tab[2 * lvl];
            FuseCtx fc;
            fc.rf = rf;
            fc.R = la.x;
            fc.invR = la.y;
            const float s2 = la.z;

            // INTER_AREA: analytic weights, or this env's tap tables (clipped non-square footprints)
            TapView tapv;
            tapv.rows = taps;
            tapv.cols = taps + 3 * kTapCap;
            int tap_mode = TAPS_FAST;
            int out_c = 1;
            float inv_outc = 0.f;
            float inv_nx = 0.f, inv_ny = 0.f, wmid_x = 0.f, wmid_y = 0.f;
            if (MODE != MODE_PREDICT && rf == 2 && !unsupported) {
                if (analytic && (pflags & 8)) {  // the unclipped footprint of this level: weights from the table
                    const float4 lb = lvl_tab[2 * lvl + 1];
                    inv_nx = lb.x;
                    wmid_x = lb.y;
                    inv_ny = lb.z;
                    wmid_y = lb.w;
                } else if (analytic) {
                    inv_nx = __frcp_rn((float)nx);  // == 1.0f / (float)n_in of make_tap_entry
                    inv_ny = __frcp_rn((float)ny);
                    wmid_x = (float)nqx * inv_nx;
                    wmid_y = (float)nqy * inv_ny;
                } else {
                    const int out_r = pc.w & 255;
                    out_c = (pc.w >> 8) & 255;
                    inv_outc = __frcp_rn((float)out_c);
                    tap_mode = build_tap_tables<kTapCap>(taps, lane, ny, nx, out_r, out_c);
                }
            }
            const size_t nrow = (size_t)job * (size_t)p.noise_stride;
            float acc = 0.0f;
            float nrm_cache[4] = {0.f, 0.f, 0.f, 0.f};  // rf = 2: one Philox call serves four passes

            const uint32_t magic_x = (uint32_t)pc.x;
            const bool odd = (a4 & 1) != 0;  // warp-uniform: (c0, c0 + 1) do not share a 16-byte chunk
            const int nq = unsupported ? 0 : nqx * nqy;
            int it = 0;
            if constexpr (SPLIT) {
                // staged runs: [ntr x var run (srow bytes)] [ntr x {mean | gt} run (2 * srow bytes)]; the same geometry in HBM with
                // the arrays' tile-row strides (grow, 2 * grow)
                const int ntr_s = pb.z >> 16;
                const uint32_t slot_m = slot + (uint32_t)(ntr_s * srow);
                unsigned char *gbase_m = plane_mg + 2 * (((unsigned long long)(unsigned)pb.y << 32) | (unsigned)pb.x);
#pragma unroll 1
                for (int q = lane; q < nq; q += 32, ++it) {
                    const int qy = (int)(((uint32_t)q * magic_x) >> 16);
                    const int qx = q - qy * nqx;
                    const int c0 = 2 * qx, r0 = 2 * qy;
                    const bool cok = c0 + 1 < nx, rok = r0 + 1 < ny, ok3 = cok && rok;
                    const int cc = a4 + c0, ic = cc & 3, tcx = cc >> 2;
                    const int rr = b4 + r0, ir = rr & 3, trl = rr >> 2;
                    const bool last_c = ic == 3, last_r = ir == 3;
                    const int u = (ir << 4) + (ic << 2);                    // byte offset of cell (r0, c0) inside a 64-byte tile field
                    const int tv = trl * srow + tcx * kSplitVarTileBytes + u;  // ... inside the staged variance runs
                    const int dCv = last_c ? 52 : 4, dCm = last_c ? 116 : 4;   // to column c0 + 1 (next tile: 64 / 128 - 12)
                    const int dRv = last_r ? srow - 48 : 16, dRm = last_r ? 2 * srow - 48 : 16;  // to row r0 + 1, staged
                    const int dRgv = last_r ? grow - 48 : 16, dRgm = last_r ? 2 * grow - 48 : 16;  // ... in HBM
                    const uint32_t sv = slot + (uint32_t)tv;
                    const uint32_t sm = slot_m + (uint32_t)(2 * tv - u);    // mean of cell (r0, c0); its ground truth 64 bytes behind

                    float v[4], m[4] = {0.f, 0.f, 0.f, 0.f};
                    {
                        const uint32_t sv1 = sv + (rok ? dRv : 0);
                        if (odd) {
                            v[0] = lds32(sv);
                            v[1] = cok ? lds32(sv + dCv) : 0.0f;
                            v[2] = rok ? lds32(sv1) : 0.0f;
                            v[3] = ok3 ? lds32(sv1 + dCv) : 0.0f;
                        } else {
                            const float2 t = lds64(sv), b2 = lds64(sv1);
                            v[0] = t.x;
                            v[1] = cok ? t.y : 0.0f;
                            v[2] = rok ? b2.x : 0.0f;
                            v[3] = ok3 ? b2.y : 0.0f;
                        }
                    }
                    if (kNeedMg) {
                        const uint32_t sm1 = sm + (rok ? dRm : 0);
                        if (odd) {
                            m[0] = lds32(sm);
                            m[1] = cok ? lds32(sm + dCm) : 0.0f;
                            m[2] = rok ? lds32(sm1) : 0.0f;
                            m[3] = ok3 ? lds32(sm1 + dCm) : 0.0f;
                        } else {
                            const float2 t = lds64(sm), b2 = lds64(sm1);
                            m[0] = t.x;
                            m[1] = cok ? t.y : 0.0f;
                            m[2] = rok ? b2.x : 0.0f;
                            m[3] = ok3 ? b2.y : 0.0f;
                        }
                    }
                    const bool ok[4] = {true, cok, rok, ok3};

                    // ---- measurement ----------------------------------------------------------------------------
                    float z[4] = {0.f, 0.f, 0.f, 0.f};
                    if (MODE != MODE_PREDICT) {
                        float eps[4];
                        if (EXTRAS && p.noise != nullptr) {
                            if (rf == 1) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? __ldg(p.noise + nrow + (r0 + (k >> 1)) * nx + c0 + (k & 1)) : 0.0f;
                            } else {
                                eps[0] = __ldg(p.noise + nrow + q);
                            }
                        } else {
                            draw_normals(p, rf, q, lane, it, (uint32_t)job + p.env_id_offset, nrm_cache, eps);
                        }
                        const uint32_t gs = sm + 64u;  // gt of cell (r0, c0), staged
                        const int dCq = cok ? dCm : 0, dRq = rok ? dRm : 0;
                        if (rf == 1) {
                            float gv[4];
                            if (odd) {
                                gv[0] = lds32(gs);
                                gv[1] = lds32(gs + dCq);
                                gv[2] = lds32(gs + dRq);
                                gv[3] = lds32(gs + dRq + dCq);
                            } else {
                                const float2 g0 = lds64(gs), g1 = lds64(gs + dRq);
                                gv[0] = g0.x;
                                gv[1] = g0.y;
                                gv[2] = g1.x;
                                gv[3] = g1.y;
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) z[k] = ok[k] ? __saturatef(fmaf(s2, eps[k], gv[k])) : 0.0f;
                        } else if (analytic) {
                            // rows {r0-1, r0, r0+1} x cols {c0-1, c0, c0+1}; taps with weight 0 are clamped onto the quad's own cells
                            const int dU = qy > 0 ? ((ir == 0) ? -(2 * srow - 48) : -16) : 0;
                            const int dL = qx > 0 ? ((ic == 0) ? -116 : -4) : 0;
                            const float wl = (float)qx * inv_nx, wr_ = (float)(nqx - 1 - qx) * inv_nx;
                            const float wu = (float)qy * inv_ny, wb = (float)(nqy - 1 - qy) * inv_ny;
                            float rs[3];
                            const int dro[3] = {dU, 0, dRq};
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const uint32_t ra = gs + dro[k];
                                float g0, g1;
                                if (odd) {
                                    g0 = lds32(ra);
                                    g1 = lds32(ra + dCq);
                                } else {
                                    const float2 t = lds64(ra);
                                    g0 = t.x;
                                    g1 = t.y;
                                }
                                const float gl = lds32(ra + dL);
                                rs[k] = fmaf(wr_, g1, fmaf(wmid_x, g0, wl * gl));
                            }
                            float d = fmaf(wu, rs[0], 0.0f);
                            d = fmaf(wmid_y, rs[1], d);
                            d = fmaf(wb, rs[2], d);
                            z[0] = __saturatef(fmaf(s2, eps[0], d));
                        } else {
                            const int pr = fdiv(q, out_c, inv_outc), pcc = q - pr * out_c;
                            const float d = downsample(tap_mode, GtSplitShared{slot_m, ntx, a4, b4}, tapv, pr, pcc, ny, nx, pc.w & 255, out_c);
                            z[0] = __saturatef(fmaf(s2, eps[0], d));
                        }
                        if (EXTRAS && p.z_out != nullptr) {
                            if (rf == 1) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (ok[k]) p.z_out[nrow + (r0 + (k >> 1)) * nx + c0 + (k & 1)] = z[k];
                            } else {
                                p.z_out[nrow + q] = z[0];
                            }
                        }
                    }

                    // ---- fusion + reward, results straight to HBM --------------------------------------------------
                    float mn[4], vn[4];
                    bool msk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!ADAPTIVE || (fmaf(p.kappa, v[k], m[k]) >= p.thr));
                    acc += kalman_quad<ENTROPY, ADAPTIVE>(fc, cok, rok, m, v, z, msk, mn, vn);
                    if (kTmaStore) {
                        if (commit) {  // update the staged run in place; it goes back as a whole after the loop
                            const uint32_t sv1 = sv + dRv;
                            if (odd || !cok) {
                                sts32(sv, vn[0]);
                                if (cok) sts32(sv + dCv, vn[1]);
                                if (rok) sts32(sv1, vn[2]);
                                if (ok3) sts32(sv1 + dCv, vn[3]);
                            } else {
                                sts64(sv, vn[0], vn[1]);
                                if (rok) sts64(sv1, vn[2], vn[3]);
                            }
                        }
                    } else if (MODE != MODE_PREDICT || commit) {
                        unsigned char *gv_ = gbase + (trl * grow + tcx * kSplitVarTileBytes + u);
                        if (odd || !cok) {
                            st_row1(gv_, vn[0]);
                            if (cok) st_row1(gv_ + dCv, vn[1]);
                            if (rok) st_row1(gv_ + dRgv, vn[2]);
                            if (ok3) st_row1(gv_ + dRgv + dCv, vn[3]);
                        } else {
                            st_row2(gv_, vn[0], vn[1]);
                            if (rok) st_row2(gv_ + dRgv, vn[2], vn[3]);
                        }
                        if (MODE != MODE_PREDICT) {
                            unsigned char *gm_ = gbase_m + (trl * 2 * grow + tcx * kSplitMgTileBytes + u);
                            if (odd || !cok) {
                                st_row1(gm_, mn[0]);
                                if (cok) st_row1(gm_ + dCm, mn[1]);
                                if (rok) st_row1(gm_ + dRgm, mn[2]);
                                if (ok3) st_row1(gm_ + dRgm + dCm, mn[3]);
                            } else {
                                st_row2(gm_, mn[0], mn[1]);
                                if (rok) st_row2(gm_ + dRgm, mn[2], mn[3]);
                            }
                        }
                    }
                }
            } else {
    #pragma unroll 1
                for (int q = lane; q < nq; q += 32, ++it) {
                    const int qy = (int)(((uint32_t)q * magic_x) >> 16);
                    const int qx = q - qy * nqx;
                    // column part
                    const int c0 = 2 * qx;
                    const bool cok = c0 + 1 < nx;
                    const int cc = a4 + c0, ic = cc & 3;
                    const bool last_c = ic == 3;
                    const int col_s = (cc >> 2) * kSuperTileBytes + (ic << 3);        // {mean,var} of column c0 inside a tile row [bytes]
                    const int col_q = (cc >> 2) * kSuperTileBytes + 128 + (ic << 2);  // ground truth of column c0
                    const int dCg = last_c ? 168 : 8;                                 // to column c0 + 1 ({mean,var})
                    const int dC = cok ? dCg : 0;                                     // ... clamped inside the footprint
                    const int dCq = cok ? (last_c ? 180 : 4) : 0;                     // ... ground truth
                    // row part
                    const int r0 = 2 * qy;
                    const bool rok = r0 + 1 < ny;
                    const bool ok3 = cok && rok;
                    const int rr = b4 + r0, trl = rr >> 2, ir = rr & 3;
                    const bool last_r = ir == 3;
                    const int row_s = trl * srow;
                    const uint32_t so = slot + (uint32_t)(row_s + (ir << 5) + col_s);  // {mean,var} of cell (r0, c0), staged
                    unsigned char *go = gbase + (trl * grow + (ir << 5) + col_s);      // ... and in HBM
                    const int dRs = rok ? (last_r ? srow - 96 : 32) : 0;               // to the quad's second row (clamped inside the footprint)
                    const int dRg = last_r ? grow - 96 : 32;

                    float4 top, bot;
                    if (odd) {
                        const float2 t0 = lds64(so), t1 = lds64(so + dC), b0 = lds64(so + dRs), b1 = lds64(so + dRs + dC);
                        top = make_float4(t0.x, t0.y, t1.x, t1.y);
                        bot = make_float4(b0.x, b0.y, b1.x, b1.y);
                    } else {
                        top = lds128(so);
                        bot = lds128(so + dRs);
                    }
                    const float m[4] = {top.x, cok ? top.z : 0.0f, rok ? bot.x : 0.0f, ok3 ? bot.z : 0.0f};
                    const float v[4] = {top.y, cok ? top.w : 0.0f, rok ? bot.y : 0.0f, ok3 ? bot.w : 0.0f};
                    const bool ok[4] = {true, cok, rok, ok3};

                    // ---- measurement --------------------------------------------------------------------------------
                    float z[4] = {0.f, 0.f, 0.f, 0.f};
                    if (MODE != MODE_PREDICT) {
                        float eps[4];
                        if (EXTRAS && p.noise != nullptr) {
                            if (rf == 1) {
    #pragma unroll
                                for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? __ldg(p.noise + nrow + (r0 + (k >> 1)) * nx + c0 + (k & 1)) : 0.0f;
                            } else {
                                eps[0] = __ldg(p.noise + nrow + q);
                            }
                        } else {
                            draw_normals(p, rf, q, lane, it, (uint32_t)job + p.env_id_offset, nrm_cache, eps);
                        }
                        const uint32_t gs = slot + (uint32_t)(row_s + (ir << 4) + col_q);  // gt of cell (r0, c0), staged
                        const int dRq = rok ? (last_r ? srow - 48 : 16) : 0;
                        if (rf == 1) {
                            float gv[4];
                            if (odd) {
                                gv[0] = lds32(gs);
                                gv[1] = lds32(gs + dCq);
                                gv[2] = lds32(gs + dRq);
                                gv[3] = lds32(gs + dRq + dCq);
                            } else {
                                const float2 g0 = lds64(gs), g1 = lds64(gs + dRq);
                                gv[0] = g0.x;
                                gv[1] = g0.y;
                                gv[2] = g1.x;
                                gv[3] = g1.y;
                            }
    #pragma unroll
                            for (int k = 0; k < 4; ++k) z[k] = ok[k] ? __saturatef(fmaf(s2, eps[k], gv[k])) : 0.0f;
                        } else if (analytic) {
                            // rows {r0-1, r0, r0+1} x cols {c0-1, c0, c0+1}; taps with weight 0 are clamped onto the quad's own cells
                            const int dU = qy > 0 ? ((ir == 0) ? -(srow - 48) : -16) : 0;
                            const int dL = qx > 0 ? ((ic == 0) ? -180 : -4) : 0;  // to column c0 - 1 (ground truth), clamped
                            const float wl = (float)qx * inv_nx, wr_ = (float)(nqx - 1 - qx) * inv_nx;
                            const float wu = (float)qy * inv_ny, wb = (float)(nqy - 1 - qy) * inv_ny;
                            float rs[3];
                            const int dro[3] = {dU, 0, dRq};
    #pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const uint32_t ra = gs + dro[k];
                                float g0, g1;
                                if (odd) {
                                    g0 = lds32(ra);
                                    g1 = lds32(ra + dCq);
                                } else {
                                    const float2 t = lds64(ra);
                                    g0 = t.x;
                                    g1 = t.y;
                                }
                                const float gl = lds32(ra + dL);
                                rs[k] = fmaf(wr_, g1, fmaf(wmid_x, g0, wl * gl));
                            }
                            float d = fmaf(wu, rs[0], 0.0f);
                            d = fmaf(wmid_y, rs[1], d);
                            d = fmaf(wb, rs[2], d);
                            z[0] = __saturatef(fmaf(s2, eps[0], d));
                        } else {
                            const int pr = fdiv(q, out_c, inv_outc), pcc = q - pr * out_c;
                            const float d = downsample(tap_mode, GtSuperShared{slot, ntx, a4, b4}, tapv, pr, pcc, ny, nx, pc.w & 255, out_c);
                            z[0] = __saturatef(fmaf(s2, eps[0], d));
                        }
                        if (EXTRAS && p.z_out != nullptr) {
                            if (rf == 1) {
    #pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (ok[k]) p.z_out[nrow + (r0 + (k >> 1)) * nx + c0 + (k & 1)] = z[k];
                            } else {
                                p.z_out[nrow + q] = z[0];
                            }
                        }
                    }

                    // ---- fusion + reward, results straight to HBM ------------------------------------------------------
                    float mn[4], vn[4];
                    bool msk[4];
    #pragma unroll
                    for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!ADAPTIVE || (fmaf(p.kappa, v[k], m[k]) >= p.thr));
                    acc += kalman_quad<ENTROPY, ADAPTIVE>(fc, cok, rok, m, v, z, msk, mn, vn);
                    if (MODE == MODE_PREDICT) {  // covariance only: the mean goes back as it came
    #pragma unroll
                        for (int k = 0; k < 4; ++k) mn[k] = m[k];
                    }
                    if (IPP_BULK_NOSTORE != 0) acc += 1.0e-30f * ((mn[0] + mn[1]) + (mn[2] + mn[3]));  // keep the mean update alive
                    if ((IPP_BULK_NOSTORE == 0) && (MODE != MODE_PREDICT || commit)) {
                        if (odd) {
                            st_row2(go, mn[0], vn[0]);
                            if (cok) st_row2(go + dCg, mn[1], vn[1]);
                            if (rok) st_row2(go + dRg, mn[2], vn[2]);
                            if (ok3) st_row2(go + dRg + dCg, mn[3], vn[3]);
                        } else if (cok) {
                            st_row4(go, mn[0], vn[0], mn[1], vn[1]);
                            if (rok) st_row4(go + dRg, mn[2], vn[2], mn[3], vn[3]);
                        } else {
                            st_row2(go, mn[0], vn[0]);
                            if (rok) st_row2(go + dRg, mn[2], vn[2]);
                        }
                    }
                }

            }

            if (kTmaStore && commit && !unsupported) {  // the updated variance runs, one bulk store per tile row
                fence_proxy_async_smem();
                __syncwarp();
                const int ntr_s = pb.z >> 16;
                if (lane < ntr_s) {
                    bulk_s2g(gbase + (size_t)lane * grow, slot + (uint32_t)(lane * srow), (uint32_t)srow);
                    bulk_commit();
                }
            }
            // per-env information gain: fp32 partials per lane, fp32 tree across the warp; the cost term comes from the plan
            float accd = acc;
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, sft);
            if (lane == 0 && p.reward != nullptr) p.reward[job] = accd * __int_as_float(pc.z);
            __syncwarp();  // every lane is done with the staged footprint, the plan and the tap tables

            // (C) release the footprint
            ++ci;
            c_pos = (c_pos + 1) & (kBulkPlanRing - 1);
            --n_if;
            if (n_if > 0) r_head = plans[c_pos].ring_off;
        }

        // (D) plan the chunk requested at (A): lane i decodes ticket fresh + i
        if (request) {
            chunk_base = __shfl_sync(0xffffffffu, fresh, 0);
            exhausted = chunk_base + (unsigned)req >= (unsigned)n_jobs;
            if (chunk_base < (unsigned)n_jobs) {
                if (fetch_ids) {  // the slices this chunk's ids live in (ready long ago, except right after the start)
                    const int s0 = (int)chunk_base / kIdSlice, s1 = (int)(min(chunk_base + (unsigned)req, (unsigned)n_jobs) - 1u) / kIdSlice;
                    bulk_fetch_ids(bp, s0, lane);
                    if (s1 != s0) bulk_fetch_ids(bp, s1, lane);
                }
                if (lane < req) {
                    const unsigned int t = chunk_base + (unsigned)lane;
                    bulk_plan_env<MODE, SPLIT>(p, quirk, write_prev, t < (unsigned)n_jobs ? (int)t : -1, plans + ((q_tail + lane) & (kBulkPlanRing - 1)));
                }
                q_tail = (q_tail + req) & (kBulkPlanRing - 1);
                n_wait += req;
                __syncwarp();
            }
        }
        // (E) stage what fits: one footprint per consumed one, more while the pipeline is shallow
        if (try_fill()) {
            while (n_if < IPP_BULK_TARGET_DEPTH && try_fill()) {
            }
        }
        __syncwarp();  // ring_off of the freshly staged footprints is visible to every lane
    }
    if (kTmaStore) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory must outlive the bulk stores reading it
}

}  // namespace ipp
