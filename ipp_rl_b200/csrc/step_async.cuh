// step_async.cuh — the throughput path of the fused step (sm_100a): a persistent kernel whose warps
// stage whole footprints in shared memory with asynchronous copies (cp.async / LDGSTS) for the
// discrete action set (action ids; MV and TILED layouts).
//
// Why: the footprint gather is latency bound when a warp waits for its own loads quad by quad
// (ncu, v1: 21 % of HBM peak, 17 resident warps/SM, one third of a footprint in flight per warp).
// Here a warp keeps the COMPLETE footprint of its next env in flight — 16-byte asynchronous copies of the
// aligned superset of every footprint row ({mean,var} tile and ground-truth tile), all 32 lanes issuing,
// no registers held — while it fuses the env whose tiles have already landed; cp.async group
// accounting, no block-level synchronisation in the loop.  Shared memory is a pool of footprint slots:
// every warp owns one, the rest are second (prefetch) slots of the first `double_warps` warps.
//
// Work distribution: a global ticket counter with guided self-scheduling (chunks shrink with the work that
// is left).  When a warp takes a chunk, lane i PLANS ticket base + i: everything about an env-step that is
// warp-uniform (action decode, clipped footprint, staging geometry, division magics, tap-table choice, the
// cost term, the stored previous action) is computed once by one lane and kept as a 48-byte EnvPlan in
// shared memory, instead of redundantly by all 32 lanes per env.
//
// The loop body is kept small on purpose (ncu: the first version, 63 KB of SASS, spent most of its
// time in instruction-fetch stalls): rarely used paths are out of line, index divisions use a 16-bit
// magic multiplier, and the INTER_AREA tap tables of the unclipped footprints are precomputed per level.
//
// (A TMA variant of the same pipeline — 3-D tensor-map box copies — was measured and retired:
// the TMA unit spends ~40 cycles per 100-200 B box row, slower than even the plain LSU kernel on
// these narrow footprints; profiles/attic/step_tma.cuh.txt, tools/tma_probe.cu, DESIGN.md.)
//
//   grid  = #SMs (persistent, 1 CTA / SM), block = up to 16 warps, dynamic smem ~ 225 KB
//   smem  = [slot]{mv tile, gt tile} | [warp] plan ring | [warp] tap tables | level tap tables
#pragma once
#include "step_kernel.cuh"

namespace ipp {

#ifndef IPP_ASYNC_SLOTS
#define IPP_ASYNC_SLOTS 2
#endif
#ifndef IPP_ASYNC_MAX_WARPS
#define IPP_ASYNC_MAX_WARPS 16
#endif
// experiment switch: alias every env onto the first IPP_ASYNC_ALIAS + 1 envs (all traffic hits L2) to time the
// compute side of the pipeline alone; results are meaningless in that build
#ifdef IPP_ASYNC_ALIAS
#define IPP_ENV_OF(job) ((job) & IPP_ASYNC_ALIAS)
#else
#define IPP_ENV_OF(job) (job)
#endif
constexpr int kAsyncSlots = IPP_ASYNC_SLOTS;         // footprints in flight per warp
constexpr int kAsyncMaxWarps = IPP_ASYNC_MAX_WARPS;  // warps per CTA (1 CTA / SM), further limited by shared memory
#ifndef IPP_TICKET_CHUNK
#define IPP_TICKET_CHUNK 8
#endif
#ifndef IPP_TICKET_GUIDE
#define IPP_TICKET_GUIDE 3  // 0: fixed chunks of IPP_TICKET_CHUNK (measured: 2 -> 520, 3 -> 523, 4 -> 512, 6 -> 498 M env-steps/s)
#endif
// Experiment switch (TILED layout): 1 = only the ground truth is staged in shared memory; the belief — every cell is read
// exactly once, by the lane that rewrites it — is pulled into L2 one env ahead (IPP_PF_MODE) and read from there with
// ld.global.cg.  Measured and NOT adopted: slots shrink 7.4 -> 2.6 KB and 20+ warps fit with two slots each, but the L2
// latency of the per-quad loads is exposed: 445 M env-steps/s at 20 warps (405 M without the prefetch) vs 524 M staged.
#ifndef IPP_DIRECT_MV
#define IPP_DIRECT_MV 0
#endif
#ifndef IPP_PF_MODE
#define IPP_PF_MODE 1  // IPP_DIRECT_MV: 0 no prefetch, 1 bulk prefetch per tile row, 2 prefetch.global.L2 per tile, 3 per 32 B sector
#endif
#ifndef IPP_QUAD_UNROLL
#define IPP_QUAD_UNROLL 1
#endif
#define IPP_PRAGMA_(x) _Pragma(#x)
#define IPP_UNROLL(n) IPP_PRAGMA_(unroll n)
constexpr int kTicketChunk = IPP_TICKET_CHUNK;       // tickets taken per atomic
constexpr int kLevelTabs = 4;       // altitude levels whose interior tap tables are staged in smem
constexpr int kTapFloats2 = 2 * kTapCap * 3;  // float2 per tap-table pair (rows + cols)

struct AsyncParams {
    StepParams base;
    unsigned int *tickets;     // [2] ping-pong work counters
    const float2 *level_taps;  // [kLevelTabs][kTapFloats2] tap tables of the unclipped footprints
    int parity;                // counter consumed by this launch; the other one is zeroed for the next
    int warps;                 // warps per CTA
    int double_warps;          // warps [0, double_warps) own two staging slots (prefetch one env ahead), the rest one
    int mv_tile_bytes;         // per-slot tile capacities (multiples of 16 B)
    int gt_tile_bytes;
    int vec16;                 // 1: 16-byte L1-bypassing staging copies (x_dim % 4 == 0)
    int level_tap_mode[kLevelTabs];  // TAPS_FAST / TAPS_WIDE, or -1: no table for this level
};

// Per-env plan: everything about an env-step that is the same for all 32 lanes — decoded action, clipped footprint,
// staging geometry, division magics, tap-table choice, the cost term — computed ONCE, by one lane, when a warp takes
// a chunk of tickets (lane i plans ticket base + i), instead of redundantly by the whole warp for every env
// (ncu: ~250 of 1 400 warp-instructions per env-step were such warp-uniform arithmetic).  48 B in shared memory.
struct __align__(16) EnvPlan {
    int job;           // < 0: out of work
    int geo;           // xl | yu << 16
    int dims;          // nx | ny << 8 | nqx << 16 | nqy << 24
    int outs;          // out_r | out_c << 8 | cm << 16 | cg << 24   (cm / cg: 16-byte chunks per staged row)
    int magic_x;       // floor(65536 / nqx) + 1
    int magic_c;       // floor(65536 / out_c) + 1
    int misc;          // lvl | rf << 8 | (tap_mode + 1) << 16 (0: build the tables per env) | unsupported << 24
    float inv_cost1;   // 1 / (cost + 1)
    float R, invR, s2;
    int pad;
};
static_assert(sizeof(EnvPlan) == 48, "EnvPlan layout");
#ifndef IPP_PLAN_RING
#define IPP_PLAN_RING 12
#endif
constexpr int kPlanRing = IPP_PLAN_RING;  // live plans per warp: <= 2 in slots + <= 1 queued + a fresh chunk of <= 8 (+ 1 sentinel after the ring)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
#ifdef IPP_ASYNC_CA16
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {  // src 16 B aligned, bytes % 16 == 0
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// per-level tap tables of the unclipped (interior) footprints, built once at engine creation with the
// same device code the kernels use per env (bit-identical entries)
__global__ void build_level_taps_kernel(float2 *tabs, const int *dims /* [levels][4] = ny, nx, out_r, out_c */, int *modes) {
    const int k = blockIdx.x;
    const int ny = dims[4 * k], nx = dims[4 * k + 1], out_r = dims[4 * k + 2], out_c = dims[4 * k + 3];
    int mode = -1;
    if (out_r >= 1 && out_c >= 1 && out_r <= ny && out_c <= nx) {
        mode = build_tap_tables<kTapCap>(tabs + (size_t)k * kTapFloats2, (int)threadIdx.x, ny, nx, out_r, out_c);
        if (mode == TAPS_GENERIC) mode = -1;
    }
    if (threadIdx.x == 0) modes[k] = mode;
}

// Plan one env-step (one lane): see EnvPlan.  Also advances the env's stored previous action (unless KEEP_PREV) and
// raises the status word for footprints the INTER_AREA path cannot serve.
__device__ __forceinline__ void plan_env(const AsyncParams &ap, bool quirk, bool keep_prev, int job, EnvPlan *out) {
    const StepParams &p = ap.base;
    if (job < 0) {
        out->job = -1;
        return;
    }
    const int id = __ldg(p.action_ids + job);
    double *ps = p.prev_state + 3 * (size_t)job;
    const double q0 = ps[0], q1 = ps[1], q2 = ps[2];
    int lvl, col, row;
    decode_id(p, id, lvl, col, row);
    const AltLevel &L = p.lut[lvl];
    const int xl = max(col - L.rx, 0), xr = min(col + L.rx, p.X - 1);
    const int yu = max(row - L.ry, 0), yd = min(row + L.ry, p.Y - 1);
    const int nx = xr - xl + 1, ny = yd - yu + 1;
    const int cm = ((xl & 1) + nx + 1) >> 1, cg = ((xl & 3) + nx + 3) >> 2;
    const int nqx = (nx + 1) >> 1, nqy = (ny + 1) >> 1;
    const int out_r = quirk ? nqx : nqy, out_c = quirk ? nqy : nqx;
    const int rf = L.rf;
    int tap_mode = TAPS_FAST;
    bool unsupported = false;
    if (rf == 2) {
        unsupported = out_r > ny || out_c > nx;
        const int lmode = lvl < kLevelTabs ? ap.level_tap_mode[lvl] : -1;
        const bool interior = lmode >= 0 && nx == 2 * L.rx + 1 && ny == 2 * L.ry + 1 && (quirk || nqx == nqy);
        tap_mode = interior ? lmode : -1;  // -1: the warp builds this env's tables when it fuses it
    }
    // q / nqx and q / out_c with a 16-bit magic multiplier floor(65536/d)+1: exact for q*d < 65536, which
    // the host guarantees for every footprint that fits the shared-memory tiles (setup_async)
    const int magic_x = (int)((uint32_t)(65536.0f * fast_rcp((float)nqx) * 1.00000012f) + 1u);
    const int magic_c = (int)((uint32_t)(65536.0f * fast_rcp((float)out_c) * 1.00000012f) + 1u);
    const double px = __dadd_rn(__dmul_rn(p.res, (double)col), __dmul_rn(0.5, p.res));
    const double py = __dadd_rn(__dmul_rn(p.res, (double)row), __dmul_rn(0.5, p.res));
    const float cost = job_cost(p, px, py, L.alt, q0, q1, q2);
    if (!keep_prev) {
        ps[0] = px;
        ps[1] = py;
        ps[2] = L.alt;
    }
    if (unsupported) *(volatile int *)p.status = 1;  // mapped host word, bit 0 is the only bit
    int4 *o = reinterpret_cast<int4 *>(out);
    o[0] = make_int4(job, xl | (yu << 16), nx | (ny << 8) | (nqx << 16) | (nqy << 24), out_r | (out_c << 8) | (cm << 16) | (cg << 24));
    o[1] = make_int4(magic_x, magic_c, lvl | (rf << 8) | ((tap_mode + 1) << 16) | ((unsupported ? 1 : 0) << 24),
                     __float_as_int(fast_rcp(cost + 1.0f)));
    o[2] = make_int4(__float_as_int(L.R), __float_as_int(fast_rcp(L.R)), __float_as_int(L.s2), 0);
}

// ENTROPY / ADAPTIVE: reward variant and adaptive mask as compile-time switches; EXTRAS: host-supplied noise
// and measurement read-back (parity / test features, not on the throughput path).
template <bool ENTROPY, bool ADAPTIVE, bool EXTRAS, bool TILED>
__global__ void __launch_bounds__(kAsyncMaxWarps * 32, 1) ipp_step_async_kernel(const __grid_constant__ AsyncParams ap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const StepParams &p = ap.base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr bool kDirect = TILED && (IPP_DIRECT_MV != 0);  // belief read from L2, not staged (mv_tile_bytes == 0)
    const int stage_bytes = ap.mv_tile_bytes + ap.gt_tile_bytes;
    // slot pool: slot w is warp w's first slot, slot warps + w its second one (warps below double_warps only).  Shared
    // memory holds fewer than 2 x warps footprints of the largest size; more resident warps with some of them
    // un-prefetched hide more latency than fewer warps that all prefetch.
    const int n_slots_cta = ap.warps + ap.double_warps;
    const int nsl = w < ap.double_warps ? 2 : 1;  // staging slots of this warp
    unsigned char *after = smem_raw + (size_t)n_slots_cta * stage_bytes;
    EnvPlan *plans = reinterpret_cast<EnvPlan *>(after) + w * (kPlanRing + 1);  // ring + the "out of work" sentinel
    float2 *tap_base = reinterpret_cast<float2 *>(after + (size_t)ap.warps * (kPlanRing + 1) * sizeof(EnvPlan));
    float2 *taps = tap_base + (size_t)w * kTapFloats2;                 // this warp's per-env tables
    const float2 *lvl_taps = tap_base + (size_t)ap.warps * kTapFloats2;  // shared, read-only after the barrier
    auto slot_of = [&](int s) { return s == 0 ? w : ap.warps + w; };

    // stage the per-level tap tables (a few hundred bytes each) once per CTA
    for (int i = threadIdx.x; i < kLevelTabs * kTapFloats2; i += blockDim.x) const_cast<float2 *>(lvl_taps)[i] = __ldg(ap.level_taps + i);
    if (lane == 0) plans[kPlanRing].job = -1;
    __syncthreads();

    unsigned int *ticket = ap.tickets + ap.parity;
    if (blockIdx.x == 0 && threadIdx.x == 0) ap.tickets[ap.parity ^ 1] = 0u;  // for the next launch

    const int n_jobs = p.n_jobs;
    const bool quirk = (p.flags & IPP_FLAG_NO_DSIZE_QUIRK) == 0;
    const bool keep_prev = (p.flags & IPP_FLAG_KEEP_PREV) != 0;
    const int X = p.X;
    float2 *mv_base = reinterpret_cast<float2 *>(p.mean);

    // ---- plan queue: ring of planned, not yet staged env-steps ------------------------------------------
    int q_head = 0, q_tail = 0, q_cnt = 0;
    auto plan_chunk = [&](unsigned int base, int cnt) {  // lane i plans ticket base + i
        if (lane < cnt) {
            const unsigned int t = base + (unsigned)lane;
            int pos = q_tail + lane;
            pos -= pos >= kPlanRing ? kPlanRing : 0;
            plan_env(ap, quirk, keep_prev, t < (unsigned)n_jobs ? (int)t : -1, plans + pos);
        }
        q_tail += cnt;
        q_tail -= q_tail >= kPlanRing ? kPlanRing : 0;
        q_cnt += cnt;
        __syncwarp();
    };
    auto pop_plan = [&]() {
        const int pos = q_head;
        q_head = q_head + 1 == kPlanRing ? 0 : q_head + 1;
        --q_cnt;
        return pos;
    };

    // Start the asynchronous copies of a planned env's footprint into slot s (all lanes), one commit group.
    auto fill = [&](int s, const EnvPlan *pl) {
        const int4 a = *reinterpret_cast<const int4 *>(pl);
        const int job = a.x;
        if (job >= 0) {
            const int xl = a.y & 0xffff, yu = a.y >> 16;
            const int nx = a.z & 255, ny = (a.z >> 8) & 255;
            const uint32_t tile = smem_u32(smem_raw + (size_t)slot_of(s) * stage_bytes);
            if (TILED) {
                // IPP_LAYOUT_TILED: lanes = RP row-segments of W chunks, rows advance by RP; chunk (R, cc) of a plane with `tx`
                // tiles per tile-row sits at float4 index ((R>>2)*tx + tile(cc))*8 + (R&3)*2 + half(cc).  Stepping R by
                // RP = 2 alternates between "+4" (inside a tile) and "+tx*8 - 4" (into the tile below); RP = 4 always adds
                // tx*8: one add and one xor per row (step ^= step_a ^ step_b).  RP = 1 (footprints wider than 31 cells)
                // takes the generic stride.
                const int ox = xl & 1, oxg = xl & 3;
                const int cm = (a.w >> 16) & 255, cg = (a.w >> 24) & 255;  // 16 B chunks per row
                auto stage = [&](const unsigned char *plane, int tx, int cw, int cc_first, int cc_step, int tile_shift, uint32_t dst0) {
                    const int sh = cw <= 8 ? 3 : (cw <= 16 ? 4 : 5);  // W = 1 << sh lanes per row
                    const int RP = 32 >> sh;
                    const int lr = lane >> sh, lc = lane & ((1 << sh) - 1);
                    if (lc >= cw) return;
                    const int cc = cc_first + cc_step * lc;
                    int R = yu + lr;
                    uint32_t off = 16u * (uint32_t)(((R >> 2) * tx + (cc >> tile_shift)) * 8 + (R & 3) * 2 + ((cc >> (tile_shift - 1)) & 1));
                    uint32_t dst = dst0 + 16u * (uint32_t)(lr * cw + lc);
                    const uint32_t dstep = 16u * (uint32_t)(RP * cw);
                    if (RP >= 2) {
                        const uint32_t in_tile = 32u * (uint32_t)RP, cross = 16u * (uint32_t)(tx * 8) - 32u * (uint32_t)(4 - RP);
                        uint32_t step = ((R & 3) + RP >= 4) ? cross : in_tile;
                        const uint32_t flip = RP == 4 ? 0u : (in_tile ^ cross);
#pragma unroll 1
                        for (int r = lr; r < ny; r += RP) {
                            cp_async_16(dst, plane + off);
                            off += step;
                            step ^= flip;
                            dst += dstep;
                        }
                    } else {
#pragma unroll 1
                        for (int r = lr; r < ny; ++r) {
                            cp_async_16(dst, plane + off);
                            off += ((R & 3) == 3) ? 16u * (uint32_t)(tx * 8) - 96u : 32u;
                            ++R;
                            dst += dstep;
                        }
                    }
                };
                if (kDirect) {
                    // belief tiles -> L2: tile row k of the footprint is one contiguous run of ntx 128-byte tiles (warp-uniform)
                    const int ty0 = yu >> 2, ntr = ((yu + ny - 1) >> 2) - ty0 + 1;
                    const int tx0 = xl >> 2, ntx = ((xl + nx - 1) >> 2) - tx0 + 1;
                    const unsigned char *run = reinterpret_cast<const unsigned char *>(mv_base + (size_t)IPP_ENV_OF(job) * p.plane) +
                                               ((size_t)(ty0 * p.txm + tx0) << 7);
#if IPP_PF_MODE == 1
#pragma unroll 1
                    for (int k = 0; k < ntr; ++k) {
                        bulk_prefetch_l2(run, (uint32_t)ntx << 7);
                        run += (size_t)p.txm << 7;
                    }
#elif IPP_PF_MODE == 2 || IPP_PF_MODE == 3
                    // one prefetch per 128-byte tile (2) or per 32-byte sector (3), lanes side by side
                    constexpr int kPer = IPP_PF_MODE == 3 ? 4 : 1;
                    const uint32_t inv_ntx = 65536u / (uint32_t)ntx + 1u;
                    for (int u = lane; u < ntr * ntx * kPer; u += 32) {
                        const int t = u / kPer, sec = u - t * kPer;
                        const int tr = (int)(((uint32_t)t * inv_ntx) >> 16), tc = t - tr * ntx;
                        const unsigned char *a2 = run + (size_t)tr * ((size_t)p.txm << 7) + ((size_t)tc << 7) + sec * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a2) : "memory");
                    }
#endif
                } else {
                    stage(reinterpret_cast<const unsigned char *>(mv_base + (size_t)IPP_ENV_OF(job) * p.plane), p.txm, cm, xl - ox, 2, 2, tile);
                }
                stage(reinterpret_cast<const unsigned char *>(p.gt + (size_t)IPP_ENV_OF(job) * p.plane_gt), p.txg, cg, xl - oxg, 4, 3,
                      tile + (uint32_t)ap.mv_tile_bytes);
            } else if (ap.vec16) {
                // 16-byte copies that bypass L1 (cp.async.cg): every row is fetched as the 16-byte aligned
                // superset of its footprint segment — the same 32 B sectors, a quarter of the copy instructions,
                // and no L1 line allocation limiting the copies in flight.  Needs x_dim % 4 == 0.
                const int ox = xl & 1, oxg = xl & 3;
                const int cm = (a.w >> 16) & 255, cg = (a.w >> 24) & 255;  // 16 B chunks per row
                const size_t row0 = (size_t)IPP_ENV_OF(job) * p.plane + (size_t)(yu * X);
                {
                    const int W = cm <= 8 ? 8 : (cm <= 16 ? 16 : 32), RP = 32 / W;
                    const int lr = lane / W, lc = lane - lr * W;
                    for (int c0 = lc; c0 < cm; c0 += 32) {
                        const float4 *src = reinterpret_cast<const float4 *>(mv_base + row0 + (xl - ox)) + (size_t)(lr * (X >> 1) + c0);
                        uint32_t dst = tile + 16u * (uint32_t)(lr * cm + c0);
#pragma unroll 2
                        for (int r = lr; r < ny; r += RP) {
                            cp_async_16(dst, src);
                            src += RP * (X >> 1);
                            dst += 16u * (uint32_t)(RP * cm);
                        }
                    }
                }
                {
                    const int W = cg <= 8 ? 8 : (cg <= 16 ? 16 : 32), RP = 32 / W;
                    const int lr = lane / W, lc = lane - lr * W;
                    for (int c0 = lc; c0 < cg; c0 += 32) {
                        const float4 *src = reinterpret_cast<const float4 *>(p.gt + row0 + (xl - oxg)) + (size_t)(lr * (X >> 2) + c0);
                        uint32_t dst = tile + (uint32_t)ap.mv_tile_bytes + 16u * (uint32_t)(lr * cg + c0);
#pragma unroll 2
                        for (int r = lr; r < ny; r += RP) {
                            cp_async_16(dst, src);
                            src += RP * (X >> 2);
                            dst += 16u * (uint32_t)(RP * cg);
                        }
                    }
                }
            } else {
                // lanes are laid out as RP row-segments of W columns: a warp instruction covers RP rows of the
                // footprint, each a coalesced run of 32 B sectors; addresses advance by constant strides
                const int pitch = (nx + 1) & ~1;
                const int W = nx <= 8 ? 8 : (nx <= 16 ? 16 : 32);
                const int RP = 32 / W;
                const int lr = lane / W, lc = lane - lr * W;
                const size_t org = (size_t)IPP_ENV_OF(job) * p.plane + (size_t)(yu * X + xl);
                for (int c0 = lc; c0 < nx; c0 += 32) {  // one pass unless the footprint is wider than 32 cells
                    const float2 *src_mv = mv_base + org + (size_t)(lr * X + c0);
                    const float *src_gt = p.gt + org + (size_t)(lr * X + c0);
                    uint32_t dst_mv = tile + 8u * (uint32_t)(lr * pitch + c0);
                    uint32_t dst_gt = tile + (uint32_t)ap.mv_tile_bytes + 4u * (uint32_t)(lr * pitch + c0);
#pragma unroll 2
                    for (int r = lr; r < ny; r += RP) {
                        cp_async_8(dst_mv, src_mv);
                        cp_async_4(dst_gt, src_gt);
                        src_mv += RP * X;
                        src_gt += RP * X;
                        dst_mv += 8u * (uint32_t)(RP * pitch);
                        dst_gt += 4u * (uint32_t)(RP * pitch);
                    }
                }
            }
        }
        cp_async_commit();  // always: keeps the group count in step with the slot rotation
    };

    // ---- prologue: nsl + 1 tickets, planned at once; fill every slot -------------------------------------
    // Later chunks shrink with the work that is left ("guided" self-scheduling): chunk = clamp(remaining /
    // (IPP_TICKET_GUIDE * warps in the grid), 1, kTicketChunk).  An env takes a warp ~5 us, so fixed chunks of 8 left warps up
    // to 40 us of work after the counter ran dry while the rest of the GPU idled (a 145 us launch).
    unsigned int chunk_base = 0;
    if (lane == 0) chunk_base = atomicAdd(ticket, (unsigned)(nsl + 1));
    chunk_base = __shfl_sync(0xffffffffu, chunk_base, 0);
    bool exhausted = chunk_base + (unsigned)(nsl + 1) >= (unsigned)n_jobs;  // no ticket left behind this chunk
    plan_chunk(chunk_base, nsl + 1);
    const float inv_guide = 1.0f / (float)(max(IPP_TICKET_GUIDE, 1) * (int)gridDim.x * ap.warps);
    int sp0 = pop_plan(), sp1 = kPlanRing;  // plan held by slot 0 / slot 1
    fill(0, plans + sp0);
    if (nsl == 2) {
        sp1 = pop_plan();
        fill(1, plans + sp1);
    }
    int s = 0;

#pragma unroll 1
    while (true) {
        // (A) ask for the next chunk of tickets early — the atomic's result is consumed only after this env has been fused
        const bool request = q_cnt <= 1 && !exhausted;
        unsigned int fresh = 0;
        int req = kTicketChunk;
        if (request) {
            if (IPP_TICKET_GUIDE > 0) {
                const int left = n_jobs - (int)min(chunk_base, (unsigned)n_jobs);  // as of this warp's previous chunk
                req = min(kTicketChunk, max(1, (int)((float)left * inv_guide)));
            }
            if (lane == 0) fresh = atomicAdd(ticket, (unsigned)req);
        }

        if (nsl == 2)  // this lane's copies into slot s have landed (the younger group may still be in flight)
            cp_async_wait<1>();
        else
            cp_async_wait<0>();
        __syncwarp();  // ... and so have every other lane's

        // (B) fuse the env whose tiles sit in slot s
        const EnvPlan *pl = plans + (s == 0 ? sp0 : sp1);
        const int4 pa = *reinterpret_cast<const int4 *>(pl);
        const int job = pa.x;
        if (job < 0) break;  // warp-uniform: tickets are monotonic, every later slot is empty too
        const int4 pb = *reinterpret_cast<const int4 *>(&pl->magic_x);
        const float4 pc4 = *reinterpret_cast<const float4 *>(&pl->R);
        const int xl = pa.y & 0xffff, yu = pa.y >> 16;
        const int nx = pa.z & 255, ny = (pa.z >> 8) & 255, nqx = (pa.z >> 16) & 255, nqy = (pa.z >> 24) & 255;
        const int out_r = pa.w & 255, out_c = (pa.w >> 8) & 255;
        const int lvl = pb.z & 255, rf = (pb.z >> 8) & 255;
        const bool unsupported = (pb.z >> 24) != 0;
        const float s2 = pc4.z;
        // tile geometry: with 16-byte staging the tiles start at the aligned cell left of the footprint
        const bool aligned16 = TILED || ap.vec16;
        const int ox = aligned16 ? (xl & 1) : 0, oxg = aligned16 ? (xl & 3) : 0;
        const int pm = aligned16 ? ((ox + nx + 1) & ~1) : ((nx + 1) & ~1);  // {mean,var} tile pitch [cells]
        const int pg = aligned16 ? ((oxg + nx + 3) & ~3) : pm;              // ground-truth tile pitch [floats]
        const int nq = nqx * nqy;
        const unsigned char *st = smem_raw + (size_t)slot_of(s) * stage_bytes;
        const float2 *mv_t = reinterpret_cast<const float2 *>(st) + ox;
        const float *gt_t = reinterpret_cast<const float *>(st + ap.mv_tile_bytes) + oxg;

        // INTER_AREA tap tables: the per-level table when the footprint is unclipped, else built here
        TapView tapv;
        tapv.rows = taps;
        tapv.cols = taps + 3 * kTapCap;
        int tap_mode = ((pb.z >> 16) & 255) - 1;
        if (rf == 2) {
            if (tap_mode >= 0) {
                tapv.rows = lvl_taps + (size_t)lvl * kTapFloats2;
                tapv.cols = tapv.rows + 3 * kTapCap;
            } else if (!unsupported) {
                tap_mode = build_tap_tables<kTapCap>(taps, lane, ny, nx, out_r, out_c);
            }
        }

        FuseCtx fc;
        fc.rf = rf;
        fc.R = pc4.x;
        fc.invR = pc4.y;
        const uint32_t magic_x = (uint32_t)pb.x, magic_c = (uint32_t)pb.y;
        float2 *mv_g = mv_base + (size_t)IPP_ENV_OF(job) * p.plane + (TILED ? (size_t)0 : (size_t)(yu * X + xl));
        const size_t nrow = (size_t)job * (size_t)p.noise_stride;
        float acc = 0.0f;
        float nrm_cache[4] = {0.f, 0.f, 0.f, 0.f};  // rf = 2: one Philox call serves four passes

        if (!unsupported) {
            int it = 0;
            IPP_UNROLL(IPP_QUAD_UNROLL)
            for (int q = lane; q < nq; q += 32, ++it) {
                const int qy = (int)(((uint32_t)q * magic_x) >> 16);
                const int qx = q - qy * nqx;
                const int r0 = 2 * qy, c0 = 2 * qx;
                const bool cok = c0 + 1 < nx, rok = r0 + 1 < ny;
                const bool ok[4] = {true, cok, rok, cok && rok};
                const int r1 = min(r0 + 1, ny - 1);  // clamped: the tile has exactly ny rows

                // ---- belief: from L2 at the very addresses the results go back to (kDirect), else from the staged tile ----
                float4 top, bot;
                const int R0 = yu + r0, C0 = xl + c0;
                float2 *o = mv_g + (TILED ? tiled_mv_index(p.txm, R0, C0) : r0 * X + c0);
                const int dR = (R0 & 3) == 3 ? 16 * p.txm - 12 : 4;  // next row: inside the tile, or the tile below
                const int dC = (C0 & 3) == 3 ? 13 : 1;               // next column: inside the tile, or the tile to the right
                if (kDirect) {
                    if (ox == 0) {  // warp-uniform: (C0, C0+1) share a 16-byte chunk
                        top = __ldcg(reinterpret_cast<const float4 *>(o));
                        bot = rok ? __ldcg(reinterpret_cast<const float4 *>(o + dR)) : top;
                    } else {
                        const float2 zz = make_float2(0.f, 0.f);
                        const float2 a = __ldcg(o), b = cok ? __ldcg(o + dC) : zz, c2 = rok ? __ldcg(o + dR) : zz,
                                     d2 = ok[3] ? __ldcg(o + dR + dC) : zz;
                        top = make_float4(a.x, a.y, b.x, b.y);
                        bot = make_float4(c2.x, c2.y, d2.x, d2.y);
                    }
                } else if (ox == 0) {  // warp-uniform: 16-byte aligned quad rows
                    top = *reinterpret_cast<const float4 *>(mv_t + r0 * pm + c0);
                    bot = *reinterpret_cast<const float4 *>(mv_t + r1 * pm + c0);
                } else {
                    const float2 a = mv_t[r0 * pm + c0], b = mv_t[r0 * pm + c0 + 1], c2 = mv_t[r1 * pm + c0], d2 = mv_t[r1 * pm + c0 + 1];
                    top = make_float4(a.x, a.y, b.x, b.y);
                    bot = make_float4(c2.x, c2.y, d2.x, d2.y);
                }
                const float m[4] = {top.x, cok ? top.z : 0.0f, rok ? bot.x : 0.0f, ok[3] ? bot.z : 0.0f};
                const float v[4] = {top.y, cok ? top.w : 0.0f, rok ? bot.y : 0.0f, ok[3] ? bot.w : 0.0f};

                // ---- measurement -----------------------------------------------------------------
                float z[4] = {0.f, 0.f, 0.f, 0.f};
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
                if (rf == 1) {
                    const float gv[4] = {gt_t[r0 * pg + c0], gt_t[r0 * pg + c0 + 1], gt_t[r1 * pg + c0], gt_t[r1 * pg + c0 + 1]};
#pragma unroll
                    for (int k = 0; k < 4; ++k) z[k] = ok[k] ? __saturatef(fmaf(s2, eps[k], gv[k])) : 0.0f;
                } else {
                    int pr = qy, pc = qx;
                    if (out_c != nqx) {
                        pr = (int)(((uint32_t)q * magic_c) >> 16);
                        pc = q - pr * out_c;
                    }
                    const float d = downsample(tap_mode, GtShared{gt_t, pg}, tapv, pr, pc, ny, nx, out_r, out_c);
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

                // ---- fusion + reward, results straight to HBM --------------------------------------
                float mn[4], vn[4];
                bool msk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!ADAPTIVE || (fmaf(p.kappa, v[k], m[k]) >= p.thr));
                acc += kalman_quad<ENTROPY, ADAPTIVE>(fc, cok, rok, m, v, z, msk, mn, vn);
                if (TILED) {
                    if (ox == 0) {  // warp-uniform: (C0, C0+1) share a 16-byte chunk
                        if (cok) {
                            *reinterpret_cast<float4 *>(o) = make_float4(mn[0], vn[0], mn[1], vn[1]);
                            if (rok) *reinterpret_cast<float4 *>(o + dR) = make_float4(mn[2], vn[2], mn[3], vn[3]);
                        } else {
                            o[0] = make_float2(mn[0], vn[0]);
                            if (rok) o[dR] = make_float2(mn[2], vn[2]);
                        }
                    } else {
                        o[0] = make_float2(mn[0], vn[0]);
                        if (cok) o[dC] = make_float2(mn[1], vn[1]);
                        if (rok) o[dR] = make_float2(mn[2], vn[2]);
                        if (ok[3]) o[dR + dC] = make_float2(mn[3], vn[3]);
                    }
                } else {
                    o[0] = make_float2(mn[0], vn[0]);
                    if (cok) o[1] = make_float2(mn[1], vn[1]);
                    if (rok) o[X] = make_float2(mn[2], vn[2]);
                    if (ok[3]) o[X + 1] = make_float2(mn[3], vn[3]);
                }
            }
        }

        // per-env information gain: fp32 partials per lane (<= 17 quads), fp32 tree across the warp
        // (<= 32 partials of similar size, relative error ~3e-7); the cost term comes from the plan
        float accd = acc;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, sft);
        if (lane == 0 && p.reward != nullptr) p.reward[job] = accd * __int_as_float(pb.w);
        __syncwarp();  // every lane is done with slot s (tiles, plan, tap tables)

        // (C) refill slot s with the next planned env (the sentinel once the tickets have run out)
        const int pos_n = q_cnt > 0 ? pop_plan() : kPlanRing;
        if (s == 0)
            sp0 = pos_n;
        else
            sp1 = pos_n;
        fill(s, plans + pos_n);

        // (D) plan the chunk requested at (A): lane i decodes ticket fresh + i
        if (request) {
            chunk_base = __shfl_sync(0xffffffffu, fresh, 0);
            exhausted = chunk_base + (unsigned)req >= (unsigned)n_jobs;
            if (chunk_base < (unsigned)n_jobs) plan_chunk(chunk_base, req);
        }
        s = (s + 1 == nsl) ? 0 : s + 1;
    }
    cp_async_wait<0>();
}

}  // namespace ipp
