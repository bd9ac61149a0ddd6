// step_tma.cuh — the throughput path of the fused step (sm_100a): a persistent, TMA-staged,
// mbarrier-pipelined kernel for the discrete action set (action ids, MV layout).
//
// Why: the footprint gather is latency bound when every warp waits for its own LSU loads (ncu, v1:
// 21 % of HBM peak with 17 resident warps/SM).  Here one elected lane per warp asks the Tensor Memory
// Accelerator for the whole footprint of the env it will process NEXT — two 3-D box copies
// (interleaved {mean,var} tile + ground-truth tile) from tensor maps over [env][row][col] — while the
// warp fuses the env whose tiles have already landed in shared memory.  Completion is tracked with
// one mbarrier per (warp, slot) (expect_tx / try_wait.parity), so a warp never synchronises with any
// other warp.  Work is handed out through a global ticket counter (dynamic load balance: footprints
// are 81 / 289 / 529 cells).  Up to 2 x 16 footprints (~140 KB) are in flight per SM, independent of
// register pressure.  Results go back with plain 64-bit stores (write-back L2).
//
//   grid  = #SMs (persistent, 1 CTA / SM), block = up to 16 warps, dynamic smem ~ 220 KB
//   smem  = [warp][slot]{mv tile, gt tile} | mbarriers | per-warp INTER_AREA tap tables
#pragma once
#include <cuda.h>

#include "step_kernel.cuh"

namespace ipp {

constexpr int kTmaSlots = 2;     // tiles in flight per warp
constexpr int kTmaTapCap = 16;   // tap-table entries per axis (footprints up to 32 cells wide)
constexpr int kTmaMaxWarps = 16;

struct TmaParams {
    StepParams base;
    const CUtensorMap *maps;  // [n_levels][2] in global memory: {mean/var map, ground-truth map}
    unsigned int *tickets;    // [2] ping-pong work counters
    int parity;               // counter consumed by this launch; the other one is zeroed for the next
    int warps;                // warps per CTA
    int mv_tile_bytes;        // per-slot tile capacities (multiples of 128 B)
    int gt_tile_bytes;
    short bw_mv[IPP_MAX_ALTITUDE_LEVELS];  // box width  of the mean/var tile [cells]  (even, >= footprint + 1)
    short bw_gt[IPP_MAX_ALTITUDE_LEVELS];  // box width  of the ground-truth tile [floats] (multiple of 4, >= footprint + 3)
    short bh[IPP_MAX_ALTITUDE_LEVELS];     // box height [rows]
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_addr(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
        : "memory");
}

struct SlotJob {
    int job;  // -1: none
    int id;   // action id
};

__global__ void __launch_bounds__(kTmaMaxWarps * 32, 1) ipp_step_tma_kernel(const __grid_constant__ TmaParams tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const StepParams &p = tp.base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int stage_bytes = tp.mv_tile_bytes + tp.gt_tile_bytes;
    unsigned char *my_stages = smem_raw + (size_t)w * kTmaSlots * stage_bytes;
    unsigned char *after = smem_raw + (size_t)tp.warps * kTmaSlots * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after) + w * kTmaSlots;
    float4 *taps = reinterpret_cast<float4 *>(after + (size_t)tp.warps * kTmaSlots * sizeof(uint64_t)) + w * 2 * kTmaTapCap;

    unsigned int *ticket = tp.tickets + tp.parity;
    if (blockIdx.x == 0 && threadIdx.x == 0) tp.tickets[tp.parity ^ 1] = 0u;  // for the next launch

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaSlots; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int n_jobs = p.n_jobs;
    const bool quirk = (p.flags & IPP_FLAG_NO_DSIZE_QUIRK) == 0;
    const bool adaptive = (p.flags & IPP_FLAG_ADAPTIVE) != 0;
    const bool entropy = (p.flags & IPP_REWARD_MASK) == IPP_REWARD_GAUSS_ENTROPY;
    const bool keep_prev = (p.flags & IPP_FLAG_KEEP_PREV) != 0;
    const int X = p.X;
    float2 *mv_base = reinterpret_cast<float2 *>(p.mean);

    // Issue the two box copies of `job` into slot s (all lanes compute the geometry, lane 0 issues).
    // Returns the previous-action pose of the env (meaningful in lane 0) for the cost term.
    auto issue = [&](int s, int job, int id, double (&pq)[3]) {
        int lvl, col, row;
        decode_id(p, id, lvl, col, row);
        const AltLevel &L = p.lut[lvl];
        Geom g;
        clip_footprint(p, col, row, L.rx, L.ry, g);
        if (lane == 0) {
            unsigned char *st = my_stages + (size_t)s * stage_bytes;
            const uint32_t bytes = (uint32_t)tp.bh[lvl] * ((uint32_t)tp.bw_mv[lvl] * 8u + (uint32_t)tp.bw_gt[lvl] * 4u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bars[s], bytes);
            // the innermost box coordinate must be 16-byte aligned (probed: tools/tma_probe.cu), so the
            // boxes start at the enclosing even cell / multiple-of-4 column; same 32 B sectors either way
            tma_load_3d(st, tp.maps + 2 * lvl, 2 * (g.xl & ~1), g.yu, job, &bars[s]);
            tma_load_3d(st + tp.mv_tile_bytes, tp.maps + 2 * lvl + 1, g.xl & ~3, g.yu, job, &bars[s]);
            const double *pv = p.prev_state + 3 * (size_t)job;
            pq[0] = pv[0];
            pq[1] = pv[1];
            pq[2] = pv[2];
        }
    };

    // ---- prologue: three tickets in one atomic; fill both slots ---------------------------------
    unsigned int t0 = 0;
    if (lane == 0) t0 = atomicAdd(ticket, 3u);
    t0 = __shfl_sync(0xffffffffu, t0, 0);
    SlotJob slot[kTmaSlots];
    double pq[kTmaSlots][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int s = 0; s < kTmaSlots; ++s) {
        const unsigned int t = t0 + s;
        slot[s].job = t < (unsigned)n_jobs ? (int)t : -1;
        slot[s].id = slot[s].job >= 0 ? __ldg(p.action_ids + slot[s].job) : 0;
    }
#pragma unroll
    for (int s = 0; s < kTmaSlots; ++s)
        if (slot[s].job >= 0) issue(s, slot[s].job, slot[s].id, pq[s]);
    unsigned int tk = t0 + 2;  // ticket whose action id has not been loaded yet
    uint32_t phase = 0;        // bit s = parity to wait for on slot s

    bool done = false;
#pragma unroll 1
    while (!done) {
#pragma unroll
      for (int s = 0; s < kTmaSlots; ++s) {  // compile-time slot index: slot[], pq[] stay in registers
        if (slot[s].job < 0) {
            done = true;
            break;
        }

        // (A) start the next fetches early: action id of ticket tk, and a fresh ticket
        const int job_n = tk < (unsigned)n_jobs ? (int)tk : -1;
        const int id_n = job_n >= 0 ? __ldg(p.action_ids + job_n) : 0;
        unsigned int tk2 = 0;
        if (lane == 0) tk2 = atomicAdd(ticket, 1u);

        // (B) fuse the env whose tiles sit in slot s
        const int job = slot[s].job;
        int lvl, col, row;
        decode_id(p, slot[s].id, lvl, col, row);
        const Geom g = geom_from_cell(p, lvl, col, row);
        const int bw = tp.bw_mv[lvl], bwg = tp.bw_gt[lvl];
        const int nqx = (g.nx + 1) >> 1, nqy = (g.ny + 1) >> 1;
        const int nq = nqx * nqy;
        const int out_r = quirk ? nqx : nqy, out_c = quirk ? nqy : nqx;
        const unsigned char *st = my_stages + (size_t)s * stage_bytes;
        // tile origins are the 16-byte aligned cells left of the footprint: shift to the footprint
        const float2 *mv_t = reinterpret_cast<const float2 *>(st) + (g.xl & 1);
        const float *gt_t = reinterpret_cast<const float *>(st + tp.mv_tile_bytes) + (g.xl & 3);

        bool generic_taps = false;
        bool unsupported = false;
        if (g.rf == 2) {
            unsupported = out_r > g.ny || out_c > g.nx;
            if (!unsupported) generic_taps = build_tap_tables<kTmaTapCap>(taps, lane, g.ny, g.nx, out_r, out_c);
        }

        mbar_wait(&bars[s], (phase >> s) & 1u);
        phase ^= 1u << s;

        FuseCtx fc;
        fc.rf = g.rf;
        fc.R = g.R;
        fc.invR = __frcp_rn(g.R);
        fc.entropy = entropy;
        const float inv_nqx = __frcp_rn((float)nqx);
        const float inv_outc = __frcp_rn((float)out_c);
        float2 *mv_g = mv_base + (size_t)job * p.plane + (size_t)(g.yu * X + g.xl);
        const size_t nrow = (size_t)job * (size_t)p.noise_stride;
        float acc = 0.0f;  // per-lane partial (<= a few dozen quads); fp64 tree across the warp

        if (unsupported) {
            if (lane == 0) atomicOr(p.status, 1);
        } else {
            for (int q = lane; q < nq; q += 32) {
                const int qy = fdiv(q, nqx, inv_nqx), qx = q - qy * nqx;
                const int r0 = 2 * qy, c0 = 2 * qx;
                const bool cok = c0 + 1 < g.nx, rok = r0 + 1 < g.ny;
                const bool ok[4] = {true, cok, rok, cok && rok};

                // ---- belief from the staged tile (four 64-bit shared loads) ------------------------
                const float2 t00 = mv_t[r0 * bw + c0], t01 = mv_t[r0 * bw + c0 + 1];
                const float2 t10 = mv_t[(r0 + 1) * bw + c0], t11 = mv_t[(r0 + 1) * bw + c0 + 1];
                const float m[4] = {t00.x, cok ? t01.x : 0.0f, rok ? t10.x : 0.0f, ok[3] ? t11.x : 0.0f};
                const float v[4] = {t00.y, cok ? t01.y : 0.0f, rok ? t10.y : 0.0f, ok[3] ? t11.y : 0.0f};

                // ---- measurement ---------------------------------------------------------------------
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                float eps[4];
                if (p.noise != nullptr) {
                    if (g.rf == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) eps[k] = ok[k] ? __ldg(p.noise + nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)) : 0.0f;
                    } else {
                        eps[0] = __ldg(p.noise + nrow + q);
                    }
                } else {
                    uint32_t rnd[4];
                    philox4x32_10((uint32_t)q, (uint32_t)job + p.env_id_offset, p.step_lo, p.step_hi, p.seed_lo, p.seed_hi, rnd);
                    box_muller(rnd[0], rnd[1], eps[0], eps[1]);
                    if (g.rf == 1) box_muller(rnd[2], rnd[3], eps[2], eps[3]);
                }
                if (g.rf == 1) {
                    const float gv[4] = {gt_t[r0 * bwg + c0], gt_t[r0 * bwg + c0 + 1], gt_t[(r0 + 1) * bwg + c0], gt_t[(r0 + 1) * bwg + c0 + 1]};
#pragma unroll
                    for (int k = 0; k < 4; ++k) z[k] = ok[k] ? __saturatef(fmaf(g.s2, eps[k], gv[k])) : 0.0f;
                } else {
                    const int pr = fdiv(q, out_c, inv_outc), pc = q - pr * out_c;
                    float d = 0.0f;
                    if (!generic_taps) {
                        const float4 tr = taps[pr], tc = taps[kTmaTapCap + pc];
                        const int rs = __float_as_int(tr.x), cs = __float_as_int(tc.x);
                        const float wr[3] = {tr.y, tr.z, tr.w};
                        const int cb[3] = {cs, min(cs + 1, g.nx - 1), min(cs + 2, g.nx - 1)};
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float *rowp = gt_t + min(rs + a, g.ny - 1) * bwg;
                            const float rowsum = fmaf(tc.w, rowp[cb[2]], fmaf(tc.z, rowp[cb[1]], tc.y * rowp[cb[0]]));
                            d = fmaf(wr[a], rowsum, d);
                        }
                    } else {
                        const int rs = (pr * g.ny) / out_r, re = ((pr + 1) * g.ny + out_r - 1) / out_r;
                        const int cs = (pc * g.nx) / out_c, ce = ((pc + 1) * g.nx + out_c - 1) / out_c;
                        for (int a = rs; a < re; ++a) {
                            const float *rowp = gt_t + min(a, g.ny - 1) * bwg;
                            float rowsum = 0.0f;
                            for (int b = cs; b < ce; ++b) rowsum = fmaf(tap_weight_generic(pc, b, g.nx, out_c), rowp[min(b, g.nx - 1)], rowsum);
                            d = fmaf(tap_weight_generic(pr, a, g.ny, out_r), rowsum, d);
                        }
                    }
                    z[0] = __saturatef(fmaf(g.s2, eps[0], d));
                }
                if (p.z_out != nullptr) {
                    if (g.rf == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (ok[k]) p.z_out[nrow + (r0 + (k >> 1)) * g.nx + c0 + (k & 1)] = z[k];
                    } else {
                        p.z_out[nrow + q] = z[0];
                    }
                }

                // ---- fusion + reward, results straight to HBM ------------------------------------
                float mn[4], vn[4];
                bool msk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) msk[k] = ok[k] && (!adaptive || (fmaf(p.kappa, v[k], m[k]) >= p.thr));
                acc += kalman_quad(fc, cok, rok, m, v, z, msk, mn, vn);
                float2 *o = mv_g + r0 * X + c0;
                o[0] = make_float2(mn[0], vn[0]);
                if (cok) o[1] = make_float2(mn[1], vn[1]);
                if (rok) o[X] = make_float2(mn[2], vn[2]);
                if (ok[3]) o[X + 1] = make_float2(mn[3], vn[3]);
            }
        }

        double accd = (double)acc;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, sft);
        if (lane == 0) {
            const float cost = job_cost(p, g.px, g.py, g.ph, pq[s][0], pq[s][1], pq[s][2]);
            if (p.reward != nullptr) p.reward[job] = (float)accd * __frcp_rn(cost + 1.0f);
            if (!keep_prev) {
                double *ps = p.prev_state + 3 * (size_t)job;
                ps[0] = g.px;
                ps[1] = g.py;
                ps[2] = g.ph;
            }
        }
        __syncwarp();  // every lane is done with slot s (tiles + tap tables)

        // (C) refill slot s with the job fetched at (A)
        slot[s].job = job_n;
        slot[s].id = id_n;
        if (job_n >= 0) issue(s, job_n, id_n, pq[s]);
        // (D) the ticket requested at (A) becomes the next one to resolve
        tk = __shfl_sync(0xffffffffu, tk2, 0);
      }
    }
}

}  // namespace ipp
