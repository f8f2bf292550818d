// tf_stream_dev.cuh — TMA-staged form of the FP32 fitness sweep / fused generation for fixed-length
// chromosomes (no ragged lengths), in two warp schedules:
//   stream_kernel  one chromosome per warp step; W = 32 NW exactly (W = 32, 64); the variants WITH an obstacle table,
//                  uniform crossover (UX: both parent rows staged, two ring slots per child) and S > 1 samples (XF 2);
//   duo_kernel     two chromosomes per warp step (one per half-warp), dynamic tile tickets; the variants WITHOUT a
//                  table; runtime row length (3 <= W <= 256, masked lanes) on top of the exact-W form.
//
// Same tile structure and the SAME arithmetic (bit-identical results) as fitness_kernel /
// generation_kernel, but the rows never pass through the LSU on the way in:
//   * each warp owns a ring of kStages row buffers in shared memory with mbarriers;
//   * the lane that owns a child issues, from its own plan registers, the 1-D TMA bulk copies that gather the
//     child's two parent pieces ([0,cp) of one parent, [cp,W) of the other) straight into a ring
//     slot (cp.async.bulk ... mbarrier::complete_tx — SASS UBLKCP);
//   * the mutation is applied in shared memory and the neighbour waypoint is a second LDS instead of
//     three shuffles; stream_kernel also issues the voxel gathers of chromosome c+1 before c is evaluated.
// What is in flight is bounded by shared memory (kStages KB-sized rows per warp), not by
// registers, so the DRAM latency of the random parent gathers is hidden at 32 warps / SM.
#pragma once
#include "tf_fitness_dev.cuh"
#include "tf_ga_dev.cuh"
#include "tf_stream_host.cuh"

namespace tf {

// ---- fused ranking of the children (tf_generation_ranked): HR:114 sorts once per generation; here the generation
// kernel keeps the next_k (1 or 2) best ranking keys of the children it writes.  wbest: two keys per warp in shared
// memory; cta_done: the CTA's count of warps that ran dry; rank_part: one pair of keys per CTA in global memory.
// A warp notes the keys of a tile (one per owning lane, ~0 for lanes without a child): a ballot finds the rare
// child that beats the warp's next_k-th best so far.
__device__ __forceinline__ void rank_note(unsigned long long key, unsigned long long* wbest, int next_k, int lane) {
    unsigned int cand = __ballot_sync(0xffffffffu, key < wbest[next_k - 1]);
    while (cand) {
        const unsigned long long k = __shfl_sync(0xffffffffu, key, __ffs(cand) - 1);
        cand &= cand - 1u;
        if (lane == 0) {
            const unsigned long long a = wbest[0], b = wbest[1];
            if (k < a) { wbest[0] = k; wbest[1] = a; }
            else if (k < b) wbest[1] = k;
        }
    }
    __syncwarp();
}
// Called by every warp when it has run dry: the last one of the CTA merges the CTA's wpb x 2 keys into rank_part.
__device__ __forceinline__ void rank_merge_cta(const unsigned long long* wbest_all, unsigned int* cta_done,
                                               unsigned long long* __restrict__ rank_part, int lane, int wpb) {
    __threadfence_block();
    unsigned int nth = 0;
    if (lane == 0) nth = atomicAdd(cta_done, 1u);
    if (__shfl_sync(0xffffffffu, nth, 0) == (unsigned int)wpb - 1u) {
        __threadfence_block();
        unsigned long long a = ~0ull, b = ~0ull, first, second;
        if (lane < wpb) { a = wbest_all[2 * lane]; b = wbest_all[2 * lane + 1]; }
        warp_two_smallest(a, b, first, second);
        if (lane == 0) {
            rank_part[2 * blockIdx.x] = first;
            rank_part[2 * blockIdx.x + 1] = second;
            __threadfence();
        }
    }
    __syncwarp();
}
// Called by the last warp of the GRID to run dry (every CTA wrote its pair before its last warp counted itself).
__device__ __forceinline__ void rank_merge_grid(const unsigned long long* __restrict__ rank_part, int32_t* __restrict__ next_top,
                                                int next_k, int lane) {
    __threadfence();
    unsigned long long a = ~0ull, b = ~0ull, first, second;
    for (unsigned int i = lane; i < 2u * gridDim.x; i += 32u) {
        const unsigned long long k = __ldcg(rank_part + i);
        if (k < a) { b = a; a = k; }
        else if (k < b) b = k;
    }
    warp_two_smallest(a, b, first, second);
    if (lane == 0) {
        next_top[0] = (int32_t)(uint32_t)first;
        if (next_k > 1) next_top[1] = (int32_t)(uint32_t)second;
    }
}

// Score bookkeeping of one full-length chromosome whose waypoints sit in shared memory.
template <int NW, int SEM, bool VOX, int XF>
__device__ __forceinline__ void reduce_fullw(const EnvDev& e, const tf_params& prm, const Tables& t,
                                             const Lanes<NW>& cur, const Lanes<NW>& prev, const uint32_t (&vword)[NW],
                                             const uint32_t (&vmask)[NW], PointAcc<NW, SEM>& acc, int c,
                                             TileSums<NW>& out) {
    constexpr int W = 32 * NW;
    const int lane = threadIdx.x & 31;
    float a_prox = 0.0f, a_dist = 0.0f, a_alt = 0.0f, a_haz = 0.0f;
    int turns = 0, ncoll = 0;
    uint32_t hit = acc.hit;
    const bool has_dist = (SEM != TF_SEM_NONE) || (XF >= 1 && e.n_dyn > 0);
    float dn[NW];
#pragma unroll
    for (int s = 0; s < NW; ++s) dn[s] = (SEM == TF_SEM_HR) ? sqrt_rn(acc.best[s]) : acc.best[s];
    if constexpr (XF >= 1) {
        if (e.n_dyn > 0) {
            float m2[NW];
#pragma unroll
            for (int s = 0; s < NW; ++s) m2[s] = INFINITY;
            eval_dynamic<NW>(e, t, cur.x, cur.y, cur.z, lane, 32, hit, m2);
#pragma unroll
            for (int s = 0; s < NW; ++s) dn[s] = fminf(dn[s], sqrt_rn(m2[s]));
        }
    }
    const int S = (XF >= 2) ? max(prm.samples_per_segment, 1) : 1;
#pragma unroll
    for (int s = 0; s < NW; ++s) {
        const int k = lane + 32 * s;
        bool h = (hit >> s) & 1u;
        if constexpr (VOX) h = h || (vword[s] & vmask[s]) != 0u;
        const uint32_t word = __ballot_sync(0xffffffffu, h);
        if (lane == c) out.cw[s] = word;
        ncoll += h ? 1 : 0;
        if (has_dist) a_prox = __fadd_rn(a_prox, rcp_rn(__fadd_rn(1.0f, dn[s])));  // HR:46
        if constexpr (XF >= 1) {
            if (e.hazard != nullptr) a_haz = __fadd_rn(a_haz, hazard_sample(e, cur.x[s], cur.y[s], (float)k));
        }
        if (s > 0 || lane > 0) {  // k >= 1
            const float px = prev.x[s], py = prev.y[s], pz = prev.z[s];
            const float dx = cur.x[s] - px, dy = cur.y[s] - py, dz = cur.z[s] - pz;
            a_dist = __fadd_rn(a_dist, sqrt_rn(dist2(dx, dy, dz)));  // HR:51-52
            a_alt = __fadd_rn(a_alt, fabsf(dz));                         // HR:44
            if (k <= W - 2) turns += (cur.x[s] != px || cur.y[s] != py) ? 1 : 0;  // HR:38-40,43
            if constexpr (XF >= 2) {
                for (int j = 1; j < S; ++j) {  // extension: samples inside segment (k-1, k)
                    float tj = __fdiv_rn((float)j, (float)S);
                    bool sh;
                    float sdn;
                    eval_point_full<SEM, VOX, XF>(e, t, fmaf(dx, tj, px), fmaf(dy, tj, py), fmaf(dz, tj, pz),
                                                     __fadd_rn((float)(k - 1), tj), sh, sdn);
                    ncoll += sh ? 1 : 0;
                    if (has_dist) a_prox = __fadd_rn(a_prox, rcp_rn(__fadd_rn(1.0f, sdn)));
                }
            }
        }
    }
    const float dist = warp_sum_xor(a_dist), alt = warp_sum_xor(a_alt);
    const int packed = __reduce_add_sync(0xffffffffu, turns | (ncoll << 10));
    if (lane == c) { out.dist = dist; out.alt = alt; out.packed = packed; }
    if (has_dist) {
        const float prox = warp_sum_xor(a_prox);
        if (lane == c) out.prox = prox;
    }
    if constexpr (XF >= 1) {
        if (e.hazard != nullptr) {
            const float haz = warp_sum_xor(a_haz);
            if (lane == c) out.haz = haz;
        }
    }
}

// GA = true : fused generation (plan, gather both parents, mutate, store child, evaluate)
// GA = false: fitness sweep (row p, no store)
// UX = true : uniform crossover (north-star extension).  A child takes every waypoint from one of its two parents by a
//             Philox mask, so BOTH parent rows are staged — a child occupies two ring slots (the ring holds 2 children
//             instead of 4) — and the lanes move the masked waypoints of the second row into the first before the
//             mutation; from there on the slot is read exactly like a single-point child.
template <int NW, int SEM, bool VOX, int XF, bool GA, bool UX = false>
__global__ void __launch_bounds__(kBlockThreads, (XF == 0 && SEM == TF_SEM_NONE) ? 4 : (XF == 1 ? 3 : 0))
// 32 warps / SM for the HBM-bound plain variant (<= 64 registers), 24 for the gather-heavy XF 1
// variants (<= 80); the obstacle-loop (FP32-bound) and sampling variants are left to ptxas
// (measured: forcing 3 or 4 CTAs / SM on the PF variants costs 4-12 %)
stream_kernel(const __grid_constant__ EnvDev e, const __grid_constant__ tf_params prm,
              const __grid_constant__ tf_ga_params ga, const float4* __restrict__ pop,
              const float* __restrict__ fit, const int32_t* __restrict__ top_idx,
              const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen, int P, int T,
              float4* __restrict__ next, float* __restrict__ out_fit, uint32_t* __restrict__ out_coll,
              int32_t* __restrict__ sel, int32_t* __restrict__ cpo) {
    constexpr int W = 32 * NW;
    constexpr int kSl = UX ? 2 : 1, kSt = kStages / kSl;   // ring slots per child, children in flight
    static_assert(!UX || GA, "uniform crossover exists only in the fused generation");
    extern __shared__ __align__(128) unsigned char smem_stream[];
    unsigned char* smem = smem_stream;
    const SmemLayout sm = carve_smem(e, smem);
    const Tables tabs{sm.A, sm.B, sm.dA, sm.dB};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t tables = (smem_bytes_for(e) + 127) & ~(size_t)127;
    float4* ring = reinterpret_cast<float4*>(smem + tables) + (size_t)warp * kStages * W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tables + (size_t)kWarpsPerBlock * kStages * W * 16) + warp * kStages;
    const int ne = GA ? min(ga.n_elite, P) : 0;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
    }
    if ((SEM != TF_SEM_NONE && e.n_smem > 0) || (XF >= 1 && e.n_dyn_smem > 0)) {
        if (threadIdx.x == 0) mbar_init(sm.bar, 1);
        fence_mbar_init();
        __syncthreads();
        if (threadIdx.x == 0)
            stage_chunk(e, sm.A, sm.B, 0, min(e.n_smem, kTableCap), sm.bar, XF >= 1 && e.n_dyn_smem > 0, sm.dA, sm.dB);
        mbar_wait(sm.bar, 0);
    } else {
        fence_mbar_init();
        __syncwarp();
    }

    // ring bookkeeping without divisions: `head` = ring slot of the chromosome being evaluated,
    // `phase` = one parity bit per slot (flips every time the slot's mbarrier completes a phase)
    int head = 0;
    uint32_t phase = 0;
    const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    auto next_slot = [](int st) { return st + 1 == kSt ? 0 : st + 1; };

    const int ntiles = (P + T - 1) / T;
    for (int tile = blockIdx.x * kWarpsPerBlock + warp; tile < ntiles; tile += gridDim.x * kWarpsPerBlock) {
        const int slot0 = tile * T;
        const int cnt = min(T, P - slot0);
        // ---- phase A (lane = slot)
        ChildPlan<NW> mine;
        if constexpr (GA) {
            mine = plan_slot<NW, float>(lane < T ? slot0 + lane : P, ne, fit, nullptr, top_idx, tape, seed, gen, ga, P, W, sel, cpo);
        } else {
            mine = plan_elite<NW>(min(slot0 + lane, P - 1), W);
        }
        TileSums<NW> sums;
        sums.reset();

        // gather child c's two parent pieces into ring slot st: the lane that OWNS child c issues the two TMA
        // copies from its own plan registers (no broadcast of the plan to an elected lane)
        auto issue = [&](int c, int st) {
            if (lane == c) {
                const uint32_t dst = ring_u32 + (uint32_t)(st * kSl) * (W * 16u), bar = bars_u32 + (uint32_t)st * 8u;
                if constexpr (UX) {   // both parents, whole rows
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * W * 16u) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                                 "l"(pop + (size_t)mine.first * W), "r"((uint32_t)W * 16u), "r"(bar), "l"(pol_stream) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                                     dst + (uint32_t)W * 16u),
                                 "l"(pop + (size_t)mine.second * W), "r"((uint32_t)W * 16u), "r"(bar), "l"(pol_stream) : "memory");
                } else {
                    const int cpc = min(meta_cp(mine.meta), W);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)W * 16u) : "memory");
                    if (cpc > 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                                     "l"(pop + (size_t)mine.first * W), "r"((uint32_t)cpc * 16u), "r"(bar), "l"(pol_stream) : "memory");
                    if (cpc < W)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                                         dst + (uint32_t)cpc * 16u),
                                     "l"(pop + (size_t)mine.second * W + cpc), "r"((uint32_t)(W - cpc) * 16u), "r"(bar), "l"(pol_stream) : "memory");
                }
            }
        };

        // wait for child c's row (ring slot st), apply its mutation in shared memory, start its voxel gathers
        auto prepare = [&](int c, int st, uint32_t (&vword)[NW], uint32_t (&vmask)[NW]) {
            float4* row = ring + st * kSl * W;
            mbar_wait(&bars[st], (phase >> st) & 1u);
            phase ^= 1u << st;
            if constexpr (UX) {   // masked waypoints come from the second parent (staged right behind the first)
#pragma unroll
                for (int s = 0; s < NW; ++s) {
                    const uint32_t m = __shfl_sync(0xffffffffu, mine.mask[s], c);
                    if ((m >> lane) & 1u) row[lane + 32 * s] = row[W + lane + 32 * s];
                }
                __syncwarp();
            }
            if constexpr (GA) {
                if (lane == c) {  // the owning lane applies its child's mutation (HR:97-98)
                    const int mp = meta_mp(mine.meta);
                    if (mp >= 0) {
                        float4 v = row[mp];
                        v.x = __fadd_rn(v.x, mine.dx);
                        v.y = __fadd_rn(v.y, mine.dy);
                        v.z = __fadd_rn(v.z, mine.dz);
                        row[mp] = v;
                    }
                }
                __syncwarp();
            }
            if constexpr (VOX) {
                Lanes<NW> p;
#pragma unroll
                for (int s = 0; s < NW; ++s) {
                    const float4 v = row[lane + 32 * s];
                    p.x[s] = v.x; p.y[s] = v.y; p.z[s] = v.z;
                }
                VoxRegs<NW> vr;
                voxel_prefetch<NW>(e, p, vr, pol_keep);
#pragma unroll
                for (int s = 0; s < NW; ++s) { vword[s] = vr.word[s]; vmask[s] = vr.mask[s]; }
            }
        };

        // all ring slots are free at a tile boundary: slots head, head+1, ... take children 0, 1, ...
        const int pre = min(cnt, kSt);
        {
            int st = head;
            for (int c = 0; c < pre; ++c) { issue(c, st); st = next_slot(st); }
        }
        uint32_t v0w[NW], v0m[NW], v1w[NW], v1m[NW];
#pragma unroll
        for (int s = 0; s < NW; ++s) v0w[s] = v0m[s] = v1w[s] = v1m[s] = 0u;
        prepare(0, head, v0w, v0m);
#pragma unroll 1
        for (int c = 0; c < cnt; ++c) {
            const int st_next = next_slot(head);
            if (c + 1 < cnt) prepare(c + 1, st_next, v1w, v1m);
            const float4* row = ring + head * kSl * W;
            Lanes<NW> cur, prev;
            float4* dst = GA ? next + (size_t)(slot0 + c) * W : nullptr;
#pragma unroll
            for (int s = 0; s < NW; ++s) {
                const int k = lane + 32 * s;
                const float4 v = row[k];
                const float4 q = row[k > 0 ? k - 1 : 0];
                cur.x[s] = v.x; cur.y[s] = v.y; cur.z[s] = v.z;
                prev.x[s] = q.x; prev.y[s] = q.y; prev.z[s] = q.z;
                if constexpr (GA) stg_stream_hint(dst + k, v, pol_stream);
            }
            PointAcc<NW, SEM> acc;
            acc.reset();
            if constexpr (SEM != TF_SEM_NONE) eval_static<SEM, NW>(e, tabs, cur.x, cur.y, cur.z, acc.hit, acc.best);
            reduce_fullw<NW, SEM, VOX, XF>(e, prm, tabs, cur, prev, v0w, v0m, acc, c, sums);
            __syncwarp();  // every lane is done with this ring slot: refill it
            if (c + kSt < cnt) issue(c + kSt, head);
            head = st_next;
#pragma unroll
            for (int s = 0; s < NW; ++s) { v0w[s] = v1w[s]; v0m[s] = v1m[s]; }
        }
        // ---- phase C (lane = slot)
        if (lane < cnt) {
            const int slot = slot0 + lane;
            out_fit[slot] = combine_fitness<NW, XF>(e, prm, sums);
            if (out_coll) {
#pragma unroll
                for (int s = 0; s < NW; ++s) out_coll[(size_t)slot * NW + s] = sums.cw[s];
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// "duo" schedule of the same kernel: a warp evaluates TWO chromosomes at a time, one per half-warp.
// Lane h = lane & 15 of a half owns waypoints h, h+16, h+32, ... of its half's chromosome (W/16 per lane).
// Why: with one chromosome per warp-iteration the per-chromosome overhead (mbarrier wait, TMA issue, five
// butterfly stages per total, loop bookkeeping) is about half of the 330 warp-instructions a chromosome
// costs, and the kernel is issue-bound (ncu round 1: issue-active 78 %).  Two chromosomes per iteration
// share the wait / loop / reduction instructions, and the butterfly needs only stages 8, 4, 2, 1 across
// lanes: stage 16 is a register add, because lane h holds the partial of "virtual lane" h (even m) and of
// virtual lane h + 16 (odd m) of the FP32 model (oracle: lane l owns waypoints l, l+32, ...; tree32).
// Every floating-point add has the same two operands as in the one-chromosome schedule, so the results
// are bit-identical (tests run both through TF_OPT_STREAM_SCHEDULE).
// Tile bookkeeping: lanes 0-15 own children [0, T/2) of the tile, lanes 16-31 own [T/2, T); pair j =
// children j and T/2 + j = lanes j and 16 + j, so every total ends in the half where its owner sits.
// XF 1 (hazard map / moving spheres): 3 CTAs / SM (80 registers; the tile sums spill 128 bytes to local memory,
// outside the waypoint loops).  Round 2, cfg4 in the driver's regime (generation 6): 1.32 ms at 2 CTAs / SM
// (128 registers), 1.25 ms at 3 — the kernel is latency- / issue-bound, not bandwidth-bound (profiles/README.md).
// TF_DUO_XF1_CTAS is an A/B knob for build.py (TF_NVCC_EXTRA).
// RAGW: the row length is a runtime value Wr <= 32 NW (any chromosome length up to 64 on the TMA-staged path);
// lanes whose waypoint index is >= Wr are masked.  RAGW = false: W = 32 NW, the code of the BASELINE configs.
#ifndef TF_PROBE_FIRST
#define TF_PROBE_FIRST 1
#endif
#ifndef TF_XF1_L2_HINTS // A/B: 0 = the hazard / moving-sphere variants of duo_kernel drop the L2 eviction hints (the two
#define TF_XF1_L2_HINTS 1 //      createpolicy results do not stay in registers there: 18 instructions per pair to re-create them)
#endif
#ifndef TF_BATCH_MATH   // A/B: 0 = one sqrt_rn / rcp_rn call per waypoint, 1 = batched per lane in the XF 1 variants, 2 = everywhere
#define TF_BATCH_MATH 1
#endif
// Row staging of duo_kernel: the owning lane's two TMA bulk copies, or — kGather — per-lane 16-byte cp.async copies
// (every lane fetches its own waypoints from whichever parent covers them; the plan comes from the owning lane by three
// shuffles).  Bit-identical.  Measured (round 2, same box, default run): cfg3 0.3757 -> 0.3733 ms, cfg4 0.9529 -> 0.9175 ms:
// the hazard / moving-sphere variant (XF 1) is instruction-bound and short of registers (the TMA issue sequence costs 78
// instructions per pair and 124 bytes of spills), so XF 1 gathers per lane; the plain variant keeps TMA.
// TF_ROW_GATHER: 0 = that default, 1 = per-lane everywhere, -1 = TMA everywhere (A/B).
#ifndef TF_ROW_GATHER
#define TF_ROW_GATHER 0
#endif
#ifndef TF_GA_TAIL_TILES   // A/B: 0 = the plain fused generation runs on half-size tiles only; n = tiles of T with n
#define TF_GA_TAIL_TILES 0 //      half-size tiles per resident warp at the end (re-measured with the 4-buffer ring: 0 wins)
#endif
#ifndef TF_DUO_XF1_CTAS
#define TF_DUO_XF1_CTAS 3
#endif
#ifndef TF_DUO_XF1_WPB   // A/B: warps per CTA of the hazard / moving-sphere variants (W <= 64), e.g. 5 with TF_DUO_XF1_CTAS 4 = 20 warps at 96
#define TF_DUO_XF1_WPB 8 // registers instead of 8 x 3 = 24 at 80 — measured: 5 x 4 CTAs (20 warps, 96 registers) 0.807 vs 0.801 ms, 6 x 3 0.857 ms
#endif
// CTA shape of duo_kernel: 8 warps; the XF 1 variants on rows of up to 64 waypoints take theirs from the knob above
template <int NW, int XF>
constexpr int duo_wpb() { return (XF == 1 && NW <= 2) ? TF_DUO_XF1_WPB : kWarpsPerBlock; }
template <int NW, int SEM, bool VOX, int XF, bool GA, bool RAGW>
__global__ void __launch_bounds__(duo_wpb<NW, XF>() * 32, NW >= 8 ? 1 : NW >= 4 ? 2 : ((XF == 0 && SEM == TF_SEM_NONE) ? 4 : (XF == 1 ? TF_DUO_XF1_CTAS : 0)))
duo_kernel(const __grid_constant__ EnvDev e, const __grid_constant__ tf_params prm,
           const __grid_constant__ tf_ga_params ga, const float4* __restrict__ pop,
           const float* __restrict__ fit, const int32_t* __restrict__ top_idx,
           const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen, int P, int T,
           float4* __restrict__ next, float* __restrict__ out_fit, uint32_t* __restrict__ out_coll,
           int32_t* __restrict__ sel, int32_t* __restrict__ cpo, int n_big, int n_tickets, int* __restrict__ sched,
           int Wr, unsigned long long* __restrict__ rank_part, int32_t* __restrict__ next_top, int next_k) {
    constexpr int NL = 2 * NW, kPairStages = kStages / 2, kWpb = duo_wpb<NW, XF>();
    constexpr bool kGather = TF_ROW_GATHER > 0 || (TF_ROW_GATHER == 0 && XF == 1);
    const int W = RAGW ? Wr : 32 * NW;
    extern __shared__ __align__(128) unsigned char smem_stream[];
    unsigned char* smem = smem_stream;
    const SmemLayout sm = carve_smem(e, smem);
    const Tables tabs{sm.A, sm.B, sm.dA, sm.dB};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane & 15, half = lane >> 4;
    const size_t tables = (smem_bytes_for(e) + 127) & ~(size_t)127;
    float4* ring = reinterpret_cast<float4*>(smem + tables) + (size_t)warp * kStages * W;   // kPairStages x 2 rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tables + (size_t)kWpb * kStages * W * 16) + warp * kStages;
    const int ne = GA ? min(ga.n_elite, P) : 0;
    // fused ranking (GA, next_k = 1 or 2, dynamic schedule): the best keys this warp has produced so far, and the
    // CTA's count of warps that ran dry — behind the mbarriers
    unsigned long long* wbest = reinterpret_cast<unsigned long long*>(bars - warp * kStages + kWpb * kStages) + warp * 2;
    unsigned int* cta_done = reinterpret_cast<unsigned int*>(wbest - warp * 2 + kWpb * 2);
    const bool ranking = GA && next_k > 0;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kPairStages; ++i) mbar_init(&bars[i], 2);   // one arrival per half
        if (ranking) { wbest[0] = ~0ull; wbest[1] = ~0ull; }
    }
    if (ranking && threadIdx.x == 0) *cta_done = 0u;
    if ((SEM != TF_SEM_NONE && e.n_smem > 0) || (XF >= 1 && e.n_dyn_smem > 0)) {
        if (threadIdx.x == 0) mbar_init(sm.bar, 1);
        fence_mbar_init();
        __syncthreads();
        if (threadIdx.x == 0)
            stage_chunk(e, sm.A, sm.B, 0, min(e.n_smem, kTableCap), sm.bar, XF >= 1 && e.n_dyn_smem > 0, sm.dA, sm.dB);
        mbar_wait(sm.bar, 0);
    } else {
        fence_mbar_init();
        if (ranking) __syncthreads();   // the CTA's counter is zeroed before any warp can run dry
        else __syncwarp();
    }

    int head = 0;
    uint32_t phase = 0;
    const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
    constexpr bool kHints = !(XF == 1 && TF_XF1_L2_HINTS == 0);
    const uint64_t pol_stream = kHints ? l2_policy_evict_first() : 0ull, pol_keep = kHints ? l2_policy_evict_last() : 0ull;
    auto next_stage = [](int st) { return st + 1 == kPairStages ? 0 : st + 1; };
    const bool has_dist = (SEM != TF_SEM_NONE) || (XF >= 1 && e.n_dyn > 0);

    // Dynamic schedule: the grid is persistent (one CTA per resident slot) and every warp draws tile tickets
    // from a global counter.  Tickets [0, n_big) are tiles of T children, the rest are tiles of T/2 that
    // cover the end of the population, so all warps run dry within half a tile of each other (ncu round 1,
    // static schedule: SMs active 92 % of the kernel, 28.7 of 32 warp slots filled while active).  The next
    // ticket is drawn at the top of a tile, so the atomic's latency hides behind the tile's work.
    // Short runs (sched == nullptr) keep the static assignment: ticket = global warp index, then stride.
    const bool dynamic = sched != nullptr;
    auto draw = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(sched, 1);
        return t;
    };
    int ticket = dynamic ? __shfl_sync(0xffffffffu, draw(), 0) : (int)(blockIdx.x * kWpb) + warp;
    while (ticket < n_tickets) {
        const int ahead = dynamic ? draw() : ticket + (int)(gridDim.x * kWpb);
        const int Tt = ticket < n_big ? T : (T >> 1), hT = Tt >> 1;
        const int slot0 = ticket < n_big ? ticket * T : n_big * T + (ticket - n_big) * Tt;
        const int cnt = min(Tt, P - slot0);
        // ---- phase A (lane = child of the tile)
        const int myc = half * hT + h;
        const bool own = h < hT && myc < cnt;
        ChildPlan<NW> mine;
        if constexpr (GA) {
            mine = plan_slot<NW, float>(own ? slot0 + myc : P, ne, fit, nullptr, top_idx, tape, seed, gen, ga, P, W, sel, cpo);
        } else {
            mine = plan_elite<NW>(min(slot0 + myc, P - 1), W);
        }
        TileSums<NW> sums;
        sums.reset();

        // pair j -> stage st: each owning lane gathers its own child's two parent pieces (or just arrives
        // when the pair has no second child)
        auto issue = [&](int j, int st) {
          if constexpr (kGather) {
            // every lane copies its own waypoints of its half's child (16 bytes each, L1 bypassed) from whichever
            // parent covers them; the child's plan comes from its owning lane by three shuffles
            const int src_lane = (lane & 16) | j;
            const int pfirst = __shfl_sync(0xffffffffu, mine.first, src_lane), psecond = __shfl_sync(0xffffffffu, mine.second, src_lane);
            const int cpc = min(meta_cp(__shfl_sync(0xffffffffu, mine.meta, src_lane)), W);
            const bool has = __shfl_sync(0xffffffffu, (int)own, src_lane) != 0;
            if (has) {
                const uint32_t dst = ring_u32 + (uint32_t)(st * 2 + half) * (W * 16u) + (uint32_t)h * 16u;
                const float4* pa = pop + (size_t)pfirst * W + h;
                const float4* pb = pop + (size_t)psecond * W + h;
#pragma unroll
                for (int m = 0; m < NL; ++m) {
                    const int k = h + 16 * m;
                    if (RAGW && k >= W) continue;
                    const float4* src = (k < cpc ? pa : pb) + 16 * m;
                    if constexpr (kHints)
                        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)m * 256u), "l"(src), "l"(pol_stream) : "memory");
                    else
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)m * 256u), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
          } else {
            if (h == j) {
                const uint32_t bar = bars_u32 + (uint32_t)st * 8u;
                if (own) {
                    const int cpc = min(meta_cp(mine.meta), W);
                    const uint32_t dst = ring_u32 + (uint32_t)(st * 2 + half) * (W * 16u);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)W * 16u) : "memory");
                    if (cpc > 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                                     "l"(pop + (size_t)mine.first * W), "r"((uint32_t)cpc * 16u), "r"(bar), "l"(pol_stream) : "memory");
                    if (cpc < W)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                                         dst + (uint32_t)cpc * 16u),
                                     "l"(pop + (size_t)mine.second * W + cpc), "r"((uint32_t)(W - cpc) * 16u), "r"(bar), "l"(pol_stream) : "memory");
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                }
            }
          }
        };

        const int npairs = min(cnt, hT);
        {
            int st = head;
            if constexpr (kGather) {
                for (int j = 0; j < kPairStages; ++j) {   // one group per stage, empty past the end of the tile
                    if (j < npairs) issue(j, st);
                    else asm volatile("cp.async.commit_group;" ::: "memory");
                    st = next_stage(st);
                }
            } else {
                for (int j = 0; j < min(npairs, kPairStages); ++j) { issue(j, st); st = next_stage(st); }
            }
        }
#pragma unroll 1
        for (int j = 0; j < npairs; ++j) {
            const bool hasB = hT + j < cnt;                       // warp-uniform
            const bool live = half == 0 || hasB;                  // this half has a chromosome of its own
            float4* row = ring + (head * 2 + (hasB ? half : 0)) * W;
            if constexpr (kGather) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kPairStages - 1) : "memory");   // pair j's group has landed
                __syncwarp();
            } else {
                mbar_wait(&bars[head], (phase >> head) & 1u);
                phase ^= 1u << head;
            }
            if constexpr (GA) {
                if (h == j && own) {  // the owning lane applies its child's mutation (HR:97-98)
                    const int mp = meta_mp(mine.meta);
                    if (mp >= 0) {
                        float4 v = row[mp];
                        v.x = __fadd_rn(v.x, mine.dx);
                        v.y = __fadd_rn(v.y, mine.dy);
                        v.z = __fadd_rn(v.z, mine.dz);
                        row[mp] = v;
                    }
                }
                __syncwarp();
            }
            // per-lane partials of virtual lanes h (even m) and h + 16 (odd m), accumulated in waypoint order.
            // XF 1: segment sums first (length, altitude, turns: they only need the row), so that the neighbour
            // waypoints are dead before the gather-heavy part (voxels, moving spheres, hazard map) starts.
            // XF 0 with a voxel grid: the probes go out FIRST and the segment sums (4 square roots per lane) run
            // under their latency — the probe is the one long-latency load of the step and its words are only
            // needed by the ballots at the end.
            constexpr bool kProbeFirst = VOX && XF == 0 && (TF_PROBE_FIRST != 0);
            Lanes<NL> cur;
            float a_prox[2] = {0.0f, 0.0f}, a_dist[2] = {0.0f, 0.0f}, a_alt[2] = {0.0f, 0.0f}, a_haz[2] = {0.0f, 0.0f};
            int turns = 0, ncoll = 0;
            float4* dst = GA ? next + (size_t)(slot0 + half * hT + j) * W : nullptr;
            auto fetch = [&](int m) {     // waypoint h + 16 m: into registers and out to the child row
                const int k = h + 16 * m;
                const bool ok = !RAGW || k < W;   // masked lanes sit at (-1, -1, -1): outside every grid
                const float4 c4 = ok ? row[k] : make_float4(-1.0f, -1.0f, -1.0f, 0.0f);
                cur.x[m] = c4.x; cur.y[m] = c4.y; cur.z[m] = c4.z;
                if constexpr (GA) {
                    if (live && ok) {
                        if constexpr (kHints) stg_stream_hint(dst + k, c4, pol_stream);
                        else stg_stream(dst + k, c4);
                    }
                }
            };
            // kBatch (XF 1): the square roots / reciprocals of a lane's NL waypoints are taken together (sqrt_rn_n,
            // rcp_rn_n: one straight-line block behind one rare-input branch instead of NL branches around NL calls);
            // the sums still take their terms in waypoint order, so every add has the same operands
            constexpr bool kBatch = TF_BATCH_MATH > 1 || (TF_BATCH_MATH == 1 && XF == 1);
            float seg2[NL];
            auto segment = [&](int m) {   // segment (k - 1, k)
                const int k = h + 16 * m, v = m & 1;
                const bool ok = !RAGW || k < W;
                if constexpr (kBatch) seg2[m] = 0.0f;
                if (ok && (m > 0 || h > 0)) {  // k >= 1
                    const float4 q = row[k - 1];
                    const float dx = cur.x[m] - q.x, dy = cur.y[m] - q.y, dz = cur.z[m] - q.z;
                    if constexpr (kBatch) seg2[m] = dist2(dx, dy, dz);
                    else a_dist[v] = __fadd_rn(a_dist[v], sqrt_rn(dist2(dx, dy, dz)));  // HR:51-52
                    a_alt[v] = __fadd_rn(a_alt[v], fabsf(dz));                      // HR:44
                    if (RAGW ? k <= W - 2 : (m < NL - 1 || h < 15)) turns += (cur.x[m] != q.x || cur.y[m] != q.y) ? 1 : 0;  // k <= W-2, HR:38-40,43
                }
            };
            auto segment_lengths = [&]() {
                if constexpr (kBatch) {
                    float len[NL];
                    sqrt_rn_n<NL>(seg2, len);
#pragma unroll
                    for (int m = 0; m < NL; ++m) {
                        const int k = h + 16 * m;
                        if ((!RAGW || k < W) && (m > 0 || h > 0)) a_dist[m & 1] = __fadd_rn(a_dist[m & 1], len[m]);  // HR:51-52
                    }
                }
            };
            VoxRegs<NL> vr;
            if constexpr (kProbeFirst) {
#pragma unroll
                for (int m = 0; m < NL; ++m) fetch(m);
                voxel_prefetch<NL, kHints>(e, cur, vr, pol_keep);
#pragma unroll
                for (int m = 0; m < NL; ++m) segment(m);
                segment_lengths();
            } else {
#pragma unroll
                for (int m = 0; m < NL; ++m) { fetch(m); segment(m); }
                segment_lengths();
                if constexpr (VOX) voxel_prefetch<NL, kHints>(e, cur, vr, pol_keep);
            }
            PointAcc<NL, SEM> acc;
            acc.reset();
            if constexpr (SEM != TF_SEM_NONE) eval_static<SEM, NL>(e, tabs, cur.x, cur.y, cur.z, acc.hit, acc.best);
            uint32_t hit = acc.hit;
            if constexpr (XF >= 1) {
                // the hazard taps go out behind the voxel probes; then the probes are folded into `hit`, so that
                // neither their words nor their masks are live during the candidate walk of the moving spheres
                if (e.hazard != nullptr) {
#pragma unroll
                    for (int m = 0; m < NL; ++m) {
                        const int k = h + 16 * m;
                        if (RAGW && k >= W) continue;
                        a_haz[m & 1] = __fadd_rn(a_haz[m & 1], hazard_sample(e, cur.x[m], cur.y[m], (float)k));
                    }
                }
                if constexpr (VOX) {
#pragma unroll
                    for (int m = 0; m < NL; ++m) hit |= (vr.word[m] & vr.mask[m]) != 0u ? (1u << m) : 0u;
                }
            }
            float dn[NL];
#pragma unroll
            for (int m = 0; m < NL; ++m) dn[m] = (SEM == TF_SEM_HR) ? sqrt_rn(acc.best[m]) : acc.best[m];
            if constexpr (XF >= 1) {
                if (e.n_dyn > 0) {
                    float m2[NL];
#pragma unroll
                    for (int m = 0; m < NL; ++m) m2[m] = INFINITY;
                    eval_dynamic<NL>(e, tabs, cur.x, cur.y, cur.z, h, 16, hit, m2);
                    if constexpr (kBatch) {
                        float d[NL];
                        sqrt_rn_n<NL>(m2, d);
#pragma unroll
                        for (int m = 0; m < NL; ++m) dn[m] = fminf(dn[m], d[m]);
                    } else {
#pragma unroll
                        for (int m = 0; m < NL; ++m) dn[m] = fminf(dn[m], sqrt_rn(m2[m]));
                    }
                }
            }
            uint32_t cwords[NL];
            if (has_dist) {
                float term[NL];
                if constexpr (kBatch) {
                    float den[NL];
#pragma unroll
                    for (int m = 0; m < NL; ++m) den[m] = __fadd_rn(1.0f, dn[m]);
                    rcp_rn_n<NL>(den, term);
                } else {
#pragma unroll
                    for (int m = 0; m < NL; ++m) term[m] = rcp_rn(__fadd_rn(1.0f, dn[m]));
                }
#pragma unroll
                for (int m = 0; m < NL; ++m) {
                    if (RAGW && h + 16 * m >= W) continue;
                    a_prox[m & 1] = __fadd_rn(a_prox[m & 1], term[m]);  // HR:46
                }
            }
#pragma unroll
            for (int m = 0; m < NL; ++m) {
                bool hm = (hit >> m) & 1u;
                if constexpr (VOX && XF == 0) hm = hm || (vr.word[m] & vr.mask[m]) != 0u;
                if (RAGW) hm = hm && h + 16 * m < W;
                cwords[m] = __ballot_sync(0xffffffffu, hm);
                ncoll += hm ? 1 : 0;
            }
            // totals: stage 16 of the butterfly is the add of the two virtual lanes, stages 8..1 stay in the half
            auto half_sum = [](const float (&a)[2]) {
                float r = __fadd_rn(a[0], a[1]);
#pragma unroll
                for (int x = 8; x >= 1; x >>= 1) r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, x));
                return r;
            };
            const float dist = half_sum(a_dist), alt = half_sum(a_alt);
            // integer totals of both halves in one reduction: 16 bits per half (7 + 9 bits up to W = 128; longer
            // rows take one reduction per count)
            uint32_t turns_tot, ncoll_tot;
            if constexpr (NW <= 4) {
                const uint32_t both = __reduce_add_sync(0xffffffffu, (uint32_t)(turns | (ncoll << 7)) << (half * 16));
                const uint32_t t16 = (both >> (half * 16)) & 0xffffu;
                turns_tot = t16 & 127u;
                ncoll_tot = t16 >> 7;
            } else {
                turns_tot = (__reduce_add_sync(0xffffffffu, (uint32_t)turns << (half * 16)) >> (half * 16)) & 0xffffu;
                ncoll_tot = (__reduce_add_sync(0xffffffffu, (uint32_t)ncoll << (half * 16)) >> (half * 16)) & 0xffffu;
            }
            if (h == j) {
                sums.dist = dist; sums.alt = alt;
                sums.packed = (int)(turns_tot | (ncoll_tot << 10));
#pragma unroll
                for (int w = 0; w < NW; ++w)
                    sums.cw[w] = __byte_perm(cwords[2 * w], cwords[2 * w + 1], half ? 0x7632 : 0x5410);
            }
            if (has_dist) {
                const float prox = half_sum(a_prox);
                if (h == j) sums.prox = prox;
            }
            if constexpr (XF >= 1) {
                if (e.hazard != nullptr) {
                    const float haz = half_sum(a_haz);
                    if (h == j) sums.haz = haz;
                }
            }
            __syncwarp();  // every lane is done with this stage: refill it
            if (j + kPairStages < npairs) issue(j + kPairStages, head);
            else if constexpr (kGather) asm volatile("cp.async.commit_group;" ::: "memory");
            head = next_stage(head);
        }
        // ---- phase C (lane = child of the tile)
        unsigned long long key = ~0ull;
        if (own) {
            const int slot = slot0 + myc;
            const float f = combine_fitness<NW, XF>(e, prm, sums);
            out_fit[slot] = f;
            if (ranking) key = rank_key32(f, (uint32_t)slot);
            if (out_coll) {
                const int nwords = RAGW ? (W + 31) >> 5 : NW;   // the flag array is [P][ceil(W / 32)]
#pragma unroll
                for (int s = 0; s < NW; ++s)
                    if (s < nwords) out_coll[(size_t)slot * nwords + s] = sums.cw[s];
            }
        }
        if constexpr (GA) {
            if (ranking) rank_note(key, wbest, next_k, lane);
        }
        ticket = dynamic ? __shfl_sync(0xffffffffu, ahead, 0) : ahead;
    }
    if constexpr (GA) {
        if (ranking) rank_merge_cta(wbest - warp * 2, cta_done, rank_part, lane, kWpb);
    }
    // the last warp of the grid to run dry hands the counters back zeroed for the next launch and merges the CTAs'
    // ranking keys
    if (dynamic) {
        int last = 0;
        if (lane == 0) last = atomicAdd(sched + 1, 1) == (int)(gridDim.x * kWpb) - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if constexpr (GA) {
                if (ranking) rank_merge_grid(rank_part, next_top, next_k, lane);
            }
            if (lane == 0) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
    }
}

template <int NW, int SEM, bool VOX, int XF, bool GA, bool RAGW>
static int launch_duo(const StreamArgs& a, int T) {
    auto kern = duo_kernel<NW, SEM, VOX, XF, GA, RAGW>;
    constexpr int kWpb = duo_wpb<NW, XF>();
    const size_t smem = stream_smem_bytes(*a.e, a.W, kWpb);
    if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // resident CTAs per SM of this variant with this much shared memory.  One packed word (smem << 8 | occ): a
    // racing thread at worst repeats the query; the answer is the same on every B200 of a node.
    static unsigned long long occ_cache = 0;
    const unsigned long long cached = __atomic_load_n(&occ_cache, __ATOMIC_RELAXED);
    int occ = (int)(cached & 0xffu);
    if ((cached >> 8) != (unsigned long long)smem || occ <= 0) {
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWpb * 32, smem));
        __atomic_store_n(&occ_cache, ((unsigned long long)smem << 8) | (unsigned long long)(occ & 0xff), __ATOMIC_RELAXED);
    }
    const int resident_warps = sm_count() * max(occ, 1) * kWpb;
    // long runs (>= 4 tiles per resident warp): persistent grid, tickets from a global counter.  The plain fused
    // generation runs entirely on half-size tiles (16 children; measured at cfg3: tiles of 32 0.387 ms, 32 with
    // a tail of 2 / 4 / 8 half-size tiles per resident warp 0.385 / 0.378 / 0.377 ms, all 16 0.374 ms, all 8
    // 0.391 ms); the fitness sweeps and the hazard / moving-sphere variants keep tiles of 32 and end in two
    // half-size tiles per resident warp (all-16 there: fitness sweep without environment 0.192 -> 0.252 ms,
    // cfg4 1.17 -> 1.23 ms).  Short runs: static assignment.
    int n_big = (a.P + T - 1) / T, n_small = 0;
    int* sched = nullptr;
    int grid = min((n_big + kWpb - 1) / kWpb, sm_count() * 8);
    if (T >= 8 && a.P / T >= 4 * resident_warps) {
        const int Ts = T >> 1;
        n_big = (GA && XF == 0 && TF_GA_TAIL_TILES == 0) ? 0 : (a.P - max(TF_GA_TAIL_TILES, 2) * resident_warps * Ts) / T;
        n_small = (a.P - n_big * T + Ts - 1) / Ts;
        sched = sched_slot();
        if (sched == nullptr) return TF_ERR_NOMEM;
        // TF_OPT_RESERVE_SMS: leave some SMs to kernels of other streams (the island model's elite exchange
        // runs beside the generation; a full persistent grid would hold every CTA slot until it ends)
        const int reserve = max(g_reserve_sms, (GA && a.ga) ? a.ga->reserve_sms : 0);
        grid = max(sm_count() - max(reserve, 0), sm_count() / 2) * max(occ, 1);
    }
    // fused ranking of the children: only on the persistent grid (its last warp is the one that merges)
    const bool rank = GA && sched != nullptr && a.next_top != nullptr && (a.next_k == 1 || a.next_k == 2) &&
                      a.rank_part != nullptr && a.rank_part_bytes >= (int64_t)grid * 16 && a.P >= a.next_k;
    kern<<<grid, kWpb * 32, smem, a.st>>>(*a.e, *a.prm, *a.ga, reinterpret_cast<const float4*>(a.pop), a.fit,
                                              a.top_idx, a.tape, a.seed, a.gen, a.P, T,
                                              reinterpret_cast<float4*>(a.next), a.out_fit, a.out_coll, a.sel, a.cpo,
                                              n_big, n_big + n_small, sched, a.W, a.rank_part, a.next_top, rank ? a.next_k : 0);
    TF_LAUNCH_CHECK();
    a.ranked = rank;
    return TF_OK;
}

template <int NW, int SEM, bool VOX, int XF, bool GA>
static int launch_stream(const StreamArgs& a) {
    const size_t smem = stream_smem_bytes(*a.e, a.W);
    const int T = tile_size_for(a.P);
    // schedule: two chromosomes per warp step wherever there is no obstacle-table loop (measured faster on the
    // voxel / hazard / moving-sphere variants); with a table loop the extra registers (104 vs 64) halve the
    // occupancy and one chromosome per warp step stays ahead (cfg2: 56 us vs 71 us)
    // uniform crossover and S > 1 samples per segment (XF 2) exist in the one-chromosome-per-step kernel only
    const bool duo = XF < 2 && !a.uniform && (g_stream_schedule == 0 ? SEM == TF_SEM_NONE : g_stream_schedule == 2);
    if (!duo) {  // one chromosome per warp step, static tile assignment
        auto kern = stream_kernel<NW, SEM, VOX, XF, GA>;
        if constexpr (GA && XF < 2) {
            if (a.uniform) kern = stream_kernel<NW, SEM, VOX, XF, GA, true>;
        }
        if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int ntiles = (a.P + T - 1) / T;
        const int grid = min((ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
        // (no fused ranking here: the ranking state costs this kernel 16 registers — 80 instead of 64, 3 instead
        // of 4 CTAs / SM, cfg2 generation 61 -> 87 us — more than the 14 us ranking launch it would save;
        // tf_generation_ranked runs tf_rank_topk behind these variants)
        kern<<<grid, kBlockThreads, smem, a.st>>>(*a.e, *a.prm, *a.ga, reinterpret_cast<const float4*>(a.pop), a.fit,
                                                  a.top_idx, a.tape, a.seed, a.gen, a.P, T,
                                                  reinterpret_cast<float4*>(a.next), a.out_fit, a.out_coll, a.sel, a.cpo);
        TF_LAUNCH_CHECK();
        return TF_OK;
    }
    if constexpr (XF < 2) {
        if (a.W != 32 * NW) {
            if constexpr (SEM == TF_SEM_NONE) return launch_duo<NW, SEM, VOX, XF, GA, true>(a, T);
            else return TF_ERR_UNSUPPORTED;  // stream_eligible() never sends these here
        }
        return launch_duo<NW, SEM, VOX, XF, GA, false>(a, T);
    } else {
        return TF_ERR_UNSUPPORTED;
    }
}
template <int NW, int SEM, bool GA>
static int stream_flags(const StreamArgs& a) {
    const bool vox = a.e->bricks != nullptr;
    const bool xf1 = a.e->n_dyn > 0 || a.e->hazard != nullptr;
    if (a.prm->samples_per_segment > 1)   // XF 2: the segment-sampling form (W = 32 / 64, no uniform crossover: stream_eligible)
        return vox ? launch_stream<NW, SEM, true, 2, GA>(a) : launch_stream<NW, SEM, false, 2, GA>(a);
    if (vox) return xf1 ? launch_stream<NW, SEM, true, 1, GA>(a) : launch_stream<NW, SEM, true, 0, GA>(a);
    return xf1 ? launch_stream<NW, SEM, false, 1, GA>(a) : launch_stream<NW, SEM, false, 0, GA>(a);
}
// 64 < W <= 256 (NW = 4, 8): only the two-chromosome kernel without an obstacle table is built (8 / 16
// waypoints per lane; the ring is 12 / 24 KB per warp, so 2 CTAs / SM and 128 registers, or 1 and 255)
template <int NW, bool GA>
static int stream_long_rows(const StreamArgs& a) {
    const int T = tile_size_for(a.P);
    const bool vox = a.e->bricks != nullptr;
    const bool xf1 = a.e->n_dyn > 0 || a.e->hazard != nullptr;
    if (a.W == 32 * NW) {
        if (vox) return xf1 ? launch_duo<NW, TF_SEM_NONE, true, 1, GA, false>(a, T) : launch_duo<NW, TF_SEM_NONE, true, 0, GA, false>(a, T);
        return xf1 ? launch_duo<NW, TF_SEM_NONE, false, 1, GA, false>(a, T) : launch_duo<NW, TF_SEM_NONE, false, 0, GA, false>(a, T);
    }
    if (vox) return xf1 ? launch_duo<NW, TF_SEM_NONE, true, 1, GA, true>(a, T) : launch_duo<NW, TF_SEM_NONE, true, 0, GA, true>(a, T);
    return xf1 ? launch_duo<NW, TF_SEM_NONE, false, 1, GA, true>(a, T) : launch_duo<NW, TF_SEM_NONE, false, 0, GA, true>(a, T);
}
template <int NW, bool GA>
static int stream_sem(const StreamArgs& a) {
    switch (a.e->n > 0 ? a.e->semantics : TF_SEM_NONE) {
    case TF_SEM_HR: return stream_flags<NW, TF_SEM_HR, GA>(a);
    case TF_SEM_PF: return stream_flags<NW, TF_SEM_PF, GA>(a);
    default: return stream_flags<NW, TF_SEM_NONE, GA>(a);
    }
}
}  // namespace tf
