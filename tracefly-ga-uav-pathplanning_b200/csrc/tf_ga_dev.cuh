// tf_ga_dev.cuh — device code of selection + crossover + mutation.
//
// Replaces HRMAGAGenerator.evolve's inner loop HR:114-122 (elitism HR:115, truncation
// selection HR:118, single-point crossover HR:85-89, lattice mutation HR:91-99) and the PF
// variants (PF:192-208), plus the north-star extensions (tournament selection, uniform
// crossover, Gaussian mutation) driven by counter-based Philox4x32-10.
// oracle/tf_oracle.c ga_step_any is the CPU twin; draws are bit-identical by construction.
//
// A warp owns a TILE of 32 consecutive slots of the next population.  Phase A (lane = slot)
// decides each child's plan — parents, crossover point, mutation — with all Philox / tournament
// work done once per child in its own lane; phase B (lane = waypoint) broadcasts one plan at a
// time with shuffles and moves the waypoints.
#pragma once
#include "tf_common.cuh"

namespace tf {

// The plan of one child, held by the lane that owns the slot.
template <int NW>
struct ChildPlan {
    int first, second;   // population rows: waypoint k comes from `first` iff from_first(k)
    int ia, ib;          // both parents of the pair (parity output)
    int meta;            // cp | (mp + 1) << 9 | L << 18
    float dx, dy, dz;    // mutation delta
    uint32_t mask[NW];   // uniform-crossover mask words (bit l of word s -> waypoint 32 s + l)
};
__device__ __forceinline__ int pack_meta(int cp, int mp, int L) { return cp | ((mp + 1) << 9) | (L << 18); }
__device__ __forceinline__ int meta_cp(int m) { return m & 511; }
__device__ __forceinline__ int meta_mp(int m) { return ((m >> 9) & 511) - 1; }
__device__ __forceinline__ int meta_len(int m) { return m >> 18; }

template <typename F>
__device__ __forceinline__ bool better(F fa, int a, F fb, int b) {  // stable order: fitness desc, index asc
    return (fa > fb) || (fa == fb && a < b);
}

__device__ __forceinline__ float gauss_ih8(const U4& r) {  // Irwin-Hall(8) of 16-bit uniforms
    uint32_t s = (r.x & 0xffffu) + (r.x >> 16) + (r.y & 0xffffu) + (r.y >> 16) + (r.z & 0xffffu) + (r.z >> 16) +
                 (r.w & 0xffffu) + (r.w >> 16);
    return __fmul_rn(__fsub_rn((float)s, 262140.0f), 1.8688131e-5f);
}

// Phase A: the plan of slot `slot` (>= ne), computed entirely by the calling lane.
template <int NW, typename F>
__device__ __forceinline__ ChildPlan<NW> plan_child(int slot, int ne, const F* __restrict__ fit,
                                                   const int32_t* __restrict__ lens,
                                                   const int32_t* __restrict__ top_idx,
                                                   const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen,
                                                   const tf_ga_params& ga, int P, int W) {
    const int j = (slot - ne) >> 1, c = (slot - ne) & 1;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    ChildPlan<NW> pl;
#pragma unroll
    for (int s = 0; s < NW; ++s) pl.mask[s] = 0;
    pl.dx = pl.dy = pl.dz = 0.0f;
    int ia, ib, cp, mp = -1;
    U4 r = U4{0, 0, 0, 0};
    if (tape != nullptr) {
        const tf_draw t = tape[j];
        // a host-supplied tape is clamped, never trusted: ranks to the ranked pool, indices to the population, the
        // crossover point to [0, W] (a valid tape passes through unchanged; the oracle clamps the same way)
        if (ga.selection != TF_SEL_TOURNAMENT) {   // truncation / top-2: ranks into top_idx
            const int pool = max(1, min(P, ga.selection == TF_SEL_TOP2 ? 2 : max(ga.n_top, 1)));
            ia = top_idx[min(max(t.pa, 0), pool - 1)];
            ib = top_idx[min(max(t.pb, 0), pool - 1)];
        } else {                                   // tournament: population indices
            ia = min(max(t.pa, 0), P - 1);
            ib = min(max(t.pb, 0), P - 1);
        }
        cp = min(max(t.cp, 0), W);
        mp = c == 0 ? t.mp1 : t.mp2;
        const int d = c == 0 ? t.d1 : t.d2;
        pl.dx = (float)((d & 3) - 1);
        pl.dy = (float)(((d >> 2) & 3) - 1);
        pl.dz = (float)(((d >> 4) & 3) - 1);
    } else {
        r = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_PAIR);
        if (ga.selection == TF_SEL_TOP2) {  // PF:192-194 select_parents: the two best, in rank order
            ia = top_idx[0];
            ib = top_idx[P > 1 ? 1 : 0];
        } else if (ga.selection == TF_SEL_TRUNCATION) {  // HR:118 two distinct uniform ranks of the top pool
            const int n = min(ga.n_top, P);
            int a = (int)mulhi32(r.x, (uint32_t)n), b = (int)mulhi32(r.y, (uint32_t)(n - 1));
            if (b >= a) ++b;
            ia = top_idx[a];
            ib = top_idx[b];
        } else {  // tournament: k uniform entrants per parent, winner under the stable order
            ia = ib = -1;
            F fa = (F)0, fb = (F)0;
            // entrants e = 0, 1, ... of parent a are q.x, q.z of Philox call CALL_TOURN + (e >> 1), parent b's are
            // q.y, q.w.  Four entrants per parent at a time: all draws, then all 8 fitness gathers, then the
            // comparisons in entrant order, so a batch costs one memory round trip instead of two.
            for (int e4 = 0; e4 < ga.tournament_k; e4 += 4) {
                const int n = min(4, ga.tournament_k - e4);
                const U4 q0 = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_TOURN + (uint32_t)(e4 >> 1));
                U4 q1 = U4{0, 0, 0, 0};
                if (n > 2) q1 = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_TOURN + (uint32_t)(e4 >> 1) + 1u);
                const int ea[4] = {(int)mulhi32(q0.x, (uint32_t)P), (int)mulhi32(q0.z, (uint32_t)P),
                                   (int)mulhi32(q1.x, (uint32_t)P), (int)mulhi32(q1.z, (uint32_t)P)};
                const int eb[4] = {(int)mulhi32(q0.y, (uint32_t)P), (int)mulhi32(q0.w, (uint32_t)P),
                                   (int)mulhi32(q1.y, (uint32_t)P), (int)mulhi32(q1.w, (uint32_t)P)};
                F fea[4], feb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // indices of unused entrants are 0: a harmless gather
                    fea[i] = fit[ea[i]];
                    feb[i] = fit[eb[i]];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < n) {
                        if (ia < 0 || better(fea[i], ea[i], fa, ia)) { ia = ea[i]; fa = fea[i]; }
                        if (ib < 0 || better(feb[i], eb[i], fb, ib)) { ib = eb[i]; fb = feb[i]; }
                    }
                }
            }
        }
    }
    const int La = lens ? lens[ia] : W, Lb = lens ? lens[ib] : W;
    const int Lmin = min(La, Lb);
    int L = c == 0 ? Lb : La;  // child1 = p1[:cp] + p2[cp:] has len(p2) whenever cp <= min(len) (HR:86-88)
    if (tape != nullptr) {
        // Python slicing in general (PF:197-199 draws cp from len(parent1) alone, so it may exceed len(parent2)):
        // child = first[:cp] + second[cp:] has min(cp, len(first)) + max(0, len(second) - cp) waypoints
        const int Lf = c == 0 ? La : Lb, Ls = c == 0 ? Lb : La;
        L = min(cp, Lf) + max(0, Ls - cp);
    }
    if (tape == nullptr) {
        cp = Lmin >= 3 ? 1 + (int)mulhi32(r.z, (uint32_t)(Lmin - 2)) : Lmin;  // [1, Lmin-2] (HR:86)
        const U4 s = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_CHILD + (uint32_t)c);
        const float u = __fmul_rn((float)(s.x >> 8), 1.0f / 16777216.0f);
        if (u < ga.mutation_rate) {  // HR:94
            if (ga.mutation == TF_MUT_LATTICE_PF) mp = (int)mulhi32(s.y, (uint32_t)L);   // PF:204
            else if (L >= 3) mp = 1 + (int)mulhi32(s.y, (uint32_t)(L - 2));              // HR:95
        }
        if (ga.mutation == TF_MUT_GAUSSIAN) {
            if (mp >= 0) {
                pl.dx = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 0u)));
                pl.dy = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 1u)));
                pl.dz = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 2u)));
            }
        } else {
            const uint32_t fx = s.z & 1023u, fy = (s.z >> 10) & 1023u, fz = (s.z >> 20) & 1023u;
            if (ga.mutation == TF_MUT_LATTICE_PF) {  // PF:205-207 each axis +-1
                pl.dx = (fx & 1) ? 1.0f : -1.0f;
                pl.dy = (fy & 1) ? 1.0f : -1.0f;
                pl.dz = (fz & 1) ? 1.0f : -1.0f;
            } else {  // HR:96 each axis in {-1,0,1}
                pl.dx = (float)((int)((fx * 3u) >> 10) - 1);
                pl.dy = (float)((int)((fy * 3u) >> 10) - 1);
                pl.dz = (float)((int)((fz * 3u) >> 10) - 1);
            }
        }
        if (ga.crossover == TF_CX_UNIFORM) {
            // mask bit set -> both children take waypoint k from the "other" parent; endpoints never swap
#pragma unroll
            for (int s4 = 0; s4 < NW; s4 += 4) {
                const U4 m = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_MASK + (uint32_t)(s4 >> 2));
                pl.mask[s4] = m.x;
                if (s4 + 1 < NW) pl.mask[s4 + 1] = m.y;
                if (s4 + 2 < NW) pl.mask[s4 + 2] = m.z;
                if (s4 + 3 < NW) pl.mask[s4 + 3] = m.w;
            }
#pragma unroll
            for (int s = 0; s < NW; ++s) {  // clear the endpoint bits: k == 0 and k >= L-1
                uint32_t keep = 0xffffffffu;
                const int lo = 32 * s;
                if (lo == 0) keep &= ~1u;
                if (L - 1 <= lo) keep = 0;
                else if (L - 1 < lo + 32) keep &= (1u << (L - 1 - lo)) - 1u;
                pl.mask[s] &= keep;
            }
            cp = 511;  // every waypoint comes from `first` unless its mask bit is set
        }
    }
    pl.ia = ia;
    pl.ib = ib;
    pl.first = c == 0 ? ia : ib;
    pl.second = c == 0 ? ib : ia;
    pl.meta = pack_meta(cp, (mp >= 0 && mp < L) ? mp : -1, L);
    return pl;
}

// Elite slot: copy row `src` unchanged.
template <int NW>
__device__ __forceinline__ ChildPlan<NW> plan_elite(int src, int L) {
    ChildPlan<NW> pl;
#pragma unroll
    for (int s = 0; s < NW; ++s) pl.mask[s] = 0;
    pl.dx = pl.dy = pl.dz = 0.0f;
    pl.first = pl.second = pl.ia = pl.ib = src;
    pl.meta = pack_meta(511, -1, L);
    return pl;
}

// Phase-B plan broadcast: everything the 32 lanes need to move child c's waypoints.
template <int NW>
struct PlanB {
    int first, second, cp, mp, L;
    float dx, dy, dz;
    uint32_t mask[NW];
};
template <int NW>
__device__ __forceinline__ PlanB<NW> broadcast_plan(const ChildPlan<NW>& mine, int c, bool uniform) {
    PlanB<NW> b;
    b.first = __shfl_sync(0xffffffffu, mine.first, c);
    b.second = __shfl_sync(0xffffffffu, mine.second, c);
    const int meta = __shfl_sync(0xffffffffu, mine.meta, c);
    b.cp = meta_cp(meta);
    b.mp = meta_mp(meta);
    b.L = meta_len(meta);
    b.dx = b.dy = b.dz = 0.0f;
    if (b.mp >= 0) {  // warp-uniform
        b.dx = __shfl_sync(0xffffffffu, mine.dx, c);
        b.dy = __shfl_sync(0xffffffffu, mine.dy, c);
        b.dz = __shfl_sync(0xffffffffu, mine.dz, c);
    }
#pragma unroll
    for (int s = 0; s < NW; ++s) b.mask[s] = uniform ? __shfl_sync(0xffffffffu, mine.mask[s], c) : 0u;
    return b;
}
template <int NW>
__device__ __forceinline__ bool from_first(const PlanB<NW>& b, int k, int s, int lane) {
    return (k < b.cp) && !((b.mask[s] >> lane) & 1u);  // HR:87-88; uniform: cp = 511 and the mask decides
}

// Phase A for a tile: this lane's slot.
template <int NW, typename F>
__device__ __forceinline__ ChildPlan<NW> plan_slot(int slot, int ne, const F* fit, const int32_t* lens,
                                                  const int32_t* top_idx, const tf_draw* tape, uint64_t seed,
                                                  uint32_t gen, const tf_ga_params& ga, int P, int W,
                                                  int32_t* sel, int32_t* cpo) {
    if (slot >= P) return plan_elite<NW>(0, 0);
    if (slot < ne) {  // HR:115 elitism
        const int src = top_idx[slot];
        return plan_elite<NW>(src, lens ? lens[src] : W);
    }
    ChildPlan<NW> pl = plan_child<NW, F>(slot, ne, fit, lens, top_idx, tape, seed, gen, ga, P, W);
    if (((slot - ne) & 1) == 0) {
        const int j = (slot - ne) >> 1;
        if (sel) { sel[2 * j] = pl.ia; sel[2 * j + 1] = pl.ib; }
        if (cpo) cpo[j] = (ga.crossover == TF_CX_UNIFORM && tape == nullptr) ? 0 : meta_cp(pl.meta);
    }
    return pl;
}


}  // namespace tf
