// tf_ga_dev.cuh — device code of selection + crossover + mutation, one warp per child.
//
// Replaces HRMAGAGenerator.evolve's inner loop HR:114-122 (elitism HR:115, truncation
// selection HR:118, single-point crossover HR:85-89, lattice mutation HR:91-99) and the PF
// variants (PF:192-208), plus the north-star extensions (tournament selection by warp
// shuffles, uniform crossover, Gaussian mutation) driven by counter-based Philox4x32-10.
// oracle/tf_oracle.c ga_step_any is the CPU twin; draws are bit-identical by construction.
#pragma once
#include "tf_common.cuh"

namespace tf {

// Everything a warp needs to know to build child `c` of pair `j` (uniform across the warp).
struct ChildPlan {
    int first, second;  // population rows: waypoint k comes from `first` iff from_first(k)
    int ia, ib;         // both parents (parity output)
    int cp;             // crossover point
    int L;              // child length
    int mp;             // mutated waypoint or -1
    float dx, dy, dz;   // mutation delta
    uint32_t lane_bits; // uniform-crossover mask bits of THIS lane (bit s -> waypoint lane+32s)
    bool uniform;
};

template <typename F>
__device__ __forceinline__ bool better(F fa, int a, F fb, int b) {  // stable order: fitness desc, index asc
    return (fa > fb) || (fa == fb && a < b);
}

// Tournament winner across the first k lanes (every lane returns the winner).
template <typename F>
__device__ __forceinline__ int warp_tournament(const F* __restrict__ fit, int entrant, bool in_play) {
    F f = in_play ? fit[entrant] : (F)0;
    int idx = in_play ? entrant : -1;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        F of = __shfl_xor_sync(0xffffffffu, f, m);
        int oi = __shfl_xor_sync(0xffffffffu, idx, m);
        bool take = oi >= 0 && (idx < 0 || better(of, oi, f, idx));
        f = take ? of : f;
        idx = take ? oi : idx;
    }
    return idx;
}

__device__ __forceinline__ float gauss_ih8(const U4& r) {  // Irwin-Hall(8) of 16-bit uniforms
    uint32_t s = (r.x & 0xffffu) + (r.x >> 16) + (r.y & 0xffffu) + (r.y >> 16) + (r.z & 0xffffu) + (r.z >> 16) +
                 (r.w & 0xffffu) + (r.w >> 16);
    return __fmul_rn(__fsub_rn((float)s, 262140.0f), 1.8688131e-5f);
}

template <typename F>
__device__ __forceinline__ ChildPlan plan_child(int j, int c, const F* __restrict__ fit,
                                                const int32_t* __restrict__ lens, const int32_t* __restrict__ top_idx,
                                                const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen,
                                                const tf_ga_params& ga, int P, int W) {
    const int lane = threadIdx.x & 31;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    ChildPlan pl;
    pl.uniform = false;
    pl.lane_bits = 0;
    pl.dx = pl.dy = pl.dz = 0.0f;
    int ia, ib, mp = -1;
    if (tape != nullptr) {
        const tf_draw t = tape[j];
        ia = ga.selection == TF_SEL_TRUNCATION ? top_idx[t.pa] : t.pa;
        ib = ga.selection == TF_SEL_TRUNCATION ? top_idx[t.pb] : t.pb;
        pl.cp = t.cp;
        mp = c == 0 ? t.mp1 : t.mp2;
        const int d = c == 0 ? t.d1 : t.d2;
        pl.dx = (float)((d & 3) - 1);
        pl.dy = (float)(((d >> 2) & 3) - 1);
        pl.dz = (float)(((d >> 4) & 3) - 1);
    } else {
        const U4 r = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_PAIR);
        U4 q = U4{0, 0, 0, 0};
        if (ga.selection == TF_SEL_TOURNAMENT || ga.crossover == TF_CX_UNIFORM)
            q = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_LANE + (uint32_t)lane);
        if (ga.selection == TF_SEL_TRUNCATION) {  // HR:118 two distinct uniform ranks of the top pool
            const int n = min(ga.n_top, P);
            int a = (int)mulhi32(r.x, (uint32_t)n), b = (int)mulhi32(r.y, (uint32_t)(n - 1));
            if (b >= a) ++b;
            ia = top_idx[a];
            ib = top_idx[b];
        } else {  // tournament: lane l < k enters one candidate per parent, winner by warp shuffles
            const bool in_play = lane < ga.tournament_k;
            ia = warp_tournament<F>(fit, (int)mulhi32(q.x, (uint32_t)P), in_play);
            ib = warp_tournament<F>(fit, (int)mulhi32(q.y, (uint32_t)P), in_play);
        }
        pl.lane_bits = q.z;
        pl.uniform = ga.crossover == TF_CX_UNIFORM;
    }
    const int La = lens ? lens[ia] : W, Lb = lens ? lens[ib] : W;
    const int Lmin = min(La, Lb);
    pl.L = c == 0 ? Lb : La;  // child1 = p1[:cp] + p2[cp:] has len(p2) (HR:87-88)
    if (tape == nullptr) {
        const U4 r = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_PAIR);
        pl.cp = Lmin >= 3 ? 1 + (int)mulhi32(r.z, (uint32_t)(Lmin - 2)) : Lmin;  // [1, Lmin-2] (HR:86)
        const U4 s = philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_CHILD + (uint32_t)c);
        const float u = __fmul_rn((float)(s.x >> 8), 1.0f / 16777216.0f);
        if (u < ga.mutation_rate) {  // HR:94
            if (ga.mutation == TF_MUT_LATTICE_PF) mp = (int)mulhi32(s.y, (uint32_t)pl.L);       // PF:204
            else if (pl.L >= 3) mp = 1 + (int)mulhi32(s.y, (uint32_t)(pl.L - 2));               // HR:95
        }
        if (ga.mutation == TF_MUT_GAUSSIAN) {
            pl.dx = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 0u)));
            pl.dy = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 1u)));
            pl.dz = __fmul_rn(ga.sigma, gauss_ih8(philox4x32_10(k0, k1, (uint32_t)j, gen, ga.stream_id, CALL_GAUSS + 3u * c + 2u)));
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
    }
    pl.ia = ia;
    pl.ib = ib;
    pl.first = c == 0 ? ia : ib;
    pl.second = c == 0 ? ib : ia;
    pl.mp = (mp >= 0 && mp < pl.L) ? mp : -1;
    return pl;
}

// Which parent supplies waypoint k (slot s of this lane)?
__device__ __forceinline__ bool from_first(const ChildPlan& pl, int k, int s) {
    if (pl.uniform) return (k == 0 || k >= pl.L - 1) ? true : !((pl.lane_bits >> s) & 1u);
    return k < pl.cp;  // HR:87-88
}

}  // namespace tf
