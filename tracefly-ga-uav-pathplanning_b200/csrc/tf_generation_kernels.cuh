// tf_generation_kernels.cuh — the LSU-path fused generation kernel and its template dispatch
// (instantiated per NW in tf_inst.cu).
#pragma once
#include "tf_fitness_dev.cuh"
#include "tf_ga_dev.cuh"
#include "tf_stream_host.cuh"

namespace tf {

// ------------------------------------------------------------------------------------------
// Fused generation: child built in registers -> streamed to next_pop -> evaluated by the same
// warp.  Same results as tf_ga_step followed by tf_fitness (fixed W, single table chunk).
// Tile structure: phase A (lane = slot) plans 32 children; phase B (lane = waypoint) builds,
// stores and reduces one child at a time with the next child's parents already in flight;
// phase C (lane = slot) scores the 32 children and writes fitness / flags coalesced.
template <int NW, int SEM, bool VOX, int XF>
__global__ void __launch_bounds__(kBlockThreads)
generation_kernel(const __grid_constant__ EnvDev e, const __grid_constant__ tf_params prm,
                  const __grid_constant__ tf_ga_params ga, const float4* __restrict__ pop,
                  const float* __restrict__ fit, const int32_t* __restrict__ top_idx,
                  const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen, int P, int W, int T,
                  float4* __restrict__ next, float* __restrict__ next_fit, uint32_t* __restrict__ next_coll,
                  int32_t* __restrict__ sel, int32_t* __restrict__ cpo) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout sm = carve_smem(e, smem);
    const Tables tabs{sm.A, sm.B, sm.dA, sm.dB};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwords = (W + 31) >> 5;
    const int ne = min(ga.n_elite, P);
    const bool uniform = ga.crossover == TF_CX_UNIFORM && tape == nullptr;
    if ((SEM != TF_SEM_NONE && e.n_smem > 0) || (XF >= 1 && e.n_dyn_smem > 0)) {
        if (threadIdx.x == 0) {
            mbar_init(sm.bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            stage_chunk(e, sm.A, sm.B, 0, min(e.n_smem, kTableCap), sm.bar, XF >= 1 && e.n_dyn_smem > 0, sm.dA, sm.dB);
        mbar_wait(sm.bar, 0);
    }
    const int ntiles = (P + T - 1) / T;
    for (int tile = blockIdx.x * kWarpsPerBlock + warp; tile < ntiles; tile += gridDim.x * kWarpsPerBlock) {
        const int slot0 = tile * T;
        const int cnt = min(T, P - slot0);
        // phase A
        const ChildPlan<NW> mine = plan_slot<NW, float>(lane < T ? slot0 + lane : P, ne, fit, nullptr, top_idx, tape, seed, gen, ga, P, W, sel, cpo);
        TileSums<NW> sums;
        sums.reset();

        auto fetch = [&](int c, Lanes<NW>& dst, float (&w4)[NW]) {  // gather + mutate child c into registers
            const PlanB<NW> b = broadcast_plan<NW>(mine, c, uniform);
            const float4* ra = pop + (size_t)b.first * W;
            const float4* rb = pop + (size_t)b.second * W;
#pragma unroll
            for (int s = 0; s < NW; ++s) {
                const int k = lane + 32 * s;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < W) {
                    v = ldg_stream(from_first<NW>(b, k, s, lane) ? ra + k : rb + k);
                    if (k == b.mp) {  // HR:97-98
                        v.x = __fadd_rn(v.x, b.dx);
                        v.y = __fadd_rn(v.y, b.dy);
                        v.z = __fadd_rn(v.z, b.dz);
                    }
                }
                dst.x[s] = v.x; dst.y[s] = v.y; dst.z[s] = v.z; w4[s] = v.w;
            }
        };

        Lanes<NW> cur, nxt;
        float cw4[NW], nw4[NW];
        fetch(0, cur, cw4);
#pragma unroll 1
        for (int c = 0; c < cnt; ++c) {
            if (c + 1 < cnt) fetch(c + 1, nxt, nw4);
            float4* dst = next + (size_t)(slot0 + c) * W;
#pragma unroll
            for (int s = 0; s < NW; ++s) {
                const int k = lane + 32 * s;
                if (k < W) stg_stream(dst + k, make_float4(cur.x[s], cur.y[s], cur.z[s], cw4[s]));
            }
            PointAcc<NW, SEM> acc;
            acc.reset();
            if constexpr (SEM != TF_SEM_NONE) eval_static<SEM, NW>(e, tabs, cur.x, cur.y, cur.z, acc.hit, acc.best);
            reduce_chromosome<NW, SEM, VOX, XF>(e, prm, tabs, cur, W, acc, c, sums);
            cur = nxt;
#pragma unroll
            for (int s = 0; s < NW; ++s) cw4[s] = nw4[s];
        }
        // phase C
        if (lane < cnt) {
            const int slot = slot0 + lane;
            next_fit[slot] = combine_fitness<NW, XF>(e, prm, sums);
            if (next_coll) {
#pragma unroll
                for (int s = 0; s < NW; ++s)
                    if (s < nwords) next_coll[(size_t)slot * nwords + s] = sums.cw[s];
            }
        }
    }
}


template <int NW, int SEM, bool VOX, int XF>
static int launch_generation(const GenArgs& a) {
    auto kern = generation_kernel<NW, SEM, VOX, XF>;
    size_t smem = smem_bytes_for(*a.e);
    if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = tile_size_for(a.P);
    const int ntiles = (a.P + T - 1) / T;
    int grid = min((ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
    kern<<<grid, kBlockThreads, smem, a.st>>>(*a.e, *a.prm, *a.ga, reinterpret_cast<const float4*>(a.pop), a.fit,
                                              a.top_idx, a.tape, a.seed, a.gen, a.P, a.W, T,
                                              reinterpret_cast<float4*>(a.next), a.next_fit, a.next_coll, a.sel, a.cpo);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
template <int NW, int SEM>
static int gen_flags(const GenArgs& a) {
    const bool vox = a.e->bricks != nullptr;
    const int xf = a.prm->samples_per_segment > 1 ? 2 : ((a.e->n_dyn > 0 || a.e->hazard != nullptr) ? 1 : 0);
    if (vox) {
        if (xf == 2) return launch_generation<NW, SEM, true, 2>(a);
        if (xf == 1) return launch_generation<NW, SEM, true, 1>(a);
        return launch_generation<NW, SEM, true, 0>(a);
    }
    if (xf == 2) return launch_generation<NW, SEM, false, 2>(a);
    if (xf == 1) return launch_generation<NW, SEM, false, 1>(a);
    return launch_generation<NW, SEM, false, 0>(a);
}
template <int NW>
static int gen_sem(const GenArgs& a) {
    switch (a.e->n > 0 ? a.e->semantics : TF_SEM_NONE) {
    case TF_SEM_HR: return gen_flags<NW, TF_SEM_HR>(a);
    case TF_SEM_PF: return gen_flags<NW, TF_SEM_PF>(a);
    default: return gen_flags<NW, TF_SEM_NONE>(a);
    }
}

}  // namespace tf
