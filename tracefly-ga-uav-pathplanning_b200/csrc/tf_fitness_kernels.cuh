// tf_fitness_kernels.cuh — the LSU-path FP32 fitness kernel and its template dispatch.  Instantiated per
// NW (waypoints per lane) in separate translation units (tf_inst.cu) to keep build times parallel.
#pragma once
#include "tf_fitness_dev.cuh"

namespace tf {

// ------------------------------------------------------------------------------------------
// FP32 throughput kernel.  Persistent grid; every CTA stages the obstacle table into shared
// memory with one TMA bulk copy; each of its 8 warps then takes TILES of 32 chromosomes:
// phase B streams one chromosome at a time (coalesced float4 loads, L1 no-allocate, next
// chromosome prefetched) and reduces it across the warp, phase C scores 32 chromosomes at once.
template <int NW, int SEM, bool VOX, int XF>
__global__ void __launch_bounds__(kBlockThreads)
fitness_kernel(const __grid_constant__ EnvDev e, const __grid_constant__ tf_params prm,
               const float4* __restrict__ pop, const int32_t* __restrict__ lens, int P, int W, int T,
               float* __restrict__ fit, uint32_t* __restrict__ coll) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout sm = carve_smem(e, smem);
    const Tables tabs{sm.A, sm.B, sm.dA, sm.dB};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwords = (W + 31) >> 5;
    const int nchunks = (SEM == TF_SEM_NONE) ? 0 : (e.n_smem + kTableCap - 1) / kTableCap;
    const bool need_stage = (SEM != TF_SEM_NONE && e.n_smem > 0) || (XF >= 1 && e.n_dyn_smem > 0);
    uint32_t phase = 0;
    bool first_chunk_resident = false;

    if (need_stage) {
        if (threadIdx.x == 0) {
            mbar_init(sm.bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            stage_chunk(e, sm.A, sm.B, 0, min(e.n_smem, kTableCap), sm.bar, XF >= 1 && e.n_dyn_smem > 0, sm.dA, sm.dB);
        mbar_wait(sm.bar, phase);
        phase ^= 1;
        first_chunk_resident = true;
    }

    const int ntiles = (P + T - 1) / T;  // T = chromosomes per tile (host picks 32 unless P is small)
    // CTA-uniform trip count (the multi-chunk path synchronises the whole CTA per chunk)
    for (int tbase = blockIdx.x * kWarpsPerBlock; tbase < ntiles; tbase += gridDim.x * kWarpsPerBlock) {
        const int tile = tbase + warp;
        const int p0 = tile * T;
        const int cnt = max(0, min(T, P - p0));  // chromosomes of this warp's tile (0: idle warp)
        const int myL = (lane < cnt) ? (lens ? min(lens[p0 + lane], W) : W) : 0;
        TileSums<NW> sums;
        sums.reset();

        auto load = [&](int c, Lanes<NW>& dst) {
            const int L = __shfl_sync(0xffffffffu, myL, c);
            const float4* row = pop + (size_t)(p0 + (c < cnt ? c : 0)) * W;
#pragma unroll
            for (int s = 0; s < NW; ++s) {
                const int k = lane + 32 * s;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < L) v = ldg_stream(row + k);
                dst.x[s] = v.x; dst.y[s] = v.y; dst.z[s] = v.z;
            }
            return L;
        };

        Lanes<NW> cur, nxt;
        int Lc = load(0, cur), Ln = 0;
        const int climit = (nchunks > 1) ? T : cnt;  // multi-chunk: every warp walks all T (syncs inside)
#pragma unroll 1
        for (int c = 0; c < climit; ++c) {
            if (c + 1 < climit) Ln = load(c + 1, nxt);
            PointAcc<NW, SEM> acc;
            acc.reset();
            if constexpr (SEM != TF_SEM_NONE) {
                if (nchunks <= 1) {
                    eval_static<SEM, NW>(e, tabs, cur.x, cur.y, cur.z, acc.hit, acc.best);
                } else {
                    // table larger than one chunk: the CTA walks the chunks in lock-step
                    for (int ch = 0; ch < nchunks; ++ch) {
                        const int row0 = ch * kTableCap, n = min(kTableCap, e.n - row0);
                        if (!(ch == 0 && first_chunk_resident)) {
                            __syncthreads();  // everyone is done reading the previous chunk
                            if (threadIdx.x == 0) stage_chunk(e, sm.A, sm.B, row0, n, sm.bar, false, sm.dA, sm.dB);
                            mbar_wait(sm.bar, phase);
                            phase ^= 1;
                        }
                        first_chunk_resident = false;
                        eval_static_chunk<SEM, NW>(e, tabs, row0, n, cur.x, cur.y, cur.z, acc.hit, acc.best);
                    }
                }
            }
            reduce_chromosome<NW, SEM, VOX, XF>(e, prm, tabs, cur, Lc, acc, c, sums);
            cur = nxt;
            Lc = Ln;
        }
        // phase C: lane = chromosome
        if (lane < cnt) {
            const int p = p0 + lane;
            fit[p] = combine_fitness<NW, XF>(e, prm, sums);
            if (coll) {
#pragma unroll
                for (int s = 0; s < NW; ++s)
                    if (s < nwords) coll[(size_t)p * nwords + s] = sums.cw[s];
            }
        }
    }
}

template <int NW, int SEM, bool VOX, int XF>
static int launch_fitness(const EnvDev& e, const tf_params& prm, const float* pop, const int32_t* lens, int P, int W,
                          float* fit, uint32_t* coll, cudaStream_t st) {
    auto kern = fitness_kernel<NW, SEM, VOX, XF>;
    size_t smem = smem_bytes_for(e);
    if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = tile_size_for(P);
    const int ntiles = (P + T - 1) / T;
    int grid = min((ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
    kern<<<grid, kBlockThreads, smem, st>>>(e, prm, reinterpret_cast<const float4*>(pop), lens, P, W, T, fit, coll);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

template <int NW, int SEM>
static int dispatch_flags(const EnvDev& e, const tf_params& prm, const float* pop, const int32_t* lens, int P, int W,
                          float* fit, uint32_t* coll, cudaStream_t st) {
    const bool vox = e.bricks != nullptr;
    const int xf = prm.samples_per_segment > 1 ? 2 : ((e.n_dyn > 0 || e.hazard != nullptr) ? 1 : 0);
    if (vox) {
        if (xf == 2) return launch_fitness<NW, SEM, true, 2>(e, prm, pop, lens, P, W, fit, coll, st);
        if (xf == 1) return launch_fitness<NW, SEM, true, 1>(e, prm, pop, lens, P, W, fit, coll, st);
        return launch_fitness<NW, SEM, true, 0>(e, prm, pop, lens, P, W, fit, coll, st);
    }
    if (xf == 2) return launch_fitness<NW, SEM, false, 2>(e, prm, pop, lens, P, W, fit, coll, st);
    if (xf == 1) return launch_fitness<NW, SEM, false, 1>(e, prm, pop, lens, P, W, fit, coll, st);
    return launch_fitness<NW, SEM, false, 0>(e, prm, pop, lens, P, W, fit, coll, st);
}

template <int NW>
static int dispatch_sem(const EnvDev& e, const tf_params& prm, const float* pop, const int32_t* lens, int P, int W,
                        float* fit, uint32_t* coll, cudaStream_t st) {
    switch (e.n > 0 ? e.semantics : TF_SEM_NONE) {
    case TF_SEM_HR: return dispatch_flags<NW, TF_SEM_HR>(e, prm, pop, lens, P, W, fit, coll, st);
    case TF_SEM_PF: return dispatch_flags<NW, TF_SEM_PF>(e, prm, pop, lens, P, W, fit, coll, st);
    default: return dispatch_flags<NW, TF_SEM_NONE>(e, prm, pop, lens, P, W, fit, coll, st);
    }
}

}  // namespace tf
