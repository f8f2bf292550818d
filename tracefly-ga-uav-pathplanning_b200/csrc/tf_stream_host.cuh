// tf_stream_host.cuh — argument packs and host-side eligibility of the TMA-staged stream kernels and the
// LSU-path generation kernel (no kernel templates here: those are instantiated in tf_inst.cu).
#pragma once
#include "tf_fitness_dev.cuh"

namespace tf {

// Row buffers per warp (the two-chromosome schedule uses them as kStages / 2 pair stages).  4, not more: shared
// memory comes out of the same 256 KB as L1, and the voxel probes live on L1 — with 6 buffers (4 CTAs x 50 KB) the
// SM keeps 28 KB of L1 and a spread-out population thrashes it; with 4 (4 x 33 KB) it keeps 92 KB.  cfg3 generation
// kernel, generations 5..24: 0.421 ms with 6, 0.375 ms with 4, 0.418 ms with 2 (TF_STAGES is the A/B knob).
#ifndef TF_STAGES
#define TF_STAGES 4
#endif
constexpr int kStages = TF_STAGES;

__host__ __device__ inline size_t stream_smem_bytes(const EnvDev& e, int W, int wpb = kWarpsPerBlock) {
    size_t tables = (smem_bytes_for(e) + 127) & ~(size_t)127;
    // tables | row ring | mbarriers | fused ranking: two best keys per warp + the CTA's finished-warp counter
    return tables + (size_t)wpb * kStages * W * 16 + (size_t)wpb * kStages * 8 + (size_t)wpb * 16 + 16;
}

struct StreamArgs {
    const EnvDev* e; const tf_params* prm; const tf_ga_params* ga; const float* pop; const float* fit;
    const int32_t* top_idx; const tf_draw* tape; uint64_t seed; uint32_t gen; int P, W;
    float* next; float* out_fit; uint32_t* out_coll; int32_t* sel; int32_t* cpo; cudaStream_t st;
    // fused ranking of the children (tf_generation_ranked): the kernel writes the next_k (1 or 2) best slots of
    // out_fit to next_top; rank_part = scratch for 2 keys per CTA; `ranked` is set by the launcher that did it
    int32_t* next_top = nullptr; int next_k = 0; unsigned long long* rank_part = nullptr; int64_t rank_part_bytes = 0;
    mutable bool ranked = false;
    bool uniform = false;   // uniform crossover from Philox masks: both parent rows are staged (stream_kernel<..., UX>)
};

struct GenArgs {
    const EnvDev* e; const tf_params* prm; const tf_ga_params* ga; const float* pop; const float* fit;
    const int32_t* top_idx; const tf_draw* tape; uint64_t seed; uint32_t gen; int P, W;
    float* next; float* next_fit; uint32_t* next_coll; int32_t* sel; int32_t* cpo; cudaStream_t st;
};

// Host-side eligibility: full-length chromosomes, single table chunk; parent pieces contiguous (single-point
// crossover, any staged W) or whole rows (uniform crossover, W = 32 / 64).
inline bool stream_eligible(const EnvDev& e, const tf_params& prm, int W, bool uniform_mask) {
    if (e.n_smem > kTableCap) return false;
    // uniform crossover / S > 1 samples per segment: the one-chromosome-per-step kernel, W = 32 NW exactly, not both
    if ((uniform_mask || prm.samples_per_segment > 1) && W != 32 && W != 64) return false;
    if (uniform_mask && prm.samples_per_segment > 1) return false;
    if (W != 32 && W != 64) {  // other lengths: the two-chromosome schedule with masked lanes, no table loop
        const bool table = e.n > 0 && e.semantics != TF_SEM_NONE;
        if (W < 3 || W > 256 || table || g_stream_schedule == 1) return false;
    }
    return stream_smem_bytes(e, W) <= 200 * 1024;
}

}  // namespace tf
