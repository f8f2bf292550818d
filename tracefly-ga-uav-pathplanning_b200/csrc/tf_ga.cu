// tf_ga.cu — tf_ga_step (f32 / f64), the fused tf_generation, best tracking, initial
// populations and the island-migration pack / replace kernels.
#include <math.h>

#include <vector>

#include "tf_fitness_dev.cuh"
#include "tf_ga_dev.cuh"
#include "tf_stream_host.cuh"

namespace tf {

// ------------------------------------------------------------------------------------------
// Unfused GA step: one warp per tile of 32 slots.  V4 = float4 / double4.  (W <= 256: NW = 8.)
template <typename V4, typename F>
__global__ void __launch_bounds__(kBlockThreads)
ga_step_kernel(const V4* __restrict__ pop, const int32_t* __restrict__ lens, const F* __restrict__ fit,
               const int32_t* __restrict__ top_idx, const tf_draw* __restrict__ tape, uint64_t seed, uint32_t gen,
               const __grid_constant__ tf_ga_params ga, int P, int W, int T, V4* __restrict__ next,
               int32_t* __restrict__ next_lens, int32_t* __restrict__ sel, int32_t* __restrict__ cpo) {
    constexpr int NW = 8;
    const int lane = threadIdx.x & 31;
    const int ne = min(ga.n_elite, P);
    const bool uniform = ga.crossover == TF_CX_UNIFORM && tape == nullptr;
    const int ntiles = (P + T - 1) / T;
    for (int tile = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); tile < ntiles; tile += gridDim.x * kWarpsPerBlock) {
        const int slot0 = tile * T;
        const int cnt = min(T, P - slot0);
        const ChildPlan<NW> mine = plan_slot<NW, F>(lane < T ? slot0 + lane : P, ne, fit, lens, top_idx, tape, seed, gen, ga, P, W, sel, cpo);
        if (next_lens && lane < cnt) next_lens[slot0 + lane] = meta_len(mine.meta);
        for (int c = 0; c < cnt; ++c) {
            const PlanB<NW> b = broadcast_plan<NW>(mine, c, uniform);
            const V4* ra = pop + (size_t)b.first * W;
            const V4* rb = pop + (size_t)b.second * W;
            V4* dst = next + (size_t)(slot0 + c) * W;
            // child = first[:cp] + second[cp:] (Python slicing): n1 waypoints of `first`, then `second` from index cp
            // on — the same index k unless cp exceeds len(first), which only a PF tape can ask for (PF:197-199)
            // (cp >= W — elite slots carry cp = 511 — copies the whole row of `first`, padding included)
            const int n1 = uniform ? 0 : (b.cp >= W ? W : min(b.cp, lens ? lens[b.first] : W));
            for (int k = lane, s = 0; k < W; k += 32, ++s) {
                V4 v;
                if (uniform) {
                    v = from_first<NW>(b, k, s, lane) ? ra[k] : rb[k];
                } else if (k < n1) {
                    v = ra[k];
                } else {
                    const int si = b.cp + (k - n1);
                    v = V4{};
                    if (si < W) v = rb[si];
                }
                if (k == b.mp) {  // HR:97-98
                    v.x = v.x + (decltype(v.x))b.dx;
                    v.y = v.y + (decltype(v.y))b.dy;
                    v.z = v.z + (decltype(v.z))b.dz;
                }
                dst[k] = v;
            }
        }
    }
}

template <typename V4, typename F, typename S>
static int run_ga_step(const S* pop, const int32_t* lens, const F* fit, const int32_t* top_idx, const tf_draw* tape,
                       uint64_t seed, uint32_t gen, const tf_ga_params* ga, int P, int W, S* next, int32_t* next_lens,
                       int32_t* sel, int32_t* cpo, cudaStream_t st) {
    if (!pop || !next || !ga || !top_idx) return TF_ERR_INVALID;
    if (ga->selection == TF_SEL_TOURNAMENT && !tape && !fit) return TF_ERR_INVALID;
    if (P < 1 || W < 1 || W > TF_W_MAX) return TF_ERR_SHAPE;
    if (ga->n_elite < 0 || ga->tournament_k < 1 || ga->tournament_k > 32) return TF_ERR_SHAPE;
    if (ga->selection < TF_SEL_TRUNCATION || ga->selection > TF_SEL_TOP2) return TF_ERR_INVALID;
    if (ga->selection == TF_SEL_TRUNCATION && (ga->n_top < 2 || P < 2)) return TF_ERR_SHAPE;
    if (lens && ga->crossover == TF_CX_UNIFORM) return TF_ERR_UNSUPPORTED;  // uniform crossover needs fixed W
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    const int T = tile_size_for(P);
    const int ntiles = (P + T - 1) / T;
    int grid = min((ntiles + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
    ga_step_kernel<V4, F><<<grid, kBlockThreads, 0, st>>>(reinterpret_cast<const V4*>(pop), lens, fit, top_idx, tape,
                                                         seed, gen, *ga, P, W, T, reinterpret_cast<V4*>(next),
                                                         next_lens, sel, cpo);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// per-NW instantiations live in tf_inst.cu
int generation_lsu_nw1(const GenArgs&);
int generation_lsu_nw2(const GenArgs&);
int generation_lsu_nw4(const GenArgs&);
int generation_lsu_nw8(const GenArgs&);
int generation_stream_nw1(const StreamArgs&);
int generation_stream_nw2(const StreamArgs&);
int generation_stream_nw4(const StreamArgs&);
int generation_stream_nw8(const StreamArgs&);
static int generation_stream(const StreamArgs& a) { return a.W <= 32 ? generation_stream_nw1(a) : (a.W <= 64 ? generation_stream_nw2(a) : (a.W <= 128 ? generation_stream_nw4(a) : generation_stream_nw8(a))); }

// ------------------------------------------------------------------------------------------
template <typename V4, typename F>
__global__ void track_best_kernel(const V4* __restrict__ pop, const F* __restrict__ fit, int slot, int W, uint32_t gen,
                                  F* best_fit, V4* best_path, int32_t* best_gen) {
    __shared__ int take;
    if (threadIdx.x == 0) take = fit[slot] > *best_fit;  // HR:126
    __syncthreads();
    if (!take) return;
    for (int k = threadIdx.x; k < W; k += blockDim.x) best_path[k] = pop[(size_t)slot * W + k];
    __syncthreads();
    if (threadIdx.x == 0) {
        *best_fit = fit[slot];
        if (best_gen) *best_gen = (int32_t)gen;
    }
}

template <typename V4>
__global__ void replicate_row_kernel(V4* pop, int P, int W) {  // rows 1..P-1 <- row 0 (HR:82-83)
    const size_t n = (size_t)P * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + W; i < n; i += (size_t)gridDim.x * blockDim.x)
        pop[i] = pop[i % W];
}

__global__ void init_walks_kernel(uint64_t seed, const float4* __restrict__ start_pp, float3 start, float3 lo, float3 hi,
                                  int P, int W, float4* __restrict__ pop) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    float cx = start.x, cy = start.y, cz = start.z;
    if (start_pp) { const float4 s = start_pp[p]; cx = s.x; cy = s.y; cz = s.z; }
    float4* row = pop + (size_t)p * W;
    for (int i = 0; i < W; ++i) {
        if (i > 0) {
            const U4 r = philox4x32_10(k0, k1, (uint32_t)p, (uint32_t)i, 0u, CALL_INIT);
            cx = fminf(fmaxf(__fadd_rn(cx, (float)((int)mulhi32(r.x, 3u) - 1)), lo.x), hi.x);
            cy = fminf(fmaxf(__fadd_rn(cy, (float)((int)mulhi32(r.y, 3u) - 1)), lo.y), hi.y);
            cz = fminf(fmaxf(__fadd_rn(cz, (float)((int)mulhi32(r.z, 3u) - 1)), lo.z), hi.z);
        }
        row[i] = make_float4(cx, cy, cz, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------
__global__ void pack_elites_kernel(const float4* __restrict__ pop, const float* __restrict__ fit,
                                   const int32_t* __restrict__ idx, int E, int W, float4* __restrict__ rows,
                                   float* __restrict__ rfit) {
    const int lane = threadIdx.x & 31;
    for (int e = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); e < E; e += gridDim.x * kWarpsPerBlock) {
        const int src = idx[e];
        for (int k = lane; k < W; k += 32) rows[(size_t)e * W + k] = pop[(size_t)src * W + k];
        if (lane == 0) rfit[e] = fit[src];
    }
}

__global__ void replace_worst_kernel(float4* __restrict__ pop, float* __restrict__ fit,
                                     const int32_t* __restrict__ worst, int W, const float4* __restrict__ grows,
                                     const float* __restrict__ gfit, int world, int self_rank, int E) {
    const int lane = threadIdx.x & 31;
    const int total = (world - 1) * E;
    for (int j = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); j < total; j += gridDim.x * kWarpsPerBlock) {
        int r = j / E;
        const int e = j - r * E;
        if (r >= self_rank) ++r;  // foreign ranks in ascending order, skipping self
        const int dst = worst[j];
        const size_t src = (size_t)r * E + e;
        for (int k = lane; k < W; k += 32) pop[(size_t)dst * W + k] = grows[src * W + k];
        if (lane == 0) fit[dst] = gfit[src];
    }
}

// HR:68-80 generate_initial_path, evaluated on the host in float64 (it is a <= W-step scalar walk).
static int diagonal_walk(const double start[3], const double end[3], std::vector<double>* out) {
    double cur[3] = {start[0], start[1], start[2]};
    int n = 1;
    if (out) out->assign({cur[0], cur[1], cur[2], 0.0});
    while (!(cur[0] == end[0] && cur[1] == end[1] && cur[2] == end[2])) {
        double nxt[3];
        bool same = true;
        for (int a = 0; a < 3; ++a) {
            double st = end[a] - cur[a];
            st = st < -1.0 ? -1.0 : (st > 1.0 ? 1.0 : st);  // np.clip(end - current, -1, 1) HR:73
            nxt[a] = cur[a] + st;
            same = same && (nxt[a] == cur[a]);
        }
        if (same) break;  // HR:75-76
        if (out) { out->push_back(nxt[0]); out->push_back(nxt[1]); out->push_back(nxt[2]); out->push_back(0.0); }
        ++n;
        cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2];
        if (n > (1 << 20)) return TF_ERR_SHAPE;
    }
    return n;
}

template <typename S, typename V4>
static int init_diagonal(const double start[3], const double end[3], int P, int W, S* pop, int* w_out, cudaStream_t st) {
    if (!start || !end) return TF_ERR_INVALID;
    std::vector<double> walk;
    int n = diagonal_walk(start, end, pop ? &walk : nullptr);
    if (n < 0) return n;
    if (w_out) *w_out = n;
    if (!pop) return TF_OK;
    if (P < 1 || W < n) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    std::vector<S> row((size_t)W * 4, (S)0);
    for (size_t i = 0; i < walk.size(); ++i) row[i] = (S)walk[i];
    TF_CUDA(cudaMemcpyAsync(pop, row.data(), row.size() * sizeof(S), cudaMemcpyHostToDevice, st));
    TF_CUDA(cudaStreamSynchronize(st));  // `row` is a stack-owned pageable buffer
    if (P > 1) {
        size_t n4 = (size_t)P * W;
        int grid = (int)min((n4 + 255) / 256, (size_t)sm_count() * 16);
        replicate_row_kernel<V4><<<grid, 256, 0, st>>>(reinterpret_cast<V4*>(pop), P, W);
        TF_LAUNCH_CHECK();
    }
    return TF_OK;
}

}  // namespace tf

using namespace tf;

extern "C" int tf_ga_step(const float* pop_dev, const int32_t* lens_dev, const float* fit_dev,
                          const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
                          const tf_ga_params* ga, int P, int W, float* next_pop_dev, int32_t* next_lens_dev,
                          int32_t* sel_dev, int32_t* cp_dev, void* stream) {
    return run_ga_step<float4, float, float>(pop_dev, lens_dev, fit_dev, top_idx_dev, tape_dev, seed, generation, ga, P,
                                             W, next_pop_dev, next_lens_dev, sel_dev, cp_dev, (cudaStream_t)stream);
}

extern "C" int tf_ga_step_f64(const double* pop64_dev, const int32_t* lens_dev, const double* fit64_dev,
                              const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed, uint32_t generation,
                              const tf_ga_params* ga, int P, int W, double* next_pop64_dev, int32_t* next_lens_dev,
                              int32_t* sel_dev, int32_t* cp_dev, void* stream) {
    return run_ga_step<double4, double, double>(pop64_dev, lens_dev, fit64_dev, top_idx_dev, tape_dev, seed, generation,
                                                ga, P, W, next_pop64_dev, next_lens_dev, sel_dev, cp_dev,
                                                (cudaStream_t)stream);
}

// tf_generation / tf_generation_ranked.  With next_top_dev the ranking of the CHILDREN (tf_rank_topk of next_fit_dev,
// next_k best) is part of the call: fused into the persistent generation kernel where that kernel runs (the last warp
// of the grid merges per-CTA candidates), a tf_rank_topk launch behind the generation otherwise.
static int run_generation(const tf_env* env, const float* pop_dev, const float* fit_dev, const int32_t* top_idx_dev,
                          const tf_draw* tape_dev, uint64_t seed, uint32_t generation, const tf_ga_params* ga,
                          const tf_params* params, int P, int W, float* next_pop_dev, float* next_fit_dev,
                          uint32_t* next_coll_dev, int32_t* sel_dev, int32_t* cp_dev, int32_t* next_top_dev, int next_k,
                          void* ws, int64_t ws_bytes, void* stream) {
    if (!env || !pop_dev || !ga || !params || !next_pop_dev || !next_fit_dev || !top_idx_dev) return TF_ERR_INVALID;
    if (ga->selection == TF_SEL_TOURNAMENT && !tape_dev && !fit_dev) return TF_ERR_INVALID;
    if (P < 1 || W < 1 || W > TF_W_MAX) return TF_ERR_SHAPE;
    if (ga->n_elite < 0 || ga->tournament_k < 1 || ga->tournament_k > 32) return TF_ERR_SHAPE;
    if (ga->selection < TF_SEL_TRUNCATION || ga->selection > TF_SEL_TOP2) return TF_ERR_INVALID;
    if (ga->selection == TF_SEL_TRUNCATION && (ga->n_top < 2 || P < 2)) return TF_ERR_SHAPE;
    if (env->d.n_smem > kTableCap) return TF_ERR_UNSUPPORTED;  // use tf_ga_step + tf_fitness (same results)
    if (params->samples_per_segment > 64) return TF_ERR_SHAPE;
    if (next_top_dev) {
        if (next_k < 1 || next_k > TF_TOPK_MAX || next_k > P) return TF_ERR_SHAPE;
        if (!ws || ws_bytes < tf_topk_workspace_bytes(P, next_k, 0)) return TF_ERR_WORKSPACE;
    }
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    bool ranked = false;
    const bool uniform = ga->crossover == TF_CX_UNIFORM && tape_dev == nullptr;
    if (stream_eligible(env->d, *params, W, uniform) && !g_disable_stream) {
        StreamArgs sa{&env->d, params, ga, pop_dev, fit_dev, top_idx_dev, tape_dev, seed, generation, P, W,
                      next_pop_dev, next_fit_dev, next_coll_dev, sel_dev, cp_dev, (cudaStream_t)stream};
        sa.uniform = uniform;
        if (next_top_dev) {
            sa.next_top = next_top_dev;
            sa.next_k = next_k;
            sa.rank_part = reinterpret_cast<unsigned long long*>(((uintptr_t)ws + 15) & ~(uintptr_t)15);
            sa.rank_part_bytes = ws_bytes - 16;
        }
        rc = generation_stream(sa);
        ranked = sa.ranked;
    } else {
        GenArgs a{&env->d, params, ga, pop_dev, fit_dev, top_idx_dev, tape_dev, seed, generation, P, W,
                  next_pop_dev, next_fit_dev, next_coll_dev, sel_dev, cp_dev, (cudaStream_t)stream};
        const int nw = (W + 31) / 32;
        rc = nw <= 1 ? generation_lsu_nw1(a) : nw <= 2 ? generation_lsu_nw2(a) : nw <= 4 ? generation_lsu_nw4(a) : generation_lsu_nw8(a);
    }
    if (rc != TF_OK || !next_top_dev || ranked) return rc;
    return tf_rank_topk(next_fit_dev, P, next_k, 0, next_top_dev, ws, ws_bytes, stream);
}

extern "C" int tf_generation(const tf_env* env, const float* pop_dev, const float* fit_dev, const int32_t* top_idx_dev,
                             const tf_draw* tape_dev, uint64_t seed, uint32_t generation, const tf_ga_params* ga,
                             const tf_params* params, int P, int W, float* next_pop_dev, float* next_fit_dev,
                             uint32_t* next_coll_dev, int32_t* sel_dev, int32_t* cp_dev, void* stream) {
    return run_generation(env, pop_dev, fit_dev, top_idx_dev, tape_dev, seed, generation, ga, params, P, W, next_pop_dev,
                          next_fit_dev, next_coll_dev, sel_dev, cp_dev, nullptr, 0, nullptr, 0, stream);
}

extern "C" int tf_generation_ranked(const tf_env* env, const float* pop_dev, const float* fit_dev,
                                    const int32_t* top_idx_dev, const tf_draw* tape_dev, uint64_t seed,
                                    uint32_t generation, const tf_ga_params* ga, const tf_params* params, int P, int W,
                                    float* next_pop_dev, float* next_fit_dev, uint32_t* next_coll_dev, int32_t* sel_dev,
                                    int32_t* cp_dev, int32_t* next_top_dev, int next_k, void* workspace_dev,
                                    int64_t workspace_bytes, void* stream) {
    if (!next_top_dev) return TF_ERR_INVALID;
    return run_generation(env, pop_dev, fit_dev, top_idx_dev, tape_dev, seed, generation, ga, params, P, W, next_pop_dev,
                          next_fit_dev, next_coll_dev, sel_dev, cp_dev, next_top_dev, next_k, workspace_dev,
                          workspace_bytes, stream);
}

extern "C" int tf_track_best(const float* pop_dev, const float* fit_dev, int slot, int W, uint32_t generation,
                             float* best_fit_dev, float* best_path_dev, int32_t* best_gen_dev, void* stream) {
    if (!pop_dev || !fit_dev || !best_fit_dev || !best_path_dev) return TF_ERR_INVALID;
    if (slot < 0 || W < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    track_best_kernel<float4, float><<<1, 64, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(pop_dev), fit_dev, slot, W, generation, best_fit_dev,
        reinterpret_cast<float4*>(best_path_dev), best_gen_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_track_best_f64(const double* pop64_dev, const double* fit64_dev, int slot, int W, uint32_t generation,
                                 double* best_fit_dev, double* best_path_dev, int32_t* best_gen_dev, void* stream) {
    if (!pop64_dev || !fit64_dev || !best_fit_dev || !best_path_dev) return TF_ERR_INVALID;
    if (slot < 0 || W < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    track_best_kernel<double4, double><<<1, 64, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double4*>(pop64_dev), fit64_dev, slot, W, generation, best_fit_dev,
        reinterpret_cast<double4*>(best_path_dev), best_gen_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_init_diagonal(const double start[3], const double end[3], int P, int W, float* pop_dev, int* w_out,
                                void* stream) {
    return init_diagonal<float, float4>(start, end, P, W, pop_dev, w_out, (cudaStream_t)stream);
}
extern "C" int tf_init_diagonal_f64(const double start[3], const double end[3], int P, int W, double* pop64_dev,
                                    int* w_out, void* stream) {
    return init_diagonal<double, double4>(start, end, P, W, pop64_dev, w_out, (cudaStream_t)stream);
}

extern "C" int tf_init_walks(uint64_t seed, const float* start_dev, const float start[3], const float lo[3],
                             const float hi[3], int P, int W, float* pop_dev, void* stream) {
    if (!pop_dev || !lo || !hi || (!start_dev && !start)) return TF_ERR_INVALID;
    if (P < 1 || W < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    float3 s = start ? make_float3(start[0], start[1], start[2]) : make_float3(0.f, 0.f, 0.f);
    init_walks_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        seed, reinterpret_cast<const float4*>(start_dev), s, make_float3(lo[0], lo[1], lo[2]),
        make_float3(hi[0], hi[1], hi[2]), P, W, reinterpret_cast<float4*>(pop_dev));
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_pack_elites(const float* pop_dev, const float* fit_dev, const int32_t* elite_idx_dev, int E, int W,
                              float* rows_dev, float* rfit_dev, void* stream) {
    if (!pop_dev || !fit_dev || !elite_idx_dev || !rows_dev || !rfit_dev) return TF_ERR_INVALID;
    if (E < 1 || W < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    int grid = min((E + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
    pack_elites_kernel<<<grid, kBlockThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(pop_dev), fit_dev, elite_idx_dev, E, W, reinterpret_cast<float4*>(rows_dev),
        rfit_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_replace_worst(float* pop_dev, float* fit_dev, const int32_t* worst_idx_dev, int W,
                                const float* gathered_rows_dev, const float* gathered_fit_dev, int world,
                                int self_rank, int E, void* stream) {
    if (!pop_dev || !fit_dev || !worst_idx_dev || !gathered_rows_dev || !gathered_fit_dev) return TF_ERR_INVALID;
    if (world < 1 || self_rank < 0 || self_rank >= world || E < 1 || W < 1) return TF_ERR_SHAPE;
    if (world == 1) return TF_OK;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    int total = (world - 1) * E;
    int grid = min((total + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 8);
    replace_worst_kernel<<<grid, kBlockThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(pop_dev), fit_dev, worst_idx_dev, W,
        reinterpret_cast<const float4*>(gathered_rows_dev), gathered_fit_dev, world, self_rank, E);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
