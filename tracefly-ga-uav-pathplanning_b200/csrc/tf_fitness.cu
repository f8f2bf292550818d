// tf_fitness.cu — tf_fitness (FP32, one warp per chromosome) and tf_fitness_f64
// (reference-order float64, one thread per chromosome).
#include "tf_fitness_dev.cuh"
#include "tf_stream_dev.cuh"

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

int check_fitness_args(const tf_env* env, const void* pop, int P, int W, const tf_params* prm, const void* fit) {
    if (!env || !pop || !prm || !fit) return TF_ERR_INVALID;
    if (P < 0 || W < 1 || W > TF_W_MAX) return TF_ERR_SHAPE;
    if (prm->samples_per_segment > 1 && env->d.n_smem > kTableCap) return TF_ERR_UNSUPPORTED;
    if (prm->samples_per_segment > 64) return TF_ERR_SHAPE;
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// FP64 reference-order kernel: one thread per chromosome, the reference's own loops.
// Every operation is the IEEE double operation the reference performs, in its order
// (oracle/tf_oracle.c orc_fitness_one_f64 is the CPU twin); norms are
// sqrt(fma(dz,dz,fma(dy,dy,dx*dx))) as NumPy's dot evaluates them.
__device__ __forceinline__ double norm3d(double dx, double dy, double dz) {
    return __dsqrt_rn(__fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx))));
}

// HR:17-28 / PF:28-59 for one point, in the reference's float64 arithmetic.
__device__ __forceinline__ void point_eval_f64(const EnvDev& e, double x, double y, double z, bool& hit_out,
                                               double& best_out) {
    bool hit = false;
    double best = INFINITY;
    for (int k = 0; k < e.n; ++k) {
        const double* o = e.obs64 + 8 * k;
        const double dx = __dsub_rn(x, o[0]), dy = __dsub_rn(y, o[1]), dz = __dsub_rn(z, o[2]);
        if (e.semantics == TF_SEM_HR) {  // HR:17-28
            const double size = o[3] < 0.0 ? 1.0 : o[3];
            const double d = norm3d(dx, dy, dz);
            hit = hit || (d <= __dadd_rn(size, e.safety64));
            if (d < best) best = d;
        } else {  // PF:28-59
            const double s = o[3];
            const int kind = (int)o[7];
            double dd;
            if (kind == TF_KIND_SPHERE) {
                const double d = norm3d(dx, dy, dz);
                hit = hit || (d <= s);
                dd = __dsub_rn(d, s);
                if (!(dd > 0.0)) dd = 0.0;
            } else if (kind == TF_KIND_CUBE) {
                hit = hit || (fabs(dx) <= s && fabs(dy) <= s && fabs(dz) <= s);
                dd = __dsub_rn(norm3d(dx, dy, dz), s);
                if (!(dd > 0.0)) dd = 0.0;
            } else if (kind == TF_KIND_CYLINDER) {
                const double h = __dmul_rn(2.0, s);
                hit = hit || (fabs(dx) <= s && fabs(dy) <= s && fabs(dz) <= h);
                dd = __dsub_rn(fabs(dz), h);
                if (!(dd > 0.0)) dd = 0.0;
            } else {
                const double ax = fabs(dx), ay = fabs(dy), az = fabs(dz);
                hit = hit || (ax <= o[4] && ay <= o[5] && az <= o[6]);
                dd = norm3d(fmax(__dsub_rn(ax, o[4]), 0.0), fmax(__dsub_rn(ay, o[5]), 0.0),
                            fmax(__dsub_rn(az, o[6]), 0.0));
            }
            if (dd < best) best = dd;
        }
    }
    hit_out = hit;
    best_out = best;
}

__global__ void point_query_f64_kernel(const __grid_constant__ EnvDev e, const double* __restrict__ pts, int n,
                                       int32_t* __restrict__ collide, double* __restrict__ dnear) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool hit;
    double best;
    point_eval_f64(e, pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], hit, best);
    if (collide) collide[i] = hit ? 1 : 0;
    if (dnear) dnear[i] = best;
}

__global__ void __launch_bounds__(128)
fitness_f64_kernel(const __grid_constant__ EnvDev e, const __grid_constant__ tf_params prm,
                   const double* __restrict__ pop, const int32_t* __restrict__ lens, int P, int W,
                   double* __restrict__ fit, uint32_t* __restrict__ coll) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int L = lens ? min(lens[p], W) : W;
    const double* wp = pop + (size_t)p * W * 4;
    const int nwords = (W + 31) >> 5;
    int turns = 0, ncoll = 0;
    double alt = 0.0, prox = 0.0, dist = 0.0;
    for (int i = 1; i <= L - 2; ++i)  // HR:43
        turns += (wp[4 * i] != wp[4 * (i - 1)] || wp[4 * i + 1] != wp[4 * (i - 1) + 1]) ? 1 : 0;
    for (int i = 1; i < L; ++i) alt = __dadd_rn(alt, fabs(__dsub_rn(wp[4 * i + 2], wp[4 * (i - 1) + 2])));  // HR:44
    uint32_t word = 0;
    for (int i = 0; i < L; ++i) {  // HR:45-46
        const double x = wp[4 * i], y = wp[4 * i + 1], z = wp[4 * i + 2];
        bool hit;
        double best;
        point_eval_f64(e, x, y, z, hit, best);
        ncoll += hit ? 1 : 0;
        if (hit) word |= 1u << (i & 31);
        if (coll && ((i & 31) == 31 || i == L - 1)) {
            coll[(size_t)p * nwords + (i >> 5)] = word;
            word = 0;
        }
        prox = __dadd_rn(prox, __ddiv_rn(1.0, __dadd_rn(1.0, best)));
    }
    if (coll)
        for (int w = (L + 31) >> 5; w < nwords; ++w) coll[(size_t)p * nwords + w] = 0;
    const double pen = __dadd_rn((double)ncoll, prox);
    const double safety = __ddiv_rn(1.0, __dadd_rn(1.0, pen));
    const double energy = __ddiv_rn(1.0, __dadd_rn((double)(1 + turns), alt));
    const double wind = __ddiv_rn(1.0, __dadd_rn(1.0, alt));
    for (int i = 0; i + 1 < L; ++i)  // HR:51-52
        dist = __dadd_rn(dist, norm3d(__dsub_rn(wp[4 * i], wp[4 * (i + 1)]), __dsub_rn(wp[4 * i + 1], wp[4 * (i + 1) + 1]),
                                      __dsub_rn(wp[4 * i + 2], wp[4 * (i + 1) + 2])));
    const double time_s = (dist != 0.0) ? __ddiv_rn(1.0, dist) : (prm.dist_zero_guard ? 0.0 : (double)INFINITY);
    double f = __dmul_rn(0.4, safety);  // HR:55-59 — literal weights, left to right, no fusion
    f = __dadd_rn(f, __dmul_rn(0.2, energy));
    f = __dadd_rn(f, __dmul_rn(0.2, wind));
    f = __dadd_rn(f, __dmul_rn(0.1, time_s));
    f = __dadd_rn(f, __dmul_rn(0.1, __ddiv_rn(1.0, (double)(1 + turns))));
    fit[p] = f;
}

}  // namespace tf

using namespace tf;

extern "C" int tf_fitness(const tf_env* env, const float* pop_dev, const int32_t* lens_dev, int P, int W,
                          const tf_params* params, float* fit_dev, uint32_t* coll_dev, void* stream) {
    int rc = check_fitness_args(env, pop_dev, P, W, params, fit_dev);
    if (rc != TF_OK) return rc;
    if ((rc = ensure_device()) != TF_OK) return rc;
    if (P == 0) return TF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (lens_dev == nullptr && stream_eligible(env->d, W, false) && !g_disable_stream) {
        tf_ga_params ga_unused;
        tf_ga_params_default(&ga_unused);
        StreamArgs sa{&env->d, params, &ga_unused, pop_dev, nullptr, nullptr, nullptr, 0, 0, P, W,
                      nullptr, fit_dev, coll_dev, nullptr, nullptr, st};
        return stream_dispatch<false>(sa);
    }
    const int nw = (W + 31) / 32;
    if (nw <= 1) return dispatch_sem<1>(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    if (nw <= 2) return dispatch_sem<2>(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    if (nw <= 4) return dispatch_sem<4>(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    return dispatch_sem<8>(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
}

extern "C" int tf_fitness_f64(const tf_env* env, const double* pop64_dev, const int32_t* lens_dev, int P, int W,
                              const tf_params* params, double* fit64_dev, uint32_t* coll_dev, void* stream) {
    if (!env || !pop64_dev || !params || !fit64_dev) return TF_ERR_INVALID;
    if (P < 0 || W < 1) return TF_ERR_SHAPE;
    // the reference-order model covers exactly what the reference computes
    if (env->d.n_dyn > 0 || env->d.bricks || env->d.hazard || params->samples_per_segment > 1) return TF_ERR_UNSUPPORTED;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (P == 0) return TF_OK;
    fitness_f64_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->d, *params, pop64_dev, lens_dev, P, W,
                                                                         fit64_dev, coll_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_point_query_f64(const tf_env* env, const double* pts64_dev, int n, int32_t* collide_dev,
                                  double* dnear_dev, void* stream) {
    if (!env || !pts64_dev || n < 0) return TF_ERR_INVALID;
    if (env->d.n_dyn > 0 || env->d.bricks || env->d.hazard) return TF_ERR_UNSUPPORTED;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n == 0) return TF_OK;
    point_query_f64_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->d, pts64_dev, n, collide_dev, dnear_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
