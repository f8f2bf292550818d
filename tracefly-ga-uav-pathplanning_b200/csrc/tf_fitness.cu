// tf_fitness.cu — tf_fitness (FP32, one warp per chromosome) and tf_fitness_f64
// (reference-order float64, one thread per chromosome).
#include "tf_fitness_dev.cuh"
#include "tf_stream_host.cuh"

namespace tf {

// per-NW instantiations live in tf_inst.cu (compiled once per NW)
int fitness_lsu_nw1(const EnvDev&, const tf_params&, const float*, const int32_t*, int, int, float*, uint32_t*, cudaStream_t);
int fitness_lsu_nw2(const EnvDev&, const tf_params&, const float*, const int32_t*, int, int, float*, uint32_t*, cudaStream_t);
int fitness_lsu_nw4(const EnvDev&, const tf_params&, const float*, const int32_t*, int, int, float*, uint32_t*, cudaStream_t);
int fitness_lsu_nw8(const EnvDev&, const tf_params&, const float*, const int32_t*, int, int, float*, uint32_t*, cudaStream_t);
int fitness_stream_nw1(const StreamArgs&);
int fitness_stream_nw2(const StreamArgs&);
int fitness_stream_nw4(const StreamArgs&);
int fitness_stream_nw8(const StreamArgs&);
static int fitness_stream(const StreamArgs& a) { return a.W <= 32 ? fitness_stream_nw1(a) : (a.W <= 64 ? fitness_stream_nw2(a) : (a.W <= 128 ? fitness_stream_nw4(a) : fitness_stream_nw8(a))); }

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

// PF:93-132 `PathPlanner.find_path`, one thread per path (the walk is sequential, the paths are independent).
__device__ __forceinline__ double sign_d(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }
__global__ void find_path_pf_kernel(const __grid_constant__ EnvDev e, double sx, double sy, double sz, double tx,
                                    double ty, double tz, const int32_t* __restrict__ hv, int n, double mx, double my,
                                    double mz, int cap, double* __restrict__ paths, int32_t* __restrict__ lens,
                                    int32_t* __restrict__ overflow) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double* out = paths + (size_t)p * cap * 4;
    const double bias = (double)hv[p];
    double cx = sx, cy = sy, cz = sz;
    int len = 1;
    out[0] = cx; out[1] = cy; out[2] = cz; out[3] = 0.0;
    while (!(cx == tx && cy == ty && cz == tz)) {  // PF:97
        const double nx = __dadd_rn(cx, sign_d(__dsub_rn(tx, cx))), ny = __dadd_rn(cy, sign_d(__dsub_rn(ty, cy))),
                     nz = __dadd_rn(cz, sign_d(__dadd_rn(__dsub_rn(tz, cz), bias)));  // PF:111-116
        bool hit;
        double dn;
        point_eval_f64(e, nx, ny, nz, hit, dn);
        if (hit) {  // PF:100-101, 118-129: one dodge move, or stay put
            const double qx = __dadd_rn(cx, mx), qy = __dadd_rn(cy, my), qz = __dadd_rn(cz, mz);
            point_eval_f64(e, qx, qy, qz, hit, dn);
            if (!hit) { cx = qx; cy = qy; cz = qz; }
        } else {
            cx = nx; cy = ny; cz = nz;
        }
        bool seen = false;  // PF:105-106 loop guard
        for (int i = len - 1; i >= 0 && !seen; --i) seen = out[4 * i] == cx && out[4 * i + 1] == cy && out[4 * i + 2] == cz;
        if (seen) break;
        if (len == cap) {
            if (overflow) atomicAdd(overflow, 1);
            break;
        }
        out[4 * len] = cx; out[4 * len + 1] = cy; out[4 * len + 2] = cz; out[4 * len + 3] = 0.0;
        ++len;
    }
    lens[p] = len;
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
    if (lens_dev == nullptr && stream_eligible(env->d, *params, W, false) && !g_disable_stream) {
        tf_ga_params ga_unused;
        tf_ga_params_default(&ga_unused);
        StreamArgs sa{&env->d, params, &ga_unused, pop_dev, nullptr, nullptr, nullptr, 0, 0, P, W,
                      nullptr, fit_dev, coll_dev, nullptr, nullptr, st};
        return fitness_stream(sa);
    }
    const int nw = (W + 31) / 32;
    if (nw <= 1) return fitness_lsu_nw1(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    if (nw <= 2) return fitness_lsu_nw2(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    if (nw <= 4) return fitness_lsu_nw4(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
    return fitness_lsu_nw8(env->d, *params, pop_dev, lens_dev, P, W, fit_dev, coll_dev, st);
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

extern "C" int tf_find_path_pf(const tf_env* env, const double source[3], const double destination[3],
                               const int32_t* height_variation_dev, int n, int dodge, int cap, double* paths64_dev,
                               int32_t* lens_dev, int32_t* overflow_dev, void* stream) {
    static const double kMoves[6][3] = {{0, 0, 1}, {0, 0, -1}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}};  // PF:119-126
    if (!env || !source || !destination || !height_variation_dev || !paths64_dev || !lens_dev) return TF_ERR_INVALID;
    if (n < 0 || cap < 1 || dodge < 0 || dodge > 5) return TF_ERR_INVALID;
    if (env->d.n_dyn > 0 || env->d.bricks || env->d.hazard) return TF_ERR_UNSUPPORTED;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n == 0) return TF_OK;
    find_path_pf_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        env->d, source[0], source[1], source[2], destination[0], destination[1], destination[2], height_variation_dev, n,
        kMoves[dodge][0], kMoves[dodge][1], kMoves[dodge][2], cap, paths64_dev, lens_dev, overflow_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
