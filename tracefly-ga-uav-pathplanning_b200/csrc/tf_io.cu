// tf_io.cu — compact host formats either side of the path ("data formats").
// Lattice chromosomes (what HR:68-99 produces: integer coordinates) fit int16: 8 B per waypoint
// (x, y, z, pad) instead of 16 B, which halves the PCIe traffic of a host round trip.
#include "tf_common.cuh"

namespace tf {

__global__ void unpack_i16_kernel(const short4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const short4 v = in[i];
        out[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
    }
}

// round to nearest even, saturate to [-32768, 32767]; exact on lattice coordinates
__global__ void pack_i16_kernel(const float4* __restrict__ in, short4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        auto cv = [](float f) { return (short)max(-32768, min(32767, __float2int_rn(f))); };
        out[i] = make_short4(cv(v.x), cv(v.y), cv(v.z), cv(v.w));
    }
}

// (x, y, z) triples <-> the device's float4 waypoints: what a caller holding the reference's 3-tuples as a
// float32[P, W, 3] array sends; 12 B per waypoint over PCIe instead of 16
__global__ void unpack_xyz_kernel(const float* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.0f);
}
__global__ void pack_xyz_kernel(const float4* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        out[3 * i] = v.x;
        out[3 * i + 1] = v.y;
        out[3 * i + 2] = v.z;
    }
}

}  // namespace tf

using namespace tf;

extern "C" int tf_unpack_xyz(const float* in_dev, float* pop_dev, int64_t n_waypoints, void* stream) {
    if (!in_dev || !pop_dev || n_waypoints < 0) return TF_ERR_INVALID;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n_waypoints == 0) return TF_OK;
    const int64_t want = (n_waypoints + 255) / 256, cap = (int64_t)sm_count() * 16;
    unpack_xyz_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        in_dev, reinterpret_cast<float4*>(pop_dev), (size_t)n_waypoints);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_pack_xyz(const float* pop_dev, float* out_dev, int64_t n_waypoints, void* stream) {
    if (!pop_dev || !out_dev || n_waypoints < 0) return TF_ERR_INVALID;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n_waypoints == 0) return TF_OK;
    const int64_t want = (n_waypoints + 255) / 256, cap = (int64_t)sm_count() * 16;
    pack_xyz_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(pop_dev), out_dev, (size_t)n_waypoints);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_unpack_i16(const int16_t* in_dev, float* pop_dev, int64_t n_waypoints, void* stream) {
    if (!in_dev || !pop_dev || n_waypoints < 0) return TF_ERR_INVALID;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n_waypoints == 0) return TF_OK;
    const int64_t want = (n_waypoints + 255) / 256, cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    unpack_i16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const short4*>(in_dev),
                                                              reinterpret_cast<float4*>(pop_dev), (size_t)n_waypoints);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_pack_i16(const float* pop_dev, int16_t* out_dev, int64_t n_waypoints, void* stream) {
    if (!pop_dev || !out_dev || n_waypoints < 0) return TF_ERR_INVALID;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n_waypoints == 0) return TF_OK;
    const int64_t want = (n_waypoints + 255) / 256, cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    pack_i16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(pop_dev),
                                                            reinterpret_cast<short4*>(out_dev), (size_t)n_waypoints);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// Backend GA (uav-backend/app/routers/GA_enhanced.py) integer fitness, GAE:20-26:
//   fitness = -len(path) - #{ i in [1, len-2] : path[i-1] != path[i+1] }      (tuple inequality)
// One warp per path; lanes stride over i.
namespace tf {
__global__ void __launch_bounds__(256)
fitness_gae_kernel(const float4* __restrict__ pop, const int32_t* __restrict__ lens, int P, int W, int32_t* __restrict__ fit) {
    const int lane = threadIdx.x & 31;
    for (int p = blockIdx.x * 8 + (threadIdx.x >> 5); p < P; p += gridDim.x * 8) {
        const int L = lens ? min(lens[p], W) : W;
        const float4* row = pop + (size_t)p * W;
        int n = 0;
        for (int i = 1 + lane; i <= L - 2; i += 32) {
            const float4 a = row[i - 1], b = row[i + 1];
            n += (a.x != b.x || a.y != b.y || a.z != b.z) ? 1 : 0;
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) fit[p] = -L - n;
    }
}
}  // namespace tf

extern "C" int tf_fitness_gae(const float* pop_dev, const int32_t* lens_dev, int P, int W, int32_t* fit_dev, void* stream) {
    if (!pop_dev || !fit_dev) return TF_ERR_INVALID;
    if (P < 0 || W < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (P == 0) return TF_OK;
    const int grid = (P + 7) / 8 < sm_count() * 8 ? (P + 7) / 8 : sm_count() * 8;
    tf::fitness_gae_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(pop_dev), lens_dev, P, W,
                                                                  fit_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// ------------------------------------------------------------------------------------------
// LiDAR point cloud -> bit-packed occupancy: every point marks the voxel
// int((p - origin) / voxel_size) — float64 division, truncation toward zero: the voxel assignment of
// `downsample_lidar` (UAV_Navigation/data_compression/lidar_compression.py:242-246, origin = 0 there), kept as
// occupancy bits in the brick layout the fitness kernels probe; its np.unique is the set of bits.  Voxels outside
// the grid are dropped.
namespace tf {
__global__ void voxelize_points_kernel(const float4* __restrict__ pts, int64_t n, double ox, double oy, double oz,
                                       double voxel, uint32_t* __restrict__ bricks, int nx, int ny, int nz) {
    const int nbx = nx >> 3, nby = ny >> 3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const double qx = __ddiv_rn(__dsub_rn((double)p.x, ox), voxel), qy = __ddiv_rn(__dsub_rn((double)p.y, oy), voxel),
                     qz = __ddiv_rn(__dsub_rn((double)p.z, oz), voxel);
        if (!(fabs(qx) < 2147483648.0 && fabs(qy) < 2147483648.0 && fabs(qz) < 2147483648.0)) continue;
        const long long x = (long long)qx, y = (long long)qy, z = (long long)qz;   // astype(int): toward zero
        if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz) continue;
        const size_t brick = (size_t)(x >> 3) + (size_t)nbx * ((size_t)(y >> 3) + (size_t)nby * (size_t)(z >> 3));
        atomicOr(bricks + brick * 16 + ((z & 7) * 2 + ((y & 7) >> 2)), 1u << (((y & 3) << 3) | (x & 7)));
    }
}

// GAE:28-34 roulette: cumulative weights exactly as itertools.accumulate builds them (one thread walks the
// population left to right in float64; the integer total is reduced by the whole block first), then one bisect per
// uniform draw.
__global__ void __launch_bounds__(256) roulette_cumweights_kernel(const int32_t* __restrict__ score, int P, double* __restrict__ cum) {
    __shared__ long long part[256];
    long long t = 0;
    for (int i = threadIdx.x; i < P; i += 256) t += score[i];
    part[threadIdx.x] = t;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) part[threadIdx.x] += part[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double total = (double)part[0];
    double run = 0.0;
    for (int i = 0; i < P; ++i) {
        const double w = __ddiv_rn((double)score[i], total);
        run = i == 0 ? w : __dadd_rn(run, w);
        cum[i] = run;
    }
}
__global__ void select_roulette_kernel(const double* __restrict__ cum, int P, const double* __restrict__ u, int n,
                                       int32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double total = __dadd_rn(cum[P - 1], 0.0);
    const double x = __dmul_rn(u[i], total);
    int lo = 0, hi = P - 1;
    while (lo < hi) {  // bisect_right(cum, x, 0, P - 1)
        const int mid = (lo + hi) >> 1;
        if (x < cum[mid]) hi = mid; else lo = mid + 1;
    }
    idx[i] = lo;
}
}  // namespace tf

extern "C" int tf_voxelize_points(const float* points_dev, int64_t n, const double origin[3], double voxel_size,
                                  uint32_t* bricks_dev, int nx, int ny, int nz, void* stream) {
    if (!bricks_dev || !origin || (n > 0 && !points_dev) || n < 0) return TF_ERR_INVALID;
    if (!(voxel_size > 0.0) || tf_voxel_grid_bytes(nx, ny, nz) < 0) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n == 0) return TF_OK;
    const int64_t want = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
    tf::voxelize_points_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(points_dev), n, origin[0], origin[1], origin[2], voxel_size, bricks_dev, nx, ny, nz);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_roulette_cumweights(const int32_t* score_dev, int P, double* cum_dev, void* stream) {
    if (!score_dev || !cum_dev) return TF_ERR_INVALID;
    if (P < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    tf::roulette_cumweights_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(score_dev, P, cum_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_select_roulette(const double* cum_dev, int P, const double* u_dev, int n, int32_t* idx_dev, void* stream) {
    if (!cum_dev || (n > 0 && (!u_dev || !idx_dev))) return TF_ERR_INVALID;
    if (P < 1 || n < 0) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n == 0) return TF_OK;
    tf::select_roulette_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cum_dev, P, u_dev, n, idx_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}


// ------------------------------------------------------------------------------------------
// Result egress: the lossy front ends of ga_waypoints_compression.py on the device, one thread per
// coordinate, float64 arithmetic with explicit roundings (np.round = rint, half to even; decimals d: rint(x 10^d) / 10^d).
namespace tf {
template <typename T>
__global__ void encode_waypoints_kernel(const T* __restrict__ paths, int64_t n, int W, int mode, double param, double scale,
                                        void* __restrict__ out) {
    const int64_t total = n * W * 3;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(o % 3);
        const int64_t wp = o / 3;
        const int i = (int)(wp % W);
        const double w = (double)paths[wp * 4 + a];
        if (mode == TF_CODEC_QUANT_F32) {
            reinterpret_cast<float*>(out)[o] = __double2float_rn(__dmul_rn(rint(__ddiv_rn(w, param)), param));
        } else {
            const double prev = i == 0 ? w : (double)paths[(wp - 1) * 4 + a];
            reinterpret_cast<double*>(out)[o] = __ddiv_rn(rint(__dmul_rn(__dsub_rn(w, prev), scale)), scale);
        }
    }
}
}  // namespace tf

extern "C" int tf_encode_waypoints(const void* paths_dev, int f32, int64_t n_paths, int W, int mode, double param,
                                   void* out_dev, void* stream) {
    if (!paths_dev || !out_dev) return TF_ERR_INVALID;
    if (n_paths < 0 || W < 1) return TF_ERR_SHAPE;
    if (mode != TF_CODEC_QUANT_F32 && mode != TF_CODEC_DELTA_ROUND) return TF_ERR_INVALID;
    if (mode == TF_CODEC_QUANT_F32 && !(param > 0.0)) return TF_ERR_SHAPE;
    if (mode == TF_CODEC_DELTA_ROUND && (param < 0.0 || param > 15.0 || param != (double)(int)param)) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n_paths == 0) return TF_OK;
    double scale = 1.0;
    if (mode == TF_CODEC_DELTA_ROUND) for (int i = 0; i < (int)param; ++i) scale *= 10.0;
    const int64_t total = n_paths * W * 3, want = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    if (f32)
        tf::encode_waypoints_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)paths_dev, n_paths, W, mode, param, scale, out_dev);
    else
        tf::encode_waypoints_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)paths_dev, n_paths, W, mode, param, scale, out_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}


// ------------------------------------------------------------------------------------------
// Backend GA initial population: UAVPathPlanner.find_path (GAE:166-179) for n paths in sequence, one
// warp; the 32 lanes share the voxel-set membership scan (GAE:140-141: `point in obstacles`).  float64 straight lines
// (GAE:143-151, np.linspace + round-half-even), shuffles from the host tape (GAE:157).
namespace tf {
__device__ __forceinline__ bool gae_hit(const int4* __restrict__ obs, int n_obs, int x, int y, int z) {
    bool hit = false;
    for (int k = threadIdx.x; k < n_obs && !hit; k += 32) {
        const int4 o = __ldg(obs + k);
        hit = o.x == x && o.y == y && o.z == z;
    }
    return __any_sync(0xffffffffu, hit);
}
__global__ void __launch_bounds__(32)
find_path_gae_kernel(const int4* __restrict__ obs, int n_obs, int xlo, int xhi, int ylo, int yhi, int zlo, int zhi, int sx,
                     int sy, int sz, int dx, int dy, int dz, const int32_t* __restrict__ shuffles, int n_shuffles, int n_paths,
                     int cap, int4* __restrict__ paths, int32_t* __restrict__ lens, int32_t* __restrict__ result) {
    const int mvx[6] = {1, -1, 0, 0, 0, 0}, mvy[6] = {0, 0, 1, -1, 0, 0}, mvz[6] = {0, 0, 0, 0, 1, -1};
    int used = 0, status = 0;
    for (int p = 0; p < n_paths && status == 0; ++p) {
        int4* out = paths + (size_t)p * cap;
        int cx = sx, cy = sy, cz = sz, len = 1;
        if (threadIdx.x == 0) out[0] = make_int4(cx, cy, cz, 0);
        while (!(cx == dx && cy == dy && cz == dz) && status == 0) {
            const int ax = cx, ay = cy, az = cz;                       // the line starts where the planner stands
            const int steps = max(max(abs(dx - ax), abs(dy - ay)), abs(dz - az));
            const double inv = __ddiv_rn(1.0, (double)steps);
            for (int i = 0; i <= steps; ++i) {
                const double t = i == steps ? 1.0 : __dmul_rn((double)i, inv);
                const int qx = (int)rint(__dadd_rn((double)ax, __dmul_rn(t, (double)(dx - ax))));
                const int qy = (int)rint(__dadd_rn((double)ay, __dmul_rn(t, (double)(dy - ay))));
                const int qz = (int)rint(__dadd_rn((double)az, __dmul_rn(t, (double)(dz - az))));
                if (gae_hit(obs, n_obs, qx, qy, qz)) {                   // GAE:170-172
                    if (used >= n_shuffles) { status = 2; break; }
                    const int order = __ldg(shuffles + used++);
                    for (int m = 0; m < 6; ++m) {                        // GAE:157-163
                        const int k = (order >> (3 * m)) & 7;
                        const int nx = cx + mvx[k], ny = cy + mvy[k], nz = cz + mvz[k];
                        if (nx >= xlo && nx <= xhi && ny >= ylo && ny <= yhi && nz >= zlo && nz <= zhi &&
                            !gae_hit(obs, n_obs, nx, ny, nz)) {
                            cx = nx; cy = ny; cz = nz;
                            break;
                        }
                    }
                    break;
                }
                if (len == cap) { status = 1; break; }
                if (threadIdx.x == 0) out[len] = make_int4(qx, qy, qz, 0);
                ++len;
                cx = qx; cy = qy; cz = qz;
                if (cx == dx && cy == dy && cz == dz) break;
            }
        }
        if (threadIdx.x == 0) lens[p] = len;
    }
    if (threadIdx.x == 0) { result[0] = used; result[1] = status; }
}
}  // namespace tf

extern "C" int tf_find_path_gae(const int32_t* obstacles_dev, int n_obs, const int32_t bounds[6], const int32_t source[3],
                                const int32_t destination[3], const int32_t* shuffles_dev, int n_shuffles, int n_paths,
                                int cap, int32_t* paths_dev, int32_t* lens_dev, int32_t* result_dev, void* stream) {
    if (!bounds || !source || !destination || !paths_dev || !lens_dev || !result_dev) return TF_ERR_INVALID;
    if ((n_obs > 0 && !obstacles_dev) || (n_shuffles > 0 && !shuffles_dev)) return TF_ERR_INVALID;
    if (n_obs < 0 || n_shuffles < 0 || n_paths < 0 || cap < 1) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    tf::find_path_gae_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int4*>(obstacles_dev), n_obs, bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5],
        source[0], source[1], source[2], destination[0], destination[1], destination[2], shuffles_dev, n_shuffles, n_paths,
        cap, reinterpret_cast<int4*>(paths_dev), lens_dev, result_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
