// tf_env.cu — environment objects (obstacle tables, voxel grid, hazard map), the voxeliser,
// and the small misc entry points of the C-ABI.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <mutex>

#include "tf_common.cuh"

namespace tf {

thread_local int g_last_cuda_error = 0;
long long g_launch_count = 0;
int g_disable_stream = 0;
int g_stream_schedule = 0;
int g_reserve_sms = 0;
// per-device caches, keyed by the CURRENT device of the calling thread (one process may drive several GPUs)
constexpr int kMaxDev = 64;
static int g_sm_count[kMaxDev] = {0};
static signed char g_dev_checked[kMaxDev] = {0};

int ensure_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) {
        cudaGetLastError();
        return TF_ERR_NODEVICE;
    }
    if (g_dev_checked[dev]) return g_dev_checked[dev] > 0 ? TF_OK : TF_ERR_NODEVICE;
    int major = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || major != 10) {
        cudaGetLastError();
        g_dev_checked[dev] = -1;  // no sm_100 device: there is deliberately no fallback
        return TF_ERR_NODEVICE;
    }
    g_sm_count[dev] = sms;
    g_dev_checked[dev] = 1;
    return TF_OK;
}
int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
    return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

int* sched_slot() {
    constexpr int kSlots = 4096;
    static int* rings[kMaxDev] = {nullptr};
    static unsigned next[kMaxDev] = {0};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    if (rings[dev] == nullptr) {
        std::lock_guard<std::mutex> lock(mu);
        if (rings[dev] == nullptr) {
            int* p = nullptr;
            if (cudaMalloc(&p, (size_t)kSlots * 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            if (cudaMemset(p, 0, (size_t)kSlots * 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return nullptr; }
            rings[dev] = p;
        }
    }
    const unsigned i = __atomic_fetch_add(&next[dev], 1u, __ATOMIC_RELAXED) % kSlots;
    return rings[dev] + 2 * i;
}

// ------------------------------------------------------------------ voxel kernels
__device__ __forceinline__ size_t brick_word(int x, int y, int z, int nbx, int nby) {
    size_t b = (size_t)(x >> 3) + (size_t)nbx * ((size_t)(y >> 3) + (size_t)nby * (size_t)(z >> 3));
    return b * 16 + ((z & 7) * 2 + ((y & 7) >> 2));
}

// One warp per obstacle: walk its bounding box, set the bit of every colliding lattice point.
// Collision rules and arithmetic are those of the FP32 fitness model (HR:17-23 / PF:28-44).
__global__ void voxelize_kernel(int semantics, const tf_obstacle* __restrict__ obs, int n, float safety,
                                uint32_t* __restrict__ bricks, int nx, int ny, int nz) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const tf_obstacle o = obs[w];
    const float s = o.size < 0.0f ? 1.0f : o.size;
    float rx = s, ry = s, rz = s;
    if (semantics == TF_SEM_HR) rx = ry = rz = __fadd_rn(s, safety);
    else if (o.kind == TF_KIND_CYLINDER) rz = __fmul_rn(2.0f, s);
    else if (o.kind == TF_KIND_AABB) { rx = o.hx; ry = o.hy; rz = o.hz; }
    const int x0 = max(0, (int)floorf(o.cx - rx)), x1 = min(nx - 1, (int)ceilf(o.cx + rx));
    const int y0 = max(0, (int)floorf(o.cy - ry)), y1 = min(ny - 1, (int)ceilf(o.cy + ry));
    const int z0 = max(0, (int)floorf(o.cz - rz)), z1 = min(nz - 1, (int)ceilf(o.cz + rz));
    if (x1 < x0 || y1 < y0 || z1 < z0) return;
    const int ex = x1 - x0 + 1, ey = y1 - y0 + 1, ez = z1 - z0 + 1;
    const long long total = (long long)ex * ey * ez;
    const float r = __fadd_rn(s, safety), r2 = __fmul_rn(r, r), s2 = __fmul_rn(s, s);
    for (long long i = lane; i < total; i += 32) {
        const int x = x0 + (int)(i % ex), y = y0 + (int)((i / ex) % ey), z = z0 + (int)(i / ((long long)ex * ey));
        const float dx = (float)x - o.cx, dy = (float)y - o.cy, dz = (float)z - o.cz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
        bool hit;
        if (semantics == TF_SEM_HR) hit = d2 <= r2;
        else if (o.kind == TF_KIND_SPHERE) hit = d2 <= s2;
        else hit = fabsf(dx) <= rx && fabsf(dy) <= ry && fabsf(dz) <= rz;
        if (hit) atomicOr(bricks + brick_word(x, y, z, nx >> 3, ny >> 3), 1u << (((y & 3) << 3) | (x & 7)));
    }
}

// One thread per brick word: gathers its 8x4 voxels from the linear bit grid.
__global__ void pack_linear_kernel(const uint32_t* __restrict__ lin, uint32_t* __restrict__ bricks, int nx, int ny,
                                   int nz, size_t nwords) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const int nbx = nx >> 3, nby = ny >> 3;
    const size_t brick = i >> 4;
    const int wi = (int)(i & 15);
    const int bx = (int)(brick % nbx), by = (int)((brick / nbx) % nby), bz = (int)(brick / ((size_t)nbx * nby));
    const int z = bz * 8 + (wi >> 1), ybase = by * 8 + (wi & 1) * 4, xbase = bx * 8;
    uint32_t out = 0;
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const size_t bit0 = (size_t)xbase + (size_t)nx * ((size_t)(ybase + yy) + (size_t)ny * z);  // multiple of 8
        const uint32_t byte = (lin[bit0 >> 5] >> (bit0 & 31)) & 0xffu;
        out |= byte << (yy * 8);
    }
    bricks[i] = out;
}

// Candidate index of the moving spheres.  One CTA per (time band, cell); a band covers `band` consecutive integer
// times.  For a sphere centre c(t), near(c) / far(c) = its distance to the nearest / farthest point of the cell (an
// axis-aligned box).  Whatever point p of the cell is asked about at time t, its nearest sphere is at most
// U_t = min_j far(c_j(t)) away, so a sphere can be NEAREST to some point of the cell only if near(c(t)) <= U_t, and it
// can COLLIDE with one only if near(c(t)) <= R.  A sphere is listed if either holds for any t of the band.  The cell's
// half extent carries a relative 1e-5 and an absolute 1e-3 margin for the float rounding of the cell assignment, and
// so do both comparisons.  (Round 2 first bounded the cell by its circumscribed sphere — |m - c| <= dmin + 2 rho; the
// box bounds list fewer spheres — cfg4: 1.68 instead of 1.78 per list — and a list of up to three needs no side table.)
constexpr int kIdxThreads = 128;
constexpr int kIdxMaxPerThread = 16;  // n_dyn <= 2048
__global__ void __launch_bounds__(kIdxThreads)
build_dyn_index_kernel(const float4* __restrict__ dynA, const float4* __restrict__ dynB, int n, int T, int band, int gx,
                       int gy, int gz, float ox, float oy, float oz, float h, int cap, uint2* __restrict__ heads,
                       uint16_t* __restrict__ ext, unsigned int ext_slots, unsigned int* __restrict__ ext_used) {
    // cap = words per list in ext (a multiple of 4); heads / ext: EnvDev::dcand / dext
    __shared__ float red[kIdxThreads];
    __shared__ int n_list;
    __shared__ unsigned int slot_s;
    __shared__ uint16_t lst[1024];
    const size_t ncell = (size_t)gx * gy * gz;
    const size_t cellid = blockIdx.x;
    const int tb = (int)(cellid / ncell);
    const size_t c = cellid % ncell;
    const int ix = (int)(c % gx), iy = (int)((c / gx) % gy), iz = (int)(c / ((size_t)gx * gy));
    const float mx = ox + ((float)ix + 0.5f) * h, my = oy + ((float)iy + 0.5f) * h, mz = oz + ((float)iz + 0.5f) * h;
    const float hh = 0.5f * h * 1.00001f + 1e-3f;   // half extent of the cell, with the margin
    uint32_t marked = 0;  // bit j: sphere threadIdx.x + j * kIdxThreads is a candidate
    if (threadIdx.x == 0) n_list = 0;
    for (int t = tb * band; t < min(T, (tb + 1) * band); ++t) {
        const float tm = (float)t;
        float umin = INFINITY;   // smallest far distance of the cell (to its farthest corner) over the spheres
        for (int k = threadIdx.x; k < n; k += kIdxThreads) {
            const float4 o = dynA[k], v = dynB[k];
            const float fx = fabsf(mx - fmaf(v.x, tm, o.x)) + hh, fy = fabsf(my - fmaf(v.y, tm, o.y)) + hh,
                        fz = fabsf(mz - fmaf(v.z, tm, o.z)) + hh;
            umin = fminf(umin, sqrtf(fx * fx + fy * fy + fz * fz));
        }
        __syncthreads();
        red[threadIdx.x] = umin;
        __syncthreads();
        for (int off = kIdxThreads / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + off]);
            __syncthreads();
        }
        const float near_lim = red[0] * 1.00001f + 1e-3f;
        int j = 0;
        for (int k = threadIdx.x; k < n; k += kIdxThreads, ++j) {
            const float4 o = dynA[k], v = dynB[k];
            const float nx = fmaxf(fabsf(mx - fmaf(v.x, tm, o.x)) - hh, 0.0f), ny = fmaxf(fabsf(my - fmaf(v.y, tm, o.y)) - hh, 0.0f),
                        nz = fmaxf(fabsf(mz - fmaf(v.z, tm, o.z)) - hh, 0.0f);
            const float d = sqrtf(nx * nx + ny * ny + nz * nz);   // near distance of the cell
            const float coll_lim = sqrtf(o.w) * 1.00001f + 1e-3f;
            if (d <= near_lim || d <= coll_lim) marked |= 1u << j;
        }
    }
    __syncthreads();
    int j = 0;
    for (int k = threadIdx.x; k < n; k += kIdxThreads, ++j) {
        if ((marked >> j) & 1u) {
            const int pos = atomicAdd(&n_list, 1);
            if (pos < cap) lst[pos] = (uint16_t)(2 * k);
        }
    }
    __syncthreads();
    // (a list is never empty — the nearest sphere of every time slice is on it — and the walk relies on that: an
    // empty one would be marked as overflowed)
    const int nl = n_list;
    const bool fits = nl >= 1 && nl <= cap;
    if (fits && nl > 3) {   // a slot in ext; when ext is full the list counts as overflowed and the host retries
        if (threadIdx.x == 0) slot_s = atomicAdd(ext_used, 1u);
        __syncthreads();
        if (slot_s < ext_slots)
            for (int i = threadIdx.x; i < cap; i += kIdxThreads) ext[(size_t)slot_s * cap + i] = lst[i < nl ? i : 0];
    }
    if (threadIdx.x == 0) {
        uint2 hd = make_uint2(0xffffu, 0u);
        if (fits && nl <= 3) {
            const uint32_t e1 = lst[0], e2 = nl > 1 ? lst[1] : e1, e3 = nl > 2 ? lst[2] : e1;
            hd = make_uint2((uint32_t)nl | (e1 << 16), e2 | (e3 << 16));
        } else if (fits && slot_s < ext_slots) {
            hd = make_uint2((uint32_t)nl | ((uint32_t)lst[0] << 16), slot_s);
        }
        heads[cellid] = hd;
    }
}

// The moving spheres as interleaved 32-byte records (c0, r2 | v, 0): one sector per sphere.
__global__ void dyn_records_kernel(const float4* __restrict__ dynA, const float4* __restrict__ dynB, int n,
                                   float4* __restrict__ rec) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    rec[2 * k] = dynA[k];
    rec[2 * k + 1] = dynB[k];
}

// ------------------------------------------------------------------ static candidate index
// 3-D cells.  For a cell with centre m and half diagonal rho, an obstacle whose (unclamped) distance
// function f_k is 1-Lipschitz satisfies  max(0, f_k(m) - rho) <= dist_k(p) <= max(0, f_k(m)) + rho  for every
// point p of the cell; k can be the nearest only if its lower bound does not exceed the smallest upper
// bound.  Eligible here: HR spheres (f = |m - c|), PF spheres / cubes (f = |m - c| - s) and aabbs (box
// distance).  PF cylinders measure distance along z only and go to the 1-D index below.  In addition every
// obstacle whose collision box, grown by half a cell, contains m is listed (it may collide).
__device__ __forceinline__ float static_f(const EnvDev& e, int semantics, int id, float mx, float my, float mz,
                                          bool& eligible, float (&half)[3]) {
    const float4 o = e.tabA[id];
    const float dx = mx - o.x, dy = my - o.y, dz = mz - o.z;
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    eligible = true;
    if (semantics == TF_SEM_HR) {
        const float R = sqrtf(o.w);
        half[0] = half[1] = half[2] = R;
        return d;
    }
    const int s_end = e.n_sphere, c_end = s_end + e.n_cube, y_end = c_end + e.n_cyl;
    if (id < c_end) { half[0] = half[1] = half[2] = o.w; return d - o.w; }
    if (id < y_end) { half[0] = half[1] = o.w; half[2] = 2.0f * o.w; eligible = false; return 0.0f; }
    const float4 h = e.tabB[id];
    half[0] = h.x; half[1] = h.y; half[2] = h.z;
    const float qx = fmaxf(fabsf(dx) - h.x, 0.0f), qy = fmaxf(fabsf(dy) - h.y, 0.0f), qz = fmaxf(fabsf(dz) - h.z, 0.0f);
    return sqrtf(qx * qx + qy * qy + qz * qz);
}

__global__ void __launch_bounds__(128)
build_static_index_kernel(const __grid_constant__ EnvDev e, int gx, int gy, int gz, float ox, float oy, float oz, float h,
                          int cap, int32_t* __restrict__ cand, int32_t* __restrict__ count) {
    __shared__ float red[128];
    __shared__ int n_list;
    const size_t c = blockIdx.x;
    const int ix = (int)(c % gx), iy = (int)((c / gx) % gy), iz = (int)(c / ((size_t)gx * gy));
    const float mx = ox + ((float)ix + 0.5f) * h, my = oy + ((float)iy + 0.5f) * h, mz = oz + ((float)iz + 0.5f) * h;
    const float rho = 0.8660254f * h * 1.00001f + 1e-3f, hh = 0.5f * h * 1.00001f + 1e-3f;
    float ub = INFINITY;
    for (int k = threadIdx.x; k < e.n; k += 128) {
        bool el;
        float half[3];
        const float f = static_f(e, e.semantics, k, mx, my, mz, el, half);
        if (el) ub = fminf(ub, fmaxf(f, 0.0f) + rho);
    }
    red[threadIdx.x] = ub;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + off]);
        __syncthreads();
    }
    const float lim = red[0] * 1.00001f + 1e-3f;
    for (int k = threadIdx.x; k < e.n; k += 128) {
        bool el;
        float half[3];
        const float f = static_f(e, e.semantics, k, mx, my, mz, el, half);
        const float4 o = e.tabA[k];
        const bool near = el && fmaxf(f - rho, 0.0f) <= lim;
        const bool coll = fabsf(mx - o.x) <= half[0] + hh && fabsf(my - o.y) <= half[1] + hh && fabsf(mz - o.z) <= half[2] + hh;
        if (near || coll) {
            const int pos = atomicAdd(&n_list, 1);
            if (pos < cap) cand[c * cap + pos] = k;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) count[c] = n_list <= cap ? n_list : -1;
}

// 1-D z cells for the PF cylinders (distance max(0, |dz| - 2 s), PF:56-57).  If some cylinder's slab
// contains the whole cell its distance is exactly 0 everywhere in the cell — the global minimum — and
// that single cylinder is a sufficient candidate; otherwise the lower / upper bound rule as above.
__global__ void __launch_bounds__(128)
build_cyl_index_kernel(const __grid_constant__ EnvDev e, int nz, float oz, float hz, int cap,
                       int32_t* __restrict__ cand, int32_t* __restrict__ count) {
    __shared__ float red[128];
    __shared__ int covered[128];
    __shared__ int n_list;
    const int c = blockIdx.x;
    const float mz = oz + ((float)c + 0.5f) * hz;
    const float dlt = 0.5f * hz * 1.00001f + 1e-3f;
    const int y0 = e.n_sphere + e.n_cube, y1 = y0 + e.n_cyl;
    float ub = INFINITY;
    int cov = 0x7fffffff;
    for (int k = y0 + threadIdx.x; k < y1; k += 128) {
        const float4 o = e.tabA[k];
        const float f = fabsf(mz - o.z) - 2.0f * o.w;
        ub = fminf(ub, fmaxf(f, 0.0f) + dlt);
        if (f + dlt <= -1e-3f) cov = min(cov, k);
    }
    red[threadIdx.x] = ub;
    covered[threadIdx.x] = cov;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            red[threadIdx.x] = fminf(red[threadIdx.x], red[threadIdx.x + off]);
            covered[threadIdx.x] = min(covered[threadIdx.x], covered[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (covered[0] != 0x7fffffff) {
        if (threadIdx.x == 0) { cand[(size_t)c * cap] = covered[0]; count[c] = 1; }
        return;
    }
    const float lim = red[0] * 1.00001f + 1e-3f;
    for (int k = y0 + threadIdx.x; k < y1; k += 128) {
        const float4 o = e.tabA[k];
        const float f = fabsf(mz - o.z) - 2.0f * o.w;
        if (fmaxf(f - dlt, 0.0f) <= lim) {
            const int pos = atomicAdd(&n_list, 1);
            if (pos < cap) cand[(size_t)c * cap + pos] = k;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) count[c] = n_list <= cap ? n_list : -1;
}

// [t][y][x] -> corner-packed [y][x][t]: (v[t][y][x], v[t][y][x1] - v[t][y][x], v[t1][y][x], v[t1][y][x1] - v[t1][y][x]),
// x1 / t1 clamped.  The x differences are the FP32 subtractions hazard_sample's lerp would do per sample (same
// operands, same rounding): stored once, they turn its four x-lerps into single fmas.
__global__ void hazard_pack_kernel(const float* __restrict__ in, float4* __restrict__ out, int nt, int ny, int nx) {
    const size_t nyx = (size_t)ny * nx, n = nyx * (size_t)nt;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / nt;
        const int t = (int)(i - c * nt), x = (int)(c % nx);
        const size_t c1 = x + 1 < nx ? c + 1 : c;
        const size_t t1 = t + 1 < nt ? t + 1 : t;
        const float v00 = in[(size_t)t * nyx + c], v01 = in[(size_t)t * nyx + c1], v10 = in[t1 * nyx + c], v11 = in[t1 * nyx + c1];
        out[i] = make_float4(v00, __fsub_rn(v01, v00), v10, __fsub_rn(v11, v10));
    }
}

}  // namespace tf

using namespace tf;

// ------------------------------------------------------------------ env lifecycle
static int upload_static(tf_env* env, int semantics, const std::vector<double>& obs64, int n, double safety) {
    // kind-sort (stable) and pre-square for the FP32 tables
    std::vector<int> order;
    int cnt[4] = {0, 0, 0, 0};
    if (semantics == TF_SEM_PF) {
        for (int kind = 0; kind < 4; ++kind)
            for (int i = 0; i < n; ++i) {
                int k = (int)obs64[8 * i + 7];
                if (k < 0 || k > 3) return TF_ERR_INVALID;
                if (k == kind) { order.push_back(i); ++cnt[kind]; }
            }
    } else {
        for (int i = 0; i < n; ++i) order.push_back(i);
        cnt[0] = n;
    }
    std::vector<float> A((size_t)n * 4), B((size_t)n * 4, 0.0f);
    for (int j = 0; j < n; ++j) {
        const double* o = &obs64[8 * order[j]];
        A[4 * j] = (float)o[0]; A[4 * j + 1] = (float)o[1]; A[4 * j + 2] = (float)o[2];
        if (semantics == TF_SEM_HR) {
            float size = (float)o[3] < 0.0f ? 1.0f : (float)o[3];  // HR:20 default size 1.0
            float r = size + (float)safety;
            A[4 * j + 3] = r * r;  // compared against d^2 (numerics)
        } else {
            A[4 * j + 3] = (float)o[3];
            B[4 * j] = (float)o[4]; B[4 * j + 1] = (float)o[5]; B[4 * j + 2] = (float)o[6];
        }
    }
    size_t bytesA = (size_t)n * 16, bytes64 = (size_t)n * 64;
    void* block = nullptr;
    if (n > 0) {
        TF_CUDA(cudaMalloc(&block, bytesA * 2 + bytes64));
        TF_CUDA(cudaMemcpy(block, A.data(), bytesA, cudaMemcpyHostToDevice));
        TF_CUDA(cudaMemcpy((char*)block + bytesA, B.data(), bytesA, cudaMemcpyHostToDevice));
        TF_CUDA(cudaMemcpy((char*)block + 2 * bytesA, obs64.data(), bytes64, cudaMemcpyHostToDevice));
    }
    env->dev_block = block;
    EnvDev& d = env->d;
    d.semantics = semantics;
    d.n = n;
    d.n_sphere = cnt[0]; d.n_cube = cnt[1]; d.n_cyl = cnt[2]; d.n_aabb = cnt[3];
    d.n_smem = n;
    d.smax_sc = 0.0f;
    if (semantics == TF_SEM_PF)
        for (int j = 0; j < cnt[0] + cnt[1]; ++j) d.smax_sc = fmaxf(d.smax_sc, A[4 * j + 3]);
    d.tabA = (const float4*)block;
    d.tabB = n > 0 ? (const float4*)((char*)block + bytesA) : nullptr;
    d.obs64 = n > 0 ? (const double*)((char*)block + 2 * bytesA) : nullptr;
    d.safety64 = safety;
    return TF_OK;
}

static int env_create_common(int semantics, const std::vector<double>& obs64, int n, double safety, tf_env** out) {
    if (!out) return TF_ERR_INVALID;
    *out = nullptr;
    if (semantics < TF_SEM_NONE || semantics > TF_SEM_PF || n < 0) return TF_ERR_INVALID;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    tf_env* env = (tf_env*)calloc(1, sizeof(tf_env));
    if (!env) return TF_ERR_NOMEM;
    if (semantics == TF_SEM_NONE) n = 0;
    rc = upload_static(env, semantics, obs64, n, safety);
    if (rc != TF_OK) { free(env); return rc; }
    *out = env;
    return TF_OK;
}

extern "C" int tf_env_create(int semantics, const tf_obstacle* obs_host, int n, float safety_radius, tf_env** out) {
    if (n > 0 && !obs_host) return TF_ERR_INVALID;
    std::vector<double> o64((size_t)(n > 0 ? n : 0) * 8);
    for (int i = 0; i < n; ++i) {
        const tf_obstacle& o = obs_host[i];
        double* d = &o64[8 * (size_t)i];
        d[0] = o.cx; d[1] = o.cy; d[2] = o.cz; d[3] = o.size; d[4] = o.hx; d[5] = o.hy; d[6] = o.hz; d[7] = (double)o.kind;
    }
    return env_create_common(semantics, o64, n, (double)safety_radius, out);
}

// f64 records: n x 8 doubles (cx cy cz size hx hy hz kind) — keeps positions/sizes that are not
// representable in FP32 exact for the reference-order kernels.
extern "C" int tf_env_create_f64(int semantics, const double* obs64_host, int n, double safety_radius, tf_env** out) {
    if (n > 0 && !obs64_host) return TF_ERR_INVALID;
    std::vector<double> o64(obs64_host, obs64_host + (size_t)(n > 0 ? n : 0) * 8);
    return env_create_common(semantics, o64, n, safety_radius, out);
}

extern "C" int tf_env_set_dynamic(tf_env* env, const tf_dyn_obstacle* dyn_host, int n, float safety_radius) {
    if (!env || n < 0 || (n > 0 && !dyn_host)) return TF_ERR_INVALID;
    if (n > 2048) return TF_ERR_SHAPE;
    if (env->dyn_block) { cudaFree(env->dyn_block); env->dyn_block = nullptr; }
    env->d.n_dyn = env->d.n_dyn_smem = 0; env->d.dynA = env->d.dynB = nullptr;
    if (env->dindex_block) { cudaFree(env->dindex_block); env->dindex_block = nullptr; }
    env->d.dcand = nullptr; env->d.dext = nullptr; env->d.drec = nullptr;
    if (n == 0) return TF_OK;
    std::vector<float> A((size_t)n * 4), B((size_t)n * 4);
    for (int i = 0; i < n; ++i) {
        const tf_dyn_obstacle& o = dyn_host[i];
        float size = o.size < 0.0f ? 1.0f : o.size;
        float r = size + safety_radius;
        A[4 * i] = o.cx; A[4 * i + 1] = o.cy; A[4 * i + 2] = o.cz; A[4 * i + 3] = r * r;
        B[4 * i] = o.vx; B[4 * i + 1] = o.vy; B[4 * i + 2] = o.vz; B[4 * i + 3] = 0.0f;
    }
    TF_CUDA(cudaMalloc(&env->dyn_block, (size_t)n * 32));
    TF_CUDA(cudaMemcpy(env->dyn_block, A.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    TF_CUDA(cudaMemcpy((char*)env->dyn_block + (size_t)n * 16, B.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    env->d.n_dyn = env->d.n_dyn_smem = n;
    env->d.dynA = (const float4*)env->dyn_block;
    env->d.dynB = (const float4*)((char*)env->dyn_block + (size_t)n * 16);
    return TF_OK;
}

extern "C" int tf_env_build_dynamic_index(tf_env* env, int time_slices, int band, const float box_lo[3],
                                          const float box_hi[3], float cell, int cap, void* stream) {
    if (!env || !box_lo || !box_hi) return TF_ERR_INVALID;
    if (env->d.n_dyn <= 0) return TF_ERR_INVALID;
    if (time_slices < 1 || time_slices > 4096 || band < 1 || cell <= 0.0f || cap < 1 || cap > 1024) return TF_ERR_SHAPE;
    if (env->d.n_dyn > kIdxThreads * kIdxMaxPerThread) return TF_ERR_SHAPE;
    // the kernels take the band of a time by multiply-shift; where that is not exact for every t < time_slices the
    // band is widened to the next power of two (always exact; the index is built for the band actually used)
    auto band_magic = [&](int b) {
        uint32_t m = (uint32_t)(((1u << 20) + (uint32_t)b - 1u) / (uint32_t)b);
        for (int t = 0; t < time_slices && m; ++t)
            if ((((uint32_t)t * m) >> 20) != (uint32_t)(t / b)) m = 0;
        return m;
    };
    uint32_t magic = band_magic(band);
    if (magic == 0) {
        int b2 = 1;
        while (b2 < band) b2 <<= 1;
        band = b2;
        magic = band_magic(band);
    }
    const int nbands = (time_slices + band - 1) / band;
    cap = (cap + 3) & ~3;  // words per list in ext: groups of four entries
    int g[3];
    for (int a = 0; a < 3; ++a) {
        if (!(box_hi[a] > box_lo[a])) return TF_ERR_SHAPE;
        g[a] = (int)ceilf((box_hi[a] - box_lo[a]) / cell);
        if (g[a] < 1 || g[a] > 256) return TF_ERR_SHAPE;
    }
    const size_t ncell = (size_t)g[0] * g[1] * g[2];
    const size_t nlist = ncell * (size_t)nbands;
    if (nlist > (size_t)1 << 26) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (env->dindex_block) { cudaFree(env->dindex_block); env->dindex_block = nullptr; }
    env->d.dcand = nullptr; env->d.dext = nullptr; env->d.drec = nullptr;
    // [records | heads | ext | ext counter]; ext is sized for an eighth of the lists being longer than three entries and
    // the build is repeated with the demanded size when that was not enough
    auto r256 = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t bytes_rec = r256((size_t)env->d.n_dyn * 32), bytes_heads = r256(nlist * 8);
    size_t ext_slots = nlist / 8 + 256;
    float4* rec = nullptr;
    uint2* heads = nullptr;
    uint16_t* ext = nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t bytes_ext = r256(ext_slots * (size_t)cap * 2);
        TF_CUDA(cudaMalloc(&env->dindex_block, bytes_rec + bytes_heads + bytes_ext + 256));
        rec = (float4*)env->dindex_block;
        heads = (uint2*)((char*)env->dindex_block + bytes_rec);
        ext = (uint16_t*)((char*)env->dindex_block + bytes_rec + bytes_heads);
        unsigned int* used = (unsigned int*)((char*)env->dindex_block + bytes_rec + bytes_heads + bytes_ext);
        TF_CUDA(cudaMemsetAsync(used, 0, 4, (cudaStream_t)stream));
        build_dyn_index_kernel<<<(unsigned)nlist, kIdxThreads, 0, (cudaStream_t)stream>>>(
            env->d.dynA, env->d.dynB, env->d.n_dyn, time_slices, band, g[0], g[1], g[2], box_lo[0], box_lo[1], box_lo[2], cell,
            cap, heads, ext, (unsigned int)ext_slots, used);
        TF_LAUNCH_CHECK();
        unsigned int demanded = 0;
        TF_CUDA(cudaMemcpyAsync(&demanded, used, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        TF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        if ((size_t)demanded <= ext_slots) break;
        cudaFree(env->dindex_block);
        env->dindex_block = nullptr;
        ext_slots = demanded;
    }
    dyn_records_kernel<<<(env->d.n_dyn + 255) / 256, 256, 0, (cudaStream_t)stream>>>(env->d.dynA, env->d.dynB, env->d.n_dyn, rec);
    TF_LAUNCH_CHECK();
    TF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    env->d.dcand = heads; env->d.dext = ext; env->d.drec = rec;
    env->d.dT = time_slices; env->d.dband = band; env->d.dgx = g[0]; env->d.dgy = g[1]; env->d.dgz = g[2]; env->d.dcap = cap;
    env->d.dox = box_lo[0]; env->d.doy = box_lo[1]; env->d.doz = box_lo[2]; env->d.dinv_h = 1.0f / cell;
    env->d.dnox = -box_lo[0] * env->d.dinv_h; env->d.dnoy = -box_lo[1] * env->d.dinv_h; env->d.dnoz = -box_lo[2] * env->d.dinv_h;
    env->d.dband_magic = magic;
    env->d.n_dyn_smem = 0;  // the kernels read the records through L1 / L2 now
    return TF_OK;
}

extern "C" int tf_env_dynamic_index_stats(const tf_env* env, int64_t* n_lists, int64_t* n_overflow, double* mean_len,
                                          int* max_len) {
    if (!env || !env->d.dcand) return TF_ERR_INVALID;
    const size_t n = (size_t)((env->d.dT + env->d.dband - 1) / env->d.dband) * env->d.dgx * env->d.dgy * env->d.dgz;
    std::vector<uint16_t> h(n);
    TF_CUDA(cudaMemcpy2D(h.data(), 2, env->d.dcand, 8, 2, n, cudaMemcpyDeviceToHost));   // word 0 of every 8-byte head
    int64_t over = 0;
    double sum = 0.0;
    int mx = 0;
    for (size_t i = 0; i < n; ++i) {
        if (h[i] == 0xffff) ++over;
        else { sum += h[i]; if (h[i] > mx) mx = h[i]; }
    }
    if (n_lists) *n_lists = (int64_t)n;
    if (n_overflow) *n_overflow = over;
    if (mean_len) *mean_len = n > (size_t)over ? sum / (double)(n - over) : 0.0;
    if (max_len) *max_len = mx;
    return TF_OK;
}

extern "C" int tf_env_build_static_index(tf_env* env, const float box_lo[3], const float box_hi[3], float cell, int cap,
                                         float z_cell, int z_cap, void* stream) {
    if (!env || !box_lo || !box_hi) return TF_ERR_INVALID;
    EnvDev& d = env->d;
    if (d.n <= 0 || (d.semantics != TF_SEM_HR && d.semantics != TF_SEM_PF)) return TF_ERR_INVALID;
    if (cell <= 0.0f || cap < 1 || cap > 1024 || z_cell <= 0.0f || z_cap < 1 || z_cap > 1024) return TF_ERR_SHAPE;
    int g[3];
    for (int a = 0; a < 3; ++a) {
        if (!(box_hi[a] > box_lo[a])) return TF_ERR_SHAPE;
        g[a] = (int)ceilf((box_hi[a] - box_lo[a]) / cell);
        if (g[a] < 1 || g[a] > 512) return TF_ERR_SHAPE;
    }
    const size_t ncell = (size_t)g[0] * g[1] * g[2];
    if (ncell > (size_t)1 << 24) return TF_ERR_SHAPE;
    const int nz = (int)ceilf((box_hi[2] - box_lo[2]) / z_cell);
    if (nz < 1 || nz > (1 << 20)) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (env->sindex_block) { cudaFree(env->sindex_block); env->sindex_block = nullptr; }
    d.scand = d.scount = d.zcand = d.zcount = nullptr;
    d.n_smem = d.n;
    auto r256 = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t b_scount = r256(ncell * 4), b_scand = r256(ncell * (size_t)cap * 4);
    const size_t b_zcount = r256((size_t)nz * 4), b_zcand = r256((size_t)nz * z_cap * 4);
    TF_CUDA(cudaMalloc(&env->sindex_block, b_scount + b_scand + b_zcount + b_zcand));
    char* base = (char*)env->sindex_block;
    int32_t* scount = (int32_t*)base;
    int32_t* scand = (int32_t*)(base + b_scount);
    int32_t* zcount = (int32_t*)(base + b_scount + b_scand);
    int32_t* zcand = (int32_t*)(base + b_scount + b_scand + b_zcount);
    cudaStream_t st = (cudaStream_t)stream;
    build_static_index_kernel<<<(unsigned)ncell, 128, 0, st>>>(d, g[0], g[1], g[2], box_lo[0], box_lo[1], box_lo[2], cell, cap,
                                                              scand, scount);
    TF_LAUNCH_CHECK();
    if (d.semantics == TF_SEM_PF && d.n_cyl > 0) {
        build_cyl_index_kernel<<<nz, 128, 0, st>>>(d, nz, box_lo[2], z_cell, z_cap, zcand, zcount);
        TF_LAUNCH_CHECK();
    }
    TF_CUDA(cudaStreamSynchronize(st));
    d.scand = scand; d.scount = scount; d.zcand = zcand; d.zcount = zcount;
    d.sgx = g[0]; d.sgy = g[1]; d.sgz = g[2]; d.scap = cap;
    d.sox = box_lo[0]; d.soy = box_lo[1]; d.soz = box_lo[2]; d.sinv_h = 1.0f / cell;
    d.znz = nz; d.zcap = z_cap; d.zoz = box_lo[2]; d.zinv_h = 1.0f / z_cell;
    d.n_smem = 0;  // the kernels read the candidate records through L1 / L2
    return TF_OK;
}

extern "C" int tf_env_static_index_stats(const tf_env* env, int64_t* n_lists, int64_t* n_overflow, double* mean_len,
                                         int* max_len, double* z_mean_len) {
    if (!env || !env->d.scount) return TF_ERR_INVALID;
    const size_t n = (size_t)env->d.sgx * env->d.sgy * env->d.sgz;
    std::vector<int32_t> h(n), hz(env->d.znz);
    TF_CUDA(cudaMemcpy(h.data(), env->d.scount, n * 4, cudaMemcpyDeviceToHost));
    int64_t over = 0;
    double sum = 0.0, zsum = 0.0;
    int mx = 0;
    for (size_t i = 0; i < n; ++i) {
        if (h[i] < 0) ++over;
        else { sum += h[i]; if (h[i] > mx) mx = h[i]; }
    }
    if (env->d.semantics == TF_SEM_PF && env->d.n_cyl > 0) {
        TF_CUDA(cudaMemcpy(hz.data(), env->d.zcount, hz.size() * 4, cudaMemcpyDeviceToHost));
        for (int v : hz) { if (v < 0) ++over; else zsum += v; }
    }
    if (n_lists) *n_lists = (int64_t)n;
    if (n_overflow) *n_overflow = over;
    if (mean_len) *mean_len = sum / (double)n;
    if (max_len) *max_len = mx;
    if (z_mean_len) *z_mean_len = hz.empty() ? 0.0 : zsum / (double)hz.size();
    return TF_OK;
}

extern "C" int tf_env_set_voxels(tf_env* env, const uint32_t* bricks_dev, int nx, int ny, int nz) {
    if (!env) return TF_ERR_INVALID;
    if (!bricks_dev) { env->d.bricks = nullptr; return TF_OK; }
    if (nx <= 0 || ny <= 0 || nz <= 0 || (nx & 7) || (ny & 7) || (nz & 7)) return TF_ERR_SHAPE;
    env->d.bricks = bricks_dev;
    env->d.nx = nx; env->d.ny = ny; env->d.nz = nz;
    env->d.nbx = nx >> 3; env->d.nby = ny >> 3;
    return TF_OK;
}

extern "C" int tf_env_set_hazard(tf_env* env, const float* hazard_dev, int nt, int ny, int nx, float sx, float sy,
                                 float ox, float oy, float st) {
    if (!env) return TF_ERR_INVALID;
    if (env->hazard_block) { cudaFree(env->hazard_block); env->hazard_block = nullptr; }
    if (!hazard_dev) { env->d.hazard = nullptr; return TF_OK; }
    if (nt <= 0 || ny <= 0 || nx <= 0) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    // the env owns a corner-packed copy in [y][x][t] order (see hazard_sample); the caller's tensor is not kept
    const size_t n = (size_t)nt * ny * nx;
    if (n >= ((size_t)1 << 29)) return TF_ERR_SHAPE;  // the kernels index the packed map with 32 bits
    TF_CUDA(cudaMalloc(&env->hazard_block, n * 16));
    hazard_pack_kernel<<<(unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256>>>(
        hazard_dev, (float4*)env->hazard_block, nt, ny, nx);
    TF_LAUNCH_CHECK();
    TF_CUDA(cudaDeviceSynchronize());
    env->d.hazard = (const float4*)env->hazard_block;
    env->d.nt = nt; env->d.hny = ny; env->d.hnx = nx;
    env->d.sx = sx; env->d.sy = sy; env->d.ox = ox; env->d.oy = oy; env->d.st = st;
    return TF_OK;
}

extern "C" int tf_env_destroy(tf_env* env) {
    if (!env) return TF_OK;
    if (env->dev_block) cudaFree(env->dev_block);
    if (env->dyn_block) cudaFree(env->dyn_block);
    if (env->dindex_block) cudaFree(env->dindex_block);
    if (env->hazard_block) cudaFree(env->hazard_block);
    if (env->sindex_block) cudaFree(env->sindex_block);
    free(env);
    return TF_OK;
}

// ------------------------------------------------------------------ voxel helpers
extern "C" int64_t tf_voxel_grid_bytes(int nx, int ny, int nz) {
    if (nx <= 0 || ny <= 0 || nz <= 0 || (nx & 7) || (ny & 7) || (nz & 7)) return TF_ERR_SHAPE;
    return (int64_t)(nx >> 3) * (ny >> 3) * (nz >> 3) * 64;
}

extern "C" int tf_voxelize(int semantics, const tf_obstacle* obs_dev, int n, float safety_radius, uint32_t* bricks_dev,
                           int nx, int ny, int nz, void* stream) {
    if (!bricks_dev || (n > 0 && !obs_dev) || n < 0) return TF_ERR_INVALID;
    if (semantics != TF_SEM_HR && semantics != TF_SEM_PF) return TF_ERR_INVALID;
    if (tf_voxel_grid_bytes(nx, ny, nz) < 0) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    if (n == 0) return TF_OK;
    const int threads = 256, warps_per_block = threads / 32;
    voxelize_kernel<<<(n + warps_per_block - 1) / warps_per_block, threads, 0, (cudaStream_t)stream>>>(
        semantics, obs_dev, n, safety_radius, bricks_dev, nx, ny, nz);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_voxel_pack_linear(const uint32_t* linear_dev, uint32_t* bricks_dev, int nx, int ny, int nz,
                                    void* stream) {
    if (!linear_dev || !bricks_dev) return TF_ERR_INVALID;
    int64_t bytes = tf_voxel_grid_bytes(nx, ny, nz);
    if (bytes < 0) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    size_t nwords = (size_t)bytes / 4;
    pack_linear_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, (cudaStream_t)stream>>>(linear_dev, bricks_dev, nx,
                                                                                         ny, nz, nwords);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// ------------------------------------------------------------------ misc
extern "C" void tf_params_default(tf_params* p) {
    if (!p) return;
    p->w_safety = 0.4f; p->w_energy = 0.2f; p->w_wind = 0.2f; p->w_time = 0.1f; p->w_turn = 0.1f;  // HR:55-59
    p->dist_zero_guard = 1;  // HR:53
    p->hazard_weight = 0.0f;
    p->samples_per_segment = 1;
}

extern "C" void tf_ga_params_default(tf_ga_params* p) {
    if (!p) return;
    p->n_elite = 2;  // HR:115
    p->selection = TF_SEL_TRUNCATION;
    p->n_top = 10;  // HR:118
    p->tournament_k = 4;
    p->crossover = TF_CX_SINGLE_POINT;
    p->mutation = TF_MUT_LATTICE;
    p->mutation_rate = 0.1f;  // HR:62
    p->sigma = 1.0f;
    p->stream_id = 0;
    p->reserve_sms = 0;
}

extern "C" int tf_version(void) { return TF_VERSION; }

extern "C" const char* tf_strerror(int status) {
    switch (status) {
    case TF_OK: return "ok";
    case TF_ERR_INVALID: return "invalid argument (null pointer or bad enum)";
    case TF_ERR_SHAPE: return "shape out of supported range";
    case TF_ERR_UNSUPPORTED: return "unsupported combination";
    case TF_ERR_CUDA: return "CUDA runtime error (see tf_last_cuda_error)";
    case TF_ERR_NOMEM: return "out of host memory";
    case TF_ERR_NODEVICE: return "no sm_100 (B200) device: libtracefly_b200 has no fallback path";
    case TF_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown tf_status";
    }
}

extern "C" int tf_last_cuda_error(void) { return tf::g_last_cuda_error; }

extern "C" int tf_device_info(int* sms, int* cc_major, int* cc_minor, int64_t* hbm_bytes) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        return TF_ERR_NODEVICE;
    }
    if (sms) *sms = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    return TF_OK;
}

extern "C" void tf_philox4x32_10(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t out[4]) {
    tf::U4 r = tf::philox4x32_10(key0, key1, c0, c1, c2, c3);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

namespace tf {
__global__ void selftest_sqrt_kernel(uint32_t lo, uint64_t n, unsigned long long* bad) {
    unsigned long long mine = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float(lo + (uint32_t)i);
        const float xs[2] = {x, 2.0f};
        float rs[2];
        sqrt_rn_n<2>(xs, rs);   // the batched form (it folds x == 0 in differently)
        const uint32_t a = __float_as_uint(sqrt_rn(x)), a2 = __float_as_uint(rs[0]), b = __float_as_uint(__fsqrt_rn(x));
        const bool nan_b = isnan(__uint_as_float(b));
        mine += (a != b && !(isnan(__uint_as_float(a)) && nan_b)) ? 1 : 0;
        mine += (a2 != b && !(isnan(__uint_as_float(a2)) && nan_b)) ? 1 : 0;
    }
    if (mine) atomicAdd(bad, mine);
}
}  // namespace tf

namespace tf {
__global__ void selftest_rcp_kernel(uint32_t lo, uint64_t n, unsigned long long* bad) {
    unsigned long long mine = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float y = __uint_as_float(lo + (uint32_t)i);
        const float ys[2] = {y, 1.5f};
        float rs[2];
        rcp_rn_n<2>(ys, rs);
        const uint32_t a = __float_as_uint(rcp_rn(y)), a2 = __float_as_uint(rs[0]), b = __float_as_uint(__fdiv_rn(1.0f, y));
        const bool nan_b = isnan(__uint_as_float(b));
        mine += (a != b && !(isnan(__uint_as_float(a)) && nan_b)) ? 1 : 0;
        mine += (a2 != b && !(isnan(__uint_as_float(a2)) && nan_b)) ? 1 : 0;
    }
    if (mine) atomicAdd(bad, mine);
}
}  // namespace tf

extern "C" int tf_selftest_rcp(uint32_t first_bits, uint64_t count, uint64_t* mismatches_dev, void* stream) {
    if (!mismatches_dev || count > (1ull << 32)) return TF_ERR_INVALID;
    if (count == 0) return TF_OK;
    tf::selftest_rcp_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(first_bits, count,
                                                                     (unsigned long long*)mismatches_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_selftest_sqrt(uint32_t first_bits, uint64_t count, uint64_t* mismatches_dev, void* stream) {
    if (!mismatches_dev || count > (1ull << 32)) return TF_ERR_INVALID;
    if (count == 0) return TF_OK;
    tf::selftest_sqrt_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(first_bits, count,
                                                                      (unsigned long long*)mismatches_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

extern "C" int tf_set_option(int option, int value) {
    if (option == TF_OPT_DISABLE_STREAM) { tf::g_disable_stream = value != 0; return TF_OK; }
    if (option == TF_OPT_RESERVE_SMS && value >= 0) { tf::g_reserve_sms = value; return TF_OK; }
    if (option == TF_OPT_STREAM_SCHEDULE && value >= 0 && value <= 2) { tf::g_stream_schedule = value; return TF_OK; }
    return TF_ERR_INVALID;
}

extern "C" int64_t tf_launch_count(void) { return (int64_t)tf::g_launch_count; }
