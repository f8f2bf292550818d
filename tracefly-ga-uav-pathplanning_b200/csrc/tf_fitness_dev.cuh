// tf_fitness_dev.cuh — device code of the FP32 fitness evaluation, one warp per chromosome.
//
// Replaces Path.calculate_fitness (HR:42-59 / PF:144-170) with is_collision (HR:17-23),
// distance_to_nearest_obstacle (HR:25-28), ComplexEnvironment.is_obstacle (PF:28-44) and
// distance_to_obstacles (PF:46-59).  The arithmetic order is the FP32 fixed-order model that
// oracle/tf_oracle.c restates (orc_fitness_one_f32): lane l owns waypoints l, l+32, ...;
// per-lane partial sums are accumulated in slot order and reduced with an xor butterfly.
// Compiled with --fmad=false: every FMA below is explicit.
#pragma once
#include "tf_common.cuh"

namespace tf {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlockThreads = kWarpsPerBlock * 32;
constexpr int kTableCap = 2048;  // static obstacles staged per chunk (32 KB of float4)
constexpr int kDynCap = 2048;

struct Tables {  // shared-memory staging area handed to the evaluators
    const float4* A;
    const float4* B;
    const float4* dA;
    const float4* dB;
};

// One lane's slice of a chromosome: waypoint k = lane + 32*s lives in slot s.
template <int NW>
struct Lanes {
    float x[NW], y[NW], z[NW];
};

__device__ __forceinline__ float dist2(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dy, dy, __fmul_rn(dx, dx)));
}

// Pruning bound for the sphere / cube distance (see eval_static_chunk): candidates with
// d^2 >= bound cannot lower `best`.  best = inf -> bound = inf (nothing is pruned).
__device__ __forceinline__ float prune_bound(float best, float smax) {
    const float t = __fadd_rn(best, smax);
    return __fmul_rn(__fmul_rn(t, t), 1.00000095367431640625f);  // 1 + 2^-20
}

// Static analytic obstacles, table rows [lo, hi) of the staged chunk whose first row has global
// index `base`.  Kind buckets are global index ranges (tables are kind-sorted).
template <int SEM, int NP>
__device__ __forceinline__ void eval_static_chunk(const EnvDev& e, const Tables& t, int base, int cnt,
                                                  const float (&x)[NP], const float (&y)[NP], const float (&z)[NP],
                                                  uint32_t& hit, float (&best)[NP]) {
    if constexpr (SEM == TF_SEM_HR) {
        // HR:17-28.  best[] carries min d^2 here; the caller takes one sqrt at the end.
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float4 o = t.A[k];
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                float d2 = dist2(x[s] - o.x, y[s] - o.y, z[s] - o.z);
                hit |= (d2 <= o.w) ? (1u << s) : 0u;
                best[s] = fminf(best[s], d2);
            }
        }
    } else if constexpr (SEM == TF_SEM_PF) {
        // chunk-local row ranges of each kind (tables are kind-sorted: spheres, cubes, cylinders, aabbs)
        const int s_end = e.n_sphere, c_end = s_end + e.n_cube, y_end = c_end + e.n_cyl;
        const int s_hi = min(cnt, max(0, s_end - base));
        const int c_hi = min(cnt, max(0, c_end - base));
        const int y_hi = min(cnt, max(0, y_end - base));
        // cylinders first PF:40-42 (AABB s,s,2s), PF:56-57 (z-only distance): they are cheap and pull
        // `best` down early, which makes the pruning below bite (min / or are order-independent)
#pragma unroll 2
        for (int k = c_hi; k < y_hi; ++k) {
            const float4 o = t.A[k];
            const float h = __fmul_rn(2.0f, o.w);
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                float az = fabsf(z[s] - o.z);
                bool in = fabsf(x[s] - o.x) <= o.w && fabsf(y[s] - o.y) <= o.w && az <= h;
                hit |= in ? (1u << s) : 0u;
                best[s] = fminf(best[s], fmaxf(0.0f, __fsub_rn(az, h)));
            }
        }
        // Spheres and cubes share the distance max(0, ||p-c|| - s).  The sqrt is skipped whenever
        // d^2 >= fl(fl(t*t) * (1 + 2^-20)), t = fl(best + s_max): then fl(sqrt(d^2) - s) >= best is
        // guaranteed, so the minimum — and every bit of the result — is unchanged.
        const float smax = e.smax_sc;
        float bound[NP];
#pragma unroll
        for (int s = 0; s < NP; ++s) bound[s] = prune_bound(best[s], smax);
        // spheres PF:37-39, PF:54-55
#pragma unroll 2
        for (int k = 0; k < s_hi; ++k) {
            const float4 o = t.A[k];
            const float s2 = __fmul_rn(o.w, o.w);
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                float d2 = dist2(x[s] - o.x, y[s] - o.y, z[s] - o.z);
                hit |= (d2 <= s2) ? (1u << s) : 0u;
                if (d2 < bound[s]) {
                    best[s] = fminf(best[s], fmaxf(0.0f, __fsub_rn(sqrt_rn(d2), o.w)));
                    bound[s] = prune_bound(best[s], smax);
                }
            }
        }
        // cubes PF:34-36 (Chebyshev collision), PF:52-53 (spherical distance)
#pragma unroll 2
        for (int k = s_hi; k < c_hi; ++k) {
            const float4 o = t.A[k];
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                float dx = x[s] - o.x, dy = y[s] - o.y, dz = z[s] - o.z;
                bool in = fabsf(dx) <= o.w && fabsf(dy) <= o.w && fabsf(dz) <= o.w;
                hit |= in ? (1u << s) : 0u;
                float d2 = dist2(dx, dy, dz);
                if (d2 < bound[s]) {
                    best[s] = fminf(best[s], fmaxf(0.0f, __fsub_rn(sqrt_rn(d2), o.w)));
                    bound[s] = prune_bound(best[s], smax);
                }
            }
        }
        int k = y_hi;
        // AABBs (extension): true box distance
        for (; k < cnt; ++k) {
            const float4 o = t.A[k];
            const float4 h = t.B[k];
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                float ax = fabsf(x[s] - o.x), ay = fabsf(y[s] - o.y), az = fabsf(z[s] - o.z);
                bool in = ax <= h.x && ay <= h.y && az <= h.z;
                hit |= in ? (1u << s) : 0u;
                float qx = fmaxf(__fsub_rn(ax, h.x), 0.0f), qy = fmaxf(__fsub_rn(ay, h.y), 0.0f),
                      qz = fmaxf(__fsub_rn(az, h.z), 0.0f);
                best[s] = fminf(best[s], sqrt_rn(dist2(qx, qy, qz)));
            }
        }
    }
}

// One obstacle (global table row `id`) against one point — the same arithmetic as the table loops above.
template <int SEM>
__device__ __forceinline__ void static_test(const EnvDev& e, int id, float x, float y, float z, bool& hit, float& best) {
    const float4 o = __ldg(e.tabA + id);
    const float dx = x - o.x, dy = y - o.y, dz = z - o.z;
    if constexpr (SEM == TF_SEM_HR) {
        const float d2 = dist2(dx, dy, dz);
        hit = hit || (d2 <= o.w);
        best = fminf(best, d2);
    } else {
        const int s_end = e.n_sphere, c_end = s_end + e.n_cube, y_end = c_end + e.n_cyl;
        if (id < s_end) {
            const float d2 = dist2(dx, dy, dz);
            hit = hit || (d2 <= __fmul_rn(o.w, o.w));
            best = fminf(best, fmaxf(0.0f, __fsub_rn(sqrt_rn(d2), o.w)));
        } else if (id < c_end) {
            hit = hit || (fabsf(dx) <= o.w && fabsf(dy) <= o.w && fabsf(dz) <= o.w);
            best = fminf(best, fmaxf(0.0f, __fsub_rn(sqrt_rn(dist2(dx, dy, dz)), o.w)));
        } else if (id < y_end) {
            const float h = __fmul_rn(2.0f, o.w), az = fabsf(dz);
            hit = hit || (fabsf(dx) <= o.w && fabsf(dy) <= o.w && az <= h);
            best = fminf(best, fmaxf(0.0f, __fsub_rn(az, h)));
        } else {
            const float4 h = __ldg(e.tabB + id);
            const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
            hit = hit || (ax <= h.x && ay <= h.y && az <= h.z);
            const float qx = fmaxf(__fsub_rn(ax, h.x), 0.0f), qy = fmaxf(__fsub_rn(ay, h.y), 0.0f),
                        qz = fmaxf(__fsub_rn(az, h.z), 0.0f);
            best = fminf(best, sqrt_rn(dist2(qx, qy, qz)));
        }
    }
}

// Static obstacles through the candidate index (tf_env_build_static_index): every point visits the
// list of its 3-D cell and (PF) of its z cell — supersets of "nearest" and "colliding", so min / or
// equal the full table loop bit for bit.  Outside the box / overflowed lists: the full loop.
template <int SEM, int NP>
__device__ __forceinline__ void eval_static_indexed(const EnvDev& e, const float (&x)[NP], const float (&y)[NP],
                                                    const float (&z)[NP], uint32_t& hit, float (&best)[NP]) {
#pragma unroll
    for (int s = 0; s < NP; ++s) {
        const int ix = __float2int_rd((x[s] - e.sox) * e.sinv_h), iy = __float2int_rd((y[s] - e.soy) * e.sinv_h),
                  iz = __float2int_rd((z[s] - e.soz) * e.sinv_h);
        bool ok = (unsigned)ix < (unsigned)e.sgx && (unsigned)iy < (unsigned)e.sgy && (unsigned)iz < (unsigned)e.sgz;
        int cnt = -1, zcnt = 0;
        size_t cell = 0;
        int zc = 0;
        if (ok) {
            cell = ((size_t)iz * e.sgy + iy) * e.sgx + ix;
            cnt = __ldg(e.scount + cell);
            if (SEM == TF_SEM_PF && e.n_cyl > 0) {
                zc = __float2int_rd((z[s] - e.zoz) * e.zinv_h);
                zcnt = (unsigned)zc < (unsigned)e.znz ? __ldg(e.zcount + zc) : -1;
            }
        }
        bool h = false;
        float b = best[s];
        if (cnt >= 0 && zcnt >= 0) {
            const int32_t* list = e.scand + cell * e.scap;
            for (int j = 0; j < cnt; ++j) static_test<SEM>(e, __ldg(list + j), x[s], y[s], z[s], h, b);
            if (SEM == TF_SEM_PF) {
                const int32_t* zl = e.zcand + (size_t)zc * e.zcap;
                for (int j = 0; j < zcnt; ++j) static_test<SEM>(e, __ldg(zl + j), x[s], y[s], z[s], h, b);
            }
        } else {
            for (int k = 0; k < e.n; ++k) static_test<SEM>(e, k, x[s], y[s], z[s], h, b);
        }
        best[s] = b;
        hit |= h ? (1u << s) : 0u;
    }
}

// Dispatcher used by the single-chunk kernels.
template <int SEM, int NP>
__device__ __forceinline__ void eval_static(const EnvDev& e, const Tables& t, const float (&x)[NP],
                                            const float (&y)[NP], const float (&z)[NP], uint32_t& hit,
                                            float (&best)[NP]) {
    if (e.scand != nullptr) eval_static_indexed<SEM, NP>(e, x, y, z, hit, best);
    else eval_static_chunk<SEM, NP>(e, t, 0, min(e.n, kTableCap), x, y, z, hit, best);
}

// Moving spheres (extension): HR rules at c(t) = c0 + v t.  Returns min d^2 per point.
// With a candidate index (tf_env_build_dynamic_index) a point at an integer time slice inside the
// index box only visits the spheres listed for its cell — a superset of "nearest" and of "colliding"
// by construction, so min / or (and every bit of the result) equal the brute-force loop.
__device__ __forceinline__ void dyn_test(const float4 o, const float4 v, float x, float y, float z, float tm,
                                         bool& hit, float& m2) {
    const float cx = fmaf(v.x, tm, o.x), cy = fmaf(v.y, tm, o.y), cz = fmaf(v.z, tm, o.z);
    const float d2 = dist2(x - cx, y - cy, z - cz);
    hit = hit || (d2 <= o.w);
    m2 = fminf(m2, d2);
}

// One listed sphere against one point: `rec` is the sphere's offset into the interleaved record table
// drec[2 k] = (c0, r2), drec[2 k + 1] = (v, 0) — 32 bytes, one sector — which is what the lists store.  The
// table (32 KB for 1024 spheres) stays in L1, and the lanes of a half-warp (consecutive waypoints, mostly the
// same cell) read the same record.  Same arithmetic as dyn_test.
__device__ __forceinline__ void dyn_test_rec(const float4* __restrict__ drec, int rec, float x, float y, float z, float tm,
                                             bool& hit, float& m2) {
    dyn_test(__ldg(drec + (uint32_t)rec), __ldg(drec + (uint32_t)rec + 1), x, y, z, tm, hit, m2);
}

#ifndef TF_DYN_CELL_FMA   // A/B: cell coordinates of the moving-sphere index by one fma per axis
#define TF_DYN_CELL_FMA 1
#endif
// Cell of the candidate index for a point at integer time ti, or -1 (outside the index box / time range).
__device__ __forceinline__ int dyn_cell(const EnvDev& e, float x, float y, float z, int ti) {
#if TF_DYN_CELL_FMA
    // one fma per axis.  It may put a point that sits within a few ulps of a cell wall into the neighbouring cell —
    // which lists a superset of its nearest / colliding spheres too (the 1e-3 margin of tf_env_build_dynamic_index)
    const int ix = __float2int_rd(fmaf(x, e.dinv_h, e.dnox)), iy = __float2int_rd(fmaf(y, e.dinv_h, e.dnoy)),
              iz = __float2int_rd(fmaf(z, e.dinv_h, e.dnoz));
#else
    const int ix = __float2int_rd((x - e.dox) * e.dinv_h), iy = __float2int_rd((y - e.doy) * e.dinv_h),
              iz = __float2int_rd((z - e.doz) * e.dinv_h);
#endif
    const bool indexed = ti >= 0 && ti < e.dT && (unsigned)ix < (unsigned)e.dgx && (unsigned)iy < (unsigned)e.dgy &&
                         (unsigned)iz < (unsigned)e.dgz;
    if (!indexed) return -1;
    const uint32_t tb = ((uint32_t)ti * e.dband_magic) >> 20;   // == ti / dband for every ti < dT (checked at build time)
    // < 2^26 lists (tf_env_build_dynamic_index)
    return (int)(((tb * (uint32_t)e.dgz + (uint32_t)iz) * (uint32_t)e.dgy + (uint32_t)iy) * (uint32_t)e.dgx + (uint32_t)ix);
}

// A point outside the index (or in an overflowed cell): every sphere, from global memory.  Rare by construction
// (tf_env_build_dynamic_index asks for a generous box): a rolled loop, so that the hot kernels stay small.
__device__ __forceinline__ void dyn_brute_force(const float4* __restrict__ dynA, const float4* __restrict__ dynB, int n, float x,
                                             float y, float z, float tm, bool& hit, float& m2) {
    bool h = hit;
    float m = m2;
#pragma unroll 1
    for (int k = 0; k < n; ++k) dyn_test(__ldg(dynA + k), __ldg(dynB + k), x, y, z, tm, h, m);
    hit = h;
    m2 = m;
}

// Four listed spheres (two words of a list) against one point: a group is walked without per-entry predicates and its
// loads go out together (the last group of a list is filled up with copies of the list's first entry; min / or are
// idempotent).
__device__ __forceinline__ void dyn_group(const float4* __restrict__ drec, uint32_t w0, uint32_t w1, float x, float y, float z,
                                          float tm, bool& h, float& m2) {
    dyn_test_rec(drec, (int)(w0 & 0xffffu), x, y, z, tm, h, m2);
    dyn_test_rec(drec, (int)(w0 >> 16), x, y, z, tm, h, m2);
    dyn_test_rec(drec, (int)(w1 & 0xffffu), x, y, z, tm, h, m2);
    dyn_test_rec(drec, (int)(w1 >> 16), x, y, z, tm, h, m2);
}
// The indexed walk in three phases, so that the common case is ONE straight-line block:
// (1) dyn_heads: the cells and the 8-byte heads of the NP lists (length + three words, EnvDev::dcand) — NP independent
//     loads in flight;
// (2) dyn_eval_heads: three tests per list, unconditionally: the entries of a list of up to three spheres (an indexed
//     list holds at least the nearest sphere, and its head is filled up with copies of the first entry), the first entry
//     three times for a longer list, sphere 0 three times for a point outside the index / an overflowed list (head
//     (0xffff, 0); the loop over all spheres repeats it — min / or are idempotent);
// (3) behind one branch, the lists longer than three entries (from dext, in groups of four) and the loop over all spheres.
template <int NP>
struct DynHeads {
    uint2 hd[NP];
};
template <int NP>
__device__ __forceinline__ void dyn_heads(const EnvDev& e, const float (&x)[NP], const float (&y)[NP], const float (&z)[NP],
                                          int t0, int dt, DynHeads<NP>& d) {
#pragma unroll
    for (int s = 0; s < NP; ++s) {
        const int cell = dyn_cell(e, x[s], y[s], z[s], t0 + dt * s);
        d.hd[s] = make_uint2(0xffffu, 0u);
        if (cell >= 0) d.hd[s] = __ldg(e.dcand + (uint32_t)cell);
    }
}
template <int NP>
__device__ __forceinline__ void dyn_eval_heads(const EnvDev& e, const float (&x)[NP], const float (&y)[NP], const float (&z)[NP],
                                               int t0, int dt, const DynHeads<NP>& d, uint32_t& hit, float (&m2)[NP]) {
    float tm[NP];
#pragma unroll
    for (int s = 0; s < NP; ++s) tm[s] = (float)(t0 + dt * s);
    bool rare = false;
#pragma unroll
    for (int s = 0; s < NP; ++s) {
        const uint32_t n = d.hd[s].x & 0xffffu, e1 = d.hd[s].x >> 16;
        const uint32_t w23 = n > 3u ? e1 * 0x10001u : d.hd[s].y;   // words 2..3 of a longer list are its slot in dext
        bool h = false;
        dyn_test_rec(e.drec, (int)e1, x[s], y[s], z[s], tm[s], h, m2[s]);
        dyn_test_rec(e.drec, (int)(w23 & 0xffffu), x[s], y[s], z[s], tm[s], h, m2[s]);
        dyn_test_rec(e.drec, (int)(w23 >> 16), x[s], y[s], z[s], tm[s], h, m2[s]);
        hit |= h ? (1u << s) : 0u;
        rare = rare || n > 3u;
    }
    if (rare) {
#pragma unroll
        for (int s = 0; s < NP; ++s) {
            const int n = (int)(d.hd[s].x & 0xffffu);
            if (n <= 3) continue;
            bool h = false;
            if (n != 0xffff) {
                const uint2* list = reinterpret_cast<const uint2*>(e.dext + (size_t)d.hd[s].y * (uint32_t)e.dcap);
#pragma unroll 1
                for (int j = 0; j < n; j += 4) {
                    const uint2 q = __ldg(list++);
                    dyn_group(e.drec, q.x, q.y, x[s], y[s], z[s], tm[s], h, m2[s]);
                }
            } else {
                dyn_brute_force(e.dynA, e.dynB, e.n_dyn, x[s], y[s], z[s], tm[s], h, m2[s]);
            }
            hit |= h ? (1u << s) : 0u;
        }
    }
}

// Moving spheres against NP points whose times are the INTEGERS t0, t0 + dt, ... (waypoint indices).
// Without a candidate index: brute force over the records staged in shared memory.  With one: the phased walk above
// (a chain of two dependent loads per point, list head -> records; the NP heads go out together).
template <int NP>
__device__ __forceinline__ void eval_dynamic(const EnvDev& e, const Tables& t, const float (&x)[NP],
                                             const float (&y)[NP], const float (&z)[NP], int t0, int dt,
                                             uint32_t& hit, float (&m2)[NP]) {
    if (e.dcand == nullptr) {
        for (int k = 0; k < e.n_dyn; ++k) {
            const float4 o = t.dA[k];
            const float4 v = t.dB[k];
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                bool h = false;
                dyn_test(o, v, x[s], y[s], z[s], (float)(t0 + dt * s), h, m2[s]);
                hit |= h ? (1u << s) : 0u;
            }
        }
        return;
    }
    DynHeads<NP> d;
    dyn_heads<NP>(e, x, y, z, t0, dt, d);
    dyn_eval_heads<NP>(e, x, y, z, t0, dt, d, hit, m2);
}

// One point at an arbitrary (fractional) time — the sample points of the S > 1 extension: brute force unless the
// time happens to be an integer.
__device__ __forceinline__ void eval_dynamic_point(const EnvDev& e, const Tables& t, float x, float y, float z, float tm,
                                                   bool& hit, float& m2) {
    const int ti = (int)tm;
    if (e.dcand == nullptr || (float)ti == tm) {
        const float xs[1] = {x}, ys[1] = {y}, zs[1] = {z};
        float m[1] = {m2};
        uint32_t hb = 0;
        if (e.dcand == nullptr) {   // staged records: dyn_test at the exact time
            for (int k = 0; k < e.n_dyn; ++k) {
                bool h = false;
                dyn_test(t.dA[k], t.dB[k], x, y, z, tm, h, m[0]);
                hb |= h ? 1u : 0u;
            }
        } else {
            eval_dynamic<1>(e, t, xs, ys, zs, ti, 0, hb, m);
        }
        hit = hit || hb != 0u;
        m2 = m[0];
        return;
    }
    for (int k = 0; k < e.n_dyn; ++k) dyn_test(__ldg(e.dynA + k), __ldg(e.dynB + k), x, y, z, tm, hit, m2);
}

// Bit-packed occupancy probe through L2 (brick layout, include/tracefly_b200.h).
__device__ __forceinline__ bool voxel_hit(const EnvDev& e, float x, float y, float z) {
    int ix = __float2int_rd(x), iy = __float2int_rd(y), iz = __float2int_rd(z);
    bool inb = (unsigned)ix < (unsigned)e.nx && (unsigned)iy < (unsigned)e.ny && (unsigned)iz < (unsigned)e.nz;
    if (!inb) return false;
    size_t brick = (size_t)(ix >> 3) + (size_t)e.nbx * ((size_t)(iy >> 3) + (size_t)e.nby * (size_t)(iz >> 3));
    uint32_t w = __ldg(e.bricks + brick * 16 + ((iz & 7) * 2 + ((iy & 7) >> 2)));
    return (w >> (((iy & 3) << 3) | (ix & 7))) & 1u;
}

// The same probe split in two so that the gather can be issued one chromosome ahead of its use:
// voxel_prefetch issues the loads (word + single-bit mask per slot), the consumer tests word & mask.
template <int NW>
struct VoxRegs {
    uint32_t word[NW], mask[NW];
};
template <int NW, bool HINT = true>
__device__ __forceinline__ void voxel_prefetch(const EnvDev& e, const Lanes<NW>& p, VoxRegs<NW>& v, uint64_t keep_policy) {
#pragma unroll
    for (int s = 0; s < NW; ++s) {
        const int ix = __float2int_rd(p.x[s]), iy = __float2int_rd(p.y[s]), iz = __float2int_rd(p.z[s]);
        const bool inb = (unsigned)ix < (unsigned)e.nx && (unsigned)iy < (unsigned)e.ny && (unsigned)iz < (unsigned)e.nz;
        const uint32_t brick = (uint32_t)(ix >> 3) + (uint32_t)e.nbx * ((uint32_t)(iy >> 3) + (uint32_t)e.nby * (uint32_t)(iz >> 3));
        v.mask[s] = inb ? (1u << (((iy & 3) << 3) | (ix & 7))) : 0u;
        const uint32_t* wp = e.bricks + (size_t)brick * 16 + ((iz & 7) * 2 + ((iy & 7) >> 2));
        if constexpr (HINT) v.word[s] = inb ? ldg_u32_hint(wp, keep_policy) : 0u;
        else v.word[s] = inb ? __ldg(wp) : 0u;
    }
}

__device__ __forceinline__ float lerp_f(float a, float b, float f) { return fmaf(f, __fsub_rn(b, a), a); }

// Tri-linear hazard lookup (extension; the unused wind_data hook of PF:144).
__device__ __forceinline__ float hazard_sample(const EnvDev& e, float x, float y, float t) {
    float gx = fmaf(x, e.sx, e.ox), gy = fmaf(y, e.sy, e.oy), gt = __fmul_rn(t, e.st);
    gx = fminf(fmaxf(gx, 0.0f), (float)(e.hnx - 1));
    gy = fminf(fmaxf(gy, 0.0f), (float)(e.hny - 1));
    gt = fminf(fmaxf(gt, 0.0f), (float)(e.nt - 1));
    const int x0 = (int)gx, y0 = (int)gy, t0 = (int)gt;
    const int y1 = min(y0 + 1, e.hny - 1);
    const float fx = __fsub_rn(gx, (float)x0), fy = __fsub_rn(gy, (float)y0), ft = __fsub_rn(gt, (float)t0);
    // the env keeps the map corner-packed in [y][x][t] order: one float4 holds the (x0, x1) x (t0, t1) taps of
    // a row y (clamping of x1 / t1 folded in at pack time; x1 as its difference to x0), so the 8 taps are 2 loads; the lanes of a warp are
    // consecutive waypoints, i.e. nearly the same (x, y) at consecutive times: their taps share cache lines.
    // 32-bit element indices (the packed map has < 2^29 elements)
    const uint32_t col = (uint32_t)(x0 * e.nt + t0), rowlen = (uint32_t)(e.hnx * e.nt);
    const float4 a = __ldg(e.hazard + ((uint32_t)y0 * rowlen + col));
    const float4 b = __ldg(e.hazard + ((uint32_t)y1 * rowlen + col));
    // .y / .w hold the x differences (v[x1] - v[x0], subtracted in FP32 at pack time): lerp_f(v0, v1, fx) = fmaf(fx, v1 - v0, v0)
    const float a0 = lerp_f(fmaf(fx, a.y, a.x), fmaf(fx, b.y, b.x), fy);
    const float a1 = lerp_f(fmaf(fx, a.w, a.z), fmaf(fx, b.w, b.z), fy);
    return lerp_f(a0, a1, ft);
}

// Full collision / nearest-distance evaluation of ONE point against a fully staged table
// (sample points of the S > 1 extension; requires a single chunk).
template <int SEM, bool VOX, int XF>
__device__ __forceinline__ void eval_point_full(const EnvDev& e, const Tables& t, float x, float y, float z, float tm,
                                                bool& hit_out, float& dnear) {
    float xs[1] = {x}, ys[1] = {y}, zs[1] = {z};
    uint32_t hit = 0;
    float best[1] = {INFINITY};
    eval_static<SEM, 1>(e, t, xs, ys, zs, hit, best);
    float dn = (SEM == TF_SEM_HR) ? sqrt_rn(best[0]) : best[0];
    if constexpr (XF >= 1) {
        if (e.n_dyn > 0) {
            float m2 = INFINITY;
            bool dh = false;
            eval_dynamic_point(e, t, x, y, z, tm, dh, m2);
            hit |= dh ? 1u : 0u;
            dn = fminf(dn, sqrt_rn(m2));
        }
    }
    if constexpr (VOX) hit |= voxel_hit(e, x, y, z) ? 1u : 0u;
    hit_out = hit != 0;
    dnear = dn;
}

// Stage `cnt` table rows starting at global row `row0` into shared memory with the TMA engine.
__device__ __forceinline__ void stage_chunk(const EnvDev& e, float4* sA, float4* sB, int row0, int cnt,
                                            uint64_t* bar, bool with_dyn, float4* sdA, float4* sdB) {
    // one elected thread arms the barrier with the byte count and issues the bulk copies
    uint32_t bytes = (uint32_t)cnt * 16u * (e.n_aabb > 0 ? 2u : 1u);
    if (with_dyn) bytes += (uint32_t)e.n_dyn_smem * 32u;
    mbar_expect_tx(bar, bytes);
    if (cnt > 0) {
        tma_load_1d(sA, e.tabA + row0, (uint32_t)cnt * 16u, bar);
        if (e.n_aabb > 0) tma_load_1d(sB, e.tabB + row0, (uint32_t)cnt * 16u, bar);
    }
    if (with_dyn) {
        tma_load_1d(sdA, e.dynA, (uint32_t)e.n_dyn_smem * 16u, bar);
        tma_load_1d(sdB, e.dynB, (uint32_t)e.n_dyn_smem * 16u, bar);
    }
}

struct SmemLayout {
    uint64_t* bar;
    float4 *A, *B, *dA, *dB;
};
__host__ __device__ inline size_t smem_bytes_for(const EnvDev& e) {
    size_t cap = (size_t)(e.n_smem < kTableCap ? e.n_smem : kTableCap);
    return 16 + cap * 16 * (e.n_aabb > 0 ? 2 : 1) + (size_t)e.n_dyn_smem * 32;
}
__device__ __forceinline__ SmemLayout carve_smem(const EnvDev& e, unsigned char* smem) {
    SmemLayout s;
    size_t cap = (size_t)min(e.n_smem, kTableCap);
    s.bar = reinterpret_cast<uint64_t*>(smem);
    s.A = reinterpret_cast<float4*>(smem + 16);
    s.B = s.A + (e.n_aabb > 0 ? cap : 0);
    s.dA = s.B + cap;
    if (e.n_aabb == 0) s.dA = s.A + cap;
    s.dB = s.dA + e.n_dyn_smem;
    return s;
}

// Per-warp evaluation state shared by tf_fitness and the fused tf_generation kernel.
//   cur      : this lane's waypoints (slot s = waypoint lane + 32 s), already in registers
//   L        : valid waypoint count of the chromosome
// The CTA-uniform chunk loop lives in the caller when the table needs more than one chunk;
// this struct holds the accumulators across chunks.
template <int NW, int SEM>
struct PointAcc {
    uint32_t hit;
    float best[NW];
    __device__ __forceinline__ void reset() {
        hit = 0;
#pragma unroll
        for (int s = 0; s < NW; ++s) best[s] = INFINITY;
    }
};

// ---- tile structure -----------------------------------------------------------------------
// A warp works on a TILE of 32 chromosomes.  Phase B (lane = waypoint) walks the tile one
// chromosome at a time and leaves that chromosome's warp-reduced sums in lane c of a TileSums;
// phase C (lane = chromosome) then does the final score arithmetic (5 IEEE divisions) for all 32
// at once and writes fitness / collision words coalesced.  Per-chromosome scalar work is thereby
// paid once per 32 chromosomes instead of once per warp (ncu round 1: 842 warp-instr/chromosome
// when every lane redundantly did the scalar work).
template <int NW>
struct TileSums {
    float prox, dist, alt, haz;
    int packed;            // turns | ncoll << 10
    uint32_t cw[NW];       // collision words
    __device__ __forceinline__ void reset() {
        prox = dist = alt = haz = 0.0f;
        packed = 0;
#pragma unroll
        for (int s = 0; s < NW; ++s) cw[s] = 0;
    }
};

// Chromosomes per warp tile: 32 when the population fills the machine, fewer for small P so
// that there are still enough tiles to occupy every SM (lanes >= T idle in phases A / C).
inline int tile_size_for(int P) {
    const long long target_warps = (long long)sm_count() * 48;
    int T = 32;
    while (T > 4 && (long long)P / T < target_warps) T >>= 1;
    return T;
}

// Phase B for one chromosome (all 32 lanes cooperate; `c` = its index in the tile).
template <int NW, int SEM, bool VOX, int XF>
__device__ __forceinline__ void reduce_chromosome(const EnvDev& e, const tf_params& prm, const Tables& t,
                                                  const Lanes<NW>& cur, int L, PointAcc<NW, SEM>& acc, int c,
                                                  TileSums<NW>& out) {
    const int lane = threadIdx.x & 31;
    float a_prox = 0.0f, a_dist = 0.0f, a_alt = 0.0f, a_haz = 0.0f;
    int turns = 0, ncoll = 0;
    uint32_t hit = acc.hit;
    // without any distance source d_near = inf and every proximity term is exactly 0 (HR:26-27)
    const bool has_dist = (SEM != TF_SEM_NONE) || (XF >= 1 && e.n_dyn > 0);
    float dn[NW];
#pragma unroll
    for (int s = 0; s < NW; ++s) dn[s] = (SEM == TF_SEM_HR) ? sqrt_rn(acc.best[s]) : acc.best[s];
    if constexpr (XF >= 1) {
        if (e.n_dyn > 0) {
            float m2[NW];
#pragma unroll
            for (int s = 0; s < NW; ++s) m2[s] = INFINITY;
            eval_dynamic<NW>(e, t, cur.x, cur.y, cur.z, lane, 32, hit, m2);
#pragma unroll
            for (int s = 0; s < NW; ++s) dn[s] = fminf(dn[s], sqrt_rn(m2[s]));
        }
    }
    const int S = (XF >= 2) ? max(prm.samples_per_segment, 1) : 1;
    // previous waypoint of each slot: rotate by one lane; lane 0 takes lane 31 of the previous slot
    float rx_prev = 0.f, ry_prev = 0.f, rz_prev = 0.f;
#pragma unroll
    for (int s = 0; s < NW; ++s) {
        const int k = lane + 32 * s;
        const bool valid = k < L;
        const int src = (lane + 31) & 31;
        float rx = __shfl_sync(0xffffffffu, cur.x[s], src);
        float ry = __shfl_sync(0xffffffffu, cur.y[s], src);
        float rz = __shfl_sync(0xffffffffu, cur.z[s], src);
        float px = lane == 0 ? rx_prev : rx, py = lane == 0 ? ry_prev : ry, pz = lane == 0 ? rz_prev : rz;
        rx_prev = rx; ry_prev = ry; rz_prev = rz;
        bool h = (hit >> s) & 1u;
        if constexpr (VOX) h = h || voxel_hit(e, cur.x[s], cur.y[s], cur.z[s]);
        h = h && valid;
        const uint32_t word = __ballot_sync(0xffffffffu, h);
        if (lane == c) out.cw[s] = word;
        ncoll += h ? 1 : 0;
        if (has_dist && valid) a_prox = __fadd_rn(a_prox, rcp_rn(__fadd_rn(1.0f, dn[s])));  // HR:46
        if constexpr (XF >= 1) {
            if (e.hazard != nullptr && valid) a_haz = __fadd_rn(a_haz, hazard_sample(e, cur.x[s], cur.y[s], (float)k));
        }
        if (valid && k >= 1) {
            float dx = cur.x[s] - px, dy = cur.y[s] - py, dz = cur.z[s] - pz;
            a_dist = __fadd_rn(a_dist, sqrt_rn(dist2(dx, dy, dz)));  // HR:51-52
            a_alt = __fadd_rn(a_alt, fabsf(dz));                         // HR:44
            if (k <= L - 2) turns += (cur.x[s] != px || cur.y[s] != py) ? 1 : 0;  // HR:38-40,43
            if constexpr (XF >= 2) {
                for (int j = 1; j < S; ++j) {  // extension: samples inside segment (k-1, k)
                    float tj = __fdiv_rn((float)j, (float)S);
                    bool sh;
                    float sdn;
                    eval_point_full<SEM, VOX, XF>(e, t, fmaf(dx, tj, px), fmaf(dy, tj, py), fmaf(dz, tj, pz),
                                                     __fadd_rn((float)(k - 1), tj), sh, sdn);
                    ncoll += sh ? 1 : 0;
                    if (has_dist) a_prox = __fadd_rn(a_prox, rcp_rn(__fadd_rn(1.0f, sdn)));
                }
            }
        }
    }
    const float dist = warp_sum_xor(a_dist), alt = warp_sum_xor(a_alt);
    const int packed = __reduce_add_sync(0xffffffffu, turns | (ncoll << 10));
    if (lane == c) { out.dist = dist; out.alt = alt; out.packed = packed; }
    if (has_dist) {
        const float prox = warp_sum_xor(a_prox);
        if (lane == c) out.prox = prox;
    }
    if constexpr (XF >= 1) {
        if (e.hazard != nullptr) {
            const float haz = warp_sum_xor(a_haz);
            if (lane == c) out.haz = haz;
        }
    }
}

// Phase C: the score arithmetic of HR:45-59 for this lane's chromosome.
template <int NW, int XF>
__device__ __forceinline__ float combine_fitness(const EnvDev& e, const tf_params& prm, const TileSums<NW>& s) {
    const float tf_ = (float)(s.packed & 1023);
    const float pen = __fadd_rn((float)(s.packed >> 10), s.prox);            // HR:45-46
    const float safety = __fdiv_rn(1.0f, __fadd_rn(1.0f, pen));              // HR:47
    const float energy = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(1.0f, tf_), s.alt));  // HR:48
    float windden = __fadd_rn(1.0f, s.alt);                                  // HR:49
    if constexpr (XF >= 1) {
        if (e.hazard != nullptr) windden = fmaf(prm.hazard_weight, s.haz, windden);
    }
    const float wind = __fdiv_rn(1.0f, windden);
    const float time_s = (s.dist != 0.0f) ? __fdiv_rn(1.0f, s.dist) : (prm.dist_zero_guard ? 0.0f : INFINITY);  // HR:53 / PF:163
    const float ts = __fdiv_rn(1.0f, __fadd_rn(1.0f, tf_));
    float f = __fmul_rn(prm.w_safety, safety);                               // HR:55-59
    f = fmaf(prm.w_energy, energy, f);
    f = fmaf(prm.w_wind, wind, f);
    f = fmaf(prm.w_time, time_s, f);
    f = fmaf(prm.w_turn, ts, f);
    return f;
}

}  // namespace tf
