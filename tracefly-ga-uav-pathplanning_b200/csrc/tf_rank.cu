// tf_rank.cu — stable descending top-k (replaces population.sort(key=fitness, reverse=True),
// HR:114 / PF:193; only the head of the ranking is consumed: HR:115,118 top-2 / top-10).
//
// Order = (fitness desc, index asc) — exactly what a STABLE descending sort gives.  Each CTA
// bitonic-sorts a tile of composite keys in shared memory and keeps its first k; levels repeat
// on the survivors until one tile is left.  `worst` selects the tail of the same order.
#include "tf_common.cuh"

namespace tf {

struct Key32 {  // fitness f32: one u64 = (~orderable(f) << 32) | idx
    unsigned long long v;
    __device__ __forceinline__ static Key32 make(float f, uint32_t idx, bool worst) {
        if (f == 0.0f) f = 0.0f;  // -0.0 == +0.0 for Python's comparison
        uint32_t u = __float_as_uint(f);
        u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;  // ascending u <=> ascending f
        unsigned long long k = ((unsigned long long)(~u) << 32) | idx;
        return Key32{worst ? ~k : k};
    }
    __device__ __forceinline__ static Key32 pad() { return Key32{~0ull}; }
    __device__ __forceinline__ uint32_t index(bool worst) const { return (uint32_t)(worst ? ~v : v); }
    __device__ __forceinline__ bool operator<(const Key32& o) const { return v < o.v; }
};

struct Key64 {  // fitness f64: (~orderable(f), idx)
    unsigned long long hi;
    unsigned long long lo;
    __device__ __forceinline__ static Key64 make(double f, uint32_t idx, bool worst) {
        if (f == 0.0) f = 0.0;
        unsigned long long u = (unsigned long long)__double_as_longlong(f);
        u ^= (u >> 63) ? ~0ull : 0x8000000000000000ull;
        Key64 k{~u, (unsigned long long)idx};
        if (worst) { k.hi = ~k.hi; k.lo = ~k.lo; }
        return k;
    }
    __device__ __forceinline__ static Key64 pad() { return Key64{~0ull, ~0ull}; }
    __device__ __forceinline__ uint32_t index(bool worst) const { return (uint32_t)(worst ? ~lo : lo); }
    __device__ __forceinline__ bool operator<(const Key64& o) const { return hi < o.hi || (hi == o.hi && lo < o.lo); }
};

constexpr int kRankThreads = 256;

// values != nullptr : level 0, build keys from fitness values; else keys_in holds survivors.
template <typename Key, typename T>
__global__ void __launch_bounds__(kRankThreads)
topk_tile_kernel(const T* __restrict__ values, const Key* __restrict__ keys_in, int n, int tile, int k, int worst,
                 Key* __restrict__ keys_out, int32_t* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key* s = reinterpret_cast<Key*>(smem_raw);
    const int base = blockIdx.x * tile;
    for (int i = threadIdx.x; i < tile; i += kRankThreads) {
        const int g = base + i;
        Key key = Key::pad();
        if (g < n) key = values ? Key::make(values[g], (uint32_t)g, worst != 0) : keys_in[g];
        s[i] = key;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= tile; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (tile >> 1); i += kRankThreads) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool up = (lo & k2) == 0;
                Key a = s[lo], b = s[hi];
                if ((b < a) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    const int kk = min(k, tile);
    if (idx_out) {
        for (int i = threadIdx.x; i < k; i += kRankThreads) idx_out[i] = i < kk ? (int32_t)s[i].index(worst != 0) : -1;
    } else {
        for (int i = threadIdx.x; i < kk; i += kRankThreads) keys_out[(size_t)blockIdx.x * k + i] = s[i];
    }
}

static int tile_for(int k) {
    int t = 2048;
    while (t < 2 * k) t <<= 1;
    return t;
}

template <typename Key>
static int64_t workspace_bytes(int P, int k) {
    const int tile = tile_for(k);
    int64_t n1 = (int64_t)((P + tile - 1) / tile) * k;
    int64_t n2 = ((n1 + tile - 1) / tile) * k;
    return (n1 + n2) * (int64_t)sizeof(Key) + 256;
}

template <typename Key, typename T>
static int run_topk(const T* values, int P, int k, int worst, int32_t* idx, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if (!values || !idx) return TF_ERR_INVALID;
    if (P < 1 || k < 1 || k > TF_TOPK_MAX || k > P) return TF_ERR_SHAPE;
    const int tile = tile_for(k);
    const size_t smem = (size_t)tile * sizeof(Key);
    if (smem > 200 * 1024) return TF_ERR_SHAPE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    auto kern = topk_tile_kernel<Key, T>;
    if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nblocks = (P + tile - 1) / tile;
    if (nblocks == 1) {
        kern<<<1, kRankThreads, smem, st>>>(values, nullptr, P, tile, k, worst, nullptr, idx);
        TF_LAUNCH_CHECK();
        return TF_OK;
    }
    if (!ws || ws_bytes < workspace_bytes<Key>(P, k)) return TF_ERR_WORKSPACE;
    Key* bufA = reinterpret_cast<Key*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    int64_t n1 = (int64_t)nblocks * k;
    Key* bufB = bufA + n1;
    kern<<<nblocks, kRankThreads, smem, st>>>(values, nullptr, P, tile, k, worst, bufA, nullptr);
    TF_LAUNCH_CHECK();
    int n = (int)n1;
    Key *in = bufA, *out = bufB;
    while (true) {
        nblocks = (n + tile - 1) / tile;
        if (nblocks == 1) {
            kern<<<1, kRankThreads, smem, st>>>(nullptr, in, n, tile, k, worst, nullptr, idx);
            TF_LAUNCH_CHECK();
            return TF_OK;
        }
        kern<<<nblocks, kRankThreads, smem, st>>>(nullptr, in, n, tile, k, worst, out, nullptr);
        TF_LAUNCH_CHECK();
        n = nblocks * k;
        Key* t = in; in = out; out = t;
    }
}

}  // namespace tf

using namespace tf;

extern "C" int64_t tf_topk_workspace_bytes(int P, int k, int f64) {
    if (P < 1 || k < 1 || k > TF_TOPK_MAX) return TF_ERR_SHAPE;
    return f64 ? workspace_bytes<Key64>(P, k) : workspace_bytes<Key32>(P, k);
}
extern "C" int tf_rank_topk(const float* fit_dev, int P, int k, int worst, int32_t* idx_dev, void* workspace_dev,
                            int64_t workspace_bytes, void* stream) {
    return run_topk<Key32, float>(fit_dev, P, k, worst, idx_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}
extern "C" int tf_rank_topk_f64(const double* fit64_dev, int P, int k, int worst, int32_t* idx_dev,
                                void* workspace_dev, int64_t workspace_bytes, void* stream) {
    return run_topk<Key64, double>(fit64_dev, P, k, worst, idx_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}
