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
        const unsigned long long k = rank_key32(f, idx);  // -0.0 == +0.0 for Python's comparison
        return Key32{worst ? ~k : k};
    }
    __device__ __forceinline__ static Key32 pad() { return Key32{~0ull}; }
    __device__ __forceinline__ uint32_t index(bool worst) const { return (uint32_t)(worst ? ~v : v); }
    __device__ __forceinline__ bool operator<(const Key32& o) const { return v < o.v; }
    __device__ __forceinline__ bool operator==(const Key32& o) const { return v == o.v; }
    __device__ __forceinline__ bool is_pad() const { return v == ~0ull; }
    __device__ __forceinline__ void shfl_xor(int m) { v = __shfl_xor_sync(0xffffffffu, v, m); }
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
    __device__ __forceinline__ bool operator==(const Key64& o) const { return hi == o.hi && lo == o.lo; }
    __device__ __forceinline__ bool is_pad() const { return hi == ~0ull && lo == ~0ull; }
    __device__ __forceinline__ void shfl_xor(int m) {
        hi = __shfl_xor_sync(0xffffffffu, hi, m);
        lo = __shfl_xor_sync(0xffffffffu, lo, m);
    }
};

constexpr int kRankThreads = 256;

// values != nullptr : level 0, build keys from fitness values; else keys_in holds survivors.
template <typename Key, typename T>
__global__ void __launch_bounds__(1024)
topk_tile_kernel(const T* __restrict__ values, const Key* __restrict__ keys_in, int n, int tile, int k, int worst,
                 Key* __restrict__ keys_out, int32_t* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key* s = reinterpret_cast<Key*>(smem_raw);
    const int base = blockIdx.x * tile;
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
        const int g = base + i;
        Key key = Key::pad();
        if (g < n) key = values ? Key::make(values[g], (uint32_t)g, worst != 0) : keys_in[g];
        s[i] = key;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= tile; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (tile >> 1); i += blockDim.x) {
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
        for (int i = threadIdx.x; i < k; i += blockDim.x) idx_out[i] = i < kk ? (int32_t)s[i].index(worst != 0) : -1;
    } else {
        for (int i = threadIdx.x; i < kk; i += blockDim.x) keys_out[(size_t)blockIdx.x * k + i] = s[i];
    }
}

// ------------------------------------------------------------------------------------------
// Small k (the per-generation ranking: HR:115 top-2, HR:118 top-10): k rounds of a block-wide
// arg-min over keys held in registers (8 per thread, 2048 per block) instead of a full sort.
constexpr int kSmallItems = 8;
constexpr int kSmallTile = kRankThreads * kSmallItems;
constexpr int kSmallKMax = 16;

template <typename Key>
__device__ __forceinline__ Key load_key_cg(const Key* p) {  // L1-bypassing read of a key another CTA wrote
    Key k;
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(p);
    unsigned long long* o = reinterpret_cast<unsigned long long*>(&k);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(Key) / 8); ++i) o[i] = __ldcg(w + i);
    return k;
}

// k rounds of block-wide arg-min over the keys in `mine`; round r's winner goes to idx_out[r] or keys_out[r]
template <typename Key>
__device__ __forceinline__ void small_rounds(Key (&mine)[kSmallItems], int k, int worst, Key* keys_out, int32_t* idx_out,
                                             Key* wbest, Key* winner) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < k; ++r) {
        Key best = mine[0];
#pragma unroll
        for (int i = 1; i < kSmallItems; ++i)
            if (mine[i] < best) best = mine[i];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            Key o = best;
            o.shfl_xor(m);
            if (o < best) best = o;
        }
        if (lane == 0) wbest[warp] = best;
        __syncthreads();
        if (warp == 0) {
            Key b = lane < kRankThreads / 32 ? wbest[lane] : Key::pad();
#pragma unroll
            for (int m = 4; m >= 1; m >>= 1) {
                Key o = b;
                o.shfl_xor(m);
                if (o < b) b = o;
            }
            if (lane == 0) {
                *winner = b;
                if (idx_out) idx_out[r] = b.is_pad() ? -1 : (int32_t)b.index(worst != 0);
                else keys_out[r] = b;
            }
        }
        __syncthreads();
        const Key w = *winner;
#pragma unroll
        for (int i = 0; i < kSmallItems; ++i)
            if (mine[i] == w) mine[i] = Key::pad();  // keys are unique: exactly one owner retires it
    }
}

// `done` != nullptr: single-launch form.  Every CTA leaves its k winners in keys_out, and the CTA that arrives
// last at the counter (gridDim.x * k <= kSmallTile keys) runs the final rounds itself and zeroes the counter.
template <typename Key, typename T>
__global__ void __launch_bounds__(kRankThreads)
topk_small_kernel(const T* __restrict__ values, const Key* __restrict__ keys_in, int n, int k, int worst,
                  Key* __restrict__ keys_out, int32_t* __restrict__ idx_out, int* __restrict__ done) {
    __shared__ Key wbest[kRankThreads / 32];
    __shared__ Key winner;
    __shared__ int last;
    Key mine[kSmallItems];
    const int base = blockIdx.x * kSmallTile;
#pragma unroll
    for (int i = 0; i < kSmallItems; ++i) {
        const int g = base + i * kRankThreads + threadIdx.x;
        mine[i] = Key::pad();
        if (g < n) mine[i] = values ? Key::make(values[g], (uint32_t)g, worst != 0) : keys_in[g];
    }
    const bool partial = done != nullptr || idx_out == nullptr;
    small_rounds(mine, k, worst, partial ? keys_out + (size_t)blockIdx.x * k : nullptr, partial ? nullptr : idx_out, wbest, &winner);
    if (done == nullptr) return;
    if (threadIdx.x == 0) {  // thread 0 wrote this CTA's winners
        __threadfence();
        last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int n2 = (int)gridDim.x * k;
#pragma unroll
    for (int i = 0; i < kSmallItems; ++i) {
        const int g = i * kRankThreads + threadIdx.x;
        mine[i] = g < n2 ? load_key_cg(keys_out + g) : Key::pad();
    }
    small_rounds(mine, k, worst, (Key*)nullptr, idx_out, wbest, &winner);
    if (threadIdx.x == 0) *done = 0;
}

static int tile_for(int k) {
    int t = 2048;
    while (t < 2 * k) t <<= 1;
    return t;
}

// ------------------------------------------------------------------------------------------
// Large k (elite migration: E = 1024 elites, (world-1)*E worst): MSD radix SELECT on the 64-bit
// composite key.  The keys are unique (the index is the low word), so six digit passes pin the
// exact k-th key K; one compaction pass collects the k keys <= K and a single tile sort orders
// them.  ~15 short launches that each read 4 B / chromosome, independent of k.
constexpr int kSelPasses = 6;
constexpr int kSelBins = 2048;
struct SelState {
    unsigned long long prefix;  // digits chosen so far (high bits of K)
    int need;                   // rank of K inside the current prefix bucket (1-based)
    unsigned int count;         // compaction cursor
    unsigned int done;          // CTAs that have added their histogram of the current pass
};
__host__ __device__ __forceinline__ int sel_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
__host__ __device__ __forceinline__ int sel_width(int pass) { return pass < 5 ? 11 : 9; }

__global__ void sel_init_kernel(SelState* st, unsigned int* hist, int k) {
    for (int i = threadIdx.x; i < kSelPasses * kSelBins; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) { st->prefix = 0; st->need = k; st->count = 0; st->done = 0; }
}

// one block of 256 threads (the last CTA of a histogram pass): pick the digit whose cumulative count reaches `need`
__device__ __forceinline__ void sel_scan_block(int pass, SelState* st, const unsigned int* hist) {
    __shared__ unsigned int part[256];
    const unsigned int* h = hist + pass * kSelBins;
    const int nb = 1 << sel_width(pass);
    const int per = kSelBins / 256;  // 8 bins per thread
    unsigned int loc[per];
    unsigned int sum = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        const int b = threadIdx.x * per + i;
        loc[i] = b < nb ? __ldcg(h + b) : 0u;  // other CTAs' atomics: read through L2
        sum += loc[i];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {  // inclusive Hillis-Steele scan
        unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned int need = (unsigned int)st->need;
    const unsigned int before = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    const bool last = threadIdx.x == 255;
    __syncthreads();
    if ((before < need && need <= part[threadIdx.x]) || (last && need > part[255])) {
        unsigned int run = before;
        int i = 0;
        for (; i < per - 1; ++i) {
            if (run + loc[i] >= need) break;
            run += loc[i];
        }
        int b = threadIdx.x * per + i;
        if (b > nb - 1) b = nb - 1;
        st->need = (int)(need - run);
        st->prefix = (st->prefix << sel_width(pass)) | (unsigned long long)b;
    }
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(const float* __restrict__ values, int P, int worst, int pass, SelState* st, unsigned int* hist) {
    __shared__ unsigned int sh[kSelBins];
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int shift = sel_shift(pass), width = sel_width(pass);
    const unsigned long long prefix = st->prefix;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const unsigned long long key = Key32::make(values[i], (uint32_t)i, worst != 0).v;
        const bool match = pass == 0 || (key >> (shift + width)) == prefix;
        if (match) atomicAdd(&sh[(unsigned)(key >> shift) & ((1u << width) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[pass * kSelBins + i], sh[i]);
    // the CTA that arrives last has the complete histogram in front of it: it picks the digit (no second launch)
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    sel_scan_block(pass, st, hist);
    if (threadIdx.x == 0) st->done = 0;
}

__global__ void __launch_bounds__(256)
sel_collect_kernel(const float* __restrict__ values, int P, int worst, int k, SelState* st, Key32* __restrict__ out) {
    const unsigned long long K = st->prefix;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const Key32 key = Key32::make(values[i], (uint32_t)i, worst != 0);
        if (key.v <= K) {
            const unsigned int pos = atomicAdd(&st->count, 1u);
            if (pos < (unsigned int)k) out[pos] = key;
        }
    }
}

static int64_t select_workspace_bytes(int k) {
    return 256 + 256 + (int64_t)kSelPasses * kSelBins * 4 + (int64_t)k * 8 + 256;
}

template <typename Key>
static int64_t workspace_bytes(int P, int k) {
    const int tile = tile_for(k);
    int64_t n1 = (int64_t)((P + tile - 1) / tile) * k;
    int64_t n2 = ((n1 + tile - 1) / tile) * k;
    int64_t tiles = (n1 + n2) * (int64_t)sizeof(Key) + 256;
    int64_t sel = select_workspace_bytes(k);
    return tiles > sel ? tiles : sel;
}

static bool use_select(int P, int k) { return k > 64 && P > 4 * tile_for(k); }

static int run_select(const float* values, int P, int k, int worst, int32_t* idx, void* ws, int64_t ws_bytes,
                      cudaStream_t st) {
    if (!ws || ws_bytes < select_workspace_bytes(k)) return TF_ERR_WORKSPACE;
    char* base = reinterpret_cast<char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    SelState* state = reinterpret_cast<SelState*>(base);
    unsigned int* hist = reinterpret_cast<unsigned int*>(base + 256);
    Key32* keys = reinterpret_cast<Key32*>(base + 256 + kSelPasses * kSelBins * 4);
    sel_init_kernel<<<1, 256, 0, st>>>(state, hist, k);
    TF_LAUNCH_CHECK();
    const int grid = min((P + 255) / 256, sm_count() * 8);
    for (int pass = 0; pass < kSelPasses; ++pass) {
        sel_hist_kernel<<<grid, 256, 0, st>>>(values, P, worst, pass, state, hist);
        TF_LAUNCH_CHECK();
    }
    sel_collect_kernel<<<grid, 256, 0, st>>>(values, P, worst, k, state, keys);
    TF_LAUNCH_CHECK();
    const int tile = tile_for(k);
    const size_t smem = (size_t)tile * sizeof(Key32);
    auto kern = topk_tile_kernel<Key32, float>;
    if (smem > 48 * 1024) TF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<1, tile >= 4096 ? 1024 : kRankThreads, smem, st>>>(nullptr, keys, k, tile, k, worst, nullptr, idx);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
static int run_select(const double*, int, int, int, int32_t*, void*, int64_t, cudaStream_t) { return TF_ERR_UNSUPPORTED; }

// ------------------------------------------------------------------------------------------
// The k best (or worst) chromosomes as a SET, returned in ascending index order (the victims and the elites of
// island migration, where only a deterministic assignment matters, not the rank order; k = (world-1) * E = 7168
// victims at 8 GPUs).
constexpr int kSetBlock = 1024;  // sizes the caller's workspace (kept from the multi-launch version: ABI-stable)

static int64_t set_workspace_bytes(int P) {
    return 256 + 256 + (int64_t)kSelPasses * kSelBins * 4 + (int64_t)((P + kSetBlock - 1) / kSetBlock) * 4 + 256;
}

// ------------------------------------------------------------------------------------------
// tf_select_set in ONE launch (round 2).  The victims of an elite exchange sit on the critical path of the island
// model (and so do the elites, once their selection no longer competes with the generation kernel for SMs): ten
// short launches cost ~90 us at P = 1 M although the fitness vector is 4 MB.  One cooperative kernel (every CTA
// resident, grid barriers through a global counter) does the whole job:
//   passes 0-2  MSD radix select on the 32-bit orderable FITNESS key h (11 + 11 + 10 bits): every CTA adds its
//               shared-memory histogram to the global one, waits at the barrier, and scans the 2048 bins itself
//               (no designated CTA, no second barrier) -> threshold T, the number of keys below it, the number
//               of ties at it;
//   pass 3      ordered compaction.  A CTA owns a contiguous chunk of the index range; counts of (h < T, h == T)
//               per CTA -> barrier -> every CTA sums the counts of the CTAs before it -> scatter.  Ties at T are
//               resolved by index exactly as the composite 64-bit key did: the t smallest indices for `best`,
//               the t largest for `worst`.
// Output: the k selected indices in ascending index order — bit-identical to the multi-launch version.
constexpr int kOneThreads = 256;
constexpr int kOneMaxGrid = 1024;
struct OneCtl {
    unsigned int arrived;      // grid barrier: monotonically increasing
    unsigned int finished;     // CTAs that left the kernel (the last one hands the workspace back zeroed)
    unsigned int pad[2];
};
__device__ __forceinline__ uint32_t sel_key32(float f, bool worst) {
    if (f == 0.0f) f = 0.0f;  // -0.0 == +0.0 for Python's comparison
    uint32_t u = __float_as_uint(f);
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;  // ascending u <=> ascending f
    return worst ? u : ~u;                        // the k SMALLEST keys are selected
}
__device__ __forceinline__ void one_barrier(OneCtl* ctl, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ctl->arrived, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&ctl->arrived) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}
// every CTA: digit of this level whose cumulative count reaches `need` (1-based rank inside the current prefix)
__device__ __forceinline__ void one_pick(const unsigned int* __restrict__ hist, int nbins, unsigned int need, unsigned int* part,
                                         unsigned int& digit, unsigned int& before_out, unsigned int& at_out) {
    constexpr int per = kSelBins / kOneThreads;  // 8 bins per thread
    unsigned int loc[per], sum = 0;
#pragma unroll
    for (int i = 0; i < per; ++i) {
        const int bin = threadIdx.x * per + i;
        loc[i] = bin < nbins ? __ldcg(hist + bin) : 0u;  // other CTAs' atomics: read through L2
        sum += loc[i];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < kOneThreads; off <<= 1) {  // inclusive Hillis-Steele scan
        const unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned int before = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    __shared__ unsigned int res[3];
    if (before < need && need <= part[threadIdx.x]) {
        unsigned int run = before;
        int i = 0;
        for (; i < per - 1; ++i) {
            if (run + loc[i] >= need) break;
            run += loc[i];
        }
        res[0] = (unsigned int)(threadIdx.x * per + i);
        res[1] = run;
        res[2] = loc[i];
    }
    __syncthreads();
    digit = res[0];
    before_out = res[1];
    at_out = res[2];
    __syncthreads();
}

__global__ void __launch_bounds__(kOneThreads)
select_set_one_kernel(const float* __restrict__ values, int P, int k, int worst, OneCtl* ctl, unsigned int* __restrict__ hist,
                      unsigned int* __restrict__ cta_counts, int32_t* __restrict__ idx) {
    __shared__ unsigned int sh[kSelBins];
    __shared__ unsigned int part[kOneThreads];
    __shared__ unsigned int warp_tot[2][kOneThreads / 32];
    const unsigned int G = gridDim.x;
    const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
    uint32_t prefix = 0;          // digits chosen so far (high bits of T)
    unsigned int need = (unsigned int)k, n_eq = 0;
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = threadIdx.x; i < kSelBins; i += kOneThreads) sh[i] = 0;
        __syncthreads();
        const int shift = shifts[pass], width = widths[pass];
        for (int i = blockIdx.x * kOneThreads + threadIdx.x; i < P; i += G * kOneThreads) {
            const uint32_t h = sel_key32(values[i], worst != 0);
            if (pass == 0 || (h >> (shift + width)) == prefix) atomicAdd(&sh[(h >> shift) & ((1u << width) - 1u)], 1u);
        }
        __syncthreads();
        unsigned int* hg = hist + pass * kSelBins;
        for (int i = threadIdx.x; i < kSelBins; i += kOneThreads)
            if (sh[i]) atomicAdd(&hg[i], sh[i]);
        one_barrier(ctl, (unsigned int)(pass + 1) * G);
        unsigned int digit, before, at;
        one_pick(hg, 1 << width, need, part, digit, before, at);
        prefix = (prefix << width) | digit;
        need -= before;           // rank among the keys of the chosen bin
        n_eq = at;
    }
    const uint32_t T = prefix;                     // the k-th smallest key
    const unsigned int t = need;                   // ties at T that are selected (1 <= t <= n_eq)
    // ---- ordered compaction over contiguous chunks
    const int chunk = ((P + (int)G - 1) / (int)G + kOneThreads - 1) / kOneThreads * kOneThreads;
    const int lo = blockIdx.x * chunk, hi = min(P, lo + chunk);
    unsigned int c_less = 0, c_eq = 0;
    for (int i = lo + threadIdx.x; i < hi; i += kOneThreads) {
        const uint32_t h = sel_key32(values[i], worst != 0);
        c_less += h < T ? 1u : 0u;
        c_eq += h == T ? 1u : 0u;
    }
    c_less = __reduce_add_sync(0xffffffffu, c_less);
    c_eq = __reduce_add_sync(0xffffffffu, c_eq);
    if ((threadIdx.x & 31) == 0) { warp_tot[0][threadIdx.x >> 5] = c_less; warp_tot[1][threadIdx.x >> 5] = c_eq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0, e = 0;
        for (int w = 0; w < kOneThreads / 32; ++w) { a += warp_tot[0][w]; e += warp_tot[1][w]; }
        cta_counts[2 * blockIdx.x] = a;
        cta_counts[2 * blockIdx.x + 1] = e;
    }
    one_barrier(ctl, 4u * G);
    unsigned int less_before = 0, eq_before = 0;
    for (unsigned int b = threadIdx.x; b < blockIdx.x; b += kOneThreads) {
        less_before += __ldcg(cta_counts + 2 * b);
        eq_before += __ldcg(cta_counts + 2 * b + 1);
    }
    __syncthreads();
    part[threadIdx.x] = less_before;
    sh[threadIdx.x] = eq_before;
    __syncthreads();
    for (int off = kOneThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) { part[threadIdx.x] += part[threadIdx.x + off]; sh[threadIdx.x] += sh[threadIdx.x + off]; }
        __syncthreads();
    }
    less_before = part[0];
    eq_before = sh[0];
    __syncthreads();
    // ties: `best` keeps eq ranks [0, t), `worst` keeps [n_eq - t, n_eq)  (ranks in ascending index order)
    const unsigned int eq_lo = worst ? n_eq - t : 0u, eq_hi = worst ? n_eq : t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = lo; base < hi; base += kOneThreads) {
        const int i = base + threadIdx.x;
        bool is_less = false, is_eq = false;
        if (i < hi) {
            const uint32_t h = sel_key32(values[i], worst != 0);
            is_less = h < T;
            is_eq = h == T;
        }
        const unsigned int m_less = __ballot_sync(0xffffffffu, is_less), m_eq = __ballot_sync(0xffffffffu, is_eq);
        if (lane == 0) { warp_tot[0][warp] = __popc(m_less); warp_tot[1][warp] = __popc(m_eq); }
        __syncthreads();
        unsigned int lb = less_before, eb = eq_before, tl = 0, te = 0;
        for (int w = 0; w < kOneThreads / 32; ++w) {
            if (w < warp) { lb += warp_tot[0][w]; eb += warp_tot[1][w]; }
            tl += warp_tot[0][w];
            te += warp_tot[1][w];
        }
        lb += __popc(m_less & ((1u << lane) - 1u));
        eb += __popc(m_eq & ((1u << lane) - 1u));
        // selected ties before this element: |[eq_lo, eq_hi) ∩ [0, eb)|
        const unsigned int sel_eq_before = eb <= eq_lo ? 0u : (eb < eq_hi ? eb - eq_lo : eq_hi - eq_lo);
        if (is_less || (is_eq && eb >= eq_lo && eb < eq_hi)) {
            const unsigned int pos = lb + sel_eq_before;
            if (pos < (unsigned int)k) idx[pos] = i;
        }
        less_before += tl;
        eq_before += te;
        __syncthreads();
    }
    // the last CTA to leave hands the workspace back zeroed (all CTAs are past their last read of it)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        part[0] = atomicAdd(&ctl->finished, 1u) == G - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (part[0]) {
        for (int i = threadIdx.x; i < 3 * kSelBins; i += kOneThreads) hist[i] = 0;
        if (threadIdx.x == 0) { ctl->arrived = 0; ctl->finished = 0; }
    }
}

static int run_select_set(const float* values, int P, int k, int worst, int32_t* idx, void* ws, int64_t ws_bytes,
                          cudaStream_t st) {
    if (!values || !idx) return TF_ERR_INVALID;
    if (P < 1 || k < 1 || k > P) return TF_ERR_SHAPE;
    if (!ws || ws_bytes < set_workspace_bytes(P)) return TF_ERR_WORKSPACE;
    int rc = ensure_device();
    if (rc != TF_OK) return rc;
    char* base = reinterpret_cast<char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    OneCtl* ctl = reinterpret_cast<OneCtl*>(base);
    unsigned int* hist = reinterpret_cast<unsigned int*>(base + 256);
    unsigned int* counts = reinterpret_cast<unsigned int*>(base + 256 + 3 * kSelBins * 4);   // 2 x grid words <= 8 KB
    // control block + histograms zeroed on the stream: the kernel also hands them back zeroed, but the caller's
    // workspace is uninitialised memory the first time and after an aborted launch
    TF_CUDA(cudaMemsetAsync(base, 0, 256 + 3 * kSelBins * 4, st));
    static int occ_cache = 0;
    int occ = __atomic_load_n(&occ_cache, __ATOMIC_RELAXED);
    if (occ <= 0) {
        TF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, select_set_one_kernel, kOneThreads, 0));
        __atomic_store_n(&occ_cache, occ, __ATOMIC_RELAXED);
    }
    // two CTAs per SM keep the barriers cheap; never more than can be resident (cooperative launch)
    int grid = min(min(2, max(occ, 1)) * sm_count(), min(kOneMaxGrid, (P + kOneThreads - 1) / kOneThreads));
    void* args[] = {(void*)&values, (void*)&P, (void*)&k, (void*)&worst, (void*)&ctl, (void*)&hist, (void*)&counts, (void*)&idx};
    TF_CUDA(cudaLaunchCooperativeKernel((const void*)select_set_one_kernel, dim3(grid), dim3(kOneThreads), args, 0, st));
    TF_LAUNCH_CHECK();
    return TF_OK;
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
    if (sizeof(Key) == 8 && use_select(P, k)) return run_select(values, P, k, worst, idx, ws, ws_bytes, st);
    if (k <= kSmallKMax) {
        auto small = topk_small_kernel<Key, T>;
        int nb = (P + kSmallTile - 1) / kSmallTile;
        if (nb == 1) {
            small<<<1, kRankThreads, 0, st>>>(values, nullptr, P, k, worst, nullptr, idx, nullptr);
            TF_LAUNCH_CHECK();
            return TF_OK;
        }
        if (!ws || ws_bytes < workspace_bytes<Key>(P, k)) return TF_ERR_WORKSPACE;
        Key* a = reinterpret_cast<Key*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
        Key* b = a + (int64_t)nb * k;
        if ((int64_t)nb * k <= kSmallTile) {  // one launch: the last CTA to finish merges the nb * k winners
            int* done = sched_slot();
            if (done == nullptr) return TF_ERR_NOMEM;
            small<<<nb, kRankThreads, 0, st>>>(values, nullptr, P, k, worst, a, idx, done);
            TF_LAUNCH_CHECK();
            return TF_OK;
        }
        small<<<nb, kRankThreads, 0, st>>>(values, nullptr, P, k, worst, a, nullptr, nullptr);
        TF_LAUNCH_CHECK();
        int n = nb * k;
        while (true) {
            nb = (n + kSmallTile - 1) / kSmallTile;
            if (nb == 1) {
                small<<<1, kRankThreads, 0, st>>>(nullptr, a, n, k, worst, nullptr, idx, nullptr);
                TF_LAUNCH_CHECK();
                return TF_OK;
            }
            small<<<nb, kRankThreads, 0, st>>>(nullptr, a, n, k, worst, b, nullptr, nullptr);
            TF_LAUNCH_CHECK();
            n = nb * k;
            Key* t = a; a = b; b = t;
        }
    }
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

extern "C" int64_t tf_select_set_workspace_bytes(int P) { return P < 1 ? TF_ERR_SHAPE : set_workspace_bytes(P); }
extern "C" int tf_select_set(const float* fit_dev, int P, int k, int worst, int32_t* idx_dev, void* workspace_dev,
                             int64_t workspace_bytes, void* stream) {
    return run_select_set(fit_dev, P, k, worst, idx_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}
