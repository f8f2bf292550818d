// tf_tape.cu — host-side reproduction of the np.random draws HRMAGAGenerator.evolve consumes (HR:117-120).
//
// The drop-in must consume NumPy's legacy global RandomState (MT19937) exactly like the reference does, so that
// `np.random.seed(s); evolve(...)` returns the reference's path.  Drawing through Python costs ~8 us per child pair
// (np.random.choice(..., replace=False) alone is ~5 us), which was half of the drop-in's wall time at BASELINE
// configs[0].  The draws depend only on (P, W, mutation_rate) — never on fitness — so the whole [G][npairs] tape can
// be produced up front from a copy of the generator state.  This file restates, bit for bit, the four legacy
// RandomState methods on the path (NumPy 2.3: numpy/random/mtrand.pyx, src/distributions/distributions.c,
// src/mt19937/mt19937.c):
//   choice(n, 2, replace=False) = permutation(n)[:2] = Fisher-Yates from the top with random_interval(i)
//   randint(lo, hi)             = lo + masked rejection on 32-bit draws (no draw when hi - lo == 1)
//   rand()                      = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two 32-bit draws
//   choice([-1,0,1], 3)         = randint(0, 3, size=3) - 1
// The Python side verifies the restatement against np.random itself before trusting it (path_planner_hrmaga.py).
#include <stdint.h>
#include <string.h>

#include "../../include/tracefly_b200.h"

namespace {

struct MT {
    uint32_t key[624];
    int pos;
};
inline void mt_gen(MT& s) {
    const uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; ++i) {
        y = (s.key[i] & UP) | (s.key[i + 1] & LO);
        s.key[i] = s.key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    for (; i < 623; ++i) {
        y = (s.key[i] & UP) | (s.key[i + 1] & LO);
        s.key[i] = s.key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    y = (s.key[623] & UP) | (s.key[0] & LO);
    s.key[623] = s.key[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    s.pos = 0;
}
inline uint32_t mt_next(MT& s) {
    if (s.pos == 624) mt_gen(s);
    uint32_t y = s.key[s.pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
inline uint32_t mask_for(uint32_t max) {
    uint32_t m = max;
    m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
    return m;
}
// random_interval / buffered_bounded_masked_uint32: uniform on [0, max], max < 2^32 - 1
inline uint32_t mt_interval(MT& s, uint32_t max) {
    if (max == 0) return 0;
    const uint32_t mask = mask_for(max);
    uint32_t v;
    while ((v = (mt_next(s) & mask)) > max) {}
    return v;
}
inline double mt_double(MT& s) {
    const uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

}  // namespace

extern "C" int tf_np_hr_tape(uint32_t* mt_key, int32_t* mt_pos, int generations, int npairs, int n_top, int W,
                             double mutation_rate, tf_draw* tape_out) {
    if (!mt_key || !mt_pos || (generations > 0 && npairs > 0 && !tape_out)) return TF_ERR_INVALID;
    if (generations < 0 || npairs < 0 || n_top < 2 || n_top > 64 || W < 3 || *mt_pos < 0 || *mt_pos > 624) return TF_ERR_SHAPE;
    MT s;
    memcpy(s.key, mt_key, sizeof(s.key));
    s.pos = *mt_pos;
    const uint32_t cp_rng = (uint32_t)(W - 3);   // randint(1, W - 1): values 1 .. W - 2
    for (long long r = 0; r < (long long)generations * npairs; ++r) {
        tf_draw& t = tape_out[r];
        int perm[64];
        for (int i = 0; i < n_top; ++i) perm[i] = i;
        for (int i = n_top - 1; i >= 1; --i) {   // permutation(n_top): HR:118
            const int j = (int)mt_interval(s, (uint32_t)i);
            const int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
        }
        t.pa = perm[0];
        t.pb = perm[1];
        t.cp = 1 + (int)mt_interval(s, cp_rng);  // HR:86
        t._pad = 0;
        for (int c = 0; c < 2; ++c) {
            int mp = -1, d = 0;
            if (mt_double(s) < mutation_rate) {  // HR:94
                mp = 1 + (int)mt_interval(s, cp_rng);                       // HR:95
                const int dx = (int)mt_interval(s, 2u), dy = (int)mt_interval(s, 2u), dz = (int)mt_interval(s, 2u);  // HR:96
                d = dx | (dy << 2) | (dz << 4);
            }
            if (c == 0) { t.mp1 = mp; t.d1 = d; } else { t.mp2 = mp; t.d2 = d; }
        }
    }
    memcpy(mt_key, s.key, sizeof(s.key));
    *mt_pos = s.pos;
    return TF_OK;
}

// random.shuffle of the six axis moves in UAVPathPlanner.dodge_obstacle (GAE:150-157), `n` times, from a COPY of
// CPython's global Mersenne Twister (random.getstate()[1]: 624 key words + position):
//   shuffle(x): for i in reversed(range(1, len(x))): j = _randbelow(i + 1); x[i], x[j] = x[j], x[i]
//   _randbelow(n): k = n.bit_length(); r = getrandbits(k); while r >= n: r = getrandbits(k)
//   getrandbits(k <= 32) = genrand_uint32() >> (32 - k)
// out[s] packs the shuffled order of moves 0..5 (GAE:151-155 list order), 3 bits per position, first move lowest.
extern "C" int tf_py_shuffle6_tape(uint32_t* mt_key, int32_t* mt_pos, int n, int32_t* out) {
    if (!mt_key || !mt_pos || (n > 0 && !out)) return TF_ERR_INVALID;
    if (n < 0 || *mt_pos < 0 || *mt_pos > 624) return TF_ERR_SHAPE;
    MT s;
    memcpy(s.key, mt_key, sizeof(s.key));
    s.pos = *mt_pos;
    for (int r = 0; r < n; ++r) {
        int x[6] = {0, 1, 2, 3, 4, 5};
        for (int i = 5; i >= 1; --i) {
            const uint32_t lim = (uint32_t)(i + 1);
            int k = 0;
            while ((lim >> k) != 0u) ++k;                  // bit_length
            uint32_t j;
            do { j = mt_next(s) >> (32 - k); } while (j >= lim);
            const int tmp = x[i]; x[i] = x[(int)j]; x[(int)j] = tmp;
        }
        int32_t packed = 0;
        for (int i = 0; i < 6; ++i) packed |= x[i] << (3 * i);
        out[r] = packed;
    }
    memcpy(mt_key, s.key, sizeof(s.key));
    *mt_pos = s.pos;
    return TF_OK;
}
