// Host side of the duration-matched triplet sampler (pig/triplet.py:99-104 `_triplets` over pig/util.py:`shuffled` /
// `grouped` / `pairs`): the reference draws, per evaluation, n_samples x (one random.random() per clip + one
// random.sample(pair, 2) per pair) on Python's global generator -- 1.5 million interpreter-level calls for 500 samples of
// a 1467-clip gallery, more host time than everything the GPU does for that evaluation.  Seeded parity needs exactly
// those draws, in that order, and the generator left in the state the reference would leave it in; it does not need
// the interpreter.  This file replays them on a copy of the generator's MT19937 state (random.getstate()), which the
// caller writes back (random.setstate()): same triplets, same final state, bit for bit.
//
// CPython's generator (Modules/_randommodule.c, unchanged since 2.3 apart from getrandbits' argument handling):
//   genrand_uint32()  MT19937, 624-word state + index, standard tempering
//   random()          a = genrand_uint32() >> 5, b = genrand_uint32() >> 6 -> (a * 2^26 + b) / 2^53
//   getrandbits(k)    k <= 32: genrand_uint32() >> (32 - k)
//   random.sample(p, 2) on a pair (Lib/random.py): j = _randbelow(2); _randbelow(1) -> [p[j], p[1 - j]], with
//   _randbelow(n) = rejection sampling on getrandbits(n.bit_length()): for n = 2 two bits until < 2, for n = 1 one bit
//   until 0.  peppa_b200/triplet.py checks these semantics against the running interpreter before taking this path.
// No CUDA in here: plain C++ behind the C ABI, compiled into the same library.
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "host_util.h"
#include "peppa_b200.h"

namespace {
struct Mt19937 {
    uint32_t* mt;  // 624 words
    uint32_t idx;  // 0..624
    static constexpr int N = 624, M = 397;
    void refill() {
        constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
        int kk = 0;
        for (; kk < N - M; ++kk) {
            const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        }
        for (; kk < N - 1; ++kk) {
            const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
            mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        }
        const uint32_t y = (mt[N - 1] & kUpper) | (mt[0] & kLower);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        idx = 0;
    }
    uint32_t next() {
        if (idx >= (uint32_t)N) refill();
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double random() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
};
}  // namespace

extern "C" int pb2_host_random_doubles(uint32_t* mt_state, int64_t n, double* out) {
    if (!mt_state || (n > 0 && !out)) return pb2::set_error(PB2_ERR_ARG, "host_random_doubles: null");
    if (mt_state[624] > 624u) return pb2::set_error(PB2_ERR_ARG, "host_random_doubles: state index out of range");
    Mt19937 g{mt_state, mt_state[624]};
    for (int64_t i = 0; i < n; ++i) out[i] = g.random();
    mt_state[624] = g.idx;
    return PB2_OK;
}

extern "C" int pb2_host_sample_pairs(uint32_t* mt_state, const int64_t* items, const int64_t* group_start, int64_t n_groups,
                                     int64_t n_samples, int64_t* pos, int64_t* neg) {
    if (n_samples <= 0 || n_groups <= 0) return PB2_OK;
    if (!mt_state || !items || !group_start || !pos || !neg) return pb2::set_error(PB2_ERR_ARG, "host_sample_pairs: null");
    if (mt_state[624] > 624u) return pb2::set_error(PB2_ERR_ARG, "host_sample_pairs: state index out of range");
    int64_t longest = 0;
    for (int64_t g = 0; g < n_groups; ++g) {
        const int64_t len = group_start[g + 1] - group_start[g];
        if (len <= 0) return pb2::set_error(PB2_ERR_ARG, "host_sample_pairs: empty group");
        longest = std::max(longest, len);
    }
    Mt19937 gen{mt_state, mt_state[624]};
    // A key is random.random() = k53 / 2^53 with the 53-bit integer k53 = (a >> 5) * 2^26 + (b >> 6): sorting the integers
    // sorts the doubles.  One 64-bit word per item -- k53 in the high bits, the item's position in the low 11 (groups of
    // up to 2048 clips; longer ones sort (key, position) pairs) -- makes the plain sort the stable one that
    // sorted(key=...) is: equal keys keep their order.
    constexpr int kPosBits = 11;
    const bool packed = longest <= (int64_t(1) << kPosBits);
    std::vector<uint64_t> words(packed ? (size_t)longest : 0), sorted(packed ? (size_t)longest : 0);
    std::vector<std::pair<uint64_t, int64_t>> keyed(packed ? 0 : (size_t)longest);
    auto key53 = [&]() {
        const uint64_t a = gen.next() >> 5, b = gen.next() >> 6;
        return (a << 26) | b;
    };
    int64_t out = 0;
    for (int64_t s = 0; s < n_samples; ++s) {
        for (int64_t g = 0; g < n_groups; ++g) {
            const int64_t* grp = items + group_start[g];
            const int64_t len = group_start[g + 1] - group_start[g];
            // shuffled(items) = sorted(items, key=lambda _: random.random()): one key per item, in item order -- also for
            // a group of one --, then a STABLE ascending sort on the keys alone
            if (packed) {
                for (int64_t k = 0; k < len; ++k) words[(size_t)k] = (key53() << kPosBits) | (uint64_t)k;
                // The keys are uniform 53-bit numbers: 64 buckets on their top six bits put every word within a slot or
                // two of its place (counting pass, prefix sums, stable scatter), and one insertion pass finishes -- a
                // few hundred branch-predictable operations for the ~40 clips of a duration group, where the comparison
                // sort spends its time on mispredicted branches (17 of 32 ms per 500 samples of 1467 clips).
                uint32_t cnt[65] = {0};
                for (int64_t k = 0; k < len; ++k) ++cnt[(words[(size_t)k] >> 58) + 1];
                for (int b = 0; b < 64; ++b) cnt[b + 1] += cnt[b];
                for (int64_t k = 0; k < len; ++k) sorted[cnt[words[(size_t)k] >> 58]++] = words[(size_t)k];
                for (int64_t i = 1; i < len; ++i) {
                    const uint64_t w = sorted[(size_t)i];
                    int64_t j = i;
                    for (; j > 0 && sorted[(size_t)(j - 1)] > w; --j) sorted[(size_t)j] = sorted[(size_t)(j - 1)];
                    sorted[(size_t)j] = w;
                }
            } else {
                for (int64_t k = 0; k < len; ++k) keyed[(size_t)k] = {key53(), k};
                std::sort(keyed.begin(), keyed.begin() + len);
            }
            auto at = [&](int64_t k) {
                return grp[packed ? (int64_t)(sorted[(size_t)k] & ((uint64_t(1) << kPosBits) - 1)) : keyed[(size_t)k].second];
            };
            for (int64_t k = 0; k + 1 < len; k += 2) {  // pairs(xs); random.sample(pair, 2)
                uint32_t r = gen.next() >> 30;
                while (r >= 2u) r = gen.next() >> 30;
                uint32_t q = gen.next() >> 31;
                while (q) q = gen.next() >> 31;
                pos[out] = at(k + r);
                neg[out] = at(k + 1 - r);
                ++out;
            }
        }
    }
    mt_state[624] = gen.idx;
    return PB2_OK;
}
