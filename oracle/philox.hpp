// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11; Random123 v1.14 published algorithm and
// known-answer vectors, checked in tests/test_oracle_philox.py).  The reference (Merzbild.jl) draws
// from one sequential `AbstractRNG` passed as the first argument of every operator
// (e.g. src/collisions/collision_ntc.jl:338, src/convection/convection_1D.jl:130); a GPU cannot replay
// one sequential stream, so both the oracle and the CUDA path key an independent stream per
// (seed, operator, substream, timestep, entity) -- see DESIGN.md "RNG convention".
#pragma once
#include <cmath>
#include <cstdint>

namespace mbo {

struct Philox4x32 {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k[0];
        const uint32_t n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; r++) {
            round(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// operator ids (the "stream" word of the counter); must match merzbild.jl_b200/csrc/mb_philox.cuh
enum : uint32_t {
    OP_NTC = 1, OP_CONVECT = 2, OP_MERGE = 3, OP_SWPM = 4, OP_FP = 5, OP_SAMPLE = 6, OP_USER = 7, OP_MERGE_GRID = 8
};

// One stream: key = 64-bit seed; counter = (block, entity, timestep, op | substream << 8).
// Draw d (0-based) of the stream is the (d & 1)-th double of block d >> 1:
// u = ((hi32 << 32 | lo32) >> 11) * 2^-53 in [0, 1), words (0,1) then (2,3) with word 2k+1 the high half.
struct PhiloxStream {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have;  // doubles left in buf (0, 1 or 2)

    PhiloxStream() : have(0) { key[0] = key[1] = 0; ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0; }
    PhiloxStream(uint64_t seed, uint32_t op, uint32_t substream, uint32_t timestep, uint32_t entity) {
        reset(seed, op, substream, timestep, entity);
    }
    void reset(uint64_t seed, uint32_t op, uint32_t substream, uint32_t timestep, uint32_t entity) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0;
        ctr[1] = entity;
        ctr[2] = timestep;
        ctr[3] = (op & 0xFFu) | (substream << 8);
        have = 0;
    }
    static inline double to_double(uint32_t lo, uint32_t hi) {
        const uint64_t u = ((uint64_t)hi << 32) | lo;
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    inline double rand() {
        if (have == 0) {
            Philox4x32::block(ctr, key, buf);
            ctr[0]++;
            have = 2;
            have--;
            return to_double(buf[0], buf[1]);
        }
        have--;
        return to_double(buf[2], buf[3]);
    }
};

// The "one rng object passed around" mode of the reference: a plain sequential generator with two engines.
//   engine 0: xoshiro256++ seeded with splitmix64 (CPU baseline driver, statistical tests);
//   engine 1: the LehmerRNG of StableRNGs.jl (the generator of the reference's test suite, `StableRNG(seed)`): state =
//             (seed << 1) | 1 as UInt128, each draw multiplies the state by 0x45a31efc5a35d971261fd0407a968add and returns its high
//             64 bits; rand(Float64) = reinterpret(Float64, 0x3ff0000000000000 | (u & 0x000fffffffffffff)) - 1.0 (Julia's generic
//             CloseOpen12 path for an rng whose native 52-bit type is UInt64); rand(rng, [-1.0, 1.0]) indexes the 2-element array with
//             the low bit of one draw (SamplerRangeFast, mask 1).  Pinned by the reference's golden runs: with this engine the oracle
//             reproduces test/data/*.nc of the reference to round-off (tests/test_oracle_reference_bitlevel.py).
struct SeqRng {
    uint64_t s[4];
    int engine = 0;
    unsigned __int128 lehmer = 1;
    explicit SeqRng(uint64_t seed = 1234) {
        uint64_t z = seed;
        for (int i = 0; i < 4; i++) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            s[i] = x ^ (x >> 31);
        }
    }
    static SeqRng stable_rng(uint64_t seed) {
        SeqRng r(seed);
        r.engine = 1;
        r.lehmer = ((unsigned __int128)seed << 1) | 1;
        return r;
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    inline uint64_t next() {
        if (engine == 1) {
            lehmer *= ((unsigned __int128)0x45a31efc5a35d971ull << 64) | 0x261fd0407a968addull;
            return (uint64_t)(lehmer >> 64);
        }
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    inline double rand() {
        const uint64_t u = next();
        if (engine == 1) {
            const uint64_t bits = 0x3ff0000000000000ull | (u & 0x000fffffffffffffull);
            double d;
            __builtin_memcpy(&d, &bits, 8);
            return d - 1.0;
        }
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
    // randn(rng) of Julia's Random (stdlib normal.jl) for engine 1: 256-strip ziggurat of Marsaglia & Tsang as tabulated by Octave's
    // randmtzig (r = 3.6541528853610088, 51-bit magnitude + sign bit from one 52-bit draw).  The tables are regenerated here from the
    // construction (Julia hard-codes the resulting literals).  With the strip area at full precision (0.004928673233974655; randmtzig's
    // source quotes it to 12 digits, 0.00492867323399, which shifts every sample by 3e-12 relative) the reference's golden Fokker-Planck
    // run is reproduced to ~1e-14 relative with every cell count exact over 3000 steps; a residual last-bit difference in a table entry
    // moves a sample by an ulp and a decision with probability ~2^-51.
    struct Ziggurat {
        uint64_t ki[256];
        double wi[256], fi[256];
        Ziggurat() {
            const double R = 3.6541528853610088, NM = 2251799813685248.0;  // 2^51
            const double AREA = 0.004928673233974655;  // r f(r) + int_r^inf f: the strip area to full precision (see above)
            double x1 = R;
            wi[255] = x1 / NM;
            fi[255] = std::exp(-0.5 * x1 * x1);
            ki[0] = (uint64_t)(x1 * fi[255] / AREA * NM);
            wi[0] = AREA / fi[255] / NM;
            fi[0] = 1.0;
            for (int i = 254; i > 0; i--) {
                const double x = std::sqrt(-2.0 * std::log(AREA / x1 + fi[i + 1]));
                ki[i + 1] = (uint64_t)(x / x1 * NM);
                wi[i] = x / NM;
                fi[i] = std::exp(-0.5 * x * x);
                x1 = x;
            }
            ki[1] = 0;
        }
    };
    inline double randn() {
        static const Ziggurat z;
        const double R = 3.6541528853610087963519472518, INV_R = 1.0 / R;
        while (true) {
            const uint64_t r = next() & 0x000fffffffffffffull;
            const int64_t rabs = (int64_t)(r >> 1);
            const int idx = (int)(rabs & 0xFF);
            const double x = (double)((r & 1) ? -rabs : rabs) * z.wi[idx];
            if ((uint64_t)rabs < z.ki[idx]) return x;
            if (idx == 0) {
                while (true) {
                    const double xx = -INV_R * std::log(rand());
                    const double yy = -std::log(rand());
                    if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -R - xx : R + xx;
                }
            } else if ((z.fi[idx - 1] - z.fi[idx]) * rand() + z.fi[idx] < std::exp(-0.5 * x * x)) {
                return x;
            }
        }
    }
    inline double sign() {  // one element of direction_signs = [-1.0, 1.0] (constants.jl:39)
        if (engine == 1) return (next() & 1) ? 1.0 : -1.0;  // SamplerRangeFast over 1:2: index = 1 + (u & 1)
        return rand() < 0.5 ? -1.0 : 1.0;
    }
};
using Xoshiro256pp = SeqRng;  // the name the operator templates and the C API were written against

}  // namespace mbo
