// Poseidon-Goldilocks round constants, regenerated at context creation.
//
// plonky2's ALL_ROUND_CONSTANTS (plonky2::hash::poseidon) are the first 360 draws of
// ChaCha8Rng::seed_from_u64(0).gen_range(0..p) (rand 0.8.5 / rand_chacha 0.3.1, the versions the
// reference locks at Cargo.lock:1135-1149).  No table of them exists in /root/reference, so the
// library derives them and refuses to start unless the derived table reproduces the fingerprints
// that the reference's own known-answer test implies (src/zkdsa/circuits/mod.rs:85-105 passes only
// with exactly this table; XOR / first / last words recorded in SURVEY.md 8c).
#pragma once
#include <cstdint>
#include <cstring>

namespace poseidon_constants {

struct ChaCha8Stream {
    uint32_t key[8];
    uint64_t block_counter = 0;
    uint32_t buf[16];
    int used = 16;

    static uint32_t rol(uint32_t v, int k) { return (v << k) | (v >> (32 - k)); }
    static void quarter(uint32_t* x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rol(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rol(x[b] ^ x[c], 7);
    }
    // rand_core::SeedableRng::seed_from_u64: a PCG32 stream stretches the u64 into the 256-bit key
    explicit ChaCha8Stream(uint64_t seed) {
        uint64_t pcg = seed;
        for (int i = 0; i < 8; i++) {
            pcg = pcg * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xorshifted = (uint32_t)(((pcg >> 18) ^ pcg) >> 27);
            unsigned rot = (unsigned)(pcg >> 59) & 31u;
            key[i] = rot ? ((xorshifted >> rot) | (xorshifted << (32 - rot))) : xorshifted;
        }
    }
    void refill() {
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};  // "expand 32-byte k"
        for (int i = 0; i < 8; i++) init[4 + i] = key[i];
        init[12] = (uint32_t)block_counter;
        init[13] = (uint32_t)(block_counter >> 32);
        init[14] = init[15] = 0;
        uint32_t x[16];
        std::memcpy(x, init, sizeof x);
        for (int dr = 0; dr < 4; dr++) {  // 8 rounds
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf[i] = x[i] + init[i];
        block_counter++;
        used = 0;
    }
    uint32_t next_u32() {
        if (used == 16) refill();
        return buf[used++];
    }
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return lo | (hi << 32);
    }
};

// Fills rc[360]; returns false if the fingerprints do not match.
inline bool generate(uint64_t* rc) {
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    ChaCha8Stream rng(0);
    int n = 0;
    while (n < 360) {
        // rand::distributions::uniform::UniformInt<u64>::sample_single(0, p): widening multiply with
        // rejection zone (p << leading_zeros(p)) - 1 = p - 1
        unsigned __int128 wide = (unsigned __int128)rng.next_u64() * P;
        if ((uint64_t)wide <= P - 1) rc[n++] = (uint64_t)(wide >> 64);
    }
    uint64_t x = 0;
    for (int i = 0; i < 360; i++) x ^= rc[i];
    return x == 0xd95d3c3bb2fe42e3ULL && rc[0] == 0xb585f766f2144405ULL && rc[359] == 0xbc8dfb627fe558fcULL;
}

}  // namespace poseidon_constants
