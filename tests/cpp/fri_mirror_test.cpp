// PolynomialBatch::prove_openings through the C++ host mirror (include/plonky2_b200.hpp): commits three oracles,
// proves the openings of a two-batch instance and writes every field of the FriProof, flattened, to argv[1].
// tests/test_gpu_cpp_mirror.py rebuilds the same proof with the oracle and compares the two streams.
#include <cstdio>

#include "plonky2_b200.hpp"

using namespace plonky2_b200;

static uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x ^= x >> 31;
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    return x >= P ? x - P : x;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    Context ctx(0);
    const uint32_t degree_bits = 8, cols[3] = {4, 9, 3};
    FriConfig cfg;
    cfg.proof_of_work_bits = 10;
    const FriParams fp = FriParams::for_degree(cfg, degree_bits);
    std::vector<PolynomialBatch> batches;
    for (uint32_t k = 0; k < 3; k++) {
        std::vector<std::vector<F>> values(cols[k], std::vector<F>(1u << degree_bits));
        for (uint32_t j = 0; j < cols[k]; j++)
            for (uint32_t i = 0; i < (1u << degree_bits); i++) values[j][i] = splitmix((700 + k) ^ (((uint64_t)j << 32) + i));
        batches.push_back(PolynomialBatch::from_values(ctx, values, cfg.rate_bits, false, cfg.cap_height));
    }
    std::vector<const PolynomialBatch*> oracles = {&batches[0], &batches[1], &batches[2]};
    const F zeta[2] = {0x0123456789abcdefULL, 0x0fedcba987654321ULL};
    const F g = detail::pow_mod(detail::pow_mod(7, (0xFFFFFFFF00000001ULL - 1) >> 32), 1ull << (32 - degree_bits));
    FriBatchInfo b0, b1;
    b0.point[0] = zeta[0];
    b0.point[1] = zeta[1];
    for (uint32_t oi = 0; oi < 3; oi++)
        for (uint32_t pi = 0; pi < cols[oi]; pi++) b0.polynomials.push_back({oi, pi});
    b1.point[0] = (F)((unsigned __int128)zeta[0] * g % 0xFFFFFFFF00000001ULL);
    b1.point[1] = (F)((unsigned __int128)zeta[1] * g % 0xFFFFFFFF00000001ULL);
    b1.polynomials = {{2, 0}, {2, 1}};
    Challenger ch(ctx);
    for (auto* o : oracles) ch.observe_cap(o->cap);
    FriProof proof = prove_openings(ctx, {b0, b1}, oracles, ch, fp);

    FILE* f = std::fopen(argv[1], "w");
    auto put = [&](uint64_t v) { std::fprintf(f, "%llu\n", (unsigned long long)v); };
    for (auto& cap : proof.commit_phase_merkle_caps)
        for (auto& h : cap)
            for (int e = 0; e < 4; e++) put(h.elements[e]);
    for (auto v : proof.final_poly) put(v);
    put(proof.pow_witness);
    for (auto& r : proof.query_round_proofs) {
        put(r.x_index);
        for (auto& ip : r.initial_trees_proof) {
            for (auto v : ip.first) put(v);
            for (auto& h : ip.second.siblings)
                for (int e = 0; e < 4; e++) put(h.elements[e]);
        }
        for (auto& s : r.steps) {
            for (auto v : s.evals) put(v);
            for (auto& h : s.merkle_proof.siblings)
                for (int e = 0; e < 4; e++) put(h.elements[e]);
        }
    }
    put(ch.get_challenge());  // the transcript ends in the same state
    std::fclose(f);
    std::puts("fri_mirror_test ok");
    return 0;
}
