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
    const F last = ch.get_challenge();
    put(last);  // the transcript ends in the same state
    std::fclose(f);

    // the one-call prover (gl_fri_prove, transcript on the device) returns the same FriProof and the same transcript
    {
        Challenger ch1(ctx);
        for (auto* o : oracles) ch1.observe_cap(o->cap);
        FriProof one = prove_openings_device(ctx, {b0, b1}, oracles, ch1, fp);
        bool same = one.pow_witness == proof.pow_witness && one.final_poly == proof.final_poly &&
                    one.commit_phase_merkle_caps.size() == proof.commit_phase_merkle_caps.size() &&
                    one.query_round_proofs.size() == proof.query_round_proofs.size() && ch1.get_challenge() == last;
        for (size_t i = 0; same && i < one.commit_phase_merkle_caps.size(); i++)
            for (size_t j = 0; j < one.commit_phase_merkle_caps[i].size(); j++)
                for (int e = 0; e < 4; e++)
                    same = same && one.commit_phase_merkle_caps[i][j].elements[e] == proof.commit_phase_merkle_caps[i][j].elements[e];
        auto same_path = [](const MerkleProof& a, const MerkleProof& b) {
            if (a.siblings.size() != b.siblings.size()) return false;
            for (size_t i = 0; i < a.siblings.size(); i++)
                for (int e = 0; e < 4; e++)
                    if (a.siblings[i].elements[e] != b.siblings[i].elements[e]) return false;
            return true;
        };
        for (size_t q = 0; same && q < one.query_round_proofs.size(); q++) {
            auto &x = one.query_round_proofs[q], &y = proof.query_round_proofs[q];
            same = x.x_index == y.x_index && x.initial_trees_proof.size() == y.initial_trees_proof.size() && x.steps.size() == y.steps.size();
            for (size_t i = 0; same && i < x.initial_trees_proof.size(); i++)
                same = x.initial_trees_proof[i].first == y.initial_trees_proof[i].first &&
                       same_path(x.initial_trees_proof[i].second, y.initial_trees_proof[i].second);
            for (size_t i = 0; same && i < x.steps.size(); i++)
                same = x.steps[i].evals == y.steps[i].evals && same_path(x.steps[i].merkle_proof, y.steps[i].merkle_proof);
        }
        if (!same) {
            std::puts("prove_openings_device differs from prove_openings");
            return 3;
        }
    }
    // a one-rank Group runs the sharded commit path (no collective needed) and must reproduce oracle 1's cap and rows
    try {
        Group grp({&ctx});
        const uint32_t c1 = cols[1], n1 = 1u << degree_bits;
        std::vector<F> flat((size_t)c1 * n1);
        for (uint32_t j = 0; j < c1; j++)
            for (uint32_t i = 0; i < n1; i++) flat[(size_t)j * n1 + i] = splitmix((700 + 1) ^ (((uint64_t)j << 32) + i));
        Group::Sharded sh = grp.commit_from_values(flat, c1, degree_bits, cfg.rate_bits, cfg.cap_height);
        bool ok = sh.cap.size() == batches[1].cap.size();
        for (size_t i = 0; ok && i < sh.cap.size(); i++)
            for (int e = 0; e < 4; e++) ok = ok && sh.cap[i].elements[e] == batches[1].cap[i].elements[e];
        std::vector<uint64_t> idx = {0, 5, (uint64_t)(n1 << cfg.rate_bits) - 1};
        std::vector<F> rows, paths;
        grp.open(sh, idx, &rows, &paths);
        auto ref = batches[1].open(idx);
        for (size_t q = 0; ok && q < idx.size(); q++)
            for (uint32_t j = 0; j < c1; j++) ok = ok && rows[q * c1 + j] == ref.first[q][j];
        for (uint32_t j = 0; ok && j < c1; j++)
            for (uint32_t i = 0; i < n1; i++) ok = ok && sh.coefficients[(size_t)j * n1 + i] == batches[1].polynomials[j][i];
        Group::free(sh);
        if (!ok) {
            std::puts("Group commit differs from PolynomialBatch::from_values");
            return 4;
        }
        std::puts("group ok");
    } catch (const Panic& e) {
        if (e.code != GL_E_NCCL) throw;
        std::printf("group skipped: %s\n", e.what());   // no libnccl.so.2 for a stand-alone binary on this machine
    }
    std::puts("fri_mirror_test ok");
    return 0;
}
