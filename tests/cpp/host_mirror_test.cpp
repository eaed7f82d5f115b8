// Exercises the C++ host mirror (include/plonky2_b200.hpp) against the reference's own known answers, the way
// the reference's in-file #[test]s do.  Built and run by tests/test_gpu_cpp_mirror.py on the GPU box:
//   g++ -std=c++17 -I include tests/cpp/host_mirror_test.cpp -L plonky2-lib_b200 -lgl_b200 -Wl,-rpath,...
#include <cstdio>
#include <cstdlib>

#include "plonky2_b200.hpp"

using namespace plonky2_b200;

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                    \
        }                                                                \
    } while (0)

static uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x ^= x >> 31;
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    return x >= P ? x - P : x;
}

int main() {
    Context ctx(0);
    // src/zkdsa/circuits/mod.rs:85-105 test_default_simple_signature
    HashOut zero{{0, 0, 0, 0}};
    HashOut h = PoseidonHash::two_to_one(ctx, zero, zero);
    HashOut want{{4330397376401421145ULL, 14124799381142128323ULL, 8742572140681234676ULL, 14345658006221440202ULL}};
    CHECK(h == want);
    // src/zkdsa/circuits/mod.rs:136-153: the same digest as the JSON of the public inputs spells it
    CHECK(h.to_hex() == "0xc71603f33a1144ca7953db0ab48808f4c4055e3364a246c33c18a9786cb0b359");
    // src/smt/gadgets/common.rs:67-101 vs src/smt/goldilocks_poseidon/mod.rs:167-181
    HashOut k{{1, 0, 0, 0}}, v{{2, 0, 0, 0}};
    HashOut leaf = PoseidonNodeHash::calc_leaf_hash_batch(ctx, &k, &v, 1)[0];
    HashOut a = PoseidonHash::hash_pad(ctx, {1, 0, 0, 0, 2, 0, 0, 0, 1});
    HashOut b = PoseidonHash::hash_no_pad(ctx, {1, 0, 0, 0, 2, 0, 0, 0, 1, 1, 0, 1});
    CHECK(leaf == a && a == b);
    CHECK(leaf.elements[0] == 9613647271972624781ULL);

    // src/smt: (1 -> 2), (12 -> 1), (5 -> 51) as in src/smt/gadgets/verify/mod.rs:24-34 -- bulk root, the process proofs
    // of the three inserts, and the batch verifier over them
    {
        std::vector<HashOut> keys = {{{1, 0, 0, 0}}, {{12, 0, 0, 0}}, {{5, 0, 0, 0}}};
        std::vector<HashOut> vals = {{{2, 0, 0, 0}}, {{1, 0, 0, 0}}, {{51, 0, 0, 0}}};
        HashOut root = SparseMerkleTreeBatch::root_of(ctx, keys, vals);
        HashOut want_root{{16994558480514381166ULL, 8559105504417206749ULL, 13458782878755336329ULL, 17099432696459526118ULL}};
        CHECK(root == want_root);
        std::vector<SparseMerkleProcessProof> proofs = SparseMerkleTreeBatch::set_proofs(ctx, keys, vals);
        CHECK(proofs.size() == 3 && proofs[2].new_root == want_root && proofs[0].is_old0 && !proofs[1].is_old0);
        CHECK(proofs[0].old_root == zero && proofs[1].old_root == proofs[0].new_root && proofs[2].old_root == proofs[1].new_root);
        CHECK(proofs[0].new_root == PoseidonNodeHash::calc_leaf_hash_batch(ctx, &keys[0], &vals[0], 1)[0]);
        CHECK(proofs[1].old_key == keys[0] && proofs[1].old_value == vals[0] && proofs[1].fnc == 2);
        std::vector<int32_t> st = SparseMerkleTreeBatch::check_process_proofs(ctx, proofs);
        CHECK(st[0] == 0 && st[1] == 0 && st[2] == 0);
        proofs[2].new_root.elements[1] ^= 1;      // a wrong new root is assert 5 of verify_smt_process_proof
        proofs[1].old_value.elements[0] ^= 1;     // a wrong old leaf changes the old root: assert 4
        st = SparseMerkleTreeBatch::check_process_proofs(ctx, proofs);
        CHECK(st[0] == 0 && st[1] == 4 && st[2] == 5);
    }

    // commit + open: every opened row hashes up to the cap through MerkleTree semantics
    const CircuitConfig cfg = CircuitConfig::standard_recursion_config();
    const uint32_t lg = 8, cols = 20;
    std::vector<std::vector<F>> values(cols, std::vector<F>(1u << lg));
    for (uint32_t j = 0; j < cols; j++)
        for (uint32_t i = 0; i < (1u << lg); i++) values[j][i] = splitmix(0x706C6F6E6B7932ULL ^ (((uint64_t)j << 32) + i));
    PolynomialBatch pb = PolynomialBatch::from_values(ctx, values, cfg.fri_config.rate_bits, false, cfg.fri_config.cap_height);
    CHECK(pb.cap.size() == 16 && pb.polynomials.size() == cols && pb.degree_log == lg);
    std::vector<F> leaves;
    std::vector<HashOut> digests;
    pb.download(&leaves, &digests);
    const uint64_t N = 1ull << (lg + 3);
    CHECK(leaves.size() == N * cols && digests.size() == 2 * (N - 16));
    // mirror: MerkleTree::new on the downloaded leaves reproduces digests and cap
    std::vector<std::vector<F>> rows(N);
    for (uint64_t i = 0; i < N; i++) rows[i].assign(leaves.begin() + i * cols, leaves.begin() + (i + 1) * cols);
    MerkleTree t = MerkleTree::build(ctx, rows, 4);
    CHECK(t.cap == pb.cap);
    CHECK(t.digests == digests);
    auto opened = pb.open({0, 77, N - 1});
    for (size_t q = 0; q < 3; q++) {
        uint64_t idx = q == 0 ? 0 : q == 1 ? 77 : N - 1;
        CHECK(opened.first[q] == rows[idx]);
        CHECK(opened.second[q].siblings == t.prove(idx).siblings);
        // verify_merkle_proof_to_cap
        HashOut d = PoseidonHash::hash_or_noop(ctx, opened.first[q]);
        uint64_t i = idx;
        for (auto& s : opened.second[q].siblings) {
            d = (i & 1) ? PoseidonHash::two_to_one(ctx, s, d) : PoseidonHash::two_to_one(ctx, d, s);
            i >>= 1;
        }
        CHECK(d == pb.cap[i]);
    }
    {   // the same check for the whole batch on the device, and a tampered row is caught
        std::vector<uint64_t> idx = {0, 77, N - 1};
        std::vector<bool> ok = MerkleTree::verify_batch(ctx, opened.first, idx, opened.second, pb.cap);
        CHECK(ok[0] && ok[1] && ok[2]);
        auto tampered = opened.first;
        tampered[1][3] ^= 1;
        ok = MerkleTree::verify_batch(ctx, tampered, idx, opened.second, pb.cap);
        CHECK(ok[0] && !ok[1] && ok[2]);
    }
    {   // OpeningSet::new at the base-field point 1: the sum of the coefficients (2^64 = 2^32 - 1 mod p)
        const F one[2] = {1, 0};
        auto at1 = pb.eval_at(one);
        CHECK(at1.size() == cols);
        const unsigned __int128 p = 0xFFFFFFFF00000001ULL;
        unsigned __int128 sum = 0;
        for (F cf : pb.polynomials[7]) sum = (sum + cf) % p;
        CHECK(at1[7][0] == (F)sum && at1[7][1] == 0);
    }
    CHECK(pb.get_lde_values(3, 8) == rows[192]);  // leaves[reverse_bits(3 * 8, 11)] = leaves[192]
    // panics like upstream
    bool threw = false;
    try {
        MerkleTree::build(ctx, std::vector<std::vector<F>>(8, std::vector<F>(5, 1)), 4);
    } catch (const Panic&) {
        threw = true;
    }
    CHECK(threw);
    std::puts("host_mirror_test ok");
    return 0;
}
