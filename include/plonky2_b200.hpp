// plonky2_b200.hpp -- header-only C++ host mirror of the plonky2 plugin surface that Plonky2-lib reaches,
// over the C ABI of gl_b200.h.  The reference's host language is Rust (no cargo/rustc in this image), so the
// host side above the C ABI is C++ with the SAME names, argument meaning and error behaviour as upstream:
//
//   plonky2::fri::oracle::PolynomialBatch::{from_values, from_coeffs, get_lde_values}
//   plonky2::hash::merkle_tree::MerkleTree::{new, get, prove}        (MerkleTree::build here: `new` is a keyword)
//   plonky2::hash::poseidon::PoseidonHash::{hash_no_pad, hash_pad, hash_or_noop, two_to_one}
//   plonky2::fri::FriConfig / plonk::circuit_data::CircuitConfig presets
//   src/smt/goldilocks_poseidon/mod.rs:158-184  PoseidonNodeHash::calc_node_hash
//
// Upstream is infallible and panics on contract violations; here a violation throws plonky2_b200::Panic.
// Nothing in this header computes field arithmetic: every result comes from libgl_b200.so.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gl_b200.h"

namespace plonky2_b200 {

using F = uint64_t;  // GoldilocksField, canonical u64

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct HashOut {
    F elements[4];
    bool operator==(const HashOut& o) const {
        return elements[0] == o.elements[0] && elements[1] == o.elements[1] && elements[2] == o.elements[2] &&
               elements[3] == o.elements[3];
    }
};
using MerkleCap = std::vector<HashOut>;
struct MerkleProof {
    std::vector<HashOut> siblings;
};

class Context {
   public:
    explicit Context(int device = 0) {
        int rc = gl_ctx_create(device, &ctx_);
        if (rc) throw Panic(rc, gl_last_error(nullptr));
    }
    ~Context() { gl_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    gl_ctx* raw() const { return ctx_; }
    void check(int rc) const {
        if (rc) throw Panic(rc, gl_last_error(ctx_));
    }
    void set_shard(uint32_t index, uint32_t count) { check(gl_ctx_set_shard(ctx_, index, count)); }

   private:
    gl_ctx* ctx_ = nullptr;
};

// ---- FriConfig / CircuitConfig presets (the only configuration Plonky2-lib passes down) -------------
struct FriReductionStrategy {  // ConstantArityBits(arity_bits, final_poly_bits)
    uint32_t arity_bits = 4, final_poly_bits = 5;
    std::vector<uint32_t> reduction_arity_bits(uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height) const {
        std::vector<uint32_t> out;
        while (degree_bits > final_poly_bits && degree_bits + rate_bits - arity_bits >= cap_height) {
            out.push_back(arity_bits);
            degree_bits -= arity_bits;
        }
        return out;
    }
};
struct FriConfig {
    uint32_t rate_bits = 3, cap_height = 4, proof_of_work_bits = 16;
    FriReductionStrategy reduction_strategy;
    uint32_t num_query_rounds = 28;
};
struct CircuitConfig {
    uint32_t num_wires = 135, num_routed_wires = 80, num_constants = 2, num_challenges = 2;
    bool zero_knowledge = false;
    uint32_t max_quotient_degree_factor = 8;
    FriConfig fri_config;
    static CircuitConfig standard_recursion_config() { return {}; }
    static CircuitConfig standard_ecc_config() {
        CircuitConfig c;
        c.num_wires = 136;
        return c;
    }
    static CircuitConfig wide_ecc_config() {
        CircuitConfig c;
        c.num_wires = 234;
        return c;
    }
};

// ---- PoseidonHash (Hasher<F>) -----------------------------------------------------------------------
struct PoseidonHash {
    static constexpr size_t HASH_SIZE = 32, SPONGE_WIDTH = 12, SPONGE_RATE = 8;

    // batch forms: the single-shot forms below call them with m = 1
    static std::vector<HashOut> hash_no_pad_batch(const Context& c, const F* inputs, uint32_t len_each, uint64_t m) {
        std::vector<HashOut> out(m);
        c.check(gl_poseidon_hash_no_pad_batch(c.raw(), inputs, len_each, m, &out[0].elements[0], GL_HOST));
        return out;
    }
    static std::vector<HashOut> two_to_one_batch(const Context& c, const HashOut* l, const HashOut* r, uint64_t m) {
        std::vector<HashOut> out(m);
        c.check(gl_poseidon_two_to_one_batch(c.raw(), &l[0].elements[0], &r[0].elements[0], &out[0].elements[0], m, GL_HOST));
        return out;
    }
    static HashOut hash_no_pad(const Context& c, const std::vector<F>& input) {
        return hash_no_pad_batch(c, input.data(), (uint32_t)input.size(), 1)[0];
    }
    static HashOut two_to_one(const Context& c, const HashOut& l, const HashOut& r) { return two_to_one_batch(c, &l, &r, 1)[0]; }
    // pad 1, 0*, 1 to a multiple of SPONGE_WIDTH (this fork generation: src/smt/goldilocks_poseidon/mod.rs:167-181
    // must equal src/smt/gadgets/common.rs:87-101)
    static HashOut hash_pad(const Context& c, const std::vector<F>& input) {
        std::vector<F> p(input);
        p.push_back(1);
        while ((p.size() + 1) % SPONGE_WIDTH) p.push_back(0);
        p.push_back(1);
        return hash_no_pad(c, p);
    }
    static HashOut hash_or_noop(const Context& c, const std::vector<F>& input) {
        if (input.size() <= 4) {
            HashOut h{{0, 0, 0, 0}};
            for (size_t i = 0; i < input.size(); i++) h.elements[i] = input[i] % 0xFFFFFFFF00000001ULL;
            return h;
        }
        return hash_no_pad(c, input);
    }
};

// src/smt/goldilocks_poseidon/mod.rs:158-184
struct PoseidonNodeHash {
    static std::vector<HashOut> calc_leaf_hash_batch(const Context& c, const HashOut* keys, const HashOut* values, uint64_t m) {
        std::vector<HashOut> out(m);
        c.check(gl_smt_leaf_hash_batch(c.raw(), &keys[0].elements[0], &values[0].elements[0], &out[0].elements[0], m, GL_HOST));
        return out;
    }
    static HashOut calc_internal_hash(const Context& c, const HashOut& l, const HashOut& r) {
        return PoseidonHash::two_to_one(c, l, r);
    }
};

// ---- MerkleTree ---------------------------------------------------------------------------------------
struct MerkleTree {
    std::vector<std::vector<F>> leaves;
    std::vector<HashOut> digests;
    MerkleCap cap;
    uint32_t cap_height = 0;

    // MerkleTree::new(leaves, cap_height)
    static MerkleTree build(const Context& c, std::vector<std::vector<F>> leaves, uint32_t cap_height) {
        MerkleTree t;
        const uint64_t n = leaves.size();
        const uint32_t len = n ? (uint32_t)leaves[0].size() : 0;
        std::vector<F> flat;
        flat.reserve(n * len);
        for (auto& row : leaves) {
            if (row.size() != len) throw Panic(GL_E_ARG, "MerkleTree::new: ragged leaves");
            flat.insert(flat.end(), row.begin(), row.end());
        }
        uint64_t nd = n >= (1ull << cap_height) ? 2 * (n - (1ull << cap_height)) : 0;
        t.digests.resize(nd);
        t.cap.resize(1ull << cap_height);
        c.check(gl_merkle_build(c.raw(), flat.data(), n, len, cap_height, nd ? &t.digests[0].elements[0] : nullptr,
                                &t.cap[0].elements[0], GL_HOST));
        t.leaves = std::move(leaves);
        t.cap_height = cap_height;
        return t;
    }
    const std::vector<F>& get(size_t i) const { return leaves[i]; }
    MerkleProof prove(size_t leaf_index) const {
        MerkleProof p;
        size_t n = leaves.size();
        uint32_t lg = 0;
        while ((1ull << lg) < n) lg++;
        uint32_t L = lg - cap_height;
        if (L == 0) return p;
        size_t per = 2 * ((1ull << L) - 1), sub = leaf_index >> L, pair = leaf_index & ((1ull << L) - 1);
        const HashOut* buf = digests.data() + sub * per;
        for (uint32_t i = 0; i < L; i++) {
            size_t parity = pair & 1;
            pair >>= 1;
            p.siblings.push_back(buf[2 * ((pair << (i + 1)) + (1ull << i) - 1) + (1 - parity)]);
        }
        return p;
    }
};

// ---- PolynomialBatch (resident mode) ------------------------------------------------------------------
class PolynomialBatch {
   public:
    std::vector<std::vector<F>> polynomials;  // coefficients
    MerkleCap cap;                            // merkle_tree.cap
    uint32_t degree_log = 0, rate_bits = 0, cap_height = 0;
    bool blinding = false;

    // values: one Vec per polynomial (Vec<PolynomialValues<F>>); `timing` / `fft_root_table` of the upstream
    // signature have no meaning on the device and are not taken.
    static PolynomialBatch from_values(const Context& c, const std::vector<std::vector<F>>& values, uint32_t rate_bits,
                                       bool blinding, uint32_t cap_height) {
        return make(c, values, true, rate_bits, blinding, cap_height);
    }
    static PolynomialBatch from_coeffs(const Context& c, const std::vector<std::vector<F>>& coeffs, uint32_t rate_bits,
                                       bool blinding, uint32_t cap_height) {
        return make(c, coeffs, false, rate_bits, blinding, cap_height);
    }
    // (tree.get(i), tree.prove(i)) for every i
    std::pair<std::vector<std::vector<F>>, std::vector<MerkleProof>> open(const std::vector<uint64_t>& idx) const {
        const uint32_t k = (uint32_t)idx.size(), cols = (uint32_t)polynomials.size();
        const uint32_t L = degree_log + rate_bits - cap_height;
        std::vector<F> rows((size_t)k * cols), paths((size_t)k * L * 4);
        ctx_->check(gl_commit_open(h_, idx.data(), k, rows.data(), paths.data(), GL_HOST));
        std::pair<std::vector<std::vector<F>>, std::vector<MerkleProof>> out;
        for (uint32_t q = 0; q < k; q++) {
            out.first.emplace_back(rows.begin() + (size_t)q * cols, rows.begin() + (size_t)(q + 1) * cols);
            MerkleProof p;
            for (uint32_t l = 0; l < L; l++) {
                HashOut d;
                for (int e = 0; e < 4; e++) d.elements[e] = paths[((size_t)q * L + l) * 4 + e];
                p.siblings.push_back(d);
            }
            out.second.push_back(std::move(p));
        }
        return out;
    }
    std::vector<F> get_lde_values(uint64_t index, uint64_t step) const {
        std::vector<F> row(polynomials.size());
        ctx_->check(gl_commit_get_lde_values(h_, &index, 1, step, row.data(), GL_HOST));
        return row;
    }
    // mirror mode: MerkleTree.leaves / .digests exactly as upstream lays them out
    void download(std::vector<F>* leaves_row_major, std::vector<HashOut>* digests) const {
        uint64_t lo, hi;
        gl_commit_info(h_, nullptr, nullptr, nullptr, nullptr, &lo, &hi);
        if (leaves_row_major) leaves_row_major->resize((hi - lo) * polynomials.size());
        uint64_t caps_local = (uint64_t)cap.size() * (hi - lo) >> (degree_log + rate_bits);
        if (digests) digests->resize(2 * ((hi - lo) - caps_local));
        ctx_->check(gl_commit_download(h_, leaves_row_major ? leaves_row_major->data() : nullptr,
                                       digests && !digests->empty() ? &(*digests)[0].elements[0] : nullptr, GL_HOST));
    }
    PolynomialBatch(PolynomialBatch&& o) noexcept { *this = std::move(o); }
    PolynomialBatch& operator=(PolynomialBatch&& o) noexcept {
        std::swap(polynomials, o.polynomials);
        std::swap(cap, o.cap);
        std::swap(degree_log, o.degree_log);
        std::swap(rate_bits, o.rate_bits);
        std::swap(cap_height, o.cap_height);
        std::swap(h_, o.h_);
        std::swap(ctx_, o.ctx_);
        return *this;
    }
    ~PolynomialBatch() {
        if (h_) gl_commit_free(h_);
    }

   private:
    PolynomialBatch() = default;
    static PolynomialBatch make(const Context& c, const std::vector<std::vector<F>>& polys, bool is_values,
                                uint32_t rate_bits, bool blinding, uint32_t cap_height) {
        if (blinding) throw Panic(GL_E_ARG, "blinding (zero_knowledge) is not supported: every reference config uses false");
        if (polys.empty()) throw Panic(GL_E_ARG, "PolynomialBatch: empty batch");
        const size_t n = polys[0].size();
        if (n == 0 || (n & (n - 1))) throw Panic(GL_E_ARG, "log2_strict: polynomial length is not a power of two");
        uint32_t lg = 0;
        while ((1ull << lg) < n) lg++;
        std::vector<F> flat;
        flat.reserve(n * polys.size());
        for (auto& p : polys) {
            if (p.size() != n) throw Panic(GL_E_ARG, "assert_eq!(p.len(), degree)");
            flat.insert(flat.end(), p.begin(), p.end());
        }
        PolynomialBatch b;
        b.ctx_ = &c;
        b.degree_log = lg;
        b.rate_bits = rate_bits;
        b.cap_height = cap_height;
        b.cap.resize(1ull << cap_height);
        std::vector<F> coeffs(is_values ? flat.size() : 0);
        int rc = is_values ? gl_commit_from_values(c.raw(), flat.data(), lg, (uint32_t)polys.size(), rate_bits, cap_height,
                                                   coeffs.data(), &b.cap[0].elements[0], &b.h_, GL_HOST)
                           : gl_commit_from_coeffs(c.raw(), flat.data(), lg, (uint32_t)polys.size(), rate_bits, cap_height,
                                                   &b.cap[0].elements[0], &b.h_, GL_HOST);
        c.check(rc);
        const std::vector<F>& src = is_values ? coeffs : flat;
        for (size_t j = 0; j < polys.size(); j++) b.polynomials.emplace_back(src.begin() + j * n, src.begin() + (j + 1) * n);
        return b;
    }
    gl_commit* h_ = nullptr;
    const Context* ctx_ = nullptr;
};

}  // namespace plonky2_b200
