// Throughput of the Poseidon permutation alone (state in registers, no memory traffic) + KAT check.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I plonky2-lib_b200/csrc -o tools/poseidon_bench tools/poseidon_bench.cu
#include <cuda_runtime.h>

#include <cstdio>

#include "poseidon.cuh"
#include "poseidon_constants.h"

#ifndef REPS
#define REPS 64
#endif
#ifndef BLOCK
#define BLOCK 256
#endif
#ifndef MINB
#define MINB 1
#endif

#ifdef TWO_STATES
// experiment: two independent states per thread, their rounds in the same basic block (does ptxas overlap the FP64 layers
// of one with the integer S-boxes of the other?)
__device__ __forceinline__ void poseidon_permute2(u64 s[12], u64 t[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) { s[i] = gl_add_c(s[i], c_poseidon_rc[i]); t[i] = gl_add_c(t[i], c_poseidon_rc[i]); }
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const double2* rc = c_poseidon_rc_split + 12 * (phase ? POSEIDON_FULL_HALF + POSEIDON_PARTIAL + 1 : 1);
#pragma unroll 1
        for (int r = 0; r < POSEIDON_FULL_HALF; r++, rc += 12) { poseidon_full_round(s, rc); poseidon_full_round(t, rc); }
        if (phase == 0) {
#pragma unroll 1
            for (int pair = 0; pair < POSEIDON_PARTIAL / 2; pair++) { poseidon_partial_pair(s, pair); poseidon_partial_pair(t, pair); }
        }
    }
}
__global__ void __launch_bounds__(BLOCK, MINB) k_bench(u64* out, u64 seed) {
    u64 s[12], t[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { s[i] = seed * (threadIdx.x + 1 + blockIdx.x * (u64)blockDim.x) + i; t[i] = s[i] ^ 0x5555; }
#pragma unroll 1
    for (int r = 0; r < REPS / 2; r++) poseidon_permute2(s, t);
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= s[i] ^ t[i];
    out[blockIdx.x * (u64)blockDim.x + threadIdx.x] = acc;
}
#else
__global__ void __launch_bounds__(BLOCK, MINB) k_bench(u64* out, u64 seed) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = seed * (threadIdx.x + 1 + blockIdx.x * (u64)blockDim.x) + i;
#pragma unroll 1
    for (int r = 0; r < REPS; r++) poseidon_permute(s);
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= s[i];
    out[blockIdx.x * (u64)blockDim.x + threadIdx.x] = acc;
}

#endif

__global__ void k_kat(u64* io) {
    u64 s[12];
    for (int i = 0; i < 12; i++) s[i] = io[i];
    poseidon_permute(s);
    for (int i = 0; i < 12; i++) io[i] = gl_canon(s[i]);
}

// exactness of the FP64 linear layers at the extremes: every 32-bit half at 0xffffffff / 0 / random, compared with
// 128-bit integer arithmetic on the device
__global__ void k_circ_extremes(const u64* halves, int cases, int* bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cases) return;
    const u64* h = halves + 12 * t;
    double x[12], y1[12], y2[12];
    for (int i = 0; i < 12; i++) x[i] = u32_as_denormal((u32)h[i]);
    poseidon_circ12<0>(x, y1);
    poseidon_circ12<1>(x, y2);
    const int C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 c1[12], c2[12];
    for (int r = 0; r < 12; r++) {
        u64 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u64)C[i] * h[(i + r) % 12];
        c1[r] = acc;
    }
    for (int r = 0; r < 12; r++) {
        u64 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u64)C[i] * c1[(i + r) % 12];
        c2[r] = acc;
    }
    for (int r = 0; r < 12; r++)
        if (double_bits(y1[r]) != c1[r] || double_bits(y2[r]) != c2[r]) atomicAdd(bad, 1);
}

// The same check for the SIGNED inputs of the split S-box: lanes in (-2^33, 2^32) for one layer (SQ = 0); lane 0 in that
// range and lanes 1..11 in [0, 2^32) for the fused pair (SQ = 1).  v[t][i] = (a, b): the lane is denormal(a) - 2 denormal(b).
__global__ void k_circ_signed(const u32* ab, int cases, int* bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cases) return;
    const u32* h = ab + 24 * t;
    const int C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    for (int sq = 0; sq < 2; sq++) {
        double x[12], y[12];
        long long xi[12], c1[12], c2[12];
        for (int i = 0; i < 12; i++) {
            const bool sgn = sq == 0 || i == 0;
            const double a = u32_as_denormal(h[2 * i]), b = u32_as_denormal(h[2 * i + 1]);
            x[i] = sgn ? (a - b) - b : a;
            xi[i] = sgn ? (long long)h[2 * i] - 2 * (long long)h[2 * i + 1] : (long long)h[2 * i];
        }
        if (sq == 0) poseidon_circ12<0>(x, y);
        else poseidon_circ12<1>(x, y);
        for (int r = 0; r < 12; r++) {
            long long acc = 0;
            for (int i = 0; i < 12; i++) acc += (long long)C[i] * xi[(i + r) % 12];
            c1[r] = acc;
        }
        for (int r = 0; r < 12; r++) {
            long long acc = 0;
            for (int i = 0; i < 12; i++) acc += (long long)C[i] * c1[(i + r) % 12];
            c2[r] = acc;
        }
        for (int r = 0; r < 12; r++) {
            // y is an integer multiple of 2^-1074 below 2^53 in magnitude: scale it up exactly and compare as an integer
            const long long got = (long long)((y[r] * 0x1p537) * 0x1p537);
            if (got != (sq ? c2[r] : c1[r])) atomicAdd(bad, 1);
        }
    }
}

int main() {
    u64 rc[360];
    if (!poseidon_constants::generate(rc)) { printf("constants fingerprint mismatch\n"); return 1; }
    cudaMemcpyToSymbol(c_poseidon_rc, rc, sizeof rc);
    static u64 split[31 * 12 * 2], pk[11 * 12 * 2];
    static const int circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
#ifdef POSEIDON_SPLIT_SBOX
    const bool signed_sbox = true;
#else
    const bool signed_sbox = false;
#endif
    poseidon_constants::linear_layer_tables(rc, circ, 8, signed_sbox, split, pk);
    cudaMemcpyToSymbol(c_poseidon_rc_split, split, sizeof split);
    cudaMemcpyToSymbol(c_poseidon_pair_k, pk, sizeof pk);
    u64 h[12], *d;
    cudaMalloc(&d, 96);
    const u64 want_iota[4] = {0xd64e1e3efc5b8e9eULL, 0x53666633020aaa47ULL, 0xd40285597c6a8825ULL, 0x613a4f81e81231d2ULL};
    const u64 want_max[4] = {0xbe0085cfc57a8357ULL, 0xd95af71847d05c09ULL, 0xcf55a13d33c1c953ULL, 0x95803a74f4530e82ULL};
    int ok = 1;
    for (int i = 0; i < 12; i++) h[i] = i;
    cudaMemcpy(d, h, 96, cudaMemcpyHostToDevice);
    k_kat<<<1, 1>>>(d);
    cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; i++) ok &= h[i] == want_iota[i];
    for (int i = 0; i < 12; i++) h[i] = 0xFFFFFFFF00000000ULL;
    cudaMemcpy(d, h, 96, cudaMemcpyHostToDevice);
    k_kat<<<1, 1>>>(d);
    cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 4; i++) ok &= h[i] == want_max[i];
    {
        const int cases = 1 << 16;
        u64* hh = (u64*)malloc((size_t)cases * 12 * 8);
        u64 st = 88172645463325252ULL;
        for (int t = 0; t < cases; t++)
            for (int i = 0; i < 12; i++) {
                st ^= st << 13; st ^= st >> 7; st ^= st << 17;
                u64 v = st & 0xFFFFFFFFULL;
                int mode = t & 7;   // 0: all max, 1: all zero, 2: alternating, 3: one hot max, 4..7: random with extremes
                if (mode == 0) v = 0xFFFFFFFFULL;
                else if (mode == 1) v = 0;
                else if (mode == 2) v = ((i + (t >> 3)) & 1) ? 0xFFFFFFFFULL : 0;
                else if (mode == 3) v = (i == (t >> 3) % 12) ? 0xFFFFFFFFULL : 0;
                else if (mode == 4 && (st >> 40) % 3 == 0) v = 0xFFFFFFFFULL;
                hh[(size_t)t * 12 + i] = v;
            }
        u64* dh;
        int *dbad, hbad = 0;
        cudaMalloc(&dh, (size_t)cases * 96);
        cudaMalloc(&dbad, 4);
        cudaMemcpy(dh, hh, (size_t)cases * 96, cudaMemcpyHostToDevice);
        cudaMemset(dbad, 0, 4);
        k_circ_extremes<<<cases / 128, 128>>>(dh, cases, dbad);
        cudaMemcpy(&hbad, dbad, 4, cudaMemcpyDeviceToHost);
        printf("{\"circ12_extreme_cases\": %d, \"mismatches\": %d}\n", cases, hbad);
        ok &= hbad == 0;
        free(hh);
    }
    {
        const int cases = 1 << 16;
        u32* hh = (u32*)malloc((size_t)cases * 24 * 4);
        u64 st = 0x9E3779B97F4A7C15ULL;
        for (int t = 0; t < cases; t++)
            for (int i = 0; i < 24; i++) {
                st ^= st << 13; st ^= st >> 7; st ^= st << 17;
                u32 v = (u32)st;
                const int mode = t & 7;   // extremes of a - 2b: all most negative, all most positive, alternating, random
                if (mode == 0) v = (i & 1) ? 0xFFFFFFFFu : 0u;
                else if (mode == 1) v = (i & 1) ? 0u : 0xFFFFFFFFu;
                else if (mode == 2) v = (((i >> 1) + (t >> 3)) & 1) ? ((i & 1) ? 0xFFFFFFFFu : 0u) : ((i & 1) ? 0u : 0xFFFFFFFFu);
                else if (mode == 3 && (st >> 40) % 3 == 0) v = 0xFFFFFFFFu;
                hh[(size_t)t * 24 + i] = v;
            }
        u32* dh;
        int *dbad, hbad = 0;
        cudaMalloc(&dh, (size_t)cases * 96);
        cudaMalloc(&dbad, 4);
        cudaMemcpy(dh, hh, (size_t)cases * 96, cudaMemcpyHostToDevice);
        cudaMemset(dbad, 0, 4);
        k_circ_signed<<<cases / 128, 128>>>(dh, cases, dbad);
        cudaMemcpy(&hbad, dbad, 4, cudaMemcpyDeviceToHost);
        printf("{\"circ12_signed_cases\": %d, \"mismatches\": %d}\n", cases, hbad);
        ok &= hbad == 0;
        free(hh);
    }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 16;
    u64* out;
    cudaMalloc(&out, (size_t)blocks * BLOCK * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_bench<<<blocks, BLOCK>>>(out, 3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k_bench<<<blocks, BLOCK>>>(out, 5 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    double perms = (double)blocks * BLOCK * REPS;
    printf("{\"kat_ok\": %d, \"perms_per_s\": %.4e, \"ms\": %.3f, \"block\": %d, \"minb\": %d, \"err\": \"%s\"}\n", ok, perms / (best * 1e-3), best,
           BLOCK, MINB, cudaGetErrorString(e));
    return ok ? 0 : 2;
}
