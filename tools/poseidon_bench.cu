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

int main() {
    u64 rc[360];
    if (!poseidon_constants::generate(rc)) { printf("constants fingerprint mismatch\n"); return 1; }
    cudaMemcpyToSymbol(c_poseidon_rc, rc, sizeof rc);
    static u64 split[31 * 12 * 2];
    for (int i = 0; i < 360; i++) { split[2 * i] = rc[i] & 0xFFFFFFFFULL; split[2 * i + 1] = rc[i] >> 32; }
    cudaMemcpyToSymbol(c_poseidon_rc_split, split, sizeof split);
    static u64 pk[11 * 12 * 2];
    for (int pair = 0; pair < 11; pair++) {
        const u64* r1 = rc + (4 + 2 * pair + 1) * 12;
        const u64* r2 = rc + (4 + 2 * pair + 2) * 12;
        for (int lane = 0; lane < 12; lane++) {
            u64 lo = r2[lane] & 0xFFFFFFFFULL, hi = r2[lane] >> 32;
            for (int i = 1; i < 12; i++) {
                lo += (u64)poseidon_mds_entry(lane, i) * (r1[i] & 0xFFFFFFFFULL);
                hi += (u64)poseidon_mds_entry(lane, i) * (r1[i] >> 32);
            }
            pk[(pair * 12 + lane) * 2] = lo;
            pk[(pair * 12 + lane) * 2 + 1] = hi;
        }
    }
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
