// Measures the B200 instruction-pipe rates that bound the Poseidon kernels: IMAD.WIDE.U32, IMAD (32-bit),
// DFMA, DADD, IADD3, LOP3/SHF, and pairs of them issued together (are the pipes independent?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu
// Output: one JSON object (ops per second per GPU and per clock per SM, with the SM clock each figure was measured at:
// clock64 ticks of one block over the CUDA-event time of its launch, and nvidia-smi's view before / after)
// -> profiles/r2_pipe_peaks.json
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(unsigned long long* out, unsigned seed, long long* ticks) {
    const long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    unsigned a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
    unsigned long long w[CHAINS];
    double d[CHAINS];
    unsigned u[CHAINS], v[CHAINS];
    unsigned long long x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        w[i] = (unsigned long long)a * (i + 3);
        d[i] = (double)(a & 0xffff) + i;
        u[i] = a + i * 77u;
        v[i] = b + i * 13u;
        x[i] = w[i] ^ 0x1234567ull;
    }
    const double c = (double)(seed & 7) + 1.0, c2 = (double)(seed & 3) + 0.5;
    const unsigned long long w0 = w[0] | 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE & 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
            if (MODE & 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(a), "r"(b));
            if (MODE & 4) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(c), "d"(c2));
            if (MODE & 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c));
            if (MODE & 16) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (MODE & 32) asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (MODE & 64) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(a), "r"(b));
            if (MODE & 128) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(v[i]) : "r"(b));
            if (MODE & 256) asm volatile("add.u64 %0, %0, %1;" : "+l"(x[i]) : "l"(w0));
        }
    }
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += w[i] + (unsigned long long)d[i] + u[i] + v[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g1;
        const long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        ticks[0] = t1 - t0;
        ticks[1] = (long long)(g1 - g0);   // ns
    }
}

static double g_mhz_sum = 0, g_mhz_min = 1e30, g_mhz_max = 0;
static int g_mhz_n = 0;
template <int MODE>
static double run(unsigned long long* out, int blocks, int ops_per_iter) {
    static long long* ticks = nullptr;
    if (!ticks) cudaMallocManaged(&ticks, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<MODE><<<blocks, 256>>>(out, 1, ticks);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_pipe<MODE><<<blocks, 256>>>(out, r + 2, ticks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        // SM clock while block 0 ran: its clock64 ticks over its %globaltimer nanoseconds
        const double mhz = (double)ticks[0] * 1e3 / (double)ticks[1];
        g_mhz_sum += mhz; g_mhz_n++;
        if (mhz < g_mhz_min) g_mhz_min = mhz;
        if (mhz > g_mhz_max) g_mhz_max = mhz;
    }
    double ops = (double)blocks * 256 * ITERS * CHAINS * ops_per_iter;
    return ops / (best * 1e-3);
}

static void smi(const char* key) {
    FILE* f = popen("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv,noheader -i 0 2>/dev/null", "r");
    char buf[256] = "";
    if (f) {
        if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
        pclose(f);
    }
    for (char* q = buf; *q; q++)
        if (*q == '\n' || *q == '"') *q = ' ';
    printf(", \"%s\": \"%s\"", key, buf);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int blocks = sms * 8;
    unsigned long long* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    printf("%s", "");
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double clk = khz * 1e3;
    struct { const char* name; double v; } res[] = {
        {"imad_wide", run<1>(out, blocks, 1)}, {"imad_lo", run<2>(out, blocks, 1)}, {"dfma", run<4>(out, blocks, 1)},
        {"dadd", run<8>(out, blocks, 1)}, {"iadd3", run<16>(out, blocks, 1)}, {"lop3", run<32>(out, blocks, 1)},
        {"imad_hi", run<64>(out, blocks, 1)}, {"shf", run<128>(out, blocks, 1)}, {"iadd64", run<256>(out, blocks, 1)},
        {"imad_wide+dfma", run<1 | 4>(out, blocks, 2)}, {"imad_wide+iadd3", run<1 | 16>(out, blocks, 2)},
        {"imad_wide+lop3", run<1 | 32>(out, blocks, 2)}, {"imad_lo+lop3", run<2 | 32>(out, blocks, 2)},
        {"imad_lo+iadd3", run<2 | 16>(out, blocks, 2)}, {"dfma+lop3", run<4 | 32>(out, blocks, 2)},
        {"dfma+iadd3", run<4 | 16>(out, blocks, 2)}, {"dfma+imad_lo", run<4 | 2>(out, blocks, 2)},
        {"imad_wide+dfma+iadd3", run<1 | 4 | 16>(out, blocks, 3)}, {"imad_wide+dfma+lop3+iadd3", run<1 | 4 | 16 | 128>(out, blocks, 4)},
        {"iadd3+lop3", run<16 | 128>(out, blocks, 2)}, {"imad_wide+iadd64", run<1 | 256>(out, blocks, 2)},
        {"imad_lo+dfma+iadd3", run<2 | 4 | 16>(out, blocks, 3)},
    };
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %.0f", p.name, sms, clk / 1e6);
    printf(", \"clocks\": {\"sm_mhz_measured_mean\": %.1f, \"sm_mhz_measured_min\": %.1f, \"sm_mhz_measured_max\": %.1f, \"launches\": %d, "
           "\"how\": \"clock64 ticks of block 0 over its %%globaltimer nanoseconds, every timed launch\"", g_mhz_sum / g_mhz_n, g_mhz_min, g_mhz_max, g_mhz_n);
    smi("nvidia_smi_after_sm_max_reasons");
    printf("}");
    for (auto& r : res)
        printf(", \"%s\": {\"ops_per_s\": %.4e, \"per_clk_per_sm\": %.2f}", r.name, r.v, r.v / clk / sms);
    printf("}\n");
    return 0;
}
