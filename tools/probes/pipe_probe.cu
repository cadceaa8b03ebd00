// Integer pipe throughput probe (per SM, warp-instructions per clock):
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_probe.cu -o pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(1024) probe(uint32_t *out, unsigned long long *cyc, uint32_t seed)
{
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
    uint32_t y = seed ^ 0x9e3779b9u, z = seed | 5u;
    unsigned long long acc64[4] = {1, 2, 3, 4};
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = x[i] * y + z;                                  // IMAD
            else if (OP == 1) x[i] = __umulhi(x[i], y) + z;                    // IMAD.HI
            else if (OP == 2) asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(x[i]) : "r"(x[i]), "r"(y), "r"(z & 7u));   // SHF
            else if (OP == 3) x[i] = max(x[i], y) ^ 0;                         // VIMNMX (the ^0 folds away)
            else if (OP == 4) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(x[i]), "r"(y), "r"(z));        // LOP3
            else if (OP == 5) acc64[i & 3] += (unsigned long long)x[i] * y;    // IMAD.WIDE (accumulating)
            else if (OP == 6) x[i] = __dp4a(x[i], y, z);                       // IDP.4A
            else if (OP == 7) x[i] = __byte_perm(x[i], y, z);                  // PRMT
            else if (OP == 8) x[i] = x[i] + y + z;                             // IADD3
            else if (OP == 9) { x[i] = max(x[i], y); x[(i + 1) & 7] = x[(i + 1) & 7] * y + z; }   // VIMNMX + IMAD mix
            else if (OP == 10) x[i] = __dp2a_lo(x[i], y, z);                   // IDP.2A
            else if (OP == 11) { x[i] = __dp2a_lo(x[i], y, z); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[(i + 1) & 7]) : "r"(x[(i + 1) & 7]), "r"(y), "r"(z)); }   // IDP.2A + LOP3
            else if (OP == 12) { x[i] = __dp2a_lo(x[i], y, z); x[(i + 1) & 7] = x[(i + 1) & 7] * y + z; }   // IDP.2A + IMAD
            else if (OP == 13) { x[i] = __dp4a(x[i], y, z); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[(i + 1) & 7]) : "r"(x[(i + 1) & 7]), "r"(y), "r"(z)); }   // IDP.4A + LOP3
            else if (OP == 14) { x[i] = __dp4a(x[i], y, z); x[(i + 1) & 7] = x[(i + 1) & 7] * y + z; }   // IDP.4A + IMAD
            else if (OP == 16) x[i] = __vmaxu2(x[i], y);                       // VIMNMX.U16x2
            else if (OP == 17) x[i] = __vadd2(x[i], y);                        // VIADD.16x2
            else if (OP == 18) x[i] = __vimax3_u32(x[i], y, z);                // VIMNMX3
            else if (OP == 19) { x[i] = __vmaxu2(x[i], y); x[(i + 1) & 7] = x[(i + 1) & 7] * y + z; }   // VIMNMX.U16x2 + IMAD
            else if (OP == 15) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(x[i]), "r"(y), "r"(z)); x[(i + 1) & 7] = x[(i + 1) & 7] * y + z; }   // LOP3 + IMAD
        }
    }
    unsigned long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    for (int i = 0; i < 4; ++i) s += (uint32_t)acc64[i] + (uint32_t)(acc64[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter)
{
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    probe<OP><<<148, 1024>>>(out, cyc, 12345u);
    probe<OP><<<148, 1024>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double winst = (double)ITERS * per_iter * 32;            // warp-instructions per SM (32 warps)
    printf("%-22s %.2f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, winst / c, c / (winst / 4));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("IMAD", 8); run<1>("IMAD.HI (+add)", 8); run<2>("SHF", 8); run<3>("VIMNMX", 8); run<4>("LOP3", 8);
    run<5>("IMAD.WIDE acc", 8); run<6>("IDP.4A", 8); run<7>("PRMT", 8); run<8>("IADD3", 8); run<9>("VIMNMX+IMAD", 16);
    run<10>("IDP.2A", 8); run<11>("IDP.2A+LOP3", 16); run<12>("IDP.2A+IMAD", 16); run<13>("IDP.4A+LOP3", 16); run<14>("IDP.4A+IMAD", 16); run<15>("LOP3+IMAD", 16);
    run<16>("VIMNMX.U16x2", 8); run<17>("VIADD.16x2", 8); run<18>("VIMNMX3", 8); run<19>("VIMNMX.U16x2+IMAD", 16);
    return 0;
}
