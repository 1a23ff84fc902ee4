// pipes.cu -- throughput of the integer instructions the hot-path kernel is made of, on sm_100a.
// Each test runs N independent dependency chains per thread, fully unrolled, all SMs busy, and reports
// warp-instructions / clock / SM (4 SMSPs: 4.0 = full issue rate, 2.0 = half-rate pipe).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 32
#define ITERS 256

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b)
{
    uint32_t r;
    if constexpr (OP == 0) r = __byte_perm(a, b, 0x2140);                               // PRMT
    else if constexpr (OP == 1) asm("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(r) : "r"(a), "r"(b), "r"(a ^ 0x55u));  // LOP3 (+1 lop)
    else if constexpr (OP == 2) r = __funnelshift_r(a, b, 1);                                       // SHF
    else if constexpr (OP == 3) r = a * 152618u + b;                                     // IMAD
    else if constexpr (OP == 4) r = __umulhi(a, b | 0x80000001u);                        // IMAD.HI
    else if constexpr (OP == 5) r = __dp4a(a, b, a);                                     // IDP4A
    else if constexpr (OP == 6) r = __viaddmin_s16x2_relu(a, b, 0x01ff01ffu);            // VIADDMNMX
    else if constexpr (OP == 7) r = __vadd2(a, b);                                       // VIADD.16x2
    else if constexpr (OP == 8) r = __vabsdiffu4(a, b);                                  // VABSDIFF4
    else if constexpr (OP == 9) asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(a));  // I2IP
    else if constexpr (OP == 10) r = __popc(a) + b;                                      // POPC (+iadd)
    else if constexpr (OP == 11) r = a + b;                                              // IADD3
    else if constexpr (OP == 12) r = a * 128u + b;                                          // IMAD.SHL / SHL
    else if constexpr (OP == 13) r = __vimin_s16x2_relu(a, b);                           // VIMNMX
    else if constexpr (OP == 14) r = __dp2a_lo(a, b, a);                                 // IDP2A
    else if constexpr (OP == 15) { r = __byte_perm(a, b, 0x2140); r = r * 152618u + b; } // PRMT+IMAD pair
    else if constexpr (OP == 16) { r = __byte_perm(a, b, 0x2140); r = __dp4a(r, b, a); } // PRMT+IDP4A pair
    else if constexpr (OP == 17) { r = __viaddmin_s16x2_relu(a, b, 0x01ff01ffu); r = r * 128u + b; } // VIADDMNMX+IMAD
    else if constexpr (OP == 18) r = __vabsdiffu4(a, b) + a;                             // VABSDIFF4.ACC?
    else r = a;
    return r;
}

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = seed * (threadIdx.x + c + 1);
    uint32_t b = seed ^ threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) v[c] = op<OP>(v[c], b);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= v[c];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int ops_per_call, uint32_t *d, int sms, double mhz)
{
    const int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d, 3);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * ITERS * UNROLL * CHAINS * ops_per_call;
    const double clocks = ms * 1e-3 * mhz * 1e6;
    printf("%-22s %8.3f ms  %6.3f warp-instr/clk/SM (counting %d instr per op)\n", name, ms, warp_instr / clocks / sms, ops_per_call);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, nominal %0.f MHz (rates assume the nominal clock)\n", p.name, p.multiProcessorCount, mhz);
    uint32_t *d;
    cudaMalloc(&d, 4096);
    const int sms = p.multiProcessorCount;
    run<11>("IADD3", 1, d, sms, mhz);
    run<0>("PRMT", 1, d, sms, mhz);
    run<1>("LOP3 (x2)", 2, d, sms, mhz);
    run<2>("SHF", 1, d, sms, mhz);
    run<3>("IMAD", 1, d, sms, mhz);
    run<4>("IMAD.HI (+lop)", 1, d, sms, mhz);
    run<12>("x*128", 1, d, sms, mhz);
    run<5>("IDP4A", 1, d, sms, mhz);
    run<14>("IDP2A", 1, d, sms, mhz);
    run<6>("VIADDMNMX.S16x2", 1, d, sms, mhz);
    run<13>("VIMNMX.S16x2", 1, d, sms, mhz);
    run<7>("VIADD.16x2", 1, d, sms, mhz);
    run<8>("VABSDIFF4", 1, d, sms, mhz);
    run<18>("VABSDIFF4+add", 1, d, sms, mhz);
    run<9>("I2IP", 1, d, sms, mhz);
    run<10>("POPC+IADD", 2, d, sms, mhz);
    run<15>("PRMT+IMAD", 2, d, sms, mhz);
    run<16>("PRMT+IDP4A", 2, d, sms, mhz);
    run<17>("VIADDMNMX+IMAD", 2, d, sms, mhz);
    return 0;
}
