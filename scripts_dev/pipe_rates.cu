// Dispatch cost of the instruction classes the Philox rollout kernel is made of (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipe_rates scripts_dev/pipe_rates.cu && ./pipe_rates
// Every kernel runs 148 x 4 CTAs of 256 threads (8 warps per SM sub-partition) through ITER iterations of 8
// independent dependency chains x 4 instructions of one class (or a mix), and reports SM sub-partition cycles per
// warp-instruction = time x clock x (148 x 4) / warp-instructions.  1.0 = one instruction per cycle per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

#define CHAINS8(OP) OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

template <int KIND>
__global__ void __launch_bounds__(256) k(float *out, float seed, uint32_t useed)
{
    float f[8];
    uint32_t u[8], v[8];
    unsigned long long w[8], fp[8];
    for (int i = 0; i < 8; i++) {
        f[i] = seed + i + threadIdx.x * 1e-3f;
        u[i] = useed + i * 77u + threadIdx.x;
        v[i] = useed * 3u + i;
        w[i] = u[i];
        fp[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f);
    }
    const float a = seed * 0.999f, b = seed * 1e-3f;
    const unsigned long long ab = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (KIND == 0) {          // FFMA
#define OP(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 1) {   // FFMA2
#define OP(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(fp[i]) : "l"(ab), "l"(bb));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 2) {   // IMAD.WIDE
#define OP(i) asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[i]) : "r"(u[i]));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 3) {   // MUFU.EX2
#define OP(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 4) {   // LOP3
#define OP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(useed));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 9) {   // LOP3 with two register sources and an immediate (the Philox round-key form)
#define OP(i) asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(u[i]) : "r"(v[i]));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 10) {  // FMUL
#define OP(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(a));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 11) {  // shift-or (bits -> [1,2) float), LEA.HI / SHF + LOP3
#define OP(i) asm volatile("{ .reg .b32 t; shr.b32 t, %0, 9; or.b32 %0, t, 0x3f800000; }" : "+r"(u[i]));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 5) {   // FFMA + LOP3 interleaved (different pipes)
#define OP(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(b)); \
              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(useed));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 6) {   // MUFU + 3 FFMA interleaved
#define OP(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); \
              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + 1) & 7]) : "f"(a), "f"(b)); \
              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + 2) & 7]) : "f"(a), "f"(b)); \
              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + 3) & 7]) : "f"(a), "f"(b));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 7) {   // IMAD.WIDE + 2 LOP3 (a Philox half-round)
#define OP(i) asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[i]) : "r"(u[i])); \
              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(useed)); \
              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(u[i]), "r"(useed));
                CHAINS8(OP)
#undef OP
            } else if (KIND == 8) {   // MUFU + IMAD.WIDE + LOP3 + FFMA2: the Philox kernel's mix in miniature
#define OP(i) asm volatile("mad.wide.u32 %0, %1, 0xD2511F53, %0;" : "+l"(w[i]) : "r"(u[i])); \
              asm volatile("mad.wide.u32 %0, %1, 0xCD9E8D57, %0;" : "+l"(w[(i + 1) & 7]) : "r"(v[i])); \
              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(useed)); \
              asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(u[i]), "r"(useed)); \
              asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(fp[i]) : "l"(ab), "l"(bb)); \
              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i + 2) & 7]) : "f"(a), "f"(b));
                CHAINS8(OP)
#undef OP
                if ((r & 1) == 0) {
#define OP(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                    CHAINS8(OP)
#undef OP
                }
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += f[i] + (float)u[i] + (float)v[i] + (float)(w[i] >> 7) + __uint_as_float((uint32_t)fp[i]);
    if (s == 12345.678f) out[0] = s;
}

template <int KIND>
static void run(const char *name, double instr_per_iter, float *out, double clock_hz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 4;
    k<KIND><<<grid, 256>>>(out, 1.0001f, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND><<<grid, 256>>>(out, 1.0001f, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * 8 * ITER * instr_per_iter;
    const double smsp_cycles = ms * 1e-3 * clock_hz * 148 * 4;
    printf("%-46s %8.3f ms   %6.3f SMSP-cycles per warp-instruction\n", name, ms, smsp_cycles / warp_instr);
}

int main()
{
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double hz = khz * 1e3;
    printf("SM clock (max) %.0f MHz; 8 warps per sub-partition, 8 independent chains per thread\n", hz / 1e6);
    float *out;
    cudaMalloc(&out, 4);
    run<0>("FFMA", 32, out, hz);
    run<1>("FFMA2", 32, out, hz);
    run<2>("IMAD.WIDE.U32", 32, out, hz);
    run<3>("MUFU.EX2", 32, out, hz);
    run<4>("LOP3", 32, out, hz);
    run<9>("LOP3 (2 registers + immediate)", 32, out, hz);
    run<10>("FMUL", 32, out, hz);
    run<11>("shr + or (one or two instructions)", 32, out, hz);
    run<5>("FFMA + LOP3 (1:1)", 64, out, hz);
    run<6>("MUFU + 3 FFMA", 128, out, hz);
    run<7>("IMAD.WIDE + 2 LOP3", 96, out, hz);
    run<8>("2 IMAD.WIDE + 2 LOP3 + FFMA2 + FFMA + 0.5 MUFU", 6 * 32 + 16, out, hz);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
