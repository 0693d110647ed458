// Instruction-throughput microbenchmarks for the integer / tensor paths the ladder kernel can use (sm_100a).
// Each test runs ITER iterations of UNR independent chains per thread; prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int OP>
__device__ __forceinline__ void body(uint32_t (&a)[8], uint32_t b, uint32_t c)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 1) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 3) asm volatile("dp2a.lo.s32.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 4) asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 5) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (OP == 7) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c & 31));
        if (OP == 8) { // a + (b >> 7): LEA.HI candidate
            a[k] = a[k] + (b >> 7);
            asm volatile("" : "+r"(a[k]));
        }
        if (OP == 9) { // masked accumulate: (a & 0xffff0000) + b*c   -> LOP3 + IMAD
            a[k] = (a[k] & 0xffff0000u) + b * c;
            asm volatile("" : "+r"(a[k]));
        }
        if (OP == 10) { // cvt.pack.sat.s16.s32
            asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
        }
        if (OP == 11) { // arithmetic shift right
            asm volatile("shr.s32 %0, %0, 3;" : "+r"(a[k]));
        }
        if (OP == 12) { // vabsdiff4
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
        }
        if (OP == 13) { // mul.wide
            uint64_t w;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[k]), "r"(b));
            a[k] = (uint32_t)(w >> 32) + (uint32_t)w;
        }
        if (OP == 14) { // hi-half pmulhw via mask + mad.hi : LOP3 + IMAD.HI
            uint32_t x = b & 0xffff0000u;
            asm volatile("" : "+r"(x));
            asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(a[k]) : "r"(x), "r"(c));
        }
        if (OP == 16) { // IDP + LOP3 pair
            asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 17) { // IDP + IMAD pair
            asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 18) { // VABSDIFF4 + LOP3 pair
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 19) { // VABSDIFF4 + IMAD pair
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 20) { // scalar sad (VABSDIFF) + LOP3
            asm volatile("sad.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 21) { // scalar sad + IMAD
            asm volatile("sad.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 22) { // PRMT + IMAD
            asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 23) { // SHF + LOP3 (both alu?)
            asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c & 31));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[(k + 4) & 7]) : "r"(b), "r"(c));
        }
        if (OP == 15) { // vimnmx s16x2
            asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
        }
    }
}

template <int OP>
__global__ void k_alu(uint32_t *out, uint32_t b, uint32_t c, long long *cyc)
{
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) body<OP>(a, b, c);
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// IMMA m16n8k32: NACC independent accumulators per warp
template <int NACC, int KIND>
__global__ void k_mma(uint32_t *out, uint32_t b, long long *cyc)
{
    int d[NACC][4];
#pragma unroll
    for (int k = 0; k < NACC; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = b, a3 = b + 1, b0 = b * 5, b1 = b * 7;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[k][0]), "+r"(d[k][1]), "+r"(d[k][2]), "+r"(d[k][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(d[k][0]), "+r"(d[k][1]), "+r"(d[k][2]), "+r"(d[k][3])
                             : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared loads: MODE 0 LDS.128 (8 x s16), 1 LDS.S16, 2 LDS.32, 3 ldmatrix.x4
template <int MODE>
__global__ void k_lds(uint32_t *out, long long *cyc)
{
    __shared__ __align__(16) int16_t sm[8192 + 64];
    for (int i = threadIdx.x; i < 8192 + 64; i += blockDim.x) sm[i] = (int16_t)i;
    __syncthreads();
    uint32_t s = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int base = ((i * 8 + k) * 72 + w * 264) & 4095;
            if (MODE == 0) {
                uint4 q = *reinterpret_cast<const uint4 *>(&sm[(base & ~7) + lane * 8]);
                s += q.x ^ q.y ^ q.z ^ q.w;
            } else if (MODE == 1) {
                s += (int)sm[base + lane];
            } else if (MODE == 2) {
                s += *reinterpret_cast<const uint32_t *>(&sm[(base & ~1) + lane * 2]);
            } else {
                uint32_t r0, r1, r2, r3;
                // rows of 16 bytes, pitch 144 bytes (72 int16): conflict-free
                uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[((base & ~7) & 1023) + lane * 72]);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
                s += r0 ^ r1 ^ r2 ^ r3;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static uint32_t *d_out;
static long long *d_cyc;

template <typename F>
static void report(const char *name, int threads, int instr_per_iter, F launch)
{
    launch();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[148];
    cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)cyc[i];
    avg /= 148;
    const double warps_per_smsp = threads / 32 / 4.0;
    const double per = avg / ((double)ITER * instr_per_iter * warps_per_smsp);
    printf("%-34s threads/SM %4d  cycles %10.0f  cyc/warp-instr/SMSP %6.3f  (%.3f ms) %s\n", name, threads, avg, per, ms,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

#define RUN_ALU(OP, NAME, T) report(NAME, T, 8, [&] { k_alu<OP><<<148, T>>>(d_out, b, c, d_cyc); })

int main()
{
    cudaMalloc(&d_out, 148 * 1024 * 4);
    cudaMalloc(&d_cyc, 148 * 8);
    uint32_t b = 0x01020304u, c = 0x00030001u;
    for (int T : {512}) {
        RUN_ALU(0, "IMAD (mad.lo)", T);
        RUN_ALU(1, "IMAD.HI (mad.hi)", T);
        RUN_ALU(2, "LOP3", T);
        RUN_ALU(3, "IDP.2A", T);
        RUN_ALU(4, "IDP.4A", T);
        RUN_ALU(5, "I2IP u8 pack.sat", T);
        RUN_ALU(6, "PRMT", T);
        RUN_ALU(7, "SHF", T);
        RUN_ALU(8, "a + (b>>7)", T);
        RUN_ALU(9, "masked acc (LOP3+IMAD) x1", T);
        RUN_ALU(10, "I2IP s16 pack.sat", T);
        RUN_ALU(11, "SHR.S32", T);
        RUN_ALU(12, "VABSDIFF4.add", T);
        RUN_ALU(13, "mul.wide + add", T);
        RUN_ALU(14, "LOP3 + IMAD.HI", T);
        RUN_ALU(15, "VIMNMX.S16x2", T);
        RUN_ALU(16, "pair IDP+LOP3 (x2 instr)", T);
        RUN_ALU(17, "pair IDP+IMAD", T);
        RUN_ALU(18, "pair VABSDIFF4+LOP3", T);
        RUN_ALU(19, "pair VABSDIFF4+IMAD", T);
        RUN_ALU(20, "pair SAD+LOP3", T);
        RUN_ALU(21, "pair SAD+IMAD", T);
        RUN_ALU(22, "pair PRMT+IMAD", T);
        RUN_ALU(23, "pair SHF+LOP3", T);
        report("IMMA.16832 u8.s8 x4acc", T, 4, [&] { k_mma<4, 0><<<148, T>>>(d_out, b, d_cyc); });
        report("IMMA.16832 u8.s8 x8acc", T, 8, [&] { k_mma<8, 0><<<148, T>>>(d_out, b, d_cyc); });
        report("IMMA.16816 u8.s8 x8acc", T, 8, [&] { k_mma<8, 1><<<148, T>>>(d_out, b, d_cyc); });
        report("LDS.128", T, 8, [&] { k_lds<0><<<148, T>>>(d_out, d_cyc); });
        report("LDS.S16", T, 8, [&] { k_lds<1><<<148, T>>>(d_out, d_cyc); });
        report("LDS.32", T, 8, [&] { k_lds<2><<<148, T>>>(d_out, d_cyc); });
        report("LDSM.x4", T, 8, [&] { k_lds<3><<<148, T>>>(d_out, d_cyc); });
    }
    return 0;
}
