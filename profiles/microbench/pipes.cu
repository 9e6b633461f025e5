// pipes.cu -- issue-rate microbenchmarks that decide the prefill epilogue design on B200:
// cycles per warp-instruction per SM sub-partition for the candidate fold instructions, and the
// legacy mma.sync s8 rate.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 200
#define N 16   // independent chains per thread

template <int MODE>
__global__ void k(float* out, const float* in, long long* cyc, int warps) {
    float a[N], b[N], c[N];
    int xi[N];
    for (int i = 0; i < N; i++) { a[i] = in[i + threadIdx.x]; b[i] = in[64 + i + threadIdx.x]; c[i] = in[128 + i]; xi[i] = (int)in[200 + i]; }
    const float s0 = in[300 + (threadIdx.x & 1)], s1 = in[301 + (threadIdx.x & 1)];
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (MODE == 0) a[i] = fmaf(b[i], c[i], a[i]);                 // FFMA, 3 distinct regs
                if (MODE == 1) a[i] = fmaf(s0, a[i], s1);                     // FFMA, 2 regs shared by all (reuse)
                if (MODE == 2) a[i] = fmaf(s0, b[i], a[i]);                   // FFMA, 1 shared + 2 distinct
                if (MODE == 3) a[i] = a[i] * s0;                              // FMUL 2 regs
                if (MODE == 4) a[i] = __int2float_rn(xi[i]) + a[i];           // I2FP + FADD
                if (MODE == 5) xi[i] = xi[i] + 0x4B400000;                    // IADD imm
                if (MODE == 6) a[i] = a[i] + -12582912.0f;                    // FADD imm
                if (MODE == 9) a[i] = __int2float_rn(xi[i] + it);             // IADD + I2FP only
            }
            if (MODE == 7 || MODE == 8 || MODE == 10 || MODE == 11) {
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    uint64_t A, B, C, D;
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(A) : "f"(a[i]), "f"(a[i + 1]));
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b[i]), "f"(b[i + 1]));
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(c[i]), "f"(c[i + 1]));
                    uint64_t S, S2;
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(S) : "f"(s0), "f"(s0));
                    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(S2) : "f"(s1), "f"(s1));
                    if (MODE == 7) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(B), "l"(C), "l"(A));   // FFMA2 3 distinct
                    if (MODE == 8) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(S), "l"(A), "l"(S2));  // FFMA2 scalar,reg,scalar
                    if (MODE == 10) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(S), "l"(B), "l"(A));  // FFMA2 scalar + 2 distinct
                    if (MODE == 11) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(S));              // FADD2
                    asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(D));
                }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < N; i++) acc += a[i] + b[i] + (float)xi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// legacy tensor path: mma.sync.m16n8k32 s32 += u8 x s8, independent accumulators
__global__ void k_imma(int* out, const int* in, long long* cyc) {
    int a0 = in[threadIdx.x], a1 = in[32 + threadIdx.x], a2 = in[64 + threadIdx.x], a3 = in[96 + threadIdx.x];
    int b0 = in[128 + threadIdx.x], b1 = in[160 + threadIdx.x];
    int c[8][4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    int acc = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float *in, *out; long long* cyc; int *iin, *iout;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    cudaMalloc(&iin, 4096 * 4); cudaMalloc(&iout, 1 << 20);
    cudaMemset(in, 0, 4096 * 4); cudaMemset(iin, 1, 4096 * 4);
    const char* names[] = {"FFMA 3 distinct regs", "FFMA scalar,reg,scalar", "FFMA scalar+2 distinct", "FMUL reg*scalar",
                           "I2FP+FADD", "IADD imm", "FADD imm", "FFMA2 3 distinct", "FFMA2 scalar,reg,scalar",
                           "IADD+I2FP", "FFMA2 scalar+2 distinct", "FADD2 reg+scalar"};
    for (int warps : {4, 8, 16}) {
        printf("---- %d warps per SM (1 CTA), cycles per warp-instruction per sub-partition ----\n", warps);
        for (int mode = 0; mode < 12; mode++) {
            long long h = 0;
            auto run = [&](auto kern) { kern<<<1, warps * 32>>>(out, in, cyc, warps); kern<<<1, warps * 32>>>(out, in, cyc, warps); };
            switch (mode) {
            case 0: run(k<0>); break; case 1: run(k<1>); break; case 2: run(k<2>); break; case 3: run(k<3>); break;
            case 4: run(k<4>); break; case 5: run(k<5>); break; case 6: run(k<6>); break; case 7: run(k<7>); break;
            case 8: run(k<8>); break; case 9: run(k<9>); break; case 10: run(k<10>); break; case 11: run(k<11>); break;
            }
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double per_thread_instr = (double)ITERS * 4 * N * ((mode == 4 || mode == 9) ? 2 : 1) / ((mode == 7 || mode == 8 || mode >= 10) ? 2 : 1);
            double warp_instr_per_smsp = per_thread_instr * warps / 4.0;
            printf("%-26s %8.3f cyc/warp-instr/SMSP   (%lld cycles)\n", names[mode], h / warp_instr_per_smsp, h);
        }
        long long h = 0;
        k_imma<<<1, warps * 32>>>(iout, iin, cyc); k_imma<<<1, warps * 32>>>(iout, iin, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double mmas = (double)ITERS * 4 * 8 * warps;
        printf("%-26s %8.3f cyc/mma/SM -> %.0f MAC/clk/SM   (%lld cycles)\n", "mma.sync m16n8k32 u8s8", h / mmas, mmas * 4096.0 / h, h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
