// Microbenchmark of the JSD per-dimension term recipe: scalar FP32 vs packed f32x2 (FFMA2), register resident.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/ubench_jsd tools/ubench_jsd.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#define G7 5.809747504e-02f
#define G6 -4.418099709e-02f
#define G5 4.228754936e-02f
#define G4 1.528030711e-02f
#define G3 3.664686569e-02f
#define G2 6.660644403e-02f
#define G1 1.666681249e-01f
#define G0 9.999999943e-01f

__device__ __forceinline__ float jsd_G(float u) {
    float G = G7;
    G = fmaf(G, u, G6); G = fmaf(G, u, G5); G = fmaf(G, u, G4); G = fmaf(G, u, G3);
    G = fmaf(G, u, G2); G = fmaf(G, u, G1); G = fmaf(G, u, G0);
    return G;
}

__global__ void __launch_bounds__(128, 4) k_scalar(float* sink, int iters) {
    float a[4], b[4], c[4][4];
    for (int i = 0; i < 4; ++i) { a[i] = 0.003f + 1e-4f * (threadIdx.x + i); b[i] = 0.004f + 1e-4f * ((threadIdx.x * 7 + i) & 31); }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    float umax = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = a[i] - b[j];
                const float x = d * rcp_approx(a[i] + b[j]);
                const float u = x * x;
                umax = fmaxf(umax, u);
                c[i][j] = fmaf(d * x, jsd_G(u), c[i][j]);
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] += 1e-7f; b[i] += 2e-7f; }
    }
    float s = umax;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456f) sink[0] = s;
}

template <int ORDER>
__global__ void __launch_bounds__(128, 4) k_packed(float* sink, int iters) {
    float a[4], b[4];
    u64 c[4][2];
    for (int i = 0; i < 4; ++i) { a[i] = 0.003f + 1e-4f * (threadIdx.x + i); b[i] = 0.004f + 1e-4f * ((threadIdx.x * 7 + i) & 31); }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) c[i][j] = 0ull;
    const u64 g7 = pk(G7, G7), g6 = pk(G6, G6), g5 = pk(G5, G5), g4 = pk(G4, G4), g3 = pk(G3, G3), g2 = pk(G2, G2),
              g1 = pk(G1, G1), g0 = pk(G0, G0);
    float umax = 0.f;
    for (int it = 0; it < iters; ++it) {
        u64 a2[4], b2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) a2[i] = pk(a[i], a[i]);
        b2[0] = pk(b[0], b[1]);
        b2[1] = pk(b[2], b[3]);
        if (ORDER == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const u64 s = add2(a2[i], b2[j]);
                    const u64 d = sub2(a2[i], b2[j]);
                    float s0, s1; upk(s, s0, s1);
                    const u64 r = pk(rcp_approx(s0), rcp_approx(s1));
                    const u64 x = mul2(d, r);
                    const u64 u = mul2(x, x);
                    float u0, u1; upk(u, u0, u1);
                    umax = fmaxf(umax, fmaxf(u0, u1));
                    u64 G = fma2(g7, u, g6);
                    G = fma2(G, u, g5); G = fma2(G, u, g4); G = fma2(G, u, g3);
                    G = fma2(G, u, g2); G = fma2(G, u, g1); G = fma2(G, u, g0);
                    c[i][j] = fma2(mul2(d, x), G, c[i][j]);
                }
        } else {
            // coefficient-major Horner: the constant operand repeats over consecutive instructions
            u64 d[8], x[8], u[8], G[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = i * 2 + j;
                    const u64 s = add2(a2[i], b2[j]);
                    d[k] = sub2(a2[i], b2[j]);
                    float s0, s1; upk(s, s0, s1);
                    const u64 r = pk(rcp_approx(s0), rcp_approx(s1));
                    x[k] = mul2(d[k], r);
                    u[k] = mul2(x[k], x[k]);
                    float u0, u1; upk(u[k], u0, u1);
                    umax = fmaxf(umax, fmaxf(u0, u1));
                }
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(g7, u[k], g6);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g5);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g4);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g3);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g2);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g1);
#pragma unroll
            for (int k = 0; k < 8; ++k) G[k] = fma2(G[k], u[k], g0);
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k >> 1][k & 1] = fma2(mul2(d[k], x[k]), G[k], c[k >> 1][k & 1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] += 1e-7f; b[i] += 2e-7f; }
    }
    float s = umax;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) { float lo, hi; upk(c[i][j], lo, hi); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}

int main() {
    float* sink; CK(cudaMalloc(&sink, 64));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = sms * 16, iters = 4096;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int v = 0; v < 3; ++v) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0));
            if (v == 0) k_scalar<<<blocks, 128>>>(sink, iters);
            if (v == 1) k_packed<0><<<blocks, 128>>>(sink, iters);
            if (v == 2) k_packed<1><<<blocks, 128>>>(sink, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        const double terms = (double)blocks * 128 * iters * 16;
        printf("%s : %.3f ms, %.3f T terms/s -> %.2f G pairs/s at D=256\n",
               v == 0 ? "scalar FP32     " : v == 1 ? "packed pair-major" : "packed coef-major", best, terms / (best * 1e-3) / 1e12,
               terms / (best * 1e-3) / 256 / 1e9);
    }
    return 0;
}
