// Standalone pipe/shared-memory microbenchmarks for design decisions (not part of the library).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_ffma(float* sink, int iters) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
    const float m = 0.999999f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* sink, int iters) {
    unsigned long long a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float lo = 1.0f + 0.001f * (float)(threadIdx.x + i), hi = lo + 0.5f;
        a[i] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
    }
    const float mf = 0.999999f, cf = 1e-7f;
    const unsigned long long m = ((unsigned long long)__float_as_uint(mf) << 32) | __float_as_uint(mf);
    const unsigned long long c = ((unsigned long long)__float_as_uint(cf) << 32) | __float_as_uint(cf);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123456ull) sink[0] = (float)s;
}

// FFMA2 with different multiplier registers per op (3 distinct 64-bit sources)
__global__ void __launch_bounds__(256) k_ffma2_mix(float* sink, int iters) {
    unsigned long long a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo = 1.0f + 0.001f * (float)(threadIdx.x + i), hi = lo + 0.5f;
        a[i] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
        b[i] = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.9999f + 1e-6f * i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b[i]), "l"(b[(i + 1) & 7]));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123456ull) sink[0] = (float)s;
}

// FFMA + FMNMX(alu) + MUFU mix: 13 FFMA : 1 FMNMX : 1 MUFU (the JSD term recipe)
__global__ void __launch_bounds__(256) k_mix(float* sink, int iters) {
    float a[8], mx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
    const float m = 0.999999f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float r;
            asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i]));
            float t = a[i];
#pragma unroll
            for (int q = 0; q < 12; ++q) t = fmaf(t, m, c);
            mx = fmaxf(mx, t);
            a[i] = fmaf(t, r, c);
        }
    }
    float s = mx;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;
}

// shared-memory atomics: per-warp histogram copies, indices from a rolling 2-bit stream
template <int MODE>  // 0 = no histogram op (loop cost), 1 = atomicAdd (RED), 2 = private byte counters
__global__ void __launch_bounds__(128) k_hist(unsigned* out, int iters, int bins_log2, int per_warp) {
    extern __shared__ unsigned sm[];
    const int bins = 1 << bins_log2;
    const int tid = threadIdx.x;
    if (MODE == 2) { for (int i = tid; i < bins * 128 / 4; i += 128) sm[i] = 0; }
    else { for (int i = tid; i < bins * (per_warp ? 4 : 1); i += 128) sm[i] = 0; }
    __syncthreads();
    unsigned* h = sm + (per_warp ? (tid >> 5) * bins : 0);
    unsigned char* hb = reinterpret_cast<unsigned char*>(sm);
    unsigned x = 0x9E3779B9u * (blockIdx.x * 128 + tid + 1);
    unsigned w = x, acc = 0;
    const unsigned mask = bins - 1;
    for (int it = 0; it < iters; ++it) {
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        unsigned bits = x;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            w = (w << 2) | (bits & 3u);
            bits >>= 2;
            const unsigned idx = w & mask;
            if (MODE == 0) acc ^= idx + j;
            if (MODE == 1) atomicAdd(&h[idx], 1u);
            if (MODE == 2) { unsigned char* p = hb + idx * 128 + tid; *p = (unsigned char)(*p + 1); }
        }
    }
    __syncthreads();
    unsigned s = acc;
    if (MODE == 2) { for (int i = tid; i < bins * 128 / 4; i += 128) s += sm[i]; }
    else { for (int i = tid; i < bins * (per_warp ? 4 : 1); i += 128) s += sm[i]; }
    if (s == 0x12345u) out[0] = s;
}

static double time_ms(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch(arg);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

static float* g_sink; static int g_sms; static int g_iters;
int main(int argc, char** argv) {
    CK(cudaMalloc(&g_sink, 1024));
    CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max clock %d kHz\n", g_sms, clk);
    const int blocks = g_sms * 8;
    g_iters = 2048;
    auto rate = [&](double ms, double ops_per_thread_iter, int threads) {
        return (double)blocks * threads * g_iters * ops_per_thread_iter / (ms * 1e-3) / 1e12;
    };
    double ms;
    ms = time_ms([](void*) { k_ffma<<<g_sms * 8, 256>>>(g_sink, g_iters); }, 0, 5);
    printf("FFMA burst      : %.2f T FFMA/s  (%.1f TFLOP/s)  %.3f ms\n", rate(ms, 64, 256), 2 * rate(ms, 64, 256), ms);
    ms = time_ms([](void*) { k_ffma2<<<g_sms * 8, 256>>>(g_sink, g_iters); }, 0, 5);
    printf("FFMA2 burst     : %.2f T FFMA2/s (%.1f TFLOP/s)  %.3f ms\n", rate(ms, 64, 256), 4 * rate(ms, 64, 256), ms);
    ms = time_ms([](void*) { k_ffma2_mix<<<g_sms * 8, 256>>>(g_sink, g_iters); }, 0, 5);
    printf("FFMA2 3-src     : %.2f T FFMA2/s (%.1f TFLOP/s)  %.3f ms\n", rate(ms, 64, 256), 4 * rate(ms, 64, 256), ms);
    ms = time_ms([](void*) { k_mix<<<g_sms * 8, 256>>>(g_sink, g_iters); }, 0, 5);
    printf("JSD-like mix    : %.3f T terms/s (13 FFMA+1 FMNMX+1 MUFU per term) %.3f ms\n", rate(ms, 8, 256), ms);

    // sustained FFMA: back-to-back launches for ~4 s, rate per 0.5 s window
    {
        g_iters = 8192;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        double t_acc = 0; int win = 0;
        while (t_acc < 4000.0) {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 8; ++i) k_ffma<<<g_sms * 8, 256>>>(g_sink, g_iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float t; CK(cudaEventElapsedTime(&t, e0, e1));
            t_acc += t;
            if (t_acc > 500.0 * (win + 1)) {
                ++win;
                printf("FFMA sustained t=%.1fs : %.1f TFLOP/s\n", t_acc / 1e3, 2.0 * 8 * blocks * 256.0 * g_iters * 64 / (t * 1e-3) / 1e12);
            }
        }
        t_acc = 0; win = 0;
        while (t_acc < 3000.0) {
            CK(cudaEventRecord(e0));
            for (int i = 0; i < 8; ++i) k_mix<<<g_sms * 8, 256>>>(g_sink, g_iters / 4);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float t; CK(cudaEventElapsedTime(&t, e0, e1));
            t_acc += t;
            if (t_acc > 500.0 * (win + 1)) {
                ++win;
                printf("mix sustained t=%.1fs : %.3f T terms/s\n", t_acc / 1e3, 8.0 * blocks * 256.0 * (g_iters / 4) * 8 / (t * 1e-3) / 1e12);
            }
        }
        g_iters = 2048;
    }

    // histogram updates
    unsigned* out = (unsigned*)g_sink;
    CK(cudaFuncSetAttribute(k_hist<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaFuncSetAttribute(k_hist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaFuncSetAttribute(k_hist<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    const int hb = g_sms * 16, hiters = 512;
    for (int bl = 8; bl <= 12; bl += 2) {
        for (int pw = 0; pw < 2; ++pw) {
            const size_t smem = (size_t)(4u << bl) * (pw ? 4 : 1);
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float t0 = 0, t1 = 0;
            for (int r = 0; r < 3; ++r) {
                CK(cudaEventRecord(e0)); k_hist<0><<<hb, 128, smem>>>(out, hiters, bl, pw); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&t0, e0, e1));
                CK(cudaEventRecord(e0)); k_hist<1><<<hb, 128, smem>>>(out, hiters, bl, pw); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&t1, e0, e1));
            }
            const double n = (double)hb * 128 * hiters * 16;
            printf("hist bins=%5d per_warp=%d : loop %.3f ms, +atomics %.3f ms -> %.1f G upd/s total, %.2f upd/clk/SM (at 1.9GHz)\n",
                   1 << bl, pw, t0, t1, n / (t1 * 1e-3) / 1e9, n / (t1 * 1e-3) / g_sms / 1.9e9);
        }
    }
    {   // private byte counters, 256 bins x 128 threads = 32 KB
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float t2 = 0;
        for (int r = 0; r < 3; ++r) {
            CK(cudaEventRecord(e0)); k_hist<2><<<hb, 128, 256 * 128>>>(out, hiters, 8, 0); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&t2, e0, e1));
        }
        const double n = (double)hb * 128 * hiters * 16;
        printf("private u8 counters bins=256 : %.3f ms -> %.1f G upd/s, %.2f upd/clk/SM\n", t2, n / (t2 * 1e-3) / 1e9,
               n / (t2 * 1e-3) / g_sms / 1.9e9);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
