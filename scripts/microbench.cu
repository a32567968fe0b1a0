// Micro-benchmarks for the roofline denominators of the swarm kernels (run on the B200 box):
// MUFU (XU pipe) / FFMA / SHFL issue rates per SM, FP64 add latency, write-only and copy HBM bandwidth.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench scripts/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int OP>
__global__ void k_pipe(float* out, float seed) {
    float v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = seed + 0.001f * (threadIdx.x + c);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
            if (OP == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[c]) : "f"(seed));
            if (OP == 4) v[c] = __shfl_sync(0xffffffffu, v[c], (threadIdx.x + 1) & 31);
            if (OP == 5) {   // the pair-force mix: 4 MUFU + 15 FP32
                float r2 = fmaf(v[c], v[c], 1e-30f), ri, e1, e2, iv;
                asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"(r2));
                float r = r2 * ri;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-r));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(r * -0.1f));
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(iv) : "f"(r + 1e-6f));
                v[c] = fmaf(fmaf(0.5f, e2, -e1) * iv, v[c], seed);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;
}

// packed FP32 pairs (sm_100+): one issue slot, two FP32 results
template <int OP>
__global__ void k_pipe2(float* out, float seed) {
    unsigned long long v[CHAINS], sd;
    float2 s2 = make_float2(seed, seed * 0.5f);
    sd = *reinterpret_cast<unsigned long long*>(&s2);
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        float2 t = make_float2(seed + 0.001f * (threadIdx.x + c), seed);
        v[c] = *reinterpret_cast<unsigned long long*>(&t);
    }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v[c]) : "l"(sd));
            if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(sd));
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { float2 t = *reinterpret_cast<float2*>(&v[c]); s += t.x + t.y; }
    if (s == 123.456f) out[0] = s;
}

// FP64 -> FP32 conversion: which pipe?  OP 0: DADD + cvt.rn.f32.f64 chain; OP 1: MUFU.EX2 chain; OP 2: both interleaved
template <int OP>
__global__ void k_cvt(float* out, double seed) {
    double d[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c] = seed + 0.001 * (threadIdx.x + c); f[c] = 0.5f + 0.001f * c; }
    for (int i = 0; i < ITERS / 4; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0 || OP == 2) {
                float t;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[c]));
                d[c] = __dadd_rn(d[c], (double)1e-9);
                f[c] += t * 1e-30f;
            }
            if (OP == 1 || OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c] + (float)d[c];
    if (s == 123.456f) out[0] = s;
}

template <int OP>
__global__ void k_pipe64(double* out, double seed) {
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = seed + 0.001 * (threadIdx.x + c);
    for (int i = 0; i < ITERS / 8; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) v[c] = __dadd_rn(v[c], seed);
            if (OP == 1) v[c] = __fma_rn(v[c], seed, seed);
            if (OP == 2) v[c] = (double)(float)v[c] + 0.0;        // F2F.F32.F64 + F2F.F64.F32
            if (OP == 3) v[c] = (double)(int)v[c];                // F2I + I2F
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    if (s == 123.456) out[0] = s;
}

__global__ void k_dadd_latency(double* out, long long* cyc, double seed) {
    double s = seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) s = __dadd_rn(s, seed);
    long long t1 = clock64();
    out[0] = s;
    cyc[0] = t1 - t0;
}

__global__ void k_fill(float4* p, size_t n4, int streaming) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (streaming) __stcs(p + i, z); else p[i] = z;
    }
}

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, clk_khz);
    float* out;
    CK(cudaMalloc(&out, 1024));
    const char* names[] = {"mufu_ex2", "mufu_rsq", "mufu_rcp", "ffma", "shfl", "pair_mix"};
    const double per_iter[] = {1, 1, 1, 1, 1, 1};
    for (int op = 0; op < 6; ++op) {
        const int blocks = sms * 8, threads = 256;
        float ms = 0;
        switch (op) {
            case 0: ms = time_ms([&] { k_pipe<0><<<blocks, threads>>>(out, 0.5f); }); break;
            case 1: ms = time_ms([&] { k_pipe<1><<<blocks, threads>>>(out, 0.5f); }); break;
            case 2: ms = time_ms([&] { k_pipe<2><<<blocks, threads>>>(out, 0.5f); }); break;
            case 3: ms = time_ms([&] { k_pipe<3><<<blocks, threads>>>(out, 0.5f); }); break;
            case 4: ms = time_ms([&] { k_pipe<4><<<blocks, threads>>>(out, 0.5f); }); break;
            case 5: ms = time_ms([&] { k_pipe<5><<<blocks, threads>>>(out, 0.5f); }); break;
        }
        const double ops = (double)blocks * threads * ITERS * CHAINS * per_iter[op];
        const double per_s = ops / (ms * 1e-3);
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"thread_ops_per_s\": %.4e, \"per_clk_per_sm_at_1965MHz\": %.2f}\n",
               names[op], ms, per_s, per_s / sms / 1.965e9);
    }
    {
        const int blocks = sms * 8, threads = 256;
        float ms = time_ms([&] { k_pipe2<0><<<blocks, threads>>>(out, 0.5f); });
        double per_s = (double)blocks * threads * ITERS * CHAINS / (ms * 1e-3);
        printf("{\"bench\": \"ffma2\", \"ms\": %.4f, \"thread_inst_per_s\": %.4e, \"inst_per_clk_per_sm_at_1965MHz\": %.2f}\n", ms, per_s, per_s / sms / 1.965e9);
        ms = time_ms([&] { k_pipe2<1><<<blocks, threads>>>(out, 0.5f); });
        per_s = (double)blocks * threads * ITERS * CHAINS / (ms * 1e-3);
        printf("{\"bench\": \"fadd2\", \"ms\": %.4f, \"thread_inst_per_s\": %.4e, \"inst_per_clk_per_sm_at_1965MHz\": %.2f}\n", ms, per_s, per_s / sms / 1.965e9);
    }
    {
        const int blocks = sms * 8, threads = 256;
        const char* nm[] = {"cvt_f32_f64+dadd", "mufu_ex2_quarter", "cvt_f32_f64+dadd+mufu_ex2"};
        for (int op = 0; op < 3; op += 2) {
            float ms = 0;
            if (op == 0) ms = time_ms([&] { k_cvt<0><<<blocks, threads>>>(out, 0.5); });
            if (op == 2) ms = time_ms([&] { k_cvt<2><<<blocks, threads>>>(out, 0.5); });
            const double per_s = (double)blocks * threads * (ITERS / 4) * CHAINS / (ms * 1e-3);
            printf("{\"bench\": \"%s\", \"ms\": %.4f, \"iters_per_clk_per_sm_at_1965MHz\": %.2f}\n", nm[op], ms, per_s / sms / 1.965e9);
        }
    }
    {
        double* d;
        CK(cudaMalloc(&d, 1024));
        const char* n64[] = {"dadd", "dfma", "cvt_f64_f32_roundtrip", "cvt_f64_i32_roundtrip"};
        for (int op = 0; op < 4; ++op) {
            const int blocks = sms * 8, threads = 256;
            float ms = 0;
            switch (op) {
                case 0: ms = time_ms([&] { k_pipe64<0><<<blocks, threads>>>(d, 0.5); }); break;
                case 1: ms = time_ms([&] { k_pipe64<1><<<blocks, threads>>>(d, 0.5); }); break;
                case 2: ms = time_ms([&] { k_pipe64<2><<<blocks, threads>>>(d, 0.5); }); break;
                case 3: ms = time_ms([&] { k_pipe64<3><<<blocks, threads>>>(d, 0.5); }); break;
            }
            const double ops = (double)blocks * threads * (ITERS / 8) * CHAINS;
            const double per_s = ops / (ms * 1e-3);
            printf("{\"bench\": \"%s\", \"ms\": %.4f, \"thread_ops_per_s\": %.4e, \"per_clk_per_sm_at_1965MHz\": %.2f}\n",
                   n64[op], ms, per_s, per_s / sms / 1.965e9);
        }
    }
    {
        double* d; long long* c;
        CK(cudaMalloc(&d, 8)); CK(cudaMalloc(&c, 8));
        k_dadd_latency<<<1, 1>>>(d, c, 1.0);
        long long h;
        CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
        printf("{\"bench\": \"dadd_dependent_latency\", \"cycles_per_add\": %.2f}\n", h / 1024.0);
    }
    {
        const size_t bytes = (size_t)2 << 30;
        float4 *a, *b;
        CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        const size_t n4 = bytes / 16;
        for (int streaming = 0; streaming < 2; ++streaming) {
            float ms = time_ms([&] { k_fill<<<sms * 16, 256>>>(a, n4, streaming); });
            printf("{\"bench\": \"fill_%s\", \"gbs\": %.1f}\n", streaming ? "stcs" : "st", bytes / (ms * 1e-3) / 1e9);
        }
        float ms = time_ms([&] { cudaMemsetAsync(a, 0, bytes); });
        printf("{\"bench\": \"cudaMemset\", \"gbs\": %.1f}\n", bytes / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_copy<<<sms * 16, 256>>>(a, b, n4); });
        printf("{\"bench\": \"copy_rw\", \"gbs_read_plus_write\": %.1f}\n", 2.0 * bytes / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
        printf("{\"bench\": \"cudaMemcpyD2D\", \"gbs_read_plus_write\": %.1f}\n", 2.0 * bytes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
