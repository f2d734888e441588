// Micro-benchmarks that decide the FFT kernel design on B200 (run under gpurun):
// FP32 pipe issue rates for scalar vs packed (f32x2) instructions, shared-memory and shuffle rates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float a[8], b = seed, c = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // scalar FFMA, 8 independent chains
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {  // FFMA2: 4 packed chains (same flop count as MODE 0)
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2}; mov.b64 z, {%3,%3};"
                             " fma.rn.f32x2 x, x, y, z; mov.b64 {%0,%1}, x; }"
                             : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b), "f"(c));
        } else if (MODE == 2) {  // scalar FADD
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a[i] + b;
        } else if (MODE == 3) {  // FADD2
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2};"
                             " add.rn.f32x2 x, x, y; mov.b64 {%0,%1}, x; }"
                             : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b));
        } else if (MODE == 4) {  // mixed FFMA + FADD (4 + 4)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
            for (int i = 4; i < 8; ++i) a[i] = a[i] + b;
        } else if (MODE == 5) {  // FFMA with immediate operand
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, 0.001f);
        } else if (MODE == 6) {  // FMUL scalar
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a[i] * b;
        } else if (MODE == 7) {  // 16 FFMA2 worth = test 2x more packed work per iteration
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; i += 2)
                    asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%2}; mov.b64 z, {%3,%3};"
                                 " fma.rn.f32x2 x, x, y, z; mov.b64 {%0,%1}, x; }"
                                 : "+f"(a[i]), "+f"(a[i + 1]) : "f"(b), "f"(c));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) kmem(float* out, float seed) {
    __shared__ __align__(16) float sm[256 * 4 + 64];
    for (int i = threadIdx.x; i < 256 * 4 + 64; i += 256) sm[i] = seed + i;
    __syncthreads();
    float acc = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // LDS.32 conflict free
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += sm[((it + i) & 3) * 256 + threadIdx.x];
        } else if (MODE == 1) {  // LDS.64
#pragma unroll
            for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&sm[((it + i) & 1) * 512 + 2 * threadIdx.x]); acc += v.x + v.y; }
        } else if (MODE == 2) {  // LDS.128
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 v = *reinterpret_cast<float4*>(&sm[4 * ((threadIdx.x + it + i) & 255)]); acc += v.x + v.y + v.z + v.w; }
        } else if (MODE == 3) {  // SHFL
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += __shfl_xor_sync(0xffffffffu, acc, 1 + (i & 15));
        } else if (MODE == 4) {  // STS.64
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float2*>(&sm[((it + i) & 1) * 512 + 2 * threadIdx.x]) = make_float2(acc, acc + i);
            acc += 1.0f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[lane + w];
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, sms, clk_khz / 1000);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int grid = sms * 8;  // 8 CTAs x 256 threads = 64 warps / SM
    const char* names[] = {"FFMA scalar (8/thr/iter)", "FFMA2 (4 packed = 8 flop-lanes)", "FADD scalar", "FADD2", "FFMA+FADD mixed", "FFMA imm", "FMUL scalar", "FFMA2 x8 (16 flop-lanes)"};
    float lanes_per_iter[] = {8, 8, 8, 8, 8, 8, 8, 16};
    float ms[8];
    ms[0] = time_it([&] { k<0><<<grid, 256>>>(out, 1.0001f); });
    ms[1] = time_it([&] { k<1><<<grid, 256>>>(out, 1.0001f); });
    ms[2] = time_it([&] { k<2><<<grid, 256>>>(out, 1.0001f); });
    ms[3] = time_it([&] { k<3><<<grid, 256>>>(out, 1.0001f); });
    ms[4] = time_it([&] { k<4><<<grid, 256>>>(out, 1.0001f); });
    ms[5] = time_it([&] { k<5><<<grid, 256>>>(out, 1.0001f); });
    ms[6] = time_it([&] { k<6><<<grid, 256>>>(out, 1.0001f); });
    ms[7] = time_it([&] { k<7><<<grid, 256>>>(out, 1.0001f); });
    for (int i = 0; i < 8; ++i) {
        double ops = (double)grid * 256 * ITERS * lanes_per_iter[i];
        printf("%-36s %8.3f ms  %8.2f Gop-lanes/s  = %6.1f lanes/clk/SM @%d MHz nominal (TFLOP-equiv x2 for FMA: %.1f)\n", names[i], ms[i],
               ops / ms[i] / 1e6, ops / (ms[i] * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, 2 * ops / ms[i] / 1e9);
    }
    const char* mn[] = {"LDS.32", "LDS.64", "LDS.128", "SHFL", "STS.64"};
    float bytes[] = {4, 8, 16, 4, 8};
    float mm[5];
    mm[0] = time_it([&] { kmem<0><<<grid, 256>>>(out, 1.f); });
    mm[1] = time_it([&] { kmem<1><<<grid, 256>>>(out, 1.f); });
    mm[2] = time_it([&] { kmem<2><<<grid, 256>>>(out, 1.f); });
    mm[3] = time_it([&] { kmem<3><<<grid, 256>>>(out, 1.f); });
    mm[4] = time_it([&] { kmem<4><<<grid, 256>>>(out, 1.f); });
    for (int i = 0; i < 5; ++i) {
        double n = (double)grid * 256 * ITERS * 8;
        printf("%-8s %8.3f ms  %7.1f lane-ops/clk/SM  %7.1f B/clk/SM (nominal clock)\n", mn[i], mm[i],
               n / (mm[i] * 1e-3) / sms / (clk_khz * 1e3), n * bytes[i] / (mm[i] * 1e-3) / sms / (clk_khz * 1e3));
    }
    return 0;
}
