// Can packed (FFMA2) and scalar (FFMA) FP32 instructions run concurrently on B200's two FMA pipes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_fmamix tools/ubench_fmamix.cu && /tmp/ubench_fmamix
// P packed + S scalar independent chains per thread per iteration; reports flop-lanes / clk / SM.  If packed work only
// occupied the "heavy" pipe, a mix would exceed the 128 lanes/clk/SM that either kind reaches alone.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int P, int S>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float2 a[P > 0 ? P : 1];
    float s[S > 0 ? S : 1];
    const float b = seed, c = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < P; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = seed + 3 * i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < (P > S ? P : S); ++i) {
            if (i < P) a[i] = __ffma2_rn(a[i], make_float2(b, b), make_float2(c, c));
            if (i < S) s[i] = fmaf(s[i], b, c);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) r += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < S; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int P, int S>
void run(float* out, int sms, int clk_khz, int warps_per_sm) {
    const int grid = sms * warps_per_sm / 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<P, S><<<grid, 256>>>(out, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<P, S><<<grid, 256>>>(out, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double lanes = (double)grid * 256 * ITERS * (2 * P + S);
    const double instr = (double)grid * 8 * ITERS * (P + S);
    printf("%2d warps/SM  %d FFMA2 + %d FFMA per iteration: %7.3f ms  %6.1f flop-lanes/clk/SM  %5.2f warp-instr/clk/SM\n", warps_per_sm, P, S, ms,
           lanes / (ms * 1e-3) / sms / (clk_khz * 1e3), instr / (ms * 1e-3) / sms / (clk_khz * 1e3));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    for (int w : {64, 16}) {
        run<8, 0>(out, sms, clk_khz, w);
        run<0, 8>(out, sms, clk_khz, w);
        run<8, 8>(out, sms, clk_khz, w);
        run<8, 4>(out, sms, clk_khz, w);
        run<6, 6>(out, sms, clk_khz, w);
        run<4, 8>(out, sms, clk_khz, w);
    }
    return 0;
}
