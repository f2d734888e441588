// Does SHFL share the shared-memory crossbar with LDS/STS on B200?  Times loops of 64-bit LDS, SHFL and both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_shfl tools/ubench_shfl.cu && tools/ubench_shfl
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int NLDS, int NSHFL>
__global__ void __launch_bounds__(512, 1) k(float* out, int lane_xor) {
    __shared__ __align__(16) float2 sm[512 * 9];
    for (int i = threadIdx.x; i < 512 * 9; i += 512) sm[i] = make_float2(i, -i);
    __syncthreads();
    float2 acc[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = make_float2(0, 0); s[i] = threadIdx.x + i; }
    const float2* p = sm + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NLDS; ++i) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p + 512 * i)));
            acc[i].x += v.x; acc[i].y += v.y;
        }
#pragma unroll
        for (int i = 0; i < NSHFL; ++i) s[i] = __shfl_xor_sync(0xffffffffu, s[i], lane_xor);
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y + s[i];
    out[blockIdx.x * 512 + threadIdx.x] = r;
}
template <typename F> float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 512);
    auto rep = [&](const char* n, float ms, int nl, int ns) {
        double cyc = ms * 1e-3 * khz * 1e3 / ITERS;  // cycles per iteration per SM (16 warps)
        printf("%-22s %7.3f ms  %7.1f cyc/iter/SM  (16 warps x %d LDS.64 = %d wavefronts, x %d SHFL = %d warp-shfl)\n", n, ms, cyc, nl, 16 * nl * 2, ns, 16 * ns);
    };
    rep("LDS.64 x8", time_it([&] { k<8, 0><<<sms, 512>>>(out, 5); }), 8, 0);
    rep("SHFL x8", time_it([&] { k<0, 8><<<sms, 512>>>(out, 5); }), 0, 8);
    rep("LDS.64 x8 + SHFL x8", time_it([&] { k<8, 8><<<sms, 512>>>(out, 5); }), 8, 8);
    rep("LDS.64 x4 + SHFL x8", time_it([&] { k<4, 8><<<sms, 512>>>(out, 5); }), 4, 8);
    rep("LDS.64 x8 + SHFL x4", time_it([&] { k<8, 4><<<sms, 512>>>(out, 5); }), 8, 4);
    return 0;
}
