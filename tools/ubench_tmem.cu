// How fast can 16 warps read per-lane constants from tensor memory (tcgen05.ld 32x32b), alone and next to 64-bit
// shared-memory loads?  Decides whether the Griffin-Lim kernel's constant tables can move from shared memory to TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tmem tools/ubench_tmem.cu && tools/ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD8(r, addr)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])       \
                 : "r"(addr))

// MODE 0: TMEM loads only (NT x8 loads = NT * 1 KB per warp per iteration); MODE 1: + NL LDS.64; MODE 2: LDS.64 only
template <int NT, int NL>
__global__ void __launch_bounds__(512, 1) k(float* out) {
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float2 sm[512 * 8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 512 * 8; i += 512) sm[i] = make_float2(i, -i);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // fill: every warp writes 64 columns of its lane quarter (4 warps share a quarter: same values, harmless)
    {
        uint32_t v[8];
        for (int c = 0; c < 256; c += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float)(lane * 1000 + c + j));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + c), "r"(v[0]),
                         "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float acc = 0.0f;
    float2 a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a2[i] = make_float2(0, 0);
    const float2* p = sm + tid;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            uint32_t r[8];
            LD8(r, base + ((it * 8 + i * 32) & 255));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[0]) + __uint_as_float(r[7]);
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p + 512 * (i & 7))));
            a2[i & 7].x += v.x;
            a2[i & 7].y += v.y;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a2[i].x + a2[i].y;
    out[blockIdx.x * 512 + tid] = acc;
    if (warp == 0 && blockIdx.x == 0) {  // warp-collective instruction: the whole warp executes it
        uint32_t r[8];
        LD8(r, base + 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 1) out[0] = __uint_as_float(r[3]);  // expect lane*1000 + 11 = 1011
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512));
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
    auto rep = [&](const char* n, float ms, int nt, int nl) {
        double cyc = ms * 1e-3 * khz * 1e3 / ITERS;
        printf("%-28s %7.3f ms %7.1f cyc/iter/SM: TMEM %5.1f B/clk/SM, LDS %5.1f B/clk/SM\n", n, ms, cyc, 16.0 * nt * 1024 / cyc, 16.0 * nl * 256 / cyc);
        fflush(stdout);
    };
    rep("LDTM.x8 x4", time_it([&] { k<4, 0><<<sms, 512>>>(out); }), 4, 0);
    float h; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost); printf("check value %.1f (expect 1011.0)\n", h); fflush(stdout);
    rep("LDTM.x8 x8", time_it([&] { k<8, 0><<<sms, 512>>>(out); }), 8, 0);
    rep("LDS.64 x8", time_it([&] { k<0, 8><<<sms, 512>>>(out); }), 0, 8);
    rep("LDTM.x8 x4 + LDS.64 x8", time_it([&] { k<4, 8><<<sms, 512>>>(out); }), 4, 8);
    rep("LDTM.x8 x8 + LDS.64 x8", time_it([&] { k<8, 8><<<sms, 512>>>(out); }), 8, 8);
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
