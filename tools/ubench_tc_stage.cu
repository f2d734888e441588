// Measured prototype: ONE radix-32 stage of the Griffin-Lim forward transform as a split-precision tcgen05 GEMM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/ubench_tc_stage tools/ubench_tc_stage.cu
//   /tmp/ubench_tc_stage            (numerics + cycles per frame per SM for every variant)
//
// Context: k_gl_pass runs a 1024-point complex DFT per frame as in-lane FFT-32 -> twiddle -> 32 x 32
// transpose through shared memory -> in-lane FFT-32, one warp per frame, 16 warps per SM.  The second stage (the
// contraction over the lane index n1, 22 live output rows k1 of 32) is the one a tensor core could take over without
// an extra data movement: the transpose already writes every value to shared memory once.  This benchmark builds
// that stage both ways on the same data flow and measures them on the same loop:
//
//   stage 1 (both variants): in-lane FFT-32 (packed f32x2) + twiddle table multiply          [FP32 pipe]
//   stage 2, SIMT : warp_transpose (32 STS.64 + 16 LDS.128) + in-lane FFT-32 pruned to 22 outputs  [FP32 pipe + LSU]
//   stage 2, TC   : fp16 head/tail split of the lane's 32 complex values (F2FP + FHADD), 64 STS.32 into the K-major
//                   UMMA operand layout (the transpose IS the operand write), 12 x tcgen05.mma kind::f16
//                   M = 128 (4 frames x 32 columns k2), N = 48 (22 complex outputs), K = 64 (32 complex inputs) =
//                   A_hi B_hi + A_hi B_lo + A_lo B_hi with fp32 accumulation in TMEM, issued by one elected thread per
//                   warpgroup, tcgen05.ld of the 44 live floats back into the lane's registers.
//
// Every warp of a warpgroup (4 frames) must meet at the MMA: one named barrier + one mbarrier wait per frame, which
// the SIMT kernel does not have.  While a warpgroup waits, the other three run their stage 1 on the FP32 pipe: that is
// the overlap the tensor stage can get.  fp16 x 3 and not tf32 x 3: the A operand of four frames is 2 x 16 KB in
// fp16 but 2 x 32 KB in tf32, and k_gl_pass has 8 KB of scratch per warp (230 KB of shared memory are in use).
#include <cuda_fp16.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../speech-to-speech-translation_b200/csrc/frame_fft.cuh"

using namespace s2st;

constexpr int kWarps = 16, kThreads = 32 * kWarps, kGroups = 4;
constexpr int kLive = 22;              // live output rows k1 (bins 32 k1 + k2 < 704)
constexpr int kN = 48;                 // UMMA N: 2 * 22 = 44 padded to a multiple of 16
constexpr int kK = 64;                 // 32 complex inputs as (re, im)
constexpr int kLboA = 16 * 128 + 16;   // bytes between the K chunks (8 halves) of the A tile: 16 row groups + 16 B of
                                       // padding, which spreads a warp's 32 stores (one row, 32 different n1) over 32 banks
constexpr int kLboB = (kN / 8) * 128;  // B tile: 6 row groups per K chunk
constexpr int kATileBytes = (kK / 8) * kLboA;  // 16 512 B per head / tail tile of one warpgroup
constexpr int kBTileBytes = (kK / 8) * kLboB;  //  6 144 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// head / tail split of one complex value: hi = rn_f16(v), nlo = rn_f16(hi - v)  (the NEGATED tail: FHADD computes
// f16 - f32 in one instruction; the MMA that uses the tail negates A in its instruction descriptor)
__device__ __forceinline__ void split_f16(const float2 v, uint32_t& hi, uint32_t& nlo) {
    const __half2 h = __float22half2_rn(v);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    float l0, l1;
    const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16);
    asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(l0) : "h"(h0), "f"(v.x));
    asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(l1) : "h"(h1), "f"(v.y));
    const __half2 l = __float22half2_rn(make_float2(l0, l1));
    nlo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Params {
    const float2* in;     // [n_frames][1024]: z[n1 + 32 n2] of frame f at [f][n2 * 32 + n1]
    float2* out;          // [n_frames][22 * 32]: Z[k1][k2] (second-stage output) at [f][k1 * 32 + k2]; nullptr in timing runs
    const float2* tw;     // [32 * 32] exp(-2 pi i r l / 1024)
    const uint32_t* b_hi; // B tiles in the canonical K-major layout (built on the host)
    const uint32_t* b_lo;
    int iters;
    float* sink;
    long long* cycles;    // [gridDim.x]
};

// MODE 0: stage 1 only; 1: stage 1 + SIMT stage 2; 2: stage 1 + tensor-core stage 2
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_stage(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                       // 8 KB
    unsigned char* s_b = smem_raw + 8192;                                      // B head, B tail
    unsigned char* s_work = s_b + 2 * kBTileBytes;                             // SIMT: 8 KB scratch per warp; TC: A head + tail per warpgroup
    __shared__ __align__(8) uint64_t s_bar[kGroups];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = warp >> 2, wq = warp & 3;
    for (int i = tid; i < 1024; i += kThreads) s_tw[i] = p.tw[i];
    if constexpr (MODE == 2) {
        for (int i = tid; i < kBTileBytes / 4; i += kThreads) {
            reinterpret_cast<uint32_t*>(s_b)[i] = p.b_hi[i];
            reinterpret_cast<uint32_t*>(s_b + kBTileBytes)[i] = p.b_lo[i];
        }
        if (tid < kGroups) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[tid])));
        if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if constexpr (MODE == 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = MODE == 2 ? s_tmem : 0;
    float* scratch = reinterpret_cast<float*>(s_work + warp * 8192);
    unsigned char* a_hi = s_work + grp * 2 * kATileBytes;
    unsigned char* a_lo = a_hi + kATileBytes;
    // instruction descriptors: D = F32 (bit 4), A = B = F16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29); the
    // tail term negates A (bit 13) because the split stores hi - v
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_neg = idesc | (1u << 13);
    uint32_t phase = 0;

    const long long frame = (long long)blockIdx.x * kWarps + warp;
    float2 a[32];
    {
        const float2* src = p.in + frame * 1024;
#pragma unroll
        for (int r = 0; r < 32; ++r) a[brev5(r)] = src[r * 32 + lane];  // a[brev5(n2)] = z[lane + 32 n2]
    }
    float2 acc = make_float2(0.0f, 0.0f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < p.iters; ++it) {
        // ---- stage 1: in-lane FFT over n2, twiddle exp(-2 pi i n1 k2 / 1024), scaled by 1/32 to keep the loop bounded
        fft32_inplace_br<false, 32, 32>(a);
#pragma unroll
        for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], s_tw[r * 32 + lane]);
#pragma unroll
        for (int r = 0; r < 32; ++r) a[r] = mul2(a[r], bcast2(1.0f / 32.0f));
        if constexpr (MODE == 1) {
            // ---- stage 2, SIMT: transpose (lane <-> k2), in-lane FFT over n1, 22 live outputs
            warp_transpose<true>(a, scratch, lane);
            fft32_inplace_br<false, 32, kLive>(a);
#pragma unroll
            for (int r = kLive; r < 32; ++r) a[r] = make_float2(0.0f, 0.0f);
        } else if constexpr (MODE == 2) {
            // ---- stage 2, tensor core.  A[row = 32 wq + k2][k = 2 lane + c] = Y'[n1 = lane][k2].c
            {
                // element (row, k) of the K-major tile: (k >> 3) * LBO + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2 bytes
                unsigned char* base_hi = a_hi + (lane >> 2) * kLboA + (wq * 4) * 128 + (lane & 3) * 4;
                const int d_lo = kATileBytes;
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2) {
                    uint32_t hi, nlo;
                    split_f16(a[k2], hi, nlo);
                    unsigned char* q = base_hi + (k2 >> 3) * 128 + (k2 & 7) * 16;
                    *reinterpret_cast<uint32_t*>(q) = hi;
                    *reinterpret_cast<uint32_t*>(q + d_lo) = nlo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (wq == 0 && lane == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + 64u * grp;
                const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(s_b), bl = smem_u32(s_b + kBTileBytes);
                uint32_t accf = 0;
#pragma unroll
                for (int ks = 0; ks < kK / 16; ++ks) {  // A_hi B_hi + A_hi B_lo - (hi - v) B_hi
                    mma_f16(d, make_desc(ah + ks * 2 * kLboA, kLboA, 128), make_desc(bh + ks * 2 * kLboB, kLboB, 128), idesc, accf);
                    accf = 1;
                    mma_f16(d, make_desc(ah + ks * 2 * kLboA, kLboA, 128), make_desc(bl + ks * 2 * kLboB, kLboB, 128), idesc, 1);
                    mma_f16(d, make_desc(al + ks * 2 * kLboA, kLboA, 128), make_desc(bh + ks * 2 * kLboB, kLboB, 128), idesc_neg, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[grp])) : "memory");
            }
            mbar_wait(smem_u32(&s_bar[grp]), phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + 64u * grp;
            uint32_t r[48];
#pragma unroll
            for (int c0 = 0; c0 < 48; c0 += 16)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[c0 + 0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]),
                      "=r"(r[c0 + 6]), "=r"(r[c0 + 7]), "=r"(r[c0 + 8]), "=r"(r[c0 + 9]), "=r"(r[c0 + 10]), "=r"(r[c0 + 11]),
                      "=r"(r[c0 + 12]), "=r"(r[c0 + 13]), "=r"(r[c0 + 14]), "=r"(r[c0 + 15])
                    : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k1 = 0; k1 < kLive; ++k1) a[k1] = make_float2(__uint_as_float(r[2 * k1]), __uint_as_float(r[2 * k1 + 1]));
#pragma unroll
            for (int k1 = kLive; k1 < 32; ++k1) a[k1] = make_float2(0.0f, 0.0f);
        }
        if (p.out && it == 0) {
            // lane = k2 now (MODE 1, 2): Z[k1][k2]
            float2* dst = p.out + frame * (kLive * 32);
#pragma unroll
            for (int k1 = 0; k1 < kLive; ++k1) dst[k1 * 32 + lane] = a[k1];
        }
        // the next round consumes the outputs (slot r as row brev5(r): any fixed relabelling keeps the work identical)
        acc = add2(acc, a[0]);
#pragma unroll
        for (int r = kLive; r < 32; ++r) a[r] = a[r - kLive];
    }
    const long long t1 = clock64();
    if (tid == 0) p.cycles[blockIdx.x] = t1 - t0;
    if (acc.x == 1.2345f) p.sink[0] = acc.y;
    if constexpr (MODE == 2) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
}

static uint16_t f2h(float f) {  // round to nearest even, host side
    __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
static float h2f(uint16_t u) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_frames = sms * kWarps;
    // inputs with the dynamic range of speech spectra: a few strong components + a 60 dB weaker floor
    std::vector<float2> in((size_t)n_frames * 1024);
    srand(1);
    for (size_t i = 0; i < in.size(); ++i) {
        const float u = (float)rand() / RAND_MAX - 0.5f, v = (float)rand() / RAND_MAX - 0.5f;
        const float s = (i % 37 == 0) ? 20.0f : ((i % 5 == 0) ? 1.0f : 0.02f);
        in[i] = make_float2(s * u, s * v);
    }
    std::vector<float2> tw(1024);
    for (int r = 0; r < 32; ++r)
        for (int l = 0; l < 32; ++l) {
            const double th = -2.0 * M_PI * r * l / 1024.0;
            tw[r * 32 + l] = make_float2((float)cos(th), (float)sin(th));
        }
    // B[n = (k1, c')][k = (n1, c)]: (re + i im)(cos - i sin): out_re = re cos + im sin, out_im = -re sin + im cos
    std::vector<uint32_t> bh(kBTileBytes / 4, 0), bl(kBTileBytes / 4, 0);
    auto put = [&](std::vector<uint32_t>& t, int n, int k, uint16_t v) {
        const int off = (k >> 3) * kLboB + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
        reinterpret_cast<uint16_t*>(t.data())[off / 2] = v;
    };
    for (int k1 = 0; k1 < kLive; ++k1)
        for (int n1 = 0; n1 < 32; ++n1) {
            const double th = 2.0 * M_PI * ((n1 * k1) % 32) / 32.0;
            const double w[2][2] = {{cos(th), sin(th)}, {-sin(th), cos(th)}};
            for (int cp = 0; cp < 2; ++cp)
                for (int c = 0; c < 2; ++c) {
                    const float v = (float)w[cp][c];
                    const uint16_t hi = f2h(v);
                    put(bh, 2 * k1 + cp, 2 * n1 + c, hi);
                    put(bl, 2 * k1 + cp, 2 * n1 + c, f2h((float)(w[cp][c] - (double)h2f(hi))));
                }
        }
    float2 *d_in, *d_out, *d_tw;
    uint32_t *d_bh, *d_bl;
    float* d_sink;
    long long* d_cyc;
    CK(cudaMalloc(&d_in, in.size() * sizeof(float2)));
    CK(cudaMalloc(&d_out, (size_t)n_frames * kLive * 32 * sizeof(float2)));
    CK(cudaMalloc(&d_tw, 1024 * sizeof(float2)));
    CK(cudaMalloc(&d_bh, kBTileBytes));
    CK(cudaMalloc(&d_bl, kBTileBytes));
    CK(cudaMalloc(&d_sink, 4));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    CK(cudaMemcpy(d_in, in.data(), in.size() * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tw, tw.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bh, bh.data(), kBTileBytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bl, bl.data(), kBTileBytes, cudaMemcpyHostToDevice));
    const size_t smem_simt = 8192 + 2 * kBTileBytes + kWarps * 8192;
    const size_t smem_tc = 8192 + 2 * kBTileBytes + kGroups * 2 * kATileBytes;
    CK(cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_simt));
    CK(cudaFuncSetAttribute(k_stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_simt));
    CK(cudaFuncSetAttribute(k_stage<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
    Params p{d_in, d_out, d_tw, d_bh, d_bl, 1, d_sink, d_cyc};

    // ---- numerics: one round of stage 1 + stage 2, both variants, against a float64 DFT of the float32 stage-1 output
    std::vector<float2> o1((size_t)n_frames * kLive * 32), o2(o1.size());
    k_stage<1><<<sms, kThreads, smem_simt>>>(p);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o1.data(), d_out, o1.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    CK(cudaMemset(d_out, 0, o1.size() * sizeof(float2)));
    k_stage<2><<<sms, kThreads, smem_tc>>>(p);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(o2.data(), d_out, o2.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    double num1 = 0, num2 = 0, den = 0, maxabs = 0, maxe1 = 0, maxe2 = 0;
    for (int f = 0; f < 64; ++f) {
        // exact 1024-point DFT restricted to the live rows: Z[k1][k2] = (1/32) sum_n z[n] exp(-2 pi i n (k2 + 32 k1) / 1024)
        for (int k1 = 0; k1 < kLive; k1 += 3)
            for (int k2 = 0; k2 < 32; k2 += 5) {
                double re = 0, im = 0;
                const int k = k2 + 32 * k1;
                for (int n = 0; n < 1024; ++n) {
                    const float2 z = in[(size_t)f * 1024 + (n >> 5) * 32 + (n & 31)];
                    const double th = -2.0 * M_PI * ((long long)n * k % 1024) / 1024.0;
                    re += z.x * cos(th) - z.y * sin(th);
                    im += z.x * sin(th) + z.y * cos(th);
                }
                re /= 32.0;
                im /= 32.0;
                const float2 a = o1[(size_t)f * kLive * 32 + k1 * 32 + k2], b = o2[(size_t)f * kLive * 32 + k1 * 32 + k2];
                num1 += (a.x - re) * (a.x - re) + (a.y - im) * (a.y - im);
                num2 += (b.x - re) * (b.x - re) + (b.y - im) * (b.y - im);
                den += re * re + im * im;
                maxabs = fmax(maxabs, fmax(fabs(re), fabs(im)));
                maxe1 = fmax(maxe1, fmax(fabs(a.x - re), fabs(a.y - im)));
                maxe2 = fmax(maxe2, fmax(fabs(b.x - re), fabs(b.y - im)));
            }
    }
    printf("numerics vs float64 DFT (64 frames, sampled bins): rel-L2 SIMT fp32 %.3e | tcgen05 fp16x3 %.3e ; max |err| / max |Z|: %.3e | %.3e\n",
           sqrt(num1 / den), sqrt(num2 / den), maxe1 / maxabs, maxe2 / maxabs);

    // ---- timing: cycles per frame per SM (16 frames in flight per SM, like k_gl_pass)
    p.out = nullptr;
    p.iters = iters;
    double cyc[3];
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k_stage<0><<<sms, kThreads, smem_simt>>>(p);
            if (mode == 1) k_stage<1><<<sms, kThreads, smem_simt>>>(p);
            if (mode == 2) k_stage<2><<<sms, kThreads, smem_tc>>>(p);
            CK(cudaDeviceSynchronize());
        }
        std::vector<long long> c(sms);
        CK(cudaMemcpy(c.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double s = 0;
        for (long long v : c) s += (double)v;
        cyc[mode] = s / sms / iters / kWarps;  // cycles per frame per SM
    }
    printf("cycles per frame per SM (16 warps/SM): stage 1 only %.1f | stage 1 + SIMT stage 2 %.1f | stage 1 + tcgen05 stage 2 %.1f\n",
           cyc[0], cyc[1], cyc[2]);
    printf("stage 2 alone: SIMT %.1f cycles | tcgen05 fp16x3 %.1f cycles  (k_gl_pass spends ~1210 cycles per frame-iteration per SM in total)\n",
           cyc[1] - cyc[0], cyc[2] - cyc[0]);
    return 0;
}
