// In-lane (per-thread, all-register) 32-point complex FFT with compile-time twiddles.
//
// Radix-2 decimation-in-time, written as a template recursion so that every index is a
// compile-time constant: the arrays live entirely in registers and the twiddle constants
// become FFMA immediates.  Twiddled butterflies use the 6-FMA "factor out the cosine" form
//     e +- w*o = e +- c * (o + i*tau*o'),  tau = s/c  (or the sine form when |s| > |c|).
// NZ = number of leading non-zero inputs: Griffin-Lim frames are a 1200-sample window inside a
// 2048-point transform, so 13 of the 32 strided inputs each lane sees are structurally zero and
// the leaf butterflies that only copy are never emitted.
#pragma once
#include <cuda_runtime.h>

namespace s2st {

// cos(2*pi*j/32), j = 0..8
__host__ __device__ constexpr float quarter_cos32(int j) {
    constexpr float t[9] = {1.0f,
                            0.98078528040323044913f,
                            0.92387953251128675613f,
                            0.83146961230254523708f,
                            0.70710678118654752440f,
                            0.55557023301960222474f,
                            0.38268343236508977173f,
                            0.19509032201612826785f,
                            0.0f};
    return t[j];
}
__host__ __device__ constexpr float cos32(int j) {
    j &= 31;
    return j <= 8 ? quarter_cos32(j) : j <= 16 ? -quarter_cos32(16 - j) : j <= 24 ? -quarter_cos32(j - 16)
                                                                                  : quarter_cos32(32 - j);
}
__host__ __device__ constexpr float sin32(int j) { return cos32(j + 24); }

// lo = e + w*o, hi = e - w*o with w = exp(-+ 2*pi*i*J/32) (minus for the forward transform).
template <int J, bool INV>
__device__ __forceinline__ void bfly(const float2 e, const float2 o, float2& lo, float2& hi) {
    static_assert(J >= 0 && J < 16, "twiddle index");
    if constexpr (J == 0) {
        lo = make_float2(e.x + o.x, e.y + o.y);
        hi = make_float2(e.x - o.x, e.y - o.y);
    } else if constexpr (J == 8) {
        // w = -+ i  ->  w*o = (+-o.y, -+o.x)
        if constexpr (!INV) {
            lo = make_float2(e.x + o.y, e.y - o.x);
            hi = make_float2(e.x - o.y, e.y + o.x);
        } else {
            lo = make_float2(e.x - o.y, e.y + o.x);
            hi = make_float2(e.x + o.y, e.y - o.x);
        }
    } else {
        constexpr float c = cos32(J);
        constexpr float s = INV ? -sin32(J) : sin32(J);  // w*o = (o.x c + o.y s, o.y c - o.x s)
        if constexpr ((c < 0 ? -c : c) >= (s < 0 ? -s : s)) {
            constexpr float tau = s / c;
            const float tx = fmaf(o.y, tau, o.x);
            const float ty = fmaf(-o.x, tau, o.y);
            lo = make_float2(fmaf(c, tx, e.x), fmaf(c, ty, e.y));
            hi = make_float2(fmaf(-c, tx, e.x), fmaf(-c, ty, e.y));
        } else {
            constexpr float kap = c / s;
            const float tx = fmaf(o.x, kap, o.y);
            const float ty = fmaf(o.y, kap, -o.x);
            lo = make_float2(fmaf(s, tx, e.x), fmaf(s, ty, e.y));
            hi = make_float2(fmaf(-s, tx, e.x), fmaf(-s, ty, e.y));
        }
    }
}

template <int K, int N, bool INV>
struct BflyLoop {
    static __device__ __forceinline__ void run(const float2 (&e)[N / 2], const float2 (&o)[N / 2],
                                               float2 (&out)[N]) {
        bfly<K*(32 / N), INV>(e[K], o[K], out[K], out[K + N / 2]);
        if constexpr (K + 1 < N / 2) BflyLoop<K + 1, N, INV>::run(e, o, out);
    }
};

// out[k] = sum_j in[OFF + j*STRIDE] * exp(-+2*pi*i*j*k/N), only j < NZ non-zero.
template <int N, int NZ, bool INV, int STRIDE, int OFF>
struct FftDit {
    static __device__ __forceinline__ void run(const float2 (&in)[32], float2 (&out)[N]) {
        if constexpr (N == 1) {
            out[0] = in[OFF];
        } else if constexpr (NZ == 1) {
#pragma unroll
            for (int k = 0; k < N; ++k) out[k] = in[OFF];
        } else {
            constexpr int NZE = (NZ + 1) / 2, NZO = NZ / 2;
            float2 e[N / 2];
            FftDit<N / 2, NZE, INV, 2 * STRIDE, OFF>::run(in, e);
            if constexpr (NZO > 0) {
                float2 o[N / 2];
                FftDit<N / 2, NZO, INV, 2 * STRIDE, OFF + STRIDE>::run(in, o);
                BflyLoop<0, N, INV>::run(e, o, out);
            } else {
#pragma unroll
                for (int k = 0; k < N / 2; ++k) out[k] = out[k + N / 2] = e[k];
            }
        }
    }
};

// In-place (from the caller's point of view) 32-point FFT, natural order in and out.
template <int NZ, bool INV>
__device__ __forceinline__ void fft32(float2 (&a)[32]) {
    float2 out[32];
    FftDit<32, NZ, INV, 1, 0>::run(a, out);
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = out[k];
}

}  // namespace s2st
