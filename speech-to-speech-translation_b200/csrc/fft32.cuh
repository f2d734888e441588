// In-lane (per-thread, all-register) 32-point complex FFT with compile-time twiddles.
//
// Radix-2 decimation-in-time, written as a template recursion so that every index is a
// compile-time constant: the arrays live entirely in registers and the twiddle constants
// become instruction immediates.  Twiddled butterflies use the 6-FMA "factor out the cosine" form
//     e +- w*o = e +- c * (o + tau * (-i o)),  tau = s/c  (or the sine form when |s| > |c|).
// All complex arithmetic is issued as Blackwell packed-pair instructions (FFMA2 / FADD2 / FMUL2 on a
// (re, im) register pair): one issue slot per complex operation instead of two.  The "-i o" operand
// (swap the halves, negate one) is an operand modifier of those instructions (.LO_HI.NP in SASS), so
// it costs nothing; ptxas folds it from the make_float2(o.y, -o.x) expression below.
// NZ = number of leading non-zero inputs: Griffin-Lim frames are a 1200-sample window inside a
// 2048-point transform, so 13 of the 32 strided inputs each lane sees are structurally zero and
// the leaf butterflies that only copy are never emitted.
#pragma once
#include <cuda_runtime.h>

namespace s2st {

// Packed-pair helpers (sm_100a: fma/add/mul.rn.f32x2).  Results are bit-identical to the scalar fmaf forms.
__device__ __forceinline__ float2 bcast2(const float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 neg2(const float2 o) { return make_float2(-o.x, -o.y); }
__device__ __forceinline__ float2 conj2(const float2 o) { return make_float2(o.x, -o.y); }
__device__ __forceinline__ float2 swap2(const float2 o) { return make_float2(o.y, o.x); }
__device__ __forceinline__ float2 mul_mi(const float2 o) { return make_float2(o.y, -o.x); }  // -i * o
__device__ __forceinline__ float2 mul_pi(const float2 o) { return make_float2(-o.y, o.x); }  // +i * o
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) { return __fmul2_rn(a, b); }

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
        lo = add2(e, o);
        hi = add2(e, neg2(o));
    } else if constexpr (J == 8) {
        // w = -+ i  ->  w*o = (+-o.y, -+o.x)
        if constexpr (!INV) {
            lo = add2(e, mul_mi(o));
            hi = add2(e, mul_pi(o));
        } else {
            lo = add2(e, mul_pi(o));
            hi = add2(e, mul_mi(o));
        }
    } else {
        constexpr float c = cos32(J);
        constexpr float s = INV ? -sin32(J) : sin32(J);  // w*o = (o.x c + o.y s, o.y c - o.x s)
        if constexpr ((c < 0 ? -c : c) >= (s < 0 ? -s : s)) {
            constexpr float tau = s / c;
            const float2 t = fma2(mul_mi(o), bcast2(tau), o);  // (o.x + tau o.y, o.y - tau o.x)
            lo = fma2(t, bcast2(c), e);
            hi = fma2(t, bcast2(-c), e);
        } else {
            constexpr float kap = c / s;
            const float2 t = fma2(o, bcast2(kap), mul_mi(o));  // (kap o.x + o.y, kap o.y - o.x)
            lo = fma2(t, bcast2(s), e);
            hi = fma2(t, bcast2(-s), e);
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

// ---- in-place variants on bit-reversed input ------------------------------------------------------------
// brev5(p) / brev4(p): 5- and 4-bit bit reversal.
__host__ __device__ constexpr int brev5(int p) {
    return ((p & 1) << 4) | ((p & 2) << 2) | (p & 4) | ((p & 8) >> 2) | ((p & 16) >> 4);
}
__host__ __device__ constexpr int brev4(int p) { return ((p & 1) << 3) | ((p & 2) << 1) | ((p & 4) >> 1) | ((p & 8) >> 3); }
template <int N>
__host__ __device__ constexpr int brev_n(int p) {
    static_assert(N == 16 || N == 32, "sizes in use");
    return N == 32 ? brev5(p) : brev4(p);
}

// Butterfly I of the (N/2) log2(N) of an N-point radix-2 DIT run in place: stage s = I / (N/2) has half-span
// m = 2^s and twiddle exp(-+2 pi i k / (2m)) = W_32^(k * 16 / m) for every N.
// NZ_IN: only input rows < NZ_IN are non-zero (slot p holds row brev(p)): first-stage butterflies whose odd
// operand is structurally zero are copies.  NZ_OUT: only outputs k < NZ_OUT are used: last-stage butterflies
// skip the unused half.  LOOPED: the array is carried around a loop, so the last stage must leave its results
// in the operand slots without a spare register (hi overwrites o, then lo = 2e - hi overwrites e).
template <int N, int I, bool INV, int NZ_IN, int NZ_OUT, bool LOOPED>
struct InplaceStep {
    static __device__ __forceinline__ void run(float2 (&b)[N]) {
        constexpr int H = N / 2, LAST = (N == 32 ? 4 : 3), TOTAL = H * (LAST + 1);
        constexpr int s = I / H, idx = I % H, m = 1 << s;
        constexpr int k = idx % m, blk = idx / m, i = blk * 2 * m + k;
        if constexpr (s == 0 && brev_n<N>(i + m) >= NZ_IN) {
            b[i + m] = b[i];
        } else if constexpr (s == LAST && k + H >= NZ_OUT) {
            float2 hi;
            bfly<k*(16 / m), INV>(b[i], b[i + m], b[i], hi);
        } else if constexpr (s == LAST && LOOPED) {
            float2 lo, hi;
            bfly<k*(16 / m), INV>(b[i], b[i + m], lo, hi);
            b[i + m] = hi;
            b[i] = fma2(b[i], bcast2(2.0f), neg2(hi));
        } else {
            bfly<k*(16 / m), INV>(b[i], b[i + m], b[i], b[i + m]);
        }
        if constexpr (I + 1 < TOTAL) InplaceStep<N, I + 1, INV, NZ_IN, NZ_OUT, LOOPED>::run(b);
    }
};

// N-point FFT in place (N = 32 or 16): in b[p] = x[brev(p)], out b[k] = X[k].  Every value stays in the array
// slot (register) it was computed into; producers write their values to the bit-reversed slot (a compile-time
// renaming), consumers read natural order.
template <bool INV, int NZ_IN = 32, int NZ_OUT = 32, bool LOOPED = false>
__device__ __forceinline__ void fft32_inplace_br(float2 (&b)[32]) {
    static_assert(NZ_IN > 16 && NZ_OUT > 16, "pruning only covers the upper half");
    InplaceStep<32, 0, INV, NZ_IN, NZ_OUT, LOOPED>::run(b);
}
template <int NZ_IN = 16, int NZ_OUT = 16>
__device__ __forceinline__ void fft16_inplace_br(float2 (&b)[16]) {
    static_assert(NZ_IN > 8 && NZ_OUT > 8, "pruning only covers the upper half");
    InplaceStep<16, 0, false, NZ_IN, NZ_OUT, false>::run(b);
}

// In-place (from the caller's point of view) 32-point FFT, natural order in and out.
template <int NZ, bool INV>
__device__ __forceinline__ void fft32(float2 (&a)[32]) {
    float2 out[32];
    FftDit<32, NZ, INV, 1, 0>::run(a, out);
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = out[k];
}

}  // namespace s2st
