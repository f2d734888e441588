// Warp-level 2048-point real FFT / inverse: one warp per frame, 32 complex values per lane.
//
// A real frame x[0..2048) is packed as z[n] = x[2n] + i x[2n+1] (1024 complex points) and
// transformed as a 32 x 32 Cooley-Tukey: an in-lane FFT-32 (fft32.cuh), one twiddle multiply,
// one 32x32 transpose through a warp-private shared-memory scratch, a second in-lane FFT-32.
// The real-signal spectrum is then recovered from Z[k] and conj(Z[1024-k]); the partner lives in
// lane (32-k2)%32, so the pairs are exchanged through the same scratch, stored linearly in k.
//
// Layouts (lane = threadIdx.x & 31):
//   time      a[n2] = z[lane + 32*n2]        (samples 2*lane + 64*n2 and +1)
//   frequency a[k1] = X[32*k1 + lane],  plus the Nyquist bin X[1024] (real) in lane 0
//
// Scratch per warp: kScratchFloats floats (8 KB) = one 32 x 32 tile of complex values.  The transposes move whole
// (re, im) pairs -- the packed-pair arithmetic wants them in adjacent registers -- with 8-byte column writes
// and 16-byte row reads; instead of padding, the 16-byte chunks of row j are XOR-swizzled with (j & 7),
// which makes both directions bank-conflict free.  The pair exchange only moves the rows that are needed:
//   PRUNED (all live bins < 704, the f_max = 8 kHz vocoder case): one round, rows 10..31 forward /
//          rows 0..21 inverse, 704 complex values + the bin-1024 alias;
//   generic: two rounds of 16 rows with the first round's partners parked in registers.
#pragma once
#include "fft32.cuh"

namespace s2st {

constexpr int kPrunedRows = 22;         // rows (of 32 bins) that can be live in PRUNED mode
constexpr int kScratchFloats = 2048;    // 32*32 complex (transpose) >= 2*705 (pruned exchange), 2*513 (generic)

// Complex products as two packed-pair instructions (FMUL2 + FFMA2).  "Swap the halves / negate one half" is a
// free modifier of the FIRST source operand only (SASS .LO_HI.NP); ptxas does not commute it past a broadcast
// scalar, so the modified operand is always written first.
__device__ __forceinline__ float2 cmul(const float2 a, const float2 b) {
    return fma2(mul_pi(b), bcast2(a.y), mul2(b, bcast2(a.x)));
}
__device__ __forceinline__ float2 cmul_conj(const float2 a, const float2 b) {  // a * conj(b)
    return fma2(swap2(b), bcast2(a.y), mul2(conj2(b), bcast2(a.x)));
}

// 32x32 transpose of the per-lane register arrays: out a[r] = (lane r's) a[lane].
// Element (row j, column l) lives in 16-byte chunk ((l >> 1) ^ (j & 7)) of row j (256 B per row), half (l & 1).
template <bool BR = false>  // BR: deliver element r into slot brev5(r) (input order of fft32_inplace_br)
__device__ __forceinline__ void warp_transpose(float2 (&a)[32], float* scratch, int lane) {
    char* base = reinterpret_cast<char*>(scratch);
    const int wofs = lane * 8;  // ((lane >> 1) << 4) | ((lane & 1) << 3)
#pragma unroll
    for (int j = 0; j < 32; ++j)
        *reinterpret_cast<float2*>(base + j * 256 + (wofs ^ ((j & 7) << 4))) = a[j];
    __syncwarp();
    const char* row = base + lane * 256;
    const int sw = (lane & 7) << 4;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(row + ((((c & 7) << 4) ^ sw) | ((c & 8) << 4)));
        a[BR ? brev5(2 * c) : 2 * c] = make_float2(v.x, v.y);
        a[BR ? brev5(2 * c + 1) : 2 * c + 1] = make_float2(v.z, v.w);
    }
    __syncwarp();
}

// X2 = (Z + conj P) + V (Z - conj P)     (V = -i exp(-2 pi i k / 2048): forward split)
__device__ __forceinline__ float2 split_fwd(const float2 z, const float2 p, const float2 v) {
    const float2 s = add2(z, conj2(p)), d = add2(z, neg2(conj2(p)));
    return fma2(d, bcast2(v.x), fma2(mul_pi(d), bcast2(v.y), s));
}
// Z' = (Y + conj P) + conj(V) (Y - conj P)   (inverse merge).  SW: return Z' with its halves exchanged; the
// exchange is pushed down to the operands (where it is a free operand modifier) instead of moving registers.
template <bool SW>
__device__ __forceinline__ float2 sw2(const float2 o) { return SW ? swap2(o) : o; }
template <bool SW>
__device__ __forceinline__ float2 merge_inv(const float2 y, const float2 p, const float2 v) {
    const float2 s = add2(sw2<SW>(y), sw2<SW>(conj2(p))), d = add2(sw2<SW>(y), sw2<SW>(neg2(conj2(p))));
    return fma2(d, bcast2(v.x), fma2(SW ? mul_pi(d) : mul_mi(d), bcast2(v.y), s));
}
// the same with P = 0  (Z' = Y + conj(V) Y)  and with Y = 0  (Z' = conj(P) - conj(V) conj(P))
template <bool SW>
__device__ __forceinline__ float2 merge_inv_y(const float2 y, const float2 v) {
    const float2 d = sw2<SW>(y);
    return fma2(d, bcast2(v.x), fma2(SW ? mul_pi(d) : mul_mi(d), bcast2(v.y), d));
}
template <bool SW>
__device__ __forceinline__ float2 merge_inv_p(const float2 p, const float2 v) {
    const float2 s = sw2<SW>(conj2(p)), d = sw2<SW>(neg2(conj2(p)));
    return fma2(d, bcast2(v.x), fma2(SW ? mul_pi(d) : mul_mi(d), bcast2(v.y), s));
}

// Forward, first half: in-lane FFT over n2, twiddle, transpose.  In: a[n2] for n2 < NZ (others ignored).
// tw[r*32 + lane] = exp(-2*pi*i*r*lane/1024).
template <int NZ>
__device__ __forceinline__ void frame_fwd_a(float2 (&a)[32], float* scratch, const float2* __restrict__ tw, int lane) {
    fft32<NZ, false>(a);
#pragma unroll
    for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], tw[r * 32 + lane]);
    warp_transpose(a, scratch, lane);
}

// Pair exchange + real-signal split after the second in-lane FFT.  In: a[k1] = Z[32*k1+lane].
// Out: a[k1] = 2*X[32*k1+lane] (PRUNED: only k1 < 22, the rest undefined), nyq = 2*X[1024] (valid in
// lane 0).  vtab[k] = -i*exp(-2*pi*i*k/2048).
template <bool PRUNED>
__device__ __forceinline__ void fwd_split(float2 (&a)[32], float& nyq, float* scratch,
                                          const float2* __restrict__ vtab, int lane) {
    nyq = 2.0f * (a[0].x - a[0].y);
    float2* sc = reinterpret_cast<float2*>(scratch);
    if constexpr (PRUNED) {
        // bins k < 704 need partners 1024-k in [321, 1024]: rows 10..31 (stored from bin 320) + alias of bin 0
#pragma unroll
        for (int r = 32 - kPrunedRows; r < 32; ++r) sc[32 * r + lane - 320] = a[r];
        if (lane == 0) sc[704] = a[0];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kPrunedRows; ++r) {
            const int k = 32 * r + lane;
            a[r] = split_fwd(a[r], sc[704 - k], vtab[k]);
        }
        __syncwarp();
    } else {
        float2 t[16];
        // round 1: upper rows -> scratch (from bin 512), bin 1024 aliases bin 0; partners of the lower rows
#pragma unroll
        for (int r = 16; r < 32; ++r) sc[32 * (r - 16) + lane] = a[r];
        if (lane == 0) sc[512] = a[0];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) t[r] = sc[512 - 32 * r - lane];
        __syncwarp();
        // round 2: lower rows -> scratch; upper rows are finished in place (bin 512 pairs with itself)
#pragma unroll
        for (int r = 0; r < 16; ++r) sc[32 * r + lane] = a[r];
        __syncwarp();
#pragma unroll
        for (int r = 16; r < 32; ++r) {
            const int k = 32 * r + lane;
            const float2 p = (k == 512) ? a[r] : sc[1024 - k];
            a[r] = split_fwd(a[r], p, vtab[k]);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r] = split_fwd(a[r], t[r], vtab[32 * r + lane]);
        __syncwarp();
    }
}

// Forward, second half: in-lane FFT over n1, then the split.
template <bool PRUNED>
__device__ __forceinline__ void frame_fwd_b(float2 (&a)[32], float& nyq, float* scratch,
                                            const float2* __restrict__ vtab, int lane) {
    fft32<32, false>(a);
    fwd_split<PRUNED>(a, nyq, scratch, vtab, lane);
}

// Forward = both halves, all bins.
template <int NZ>
__device__ __forceinline__ void frame_fwd(float2 (&a)[32], float& nyq, float* scratch,
                                          const float2* __restrict__ tw,
                                          const float2* __restrict__ vtab, int lane) {
    frame_fwd_a<NZ>(a, scratch, tw, lane);
    frame_fwd_b<false>(a, nyq, scratch, vtab, lane);
}

// Hermitian merge before the inverse transform.  In: a[k1] = Y[32*k1+lane] (half-spectrum, imag of DC
// ignored; PRUNED: rows >= 22 are taken as zero whatever they hold), ynyq = Y[1024] (real, from lane 0).
// Out: a[k1] = Z'[32*k1+lane], the packed 1024-point spectrum; SWAP: real and imaginary parts exchanged,
// which turns the inverse transform into a forward one (IDFT(Z) = swap(DFT(swap(Z)))).
template <bool PRUNED, bool SWAP, bool BR = false>  // BR: Z' row r goes to slot brev5(r)
__device__ __forceinline__ void inv_merge(float2 (&a)[32], float ynyq, float* scratch,
                                          const float2* __restrict__ vtab, int lane) {
    if (lane == 0) a[0].y = 0.0f;
    float2 z[32];
    float2* sc = reinterpret_cast<float2*>(scratch);
    if constexpr (PRUNED) {
#pragma unroll
        for (int r = 0; r < kPrunedRows; ++r) sc[32 * r + lane] = a[r];
        if (lane == 0) sc[704] = make_float2(ynyq, 0.0f);
        __syncwarp();
        // partner bin 1024-k is live (< 704) iff k > 320: never for rows < 10, always for rows > 10, row 10
        // except lane 0; bin 0 pairs with the Nyquist alias.  Rows >= 22 hold no Y of their own.
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int k = 32 * r + lane;
            const float2 v = vtab[k];
            if (r < 10) {
                float2 p = make_float2(0.0f, 0.0f);
                if (r == 0 && lane == 0) p = sc[704];
                z[r] = (r == 0) ? merge_inv<SWAP>(a[r], p, v) : merge_inv_y<SWAP>(a[r], v);
            } else if (r == 10) {
                float2 p = make_float2(0.0f, 0.0f);
                if (lane > 0) p = sc[1024 - k];
                z[r] = merge_inv<SWAP>(a[r], p, v);
            } else if (r < kPrunedRows) {
                z[r] = merge_inv<SWAP>(a[r], sc[1024 - k], v);
            } else {
                z[r] = merge_inv_p<SWAP>(sc[1024 - k], v);
            }
        }
        __syncwarp();
    } else {
        float2 t[16];
#pragma unroll
        for (int r = 16; r < 32; ++r) sc[32 * (r - 16) + lane] = a[r];
        if (lane == 0) sc[512] = make_float2(ynyq, 0.0f);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) t[r] = sc[512 - 32 * r - lane];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 16; ++r) sc[32 * r + lane] = a[r];
        __syncwarp();
#pragma unroll
        for (int r = 16; r < 32; ++r) {
            const int k = 32 * r + lane;
            const float2 p = (k == 512) ? a[r] : sc[1024 - k];
            z[r] = merge_inv<SWAP>(a[r], p, vtab[k]);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            z[r] = merge_inv<SWAP>(a[r], t[r], vtab[32 * r + lane]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) a[BR ? brev5(r) : r] = z[r];
}

// The 1024-point complex forward DFT of the per-lane arrays, in place: in a[brev5(j)] = z[lane + 32*j] (the
// bit-reversed slots fft32_inplace_br wants), out a[j] = Z[lane + 32*j] -- input and output use the same
// (index mod 32, index div 32) layout, which is what lets one routine serve both directions.  The in-lane FFT
// is emitted once and run twice; it works in place, so the loop carries no register shuffling.
// NZ_IN: input rows >= NZ_IN are zero (their slots need not be initialised); NZ_OUT: only output rows
// < NZ_OUT are needed (the others are left undefined).  SHARED_FFT: emit the in-lane FFT once and run it twice
// (smaller code, but the array becomes loop-carried and no pruning is possible).
template <int NZ_IN = 32, int NZ_OUT = 32, bool SHARED_FFT = false>
__device__ __forceinline__ void fwd1024(float2 (&a)[32], float* scratch, const float2* __restrict__ tw, int lane) {
    if constexpr (SHARED_FFT) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            fft32_inplace_br<false, 32, 32, true>(a);
            if (h == 0) {
#pragma unroll
                for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], tw[r * 32 + lane]);
                warp_transpose<true>(a, scratch, lane);
            }
        }
    } else {
        fft32_inplace_br<false, NZ_IN, 32>(a);
#pragma unroll
        for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], tw[r * 32 + lane]);
        warp_transpose<true>(a, scratch, lane);
        fft32_inplace_br<false, 32, NZ_OUT>(a);
    }
}

// Inverse.  In: as inv_merge.  Out: a[n2] = 2048 * y[2*(lane+32*n2)] + i * 2048 * y[..+1].
template <bool PRUNED>
__device__ __forceinline__ void frame_inv(float2 (&a)[32], float ynyq, float* scratch,
                                          const float2* __restrict__ tw,
                                          const float2* __restrict__ vtab, int lane) {
    inv_merge<PRUNED, false>(a, ynyq, scratch, vtab, lane);
    fft32<32, true>(a);
#pragma unroll
    for (int r = 1; r < 32; ++r) a[r] = cmul_conj(a[r], tw[r * 32 + lane]);
    warp_transpose(a, scratch, lane);
    fft32<32, true>(a);
}

}  // namespace s2st
