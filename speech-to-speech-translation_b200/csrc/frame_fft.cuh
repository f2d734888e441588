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
// Scratch per warp: kScratchFloats floats (5648 B).  To stay that small the transposes move the real
// and imaginary parts one after the other (rows of 36 floats: 4-byte column writes and 16-byte row
// reads are both bank-conflict free), and the pair exchange only moves the rows that are needed:
//   PRUNED (all live bins < 704, the f_max = 8 kHz vocoder case): one round, rows 10..31 forward /
//          rows 0..21 inverse, 704 complex values + the bin-1024 alias;
//   generic: two rounds of 16 rows with the first round's partners parked in registers.
#pragma once
#include "fft32.cuh"

namespace s2st {

constexpr int kScratchPitch = 36;       // floats per transpose row (144 B)
constexpr int kPrunedRows = 22;         // rows (of 32 bins) that can be live in PRUNED mode
constexpr int kScratchFloats = 1412;    // >= 32*36 (transpose), 2*705 (pruned exchange), 2*513 (generic)

__device__ __forceinline__ float2 cmul(const float2 a, const float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(const float2 a, const float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// 32x32 transpose of the per-lane register arrays: out a[r] = (lane r's) a[lane].
__device__ __forceinline__ void warp_transpose(float2 (&a)[32], float* scratch, int lane) {
    const float4* row = reinterpret_cast<const float4*>(scratch + lane * kScratchPitch);
#pragma unroll
    for (int r = 0; r < 32; ++r) scratch[r * kScratchPitch + lane] = a[r].x;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = row[c];
        a[4 * c].x = v.x;
        a[4 * c + 1].x = v.y;
        a[4 * c + 2].x = v.z;
        a[4 * c + 3].x = v.w;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) scratch[r * kScratchPitch + lane] = a[r].y;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = row[c];
        a[4 * c].y = v.x;
        a[4 * c + 1].y = v.y;
        a[4 * c + 2].y = v.z;
        a[4 * c + 3].y = v.w;
    }
    __syncwarp();
}

// X2 = (Z + conj P) + V (Z - conj P)     (V = -i exp(-2 pi i k / 2048): forward split)
__device__ __forceinline__ float2 split_fwd(const float2 z, const float2 p, const float2 v) {
    const float sx = z.x + p.x, sy = z.y - p.y, dx = z.x - p.x, dy = z.y + p.y;
    return make_float2(fmaf(v.x, dx, fmaf(-v.y, dy, sx)), fmaf(v.x, dy, fmaf(v.y, dx, sy)));
}
// Z' = (Y + conj P) + conj(V) (Y - conj P)   (inverse merge)
__device__ __forceinline__ float2 merge_inv(const float2 y, const float2 p, const float2 v) {
    const float sx = y.x + p.x, sy = y.y - p.y, dx = y.x - p.x, dy = y.y + p.y;
    return make_float2(fmaf(v.x, dx, fmaf(v.y, dy, sx)), fmaf(v.x, dy, fmaf(-v.y, dx, sy)));
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
template <bool PRUNED, bool SWAP>
__device__ __forceinline__ void inv_merge(float2 (&a)[32], float ynyq, float* scratch,
                                          const float2* __restrict__ vtab, int lane) {
    if (lane == 0) a[0].y = 0.0f;
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
            float2 z;
            if (r < 10) {
                float2 p = make_float2(0.0f, 0.0f);
                if (r == 0 && lane == 0) p = sc[704];
                z = (r == 0) ? merge_inv(a[r], p, v)
                             : make_float2(fmaf(v.x, a[r].x, fmaf(v.y, a[r].y, a[r].x)), fmaf(v.x, a[r].y, fmaf(-v.y, a[r].x, a[r].y)));
            } else if (r == 10) {
                float2 p = make_float2(0.0f, 0.0f);
                if (lane > 0) p = sc[1024 - k];
                z = merge_inv(a[r], p, v);
            } else if (r < kPrunedRows) {
                z = merge_inv(a[r], sc[1024 - k], v);
            } else {
                // Y = 0: Z' = conj(P) - conj(V) conj(P)
                const float2 p = sc[1024 - k];
                z = make_float2(fmaf(-v.x, p.x, fmaf(v.y, p.y, p.x)), fmaf(v.x, p.y, fmaf(v.y, p.x, -p.y)));
            }
            a[r] = SWAP ? make_float2(z.y, z.x) : z;
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
            const float2 z = merge_inv(a[r], p, vtab[k]);
            a[r] = SWAP ? make_float2(z.y, z.x) : z;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float2 z = merge_inv(a[r], t[r], vtab[32 * r + lane]);
            a[r] = SWAP ? make_float2(z.y, z.x) : z;
        }
        __syncwarp();
    }
}

// The 1024-point complex forward DFT of the per-lane arrays, in place: in a[j] = z[lane + 32*j], out
// a[j] = Z[lane + 32*j] -- input and output use the same (index mod 32, index div 32) layout, which is
// what lets one routine serve both directions.  The in-lane FFT is emitted once and run twice.
__device__ __forceinline__ void fwd1024(float2 (&a)[32], float* scratch, const float2* __restrict__ tw, int lane) {
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        fft32<32, false>(a);
        if (h == 0) {
#pragma unroll
            for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], tw[r * 32 + lane]);
            warp_transpose(a, scratch, lane);
        }
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
