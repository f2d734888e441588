// Warp-level 2048-point real FFT / inverse: one warp per frame, 32 complex values per lane.
//
// A real frame x[0..2048) is packed as z[n] = x[2n] + i x[2n+1] (1024 complex points) and
// transformed as a 32 x 32 Cooley-Tukey: an in-lane FFT-32 (fft32.cuh), one twiddle multiply,
// one 32x32 transpose through a warp-private shared-memory scratch, a second in-lane FFT-32.
// The real-signal spectrum is then recovered from Z[k] and conj(Z[1024-k]) (the partner lives in
// lane (32-k2)%32, so the pair is exchanged through the same scratch, stored linearly in k).
//
// Layouts (lane = threadIdx.x & 31):
//   time      a[n2] = z[lane + 32*n2]        (samples 2*lane + 64*n2 and +1)
//   frequency a[k1] = X[32*k1 + lane],  plus the Nyquist bin X[1024] (real) in lane 0
//
// Scratch per warp: 32 rows of 34 float2 (row pitch 272 B keeps both the 8-byte column writes and
// the 16-byte row reads bank-conflict free) = 8704 B; the linear [1025] exchange array aliases it.
#pragma once
#include "fft32.cuh"

namespace s2st {

constexpr int kScratchPitch = 34;                    // float2 per transpose row
constexpr int kScratchFloat2 = 32 * kScratchPitch;   // 1088 float2 = 8704 B per warp

__device__ __forceinline__ float2 cmul(const float2 a, const float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(const float2 a, const float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// 32x32 transpose of the per-lane register arrays: out a[r] = (lane r's) a[lane].
__device__ __forceinline__ void warp_transpose(float2 (&a)[32], float2* scratch, int lane) {
#pragma unroll
    for (int r = 0; r < 32; ++r) scratch[r * kScratchPitch + lane] = a[r];
    __syncwarp();
    const float4* row = reinterpret_cast<const float4*>(scratch + lane * kScratchPitch);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float4 v = row[r];
        a[2 * r] = make_float2(v.x, v.y);
        a[2 * r + 1] = make_float2(v.z, v.w);
    }
    __syncwarp();
}

// Forward, first half: in-lane FFT over n2, twiddle, transpose.  In: a[n2] for n2 < NZ (others ignored).
template <int NZ>
__device__ __forceinline__ void frame_fwd_a(float2 (&a)[32], float2* scratch, const float2* __restrict__ tw, int lane) {
    fft32<NZ, false>(a);
#pragma unroll
    for (int r = 1; r < 32; ++r) a[r] = cmul(a[r], tw[r * 32 + lane]);
    warp_transpose(a, scratch, lane);
}

// Forward, second half.  Out: a[k1] = 2*X[32*k1+lane] for 32*k1 < kb (the other registers are left
// undefined), nyq = 2*X[1024] (valid in lane 0).  vtab[k] = -i*exp(-2*pi*i*k/2048).
__device__ __forceinline__ void frame_fwd_b(float2 (&a)[32], float& nyq, float2* scratch,
                                            const float2* __restrict__ vtab, int lane, int kb) {
    fft32<32, false>(a);
    // pair exchange: Zs[k] = Z[k], Zs[1024] = Z[0]
#pragma unroll
    for (int r = 0; r < 32; ++r) scratch[32 * r + lane] = a[r];
    if (lane == 0) scratch[1024] = a[0];
    __syncwarp();
    nyq = 2.0f * (a[0].x - a[0].y);
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        if (32 * r >= kb) continue;  // bins the caller does not need (warp-uniform)
        const int k = 32 * r + lane;
        const float2 p = scratch[1024 - k];  // Z[1024-k]; conj applied below
        const float2 v = vtab[k];
        const float sx = a[r].x + p.x, sy = a[r].y - p.y;  // Z + conj(Zp)
        const float dx = a[r].x - p.x, dy = a[r].y + p.y;  // Z - conj(Zp)
        a[r] = make_float2(fmaf(v.x, dx, fmaf(-v.y, dy, sx)), fmaf(v.x, dy, fmaf(v.y, dx, sy)));
    }
    __syncwarp();
}

// Forward = both halves.  tw[r*32 + lane] = exp(-2*pi*i*r*lane/1024).
template <int NZ>
__device__ __forceinline__ void frame_fwd(float2 (&a)[32], float& nyq, float2* scratch,
                                          const float2* __restrict__ tw,
                                          const float2* __restrict__ vtab, int lane, int kb = 1024) {
    frame_fwd_a<NZ>(a, scratch, tw, lane);
    frame_fwd_b(a, nyq, scratch, vtab, lane, kb);
}

// Inverse.  In: a[k1] = Y[32*k1+lane] (Hermitian half-spectrum, imag of DC ignored), ynyq = Y[1024]
// (real, read from lane 0).  Out: a[n2] = 2048 * y[2*(lane+32*n2)] + i * 2048 * y[..+1].
__device__ __forceinline__ void frame_inv(float2 (&a)[32], float ynyq, float2* scratch,
                                          const float2* __restrict__ tw,
                                          const float2* __restrict__ vtab, int lane) {
    if (lane == 0) a[0].y = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r) scratch[32 * r + lane] = a[r];
    if (lane == 0) scratch[1024] = make_float2(ynyq, 0.0f);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        const int k = 32 * r + lane;
        const float2 p = scratch[1024 - k];
        const float2 v = vtab[k];  // U[k] = conj(V[k])
        const float sx = a[r].x + p.x, sy = a[r].y - p.y;
        const float dx = a[r].x - p.x, dy = a[r].y + p.y;
        // S + conj(V) * D
        a[r] = make_float2(fmaf(v.x, dx, fmaf(v.y, dy, sx)), fmaf(v.x, dy, fmaf(-v.y, dx, sy)));
    }
    __syncwarp();
    fft32<32, true>(a);
#pragma unroll
    for (int r = 1; r < 32; ++r) a[r] = cmul_conj(a[r], tw[r * 32 + lane]);
    warp_transpose(a, scratch, lane);
    fft32<32, true>(a);
}

}  // namespace s2st
