// Warp-level 2048-point real FFT / inverse, second formulation: 32 lanes x (in-lane real FFT-64).
//
// frame_fft.cuh packs a real frame as 1024 complex points and recovers the real-signal spectrum from Z[k] and
// conj(Z[1024-k]) AFTER the last in-lane FFT -- the partner lives in another lane, so every transform pays a
// pair exchange through shared memory plus a table of split factors (284 of the 888 shared-memory wavefronts of
// a Griffin-Lim frame-iteration).  Here the real-signal symmetry is used one stage earlier, where it is an
// in-lane operation with compile-time constants:
//
//   time      lane l holds x[l + 32 j], j = 0..63, as pairs  a[j'] = (x[l + 64 j'], x[l + 32 + 64 j'])
//   stage 1   in-lane complex FFT-32 over j' + in-lane split  ->  A_l[m] = sum_j x[l + 32 j] W64^(j m),  m = 0..32
//             (a real FFT-64; A_l[0], A_l[32] are real and stay packed as Z_l[0] = (u[l], u[l+32]),
//              u[s] = sum_j x[s + 64 j])
//   twiddle   B_l[m] = A_l[m] W2048^(l m),  m = 1..31
//   transpose lane m receives B_l[m] for all l (m = 1..31); lane 0 receives z[n] = (u[2n], u[2n+1])
//   stage 2   in-lane complex FFT-32 over l.  The twiddle table carries an extra factor W32^(-11 l), which moves output
//             q of lane m >= 1 to slot q + 11 (mod 32):  slot s holds X[m + 64 (s - 11)]; for s <= 10 that is index
//             2048 - (64 (11 - s) - m), i.e. the conjugate of bin 64 (11 - s) - m.  Lane 0 (not twiddled) gets the
//             packed real FFT-64 of u, whose split / merge is done by the whole warp (r64_column_mid).
//
// so bins k with k mod 64 in [1, 31] sit in lane (k mod 64), bins with k mod 64 in [33, 63] in lane 64 - (k mod 64)
// as conjugates, the multiples of 32 in lane 0, and with live bins < 704 (the pruned vocoder case, the only one
// implemented; input rows j' < 19, i.e. window support <= 1216 samples) every lane's live slots are 0..21:
//   lane m >= 1:  slot s <= 10: bin 64 (11 - s) - m (conjugated);  slot s >= 11: bin 64 (s - 11) + m
//   column U:     bin 32 s, stored as "lane 0, slot s" in the magnitude rows
// There is no lane-dependent code.  The inverse runs the same steps backwards.  The Griffin-Lim kernel keeps its
// target magnitudes in this order ([frame][slot][lane], written directly by the inverse-mel projection with a
// permuted basis).
#pragma once
#include "frame_fft.cuh"

namespace s2st {

constexpr int kR64Live = 22;  // live slots per lane
constexpr int kR64Pitch = 272;                         // bytes per row of the transpose scratch (see r64_transpose)
constexpr int kR64ScratchFloats = 32 * kR64Pitch / 4;  // 2176

// cos(2 pi j / 64), j = 0..16
__host__ __device__ constexpr float quarter_cos64(int j) {
    constexpr float t[17] = {1.0f,
                             0.99518472667219688624f,
                             0.98078528040323044913f,
                             0.95694033573220886494f,
                             0.92387953251128675613f,
                             0.88192126434835502971f,
                             0.83146961230254523708f,
                             0.77301045336273696081f,
                             0.70710678118654752440f,
                             0.63439328416364549822f,
                             0.55557023301960222474f,
                             0.47139673682599764856f,
                             0.38268343236508977173f,
                             0.29028467725446236764f,
                             0.19509032201612826785f,
                             0.09801714032956060199f,
                             0.0f};
    return t[j];
}
__host__ __device__ constexpr float cos64(int j) {
    j &= 63;
    return j <= 16 ? quarter_cos64(j) : j <= 32 ? -quarter_cos64(32 - j) : j <= 48 ? -quarter_cos64(j - 32)
                                                                                  : quarter_cos64(64 - j);
}
__host__ __device__ constexpr float sin64(int j) { return cos64(j + 48); }

// lo = s + v d, hi = s - v d for a compile-time complex v = (VX, VY): 3 packed instructions ("factor out the larger
// component" form of fft32.cuh's butterfly).  LO_ONLY skips hi.
template <bool LO_ONLY = false>
__device__ __forceinline__ void rot_bfly(const float vx, const float vy, const float2 s, const float2 d, float2& lo,
                                         float2& hi) {
    if ((vx < 0 ? -vx : vx) >= (vy < 0 ? -vy : vy)) {
        const float2 t = fma2(mul_pi(d), bcast2(vy / vx), d);  // d + (vy/vx) i d
        lo = fma2(t, bcast2(vx), s);
        if (!LO_ONLY) hi = fma2(t, bcast2(-vx), s);
    } else {
        const float2 t = fma2(d, bcast2(vx / vy), mul_pi(d));  // (vx/vy) d + i d
        lo = fma2(t, bcast2(vy), s);
        if (!LO_ONLY) hi = fma2(t, bcast2(-vy), s);
    }
}

// Pair (M, 32 - M) of the in-lane real-FFT-64 split (SYNTH = false) or Hermitian merge (SYNTH = true):
//   s = x + conj(y), d = x - conj(y),  out_m = s + v d,  out_n = conj(s - v d)
//   v = -i W64^M = (-sin, -cos)(2 pi M / 64)   (split: out = 2 A[M], 2 A[32-M] from x = Z[M], y = Z[32-M])
//   v = +i W64^-M = (-sin, +cos)               (merge: out = Z'[M], Z'[32-M] from x = D[M], y = D[32-M])
// Y_ZERO: y is structurally zero; LO_ONLY: out_n is not needed.
template <int M, bool SYNTH, bool Y_ZERO = false, bool LO_ONLY = false>
__device__ __forceinline__ void r64_pair(const float2 x, const float2 y, float2& out_m, float2& out_n) {
    static_assert(M >= 1 && M <= 15, "pair index");
    constexpr float vx = -sin64(M), vy = SYNTH ? cos64(M) : -cos64(M);
    float2 s = x, d = x;
    if constexpr (!Y_ZERO) {
        s = add2(x, conj2(y));
        d = add2(x, neg2(conj2(y)));
    }
    float2 hi = make_float2(0.0f, 0.0f);
    rot_bfly<LO_ONLY>(vx, vy, s, d, out_m, hi);
    if constexpr (!LO_ONLY) out_n = conj2(hi);
}

template <int M, bool SYNTH>
struct R64PairLoop {
    static __device__ __forceinline__ void run(float2 (&a)[32]) {
        r64_pair<M, SYNTH>(a[M], a[32 - M], a[M], a[32 - M]);
        if constexpr (M < 15) R64PairLoop<M + 1, SYNTH>::run(a);
    }
};

// In-lane split after the first FFT-32 (natural slots): a[m] <- 2 A[m] for m = 1..31 (a[16] up to a factor 2, which
// the re-normalisation of every bin makes irrelevant); a[0] = Z[0] = (u[l], u[l+32]) is left alone.
__device__ __forceinline__ void r64_split_mid(float2 (&a)[32]) {
    R64PairLoop<1, false>::run(a);
    a[16] = conj2(a[16]);
}
// In-lane Hermitian merge before the last inverse FFT-32: a[m] = D[m] (m = 1..31), a[0] = Z'[0] given.
__device__ __forceinline__ void r64_merge_mid(float2 (&a)[32]) {
    R64PairLoop<1, true>::run(a);
    a[16] = mul2(conj2(a[16]), bcast2(2.0f));
}

// ---- the column of bins that are multiples of 32 ---------------------------------------------------------------
// U[p] = X[32 p] is the real FFT-64 of u (see the header), whose samples arrive as slot 0 of every lane: (u[l], u[l+32]).
// Packed as z[n] = (u[2n], u[2n+1]) it is one more 32-point column, and lane 0 -- which has no column of its own --
// runs it through the common FFT-32 code for free.  What must NOT happen in lane 0 is the real-signal split / merge
// of that column: as lane-dependent code it costs ~120 predicated instructions plus as many register moves per
// frame (measured: slower than the packed-complex kernel).  Instead lane 0 publishes its 32 outputs to a 256-byte
// buffer, every lane p takes Zhat[p] and Zhat[32 - p] and does split, magnitude re-imposition (exact reference
// semantics) and Hermitian merge for ITS bin (the merge partner comes by shuffle), writes Zhat'[p] back, and lane 0
// reloads the 32 values in front of the common inverse FFT-32.  (A 32-point FFT across lanes by butterfly shuffles
// was also measured: its 12 dependent shuffle round trips stall the warp longer than this.)
// vp[p] = -i exp(-2 pi i p / 64).
__device__ __forceinline__ float2 shfl_idx2(const float2 v, const int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// m: the target magnitude of bin 32 * lane (0 for lane >= 22), loaded by the caller well ahead of time.
__device__ __forceinline__ void r64_column_mid(const float2 (&a)[32], float* __restrict__ upub,
                                               const float2* __restrict__ vp, const float m, int lane) {
    // entry q of the column buffer = the padding of scratch row q
    char* pub = reinterpret_cast<char*>(upub) + 256;
    const float2 vv = vp[lane];
    const int partner = (32 - lane) & 31;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 32; ++q) *reinterpret_cast<float2*>(pub + q * kR64Pitch) = a[q];
    }
    __syncwarp();
    float2* mine = reinterpret_cast<float2*>(pub + lane * kR64Pitch);
    float2 v = split_fwd(*mine, *reinterpret_cast<const float2*>(pub + partner * kR64Pitch), vv);  // 2 U[p]
    {
        const float r2 = fmaf(v.x, v.x, v.y * v.y);
        float rs;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(r2));
        const float2 sa = mul2(v, bcast2(m * rs));
        v = r2 >= 1.1754944e-38f ? sa : make_float2(copysignf(m, v.x), 0.0f);
    }
    // the shuffle doubles as the barrier between the reads above and the write below: a lane's shuffle input
    // depends on both of its reads, and nobody gets a shuffle result before every lane has delivered its input
    float2 y = shfl_idx2(v, partner);
    if (lane == 0) y = make_float2(0.0f, 0.0f);  // bin 0 pairs with the Nyquist bin, which is not live
    *mine = merge_inv<false>(v, y, vv);
    __syncwarp();
}
// Lane 0 takes the merged column back (natural slots) in front of the inverse FFT-32.  Predicated loads written in
// PTX with read-write operands: the values land in the registers the other lanes keep theirs in (an `if (lane == 0)`
// block makes ptxas load into fresh registers and copy 64 of them at the join).
template <int Q>
struct R64Reload {
    static __device__ __forceinline__ void run(float2 (&a)[32], unsigned addr, int lane) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, 0;\n\t@p ld.shared.v2.f32 {%0, %1}, [%3+%4];\n\t}"
                     : "+f"(a[Q].x), "+f"(a[Q].y)
                     : "r"(lane), "r"(addr), "n"(Q * kR64Pitch + 256));
        if constexpr (Q < 31) R64Reload<Q + 1>::run(a, addr, lane);
    }
};
__device__ __forceinline__ void r64_column_reload(float2 (&a)[32], const float* __restrict__ upub, int lane) {
    R64Reload<0>::run(a, (unsigned)__cvta_generic_to_shared(upub), lane);
}

// 32x32 transpose of (re, im) pairs through the warp's scratch: out a[c] = (lane c's) a[lane]; BR: deliver element c
// into slot brev5(c).  Rows are padded to kR64Pitch = 272 bytes instead of XOR-swizzled: the 8-byte column writes of a
// row are contiguous and the 16-byte row reads of 8 consecutive lanes start 272 bytes apart (banks 4 l), so both
// directions are conflict-free AND every address is one per-lane base plus a compile-time offset -- the swizzle cost
// ~45 address instructions per transpose, which matters here: the frame loop is straight-line code sitting right at
// the 32 KB instruction cache, and it runs ~15 % slower once it grows past it.  The 16 padding bytes of row q hold
// entry q of the column buffer (r64_column_mid).
// FWD (analysis): slot 0 is not part of the matrix: its halves (u[l], u[l+32]) go to floats l and 32 + l of row 0,
//   which lane 0 then reads as its column z[n] = (u[2n], u[2n+1]) with the same row loads as everybody else.
// !FWD (synthesis): lane 0 writes z'[j] = (u'[2j], u'[2j+1]) into column 0 like any other lane, and reader l fetches
//   its slot 0 = (u'[l], u'[l+32]) from column 0 of rows l/2 and 16 + l/2.
template <bool FWD, bool BR>
__device__ __forceinline__ void r64_transpose(float2 (&a)[32], float* scratch, int lane) {
    char* col = reinterpret_cast<char*>(scratch) + lane * 8;
    if constexpr (FWD) {
        scratch[lane] = a[0].x;
        scratch[32 + lane] = a[0].y;
    }
#pragma unroll
    for (int j = FWD ? 1 : 0; j < 32; ++j) *reinterpret_cast<float2*>(col + j * kR64Pitch) = a[j];
    __syncwarp();
    const char* row = reinterpret_cast<const char*>(scratch) + lane * kR64Pitch;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(row + 16 * c);
        a[BR ? brev5(2 * c) : 2 * c] = make_float2(v.x, v.y);
        a[BR ? brev5(2 * c + 1) : 2 * c + 1] = make_float2(v.z, v.w);
    }
    if constexpr (!FWD) {
        const float* sc = scratch + (lane >> 1) * (kR64Pitch / 4) + (lane & 1);  // (row l/2, column 0)
        a[0] = make_float2(sc[0], sc[16 * (kR64Pitch / 4)]);
    }
    __syncwarp();
}

// Constant tables (shared memory) and per-warp buffers of the transform.
struct R64Ctx {
    const float2* tw;    // [32*32] exp(-2 pi i (m l - 704 l) / 2048) at [m * 32 + l]
    const float2* vp;    // [32]    -i exp(-2 pi i p / 64)
    float* scratch;      // kR64ScratchFloats per warp (the row padding doubles as the column buffer)
};

// Analysis.  In: a[brev5(j')] = windowed (x[l + 64 j'], x[l + 32 + 64 j']), j' < 19 (the other slots are ignored).
// Out: slots 0..21 as in the header for lanes >= 1 (scaled by 2: irrelevant to the phase); lane 0: slot p =
// Zhat[p], the packed FFT of the multiples-of-32 column (see r64_column_mid).
__device__ __forceinline__ void r64_analysis(float2 (&a)[32], const R64Ctx& ctx, int lane) {
    fft32_inplace_br<false, 19, 32>(a);
    r64_split_mid(a);
#pragma unroll
    for (int m = 1; m < 32; ++m) a[m] = cmul(a[m], ctx.tw[m * 32 + lane]);
    r64_transpose<true, true>(a, ctx.scratch, lane);
    fft32_inplace_br<false, 32, 32>(a);
}

// Synthesis.  In: lanes >= 1: slots 0..21 hold the (Hermitian-consistent) spectrum, slots >= 22 are taken as zero;
// lane 0's column comes from ctx.scratch (r64_column_mid).
// Out: a[j'] = (y[l + 64 j'], y[l + 32 + 64 j']) for j' < 19, y[n] = sum_{k < 2048} Y[k] exp(+2 pi i n k / 2048).
__device__ __forceinline__ void r64_synthesis(float2 (&a)[32], const R64Ctx& ctx, int lane) {
#pragma unroll
    for (int q = kR64Live; q < 32; ++q) a[q] = make_float2(0.0f, 0.0f);
    r64_column_reload(a, ctx.scratch, lane);
    float2 b[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) b[brev5(q)] = a[q];
    fft32_inplace_br<true, 32, 32>(b);
    r64_transpose<false, false>(b, ctx.scratch, lane);
#pragma unroll
    for (int m = 1; m < 32; ++m) b[m] = cmul_conj(b[m], ctx.tw[m * 32 + lane]);
    r64_merge_mid(b);
#pragma unroll
    for (int m = 0; m < 32; ++m) a[brev5(m)] = b[m];
    fft32_inplace_br<true, 32, 19>(a);
}

}  // namespace s2st
