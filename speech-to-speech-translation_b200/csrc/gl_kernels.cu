// Griffin-Lim synthesis kernels (sm_100a).
//
// Replaces GriffinLim.forward / GriffinLim.inverse / TTSSpectrogram.forward of the reference
// (fairseq/models/text_to_speech/vocoder.py:84-110, fairseq/data/audio/audio_utils.py:259-271):
// the reference runs each STFT / ISTFT as a dense 2050x2048 convolution plus a host-side Python
// loop for the window-sum-square; here one persistent kernel launch per Griffin-Lim iteration does
//   frame load (reflect) -> window -> rFFT-2048 -> magnitude re-imposition -> irFFT-2048 -> window
//   -> overlap-add -> 1/window-sum-square -> store
// for a ragged batch of utterances.  The only state between iterations is the waveform.
//
// Work decomposition: a STRIP = S consecutive frames of one utterance, owned by ONE warp, which runs
// them in order.  The warp keeps the overlap-add of its frames in a private shared-memory ring of
// `ws` samples: once frame f has been added, hop f of the strip is final (later frames start after
// it), so it is normalised and written out immediately and its ring slots are cleared for reuse.
// Warps never synchronise with each other: there is no __syncthreads in the main loop.
//
// Only the first / last (ws - hop) samples of a strip are shared with the neighbouring strip.  Such
// "seam" samples receive exactly two contributions, both added with red.global.add onto a zeroed
// location, so the result does not depend on arrival order (a + b == b + a) and the output stays
// bitwise deterministic.  Three waveform buffers rotate: pass i reads buf[(i-1)%3], writes buf[i%3]
// and zeroes the seams of buf[(i+1)%3] for the next pass.
#include <math_constants.h>

#include <algorithm>
#include <vector>
#ifdef S2ST_FRAMES_PROF
#include <cstdio>
#endif

#include "../../include/s2st_b200.h"
#include "frame_fft.cuh"
#include "plan.h"

namespace s2st {

namespace {

constexpr float kTiny = 1.1754944e-38f;  // vocoder.py:69
constexpr int kGlWarps = kGlThreads / 32;

// Counter-based uniform in [0, 1) for the device-side initial phase (used when the caller passes no phase):
// two rounds of a 64-bit mix (splitmix64 finaliser) of (seed, element index); 24 random bits.
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
}

// Frame load, cold path: reflect padding (audio_utils.py:262-263) at utterance edges / unaligned data.
template <int NZ>
__device__ __forceinline__ void load_frame_edge(float2 (&a)[32], const float* __restrict__ y, int j, int L) {
#pragma unroll 1
    for (int r = 0; r < NZ; ++r, j += 64) {
        int j0 = j < 0 ? -j : j, j1 = j + 1 < 0 ? -(j + 1) : j + 1;
        j0 = j0 >= L ? 2 * (L - 1) - j0 : j0;
        j1 = j1 >= L ? 2 * (L - 1) - j1 : j1;
        j0 = min(max(j0, 0), L - 1);  // only reachable where the window is zero
        j1 = min(max(j1, 0), L - 1);
        const float2 v = make_float2(y[j0], y[j1]);
        // registers cannot be indexed dynamically: scatter through a switch-free unrolled select
#pragma unroll
        for (int q = 0; q < NZ; ++q)
            if (q == r) a[brev5(q)] = v;  // frame row q lives in slot brev5(q) (fwd1024's input order)
    }
}

// Generic (cold) write-out of strip samples [i0, i0 + n): seams, utterance edges, unaligned layouts.
// Kept out of line; everything it needs is passed by value.
__device__ __noinline__ void emit_generic(float* __restrict__ ring, const float* __restrict__ s_inv_wss,
                                          const float* __restrict__ w2, float inv_nfft, float* __restrict__ out,
                                          float* __restrict__ znext, int hop, int ws, int f0, int nf, int T, int L,
                                          int j_base, int i0, int n, int slot0, int lane) {
    const bool has_prev = f0 > 0, has_next = f0 + nf < T;
    const int left_end = ws - hop;     // i < left_end : the previous strip's frames cover it too
    const int right_beg = nf * hop;    // i >= right_beg: the next strip's frames cover it too
    const int clip_beg = (T - f0) * hop;  // a frame index >= T would be needed from here on
    for (int e = lane; e < n; e += 32) {
        int slot = slot0 + e;
        while (slot >= ws) slot -= ws;
        const float acc = ring[slot];
        ring[slot] = 0.0f;
        const int i = i0 + e;
        const int j = j_base + i;
        if (j < 0 || j >= L) continue;
        const bool in_left = i < left_end, in_right = i >= right_beg;
        float inv;
        if ((in_left && !has_prev) || i >= clip_beg) {
            // utterance edges: only the frames that exist contribute (vocoder.py:78-81, frame order)
            const int t_lo = i < ws ? f0 - (ws - 1 - i) / hop : f0 + (i - ws) / hop + 1;
            const int t_hi = f0 + i / hop;
            const int qq = f0 * hop + i;
            float w = 0.0f;
            for (int t = max(t_lo, 0); t <= min(t_hi, T - 1); ++t) w += __ldg(w2 + (qq - t * hop));
            inv = (w > kTiny ? 1.0f / w : 1.0f) * inv_nfft;
        } else {
            inv = s_inv_wss[i % hop];
        }
        const float v = acc * inv;
        const bool right_seam = in_right && has_next;
        if (right_seam || (in_left && has_prev)) atomicAdd(out + j, v);
        else out[j] = v;
        if (right_seam) znext[j] = 0.0f;
    }
}

constexpr int kStdHop = 300, kStdWs = 1200, kStdRot = 424;  // the vocoder's standard geometry (see k_gl_pass STD)

// One Griffin-Lim pass.  FIRST: spectra come from (mag, initial phase) -> inverse only.
// PRUNED: every live bin is below 704 (kb <= 704): magnitudes are prefetched into registers and the
// pair exchanges move 22 rows instead of 32.
// STD: the vocoder's standard geometry (hop 300, window support 1200 starting at sample 424 of the 2048-point
// frame, magnitude rows at least 704 wide) as compile-time constants: no geometry tests in the frame loop.
// S2ST_GL_SHARED_FFT=1 builds the compact variant: analysis and synthesis share one copy of the 1024-point routine
// and that routine runs one copy of the in-lane FFT twice (23 KB hot loop instead of 40 KB).  Measured on B200 the
// straight-line variant is 4.5 % faster: its loop-carried register shuffles cost more than its I-cache misses.
#ifndef S2ST_GL_SHARED_FFT
#define S2ST_GL_SHARED_FFT 0
#endif
constexpr bool kGlSharedFft = S2ST_GL_SHARED_FFT != 0;
constexpr int kGlUnrollPass = kGlSharedFft ? 1 : 2;
// Two variants of this kernel were measured and removed in round 2 (profiles/r02_small_calls.txt, git history): a
// persistent launch of all iterations with per-strip neighbour flags (never faster: 15.82 vs 15.57 ms on the config-2
// batch, 2.64 vs 2.16 ms for one 500-frame utterance) and a "team" mode with four warps per strip for small calls
// (1.67 ms).  Small calls now run the frame-parallel kernel of gl_frames.cuh (0.50 ms), which uses the two helpers below.
__device__ __forceinline__ int ld_acquire(const int* ptr) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int* ptr) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* ptr, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

template <int NZ, bool FIRST, bool PRUNED, bool STD>
__global__ void __launch_bounds__(kGlThreads, 1) k_gl_pass(const __grid_constant__ GlParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);            // 1024
    float2* s_vtab = s_tw + 1024;                                  // 1024
    float* s_win_a = reinterpret_cast<float*>(s_vtab + 1024);
    float* s_inv_wss = s_win_a + 64 * NZ;                          // hop (rounded up to 4)
    const int hop = STD ? kStdHop : p.hop, ws = STD ? kStdWs : p.ws;
    const int rot_half = STD ? kStdRot - kNfft / 2 : p.rot - p.half;
    float* s_warp = s_inv_wss + ((hop + 3) & ~3);                  // per warp: scratch + ring
    const int ring_floats = (ws + 3) & ~3;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* scratch = s_warp + warp * (kScratchFloats + ring_floats);
    float* ring = scratch + kScratchFloats;
    for (int i = tid; i < 1024; i += kGlThreads) {
        s_tw[i] = p.tw[i];
        s_vtab[i] = p.vtab[i];
    }
    for (int i = tid; i < 64 * NZ; i += kGlThreads) s_win_a[i] = p.win_a[i];
    for (int i = tid; i < hop; i += kGlThreads) s_inv_wss[i] = p.inv_wss[i];
    for (int i = lane; i < ring_floats; i += 32) ring[i] = 0.0f;
    __syncthreads();  // the only block-wide barrier: constant tables are in place
    // Programmatic dependent launch: the pass is launched while the previous pass is still draining, so everything
    // above (CTA start-up, 21 KB of constant tables, ring clearing) overlaps its tail; from here on the previous
    // kernel's output is read.  (A no-op when the kernel is launched the ordinary way.)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int n_strips = *p.n_tiles;
    const int kb = PRUNED ? min(p.kb, 32 * kPrunedRows) : p.kb;
    // fast paths need 16-byte friendly geometry (true for hop 300 / win 1200 / n_fft 2048)
    const bool geom4 = STD || ((hop % 4 == 0) && (hop >= 64) && (ws % hop == 0) && (rot_half % 4 == 0));

    // Strips are dealt to SMs first, then to warps, and -- the table is sorted by descending length -- in SNAKE order
    // when there are more strips than warps: round 0 gives warp slot w strip w, round 1 strip 2 W - 1 - w, ... so the
    // warps that hold the shortest strips of one round get the longest of the next (the shortest-processing-time fold
    // of longest-first scheduling).  choose_strip() on the host evaluates exactly this assignment: on the config-2
    // batch S = 25 with 82 folded utterance tails (longest warp: 25 frames) instead of S = 26 with one strip per warp.
    const int w_slot = blockIdx.x + gridDim.x * warp, n_slots = gridDim.x * kGlWarps;
    for (int round = 0, strip = w_slot; strip < n_strips;
         ++round, strip = (round & 1) ? (round + 1) * n_slots - 1 - w_slot : round * n_slots + w_slot) {
        const TileDesc td = p.tiles[strip];
        UttDesc ud;
        ud.wave_off = td.wave_off;
        ud.frame_off = td.frame_off;
        ud.n_frames = td.n_frames;
        const int T = ud.n_frames, L = (T - 1) * hop;
        const int j_base = td.f0 * hop + rot_half;  // output sample index of strip-relative sample 0
        const bool aligned = STD || (geom4 && ((ud.wave_off & 3) == 0));  // STD: wave_off is a multiple of hop
        float* out = p.out + ud.wave_off;
        float* znext = p.zero_next + ud.wave_off;
        const float* y = p.in + ud.wave_off;
        const float* magrow = p.mag + ((size_t)ud.frame_off + td.f0) * p.mag_stride;
        const float* phrow = (FIRST && p.phase) ? p.phase + ((size_t)ud.frame_off + td.f0) * p.phase_stride : nullptr;

        float2 a[32];
        int slot0 = 0;  // ring slot of strip-relative sample f * hop
        // f = -1 only fetches frame 0; iteration f processes frame f and fetches frame f + 1, so the
        // (single) copy of the load code overlaps with the write-out of the previous hop.
        auto fetch_frame = [&](int fi) {
            const int jf = j_base + fi * hop;  // first sample of the frame (lane 0)
            if (aligned && jf >= 0 && jf + 64 * NZ <= L) {
                const float2* src = reinterpret_cast<const float2*>(y + jf) + lane;
#pragma unroll
                for (int r = 0; r < NZ; ++r) a[brev5(r)] = src[32 * r];
                // the hop the frame after that adds is not in cache yet: ask L2 for it now (11 lines)
                const int jp = jf + 64 * NZ - 16 + 32 * lane;
                if (lane < 11 && jp < L) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + jp));
            } else {
                load_frame_edge<NZ>(a, y, jf + 2 * lane, L);
            }
        };
#pragma unroll 1
        for (int f = FIRST ? 0 : -1; f < td.nf; ++f) {
            if (f >= 0) {
                float ynyq = 0.0f;
                float mg[kPrunedRows];
                if constexpr (!FIRST) {
                    const float* w = s_win_a + 2 * lane;
#pragma unroll
                    for (int r = 0; r < NZ; ++r) {
                        const float2 ww = *reinterpret_cast<const float2*>(w + 64 * r);
                        a[brev5(r)] = mul2(a[brev5(r)], ww);  // frame row r lives in slot brev5(r)
                    }
                    if constexpr (kGlSharedFft) {
#pragma unroll
                        for (int r = NZ; r < 32; ++r) a[brev5(r)] = make_float2(0.0f, 0.0f);
                    }
                    // target magnitudes: requested now, the forward transform hides their latency
                    if constexpr (PRUNED) {
#pragma unroll
                        for (int r = 0; r < kPrunedRows; ++r)
                            mg[r] = (STD || 32 * r < p.mag_stride) ? __ldcs(magrow + 32 * r + lane) : 0.0f;  // uniform test; read once per pass
                    }
                }
                // pass 0: analysis (frame -> spectrum -> re-imposed magnitude); pass 1: synthesis.  The
                // 1024-point transform is the same code in both directions (see inv_merge), emitted once.
#pragma unroll kGlUnrollPass
                for (int pass = FIRST ? 1 : 0; pass < 2; ++pass) {
                    if (pass == 1) {
                        if constexpr (FIRST) {
                            // spectrum from (magnitude, initial phase).  All loads of the frame are issued first (22-32
                            // magnitudes + phases per lane): with the loads inside the per-row code the pass was bound
                            // by global-load latency, 0.46 ms against 0.24 ms for a full iteration.  The phasor stays
                            // the accurate sincospif: a MUFU sin / cos pair (absolute error ~5e-7) is 0.03 ms faster, but
                            // the iterations amplify a perturbed start -- on the 4800-frame utterance the distance to the
                            // CPU restatement after 64 iterations went from ~1e-4 to 6.7e-4 of the 1e-3 budget (measured).
                            constexpr int kRows = PRUNED ? kPrunedRows : 32;
                            float mgv[kRows], phv[kRows];
                            if (f + 1 < td.nf) {
                                // the next frame's magnitude and phase rows (4.1 + 2.8 KB) are read exactly once, from
                                // DRAM: ask L2 for them now, one transform ahead of their use (without this the pass
                                // waits a full DRAM round trip per frame: 63 % of its stall samples)
                                const char* nm = reinterpret_cast<const char*>(magrow + p.mag_stride);
                                for (int o = 128 * lane; o < 4 * kb; o += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nm + o));
                                if (p.phase) {
                                    const char* np = reinterpret_cast<const char*>(phrow + p.phase_stride);
                                    for (int o = 128 * lane; o < 4 * kb; o += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + o));
                                }
                            }
#pragma unroll
                            for (int r = 0; r < kRows; ++r) {
                                const int k = 32 * r + lane;
                                mgv[r] = k < kb ? __ldg(magrow + k) : 0.0f;
                                if (p.phase) {
                                    phv[r] = k < kb ? __ldg(phrow + k) : 0.0f;  // radians
                                } else {  // phi = 2 pi u - pi, u ~ U[0, 1): same law as vocoder.py:103 (kept in units of pi)
                                    const unsigned long long e = ((unsigned long long)ud.frame_off + td.f0 + f) * kBins + k;
                                    phv[r] = 2.0f * uniform01(p.phase_seed, e) - 1.0f;
                                }
                            }
#pragma unroll
                            for (int r = 0; r < kRows; ++r) {
                                const int k = 32 * r + lane;
                                // the caller's phase refers to the un-rotated frame; frames are processed rotated by
                                // p.rot samples:  Y'[k] = Y[k] * exp(+2 pi i k rot / 2048).  The rotation angle (an exact
                                // multiple of pi / 1024, reduced to [-pi, pi)) is added to the phase before the one sincos
                                const float rot_pi = (float)(((k * p.rot + 1024) & 2047) - 1024) * (1.0f / 1024.0f);
                                // in units of pi: sincospi reduces its argument exactly (no slow path); the product
                                // phase / pi rounds to within 1e-7 rad
                                const float v = p.phase ? fmaf(phv[r], 0.31830988618379067154f, rot_pi) : phv[r] + rot_pi;
                                float sn, cs;
                                sincospif(v, &sn, &cs);
                                a[r] = make_float2(mgv[r] * cs, mgv[r] * sn);
                            }
                            if (kb > 1024) {
                                const unsigned long long e = ((unsigned long long)ud.frame_off + td.f0 + f) * kBins + 1024;
                                ynyq = __ldg(magrow + 1024) * (p.phase ? cosf(__ldg(phrow + 1024)) : cospif(2.0f * uniform01(p.phase_seed, e) - 1.0f));
                            }
                            if (p.phase) phrow += p.phase_stride;
                        }
                        inv_merge<PRUNED, true, true>(a, ynyq, scratch, s_vtab, lane);
                    }
                    if constexpr (kGlSharedFft) {
                        fwd1024<32, 32, true>(a, scratch, s_tw, lane);
                    } else {
                        // analysis: rows >= NZ of the frame are zero; synthesis: only rows < NZ are overlap-added
                        if (pass == 0) fwd1024<(NZ > 16 ? NZ : 32), 32>(a, scratch, s_tw, lane);
                        else fwd1024<32, (NZ > 16 ? NZ : 32)>(a, scratch, s_tw, lane);
                    }
                    if constexpr (!FIRST) {
                        if (pass == 0) {
                            float nyq;
                            fwd_split<PRUNED>(a, nyq, scratch, s_vtab, lane);
                            if constexpr (PRUNED) {
                                // scale factors first (into mg[]); spectra whose |X|^2 is below the normal range
                                // (where the reference's atan2(0, +-0) = 0 / pi gives (+-mag, 0)) are only flagged
                                bool degenerate = false;
#pragma unroll
                                for (int r = 0; r < kPrunedRows; ++r) {
                                    const float r2 = fmaf(a[r].x, a[r].x, a[r].y * a[r].y);
                                    float rs;
                                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(r2));
                                    mg[r] *= rs;
                                    degenerate |= !(r2 >= 1.1754944e-38f);
                                }
                                if (__builtin_expect(__any_sync(0xffffffffu, degenerate), 0)) {
#pragma unroll 1
                                    for (int r = 0; r < kPrunedRows; ++r) {
                                        const float m = __ldg(magrow + 32 * r + lane);
                                        float2 v = make_float2(0.0f, 0.0f);
                                        float sc = 0.0f;
#pragma unroll
                                        for (int q = 0; q < kPrunedRows; ++q)
                                            if (q == r) {
                                                v = a[q];
                                                sc = mg[q];
                                            }
                                        const float r2 = fmaf(v.x, v.x, v.y * v.y);
                                        v = r2 >= 1.1754944e-38f ? mul2(v, bcast2(sc)) : make_float2(copysignf(m, v.x), 0.0f);
#pragma unroll
                                        for (int q = 0; q < kPrunedRows; ++q)
                                            if (q == r) a[q] = v;
                                    }
                                } else {
#pragma unroll
                                    for (int r = 0; r < kPrunedRows; ++r) a[r] = mul2(a[r], bcast2(mg[r]));
                                }
                            } else {
#pragma unroll
                                for (int r = 0; r < 32; ++r) {
                                    const int k = 32 * r + lane;
                                    const float m = (k < kb) ? __ldg(magrow + k) : 0.0f;
                                    const float x = a[r].x, yy = a[r].y;
                                    const float r2 = fmaf(x, x, yy * yy);
                                    float rs;
                                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(r2));
                                    const float sc = m * rs;
                                    const float2 sa = mul2(a[r], bcast2(sc));
                                    a[r] = r2 >= 1.1754944e-38f ? sa : make_float2(copysignf(m, x), 0.0f);
                                }
                            }
                            if (kb > 1024) ynyq = copysignf(__ldg(magrow + 1024), nyq);
                        }
                    }
                }
                magrow += p.mag_stride;
                // a[] holds the synthesis frame with the parts swapped (.y = even sample, .x = odd sample):
                // window and overlap-add into the private ring
                const float* w = s_win_a + 2 * lane;
                if (geom4) {
                    // every row accumulates without a bounds test (slots stay even -> 8-byte RMW); only the lanes of
                    // the last row that lie past the window support are predicated off
                    // (Measured and rejected, round 2: issuing the ring / window loads of a batch of rows ahead of the
                    // stores -- the compiler serialises them as written, possible alias -- costs registers the loop does
                    // not have: 20-32 bytes of spills, 0.2425 ms per pass instead of 0.2369; specialising the section on
                    // the four ring positions, immediate offsets and no wrap arithmetic, quadruples it and pushes the hot
                    // loop past the instruction cache: 0.2450 ms.)
                    const int first = slot0 + 2 * lane;
                    const int until_wrap = ws - first;  // rows with 64 r >= until_wrap wrap around once
#pragma unroll
                    for (int r = 0; r < NZ; ++r) {
                        // the lanes of the last row that lie past the window support (zero weight) would wrap onto slots
                        // the first row's lanes have just updated: skipped (for STD this is one predicate, on row 18)
                        if (64 * r + 62 >= ws && 64 * r + 2 * lane >= ws) continue;
                        float* dst = ring + first + 64 * r - (64 * r >= until_wrap ? ws : 0);
                        float2 o = *reinterpret_cast<float2*>(dst);
                        const float2 ww = *reinterpret_cast<const float2*>(w + 64 * r);
                        *reinterpret_cast<float2*>(dst) = fma2(swap2(a[r]), ww, o);
                    }
                } else {
#pragma unroll 1
                    for (int r = 0; r < NZ; ++r) {
                        float ex = 0.0f, ey = 0.0f;
#pragma unroll
                        for (int q = 0; q < NZ; ++q)
                            if (q == r) {
                                ex = a[q].y;
                                ey = a[q].x;
                            }
                        const int m = 64 * r + 2 * lane;
                        if (m < ws) {
                            int slot = slot0 + m;
                            while (slot >= ws) slot -= ws;
                            int slot1 = slot + 1;
                            if (slot1 >= ws) slot1 -= ws;
                            ring[slot] = fmaf(ex, w[64 * r], ring[slot]);
                            if (m + 1 < ws) ring[slot1] = fmaf(ey, w[64 * r + 1], ring[slot1]);
                        }
                    }
                }
            }
            // fetch the next frame (the loads fly while the finished hop is written out)
            if constexpr (!FIRST) {
                if (f + 1 < td.nf) fetch_frame(f + 1);
            }
            if (f >= 0) {
                __syncwarp();
                // hop f of the strip is final: normalise, store, clear its ring slots
                const int i0 = f * hop;
                const int j0 = j_base + i0;
                if (aligned && i0 >= ws - hop && j0 >= 0 && j0 + hop <= L) {
                    // steady state: no seam, no edge -> 16-byte wide, straight stores
                    float4* dst = reinterpret_cast<float4*>(out + j0);
                    float4* rg = reinterpret_cast<float4*>(ring + slot0);
                    const float4* iw = reinterpret_cast<const float4*>(s_inv_wss);
                    for (int q = lane; q < (hop >> 2); q += 32) {
                        const float4 v = rg[q], wv = iw[q];
                        rg[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        const float2 lo = mul2(make_float2(v.x, v.y), make_float2(wv.x, wv.y));
                        const float2 hi = mul2(make_float2(v.z, v.w), make_float2(wv.z, wv.w));
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                } else if (STD && td.f0 > 0 && i0 < ws - hop && j0 + hop <= L && i0 + hop <= (T - td.f0) * hop) {
                    // left seam of an interior strip, steady-state window sum: the previous strip adds the other
                    // part of these samples; two contributions onto a zeroed location commute -> deterministic
                    float4* dst = reinterpret_cast<float4*>(out + j0);
                    float4* rg = reinterpret_cast<float4*>(ring + slot0);
                    const float4* iw = reinterpret_cast<const float4*>(s_inv_wss);
                    for (int q = lane; q < (hop >> 2); q += 32) {
                        const float4 v = rg[q], wv = iw[q];
                        rg[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        atomicAdd(dst + q, make_float4(v.x * wv.x, v.y * wv.y, v.z * wv.z, v.w * wv.w));
                    }
                } else {
                    emit_generic(ring, s_inv_wss, p.w2, p.inv_nfft, out, znext, hop, ws, td.f0, td.nf, T, L, j_base, i0, hop, slot0, lane);
                }
                slot0 += hop;
                if (slot0 >= ws) slot0 -= ws;
                __syncwarp();
            }
        }
        // the tail of the last frame: [nf*hop, (nf-1)*hop + ws)
        if (STD && T - td.f0 - td.nf >= 3) {
            // right seam of an interior strip (the next strip has at least three frames, so the window sum is the
            // steady one): add our part, and clear the same samples of the buffer the NEXT pass accumulates into
            const float4* iw = reinterpret_cast<const float4*>(s_inv_wss);
#pragma unroll 1
            for (int h = 0; h < (kStdWs - kStdHop) / kStdHop; ++h) {
                const int j0 = j_base + (td.nf + h) * hop;
                float4* dst = reinterpret_cast<float4*>(out + j0);
                float4* zn = reinterpret_cast<float4*>(znext + j0);
                float4* rg = reinterpret_cast<float4*>(ring + slot0);
                for (int q = lane; q < (hop >> 2); q += 32) {
                    const float4 v = rg[q], wv = iw[q];
                    rg[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    atomicAdd(dst + q, make_float4(v.x * wv.x, v.y * wv.y, v.z * wv.z, v.w * wv.w));
                    zn[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
                slot0 += hop;
                if (slot0 >= ws) slot0 -= ws;
            }
        } else {
            emit_generic(ring, s_inv_wss, p.w2, p.inv_nfft, out, znext, hop, ws, td.f0, td.nf, T, L, j_base, td.nf * hop, ws - hop, slot0, lane);
        }
        __syncwarp();
    }
}

// Strip table.  Strips are first listed utterance by utterance (tiles_tmp), then written to `tiles` ordered by
// DESCENDING length (a counting sort on nf <= kMaxStrip): warps take strips round-robin, SM first, so every SM gets the
// same mix of full strips and short utterance tails.  In utterance order the tails land on random SMs and the SMs
// without one (16 full strips) finish last: 416 frames against an average of 394 on the config-2 batch.  The position
// inside a length class is claimed with an atomic counter -- which warp runs a strip has no influence on the result.
constexpr int kMaxStrip = 64;
__global__ void __launch_bounds__(1024) k_build_tiles(const int32_t* __restrict__ fo, int n_utts, int hop, int S,
                                                       UttDesc* __restrict__ utts, TileDesc* __restrict__ tiles_tmp,
                                                       TileDesc* __restrict__ tiles, int* __restrict__ tile_pos,
                                                       int* __restrict__ n_tiles) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    __shared__ int s_hist[kMaxStrip + 1], s_start[kMaxStrip + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    if (tid <= kMaxStrip) s_hist[tid] = 0;
    __syncthreads();
    const bool sorted = S <= kMaxStrip;
    TileDesc* listing = sorted ? tiles_tmp : tiles;
    for (int base = 0; base < n_utts; base += 1024) {
        const int u = base + tid;
        const int T = u < n_utts ? fo[u + 1] - fo[u] : 0;
        const int nt = (T + S - 1) / S;
        int incl = nt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += v;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int first = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - nt;
        if (u < n_utts) {
            UttDesc d;
            d.wave_off = (long long)(fo[u] - u) * hop;
            d.frame_off = fo[u];
            d.n_frames = T;
            utts[u] = d;
            for (int k = 0; k < nt; ++k) {
                TileDesc t;
                t.wave_off = d.wave_off;
                t.frame_off = d.frame_off;
                t.n_frames = T;
                t.f0 = k * S;
                t.nf = min(S, T - k * S);
                t.prev = k > 0 ? first + k - 1 : -1;   // listing indices; remapped below when the table is sorted
                t.next = k + 1 < nt ? first + k + 1 : -1;
                listing[first + k] = t;
            }
            if (sorted && nt > 0) {
                if (nt > 1) atomicAdd(&s_hist[S], nt - 1);
                atomicAdd(&s_hist[T - (nt - 1) * S], 1);
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += s_warp[31];
        __syncthreads();
    }
    const int total = s_carry;
    if (tid == 0) *n_tiles = total;
    if (!sorted) return;
    if (tid == 0) {
        int acc = 0;
        for (int len = kMaxStrip; len >= 0; --len) {
            s_start[len] = acc;
            acc += s_hist[len];
        }
    }
    __syncthreads();
    for (int i = tid; i < total; i += 1024) {
        const TileDesc t = tiles_tmp[i];
        const int np = atomicAdd(&s_start[t.nf], 1);
        tile_pos[i] = np;
        tiles[np] = t;
    }
    __syncthreads();
    for (int i = tid; i < total; i += 1024) {
        TileDesc* t = tiles + tile_pos[i];
        if (t->prev >= 0) t->prev = tile_pos[t->prev];
        if (t->next >= 0) t->next = tile_pos[t->next];
    }
}

// mag[t, f] = max(0, sum_m inv_mel[f, m] * g(mel[t, m]))     (vocoder.py:42, 141)
constexpr int kImFrames = 16;
__global__ void __launch_bounds__(256) k_inverse_mel(const float* __restrict__ logmel, bool is_log, long long n_frames,
                                                      int n_mels, const float* __restrict__ inv_mel_t,
                                                      int kb, int kb_pad, float* __restrict__ mag,
                                                      int out_stride, int n_out) {
    extern __shared__ float s_e[];  // [kImFrames][n_mels]
    const long long t0 = (long long)blockIdx.x * kImFrames;
    const int nt = (int)min((long long)kImFrames, n_frames - t0);
    for (int i = threadIdx.x; i < kImFrames * n_mels; i += blockDim.x) {
        const int t = i / n_mels;
        s_e[i] = t < nt ? (is_log ? expf(logmel[t0 * n_mels + i]) : logmel[t0 * n_mels + i]) : 0.0f;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < n_out; f += blockDim.x) {
        float acc[kImFrames];
#pragma unroll
        for (int t = 0; t < kImFrames; ++t) acc[t] = 0.0f;
        if (f < kb) {
            for (int m = 0; m < n_mels; ++m) {
                const float w = __ldg(inv_mel_t + (size_t)m * kb_pad + f);
#pragma unroll
                for (int t = 0; t < kImFrames; ++t) acc[t] = fmaf(w, s_e[t * n_mels + m], acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < kImFrames; ++t)
            if (t < nt) mag[(t0 + t) * out_stride + f] = fmaxf(acc[t], 0.0f);
    }
}

// batched rfft / irfft of full 2048-sample frames (test surface of the warp-level transform)
template <bool INVERSE>
__global__ void __launch_bounds__(256) k_rfft2048(const float2* __restrict__ tw_g, const float2* __restrict__ vtab_g,
                                                   long long n, const float* __restrict__ in,
                                                   float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);
    float2* s_vtab = s_tw + 1024;
    float* s_scratch = reinterpret_cast<float*>(s_vtab + 1024);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) {
        s_tw[i] = tw_g[i];
        s_vtab[i] = vtab_g[i];
    }
    __syncthreads();
    float* scratch = s_scratch + warp * kScratchFloats;
    for (long long f = (long long)blockIdx.x * 8 + warp; f < n; f += (long long)gridDim.x * 8) {
        float2 a[32];
        if constexpr (!INVERSE) {
            const float2* src = reinterpret_cast<const float2*>(in + f * kNfft);
#pragma unroll
            for (int r = 0; r < 32; ++r) a[r] = src[lane + 32 * r];
            float nyq;
            frame_fwd<32>(a, nyq, scratch, s_tw, s_vtab, lane);
            float2* dst = reinterpret_cast<float2*>(out + f * (2 * kBins));
#pragma unroll
            for (int r = 0; r < 32; ++r) dst[32 * r + lane] = make_float2(0.5f * a[r].x, 0.5f * a[r].y);
            if (lane == 0) dst[1024] = make_float2(0.5f * nyq, 0.0f);
        } else {
            const float2* src = reinterpret_cast<const float2*>(in + f * (2 * kBins));
#pragma unroll
            for (int r = 0; r < 32; ++r) a[r] = src[32 * r + lane];
            const float ynyq = src[1024].x;
            frame_inv<false>(a, ynyq, scratch, s_tw, s_vtab, lane);
            float2* dst = reinterpret_cast<float2*>(out + f * kNfft);
            const float sc = 1.0f / 2048.0f;
#pragma unroll
            for (int r = 0; r < 32; ++r) dst[lane + 32 * r] = make_float2(a[r].x * sc, a[r].y * sc);
        }
    }
}

// Initial phase from the reference's uniform draw (vocoder.py:103: angles = angle(exp(2j * pi * rand(F, T))), cast to
// the spectrogram's dtype).  The host shim draws u with numpy's global RNG -- the part that must be numpy's stream --
// and uploads the float64 uniforms; angle(exp(i theta)) with theta = 2 pi u in [0, 2 pi) is theta for theta <= pi and
// theta - 2 pi otherwise, evaluated here in float64 with a two-term 2 pi (the reference's libm sin / cos / atan2 round
// trip agrees to < 3e-16, i.e. the float32 results are identical up to ~1e-9 of the elements), then transposed from
// the reference's [B, F, T] to the frame-major [B * T, F] the synthesis kernels read.  HBM-bound byte mover: 8 B in +
// 4 B out per element, both sides coalesced through a 32 x 33 shared-memory tile.
__global__ void __launch_bounds__(256) k_phase_from_uniform(const double* __restrict__ u, int n_bins, int n_frames,
                                                             float* __restrict__ phase) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const double* ub = u + (size_t)b * n_bins * n_frames;
    float* pb = phase + (size_t)b * n_frames * n_bins;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int f = f0 + ty + k, t = t0 + tx;
        if (f < n_bins && t < n_frames) {
            const double th = 6.283185307179586 * ub[(size_t)f * n_frames + t];  // fl(2 pi) * u, one rounding like numpy
            const double a = th > 3.141592653589793 ? (th - 6.283185307179586) - 2.4492935982947064e-16 : th;
            tile[ty + k][tx] = (float)a;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int t = t0 + ty + k, f = f0 + tx;
        if (f < n_bins && t < n_frames) pb[(size_t)t * n_bins + f] = tile[tx][ty + k];
    }
}

#include "gl_frames.cuh"

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// The frame-parallel path keeps two raw synthesis frames per frame in the workspace (9.7 KB): only calls of up to this
// many frames carry that space (and can take the path).
#ifndef S2ST_FRAMES_PATH_MAX
#define S2ST_FRAMES_PATH_MAX (16 * 148 * 4)
#endif
constexpr long long kFramesPathMax = S2ST_FRAMES_PATH_MAX;

struct GlWorkspace {
    int4* frames;    // frame-parallel path: (T, f, utterance) per frame, or NULL
    float* y[2];     //                      raw synthesis frames, with guard rows
    int* fdone;      //                      last published iteration + 1 per frame
    UttDesc* utts;
    TileDesc* tiles;
    TileDesc* tiles_tmp;
    int* tile_pos;   // listing index -> position in the sorted table
    int* n_tiles;
    float* mag;
    float* buf[2];
    size_t total;
    long long max_tiles;
    long long wave_samples;
    int mag_stride;
};

GlWorkspace carve(const s2st_plan* plan, int n_utts, long long total_frames, void* base) {
    GlWorkspace w;
    char* ptr = static_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* r = ptr ? ptr + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    w.max_tiles = total_frames / kMinStrip + n_utts;
    w.wave_samples = (total_frames - n_utts) * (long long)plan->hop;
    w.mag_stride = (int)align_up((size_t)plan->kb, 32);  // zero padded: row loads need no bounds test
    w.utts = reinterpret_cast<UttDesc*>(take(sizeof(UttDesc) * (size_t)n_utts));
    w.tiles = reinterpret_cast<TileDesc*>(take(sizeof(TileDesc) * (size_t)w.max_tiles));
    w.tiles_tmp = reinterpret_cast<TileDesc*>(take(sizeof(TileDesc) * (size_t)w.max_tiles));
    w.tile_pos = reinterpret_cast<int*>(take(sizeof(int) * (size_t)w.max_tiles));
    w.n_tiles = reinterpret_cast<int*>(take(sizeof(int)));
    w.mag = reinterpret_cast<float*>(take(sizeof(float) * (size_t)total_frames * w.mag_stride));
    for (int i = 0; i < 2; ++i)
        w.buf[i] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)(w.wave_samples > 0 ? w.wave_samples : 1)));
    w.frames = nullptr;
    w.y[0] = w.y[1] = nullptr;
    w.fdone = nullptr;
    if (total_frames <= kFramesPathMax) {
        const size_t rows = (size_t)total_frames + 2 * kFrGuard * (size_t)n_utts + 2;
        w.frames = reinterpret_cast<int4*>(take(sizeof(int4) * (size_t)total_frames));
        w.fdone = reinterpret_cast<int*>(take(sizeof(int) * (size_t)total_frames));
        for (int i = 0; i < 2; ++i) w.y[i] = reinterpret_cast<float*>(take(sizeof(float) * rows * kFrPitch));
    }
    w.total = off;
    return w;
}

size_t gl_pass_smem(const s2st_plan* plan) {
    return sizeof(float2) * 2048 +
           sizeof(float) * (plan->wp + ((plan->hop + 3) & ~3) + kGlWarps * (kScratchFloats + ((plan->ws + 3) & ~3)));
}

// Strip length: all strips cost the same, warps take them round-robin, so the pass lasts
// ceil(n_strips / n_warps) strip times.  Pick the S that minimises that (exactly when the host knows the
// utterance lengths, from the average otherwise).
int choose_strip(const s2st_plan* plan, int n_utts, long long total_frames, const int32_t* fo_host) {
    const long long n_warps = (long long)plan->num_sms * kGlWarps;
    const int s_min = plan->nphase > kMinStrip ? plan->nphase : kMinStrip;
    if (!fo_host) {  // lengths unknown on the host: from the average
        int best = s_min;
        double best_cost = 1e300;
        for (int S = s_min; S <= kMaxStrip; ++S) {
            const long long strips = total_frames / S + (n_utts + 1) / 2;
            const double cost = (double)((strips + n_warps - 1) / n_warps) * (S + 0.35);  // + the flush / seam work of a strip
            if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && S > best)) best_cost = cost, best = S;
        }
        return best;
    }
    // Exact: a strip costs its frames + 0.35 (flush / seams); the strips (full ones of S frames + one tail per utterance)
    // are sorted by descending length like k_build_tiles does and dealt in the kernel's snake order; the longest warp
    // decides.  Candidates are visited by increasing lower bound (average load, one full strip) and the search stops
    // when the bound reaches the best cost found, so only a few S are simulated (the host time of a call matters: the
    // first version simulated all 61 and doubled it).
    struct Cand { int S; long long strips; double bound; };
    Cand cand[kMaxStrip + 1];
    int n_cand = 0, t_max = 0;
    for (int u = 0; u < n_utts; ++u) t_max = std::max(t_max, (int)(fo_host[u + 1] - fo_host[u]));
    for (int S = s_min; S <= kMaxStrip; ++S) {
        long long strips = 0;
        for (int u = 0; u < n_utts; ++u) strips += (fo_host[u + 1] - fo_host[u] + S - 1) / S;
        const long long slots = std::min<long long>(n_warps, std::max<long long>(strips, 1));
        const double avg = ((double)total_frames + 0.35 * (double)strips) / (double)slots;
        cand[n_cand++] = {S, strips, std::max(avg, (double)std::min(S, t_max) + 0.35)};
    }
    std::sort(cand, cand + n_cand, [](const Cand& a, const Cand& b) { return a.bound < b.bound || (a.bound == b.bound && a.S > b.S); });
    int best = cand[0].S;
    double best_cost = 1e300;
    std::vector<float> load;
    for (int c = 0; c < n_cand && cand[c].bound < best_cost - 1e-9; ++c) {
        const int S = cand[c].S;
        const long long strips = cand[c].strips;
        const long long slots = std::min<long long>(n_warps, std::max<long long>(strips, 1));
        double cost;
        if (strips > 8 * slots) {
            cost = (double)((strips + n_warps - 1) / n_warps) * (S + 0.35);  // many rounds: the average decides
        } else {
            long long hist[kMaxStrip + 1] = {};
            for (int u = 0; u < n_utts; ++u) {
                const int T = fo_host[u + 1] - fo_host[u];
                if (T <= 0) continue;
                const int nt = (T + S - 1) / S;
                hist[S] += nt - 1;
                hist[T - (nt - 1) * S] += 1;
            }
            load.assign((size_t)slots, 0.0f);
            long long k = 0;
            for (int len = kMaxStrip; len >= 1; --len)
                for (long long n = 0; n < hist[len]; ++n, ++k) {
                    const long long round = k / slots, pos = k % slots;
                    load[(size_t)((round & 1) ? slots - 1 - pos : pos)] += (float)len + 0.35f;
                }
            cost = *std::max_element(load.begin(), load.end());
        }
        if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && S > best)) best_cost = cost, best = S;
    }
    return best;
}

// cudaFuncSetAttribute is a driver call per launch otherwise (65 per synthesis step): do it once per kernel, device
// and size.  Not thread-safe by design (the worst case is a redundant call).
template <auto Kernel>
int allow_dynamic_smem(size_t smem, int device) {
    static size_t granted[64] = {};  // one table per kernel (the kernel is a template argument)
    if (device < 0 || device >= 64 || granted[device] < smem) {
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (device >= 0 && device < 64) granted[device] = smem;
    }
    return S2ST_OK;
}

// Frame-parallel path: one cooperative launch for the whole call (gl_frames.cuh).  Returns S2ST_OK with *launched = false
// when the launch is refused (not co-resident): the caller then runs the strip kernels.
template <int WARPS>
int launch_frames_t(const FrameGlParams& fp, int device, int grid, cudaStream_t stream, bool* launched) {
    const size_t smem = sizeof(float2) * 2048 + sizeof(float) * (64 * 19 + kStdWs + kStdHop + 2 * (kStdWs - kStdHop) + WARPS * kScratchFloats);
    if (int rc = allow_dynamic_smem<k_gl_frames<WARPS>>(smem, device)) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(32 * WARPS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // every warp must be resident: frames wait for their neighbours
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    *launched = cudaLaunchKernelEx(&cfg, k_gl_frames<WARPS>, fp) == cudaSuccess;
    if (!*launched) (void)cudaGetLastError();
    return S2ST_OK;
}

template <int NZ, bool FIRST, bool PRUNED, bool STD = false>
int launch_pass_t(const GlParams& p, int grid, size_t smem, cudaStream_t stream) {
    if (int rc = allow_dynamic_smem<k_gl_pass<NZ, FIRST, PRUNED, STD>>(smem, p.device)) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGlThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    S2ST_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gl_pass<NZ, FIRST, PRUNED, STD>, p));
    return S2ST_OK;
}

template <int NZ>
int launch_pass(const GlParams& p, bool first, bool pruned, int grid, size_t smem, cudaStream_t stream) {
    if (first) return pruned ? launch_pass_t<NZ, true, true>(p, grid, smem, stream) : launch_pass_t<NZ, true, false>(p, grid, smem, stream);
    return pruned ? launch_pass_t<NZ, false, true>(p, grid, smem, stream) : launch_pass_t<NZ, false, false>(p, grid, smem, stream);
}

}  // namespace

size_t gl_workspace_bytes(const s2st_plan* plan, int n_utts, long long total_frames) {
    return carve(plan, n_utts, total_frames, nullptr).total;
}

int launch_inverse_mel(const s2st_plan* plan, long long n_frames, const float* logmel, bool is_log, float* mag,
                       int out_stride, int n_out, cudaStream_t stream) {
    if (!plan->inv_mel_t) {
        set_error("plan was created without an inverse-mel basis");
        return S2ST_EINVAL;
    }
    if (n_frames <= 0) return S2ST_OK;
    // dense contraction -> tensor cores (tcgen05, 3xTF32) whenever the shape allows; S2ST_OPT_INVERSE_MEL = 1
    // selects the FP32 SIMT kernel (kept for other shapes and for A/B checks)
    const bool force_simt = plan->opt_inverse_mel_simt != 0;
    if (!force_simt && inverse_mel_tc_supported(plan) && ((uintptr_t)logmel & 15) == 0)
        return launch_inverse_mel_tc(plan, n_frames, logmel, is_log, mag, out_stride, n_out, stream);
    const long long blocks = (n_frames + kImFrames - 1) / kImFrames;
    k_inverse_mel<<<(unsigned)blocks, 256, sizeof(float) * kImFrames * plan->n_mels, stream>>>(
        logmel, is_log, n_frames, plan->n_mels, plan->inv_mel_t, plan->kb, plan->kb_pad, mag, out_stride, n_out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_rfft2048(const s2st_plan* plan, long long n, const float* in, float* out, bool inverse,
                    cudaStream_t stream) {
    if (n <= 0) return S2ST_OK;
    const size_t smem = sizeof(float2) * 2048 + sizeof(float) * 8 * kScratchFloats;
    const int grid = (int)min((long long)plan->num_sms * 2, (n + 7) / 8);
    if (inverse) {
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_rfft2048<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_rfft2048<true><<<grid, 256, smem, stream>>>(plan->tw, plan->vtab, n, in, out);
    } else {
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_rfft2048<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_rfft2048<false><<<grid, 256, smem, stream>>>(plan->tw, plan->vtab, n, in, out);
    }
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_phase_from_uniform(int n_batch, int n_bins, int n_frames, const double* u, float* phase, cudaStream_t stream) {
    if (n_batch <= 0 || n_bins <= 0 || n_frames <= 0) return S2ST_OK;
    dim3 grid((unsigned)((n_frames + 31) / 32), (unsigned)((n_bins + 31) / 32), (unsigned)n_batch);
    k_phase_from_uniform<<<grid, 256, 0, stream>>>(u, n_bins, n_frames, phase);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int gl_run(s2st_plan* plan, int n_utts, long long total_frames, const int32_t* frame_offsets,
           const int32_t* frame_offsets_host, const float* logmel, const float* mag, int mag_kb, const float* phase,
           unsigned long long phase_seed, int n_iter, float* wave_out, void* workspace, size_t workspace_bytes,
           cudaStream_t stream) {
    if (n_utts <= 0 || total_frames < n_utts) {
        set_error("bad batch: n_utts=%d total_frames=%lld", n_utts, total_frames);
        return S2ST_EINVAL;
    }
    const bool timed = plan->timing_enabled && n_iter + 1 <= kMaxTimedPasses;
    plan->timing_recorded = 0;
    GlWorkspace w = carve(plan, n_utts, total_frames, workspace);
    if (workspace_bytes < w.total || !workspace) {
        set_error("workspace too small: have %zu need %zu", workspace_bytes, w.total);
        return S2ST_EWORKSPACE;
    }
    if (w.wave_samples <= 0) return S2ST_OK;  // every utterance has a single frame: nothing to write
    const int s_floor = plan->nphase > kMinStrip ? plan->nphase : kMinStrip;
    const int S = plan->strip_frames > 0 ? (plan->strip_frames < s_floor ? s_floor : plan->strip_frames)
                                         : choose_strip(plan, n_utts, total_frames, frame_offsets_host);

    GlParams p;
    p.device = plan->device;
    p.pdl = plan->opt_pdl;
    p.hop = plan->hop;
    p.half = plan->n_fft / 2;
    p.rot = plan->rot;
    p.ws = plan->ws;
    p.wp = plan->wp;
    p.nphase = plan->nphase;
    p.win_a = plan->win_a;
    p.inv_nfft = 1.0f / (float)plan->n_fft;
    p.w2 = plan->w2;
    p.inv_wss = plan->inv_wss;
    p.tw = plan->tw;
    p.vtab = plan->vtab;
    p.utts = w.utts;
    p.tiles = w.tiles;
    p.n_tiles = w.n_tiles;
    p.phase = phase;
    p.phase_seed = phase_seed;
    p.phase_stride = kBins;
    if (logmel) {
        int rc = launch_inverse_mel(plan, total_frames, logmel, true, w.mag, w.mag_stride, w.mag_stride, stream);
        if (rc != S2ST_OK) return rc;
        p.mag = w.mag;
        p.mag_stride = w.mag_stride;
        p.kb = plan->kb;
    } else {
        p.mag = mag;
        p.mag_stride = kBins;
        p.kb = mag_kb;
    }
    const size_t smem = gl_pass_smem(plan);
    const long long strips_ub = total_frames / S + n_utts;  // <= w.max_tiles because S >= kMinStrip
    const int grid = (int)min((long long)plan->num_sms, strips_ub);
    const bool pruned = p.kb <= 32 * kPrunedRows;
    // three rotating waveform buffers; the one the last pass writes is the caller's output
    float* ring[3];
    ring[n_iter % 3] = wave_out;
    ring[(n_iter + 1) % 3] = w.buf[0];
    ring[(n_iter + 2) % 3] = w.buf[1];
    const bool std_geom0 = plan->nz == 19 && plan->hop == kStdHop && plan->ws == kStdWs && plan->rot == kStdRot &&
                           plan->n_fft == kNfft && pruned && p.mag_stride >= 32 * kPrunedRows;
    // Calls that fit (a few frames per resident warp): the frame-parallel kernel, all iterations in one launch.
    // A pinned strip length (s2st_plan_set_strip_frames) asks for the strip decomposition and keeps it.
    if (std_geom0 && plan->opt_frames != 0 && plan->strip_frames == 0 && w.frames && n_iter >= 0 &&
        total_frames <= (long long)plan->opt_frames_max) {
        const int blocks = (int)min((long long)148 * 4, (total_frames + 255) / 256);
        k_build_frames<<<blocks, 256, 0, stream>>>(frame_offsets, n_utts, (int)total_frames, w.frames, w.y[0], w.y[1], w.fdone);
        S2ST_CUDA_CHECK(cudaGetLastError());
        FrameGlParams fp;
        fp.win_a = plan->win_a;
        fp.w2 = plan->w2;
        fp.inv_wss = plan->inv_wss;
        fp.inv_nfft = p.inv_nfft;
        fp.tw = plan->tw;
        fp.vtab = plan->vtab;
        fp.rot = plan->rot;
        fp.frames = w.frames;
        fp.n_frames = (int)total_frames;
        fp.n_iter = n_iter;
        fp.mag = p.mag;
        fp.mag_stride = p.mag_stride;
        fp.kb = p.kb;
        fp.phase = phase;
        fp.phase_stride = kBins;
        fp.phase_seed = phase_seed;
        fp.Y[0] = w.y[0];
        fp.Y[1] = w.y[1];
        fp.done = w.fdone;
        fp.out = wave_out;
        const int fgrid = (int)min((long long)plan->num_sms, total_frames);
        const long long per_cta = (total_frames + fgrid - 1) / fgrid;
        bool launched = false;
        if (timed) {
            if (!plan->timing_events[0]) S2ST_CUDA_CHECK(cudaEventCreate(&plan->timing_events[0]));
            S2ST_CUDA_CHECK(cudaEventRecord(plan->timing_events[0], stream));
        }
        int rc = per_cta <= 1 ? launch_frames_t<1>(fp, plan->device, fgrid, stream, &launched)
               : per_cta <= 2 ? launch_frames_t<2>(fp, plan->device, fgrid, stream, &launched)
               : per_cta <= 4 ? launch_frames_t<4>(fp, plan->device, fgrid, stream, &launched)
               : per_cta <= 8 ? launch_frames_t<8>(fp, plan->device, fgrid, stream, &launched)
                              : launch_frames_t<16>(fp, plan->device, fgrid, stream, &launched);
        if (rc != S2ST_OK) return rc;
        if (launched) {
            plan->last_launches = (logmel ? 1 : 0) + 2;  // [inverse_mel], build_frames, the frame kernel
            if (timed) {
                if (!plan->timing_events[1]) S2ST_CUDA_CHECK(cudaEventCreate(&plan->timing_events[1]));
                S2ST_CUDA_CHECK(cudaEventRecord(plan->timing_events[1], stream));
                plan->timing_recorded = 2;
            }
            return S2ST_OK;
        }
    }
    k_build_tiles<<<1, 1024, 0, stream>>>(frame_offsets, n_utts, plan->hop, S, w.utts, w.tiles_tmp, w.tiles, w.tile_pos, w.n_tiles);
    S2ST_CUDA_CHECK(cudaGetLastError());
    // pass 0 accumulates its seams into ring[0]: clear it (later passes get theirs cleared by the pass before)
    S2ST_CUDA_CHECK(cudaMemsetAsync(ring[0], 0, sizeof(float) * (size_t)w.wave_samples, stream));
    plan->last_launches = 1 + (logmel ? 1 : 0) + (n_iter + 1);  // build_tiles, [inverse_mel], the passes
    for (int it = 0; it <= n_iter; ++it) {
        p.in = ring[(it + 2) % 3];
        p.out = ring[it % 3];
        p.zero_next = ring[(it + 1) % 3];
        if (timed) {
            if (!plan->timing_events[it]) S2ST_CUDA_CHECK(cudaEventCreate(&plan->timing_events[it]));
            S2ST_CUDA_CHECK(cudaEventRecord(plan->timing_events[it], stream));
        }
        // the iteration kernel specialised for the vocoder's standard geometry, else the generic one
        const bool std_geom = plan->nz == 19 && plan->hop == kStdHop && plan->ws == kStdWs && plan->rot == kStdRot &&
                              plan->n_fft == kNfft && pruned && p.mag_stride >= 32 * kPrunedRows;
        const int rc = (it > 0 && std_geom) ? launch_pass_t<19, false, true, true>(p, grid, smem, stream)
                       : std_geom             ? launch_pass_t<19, true, true, true>(p, grid, smem, stream)
                       : (plan->nz == 19)     ? launch_pass<19>(p, it == 0, pruned, grid, smem, stream)
                                              : launch_pass<32>(p, it == 0, pruned, grid, smem, stream);
        if (rc != S2ST_OK) return rc;
    }
    if (timed) {
        if (!plan->timing_events[n_iter + 1]) S2ST_CUDA_CHECK(cudaEventCreate(&plan->timing_events[n_iter + 1]));
        S2ST_CUDA_CHECK(cudaEventRecord(plan->timing_events[n_iter + 1], stream));
        plan->timing_recorded = n_iter + 2;
    }
    return S2ST_OK;
}

}  // namespace s2st
