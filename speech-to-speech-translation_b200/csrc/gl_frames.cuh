// Frame-parallel Griffin-Lim for calls that fit the resident warps (included by gl_kernels.cu, inside its namespace).
//
// The reference's own call shape is ONE utterance per GriffinLimVocoder.forward (speech_generator_for_s2st.py:115-124):
// a few hundred frames, 64 dependent iterations.  The strip kernel (k_gl_pass) is built for throughput -- a warp walks
// S frames in sequence and a pass is a kernel launch -- so a small call costs ~26 us per iteration whatever its size
// (profiles/r02_small_calls.txt).  Here ALL iterations run in ONE cooperative launch and every frame has a warp of its
// own for the whole call:
//   * iteration `it` of frame f writes its raw synthesis frame (19 rows x 64 samples, before windowing) to row f of
//     Y[it & 1], publishes done[f] = it + 1 (release), and starts iteration it + 1 as soon as the frames f-4 .. f+4
//     have published iteration it (acquire) -- no grid-wide barrier, no kernel boundary;
//   * the overlap-add moves to the consumer: a sample of the waveform is
//         x[j] = (sum over the <= 4 frames t covering j, ascending t, of fma(y_t[j - t hop], w[j - t hop], .)) * inv[j]
//     gathered straight from the neighbours' Y rows (L2 resident: 4.9 KB per frame and buffer), windowed and
//     transformed.  The additions are the same, in the same order, as the strip kernel's shared-memory ring performs
//     for a strip that covers the whole utterance, and the edge normalisation follows emit_generic: the arithmetic
//     of k_gl_pass with one strip per utterance.  Every stage of every frame is bitwise equal to it in an instrumented
//     build; in the release build the two kernels differ in the last bit, because ptxas (12.9) fuses mul.rn.f32x2 +
//     add.rn.f32x2 pairs into FFMA2 as its scheduling sees fit (even with --fmad=false), i.e. two separately compiled
//     kernels round the same packed source differently (tests/test_gl_gpu.py bounds the difference).
//     Within this kernel results are deterministic and, unlike the automatic strip length, bitwise independent of
//     what else is in the batch.
//   * utterances get three zero guard rows on both sides in Y, so "frame does not exist" needs no test in the gather;
//     the reflect padding of the first / last two frames mirrors the frame's OWN samples through the warp scratch.
// Calls with more frames than resident warps give every warp several frames per iteration (same dependencies, no
// deadlock: all warps are co-resident and a warp never waits for a later iteration); the gather reads 4x the waveform
// from L2, so large batches stay on the strip kernel (gl_run: kFramesPathMax, S2ST_OPT_GL_FRAMES).
#pragma once

constexpr int kFrPitch = 64 * 19;     // floats per Y row (19 rows of 32 (odd, even) sample pairs)
constexpr int kFrGuard = 3;           // zero rows before / after each utterance in Y
constexpr int kFrReach = 4;           // frames f-4 .. f+4 are read (3 by the overlap, 1 more through the reflect padding)

struct FrameGlParams {
    // plan constants
    const float* win_a;      // [1216]
    const float* w2;         // [1200]
    const float* inv_wss;    // [300]
    float inv_nfft;
    const float2* tw;
    const float2* vtab;
    int rot;
    // batch
    const int4* frames;      // per global frame g: (T, f, utterance, 0)
    int n_frames;
    int n_iter;
    const float* mag;        // [n_frames, mag_stride]
    int mag_stride;
    int kb;
    const float* phase;      // [n_frames, phase_stride] or NULL -> device RNG
    int phase_stride;
    unsigned long long phase_seed;
    float* Y[2];             // [n_frames + 2 * kFrGuard * n_utts + 2 rows][kFrPitch] (the gather of a last frame's zero-weight tail reaches one row past its guards)
    int* done;               // [n_frames], zeroed
    float* out;              // concatenated waveforms, utterance u at (fo[u] - u) * hop
};

// (T, f, u) per frame, zero guard rows, zero flags.  One block per 256 frames; grid-stride for the clearing.
__global__ void __launch_bounds__(256) k_build_frames(const int32_t* __restrict__ fo, int n_utts, int n_frames,
                                                       int4* __restrict__ frames, float* __restrict__ y0,
                                                       float* __restrict__ y1, int* __restrict__ done) {
    const int stride = gridDim.x * blockDim.x, tid0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int g = tid0; g < n_frames; g += stride) {
        int lo = 0, hi = n_utts - 1;  // last u with fo[u] <= g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (fo[mid] <= g) lo = mid; else hi = mid - 1;
        }
        frames[g] = make_int4(fo[lo + 1] - fo[lo], g - fo[lo], lo, 0);
        done[g] = 0;
    }
    // guard rows: utterance u owns Y rows [fo[u] + 6u, fo[u+1] + 6(u+1)): 3 guards, T frames, 3 guards
    const int per_utt = 2 * kFrGuard * (kFrPitch / 4);  // float4 per utterance and buffer
    for (long long e = tid0; e < (long long)n_utts * per_utt; e += stride) {
        const int u = (int)(e / per_utt), q = (int)(e - (long long)u * per_utt);
        const int gr = q / (kFrPitch / 4), c = q - gr * (kFrPitch / 4);
        const long long row = gr < kFrGuard ? (long long)fo[u] + 2 * kFrGuard * u + gr
                                            : (long long)fo[u + 1] + 2 * kFrGuard * u + gr;  // = fo[u+1] + 6u + 3 + (gr - 3)
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        reinterpret_cast<float4*>(y0 + row * kFrPitch)[c] = z;
        reinterpret_cast<float4*>(y1 + row * kFrPitch)[c] = z;
    }
}

// One waveform sample at position j in [0, L) for utterances of fewer than 4 frames, where both edges interact: the
// clipped window sum itself (the arithmetic of emit_generic).  Cold.
__device__ __noinline__ float frames_sample_slow(const float* __restrict__ Yin, long long yrow0, int T, int j,
                                                  const float* __restrict__ s_win, const float* __restrict__ s_w2,
                                                  const float* __restrict__ s_inv_wss, float inv_nfft) {
    constexpr int HOP = kStdHop, WS = kStdWs;
    const int i = j + WS / 2, th = i / HOP, o = i - th * HOP;
    const float* src = Yin + (yrow0 + th - 3) * kFrPitch;
    float acc = 0.0f, wsum = 0.0f;
    for (int dt = 0; dt < 4; ++dt) {
        const int q = o + (3 - dt) * HOP, t = th - 3 + dt;
        acc = fmaf(__ldcg(src + dt * kFrPitch + (q ^ 1)), s_win[q], acc);
        if (t >= 0 && t < T) wsum += s_w2[q];
    }
    const float inv = (i < WS - HOP || i >= T * HOP) ? (wsum > kTiny ? 1.0f / wsum : 1.0f) * inv_nfft : s_inv_wss[o];
    return acc * inv;
}

// -DS2ST_FRAMES_PROF: the warp of the middle frame prints the SM-clock time stamps of its phases (tools/dbg)
#ifdef S2ST_FRAMES_PROF
#define FR_STAMP(i) do { if (prof_on) stamp[i] = clock64(); } while (0)
#else
#define FR_STAMP(i) do { } while (0)
#endif

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) k_gl_frames(const __grid_constant__ FrameGlParams p) {
    constexpr int NZ = 19, HOP = kStdHop, WS = kStdWs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);   // 1024
    float2* s_vtab = s_tw + 1024;                          // 1024
    float* s_win = reinterpret_cast<float*>(s_vtab + 1024);  // 1216
    float* s_w2 = s_win + 64 * NZ;                         // 1200
    // normalisation 1 / (n_fft * window sum), one table: [0, 900) the first 900 samples frame 0 covers, [900, 1200) the
    // steady state by (sample mod hop), [1200, 2100) the samples from T * hop on
    float* s_inv_head = s_w2 + WS;
    float* s_inv_wss = s_inv_head + (WS - HOP);
    float* s_inv_tail = s_inv_wss + HOP;
    float* s_scratch = s_inv_tail + (WS - HOP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += 32 * WARPS) {
        s_tw[i] = p.tw[i];
        s_vtab[i] = p.vtab[i];
    }
    for (int i = tid; i < 64 * NZ; i += 32 * WARPS) s_win[i] = p.win_a[i];
    for (int i = tid; i < WS; i += 32 * WARPS) s_w2[i] = p.w2[i];
    for (int i = tid; i < HOP; i += 32 * WARPS) s_inv_wss[i] = p.inv_wss[i];
    __syncthreads();
    // Utterance edges: only the frames that exist contribute to the window sum (vocoder.py:78-81; emit_generic of the
    // strip kernel).  With T >= 4 the clipped sums depend only on the distance to the edge: tabulate them once per CTA,
    // same additions in the same (frame) order, same division.
    for (int e = tid; e < WS - HOP; e += 32 * WARPS) {
        const int hh = e / HOP, o = e - hh * HOP;
        float wh = 0.0f, wt = 0.0f;
        for (int dt = 0; dt < 4; ++dt) {
            const int q = o + (3 - dt) * HOP;
            if (dt >= 3 - hh) wh += s_w2[q];   // head: frames 0 .. hh
            if (dt <= 2 - hh) wt += s_w2[q];   // tail: frames T + hh - 3 .. T - 1
        }
        s_inv_head[e] = (wh > kTiny ? 1.0f / wh : 1.0f) * p.inv_nfft;
        s_inv_tail[e] = (wt > kTiny ? 1.0f / wt : 1.0f) * p.inv_nfft;
    }
    __syncthreads();
    float* scratch = s_scratch + warp * kScratchFloats;
    const int kb = min(p.kb, 32 * kPrunedRows);
    // Frames are dealt round-robin over the warp slots (SM first).  When the last round would be nearly empty (< 5 % of
    // the slots: 4 800 frames = 2 x 2 368 + 64) the slots are reduced so that every warp gets the same number of frames
    // -- 1 600 warps x 3, ~11 per SM, instead of two full rounds plus a third that every frame of the next iteration
    // waits for: 1.86 -> 1.75 ms for the 60 s utterance.  With a fuller last round all 16 warps per SM win (measured).
    const int max_slots = gridDim.x * WARPS;
    const int per_slot = (p.n_frames + max_slots - 1) / max_slots;
    const bool balance = per_slot > 1 && (p.n_frames - (per_slot - 1) * max_slots) * 20 < max_slots;
    const int slots = balance ? (p.n_frames + per_slot - 1) / per_slot : max_slots;
    const int my_slot = blockIdx.x + gridDim.x * warp;

#pragma unroll 1
    for (int it = 0; it <= p.n_iter + 1; ++it) {  // 0: initial inverse; 1..n_iter: iterations; n_iter + 1: write-out
        const float* Yin = p.Y[(it + 1) & 1];
        float* Yout = p.Y[it & 1];
#pragma unroll 1
        for (int g = my_slot < slots ? my_slot : p.n_frames; g < p.n_frames; g += slots) {
            const int4 fd = __ldg(p.frames + g);
            const int T = fd.x, f = fd.y, u = fd.z;
            if (T < 2) continue;  // a single frame: the utterance has no samples
#ifdef S2ST_FRAMES_PROF
            const bool prof_on = lane == 0 && (g == p.n_frames / 2 || g == 1) && it >= 10 && it <= 11;
            long long stamp[8];
#endif
            FR_STAMP(0);
            const int L = (T - 1) * HOP;
            const long long yrow0 = (long long)(g - f) + 2 * kFrGuard * u + kFrGuard;  // Y row of the utterance's frame 0
            if (it > 0) {
                // the frames whose rows are read now -- and which read the row written below one iteration ago
                const int t = f - kFrReach + lane;
                if (lane <= 2 * kFrReach && t >= 0 && t < T) {
                    // poll relaxed, acquire once: every ld.acquire.gpu invalidates the L1 (CCTL.IVALL was 19 % of the
                    // kernel's stall samples when each poll was an acquire)
                    const int* flag = p.done + (g - f + t);
                    while (ld_relaxed(flag) < it) __nanosleep(20);
                    (void)ld_acquire(flag);
                }
                __syncwarp();
            }
            FR_STAMP(1);
            // one waveform sample of the previous iteration at (already reflected) position j in [0, L); T >= 4.
            // Branch-free, so that the loads of many samples can be in flight together.
            auto sample_at = [&](int j) -> float {
                const int i = j + (WS / 2);          // relative to the first sample frame 0 covers (j = -600)
                const int th = i / HOP, o = i - th * HOP;
                const float* src = Yin + (yrow0 + th - 3) * kFrPitch;
                float y[4];
#pragma unroll
                for (int dt = 0; dt < 4; ++dt)  // index q = o + (3 - dt) hop inside frame t = th - 3 + dt; rows hold (odd, even) pairs
                    y[dt] = __ldcg(src + dt * kFrPitch + ((o + (3 - dt) * HOP) ^ 1));
                const int e = i - T * HOP;
                const float inv = s_inv_head[i < WS - HOP ? i : e >= 0 ? e + WS : o + (WS - HOP)];
                float acc = 0.0f;
#pragma unroll
                for (int dt = 0; dt < 4; ++dt) acc = fmaf(y[dt], s_win[o + (3 - dt) * HOP], acc);
                return acc * inv;
            };
            if (it == p.n_iter + 1) {
                if (f < T - 1) {
                    float* dst = p.out + (long long)(g - f - u) * HOP + f * HOP;
                    float v[10];
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const int e = lane + 32 * k;
                        v[k] = e >= HOP ? 0.0f : T >= 4 ? sample_at(f * HOP + e)
                                                        : frames_sample_slow(Yin, yrow0, T, f * HOP + e, s_win, s_w2, s_inv_wss, p.inv_nfft);
                    }
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const int e = lane + 32 * k;
                        if (e < HOP) dst[e] = v[k];
                    }
                }
                continue;
            }

            float2 a[32];
            const float* magrow = p.mag + (size_t)g * p.mag_stride;
            if (it == 0) {
                // spectrum from (magnitude, initial phase): the code of k_gl_pass<FIRST>
                const float* phrow = p.phase ? p.phase + (size_t)g * p.phase_stride : nullptr;
                float mgv[kPrunedRows], phv[kPrunedRows];
#pragma unroll
                for (int r = 0; r < kPrunedRows; ++r) {
                    const int k = 32 * r + lane;
                    mgv[r] = k < kb ? __ldg(magrow + k) : 0.0f;
                    if (p.phase) {
                        phv[r] = k < kb ? __ldg(phrow + k) : 0.0f;
                    } else {
                        const unsigned long long e = (unsigned long long)g * kBins + k;
                        phv[r] = 2.0f * uniform01(p.phase_seed, e) - 1.0f;
                    }
                }
#pragma unroll
                for (int r = 0; r < kPrunedRows; ++r) {
                    const int k = 32 * r + lane;
                    const float rot_pi = (float)(((k * p.rot + 1024) & 2047) - 1024) * (1.0f / 1024.0f);
                    const float v = p.phase ? fmaf(phv[r], 0.31830988618379067154f, rot_pi) : phv[r] + rot_pi;
                    float sn, cs;
                    sincospif(v, &sn, &cs);
                    a[r] = make_float2(mgv[r] * cs, mgv[r] * sn);
                }
            } else {
                if (T < 4) {
                    // (reachable through the C ABI only: the reference's reflect padding needs T >= 5)
#pragma unroll 1
                    for (int r = 0; r < NZ; ++r) {
                        const int j = f * HOP - WS / 2 + 64 * r + 2 * lane;
                        int j0 = j < 0 ? -j : j, j1 = j + 1 < 0 ? -(j + 1) : j + 1;
                        j0 = j0 >= L ? 2 * (L - 1) - j0 : j0;
                        j1 = j1 >= L ? 2 * (L - 1) - j1 : j1;
                        j0 = min(max(j0, 0), L - 1);
                        j1 = min(max(j1, 0), L - 1);
                        const float2 x = make_float2(frames_sample_slow(Yin, yrow0, T, j0, s_win, s_w2, s_inv_wss, p.inv_nfft),
                                                     frames_sample_slow(Yin, yrow0, T, j1, s_win, s_w2, s_inv_wss, p.inv_nfft));
                        const float2 xw = mul2(x, *reinterpret_cast<const float2*>(s_win + 64 * r + 2 * lane));
#pragma unroll
                        for (int q = 0; q < NZ; ++q)
                            if (q == r) a[brev5(q)] = xw;
                    }
                } else {
                    // Every sample pair from the four frames that can cover it (missing frames are zero guard rows),
                    // normalised from the one table: pairs stay pairs, 76 independent 8-byte loads per lane.
#pragma unroll
                    for (int r = 0; r < NZ; ++r) {
                        const int m = 64 * r + 2 * lane;
                        const int h = m / HOP, o = m - h * HOP;
                        const float* src = Yin + (yrow0 + f + h - 3) * kFrPitch + o + 3 * HOP;
                        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                        for (int dt = 0; dt < 4; ++dt) {
                            const float2 y = __ldcg(reinterpret_cast<const float2*>(src + dt * (kFrPitch - HOP)));
                            const float2 w = *reinterpret_cast<const float2*>(s_win + o + (3 - dt) * HOP);
                            acc = fma2(swap2(y), w, acc);
                        }
                        const int i = f * HOP + m, e = i - T * HOP;
                        // (the last frame's zero-weight samples m >= 1200 lie past the tail table: clamped, their value is multiplied by 0)
                        a[brev5(r)] = mul2(acc, *reinterpret_cast<const float2*>(s_inv_head + (i < WS - HOP ? i : e >= 0 ? min(e, WS - HOP - 2) + WS : o + (WS - HOP))));
                    }
                    // Reflect padding (audio_utils.py:262-263): the first two and the last two frames of an utterance read
                    // mirrored samples in their first / last rows (frame 0: rows 0-9, frame 1: 0-4, frame T-1: 9-18,
                    // frame T-2: 14-18).  The mirror image of such a sample is one of the frame's OWN samples -- sample m
                    // of the frame takes the value of sample c - m with c = 1200, 600, 1198, 1798 -- so the rows are
                    // exchanged through the warp scratch, no further loads.
                    const int fix_lo = f <= 1 ? 0 : f == T - 1 ? 9 : 14;
                    const int fix_hi = f == 0 ? 10 : f == 1 ? 5 : (f >= T - 2 ? NZ : 0);  // rows [fix_lo, fix_hi)
                    if (fix_hi > fix_lo) {
                        const int c = f == 0 ? 4 * HOP : f == 1 ? 2 * HOP : f == T - 1 ? 4 * HOP - 2 : 6 * HOP - 2;
                        float* xs = scratch;  // the frame's samples, linear
#pragma unroll
                        for (int r = 0; r < NZ; ++r) *reinterpret_cast<float2*>(xs + 64 * r + 2 * lane) = a[brev5(r)];
                        __syncwarp();
#pragma unroll
                        for (int r = 0; r < NZ; ++r) {
                            const int m = 64 * r + 2 * lane;
                            // which samples of the row are mirrored: left edge j < 0 <=> m < 600 - f hop; right edge j >= L
                            const int j = f * HOP - WS / 2 + m;
                            if (r >= fix_lo && r < fix_hi) {
                                const int s0 = min(max(c - m, 0), 64 * NZ - 1), s1 = min(max(c - m - 1, 0), 64 * NZ - 1);
                                const float x0 = (j < 0 || j >= L) ? xs[s0] : a[brev5(r)].x;
                                const float x1 = (j + 1 < 0 || j + 1 >= L) ? xs[s1] : a[brev5(r)].y;
                                a[brev5(r)] = make_float2(x0, x1);
                            }
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int r = 0; r < NZ; ++r)
                        a[brev5(r)] = mul2(a[brev5(r)], *reinterpret_cast<const float2*>(s_win + 64 * r + 2 * lane));
                }
                float mg[kPrunedRows];
#pragma unroll
                for (int r = 0; r < kPrunedRows; ++r) mg[r] = __ldg(magrow + 32 * r + lane);
                FR_STAMP(2);
                fwd1024<NZ, 32>(a, scratch, s_tw, lane);
                float nyq;
                fwd_split<true>(a, nyq, scratch, s_vtab, lane);
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
            }
            FR_STAMP(3);
            inv_merge<true, true, true>(a, 0.0f, scratch, s_vtab, lane);
            fwd1024<32, NZ>(a, scratch, s_tw, lane);
            float2* dst = reinterpret_cast<float2*>(Yout + (yrow0 + f) * kFrPitch) + lane;
            FR_STAMP(4);
#pragma unroll
            for (int r = 0; r < NZ; ++r) dst[32 * r] = a[r];
            FR_STAMP(5);
            // publish: the warp barrier orders every lane's stores before lane 0's release (cumulative at gpu scope) -- the
            // pattern of a cooperative-groups grid barrier (block barrier, then one thread fences and signals)
            __syncwarp();
            if (lane == 0) st_release(p.done + g, it + 1);
            FR_STAMP(6);
#ifdef S2ST_FRAMES_PROF
            if (prof_on)
                printf("it %d g %d: wait %lld gather %lld analysis %lld synthesis %lld store %lld fence %lld release %lld (cycles)\n", it, g,
                       stamp[1] - stamp[0], stamp[2] - stamp[1], stamp[3] - stamp[2], stamp[4] - stamp[3], stamp[5] - stamp[4],
                       stamp[6] - stamp[5], clock64() - stamp[6]);
#endif
        }
    }
}
