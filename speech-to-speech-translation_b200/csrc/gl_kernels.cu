// Griffin-Lim synthesis kernels (sm_100a).
//
// Replaces GriffinLim.forward / GriffinLim.inverse / TTSSpectrogram.forward of the reference
// (fairseq/models/text_to_speech/vocoder.py:84-110, fairseq/data/audio/audio_utils.py:259-271):
// the reference runs each STFT / ISTFT as a dense 2050x2048 convolution plus a host-side Python
// loop for the window-sum-square; here one persistent kernel launch per Griffin-Lim iteration does
//   frame load (reflect) -> window -> rFFT-2048 -> magnitude re-imposition -> irFFT-2048 -> window
//   -> overlap-add in shared memory -> 1/window-sum-square -> store
// for a ragged batch of utterances.  The only state between iterations is the waveform.
//
// Tiling: a tile = 8 consecutive frames of one utterance = one CTA pass (8 warps, one frame per
// warp).  A tile owns the samples only its frames touch and stores them normalised with plain
// stores.  The "seam" samples it shares with the next / previous tile receive exactly two
// contributions; both are added with red.global.add onto a zeroed location, so the result is the
// same whichever lands first (a + b == b + a) and the output stays bitwise deterministic.  Three
// waveform buffers rotate: pass i reads buf[(i-1)%3], writes buf[i%3] and zeroes the seams of
// buf[(i+1)%3] for the next pass.  Readers therefore see a finished, normalised waveform and a
// frame load is 19 independent 8-byte loads per lane.
#include <math_constants.h>

#include "../../include/s2st_b200.h"
#include "frame_fft.cuh"
#include "plan.h"

namespace s2st {

namespace {

constexpr float kTiny = 1.1754944e-38f;  // vocoder.py:69
constexpr int kMagRegs = 22;             // magnitude rows prefetched into registers (covers kb <= 704)

// two consecutive samples of the reflect-padded waveform (audio_utils.py:262-263), j even
__device__ __forceinline__ float2 load_pair(const float* __restrict__ y, int j, int L, bool vec_ok) {
    if (vec_ok && j >= 0 && j + 1 < L) return *reinterpret_cast<const float2*>(y + j);
    int j0 = j < 0 ? -j : j, j1 = j + 1 < 0 ? -(j + 1) : j + 1;
    j0 = j0 >= L ? 2 * (L - 1) - j0 : j0;
    j1 = j1 >= L ? 2 * (L - 1) - j1 : j1;
    j0 = min(max(j0, 0), L - 1);  // only reachable where the window is zero
    j1 = min(max(j1, 0), L - 1);
    return make_float2(y[j0], y[j1]);
}

// One Griffin-Lim pass.  FIRST: spectra come from (mag, initial phase) -> inverse only.
template <int NZ, bool FIRST>
__global__ void __launch_bounds__(kGlThreads, 2) k_gl_pass(const __grid_constant__ GlParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);            // 1024
    float2* s_vtab = s_tw + 1024;                                  // 1024
    float2* s_scratch = s_vtab + 1024;                             // 8 * kScratchFloat2
    float* s_win_a = reinterpret_cast<float*>(s_scratch + kTileFrames * kScratchFloat2);
    float* s_win_s = s_win_a + 64 * NZ;
    float* s_inv_wss = s_win_s + 64 * NZ;                          // hop (rounded up to 4)
    float* s_ola = s_inv_wss + ((p.hop + 3) & ~3);                 // (kTileFrames-1)*hop + ws

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += kGlThreads) {
        s_tw[i] = p.tw[i];
        s_vtab[i] = p.vtab[i];
    }
    for (int i = tid; i < 64 * NZ; i += kGlThreads) {
        s_win_a[i] = p.win_a[i];
        s_win_s[i] = p.win_s[i];
    }
    for (int i = tid; i < p.hop; i += kGlThreads) s_inv_wss[i] = p.inv_wss[i];
    float2* scratch = s_scratch + warp * kScratchFloat2;
    const int n_tiles = *p.n_tiles;
    const bool hop_even = (p.hop & 1) == 0;
    // increments of (i / hop, i % hop) when i advances by the CTA size
    const int step_q = kGlThreads / p.hop, step_r = kGlThreads % p.hop;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileDesc td = p.tiles[tile];
        const UttDesc ud = p.utts[td.utt];
        const int T = ud.n_frames, L = (T - 1) * p.hop;
        const int span_out = (td.nf - 1) * p.hop + p.ws;
        const int j_base = td.f0 * p.hop + p.rot - p.half;  // output sample of s_ola[0]
        const bool active = warp < td.nf;
        const size_t row = (size_t)ud.frame_off + td.f0 + warp;
        const float* magrow = p.mag + row * p.mag_stride;
        float2 a[32];

        // ---- issue this frame's global loads first: they fly while the overlap-add buffer is cleared
        if constexpr (!FIRST) {
            if (active) {
                const float* y = p.in + ud.wave_off;
                const int j0 = j_base + warp * p.hop + 2 * lane;
                const bool vec_ok = hop_even && ((ud.wave_off & 1) == 0);
#pragma unroll
                for (int r = 0; r < NZ; ++r) a[r] = load_pair(y, j0 + 64 * r, L, vec_ok);
            }
        }
        __syncthreads();  // tables loaded / previous tile's write-out finished
        for (int i = tid; i < span_out; i += kGlThreads) s_ola[i] = 0.0f;

        if (active) {
            float ynyq = 0.0f;
            if constexpr (!FIRST) {
                const float* w = s_win_a + 2 * lane;
#pragma unroll
                for (int r = 0; r < NZ; ++r) {
                    const float2 ww = *reinterpret_cast<const float2*>(w + 64 * r);
                    a[r] = make_float2(a[r].x * ww.x, a[r].y * ww.y);
                }
                frame_fwd_a<NZ>(a, scratch, s_tw, lane);
                // target magnitudes: prefetched here so the second in-lane FFT hides their latency
                float mg[kMagRegs];
                const bool pre = p.kb <= 32 * kMagRegs;
                if (pre) {
#pragma unroll
                    for (int r = 0; r < kMagRegs; ++r) {
                        const int k = 32 * r + lane;
                        mg[r] = (k < p.kb) ? __ldg(magrow + k) : 0.0f;
                    }
                }
                float nyq;
                frame_fwd_b(a, nyq, scratch, s_vtab, lane, p.kb);
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    if (32 * r < p.kb) {
                        const int k = 32 * r + lane;
                        float m;
                        if (r < kMagRegs && pre) m = mg[r < kMagRegs ? r : 0];
                        else m = (k < p.kb) ? __ldg(magrow + k) : 0.0f;
                        float x = a[r].x, y = a[r].y;
                        float r2 = fmaf(x, x, y * y);
                        if (r2 < 1e-30f) {  // keep the phase of tiny (possibly denormal) bins
                            x *= 1.1529215e18f;
                            y *= 1.1529215e18f;
                            r2 = fmaf(x, x, y * y);
                        }
                        const float sc = m * rsqrtf(r2);
                        // atan2(0, +-0) = 0 / pi  ->  (+-mag, 0)
                        a[r] = r2 > 0.0f ? make_float2(x * sc, y * sc) : make_float2(copysignf(m, x), 0.0f);
                    } else {
                        a[r] = make_float2(0.0f, 0.0f);
                    }
                }
                if (p.kb > 1024) ynyq = copysignf(__ldg(magrow + 1024), nyq);
            } else {
                const float* phrow = p.phase + row * p.phase_stride;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const int k = 32 * r + lane;
                    if (32 * r < p.kb && k < p.kb) {
                        const float m = __ldg(magrow + k);
                        float sn, cs, rs, rc;
                        sincosf(__ldg(phrow + k), &sn, &cs);
                        // the caller's phase refers to the un-rotated frame; frames are processed rotated
                        // by p.rot samples:  Y'[k] = Y[k] * exp(+2 pi i k rot / 2048)
                        sincospif((float)((k * p.rot) & 2047) * (1.0f / 1024.0f), &rs, &rc);
                        a[r] = make_float2(m * fmaf(cs, rc, -sn * rs), m * fmaf(cs, rs, sn * rc));
                    } else {
                        a[r] = make_float2(0.0f, 0.0f);
                    }
                }
                if (p.kb > 1024) ynyq = __ldg(magrow + 1024) * cosf(__ldg(phrow + 1024));
            }
            frame_inv(a, ynyq, scratch, s_tw, s_vtab, lane);
        }
        __syncthreads();  // zero-fill done
        // overlap-add: frames of the same phase (w mod nphase) never overlap
        for (int c = 0; c < p.nphase; ++c) {
            if (active && (warp % p.nphase) == c) {
                float* dst = s_ola + warp * p.hop + 2 * lane;
                const float* w = s_win_s + 2 * lane;
                if (hop_even) {
#pragma unroll
                    for (int r = 0; r < NZ; ++r) {
                        if (64 * r + 2 * lane < p.ws) {
                            float2 o = *reinterpret_cast<float2*>(dst + 64 * r);
                            const float2 ww = *reinterpret_cast<const float2*>(w + 64 * r);
                            o.x = fmaf(a[r].x, ww.x, o.x);
                            o.y = fmaf(a[r].y, ww.y, o.y);
                            *reinterpret_cast<float2*>(dst + 64 * r) = o;
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < NZ; ++r) {
                        if (64 * r + 2 * lane < p.ws) {
                            dst[64 * r] = fmaf(a[r].x, w[64 * r], dst[64 * r]);
                            dst[64 * r + 1] = fmaf(a[r].y, w[64 * r + 1], dst[64 * r + 1]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // ---- normalise and write.  Sample i of the tile sits at q = f0*hop + i from frame 0's origin,
        // so q mod hop == i mod hop and the last frame that can cover it is f0 + i / hop.
        {
            float* out = p.out + ud.wave_off;
            float* znext = p.zero_next + ud.wave_off;
            const bool has_prev = td.f0 > 0, has_next = td.f0 + td.nf < T;
            const int left_end = p.ws - p.hop;        // i < left_end  : also covered by the previous tile
            const int right_beg = td.nf * p.hop;      // i >= right_beg: also covered by the next tile
            int q = tid / p.hop, r = tid % p.hop;
            for (int i = tid; i < span_out; i += kGlThreads) {
                const int j = j_base + i;
                if (j >= 0 && j < L) {
                    const int t_hi_u = td.f0 + q;
                    const int t_lo_u = i < p.ws ? td.f0 - (p.ws - 1 - i) / p.hop : td.f0 + (i - p.ws) / p.hop + 1;
                    float inv;
                    if (t_lo_u >= 0 && t_hi_u <= T - 1) {
                        inv = s_inv_wss[r];
                    } else {  // utterance edges: only the frames that exist (vocoder.py:78-81 order)
                        float acc = 0.0f;
                        const int qq = td.f0 * p.hop + i;
                        for (int t = max(t_lo_u, 0); t <= min(t_hi_u, T - 1); ++t) acc += __ldg(p.w2 + (qq - t * p.hop));
                        inv = acc > kTiny ? 1.0f / acc : 1.0f;
                    }
                    const float v = s_ola[i] * inv;
                    const bool right_seam = has_next && i >= right_beg;
                    if (right_seam || (has_prev && i < left_end)) atomicAdd(out + j, v);
                    else out[j] = v;
                    if (right_seam) znext[j] = 0.0f;
                }
                q += step_q;
                r += step_r;
                if (r >= p.hop) {
                    r -= p.hop;
                    ++q;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(1024) k_build_tiles(const int32_t* __restrict__ fo, int n_utts, int hop,
                                                       UttDesc* __restrict__ utts,
                                                       TileDesc* __restrict__ tiles, int* __restrict__ n_tiles) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_utts; base += 1024) {
        const int u = base + tid;
        const int T = u < n_utts ? fo[u + 1] - fo[u] : 0;
        const int nt = (T + kTileFrames - 1) / kTileFrames;
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
                t.utt = u;
                t.f0 = k * kTileFrames;
                t.nf = min(kTileFrames, T - k * kTileFrames);
                t.pad = 0;
                tiles[first + k] = t;
            }
        }
        __syncthreads();
        if (tid == 0) s_carry += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) *n_tiles = s_carry;
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
    float2* s_scratch = s_vtab + 1024;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) {
        s_tw[i] = tw_g[i];
        s_vtab[i] = vtab_g[i];
    }
    __syncthreads();
    float2* scratch = s_scratch + warp * kScratchFloat2;
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
            frame_inv(a, ynyq, scratch, s_tw, s_vtab, lane);
            float2* dst = reinterpret_cast<float2*>(out + f * kNfft);
            const float sc = 1.0f / 2048.0f;
#pragma unroll
            for (int r = 0; r < 32; ++r) dst[lane + 32 * r] = make_float2(a[r].x * sc, a[r].y * sc);
        }
    }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct GlWorkspace {
    UttDesc* utts;
    TileDesc* tiles;
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
    w.max_tiles = total_frames / kTileFrames + n_utts;
    w.wave_samples = (total_frames - n_utts) * (long long)plan->hop;
    w.mag_stride = (int)align_up((size_t)plan->kb, 4);
    w.utts = reinterpret_cast<UttDesc*>(take(sizeof(UttDesc) * (size_t)n_utts));
    w.tiles = reinterpret_cast<TileDesc*>(take(sizeof(TileDesc) * (size_t)w.max_tiles));
    w.n_tiles = reinterpret_cast<int*>(take(sizeof(int)));
    w.mag = reinterpret_cast<float*>(take(sizeof(float) * (size_t)total_frames * w.mag_stride));
    for (int i = 0; i < 2; ++i)
        w.buf[i] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)(w.wave_samples > 0 ? w.wave_samples : 1)));
    w.total = off;
    return w;
}

size_t gl_pass_smem(const s2st_plan* plan) {
    return sizeof(float2) * (2048 + kTileFrames * kScratchFloat2) +
           sizeof(float) * (2 * plan->wp + ((plan->hop + 3) & ~3) + (kTileFrames - 1) * plan->hop + plan->wp);
}

template <int NZ>
int launch_pass(const GlParams& p, bool first, int grid, size_t smem, cudaStream_t stream) {
    if (first) {
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_gl_pass<NZ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gl_pass<NZ, true><<<grid, kGlThreads, smem, stream>>>(p);
    } else {
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_gl_pass<NZ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_gl_pass<NZ, false><<<grid, kGlThreads, smem, stream>>>(p);
    }
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
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
    const long long blocks = (n_frames + kImFrames - 1) / kImFrames;
    k_inverse_mel<<<(unsigned)blocks, 256, sizeof(float) * kImFrames * plan->n_mels, stream>>>(
        logmel, is_log, n_frames, plan->n_mels, plan->inv_mel_t, plan->kb, plan->kb_pad, mag, out_stride, n_out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_rfft2048(const s2st_plan* plan, long long n, const float* in, float* out, bool inverse,
                    cudaStream_t stream) {
    if (n <= 0) return S2ST_OK;
    const size_t smem = sizeof(float2) * (2048 + 8 * kScratchFloat2);
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

int gl_run(const s2st_plan* plan_c, int n_utts, long long total_frames, const int32_t* frame_offsets,
           const float* logmel, const float* mag, int mag_kb, const float* phase, int n_iter,
           float* wave_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    s2st_plan* plan = const_cast<s2st_plan*>(plan_c);  // only the profiling state is mutated
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
    k_build_tiles<<<1, 1024, 0, stream>>>(frame_offsets, n_utts, plan->hop, w.utts, w.tiles, w.n_tiles);
    S2ST_CUDA_CHECK(cudaGetLastError());

    GlParams p;
    p.hop = plan->hop;
    p.half = plan->n_fft / 2;
    p.rot = plan->rot;
    p.ws = plan->ws;
    p.wp = plan->wp;
    p.nphase = plan->nphase;
    p.win_a = plan->win_a;
    p.win_s = plan->win_s;
    p.w2 = plan->w2;
    p.inv_wss = plan->inv_wss;
    p.tw = plan->tw;
    p.vtab = plan->vtab;
    p.utts = w.utts;
    p.tiles = w.tiles;
    p.n_tiles = w.n_tiles;
    p.phase = phase;
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
    const int grid = (int)min((long long)plan->num_sms * 2, w.max_tiles);
    // three rotating waveform buffers; the one the last pass writes is the caller's output
    float* ring[3];
    ring[n_iter % 3] = wave_out;
    ring[(n_iter + 1) % 3] = w.buf[0];
    ring[(n_iter + 2) % 3] = w.buf[1];
    // pass 0 accumulates its seams into ring[0]: clear it (later passes get theirs cleared by the pass before)
    S2ST_CUDA_CHECK(cudaMemsetAsync(ring[0], 0, sizeof(float) * (size_t)w.wave_samples, stream));
    for (int it = 0; it <= n_iter; ++it) {
        p.in = ring[(it + 2) % 3];
        p.out = ring[it % 3];
        p.zero_next = ring[(it + 1) % 3];
        if (timed) {
            if (!plan->timing_events[it]) S2ST_CUDA_CHECK(cudaEventCreate(&plan->timing_events[it]));
            S2ST_CUDA_CHECK(cudaEventRecord(plan->timing_events[it], stream));
        }
        int rc = (plan->nz == 19) ? launch_pass<19>(p, it == 0, grid, smem, stream)
                                  : launch_pass<32>(p, it == 0, grid, smem, stream);
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
