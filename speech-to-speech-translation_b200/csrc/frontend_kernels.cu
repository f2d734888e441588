// Feature front-end kernels (sm_100a): STFT magnitude/phase, logmelspec80, Kaldi fbank80, global CMVN.
//
// Replaces, for ragged batches resident in HBM:
//   TTSSpectrogram.forward            fairseq/data/audio/audio_utils.py:259-271
//   TTSMelScale.forward               audio_utils.py:284-285
//   extract_logmel_spectrogram        examples/speech_synthesis/data_utils.py:46-76
//   _get_torchaudio_fbank             audio_utils.py:136-149 (torchaudio.compliance.kaldi.fbank defaults)
//   GlobalCMVN.__call__               fairseq/data/audio/feature_transforms/global_cmvn.py:26-29
//   gcmvn_denormalize                 fairseq/speech_generator_for_s2st.py:21-29
//   get_global_cmvn (accumulation)    examples/speech_synthesis/data_utils.py:190-220
// One warp per frame; everything after the waveform read stays on-chip until the [T, n_mels] row is
// written, so the HBM traffic is the hop of new samples in and the feature row out.
#include <math_constants.h>

#include <cstdlib>

#include "../../include/s2st_b200.h"
#include "frame_fft.cuh"
#include "plan.h"

namespace s2st {

namespace {

// Sum of the slab floats a mel bin's gather list names (8 ints per bin, unused entries point at a float kept at
// zero).  Fixed trip count (4 or 8, warp-uniform).  The table is stored as two planes of int4 ([2][n_bins]): lanes
// with consecutive bins fetch consecutive 16-byte entries (the [n_bins][8] layout cost 4 extra wavefronts per load).
__device__ __forceinline__ float gather_sum(const float* __restrict__ slab, const int4* __restrict__ list4, int m,
                                            int n_bins, int terms) {
    const int4 g = list4[m];
    float acc = ((slab[g.x] + slab[g.y]) + slab[g.z]) + slab[g.w];
    if (terms > 4) {
        const int4 h = list4[n_bins + m];
        acc = (((acc + slab[h.x]) + slab[h.y]) + slab[h.z]) + slab[h.w];
    }
    return acc;
}

__device__ __forceinline__ int find_utt(const int32_t* __restrict__ fo, int n_utts, long long f) {
    int lo = 0, hi = n_utts - 1;  // last u with fo[u] <= f (utterances with zero frames are skipped)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (fo[mid] <= f) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct StftParams {
    int hop, half, rot, ws, n_mels, n_utts;
    long long total_frames;
    const float* win_a;
    const float2* tw;
    const float2* vtab;
    const int64_t* wave_offsets;
    const int32_t* frame_offsets;
    const float* wave;
    float* mag_out;
    float* phase_out;
    float* logmel_out;
    float eps;
    const float* cmvn_mean;
    const float* cmvn_std;
    double* sums;           // optional [2][n_mels]: += (sum, sum of squares) of the features BEFORE the CMVN
    const int* mel_ptr;
    const int* mel_idx;
    const float* mel_val;
    const float4* mel_col;  // k_logmel_fast: [22 * 32] column view of the mel bank (see s2st_plan::mel_col)
    const int* mel_gather;  //                [n_mels * 8] slab floats per mel bin
    int mel_terms;
};

// MODE 0: magnitude (+ optional phase) [F,] rows; MODE 1: log-mel (+ optional CMVN) rows.
template <int NZ, int MODE>
__global__ void __launch_bounds__(256, 2) k_stft(const __grid_constant__ StftParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);
    float2* s_vtab = s_tw + 1024;
    float* s_scratch = reinterpret_cast<float*>(s_vtab + 1024);
    float* s_win = s_scratch + 8 * kScratchFloats;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw[i];
        s_vtab[i] = p.vtab[i];
    }
    for (int i = tid; i < 64 * NZ; i += blockDim.x) s_win[i] = p.win_a[i];
    __syncthreads();
    float* scratch = s_scratch + warp * kScratchFloats;

    for (long long f = (long long)blockIdx.x * 8 + warp; f < p.total_frames; f += (long long)gridDim.x * 8) {
        const int u = find_utt(p.frame_offsets, p.n_utts, f);
        const int t = (int)(f - p.frame_offsets[u]);
        const long long woff = p.wave_offsets[u];
        const int n = (int)(p.wave_offsets[u + 1] - woff);
        const float* src = p.wave + woff;
        const int base = t * p.hop + p.rot - p.half + 2 * lane;
        float2 a[32];
#pragma unroll
        for (int r = 0; r < NZ; ++r) {
            int j0 = base + 64 * r, j1 = j0 + 1;
            j0 = j0 < 0 ? -j0 : j0;
            j1 = j1 < 0 ? -j1 : j1;
            j0 = j0 >= n ? 2 * (n - 1) - j0 : j0;
            j1 = j1 >= n ? 2 * (n - 1) - j1 : j1;
            j0 = min(max(j0, 0), n - 1);  // only reachable where the window is zero
            j1 = min(max(j1, 0), n - 1);
            a[r] = make_float2(__ldg(src + j0) * s_win[64 * r + 2 * lane], __ldg(src + j1) * s_win[64 * r + 2 * lane + 1]);
        }
        float nyq;
        frame_fwd<NZ>(a, nyq, scratch, s_tw, s_vtab, lane);
        if constexpr (MODE == 0) {
            float* mrow = p.mag_out + f * kBins;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int k = 32 * r + lane;
                const float x = 0.5f * a[r].x, y = 0.5f * a[r].y;
                mrow[k] = sqrtf(fmaf(x, x, y * y));
                if (p.phase_out) {
                    // undo the circular rotation: X[k] = X'[k] * exp(-2 pi i k rot / 2048)
                    float sn, cs;
                    sincospif(-(float)((k * p.rot) & 2047) * (1.0f / 1024.0f), &sn, &cs);
                    p.phase_out[f * kBins + k] = atan2f(fmaf(x, sn, y * cs), fmaf(x, cs, -y * sn));
                }
            }
            if (lane == 0) {
                mrow[1024] = fabsf(0.5f * nyq);
                if (p.phase_out) p.phase_out[f * kBins + 1024] = atan2f(0.0f, nyq)  /* exp(-i pi rot) = 1, rot is even */;
            }
        } else {
            float* spec = scratch;  // 1025 floats fit in the warp scratch
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float x = 0.5f * a[r].x, y = 0.5f * a[r].y;
                spec[32 * r + lane] = sqrtf(fmaf(x, x, y * y));
            }
            if (lane == 0) spec[1024] = fabsf(0.5f * nyq);
            __syncwarp();
            for (int m = lane; m < p.n_mels; m += 32) {
                float acc = 0.0f;
                const int e1 = __ldg(p.mel_ptr + m + 1);
                for (int e = __ldg(p.mel_ptr + m); e < e1; ++e)
                    acc = fmaf(__ldg(p.mel_val + e), spec[__ldg(p.mel_idx + e)], acc);
                float v = logf(fmaxf(acc, p.eps));
                if (p.sums) {  // generic path: straight to the accumulators (the fast kernels keep per-warp partials)
                    atomicAdd(p.sums + m, (double)v);
                    atomicAdd(p.sums + p.n_mels + m, (double)v * (double)v);
                }
                if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + m)) / __ldg(p.cmvn_std + m);
                p.logmel_out[f * p.n_mels + m] = v;
            }
            __syncwarp();
        }
    }
}

// log of a positive normal float as MUFU.LG2 + one multiply (2 instructions instead of the ~13 of logf): absolute error
// <= 2^-21.4 near 1, <= 3 ulp of the result elsewhere -- a few 1e-7 relative on features of magnitude 1 .. 16, inside the
// 1e-5 feature tolerance with a wide margin.  The register-resident extraction kernels spend a third of their
// instructions after the FFT; log and CMVN were 20 of the ~42 instructions per feature there.
__device__ __forceinline__ float fast_log(float x) { return __logf(x); }

// Fused global CMVN (feature_transforms/global_cmvn.py:26-29) inside the extraction kernels: (x - mean) / std evaluated
// as fma(x, 1 / std, -mean / std) with the two constants per feature prepared once per block (identity when no
// statistics are given).  Differs from the IEEE subtract-divide of the stand-alone kernel (which is bit-exact with
// numpy) by at most ~1.5 ulp; the stand-alone s2st_cmvn_apply stays bit-exact.
__device__ __forceinline__ void load_cmvn_table(float2* s_cm, const float* __restrict__ mean, const float* __restrict__ std, int n) {
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        float2 c = make_float2(1.0f, 0.0f);
        if (mean) {
            const float r = 1.0f / __ldg(std + m);
            c = make_float2(r, -__ldg(mean + m) * r);
        }
        s_cm[m] = c;
    }
    __syncthreads();
}

// Fused global-CMVN statistics (get_global_cmvn, examples/speech_synthesis/data_utils.py:190-220, without re-reading
// the corpus): a lane keeps float partial sums of the features it writes over one chunk of frames (8 or 16), folds them
// into a per-block double accumulator in shared memory at the end of the chunk, and the block adds that to
// sums[2][n] once at the end of the launch (one double atomicAdd per block and feature).
constexpr int kMaxStatCols = 128;
template <int N>
struct FeatureSums {
    float s[N], q[N];
    __device__ __forceinline__ void init(double* s_sums) {
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = q[i] = 0.0f;
        for (int i = threadIdx.x; i < 2 * kMaxStatCols; i += blockDim.x) s_sums[i] = 0.0;
        __syncthreads();
    }
    __device__ __forceinline__ void add(int i, float v) {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k == i) {
                s[k] += v;
                q[k] = fmaf(v, v, q[k]);
            }
    }
    __device__ __forceinline__ void fold(double* s_sums, int first, int stride, int n) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int m = first + stride * i;
            if (m < n) {
                atomicAdd(s_sums + m, (double)s[i]);
                atomicAdd(s_sums + kMaxStatCols + m, (double)q[i]);
            }
            s[i] = q[i] = 0.0f;
        }
    }
    static __device__ __forceinline__ void flush(const double* s_sums, double* sums, int n) {
        __syncthreads();
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            atomicAdd(sums + m, s_sums[m]);
            atomicAdd(sums + n + m, s_sums[kMaxStatCols + m]);
        }
    }
};

// logmelspec80, fast path (mel bank confined to bins < 704 with pairwise-overlapping triangles, i.e. the
// recipe's f_max = 8 kHz Slaney bank): a warp walks kLmChunk consecutive frames of the ragged batch (one
// utterance lookup per chunk); per frame: window -> in-place pruned 1024-point transform (fwd1024) -> pruned
// Hermitian split (22 rows) -> |X| -> column-wise mel (each lane owns 22 consecutive bins, see k_fbank_fast) ->
// log(max(., eps)) -> optional CMVN -> one 320-byte row.
constexpr int kLmChunk = 8;
constexpr int kLmCols = 32 * kPrunedRows;      // 704 spectrum bins
constexpr int kLmSlots = 17;                   // partial-sum slots per lane; slab[slot][lane] (lo, hi): a step's 32 stores hit 32 different bank pairs whatever the slots
constexpr int kLmSlabFloats = 2 * 32 * kLmSlots + 4;  // + the always-zero float unused gather entries point at
static_assert(kLmCols + 4 + kLmSlabFloats <= kScratchFloats, "spectrum + slab live in the warp scratch");
template <int NZ, bool SUMS>
__global__ void __launch_bounds__(256, 2) k_logmel_fast(const __grid_constant__ StftParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);
    float2* s_vtab = s_tw + 1024;
    float4* s_col = reinterpret_cast<float4*>(s_vtab + 1024);            // [22][32]: entry of bin 22 * lane + j at [j][lane]
    int* s_gather = reinterpret_cast<int*>(s_col + kLmCols);            // [n_mels][8]
    float* s_win = reinterpret_cast<float*>(s_gather + 8 * 128);
    float* s_warp = s_win + 64 * NZ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) {
        s_tw[i] = p.tw[i];
        s_vtab[i] = p.vtab[i];
    }
    for (int i = tid; i < kLmCols; i += blockDim.x) s_col[i] = p.mel_col[i];
    for (int i = tid; i < 8 * p.n_mels; i += blockDim.x) s_gather[i] = p.mel_gather[i];
    for (int i = tid; i < 64 * NZ; i += blockDim.x) s_win[i] = p.win_a[i];
    float* scratch = s_warp + warp * kScratchFloats;
    float* slab = scratch + kLmCols + 4;  // [17][32] (lo, hi) pairs, then the zero float unused gather entries read
    __syncthreads();
    const long long n_chunks = (p.total_frames + kLmChunk - 1) / kLmChunk;
    FeatureSums<SUMS ? 4 : 1> fs;  // mel bins lane, lane + 32, ... (n_mels <= 128)
    __shared__ double s_sums[SUMS ? 2 * kMaxStatCols : 1];
    __shared__ float2 s_cm[kMaxStatCols];  // fused CMVN as one FMA per feature: (1 / std, -mean / std)
    load_cmvn_table(s_cm, p.cmvn_mean, p.cmvn_std, p.n_mels);
    if constexpr (SUMS) fs.init(s_sums);
    for (long long chunk = (long long)blockIdx.x * 8 + warp; chunk < n_chunks; chunk += (long long)gridDim.x * 8) {
        long long f = chunk * kLmChunk;
        int u = find_utt(p.frame_offsets, p.n_utts, f);
        // the utterance's table entries stay in registers while the chunk stays inside it (no dependent loads per frame)
        long long fo_u = __ldg(p.frame_offsets + u), fo_next = __ldg(p.frame_offsets + u + 1);
        long long woff = __ldg(p.wave_offsets + u);
        int n = (int)(__ldg(p.wave_offsets + u + 1) - woff);
        {
            // the chunk's frames are (mostly) consecutive hops of one utterance: ask L2 for their samples now
            const long long b0 = (f - fo_u) * p.hop + p.rot - p.half;
            const long long span = (long long)(kLmChunk - 1) * p.hop + 64 * NZ;
            for (long long j = b0 + 32 * lane; j < b0 + span; j += 32 * 32)
                if (j >= 0 && j < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.wave + woff + j));
        }
#pragma unroll 1
        for (int it = 0; it < kLmChunk && f < p.total_frames; ++it, ++f) {
            while (u + 1 < p.n_utts && f >= fo_next) {
                ++u;
                fo_u = fo_next;
                fo_next = __ldg(p.frame_offsets + u + 1);
                woff = __ldg(p.wave_offsets + u);
                n = (int)(__ldg(p.wave_offsets + u + 1) - woff);
            }
            const int t = (int)(f - fo_u);
            const float* src = p.wave + woff;
            const int base0 = t * p.hop + p.rot - p.half;
            float2 a[32];
            if (base0 >= 0 && base0 + 64 * NZ <= n) {
                // interior frame: no reflection.  An utterance that starts at an odd sample of the concatenated
                // buffer cannot use 8-byte loads; it still skips the index arithmetic of the edge path
                if (((woff + base0) & 1) == 0) {
                    const float2* s2 = reinterpret_cast<const float2*>(src + base0) + lane;
#pragma unroll
                    for (int r = 0; r < NZ; ++r) a[brev5(r)] = __ldg(s2 + 32 * r);
                } else {
                    const float* s1 = src + base0 + 2 * lane;
#pragma unroll
                    for (int r = 0; r < NZ; ++r) a[brev5(r)] = make_float2(__ldg(s1 + 64 * r), __ldg(s1 + 64 * r + 1));
                }
            } else {
                const int base = base0 + 2 * lane;
#pragma unroll
                for (int r = 0; r < NZ; ++r) {
                    int j0 = base + 64 * r, j1 = j0 + 1;
                    j0 = j0 < 0 ? -j0 : j0;
                    j1 = j1 < 0 ? -j1 : j1;
                    j0 = j0 >= n ? 2 * (n - 1) - j0 : j0;
                    j1 = j1 >= n ? 2 * (n - 1) - j1 : j1;
                    j0 = min(max(j0, 0), n - 1);  // only reachable where the window is zero
                    j1 = min(max(j1, 0), n - 1);
                    a[brev5(r)] = make_float2(__ldg(src + j0), __ldg(src + j1));
                }
            }
#pragma unroll
            for (int r = 0; r < NZ; ++r)
                a[brev5(r)] = mul2(a[brev5(r)], *reinterpret_cast<const float2*>(s_win + 64 * r + 2 * lane));
            fwd1024<(NZ > 16 ? NZ : 32), 32>(a, scratch, s_tw, lane);
            float nyq;
            fwd_split<true>(a, nyq, scratch, s_vtab, lane);
            // |X| (the split returns 2 X) to shared memory, linear in k
#pragma unroll
            for (int r = 0; r < kPrunedRows; ++r) {
                float mag;  // |2 X|: MUFU.SQRT (2 ulp) instead of the IEEE sequence; the features are compared at 1e-5
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(a[r].x, a[r].x, a[r].y * a[r].y)));
                scratch[32 * r + lane + (r >= 11 ? 1 : 0)] = 0.5f * mag;  // one float of padding after bin 351: see sp below
            }
            __syncwarp();
            {
                // column-wise mel: running (lo, hi) partial sums per run of bins that feed the same mel bin, written
                // to the run's slot after every step (the last write of a run is its total): no branches, no atomics
                float2 lh = make_float2(0.0f, 0.0f);  // (lo, hi) as one packed pair: 2 instructions per step
                // lane l owns bins 22 l .. 22 l + 21; lanes l and l + 16 would start 352 floats = 0 banks apart, the
                // padding float after bin 351 moves the upper half-warp by one bank: conflict-free reads
                const float* sp = scratch + kPrunedRows * lane + (lane >> 4);
                float2* my = reinterpret_cast<float2*>(slab) + lane;  // slab[slot][lane]: the table holds slot * 32
                // two batches of 11 steps: all loads of a batch are issued before its first store -- written as one
                // load / FMA / store loop the compiler must keep every load behind the previous step's slab store (same
                // address space, possible alias) and each step waits a full shared-memory round trip: 22 exposed
                // latencies per frame instead of 2
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    constexpr int kB = kPrunedRows / 2;
                    float4 c[kB];
                    float pv[kB];
                    float2 run[kB];
#pragma unroll
                    for (int j = 0; j < kB; ++j) {
                        c[j] = s_col[(h * kB + j) * 32 + lane];
                        pv[j] = sp[h * kB + j];
                    }
#pragma unroll
                    for (int j = 0; j < kB; ++j) {
                        lh = fma2(make_float2(c[j].x, c[j].y), bcast2(pv[j]), mul2(lh, bcast2(c[j].z)));
                        run[j] = lh;
                    }
#pragma unroll
                    for (int j = 0; j < kB; ++j) my[__float_as_int(c[j].w)] = run[j];
                }
                if (lane == 0) slab[2 * 32 * kLmSlots] = 0.0f;  // (the transposes use the whole scratch)
            }
            __syncwarp();
            for (int m = lane, i = 0; m < p.n_mels; m += 32, ++i) {
                const float acc = gather_sum(slab, reinterpret_cast<const int4*>(s_gather), m, p.n_mels, p.mel_terms);
                float v = fast_log(fmaxf(acc, p.eps));
                if constexpr (SUMS) fs.add(i, v);
                const float2 cm = s_cm[m];
                p.logmel_out[f * p.n_mels + m] = fmaf(v, cm.x, cm.y);
            }
            __syncwarp();
        }
        if constexpr (SUMS) fs.fold(s_sums, lane, 32, p.n_mels);
    }
    if constexpr (SUMS) fs.flush(s_sums, p.sums, p.n_mels);
}

// ---------------------------------------------------------------------------------------------
// STFT / log-mel for any power-of-two n_fft in [64, 4096] (the reference's extractors default to n_fft = 1024,
// examples/speech_synthesis/data_utils.py:46-52; its dense-basis convolution takes any size).  One warp per frame:
// reflect padding + window (audio_utils.py:259-263), the real frame packed as n_fft / 2 complex points, radix-2
// Stockham FFT in two shared-memory ping-pong buffers, Hermitian split, then magnitude / phase rows or mel -> log ->
// CMVN rows exactly like k_stft.  Not tuned: the 2048-point recipe geometry has its own register-resident kernels.
struct StftGenericParams {
    int n_fft, n_bins, hop, n_mels, n_utts;
    long long total_frames;
    const float* win;       // [n_fft]
    const float2* tw;       // [n_fft / 2]
    const int64_t* wave_offsets;
    const int32_t* frame_offsets;
    const float* wave;
    float* mag_out;
    float* phase_out;
    float* logmel_out;
    float eps;
    const float* cmvn_mean;
    const float* cmvn_std;
    double* sums;
    const int* mel_ptr;
    const int* mel_idx;
    const float* mel_val;
};

template <int MODE>
__global__ void k_stft_generic(const __grid_constant__ StftGenericParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int half = p.n_fft >> 1;
    float2* buf0 = reinterpret_cast<float2*>(smem_raw) + (size_t)warp * (2 * half + 2);
    float2* buf1 = buf0 + half + 1;  // + 1: the spectrum (half + 1 values) is staged here in MODE 1
    for (long long f = (long long)blockIdx.x * warps + warp; f < p.total_frames; f += (long long)gridDim.x * warps) {
        const int u = find_utt(p.frame_offsets, p.n_utts, f);
        const int t = (int)(f - p.frame_offsets[u]);
        const long long woff = p.wave_offsets[u];
        const int n = (int)(p.wave_offsets[u + 1] - woff);
        const float* src = p.wave + woff;
        const int base = t * p.hop - half;
        float* y = reinterpret_cast<float*>(buf0);
        for (int i = lane; i < p.n_fft; i += 32) {
            int j = base + i;
            j = j < 0 ? -j : j;
            j = j >= n ? 2 * (n - 1) - j : j;
            j = min(max(j, 0), n - 1);
            y[i] = __ldg(src + j) * __ldg(p.win + i);
        }
        __syncwarp();
        float2* in = buf0;
        float2* outb = buf1;
        for (int ns = 1; ns < half; ns <<= 1) {
            const int tstride = p.n_fft / (2 * ns);  // W_{2 ns}^k = W_{n_fft}^{k * n_fft / (2 ns)}
            for (int j = lane; j < (half >> 1); j += 32) {
                const int k = j & (ns - 1);
                const float2 w = __ldg(p.tw + k * tstride);
                const float2 a = in[j];
                const float2 b = cmul(in[j + (half >> 1)], w);
                const int j0 = ((j - k) << 1) + k;
                outb[j0] = make_float2(a.x + b.x, a.y + b.y);
                outb[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncwarp();
            float2* tmp = in;
            in = outb;
            outb = tmp;
        }
        // X[k] = E + W^k O,  E = (Z[k] + conj Z[half - k]) / 2,  O = (Z[k] - conj Z[half - k]) / (2 i);  X[half] = Re Z[0] - Im Z[0]
        float* spec = reinterpret_cast<float*>(outb);  // MODE 1: |X| for the mel projection (half + 1 floats fit)
        for (int k = lane; k <= half; k += 32) {
            float xr, xi;
            if (k == half) {
                xr = in[0].x - in[0].y;
                xi = 0.0f;
            } else {
                const float2 z = in[k];
                const float2 zp = in[(half - k) & (half - 1)];
                const float2 w = __ldg(p.tw + k);
                const float ex = 0.5f * (z.x + zp.x), ey = 0.5f * (z.y - zp.y);
                const float ox = 0.5f * (z.y + zp.y), oy = -0.5f * (z.x - zp.x);
                xr = ex + (w.x * ox - w.y * oy);
                xi = ey + (w.x * oy + w.y * ox);
            }
            const float mag = sqrtf(fmaf(xr, xr, xi * xi));
            if constexpr (MODE == 0) {
                p.mag_out[f * p.n_bins + k] = mag;
                if (p.phase_out) p.phase_out[f * p.n_bins + k] = atan2f(xi, xr);
            } else {
                spec[k] = mag;
            }
        }
        __syncwarp();
        if constexpr (MODE == 1) {
            for (int m = lane; m < p.n_mels; m += 32) {
                float acc = 0.0f;
                const int e1 = __ldg(p.mel_ptr + m + 1);
                for (int e = __ldg(p.mel_ptr + m); e < e1; ++e)
                    acc = fmaf(__ldg(p.mel_val + e), spec[__ldg(p.mel_idx + e)], acc);
                float v = logf(fmaxf(acc, p.eps));
                if (p.sums) {
                    atomicAdd(p.sums + m, (double)v);
                    atomicAdd(p.sums + p.n_mels + m, (double)v * (double)v);
                }
                if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + m)) / __ldg(p.cmvn_std + m);
                p.logmel_out[f * p.n_mels + m] = v;
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(256) k_mel_project(long long n_frames, int n_mels, int n_bins, const float* __restrict__ spec,
                                                      const int* __restrict__ mel_ptr, const int* __restrict__ mel_idx,
                                                      const float* __restrict__ mel_val, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long f = (long long)blockIdx.x * 8 + warp; f < n_frames; f += (long long)gridDim.x * 8) {
        const float* row = spec + f * n_bins;
        for (int m = lane; m < n_mels; m += 32) {
            float acc = 0.0f;
            const int e1 = mel_ptr[m + 1];
            for (int e = mel_ptr[m]; e < e1; ++e) acc = fmaf(mel_val[e], __ldg(row + mel_idx[e]), acc);
            out[f * n_mels + m] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Kaldi fbank: one warp per frame, radix-2 Stockham FFT of padded/2 complex points in shared memory.
struct FbankParams {
    int win, shift, padded, n_bins, n_utts;
    long long total_frames;
    const float* window;
    const float2* tw;  // [padded/2] exp(-2 pi i j / padded)
    const int64_t* wave_offsets;
    const int32_t* frame_offsets;
    const float* wave;
    const float* cmvn_mean;
    const float* cmvn_std;
    const int* mel_ptr;
    const int* mel_idx;
    const float* mel_val;
    float* out;
    double* sums;  // optional [2][n_bins] (see StftParams::sums)
};

constexpr int kFbankWarps = 8;

__global__ void __launch_bounds__(32 * kFbankWarps) k_fbank(const __grid_constant__ FbankParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = p.padded >> 1;
    // per warp: two ping-pong buffers of `half` float2
    float2* buf0 = reinterpret_cast<float2*>(smem_raw) + (size_t)warp * 2 * half;
    float2* buf1 = buf0 + half;
    float* raw = reinterpret_cast<float*>(buf1);  // raw samples staged in buf1 first
    float* y = reinterpret_cast<float*>(buf0);    // windowed frame == packed complex input

    for (long long f = (long long)blockIdx.x * kFbankWarps + warp; f < p.total_frames;
         f += (long long)gridDim.x * kFbankWarps) {
        const int u = find_utt(p.frame_offsets, p.n_utts, f);
        const int t = (int)(f - p.frame_offsets[u]);
        const float* src = p.wave + p.wave_offsets[u] + (long long)t * p.shift;
        float sum = 0.0f;
        for (int j = lane; j < p.win; j += 32) {
            const float v = __ldg(src + j);
            raw[j] = v;
            sum += v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        const float mean = sum / (float)p.win;
        __syncwarp();
        // remove DC, pre-emphasis 0.97 with replicate padding, povey window, zero pad
        for (int j = lane; j < p.padded; j += 32) {
            float v = 0.0f;
            if (j < p.win) {
                const float c = raw[j] - mean;
                const float pr = raw[j > 0 ? j - 1 : 0] - mean;
                v = (c - 0.97f * pr) * __ldg(p.window + j);
            }
            y[j] = v;
        }
        __syncwarp();
        // Stockham radix-2, half complex points
        float2* in = buf0;
        float2* outb = buf1;
        for (int ns = 1; ns < half; ns <<= 1) {
            const int tstride = p.padded / (2 * ns);  // W_{2ns}^k = W_padded^{k * padded / (2 ns)}
            for (int j = lane; j < (half >> 1); j += 32) {
                const int k = j & (ns - 1);
                const float2 w = __ldg(p.tw + k * tstride);
                const float2 a = in[j];
                const float2 b = cmul(in[j + (half >> 1)], w);
                const int j0 = ((j - k) << 1) + k;
                outb[j0] = make_float2(a.x + b.x, a.y + b.y);
                outb[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
            }
            __syncwarp();
            float2* tmp = in;
            in = outb;
            outb = tmp;
        }
        // real split + power spectrum into the free buffer (bins 0 .. half-1; Kaldi drops Nyquist)
        float* power = reinterpret_cast<float*>(outb);
        for (int k = lane; k < half; k += 32) {
            const float2 z = in[k];
            const float2 zp = in[(half - k) & (half - 1)];
            const float2 w = __ldg(p.tw + k);
            // X[k] = E + W^k O,  E = (Z + conj Zp)/2,  O = (Z - conj Zp)/(2i)
            const float ex = 0.5f * (z.x + zp.x), ey = 0.5f * (z.y - zp.y);
            const float ox = 0.5f * (z.y + zp.y), oy = -0.5f * (z.x - zp.x);
            const float xr = ex + (w.x * ox - w.y * oy);
            const float xi = ey + (w.x * oy + w.y * ox);
            power[k] = fmaf(xr, xr, xi * xi);
        }
        __syncwarp();
        for (int m = lane; m < p.n_bins; m += 32) {
            float acc = 0.0f;
            const int e1 = __ldg(p.mel_ptr + m + 1);
            for (int e = __ldg(p.mel_ptr + m); e < e1; ++e)
                acc = fmaf(__ldg(p.mel_val + e), power[__ldg(p.mel_idx + e)], acc);
            float v = logf(fmaxf(acc, 1.1920928955078125e-07f));
            if (p.sums) {
                atomicAdd(p.sums + m, (double)v);
                atomicAdd(p.sums + p.n_bins + m, (double)v * (double)v);
            }
            if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + m)) / __ldg(p.cmvn_std + m);
            p.out[f * p.n_bins + m] = v;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Kaldi fbank, register-resident version for the two speech rates of the recipe: 16 kHz (25 ms = 400 samples,
// FFT 512) and 8 kHz (200 samples, FFT 256).  A half-warp (16 lanes) owns a frame; everything between the
// waveform read and the feature row stays in registers / 2 KB of shared memory per half-warp:
//   MODE 0 (FFT 512): the real frame is packed as 256 complex points z[n] = y[2n] + i y[2n+1];
//   MODE 1 (FFT 256): TWO frames share one transform, z[n] = yA[n] + i yB[n] (two-for-one real FFT);
// either way a 256-point complex DFT = in-lane FFT-16 (fft32.cuh, packed f32x2 arithmetic), twiddle,
// 16 x 16 transpose through the swizzled scratch, in-lane FFT-16, then the Hermitian split, |X|^2,
// sparse mel, log, optional CMVN.  A half-warp walks kFbChunk consecutive frames of the ragged batch, so
// the utterance lookup is one binary search per chunk.
// Mel projection: Kaldi's triangles overlap pairwise, so every FFT bin k feeds at most two ADJACENT mel bins
// (checked when the plan is built).  Each sub-lane walks a contiguous run of bins with a column table
// (weight into bin b, weight into bin b + 1, b), accumulates in registers while b stays the same and adds
// the partial sums to a small shared-memory accumulator when it changes: balanced across lanes (the row-wise
// gather is not: high mel bins are ten times wider than low ones) and free of dependent index loads.
#ifndef S2ST_FB_BLOCKS
#define S2ST_FB_BLOCKS 3
#endif
constexpr int kFbBlocks0 = S2ST_FB_BLOCKS;  // resident blocks per SM of the 16 kHz kernel: 3 (80 registers, no spills) measured 2 % faster than 4 (64 registers, 44 B spilled)
constexpr int kFbWarps = 8;
constexpr int kFbChunk = 16;
constexpr int kFbRows = 13;                 // rows of 16 elements that can hold window samples (13 * 16 >= 200)
constexpr int kFbScratchBytes = 2048;       // per half-warp: 16 x 16 complex
constexpr int kFbPwrFloats = 16 * 17;          // MODE 0: power spectrum staging; the mel slab follows it
constexpr int kFbSlabRows = 13;                // MODE 0: most runs (mel bin changes) one sub-lane goes through; slab = [rows][16] (lo, hi)
// MODE 0 bytes per half-warp: FFT scratch (2048) overlaid by power spectrum + slab (+ the always-zero float), rounded so
// that the regions of a warp's two half-warps start 16 banks apart (their 16-lane accesses then never share a bank)
constexpr int kFbRegion0Floats = 720;
static_assert(kFbRegion0Floats >= kFbPwrFloats + 2 * 16 * kFbSlabRows + 1 && kFbRegion0Floats * 4 >= 2048 && kFbRegion0Floats % 32 == 16, "region layout");

struct FbankFastParams {
    int win, shift, n_bins, n_utts, mel_nnz;
    long long total_frames;
    const float2* tw16;    // [256] exp(-2 pi i k1 n2 / 256) at [k1 * 16 + n2]
    const float2* vsplit;  // [256] -i exp(-2 pi i k / 512)            (MODE 0)
    const float* winp;     // MODE 0: [13 * 16 * 2] (w[2n], w[2n+1]) zero padded; MODE 1: [13 * 16] w[n] zero padded
    const int64_t* wave_offsets;
    const int32_t* frame_offsets;
    const float* wave;
    const float* cmvn_mean;
    const float* cmvn_std;
    const float4* mel_col;  // MODE 0: [256] (w into bin b, w into bin b + 1, run continues, slot) in the order the sub-lanes read it
    const int* mel_gather;  // MODE 0: [n_bins * 8] slab floats that add up to each mel bin (s2st_fbank_plan::mel_gather)
    int mel_terms, mel_zero;
    const int* mel_ptr;     // MODE 1: CSR rows of the mel bank (its 128-bin rows are short: the row gather wins)
    const int* mel_idx;
    const float* mel_val;
    float* out;
    double* sums;           // optional [2][n_bins] (see StftParams::sums)
};

// 16 x 16 transpose inside a half-warp: out a[brev4(r)] = (sub-lane r's) a[sub].  Same XOR swizzle as
// warp_transpose: element (row j, column l) sits in 16-byte chunk ((l >> 1) ^ (j & 7)) of its 128-byte row.
__device__ __forceinline__ void group_transpose16(float2 (&a)[16], char* scratch, int sub) {
    const int wofs = sub * 8;
#pragma unroll
    for (int j = 0; j < 16; ++j) *reinterpret_cast<float2*>(scratch + j * 128 + (wofs ^ ((j & 7) << 4))) = a[j];
    __syncwarp();
    const char* row = scratch + sub * 128;
    const int sw = (sub & 7) << 4;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(row + ((c << 4) ^ sw));
        a[brev4(2 * c)] = make_float2(v.x, v.y);
        a[brev4(2 * c + 1)] = make_float2(v.z, v.w);
    }
    __syncwarp();
}

template <int MODE, bool SUMS>
__global__ void __launch_bounds__(32 * kFbWarps, MODE == 0 ? kFbBlocks0 : 3) k_fbank_fast(const __grid_constant__ FbankFastParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                 // 256
    float2* s_vs = s_tw + 256;                                          // 256
    float* s_win = reinterpret_cast<float*>(s_vs + 256);                // 13 * 16 * 2
    float4* s_col = reinterpret_cast<float4*>(s_win + kFbRows * 32);    // MODE 0: 256 column entries
    int2* s_mel = reinterpret_cast<int2*>(s_col);                       // MODE 1: mel_nnz CSR entries (idx, val bits)
    int* s_ptr = reinterpret_cast<int*>(s_mel + ((p.mel_nnz + 1) & ~1));  //         n_bins + 1 row pointers
    int* s_gather = reinterpret_cast<int*>(s_col + 256);                // MODE 0: [n_bins][8]
    char* s_scr = MODE == 0 ? reinterpret_cast<char*>(s_gather + ((8 * p.n_bins + 3) & ~3))
                            : reinterpret_cast<char*>(s_ptr + ((p.n_bins + 1 + 3) & ~3));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane & 15, grp = lane >> 4;
    for (int i = tid; i < 256; i += blockDim.x) {
        s_tw[i] = p.tw16[i];
        if (MODE == 0) s_vs[i] = p.vsplit[i];
    }
    for (int i = tid; i < kFbRows * 16 * (MODE == 0 ? 2 : 1); i += blockDim.x) s_win[i] = p.winp[i];
    if constexpr (MODE == 0) {
        for (int i = tid; i < 256; i += blockDim.x) s_col[i] = p.mel_col[i];
        for (int i = tid; i < 8 * p.n_bins; i += blockDim.x) s_gather[i] = p.mel_gather[i];
    } else {
        for (int i = tid; i < p.mel_nnz; i += blockDim.x) s_mel[i] = make_int2(p.mel_idx[i], __float_as_int(p.mel_val[i]));
        for (int i = tid; i <= p.n_bins; i += blockDim.x) s_ptr[i] = p.mel_ptr[i];
    }
    __syncthreads();
    const int acc_floats = ((MODE + 1) * (p.n_bins + 1) + 3) & ~3;
    const size_t region = MODE == 0 ? (size_t)4 * kFbRegion0Floats : (size_t)(kFbScratchBytes + 4 * acc_floats);
    char* scratch = s_scr + (size_t)(warp * 2 + grp) * region;
    float* pwr = reinterpret_cast<float*>(scratch);
    float* macc = reinterpret_cast<float*>(scratch + kFbScratchBytes);
    const int prev_lane = (lane & 16) | ((sub + 15) & 15);
    const int n_mel_iter = (p.n_bins + 15) >> 4;

    const long long n_chunks = (p.total_frames + kFbChunk - 1) / kFbChunk;
    FeatureSums<SUMS ? 8 : 1> fs;  // mel bins sub, sub + 16, ... (n_bins <= 128)
    __shared__ double s_sums[SUMS ? 2 * kMaxStatCols : 1];
    __shared__ float2 s_cm[kMaxStatCols];  // fused CMVN as one FMA per feature: (1 / std, -mean / std)
    load_cmvn_table(s_cm, p.cmvn_mean, p.cmvn_std, p.n_bins);
    if constexpr (SUMS) fs.init(s_sums);
    for (long long cp = (long long)blockIdx.x * kFbWarps + warp; 2 * cp < n_chunks; cp += (long long)gridDim.x * kFbWarps) {
        long long f = (2 * cp + grp) * kFbChunk;
        int u = find_utt(p.frame_offsets, p.n_utts, f < p.total_frames ? f : p.total_frames - 1);
        // the utterance's table entries stay in registers while the chunk stays inside it (no dependent loads per frame)
        long long fo_u = __ldg(p.frame_offsets + u), fo_next = __ldg(p.frame_offsets + u + 1);
        const float* wave_u = p.wave + __ldg(p.wave_offsets + u);
#pragma unroll 1
        for (int it = 0; it < kFbChunk / (MODE + 1); ++it) {
            // ---- resolve the frame(s) of this step
            const float* src[MODE + 1];
            bool valid[MODE + 1];
#pragma unroll
            for (int q = 0; q <= MODE; ++q) {
                valid[q] = f < p.total_frames;
                if (valid[q]) {
                    while (u + 1 < p.n_utts && f >= fo_next) {
                        ++u;
                        fo_u = fo_next;
                        fo_next = __ldg(p.frame_offsets + u + 1);
                        wave_u = p.wave + __ldg(p.wave_offsets + u);
                    }
                    src[q] = wave_u + (f - fo_u) * p.shift;
                } else {
                    src[q] = p.wave;  // a readable address; nothing is stored for this frame
                }
                ++f;
            }
            // ---- load, frame mean
            float2 a[16];
            if constexpr (MODE == 0) {
                const bool al = (reinterpret_cast<uintptr_t>(src[0]) & 7) == 0;
                if (valid[0] && (p.win & 1) == 0) {
                    // the common case (even window length): predicated loads, no nested branches -- one 8-byte load
                    // per row when the frame starts 8-byte aligned, two 4-byte loads otherwise (an utterance that
                    // begins at an odd sample of the concatenated buffer)
                    const int n2 = (p.win >> 1) - sub;  // element 16 r + sub exists iff 16 r < n2
                    if (al) {
                        const float2* s2 = reinterpret_cast<const float2*>(src[0]) + sub;
#pragma unroll
                        for (int r = 0; r < kFbRows; ++r) a[r] = 16 * r < n2 ? __ldg(s2 + 16 * r) : make_float2(0.0f, 0.0f);
                    } else {
                        const float* s1 = src[0] + 2 * sub;
#pragma unroll
                        for (int r = 0; r < kFbRows; ++r)
                            a[r] = 16 * r < n2 ? make_float2(__ldg(s1 + 32 * r), __ldg(s1 + 32 * r + 1)) : make_float2(0.0f, 0.0f);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kFbRows; ++r) {
                        const int j = 32 * r + 2 * sub;
                        float2 v = make_float2(0.0f, 0.0f);
                        if (valid[0] && j + 1 < p.win) {
                            if (al) v = __ldg(reinterpret_cast<const float2*>(src[0] + j));
                            else v = make_float2(__ldg(src[0] + j), __ldg(src[0] + j + 1));
                        } else if (valid[0] && j < p.win) {
                            v.x = __ldg(src[0] + j);
                        }
                        a[r] = v;
                    }
                }
                float sum = 0.0f;
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) sum += a[r].x + a[r].y;
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
                const float mean = sum / (float)p.win;
                // DC removal, pre-emphasis 0.97 with replicate padding, povey window.  The previous sample of
                // an element's first half is the second half of the element one sub-lane to the left.
                float carry = 0.0f;
                float2 y[kFbRows];
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) {
                    const float2 c = make_float2(a[r].x - mean, a[r].y - mean);
                    const float left = __shfl_sync(0xffffffffu, c.y, prev_lane);
                    const float prev = sub > 0 ? left : (r == 0 ? c.x : carry);
                    carry = left;  // for sub == 0: the last element of the row above
                    const float2 w = *reinterpret_cast<const float2*>(s_win + 32 * r + 2 * sub);
                    y[r] = make_float2(fmaf(-0.97f, prev, c.x) * w.x, fmaf(-0.97f, c.x, c.y) * w.y);
                }
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) a[brev4(r)] = y[r];
            } else {
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) {
                    const int j = 16 * r + sub;
                    float2 v = make_float2(0.0f, 0.0f);
                    if (j < p.win) {
                        if (valid[0]) v.x = __ldg(src[0] + j);
                        if (valid[1]) v.y = __ldg(src[1] + j);
                    }
                    a[r] = v;
                }
                float2 sum = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) sum = add2(sum, a[r]);
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) {
                    sum.x += __shfl_xor_sync(0xffffffffu, sum.x, d);
                    sum.y += __shfl_xor_sync(0xffffffffu, sum.y, d);
                }
                const float2 nmean = make_float2(-(sum.x / (float)p.win), -(sum.y / (float)p.win));
                float2 carry = make_float2(0.0f, 0.0f);
                float2 y[kFbRows];
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) {
                    const float2 c = add2(a[r], nmean);
                    const float2 left = make_float2(__shfl_sync(0xffffffffu, c.x, prev_lane),
                                                    __shfl_sync(0xffffffffu, c.y, prev_lane));
                    const float2 prev = sub > 0 ? left : (r == 0 ? c : carry);
                    carry = left;
                    y[r] = mul2(fma2(prev, bcast2(-0.97f), c), bcast2(s_win[16 * r + sub]));
                }
#pragma unroll
                for (int r = 0; r < kFbRows; ++r) a[brev4(r)] = y[r];
            }
            // ---- 256-point complex DFT (rows >= 13 of the input are the zero padding)
            fft16_inplace_br<kFbRows, 16>(a);
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) a[k1] = cmul(a[k1], s_tw[k1 * 16 + sub]);
            group_transpose16(a, scratch, sub);
            fft16_inplace_br<16, 16>(a);
            // ---- Hermitian split through the scratch (linear in k = sub + 16 * k2), power spectrum
            float2* sc = reinterpret_cast<float2*>(scratch);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) sc[16 * k2 + sub] = a[k2];
            __syncwarp();
            float pw[16];
            if constexpr (MODE == 0) {
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    const int k = 16 * k2 + sub;
                    const float2 x2 = split_fwd(a[k2], sc[(256 - k) & 255], s_vs[k]);  // 2 X[k]
                    pw[k2] = 0.25f * fmaf(x2.x, x2.x, x2.y * x2.y);
                }
            } else {
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) {
                    const int k = 16 * k2 + sub;
                    const float2 pz = sc[(256 - k) & 255];
                    const float2 sa = add2(a[k2], conj2(pz)), sb = add2(a[k2], neg2(conj2(pz)));  // 2 XA, 2i XB
                    pw[k2] = 0.25f * fmaf(sa.x, sa.x, sa.y * sa.y);
                    pw[k2 + 8] = 0.25f * fmaf(sb.x, sb.x, sb.y * sb.y);
                }
            }
            __syncwarp();
            if constexpr (MODE == 0) {
                // power spectrum to shared memory, bin k at [(k & 15) * 17 + (k >> 4)] (conflict-free for this write
                // and for the run-per-lane read below)
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) pwr[sub * 17 + k2] = pw[k2];
                __syncwarp();
                // mel: sub-lane s owns FFT bins [16 s, 16 s + 16).  Running (lo, hi) partial sums per run of bins that
                // feed the same mel bin, stored to the run's slot after every step (the last store of a run is its
                // total): no branches, no atomics.  The slab (one (lo, hi) pair per run, ~100 runs per frame) sits
                // behind the power spectrum.
                float2 lh = make_float2(0.0f, 0.0f);  // (lo, hi) as one packed pair: 2 instructions per step
                float2* slab2 = reinterpret_cast<float2*>(pwr + kFbPwrFloats);
                // two batches of 8 steps, loads first: a load / FMA / store loop serialises on the slab stores (possible
                // alias with the next step's loads), i.e. one exposed shared-memory round trip per step
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 c[8];
                    float pv[8];
                    float2 run[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        c[j] = s_col[(8 * h + j) * 16 + sub];
                        pv[j] = pwr[(8 * h + j) * 17 + sub];  // bin 16 sub + 8 h + j
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        lh = fma2(make_float2(c[j].x, c[j].y), bcast2(pv[j]), mul2(lh, bcast2(c[j].z)));
                        run[j] = lh;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) slab2[__float_as_int(c[j].w)] = run[j];
                }
                if (sub == 0) pwr[kFbPwrFloats + p.mel_zero] = 0.0f;  // the float unused gather entries point at
            } else {
                // frame A at [0, 128), frame B at [128, 256); row gather: sub-lane s owns mel bins s, s + 16, ...
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) pwr[16 * k2 + sub] = pw[k2];
                __syncwarp();
                for (int i = 0; i < n_mel_iter; ++i) {
                    const int m = sub + 16 * i;
                    if (m < p.n_bins) {
                        float acc0 = 0.0f, acc1 = 0.0f;
                        const int e1 = s_ptr[m + 1];
                        for (int e = s_ptr[m]; e < e1; ++e) {
                            const int2 ent = s_mel[e];
                            const float w = __int_as_float(ent.y);
                            acc0 = fmaf(w, pwr[ent.x], acc0);
                            acc1 = fmaf(w, pwr[128 + ent.x], acc1);
                        }
                        macc[m] = acc0;
                        macc[p.n_bins + 1 + m] = acc1;
                    }
                }
            }
            __syncwarp();
            // ---- log, CMVN, store: sub-lane s owns bins s, s + 16, ...
            for (int i = 0; i < n_mel_iter; ++i) {
                const int m = sub + 16 * i;
                if (m < p.n_bins) {
                    float e0;
                    if constexpr (MODE == 0) {
                        e0 = gather_sum(pwr + kFbPwrFloats, reinterpret_cast<const int4*>(s_gather), m, p.n_bins, p.mel_terms);
                    } else {
                        e0 = macc[m];
                    }
                    float v0 = fast_log(fmaxf(e0, 1.1920928955078125e-07f));
                    float v1 = MODE == 1 ? fast_log(fmaxf(macc[p.n_bins + 1 + m], 1.1920928955078125e-07f)) : 0.0f;
                    if constexpr (SUMS) {
                        if (valid[0]) fs.add(i, v0);
                        if (MODE == 1 && valid[MODE]) fs.add(i, v1);
                    }
                    const float2 cm = s_cm[m];
                    v0 = fmaf(v0, cm.x, cm.y);
                    v1 = fmaf(v1, cm.x, cm.y);
                    const long long f0 = f - (MODE + 1);
                    if (valid[0]) p.out[f0 * p.n_bins + m] = v0;
                    if (MODE == 1 && valid[MODE]) p.out[(f0 + 1) * p.n_bins + m] = v1;
                }
            }
            __syncwarp();
        }
        if constexpr (SUMS) fs.fold(s_sums, sub, 16, p.n_bins);
    }
    if constexpr (SUMS) fs.flush(s_sums, p.sums, p.n_bins);
}

template <bool DENORM>
__global__ void __launch_bounds__(256) k_cmvn(long long n, int n_cols, const float* __restrict__ x,
                                               const float* __restrict__ mean, const float* __restrict__ std,
                                               float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % n_cols);
        const float v = x[i];
        // x * std, then + mean: two roundings like the reference (speech_generator_for_s2st.py:27-28)
        out[i] = DENORM ? __fadd_rn(__fmul_rn(v, __ldg(std + c)), __ldg(mean + c)) : (v - __ldg(mean + c)) / __ldg(std + c);
    }
}

// vectorised variant for n_cols % 4 == 0 and 16-byte aligned pointers
template <bool DENORM>
__global__ void __launch_bounds__(256) k_cmvn4(long long n4, int n_cols4, const float4* __restrict__ x,
                                                const float4* __restrict__ mean, const float4* __restrict__ std,
                                                float4* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % n_cols4);
        const float4 v = x[i], m = __ldg(mean + c), s = __ldg(std + c);
        float4 o;
        if (DENORM) {
            o = make_float4(__fadd_rn(__fmul_rn(v.x, s.x), m.x), __fadd_rn(__fmul_rn(v.y, s.y), m.y),
                            __fadd_rn(__fmul_rn(v.z, s.z), m.z), __fadd_rn(__fmul_rn(v.w, s.w), m.w));
        } else {
            o = make_float4((v.x - m.x) / s.x, (v.y - m.y) / s.y, (v.z - m.z) / s.z, (v.w - m.w) / s.w);
        }
        out[i] = o;
    }
}

constexpr int kAccRows = 64;
__global__ void __launch_bounds__(128) k_cmvn_accumulate(long long n_rows, int n_cols, const float* __restrict__ x,
                                                          double* __restrict__ sums) {
    const long long r0 = (long long)blockIdx.x * kAccRows;
    const int nr = (int)min((long long)kAccRows, n_rows - r0);
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
        double s = 0.0, s2 = 0.0;
        for (int r = 0; r < nr; ++r) {
            const double v = (double)x[(r0 + r) * n_cols + c];
            s += v;
            s2 += v * v;
        }
        atomicAdd(sums + c, s);
        atomicAdd(sums + n_cols + c, s2);
    }
}

}  // namespace

int launch_stft(const s2st_plan* plan, int n_utts, long long total_frames, const int64_t* wave_offsets,
                const int32_t* frame_offsets, const float* wave, float* mag_out, float* phase_out,
                float* logmel_out, float eps, const float* cmvn_mean, const float* cmvn_std,
                cudaStream_t stream, double* sums) {
    if (total_frames <= 0) return S2ST_OK;
    if (logmel_out && !plan->mel_ptr) {
        set_error("plan was created without a mel filterbank");
        return S2ST_EINVAL;
    }
    if (plan->generic) {
        StftGenericParams g;
        g.n_fft = plan->n_fft;
        g.n_bins = plan->n_bins;
        g.hop = plan->hop;
        g.n_mels = plan->n_mels;
        g.n_utts = n_utts;
        g.total_frames = total_frames;
        g.win = plan->gwin;
        g.tw = plan->gtw;
        g.wave_offsets = wave_offsets;
        g.frame_offsets = frame_offsets;
        g.wave = wave;
        g.mag_out = mag_out;
        g.phase_out = phase_out;
        g.logmel_out = logmel_out;
        g.eps = eps;
        g.cmvn_mean = cmvn_mean;
        g.cmvn_std = cmvn_std;
        g.sums = sums;
        g.mel_ptr = plan->mel_ptr;
        g.mel_idx = plan->mel_idx;
        g.mel_val = plan->mel_val;
        const size_t per_warp = sizeof(float2) * (size_t)(plan->n_fft + 2);
        const int warps = (int)max((size_t)1, min((size_t)8, (size_t)96 * 1024 / per_warp));
        const size_t gsmem = per_warp * warps;
        const int ggrid = (int)min((long long)plan->num_sms * max(1, (int)(200 * 1024 / gsmem)), (total_frames + warps - 1) / warps);
        if (logmel_out) {
            S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_stft_generic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            k_stft_generic<1><<<ggrid, 32 * warps, gsmem, stream>>>(g);
        } else {
            S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_stft_generic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            k_stft_generic<0><<<ggrid, 32 * warps, gsmem, stream>>>(g);
        }
        S2ST_CUDA_CHECK(cudaGetLastError());
        return S2ST_OK;
    }
    StftParams p;
    p.hop = plan->hop;
    p.half = plan->n_fft / 2;
    p.rot = plan->rot;
    p.ws = plan->ws;
    p.n_mels = plan->n_mels;
    p.n_utts = n_utts;
    p.total_frames = total_frames;
    p.win_a = plan->win_a;
    p.tw = plan->tw;
    p.vtab = plan->vtab;
    p.wave_offsets = wave_offsets;
    p.frame_offsets = frame_offsets;
    p.wave = wave;
    p.mag_out = mag_out;
    p.phase_out = phase_out;
    p.logmel_out = logmel_out;
    p.eps = eps;
    p.cmvn_mean = cmvn_mean;
    p.cmvn_std = cmvn_std;
    p.sums = sums;
    p.mel_ptr = plan->mel_ptr;
    p.mel_idx = plan->mel_idx;
    p.mel_val = plan->mel_val;
    p.mel_col = plan->mel_col;
    p.mel_gather = plan->mel_gather;
    p.mel_terms = plan->mel_terms;
    if (logmel_out && plan->mel_col && plan->mel_gather && !plan->opt_frontend_generic) {
        const size_t fsmem = sizeof(float2) * 2048 + sizeof(float4) * kLmCols + sizeof(int) * 8 * 128 +
                             sizeof(float) * (plan->wp + 8 * kScratchFloats);
        const long long chunks = (total_frames + kLmChunk - 1) / kLmChunk;
        const int fgrid = (int)min((long long)plan->num_sms * 2, (chunks + 7) / 8);
#define S2ST_LAUNCH_LOGMEL(NZV, SUMSV)                                                                                \
    do {                                                                                                             \
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_logmel_fast<NZV, SUMSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)fsmem));                                                           \
        k_logmel_fast<NZV, SUMSV><<<fgrid, 256, fsmem, stream>>>(p);                                                 \
    } while (0)
        if (plan->nz == 19) {
            if (sums) S2ST_LAUNCH_LOGMEL(19, true); else S2ST_LAUNCH_LOGMEL(19, false);
        } else {
            if (sums) S2ST_LAUNCH_LOGMEL(32, true); else S2ST_LAUNCH_LOGMEL(32, false);
        }
#undef S2ST_LAUNCH_LOGMEL
        S2ST_CUDA_CHECK(cudaGetLastError());
        return S2ST_OK;
    }
    const size_t smem = sizeof(float2) * 2048 + sizeof(float) * (8 * kScratchFloats + plan->wp);
    const int grid = (int)min((long long)plan->num_sms * 2, (total_frames + 7) / 8);
#define S2ST_LAUNCH_STFT(NZV, MODEV)                                                                          \
    do {                                                                                                      \
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_stft<NZV, MODEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                     \
        k_stft<NZV, MODEV><<<grid, 256, smem, stream>>>(p);                                                   \
    } while (0)
    if (logmel_out) {
        if (plan->nz == 19) S2ST_LAUNCH_STFT(19, 1); else S2ST_LAUNCH_STFT(32, 1);
    } else {
        if (plan->nz == 19) S2ST_LAUNCH_STFT(19, 0); else S2ST_LAUNCH_STFT(32, 0);
    }
#undef S2ST_LAUNCH_STFT
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_mel_project(const s2st_plan* plan, long long n_frames, const float* spec, float* out,
                       cudaStream_t stream) {
    if (!plan->mel_ptr) {
        set_error("plan was created without a mel filterbank");
        return S2ST_EINVAL;
    }
    if (n_frames <= 0) return S2ST_OK;
    // dense 80 x 1025 contraction -> tensor cores (tcgen05, 3 x TF32); S2ST_OPT_MEL_PROJECT = 1 keeps the SIMT CSR kernel
    if (!plan->opt_mel_simt && mel_project_tc_supported(plan)) return launch_mel_project_tc(plan, n_frames, spec, out, stream);
    const int grid = (int)min((long long)plan->num_sms * 8, (n_frames + 7) / 8);
    k_mel_project<<<grid, 256, 0, stream>>>(n_frames, plan->n_mels, plan->n_bins, spec, plan->mel_ptr, plan->mel_idx,
                                            plan->mel_val, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_fbank(const s2st_fbank_plan* plan, int n_utts, long long total_frames,
                 const int64_t* wave_offsets, const int32_t* frame_offsets, const float* wave,
                 const float* cmvn_mean, const float* cmvn_std, float* out, cudaStream_t stream, double* sums) {
    if (total_frames <= 0) return S2ST_OK;
    FbankParams p;
    p.win = plan->win;
    p.shift = plan->shift;
    p.padded = plan->padded;
    p.n_bins = plan->n_bins;
    p.n_utts = n_utts;
    p.total_frames = total_frames;
    p.window = plan->window;
    p.tw = plan->tw;
    p.wave_offsets = wave_offsets;
    p.frame_offsets = frame_offsets;
    p.wave = wave;
    p.cmvn_mean = cmvn_mean;
    p.cmvn_std = cmvn_std;
    p.mel_ptr = plan->mel_ptr;
    p.mel_idx = plan->mel_idx;
    p.mel_val = plan->mel_val;
    p.out = out;
    p.sums = sums;
    if (plan->fast_mode >= 0 && !plan->opt_generic) {
        FbankFastParams q;
        q.win = plan->win;
        q.shift = plan->shift;
        q.n_bins = plan->n_bins;
        q.n_utts = n_utts;
        q.total_frames = total_frames;
        q.tw16 = plan->tw16;
        q.vsplit = plan->vsplit;
        q.winp = plan->winp;
        q.wave_offsets = wave_offsets;
        q.frame_offsets = frame_offsets;
        q.wave = wave;
        q.cmvn_mean = cmvn_mean;
        q.cmvn_std = cmvn_std;
        q.mel_col = plan->mel_col;
        q.mel_gather = plan->mel_gather;
        q.mel_terms = plan->mel_terms;
        q.mel_zero = plan->mel_zero;
        q.mel_ptr = plan->mel_ptr;
        q.mel_idx = plan->mel_idx;
        q.mel_val = plan->mel_val;
        q.mel_nnz = plan->mel_nnz;
        q.out = out;
        q.sums = sums;
        const size_t acc_floats = (size_t)(((plan->fast_mode + 1) * (plan->n_bins + 1) + 3) & ~3);
        const size_t tab = plan->fast_mode == 0 ? sizeof(float4) * 256 + sizeof(int) * (size_t)((8 * plan->n_bins + 3) & ~3)
                                                : sizeof(int2) * ((plan->mel_nnz + 1) & ~1) + sizeof(int) * ((plan->n_bins + 1 + 3) & ~3);
        const size_t region = plan->fast_mode == 0 ? (size_t)4 * kFbRegion0Floats : (size_t)(kFbScratchBytes + 4 * acc_floats);
        const size_t fsmem = sizeof(float2) * 512 + sizeof(float) * kFbRows * 32 + tab + (size_t)kFbWarps * 2 * region;
        const long long pair_chunks = ((total_frames + kFbChunk - 1) / kFbChunk + 1) / 2;
        const int fgrid = (int)min((long long)plan->num_sms * (plan->fast_mode == 0 ? kFbBlocks0 : 3), (pair_chunks + kFbWarps - 1) / kFbWarps);
#define S2ST_LAUNCH_FBANK(MODEV, SUMSV)                                                                                \
    do {                                                                                                               \
        S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_fbank_fast<MODEV, SUMSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)fsmem));                                                             \
        k_fbank_fast<MODEV, SUMSV><<<fgrid, 32 * kFbWarps, fsmem, stream>>>(q);                                        \
    } while (0)
        if (plan->fast_mode == 0) {
            if (sums) S2ST_LAUNCH_FBANK(0, true); else S2ST_LAUNCH_FBANK(0, false);
        } else {
            if (sums) S2ST_LAUNCH_FBANK(1, true); else S2ST_LAUNCH_FBANK(1, false);
        }
#undef S2ST_LAUNCH_FBANK
        S2ST_CUDA_CHECK(cudaGetLastError());
        return S2ST_OK;
    }
    const size_t smem = sizeof(float2) * (size_t)plan->padded * kFbankWarps;  // 2 * half per warp
    S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_fbank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)200 * 1024 / smem));
    const int grid = (int)min((long long)plan->num_sms * per_sm, (total_frames + kFbankWarps - 1) / kFbankWarps);
    k_fbank<<<grid, 32 * kFbankWarps, smem, stream>>>(p);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_cmvn(long long n_rows, int n_cols, const float* x, const float* mean, const float* std,
                float* out, bool denorm, cudaStream_t stream) {
    const long long n = n_rows * n_cols;
    if (n <= 0) return S2ST_OK;
    const bool vec = (n_cols % 4 == 0) && ((((uintptr_t)x | (uintptr_t)out | (uintptr_t)mean | (uintptr_t)std) & 15) == 0);
    if (vec) {
        const long long n4 = n / 4;
        const int grid = (int)min((long long)148 * 16, (n4 + 255) / 256);
        if (denorm)
            k_cmvn4<true><<<grid, 256, 0, stream>>>(n4, n_cols / 4, (const float4*)x, (const float4*)mean, (const float4*)std, (float4*)out);
        else
            k_cmvn4<false><<<grid, 256, 0, stream>>>(n4, n_cols / 4, (const float4*)x, (const float4*)mean, (const float4*)std, (float4*)out);
    } else {
        const int grid = (int)min((long long)148 * 16, (n + 255) / 256);
        if (denorm)
            k_cmvn<true><<<grid, 256, 0, stream>>>(n, n_cols, x, mean, std, out);
        else
            k_cmvn<false><<<grid, 256, 0, stream>>>(n, n_cols, x, mean, std, out);
    }
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_cmvn_accumulate(long long n_rows, int n_cols, const float* x, double* sums, cudaStream_t stream) {
    if (n_rows <= 0) return S2ST_OK;
    const long long blocks = (n_rows + kAccRows - 1) / kAccRows;
    k_cmvn_accumulate<<<(unsigned)blocks, 128, 0, stream>>>(n_rows, n_cols, x, sums);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

}  // namespace s2st
