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

#include "../../include/s2st_b200.h"
#include "frame_fft.cuh"
#include "plan.h"

namespace s2st {

namespace {

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
    const int* mel_ptr;
    const int* mel_idx;
    const float* mel_val;
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
                if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + m)) / __ldg(p.cmvn_std + m);
                p.logmel_out[f * p.n_mels + m] = v;
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(256) k_mel_project(long long n_frames, int n_mels, const float* __restrict__ spec,
                                                      const int* __restrict__ mel_ptr, const int* __restrict__ mel_idx,
                                                      const float* __restrict__ mel_val, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long f = (long long)blockIdx.x * 8 + warp; f < n_frames; f += (long long)gridDim.x * 8) {
        const float* row = spec + f * kBins;
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
            if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + m)) / __ldg(p.cmvn_std + m);
            p.out[f * p.n_bins + m] = v;
        }
        __syncwarp();
    }
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
                cudaStream_t stream) {
    if (total_frames <= 0) return S2ST_OK;
    if (logmel_out && !plan->mel_ptr) {
        set_error("plan was created without a mel filterbank");
        return S2ST_EINVAL;
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
    p.mel_ptr = plan->mel_ptr;
    p.mel_idx = plan->mel_idx;
    p.mel_val = plan->mel_val;
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
    const int grid = (int)min((long long)plan->num_sms * 8, (n_frames + 7) / 8);
    k_mel_project<<<grid, 256, 0, stream>>>(n_frames, plan->n_mels, spec, plan->mel_ptr, plan->mel_idx,
                                            plan->mel_val, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_fbank(const s2st_fbank_plan* plan, int n_utts, long long total_frames,
                 const int64_t* wave_offsets, const int32_t* frame_offsets, const float* wave,
                 const float* cmvn_mean, const float* cmvn_std, float* out, cudaStream_t stream) {
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
