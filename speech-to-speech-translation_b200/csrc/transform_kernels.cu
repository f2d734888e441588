// Per-utterance feature transforms of the registry (sm_100a): utterance CMVN and SpecAugment mask fill.
//
// Replaces UtteranceCMVN.__call__ (fairseq/data/audio/feature_transforms/utterance_cmvn.py:29-40) and the masking
// part of SpecAugmentTransform.__call__ (specaugment.py:111-131) for ragged batches of [T_i, n_cols] feature
// matrices concatenated row-major.  Both are HBM-bound byte movers (read 4 B + write 4 B per element; the mask
// fill writes only the masked cells).
#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {
namespace {

// Statistics: one thread per (utterance, column).  numpy reduces a C-contiguous [T, n] float32 array over axis 0 row by
// row, in float32 (x.mean(axis=0), (x ** 2).sum(axis=0)), so the reference's statistics are a plain sequential float32
// accumulation per column: the same order and roundings are used here (no FMA contraction), which makes the result
// bit-identical.  Lanes are consecutive columns: every row access of a warp is one coalesced segment; eight rows are
// in flight per thread.  stats[u] = (mean[n_cols], std[n_cols]).
template <bool RAW>  // RAW: stats[u] = (sum[n_cols], sum of squares[n_cols]) -- get_global_cmvn's per-file terms
__global__ void __launch_bounds__(128) k_utterance_stats(const int32_t* __restrict__ fo, int n_cols,
                                                          const float* __restrict__ x, float* __restrict__ stats) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int r0 = fo[blockIdx.x], T = fo[blockIdx.x + 1] - r0;
    if (T <= 0) {
        if (RAW) stats[(size_t)blockIdx.x * 2 * n_cols + c] = stats[(size_t)blockIdx.x * 2 * n_cols + n_cols + c] = 0.0f;
        return;
    }
    const float* px = x + (size_t)r0 * n_cols + c;
    float s = 0.0f, s2 = 0.0f;
    int r = 0;
    for (; r + 8 <= T; r += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(px + (size_t)(r + k) * n_cols);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s = __fadd_rn(s, v[k]);
            s2 = __fadd_rn(s2, __fmul_rn(v[k], v[k]));
        }
    }
    for (; r < T; ++r) {
        const float v = __ldg(px + (size_t)r * n_cols);
        s = __fadd_rn(s, v);
        s2 = __fadd_rn(s2, __fmul_rn(v, v));
    }
    if constexpr (RAW) {
        stats[(size_t)blockIdx.x * 2 * n_cols + c] = s;
        stats[(size_t)blockIdx.x * 2 * n_cols + n_cols + c] = s2;
        return;
    }
    const float n = (float)T;
    const float mean = __fdiv_rn(s, n);
    const float var = __fsub_rn(__fdiv_rn(s2, n), __fmul_rn(mean, mean));
    stats[(size_t)blockIdx.x * 2 * n_cols + c] = mean;
    stats[(size_t)blockIdx.x * 2 * n_cols + n_cols + c] = sqrtf(fmaxf(var, 1e-10f));
}

// Normalisation: a streaming pass over the concatenated rows (16-byte accesses when n_cols is a multiple of 4, which
// the 80-bin features are).  A block owns kApplyRows consecutive rows; a thread walks down its column group and
// follows the utterance boundaries with a running index (one binary search per thread).
constexpr int kApplyRows = 64;
template <int VEC>
__global__ void __launch_bounds__(256) k_utterance_apply(const int32_t* __restrict__ fo, int n_utts, long long n_rows,
                                                          int n_cols, const float* __restrict__ x,
                                                          const float* __restrict__ stats, float* __restrict__ out,
                                                          int norm_means, int norm_vars) {
    const int groups = n_cols / VEC;              // column groups per row
    const int rows_per_pass = blockDim.x / groups;  // rows the block touches at once
    const int rl = threadIdx.x / groups, g = threadIdx.x - rl * groups;
    if (rl >= rows_per_pass) return;
    const long long row0 = (long long)blockIdx.x * kApplyRows;
    long long row = row0 + rl;
    if (row >= n_rows) return;
    int lo = 0, hi = n_utts - 1;  // last u with fo[u] <= row
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (fo[mid] <= row) lo = mid; else hi = mid - 1;
    }
    int u = lo;
    long long next = fo[u + 1];
    float mean[VEC], sd[VEC];
    auto load_stats = [&]() {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            mean[k] = stats[(size_t)u * 2 * n_cols + g * VEC + k];
            sd[k] = stats[(size_t)u * 2 * n_cols + n_cols + g * VEC + k];
        }
    };
    load_stats();
    const long long row_end = min(row0 + kApplyRows, n_rows);
    for (; row < row_end; row += rows_per_pass) {
        if (row >= next) {
            while (u + 1 < n_utts && row >= next) next = fo[++u + 1];
            load_stats();
        }
        const size_t off = (size_t)row * n_cols + g * VEC;
        float v[VEC];
        if constexpr (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(x + off));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = __ldcs(x + off);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            if (norm_means) v[k] = __fsub_rn(v[k], mean[k]);
            if (norm_vars) v[k] = __fdiv_rn(v[k], sd[k]);
        }
        if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(out + off), make_float4(v[0], v[1], v[2], v[3]));
        else __stcs(out + off, v[0]);
    }
}

// sums[u] = sum of all elements of utterance u (double accumulation; block per utterance)
__global__ void __launch_bounds__(256) k_utterance_sum(const int32_t* __restrict__ fo, int n_cols,
                                                        const float* __restrict__ x, double* __restrict__ sums) {
    __shared__ double s_part[8];
    const int r0 = fo[blockIdx.x], T = fo[blockIdx.x + 1] - r0;
    const long long n = (long long)T * n_cols;
    const float* px = x + (size_t)r0 * n_cols;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)px[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];  // fixed order: deterministic
        sums[blockIdx.x] = t;
    }
}

// rect = (row0, row1, col0, col1) in rows of the concatenated matrix; x[row0:row1, col0:col1] = value
__global__ void __launch_bounds__(256) k_fill_rects(const int4* __restrict__ rects, const float* __restrict__ values,
                                                     int n_cols, float* __restrict__ x) {
    const int4 r = rects[blockIdx.x];
    const float v = values[blockIdx.x];
    const int w = r.w - r.z, h = r.y - r.x;
    if (w <= 0 || h <= 0) return;
    const long long n = (long long)w * h;
    for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
        const int rr = (int)(i / w), cc = (int)(i - (long long)rr * w);
        x[(size_t)(r.x + rr) * n_cols + r.z + cc] = v;
    }
}

// SpecAugment time warping (feature_transforms/specaugment.py:96-110): the rows [0, w0) of an utterance are resized to
// w0 + w rows and the rows [w0, T) to T - w0 - w rows with cv2.resize(..., INTER_LINEAR) -- the width stays, so the
// horizontal pass is the identity and every output row is a blend of two source rows (clamped to the part).
//   IPP = false  OpenCV's own float32 resize: position fy = (float)((dy + 0.5) * scale - 0.5) with scale = 1 / (dst_h /
//                src_h) in double, sy = floor(fy), weights (1 - fy, fy) in float, dst = S0 * b0 + S1 * b1 (two rounded
//                products: the baseline-SSE build has no FMA);
//   IPP = true   the x86-64 opencv-python wheels (IPP on): position in double, weight rounded to float,
//                dst = fma(S1 - S0, w, S0).
// warp[u] = (w0, w); w0 < 0: the utterance is copied.  One thread per (row, column); HBM-bound (4 B in + 4 B out, the
// second source row comes from L1/L2).
template <bool IPP>
__global__ void __launch_bounds__(256) k_time_warp(const int32_t* __restrict__ fo, int n_utts, long long n_rows, int n_cols,
                                                    const int2* __restrict__ warp, const float* __restrict__ x,
                                                    float* __restrict__ out) {
    const long long total = n_rows * n_cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / n_cols;
        const int c = (int)(idx - row * n_cols);
        int lo = 0, hi = n_utts - 1;  // last u with fo[u] <= row
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (fo[mid] <= row) lo = mid; else hi = mid - 1;
        }
        const int r0 = fo[lo], T = fo[lo + 1] - r0;
        const int2 wp = warp[lo];
        const int i = (int)(row - r0);
        float v;
        if (wp.x < 0) {
            v = x[idx];
        } else {
            const int w0 = wp.x, w = wp.y;
            int src0, src_h, dst_h, dy;
            if (i < w0 + w) {
                src0 = 0, src_h = w0, dst_h = w0 + w, dy = i;
            } else {
                src0 = w0, src_h = T - w0, dst_h = T - w0 - w, dy = i - (w0 + w);
            }
            int sy;
            float fy;
            if (IPP) {
                const double pos = __dsub_rn(__dmul_rn(dy + 0.5, (double)src_h / (double)dst_h), 0.5);
                const double fl = floor(pos);
                sy = (int)fl;
                fy = (float)(pos - fl);
            } else {
                const double scale = 1.0 / ((double)dst_h / (double)src_h);  // OpenCV: scale_y = 1. / inv_scale_y
                fy = (float)__dsub_rn(__dmul_rn(dy + 0.5, scale), 0.5);
                sy = (int)floorf(fy);
                fy -= (float)sy;
            }
            const int s0 = min(max(sy, 0), src_h - 1), s1 = min(max(sy + 1, 0), src_h - 1);
            const float a = x[(size_t)(r0 + src0 + s0) * n_cols + c], b = x[(size_t)(r0 + src0 + s1) * n_cols + c];
            v = IPP ? fmaf(__fsub_rn(b, a), fy, a) : __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, fy)), __fmul_rn(b, fy));
        }
        out[idx] = v;
    }
}

// Waveform post-processing of generate_waveform.py:115-124 (soundfile writes float data to a WAV / FLAC file as 16-bit
// PCM): sample = lrint(x * 32767) like libsndfile's float -> short conversion, saturated to the int16 range.  Done on
// the device for the whole batch so the download is 2 bytes per sample instead of 4.  HBM-bound: 4 B in + 2 B out.
__global__ void __launch_bounds__(256) k_wave_to_pcm16(long long n, const float* __restrict__ x, short* __restrict__ out) {
    const long long n8 = n >> 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    auto cvt = [](float v) -> unsigned {
        const float s = v == v ? fminf(fmaxf(v * 32767.0f, -32768.0f), 32767.0f) : 0.0f;  // NaN -> 0
        return (unsigned)(unsigned short)(short)__float2int_rn(s);
    };
    if (vec) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i + 1);
            uint4 o;
            o.x = cvt(a.x) | (cvt(a.y) << 16);
            o.y = cvt(a.z) | (cvt(a.w) << 16);
            o.z = cvt(b.x) | (cvt(b.y) << 16);
            o.w = cvt(b.z) | (cvt(b.w) << 16);
            __stcs(reinterpret_cast<uint4*>(out) + i, o);
        }
    }
    const long long tail0 = vec ? n8 << 3 : 0;
    for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (short)cvt(x[i]);
}

// The other direction, for the front-end's input (audio_utils.py:65-109: soundfile reads 16-bit PCM files as float32,
// value / 32768 when normalised, the int16 value itself for the Kaldi-style fbank of get_fbank): samples cross PCIe as
// the 2 bytes they occupy on disk and become float32 here, wave[i] = (float)pcm[i] * scale (exact).  HBM-bound: 2 B in +
// 4 B out, 16-byte loads / stores.
__global__ void __launch_bounds__(256) k_pcm16_to_wave(long long n, const short* __restrict__ pcm, float scale, float* __restrict__ out) {
    const long long n8 = n >> 3;
    const bool vec = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(pcm) + i);
            auto lo = [](unsigned w) { return (float)(short)(w & 0xFFFFu); };
            auto hi = [](unsigned w) { return (float)(short)(w >> 16); };
            __stcs(reinterpret_cast<float4*>(out) + 2 * i, make_float4(lo(v.x) * scale, hi(v.x) * scale, lo(v.y) * scale, hi(v.y) * scale));
            __stcs(reinterpret_cast<float4*>(out) + 2 * i + 1, make_float4(lo(v.z) * scale, hi(v.z) * scale, lo(v.w) * scale, hi(v.w) * scale));
        }
    }
    const long long tail0 = vec ? n8 << 3 : 0;
    for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)pcm[i] * scale;
}

}  // namespace

int launch_pcm16_to_wave(long long n, const short* pcm, float scale, float* out, cudaStream_t stream) {
    if (n <= 0) return S2ST_OK;
    const int grid = (int)min((long long)148 * 8, (n / 8 + 255) / 256 + 1);
    k_pcm16_to_wave<<<grid, 256, 0, stream>>>(n, pcm, scale, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_time_warp(int n_utts, long long n_rows, const int32_t* fo, int n_cols, const int32_t* warp, int arithmetic,
                     const float* x, float* out, cudaStream_t stream) {
    if (n_utts <= 0 || n_rows <= 0) return S2ST_OK;
    const long long total = n_rows * n_cols;
    const int grid = (int)min((long long)148 * 8, (total + 255) / 256);
    if (arithmetic)
        k_time_warp<true><<<grid, 256, 0, stream>>>(fo, n_utts, n_rows, n_cols, reinterpret_cast<const int2*>(warp), x, out);
    else
        k_time_warp<false><<<grid, 256, 0, stream>>>(fo, n_utts, n_rows, n_cols, reinterpret_cast<const int2*>(warp), x, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_wave_to_pcm16(long long n, const float* x, short* out, cudaStream_t stream) {
    if (n <= 0) return S2ST_OK;
    const int grid = (int)min((long long)148 * 8, (n / 8 + 255) / 256 + 1);
    k_wave_to_pcm16<<<grid, 256, 0, stream>>>(n, x, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_utterance_cmvn(int n_utts, long long n_rows, const int32_t* fo, int n_cols, const float* x, float* out,
                          bool norm_means, bool norm_vars, float* stats, cudaStream_t stream) {
    if (n_utts <= 0 || n_rows <= 0) return S2ST_OK;
    dim3 grid((unsigned)n_utts, (unsigned)((n_cols + 127) / 128));
    k_utterance_stats<false><<<grid, 128, 0, stream>>>(fo, n_cols, x, stats);
    S2ST_CUDA_CHECK(cudaGetLastError());
    const unsigned blocks = (unsigned)((n_rows + kApplyRows - 1) / kApplyRows);
    const bool vec = n_cols % 4 == 0 && n_cols / 4 <= 256 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
    if (vec) {
        k_utterance_apply<4><<<blocks, 256, 0, stream>>>(fo, n_utts, n_rows, n_cols, x, stats, out, norm_means ? 1 : 0, norm_vars ? 1 : 0);
    } else if (n_cols <= 256) {
        k_utterance_apply<1><<<blocks, 256, 0, stream>>>(fo, n_utts, n_rows, n_cols, x, stats, out, norm_means ? 1 : 0, norm_vars ? 1 : 0);
    } else {
        set_error("s2st_utterance_cmvn supports at most 256 feature columns (got %d)", n_cols);
        return S2ST_EINVAL;
    }
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_utterance_sums(int n_utts, const int32_t* fo, int n_cols, const float* x, float* sums, cudaStream_t stream) {
    if (n_utts <= 0) return S2ST_OK;
    dim3 grid((unsigned)n_utts, (unsigned)((n_cols + 127) / 128));
    k_utterance_stats<true><<<grid, 128, 0, stream>>>(fo, n_cols, x, sums);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_utterance_sum(int n_utts, const int32_t* fo, int n_cols, const float* x, double* sums, cudaStream_t stream) {
    if (n_utts <= 0) return S2ST_OK;
    k_utterance_sum<<<(unsigned)n_utts, 256, 0, stream>>>(fo, n_cols, x, sums);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_fill_rects(int n_rects, const int32_t* rects, const float* values, int n_cols, float* x, cudaStream_t stream) {
    if (n_rects <= 0) return S2ST_OK;
    dim3 grid((unsigned)n_rects, 8);
    k_fill_rects<<<grid, 256, 0, stream>>>(reinterpret_cast<const int4*>(rects), values, n_cols, x);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

}  // namespace s2st
