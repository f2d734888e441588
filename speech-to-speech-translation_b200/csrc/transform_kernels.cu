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

// One thread per (utterance, column).  numpy reduces a C-contiguous [T, n] float32 array over axis 0 row by row, in
// float32 (x.mean(axis=0), (x ** 2).sum(axis=0)), so the reference's statistics are a plain sequential float32
// accumulation per column: the same order and roundings are used here (no FMA contraction), which makes the result
// bit-identical.  Lanes are consecutive columns: every row access of a warp is one coalesced segment.
__global__ void __launch_bounds__(128) k_utterance_cmvn(const int32_t* __restrict__ fo, int n_cols,
                                                         const float* __restrict__ x, float* __restrict__ out,
                                                         int norm_means, int norm_vars) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int r0 = fo[blockIdx.x], T = fo[blockIdx.x + 1] - r0;
    if (T <= 0) return;
    const float* px = x + (size_t)r0 * n_cols + c;
    float* po = out + (size_t)r0 * n_cols + c;
    float s = 0.0f, s2 = 0.0f;
#pragma unroll 8
    for (int r = 0; r < T; ++r) {
        const float v = px[(size_t)r * n_cols];
        s = __fadd_rn(s, v);
        s2 = __fadd_rn(s2, __fmul_rn(v, v));
    }
    const float n = (float)T;
    const float mean = __fdiv_rn(s, n);
    const float var = __fsub_rn(__fdiv_rn(s2, n), __fmul_rn(mean, mean));
    const float sd = sqrtf(fmaxf(var, 1e-10f));
#pragma unroll 8
    for (int r = 0; r < T; ++r) {
        float v = px[(size_t)r * n_cols];
        if (norm_means) v = __fsub_rn(v, mean);
        if (norm_vars) v = __fdiv_rn(v, sd);
        po[(size_t)r * n_cols] = v;
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

}  // namespace

int launch_utterance_cmvn(int n_utts, const int32_t* fo, int n_cols, const float* x, float* out, bool norm_means,
                          bool norm_vars, cudaStream_t stream) {
    if (n_utts <= 0) return S2ST_OK;
    dim3 grid((unsigned)n_utts, (unsigned)((n_cols + 127) / 128));
    k_utterance_cmvn<<<grid, 128, 0, stream>>>(fo, n_cols, x, out, norm_means ? 1 : 0, norm_vars ? 1 : 0);
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
