// Batched dynamic time warping for the MCD validation metric (sm_100a).
//
// Replaces batch_dynamic_time_warping (examples/s2s_trans/tasks/s2s_translation.py:414-464; the same function lives in
// s2s_translation_mtl.py:364 and fairseq/tasks/text_to_speech.py), which runs the anti-diagonal recurrence as O(M+N)
// rounds of small torch launches and then walks the back pointers on the host with one .item() synchronisation per
// step, and compute_rms_dist (:473-475).  Here one CTA owns one pair of the batch: the recurrence runs diagonal by
// diagonal with a block barrier in between, the back trace follows on the device.
// Arithmetic is the reference's on the CPU: first row / column are sequential cumulative sums (double accumulator,
// float32 results), inner cells float32
// min(left, up-left, up) + distance with the FIRST minimum winning ties (pointer 0 = left, 1 = up-left, 2 = up).
#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {
namespace {

__global__ void __launch_bounds__(256) k_dtw(int m, int n, const float* __restrict__ dist_all,
                                              const long long* __restrict__ shapes, float* __restrict__ cum_all,
                                              int* __restrict__ bp_all, int* __restrict__ path_all) {
    const size_t plane = (size_t)m * n;
    const float* dist = dist_all + plane * blockIdx.x;
    float* cum = cum_all + plane * blockIdx.x;
    int* bp = bp_all + plane * blockIdx.x;
    int* path = path_all + plane * blockIdx.x;
    const int tid = threadIdx.x;
    for (size_t i = tid; i < plane; i += blockDim.x) path[i] = 0;
    // first row and first column: torch.cumsum on the CPU = sequential sum in a double accumulator, every prefix
    // rounded to float32; two warps in parallel
    if (tid == 0) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            acc += (double)dist[j];
            cum[j] = (float)acc;
            bp[j] = 0;
        }
    }
    __syncthreads();  // column 0 is written after row 0 (cell (0, 0) ends up with pointer 2, :430-431)
    if (tid == 32 % blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) {
            acc += (double)dist[(size_t)i * n];
            cum[(size_t)i * n] = (float)acc;
            bp[(size_t)i * n] = 2;
        }
    }
    __syncthreads();
    for (int off = 2; off < m + n - 1; ++off) {
        const int i_lo = max(1, off - n + 1), i_hi = min(m - 1, off - 1);
        for (int i = i_lo + tid; i <= i_hi; i += blockDim.x) {
            const int j = off - i;
            const size_t c = (size_t)i * n + j;
            const float left = cum[c - 1], diag = cum[c - n - 1], up = cum[c - n];
            float v = left;
            int b = 0;
            if (diag < v) {
                v = diag;
                b = 1;
            }
            if (up < v) {
                v = up;
                b = 2;
            }
            bp[c] = b;
            cum[c] = __fadd_rn(v, dist[c]);
        }
        __syncthreads();
    }
    // back trace (:449-460), capped at 10000 path entries like the reference
    if (tid == 0) {
        int i = shapes ? (int)shapes[2 * blockIdx.x] - 1 : m - 1;
        int j = shapes ? (int)shapes[2 * blockIdx.x + 1] - 1 : n - 1;
        if (i >= 0 && j >= 0 && i < m && j < n) {
            int len = 1;
            path[(size_t)i * n + j] = 1;
            while ((i != 0 || j != 0) && len < 10000) {
                const int b = bp[(size_t)i * n + j];
                if (b == 0) --j;
                else if (b == 1) { --i; --j; }
                else --i;
                path[(size_t)i * n + j] = 1;
                ++len;
            }
        }
    }
}

// out[i, j] = sqrt(sum_k (x1[i, k] - x2[j, k])^2 / d)      (compute_rms_dist, s2s_translation.py:467-475)
__global__ void __launch_bounds__(256) k_rms_dist(int m, int n, int d, const float* __restrict__ x1,
                                                   const float* __restrict__ x2, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * n) return;
    const int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
    float acc = 0.0f;
    for (int k = 0; k < d; ++k) {
        const float t = x1[(size_t)i * d + k] - x2[(size_t)j * d + k];
        acc = fmaf(t, t, acc);
    }
    out[idx] = sqrtf(acc / (float)d);
}

// The padded distance batch of batch_compute_distortion (s2s_translation.py:489-505) in ONE launch: pair b owns rows
// off1[b] .. off1[b + 1] of x1 [sum M, d] and off2[b] .. off2[b + 1] of x2 [sum N, d]; out [bsz, max_m, max_n] gets
// the pair's RMS distance matrix in its top-left corner and zeros elsewhere (what F.pad + torch.stack build on the
// host in the reference, one matrix at a time).
__global__ void __launch_bounds__(256) k_rms_dist_batch(int max_m, int max_n, int d, const float* __restrict__ x1,
                                                         const float* __restrict__ x2, const int* __restrict__ off1,
                                                         const int* __restrict__ off2, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int r1 = off1[b], m = off1[b + 1] - r1, r2 = off2[b], n = off2[b + 1] - r2;
    const long long cells = (long long)max_m * max_n;
    float* o = out + (size_t)b * cells;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < cells; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / max_n), j = (int)(idx - (long long)i * max_n);
        float v = 0.0f;
        if (i < m && j < n) {
            float acc = 0.0f;
            for (int k = 0; k < d; ++k) {
                const float t = x1[(size_t)(r1 + i) * d + k] - x2[(size_t)(r2 + j) * d + k];
                acc = fmaf(t, t, acc);
            }
            v = sqrtf(acc / (float)d);
        }
        o[idx] = v;
    }
}

}  // namespace

int launch_rms_dist_batch(int bsz, int max_m, int max_n, int d, const float* x1, const float* x2, const int* off1,
                          const int* off2, float* out, cudaStream_t stream) {
    if (bsz <= 0 || max_m <= 0 || max_n <= 0) return S2ST_OK;
    const long long cells = (long long)max_m * max_n;
    dim3 grid((unsigned)min((cells + 255) / 256, (long long)1024), (unsigned)bsz);
    k_rms_dist_batch<<<grid, 256, 0, stream>>>(max_m, max_n, d, x1, x2, off1, off2, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_dtw(int bsz, int m, int n, const float* dist, const long long* shapes, float* cum, int* bp, int* path,
               cudaStream_t stream) {
    if (bsz <= 0 || m <= 0 || n <= 0) return S2ST_OK;
    k_dtw<<<(unsigned)bsz, 256, 0, stream>>>(m, n, dist, shapes, cum, bp, path);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

int launch_rms_dist(int m, int n, int d, const float* x1, const float* x2, float* out, cudaStream_t stream) {
    if (m <= 0 || n <= 0) return S2ST_OK;
    const long long total = (long long)m * n;
    k_rms_dist<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(m, n, d, x1, x2, out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

}  // namespace s2st
