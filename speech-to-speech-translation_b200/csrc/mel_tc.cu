// Inverse-mel projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   mag[t, f] = max(0, sum_m pinv[f, m] * g(mel[t, m]))      g = exp (vocoder.py:141) or identity (vocoder.py:42)
//
// is the one dense contraction of the synthesis path: D[M = frames, N = live bins] = A[M, K = n_mels] * B[N, K]^T.
// The reference runs it as an fp32 matmul and the features must stay within ~1e-6, so single-pass TF32 (10-bit
// mantissa) is not enough: operands are split into a TF32 head and tail (a = a_hi + a_lo, |a_lo| <= 2^-11 |a|)
// and three MMAs are accumulated in fp32 in TMEM:  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (the dropped tail*tail
// term is ~2^-22 relative).
//
// One CTA = 128 frames (UMMA M = 128, cta_group::1), 128 threads, software-pipelined over 8 chunks of 96 bins:
//   - the pre-split, pre-laid-out B chunks (built once by the plan, L2 resident) arrive by BULK ASYNC COPY
//     (cp.async.bulk -> mbarrier complete_tx; SASS UBLKCP) into two alternating shared-memory buffers; the first two
//     are in flight while all threads do exp + hi/lo split of the A tile into the canonical K-major, no-swizzle UMMA
//     layout (8-row x 16-byte core matrices; SBO = 128 B between 8-row groups, LBO between 4-element K chunks);
//   - one elected thread issues the 3 x (K / 8) tcgen05.mma kind::tf32 of chunk c + 1 into the OTHER of two 96-column
//     fp32 accumulators in TMEM before the CTA reads chunk c back, so the tensor pipe works under the epilogue;
//   - epilogue: every thread reads its own row (TMEM lane) with tcgen05.ld, clamps at zero and stores 16 bytes at a
//     time; when the chunk's MMAs have retired its B buffer is refilled with chunk c + 2.
#include <stdint.h>

#include <cstring>

#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {

namespace {

constexpr int kTcM = 128;        // frames per CTA (UMMA M)
constexpr int kTcNChunk = 96;    // bins per MMA (UMMA N, multiple of 16)
constexpr int kTcChunks = 8;     // 8 * 96 = 768 >= 704 live bins (rows beyond the basis are zero)
constexpr int kTcTmemCols = 256; // power of two >= 2 accumulators of kTcNChunk columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SM100 shared-memory matrix descriptor: start address [0,14), LBO [16,30), SBO [32,46) (all >> 4),
    // version 1 at [46,48), layout type 0 (no swizzle) at [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// canonical K-major / no-swizzle offset (in floats) of element (row, k) of a tile with `groups` 8-row groups
__host__ __device__ __forceinline__ int canon_off(int row, int k, int groups) {
    return (((k >> 2) * groups + (row >> 3)) * 8 + (row & 7)) * 4 + (k & 3);
}

struct TcParams {
    const float* mel;       // [n_frames, K]
    const float* b_tc;      // [kTcChunks][2 (hi, lo)][K/4][kTcNChunk/8][8][4]
    float* mag;             // [n_frames, out_stride]
    long long n_frames;
    int K, is_log, out_stride, n_out;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) k_inverse_mel_tc(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int K = p.K, kc = K >> 2;                      // K chunks of 4 elements (16 bytes)
    const int a_floats = kTcM * K, b_floats = kTcNChunk * K;
    float* sA_hi = reinterpret_cast<float*>(smem_raw);
    float* sA_lo = sA_hi + a_floats;
    float* sB[2] = {sA_lo + a_floats, sA_lo + a_floats + 2 * b_floats};  // each: head then tail of one chunk
    __shared__ __align__(8) uint64_t s_full[2], s_done[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    const long long t0 = (long long)blockIdx.x * kTcM;
    const int rows = (int)min((long long)kTcM, p.n_frames - t0);
    const uint32_t chunk_bytes = 2u * (uint32_t)b_floats * 4u;
    // gridDim.y > 1 (small calls): the CTAs of one frame tile share its bin chunks, [cb, ce) each -- a call of a few
    // hundred frames is bound by the latency of one CTA walking all eight chunks (41 us), not by throughput
    const int per = (kTcChunks + (int)gridDim.y - 1) / (int)gridDim.y;
    const int cb = (int)blockIdx.y * per, ce = min(kTcChunks, cb + per);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_done[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first two B chunks fly while the CTA prepares A
        for (int i = 0; i < 2; ++i)
            if (cb + i < ce) bulk_load(sB[i], p.b_tc + (size_t)(cb + i) * 2 * b_floats, chunk_bytes, smem_u32(&s_full[i]));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // A tile: g(mel) split into TF32 head / tail, canonical layout.  The [rows, K] block is contiguous in global.
    const float4* src = reinterpret_cast<const float4*>(p.mel + t0 * K);
    for (int i = tid; i < kTcM * kc; i += 128) {
        const int row = i / kc, c = i - row * kc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
            v = src[i];
            if (p.is_log) v = make_float4(expf(v.x), expf(v.y), expf(v.z), expf(v.w));
        }
        const float4 hi = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        const int off = canon_off(row, 4 * c, kTcM / 8);
        *reinterpret_cast<float4*>(sA_hi + off) = hi;
        *reinterpret_cast<float4*>(sA_lo + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of A -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcNChunk >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
    const uint32_t lbo_a = (kTcM / 8) * 128, lbo_b = (kTcNChunk / 8) * 128, sbo = 128;
    const bool vec_store = (p.out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mag) & 15) == 0;

    // the CTA's chunk number lc = c - cb: B buffer lc & 1, accumulator columns (lc & 1) * kTcNChunk; mbarrier phase (lc >> 1) & 1
    auto issue_mma = [&](int c) {  // thread 0 only
        const int buf = (c - cb) & 1;
        mbar_wait(smem_u32(&s_full[buf]), (uint32_t)((c - cb) >> 1) & 1);  // the chunk's B has landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + (uint32_t)(buf * kTcNChunk);
        uint32_t acc = 0;
#pragma unroll 1
        for (int term = 0; term < 3; ++term) {
            const uint32_t a_base = smem_u32(term == 2 ? sA_lo : sA_hi);
            const uint32_t b_base = smem_u32(term == 1 ? sB[buf] + b_floats : sB[buf]);
            for (int ks = 0; ks < K / 8; ++ks) {
                mma_tf32(d, make_desc(a_base + ks * 2 * lbo_a, lbo_a, sbo), make_desc(b_base + ks * 2 * lbo_b, lbo_b, sbo), idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_done[buf])) : "memory");
    };
    if (tid == 0 && cb < ce) issue_mma(cb);
    for (int chunk = cb; chunk < ce; ++chunk) {
        const int buf = (chunk - cb) & 1;
        // the next chunk's MMAs go to the other accumulator and run under this chunk's epilogue
        if (tid == 0 && chunk + 1 < ce) issue_mma(chunk + 1);
        mbar_wait(smem_u32(&s_done[buf]), (uint32_t)((chunk - cb) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // this chunk's MMAs have retired: its B buffer is free for chunk + 2
        if (tid == 0 && chunk + 2 < ce)
            bulk_load(sB[buf], p.b_tc + (size_t)(chunk + 2) * 2 * b_floats, chunk_bytes, smem_u32(&s_full[buf]));
        // epilogue: thread tid owns row tid (TMEM lane tid); 16 columns per tcgen05.ld
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * kTcNChunk);
        float* orow = p.mag + (t0 + tid) * (long long)p.out_stride + chunk * kTcNChunk;
        for (int c0 = 0; c0 < kTcNChunk; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(lane_addr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int col0 = chunk * kTcNChunk + c0;
            if (tid < rows && col0 < p.n_out) {
                if (vec_store && col0 + 16 <= p.n_out) {
                    // the thread's 16 consecutive bins as four 16-byte stores: scalar stores from 32 different rows per
                    // instruction made this epilogue the kernel's bottleneck
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(orow + c0 + 4 * q) =
                            make_float4(fmaxf(__uint_as_float(r[4 * q]), 0.0f), fmaxf(__uint_as_float(r[4 * q + 1]), 0.0f),
                                        fmaxf(__uint_as_float(r[4 * q + 2]), 0.0f), fmaxf(__uint_as_float(r[4 * q + 3]), 0.0f));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (col0 + j < p.n_out) orow[c0 + j] = fmaxf(__uint_as_float(r[j]), 0.0f);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every thread has read this accumulator: chunk + 2 may overwrite it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // columns beyond the chunks are exactly zero
    if (tid < rows && blockIdx.y == gridDim.y - 1)
        for (int col = kTcChunks * kTcNChunk; col < p.n_out; ++col) p.mag[(t0 + tid) * (long long)p.out_stride + col] = 0.0f;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward mel projection (TTSMelScale.forward, audio_utils.py:284-285) on the tensor cores:
//   mel_out[t, m] = sum_f mel[m, f] * spec[t, f]      D[M = 128 frames, N = n_mels] = A[M, K = bins] * B[N, K]^T
// the other dense contraction of the path (80 x 1025).  Same 3 x TF32 split as above.  K is walked in chunks of 64 bins:
// the CTA stages the chunk of its 128 spectrogram rows (coalesced 128-byte row segments from global, hi / lo split,
// canonical K-major layout) and the pre-split chunk of the filterbank (built once by the plan, L2 resident), one
// elected thread issues 3 x 8 tcgen05.mma into ONE accumulator that stays in TMEM across all chunks, and after the
// last chunk every thread reads its frame's n_mels sums back with tcgen05.ld.  104 KB of shared memory per CTA: two
// CTAs per SM overlap one's loads with the other's MMAs.  Bins beyond the last non-zero filterbank column are skipped.
constexpr int kMpKChunk = 64;
constexpr int kMpMaxN = 128;

struct MelProjParams {
    const float* spec;      // [n_frames, n_bins]
    const float* b_tc;      // [k_chunks][2 (hi, lo)][kMpKChunk / 4][n_pad / 8][8][4]
    float* out;             // [n_frames, n_mels]
    long long n_frames;
    int n_bins, n_mels, n_pad, k_chunks;
};

__global__ void __launch_bounds__(128, 2) k_mel_project_tc(const __grid_constant__ MelProjParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int a_floats = kTcM * kMpKChunk;
    const int b_floats = p.n_pad * kMpKChunk;
    float* sA_hi = reinterpret_cast<float*>(smem_raw);
    float* sA_lo = sA_hi + a_floats;
    float* sB_hi = sA_lo + a_floats;
    float* sB_lo = sB_hi + b_floats;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = (long long)blockIdx.x * kTcM;
    const int rows = (int)min((long long)kTcM, p.n_frames - t0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kMpMaxN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
    const uint32_t lbo_a = (kTcM / 8) * 128, lbo_b = (uint32_t)(p.n_pad / 8) * 128, sbo = 128;
    uint32_t phase = 0, acc = 0;
    for (int kc = 0; kc < p.k_chunks; ++kc) {
        const int k0 = kc * kMpKChunk;
        // A chunk: warp w stages rows w, w + 4, ...; a row segment is 64 consecutive floats (rows are 4-byte aligned only)
        for (int row = warp; row < kTcM; row += 4) {
            const float* src = p.spec + (t0 + row) * (long long)p.n_bins + k0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 32 * h + lane;
                const float v = (row < rows && k0 + k < p.n_bins) ? __ldg(src + k) : 0.0f;
                const float hi = tf32_round(v);
                const int off = canon_off(row, k, kTcM / 8);
                sA_hi[off] = hi;
                sA_lo[off] = v - hi;
            }
        }
        const float4* bsrc = reinterpret_cast<const float4*>(p.b_tc + (size_t)kc * 2 * b_floats);
        float4* bdst = reinterpret_cast<float4*>(sB_hi);
        for (int i = tid; i < 2 * b_floats / 4; i += 128) bdst[i] = __ldg(bsrc + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int term = 0; term < 3; ++term) {
                const uint32_t a_base = smem_u32(term == 2 ? sA_lo : sA_hi);
                const uint32_t b_base = smem_u32(term == 1 ? sB_lo : sB_hi);
                for (int ks = 0; ks < kMpKChunk / 8; ++ks) {
                    mma_tf32(tmem, make_desc(a_base + ks * 2 * lbo_a, lbo_a, sbo), make_desc(b_base + ks * 2 * lbo_b, lbo_b, sbo), idesc, acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        mbar_wait(smem_u32(&s_bar), phase);  // the MMAs have read this chunk: its buffers may be refilled
        phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float* orow = p.out + (t0 + tid) * (long long)p.n_mels;
    for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(lane_addr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (tid < rows) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < p.n_mels) orow[c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kMpMaxN));
}

}  // namespace

// Host: the mel filterbank [n_mels, n_bins] pre-split into TF32 head / tail, one canonical K-major tile of n_pad rows x
// 64 bins per K chunk.  Returns the number of chunks that hold non-zero weights (bins beyond them are never read).
int mel_project_tc_chunks(const float* mel, int n_mels, int n_bins) {
    int last = -1;
    for (int m = 0; m < n_mels; ++m)
        for (int k = 0; k < n_bins; ++k)
            if (mel[(size_t)m * n_bins + k] != 0.0f && k > last) last = k;
    return last < 0 ? 0 : last / kMpKChunk + 1;
}
int mel_project_tc_npad(int n_mels) { return (n_mels + 15) & ~15; }
size_t mel_project_tc_floats(int n_mels, int k_chunks) { return (size_t)k_chunks * 2 * mel_project_tc_npad(n_mels) * kMpKChunk; }
void build_mel_project_tc(const float* mel, int n_mels, int n_bins, int k_chunks, float* out) {
    const int n_pad = mel_project_tc_npad(n_mels), b_floats = n_pad * kMpKChunk;
    for (size_t i = 0; i < mel_project_tc_floats(n_mels, k_chunks); ++i) out[i] = 0.0f;
    for (int m = 0; m < n_mels; ++m)
        for (int k = 0; k < n_bins && k < k_chunks * kMpKChunk; ++k) {
            const float v = mel[(size_t)m * n_bins + k];
            uint32_t u;
            memcpy(&u, &v, 4);
            uint32_t h = (u + 0x1000u) & 0xFFFFE000u;  // cvt.rna.tf32.f32
            if ((u & 0x7F800000u) == 0x7F800000u) h = u;
            float hi;
            memcpy(&hi, &h, 4);
            const int chunk = k / kMpKChunk, off = canon_off(m, k % kMpKChunk, n_pad / 8);
            out[(size_t)(chunk * 2 + 0) * b_floats + off] = hi;
            out[(size_t)(chunk * 2 + 1) * b_floats + off] = v - hi;
        }
}

bool mel_project_tc_supported(const s2st_plan* plan) {
    return plan->mel_tc != nullptr && plan->mel_tc_chunks > 0 && plan->n_mels <= kMpMaxN;
}

int launch_mel_project_tc(const s2st_plan* plan, long long n_frames, const float* spec, float* out, cudaStream_t stream) {
    if (n_frames <= 0) return S2ST_OK;
    MelProjParams p;
    p.spec = spec;
    p.b_tc = plan->mel_tc;
    p.out = out;
    p.n_frames = n_frames;
    p.n_bins = plan->n_bins;
    p.n_mels = plan->n_mels;
    p.n_pad = mel_project_tc_npad(plan->n_mels);
    p.k_chunks = plan->mel_tc_chunks;
    const size_t smem = sizeof(float) * (size_t)(2 * kTcM * kMpKChunk + 2 * p.n_pad * kMpKChunk) + 1024;
    S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_mel_project_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (n_frames + kTcM - 1) / kTcM;
    k_mel_project_tc<<<(unsigned)blocks, 128, smem, stream>>>(p);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

// Host: pre-split the pseudo-inverse basis into TF32 head / tail in the canonical per-chunk layout.
// inv_mel [n_bins_total, K] row-major (K-major), rows >= kb are zero.
void build_inverse_mel_tc(const float* inv_mel, int kb, int K, float* out /* kTcChunks*2*kTcNChunk*K */) {
    const int b_floats = kTcNChunk * K;
    for (int i = 0; i < kTcChunks * 2 * b_floats; ++i) out[i] = 0.0f;
    for (int n = 0; n < kb && n < kTcChunks * kTcNChunk; ++n) {
        const int chunk = n / kTcNChunk, row = n % kTcNChunk;
        for (int k = 0; k < K; ++k) {
            const float v = inv_mel[(size_t)n * K + k];
            uint32_t u;
            memcpy(&u, &v, 4);
            // round to nearest, ties away (cvt.rna.tf32.f32): add half an ulp of the 13 dropped bits, truncate
            uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
            if ((u & 0x7F800000u) == 0x7F800000u) h = u;  // inf / nan untouched
            float hi;
            memcpy(&hi, &h, 4);
            const float lo = v - hi;
            const int off = canon_off(row, k, kTcNChunk / 8);
            out[(size_t)(chunk * 2 + 0) * b_floats + off] = hi;
            out[(size_t)(chunk * 2 + 1) * b_floats + off] = lo;
        }
    }
}

size_t inverse_mel_tc_floats(int K) { return (size_t)kTcChunks * 2 * kTcNChunk * K; }

bool inverse_mel_tc_supported(const s2st_plan* plan) {
    return plan->inv_mel_tc != nullptr && plan->kb <= kTcChunks * kTcNChunk && plan->n_mels % 8 == 0 && plan->n_mels <= 80;
}

int launch_inverse_mel_tc(const s2st_plan* plan, long long n_frames, const float* mel, bool is_log, float* mag,
                          int out_stride, int n_out, cudaStream_t stream, const float* basis_tc) {
    if (n_frames <= 0) return S2ST_OK;
    TcParams p;
    p.mel = mel;
    p.b_tc = basis_tc ? basis_tc : plan->inv_mel_tc;
    p.mag = mag;
    p.n_frames = n_frames;
    p.K = plan->n_mels;
    p.is_log = is_log ? 1 : 0;
    p.out_stride = out_stride;
    p.n_out = n_out;
    const size_t smem = sizeof(float) * (size_t)(2 * kTcM * p.K + 4 * kTcNChunk * p.K) + 1024;  // A head + tail, two B chunk buffers
    S2ST_CUDA_CHECK(cudaFuncSetAttribute(k_inverse_mel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long blocks = (n_frames + kTcM - 1) / kTcM;
    // small calls: split the eight bin chunks of a frame tile over up to eight CTAs (as long as the grid fits one wave)
    const int split = blocks * 8 <= plan->num_sms ? 8 : blocks * 4 <= plan->num_sms ? 4 : blocks * 2 <= plan->num_sms ? 2 : 1;
    k_inverse_mel_tc<<<dim3((unsigned)blocks, (unsigned)split), 128, smem, stream>>>(p);
    S2ST_CUDA_CHECK(cudaGetLastError());
    return S2ST_OK;
}

}  // namespace s2st
