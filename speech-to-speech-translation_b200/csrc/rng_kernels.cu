// numpy's legacy global generator continued on the device (sm_100a).
//
// GriffinLim.forward draws its initial phase with np.random.rand(*specgram.shape) (fairseq/models/text_to_speech/
// vocoder.py:103): 1025 x T float64 uniforms from numpy's GLOBAL RandomState -- seeding numpy is how a user reproduces
// the reference.  On the host that draw costs more than the whole synthesis of the utterance (config 1: ~1.2 ms for
// 512 500 doubles plus a 4 MB upload against 0.45 ms of kernels), so the host shim hands the generator STATE
// (624 words + position) to the device instead: k_mt19937_advance produces exactly the next 2 n outputs and the state
// numpy would be left in, k_phase_from_words turns the outputs into the frame-major float32 phase, and the shim puts
// the new state back with np.random.set_state.  Bit-for-bit numpy's stream (tests: against numpy itself).
//
// MT19937 as ONE untempered sequence: X[0:624] = key, X[n] = X[n-227] ^ twist(X[n-624], X[n-623]); the outputs are
// temper(X[pos]), temper(X[pos+1]), ...  The closest dependency lies 227 back and is the SAME thread's previous
// element when 227 threads walk the sequence in steps of 227 (kept in a register); the other two operands are at
// least 397 back, i.e. at least one step old after TWO steps -- so a block barrier is needed only every 454 elements.
// The recurrence is sequential by nature (one block, latency bound: ~50 ns per 454 words); everything else
// (tempering, the conversion to doubles, the phase) is done by the second, fully parallel kernel.
#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {

namespace {

constexpr int kMtN = 624, kMtM = 227, kMtRing = 2048;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7FFFFFFFu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    return y ^ (y >> 18);
}

// words_out[k] = X[pos + k] for k < n_words (untempered); key_out = the block numpy's state holds afterwards.
// One barrier interval = 454 elements: thread tid < 227 computes X[base + tid] and X[base + 227 + tid].  Its operand
// X[n - 227] is its own previous result (a register); the four other operands are at least 397 back, i.e. written
// before the last barrier -- their loads are issued together at the top of the interval.
__global__ void __launch_bounds__(256) k_mt19937_advance(const uint32_t* __restrict__ key_in, int pos, long long n_words,
                                                          uint32_t* __restrict__ words_out, uint32_t* __restrict__ key_out) {
    __shared__ uint32_t ring[kMtRing];  // X[n] at ring[n mod 2048]: an interval reads 624 back and writes 454 ahead
    const int tid = threadIdx.x;
    const long long end = pos + n_words;
    const long long q = end > 0 ? (end - 1) / kMtN : 0;
    const long long n_end = kMtN * (q + 1);
    for (int i = tid; i < kMtN; i += 256) {
        const uint32_t v = key_in[i];
        ring[i] = v;
        if (i >= pos && i < end) words_out[i - pos] = v;
    }
    __syncthreads();
    constexpr unsigned kMask = kMtRing - 1;
    // full intervals: every element exists (< n_end); almost all of them are outputs too
    const long long n_full = (n_end - kMtN) / (2 * kMtM);
    {
        const bool worker = tid < kMtM;
        const int wt = worker ? tid : 0;                        // (idle threads shadow thread 0 without storing)
        unsigned r0 = (unsigned)(kMtN + wt) & kMask;            // ring slot of X[base + tid]
        uint32_t prev = ring[(kMtN + wt - kMtM) & kMask];       // X[base + tid - 227]
        uint32_t* out = words_out + ((long long)kMtN + wt - pos);  // &words_out[n0 - pos] (only dereferenced when valid)
        // every interval starts at or after element 624 >= pos, so only "n < end" can exclude an element from the
        // outputs, and only in the last block: the first n_body intervals store unconditionally
        const long long n_body = end > kMtN ? min(n_full, (end - kMtN) / (2 * kMtM)) : 0;
        long long it = 0;
        for (; it < n_body; ++it) {
            const unsigned r1 = (r0 + kMtM) & kMask;
            const uint32_t a0 = ring[(r0 - kMtN) & kMask], a1 = ring[(r0 - kMtN + 1) & kMask];
            const uint32_t b0 = ring[(r1 - kMtN) & kMask], b1 = ring[(r1 - kMtN + 1) & kMask];
            const uint32_t x0 = prev ^ mt_twist(a0, a1);
            const uint32_t x1 = x0 ^ mt_twist(b0, b1);
            if (worker) {
                ring[r0] = x0;
                ring[r1] = x1;
                out[0] = x0;
                out[kMtM] = x1;
            }
            prev = x1;
            r0 = (r0 + 2 * kMtM) & kMask;
            out += 2 * kMtM;
            __syncthreads();
        }
        long long n0 = kMtN + wt + it * 2 * kMtM;
        for (; it < n_full; ++it) {
            const unsigned r1 = (r0 + kMtM) & kMask;
            const uint32_t a0 = ring[(r0 - kMtN) & kMask], a1 = ring[(r0 - kMtN + 1) & kMask];
            const uint32_t b0 = ring[(r1 - kMtN) & kMask], b1 = ring[(r1 - kMtN + 1) & kMask];
            const uint32_t x0 = prev ^ mt_twist(a0, a1);
            const uint32_t x1 = x0 ^ mt_twist(b0, b1);
            if (worker) {
                ring[r0] = x0;
                ring[r1] = x1;
                if (n0 < end) out[0] = x0;
                if (n0 + kMtM < end) out[kMtM] = x1;
            }
            prev = x1;
            r0 = (r0 + 2 * kMtM) & kMask;
            out += 2 * kMtM;
            n0 += 2 * kMtM;
            __syncthreads();
        }
    }
    // the remaining elements of the last block (fewer than 454), by the plain two-step scheme
    for (long long base = kMtN + n_full * 2 * kMtM; base < n_end; base += 2 * kMtM) {
        if (tid < kMtM) {
            const long long n0 = base + tid, n1 = n0 + kMtM;
            if (n0 < n_end) {
                const uint32_t x0 = ring[(n0 - kMtM) & kMask] ^ mt_twist(ring[(n0 - kMtN) & kMask], ring[(n0 - kMtN + 1) & kMask]);
                ring[n0 & kMask] = x0;
                if (n0 >= pos && n0 < end) words_out[n0 - pos] = x0;
                if (n1 < n_end) {
                    const uint32_t x1 = x0 ^ mt_twist(ring[(n1 - kMtN) & kMask], ring[(n1 - kMtN + 1) & kMask]);
                    ring[n1 & kMask] = x1;
                    if (n1 >= pos && n1 < end) words_out[n1 - pos] = x1;
                }
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < kMtN; i += 256) key_out[i] = ring[(kMtN * q + i) & kMask];
}

// phase[b * T + t, f] = angle(exp(2 pi i u)), u = rand()[b, f, t] = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from the two
// tempered words of element (b, f, t) (randomkit rk_double); the closed form of the angle and the 32 x 33 transposing
// tile are those of k_phase_from_uniform (gl_kernels.cu).
__global__ void __launch_bounds__(256) k_phase_from_words(const uint2* __restrict__ words, int n_bins, int n_frames,
                                                           float* __restrict__ phase) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const uint2* wb = words + (size_t)b * n_bins * n_frames;
    float* pb = phase + (size_t)b * n_frames * n_bins;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int f = f0 + ty + k, t = t0 + tx;
        if (f < n_bins && t < n_frames) {
            const uint2 w = wb[(size_t)f * n_frames + t];
            const double u = ((double)(mt_temper(w.x) >> 5) * 67108864.0 + (double)(mt_temper(w.y) >> 6)) * (1.0 / 9007199254740992.0);
            const double th = 6.283185307179586 * u;  // fl(2 pi) * u, one rounding like numpy
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

}  // namespace

int launch_phase_from_mt19937(int n_batch, int n_bins, int n_frames, const uint32_t* key, int pos, uint32_t* words,
                              float* phase, uint32_t* key_out, cudaStream_t stream) {
    const long long n = (long long)n_batch * n_bins * n_frames;
    k_mt19937_advance<<<1, 256, 0, stream>>>(key, pos, 2 * n, words, key_out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    if (n > 0) {
        dim3 grid((unsigned)((n_frames + 31) / 32), (unsigned)((n_bins + 31) / 32), (unsigned)n_batch);
        k_phase_from_words<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint2*>(words), n_bins, n_frames, phase);
        S2ST_CUDA_CHECK(cudaGetLastError());
    }
    return S2ST_OK;
}

}  // namespace s2st
