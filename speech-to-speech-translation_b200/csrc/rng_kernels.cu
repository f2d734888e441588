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
// MT19937 regenerates its 624-word state block by block: new[i] = new-or-old[i + 397 mod 624] ^ twist(old[i], old[i + 1]),
// i.e. the words 227 .. 623 of a block depend on words 0 .. 396 of the SAME block.  Substituting those once (twice for
// the last third) makes every word of the new block a function of the OLD block alone,
//   i <  227:  new[i] = old[i + 397] ^ tw(old[i], old[i + 1])
//   i <  454:  new[i] = old[i + 170] ^ tw(old[i - 227], old[i - 226]) ^ tw(old[i], old[i + 1])
//   i <= 623:  new[i] = old[i -  57] ^ tw(old[i - 454], old[i - 453]) ^ tw(old[i - 227], old[i - 226]) ^ tw(old[i], old[i + 1])
// (old[624] := new[0]), so 624 threads produce a whole block per barrier with no dependent chain between them: the
// sequential part of the generator is one shared-memory round trip + ~20 ALU operations + one block barrier per 624
// outputs (the textbook order needs a barrier every 227, the two-step order of the first version every 454).
// Everything else (tempering, the conversion to doubles, the phase) is done by the second, fully parallel kernel.
#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {

namespace {

constexpr int kMtN = 624, kMtThreads = 640;

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

// words_out[k] = X[pos + k] for k < n_words (untempered), X = the concatenation of the state blocks starting with
// key_in; key_out = the block numpy's state holds afterwards (block (pos + n_words - 1) / 624).
__global__ void __launch_bounds__(kMtThreads) k_mt19937_advance(const uint32_t* __restrict__ key_in, int pos, long long n_words,
                                                                 uint32_t* __restrict__ words_out, uint32_t* __restrict__ key_out) {
    __shared__ uint32_t blk[2][kMtN];
    const int i = threadIdx.x;
    const bool worker = i < kMtN;
    const long long end = pos + n_words;
    const long long q = end > 0 ? (end - 1) / kMtN : 0;  // last block that is needed
    if (worker) {
        const uint32_t v = key_in[i];
        blk[0][i] = v;
        if (i >= pos && i < end) words_out[i - pos] = v;
    }
    __syncthreads();
    int cur = 0;
    uint32_t* out = words_out + ((long long)kMtN + i - pos);  // &words_out[624 b + i - pos] for b = 1 (dereferenced only when valid)
    for (long long b = 1; b <= q; ++b) {
        if (worker) {
            const uint32_t* o = blk[cur];
            uint32_t x;
            if (i < 227) {
                x = o[i + 397] ^ mt_twist(o[i], o[i + 1]);
            } else if (i < 454) {
                x = o[i + 170] ^ mt_twist(o[i - 227], o[i - 226]) ^ mt_twist(o[i], o[i + 1]);
            } else {
                const uint32_t nxt = i < kMtN - 1 ? o[i + 1] : (o[397] ^ mt_twist(o[0], o[1]));  // old[624] := new[0]
                x = o[i - 57] ^ mt_twist(o[i - 454], o[i - 453]) ^ mt_twist(o[i - 227], o[i - 226]) ^ mt_twist(o[i], nxt);
            }
            blk[cur ^ 1][i] = x;
            if (b < q || kMtN * b + i < end) *out = x;  // only the last block can reach past the outputs
        }
        out += kMtN;
        cur ^= 1;
        __syncthreads();
    }
    if (worker) key_out[i] = blk[cur][i];
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
    k_mt19937_advance<<<1, kMtThreads, 0, stream>>>(key, pos, 2 * n, words, key_out);
    S2ST_CUDA_CHECK(cudaGetLastError());
    if (n > 0) {
        dim3 grid((unsigned)((n_frames + 31) / 32), (unsigned)((n_bins + 31) / 32), (unsigned)n_batch);
        k_phase_from_words<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint2*>(words), n_bins, n_frames, phase);
        S2ST_CUDA_CHECK(cudaGetLastError());
    }
    return S2ST_OK;
}

}  // namespace s2st
