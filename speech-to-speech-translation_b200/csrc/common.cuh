// Shared declarations of the s2st_b200 CUDA library (internal; the public ABI is include/s2st_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2st {

constexpr int kNfft = 2048;
constexpr int kBins = kNfft / 2 + 1;  // 1025
constexpr int kMinStrip = 4;          // shortest strip (frames a warp runs in sequence); must cover the overlap
constexpr int kGlThreads = 512;       // 16 warps per CTA, 1 CTA per SM (128 registers per thread = the whole file)

struct UttDesc {
    long long wave_off;  // first sample of this utterance in the concatenated waveform buffers
    int frame_off;       // first row in the frame-major tensors
    int n_frames;        // T
};

// One strip = nf consecutive frames of one utterance, with everything a warp needs to run it (32 bytes,
// fetched with two 16-byte loads, no dependent lookup).
struct TileDesc {
    long long wave_off;  // first sample of the utterance in the concatenated waveform buffers
    int frame_off;       // first row of the utterance in the frame-major tensors
    int n_frames;        // T of the utterance
    int f0;              // first frame (utterance-local) of the strip
    int nf;              // frames in the strip (1..S)
    int prev;            // strip table index of the strip before / after this one in the same utterance, -1 if none
    int next;            // (kept for tools that walk an utterance's strips; the kernels do not need them)
};

// Everything the Griffin-Lim kernels need, passed by value.
struct GlParams {
    int device;    // host-side bookkeeping only
    int pdl;       // host-side: launch the passes with programmatic dependent launch
    // geometry
    int hop;       // H
    int half;      // n_fft / 2 (reflect padding, trimmed from both ends)
    int rot;       // s: first frame sample kept (even); frames are processed circularly rotated by s
    int ws;        // even upper bound of the window support in rotated coordinates (<= wp)
    int wp;        // 64 * NZ: samples each lane-register layout covers
    int nphase;    // frames overlapping one sample = ceil(ws / hop)
    int kb;        // bins >= kb have zero target magnitude
    int mag_stride, phase_stride;
    // constants (device global memory, owned by the plan)
    const float* win_a;      // [wp] window, rotated (analysis and synthesis; 1 / n_fft is folded into inv_wss)
    const float* w2;         // [ws] squared window, rotated (edge window-sum-square)
    const float* inv_wss;    // [hop] 1 / (n_fft * steady-state window-sum-square) by (q mod hop)
    float inv_nfft;          // 1 / n_fft (a power of two: the scaling is exact wherever it is applied)
    const float2* tw;        // [32*32] exp(-2 pi i r l / 1024)
    const float2* vtab;      // [1024]  -i exp(-2 pi i k / 2048)
    // batch tables (workspace)
    const UttDesc* utts;
    const TileDesc* tiles;
    const int* n_tiles;
    // data
    const float* mag;        // [total_frames, mag_stride]
    const float* phase;      // [total_frames, phase_stride] (first pass only); NULL -> device RNG
    unsigned long long phase_seed;
    const float* in;         // normalised waveforms written by the previous pass
    float* out;              // waveforms written by this pass (seams pre-zeroed)
    float* zero_next;        // buffer the NEXT pass writes: this pass zeroes its seams
};

struct s2st_error_state;
void set_error(const char* fmt, ...);

#define S2ST_CUDA_CHECK(expr)                                                               \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::s2st::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                              __LINE__);                                                    \
            return S2ST_ECUDA;                                                              \
        }                                                                                   \
    } while (0)

}  // namespace s2st
