// Plan objects: device-resident constants only (internal).
#pragma once
#include "common.cuh"

constexpr int kMaxTimedPasses = 1025;

struct s2st_plan {
    int device;
    int n_fft, win_length, hop, n_mels;
    int rot, ws, wp, nz, nphase;
    // generic geometry (n_fft != 2048, any power of two in [64, 4096]): STFT / log-mel / mel projection only, run by
    // k_stft_generic (radix-2 Stockham FFT in shared memory); the warp-level 2048-point tables below are absent
    int generic;
    int n_bins;    // n_fft / 2 + 1
    float* gwin;   // [n_fft] padded window (audio_utils.py:218-223)
    float2* gtw;   // [n_fft / 2] exp(-2 pi i j / n_fft)
    int kb;        // active bins of the inverse-mel basis (kBins when none was given)
    int kb_pad;    // row pitch of inv_mel_t
    int num_sms;
    // device constants
    float* win_a;
    float* win_s;
    float* w2;
    float* inv_wss;
    float2* tw;
    float2* vtab;
    float* inv_mel_t;   // [n_mels, kb_pad] transposed pseudo-inverse (NULL if absent)
    float* inv_mel_tc;  // the same basis pre-split into TF32 head / tail in UMMA layout (mel_tc.cu), or NULL
    float* mel_tc;      // the mel filterbank pre-split into TF32 head / tail in UMMA layout, per 64-bin K chunk (mel_tc.cu), or NULL
    int mel_tc_chunks;  // K chunks that hold non-zero weights
    int opt_mel_simt;   // 0 = tcgen05 mel projection (default), 1 = FP32 SIMT CSR kernel
    int* mel_ptr;       // CSR of the mel filterbank: [n_mels + 1]
    int* mel_idx;       // [nnz] bin indices, ascending per row
    float* mel_val;     // [nnz]
    int mel_nnz;
    int mel_max_row;    // longest CSR row
    // column view for k_logmel_fast, or NULL when the bank does not qualify (every bin < 704 must feed at most two
    // adjacent mel bins, bins >= 704 none).  Lane l owns bins 22 l + j; entry [j * 32 + l] = (weight into mel bin b,
    // weight into b + 1, 1 if b is the same as for j - 1 else 0, slot as int bits): the lane keeps (lo, hi) partial
    // sums per run of equal b and writes them to slot (l * 17 + slot) of a per-warp slab.  mel_gather[m * 8 + q] are
    // the slab floats that add up to mel bin m (mel_terms = longest list; unused entries point at a zero).
    float4* mel_col;
    int* mel_gather;
    int mel_terms;
    // profiling aid (s2st_plan_set_pass_timing): CUDA events around every Griffin-Lim pass of the LAST call
    int strip_frames;             // 0 = choose per call (s2st_plan_set_strip_frames)
    // options (s2st_plan_set_option; initialised ONCE at plan creation from the S2ST_* environment variables)
    int opt_pdl;                  // programmatic dependent launch of the passes (default 1)
    int opt_frames;               // 1 (default): calls of up to opt_frames_max frames run the frame-parallel kernel (gl_frames.cuh)
    int opt_frames_max;
    int opt_inverse_mel_simt;     // 0 = tcgen05 inverse-mel (default), 1 = FP32 SIMT kernel
    int opt_frontend_generic;     // 0 = register-resident log-mel kernel (default), 1 = generic k_stft path
    int last_launches;            // kernel launches of the last gl_run (0 before the first call)
    int timing_enabled;
    int timing_recorded;          // events recorded by the last gl_run (passes + 1), 0 if none
    cudaEvent_t timing_events[kMaxTimedPasses + 1];
};

struct s2st_fbank_plan {
    int device;
    int sample_rate, n_bins;
    int win, shift, padded, log2_half;  // padded = FFT size, half = padded / 2 complex points
    int num_sms;
    float* window;      // [win] povey
    float2* tw;         // [padded / 2] exp(-2 pi i j / padded)   (covers both the half-size FFT and the split)
    int* mel_ptr;
    int* mel_idx;
    float* mel_val;
    int mel_nnz;
    // register-resident kernel (k_fbank_fast): -1 = not available for this rate, 0 = FFT 512 (one frame per
    // 256-point complex transform), 1 = FFT 256 (two frames per transform)
    int fast_mode;
    int opt_generic;    // s2st_fbank_plan_set_option: 1 = run the generic kernel even where the fast one applies
    float2* tw16;       // [256] exp(-2 pi i k1 n2 / 256) at [k1 * 16 + n2]
    float2* vsplit;     // [256] -i exp(-2 pi i k / 512)
    float* winp;        // window in the kernel's register layout, zero padded (see FbankFastParams)
    float4* mel_col;    // [256] per FFT bin; fast_mode 0: (w into mel bin b, w into b + 1, run continues ? 1 : 0, slot), see api.cu
    int* mel_gather;    // fast_mode 0: [n_bins * 8] slab floats that add up to each mel bin
    int mel_terms;      //              longest gather list
    int mel_zero;       //              index of the slab float that is kept at zero
};

namespace s2st {

// gl_kernels.cu
size_t gl_workspace_bytes(const s2st_plan* plan, int n_utts, long long total_frames);
int gl_run(s2st_plan* plan, int n_utts, long long total_frames, const int32_t* frame_offsets,
           const int32_t* frame_offsets_host, const float* logmel, const float* mag, int mag_kb, const float* phase,
           unsigned long long phase_seed, int n_iter, float* wave_out, void* workspace, size_t workspace_bytes,
           cudaStream_t stream);
int launch_inverse_mel(const s2st_plan* plan, long long n_frames, const float* logmel, bool is_log, float* mag,
                       int out_stride, int n_out, cudaStream_t stream);
int launch_rfft2048(const s2st_plan* plan, long long n, const float* in, float* out, bool inverse,
                    cudaStream_t stream);
int launch_phase_from_uniform(int n_batch, int n_bins, int n_frames, const double* u, float* phase, cudaStream_t stream);

// rng_kernels.cu
int launch_phase_from_mt19937(int n_batch, int n_bins, int n_frames, const uint32_t* key, int pos, uint32_t* words,
                              float* phase, uint32_t* key_out, cudaStream_t stream);

// mel_tc.cu (tcgen05 tensor-core path of the inverse-mel projection)
void build_inverse_mel_tc(const float* inv_mel, int kb, int K, float* out);
size_t inverse_mel_tc_floats(int K);
bool inverse_mel_tc_supported(const s2st_plan* plan);
int launch_inverse_mel_tc(const s2st_plan* plan, long long n_frames, const float* mel, bool is_log, float* mag,
                          int out_stride, int n_out, cudaStream_t stream, const float* basis_tc = nullptr);

int mel_project_tc_chunks(const float* mel, int n_mels, int n_bins);
size_t mel_project_tc_floats(int n_mels, int k_chunks);
void build_mel_project_tc(const float* mel, int n_mels, int n_bins, int k_chunks, float* out);
bool mel_project_tc_supported(const s2st_plan* plan);
int launch_mel_project_tc(const s2st_plan* plan, long long n_frames, const float* spec, float* out, cudaStream_t stream);

// frontend_kernels.cu
int launch_stft(const s2st_plan* plan, int n_utts, long long total_frames, const int64_t* wave_offsets,
                const int32_t* frame_offsets, const float* wave, float* mag_out, float* phase_out,
                float* logmel_out, float eps, const float* cmvn_mean, const float* cmvn_std,
                cudaStream_t stream, double* sums = nullptr);
int launch_mel_project(const s2st_plan* plan, long long n_frames, const float* spec, float* out,
                       cudaStream_t stream);
int launch_fbank(const s2st_fbank_plan* plan, int n_utts, long long total_frames,
                 const int64_t* wave_offsets, const int32_t* frame_offsets, const float* wave,
                 const float* cmvn_mean, const float* cmvn_std, float* out, cudaStream_t stream,
                 double* sums = nullptr);
int launch_cmvn(long long n_rows, int n_cols, const float* x, const float* mean, const float* std,
                float* out, bool denorm, cudaStream_t stream);
int launch_cmvn_accumulate(long long n_rows, int n_cols, const float* x, double* sums, cudaStream_t stream);


// transform_kernels.cu
int launch_utterance_cmvn(int n_utts, long long n_rows, const int32_t* fo, int n_cols, const float* x, float* out,
                          bool norm_means, bool norm_vars, float* stats, cudaStream_t stream);
int launch_utterance_sums(int n_utts, const int32_t* fo, int n_cols, const float* x, float* sums, cudaStream_t stream);
int launch_utterance_sum(int n_utts, const int32_t* fo, int n_cols, const float* x, double* sums, cudaStream_t stream);
int launch_fill_rects(int n_rects, const int32_t* rects, const float* values, int n_cols, float* x, cudaStream_t stream);

int launch_time_warp(int n_utts, long long n_rows, const int32_t* fo, int n_cols, const int32_t* warp, int arithmetic,
                     const float* x, float* out, cudaStream_t stream);
int launch_wave_to_pcm16(long long n, const float* x, short* out, cudaStream_t stream);
int launch_pcm16_to_wave(long long n, const short* pcm, float scale, float* out, cudaStream_t stream);

// dtw_kernels.cu
int launch_dtw(int bsz, int m, int n, const float* dist, const long long* shapes, float* cum, int* bp, int* path,
               cudaStream_t stream);
int launch_rms_dist_batch(int bsz, int max_m, int max_n, int d, const float* x1, const float* x2, const int* off1,
                          const int* off2, float* out, cudaStream_t stream);
int launch_rms_dist(int m, int n, int d, const float* x1, const float* x2, float* out, cudaStream_t stream);

}  // namespace s2st
