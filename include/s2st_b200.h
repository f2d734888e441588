/*
 * s2st_b200.h -- C ABI of the B200 (sm_100a) waveform-synthesis / feature front-end library.
 *
 * This is the drop-in boundary for the hot path of fengpeng-yue/speech-to-speech-translation
 * (a fairseq fork).  The reference's path is Python/PyTorch; the binding a maintainer adds is
 * the ctypes stub shown in INTEGRATION.md.  Every entry point cites the reference interface
 * it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers + sizes; no torch / C++ types, no exceptions across the boundary.
 *   - every function returns an int status: 0 = S2ST_OK, otherwise an S2ST_E* code;
 *     s2st_last_error() returns a thread-local human-readable message for the last failure.
 *   - pointers named *_dev are device pointers on the plan's device; *_host are host pointers.
 *   - the CALLER owns every input / output / workspace buffer; the library never allocates per
 *     call, enqueues all work on the caller's stream (cudaStream_t passed as void*) and never
 *     synchronises.  The only device memory the library owns are the plan's constants.
 *   - a plan also carries its OPTIONS (s2st_plan_set_option) and the profiling state of its last
 *     synthesis call (launch count, pass-timing events); functions that update either take a
 *     non-const plan.  There is no other mutable state: the S2ST_* environment variables are read
 *     once, when a plan is created, as the initial option values -- never per call.  One plan must
 *     not run two synthesis calls concurrently from different host threads.
 *   - all arithmetic is fp32.  Ragged batches are concatenated frame-major:
 *       frame_offsets[B+1]  (int32, device)  cumulative frame counts, frame_offsets[0] = 0
 *       features            [total_frames, n_mels]     row-major (the reference's [T, 80] layout)
 *       waveforms           concatenated; utterance i starts at (frame_offsets[i] - i) * hop and
 *                           has (T_i - 1) * hop samples (vocoder.py:98-99)
 */
#ifndef S2ST_B200_H_
#define S2ST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2ST_OK 0
#define S2ST_EINVAL 1      /* bad argument / unsupported configuration */
#define S2ST_ECUDA 2       /* a CUDA runtime call failed */
#define S2ST_EWORKSPACE 3  /* workspace too small */

#define S2ST_ABI_VERSION 2

typedef struct s2st_plan s2st_plan; /* opaque: owns only constants (windows, twiddles, mel matrices) */

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                     */
int s2st_abi_version(void);
const char* s2st_last_error(void);

/* ------------------------------------------------------------------------------------------ */
/* plan = the constants of GriffinLimVocoder.__init__ (fairseq/models/text_to_speech/vocoder.py:114-134),
 * GriffinLim.__init__ (:50-69), TTSSpectrogram.__init__ / TTSMelScale.__init__
 * (fairseq/data/audio/audio_utils.py:246-257, 275-282).
 *   window_host       [win_length]              the un-padded window (window_fn(win_length))
 *   inv_mel_host      [n_fft/2+1, n_mels]       pseudo-inverse mel basis (vocoder.py:28-32), may be NULL
 *   mel_host          [n_mels, n_fft/2+1]       mel filterbank (audio_utils.py:234-242), may be NULL
 * n_fft must be 2048 in this version (S2ST_EINVAL otherwise). */
int s2st_plan_create(s2st_plan** plan_out, int device, int n_fft, int win_length, int hop_length,
                     int n_mels, const float* window_host, const float* inv_mel_host,
                     const float* mel_host);
int s2st_plan_destroy(s2st_plan* plan);
/* number of leading spectrogram bins that can be non-zero after the inverse-mel projection
 * (rows of inv_mel beyond it are exactly zero; n_fft/2+1 when no inv_mel was given). */
int s2st_plan_active_bins(const s2st_plan* plan, int* active_bins_out);
/* Plan options (all have working defaults; they exist for A/B measurements and tests).  Initial values come from the
 * environment variable named in brackets, read once by s2st_plan_create. */
/* (options 1 and 3 -- a persistent launch of the strip kernel and a four-warps-per-strip mode for small calls -- were measured,
 * superseded by S2ST_OPT_GL_FRAMES and removed; s2st_plan_set_option rejects them) */
#define S2ST_OPT_GL_PDL 2           /* [S2ST_GL_PDL] 1 (default): programmatic dependent launch of the passes; 0: plain launches */
#define S2ST_OPT_GL_FRAMES 7        /* [S2ST_GL_FRAMES=0 | N] 1 (default): synthesis calls of up to 9 472 frames (one utterance, a small
                                       batch) run ALL iterations in ONE cooperative launch with a warp per frame, the overlap-add
                                       gathered from the neighbours' frames: the arithmetic of one strip per utterance (equal to
                                       rounding noise), bitwise independent of the batch, 3-4 x less latency than a launch per
                                       iteration.  0: always the strip kernels; N > 1: the frame-parallel kernel up to N frames.
                                       A pinned strip length (s2st_plan_set_strip_frames) keeps the strip kernels. */
#define S2ST_OPT_INVERSE_MEL 4      /* [S2ST_INVERSE_MEL=simt] 0 (default): tcgen05 tensor-core inverse-mel; 1: FP32 SIMT kernel */
#define S2ST_OPT_FRONTEND_GENERIC 5 /* [S2ST_LOGMEL_GENERIC / S2ST_FBANK_GENERIC] 0 (default): register-resident log-mel / fbank
                                       kernels where they apply; 1: always the generic kernels */
#define S2ST_OPT_MEL_PROJECT 6      /* 0 (default): tcgen05 tensor-core mel projection (s2st_mel_project); 1: FP32 SIMT CSR kernel */
int s2st_plan_set_option(s2st_plan* plan, int option, int value);

/* ------------------------------------------------------------------------------------------ */
/* Griffin-Lim ragged batch: GriffinLimVocoder.forward (vocoder.py:136-144) for B utterances at once.
 *   frame_offsets_host  optional host copy of frame_offsets (NULL allowed): lets the library size its
 *                   work decomposition exactly instead of from the average utterance length
 *   logmel_dev      [total_frames, n_mels]  denormalised log-mel, or NULL when mag_dev is given
 *   mag_dev         [total_frames, n_fft/2+1] linear magnitudes (GriffinLim.forward input,
 *                   vocoder.py:102, transposed to frame-major), or NULL when logmel_dev is given
 *   init_phase_dev  [total_frames, n_fft/2+1] initial phase in radians, frame-major.  The
 *                   reference draws it from numpy's global RNG on the host (vocoder.py:103-104);
 *                   the host shim does the same and uploads it.  NULL: the library draws
 *                   phi ~ U[-pi, pi) on the device from a counter-based generator keyed by
 *                   phase_seed (same distribution, not numpy's stream; for throughput runs).
 *   n_iter          spec_bwd_max_iter (n_iter forward + n_iter+1 inverse transforms)
 *   wave_out_dev    concatenated waveforms, (total_frames - B) * hop floats
 * Every utterance must have T >= 1; when n_iter > 0, (T-1)*hop must exceed n_fft/2 (the
 * reference's reflect padding raises otherwise) -- checked by the host shim, not here. */
int s2st_gl_workspace_bytes(const s2st_plan* plan, int n_utts, int64_t total_frames, size_t* bytes_out);
int s2st_gl_synthesize(s2st_plan* plan, int n_utts, int64_t total_frames,
                       const int32_t* frame_offsets_dev, const int32_t* frame_offsets_host,
                       const float* logmel_dev,
                       const float* mag_dev, const float* init_phase_dev, uint64_t phase_seed, int n_iter,
                       float* wave_out_dev, void* workspace_dev, size_t workspace_bytes,
                       void* stream);
/* Work decomposition of the Griffin-Lim kernel: a "strip" of `frames` consecutive frames is run by one warp.
 * 0 (default) lets every call pick the length that fills the GPU best for that batch.  Results are bitwise
 * reproducible for a given batch either way; pinning the length additionally makes each utterance's waveform
 * bitwise independent of what else is in the batch (the few samples shared by two strips are rounded slightly
 * differently from the others, so moving strip boundaries moves last-bit differences). */
int s2st_plan_set_strip_frames(s2st_plan* plan, int frames);
/* Profiling aid (not part of the reference's interface): when enabled, s2st_gl_synthesize / s2st_istft
 * record a CUDA event on the caller's stream before every Griffin-Lim pass and after the last one;
 * s2st_plan_get_pass_times waits for the last event and returns the device time of each pass of the
 * most recent call in milliseconds (pass 0 = initial inverse, passes 1..n_iter = fused iterations).  A call that ran the
 * frame-parallel kernel (S2ST_OPT_GL_FRAMES) is ONE launch: one value comes back, the whole synthesis. */
int s2st_plan_set_pass_timing(s2st_plan* plan, int enabled);
int s2st_plan_get_pass_times(s2st_plan* plan, float* ms_out_host, int capacity, int* n_passes_out);
/* number of kernel launches of the plan's most recent s2st_gl_synthesize call (for bench.py's gpu_launches); before
 * the first call: what a call with one launch per iteration would enqueue */
int s2st_gl_launch_count(const s2st_plan* plan, int n_iter, int from_logmel, int* launches_out);

/* The initial phase of GriffinLim.forward (vocoder.py:103-104): angles = angle(exp(2j * pi * rand(*shape))) cast to
 * float32.  The reference draws rand() from numpy's GLOBAL generator on the host; the host shim keeps exactly that draw
 * (so seeding numpy reproduces the reference) and uploads the float64 uniforms; this entry turns them into the phase on
 * the device and transposes the reference's [B, F, T] layout into the frame-major [B * T, F] that
 * s2st_gl_synthesize reads:  phase = theta if theta <= pi else theta - 2 pi,  theta = 2 pi u  (float64, then float32).
 *   uniform_dev [n_batch, n_bins, n_frames] float64 in [0, 1);  phase_out_dev [n_batch * n_frames, n_bins] float32 */
int s2st_phase_from_uniform(int n_batch, int n_bins, int n_frames, const double* uniform_dev, float* phase_out_dev,
                            void* stream);

/* The same initial phase with numpy's generator CONTINUED ON THE DEVICE: the host shim passes numpy's legacy global state
 * (np.random.get_state(): MT19937 key[624] + position), this entry produces the next 2 * n_batch * n_bins * n_frames
 * 32-bit outputs (what np.random.rand(*shape) consumes, randomkit's rk_double), the phase they give, and the key numpy
 * would hold afterwards (new position = pos + 2 n - 624 * max((pos + 2 n - 1) / 624, 0)); the shim puts it back with
 * np.random.set_state.  Bit-for-bit numpy's stream, without the host draw and the 8 B / element upload.
 *   key_dev [624] uint32; words_ws_dev scratch [2 * n] uint32 (8-byte aligned); key_out_dev [624] uint32 */
int s2st_phase_from_mt19937(int n_batch, int n_bins, int n_frames, const uint32_t* key_dev, int pos, uint32_t* words_ws_dev,
                            float* phase_out_dev, uint32_t* key_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* building blocks (test surface + the reference's public sub-modules)                         */

/* PseudoInverseMelScale.forward (vocoder.py:34-46), optionally fused with the exp of
 * GriffinLimVocoder.forward (vocoder.py:141):
 *   mag[t, f] = max(0, sum_m inv_mel[f, m] * g(mel[t, m])),  g = exp if input_is_log else identity,
 *   mel_dev [n_frames, n_mels], mag_dev [n_frames, n_fft/2+1] */
int s2st_inverse_mel(const s2st_plan* plan, int64_t n_frames, const float* mel_dev, int input_is_log,
                     float* mag_dev, void* stream);

/* TTSMelScale.forward (audio_utils.py:284-285) on frame-major data:
 *   mel_out[t, m] = sum_f mel[m, f] * spec[t, f]
 * The dense [n_mels x n_bins] contraction runs on the tensor cores (tcgen05, 3 x TF32 split, fp32 accumulation in TMEM;
 * ~1e-6 relative to an fp32 matmul) for filterbanks of up to 128 mel bins; S2ST_OPT_MEL_PROJECT selects the SIMT kernel. */
int s2st_mel_project(const s2st_plan* plan, int64_t n_frames, const float* spec_dev,
                     float* mel_out_dev, void* stream);

/* TTSSpectrogram.forward (audio_utils.py:259-271) for a ragged batch of waveforms:
 *   wave_offsets_dev [B+1] int64 sample offsets, frame_offsets_dev [B+1] int32 with
 *   T_i = 1 + n_i / hop.  mag_out / phase_out are [total_frames, n_fft/2+1]; phase_out may be NULL. */
int s2st_stft(const s2st_plan* plan, int n_utts, int64_t total_frames,
              const int64_t* wave_offsets_dev, const int32_t* frame_offsets_dev,
              const float* wave_dev, float* mag_out_dev, float* phase_out_dev, void* stream);

/* GriffinLim.inverse (vocoder.py:84-100): frame-major mag / phase -> concatenated waveforms.
 * Uses the same workspace as s2st_gl_synthesize. */
int s2st_istft(s2st_plan* plan, int n_utts, int64_t total_frames,
               const int32_t* frame_offsets_dev, const float* mag_dev, const float* phase_dev,
               float* wave_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* batched 2048-point real FFT / inverse (numpy rfft / irfft conventions), test surface for the
 * warp-level transform:  in [n, 2048] -> out [n, 1025, 2] (re, im)   and back. */
int s2st_rfft2048(const s2st_plan* plan, int64_t n, const float* in_dev, float* out_dev, void* stream);
int s2st_irfft2048(const s2st_plan* plan, int64_t n, const float* in_dev, float* out_dev, void* stream);

/* GriffinLim.get_window_sum_square (vocoder.py:71-82), host-side, out_host [n_fft + hop*(n_frames-1)] */
int s2st_window_sum_square(int n_frames, int hop_length, int win_length, int n_fft,
                           const float* window_host, float* out_host);

/* ------------------------------------------------------------------------------------------ */
/* feature front-end                                                                           */

/* logmelspec80: extract_logmel_spectrogram (examples/speech_synthesis/data_utils.py:46-76):
 *   out[t, m] = log(max(eps, sum_f mel[m, f] * |STFT(wave)|[t, f])), optionally followed by
 *   global CMVN (feature_transforms/global_cmvn.py:26-29) when cmvn_mean_dev != NULL.
 *   Same ragged layout as s2st_stft; out_dev [total_frames, n_mels].
 *   sums_dev (optional, NULL to skip): double [2, n_mels], the accumulators of get_global_cmvn
 *   (examples/speech_synthesis/data_utils.py:190-220) fused into the extraction -- the kernel ADDS
 *   (sum_t x, sum_t x^2) of the features it produces (before the CMVN, if any) to it, so the corpus is not
 *   read a second time for its statistics.  The caller zeroes it once and finishes mean / std on the host
 *   (double accumulation: the reference's float32 running sums are reproduced by s2st_utterance_sums). */
int s2st_logmel(const s2st_plan* plan, int n_utts, int64_t total_frames,
                const int64_t* wave_offsets_dev, const int32_t* frame_offsets_dev,
                const float* wave_dev, float eps, const float* cmvn_mean_dev,
                const float* cmvn_std_dev, double* sums_dev, float* out_dev, void* stream);

/* fbank80: _get_torchaudio_fbank (audio_utils.py:136-149) == torchaudio.compliance.kaldi.fbank with
 * num_mel_bins = n_bins, sample_frequency = sample_rate and every other option at its default.
 * The Kaldi plan owns the povey window, FFT twiddles and the Kaldi mel banks. */
typedef struct s2st_fbank_plan s2st_fbank_plan;
int s2st_fbank_plan_create(s2st_fbank_plan** plan_out, int device, int sample_rate, int n_bins);
int s2st_fbank_plan_destroy(s2st_fbank_plan* plan);
/* option S2ST_OPT_FRONTEND_GENERIC only (initial value: S2ST_FBANK_GENERIC, read once at plan creation) */
int s2st_fbank_plan_set_option(s2st_fbank_plan* plan, int option, int value);
/* window size / shift / padded FFT size of the plan: m_i = 1 + (n_i - win) / shift (0 if n_i < win) */
int s2st_fbank_frame_params(const s2st_fbank_plan* plan, int* win_out, int* shift_out, int* padded_out);
/*   wave_dev is the int16-scaled waveform (audio_utils.py:105-106); out_dev [total_frames, n_bins];
 *   optional fused global CMVN and optional fused statistics (sums_dev double [2, n_bins]) as above. */
int s2st_fbank(const s2st_fbank_plan* plan, int n_utts, int64_t total_frames,
               const int64_t* wave_offsets_dev, const int32_t* frame_offsets_dev,
               const float* wave_dev, const float* cmvn_mean_dev, const float* cmvn_std_dev,
               double* sums_dev, float* out_dev, void* stream);

/* GlobalCMVN.__call__ (feature_transforms/global_cmvn.py:26-29): out = (x - mean) / std, and its
 * inverse gcmvn_denormalize (fairseq/speech_generator_for_s2st.py:21-29): out = x * std + mean.
 * x_dev / out_dev [n_rows, n_cols] (in place allowed), mean / std [n_cols]. */
int s2st_cmvn_apply(int64_t n_rows, int n_cols, const float* x_dev, const float* mean_dev,
                    const float* std_dev, float* out_dev, void* stream);
int s2st_cmvn_denormalize(int64_t n_rows, int n_cols, const float* x_dev, const float* mean_dev,
                          const float* std_dev, float* out_dev, void* stream);
/* get_global_cmvn's accumulation (examples/speech_synthesis/data_utils.py:190-220):
 *   sums_dev [2, n_cols] += (sum_t x, sum_t x^2); the caller zeroes sums_dev and finishes
 *   mean / std on the host. */
int s2st_cmvn_accumulate(int64_t n_rows, int n_cols, const float* x_dev, double* sums_dev, void* stream);

/* UtteranceCMVN.__call__ (feature_transforms/utterance_cmvn.py:29-40) for a ragged batch: utterance u owns rows
 * frame_offsets_dev[u] .. frame_offsets_dev[u+1] of x_dev [total_rows, n_cols] (total_rows = frame_offsets[n_utts]).
 * Per utterance and column: mean = sum_t x / T, var = sum_t x^2 / T - mean^2 (float32, accumulated in row order like
 * numpy's axis-0 reduction), out = x - mean if norm_means, then / sqrt(max(var, 1e-10)) if norm_vars.  Bit-identical
 * to the reference.  stats_dev: caller-owned workspace of n_utts * 2 * n_cols floats; it returns (mean, std) per
 * utterance.  In place allowed. */
int s2st_utterance_cmvn(int n_utts, int64_t total_rows, const int32_t* frame_offsets_dev, int n_cols,
                        const float* x_dev, int norm_means, int norm_vars, float* out_dev, float* stats_dev,
                        void* stream);
/* The per-file terms of get_global_cmvn (examples/speech_synthesis/data_utils.py:197-208): for every utterance (= one
 * .npy file of the feature directory) cur_mean_x = frames.sum(axis=0) and cur_mean_x2 = (frames ** 2).sum(axis=0),
 * accumulated in float32 in row order without FMA contraction, exactly as numpy reduces a C-contiguous [T, n] array
 * over axis 0: bit-identical to the reference.  sums_dev float32 [n_utts, 2, n_cols] = (sum, sum of squares); the host
 * shim adds the files up in the reference's order (float32) and finishes mean / std. */
int s2st_utterance_sums(int n_utts, const int32_t* frame_offsets_dev, int n_cols, const float* x_dev,
                        float* sums_dev, void* stream);
/* Per-utterance sum of all elements (double): the "local mean" mask value of SpecAugmentTransform
 * (feature_transforms/specaugment.py:88-89) is sums_dev[u] / (T_u * n_cols). */
int s2st_utterance_sum(int n_utts, const int32_t* frame_offsets_dev, int n_cols, const float* x_dev,
                       double* sums_dev, void* stream);
/* The masking of SpecAugmentTransform.__call__ (specaugment.py:111-131): x[row0:row1, col0:col1] = value for every
 * rectangle; rects_dev int32 [n_rects, 4] = (row0, row1, col0, col1) in rows of the concatenated [total_rows, n_cols]
 * matrix, values_dev [n_rects].  The rectangles are drawn on the host with the reference's RNG call sequence. */
int s2st_fill_rects(int n_rects, const int32_t* rects_dev, const float* values_dev, int n_cols, float* x_dev,
                    void* stream);

/* batch_dynamic_time_warping (examples/s2s_trans/tasks/s2s_translation.py:414-464, the MCD validation metric):
 * distance_dev [bsz, m, n] float32 (zero padded), shapes_dev int64 [bsz, 2] = (M_b, N_b) or NULL (= full size)
 * -> cumdist_dev float32, backptr_dev int32 (0 = left, 1 = up-left, 2 = up), pathmap_dev int32 (1 on the optimal
 * path from (M_b-1, N_b-1) back to (0, 0)), all [bsz, m, n].  Bit-identical to the reference run on the CPU. */
int s2st_dtw(int bsz, int m, int n, const float* distance_dev, const int64_t* shapes_dev, float* cumdist_dev,
             int32_t* backptr_dev, int32_t* pathmap_dev, void* stream);
/* compute_rms_dist (s2s_translation.py:467-475): out_dev [m, n] = sqrt(|x1[i] - x2[j]|^2 / d), x1_dev [m, d],
 * x2_dev [n, d]. */
int s2st_rms_dist(int m, int n, int d, const float* x1_dev, const float* x2_dev, float* out_dev, void* stream);

/* The time warping of SpecAugmentTransform.__call__ (specaugment.py:96-110) for a ragged batch: warp_dev int32 [n_utts, 2] =
 * (w0, w) per utterance as the reference draws them (w0 < 0: the utterance is copied); rows [0, w0) are resized to w0 + w
 * rows and rows [w0, T) to T - w0 - w rows like cv2.resize(..., INTER_LINEAR) on float32.  arithmetic = 1: the x86-64
 * opencv-python wheels (IPP: source position in double, dst = fma(S1 - S0, w, S0)); arithmetic = 0: OpenCV's own code
 * (position rounded to float, dst = S0 * (1 - w) + S1 * w, both products rounded).  Both are bit-exact against fixtures made
 * with the reference class.  out_dev must not alias x_dev. */
int s2st_time_warp(int n_utts, int64_t total_rows, const int32_t* frame_offsets_dev, int n_cols, const int32_t* warp_dev,
                   int arithmetic, const float* x_dev, float* out_dev, void* stream);

/* Waveform post-processing for file output (examples/s2s_trans/generate_waveform.py:115-124: sf.write of the float
 * waveform, which soundfile stores as 16-bit PCM): pcm[i] = saturate_int16(lrint(wave[i] * 32767)), NaN -> 0.  Runs on
 * the concatenated batch so the device-to-host copy carries 2 bytes per sample. */
int s2st_wave_to_pcm16(int64_t n_samples, const float* wave_dev, int16_t* pcm_out_dev, void* stream);
/* The input side (fairseq/data/audio/audio_utils.py:65-109, get_waveform: soundfile reads a 16-bit PCM file as float32,
 * value / 32768 with normalization=True, the int16 value itself for get_fbank's Kaldi-style input): the samples are uploaded
 * as the 2 bytes they occupy on disk and converted on the device, wave[i] = (float)pcm[i] * scale  (scale = 1 / 32768 or 1;
 * exact).  Halves the host-to-device traffic of the PCIe-bound feature front-end. */
int s2st_pcm16_to_wave(int64_t n_samples, const int16_t* pcm_dev, float scale, float* wave_out_dev, void* stream);

/* The padded distance batch of batch_compute_distortion (s2s_translation.py:489-505) in one launch: pair b compares rows
 * offsets1[b] .. offsets1[b+1] of x1_dev [sum M_b, d] with rows offsets2[b] .. offsets2[b+1] of x2_dev [sum N_b, d];
 * out_dev [bsz, max_m, max_n] receives compute_rms_dist of the pair in its top-left M_b x N_b corner, zeros elsewhere
 * (the reference pads and stacks the per-pair matrices on the host). */
int s2st_rms_dist_batch(int bsz, int max_m, int max_n, int d, const float* x1_dev, const float* x2_dev,
                        const int32_t* offsets1_dev, const int32_t* offsets2_dev, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2ST_B200_H_ */
