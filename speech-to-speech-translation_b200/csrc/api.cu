// C ABI of libs2st_b200.so (declared in include/s2st_b200.h): argument checking, plan construction
// (host-side constant tables uploaded once), dispatch to the kernels.  No per-call allocation, no
// synchronisation; errors become integer status codes + a thread-local message.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/s2st_b200.h"
#include "plan.h"

namespace s2st {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    if (src.empty()) return S2ST_OK;
    S2ST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(dst), sizeof(T) * src.size()));
    S2ST_CUDA_CHECK(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return S2ST_OK;
}

// dense [rows, cols] -> CSR over non-zeros (column order kept ascending)
void to_csr(const float* dense, int rows, int cols, std::vector<int>& ptr, std::vector<int>& idx,
            std::vector<float>& val, int* max_row) {
    ptr.assign(rows + 1, 0);
    idx.clear();
    val.clear();
    *max_row = 0;
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            const float v = dense[(size_t)r * cols + c];
            if (v != 0.0f) {
                idx.push_back(c);
                val.push_back(v);
            }
        }
        ptr[r + 1] = (int)idx.size();
        if (ptr[r + 1] - ptr[r] > *max_row) *max_row = ptr[r + 1] - ptr[r];
    }
}

// the mel filterbank in tensor-core form (mel_tc.cu); skipped for banks wider than the kernel's accumulator
int upload_mel_tc(s2st_plan* p, const float* mel_host) {
    if (p->n_mels > 128) return S2ST_OK;
    p->mel_tc_chunks = mel_project_tc_chunks(mel_host, p->n_mels, p->n_bins);
    if (p->mel_tc_chunks == 0) return S2ST_OK;
    std::vector<float> tc(mel_project_tc_floats(p->n_mels, p->mel_tc_chunks));
    build_mel_project_tc(mel_host, p->n_mels, p->n_bins, p->mel_tc_chunks, tc.data());
    return upload(&p->mel_tc, tc);
}

void free_plan_members(s2st_plan* p) {
    for (int i = 0; i <= kMaxTimedPasses; ++i)
        if (p->timing_events[i]) cudaEventDestroy(p->timing_events[i]);
    cudaFree(p->gwin);
    cudaFree(p->gtw);
    cudaFree(p->win_a);
    cudaFree(p->win_s);
    cudaFree(p->w2);
    cudaFree(p->inv_wss);
    cudaFree(p->tw);
    cudaFree(p->vtab);
    cudaFree(p->inv_mel_t);
    cudaFree(p->inv_mel_tc);
    cudaFree(p->mel_tc);
    cudaFree(p->mel_ptr);
    cudaFree(p->mel_idx);
    cudaFree(p->mel_val);
    cudaFree(p->mel_col);
    cudaFree(p->mel_gather);
}

}  // namespace
}  // namespace s2st

using namespace s2st;

extern "C" {

int s2st_abi_version(void) { return S2ST_ABI_VERSION; }

const char* s2st_last_error(void) { return g_err; }

int s2st_plan_create(s2st_plan** plan_out, int device, int n_fft, int win_length, int hop_length,
                     int n_mels, const float* window_host, const float* inv_mel_host,
                     const float* mel_host) {
    if (!plan_out || !window_host) {
        set_error("null argument");
        return S2ST_EINVAL;
    }
    *plan_out = nullptr;
    if (n_fft != kNfft && (n_fft < 64 || n_fft > 4096 || (n_fft & (n_fft - 1)) != 0)) {
        set_error("n_fft=%d is not supported: the STFT / log-mel / mel-projection entry points take any power of two in "
                  "[64, 4096], Griffin-Lim synthesis takes n_fft = %d only", n_fft, kNfft);
        return S2ST_EINVAL;
    }
    if (n_fft != kNfft && inv_mel_host) {
        set_error("an inverse-mel basis (Griffin-Lim synthesis) needs n_fft = %d, got %d", kNfft, n_fft);
        return S2ST_EINVAL;
    }
    if (win_length < 1 || win_length > n_fft || hop_length < 1 || n_mels < 1) {
        set_error("bad geometry: win_length=%d hop_length=%d n_mels=%d", win_length, hop_length, n_mels);
        return S2ST_EINVAL;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("cannot select CUDA device %d", device);
        return S2ST_ECUDA;
    }
    // padded window (audio_utils.py:218-223)
    std::vector<float> w(n_fft, 0.0f);
    const int pl = (n_fft - win_length) / 2;
    for (int i = 0; i < win_length; ++i) w[pl + i] = window_host[i];
    int lo = -1, hi = -1;
    for (int i = 0; i < n_fft; ++i)
        if (w[i] != 0.0f) {
            if (lo < 0) lo = i;
            hi = i;
        }
    if (lo < 0) {
        set_error("window is identically zero");
        return S2ST_EINVAL;
    }
    s2st_plan* p = new s2st_plan();
    std::memset(p, 0, sizeof(*p));
    {   // options: the S2ST_* environment variables are read HERE, once; afterwards only s2st_plan_set_option changes them
        const char* e = getenv("S2ST_GL_PDL");
        p->opt_pdl = !(e && e[0] == '0');
        e = getenv("S2ST_GL_FRAMES");  // 0 = never the frame-parallel kernel, N > 1 = up to N frames
        p->opt_frames = !(e && e[0] == '0');
        p->opt_frames_max = (e && atoi(e) > 1) ? atoi(e) : 16 * 148 * 4;
        e = getenv("S2ST_INVERSE_MEL");
        p->opt_inverse_mel_simt = (e && e[0] == 's') ? 1 : 0;
        p->opt_frontend_generic = getenv("S2ST_LOGMEL_GENERIC") ? 1 : 0;
    }
    p->device = device;
    p->n_fft = n_fft;
    p->win_length = win_length;
    p->hop = hop_length;
    p->n_mels = n_mels;
    p->n_bins = n_fft / 2 + 1;
    if (n_fft != kNfft) {
        // generic geometry: padded window, radix-2 twiddles, CSR mel bank; nothing of the 2048-point machinery
        p->generic = 1;
        p->kb = p->n_bins;
        std::vector<float2> gtw(n_fft / 2);
        for (int j = 0; j < n_fft / 2; ++j) {
            const double a = -2.0 * 3.14159265358979323846 * (double)j / (double)n_fft;
            gtw[j] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        int rc = S2ST_OK;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
            set_error("cudaGetDeviceProperties failed");
            rc = S2ST_ECUDA;
        } else {
            p->num_sms = prop.multiProcessorCount;
        }
        if (rc == S2ST_OK) rc = upload(&p->gwin, w);
        if (rc == S2ST_OK) rc = upload(&p->gtw, gtw);
        if (rc == S2ST_OK && mel_host) {
            std::vector<int> ptr, idx;
            std::vector<float> val;
            to_csr(mel_host, n_mels, p->n_bins, ptr, idx, val, &p->mel_max_row);
            p->mel_nnz = (int)idx.size();
            if (idx.empty()) {
                idx.push_back(0);
                val.push_back(0.0f);
            }
            rc = upload(&p->mel_ptr, ptr);
            if (rc == S2ST_OK) rc = upload(&p->mel_idx, idx);
            if (rc == S2ST_OK) rc = upload(&p->mel_val, val);
            if (rc == S2ST_OK) rc = upload_mel_tc(p, mel_host);
        }
        if (rc != S2ST_OK) {
            free_plan_members(p);
            delete p;
            return rc;
        }
        *plan_out = p;
        return S2ST_OK;
    }
    // frames are processed circularly rotated so that the window support starts at sample 0; when it
    // fits in 19*64 samples the 13 structurally-zero inputs of every in-lane FFT are pruned.
    int rot = lo & ~1;
    if (hi - rot < 19 * 64 && rot + 19 * 64 <= n_fft) {
        p->nz = 19;
    } else {
        p->nz = 32;
        rot = 0;
    }
    p->rot = rot;
    p->wp = 64 * p->nz;
    p->ws = ((hi - rot + 1) + 1) & ~1;
    if (p->ws > p->wp) p->ws = p->wp;
    const int overlap = (p->ws + hop_length - 1) / hop_length;
    p->nphase = overlap;

    std::vector<float> win_a(p->wp), win_s(p->wp), w2(p->ws), inv_wss(hop_length);
    for (int m = 0; m < p->wp; ++m) {
        const float v = (rot + m < n_fft) ? w[rot + m] : 0.0f;
        win_a[m] = v;
        win_s[m] = v / (float)n_fft;
    }
    for (int m = 0; m < p->ws; ++m) w2[m] = win_a[m] * win_a[m];
    for (int r = 0; r < hop_length; ++r) {
        // ascending frame order == descending offset, float accumulation like vocoder.py:78-81
        float acc = 0.0f;
        int imax = -1;
        for (int i = 0; r + i * hop_length < p->ws; ++i) imax = i;
        for (int i = imax; i >= 0; --i) acc += w2[r + i * hop_length];
        // 1 / n_fft (the inverse transform's scale, a power of two) is folded in here
        inv_wss[r] = (acc > 1.1754944e-38f ? (float)(1.0 / (double)acc) : 1.0f) / (float)n_fft;
    }
    std::vector<float2> tw(1024), vtab(1024);
    const double pi = 3.14159265358979323846;
    for (int r = 0; r < 32; ++r)
        for (int l = 0; l < 32; ++l) {
            const double a = -2.0 * pi * (double)((r * l) % 1024) / 1024.0;
            tw[r * 32 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k = 0; k < 1024; ++k) {
        const double a = 2.0 * pi * (double)k / 2048.0;  // -i * exp(-i a) = (-sin a, -cos a)
        vtab[k] = make_float2((float)(-std::sin(a)), (float)(-std::cos(a)));
    }
    vtab[0] = make_float2(0.0f, -1.0f);
    vtab[512] = make_float2(-1.0f, 0.0f);

    int rc = S2ST_OK;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        rc = S2ST_ECUDA;
    } else {
        p->num_sms = prop.multiProcessorCount;
    }
    if (rc == S2ST_OK) rc = upload(&p->win_a, win_a);
    if (rc == S2ST_OK) rc = upload(&p->win_s, win_s);
    if (rc == S2ST_OK) rc = upload(&p->w2, w2);
    if (rc == S2ST_OK) rc = upload(&p->inv_wss, inv_wss);
    if (rc == S2ST_OK) rc = upload(&p->tw, tw);
    if (rc == S2ST_OK) rc = upload(&p->vtab, vtab);
    p->kb = kBins;
    if (rc == S2ST_OK && inv_mel_host) {
        int last = 0;
        for (int f = 0; f < kBins; ++f)
            for (int m = 0; m < n_mels; ++m)
                if (inv_mel_host[(size_t)f * n_mels + m] != 0.0f) last = f;
        p->kb = last + 1;
        p->kb_pad = (p->kb + 3) & ~3;
        std::vector<float> t((size_t)n_mels * p->kb_pad, 0.0f);
        for (int f = 0; f < p->kb; ++f)
            for (int m = 0; m < n_mels; ++m) t[(size_t)m * p->kb_pad + f] = inv_mel_host[(size_t)f * n_mels + m];
        rc = upload(&p->inv_mel_t, t);
        if (rc == S2ST_OK && p->kb <= 704 && n_mels % 8 == 0 && n_mels <= 80) {
            std::vector<float> tc(inverse_mel_tc_floats(n_mels));
            build_inverse_mel_tc(inv_mel_host, p->kb, n_mels, tc.data());
            rc = upload(&p->inv_mel_tc, tc);
        }
    }
    if (rc == S2ST_OK && mel_host) {
        std::vector<int> ptr, idx;
        std::vector<float> val;
        to_csr(mel_host, n_mels, kBins, ptr, idx, val, &p->mel_max_row);
        p->mel_nnz = (int)idx.size();
        if (idx.empty()) {  // keep the device arrays non-null
            idx.push_back(0);
            val.push_back(0.0f);
        }
        rc = upload(&p->mel_ptr, ptr);
        if (rc == S2ST_OK) rc = upload(&p->mel_idx, idx);
        if (rc == S2ST_OK) rc = upload(&p->mel_val, val);
        if (rc == S2ST_OK) rc = upload_mel_tc(p, mel_host);
        // column view (k_logmel_fast): lane l owns bins 22 l .. 22 l + 21 and keeps one pair of partial sums per run
        // of bins that feed the same mel bin b ("slot"); see plan.h
        constexpr int kpl = 22, lanes = 32, max_slots = 17, max_terms = 8;
        std::vector<float4> col(kpl * lanes, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        std::vector<std::vector<int>> terms(n_mels);
        bool ok = n_mels <= 128;
        for (int k = 704; k < kBins && ok; ++k)
            for (int m = 0; m < n_mels; ++m)
                if (mel_host[(size_t)m * kBins + k] != 0.0f) ok = false;
        for (int l = 0; l < lanes && ok; ++l) {
            int cur = -1, slot = -1;
            for (int j = 0; j < kpl && ok; ++j) {
                const int k = kpl * l + j;
                int first = -1, last = -1, count = 0;
                for (int m = 0; m < n_mels; ++m)
                    if (mel_host[(size_t)m * kBins + k] != 0.0f) {
                        if (first < 0) first = m;
                        last = m;
                        ++count;
                    }
                if (count > 2 || (count == 2 && last != first + 1)) {
                    ok = false;
                    break;
                }
                int b = count ? first : (cur < 0 ? 0 : cur);  // an empty column continues the current run
                if (count == 1 && cur >= 0 && first == cur + 1) b = cur;  // only the upper bin of the run: stay in it
                const bool keep = (b == cur);
                if (!keep) {
                    ++slot;
                    cur = b;
                    if (slot >= max_slots) {
                        ok = false;
                        break;
                    }
                    terms[b].push_back((slot * lanes + l) * 2);  // slab[slot][lane]
                    if (b + 1 < n_mels) terms[b + 1].push_back((slot * lanes + l) * 2 + 1);
                }
                float w0 = 0.0f, w1 = 0.0f;
                if (count) {
                    if (first == b) w0 = mel_host[(size_t)first * kBins + k];
                    else w1 = mel_host[(size_t)first * kBins + k];
                    if (count == 2) w1 = mel_host[(size_t)last * kBins + k];
                }
                float sf;
                const int slot_index = slot * lanes;  // float2 index of slab[slot][0]; the kernel adds the lane
                std::memcpy(&sf, &slot_index, sizeof(float));
                col[j * lanes + l] = make_float4(w0, w1, keep ? 1.0f : 0.0f, sf);
            }
        }
        int n_terms = 1;
        for (int m = 0; m < n_mels && ok; ++m) {
            if ((int)terms[m].size() > max_terms) ok = false;
            n_terms = std::max(n_terms, (int)terms[m].size());
        }
        if (ok && rc == S2ST_OK) {
            // unused gather entries point at a float the kernel keeps at zero (index 2 * 32 * 17)
            // two planes of int4 per mel bin: entry q of bin m at ((q / 4) * n_mels + m) * 4 + q % 4
            std::vector<int> gather((size_t)n_mels * max_terms, 2 * lanes * max_slots);
            for (int m = 0; m < n_mels; ++m)
                for (size_t q = 0; q < terms[m].size(); ++q) gather[((q / 4) * n_mels + m) * 4 + q % 4] = terms[m][q];
            p->mel_terms = n_terms;
            rc = upload(&p->mel_col, col);
            if (rc == S2ST_OK) rc = upload(&p->mel_gather, gather);
        }
    }
    if (rc != S2ST_OK) {
        free_plan_members(p);
        delete p;
        return rc;
    }
    *plan_out = p;
    return S2ST_OK;
}

int s2st_plan_destroy(s2st_plan* plan) {
    if (!plan) return S2ST_OK;
    DeviceGuard guard(plan->device);
    free_plan_members(plan);
    delete plan;
    return S2ST_OK;
}

int s2st_plan_active_bins(const s2st_plan* plan, int* active_bins_out) {
    if (!plan || !active_bins_out) {
        set_error("null argument");
        return S2ST_EINVAL;
    }
    *active_bins_out = plan->kb;
    return S2ST_OK;
}

int s2st_plan_set_strip_frames(s2st_plan* plan, int frames) {
    if (!plan || frames < 0 || frames > 4096) {
        set_error("bad argument to s2st_plan_set_strip_frames");
        return S2ST_EINVAL;
    }
    plan->strip_frames = frames;
    return S2ST_OK;
}

int s2st_plan_set_option(s2st_plan* plan, int option, int value) {
    if (!plan) {
        set_error("null plan");
        return S2ST_EINVAL;
    }
    switch (option) {
        case S2ST_OPT_GL_PDL:
            plan->opt_pdl = value != 0;
            return S2ST_OK;
        case S2ST_OPT_GL_FRAMES:
            if (value < 0) break;
            plan->opt_frames = value != 0;
            if (value > 1) plan->opt_frames_max = value;
            return S2ST_OK;
        case S2ST_OPT_INVERSE_MEL:
            if (value < 0 || value > 1) break;
            plan->opt_inverse_mel_simt = value;
            return S2ST_OK;
        case S2ST_OPT_FRONTEND_GENERIC:
            plan->opt_frontend_generic = value != 0;
            return S2ST_OK;
        case S2ST_OPT_MEL_PROJECT:
            plan->opt_mel_simt = value != 0;
            return S2ST_OK;
        default:
            break;
    }
    set_error("s2st_plan_set_option: unknown option %d or bad value %d", option, value);
    return S2ST_EINVAL;
}

int s2st_plan_set_pass_timing(s2st_plan* plan, int enabled) {
    if (!plan) {
        set_error("null plan");
        return S2ST_EINVAL;
    }
    plan->timing_enabled = enabled != 0;
    plan->timing_recorded = 0;
    return S2ST_OK;
}

int s2st_plan_get_pass_times(s2st_plan* plan, float* ms_out_host, int capacity, int* n_passes_out) {
    if (!plan || !ms_out_host || !n_passes_out) {
        set_error("null argument");
        return S2ST_EINVAL;
    }
    const int n = plan->timing_recorded - 1;
    if (n <= 0 || n > capacity) {
        set_error("no pass timing recorded (enable it before the call) or capacity %d too small for %d", capacity, n);
        return S2ST_EINVAL;
    }
    S2ST_CUDA_CHECK(cudaEventSynchronize(plan->timing_events[n]));
    for (int i = 0; i < n; ++i)
        S2ST_CUDA_CHECK(cudaEventElapsedTime(ms_out_host + i, plan->timing_events[i], plan->timing_events[i + 1]));
    *n_passes_out = n;
    return S2ST_OK;
}

static int check_gl_geometry(const s2st_plan* plan) {
    if (plan->generic) {
        set_error("Griffin-Lim synthesis / inverse STFT need n_fft = %d (this plan has n_fft = %d: STFT, log-mel and mel "
                  "projection only)", kNfft, plan->n_fft);
        return S2ST_EINVAL;
    }
    if (plan->nphase > 64) {
        set_error("window support %d exceeds %d hops of %d samples: overlap factor not supported", plan->ws,
                  64, plan->hop);
        return S2ST_EINVAL;
    }
    return S2ST_OK;
}

int s2st_gl_workspace_bytes(const s2st_plan* plan, int n_utts, int64_t total_frames, size_t* bytes_out) {
    if (!plan || !bytes_out || n_utts <= 0 || total_frames < n_utts) {
        set_error("bad argument to s2st_gl_workspace_bytes");
        return S2ST_EINVAL;
    }
    *bytes_out = gl_workspace_bytes(plan, n_utts, total_frames);
    return S2ST_OK;
}

int s2st_phase_from_uniform(int n_batch, int n_bins, int n_frames, const double* uniform_dev, float* phase_out_dev,
                             void* stream) {
    if (n_batch < 0 || n_bins < 0 || n_frames < 0 || ((!uniform_dev || !phase_out_dev) && n_batch * n_bins * n_frames > 0)) {
        set_error("bad argument to s2st_phase_from_uniform");
        return S2ST_EINVAL;
    }
    return launch_phase_from_uniform(n_batch, n_bins, n_frames, uniform_dev, phase_out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_phase_from_mt19937(int n_batch, int n_bins, int n_frames, const uint32_t* key_dev, int pos, uint32_t* words_ws_dev,
                            float* phase_out_dev, uint32_t* key_out_dev, void* stream) {
    const long long n = (long long)n_batch * n_bins * n_frames;
    if (n_batch < 0 || n_bins < 0 || n_frames < 0 || pos < 0 || pos > 624 || !key_dev || !key_out_dev ||
        (n > 0 && (!words_ws_dev || !phase_out_dev)) || (reinterpret_cast<uintptr_t>(words_ws_dev) & 7)) {
        set_error("bad argument to s2st_phase_from_mt19937");
        return S2ST_EINVAL;
    }
    return launch_phase_from_mt19937(n_batch, n_bins, n_frames, key_dev, pos, words_ws_dev, phase_out_dev, key_out_dev,
                                     static_cast<cudaStream_t>(stream));
}

int s2st_gl_synthesize(s2st_plan* plan, int n_utts, int64_t total_frames,
                       const int32_t* frame_offsets_dev, const int32_t* frame_offsets_host, const float* logmel_dev,
                       const float* mag_dev, const float* init_phase_dev, uint64_t phase_seed, int n_iter,
                       float* wave_out_dev, void* workspace_dev, size_t workspace_bytes,
                       void* stream) {
    if (!plan || !frame_offsets_dev || !wave_out_dev || n_iter < 0 ||
        ((logmel_dev == nullptr) == (mag_dev == nullptr))) {
        set_error("bad argument to s2st_gl_synthesize (exactly one of logmel / mag must be given)");
        return S2ST_EINVAL;
    }
    int rc = check_gl_geometry(plan);
    if (rc != S2ST_OK) return rc;
    return gl_run(plan, n_utts, total_frames, frame_offsets_dev, frame_offsets_host, logmel_dev, mag_dev, kBins,
                  init_phase_dev, phase_seed, n_iter, wave_out_dev, workspace_dev, workspace_bytes,
                  static_cast<cudaStream_t>(stream));
}

int s2st_gl_launch_count(const s2st_plan* plan, int n_iter, int from_logmel, int* launches_out) {
    if (!plan || !launches_out || n_iter < 0) {
        set_error("bad argument");
        return S2ST_EINVAL;
    }
    // What the last synthesis call of this plan launched, if there was one; else the per-pass count: build_tiles +
    // [inverse_mel] + (n_iter + 1) passes (plus cudaMemsetAsync calls, not kernels of ours).
    *launches_out = plan->last_launches > 0 ? plan->last_launches : 1 + (from_logmel ? 1 : 0) + (n_iter + 1);
    return S2ST_OK;
}

int s2st_inverse_mel(const s2st_plan* plan, int64_t n_frames, const float* mel_dev, int input_is_log,
                     float* mag_dev, void* stream) {
    if (!plan || !mel_dev || !mag_dev || n_frames < 0) {
        set_error("bad argument to s2st_inverse_mel");
        return S2ST_EINVAL;
    }
    if (plan->generic) {
        set_error("s2st_inverse_mel needs n_fft = %d (plan has %d)", kNfft, plan->n_fft);
        return S2ST_EINVAL;
    }
    return launch_inverse_mel(plan, n_frames, mel_dev, input_is_log != 0, mag_dev, kBins, kBins,
                              static_cast<cudaStream_t>(stream));
}

int s2st_mel_project(const s2st_plan* plan, int64_t n_frames, const float* spec_dev, float* mel_out_dev,
                     void* stream) {
    if (!plan || !spec_dev || !mel_out_dev || n_frames < 0) {
        set_error("bad argument to s2st_mel_project");
        return S2ST_EINVAL;
    }
    return launch_mel_project(plan, n_frames, spec_dev, mel_out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_stft(const s2st_plan* plan, int n_utts, int64_t total_frames, const int64_t* wave_offsets_dev,
              const int32_t* frame_offsets_dev, const float* wave_dev, float* mag_out_dev,
              float* phase_out_dev, void* stream) {
    if (!plan || !wave_offsets_dev || !frame_offsets_dev || !wave_dev || !mag_out_dev || n_utts <= 0) {
        set_error("bad argument to s2st_stft");
        return S2ST_EINVAL;
    }
    return launch_stft(plan, n_utts, total_frames, wave_offsets_dev, frame_offsets_dev, wave_dev, mag_out_dev,
                       phase_out_dev, nullptr, 0.0f, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int s2st_istft(s2st_plan* plan, int n_utts, int64_t total_frames, const int32_t* frame_offsets_dev,
               const float* mag_dev, const float* phase_dev, float* wave_out_dev, void* workspace_dev,
               size_t workspace_bytes, void* stream) {
    if (!plan || !frame_offsets_dev || !mag_dev || !phase_dev || !wave_out_dev) {
        set_error("bad argument to s2st_istft");
        return S2ST_EINVAL;
    }
    int rc = check_gl_geometry(plan);
    if (rc != S2ST_OK) return rc;
    return gl_run(plan, n_utts, total_frames, frame_offsets_dev, nullptr, nullptr, mag_dev, kBins, phase_dev, 0ull, 0,
                  wave_out_dev, workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int s2st_rfft2048(const s2st_plan* plan, int64_t n, const float* in_dev, float* out_dev, void* stream) {
    if (!plan || !in_dev || !out_dev || n < 0) {
        set_error("bad argument to s2st_rfft2048");
        return S2ST_EINVAL;
    }
    if (plan->generic) {
        set_error("s2st_rfft2048 needs a plan with n_fft = %d", kNfft);
        return S2ST_EINVAL;
    }
    return launch_rfft2048(plan, n, in_dev, out_dev, false, static_cast<cudaStream_t>(stream));
}

int s2st_irfft2048(const s2st_plan* plan, int64_t n, const float* in_dev, float* out_dev, void* stream) {
    if (!plan || !in_dev || !out_dev || n < 0) {
        set_error("bad argument to s2st_irfft2048");
        return S2ST_EINVAL;
    }
    if (plan->generic) {
        set_error("s2st_irfft2048 needs a plan with n_fft = %d", kNfft);
        return S2ST_EINVAL;
    }
    return launch_rfft2048(plan, n, in_dev, out_dev, true, static_cast<cudaStream_t>(stream));
}

int s2st_window_sum_square(int n_frames, int hop_length, int win_length, int n_fft, const float* window_host,
                           float* out_host) {
    if (n_frames < 1 || hop_length < 1 || win_length < 1 || win_length > n_fft || !window_host || !out_host) {
        set_error("bad argument to s2st_window_sum_square");
        return S2ST_EINVAL;
    }
    std::vector<float> w2(n_fft, 0.0f);
    const int pl = (n_fft - win_length) / 2;
    for (int i = 0; i < win_length; ++i) w2[pl + i] = window_host[i] * window_host[i];
    const long long n = (long long)n_fft + (long long)hop_length * (n_frames - 1);
    for (long long i = 0; i < n; ++i) out_host[i] = 0.0f;
    for (int t = 0; t < n_frames; ++t) {
        const long long o = (long long)t * hop_length;
        for (int i = 0; i < n_fft && o + i < n; ++i) out_host[o + i] += w2[i];
    }
    return S2ST_OK;
}

int s2st_logmel(const s2st_plan* plan, int n_utts, int64_t total_frames, const int64_t* wave_offsets_dev,
                const int32_t* frame_offsets_dev, const float* wave_dev, float eps,
                const float* cmvn_mean_dev, const float* cmvn_std_dev, double* sums_dev, float* out_dev, void* stream) {
    if (!plan || !wave_offsets_dev || !frame_offsets_dev || !wave_dev || !out_dev || n_utts <= 0 ||
        ((cmvn_mean_dev == nullptr) != (cmvn_std_dev == nullptr))) {
        set_error("bad argument to s2st_logmel");
        return S2ST_EINVAL;
    }
    return launch_stft(plan, n_utts, total_frames, wave_offsets_dev, frame_offsets_dev, wave_dev, nullptr, nullptr,
                       out_dev, eps, cmvn_mean_dev, cmvn_std_dev, static_cast<cudaStream_t>(stream), sums_dev);
}

int s2st_fbank_plan_create(s2st_fbank_plan** plan_out, int device, int sample_rate, int n_bins) {
    if (!plan_out || sample_rate <= 0 || n_bins < 4) {
        set_error("bad argument to s2st_fbank_plan_create");
        return S2ST_EINVAL;
    }
    *plan_out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) {
        set_error("cannot select CUDA device %d", device);
        return S2ST_ECUDA;
    }
    // kaldi.py _get_waveform_and_window_properties: python float arithmetic, truncated
    const double sf = (double)sample_rate;
    const int shift = (int)(sf * 10.0 * 0.001);
    const int win = (int)(sf * 25.0 * 0.001);
    int padded = 1;
    while (padded < win) padded <<= 1;
    if (win < 2 || shift < 1 || padded > 4096 || padded < 64) {
        set_error("unsupported sample rate %d (window %d, padded %d)", sample_rate, win, padded);
        return S2ST_EINVAL;
    }
    s2st_fbank_plan* p = new s2st_fbank_plan();
    std::memset(p, 0, sizeof(*p));
    p->opt_generic = getenv("S2ST_FBANK_GENERIC") ? 1 : 0;  // read once, here
    p->device = device;
    p->sample_rate = sample_rate;
    p->n_bins = n_bins;
    p->win = win;
    p->shift = shift;
    p->padded = padded;
    const double pi = 3.14159265358979323846;
    std::vector<float> window(win);
    for (int i = 0; i < win; ++i) {
        const float hann = (float)(0.5 - 0.5 * std::cos(2.0 * pi * (double)i / (double)(win - 1)));
        window[i] = std::pow(hann, 0.85f);
    }
    std::vector<float2> tw(padded / 2);
    for (int j = 0; j < padded / 2; ++j) {
        const double a = -2.0 * pi * (double)j / (double)padded;
        tw[j] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    // kaldi.py get_mel_banks (no VTLN), evaluated in float32 like torchaudio does
    const int nfb = padded / 2;
    const double nyquist = 0.5 * sf;
    const double low = 20.0, high = nyquist;
    const double bin_width = sf / (double)padded;
    const double mel_low = 1127.0 * std::log(1.0 + low / 700.0);
    const double mel_high = 1127.0 * std::log(1.0 + high / 700.0);
    const double delta = (mel_high - mel_low) / (double)(n_bins + 1);
    std::vector<float> dense((size_t)n_bins * (nfb + 1), 0.0f);
    for (int b = 0; b < n_bins; ++b) {
        const float left = (float)mel_low + (float)b * (float)delta;
        const float center = (float)mel_low + ((float)b + 1.0f) * (float)delta;
        const float right = (float)mel_low + ((float)b + 2.0f) * (float)delta;
        for (int k = 0; k < nfb; ++k) {
            const float freq = (float)bin_width * (float)k;
            const float mel = 1127.0f * std::log(1.0f + freq / 700.0f);
            const float up = (mel - left) / (center - left);
            const float down = (right - mel) / (right - center);
            const float v = std::fmax(0.0f, std::fmin(up, down));
            dense[(size_t)b * (nfb + 1) + k] = v;
        }
    }
    std::vector<int> ptr, idx;
    std::vector<float> val;
    int max_row = 0;
    to_csr(dense.data(), n_bins, nfb + 1, ptr, idx, val, &max_row);
    p->mel_nnz = (int)idx.size();
    if (idx.empty()) {
        idx.push_back(0);
        val.push_back(0.0f);
    }
    int rc = S2ST_OK;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        rc = S2ST_ECUDA;
    } else {
        p->num_sms = prop.multiProcessorCount;
    }
    // tables of the register-resident kernel (FFT 512 / 256 with at most 13 * 16 window samples per transform)
    p->fast_mode = -1;
    if (padded == 512 && win <= 13 * 32) p->fast_mode = 0;
    if (padded == 256 && win <= 13 * 16) p->fast_mode = 1;
    if (p->fast_mode >= 0) {
        std::vector<float2> tw16(256), vsplit(256);
        for (int k1 = 0; k1 < 16; ++k1)
            for (int n2 = 0; n2 < 16; ++n2) {
                const double a = -2.0 * pi * (double)(k1 * n2) / 256.0;
                tw16[k1 * 16 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
            }
        for (int k = 0; k < 256; ++k) {
            const double a = 2.0 * pi * (double)k / 512.0;  // -i * exp(-i a) = (-sin a, -cos a)
            vsplit[k] = make_float2((float)(-std::sin(a)), (float)(-std::cos(a)));
        }
        vsplit[0] = make_float2(0.0f, -1.0f);
        vsplit[128] = make_float2(-1.0f, 0.0f);
        const int per_row = p->fast_mode == 0 ? 32 : 16;
        std::vector<float> winp(13 * per_row, 0.0f);
        for (int i = 0; i < win; ++i) winp[i] = window[i];
        // column view of the mel bank: every FFT bin must feed at most two adjacent mel bins (true for Kaldi's
        // pairwise-overlapping triangles); the table is stored in the order the kernel's sub-lanes read it
        const int kpl = p->fast_mode == 0 ? 16 : 8;  // FFT bins per sub-lane
        std::vector<float4> col(256, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        bool ok = n_bins + 1 <= 4096;
        for (int k = 0; k < nfb && ok; ++k) {
            int first = -1, count = 0, last = -1;
            for (int b = 0; b < n_bins; ++b)
                if (dense[(size_t)b * (nfb + 1) + k] != 0.0f) {
                    if (first < 0) first = b;
                    last = b;
                    ++count;
                }
            if (count > 2 || (count == 2 && last != first + 1)) ok = false;
            const int b = first < 0 ? 0 : first;
            const float w0 = first < 0 ? 0.0f : dense[(size_t)b * (nfb + 1) + k];
            const float w1 = (count == 2) ? dense[(size_t)(b + 1) * (nfb + 1) + k] : 0.0f;
            int bi = b;
            float bf;
            std::memcpy(&bf, &bi, sizeof(float));
            col[(k % kpl) * 16 + k / kpl] = make_float4(w0, w1, bf, 0.0f);
        }
        if (ok && p->fast_mode == 0) {
            // atomic-free accumulation (same scheme as the log-mel kernel, s2st_plan::mel_col): a sub-lane keeps one pair
            // of running sums per run of its bins that feed the same mel bin b and stores them to the run's slot of a
            // per-half-warp slab after every step; entry = (w into b, w into b + 1, 1 if the run continues else 0, slot).
            // mel_gather[m * 8 + q] lists the slab floats that add up to mel bin m.
            constexpr int max_rows = 13, max_terms = 8;  // kFbSlabRows in frontend_kernels.cu
            std::vector<std::vector<int>> terms(n_bins);
            for (int sl = 0; sl < 16 && ok; ++sl) {
                int cur = -1, run = -1, slot = 0;
                for (int j = 0; j < kpl && ok; ++j) {
                    float4& e = col[j * 16 + sl];
                    const bool empty = e.x == 0.0f && e.y == 0.0f;
                    int b;
                    std::memcpy(&b, &e.z, sizeof(int));
                    if (empty) b = cur < 0 ? 0 : cur;  // an empty column continues the current run
                    float w0 = e.x, w1 = e.y;
                    if (!empty && e.y == 0.0f && cur >= 0 && b == cur + 1) {  // only the upper bin of the run: stay in it
                        b = cur;
                        w1 = w0;
                        w0 = 0.0f;
                    }
                    const bool keep = b == cur;
                    if (!keep) {
                        ++run;
                        cur = b;
                        if (run >= max_rows) {
                            ok = false;
                            break;
                        }
                        // slab[run][sub-lane]: at every step the 16 sub-lanes store to 16 different bank pairs
                        slot = run * 16 + sl;
                        terms[b].push_back(slot * 2);
                        if (b + 1 < n_bins) terms[b + 1].push_back(slot * 2 + 1);
                    }
                    float sf;
                    std::memcpy(&sf, &slot, sizeof(float));
                    e = make_float4(w0, w1, keep ? 1.0f : 0.0f, sf);
                }
            }
            const int zero_float = 2 * 16 * max_rows;
            int n_terms = 1;
            for (int m = 0; m < n_bins && ok; ++m) {
                if ((int)terms[m].size() > max_terms) ok = false;
                n_terms = std::max(n_terms, (int)terms[m].size());
            }
            if (ok) {
                std::vector<int> gather((size_t)n_bins * max_terms, zero_float);  // unused entries -> a float kept at zero
                for (int m = 0; m < n_bins; ++m)  // two planes of int4 per mel bin, as in the log-mel plan
                    for (size_t q = 0; q < terms[m].size(); ++q) gather[((q / 4) * n_bins + m) * 4 + q % 4] = terms[m][q];
                p->mel_terms = n_terms;
                p->mel_zero = zero_float;
                if (rc == S2ST_OK) rc = upload(&p->mel_gather, gather);
            }
        }
        if (!ok) p->fast_mode = -1;
        if (rc == S2ST_OK) rc = upload(&p->tw16, tw16);
        if (rc == S2ST_OK) rc = upload(&p->vsplit, vsplit);
        if (rc == S2ST_OK) rc = upload(&p->winp, winp);
        if (rc == S2ST_OK) rc = upload(&p->mel_col, col);
    }
    if (rc == S2ST_OK) rc = upload(&p->window, window);
    if (rc == S2ST_OK) rc = upload(&p->tw, tw);
    if (rc == S2ST_OK) rc = upload(&p->mel_ptr, ptr);
    if (rc == S2ST_OK) rc = upload(&p->mel_idx, idx);
    if (rc == S2ST_OK) rc = upload(&p->mel_val, val);
    if (rc != S2ST_OK) {
        s2st_fbank_plan_destroy(p);
        return rc;
    }
    *plan_out = p;
    return S2ST_OK;
}

int s2st_fbank_plan_destroy(s2st_fbank_plan* plan) {
    if (!plan) return S2ST_OK;
    DeviceGuard guard(plan->device);
    cudaFree(plan->window);
    cudaFree(plan->tw);
    cudaFree(plan->tw16);
    cudaFree(plan->vsplit);
    cudaFree(plan->winp);
    cudaFree(plan->mel_col);
    cudaFree(plan->mel_gather);
    cudaFree(plan->mel_ptr);
    cudaFree(plan->mel_idx);
    cudaFree(plan->mel_val);
    delete plan;
    return S2ST_OK;
}

int s2st_fbank_plan_set_option(s2st_fbank_plan* plan, int option, int value) {
    if (!plan || option != S2ST_OPT_FRONTEND_GENERIC) {
        set_error("s2st_fbank_plan_set_option: null plan or unknown option %d", option);
        return S2ST_EINVAL;
    }
    plan->opt_generic = value != 0;
    return S2ST_OK;
}

int s2st_fbank_frame_params(const s2st_fbank_plan* plan, int* win_out, int* shift_out, int* padded_out) {
    if (!plan) {
        set_error("null plan");
        return S2ST_EINVAL;
    }
    if (win_out) *win_out = plan->win;
    if (shift_out) *shift_out = plan->shift;
    if (padded_out) *padded_out = plan->padded;
    return S2ST_OK;
}

int s2st_fbank(const s2st_fbank_plan* plan, int n_utts, int64_t total_frames, const int64_t* wave_offsets_dev,
               const int32_t* frame_offsets_dev, const float* wave_dev, const float* cmvn_mean_dev,
               const float* cmvn_std_dev, double* sums_dev, float* out_dev, void* stream) {
    if (!plan || !wave_offsets_dev || !frame_offsets_dev || !wave_dev || !out_dev || n_utts <= 0 ||
        ((cmvn_mean_dev == nullptr) != (cmvn_std_dev == nullptr))) {
        set_error("bad argument to s2st_fbank");
        return S2ST_EINVAL;
    }
    return launch_fbank(plan, n_utts, total_frames, wave_offsets_dev, frame_offsets_dev, wave_dev, cmvn_mean_dev,
                        cmvn_std_dev, out_dev, static_cast<cudaStream_t>(stream), sums_dev);
}

int s2st_cmvn_apply(int64_t n_rows, int n_cols, const float* x_dev, const float* mean_dev, const float* std_dev,
                    float* out_dev, void* stream) {
    if (!x_dev || !mean_dev || !std_dev || !out_dev || n_rows < 0 || n_cols < 1) {
        set_error("bad argument to s2st_cmvn_apply");
        return S2ST_EINVAL;
    }
    return launch_cmvn(n_rows, n_cols, x_dev, mean_dev, std_dev, out_dev, false, static_cast<cudaStream_t>(stream));
}

int s2st_cmvn_denormalize(int64_t n_rows, int n_cols, const float* x_dev, const float* mean_dev,
                          const float* std_dev, float* out_dev, void* stream) {
    if (!x_dev || !mean_dev || !std_dev || !out_dev || n_rows < 0 || n_cols < 1) {
        set_error("bad argument to s2st_cmvn_denormalize");
        return S2ST_EINVAL;
    }
    return launch_cmvn(n_rows, n_cols, x_dev, mean_dev, std_dev, out_dev, true, static_cast<cudaStream_t>(stream));
}

int s2st_cmvn_accumulate(int64_t n_rows, int n_cols, const float* x_dev, double* sums_dev, void* stream) {
    if (!x_dev || !sums_dev || n_rows < 0 || n_cols < 1) {
        set_error("bad argument to s2st_cmvn_accumulate");
        return S2ST_EINVAL;
    }
    return launch_cmvn_accumulate(n_rows, n_cols, x_dev, sums_dev, static_cast<cudaStream_t>(stream));
}

int s2st_utterance_cmvn(int n_utts, int64_t total_rows, const int32_t* frame_offsets_dev, int n_cols, const float* x_dev,
                        int norm_means, int norm_vars, float* out_dev, float* stats_dev, void* stream) {
    if (n_utts < 0 || total_rows < 0 || n_cols < 1 ||
        (n_utts > 0 && total_rows > 0 && (!frame_offsets_dev || !x_dev || !out_dev || !stats_dev))) {
        set_error("bad argument to s2st_utterance_cmvn");
        return S2ST_EINVAL;
    }
    return launch_utterance_cmvn(n_utts, total_rows, frame_offsets_dev, n_cols, x_dev, out_dev, norm_means != 0,
                                 norm_vars != 0, stats_dev, static_cast<cudaStream_t>(stream));
}

int s2st_utterance_sums(int n_utts, const int32_t* frame_offsets_dev, int n_cols, const float* x_dev, float* sums_dev,
                        void* stream) {
    if (n_utts < 0 || n_cols <= 0 || (n_utts > 0 && (!frame_offsets_dev || !x_dev || !sums_dev))) {
        set_error("bad argument to s2st_utterance_sums");
        return S2ST_EINVAL;
    }
    return launch_utterance_sums(n_utts, frame_offsets_dev, n_cols, x_dev, sums_dev, static_cast<cudaStream_t>(stream));
}

int s2st_utterance_sum(int n_utts, const int32_t* frame_offsets_dev, int n_cols, const float* x_dev, double* sums_dev,
                       void* stream) {
    if (n_utts < 0 || n_cols < 1 || (n_utts > 0 && (!frame_offsets_dev || !x_dev || !sums_dev))) {
        set_error("bad argument to s2st_utterance_sum");
        return S2ST_EINVAL;
    }
    return launch_utterance_sum(n_utts, frame_offsets_dev, n_cols, x_dev, sums_dev, static_cast<cudaStream_t>(stream));
}

int s2st_fill_rects(int n_rects, const int32_t* rects_dev, const float* values_dev, int n_cols, float* x_dev,
                    void* stream) {
    if (n_rects < 0 || n_cols < 1 || (n_rects > 0 && (!rects_dev || !values_dev || !x_dev))) {
        set_error("bad argument to s2st_fill_rects");
        return S2ST_EINVAL;
    }
    if (((uintptr_t)rects_dev & 15) != 0) {
        set_error("s2st_fill_rects: rects_dev must be 16-byte aligned");
        return S2ST_EINVAL;
    }
    return launch_fill_rects(n_rects, rects_dev, values_dev, n_cols, x_dev, static_cast<cudaStream_t>(stream));
}

int s2st_dtw(int bsz, int m, int n, const float* distance_dev, const int64_t* shapes_dev, float* cumdist_dev,
             int32_t* backptr_dev, int32_t* pathmap_dev, void* stream) {
    if (bsz < 0 || m < 0 || n < 0 || (bsz > 0 && m > 0 && n > 0 && (!distance_dev || !cumdist_dev || !backptr_dev || !pathmap_dev))) {
        set_error("bad argument to s2st_dtw");
        return S2ST_EINVAL;
    }
    return launch_dtw(bsz, m, n, distance_dev, reinterpret_cast<const long long*>(shapes_dev), cumdist_dev, backptr_dev,
                      pathmap_dev, static_cast<cudaStream_t>(stream));
}

int s2st_rms_dist(int m, int n, int d, const float* x1_dev, const float* x2_dev, float* out_dev, void* stream) {
    if (m < 0 || n < 0 || d < 1 || (m > 0 && n > 0 && (!x1_dev || !x2_dev || !out_dev))) {
        set_error("bad argument to s2st_rms_dist");
        return S2ST_EINVAL;
    }
    return launch_rms_dist(m, n, d, x1_dev, x2_dev, out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_time_warp(int n_utts, int64_t total_rows, const int32_t* frame_offsets_dev, int n_cols, const int32_t* warp_dev,
                   int arithmetic, const float* x_dev, float* out_dev, void* stream) {
    if (n_utts < 0 || total_rows < 0 || n_cols <= 0 || x_dev == out_dev || (arithmetic != 0 && arithmetic != 1) ||
        (n_utts > 0 && (!frame_offsets_dev || !warp_dev || !x_dev || !out_dev))) {
        set_error("bad argument to s2st_time_warp (in place is not supported)");
        return S2ST_EINVAL;
    }
    return launch_time_warp(n_utts, total_rows, frame_offsets_dev, n_cols, warp_dev, arithmetic, x_dev, out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_pcm16_to_wave(int64_t n_samples, const int16_t* pcm_dev, float scale, float* wave_out_dev, void* stream) {
    if (n_samples < 0 || (n_samples > 0 && (!pcm_dev || !wave_out_dev))) {
        set_error("bad argument to s2st_pcm16_to_wave");
        return S2ST_EINVAL;
    }
    return launch_pcm16_to_wave(n_samples, pcm_dev, scale, wave_out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_wave_to_pcm16(int64_t n_samples, const float* wave_dev, int16_t* pcm_out_dev, void* stream) {
    if (n_samples < 0 || (n_samples > 0 && (!wave_dev || !pcm_out_dev))) {
        set_error("bad argument to s2st_wave_to_pcm16");
        return S2ST_EINVAL;
    }
    return launch_wave_to_pcm16(n_samples, wave_dev, pcm_out_dev, static_cast<cudaStream_t>(stream));
}

int s2st_rms_dist_batch(int bsz, int max_m, int max_n, int d, const float* x1_dev, const float* x2_dev,
                        const int32_t* offsets1_dev, const int32_t* offsets2_dev, float* out_dev, void* stream) {
    if (bsz < 0 || max_m < 0 || max_n < 0 || d <= 0 ||
        (bsz > 0 && (!x1_dev || !x2_dev || !offsets1_dev || !offsets2_dev || !out_dev))) {
        set_error("bad argument to s2st_rms_dist_batch");
        return S2ST_EINVAL;
    }
    return launch_rms_dist_batch(bsz, max_m, max_n, d, x1_dev, x2_dev, offsets1_dev, offsets2_dev, out_dev,
                                 static_cast<cudaStream_t>(stream));
}

}  // extern "C"
