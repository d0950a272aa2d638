// mshds_api.cu -- C ABI (include/mshds_b200.h), handle / scratch management and the per-chunk orchestration that mirrors
// the body of extract_mshds_features (src/mshds_extractor.py:408-448), stage by stage.
#include "../../include/mshds_b200.h"
#include "internal.h"
#include "common.cuh"
#include "num.cuh"

#include <nvtx3/nvToolsExt.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>
#include <utility>
#include <algorithm>

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char b_[512];                                                                             \
            snprintf(b_, sizeof b_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            h->err = b_;                                                                              \
            return MSHDS_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

struct DebugEntry {
    const void* base;
    const int* start;       // device prefix array (per clip) or nullptr => start = clip * fixed
    const int* count;       // device per-clip count or nullptr => fixed
    int fixed;
    int elem_size;
    int stride;             // elements per item (e.g. 15 for candidates)
    std::vector<long long> host_prefix;   // optional host-side segment offsets (n+1) instead of start / count
};

struct mshds_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // latency-bound single-warp-per-clip kernels (Viterbi, pulse walks, interval logic) are issued on a side stream so that
    // they run underneath the frame kernels of the next analysis; `cur` is the stream work is being issued on right now
    cudaStream_t side[4] = {}, cur = nullptr, copy = nullptr;
    cudaEvent_t ev[16] = {};
    bool overlap = true;
    std::string err;
    long long launches = 0;
    long long chunk_samples = 1LL << 27;
    bool chunk_auto = true;             // no mshds_set_chunk_samples call yet: long recordings may get larger chunks (mshds_extract)
    // persistent tables
    double2* tw = nullptr;
    double2* twb512 = nullptr;          // pass twiddles of the warp-resident transforms (fftreg.cuh), [32][M / 32]
    double2* twb1024 = nullptr;
    int legacy_fft = 0;
    int legacy_cc = 0;
    std::map<int, std::pair<double*, double*>> ac_windows;      // nsamp_window -> (window, windowR)
    std::map<long long, double*> kaiser;                        // key -> window
    std::map<int, double*> gauss_spec;
    std::map<int, double*> gauss_formant;
    // growable scratch for the CPP stage (sized after the voiced-segment list is known)
    char* cpp_buf = nullptr;
    size_t cpp_cap = 0;
    // scratch of the resample-to-16-kHz front-end
    char* front_buf = nullptr;
    size_t front_cap = 0;
    // per-frame contours of the chunk processed last (sources on the device) and the sink mshds_extract_contours fills
    struct ContourSrc { const double* a; const double* b; const int* nform; const int* fstart; const int* nF; const double* t1; double dt; };
    ContourSrc csrc[5] = {};
    struct ContourSink { int which; double* values; size_t cap_rows; int64_t* frame_offsets; double* t1; long long rows; bool overflow; };
    ContourSink* sink = nullptr;
    // scratch of the frame-level descriptor path (mshds_lld_extract)
    char* lld_buf = nullptr;
    size_t lld_cap = 0;
    // scratch of mshds_aggregate_sessions (grow-only)
    char* agg_buf = nullptr;
    size_t agg_cap = 0;
    // development / test switches (mshds_set_option)
    int nvtx = 0;
    // arena
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    bool arena_overflow = false;
    std::map<std::string, DebugEntry> debug;
    int last_n = 0;
    // optional per-stage timing with CUDA events on the handle's stream
    bool prof_on = false;
    struct ProfSpan { std::string name; cudaEvent_t a, b; cudaStream_t st; };
    std::vector<ProfSpan> prof_open;
    std::vector<std::string> prof_order;
    std::map<std::string, std::pair<double, long long>> prof_acc;    // name -> (ms, spans)
};

static void prof_begin(mshds_handle* h, const char* name) {
    if (h->nvtx) nvtxRangePushA(name);          // NVTX range per stage (host-side issue span; "nvtx" option)
    if (!h->prof_on) return;
    mshds_handle::ProfSpan sp;
    sp.name = name;
    cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, h->cur);
    sp.st = h->cur;
    h->prof_open.push_back(sp);
}
static void prof_end(mshds_handle* h) {
    if (h->nvtx) nvtxRangePop();
    if (!h->prof_on || h->prof_open.empty()) return;
    // close the most recent span that has not been closed yet
    for (size_t i = h->prof_open.size(); i-- > 0;) {
        if ((h->prof_open[i].name.empty() || h->prof_open[i].name[0] != '\x01') && h->prof_open[i].st == h->cur) {
            cudaEventRecord(h->prof_open[i].b, h->prof_open[i].st);
            h->prof_open[i].name.insert(h->prof_open[i].name.begin(), '\x01');
            return;
        }
    }
}
static void prof_collect(mshds_handle* h) {
    for (auto& sp : h->prof_open) {
        std::string name = sp.name;
        bool closed = !name.empty() && name[0] == '\x01';
        if (closed) {
            name.erase(name.begin());
            float ms = 0.f;
            if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
                if (!h->prof_acc.count(name)) h->prof_order.push_back(name);
                auto& acc = h->prof_acc[name];
                acc.first += ms; acc.second += 1;
            }
        }
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    h->prof_open.clear();
}
#define PB(name) prof_begin(h, name)
#define PE() prof_end(h)


// ------------------------------------------------------------------------------------------------ arena
static void* arena_take(mshds_handle* h, size_t bytes) {
    size_t o = (h->arena_off + 255) & ~(size_t)255;
    if (o + bytes > h->arena_cap) { h->arena_overflow = true; h->arena_off = o + bytes; return nullptr; }
    h->arena_off = o + bytes;
    return h->arena + o;
}
template <class T>
static T* take(mshds_handle* h, size_t n) { return (T*)arena_take(h, sizeof(T) * (n ? n : 1)); }

// ------------------------------------------------------------------------------------------------ host FFT (set-up only)
static void host_fft(std::vector<double>& re, std::vector<double>& im, int sign) {
    size_t n = re.size();
    for (size_t i = 0, j = 0; i < n - 1; i++) {
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
        size_t m = n >> 1;
        while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
        j |= m;
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1;
        for (size_t k = 0; k < half; k++) {
            double ang = sign * 2.0 * MSHDS_PI * (double)k / (double)len;
            double wr = cos(ang), wi = sin(ang);
            for (size_t i = k; i < n; i += len) {
                size_t b = i + half;
                double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[i] - xr; im[b] = im[i] - xi;
                re[i] += xr; im[i] += xi;
            }
        }
    }
}

static int upload(mshds_handle* h, const std::vector<double>& v, double** out) {
    CK(cudaMalloc((void**)out, sizeof(double) * v.size()));
    CK(cudaMemcpy(*out, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
    return MSHDS_OK;
}

// ------------------------------------------------------------------------------------------------ configurations
// fon/Sound_to_Pitch.cpp Sound_to_Pitch_any: everything before the frame loop.
static int make_pitch_cfg(mshds_handle* h, double fs, double dt, double floor_hz, double ppw, int maxn, int method,
                          double sil, double vt, double oct, double jump, double vuv, double ceiling, PitchCfg* g) {
    double dx = 1.0 / fs;
    memset(g, 0, sizeof(*g));
    if (maxn < ceiling / floor_hz) maxn = (int)floor(ceiling / floor_hz);
    if (dt <= 0.0) dt = ppw / floor_hz / 4.0;
    g->floor_hz = floor_hz; g->dt = dt; g->ppw = ppw; g->method = method;
    g->sil = sil; g->vt = vt; g->octave_cost = oct; g->jump_cost = jump; g->vuv_cost = vuv;
    g->nsamp_period = (int)floor(1.0 / dx / floor_hz);
    g->halfnsamp_period = g->nsamp_period / 2 + 1;
    if (ceiling > 0.5 / dx) ceiling = 0.5 / dx;
    g->ceiling = ceiling;
    g->dt_window = ppw / floor_hz;
    g->nsamp_window = (int)floor(g->dt_window / dx);
    g->halfnsamp_window = g->nsamp_window / 2 - 1;
    g->nsamp_window = g->halfnsamp_window * 2;
    g->maximumLag = (int)floor(g->nsamp_window / ppw) + 2;
    if (g->maximumLag > g->nsamp_window) g->maximumLag = g->nsamp_window;
    g->grid_window = method >= 2 ? 1.0 / floor_hz + g->dt_window : g->dt_window;
    g->maxn = maxn;
    if (method >= 2) {
        g->brent_ixmax = g->nsamp_window;
        return MSHDS_OK;
    }
    int nfft = 1;
    while (nfft < g->nsamp_window * 1.5) nfft *= 2;
    g->nsampFFT = nfft; g->M = nfft / 2;
    g->logM = 0;
    while ((1 << g->logM) < g->M) g->logM++;
    g->brent_ixmax = (int)floor(g->nsamp_window * 0.5);
    auto it = h->ac_windows.find(g->nsamp_window);
    if (it == h->ac_windows.end()) {
        int W = g->nsamp_window;
        std::vector<double> win(W), re(nfft, 0.0), im(nfft, 0.0);
        for (int i = 1; i <= W; i++) win[i - 1] = 0.5 - 0.5 * cos(i * 2 * MSHDS_PI / (W + 1));
        for (int i = 0; i < W; i++) re[i] = win[i];
        host_fft(re, im, -1);
        for (int i = 0; i < nfft; i++) { re[i] = re[i] * re[i] + im[i] * im[i]; im[i] = 0.0; }
        host_fft(re, im, +1);
        std::vector<double> wr(nfft);
        for (int i = 0; i < nfft; i++) wr[i] = re[i];
        for (int i = 1; i < W; i++) wr[i] /= wr[0];
        wr[0] = 1.0;
        double *dw, *dr;
        int rc = upload(h, win, &dw); if (rc) return rc;
        rc = upload(h, wr, &dr); if (rc) return rc;
        it = h->ac_windows.emplace(W, std::make_pair(dw, dr)).first;
    }
    g->window = it->second.first;
    g->windowR = it->second.second;
    return MSHDS_OK;
}

static int make_kaiser(mshds_handle* h, double fs, double minPitch, int* halfN, const double** win) {
    double dx = 1.0 / fs;
    double halfWindowDuration = 0.5 * (6.4 / minPitch);
    int hn = (int)floor(halfWindowDuration / dx);
    long long key = (long long)llround(minPitch * 1000.0) * 100000 + (long long)llround(fs);
    auto it = h->kaiser.find(key);
    if (it == h->kaiser.end()) {
        std::vector<double> w(2 * hn + 1);
        for (int i = -hn; i <= hn; i++) {
            double x = i * dx / halfWindowDuration;
            double root = 1.0 - x * x;
            w[i + hn] = root <= 0.0 ? 0.0 : bessel_i0_f((2.0 * MSHDS_PI * MSHDS_PI + 0.5) * sqrt(root));
        }
        double* d;
        int rc = upload(h, w, &d); if (rc) return rc;
        it = h->kaiser.emplace(key, d).first;
    }
    *halfN = hn;
    *win = it->second;
    return MSHDS_OK;
}

// fon/Sound_and_Spectrogram.cpp Sound_to_Spectrogram (0.025, 5000, 0.005, 20, GAUSSIAN, 8, 8): set-up section
static int make_spec_cfg(mshds_handle* h, double fs, SpecPass* p) {
    double dx = 1.0 / fs, nyquist = 0.5 / dx;
    double effectiveAnalysisWidth = 0.025, fmax = 5000.0, minimumTimeStep1 = 0.005, minimumFreqStep1 = 20.0;
    memset(p, 0, sizeof(*p));
    double physicalAnalysisWidth = 2 * effectiveAnalysisWidth;
    double effectiveTimeWidth = effectiveAnalysisWidth / sqrt(MSHDS_PI);
    double effectiveFreqWidth = 1 / effectiveTimeWidth;
    double minimumTimeStep2 = effectiveTimeWidth / 8.0, minimumFreqStep2 = effectiveFreqWidth / 8.0;
    double timeStep = minimumTimeStep1 > minimumTimeStep2 ? minimumTimeStep1 : minimumTimeStep2;
    double freqStep = minimumFreqStep1 > minimumFreqStep2 ? minimumFreqStep1 : minimumFreqStep2;
    int nsamp_window = (int)floor(physicalAnalysisWidth / dx);
    int halfnsamp_window = nsamp_window / 2 - 1;
    nsamp_window = halfnsamp_window * 2;
    if (fmax <= 0.0 || fmax > nyquist) fmax = nyquist;
    int numberOfFreqs = (int)floor(fmax / freqStep);
    int nsampFFT = 1;
    while (nsampFFT < nsamp_window || nsampFFT < 2 * numberOfFreqs * (nyquist / fmax)) nsampFFT *= 2;
    int binWidth_samples = (int)floor(freqStep * dx * nsampFFT);
    if (binWidth_samples < 1) binWidth_samples = 1;
    double binWidth_hertz = 1.0 / (dx * nsampFFT);
    freqStep = binWidth_samples * binWidth_hertz;
    numberOfFreqs = (int)floor(fmax / freqStep);
    p->physicalAnalysisWidth = physicalAnalysisWidth; p->timeStep = timeStep; p->freqStep = freqStep;
    p->y1 = 0.5 * (freqStep - binWidth_hertz);
    p->nsamp_window = nsamp_window; p->halfnsamp_window = halfnsamp_window; p->nsampFFT = nsampFFT; p->M = nsampFFT / 2;
    p->logM = 0;
    while ((1 << p->logM) < p->M) p->logM++;
    p->numberOfFreqs = numberOfFreqs; p->binWidth_samples = binWidth_samples;
    std::vector<double> w(nsamp_window);
    double windowssq = 0.0;
    for (int i = 1; i <= nsamp_window; i++) {
        double nSamplesPerWindow_f = physicalAnalysisWidth / dx;
        double imid = 0.5 * (double)(nsamp_window + 1), edge = exp(-12.0);
        double phase = ((double)i - imid) / nSamplesPerWindow_f;
        double value = (exp(-48.0 * phase * phase) - edge) / (1.0 - edge);
        w[i - 1] = (double)(float)value;
        windowssq += value * value;
    }
    p->oneByBinWidth = 1.0 / windowssq / binWidth_samples;
    auto it = h->gauss_spec.find(nsamp_window);
    if (it == h->gauss_spec.end()) {
        double* d;
        int rc = upload(h, w, &d); if (rc) return rc;
        it = h->gauss_spec.emplace(nsamp_window, d).first;
    }
    p->window = it->second;
    return MSHDS_OK;
}

// ------------------------------------------------------------------------------------------------ pass allocation
static long long frames_upper_bound(const std::vector<long long>& lens, double dx, double dt) {
    long long t = 0;
    for (long long n : lens) t += (long long)floor((double)n * dx / dt) + 2;
    return t;
}

struct CandScratchBuf { double *f, *s, *score, *lf, *inten, *rbuf; uint8_t *ncand, *psi; unsigned short* imax; int *queue, *qcount; long long cap; };

static void alloc_pitch_pass(mshds_handle* h, PitchPass* p, int n, long long fub, const CandScratchBuf& cs) {
    p->nF = take<int>(h, n);
    p->turn_counter = take<int>(h, 1);
    p->t1 = take<double>(h, n);
    p->fstart = take<int>(h, n + 1);
    p->sel_f = take<double>(h, fub);
    p->sel_s = take<double>(h, fub);
    p->cand_f = cs.f; p->cand_s = cs.s; p->cand_score = cs.score; p->cand_lf = cs.lf; p->ncand = cs.ncand; p->psi = cs.psi;
    p->cand_imax = cs.imax; p->inten = cs.inten; p->rbuf = cs.rbuf; p->queue = cs.queue; p->qcount = cs.qcount;
    int rs = 0;
    for (int k = 0; k < 3; k++) { int l = stored_lags(p->cfg[k]); if (l > rs) rs = l; }
    p->rstride = (rs + 7) & ~7;
}
static long long rbuf_need(const PitchPass& p, long long fub) {
    int rs = 0;
    for (int k = 0; k < 3; k++) { int l = stored_lags(p.cfg[k]); if (l > rs) rs = l; }
    return fub * (long long)((rs + 7) & ~7);
}

static void reg_debug(mshds_handle* h, const char* name, const void* base, const int* start, const int* count, int fixed,
                      int elem_size, int stride = 1) {
    DebugEntry e{base, start, count, fixed, elem_size, stride, {}};
    h->debug[name] = e;
}

// ------------------------------------------------------------------------------------------------ resample planning
static int ilog2_ceil(long long v) { int k = 0; while ((1LL << k) < v) k++; return k; }

// Sound_resample bookkeeping for one source sound: xmin = 0, xmax given
static void fill_resample_job(ResampleJob* J, long long clip_off, long long nclip, long long ix1, long long nx, double x1,
                              double xmax, double fs_new) {
    memset(J, 0, sizeof(*J));
    J->clip_off = clip_off; J->nclip = nclip; J->ix1 = ix1; J->nx = nx; J->x1 = x1;
    long long nout = (long long)floor((xmax - 0.0) * fs_new + 0.5);
    J->nout = nout;
    J->out_dx = 1.0 / fs_new;
    J->out_x1 = 0.5 * (0.0 + xmax - (double)(nout - 1) / fs_new);
}

struct ResamplePlan {
    std::vector<ResampleJob> jobs;
    std::vector<int> ids;                 // job indices grouped by FFT size
    std::vector<std::pair<int, int>> groups;   // (logn, count) in ids order
    std::vector<long long> out_prefix;
    std::vector<int> table_rep;
    std::vector<int> tile_prefix;         // polyphase FIR tiles per job (exclusive scan)
    int phases = 0, qstep = 0;
    long long ztotal = 0, ftotal = 0, ototal = 0;
};

static void finish_plan(ResamplePlan* P, bool share_tables_by_length, double fs, double fs_new) {
    const int nj = (int)P->jobs.size();
    std::map<int, std::vector<int>> bylog;
    std::map<std::tuple<long long, long long, double, double>, int> table_of_len;   // same geometry -> same coefficient rows
    P->out_prefix.assign(nj + 1, 0);
    P->tile_prefix.assign(nj + 1, 0);
    {   // polyphase period of the rate change (16 kHz -> 10 kHz: 5 phases, input advances 8 samples per period)
        long long a = (long long)llround(fs), b = (long long)llround(fs_new), gg = a, r = b;
        while (r) { long long t = gg % r; gg = r; r = t; }
        P->phases = (int)(b / gg); P->qstep = (int)(a / gg);
        if (P->phases > 640) { P->phases = 0; P->qstep = 0; }     // no useful period: every output evaluates its own window
    }
    for (int j = 0; j < nj; j++) {
        ResampleJob& J = P->jobs[j];
        int logn = ilog2_ceil(J.nx + 2000);
        bylog[logn].push_back(j);
        J.zoff = P->ztotal; P->ztotal += (logn > 12) ? (1LL << logn) : 0;
        J.filt_off = P->ftotal; P->ftotal += J.nx;
        J.out_off = P->ototal; P->ototal += J.nout;
        P->out_prefix[j + 1] = P->out_prefix[j] + J.nout;
        if (P->phases > 0 && P->phases <= 16 && P->qstep <= 64) {
            long long groups = (J.nout + P->phases - 1) / P->phases;
            P->tile_prefix[j + 1] = P->tile_prefix[j] + (int)((groups + 32 * SINC_FIR_R - 1) / (32 * SINC_FIR_R));
        }
        if (share_tables_by_length) {
            auto key = std::make_tuple(J.nx, J.nout, J.x1, J.out_x1);
            auto it = table_of_len.find(key);
            if (it == table_of_len.end()) { it = table_of_len.emplace(key, (int)P->table_rep.size()).first; P->table_rep.push_back(j); }
            J.table_id = it->second;
        } else { J.table_id = (int)P->table_rep.size(); P->table_rep.push_back(j); }
    }
    for (auto& kv : bylog) {
        P->groups.push_back(std::make_pair(kv.first, (int)kv.second.size()));
        for (int j : kv.second) P->ids.push_back(j);
    }
}

struct ResampleDev { ResampleJob* jobs; int* ids; long long* out_prefix; int* table_rep; int* tile_prefix; double2* zbuf; double* filt; double* out; double* table; };

static void run_resample(mshds_handle* h, const ResamplePlan& P, const ResampleDev& D, SPtr pcm, double fs, double fs_new,
                         int precision, cudaStream_t s) {
    const double dx = 1.0 / fs;
    const double upfactor = fs_new * dx;
    int pos = 0;
    PB(precision >= 500 ? "  resample500: fft low-pass" : "  resample50: fft low-pass");
    if (upfactor >= 1.0) {
        long long mx = 0;
        for (auto& J : P.jobs) mx = J.nx > mx ? J.nx : mx;
        launch_resample_copy(D.jobs, (int)P.jobs.size(), mx, pcm, D.filt, s, &h->launches);
    } else for (auto& g : P.groups) {
        launch_resample_fft_group(D.jobs, D.ids + pos, g.second, g.first, pcm, D.zbuf, D.filt, h->tw, upfactor, s, &h->launches);
        pos += g.second;
    }
    PE();
    PB(precision >= 500 ? "  resample500: sinc interpolation" : "  resample50: sinc interpolation");
    launch_sinc_resample(D.jobs, D.out_prefix, (int)P.jobs.size(), P.ototal, D.table_rep, (int)P.table_rep.size(), D.tile_prefix,
                         P.tile_prefix.back(), D.filt, D.table, D.out, P.phases, P.qstep, precision, dx, s, &h->launches);
    PE();
}

static int upload_plan(mshds_handle* h, const ResamplePlan& P, const ResampleDev& D, cudaStream_t s) {
    const int nj = (int)P.jobs.size();
    if (nj == 0) return MSHDS_OK;
    CK(cudaMemcpyAsync(D.jobs, P.jobs.data(), sizeof(ResampleJob) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(D.ids, P.ids.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(D.out_prefix, P.out_prefix.data(), sizeof(long long) * (nj + 1), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(D.table_rep, P.table_rep.data(), sizeof(int) * P.table_rep.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(D.tile_prefix, P.tile_prefix.data(), sizeof(int) * (nj + 1), cudaMemcpyHostToDevice, s));
    return MSHDS_OK;
}

static int make_formant_window(mshds_handle* h, int nsw, const double** win) {
    auto it = h->gauss_formant.find(nsw);
    if (it == h->gauss_formant.end()) {
        std::vector<double> w(nsw);
        for (int i = 1; i <= nsw; i++) {
            double imid = 0.5 * (nsw + 1), edge = exp(-12.0);
            w[i - 1] = (exp(-48.0 * (i - imid) * (i - imid) / (nsw + 1) / (nsw + 1)) - edge) / (1.0 - edge);
        }
        double* d;
        int rc = upload(h, w, &d); if (rc) return rc;
        it = h->gauss_formant.emplace(nsw, d).first;
    }
    *win = it->second;
    return MSHDS_OK;
}

// ------------------------------------------------------------------------------------------------ CPP stage
// The voiced-segment list only exists on the device; one small read-back per chunk sizes the per-segment FFTs.
static int run_cpp_stage(mshds_handle* h, const Clips& c, const std::vector<long long>& off_host, const std::vector<long long>& lens,
                         const std::vector<double>& x1_host, const CppSegs& sg, const std::vector<int>& scap, int* d_seg_prefix, double fs, cudaStream_t s,
                         cudaStream_t rb /* read-back stream: already ordered after the voiced-interval kernel */) {
    const int n = c.n;
    const double dx = 1.0 / fs, fs10 = 10000.0;
    // The host only waits for `rb`; whatever is still queued on `s` keeps the GPU busy while the segment plan is built.
    std::vector<int> cnt(n);
    CK(cudaMemcpyAsync(cnt.data(), sg.count, sizeof(int) * n, cudaMemcpyDeviceToHost, rb));
    CK(cudaStreamSynchronize(rb));
    std::vector<int> sprefix(n + 1, 0);
    for (int i = 0; i < n; i++) sprefix[i + 1] = sprefix[i] + cnt[i];
    const int nsegs = sprefix[n];
    CK(cudaMemcpyAsync(d_seg_prefix, sprefix.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, s));
    std::vector<double> tmin(scap[n]), tmax(scap[n]);
    std::vector<long long> ix1(scap[n]);
    std::vector<int> nseg(scap[n]);
    if (nsegs > 0) {
        CK(cudaMemcpyAsync(tmin.data(), sg.tmin, sizeof(double) * scap[n], cudaMemcpyDeviceToHost, rb));
        CK(cudaMemcpyAsync(tmax.data(), sg.tmax, sizeof(double) * scap[n], cudaMemcpyDeviceToHost, rb));
        CK(cudaMemcpyAsync(ix1.data(), sg.ix1, sizeof(long long) * scap[n], cudaMemcpyDeviceToHost, rb));
        CK(cudaMemcpyAsync(nseg.data(), sg.nseg, sizeof(int) * scap[n], cudaMemcpyDeviceToHost, rb));
        CK(cudaStreamSynchronize(rb));
    }
    ResamplePlan plan;
    std::vector<CepSeg> cseg(nsegs);
    std::vector<int> fprefix(nsegs + 1, 0);
    plan.jobs.resize(nsegs);
    const double dt = 0.002, pitchFloor = 60.0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < cnt[i]; k++) {
            const int src = scap[i] + k, j = sprefix[i] + k;
            // Sound_extractPart: xmin = 0, xmax = t2 - t1, x1 = my x1 + (ix1 - 1) dx - t1
            const double xmax = tmax[src] - tmin[src];
            const double x1 = (x1_host[i] + (double)(ix1[src] - 1) * dx) - tmin[src];
            fill_resample_job(&plan.jobs[j], off_host[i], lens[i], ix1[src], nseg[src], x1, xmax, fs10);
            // Sound_to_PowerCepstrogram set-up
            CepSeg& S = cseg[j];
            double windowDuration = 2.0 * (3.0 / pitchFloor);
            if (windowDuration > dx * (double)nseg[src]) windowDuration = dx * (double)nseg[src];
            int nFrames = 0;
            double t1 = 0.0;
            if (!short_term_analysis(nseg[src], dx, x1, windowDuration, dt, &nFrames, &t1)) nFrames = 0;
            S.t1 = t1; S.windowDuration = windowDuration;
            S.nwin = (int)floor(windowDuration * fs10 + 0.5);
            int nfft = 2;
            while (nfft < S.nwin) nfft *= 2;
            if (S.nwin < 1 || plan.jobs[j].nout < 1 || nfft > 1024) nFrames = 0;     // (nwin <= 1000 by construction)
            S.nfft = nfft;
            S.logM = 0;
            while ((1 << S.logM) < nfft / 2) S.logM++;
            const int nq = nfft / 2 + 1;
            const double qmax = 0.5 * nfft / fs10;
            S.dq = qmax / (nq - 1);
            S.pad = 0;
            fprefix[j + 1] = fprefix[j] + nFrames;
        }
    finish_plan(&plan, false, fs, fs10);
    const int totalFrames = fprefix[nsegs];
    const int nqmax = 513;
    // carve the CPP scratch
    size_t need = 0;
    auto sz = [&](size_t bytes) { size_t o = (need + 255) & ~(size_t)255; need = o + bytes; return o; };
    size_t o_jobs = sz(sizeof(ResampleJob) * (nsegs + 1)), o_ids = sz(sizeof(int) * (nsegs + 1));
    size_t o_opre = sz(sizeof(long long) * (nsegs + 2)), o_trep = sz(sizeof(int) * (nsegs + 1));
    size_t o_z = sz(sizeof(double2) * (size_t)(plan.ztotal + 1)), o_filt = sz(sizeof(double) * (size_t)(plan.ftotal + 1));
    size_t o_out = sz(sizeof(double) * (size_t)(plan.ototal + 1)), o_tab = sz(sizeof(double) * ((size_t)nsegs * (5 * 100 + 32) + 1));
    size_t o_tile = sz(sizeof(int) * (nsegs + 2));
    size_t o_cseg = sz(sizeof(CepSeg) * (nsegs + 1)), o_fpre = sz(sizeof(int) * (nsegs + 2));
    size_t o_cep = sz(sizeof(double) * ((size_t)totalFrames * nqmax + 1)), o_cppf = sz(sizeof(double) * ((size_t)totalFrames + 1));
    size_t o_turn = sz(sizeof(int) * 4);
    if (need > h->cpp_cap) {
        CK(cudaStreamSynchronize(s));
        if (h->cpp_buf) CK(cudaFree(h->cpp_buf));
        h->cpp_buf = nullptr; h->cpp_cap = 0;
        size_t cap = need + (need >> 2);
        CK(cudaMalloc((void**)&h->cpp_buf, cap));
        h->cpp_cap = cap;
    }
    char* B = h->cpp_buf;
    ResampleDev D;
    D.jobs = (ResampleJob*)(B + o_jobs); D.ids = (int*)(B + o_ids); D.out_prefix = (long long*)(B + o_opre);
    D.table_rep = (int*)(B + o_trep); D.zbuf = (double2*)(B + o_z); D.filt = (double*)(B + o_filt); D.out = (double*)(B + o_out);
    D.table = (double*)(B + o_tab); D.tile_prefix = (int*)(B + o_tile);
    CepSeg* d_cseg = (CepSeg*)(B + o_cseg);
    int* d_fprefix = (int*)(B + o_fpre);
    double* d_cep = (double*)(B + o_cep);
    double* d_cppf = (double*)(B + o_cppf);
    int* cep_turn = (int*)(B + o_turn);
    if (nsegs > 0) {
        int rc = upload_plan(h, plan, D, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(d_cseg, cseg.data(), sizeof(CepSeg) * nsegs, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemcpyAsync(d_fprefix, fprefix.data(), sizeof(int) * (nsegs + 1), cudaMemcpyHostToDevice, s));
    if (nsegs > 0 && totalFrames > 0) {
        PB("resample_segments_10k[fft+sinc50]"); run_resample(h, plan, D, c.pcm, fs, fs10, 50, s); PE();
        const double emphasis = exp(-2.0 * MSHDS_PI * 50.0 * (1.0 / fs10));
        const double* cep_win = nullptr;                 // same formula as the formant window (Sound_createGaussian)
        { int rcw = make_formant_window(h, 1000, &cep_win); if (rcw) return rcw; }
        PB("cepstrogram_frames"); launch_cepstrogram(d_cseg, d_fprefix, nsegs, D.jobs, D.out, h->tw, emphasis, dt, d_cep, nqmax, totalFrames, cep_win, 1000, s,
                                                      h->legacy_fft ? nullptr : h->twb512, cep_turn); h->launches += 2; PE();
        const int nTimeAvg = (int)floor(0.01 / dt);
        PB("cpps_frames"); launch_cpp_frames(d_cseg, d_fprefix, nsegs, d_cep, nqmax, nTimeAvg, 0.001, d_cppf, totalFrames, s); h->launches += 1; PE();
    }
    launch_cpp_reduce(c, sg, d_seg_prefix, d_fprefix, d_cppf, s); h->launches += 1;
    CK(cudaStreamSynchronize(s));       // host vectors above must outlive the async copies
    return MSHDS_OK;
}

// ------------------------------------------------------------------------------------------------ one chunk
static int process_chunk(mshds_handle* h, SPtr d_pcm, const std::vector<long long>& off_host, double fs,
                         const std::vector<double>& x1_host, const std::vector<double>& xmax_host,
                         double* d_feat, uint32_t* d_status, bool dry_run) {
    cudaStream_t s = h->stream;
    const int n = (int)off_host.size() - 1;
    const double dx = 1.0 / fs;
    std::vector<long long> lens(n);
    long long maxlen = 0;
    for (int i = 0; i < n; i++) { lens[i] = off_host[i + 1] - off_host[i]; if (lens[i] > maxlen) maxlen = lens[i]; }

    h->arena_off = 0;
    h->arena_overflow = false;
    h->debug.clear();

    Clips c;
    c.n = n; c.pcm = d_pcm; c.fs = fs; c.dx = dx;
    long long* d_off = take<long long>(h, n + 1);
    c.off = d_off;
    c.mean = take<double>(h, n); c.gpeak = take<double>(h, n); c.apeak = take<double>(h, n);
    double* d_x1 = take<double>(h, n); double* d_xmax = take<double>(h, n);
    c.x1 = d_x1; c.xmax = d_xmax;
    c.cls = take<int>(h, n);
    c.status = d_status; c.feat = d_feat;
    c.total_samples = off_host[n]; c.twb512 = h->twb512; c.twb1024 = h->twb1024; c.legacy_fft = h->legacy_fft; c.legacy_cc = h->legacy_cc;
    void* stat_scratch = arena_take(h, (size_t)n * 16);

    // ---- pitch pass configurations (parselmouth defaults unless mshds_extractor.py passes a value)
    PitchPass wide, mainp, hnr, srp, ltp, ccp, cpp_p;
    memset(&wide, 0, sizeof wide); memset(&mainp, 0, sizeof mainp); memset(&hnr, 0, sizeof hnr);
    memset(&srp, 0, sizeof srp); memset(&ltp, 0, sizeof ltp); memset(&ccp, 0, sizeof ccp); memset(&cpp_p, 0, sizeof cpp_p);
    int rc;
    for (int k = 0; k < 3; k++) {
        // :143 to_pitch_ac(time_step=0.005, pitch_floor=50, pitch_ceiling=600)
        if ((rc = make_pitch_cfg(h, fs, 0.005, 50.0, 3.0, 15, 0, 0.03, 0.45, 0.01, 0.35, 0.14, 600.0, &wide.cfg[k]))) return rc;
        // :178 / :355 to_pitch_ac(0.005, floor, ceiling)
        if ((rc = make_pitch_cfg(h, fs, 0.005, cls_floor(k), 3.0, 15, 0, 0.03, 0.45, 0.01, 0.35, 0.14, cls_ceiling(k), &mainp.cfg[k]))) return rc;
        // :221 to_harmonicity_cc(0.005, floor, 0.1, 4.5): FCC, 15 candidates, all path costs 0, ceiling = Nyquist
        if ((rc = make_pitch_cfg(h, fs, 0.005, cls_floor(k), 4.5, 15, 2, 0.1, 0.0, 0.0, 0.0, 0.0, 0.5 * fs, &hnr.cfg[k]))) return rc;
        // :104 to_pitch_ac(0.02, 30, 4, False, 0.03, 0.25, 0.01, 0.35, 0.25, 450)
        if ((rc = make_pitch_cfg(h, fs, 0.02, 30.0, 3.0, 4, 0, 0.03, 0.25, 0.01, 0.35, 0.25, 450.0, &srp.cfg[k]))) return rc;
        // :241 Sound_to_PointProcess_periodic_cc -> Sound_to_Pitch (0.0, floor, ceiling): dt = 0.75 / floor
        if ((rc = make_pitch_cfg(h, fs, 0.0, cls_floor(k), 3.0, 15, 0, 0.03, 0.45, 0.01, 0.35, 0.14, cls_ceiling(k), &ltp.cfg[k]))) return rc;
        // :320 to_pitch_cc(0.005, floor, ceiling): forward cross-correlation, 1 period per window
        if ((rc = make_pitch_cfg(h, fs, 0.005, cls_floor(k), 1.0, 15, 2, 0.03, 0.45, 0.01, 0.35, 0.14, cls_ceiling(k), &ccp.cfg[k]))) return rc;
        // :270 to_pitch_ac(0.005, floor, ceiling, voicing_threshold=0.3)
        if ((rc = make_pitch_cfg(h, fs, 0.005, cls_floor(k), 3.0, 15, 0, 0.03, 0.3, 0.01, 0.35, 0.14, cls_ceiling(k), &cpp_p.cfg[k]))) return rc;
    }
    hnr.hnr_mode = 1;

    const long long fub5 = frames_upper_bound(lens, dx, 0.005);
    CandScratchBuf cs;
    cs.cap = fub5;
    cs.f = take<double>(h, fub5 * MAXCAND); cs.s = take<double>(h, fub5 * MAXCAND);
    cs.score = take<double>(h, fub5 * MAXCAND); cs.lf = take<double>(h, fub5 * MAXCAND);
    cs.ncand = take<uint8_t>(h, fub5); cs.psi = take<uint8_t>(h, fub5 * 16);
    cs.imax = take<unsigned short>(h, fub5 * MAXCAND); cs.inten = take<double>(h, fub5);
    cs.queue = take<int>(h, fub5 * (MAXCAND - 1)); cs.qcount = take<int>(h, 1);
    const long long fub20 = frames_upper_bound(lens, dx, 0.02);
    const long long fub75 = frames_upper_bound(lens, dx, 0.75 / 100.0);
    {   // correlation rows kept between the frame kernel and the refinement kernel (largest pass decides)
        long long need = rbuf_need(wide, fub5);
        need = std::max(need, rbuf_need(mainp, fub5)); need = std::max(need, rbuf_need(hnr, fub5));
        need = std::max(need, rbuf_need(srp, fub20)); need = std::max(need, rbuf_need(ltp, fub75));
        need = std::max(need, rbuf_need(ccp, fub5)); need = std::max(need, rbuf_need(cpp_p, fub5));
        cs.rbuf = take<double>(h, need);
    }
    // the CPP pitch (:270) differs from the main one (:178) only in voicing_threshold: it shares the main pass's frames
    // and correlation rows, so its candidates live in a second scratch set
    CandScratchBuf cs2 = cs;
    cs2.f = take<double>(h, fub5 * MAXCAND); cs2.s = take<double>(h, fub5 * MAXCAND);
    cs2.score = take<double>(h, fub5 * MAXCAND); cs2.lf = take<double>(h, fub5 * MAXCAND);
    cs2.ncand = take<uint8_t>(h, fub5); cs2.psi = take<uint8_t>(h, fub5 * 16);
    cs2.imax = take<unsigned short>(h, fub5 * MAXCAND); cs2.inten = take<double>(h, fub5);
    cs2.queue = take<int>(h, fub5 * (MAXCAND - 1)); cs2.qcount = take<int>(h, 1);
    // A pass whose Viterbi runs on the side stream while the main stream already fills the next pass's candidates owns the
    // arrays the path finder reads (f, s, score, lf, ncand, psi); the frame -> refine -> score scratch stays shared.
    auto own_set = [&](long long fub) {
        CandScratchBuf o = cs;
        o.f = take<double>(h, fub * MAXCAND); o.s = take<double>(h, fub * MAXCAND);
        o.score = take<double>(h, fub * MAXCAND); o.lf = take<double>(h, fub * MAXCAND);
        o.ncand = take<uint8_t>(h, fub); o.psi = take<uint8_t>(h, fub * 16);
        return o;
    };
    const CandScratchBuf cs_wide = own_set(fub5), cs_sr = own_set(fub20), cs_lt = own_set(fub75), cs_cc = own_set(fub5);
    alloc_pitch_pass(h, &wide, n, fub5, cs_wide);
    alloc_pitch_pass(h, &mainp, n, fub5, cs);
    alloc_pitch_pass(h, &hnr, n, fub5, cs_wide);          // no path finder; the wide pass is long finished by then
    // harmonicity: worst case one maximum every other lag (maximumLag / 2 per frame)
    hnr.q64_cap = (unsigned long long)fub5 * 136ull;
    hnr.queue64 = take<unsigned long long>(h, hnr.q64_cap);
    hnr.qcount64 = take<unsigned long long>(h, 1);
    hnr.best_bits = take<unsigned long long>(h, fub5);
    alloc_pitch_pass(h, &srp, n, fub20, cs_sr);
    alloc_pitch_pass(h, &ltp, n, fub75, cs_lt);
    alloc_pitch_pass(h, &ccp, n, fub5, cs_cc);
    alloc_pitch_pass(h, &cpp_p, n, fub5, cs2);
    // cpp_p reuses the frame grid of mainp (identical dt, floor, window)
    cpp_p.nF = mainp.nF; cpp_p.t1 = mainp.t1; cpp_p.fstart = mainp.fstart;
    mainp.dual_vt = cpp_p.cfg[0].vt;
    mainp.dual_cand_f = cpp_p.cand_f; mainp.dual_cand_s = cpp_p.cand_s; mainp.dual_cand_imax = cpp_p.cand_imax;
    mainp.dual_ncand = cpp_p.ncand; mainp.dual_inten = cpp_p.inten; mainp.dual_queue = cpp_p.queue; mainp.dual_qcount = cpp_p.qcount;

    // ---- glottal pulses: raw / final capacity per clip (shared raw scratch, one final set per consumer)
    const double cprime = 500.0 / 0.8 * 1.15;
    std::vector<int> pcap(n + 1, 0);
    for (int i = 0; i < n; i++) {
        long long nfr = (long long)floor((double)lens[i] * dx / 0.005) + 2;
        long long cap = (long long)floor((double)lens[i] * dx * cprime) + 8 * (nfr / 2 + 2) + 16;
        pcap[i + 1] = pcap[i] + (int)cap;
    }
    const long long ptotal = pcap[n];
    PulseSet pl_lt;
    memset(&pl_lt, 0, sizeof pl_lt);
    pl_lt.cprime = cprime;
    int* d_pcap = take<int>(h, n + 1);
    // every pulse set owns its raw scratch: the three walks run on different side streams
    auto alloc_pulses = [&](PulseSet* ps) {
        ps->cprime = cprime; ps->cap_start = d_pcap;
        ps->t = take<double>(h, ptotal); ps->count = take<int>(h, n); ps->valid = take<int>(h, n);
        ps->st_count = take<int>(h, n); ps->st_start = take<int>(h, n + 1);
        ps->st_ileft = take<int>(h, fub5); ps->st_iright = take<int>(h, fub5);
        ps->raw_t = take<double>(h, 2 * ptotal); ps->raw_thr = take<double>(h, 2 * ptotal);      // left / right walk halves
        ps->raw_nleft = take<int>(h, fub5); ps->raw_nright = take<int>(h, fub5);
        ps->raw_added_right = take<double>(h, fub5); ps->raw_region = take<long long>(h, fub5);
    };
    alloc_pulses(&pl_lt);
    PulseSet pl_fm, pl_cp;
    memset(&pl_fm, 0, sizeof pl_fm); memset(&pl_cp, 0, sizeof pl_cp);
    alloc_pulses(&pl_fm);
    alloc_pulses(&pl_cp);

    // ---- formants (:319): whole-clip resample to 10 kHz (precision 500) + Burg frames
    const double fs10 = 10000.0;
    ResamplePlan fplan;
    fplan.jobs.resize(n);
    for (int i = 0; i < n; i++)
        fill_resample_job(&fplan.jobs[i], off_host[i], lens[i], 1, lens[i], x1_host[i], xmax_host[i], fs10);
    finish_plan(&fplan, true, fs, fs10);
    ResampleDev fdev;
    fdev.jobs = take<ResampleJob>(h, n); fdev.ids = take<int>(h, n); fdev.out_prefix = take<long long>(h, n + 1);
    fdev.table_rep = take<int>(h, fplan.table_rep.size()); fdev.tile_prefix = take<int>(h, n + 1);
    fdev.zbuf = take<double2>(h, fplan.ztotal); fdev.filt = take<double>(h, fplan.ftotal); fdev.out = take<double>(h, fplan.ototal);
    fdev.table = take<double>(h, fplan.table_rep.size() * (5 * 1000 + 32));
    FormantPass fm;
    memset(&fm, 0, sizeof fm);
    fm.jobs = fdev.jobs; fm.dt = 0.005; fm.dt_window = 2.0 * 0.025;
    {
        double dx10 = 1.0 / fs10;
        fm.nsamp_window = (int)floor(fm.dt_window / dx10);
        fm.emphasis = exp(-2.0 * MSHDS_PI * 50.0 * dx10);
        fm.nyquist = 0.5 / dx10;
        if ((rc = make_formant_window(h, fm.nsamp_window, &fm.window))) return rc;
    }
    fm.nF = take<int>(h, n); fm.t1 = take<double>(h, n); fm.fstart = take<int>(h, n + 1);
    fm.nform = take<int>(h, fub5); fm.freq = take<double>(h, fub5 * 5); fm.bw = take<double>(h, fub5 * 5);

    // ---- CPP (:253-301): voiced segment list (capacity: one segment per 20 ms at most)
    std::vector<int> scap(n + 1, 0);
    for (int i = 0; i < n; i++) scap[i + 1] = scap[i] + (int)floor((double)lens[i] * dx * 50.0) + 2;
    CppSegs sg;
    sg.cap_start = take<int>(h, n + 1); sg.count = take<int>(h, n); sg.fail = take<int>(h, n);
    sg.tmin = take<double>(h, scap[n]); sg.tmax = take<double>(h, scap[n]);
    sg.ix1 = take<long long>(h, scap[n]); sg.nseg = take<int>(h, scap[n]);
    int* seg_prefix = take<int>(h, n + 1);

    // ---- LTAS
    LtasPass lt;
    lt.part_count = take<int>(h, n); lt.part_start = take<int>(h, n + 1); lt.fail = take<int>(h, n);
    lt.partial = take<double>(h, (size_t)(ptotal / 64 + n + 1) * 100);
    double* ltas_bands = take<double>(h, (size_t)n * 50);

    // ---- speech rate: 16 ms intensity contour (:41) and its scratch
    IntensityPass isr;
    memset(&isr, 0, sizeof isr);
    isr.class_dep = 0; isr.dt = 0.016;
    for (int k = 0; k < 3; k++) {
        isr.min_pitch[k] = 50.0;
        if ((rc = make_kaiser(h, fs, 50.0, &isr.halfN[k], &isr.win[k]))) return rc;
    }
    const long long fub16 = frames_upper_bound(lens, dx, 0.016);
    isr.nF = take<int>(h, n); isr.t1 = take<double>(h, n); isr.fstart = take<int>(h, n + 1);
    isr.out = take<double>(h, fub16);
    double* isr_stats = take<double>(h, (size_t)n * 4);
    SpeechRateScratch srs;
    srs.ivl = take<Ivl>(h, fub16 + 2 * n + 2); srs.pk_t = take<double>(h, fub16 + 2 * n + 2);
    srs.pk_v = take<double>(h, fub16 + 2 * n + 2); srs.pk_i = take<int>(h, fub16 + 2 * n + 2);

    // ---- intensity (:198) and spectrogram (:356)
    IntensityPass imain;
    memset(&imain, 0, sizeof imain);
    imain.class_dep = 1; imain.dt = 0.005;
    for (int k = 0; k < 3; k++) {
        imain.min_pitch[k] = cls_floor(k);
        if ((rc = make_kaiser(h, fs, cls_floor(k), &imain.halfN[k], &imain.win[k]))) return rc;
    }
    imain.nF = take<int>(h, n); imain.t1 = take<double>(h, n); imain.fstart = take<int>(h, n + 1);
    imain.out = take<double>(h, fub5);
    double* imain_stats = take<double>(h, (size_t)n * 4);

    SpecPass spec;
    if ((rc = make_spec_cfg(h, fs, &spec))) return rc;
    spec.nF = take<int>(h, n); spec.t1 = take<double>(h, n); spec.fstart = take<int>(h, n + 1);
    spec.mom = take<double>(h, fub5 * 4);
    spec.turn_counter = take<int>(h, 1);

    if (dry_run) return MSHDS_OK;           // sizing pass only
    if (h->arena_overflow) { h->err = "internal: arena overflow after sizing"; return MSHDS_ERR_CUDA; }

    CK(cudaMemcpyAsync(d_off, off_host.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_x1, x1_host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_xmax, xmax_host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_pcap, pcap.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(sg.cap_start, scap.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, s));
    if ((rc = upload_plan(h, fplan, fdev, s))) return rc;
    const int fhint = (int)(fub5 > 0x3fffffff ? 0x3fffffff : fub5);

    // Two streams.  `s` carries everything that fills the GPU; `side` carries the kernels that are one warp (or thread) per
    // clip and therefore latency-bound: Viterbi, glottal-pulse walks, interval logic.  SIDE(ev) makes the side stream wait for
    // what `s` has issued so far and switches issue to it; MAIN() switches back; MARK(ev) records a point on the side stream
    // that `s` can later WAIT(ev) for.  With MSHDS_NO_OVERLAP=1 everything is issued on `s` in the same order.  Profiling spans
    // of side-stream work are named "~...": their durations include time spent sharing the SMs with the main stream.
    // Four side streams: the path finders / pulse walks of different analyses do not depend on each other, and with a single
    // recording in the batch (BASELINE.json configs[0]) their serial chains ARE the latency of the call.
    const bool ovl = h->overlap;
    cudaStream_t q = s;                          // stream the next launch goes to
    h->cur = s;
    int evn = 0;
    auto SIDE = [&](int k) {
        cudaStream_t side = ovl ? h->side[k] : s;
        if (side != s) { cudaEvent_t e = h->ev[evn++]; cudaEventRecord(e, s); cudaStreamWaitEvent(side, e, 0); }
        q = side; h->cur = side;
    };
    auto MAIN = [&]() { q = s; h->cur = s; };
    auto MARK = [&]() { cudaEvent_t e = h->ev[evn++]; if (q != s) cudaEventRecord(e, q); return e; };     // on the current side stream
    auto WAIT = [&](cudaEvent_t e) { if (ovl) cudaStreamWaitEvent(s, e, 0); };
    const int fhint20 = (int)(fub20 > 0x3fffffff ? 0x3fffffff : fub20), fhint75 = (int)(fub75 > 0x3fffffff ? 0x3fffffff : fub75);

    // ---- per-clip statistics (mean, global peaks); also clears status / feature rows
    PB("clip_stats"); launch_clip_stats(c, maxlen, stat_scratch, q); h->launches += 3; PE();

    // ---- _pitch_values (:127-162): wide AC pass -> speaker class
    launch_pitch_grid(c, wide, q); h->launches += 2;
    PB("pitch_ac_frames[wide 50-600Hz]"); launch_pitch_frames(c, wide, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, wide, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, wide, fhint, q); h->launches += 1; PE();
    SIDE(0);
    PB("~viterbi"); launch_pitch_viterbi(c, wide, q); h->launches += 1; PE();
    const cudaEvent_t ev_wide = MARK();
    MAIN();

    // ---- _speechrate (:11-125)
    PB("intensity_sr"); launch_intensity(c, isr, (int)(fub16 > 0x3fffffff ? 0x3fffffff : fub16), q); h->launches += 3;
    launch_contour_stats(c, isr, isr_stats, 1, q); h->launches += 1; PE();
    launch_pitch_grid(c, srp, q); h->launches += 2;
    PB("pitch_ac_frames[speechrate 30-450Hz]"); launch_pitch_frames(c, srp, h->tw, fhint20, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, srp, h->tw, fhint20, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, srp, fhint20, q); h->launches += 1; PE();
    SIDE(1);
    PB("~viterbi"); launch_pitch_viterbi(c, srp, q); h->launches += 1; PE();
    PB("~speechrate_logic"); launch_speechrate(c, isr, isr_stats, srp, srs, h->tw, q); h->launches += 1; PE();
    MAIN();

    WAIT(ev_wide);
    launch_pitch_class(c, wide, q); h->launches += 1;

    // ---- _extract_pitch (:164-183); the CPP pitch pass (:270) rides on the same frames and correlation rows
    launch_pitch_grid(c, mainp, q); h->launches += 2;
    PB("pitch_ac_frames[main + cpp vt=0.3]"); launch_pitch_frames(c, mainp, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, mainp, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, mainp, fhint, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, cpp_p, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, cpp_p, fhint, q); h->launches += 1; PE();
    SIDE(0);
    PB("~viterbi"); launch_pitch_viterbi(c, mainp, q); h->launches += 1; PE();
    launch_pitch_stats(c, mainp, q); h->launches += 1;
    const cudaEvent_t ev_mainpitch = MARK();
    MAIN();
    SIDE(2);
    PB("~viterbi"); launch_pitch_viterbi(c, cpp_p, q); h->launches += 1; PE();
    // ---- _extract_CPP (:253-301), first half: pulses of the vt=0.3 pitch and the voiced intervals
    PB("~pulses"); launch_pulses(c, cpp_p, pl_cp, q); h->launches += 5; PE();
    launch_vuv_segments(c, pl_cp, sg, q); h->launches += 1;
    const cudaEvent_t ev_vuv = MARK();
    MAIN();

    // ---- _extract_intensity (:185-205)
    PB("intensity_main"); launch_intensity(c, imain, fhint, q); h->launches += 3;
    launch_contour_stats(c, imain, imain_stats, 0, q); h->launches += 1; PE();
    launch_intensity_features(c, imain, imain_stats, q); h->launches += 1;

    // ---- _extract_harmonicity (:207-225)
    launch_pitch_grid(c, hnr, q); h->launches += 2;
    PB("pitch_cc_frames[hnr]"); launch_pitch_frames(c, hnr, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_hnr_refine"); launch_pitch_refine(c, hnr, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, hnr, fhint, q); h->launches += 1; PE();
    launch_hnr_mean(c, hnr, q); h->launches += 1;

    // ---- _extract_Slope_Tilt (:227-251)
    launch_pitch_grid(c, ltp, q); h->launches += 2;
    PB("pitch_ac_frames[ltas]"); launch_pitch_frames(c, ltp, h->tw, fhint75, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, ltp, h->tw, fhint75, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, ltp, fhint75, q); h->launches += 1; PE();
    SIDE(1);
    PB("~viterbi"); launch_pitch_viterbi(c, ltp, q); h->launches += 1; PE();
    PB("~pulses"); launch_pulses(c, ltp, pl_lt, q); h->launches += 5; PE();
    const cudaEvent_t ev_ltas = MARK();
    MAIN();

    // ---- _measureFormants (:303-338)
    PB("resample_clip_10k[fft+sinc500]"); run_resample(h, fplan, fdev, d_pcm, fs, fs10, 500, q); PE();
    PB("formant_burg_frames"); launch_formants(c, fm, n, fdev.out, fhint, q); h->launches += 3; PE();
    launch_pitch_grid(c, ccp, q); h->launches += 2;
    PB("pitch_cc_frames[formant]"); launch_pitch_frames(c, ccp, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_refine"); launch_pitch_refine(c, ccp, h->tw, fhint, q); h->launches += 1; PE();
    PB("k_pitch_score"); launch_pitch_score(c, ccp, fhint, q); h->launches += 1; PE();
    SIDE(3);
    PB("~viterbi"); launch_pitch_viterbi(c, ccp, q); h->launches += 1; PE();
    PB("~pulses"); launch_pulses(c, ccp, pl_fm, q); h->launches += 5; PE();
    const cudaEvent_t ev_fmt = MARK();
    MAIN();

    // ---- _extract_Spectral_Moments (:340-376); its pitch object is identical to _extract_pitch's
    WAIT(ev_mainpitch);
    PB("spectrogram_moments"); launch_moments(c, spec, mainp, h->tw, fhint, q); h->launches += 4; PE();

    WAIT(ev_ltas);
    PB("ltas"); launch_ltas(c, pl_lt, lt, ltas_bands, q); h->launches += 5; PE();

    // ---- _extract_CPP, second half (one small read-back, then resample / cepstrogram / CPPS of every voiced interval)
    WAIT(ev_vuv);
    cudaStream_t rb = s;
    if (ovl) { rb = h->copy; cudaStreamWaitEvent(rb, ev_vuv, 0); }
    if ((rc = run_cpp_stage(h, c, off_host, lens, x1_host, sg, scap, seg_prefix, fs, q, rb))) return rc;

    WAIT(ev_fmt);
    launch_formant_stats(c, fm, pl_fm, q); h->launches += 1;

    launch_finalize_status(c, s); h->launches += 1;
    CK(cudaGetLastError());

    reg_debug(h, "pitch_wide_f", wide.sel_f, wide.fstart, wide.nF, 0, 8);
    reg_debug(h, "pitch_main_f", mainp.sel_f, mainp.fstart, mainp.nF, 0, 8);
    reg_debug(h, "pitch_main_s", mainp.sel_s, mainp.fstart, mainp.nF, 0, 8);
    reg_debug(h, "hnr_r", hnr.sel_s, hnr.fstart, hnr.nF, 0, 8);
    reg_debug(h, "intensity_main", imain.out, imain.fstart, imain.nF, 0, 8);
    reg_debug(h, "moments", spec.mom, spec.fstart, spec.nF, 0, 8, 4);
    reg_debug(h, "class", c.cls, nullptr, nullptr, 1, 4);
    reg_debug(h, "pitch_sr_f", srp.sel_f, srp.fstart, srp.nF, 0, 8);
    reg_debug(h, "pitch_ltas_f", ltp.sel_f, ltp.fstart, ltp.nF, 0, 8);
    reg_debug(h, "intensity_sr", isr.out, isr.fstart, isr.nF, 0, 8);
    reg_debug(h, "pulses_ltas", pl_lt.t, pl_lt.cap_start, pl_lt.count, 0, 8);
    reg_debug(h, "ltas_bands", ltas_bands, nullptr, nullptr, 50, 8);
    reg_debug(h, "pitch_cc_f", ccp.sel_f, ccp.fstart, ccp.nF, 0, 8);
    reg_debug(h, "pitch_cpp_f", cpp_p.sel_f, cpp_p.fstart, cpp_p.nF, 0, 8);
    reg_debug(h, "pulses_fmt", pl_fm.t, pl_fm.cap_start, pl_fm.count, 0, 8);
    reg_debug(h, "pulses_cpp", pl_cp.t, pl_cp.cap_start, pl_cp.count, 0, 8);
    reg_debug(h, "formant_f", fm.freq, fm.fstart, fm.nF, 0, 8, 5);
    reg_debug(h, "formant_b", fm.bw, fm.fstart, fm.nF, 0, 8, 5);
    reg_debug(h, "formant_n", fm.nform, fm.fstart, fm.nF, 0, 4);
    reg_debug(h, "resampled10k", fdev.out, nullptr, nullptr, 0, 8);
    h->debug["resampled10k"].host_prefix = fplan.out_prefix;
    h->last_n = n;
    h->csrc[MSHDS_CONTOUR_F0] = {mainp.sel_f, mainp.sel_s, nullptr, mainp.fstart, mainp.nF, mainp.t1, mainp.cfg[0].dt};
    h->csrc[MSHDS_CONTOUR_INTENSITY] = {imain.out, nullptr, nullptr, imain.fstart, imain.nF, imain.t1, imain.dt};
    h->csrc[MSHDS_CONTOUR_HNR] = {hnr.sel_s, nullptr, nullptr, hnr.fstart, hnr.nF, hnr.t1, hnr.cfg[0].dt};
    h->csrc[MSHDS_CONTOUR_FORMANTS] = {fm.freq, fm.bw, fm.nform, fm.fstart, fm.nF, fm.t1, fm.dt};
    h->csrc[MSHDS_CONTOUR_MOMENTS] = {spec.mom, nullptr, nullptr, spec.fstart, spec.nF, spec.t1, spec.timeStep};
    return MSHDS_OK;
}

static const int kContourWidth[5] = {2, 1, 1, 4, 4};

// copy one contour of the chunk just processed (clips [c0, c0 + n) of the call) into the caller's buffers
static int collect_contours(mshds_handle* h, int c0, int n) {
    mshds_handle::ContourSink& K = *h->sink;
    const mshds_handle::ContourSrc& S = h->csrc[K.which];
    cudaStream_t s = h->stream;
    const int W = kContourWidth[K.which];
    std::vector<int> fstart(n + 1);
    std::vector<double> t1(n);
    CK(cudaMemcpyAsync(fstart.data(), S.fstart, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(t1.data(), S.t1, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int total = fstart[n];
    std::vector<double> a, b;
    std::vector<int> nform;
    const int src_w = K.which == MSHDS_CONTOUR_FORMANTS ? 5 : (K.which == MSHDS_CONTOUR_MOMENTS ? 4 : 1);
    if (total > 0) {
        a.resize((size_t)total * src_w);
        CK(cudaMemcpyAsync(a.data(), S.a, sizeof(double) * a.size(), cudaMemcpyDeviceToHost, s));
        if (S.b) { b.resize((size_t)total * src_w); CK(cudaMemcpyAsync(b.data(), S.b, sizeof(double) * b.size(), cudaMemcpyDeviceToHost, s)); }
        if (S.nform) { nform.resize(total); CK(cudaMemcpyAsync(nform.data(), S.nform, sizeof(int) * total, cudaMemcpyDeviceToHost, s)); }
        CK(cudaStreamSynchronize(s));
    }
    const double nan = std::nan("");
    for (int i = 0; i < n; i++) {
        if (K.t1) K.t1[c0 + i] = fstart[i + 1] > fstart[i] ? t1[i] : nan;
        if (K.frame_offsets) K.frame_offsets[c0 + i] = K.rows + fstart[i];
    }
    for (int f = 0; f < total; f++) {
        const long long r = K.rows + f;
        if ((size_t)r >= K.cap_rows) { K.overflow = true; break; }
        double* o = K.values + (size_t)r * W;
        switch (K.which) {
        case MSHDS_CONTOUR_F0: o[0] = a[f]; o[1] = b[f]; break;                              // Hz (0 = unvoiced), strength
        case MSHDS_CONTOUR_INTENSITY: o[0] = a[f]; break;                                    // dB
        case MSHDS_CONTOUR_HNR: o[0] = a[f] == a[f] ? 10.0 * log10(a[f] / (1.0 - a[f])) : -200.0; break;   // dB, -200 = voiceless
        case MSHDS_CONTOUR_FORMANTS:
            o[0] = nform[f] >= 1 ? a[(size_t)f * 5] : nan; o[1] = nform[f] >= 1 ? b[(size_t)f * 5] : nan;
            o[2] = nform[f] >= 2 ? a[(size_t)f * 5 + 1] : nan; o[3] = nform[f] >= 2 ? b[(size_t)f * 5 + 1] : nan;
            break;
        default: for (int k = 0; k < 4; k++) o[k] = a[(size_t)f * 4 + k];
        }
    }
    K.rows += total;
    if (K.frame_offsets) K.frame_offsets[c0 + n] = K.rows;
    return MSHDS_OK;
}

double run_dfma_peak(double* scratch, cudaStream_t s, int reps);      // k_microbench.cu

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int mshds_create(int device, mshds_handle** out) {
    if (!out) return MSHDS_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return MSHDS_ERR_CUDA;
    mshds_handle* h = new mshds_handle();
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return MSHDS_ERR_CUDA;
    }
    h->stream = h->own_stream;
    for (auto& st : h->side) if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MSHDS_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MSHDS_ERR_CUDA; }
    for (auto& e : h->ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete h; return MSHDS_ERR_CUDA; }
    { const char* e = getenv("MSHDS_NO_OVERLAP"); h->overlap = !(e && atoi(e)); }     // development switch
    // twiddle table exp(-2 pi i j / TW_N), j < TW_N/2
    const int TWN = 8192;
    std::vector<double> tw(TWN);
    for (int j = 0; j < TWN / 2; j++) {
        tw[2 * j] = cos(2.0 * MSHDS_PI * (double)j / (double)TWN);
        tw[2 * j + 1] = -sin(2.0 * MSHDS_PI * (double)j / (double)TWN);
    }
    if (cudaMalloc((void**)&h->tw, sizeof(double) * TWN) != cudaSuccess ||
        cudaMemcpy(h->tw, tw.data(), sizeof(double) * TWN, cudaMemcpyHostToDevice) != cudaSuccess) {
        delete h;
        return MSHDS_ERR_CUDA;
    }
    for (int M : {512, 1024}) {     // exp(-2 pi i j q / M), row q, L = M / 32 entries per row
        const int L = M / 32;
        std::vector<double> t((size_t)2 * M);
        for (int q = 0; q < 32; q++)
            for (int j = 0; j < L; j++) {
                const double ang = 2.0 * MSHDS_PI * (double)(j * q % M) / (double)M;
                t[2 * (size_t)(q * L + j)] = cos(ang);
                t[2 * (size_t)(q * L + j) + 1] = -sin(ang);
            }
        double2** dst = M == 512 ? &h->twb512 : &h->twb1024;
        if (cudaMalloc((void**)dst, sizeof(double) * t.size()) != cudaSuccess ||
            cudaMemcpy(*dst, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice) != cudaSuccess) { delete h; return MSHDS_ERR_CUDA; }
    }
    { const char* e = getenv("MSHDS_LEGACY_FFT"); h->legacy_fft = e && atoi(e); }
    { const char* e = getenv("MSHDS_LEGACY_CC"); h->legacy_cc = e ? atoi(e) : 0; }
    *out = h;
    return MSHDS_OK;
}

void mshds_destroy(mshds_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->ac_windows) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
    for (auto& kv : h->kaiser) cudaFree(kv.second);
    for (auto& kv : h->gauss_spec) cudaFree(kv.second);
    for (auto& kv : h->gauss_formant) cudaFree(kv.second);
    cudaFree(h->cpp_buf);
    cudaFree(h->front_buf);
    cudaFree(h->lld_buf);
    cudaFree(h->agg_buf);
    cudaFree(h->tw);
    cudaFree(h->twb512);
    cudaFree(h->twb1024);
    cudaFree(h->arena);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (auto& st : h->side) if (st) cudaStreamDestroy(st);
    if (h->copy) cudaStreamDestroy(h->copy);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    delete h;
}

int mshds_set_stream(mshds_handle* h, void* cuda_stream) {
    if (!h) return MSHDS_ERR_ARG;
    // a NULL cudaStream_t IS a stream: the legacy default stream (what torch.cuda.current_stream().cuda_stream is when no
    // stream context is active).  Work the caller queued there is ordered before the extraction, like on any other stream.
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamLegacy;
    return MSHDS_OK;
}

int mshds_reset_stream(mshds_handle* h) {
    if (!h) return MSHDS_ERR_ARG;
    h->stream = h->own_stream;
    return MSHDS_OK;
}

int mshds_set_option(mshds_handle* h, const char* name, long long value) {
    if (!h || !name) return MSHDS_ERR_ARG;
    const std::string n(name);
    if (n == "legacy_fft") h->legacy_fft = value != 0;
    else if (n == "legacy_cc") h->legacy_cc = (int)value;
    else if (n == "overlap") h->overlap = value != 0;
    else if (n == "nvtx") h->nvtx = value != 0;
    else { h->err = "unknown option: " + n; return MSHDS_ERR_ARG; }
    return MSHDS_OK;
}

int mshds_set_chunk_samples(mshds_handle* h, long long max_samples) {
    if (!h || max_samples < 1) return MSHDS_ERR_ARG;
    h->chunk_samples = max_samples;
    h->chunk_auto = false;
    return MSHDS_OK;
}

const char* mshds_last_error(const mshds_handle* h) { return h ? h->err.c_str() : "null handle"; }
long long mshds_launch_count(const mshds_handle* h) { return h ? h->launches : 0; }

int mshds_extract(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate,
                  double* features, uint32_t* status, unsigned flags) {
    if (!h) return MSHDS_ERR_ARG;
    h->err.clear();
    if (n_clips < 0 || (n_clips > 0 && (!pcm || !offsets || !features))) { h->err = "null pointer argument"; return MSHDS_ERR_ARG; }
    if (sample_rate < 4000 || sample_rate > 384000) { h->err = "sample_rate out of range"; return MSHDS_ERR_ARG; }
    if (n_clips == 0) return MSHDS_OK;
    for (int i = 0; i < n_clips; i++)
        if (offsets[i + 1] < offsets[i]) { h->err = "offsets must be non-decreasing"; return MSHDS_ERR_ARG; }
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    h->cur = s;
    const bool pcm_dev = flags & MSHDS_PCM_ON_DEVICE, out_dev = flags & MSHDS_OUT_ON_DEVICE;
    // MSHDS_PCM_FLOAT64: `pcm` really points to float64 samples in [-1, 1) (what parselmouth.Sound holds for a 24/32-bit or
    // float file); every kernel reads them through the same SPtr accessor the resampling front-end uses
    const bool f64_in = (flags & MSHDS_PCM_FLOAT64) != 0;
    const size_t ssz = f64_in ? 8 : 2;
    const double fs_in = (double)sample_rate;
    const bool front = sample_rate != 16000;          // mshds_extractor.py:418-419  snd.resample(16000, 50)
    const bool doubling = sample_rate == 8000;        // Sound_resample hands an exact doubling to Sound_upsample
    const double fs = 16000.0;

    // Chunk size.  The path finders and pulse walks are sequential per recording (one CTA / a few warps each) and only fill the
    // GPU through the number of recordings in flight: 2^27 samples are 279 clips of 30 s but only ~25 recordings of 5 minutes
    // (BASELINE.json configs[2]: 13.1 k audio-s/s at 2^27, 16.2 k at 2^29).  Unless the caller fixed the size, long
    // recordings therefore get chunks of up to 128 of them, bounded by 2^29 samples and by half of the free device memory
    // (the scratch arena takes ~170 B per sample).
    long long chunk_samples = h->chunk_samples;
    if (h->chunk_auto && !front && n_clips > 0) {
        const long long total = offsets[n_clips] - offsets[0];
        const long long want = (total / n_clips) * 128;
        if (want > chunk_samples) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const long long fit = (long long)((free_b + h->arena_cap) / 2 / 200);
                long long grown = want < (1LL << 29) ? want : (1LL << 29);
                if (grown > fit) grown = fit;
                if (grown > total) grown = total;
                if (grown > chunk_samples) chunk_samples = grown;
            }
        }
    }
    int c0 = 0;
    while (c0 < n_clips) {
        // clips [c0, c1) form one chunk
        int c1 = c0;
        long long tot = 0;
        while (c1 < n_clips && (c1 == c0 || tot + (offsets[c1 + 1] - offsets[c1]) <= chunk_samples)) {
            tot += offsets[c1 + 1] - offsets[c1];
            c1++;
        }
        const int n = c1 - c0;
        std::vector<long long> off_in(n + 1);
        for (int i = 0; i <= n; i++) off_in[i] = offsets[c0 + i] - offsets[c0];

        // time domain of every clip as the analyses see it, and (front-end) the resample plan to 16 kHz
        std::vector<long long> off(n + 1, 0);
        std::vector<double> x1v(n), xmaxv(n);
        ResamplePlan fe;
        if (!front) {
            off = off_in;
            for (int i = 0; i < n; i++) { x1v[i] = 0.5 * (1.0 / fs); xmaxv[i] = (double)(off[i + 1] - off[i]) * (1.0 / fs); }
        } else {
            const double dx_in = 1.0 / fs_in;
            fe.jobs.resize(n);
            for (int i = 0; i < n; i++) {
                const long long len = off_in[i + 1] - off_in[i];
                fill_resample_job(&fe.jobs[i], off_in[i], len, 1, len, 0.5 * dx_in, (double)len * dx_in, fs);
                if (doubling) {     // Sound_upsample: nx' = 2 nx, dx' = dx / 2, x1' = x1 - dx / 4
                    fe.jobs[i].nout = 2 * len;
                    fe.jobs[i].out_x1 = 0.5 * dx_in - dx_in / 4.0;
                }
                if (len <= 0 || fe.jobs[i].nout < 1) fe.jobs[i].nout = 0;    // empty clip (or too short to give one sample): whole row NaN
                x1v[i] = fe.jobs[i].out_x1; xmaxv[i] = (double)len * dx_in;
            }
            finish_plan(&fe, true, fs_in, fs);
            for (int i = 0; i < n; i++) off[i + 1] = off[i] + fe.jobs[i].nout;
        }

        // sizing pass, then (re)allocate the arena
        char* saved = h->arena;
        size_t saved_cap = h->arena_cap;
        h->arena = nullptr; h->arena_cap = 0;
        int rc = process_chunk(h, SPtr{nullptr, nullptr}, off, fs, x1v, xmaxv, nullptr, nullptr, true);
        h->arena = saved; h->arena_cap = saved_cap;
        if (rc) return rc;
        size_t need = h->arena_off + (pcm_dev ? 0 : (size_t)tot * ssz + 512) + (out_dev ? 0 : (size_t)n * (25 * 8 + 4) + 512) + 4096;
        if (need > h->arena_cap) {
            CK(cudaStreamSynchronize(s));
            if (h->arena) CK(cudaFree(h->arena));
            h->arena = nullptr; h->arena_cap = 0;
            CK(cudaMalloc((void**)&h->arena, need + (need >> 3)));
            h->arena_cap = need + (need >> 3);
            if (getenv("MSHDS_DEBUG_MEM")) fprintf(stderr, "mshds: scratch arena %.2f GB for a chunk of %d clips / %lld samples\n", (double)h->arena_cap / 1e9, n, tot);
        }
        // tail of the arena: staging for pcm / outputs when the caller's buffers live on the host
        size_t tail = h->arena_cap;
        const char* d_pcm;
        const char* pcm_bytes = (const char*)pcm + (size_t)offsets[c0] * ssz;
        if (pcm_dev) d_pcm = pcm_bytes;
        else {
            tail = (tail - (size_t)tot * ssz - 256) & ~(size_t)255;
            CK(cudaMemcpyAsync(h->arena + tail, pcm_bytes, (size_t)tot * ssz, cudaMemcpyHostToDevice, s));
            d_pcm = h->arena + tail;
        }
        double* d_feat;
        uint32_t* d_status;
        if (out_dev) {
            d_feat = features + (size_t)c0 * 25;
            d_status = status ? status + c0 : nullptr;
        } else {
            tail = (tail - (size_t)n * 25 * 8 - 256) & ~(size_t)255;
            d_feat = (double*)(h->arena + tail);
            d_status = nullptr;
        }
        if (!d_status) {
            tail = (tail - (size_t)n * 4 - 256) & ~(size_t)255;
            d_status = (uint32_t*)(h->arena + tail);
        }

        SPtr src{f64_in ? nullptr : (const int16_t*)d_pcm, f64_in ? (const double*)d_pcm : nullptr};
        const double* d_front_out = nullptr;
        if (front) {
            // ---- front-end: every clip -> 16 kHz float64 (FFT low-pass when down-sampling, sinc depth 50), own scratch
            size_t fneed = 0;
            auto sz = [&](size_t bytes) { size_t o = (fneed + 255) & ~(size_t)255; fneed = o + bytes; return o; };
            size_t o_jobs = sz(sizeof(ResampleJob) * (n + 1)), o_ids = sz(sizeof(int) * (n + 1));
            size_t o_opre = sz(sizeof(long long) * (n + 2)), o_trep = sz(sizeof(int) * (n + 1)), o_tile = sz(sizeof(int) * (n + 2));
            size_t o_z = sz(sizeof(double2) * (size_t)(fe.ztotal + 1)), o_filt = sz(sizeof(double) * (size_t)(fe.ftotal + 1));
            size_t o_out = sz(sizeof(double) * (size_t)(fe.ototal + 1));
            size_t o_tab = sz(sizeof(double) * (fe.table_rep.size() * (size_t)((fe.phases > 0 ? fe.phases : 1) * 102 + 8) + 8));
            if (fneed > h->front_cap) {
                CK(cudaStreamSynchronize(s));
                if (h->front_buf) CK(cudaFree(h->front_buf));
                h->front_buf = nullptr; h->front_cap = 0;
                CK(cudaMalloc((void**)&h->front_buf, fneed + (fneed >> 3)));
                h->front_cap = fneed + (fneed >> 3);
            }
            char* B = h->front_buf;
            ResampleDev D;
            D.jobs = (ResampleJob*)(B + o_jobs); D.ids = (int*)(B + o_ids); D.out_prefix = (long long*)(B + o_opre);
            D.table_rep = (int*)(B + o_trep); D.tile_prefix = (int*)(B + o_tile); D.zbuf = (double2*)(B + o_z);
            D.filt = (double*)(B + o_filt); D.out = (double*)(B + o_out); D.table = (double*)(B + o_tab);
            if ((rc = upload_plan(h, fe, D, s))) return rc;
            PB("frontend_resample_to_16k[fft+sinc50]");
            if (doubling) {
                // even output samples = inverse of the tapered spectrum, odd ones = inverse of the half-sample-advanced spectrum
                int pos = 0;
                for (auto& g : fe.groups) {
                    launch_resample_fft_group(D.jobs, D.ids + pos, g.second, g.first, src, D.zbuf, D.out, h->tw, 2.0, s, &h->launches, 1);
                    launch_resample_fft_group(D.jobs, D.ids + pos, g.second, g.first, src, D.zbuf, D.out, h->tw, 2.0, s, &h->launches, 2);
                    pos += g.second;
                }
            } else run_resample(h, fe, D, src, fs_in, fs, 50, s);
            PE();
            src = SPtr{nullptr, D.out};
            d_front_out = D.out;
        }

        size_t cap_saved = h->arena_cap;
        h->arena_cap = tail;                       // scratch may not run into the staging area
        rc = process_chunk(h, src, off, fs, x1v, xmaxv, d_feat, d_status, false);
        h->arena_cap = cap_saved;
        if (rc) return rc;
        if (front) {
            reg_debug(h, "resampled16k", d_front_out, nullptr, nullptr, 0, 8);
            h->debug["resampled16k"].host_prefix = fe.out_prefix;
        }
        if (!out_dev) {
            CK(cudaMemcpyAsync(features + (size_t)c0 * 25, d_feat, (size_t)n * 25 * 8, cudaMemcpyDeviceToHost, s));
            if (status) CK(cudaMemcpyAsync(status + c0, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
        prof_collect(h);
        if (h->sink && (rc = collect_contours(h, c0, n))) return rc;
        c0 = c1;
    }
    return MSHDS_OK;
}

int mshds_extract_contours(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate, int contour,
                           double* values, size_t capacity_rows, int64_t* frame_offsets, double* t1, double* dt, int* width,
                           double* features, unsigned flags) {
    if (!h) return MSHDS_ERR_ARG;
    h->err.clear();
    if (contour < 0 || contour > 4 || (capacity_rows > 0 && !values) || (flags & MSHDS_OUT_ON_DEVICE)) { h->err = "bad contour argument"; return MSHDS_ERR_ARG; }
    if (width) *width = kContourWidth[contour];
    std::vector<double> feat_tmp;
    if (!features) { feat_tmp.resize((size_t)(n_clips > 0 ? n_clips : 1) * 25); features = feat_tmp.data(); }
    mshds_handle::ContourSink K{contour, values, capacity_rows, frame_offsets, t1, 0, false};
    if (frame_offsets && n_clips >= 0) frame_offsets[0] = 0;
    h->sink = &K;
    int rc = mshds_extract(h, pcm, offsets, n_clips, sample_rate, features, nullptr, flags);
    h->sink = nullptr;
    if (rc) return rc;
    if (dt) *dt = n_clips > 0 ? h->csrc[contour].dt : 0.0;
    if (K.overflow) { h->err = "values buffer too small (capacity_rows)"; return MSHDS_ERR_ARG; }
    return MSHDS_OK;
}

void mshds_lld_default_params(mshds_lld_params* p) {
    if (!p) return;
    p->frame_size = 0.025; p->frame_step = 0.010; p->preemph = 0.97; p->n_fft = 0; p->n_mel = 26;
    p->mel_lo = 20.0; p->mel_hi = 8000.0; p->n_mfcc = 12; p->cep_lifter = 22.0; p->smooth_win = 3; p->delta_win = 2;
    p->descriptor_set = 0; p->functional_set = 0;
}

int mshds_lld_extract(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate,
                      const mshds_lld_params* prm, double* functionals, double* frames_out, int64_t* frame_offsets,
                      unsigned flags) {
    if (!h) return MSHDS_ERR_ARG;
    h->err.clear();
    mshds_lld_params P;
    if (prm) P = *prm; else mshds_lld_default_params(&P);
    if (n_clips < 0 || !offsets || (n_clips > 0 && !functionals) || sample_rate < 1000 || sample_rate > 384000) { h->err = "bad argument"; return MSHDS_ERR_ARG; }
    for (int i = 0; i < n_clips; i++) if (offsets[i + 1] < offsets[i]) { h->err = "offsets must be non-decreasing"; return MSHDS_ERR_ARG; }
    if (n_clips > 0 && offsets[n_clips] > offsets[0] && !pcm) { h->err = "pcm is NULL"; return MSHDS_ERR_ARG; }
    const double fs = (double)sample_rate;
    const int nf = (int)floor(P.frame_size * fs + 0.5), ns = (int)floor(P.frame_step * fs + 0.5);
    if (nf < 2 || ns < 1 || nf > 8192) { h->err = "frame_size / frame_step out of range"; return MSHDS_ERR_ARG; }
    int n_fft = P.n_fft;
    if (n_fft == 0) { n_fft = 64; while (n_fft < nf) n_fft <<= 1; }
    if (n_fft < nf || n_fft < 64 || n_fft > 8192 || (n_fft & (n_fft - 1))) { h->err = "n_fft must be a power of two in [max(64, frame length), 8192]"; return MSHDS_ERR_ARG; }
    if (P.n_mel < 2 || P.n_mel > 256 || P.n_mfcc < 1 || P.n_mfcc >= P.n_mel || !(P.preemph >= 0.0 && P.preemph < 1.0)) { h->err = "bad n_mel / n_mfcc / preemph"; return MSHDS_ERR_ARG; }
    if (P.smooth_win < 0 || P.smooth_win > 101 || (P.smooth_win > 1 && P.smooth_win % 2 == 0) || P.delta_win < 0 || P.delta_win > 50) { h->err = "smooth_win must be odd (or <= 1), delta_win in [0, 50]"; return MSHDS_ERR_ARG; }
    const double hi = P.mel_hi < 0.5 * fs ? P.mel_hi : 0.5 * fs;
    if (!(P.mel_lo >= 0.0 && P.mel_lo < hi)) { h->err = "bad mel range"; return MSHDS_ERR_ARG; }
    if (P.descriptor_set < 0 || P.descriptor_set > 1 || P.functional_set < 0 || P.functional_set > 1) { h->err = "descriptor_set / functional_set must be 0 or 1"; return MSHDS_ERR_ARG; }
    if (n_clips == 0) return MSHDS_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    h->cur = s;
    const bool pcm_dev = (flags & MSHDS_PCM_ON_DEVICE) != 0, out_dev = (flags & MSHDS_OUT_ON_DEVICE) != 0;
    const int M = n_fft / 2, D0 = P.n_mfcc + 2 + (P.descriptor_set ? 16 : 0);
    const int NF = P.functional_set ? 12 : 2;          // functionals per contour
    const bool post = P.smooth_win > 1 || P.delta_win > 0;
    const int D = D0 * (P.delta_win > 0 ? 2 : 1);          // width of a final row
    int logM = 0; while ((1 << logM) < M) logM++;
    // ---- host tables
    std::vector<double> window(nf), melbin(M + 1), centres(P.n_mel + 2), dct((size_t)P.n_mfcc * P.n_mel);
    std::vector<int> klo(P.n_mel), khi(P.n_mel);
    for (int j = 0; j < nf; j++) window[j] = 0.54 - 0.46 * cos(2.0 * MSHDS_PI * (double)j / (double)(nf - 1));
    auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
    for (int k = 0; k <= M; k++) melbin[k] = mel((double)k * fs / (double)n_fft);
    const double m_lo = mel(P.mel_lo), m_hi = mel(hi);
    for (int m = 0; m < P.n_mel + 2; m++) centres[m] = m_lo + (m_hi - m_lo) * (double)m / (double)(P.n_mel + 1);
    for (int m = 0; m < P.n_mel; m++) {
        int a = M + 1, b = -1;
        for (int k = 0; k <= M; k++) if (melbin[k] > centres[m] && melbin[k] < centres[m + 2]) { if (k < a) a = k; b = k; }
        klo[m] = a; khi[m] = b;                      // empty band: klo > khi
    }
    for (int i = 0; i < P.n_mfcc; i++)
        for (int m = 0; m < P.n_mel; m++) dct[(size_t)i * P.n_mel + m] = cos(MSHDS_PI * (double)(i + 1) * ((double)m + 0.5) / (double)P.n_mel);
    // ---- frame counts
    std::vector<long long> off(n_clips + 1);
    std::vector<long long> fo(n_clips + 1, 0);
    for (int i = 0; i <= n_clips; i++) off[i] = offsets[i] - offsets[0];
    for (int i = 0; i < n_clips; i++) {
        const long long nx = off[i + 1] - off[i];
        fo[i + 1] = fo[i] + (nx >= nf ? (nx - nf) / ns + 1 : 0);
    }
    const long long total_frames = fo[n_clips], total = off[n_clips];
    if (total_frames > 0x7fffff00LL) { h->err = "too many frames for one call"; return MSHDS_ERR_ARG; }
    if (frame_offsets) for (int i = 0; i <= n_clips; i++) frame_offsets[i] = fo[i];
    // ---- device buffer (grow-only)
    size_t need = 0;
    auto sz = [&](size_t bytes) { size_t o = (need + 255) & ~(size_t)255; need = o + bytes; return o; };
    const size_t o_off = sz(sizeof(long long) * (n_clips + 1)), o_nF = sz(sizeof(int) * n_clips), o_fs = sz(sizeof(int) * (n_clips + 1));
    const size_t o_win = sz(sizeof(double) * nf), o_mb = sz(sizeof(double) * (M + 1)), o_c = sz(sizeof(double) * (P.n_mel + 2));
    const size_t o_klo = sz(sizeof(int) * P.n_mel), o_khi = sz(sizeof(int) * P.n_mel), o_dct = sz(sizeof(double) * dct.size());
    const size_t o_raw = post ? sz(sizeof(double) * (size_t)(total_frames + 1) * D0) : 0;
    const size_t o_fr = (out_dev && frames_out) ? 0 : sz(sizeof(double) * (size_t)(total_frames + 1) * D);
    const size_t o_fun = out_dev ? 0 : sz(sizeof(double) * (size_t)n_clips * NF * D);
    const size_t o_pcm = pcm_dev ? 0 : sz((size_t)total * 2 + 16);
    if (need > h->lld_cap) {
        CK(cudaStreamSynchronize(s));
        if (h->lld_buf) CK(cudaFree(h->lld_buf));
        h->lld_buf = nullptr; h->lld_cap = 0;
        CK(cudaMalloc((void**)&h->lld_buf, need + (need >> 3)));
        h->lld_cap = need + (need >> 3);
    }
    char* Bf = h->lld_buf;
    CK(cudaMemcpyAsync(Bf + o_off, off.data(), sizeof(long long) * (n_clips + 1), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_win, window.data(), sizeof(double) * nf, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_mb, melbin.data(), sizeof(double) * (M + 1), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_c, centres.data(), sizeof(double) * (P.n_mel + 2), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_klo, klo.data(), sizeof(int) * P.n_mel, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_khi, khi.data(), sizeof(int) * P.n_mel, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(Bf + o_dct, dct.data(), sizeof(double) * dct.size(), cudaMemcpyHostToDevice, s));
    const int16_t* d_pcm = pcm_dev ? pcm + offsets[0] : (const int16_t*)(Bf + o_pcm);
    if (!pcm_dev && total > 0) CK(cudaMemcpyAsync(Bf + o_pcm, pcm + offsets[0], (size_t)total * 2, cudaMemcpyHostToDevice, s));
    LldPass L;
    L.nf = nf; L.ns = ns; L.n_fft = n_fft; L.M = M; L.logM = logM; L.n_mel = P.n_mel; L.n_mfcc = P.n_mfcc;
    L.preemph = P.preemph; L.lifter = P.cep_lifter; L.dct_scale = sqrt(2.0 / (double)P.n_mel); L.log_floor = 1e-10;
    L.window = (const double*)(Bf + o_win); L.melbin = (const double*)(Bf + o_mb); L.centres = (const double*)(Bf + o_c);
    L.klo = (const int*)(Bf + o_klo); L.khi = (const int*)(Bf + o_khi); L.dct = (const double*)(Bf + o_dct);
    L.nF = (int*)(Bf + o_nF); L.fstart = (int*)(Bf + o_fs);
    L.final = (out_dev && frames_out) ? frames_out : (double*)(Bf + o_fr);
    L.frames = post ? (double*)(Bf + o_raw) : L.final;
    L.smooth_win = P.smooth_win; L.delta_win = P.delta_win; L.W = D;
    L.desc = P.descriptor_set; L.fset = P.functional_set; L.D0 = D0; L.fs = fs;
    L.win_sum = 0.0;
    for (int j = 0; j < nf; j++) L.win_sum += window[j];
    L.band_lo[0] = 250.0; L.band_hi[0] = 650.0; L.band_lo[1] = 1000.0; L.band_hi[1] = 4000.0;       // Androids.conf:261-262
    L.rolloff[0] = 0.25; L.rolloff[1] = 0.50; L.rolloff[2] = 0.75; L.rolloff[3] = 0.90;                // Androids.conf:263-266
    double* d_fun = out_dev ? functionals : (double*)(Bf + o_fun);
    PB(P.descriptor_set ? "lld_frames[mfcc+energy+zcr+intensity+spectral]" : "lld_frames[mfcc+energy+zcr]");
    launch_lld_grid(n_clips, (const long long*)(Bf + o_off), nf, ns, L.nF, L.fstart, s);
    launch_lld_frames(L, d_pcm, (const long long*)(Bf + o_off), n_clips, h->tw, total_frames, s);
    PE();
    if (post) { PB("lld_smooth_delta"); launch_lld_post(L, n_clips, total_frames, s); h->launches += 1; PE(); }
    PB("lld_functionals");
    launch_lld_functionals(L, n_clips, d_fun, s);
    PE();
    h->launches += 4;
    CK(cudaGetLastError());
    if (!out_dev) {
        CK(cudaMemcpyAsync(functionals, d_fun, sizeof(double) * (size_t)n_clips * NF * D, cudaMemcpyDeviceToHost, s));
        if (frames_out && total_frames > 0)
            CK(cudaMemcpyAsync(frames_out, L.final, sizeof(double) * (size_t)total_frames * D, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    prof_collect(h);
    return MSHDS_OK;
}

int mshds_aggregate_sessions(mshds_handle* h, const double* features, int n_rows, int n_cols, const int32_t* row_group,
                             int n_groups, double* mean_out, double* std_out, unsigned flags) {
    if (!h) return MSHDS_ERR_ARG;
    h->err.clear();
    if (n_rows < 0 || n_cols < 1 || n_groups < 0 || (n_rows > 0 && (!features || !row_group)) ||
        (n_groups > 0 && (!mean_out || !std_out))) { h->err = "bad argument"; return MSHDS_ERR_ARG; }
    if ((long long)n_groups * n_cols > 0x7fffffffLL) { h->err = "too many sessions x columns"; return MSHDS_ERR_ARG; }
    if (n_groups == 0) return MSHDS_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return MSHDS_ERR_CUDA; }
    cudaStream_t s = h->stream;
    const bool on_dev = (flags & MSHDS_AGG_ON_DEVICE) != 0;
    // rows of every session in index order (counting sort on the host; the list is a few hundred entries)
    std::vector<int> start(n_groups + 1, 0), rows;
    for (int i = 0; i < n_rows; i++) {
        if (row_group[i] >= n_groups) { h->err = "row_group out of range"; return MSHDS_ERR_ARG; }
        if (row_group[i] >= 0) start[row_group[i] + 1]++;
    }
    for (int g = 0; g < n_groups; g++) start[g + 1] += start[g];
    rows.resize(start[n_groups] > 0 ? start[n_groups] : 1);
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int i = 0; i < n_rows; i++) if (row_group[i] >= 0) rows[fill[row_group[i]]++] = i;
    }
    const size_t fbytes = sizeof(double) * (size_t)(n_rows > 0 ? n_rows : 1) * n_cols, obytes = sizeof(double) * (size_t)n_groups * n_cols;
    char* buf = nullptr;
    const size_t o_start = 0, o_rows = (sizeof(int) * (n_groups + 1) + 255) & ~(size_t)255;
    const size_t o_feat = o_rows + ((sizeof(int) * rows.size() + 255) & ~(size_t)255);
    const size_t o_mean = o_feat + (on_dev ? 0 : ((fbytes + 255) & ~(size_t)255));
    const size_t o_std = o_mean + (on_dev ? 0 : ((obytes + 255) & ~(size_t)255));
    const size_t total = o_std + (on_dev ? 0 : obytes) + 256;
    if (total > h->agg_cap) {
        CK(cudaStreamSynchronize(s));
        if (h->agg_buf) CK(cudaFree(h->agg_buf));
        h->agg_buf = nullptr; h->agg_cap = 0;
        CK(cudaMalloc((void**)&h->agg_buf, total + (total >> 2)));
        h->agg_cap = total + (total >> 2);
    }
    buf = h->agg_buf;
    int rc = MSHDS_OK;
    do {
        if (cudaMemcpyAsync(buf + o_start, start.data(), sizeof(int) * (n_groups + 1), cudaMemcpyHostToDevice, s) != cudaSuccess ||
            cudaMemcpyAsync(buf + o_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = MSHDS_ERR_CUDA; break; }
        const double* d_feat = features;
        double* d_mean = mean_out; double* d_std = std_out;
        if (!on_dev) {
            if (n_rows > 0 && cudaMemcpyAsync(buf + o_feat, features, fbytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = MSHDS_ERR_CUDA; break; }
            d_feat = (const double*)(buf + o_feat); d_mean = (double*)(buf + o_mean); d_std = (double*)(buf + o_std);
        }
        launch_session_agg(d_feat, n_cols, (const int*)(buf + o_start), (const int*)(buf + o_rows), n_groups, d_mean, d_std, s);
        h->launches += 1;
        if (!on_dev) {
            if (cudaMemcpyAsync(mean_out, d_mean, obytes, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaMemcpyAsync(std_out, d_std, obytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) { rc = MSHDS_ERR_CUDA; break; }
        }
        if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = MSHDS_ERR_CUDA; break; }
    } while (0);
    if (rc) h->err = "CUDA failure in mshds_aggregate_sessions";
    return rc;
}

int mshds_fp64_peak(mshds_handle* h, double* tflops) {
    if (!h || !tflops) return MSHDS_ERR_ARG;
    h->err.clear();
    CK(cudaSetDevice(h->device));
    double* scratch = nullptr;
    CK(cudaMalloc((void**)&scratch, sizeof(double) * (size_t)sm_count() * 8 * 256));
    const double tf = run_dfma_peak(scratch, h->stream, 5);
    cudaFree(scratch);
    h->launches += 6;
    if (tf < 0.0) { h->err = "DFMA microbenchmark failed"; return MSHDS_ERR_CUDA; }
    *tflops = tf;
    return MSHDS_OK;
}

int mshds_profile_enable(mshds_handle* h, int on) {
    if (!h) return MSHDS_ERR_ARG;
    h->prof_on = on != 0;
    if (on) { h->prof_acc.clear(); h->prof_order.clear(); }
    return MSHDS_OK;
}

int mshds_profile_report(mshds_handle* h, char* buf, size_t cap) {
    if (!h || !buf || cap == 0) return MSHDS_ERR_ARG;
    std::string out;
    for (auto& name : h->prof_order) {
        auto& acc = h->prof_acc[name];
        char line[256];
        snprintf(line, sizeof line, "%s\t%.6f\t%lld\n", name.c_str(), acc.first, acc.second);
        out += line;
    }
    snprintf(buf, cap, "%s", out.c_str());
    return MSHDS_OK;
}

int mshds_debug_fetch(mshds_handle* h, const char* name, int clip, void* host_buf, size_t cap_elems, size_t* n_out) {
    if (!h || !name || !n_out) return MSHDS_ERR_ARG;
    auto it = h->debug.find(name);
    if (it == h->debug.end() || clip < 0 || clip >= h->last_n) { h->err = "unknown debug name or clip"; return MSHDS_ERR_ARG; }
    const DebugEntry& e = it->second;
    CK(cudaSetDevice(h->device));
    long long start = (long long)clip * e.fixed;
    int count = e.fixed;
    if (!e.host_prefix.empty()) { start = e.host_prefix[clip]; count = (int)(e.host_prefix[clip + 1] - e.host_prefix[clip]); }
    int start_i = 0;
    if (e.start) { CK(cudaMemcpy(&start_i, e.start + clip, sizeof(int), cudaMemcpyDeviceToHost)); start = start_i; }
    if (e.count) CK(cudaMemcpy(&count, e.count + clip, sizeof(int), cudaMemcpyDeviceToHost));
    size_t nel = (size_t)count * e.stride;
    *n_out = nel;
    size_t ncopy = nel < cap_elems ? nel : cap_elems;
    if (ncopy && host_buf)
        CK(cudaMemcpy(host_buf, (const char*)e.base + (size_t)start * e.stride * e.elem_size, ncopy * e.elem_size, cudaMemcpyDeviceToHost));
    return MSHDS_OK;
}

}  // extern "C"
