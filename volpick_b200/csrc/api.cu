// Host integer math and the whole-record entry point.
//
// vp_window_starts  <- WaveformModel._cut_fragments_array (SeisBench)
// vp_annotate       <- WaveformModel.annotate + classify_aggregate for one gap-free record
//                      (call sites /root/reference/README.md:54-66, Final_models/demo.ipynb cell 13-15)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace vp;

extern "C" int64_t vp_window_count(int64_t n, int64_t L, int64_t overlap) {
    const int64_t stride = L - overlap;
    if (L <= 0 || stride <= 0 || n < L) return 0;
    int64_t cnt = (n - L) / stride + 1;              // len(arange(0, n - L + 1, stride))
    if ((cnt - 1) * stride + L < n) ++cnt;           // tail window at n - L
    return cnt;
}

extern "C" int vp_window_starts(int64_t n, int64_t L, int64_t overlap, int64_t *starts, int64_t capacity,
                                int64_t *count) {
    VP_REQUIRE(count != nullptr, VP_ERR_ARG, "vp_window_starts: null count");
    VP_REQUIRE(L > 0 && overlap >= 0 && overlap < L, VP_ERR_ARG,
               "vp_window_starts: need 0 <= overlap (%lld) < in_samples (%lld)", (long long)overlap, (long long)L);
    const int64_t cnt = vp_window_count(n, L, overlap);
    *count = cnt;
    if (cnt == 0) return VP_OK;
    VP_REQUIRE(starts != nullptr && capacity >= cnt, VP_ERR_CAPACITY, "vp_window_starts: capacity %lld < %lld windows",
               (long long)capacity, (long long)cnt);
    const int64_t stride = L - overlap;
    const int64_t nreg = (n - L) / stride + 1;
    for (int64_t i = 0; i < nreg; ++i) starts[i] = i * stride;
    if (cnt > nreg) starts[nreg] = n - L;
    return VP_OK;
}

namespace {

// device twin of vp_window_starts: starts[i] = i * stride for the regular windows, the last entry is the tail window n - L
// when the regular grid does not end at the record end (cnt regular + optional tail == n_win)
__global__ void window_starts_kernel(int64_t *__restrict__ starts, int64_t n_win, int64_t stride, int64_t tail) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_win) return;
    const int64_t v = i * stride;
    starts[i] = v <= tail ? v : tail;
}

struct Layout {
    int64_t off_trace, off_starts, off_x, off_fwd, off_y, off_annot, off_bounds, off_count, off_picks, off_scratch;
    int64_t off_x2, off_fwd2;  // extra forward lanes (chunk c runs on lane c % n_lanes): MAX_LANES - 1 slots of x_bytes / fwd_bytes
    int64_t x_bytes;
    int n_lanes;
    int64_t fwd_bytes, total;
    bool two_lanes;
    int64_t nwin, chunk, pred_len;
};

// Per host thread and device: a second stream + events for the piecewise H2D copy of host records.
struct CopyPipe {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_piece = nullptr;
    cudaStream_t lane[3] = {nullptr, nullptr, nullptr};  // extra forward lanes
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    int device = -1;
};
}  // namespace

// One record in flight between vp_annotate_begin and vp_annotate_end.
struct vp_pending {
    cudaStream_t stream = nullptr;
    const int64_t *d_count = nullptr, *d_bounds = nullptr;
    const vp_trigger *d_picks = nullptr;
    int64_t *h_pinned = nullptr;  // pinned: [0] pick count, [1..6] trim bounds
    int64_t pred_len = 0, pick_capacity = 0;
    bool empty = false;           // record shorter than one window
};

namespace {
// pinned 64-byte result blocks, recycled per host thread (cudaHostAlloc costs tens of microseconds)
thread_local std::vector<int64_t *> g_pinned_blocks;
int64_t *pinned_block_get() {
    if (!g_pinned_blocks.empty()) {
        int64_t *b = g_pinned_blocks.back();
        g_pinned_blocks.pop_back();
        return b;
    }
    int64_t *b = nullptr;
    return cudaHostAlloc((void **)&b, 8 * sizeof(int64_t), cudaHostAllocDefault) == cudaSuccess ? b : nullptr;
}
void pinned_block_put(int64_t *b) {
    if (b) g_pinned_blocks.push_back(b);
}

// One pipe per record in flight: consecutive vp_annotate_begin calls of a host thread rotate through N_PIPES pipes per
// device, so the copy stream and the extra forward lanes of record i + 1 are not queued behind those of record i.
constexpr int N_PIPES = 4;
CopyPipe *copy_pipe(unsigned rotation) {
    static thread_local CopyPipe pipes[16][N_PIPES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    CopyPipe &cp = pipes[dev][rotation % N_PIPES];
    if (cp.device != dev) {
        if (cudaStreamCreateWithFlags(&cp.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&cp.ev_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&cp.ev_piece, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&cp.ev_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < 3; ++i) {
            if (cudaStreamCreateWithFlags(&cp.lane[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&cp.ev_join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        cp.device = dev;
    }
    return &cp;
}

// measured: EQTransformer 1024 -> 4096 windows per launch group = -10 % forward time; PhaseNet host records: 2048 overlaps
// the H2D pieces best (158 vs 152 station-days/s end to end, round 1), device-resident 4096 is 3 % faster.  Round 2 (first chunk a
// quarter chunk, two records in flight): 2880 = half a station-day's 5,756 windows per lane: 319.6 device-resident / 276.9 end to end
// against 302.7 / 272.4 with 2048; 4096 is as fast device-resident but starts later behind the H2D copy
int64_t default_chunk(const vp_model *m) { return vp_model_kind(m) == VP_KIND_PHASENET ? 2880 : 4096; }

// Two forward lanes: consecutive chunks of a record are independent until the stacker, so odd chunks run on a second
// stream with their own forward workspace.  The latency-bound kernels of one chunk (LSTM recurrences, attention, the
// drain / fill of ~40 launches) then overlap the tensor-core kernels of the other.  VP_LANES=1 disables it; the
// per-kernel timing pass (vp_kernel_timing) runs single-lane so that kernel durations are not inflated by co-runners.
constexpr int MAX_LANES = 4;
int lanes_wanted(int64_t nwin, int64_t chunk) {
    static const int env = getenv("VP_LANES") ? atoi(getenv("VP_LANES")) : 2;
    const int want = std::min(std::max(env, 1), MAX_LANES);
    if (vp::g_ktimer_on || nwin <= chunk) return 1;
    return (int)std::min<int64_t>(want, (nwin + chunk - 1) / chunk);
}

int make_layout(const vp_model *m, int64_t n, const vp_annotate_params *p, int trace_on_host, int64_t pick_cap,
                Layout *lo) {
    const int64_t L = vp_model_in_samples(m);
    VP_REQUIRE(p->overlap >= 0 && p->overlap < L, VP_ERR_ARG, "annotate: need 0 <= overlap (%lld) < in_samples (%lld)",
               (long long)p->overlap, (long long)L);
    lo->nwin = vp_window_count(n, L, p->overlap);
    lo->chunk = p->chunk_windows > 0 ? p->chunk_windows : default_chunk(m);
    lo->chunk = std::min<int64_t>(std::max<int64_t>(lo->chunk, 1), 4096);
    lo->pred_len = lo->nwin ? n : 0;  // max(starts) + L == n whenever at least one window exists
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        const int64_t o = align_up(off, 256);
        off = o + bytes;
        return o;
    };
    lo->off_trace = take(trace_on_host ? 3 * n * 4 : 0);
    lo->off_starts = take(std::max<int64_t>(lo->nwin, 1) * 8);
    lo->off_x = take(std::min(lo->chunk, std::max<int64_t>(lo->nwin, 1)) * 3 * L * 4);
    lo->fwd_bytes = vp_forward_workspace_bytes(m, std::min(lo->chunk, std::max<int64_t>(lo->nwin, 1)), p->precision);
    if (lo->fwd_bytes < 0) return (int)lo->fwd_bytes;
    lo->fwd_bytes = align_up(lo->fwd_bytes, 256);
    lo->off_fwd = take(lo->fwd_bytes);
    lo->n_lanes = lanes_wanted(lo->nwin, lo->chunk);
    // the extra lanes are laid out whenever chunking happens (a workspace sized with the timing pass on must fit a later untimed call)
    static const int env_lanes = getenv("VP_LANES") ? std::min(std::max(atoi(getenv("VP_LANES")), 1), MAX_LANES) : 2;
    const int extra = lo->nwin > lo->chunk ? env_lanes - 1 : 0;
    lo->x_bytes = align_up(lo->chunk * 3 * L * 4, 256);
    lo->off_x2 = take(extra * lo->x_bytes);
    lo->off_fwd2 = take(extra * lo->fwd_bytes);
    lo->off_y = take(std::max<int64_t>(lo->nwin, 1) * 3 * L * 4);
    lo->off_annot = take(3 * std::max<int64_t>(lo->pred_len, 1) * 4);
    lo->off_bounds = take(6 * 8);
    lo->off_count = take(8);
    lo->off_picks = take(std::max<int64_t>(pick_cap, 1) * (int64_t)sizeof(vp_trigger));
    lo->off_scratch = take(vp_pick_scratch_bytes(std::max<int64_t>(lo->pred_len, 1)));
    lo->total = align_up(off, 256);
    return VP_OK;
}

}  // namespace

extern "C" int64_t vp_annotate_workspace_bytes(const vp_model *m, int64_t n_samples, const vp_annotate_params *p,
                                               int trace_on_host, int64_t pick_capacity) {
    if (!m || !p || n_samples < 0) return VP_ERR_ARG;
    Layout lo;
    int rc = make_layout(m, n_samples, p, trace_on_host, pick_capacity, &lo);
    return rc == VP_OK ? lo.total : rc;
}

extern "C" int vp_annotate_begin(vp_model *m, const void *trace, int trace_on_host, int dtype, int64_t n, int64_t ch_stride,
                                 const vp_annotate_params *p, float *annotation, int annotation_on_host,
                                 int64_t pick_capacity, void *workspace, int64_t workspace_bytes, void *stream,
                                 vp_pending **pending) {
    VP_REQUIRE(m && trace && p && workspace && pending, VP_ERR_ARG, "vp_annotate: null pointer");
    *pending = nullptr;
    VP_REQUIRE(dtype == VP_DTYPE_F32 || dtype == VP_DTYPE_I32, VP_ERR_ARG, "vp_annotate: unknown dtype %d", dtype);
    VP_REQUIRE(p->stacking == VP_STACK_AVG || p->stacking == VP_STACK_MAX, VP_ERR_ARG,
               "Stacking method %d unknown. Known methods are: 'avg' (0), 'max' (1)", p->stacking);
    VP_REQUIRE(pick_capacity >= 0, VP_ERR_ARG, "vp_annotate: negative pick capacity");
    cudaStream_t s = (cudaStream_t)stream;
    const int kind = vp_model_kind(m);
    const int64_t L = vp_model_in_samples(m);
    Layout lo;
    int rc = make_layout(m, n, p, trace_on_host, pick_capacity, &lo);
    if (rc != VP_OK) return rc;
    VP_REQUIRE(workspace_bytes >= lo.total, VP_ERR_WORKSPACE, "vp_annotate: workspace too small (%lld < %lld bytes)",
               (long long)workspace_bytes, (long long)lo.total);
    vp_pending *pd = new vp_pending();
    pd->stream = s;
    pd->pred_len = lo.pred_len;
    pd->pick_capacity = pick_capacity;
    pd->h_pinned = pinned_block_get();
    if (pd->h_pinned == nullptr) {
        delete pd;
        set_error("vp_annotate: cannot allocate the pinned result block");
        return VP_ERR_CUDA;
    }
    struct Guard {  // the pending record is released on every error path
        vp_pending *&pd;
        bool armed = true;
        ~Guard() {
            if (armed && pd) {
                pinned_block_put(pd->h_pinned);
                delete pd;
                pd = nullptr;
            }
        }
    } guard{pd};
    if (lo.nwin == 0) {  // record shorter than one window: empty output (SeisBench warns)
        pd->empty = true;
        guard.armed = false;
        *pending = pd;
        return VP_OK;
    }

    char *ws = (char *)workspace;
    const void *d_trace = trace;
    int64_t d_stride = ch_stride;
    // Host record: (3, n) with channel stride ch_stride -> packed (3, n) on the device.  The copy is issued piecewise
    // on a second stream, one piece per forward chunk (the samples that chunk's windows read), so that the H2D
    // transfer of the record (104 MB for a station-day, ~2 ms over PCIe) overlaps the network instead of preceding it.
    static thread_local unsigned g_rotation = 0;
    const unsigned rotation = g_rotation++;
    CopyPipe *pipe = nullptr;
    int64_t copied = 0;  // samples [0, copied) of every channel are on their way
    if (trace_on_host) {
        pipe = copy_pipe(rotation);
        VP_REQUIRE(pipe != nullptr, VP_ERR_CUDA, "vp_annotate: cannot create the copy stream");
        VP_CUDA_CHECK(cudaEventRecord(pipe->ev_ready, s));  // workspace reuse: earlier work of `s` reads the old record
        VP_CUDA_CHECK(cudaStreamWaitEvent(pipe->stream, pipe->ev_ready, 0));
        d_trace = ws + lo.off_trace;
        d_stride = n;
    }
    auto copy_upto = [&](int64_t end, cudaStream_t waiter) -> int {  // make samples [0, end) available to `waiter`
        if (!trace_on_host) return VP_OK;
        if (end <= copied) {  // already issued (by the other lane's chunk): order this lane after the latest piece
            if (copied > 0) VP_CUDA_CHECK(cudaStreamWaitEvent(waiter, pipe->ev_piece, 0));
            return VP_OK;
        }
        end = std::min(end, n);
        char *dst = ws + lo.off_trace;
        for (int c = 0; c < 3; ++c)
            VP_CUDA_CHECK(cudaMemcpyAsync(dst + ((int64_t)c * n + copied) * 4, (const char *)trace + ((int64_t)c * ch_stride + copied) * 4,
                                          (end - copied) * 4, cudaMemcpyHostToDevice, pipe->stream));
        copied = end;
        VP_CUDA_CHECK(cudaEventRecord(pipe->ev_piece, pipe->stream));
        VP_CUDA_CHECK(cudaStreamWaitEvent(waiter, pipe->ev_piece, 0));
        return VP_OK;
    };
    int64_t *d_starts = (int64_t *)(ws + lo.off_starts);
    std::vector<int64_t> h_starts((size_t)lo.nwin);
    {
        int64_t cnt = 0;
        rc = vp_window_starts(n, L, p->overlap, h_starts.data(), lo.nwin, &cnt);
        if (rc != VP_OK) return rc;
        // the device copy is generated in place (same integer rule): no pageable H2D copy on the path of a record
        window_starts_kernel<<<(unsigned)((lo.nwin + 255) / 256), 256, 0, s>>>(d_starts, lo.nwin, L - p->overlap, n - L);
        VP_LAUNCH_CHECK();
    }
    float *d_y = (float *)(ws + lo.off_y);
    CopyPipe *lanes = lo.n_lanes > 1 ? copy_pipe(rotation) : nullptr;
    VP_REQUIRE(lo.n_lanes == 1 || lanes != nullptr, VP_ERR_CUDA, "vp_annotate: cannot create the extra forward lanes");
    if (lanes) {  // fork: the extra lanes start after the window starts are uploaded (and after earlier users of the workspace)
        VP_CUDA_CHECK(cudaEventRecord(lanes->ev_fork, s));
        for (int i = 1; i < lo.n_lanes; ++i) VP_CUDA_CHECK(cudaStreamWaitEvent(lanes->lane[i - 1], lanes->ev_fork, 0));
    }
    const int taper = ((kind == VP_KIND_EQTRANSFORMER) ? VP_PRE_TAPER : 0) | (p->norm_detrend ? VP_PRE_DETREND : 0);
    static const bool fused_off = getenv("VP_FUSED_SLICE") && atoi(getenv("VP_FUSED_SLICE")) == 0;  // debugging aid
    const bool fused_slice = !fused_off && (p->precision == VP_PREC_F16X3 || p->precision == VP_PREC_BF16);
    int64_t chunk_no = 0;
    // PhaseNet host records: the first chunk is a quarter chunk, so that the network starts after a quarter of the first H2D
    // piece (measured: PhaseNet 185 -> 189 station-days/s end to end; EQTransformer, whose step is six times longer: no gain).
    static const bool ramp_off = getenv("VP_RAMP") && atoi(getenv("VP_RAMP")) == 0;
    const int64_t first_chunk =
        (trace_on_host && !ramp_off && kind == VP_KIND_PHASENET && lo.nwin > lo.chunk && lo.chunk >= 64) ? lo.chunk / 4 : lo.chunk;
    int64_t nw = 0;
    for (int64_t w0 = 0; w0 < lo.nwin; w0 += nw, ++chunk_no) {
        nw = std::min(chunk_no == 0 ? first_chunk : lo.chunk, lo.nwin - w0);
        const int ln = lanes ? (int)(chunk_no % lo.n_lanes) : 0;
        cudaStream_t cs = ln ? lanes->lane[ln - 1] : s;
        float *d_x = (float *)(ws + (ln ? lo.off_x2 + (ln - 1) * lo.x_bytes : lo.off_x));
        char *fwd_ws = ws + (ln ? lo.off_fwd2 + (ln - 1) * lo.fwd_bytes : lo.off_fwd);
        {   // the record samples this chunk's windows read (starts ascend; the tail window ends at n)
            int64_t need = 0;
            for (int64_t i = w0 + nw - 1; i >= w0 && i >= w0 + nw - 2; --i) need = std::max(need, h_starts[(size_t)i] + L);
            rc = copy_upto(w0 + nw >= lo.nwin ? n : need, cs);
            if (rc != VP_OK) return rc;
        }
        // vp_stack discards the blinded margins of every window: the forward need not compute them
        const int64_t keep_lo = std::min(std::max<int64_t>(p->blinding[0], 0), L);
        const int64_t keep_hi = std::min(std::max<int64_t>(L - p->blinding[1], keep_lo), L);
        if (fused_slice) {  // K1 lives inside the first encoder kernel: the fp32 windows are never materialised
            rc = vp_slice_forward(m, d_trace, dtype, n, d_stride, d_starts + w0, nw, p->peak_scope, taper, d_y + w0 * 3 * L,
                                  fwd_ws, lo.fwd_bytes, p->precision, keep_lo, keep_hi, cs);
            if (rc != VP_OK) return rc;
            continue;
        }
        rc = vp_slice_normalize(d_trace, dtype, n, d_stride, d_starts + w0, nw, L, p->peak_scope, taper, d_x, cs);
        if (rc != VP_OK) return rc;
        rc = vp_forward_range(m, d_x, nw, d_y + w0 * 3 * L, fwd_ws, lo.fwd_bytes, p->precision, keep_lo, keep_hi, cs);
        if (rc != VP_OK) return rc;
    }
    for (int i = 1; lanes && i < lo.n_lanes; ++i) {  // join before the stacker reads every window
        VP_CUDA_CHECK(cudaEventRecord(lanes->ev_join[i - 1], lanes->lane[i - 1]));
        VP_CUDA_CHECK(cudaStreamWaitEvent(s, lanes->ev_join[i - 1], 0));
    }
    float *d_annot = (float *)(ws + lo.off_annot);
    rc = vp_stack(d_y, d_starts, lo.nwin, L, 3, p->overlap, p->blinding[0], p->blinding[1], p->stacking, d_annot,
                  lo.pred_len, s);
    if (rc != VP_OK) return rc;
    // _trim_nan bounds + trigger_onset / first-argmax picks of all labels: one pass over the annotation
    int64_t *d_bounds = (int64_t *)(ws + lo.off_bounds);
    int64_t *d_count = (int64_t *)(ws + lo.off_count);
    vp_trigger *d_picks = (vp_trigger *)(ws + lo.off_picks);
    VP_CUDA_CHECK(cudaMemsetAsync(d_count, 0, 8, s));
    {
        float thr_on[3], thr_off[3];
        for (int c = 0; c < 3; ++c) {
            thr_on[c] = p->threshold[c];
            thr_off[c] = p->threshold[c] / 2;
        }
        rc = vp_pick_labels(d_annot, 3, lo.pred_len, thr_on, thr_off, d_picks, pick_capacity, d_count, d_bounds, s);
        if (rc != VP_OK) return rc;
    }
    VP_CUDA_CHECK(cudaMemcpyAsync(pd->h_pinned, d_count, 8, cudaMemcpyDeviceToHost, s));
    VP_CUDA_CHECK(cudaMemcpyAsync(pd->h_pinned + 1, d_bounds, 6 * 8, cudaMemcpyDeviceToHost, s));
    if (annotation)
        VP_CUDA_CHECK(cudaMemcpyAsync(annotation, d_annot, 3 * lo.pred_len * 4,
                                      annotation_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
    pd->d_count = d_count;
    pd->d_bounds = d_bounds;
    pd->d_picks = d_picks;
    guard.armed = false;
    *pending = pd;
    return VP_OK;
}

extern "C" int vp_annotate_end(vp_pending *pd, vp_trigger *picks, int64_t pick_capacity, int64_t *n_picks, int64_t *trim) {
    VP_REQUIRE(pd && n_picks && trim, VP_ERR_ARG, "vp_annotate_end: null pointer");
    struct Release {
        vp_pending *pd;
        ~Release() {
            pinned_block_put(pd->h_pinned);
            delete pd;
        }
    } release{pd};
    *n_picks = 0;
    for (int i = 0; i < 3; ++i) {
        trim[2 * i] = pd->pred_len;
        trim[2 * i + 1] = -1;
    }
    if (pd->empty) return VP_OK;
    VP_CUDA_CHECK(cudaStreamSynchronize(pd->stream));
    const int64_t h_count = pd->h_pinned[0];
    std::memcpy(trim, pd->h_pinned + 1, 6 * sizeof(int64_t));
    *n_picks = h_count;
    const int64_t cap = std::min(pick_capacity, pd->pick_capacity);
    VP_REQUIRE(h_count <= cap, VP_ERR_CAPACITY, "vp_annotate: %lld picks exceed the pick capacity %lld", (long long)h_count,
               (long long)cap);
    if (h_count > 0) {
        VP_REQUIRE(picks != nullptr, VP_ERR_ARG, "vp_annotate: pick buffer missing");
        VP_CUDA_CHECK(cudaMemcpyAsync(picks, pd->d_picks, (size_t)h_count * sizeof(vp_trigger), cudaMemcpyDeviceToHost, pd->stream));
        VP_CUDA_CHECK(cudaStreamSynchronize(pd->stream));
        std::sort(picks, picks + h_count, [](const vp_trigger &a, const vp_trigger &b) {
            return a.label != b.label ? a.label < b.label : a.s0 < b.s0;
        });
    }
    return VP_OK;
}

extern "C" int vp_annotate(vp_model *m, const void *trace, int trace_on_host, int dtype, int64_t n, int64_t ch_stride,
                           const vp_annotate_params *p, float *annotation, int annotation_on_host, vp_trigger *picks,
                           int64_t pick_capacity, int64_t *n_picks, int64_t *trim, void *workspace,
                           int64_t workspace_bytes, void *stream) {
    VP_REQUIRE(n_picks && trim, VP_ERR_ARG, "vp_annotate: null pointer");
    VP_REQUIRE(pick_capacity == 0 || picks, VP_ERR_ARG, "vp_annotate: pick buffer missing");
    vp_pending *pd = nullptr;
    int rc = vp_annotate_begin(m, trace, trace_on_host, dtype, n, ch_stride, p, annotation, annotation_on_host, pick_capacity,
                               workspace, workspace_bytes, stream, &pd);
    if (rc != VP_OK) return rc;
    return vp_annotate_end(pd, picks, pick_capacity, n_picks, trim);
}
