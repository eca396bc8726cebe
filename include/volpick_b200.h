/*
 * volpick_b200 -- C ABI of the B200-native continuous-waveform picking path.
 *
 * This is the drop-in boundary for the path that the reference runs through SeisBench:
 *   picker = sbm.EQTransformer.from_pretrained("volpick"); picker.classify(stream, ...).picks
 *   (/root/reference/README.md:46-66, /root/reference/Final_models/demo.ipynb cells 7-15).
 * The reference has no C plugin ABI (it is pure Python calling SeisBench); each entry point
 * below names the SeisBench / reference function it replaces.  All signatures are plain C:
 * raw pointers, sizes and a CUDA stream passed as void* (cudaStream_t).  No torch types.
 *
 * Conventions
 *   - every function returns VP_OK (0) or a negative VP_ERR_* code; vp_last_error() gives the
 *     thread-local message of the last failure.  Pick-buffer overflow is an error, never a
 *     silent truncation.  There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with VP_ERR_CUDA.
 *   - "device" pointers are CUDA device memory of the model's device; the caller owns every
 *     buffer (inputs are never modified); the model handle owns only the folded weights.
 *   - activations / predictions are planar: windows (B, 3, L), predictions (B, 3, L), stacked
 *     annotation (3, pred_len).  Label order: EQTransformer Detection,P,S; PhaseNet P,S,N
 *     (/root/reference/volpick/model/eval_taks0.py:68-72,131-134).
 *   - launches go to the caller's stream; a handle is immutable after creation and may be
 *     shared by threads that use different streams and different workspaces.
 */
#ifndef VOLPICK_B200_H
#define VOLPICK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VP_API __attribute__((visibility("default")))
#else
#define VP_API
#endif

#define VP_OK 0
#define VP_ERR_ARG (-1)
#define VP_ERR_CUDA (-2)
#define VP_ERR_CAPACITY (-3)
#define VP_ERR_WORKSPACE (-4)
#define VP_ERR_UNSUPPORTED (-5)

enum { VP_KIND_EQTRANSFORMER = 0, VP_KIND_PHASENET = 1 };
enum { VP_STACK_AVG = 0, VP_STACK_MAX = 1 };
/* fp32: CUDA-core FFMA kernels.  f16x3: tcgen05 tensor cores on fp16 hi/lo split operands, three MMAs per
 * K step into fp32 TMEM accumulators (fp32-equivalent).  bf16: tcgen05, one bf16 pass. */
enum { VP_PREC_FP32 = 0, VP_PREC_F16X3 = 1, VP_PREC_BF16 = 2 };
enum { VP_DTYPE_F32 = 0, VP_DTYPE_I32 = 1 };
/* Amplitude scope of the "peak" window normalisation (SeisBench annotate_batch_pre).  PER_CHANNEL: every component by
 * its own peak -- EQTransformer(norm_amp_per_comp=True), PhaseNet, and this library's reading of norm="peak" (it is the
 * normalisation the volpick weights were trained with, /root/reference/volpick/model/models.py:449-451,853-855).
 * PER_WINDOW: one peak over the three components of a window (the alternative reading). */
enum { VP_PEAK_PER_CHANNEL = 0, VP_PEAK_PER_WINDOW = 1 };
/* Window pre-processing flags: 6-sample cosine taper (EQTransformer), linear detrend after the demean
 * (SeisBench norm_detrend=True: scipy.signal.detrend per component). */
enum { VP_PRE_TAPER = 1, VP_PRE_DETREND = 2 };

typedef struct vp_model vp_model;

/* One trigger: SeisBench Pick / Detection before the index->time conversion. */
typedef struct vp_trigger {
    int64_t s0;     /* first sample > threshold inside the run            (start_time) */
    int64_t s1;     /* last sample of the run of samples > threshold / 2  (end_time)   */
    int64_t s_peak; /* s0 + argmax(x[s0..s1]), first maximum              (peak_time)  */
    float value;    /* x[s_peak]                                          (peak_value) */
    int32_t label;  /* label index (column of the annotation)                          */
} vp_trigger;

/* kwargs of WaveformModel.annotate / classify (/root/reference/README.md:54-66). */
typedef struct vp_annotate_params {
    int64_t overlap;       /* samples                                                   */
    int64_t blinding[2];   /* (pre, post) samples set to NaN in every window            */
    int32_t stacking;      /* VP_STACK_AVG | VP_STACK_MAX                               */
    int32_t precision;     /* VP_PREC_*                                                 */
    int32_t peak_scope;    /* VP_PEAK_PER_CHANNEL (default) | VP_PEAK_PER_WINDOW        */
    int32_t chunk_windows; /* windows per forward launch group; <= 0: library default   */
    float threshold[3];    /* per label; <= 0 or NaN: no picks for that label           */
    int32_t norm_detrend;  /* != 0: linear detrend of every window component (VP_PRE_DETREND) */
} vp_annotate_params;

/* ---- misc ---------------------------------------------------------------------------- */
VP_API int vp_version(void);
VP_API const char *vp_last_error(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset. */
VP_API int64_t vp_launch_count(int reset);
/* Per-kernel-class timing with CUDA events on the launching stream (calling thread; measurement aid of bench.py).
 * vp_kernel_timing(1) clears the record and starts bracketing every launch; vp_kernel_timing(0) stops and clears.
 * vp_kernel_timing_read synchronises on the recorded events and returns the summed duration (ms) and the
 * number of launches of class `kclass` (index into the comma-separated vp_kernel_class_names()). */
VP_API int vp_kernel_timing(int enable);
VP_API const char *vp_kernel_class_names(void);
VP_API int vp_kernel_timing_read(int kclass, double *total_ms, int64_t *launches);

/* ---- model handle: SeisBenchModel.from_pretrained -> load_state_dict ------------------- */
/* weights: every float tensor of the SeisBench state dict, concatenated in state-dict order
 * (num_batches_tracked dropped): 378,823 floats for EQTransformer, 269,675 for PhaseNet
 * (/root/reference/Final_models/volpick/{eqtransformer,phasenet}/volpick.pt.v1).  BatchNorm
 * (eps 1e-3) is folded in double precision on the host. */
VP_API int vp_model_create(int kind, const float *weights, int64_t n_floats, int device, vp_model **out);
VP_API int vp_model_destroy(vp_model *m);
VP_API int vp_model_kind(const vp_model *m);
VP_API int vp_model_in_samples(const vp_model *m); /* 6000 / 3001 */
VP_API int64_t vp_model_expected_floats(int kind);

/* ---- host integer math: WaveformModel._cut_fragments_array ----------------------------- */
VP_API int64_t vp_window_count(int64_t n_samples, int64_t in_samples, int64_t overlap);
VP_API int vp_window_starts(int64_t n_samples, int64_t in_samples, int64_t overlap, int64_t *starts, int64_t capacity,
                     int64_t *count);
VP_API int64_t vp_coverage(int64_t in_samples, int64_t overlap); /* ceil(L / (L - overlap) + 1) */

/* ---- stage kernels (all pointers are device pointers) ---------------------------------- */
/* Stream pre-filter (annotate_stream_pre with filter_args / filter_kwargs; reference use:
 * /root/reference/model_training/test_onephase.ipynb cell 43, /root/reference/volpick/data/utils.py:702-704).
 * Cascaded second-order sections with scipy.signal.sosfilt semantics in float64 (what obspy.signal.filter.highpass /
 * lowpass / bandpass / bandstop evaluate); zerophase != 0: forward pass, then the same cascade over the reversed signal.
 * x: device (n_channels, n_samples) f32 / i32 with channel stride ch_stride; sos: HOST double[n_sections][6] =
 * (b0, b1, b2, a0, a1, a2) per section (<= 8 sections); y: device float32 (n_channels, n_samples), packed. */
VP_API int64_t vp_sosfilt_workspace_bytes(int64_t n_samples, int n_channels, int n_sections);
VP_API int vp_sosfilt(const void *x, int dtype, int64_t n_samples, int64_t ch_stride, int n_channels, const double *sos,
               int n_sections, int zerophase, float *y, void *workspace, int64_t workspace_bytes, void *stream);
/* _cut_fragments_array + annotate_batch_pre: gather windows, demean, peak-normalise (+1e-10),
 * EQTransformer 6-sample cosine taper.  trace: (3, n) with channel stride ch_stride elements,
 * f32 or i32 counts.  pre_flags: VP_PRE_TAPER | VP_PRE_DETREND.  out: (n_windows, 3, L) f32. */
VP_API int vp_slice_normalize(const void *trace, int dtype, int64_t n_samples, int64_t ch_stride, const int64_t *starts,
                       int64_t n_windows, int64_t in_samples, int peak_scope, int pre_flags, float *out, void *stream);

/* EQTransformer.forward / PhaseNet.forward on pre-normalised windows.
 * x: (n_windows, 3, L) -> y: (n_windows, 3, L) probabilities (sigmoid heads / channel softmax). */
VP_API int64_t vp_forward_workspace_bytes(const vp_model *m, int64_t n_windows, int precision);
VP_API int vp_forward(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace, int64_t workspace_bytes,
               int precision, void *stream);
/* vp_forward for a caller that discards the first keep_lo and the samples from keep_hi on of every window
 * (annotate_batch_post blinding, /root/reference/README.md:58): only y[:, :, keep_lo:keep_hi] is guaranteed to be
 * written; work that only feeds the discarded samples may be skipped. */
VP_API int vp_forward_range(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace,
                            int64_t workspace_bytes, int precision, int64_t keep_lo, int64_t keep_hi, void *stream);
/* vp_slice_normalize + vp_forward_range in one call, for the tensor-core modes of both models: the windows are cut
 * and normalised from the record (device pointer, (3, n_samples) with channel stride ch_stride) inside the kernel
 * that also runs the first conv (EQTransformer encoder.convs.0 + pool, PhaseNet inc + in_bn), so the fp32 windows
 * never reach HBM.  Same results as the two-call sequence
 * within the mode's tolerance (the first conv runs in fp32 instead of f16x3 / bf16). */
VP_API int vp_slice_forward(vp_model *m, const void *trace, int dtype, int64_t n_samples, int64_t ch_stride,
                            const int64_t *starts, int64_t n_windows, int peak_scope, int pre_flags, float *y, void *workspace,
                            int64_t workspace_bytes, int precision, int64_t keep_lo, int64_t keep_hi, void *stream);
/* Debug/parity: run the forward and copy the named intermediate activation (device->device) into
 * tap_out (capacity in floats); *tap_floats receives its size.  Names: see vp_forward_tap_names(). */
VP_API int vp_forward_tap(vp_model *m, const float *x, int64_t n_windows, float *y, void *workspace,
                   int64_t workspace_bytes, int precision, const char *tap_name, float *tap_out,
                   int64_t tap_capacity, int64_t *tap_floats, void *stream);
VP_API const char *vp_forward_tap_names(const vp_model *m); /* comma separated */

/* Debug/parity: ONE Conv1d through the tcgen05 path (volpick_b200/csrc/tcconv.cu).  x: device fp32
 * (NS, CIN, T_in); w_host / bias_host: HOST fp32 (COUT, CIN, K) / (COUT) or NULL; y: device fp32
 * (NS, COUT, T_out).  mode 0: 'same' conv; 1: x2 nearest up-sampling folded into the weights; 2: conv on
 * the x2 up-sampled input minus `crop` trailing samples; 3: the 'same' conv + ReLU + MaxPool1d(2) on the time-folded
 * [T / 4][4 C] view (encoder.convs.1 / .2: CIN 8 | 16, COUT 16, odd K <= 9, T_in % 4 == 0, act 1, pool 2).
 * act: 0 none, 1 ReLU, 2 sigmoid.  pool: 1 | 2. */
VP_API int vp_tcconv_debug(const float *x, int NS, int CIN, int T_in, const float *w_host, const float *bias_host, int COUT,
                           int K, int mode, int crop, int act, int pool, int precision, float *y, void *stream);

/* Profiling aid: device time (ms per launch) of one tcgen05 conv layer on synthetic data. */
VP_API int vp_tcconv_bench(int NS, int CIN, int T_in, int COUT, int K, int mode, int crop, int pool, int precision,
                           int out_fmt, int iters, float *ms_out);

/* annotate_batch_post (blinding) + _reassemble_blocks_array: y (n_windows,3,L) -> out (3,pred_len). */
VP_API int vp_stack(const float *y, const int64_t *starts, int64_t n_windows, int64_t in_samples, int n_labels,
             int64_t overlap, int64_t blind0, int64_t blind1, int mode, float *out, int64_t pred_len, void *stream);

/* _trim_nan: first / last non-NaN index per label (first = pred_len, last = -1 when all NaN).
 * bounds: device int64[2 * n_labels] = {first_0, last_0, first_1, ...}. */
VP_API int vp_nan_bounds(const float *annotation, int n_labels, int64_t pred_len, int64_t *bounds, void *stream);

/* picks_from_annotations / detections_from_annotations (trigger_onset + first argmax;
 * /root/reference/volpick/model/eval_taks0.py:46-56).  Appends to picks[*count...] (device),
 * unordered; *count (device int64) may exceed capacity -> caller must treat as overflow.
 * One pass over the trace (tiles staged in shared memory); scratch is unused and kept for ABI
 * stability (vp_pick_scratch_bytes returns a token size, NULL is accepted). */
VP_API int64_t vp_pick_scratch_bytes(int64_t n_samples);
VP_API int vp_pick(const float *trace, int64_t n_samples, float thr_on, float thr_off, int label, vp_trigger *picks,
            int64_t capacity, int64_t *count, void *scratch, int64_t scratch_bytes, void *stream);

/* classify_aggregate for a whole (n_labels, pred_len) annotation in ONE launch: _trim_nan bounds
 * (as vp_nan_bounds; bounds may be NULL) and the picks of every label whose thr_on[c] > 0
 * (thr_on / thr_off: HOST arrays of n_labels floats; n_labels <= 4).  Appends like vp_pick. */
VP_API int vp_pick_labels(const float *annotation, int n_labels, int64_t pred_len, const float *thr_on,
                   const float *thr_off, vp_trigger *picks, int64_t capacity, int64_t *count, int64_t *bounds,
                   void *stream);

/* Window-level picks of the reference's evaluate() path (/root/reference/volpick/model/eval_taks0.py:20-140): for every
 * window b and every label c with thr_on[c] > 0, trigger_onset + first argmax on y[b, c, lo_b:hi_b] (y: device
 * (n_windows, n_labels, in_samples) as written by vp_forward; borders: DEVICE int64 (n_windows, 2) = window_borders, or
 * NULL for whole windows).  Appends vp_trigger records with indices relative to lo_b and
 * label = b * n_labels + c; unordered, *count may exceed capacity (overflow) as for vp_pick.
 * thr_on / thr_off: HOST arrays of n_labels floats. */
VP_API int vp_pick_windows(const float *y, int64_t n_windows, int n_labels, int64_t in_samples, const int64_t *borders,
                    const float *thr_on, const float *thr_off, vp_trigger *picks, int64_t capacity, int64_t *count,
                    void *stream);

/* ---- the whole path for one gap-free record: WaveformModel.annotate + classify_aggregate -- */
VP_API int64_t vp_annotate_workspace_bytes(const vp_model *m, int64_t n_samples, const vp_annotate_params *p,
                                    int trace_on_host, int64_t pick_capacity);
/* trace: (3, n) f32/i32, on the host (pinned recommended; copied inside) or on the device.
 * annotation: optional (3, pred_len) output, host or device (NULL to skip).
 * picks: HOST buffer; sorted by (label, s0); indices relative to the record start.
 * trim: HOST int64[6] = per label {first non-NaN, last non-NaN} of the stacked annotation.
 * Synchronises the stream before returning. */
VP_API int vp_annotate(vp_model *m, const void *trace, int trace_on_host, int dtype, int64_t n_samples, int64_t ch_stride,
                const vp_annotate_params *p, float *annotation, int annotation_on_host, vp_trigger *picks,
                int64_t pick_capacity, int64_t *n_picks, int64_t *trim, void *workspace, int64_t workspace_bytes,
                void *stream);

/* The same in two halves, to keep several records in flight on different streams (H2D of record i + 1 under the compute
 * of record i): vp_annotate_begin enqueues everything on `stream` and returns without waiting (host buffers must be pinned
 * for that; workspace and buffers stay owned by the caller until _end); vp_annotate_end waits for the stream, copies the
 * picks (HOST buffer, sorted by (label, s0)), fills n_picks / trim and releases the pending record -- also on error.
 * Every begin must be matched by exactly one end on the same host thread. */
typedef struct vp_pending vp_pending;
VP_API int vp_annotate_begin(vp_model *m, const void *trace, int trace_on_host, int dtype, int64_t n_samples, int64_t ch_stride,
                      const vp_annotate_params *p, float *annotation, int annotation_on_host, int64_t pick_capacity,
                      void *workspace, int64_t workspace_bytes, void *stream, vp_pending **pending);
VP_API int vp_annotate_end(vp_pending *pending, vp_trigger *picks, int64_t pick_capacity, int64_t *n_picks, int64_t *trim);

#ifdef __cplusplus
}
#endif
#endif /* VOLPICK_B200_H */
