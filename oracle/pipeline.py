"""NumPy restatement of SeisBench's annotate/classify pipeline (TEST INFRASTRUCTURE).

Follows ``seisbench/models/base.py`` (``WaveformModel._cut_fragments_array``,
``_reassemble_blocks_array``, ``_predictions_to_stream``, ``_trim_nan``,
``picks_from_annotations``, ``detections_from_annotations``),
``seisbench/models/{eqtransformer,phasenet}.py`` (``annotate_batch_pre/post``) and
ObsPy ``obspy/signal/trigger.py::trigger_onset`` as restated in SURVEY.md
Appendix C.  Reference-side pins: the kwargs at /root/reference/README.md:54-66,
the pick rule at /root/reference/volpick/model/eval_taks0.py:46-56, window
lengths and normalisation at /root/reference/volpick/model/models.py:445-452,849-856.

PARITY UNPINNED (see oracle/__init__.py): no golden vectors exist in the reference.
"""
from __future__ import annotations

import warnings
from collections import deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------- C.1 window starts


def window_starts(n_samples: int, in_samples: int, overlap: int) -> np.ndarray:
    """``_cut_fragments_array``: arange(0, N-L+1, L-overlap) plus a tail window at N-L."""
    starts = np.arange(0, n_samples - in_samples + 1, in_samples - overlap, dtype=np.int64)
    if len(starts) == 0:
        return starts
    if starts[-1] + in_samples < n_samples:
        starts = np.concatenate([starts, np.array([n_samples - in_samples], dtype=np.int64)])
    return starts


def cut_windows(trace: np.ndarray, starts: np.ndarray, in_samples: int) -> np.ndarray:
    """(C, N) -> (B, C, L) float32 windows (the reference stacks float32 batches before the H2D copy)."""
    return np.stack([trace[:, s : s + in_samples] for s in starts], axis=0).astype(np.float32)


# --------------------------------------------------------------------------- C.2 pre-processing

EQT_TAPER_LEN = 6


def eqt_taper() -> np.ndarray:
    return (0.5 * (1 + np.cos(np.linspace(np.pi, 2 * np.pi, EQT_TAPER_LEN)))).astype(np.float32)


def prenorm(batch: np.ndarray, kind: str, norm: str = "peak", peak_scope: str = "channel", detrend: bool = False) -> np.ndarray:
    """``annotate_batch_pre`` in fp32: demean, [linear detrend], amplitude normalise, (EQT) 6-sample cosine taper.

    ``peak_scope``: "channel" (per window and channel, Appendix C.2 -- and the training-side pin
    models.py:449-451,853-855; SeisBench ``norm_amp_per_comp=True``) or "window" (one peak/std over all
    channels of a window; the uncertain SeisBench variant listed in Appendix D #6).
    ``detrend``: SeisBench ``norm_detrend=True`` -- ``scipy.signal.detrend`` (least-squares line, float64)
    of every component after the demean.
    """
    x = batch.astype(np.float32, copy=True)
    # mean accumulated in float64, applied in float32: SeisBench 0.4 takes it in NumPy's float64 (int32 / float64 traces),
    # torch's CPU mean of later releases accumulates float32 input in double
    x = x - x.mean(axis=-1, keepdims=True, dtype=np.float64).astype(np.float32)
    if detrend:
        from scipy.signal import detrend as _detrend

        x = _detrend(x.astype(np.float64), axis=-1, type="linear").astype(np.float32)
    axes = (-1,) if peak_scope == "channel" else (-1, -2)
    if norm == "peak":
        amp = np.abs(x).max(axis=axes, keepdims=True)
    elif norm == "std":
        amp = x.std(axis=axes, keepdims=True, dtype=np.float32)
    else:
        raise ValueError(norm)
    x = x / (amp + np.float32(1e-10))
    if kind == "eqtransformer":
        tap = eqt_taper()
        x[:, :, :EQT_TAPER_LEN] *= tap
        x[:, :, -EQT_TAPER_LEN:] *= tap[::-1]
    return x.astype(np.float32)


# --------------------------------------------------------------------------- C.3 batch post


def blind(y: np.ndarray, blinding: Tuple[int, int]) -> np.ndarray:
    """``annotate_batch_post``: y is (B, L, C) heads-last; first b0 / last b1 samples := NaN."""
    y = y.copy()
    b0, b1 = blinding
    if b0 > 0:
        y[:, :b0] = np.nan
    if b1 > 0:
        y[:, -b1:] = np.nan
    return y


# --------------------------------------------------------------------------- C.4 stacking


def coverage(in_samples: int, overlap: int) -> int:
    return int(np.ceil(in_samples / (in_samples - overlap) + 1))


def reassemble(preds: np.ndarray, starts: np.ndarray, in_samples: int, overlap: int, stacking: str) -> np.ndarray:
    """``_reassemble_blocks_array``: (B, L, C) + starts -> (pred_len, C) via the NaN slot buffer."""
    cov = coverage(in_samples, overlap)
    pred_length = int(np.max(starts) + in_samples)
    merge = np.zeros_like(preds[0], shape=(pred_length, preds.shape[2], cov)) * np.nan
    for i, (pred, start) in enumerate(zip(preds, starts)):
        s = int(start)
        merge[s : s + pred.shape[0], :, i % cov] = pred
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        if stacking == "avg":
            return np.nanmean(merge, axis=-1)
        if stacking == "max":
            return np.nanmax(merge, axis=-1)
    raise ValueError(f"Stacking method {stacking} unknown. Known methods are: 'avg', 'max'")


# --------------------------------------------------------------------------- C.5 trim


def trim_nan(x: np.ndarray) -> Tuple[np.ndarray, int, int]:
    """``_trim_nan``: strip leading/trailing NaNs; returns (trimmed, n_front, n_back)."""
    mask_forward = np.cumprod(np.isnan(x)).astype(bool)
    x = x[~mask_forward]
    mask_backward = np.cumprod(np.isnan(x)[::-1])[::-1].astype(bool)
    x = x[~mask_backward]
    return x, int(np.sum(mask_forward.astype(int))), int(np.sum(mask_backward.astype(int)))


# --------------------------------------------------------------------------- C.6 picks


def trigger_onset(charfct: np.ndarray, thres1: float, thres2: float) -> np.ndarray:
    """ObsPy ``trigger_onset`` (classic deque implementation, max_len = inf)."""
    ind1 = np.where(charfct > thres1)[0]
    if len(ind1) == 0:
        return np.empty((0, 2), dtype=np.int64)
    ind2 = np.where(charfct > thres2)[0]
    on = deque([ind1[0]])
    of = deque([-1])
    ind2_ = np.empty_like(ind2, dtype=bool)
    ind2_[:-1] = np.diff(ind2) > 1
    ind2_[-1] = True
    of.extend(ind2[ind2_].tolist())
    on.extend(ind1[np.where(np.diff(ind1) > 1)[0] + 1].tolist())
    of.extend([ind2[-1]])
    pick = []
    while on[-1] > of[0]:
        while on[0] <= of[0]:
            on.popleft()
        while of[0] < on[0]:
            of.popleft()
        pick.append([on[0], of[0]])
    return np.array(pick, dtype=np.int64).reshape(-1, 2)


def trigger_runs(x: np.ndarray, thres1: float, thres2: float) -> np.ndarray:
    """Run form of the same rule (Appendix C.6): one trigger per maximal run of ``x > thres2``
    containing a sample ``> thres1``; on = first such sample, off = last sample of the run."""
    out = []
    n = len(x)
    i = 0
    while i < n:
        if x[i] > thres2:
            j = i
            on = -1
            while j < n and x[j] > thres2:
                if on < 0 and x[j] > thres1:
                    on = j
                j += 1
            if on >= 0:
                out.append([on, j - 1])
            i = j
        else:
            i += 1
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def picks_from_trace(x: np.ndarray, threshold: float) -> List[Tuple[int, int, int, float]]:
    """``picks_from_annotations`` body (eval_taks0.py:46-56): (s0, s1, s_peak, peak_value) per trigger."""
    res = []
    for s0, s1 in trigger_onset(x, threshold, threshold / 2):
        seg = x[s0 : s1 + 1]
        res.append((int(s0), int(s1), int(s0 + np.argmax(seg)), float(np.max(seg))))
    return res


# --------------------------------------------------------------------------- stream pre-filter
def design_sos(ftype: str, df: float, corners: int = 4, **kw) -> np.ndarray:
    """Second-order sections of obspy.signal.filter.{highpass,lowpass,bandpass,bandstop} (Butterworth;
    ``iirfilter(corners, f / (0.5 df), btype, ftype="butter", output="zpk")`` + ``zpk2sos``), the filters a SeisBench model
    applies when it carries ``filter_args`` / ``filter_kwargs`` (reference use: model_training/test_onephase.ipynb cell 43:
    ``filter_args=["highpass"], filter_kwargs={"freq": 0.5, "corners": 2, "zerophase": True}``)."""
    from scipy.signal import iirfilter, zpk2sos

    fe = 0.5 * df
    if ftype in ("highpass", "lowpass"):
        wn = kw["freq"] / fe
        btype = ftype
    elif ftype in ("bandpass", "bandstop"):
        wn = [kw["freqmin"] / fe, kw["freqmax"] / fe]
        btype = "band" if ftype == "bandpass" else "bandstop"
    else:
        raise ValueError(f"unknown filter type {ftype!r}")
    z, p, k = iirfilter(corners, wn, btype=btype, ftype="butter", output="zpk")
    return np.ascontiguousarray(zpk2sos(z, p, k), dtype=np.float64)


def sosfilt_record(x: np.ndarray, sos: np.ndarray, zerophase: bool = False) -> np.ndarray:
    """obspy filter body: ``sosfilt(sos, data)``; zerophase: ``sosfilt(sos, firstpass[::-1])[::-1]``.  float64 -> float32."""
    from scipy.signal import sosfilt

    y = sosfilt(sos, np.asarray(x, dtype=np.float64), axis=-1)
    if zerophase:
        y = sosfilt(sos, y[..., ::-1], axis=-1)[..., ::-1]
    return np.ascontiguousarray(y, dtype=np.float32)


def fft_resample(data: np.ndarray, rate: float, new_rate: float, window: str = "hann") -> np.ndarray:
    """ObsPy ``Trace.resample(new_rate, window="hann", no_filter=True)`` -- what SeisBench's ``WaveformModel.resample`` calls for a
    sampling rate that is NOT an integer multiple of the model's -- restated with ``numpy.fft`` (ObsPy is not installable here: the
    steps are recalled from obspy/core/trace.py, "parity unpinned" like the rest of this package; tests/test_vs_seisbench.py
    compares with the real one when it imports).  Steps: real FFT; the spectrum times ``ifftshift(get_window(window, npts))`` (1 at
    DC, 0 at Nyquist for "hann"); real and imaginary parts interpolated linearly (``np.interp``) from the frequency grid
    ``k / (npts * delta)`` onto ``k * new_rate / num``, ``num = int(npts / (rate / new_rate))``; inverse real FFT of length num;
    times ``num / npts``.  float64 in, float64 out."""
    from scipy.signal import get_window

    x = np.asarray(data, dtype=np.float64)
    npts = x.shape[-1]
    factor = rate / float(new_rate)
    spec = np.fft.rfft(x)  # scipy.fftpack.rfft's [y0, Re y1, Im y1, ...] regrouped: x_r = Re, x_i = Im (Im y0 = Im y_nyq = 0)
    x_r, x_i = spec.real.copy(), spec.imag.copy()
    large_w = np.fft.ifftshift(get_window(window, npts))
    x_r *= large_w[: npts // 2 + 1]
    x_i *= large_w[: npts // 2 + 1]
    num = int(npts / factor)
    df = 1.0 / (npts * (1.0 / rate))
    d_large_f = 1.0 / num * new_rate
    f = df * np.arange(0, npts // 2 + 1, dtype=np.int32)
    large_f = d_large_f * np.arange(0, num // 2 + 1, dtype=np.int32)
    y = np.interp(large_f, f, x_r) + 1j * np.interp(large_f, f, x_i)
    return np.fft.irfft(y, n=num) * (float(num) / float(npts))  # irfft ignores Im y0 and (num even) Im y_nyq, as fftpack's layout does


# --------------------------------------------------------------------------- whole path on arrays

LABELS = {"eqtransformer": ["Detection", "P", "S"], "phasenet": ["P", "S", "N"]}
IN_SAMPLES = {"eqtransformer": 6000, "phasenet": 3001}
DEFAULT_OVERLAP = {"eqtransformer": 1800, "phasenet": 1500}
DEFAULT_BLINDING = {"eqtransformer": (500, 500), "phasenet": (0, 0)}


def forward_batches(kind: str, sd, windows: np.ndarray, batch_size: int = 256) -> np.ndarray:
    """Run the oracle network over (B, 3, L) pre-normalised windows -> (B, L, 3) heads-last fp32."""
    import torch

    from . import nets

    outs = []
    for i in range(0, len(windows), batch_size):
        xb = torch.from_numpy(np.ascontiguousarray(windows[i : i + batch_size]))
        if kind == "eqtransformer":
            det, p, s = nets.eqtransformer_forward(sd, xb)
            y = torch.stack((det, p, s), dim=-1)
        else:
            y = nets.phasenet_forward(sd, xb).permute(0, 2, 1)
        outs.append(y.numpy())
    return np.concatenate(outs, axis=0).astype(np.float32)


def annotate_array(
    kind: str,
    sd,
    trace: np.ndarray,
    overlap: Optional[int] = None,
    blinding: Optional[Tuple[int, int]] = None,
    stacking: str = "avg",
    batch_size: int = 256,
    peak_scope: str = "channel",
    return_windows: bool = False,
    detrend: bool = False,
):
    """(3, N) gap-free segment -> (pred_len, 3) stacked annotation (NaN where blinded / uncovered)."""
    L = IN_SAMPLES[kind]
    overlap = DEFAULT_OVERLAP[kind] if overlap is None else overlap
    blinding = DEFAULT_BLINDING[kind] if blinding is None else blinding
    starts = window_starts(trace.shape[1], L, overlap)
    if len(starts) == 0:
        return (np.empty((0, 3), np.float32), starts, None) if return_windows else np.empty((0, 3), np.float32)
    win = prenorm(cut_windows(trace, starts, L), kind, "peak", peak_scope, detrend)
    y = forward_batches(kind, sd, win, batch_size)
    yb = blind(y, blinding)
    out = reassemble(yb, starts, L, overlap, stacking)
    if return_windows:
        return out, starts, y
    return out


def classify_array(kind: str, annotation: np.ndarray, thresholds: Dict[str, float]):
    """Stacked (pred_len, 3) annotation -> {label: [(s0, s1, s_peak, value)]} with indices relative
    to the un-trimmed annotation (i.e. to the segment start), plus the per-label trim offsets."""
    res = {}
    offsets = {}
    for i, label in enumerate(LABELS[kind]):
        col, f, _ = trim_nan(annotation[:, i])
        offsets[label] = f
        if label == "N":
            continue
        key = "detection_threshold" if label == "Detection" else f"{label}_threshold"
        res[label] = [(s0 + f, s1 + f, sp + f, v) for (s0, s1, sp, v) in picks_from_trace(col, thresholds[key])]
    return res, offsets
