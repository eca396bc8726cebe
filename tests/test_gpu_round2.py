"""GPU parity tests added in round 2: the holes VERDICT r01 listed.

* the window normalisation under BOTH amplitude scopes (and SeisBench's norm_detrend) through the fused product path
  (vp_slice_forward and the whole annotate), not only the stand-alone K1;
* BASELINE.json's full-size configurations compared with the oracle itself (EQTransformer and PhaseNet station-day);
* the bf16 mode's +-1-sample pick-match rate as an assertion;
* stream ingest per segment against the oracle (gaps, missing components, int32 counts, 200 Hz traces);
* the boundary's promises: clean VP_ERR_WORKSPACE, pick-capacity retry, two host threads on one handle, handles on two
  devices in one process.
"""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

import volpick_b200 as vb
from oracle import pipeline
from volpick_b200 import _lib, weights_io
from volpick_b200.synthetic import station_start, synthetic_record, synthetic_stream

pytestmark = pytest.mark.gpu

PROB_ATOL = 1e-4  # north_star: fp32 probability traces agree within 1e-4 absolute


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.load()


@pytest.fixture(scope="module")
def eqt(lib):
    return vb.EQTransformer.from_pretrained("volpick").cuda()


@pytest.fixture(scope="module")
def pn(lib):
    return vb.PhaseNet.from_pretrained("volpick").cuda()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _match_rate(ref_trig, got_trig, tol=1):
    """Fraction of reference triggers with a same-label trigger whose peak lies within +-tol samples."""
    if len(ref_trig) == 0:
        return 1.0
    hit = 0
    for t in ref_trig:
        cand = got_trig[got_trig["label"] == t["label"]]
        hit += int(len(cand) > 0 and np.abs(cand["s_peak"] - t["s_peak"]).min() <= tol)
    return hit / len(ref_trig)


# ------------------------------------------------------------------------------------------ normalisation scopes
@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
@pytest.mark.parametrize("scope", ["channel", "window"])
@pytest.mark.parametrize("detrend", [False, True])
def test_slice_forward_scopes(lib, eqt, pn, sd_eqt, sd_pn, kind, scope, detrend):
    """The fused slicer + first conv (the product path of the tensor-core modes) under both peak scopes and with the
    linear detrend, against the oracle's annotate_batch_pre + forward on the same record."""
    model, sd = (eqt, sd_eqt) if kind == "eqtransformer" else (pn, sd_pn)
    L = model.in_samples
    x = synthetic_record(45, 40_000)
    x[1] *= 0.2  # unequal component amplitudes: the two scopes differ visibly
    # a drift of about one noise sigma per window (with a drift that dwarfs the signal the demeaned window is a ramp, far
    # from anything the network was trained on, and the f16x3 forward sits right at 1e-4 of the fp32 oracle: measured 1.04e-4)
    x += (np.linspace(-300.0, 500.0, x.shape[1], dtype=np.float32) * np.array([[1.0], [-0.1], [2.0]], dtype=np.float32))
    starts = pipeline.window_starts(x.shape[1], L, L - 900)
    nw = len(starts)
    d_tr, d_st = torch.from_numpy(x).cuda(), torch.from_numpy(starts).cuda()
    prec = _lib.PRECISION["f16x3"]
    need = _lib.check(lib.vp_forward_workspace_bytes(model._handle, nw, prec))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    y = torch.empty((nw, 3, L), dtype=torch.float32, device="cuda")
    flags = (_lib.PRE_TAPER if kind == "eqtransformer" else 0) | (_lib.PRE_DETREND if detrend else 0)
    _lib.check(lib.vp_slice_forward(model._handle, d_tr.data_ptr(), 0, x.shape[1], x.shape[1], d_st.data_ptr(), nw,
                                    _lib.PEAK_SCOPE[scope], flags, y.data_ptr(), ws.data_ptr(), need, prec, 0, L, _stream()))
    xw = torch.empty((nw, 3, L), dtype=torch.float32, device="cuda")
    _lib.check(lib.vp_slice_normalize(d_tr.data_ptr(), 0, x.shape[1], x.shape[1], d_st.data_ptr(), nw, L, _lib.PEAK_SCOPE[scope],
                                      flags, xw.data_ptr(), _stream()))
    torch.cuda.synchronize()
    win = pipeline.prenorm(pipeline.cut_windows(x, starts, L), kind, "peak", scope, detrend)
    np.testing.assert_allclose(xw.cpu().numpy(), win, atol=(2e-5 if detrend else 3e-6), rtol=0)
    ref = pipeline.forward_batches(kind, sd, win, 64).transpose(0, 2, 1)
    err = float(np.abs(y.cpu().numpy() - ref).max())
    print(f"{kind} scope={scope} detrend={detrend}: max|prob - oracle| = {err:.2e}")
    assert err <= PROB_ATOL


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_annotate_norm_kwargs_vs_oracle(lib, sd_eqt, sd_pn, kind):
    """SeisBench's constructor kwargs (norm_amp_per_comp, norm_detrend) and the per-window reading of norm="peak" through
    the whole annotate; the exposure between the two scopes is printed (DESIGN.md section 0c quotes it)."""
    cls, sd = (vb.EQTransformer, sd_eqt) if kind == "eqtransformer" else (vb.PhaseNet, sd_pn)
    x = synthetic_record(46, 45_000)
    x[1] *= 0.2
    anns = {}
    for name, kw, okw in [("default", {}, dict(peak_scope="channel")),
                          ("per_comp", dict(norm_amp_per_comp=True, peak_scope="window"), dict(peak_scope="channel")),
                          ("window", dict(peak_scope="window"), dict(peak_scope="window")),
                          ("detrend", dict(norm_detrend=True), dict(peak_scope="channel", detrend=True))]:
        model = cls(**kw)
        model.load_state_dict(weights_io.load_weights(weights_io.find_weights(kind, "volpick")[1]))
        model.cuda()
        argdict = model._argdict({})
        ann, _, _ = model.annotate_array(x, argdict, True, [0.0, 0.0, 0.0])
        ref = pipeline.annotate_array(kind, sd, x, argdict["overlap"], argdict["blinding"], "avg", **okw)
        ok = ~np.isnan(ref)
        np.testing.assert_array_equal(np.isnan(ann.T), np.isnan(ref))
        err = float(np.abs(ann.T[ok] - ref[ok]).max())
        print(f"{kind} {name}: max|prob - oracle| = {err:.2e}")
        assert err <= PROB_ATOL
        anns[name] = ann
    np.testing.assert_array_equal(anns["default"], anns["per_comp"])
    ok = ~np.isnan(anns["default"])
    print(f"{kind}: exposure max|prob(channel scope) - prob(window scope)| = "
          f"{float(np.abs(anns['default'][ok] - anns['window'][ok]).max()):.3f}")


# ------------------------------------------------------------------------------------------ full-size configurations
def _oracle_annotate_streaming(kind, sd, x, overlap, blinding, oracle_c, batch=512):
    """oracle.pipeline.annotate_array without its O(n * coverage) NaN buffer: windows are cut / normalised / run in
    batches and stacked by the C restatement (bit-identical to np.nanmean, tests/test_oracle.py)."""
    L = pipeline.IN_SAMPLES[kind]
    n = x.shape[1]
    starts = pipeline.window_starts(n, L, overlap)
    y = np.empty((len(starts), 3, L), dtype=np.float32)
    for i in range(0, len(starts), batch):
        win = pipeline.prenorm(pipeline.cut_windows(x, starts[i:i + batch], L), kind)
        y[i:i + batch] = pipeline.forward_batches(kind, sd, win, 256).transpose(0, 2, 1)
    out = np.empty((3, n), dtype=np.float32)
    rc = oracle_c.vpo_stack(y.ctypes.data, starts.ctypes.data, len(starts), L, 3, pipeline.coverage(L, overlap), blinding[0],
                            blinding[1], 0, out.ctypes.data, n)
    assert rc == 0
    return out


@pytest.mark.parametrize("kind", ["phasenet", "eqtransformer"])
def test_station_day_matches_oracle(eqt, pn, sd_eqt, sd_pn, oracle_c, kind):
    """BASELINE.json configs[1] (EQTransformer station-day, 17,269 windows) and a PhaseNet station-day (5,756 windows)
    against the oracle itself: probabilities <= 1e-4, pick indices bit-exact on the device's own trace, and the picks of
    the two traces paired one to one within one sample."""
    n = 8_640_000
    x = synthetic_record(1000, n)
    if kind == "eqtransformer":
        model, sd, kw = eqt, sd_eqt, dict(overlap=5500, blinding=(500, 500), stacking="avg", P_threshold=0.2, S_threshold=0.2)
    else:
        model, sd, kw = pn, sd_pn, {}
    argdict = model._argdict(kw)
    thr = model._thresholds(argdict)
    ann, trig, trim = model.annotate_array(x, argdict, True, thr)
    ref = _oracle_annotate_streaming(kind, sd, x, argdict["overlap"], argdict["blinding"], oracle_c)
    np.testing.assert_array_equal(np.isnan(ann), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = float(np.abs(ann[ok] - ref[ok]).max())
    print(f"{kind} station-day: max|prob - oracle| = {err:.3e} over {int(ok.sum())} samples, {len(trig)} triggers")
    assert err <= PROB_ATOL
    for li, label in enumerate(model.labels):
        if thr[li] <= 0:
            continue
        mine = trig[trig["label"] == li]
        f, l = int(trim[li, 0]), int(trim[li, 1])
        own = pipeline.picks_from_trace(ann[li, f:l + 1], np.float32(thr[li]))  # the pick rule on the device's own trace
        assert [(int(g["s0"]), int(g["s1"]), int(g["s_peak"])) for g in mine] == [(a + f, b + f, c + f) for a, b, c, _ in own]
        theirs = pipeline.picks_from_trace(ref[li, f:l + 1], np.float32(thr[li]))
        # a probability that differs by 1e-5 can move a threshold crossing; the peaks must still pair up
        assert abs(len(theirs) - len(mine)) <= max(2, len(theirs) // 200) and len(theirs) > 50
        peaks = np.array([c + f for _, _, c, _ in theirs])
        d = np.array([np.abs(peaks - int(g["s_peak"])).min() for g in mine])
        assert float((d <= 1).mean()) >= 0.995, f"{label}: {(d > 1).sum()} of {len(d)} peaks off by more than one sample"


def test_bf16_pick_match_rate(eqt, pn):
    """bf16 mode, reported separately (north_star) with ITS tolerance: probabilities within 5e-2 of the exact mode, >= 85 % of
    the exact mode's picks reproduced within +-1 sample and >= 97 % within +-10 samples (0.1 s) on a station-hour.  Measured:
    EQTransformer 0.92 / 1.00, PhaseNet see the log; the bench line of --precision bf16 carries the station-day rate."""
    x = synthetic_record(1001, 360_000)
    for model, kw in ((eqt, dict(overlap=5500, blinding=(500, 500), P_threshold=0.2, S_threshold=0.2)), (pn, {})):
        res = {}
        for prec in ("f16x3", "bf16"):
            a = model._argdict(dict(precision=prec, **kw))
            res[prec] = model.annotate_array(x, a, True, model._thresholds(a))
        ok = ~np.isnan(res["f16x3"][0])
        err = float(np.abs(res["bf16"][0][ok] - res["f16x3"][0][ok]).max())
        rate = _match_rate(res["f16x3"][1], res["bf16"][1])
        rate10 = _match_rate(res["f16x3"][1], res["bf16"][1], tol=10)
        print(f"{model.name} bf16: max|prob - exact| = {err:.3e}; {len(res['bf16'][1])} vs {len(res['f16x3'][1])} triggers, "
              f"match rate within 1 sample {rate:.4f}, within 10 samples {rate10:.4f}")
        assert err <= 5e-2 and len(res["f16x3"][1]) >= 20
        assert rate >= 0.85 and rate10 >= 0.97


# ------------------------------------------------------------------------------------------ stream ingest per segment
def _oracle_picks_for_record(kind, sd, arr, thresholds, overlap=None, blinding=None):
    ann = pipeline.annotate_array(kind, sd, arr.astype(np.float32), overlap, blinding)
    if ann.shape[0] == 0:
        return {}, ann
    return pipeline.classify_array(kind, ann, thresholds)[0], ann


def test_multi_segment_stream_vs_oracle(pn, sd_pn):
    """classify / annotate on a stream with a gap, a component that starts late and a missing component, compared per
    segment with the oracle run on the (3, n) records SeisBench's stream_to_array would build (zero-filled)."""
    x = synthetic_record(7, 60_000)
    t0 = station_start(7)
    hdr = dict(network="XX", station="SEG", location="", sampling_rate=100.0)
    st = vb.Stream([
        vb.Trace(x[0, :26_000], dict(hdr, channel="HHZ", starttime=t0)),
        vb.Trace(x[1, 500:26_000], dict(hdr, channel="HHN", starttime=t0 + 5.0)),      # starts 5 s late: zero-filled
        vb.Trace(x[2, :26_000], dict(hdr, channel="HHE", starttime=t0)),
        vb.Trace(x[0, 30_000:], dict(hdr, channel="HHZ", starttime=t0 + 300.0)),       # second segment: E is missing
        vb.Trace(x[1, 30_000:], dict(hdr, channel="HHN", starttime=t0 + 300.0)),
    ])
    rec1 = x[:, :26_000].copy()
    rec1[1, :500] = 0
    rec2 = x[:, 30_000:].copy()
    rec2[2] = 0
    thr = {"P_threshold": 0.39, "S_threshold": 0.34}
    ann = pn.annotate(st)
    picks = pn.classify(st).picks
    assert len(ann) == 6
    n_ref = 0
    for rec, start in ((rec1, t0), (rec2, t0 + 300.0)):
        ref_picks, ref_ann = _oracle_picks_for_record("phasenet", sd_pn, rec, thr)
        for li, label in enumerate("PSN"):
            tr = [t for t in ann if t.stats.channel == f"PhaseNet_{label}" and t.stats.starttime == start]
            assert len(tr) == 1 and tr[0].stats.npts == rec.shape[1]
            assert float(np.abs(tr[0].data - ref_ann[:, li]).max()) <= PROB_ATOL
        for phase in "PS":
            mine = [p for p in picks if p.phase == phase and start <= p.peak_time < start + rec.shape[1] / 100.0]
            assert len(mine) == len(ref_picks[phase])
            n_ref += len(mine)
            for p, r in zip(mine, ref_picks[phase]):
                assert abs((p.peak_time - start) * 100 - r[2]) <= 1 + 1e-6
    assert n_ref >= 4


def test_records_assembled_by_the_helper_thread(pn, monkeypatch):
    """Long streams have their records assembled one ahead by a helper thread into a ring of three pinned buffers (models._run);
    forced here on a short multi-station stream with a gap and a too-short fragment: same picks, same annotation traces and the
    same warnings as the inline path; an error in the middle of the stream leaves no thread and no record behind."""
    import threading

    traces = []
    for j in range(5):
        x = synthetic_record(20 + j, 30_000)
        t0 = station_start(20 + j)
        hdr = dict(network="XX", station=f"T{j}", location="", sampling_rate=100.0)
        for i, c in enumerate("ZNE"):
            traces.append(vb.Trace(x[i, :14_000], dict(hdr, channel="HH" + c, starttime=t0)))
            traces.append(vb.Trace(x[i, 16_000:], dict(hdr, channel="HH" + c, starttime=t0 + 160.0)))
        traces.append(vb.Trace(x[0, :1000], dict(hdr, channel="HHZ", starttime=t0 + 1000.0)))  # shorter than a window
    st = vb.Stream(traces)
    inline_picks, inline_ann = pn.classify(st).picks, pn.annotate(st)
    assert len(inline_picks) >= 5
    pn._prefetch_min_samples = 0
    try:
        for _ in range(3):
            picks, ann = pn.classify(st).picks, pn.annotate(st)
            assert picks == inline_picks and len(ann) == len(inline_ann) == 30
            for u, v in zip(ann, inline_ann):
                assert u.stats.channel == v.stats.channel and u.stats.starttime == v.stats.starttime
                np.testing.assert_array_equal(u.data, v.data)
        from volpick_b200 import models

        calls, real = [0], models._copy_jobs

        def failing(jobs):  # the fourth record fails while the helper thread assembles it (two records are in flight then)
            calls[0] += 1
            if calls[0] == 4:
                raise OSError("simulated read error")
            return real(jobs)

        monkeypatch.setattr(models, "_copy_jobs", failing)
        with pytest.raises(OSError, match="simulated read error"):
            pn.classify(st)
        monkeypatch.setattr(models, "_copy_jobs", real)
        assert pn.classify(st).picks == inline_picks  # the ring and the slots are usable after the failure
    finally:
        del pn._prefetch_min_samples
    import time
    time.sleep(0.3)
    assert not [t for t in threading.enumerate() if t.name == "vp-assemble" and t.is_alive()]


def test_int32_stream_equals_float_stream(eqt):
    """Traces of int32 counts stay int32 on the wire (converted by the slicer on the device): same picks and
    probabilities as the float32 stream of the same counts."""
    x = np.round(synthetic_record(8, 40_000)).astype(np.int32)
    t0 = station_start(8)
    hdr = dict(network="XX", station="I32", location="", sampling_rate=100.0)
    st_i = vb.Stream([vb.Trace(x[i], dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("ZNE")])
    st_f = vb.Stream([vb.Trace(x[i].astype(np.float32), dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("ZNE")])
    recs = eqt.stream_to_arrays(list(st_i), eqt._argdict({}))
    assert recs[0][1].dtype == np.int32
    a_i, a_f = eqt.annotate(st_i), eqt.annotate(st_f)
    for u, v in zip(a_i, a_f):
        np.testing.assert_array_equal(u.data, v.data)
    assert eqt.classify(st_i).picks == eqt.classify(st_f).picks


def test_resample_200hz_on_device(pn, sd_pn):
    """A 200 Hz stream: zero-phase 4-corner low-pass at 50 Hz + every second sample (SeisBench WaveformModel.resample
    for integer ratios) on the device, against scipy on the host; then the picks against the oracle."""
    from scipy.signal import iirfilter, sosfilt, zpk2sos

    rng = np.random.default_rng(5)
    x100 = synthetic_record(9, 30_000)
    x200 = np.repeat(x100, 2, axis=1) + rng.standard_normal((3, 60_000)).astype(np.float32)
    t0 = station_start(9)
    hdr = dict(network="XX", station="R200", location="", sampling_rate=200.0)
    st = vb.Stream([vb.Trace(x200[i], dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("ZNE")])
    z, p, k = iirfilter(4, 0.5, btype="lowpass", ftype="butter", output="zpk")
    sos = zpk2sos(z, p, k)
    y = sosfilt(sos, x200.astype(np.float64), axis=-1)
    y = sosfilt(sos, y[:, ::-1], axis=-1)[:, ::-1]
    want = np.ascontiguousarray(y[:, ::2]).astype(np.float32)
    st2 = st.copy()
    for tr in st2:
        pn.resample_trace(tr)
        assert tr.stats.sampling_rate == 100.0 and len(tr.data) == 30_000
    got = np.stack([tr.data for tr in st2])
    assert float(np.abs(got - want).max()) <= 1e-3 * float(np.abs(want).max())
    picks = pn.classify(st).picks
    ref, _ = _oracle_picks_for_record("phasenet", sd_pn, want, {"P_threshold": 0.39, "S_threshold": 0.34})
    for phase in "PS":
        mine = [q for q in picks if q.phase == phase]
        assert len(mine) == len(ref[phase]) and len(mine) > 0
        for q, r in zip(mine, ref[phase]):
            assert abs((q.peak_time - t0) * 100 - r[2]) <= 1 + 1e-6


def test_resample_fractional_ratio_on_device(pn, sd_pn):
    """A 125 Hz stream (not an integer multiple of 100 Hz): ObsPy's FFT resampling restated on the device (float64, torch.fft)
    against the oracle's NumPy restatement, then classify() on the stream against the oracle on the resampled record."""
    rng = np.random.default_rng(6)
    x100 = synthetic_record(9, 30_000)  # the oracle finds 2 P and 2 S picks on its resampled version (checked on the CPU)
    # a 125 Hz version of the same ground motion: band-limited interpolation of the 100 Hz record + a little noise
    x125 = np.stack([pipeline.fft_resample(x100[i], 100.0, 125.0) for i in range(3)]).astype(np.float32)
    x125 += 0.01 * float(np.abs(x125).max()) * rng.standard_normal(x125.shape).astype(np.float32)
    want = np.stack([pipeline.fft_resample(x125[i], 125.0, 100.0) for i in range(3)])
    t0 = station_start(9)
    hdr = dict(network="XX", station="R125", location="", sampling_rate=125.0)
    st = vb.Stream([vb.Trace(x125[i], dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("ZNE")])
    st2 = st.copy()
    for tr in st2:
        pn.resample_trace(tr)
        assert tr.stats.sampling_rate == 100.0 and len(tr.data) == want.shape[1]
    got = np.stack([tr.data for tr in st2])
    assert got.dtype == np.float64
    assert float(np.abs(got - want).max()) <= 1e-9 * float(np.abs(want).max())
    for odd in (x125[0, :-1], x125[0, :4097]):  # odd lengths: no Nyquist bin
        g = pn._fft_resample(odd, 125.0)
        w = pipeline.fft_resample(odd, 125.0, 100.0)
        assert g.shape == w.shape and float(np.abs(g - w).max()) <= 1e-9 * float(np.abs(w).max())
    picks = pn.classify(st).picks
    ref, _ = _oracle_picks_for_record("phasenet", sd_pn, want.astype(np.float32), {"P_threshold": 0.39, "S_threshold": 0.34})
    for phase in "PS":
        mine = [q for q in picks if q.phase == phase]
        assert len(mine) == len(ref[phase]) and len(mine) > 0
        for q, r in zip(mine, ref[phase]):
            assert abs((q.peak_time - t0) * 100 - r[2]) <= 1 + 1e-6


# ------------------------------------------------------------------------------------------ boundary promises
@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
@pytest.mark.parametrize("precision", ["f16x3", "fp32"])
def test_undersized_workspace_is_a_clean_error(lib, eqt, pn, kind, precision):
    """VP_ERR_WORKSPACE before any launch (the header documents it as a clean error): a guard region behind the short
    workspace stays untouched and the next correct call still works."""
    model = eqt if kind == "eqtransformer" else pn
    L, B = model.in_samples, 16
    prec = _lib.PRECISION[precision]
    need = _lib.check(lib.vp_forward_workspace_bytes(model._handle, B, prec))
    x = torch.zeros((B, 3, L), dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    buf = torch.full((need,), 0x5A, dtype=torch.uint8, device="cuda")
    short = need // 3
    with pytest.raises(_lib.VolpickError, match="workspace too small"):
        _lib.check(lib.vp_forward(model._handle, x.data_ptr(), B, y.data_ptr(), buf.data_ptr(), short, prec, _stream()))
    torch.cuda.synchronize()
    assert bool((buf[short:] == 0x5A).all()), "a refused call wrote behind its workspace"
    _lib.check(lib.vp_forward(model._handle, x.data_ptr(), B, y.data_ptr(), buf.data_ptr(), need, prec, _stream()))
    torch.cuda.synchronize()


def test_pick_capacity_retry(lib, pn):
    """The Python API re-runs a record whose triggers exceed the pick buffer with the reported count; the C ABI reports
    the overflow as VP_ERR_CAPACITY (never a truncation)."""
    x = synthetic_record(4, 30_000)
    thr = [1e-6, 1e-6, 0.0]
    _, big, _ = pn.annotate_array(x, None, False, thr)
    _, small, _ = pn.annotate_array(x, None, False, thr, pick_capacity=1)
    assert len(big) > 1
    np.testing.assert_array_equal(big, small)
    argdict = pn._argdict({})
    params = pn._params(argdict, thr)
    need = _lib.check(lib.vp_annotate_workspace_bytes(pn._handle, x.shape[1], C.byref(params), 1, 1))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    trig = (_lib.Trigger * 1)()
    n_picks = C.c_int64(0)
    trim = np.zeros(6, dtype=np.int64)
    rc = lib.vp_annotate(pn._handle, x.ctypes.data, 1, 0, x.shape[1], x.shape[1], C.byref(params), None, 1, C.cast(trig, C.c_void_p), 1,
                         C.byref(n_picks), trim.ctypes.data, ws.data_ptr(), need, _stream())
    assert rc == _lib.VP_ERR_CAPACITY and n_picks.value == len(big)
    assert b"exceed the pick capacity" in lib.vp_last_error()


def test_two_host_threads_share_one_handle(eqt):
    """A handle is immutable after creation: two host threads, each with its own stream and workspace, annotate
    different records at the same time and get what a single thread gets."""
    recs = [synthetic_record(50 + i, 50_000) for i in range(2)]
    argdict = eqt._argdict(dict(overlap=5500, blinding=(500, 500)))
    want = [eqt.annotate_array(r, argdict, True, [0.3, 0.2, 0.2]) for r in recs]
    got = [None, None]
    errs = []

    def work(i):
        try:
            torch.cuda.set_device(eqt._device_index)
            stream = torch.cuda.Stream()
            ws = torch.empty(eqt.annotate_workspace_bytes(recs[i].shape[1], argdict, True), dtype=torch.uint8, device="cuda")
            for _ in range(3):
                got[i] = eqt.annotate_array_async(recs[i], argdict, True, [0.3, 0.2, 0.2], stream=stream, workspace=ws).result()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g[0], w[0])
        np.testing.assert_array_equal(g[1], w[1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_handles_on_two_devices_in_one_process():
    """The opt-in of dynamic shared memory is per device (not per process): a second device runs every kernel class."""
    x = synthetic_record(60, 40_000)
    res = []
    prev = torch.cuda.current_device()
    for dev in (0, 1):
        for cls in (vb.EQTransformer, vb.PhaseNet):
            m = cls.from_pretrained("volpick").to(f"cuda:{dev}")
            assert torch.cuda.current_device() == prev  # creating a handle leaves the caller's device alone
            a = m._argdict({})
            res.append(m.annotate_array(x, a, True, m._thresholds(a)))
    for (a0, t0, _), (a1, t1, _) in zip(res[:2], res[2:]):
        np.testing.assert_array_equal(a0, a1)
        np.testing.assert_array_equal(t0, t1)


@pytest.mark.parametrize("B", [1, 7, 64])
def test_attention_kernels_agree(eqt, B, monkeypatch):
    """seq.cu: attention2_kernel (two lanes per query time step, halves joined by shuffles; the default) against
    attention_kernel (one thread per query, VP_ATTN_V1=1): the transformer blocks (full attention + LayerNorm + feed-forward)
    and the banded pick attentions differ only in summation order -- bottleneck tap and probabilities within fp32 rounding."""
    rng = np.random.default_rng(40 + B)
    x = rng.standard_normal((B, 3, 6000)).astype(np.float32)
    x[:, :, 3100:3500] *= 8.0
    xd = torch.from_numpy(x).cuda()
    monkeypatch.setenv("VP_ATTN_V1", "0")
    new = torch.stack(eqt.forward(xd, precision="f16x3"), dim=1).cpu().numpy()
    monkeypatch.setenv("VP_ATTN_V1", "1")
    old = torch.stack(eqt.forward(xd, precision="f16x3"), dim=1).cpu().numpy()
    assert np.isfinite(new).all() and new.shape == old.shape == (B, 3, 6000)
    d = float(np.abs(new - old).max())
    print(f"B={B}: attention v2 vs v1 max|dprob| = {d:.3e}")
    assert d <= 2e-6


# ------------------------------------------------------------------------------------------ decoder tail: TMEM-operand kernel
@pytest.mark.parametrize("precision,atol", [("f16x3", 2e-5), ("bf16", 5e-2)])
@pytest.mark.parametrize("B", [1, 3, 50])
def test_decoder_tail_generations_agree(eqt, precision, atol, B, monkeypatch):
    """fused_dec2.cu (1500- / 3000-sample levels in tensor memory, the default) against fused_dec.cu (every operand from shared
    memory, VP_DECB_V1=1) on the same windows: two independent index maps (lane folding, halo shuffles, banded decoder.convs.4
    weights vs. row-tap tiles) must give the same probabilities; both are compared with the oracle elsewhere."""
    rng = np.random.default_rng(5 + B)
    x = rng.standard_normal((B, 3, 6000)).astype(np.float32)
    x[:, :, 2000:2300] *= 6.0
    xd = torch.from_numpy(x).cuda()
    monkeypatch.setenv("VP_DECB_V1", "0")
    new = torch.stack(eqt.forward(xd, precision=precision), dim=1).cpu().numpy()
    monkeypatch.setenv("VP_DECB_V1", "1")
    old = torch.stack(eqt.forward(xd, precision=precision), dim=1).cpu().numpy()
    assert new.shape == old.shape == (B, 3, 6000)
    assert np.isfinite(new).all()
    d = float(np.abs(new - old).max())
    print(f"{precision} B={B}: decoder tail v2 vs v1 max|diff| = {d:.3e}")
    assert d <= atol


@pytest.mark.parametrize("precision,atol", [("f16x3", 2e-5), ("bf16", 5e-2)])
@pytest.mark.parametrize("B", [1, 3, 50])
def test_fused_encoder_front_agrees_with_layers(eqt, precision, atol, B, monkeypatch):
    """fused_enc.cu (encoder.convs.1-3 in one kernel, the 1500- / 750-sample levels in tensor memory) against the same three layers
    run one by one through tcconv.cu (VP_ENC_FUSED=0): encoder output (enc6 tap) and probabilities."""
    rng = np.random.default_rng(11 + B)
    x = rng.standard_normal((B, 3, 6000)).astype(np.float32)
    x[:, :, :40] *= 5.0   # energy at both window ends: the convs' zero padding and the item edges
    x[:, :, -40:] *= 5.0
    xd = torch.from_numpy(x).cuda()
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("VP_ENC_FUSED", flag)
        out[flag] = (eqt.forward_tap(xd, "enc6", precision=precision).cpu().numpy(),
                     torch.stack(eqt.forward(xd, precision=precision), dim=1).cpu().numpy())
    scale = float(np.abs(out["0"][0]).max())
    d6 = float(np.abs(out["1"][0] - out["0"][0]).max())
    dp = float(np.abs(out["1"][1] - out["0"][1]).max())
    print(f"{precision} B={B}: fused encoder front vs layers: enc6 max|diff| = {d6:.3e} (scale {scale:.2f}), probabilities {dp:.3e}")
    assert np.isfinite(out["1"][1]).all()
    assert d6 <= (1e-5 if precision == "f16x3" else 5e-2) * max(scale, 1.0)
    assert dp <= atol
