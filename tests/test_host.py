"""CPU tests of the host side: containers, stream bookkeeping, argument handling, weight layout,
and that the C-ABI library loads and exports every symbol of include/volpick_b200.h."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import volpick_b200 as vb
from oracle import pipeline
from volpick_b200 import _lib, models, weights_io
from volpick_b200.stream import Stream, Trace, UTCDateTime
from volpick_b200.synthetic import synthetic_record, synthetic_stream, station_start


def test_header_symbols_are_exported(repo_root, built_lib):
    hdr = open(os.path.join(repo_root, "include", "volpick_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(vp_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(built_lib, n), n
    assert built_lib.vp_version() >= 100
    assert C.sizeof(_lib.Trigger) == 32 and C.sizeof(_lib.AnnotateParams) == 56


@pytest.mark.parametrize("n,L,ov", [(360_000, 6000, 5500), (8_640_000, 3001, 1500), (6001, 6000, 100), (5999, 6000, 0),
                                     (12345, 3001, 3000), (3001, 3001, 1500)])
def test_vp_window_starts_matches_oracle(built_lib, n, L, ov):
    ref = pipeline.window_starts(n, L, ov)
    assert built_lib.vp_window_count(n, L, ov) == len(ref)
    buf = np.zeros(max(len(ref), 1), dtype=np.int64)
    cnt = C.c_int64(-1)
    assert built_lib.vp_window_starts(n, L, ov, buf.ctypes.data, len(buf), C.byref(cnt)) == 0
    assert cnt.value == len(ref)
    np.testing.assert_array_equal(buf[: cnt.value], ref)
    assert built_lib.vp_coverage(L, ov) == pipeline.coverage(L, ov)


def test_vp_errors_are_loud(built_lib):
    cnt = C.c_int64(0)
    assert built_lib.vp_window_starts(100, 10, 10, None, 0, C.byref(cnt)) == _lib.VP_ERR_ARG
    assert b"overlap" in built_lib.vp_last_error()
    buf = np.zeros(1, dtype=np.int64)
    assert built_lib.vp_window_starts(100, 10, 5, buf.ctypes.data, 1, C.byref(cnt)) == _lib.VP_ERR_CAPACITY
    with pytest.raises(_lib.VolpickError):
        _lib.check(_lib.VP_ERR_CAPACITY)
    assert built_lib.vp_model_expected_floats(0) == 378_823 and built_lib.vp_model_expected_floats(1) == 269_675


def test_utcdatetime():
    t = UTCDateTime("2005-05-31T21:04:52.110000Z")
    assert str(t + 18.86) == "2005-05-31T21:05:10.970000Z"
    assert (t + 1.5) - t == 1.5 and t < t + 0.01 and t == UTCDateTime(t)
    assert UTCDateTime("2020-01-01") + 86400.0 == UTCDateTime("2020-01-02T00:00:00")
    assert str(station_start(3)) == "2020-01-04T00:00:00.000000Z"
    # 100 Hz sample arithmetic is exact in integer nanoseconds
    assert (t + 8_639_999 / 100.0).ns - t.ns == 8_639_999 * 10_000_000


def test_stream_merge_and_segments():
    m = vb.PhaseNet.from_pretrained("volpick")
    x = synthetic_record(1, 9000)
    t0 = UTCDateTime("2021-03-04T05:06:07.5")
    hdr = dict(network="XX", station="A", location="00", sampling_rate=100.0)
    st = Stream([
        Trace(x[0, :4000], dict(hdr, channel="EHZ", starttime=t0)),
        Trace(x[0, 4000:], dict(hdr, channel="EHZ", starttime=t0 + 40.0)),       # contiguous -> merged
        Trace(x[1, :3000], dict(hdr, channel="EHN", starttime=t0)),              # N ends early, gap, resumes
        Trace(x[1, 5000:], dict(hdr, channel="EHN", starttime=t0 + 50.0)),
    ])
    st.merge(-1)
    assert len(st) == 3
    segs = m.stream_to_arrays(list(st), m._argdict({}))
    assert len(segs) == 1
    s0, arr = segs[0]
    assert s0 == t0 and arr.shape == (3, 9000) and arr.dtype == np.float32
    np.testing.assert_array_equal(arr[0], x[0])
    np.testing.assert_array_equal(arr[1, :3000], x[1, :3000])
    assert np.all(arr[1, 3000:5000] == 0) and np.all(arr[2] == 0)  # missing data / component zero-filled
    np.testing.assert_array_equal(arr[1, 5000:], x[1, 5000:])
    # a real gap on all components splits the record
    st2 = Stream([Trace(x[0, :3500], dict(hdr, channel="EHZ", starttime=t0)),
                  Trace(x[0, 4000:], dict(hdr, channel="EHZ", starttime=t0 + 40.0))])
    segs = m.stream_to_arrays(list(st2), m._argdict({}))
    assert [a.shape[1] for _, a in segs] == [3500, 5000] and segs[1][0] == t0 + 40.0
    # strict: only spans with all three components
    st3 = Stream([Trace(x[i, o:], dict(hdr, channel="EH" + c, starttime=t0 + o / 100.0)) for i, (c, o) in enumerate(zip("ZNE", (0, 100, 250)))])
    segs = m.stream_to_arrays(list(st3), m._argdict({"strict": True}))
    assert len(segs) == 1 and segs[0][1].shape[1] == 8750 and segs[0][0] == t0 + 2.5
    # Z12 instead of ZNE
    st4 = Stream([Trace(x[i], dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("Z12")])
    (_, arr4), = m.stream_to_arrays(list(st4), m._argdict({}))
    np.testing.assert_array_equal(arr4, x)


def test_stream_records_into_caller_buffers():
    """_iter_stream_arrays(alloc=...): records are assembled straight into the buffers the caller hands out (models._run: a ring of
    pinned buffers) -- same content as the fresh-array form, int32 counts stay int32, zero fill of what no trace covers even when the
    buffer holds an older record, one alloc call per non-strict record and none in strict mode."""
    m = vb.PhaseNet.from_pretrained("volpick")
    x = np.round(synthetic_record(3, 12_000)).astype(np.int32)
    t0 = UTCDateTime("2022-02-02T00:00:00")
    hdr = dict(network="XX", station="R", location="", sampling_rate=100.0)
    traces = [Trace(x[0, :5000], dict(hdr, channel="HHZ", starttime=t0)),
              Trace(x[1, 200:5000], dict(hdr, channel="HHN", starttime=t0 + 2.0)),    # late start: zero fill in front
              Trace(x[0, 6000:], dict(hdr, channel="HHZ", starttime=t0 + 60.0)),     # second segment: N and E missing
              Trace(x[2, 6000:11_000], dict(hdr, channel="HHE", starttime=t0 + 60.0))]
    ring = [np.full(3 * 12_000, 7, dtype=np.int32) for _ in range(3)]  # "older records" in the buffers
    calls = []

    def alloc(shape, dtype):
        k = len(calls) % 3
        calls.append((shape, np.dtype(dtype)))
        return ring[k][: shape[0] * shape[1]].view(dtype).reshape(shape)

    want = m.stream_to_arrays(traces, m._argdict({}))
    got = [(t, a.copy()) for t, a in m._iter_stream_arrays(traces, m._argdict({}), alloc)]
    assert len(got) == len(want) == 2 and len(calls) == 2 and all(d == np.int32 for _, d in calls)
    for (t_g, a_g), (t_w, a_w) in zip(got, want):
        assert t_g == t_w and a_g.dtype == a_w.dtype == np.int32
        np.testing.assert_array_equal(a_g, a_w)
    assert not got[0][1][1, :200].any() and not got[0][1][2].any() and not got[1][1][1].any() and not got[1][1][2, 5000:].any()
    np.testing.assert_array_equal(got[1][1][2, :5000], x[2, 6000:11_000])
    del calls[:]
    strict = list(m._iter_stream_arrays(traces, m._argdict({"strict": True}), alloc))
    assert not calls and strict == []  # no span holds all three components


def test_argdict_defaults_and_validation():
    e = vb.EQTransformer.from_pretrained("volpick")
    p = vb.PhaseNet.from_pretrained("volpick")
    a = e._argdict({})
    assert (a["overlap"], a["blinding"], a["stacking"], a["batch_size"]) == (1800, (500, 500), "avg", 256)
    assert e._thresholds(a) == [pytest.approx(0.10141666), 0.22, 0.22]  # JSON default_args
    a = e._argdict(dict(overlap=5500, blinding=[500, 500], P_threshold=0.2, S_threshold=0.2, parallelism=None, copy=True))
    assert e._thresholds(a)[1:] == [0.2, 0.2] and a["blinding"] == (500, 500)
    b = p._argdict({})
    assert (b["overlap"], b["blinding"]) == (1500, (0, 0)) and p._thresholds(b) == [0.39, 0.34, 0.0]
    assert e.labels == ["Detection", "P", "S"] and p.labels == ["P", "S", "N"]
    assert e.in_samples == 6000 and p.in_samples == 3001 and e.norm == "peak" and e.component_order == "ZNE"
    with pytest.raises(ValueError, match="Stacking"):
        e._argdict(dict(stacking="median"))
    with pytest.raises(ValueError):
        e._argdict(dict(overlap=6000))
    with pytest.warns(UserWarning, match="Unknown argument"):
        e._argdict(dict(bogus=1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        e.classify(synthetic_stream(0, 7000))
    assert "Zhong" in e.weights_docstring and str(e.device) == "cpu"
    assert vb.PhaseNet.list_pretrained() == ["volpick", "volpick_95train"]


def test_weight_layout_roundtrip(tmp_path, sd_eqt):
    w = weights_io.load_weights(weights_io.find_weights("eqtransformer", "volpick")[1])
    flat = models.flatten_weights(w, models.eqtransformer_spec())
    assert flat.size == 378_823 and flat.dtype == np.float32
    np.testing.assert_array_equal(flat[: 8 * 3 * 11], sd_eqt["encoder.convs.0.weight"].numpy().reshape(-1))
    np.testing.assert_array_equal(flat[-1:], sd_eqt["pick_convs.1.bias"].numpy())
    wp = weights_io.load_weights(weights_io.find_weights("phasenet", "volpick_95train")[1])
    assert models.flatten_weights(wp, models.phasenet_spec()).size == 269_675
    path = tmp_path / "x.vpw.v1"
    weights_io.write_vpw(str(path), w)
    w2 = weights_io.read_vpw(str(path))
    assert list(w2) == list(w) and all(np.array_equal(w[k], w2[k]) for k in w)
    bad = dict(w)
    bad["encoder.convs.0.weight"] = np.zeros((8, 3, 9), np.float32)
    with pytest.raises(ValueError, match="expected shape"):
        models.flatten_weights(bad, models.eqtransformer_spec())
    m = vb.EQTransformer()
    m.load_state_dict(sd_eqt)  # torch tensors are accepted like SeisBench's load_state_dict
    np.testing.assert_array_equal(m._flat, flat)


def test_reference_pt_files_convert_identically():
    ref = "/root/reference/Final_models/volpick/phasenet/volpick.pt.v1"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    a = weights_io.read_pt(ref)
    b = weights_io.load_weights(weights_io.find_weights("phasenet", "volpick")[1])
    assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a)


def test_pick_containers():
    t = UTCDateTime("2020-01-01T00:00:10")
    p1 = vb.Pick("XX.A.", t + 1, t + 3, t + 2, 0.9, "P")
    p2 = vb.Pick("XX.A.", t, t + 3, t + 1, 0.5, "S")
    pl = vb.PickList(sorted([p1, p2]))
    assert pl[0] is p2 and len(pl.select(phase="P")) == 1 and len(pl.select(min_confidence=0.6)) == 1
    df = pl.to_dataframe()
    assert list(df.columns) == ["trace_id", "start_time", "end_time", "peak_time", "peak_value", "phase"]  # README.md:69-80
    out = vb.ClassifyOutput("EQTransformer", picks=pl, detections=vb.DetectionList())
    assert out.picks is pl and "picks" in str(out)
    with pytest.raises(ValueError):
        vb.Pick("XX.A.", t + 2, t + 3, t + 1, 0.9, "P")


def test_pick_lists_come_out_in_sorted_order():
    """models._run builds the objects per record and orders them with one lexsort over (start ns, end ns, trace_id, phase):
    the same list as sorted() over the objects (Pick.__lt__ / Detection.__lt__), ties included."""
    rng = np.random.default_rng(5)
    base = 1_600_000_000_000_000_000
    parts, dparts, plain, dplain = [], [], [], []
    for r in range(6):
        for lab in ("P", "S"):
            s0 = np.sort(rng.integers(0, 2000, 300))  # many equal start times across records, stations and phases
            ns0, ns1, nsp = base + s0 * 10**7, base + (s0 + rng.integers(1, 4, 300)) * 10**7, base + s0 * 10**7 + 5
            vals = rng.random(300).astype(np.float32)
            tid = "XX.S%d." % (r % 3)
            parts.append((tid, lab, ns0, ns1, models.WaveformModel._build_objects(tid, lab, UTCDateTime, ns0, ns1, nsp, vals)))
            plain += [vb.Pick(tid, UTCDateTime(ns=a), UTCDateTime(ns=b), UTCDateTime(ns=c), v, lab)
                      for a, b, c, v in zip(ns0.tolist(), ns1.tolist(), nsp.tolist(), vals.tolist())]
        dparts.append((tid, "", ns0, ns1, models.WaveformModel._build_objects(tid, "", UTCDateTime, ns0, ns1, None, vals)))
        dplain += [vb.Detection(tid, UTCDateTime(ns=a), UTCDateTime(ns=b), v) for a, b, v in zip(ns0.tolist(), ns1.tolist(), vals.tolist())]
    got, want = models.WaveformModel._sorted_objects(parts), sorted(plain)
    assert len(got) == len(want) == 3600 and all(a == b for a, b in zip(got, want))
    got, want = models.WaveformModel._sorted_objects(dparts), sorted(dplain)
    assert len(got) == 1800 and all(a == b for a, b in zip(got, want))
    assert models.WaveformModel._sorted_objects([]) == []

    class OtherTime:  # a foreign time class (obspy.UTCDateTime) goes through its ns= keyword
        def __init__(self, ns=None):
            self.ns = ns

    objs = models.WaveformModel._build_objects("XX.A.", "P", OtherTime, ns0[:5], ns1[:5], nsp[:5], vals[:5])
    assert all(type(o.start_time) is OtherTime and o.start_time.ns == int(a) for o, a in zip(objs, ns0[:5]))
    with pytest.raises(ValueError):  # Pick.__init__'s ordering check survives the bulk constructor
        models.WaveformModel._build_objects("XX.A.", "P", UTCDateTime, ns0, ns1, ns1 + 1, vals)


def test_prefetch_helper_thread():
    """models._prefetched: same items in the same order, at most one item ahead of the consumer plus the one being produced, an
    exception re-raised at its place, and no helper left behind once the stop event is set."""
    import threading
    import time

    produced = []

    def gen(n, fail_at=None):
        for i in range(n):
            if i == fail_at:
                raise OSError("boom")
            produced.append(i)
            yield i

    stop = threading.Event()
    seen = []
    for x in models._prefetched(gen(20), stop):
        time.sleep(0.002)
        seen.append(x)
        assert len(produced) <= len(seen) + 2  # one in the queue, one being handed over
    assert seen == list(range(20))
    del produced[:]
    stop = threading.Event()
    got = []
    with pytest.raises(OSError, match="boom"):
        for x in models._prefetched(gen(10, fail_at=4), stop):
            got.append(x)
    assert got == [0, 1, 2, 3]
    # the consumer walks away in the middle: the helper must not wait for the queue for ever
    stop = threading.Event()
    it = models._prefetched(gen(1000), stop)
    assert next(it) == 0
    stop.set()
    t_end = time.time() + 2.0
    while time.time() < t_end and [t for t in threading.enumerate() if t.name == "vp-assemble" and t.is_alive()]:
        time.sleep(0.02)
    assert not [t for t in threading.enumerate() if t.name == "vp-assemble" and t.is_alive()]


def test_copy_jobs_in_pieces():
    """models._copy_jobs: long records are copied in 2 M-sample pieces on a thread pool; dtype conversion and offsets as plain
    slice assignment."""
    rng = np.random.default_rng(9)
    n = (1 << 22) + 12345
    src = [rng.integers(-1000, 1000, n).astype(np.int32), rng.standard_normal(n).astype(np.float32), rng.standard_normal(n - 777)]
    dst = np.zeros((3, n), dtype=np.float32)
    models._copy_jobs([(dst[0], src[0]), (dst[1], src[1]), (dst[2, 777:], src[2])])
    np.testing.assert_array_equal(dst[0], src[0].astype(np.float32))
    np.testing.assert_array_equal(dst[1], src[1])
    np.testing.assert_array_equal(dst[2, 777:], src[2].astype(np.float32))
    assert not dst[2, :777].any()
    small = np.zeros(10, np.float32)
    models._copy_jobs([(small, np.arange(10))])
    np.testing.assert_array_equal(small, np.arange(10, dtype=np.float32))


def test_filter_design_matches_oracle():
    """filter_args / filter_kwargs (model_training/test_onephase.ipynb cell 43): the host designs the same second-order
    sections as the oracle's ObsPy restatement; the records are filtered on the device (GPU test)."""
    from oracle import pipeline

    m = vb.EQTransformer.from_pretrained("volpick")
    assert m.design_filter() is None
    m.filter_args = ["highpass"]
    m.filter_kwargs = {"freq": 0.5, "corners": 2, "zerophase": True}
    sos, zerophase = m.design_filter()
    np.testing.assert_array_equal(sos, pipeline.design_sos("highpass", 100.0, corners=2, freq=0.5))
    assert zerophase and sos.shape == (1, 6)
    m.filter_args, m.filter_kwargs = ["bandpass"], {"freqmin": 1, "freqmax": 20}
    sos, zerophase = m.design_filter()
    assert sos.shape == (4, 6) and not zerophase
    m.filter_args = ["notch"]
    with pytest.raises(NotImplementedError):
        m.design_filter()
    # oracle sosfilt: one pass is causal, zero-phase is symmetric for a symmetric input
    x = np.zeros((1, 401), np.float32)
    x[0, 200] = 1.0
    y = pipeline.sosfilt_record(x, pipeline.design_sos("lowpass", 100.0, corners=2, freq=5.0), True)
    np.testing.assert_allclose(y[0, :200], y[0, :200:-1], atol=1e-6)


def test_seisbench_norm_kwargs_and_int32_ingest(tmp_path):
    """SeisBench's constructor kwargs of the window pre-processing are accepted from ``model_args``; int32 counts stay
    int32 on the wire; the version sort of from_pretrained is type-stable."""
    e = vb.EQTransformer(norm_amp_per_comp=True, norm_detrend=True, peak_scope="window")
    p = e._params(e._argdict({}), [0.0, 0.0, 0.0])
    assert p.peak_scope == 0 and p.norm_detrend == 1  # norm_amp_per_comp forces the per-component peak
    d = vb.PhaseNet()
    p = d._params(d._argdict({}), [0.0, 0.0, 0.0])
    assert (d.norm_amp_per_comp, d.norm_detrend, p.peak_scope, p.norm_detrend) == (False, False, 0, 0)
    w = vb.PhaseNet(peak_scope="window")
    assert w._params(w._argdict({}), [0.0, 0.0, 0.0]).peak_scope == 1
    with pytest.raises(ValueError, match="peak_scope"):
        vb.PhaseNet(peak_scope="trace")
    x = np.round(synthetic_record(2, 5000)).astype(np.int32)
    t0 = UTCDateTime("2021-01-01")
    hdr = dict(network="XX", station="I", location="", sampling_rate=100.0)
    st = Stream([Trace(x[i], dict(hdr, channel="HH" + c, starttime=t0)) for i, c in enumerate("ZN")])
    (_, arr), = d.stream_to_arrays(list(st), d._argdict({}))
    assert arr.dtype == np.int32 and np.array_equal(arr[:2], x[:2]) and not arr[2].any()
    st[1].data = st[1].data.astype(np.float64)
    (_, arr), = d.stream_to_arrays(list(st), d._argdict({}))
    assert arr.dtype == np.float32
    # records assembled in a caller-provided buffer (the pinned ring of _run)
    bufs = []

    def alloc(shape, dtype):
        bufs.append(np.full(shape, 7, dtype=dtype))
        return bufs[-1]

    (_, arr), = list(d._iter_stream_arrays(list(st), d._argdict({}), alloc))
    assert arr is bufs[0] and not arr[2].any() and np.array_equal(arr[0], x[0])
    # a 125 Hz trace cannot be resampled without ObsPy: loud, not silent
    with pytest.raises((NotImplementedError, RuntimeError)):
        d.resample_trace(Trace(x[0], dict(hdr, channel="HHZ", starttime=t0, sampling_rate=125.0)))
    # version sort: "1", "2", "10", "2b" must not raise and "10" wins over "2"
    import json as _json, os as _os

    root = tmp_path / "phasenet"
    root.mkdir()
    for v in ("1", "2", "10", "2b"):
        (root / f"x.json.v{v}").write_text(_json.dumps({"model_args": {}}))
        (root / f"x.vpw.v{v}").write_bytes(b"")
    old = _os.environ.get("VOLPICK_B200_CACHE")
    _os.environ["VOLPICK_B200_CACHE"] = str(tmp_path)
    try:
        js, _ = weights_io.find_weights("phasenet", "x")
        assert js.endswith(".v2b") or js.endswith(".v10")
        assert weights_io.find_weights("phasenet", "x", "10")[0].endswith(".v10")
    finally:
        if old is None:
            del _os.environ["VOLPICK_B200_CACHE"]
        else:
            _os.environ["VOLPICK_B200_CACHE"] = old


def test_integration_doc_line_numbers(repo_root):
    """INTEGRATION.md cites the header line of every entry point it maps a SeisBench function to: keep them exact."""
    import os
    import re

    hdr = open(os.path.join(repo_root, "include", "volpick_b200.h")).read().splitlines()
    doc = open(os.path.join(repo_root, "INTEGRATION.md")).read()
    refs = re.findall(r"`(vp_\w+)`:(\d+)", doc)
    assert len(refs) >= 15
    for fn, line in refs:
        decl = hdr[int(line) - 1]
        assert decl.startswith("VP_API") and re.search(r"\b%s\(" % fn, decl), f"{fn} is not declared at include/volpick_b200.h:{line}: {decl!r}"
