"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): window / stacking / pick indices bit-exact given identical
probability traces; fp32 probabilities within 1e-4 absolute of the oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import volpick_b200 as vb
from oracle import nets, pipeline
from volpick_b200 import _lib
from volpick_b200.synthetic import synthetic_record, synthetic_stream, station_start

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


def _windows(kind, n_windows, seed):
    """Pre-normalised windows cut around the first synthetic event of a record (so that the
    probabilities are not all ~0), starts 250 samples apart."""
    L = pipeline.IN_SAMPLES[kind]
    x, events = synthetic_record(seed, 120_000, return_events=True)
    ip = next(p for p, _ in events if p > L)
    starts = np.clip(ip - L // 2 - 250 * (n_windows // 2) + 250 * np.arange(n_windows, dtype=np.int64), 0, x.shape[1] - L)
    return pipeline.prenorm(pipeline.cut_windows(x, starts, L), kind)


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("kind,L,taper", [("eqtransformer", 6000, 1), ("phasenet", 3001, 0)])
@pytest.mark.parametrize("scope", ["channel", "window"])
@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_slice_normalize(lib, kind, L, taper, scope, dtype):
    n = 40_123
    x = synthetic_record(11, n)
    x[2, :] = 0.0  # zero-filled missing component must stay exactly 0
    if dtype == "i32":
        x = np.round(x).astype(np.int32)
    starts = pipeline.window_starts(n, L, L - 777)
    ref = pipeline.prenorm(pipeline.cut_windows(x.astype(np.float32), starts, L), kind, "peak", scope)
    d_x = torch.from_numpy(x).cuda()
    d_s = torch.from_numpy(starts).cuda()
    out = torch.empty((len(starts), 3, L), dtype=torch.float32, device="cuda")
    _lib.check(lib.vp_slice_normalize(d_x.data_ptr(), 0 if dtype == "f32" else 1, n, n, d_s.data_ptr(), len(starts), L,
                                      _lib.PEAK_SCOPE[scope], taper, out.data_ptr(), _stream()))
    got = out.cpu().numpy()
    assert np.all(got[:, 2] == 0)
    np.testing.assert_allclose(got, ref, atol=2e-6, rtol=0)


# ------------------------------------------------------------------------------------------ forward
def _tap_report(model, kind, sd, x, groups_of):
    taps = {}
    xt = torch.from_numpy(x)
    if kind == "eqtransformer":
        ref_out = torch.stack(nets.eqtransformer_forward(sd, xt, taps), dim=1).numpy()
    else:
        ref_out = nets.phasenet_forward(sd, xt, taps).numpy()
    xd = xt.cuda()
    B = x.shape[0]
    report = []
    for name in model.tap_names():
        got = model.forward_tap(xd, name, precision="fp32").cpu().numpy()
        ref = groups_of(name, taps, B)
        if ref is None:
            continue
        got = got.reshape(ref.shape) if got.size == ref.size else got
        if got.shape != ref.shape:  # concat buffers: compare the valid half
            got = got.reshape(B, -1, ref.shape[-1])[:, : ref.shape[1]]
        err = float(np.abs(got - ref).max())
        scale = float(np.abs(ref).max())
        report.append((name, err, scale))
    out = model.forward(xd, precision="fp32")
    got_out = (torch.stack(out, dim=1) if isinstance(out, tuple) else out).cpu().numpy()
    return report, got_out, ref_out


def _eqt_ref(name, taps, B):
    if name in taps:
        return taps[name].numpy()
    if name == "pick_lstm":
        return np.stack([taps["pick0_lstm"].numpy(), taps["pick1_lstm"].numpy()])
    if name == "pick_attn":
        return np.stack([taps["pick0_attn"].numpy(), taps["pick1_attn"].numpy()])
    if name.startswith("dec"):
        return np.stack([taps[f"d{g}_{name}"].numpy() for g in range(3)])
    return None


def _pn_ref(name, taps, B):
    return taps[name].numpy() if name in taps else None


def test_eqt_forward_layerwise(eqt, sd_eqt):
    x = _windows("eqtransformer", 6, seed=21)
    report, got, ref = _tap_report(eqt, "eqtransformer", sd_eqt, x, _eqt_ref)
    lines = [f"{n:16s} max|diff|={e:.3e}  max|ref|={s:.3e}" for n, e, s in report]
    print("\n".join(lines))
    assert len(report) >= 31
    worst = [(n, e, s) for n, e, s in report if e > 1e-4 * max(1.0, s)]
    assert not worst, "\n".join(lines)
    assert ref.max() > 0.5, "test windows must contain detections"
    assert float(np.abs(got - ref).max()) <= PROB_ATOL, float(np.abs(got - ref).max())


def test_pn_forward_layerwise(pn, sd_pn):
    x = _windows("phasenet", 6, seed=22)
    report, got, ref = _tap_report(pn, "phasenet", sd_pn, x, _pn_ref)
    lines = [f"{n:16s} max|diff|={e:.3e}  max|ref|={s:.3e}" for n, e, s in report]
    print("\n".join(lines))
    assert len(report) >= 17
    worst = [(n, e, s) for n, e, s in report if e > 1e-4 * max(1.0, s)]
    assert not worst, "\n".join(lines)
    assert ref[:, :2].max() > 0.5
    assert float(np.abs(got - ref).max()) <= PROB_ATOL
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("kind,B", [("eqtransformer", 1), ("eqtransformer", 37), ("phasenet", 1), ("phasenet", 130)])
def test_forward_batch_sizes(eqt, pn, sd_eqt, sd_pn, kind, B):
    model, sd = (eqt, sd_eqt) if kind == "eqtransformer" else (pn, sd_pn)
    x = _windows(kind, B, seed=30 + B)
    out = model(torch.from_numpy(x).cuda())
    got = (torch.stack(out, dim=1) if isinstance(out, tuple) else out).cpu().numpy()
    ref = pipeline.forward_batches(kind, sd, x, 64).transpose(0, 2, 1)
    assert float(np.abs(got - ref).max()) <= PROB_ATOL


@pytest.mark.parametrize("precision,atol", [("f16x3", PROB_ATOL), ("bf16", 5e-2)])
def test_eqt_forward_tensor_core(eqt, sd_eqt, precision, atol):
    """tcgen05 path: f16x3 (fp16 hi/lo split, 3 MMAs) must hold the fp32 tolerance; bf16 is reported with its own."""
    x = _windows("eqtransformer", 9, seed=21)
    xd = torch.from_numpy(x).cuda()
    ref = torch.stack(nets.eqtransformer_forward(sd_eqt, torch.from_numpy(x)), dim=1).numpy()
    got = torch.stack(eqt.forward(xd, precision=precision), dim=1).cpu().numpy()
    taps = {}
    nets.eqtransformer_forward(sd_eqt, torch.from_numpy(x), taps)
    e6 = eqt.forward_tap(xd, "enc6", precision=precision).cpu().numpy().reshape(taps["enc6"].shape)
    r6 = eqt.forward_tap(xd, "res6", precision=precision).cpu().numpy().reshape(taps["res6"].shape)
    print(f"{precision}: enc6 max|diff| = {np.abs(e6 - taps['enc6'].numpy()).max():.3e} (max|ref| {taps['enc6'].abs().max():.2f}); "
          f"res6 max|diff| = {np.abs(r6 - taps['res6'].numpy()).max():.3e} (max|ref| {taps['res6'].abs().max():.2f}); "
          f"probabilities max|diff| = {np.abs(got - ref).max():.3e}")
    assert ref.max() > 0.5
    if precision == "f16x3":  # the res-CNN stack on the tensor cores (k = 3 and right-padded k = 2 convs, fp32 residual stream)
        assert float(np.abs(r6 - taps["res6"].numpy()).max()) <= 1e-4 * float(taps["res6"].abs().max())
    assert float(np.abs(got - ref).max()) <= atol


PN_TC_TAPS = ["inc", "down0_same", "down0_down", "down1_same", "down1_down", "down2_same", "down2_down", "down3_same",
              "down3_down", "down4_same", "up0_same", "up1_same", "up2_same", "up3_same"]


@pytest.mark.parametrize("precision,atol", [("f16x3", PROB_ATOL), ("bf16", 5e-2)])
def test_pn_forward_tensor_core(pn, sd_pn, precision, atol):
    """PhaseNet on tcgen05: stride-4 convs / ConvTranspose1d as k = 2 convs on the row-reshaped buffers, two-source
    concat conv, fused 1x1 + softmax head.  Layer by layer against the oracle, then the probabilities."""
    x = _windows("phasenet", 7, seed=22)
    xt = torch.from_numpy(x)
    xd = xt.cuda()
    taps = {}
    ref = nets.phasenet_forward(sd_pn, xt, taps).numpy()
    lines, worst = [], []
    for name in PN_TC_TAPS:
        r = taps[name].numpy()
        got = pn.forward_tap(xd, name, precision=precision).cpu().numpy().reshape(r.shape)
        err, scale = float(np.abs(got - r).max()), float(np.abs(r).max())
        lines.append(f"{name:12s} max|diff|={err:.3e}  max|ref|={scale:.3e}")
        if precision == "f16x3" and err > 1e-4 * max(1.0, scale):
            worst.append(name)
    print("\n".join(lines))
    got = pn.forward(xd, precision=precision).cpu().numpy()
    print(f"{precision}: probabilities max|diff| = {np.abs(got - ref).max():.3e}")
    assert not worst, "\n".join(lines)
    assert ref[:, :2].max() > 0.5
    assert float(np.abs(got - ref).max()) <= atol
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("B", [1, 130])
def test_pn_tensor_core_batch_sizes(pn, sd_pn, B):
    x = _windows("phasenet", B, seed=60 + B)
    got = pn.forward(torch.from_numpy(x).cuda(), precision="f16x3").cpu().numpy()
    ref = pipeline.forward_batches("phasenet", sd_pn, x, 64).transpose(0, 2, 1)
    assert float(np.abs(got - ref).max()) <= PROB_ATOL


def test_pn_annotate_tensor_core_exact_mode(pn, sd_pn):
    """f16x3 end to end for PhaseNet: probabilities within 1e-4 of the oracle, picks identical to the fp32 CUDA-core path."""
    x = synthetic_record(41, 45_000)
    thr = {"P_threshold": 0.2, "S_threshold": 0.2}
    a32 = pn._argdict(dict(overlap=1500, blinding=(0, 0), stacking="avg", precision="fp32", **thr))
    atc = pn._argdict(dict(overlap=1500, blinding=(0, 0), stacking="avg", precision="f16x3", **thr))
    ann32, trig32, _ = pn.annotate_array(x, a32, True, pn._thresholds(a32))
    anntc, trigtc, _ = pn.annotate_array(x, atc, True, pn._thresholds(atc))
    ref = pipeline.annotate_array("phasenet", sd_pn, x, 1500, (0, 0), "avg")
    ok = ~np.isnan(ref)
    assert float(np.abs(anntc.T[ok] - ref[ok]).max()) <= PROB_ATOL
    assert len(trig32) == len(trigtc) and len(trigtc) > 0
    assert np.array_equal(trigtc["label"], trig32["label"])
    assert np.abs(trigtc["s_peak"] - trig32["s_peak"]).max() <= 1


@pytest.mark.parametrize("precision", ["f16x3", "bf16", "fp32"])
@pytest.mark.parametrize("lo,hi", [(500, 5500), (0, 6000), (496, 5505), (1234, 2345), (5999, 6000), (0, 1), (3000, 3000)])
def test_forward_range_matches_full_forward(lib, eqt, precision, lo, hi):
    """vp_forward_range (blinding-aware decoder tail): the kept samples are bit-identical to vp_forward's."""
    x = torch.from_numpy(_windows("eqtransformer", 5, seed=23)).cuda()
    full = torch.stack(eqt.forward(x, precision=precision), dim=1)
    prec = _lib.PRECISION[precision]
    y = torch.full((5, 3, 6000), -7.0, dtype=torch.float32, device="cuda")
    need = _lib.check(lib.vp_forward_workspace_bytes(eqt._handle, 5, prec))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    _lib.check(lib.vp_forward_range(eqt._handle, C.c_void_p(x.data_ptr()), 5, C.c_void_p(y.data_ptr()), C.c_void_p(ws.data_ptr()),
                                    need, prec, lo, hi, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(y[:, :, lo:hi], full[:, :, lo:hi])
    with pytest.raises(_lib.VolpickError, match="bad sample range"):
        _lib.check(lib.vp_forward_range(eqt._handle, C.c_void_p(x.data_ptr()), 5, C.c_void_p(y.data_ptr()), C.c_void_p(ws.data_ptr()),
                                        need, prec, 10, 5, _stream()))


def test_annotate_tensor_core_exact_mode(eqt, sd_eqt):
    """f16x3 end to end: probabilities within 1e-4 of the oracle and identical picks to the fp32 CUDA-core path."""
    x = synthetic_record(40, 60_000)
    thr = {"P_threshold": 0.2, "S_threshold": 0.2, "detection_threshold": 0.3}
    a32 = eqt._argdict(dict(overlap=5500, blinding=(500, 500), stacking="avg", precision="fp32", **thr))
    atc = eqt._argdict(dict(overlap=5500, blinding=(500, 500), stacking="avg", precision="f16x3", **thr))
    ann32, trig32, _ = eqt.annotate_array(x, a32, True, eqt._thresholds(a32))
    anntc, trigtc, _ = eqt.annotate_array(x, atc, True, eqt._thresholds(atc))
    ref = pipeline.annotate_array("eqtransformer", sd_eqt, x, 5500, (500, 500), "avg")
    ok = ~np.isnan(ref)
    np.testing.assert_array_equal(np.isnan(anntc.T), np.isnan(ref))
    err = float(np.abs(anntc.T[ok] - ref[ok]).max())
    print(f"f16x3 annotate: max|prob - oracle| = {err:.3e}; fp32 path: {np.abs(ann32.T[ok] - ref[ok]).max():.3e}")
    assert err <= PROB_ATOL
    assert len(trigtc) == len(trig32) and len(trig32) > 0
    assert np.array_equal(trigtc["label"], trig32["label"])
    assert np.abs(trigtc["s_peak"] - trig32["s_peak"]).max() <= 1
    abf = eqt._argdict(dict(overlap=5500, blinding=(500, 500), stacking="avg", precision="bf16", **thr))
    annbf, trigbf, _ = eqt.annotate_array(x, abf, True, eqt._thresholds(abf))
    errbf = float(np.abs(annbf.T[ok] - ref[ok]).max())
    match = 0
    for t in trig32:
        cand = trigbf[trigbf["label"] == t["label"]]
        match += int(len(cand) > 0 and np.abs(cand["s_peak"] - t["s_peak"]).min() <= 1)
    print(f"bf16 annotate: max|prob - oracle| = {errbf:.3e}; picks {len(trigbf)} vs {len(trig32)}, matched within 1 sample: {match}")
    assert errbf <= 5e-2


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
@pytest.mark.parametrize("dtype", ["f32", "i32"])
def test_slice_forward_matches_two_call_sequence(lib, eqt, pn, sd_eqt, sd_pn, kind, dtype):
    """vp_slice_forward (K1 inside the first conv kernel, fp32 first layer) against vp_slice_normalize + vp_forward and
    against the oracle on the same record."""
    model, sd = (eqt, sd_eqt) if kind == "eqtransformer" else (pn, sd_pn)
    L = model.in_samples
    x = synthetic_record(44, 40_000)
    if dtype == "i32":
        x = np.round(x).astype(np.int32)
    starts = pipeline.window_starts(x.shape[1], L, L - 700)
    nw = len(starts)
    d_tr = torch.from_numpy(x).cuda()
    d_st = torch.from_numpy(starts).cuda()
    prec = _lib.PRECISION["f16x3"]
    need = _lib.check(lib.vp_forward_workspace_bytes(model._handle, nw, prec))
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    y1 = torch.empty((nw, 3, L), dtype=torch.float32, device="cuda")
    y2 = torch.empty_like(y1)
    taper = 1 if kind == "eqtransformer" else 0
    dt = 0 if dtype == "f32" else 1
    _lib.check(lib.vp_slice_forward(model._handle, d_tr.data_ptr(), dt, x.shape[1], x.shape[1], d_st.data_ptr(), nw, 0, taper,
                                    y1.data_ptr(), ws.data_ptr(), need, prec, 0, L, _stream()))
    xw = torch.empty((nw, 3, L), dtype=torch.float32, device="cuda")
    _lib.check(lib.vp_slice_normalize(d_tr.data_ptr(), dt, x.shape[1], x.shape[1], d_st.data_ptr(), nw, L, 0, taper, xw.data_ptr(), _stream()))
    _lib.check(lib.vp_forward(model._handle, xw.data_ptr(), nw, y2.data_ptr(), ws.data_ptr(), need, prec, _stream()))
    torch.cuda.synchronize()
    assert float((y1 - y2).abs().max()) <= 2e-5
    ref = pipeline.forward_batches(kind, sd, xw.cpu().numpy(), 64).transpose(0, 2, 1)
    assert float(np.abs(y1.cpu().numpy() - ref).max()) <= PROB_ATOL


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
@pytest.mark.parametrize("on_host", [True, False])
def test_annotate_chunked_two_lanes_bit_identical(eqt, pn, kind, on_host):
    """Chunks of a record alternate between two streams (own forward workspace each, piecewise H2D of host records):
    any chunking must give bit-identical annotations and triggers to the single-chunk run."""
    model = eqt if kind == "eqtransformer" else pn
    x = synthetic_record(43, 90_000)
    rec = x if on_host else torch.from_numpy(x).cuda()
    base = dict(overlap=model.in_samples - 500, blinding=(250, 250), stacking="avg", P_threshold=0.2, S_threshold=0.2)
    a1 = model._argdict(dict(base, chunk_windows=4096))
    ann1, trig1, trim1 = model.annotate_array(rec, a1, True, model._thresholds(a1))
    for chunk in (7, 32, 100):
        ac = model._argdict(dict(base, chunk_windows=chunk))
        annc, trigc, trimc = model.annotate_array(rec, ac, True, model._thresholds(ac))
        np.testing.assert_array_equal(annc.view(np.uint32), ann1.view(np.uint32))
        assert np.array_equal(trigc, trig1) and np.array_equal(np.asarray(trimc), np.asarray(trim1))


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_annotate_async_two_records_in_flight(eqt, pn, kind):
    """vp_annotate_begin / vp_annotate_end: two records in flight on two streams with their own workspaces give the results
    of the blocking call, whatever the order the handles are collected in."""
    model = eqt if kind == "eqtransformer" else pn
    recs = [torch.from_numpy(synthetic_record(50 + i, 70_000 + 1000 * i)).pin_memory() for i in range(4)]
    a = model._argdict(dict(overlap=model.in_samples - 600, P_threshold=0.2, S_threshold=0.2, chunk_windows=32))
    thr = model._thresholds(a)
    ref = [model.annotate_array(r, a, True, thr) for r in recs]
    streams = [torch.cuda.Stream() for _ in range(2)]
    wss = [torch.empty(model.annotate_workspace_bytes(80_000, a, True), dtype=torch.uint8, device="cuda") for _ in range(2)]
    h01 = [model.annotate_array_async(recs[i], a, True, thr, stream=streams[i], workspace=wss[i]) for i in range(2)]
    out = {1: h01[1].result(), 0: h01[0].result()}  # collected out of order
    h23 = [model.annotate_array_async(recs[2 + i], a, True, thr, stream=streams[i], workspace=wss[i]) for i in range(2)]
    out[2], out[3] = h23[0].result(), h23[1].result()
    for i in range(4):
        np.testing.assert_array_equal(out[i][0].view(np.uint32), ref[i][0].view(np.uint32))
        assert np.array_equal(out[i][1], ref[i][1]) and np.array_equal(out[i][2], ref[i][2])
    assert h01[0].result() is out[0]  # idempotent
    with pytest.raises(ValueError, match="workspace too small"):
        model.annotate_array_async(recs[0], a, True, thr, workspace=torch.empty(16, dtype=torch.uint8, device="cuda"))


def test_golden_window_probabilities(eqt, pn, golden):
    for name, g in golden.items():
        model = eqt if str(g["kind"]) == "eqtransformer" else pn
        out = model(torch.from_numpy(g["windows01"]).cuda())
        got = (torch.stack(out, dim=1) if isinstance(out, tuple) else out).cpu().numpy()
        ref = g["probs01"].transpose(0, 2, 1)
        assert float(np.abs(got - ref).max()) <= PROB_ATOL, name


# ------------------------------------------------------------------------------------------ K9
@pytest.mark.parametrize("L,ov,n,blind,mode", [
    (6000, 5500, 30_000, (500, 500), "avg"), (6000, 5500, 30_123, (500, 500), "max"), (3001, 1500, 20_000, (0, 0), "avg"),
    (3001, 1500, 20_000, (100, 50), "max"), (6000, 1800, 25_000, (500, 500), "avg"), (3001, 2900, 9_000, (0, 0), "avg"),
    (6000, 0, 18_500, (500, 500), "avg"), (6000, 5500, 6_000, (0, 0), "avg"),
])
def test_stack_bit_exact(lib, L, ov, n, blind, mode):
    rng = np.random.default_rng(n + ov)
    starts = pipeline.window_starts(n, L, ov)
    y = rng.random((len(starts), L, 3), dtype=np.float32)
    if mode == "max":
        y[rng.random(y.shape) < 0.001] = np.nan
    ref = pipeline.reassemble(pipeline.blind(y, blind), starts, L, ov, mode).astype(np.float32)
    d_y = torch.from_numpy(np.ascontiguousarray(y.transpose(0, 2, 1))).cuda()
    d_s = torch.from_numpy(starts).cuda()
    out = torch.empty((3, n), dtype=torch.float32, device="cuda")
    _lib.check(lib.vp_stack(d_y.data_ptr(), d_s.data_ptr(), len(starts), L, 3, ov, blind[0], blind[1], _lib.STACK[mode],
                            out.data_ptr(), n, _stream()))
    got = out.cpu().numpy().T
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_array_equal(got[ok].view(np.uint32), ref[ok].view(np.uint32))  # NumPy summation order reproduced
    bounds = torch.empty(6, dtype=torch.int64, device="cuda")
    _lib.check(lib.vp_nan_bounds(out.data_ptr(), 3, n, bounds.data_ptr(), _stream()))
    b = bounds.cpu().numpy().reshape(3, 2)
    for c in range(3):
        _, f, back = pipeline.trim_nan(ref[:, c])
        assert (b[c, 0], b[c, 1]) == (f, n - 1 - back)


def test_stack_rejects_bad_arguments(lib):
    t = torch.zeros(16, device="cuda")
    s = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert lib.vp_stack(t.data_ptr(), s.data_ptr(), 1, 4, 1, 0, 0, 0, 7, t.data_ptr(), 4, _stream()) == _lib.VP_ERR_ARG
    assert b"Stacking method" in lib.vp_last_error()
    assert lib.vp_stack(t.data_ptr(), s.data_ptr(), 1, 4, 1, 4, 0, 0, 0, t.data_ptr(), 4, _stream()) == _lib.VP_ERR_ARG


# ------------------------------------------------------------------------------------------ K10
def _gpu_picks(lib, x, thr, cap=None):
    n = len(x)
    cap = cap or max(n, 1)
    d_x = torch.from_numpy(x).cuda()
    picks = torch.zeros(cap * 32, dtype=torch.uint8, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    sb = lib.vp_pick_scratch_bytes(n)
    scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.vp_pick(d_x.data_ptr(), n, thr, np.float32(thr) / np.float32(2), 1, picks.data_ptr(), cap, count.data_ptr(),
                           scratch.data_ptr(), sb, _stream()))
    k = int(count.item())
    dt = np.dtype([("s0", "<i8"), ("s1", "<i8"), ("s_peak", "<i8"), ("value", "<f4"), ("label", "<i4")])
    arr = np.frombuffer(picks.cpu().numpy().tobytes(), dtype=dt, count=min(k, cap))
    return k, np.sort(arr, order="s0")


def test_picks_bit_exact_random(lib):
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(1, 5000))
        if trial % 4 == 0:
            x = rng.random(n, dtype=np.float32)
        else:
            x = np.abs(np.cumsum(rng.standard_normal(n))).astype(np.float32)
            x /= max(float(x.max()), 1e-6)
            x[rng.random(n) < 0.01] = np.nan
        if trial % 5 == 0:
            x = np.round(x, 1)  # plateaus: first-argmax tie rule
        thr = float(np.float32(rng.uniform(0.05, 0.9)))
        ref = pipeline.picks_from_trace(x, np.float32(thr))
        k, got = _gpu_picks(lib, x, thr)
        assert k == len(ref), (trial, k, len(ref))
        for g, r in zip(got, ref):
            assert (g["s0"], g["s1"], g["s_peak"]) == r[:3] and g["value"] == np.float32(r[3]) and g["label"] == 1


def test_picks_long_runs_and_edges(lib):
    n = 1_000_003
    x = np.full(n, 0.5, np.float32)  # one run spanning the whole trace, peak in the middle of a chunk
    x[777_777] = 0.9
    x[777_778] = 0.9  # tie -> first
    k, got = _gpu_picks(lib, x, 0.4)
    assert k == 1 and (got[0]["s0"], got[0]["s1"], got[0]["s_peak"]) == (0, n - 1, 777_777)
    k, got = _gpu_picks(lib, np.zeros(100, np.float32), 0.3)
    assert k == 0
    x = np.zeros(64, np.float32)
    x[63] = 1.0  # trigger on the last sample
    k, got = _gpu_picks(lib, x, 0.3)
    assert k == 1 and (got[0]["s0"], got[0]["s1"], got[0]["s_peak"]) == (63, 63, 63)
    # overflow is reported through the count, never silently truncated
    x = np.tile(np.array([1.0, 0.0], np.float32), 50)
    k, got = _gpu_picks(lib, x, 0.3, cap=8)
    assert k == 50 and len(got) == 8


def test_pick_labels_tiles_alignment_and_trim(lib):
    """vp_pick_labels: all labels + _trim_nan bounds in one launch; runs crossing the 4096-sample tiles, label rows at
    every 16-byte misalignment (pred_len % 4 != 0), a label without picks (threshold 0)."""
    rng = np.random.default_rng(11)
    dt = np.dtype([("s0", "<i8"), ("s1", "<i8"), ("s_peak", "<i8"), ("value", "<f4"), ("label", "<i4")])
    for n in (4093, 4096, 12_289, 30_001, 70_002, 3):
        ann = np.empty((3, n), np.float32)
        for c in range(3):
            w = np.abs(np.cumsum(rng.standard_normal(n))).astype(np.float32)
            w /= max(float(w.max()), 1e-6)
            if c == 1:
                w = np.round(w, 1)  # plateaus and long runs
            w[rng.random(n) < 0.002] = np.nan
            lead, trail = int(rng.integers(0, min(n, 600))), int(rng.integers(0, min(n, 600)))
            w[:lead] = np.nan
            w[n - trail:] = np.nan
            ann[c] = w
        thr = np.array([0.3, 0.45, 0.0], np.float32)
        thr_off = thr / np.float32(2)
        d_a = torch.from_numpy(ann).cuda()
        cap = n
        picks = torch.zeros(cap * 32, dtype=torch.uint8, device="cuda")
        count = torch.zeros(1, dtype=torch.int64, device="cuda")
        bounds = torch.empty(6, dtype=torch.int64, device="cuda")
        _lib.check(lib.vp_pick_labels(d_a.data_ptr(), 3, n, thr.ctypes.data, thr_off.ctypes.data, picks.data_ptr(), cap,
                                      count.data_ptr(), bounds.data_ptr(), _stream()))
        k = int(count.item())
        arr = np.sort(np.frombuffer(picks.cpu().numpy().tobytes(), dtype=dt, count=k), order=["label", "s0"])
        b = bounds.cpu().numpy().reshape(3, 2)
        ref_all = []
        for c in range(3):
            _, f, back = pipeline.trim_nan(ann[c])
            if np.isnan(ann[c]).all():
                assert (b[c, 0], b[c, 1]) == (n, -1)
            else:
                assert (b[c, 0], b[c, 1]) == (f, n - 1 - back), (n, c)
            if thr[c] > 0:
                ref_all += [(c,) + tuple(r) for r in pipeline.picks_from_trace(ann[c], thr[c])]
        assert k == len(ref_all), (n, k, len(ref_all))
        for g, r in zip(arr, ref_all):
            assert (g["label"], g["s0"], g["s1"], g["s_peak"]) == r[:4] and g["value"] == np.float32(r[4]), (n, g, r)
        # bounds are optional
        count.zero_()
        _lib.check(lib.vp_pick_labels(d_a.data_ptr(), 3, n, thr.ctypes.data, thr_off.ctypes.data, picks.data_ptr(), cap,
                                      count.data_ptr(), None, _stream()))
        assert int(count.item()) == k


def test_picks_full_size_properties(lib):
    """Station-day size: one launch, picks sorted/disjoint, every pick satisfies the trigger rule (size-independent)."""
    n = 8_640_000
    rng = np.random.default_rng(5)
    x = np.abs(np.cumsum(rng.standard_normal(n))).astype(np.float32)
    x = (x % 97.0) / 97.0
    thr = np.float32(0.6)
    k, got = _gpu_picks(lib, x, float(thr))
    assert k == len(got) and k > 0
    above = x > thr / np.float32(2)
    n_runs_on = 0
    edges = np.flatnonzero(np.diff(np.concatenate(([0], above.view(np.int8), [0]))))
    for a, b in zip(edges[::2], edges[1::2]):
        if (x[a:b] > thr).any():
            n_runs_on += 1
    assert k == n_runs_on
    assert np.all(got["s0"][1:] > got["s1"][:-1])
    for g in got[:: max(1, k // 500)]:
        s0, s1, sp = int(g["s0"]), int(g["s1"]), int(g["s_peak"])
        a = s0
        while a > 0 and above[a - 1]:
            a -= 1
        assert x[s0] > thr and not (x[a:s0] > thr).any()
        assert above[s0:s1 + 1].all() and (s1 + 1 == n or not above[s1 + 1])
        assert sp == s0 + int(np.argmax(x[s0:s1 + 1])) and g["value"] == x[sp]


@pytest.mark.parametrize("kind", ["eqtransformer", "phasenet"])
def test_pick_windows_evaluate_path(lib, eqt, pn, kind):
    """Window-level picks (the reference's evaluate(): eval_taks0.py:46-56 per window inside window_borders): bit-exact
    against the oracle's trigger rule applied to the same probability rows."""
    model = eqt if kind == "eqtransformer" else pn
    L = model.in_samples
    B = 23
    x = torch.from_numpy(_windows(kind, B, seed=71)).cuda()
    rng = np.random.default_rng(7)
    lo = rng.integers(0, L // 2, B)
    hi = lo + rng.integers(1, L // 2, B)
    hi[3] = lo[3]  # empty border
    lo[5], hi[5] = 0, L
    borders = np.stack([lo, hi], axis=1).astype(np.int64)
    for wb in (borders, None):
        got, y = model.pick_windows(x, wb, P_threshold=0.12, S_threshold=0.2, probabilities=True)
        yh = y.cpu().numpy()
        assert sorted(got) == ["P", "S"]
        n_total = 0
        for lab, thr in (("P", 0.12), ("S", 0.2)):
            c = model.labels.index(lab)
            for b in range(B):
                a, e = (0, L) if wb is None else (int(borders[b, 0]), int(borders[b, 1]))
                ref = pipeline.picks_from_trace(yh[b, c, a:e], np.float32(thr))
                picks, scores = got[lab][b]
                assert list(picks) == [r[2] for r in ref], (lab, b)
                assert list(scores) == [np.float32(r[3]) for r in ref]
                n_total += len(ref)
        assert n_total > 0
    # one threshold for all phases, detections on request
    got = model.pick_windows(x, None, threshold=0.3)
    assert sorted(got) == ["P", "S"]
    if kind == "eqtransformer":
        assert "Detection" in model.pick_windows(x, None, threshold=0.3, detection_threshold=0.5)


@pytest.mark.parametrize("spec", [
    ("highpass", dict(freq=0.5, corners=2, zerophase=True)),   # model_training/test_onephase.ipynb cell 43
    ("highpass", dict(freq=0.3)),                                # volpick/data/utils.py:702 (ObsPy defaults: 4 corners, one pass)
    ("bandpass", dict(freqmin=1.0, freqmax=20.0)),               # volpick/data/utils.py:704
    ("lowpass", dict(freq=10.0, corners=3, zerophase=True)),
])
@pytest.mark.parametrize("n,dtype", [(1, "f32"), (2047, "f32"), (2048, "i32"), (100_003, "f32"), (8_640_000, "f32")])
def test_sosfilt_matches_scipy(pn, spec, n, dtype):
    """Stream pre-filter on the device (block-parallel biquad cascade in float64) against scipy.signal.sosfilt, the routine
    ObsPy's filters call.  Tolerance: float32 rounding of the output (the float64 results agree to ~1e-13)."""
    ftype, kw = spec
    kw = dict(kw)
    zerophase = kw.pop("zerophase", False)
    sos = pipeline.design_sos(ftype, 100.0, **kw)
    rng = np.random.default_rng(n % 97)
    x = (rng.standard_normal((3, n)) * 1000.0 + 250.0).astype(np.float32)  # counts-like with an offset (step response)
    if dtype == "i32":
        x = np.round(x).astype(np.int32)
    ref = pipeline.sosfilt_record(x, sos, zerophase)
    got = pn.filter_record(x, sos, zerophase).cpu().numpy()
    scale = float(np.abs(ref).max()) + 1.0
    assert got.shape == ref.shape and float(np.abs(got - ref).max()) <= 2e-7 * scale
    # the host design of the model mirror equals the oracle's
    m = vb.PhaseNet.from_pretrained("volpick")
    m.filter_args, m.filter_kwargs = [ftype], dict(spec[1])  # set as attributes, as model_training/test_onephase.ipynb cell 43 does
    sos2, zp2 = m.design_filter()
    np.testing.assert_array_equal(sos2, sos)
    assert zp2 == zerophase


def test_annotate_with_filter_matches_oracle(pn, sd_pn):
    """filter_args / filter_kwargs on the model: the record is filtered on the device before it is cut into windows."""
    x = synthetic_record(45, 36_000) + np.float32(500.0)  # offset: the high-pass matters
    m = vb.PhaseNet.from_pretrained("volpick").cuda()
    m.filter_args, m.filter_kwargs = ["highpass"], {"freq": 0.5, "corners": 2, "zerophase": True}
    a = m._argdict(dict(overlap=1500, P_threshold=0.2, S_threshold=0.2))
    m.annotate_stream_pre([], a)  # designs the sections into the argdict (no trace to resample)
    assert a["_sos"] is not None
    ann, trig, _ = m.annotate_array(x, a, True, m._thresholds(a))
    xf = pipeline.sosfilt_record(x, pipeline.design_sos("highpass", 100.0, corners=2, freq=0.5), True)
    ref = pipeline.annotate_array("phasenet", sd_pn, xf, 1500, (0, 0), "avg")
    ok = ~np.isnan(ref)
    assert float(np.abs(ann.T[ok] - ref[ok]).max()) <= PROB_ATOL
    assert len(trig) > 0


# ------------------------------------------------------------------------------------------ whole path
def _oracle_triggers(kind, sd, x, overlap, blinding, stacking, thr):
    ann = pipeline.annotate_array(kind, sd, x, overlap, blinding, stacking)
    picks, offsets = pipeline.classify_array(kind, ann, thr)
    return ann, picks, offsets


@pytest.mark.parametrize("kind,n,overlap,blinding,stacking", [
    ("phasenet", 360_000, 1500, (0, 0), "avg"),          # BASELINE.json configs[0]: one station-hour, defaults
    ("eqtransformer", 60_000, 5500, (500, 500), "avg"),  # configs[1] settings on ten minutes
    ("eqtransformer", 21_234, 1800, (500, 500), "max"),
    ("phasenet", 10_000, 2000, (200, 300), "max"),
])
@pytest.mark.parametrize("precision", ["default", "fp32"])
def test_annotate_matches_oracle(eqt, pn, sd_eqt, sd_pn, kind, n, overlap, blinding, stacking, precision):
    model, sd = (eqt, sd_eqt) if kind == "eqtransformer" else (pn, sd_pn)
    if precision == "fp32" and model.precision == "fp32":
        pytest.skip("fp32 is this model's default")
    x = synthetic_record(40, n)
    thr = {"P_threshold": 0.2, "S_threshold": 0.2, "detection_threshold": 0.3}
    ann, picks, offsets = _oracle_triggers(kind, sd, x, overlap, blinding, stacking, thr)
    extra = {} if precision == "default" else {"precision": precision}
    argdict = model._argdict(dict(overlap=overlap, blinding=blinding, stacking=stacking, **thr, **extra))
    got_ann, trig, trim = model.annotate_array(x, argdict, True, model._thresholds(argdict))
    # also from a CUDA-resident trace
    got_ann2, trig2, _ = model.annotate_array(torch.from_numpy(x).cuda(), argdict, True, model._thresholds(argdict))
    np.testing.assert_array_equal(got_ann, got_ann2)
    np.testing.assert_array_equal(trig, trig2)
    np.testing.assert_array_equal(np.isnan(got_ann.T), np.isnan(ann))
    ok = ~np.isnan(ann)
    assert float(np.abs(got_ann.T[ok] - ann[ok]).max()) <= PROB_ATOL
    labels = pipeline.LABELS[kind]
    for li, lab in enumerate(labels):
        assert trim[li, 0] == offsets[lab]
    # picks from the GPU trace == oracle pick rule applied to the SAME (GPU) probability trace: bit-exact
    th_by_label = model._thresholds(argdict)
    for li, lab in enumerate(labels):
        if th_by_label[li] <= 0:
            continue
        col, f, _ = pipeline.trim_nan(got_ann[li])
        ref = [(s0 + f, s1 + f, sp + f, v) for s0, s1, sp, v in pipeline.picks_from_trace(col, np.float32(th_by_label[li]))]
        mine = trig[trig["label"] == li]
        assert len(mine) == len(ref)
        for g, r in zip(mine, ref):
            assert (g["s0"], g["s1"], g["s_peak"]) == r[:3] and g["value"] == np.float32(r[3])
        # and against the oracle's own end-to-end picks: same count, peaks within 1 sample
        assert len(mine) == len(picks[lab]), (lab, len(mine), len(picks[lab]))
        for g, r in zip(mine, picks[lab]):
            assert abs(int(g["s_peak"]) - r[2]) <= 1 and abs(float(g["value"]) - r[3]) <= PROB_ATOL


def test_golden_annotations(eqt, pn, golden):
    for name, g in golden.items():
        model = eqt if str(g["kind"]) == "eqtransformer" else pn
        x = synthetic_record(int(g["station"]), int(g["n_samples"]))
        thr = {"P_threshold": 0.2, "S_threshold": 0.2, "detection_threshold": float(g["thresholds"][0])}
        argdict = model._argdict(dict(overlap=int(g["overlap"]), blinding=tuple(int(v) for v in g["blinding"]),
                                      stacking=str(g["stacking"]), **thr))
        ann, trig, trim = model.annotate_array(x, argdict, True, model._thresholds(argdict))
        ref = g["annotation"]
        np.testing.assert_array_equal(np.isnan(ann.T), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert float(np.abs(ann.T[ok] - ref[ok]).max()) <= PROB_ATOL, name
        np.testing.assert_array_equal(trim[:, 0], g["trim"])
        gt = g["triggers"]
        assert len(trig) == len(gt), name
        for a, b in zip(trig, gt):
            assert a["label"] == int(b[0]) and abs(int(a["s_peak"]) - int(b[3])) <= 1 and abs(float(a["value"]) - b[4]) <= PROB_ATOL


def test_classify_stream_api(eqt, pn, sd_pn):
    """The README call (/root/reference/README.md:54-66) end to end on a stream."""
    st = synthetic_stream(0, 60_000)
    out = eqt.classify(st, batch_size=256, overlap=5500, blinding=(500, 500), stacking="avg", parallelism=None,
                       P_threshold=0.2, S_threshold=0.2, copy=True)
    assert str(out).startswith("ClassifyOutput from EQTransformer")
    assert len(out.picks) > 0 and {p.phase for p in out.picks} <= {"P", "S"}
    assert out.picks == sorted(out.picks) and all(p.trace_id == "XX.S0000." for p in out.picks)
    assert all(p.start_time <= p.peak_time <= p.end_time for p in out.picks)
    assert len(out.detections) > 0
    ann = eqt.annotate(st, overlap=4500, blinding=[1000, 1000])  # demo.ipynb cell 14
    assert [t.stats.channel for t in ann] == ["EQTransformer_Detection", "EQTransformer_P", "EQTransformer_S"]
    assert all(t.stats.npts == 60_000 - 2000 and t.stats.starttime == station_start(0) + 10.0 for t in ann)
    assert not np.isnan(ann[1].data).any() and ann[1].data.dtype == np.float32
    # PhaseNet: picks against the oracle run on the same record, as times
    x = synthetic_record(0, 60_000)
    picks = pn.classify(st).picks  # JSON thresholds P 0.39 / S 0.34, overlap 1500
    ref_ann = pipeline.annotate_array("phasenet", sd_pn, x)
    ref, _ = pipeline.classify_array("phasenet", ref_ann, {"P_threshold": 0.39, "S_threshold": 0.34})
    for phase in "PS":
        mine = [p for p in picks if p.phase == phase]
        assert len(mine) == len(ref[phase])
        for p, r in zip(mine, ref[phase]):
            assert abs((p.peak_time - station_start(0)) * 100 - r[2]) <= 1 + 1e-6
    pn_ann = pn.annotate(st, overlap=2500, blinding=[500, 500])  # demo.ipynb cell 13
    assert [t.stats.channel for t in pn_ann] == ["PhaseNet_P", "PhaseNet_S", "PhaseNet_N"]


def test_edge_records(eqt, pn):
    # shorter than one window: empty output, no error
    ann, trig, trim = pn.annotate_array(synthetic_record(1, 3000), None, True, [0.3, 0.3, 0.0])
    assert ann.shape == (3, 0) and len(trig) == 0
    out = pn.classify(synthetic_stream(1, 2000))
    assert len(out.picks) == 0
    # exactly one window
    ann, trig, trim = pn.annotate_array(synthetic_record(1, 3001), None, True, [0.3, 0.3, 0.0])
    assert ann.shape == (3, 3001) and not np.isnan(ann).any()
    # single-component stream (demo.ipynb cell 12 feeds one EHZ trace): the others are zero-filled
    st = synthetic_stream(2, 20_000).select(channel="HHZ")
    a = eqt.annotate(st)
    assert len(a) == 3 and a[0].stats.npts == 20_000 - 1000
    # gap: two segments -> two annotation traces per label
    full = synthetic_stream(3, 30_000)
    parts = vb.Stream()
    for tr in full:
        parts.append(vb.Trace(tr.data[:12_000], dict(network="XX", station="G", location="", channel=tr.stats.channel,
                                                      starttime=tr.stats.starttime, sampling_rate=100.0)))
        parts.append(vb.Trace(tr.data[15_000:], dict(network="XX", station="G", location="", channel=tr.stats.channel,
                                                      starttime=tr.stats.starttime + 150.0, sampling_rate=100.0)))
    a = pn.annotate(parts)
    assert len(a) == 6 and sorted({t.stats.npts for t in a}) == [12_000, 15_000]
    # pick buffer overflow is never a truncation: the C ABI reports it (tests/test_gpu_round2.py::test_pick_capacity_retry)
    # and the Python API re-runs the record with the reported count
    _, trig, _ = pn.annotate_array(synthetic_record(4, 30_000), None, False, [1e-6, 1e-6, 0.0], pick_capacity=1)
    assert len(trig) > 1


def test_station_day_properties(eqt):
    """Full-size config (BASELINE.json configs[1]) through size-independent properties."""
    n = 8_640_000
    x = synthetic_record(1000, n)
    argdict = eqt._argdict(dict(overlap=5500, blinding=(500, 500), stacking="avg", P_threshold=0.2, S_threshold=0.2))
    ann, trig, trim = eqt.annotate_array(x, argdict, True, eqt._thresholds(argdict))
    assert ann.shape == (3, n)
    assert np.all(trim[:, 0] == 500) and np.all(trim[:, 1] == n - 501)  # blinding trims 5 s on both ends
    body = ann[:, 500 : n - 500]
    assert not np.isnan(body).any() and body.min() >= 0 and body.max() <= 1
    assert np.isnan(ann[:, :500]).all() and np.isnan(ann[:, n - 500 :]).all()
    # idempotence / determinism
    ann2, trig2, _ = eqt.annotate_array(x, argdict, True, eqt._thresholds(argdict))
    np.testing.assert_array_equal(ann, ann2)
    np.testing.assert_array_equal(trig, trig2)
    # picks are exactly what the pick rule gives on the returned trace (bit-exact indices)
    for li in (1, 2):
        ref = pipeline.picks_from_trace(body[li], np.float32(0.2))
        mine = trig[trig["label"] == li]
        assert len(mine) == len(ref) and len(ref) > 100
        assert [(int(g["s0"]), int(g["s1"]), int(g["s_peak"])) for g in mine] == [(a + 500, b + 500, c + 500) for a, b, c, _ in ref]
    # translation property: the first hour annotated alone agrees with the day wherever the same windows cover it
    sub, _, _ = eqt.annotate_array(x[:, :360_000], argdict, True, [0, 0, 0])
    np.testing.assert_allclose(sub[:, 6000:354_000], ann[:, 6000:354_000], atol=1e-6)
