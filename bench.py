#!/usr/bin/env python
"""Benchmark of the continuous-waveform picking path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A "step" is one station-day (8,640,000 samples x 3 components, synthetic, 100 Hz) per rank through
the whole path of BASELINE.json configs[1]: EQTransformer-volpick, 6000-sample windows,
overlap=5500 (17,269 windows), blinding=(500,500), stacking="avg", P/S threshold 0.2.
``value`` = station-days/s over all ranks with the record resident in HBM; ``e2e`` = the same through
the host-buffer entry point (pinned host record -> vp_annotate -> picks on the host).
Prints ONE JSON line on rank 0.

Other BASELINE.json configurations (bench lines of every one are tracked under profiles/):
    --model phasenet [--samples 360000]        configs[0] / the unit of configs[2] (JSON thresholds, overlap 1500)
    --precision bf16                            bf16 mode of configs[3]; the line carries the +-1-sample pick-match rate
    --records 16 --steps 1000/N --quick         configs[2] / [3]: 1000 station-days sharded over N GPUs (tools/gpu_r02_scale.sh)
    --sweep                                     configs[4]: raw forward sweep B = 256 ... 8192, per-kernel-class times
    --classify-stream K                         the drop-in call itself: picker.classify(stream) on K station-days (``e2e_classify``)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DAY = 8_640_000
N_HOUR = 360_000
FLOP_PER_WINDOW = {"eqtransformer": 257.04e6, "phasenet": 38.92e6}
# Algorithmic FLOPs per window of the kernel classes that can dominate a step (2 x the reference's conv MACs on the
# up-sampled / full-length input; blinded margins and polyphase savings are NOT subtracted):
#   decb = decoder.convs.3-6 + head of the three EQTransformer decoders:
#          3 x (32*32*7*750 + 16*32*7*1500 + 16*16*9*3000 + 8*16*11*6000 + 1*8*11*6000) MAC = 79.92 MMAC
#   conv1d_f32 (PhaseNet fp32 path) = all Conv1d layers = 19.46 - 2.71 (ConvTranspose1d) MMAC
KCLASS_FLOP_PER_WINDOW = {
    ("eqtransformer", "decb"): 2 * 3 * (32 * 32 * 7 * 750 + 16 * 32 * 7 * 1500 + 16 * 16 * 9 * 3000 + 8 * 16 * 11 * 6000 + 8 * 11 * 6000),
    ("phasenet", "conv1d_f32"): 2 * (19.46e6 - (128 * 64 * 12 + 64 * 32 * 47 + 32 * 16 * 188 + 16 * 8 * 751) * 7),
    ("phasenet", "tcconv"): 38.92e6,  # every Conv1d / ConvTranspose1d of the network runs in tcconv_kernel
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/):
KCLASS_NCU_TRAFFIC = {("eqtransformer", "decb", "f16x3"): {"bytes_per_launch": 552.65e6 + 220.10e6, "windows_per_launch": 4096,
                                                            "source": "profiles/r02g_f16x3_top_kernels_ncu_full.md (decb2_kernel)"}}
CONFIGS = {
    # BASELINE.json configs[1] / configs[3]: EQTransformer, overlap 5500, blinding (500, 500), avg, P/S threshold 0.2
    "eqtransformer": dict(overlap=5500, blinding=(500, 500), stacking="avg", P_threshold=0.2, S_threshold=0.2),
    # BASELINE.json configs[0] / configs[2]: PhaseNet, "overlap default" (1500), no blinding, thresholds of the shipped JSON
    # (/root/reference/Final_models/volpick/phasenet/volpick.json.v1:10-13)
    "phasenet": dict(overlap=1500, blinding=(0, 0), stacking="avg", P_threshold=0.39, S_threshold=0.34),
}
WORKLOAD_LABEL = {
    "eqtransformer": "EQTransformer-volpick classify() on one synthetic 3-C 100 Hz station-day per GPU per step "
                     "(BASELINE.json configs[1]; the unit of configs[3])",
    "phasenet": "PhaseNet-volpick classify() on one synthetic 3-C 100 Hz station-day per GPU per step "
                "(the unit of BASELINE.json configs[2]; configs[0] is the same call on a station-hour)",
}


def measured_peaks():
    fallback = dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")  # B200_PROFILING.md
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            burst = float(p.get("bf16_tflops", p.get("bf16_tflops_burst", fallback["tf_burst"])))
            return dict(hbm=float(p["hbm_gbs"]), tf_burst=burst, tf_sustained=float(p.get("bf16_tflops_sustained", burst)),
                        src="measured")
        except (OSError, ValueError, KeyError, TypeError):
            pass
    return fallback


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                self._stop_evt.wait(0.02)
        except Exception:
            self._fallback()

    def _fallback(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(int(out[0]))
                self.sm_max = int(out[1])
                for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def oracle_station_hours_per_s(kind: str, steps: int, warmup: int, n_samples: int = N_HOUR):
    """The reference's CPU path (oracle port: torch-CPU nets + NumPy pipeline) on one station-hour per step."""
    import torch

    from oracle import nets, pipeline
    from volpick_b200 import weights_io
    from volpick_b200.synthetic import synthetic_record

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = CONFIGS[kind]
    sd = nets.state_dict_from_numpy(weights_io.load_weights(weights_io.find_weights(kind, "volpick")[1]))
    x = synthetic_record(1000, n_samples)
    thr = {"P_threshold": cfg["P_threshold"], "S_threshold": cfg["S_threshold"], "detection_threshold": 0.3}
    times = []
    n_picks = 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ann = pipeline.annotate_array(kind, sd, x, cfg["overlap"], cfg["blinding"], cfg["stacking"], batch_size=256)
        picks, _ = pipeline.classify_array(kind, ann, thr)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        n_picks = sum(len(v) for v in picks.values())
    nwin = len(pipeline.window_starts(n_samples, pipeline.IN_SAMPLES[kind], cfg["overlap"]))
    return times, nwin, cores, n_picks


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = args.model
    times, nwin, cores, n_picks = oracle_station_hours_per_s(kind, args.steps, args.warmup)
    total = sum(times)
    frac_day = N_HOUR / N_DAY
    value = frac_day * len(times) / total
    line = {
        "impl": "reference", "metric": "station_days_per_s", "value": value, "unit": "station-days/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "windows_per_s": nwin * len(times) / total,
        "config": workload_config(kind, "one station-hour (1/24 of the station-day) per step on the host CPU, extrapolated to station-days"),
        "cpu_baseline": {"value": value, "unit": "station-days/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} x 1 synthetic station-hour ({nwin} windows each), oracle port "
                                   f"(torch-CPU fp32 + NumPy), {cpu_model()}"},
        "e2e": {"value": value, "unit": "station-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(kind: str, note: str = ""):
    cfg = CONFIGS[kind]
    L = 6000 if kind == "eqtransformer" else 3001
    stride = L - cfg["overlap"]
    nwin = (N_DAY - L) // stride + 1
    if (nwin - 1) * stride + L < N_DAY:
        nwin += 1
    d = {"workload": WORKLOAD_LABEL[kind],
         "samples_per_record": N_DAY, "window": L, "overlap": cfg["overlap"], "blinding": list(cfg["blinding"]),
         "stacking": cfg["stacking"], "P_threshold": cfg["P_threshold"], "S_threshold": cfg["S_threshold"],
         "windows_per_record": nwin, "l2": "per-step working set (1.2 GB of window predictions) exceeds the 126 MB L2"}
    if note:
        d["note"] = note
    return d


def run_ours(args):
    import torch
    import torch.distributed as dist

    import volpick_b200 as vb
    from volpick_b200 import _lib, shard
    from volpick_b200.synthetic import synthetic_record

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    kind = args.model
    cls = vb.EQTransformer if kind == "eqtransformer" else vb.PhaseNet
    model = cls.from_pretrained("volpick").cuda(local_rank)
    lib = _lib.load()
    cfg = dict(CONFIGS[kind], precision=args.precision, chunk_windows=args.chunk)
    argdict = model._argdict(cfg)
    thresholds = model._thresholds(argdict)
    thresholds[0] = 0.3 if kind == "eqtransformer" else thresholds[0]
    n = args.samples
    # R distinct records per rank (station seeds 1000 + rank + world * j, the recipe of SURVEY.md 8(d)), cycled over the
    # steps: pinned host copies for the end-to-end arm, device-resident copies for `value`
    R = max(2, args.records)
    recs_host = [torch.from_numpy(synthetic_record(1000 + rank + world * j, n)).pin_memory() for j in range(R)]
    recs_dev = [r.cuda(non_blocking=True) for r in recs_host]
    torch.cuda.synchronize()

    def step_device(i):
        return model.annotate_array(recs_dev[i % R], argdict, False, thresholds)

    def step_host(i):
        return model.annotate_array(recs_host[i % R], argdict, False, thresholds)

    if args.profile_steps > 0:
        for i in range(args.profile_steps):
            step_device(i)
        torch.cuda.synchronize()
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        lib.vp_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            last = fn(i)
        e1.record()
        barrier()
        launches = lib.vp_launch_count(1)
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks, last

    # Two records in flight (vp_annotate_begin / vp_annotate_end on two streams, one workspace each): the tail of record i
    # (stacker, picker, result read-back, host gaps) and -- for host records -- the H2D copy of record i + 1 run under the
    # compute of the neighbouring record.  Every step still processes its own record completely inside the timed region;
    # for the end-to-end number that includes the copy of the record from pinned host memory and the read-back of its picks.
    ws_bytes = model.annotate_workspace_bytes(n, argdict, True)
    side = [torch.cuda.Stream() for _ in range(2)]
    side_ws = [torch.empty(ws_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]

    all_results = []  # (record index, triggers) of every step of the last pipelined run: gathered after the timed region

    def run_pipelined(recs, steps):
        pend, res = [], None
        del all_results[:]
        for i in range(steps):
            k = i & 1
            pend.append(model.annotate_array_async(recs[i % R], argdict, False, thresholds, stream=side[k], workspace=side_ws[k]))
            if len(pend) == 2:
                res = pend.pop(0).result()
                all_results.append(res[1])
        while pend:
            res = pend.pop(0).result()
            all_results.append(res[1])
        return res

    def timed_pipelined(recs, steps, warmup, sample_clocks=False):
        # W warm-up steps, and more of them until 0.4 s have passed: the first arm of a fresh process read 2 % low after three
        # 17 ms steps (lazy module loading, clock ramp) whichever arm came first
        t_w = time.perf_counter()
        run_pipelined(recs, max(2, warmup))
        while time.perf_counter() - t_w < 0.4:
            run_pipelined(recs, 2)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        lib.vp_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in side:
            st.wait_event(e0)
        last = run_pipelined(recs, steps)
        for st in side:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        barrier()
        launches = lib.vp_launch_count(1)
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks, last

    # device-resident arm first: it also brings the process to its steady state (lazy module loading, shared-memory opt-ins, SM
    # clocks) before the end-to-end arm, which gets its own full set of warm-up steps (with two warm-up steps and the cold start in
    # front of it the end-to-end number of a 5-step run read 49.7 against 58 - 59 station-days/s in longer runs)
    ms, launches, clocks, last = timed_pipelined(recs_dev, args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _, _, last_e2e = timed_pipelined(recs_host, args.steps, max(3, args.warmup))
    e2e_results = list(all_results)
    # the same with one blocking vp_annotate call per record (no overlap between records)
    if args.quick:
        ms_seq = ms_e2e_seq = float("nan")
    else:
        ms_seq, _, _, _ = timed(step_device, args.steps, max(1, args.warmup // 2))
        ms_e2e_seq, _, _, _ = timed(step_host, args.steps, max(1, args.warmup // 2))

    # ---- per-kernel-class CUDA-event timing over K more device-resident steps (events bracket every launch on the
    # launching stream inside the library; a separate pass so that the headline numbers above carry no event overhead)
    kernels = {}
    kernels_step_ms = None
    if rank == 0 and not args.quick:
        lib.vp_kernel_timing(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            step_device(i)
        e1.record()
        torch.cuda.synchronize()
        for k, name in enumerate(lib.vp_kernel_class_names().decode().split(",")):
            tot, cnt = C.c_double(0.0), C.c_int64(0)
            _lib.check(lib.vp_kernel_timing_read(k, C.byref(tot), C.byref(cnt)))
            if cnt.value:
                kernels[name] = {"ms_per_step": tot.value / args.steps, "launches_per_step": cnt.value / args.steps,
                                 "ms_per_launch": tot.value / cnt.value}
        lib.vp_kernel_timing(0)
        kernels_step_ms = e0.elapsed_time(e1) / args.steps
    n_trig = len(last[1])
    nwin = int(lib.vp_window_count(n, model.in_samples, argdict["overlap"]))
    days = n / N_DAY
    value = world * args.steps * days / (ms / 1e3)
    e2e_value = world * args.steps * days / (ms_e2e / 1e3)

    # final gather of picks -- the only exchange of the path: the triggers of EVERY record this rank processed in the
    # end-to-end run go to rank 0 (host gather; KBs).  Timed on the host next to the device-timed steps.
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    merged = shard.gather_triggers([(rank + world * i, trig) for i, trig in enumerate(e2e_results)], rank, world)
    gather_ms = 1e3 * (time.perf_counter() - t0)
    n_gathered = sum(len(t) for _, t in merged) if merged else 0

    # bf16 is reported separately (north_star): probabilities and picks of record 0 against the exact mode
    bf16_report = None
    if rank == 0 and args.precision == "bf16":
        a_ex = model._argdict(dict(CONFIGS[kind], precision="f16x3", chunk_windows=args.chunk))
        ann_b, trig_b, _ = model.annotate_array(recs_dev[0], argdict, True, thresholds)
        ann_x, trig_x, _ = model.annotate_array(recs_dev[0], a_ex, True, thresholds)
        ok = ~np.isnan(ann_x)
        hit = 0
        for t in trig_x:
            cand = trig_b[trig_b["label"] == t["label"]]
            hit += int(len(cand) > 0 and np.abs(cand["s_peak"] - t["s_peak"]).min() <= 1)
        bf16_report = {"max_abs_dprob_vs_exact": float(np.abs(ann_b[ok] - ann_x[ok]).max()), "triggers_exact": int(len(trig_x)),
                       "triggers_bf16": int(len(trig_b)), "pick_match_within_1_sample": hit / max(1, len(trig_x)),
                       "note": "record 0 of this run; exact = f16x3 mode (<= 1e-4 of the fp32 oracle)"}

    # the drop-in call itself: picker.classify(stream, ...) on ObsPy-like streams of 3 traces per station-day
    classify_report = None
    if args.classify_stream > 0 and n == N_DAY:
        classify_report = classify_stream_timing(model, kind, args, rank, world, barrier, dist if world > 1 else None)

    # ---- per-stage device timings (rank 0) for the roofline objects ---------------------------
    stages = {}
    if rank == 0 and not args.quick:
        stages = stage_timings(model, lib, recs_dev[0], argdict, thresholds, kind, precision=_lib.PRECISION[args.precision],
                               chunk=args.chunk)
    peaks = measured_peaks()
    line = None
    if rank == 0:
        fwd_ms = stages.get("forward_ms")
        flops = FLOP_PER_WINDOW[kind] * nwin
        fwd_tf = flops / (fwd_ms / 1e3) / 1e12 if fwd_ms else None
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
        share = {k: v["ms_per_step"] / sum(x["ms_per_step"] for x in kernels.values()) for k, v in kernels.items()}
        if dom and (kind, dom) in KCLASS_FLOP_PER_WINDOW:
            kf = KCLASS_FLOP_PER_WINDOW[(kind, dom)] * nwin  # per step
            achieved = kf / (kernels[dom]["ms_per_step"] / 1e3) / 1e12
            tr = KCLASS_NCU_TRAFFIC.get((kind, dom, args.precision))
            roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["tf_sustained"],
                        "traffic": (tr["bytes_per_launch"] if tr else None),
                        "traffic_note": (f"ncu dram bytes of one {tr['windows_per_launch']}-window launch, {tr['source']}" if tr else None),
                        "ms_per_launch": kernels[dom]["ms_per_launch"], "launches_per_step": kernels[dom]["launches_per_step"],
                        "share_of_kernel_time": share[dom],
                        "algorithmic_flop_per_window": KCLASS_FLOP_PER_WINDOW[(kind, dom)],
                        "peak_source": f"bf16_tflops_sustained ({peaks['src']}); kernel timed inside a long step",
                        "note": ("f16x3 issues 3 fp16 MMAs per algorithmic product (hi*hi + hi*lo + lo*hi) for fp32-grade results; "
                                 "the fraction is algorithmic FLOP/s against the single-pass bf16 peak") if args.precision == "f16x3" else
                                ("fp32 mode runs on the CUDA cores (FFMA); fraction is against the bf16 tensor peak" if args.precision == "fp32" else "")}
        else:
            roofline = {"kernel": "forward (all network kernels of vp_forward)", "bound": "tensor", "achieved": fwd_tf,
                        "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": (fwd_tf / peaks["tf_sustained"]) if fwd_tf else None,
                        "traffic": None, "peak_source": f"bf16_tflops_sustained ({peaks['src']})"}
        roofline["forward_tflops"] = fwd_tf  # whole network: 2 x MACs of SURVEY 8(d) / CUDA-event time of the forward stage
        roofline["forward_frac"] = (fwd_tf / peaks["tf_sustained"]) if fwd_tf else None
        line = {
            "metric": "station_days_per_s", "value": value, "unit": "station-days/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "records_in_flight": 2,
            "sequential": {"value": world * args.steps * days / (ms_seq / 1e3), "ms_per_step": ms_seq / args.steps,
                           "note": "one blocking vp_annotate call per record"},
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else args.precision,
            "data": "synthetic", "windows_per_s": value * nwin / days,
            "config": workload_config(kind),
            "e2e": {"value": e2e_value, "unit": "station-days/s", "h2d_bytes_per_step": 3 * n * 4,
                    "d2h_bytes_per_step": int(len(last_e2e[1]) * 32 + 56), "ms_per_step": ms_e2e / args.steps,
                    "records_in_flight": 2,
                    "sequential": {"value": world * args.steps * days / (ms_e2e_seq / 1e3), "ms_per_step": ms_e2e_seq / args.steps,
                                   "note": "one blocking vp_annotate call per record"}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stages,
            "kernels": {"per_class": kernels, "ms_per_step_with_events": kernels_step_ms,
                        "note": "CUDA events around every launch of the class on the launching stream, K extra steps"},
            "picks_per_record": n_trig, "gather_ms": gather_ms, "records_per_rank": R, "records_total": world * args.steps,
            "gather": {"ms": gather_ms, "records": world * len(e2e_results), "triggers": int(n_gathered),
                       "e2e_value_incl_gather": world * args.steps * days / ((ms_e2e + gather_ms) / 1e3),
                       "note": "triggers of every record of the end-to-end run gathered on rank 0 (two all_gather calls of byte tensors)"},
        }
        if bf16_report:
            line["bf16"] = bf16_report
        if classify_report:
            line["e2e_classify"] = classify_report
        if args.quick:  # long multi-GPU evidence runs: only the pipelined device-resident and end-to-end arms
            line["quick"] = True
            for k in ("sequential",):
                line.pop(k, None)
            line["e2e"].pop("sequential", None)
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            times, cw, cores, _ = oracle_station_hours_per_s(kind, 1, 1)
            v = (N_HOUR / N_DAY) * len(times) / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": "station-days/s", "cores": cores, "kind": "port",
                                    "sample": f"1 synthetic station-hour ({cw} windows, {sum(times):.1f} s) after 1 warm-up, "
                                              f"oracle port (torch-CPU fp32 + NumPy), extrapolated x24; {cpu_model()}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def classify_stream_timing(model, kind, args, rank, world, barrier, dist):
    """The README call (/root/reference/README.md:54-66) measured as a user makes it: ``picker.classify(stream, **kw)``
    on a stream of ``args.classify_stream`` station-days (3 traces each, pageable NumPy data as ObsPy holds it).  Host
    wall clock around the call (it contains host work: stream copy / merge, record assembly into the pinned ring, pick
    objects), max over ranks."""
    import torch

    from volpick_b200.stream import Stream, Trace
    from volpick_b200.synthetic import station_start, synthetic_record

    k = args.classify_stream
    traces = []
    for j in range(k):
        st_no = 1000 + rank + world * j
        x = synthetic_record(st_no, N_DAY)
        for i, c in enumerate("ZNE"):
            traces.append(Trace(x[i].copy(), {"network": "XX", "station": f"S{st_no:04d}", "location": "", "channel": "HH" + c,
                                              "starttime": station_start(st_no), "sampling_rate": 100.0}))
    stream = Stream(traces)
    kw = dict(CONFIGS[kind], precision=args.precision, batch_size=256, parallelism=None)
    out = {}
    import gc

    for copy in (True, False):
        res = model.classify(stream, copy=copy, **kw)  # warm-up: pins the staging ring, sizes the workspaces
        dts = []
        for _ in range(3):
            # a call leaves ~50,000 new objects per 8 station-days (picks and their time stamps): collect the generations
            # before the clock starts so that a full collection of the interpreter's ~10^6 long-lived objects (tens of ms,
            # every few calls) lands outside the timed call -- timeit's policy; the call itself runs with the collector on
            del res
            gc.collect()
            barrier()
            t0 = time.perf_counter()
            res = model.classify(stream, copy=copy, **kw)
            torch.cuda.synchronize()
            dts.append(time.perf_counter() - t0)
        dt = sorted(dts)[1]  # median of three calls
        if dist is not None:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out["copy_true" if copy else "copy_false"] = {"value": world * k / dt, "ms_per_record": 1e3 * dt / k, "picks": len(res.picks)}
    out.update({"unit": "station-days/s", "records_per_call": k, "h2d_bytes_per_record": 3 * N_DAY * 4,
                "note": "host wall clock of picker.classify(stream) per rank (median of 3 calls), max over ranks; copy=True is the README call "
                        "(deep copy of the stream first, as SeisBench does)"})
    return out


def run_sweep(args):
    """BASELINE.json configs[4]: raw forward sweep, B in {256 ... 8192} windows of (B, 3, 6000) EQTransformer and
    (B, 3, 3001) PhaseNet (standard_normal seed 0, per-channel demean + peak normalisation), windows/s and per-kernel-class
    milliseconds (CUDA events around every launch inside the library).  One JSON line per (model, precision, B)."""
    import torch

    import volpick_b200 as vb
    from volpick_b200 import _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(0)
    lib = _lib.load()
    peaks = measured_peaks()
    names = lib.vp_kernel_class_names().decode().split(",")
    for kind in ([args.model] if args.sweep_one_model else ["eqtransformer", "phasenet"]):
        model = (vb.EQTransformer if kind == "eqtransformer" else vb.PhaseNet).from_pretrained("volpick").cuda(0)
        L = model.in_samples
        g = torch.Generator(device="cpu").manual_seed(0)
        for prec in args.sweep_precisions.split(","):
            for B in (256, 512, 1024, 2048, 4096, 8192):
                x = torch.randn((B, 3, L), generator=g, dtype=torch.float32)
                x = x - x.mean(-1, keepdim=True)
                x = (x / (x.abs().amax(-1, keepdim=True) + 1e-10)).cuda()
                for _ in range(max(3, args.warmup)):
                    model.forward(x, precision=prec)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    model.forward(x, precision=prec)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                lib.vp_kernel_timing(1)
                for _ in range(args.steps):
                    model.forward(x, precision=prec)
                torch.cuda.synchronize()
                per = {}
                for k, name in enumerate(names):
                    tot, cnt = C.c_double(0.0), C.c_int64(0)
                    _lib.check(lib.vp_kernel_timing_read(k, C.byref(tot), C.byref(cnt)))
                    if cnt.value:
                        per[name] = {"ms": tot.value / args.steps, "launches": cnt.value / args.steps}
                lib.vp_kernel_timing(0)
                tf = FLOP_PER_WINDOW[kind] * B / (ms / 1e3) / 1e12
                print(json.dumps({"sweep": "forward", "model": kind, "precision": prec, "windows": B, "ms": ms,
                                  "windows_per_s": B / (ms / 1e3), "tflops": tf, "frac_of_bf16_sustained": tf / peaks["tf_sustained"],
                                  "kernels": per, "l2": "activations of B >= 256 windows exceed the 126 MB L2"}), flush=True)


def stage_timings(model, lib, rec_dev, argdict, thresholds, kind, reps: int = 3, precision: int = 0, chunk: int = 0):
    """CUDA-event timings of the four stages on the current stream, through the stage-level C ABI."""
    import torch

    from volpick_b200 import _lib

    L = model.in_samples
    n = rec_dev.shape[1]
    ov = argdict["overlap"]
    nwin = int(lib.vp_window_count(n, L, ov))
    starts = np.zeros(nwin, dtype=np.int64)
    cnt = C.c_int64(0)
    lib.vp_window_starts(n, L, ov, starts.ctypes.data, nwin, C.byref(cnt))
    d_starts = torch.from_numpy(starts).cuda()
    chunk = chunk if chunk > 0 else 4096
    d_x = torch.empty((chunk, 3, L), dtype=torch.float32, device="cuda")
    d_y = torch.empty((nwin, 3, L), dtype=torch.float32, device="cuda")
    d_ann = torch.empty((3, n), dtype=torch.float32, device="cuda")
    ws_bytes = int(lib.vp_forward_workspace_bytes(model._handle, chunk, precision))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    thr_on = np.array([max(float(v), 0.0) for v in thresholds], dtype=np.float32)
    thr_off = thr_on / np.float32(2)
    bounds = torch.empty(6, dtype=torch.int64, device="cuda")
    picks = torch.empty(65536 * 32, dtype=torch.uint8, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peaks = measured_peaks()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # tensor-core modes: vp_annotate runs K1 inside the first conv kernel (vp_slice_forward); the
    # stand-alone K1 is still timed (it is the stage-level API and the PhaseNet / fp32 path) but is not part of "forward"
    fused = precision != _lib.PRECISION["fp32"]
    keep_lo, keep_hi = argdict["blinding"][0], L - argdict["blinding"][1]
    acc = {"slice": 0.0, "forward": 0.0, "stack": 0.0, "pick": 0.0}
    for rep in range(reps + 1):
        t = {k: 0.0 for k in acc}
        evs = []
        for w0 in range(0, nwin, chunk):
            nw = min(chunk, nwin - w0)
            a, b, c = ev(), ev(), ev()
            a.record()
            _lib.check(lib.vp_slice_normalize(rec_dev.data_ptr(), 0, n, rec_dev.stride(0), d_starts.data_ptr() + 8 * w0, nw, L,
                                              0, 1 if kind == "eqtransformer" else 0, d_x.data_ptr(), stream))
            b.record()
            if fused:
                _lib.check(lib.vp_slice_forward(model._handle, rec_dev.data_ptr(), 0, n, rec_dev.stride(0), d_starts.data_ptr() + 8 * w0, nw,
                                                0, 1 if kind == "eqtransformer" else 0, d_y.data_ptr() + 4 * 3 * L * w0, ws.data_ptr(), ws_bytes, precision, keep_lo, keep_hi, stream))
            else:
                _lib.check(lib.vp_forward_range(model._handle, d_x.data_ptr(), nw, d_y.data_ptr() + 4 * 3 * L * w0, ws.data_ptr(), ws_bytes,
                                                precision, keep_lo, keep_hi, stream))
            c.record()
            evs.append((a, b, c))
        a, b, c = ev(), ev(), ev()
        a.record()
        _lib.check(lib.vp_stack(d_y.data_ptr(), d_starts.data_ptr(), nwin, L, 3, ov, argdict["blinding"][0], argdict["blinding"][1],
                                _lib.STACK[argdict["stacking"]], d_ann.data_ptr(), n, stream))
        b.record()
        count.zero_()
        _lib.check(lib.vp_pick_labels(d_ann.data_ptr(), 3, n, thr_on.ctypes.data, thr_off.ctypes.data, picks.data_ptr(), 65536,
                                      count.data_ptr(), bounds.data_ptr(), stream))
        c.record()
        torch.cuda.synchronize()
        for (x0, x1, x2) in evs:
            t["slice"] += x0.elapsed_time(x1)
            t["forward"] += x1.elapsed_time(x2)
        t["stack"] = a.elapsed_time(b)
        t["pick"] = b.elapsed_time(c)
        if rep > 0:
            for k in acc:
                acc[k] += t[k] / reps
    bytes_slice = 3 * n * 4 + nwin * 3 * L * 4
    bytes_stack = nwin * 3 * L * 4 + 3 * n * 4
    bytes_pick = 3 * n * 4  # one pass over all three labels: picks of the thresholded ones + the _trim_nan bounds of each
    out = {f"{k}_ms": v for k, v in acc.items()}
    out["forward_includes_slicing"] = bool(fused)  # True: K1 runs inside the first encoder kernel; slice_ms is the stand-alone K1
    for k, by in (("slice", bytes_slice), ("stack", bytes_stack), ("pick", bytes_pick)):
        gbs = by / (acc[k] / 1e3) / 1e9 if acc[k] > 0 else None
        out[f"{k}_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                           "frac": gbs / peaks["hbm"] if gbs else None, "algorithmic_bytes": by}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="eqtransformer", choices=["eqtransformer", "phasenet"])
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "f16x3", "bf16"],
                    help="f16x3 (default): tcgen05 on fp16 hi/lo split operands, fp32 accumulate -- the exact mode (<= 1e-4 vs the oracle); "
                         "fp32: CUDA-core FFMA; bf16: single-pass tensor cores (reported separately, looser tolerance)")
    ap.add_argument("--samples", type=int, default=N_DAY, help="samples per record (default: one station-day)")
    ap.add_argument("--chunk", type=int, default=0, help="windows per forward launch group (0: library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true",
                    help="only the two headline arms (device-resident value, end-to-end e2e): no blocking arms, per-kernel pass, "
                         "stage timings or CPU baseline -- for the long sharded runs of BASELINE.json configs[2] / [3]")
    ap.add_argument("--records", type=int, default=2,
                    help="distinct synthetic records per rank, cycled over the steps (long mode: >= 16)")
    ap.add_argument("--classify-stream", type=int, default=0,
                    help="also time picker.classify(stream) on a stream of this many station-days (the drop-in call)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE.json configs[4]: raw forward sweep B = 256 ... 8192")
    ap.add_argument("--sweep-precisions", default="f16x3,bf16")
    ap.add_argument("--sweep-one-model", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=0,
                    help="profiling aid (ncu): run this many device-resident steps and exit without timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep:
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
