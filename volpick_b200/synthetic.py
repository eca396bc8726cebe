"""Seeded synthetic 3-component 100 Hz records (the recipe of SURVEY.md section 8(d)).

Per station ``i``: Gaussian noise (sigma = 100 counts) on Z, N, E plus damped-sinusoid events at
Poisson times (mean spacing 120 s): P on Z ``A sin(2 pi 8 tau) exp(-tau / 1.0)``, S on N and E
``2 A sin(2 pi 4 tau) exp(-tau / 2.0)`` starting U(3, 10) s later, ``A = 100 U(3, 30)``.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .stream import Stream, Trace, UTCDateTime

SAMPLING_RATE = 100.0
STATION_HOUR = 360_000
STATION_DAY = 8_640_000


def synthetic_record(station: int, n_samples: int, return_events: bool = False):
    """(3, n_samples) float32, component order Z, N, E; deterministic in ``station``."""
    rng = np.random.default_rng(1000 + station)
    x = rng.standard_normal((3, n_samples)).astype(np.float32) * np.float32(100.0)
    events: List[Tuple[int, int]] = []
    t = 0.0
    dur_p, dur_s = int(8 * SAMPLING_RATE), int(16 * SAMPLING_RATE)
    tau_p = np.arange(dur_p) / SAMPLING_RATE
    tau_s = np.arange(dur_s) / SAMPLING_RATE
    while True:
        t += rng.exponential(120.0)
        ip = int(t * SAMPLING_RATE)
        if ip >= n_samples:
            break
        sp = rng.uniform(3.0, 10.0)
        amp = 100.0 * rng.uniform(3.0, 30.0)
        i_s = ip + int(sp * SAMPLING_RATE)
        p_wave = (amp * np.sin(2 * np.pi * 8 * tau_p) * np.exp(-tau_p / 1.0)).astype(np.float32)
        s_wave = (2 * amp * np.sin(2 * np.pi * 4 * tau_s) * np.exp(-tau_s / 2.0)).astype(np.float32)
        m = min(dur_p, n_samples - ip)
        x[0, ip : ip + m] += p_wave[:m]
        if i_s < n_samples:
            m = min(dur_s, n_samples - i_s)
            x[1, i_s : i_s + m] += s_wave[:m]
            x[2, i_s : i_s + m] += s_wave[:m]
        events.append((ip, i_s))
    return (x, events) if return_events else x


def station_start(station: int) -> UTCDateTime:
    return UTCDateTime("2020-01-01T00:00:00") + 86400.0 * station


def synthetic_stream(station: int, n_samples: int) -> Stream:
    """The same record as a 3-trace stream ``XX.S{station:04d}..HH{Z,N,E}`` starting 2020-01-01 + station days."""
    x = synthetic_record(station, n_samples)
    t0 = station_start(station)
    return Stream(
        Trace(x[i], {"network": "XX", "station": f"S{station:04d}", "location": "", "channel": "HH" + c,
                     "starttime": t0, "sampling_rate": SAMPLING_RATE})
        for i, c in enumerate("ZNE")
    )
