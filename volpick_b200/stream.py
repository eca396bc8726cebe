"""Minimal ObsPy-compatible waveform containers.

The reference feeds ``obspy.Stream`` objects to SeisBench (/root/reference/README.md:43-55,
/root/reference/Final_models/demo.ipynb cell 12).  ObsPy is not part of this image, so this module
provides duck-typed stand-ins with the attribute surface the picking path touches
(``trace.data``, ``trace.stats.{network,station,location,channel,starttime,sampling_rate,npts,delta}``,
``trace.id``, ``trace.times()``, ``stream.copy()``, ``stream.merge(-1)``, iteration, ``len``).  Real
ObsPy streams are accepted everywhere these are: the host code only relies on that surface.
"""
from __future__ import annotations

import copy as _copy
import datetime as _dt
from typing import Iterable, Iterator, List, Optional, Union

import numpy as np

_EPOCH = _dt.datetime(1970, 1, 1, tzinfo=_dt.timezone.utc)


class UTCDateTime:
    """UTC time stamp with integer-nanosecond resolution (like ``obspy.UTCDateTime``)."""

    __slots__ = ("ns",)

    def __init__(self, value: Union[str, float, int, "UTCDateTime", _dt.datetime, None] = 0, *, ns: Optional[int] = None):
        if ns is not None:
            self.ns = int(ns)
        elif isinstance(value, UTCDateTime):
            self.ns = value.ns
        elif hasattr(value, "ns") and not isinstance(value, (int, float)):  # obspy.UTCDateTime
            self.ns = int(value.ns)
        elif isinstance(value, str):
            s = value.strip().rstrip("Z").replace(" ", "T")
            if "T" not in s:
                s += "T00:00:00"
            date, _, tm = s.partition("T")
            frac = "0"
            if "." in tm:
                tm, frac = tm.split(".")
            parts = tm.split(":") + ["0", "0"]
            y, mo, d = (int(v) for v in date.split("-"))
            base = _dt.datetime(y, mo, d, int(parts[0]), int(parts[1]), int(parts[2]), tzinfo=_dt.timezone.utc)
            whole = int((base - _EPOCH).total_seconds())
            self.ns = whole * 10**9 + int((frac + "0" * 9)[:9])
        elif isinstance(value, _dt.datetime):
            if value.tzinfo is None:
                value = value.replace(tzinfo=_dt.timezone.utc)
            delta = value - _EPOCH
            self.ns = (delta.days * 86400 + delta.seconds) * 10**9 + delta.microseconds * 1000
        else:
            self.ns = int(round(float(value) * 1e9))

    @classmethod
    def _from_ns(cls, ns: int) -> "UTCDateTime":
        """Straight from integer nanoseconds, without __init__'s type dispatch (the pick lists of a station-day hold ~4,500
        time stamps)."""
        t = cls.__new__(cls)
        t.ns = ns
        return t

    # arithmetic -----------------------------------------------------------------------------
    def __add__(self, seconds: float) -> "UTCDateTime":
        return UTCDateTime(ns=self.ns + int(round(float(seconds) * 1e9)))

    def __sub__(self, other):
        if isinstance(other, UTCDateTime) or hasattr(other, "ns"):
            return (self.ns - int(other.ns)) / 1e9
        return UTCDateTime(ns=self.ns - int(round(float(other) * 1e9)))

    @property
    def timestamp(self) -> float:
        return self.ns / 1e9

    # comparisons ----------------------------------------------------------------------------
    def _key(self, other) -> int:
        return int(other.ns) if hasattr(other, "ns") else int(round(float(other) * 1e9))

    def __eq__(self, other) -> bool:
        try:
            return self.ns == self._key(other)
        except (TypeError, ValueError):
            return NotImplemented

    def __lt__(self, other) -> bool:
        return self.ns < self._key(other)

    def __le__(self, other) -> bool:
        return self.ns <= self._key(other)

    def __gt__(self, other) -> bool:
        return self.ns > self._key(other)

    def __ge__(self, other) -> bool:
        return self.ns >= self._key(other)

    def __hash__(self) -> int:
        return hash(self.ns)

    # formatting -----------------------------------------------------------------------------
    @property
    def datetime(self) -> _dt.datetime:
        return (_EPOCH + _dt.timedelta(seconds=self.ns // 10**9, microseconds=(self.ns % 10**9) // 1000)).replace(tzinfo=None)

    def isoformat(self) -> str:
        sec, rem = divmod(self.ns, 10**9)
        base = _EPOCH + _dt.timedelta(seconds=sec)
        return base.strftime("%Y-%m-%dT%H:%M:%S") + f".{rem // 1000:06d}Z"

    def __str__(self) -> str:
        return self.isoformat()

    def __repr__(self) -> str:
        return f"UTCDateTime({self.isoformat()!r})"


class Stats:
    """``obspy.core.trace.Stats`` subset; ``npts``/``delta``/``endtime`` are derived."""

    def __init__(self, header: Optional[dict] = None):
        self.network = ""
        self.station = ""
        self.location = ""
        self.channel = ""
        self.starttime = UTCDateTime(0)
        self.sampling_rate = 1.0
        self.npts = 0
        for k, v in (header or {}).items():
            if k == "starttime":
                v = UTCDateTime(v)
            if k == "delta":
                self.sampling_rate = 1.0 / float(v)
                continue
            setattr(self, k, v)

    @property
    def delta(self) -> float:
        return 1.0 / float(self.sampling_rate)

    @property
    def endtime(self) -> UTCDateTime:
        return self.starttime + max(self.npts - 1, 0) * self.delta

    def __repr__(self) -> str:
        return (f"Stats({self.network}.{self.station}.{self.location}.{self.channel} | {self.starttime} | "
                f"{self.sampling_rate} Hz, {self.npts} samples)")


class Trace:
    def __init__(self, data: Optional[np.ndarray] = None, header: Optional[dict] = None):
        self.data = np.asarray(data if data is not None else np.zeros(0, dtype=np.float32))
        self.stats = header if isinstance(header, Stats) else Stats(header)
        self.stats.npts = len(self.data)

    @property
    def id(self) -> str:
        s = self.stats
        return f"{s.network}.{s.station}.{s.location}.{s.channel}"

    def times(self) -> np.ndarray:
        return np.arange(self.stats.npts) / self.stats.sampling_rate

    def copy(self) -> "Trace":
        return _copy.deepcopy(self)

    def __len__(self) -> int:
        return len(self.data)

    def __repr__(self) -> str:
        s = self.stats
        return f"{self.id} | {s.starttime} - {s.endtime} | {s.sampling_rate} Hz, {s.npts} samples"


class Stream:
    def __init__(self, traces: Optional[Iterable[Trace]] = None):
        self.traces: List[Trace] = list(traces) if traces is not None else []

    def __iter__(self) -> Iterator[Trace]:
        return iter(self.traces)

    def __len__(self) -> int:
        return len(self.traces)

    def __getitem__(self, i):
        return Stream(self.traces[i]) if isinstance(i, slice) else self.traces[i]

    def __add__(self, other: "Stream") -> "Stream":
        return Stream(self.traces + list(other))

    def append(self, tr: Trace) -> "Stream":
        self.traces.append(tr)
        return self

    def copy(self) -> "Stream":
        return _copy.deepcopy(self)

    def select(self, network=None, station=None, location=None, channel=None) -> "Stream":
        import fnmatch

        def ok(val, pat):
            return pat is None or fnmatch.fnmatch(val, pat)

        return Stream(
            t for t in self.traces
            if ok(t.stats.network, network) and ok(t.stats.station, station)
            and ok(t.stats.location, location) and ok(t.stats.channel, channel)
        )

    def merge(self, method: int = -1) -> "Stream":
        """``method=-1`` clean-up only (what SeisBench calls): join traces of the same id that are
        exactly contiguous in time and share sampling rate and dtype; drop exact duplicates."""
        if method != -1:
            raise NotImplementedError("only Stream.merge(-1) is provided")
        by_id = {}
        for tr in self.traces:
            by_id.setdefault((tr.id, float(tr.stats.sampling_rate), tr.data.dtype.str), []).append(tr)
        merged: List[Trace] = []
        for (_, rate, _), trs in by_id.items():
            trs.sort(key=lambda t: t.stats.starttime.ns)
            cur = trs[0]
            for nxt in trs[1:]:
                expected = cur.stats.starttime.ns + int(round(cur.stats.npts / rate * 1e9))
                if nxt.stats.starttime.ns == cur.stats.starttime.ns and nxt.stats.npts == cur.stats.npts and np.array_equal(nxt.data, cur.data):
                    continue
                if abs(nxt.stats.starttime.ns - expected) <= 0.01 / rate * 1e9 and cur.stats.npts > 0:
                    cur = Trace(np.concatenate([cur.data, nxt.data]), _copy.deepcopy(cur.stats))
                else:
                    merged.append(cur)
                    cur = nxt
            merged.append(cur)
        self.traces = merged
        return self

    def __repr__(self) -> str:
        return f"{len(self.traces)} Trace(s) in Stream:\n" + "\n".join(repr(t) for t in self.traces)
