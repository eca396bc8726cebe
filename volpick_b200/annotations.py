"""Result containers with the field surface of ``seisbench.util.annotations``.

The reference consumes ``classify(...).picks`` as an iterable of objects with the fields
``trace_id, start_time, end_time, peak_time, peak_value, phase`` (/root/reference/README.md:66-84).
"""
from __future__ import annotations

import fnmatch
from typing import Iterable, List, Optional


class Pick:
    """A phase pick.  ``trace_id`` is ``"NET.STA.LOC"``; times are UTCDateTime-like."""

    def __init__(self, trace_id, start_time, end_time=None, peak_time=None, peak_value=None, phase=None):
        self.trace_id = trace_id
        self.start_time = start_time
        self.end_time = end_time
        self.peak_time = peak_time
        self.peak_value = peak_value
        self.phase = phase
        if end_time is not None and peak_time is not None:
            if not start_time <= peak_time <= end_time:
                raise ValueError("peak_time must be between start_time and end_time.")

    @classmethod
    def _trusted(cls, trace_id, start_time, end_time, peak_time, peak_value, phase):
        """Constructor without the ordering check, for callers that checked start <= peak <= end on whole arrays."""
        p = cls.__new__(cls)
        p.trace_id, p.start_time, p.end_time, p.peak_time, p.peak_value, p.phase = trace_id, start_time, end_time, peak_time, peak_value, phase
        return p

    def _key(self):
        end = self.end_time if self.end_time is not None else self.start_time
        return (self.start_time.ns if hasattr(self.start_time, "ns") else self.start_time,
                end.ns if hasattr(end, "ns") else end, str(self.trace_id), str(self.phase))

    def __lt__(self, other) -> bool:
        return self._key() < other._key()

    def __eq__(self, other) -> bool:
        return isinstance(other, Pick) and self._key() == other._key() and self.peak_time == other.peak_time and self.peak_value == other.peak_value

    def __hash__(self):
        return hash(self._key())

    def __str__(self) -> str:
        parts = [self.trace_id]
        if self.peak_time is None:
            parts.append(str(self.start_time) if self.end_time is None else f"{self.start_time}-{self.end_time}")
        else:
            parts.append(str(self.peak_time))
        if self.phase is not None:
            parts.append(str(self.phase))
        return "\t".join(parts)

    __repr__ = __str__


class Detection:
    """An event detection (EQTransformer ``Detection`` trace)."""

    def __init__(self, trace_id, start_time, end_time, peak_value=None):
        self.trace_id = trace_id
        self.start_time = start_time
        self.end_time = end_time
        self.peak_value = peak_value

    def _key(self):
        return (self.start_time.ns if hasattr(self.start_time, "ns") else self.start_time,
                self.end_time.ns if hasattr(self.end_time, "ns") else self.end_time, str(self.trace_id))

    def __lt__(self, other) -> bool:
        return self._key() < other._key()

    def __eq__(self, other) -> bool:
        return isinstance(other, Detection) and self._key() == other._key() and self.peak_value == other.peak_value

    def __hash__(self):
        return hash(self._key())

    def __str__(self) -> str:
        return f"{self.trace_id}\t{self.start_time}\t{self.end_time}"

    __repr__ = __str__


class PickList(list):
    """List of picks with the convenience API of ``seisbench.util.PickList``."""

    def __str__(self) -> str:
        return f"PickList with {len(self)} entries:\n\n" + self._rep_entries()

    def _rep_entries(self) -> str:
        if len(self) <= 6:
            return "\n".join(str(x) for x in self)
        return "\n".join([str(x) for x in self[:3]] + ["..."] + [str(x) for x in self[-3:]])

    __repr__ = __str__

    def select(self, trace_id: Optional[str] = None, min_confidence: Optional[float] = None, phase: Optional[str] = None) -> "PickList":
        def ok(p):
            if trace_id is not None and not fnmatch.fnmatch(p.trace_id, trace_id):
                return False
            if min_confidence is not None and (p.peak_value is None or p.peak_value < min_confidence):
                return False
            if phase is not None and p.phase != phase:
                return False
            return True

        return PickList(p for p in self if ok(p))

    def to_dataframe(self):
        """The ``picklist2df`` helper of /root/reference/README.md:69-84."""
        import pandas as pd

        return pd.DataFrame([
            {"trace_id": p.trace_id, "start_time": p.start_time, "end_time": p.end_time,
             "peak_time": p.peak_time, "peak_value": p.peak_value, "phase": p.phase}
            for p in self
        ])


class DetectionList(PickList):
    def __str__(self) -> str:
        return f"DetectionList with {len(self)} entries:\n\n" + self._rep_entries()

    __repr__ = __str__

    def to_dataframe(self):
        import pandas as pd

        return pd.DataFrame([
            {"trace_id": d.trace_id, "start_time": d.start_time, "end_time": d.end_time, "peak_value": d.peak_value}
            for d in self
        ])


class ClassifyOutput:
    """``seisbench.util.ClassifyOutput``: attribute bag with a ``creator``."""

    def __init__(self, creator: str, **kwargs):
        self.creator = creator
        self._keys = list(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __str__(self) -> str:
        return f"ClassifyOutput from {self.creator} with: " + ", ".join(self._keys)

    __repr__ = __str__
