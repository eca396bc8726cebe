"""volpick_b200 -- B200-native continuous-waveform picking with the volpick weights.

A from-scratch sm_100a implementation of the path the reference (zhong-yy/volpick) runs through
SeisBench: ``Model.from_pretrained("volpick")`` -> ``annotate`` / ``classify`` -> ``PickList``
(/root/reference/README.md:36-84).  Python holds the API surface and the stream bookkeeping;
all arithmetic is hand-written CUDA behind the C ABI in ``include/volpick_b200.h``.
"""
from .annotations import ClassifyOutput, Detection, DetectionList, Pick, PickList
from .models import EQTransformer, PhaseNet, WaveformModel
from .stream import Stream, Trace, UTCDateTime

__all__ = [
    "EQTransformer", "PhaseNet", "WaveformModel",
    "Pick", "Detection", "PickList", "DetectionList", "ClassifyOutput",
    "Stream", "Trace", "UTCDateTime",
]
__version__ = "0.1.0"
