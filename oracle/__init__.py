"""CPU oracle for the volpick continuous-waveform picking path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``volpick_b200/`` imports this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline -- never as the thing shipped.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
SeisBench (``seisbench>=0.4.0``, un-pinned in the reference: no ``dependencies``
key in /root/reference/pyproject.toml:5-12, lower bound
``"seisbench_requirement": "0.4.0"`` in
/root/reference/Final_models/volpick/*/volpick.json.v1:7) plus ObsPy's
``trigger_onset`` and PyTorch.  Neither SeisBench nor ObsPy is vendored under
/root/reference or installable here (no network), and the reference has no tests
or golden vectors for the path.  This oracle therefore
restates the published SeisBench algorithm (``seisbench/models/base.py``,
``eqtransformer.py``, ``phasenet.py``, ``util/annotations.py``) and ObsPy's
``obspy/signal/trigger.py::trigger_onset`` and anchors on what the reference
itself pins:

* the call sites            /root/reference/README.md:46-66,
                            /root/reference/Final_models/demo.ipynb cells 7-15
* the state-dict shapes     /root/reference/Final_models/volpick/*/volpick.pt.v1
* window lengths + norm     /root/reference/volpick/model/models.py:445-452,849-856
* forward signature/labels  /root/reference/volpick/model/eval_taks0.py:68-72,85-89
* the pick rule             /root/reference/volpick/model/eval_taks0.py:46-56
* default thresholds        /root/reference/Final_models/volpick/*/volpick.json.v1:9-13

Modules
-------
nets      fp32 PyTorch-CPU forwards of EQTransformer / PhaseNet driven straight
          from a state dict.
pipeline  NumPy restatement of annotate/classify:
          window starts, pre-normalisation, blinding, NaN-buffer stacking,
          NaN trim, hysteresis picks.
c/        plain-C restatement of the integer / index pieces (window starts,
          stacking slot math with NumPy's summation order, trigger_onset runs),
          compiled by ``__graft_entry__.build()`` into ``oracle/_build/``.
"""
