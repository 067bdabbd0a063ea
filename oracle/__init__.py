"""CPU oracle for the WaveformAnalysis per-record hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy, plus a small C
restatement under ``oracle/c``) of the reference algorithms that the CUDA kernels in
``waveformanalysis_b200/csrc`` replace.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it - and there only
as the checker or the timed CPU baseline, never as the thing shipped.  Nothing under
``waveformanalysis_b200/`` imports ``oracle``.

Parity pinning: every function here is checked (a) against the LIVE reference imported
from ``/root/reference`` in the build container (``tests/golden/make_golden.py`` runs the
reference plugins and writes ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py``
re-runs the comparison whenever ``/root/reference`` is present) and (b) against the
known-answer vectors of the reference's own tests (``tests/test_oracle_known_answers.py``,
citing the reference test file:line for each vector).
"""
