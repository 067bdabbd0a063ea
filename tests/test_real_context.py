"""GPU: the B200 plugins inside a REAL reference Context (core/context_execution.py:140-183).

Two fresh interpreters run tests/real_context_run.py - once with the reference's CPU plugins, once with
``ctx.register(*b200_default(), allow_override=True)`` - on the same V1725 ``.bin`` raw files, and every
result pulled with ``ctx.get_data`` (records ... hit_grouped) is compared.  The reference comes from
``baseline/_ref`` on the GPU box (tools/install_reference.sh; git-ignored, travels with the snapshot)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ATOL, RTOL
from refctx import reference_root

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "real_context_run.py")


def _run(mode, tmp_path, scenario="records"):
    out = os.path.join(str(tmp_path), f"{mode}_{scenario}.npz")
    res = subprocess.run([sys.executable, SCRIPT, mode, out, str(tmp_path), scenario], capture_output=True, text=True, timeout=900)
    assert "REAL_CONTEXT_DONE" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    return np.load(out, allow_pickle=False)


def _check_event_frames(cpu, gpu):
    """df_events / df_paired.  The reference orders the members of an event with np.argsort(channels) - quicksort, which is
    stable only up to 16 elements - on rows whose order among EQUAL timestamps comes from pandas' quicksort.  The raw files of
    this test hold both (ties in time, the same channel number on two boards), so for an event in which a channel occurs
    more than once the member order - and t_min / t_max / dt, which are the first / last member's time - is not defined by
    the reference.  Compared: every event's members as a multiset; order-dependent columns where the order is defined."""
    for frame in ("df_events", "df_paired"):
        for col in ("event_id", "n_hits", "channels.len", "channels.flat"):
            assert np.array_equal(cpu[f"{frame}.{col}"], gpu[f"{frame}.{col}"]), f"{frame}.{col}"
        lens = cpu[f"{frame}.channels.len"]
        off = np.concatenate([[0], np.cumsum(lens)])
        ch = cpu[f"{frame}.channels.flat"]
        unique = np.array([len(np.unique(ch[off[k]:off[k + 1]])) == lens[k] for k in range(len(lens))])
        assert unique.sum() > len(lens) // 3
        for col in ("areas", "heights", "timestamps"):
            w, g = cpu[f"{frame}.{col}.flat"], gpu[f"{frame}.{col}.flat"]
            for k in range(len(lens)):
                a, b = w[off[k]:off[k + 1]], g[off[k]:off[k + 1]]
                if unique[k]:
                    assert np.allclose(a, b, rtol=RTOL, atol=ATOL), (frame, col, k)
                else:  # same members, channel by channel
                    c = ch[off[k]:off[k + 1]]
                    for v in np.unique(c):
                        assert np.allclose(np.sort(a[c == v]), np.sort(b[c == v]), rtol=RTOL, atol=ATOL), (frame, col, k, int(v))
        for name in cpu.files:
            if not name.startswith(frame + ".") or ".flat" in name or ".len" in name or name.endswith((".event_id", ".n_hits")):
                continue
            w, g = cpu[name][unique], gpu[name][unique]
            assert np.allclose(g, w, rtol=RTOL, atol=ATOL, equal_nan=True), name


@pytest.mark.skipif(reference_root() is None, reason="reference package not installed (baseline/_ref)")
def test_records_route_through_a_real_context(tmp_path):
    cpu = _run("cpu", tmp_path)
    gpu = _run("b200", tmp_path)
    assert sorted(cpu.files) == sorted(gpu.files)
    # `df` is sorted by timestamp with pandas' default (unstable) quicksort in the reference and with a stable device sort
    # here: rows of EQUAL timestamps (the raw files contain ties on purpose) may come in another order, so the columns of
    # `df` are compared after a canonical (timestamp, record_id) order
    df_perm = {tag: np.lexsort((d["df.record_id"], d["df.timestamp"])) for tag, d in (("cpu", cpu), ("gpu", gpu))}
    _check_event_frames(cpu, gpu)
    for name in cpu.files:
        if name.startswith(("df_events.", "df_paired.")):
            continue  # compared event by event above
        want, got = cpu[name], gpu[name]
        if name.startswith("df."):
            want, got = want[df_perm["cpu"]], got[df_perm["gpu"]]
        assert want.shape == got.shape and want.dtype == got.dtype, (name, want.shape, got.shape, want.dtype, got.dtype)
        if want.dtype.names:
            for f in want.dtype.names:
                w, g = want[f], got[f]
                if w.dtype.kind == "f":
                    assert np.allclose(g, w, rtol=RTOL, atol=ATOL, equal_nan=True), f"{name}.{f}"
                else:
                    assert np.array_equal(g, w), f"{name}.{f}"
        elif want.dtype.kind == "f":
            assert np.allclose(got, want, rtol=RTOL, atol=ATOL, equal_nan=True), name
        else:
            assert np.array_equal(got, want), name
    assert len(cpu["hit_threshold"]) > 1000 and len(cpu["hit_merged"]) < len(cpu["hit_threshold"])



def _compare(cpu, gpu, skip=(), loose=None):
    """``loose``: {(name, field): (row mask, atol)} - rows compared with a wider absolute tolerance."""
    assert sorted(cpu.files) == sorted(gpu.files)
    for name in cpu.files:
        if name.startswith(skip):
            continue
        want, got = cpu[name], gpu[name]
        assert want.shape == got.shape and want.dtype == got.dtype, (name, want.shape, got.shape, want.dtype, got.dtype)
        for f in want.dtype.names or (None,):
            w, g = (want, got) if f is None else (want[f], got[f])
            if w.dtype.kind == "f":
                atol = np.full(w.shape, ATOL)
                if loose and (name, f) in loose:
                    mask, wide = loose[(name, f)]
                    atol[mask] = wide
                assert np.all(np.abs(g - w) <= atol + RTOL * np.abs(w)) or np.allclose(g, w, rtol=RTOL, atol=ATOL, equal_nan=True), f"{name}.{f}"
            else:
                assert np.array_equal(g, w), f"{name}.{f}"


@pytest.mark.skipif(reference_root() is None, reason="reference package not installed (baseline/_ref)")
def test_structured_wave_sources_through_a_real_context(tmp_path):
    """The same raw files with the structured rows as the wave source: basic_features on st_waveforms, hit_threshold and
    waveform_width_integral on filtered_waveforms (per-channel threshold + fixed baseline from channel_config), `hit` on
    records + wave_pool_filtered, waveform_width behind it."""
    cpu = _run("cpu", tmp_path, "waves")
    gpu = _run("b200", tmp_path, "waves")
    # `hit` reads the Savitzky-Golay filtered pool: scipy fits the edge polynomial of float32 input in float32, which leaves
    # ~6 ulp (here up to 4e-3 ADC) of noise on the first / last samples of a record; a peak height taken there (max - min of
    # the window) carries twice that.  Everything else - positions, edges, widths, the other rows - meets the normal bar.
    pos = cpu["hit"]["position"]
    at_edge = (pos < 16) | (pos >= 300 - 16)  # 5 edge samples + the height window + the derivative's shift
    _compare(cpu, gpu, loose={("hit", "height"): (at_edge, 0.02)})
    assert at_edge.sum() < len(pos) // 10
    assert len(cpu["hit_threshold"]) > 1000 and len(cpu["hit"]) > 1000


@pytest.mark.skipif(reference_root() is None, reason="reference package not installed (baseline/_ref)")
def test_vx2730_csv_route_through_a_real_context(tmp_path):
    """VX2730 CSV files (the reference's default adapter): the reference's readers parse the text, B200WaveformsPlugin
    structures the rows on the device (dual baseline), B200RecordsPlugin builds records + wave_pool (device sort with ties
    across channels), then features, hits, merge and grouping."""
    cpu = _run("cpu", tmp_path, "csv")
    gpu = _run("b200", tmp_path, "csv")
    _compare(cpu, gpu)
    assert len(cpu["records"]) == 640 and len(cpu["hit_threshold"]) > 500


@pytest.mark.skipif(reference_root() is None, reason="reference package not installed (baseline/_ref)")
def test_polarity_metadata_and_mixed_filters_through_a_real_context(tmp_path):
    """CSV route with channel metadata (one negative, one positive channel: the float32 signal branch of the features and
    the polarity of the hits), a Butterworth override for one channel next to the default Savitzky-Golay filter, hits on
    the filtered pool, charge widths on records."""
    cpu = _run("cpu", tmp_path, "mixed")
    gpu = _run("b200", tmp_path, "mixed")
    assert set(np.unique(cpu["records"]["polarity"])) == {"negative", "positive", "unknown"}
    _compare(cpu, gpu)
