"""hit_merge_clusters / hit_merged / hit_merged_components on the B200 (reference:
core/plugins/builtin/cpu/hit_merge.py:325-534).

With the default ``merge_gap_ns <= 0`` nothing is merged: every threshold hit is its own cluster,
but rows are regrouped by (board, channel) and ordered by absolute window start.  With
``merge_gap_ns > 0`` hits are chained greedily (gap and max-total-width rule); the device cuts the
ordered hits where a break is certain and one thread replays the reference loop inside each
piece (``wfb_hit_merge``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from .. import ops
from ..dtypes import HIT_MERGE_CLUSTERS_DTYPE, HIT_MERGED_COMPONENTS_DTYPE, HIT_MERGED_DTYPE
from ..plugin_api import Option, Plugin, check_dt_array, resolve_dt_config

_MERGE_OPTIONS = {
    "merge_gap_ns": Option(default=0.0, type=float, help="max edge gap (ns); <= 0 disables merging"),
    "max_total_width_ns": Option(default=10000.0, type=float, help="max total width of a merged chain (ns)"),
    "dt": Option(default=None, type=int, help="sample interval (ns), only used when hit_threshold has no dt field"),
}


def _merge_all(context: Any, plugin: Plugin, run_id: str):
    hits = context.get_data(run_id, "hit_threshold")
    if not isinstance(hits, np.ndarray):
        raise ValueError(f"{plugin.provides} expects hit_threshold as a single structured array")
    merge_plugin = plugin
    if "merge_gap_ns" not in plugin.options:  # hit_merged_components reads hit_merged's options
        getter = getattr(context, "get_plugin", None)
        merge_plugin = getter("hit_merged") if callable(getter) else B200HitMergePlugin()
    merge_gap_ns = float(context.get_config(merge_plugin, "merge_gap_ns"))
    max_total_width_ns = float(context.get_config(merge_plugin, "max_total_width_ns"))
    if len(hits):
        explicit_dt = resolve_dt_config(context, merge_plugin, deprecated_keys=("sampling_interval_ns", "dt_ns"))
        check_dt_array(hits, explicit_dt, plugin.provides, "hit_threshold[channel]")
    key = (run_id, "_b200_hit_merge", id(hits), merge_gap_ns, max_total_width_ns)
    cache = getattr(context, "_b200_cache", None)
    if cache is None:
        cache = {}
        try:
            context._b200_cache = cache
        except Exception:
            pass
    if key not in cache:
        cache.clear()
        cache[key] = ops.hit_merge(hits, merge_gap_ns=merge_gap_ns, max_total_width_ns=max_total_width_ns)
    return cache[key]


class B200HitMergeClustersPlugin(Plugin):
    provides = "hit_merge_clusters"
    depends_on = ["hit_threshold"]
    description = "Internal cluster membership rows shared by hit_merged outputs."
    version = "0.1.0"
    save_when = "always"
    output_dtype = HIT_MERGE_CLUSTERS_DTYPE
    options = dict(_MERGE_OPTIONS)

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        return _merge_all(context, self, run_id)[0]


class B200HitMergePlugin(Plugin):
    provides = "hit_merged"
    depends_on = ["hit_threshold", "hit_merge_clusters"]
    description = "Merge nearby threshold hits per channel with time-gap and max-width constraints."
    version = "0.8.0"
    save_when = "always"
    output_dtype = HIT_MERGED_DTYPE
    options = dict(_MERGE_OPTIONS)

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        return _merge_all(context, self, run_id)[1]


class B200HitMergedComponentsPlugin(Plugin):
    provides = "hit_merged_components"
    depends_on = ["hit_merge_clusters", "hit_merged"]
    description = "Return per-cluster component hit indices for hit_merged rows."
    version = "0.1.0"
    save_when = "always"
    output_dtype = HIT_MERGED_COMPONENTS_DTYPE

    def compute(self, context: Any, run_id: str, **_kwargs) -> np.ndarray:
        return _merge_all(context, self, run_id)[2]
