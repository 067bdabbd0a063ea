"""Per-(board, channel) plugin overrides: host-side mirror of the reference's layered
``channel_config`` resolution (core/hardware/channel.py:212-313, 412-431).

Layers, later wins: ``defaults`` -> matching ``groups`` (``config`` mapping) -> ``channels``
(or bare top-level channel keys).  Keys are HardwareChannel-like objects, ``(board, channel)``
pairs or ``"board:channel"`` strings; anything else raises ValueError("Invalid channel key ...").
A top-level block keyed by run_id selects the run's sub-tree first.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Any

import numpy as np


def parse_channel_ref(key: Any) -> tuple[int, int] | None:
    if hasattr(key, "board") and hasattr(key, "channel"):
        return int(key.board), int(key.channel)
    if isinstance(key, (tuple, list)) and len(key) == 2:
        try:
            return int(key[0]), int(key[1])
        except (TypeError, ValueError):
            return None
    if isinstance(key, str) and ":" in key:
        left, right = key.strip().split(":", 1)
        try:
            return int(left.strip()), int(right.strip())
        except (TypeError, ValueError):
            return None
    return None


def _groups(groups: Any) -> list[Mapping]:
    if isinstance(groups, Mapping):
        out = []
        for name, g in groups.items():
            if isinstance(g, Mapping):
                out.append(g if "name" in g else {"name": str(name), **g})
        return out
    if isinstance(groups, Sequence) and not isinstance(groups, (str, bytes)):
        return [g for g in groups if isinstance(g, Mapping)]
    return []


def _in_selector(channel: tuple[int, int], selectors: Any) -> bool:
    if not isinstance(selectors, Sequence) or isinstance(selectors, (str, bytes)):
        return False
    return any(parse_channel_ref(item) == channel for item in selectors)


def resolve_channel_values(channel_config: Any, run_id: str, board: int, channel: int, base_values: Mapping | None = None) -> dict:
    """Effective option values of one hardware channel (resolve_effective_channel_config)."""
    resolved: dict = dict(base_values or {})
    if not isinstance(channel_config, Mapping):
        return resolved
    block = channel_config
    run_block = block.get(run_id)
    if isinstance(run_block, Mapping):
        block = run_block
    key = (int(board), int(channel))
    defaults = block.get("defaults")
    if isinstance(defaults, Mapping):
        resolved.update(defaults)
    for g in _groups(block.get("groups")):
        if not _in_selector(key, g.get("channels")):
            continue
        values = g.get("config")
        if isinstance(values, Mapping):
            resolved.update(values)
    channels_block = block.get("channels")
    if not isinstance(channels_block, Mapping):
        channels_block = block
    for k, values in channels_block.items():
        if isinstance(k, str) and k in {"defaults", "groups", "channels"}:
            continue
        parsed = parse_channel_ref(k)
        if parsed is None:
            raise ValueError(f'Invalid channel key {k!r}; expected HardwareChannel, (board, channel), or "board:channel".')
        if parsed != key:
            continue
        if not isinstance(values, Mapping):
            raise ValueError(f"Invalid channel config for {k!r}; expected a mapping, got {type(values).__name__}.")
        resolved.update(values)
        break
    return resolved


def unique_channels(boards: np.ndarray, channels: np.ndarray) -> list[tuple[int, int]]:
    if len(boards) == 0:
        return []
    keys = np.asarray(boards, dtype=np.int64) * 65536 + (np.asarray(channels, dtype=np.int64) & 0xFFFF)
    out = []
    for k in np.unique(keys).tolist():
        out.append((int(k >> 16), ((int(k) & 0xFFFF) ^ 0x8000) - 0x8000))  # sign-extend the 16-bit channel
    return sorted(out)


def per_channel_option(channel_config: Any, run_id: str, boards: np.ndarray, channels: np.ndarray, name: str, base_value: Any) -> dict:
    """{(board, channel): effective value of option ``name``} for every channel present."""
    if not channel_config:  # no overrides configured: every channel takes the base value (and nothing needs the channel list)
        return {}
    out = {}
    for b, c in unique_channels(boards, channels):
        out[(b, c)] = resolve_channel_values(channel_config, run_id, b, c, {name: base_value}).get(name, base_value)
    return out
