"""The drop-in boundary: the reference's ``Plugin`` / ``Option`` classes.

When the reference package (``waveform_analysis``) is importable the B200 plugins subclass ITS
``Plugin`` and use ITS ``Option`` (core/plugins/core/base.py:37-275, 320-663), so they register in
a real ``Context`` with ``ctx.register(plugin, allow_override=True)`` and take part in lineage /
cache keys.  When it is not importable (the GPU test box), a minimal mirror with the same
attributes keeps the plugins usable with any object that offers ``get_config`` / ``get_data``.
"""

from __future__ import annotations

from typing import Any

try:  # pragma: no cover - depends on the environment
    from waveform_analysis.core.plugins.core.base import Option, Plugin  # type: ignore

    HAVE_REFERENCE = True
except Exception:  # reference not installed: same attribute surface, nothing else
    HAVE_REFERENCE = False

    class Option:  # type: ignore[no-redef]
        """Mirror of waveform_analysis.core.plugins.core.base.Option (base.py:37-120)."""

        def __init__(self, default: Any = None, type: Any = None, help: str = "", validate=None, track: bool = True,
                     unit=None, internal_unit=None, choices=None, min_value=None, max_value=None, deprecated: bool = False,
                     deprecated_message: str = "", alias=None):
            self.default = default
            self.type = type
            self.help = help
            self.validate = validate
            self.track = track
            self.unit = unit
            self.internal_unit = internal_unit
            self.choices = choices
            self.min_value = min_value
            self.max_value = max_value
            self.deprecated = deprecated
            self.deprecated_message = deprecated_message
            self.alias = alias

    class Plugin:  # type: ignore[no-redef]
        """Mirror of waveform_analysis.core.plugins.core.base.Plugin (base.py:320-344, 407-416, 590-613)."""

        provides: str = ""
        depends_on: list = []
        options: dict = {}
        save_when: str = "never"
        output_dtype = None
        input_dtype: dict = {}
        output_kind = "static"
        description: str = ""
        version: str = "0.0.0"
        is_side_effect: bool = False
        uses_run_config: bool = False
        timeout = None

        def __init_subclass__(cls, **kwargs):
            super().__init_subclass__(**kwargs)
            merged: dict = {}
            for base in reversed(cls.__mro__):
                if isinstance(getattr(base, "options", None), dict):
                    merged.update(base.options)
            cls.options = merged

        @property
        def config_keys(self):
            return list(self.options.keys())

        def resolve_depends_on(self, context: Any, run_id: str | None = None) -> list:
            return list(self.depends_on) if self.depends_on else []

        def compute(self, context: Any, run_id: str, **kwargs):  # pragma: no cover
            raise NotImplementedError

        def on_error(self, context: Any, exception: Exception):
            pass

        def cleanup(self, context: Any):
            pass


def get_raw_config_value(context: Any, plugin: Any, name: str) -> Any:
    """Config lookup that does not require ``name`` in plugin.options
    (core/plugins/builtin/cpu/_dt_compat.py:12-24)."""
    provides = plugin.provides
    cfg = getattr(context, "config", {}) or {}
    if provides in cfg and isinstance(cfg[provides], dict) and name in cfg[provides]:
        return cfg[provides][name]
    if f"{provides}.{name}" in cfg:
        return cfg[f"{provides}.{name}"]
    return cfg.get(name)


def resolve_dt_config(context: Any, plugin: Any, deprecated_keys=()) -> Any:
    """``dt`` with fallback to deprecated keys (+ DeprecationWarning) (_dt_compat.py:27-48)."""
    import warnings

    dt = get_raw_config_value(context, plugin, "dt")
    if dt is not None:
        return dt
    for old in deprecated_keys:
        v = get_raw_config_value(context, plugin, old)
        if v is None:
            continue
        warnings.warn(f"[{plugin.provides}] Config '{old}' is deprecated and will be removed in a future release. Use 'dt' instead.",
                      DeprecationWarning, stacklevel=3)
        return v
    return None


def check_dt_array(data, explicit_dt, plugin_name: str, data_name: str) -> int | None:
    """Validation part of require_dt_array (_dt_compat.py:51-81).  Returns the scalar dt to use when
    the data has no ``dt`` field, None when it has one (and it is valid)."""
    import numpy as np

    names = data.dtype.names or ()
    if "dt" in names:
        if len(data):
            dt = np.ascontiguousarray(data["dt"])  # one pass over the strided field
            if int(dt.min()) <= 0:
                raise ValueError(f"[{plugin_name}] {data_name}.dt must be positive for every row")
            if int(dt.max()) > np.iinfo(np.int32).max:
                raise ValueError(f"[{plugin_name}] {data_name}.dt exceeds int32 range")
        return None
    if explicit_dt is None:
        raise ValueError(f"[{plugin_name}] Input '{data_name}' is missing required field 'dt'; "
                         "provide explicit config 'dt' for this migration period.")
    dt_scalar = int(explicit_dt)
    if dt_scalar <= 0:
        raise ValueError(f"[{plugin_name}] dt must be > 0")
    if dt_scalar > np.iinfo(np.int32).max:
        raise ValueError(f"[{plugin_name}] dt exceeds int32 range: {dt_scalar}")
    return dt_scalar
