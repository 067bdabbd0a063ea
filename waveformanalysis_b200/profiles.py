"""Plugin sets mirroring the reference's ``profiles.cpu_default()`` (core/plugins/profiles.py:19-29)
for the rows of the hot path: ``ctx.register(*b200_default(), allow_override=True)``."""

from __future__ import annotations


def b200_default() -> list:
    from . import plugins as P

    return [
        P.B200WaveformsPlugin(), P.B200RecordsPlugin(), P.B200WavePoolPlugin(), P.B200WavePoolFilteredPlugin(), P.B200BasicFeaturesPlugin(),
        P.B200ThresholdHitPlugin(), P.B200HitFinderPlugin(), P.B200WaveformWidthPlugin(), P.B200WaveformWidthIntegralPlugin(),
        P.B200HitMergeClustersPlugin(), P.B200HitMergePlugin(), P.B200HitMergedComponentsPlugin(),
        P.B200HitGroupedPlugin(), P.B200DataFramePlugin(), P.B200GroupedEventsPlugin(), P.B200PairedEventsPlugin(),
        P.B200S1S2ClassifierPlugin(),
    ]


def b200_hot_path() -> list:
    """Only the per-record plugins (no records builder, no grouping)."""
    from . import plugins as P

    return [P.B200WavePoolFilteredPlugin(), P.B200BasicFeaturesPlugin(), P.B200ThresholdHitPlugin(), P.B200HitFinderPlugin(),
            P.B200WaveformWidthPlugin(), P.B200WaveformWidthIntegralPlugin()]
