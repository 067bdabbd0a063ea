"""B200 drop-in plugins: same ``provides`` names, options, versions and output dtypes as the
reference's CPU plugins (core/plugins/builtin/cpu/*.py); only ``compute`` differs - it calls the
sm_100a kernels through libwfb200.so.  Register with
``ctx.register(*waveformanalysis_b200.profiles.b200_default(), allow_override=True)``.
"""

from .dataframe import B200DataFramePlugin, B200PairedEventsPlugin, B200S1S2ClassifierPlugin
from .features import B200BasicFeaturesPlugin
from .filtering import B200WavePoolFilteredPlugin
from .grouping import B200GroupedEventsPlugin, B200HitGroupedPlugin
from .hits import B200HitFinderPlugin, B200ThresholdHitPlugin
from .merge import B200HitMergeClustersPlugin, B200HitMergedComponentsPlugin, B200HitMergePlugin
from .records import B200RecordsPlugin, B200WavePoolPlugin
from .streaming import B200HitThresholdStreamPlugin, B200SignalPeaksStreamPlugin
from .waveforms import B200WaveformsPlugin
from .widths import B200WaveformWidthIntegralPlugin, B200WaveformWidthPlugin

__all__ = [
    "B200BasicFeaturesPlugin",
    "B200ThresholdHitPlugin",
    "B200HitFinderPlugin",
    "B200WavePoolFilteredPlugin",
    "B200WaveformWidthPlugin",
    "B200WaveformWidthIntegralPlugin",
    "B200HitMergeClustersPlugin",
    "B200HitMergePlugin",
    "B200HitMergedComponentsPlugin",
    "B200HitGroupedPlugin",
    "B200GroupedEventsPlugin",
    "B200RecordsPlugin",
    "B200WavePoolPlugin",
    "B200SignalPeaksStreamPlugin",
    "B200HitThresholdStreamPlugin",
    "B200WaveformsPlugin",
    "B200DataFramePlugin",
    "B200PairedEventsPlugin",
    "B200S1S2ClassifierPlugin",
]
