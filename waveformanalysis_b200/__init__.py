"""B200-native hot path of WaveformAnalysis: drop-in plugins over libwfb200.so."""


def release_device_cache(run_id=None) -> None:
    """Free the device-resident copies kept between plugin calls (``residency``) and the library's staging buffers."""
    from . import _lib, residency

    residency.release(run_id)
    if run_id is None:
        try:
            _lib.load().wfb_release_cache()
        except Exception:
            pass
