"""Rendering is outside the hot path this package replaces.

``Environment.data_rods`` keeps the reference's callback keys and array shapes, so the reference's
matplotlib/ffmpeg post-processing (``br2.visualize.post_processing`` / ``twist_angle`` upstream) can be
pointed at it unchanged; it is not re-implemented here.
"""


def _unavailable(*args, **kwargs):
    raise NotImplementedError(
        "video rendering is not part of the B200 hot-path replacement; feed Environment.data_rods "
        "(same keys/shapes as the reference's FreeCallback) to the reference's br2.visualize functions")


plot_video_with_surface = _unavailable
visual_twist_with_surface = _unavailable
