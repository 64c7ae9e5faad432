from .detection_checkpoint import DetectionCheckpointer, DetectionTSCheckpointer  # noqa: F401
