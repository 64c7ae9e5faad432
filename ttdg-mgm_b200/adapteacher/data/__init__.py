from .build import build_detection_test_loader, DatasetCatalog, DatasetMapper, InferenceSampler  # noqa: F401
from .datasets.builtin import register_coco_instances, register_synthetic  # noqa: F401
