"""Golden values for the Dice / E-measure / S-measure definitions from the reference's OWN functions
(adapteacher/evaluation/dice_metric.py, imported with Detectron2 / pycocotools stubbed out).  Build container only:
    python -m oracle.gen_metric_golden      -> tests/golden/metrics.npz"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_metrics():
    for name in ("detectron2", "detectron2.evaluation", "detectron2.data", "pycocotools", "pycocotools.mask"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["detectron2.evaluation"].DatasetEvaluator = object
    sys.modules["detectron2.data"].MetadataCatalog = sys.modules["detectron2.data"].DatasetCatalog = object
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dice_metric", "/root/reference/adapteacher/evaluation/dice_metric.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cases():
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:64, 0:64]
    out = []
    for k in range(8):
        cy, cx, r = rng.uniform(20, 44, 2).tolist() + [rng.uniform(6, 18)]
        gt = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        dy, dx, dr = rng.uniform(-5, 5, 2).tolist() + [rng.uniform(-3, 3)]
        pred = (yy - cy - dy) ** 2 + (xx - cx - dx) ** 2 <= (r + dr) ** 2
        out.append((pred, gt))
    out.append((np.zeros((64, 64), bool), out[0][1]))
    out.append((out[1][0], np.zeros((64, 64), bool)))
    out.append((np.ones((64, 64), bool), np.ones((64, 64), bool)))
    return out


def main():
    ref = load_reference_metrics()
    res = {}
    for i, (pred, gt) in enumerate(cases()):
        inter = np.logical_and(pred, gt).sum()
        res[f"dice_{i}"] = np.array(2 * inter / (pred.sum() + gt.sum() + 1e-6))
        res[f"ea_{i}"] = np.array(ref.enhanced_align(pred, gt))
        res[f"sm_{i}"] = np.array(ref.Structure_measure().get_score(pred, gt))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **res)
    print({k: float(v) for k, v in res.items() if k.endswith("_0") or k.endswith("_8")})


if __name__ == "__main__":
    main()
