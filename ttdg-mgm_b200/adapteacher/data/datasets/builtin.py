"""Dataset registration for the test path; mirrors reference adapteacher/data/datasets/builtin.py:193-225 (the Fundus /
Polyp / COVID splits, COCO-format json + image directory) on a minimal catalog (Detectron2's DatasetCatalog /
``register_coco_instances`` / pycocotools are not dependencies here).  ``register_synthetic`` adds the seeded synthetic
datasets of SURVEY 8(d) under names like ``synthetic_fundus_16`` (no files needed)."""
import json
import os

import numpy as np

from ..build import DatasetCatalog

# name -> (json, image dir), the reference's table (builtin.py:196-222), relative to the working directory
SPLITS = {}
for _fam, _names in (("Fundus", ("Drishti_GS", "ORIGA", "REFUGE", "RIM_ONE_r3")),):
    for _n in _names:
        for _s in ("train", "test"):
            SPLITS[f"{_n}_{_s}"] = (f"datasets/{_fam}/{_n}_{_s}.json", f"datasets/{_fam}/{_n}/{_s}/image")
SPLITS["REFUGE_Valid"] = ("datasets/Fundus/REFUGE_Valid.json", "datasets/Fundus/REFUGE_Valid/image")
for _key, _n in (("BKAI", "BKAI"), ("CVC_ClinicDB", "CVC-ClinicDB"), ("ETIS_LaribPolypDB", "ETIS-LaribPolypDB"), ("Kvasir_SEG", "Kvasir-SEG")):
    for _s in ("train", "test"):
        SPLITS[f"{_key}_{_s}"] = (f"datasets/Polyp/{_n}_{_s}.json", f"datasets/Polyp/{_n}/{_s}/image")
SPLITS["COVID_train"] = ("datasets/covid19/0_train.json", "datasets/covid19/0/train/images")
SPLITS["COVID_test"] = ("datasets/covid19/0_test.json", "datasets/covid19/0/test/images")


def load_coco_json(json_file, image_root):
    """COCO instances json -> list of Detectron2-style dataset dicts (category ids remapped to 0..K-1 in id order)."""
    with open(json_file) as f:
        coco = json.load(f)
    cat_ids = sorted(c["id"] for c in coco.get("categories", []))
    remap = {c: i for i, c in enumerate(cat_ids)}
    anns = {}
    for a in coco.get("annotations", []):
        if a.get("iscrowd", 0):
            continue
        anns.setdefault(a["image_id"], []).append({"bbox": a["bbox"], "bbox_mode": "XYWH_ABS", "category_id": remap[a["category_id"]],
                                                   "segmentation": a.get("segmentation", [])})
    out = []
    for im in coco["images"]:
        out.append({"file_name": os.path.join(image_root, im["file_name"]), "height": im["height"], "width": im["width"],
                    "image_id": im["id"], "annotations": anns.get(im["id"], [])})
    return out


def register_coco_instances(name, metadata, json_file, image_root):
    DatasetCatalog.register(name, lambda: load_coco_json(json_file, image_root))


def register_all_fetus():
    for key, (json_file, image_root) in SPLITS.items():
        register_coco_instances(key, {}, json_file, image_root)


def register_synthetic(name):
    """``synthetic_fundus_<N>`` / ``synthetic_polyp_<N>[_<S>]``: N seeded images of size S (default 512 / 384)."""
    parts = name.split("_")
    kind, n = parts[1], int(parts[2])
    size = int(parts[3]) if len(parts) > 3 else (384 if kind == "polyp" else 512)

    def make():
        from ttdg_b200 import synth
        out = []
        for i in range(n):
            d = synth.fundus_like_image(i, size, polyp=(kind == "polyp"))
            out.append({"image_id": i, "height": size, "width": size, "image": d["image"],
                        "annotations": [{"category_id": int(c), "bbox": [float(v) for v in b], "bbox_mode": "XYXY_ABS",
                                         "mask": np.asarray(m)} for b, c, m in zip(d["gt_boxes"], d["gt_classes"], d["gt_masks"])]})
        return out
    DatasetCatalog.register(name, make)


register_all_fetus()
