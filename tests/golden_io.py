"""Rebuild the inputs stored in tests/golden/*.npz (see oracle/make_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(case):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    data = {
        "query_image": torch.from_numpy(z["image_u8"]).float() / 255,
        "keypoints3d": torch.from_numpy(z["keypoints3d"]),
        "descriptors3d_db": torch.from_numpy(z["descriptors3d_db_f16"]).float(),
        "descriptors3d_coarse_db": torch.from_numpy(z["descriptors3d_coarse_db_f16"]).float(),
    }
    if "query_image_scale" in z.files:
        data["query_image_scale"] = torch.from_numpy(z["query_image_scale"])
    return data, z
