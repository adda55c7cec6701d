"""Image masks — thin host-side counterpart of the reference's mask_dataloader.py (:5-25).  The python renderer of the
reference ignores masks; they only change the camera resolution rule (camera_handler.py:55-61)."""
import os

import torch


def load_image_masks(directory_path):
    import cv2
    masks = {}
    for filename in os.listdir(directory_path):
        path = os.path.join(directory_path, filename)
        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if img is None:
            print(f"WARNING: Could not load mask with name {filename}")
            continue
        masks[str(os.path.basename(path).split(".")[0])] = torch.tensor(img).to(torch.int)
    return masks
