"""Small pieces of glue the reference's callers expect around the path (not on the path itself)."""
from __future__ import annotations

import numpy as np
import torch.utils.data


class DataWrapper(torch.utils.data.Dataset):
    """Dataset view that skips annotations (detector.py:125-145): yields {'image','image_meta'}."""

    def __init__(self, dataset):
        super().__init__()
        self.dataset = dataset

    def __getitem__(self, index):
        image, image_id = self.dataset.load_image(index)
        image_meta = {"index": index, "image_id": image_id, "orig_size": np.array(image.shape, dtype=np.int32)}
        image, image_meta, _ = self.dataset.preprocess(image, image_meta)
        return {"image": image.transpose(2, 0, 1), "image_meta": image_meta}

    def __len__(self):
        return len(self.dataset)
