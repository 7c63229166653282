"""Mirror of the reference's nerf/dataset.py: the Blender-synthetic loader used by train.py / render_only.

Host-side I/O only (PIL + torchvision); nothing here touches the GPU.  Same public names and behaviour as the reference
(nerf/dataset.py:22-114): `AdaptiveResize(ratio)` (bilinear resize to int(h * ratio) x int(w * ratio)), `CustomDataSet`
(natural-sorted `*.png` of `<root>/train|test/` minus `*normal*` / `*alpha*`, `transforms_{train,test}[_div].json` with
`camera_angle_x[,_y]` and per-frame `transform_matrix`, white-background alpha compositing, scene_scale on the translation).
natsort is not a dependency: the natural order is implemented here.
"""
import json
import os
import re

import numpy as np
import torch
from torch import nn
from torch.utils import data


def natural_key(name):
    """'r_10.png' after 'r_9.png': digit runs compare as integers (what natsort.natsorted does for these file names)."""
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", name)]


class AdaptiveResize(nn.Module):
    def __init__(self, ratio):
        super().__init__()
        self.ratio = ratio

    def forward(self, image):
        from torchvision import transforms
        from torchvision.transforms import functional as TF
        size = (int(image.size[1] * self.ratio), int(image.size[0] * self.ratio))       # PIL size is (w, h)
        return TF.resize(image, size, transforms.InterpolationMode.BILINEAR, None, None)


class CustomDataSet(data.Dataset):
    def __init__(self, root_dir, transform, scene_scale=1.0, is_train=True, use_alpha=False, white_bkg=False, use_div=False):
        self.is_train = is_train
        self.root_dir = root_dir
        self.main_dir = root_dir + ("train/" if is_train else "test/")
        self.transform = transform
        names = [n for n in os.listdir(self.main_dir) if n.endswith("png") and "normal" not in n and "alpha" not in n]
        self.total_imgs = sorted(names, key=natural_key)
        self.use_alpha = use_alpha
        self.scene_scale = scene_scale
        self.white_bkg = white_bkg
        self.use_div = use_div
        stem = f"{self.root_dir}transforms_{'train' if is_train else 'test'}"
        self.cam_fov, self.tfs, self.divisions, self.weights = CustomDataSet.readFromJson(stem + ("_div.json" if use_div else ".json"), use_div)

    def __len__(self):
        return len(self.total_imgs)

    def _open(self, idx, rgba):
        from PIL import Image
        return Image.open(os.path.join(self.main_dir, self.total_imgs[idx]), mode="r").convert("RGBA" if rgba else "RGB")

    def __getitem__(self, idx):
        tensor_image = self.transform(self._open(idx, self.use_alpha or self.white_bkg))
        tf = self.tfs[idx].clone()
        if self.white_bkg:
            tensor_image = tensor_image[:3, ...] * tensor_image[-1:, ...] + (1. - tensor_image[-1:, ...])
        tf[:3, -1] *= self.scene_scale
        return tensor_image, tf

    def r_c(self):
        image, _ = self.__getitem__(0)
        return image.shape[1], image.shape[2]

    def cuda(self, flag=True):
        self.is_cuda = flag

    @staticmethod
    def readFromJson(path: str, use_div=False):
        with open(path, "r") as file:
            items = json.load(file)
        cam_fov = items["camera_angle_x"]
        if "camera_angle_y" in items:
            cam_fov = (cam_fov, items["camera_angle_y"])
        tfs = torch.from_numpy(np.stack([frame["transform_matrix"] for frame in items["frames"]], axis=0))[:, :3, :]
        division = items.get("division", None) if use_div else None
        weights = items.get("weights", None) if use_div else None
        return cam_fov, tfs.float(), division, weights

    def getCameraParam(self):
        return self.cam_fov, self.tfs

    def get_dataset(self, to_cuda: bool):
        """(camera fov, per-image transforms, images (N, C, H, W))."""
        images = torch.stack([self.transform(self._open(i, self.use_alpha)) for i in range(len(self))], dim=0).float()
        cam_fov, tfs = self.getCameraParam()
        if to_cuda:
            return cam_fov, tfs.cuda(), images.cuda()
        return cam_fov, tfs, images
