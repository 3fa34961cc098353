"""Drop-in for the reference's data_loader_kth (pt/data_loader_kth.py:7-47): KTH clips stored as
<root>/<index>/frame%02d.jpg with class.txt (class number) and label.txt (tag such as
person17_boxing_d1_1) -> (data [3,T,H,W] 0..255, label[, tag])."""
import os

import torch

from _frames import read_clip


class KTHImLoader(torch.utils.data.Dataset):
    def __init__(self, root, json_file_input='', json_file_labels='', clip_size=16, is_val=False, get_item_id=False,
                 is_test=False, as_uint8=False):
        self.root = root
        self.clip_size = clip_size
        self.nclips = 1
        self.step_size = 1
        self.is_val = is_val
        self.get_item_id = get_item_id
        self.as_uint8 = as_uint8

    def __getitem__(self, index):
        base = os.path.join(self.root, str(index))
        data = read_clip(base, self.clip_size, self.as_uint8)
        with open(os.path.join(base, "class.txt")) as f:
            label = f.readline()
        if self.get_item_id:
            with open(os.path.join(base, "label.txt")) as f:
                tag = f.readline()
            return data, int(label), tag
        return data, int(label)

    def __len__(self):
        return len(os.listdir(self.root))
