"""Drop-in for the reference's data_loader_jpg (pt/data_loader_jpg.py:8-41): Something-Something clips stored as
<root>/<class>/<clip id>/frame%02d.jpg -> (data [3,T,H,W] 0..255, label[, id])."""
import torch

from data_parser import PicDatabase
from _frames import read_clip


class ImLoader(torch.utils.data.Dataset):
    def __init__(self, root, json_file_input="", json_file_labels="", clip_size=16, is_val=False, get_item_id=False,
                 is_test=False, as_uint8=False):
        self.dataset_object = PicDatabase(root)
        self.path_data = self.dataset_object.input_data
        self.classes = self.dataset_object.classes
        self.root = root
        self.clip_size = clip_size
        self.nclips = 1
        self.step_size = 1
        self.is_val = is_val
        self.get_item_id = get_item_id
        self.as_uint8 = as_uint8

    def __getitem__(self, index):
        item = self.path_data[index]
        data = read_clip(item.path, self.clip_size, self.as_uint8)
        if self.get_item_id:
            return data, int(item.label), item.id
        return data, int(item.label)

    def __len__(self):
        return len(self.path_data)
