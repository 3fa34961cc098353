"""Drop-in for the part of the reference's data_parser the interpretation drivers use (pt/data_parser.py:102-160):
PicDatabase walks <root>/<class number>/<clip id>/ and lists (id, label, path) items."""
import os
from collections import namedtuple

ListData = namedtuple("ListData", ["id", "label", "path"])


class PicDatabase(object):
    def __init__(self, data_root, is_test=False):
        self.data_root = data_root
        self.is_test = is_test
        self.classes = []
        self.input_data = self.read_json_input()

    def read_json_input(self):
        items, classes = [], []
        if not self.is_test:
            for class_dir in sorted(next(os.walk(self.data_root))[1]):
                classes.append(int(class_dir))
                base = os.path.join(self.data_root, class_dir)
                for clip_dir in sorted(next(os.walk(base))[1]):
                    items.append(ListData(clip_dir, class_dir, os.path.join(base, clip_dir)))
        self.classes = classes
        return items
