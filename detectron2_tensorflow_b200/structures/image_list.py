"""Minimal `ImageList` (lib/structures/image_list.py:7): the hot path only reads `image_shapes`."""


class ImageList(object):
    def __init__(self, tensor, image_shapes):
        self.tensor = tensor
        self.image_shapes = image_shapes  # [N, 2] in (h, w) order
