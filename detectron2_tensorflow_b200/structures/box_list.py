"""`BoxList` / `SparseBoxList` containers (lib/structures/box_list.py:7-171, 174-264).

These are the data contract at the operator boundary: a dict of tensors keyed by
field plus "trackings".  They carry torch tensors (host or device) instead of
tf.Tensors; semantics (field names, row-major `from_dense`, zero-padded `to_dense`
with an `is_valid` field) follow the reference.
"""
import torch


class _Trackings(object):
    """Per-batch side data ("trackings", e.g. image_shape) shared by both containers (box_list.py:30-44, 192-202)."""

    trackings = None

    def get_all_trackings(self):
        return self.trackings.keys()

    def set_tracking(self, name, value):
        self.trackings[name] = value

    def get_tracking(self, name):
        return self.trackings[name]

    def has_tracking(self, name):
        return name in self.trackings


class BoxList(_Trackings):
    """Box list collection."""

    def __init__(self, boxes):
        self.data = {}
        self.trackings = {}
        self.boxes = boxes

    @property
    def boxes(self):
        return self.get_field('boxes')

    @boxes.setter
    def boxes(self, boxes):
        assert self.data == {}, "BoxList is immutable."
        if boxes.shape[-1] != 4:
            raise ValueError('Invalid dimensions for box data.')
        if boxes.dtype != torch.float32:
            raise ValueError('Invalid tensor type: should be float32')
        self.data['boxes'] = boxes

    def num_boxes(self):
        return self.data['boxes'].shape[0]

    def get_all_fields(self):
        return self.data.keys()

    def get_extra_fields(self):
        return [k for k in self.data.keys() if k != 'boxes']

    def add_field(self, field, field_data):
        self.data[field] = field_data

    def has_field(self, field):
        return field in self.data

    def set_field(self, field, field_data):
        """box_list.py:106-121: replace the value of an existing field."""
        if not self.has_field(field):
            raise ValueError('field %s does not exist' % field)
        self.data[field] = field_data

    def get_field(self, field):
        if not self.has_field(field):
            raise ValueError('field ' + str(field) + ' does not exist')
        return self.data[field]

    def as_tensor_dict(self, fields=None):
        if fields is None:
            fields = self.get_all_fields()
        return {f: self.get_field(f) for f in fields}

    @classmethod
    def from_tensor_dict(cls, tensor_dict):
        boxlist = cls(tensor_dict['boxes'])
        for field in tensor_dict:
            if field != 'boxes':
                boxlist.add_field(field, tensor_dict[field])
        return boxlist


class SparseBoxList(_Trackings):
    """COO batch -> ragged view of a dense [N, R] BoxList (box_list.py:174-264)."""

    def __init__(self, indices, data, dense_shape):
        assert isinstance(data, BoxList), type(data)
        self.data = data
        self.dense_shape = dense_shape
        assert indices.shape[0] == data.boxes.shape[0]
        self.indices = indices
        self.trackings = {}

    def to_dense(self):
        """box_list.py:204-246: scatter rows to [N, R, ...], zero padding, adds `is_valid`."""
        N, R = int(self.dense_shape[0]), int(self.dense_shape[1])
        flat = self.indices[:, 0] * R + self.indices[:, 1]
        tensor_dict = {}
        for field in self.data.get_all_fields():
            if field == "is_valid":
                continue
            v = self.data.get_field(field)
            dense = torch.zeros((N * R,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            dense[flat] = v
            tensor_dict[field] = dense.reshape((N, R) + tuple(v.shape[1:]))
        mask = torch.zeros(N * R, dtype=torch.bool, device=self.indices.device)
        mask[flat] = True
        tensor_dict['is_valid'] = mask.reshape(N, R)
        dense = BoxList.from_tensor_dict(tensor_dict)
        for tracking in self.get_all_trackings():
            dense.set_tracking(tracking, self.get_tracking(tracking))
        return dense

    @classmethod
    def from_dense(cls, boxlist):
        """box_list.py:249-264: keep rows where `is_valid`, row-major (tf.where) order."""
        assert boxlist.has_field('is_valid')
        valid = boxlist.get_field('is_valid')
        dense_shape = tuple(boxlist.boxes.shape[:-1])
        indices = torch.nonzero(valid)  # row-major, int64 [M, 2]
        tensor_dict = {}
        for field in boxlist.get_all_fields():
            if field != 'is_valid':
                tensor_dict[field] = boxlist.get_field(field)[valid]
        data = BoxList.from_tensor_dict(tensor_dict)
        sparse = cls(indices, data, dense_shape)
        for tracking in boxlist.get_all_trackings():
            sparse.set_tracking(tracking, boxlist.get_tracking(tracking))
        return sparse
