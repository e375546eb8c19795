"""A numpy-backed, eager stand-in for the slice of the TensorFlow 1.x API that the reference's
post-backbone path uses -- GOLDEN-VECTOR GENERATION ONLY (tests/golden/make_golden.py).

TensorFlow cannot be installed in the build container, so the reference's Python cannot run as shipped.
But on this path the reference is *composition*: Python that chains ~90 stock TF ops.  Installing this
module as ``sys.modules['tensorflow']`` lets the reference's OWN source files
(/root/reference/lib/{layers,modeling,structures,utils}/...) execute unmodified, every ``tf.*`` call
being evaluated immediately on numpy arrays in float32.  The fixtures produced that way pin the
oracle's restatement of the reference's composition (operation order, tie rules of its own code,
padding, class offsets, level routing, ...) to the reference's actual code.

What this does NOT pin: the arithmetic inside the stock TF kernels, which is restated here a second
time, in numpy and independently of oracle/d2b_oracle.c (crop_and_resize, non_max_suppression, top_k,
avg_pool, SYMMETRIC pad = ``np.pad(mode='symmetric')``).  exp/log/sigmoid are numpy's (libm-grade),
not Eigen's Cephes forms, so values that pass through them agree to ~1 ulp, not bit for bit.

Semantics notes: elementwise ops keep float32 (numpy weak-scalar promotion == TF constant
conversion); ``tf.where(cond)`` returns int64 [n, rank] in row-major order; ``tf.argmax`` returns the
first maximum; ``tf.nn.top_k`` and ``non_max_suppression`` break ties toward the lower index
(SURVEY.md A.8/A.9); ``tf.dynamic_stitch`` lets later entries win; ``tf.map_fn`` loops over axis 0.
"""
import collections
import contextlib
import sys
import types

import numpy as np


# ------------------------------------------------------------------ tensor type
class Dim(int):
    @property
    def value(self):
        return int(self)

    def assert_is_compatible_with(self, other):
        assert other is None or int(other) == int(self), (self, other)


class TShape(tuple):
    def as_list(self):
        return [int(v) for v in self]

    @property
    def ndims(self):
        return len(self)

    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return TShape(r) if isinstance(i, slice) else Dim(r)

    def assert_is_compatible_with(self, other):
        assert tuple(self) == tuple(other), (self, other)

    def assert_has_rank(self, rank):
        assert len(self) == rank, (self, rank)


class T(np.ndarray):
    """ndarray with the tf.Tensor methods the reference touches."""

    def get_shape(self):
        return TShape(np.ndarray.shape.__get__(self))

    @property
    def shape(self):
        return TShape(np.ndarray.shape.__get__(self))

    def set_shape(self, shape):
        pass


def t(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    if a.dtype == np.float64 and dtype is None:
        a = a.astype(np.float32)  # TF's default float is float32
    return a.view(T)


# ------------------------------------------------------------------ nest helpers
def _flatten(s):
    if isinstance(s, dict):
        return [v for k in s for v in _flatten(s[k])]
    if isinstance(s, (list, tuple)):
        return [v for e in s for v in _flatten(e)]
    return [s]


def _pack(s, flat):
    it = iter(flat)

    def rec(x):
        if isinstance(x, dict):
            return {k: rec(v) for k, v in x.items()}
        if isinstance(x, tuple):
            return tuple(rec(v) for v in x)
        if isinstance(x, list):
            return [rec(v) for v in x]
        return next(it)
    return rec(s)


# ------------------------------------------------------------------ restated stock kernels (numpy, fp32)
def _crop_and_resize(image, boxes, box_ind, crop_size, method='bilinear', extrapolation_value=0, name=None):
    """tf.image.crop_and_resize, CPU kernel semantics; vectorised over channels and x."""
    image = np.asarray(image, np.float32)
    boxes = np.asarray(boxes, np.float32)
    box_ind = np.asarray(box_ind)
    N, H, W, C = image.shape
    ch, cw = int(crop_size[0]), int(crop_size[1])
    f = np.float32
    out = np.full((boxes.shape[0], ch, cw, C), extrapolation_value, np.float32)
    for b in range(boxes.shape[0]):
        y1, x1, y2, x2 = (f(v) for v in boxes[b])
        bi = int(box_ind[b])
        if bi < 0 or bi >= N:
            out[b] = 0
            continue
        hs = (y2 - y1) * f(H - 1) / f(ch - 1) if ch > 1 else f(0)
        ws = (x2 - x1) * f(W - 1) / f(cw - 1) if cw > 1 else f(0)
        xs = np.arange(cw, dtype=np.float32)
        in_x = x1 * f(W - 1) + xs * ws if cw > 1 else np.full(cw, f(0.5) * (x1 + x2) * f(W - 1), np.float32)
        okx = (in_x >= 0) & (in_x <= f(W - 1))
        lft = np.floor(in_x)
        lx = (in_x - lft).astype(np.float32)
        li = np.clip(lft.astype(np.int64), 0, W - 1)
        ri = np.clip(np.ceil(in_x).astype(np.int64), 0, W - 1)
        for y in range(ch):
            in_y = y1 * f(H - 1) + f(y) * hs if ch > 1 else f(0.5) * (y1 + y2) * f(H - 1)
            if not (in_y >= 0 and in_y <= f(H - 1)):
                continue
            top, bot = int(np.floor(in_y)), int(np.ceil(in_y))
            ly = f(in_y - f(top))
            tl, tr = image[bi, top, li], image[bi, top, ri]
            bl, br = image[bi, bot, li], image[bi, bot, ri]
            tv = tl + (tr - tl) * lx[:, None]
            bv = bl + (br - bl) * lx[:, None]
            row = tv + (bv - tv) * ly
            out[b, y][okx] = row[okx]
    return t(out)


def _iou_tf(a, b):
    """NonMaxSuppression CPU kernel IoU, fp32."""
    f = np.float32
    ymin_i, xmin_i, ymax_i, xmax_i = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
    ymin_j, xmin_j, ymax_j, xmax_j = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    ih = max(f(min(ymax_i, ymax_j) - max(ymin_i, ymin_j)), f(0))
    iw = max(f(min(xmax_i, xmax_j) - max(xmin_i, xmin_j)), f(0))
    inter = f(ih * iw)
    return f(inter / f(f(area_i + area_j) - inter))


def _non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float('-inf'),
                         name=None):
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    order = [i for i in np.argsort(-scores, kind='stable') if scores[i] > score_threshold]
    sel = []
    thr = np.float32(iou_threshold)
    for i in order:
        if len(sel) >= int(max_output_size):
            break
        if all(not (_iou_tf(boxes[i], boxes[j]) > thr) for j in reversed(sel)):
            sel.append(i)
    return t(np.array(sel, np.int32))


TopK = collections.namedtuple("TopKV2", ["values", "indices"])


def _top_k(x, k=1, sorted=True, name=None):
    x = np.asarray(x)
    assert x.ndim == 1
    idx = np.argsort(-x, kind='stable')[:int(k)].astype(np.int32)
    return TopK(t(x[idx]), t(idx))


def _avg_pool2d(x, kernel_size, stride=2, padding='VALID', **kw):
    x = np.asarray(x, np.float32)
    kh, kw_ = (kernel_size, kernel_size) if np.isscalar(kernel_size) else kernel_size
    sh, sw = (stride, stride) if np.isscalar(stride) else stride
    assert (sh, sw) == (kh, kw_) and x.shape[1] % kh == 0 and x.shape[2] % kw_ == 0
    n, h, w, c = x.shape
    acc = np.zeros((n, h // kh, w // kw_, c), np.float32)
    for dy in range(kh):
        for dx in range(kw_):
            acc = acc + x[:, dy::kh, dx::kw_]
    return t(acc / np.float32(kh * kw_))


# ------------------------------------------------------------------ the module
class _Missing(object):
    """Attribute sink for names only touched at import time (default arguments, enum constants)."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        return _Missing(self._name + "." + k)

    def __call__(self, *a, **k):
        raise NotImplementedError("tf shim: %s is not on the post-backbone path" % self._name)


class _TFModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Missing("tf." + k)


def build_tf():
    tf = _TFModule("tensorflow")
    tf.__version__ = "1.15.0"
    for n in ("float32", "float64", "int32", "int64", "uint8", "bool"):
        setattr(tf, n, getattr(np, n if n != "bool" else "bool_"))
    tf.Tensor = T
    tf.Variable = type("Variable", (), {})

    class SparseTensor(object):
        def __init__(self, indices, values, dense_shape):
            self.indices, self.values, self.dense_shape = indices, values, dense_shape
    tf.SparseTensor = SparseTensor

    class IndexedSlices(object):
        def __init__(self, values, indices, dense_shape=None):
            self.values, self.indices, self.dense_shape = values, indices, dense_shape
    tf.IndexedSlices = IndexedSlices

    @contextlib.contextmanager
    def scope(*a, **k):
        yield
    tf.name_scope = scope
    tf.variable_scope = scope
    tf.control_dependencies = scope

    def convert_to_tensor(x, dtype=None, name=None):
        if isinstance(x, IndexedSlices):  # == unsorted_segment_sum into the dense shape
            out = np.zeros([int(v) for v in np.asarray(x.dense_shape)], np.asarray(x.values).dtype)
            np.add.at(out, np.asarray(x.indices), np.asarray(x.values))
            return t(out)
        return t(x, dtype)
    tf.convert_to_tensor = convert_to_tensor
    tf.constant = lambda x, dtype=None, shape=None, name=None: t(x, dtype)
    tf.identity = lambda x, name=None: x
    tf.stop_gradient = lambda x, name=None: x

    def cast(x, dtype, name=None):
        return t(np.asarray(x).astype(dtype))
    tf.cast = cast
    tf.shape = lambda x, name=None, out_type=np.int32: t(np.array(np.asarray(x).shape, out_type))
    tf.size = lambda x, name=None, out_type=np.int32: out_type(np.asarray(x).size)
    tf.rank = lambda x, name=None: np.int32(np.asarray(x).ndim)
    tf.reshape = lambda x, shape, name=None: t(np.reshape(np.asarray(x), [int(v) for v in _flatten(list(shape))]
                                                          if not isinstance(shape, np.ndarray)
                                                          else [int(v) for v in shape]))
    tf.transpose = lambda x, perm=None, name=None: t(np.transpose(np.asarray(x), perm))
    tf.expand_dims = lambda x, axis=None, name=None, dim=None: t(np.expand_dims(np.asarray(x), axis if axis is not None else dim))
    tf.squeeze = lambda x, axis=None, name=None, squeeze_dims=None: t(np.squeeze(
        np.asarray(x), axis=None if (axis if axis is not None else squeeze_dims) is None
        else tuple(np.atleast_1d(axis if axis is not None else squeeze_dims))))
    tf.concat = lambda values, axis, name=None: t(np.concatenate([np.asarray(v) for v in values], axis=int(axis)))
    tf.stack = lambda values, axis=0, name=None: t(np.stack([np.asarray(v) for v in values], axis=axis))
    tf.unstack = lambda x, num=None, axis=0, name=None: [t(v) for v in np.moveaxis(np.asarray(x), axis, 0)]
    tf.tile = lambda x, multiples, name=None: t(np.tile(np.asarray(x), [int(m) for m in multiples]))
    def range_(start, limit=None, delta=1, dtype=None, name=None):
        if limit is None:
            start, limit = 0, start
        return t(np.arange(int(start), int(limit), int(delta), dtype=dtype or np.int32))
    tf.range = range_
    tf.meshgrid = lambda *a, **k: [t(m) for m in np.meshgrid(*[np.asarray(v) for v in a], indexing=k.get("indexing", "xy"))]
    tf.zeros = lambda shape, dtype=np.float32, name=None: t(np.zeros([int(v) for v in np.atleast_1d(shape)], dtype))
    tf.ones = lambda shape, dtype=np.float32, name=None: t(np.ones([int(v) for v in np.atleast_1d(shape)], dtype))
    tf.zeros_like = lambda x, dtype=None, name=None, optimize=True: t(np.zeros_like(np.asarray(x), dtype=dtype))
    tf.ones_like = lambda x, dtype=None, name=None, optimize=True: t(np.ones_like(np.asarray(x), dtype=dtype))

    def split(value, num_or_size_splits, axis=0, num=None, name=None):
        v = np.asarray(value)
        if isinstance(num_or_size_splits, int):
            return [t(p) for p in np.split(v, num_or_size_splits, axis=axis)]
        return [t(p) for p in np.split(v, np.cumsum(num_or_size_splits)[:-1], axis=axis)]
    tf.split = split

    def slice_(x, begin, size, name=None):
        x = np.asarray(x)
        idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
        return t(x[idx])
    tf.slice = slice_

    def where(condition, x=None, y=None, name=None):
        c = np.asarray(condition)
        if x is None:
            return t(np.argwhere(c).astype(np.int64).reshape(-1, c.ndim))
        xa, ya = np.asarray(x), np.asarray(y)
        if c.ndim == 1 and xa.ndim > 1:  # TF1: a vector condition selects rows
            c = c.reshape((-1,) + (1,) * (xa.ndim - 1))
        return t(np.where(c, xa, ya))
    tf.where = where
    tf.gather = lambda params, indices, validate_indices=None, name=None, axis=0: t(
        np.take(np.asarray(params), np.asarray(indices).astype(np.int64), axis=axis))

    def gather_nd(params, indices, name=None):
        p, i = np.asarray(params), np.asarray(indices).astype(np.int64)
        return t(p[tuple(i[..., k] for k in range(i.shape[-1]))])
    tf.gather_nd = gather_nd
    tf.boolean_mask = lambda tensor, mask, name=None, axis=None: t(np.asarray(tensor)[np.asarray(mask, bool)])

    def dynamic_stitch(indices, data, name=None):
        n = max([int(np.max(i)) + 1 for i in indices if np.size(i)], default=0)
        first = next(np.asarray(d) for d, i in zip(data, indices) if True)
        rest = np.asarray(data[0]).shape[np.asarray(indices[0]).ndim:]
        out = np.zeros((n,) + tuple(rest), first.dtype)
        for i, d in zip(indices, data):
            out[np.asarray(i).astype(np.int64)] = np.asarray(d)
        return t(out)
    tf.dynamic_stitch = dynamic_stitch
    tf.invert_permutation = lambda x, name=None: t(np.argsort(np.asarray(x), kind='stable').astype(np.asarray(x).dtype))
    tf.reverse_v2 = lambda x, axis, name=None: t(np.flip(np.asarray(x), axis=tuple(np.atleast_1d(axis))))

    tf.cond = lambda pred, true_fn=None, false_fn=None, name=None, strict=False: true_fn() if bool(pred) else false_fn()

    def map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=True, swap_memory=False,
               infer_shape=True, name=None):
        flat = _flatten(elems)
        n = np.asarray(flat[0]).shape[0]
        outs = []
        for i in range(n):
            outs.append(fn(_pack(elems, [t(np.asarray(f)[i]) for f in flat])))
        struct = outs[0]
        cols = list(zip(*[_flatten(o) for o in outs]))
        return _pack(struct, [t(np.stack([np.asarray(v) for v in c], 0)) for c in cols])
    tf.map_fn = map_fn

    # ---- elementwise / reductions (float32 stays float32)
    def un(f):
        return lambda x, name=None: t(f(np.asarray(x)))

    def bi(f):
        return lambda x, y, name=None: t(f(np.asarray(x) if not np.isscalar(x) else x,
                                           np.asarray(y) if not np.isscalar(y) else y))
    tf.sqrt, tf.log, tf.exp, tf.floor, tf.square, tf.atan = (un(np.sqrt), un(np.log), un(np.exp), un(np.floor),
                                                             un(np.square), un(np.arctan))
    tf.logical_not = un(np.logical_not)
    tf.is_finite = un(np.isfinite)
    tf.minimum, tf.maximum, tf.add, tf.truediv, tf.div = (bi(np.minimum), bi(np.maximum), bi(np.add),
                                                          bi(np.true_divide), bi(np.true_divide))
    tf.equal, tf.not_equal, tf.greater, tf.greater_equal, tf.less, tf.less_equal = (
        bi(np.equal), bi(np.not_equal), bi(np.greater), bi(np.greater_equal), bi(np.less), bi(np.less_equal))
    tf.logical_and, tf.logical_or = bi(np.logical_and), bi(np.logical_or)
    tf.clip_by_value = lambda x, clip_value_min, clip_value_max, name=None: t(
        np.minimum(np.maximum(np.asarray(x), clip_value_min), clip_value_max))

    def red(f):
        def r(x, axis=None, keepdims=False, name=None, keep_dims=None, reduction_indices=None):
            ax = axis if axis is not None else reduction_indices
            ax = tuple(np.atleast_1d(ax)) if ax is not None else None
            return t(f(np.asarray(x), axis=ax, keepdims=bool(keepdims or keep_dims)))
        return r
    tf.reduce_max, tf.reduce_min, tf.reduce_sum, tf.reduce_any, tf.reduce_all, tf.reduce_mean, tf.reduce_prod = (
        red(np.max), red(np.min), red(np.sum), red(np.any), red(np.all), red(np.mean), red(np.prod))
    tf.count_nonzero = lambda x, axis=None, dtype=np.int64, **k: dtype(np.count_nonzero(np.asarray(x), axis=axis))
    tf.argmax = lambda x, axis=None, name=None, dimension=None, output_type=np.int64: t(
        np.argmax(np.asarray(x), axis=axis if axis is not None else dimension).astype(output_type))
    tf.matmul = lambda a, b, transpose_a=False, transpose_b=False, name=None: t(
        (np.asarray(a).T if transpose_a else np.asarray(a)) @ (np.asarray(b).T if transpose_b else np.asarray(b)))

    def band_part(x, num_lower, num_upper, name=None):
        x = np.asarray(x)
        m, n = x.shape[-2:]
        i, j = np.arange(m)[:, None], np.arange(n)[None, :]
        keep = ((num_lower < 0) | ((i - j) <= num_lower)) & ((num_upper < 0) | ((j - i) <= num_upper))
        return t(np.where(keep, x, np.zeros_like(x)))
    tf.matrix_band_part = band_part
    tf.linalg = types.SimpleNamespace(band_part=band_part)

    def pad(tensor, paddings, mode='CONSTANT', name=None, constant_values=0):
        x = np.asarray(tensor)
        pw = [(int(a), int(b)) for a, b in np.asarray(paddings)]
        if mode.upper() == 'CONSTANT':
            return t(np.pad(x, pw, mode='constant', constant_values=constant_values))
        return t(np.pad(x, pw, mode={'SYMMETRIC': 'symmetric', 'REFLECT': 'reflect'}[mode.upper()]))
    tf.pad = pad

    def sparse_to_dense(sp, default_value=0, validate_indices=True, name=None):
        vals = np.asarray(sp.values)
        shape = [int(v) for v in np.asarray(sp.dense_shape)]
        out = np.full(shape, default_value, vals.dtype)
        idx = np.asarray(sp.indices).astype(np.int64)
        out[tuple(idx[:, k] for k in range(idx.shape[1]))] = vals
        return t(out)
    tf.sparse = types.SimpleNamespace(to_dense=sparse_to_dense, SparseTensor=SparseTensor)

    def softmax(x, axis=-1, name=None):
        x = np.asarray(x, np.float32)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return t(e / e.sum(axis=axis, keepdims=True))
    def max_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
        x = np.asarray(value)
        kh, kw = ksize[1], ksize[2]
        assert list(strides) == [1, 1, 1, 1] and padding == "VALID"
        oh, ow = x.shape[1] - kh + 1, x.shape[2] - kw + 1
        out = np.full((x.shape[0], oh, ow, x.shape[3]), -np.inf, x.dtype)
        for dy in range(kh):
            for dx in range(kw):
                out = np.maximum(out, x[:, dy:dy + oh, dx:dx + ow])
        return t(out)
    def conv2d(x, filters, strides=None, padding="VALID", **kw):
        x, f = np.asarray(x, np.float32), np.asarray(filters, np.float32)
        assert f.shape[0] == 1 and f.shape[1] == 1 and padding == "VALID", "shim: 1x1 VALID convolutions only"
        return t(np.einsum("nhwc,co->nhwo", x, f[0, 0]).astype(np.float32))

    def resize_same(images, size, method=None, **kw):
        """tf.compat.v2.image.resize(images NHWC, size, 'bilinear'): ResizeBilinear with half-pixel centres, fp32,
        one rounding per written operation (numpy float32 arithmetic)."""
        x = np.asarray(images, np.float32)
        oh, ow = (int(v) for v in size)
        if (oh, ow) == x.shape[1:3]:
            return t(x)
        assert method in (None, "bilinear"), "shim: bilinear resize only"
        f32 = np.float32

        def taps(n_in, n_out):
            scale = f32(n_in) / f32(n_out)
            src = (np.arange(n_out, dtype=np.float32) + f32(0.5)) * scale - f32(0.5)
            fl = np.floor(src)
            lo = np.maximum(fl.astype(np.int64), 0)
            hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
            return lo, hi, (src - fl).astype(np.float32)
        ylo, yhi, yl = taps(x.shape[1], oh)
        xlo, xhi, xl = taps(x.shape[2], ow)
        xl_ = xl[None, None, :, None]
        yl_ = yl[None, :, None, None]
        r0, r1 = x[:, ylo], x[:, yhi]
        top = r0[:, :, xlo] + (r0[:, :, xhi] - r0[:, :, xlo]) * xl_
        bot = r1[:, :, xlo] + (r1[:, :, xhi] - r1[:, :, xlo]) * xl_
        return t((top + (bot - top) * yl_).astype(np.float32))
    tf.nn = types.SimpleNamespace(conv2d=conv2d, max_pool=max_pool, top_k=_top_k, sigmoid=un(lambda x: (np.float32(1) / (np.float32(1) + np.exp(-x)))),
                                  softmax=softmax, relu=un(lambda x: np.maximum(x, 0)))
    tf.image = types.SimpleNamespace(crop_and_resize=_crop_and_resize, non_max_suppression=_non_max_suppression,
                                     resize_images=None, ResizeMethod=types.SimpleNamespace(BILINEAR=0, NEAREST_NEIGHBOR=1))
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None, histogram=lambda *a, **k: None)
    tf.Assert = lambda cond, data=None, **k: None
    tf.assert_equal = lambda *a, **k: None
    tf.assert_positive = lambda *a, **k: None
    tf.math = types.SimpleNamespace(is_finite=tf.is_finite, log=tf.log, exp=tf.exp)
    slim = types.SimpleNamespace(avg_pool2d=_avg_pool2d, max_pool2d=None, flatten=None, model_variable=None,
                                 add_arg_scope=lambda f: f)
    tf.contrib = types.SimpleNamespace(slim=slim, framework=types.SimpleNamespace(add_arg_scope=lambda f: f))
    tf.compat = types.SimpleNamespace(v2=types.SimpleNamespace(image=types.SimpleNamespace(resize=resize_same)),
                                      v1=tf)
    tf.random_normal_initializer = lambda *a, **k: None
    tf.constant_initializer = lambda *a, **k: None
    return tf


def install():
    """Puts the shim at sys.modules['tensorflow'] (+ the two private submodules the reference imports)."""
    sys.dont_write_bytecode = True  # /root/reference is read-only
    tf = build_tf()
    sys.modules["tensorflow"] = tf
    for name in ("tensorflow.python", "tensorflow.python.client", "tensorflow.python.client.device_lib",
                 "tensorflow.python.training", "tensorflow.python.training.moving_averages"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.python.client"].device_lib = sys.modules["tensorflow.python.client.device_lib"]
    sys.modules["tensorflow.python.training"].moving_averages = sys.modules["tensorflow.python.training.moving_averages"]
    return tf
