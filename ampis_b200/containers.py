"""Container types used by the AMPIS API: detectron2's ``Instances``, ``Boxes``, ``BitMasks``,
``PolygonMasks`` and ``BoxMode``.

AMPIS uses them purely as containers.  If detectron2 is installed the
real classes are re-exported; otherwise the minimal stand-ins below provide the same attribute
surface (``image_size``, ``_fields``, ``has``, ``get``, ``set``, ``__getitem__``, ``__len__``,
``.polygons``, ``.tensor``), which is also enough to unpickle prediction files written by the
reference (``load_pickle``).
"""
import enum
import pickle

import numpy as np
import torch

try:  # pragma: no cover - detectron2 is not part of this image
    from detectron2.structures import BitMasks, Boxes, BoxMode, Instances, PolygonMasks
    HAVE_DETECTRON2 = True
except Exception:
    HAVE_DETECTRON2 = False

    class BoxMode(enum.IntEnum):
        XYXY_ABS = 0
        XYWH_ABS = 1
        XYXY_REL = 2
        XYWH_REL = 3
        XYWHA_ABS = 4

    class Boxes(object):
        def __init__(self, tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
            if tensor.numel() == 0:
                tensor = tensor.reshape((-1, 4))
            assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
            self.tensor = tensor

        def __getitem__(self, item):
            if isinstance(item, int):
                return Boxes(self.tensor[item].view(1, -1))
            b = self.tensor[item]
            assert b.dim() == 2, 'Indexing on Boxes with {} failed to return a matrix!'.format(item)
            return Boxes(b)

        def __len__(self):
            return self.tensor.shape[0]

        def area(self):
            b = self.tensor
            return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

        def to(self, *args, **kwargs):
            return Boxes(self.tensor.to(*args, **kwargs))

    class BitMasks(object):
        def __init__(self, tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.bool)
            assert tensor.dim() == 3, tensor.size()
            self.image_size = tensor.shape[1:]
            self.tensor = tensor

        def __getitem__(self, item):
            if isinstance(item, int):
                return BitMasks(self.tensor[item].unsqueeze(0))
            m = self.tensor[item]
            assert m.dim() == 3
            return BitMasks(m)

        def __len__(self):
            return self.tensor.shape[0]

    class PolygonMasks(object):
        def __init__(self, polygons):
            if not isinstance(polygons, list):
                raise ValueError("Cannot create PolygonMasks: Expect a list of list of polygons per image. "
                                 "Got '{}' instead.".format(type(polygons)))

            def _make_array(t):
                if isinstance(t, torch.Tensor):
                    t = t.cpu().numpy()
                return np.asarray(t).astype('float64')

            def process(polygons_per_instance):
                if not isinstance(polygons_per_instance, list):
                    raise ValueError("Cannot create polygons: Expect a list of polygons per instance. "
                                     "Got '{}' instead.".format(type(polygons_per_instance)))
                polygons_per_instance = [_make_array(p) for p in polygons_per_instance]
                for polygon in polygons_per_instance:
                    if len(polygon) % 2 != 0 or len(polygon) < 6:
                        raise ValueError('Cannot create a polygon from {} coordinates.'.format(len(polygon)))
                return polygons_per_instance

            self.polygons = [process(p) for p in polygons]

        def __len__(self):
            return len(self.polygons)

        def __iter__(self):
            return iter(self.polygons)

        def __getitem__(self, item):
            if isinstance(item, int):
                selected = [self.polygons[item]]
            elif isinstance(item, slice):
                selected = self.polygons[item]
            elif isinstance(item, list):
                selected = [self.polygons[i] for i in item]
            elif isinstance(item, (torch.Tensor, np.ndarray)):
                item = torch.as_tensor(item)
                if item.dtype == torch.bool:
                    assert item.dim() == 1, item.shape
                    item = item.nonzero().squeeze(1).cpu().numpy().tolist()
                elif item.dtype in [torch.int32, torch.int64]:
                    item = item.cpu().numpy().tolist()
                else:
                    raise ValueError('Unsupported tensor dtype={} for indexing!'.format(item.dtype))
                selected = [self.polygons[i] for i in item]
            else:
                raise ValueError('Unsupported index type %s' % type(item))
            return PolygonMasks(selected)

    class Instances(object):
        def __init__(self, image_size, **kwargs):
            self._image_size = image_size
            self._fields = {}
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith('_'):
                super().__setattr__(name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            if name == '_fields' or name not in self._fields:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
            return self._fields[name]

        def set(self, name, value):
            data_len = len(value)
            if len(self._fields):
                assert len(self) == data_len, \
                    'Adding a field of length {} to a Instances of length {}'.format(data_len, len(self))
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def remove(self, name):
            del self._fields[name]

        def get(self, name):
            return self._fields[name]

        def get_fields(self):
            return self._fields

        def __getitem__(self, item):
            if type(item) == int:
                if item >= len(self) or item < -len(self):
                    raise IndexError('Instances index out of range!')
                else:
                    item = slice(item, None, len(self))
            ret = Instances(self._image_size)
            for k, v in self._fields.items():
                ret.set(k, v[item])
            return ret

        def __len__(self):
            for v in self._fields.values():
                return v.__len__()
            raise NotImplementedError('Empty Instances does not support __len__!')

        def __iter__(self):
            raise NotImplementedError('`Instances` object is not iterable!')

        def __str__(self):
            s = self.__class__.__name__ + '('
            s += 'num_instances={}, '.format(len(self))
            s += 'image_height={}, '.format(self._image_size[0])
            s += 'image_width={}, '.format(self._image_size[1])
            s += 'fields=[{}])'.format(', '.join('{}: {}'.format(k, v) for k, v in self._fields.items()))
            return s

        __repr__ = __str__


#: exact (module, name) pairs a prediction pickle may reference besides detectron2's Instances: what numpy arrays,
#: numpy scalars, paths and plain containers pickle to.  Nothing callable with side effects (no builtins.eval / exec /
#: getattr / __import__, no numpy.* functions beyond the array reconstructors).
_PICKLE_ALLOWED = frozenset([
    ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
    ('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'),
    ('numpy.core.numeric', '_frombuffer'), ('numpy._core.numeric', '_frombuffer'),
    ('numpy', 'ndarray'), ('numpy', 'dtype'),
    ('builtins', 'set'), ('builtins', 'frozenset'), ('builtins', 'slice'), ('builtins', 'complex'),
    ('builtins', 'bytearray'), ('builtins', 'range'), ('builtins', 'list'), ('builtins', 'dict'),
    ('builtins', 'tuple'), ('builtins', 'int'), ('builtins', 'float'), ('builtins', 'bool'), ('builtins', 'str'),
    ('builtins', 'bytes'),
    ('collections', 'OrderedDict'), ('collections', 'defaultdict'),
    ('_codecs', 'encode'),
    ('pathlib', 'PosixPath'), ('pathlib', 'WindowsPath'), ('pathlib', 'PurePosixPath'),
    ('pathlib', 'PureWindowsPath'), ('pathlib', 'PurePath'), ('pathlib', 'Path'),
])


class _Unpickler(pickle.Unpickler):
    """Loads prediction pickles written by the reference (data_utils.format_outputs output).  Only a detectron2
    ``Instances`` (mapped onto the class above) and the exact names in _PICKLE_ALLOWED resolve; anything else --
    in particular every other builtin and numpy function -- raises UnpicklingError."""

    def find_class(self, module, name):
        if (module.startswith('detectron2') or module == __name__) and name == 'Instances':
            return Instances
        if (module, name) in _PICKLE_ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError('refusing to load %s.%s' % (module, name))


def load_pickle(path):
    with open(path, 'rb') as f:
        return _Unpickler(f).load()
