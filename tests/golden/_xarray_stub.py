"""A minimal stand-in for the part of xarray the reference's codes_revision/BasicsSetup.py and NetworkSetup.py use
(labelled n-d arrays with `.loc` get/set, orthogonal label indexing, scalar-broadcast arithmetic).

TEST INFRASTRUCTURE ONLY: it exists so that tests/golden/make_golden_revision.py can import and run the REFERENCE's own
revision generator in this container (xarray and brian2 are not installed) and record what it draws.  The product
(rescience_hass_hertag_durstewitz_2016_b200/codes_revision) does not import it: it holds the same tables as plain numpy
arrays indexed by group number.  Its correctness is cross-checked in the generator by comparing every table it
assembles with the tables of the reference's codes/ParamSetup.py, which are built with plain numpy.
"""
import numpy as np


def _positions(coord, labels):
    """label(s) -> integer position(s) along one dimension (labels must exist)."""
    coord = np.asarray(coord)
    scalar = np.ndim(labels) == 0
    lab = np.atleast_1d(np.asarray(labels))
    if coord.dtype.kind in "iuf" and lab.dtype.kind in "iuf":
        order = np.argsort(coord, kind="stable")
        pos = order[np.clip(np.searchsorted(coord[order], lab), 0, len(coord) - 1)] if len(coord) else np.zeros(0, dtype=int)
        if len(lab) and not np.array_equal(coord[pos], lab):
            raise KeyError("labels not in coordinate")
    else:
        index = {str(c): i for i, c in enumerate(coord)}
        pos = np.array([index[str(x)] for x in lab], dtype=int)
    return int(pos[0]) if scalar else pos


class _Loc:
    def __init__(self, arr):
        self.arr = arr

    def _key(self, key):
        a = self.arr
        if isinstance(key, dict):
            sel = dict(key)
        else:
            if not isinstance(key, tuple):
                key = (key,)
            sel = {a.dims[i]: k for i, k in enumerate(key)}
        return {d: (v if isinstance(v, slice) else _positions(a.coords_list[a.dims.index(d)], v.values if isinstance(v, DataArray) else v))
                for d, v in sel.items()}

    def __getitem__(self, key):
        return self.arr.isel(self._key(key))

    def __setitem__(self, key, value):
        self.arr._assign(self._key(key), value)


class _Coord:
    def __init__(self, values):
        self.values = np.asarray(values)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)


class DataArray:
    __array_priority__ = 100

    def __init__(self, data=(), coords=None, dims=None, name=None):
        self.values = np.array(data.values if isinstance(data, DataArray) else data)
        nd = self.values.ndim
        if dims is None:
            dims = ["dim_%d" % i for i in range(nd)]
        if isinstance(dims, str):
            dims = [dims]
        self.dims = list(dims)
        if coords is None:
            coords = [np.arange(s) for s in self.values.shape]
        self.coords_list = [np.asarray(c) for c in coords]
        self.name = name

    # ---- basic protocol ----------------------------------------------------------------------------------------------
    __hash__ = object.__hash__

    @property
    def coords(self):
        return {d: _Coord(c) for d, c in zip(self.dims, self.coords_list)}

    @property
    def shape(self):
        return self.values.shape

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def loc(self):
        return _Loc(self)

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        for i in range(len(self.values)):
            yield self.isel({self.dims[0]: i})

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __float__(self):
        return float(self.values)

    def __int__(self):
        return int(self.values)

    def __index__(self):
        return int(self.values)

    def __bool__(self):
        return bool(self.values)

    def to_numpy(self):
        return np.asarray(self.values)

    def copy(self):
        return DataArray(self.values.copy(), [c.copy() for c in self.coords_list], list(self.dims), self.name)

    def astype(self, t):
        return DataArray(self.values.astype(t), self.coords_list, self.dims, self.name)

    def transpose(self):
        return DataArray(self.values.T, self.coords_list[::-1], self.dims[::-1], self.name)

    def assign_coords(self, **kw):
        out = self.copy()
        for d, c in kw.items():
            out.coords_list[out.dims.index(d)] = np.asarray(c)
        return out

    def sum(self, axis=None, dtype=None, out=None, **kw):
        v = self.values.sum(axis=axis, dtype=dtype)
        if axis is None:
            return DataArray(v)
        keep = [i for i in range(self.ndim) if i != axis % self.ndim]
        return DataArray(v, [self.coords_list[i] for i in keep], [self.dims[i] for i in keep])

    # ---- indexing ----------------------------------------------------------------------------------------------------
    def _indexer(self, sel):
        """dict dim -> position(s)  ->  (numpy index doing ORTHOGONAL selection, kept dims, kept coords)"""
        idx, dims, coords, arrays = [], [], [], []
        for d, c in zip(self.dims, self.coords_list):
            k = sel.get(d, slice(None))
            if isinstance(k, DataArray):
                k = k.values
            if isinstance(k, slice):
                idx.append(np.arange(len(c))[k]); dims.append(d); coords.append(c[k]); arrays.append(True)
            elif np.ndim(k) == 0:
                idx.append(int(k)); arrays.append(False)
            else:
                k = np.asarray(k)
                if k.dtype == bool:
                    k = np.nonzero(k)[0]
                idx.append(k.astype(int)); dims.append(d); coords.append(c[k.astype(int)]); arrays.append(True)
        vec = [i for i, a in zip(idx, arrays) if a]
        mesh = iter(np.ix_(*vec)) if vec else iter(())
        full = tuple(next(mesh) if a else i for i, a in zip(idx, arrays))
        return full, dims, coords

    def isel(self, sel):
        full, dims, coords = self._indexer(sel)
        return DataArray(self.values[full], coords, dims, self.name)

    def _assign(self, sel, value):
        full, dims, coords = self._indexer(sel)
        v = np.asarray(value.values if isinstance(value, DataArray) else value)
        target_shape = np.broadcast(*[np.asarray(f) for f in full]).shape if full else ()
        if v.ndim and v.shape != target_shape:
            v = v.reshape(target_shape) if v.size == int(np.prod(target_shape)) else v
        self.values[full] = v

    def _positional(self, key):
        if isinstance(key, dict):
            return key
        if not isinstance(key, tuple):
            key = (key,)
        return {self.dims[i]: k for i, k in enumerate(key)}

    def __getitem__(self, key):
        return self.isel(self._positional(key))

    def __setitem__(self, key, value):
        self._assign(self._positional(key), value)

    # ---- arithmetic: scalars and same-shaped operands only (all the reference needs) ---------------------------------
    def _wrap(self, v, other=None):
        src = self
        if isinstance(other, DataArray) and other.ndim > self.ndim:
            src = other
        if np.shape(v) == src.values.shape:
            return DataArray(v, src.coords_list, src.dims)
        return DataArray(v)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        raw = [x.values if isinstance(x, DataArray) else x for x in inputs]
        v = ufunc(*raw, **kw)
        big = max((x for x in inputs if isinstance(x, DataArray)), key=lambda x: x.ndim)
        return big._wrap(v)


def _binop(name, ufunc, reflected=False):
    def op(self, other):
        o = other.values if isinstance(other, DataArray) else other
        v = ufunc(o, self.values) if reflected else ufunc(self.values, o)
        return self._wrap(v, other)
    op.__name__ = name
    return op


for _n, _u in (("add", np.add), ("sub", np.subtract), ("mul", np.multiply), ("truediv", np.true_divide), ("pow", np.power)):
    setattr(DataArray, "__%s__" % _n, _binop(_n, _u))
    setattr(DataArray, "__r%s__" % _n, _binop("r" + _n, _u, True))
for _n, _u in (("lt", np.less), ("le", np.less_equal), ("gt", np.greater), ("ge", np.greater_equal), ("eq", np.equal),
               ("ne", np.not_equal), ("or", np.logical_or), ("and", np.logical_and)):
    setattr(DataArray, "__%s__" % _n, _binop(_n, _u))
DataArray.__neg__ = lambda self: self._wrap(-self.values)
DataArray.__abs__ = lambda self: self._wrap(abs(self.values))
