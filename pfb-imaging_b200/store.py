"""Minimal reader / writer of the zarr directory stores pfb-imaging keeps its per-band data in:
the ``.dds`` datasets ``compute_residual`` / ``pcg_dds`` open (``/root/reference/src/pfb_imaging/operators/
gridder.py:1046-1059``, ``opt/pcg.py:473-474``) and the ``.dt`` DataTree a band worker loads its band from
(``operators/band_worker.py:61-106``: node ``DIRTY`` + children ``UVW / WEIGHT / MASK / FREQ / BEAM / PSFHAT``,
attrs ``wsum, l0, m0, ...``).

xarray and zarr are not importable in this image, so the layout is read directly: zarr v2 (``.zgroup`` /
``.zarray`` / ``.zattrs``, chunk files ``i.j.k`` or ``i/j/k``) and the v3 subset xarray writes (``zarr.json``,
``bytes`` codec, chunk keys ``c/i/j``).  Chunk compression: none, zlib, gzip from the standard library; zstd and
blosc when ``zstandard`` / ``numcodecs`` happen to be importable, a clear error otherwise.  What comes back is
duck-typed like the xarray objects the reference code touches: ``ds.UVW.values``, ``ds.attrs`` (also as
attributes: ``ds.flip_u``), ``name in ds``, ``ds[name] = (dims, array)``, ``ds[[names]]``, ``ds.drop_vars``,
``ds.assign``, ``ds.to_zarr(path, mode="a")``; ``open_datatree(path)[node].ds / .children``.
Host logic only: arrays are numpy, the device never sees this module.
"""

from __future__ import annotations

import json
import os
import zlib

import numpy as np


# ---------------------------------------------------------------------------
# chunk codecs
# ---------------------------------------------------------------------------
def _decompress(cfg, raw: bytes) -> bytes:
    if cfg is None:
        return raw
    cid = cfg.get("id") or cfg.get("name")
    if cid == "zlib":
        return zlib.decompress(raw)
    if cid == "gzip":
        import gzip

        return gzip.decompress(raw)
    if cid == "zstd":
        try:
            import zstandard

            return zstandard.ZstdDecompressor().decompress(raw)
        except ImportError:
            try:
                from compression import zstd  # Python >= 3.14

                return zstd.decompress(raw)
            except ImportError as e:
                raise RuntimeError("zstd-compressed chunks need the `zstandard` module, which is not installed") from e
    if cid == "blosc":
        try:
            import numcodecs

            return numcodecs.Blosc().decode(raw)
        except ImportError as e:
            raise RuntimeError("blosc-compressed chunks need `numcodecs`, which is not installed; re-write the "
                               "store with compressor=None or zlib") from e
    raise RuntimeError(f"unsupported chunk codec {cid!r}")


class Var:
    """One array of a group: shape / dtype / dims from the metadata, ``.values`` reads the chunks."""

    def __init__(self, path=None, values=None, dims=None, attrs=None):
        self._path, self._values = path, None if values is None else np.asarray(values)
        self.dims = tuple(dims) if dims is not None else None
        self.attrs = dict(attrs or {})
        if path is not None:
            self._read_meta()

    def _read_meta(self):
        p = self._path
        if os.path.exists(os.path.join(p, ".zarray")):
            m = json.load(open(os.path.join(p, ".zarray")))
            self._v = 2
            self.shape, self._chunks = tuple(m["shape"]), tuple(m["chunks"])
            self.dtype = np.dtype(m["dtype"])
            self._comp, self._fill = m.get("compressor"), m.get("fill_value")
            self._order = m.get("order", "C")
            self._sep = m.get("dimension_separator", ".")
            if m.get("filters"):
                raise RuntimeError(f"{p}: zarr filters are not supported")
            za = os.path.join(p, ".zattrs")
            if os.path.exists(za):
                self.attrs = json.load(open(za))
            if self.dims is None:
                self.dims = tuple(self.attrs.get("_ARRAY_DIMENSIONS", [f"dim_{i}" for i in range(len(self.shape))]))
        else:
            m = json.load(open(os.path.join(p, "zarr.json")))
            if m.get("node_type") != "array":
                raise RuntimeError(f"{p} is not a zarr array")
            self._v = 3
            self.shape = tuple(m["shape"])
            self._chunks = tuple(m["chunk_grid"]["configuration"]["chunk_shape"])
            self.dtype = np.dtype(m["data_type"])
            self._fill = m.get("fill_value")
            self._order = "C"
            self._sep = m.get("chunk_key_encoding", {}).get("configuration", {}).get("separator", "/")
            self._comp = None
            for c in m.get("codecs", []):
                if c["name"] == "bytes":
                    if c.get("configuration", {}).get("endian", "little") != "little":
                        self.dtype = self.dtype.newbyteorder(">")
                elif c["name"] in ("gzip", "zstd", "blosc", "zlib"):
                    self._comp = dict(c.get("configuration", {}), id=c["name"])
                else:
                    raise RuntimeError(f"{p}: zarr v3 codec {c['name']!r} is not supported")
            self.attrs = m.get("attributes", {})
            if self.dims is None:
                self.dims = tuple(m.get("dimension_names") or [f"dim_{i}" for i in range(len(self.shape))])

    @property
    def size(self):
        return int(np.prod(self.shape)) if self._values is None else self._values.size

    @property
    def values(self):
        if self._values is not None:
            return self._values
        shape, chunks = self.shape, self._chunks
        fill = 0 if self._fill in (None, "NaN") and self.dtype.kind != "f" else (np.nan if self._fill in (None, "NaN") else self._fill)
        if self.dtype.kind == "c" and isinstance(fill, (list, tuple)):
            fill = complex(*fill)
        out = np.full(shape, fill, dtype=self.dtype) if shape else np.zeros((), self.dtype)
        if not shape:  # scalar array
            key = "0" if self._v == 2 else "c"
            f = os.path.join(self._path, key)
            if os.path.exists(f):
                out[...] = np.frombuffer(_decompress(self._comp, open(f, "rb").read()), dtype=self.dtype)[0]
            self._values = out
            return out
        nchunks = [max(1, -(-s // c)) for s, c in zip(shape, chunks)]
        for idx in np.ndindex(*nchunks):
            key = self._sep.join(str(i) for i in idx)
            f = os.path.join(self._path, key if self._v == 2 else os.path.join("c", *[str(i) for i in idx]))
            if not os.path.exists(f):
                continue
            buf = np.frombuffer(_decompress(self._comp, open(f, "rb").read()), dtype=self.dtype)
            blk = buf.reshape(chunks, order=self._order)
            sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
            out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        self._values = out
        return out

    def load(self):
        self.values
        return self


def _write_array(path, arr, dims, attrs=None, compress=False):
    arr = np.ascontiguousarray(arr)
    os.makedirs(path, exist_ok=True)
    for f in os.listdir(path):  # replace a previous version of the array
        fp = os.path.join(path, f)
        if os.path.isfile(fp):
            os.remove(fp)
    meta = dict(zarr_format=2, shape=list(arr.shape), chunks=list(arr.shape) if arr.ndim else [], dtype=arr.dtype.str,
                compressor=dict(id="zlib", level=1) if compress else None, fill_value=None, order="C", filters=None)
    json.dump(meta, open(os.path.join(path, ".zarray"), "w"))
    at = dict(attrs or {})
    at["_ARRAY_DIMENSIONS"] = list(dims)
    json.dump(at, open(os.path.join(path, ".zattrs"), "w"))
    raw = arr.tobytes()
    key = ".".join("0" for _ in arr.shape) if arr.ndim else "0"
    open(os.path.join(path, key), "wb").write(zlib.compress(raw, 1) if compress else raw)


def _jsonable(v):
    if isinstance(v, np.generic):
        return v.item()
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    return v


class Dataset:
    """The slice of the xarray.Dataset interface the reference's operator-level code uses."""

    def __init__(self, variables=None, attrs=None, path=None):
        object.__setattr__(self, "_vars", dict(variables or {}))
        object.__setattr__(self, "attrs", dict(attrs or {}))
        object.__setattr__(self, "_path", path)

    # -- access ------------------------------------------------------------------------------------------------
    def __getattr__(self, name):
        v = object.__getattribute__(self, "_vars")
        if name in v:
            return v[name]
        a = object.__getattribute__(self, "attrs")
        if name in a:
            return a[name]
        raise AttributeError(name)

    def __contains__(self, name):
        return name in self._vars

    def __iter__(self):
        return iter(self._vars)

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self._vars[k] for k in key}, self.attrs, self._path)
        return self._vars[key]

    def __setitem__(self, name, value):
        if isinstance(value, tuple) and len(value) == 2:
            dims, arr = value
            self._vars[name] = Var(values=arr, dims=dims)
        elif isinstance(value, Var):
            self._vars[name] = value
        else:
            arr = np.asarray(value)
            self._vars[name] = Var(values=arr, dims=[f"dim_{i}" for i in range(arr.ndim)])

    def assign(self, **kw):
        out = Dataset(self._vars, self.attrs, self._path)
        for k, v in kw.items():
            out[k] = v
        return out

    def drop_vars(self, names):
        names = [names] if isinstance(names, str) else list(names)
        return Dataset({k: v for k, v in self._vars.items() if k not in names}, self.attrs, self._path)

    def load(self):
        for v in self._vars.values():
            v.load()
        return self

    @property
    def data_vars(self):
        return self._vars

    # -- write ---------------------------------------------------------------------------------------------------
    def to_zarr(self, path, mode="a", compress=False):
        """Write / replace this dataset's arrays and attrs in the group at `path` (zarr v2); arrays the group
        already holds and this dataset does not are left alone (mode "a")."""
        if mode == "w" and os.path.isdir(path):
            import shutil

            shutil.rmtree(path)
        os.makedirs(path, exist_ok=True)
        zg = os.path.join(path, ".zgroup")
        if not os.path.exists(zg) and not os.path.exists(os.path.join(path, "zarr.json")):
            json.dump(dict(zarr_format=2), open(zg, "w"))
        za = os.path.join(path, ".zattrs")
        old = json.load(open(za)) if os.path.exists(za) else {}
        old.update(_jsonable(self.attrs))
        json.dump(old, open(za, "w"))
        for name, v in self._vars.items():
            _write_array(os.path.join(path, name), v.values, v.dims or [f"dim_{i}" for i in range(v.values.ndim)],
                         {k: val for k, val in v.attrs.items() if k != "_ARRAY_DIMENSIONS"}, compress)
        return self


def _is_array(p):
    if os.path.exists(os.path.join(p, ".zarray")):
        return True
    zj = os.path.join(p, "zarr.json")
    return os.path.exists(zj) and json.load(open(zj)).get("node_type") == "array"


def _is_group(p):
    if os.path.exists(os.path.join(p, ".zgroup")):
        return True
    zj = os.path.join(p, "zarr.json")
    return os.path.exists(zj) and json.load(open(zj)).get("node_type") == "group"


def _group_attrs(p):
    za = os.path.join(p, ".zattrs")
    if os.path.exists(za):
        return json.load(open(za))
    zj = os.path.join(p, "zarr.json")
    if os.path.exists(zj):
        return json.load(open(zj)).get("attributes", {})
    return {}


def open_zarr(path, drop_vars=None) -> Dataset:
    """One zarr group as a Dataset (what ``xds_from_list([path])[0]`` hands the reference code)."""
    if not _is_group(path):
        raise FileNotFoundError(f"{path} is not a zarr group")
    drop = set(drop_vars or [])
    vs = {}
    for name in sorted(os.listdir(path)):
        p = os.path.join(path, name)
        if os.path.isdir(p) and _is_array(p) and name not in drop:
            vs[name] = Var(p)
    return Dataset(vs, _group_attrs(path), path)


def xds_from_list(paths, nthreads=1, drop_vars=None):
    return [open_zarr(p, drop_vars=drop_vars) for p in paths]


class TreeNode:
    """A node of a DataTree store: ``.ds`` (its own arrays), ``.children`` (name -> node), ``node[name]``."""

    def __init__(self, path):
        self.path = path
        self.ds = open_zarr(path)
        self.children = {n: None for n in sorted(os.listdir(path))
                         if os.path.isdir(os.path.join(path, n)) and _is_group(os.path.join(path, n))}

    def __getitem__(self, name):
        node = self
        for part in str(name).strip("/").split("/"):
            if part not in node.children:
                raise KeyError(name)
            if node.children[part] is None:
                node.children[part] = TreeNode(os.path.join(node.path, part))
            node = node.children[part]
        return node

    @property
    def attrs(self):
        return self.ds.attrs


def open_datatree(path) -> TreeNode:
    return TreeNode(path)
