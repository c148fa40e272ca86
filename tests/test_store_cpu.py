"""CPU: the zarr directory reader / writer behind compute_residual / pcg_dds / BandWorkerPool.load_bands
(pfb_imaging_b200/store.py): v2 round trip, chunked + compressed arrays as xarray writes them, the v3 subset,
DataTree nodes, and load_band's selective reads (operators/band_worker.py:61-106)."""
import json
import os
import zlib

import numpy as np
import pytest

from pfb_imaging_b200 import store


def test_dataset_round_trip_and_append(tmp_path):
    rng = np.random.default_rng(0)
    ds = store.Dataset(attrs=dict(x0=0.1, y0=-0.2, flip_v=True, wsum=np.float64(3.5), bandid=np.int64(4)))
    ds["UVW"] = (("row", "three"), rng.standard_normal((50, 3)))
    ds["MASK"] = (("row", "chan"), rng.integers(0, 2, (50, 4)).astype(np.uint8))
    ds["PSFHAT"] = (("corr", "x_psf", "yo2"), rng.standard_normal((1, 8, 5)) + 1j * rng.standard_normal((1, 8, 5)))
    p = str(tmp_path / "band.zarr")
    ds.to_zarr(p, mode="w")
    back = store.open_zarr(p)
    assert set(back) == {"UVW", "MASK", "PSFHAT"}
    for k in ds:
        assert np.array_equal(back[k].values, ds[k].values) and back[k].values.dtype == ds[k].values.dtype
    assert back.UVW.dims == ("row", "three")
    assert back.x0 == 0.1 and back.flip_v is True and back.attrs["wsum"] == 3.5 and back.bandid == 4
    # append two arrays like compute_residual does (only RESIDUAL / MODEL are written, the rest stays)
    back["MODEL"] = (("corr", "x", "y"), np.ones((1, 6, 6)))
    back["RESIDUAL"] = (("corr", "x", "y"), np.full((1, 6, 6), 2.0))
    back[["RESIDUAL", "MODEL"]].to_zarr(p, mode="a", compress=True)
    again = store.open_zarr(p, drop_vars=["PSFHAT"])
    assert set(again) == {"UVW", "MASK", "MODEL", "RESIDUAL"} and again.RESIDUAL.values[0, 0, 0] == 2.0
    assert "MODEL" in again and "PSFHAT" not in again
    assert set(again.drop_vars("MODEL")) == {"UVW", "MASK", "RESIDUAL"}


def test_reads_chunked_compressed_arrays_as_xarray_writes_them(tmp_path):
    """zarr v2, 2-D chunks that do not divide the shape, zlib, nested and flat chunk keys, fill value."""
    a = np.arange(7 * 10, dtype="<f8").reshape(7, 10)
    for sep in (".", "/"):
        g = tmp_path / f"g{ord(sep)}"
        (g / "A").mkdir(parents=True)
        json.dump({"zarr_format": 2}, open(g / ".zgroup", "w"))
        json.dump({"foo": 1}, open(g / ".zattrs", "w"))
        meta = dict(zarr_format=2, shape=[7, 10], chunks=[4, 4], dtype="<f8", compressor={"id": "zlib", "level": 1},
                    fill_value=-1.0, order="C", filters=None, dimension_separator=sep)
        json.dump(meta, open(g / "A" / ".zarray", "w"))
        json.dump({"_ARRAY_DIMENSIONS": ["x", "y"]}, open(g / "A" / ".zattrs", "w"))
        for i in range(2):
            for j in range(3):
                if (i, j) == (1, 2):
                    continue  # a missing chunk reads as the fill value
                blk = np.full((4, 4), -1.0)
                sub = a[4 * i:4 * i + 4, 4 * j:4 * j + 4]
                blk[:sub.shape[0], :sub.shape[1]] = sub
                f = g / "A" / (f"{i}.{j}" if sep == "." else os.path.join(str(i), str(j)))
                f.parent.mkdir(parents=True, exist_ok=True)
                f.write_bytes(zlib.compress(blk.tobytes()))
        ds = store.open_zarr(str(g))
        want = a.copy()
        want[4:, 8:] = -1.0
        assert np.array_equal(ds.A.values, want) and ds.A.dims == ("x", "y") and ds.foo == 1


def test_reads_the_v3_subset(tmp_path):
    import gzip

    g = tmp_path / "v3"
    (g / "B" / "c" / "0").mkdir(parents=True)
    json.dump(dict(zarr_format=3, node_type="group", attributes={"wsum": [2.0]}), open(g / "zarr.json", "w"))
    b = np.arange(6, dtype="<i4").reshape(2, 3)
    meta = dict(zarr_format=3, node_type="array", shape=[2, 3], data_type="int32", fill_value=0,
                chunk_grid=dict(name="regular", configuration=dict(chunk_shape=[2, 3])),
                chunk_key_encoding=dict(name="default", configuration=dict(separator="/")),
                codecs=[dict(name="bytes", configuration=dict(endian="little")), dict(name="gzip", configuration=dict(level=1))],
                dimension_names=["row", "chan"], attributes={})
    json.dump(meta, open(g / "B" / "zarr.json", "w"))
    (g / "B" / "c" / "0" / "0").write_bytes(gzip.compress(b.tobytes()))
    ds = store.open_zarr(str(g))
    assert np.array_equal(ds.B.values, b) and ds.B.dims == ("row", "chan") and ds.wsum == [2.0]
    # an unsupported codec is a clear error, not garbage
    meta["codecs"].append(dict(name="crc32c"))
    json.dump(meta, open(g / "B" / "zarr.json", "w"))
    with pytest.raises(RuntimeError, match="codec"):
        store.open_zarr(str(g))


def make_dt_store(path, nband=2, npart=2, nx=32, nx_psf=48, nrow=120, nchan=3, seed=0, ncorr=1):
    """A .dt store laid out like `pfb imager` writes it: band nodes with DIRTY, partition children with the
    gridding inputs, PSFHAT and attrs (docs/wiki/imager-pipeline.md:39-62)."""
    rng = np.random.default_rng(seed)
    store.Dataset(attrs=dict(nband=nband)).to_zarr(path, mode="w")
    truth = {}
    for b in range(nband):
        node = os.path.join(path, f"band{b:04d}")
        dirty = rng.standard_normal((ncorr, nx, nx))
        bd = store.Dataset(attrs=dict(bandid=b))
        bd["DIRTY"] = (("corr", "x", "y"), dirty)
        bd.to_zarr(node)
        truth[b] = dict(dirty=dirty, parts=[])
        for p in range(npart):
            freq = np.linspace(1.0e9, 1.1e9, nchan) * (1 + 0.1 * b)
            cell = 0.2 / nx
            umax = 0.5 / cell * 299792458.0 / freq.max()
            part = dict(UVW=rng.uniform(-1, 1, (nrow, 3)) * umax * np.array([0.9, 0.9, 0.05]),
                        WEIGHT=rng.uniform(0.5, 1.5, (ncorr, nrow, nchan)),
                        MASK=(rng.uniform(size=(nrow, nchan)) > 0.1).astype(np.uint8), FREQ=freq,
                        BEAM=rng.uniform(0.7, 1.0, (ncorr, nx, nx)),
                        # the spectrum of a REAL point-spread function (Hermitian along x in the k_y = 0 / Nyquist columns)
                        PSFHAT=np.fft.rfft2(rng.standard_normal((ncorr, nx_psf, nx_psf)) *
                                            np.outer(np.hanning(nx_psf), np.hanning(nx_psf))[None], axes=(1, 2)),
                        VIS=rng.standard_normal((ncorr, nrow, nchan)) + 0j)
            pd = store.Dataset(attrs=dict(wsum=[float(part["WEIGHT"][c][part["MASK"] > 0].sum()) for c in range(ncorr)],
                                          l0=0.0, m0=0.0, cell_rad=cell))
            dims = dict(UVW=("row", "three"), WEIGHT=("corr", "row", "chan"), MASK=("row", "chan"), FREQ=("chan",),
                        BEAM=("corr", "x", "y"), PSFHAT=("corr", "x_psf", "yo2"), VIS=("corr", "row", "chan"))
            for k, v in part.items():
                pd[k] = (dims[k], v)
            pd.to_zarr(os.path.join(node, f"part{p:04d}"), compress=(p % 2 == 1))
            part["cell"] = cell
            truth[b]["parts"].append(part)
    return truth


def test_datatree_nodes_and_load_band(tmp_path):
    from pfb_imaging_b200.band_worker import _BandWorkerImpl

    path = str(tmp_path / "img.dt")
    truth = make_dt_store(path)
    tree = store.open_datatree(path)
    assert sorted(tree.children) == ["band0000", "band0001"]
    band = tree["band0001"]
    assert sorted(band.children) == ["part0000", "part0001"] and band.ds.bandid == 1
    assert np.array_equal(tree["band0001/part0001"].ds.FREQ.values, truth[1]["parts"][1]["FREQ"])
    w = _BandWorkerImpl(1)
    w.load_band(path, "band0001")
    assert np.array_equal(w._dirty, truth[1]["dirty"])
    assert len(w._parts) == 2 and len(w._hess_parts) == 2
    for p in range(2):
        t = truth[1]["parts"][p]
        assert set(w._parts[p]) == {"UVW", "WEIGHT", "MASK", "FREQ", "BEAM"}  # selective load: no VIS, no PSFHAT
        assert np.array_equal(w._parts[p].UVW.values, t["UVW"]) and w._parts[p].attrs["l0"] == 0.0
        assert np.array_equal(w._hess_parts[p]["psfhat"], np.abs(t["PSFHAT"]))
        assert np.allclose(w._hess_parts[p]["wsum"], [t["WEIGHT"][0][t["MASK"] > 0].sum()])
    with pytest.raises(RuntimeError, match="load_band"):
        _BandWorkerImpl(1).init_hess(None, 32, 32, 48, 48, 0.1, None)
