"""On-disk catalogs -> the column dictionaries the model classes and ``capi.Catalog`` take.

The reference loads its inputs with ArviZ / xarray (``gwinferno/pipeline/utils.py:51-96``:
``az.from_netcdf(file)`` -> ``pe_data.posteriors[event, param, samples]``, ``inj_data.injections[param, injection]``,
attributes ``total_generated`` and ``analysis_time``; the writer is ``preprocess/data_collection.py:203-207``) and its
tests with ``xr.load_dataset`` (``tests/inference_test.py:74-82``: one ``[param, sample]`` variable per event).  Neither
library (nor h5py / netCDF4) exists in this environment, so this module carries its own reader:

* :func:`read_netcdf3` -- the NetCDF classic container (CDF-1, CDF-2 "64-bit offset", CDF-5) parsed with NumPy only.
  This IS the format of the reference's vendored PE fixture (``tests/data/xarray_GWTC3_BBH_69evs_..._nospin.h5`` is a
  CDF-2 file despite its extension) and what ``xarray.Dataset.to_netcdf(format="NETCDF3_64BIT")`` writes.
* :func:`load_pe_dataset` -- the fixture layout (a variable per event) -> ``{param: (E, S) float64}``.
* :func:`load_pe_and_injections_as_dict` -- same return value as the reference function, from a *flattened* catalog
  file: the two idata groups side by side in one classic file (``posteriors[event, param, sample]``,
  ``injections[inj_param, injection]``, global attributes ``total_generated`` / ``analysis_time``) or the ``.npz`` twin
  written by :func:`save_catalog`.  NetCDF-4 / HDF5 idata files (groups, chunked + deflated variables) are not parsed
  here: convert them once where ArviZ exists with ``tools/idata_to_gwi.py`` (a dozen lines of xarray).
* :func:`write_netcdf3` / :func:`save_catalog` -- writers for the flattened layout (fixed-size variables only).

Everything here is host-side ingest that runs once before the sampler starts; the hot path only sees the resulting
column arrays.
"""

import os
import struct

import numpy as np

__all__ = ["read_netcdf3", "write_netcdf3", "load_pe_dataset", "load_pe_and_injections_as_dict", "save_catalog", "NetCDFError"]

# nc_type -> (NumPy big-endian dtype, item size)
_NC_TYPES = {1: (">i1", 1), 2: ("S1", 1), 3: (">i2", 2), 4: (">i4", 4), 5: (">f4", 4), 6: (">f8", 8),
             7: (">u1", 1), 8: (">u2", 2), 9: (">u4", 4), 10: (">i8", 8), 11: (">u8", 8)}
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C


class NetCDFError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.wide = False  # CDF-5: 64-bit counts

    def u32(self):
        (v,) = struct.unpack_from(">I", self.b, self.p)
        self.p += 4
        return v

    def u64(self):
        (v,) = struct.unpack_from(">Q", self.b, self.p)
        self.p += 8
        return v

    def count(self):
        return self.u64() if self.wide else self.u32()

    def name(self):
        n = self.count()
        s = bytes(self.b[self.p : self.p + n]).decode("utf-8")
        self.p += (n + 3) & ~3
        return s

    def values(self, nc_type, n):
        if nc_type not in _NC_TYPES:
            raise NetCDFError(f"unknown nc_type {nc_type}")
        dt, size = _NC_TYPES[nc_type]
        raw = self.b[self.p : self.p + n * size]
        self.p += (n * size + 3) & ~3
        if nc_type == 2:
            return bytes(raw).decode("utf-8", errors="replace").rstrip("\x00")
        return np.frombuffer(raw, dtype=dt).astype(np.dtype(dt).newbyteorder("="))

    def attributes(self):
        tag = self.u32()
        n = self.count()
        if tag == 0 and n == 0:
            return {}
        if tag != _NC_ATTRIBUTE:
            raise NetCDFError("attribute list expected")
        out = {}
        for _ in range(n):
            k = self.name()
            t = self.u32()
            m = self.count()
            v = self.values(t, m)
            out[k] = v if isinstance(v, str) or v.size != 1 else v[0]
        return out


def read_netcdf3(path, variables=None):
    """Parse a NetCDF classic file.  Returns ``(dims, attrs, vars)``: ``dims`` = {name: length} (0 = the record
    dimension, whose current length is ``attrs_['__numrecs__']``), ``attrs`` = global attributes, ``vars`` =
    {name: (dim names, attributes, native-endian array)}.  ``variables`` restricts which arrays are materialised."""
    buf = np.fromfile(path, dtype=np.uint8) if isinstance(path, (str, os.PathLike)) else np.frombuffer(path, dtype=np.uint8)
    mv = memoryview(buf)
    if len(buf) < 8 or bytes(mv[:3]) != b"CDF":
        if bytes(mv[:4]) == b"\x89HDF":
            raise NetCDFError("NetCDF-4 / HDF5 container: convert it once with tools/idata_to_gwi.py (needs xarray) -- this reader parses the classic formats")
        raise NetCDFError("not a NetCDF classic file")
    version = int(buf[3])
    if version not in (1, 2, 5):
        raise NetCDFError(f"unsupported NetCDF classic version {version}")
    r = _Reader(mv)
    r.p = 4
    r.wide = version == 5
    numrecs = r.count()
    # dimensions
    tag, n = r.u32(), r.count()
    dim_names, dim_lens = [], []
    if not (tag == 0 and n == 0):
        if tag != _NC_DIMENSION:
            raise NetCDFError("dimension list expected")
        for _ in range(n):
            dim_names.append(r.name())
            dim_lens.append(r.count())
    gatts = r.attributes()
    # variables
    tag, n = r.u32(), r.count()
    heads = []
    if not (tag == 0 and n == 0):
        if tag != _NC_VARIABLE:
            raise NetCDFError("variable list expected")
        for _ in range(n):
            name = r.name()
            nd = r.count()
            dimids = [r.count() for _ in range(nd)]
            vatts = r.attributes()
            t = r.u32()
            vsize = r.count()
            begin = r.u32() if version == 1 else r.u64()
            heads.append((name, dimids, vatts, t, vsize, begin))
    is_rec = lambda dimids: len(dimids) > 0 and dim_lens[dimids[0]] == 0
    rec_vars = [h for h in heads if is_rec(h[1])]
    recsize = sum(h[4] for h in rec_vars)
    out = {}
    for name, dimids, vatts, t, vsize, begin in heads:
        dnames = tuple(dim_names[d] for d in dimids)
        if variables is not None and name not in variables:
            continue
        if t not in _NC_TYPES:
            raise NetCDFError(f"variable {name}: unknown nc_type {t}")
        dt, size = _NC_TYPES[t]
        if is_rec(dimids):
            shape1 = tuple(dim_lens[d] for d in dimids[1:])
            n1 = int(np.prod(shape1, dtype=np.int64)) if shape1 else 1
            stride = recsize if len(rec_vars) > 1 else n1 * size  # a lone record variable is stored without padding
            arr = np.empty((numrecs,) + shape1, dtype=np.dtype(dt).newbyteorder("="))
            for k in range(numrecs):
                o = begin + k * stride
                arr[k] = np.frombuffer(mv[o : o + n1 * size], dtype=dt).reshape(shape1)
        else:
            shape = tuple(dim_lens[d] for d in dimids)
            cnt = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if begin + cnt * size > len(buf):
                raise NetCDFError(f"variable {name} runs past the end of the file")
            arr = np.frombuffer(mv[begin : begin + cnt * size], dtype=dt).reshape(shape)
            if t != 2:
                arr = arr.astype(np.dtype(dt).newbyteorder("="))
        out[name] = (dnames, vatts, arr)
    dims = dict(zip(dim_names, dim_lens))
    gatts = dict(gatts)
    gatts["__numrecs__"] = numrecs
    return dims, gatts, out


def _char_rows_to_str(arr):
    """``[n, strlen]`` array of single characters (how the classic format stores string coordinates) -> list of str."""
    a = np.asarray(arr)
    if a.ndim == 1:
        return [b"".join(a.tolist()).decode("utf-8").rstrip("\x00 ")]
    return [b"".join(row.tolist()).decode("utf-8").rstrip("\x00 ") for row in a]


def load_pe_dataset(path, n_samples=None, rng=None):
    """The reference's PE fixture layout (``tests/inference_test.py:74-82``): one ``[param, sample]`` variable per
    event, a ``param`` character coordinate.  Returns ``(pedict, event_names, param_names)`` with ``pedict[param]`` an
    ``(E, S)`` float64 array.  ``n_samples``: draw that many samples per event without replacement, as the reference's
    tests do (``rng``: a ``numpy.random.Generator``; the reference uses the unseeded global state)."""
    dims, _, var = read_netcdf3(path)
    if "param" not in var:
        raise NetCDFError("no `param` coordinate in the file")
    params = _char_rows_to_str(var["param"][2])
    events = [k for k, (d, _, _) in var.items() if d == ("param", "sample")]
    if not events:
        raise NetCDFError("no [param, sample] event variables in the file")
    cube = np.stack([np.asarray(var[e][2], dtype=np.float64) for e in events])  # (E, P, S)
    if n_samples is not None:
        rng = rng or np.random.default_rng()
        idx = rng.choice(cube.shape[2], size=int(n_samples), replace=False)
        cube = cube[:, :, idx]
    pedict = {p: np.ascontiguousarray(cube[:, i, :]) for i, p in enumerate(params)}
    return pedict, events, params


def load_pe_and_injections_as_dict(file, ignore=None):
    """Same return value as ``gwinferno.pipeline.utils.load_pe_and_injections_as_dict`` (utils.py:51-96):
    ``(pedict, injdict, constants, param_names)`` with ``constants = {total_inj, obs_time, nObs}``.
    ``file``: a flattened classic NetCDF catalog or the ``.npz`` written by :func:`save_catalog`."""
    if str(file).endswith(".npz"):
        z = np.load(file, allow_pickle=False)
        post, inj = z["posteriors"], z["injections"]
        params, inj_params = [str(s) for s in z["param"]], [str(s) for s in z["inj_param"]]
        events = [str(s) for s in z["event"]]
        total_inj, obs_time = float(z["total_generated"]), float(z["analysis_time"])
    else:
        dims, att, var = read_netcdf3(file)
        for need in ("posteriors", "injections", "param"):
            if need not in var:
                raise NetCDFError(f"`{need}` missing: not a flattened catalog file")
        post = np.asarray(var["posteriors"][2], dtype=np.float64)
        inj = np.asarray(var["injections"][2], dtype=np.float64)
        params = _char_rows_to_str(var["param"][2])
        inj_params = _char_rows_to_str(var["inj_param"][2]) if "inj_param" in var else params
        events = _char_rows_to_str(var["event"][2]) if "event" in var else [str(i) for i in range(post.shape[0])]
        total_inj, obs_time = float(att["total_generated"]), float(att["analysis_time"])
    if post.ndim != 3 or post.shape[1] != len(params) or inj.ndim != 2 or inj.shape[0] != len(inj_params):
        raise NetCDFError("posteriors must be [event, param, sample] and injections [param, injection]")
    if ignore is not None:
        sel = np.array([e not in set(ignore) for e in events])
        post = post[sel]
    pedict = {k: np.ascontiguousarray(post[:, i, :], dtype=np.float64) for i, k in enumerate(params)}
    injdict = {k: np.ascontiguousarray(inj[i], dtype=np.float64) for i, k in enumerate(inj_params)}
    constants = {"total_inj": total_inj, "obs_time": obs_time, "nObs": int(post.shape[0])}
    return pedict, injdict, constants, list(params)


# ---- writers ---------------------------------------------------------------------------------------
def _pad4(b):
    return b + b"\x00" * (-len(b) % 4)


def _nc_name(s):
    e = s.encode("utf-8")
    return struct.pack(">I", len(e)) + _pad4(e)


def _nc_att(k, v):
    if isinstance(v, str):
        e = v.encode("utf-8")
        return _nc_name(k) + struct.pack(">II", 2, len(e)) + _pad4(e)
    a = np.atleast_1d(np.asarray(v))
    t = 6 if a.dtype.kind == "f" else 4
    return _nc_name(k) + struct.pack(">II", t, a.size) + _pad4(a.astype(_NC_TYPES[t][0]).tobytes())


def write_netcdf3(path, dims, attrs, variables):
    """Write a CDF-2 (64-bit offset) file with fixed-size variables.  ``dims``: {name: length};
    ``variables``: {name: (dim names, array)} -- float64 / float32 / int32 arrays, or ``S1`` character arrays."""
    dim_names = list(dims)
    head = b"CDF\x02" + struct.pack(">I", 0)
    head += struct.pack(">II", _NC_DIMENSION, len(dim_names)) if dim_names else struct.pack(">II", 0, 0)
    for d in dim_names:
        head += _nc_name(d) + struct.pack(">I", int(dims[d]))
    head += (struct.pack(">II", _NC_ATTRIBUTE, len(attrs)) + b"".join(_nc_att(k, v) for k, v in attrs.items())) if attrs else struct.pack(">II", 0, 0)
    entries, blobs = [], []
    for name, (dnames, arr) in variables.items():
        a = np.asarray(arr)
        t = {"f8": 6, "f4": 5, "i4": 4, "i2": 3, "i1": 1}.get(a.dtype.str[1:], 2 if a.dtype.kind == "S" else None)
        if t is None:
            raise NetCDFError(f"variable {name}: dtype {a.dtype} is not supported by the classic format writer")
        if tuple(int(dims[d]) for d in dnames) != a.shape:
            raise NetCDFError(f"variable {name}: shape {a.shape} does not match its dimensions")
        raw = _pad4(a.astype(_NC_TYPES[t][0]).tobytes() if t != 2 else a.tobytes())
        entries.append((name, [dim_names.index(d) for d in dnames], t, len(raw)))
        blobs.append(raw)
    var_head_size = 8 + sum(len(_nc_name(n)) + 4 + 4 * len(ids) + 8 + 4 + 4 + 8 for n, ids, _, _ in entries)
    offset = len(head) + var_head_size
    vh = struct.pack(">II", _NC_VARIABLE, len(entries)) if entries else struct.pack(">II", 0, 0)
    for (name, ids, t, size) in entries:
        vh += _nc_name(name) + struct.pack(">I", len(ids)) + b"".join(struct.pack(">I", i) for i in ids) + struct.pack(">II", 0, 0)
        vh += struct.pack(">II", t, min(size, 0xFFFFFFFF)) + struct.pack(">Q", offset)
        offset += size
    with open(path, "wb") as f:
        f.write(head + vh)
        for b in blobs:
            f.write(b)


def _char_rows(strings):
    n = max(1, max((len(s.encode("utf-8")) for s in strings), default=1))
    out = np.zeros((len(strings), n), dtype="S1")
    for i, s in enumerate(strings):
        e = s.encode("utf-8")
        out[i, : len(e)] = np.frombuffer(e, dtype="S1")
    return out


def save_catalog(path, pedict, injdict, total_inj, obs_time, events=None):
    """Write the flattened catalog (``.npz`` or classic NetCDF by extension) that
    :func:`load_pe_and_injections_as_dict` reads."""
    params, inj_params = list(pedict), list(injdict)
    post = np.stack([np.asarray(pedict[k], dtype=np.float64) for k in params], axis=1)  # (E, P, S)
    inj = np.stack([np.asarray(injdict[k], dtype=np.float64) for k in inj_params])
    events = list(events) if events is not None else [f"event{i}" for i in range(post.shape[0])]
    if str(path).endswith(".npz"):
        np.savez(path, posteriors=post, injections=inj, param=np.array(params), inj_param=np.array(inj_params), event=np.array(events),
                 total_generated=float(total_inj), analysis_time=float(obs_time))
        return
    pc, ic, ec = _char_rows(params), _char_rows(inj_params), _char_rows(events)
    dims = {"event": post.shape[0], "param": post.shape[1], "sample": post.shape[2], "inj_param": inj.shape[0], "injection": inj.shape[1],
            "param_strlen": pc.shape[1], "inj_param_strlen": ic.shape[1], "event_strlen": ec.shape[1]}
    write_netcdf3(path, dims, {"total_generated": float(total_inj), "analysis_time": float(obs_time)},
                  {"posteriors": (("event", "param", "sample"), post), "injections": (("inj_param", "injection"), inj),
                   "param": (("param", "param_strlen"), pc), "inj_param": (("inj_param", "inj_param_strlen"), ic), "event": (("event", "event_strlen"), ec)})
