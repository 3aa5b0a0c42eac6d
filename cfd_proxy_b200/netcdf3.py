"""NetCDF-3 "classic" (CDF-1 / CDF-2) writer and reader in numpy.

Writes the per-domain mesh files of the F6 schema the reference loader reads
(reference: src/solver_data.c:98-144, src/comm_data.c:79-112, file name
``<PREFIX>_domain_<rank>_lvl_<L>`` from src/hybrid.f6.c:57-62).  There is no libnetcdf /
netCDF4 in this image; the container format is simple enough to emit directly
(SURVEY.md Appendix A).  The reader is only used by tests to cross-check the C loader.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

_NC_INT, _NC_DOUBLE = 4, 6
_TAG_DIM, _TAG_VAR = 0x0A, 0x0B


def _name(s: str) -> bytes:
    b = s.encode()
    return struct.pack(">I", len(b)) + b + b"\0" * ((-len(b)) % 4)


def write_cdf(path, dims: "OrderedDict[str,int]", variables: "OrderedDict[str,tuple]", version: int = 2) -> None:
    """variables: name -> (tuple of dim names, ndarray int32|float64)."""
    assert version in (1, 2)
    dim_ids = {n: i for i, n in enumerate(dims)}
    off_fmt = ">I" if version == 1 else ">Q"
    # header size is independent of the offsets' values, so lay it out once with zeros
    def header(begins):
        h = b"CDF" + bytes([version]) + struct.pack(">I", 0)
        h += struct.pack(">II", _TAG_DIM, len(dims)) if dims else struct.pack(">II", 0, 0)
        for n, ln in dims.items():
            h += _name(n) + struct.pack(">I", ln)
        h += struct.pack(">II", 0, 0)  # no global attributes
        h += struct.pack(">II", _TAG_VAR, len(variables)) if variables else struct.pack(">II", 0, 0)
        for (n, (vd, arr)), beg in zip(variables.items(), begins):
            ty = _NC_INT if arr.dtype.kind in "iu" else _NC_DOUBLE
            nbytes = arr.size * (4 if ty == _NC_INT else 8)
            vsize = min((nbytes + 3) & ~3, 0xFFFFFFFF)
            h += _name(n) + struct.pack(">I", len(vd))
            for d in vd:
                h += struct.pack(">I", dim_ids[d])
            h += struct.pack(">II", 0, 0)  # no attributes
            h += struct.pack(">II", ty, vsize) + struct.pack(off_fmt, beg)
        return h

    hlen = len(header([0] * len(variables)))
    begins, pos = [], hlen
    for n, (vd, arr) in variables.items():
        expect = int(np.prod([dims[d] for d in vd])) if vd else 1
        assert arr.size == expect, f"{n}: {arr.size} elements, dims say {expect}"
        begins.append(pos)
        nbytes = arr.size * (4 if arr.dtype.kind in "iu" else 8)
        pos += (nbytes + 3) & ~3
    if version == 1:
        assert pos < 2**31, "CDF-1 offsets overflow; use version=2"
    with open(path, "wb") as f:
        f.write(header(begins))
        for n, (vd, arr) in variables.items():
            if arr.dtype.kind in "iu":
                f.write(np.ascontiguousarray(arr, dtype=">i4").tobytes())
            else:
                f.write(np.ascontiguousarray(arr, dtype=">f8").tobytes())


def read_cdf(path):
    """Return (dims, variables) of a CDF-1/2 file (int32/float64 variables only)."""
    buf = open(path, "rb").read()
    assert buf[:3] == b"CDF" and buf[3] in (1, 2)
    wide = buf[3] == 2
    pos = 8

    def u32():
        nonlocal pos
        v = struct.unpack_from(">I", buf, pos)[0]
        pos += 4
        return v

    def name():
        nonlocal pos
        n = u32()
        s = buf[pos:pos + n].decode()
        pos += (n + 3) & ~3
        return s

    dims = OrderedDict()
    tag, n = u32(), u32()
    for _ in range(n if tag == _TAG_DIM else 0):
        k = name()
        dims[k] = u32()
    tag, n = u32(), u32()
    assert tag == 0 and n == 0, "attributes not supported by this test reader"
    out = OrderedDict()
    tag, n = u32(), u32()
    dl = list(dims.values())
    for _ in range(n if tag == _TAG_VAR else 0):
        k = name()
        nd = u32()
        shape = [dl[u32()] for _ in range(nd)]
        a_tag, a_n = u32(), u32()
        assert a_tag == 0 and a_n == 0
        ty, _vsize = u32(), u32()
        if wide:
            beg = struct.unpack_from(">Q", buf, pos)[0]
            pos += 8
        else:
            beg = u32()
        cnt = int(np.prod(shape)) if shape else 1
        dt = ">i4" if ty == _NC_INT else ">f8"
        out[k] = np.frombuffer(buf, dtype=dt, count=cnt, offset=beg).reshape(shape).astype(dt[1:])
    return dims, out


def write_domain_file(path, dom: dict, version: int = 2) -> None:
    """dom: arrays of one mesh domain (see mesh.gen_domain) -> F6-schema NetCDF file."""
    nall, nown, nadd = int(dom["nall"]), int(dom["nown"]), int(dom["nadd"])
    nd = int(dom["ndomains"])
    dims = OrderedDict()
    dims["ncolors"] = 1
    dims["nfaces"] = int(dom["nfaces"])
    dims["nownpoints"] = nown
    dims["nallpoints"] = nall
    dims["ndomains"] = nd
    dims["two"] = 2
    dims["three"] = 3
    v = OrderedDict()
    v["fpoint"] = (("nfaces", "two"), np.asarray(dom["fpoint"], dtype=np.int32))
    v["fnormal"] = (("nfaces", "three"), np.asarray(dom["fnormal"], dtype=np.float64))
    v["pvolume"] = (("nallpoints",), np.asarray(dom["pvolume"], dtype=np.float64))
    # simplest legal colouring; the reference reads and discards it (threads.c:748-749)
    v["fcolor_npoints"] = (("ncolors",), np.array([nall], dtype=np.int32))
    v["fcolor_points"] = (("nallpoints",), np.arange(nall, dtype=np.int32))
    if nd > 1:
        dims["naddpoints"] = nadd
        dims["ncommdomains"] = int(dom["ncommdomains"])
        v["commpartner"] = (("ncommdomains",), np.asarray(dom["commpartner"], dtype=np.int32))
        v["sendcount"] = (("ndomains",), np.asarray(dom["sendcount"], dtype=np.int32))
        v["recvcount"] = (("ndomains",), np.asarray(dom["recvcount"], dtype=np.int32))
        v["addpoint_owner"] = (("naddpoints",), np.asarray(dom["addpoint_owner"], dtype=np.int32))
        v["addpoint_idx"] = (("naddpoints",), np.asarray(dom["addpoint_idx"], dtype=np.int32))
    write_cdf(path, dims, v, version=version)
