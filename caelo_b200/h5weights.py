"""Minimal read-only HDF5 walker for the Keras 2.3 ``.h5`` weight files CAE-LO ships.

The reference loads its two inference networks with ``keras.models.load_model``
(/root/reference/Match.py:313,324; PoseEstimation.py:73; BatchPreprocess.py:168).
Neither Keras nor h5py exists on a B200 box, and the files are the simplest HDF5
there is: superblock v0, v1 object headers, old-style symbol-table groups and
contiguous little-endian float32 datasets.  This module walks exactly that subset
and returns ``{"layer/weight": ndarray}``; anything else raises ``H5FormatError``.
"""
from __future__ import annotations

import hashlib
import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5FormatError(ValueError):
    pass


class _H5:
    def __init__(self, buf: bytes):
        if buf[:8] != _SIG:
            raise H5FormatError("not an HDF5 file")
        if buf[8] != 0:
            raise H5FormatError("only superblock v0 is supported")
        if buf[13] != 8 or buf[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths are supported")
        self.b = buf
        base, _free, _eof, _drv = struct.unpack_from("<4Q", buf, 24)
        self.base = base
        # root symbol-table entry follows the four addresses
        _name_off, self.root_hdr, _cache, _res = struct.unpack_from("<QQII", buf, 56)

    # -- object headers -----------------------------------------------------
    def messages(self, addr: int):
        """Yield (type, payload bytes) of a v1 object header incl. continuations."""
        b = self.b
        addr += self.base
        ver, _r, nmsgs, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5FormatError("only v1 object headers are supported")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                payload = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((caddr + self.base, clen))
                else:
                    yield mtype, payload

    # -- groups ---------------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self.b
        heap_addr += self.base
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap")
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0] + self.base
        end = b.index(b"\x00", data_addr + off)
        return b[data_addr + off:end].decode("ascii")

    def _walk_btree(self, node_addr: int, heap_addr: int, out: Dict[str, int]):
        b = self.b
        a = node_addr + self.base
        sig = b[a:a + 4]
        if sig == b"TREE":
            _ntype, level, nent = struct.unpack_from("<BBH", b, a + 4)
            pos = a + 8 + 16  # skip two sibling addresses
            for _ in range(nent):
                pos += 8  # key
                child = struct.unpack_from("<Q", b, pos)[0]
                pos += 8
                self._walk_btree(child, heap_addr, out)
        elif sig == b"SNOD":
            nsym = struct.unpack_from("<H", b, a + 6)[0]
            pos = a + 8
            for _ in range(nsym):
                name_off, hdr = struct.unpack_from("<QQ", b, pos)
                out[self._heap_name(heap_addr, name_off)] = hdr
                pos += 40
        else:
            raise H5FormatError("unexpected group node signature %r" % sig)

    def children(self, hdr_addr: int) -> Dict[str, int]:
        for mtype, payload in self.messages(hdr_addr):
            if mtype == 0x11:  # symbol table
                btree, heap = struct.unpack_from("<QQ", payload, 0)
                out: Dict[str, int] = {}
                self._walk_btree(btree, heap, out)
                return out
        return {}

    # -- datasets -------------------------------------------------------------
    def dataset(self, hdr_addr: int):
        shape = None
        dtype = None
        addr = size = None
        for mtype, p in self.messages(hdr_addr):
            if mtype == 0x01:  # dataspace v1
                rank = p[1]
                shape = struct.unpack_from("<%dQ" % rank, p, 8)
            elif mtype == 0x03:  # datatype
                cls = p[0] & 0x0F
                dsize = struct.unpack_from("<I", p, 4)[0]
                dtype = (cls, dsize)
            elif mtype == 0x08:  # layout
                if p[0] != 3 or p[1] != 1:
                    raise H5FormatError("only contiguous v3 layouts are supported")
                addr, size = struct.unpack_from("<QQ", p, 2)
        if shape is None or dtype is None or addr is None:
            return None
        if dtype != (1, 4):
            return None  # not float32
        n = int(np.prod(shape)) if len(shape) else 1
        if size != 4 * n:
            raise H5FormatError("dataset size mismatch")
        off = addr + self.base
        arr = np.frombuffer(self.b, dtype="<f4", count=n, offset=off).reshape(shape)
        return arr.copy(), off


def read_keras_weights(path: str, with_offsets: bool = False):
    """Return ``{"conv2d_1/kernel:0": ndarray, ...}`` for every float32 dataset under
    ``/model_weights`` of a Keras ``.h5`` file (``(dict, sha256)``; offsets optional)."""
    with open(path, "rb") as f:
        buf = f.read()
    h5 = _H5(buf)
    root = h5.children(h5.root_hdr)
    if "model_weights" not in root:
        raise H5FormatError("no /model_weights group")
    out: Dict[str, np.ndarray] = {}
    offs: Dict[str, int] = {}

    def rec(hdr: int, prefix: str):
        kids = h5.children(hdr)
        if not kids:
            ds = h5.dataset(hdr)
            if ds is not None:
                out[prefix] = ds[0]
                offs[prefix] = ds[1]
            return
        for name, child in sorted(kids.items()):
            rec(child, name if not prefix else prefix + "/" + name)

    for lname, lhdr in sorted(h5.children(root["model_weights"]).items()):
        # layout is /model_weights/<layer>/<layer>/<weight>:0
        inner = h5.children(lhdr)
        for iname, ihdr in sorted(inner.items()):
            for wname, whdr in sorted(h5.children(ihdr).items()):
                ds = h5.dataset(whdr)
                if ds is not None:
                    out[lname + "/" + wname] = ds[0]
                    offs[lname + "/" + wname] = ds[1]
    sha = hashlib.sha256(buf).hexdigest()
    if with_offsets:
        return out, sha, offs
    return out, sha


def read_model_config(path: str) -> dict:
    """Extract the Keras ``model_config`` JSON (a variable-length string attribute kept in
    the file's global heap) by bracket-matching the first ``{"class_name"`` occurrence."""
    import json

    with open(path, "rb") as f:
        buf = f.read()
    start = buf.find(b'{"class_name"')
    if start < 0:
        raise H5FormatError("no model_config JSON found")
    depth = 0
    in_str = False
    esc = False
    for i in range(start, len(buf)):
        c = buf[i]
        if in_str:
            if esc:
                esc = False
            elif c == 0x5C:
                esc = True
            elif c == 0x22:
                in_str = False
            continue
        if c == 0x22:
            in_str = True
        elif c == 0x7B:
            depth += 1
        elif c == 0x7D:
            depth -= 1
            if depth == 0:
                return json.loads(buf[start:i + 1].decode("utf-8"))
    raise H5FormatError("unterminated model_config JSON")


def layer_summary(cfg: dict):
    """[(class_name, name, activation or None, extra)] for a Sequential/Model config."""
    layers = cfg["config"]["layers"] if isinstance(cfg["config"], dict) else cfg["config"]
    out = []
    for l in layers:
        c = l["config"]
        out.append((l["class_name"], c.get("name"), c.get("activation"),
                    {k: c[k] for k in ("filters", "units", "kernel_size", "padding", "pool_size",
                                       "strides", "batch_input_shape") if k in c}))
    return out
