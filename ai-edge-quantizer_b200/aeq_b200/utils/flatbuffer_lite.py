"""A small FlatBuffers reader and back-to-front builder (wire format only, no schema compiler).

The reference reaches `.tflite` bytes through the `flatbuffers` runtime and the generated
object API of `ai_edge_litert` (utils/tfl_flatbuffer_utils.py:118-263); neither wheel is part
of this package's dependencies, so the two pieces of the wire format that the quantizer needs
are implemented here: reading tables / vectors / strings in place (zero-copy NumPy views for
scalar vectors) and writing them back.

Wire format recap (all little endian):
  file    : uoffset32 to the root table, then an optional 4-byte file identifier
  table   : soffset32 to its vtable (vtable = table - soffset), then the inline fields
  vtable  : u16 vtable bytes, u16 table bytes, then one u16 per field id (0 = absent -> default)
  vector  : u32 count, then the elements; string = byte vector + a 0 terminator
  offsets : uoffset32 stored AT the field, target = field position + value
"""
from __future__ import annotations

import struct

import numpy as np

_SCALAR = {
    "bool": ("<?", 1), "byte": ("<b", 1), "ubyte": ("<B", 1), "short": ("<h", 2), "ushort": ("<H", 2),
    "int": ("<i", 4), "uint": ("<I", 4), "long": ("<q", 8), "ulong": ("<Q", 8),
    "float": ("<f", 4), "double": ("<d", 8),
}
_NP = {"bool": np.bool_, "byte": np.int8, "ubyte": np.uint8, "short": np.int16, "ushort": np.uint16,
       "int": np.int32, "uint": np.uint32, "long": np.int64, "ulong": np.uint64,
       "float": np.float32, "double": np.float64}


def scalar_size(kind: str) -> int:
  return _SCALAR[kind][1]


# ------------------------------------------------------------------------------ reading
class Table:
  """A table inside `buf` (any buffer-protocol object) at absolute position `pos`."""

  __slots__ = ("buf", "pos", "_vt", "_vt_len")

  def __init__(self, buf, pos: int):
    self.buf = buf
    self.pos = pos
    self._vt = pos - struct.unpack_from("<i", buf, pos)[0]
    self._vt_len = struct.unpack_from("<H", buf, self._vt)[0]

  @classmethod
  def root(cls, buf) -> "Table":
    return cls(buf, struct.unpack_from("<I", buf, 0)[0])

  def _field(self, field_id: int) -> int:
    """Absolute position of field `field_id`, or 0 when absent."""
    slot = 4 + 2 * field_id
    if slot >= self._vt_len:
      return 0
    off = struct.unpack_from("<H", self.buf, self._vt + slot)[0]
    return self.pos + off if off else 0

  def has(self, field_id: int) -> bool:
    return self._field(field_id) != 0

  def scalar(self, field_id: int, kind: str, default=0):
    p = self._field(field_id)
    return struct.unpack_from(_SCALAR[kind][0], self.buf, p)[0] if p else default

  def _indirect(self, field_id: int) -> int:
    p = self._field(field_id)
    return p + struct.unpack_from("<I", self.buf, p)[0] if p else 0

  def table(self, field_id: int):
    p = self._indirect(field_id)
    return Table(self.buf, p) if p else None

  def string(self, field_id: int):
    """bytes (like the flatbuffers object API) or None."""
    p = self._indirect(field_id)
    if not p:
      return None
    n = struct.unpack_from("<I", self.buf, p)[0]
    return bytes(self.buf[p + 4:p + 4 + n])

  def vector_len(self, field_id: int) -> int:
    p = self._indirect(field_id)
    return struct.unpack_from("<I", self.buf, p)[0] if p else 0

  def scalar_vector(self, field_id: int, kind: str):
    """Zero-copy NumPy view of a vector of scalars, or None when the field is absent."""
    p = self._indirect(field_id)
    if not p:
      return None
    n = struct.unpack_from("<I", self.buf, p)[0]
    return np.frombuffer(self.buf, dtype=np.dtype(_NP[kind]).newbyteorder("<"), count=n, offset=p + 4)

  def table_vector(self, field_id: int) -> list:
    p = self._indirect(field_id)
    if not p:
      return []
    n = struct.unpack_from("<I", self.buf, p)[0]
    out = []
    for i in range(n):
      e = p + 4 + 4 * i
      out.append(Table(self.buf, e + struct.unpack_from("<I", self.buf, e)[0]))
    return out

  def raw(self) -> tuple[bytes, bytes, int]:
    """(vtable bytes, table bytes, position mod 8): enough to re-emit a table that holds only
    scalars with every field's alignment intact."""
    table_len = struct.unpack_from("<H", self.buf, self._vt + 2)[0]
    return (bytes(self.buf[self._vt:self._vt + self._vt_len]),
            bytes(self.buf[self.pos:self.pos + table_len]), self.pos % 8)


def file_identifier(buf) -> bytes:
  return bytes(buf[4:8])


# ------------------------------------------------------------------------------ writing
class Builder:
  """Back-to-front builder: data is written from the end of a growing buffer towards its
  start, so every offset points forward (towards higher addresses) as the format requires."""

  def __init__(self, initial: int = 1 << 16):
    self._buf = bytearray(initial)
    self._head = initial     # first used byte
    self._minalign = 1
    self._vtables: dict[bytes, int] = {}   # vtable bytes -> offset-from-end, for sharing
    self._fields = None      # current table: {field id: offset-from-end of the inline value}
    self._table_start = 0

  # ---- space management
  def offset(self) -> int:
    """Bytes written so far = distance of the head from the END of the buffer."""
    return len(self._buf) - self._head

  def _grow(self, need: int) -> None:
    while self._head < need:
      old = len(self._buf)
      self._buf = bytearray(old) + self._buf
      self._head += old

  def prep(self, align: int, extra: int) -> None:
    """Pads so that, after `extra` more bytes are written, the head is `align`-aligned."""
    self._minalign = max(self._minalign, align)
    pad = (-(self.offset() + extra)) % align
    self._grow(pad + extra + align)
    self._head -= pad  # bytearray is zero-initialised: padding is zeros

  def _place(self, fmt: str, size: int, value) -> None:
    self._head -= size
    struct.pack_into(fmt, self._buf, self._head, value)

  def place_bytes(self, data) -> None:
    n = len(data)
    self._grow(n)
    self._head -= n
    self._buf[self._head:self._head + n] = data

  # ---- leaves
  def create_byte_vector(self, data, align: int = 4, elem_size: int = 1) -> int:
    """Vector whose payload is `data` (bytes-like); `len(data) / elem_size` elements.  `align`
    is the alignment of the FIRST ELEMENT (16 for TFLite buffers, 8 for [long])."""
    data = memoryview(data).cast("B") if not isinstance(data, (bytes, bytearray)) else data
    n = len(data)
    self.prep(max(align, 4), n)      # payload start aligned
    self.place_bytes(data)
    self.prep(4, 0)
    self._place("<I", 4, n // elem_size)
    return self.offset()

  def create_numpy_vector(self, arr: np.ndarray) -> int:
    a = np.ascontiguousarray(arr)
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)
    return self.create_byte_vector(a.tobytes() if a.size else b"", align=max(a.itemsize, 4),
                                   elem_size=a.itemsize)

  def create_string(self, s) -> int:
    data = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    self.prep(4, len(data) + 1)
    self._grow(1)
    self._head -= 1
    self._buf[self._head] = 0
    self.place_bytes(data)
    self._place("<I", 4, len(data))
    return self.offset()

  def create_offset_vector(self, offsets: list[int]) -> int:
    """Vector of tables / strings given their offsets-from-end."""
    self.prep(4, 4 * len(offsets))
    for off in reversed(offsets):
      self.prep(4, 0)
      self._place("<I", 4, self.offset() + 4 - off)  # relative to the element's own position
    self._place("<I", 4, len(offsets))
    return self.offset()

  # ---- tables
  def start_table(self) -> None:
    assert self._fields is None, "tables cannot nest: build children first"
    self._fields = {}
    self._table_start = self.offset()

  def add_scalar(self, field_id: int, kind: str, value, default=0) -> None:
    if value == default or value is None:
      return
    fmt, size = _SCALAR[kind]
    self.prep(size, 0)
    self._place(fmt, size, value)
    self._fields[field_id] = self.offset()

  def add_offset(self, field_id: int, off: int) -> None:
    """Reference to a previously built string / vector / table (0 / None: leave absent)."""
    if not off:
      return
    self.prep(4, 0)
    self._place("<I", 4, self.offset() + 4 - off)
    self._fields[field_id] = self.offset()

  def end_table(self) -> int:
    fields, self._fields = self._fields, None
    self.prep(4, 0)
    self._place("<i", 4, 0)  # soffset to the vtable, patched below
    table = self.offset()
    n = (max(fields) + 1) if fields else 0
    vt = bytearray(4 + 2 * n)
    struct.pack_into("<HH", vt, 0, len(vt), table - self._table_start)
    for fid, off in fields.items():
      struct.pack_into("<H", vt, 4 + 2 * fid, table - off)
    key = bytes(vt)
    shared = self._vtables.get(key)
    if shared is None:
      self.prep(2, len(vt))
      self.place_bytes(vt)
      shared = self.offset()
      self._vtables[key] = shared
    # soffset = table position - vtable position = (end - table) .. in offsets-from-end terms:
    pos = len(self._buf) - table
    struct.pack_into("<i", self._buf, pos, shared - table)
    return table

  def add_raw_table(self, vtable: bytes, table: bytes, pos_mod8: int = 0) -> int:
    """Re-emits a scalar-only table verbatim (its vtable is shared when possible) at a position
    with the same residue mod 8 as the original, so 8-byte scalars stay aligned."""
    assert self._fields is None
    self._minalign = max(self._minalign, 8)
    want = (-pos_mod8) % 8   # final position = total - offset, total is a multiple of 8
    pad = (want - (self.offset() + len(table))) % 8
    self._grow(pad + len(table) + 8)
    self._head -= pad
    self.place_bytes(table)
    tab = self.offset()
    shared = self._vtables.get(vtable)
    if shared is None:
      self.prep(2, len(vtable))
      self.place_bytes(vtable)
      shared = self.offset()
      self._vtables[vtable] = shared
    struct.pack_into("<i", self._buf, len(self._buf) - tab, shared - tab)
    return tab

  def finish(self, root: int, identifier: bytes = b"") -> bytes:
    assert len(identifier) in (0, 4)
    self.prep(self._minalign, 4 + len(identifier))
    if identifier:
      self.place_bytes(identifier)
    self._place("<I", 4, self.offset() + 4 - root)
    return bytes(self._buf[self._head:])
