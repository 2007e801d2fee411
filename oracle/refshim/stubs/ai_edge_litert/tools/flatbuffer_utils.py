"""Attribute-permissive stand-in for ai_edge_litert.tools.flatbuffer_utils.

Only the schema *names* are needed for the reference's numeric modules to
import; nothing here can parse a .tflite file. Test infrastructure only.
"""
import os
import types
from typing import Union


class TensorType:
  FLOAT32 = 0
  FLOAT16 = 1
  INT32 = 2
  UINT8 = 3
  INT64 = 4
  STRING = 5
  BOOL = 6
  INT16 = 7
  COMPLEX64 = 8
  INT8 = 9
  FLOAT64 = 10
  COMPLEX128 = 11
  UINT64 = 12
  RESOURCE = 13
  VARIANT = 14
  UINT32 = 15
  UINT16 = 16
  INT4 = 17
  BFLOAT16 = 18
  INT2 = 20


class _Bag:
  """Placeholder for a generated flatbuffer object-API class."""

  def __init__(self, **kw):
    self.__dict__.update(kw)

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    return None


class _Meta(type):
  """Class-level attribute access yields a distinct hashable token per name."""

  def __getattr__(cls, name):
    if name.startswith("__"):
      raise AttributeError(name)
    return f"{cls.__name__}.{name}"


def _make(name):
  return _Meta(name, (_Bag,), {})


_cache = {}
Path = Union[str, os.PathLike]
BufferType = Union[bytes, bytearray, memoryview]


class _Schema(types.SimpleNamespace):

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    return __getattr__(name)


schema_fb = _Schema()


def __getattr__(name):
  if name.startswith("__"):
    raise AttributeError(name)
  if name not in _cache:
    _cache[name] = _make(name)
  return _cache[name]
