"""Synthetic TFLite models built with aeq_b200's own writer (no binary fixtures, no reference
assets): a stack of FULLY_CONNECTED layers with optional bias, optionally sharing one weight
buffer, plus an EMBEDDING_LOOKUP front."""
import struct

import numpy as np

from aeq_b200.utils import tfl_model as T


def _fc_options(fused_activation=0, keep_num_dims=False):
  """FullyConnectedOptions{fused_activation_function:byte(0), weights_format:byte(1),
  keep_num_dims:bool(2)} as the raw scalar-only table the reader would have produced."""
  # vtable: size, table size, 3 field offsets; table: soffset + 2 bytes + padding
  vt = struct.pack("<HHHHH", 10, 8, 4, 0, 5)
  tab = struct.pack("<ibbH", 0, fused_activation, int(keep_num_dims), 0)
  return T.RawTable(vt, tab, 0)


def fc_stack(weights, biases=None, share_first_weight_with_last=False, embedding=None, batch=2,
             conv_front=None):
  """weights: list of [out, in] fp32 arrays chained in -> out; returns ModelT.
  conv_front: optional [out_ch, kh, kw, in_ch] fp32 CONV_2D weight placed in front of the stack
  (shapes are not meant to run in an interpreter: only the quantizer walks this graph)."""
  m = T.ModelT(version=3, description=b"aeq_b200 synthetic")
  m.buffers.append(T.BufferT())  # buffer 0 is the empty sentinel
  g = T.SubGraphT(name=b"main")
  m.operatorCodes.append(T.OperatorCodeT(deprecatedBuiltinCode=T.BuiltinOperator.FULLY_CONNECTED,
                                         builtinCode=T.BuiltinOperator.FULLY_CONNECTED, version=1))

  def const(name, arr, ttype):
    m.buffers.append(T.BufferT(data=np.frombuffer(np.ascontiguousarray(arr).tobytes(), np.uint8)))
    g.tensors.append(T.TensorT(shape=np.array(arr.shape, np.int32), type=ttype, buffer=len(m.buffers) - 1,
                               name=name))
    return len(g.tensors) - 1

  def act(name, shape):
    g.tensors.append(T.TensorT(shape=np.array(shape, np.int32), type=T.TensorType.FLOAT32, buffer=0, name=name))
    return len(g.tensors) - 1

  cur = act(b"input", (batch, weights[0].shape[1]))
  g.inputs = np.array([cur], np.int32)
  if embedding is not None:
    m.operatorCodes.append(T.OperatorCodeT(deprecatedBuiltinCode=T.BuiltinOperator.EMBEDDING_LOOKUP,
                                           builtinCode=T.BuiltinOperator.EMBEDDING_LOOKUP, version=1))
    g.tensors[cur].type = T.TensorType.INT32
    g.tensors[cur].shape = np.array([batch], np.int32)
    table = const(b"embedding/table", embedding, T.TensorType.FLOAT32)
    out = act(b"embedding/out", (batch, embedding.shape[1]))
    g.operators.append(T.OperatorT(opcodeIndex=1, inputs=np.array([cur, table], np.int32),
                                   outputs=np.array([out], np.int32)))
    cur = out
  if conv_front is not None:
    m.operatorCodes.append(T.OperatorCodeT(deprecatedBuiltinCode=T.BuiltinOperator.CONV_2D,
                                           builtinCode=T.BuiltinOperator.CONV_2D, version=1))
    cw = const(b"conv/w", conv_front, T.TensorType.FLOAT32)
    out = act(b"conv/out", (batch, weights[0].shape[1]))
    g.operators.append(T.OperatorT(opcodeIndex=len(m.operatorCodes) - 1,
                                   inputs=np.array([cur, cw, -1], np.int32),
                                   outputs=np.array([out], np.int32)))
    cur = out
  first_w = None
  for i, w in enumerate(weights):
    if share_first_weight_with_last and i == len(weights) - 1 and first_w is not None:
      g.tensors.append(T.TensorT(shape=np.array(w.shape, np.int32), type=T.TensorType.FLOAT32,
                                 buffer=g.tensors[first_w].buffer, name=b"layer%d/w_shared" % i))
      wid = len(g.tensors) - 1
    else:
      wid = const(b"layer%d/w" % i, w, T.TensorType.FLOAT32)
    if first_w is None:
      first_w = wid
    bid = -1
    if biases is not None and biases[i] is not None:
      bid = const(b"layer%d/b" % i, biases[i], T.TensorType.FLOAT32)
    out = act(b"layer%d/out" % i, (batch, w.shape[0]))
    g.operators.append(T.OperatorT(opcodeIndex=0, inputs=np.array([cur, wid, bid], np.int32),
                                   outputs=np.array([out], np.int32), builtinOptionsType=8,
                                   builtinOptions=_fc_options()))
    cur = out
  g.outputs = np.array([cur], np.int32)
  m.subgraphs.append(g)
  m.signatureDefs.append(T.SignatureDefT(
      inputs=[T.TensorMapT(b"x", int(g.inputs[0]))], outputs=[T.TensorMapT(b"y", int(cur))],
      signatureKey=b"serving_default", subgraphIndex=0))
  return m
