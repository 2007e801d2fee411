"""Duck-typed stand-ins for the flatbuffer object API (tensor / buffer / op), enough
to drive materialisers and calibration functions without a .tflite file."""
import types

import numpy as np


def tensor(name, shape, buffer=0, ttype=0):
  return types.SimpleNamespace(name=name.encode(), shape=list(shape), buffer=buffer, type=ttype,
                               quantization=None)


def fc_graph(weight: np.ndarray, batch=4, bias=False):
  """INPUT[batch, in] x W[out, in] (+bias) -> OUTPUT[batch, out]."""
  from aeq_b200 import qtyping
  buffers = [types.SimpleNamespace(data=None),
             types.SimpleNamespace(data=weight.tobytes())]
  tensors = [tensor("input", (batch, weight.shape[1]), 0),
             tensor("weight", weight.shape, 1),
             tensor("output", (batch, weight.shape[0]), 0)]
  inputs = [0, 1, -1]
  if bias:
    b = np.zeros(weight.shape[0], np.float32)
    buffers.append(types.SimpleNamespace(data=b.tobytes()))
    tensors.append(tensor("bias", b.shape, 2))
    inputs = [0, 1, 3]
  op = types.SimpleNamespace(inputs=inputs, outputs=[2], builtinOptions=None)
  return op, qtyping.GraphInfo(subgraph_tensors=tensors, buffers=buffers)


def op_info(op, weight_cfg, op_name="FULLY_CONNECTED", **kw):
  from aeq_b200 import qtyping
  return qtyping.OpInfo(op, qtyping.TFLOperationName(op_name), 0,
                        qtyping.OpQuantizationConfig(weight_tensor_config=weight_cfg, **kw))
