"""Batched weight driver: every min-max weight of a synthetic model in a few
launches, cache pre-populated, per-op materialise calls become hits with identical results."""
import types

import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu


def _model(shapes, cfg_of):
  """One subgraph of FC ops; op k reads weight k (ops 1 and 2 SHARE a buffer)."""
  from aeq_b200 import qtyping
  buffers = [types.SimpleNamespace(data=None)]
  tensors, ops, weights = [], [], []
  for k, (r, c) in enumerate(shapes):
    if k == 2:
      buf = len(buffers) - 1          # same buffer as op 1
      w = weights[-1]
    else:
      w = O.synthetic_weight(r, c, 300 + k)
      buffers.append(types.SimpleNamespace(data=w.tobytes()))
      buf = len(buffers) - 1
    weights.append(w)
    base = len(tensors)
    tensors += [sg.tensor(f"in{k}", (4, w.shape[1]), 0), sg.tensor(f"w{k}", w.shape, buf),
                sg.tensor(f"out{k}", (4, w.shape[0]), 0)]
    ops.append(types.SimpleNamespace(inputs=[base, base + 1, -1], outputs=[base + 2], builtinOptions=None))
  graph = qtyping.GraphInfo(subgraph_tensors=tensors, buffers=buffers)
  infos = [qtyping.OpInfo(op, qtyping.TFLOperationName.FULLY_CONNECTED, k,
                          qtyping.OpQuantizationConfig(
                              weight_tensor_config=cfg_of(k),
                              compute_precision=qtyping.ComputePrecision.INTEGER))
           for k, op in enumerate(ops)]
  return graph, infos, weights


def test_prefetch_fills_cache_and_matches_per_op_path(cuda):
  from aeq_b200 import _lib, prefetch, qtyping
  from aeq_b200 import algorithm_manager as am
  from aeq_b200.algorithms.utils import common_utils
  G = qtyping.QuantGranularity
  shapes = [(64, 256), (130, 1024), (130, 1024), (32, 4096), (16, 2048), (8, 11008), (48, 64)]
  int8 = qtyping.TensorQuantizationConfig(8, True, G.CHANNELWISE)
  int4b = qtyping.TensorQuantizationConfig(4, True, G.BLOCKWISE_32)
  per_tensor = qtyping.TensorQuantizationConfig(8, True, G.TENSORWISE)
  cfg_of = lambda k: int4b if k in (3, 4) else (per_tensor if k == 6 else int8)
  graph, infos, weights = _model(shapes, cfg_of)

  cache = common_utils.TensorQuantParamsCache()
  lib = _lib.load()
  l0 = lib.aeqb_launch_count()
  stats = prefetch.prefetch_weights([(i, graph) for i in infos], cache)
  launches = lib.aeqb_launch_count() - l0
  assert stats == {"quantized": 5, "already_cached": 0, "left_to_per_op_path": 1, "batched_calls": 2}
  assert launches <= 6, launches  # two groups, a few stream-class / generic launches each
  assert len(cache) == 5

  fn = am.get_quantization_func("min_max_uniform_quantize", qtyping.TFLOperationName.FULLY_CONNECTED,
                                qtyping.QuantizeMode.MATERIALIZE)
  for k, info in enumerate(infos):
    hit = cache.lookup(graph.subgraph_tensors[info.op.inputs[1]].buffer, cfg_of(k))
    out = fn(op_info=info, graph_info=graph, tensor_name_to_qsv={}, tensor_quant_params_cache=cache)
    params = [t for t in out if t.tensor_name == f"w{k}"][0].consumers[0].parameters
    if k != 6:
      assert params is hit  # served from the prefetched cache
    block = 32 if k in (3, 4) else 0
    ref = O.minmax_requant(weights[k], cfg_of(k).num_bits, True, block=block, per_channel=(k != 6))
    np.testing.assert_array_equal(params.quantized_data, ref["q"])
    np.testing.assert_array_equal(params.scale, ref["scale"])
    assert params.quantized_dimension == (1 if block else (None if k == 6 else 0))
  # a second prefetch has nothing left to do
  again = prefetch.prefetch_weights([(i, graph) for i in infos], cache)
  assert again["quantized"] == 0 and again["already_cached"] == 6  # incl. the per-tensor one
