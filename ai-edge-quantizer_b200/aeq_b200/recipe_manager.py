"""Scope-regex -> per-op quantisation configs.

Mirror of the public surface of ai_edge_quantizer/recipe_manager.py that a weight
quantisation run needs: `add_quantization_config` (:88-155), `add_dynamic_config` /
`add_weight_only_config` (:157-249), `load_quantization_recipe` / `get_quantization_recipe`
(:312-371) and `get_quantization_configs` (:373-409).  Later entries override earlier ones for
the same scope; a `*` entry is kept as one entry and resolved per op at lookup, where an op
whose config check fails is left float, like the reference.
"""
from __future__ import annotations

import collections
import dataclasses
import re
from typing import Optional

from . import algorithm_manager
from . import qtyping

_Op = qtyping.TFLOperationName
AlgorithmName = algorithm_manager.AlgorithmName


@dataclasses.dataclass
class OpQuantizationRecipe:
  regex: str
  operation: _Op
  algorithm_key: str
  op_config: qtyping.OpQuantizationConfig


class RecipeManager:

  def __init__(self):
    self._scope_configs: collections.OrderedDict[str, list[OpQuantizationRecipe]] = collections.OrderedDict()

  def add_quantization_config(self, regex: str, operation_name, op_config: Optional[qtyping.OpQuantizationConfig] = None,
                              algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    """A `*` entry stays ONE entry (it replaces what the scope held, recipe_manager.py:120-128) and
    is matched against each op at lookup time; a named op is validated here (:130-133)."""
    operation_name = _Op(operation_name)
    if op_config is None or op_config == {}:
      op_config = qtyping.OpQuantizationConfig()
    algorithm_key = getattr(algorithm_key, "value", algorithm_key)
    try:
      AlgorithmName(algorithm_key)
    except ValueError as e:
      raise ValueError(f"Unsupported algorithm key: {algorithm_key}.") from e
    entry = OpQuantizationRecipe(regex, operation_name, algorithm_key, op_config)
    if operation_name == _Op.ALL_SUPPORTED:
      self._scope_configs[regex] = [entry]
      return
    if algorithm_key != AlgorithmName.NO_QUANTIZE.value:
      algorithm_manager.check_op_quantization_config(algorithm_key, operation_name, op_config)
    configs = self._scope_configs.setdefault(regex, [])
    for i, c in enumerate(configs):
      if c.operation == operation_name:
        configs[i] = entry  # same op under the same scope: the new config wins, position kept
        return
    configs.append(entry)

  def add_dynamic_config(self, regex: str, operation_name, num_bits: int,
                         granularity=qtyping.QuantGranularity.CHANNELWISE,
                         algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    self.add_quantization_config(regex, operation_name, qtyping.OpQuantizationConfig(
        weight_tensor_config=qtyping.TensorQuantizationConfig(
            num_bits=num_bits, symmetric=True, granularity=granularity),
        compute_precision=qtyping.ComputePrecision.INTEGER, explicit_dequantize=False), algorithm_key)

  def add_weight_only_config(self, regex: str, operation_name, num_bits: int,
                             granularity=qtyping.QuantGranularity.CHANNELWISE,
                             algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    self.add_quantization_config(regex, operation_name, qtyping.OpQuantizationConfig(
        weight_tensor_config=qtyping.TensorQuantizationConfig(
            num_bits=num_bits, symmetric=True, granularity=granularity),
        compute_precision=qtyping.ComputePrecision.FLOAT, explicit_dequantize=True), algorithm_key)

  def get_quantization_configs(self, target_op_name, scope_name: str):
    """(algorithm key, op config) of the LAST matching entry whose config the algorithm accepts
    for this op; an entry the op's config check rejects (a blockwise `*` recipe meeting a
    CONV_2D, an op the algorithm does not cover) is skipped, so that op stays float
    (recipe_manager.py:157-202)."""
    result = (AlgorithmName.NO_QUANTIZE.value, qtyping.OpQuantizationConfig())
    for regex, configs in self._scope_configs.items():
      if not re.search(regex, scope_name):
        continue
      for c in configs:
        if c.operation not in (_Op.ALL_SUPPORTED, target_op_name):
          continue
        if c.algorithm_key != AlgorithmName.NO_QUANTIZE.value:
          try:
            algorithm_manager.check_op_quantization_config(c.algorithm_key, target_op_name, c.op_config)
          except ValueError:
            continue
        result = (c.algorithm_key, c.op_config)
    return result

  def get_quantization_recipe(self) -> list[dict]:
    """JSON-serialisable recipe, one dict per entry in insertion order (recipe_manager.py:204-228)."""
    out = []
    for regex, configs in self._scope_configs.items():
      for c in configs:
        entry = {"regex": regex, "operation": c.operation.value, "algorithm_key": c.algorithm_key}
        if c.algorithm_key != AlgorithmName.NO_QUANTIZE.value:
          entry["op_config"] = c.op_config.to_dict()
        out.append(entry)
    return out

  def load_quantization_recipe(self, quantization_recipe: list[dict]) -> None:
    self._scope_configs = collections.OrderedDict()
    for entry in quantization_recipe:
      cfg = entry.get("op_config")
      self.add_quantization_config(
          entry["regex"], entry["operation"],
          qtyping.OpQuantizationConfig.from_dict(cfg) if cfg else None, entry["algorithm_key"])

  def need_calibration(self) -> bool:
    for configs in self._scope_configs.values():
      for c in configs:
        if (c.algorithm_key != AlgorithmName.NO_QUANTIZE.value
            and c.op_config.activation_tensor_config is not None):
          return True
    return False
