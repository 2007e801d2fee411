"""Scope-regex -> per-op quantisation configs.

Mirror of the public surface of ai_edge_quantizer/recipe_manager.py that a weight
quantisation run needs: `add_quantization_config` (:88-155), `add_dynamic_config` /
`add_weight_only_config` (:157-249), `load_quantization_recipe` / `get_quantization_recipe`
(:312-371) and `get_quantization_configs` (:373-409).  Later entries override earlier ones for
the same scope; `*` covers every op the algorithm supports.
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
    operation_name = _Op(operation_name)
    if op_config is None or op_config == {}:
      op_config = qtyping.OpQuantizationConfig()
    algorithm_key = getattr(algorithm_key, "value", algorithm_key)
    if algorithm_key != AlgorithmName.NO_QUANTIZE.value:
      if not algorithm_manager.is_algorithm_registered(algorithm_key):
        raise ValueError(f"Unregistered algorithm: {algorithm_key}")
      targets = (algorithm_manager.get_supported_ops(algorithm_key)
                 if operation_name == _Op.ALL_SUPPORTED else [operation_name])
      if operation_name != _Op.ALL_SUPPORTED:
        algorithm_manager.check_op_quantization_config(algorithm_key, operation_name, op_config)
    else:
      targets = [operation_name]
    configs = self._scope_configs.setdefault(regex, [])
    if operation_name == _Op.ALL_SUPPORTED:
      configs.clear()  # a blanket entry replaces what the scope held (reference :120-128)
    for op in targets:
      configs[:] = [c for c in configs if c.operation != op]
      configs.append(OpQuantizationRecipe(regex, op, algorithm_key, op_config))

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
    """(algorithm key, op config) of the LAST scope whose regex matches, like the reference."""
    result = (AlgorithmName.NO_QUANTIZE.value, qtyping.OpQuantizationConfig())
    for regex, configs in self._scope_configs.items():
      if re.search(regex, scope_name):
        for c in configs:
          if c.operation == target_op_name:
            result = (c.algorithm_key, c.op_config)
    return result

  def get_quantization_recipe(self) -> list[dict]:
    """JSON-serialisable recipe; `*` entries are folded back when a scope covers every op of
    its algorithm with one config."""
    out = []
    for regex, configs in self._scope_configs.items():
      by_cfg = collections.OrderedDict()
      for c in configs:
        by_cfg.setdefault((c.algorithm_key, repr(c.op_config.to_dict())), []).append(c)
      for (alg, _), group in by_cfg.items():
        ops = {c.operation for c in group}
        blanket = (alg != AlgorithmName.NO_QUANTIZE.value
                   and ops == set(algorithm_manager.get_supported_ops(alg)) and len(ops) > 1)
        for c in ([group[0]] if blanket else group):
          entry = {"regex": regex, "operation": "*" if blanket else c.operation.value,
                   "algorithm_key": alg}
          if alg != AlgorithmName.NO_QUANTIZE.value:
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
