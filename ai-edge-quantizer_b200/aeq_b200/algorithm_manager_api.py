"""Plug-in registry: (algorithm key, TFLite op) -> calibrate / materialize / QSV functions.

Behavioural mirror of the reference's `AlgorithmManagerApi`
(ai_edge_quantizer/algorithm_manager_api.py:170-440): same method names,
argument order, defaults and error messages, so code written against the
reference's registry runs against this one.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional

from . import qtyping
from .utils import qsv_utils

# Callable shapes (algorithm_manager_api.py:29-150), kept as aliases:
#   InitQSVFunc(op_info, graph_info, inputs_to_ignore=None, outputs_to_ignore=None, **kw) -> QSV map
#   CalibrationFunc(tfl_op, graph_info, tensor_content_map, inputs_to_ignore=None,
#                   outputs_to_ignore=None, **kw) -> dict[str, QSV]
#   MaterializeFunc(op_info, graph_info, tensor_name_to_qsv, tensor_quant_params_cache, **kw)
#                   -> list[TensorTransformationParams]
#   UpdateQSVFunc(qsv, new_qsv, **kw) -> QSV
InitQSVFunc = Callable[..., Any]
CalibrationFunc = Callable[..., Any]
MaterializeFunc = Callable[..., Any]
UpdateQSVFunc = Callable[..., Any]
CheckOpQuantConfigFunc = Callable[..., None]


@dataclasses.dataclass
class QuantizedOperationInfo:
  tfl_op_key: qtyping.TFLOperationName
  init_qsv_func: InitQSVFunc
  calibration_func: CalibrationFunc
  materialize_func: MaterializeFunc
  update_qsv_func: UpdateQSVFunc = qsv_utils.moving_average_update


@dataclasses.dataclass
class QuantizationAlgorithmInfo:
  quantization_algorithm: str
  quantized_ops: dict


class AlgorithmManagerApi:
  """Not thread-safe, like the reference (module-global instance, serial use)."""

  def __init__(self):
    self._algorithm_registry: dict[str, QuantizationAlgorithmInfo] = {}
    self._config_check_registry: dict[str, CheckOpQuantConfigFunc] = {}
    self._config_check_policy_registry: dict[str, Optional[dict]] = {}

  # ---- registration
  def register_op_quant_config_validation_func(self, algorithm_key, config_check_func):
    self._config_check_registry[algorithm_key] = config_check_func

  def register_config_check_policy(self, algorithm_key, config_check_policy):
    self._config_check_policy_registry[algorithm_key] = config_check_policy

  def register_quantized_op(self, algorithm_key, tfl_op_name, init_qsv_func,
                            calibration_func, materialize_func,
                            update_qsv_func=qsv_utils.moving_average_update):
    info = self._algorithm_registry.setdefault(
        algorithm_key, QuantizationAlgorithmInfo(algorithm_key, {}))
    info.quantized_ops[tfl_op_name] = QuantizedOperationInfo(
        tfl_op_name, init_qsv_func, calibration_func, materialize_func, update_qsv_func)

  # ---- queries
  def is_algorithm_registered(self, quantization_algorithm) -> bool:
    return quantization_algorithm in self._algorithm_registry

  def is_op_registered(self, quantization_algorithm, tfl_op_name) -> bool:
    return (self.is_algorithm_registered(quantization_algorithm) and tfl_op_name
            in self._algorithm_registry[quantization_algorithm].quantized_ops)

  def get_supported_ops(self, alg_key):
    if alg_key not in self._algorithm_registry:
      raise ValueError(f"Unregistered algorithm: {alg_key}")
    return list(self._algorithm_registry[alg_key].quantized_ops.keys())

  def _unsupported(self, algorithm_key, tfl_op_name) -> ValueError:
    return ValueError(
        f"Unsupported operation {tfl_op_name} for Algorithm: {algorithm_key}."
        f" Supported ops for algorithm {algorithm_key}:"
        f" {self.get_supported_ops(algorithm_key)}")

  def check_op_quantization_config(self, quantization_algorithm, tfl_op_name,
                                   op_quantization_config) -> None:
    if op_quantization_config.skip_checks:
      return
    if not self.is_op_registered(quantization_algorithm, tfl_op_name):
      raise ValueError(
          f"Unsupported operation {tfl_op_name} for Algorithm:"
          f" {quantization_algorithm}.")
    if quantization_algorithm not in self._config_check_registry:
      raise ValueError(
          f"Config checking function for  algorithm {quantization_algorithm} is"
          " not registered. Please use"
          " `register_op_quant_config_validation_func` to register the"
          " validation function.")
    self._config_check_registry[quantization_algorithm](
        tfl_op_name, op_quantization_config,
        self._config_check_policy_registry[quantization_algorithm])

  def get_quantization_func(self, algorithm_key, tfl_op_name, quantize_mode):
    if not self.is_op_registered(algorithm_key, tfl_op_name):
      raise self._unsupported(algorithm_key, tfl_op_name)
    op = self._algorithm_registry[algorithm_key].quantized_ops[tfl_op_name]
    func = {
        qtyping.QuantizeMode.CALIBRATE: op.calibration_func,
        qtyping.QuantizeMode.MATERIALIZE: op.materialize_func,
    }.get(quantize_mode)
    if func is None:
      raise ValueError(
          "Cannot retrieve appropriate quantization function for"
          f" {tfl_op_name} for algorithm {algorithm_key} under quantization"
          f" mode {quantize_mode}. Check if the op is registed in"
          " algorithm_manager.")
    return func

  def get_update_qsv_func(self, algorithm_key, tfl_op_name):
    func = self._algorithm_registry[algorithm_key].quantized_ops[tfl_op_name].update_qsv_func
    if not func:
      raise self._unsupported(algorithm_key, tfl_op_name)
    return func

  def get_init_qsv_func(self, algorithm_key, tfl_op_name):
    if not self.is_op_registered(algorithm_key, tfl_op_name):
      raise self._unsupported(algorithm_key, tfl_op_name)
    return self._algorithm_registry[algorithm_key].quantized_ops[tfl_op_name].init_qsv_func
