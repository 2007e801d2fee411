"""`Quantizer`: recipe in, quantised `.tflite` bytes out, for weight quantisation recipes.

Mirror of the public surface of ai_edge_quantizer/quantizer.py (`Quantizer` :131-520,
`QuantizationResult` :59-128) over this package's own model reader / writer
(utils/tfl_model.py): load a float model, describe what to quantise (recipe JSON, or
`add_dynamic_config` / `add_weight_only_config` / `update_quantization_recipe`), `quantize()`.

Flow of `quantize()` (reference: params_generator.py:69-185 -> model_modifier.py:60-140):
  1. walk every subgraph's operators, resolve (algorithm, op config) from the recipe by the
     op's scope, and call the registered materialiser — which runs the device kernels through
     the algorithm's `get_tensor_quant_params` and the shared (buffer, config) cache;
  2. apply the resulting transformations to the object tree: QUANTIZE_TENSOR for dynamic-range
     configs, ADD_DEQUANTIZE (quantise + DEQUANTIZE op) for weight-only configs;
  3. serialise.
Recipes that quantise activations need calibration through the LiteRT interpreter, which is a
caller of this path and not part of this package: `calibrate` raises.
"""
from __future__ import annotations

import dataclasses
import json
import os
import pathlib
from typing import Optional, Union

from . import algorithm_manager
from . import prefetch
from . import qtyping
from . import recipe_manager
from .algorithms.utils import common_utils
from .transformations import quantize_tensor as qt
from .utils import tfl_flatbuffer_utils as fu
from .utils import tfl_model

AlgorithmName = algorithm_manager.AlgorithmName
_QT = qtyping.QuantTransformation


@dataclasses.dataclass
class QuantizationResult:
  recipe: list
  quantized_model: Optional[bytes]

  def export_model(self, filepath, overwrite: bool = False) -> None:
    if self.quantized_model is None:
      raise RuntimeError("No quantized model to save. Make sure .quantize() is called.")
    if os.path.exists(filepath) and not overwrite:
      raise ValueError(
          f"The model {filepath} already exists in the folder. Please"
          " consider change the model name or specify overwrite=True to"
          " overwrite the model if needed.")
    with open(filepath, "wb") as f:
      f.write(self.quantized_model)

  def save(self, save_folder, model_name: str, overwrite: bool = False) -> None:
    os.makedirs(save_folder, exist_ok=True)
    self.export_model(str(pathlib.Path(save_folder) / f"{model_name}.tflite"), overwrite)
    with open(pathlib.Path(save_folder) / (model_name + "_recipe.json"), "w") as f:
      json.dump(self.recipe, f)


class Quantizer:
  """Quantises the constant weights of a TFLite model on the device."""

  def __init__(self, float_model: Union[str, pathlib.Path, bytes, bytearray, memoryview],
               quantization_recipe=None):
    import time
    t0 = time.perf_counter()
    if isinstance(float_model, (str, pathlib.Path)):
      self._float_model = fu.read_model(str(float_model))
      self._float_size = os.path.getsize(float_model)
    else:
      self._float_model = fu.read_model_from_bytes(float_model)
      self._float_size = len(float_model)
    self._recipe_manager = recipe_manager.RecipeManager()
    if quantization_recipe is not None:
      self.load_quantization_recipe(quantization_recipe)
    self._result = QuantizationResult([{}], None)
    self.prefetch_stats: dict = {}
    self.read_seconds = time.perf_counter() - t0
    self.timings: dict = {}

  # ---- recipe
  def load_quantization_recipe(self, recipe) -> None:
    if isinstance(recipe, (str, pathlib.Path)):
      with open(recipe) as f:
        recipe = json.load(f)
    self._recipe_manager.load_quantization_recipe(recipe)

  def get_quantization_recipe(self) -> list:
    return self._recipe_manager.get_quantization_recipe()

  def update_quantization_recipe(self, regex: str, operation_name, op_config=None,
                                 algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    self._recipe_manager.add_quantization_config(regex, operation_name, op_config, algorithm_key)

  def add_dynamic_config(self, regex: str, operation_name, num_bits: int,
                         granularity=qtyping.QuantGranularity.CHANNELWISE,
                         algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    self._recipe_manager.add_dynamic_config(regex, operation_name, num_bits, granularity, algorithm_key)

  def add_weight_only_config(self, regex: str, operation_name, num_bits: int,
                             granularity=qtyping.QuantGranularity.CHANNELWISE,
                             algorithm_key: str = AlgorithmName.MIN_MAX_UNIFORM_QUANT) -> None:
    self._recipe_manager.add_weight_only_config(regex, operation_name, num_bits, granularity, algorithm_key)

  def need_calibration(self) -> bool:
    return self._recipe_manager.need_calibration()

  def calibrate(self, *args, **kwargs):
    raise NotImplementedError(
        "calibration runs the float model in the LiteRT interpreter, which is outside this"
        " package; collect QSVs with the reference Calibrator (aeq_b200.plugin.install keeps its"
        " calibration functions on the device) and pass them to quantize(calibration_result=...)")

  # ---- quantisation
  def _generate_params(self, calibration_result) -> list:
    """[(subgraph index, TensorTransformationParams)] for every op the recipe selects."""
    model = self._float_model
    cache = common_utils.TensorQuantParamsCache()
    qsvs = calibration_result or {}
    out = []
    # Batched driver first: every min-max weight of the model goes through
    # ONE pipelined host-buffer call per (granularity, bits) group and lands in the cache, so
    # the per-op walk below (params_generator.py:110-183) finds its constants already done.
    items, mse_items = [], []
    for subgraph in model.subgraphs:
      graph_info = qtyping.GraphInfo(subgraph.tensors, model.buffers)
      for op_index, op in enumerate(subgraph.operators):
        code = tfl_model.builtin_code(model.operatorCodes[op.opcodeIndex])
        if code not in fu.TFL_OP_CODE_TO_NAME:
          continue
        op_name = fu.TFL_OP_CODE_TO_NAME[code]
        alg, cfg = self._recipe_manager.get_quantization_configs(
            op_name, fu.get_op_scope(op, subgraph.tensors))
        if alg == AlgorithmName.MIN_MAX_UNIFORM_QUANT.value and cfg.weight_tensor_config is not None:
          items.append((qtyping.OpInfo(op, op_name, op_index, cfg), graph_info))
        elif alg == AlgorithmName.MSE.value and cfg.weight_tensor_config is not None:
          mse_items.append((qtyping.OpInfo(op, op_name, op_index, cfg), graph_info))
    if items:
      self.prefetch_stats = prefetch.prefetch_weights(items, cache)
    if mse_items:
      st = prefetch.prefetch_weights(mse_items, cache, algorithm="MSE")
      self.prefetch_stats = {k: self.prefetch_stats.get(k, 0) + v for k, v in st.items()}
    for sg_index, subgraph in enumerate(model.subgraphs):
      graph_info = qtyping.GraphInfo(subgraph.tensors, model.buffers)
      for op_index, op in enumerate(subgraph.operators):
        code = tfl_model.builtin_code(model.operatorCodes[op.opcodeIndex])
        if code not in fu.TFL_OP_CODE_TO_NAME:
          continue
        op_name = fu.TFL_OP_CODE_TO_NAME[code]
        alg, cfg = self._recipe_manager.get_quantization_configs(
            op_name, fu.get_op_scope(op, subgraph.tensors))
        if alg == AlgorithmName.NO_QUANTIZE.value or cfg.weight_tensor_config is None:
          continue
        if cfg.activation_tensor_config is not None and not qsvs:
          raise RuntimeError(
              "Model quantization statistics values (QSVs) are required for this recipe:"
              " run calibration first (need_calibration() is True).")
        materialize = algorithm_manager.get_quantization_func(
            alg, op_name, qtyping.QuantizeMode.MATERIALIZE)
        op_info = qtyping.OpInfo(op, op_name, op_index, cfg)
        for params in materialize(op_info, graph_info, qsvs, cache):
          out.append((sg_index, params))
    return out

  def _apply(self, params_list) -> None:
    model = self._float_model
    buffer_origin: dict = {}
    for sg_index, subgraph in enumerate(model.subgraphs):
      names = {fu.get_tensor_name(t): i for i, t in enumerate(subgraph.tensors)}
      # ADD_DEQUANTIZE inserts ops, which shifts operator indices: consumers are captured as
      # operator objects first and located again at insertion time.
      pending = []
      for sg, p in params_list:
        if sg != sg_index:
          continue
        tid = names[p.tensor_name]
        for o2t in p.consumers or []:
          if _QT.QUANTIZE_TENSOR in o2t.transformations:
            qt.quantize_tensor(model, subgraph, tid, o2t.parameters, buffer_origin)
          elif _QT.ADD_DEQUANTIZE in o2t.transformations:
            pending.append((subgraph.operators[o2t.subgraph_op_id], tid, o2t.parameters))
          elif any(t not in (_QT.NO_QUANTIZE,) for t in o2t.transformations):
            raise NotImplementedError(
                f"transformation {o2t.transformations} is a graph edit of the reference's"
                " transformation_performer and is not applied by aeq_b200.Quantizer")
      by_tensor: dict = {}
      for op, tid, prm in pending:
        by_tensor.setdefault(tid, (prm, []))[1].append(op)
      for tid, (prm, consumers) in by_tensor.items():
        qt.insert_dequant(model, subgraph, tid, prm, consumers, buffer_origin)

  def quantize(self, calibration_result=None, serialize_to_path=None,
               external_buffers: Optional[bool] = None) -> QuantizationResult:
    """`external_buffers`: None = automatic (payloads of 2 GB and more leave the flatbuffer)."""
    if not self.get_quantization_recipe():
      raise RuntimeError("Can not quantize without a quantization recipe.")
    import time
    t0 = time.perf_counter()
    params = self._generate_params(calibration_result)
    t1 = time.perf_counter()
    self._apply(params)
    t2 = time.perf_counter()
    data = tfl_model.write_model_to_bytes(self._float_model, external_buffers)
    self.timings = {"quantization_parameters": t1 - t0, "transformations": t2 - t1,
                    "serialisation": time.perf_counter() - t2}
    if serialize_to_path is not None:
      with open(serialize_to_path, "wb") as f:
        f.write(data)
    self._result = QuantizationResult(self.get_quantization_recipe(), data)
    return self._result
