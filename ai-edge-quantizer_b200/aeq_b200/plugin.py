"""Drop-in installer: rebinds the REFERENCE's registry to the device kernels.

`install(am)` takes the reference's `ai_edge_quantizer.algorithm_manager`
module and re-registers every op of every uniform algorithm with the
reference's own materialiser (all its graph bookkeeping, constraints, bias and
cache logic stay untouched) partially applied to OUR
`get_tensor_quant_params` — the exact seam the reference uses itself
(algorithm_manager.py:160-163, 316-319, 415-418, 446-449; signature
qtyping.py:702-710).  The same registrations also swap the CALIBRATION functions
(`CalibrationFunc`, algorithm_manager_api.py:97-122: `min_max_calibrate`, `gptq.calibrate`,
`oscar.calibrate`) for the device ones, and `install` rebinds the pack step of QUANTIZE_TENSOR /
ADD_DEQUANTIZE (`transformation_utils.pack_data`, looked up at call time by
transformations/quantize_tensor.py:195-200, which `TransformationPerformer` registers at
transformation_performer.py:67-93).  Afterwards `Quantizer.quantize()` / `.calibrate()` reach the
sm_100a kernels for the weights, the activation statistics and the bit packing.

    import ai_edge_quantizer.algorithm_manager as am
    import aeq_b200.plugin
    aeq_b200.plugin.install(am)
"""
from __future__ import annotations

import functools

from . import qtyping as _qt
from .algorithms.uniform_quantize import dequantized_weight_recovery
from .algorithms.uniform_quantize import gptq
from .algorithms.uniform_quantize import hadamard_rotation
from .algorithms.uniform_quantize import mse
from .algorithms.uniform_quantize import naive_min_max_quantize
from .algorithms.uniform_quantize import octav
from .algorithms.uniform_quantize import oscar

# algorithm key -> (attribute of the reference module holding {op: materialize_fn},
#                   our get_tensor_quant_params, reference module attr that owns init/calibrate)
_BINDINGS = {
    "min_max_uniform_quantize": ("MIN_MAX_OP_NAME_MATERIALIZE_FUNC_DICT",
                                 naive_min_max_quantize.get_tensor_quant_params,
                                 "naive_min_max_quantize"),
    "OCTAV": ("_OCTAV_OP_NAME_MATERIALIZE_FUNC_DICT", octav.get_tensor_quant_params,
              "naive_min_max_quantize"),
    "MSE": ("_MSE_OP_NAME_MATERIALIZE_FUNC_DICT", mse.get_tensor_quant_params,
            "naive_min_max_quantize"),
    "GPTQ": ("_GPTQ_OP_NAME_MATERIALIZE_FUNC_DICT", gptq.get_tensor_quant_params,
             "naive_min_max_quantize"),
    "dequantized_weight_recovery": ("DEQUANTIZED_WEIGHT_RECOVERY_OP_NAME_MATERIALIZE_FUNC_DICT",
                                    dequantized_weight_recovery.get_tensor_quant_params,
                                    "dequantized_weight_recovery"),
}

# Algorithms whose reference materialisers call their module's own
# `get_tensor_quant_params` global at call time instead of receiving it through
# functools.partial (hadamard_rotation.py:253, :423): the seam is that attribute.
_MODULE_SEAMS = {
    "HADAMARD_ROTATION": ("hadamard_rotation", hadamard_rotation.get_tensor_quant_params),
    "DECOMPOSED_HADAMARD_ROTATION": ("hadamard_rotation", hadamard_rotation.get_tensor_quant_params),
    # oscar._get_or_compute_weight_quant_params calls the module's get_tensor_quant_params (:553)
    "OSCAR": ("oscar", oscar.get_tensor_quant_params),
}


def register_binding(algorithm_key: str, dict_attr: str, get_tensor_quant_params, ref_module: str):
  _BINDINGS[algorithm_key] = (dict_attr, get_tensor_quant_params, ref_module)


def _adapt(fn, ref_qtyping):
  """Wraps our function so it accepts / returns the REFERENCE's dataclasses."""

  @functools.wraps(fn)
  def get_tensor_quant_params(op_info, tensor_quant_config, tensor_content=None, tensor_qsv=None):
    cfg = _qt.TensorQuantizationConfig(
        num_bits=tensor_quant_config.num_bits, symmetric=tensor_quant_config.symmetric,
        granularity=_qt.QuantGranularity(tensor_quant_config.granularity.value),
        dtype=_qt.TensorDataType(tensor_quant_config.dtype.value),
        algorithm_params=dict(tensor_quant_config.algorithm_params))
    wcfg = op_info.op_quant_config.weight_tensor_config
    ours_op = _qt.OpInfo(
        op=op_info.op, op_name=_qt.TFLOperationName(op_info.op_name.value),
        subgraph_op_index=op_info.subgraph_op_index,
        op_quant_config=_qt.OpQuantizationConfig(
            weight_tensor_config=None if wcfg is None else _qt.TensorQuantizationConfig(
                num_bits=wcfg.num_bits, symmetric=wcfg.symmetric,
                granularity=_qt.QuantGranularity(wcfg.granularity.value),
                dtype=_qt.TensorDataType(wcfg.dtype.value),
                algorithm_params=dict(wcfg.algorithm_params)),
            skip_checks=True))
    r = fn(ours_op, cfg, tensor_content, tensor_qsv)
    hadamard = None
    if r.hadamard is not None:
      hadamard = ref_qtyping.UniformQuantParams.HadamardRotationParams(
          random_binary_vector=r.hadamard.random_binary_vector,
          hadamard_size=r.hadamard.hadamard_size)
    return ref_qtyping.UniformQuantParams(
        num_bits=r.num_bits, quantized_dimension=r.quantized_dimension, scale=r.scale,
        zero_point=r.zero_point, symmetric=r.symmetric, quantized_data=r.quantized_data,
        block_size=r.block_size, hadamard=hadamard,
        custom_algorithm_param=r.custom_algorithm_param)

  return get_tensor_quant_params


_saved: list = []  # undo log of install(): (callable, args) pairs, replayed in reverse


def uninstall() -> None:
  """Restores every registration / module attribute `install` replaced."""
  while _saved:
    fn, args = _saved.pop()
    fn(*args)


def _device_calibration_func(am, ref_func):
  """Our device calibration function for one of the reference's, or None when it has no twin."""
  from .algorithms.uniform_quantize import naive_min_max_quantize as ours_nmm
  twins = []
  for mod_name, attr, ours in (("naive_min_max_quantize", "min_max_calibrate", ours_nmm.min_max_calibrate),
                               ("gptq", "calibrate", gptq.calibrate),
                               ("oscar", "calibrate", oscar.calibrate),
                               ("dequantized_weight_recovery", "calibrate",
                                dequantized_weight_recovery.calibrate)):
    mod = getattr(am, mod_name, None)
    if mod is not None and getattr(mod, attr, None) is not None:
      twins.append((getattr(mod, attr), ours))
  for theirs, ours in twins:
    if ref_func is theirs:
      return ours
  return None


def install_pack(reference_transformation_utils) -> None:
  """Rebinds `transformation_utils.pack_data` (transformations/transformation_utils.py:293-353)
  to the device pack: INT4 / INT2 bytes come from `aeqb_pack_bits`, or straight from the fused
  requantisation pass when the integers were produced by this package."""
  tu = reference_transformation_utils
  if getattr(tu.pack_data, "_aeqb200", False):
    return
  from .transformations import quantize_tensor as ours_qt

  def pack_data(bitwidth: int, data):
    return ours_qt.pack_data(bitwidth, data)

  pack_data._aeqb200 = True
  _saved.append((setattr, (tu, "pack_data", tu.pack_data)))
  tu.pack_data = pack_data


def install(reference_algorithm_manager, algorithms=None, calibration: bool = True,
            pack: bool = True) -> list[str]:
  """Re-registers the reference's ops with device-backed arithmetic; returns the keys bound.

  calibration: also swap each op's calibration function for its device twin (the reference's
               `Calibrator` then reduces activations on the GPU, calibrator.py:545-582).
  pack:        also rebind `transformation_utils.pack_data` of the reference package.
  """
  am = reference_algorithm_manager
  ref_qtyping = am.qtyping
  bound = []

  def calib_for(key, op_name):
    theirs = am.get_quantization_func(key, op_name, ref_qtyping.QuantizeMode.CALIBRATE)
    ours = _device_calibration_func(am, theirs) if calibration else None
    return theirs if ours is None else ours

  for key, (dict_attr, fn, ref_module) in _BINDINGS.items():
    if algorithms is not None and key not in algorithms:
      continue
    op_dict = getattr(am, dict_attr, None) or getattr(am, "_" + dict_attr.lstrip("_"), None)
    if op_dict is None:
      continue
    mod = getattr(am, ref_module)
    adapted = _adapt(fn, ref_qtyping)
    for op_name, materialize_func in op_dict.items():
      inner = materialize_func.func if isinstance(materialize_func, functools.partial) else materialize_func
      _saved.append((am.register_quantized_op, (
          key, op_name, am.get_init_qsv_func(key, op_name),
          am.get_quantization_func(key, op_name, ref_qtyping.QuantizeMode.CALIBRATE),
          am.get_quantization_func(key, op_name, ref_qtyping.QuantizeMode.MATERIALIZE),
          am.get_update_qsv_func(key, op_name))))
      am.register_quantized_op(
          key, op_name, mod.init_qsvs,
          calibration_func=calib_for(key, op_name),
          materialize_func=functools.partial(inner, adapted),
          update_qsv_func=am.get_update_qsv_func(key, op_name))
    bound.append(key)
  for key, (ref_module, fn) in _MODULE_SEAMS.items():
    if algorithms is not None and key not in algorithms:
      continue
    mod = getattr(am, ref_module, None)
    if mod is None:
      continue
    if not getattr(mod.get_tensor_quant_params, "_aeqb200", False):
      adapted = _adapt(fn, ref_qtyping)
      adapted._aeqb200 = True
      _saved.append((setattr, (mod, "get_tensor_quant_params", mod.get_tensor_quant_params)))
      mod.get_tensor_quant_params = adapted
    if calibration and am.is_algorithm_registered(key):
      for op_name in am.get_supported_ops(key):
        theirs = am.get_quantization_func(key, op_name, ref_qtyping.QuantizeMode.CALIBRATE)
        ours = _device_calibration_func(am, theirs)
        if ours is None:
          continue
        args = (key, op_name, am.get_init_qsv_func(key, op_name), theirs,
                am.get_quantization_func(key, op_name, ref_qtyping.QuantizeMode.MATERIALIZE),
                am.get_update_qsv_func(key, op_name))
        _saved.append((am.register_quantized_op, args))
        am.register_quantized_op(args[0], args[1], args[2], calibration_func=ours,
                                 materialize_func=args[4], update_qsv_func=args[5])
    bound.append(key)
  if pack:
    import importlib
    pkg = am.__name__.rsplit(".", 1)[0]
    install_pack(importlib.import_module(pkg + ".transformations.transformation_utils"))
  return bound


def install_histogram(reference_algorithm_manager, ops=None) -> str:
  """Registers `histogram_min_max_uniform_quantize` in the REFERENCE's registry: for every op the
  reference's min-max algorithm covers, its own materialiser partially applied to our
  histogram-aware `get_tensor_quant_params`, `histogram_calibrate` as the calibration function
  and `histogram_update` as the QSV merge (register_quantized_op, algorithm_manager_api.py:191-227).
  A recipe then names the key as its algorithm; returns the key."""
  from .algorithms.uniform_quantize import histogram_calibration as hc
  am = reference_algorithm_manager
  adapted = _adapt(hc.get_tensor_quant_params, am.qtyping)
  for op_name, materialize_func in am.MIN_MAX_OP_NAME_MATERIALIZE_FUNC_DICT.items():
    if ops is not None and op_name not in ops:
      continue
    inner = materialize_func.func if isinstance(materialize_func, functools.partial) else materialize_func
    am.register_quantized_op(
        hc.ALGORITHM_KEY, op_name, am.naive_min_max_quantize.init_qsvs,
        calibration_func=hc.histogram_calibrate,
        materialize_func=functools.partial(inner, adapted),
        update_qsv_func=hc.histogram_update)
  am.register_op_quant_config_validation_func(
      hc.ALGORITHM_KEY, am.common_quantize.check_op_quantization_config)
  policy = getattr(am, "default_policy", None)
  if policy is not None and hasattr(am, "register_config_check_policy_func"):
    am.register_config_check_policy_func(hc.ALGORITHM_KEY, policy.DEFAULT_CONFIG_CHECK_POLICY)
  return hc.ALGORITHM_KEY


def prefetch(params_generator, model_recipe_manager) -> dict:
  """Quantises all min-max weights of the reference ParamsGenerator's model in a few batched
  launches and fills its `(buffer, config)` cache; call it right before
  `params_generator.generate_quantization_parameters(model_recipe_manager, ...)`.

  Mirrors the reference's own walk (params_generator.py:110-158): same op keys, scopes, recipe
  lookups and composite / NO_QUANTIZE skipping, using the reference's modules as found in the
  generator's module namespace.  The per-op materialisers then hit the cache
  (common_utils.py:260-264).
  """
  import sys
  from . import prefetch as _prefetch
  pg = params_generator
  mod = sys.modules[type(pg).__module__]
  fu, am, rq = mod.tfl_flatbuffer_utils, mod.algorithm_manager, mod.qtyping
  policy = getattr(mod, "policy", None)
  model = pg.float_model
  op_codes = model.operatorCodes
  min_max = am.AlgorithmName.MIN_MAX_UNIFORM_QUANT
  items, skip_subgraphs = [], set()
  for sg_ind, subgraph in enumerate(model.subgraphs):
    graph_info = rq.GraphInfo(subgraph.tensors, model.buffers)
    for op_id, op in enumerate(subgraph.operators):
      code = op_codes[op.opcodeIndex].builtinCode
      if code not in fu.TFL_OP_CODE_TO_NAME:
        continue
      op_key = fu.TFL_OP_CODE_TO_NAME[code]
      algorithm_name, op_config = model_recipe_manager.get_quantization_configs(
          op_key, fu.get_op_scope(op, subgraph.tensors))
      if sg_ind in skip_subgraphs or (policy is not None and policy.is_non_quantizable_composite_op(op)):
        algorithm_name = am.AlgorithmName.NO_QUANTIZE
      if algorithm_name == am.AlgorithmName.NO_QUANTIZE:
        skip_subgraphs.update(fu.get_op_side_effect_subgraphs(op))
        continue
      if algorithm_name != min_max or op_config.weight_tensor_config is None:
        continue
      items.append((rq.OpInfo(op, op_key, op_id, op_config), graph_info))
  return _prefetch.prefetch_weights(
      items, pg._tensor_quant_params_cache,  # pylint: disable=protected-access
      make_params=lambda **kw: rq.UniformQuantParams(**kw), get_tensor_data=fu.get_tensor_data)
