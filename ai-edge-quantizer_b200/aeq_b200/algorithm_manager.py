"""Module-level registry instance + the algorithms this package implements.

Mirror of ai_edge_quantizer/algorithm_manager.py:43-75 (exposed functions and
the `AlgorithmName` keys) with registrations for the weight-carrying ops that
reach the hot path.  Registration keeps the reference's binding pattern
(algorithm_manager.py:160-163): `functools.partial(materialize, <alg>.get_tensor_quant_params)`.
"""
from __future__ import annotations

import enum
import functools

from . import algorithm_manager_api
from . import qtyping
from .algorithms.nonlinear_quantize import float_casting
from .algorithms.uniform_quantize import dequantized_weight_recovery
from .algorithms.uniform_quantize import gptq
from .algorithms.uniform_quantize import hadamard_rotation
from .algorithms.uniform_quantize import histogram_calibration
from .algorithms.uniform_quantize import mse
from .algorithms.uniform_quantize import naive_min_max_quantize
from .algorithms.uniform_quantize import octav
from .algorithms.uniform_quantize import oscar
from .algorithms.utils import common_utils
from .utils import qsv_utils

_Op = qtyping.TFLOperationName

_alg_manager_instance = algorithm_manager_api.AlgorithmManagerApi()

get_quantization_func = _alg_manager_instance.get_quantization_func
get_supported_ops = _alg_manager_instance.get_supported_ops
get_update_qsv_func = _alg_manager_instance.get_update_qsv_func
get_init_qsv_func = _alg_manager_instance.get_init_qsv_func
register_op_quant_config_validation_func = (
    _alg_manager_instance.register_op_quant_config_validation_func)
register_config_check_policy_func = _alg_manager_instance.register_config_check_policy
register_quantized_op = _alg_manager_instance.register_quantized_op
is_op_registered = _alg_manager_instance.is_op_registered
is_algorithm_registered = _alg_manager_instance.is_algorithm_registered
check_op_quantization_config = _alg_manager_instance.check_op_quantization_config


class AlgorithmName(str, enum.Enum):
  NO_QUANTIZE = "no_quantize"
  MIN_MAX_UNIFORM_QUANT = "min_max_uniform_quantize"
  FLOAT_CASTING = "float_casting"
  DEQUANTIZED_WEIGHT_RECOVERY = "dequantized_weight_recovery"
  OCTAV = "OCTAV"
  HADAMARD_ROTATION = "HADAMARD_ROTATION"
  DECOMPOSED_HADAMARD_ROTATION = "DECOMPOSED_HADAMARD_ROTATION"
  MSE = "MSE"
  GPTQ = "GPTQ"
  OSCAR = "OSCAR"


def _init_qsvs(op_info, graph_info, inputs_to_ignore=None, outputs_to_ignore=None, **kw):
  """Runtime tensors start without statistics; the first batch is stored verbatim."""
  del op_info, graph_info, inputs_to_ignore, outputs_to_ignore, kw
  return {}


def _check_config(op_name, op_quant_config, config_check_policy=None):
  """Blockwise rules of common_utils.check_subchannel_config (:80-101)."""
  del config_check_policy
  w = op_quant_config.weight_tensor_config
  if w is None or not common_utils.is_blockwise(w.granularity):
    return
  if op_name not in (_Op.FULLY_CONNECTED, _Op.EMBEDDING_LOOKUP):
    raise ValueError(f"Unsupported op for blockwise quantization: {op_name}")
  if op_quant_config.activation_tensor_config is not None:
    raise ValueError("Blockwise quantization does not support activation tensor quantization.")
  if not w.symmetric:
    raise ValueError("Blockwise quantization does not support for asymmetric weight.")


# op -> positional inputs that never carry quantisable data
WEIGHT_OPS = {
    _Op.FULLY_CONNECTED: (),
    _Op.CONV_2D: (),
    _Op.CONV_2D_TRANSPOSE: (0,),   # output-shape tensor
    _Op.EMBEDDING_LOOKUP: (0,),    # lookup indices
}


def register_weight_algorithm(algorithm_key, get_tensor_quant_params, calibration_func,
                              update_qsv_func=qsv_utils.moving_average_update):
  """Binds one algorithm's `get_tensor_quant_params` to every weight-carrying op."""
  for op_name, ignore in WEIGHT_OPS.items():
    register_quantized_op(
        algorithm_key, op_name, _init_qsvs, calibration_func=calibration_func,
        materialize_func=functools.partial(
            common_utils.materialize_weight_op, get_tensor_quant_params,
            inputs_to_ignore=ignore),
        update_qsv_func=update_qsv_func)
  register_op_quant_config_validation_func(algorithm_key, _check_config)
  register_config_check_policy_func(algorithm_key, None)


register_weight_algorithm(AlgorithmName.MIN_MAX_UNIFORM_QUANT,
                          naive_min_max_quantize.get_tensor_quant_params,
                          naive_min_max_quantize.min_max_calibrate)
# OCTAV / MSE calibrate activations exactly like min-max (algorithm_manager.py:316-336, 446-466).
register_weight_algorithm(AlgorithmName.OCTAV, octav.get_tensor_quant_params,
                          naive_min_max_quantize.min_max_calibrate)
register_weight_algorithm(AlgorithmName.MSE, mse.get_tensor_quant_params,
                          naive_min_max_quantize.min_max_calibrate)
# Histogram calibration: the reference ships utils/histogram_utils.py but
# binds it to no algorithm, so the key is ours (a plain string, not an `AlgorithmName` member:
# that enum stays equal to the reference's).  Weights quantise exactly like min-max.
register_weight_algorithm(histogram_calibration.ALGORITHM_KEY,
                          histogram_calibration.get_tensor_quant_params,
                          histogram_calibration.histogram_calibrate,
                          update_qsv_func=histogram_calibration.histogram_update)
# Weight-only rotation; the reference's own FC / EMBEDDING materialisers add the
# activation-side INSERT_HADAMARD_ROTATION (hadamard_rotation.py:206-500) and are
# reached through plugin.install.
register_weight_algorithm(AlgorithmName.HADAMARD_ROTATION,
                          hadamard_rotation.get_tensor_quant_params,
                          naive_min_max_quantize.min_max_calibrate)
# Same weight arithmetic; the variants differ only in the activation-side transformation the
# reference's materialisers add (INSERT_DECOMPOSED_HADAMARD_ROTATION, hadamard_rotation.py:283).
register_weight_algorithm(AlgorithmName.DECOMPOSED_HADAMARD_ROTATION,
                          hadamard_rotation.get_tensor_quant_params,
                          naive_min_max_quantize.min_max_calibrate)
# GPTQ: FULLY_CONNECTED only in the reference (algorithm_manager.py:434-455); the activation
# QSV carries the Hessian and merges by sample-weighted mean.
register_quantized_op(
    AlgorithmName.GPTQ, _Op.FULLY_CONNECTED, _init_qsvs, calibration_func=gptq.calibrate,
    materialize_func=functools.partial(common_utils.materialize_weight_op,
                                       gptq.get_tensor_quant_params, inputs_to_ignore=()),
    update_qsv_func=qsv_utils.gptq_and_moving_average_update)
register_op_quant_config_validation_func(AlgorithmName.GPTQ, _check_config)
register_config_check_policy_func(AlgorithmName.GPTQ, None)

# Recovery of QAT (fake-quantised) weights: activations calibrate like min-max
# (algorithm_manager.py:272-293 binds dequantized_weight_recovery.calibrate / init_qsvs).
register_weight_algorithm(AlgorithmName.DEQUANTIZED_WEIGHT_RECOVERY,
                          dequantized_weight_recovery.get_tensor_quant_params,
                          dequantized_weight_recovery.calibrate)
# fp16 weights behind DEQUANTIZE: no calibration, no statistics (algorithm_manager.py:242-269).
for _op in sorted(float_casting.SUPPORTED_WEIGHT_QUANT_OPS, key=lambda o: o.value):
  register_quantized_op(
      AlgorithmName.FLOAT_CASTING, _op, _init_qsvs, calibration_func=None,
      materialize_func=float_casting.materialize_weight_op)
register_op_quant_config_validation_func(AlgorithmName.FLOAT_CASTING,
                                         float_casting.check_op_quantization_config)
register_config_check_policy_func(AlgorithmName.FLOAT_CASTING, None)

# OSCAR: FULLY_CONNECTED only, its own materialiser (INSERT_MULTIPLY on the activation) and the
# mu2-carrying QSV merge (algorithm_manager.py:453-480).
register_quantized_op(
    AlgorithmName.OSCAR, _Op.FULLY_CONNECTED, _init_qsvs, calibration_func=oscar.calibrate,
    materialize_func=oscar.materialize_fully_connected,
    update_qsv_func=qsv_utils.oscar_and_moving_average_update)
register_op_quant_config_validation_func(AlgorithmName.OSCAR, _check_config)
register_config_check_policy_func(AlgorithmName.OSCAR, None)
