"""Ready-made weight quantisation recipes (mirror of ai_edge_quantizer/recipe.py:27-330 for the
dynamic-range and weight-only families this package can apply end to end)."""
from __future__ import annotations

from . import qtyping
from . import recipe_manager

AlgorithmName = recipe_manager.AlgorithmName
_Gran = qtyping.QuantGranularity
_ALL = qtyping.TFLOperationName.ALL_SUPPORTED


def _dynamic(num_bits, granularity, algorithm_key):
  rm = recipe_manager.RecipeManager()
  rm.add_dynamic_config(".*", _ALL, num_bits, granularity, algorithm_key)
  return rm.get_quantization_recipe()


def _weight_only(num_bits, granularity, algorithm_key):
  rm = recipe_manager.RecipeManager()
  rm.add_weight_only_config(".*", _ALL, num_bits, granularity, algorithm_key)
  return rm.get_quantization_recipe()


def dynamic_wi8_afp32(algorithm_key=AlgorithmName.MIN_MAX_UNIFORM_QUANT):
  """int8 channelwise weights, float activations quantised on the fly (recipe.py:88-105)."""
  return _dynamic(8, _Gran.CHANNELWISE, algorithm_key)


def dynamic_wi4_afp32(algorithm_key=AlgorithmName.MIN_MAX_UNIFORM_QUANT):
  """int4 channelwise weights (recipe.py:108-125)."""
  return _dynamic(4, _Gran.CHANNELWISE, algorithm_key)


def dynamic_wi4b32_afp32(algorithm_key=AlgorithmName.MIN_MAX_UNIFORM_QUANT):
  """int4 weights in blocks of 32 along the input features (recipe.py:285-305)."""
  return _dynamic(4, _Gran.BLOCKWISE_32, algorithm_key)


def weight_only_wi8_afp32(algorithm_key=AlgorithmName.MIN_MAX_UNIFORM_QUANT):
  """int8 weights behind an explicit DEQUANTIZE (recipe.py:128-147)."""
  return _weight_only(8, _Gran.CHANNELWISE, algorithm_key)


def weight_only_wi4_afp32(algorithm_key=AlgorithmName.MIN_MAX_UNIFORM_QUANT):
  return _weight_only(4, _Gran.CHANNELWISE, algorithm_key)
