"""GPU parity of OSCAR against reference-generated fixtures
(tests/golden/oscar.npz) and the oracle.

Tolerances.  The reference is float64 NumPy; the device path is float64 too, but its column /
group reductions and running sums are parallel (different association), so intermediate
values agree to ~1e-15 relative, not bitwise:
 * channel-scale multipliers (float32): <= 1e-6 relative,
 * clip bounds / per-channel scales (float64): <= 1e-9 relative,
 * blockwise scales are rounded to bf16: bit-equal except when a bound sits on a rounding
   boundary (<= 0.1 % of blocks may move by one bf16 ulp),
 * integers: |dq| <= 1 on <= 1e-4 of the elements (exactly 0 whenever the scales are equal and
   no quotient is within 1e-15 of a rounding tie).
"""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oscar.npz")


def _cfg(bits, gk, sym=True):
  from aeq_b200 import qtyping
  G = qtyping.QuantGranularity
  gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
          128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}[gk]
  return qtyping.TensorQuantizationConfig(num_bits=bits, symmetric=sym, granularity=gran)


def _run(w, cfg, mu2, op_name="FULLY_CONNECTED"):
  from aeq_b200.algorithms.uniform_quantize import oscar
  op, _ = sg.fc_graph(w)
  return oscar.get_tensor_quant_params(sg.op_info(op, cfg, op_name), cfg, w,
                                       None if mu2 is None else {"mu2": mu2})


def _check(r, want_scale, want_q, want_mult, blockwise):
  assert r.scale.dtype == want_scale.dtype and r.scale.shape == want_scale.shape
  if blockwise:
    moved = r.scale != want_scale
    assert moved.mean() <= 1e-3
    np.testing.assert_allclose(r.scale, want_scale, rtol=2.0 ** -7)
  else:
    np.testing.assert_allclose(r.scale, want_scale, rtol=1e-9)
  np.testing.assert_allclose(r.custom_algorithm_param["multiplier"], want_mult, rtol=1e-6)
  assert r.custom_algorithm_param["multiplier"].dtype == np.float32
  dq = np.abs(r.quantized_data.astype(np.int32) - want_q.astype(np.int32))
  assert dq.max() <= 1 and (dq != 0).mean() <= 1e-4, (dq.max(), (dq != 0).mean())
  assert r.quantized_data.dtype == want_q.dtype and not r.zero_point.any()


def test_oscar_golden(cuda):
  z = np.load(GOLD)
  for i, (bits, gk, with_mu2) in enumerate(z["cases"]):
    mu2 = z[f"mu2_{i}"] if with_mu2 else None
    r = _run(z[f"w{i}"], _cfg(int(bits), int(gk)), mu2)
    _check(r, z[f"scale{i}"], z[f"q{i}"], z[f"mult{i}"], gk > 0)
    assert r.quantized_dimension == (None if gk == -1 else (1 if gk > 0 else 0))
    assert r.block_size == max(int(gk), 0)


@pytest.mark.parametrize("shape,bits,gk,with_mu2", [
    ((128, 4096), 4, 0, True), ((128, 4096), 4, 32, True), ((64, 2048), 8, 256, True),
    ((32, 1000), 4, 0, True), ((16, 8192), 8, 0, False), ((256, 1024), 4, -1, True),
    ((1, 64), 4, 0, False), ((40, 96), 4, 32, False)])
def test_oscar_vs_oracle(cuda, shape, bits, gk, with_mu2):
  w = O.synthetic_weight(shape[0], shape[1], shape[1] % 11)
  rng = np.random.default_rng(shape[0])
  mu2 = None
  if with_mu2:
    mu2 = rng.standard_normal(shape[1]) ** 2 * 0.5 + 0.01
    mu2[::13] *= 50.0
  with np.errstate(all="ignore"):
    o = O.oscar_requant(w, mu2, bits, block=max(gk, 0), per_channel=(gk == 0))
  r = _run(w, _cfg(bits, gk), mu2)
  _check(r, o["scale"], o["q"], o["multiplier"], gk > 0)


def test_oscar_kernels_raw(cuda):
  """Each device stage against its oracle counterpart on the same inputs."""
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(96, 512, 5)
  w64 = w.astype(np.float64)
  rng = np.random.default_rng(1)
  s = np.exp(rng.standard_normal(512) * 0.3)
  m = O.oscar_floor(rng.standard_normal(512) ** 2)
  wd = torch.from_numpy(w).to(cuda)
  sd, md = torch.from_numpy(s).to(cuda), torch.from_numpy(m).to(cuda)
  # column second moments
  np.testing.assert_allclose(device.colsq(wd, 1.0).cpu().numpy(), (w64 * w64).sum(0), rtol=1e-13)
  x = O.synthetic_activation((3, 70, 96), 8)
  np.testing.assert_allclose(device.colsq(torch.from_numpy(x).to(cuda), 1.0 / 210).cpu().numpy(),
                             O.oscar_mu2(x), rtol=1e-13)
  # objective pass: group sums of squared maxima and the arg-max scatter
  for g in (512, 32):
    gsq, a_eff = device.oscar_pass(wd, sd, g)
    mag = np.abs(w64) * s
    want_sq = np.array([(mag[:, b * g:(b + 1) * g].max(1) ** 2).sum() for b in range(512 // g)])
    np.testing.assert_allclose(gsq.cpu().numpy(), want_sq, rtol=1e-13)
    want_eff = np.zeros(512)
    for b in range(512 // g):
      j = b * g + np.argmax(mag[:, b * g:(b + 1) * g], 1)
      np.add.at(want_eff, j, w64[np.arange(96), j] ** 2)
    np.testing.assert_allclose(a_eff.cpu().numpy(), want_eff, rtol=1e-13, atol=1e-300)
  # clip bounds: rows, blocks, whole tensor
  a = np.abs(w64 * s)
  got = device.oscar_clip(wd, sd, md, 512, 7, mass0=float(m.sum()) + 1e-12).cpu().numpy()
  np.testing.assert_allclose(got, O.oscar_group_clip(a, m, 7), rtol=1e-11)
  masses = np.array([float(m[b * 64:(b + 1) * 64].sum()) + 1e-12 for b in range(8)])
  got = device.oscar_clip(wd, sd, md, 64, 7, mass=torch.from_numpy(masses).to(cuda)).cpu().numpy()
  want = np.stack([O.oscar_group_clip(a[:, b * 64:(b + 1) * 64], m[b * 64:(b + 1) * 64], 7)
                   for b in range(8)], axis=1).ravel()
  np.testing.assert_allclose(got, want, rtol=1e-11)
  mt = np.tile(m, 96)
  got = device.oscar_clip(wd, sd, md, 96 * 512, 127, mass0=float(mt.sum()) + 1e-12).cpu().numpy()
  np.testing.assert_allclose(got, O.oscar_group_clip(a.reshape(1, -1), mt, 127), rtol=1e-10)


def test_oscar_errors_and_materialize(cuda):
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import oscar
  from aeq_b200.algorithms.utils import common_utils
  w = O.synthetic_weight(16, 64, 2)
  with pytest.raises(ValueError, match="symmetric weight quantization only"):
    _run(w, _cfg(4, 0, sym=False), None)
  with pytest.raises(ValueError, match="FULLY_CONNECTED only"):
    _run(w, _cfg(4, 0), None, op_name="CONV_2D")
  with pytest.raises(ValueError, match="mu2 has 5 channels"):
    _run(w, _cfg(4, 0), np.ones(5))
  op, graph = sg.fc_graph(w, bias=True)
  info = sg.op_info(op, _cfg(4, 0))
  mu2 = np.linspace(0.1, 5.0, 64)
  cache = common_utils.TensorQuantParamsCache()
  out = oscar.materialize_fully_connected(info, graph, {"input": {"mu2": mu2}}, cache)
  T = qtyping.QuantTransformation
  assert [p.tensor_name for p in out] == ["input", "weight", "bias", "output"]
  assert out[0].consumers[0].transformations == [T.INSERT_MULTIPLY]
  assert out[1].consumers[0].transformations == [T.QUANTIZE_TENSOR]
  assert out[0].consumers[0].parameters is out[1].consumers[0].parameters
  assert out[2].consumers[0].transformations == [T.NO_QUANTIZE]
  assert out[3].producer.transformations == [T.NO_QUANTIZE]
  with np.errstate(all="ignore"):
    o = O.oscar_requant(w, mu2, 4)
  _check(out[1].consumers[0].parameters, o["scale"], o["q"], o["multiplier"], False)
  assert len(cache) == 1
