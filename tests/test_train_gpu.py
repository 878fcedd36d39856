# coding=utf-8
"""GPU parity tests of the backward (BPTT) kernels against torch autograd on the oracle's
torch-CPU restatement (fp64)."""
import numpy as np
import pytest
import torch

import cases
from oracle import multiverse_ref_torch as RT

pytestmark = pytest.mark.gpu
GTOL = 2e-4   # gradients: relative to the largest entry of each gradient tensor


def rel(a, b):
  a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def dev():
  from multiverse_b200 import build
  build.build()
  return torch.device("cuda:0")


def T(a, dev):
  return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("name", ["dec_cx32", "enc_class_cx64", "tile_edge", "enc_reg_cx2"])
def test_cell_backward_matches_autograd(dev, name):
  from multiverse_b200 import ops
  d = cases.cell_case(name)
  ns, h, w, cx = d["x"].shape
  rng = np.random.default_rng(5)
  dh = rng.standard_normal((ns, h, w, 256)).astype(np.float32)
  dc = rng.standard_normal((ns, h, w, 256)).astype(np.float32)
  # truth: autograd through the torch restatement
  t = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in d.items()}
  c1, h1 = RT.convlstm_cell(t["x"], t["c"], t["h"], t["kernel"], t["biases"])
  ((h1 * torch.from_numpy(dh).double()).sum() + (c1 * torch.from_numpy(dc).double()).sum()).backward()
  comp = name == "enc_reg_cx2"
  planes = 2
  pk = ops.PackedCell(T(d["kernel"], dev), T(d["biases"], dev), planes, comp=comp)
  wd = ops.pack_dgrad(pk, T(d["kernel"], dev))
  xh = ops.alloc_xh(ns, h, w, pk.cpad, planes, dev)
  ops.nhwc_to_planes(T(d["x"], dev), xh, 0, h, w, comp=pk.comp)
  ops.nhwc_to_planes(T(d["h"], dev), xh, pk.cxp, h, w)
  c_in = ops.alloc_state(ns, h, w, dev); ops.nhwc_to_halo(T(d["c"], dev), c_in, h, w)
  c_out = ops.alloc_state(ns, h, w, dev); h_out = ops.alloc_state(ns, h, w, dev)
  R = ops.halo_rows(ns, h, w)
  gates = torch.zeros((R, 1024), device=dev)
  ops.cell_fwd_train(xh, pk, c_in, c_out, h_out, None, gates, h, w, ns)
  dh_h = ops.alloc_state(ns, h, w, dev); ops.nhwc_to_halo(T(dh, dev), dh_h, h, w)
  dc_h = ops.alloc_state(ns, h, w, dev); ops.nhwc_to_halo(T(dc, dev), dc_h, h, w)
  dg = torch.zeros((planes, R, 1024), dtype=torch.bfloat16, device=dev)
  dc_prev = ops.alloc_state(ns, h, w, dev)
  dbp = torch.zeros((1024,), device=dev)
  ops.lstm_gates_bwd(gates, c_in, c_out, dh_h, dc_h, dg, dc_prev, dbp, h, w, ns)
  # dc_{t-1}
  dcp = torch.empty((ns, h, w, 256), device=dev); ops.halo_to_nhwc(dc_prev, dcp, h, w)
  assert rel(dcp.cpu().numpy(), t["c"].grad.numpy()) < GTOL
  # dgrad
  dxh = torch.zeros((R, pk.cpad), device=dev)
  ops.cell_dgrad(dg, wd, dxh, h, w, ns)
  dxh_v = dxh.view(ns, h + 1, w + 1, pk.cpad)[:, :h, :w].cpu().numpy()
  assert rel(dxh_v[..., pk.cxp:], t["h"].grad.numpy()) < GTOL
  if not comp:
    assert rel(dxh_v[..., :cx], t["x"].grad.numpy()) < GTOL
  # wgrad
  Rp = (R + 7) // 8 * 8
  dgT = torch.zeros((planes, 1024, Rp), dtype=torch.bfloat16, device=dev)
  xhT = torch.zeros((planes, 9, pk.cpad, Rp), dtype=torch.bfloat16, device=dev)
  ops.transpose_planes(dg, dgT); ops.transpose_planes(xh, xhT, taps=9, w=w)
  assert torch.equal(xhT[:, 4, :, :R].transpose(1, 2).contiguous(), xh)
  assert torch.equal(dgT[:, :, :R].transpose(1, 2).contiguous(), dg)
  dwp = torch.zeros((1024, 9 * pk.cpad), device=dev)
  ops.cell_wgrad(dgT, xhT, dwp, h, w, ns)
  ops.cell_wgrad(dgT, xhT, dwp, h, w, ns)          # accumulates: twice -> 2x
  dk = torch.empty((3, 3, cx + 256, 1024), device=dev); db = torch.empty((1024,), device=dev)
  ops.unpack_cell_wgrad(dwp, dbp, dk, db, cx, comp=pk.comp)
  assert rel(0.5 * dk.cpu().numpy(), t["kernel"].grad.numpy()) < GTOL
  assert rel(db.cpu().numpy(), t["biases"].grad.numpy()) < GTOL
