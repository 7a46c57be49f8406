"""Data-parallel learner semantics on CPU with gloo, world_size 2:
rank r holds rows {r, r+2, ...} of the batch; after the gradient all-reduce
(AVG, embodied/jax/opt.py:52-54 pmean) and the all-gathered return percentiles
(embodied/jax/utils.py:83-88) every rank must hold the gradients of the single
process run on the whole batch.  Runs the device-independent torch path of the
model (fused kernels off); the NCCL path is the same code on CUDA tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dreamer_oracle as do
import dreamer_cases as cases


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _model(vals, ocfg):
  from embodied_b200.dreamerv3 import model as M, params as P
  cfg = cases.product_config(ocfg)
  cfg['fused_scan'] = False
  cfg['fused_norm'] = False
  store = P.ParamStore(cfg, 'cpu', torch.float32, 0, {k: v.numpy() for k, v in vals.items()})
  return store, M.Model(cfg, store)


def _grads(store, model, oracle, data, noise):
  carry, obs, prevact, _ = oracle.apply_replay_context(data)
  store.begin_step()
  store.grad.zero_()
  total, *_ = model.loss((carry['deter'], carry['stoch']), obs, prevact, noise)
  total.backward()
  return total.detach()


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  torch.set_num_threads(2)
  ocfg = do.tiny_config()
  vals = do.init_params(ocfg, 0, outscale_override=1.0)
  B, T = 4, 5
  data, noise = cases.batch(ocfg, B, T, seed=3), do.make_noise(ocfg, B, T, seed=4)
  rows = torch.arange(rank, B, world)
  K = T
  irows = (rows[:, None] * K + torch.arange(K)[None]).reshape(-1)     # imagination starts of these rows
  mine = {k: v[rows] for k, v in data.items()}
  mynoise = dict(observe=noise['observe'][rows], imag_stoch=noise['imag_stoch'][irows],
                 imag_act=noise['imag_act'][irows])
  store, model = _model(vals, ocfg)
  _grads(store, model, do.Dreamer(ocfg, vals), mine, mynoise)
  dist.all_reduce(store.grad, op=dist.ReduceOp.SUM)
  store.grad /= world                                                 # AVG (gloo has no AVG op)
  if rank == 0:
    np.save(out, store.grad.numpy())
    np.save(out + '.ret.npy', np.array([float(model.ret_lo), float(model.ret_hi)]))
  dist.destroy_process_group()


def test_two_ranks_equal_one_process(tmp_path):
  out = str(tmp_path / 'grads.npy')
  mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  ocfg = do.tiny_config()
  vals = do.init_params(ocfg, 0, outscale_override=1.0)
  B, T = 4, 5
  data, noise = cases.batch(ocfg, B, T, seed=3), do.make_noise(ocfg, B, T, seed=4)
  store, model = _model(vals, ocfg)
  _grads(store, model, do.Dreamer(ocfg, vals), data, noise)
  got, want = np.load(out), store.grad.numpy()
  err = np.linalg.norm(got - want) / np.linalg.norm(want)
  assert err < 1e-5, err
  ret = np.load(out + '.ret.npy')
  np.testing.assert_allclose(ret, [float(model.ret_lo), float(model.ret_hi)], rtol=1e-5, atol=1e-7)
