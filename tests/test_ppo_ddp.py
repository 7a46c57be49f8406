"""Data-parallel ppo semantics on CPU with gloo, world_size 2: rank r holds rows
{r, r+2, ...} of the batch; the meanstd normalisers average their batch moments over the ranks
(embodied/jax/utils.py:76-81 pmean) and the gradients are averaged (embodied/jax/opt.py:52-54), so
every rank must end up with the gradients and normaliser state of the single-process run on the
whole batch.  Runs the device-independent torch code of the agent; the advantage kernel (CUDA
only, no CPU path in the product) is replaced by a test double with the reference formulation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as po
import ppo_cases as cases


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _gae_double(rew, val, last, term, hor, lam):            # ppo/agent.py:194-201
  rew, val = rew.detach(), val.detach()
  live = (~term).float()[:, 1:] * (1 - 1 / hor)
  cont = (~last & ~term).float()[:, 1:] * lam
  delta = rew[:, 1:] + live * val[:, 1:] - val[:, :-1]
  advs = [torch.zeros_like(delta[:, 0])]
  for t in reversed(range(delta.shape[1])):
    advs.append(delta[:, t] + live[:, t] * cont[:, t] * advs[-1])
  adv = torch.stack(list(reversed(advs))[:-1], 1)
  return adv, adv + val[:, :-1]


def _host_agent(vals, ocfg, obs, act, world):
  """The agent's loss code on CPU tensors: an Agent object without its device-only members."""
  from embodied_b200.dreamerv3 import params as P
  from embodied_b200.ppo import agent as A
  A.gae = _gae_double
  cfg = cases.product_config(ocfg)
  cfg.setdefault('norm_eps', 1e-4)
  agent = object.__new__(A.Agent)
  agent.cfg, agent.obs_space, agent.act_space = cfg, dict(obs), dict(act)
  agent.cd, agent.world = torch.float32, world
  agent.store = P.ParamStore(cfg, 'cpu', torch.float32, 0, {k: v.numpy() for k, v in vals.items()},
                             specs=A.param_specs(cfg, obs, act))
  agent.model = A.Model(cfg, obs, act, agent.store)
  agent.advnorm = A.Normalize(cfg.norm_rate, cfg.norm_limit, 'cpu', world)
  agent.valnorm = A.Normalize(cfg.norm_rate, cfg.norm_limit, 'cpu', world)
  return agent


def _grads(agent, data, B):
  zeros = {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32)
           for k, v in agent.act_space.items()}
  memory, prevact, data = agent._context((torch.zeros(B, agent.cfg.rnn_units), zeros), data)
  agent.store.begin_step()
  agent.store.grad.zero_()
  total, _, metrics, _ = agent.loss(memory, data, prevact)
  total.backward()
  return total.detach()


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  torch.set_num_threads(2)
  obs, act = cases.vector_spaces()
  ocfg = po.tiny_config()
  _, vals = cases.oracle_for(ocfg, obs, act)
  B, T = 4, 6
  data = cases.batch(ocfg, obs, act, B, T, seed=3)
  rows = torch.arange(rank, B, world)
  agent = _host_agent(vals, ocfg, obs, act, world)
  _grads(agent, {k: v[rows] for k, v in data.items()}, len(rows))
  dist.all_reduce(agent.store.grad, op=dist.ReduceOp.SUM)
  agent.store.grad.div_(world)
  if rank == 0:
    whole = _host_agent(vals, ocfg, obs, act, 1)
    _grads(whole, data, B)
    out['grad_err'] = float((agent.store.grad - whole.store.grad).abs().max())
    out['grad_max'] = float(whole.store.grad.abs().max())
    out['norm_err'] = float(torch.cat([agent.valnorm.vars - whole.valnorm.vars,
                                       agent.advnorm.vars - whole.advnorm.vars]).abs().max())
    out['norm_max'] = float(whole.valnorm.vars.abs().max())
  dist.barrier()
  dist.destroy_process_group()


def test_two_ranks_equal_the_whole_batch():
  ctx = mp.get_context('spawn')
  out = ctx.Manager().dict()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(300)
    assert p.exitcode == 0
  assert out['grad_max'] > 0
  assert out['grad_err'] <= 1e-5 * out['grad_max'] + 1e-8, dict(out)
  assert out['norm_err'] <= 1e-6 * max(out['norm_max'], 1.0), dict(out)
