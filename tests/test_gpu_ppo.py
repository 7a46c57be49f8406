"""ppo on the device against the oracle restatement (fp32, 1e-5): policy steps, updates, the two
kernels on their own, and the agent inside Driver + Replay + run.train on the dummy env
(BASELINE config 1)."""
import math
import pathlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ppo_oracle as po
import ppo_cases as cases

GOLDEN = pathlib.Path(__file__).parent / 'golden' / 'ppo_tiny.npz'
RTOL, ATOL = 1e-5, 2e-6


def _pair(spaces=cases.dummy_spaces, **over):
  from embodied_b200 import ppo
  obs, act = spaces()
  ocfg = po.tiny_config(**over)
  oracle, vals = cases.oracle_for(ocfg, obs, act)
  agent = ppo.Agent(obs, {**act}, cases.product_config(ocfg), values={k: v.numpy() for k, v in vals.items()})
  return ocfg, obs, act, oracle, agent


def _close(a, b, name, rtol=RTOL, atol=ATOL):
  a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
  b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
  np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=name)


def test_gae_kernel_bit_exact():
  from embodied_b200.ppo import agent as agentlib
  g = torch.Generator().manual_seed(0)
  B, T = 37, 65
  rew, val = torch.randn(B, T, generator=g), torch.randn(B, T, generator=g) * 3
  last = torch.rand(B, T, generator=g) < 0.1
  term = last & (torch.rand(B, T, generator=g) < 0.5)
  hor, lam = 200, 0.8
  live = (~term).float()[:, 1:] * (1 - 1 / hor)
  cont = (~last & ~term).float()[:, 1:] * lam
  delta = rew[:, 1:] + live * val[:, 1:] - val[:, :-1]
  advs = [torch.zeros(B)]
  for t in reversed(range(T - 1)):
    advs.append(delta[:, t] + live[:, t] * cont[:, t] * advs[-1])
  adv = torch.stack(list(reversed(advs))[:-1], 1)
  a, t = agentlib.gae(rew.cuda(), val.cuda(), last.cuda(), term.cuda(), hor, lam)
  assert torch.equal(a.cpu(), adv)
  assert torch.equal(t.cpu(), adv + val[:, :-1])


@pytest.mark.parametrize('wd', [0.0, 0.03])
def test_clip_adam_kernel_matches_optax_chain(wd):
  ocfg, obs, act, oracle, agent = _pair(warmup=3, clip=0.7, wd=wd)
  g = torch.Generator().manual_seed(1)
  for step in range(5):
    grads = {k: torch.randn(v.shape, generator=g) * (0.01 if step % 2 else 1.0) for k, v in oracle.p.items()}
    for k, v in grads.items():
      agent.store.view('grad', k).copy_(v)
    mets = oracle.apply_updates(grads)
    norm = agent.opt.launch()
    _close(norm, mets['opt/grad_norm'], f'norm{step}', rtol=1e-6)
    for k in oracle.p:
      _close(agent.store.view('master', k), oracle.p[k], f'{step}/{k}', rtol=1e-5, atol=1e-7)
  assert float(agent.opt.state[0]) == 5


@pytest.mark.parametrize('spaces', [cases.dummy_spaces, cases.small_spaces])
def test_policy_matches_oracle(spaces):
  ocfg, obs, act, oracle, agent = _pair(spaces)
  B = 4
  g = torch.Generator().manual_seed(5)
  oc = (oracle.initial(B), {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32)
                            for k, v in act.items()})
  pc = agent.init_policy(B)
  for step in range(4):
    o = cases.obs_batch(obs, (B,), g)
    o['is_first'][:] = step == 0
    o['is_first'][1] = step == 2
    noise = po.make_noise(act, (B,), 10 + step)
    oc, oacts, oext = oracle.policy(oc, o, noise)
    pc, pacts, pext = agent.policy(pc, cases.to_device(o), noise=cases.to_device(noise))
    assert set(pext) == set(oext) and set(pacts) == set(oacts)
    for k, v in oacts.items():
      if v.dtype == torch.int32:
        assert torch.equal(pacts[k].cpu(), v), (step, k)
      else:
        _close(pacts[k], v, f'{step}/{k}')
    for k, v in oext.items():
      _close(pext[k], v, f'{step}/{k}')
      assert not pext[k].requires_grad


@pytest.mark.parametrize('spaces,over', [
    (cases.dummy_spaces, {}), (cases.small_spaces, {}),
    (cases.vector_spaces, dict(enc_norm='layer', enc_layers=3, wd=0.01)),
    (cases.small_spaces, dict(recurrent=False)), (cases.dummy_spaces, dict(rnnact=False, replay_context=0))])
def test_updates_match_oracle(spaces, over):
  ocfg, obs, act, oracle, agent = _pair(spaces, warmup=2, **over)
  B, T = 3, 8
  oc = (oracle.initial(B), {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32)
                            for k, v in act.items()})
  pc = agent.init_train(B)
  for step in range(3):
    data = cases.batch(ocfg, obs, act, B, T, seed=20 + step)
    oc, _, omets, ograds, olosses = oracle.train(oc, data)
    pc, outs, pmets = agent.train(pc, cases.to_device(data))
    assert outs == {}
    assert set(pmets) == set(omets), set(pmets) ^ set(omets)
    # the first update starts from identical parameters: 1e-5.  Later ones inherit the last bits in
    # which two evaluations of the x255-scaled image branch's gradients differ (run to run as well):
    # 1e-4, per-element losses relative to the largest element (squared errors near zero cancel)
    mtol = 1e-5 if step == 0 else 1e-4
    for k in omets:
      loose = 'std' in k or k == 'opt/grad_norm'       # sums over the x255-scaled encoder gradients
      _close(pmets[k], omets[k], f'{step}/{k}', rtol=10 * mtol if loose else mtol, atol=1e-6)
    for k, v in olosses.items():
      _close(agent.last_losses[k], v, f'{step}/loss/{k}', rtol=mtol, atol=mtol * float(v.abs().max()) + 1e-6)
    # per tensor, relative to its largest entry
    tol = 1e-4
    for k, v in ograds.items():
      # the image branch's gradients are sums of large cancelling terms (same x255 scale): 1e-3 there
      t = max(tol, 1e-3) if k.startswith(('enc/s', 'enc/out')) else tol
      _close(agent.store.view('grad', k), v, f'{step}/grad/{k}', rtol=0, atol=t * max(float(v.abs().max()), 1e-9))
    for k, v in oracle.p.items():
      _close(agent.store.view('master', k), v, f'{step}/param/{k}', rtol=mtol, atol=max(tol * ocfg.lr, 2e-6 if step == 0 else 1e-5))
    if ocfg.recurrent:
      _close(pc[0], oc[0], f'{step}/memory', rtol=mtol, atol=2e-6 if step == 0 else 1e-5)


def test_layer_norm_on_feature_maps_forward():
  """enc.impala.norm = layer on the convolution blocks.  The reference's var = E[x^2] - E[x]^2 over
  the channel axis of activations it scaled UP by 255 (ppo/nets.py:46) cancels catastrophically in
  fp32 whenever channels are close (debug depth: 2 channels), so two correct fp32 evaluations only
  agree loosely; this checks the path (parameter names, axis, block order), not rounding."""
  ocfg, obs, act, oracle, agent = _pair(cases.small_spaces, enc_norm='layer')
  B = 4
  g = torch.Generator().manual_seed(7)
  o = cases.obs_batch(obs, (B,), g)
  o['cam'] = o['cam'] % 3
  noise = po.make_noise(act, (B,), 3)
  oc = (oracle.initial(B), {k: torch.zeros(B, *v.shape) for k, v in act.items()})
  _, oacts, oext = oracle.policy(oc, o, noise)
  _, pacts, pext = agent.policy(agent.init_policy(B), cases.to_device(o), noise=cases.to_device(noise))
  _close(pext['memory'], oext['memory'], 'memory', rtol=5e-2, atol=5e-3)
  _close(pacts['action'], oacts['action'], 'action', rtol=5e-2, atol=5e-3)


def test_committed_golden():
  """The same three policy steps + three updates the oracle's golden was written from."""
  ocfg, obs, act, oracle, agent = _pair(warmup=2)
  want = np.load(GOLDEN)
  B, T = 3, 8
  g = torch.Generator().manual_seed(5)
  pc = agent.init_policy(B)
  for step in range(3):
    o = cases.obs_batch(obs, (B,), g)
    o['is_first'][:] = step == 0
    pc, acts, ext = agent.policy(pc, cases.to_device(o), noise=cases.to_device(po.make_noise(act, (B,), 10 + step)))
    for k, v in {**acts, **ext}.items():
      _close(v, want[f'policy{step}/{k}'], f'policy{step}/{k}')
  pc = agent.init_train(B)
  for step in range(3):
    pc, _, mets = agent.train(pc, cases.to_device(cases.batch(ocfg, obs, act, B, T, seed=20 + step)))
    base = 2e-5 if step == 0 else 2e-4                 # later updates: see test_updates_match_oracle
    for k, v in mets.items():
      _close(float(v), want[f'train{step}/{k}'], f'train{step}/{k}', rtol=base if 'std' not in k else 5 * base, atol=1e-6)


def test_config1_train_loop_on_dummy_env(tmp_path):
  """BASELINE config 1: ppo on embodied.envs.dummy, 4 envs, Driver + Replay + run.train."""
  import embodied_b200 as embodied
  from embodied_b200 import ppo
  from embodied_b200.envs import dummy
  from embodied_b200.ppo import config as configlib
  obs, act = cases.dummy_spaces()
  cfg = configlib.debug(warmup=5)
  agent = ppo.Agent(obs, act, cfg)
  replay = embodied.Replay(12 + 1, 2000, chunksize=64, online=True, seed=0, staging_rows=8)
  driver = embodied.Driver([lambda: dummy.Dummy('disc', length=20)] * 4, parallel=False)
  driver.on_step(replay.add)
  driver.reset(agent.init_policy)
  driver(agent.policy, steps=4 * 40)
  batch = replay.sample(8)
  assert {'logp/act_disc', 'logp/act_cont', 'memory', 'stepid'} <= set(batch)
  assert tuple(batch['memory'].shape) == (8, 13, cfg.rnn_units)
  carry = agent.init_train(8)
  losses = []
  for _ in range(4):
    carry, outs, mets = agent.train(carry, replay.sample(8))
    losses.append(float(mets['loss']))
    assert math.isfinite(losses[-1])
  # the stored log-probabilities are those of the acting policy: at step 0 of training the ratio is 1
  data = replay.sample(8)
  assert float(mets['ratio']) > 0
  assert agent.updates == 4


def test_config_level_entry_point(tmp_path):
  """`python -m embodied_b200.ppo.main --configs debug --task dummy_disc`: the reference's
  ppo/main.py contract (config blocks -> factories -> run.train) on the device."""
  import json
  from embodied_b200.ppo import main as mainlib
  mainlib.main(['--configs', 'debug', '--task', 'dummy_disc', '--logdir', str(tmp_path),
                '--run.steps', '400', '--run.log_every', '-1', '--run.report_every', '1000',
                '--run.save_every', '1000', '--replay.size', '4000'])
  rows = [json.loads(l) for l in (tmp_path / 'metrics.jsonl').read_text().strip().splitlines()]
  keys = {k for row in rows for k in row}
  assert 'train/loss/policy_loss' in keys or any('policy_loss' in k for k in keys), sorted(keys)
  assert (tmp_path / 'checkpoint.pkl').exists()


def _run_ppo_workers(nproc, mode, port):
  import json, os, subprocess, sys
  root = pathlib.Path(__file__).resolve().parent.parent
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
         '--master-addr', '127.0.0.1', '--master-port', str(port), 'tests/ppo_worker.py', mode]
  proc = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600,
                        env=dict(os.environ, MASTER_ADDR='127.0.0.1'))
  assert proc.returncode == 0, proc.stderr[-3000:]
  rows = [json.loads(l) for l in proc.stdout.splitlines() if l.startswith('{')]
  assert len(rows) == nproc, proc.stdout[-2000:]
  return rows


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_ranks_average_gradients_and_normalisers():
  """Data-parallel ppo: gradient mean over ranks before the optimiser chain, batch
  moments of the meanstd normalisers averaged like the reference's pmean.  Ranks stay
  bit-identical; with the SAME batch on both ranks the result is the single-process oracle's."""
  own = _run_ppo_workers(2, 'own', 29621)
  for row in own:
    assert row['max_diff_across_ranks'] == 0.0 and row['norm_diff_across_ranks'] == 0.0, row
  assert own[0]['losses'] != own[1]['losses']                      # different batches went in
  same = _run_ppo_workers(2, 'same', 29623)
  for row in same:
    assert row['max_diff_across_ranks'] == 0.0, row
  assert same[0]['rel_diff_vs_oracle'] < 1e-5, same[0]
  assert abs(same[0]['checksum'] - own[0]['checksum']) > 0


def test_captured_update_matches_eager_launches():
  """The update replayed from its CUDA graph (after two eager warm-up steps) against the same
  updates launched eagerly: same parameters, metrics and carry."""
  from embodied_b200 import ppo
  obs, act = cases.dummy_spaces()
  ocfg = po.tiny_config(warmup=2)
  _, vals = cases.oracle_for(ocfg, obs, act)
  mk = lambda graph: ppo.Agent(obs, act, cases.product_config(ocfg, graph=graph),
                               values={k: v.numpy() for k, v in vals.items()})
  a, b = mk('auto'), mk('off')
  B, T = 3, 8
  ca, cb = a.init_train(B), b.init_train(B)
  for step in range(6):
    data = cases.to_device(cases.batch(ocfg, obs, act, B, T, seed=40 + step))
    ca, _, ma = a.train(ca, data)
    cb, _, mb = b.train(cb, data)
    assert set(ma) == set(mb)
    for k in mb:         # two runs of the library's convolution gradients differ in their last bits
      _close(ma[k], mb[k], f'{step}/{k}', rtol=1e-4, atol=1e-6)
    _close(ca[0], cb[0], f'{step}/memory', rtol=1e-4, atol=1e-6)
  assert len(a._graphs) == 1 and a._graph_ok and not b._graphs
  assert a.updates == b.updates == 6
  for k in a.store.specs:
    _close(a.store.view('master', k), b.store.view('master', k), k, rtol=1e-4, atol=1e-6)
  # the policy reads the parameters the replayed graph wrote
  g = torch.Generator().manual_seed(1)
  o = cases.to_device(cases.obs_batch(obs, (2,), g))
  noise = cases.to_device(po.make_noise(act, (2,), 3))
  _, acta, _ = a.policy(a.init_policy(2), o, noise=noise)
  _, actb, _ = b.policy(b.init_policy(2), o, noise=noise)
  _close(acta['act_cont'], actb['act_cont'], 'act_cont')


def test_bf16_update_tracks_the_fp32_oracle():
  """compute_dtype=bfloat16 (the reference's default): bf16 GEMMs / convolutions / GRU with fp32
  master weights, norms, heads and losses.  Tracks the fp32 oracle loosely; every tensor that
  gets a gradient in fp32 gets one in bf16, and the bf16 copy follows the optimiser."""
  from embodied_b200 import ppo
  obs, act = cases.vector_spaces()                  # no x255 image scale: see test_layer_norm_on_feature_maps
  ocfg = po.tiny_config(warmup=0, lr=1e-3)
  oracle, vals = cases.oracle_for(ocfg, obs, act)
  agent = ppo.Agent(obs, act, cases.product_config(ocfg, compute_dtype='bfloat16'),
                    values={k: v.numpy() for k, v in vals.items()})
  B, T = 3, 8
  zeros = {k: torch.zeros(B, *v.shape, dtype=torch.int32 if v.discrete else torch.float32) for k, v in act.items()}
  oc, pc = (oracle.initial(B), zeros), agent.init_train(B)
  assert pc[0].dtype == torch.bfloat16
  for step in range(4):
    data = cases.batch(ocfg, obs, act, B, T, seed=50 + step)
    oc, _, omets, ograds, _ = oracle.train(oc, data)
    pc, _, pmets = agent.train(pc, cases.to_device(data))
    assert abs(float(pmets['loss']) - float(omets['loss'])) <= 5e-2 * abs(float(omets['loss'])) + 5e-3, step
    if step == 0:
      for k, v in ograds.items():
        got = agent.store.view('grad', k)
        assert (float(got.abs().max()) > 0) == (float(v.abs().max()) > 0), k
  low = agent.store.low_buffer().float()
  assert float((low - agent.store.master).abs().max()) <= 8e-3 * float(agent.store.master.abs().max())
  out = agent.policy(agent.init_policy(2), cases.to_device(cases.obs_batch(obs, (2,), torch.Generator().manual_seed(0))))
  assert out[2]['memory'].dtype == torch.float32 and torch.isfinite(out[2]['memory']).all()
