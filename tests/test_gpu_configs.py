"""The other BASELINE.json configs as parity cases (bench.py measures config 2):

  config 3  Atari-shape rows (84x84x1 u8, 18 actions), many envs, replay sharded by
            `worker % R`; dreamerv3 itself at the reference's 96x96x1
  config 5  replay sample sweep B in {8..128} x T in {16..256}, image-only rows

Byte-exact against the pinned oracle (oracle/host_oracle.py); the dreamerv3 case
against the fp32 oracle restatement at 1e-5."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import embodied_b200 as embodied                          # noqa: E402
from embodied_b200 import dreamerv3, elements             # noqa: E402
from oracle import dreamer_oracle as do, host_oracle      # noqa: E402
import dreamer_cases as cases                             # noqa: E402


def atari_step(rng, t, length=50):
  return {
      'image': rng.integers(0, 256, (84, 84, 1)).astype(np.uint8),
      'reward': np.float32(rng.standard_normal()),
      'is_first': np.asarray(t % length == 0), 'is_last': np.asarray(t % length == length - 1),
      'is_terminal': np.asarray(False), 'action': np.int32(rng.integers(0, 18))}


@pytest.mark.parametrize('ranks', [1, 8])
def test_config3_atari_rows_sharded_replay(ranks):
  """64 workers (a slice of the 1024 envs), shard r holds workers w % ranks == r:
  every shard equals an oracle replay fed the same workers' streams."""
  workers, steps, L = 64, 40, 17
  elements.UUID.reset(debug=True)
  shards = [embodied.Replay(L, 4096, chunksize=32, online=True, seed=0, staging_rows=16)
            for _ in range(ranks)]
  oracles = [host_oracle.OracleReplay(L, 4096, 32, True, 0, ids=itertools.count(1000 * r + 1))
             for r in range(ranks)]
  rng = np.random.default_rng(3)
  try:
    for t in range(steps):
      for w in range(workers):
        step = atari_step(rng, t)
        r = w % ranks
        shards[r].add(dict(step), w // ranks)
        oracles[r].add(dict(step), w // ranks)
    for r in range(ranks):
      assert len(shards[r]) == len(oracles[r]) == (workers // ranks) * (steps - L + 1)
      a, b = shards[r].sample(8), oracles[r].sample(8)
      for k in b:
        if k != 'stepid':            # uuids differ between the two id generators
          assert a[k].cpu().numpy().tobytes() == b[k].tobytes(), (r, k)
  finally:
    elements.UUID.reset(debug=False)


@pytest.mark.parametrize('B,T', [(8, 16), (128, 16), (16, 64), (8, 256), (32, 128)])
def test_config5_sample_sweep_image_rows(B, T):
  """Replay sample sweep: image-only 64x64x3 rows, windows of T + 1, cross-chunk."""
  L = T + 1
  workers = 8
  steps = L + max(4, 2 * B // workers)
  elements.UUID.reset(debug=True)
  try:
    rep = embodied.Replay(L, None, chunksize=64, online=False, seed=1, staging_rows=8)
    ora = host_oracle.OracleReplay(L, None, 64, False, 1, ids=itertools.count(1))
    rng = np.random.default_rng(B * T)
    for t in range(steps):
      for w in range(workers):
        step = {'image': rng.integers(0, 256, (64, 64, 3)).astype(np.uint8),
                'is_first': np.asarray(t == 0), 'is_last': np.asarray(False)}
        rep.add(dict(step), w)
        ora.add(dict(step), w)
    a, b = rep.sample(B), ora.sample(B)
    assert a['image'].shape == (B, L, 64, 64, 3)
    for k in b:
      assert a[k].cpu().numpy().tobytes() == b[k].tobytes(), k
  finally:
    elements.UUID.reset(debug=False)


def test_config3_dreamerv3_on_96x96x1():
  """dreamerv3 on the reference's Atari image shape (96x96 gray, 18 actions): one
  update against the fp32 oracle restatement (1e-5 on the loss and latents)."""
  ocfg = do.tiny_config(image=(96, 96, 1), actions=18)
  vals = do.init_params(ocfg, 4, outscale_override=1.0)
  oracle = do.Dreamer(ocfg, {k: v.clone() for k, v in vals.items()})
  S = elements.Space
  obs = {'image': S(np.uint8, ocfg.image), 'reward': S(np.float32), 'is_first': S(bool),
         'is_last': S(bool), 'is_terminal': S(bool)}
  act = {'reset': S(bool), 'action': S(np.int32, (), 0, ocfg.actions)}
  agent = dreamerv3.Agent(obs, act, cases.product_config(ocfg, 'float32'),
                          values={k: v.numpy() for k, v in vals.items()})
  B, T = 2, 4
  data, noise = cases.batch(ocfg, B, T, seed=8), do.make_noise(ocfg, B, T, seed=9)
  ocarry, oouts, omets, _, oo = oracle.train(data, noise)
  carry, outs, mets = agent.train(agent.init_train(B), cases.to_device(data), cases.to_device(noise))
  rel = lambda a, b: float((a.detach().float().cpu() - b).abs().max() / (b.abs().max() + 1e-12))
  assert rel(mets['loss'], omets['loss']) < 1e-5
  assert rel(outs['replay']['dyn/deter'], oouts['replay']['dyn/deter']) < 1e-5
  assert torch.equal(agent.last_outs['imgact'].cpu(), oo['imgact'])


def test_config4_proprio_continuous_actions_driver_and_replay():
  """DMC-proprio-shaped dummy (vector observations, one f32[6] action in [-1, 1]),
  many envs, a host policy (the director agent's role): the Driver's stacked
  observations, masked actions and the replay contents equal the oracle's."""
  import functools
  from embodied_b200.envs import synthetic
  n, L = 64, 6                                   # a slice of the 512 envs
  mk = lambda i: synthetic.SyntheticProprio(i, length=7 + i % 5)
  elements.UUID.reset(debug=True)
  try:
    replay = embodied.Replay(L, 512, chunksize=16, online=True, seed=0, staging_rows=n)
    driver = embodied.Driver([functools.partial(mk, i) for i in range(n)], parallel=False)
    driver.on_step(replay.add)
    oenvs = [mk(i) for i in range(n)]
    oreplay = host_oracle.OracleReplay(L, 512, 16, True, 0, ids=itertools.count(1))
    odriver = host_oracle.OracleDriver(oenvs, oenvs[0].act_space)
    odriver.callbacks.append(lambda row, w: oreplay.add(row, w))

    def policy(carry, obs, **kw):
      ori = np.asarray(obs['orientations'].cpu() if hasattr(obs['orientations'], 'cpu') else obs['orientations'])
      act = np.tanh(ori[:, :6] * 2).astype(np.float32)
      return carry, {'action': act}, {}
    driver.reset(lambda k: ())
    for it in range(20):
      driver(policy, steps=n)
      odriver.step(lambda carry, obs: policy(carry, obs))
      for k, v in odriver.acts.items():
        assert driver.acts[k].dtype == v.dtype and (np.asarray(driver.acts[k]) == v).all(), (it, k)
      assert len(replay) == len(oreplay)
    a, b = replay.sample(16), oreplay.sample(16)
    assert sorted(a) == sorted(b)
    for k in b:
      got = a[k].cpu().numpy()
      assert got.dtype == b[k].dtype and got.tobytes() == b[k].tobytes(), k
    assert a['action'].shape == (16, L, 6) and a['orientations'].shape == (16, L, 14)
  finally:
    elements.UUID.reset(debug=False)
