"""Driver on the GPU: masks through the device kernel (golden sequences of the
real reference) and the fused device-agent step against OracleDriver +
OracleReplay fed the same policy outputs."""
import itertools
from functools import partial as bind

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import embodied_b200 as embodied            # noqa: E402
from embodied_b200 import elements          # noqa: E402
from embodied_b200.envs import dummy, synthetic   # noqa: E402
from oracle import host_oracle              # noqa: E402
import test_driver_host                     # noqa: E402


def test_golden_sequence_with_device_mask():
  fns = [bind(dummy.Dummy, 'disc', size=(8, 8), length=3 + i) for i in range(3)]
  test_driver_host.check_driver_golden(embodied.Driver(fns, parallel=False))


class ToyDeviceAgent:
  """Deterministic device 'policy': action and latents are functions of the
  normalised image, computed with torch on the device."""

  device_obs = True

  def __init__(self, classes, latent):
    self.classes, self.latent = classes, latent
    self.ext_space = {'dyn/deter': elements.Space(np.float32, (latent,))}
    self.seen = []

  def init_policy(self, n):
    return torch.zeros(n, device='cuda')

  def policy(self, carry, obs, mode='train'):
    norm = obs.normalized['image']
    assert norm.dtype == torch.float32 and norm.is_cuda
    self.seen.append(norm.cpu().numpy())
    score = norm.flatten(1).sum(1)
    act = (score.abs() * 1000).to(torch.int64).remainder(self.classes).to(torch.int32) + 1
    deter = score[:, None] + torch.arange(self.latent, device='cuda')[None]
    return carry + 1, {'action': act}, {'dyn/deter': deter.float()}


def host_twin(latent, classes):
  def policy(carry, obs):
    norm = host_oracle.normalize_image(obs['image'])
    t = torch.from_numpy(norm).cuda()
    score = t.flatten(1).sum(1)
    act = (score.abs() * 1000).to(torch.int64).remainder(classes).to(torch.int32) + 1
    deter = score[:, None] + torch.arange(latent, device='cuda')[None]
    return carry, {'action': act.cpu().numpy()}, {
        'dyn/deter': deter.float().cpu().numpy()}
  return policy


def test_fused_step_matches_oracle_driver_and_replay():
  n, latent, classes, L = 6, 48, 5, 4
  mk = lambda i: synthetic.SyntheticImage(i, size=(16, 16, 3), classes=classes + 2, length=5 + i)
  replay = embodied.Replay(L, 40, chunksize=8, online=True, seed=0, staging_rows=8)
  driver = embodied.Driver([bind(mk, i) for i in range(n)], parallel=False)
  driver.on_step(replay.add)
  agent = ToyDeviceAgent(classes, latent)
  driver.reset(agent.init_policy)

  oenvs = [mk(i) for i in range(n)]
  oreplay = host_oracle.OracleReplay(L, 40, 8, True, 0, ids=itertools.count(1))
  odriver = host_oracle.OracleDriver(oenvs, oenvs[0].act_space)
  odriver.callbacks.append(lambda row, w: oreplay.add(row, w))
  twin = host_twin(latent, classes)

  for it in range(25):
    driver(agent.policy, steps=n)
    trans = odriver.step(twin)
    assert (agent.seen[-1] == host_oracle.normalize_image(trans['image'])).all()
    for k, v in odriver.acts.items():
      assert driver.acts[k].dtype == v.dtype, k
      assert (driver.acts[k] == v).all(), (it, k)
    assert len(replay) == len(oreplay)
    if len(oreplay) and it % 3 == 2:
      a, b = replay.sample(3), oreplay.sample(3)
      assert sorted(a) == sorted(b)
      for k in b:
        got = a[k].cpu().numpy()
        assert got.dtype == b[k].dtype and got.tobytes() == b[k].tobytes(), k
