"""Driver: steps N environments in lock-step and feeds policy + callbacks.

Same surface and step semantics as the reference ``embodied.core.Driver``
(embodied/core/driver.py:9-137): ``Driver(make_env_fns, parallel=True,
**kwargs)``, ``reset(init_policy)``, ``on_step(fn)``, ``__call__(policy, steps,
episodes)``, ``close()``; attributes ``length, act_space, acts, carry``.

Per step (driver.py:55-82): send ``acts[i]`` (incl. ``reset``) to env i, stack
the N observations, split ``log/`` keys, ``policy(carry, obs)``, zero the actions
of envs whose ``is_last`` is set, next ``reset = is_last``, then hand every
callback the per-env slice of ``{**obs, **acts, **outs, **logs}`` in env order.

What is different is where the bytes go.  The stack / cast / normalise of the
observations and the mask + scatter of the actions are two launches of the row
engine (``emb_driver_stage_obs`` / ``emb_driver_scatter_mask_actions``), fused
with the replay append when ``on_step(replay.add)`` registered a device replay:
observations go pinned-host -> HBM once, policy outputs (latents) never leave
the device.
"""
import time

import numpy as np

from .. import elements
from . import base


class _SerialEnvs:

  def __init__(self, fns):
    self.envs = [fn() for fn in fns]
    for env in self.envs:           # not a drop-in Env: say so before any device work
      lacking = base.implements_env(env)
      if lacking:
        raise TypeError(f'{type(env).__name__} does not implement the Env protocol '
                        f'(embodied/core/base.py:34-58): missing {lacking}')
    self.act_space = self.envs[0].act_space
    self.obs_space = self.envs[0].obs_space

  def step(self, acts):
    return [env.step(act) for env, act in zip(self.envs, acts)]

  def close(self):
    for env in self.envs:
      env.close()


def _env_worker(pipe, ctor, stop):
  # one process per env (reference driver.py:101-137)
  import cloudpickle
  env = None
  try:
    env = cloudpickle.loads(ctor)()
    while not stop.is_set():
      if not pipe.poll(0.1):
        continue
      try:
        msg, *args = pipe.recv()
      except EOFError:
        return
      if msg == 'step':
        pipe.send(('result', env.step(args[0])))
      elif msg == 'obs_space':
        pipe.send(('result', env.obs_space))
      elif msg == 'act_space':
        pipe.send(('result', env.act_space))
      else:
        raise ValueError(f'Invalid message {msg}')
  except ConnectionResetError:
    print('Connection to driver lost')
  except Exception as e:
    pipe.send(('error', e))
    raise
  finally:
    try:
      env and env.close()
    except Exception:
      pass
    pipe.close()


class _ProcessEnvs:

  def __init__(self, fns):
    import multiprocessing as mp
    import cloudpickle
    ctx = mp.get_context()
    self.stop = ctx.Event()
    self.pipes, remote = zip(*[ctx.Pipe() for _ in fns])
    self.procs = [
        ctx.Process(target=_env_worker, daemon=True,
                    args=(pipe, cloudpickle.dumps(fn), self.stop))
        for fn, pipe in zip(fns, remote)]
    [p.start() for p in self.procs]
    self.pipes[0].send(('act_space',))
    self.act_space = self._receive(self.pipes[0])
    self.pipes[0].send(('obs_space',))
    self.obs_space = self._receive(self.pipes[0])

  def step(self, acts):
    for pipe, act in zip(self.pipes, acts):
      pipe.send(('step', act))
    return [self._receive(pipe) for pipe in self.pipes]

  def _receive(self, pipe):
    try:
      msg, arg = pipe.recv()
      if msg == 'error':
        raise RuntimeError(arg)
      assert msg == 'result'
      return arg
    except Exception:
      print('Terminating workers due to an exception.')
      self.close()
      raise

  def close(self):
    self.stop.set()
    for proc in self.procs:
      proc.join(0.5)
      if proc.is_alive():
        proc.terminate()


class Driver:

  def __init__(self, make_env_fns, parallel=True, ops=None, fetch_outs=True,
               **kwargs):
    assert len(make_env_fns) >= 1
    self.parallel = parallel
    self.kwargs = kwargs
    self.length = len(make_env_fns)
    self._envs = (_ProcessEnvs if parallel else _SerialEnvs)(make_env_fns)
    self.act_space = self._envs.act_space
    self.callbacks = []
    self.batch_callbacks = []
    self.acts = None
    self.carry = None
    self._replay = None
    self._ops = ops
    self._fetch_outs = fetch_outs
    self.reset()

  @property
  def envs(self):
    return self._envs.envs

  def reset(self, init_policy=None):
    self.acts = {
        k: np.zeros((self.length,) + tuple(v.shape), v.dtype)
        for k, v in self.act_space.items()}
    self.acts['reset'] = np.ones(self.length, bool)
    self.carry = init_policy and init_policy(self.length)

  def close(self):
    self._envs.close()

  def on_step(self, callback):
    """Register fn(tran, worker).  A bound ``Replay.add`` of a device replay is
    recognised and served by one batched append per step instead of N calls."""
    from . import replay as replaylib
    owner = getattr(callback, '__self__', None)
    if (isinstance(owner, replaylib.Replay) and self._replay is None and
        getattr(callback, '__func__', None) is replaylib.Replay.add):
      self._replay = owner
      self.callbacks.append(owner)      # placeholder keeps registration order
    else:
      self.callbacks.append(callback)

  def on_batch(self, callback):
    """Register fn(trans, n), called once per Driver step with the stacked
    (N, ...) host transition, after the per-env callbacks."""
    self.batch_callbacks.append(callback)

  def __call__(self, policy, steps=0, episodes=0):
    step, episode = 0, 0
    while step < steps or episode < episodes:
      step, episode = self._step(policy, step, episode)

  # ----------------------------------------------------------------- one step
  def _ops_or_default(self):
    if self._ops is None:
      from . import driver_ops
      self._ops = driver_ops.DeviceOps()
    return self._ops

  def _step(self, policy, step, episode):
    acts = self.acts
    n = self.length
    assert all(len(x) == n for x in acts.values())
    assert all(isinstance(v, np.ndarray) for v in acts.values())
    per_env = [{k: v[i] for k, v in acts.items()} for i in range(n)]
    device_agent = bool(getattr(getattr(policy, '__self__', None),
                                'device_obs', False)) or bool(
                                    getattr(policy, 'device_obs', False))
    if device_agent and self._replay is not None:
      if not self._replay.store.configured:
        ext = getattr(getattr(policy, '__self__', policy), 'ext_space', None)
        self._replay.configure_spaces(
            self._envs.obs_space, self.act_space, ext, workers=self.length)
      return self._step_fused(policy, per_env, step, episode)
    obs = self._envs.step(per_env)
    obs = {k: np.stack([x[k] for x in obs]) for k in obs[0].keys()}
    logs = {k: v for k, v in obs.items() if k.startswith('log/')}
    obs = {k: v for k, v in obs.items() if not k.startswith('log/')}
    assert all(len(x) == n for x in obs.values()), obs
    self.carry, acts, outs = policy(self.carry, obs, **self.kwargs)
    assert all(k not in acts for k in outs), (list(outs), list(acts))
    is_last = obs['is_last']
    if is_last.any():                        # driver.py:72-74
      acts = self._ops_or_default().mask_actions(acts, is_last)
    self.acts = {**acts, 'reset': is_last.copy()}
    trans = {**obs, **acts, **outs, **logs}
    self._dispatch(trans, n)
    return step + n, episode + int(is_last.sum())

  def _dispatch(self, trans, n, device_trans=None):
    """Callbacks in registration order; the replay placeholder becomes ONE
    add_batch (device values passed through untouched)."""
    generic = [fn for fn in self.callbacks if fn is not self._replay]
    if self._replay is not None and device_trans is None:
      self._replay.add_batch(trans)
    if generic:
      for i in range(n):
        trn = {k: v[i] for k, v in trans.items()}
        for fn in generic:
          fn(trn, i, **self.kwargs)
    for fn in self.batch_callbacks:
      fn(trans, n)

  # ----------------------------------------------- fused device step (path D)
  def _step_fused(self, policy, per_env, step, episode):
    """Device agent + device replay: obs go host->HBM once and are appended
    and normalised by one launch; actions are masked and appended (together
    with the policy's latent outputs) by a second launch."""
    replay, n = self._replay, self.length
    batch = replay.open_batch(n)
    views = batch.views
    obs_list = self._envs.step(per_env)
    logs = {}
    for i, o in enumerate(obs_list):         # the np.stack of driver.py:65
      for k, v in o.items():
        if k.startswith('log/'):
          logs.setdefault(k, [None] * n)[i] = v
        else:
          views[k][i] = v
    logs = {k: np.stack(v) for k, v in logs.items()}
    obs_keys = [k for k in obs_list[0] if not k.startswith('log/')]
    # the host half of the reference agent's input checks (embodied/jax/agent.py:223-227): the
    # staged block is still host memory here, and only floating-point keys can be non-finite
    for k in obs_keys:
      if views[k].dtype.kind == 'f':
        assert np.isfinite(views[k][:n]).all(), (k, 'non-finite observation')
    obs_dev = replay.stage_obs(batch, obs_keys)
    self.carry, acts, outs = policy(self.carry, obs_dev, **self.kwargs)
    assert all(k not in acts for k in outs), (list(outs), list(acts))
    host_acts = replay.commit_batch(batch, acts, outs)
    for k, v in host_acts.items():           # jax/agent.py:254-259 on the masked host copy
      if v.dtype.kind in 'iu':
        assert (v >= 0).all(), (k, 'negative discrete action')
      elif v.dtype.kind == 'f':
        assert np.isfinite(v).all(), (k, 'non-finite action')
    is_last = views['is_last'][:n].copy()
    self.acts = {**host_acts, 'reset': is_last}
    generic = [fn for fn in self.callbacks if fn is not replay]
    if generic or self.batch_callbacks:
      trans = {k: views[k][:n].copy() for k in obs_keys}
      trans.update(host_acts)
      if self._fetch_outs:
        trans.update({k: v.cpu().numpy() for k, v in outs.items()})
      trans.update(logs)
      self._dispatch(trans, n, device_trans=True)
    return step + n, episode + int(is_last.sum())
