"""Single-process actor + learner loop.

Same factories, arguments, pacing and metric names as the reference
``embodied.run.train`` (embodied/run/train.py:10-118):
``train(make_agent, make_replay, make_env, make_stream, make_logger, args)``
with ``args`` holding ``steps, envs, debug, batch_size, batch_length,
train_ratio, log_every, report_every, save_every, consec_report,
report_batches, from_checkpoint, usage, logdir``.

Loop: ``driver(policy, steps=10)``; each Driver step appends N transitions to
the replay and runs ``when.Ratio(train_ratio / (B*T))`` train steps
(train.py:25-26,69-80), each followed by ``replay.update(outs['replay'])``
(train.py:77-78).  ``fps/policy`` and ``fps/train`` (train.py:110-111) are the
two numbers BASELINE.json's metric asks for.

Differences, all on the host side: the N per-transition callbacks of one
Driver step are served batched (one replay append launch, vectorised episode
statistics), so pacing is evaluated once per Driver step with the same
cumulative counts.
"""
import collections
import pickle

import numpy as np

from .. import elements
from ..core import base as baselib
from ..core import clock
from ..core import driver as driverlib


class _EpisodeStats:
  """Vectorised logfn (train.py:31-54): score / length / reward_rate per env; every scalar
  `log/<k>` key is aggregated over the episode as `log/<k>/avg|max|sum` (elements.Agg with
  agg=('avg', 'max', 'sum'), train.py:44-46)."""

  def __init__(self, logger, epstats):
    self.logger, self.epstats = logger, epstats
    self.score = self.length = self.changes = self.prev = None
    self.logsum, self.logmax = {}, {}

  def __call__(self, trans, n):
    reward = np.asarray(trans['reward'], np.float64)
    first = np.asarray(trans['is_first'], bool)
    last = np.asarray(trans['is_last'], bool)
    if self.score is None:
      self.score = np.zeros(n)
      self.length = np.zeros(n, np.int64)
      self.changes = np.zeros(n, np.int64)
      self.prev = np.zeros(n)
    self.score[first] = 0
    self.length[first] = 0
    self.changes[first] = 0
    moved = (np.abs(reward - self.prev) >= 0.01) & ~first
    self.changes += moved
    self.score += reward
    self.length += 1
    self.prev = reward
    logkeys = [k for k in trans if k.startswith('log/')]
    for k in logkeys:
      value = np.asarray(trans[k], np.float64)
      assert value.shape == (n,), (k, value.shape)       # scalars per env (train.py:45)
      if k not in self.logsum:
        self.logsum[k] = np.zeros(n)
        self.logmax[k] = np.full(n, -np.inf)
      self.logsum[k][first] = 0
      self.logmax[k][first] = -np.inf
      self.logsum[k] += value
      self.logmax[k] = np.maximum(self.logmax[k], value)
    for i in np.nonzero(last)[0]:
      self.logger.add(
          {'score': self.score[i], 'length': self.length[i]}, prefix='episode')
      result = {}
      if self.length[i] > 1:
        result['reward_rate'] = self.changes[i] / (self.length[i] - 1)
      for k in logkeys:
        result[f'{k}/avg'] = self.logsum[k][i] / self.length[i]
        result[f'{k}/max'] = self.logmax[k][i]
        result[f'{k}/sum'] = self.logsum[k][i]
      self.epstats.add(result)


class _MetricFetcher:
  """Device scalars of the learner -> host numbers, ONE STEP DELAYED like the reference's
  `pending_mets` (embodied/jax/agent.py:291-298): this call's metrics start an asynchronous copy
  into pinned memory, the previous call's (long complete) are returned -- the train loop never
  waits for the update it just enqueued."""

  def __init__(self):
    self.pending = None
    self.ring, self.turn = [None, None], 0

  def __call__(self, mets):
    try:
      import torch
    except ImportError:                                   # host-only agents
      return mets
    dev = {k: v for k, v in mets.items()
           if isinstance(v, torch.Tensor) and v.is_cuda and v.numel() == 1}
    out = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v)
           for k, v in mets.items() if k not in dev}
    prev, self.pending = self.pending, None
    if dev:
      names = list(dev)
      vec = torch.stack([dev[k].detach().to(torch.float32).reshape(()) for k in names])
      buf = self.ring[self.turn]
      if buf is None or buf.numel() < len(names):
        buf = self.ring[self.turn] = torch.empty(max(64, len(names)), dtype=torch.float32,
                                                 pin_memory=True)
      self.turn ^= 1
      buf[:len(names)].copy_(vec, non_blocking=True)
      event = torch.cuda.Event()
      event.record()
      self.pending = (names, buf, event)
    if prev:
      names, buf, event = prev
      event.synchronize()
      host = buf[:len(names)].numpy().copy()
      out.update({k: host[i] for i, k in enumerate(names)})
    return out


def train(make_agent, make_replay, make_env, make_stream, make_logger, args):

  agent = make_agent()
  lacking = baselib.implements_agent(agent)
  if lacking:
    raise TypeError(f'{type(agent).__name__} does not implement the Agent protocol '
                    f'(embodied/core/base.py:1-31): missing {lacking}')
  replay = make_replay()
  logger = make_logger()

  logdir = elements.Path(args.logdir)
  step = logger.step
  usage = elements.Usage(**args.usage)
  train_agg = elements.Agg()
  epstats = elements.Agg()
  policy_fps = elements.FPS()
  train_fps = elements.FPS()

  batch_steps = args.batch_size * args.batch_length
  should_train = elements.when.Ratio(args.train_ratio / batch_steps)
  should_log = clock.LocalClock(args.log_every)
  should_report = clock.LocalClock(args.report_every)
  should_save = clock.LocalClock(args.save_every)

  fns = [(lambda i=i: make_env(i)) for i in range(args.envs)]
  driver = driverlib.Driver(
      fns, parallel=not args.debug, fetch_outs=False,
      ops=getattr(args, 'driver_ops', None))
  episode_stats = _EpisodeStats(logger, epstats)
  driver.on_step(replay.add)

  stream_train = iter(agent.stream(make_stream(replay, 'train')))
  stream_report = iter(agent.stream(make_stream(replay, 'report')))
  carry_train = [agent.init_train(args.batch_size)]
  carry_report = agent.init_report(args.batch_size)

  def after_step(trans, n):
    step.increment(n)
    policy_fps.step(n)
    episode_stats(trans, n)
    if len(replay) < args.batch_size * args.batch_length:
      return
    for _ in range(should_train(step)):
      with elements.timer.section('stream_next'):
        batch = next(stream_train)
      carry_train[0], outs, mets = agent.train(carry_train[0], batch)
      train_fps.step(batch_steps)
      if 'replay' in outs:
        replay.update(outs['replay'])
      train_agg.add(fetch_metrics(mets), prefix='train')
  fetch_metrics = _MetricFetcher()
  driver.on_batch(after_step)

  cp = elements.Checkpoint(logdir / 'checkpoint.pkl')
  cp.step = step
  cp.agent = agent
  cp.replay = replay
  if args.from_checkpoint:
    data = pickle.loads(elements.Path(args.from_checkpoint).read(mode='rb'))
    regex = args.get('from_checkpoint_regex', None) if hasattr(args, 'get') else None
    agent.load(data['agent'], regex=regex) if regex else agent.load(data['agent'])
  cp.load_or_save()

  print('Start training loop')
  policy = _with_mode(agent, 'train')
  driver.reset(agent.init_policy)
  while step < args.steps:

    driver(policy, steps=10)

    if should_report(step) and len(replay):
      agg = elements.Agg()
      for _ in range(args.get('consec_report', 1) * args.report_batches):
        carry_report, mets = agent.report(carry_report, next(stream_report))
        agg.add({k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in mets.items()})
      logger.add(agg.result(), prefix='report')

    if should_log(step):
      logger.add(train_agg.result())
      logger.add(epstats.result(), prefix='epstats')
      logger.add(replay.stats(), prefix='replay')
      logger.add(usage.stats(), prefix='usage')
      logger.add({'fps/policy': policy_fps.result()})
      logger.add({'fps/train': train_fps.result()})
      logger.add({'timer': elements.timer.stats()['summary']})
      logger.write()

    if should_save(step):
      cp.save()

  driver.close()
  logger.close()


class _with_mode:
  """policy(carry, obs) -> agent.policy(carry, obs, mode=...), keeping the
  agent visible to the Driver (device_obs / ext_space lookups)."""

  def __init__(self, agent, mode):
    self.agent, self.mode = agent, mode
    self.device_obs = bool(getattr(agent, 'device_obs', False))
    self.ext_space = getattr(agent, 'ext_space', None)

  def __call__(self, carry, obs, **kwargs):
    return self.agent.policy(carry, obs, mode=self.mode, **kwargs)
