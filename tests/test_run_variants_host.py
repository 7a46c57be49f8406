"""The callers either side of the hot path that SURVEY 8f ranks 4th: `run.train_eval`
(embodied/run/train_eval.py:10-157), the `SamplesPerInsert` rate limiter
(embodied/core/limiters.py:19-80) and a Replay driven by a priority selector
(embodied/core/replay.py:136-139,171-179).  Numpy doubles, no GPU."""
import time

import numpy as np
import pytest

import embodied_b200 as embodied
from embodied_b200 import elements
from embodied_b200.core import limiters, selectors
from embodied_b200.envs import dummy
from oracle import refload
import doubles
from test_train_host import CountAgent

needs_ref = pytest.mark.skipif(not refload.available(), reason='/root/reference not on this machine')


def test_train_eval_loop(tmpdir):
  """Counters of the reference's tests/test_train.py:12-33 on the loop with evaluation: the
  agent sees every TRAIN env step exactly once in order (count continuity), evaluation
  episodes are played with mode='eval' into their own replay, report runs on both replays."""
  args = elements.Config(
      steps=600, train_ratio=32.0, log_every=0.1, report_every=0.2, save_every=0.2,
      report_batches=1, from_checkpoint='', usage=dict(psutil=True), debug=True,
      logdir=str(tmpdir), envs=4, eval_envs=2, eval_eps=2, batch_size=8, batch_length=16,
      replay_context=0, report_length=8)
  make_env = lambda index: dummy.Dummy('disc', size=(64, 64), length=20)
  env = make_env(0)
  agent = CountAgent(env.obs_space, env.act_space)
  modes = []
  inner = agent.policy

  def policy(carry, obs, mode='train'):
    time.sleep(0.004)
    modes.append((mode, len(obs['is_first'])))
    return inner(carry, obs, mode)
  agent.policy = policy
  replays = []

  def make_replay():
    replays.append(embodied.Replay(
        length=args.batch_length, capacity=1e4, store=doubles.HostStore(1024, staging_rows=16)))
    return replays[-1]

  def make_stream(replay, mode):
    fn = embodied.streams.Stateless(replay.sample, args.batch_size, mode)
    return embodied.streams.Consec(
        fn, length=args.batch_length, consec=1, prefix=0, strict=True, contiguous=True)

  make_logger = lambda: elements.Logger(elements.Counter(), [elements.logger.TerminalOutput()])
  args = args.update(driver_ops=doubles.HostOps())
  embodied.run.train_eval(
      lambda: agent, make_replay, make_replay, make_env, make_env, make_stream, make_logger, args)
  stats = agent.stats()
  train_steps = sum(n for m, n in modes if m == 'train')
  eval_steps = sum(n for m, n in modes if m == 'eval')
  assert np.allclose(train_steps, args.steps, 100, 0.1)
  assert eval_steps >= 2 * 20                      # at least one evaluation round of 2 episodes
  assert all(n == 4 for m, n in modes if m == 'train') and all(n == 2 for m, n in modes if m == 'eval')
  assert np.allclose(stats['replay_steps'], train_steps * args.train_ratio, 100, 0.15)
  assert stats['reports'] >= 2                     # train windows and eval windows
  assert stats['saves'] >= 2
  train_replay, eval_replay = replays
  assert len(eval_replay) > 0 and len(train_replay) > len(eval_replay)
  data = __import__('pickle').loads(elements.Path(str(tmpdir)).__truediv__('checkpoint.pkl').read(mode='rb'))
  assert {'step', 'agent', 'replay_train', 'replay_eval'} <= set(data)


def drive_limiter(lim, seed, ops=3000):
  rng = np.random.default_rng(seed)
  trace = []
  for _ in range(ops):
    if rng.random() < 0.45:
      ok = lim.want_insert()
      trace.append(ok)
      if ok:
        lim.insert()
    else:
      ok = lim.want_sample()
      trace.append(ok)
      if ok:
        lim.sample()
  return trace, lim.save()


@needs_ref
@pytest.mark.parametrize('spi,tol,minsize', [(4.0, 8, 10), (0.25, 3, 1), (0, 5, 7), (-1, 5, 2), (1.0, 1, 30)])
def test_samples_per_insert_matches_the_reference(spi, tol, minsize):
  ns = refload.load()
  got = drive_limiter(limiters.SamplesPerInsert(spi, tol, minsize), seed=int(minsize))
  want = drive_limiter(ns.limiters.SamplesPerInsert(spi, tol, minsize), seed=int(minsize))
  assert got == want
  restored = limiters.SamplesPerInsert(spi, tol, minsize)
  restored.load(got[1])
  assert restored.save() == want[1]


def test_samples_per_insert_holds_the_ratio():
  """Once past minsize the admitted samples stay within `tolerance` inserts of
  samples_per_insert x inserts, whichever side pushes (limiters.py:26-28,46-63)."""
  lim = limiters.SamplesPerInsert(4.0, 8, 10)
  assert not lim.want_sample()
  inserts = samples = 0
  rng = np.random.default_rng(0)
  for _ in range(5000):
    if rng.random() < 0.5:
      if lim.want_insert():
        lim.insert(); inserts += 1
    elif lim.want_sample():
      lim.sample(); samples += 1
    if inserts >= 10:
      balance = 4.0 * (inserts - 9) - 10 - samples
      assert -8 - 1 <= balance <= 8 * 4.0 + 4.0, balance
  assert inserts > 500 and samples > 2000


def stream_of(rng, steps, workers):
  for t in range(steps):
    for w in range(workers):
      yield {'x': rng.normal(size=(3,)).astype(np.float32), 'is_first': np.bool_(t % 17 == 0),
             'is_last': np.bool_(t % 17 == 16), 'is_terminal': np.bool_(False)}, w


@needs_ref
def test_replay_with_priority_mixture_samples_what_the_reference_samples():
  """Replay hands the window's step ids to a priority selector on insert and forwards
  `priority` from update() (replay.py:136-139,171-179): same stream, same selectors, same
  priorities -> the product and the reference's own Replay + selectors draw the same windows."""
  ns = refload.load()
  def mixture(S, recency_cls):
    return S.Mixture(dict(
        uniform=S.Uniform(seed=1),
        priority=S.Prioritized(exponent=0.8, maxfrac=0.5, initial=float('inf'), zero_on_sample=True, seed=2),
        recency=recency_cls(1.0 / np.arange(1, 201) ** 1.0, seed=3),
    ), dict(uniform=0.4, priority=0.4, recency=0.2), seed=4)

  from oracle import gen_golden
  import types
  class RefMixture(ns.selectors.Mixture):
    # the reference's Mixture has no __len__, which its own Replay.sample asks for
    # (replay.py:123): given one here so that the reference can run at all
    def __len__(self):
      return len(self.selectors[0])
    def __bool__(self):               # `selector or Uniform(seed)` (replay.py:25) must keep it
      return True
  refS = types.SimpleNamespace(Uniform=ns.selectors.Uniform, Prioritized=ns.selectors.Prioritized,
                               Mixture=RefMixture)
  ns.elements.UUID.reset(debug=True)
  ref = ns.replay.Replay(length=5, capacity=200, chunksize=16,
                         selector=mixture(refS, gen_golden.fixed_recency(ns)))
  elements.UUID.reset(debug=True)
  own = embodied.Replay(length=5, capacity=200, chunksize=16, selector=mixture(selectors, selectors.Recency),
                        store=doubles.HostStore(16, staging_rows=4))
  rng_a, rng_b = np.random.default_rng(5), np.random.default_rng(5)
  prng = np.random.default_rng(6)
  for (step_a, w), (step_b, _) in zip(stream_of(rng_a, 90, 3), stream_of(rng_b, 90, 3)):
    ref.add(step_a, w)
    own.add(step_b, w)
    if len(ref) >= 8 and prng.random() < 0.3:
      a, b = ref.sample(4), own.sample(4)
      b = {k: np.asarray(v) for k, v in b.items()}
      assert np.array_equal(a['stepid'], b['stepid'])
      assert np.array_equal(a['x'], b['x'])
      prio = prng.random((4, 5)).astype(np.float32)
      ref.update({'stepid': a['stepid'], 'priority': prio})
      own.update({'stepid': b['stepid'], 'priority': prio})
  ns.elements.UUID.reset(debug=False)
  elements.UUID.reset(debug=False)
  assert len(ref) == len(own) > 100
