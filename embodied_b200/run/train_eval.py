"""Actor + learner loop with periodic evaluation episodes (embodied/run/train_eval.py:10-157).

``train_eval(make_agent, make_replay_train, make_replay_eval, make_env_train, make_env_eval,
make_stream, make_logger, args)``: the training half is ``run.train``'s (same pacing, same
batched callbacks, same metric names); every ``report_every`` seconds a second Driver plays
``eval_eps`` episodes with ``mode='eval'`` into its own replay, and ``agent.report`` runs on a
batch of training windows (prefix ``report``) and on a batch of evaluation windows (prefix
``eval``).  ``args`` additionally holds ``eval_envs`` and ``eval_eps`` (train_eval.py:81,129).
The checkpoint keeps both replays (train_eval.py:115-118).
"""
import pickle

from .. import elements
from ..core import base as baselib
from ..core import clock
from ..core import driver as driverlib
from .train import _EpisodeStats, _MetricFetcher, _with_mode


def train_eval(make_agent, make_replay_train, make_replay_eval, make_env_train, make_env_eval,
               make_stream, make_logger, args):

  agent = make_agent()
  lacking = baselib.implements_agent(agent)
  if lacking:
    raise TypeError(f'{type(agent).__name__} does not implement the Agent protocol '
                    f'(embodied/core/base.py:1-31): missing {lacking}')
  replay_train = make_replay_train()
  replay_eval = make_replay_eval()
  logger = make_logger()

  logdir = elements.Path(args.logdir)
  step = logger.step
  usage = elements.Usage(**args.usage)
  train_agg = elements.Agg()
  train_epstats = elements.Agg()
  eval_epstats = elements.Agg()
  policy_fps = elements.FPS()
  train_fps = elements.FPS()

  batch_steps = args.batch_size * args.batch_length
  should_train = elements.when.Ratio(args.train_ratio / batch_steps)
  should_log = clock.LocalClock(args.log_every)
  should_report = clock.LocalClock(args.report_every)
  should_save = clock.LocalClock(args.save_every)
  ops = getattr(args, 'driver_ops', None)

  def make_driver(make_env, count):
    fns = [(lambda i=i: make_env(i)) for i in range(count)]
    return driverlib.Driver(fns, parallel=not args.debug, fetch_outs=False, ops=ops)

  driver_train = make_driver(make_env_train, args.envs)
  driver_train.on_step(replay_train.add)
  train_stats = _EpisodeStats(logger, train_epstats)

  driver_eval = make_driver(make_env_eval, args.eval_envs)
  driver_eval.on_step(replay_eval.add)
  eval_stats = _EpisodeStats(logger, eval_epstats)

  stream_train = iter(agent.stream(make_stream(replay_train, 'train')))
  stream_report = iter(agent.stream(make_stream(replay_train, 'report')))
  stream_eval = iter(agent.stream(make_stream(replay_eval, 'eval')))
  carry_train = [agent.init_train(args.batch_size)]
  carry_report = agent.init_report(args.batch_size)
  carry_eval = agent.init_report(args.batch_size)
  fetch_metrics = _MetricFetcher()

  def after_train_step(trans, n):
    step.increment(n)
    policy_fps.step(n)
    train_stats(trans, n)
    if len(replay_train) < args.batch_size * args.batch_length:
      return
    for _ in range(should_train(step)):
      with elements.timer.section('stream_next'):
        batch = next(stream_train)
      carry_train[0], outs, mets = agent.train(carry_train[0], batch)
      train_fps.step(batch_steps)
      if 'replay' in outs:
        replay_train.update(outs['replay'])
      train_agg.add(fetch_metrics(mets), prefix='train')
  driver_train.on_batch(after_train_step)

  def after_eval_step(trans, n):          # evaluation steps count for fps/policy only (:85)
    policy_fps.step(n)
    eval_stats(trans, n)
  driver_eval.on_batch(after_eval_step)

  def report(carry, stream):
    agg = elements.Agg()
    for _ in range(args.report_batches):
      carry, mets = agent.report(carry, next(stream))
      agg.add({k: (v.cpu().numpy() if hasattr(v, 'cpu') else v) for k, v in mets.items()})
    return carry, agg.result()

  cp = elements.Checkpoint(logdir / 'checkpoint.pkl')
  cp.step = step
  cp.agent = agent
  cp.replay_train = replay_train
  cp.replay_eval = replay_eval
  if args.from_checkpoint:
    data = pickle.loads(elements.Path(args.from_checkpoint).read(mode='rb'))
    regex = args.get('from_checkpoint_regex', None) if hasattr(args, 'get') else None
    payload = data['model'] if 'model' in data else data['agent']
    agent.load(payload, regex=regex) if regex else agent.load(payload)
  cp.load_or_save()
  should_save(step)          # the checkpoint was just written

  print('Start training loop')
  train_policy = _with_mode(agent, 'train')
  eval_policy = _with_mode(agent, 'eval')
  driver_train.reset(agent.init_policy)
  while step < args.steps:

    if should_report(step):
      print('Evaluation')
      driver_eval.reset(agent.init_policy)
      driver_eval(eval_policy, episodes=args.eval_eps)
      logger.add(eval_epstats.result(), prefix='epstats')
      if len(replay_train):
        carry_report, mets = report(carry_report, stream_report)
        logger.add(mets, prefix='report')
      if len(replay_eval):
        carry_eval, mets = report(carry_eval, stream_eval)
        logger.add(mets, prefix='eval')

    driver_train(train_policy, steps=10)

    if should_log(step):
      logger.add(train_agg.result())
      logger.add(train_epstats.result(), prefix='epstats')
      logger.add(replay_train.stats(), prefix='replay')
      logger.add(usage.stats(), prefix='usage')
      logger.add({'fps/policy': policy_fps.result()})
      logger.add({'fps/train': train_fps.result()})
      logger.add({'timer': elements.timer.stats()['summary']})
      logger.write()

    if should_save(step):
      cp.save()

  driver_train.close()
  driver_eval.close()
  logger.close()
