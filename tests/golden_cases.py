"""Replays the seeded streams behind tests/golden/*.npz into any Replay-like
object (oracle, product on a numpy double, product on the GPU) and compares with
what the REAL reference produced (oracle/gen_golden.py)."""
import pathlib

import numpy as np

from oracle import gen_golden

GOLDEN = pathlib.Path(__file__).parent / 'golden'


def tonp(x):
  return x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)


def check_batch(got, want, prefix, fix):
  keys = sorted(k[len(prefix):] for k in fix.files if k.startswith(prefix))
  assert sorted(got.keys()) == keys, (sorted(got.keys()), keys)
  for k in keys:
    a, b = tonp(got[k]), fix[prefix + k]
    assert a.dtype == b.dtype, (k, a.dtype, b.dtype)
    assert a.shape == b.shape, (k, a.shape, b.shape)
    assert a.tobytes() == b.tobytes(), f'{prefix}{k} differs from the reference'


def run_replay_case(name, make_replay, adder=None):
  """make_replay(length, capacity, chunksize, online, seed) -> replay."""
  fix = np.load(GOLDEN / f'{name}.npz')
  length, capacity, chunksize, online, workers, steps, batch, nsamples = (
      int(x) for x in fix['spec'])
  replay = make_replay(length, capacity, chunksize, bool(online), 0)
  rng = np.random.default_rng(1234)
  n, lens = 0, []
  for t in range(steps):
    if adder is None:
      for w in range(workers):
        replay.add(gen_golden.transition(rng, w, t, gen_golden.SHAPES), w)
        lens.append(len(replay))
    else:
      rows = [gen_golden.transition(rng, w, t, gen_golden.SHAPES)
              for w in range(workers)]
      adder(replay, rows)
      lens.extend([None] * (workers - 1) + [len(replay)])
    if len(replay) and t % 5 == 4 and n < nsamples:
      data = replay.sample(batch)
      check_batch(data, None, f'sample{n}/', fix)
      deter = data['deter']
      replay.update({'stepid': data['stepid'], 'deter': deter + np.float32(n + 1)})
      n += 1
  want = fix['lens']
  for i, v in enumerate(lens):
    if v is not None:
      assert v == want[i], (i, v, want[i])
  assert n == int(fix['nsamples_done'])
  check_batch(replay.sample(batch, mode='report'), None, 'final/', fix)
  return replay
