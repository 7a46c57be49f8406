"""Generates tests/golden/*.npz by running the REAL reference (TEST INFRA).

Run in the build container, where /root/reference exists:
    python -m oracle.gen_golden
Every fixture stores the seeds/parameters that define its input stream next to
the reference's outputs, so tests can rebuild the inputs without the reference.
"""
import pathlib

import numpy as np

from . import refload

OUT = pathlib.Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def transition(rng, worker, t, shapes):
  """The seeded synthetic transition stream shared by generator and tests."""
  step = {}
  for key, (dtype, shape) in shapes.items():
    dtype = np.dtype(dtype)
    if dtype == bool:
      step[key] = np.asarray(rng.integers(0, 8, shape) == 0)
    elif np.issubdtype(dtype, np.integer):
      step[key] = rng.integers(0, 200, shape).astype(dtype)
    else:
      step[key] = rng.standard_normal(shape).astype(dtype)
  step['worker'] = np.int32(worker)
  step['t'] = np.int32(t)
  return step


SHAPES = {
    'image': ('uint8', (8, 8, 3)),
    'vector': ('float32', (7,)),
    'reward': ('float32', ()),
    'is_first': ('bool', ()),
    'is_last': ('bool', ()),
    'is_terminal': ('bool', ()),
    'deter': ('float32', (32,)),
}

REPLAY_CASES = {
    # name: (length, capacity, chunksize, online, workers, steps, batch, nsamples)
    'replay_basic': (5, 40, 1024, False, 3, 30, 4, 3),
    'replay_crosschunk': (7, 27, 4, False, 2, 30, 4, 3),
    'replay_evict_online': (9, 20, 8, True, 4, 40, 6, 4),
    'replay_len1': (1, 5, 2, False, 1, 12, 3, 2),
}


def gen_replay(ns, name, spec):
  length, capacity, chunksize, online, workers, steps, batch, nsamples = spec
  ns.elements.UUID.reset(debug=True)
  replay = ns.replay.Replay(
      length=length, capacity=capacity, chunksize=chunksize, online=online,
      seed=0)
  rng = np.random.default_rng(1234)
  out = {'spec': np.array(
      [length, capacity, chunksize, int(online), workers, steps, batch,
       nsamples], np.int64)}
  n = 0
  lens = []
  for t in range(steps):
    for w in range(workers):
      replay.add(transition(rng, w, t, SHAPES), w)
      lens.append(len(replay))
    # sample between driver steps once enough items exist
    if len(replay) and t % 5 == 4 and n < nsamples:
      batchdata = replay.sample(batch)
      for k, v in batchdata.items():
        out[f'sample{n}/{k}'] = v
      # latent write-back, then read it again through a second sample later
      upd = {
          'stepid': batchdata['stepid'],
          'deter': (batchdata['deter'] + np.float32(n + 1)).astype(np.float32)}
      replay.update(upd)
      n += 1
  out['lens'] = np.array(lens, np.int64)
  final = replay.sample(batch, mode='report')
  for k, v in final.items():
    out[f'final/{k}'] = v
  out['nsamples_done'] = np.int64(n)
  np.savez_compressed(OUT / f'{name}.npz', **out)
  ns.elements.UUID.reset(debug=False)


def gen_uniform(ns):
  sel = ns.selectors.Uniform(seed=0)
  rng = np.random.default_rng(7)
  draws, live, nxt = [], [], 0
  ops = []
  for _ in range(400):
    r = rng.integers(0, 10)
    if r < 5 or len(live) < 3:
      sel[nxt] = None
      live.append(nxt)
      ops.append((0, nxt))
      nxt += 1
    elif r < 7:
      victim = live.pop(int(rng.integers(0, len(live))))
      del sel[victim]
      ops.append((1, victim))
    else:
      got = sel()
      ops.append((2, got))
      draws.append(got)
  np.savez_compressed(
      OUT / 'uniform_ops.npz', ops=np.array(ops, np.int64),
      draws=np.array(draws, np.int64))


def gen_driver(ns):
  """Flag / action sequences of the reference Driver on its Dummy env with a
  deterministic policy (cf. reference tests/test_driver.py:44-58)."""
  envs = [ns.dummy.Dummy('disc', size=(8, 8), length=3 + i) for i in range(3)]
  driver = ns.driver.Driver([(lambda e=e: e) for e in envs], parallel=False)
  rows = []
  driver.on_step(lambda tran, worker: rows.append(
      (worker, {k: np.asarray(v) for k, v in tran.items()})))
  counter = [0]

  def policy(carry, obs):
    n = len(obs['is_first'])
    counter[0] += 1
    act = {
        'act_disc': np.full(n, counter[0] % 5, np.int32) + np.arange(n, dtype=np.int32) % 2,
        'act_cont': (np.arange(n * 6, dtype=np.float32).reshape(n, 6)
                     - 7.5 + counter[0]).astype(np.float32)}
    return carry, act, {'aux': np.full((n, 2), -float(counter[0]), np.float32)}

  driver.reset()
  driver(policy, steps=45)
  keys = sorted(rows[0][1])
  out = {'workers': np.array([w for w, _ in rows], np.int64)}
  for k in keys:
    if k == 'image':
      continue
    out[f'tran/{k}'] = np.stack([r[k] for _, r in rows])
  np.savez_compressed(OUT / 'driver_seq.npz', **out)


def gen_consec(ns):
  ns.elements.UUID.reset(debug=True)
  replay = ns.replay.Replay(length=2 * 3 + 1, capacity=64, chunksize=8, seed=3)
  rng = np.random.default_rng(99)
  for t in range(40):
    for w in range(2):
      replay.add(transition(rng, w, t, SHAPES), w)
  source = ns.streams.Stateless(replay.sample, 3, 'train')
  stream = iter(ns.streams.Consec(
      source, length=3, consec=2, prefix=1, strict=True, contiguous=True))
  out = {}
  for i in range(4):
    batch = next(stream)
    for k, v in batch.items():
      out[f'b{i}/{k}'] = v
  np.savez_compressed(OUT / 'consec.npz', **out)
  ns.elements.UUID.reset(debug=False)


CHUNKDIR_SPEC = dict(length=6, capacity=60, chunksize=8, workers=3, steps=26, batch=5, seed=2)


def fill_chunkdir_stream(replay, spec, save_at=(11,)):
  """The seeded stream behind tests/golden/ref_chunks: `save()` is also called mid-stream, so the
  directory holds chunks that were completed early (length < chunksize, successor set)."""
  rng = np.random.default_rng(4321)
  for t in range(spec['steps']):
    for w in range(spec['workers']):
      replay.add(transition(rng, w, t, SHAPES), w)
    if t in save_at:
      replay.save()
  replay.save()


def gen_chunkdir(ns):
  """A replay directory written by the REFERENCE's own Replay.save (chunk.py:64-74) plus what
  the reference samples after loading it back into a fresh buffer (replay.py:312-359)."""
  import shutil
  out = OUT / 'ref_chunks'
  shutil.rmtree(out, ignore_errors=True)
  sp = CHUNKDIR_SPEC
  ns.elements.UUID.reset(debug=True)
  replay = ns.replay.Replay(length=sp['length'], capacity=sp['capacity'], chunksize=sp['chunksize'],
                            directory=str(out), seed=sp['seed'], save_wait=True)
  fill_chunkdir_stream(replay, sp)
  fresh = ns.replay.Replay(length=sp['length'], capacity=sp['capacity'], chunksize=sp['chunksize'],
                           directory=str(out), seed=sp['seed'])
  fresh.load()
  res = {'len': np.int64(len(fresh))}
  for i in range(3):
    for k, v in fresh.sample(sp['batch']).items():
      res[f'sample{i}/{k}'] = v
  np.savez_compressed(OUT / 'ref_chunks_expected.npz', **res)
  ns.elements.UUID.reset(debug=False)


def fixed_recency(ns):
  """The reference's Recency with the unbound `segment` of selectors.py:109 replaced by the
  intended `p` (as shipped the class raises UnboundLocalError on its first draw; asserted in
  tests/test_selectors_host.py).  Everything else -- table, bookkeeping, rescaling -- is the
  reference's own code."""
  class Recency(ns.selectors.Recency):
    def _sample(self, tree, rng, bfactor=16):
      path = []
      for level, prob in enumerate(tree):
        p = prob
        for segment in path:
          p = p[segment]
        path.append(rng.choice(len(p), p=p))
      return sum(index * bfactor ** (len(tree) - level - 1) for level, index in enumerate(path))
  return Recency


def gen_selectors(ns):
  """Draw sequences of the reference's SampleTree / Prioritized / Recency / Mixture on the seeded
  operation streams of tests/selector_cases.py."""
  import sys, types
  sys.path.insert(0, str(OUT.parent))
  import selector_cases as sc
  out = {}
  for name, spec in sc.TREE_CASES.items():
    out[f'tree/{name}'] = sc.drive_tree(ns.selectors.SampleTree(spec['branching'], seed=spec['seed']), spec)
  for name, spec in sc.PRIO_CASES.items():
    out[f'prio/{name}'] = sc.drive_selector(ns.selectors.Prioritized(seed=spec['seed'], **spec['kwargs']), spec['seed'])
  Recency = fixed_recency(ns)
  out['recency/draws'] = sc.drive_selector(Recency(sc.recency_uprobs(), seed=9), 9)
  for level, table in enumerate(ns.selectors.Recency(sc.recency_uprobs(300, 0.7)).tree):
    out[f'recency/table{level}'] = table
  mod = types.SimpleNamespace(Uniform=ns.selectors.Uniform, Prioritized=ns.selectors.Prioritized,
                              Recency=Recency, Mixture=ns.selectors.Mixture)
  out['mixture/draws'] = sc.drive_selector(sc.make_mixture(mod), 21)
  np.savez_compressed(OUT / 'selectors.npz', **out)


def main():
  OUT.mkdir(parents=True, exist_ok=True)
  ns = refload.load()
  for name, spec in REPLAY_CASES.items():
    gen_replay(ns, name, spec)
  gen_uniform(ns)
  gen_driver(ns)
  gen_consec(ns)
  gen_chunkdir(ns)
  gen_selectors(ns)
  for p in sorted(OUT.glob('*.npz')):
    print(p.name, p.stat().st_size)


if __name__ == '__main__':
  main()
