"""Seeded operation streams for the non-uniform selectors.  The SAME driver runs the reference's
classes (oracle/gen_golden.py, where /root/reference exists) and the product's
(tests/test_selectors_host.py); what it returns -- every drawn key, in order -- is committed as
tests/golden/selectors.npz."""
import numpy as np

TREE_CASES = {
    'b2': dict(branching=2, n=40, ops=400, seed=1),
    'b3_zero_inf': dict(branching=3, n=30, ops=400, seed=2, special=True),
    'b16': dict(branching=16, n=300, ops=1500, seed=3),
    'b5_drain': dict(branching=5, n=26, ops=300, seed=4, drain=True),
}

PRIO_CASES = {
    'default': dict(kwargs=dict(), seed=5),
    'exp_maxfrac': dict(kwargs=dict(exponent=0.8, maxfrac=0.5, initial=2.0, branching=4), seed=6),
    'zero_on_sample': dict(kwargs=dict(zero_on_sample=True, branching=3), seed=7),
}


def drive_tree(tree, spec):
  """insert / remove / update / sample on a SumTree-like object; returns the sampled keys."""
  rng = np.random.default_rng(spec['seed'])
  alive, out, nextkey = [], [], 0

  def weight():
    if spec.get('special') and rng.random() < 0.15:
      return [0.0, float('inf')][int(rng.integers(0, 2))]
    return float(rng.random() * 3)

  for _ in range(spec['n']):
    tree.insert(nextkey, weight())
    alive.append(nextkey)
    nextkey += 1
  for step in range(spec['ops']):
    op = rng.integers(0, 4)
    if spec.get('drain') and step == spec['ops'] // 2:
      for k in list(alive):           # empty the tree completely, then rebuild
        tree.remove(k)
      alive.clear()
    if op == 0 or not alive:
      tree.insert(nextkey, weight())
      alive.append(nextkey)
      nextkey += 1
    elif op == 1 and len(alive) > 1:
      k = alive.pop(int(rng.integers(0, len(alive))))
      tree.remove(k)
    elif op == 2:
      tree.update(alive[int(rng.integers(0, len(alive)))], weight())
    else:
      out.append(tree.sample())
  return np.array(out, np.int64)


def stepid(n):
  return np.frombuffer(int(n).to_bytes(20, 'big'), np.uint8)


def drive_selector(sel, seed, length=4, nitems=60, ops=500, prioritize=True):
  """The Replay-side protocol: items are windows of `length` consecutive step ids of one
  stream (overlapping, like replay.py:100-107), removed oldest first; priorities arrive as
  float32 arrays the way Replay.update hands them over (replay.py:136-139)."""
  rng = np.random.default_rng(seed)
  out, alive, nextitem = [], [], 0

  def insert():
    nonlocal nextitem
    sel[nextitem] = [stepid(nextitem + i) for i in range(length)]
    alive.append(nextitem)
    nextitem += 1

  for _ in range(nitems):
    insert()
  for _ in range(ops):
    op = rng.integers(0, 4)
    if op == 0:
      insert()
      if len(alive) > nitems:
        del sel[alive.pop(0)]
    elif op == 1 and prioritize and hasattr(sel, 'prioritize'):
      n = int(rng.integers(1, 6))
      base = alive[int(rng.integers(0, len(alive)))]
      ids = np.stack([stepid(base + i) for i in range(n)])
      prios = rng.random(n).astype(np.float32) * 2
      sel.prioritize(ids, prios)
    else:
      out.append(sel())
  return np.array(out, np.int64)


def recency_uprobs(n=70, exp=1.0):
  return 1.0 / np.arange(1, n + 1) ** exp


def make_mixture(S, seed=3):
  """S = a module with Uniform / Prioritized / Recency / Mixture."""
  return S.Mixture(dict(
      uniform=S.Uniform(seed=11), priority=S.Prioritized(branching=4, seed=12),
      recency=S.Recency(recency_uprobs(), seed=13),
  ), dict(uniform=0.5, priority=0.25, recency=0.25), seed=seed)
