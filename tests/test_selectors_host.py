"""Non-uniform selectors (embodied/core/selectors.py:60-378) against the
reference: committed draw sequences (tests/golden/selectors.npz, made by oracle/gen_golden.py
from the reference's own classes) and, where /root/reference exists, the reference's classes
driven side by side on fresh seeds.  Plus the invariants of the reference's
tests/test_sampletree.py restated for the flat SumTree."""
import collections
import types

import numpy as np
import pytest

from embodied_b200.core import selectors as S
from oracle import gen_golden, refload
import golden_cases
import selector_cases as sc

needs_ref = pytest.mark.skipif(not refload.available(), reason='/root/reference not on this machine')
GOLD = np.load(golden_cases.GOLDEN / 'selectors.npz')


@pytest.mark.parametrize('name', sorted(sc.TREE_CASES))
def test_sumtree_draws_match_the_reference_golden(name):
  spec = sc.TREE_CASES[name]
  got = sc.drive_tree(S.SumTree(spec['branching'], seed=spec['seed']), spec)
  assert np.array_equal(got, GOLD[f'tree/{name}'])


@pytest.mark.parametrize('name', sorted(sc.PRIO_CASES))
def test_prioritized_draws_match_the_reference_golden(name):
  spec = sc.PRIO_CASES[name]
  got = sc.drive_selector(S.Prioritized(seed=spec['seed'], **spec['kwargs']), spec['seed'])
  assert np.array_equal(got, GOLD[f'prio/{name}'])


def test_recency_table_and_draws_match_the_reference_golden():
  table = S.Recency(sc.recency_uprobs(300, 0.7)).table
  assert len(table) == 3
  for level, probs in enumerate(table):
    assert np.array_equal(probs, GOLD[f'recency/table{level}'])
  got = sc.drive_selector(S.Recency(sc.recency_uprobs(), seed=9), 9)
  assert np.array_equal(got, GOLD['recency/draws'])


def test_mixture_draws_match_the_reference_golden():
  assert np.array_equal(sc.drive_selector(sc.make_mixture(S), 21), GOLD['mixture/draws'])


@needs_ref
@pytest.mark.parametrize('seed', [31, 32, 33])
def test_live_reference_side_by_side(seed):
  ns = refload.load()
  spec = dict(branching=int(2 + seed % 7), n=50, ops=600, seed=seed, special=seed % 2 == 0)
  assert np.array_equal(sc.drive_tree(S.SumTree(spec['branching'], seed=seed), spec),
                        sc.drive_tree(ns.selectors.SampleTree(spec['branching'], seed=seed), spec))
  kw = dict(exponent=0.7, maxfrac=0.3, branching=5, zero_on_sample=seed % 2 == 1)
  assert np.array_equal(sc.drive_selector(S.Prioritized(seed=seed, **kw), seed),
                        sc.drive_selector(ns.selectors.Prioritized(seed=seed, **kw), seed))
  ref = types.SimpleNamespace(Uniform=ns.selectors.Uniform, Prioritized=ns.selectors.Prioritized,
                              Recency=gen_golden.fixed_recency(ns), Mixture=ns.selectors.Mixture)
  assert np.array_equal(sc.drive_selector(sc.make_mixture(S, seed), seed),
                        sc.drive_selector(sc.make_mixture(ref, seed), seed))


@needs_ref
def test_reference_recency_cannot_draw_as_shipped():
  """Documents why Recency is pinned against a one-name fix of the reference (selectors.py:109)."""
  ns = refload.load()
  sel = ns.selectors.Recency(sc.recency_uprobs())
  sel[0] = [sc.stepid(0)]
  with pytest.raises(UnboundLocalError):
    sel()


# ---- invariants of the reference's tests/test_sampletree.py, on the flat tree

@pytest.mark.parametrize('branching', [2, 3, 5, 10])
def test_total_is_the_sum_of_weights(branching):
  tree = S.SumTree(branching)
  for index in range(50):
    assert tree.total == sum(range(index))
    tree.insert(index, index)


@pytest.mark.parametrize('inserts', [1, 2, 10, 100])
@pytest.mark.parametrize('remove_every', [0, 2, 3, 4])
@pytest.mark.parametrize('branching', [2, 3, 5, 10])
def test_depth_grows_with_inserts_and_survives_removals(inserts, remove_every, branching):
  tree = S.SumTree(branching)
  for index in range(inserts):
    tree.insert(index, 1)
  removed = list(range(0, inserts, remove_every)) if remove_every and inserts > 1 else []
  for index in removed:
    tree.remove(index)
  assert len(tree) == inserts - len(removed)
  if len(tree):
    assert tree.depth == max(1, int(np.ceil(np.log(inserts) / np.log(branching))))
    assert all(len(b) == branching for b in tree.buckets[:-1]) and tree.buckets[-1]


@pytest.mark.parametrize('branching', [2, 3, 5, 10])
def test_empty_tree_restarts(branching):
  tree = S.SumTree(branching)
  rng = np.random.default_rng(0)
  for key in rng.permutation(100):
    tree.insert(int(key), 1)
  depth = tree.depth
  for key in rng.permutation(100):
    tree.remove(int(key))
  assert len(tree) == 0 and tree.depth == 1 and tree.total == 0
  for key in rng.permutation(100):
    tree.insert(int(key), 1)
  assert tree.depth == depth and tree.total == 100


@pytest.mark.parametrize('branching', [2, 3, 5, 10])
def test_single_survivor_is_always_drawn(branching):
  tree = S.SumTree(branching)
  for key in (12, 123, 42):
    tree.insert(key, 1.0)
  tree.remove(12)
  tree.remove(42)
  assert all(tree.sample() == 123 for _ in range(10))


@pytest.mark.parametrize('uprob', [1e-5, 1.0, 1e5])
@pytest.mark.parametrize('branching', [2, 3, 5, 10])
def test_equal_weights_draw_uniformly(branching, uprob):
  tree = S.SumTree(branching, seed=0)
  keys = list(range(10))
  for key in keys:
    tree.insert(key, uprob)
  for key in keys[::3]:
    tree.remove(key)
  left = [k for k in keys if k % 3]
  counts = collections.Counter(tree.sample() for _ in range(3000))
  assert set(counts) == set(left)
  assert max(abs(c / 3000 - 1 / len(left)) for c in counts.values()) < 0.04


def test_weights_shape_the_draw_frequencies():
  tree = S.SumTree(4, seed=0)
  weights = {0: 0.0, 1: 1.0, 2: 3.0, 3: 0.0, 4: 6.0}
  for key, w in weights.items():
    tree.insert(key, w)
  counts = collections.Counter(tree.sample() for _ in range(5000))
  assert counts[0] == counts[3] == 0
  for key in (1, 2, 4):
    assert abs(counts[key] / 5000 - weights[key] / 10) < 0.03
  tree.update(4, 0.0)
  counts = collections.Counter(tree.sample() for _ in range(2000))
  assert counts[4] == 0 and abs(counts[2] / 2000 - 0.75) < 0.04
  tree.update(1, float('inf'))                       # infinite weights win outright
  assert all(tree.sample() == 1 for _ in range(20))


def test_uniform_bulk_draw_is_the_scalar_sequence():
  """`Uniform.draw(n)` (one numpy call) must produce exactly the keys n scalar draws would, and
  leave the generator where they would: Replay.sample's bulk path relies on it for the bit-exact
  sampling contract (embodied/core/selectors.py:39-43)."""
  import numpy as np
  from embodied_b200.core import selectors
  for n in (1, 2, 3, 10, 1000, 70001):
    for seed in (0, 1, 7):
      a, b = selectors.Uniform(seed), selectors.Uniform(seed)
      for k in range(n):
        a[k] = None
        b[k] = None
      scalar = [a() for _ in range(133)]
      assert b.draw(100) + b.draw(1) + b.draw(32) == scalar
      if n >= 2:                                 # the reference refuses to delete the last key (:51)
        del a[n - 1], b[n - 1]
      a[-5] = b[-5] = None
      assert b.draw(17) == [a() for _ in range(17)]
  for n in (2 ** 31 - 1, 2 ** 32 - 1, 2 ** 32, 2 ** 32 + 5, 2 ** 40):     # both Lemire widths
    x, y = np.random.default_rng(3), np.random.default_rng(3)
    assert [int(x.integers(0, n)) for _ in range(65)] == y.integers(0, n, size=65).tolist()
    assert int(x.integers(0, n)) == int(y.integers(0, n))
