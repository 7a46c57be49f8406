"""embodied_b200.Replay on the GPU against the reference's golden outputs and
the pinned oracle: byte-exact batches, identical index sampling, latent
write-back, chunk save/load, and size-independent properties at the full
BASELINE config-2 size (B=16, L=65, 53 KB rows incl. latents)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import embodied_b200 as embodied            # noqa: E402
from oracle import gen_golden, host_oracle  # noqa: E402
import golden_cases                         # noqa: E402


def make(length, capacity=None, chunksize=1024, online=False, seed=0, **kw):
  return embodied.Replay(
      length, capacity, chunksize=chunksize, online=online, seed=seed, **kw)


@pytest.mark.parametrize('name', sorted(gen_golden.REPLAY_CASES))
def test_row_by_row_matches_reference_golden(name):
  golden_cases.run_replay_case(name, lambda *a: make(*a, staging_rows=8))


@pytest.mark.parametrize('name', sorted(gen_golden.REPLAY_CASES))
def test_add_batch_host_values(name):
  def adder(replay, rows):
    replay.add_batch({k: np.stack([r[k] for r in rows]) for k in rows[0]})
  golden_cases.run_replay_case(name, make, adder)


@pytest.mark.parametrize('name', sorted(gen_golden.REPLAY_CASES))
def test_add_batch_device_values(name):
  def adder(replay, rows):
    batch = {k: np.stack([r[k] for r in rows]) for k in rows[0]}
    for k in ('deter', 'image'):           # latents / images already in HBM
      batch[k] = torch.from_numpy(batch[k]).cuda()
    replay.add_batch(batch)
  golden_cases.run_replay_case(name, make, adder)


def test_consec_stream_matches_reference_golden():
  fix = np.load(golden_cases.GOLDEN / 'consec.npz')
  replay = make(7, 64, chunksize=8, seed=3)
  rng = np.random.default_rng(99)
  for t in range(40):
    for w in range(2):
      replay.add(gen_golden.transition(rng, w, t, gen_golden.SHAPES), w)
  source = embodied.streams.Stateless(replay.sample, 3, 'train')
  stream = iter(embodied.streams.Consec(
      source, length=3, consec=2, prefix=1, strict=True, contiguous=True))
  for i in range(4):
    golden_cases.check_batch(next(stream), None, f'b{i}/', fix)


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_random_streams_vs_oracle(seed):
  rng = np.random.default_rng(seed)
  length = int(rng.integers(1, 9))
  chunksize = int(rng.integers(2, 12))
  capacity = int(rng.integers(length + 1, 60))
  workers = int(rng.integers(1, 5))
  online = bool(rng.integers(0, 2))
  rep = make(length, capacity, chunksize, online, seed, staging_rows=5)
  ora = host_oracle.OracleReplay(
      length, capacity, chunksize, online, seed, ids=itertools.count(1))
  data = np.random.default_rng(seed + 100)
  for t in range(120):
    w = int(rng.integers(0, workers))
    step = {
        'x': data.standard_normal((3,)).astype(np.float32),
        'img': data.integers(0, 256, (16, 16, 3)).astype(np.uint8),
        'is_first': np.asarray(data.integers(0, 6) == 0),
        'is_last': np.asarray(data.integers(0, 6) == 0),
        'lat': np.zeros(64, np.float32)}
    rep.add(dict(step), w)
    ora.add(dict(step), w)
    assert len(rep) == len(ora)
    if len(ora) and rng.integers(0, 4) == 0:
      mode = ['train', 'report'][int(rng.integers(0, 2))]
      a, b = rep.sample(3, mode), ora.sample(3, mode)
      assert sorted(a) == sorted(b)
      for k in b:
        got = a[k].cpu().numpy()
        assert got.dtype == b[k].dtype and got.tobytes() == b[k].tobytes(), k
      T = int(rng.integers(1, length + 1))
      upd = data.standard_normal((3, T, 64)).astype(np.float32)
      # a view of the sampled stepid (fast path) or a host copy (slow path)
      sid = a['stepid'][:, :T] if t % 2 else b['stepid'][:, :T].copy()
      rep.update({'stepid': sid, 'lat': torch.from_numpy(upd).cuda()})
      ora.update({'stepid': b['stepid'][:, :T].copy(), 'lat': upd.copy()})


def test_save_load_roundtrip_on_device(tmp_path):
  replay = make(5, 25, chunksize=4, directory=tmp_path, save_wait=True)
  for step in range(30):
    replay.add({'step': np.int32(step), 'v': np.full(33, step, np.float32)})
  replay.save()
  names = sorted(p.name for p in tmp_path.glob('*.npz'))
  assert names and all(len(n.split('-')) == 4 for n in names)
  with np.load(tmp_path / names[0]) as f:
    assert set(f.keys()) == {'step', 'v', 'stepid'}
  again = make(5, 25, chunksize=4, directory=tmp_path)
  again.load()
  assert len(again) == len(replay) == 25
  for _ in range(10):
    seq = again.sample(2)
    step = seq['step'].cpu().numpy()
    assert (np.diff(step, axis=1) == 1).all()
    assert (seq['v'].cpu().numpy()[..., 0] == step).all()


def test_full_size_properties_config2():
  """BASELINE config 2 shapes: 64x64x3 image + deter f32[8192] + stoch
  f32[32,64]; B=16, L=65, 64 workers.  Checks that do not need the oracle:
  contiguity, single worker per row, stepid == (chunk, index) ordering,
  update -> sample round trip, checksum of checksums over the batch."""
  L, B, W = 65, 16, 64
  replay = make(L, 64 * 200, chunksize=1024, online=True, workers=W,
                staging_rows=64)
  rng = np.random.default_rng(0)
  sums = {}
  for t in range(140):
    img = rng.integers(0, 256, (W, 64, 64, 3)).astype(np.uint8)
    batch = {
        'image': img,
        'reward': rng.standard_normal(W).astype(np.float32),
        'is_first': np.full(W, t % 50 == 0), 'is_last': np.full(W, t % 50 == 49),
        'is_terminal': np.zeros(W, bool),
        'action': rng.integers(0, 5, W).astype(np.int32),
        't': np.full(W, t, np.int32), 'worker': np.arange(W, dtype=np.int32),
        'dyn/deter': torch.full((W, 8192), float(t), device='cuda'),
        'dyn/stoch': torch.zeros((W, 32, 64), device='cuda'),
    }
    replay.add_batch(batch)
    for w in range(W):
      sums[(w, t)] = int(img[w].sum(dtype=np.int64))
  assert len(replay) == W * (140 - L + 1)
  data = replay.sample(B)
  host = {k: v.cpu().numpy() for k, v in data.items()}
  assert host['image'].shape == (B, L, 64, 64, 3)
  assert host['dyn/deter'].shape == (B, L, 8192)
  assert (np.diff(host['t'], axis=1) == 1).all()
  assert (host['worker'] == host['worker'][:, :1]).all()
  assert (host['dyn/deter'][..., 0] == host['t']).all()
  want = np.array([[sums[(int(w), int(t))] for w, t in zip(ws, ts)]
                   for ws, ts in zip(host['worker'], host['t'])])
  assert (host['image'].reshape(B, L, -1).sum(-1, dtype=np.int64) == want).all()
  assert host['is_first'][:, 0].all()
  idx = host['stepid'][..., 16:].astype(np.int64)
  idx = (idx[..., 0] << 24) | (idx[..., 1] << 16) | (idx[..., 2] << 8) | idx[..., 3]
  assert ((np.diff(idx, axis=1) == 1) | (idx[:, 1:] == 0)).all()
  new = torch.full((B, L - 1, 8192), -3.0, device='cuda')
  replay.update({'stepid': data['stepid'][:, 1:], 'dyn/deter': new})
  key0 = host['stepid'][0, 1].tobytes()
  for _ in range(200):
    got = replay.sample(B, mode='report')
    sid = got['stepid'].cpu().numpy()
    hit = np.argwhere((sid.reshape(B, L, 20) == np.frombuffer(key0, np.uint8)).all(-1))
    if len(hit):
      b, t = hit[0]
      assert float(got['dyn/deter'][b, t, 0]) == -3.0
      break


def test_device_replay_loads_directory_written_by_the_reference(tmp_path):
  """tests/golden/ref_chunks was written by the REFERENCE's Replay.save; the HBM-backed product
  loads it (emb_rows_copy imports every slab) and samples what the reference sampled after
  loading the same directory (oracle/gen_golden.py gen_chunkdir)."""
  import shutil
  from oracle import gen_golden
  import golden_cases
  sp = gen_golden.CHUNKDIR_SPEC
  shutil.copytree(golden_cases.GOLDEN / 'ref_chunks', tmp_path / 'chunks')
  want = np.load(golden_cases.GOLDEN / 'ref_chunks_expected.npz')
  replay = embodied.Replay(sp['length'], sp['capacity'], chunksize=sp['chunksize'],
                           directory=str(tmp_path / 'chunks'), seed=sp['seed'], staging_rows=4)
  replay.load()
  assert len(replay) == int(want['len'])
  for i in range(3):
    golden_cases.check_batch(replay.sample(sp['batch']), None, f'sample{i}/', want)
