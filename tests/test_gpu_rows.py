"""Row engine (libembodied_b200.so) against numpy / the oracle, through the
C ABI, byte-exact.  Covers gather, append, update(-1 skips), annotate flags,
consec fill, u8->f32 normalise, typed action mask, empty and ragged inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
from embodied_b200 import _lib          # noqa: E402
from oracle import host_oracle          # noqa: E402


@pytest.fixture(scope='module')
def lib():
  assert torch.cuda.is_available(), 'gpu tests need CUDA'
  return _lib.load()


def dev(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rowbytes(a):
  return int(a.dtype.itemsize * np.prod(a.shape[1:], dtype=np.int64))


def stream():
  return torch.cuda.current_stream().cuda_stream


SHAPES = [
    ('flag', np.bool_, ()), ('rew', np.float32, ()), ('stepid', np.uint8, (20,)),
    ('vec7', np.float32, (7,)), ('odd', np.uint8, (12291,)),
    ('image', np.uint8, (64, 64, 3)), ('deter', np.float32, (8192,)),
    ('stoch', np.float32, (32, 64)), ('i64', np.int64, (3,)),
    ('half', np.float16, (130,)), ('big', np.uint8, (40000,)),
]


def make_table(rng, rows):
  out = {}
  for name, dt, shape in SHAPES:
    dt = np.dtype(dt)
    if dt == np.bool_:
      out[name] = rng.integers(0, 2, (rows, *shape)).astype(bool)
    elif np.issubdtype(dt, np.integer):
      out[name] = rng.integers(0, 255, (rows, *shape)).astype(dt)
    else:
      out[name] = rng.standard_normal((rows, *shape)).astype(dt)
  return out


@pytest.mark.parametrize('nrows,seed', [(1, 0), (7, 1), (130, 2), (1040, 3)])
def test_gather_matches_numpy(lib, nrows, seed):
  rng = np.random.default_rng(seed)
  table = make_table(rng, 300)
  src = rng.integers(0, 300, nrows).astype(np.int64)
  tdev = {k: dev(v) for k, v in table.items()}
  out = {k: torch.zeros((nrows, *v.shape[1:]), dtype=tdev[k].dtype, device='cuda')
         for k, v in table.items()}
  keys = [_lib.Key(src=tdev[k].data_ptr(), dst=out[k].data_ptr(),
                   src_stride=rowbytes(v), dst_stride=rowbytes(v),
                   row_bytes=rowbytes(v), op=_lib.OP_COPY)
          for k, v in table.items()]
  _lib.check(lib.emb_replay_gather(
      _lib.keys_array(keys), len(keys), dev(src).data_ptr(), nrows, nrows,
      stream()))
  torch.cuda.synchronize()
  for k, v in table.items():
    assert out[k].cpu().numpy().tobytes() == v[src].tobytes(), k


def test_append_then_update_with_skips(lib):
  rng = np.random.default_rng(5)
  table = make_table(rng, 64)
  tdev = {k: dev(v) for k, v in table.items()}
  new = make_table(rng, 20)
  ndev = {k: dev(v) for k, v in new.items()}
  dst = rng.permutation(64)[:20].astype(np.int64)
  dst[[3, 11]] = -1                      # evicted rows are skipped
  keys = [_lib.Key(src=ndev[k].data_ptr(), dst=tdev[k].data_ptr(),
                   src_stride=rowbytes(v), dst_stride=rowbytes(v),
                   row_bytes=rowbytes(v), op=_lib.OP_COPY)
          for k, v in table.items()]
  _lib.check(lib.emb_replay_scatter_update(
      _lib.keys_array(keys), len(keys), dev(dst).data_ptr(), 20, stream()))
  torch.cuda.synchronize()
  for k, v in table.items():
    want = v.copy()
    ok = dst >= 0
    want[dst[ok]] = new[k][ok]
    assert tdev[k].cpu().numpy().tobytes() == want.tobytes(), k


@pytest.mark.parametrize('batch,window', [(1, 1), (3, 2), (16, 65), (5, 7)])
def test_annotate_and_consec_match_oracle(lib, batch, window):
  rng = np.random.default_rng(batch * 100 + window)
  rows = 500
  first = rng.integers(0, 3, rows) == 0
  last = rng.integers(0, 5, rows) == 0
  src = rng.integers(0, rows, batch * window).astype(np.int64)
  f_dev, l_dev = dev(first), dev(last)
  of = torch.zeros(batch * window, dtype=torch.bool, device='cuda')
  ol = torch.zeros(batch * window, dtype=torch.bool, device='cuda')
  oc = torch.zeros(batch * window, dtype=torch.int32, device='cuda')
  keys = [
      _lib.Key(src=f_dev.data_ptr(), dst=of.data_ptr(), src_stride=1,
               dst_stride=1, row_bytes=1, op=_lib.OP_FIRST),
      _lib.Key(src=l_dev.data_ptr(), dst=ol.data_ptr(), src_stride=1,
               dst_stride=1, row_bytes=1, op=_lib.OP_LAST,
               aux=f_dev.data_ptr(), aux_stride=1),
      _lib.Key(dst=oc.data_ptr(), dst_stride=4, row_bytes=4,
               op=_lib.OP_FILL32, fill=7)]
  _lib.check(lib.emb_replay_gather(
      _lib.keys_array(keys), 3, dev(src).data_ptr(), batch * window, window,
      stream()))
  torch.cuda.synchronize()
  want = host_oracle.annotate({
      'is_first': first[src].reshape(batch, window),
      'is_last': last[src].reshape(batch, window)})
  assert (of.cpu().numpy().reshape(batch, window) == want['is_first']).all()
  assert (ol.cpu().numpy().reshape(batch, window) == want['is_last']).all()
  assert (oc.cpu().numpy() == 7).all()


@pytest.mark.parametrize('n,shape', [(1, (8, 8, 3)), (5, (64, 64, 3)), (256, (64, 64, 3)), (3, (84, 84, 1))])
def test_stage_obs_normalise_and_append(lib, n, shape):
  rng = np.random.default_rng(n)
  img = rng.integers(0, 256, (n, *shape)).astype(np.uint8)
  if n >= 5:
    img[0].reshape(-1)[:256] = np.arange(256)      # every byte value
  table = torch.zeros((2 * n + 3, int(np.prod(shape))), dtype=torch.uint8, device='cuda')
  norm = torch.zeros((n, *shape), dtype=torch.float32, device='cuda')
  dst = rng.permutation(2 * n + 3)[:n].astype(np.int64)
  rb = int(np.prod(shape))
  src = dev(img)
  keys = [_lib.Key(src=src.data_ptr(), dst=table.data_ptr(), dst2=norm.data_ptr(),
                   src_stride=rb, dst_stride=rb, dst2_stride=rb * 4,
                   row_bytes=rb, op=_lib.OP_NORM_U8_F32)]
  _lib.check(lib.emb_driver_stage_obs(
      _lib.keys_array(keys), 1, dev(dst).data_ptr(), n, stream()))
  torch.cuda.synchronize()
  want = host_oracle.normalize_image(img)
  assert norm.cpu().numpy().tobytes() == want.tobytes()      # bit-exact fp32
  assert (table.cpu().numpy()[dst] == img.reshape(n, -1)).all()


def test_mask_actions_typed_multiply(lib):
  n = 9
  rng = np.random.default_rng(0)
  acts = {
      'f': rng.standard_normal((n, 6)).astype(np.float32),
      'i': rng.integers(-5, 5, (n,)).astype(np.int32),
      'l': rng.integers(-5, 5, (n, 2)).astype(np.int64),
      'd': rng.standard_normal((n, 3)).astype(np.float64),
      'b': rng.integers(0, 2, (n, 4)).astype(bool),
      'h': rng.standard_normal((n, 5)).astype(np.float16),
  }
  acts['f'][0, :3] = [-1.5, np.inf, np.nan]        # -0.0, nan, nan when masked
  acts['f'][1, 0] = -0.0
  is_last = np.zeros(n, bool)
  is_last[[0, 4, 8]] = True
  from embodied_b200.core import driver_ops
  got = driver_ops.DeviceOps().mask_actions(acts, is_last)
  with np.errstate(invalid='ignore'):
    want = host_oracle.mask_actions(acts, is_last)
  for k in acts:
    assert got[k].dtype == want[k].dtype, k
    assert got[k].tobytes() == want[k].tobytes(), k


def test_empty_launches_are_noops(lib):
  keys = _lib.keys_array([_lib.Key(row_bytes=4, op=_lib.OP_COPY, src=8, dst=8)])
  assert lib.emb_rows_copy(keys, 1, None, None, 0, 0, stream()) == 0
  assert lib.emb_rows_copy(keys, 0, None, None, 5, 0, stream()) == 0


def test_launch_counter_counts(lib):
  before = _lib.launch_count()
  a = torch.arange(64, dtype=torch.uint8, device='cuda')
  b = torch.zeros(64, dtype=torch.uint8, device='cuda')
  keys = _lib.keys_array([_lib.Key(src=a.data_ptr(), dst=b.data_ptr(),
                                   src_stride=16, dst_stride=16, row_bytes=16)])
  _lib.check(lib.emb_rows_copy(keys, 1, None, None, 4, 0, stream()))
  torch.cuda.synchronize()
  assert _lib.launch_count() == before + 1
  assert (a == b).all()
