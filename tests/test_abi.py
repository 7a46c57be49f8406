"""The C-ABI library loads here (no GPU) and exports every symbol that
include/embodied_b200.h declares; argument validation works without a device."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def lib():
  from embodied_b200 import build, _lib
  build.build()
  return _lib.load()


def declared_symbols():
  text = (ROOT / 'include' / 'embodied_b200.h').read_text()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(emb_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(lib):
  names = declared_symbols()
  assert len(names) >= 10, names
  for name in names:
    assert hasattr(lib, name), f'{name} declared in the header but not exported'


def test_binding_lists_every_symbol(lib):
  from embodied_b200 import _lib
  assert sorted(_lib.EXPORTS) == declared_symbols()


def test_abi_version_and_struct_size(lib):
  from embodied_b200 import _lib
  assert lib.emb_abi_version() == 1
  assert ctypes.sizeof(_lib.Key) == 80   # 4 pointers + 4 u64 + 4 x 32 bit


def test_argument_validation_needs_no_device(lib):
  from embodied_b200 import _lib
  keys = _lib.keys_array([_lib.Key(row_bytes=4, op=_lib.OP_FIRST)])
  # FIRST is not a valid op for an append
  assert lib.emb_replay_append_rows(keys, 1, 8, 4, None) == -2
  assert b'op 1' in lib.emb_last_error()
  assert lib.emb_replay_gather(keys, 1, None, 4, 2, None) == -1
  assert lib.emb_replay_gather(keys, 1, 8, 5, 2, None) == -1   # 5 % 2
  assert lib.emb_rows_copy(keys, 99, None, None, 1, 0, None) == -1
  assert lib.emb_rows_copy(keys, 1, None, None, 0, 0, None) == 0   # empty: no-op
  # chunk export / import: argument checks before any copy is issued
  assert lib.emb_replay_export_chunk(keys, 1, -1, 4, None) == -1
  assert lib.emb_replay_export_chunk(keys, 1, 0, 4, None) == -2      # key without src / dst
  assert b'src and dst' in lib.emb_last_error()
  assert lib.emb_replay_import_chunk(keys, 1, 0, 0, None) == 0       # empty: no-op


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
  from embodied_b200 import build, _lib
  monkeypatch.setattr(_lib, '_LIB', None)
  monkeypatch.setattr(build, 'LIB', tmp_path / 'nope.so')
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    _lib.load()


def test_no_cuda_fails_loudly():
  import torch
  if torch.cuda.is_available():
    pytest.skip('CUDA present')
  import embodied_b200 as embodied
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    embodied.Replay(length=2, capacity=4)
  from embodied_b200.core import driver_ops
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    driver_ops.DeviceOps()
