"""Old-generation agent API (director/) on this runtime: the adapter half of SURVEY 8f rank 3.

`director/` is written against an older generation of the embodied API:
``policy(obs, state, mode) -> (outs, state)``, ``train(data, state) -> (outs, state, metrics)``,
``report(data) -> metrics``, ``dataset(generator_fn)``, and a five-factory ``run.train``
(director/jaxagent.py:104-230, director/train.py:61-65).  `adapter.OldApiAgent` presents such an
agent through the current protocol (embodied/core/base.py:1-31) so that `embodied_b200.Driver`,
`Replay` and `run.train` run it unchanged; `adapter.train` is the five-factory entry point.

The director agent's own networks (director/agent.py, hierarchy.py, nets.py -- a separate model
family on vendored ninjax 1.2.0 + tensorflow_probability) are NOT rebuilt here: BASELINE config 4
exercises Driver + Replay + this adapter with any old-API agent (tests/test_oldapi_host.py,
tests/test_gpu_configs.py::test_config4_*).
"""
from .adapter import OldApiAgent, train
