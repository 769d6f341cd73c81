"""
oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (plain PyTorch fp32, functional) restatement of the vp-suite recurrent-rollout hot path:
the Shi-et-al. ConvLSTM with peepholes, the ndrplz ConvLSTM cell, the ST-LSTM (PredRNN-V2) cell,
PhyDNet's PhyCell, the DCGAN encoder/decoder blocks and the three rollout loops that drive them.
Every function cites the reference file:line (relative to the vp-suite checkout) it follows.

Who may import this package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- always as the checker or the timed CPU baseline,
never as a compute path of ``vp_suite_b200`` (which fails loudly when its CUDA library is missing).

Parity pin: ``oracle/make_golden.py`` imports the real reference from ``/root/reference`` (authoring
container only), loads deterministic synthetic weights into the reference's own modules, runs them and
stores inputs seeds + outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this
restatement against those vectors, so the oracle is pinned to outputs of the reference itself.
The reference holds no golden vectors of its own (its tests assert shapes only).
"""
from . import blocks, models, weights, shapes  # noqa: F401
