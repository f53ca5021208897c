"""CPU oracle for the FE-training + gallery-matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product path (``pets-face-recognition_b200/``) never
imports it and fails loudly if the CUDA library is missing.

The oracle is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference's
arithmetic, written from the reference's behaviour, each function citing the
reference file:line it follows:

* ``swin_oracle``  - models/swin.py:8-241 (berniwal-variant Swin Transformer)
* ``head_oracle``  - losses/large_margin.py:10-84, losses/losses.py:7-28,
                     losses/__init__.py:37-46, SGD step of
                     configs/dog_fe/fe_dogs_config.py:123-133
* ``rank_oracle``  - engine/controller.py:48-57,77-91 (Recall@K loop) and
                     configs/dog_fe/fe_dogs_config.py:89-93 (similarity_f)

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the restatement is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run in the build container by ``tests/golden/make_golden.py`` (imports
``/root/reference/models/swin.py``, ``losses/*`` and - through stubbed
third-party imports - the unmodified ``engine/controller.py``).  The vectors it
produced are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.
"""
