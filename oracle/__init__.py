"""CPU oracle for the two STswinCL hot paths.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product package
``stswincl_b200`` must not import this package.

Parity status: PINNED.  The reference repo ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
modules themselves, imported in the authoring container by
``oracle/make_goldens.py`` (which needs ``/root/reference``) and committed as
fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks
the restatement against those fixtures on every run, with no access to the
reference tree.  The training-step kernels (OHEM cross-entropy, LARS, key-encoder
EMA) have their own restatement ``trainaux_oracle.py``, pinned the same way by
``make_goldens_trainaux.py`` -> ``tests/golden/trainaux_cases.npz``
(``tests/test_trainaux_oracle.py``).
"""
from . import index_oracle, swin_oracle, loss_oracle  # noqa: F401
