"""CPU oracle for the TM-GCN propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tmgcn_b200/`` (the product) may
import this package.  The only permitted importers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / reported CPU baseline,
never as the thing shipped.

What it is: a restatement, in numpy + PyTorch-CPU, of the arithmetic the
reference (IBM/TM-GCN, ``TensorGCN-master/``) performs on this path.  The
reference has no native code: every operation is an ATen CPU kernel driven by
Python, so "reference semantics" means "the same ATen ops in the same dtypes
and the same order".  Each function cites the reference lines it follows.

Parity pin: the reference ships NO golden vectors or tests,
so the oracle is pinned against outputs of the *unmodified reference code
imported in the build container* -- ``tests/golden/make_golden.py`` imports
``embedding_help_functions`` and exec()s ``func_MProduct`` /
``func_MProduct_dense`` / ``create_matrix_M`` straight from
``/root/reference`` and stores their outputs as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against those
files (bit-exact indices, <=1e-14 rel on fp64 values, exact fp32 equality on
the model outputs).  The fixtures travel; ``/root/reference`` does not.
"""
from .tmgcn_oracle import *  # noqa: F401,F403
