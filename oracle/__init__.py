"""CPU oracle for the ProdSearch embedding-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``prodsearch_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is the checker, never the
product: the product path raises if the CUDA library is missing.

The reference (kepingbi/ProdSearch) is pure PyTorch, so the oracle is a plain
functional fp32 restatement on CPU tensors (torch + numpy), written from the
reference's behaviour and citing the reference file:line each function follows.
Parity is PINNED: ``tests/golden/make_golden.py`` runs the reference's own modules
(imported read-only from /root/reference in the build container, uint8->bool
``masked_fill`` shim, injected negatives) and commits inputs+outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle
function against them.
"""
from .ref_models import *  # noqa: F401,F403
from .ranking import *  # noqa: F401,F403
from . import batches  # noqa: F401
