"""CPU oracle for the STFT -> UNetBaseline -> depth-loss path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and there only as the checker (or as the timed CPU baseline), never
as the thing shipped.  The product path (``audio_depth_estimation_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  Every function here is checked against outputs of the
unmodified reference modules (run in the build container from
``/root/reference`` by ``oracle/gen_golden.py``; vectors committed under
``tests/golden/``) by ``tests/test_oracle_golden.py``.  The reference itself
holds no golden vectors or tests for this path.
"""
