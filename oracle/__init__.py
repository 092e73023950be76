"""CPU oracle for the Allegro per-edge hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch/numpy restatement of the reference algorithm
(mir-group/allegro v0.7.1, files cited per function) plus the un-vendored
arithmetic it depends on (e3nn: Wigner 3j, spherical harmonics, Irreps;
nequip: ScalarMLPFunction, scatter, Bessel/cutoff embedding, scale/shift).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``allegro_b200`` never imports it; its CUDA path raises if the extension is
missing instead of falling back here.

PARITY UNPINNED: the reference's own tests hold no golden vectors and compare
against e3nn at run time (tests/nn/test_contract_basic.py:120-211,
tests/nn/test_weighter.py:12-54); neither e3nn nor nequip is installed in this
image, so the oracle cannot be executed against the real reference here.  It is
pinned instead by the known-answer values and self-consistency identities in
``tests/golden/o3_known_answers.json`` and by the
property tests the reference uses (equivariance, gradcheck, strict locality).
"""
