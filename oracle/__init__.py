"""CPU oracle for the CausalDiffAE denoising hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / reported baseline.
The shipped path (``causaldiffae_b200``) never imports ``oracle`` and raises if
its CUDA extension is missing.

What it is: a from-scratch plain-PyTorch-fp32 / numpy-float64 restatement of the
reference algorithm (Akomand/CausalDiffAE, ``improved_diffusion/*``), each
function citing the reference ``file:line`` it follows.  The arithmetic lives in
a third-party dependency of the reference (PyTorch ATen; pinned there as
torch==2.0.1+cu118, here torch 2.11) so "the reference's numbers" are defined as
the reference Python executed by this container's torch in fp32.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *reference itself* imported in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by
``tests/test_oracle_golden.py``).  ``oracle/refshim.py`` imports the unmodified reference
(in place, or from the copy ``oracle/build_ref.py`` puts under the git-ignored ``oracle/_ref``)
for fixture generation and for ``bench.py --impl reference``.
"""
from . import schedules, model, diffusion  # noqa: F401
