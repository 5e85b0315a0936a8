"""CPU oracle for the Zonos inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference path that
`zonos_b200` replaces.  It exists to CHECK the CUDA
product path; it is never the thing shipped or measured as the product.

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py`.  Nothing under `zonos_b200/` imports
it, and `zonos_b200` raises when the CUDA library is missing rather than
falling back to anything here.

Parity status (see DESIGN.md "Oracle"):
  * codebook pattern, sampler, EOS state machine, generate loop, transformer
    backbone, heads/CFG, DAC decode: PINNED against outputs of the reference
    itself (imported from /root/reference in the build container by
    `tests/golden/make_golden.py`, fixtures committed under `tests/golden/`),
    plus the one known-answer example the reference holds
    (`zonos/codebook_pattern.py:26-29`).
  * hybrid (Mamba2) backbone: PARITY UNPINNED by the reference - `mamba_ssm`
    and `causal_conv1d` are not vendored in the reference and not installed;
    the restatement follows mamba_ssm 2.2.5's published step recurrence and is
    cross-checked against `transformers`' pure-PyTorch `Mamba2Mixer`.

Each function cites the reference file:line it follows (paths relative to the
reference repository root).
"""
