"""CPU oracle for the SDVAR draft-then-verify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``sdvar_b200`` may import,
call, link or execute anything under ``oracle/``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
do, and there only as the checker or as the timed CPU reference.

Layers (see oracle/README.md):
  * ``oracle.ref_model``  -- PyTorch-CPU fp32 restatement of the reference's runnable
    functions (VAR blocks, get_logits, sampler, VQ next-input, decoder, baseline loop,
    teacher-forced forward) plus the corrected draft->verify loop spec.  Pinned against the
    real reference by ``oracle/make_golden.py`` -> ``tests/golden``.
  * ``oracle/spec_c``     -- plain-C, bit-exact arithmetic spec of the two index-producing
    kernels (sampling epilogue K3, speculative verify K4).  The CUDA kernels must match it
    bit for bit; it is pinned against the reference's torch sampler on golden vectors.
  * parity status: sampler / VQ / blocks / baseline loop = pinned by reference outputs;
    verify rule ``min(1,p/q)`` + residual resample and the repaired loop = PARITY UNPINNED
    (nothing in the reference computes them).
"""
