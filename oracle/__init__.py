"""CPU oracle for the NeuS volume-rendering hot path (TEST INFRASTRUCTURE ONLY).

This package restates, on the CPU, the algorithm of the reference path
(PJLab-ADG/neuralsim @ faba099 + nr3d_lib @ e1e87d1): LoTD hash-grid encoding,
occupancy-grid ray marching, pack_ops, the NeuS SDF->alpha maths, the tiny
autocast MLPs, the `neus_ray_query_march_occ_multi_upsample_compressed`
orchestration and the volume integration.  Every function cites the reference
file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`neuralsim_b200/`) never does: it fails loudly when its CUDA library is
missing.

Pinning status (see DESIGN.md "Oracle"):
  * pack_ops / raysample / NeuS maths / merges: pinned by the reference's own
    known-answer vectors (tests/golden/ref_kat.json) and by golden vectors
    produced by importing the reference's pure-PyTorch functions in the build
    container (tests/golden/make_golden.py).
  * LoTD / ray marching / alpha_to_vw kernels: the reference ships no numeric
    fixtures for them and its CUDA code cannot execute in the GPU-less build
    container; they are pinned on the GPU box against the reference's own
    kernels compiled into oracle/_ref (tests/test_ref_parity_gpu.py).
"""
