"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the MSDeformAttn hot path.

Nothing under ``uvhand_b200/`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg.

Contents
--------
msda_oracle.c / msda_oracle_body.inc
    plain-C restatement of the reference CUDA arithmetic (index-exact); ``c_oracle`` is
    its ctypes loader.
msda_numpy.py
    an independent vectorised numpy restatement of the same arithmetic (cross-check of
    the C file; pure integer/predicate work is compared bit-for-bit).
grid_sample_port.py
    port of the reference's CPU path ``ms_deform_attn_core_pytorch``
    (models/ops/functions/ms_deform_attn_func.py:42-62), the thing the CPU baseline times.
make_golden.py
    imports the UNMODIFIED reference from /root/reference (build container only) and
    writes the golden vectors under tests/golden/.
ref_cuda_harness.cu
    C-ABI harness that #includes the reference's own CUDA kernels from /root/reference
    and is compiled to oracle/_ref/ (git-ignored; travels to the GPU box).

Parity is pinned: see tests/test_oracle.py and tests/golden/README.md.
"""
