"""Oracle package -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot path (fpthink/PDGN) that the CUDA kernels are checked against.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; nothing under pdgn_b200/ does (tests/test_boundary.py enforces that by grepping the package).

  oracle.cpu         ctypes binding of liboracle.so (pdgn_oracle.c: exact fmaf restatement of the kernels)
  oracle.torch_ref   restatement of the reference's pure-PyTorch formulations (Chamfer, all-pairs CD,
                     MMD/COV/1-NNA, edge features, naive kNN), pinned by tests/golden/*.npz which were
                     produced by the reference's own Python code (tests/golden/make_golden.py)
  oracle.ref_kernels ctypes binding of oracle/_ref/libpdgn_ref.so = the reference's .cu files recompiled
                     for sm_100a (GPU box only)
"""
