"""CPU oracle for the SGCDet view-transform hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``sgcdet_b200/`` may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the timed CPU arm, never as
the product).

The reference ships no CPU implementation of this path (DFA3D registers CUDA only,
``packages/3D-deformable-attention/DFA3D/dfa3D/ops/csrc/cuda/cudabind.cpp:41-44,67-70``;
the plugin's ``_DFA3D`` forward has no CPU branch, ``deformable_cross_attention.py:482-489``),
so this package is a *restatement* (torch CPU, fp32 or fp64) of the reference algorithm,
each function citing the reference file:line it follows.

Pinning status (see DESIGN.md "Oracle pinning"):
  * operator level (``dfa3d_ref``): pinned against the reference's own CUDA kernels, compiled
    unmodified from /root/reference into ``oracle/_ref`` by ``oracle/build_ref.py`` and run on the
    B200 box (``tests/test_gpu_ref_ext.py``), and against golden vectors in ``tests/golden``
    generated from that run (``tests/golden/make_golden_gpu.py``).
  * module level (``path_ref``): pinned against the reference's own Python modules imported from
    /root/reference with mmcv/mmdet stand-ins (``tests/golden/make_golden_plugin.py``).
"""
