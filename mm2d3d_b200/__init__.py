"""B200-native implementation of MM2D3D's 3D-branch hot path (UNetSCN on SparseConvNet
semantics + the 2D->3D feature lift).  See DESIGN.md.

``mm2d3d_b200.scn`` is the ``sparseconvnet``-shaped module surface (CUDA only; importing it
loads ``libmm3d.so`` and raises if that library has not been built), ``mm2d3d_b200.unet`` the
``UNetSCN`` backbone, ``mm2d3d_b200.lift`` the 2D->3D lift, ``mm2d3d_b200.synth`` the
synthetic-scan generator used by the benchmark and the tests.
"""
__version__ = "0.1.0"
