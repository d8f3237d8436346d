"""CPU oracle for the MM2D3D 3D-branch hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``mm2d3d_b200/``
imports it; the product path fails loudly when its CUDA library is missing.

PARITY STATUS
-------------
* 3D path (``scn_oracle``): **parity unpinned by any reference-owned test**.  The
  arithmetic lives in facebookresearch/SparseConvNet@dcf6a7ff540e1825ffe48ba6b2c1493ba18788b2
  (``/root/reference/environment.yml:37``), which is neither vendored in
  ``/root/reference`` nor installed, and neither that project nor the reference ships a
  test with values.  The restatement follows the reference's call sites
  (``3d_net/scn_unet.py:8-126``, ``lib/dataset/__init__.py:62-67``) and
  SparseConvNet's published semantics; it is pinned
  independently by (a) a dense ``torch.nn.functional.conv3d`` equivalence on small
  grids, (b) the worked micro-example of SURVEY.md A.10 and (c) float64 gradcheck.
* 2D->3D lift (``lift_oracle``): pinned against the reference itself
  (``2d_net/model.py:131-137,163-173`` imported in the build container; fixtures
  under ``tests/golden`` with the generating script).
* points -> voxel coordinates (``augment_oracle``): pinned against the reference itself
  (``augment_and_scale_3d`` of ``lib/utils/augmentation_3d.py`` imported and run; ``augment_ref.npz``).
* RGB mask / cross-modal KL term (``heads_oracle``): the mask is pinned against the reference's own
  ``Net3DSeg.forward`` (``3d_net/model.py:44-58``, backbone replaced by a pass-through so it runs on CPU), the loss
  against the torch lines of ``train.py:157-184`` (``heads_ref.npz``).
* point values -> image maps (``raster_oracle``): the loaders build these inline in ``__getitem__`` (needs the dataset
  on disk), so the numpy statements of ``lib/dataset/nuscenes_dataloader.py:274-278`` are restated; numpy's indexed
  assignment is the definition -- **no reference-produced fixture**.
"""
