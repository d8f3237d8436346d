"""CPU checks of the oracle against the committed golden fixtures (tests/golden)."""
import os

import numpy as np
import torch

from oracle import lift_oracle, scn_cpu
from oracle import scn_oracle as O
from mm2d3d_b200.unet import UNetSCN

G = os.path.join(os.path.dirname(__file__), "golden")


def test_lift_oracle_matches_reference_fixture():
    """lift_ref.npz was produced by the reference's own L2G_classifier_2D (make_golden.py)."""
    z = np.load(os.path.join(G, "lift_ref.npz"))
    fmap = torch.from_numpy(z["fmap"]).requires_grad_(True)
    offs = np.concatenate([[0], np.cumsum(z["counts"])])
    idx = [z["idx"][offs[i]:offs[i + 1]] for i in range(len(z["counts"]))]
    out = lift_oracle.lift2d(fmap, idx)
    assert torch.equal(out.detach(), torch.from_numpy(z["lifted"]))
    (g,) = torch.autograd.grad(out, fmap, torch.from_numpy(z["grad_out"]))
    assert torch.allclose(g, torch.from_numpy(z["grad_fmap"]), atol=1e-12)


def test_structure_fixture():
    z = np.load(os.path.join(G, "structure_small.npz"))
    meta = O.Metadata(z["coords"], 4096)
    assert np.array_equal(meta.p2v, z["p2v"]) and np.array_equal(meta.npts, z["npts"])
    # literal loop restatement agrees with the fixture too
    ids, _ = O.first_occurrence_ids_loop(O.pack_keys(z["coords"]))
    assert np.array_equal(ids, z["p2v"])
    s = 4096
    for lvl in range(4):
        assert np.array_equal(meta.coords_at(s), z[f"coords_l{lvl}"])
        assert np.array_equal(meta.nbr(s), z[f"nbr_l{lvl}"])
        if lvl < 3:
            parent, off, nc = meta.down(s)
            assert np.array_equal(parent, z[f"parent_l{lvl}"]) and np.array_equal(off, z[f"off_l{lvl}"])
            assert np.array_equal(O.child_table(parent, off, nc), z[f"child_l{lvl}"])
        s //= 2


def test_unet_small_fixture_reproduces():
    z = np.load(os.path.join(G, "unet_small.npz"))
    net = UNetSCN(in_channels=3, m=4, num_planes=4, full_scale=64, backend=scn_cpu).double()
    sd = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    net.load_state_dict(sd, strict=False)
    x = torch.from_numpy(z["feats"]).double().requires_grad_(True)
    out = net([torch.from_numpy(z["coords"]), x])
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=1e-12)
    (gx,) = torch.autograd.grad(out, x, torch.from_numpy(z["grad_out"]))
    assert torch.allclose(gx, torch.from_numpy(z["grad_feats"]), atol=1e-12)


def test_augment_oracle_matches_reference_fixture():
    """augment_ref.npz was produced by the reference's own augment_and_scale_3d (make_golden.py)."""
    from oracle import augment_oracle
    z = np.load(os.path.join(G, "augment_ref.npz"))
    offs = z["offsets"]
    for i in range(len(offs) - 1):
        p = z["points"][offs[i]:offs[i + 1]]
        u = z["u"][i] if z["transl"][i] else None
        ci, keep, mn, off = augment_oracle.scale_points(p, z["rot"][i], u, int(z["scale"]), int(z["full_scale"]))
        assert np.array_equal(ci, z["coords"][offs[i]:offs[i + 1]]) and np.array_equal(keep, z["keep"][offs[i]:offs[i + 1]])
        assert np.array_equal(mn, z["min_value"][i]) and np.array_equal(off, z["offset"][i])
        ci, keep, _, _ = augment_oracle.scale_points(p, z["rot"][i], None, int(z["scale"]), int(z["small_full_scale"]))
        assert np.array_equal(ci, z["coords_small"][offs[i]:offs[i + 1]])
        assert np.array_equal(keep, z["keep_small"][offs[i]:offs[i + 1]]) and 0 < keep.mean() < 1


def test_raster_oracle_last_point_wins():
    """The duplicate-pixel rule the CUDA rasteriser has to reproduce (nuscenes_dataloader.py:274-278)."""
    from oracle import raster_oracle
    idx = np.array([[1, 2], [0, 0], [1, 2], [3, 1], [1, 2]], dtype=np.int64)
    vals = np.array([5.0, 7.0, 6.0, 8.0, 9.0], dtype=np.float32)
    m = raster_oracle.rasterize(idx, vals, 4, 3, -100.0)
    assert m[1, 2] == 9.0 and m[0, 0] == 7.0 and m[3, 1] == 8.0
    assert (m == -100.0).sum() == 4 * 3 - 3
    fidx, (fm,) = raster_oracle.fliplr(idx, [m], 3)
    assert np.array_equal(raster_oracle.rasterize(fidx, vals, 4, 3, -100.0), fm)
    img = np.arange(3 * 4 * 3, dtype=np.float32).reshape(3, 4, 3)
    f = raster_oracle.rgb_feats(img, idx)
    assert f.shape == (5, 3) and np.array_equal(f[0], img[:, 1, 2])


def test_heads_oracle_matches_reference_fixture():
    """RGB mask vs what the reference's own Net3DSeg.forward produced; KL term vs the reference's torch lines."""
    import torch
    from oracle import heads_oracle
    z = np.load(os.path.join(G, "heads_ref.npz"))
    x = torch.from_numpy(z["feats"]).requires_grad_(True)
    w, b = torch.from_numpy(z["w"]).requires_grad_(True), torch.from_numpy(z["b"]).requires_grad_(True)
    y = heads_oracle.rgb_mask(x, w, b)
    assert np.array_equal(y.detach().numpy(), z["masked"])
    y.backward(torch.from_numpy(z["g"]))
    assert np.allclose(x.grad.numpy(), z["dx"], rtol=0, atol=0) and np.allclose(w.grad.numpy(), z["dw"], rtol=1e-6)
    pred = torch.from_numpy(z["pred"]).requires_grad_(True)
    loss = heads_oracle.cross_modal_kl(pred, torch.from_numpy(z["target"]))
    assert float(loss.detach()) == float(z["loss"])
    loss.backward()
    assert np.array_equal(pred.grad.numpy(), z["dpred"])


def test_heads3d_oracle_matches_reference_fixture():
    """The two 3D heads vs what the reference's own Net3DSeg.forward produced (real `linear` / `linear_point`), the 3D
    cross-modal term and the autograd gradients of the fixture's combined scalar."""
    import torch
    from oracle import heads_oracle
    z = np.load(os.path.join(G, "heads3d_ref.npz"))
    t = lambda k: torch.from_numpy(z[k])
    feat = t("feat").requires_grad_(True)
    w1, b1, w2, b2 = (t(k).requires_grad_(True) for k in ("w1", "b1", "w2", "b2"))
    l1, l2, loss = heads_oracle.heads3d(feat, w1, b1, w2, b2, t("target"))
    assert np.array_equal(l1.detach().numpy(), z["logit1"]) and np.array_equal(l2.detach().numpy(), z["logit2"])
    assert float(loss.detach()) == float(z["loss"])
    ((l1 * t("g1")).sum() + (l2 * t("g2")).sum() + float(z["lam"]) * loss).backward()
    for got, key in ((feat.grad, "d_feat"), (w1.grad, "d_w1"), (b1.grad, "d_b1"), (w2.grad, "d_w2"), (b2.grad, "d_b2")):
        assert np.allclose(got.numpy(), z[key], rtol=1e-6, atol=1e-7), key
