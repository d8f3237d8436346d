"""CPU restatement (test infrastructure only) of the point-wise prologue and loss around the 3D network:

* RGB mask -- ``Net3DSeg.forward`` (``3d_net/model.py:46-48``): ``mask = sigmoid(linear_rgb_mask(feats))``;
  ``feats *= mask``.
* cross-modal loss -- ``TrainModel.cross_modal_loss`` (``train.py:157-184``), one of its two identical terms:
  ``F.kl_div(F.log_softmax(pred, 1), F.softmax(target.detach(), 1), reduction="none").sum(1).mean()``.

* the two heads -- ``Net3DSeg.forward`` (``3d_net/model.py:49``: ``x = self.linear(out_3D_feature)``) and
  ``L2G_classifier_3D.forward`` (``:85``: ``point_wise_pre = self.linear_point(input_3D_feature)``), with the ``loss_3d``
  term of ``cross_modal_loss`` on the second one.  Pinned against ``tests/golden/heads3d_ref.npz`` (the logits by the
  reference's own ``Net3DSeg.forward`` with its real heads, its backbone replaced by a 16-channel pass-through).

Pinned against ``tests/golden/heads_ref.npz``: the mask part was produced by the reference's own ``Net3DSeg.forward``
(its ``net_3d`` replaced by a pass-through so that the call runs on CPU), the loss part by the quoted torch lines."""
import torch
import torch.nn.functional as F


def rgb_mask(feats, weight, bias):
    """feats [N, C] float32, weight [1, C], bias [1] (``nn.Linear(C, 1)``) -> masked feats [N, C]."""
    mask = torch.sigmoid(F.linear(feats, weight, bias))              # model.py:46-47
    return feats * mask                                              # :48 (in place there)


def cross_modal_kl(pred, target):
    return F.kl_div(F.log_softmax(pred, dim=1), F.softmax(target.detach(), dim=1), reduction="none").sum(1).mean()


def heads3d(feat, w1, b1, w2, b2, target=None):
    """(seg_logit, seg_logit_point, loss_3d): model.py:49, :85 and train.py:174-182."""
    l1 = F.linear(feat, w1, b1)
    l2 = F.linear(feat, w2, b2)
    loss = cross_modal_kl(l2, target) if target is not None else l2.new_zeros(())
    return l1, l2, loss
