"""MultiCropWrapper -- drop-in for `utils.utils.MultiCropWrapper` (utils/utils.py:611-646), SURVEY 8(f) rank 2.

Same constructor, attributes (`backbone`, `head`), state_dict keys and forward contract: crops of equal resolution that
are adjacent in the list go through the backbone together, the per-group features are concatenated crop-major and the
head runs once on the `[n_crops * B, D]` matrix (the layout DINOLoss assumes).  Differences are plumbing only: the
resolution groups are found with plain Python (no tensor ops), and the features are concatenated ONCE instead of
growing `output = torch.cat((output, _out))` from an empty tensor on every group (which re-copies the rows already
gathered and makes the first copy a dtype-promoting cat with an fp32 empty tensor under autocast).
"""
from __future__ import annotations

import torch
from torch import nn


class MultiCropWrapper(nn.Module):
    def __init__(self, backbone, head, *, stage_features: bool = True):
        """`stage_features` (plumbing only): when no autograd graph is being built (the frozen teacher, main_dino_mc.py:373) the per-group backbone
        outputs are written straight into ONE persistent `[n_crops * B, D]` buffer (`torch.cat(..., out=buffer)`) that the
        head reads -- no per-step allocation and a fixed address (what a captured step needs).  With autograd (the student)
        the single concatenation stays, because its backward scatters the feature gradient back to the groups."""
        super().__init__()
        # disable layers dedicated to ImageNet labels classification (utils/utils.py:622-623)
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head
        self.stage_features = stage_features
        self._stage = None

    @staticmethod
    def _groups(x):
        """[(start, end)] of maximal runs of crops with the same last-dim size (torch.unique_consecutive + cumsum in the
        reference, :631-634)."""
        bounds, start = [], 0
        for i in range(1, len(x) + 1):
            if i == len(x) or x[i].shape[-1] != x[start].shape[-1]:
                bounds.append((start, i))
                start = i
        return bounds

    def forward(self, x):
        if not isinstance(x, list):
            x = [x]
        outs = []
        for s, e in self._groups(x):
            out = self.backbone(torch.cat(x[s:e]))
            if isinstance(out, tuple):          # XCiT returns a tuple (:639-640)
                out = out[0]
            outs.append(out)
        if len(outs) == 1:
            feats = outs[0]
        elif self.stage_features and all(o.dim() == 2 for o in outs) and not (torch.is_grad_enabled() and any(o.requires_grad for o in outs)):
            rows, dim = sum(o.shape[0] for o in outs), outs[0].shape[1]
            st = self._stage
            if st is None or st.shape != (rows, dim) or st.dtype != outs[0].dtype or st.device != outs[0].device:
                st = self._stage = torch.empty((rows, dim), dtype=outs[0].dtype, device=outs[0].device)
            feats = torch.cat(outs, out=st)
        else:
            feats = torch.cat(outs)
        return self.head(feats)
