"""The two callers of the hot path, mirrored from the reference's model wrapper so the parity tests read like its usage:

  * `optimize_parameters`  -- basicsr/models/twoImage_event_recurrent_model.py:273-310: zero_grad, `net_g(x=lq,
    event=voxel)`, Charbonnier loss (losses.py:28-30, eps 1e-12, mean), the `+ 0 * sum(p.sum())` term that gives the
    never-used parameters a zero gradient (:301), backward, `clip_grad_norm_(..., 0.01)` unless `use_grad_clip: false`
    (:304-306), optimizer step (AdamW / Adam from `train.optim_g`, :67-95) -- clip and step are one fused multi-tensor
    update (`refid_b200.optim`);
  * `test` -- :312-330: eval + no_grad, `val.max_minibatch` chunking, outputs concatenated on dim 0.

Data loading, validation bookkeeping, logging, checkpoints and schedulers stay with the reference (out of scope,
SURVEY.md 8); `feed_data` only moves the three tensors to the device (:97-104).
"""
from copy import deepcopy

import torch

from . import _lib
from . import grids
from . import losses as loss_module
from . import optim
from .plugin import define_network




class TwoImageEventRecurrentRestorationModel:
    def __init__(self, opt, device="cuda"):
        self.opt = opt
        self.device = torch.device(device)
        self.net_g = define_network(deepcopy(opt["network_g"])).to(self.device)
        self.is_train = "train" in opt
        self.optimizer_g = None
        if self.is_train:
            self.net_g.train()
            train_opt = deepcopy(opt["train"])
            pix = train_opt.get("pixel_opt") or {"type": "CharbonnierLoss"}
            if pix.get("type", "CharbonnierLoss") != "CharbonnierLoss":
                raise NotImplementedError("only CharbonnierLoss is used by the shipped option files")
            pix = dict(pix)
            self.cri_pix = getattr(loss_module, pix.pop("type", "CharbonnierLoss"))(**pix).to(self.device)  # :52-58
            og = train_opt.get("optim_g", {"type": "AdamW", "lr": 2e-4, "weight_decay": 1e-4, "betas": [0.9, 0.99]})
            og = dict(og)
            optim_type = og.pop("type")
            # the reference's two groups (:67-91): every parameter of this network lands in the first; the second
            # (`module.offsets` / `module.dcns` names, lr x 0.1) is empty but part of the optimizer's state_dict layout
            optim_params, optim_params_lowlr = [], []
            for k, v in self.net_g.named_parameters():
                if v.requires_grad:
                    (optim_params_lowlr if k.startswith(("module.offsets", "module.dcns")) else optim_params).append(v)
            groups = [{"params": optim_params}, {"params": optim_params_lowlr, "lr": og["lr"] * 0.1}]
            if optim_type == "AdamW":
                self.optimizer_g = optim.ClipAdamW(groups, **og)
            elif optim_type == "Adam":
                self.optimizer_g = optim.ClipAdam(groups, **og)
            else:
                raise NotImplementedError(f"optimizer {optim_type} is not supperted yet.")
        self.log_dict = {}

    def feed_data(self, data):
        self.lq = data["lq"].to(self.device)
        self.voxel = data["voxel"].to(self.device)
        if "gt" in data:
            self.gt = data["gt"].to(self.device)

    def optimize_parameters(self, current_iter=0):
        self.optimizer_g.zero_grad()
        pred = self.net_g(x=self.lq, event=self.voxel)
        l_pix = self.cri_pix(pred, self.gt)  # :284
        l_total = l_pix + 0 * sum(p.sum() for p in self.net_g.parameters())
        l_total.backward()
        if self.opt["train"].get("use_grad_clip", True):
            self.optimizer_g.clip_grad_norm_(0.01)  # :304-306, fused into the step below
        self.optimizer_g.step()
        self.log_dict = {"l_pix": l_pix.detach()}
        _lib.raise_if_aborted()  # a kernel that hit its bounded wait invalidates everything after it: fail loudly
        return l_pix.detach()

    # ---- full-frame validation tiling (reference :128-270; gated by `val.grids` in nondist_validation, :405-411) ----
    def grids(self):
        """`lq` (1, ..., H, W) -> overlapping crops on dim 0 (reference `grids`, :201-243)."""
        val = self.opt["val"]
        h, w = self.lq.shape[-2:]
        self.original_size = tuple(self.lq.shape)
        self.idxes = grids.crop_positions(h, w, val["crop_size"], val.get("trans_num", 1), val.get("random_crop_num", 0))
        self.origin_lq = self.lq
        self.lq = grids.crop(self.lq, self.idxes, val["crop_size"])

    def grids_voxel(self):
        """`voxel` (1, T, C, H, W) -> the same crops as `grids()` (reference `grids_voxel`, :128-199; the reference draws
        its random crops independently for lq and voxel -- here one placement list serves both)."""
        self.original_size_voxel = tuple(self.voxel.shape)
        self.origin_voxel = self.voxel
        self.voxel = grids.crop(self.voxel, self.idxes, self.opt["val"]["crop_size"])

    def grids_inverse(self):
        """Network outputs of the crops -> one (1, T, out_chn, H, W) frame, overlap-averaged (reference :245-270)."""
        h, w = self.original_size[-2:]
        self.output = grids.merge(self.output, self.idxes, h, w)
        self.lq = self.origin_lq
        self.voxel = self.origin_voxel

    def validate_frame(self, data):
        """One frame of `nondist_validation` (:395-417): feed, tile if `val.grids`, test, merge."""
        self.feed_data(data)
        tiled = (self.opt.get("val") or {}).get("grids") is not None
        if tiled:
            self.grids()
            self.grids_voxel()
        self.test()
        if tiled:
            self.grids_inverse()
        return self.output

    def test(self):
        self.net_g.eval()
        with torch.no_grad():
            n = self.lq.size(0)
            m = (self.opt.get("val") or {}).get("max_minibatch", n)
            outs, i = [], 0
            while i < n:
                j = min(i + m, n)
                outs.append(self.net_g(x=self.lq[i:j], event=self.voxel[i:j]))
                i = j
            self.output = torch.cat(outs, dim=0)
        _lib.raise_if_aborted()
        self.net_g.train()
        return self.output
