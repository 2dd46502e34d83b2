"""Fused gradient clip + AdamW over flat buffers: the optimizer step of src/engine_glassrgbd.py:155-159
(`clip_grad_norm_(model.parameters(), max_norm)` then `optimizer.step()`) with the parameter groups of
src/main_glassrgbd.py:59-67 (AdamW; backbone parameters at lr_backbone, everything else at lr; one weight decay).

B200 design: every trainable tensor of a group is re-homed into ONE flat fp32 buffer per group (parameters and their
`.grad` become views), so the whole step is 1 + G launches whatever the number of tensors (the reference model has ~900):
`gwd_sumsq` per group accumulates the global squared norm in fp64 on the device, `gwd_adamw_step` per group reads the clip
coefficient from device memory (no host sync, unlike `clip_grad_norm_`'s `.item()`-free but 2x900-kernel foreach path),
applies the update with 16-byte vectors and can refresh a bf16 mirror of the weights in the same pass.  Data-parallel
training all-reduces the flat gradient buffers (one NCCL call per group) instead of per-tensor buckets.
"""
import torch

from . import ops, parallel


class FlatAdamW(torch.optim.Optimizer):
    """Drop-in for `torch.optim.AdamW(param_groups, lr, weight_decay)` + `clip_grad_norm_` on CUDA parameters.  A
    `torch.optim.Optimizer`, so `StepLR(optimizer, lr_drop)` of src/main_glassrgbd.py:67 works on it, with
    `state_dict()` / `load_state_dict()` for the resume path (:160-163, :222).

    param_groups: list of dicts {"params": [...], "lr": optional} exactly as src/main_glassrgbd.py:59-66 builds them.
    After construction every parameter's storage is a view into the group's flat buffer and `p.grad` is a persistent
    view into the flat gradient buffer (`zero_grad` zeroes the buffers; autograd accumulates into the views)."""

    def __init__(self, param_groups, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=0.0, bf16_mirror=False):
        if isinstance(param_groups, (list, tuple)) and param_groups and not isinstance(param_groups[0], dict):
            param_groups = [{"params": list(param_groups)}]
        plain = [dict(g, params=list(g["params"])) for g in param_groups]
        torch.optim.Optimizer.__init__(self, plain, dict(lr=lr, weight_decay=weight_decay))
        self.betas, self.eps, self.max_norm, self.t = betas, eps, max_norm, 0
        self.groups = []
        for g in param_groups:
            params = [p for p in g["params"] if p.requires_grad]
            if not params:
                continue
            dev = params[0].device
            if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in params):
                raise RuntimeError("FlatAdamW runs on libgwd_b200 CUDA kernels: fp32 CUDA parameters on one device only")
            sizes = [ops.round_up(p.numel(), 4) for p in params]       # 16-byte aligned segments
            n = sum(sizes)
            P = torch.zeros(n, dtype=torch.float32, device=dev)
            G = torch.zeros_like(P)
            off = 0
            for p, sz in zip(params, sizes):
                P[off:off + p.numel()].copy_(p.detach().reshape(-1))
                p.data = P[off:off + p.numel()].view(p.shape)
                p.grad = G[off:off + p.numel()].view(p.shape)
                off += sz
            self.groups.append({"params": params, "P": P, "G": G, "M": torch.zeros_like(P), "V": torch.zeros_like(P),
                                "mirror": torch.empty(n, dtype=torch.bfloat16, device=dev) if bf16_mirror else None,
                                "lr": g.get("lr", lr), "weight_decay": g.get("weight_decay", weight_decay)})
        self.param_groups = self.groups          # the reference's lr scheduler edits param_groups[i]["lr"]
        for g in self.groups:
            g.setdefault("initial_lr", g["lr"])
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.groups[0]["P"].device) if self.groups else None

    def state_dict(self):
        """step count, learning rates and the flat Adam moments per group (the parameter order is the construction order)"""
        return {"format": "gwd_flat_adamw_v1", "t": self.t,
                "groups": [{"lr": g["lr"], "initial_lr": g.get("initial_lr", g["lr"]), "weight_decay": g["weight_decay"],
                            "numel": g["P"].numel(), "M": g["M"].detach().cpu(), "V": g["V"].detach().cpu()} for g in self.groups]}

    def load_state_dict(self, state):
        assert state.get("format") == "gwd_flat_adamw_v1" and len(state["groups"]) == len(self.groups), "optimizer state of another layout"
        self.t = int(state["t"])
        for g, s in zip(self.groups, state["groups"]):
            assert s["numel"] == g["P"].numel()
            g["M"].copy_(s["M"])
            g["V"].copy_(s["V"])
            g["lr"], g["initial_lr"], g["weight_decay"] = s["lr"], s["initial_lr"], s["weight_decay"]

    def zero_grad(self, set_to_none=False):
        for g in self.groups:
            g["G"].zero_()
            for p in g["params"]:           # autograd may have replaced a view (e.g. first backward with grad=None)
                if p.grad is None or p.grad.data_ptr() < g["G"].data_ptr() or p.grad.data_ptr() >= g["G"].data_ptr() + g["G"].numel() * 4:
                    self._rebind(g)
                    break

    def _rebind(self, g):
        off = 0
        for p in g["params"]:
            p.grad = g["G"][off:off + p.numel()].view(p.shape)
            off += ops.round_up(p.numel(), 4)

    @torch.no_grad()
    def step(self, closure=None):
        """all-reduce (if distributed) + global-norm clip + AdamW; returns the device tensor holding sum g^2 (before the
        1/world scale), so the caller can log the gradient norm without forcing a sync here"""
        world = 1
        for g in self.groups:
            world = parallel.allreduce_sum_(g["G"])
        self.t += 1
        self._sumsq.zero_()
        if self.max_norm > 0:
            for g in self.groups:
                ops.sumsq(g["G"], self._sumsq)
        for g in self.groups:
            ops.adamw_step(g["P"], g["G"], g["M"], g["V"], g["mirror"], lr=g["lr"], betas=self.betas, eps=self.eps,
                           weight_decay=g["weight_decay"], step=self.t, max_norm=self.max_norm, grad_scale=1.0 / world,
                           sumsq_buf=self._sumsq if self.max_norm > 0 else None)
        return self._sumsq
