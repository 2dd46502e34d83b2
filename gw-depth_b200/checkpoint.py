"""Checkpoint tooling of the training entry point (src/main_glassrgbd.py:104-193, 214-226) for the drop-in module: the key
surgery the reference applies before `load_state_dict(strict=False)`, and saving in the reference's layout.

    load_detr_pretrained(model, ckpt)     a DETR-R50 checkpoint (the `--resume https://...detr-r50...` path, :107-127): everything
                                          but class_embed / bbox_embed / query_embed (and input_proj when layer1_num != 3)
    load_resume(model, ckpt, ...)         a GW-Depth / LETR checkpoint (:129-164): `module.` prefixes of DataParallel /
                                          DistributedDataParallel stripped, `bbox_embed.*` of the old implementation renamed to
                                          `lines_embed.*`; optionally the optimizer / lr-scheduler / epoch
    load_frozen_letr(model, ckpt)         `--frozen_weights` (:165-189): the same surgery, then only encoder / decoder /
                                          class_embed / lines_embed tensors
    save_checkpoint(path, model, ...)     {'model', 'optimizer', 'lr_scheduler', 'epoch', 'args'} (:214-226)

Every loader returns a report {loaded, missing (in the model, not in the file), unexpected (in the file, not in the model),
shape_mismatch} instead of printing, and never touches tensors whose shape disagrees (the reference would raise there).
The module's kernel plan / training engine are invalidated so the next forward sees the new weights."""
import re

import torch


def _clean_key(k):
    """:132-142 -- `re.sub('module.', '', k)` when 'module' occurs in the key (the dot of the reference's pattern is a wildcard;
    with DataParallel prefixes it only ever matches the literal 'module.')"""
    return re.sub(re.compile("module."), "", k) if re.search("module", k) else k


def remap_resume_keys(state):
    """the key surgery of src/main_glassrgbd.py:131-142 on a {'key': tensor} mapping"""
    out = {}
    for k, v in state.items():
        if "bbox_embed" in k:          # old implementation: bbox_embed.layers.N.* -> lines_embed.layers.N.*
            out["lines_embed." + ".".join(k.split(".")[1:])] = v
        else:
            out[_clean_key(k)] = v
    return out


def filter_detr_keys(state, layer1_num=3):
    """the key filter of src/main_glassrgbd.py:109-115 for DETR-R50 weights"""
    out = {}
    for k, v in state.items():
        if ("class_embed" in k) or ("bbox_embed" in k) or ("query_embed" in k):
            continue
        if ("input_proj" in k) and layer1_num != 3:
            continue
        out[k] = v
    return out


def _apply(model, new_state):
    own = model.state_dict()
    report = {"loaded": [], "missing": [k for k in own if k not in new_state], "unexpected": [], "shape_mismatch": []}
    usable = {}
    for k, v in new_state.items():
        if k not in own:
            report["unexpected"].append(k)
        elif tuple(own[k].shape) != tuple(v.shape):
            report["shape_mismatch"].append(k)
        else:
            usable[k] = v
            report["loaded"].append(k)
    model.load_state_dict(usable, strict=False)
    if hasattr(model, "_plan"):          # the cached kernel plan / training engine hold re-laid-out copies of the weights
        model._plan = None
    if isinstance(getattr(model, "__dict__", None), dict) and model.__dict__.get("_trainer") is not None:
        model.__dict__["_trainer"].load_params(model.state_dict())
        model.__dict__["_trainer_versions"] = model._param_versions()
    return report


def _model_state(ckpt):
    return ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt


def load_detr_pretrained(model, ckpt, layer1_num=3):
    return _apply(model, filter_detr_keys(_model_state(ckpt), layer1_num))


def load_resume(model, ckpt, optimizer=None, lr_scheduler=None, lr_drop=None, no_opt=False, evaluate=False):
    """-> (report, start_epoch or None).  The optimizer / scheduler are restored exactly when the reference does (:159-163)."""
    report = _apply(model, remap_resume_keys(_model_state(ckpt)))
    start_epoch = None
    if (optimizer is not None and lr_scheduler is not None and not no_opt and not evaluate and isinstance(ckpt, dict)
            and "optimizer" in ckpt and "lr_scheduler" in ckpt and "epoch" in ckpt):
        optimizer.load_state_dict(ckpt["optimizer"])
        sched = dict(ckpt["lr_scheduler"])
        if lr_drop is not None:
            sched["step_size"] = lr_drop          # "change the lr_drop epoch" (:161)
        lr_scheduler.load_state_dict(sched)
        start_epoch = ckpt["epoch"] + 1
    return report, start_epoch


def load_frozen_letr(model, ckpt):
    state = remap_resume_keys(_model_state(ckpt))
    keep = {k: v for k, v in state.items() if any(t in k for t in ("encoder", "decoder", "class_embed", "lines_embed"))}
    return _apply(model, keep)


def save_checkpoint(path, model, optimizer=None, lr_scheduler=None, epoch=0, args=None):
    """the dictionary `save_on_master` writes every epoch (:214-226); the fused training engine's parameters are synchronised
    into the module first"""
    if hasattr(model, "sync_from_trainer"):
        model.sync_from_trainer()
    blob = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "epoch": epoch, "args": args}
    if optimizer is not None:
        blob["optimizer"] = optimizer.state_dict()
    if lr_scheduler is not None:
        blob["lr_scheduler"] = lr_scheduler.state_dict()
    torch.save(blob, path)
    return blob
