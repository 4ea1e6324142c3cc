"""Checkpoint format fidelity and resume.

`trained_model.pth` is a plain `model.state_dict()` (main.py:53) with the 80 keys of SURVEY A.2; evaluate.py:48-56 loads it
with `strict=False` into a FRESH model, whose lazily created `vertex_predictor.point_pool_proj` does not exist yet -- so the
reference silently drops the trained projection.  `load_model_state` creates the projection first when the file
has it; `save_training_state` / `load_training_state` add optimizer state and the step counter for data-parallel restarts
(every rank loads the same file, so replicas stay identical; rank 0 writes)."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

_PROJ = "vertex_predictor.point_pool_proj"


def load_model_state(model: torch.nn.Module, state_dict_or_path, strict: bool = False):
    """model.load_state_dict(...) that keeps the checkpoint's `point_pool_proj` (the reference loses it on a fresh model)."""
    sd = state_dict_or_path
    if not isinstance(sd, dict):
        sd = torch.load(sd, map_location="cpu")
    if "model" in sd and isinstance(sd["model"], dict) and "vertex_predictor.final_layer.weight" in sd["model"]:
        sd = sd["model"]
    w = sd.get(_PROJ + ".weight")
    vp = model.vertex_predictor
    if w is not None and not hasattr(vp, "point_pool_proj"):
        proj = torch.nn.Linear(w.shape[1], w.shape[0])
        ref = next(vp.parameters())
        vp.point_pool_proj = proj.to(device=ref.device, dtype=ref.dtype)
    return model.load_state_dict(sd, strict=strict)


def save_training_state(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, step: int = 0,
                        extra: Optional[dict] = None) -> None:
    """Rank 0 writes {'model': state_dict (the reference's 80 keys), 'optimizer': ..., 'step': ...} atomically; other ranks wait."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == 0:
        blob = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "step": int(step)}
        if optimizer is not None:
            blob["optimizer"] = optimizer.state_dict()
        if extra:
            blob["extra"] = extra
        tmp = path + ".tmp"
        torch.save(blob, tmp)
        os.replace(tmp, path)
    if dist.is_initialized():
        dist.barrier()


def load_training_state(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None) -> int:
    """Restores model (+ optimizer) on every rank; returns the saved step.  The optimizer must have been built AFTER the lazy
    projection exists (build it after load_model_state or after the first forward), otherwise its state has no slot for it."""
    blob = torch.load(path, map_location="cpu")
    load_model_state(model, blob["model"], strict=False)
    if optimizer is not None and "optimizer" in blob:
        optimizer.load_state_dict(blob["optimizer"])
    return int(blob.get("step", 0))


def export_reference_checkpoint(path: str, model: torch.nn.Module) -> None:
    """Exactly what main.py:53 writes: torch.save(model.state_dict(), 'trained_model.pth')."""
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
