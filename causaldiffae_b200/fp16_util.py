"""Mixed-precision helpers (ref improved_diffusion/fp16_util.py), same names.

The engine keeps fp32 master weights in one flat arena and computes the torso on bf16 tensor cores with fp32
accumulation, so there is no fp16 torso and no loss scaling: `convert_module_to_f16/f32` are recorded no-ops and the
flat master parameter of `make_master_params` IS the arena when the model has an engine (no copy)."""
import torch.nn as nn
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors

from . import ops


def convert_module_to_f16(l):
    """ref fp16_util.py:9-15 — bf16 operand copies are produced by the engine's pack kernel instead."""
    return l


def convert_module_to_f32(l):
    return l


def make_master_params(model_params):
    """ref fp16_util.py:27-37"""
    model_params = list(model_params)
    master = nn.Parameter(_flatten_dense_tensors([p.detach().float() for p in model_params]))
    master.requires_grad = True
    return [master]


def model_grads_to_master_grads(model_params, master_params):
    """ref fp16_util.py:40-47"""
    master_params[0].grad = _flatten_dense_tensors([p.grad.data.detach().float() for p in model_params])


def master_params_to_model_params(model_params, master_params):
    """ref fp16_util.py:50-62"""
    model_params = list(model_params)
    for p, mp in zip(model_params, unflatten_master_params(model_params, master_params)):
        p.detach().copy_(mp)


def unflatten_master_params(model_params, master_params):
    """ref fp16_util.py:65-68"""
    return _unflatten_dense_tensors(master_params[0].detach(), list(model_params))


def zero_grad(model_params):
    """ref fp16_util.py:71-76: in-place zeroing (gradients are views of the flat gradient arena; never set to None)."""
    for p in model_params:
        if p.grad is not None:
            p.grad.detach_()
            p.grad.zero_()
