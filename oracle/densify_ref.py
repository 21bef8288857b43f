"""ORACLE — test infrastructure only.  Restates the reference's optimiser surgery for densification / pruning
(scene/gaussian_model.py:1044-1069 `_prune_optimizer`, :1094-1123 `cat_tensors_to_optimizer`) as free functions
over a torch.optim.Adam with one named single-parameter group per tensor, statement for statement.  Pure data
movement (boolean-mask index, torch.cat), so parity is bit-exact.  Pinned against the reference's own methods
executed on a real GaussianModel in tests/test_densify.py (when /root/reference is present)."""
import torch
import torch.nn as nn


def prune_optimizer(optimizer, mask):
    optimizable_tensors = {}
    for group in optimizer.param_groups:
        if len(group["params"]) > 1 or group["name"] == "focal":
            continue
        stored_state = optimizer.state.get(group["params"][0], None)
        if stored_state is not None:
            stored_state["exp_avg"] = stored_state["exp_avg"][mask]
            stored_state["exp_avg_sq"] = stored_state["exp_avg_sq"][mask]
            del optimizer.state[group["params"][0]]
            group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            optimizer.state[group["params"][0]] = stored_state
            optimizable_tensors[group["name"]] = group["params"][0]
        elif group["name"] == "current_control_num":
            group["params"][0] = nn.Parameter(group["params"][0][mask], requires_grad=False)
            optimizable_tensors[group["name"]] = group["params"][0]
        else:
            group["params"][0] = nn.Parameter(group["params"][0][mask].requires_grad_(True))
            optimizable_tensors[group["name"]] = group["params"][0]
    return optimizable_tensors


def cat_tensors_to_optimizer(optimizer, tensors_dict):
    optimizable_tensors = {}
    for group in optimizer.param_groups:
        if len(group["params"]) > 1 or group["name"] == "focal":
            continue
        extension_tensor = tensors_dict[group["name"]]
        stored_state = optimizer.state.get(group["params"][0], None)
        if stored_state is not None:
            stored_state["exp_avg"] = torch.cat((stored_state["exp_avg"], torch.zeros_like(extension_tensor)), dim=0)
            stored_state["exp_avg_sq"] = torch.cat((stored_state["exp_avg_sq"], torch.zeros_like(extension_tensor)), dim=0)
            del optimizer.state[group["params"][0]]
            group["params"][0] = nn.Parameter(torch.cat((group["params"][0], extension_tensor), dim=0).requires_grad_(True))
            optimizer.state[group["params"][0]] = stored_state
            optimizable_tensors[group["name"]] = group["params"][0]
        elif group["name"] == "current_control_num":
            group["params"][0] = nn.Parameter(torch.cat((group["params"][0], extension_tensor), dim=0), requires_grad=False)
            optimizable_tensors[group["name"]] = group["params"][0]
        else:
            group["params"][0] = nn.Parameter(torch.cat((group["params"][0], extension_tensor), dim=0).requires_grad_(True))
            optimizable_tensors[group["name"]] = group["params"][0]
    return optimizable_tensors
