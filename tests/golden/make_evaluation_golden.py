"""Generates tests/golden/evaluation_v1.npz from the REAL reference (read-only tree /root/reference, build container only):
  * DCI (improved_diffusion/metrics.py `_compute_dci`, `disentanglement`, `completeness`) on seeded synthetic codes / factors,
  * the anti-causal regressor `GaussianConvEncoderClf` (improved_diffusion/nn.py:115-220): outputs and one SGD-free backward
    (parameter gradients of an L1 loss) on seeded weights, for the Pendulum (4 x 96 x 96, 6 conv stages) and MorphoMNIST
    (1 x 28 x 28, 4 stages) geometries the reference scripts use.
Run:  python tests/golden/make_evaluation_golden.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests.golden import evaluation_cases as ec  # noqa: E402


def main():
    ns = refshim.load()
    sys.path.insert(0, refshim.REF_ROOT)
    sys.modules.setdefault("munkres", importlib.import_module("improved_diffusion.munkres"))
    mt = importlib.import_module("improved_diffusion.metrics")
    out = {}
    for name, case in ec.DCI_CASES.items():
        xtr, ytr, xte, yte = ec.dci_inputs(case)
        np.random.seed(case["seed"] + 1)
        scores, imp, code_imp = mt._compute_dci(xtr, ytr, xte, yte)
        for k, v in scores.items():
            out[f"dci/{name}/{k}"] = np.float64(v)
        out[f"dci/{name}/importance"] = imp
        out[f"dci/{name}/code_importance"] = code_imp
    imp = ec.fixed_importance()
    out["dci/fixed/disentanglement"] = np.float64(mt.disentanglement(imp)[0])
    out["dci/fixed/completeness"] = np.float64(mt.completeness(imp))
    out["dci/fixed/per_code"] = mt.disentanglement_per_code(imp)
    out["dci/fixed/per_factor"] = mt.completeness_per_factor(imp)
    for name, case in ec.CLF_CASES.items():
        clf = ns.nn.GaussianConvEncoderClf(in_channels=case["in_channels"], latent_dim=512, num_vars=case["num_vars"])
        clf.load_state_dict(ec.clf_state_dict(clf.state_dict(), case["seed"]), strict=True)
        x, target = ec.clf_inputs(case)
        clf.eval()
        with torch.no_grad():
            out[f"clf/{name}/eval_out"] = clf(x).numpy()
            out[f"clf/{name}/mae"] = np.float64(torch.nn.L1Loss()(clf(x), target.unsqueeze(1)))
        clf.train()
        o = clf(x)
        torch.nn.L1Loss()(o, target.unsqueeze(1)).backward()
        out[f"clf/{name}/train_out"] = o.detach().numpy()
        for pn in case["grad_probe"]:
            out[f"clf/{name}/grad/{pn}"] = dict(clf.named_parameters())[pn].grad.numpy().copy()
        out[f"clf/{name}/running_mean0"] = clf.encoder[0][1].running_mean.numpy().copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "evaluation_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
