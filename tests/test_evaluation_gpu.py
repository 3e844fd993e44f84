"""SURVEY 8f N3, device part: the anti-causal regressor `GaussianConvEncoderClf` (ref nn.py:115-220) on the hand-written
encoder kernels against the REAL reference's outputs / gradients (tests/golden/evaluation_v1.npz), the effectiveness MAE
(ref scripts/image_causaldae_test.py:597-607) and latent collection + DCI end to end on the device."""
import os

import numpy as np
import pytest
import torch

from tests.golden import evaluation_cases as ec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", list(ec.CLF_CASES))
def test_regressor_matches_reference(name):
    from causaldiffae_b200.evaluation import GaussianConvEncoderClf, effectiveness_mae
    gold = np.load(os.path.join(ROOT, "tests", "golden", "evaluation_v1.npz"))
    case = ec.CLF_CASES[name]
    clf = GaussianConvEncoderClf(in_channels=case["in_channels"], latent_dim=512, num_vars=case["num_vars"])
    clf.load_state_dict(ec.clf_state_dict(clf.state_dict(), case["seed"]), strict=True)
    clf.cuda()
    x, target = ec.clf_inputs(case)
    clf.eval()
    with torch.no_grad():
        out = clf(x.cuda())
    np.testing.assert_allclose(out.cpu().numpy(), gold[f"clf/{name}/eval_out"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(float(effectiveness_mae(clf, x.cuda(), target)), gold[f"clf/{name}/mae"], rtol=2e-4)
    clf.train()
    o = clf(x.cuda())
    torch.nn.L1Loss()(o, target.cuda().unsqueeze(1)).backward()
    np.testing.assert_allclose(o.detach().cpu().numpy(), gold[f"clf/{name}/train_out"], rtol=2e-4, atol=2e-5)
    named = dict(clf.named_parameters())
    for pn in case["grad_probe"]:
        ref = gold[f"clf/{name}/grad/{pn}"]
        got = named[pn].grad.cpu().numpy()
        err = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30)
        assert err < 5e-4, (pn, err)
    np.testing.assert_allclose(clf.encoder[0][1].running_mean.cpu().numpy(), gold[f"clf/{name}/running_mean0"], rtol=1e-4, atol=1e-6)


def test_collect_latents_and_dci_on_device():
    """encode -> var := 0.001 -> DAG layer -> reparameterize over batches (ref :161-312), then DCI on the collected codes"""
    from causaldiffae_b200 import script_util as su, evaluation as ev
    from oracle import model as om
    flags = dict(image_size=32, num_channels=64, num_res_blocks=1, class_cond=False, rep_cond=True, n_vars=4, causal_modeling=True,
                 in_channels=3, learn_sigma=False)
    full = {**su.model_and_diffusion_defaults(), **flags}
    model, _ = su.create_model_and_diffusion(**full)
    cfg = om.config_from_flags(**full)
    sd = om.seeded_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    model.cuda().eval()
    g = torch.Generator().manual_seed(0)
    batches = [(torch.rand(16, 3, 32, 32, generator=g), {"c": torch.rand(16, 4, generator=g)}) for _ in range(6)]
    torch.manual_seed(1)
    rep, y = ev.collect_latents(model, batches)
    assert rep.shape == (96, 512) and y.shape == (96, 4)
    # the collected code is z_post + sqrt(0.001) xi: its mean over draws is the oracle's z_post
    x0 = batches[0][0]
    with torch.no_grad():
        mu, _ = om.encoder_encode(sd, cfg, x0, training=False)
        zp = om.nonlinearity_add_back_noise(sd, mu, om.causal_masking(mu, cfg.A, 4), 4)
    d = rep[:16] - zp.numpy()
    assert abs(d.std() - 0.001 ** 0.5) < 3e-3 and abs(d.mean()) < 3e-3
    np.random.seed(0)
    scores, imp, _ = ev.compute_dci(rep[:64, :32].T, y[:64].T, rep[64:, :32].T, y[64:].T)
    assert imp.shape == (32, 4) and 0.0 <= scores["disentanglement"] <= 1.0 and 0.0 <= scores["completeness"] <= 1.0
