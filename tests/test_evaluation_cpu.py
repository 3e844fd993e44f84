"""SURVEY 8f N3, host part: DCI (ref improved_diffusion/metrics.py:167-232) restated in causaldiffae_b200/evaluation.py against
the REAL reference's numbers on seeded codes / factors (tests/golden/evaluation_v1.npz, make_evaluation_golden.py)."""
import os

import numpy as np
import pytest

from tests.golden import evaluation_cases as ec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "evaluation_v1.npz"))


def test_dci_entropy_scores_match_reference(gold):
    from causaldiffae_b200 import evaluation as ev
    imp = ec.fixed_importance()
    np.testing.assert_allclose(ev.disentanglement_per_code(imp), gold["dci/fixed/per_code"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ev.completeness_per_factor(imp), gold["dci/fixed/per_factor"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ev.disentanglement(imp)[0], gold["dci/fixed/disentanglement"], rtol=1e-12)
    np.testing.assert_allclose(ev.completeness(imp), gold["dci/fixed/completeness"], rtol=1e-12)
    z = np.zeros((5, 3))                      # all-zero importance: the reference substitutes ones for the weights
    assert np.isfinite(ev.disentanglement(z)[0]) and np.isfinite(ev.completeness(z))


@pytest.mark.parametrize("name", list(ec.DCI_CASES))
def test_compute_dci_matches_reference(gold, name):
    from causaldiffae_b200 import evaluation as ev
    case = ec.DCI_CASES[name]
    xtr, ytr, xte, yte = ec.dci_inputs(case)
    np.random.seed(case["seed"] + 1)          # sklearn's GradientBoostingRegressor(random_state=None) draws from numpy's global state
    scores, imp, code_imp = ev.compute_dci(xtr, ytr, xte, yte)
    np.testing.assert_allclose(imp, gold[f"dci/{name}/importance"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(code_imp, gold[f"dci/{name}/code_importance"], rtol=1e-9, atol=1e-12)
    for k in ("informativeness_train", "informativeness_test", "disentanglement", "completeness"):
        np.testing.assert_allclose(scores[k], gold[f"dci/{name}/{k}"], rtol=1e-9, atol=1e-12, err_msg=k)
    assert ev._compute_dci is ev.compute_dci
