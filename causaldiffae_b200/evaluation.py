"""The evaluation step AFTER the hot path: what the reference does with trained models in
scripts/image_causaldae_test.py.

  * effectiveness of counterfactuals (ref scripts/image_causaldae_test.py:140-159,597-607): per-factor anti-causal REGRESSORS
    `GaussianConvEncoderClf` (ref nn.py:115-220: the conv encoder stack + `fc: Linear(hidden*4, 1)`) predict the factor from a
    generated image; the score is the mean absolute error (nn.L1Loss) against the intervened value.  The conv stack runs on
    the same hand-written kernels as the model's encoder (rep.EncoderRunner), forward and backward (so the regressors can be
    trained here as well, ref *_classifier.py).
  * disentanglement of the latent code (ref :161-312): z = reparameterize(z_post(encode(x)), 0.001) collected over a dataset,
    then DCI (ref metrics.py:167-232): a gradient-boosted regressor per factor (scikit-learn, host), importance matrix,
    disentanglement / completeness as entropy-weighted sums.  The DCI arithmetic is host numpy like the reference's - it is
    restated here so that the numbers are the reference's (tests/golden/evaluation_v1.npz pins them against the reference's
    own metrics.py on seeded codes)."""
import numpy as np
import torch as th
import torch.nn as nn

from . import ops


class GaussianConvEncoderClf(nn.Module):
    """ref nn.py:115-220.  state_dict keys: encoder.{k}.{0,1}.*, fc_mu.*, fc_var.*, fc.*  (the ConvTranspose blocks the
    reference appends to a local list after building `encoder` are never registered, so they are not part of the format)."""

    def __init__(self, in_channels, latent_dim, hidden_dims=None, num_vars=4, **kwargs):
        super().__init__()
        self.latent_dim, self.in_channels, self.num_vars = latent_dim, in_channels, num_vars
        if hidden_dims is None:
            hidden_dims = [16, 32, 32, 64, 64, 128] if num_vars == 4 else [16, 32, 64, 128]
        mods, cin = [], in_channels
        for h in hidden_dims:
            mods.append(nn.Sequential(nn.Conv2d(cin, h, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(h), nn.LeakyReLU()))
            cin = h
        self.encoder = nn.Sequential(*mods)
        self.fc_mu = nn.Linear(hidden_dims[-1] * 4, latent_dim)
        self.fc_var = nn.Linear(hidden_dims[-1] * 4, latent_dim)
        self.fc = nn.Linear(hidden_dims[-1] * 4, 1)

    @property
    def runner(self):
        from .rep import EncoderRunner
        r = self.__dict__.get("_runner")
        if r is None:
            r = self.__dict__["_runner"] = EncoderRunner(self)
        return r

    def encode(self, input):
        from .rep import _EncodeFn, anchor
        mu, var = _EncodeFn.apply(self, anchor(input.device), input, th.is_grad_enabled())
        return [mu, var]

    def forward(self, x):
        """ref nn.py:212-218: fc(flatten(encoder(x))) -> [B, 1]"""
        from . import _lib
        from .rep import anchor
        if not x.is_cuda:
            raise _lib.CdaeError("GaussianConvEncoderClf needs CUDA (sm_100a) tensors; there is no CPU path")
        return _ClfFn.apply(self, anchor(x.device), x, th.is_grad_enabled())


class _ClfFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, clf, _anchor, x, want_bwd):
        run = clf.runner
        xc = x.detach().float().contiguous()
        if want_bwd and not clf.training:
            raise NotImplementedError("gradients through the conv stack in eval mode (running BatchNorm statistics) are not built")
        st = run.alloc(xc.shape[0], xc.shape[1], xc.shape[2], xc.shape[3], xc.device, want_bwd)
        h = run.features(st, xc, clf.training)
        out = th.empty(xc.shape[0], 1, device=xc.device, dtype=th.float32)
        ops.linear_fwd(h, clf.fc.weight, clf.fc.bias, out)
        ctx.clf, ctx.st, ctx.x = clf, st, xc
        return out

    @staticmethod
    def backward(ctx, dout):
        from .rep import _g
        clf, st = ctx.clf, ctx.st
        ops.linear_bwd(st.hfeat, clf.fc.weight, dout.float().contiguous(), _g(clf.fc.weight), _g(clf.fc.bias), dx=st.dh)
        clf.runner.features_backward(st, ctx.x)
        return None, None, None, None


@th.no_grad()
def effectiveness_mae(clf, images, target):
    """ref scripts/image_causaldae_test.py:597-607: nn.L1Loss()(clf(sample), value.unsqueeze(1)) for one factor"""
    out = clf(images)
    tgt = th.as_tensor(target, dtype=th.float32, device=out.device).reshape(-1, 1).expand_as(out)
    return (out - tgt).abs().mean()


@th.no_grad()
def collect_latents(model, batches, A=None, device=None):
    """ref scripts/image_causaldae_test.py:161-312: z = reparameterize(z_post(encode(x)), var := 0.001) over (batch, cond)
    pairs -> (codes [N, rep_dim], factors [N, n_vars]) numpy arrays (the layout `_compute_dci` takes transposed)."""
    from .sampling import encode
    reps, ys = [], []
    dev = device if device is not None else next(model.parameters()).device
    for batch, cond in batches:
        z, _, _ = encode(model, batch.to(dev), A=A)
        reps.append(z.reshape(z.shape[0], -1).cpu().numpy())
        ys.append(cond["c"].detach().cpu().numpy())
    return np.concatenate(reps, axis=0), np.concatenate(ys, axis=0)


# ---------------------------------------------------------------------------------------------------- DCI (ref metrics.py:167-232)
def compute_importance_gbt(x_train, y_train, x_test, y_test):
    """ref metrics.py:183-201: one sklearn GradientBoostingRegressor per factor; importance = |feature_importances_|; the two
    'informativeness' numbers are the reference's (mean of exact equality between prediction and target)."""
    from sklearn import ensemble
    num_factors, num_codes = y_train.shape[0], x_train.shape[0]
    importance = np.zeros((num_codes, num_factors), dtype=np.float64)
    train_loss, test_loss = [], []
    for i in range(num_factors):
        model = ensemble.GradientBoostingRegressor()
        model.fit(x_train.T, y_train[i, :])
        importance[:, i] = np.abs(model.feature_importances_)
        train_loss.append(np.mean(model.predict(x_train.T) == y_train[i, :]))
        test_loss.append(np.mean(model.predict(x_test.T) == y_test[i, :]))
    return importance, np.mean(train_loss), np.mean(test_loss)


def _entropy(p, base):
    """scipy.stats.entropy(p, base=base) along axis 0 (columns are normalised to sum 1 first)"""
    p = np.asarray(p, dtype=np.float64)
    p = p / p.sum(axis=0, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, p * np.log(p), 0.0)
    return -t.sum(axis=0) / np.log(base)


def disentanglement_per_code(importance):
    """ref metrics.py:204-208"""
    return 1.0 - _entropy(importance.T + 1e-11, importance.shape[1])


def disentanglement(importance):
    """ref metrics.py:211-218 -> (score, code_importance)"""
    per_code = disentanglement_per_code(importance)
    if importance.sum() == 0.0:
        importance = np.ones_like(importance)
    code_importance = importance.sum(axis=1) / importance.sum()
    return np.sum(per_code * code_importance), code_importance


def completeness_per_factor(importance):
    """ref metrics.py:221-225"""
    return 1.0 - _entropy(importance + 1e-11, importance.shape[0])


def completeness(importance):
    """ref metrics.py:228-234"""
    per_factor = completeness_per_factor(importance)
    if importance.sum() == 0.0:
        importance = np.ones_like(importance)
    factor_importance = importance.sum(axis=0) / importance.sum()
    return np.sum(per_factor * factor_importance)


def compute_dci(mus_train, ys_train, mus_test, ys_test):
    """ref metrics.py:167-180 `_compute_dci`: codes [num_codes, N], factors [num_factors, N]"""
    importance, train_err, test_err = compute_importance_gbt(mus_train, ys_train, mus_test, ys_test)
    assert importance.shape == (mus_train.shape[0], ys_train.shape[0])
    scores = {"informativeness_train": train_err, "informativeness_test": test_err}
    disent, code_importance = disentanglement(importance)
    scores["disentanglement"] = disent
    scores["completeness"] = completeness(importance)
    return scores, importance, code_importance


_compute_dci = compute_dci      # the reference's name
