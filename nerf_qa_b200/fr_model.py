"""``NeRFQAModel`` -- the full-reference "NeRF-DISTS" model of the reference (nerf_qa/model_stats.py:23-102, the FR half
of nerf_qa/model.py:22-56): a DISTS flavour whose alpha/beta are fine-tuned, followed by a 2- or 4-parameter regression
head (linear / sqrt / logistic) initialised from a least-squares fit of stored DISTS scores against subjective scores,
plus the entropy regulariser the training loop adds to the loss (run_final.py:194).

This is the caller of the hot path on the FR training path (run_final.py:168-229); everything here is host-side or
a handful of torch ops on scalars -- the scoring itself goes through ``DISTSOriginal`` / ``DISTSSoftmax`` and the C-ABI
library.  The reference reads its knobs from the global ``wandb.config``; here they are constructor arguments with the
same names (a live ``wandb.config`` is used for anything left unspecified, then the reference's argparse defaults).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from .dists import DISTSOriginal, DISTSSoftmax, _FlavourConfig

_DEFAULTS = {"regression_type": "linear", "subjective_score_type": "MOS", "dists_weight_norm": "off"}


class NeRFQAModel(nn.Module):
    def __init__(self, train_df=None, dists: Optional[Sequence[float]] = None, scores: Optional[Sequence[float]] = None,
                 regression_type: Optional[str] = None, subjective_score_type: Optional[str] = None,
                 dists_weight_norm: Optional[str] = None, precision: str = "bf16", **flavour_kwargs):
        """``train_df``: a DataFrame with a ``DISTS`` column and the subjective-score column (model_stats.py:28-31), or
        pass the two arrays as ``dists`` / ``scores``."""
        super().__init__()
        cfg = _FlavourConfig(regression_type=regression_type, subjective_score_type=subjective_score_type,
                             dists_weight_norm=dists_weight_norm)
        cfg.DEFAULTS = dict(_FlavourConfig.DEFAULTS, **_DEFAULTS)
        self.regression_type = str(cfg.get("regression_type"))
        self.subjective_score_type = str(cfg.get("subjective_score_type"))
        self.dists_weight_norm = str(cfg.get("dists_weight_norm"))
        if train_df is not None:
            X = np.asarray(train_df["DISTS"].values, dtype=np.float64)
            y = np.asarray(train_df[self.subjective_score_type].values, dtype=np.float64)
        else:
            X, y = np.asarray(dists, dtype=np.float64), np.asarray(scores, dtype=np.float64)

        if self.regression_type == "logistic":                      # model_stats.py:33-47
            from scipy.optimize import curve_fit
            mos = self.subjective_score_type == "MOS"
            sign = 1.0 if mos else -1.0

            def logistic(x, b1, b2, b3, b4):
                return (b1 - b2) / (1 + np.exp(sign * (x - b3) / np.abs(b4))) + b2

            p0 = [np.max(y) if mos else np.min(y), np.min(y) if mos else np.max(y), np.median(X), np.std(X)]
            params, _ = curve_fit(logistic, X, y, p0=p0)
            for name, v in zip(("b1", "b2", "b3", "b4"), params):
                setattr(self, name, nn.Parameter(torch.tensor([v], dtype=torch.float32)))
        else:                                                        # model_stats.py:48-60
            from sklearn.linear_model import LinearRegression
            Xr = X.reshape(-1, 1)
            if self.regression_type == "sqrt":
                Xr = np.sqrt(Xr)
            fit = LinearRegression().fit(Xr, y)
            self.dists_weight = nn.Parameter(torch.tensor([fit.coef_[0]], dtype=torch.float32))
            self.dists_bias = nn.Parameter(torch.tensor([fit.intercept_], dtype=torch.float32))

        if self.dists_weight_norm == "softmax":                      # model_stats.py:63-68
            self.dists_model = DISTSSoftmax(precision=precision, **{k: v for k, v in flavour_kwargs.items() if k == "detach_beta"})
        else:
            self.dists_model = DISTSOriginal(precision=precision, dists_weight_norm=self.dists_weight_norm, **flavour_kwargs)

    # -- regression heads (model_stats.py:71-79) ------------------------------------------------------
    def logistic(self, dists_scores):
        sign = 1.0 if self.subjective_score_type == "MOS" else -1.0
        return (self.b1 - self.b2) / (1 + torch.exp(sign * (dists_scores - self.b3) / torch.abs(self.b4))) + self.b2

    def sqrt(self, dists_scores):
        return torch.sqrt(dists_scores) * self.dists_weight + self.dists_bias

    def linear(self, dists_scores):
        return dists_scores * self.dists_weight + self.dists_bias

    def entropy_loss(self):
        """Cross-entropy of the current (normalised) channel weights against the shipped ones (model_stats.py:81-90)."""
        weights = torch.cat([self.dists_model.alpha, self.dists_model.beta], dim=1)
        if self.dists_weight_norm == "softmax":
            weights = torch.softmax(weights, dim=1)
        else:
            if self.dists_weight_norm == "relu":
                weights = torch.relu(weights)
            weights = weights / weights.sum()
        original = torch.cat([self.dists_model.original_alpha, self.dists_model.original_beta], dim=1).to(weights.device)
        return -torch.sum(original * torch.log(weights + 1e-10))

    def forward(self, dist, ref):
        """``(predicted subjective score, DISTS score)`` per frame (model_stats.py:92-102)."""
        dists_scores = self.dists_model(dist, ref)
        if self.regression_type == "logistic":
            scores = self.logistic(dists_scores)
        elif self.regression_type == "sqrt":
            scores = self.sqrt(dists_scores)
        else:
            scores = self.linear(dists_scores)
        return scores, dists_scores
