"""Scenario metrics (reference: utils/metrics.py:10-36): sklearn on the host, 10-bin ECE."""
import numpy as np
from sklearn.metrics import average_precision_score, balanced_accuracy_score, brier_score_loss, f1_score, roc_auc_score


def compute_ece(y_true, y_prob, n_bins: int = 10) -> float:
    edges = np.linspace(0, 1, n_bins + 1)
    ece = 0.0
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (y_prob > lo) & (y_prob <= hi)
        frac = np.mean(sel)
        if frac > 0:
            acc = np.mean(y_true[sel] == (y_prob[sel] >= 0.5))
            ece += frac * np.abs(acc - np.mean(y_prob[sel]))
    return ece


def compute_metrics(y_true, y_prob, threshold: float = 0.5):
    y_pred = (y_prob >= threshold).astype(int)
    return {
        "roc_auc": roc_auc_score(y_true, y_prob),
        "pr_auc": average_precision_score(y_true, y_prob),
        "balanced_accuracy": balanced_accuracy_score(y_true, y_pred),
        "f1": f1_score(y_true, y_pred),
        "brier_score": brier_score_loss(y_true, y_prob),
        "ece": compute_ece(y_true, y_prob),
    }
