"""Values of the binary head's EVAL metrics (reference trainers/model_utils.py:39-54 -> tf.metrics.accuracy / auc / mean)
against closed-form small cases and an independent exact-rank restatement."""
import numpy as np

from recommender_tensorflow_b200.trainers.model_utils import get_binary_metrics, get_binary_predictions


def _logit(p):
    p = np.asarray(p, np.float64)
    return np.log(p / (1 - p))


def test_textbook_case():
    """labels 0,0,1,1 with scores .1,.4,.35,.8: ROC AUC = 3 of the 4 (negative, positive) pairs ordered = 0.75; accuracy
    at 0.5 = 3/4; average_loss = mean sigmoid cross-entropy."""
    y = np.array([0, 0, 1, 1], np.float32)
    p = np.array([0.1, 0.4, 0.35, 0.8])
    m = get_binary_metrics(y, _logit(p))
    assert abs(m["auc"] - 0.75) < 1e-5
    assert m["accuracy"] == 0.75
    want_loss = -(np.log(1 - 0.1) + np.log(1 - 0.4) + np.log(0.35) + np.log(0.8)) / 4
    assert abs(m["average_loss"] - want_loss) < 1e-9 and abs(m["loss"] - 4 * want_loss) < 1e-9
    assert m["label/mean"] == 0.5 and abs(m["prediction/mean"] - p.mean()) < 1e-12
    # PR curve through (recall, precision) = (0,1)* , (.5,1), (.5,.5)... trapezoid over tf.metrics.auc's thresholds:
    # thresholds above .8: tp=0 -> (0, 1 by the epsilon rule); (.4,.8]: tp=1, fp=0 -> (.5, 1); (.35,.4]: tp=1, fp=1 -> (.5, .5);
    # (.1,.35]: tp=2, fp=1 -> (1, 2/3); <=.1: tp=2, fp=2 -> (1, .5)
    want_pr = 0.5 * (1 + 1) / 2 + 0.0 + 0.5 * (0.5 + 2 / 3) / 2 + 0.0
    assert abs(m["auc_precision_recall"] - want_pr) < 1e-4


def test_auc_against_exact_rank_statistic():
    """tf.metrics.auc with 200 thresholds approximates the Mann-Whitney statistic: with scores on a coarse grid of
    threshold mid-points the two agree to float precision."""
    rng = np.random.default_rng(0)
    n = 5000
    y = (rng.random(n) < 0.3).astype(np.float32)
    grid = (np.arange(199) + 0.5) / 199.0                      # between consecutive thresholds k/199
    p = grid[np.clip((rng.normal(0.45 + 0.2 * y, 0.2) * 199).astype(int), 0, 198)]
    m = get_binary_metrics(y, _logit(p))
    pos, neg = p[y > 0.5], p[y < 0.5]
    exact = ((pos[:, None] > neg[None, :]).sum() + 0.5 * (pos[:, None] == neg[None, :]).sum()) / (pos.size * neg.size)
    assert abs(m["auc"] - exact) < 1e-6
    assert abs(m["accuracy"] - ((p > 0.5) == (y > 0.5)).mean()) < 1e-12


def test_perfect_and_inverted_rankings():
    y = np.array([0, 0, 0, 1, 1], np.float32)
    good = get_binary_metrics(y, np.array([-3, -2, -1, 1, 2.0]))
    bad = get_binary_metrics(y, -np.array([-3, -2, -1, 1, 2.0]))
    assert abs(good["auc"] - 1.0) < 1e-5 and good["accuracy"] == 1.0 and abs(good["auc_precision_recall"] - 1.0) < 1e-4
    assert abs(bad["auc"]) < 1e-5 and bad["accuracy"] == 0.0


def test_predictions_dict():
    z = np.array([-1.0, 0.0, 2.0], np.float32)
    p = get_binary_predictions(z)
    assert p["logits"].shape == (3, 1) and np.allclose(p["logistic"][:, 0], 1 / (1 + np.exp(-z)))
    assert np.allclose(p["probabilities"].sum(1), 1.0) and p["class_ids"][:, 0].tolist() == [0, 0, 1]
