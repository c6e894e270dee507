"""Mirror of trainers/model_utils.py: optimizer lookup and the binary head's outputs."""
import numpy as np

from ..engine import default_optimizer


def layer_summary(value):
    """trainers/model_utils.py:4-6: tf.nn.zero_fraction scalar + activation histogram of a tensor, host side."""
    v = np.asarray(value, dtype=np.float32).reshape(-1)
    hist, edges = np.histogram(v, bins=30) if v.size else (np.zeros(30, int), np.zeros(31))
    return {"fraction_of_zero_values": float((v == 0).mean()) if v.size else 0.0, "activation": (hist, edges)}


def get_optimizer(optimizer_name="Adam", learning_rate=0.001):
    """trainers/model_utils.py:57-66 — Adagrad | Adam | Ftrl | RMSProp | SGD with `learning_rate` only
    (an unknown name raises KeyError, like optimizer_classes[optimizer_name] in the reference)."""
    return default_optimizer(optimizer_name, learning_rate)


def get_binary_predictions(logits):
    """trainers/model_utils.py:9-21 / binary head predictions."""
    logits = np.asarray(logits, dtype=np.float32).reshape(-1, 1)
    logistic = 1.0 / (1.0 + np.exp(-logits))
    return {"logits": logits, "logistic": logistic, "probabilities": np.concatenate([1 - logistic, logistic], 1),
            "class_ids": (logistic > 0.5).astype(np.int64), "class_id": (logistic > 0.5).astype(np.int32)}


def _auc(labels, probs, curve="ROC", num_thresholds=200):
    """tf.metrics.auc: trapezoidal over num_thresholds evenly spaced thresholds (+/- epsilon ends)."""
    eps = 1e-7
    th = np.concatenate([[0.0 - eps], (np.arange(1, num_thresholds - 1)) / (num_thresholds - 1.0), [1.0 + eps]])
    pos = labels > 0.5
    pred = probs[None, :] > th[:, None]
    tp = (pred & pos[None, :]).sum(1).astype(np.float64)
    fp = (pred & ~pos[None, :]).sum(1).astype(np.float64)
    fn = pos.sum() - tp
    tn = (~pos).sum() - fp
    if curve == "ROC":
        x = fp / (fp + tn + eps)
        y = (tp + eps) / (tp + fn + eps)
    else:
        x = (tp + eps) / (tp + fn + eps)
        y = (tp + eps) / (tp + fp + eps)
    return float(np.sum((x[:-1] - x[1:]) * (y[:-1] + y[1:]) / 2.0))


def get_binary_metrics(labels, logits):
    """Metrics the binary head reports in EVAL mode (accuracy, auc, auc_precision_recall,
    average_loss, label/mean, prediction/mean; cf. trainers/model_utils.py:39-54)."""
    y = np.asarray(labels, dtype=np.float64).reshape(-1)
    z = np.asarray(logits, dtype=np.float64).reshape(-1)
    p = 1.0 / (1.0 + np.exp(-z))
    loss = np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))
    return {"accuracy": float(((p > 0.5) == (y > 0.5)).mean()), "auc": _auc(y, p), "auc_precision_recall": _auc(y, p, "PR"),
            "average_loss": float(loss.mean()), "loss": float(loss.sum()), "label/mean": float(y.mean()),
            "prediction/mean": float(p.mean())}
