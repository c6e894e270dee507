"""Mirror of trainers/linear_deep.py: the canned tf.estimator.DNNLinearCombinedClassifier
(trainers/linear_deep.py:32-39) = wide&deep: linear_model(linear columns) + DNN(input_layer(deep
columns)), no FM term, SUM loss, Adagrad(0.001) on the dnn side and FTRL(min(0.005, 1/sqrt(n_linear)))
on the linear side (TF-1.12 canned defaults)."""
import math
import os

import numpy as np

from ..engine import DeepFMEngine, default_optimizer
from .ml_100k import FEATURE_DTYPES, ModeKeys
from .model_utils import get_binary_metrics, get_binary_predictions

_DNN_LEARNING_RATE = 0.001
_LINEAR_LEARNING_RATE = 0.005


class _CannedBase:
    """train / evaluate / predict loops shared by the canned-estimator mirrors."""
    engine = None

    def train(self, input_fn, steps=None, max_steps=None):
        loss = None
        for feats, labels in input_fn():
            if max_steps is not None and self.engine.global_step >= max_steps:
                break
            loss = self.engine.train_step(feats, labels)
            if steps is not None:
                steps -= 1
                if steps <= 0:
                    break
        return loss

    def evaluate(self, input_fn):
        ys, zs = [], []
        for feats, labels in input_fn():
            ys.append(np.asarray(labels, dtype=np.float32).reshape(-1))
            zs.append(self.engine.predict_logits(feats))
        m = get_binary_metrics(np.concatenate(ys), np.concatenate(zs))
        m["global_step"] = self.engine.global_step
        return m

    def predict(self, input_fn):
        for item in input_fn():
            feats = item[0] if isinstance(item, tuple) else item
            preds = get_binary_predictions(self.engine.predict_logits(feats))
            for i in range(preds["logits"].shape[0]):
                yield {k: v[i] for k, v in preds.items()}


class DNNLinearCombinedClassifier(_CannedBase):
    def __init__(self, model_dir=None, linear_feature_columns=None, dnn_feature_columns=None, dnn_hidden_units=None,
                 dnn_dropout=None, config=None, max_batch=4096, device=0, feature_dtypes=FEATURE_DTYPES, tf_random_seed=None):
        linear_feature_columns = list(linear_feature_columns or [])
        dnn_feature_columns = list(dnn_feature_columns or [])
        if not linear_feature_columns and not dnn_feature_columns:
            raise ValueError("Either linear_feature_columns or dnn_feature_columns must be defined.")
        cats = [c.categorical_column for c in dnn_feature_columns] or linear_feature_columns
        if linear_feature_columns and dnn_feature_columns and [c.name for c in cats] != [c.name for c in linear_feature_columns]:
            raise NotImplementedError("linear and dnn sides must be built on the same categorical columns "
                                      "(as in trainers/ml_100k.py:37-38)")
        dims = {c.dimension for c in dnn_feature_columns}
        if len(dims) > 1:
            raise NotImplementedError("all embedding columns must share one dimension")
        lin_lr = min(_LINEAR_LEARNING_RATE, 1.0 / math.sqrt(max(len(linear_feature_columns), 1)))
        self.engine = DeepFMEngine(cats, (), embedding_size=(dims.pop() if dims else 4),
                                   hidden_units=list(dnn_hidden_units or []), use_linear=bool(linear_feature_columns),
                                   use_mf=False, use_dnn=bool(dnn_feature_columns), loss_reduction="sum",
                                   opt_deep=default_optimizer("Adagrad", _DNN_LEARNING_RATE),
                                   opt_linear=default_optimizer("Ftrl", lin_lr), max_batch=max_batch, device=device,
                                   dropout=float(dnn_dropout or 0.0),
                                   feature_dtypes=feature_dtypes)
        # variable initialisers (TF runs them on the first session.run; RunConfig.tf_random_seed)
        self.engine.init_random(int(tf_random_seed) if tf_random_seed is not None else int.from_bytes(os.urandom(4), "little"))
        self.model_dir = model_dir
