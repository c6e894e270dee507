"""Mirror of trainers/linear.py: the canned tf.estimator.LinearClassifier (trainers/linear.py:30-34) on the
"linear" feature columns = linear_model only, SUM loss, FTRL with the canned default learning rate
min(0.2, 1/sqrt(n_columns)) (TF-1.12 canned/linear.py)."""
import math

from ..engine import DeepFMEngine, default_optimizer
from .linear_deep import _CannedBase
from .ml_100k import FEATURE_DTYPES

_LEARNING_RATE = 0.2


class LinearClassifier(_CannedBase):
    def __init__(self, feature_columns, model_dir=None, config=None, max_batch=4096, device=0, feature_dtypes=FEATURE_DTYPES):
        cols = list(feature_columns)
        if not cols:
            raise ValueError("feature_columns must be defined.")
        lr = min(_LEARNING_RATE, 1.0 / math.sqrt(len(cols)))
        self.engine = DeepFMEngine(cols, (), use_linear=True, use_mf=False, use_dnn=False, loss_reduction="sum",
                                   opt_deep=default_optimizer("Ftrl", lr), opt_linear=default_optimizer("Ftrl", lr),
                                   max_batch=max_batch, device=device, feature_dtypes=feature_dtypes)
        self.model_dir = model_dir
