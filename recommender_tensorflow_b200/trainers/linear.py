"""Mirror of trainers/linear.py: the canned tf.estimator.LinearClassifier (trainers/linear.py:30-34) on the
"linear" feature columns = linear_model only, SUM loss, FTRL with the canned default learning rate
min(0.2, 1/sqrt(n_columns)) (TF-1.12 canned/linear.py)."""
import math

from ..engine import DeepFMEngine, default_optimizer
from .linear_deep import _CannedBase, canned_parser, canned_train_and_evaluate
from .ml_100k import FEATURE_DTYPES

_LEARNING_RATE = 0.2


class LinearClassifier(_CannedBase):
    def __init__(self, feature_columns, model_dir=None, config=None, max_batch=4096, device=0, feature_dtypes=FEATURE_DTYPES,
                 tf_random_seed=None):
        cols = list(feature_columns)
        if not cols:
            raise ValueError("feature_columns must be defined.")
        lr = min(_LEARNING_RATE, 1.0 / math.sqrt(len(cols)))
        self.engine = DeepFMEngine(cols, (), use_linear=True, use_mf=False, use_dnn=False, loss_reduction="sum",
                                   opt_deep=default_optimizer("Ftrl", lr), opt_linear=default_optimizer("Ftrl", lr),
                                   max_batch=max_batch, device=device, feature_dtypes=feature_dtypes)
        self._finish_init(model_dir, tf_random_seed)      # (linear weights start at zero like TF's; the seed is unused here)


def train_and_evaluate(args):
    return canned_train_and_evaluate(args, lambda fc, a: LinearClassifier(fc["linear"], model_dir=a.job_dir, max_batch=a.batch_size,
                                                                          tf_random_seed=a.seed))


if __name__ == "__main__":
    train_and_evaluate(canned_parser("checkpoints/linear").parse_args())
